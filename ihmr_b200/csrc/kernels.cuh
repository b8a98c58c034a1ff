// Host-side launchers shared between the C-ABI entry points (abi.cu) and the fused loop (opt.cu).
#pragma once
#include "common.cuh"

namespace ihmr {

// Where a hand's (orient, pose, betas) come from.
//   plain : three contiguous arrays (n,3) (n,45) (n,10)                      [MANO layer op]
//   fused : rows of the (B,122) parameter matrix; hand h = 2*frame + side, the left hand
//           (side 1) is mirrored: y,z of every axis-angle negated
//           (src/models/optimize_model.py:180-188)                           [refinement loop]
struct HandSrc {
    const float* orient = nullptr;
    const float* pose = nullptr;
    const float* betas = nullptr;
    const float* params = nullptr;
};
struct HandGrad {
    float* orient = nullptr;
    float* pose = nullptr;
    float* betas = nullptr;
    float* params_grad = nullptr;  // (B,122); pose/orient/shape slots are OVERWRITTEN
};

// Workspace carve-up for n hands (all fp32):
struct ManoWs {
    float* X;      // (n, KP)      blend coefficients [pose feature | betas | 0]
    float* A;      // (n, 16, 12)  skinning transforms, 3x4 row-major
    float* off;    // (n, LDN)     blend offsets  X @ D
    float* gposed; // (n, LDN)     d loss / d v_posed
    float* dA;     // (n, 16, 12)
    float* dX;     // (n, KP)
    float* joints; // (n, 16, 3)   scratch joints when the caller does not want them
};
size_t mano_ws_bytes(int n);
ManoWs mano_ws_carve(void* base, int n);

int launch_pose_prep(const ihmr_model* m, int n, HandSrc src, float* X, float* A, float* joints,
                     cudaStream_t st);
int launch_blend_fwd(const ihmr_model* m, int n, const float* X, float* off, cudaStream_t st);
int launch_skin_fwd(const ihmr_model* m, int n, const float* off, const float* A, float* verts,
                    cudaStream_t st, float* bbox = nullptr);
// Hands whose vertex gradient is identically zero (flagged by the penetration kernels) carry only their five
// fingertip gradients: the backward kernels take a short path for them and the blend contraction skips them.
struct SparseGrad {
    const uint8_t* gzero = nullptr;    // (n) 1 = gverts of this hand are zero and unwritten
    int* dense_list = nullptr;         // (n) hands with gzero == 0, any order; filled by the loss kernel
    int* dense_count = nullptr;        // (1)
};
// gverts (n,778,3) may be null; gtips (n,5,3) may be null (extra gradient on the 5 fingertip vertices).
// With sp.gzero: flagged hands get dA and dX (all KP entries) from the fingertips alone; their gposed rows are not written.
int launch_skin_bwd(const ihmr_model* m, int n, const float* off, const float* A, const float* gverts,
                    const float* gtips, float* gposed, float* dA, cudaStream_t st, SparseGrad sp = SparseGrad(),
                    float* dX = nullptr);
// With sp.dense_list: only the listed rows of gposed are contracted (and only their dX rows written).
// scratch: (n, 2336) floats free at this point (the blend offsets `off`, dead after the skinning backward): holds the
// K-split partial products
int launch_blend_bwd(const ihmr_model* m, int n, const float* gposed, float* dX, cudaStream_t st, SparseGrad sp = SparseGrad(),
                     float* scratch = nullptr);
// orientation-only stages (fused layout only): cache L = R0^T (x - J0), then x = R0 L + J0 and its backward
int launch_rigid_prep(int n, HandSrc src, float* verts, float* joints, float* Lv, float* Lj, cudaStream_t st);
int launch_rigid_fwd(int n, HandSrc src, float* verts, float* joints, float* Lv, float* Lj, cudaStream_t st,
                     float* bbox = nullptr);
int launch_rigid_bwd(int n, HandSrc src, const float* gverts, const float* gtips, const float* gjoints, const float* Lv,
                     const float* Lj, float* params_grad, cudaStream_t st, SparseGrad sp = SparseGrad());
// shape-only stages (fused layout only): cache (T_v | T_v c_v) per vertex, then the affine forward / backward in beta
int launch_shape_prep(const ihmr_model* m, int n, HandSrc src, const float* off, const float* A, float* cache, cudaStream_t st);
int launch_shape_fwd(const ihmr_model* m, int n, HandSrc src, const float* A, const float* cache, float* verts, cudaStream_t st,
                     float* bbox = nullptr);
int launch_shape_bwd(const ihmr_model* m, int n, const float* cache, const float* gverts, const float* gtips, float* dA,
                     float* dX, cudaStream_t st, SparseGrad sp = SparseGrad());
// skinning forward with the blend T = W . A^T on tcgen05 (blend_tc.cu)
// bbox (n,6) or null: lo xyz, hi xyz of every hand's vertices (for the penetration op)
int launch_skin_fwd_tc(const ihmr_model* m, int n, const float* off, const float* A, float* verts, cudaStream_t st,
                       float* bbox = nullptr);
// C[M,Nc] = A[M,K] . B[Nc,K]^T on tcgen05 with 3xTF32 splitting (blend_tc.cu)
// rows / nrows (device, optional): row r of the product uses row rows[r] of A and of C, for r < *nrows <= M
// Bq (device, optional): B pre-split into (hi, lo) and laid out as the tensor core reads it, one contiguous block per
// (N tile, K chunk) — gemm_presplit_b() builds it on the host.  The kernel then fetches B with two bulk copies per
// chunk (TMA, no register staging) and only stages A itself.
// ksplit > 1: K is cut into ksplit contiguous slices computed by separate CTAs into `parts` (ksplit copies of C, M * ldc
// floats each) and summed in fixed order by a second kernel — for a long K with few row tiles (the blend backward); the
// split must not depend on the batch size if results are to be independent of sharding.
int launch_gemm_tf32x3(int M, int Nc, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       cudaStream_t st, const int* rows = nullptr, const int* nrows = nullptr, const float* Bq = nullptr,
                       int ksplit = 1, float* parts = nullptr);
size_t gemm_presplit_floats(int Nc, int K);
void gemm_presplit_b(const float* B, int Nc, int K, int ldb, float* out);      // host arrays
// C[M,N] = A[M,K] . B[K,N] on the FP32 pipe: reference for the tensor-core path (tests only)
int launch_sgemm_reference(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                           cudaStream_t st);
int launch_pose_bwd(const ihmr_model* m, int n, HandSrc src, const float* dA, const float* gjoints,
                    const float* dX, HandGrad out, cudaStream_t st);

// Interpenetration loss. verts (B,2,778,3). If `joints` and `params` are given the left hand is
// stored in its mirrored model frame and is mapped to the world on load:
//   v_world = diag(-1,1,1) v + shift,  shift = trans + J_R[0] - diag(-1,1,1) J_L[0]
// (optimize_model.py:208-228); gradients are returned in the frame of the stored vertices and
// their sum over the left hand (= d/d shift, world frame) goes to gshift (B,3).
struct SdfArgs {
    const float* verts = nullptr;
    const float* bbox = nullptr;       // (B,2,6) or null: lo xyz, hi xyz of the STORED vertices of every hand, as left by
                                       // the kernel that wrote them; null = k_sdf_prep scans the vertices
    const float* joints = nullptr;     // (B,2,16,3) or null
    const float* params = nullptr;     // (B,122) or null
    const float* hand_type = nullptr;  // (B,2) or null: loss/grad masked unless both hands present
    float* per_vert = nullptr;         // (B,1556) or null
    float* origin = nullptr;           // (B,1556) or null
    float* gverts = nullptr;           // (B,2,778,3) or null
    uint8_t* gzero = nullptr;          // (B,2) or null: 1 = this hand's gverts are identically zero and were NOT written
    float* gshift = nullptr;           // (B,3) or null
    float grad_scale = 1.0f;           // gverts/gshift = grad_scale * d losses[b]/d(.)
    float robustifier = 0.0f;
    int skip_grid_mask = 0;            // bit h: skip the direction whose grid hand is h (its loss part and
                                       // the gradients of the other hand are then NOT produced)
    int* stats = nullptr;              // (B,32) debug counters / phase cycles (zeroed by the caller), tests/tools only
    uint16_t* hints = nullptr;         // (B,2,2048) or null: nearest-face seeds carried from one call to the next on the
                                       // same frames (zeroed by the caller before the first); they change the work, never
                                       // the values
    int static_grid_mask = 0;          // bit h: hand h's stored vertices are bit-identical in every call since `hints` was
                                       // zeroed; its column parities / voxel distances are then carried in pcache / phic
    uint32_t* pcache = nullptr;        // (B, 1056) or null: parity words of 1024 columns + 32 words of "known" bits, zeroed with hints
    float* phic = nullptr;             // (B, 2048) or null: finished voxel distances beside the hint table (no init needed)
    void* ws = nullptr;                // sdf_ws_bytes(B) of scratch: frame headers, work list, loss parts, spill area
    float* losses = nullptr;           // (B) or null: mask * (part_0 + part_1) / 4 (one more tiny launch)
    float box_scale = 0.6f;            // (filled by launch_sdf from the model) scale = box_scale * max bbox extent   (A2)
    int ray_axis = 0;                  // (filled by launch_sdf from the model) world axis of the parity ray          (A4)
    bool l2_prefetch = false;          // (filled by launch_sdf) verts is 16-byte aligned: the next item's frame is prefetched into L2
};
size_t sdf_ws_bytes(int B);
size_t sdf_hint_bytes(int B);
size_t sdf_pcache_bytes(int B);
size_t sdf_phic_bytes(int B);
// (B,2): the sum of rho over the query vertices of each direction of every frame (written by every launch_sdf)
const float* sdf_ws_parts(void* ws, int B);
int launch_sdf(const ihmr_model* m, int B, const SdfArgs& a, cudaStream_t st);
// the exact (grid-free) mode: same arguments and outputs, different function (see sdf.cu); hints / caches are not used
int launch_sdf_exact(const ihmr_model* m, int B, const SdfArgs& a, cudaStream_t st);

// per-frame evaluator metrics (eval.cu): out (B,6)
int launch_eval_metrics(int B, const float* pred, const float* gt, const float* origin, const float* scale, float* out,
                        cudaStream_t st);

// IHMR-MLP inference pieces (mlp.cu)
int launch_mlp_input(int n, const float* feat, const float* params, float* x, cudaStream_t st);
int launch_linear(int n, int in_dim, int out_dim, const float* x, int ldx, const float* W, const float* bias, int relu,
                  float* y, int ldy, cudaStream_t st);
int launch_mlp_apply(int n, const float* res, int ldr, int n_seg, const int* col, const int* len, const float* pin, float* pout,
                     cudaStream_t st);
int launch_select_better(int n, const float* cur, float* prev, const ihmr_stage_t* stg, const float* pnew, float* params, int* kept,
                         cudaStream_t st);

// FP32 FMA throughput of the device in TFLOP/s (synchronises the stream; scratch: >= 4 bytes of device memory)
int measure_fp32_peak(int num_sms, float* tflops, float* scratch, cudaStream_t st);

}  // namespace ihmr
