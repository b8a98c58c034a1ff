/* ihmr_b200 — C ABI of the B200-native IHMR-OPT refinement kernels (libihmr_b200.so).
 *
 * This is boundary level L2 of SURVEY.md §8(b).  The reference has exactly one native
 * precedent, the un-vendored pybind entry `sdf_cuda.sdf(phi, faces, vertices)` reached from
 * /root/reference/src/models/loss_utils.py:38,181; everything else on the path is eager
 * PyTorch.  Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C: device pointers, extents, an opaque model handle and a CUDA stream
 *     (`cudaStream_t`, passed as void* so this header needs no CUDA include);
 *   - every tensor is fp32, row-major, contiguous, resident on the model's device and owned by
 *     the caller; the library allocates only inside ihmr_model_create and never frees or keeps
 *     caller memory; scratch is a caller-provided workspace sized by the *_workspace_bytes call;
 *   - all work is enqueued on `stream`; no entry point synchronises the device;
 *   - return value 0 = ok, negative = error (IHMR_E_*); the message of the last error on the
 *     calling thread is returned by ihmr_last_error(); no C++ exception crosses the ABI;
 *   - re-entrant; a model handle is immutable after creation (except
 *     ihmr_model_update_shapedirs / ihmr_model_set_sdf_conventions) and may be shared by threads using different streams;
 *   - sm_100a only, no CPU fallback: on any other device ihmr_model_create fails.
 */
#ifndef IHMR_B200_H
#define IHMR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IHMR_NUM_VERTS 778
#define IHMR_NUM_FACES 1538
#define IHMR_NUM_JOINTS 16
#define IHMR_NUM_BETAS 10
#define IHMR_NUM_POSE_FEAT 135
#define IHMR_PARAM_DIM 122 /* [cam 3 | hand_trans 3 | pose 96 (R orient, R fingers, L orient, L fingers) | shape 20] */

#define IHMR_OK 0
#define IHMR_E_INVALID (-1)   /* bad argument */
#define IHMR_E_CUDA (-2)      /* CUDA runtime error */
#define IHMR_E_ARCH (-3)      /* device is not sm_100 */
#define IHMR_E_WORKSPACE (-4) /* workspace too small */

typedef struct ihmr_model ihmr_model_t;
typedef void* ihmr_stream_t; /* cudaStream_t */

const char* ihmr_last_error(void);
int ihmr_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long ihmr_launch_count(void);

/* ---- model constants ------------------------------------------------------------------
 * Replaces `smplx.create(path, 'mano', use_pca=False, is_rhand=..., batch_size=...)` plus
 * `.cuda()` at src/models/optimize_model.py:105-106,117 and the face buffers SDFLoss keeps
 * (src/models/loss_utils.py:34-38).  All inputs are HOST arrays, copied to `device` once:
 * v_template (778,3), shapedirs (778,3,10), posedirs (135,2334), J_regressor (16,778),
 * lbs_weights (778,16), parents (16) with parents[0] = -1, hands_mean (45),
 * faces_right / faces_left (1538,3) int32. */
int ihmr_model_create(const float* v_template, const float* shapedirs, const float* posedirs,
                      const float* J_regressor, const float* lbs_weights, const int32_t* parents,
                      const float* hands_mean, const int32_t* faces_right, const int32_t* faces_left,
                      int device, ihmr_model_t** out);
void ihmr_model_destroy(ihmr_model_t* model);
/* The reference mutates `.shapedirs` of a loaded model in place (optimize_model.py:109-113);
 * the L0 layer re-uploads through this call when its tensor changed.  shapedirs: HOST (778,3,10). */
int ihmr_model_update_shapedirs(ihmr_model_t* model, const float* shapedirs, ihmr_stream_t stream);
/* Conventions of the penetration field the upstream `sdf` package (un-vendored, unpinned) may hold differently
 * (SURVEY.md §8(c) A2, A4): the box scale is (1 + scale_factor) * 0.5 * max extent (`SDFLoss.forward(...,
 * scale_factor=0.2)` at the call site src/models/loss_utils.py:181-182 uses the default) and the axis of the
 * inside/outside parity ray (0 = +x).  Defaults 0.2 and 0; applies to every later penetration call on this model.
 * Not stream ordered: call it before enqueuing work that should see the change. */
int ihmr_model_set_sdf_conventions(ihmr_model_t* model, float scale_factor, int ray_axis);

/* ---- MANO layer (a4) ------------------------------------------------------------------
 * Replaces `mano_models['right'](global_orient=, hand_pose=, betas=)` -> .vertices/.joints at
 * src/models/optimize_model.py:194-200 (smplx 0.1.28 MANO.forward -> lbs) and its autograd
 * backward.  global_orient (n,3), hand_pose (n,45), betas (n,10) -> vertices (n,778,3),
 * joints (n,16,3).  Backward recomputes the forward from the inputs (nothing is saved):
 * grad_vertices (n,778,3) and grad_joints (n,16,3) may each be NULL (treated as zero). */
size_t ihmr_mano_workspace_bytes(int n_hands);
int ihmr_mano_forward(const ihmr_model_t* model, int n_hands, const float* global_orient,
                      const float* hand_pose, const float* betas, float* vertices, float* joints,
                      void* workspace, size_t workspace_bytes, ihmr_stream_t stream);
int ihmr_mano_backward(const ihmr_model_t* model, int n_hands, const float* global_orient,
                       const float* hand_pose, const float* betas, const float* grad_vertices,
                       const float* grad_joints, float* grad_global_orient, float* grad_hand_pose,
                       float* grad_betas, void* workspace, size_t workspace_bytes,
                       ihmr_stream_t stream);

/* The blend-shape contraction of the MANO layer (`torch.matmul(pose_feature, posedirs)` plus the
 * shape blend in smplx lbs) exposed on its own, for tests and measurements:
 * ihmr_gemm_tf32x3: C[M,Nc] = A[M,K] . B[Nc,K]^T on the tcgen05 tensor cores with 3xTF32 splitting
 * (K % 32 == 0, Nc/lda/ldb/ldc % 4 == 0); ihmr_gemm_reference_fp32: C[M,N] = A[M,K] . B[K,N] on the
 * FP32 pipe, the checker the tensor-core path is tested against (not used by the product path). */
int ihmr_gemm_tf32x3(int M, int Nc, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                     ihmr_stream_t stream);
int ihmr_gemm_reference_fp32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                             int ldc, ihmr_stream_t stream);

/* ---- interpenetration loss (a10) ------------------------------------------------------
 * Replaces `SDFLoss(faces_right, faces_left, robustifier)(hand_verts, return_per_vert_loss=True,
 * return_origin_scale_loss=True)` at src/models/loss_utils.py:181-182 including the `sdf_cuda`
 * voxel kernel and the grid_sample forward/backward behind it.  hand_verts (n,2,778,3) ->
 * losses (n), per_vert (n,1556) [may be NULL], origin_scale (n,1556) metres, right-hand
 * vertices first [may be NULL], grad_hand_verts (n,2,778,3) = d losses[b] / d hand_verts[b]
 * [may be NULL].  robustifier <= 0 means none (the IHMR-OPT path, loss_utils.py:36).
 * workspace: ihmr_sdf_workspace_bytes(n_frames) of device scratch (per-frame headers, the work list of
 * (frame, direction) items, per-direction loss sums, spill area), 256-byte aligned, not shared by
 * calls that may run concurrently. */
size_t ihmr_sdf_workspace_bytes(int n_frames);
int ihmr_sdf_loss(const ihmr_model_t* model, int n_frames, const float* hand_verts, float* losses,
                  float* per_vert, float* origin_scale, float* grad_hand_verts, float robustifier,
                  void* workspace, size_t workspace_bytes, ihmr_stream_t stream);

/* EXACT (grid-free) penetration mode — NOT the reference's function, a separately named
 * alternative: per query vertex the exact distance to the other hand's mesh if the vertex is inside it (odd +x ray
 * crossing parity) else 0, in the same normalised units / metres, same loss = sum / 4, same outputs and gradient
 * convention as ihmr_sdf_loss.  It is what the reference's 32^3 field converges to for an infinitely fine grid. */
int ihmr_sdf_loss_exact(const ihmr_model_t* model, int n_frames, const float* hand_verts, float* losses,
                        float* per_vert, float* origin_scale, float* grad_hand_verts, void* workspace,
                        size_t workspace_bytes, ihmr_stream_t stream);

/* Diagnostic variant for tools/tests: same kernels, additionally fills stats (n,32) int32 (zeroed
 * by the caller): per grid hand h in {0,1}: [2h] voxels evaluated, [4+h] query vertices inside the
 * grid box, [16+h] direction finished by the prep kernel (boxes cannot meet);
 * [6] (voxel, cluster) pairs, [7] exact candidates, [8] marked voxels, [9] ray items, [10] passes,
 * [11] / [12] ray / candidate queue segments tested early because they could not take 32 more entries,
 * [19..27] SM cycles thread 0 spent per phase (mark, face boxes, parity, scan, worklist + seeds,
 * unused, candidate search + exact tests, finish, sample + outputs). */
int ihmr_sdf_stats(const ihmr_model_t* model, int n_frames, const float* hand_verts, float* losses,
                   int32_t* stats, void* workspace, size_t workspace_bytes, ihmr_stream_t stream);

/* ---- fused refinement (a1-a13) --------------------------------------------------------
 * One call per strategy stage replaces the body of OptimizeModel.optimize's stage loop
 * (src/models/optimize_model.py:393-407): fresh optimiser state (:333-347), epoch+1 iterations
 * of forward (:254-273) -> __compute_loss (:276-330) -> snapshot every save_mid_freq (:354-374)
 * -> zero_grad/backward/step (:404-406), then filter_by_losses + select_params
 * (src/utils/opt_utils.py:104-152).  Snapshot selection is done online (origin thresholds are
 * known at snapshot 0; strict '<' keeps the first minimum), which yields the same choice as
 * stacking all snapshots. */
enum { IHMR_P_CAM = 1, IHMR_P_TRANS = 2, IHMR_P_R_ORIENT = 4, IHMR_P_R_POSE = 8, IHMR_P_L_ORIENT = 16,
       IHMR_P_L_POSE = 32, IHMR_P_R_SHAPE = 64, IHMR_P_L_SHAPE = 128 };
enum { IHMR_LOSS_JOINTS_3D_P = 0, IHMR_LOSS_COLLISION = 1, IHMR_LOSS_JOINTS_2D_P = 2 };
enum { IHMR_OPT_ADAM = 0, IHMR_OPT_SGD = 1 };
/* IHMR_STAGE_GENERIC_KERNELS: run this stage on the generic kernel chain even where a specialised rewrite
 * exists (orientation-only and shape-only stages), and push every hand through the dense backward kernels even
 * when its collision gradient is identically zero (normally such hands take a fingertip-only path); the tests
 * compare the two. */
enum { IHMR_STAGE_GENERIC_KERNELS = 1 };

typedef struct {
    uint32_t update_mask;   /* OR of IHMR_P_*: stage['update_params'] */
    float lr;
    int32_t epoch;          /* the loop runs epoch + 1 iterations */
    float w_joints_2d, w_joints_3d, w_trans, w_shape_reg, w_collision, w_finger_reg;
    int32_t n_filters;      /* stage['filter_loss'], at most 4 */
    int32_t filter_loss[4]; /* IHMR_LOSS_* */
    float filter_percent[4];/* '+0' -> 0, '-10' -> -10 */
    int32_t select_loss;    /* IHMR_LOSS_* */
    uint32_t flags;         /* OR of IHMR_STAGE_* */
} ihmr_stage_t;

/* Per-batch device inputs of the loop (what set_input copies, optimize_model.py:120-168). */
typedef struct {
    const float* init_joints_2d;    /* (B,42,3) x, y, weight   — back-propagated 2-D target */
    const float* init_joints_3d;    /* (B,42,4) x, y, z, weight — back-propagated 3-D target */
    const float* init_hand_trans_j; /* (B,1,4)  x, y, z, weight */
    const float* gt_joints_3d;      /* (B,42,4) only the wrist weight [b,0,3] is used: first root alignment */
    const float* hand_type_array;   /* (B,2) collision mask: sum > 1.5 */
} ihmr_targets_t;

/* About 159 KB per frame (10.4 GB at 65536 frames): layer intermediates, vertices and their gradients,
 * optimiser state, and the per-vertex transform cache of the shape-only stages. */
size_t ihmr_opt_workspace_bytes(int n_frames);
/* Stages that update only the global orientations or only the shape coefficients run on specialised
 * kernels (rigid / affine rewrites of the same layer, equal to rounding) unless the stage carries
 * IHMR_STAGE_GENERIC_KERNELS.
 * params (B,122) is read and updated in place; bs_norm is the batch size every batch-mean loss
 * divides by (the reference's opt.batchSize), independent of how frames are sharded. */
int ihmr_opt_stage(const ihmr_model_t* model, int n_frames, int bs_norm, float* params,
                   const ihmr_targets_t* targets, const ihmr_stage_t* stage, int save_mid_freq,
                   int optimizer, void* workspace, size_t workspace_bytes, ihmr_stream_t stream);
/* Final forward + losses with the default weights (optimize_model.py:413-414) producing what
 * get_pred_result exports (:418-435): right/left verts (B,778,3), root-aligned joints (B,42,3),
 * collision_loss (B), collision_loss_origin_scale (B,1556), and joints_3d_loss_p_batch (B). */
int ihmr_opt_final(const ihmr_model_t* model, int n_frames, const float* params,
                   const ihmr_targets_t* targets, float* right_verts, float* left_verts,
                   float* joints_3d, float* collision_loss, float* collision_origin_scale,
                   float* joints_3d_loss_p, void* workspace, size_t workspace_bytes,
                   ihmr_stream_t stream);
/* One iteration's value and gradient without an optimiser step (parity probe for tests and
 * config 2/3 style measurements): losses6 (6) = [joints_2d_p, joints_3d_p, trans_p, collision,
 * shape_reg, finger_reg] already weighted and batch-averaged, grad (B,122). */
int ihmr_opt_value_and_grad(const ihmr_model_t* model, int n_frames, int bs_norm, const float* params,
                            const ihmr_targets_t* targets, const ihmr_stage_t* stage, float* losses6,
                            float* grad, void* workspace, size_t workspace_bytes,
                            ihmr_stream_t stream);

/* The stage-end selection on its own (a13): `filter_by_losses` + `select_params`
 * (src/utils/opt_utils.py:104-152) over stacked per-snapshot criteria (S,B,3) =
 * [joints_3d_loss_p, collision_loss, joints_2d_loss_p] -> index (B) of the chosen snapshot.  It runs the same
 * device routine ihmr_opt_stage applies online after every snapshot; filter / select fields of `stage` are used. */
int ihmr_select_snapshots(int n_snapshots, int n_frames, const float* criteria, const ihmr_stage_t* stage,
                          int32_t* index, ihmr_stream_t stream);

/* ---- IHMR-MLP inference ------------------------------------------
 * The test-time path of src/models/mlp_model.py:683-699: per strategy stage a residual MLP
 * (src/models/networks.py:83-105, InterHandSubNetwork) proposes new values for the stage's parameters from
 * [image feature 1024 | final_params 122], the MANO forward + criteria are evaluated (ihmr_opt_final), and
 * select_better_params (:592-637) keeps the proposal per frame only where the criteria improved.
 * ihmr_mlp_input: x (n,1152) = [img_feat (n,1024) | cam, pose, shape, hand_trans in the reference's final_params order
 *   (:432-436) taken from params (n,122) in this library's order | zeros].
 * ihmr_linear: y[:, :out_dim] = act(x (n,in_dim) . weight^T + bias), weight (ceil4(out_dim), in_dim) row-major = the
 *   nn.Linear layout with the rows padded to a multiple of 4 (zeros), in_dim % 32 == 0; relu != 0 applies ReLU.
 * ihmr_mlp_apply: params_out = params_in + residual on the listed column segments of the (n,122) matrix; residual
 *   columns are consumed in list order (the order of stage['update_params'], :462-470).
 * ihmr_select_better: per frame, new_params replace params on the stage's update groups and cur_criteria replace
 *   prev_criteria (n,3: IHMR_LOSS_* order) iff cur[f] < prev[f] * (1 + percent_f / 100) for every filter and
 *   cur[select] <= prev[select]; kept (n) [may be NULL] receives 1 / 0.
 * ihmr_opt_criteria: forward of params (n,122) and the three per-frame criteria (n,3) = [joints_3d_loss_p * w_joints_3d,
 *   collision_loss, joints_2d_loss_p * w_joints_2d] (what compute_loss leaves in the *_batch attributes select_better_params
 *   reads, :525-538,578-582); workspace as for ihmr_opt_stage. */
int ihmr_opt_criteria(const ihmr_model_t* model, int n_frames, const float* params, const ihmr_targets_t* targets,
                      float w_joints_2d, float w_joints_3d, float* criteria, void* workspace, size_t workspace_bytes,
                      ihmr_stream_t stream);
int ihmr_mlp_input(int n, const float* img_feat, const float* params, float* x, ihmr_stream_t stream);
int ihmr_linear(int n, int in_dim, int out_dim, const float* x, int ldx, const float* weight, const float* bias, int relu,
                float* y, int ldy, ihmr_stream_t stream);
int ihmr_mlp_apply(int n, const float* residual, int ldr, int n_segments, const int32_t* seg_col, const int32_t* seg_len,
                   const float* params_in, float* params_out, ihmr_stream_t stream);
int ihmr_select_better(int n, const float* cur_criteria, float* prev_criteria, const ihmr_stage_t* stage,
                       const float* new_params, float* params, int32_t* kept, ihmr_stream_t stream);

/* ---- evaluator metrics on the device ------------------------------
 * Replaces the host-side per-frame metric code the reference runs after get_pred_result:
 * mu.get_single_joints_error and mu.get_single_pa_inter_joints_error(use_rot=False)
 * (src/utils/metric_utils.py:23-38,107-143, called at src/utils/evaluator.py:74-86) and the
 * collision mean / max of src/utils/evaluator.py:163-181.  pred_joints_3d (n,42,3), gt_joints_3d
 * (n,42,4: xyz + validity), collision_origin_scale (n,1556), scale (n) or NULL (= 1) ->
 * out (n,6) = [sum of joint errors, count, sum of no-rotation Procrustes errors, count,
 * mean collision, max collision] in the units of the inputs (metres). */
int ihmr_eval_metrics(int n_frames, const float* pred_joints_3d, const float* gt_joints_3d,
                      const float* collision_origin_scale, const float* scale, float* out,
                      ihmr_stream_t stream);

/* Measurement aid (the one entry point that synchronises `stream`): runs ONE iteration of the
 * stage (forward, losses, backward, a zero-length optimiser step) with a CUDA event after each
 * kernel class and returns the 9 device times in milliseconds:
 * [pose_prep, blend_fwd, skin_fwd, sdf, frame_loss, skin_bwd, blend_bwd, pose_bwd, step]. */
int ihmr_opt_profile_iteration(const ihmr_model_t* model, int n_frames, int bs_norm, float* params,
                               const ihmr_targets_t* targets, const ihmr_stage_t* stage,
                               float* ms_per_kernel, void* workspace, size_t workspace_bytes,
                               ihmr_stream_t stream);

/* Measurement aid (synchronises `stream`): FP32 FMA throughput of the model's device in TFLOP/s from an
 * FFMA-only micro-kernel — the denominator for the FP32-pipe fraction SURVEY.md §8(d) asks for next to GB/s
 * (MEASURED_PEAKS.json has no FP32 figure).  tflops: HOST float; scratch: >= 4 bytes of device memory. */
int ihmr_measure_fp32_peak(const ihmr_model_t* model, float* tflops, void* scratch, ihmr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IHMR_B200_H */
