// Shared definitions for the ihmr_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ihmr_b200.h"

namespace ihmr {

constexpr int NV = IHMR_NUM_VERTS;        // 778
constexpr int NF = IHMR_NUM_FACES;        // 1538
constexpr int NJ = IHMR_NUM_JOINTS;       // 16
constexpr int NB = IHMR_NUM_BETAS;        // 10
constexpr int NPF = IHMR_NUM_POSE_FEAT;   // 135
constexpr int NC = NV * 3;                // 2334 blend-shape columns
constexpr int LDN = 2336;                 // padded column count / row stride of (hands x 2334) buffers
constexpr int KP = 160;                   // blend rows: 135 pose features + 10 betas + 15 zero pad (5 chunks of 32)
constexpr int PD = IHMR_PARAM_DIM;        // 122

// offsets inside a (B,122) parameter row
constexpr int P_CAM = 0, P_TRANS = 3, P_POSE = 6, P_SHAPE = 102;
constexpr int P_R_ORIENT = 6, P_R_POSE = 9, P_L_ORIENT = 54, P_L_POSE = 57, P_R_SHAPE = 102, P_L_SHAPE = 112;

// NVTX range over a host-side entry point (visible in Nsight Systems / Compute timelines; a no-op without a tool attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

#ifdef __CUDACC__
// floats <-> integers with the same ordering (an involution on the bit pattern)
__device__ __forceinline__ int float_ordered(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// Bounding box of a hand's vertices, emitted by the kernels that write them (the penetration op needs it per frame;
// min / max are exact and order independent, so the box is bitwise the one a scan of the stored vertices gives).
// `slot`: six shared-memory ints (ordered lo xyz, hi xyz), initialised with box_slot_init and complete after a barrier.
struct BoxAcc {
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    __device__ __forceinline__ void add(float x, float y, float z) {
        lo[0] = fminf(lo[0], x); lo[1] = fminf(lo[1], y); lo[2] = fminf(lo[2], z);
        hi[0] = fmaxf(hi[0], x); hi[1] = fmaxf(hi[1], y); hi[2] = fmaxf(hi[2], z);
    }
    // whole warp: one REDUX per value, lane 0 merges into the slot
    __device__ __forceinline__ void commit(int* slot) const {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int l = __reduce_min_sync(0xffffffffu, float_ordered(lo[c]));
            const int h = __reduce_max_sync(0xffffffffu, float_ordered(hi[c]));
            if ((threadIdx.x & 31) == 0) { atomicMin(slot + c, l); atomicMax(slot + 3 + c, h); }
        }
    }
    // whole warp: lane 0 stores the warp's box (six ordered ints) into its own row; merge_rows combines the rows
    __device__ __forceinline__ void store_row(int* row) const {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int l = __reduce_min_sync(0xffffffffu, float_ordered(lo[c]));
            const int h = __reduce_max_sync(0xffffffffu, float_ordered(hi[c]));
            if ((threadIdx.x & 31) == 0) { row[c] = l; row[3 + c] = h; }
        }
    }
};
// entry k (0..5) of the box over `nrows` per-warp rows of six ordered ints
__device__ __forceinline__ float box_merge_rows(const int* rows, int nrows, int k) {
    int v = rows[k];
    for (int w = 1; w < nrows; ++w) v = (k < 3) ? min(v, rows[w * 6 + k]) : max(v, rows[w * 6 + k]);
    return ordered_float(v);
}
__device__ __forceinline__ void box_slot_init(int* slot, int k) { slot[k] = (k < 3) ? 0x7fffffff : (int)0x80000000; }
#endif

void set_error(const char* fmt, ...);
void count_launch();   // every kernel launch of this library passes through IHMR_LAUNCH_OK

// cudaFuncSetAttribute is per device: opt a kernel in to a large dynamic shared-memory size once per
// (kernel, device).  `done` is a per-kernel bitmask of devices; a race only repeats the same call.
template <typename K>
inline int ensure_dynamic_smem(K kernel, size_t bytes, unsigned long long& done) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed"); return IHMR_E_CUDA; }
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done & bit)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return IHMR_E_CUDA; }
        done |= bit;
    }
    return IHMR_OK;
}

#define IHMR_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ihmr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                            __FILE__, __LINE__);                                        \
            return IHMR_E_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define IHMR_LAUNCH_OK()                                                                \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            ihmr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),\
                            __FILE__, __LINE__);                                        \
            return IHMR_E_CUDA;                                                         \
        }                                                                               \
        ihmr::count_launch();                                                           \
    } while (0)

}  // namespace ihmr

// Device-resident constants of one hand model (all pointers are device memory).
struct ihmr_model {
    int device;
    int num_sms;
    float* D;        // (KP, LDN)  rows 0..134 posedirs, 135..144 shapedirs (k-major), rest 0
    float* DT;       // (LDN, KP)  transpose of D
    float* DTq;      // D^T pre-split hi/lo in the tensor-core operand layout: [10 N-tiles of 256][5 K-chunks][hi|lo][256 x 32]
    float* Dq;       // D likewise for the backward contraction: [73 K-chunks][hi|lo][160 x 32]
    float* vtemp;    // (LDN)      v_template flattened, padded with 0
    float* Jt;       // (48)       J_regressor @ v_template
    float* Js;       // (10, 48)   J_regressor @ shapedirs[:, :, k]
    float* Wt;       // (16, 778)  lbs weights, joint-major
    float* W4;       // (4, 778, 4) lbs weights, tiles of 4 joints: W4[t][v][i] = W[v][4t+i]
    float* hands_mean;  // (48) [0,0,0, hands_mean(45)]
    float* Jreg;     // (16, 778)  kept for update_shapedirs
    float* Sv;       // (8, 778, 4) shapedirs per vertex, 8 float4 planes: entry c*10+k (30 used) (shape-only stages)
    uint16_t* faces[2];  // (1538, 4) u16 per hand (right, left), 4th lane unused
    uint16_t* cl_tri[2]; // (49 clusters x 32, 4) u16: vertex ids of the faces of each spatial cluster, lane 3 = valid
    int parents[16];
    float sdf_box_scale = 0.6f;   // penetration-field conventions (ihmr_model_set_sdf_conventions): (1 + scale_factor) / 2
    int sdf_ray_axis = 0;         // and the axis of the inside/outside parity ray
};
