// MANO layer kernels: Rodrigues + kinematic chain per hand
// (lane = joint), blend-shape contraction [pose feature | betas] x [posedirs ; shapedirs],
// linear blend skinning over 778 vertices, and the analytic backward of each.
//
// Replaces smplx 0.1.28 MANO.forward -> lbs as called at
// /root/reference/src/models/optimize_model.py:194-200, and its autograd backward.

#include "kernels.cuh"

namespace ihmr {

// ------------------------------------------------------------------------------------ tree
struct Tree {
    int8_t parent[16];
    int8_t depth[16];
    int8_t nchild[16];          // children of every joint (all one level deeper), ascending joint order
    int8_t child[16][15];
    int8_t maxchild[16];        // [level]: most children any joint of depth level - 1 has
    int maxdepth;
};

static Tree make_tree(const int* parents) {
    Tree t;
    t.maxdepth = 0;
    for (int j = 0; j < NJ; ++j) {
        t.parent[j] = (int8_t)(parents[j] < 0 ? 0 : parents[j]);
        t.depth[j] = (int8_t)(parents[j] < 0 ? 0 : t.depth[parents[j]] + 1);
        if (t.depth[j] > t.maxdepth) t.maxdepth = t.depth[j];
        t.nchild[j] = 0;
        t.maxchild[j] = 0;
        for (int k = 0; k < 15; ++k) t.child[j][k] = -1;
    }
    for (int ch = 1; ch < NJ; ++ch) {
        const int p = t.parent[ch];
        t.child[p][t.nchild[p]++] = (int8_t)ch;
        if (t.nchild[p] > t.maxchild[t.depth[ch]]) t.maxchild[t.depth[ch]] = t.nchild[p];
    }
    return t;
}

// --------------------------------------------------------------------------- per-joint math
struct JointState {
    float r[3];     // axis-angle incl. hands_mean (and left-hand mirror in fused mode)
    float theta;    // ||r + 1e-8||
    float R[9];     // local rotation
    float J[3];     // rest joint
    float Jp[3];    // parent's rest joint (root: 0)
    float Rgp[9];   // parent's global rotation (root: identity)
    float Rg[9];    // global rotation
    float tg[3];    // global translation = posed joint
    float beta[NB];
};

__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            c[i * 3 + k] = a[i * 3 + 0] * b[0 * 3 + k] + a[i * 3 + 1] * b[1 * 3 + k] + a[i * 3 + 2] * b[2 * 3 + k];
}

// smplx.lbs.batch_rodrigues [UPSTREAM, M3]: angle = ||r + 1e-8||, n = r/angle,
// R = I + sin K + (1-cos) K K
__device__ __forceinline__ void rodrigues(const float* r, float& theta, float* R) {
    const float e = 1e-8f;
    float ax = r[0] + e, ay = r[1] + e, az = r[2] + e;
    theta = sqrtf(ax * ax + ay * ay + az * az);
    float nx = r[0] / theta, ny = r[1] / theta, nz = r[2] / theta;
    float s, c;
    sincosf(theta, &s, &c);
    float b = 1.0f - c;
    R[0] = 1.0f - b * (ny * ny + nz * nz);
    R[1] = -s * nz + b * nx * ny;
    R[2] = s * ny + b * nx * nz;
    R[3] = s * nz + b * nx * ny;
    R[4] = 1.0f - b * (nx * nx + nz * nz);
    R[5] = -s * nx + b * ny * nz;
    R[6] = -s * ny + b * nx * nz;
    R[7] = s * nx + b * ny * nz;
    R[8] = 1.0f - b * (nx * nx + ny * ny);
}

// d loss / d r given d loss / d R for the map above
__device__ __forceinline__ void rodrigues_bwd(const float* r, float theta, const float* dR, float* dr) {
    float nx = r[0] / theta, ny = r[1] / theta, nz = r[2] / theta;
    float s, c;
    sincosf(theta, &s, &c);
    float b = 1.0f - c;
    // K and K^2
    float K[9] = {0.f, -nz, ny, nz, 0.f, -nx, -ny, nx, 0.f};
    float K2[9] = {-(ny * ny + nz * nz), nx * ny, nx * nz, nx * ny, -(nx * nx + nz * nz), ny * nz,
                   nx * nz, ny * nz, -(nx * nx + ny * ny)};
    float da = 0.f, db = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        da += dR[i] * K[i];
        db += dR[i] * K2[i];
    }
    // dK = s dR + b (dR K^T + K^T dR)
    float dK[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
#pragma unroll
            for (int m = 0; m < 3; ++m) t += dR[i * 3 + m] * K[k * 3 + m] + K[m * 3 + i] * dR[m * 3 + k];
            dK[i * 3 + k] = s * dR[i * 3 + k] + b * t;
        }
    float dn[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
    float dth = da * c + db * s;
    float inv = 1.0f / theta;
    float dot = dn[0] * r[0] + dn[1] * r[1] + dn[2] * r[2];
    dth -= dot * inv * inv;
    const float e = 1e-8f;
    dr[0] = dn[0] * inv + dth * (r[0] + e) * inv;
    dr[1] = dn[1] * inv + dth * (r[1] + e) * inv;
    dr[2] = dn[2] * inv + dth * (r[2] + e) * inv;
}

// Loads joint j of hand h and runs Rodrigues + the kinematic chain with 16 lanes per hand.
// All 32 lanes of the warp must call this (shuffles).
template <bool FUSED>
__device__ __forceinline__ void joint_forward(const HandSrc& src, int h, int j, int lane,
                                              const float* __restrict__ hands_mean,
                                              const float* __restrict__ Jt, const float* __restrict__ Js,
                                              const Tree& tree, JointState& q) {
    int side = 0;
    if (FUSED) {
        const float* row = src.params + (size_t)(h >> 1) * PD;
        side = h & 1;
#pragma unroll
        for (int c = 0; c < 3; ++c) q.r[c] = row[P_POSE + 48 * side + 3 * j + c];
        if (side) { q.r[1] = -q.r[1]; q.r[2] = -q.r[2]; }
#pragma unroll
        for (int k = 0; k < NB; ++k) q.beta[k] = row[P_SHAPE + NB * side + k];
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            q.r[c] = (j == 0) ? src.orient[(size_t)h * 3 + c] : src.pose[(size_t)h * 45 + (j - 1) * 3 + c];
#pragma unroll
        for (int k = 0; k < NB; ++k) q.beta[k] = src.betas[(size_t)h * NB + k];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) q.r[c] += hands_mean[j * 3 + c];   // M1 (entry 0..2 is zero)
    rodrigues(q.r, q.theta, q.R);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = Jt[j * 3 + c];
#pragma unroll
        for (int k = 0; k < NB; ++k) acc += Js[k * 48 + j * 3 + c] * q.beta[k];
        q.J[c] = acc;
    }
    // chain, level by level (parents have smaller depth)
#pragma unroll
    for (int i = 0; i < 9; ++i) { q.Rg[i] = q.R[i]; q.Rgp[i] = (i % 4 == 0) ? 1.f : 0.f; }
#pragma unroll
    for (int c = 0; c < 3; ++c) { q.tg[c] = q.J[c]; q.Jp[c] = 0.f; }
    const int base = lane & 16;
    const int p = tree.parent[j];
    const int dep = tree.depth[j];
    for (int level = 1; level <= tree.maxdepth; ++level) {
        float pr[9], pt[3], pj[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) pr[i] = __shfl_sync(0xffffffffu, q.Rg[i], base + p);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            pt[c] = __shfl_sync(0xffffffffu, q.tg[c], base + p);
            pj[c] = __shfl_sync(0xffffffffu, q.J[c], base + p);
        }
        if (dep == level) {
#pragma unroll
            for (int i = 0; i < 9; ++i) q.Rgp[i] = pr[i];
            mat3_mul(pr, q.R, q.Rg);
#pragma unroll
            for (int c = 0; c < 3; ++c) q.Jp[c] = pj[c];
            float d[3] = {q.J[0] - pj[0], q.J[1] - pj[1], q.J[2] - pj[2]};
#pragma unroll
            for (int c = 0; c < 3; ++c) q.tg[c] = pr[c * 3 + 0] * d[0] + pr[c * 3 + 1] * d[1] + pr[c * 3 + 2] * d[2] + pt[c];
        }
    }
}

// ------------------------------------------------------------------------------- pose prep
template <bool FUSED>
__global__ void __launch_bounds__(128) k_pose_prep(int n, HandSrc src, const float* __restrict__ hands_mean,
                                                  const float* __restrict__ Jt, const float* __restrict__ Js,
                                                  Tree tree, float* __restrict__ X, float* __restrict__ A,
                                                  float* __restrict__ joints) {
    const int lane = threadIdx.x & 31, j = lane & 15;
    int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool active = h < n;
    if (!active) h = n - 1;
    JointState q;
    joint_forward<FUSED>(src, h, j, lane, hands_mean, Jt, Js, tree, q);
    if (!active) return;
    float* xr = X + (size_t)h * KP;
    if (j >= 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) xr[(j - 1) * 9 + i] = q.R[i] - ((i % 4 == 0) ? 1.f : 0.f);
    } else {
#pragma unroll
        for (int k = 0; k < NB; ++k) xr[NPF + k] = q.beta[k];
#pragma unroll
        for (int k = NPF + NB; k < KP; ++k) xr[k] = 0.f;
    }
    float4* a = reinterpret_cast<float4*>(A + ((size_t)h * NJ + j) * 12);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float ta = q.tg[r] - (q.Rg[r * 3 + 0] * q.J[0] + q.Rg[r * 3 + 1] * q.J[1] + q.Rg[r * 3 + 2] * q.J[2]);
        a[r] = make_float4(q.Rg[r * 3 + 0], q.Rg[r * 3 + 1], q.Rg[r * 3 + 2], ta);
    }
    if (joints) {
#pragma unroll
        for (int c = 0; c < 3; ++c) joints[((size_t)h * NJ + j) * 3 + c] = q.tg[c];
    }
}

// ---------------------------------------------------------------------------- pose backward
template <bool FUSED>
__global__ void __launch_bounds__(128) k_pose_bwd(int n, HandSrc src, const float* __restrict__ hands_mean,
                                                 const float* __restrict__ Jt, const float* __restrict__ Js,
                                                 Tree tree, const float* __restrict__ dA,
                                                 const float* __restrict__ gjoints,
                                                 const float* __restrict__ dX, HandGrad out) {
    const int lane = threadIdx.x & 31, j = lane & 15, base = lane & 16;
    int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool active = h < n;
    if (!active) h = n - 1;
    JointState q;
    joint_forward<FUSED>(src, h, j, lane, hands_mean, Jt, Js, tree, q);

    // seeds from the skinning transforms A_j = [Rg | tg - Rg J] and the posed joints
    float dRg[9], dtg[3], dJ[3];
    {
        const float4* a = reinterpret_cast<const float4*>(dA + ((size_t)h * NJ + j) * 12);
        float dta[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float4 v = a[r];
            dRg[r * 3 + 0] = v.x; dRg[r * 3 + 1] = v.y; dRg[r * 3 + 2] = v.z; dta[r] = v.w;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            dtg[r] = dta[r] + (gjoints ? gjoints[((size_t)h * NJ + j) * 3 + r] : 0.f);
#pragma unroll
            for (int c = 0; c < 3; ++c) dRg[r * 3 + c] -= dta[r] * q.J[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            dJ[c] = -(q.Rg[0 * 3 + c] * dta[0] + q.Rg[1 * 3 + c] * dta[1] + q.Rg[2 * 3 + c] * dta[2]);
    }
    // reverse chain: deepest level first; a parent gathers its children in ascending joint order
    float dR[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) dR[i] = dRg[i];   // root: Rg = R
    const int dep = tree.depth[j];
    for (int level = tree.maxdepth; level >= 1; --level) {
        float c[15];
        if (dep == level) {
            // dRg_parent += dRg R^T + dtg (J - Jp)^T ; dtg_parent += dtg ; dJ_parent -= Rgp^T dtg
            float d[3] = {q.J[0] - q.Jp[0], q.J[1] - q.Jp[1], q.J[2] - q.Jp[2]};
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    c[r * 3 + k] = dRg[r * 3 + 0] * q.R[k * 3 + 0] + dRg[r * 3 + 1] * q.R[k * 3 + 1] +
                                   dRg[r * 3 + 2] * q.R[k * 3 + 2] + dtg[r] * d[k];
            float rt[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                rt[k] = q.Rgp[0 * 3 + k] * dtg[0] + q.Rgp[1 * 3 + k] * dtg[1] + q.Rgp[2 * 3 + k] * dtg[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[9 + k] = dtg[k]; c[12 + k] = -rt[k]; dJ[k] += rt[k]; }
            // local rotation gradient: dR = Rgp^T dRg
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    dR[r * 3 + k] = q.Rgp[0 * 3 + r] * dRg[0 * 3 + k] + q.Rgp[1 * 3 + r] * dRg[1 * 3 + k] +
                                    q.Rgp[2 * 3 + r] * dRg[2 * 3 + k];
        } else {
#pragma unroll
            for (int i = 0; i < 15; ++i) c[i] = 0.f;
        }
        // (a joint's children are exactly one level deeper: each lane reads its own k-th child, ascending order)
        for (int k = 0; k < tree.maxchild[level]; ++k) {
            const int ch = (dep == level - 1) ? tree.child[j][k] : -1;
#pragma unroll
            for (int i = 0; i < 15; ++i) {
                float v = __shfl_sync(0xffffffffu, c[i], base + max(ch, 0));
                if (ch >= 0) {
                    if (i < 9) dRg[i] += v;
                    else if (i < 12) dtg[i - 9] += v;
                    else dJ[i - 12] += v;
                }
            }
        }
    }
    if (dep == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) dR[i] = dRg[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) dJ[k] += dtg[k];
    }
    // pose-feature gradient from the blend GEMM: f = vec(R_j - I), j >= 1
    const float* dx = dX ? dX + (size_t)h * KP : nullptr;
    if (dx && j >= 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) dR[i] += dx[(j - 1) * 9 + i];
    }
    float dr[3];
    rodrigues_bwd(q.r, q.theta, dR, dr);
    // d beta = Js^T dJ summed over the 16 joints (fixed butterfly order) + the blend-GEMM part
    float db[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        float v = Js[k * 48 + j * 3 + 0] * dJ[0] + Js[k * 48 + j * 3 + 1] * dJ[1] + Js[k * 48 + j * 3 + 2] * dJ[2];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        db[k] = v + (dx ? dx[NPF + k] : 0.f);
    }
    if (!active) return;
    if (FUSED) {
        const int side = h & 1;
        float* g = out.params_grad + (size_t)(h >> 1) * PD;
        if (side) { dr[1] = -dr[1]; dr[2] = -dr[2]; }
#pragma unroll
        for (int c = 0; c < 3; ++c) g[P_POSE + 48 * side + 3 * j + c] += dr[c];
        if (j == 0) {
#pragma unroll
            for (int k = 0; k < NB; ++k) g[P_SHAPE + NB * side + k] += db[k];
        }
    } else {
        if (j == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) out.orient[(size_t)h * 3 + c] = dr[c];
#pragma unroll
            for (int k = 0; k < NB; ++k) out.betas[(size_t)h * NB + k] = db[k];
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) out.pose[(size_t)h * 45 + (j - 1) * 3 + c] = dr[c];
        }
    }
}

// ---------------------------------------------------------------------------- SIMT SGEMM
// C[M,N] = A[M,K] B[K,N], row-major, K % BK == 0, N % 4 == 0, lda/ldb/ldc % 4 == 0.
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_sgemm(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
        float* __restrict__ C, int ldc) {
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int k = 0; k < TN; ++k) acc[i][k] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: BM rows x BK cols, float4 along K
        for (int i = tid; i < BM * (BK / 4); i += NT) {
            int r = i / (BK / 4), c4 = i % (BK / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k0 + c4 * 4);
            As[c4 * 4 + 0][r] = v.x; As[c4 * 4 + 1][r] = v.y; As[c4 * 4 + 2][r] = v.z; As[c4 * 4 + 3][r] = v.w;
        }
        for (int i = tid; i < BK * (BN / 4); i += NT) {
            int r = i / (BN / 4), c4 = i % (BN / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + c4 * 4 < N) v = *reinterpret_cast<const float4*>(B + (size_t)(k0 + r) * ldb + n0 + c4 * 4);
            *reinterpret_cast<float4*>(&Bs[r][c4 * 4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int k = 0; k < TN; ++k) b[k] = Bs[kk][tx + k * (BN / TN)];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int k = 0; k < TN; ++k) acc[i][k] += a[i] * b[k];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int r = m0 + ty * TM + i;
        if (r >= M) continue;
#pragma unroll
        for (int k = 0; k < TN; ++k) {
            int c = n0 + tx + k * (BN / TN);
            if (c < N) C[(size_t)r * ldc + c] = acc[i][k];
        }
    }
}

// -------------------------------------------------------------------------------- skinning
constexpr int SK_HPC = 8;         // hands per CTA

// Reduce-scatter of 48 per-lane partial sums over the 32 lanes of a warp (fixed order).
// Afterwards acc[0..2] of every lane holds the warp totals of elements seg..seg+2 where
// seg = 24*b4 + 12*b3 + 6*b2 + 3*b1 (b_i = bit i of the lane id).
__device__ __forceinline__ void warp_reduce_scatter_48(float (&acc)[48], int lane) {
#define IHMR_RS_STAGE(OFF, HALF)                                              \
    {                                                                         \
        const bool up = (lane & OFF) != 0;                                    \
        _Pragma("unroll") for (int i = 0; i < HALF; ++i) {                    \
            float send = up ? acc[i] : acc[i + HALF];                         \
            float keep = up ? acc[i + HALF] : acc[i];                         \
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);          \
        }                                                                     \
    }
    IHMR_RS_STAGE(16, 24)
    IHMR_RS_STAGE(8, 12)
    IHMR_RS_STAGE(4, 6)
    IHMR_RS_STAGE(2, 3)
#undef IHMR_RS_STAGE
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
}

constexpr int SKB_THREADS = 384;        // 12 warps = 4 joint tiles x 3 vertex slices
constexpr int SKB_KSLICES = 3;
constexpr int SKB_SLICE = 288;          // vertices per slice (9 x 32)

__global__ void __launch_bounds__(SKB_THREADS)
k_skin_bwd(int n, const float* __restrict__ off, const float* __restrict__ A, const float* __restrict__ vtemp,
           const float* __restrict__ W4, const float* __restrict__ gverts, const float* __restrict__ gtips,
           float* __restrict__ gposed, float* __restrict__ dA, const int* __restrict__ dense_list,
           const int* __restrict__ dense_count) {
    extern __shared__ float4 smem4[];
    float4* sW4 = smem4;                       // [4][778]
    float4* sG = sW4 + 4 * NV;                 // [778]  (g, 0)
    float4* sP = sG + NV;                      // [778]  (v_posed, 1)
    float4* sA = sP + NV;                      // [SK_HPC][48]
    float* sPart = reinterpret_cast<float*>(sA + SK_HPC * 48);   // [SKB_KSLICES][192]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // With a list (hands that do have a vertex gradient) the CTA takes 8 list entries; CTAs beyond the list leave.
    const int h0 = blockIdx.x * SK_HPC;
    const int ntot = dense_list ? min(n, *dense_count) : n;
    if (h0 >= ntot) return;
    const int nh = min(SK_HPC, ntot - h0);

    // Stage the skinning weights (49.8 KB, tile-major) and this CTA's joint transforms with the TMA engine:
    // 1-D bulk copies global -> shared that complete on an mbarrier, issued by one thread while the
    // others go straight to waiting (no register round trip, no per-thread address arithmetic).
    __shared__ __align__(8) uint64_t tma_bar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
    constexpr uint32_t W4_BYTES = 4 * NV * 16;
    const uint32_t a_bytes = (uint32_t)nh * 48 * 16;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(W4_BYTES + a_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((uint32_t)__cvta_generic_to_shared(sW4)), "l"(W4), "r"(W4_BYTES), "r"(bar) : "memory");
        if (!dense_list) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"((uint32_t)__cvta_generic_to_shared(sA)), "l"(A + (size_t)h0 * 192), "r"(a_bytes), "r"(bar) : "memory");
        } else {
            for (int hh = 0; hh < nh; ++hh)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"((uint32_t)__cvta_generic_to_shared(sA + hh * 48)), "l"(A + (size_t)dense_list[h0 + hh] * 192), "r"(768u), "r"(bar) : "memory");
        }
    }
    __syncthreads();                                   // the barrier word is initialised before anyone polls it
    {
        uint32_t done;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        } while (!done);
    }

    for (int hh = 0; hh < nh; ++hh) {
        const size_t h = dense_list ? (size_t)dense_list[h0 + hh] : (size_t)(h0 + hh);
        // phase A (thread = vertex): d v_posed = T^T g, stash (g,0) and (v_posed,1)
        for (int v = tid; v < NV; v += SKB_THREADS) {
            float g[3] = {0.f, 0.f, 0.f}, vp[3];
            if (gverts) {
#pragma unroll
                for (int c = 0; c < 3; ++c) g[c] = gverts[(h * NV + v) * 3 + c];
            }
            if (gtips) {
                const int tip = (v == 744) ? 0 : (v == 320) ? 1 : (v == 443) ? 2 : (v == 554) ? 3 : (v == 671) ? 4 : -1;
                if (tip >= 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) g[c] += gtips[(h * 5 + tip) * 3 + c];
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) vp[c] = vtemp[v * 3 + c] + off[h * LDN + v * 3 + c];
            float T[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) T[i] = 0.f;
#pragma unroll
            for (int jt = 0; jt < 4; ++jt) {
                const float4 w4 = sW4[jt * NV + v];
                const float wj[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int ji = 0; ji < 4; ++ji) {
                    const int j = jt * 4 + ji;
                    const float4 r0 = sA[hh * 48 + j * 3 + 0], r1 = sA[hh * 48 + j * 3 + 1], r2 = sA[hh * 48 + j * 3 + 2];
                    const float ww = wj[ji];
                    T[0] += ww * r0.x; T[1] += ww * r0.y; T[2] += ww * r0.z;
                    T[3] += ww * r1.x; T[4] += ww * r1.y; T[5] += ww * r1.z;
                    T[6] += ww * r2.x; T[7] += ww * r2.y; T[8] += ww * r2.z;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                gposed[h * LDN + v * 3 + c] = T[0 * 3 + c] * g[0] + T[1 * 3 + c] * g[1] + T[2 * 3 + c] * g[2];
            sG[v] = make_float4(g[0], g[1], g[2], 0.f);
            sP[v] = make_float4(vp[0], vp[1], vp[2], 1.f);
        }
        if (tid < LDN - NC) gposed[h * LDN + NC + tid] = 0.f;   // zero the pad columns (GEMM reads them)
        __syncthreads();
        // phase B (warp = joint tile x vertex slice): dA_j += W[v,j] g_v (x) [v_posed, 1]
        {
            const int t = warp & 3, ks = warp >> 2;
            float acc[48];
#pragma unroll
            for (int i = 0; i < 48; ++i) acc[i] = 0.f;
            const int vend = min(NV, (ks + 1) * SKB_SLICE);
            for (int vv = ks * SKB_SLICE + lane; vv < vend; vv += 32) {
                const float4 w4 = sW4[t * NV + vv], g = sG[vv], p = sP[vv];
                const float wj[4] = {w4.x, w4.y, w4.z, w4.w};
                const float gg[3] = {g.x, g.y, g.z};
                const float pp[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float wg = wj[i] * gg[r];
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[i * 12 + r * 4 + c] += wg * pp[c];
                    }
            }
            warp_reduce_scatter_48(acc, lane);
            if ((lane & 1) == 0) {
                const int seg = ((lane >> 4) & 1) * 24 + ((lane >> 3) & 1) * 12 + ((lane >> 2) & 1) * 6 + ((lane >> 1) & 1) * 3;
#pragma unroll
                for (int i = 0; i < 3; ++i) sPart[ks * 192 + t * 48 + seg + i] = acc[i];
            }
        }
        __syncthreads();
        if (tid < 192) {
            float s = sPart[tid];
#pragma unroll
            for (int ks = 1; ks < SKB_KSLICES; ++ks) s += sPart[ks * 192 + tid];
            dA[h * 192 + tid] = s;
        }
        // the next hand's phase A only touches sG/sP (all phase-B readers are past the barrier
        // above); its phase-B writes to sPart come after its own first barrier, i.e. after the
        // reads just above in program order of those threads and a barrier for the others.
    }
}

// Hands whose vertex gradient is identically zero (no penetration gradient: flagged by the sdf kernels): only the
// five fingertip vertices carry gradient.  One warp per hand: dA and the whole dX row come from them directly;
// gposed is not written (the blend contraction skips these hands).
__global__ void __launch_bounds__(256)
k_skin_bwd_tips(int n, const float* __restrict__ off, const float* __restrict__ A, const float* __restrict__ vtemp,
                const float* __restrict__ Wt, const float* __restrict__ gtips, const uint8_t* __restrict__ gzero,
                const float* __restrict__ DT, float* __restrict__ dA, float* __restrict__ dX) {
    __shared__ float sT[8][5][12];          // per warp and fingertip: g (3), [v_posed, 1] (4), d v_posed (3)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t h = (size_t)blockIdx.x * 8 + warp;
    if (h >= (size_t)n || !gzero[h]) return;
    const int tips[5] = {744, 320, 443, 554, 671};
    if (lane < 5) {
        const int v = tips[lane];
        float g[3] = {0.f, 0.f, 0.f}, vp[3], T[9];
        if (gtips) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = gtips[(h * 5 + lane) * 3 + c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) vp[c] = vtemp[v * 3 + c] + off[h * LDN + v * 3 + c];
#pragma unroll
        for (int i = 0; i < 9; ++i) T[i] = 0.f;
        for (int j = 0; j < NJ; ++j) {
            const float ww = Wt[j * NV + v];
            const float* a = A + h * 192 + j * 12;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) T[r * 3 + c] += ww * a[r * 4 + c];
        }
        float* o = sT[warp][lane];
        o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
        o[3] = vp[0]; o[4] = vp[1]; o[5] = vp[2]; o[6] = 1.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[7 + c] = T[0 * 3 + c] * g[0] + T[1 * 3 + c] * g[1] + T[2 * 3 + c] * g[2];
    }
    __syncwarp();
    for (int x = lane; x < 192; x += 32) {       // dA[j][r][c] = sum_tips W[v,j] g[r] [v_posed,1][c]
        const int j = x / 12, r = (x % 12) >> 2, c = x & 3;
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 5; ++t) acc += Wt[j * NV + tips[t]] * sT[warp][t][r] * sT[warp][t][3 + c];
        dA[h * 192 + x] = acc;
    }
    for (int k = lane; k < KP; k += 32) {         // dX[k] = sum_{tip,c} d v_posed[tip][c] D[k][3 tip_v + c]
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 5; ++t)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc += sT[warp][t][7 + c] * DT[(size_t)(tips[t] * 3 + c) * KP + k];
        dX[h * KP + k] = acc;
    }
}

constexpr size_t SKIN_BWD_SMEM = sizeof(float4) * (4 * NV + 2 * NV + SK_HPC * 48) + sizeof(float) * SKB_KSLICES * 192;

// ------------------------------------------------------------------ rigid (orientation-only) path
// When a stage updates nothing but the two global orientations (opt_default stage 1,
// /root/reference/src/strategies/opt_default.py:23-40) every vertex and joint of a hand is an affine
// function of its root rotation:  x = R0 L + J0  with  L = R0_init^T (x_init - J0)  fixed for the stage
// (G_j = G_0 H_j with H_j independent of the root rotation; the wrist J0 does not depend on it either).
// The stage's first iteration runs the generic layer and k_rigid_prep caches L; afterwards the forward
// is one 3x3 transform per point and the backward one 3x3 reduction per hand.
template <bool FUSED>
__device__ __forceinline__ void root_rotation(const HandSrc& src, int h, float* r, float& theta, float* R) {
    if (FUSED) {
        const float* row = src.params + (size_t)(h >> 1) * PD + P_POSE + 48 * (h & 1);
        r[0] = row[0]; r[1] = row[1]; r[2] = row[2];
        if (h & 1) { r[1] = -r[1]; r[2] = -r[2]; }
    } else {
        r[0] = src.orient[(size_t)h * 3]; r[1] = src.orient[(size_t)h * 3 + 1]; r[2] = src.orient[(size_t)h * 3 + 2];
    }
    rodrigues(r, theta, R);
}

constexpr int RG_THREADS = 256;

// L = R0^T (x - J0) from the generic forward's verts/joints (first iteration of an orientation-only stage)
__global__ void __launch_bounds__(RG_THREADS)
k_rigid_prep(int n, HandSrc src, const float* __restrict__ verts, const float* __restrict__ joints, float* __restrict__ Lv,
             float* __restrict__ Lj) {
    const int h = blockIdx.x, tid = threadIdx.x;
    float r[3], theta, R[9];
    root_rotation<true>(src, h, r, theta, R);
    const float J0[3] = {joints[(size_t)h * 48], joints[(size_t)h * 48 + 1], joints[(size_t)h * 48 + 2]};
    for (int i = tid; i < NV + NJ; i += RG_THREADS) {
        const float* x = i < NV ? verts + ((size_t)h * NV + i) * 3 : joints + ((size_t)h * NJ + (i - NV)) * 3;
        float* l = i < NV ? Lv + (size_t)h * LDN + i * 3 : Lj + (size_t)h * 192 + (i - NV) * 3;
        const float d[3] = {x[0] - J0[0], x[1] - J0[1], x[2] - J0[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) l[c] = R[0 * 3 + c] * d[0] + R[1 * 3 + c] * d[1] + R[2 * 3 + c] * d[2];
    }
}

// x = R0 L + J0, one warp per hand (no block barrier: the hand's bounding box for the penetration op is a warp
// reduction).  The wrist itself (joint 0) never moves.
constexpr int RGF_WARPS = 8;
__global__ void __launch_bounds__(RGF_WARPS * 32)
k_rigid_fwd(int n, HandSrc src, float* __restrict__ verts, float* __restrict__ joints, const float* __restrict__ Lv,
            const float* __restrict__ Lj, float* __restrict__ bbox) {
    const int lane = threadIdx.x & 31, h = blockIdx.x * RGF_WARPS + (threadIdx.x >> 5);
    if (h >= n) return;
    float r[3], theta, R[9];
    root_rotation<true>(src, h, r, theta, R);
    const float J0[3] = {joints[(size_t)h * 48], joints[(size_t)h * 48 + 1], joints[(size_t)h * 48 + 2]};
    BoxAcc box;
    const float* lv = Lv + (size_t)h * LDN;
    float* xv = verts + (size_t)h * NV * 3;
#pragma unroll 5
    for (int i = lane; i < NV; i += 32) {
        const float a[3] = {lv[i * 3], lv[i * 3 + 1], lv[i * 3 + 2]};
        float y[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { y[c] = R[c * 3 + 0] * a[0] + R[c * 3 + 1] * a[1] + R[c * 3 + 2] * a[2] + J0[c]; xv[i * 3 + c] = y[c]; }
        box.add(y[0], y[1], y[2]);
    }
    if (lane >= 1 && lane < NJ) {
        const float* l = Lj + (size_t)h * 192 + lane * 3;
        float* x = joints + ((size_t)h * NJ + lane) * 3;
        const float a[3] = {l[0], l[1], l[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = R[c * 3 + 0] * a[0] + R[c * 3 + 1] * a[1] + R[c * 3 + 2] * a[2] + J0[c];
    }
    if (bbox) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int l = __reduce_min_sync(0xffffffffu, float_ordered(box.lo[c]));
            const int hgh = __reduce_max_sync(0xffffffffu, float_ordered(box.hi[c]));
            if (lane == c) bbox[(size_t)h * 6 + c] = ordered_float(l);
            if (lane == 3 + c) bbox[(size_t)h * 6 + 3 + c] = ordered_float(hgh);
        }
    }
}

// d loss / d orient = rodrigues_bwd( sum_points g (x) L ), added to the orient slots of the gradient rows.
// One warp per hand: a hand without collision gradient (two thirds of them) has only its five fingertip vertices and
// the joints to visit, a single trip; the 3x3 sum is a fixed-order lane sum + butterfly, no block barrier.
__global__ void __launch_bounds__(RGF_WARPS * 32)
k_rigid_bwd(int n, HandSrc src, const float* __restrict__ gverts, const float* __restrict__ gtips,
            const float* __restrict__ gjoints, const float* __restrict__ Lv, const float* __restrict__ Lj,
            float* __restrict__ params_grad, const uint8_t* __restrict__ gzero) {
    const int lane = threadIdx.x & 31, h = blockIdx.x * RGF_WARPS + (threadIdx.x >> 5);
    if (h >= n) return;
    float M[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = 0.f;
    auto accumulate = [&](const float* g, const float* l) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[a * 3 + c] += g[a] * l[c];
    };
    const bool sparse = gzero && gzero[h];
    const float* lv = Lv + (size_t)h * LDN;
    if (!sparse) {
        const float* gv = gverts + (size_t)h * NV * 3;
#pragma unroll 5
        for (int i = lane; i < NV; i += 32) {
            const float g[3] = {gv[i * 3], gv[i * 3 + 1], gv[i * 3 + 2]};
            const float l[3] = {lv[i * 3], lv[i * 3 + 1], lv[i * 3 + 2]};
            accumulate(g, l);
        }
    }
    if (lane < 5) {                     // the fingertip joints' gradients arrive on their vertices
        const int tips[5] = {744, 320, 443, 554, 671};
        const int i = tips[lane];
        const float* gp = gtips + ((size_t)h * 5 + lane) * 3;
        const float g[3] = {gp[0], gp[1], gp[2]};
        const float l[3] = {lv[i * 3], lv[i * 3 + 1], lv[i * 3 + 2]};
        accumulate(g, l);
    } else if (lane >= 8 && lane < 8 + NJ) {
        const int j = lane - 8;
        const float* gp = gjoints + ((size_t)h * NJ + j) * 3;
        const float* lp = Lj + (size_t)h * 192 + j * 3;
        const float g[3] = {gp[0], gp[1], gp[2]};
        const float l[3] = {lp[0], lp[1], lp[2]};
        accumulate(g, l);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) M[i] += __shfl_xor_sync(0xffffffffu, M[i], o);
    }
    if (lane == 0) {
        float r[3], theta, R[9], dr[3];
        root_rotation<true>(src, h, r, theta, R);
        rodrigues_bwd(r, theta, M, dr);
        if (h & 1) { dr[1] = -dr[1]; dr[2] = -dr[2]; }
        float* gr = params_grad + (size_t)(h >> 1) * PD + P_POSE + 48 * (h & 1);
        gr[0] += dr[0]; gr[1] += dr[1]; gr[2] += dr[2];
    }
}

int launch_rigid_prep(int n, HandSrc src, float* verts, float* joints, float* Lv, float* Lj, cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    k_rigid_prep<<<n, RG_THREADS, 0, st>>>(n, src, verts, joints, Lv, Lj);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_rigid_fwd(int n, HandSrc src, float* verts, float* joints, float* Lv, float* Lj, cudaStream_t st, float* bbox) {
    if (n <= 0) return IHMR_OK;
    k_rigid_fwd<<<(n + RGF_WARPS - 1) / RGF_WARPS, RGF_WARPS * 32, 0, st>>>(n, src, verts, joints, Lv, Lj, bbox);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_rigid_bwd(int n, HandSrc src, const float* gverts, const float* gtips, const float* gjoints, const float* Lv,
                     const float* Lj, float* params_grad, cudaStream_t st, SparseGrad sp) {
    if (n <= 0) return IHMR_OK;
    k_rigid_bwd<<<(n + RGF_WARPS - 1) / RGF_WARPS, RGF_WARPS * 32, 0, st>>>(n, src, gverts, gtips, gjoints, Lv, Lj, params_grad, sp.gzero);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

// ------------------------------------------------------------------ shape-only path
// When a stage updates nothing but the shape coefficients (opt_default stage 3,
// /root/reference/src/strategies/opt_default.py:61-78) the rotations of the kinematic chain are fixed and the
// layer is affine in beta:
//     verts_v = T_v (c_v + S_v beta) + sum_j W[v,j] a_j(beta),
// T_v = sum_j W[v,j] Rg_j (3x3) and c_v = v_template + pose offsets fixed for the stage, a_j = the
// translation column of the joint transform A_j, which k_pose_prep recomputes every iteration together
// with the joints.  The stage's first iteration runs the generic layer and k_shape_prep caches
// (T_v | T_v c_v) as three float4 per vertex (one plane per row, lanes read consecutive words); afterwards the forward is 87 FMA per vertex and the
// backward produces exactly what k_pose_bwd needs for beta: the translation column of dA
// (sum_v W[v,j] g_v) and the beta entries of dX (sum_v S_v^T T_v^T g_v).  Both kernels stream the cache
// (37 KB per hand) and are HBM bound instead of FP32 bound.
constexpr int SH_THREADS = 416;   // 13 warps, vertices tid and tid + 416
#ifndef SH_HPC_N
#define SH_HPC_N 8
#endif
#ifndef SH_MINB_F
#define SH_MINB_F 2
#endif
#ifndef SH_MINB_B
#define SH_MINB_B 2
#endif
constexpr int SH_HPC = SH_HPC_N;  // hands per CTA
constexpr int SH_ITEMS_F = 2;      // vertices per thread in k_shape_fwd

// Sv is stored as 8 planes of float4 per vertex (plane i = entries 4i..4i+3 of the vertex's 30 values
// [c * 10 + k]), so that the lanes of a warp read consecutive 16-byte words
__device__ __forceinline__ void load_shape_row(const float* __restrict__ Sv, int v, float (&sv)[32]) {
    const float4* p = reinterpret_cast<const float4*>(Sv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = p[i * NV + v];
        sv[i * 4 + 0] = t.x; sv[i * 4 + 1] = t.y; sv[i * 4 + 2] = t.z; sv[i * 4 + 3] = t.w;
    }
}

__device__ __forceinline__ void stage_hands(int n, int h0, int nh, const HandSrc& src, const float* __restrict__ A,
                                            const float* __restrict__ W4, float4* sW4, float4 (*sA)[48], float (*sBeta)[12]) {
    const int tid = threadIdx.x;
    if (sW4) {
        const float4* w = reinterpret_cast<const float4*>(W4);
        for (int i = tid; i < 4 * NV; i += SH_THREADS) sW4[i] = w[i];
    }
    const float4* A4 = reinterpret_cast<const float4*>(A) + (size_t)h0 * 48;
    for (int i = tid; i < nh * 48; i += SH_THREADS) sA[i / 48][i % 48] = A4[i];
    for (int i = tid; i < nh * NB; i += SH_THREADS) {
        const int h = h0 + i / NB, k = i % NB;
        sBeta[i / NB][k] = src.params[(size_t)(h >> 1) * PD + P_SHAPE + NB * (h & 1) + k];
    }
}

__global__ void __launch_bounds__(SH_THREADS)
k_shape_prep(int n, HandSrc src, const float* __restrict__ off, const float* __restrict__ A,
             const float* __restrict__ vtemp, const float* __restrict__ W4, const float* __restrict__ Sv,
             float4* __restrict__ cache) {
    extern __shared__ float4 smem4[];
    float4* sW4 = smem4;                                                   // [4][778]
    float4 (*sA)[48] = reinterpret_cast<float4 (*)[48]>(sW4 + 4 * NV);       // [SH_HPC][48]
    float (*sBeta)[12] = reinterpret_cast<float (*)[12]>(sA + SH_HPC);       // [SH_HPC][12]
    const int tid = threadIdx.x;
    const int h0 = blockIdx.x * SH_HPC, nh = min(SH_HPC, n - h0);
    stage_hands(n, h0, nh, src, A, W4, sW4, sA, sBeta);
    __syncthreads();
    for (int v = tid; v < NV; v += SH_THREADS) {
        float sv[32];
        load_shape_row(Sv, v, sv);
        float w[NJ];
#pragma unroll
        for (int t = 0; t < 4; ++t) { const float4 q = sW4[t * NV + v]; w[t * 4] = q.x; w[t * 4 + 1] = q.y; w[t * 4 + 2] = q.z; w[t * 4 + 3] = q.w; }
        const float vt[3] = {vtemp[v * 3], vtemp[v * 3 + 1], vtemp[v * 3 + 2]};
        for (int hh = 0; hh < nh; ++hh) {
            const size_t h = h0 + hh;
            float T[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) T[i] = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 r0 = sA[hh][j * 3], r1 = sA[hh][j * 3 + 1], r2 = sA[hh][j * 3 + 2];
                T[0] += w[j] * r0.x; T[1] += w[j] * r0.y; T[2] += w[j] * r0.z;
                T[3] += w[j] * r1.x; T[4] += w[j] * r1.y; T[5] += w[j] * r1.z;
                T[6] += w[j] * r2.x; T[7] += w[j] * r2.y; T[8] += w[j] * r2.z;
            }
            float c[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float sb = 0.f;
#pragma unroll
                for (int k = 0; k < NB; ++k) sb += sv[a * NB + k] * sBeta[hh][k];
                c[a] = vt[a] + (off[h * LDN + v * 3 + a] - sb);        // v_template + pose offsets only
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
                cache[(h * 3 + r) * NV + v] = make_float4(T[r * 3], T[r * 3 + 1], T[r * 3 + 2],
                                                          T[r * 3] * c[0] + T[r * 3 + 1] * c[1] + T[r * 3 + 2] * c[2]);
        }
    }
}

// The cache of a hand is one contiguous block (3 planes x 778 float4 = 37,344 B): it is streamed through a
// ring of shared-memory stages by the TMA engine (one bulk copy per hand, mbarrier completion), three hands
// ahead of the arithmetic, so the kernel runs at the speed of its loads with one CTA per SM.  Every thread
// owns two vertices and keeps their skinning weights and shape directions in registers.
constexpr int SHF_STAGES = 3;
constexpr int SHF_HPC = 16;                                   // hands per CTA
constexpr uint32_t SHAPE_CACHE_BYTES = 3 * NV * sizeof(float4);
static_assert(SHAPE_CACHE_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

__device__ __forceinline__ void shape_issue(const float4* cache, size_t h, float4* stage, unsigned long long* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(SHAPE_CACHE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(stage)), "l"(cache + h * 3 * NV), "r"(SHAPE_CACHE_BYTES), "r"(b) : "memory");
}

__device__ __forceinline__ void shape_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(b), "r"(parity) : "memory");
    } while (!done);
}

__global__ void __launch_bounds__(SH_THREADS, 1)
k_shape_fwd(int n, HandSrc src, const float* __restrict__ A, const float* __restrict__ W4,
            const float* __restrict__ Sv, const float4* __restrict__ cache, float* __restrict__ verts, float* __restrict__ bbox) {
    extern __shared__ float4 smem4[];
    __shared__ int s_box[2][SH_THREADS / 32][6];                  // per-warp boxes of the current hand, two buffers
    float4* ring = smem4;                                                             // [SHF_STAGES][3 * NV]
    float4 (*sT)[NJ] = reinterpret_cast<float4 (*)[NJ]>(ring + SHF_STAGES * 3 * NV);  // [SHF_HPC][16] translation columns
    float (*sBeta)[12] = reinterpret_cast<float (*)[12]>(sT + SHF_HPC);               // [SHF_HPC][12]
    __shared__ __align__(8) unsigned long long bars[SHF_STAGES];
    const int tid = threadIdx.x;
    const int h0 = blockIdx.x * SHF_HPC, nh = min(SHF_HPC, n - h0);
    if (tid == 0) {
        for (int k = 0; k < SHF_STAGES; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bars[k])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int k = 0; k < SHF_STAGES && k < nh; ++k) shape_issue(cache, (size_t)h0 + k, ring + k * 3 * NV, &bars[k]);
    }
    // per-hand scalars while the first copies fly
    for (int i = tid; i < nh * NJ; i += SH_THREADS) {
        const int hh = i / NJ, j = i % NJ;
        const float* a = A + ((size_t)(h0 + hh) * NJ + j) * 12;
        sT[hh][j] = make_float4(a[3], a[7], a[11], 0.f);
    }
    for (int i = tid; i < nh * NB; i += SH_THREADS) {
        const int h = h0 + i / NB, k = i % NB;
        sBeta[i / NB][k] = src.params[(size_t)(h >> 1) * PD + P_SHAPE + NB * (h & 1) + k];
    }
    // this thread's two vertices: weights and shape directions in registers
    float w[SH_ITEMS_F][NJ], sv[SH_ITEMS_F][32];
    int vv[SH_ITEMS_F];
#pragma unroll
    for (int sl = 0; sl < SH_ITEMS_F; ++sl) {
        vv[sl] = min(tid + sl * SH_THREADS, NV - 1);          // out-of-range slots compute a duplicate and do not store
        load_shape_row(Sv, vv[sl], sv[sl]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float4 q = reinterpret_cast<const float4*>(W4)[t * NV + vv[sl]];
            w[sl][t * 4] = q.x; w[sl][t * 4 + 1] = q.y; w[sl][t * 4 + 2] = q.z; w[sl][t * 4 + 3] = q.w;
        }
    }
    __syncthreads();                                          // barriers initialised, sT / sBeta written
    for (int hh = 0; hh < nh; ++hh) {
        const size_t h = h0 + hh;
        const int st = hh % SHF_STAGES;
        shape_wait(&bars[st], (hh / SHF_STAGES) & 1);
        const float4* rows = ring + st * 3 * NV;
        const float4* b4 = reinterpret_cast<const float4*>(sBeta[hh]);
        const float4 b0 = b4[0], b1 = b4[1], b2 = b4[2];
        const float beta[NB] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y};
        float t[SH_ITEMS_F][3];
        BoxAcc box;
#pragma unroll
        for (int sl = 0; sl < SH_ITEMS_F; ++sl) { t[sl][0] = 0.f; t[sl][1] = 0.f; t[sl][2] = 0.f; }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 a = sT[hh][j];
#pragma unroll
            for (int sl = 0; sl < SH_ITEMS_F; ++sl) { t[sl][0] += w[sl][j] * a.x; t[sl][1] += w[sl][j] * a.y; t[sl][2] += w[sl][j] * a.z; }
        }
#pragma unroll
        for (int sl = 0; sl < SH_ITEMS_F; ++sl) {
            float u[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float sb = 0.f;
#pragma unroll
                for (int k = 0; k < NB; ++k) sb += sv[sl][a * NB + k] * beta[k];
                u[a] = sb;
            }
            if (tid + sl * SH_THREADS < NV) {
                float x[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4 row = rows[r * NV + vv[sl]];
                    x[r] = (row.w + t[sl][r]) + (row.x * u[0] + row.y * u[1] + row.z * u[2]);
                    verts[(h * NV + vv[sl]) * 3 + r] = x[r];
                }
                box.add(x[0], x[1], x[2]);
            }
        }
        if (bbox) box.store_row(s_box[hh & 1][tid >> 5]);
        __syncthreads();                                      // every thread is done with this stage (and with the hand's box)
        if (tid == 0 && hh + SHF_STAGES < nh) shape_issue(cache, h + SHF_STAGES, ring + st * 3 * NV, &bars[st]);
        // (the other buffer takes the next hand's rows; this one is rewritten after the next barrier)
        if (bbox && tid >= 32 && tid < 38) bbox[h * 6 + (tid - 32)] = box_merge_rows(&s_box[hh & 1][0][0], SH_THREADS / 32, tid - 32);
    }
}

// reduce-scatter of 16 per-lane partial sums (fixed order): afterwards acc[0] of every lane holds the warp
// total of element 8*b4 + 4*b3 + 2*b2 + b1 (b_i = bit i of the lane id)
__device__ __forceinline__ void warp_reduce_scatter_16(float (&acc)[16], int lane) {
#define IHMR_RS_STAGE(OFF, HALF)                                              \
    {                                                                         \
        const bool up = (lane & OFF) != 0;                                    \
        _Pragma("unroll") for (int i = 0; i < HALF; ++i) {                    \
            float send = up ? acc[i] : acc[i + HALF];                         \
            float keep = up ? acc[i + HALF] : acc[i];                         \
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);          \
        }                                                                     \
    }
    IHMR_RS_STAGE(16, 8)
    IHMR_RS_STAGE(8, 4)
    IHMR_RS_STAGE(4, 2)
    IHMR_RS_STAGE(2, 1)
#undef IHMR_RS_STAGE
    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
}

constexpr int SH_WARPS = SH_THREADS / 32;

// Backward of the shape-only path, one kernel: the translation column of dA (sum_v W[v,j] g_v; the rotation
// block is written as zero, it only feeds the pose gradients a shape-only stage does not use) and the beta
// entries of dX (sum_v S_v^T T_v^T g_v, what the blend contraction's backward would deliver; the
// pose-feature entries are written as zero).  Two hands per stage — their caches (2 x 37,344 B) and vertex
// gradients (18,672 B, 16-byte aligned only for an even hand) arrive by two TMA bulk copies — two stages.
constexpr int SHB_STAGES = 2;
constexpr int SHB_HPC = 16;                                                    // hands per CTA (even)
constexpr uint32_t SHB_G_BYTES = 2 * NV * 3 * sizeof(float);                   // 18,672
constexpr int SHB_STAGE_F4 = 2 * 3 * NV + (int)(SHB_G_BYTES / 16);             // float4 per stage
static_assert(SHB_G_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

__device__ __forceinline__ void shape_issue_pair(const float4* cache, const float* gverts, size_t h, float4* stage,
                                                 unsigned long long* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(2 * SHAPE_CACHE_BYTES + SHB_G_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(stage)), "l"(cache + h * 3 * NV), "r"(2 * SHAPE_CACHE_BYTES), "r"(b) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(stage + 2 * 3 * NV)), "l"(gverts + h * NV * 3), "r"(SHB_G_BYTES), "r"(b) : "memory");
}

__global__ void __launch_bounds__(SH_THREADS, 1)
k_shape_bwd(int n, const float* __restrict__ W4, const float* __restrict__ Sv, const float4* __restrict__ cache,
            const float* __restrict__ gverts, const float* __restrict__ gtips, float* __restrict__ dA,
            float* __restrict__ dX, const uint8_t* __restrict__ gzero) {
    extern __shared__ float4 smem4[];
    float4* ring = smem4;                                                          // [SHB_STAGES][SHB_STAGE_F4]
    float* sPart = reinterpret_cast<float*>(ring + SHB_STAGES * SHB_STAGE_F4);      // [2][SH_WARPS][64]: 48 dta + 16 dbeta
    float* sTips = sPart + 2 * SH_WARPS * 64;                                       // [SHB_HPC][16]
    __shared__ __align__(8) unsigned long long bars[SHB_STAGES];
    // Pairs (= frames) in which neither hand has a collision gradient (gzero) carry only fingertip gradients: they
    // are not streamed at all (s_sparse) and are finished from five cache rows per hand at the end.
    __shared__ int s_dense[SHB_HPC / 2], s_sparse[SHB_HPC / 2], s_nd, s_ns;
    __shared__ float sQ[SHB_HPC][5][8];              // per sparse hand and fingertip: g (3), q = T^T g (3)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h0 = blockIdx.x * SHB_HPC, nh = min(SHB_HPC, n - h0), npair = nh / 2;
    if (tid == 0) {
        int nd = 0, ns = 0;
        for (int p = 0; p < npair; ++p) {
            if (gzero && gzero[h0 + 2 * p] && gzero[h0 + 2 * p + 1]) s_sparse[ns++] = p;
            else s_dense[nd++] = p;
        }
        s_nd = nd; s_ns = ns;
        for (int k = 0; k < SHB_STAGES; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bars[k])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int k = 0; k < SHB_STAGES && k < nd; ++k)
            shape_issue_pair(cache, gverts, (size_t)h0 + 2 * s_dense[k], ring + k * SHB_STAGE_F4, &bars[k]);
    }
    for (int i = tid; i < nh * 15; i += SH_THREADS) sTips[(i / 15) * 16 + i % 15] = gtips ? gtips[(size_t)h0 * 15 + i] : 0.f;
    float sv[2][32];
    int vv[2], tip[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const int v = tid + sl * SH_THREADS;
        vv[sl] = min(v, NV - 1);
        load_shape_row(Sv, vv[sl], sv[sl]);
        tip[sl] = (v == 744) ? 0 : (v == 320) ? 1 : (v == 443) ? 2 : (v == 554) ? 3 : (v == 671) ? 4 : -1;
    }
    const float4* W44 = reinterpret_cast<const float4*>(W4);
    __syncthreads();                                            // barriers initialised, tips staged, pair lists built
    const int nd = s_nd, ns = s_ns;
    for (int q = 0; q < nd; ++q) {
        const int p = s_dense[q];
        const int st = q % SHB_STAGES;
        shape_wait(&bars[st], (q / SHB_STAGES) & 1);
        const float4* stage = ring + st * SHB_STAGE_F4;
        const float* G = reinterpret_cast<const float*>(stage + 2 * 3 * NV);
#pragma unroll
        for (int hl = 0; hl < 2; ++hl) {
            const int hh = 2 * p + hl;
            const float4* rows = stage + hl * 3 * NV;
            const bool hzero = gzero && gzero[h0 + hh];         // its vertex gradients were not written: zeros
            float g[2][3];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const bool ok = tid + sl * SH_THREADS < NV && !hzero;
#pragma unroll
                for (int c = 0; c < 3; ++c) g[sl][c] = ok ? G[(hl * NV + vv[sl]) * 3 + c] : 0.f;
                if (tip[sl] >= 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) g[sl][c] += sTips[hh * 16 + tip[sl] * 3 + c];
                }
            }
            // d beta: q = T^T g is the gradient of the shaped (unposed) vertex
            float accb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) accb[i] = 0.f;
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const float4 r0 = rows[vv[sl]], r1 = rows[NV + vv[sl]], r2 = rows[2 * NV + vv[sl]];
                const float q[3] = {r0.x * g[sl][0] + r1.x * g[sl][1] + r2.x * g[sl][2],
                                    r0.y * g[sl][0] + r1.y * g[sl][1] + r2.y * g[sl][2],
                                    r0.z * g[sl][0] + r1.z * g[sl][1] + r2.z * g[sl][2]};
#pragma unroll
                for (int k = 0; k < NB; ++k) accb[k] += sv[sl][k] * q[0] + sv[sl][NB + k] * q[1] + sv[sl][2 * NB + k] * q[2];
            }
            warp_reduce_scatter_16(accb, lane);
            if ((lane & 1) == 0)
                sPart[(hl * SH_WARPS + warp) * 64 + 48 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = accb[0];
            // d a_j: sum_v W[v,j] g_v
            float acc[48];
#pragma unroll
            for (int i = 0; i < 48; ++i) acc[i] = 0.f;
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 w4 = W44[t * NV + vv[sl]];
                    const float wj[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int r = 0; r < 3; ++r) acc[(t * 4 + i) * 3 + r] += wj[i] * g[sl][r];
                }
            }
            warp_reduce_scatter_48(acc, lane);
            if ((lane & 1) == 0) {
                const int seg = ((lane >> 4) & 1) * 24 + ((lane >> 3) & 1) * 12 + ((lane >> 2) & 1) * 6 + ((lane >> 1) & 1) * 3;
#pragma unroll
                for (int i = 0; i < 3; ++i) sPart[(hl * SH_WARPS + warp) * 64 + seg + i] = acc[i];
            }
        }
        __syncthreads();                                        // partial sums visible, the stage is free
        if (tid == 0 && q + SHB_STAGES < nd)
            shape_issue_pair(cache, gverts, (size_t)h0 + 2 * s_dense[q + SHB_STAGES], ring + st * SHB_STAGE_F4, &bars[st]);
        for (int i = tid; i < 2 * (192 + KP); i += SH_THREADS) {
            const int hl = i / (192 + KP), x = i % (192 + KP);
            const size_t h = (size_t)h0 + 2 * p + hl;
            const float* part = sPart + hl * SH_WARPS * 64;
            float sum = 0.f;
            if (x < 192) {
                if ((x & 3) == 3) {
                    const int e = (x / 12) * 3 + (x % 12) / 4;
#pragma unroll
                    for (int w = 0; w < SH_WARPS; ++w) sum += part[w * 64 + e];
                }
                dA[h * 192 + x] = sum;
            } else {
                const int k = x - 192;
                if (k >= NPF && k < NPF + NB) {
#pragma unroll
                    for (int w = 0; w < SH_WARPS; ++w) sum += part[w * 64 + 48 + (k - NPF)];
                }
                dX[h * KP + k] = sum;
            }
        }
        __syncthreads();                                        // sPart is rewritten by the next pair
    }
    // ---- pairs without collision gradient: the same two sums over the five fingertip vertices only
    if (ns > 0) {
        const int tips[5] = {744, 320, 443, 554, 671};
        for (int i = tid; i < ns * 2 * 5; i += SH_THREADS) {
            const int hh = 2 * s_sparse[i / 10] + (i % 10) / 5, t = i % 5, v = tips[t];
            const float4* rows = cache + ((size_t)h0 + hh) * 3 * NV;
            const float4 r0 = rows[v], r1 = rows[NV + v], r2 = rows[2 * NV + v];
            const float g[3] = {sTips[hh * 16 + t * 3], sTips[hh * 16 + t * 3 + 1], sTips[hh * 16 + t * 3 + 2]};
            float* o = sQ[hh][t];
            o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
            o[3] = r0.x * g[0] + r1.x * g[1] + r2.x * g[2];
            o[4] = r0.y * g[0] + r1.y * g[1] + r2.y * g[2];
            o[5] = r0.z * g[0] + r1.z * g[1] + r2.z * g[2];
        }
        __syncthreads();
        for (int i = tid; i < ns * 2 * (192 + KP); i += SH_THREADS) {
            const int hh = 2 * s_sparse[i / (2 * (192 + KP))] + (i / (192 + KP)) % 2, x = i % (192 + KP);
            const size_t h = (size_t)h0 + hh;
            float sum = 0.f;
            if (x < 192) {
                if ((x & 3) == 3) {
                    const int j = x / 12, r = (x % 12) / 4;
#pragma unroll
                    for (int t = 0; t < 5; ++t) sum += W4[((size_t)(j >> 2) * NV + tips[t]) * 4 + (j & 3)] * sQ[hh][t][r];
                }
                dA[h * 192 + x] = sum;
            } else {
                const int k = x - 192;
                if (k >= NPF && k < NPF + NB) {
#pragma unroll
                    for (int t = 0; t < 5; ++t)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int e = c * NB + (k - NPF);          // entry e of the vertex's shape row: plane e / 4, lane e % 4
                            sum += Sv[((size_t)(e >> 2) * NV + tips[t]) * 4 + (e & 3)] * sQ[hh][t][3 + c];
                        }
                }
                dX[h * KP + k] = sum;
            }
        }
    }
}

constexpr size_t SHAPE_SMEM = sizeof(float4) * (4 * NV + SH_HPC * 48) + sizeof(float) * SH_HPC * 12;
constexpr size_t SHAPE_BWD_SMEM = sizeof(float4) * SHB_STAGES * SHB_STAGE_F4 + sizeof(float) * (2 * SH_WARPS * 64 + SHB_HPC * 16);

int launch_shape_prep(const ihmr_model* m, int n, HandSrc src, const float* off, const float* A, float* cache, cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    static unsigned long long configured = 0ull;
    if (int rc = ensure_dynamic_smem(k_shape_prep, SHAPE_SMEM, configured)) return rc;
    k_shape_prep<<<(n + SH_HPC - 1) / SH_HPC, SH_THREADS, SHAPE_SMEM, st>>>(n, src, off, A, m->vtemp, m->W4, m->Sv,
                                                                            reinterpret_cast<float4*>(cache));
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

constexpr size_t SHAPE_FWD_SMEM = sizeof(float4) * (SHF_STAGES * 3 * NV + SHF_HPC * NJ) + sizeof(float) * SHF_HPC * 12;

int launch_shape_fwd(const ihmr_model* m, int n, HandSrc src, const float* A, const float* cache, float* verts, cudaStream_t st,
                     float* bbox) {
    if (n <= 0) return IHMR_OK;
    static unsigned long long configured = 0ull;
    if (int rc = ensure_dynamic_smem(k_shape_fwd, SHAPE_FWD_SMEM, configured)) return rc;
    k_shape_fwd<<<(n + SHF_HPC - 1) / SHF_HPC, SH_THREADS, SHAPE_FWD_SMEM, st>>>(n, src, A, m->W4, m->Sv,
                                                                               reinterpret_cast<const float4*>(cache), verts, bbox);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_shape_bwd(const ihmr_model* m, int n, const float* cache, const float* gverts, const float* gtips, float* dA,
                     float* dX, cudaStream_t st, SparseGrad sp) {
    if (n <= 0) return IHMR_OK;
    if (n & 1) { set_error("shape_bwd: the hands come in pairs (two per frame)"); return IHMR_E_INVALID; }
    static unsigned long long configured = 0ull;
    if (int rc = ensure_dynamic_smem(k_shape_bwd, SHAPE_BWD_SMEM, configured)) return rc;
    k_shape_bwd<<<(n + SHB_HPC - 1) / SHB_HPC, SH_THREADS, SHAPE_BWD_SMEM, st>>>(n, m->W4, m->Sv, reinterpret_cast<const float4*>(cache),
                                                                               gverts, gtips, dA, dX, sp.gzero);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

// -------------------------------------------------------------------------------- launchers
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t mano_ws_bytes(int n) {
    size_t b = 0;
    b += align_up((size_t)n * KP * 4, 256);          // X
    b += align_up((size_t)n * 192 * 4, 256);         // A
    b += align_up((size_t)n * LDN * 4, 256);         // off
    b += align_up((size_t)n * LDN * 4, 256);         // gposed
    b += align_up((size_t)n * 192 * 4, 256);         // dA
    b += align_up((size_t)n * KP * 4, 256);          // dX
    b += align_up((size_t)n * 48 * 4, 256);          // joints
    return b;
}

ManoWs mano_ws_carve(void* base, int n) {
    ManoWs w;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) { float* r = reinterpret_cast<float*>(p); p += align_up(bytes, 256); return r; };
    w.X = take((size_t)n * KP * 4);
    w.A = take((size_t)n * 192 * 4);
    w.off = take((size_t)n * LDN * 4);
    w.gposed = take((size_t)n * LDN * 4);
    w.dA = take((size_t)n * 192 * 4);
    w.dX = take((size_t)n * KP * 4);
    w.joints = take((size_t)n * 48 * 4);
    return w;
}

int launch_pose_prep(const ihmr_model* m, int n, HandSrc src, float* X, float* A, float* joints, cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    Tree tree = make_tree(m->parents);
    dim3 grid((n * 16 + 127) / 128);
    if (src.params)
        k_pose_prep<true><<<grid, 128, 0, st>>>(n, src, m->hands_mean, m->Jt, m->Js, tree, X, A, joints);
    else
        k_pose_prep<false><<<grid, 128, 0, st>>>(n, src, m->hands_mean, m->Jt, m->Js, tree, X, A, joints);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_blend_fwd(const ihmr_model* m, int n, const float* X, float* off, cudaStream_t st) {
    // off (n x 2336) = X (n x 160) . D  ==  X . (D^T)^T  with D^T (2336 x 160) K-major
    return launch_gemm_tf32x3(n, LDN, KP, X, KP, m->DT, KP, off, LDN, st, nullptr, nullptr, m->DTq);
}

constexpr int BLEND_BWD_KSPLIT = 4;     // fixed: the K = 2336 contraction is cut the same way whatever the batch size

int launch_blend_bwd(const ihmr_model* m, int n, const float* gposed, float* dX, cudaStream_t st, SparseGrad sp, float* scratch) {
    // dX (n x 160) = gposed (n x 2336) . D^T  with D (160 x 2336) K-major; with a dense-hand list only those rows.
    // 73 K-chunks per row tile: cut into four slices (four times the CTAs, which is what small and medium batches lack)
    static_assert((size_t)BLEND_BWD_KSPLIT * KP <= (size_t)LDN, "the partial products fit the scratch rows");
    return launch_gemm_tf32x3(n, KP, LDN, gposed, LDN, m->D, LDN, dX, KP, st, sp.dense_list, sp.dense_count, m->Dq,
                              scratch ? BLEND_BWD_KSPLIT : 1, scratch);
}

int launch_sgemm_reference(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                           cudaStream_t st) {
    if (M <= 0) return IHMR_OK;
    dim3 grid((N + 127) / 128, (M + 127) / 128);
    k_sgemm<128, 128, 8, 8, 8><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_skin_fwd(const ihmr_model* m, int n, const float* off, const float* A, float* verts, cudaStream_t st, float* bbox) {
    if (n <= 0) return IHMR_OK;
    return launch_skin_fwd_tc(m, n, off, A, verts, st, bbox);      // the blend T = W . A^T runs on tcgen05 (blend_tc.cu)
}

int launch_skin_bwd(const ihmr_model* m, int n, const float* off, const float* A, const float* gverts,
                    const float* gtips, float* gposed, float* dA, cudaStream_t st, SparseGrad sp, float* dX) {
    if (n <= 0) return IHMR_OK;
    if (sp.gzero && !dX) { set_error("skin_bwd: the sparse path needs dX"); return IHMR_E_INVALID; }
    static unsigned long long configured = 0ull;
    if (int rc = ensure_dynamic_smem(k_skin_bwd, SKIN_BWD_SMEM, configured)) return rc;
    if (sp.gzero && !sp.dense_list) { set_error("skin_bwd: the sparse path needs the list of dense hands"); return IHMR_E_INVALID; }
    k_skin_bwd<<<(n + SK_HPC - 1) / SK_HPC, SKB_THREADS, SKIN_BWD_SMEM, st>>>(n, off, A, m->vtemp, m->W4,
                                                                           gverts, gtips, gposed, dA, sp.dense_list, sp.dense_count);
    IHMR_LAUNCH_OK();
    if (sp.gzero) {
        k_skin_bwd_tips<<<(n + 7) / 8, 256, 0, st>>>(n, off, A, m->vtemp, m->Wt, gtips, sp.gzero, m->DT, dA, dX);
        IHMR_LAUNCH_OK();
    }
    return IHMR_OK;
}

int launch_pose_bwd(const ihmr_model* m, int n, HandSrc src, const float* dA, const float* gjoints,
                    const float* dX, HandGrad out, cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    Tree tree = make_tree(m->parents);
    dim3 grid((n * 16 + 127) / 128);
    if (src.params)
        k_pose_bwd<true><<<grid, 128, 0, st>>>(n, src, m->hands_mean, m->Jt, m->Js, tree, dA, gjoints, dX, out);
    else
        k_pose_bwd<false><<<grid, 128, 0, st>>>(n, src, m->hands_mean, m->Jt, m->Js, tree, dA, gjoints, dX, out);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

}  // namespace ihmr
