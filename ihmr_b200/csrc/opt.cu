// Fused IHMR-OPT iteration pieces that are not the MANO layer or the penetration kernel:
// two-hand glue, reprojection / joint / prior losses with their analytic gradients, online
// snapshot selection and the optimiser step.
#include <math.h>

#include "kernels.cuh"

namespace ihmr {

constexpr int FL_THREADS = 64;   // one frame per 64-thread block; thread t < 42 <-> joint t

struct FrameLossArgs {
    int B;
    float inv_n;                  // 1 / bs_norm
    const float* params;          // (B,122)
    const float* joints16;        // (B,2,16,3) posed joints, left hand in its mirrored model frame
    const float* verts;           // (B,2,778,3) same frames
    ihmr_targets_t tg;
    float w2d, w3d, wtrans, wshape, wfinger;
    const float* col_loss;        // (B) unweighted, masked collision loss from the sdf kernels, or
    const float* col_parts;       // (B,2) its two unmasked direction sums (loss = mask * (p0 + p1) / 4)
    const float* gshift_col;      // (B,3) collision gradient w.r.t. the left-hand shift (scaled) or null
    const uint8_t* gzero;         // (B,2) or null: hands without vertex gradient ...
    int* dense_list;              // ... the others are appended here (any order) for the blend contraction
    int* dense_count;
    // outputs
    float* gjoints16;             // (B,2,16,3) or null
    float* gtips;                 // (B,2,5,3) or null
    float* grad;                  // (B,122) or null: fully initialised here (pose/shape slots: direct terms only)
    float* loss_parts;            // (B,6) per-frame weighted contributions to the six batch losses, or null
    float* j2d_batch;             // (B) joints_2d_loss_p_batch, or null
    float* j3d_batch;             // (B) joints_3d_loss_p_batch, or null
    float* joints_out;            // (B,42,3) root-aligned joints, or null
    float* crit3;                 // (B,3) [joints_3d_loss_p, collision_loss, joints_2d_loss_p] (IHMR_LOSS_* order), or null
    // online snapshot selection
    int snap_mode;                // 0 none, 1 first snapshot of the stage, 2 later snapshot
    int n_filters;
    int filter_loss[4];
    float filter_factor[4];       // 1 + (percent + 0.1) / 100
    int select_loss;
    float* origin;                // (B,3) criteria at snapshot 0, indexed by IHMR_LOSS_*
    float* best;                  // (B)   best selected criterion so far
    int* take;                    // (B)   1 if the current parameters become the stage's best
};

__device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// fixed-order sums over the 64 threads of the block for N values behind one pair of barriers (warp butterfly, then
// warp 0 + warp 1); results in every thread
template <int N>
__device__ __forceinline__ void block64_sum_n(float (&v)[N], float* red) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) red[(threadIdx.x >> 5) * N + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = red[i] + red[N + i];
}

// Online form of filter_by_losses + select_params (src/utils/opt_utils.py:104-152) for one frame and one
// snapshot: crit = this snapshot's [joints_3d_loss_p, collision_loss, joints_2d_loss_p]; org = the same at snapshot 0;
// best = the best selected criterion so far.  Returns true when the snapshot becomes the frame's choice.
//   valid  <=> crit[f] <= org[f] * (1 + (percent + 0.1) / 100) for every filter f      (:111-113, '<=' at :131)
//   score   =  valid ? crit[select] : 1e11                                             (:134-139)
//   snapshot 0 is always eligible with its true score (:140); argmin keeps the FIRST minimum (:146), hence '<'.
__device__ __forceinline__ bool snapshot_decide(bool first, const float* crit, float* org, float* best, int n_filters,
                                                const int* filter_loss, const float* filter_factor, int select_loss) {
    if (first) {
        org[0] = crit[0]; org[1] = crit[1]; org[2] = crit[2];
        *best = crit[select_loss];
        return true;
    }
    bool ok = true;
    for (int f = 0; f < n_filters; ++f) ok = ok && (crit[filter_loss[f]] <= org[filter_loss[f]] * filter_factor[f]);
    const float score = ok ? crit[select_loss] : 100000000000.0f;
    const bool better = score < *best;
    if (better) *best = score;
    return better;
}

// test / tooling entry: the selection alone over stacked per-snapshot criteria (S,B,3) -> chosen snapshot index (B)
__global__ void k_select_snapshots(int S, int B, const float* __restrict__ crit, int n_filters, int f0, int f1, int f2, int f3,
                                   float p0, float p1, float p2, float p3, int select_loss, int* __restrict__ index) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int fl[4] = {f0, f1, f2, f3};
    const float ff[4] = {p0, p1, p2, p3};
    float org[3], best = 0.f;
    int pick = 0;
    for (int sidx = 0; sidx < S; ++sidx) {
        const float* c = crit + ((size_t)sidx * B + b) * 3;
        const float cc[3] = {c[0], c[1], c[2]};
        if (snapshot_decide(sidx == 0, cc, org, &best, n_filters, fl, ff, select_loss)) pick = sidx;
    }
    index[b] = pick;
}

__global__ void __launch_bounds__(FL_THREADS) k_frame_loss(FrameLossArgs a) {
    __shared__ float sJ[42][3];     // joints at the current alignment stage
    __shared__ float sG[42][3];     // gradient w.r.t. the aligned joints
    __shared__ float red[2 * 9];
    const int b = blockIdx.x, t = threadIdx.x;
    const bool isj = t < 42;
    const int hand = isj ? t / 21 : 0, k = isj ? t % 21 : 0;
    const float* prm = a.params + (size_t)b * PD;
    const float* jr = a.joints16 + ((size_t)b * 2 + 0) * 48;
    const float* jl = a.joints16 + ((size_t)b * 2 + 1) * 48;
    // shift = trans + J_R[0] - X J_L'[0]            (optimize_model.py:222-225)
    const float shift[3] = {prm[P_TRANS + 0] + (jr[0] + jl[0]), prm[P_TRANS + 1] + (jr[1] - jl[1]),
                            prm[P_TRANS + 2] + (jr[2] - jl[2])};
    float J[3] = {0.f, 0.f, 0.f};
    if (isj) {
        const int tips[5] = {744, 320, 443, 554, 671};
        const float* src = (k < 16) ? a.joints16 + (((size_t)b * 2 + hand) * 16 + k) * 3
                                    : a.verts + (((size_t)b * 2 + hand) * NV + tips[k - 16]) * 3;
        J[0] = src[0]; J[1] = src[1]; J[2] = src[2];
        if (hand == 1) { J[0] = -J[0] + shift[0]; J[1] += shift[1]; J[2] += shift[2]; }
    }
    // ---- 2-D reprojection, L1 against init_joints_2d     (transform_utils.py:47-53, loss_utils.py:82-87)
    const float cs = prm[P_CAM], cx = prm[P_CAM + 1], cy = prm[P_CAM + 2];
    float gJ[3] = {0.f, 0.f, 0.f};
    float l2d = 0.f, dcs = 0.f, dcx = 0.f, dcy = 0.f;
    const float k2 = a.w2d * a.inv_n / 84.0f;
    if (isj) {
        const float* tg = a.tg.init_joints_2d + ((size_t)b * 42 + t) * 3;
        const float w = tg[2];
        const float px = cs * (J[0] + cx), py = cs * (J[1] + cy);
        const float dx = tg[0] - px, dy = tg[1] - py;
        l2d = (fabsf(dx) + fabsf(dy)) * w;
        const float gx = -((dx > 0.f) - (dx < 0.f)) * w * k2, gy = -((dy > 0.f) - (dy < 0.f)) * w * k2;
        gJ[0] = cs * gx; gJ[1] = cs * gy;
        dcs = gx * (J[0] + cx) + gy * (J[1] + cy);
        dcx = cs * gx; dcy = cs * gy;
        sJ[t][0] = J[0]; sJ[t][1] = J[1]; sJ[t][2] = J[2];
        sG[t][0] = 0.f; sG[t][1] = 0.f; sG[t][2] = 0.f;
    }
    __syncthreads();
    // ---- root alignment, twice: GT weights then init weights  (loss_utils.py:90-103, optimize_model.py:292-298)
    const float wgt0 = a.tg.gt_joints_3d[((size_t)b * 42) * 4 + 3];
    const float w30 = a.tg.init_joints_3d[((size_t)b * 42) * 4 + 3];
    const int r1 = (wgt0 > 0.5f) ? 0 : ((wgt0 < 1e-7f) ? 21 : -1);
    const int r2 = (w30 > 0.5f) ? 0 : ((w30 < 1e-7f) ? 21 : -1);
    float Ja[3] = {J[0], J[1], J[2]};
    if (r1 >= 0) { Ja[0] -= sJ[r1][0]; Ja[1] -= sJ[r1][1]; Ja[2] -= sJ[r1][2]; }
    __syncthreads();
    if (isj) { sJ[t][0] = Ja[0]; sJ[t][1] = Ja[1]; sJ[t][2] = Ja[2]; }
    __syncthreads();
    if (r2 >= 0) { const float rx = sJ[r2][0], ry = sJ[r2][1], rz = sJ[r2][2]; Ja[0] -= rx; Ja[1] -= ry; Ja[2] -= rz; }
    __syncthreads();
    if (isj) { sJ[t][0] = Ja[0]; sJ[t][1] = Ja[1]; sJ[t][2] = Ja[2]; }
    __syncthreads();
    if (a.joints_out && isj) {
        float* o = a.joints_out + ((size_t)b * 42 + t) * 3;
        o[0] = Ja[0]; o[1] = Ja[1]; o[2] = Ja[2];
    }
    // ---- 3-D joints against init_joints_3d (target aligned by the same rule)
    float l3d = 0.f, ga[3] = {0.f, 0.f, 0.f};
    const float k3 = a.w3d * a.inv_n / 126.0f;
    if (isj) {
        const float* tg = a.tg.init_joints_3d + ((size_t)b * 42 + t) * 4;
        float T[3] = {tg[0], tg[1], tg[2]};
        if (r2 >= 0) {
            const float* tr = a.tg.init_joints_3d + ((size_t)b * 42 + r2) * 4;
            T[0] -= tr[0]; T[1] -= tr[1]; T[2] -= tr[2];
        }
        const float w = tg[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = T[c] - Ja[c];
            l3d += d * d * w;
            ga[c] = -2.0f * d * w * k3;
        }
    }
    // ---- finger regulariser on the aligned joints (loss_utils.py:138-171), one thread per finger
    float lfin = 0.f;
    if (t < 10 && a.wfinger != 0.f) {
        const int chains[5][4] = {{1, 2, 3, 17}, {4, 5, 6, 18}, {7, 8, 9, 20}, {10, 11, 12, 19}, {13, 14, 15, 16}};
        const int base = (t / 5) * 21;
        int id[4];
        float p[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            id[i] = base + chains[t % 5][i];
#pragma unroll
            for (int c = 0; c < 3; ++c) p[i][c] = sJ[id[i]][c];
        }
        float b0[3], b1[3], b2[3], c01[3], e[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { b0[c] = p[0][c] - p[1][c]; b1[c] = p[1][c] - p[2][c]; b2[c] = p[2][c] - p[3][c]; }
        cross3(b0, b1, c01);
        cross3(b1, b2, e);
        const float C1 = b2[0] * c01[0] + b2[1] * c01[1] + b2[2] * c01[2];
        const float C2 = c01[0] * e[0] + c01[1] * e[1] + c01[2] * e[2];
        lfin = fabsf(C1) - fminf(0.f, C2);
        const float kf = a.wfinger * a.inv_n;
        const float g1 = kf * ((C1 > 0.f) - (C1 < 0.f)), g2 = (C2 < 0.f) ? -kf : 0.f;
        float gc[3], gb2[3], ge[3], gb0[3], gb1[3], tmp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { gc[c] = g1 * b2[c] + g2 * e[c]; gb2[c] = g1 * c01[c]; ge[c] = g2 * c01[c]; }
        cross3(b2, ge, gb1);                 // e = b1 x b2 : d b1 = b2 x ge
        cross3(ge, b1, tmp);                 //               d b2 = ge x b1
#pragma unroll
        for (int c = 0; c < 3; ++c) gb2[c] += tmp[c];
        cross3(b1, gc, gb0);                 // c01 = b0 x b1 : d b0 = b1 x gc
        cross3(gc, b0, tmp);                 //                 d b1 += gc x b0
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gb1[c] += tmp[c];
            sG[id[0]][c] = gb0[c];
            sG[id[1]][c] = gb1[c] - gb0[c];
            sG[id[2]][c] = gb2[c] - gb1[c];
            sG[id[3]][c] = -gb2[c];
        }
    }
    __syncthreads();
    if (isj) { ga[0] += sG[t][0]; ga[1] += sG[t][1]; ga[2] += sG[t][2]; }
    // ---- back through the two alignments: g <- g - e_root * sum(g)
    {
        float sm[3] = {ga[0], ga[1], ga[2]};
        block64_sum_n(sm, red);
        if (r2 >= 0 && t == r2) { ga[0] -= sm[0]; ga[1] -= sm[1]; ga[2] -= sm[2]; }
        sm[0] = ga[0]; sm[1] = ga[1]; sm[2] = ga[2];
        block64_sum_n(sm, red);
        if (r1 >= 0 && t == r1) { ga[0] -= sm[0]; ga[1] -= sm[1]; ga[2] -= sm[2]; }
    }
    gJ[0] += ga[0]; gJ[1] += ga[1]; gJ[2] += ga[2];
    if (a.dense_list && t < 2 && !a.gzero[b * 2 + t]) a.dense_list[atomicAdd(a.dense_count, 1)] = b * 2 + t;
    // ---- shift gradient: all left-hand joints move with it
    //      and the per-frame scalar terms, all nine sums behind one pair of barriers
    const bool lj = isj && hand == 1;
    float sums[9] = {lj ? gJ[0] : 0.f, lj ? gJ[1] : 0.f, lj ? gJ[2] : 0.f, l2d, l3d, lfin, dcs, dcx, dcy};
    block64_sum_n(sums, red);
    float gs[3] = {sums[0], sums[1], sums[2]};
    if (a.gshift_col) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gs[c] += a.gshift_col[(size_t)b * 3 + c];
    }
    const float sum2d = sums[3], sum3d = sums[4], sumfin = sums[5], gcs = sums[6], gcx = sums[7], gcy = sums[8];

    if (isj && a.gjoints16) {
        float g[3] = {gJ[0], gJ[1], gJ[2]};
        if (k == 0) {   // wrists also carry the shift: shift = t + J_R[0] - X J_L'[0]
            if (hand == 0) { g[0] += gs[0]; g[1] += gs[1]; g[2] += gs[2]; }
            else { g[0] -= gs[0]; g[1] -= gs[1]; g[2] -= gs[2]; }
        }
        if (hand == 1) g[0] = -g[0];          // back to the mirrored model frame
        float* dst = (k < 16) ? a.gjoints16 + (((size_t)b * 2 + hand) * 16 + k) * 3
                              : a.gtips + (((size_t)b * 2 + hand) * 5 + (k - 16)) * 3;
        dst[0] = g[0]; dst[1] = g[1]; dst[2] = g[2];
    }
    // translation loss (loss_utils.py:114-118) and shape regulariser (:121-128)
    const float* tj = a.tg.init_hand_trans_j + (size_t)b * 4;
    float ltr = 0.f, gtr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float d = tj[c] - prm[P_TRANS + c];
        ltr += d * d * tj[3];
        gtr[c] = -2.0f * d * tj[3] * a.wtrans * a.inv_n / 3.0f;
    }
    float lsh = 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i) { const float d = prm[P_R_SHAPE + i] - prm[P_L_SHAPE + i]; lsh += d * d; }

    if (a.grad) {
        float* g = a.grad + (size_t)b * PD;
        for (int i = t; i < PD; i += FL_THREADS) {
            float v = 0.f;
            if (i == P_CAM) v = gcs;
            else if (i == P_CAM + 1) v = gcx;
            else if (i == P_CAM + 2) v = gcy;
            else if (i >= P_TRANS && i < P_TRANS + 3) v = gs[i - P_TRANS] + gtr[i - P_TRANS];
            else if (i >= P_R_SHAPE && i < P_R_SHAPE + NB)
                v = 2.0f * (prm[i] - prm[i + NB]) * a.wshape * a.inv_n / 10.0f;
            else if (i >= P_L_SHAPE && i < P_L_SHAPE + NB)
                v = -2.0f * (prm[i - NB] - prm[i]) * a.wshape * a.inv_n / 10.0f;
            g[i] = v;
        }
    }
    if (t == 0) {
        const float j2d_b = sum2d / 84.0f * a.w2d, j3d_b = sum3d / 126.0f * a.w3d;
        float col = a.col_loss ? a.col_loss[b] : 0.f;
        if (a.col_parts) {
            const float* ht = a.tg.hand_type_array + (size_t)b * 2;
            col = ((ht[0] + ht[1] > 1.5f) ? 1.0f : 0.0f) * (a.col_parts[b * 2] + a.col_parts[b * 2 + 1]) * 0.25f;
        }
        if (a.crit3) { a.crit3[b * 3] = j3d_b; a.crit3[b * 3 + 1] = col; a.crit3[b * 3 + 2] = j2d_b; }
        if (a.j2d_batch) a.j2d_batch[b] = j2d_b;
        if (a.j3d_batch) a.j3d_batch[b] = j3d_b;
        if (a.loss_parts) {
            float* lp = a.loss_parts + (size_t)b * 6;
            lp[0] = sum2d * k2; lp[1] = sum3d * k3; lp[2] = ltr * a.wtrans * a.inv_n / 3.0f;
            lp[3] = col * a.inv_n;            // caller applies the collision weight
            lp[4] = lsh * a.wshape * a.inv_n / 10.0f; lp[5] = sumfin * a.wfinger * a.inv_n;
        }
        if (a.snap_mode) {
            const float crit[3] = {j3d_b, col, j2d_b};
            a.take[b] = snapshot_decide(a.snap_mode == 1, crit, a.origin + (size_t)b * 3, a.best + b, a.n_filters,
                                        a.filter_loss, a.filter_factor, a.select_loss) ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------------------ optimiser
struct StepArgs {
    int B;
    uint32_t mask;
    int optimizer;
    float lr, step_size, bc2_sqrt, eps, momentum;
    int first_step;
    float* params;
    const float* grad;
    float* m;
    float* v;
    const int* take;       // null when this iteration takes no snapshot
    float* best_params;
};

__device__ __forceinline__ uint32_t param_group(int i) {
    if (i < 3) return IHMR_P_CAM;
    if (i < 6) return IHMR_P_TRANS;
    if (i < 9) return IHMR_P_R_ORIENT;
    if (i < 54) return IHMR_P_R_POSE;
    if (i < 57) return IHMR_P_L_ORIENT;
    if (i < 102) return IHMR_P_L_POSE;
    if (i < 112) return IHMR_P_R_SHAPE;
    return IHMR_P_L_SHAPE;
}

// torch.optim.Adam (betas 0.9/0.999, eps 1e-8) / SGD(momentum 0.9) on the live parameters
// (optimize_model.py:343-347, 404-406); the snapshot copy happens BEFORE the step (:402-406).
__global__ void k_step(StepArgs a) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)a.B * PD) return;
    const int b = (int)(idx / PD), i = (int)(idx % PD);
    if (!(param_group(i) & a.mask)) return;
    float p = a.params[idx];
    if (a.take && a.take[b]) a.best_params[idx] = p;
    const float g = a.grad[idx];
    if (a.optimizer == IHMR_OPT_ADAM) {
        const float m = 0.9f * a.m[idx] + 0.1f * g;
        const float v = 0.999f * a.v[idx] + 0.001f * g * g;
        a.m[idx] = m; a.v[idx] = v;
        const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
        p = p - a.step_size * (m / denom);
    } else {
        const float buf = a.first_step ? g : a.momentum * a.m[idx] + g;
        a.m[idx] = buf;
        p = p - a.lr * buf;
    }
    a.params[idx] = p;
}

__global__ void k_restore(int B, uint32_t mask, float* params, const float* best_params) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * PD) return;
    if (param_group((int)(idx % PD)) & mask) params[idx] = best_params[idx];
}

// deterministic column sums of (B,6) per-frame loss parts -> (6)
__global__ void k_loss_reduce(int B, const float* parts, float wcol, float* out) {
    __shared__ float red[8];
    const int c = blockIdx.x;
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) s += parts[(size_t)b * 6 + c];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        out[c] = (c == 3) ? tot * wcol : tot;
    }
}

// final export: world-frame vertices of both hands (optimize_model.py:204-228)
__global__ void k_export_verts(int B, const float* verts, const float* joints16, const float* params,
                               float* right, float* left) {
    const int b = blockIdx.x;
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= NV) return;
    const float* jr = joints16 + ((size_t)b * 2 + 0) * 48;
    const float* jl = joints16 + ((size_t)b * 2 + 1) * 48;
    const float* prm = params + (size_t)b * PD;
    const float* vr = verts + (((size_t)b * 2 + 0) * NV + i) * 3;
    const float* vl = verts + (((size_t)b * 2 + 1) * NV + i) * 3;
    float* r = right + ((size_t)b * NV + i) * 3;
    float* l = left + ((size_t)b * NV + i) * 3;
    r[0] = vr[0]; r[1] = vr[1]; r[2] = vr[2];
    l[0] = -vl[0] + (prm[P_TRANS + 0] + (jr[0] + jl[0]));
    l[1] = vl[1] + (prm[P_TRANS + 1] + (jr[1] - jl[1]));
    l[2] = vl[2] + (prm[P_TRANS + 2] + (jr[2] - jl[2]));
}

// percent = (float(criterion) + 0.1) / 100 ; bar = origin * (1 + percent)   (opt_utils.py:111-112)
static inline float filter_factor_of(float percent) { return (float)(1.0 + ((double)percent + 0.1) / 100.0); }

int select_snapshots(int S, int B, const float* crit, const ihmr_stage_t* stg, int* index, cudaStream_t st) {
    float ff[4] = {1.f, 1.f, 1.f, 1.f};
    int fl[4] = {0, 0, 0, 0};
    for (int f = 0; f < stg->n_filters; ++f) { ff[f] = filter_factor_of(stg->filter_percent[f]); fl[f] = stg->filter_loss[f]; }
    k_select_snapshots<<<(B + 127) / 128, 128, 0, st>>>(S, B, crit, stg->n_filters, fl[0], fl[1], fl[2], fl[3], ff[0], ff[1], ff[2], ff[3],
                                                         stg->select_loss, index);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

// ---------------------------------------------------------------------------- host driver
struct OptWs {
    ManoWs mano;
    float *verts, *gverts, *joints, *gjoints, *gtips, *grad, *m, *v, *best_params;
    float *col_loss, *gshift, *origin, *best, *j2d_b, *j3d_b, *loss_parts;
    float* shape_cache;       // (N, 778, 3) float4: (T_v | T_v c_v) of the shape-only stages
    int* take;
    void* sdf_ws;             // scratch of the penetration kernels (sdf_ws_bytes)
    uint16_t* sdf_hints;      // nearest-face seeds the penetration kernel carries from one iteration to the next
    uint32_t* sdf_pcache;     // static-grid caches of the penetration kernel (stages that move only hand_trans)
    float* sdf_phic;
    float* bbox;              // (N, 6) box of every hand's stored vertices, left by the kernel that wrote them
    uint8_t* gzero;           // (N) hands whose collision gradient is identically zero this iteration
    int* dense_list;          // (N) the other hands, [N] = their count
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t opt_ws_layout(void* base, int B, OptWs* out) {
    const int n = 2 * B;
    char* p = static_cast<char*>(base);
    size_t used = 0;
    auto take = [&](size_t bytes) {
        void* r = base ? p + used : nullptr;
        used += align_up(bytes, 256);
        return r;
    };
    const size_t mano = mano_ws_bytes(n);
    void* mano_base = take(mano);
    OptWs w;
    if (base) w.mano = mano_ws_carve(mano_base, n);
    w.verts = (float*)take((size_t)n * NV * 3 * 4);
    w.gverts = (float*)take((size_t)n * NV * 3 * 4);
    w.joints = (float*)take((size_t)n * 48 * 4);
    w.gjoints = (float*)take((size_t)n * 48 * 4);
    w.gtips = (float*)take((size_t)n * 15 * 4);
    w.grad = (float*)take((size_t)B * PD * 4);
    w.m = (float*)take((size_t)B * PD * 4);
    w.v = (float*)take((size_t)B * PD * 4);
    w.best_params = (float*)take((size_t)B * PD * 4);
    w.col_loss = (float*)take((size_t)B * 4);
    w.gshift = (float*)take((size_t)B * 3 * 4);
    w.origin = (float*)take((size_t)B * 3 * 4);
    w.best = (float*)take((size_t)B * 4);
    w.j2d_b = (float*)take((size_t)B * 4);
    w.j3d_b = (float*)take((size_t)B * 4);
    w.loss_parts = (float*)take((size_t)B * 6 * 4);
    w.take = (int*)take((size_t)B * 4);
    w.shape_cache = (float*)take((size_t)n * NV * 12 * 4);
    w.sdf_ws = take(sdf_ws_bytes(B));
    w.sdf_hints = (uint16_t*)take(sdf_hint_bytes(B));
    w.sdf_pcache = (uint32_t*)take(sdf_pcache_bytes(B));
    w.sdf_phic = (float*)take(sdf_phic_bytes(B));
    w.bbox = (float*)take((size_t)n * 6 * 4);
    w.gzero = (uint8_t*)take((size_t)n);
    w.dense_list = (int*)take((size_t)(n + 1) * 4);
    if (out) *out = w;
    return used;
}

size_t opt_ws_bytes(int B) { return opt_ws_layout(nullptr, B, nullptr); }

// Optional per-kernel timing of one iteration (ihmr_opt_profile_iteration): an event is
// recorded on the stream after each kernel class.  Off on the normal path.
constexpr int N_KERNEL_CLASSES = 9;   // pose_prep blend_fwd skin_fwd sdf frame_loss skin_bwd blend_bwd pose_bwd step
struct IterProf {
    cudaEvent_t ev[N_KERNEL_CLASSES + 1] = {};
    cudaStream_t st = nullptr;
    IterProf() = default;
    IterProf(const IterProf&) = delete;
    IterProf& operator=(const IterProf&) = delete;
    ~IterProf() { for (auto& e : ev) if (e) cudaEventDestroy(e); }      // every return path releases the events
    void tick(int i) { cudaEventRecord(ev[i], st); }
};
#define IHMR_TICK(prof, i) do { if (prof) (prof)->tick(i); } while (0)

// forward of both hands of every frame: X, A, joints, off, verts
static int forward_all(const ihmr_model* m, int B, const float* params, OptWs& w, cudaStream_t st,
                       IterProf* prof = nullptr) {
    HandSrc src;
    src.params = params;
    int rc;
    IHMR_TICK(prof, 0);
    if ((rc = launch_pose_prep(m, 2 * B, src, w.mano.X, w.mano.A, w.joints, st))) return rc;
    IHMR_TICK(prof, 1);
    if ((rc = launch_blend_fwd(m, 2 * B, w.mano.X, w.mano.off, st))) return rc;
    IHMR_TICK(prof, 2);
    if ((rc = launch_skin_fwd(m, 2 * B, w.mano.off, w.mano.A, w.verts, st, w.bbox))) return rc;
    IHMR_TICK(prof, 3);
    return IHMR_OK;
}

static FrameLossArgs base_loss_args(int B, int bs_norm, const float* params, const ihmr_targets_t* tg,
                                    const ihmr_stage_t* stg, OptWs& w) {
    FrameLossArgs a{};
    a.B = B;
    a.inv_n = 1.0f / (float)bs_norm;
    a.params = params;
    a.joints16 = w.joints;
    a.verts = w.verts;
    a.tg = *tg;
    a.w2d = stg->w_joints_2d; a.w3d = stg->w_joints_3d; a.wtrans = stg->w_trans;
    a.wshape = stg->w_shape_reg; a.wfinger = stg->w_finger_reg;
    a.col_loss = w.col_loss;
    return a;
}

// What an iteration has to recompute, given which parameter groups the stage updates
// (SURVEY.md Appendix C).  Skipped kernels would reproduce bit-identical buffers.
struct IterPlan {
    bool mano_fwd = true;     // pose_prep + skinning: orient / pose / shape live (or first iteration)
    bool blend_fwd = true;    // blend contraction: pose / shape live (or first iteration)
    bool mano_bwd = true;     // skinning + chain backward: orient / pose / shape live
    bool blend_bwd = true;    // blend contraction backward: pose / shape live
    int sdf_skip_grid = 0;    // stage that only moves the left hand rigidly: the right hand's samples
                              // carry no gradient, their loss part is needed at snapshots only
    int rigid = 0;            // orientation-only stage: 1 = first iteration (generic forward, then cache the
                              // root-local geometry), 2 = later iterations (x = R0 L + J0); backward is rigid in both
    bool static_right = false;  // the stage moves only hand_trans: the right hand (grid of the live direction) never changes
    bool dense_grad = false;  // (with IHMR_STAGE_GENERIC_KERNELS) no fingertip-only short path for hands without collision gradient
    int shape = 0;            // shape-only stage: 1 = first iteration (generic forward, then cache T_v | T_v c_v),
                              // 2 = later iterations (affine in beta); backward is the affine one in both
};

// Which kernels an iteration needs, from the parameter groups the stage updates.
// Dependency on the oracle's reading of the un-vendored `sdf` package (unpinned): the skipped
// direction below (sdf_skip_grid = 2 when only pred_hand_trans is live and no snapshot is due) is the one whose grid
// hand is the LEFT hand; its samples are the right hand's vertices, which do not move, and its box centre / scale are
// built under no_grad (A2) like the field itself (A5), so it carries no gradient to pred_hand_trans and only its
// VALUE is needed, at snapshots.  If the real package lets gradients through the box centre or scale, the trans-stage
// gradient differs: set IHMR_STAGE_GENERIC_KERNELS in ihmr_stage_t.flags (every stage on the generic chain, both
// directions every iteration) and extend the penetration backward accordingly.
static IterPlan plan_iteration(uint32_t mask, bool first, bool snapshot, bool generic) {
    const bool live_blend = mask & (IHMR_P_R_POSE | IHMR_P_L_POSE | IHMR_P_R_SHAPE | IHMR_P_L_SHAPE);
    const bool live_mano = live_blend || (mask & (IHMR_P_R_ORIENT | IHMR_P_L_ORIENT));
    IterPlan p;
    p.mano_fwd = first || live_mano;
    p.blend_fwd = first || live_blend;
    p.mano_bwd = live_mano;
    p.blend_bwd = live_blend;
    p.sdf_skip_grid = (!live_mano && !snapshot && !generic) ? 2 : 0;
    p.dense_grad = generic;
    p.static_right = !live_mano && !generic;
    if (generic) return p;      // IHMR_STAGE_GENERIC_KERNELS: every stage on the generic kernel chain, every hand dense
    if ((mask & (IHMR_P_R_SHAPE | IHMR_P_L_SHAPE)) &&
        !(mask & (IHMR_P_R_POSE | IHMR_P_L_POSE | IHMR_P_R_ORIENT | IHMR_P_L_ORIENT))) {   // opt_default stage 3
        p.shape = first ? 1 : 2;
        p.blend_fwd = first;
        p.blend_bwd = false;
    }
    if (live_mano && !live_blend) {          // only the global orientations move (opt_default stage 1)
        p.rigid = first ? 1 : 2;
        p.mano_fwd = first;
        p.mano_bwd = false;
    }
    return p;
}

// value + gradient of one iteration into w.grad (and optionally the six batch losses)
static int value_and_grad(const ihmr_model* m, int B, int bs_norm, const float* params, const ihmr_targets_t* tg,
                          const ihmr_stage_t* stg, OptWs& w, FrameLossArgs& la, cudaStream_t st,
                          IterProf* prof = nullptr, IterPlan plan = IterPlan(), bool carry_hints = false) {
    int rc;
    HandSrc src;
    src.params = params;
    IHMR_TICK(prof, 0);
    if (plan.mano_fwd && (rc = launch_pose_prep(m, 2 * B, src, w.mano.X, w.mano.A, w.joints, st))) return rc;
    IHMR_TICK(prof, 1);
    if (plan.blend_fwd && (rc = launch_blend_fwd(m, 2 * B, w.mano.X, w.mano.off, st))) return rc;
    IHMR_TICK(prof, 2);
    if (plan.mano_fwd && plan.shape != 2 && (rc = launch_skin_fwd(m, 2 * B, w.mano.off, w.mano.A, w.verts, st, w.bbox))) return rc;
    // shape-only stage: pose_prep above refreshed the joints and the translation columns of A
    if (plan.shape == 1 && (rc = launch_shape_prep(m, 2 * B, src, w.mano.off, w.mano.A, w.shape_cache, st))) return rc;
    if (plan.shape == 2 && (rc = launch_shape_fwd(m, 2 * B, src, w.mano.A, w.shape_cache, w.verts, st, w.bbox))) return rc;
    // orientation-only stage: the root-local geometry lives in the (otherwise idle) gposed / dA buffers
    if (plan.rigid == 1 && (rc = launch_rigid_prep(2 * B, src, w.verts, w.joints, w.mano.gposed, w.mano.dA, st))) return rc;
    if (plan.rigid == 2 && (rc = launch_rigid_fwd(2 * B, src, w.verts, w.joints, w.mano.gposed, w.mano.dA, st, w.bbox))) return rc;
    IHMR_TICK(prof, 3);
    SdfArgs sa;
    sa.verts = w.verts; sa.joints = w.joints; sa.params = params; sa.hand_type = tg->hand_type_array;
    sa.bbox = w.bbox;         // every writer of w.verts leaves the hands' boxes there (iterations that keep the vertices keep them)
    sa.ws = w.sdf_ws; sa.hints = carry_hints ? w.sdf_hints : nullptr;
    if (carry_hints && plan.static_right) { sa.static_grid_mask = 1; sa.pcache = w.sdf_pcache; sa.phic = w.sdf_phic; } sa.gverts = (plan.mano_bwd || plan.rigid) ? w.gverts : nullptr; sa.gshift = w.gshift;
    // generic backward chain: hands without collision gradient take the fingertip-only path (flags from the sdf kernels,
    // list of the others from the loss kernel)
    SparseGrad sp;
    if (sa.gverts && !plan.dense_grad) {
        sp.gzero = w.gzero;
        sa.gzero = w.gzero;
        // the dense skinning backward and the blend contraction's backward run on the list of the other hands
        sp.dense_list = w.dense_list; sp.dense_count = w.dense_list + 2 * B;
        IHMR_CUDA_OK(cudaMemsetAsync(sp.dense_count, 0, 4, st));
    }
    sa.grad_scale = stg->w_collision / (float)bs_norm;
    sa.skip_grid_mask = plan.sdf_skip_grid;
    if ((rc = launch_sdf(m, B, sa, st))) return rc;
    IHMR_TICK(prof, 4);
    la.gshift_col = w.gshift;
    la.gzero = sp.gzero; la.dense_list = sp.dense_list; la.dense_count = sp.dense_count;
    la.col_loss = nullptr; la.col_parts = sdf_ws_parts(w.sdf_ws, B);
    la.gjoints16 = w.gjoints; la.gtips = w.gtips; la.grad = w.grad;
    la.j2d_batch = w.j2d_b; la.j3d_batch = w.j3d_b;
    k_frame_loss<<<B, FL_THREADS, 0, st>>>(la);
    IHMR_LAUNCH_OK();
    IHMR_TICK(prof, 5);
    if (plan.mano_bwd && !plan.shape && (rc = launch_skin_bwd(m, 2 * B, w.mano.off, w.mano.A, w.gverts, w.gtips, w.mano.gposed, w.mano.dA, st, sp, w.mano.dX))) return rc;
    if (plan.shape && (rc = launch_shape_bwd(m, 2 * B, w.shape_cache, w.gverts, w.gtips, w.mano.dA, w.mano.dX, st, sp))) return rc;
    IHMR_TICK(prof, 6);
    if (plan.blend_bwd && (rc = launch_blend_bwd(m, 2 * B, w.mano.gposed, w.mano.dX, st, sp, w.mano.off))) return rc;
    IHMR_TICK(prof, 7);
    if (plan.rigid && (rc = launch_rigid_bwd(2 * B, src, w.gverts, w.gtips, w.gjoints, w.mano.gposed, w.mano.dA, w.grad, st, sp))) return rc;
    HandGrad hg;
    hg.params_grad = w.grad;
    if (plan.mano_bwd && (rc = launch_pose_bwd(m, 2 * B, src, w.mano.dA, w.gjoints, (plan.blend_bwd || plan.shape) ? w.mano.dX : nullptr, hg, st))) return rc;
    IHMR_TICK(prof, 8);
    return IHMR_OK;
}

int opt_stage(const ihmr_model* m, int B, int bs_norm, float* params, const ihmr_targets_t* tg,
              const ihmr_stage_t* stg, int save_mid_freq, int optimizer, void* ws, cudaStream_t st) {
    NvtxRange range("ihmr_opt_stage");
    OptWs w;
    opt_ws_layout(ws, B, &w);
    IHMR_CUDA_OK(cudaMemsetAsync(w.m, 0, (size_t)B * PD * 4, st));
    IHMR_CUDA_OK(cudaMemsetAsync(w.v, 0, (size_t)B * PD * 4, st));
    IHMR_CUDA_OK(cudaMemsetAsync(w.sdf_hints, 0, sdf_hint_bytes(B), st));
    IHMR_CUDA_OK(cudaMemsetAsync(w.sdf_pcache, 0, sdf_pcache_bytes(B), st));
    const int nthr = 256, nblk = (int)(((size_t)B * PD + nthr - 1) / nthr);
    int snaps = 0;
    for (int j = 0; j <= stg->epoch; ++j) {
        const bool snap = (j % save_mid_freq) == 0;
        FrameLossArgs la = base_loss_args(B, bs_norm, params, tg, stg, w);
        if (snap) {
            la.snap_mode = snaps == 0 ? 1 : 2;
            la.n_filters = stg->n_filters;
            for (int f = 0; f < stg->n_filters; ++f) {
                la.filter_loss[f] = stg->filter_loss[f];
                la.filter_factor[f] = filter_factor_of(stg->filter_percent[f]);
            }
            la.select_loss = stg->select_loss;
            la.origin = w.origin; la.best = w.best; la.take = w.take;
            ++snaps;
        }
        int rc = value_and_grad(m, B, bs_norm, params, tg, stg, w, la, st, nullptr, plan_iteration(stg->update_mask, j == 0, snap, stg->flags & IHMR_STAGE_GENERIC_KERNELS), true);
        if (rc) return rc;
        StepArgs sa{};
        sa.B = B; sa.mask = stg->update_mask; sa.optimizer = optimizer; sa.lr = stg->lr;
        const double t = (double)(j + 1);
        sa.step_size = (float)((double)stg->lr / (1.0 - pow(0.9, t)));
        sa.bc2_sqrt = (float)sqrt(1.0 - pow(0.999, t));
        sa.eps = 1e-8f; sa.momentum = 0.9f; sa.first_step = (j == 0);
        sa.params = params; sa.grad = w.grad; sa.m = w.m; sa.v = w.v;
        sa.take = snap ? w.take : nullptr; sa.best_params = w.best_params;
        k_step<<<nblk, nthr, 0, st>>>(sa);
        IHMR_LAUNCH_OK();
    }
    k_restore<<<nblk, nthr, 0, st>>>(B, stg->update_mask, params, w.best_params);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

// One iteration (forward, losses, backward, optimiser step on a scratch copy of nothing: the
// parameters ARE updated) with an event after every kernel class; synchronises the stream.
int opt_profile_iteration(const ihmr_model* m, int B, int bs_norm, float* params, const ihmr_targets_t* tg,
                          const ihmr_stage_t* stg, float* ms, void* ws, cudaStream_t st) {
    OptWs w;
    opt_ws_layout(ws, B, &w);
    IterProf prof;
    prof.st = st;
    for (auto& e : prof.ev) IHMR_CUDA_OK(cudaEventCreate(&e));
    FrameLossArgs la = base_loss_args(B, bs_norm, params, tg, stg, w);
    // one untimed full iteration fills every cached buffer, then the steady-state iteration is timed
    IHMR_CUDA_OK(cudaMemsetAsync(w.sdf_hints, 0, sdf_hint_bytes(B), st));
    IHMR_CUDA_OK(cudaMemsetAsync(w.sdf_pcache, 0, sdf_pcache_bytes(B), st));
    int rc = value_and_grad(m, B, bs_norm, params, tg, stg, w, la, st, nullptr, plan_iteration(stg->update_mask, true, true, stg->flags & IHMR_STAGE_GENERIC_KERNELS), true);
    if (rc) return rc;
    rc = value_and_grad(m, B, bs_norm, params, tg, stg, w, la, st, &prof, plan_iteration(stg->update_mask, false, false, stg->flags & IHMR_STAGE_GENERIC_KERNELS), true);
    if (rc) return rc;
    StepArgs sa{};
    sa.B = B; sa.mask = stg->update_mask; sa.optimizer = IHMR_OPT_ADAM; sa.lr = 0.f;   // lr 0: parameters unchanged
    sa.step_size = 0.f; sa.bc2_sqrt = 1.f; sa.eps = 1e-8f; sa.momentum = 0.9f; sa.first_step = 1;
    sa.params = params; sa.grad = w.grad; sa.m = w.m; sa.v = w.v; sa.take = nullptr; sa.best_params = w.best_params;
    const int nthr = 256, nblk = (int)(((size_t)B * PD + nthr - 1) / nthr);
    k_step<<<nblk, nthr, 0, st>>>(sa);
    IHMR_LAUNCH_OK();
    prof.tick(9);
    IHMR_CUDA_OK(cudaStreamSynchronize(st));
    for (int i = 0; i < N_KERNEL_CLASSES; ++i) IHMR_CUDA_OK(cudaEventElapsedTime(&ms[i], prof.ev[i], prof.ev[i + 1]));
    return IHMR_OK;
}

int opt_value_and_grad(const ihmr_model* m, int B, int bs_norm, const float* params, const ihmr_targets_t* tg,
                       const ihmr_stage_t* stg, float* losses6, float* grad, void* ws, cudaStream_t st) {
    OptWs w;
    opt_ws_layout(ws, B, &w);
    FrameLossArgs la = base_loss_args(B, bs_norm, params, tg, stg, w);
    la.loss_parts = w.loss_parts;
    int rc = value_and_grad(m, B, bs_norm, params, tg, stg, w, la, st);
    if (rc) return rc;
    if (losses6) {
        k_loss_reduce<<<6, 256, 0, st>>>(B, w.loss_parts, stg->w_collision, losses6);
        IHMR_LAUNCH_OK();
    }
    if (grad) IHMR_CUDA_OK(cudaMemcpyAsync(grad, w.grad, (size_t)B * PD * 4, cudaMemcpyDeviceToDevice, st));
    return IHMR_OK;
}

// forward + the three per-frame selection criteria, nothing else (IHMR-MLP inference, mlp_model.py:514-583 as far as
// select_better_params reads it): joints_3d_loss_p and joints_2d_loss_p carry their weights, collision_loss weight 1
int opt_criteria(const ihmr_model* m, int B, const float* params, const ihmr_targets_t* tg, float w_joints_2d, float w_joints_3d,
                 float* criteria, void* ws, cudaStream_t st) {
    OptWs w;
    opt_ws_layout(ws, B, &w);
    int rc;
    if ((rc = forward_all(m, B, params, w, st))) return rc;
    SdfArgs sa;
    sa.verts = w.verts; sa.joints = w.joints; sa.params = params; sa.hand_type = tg->hand_type_array;
    sa.losses = w.col_loss; sa.ws = w.sdf_ws; sa.bbox = w.bbox;
    if ((rc = launch_sdf(m, B, sa, st))) return rc;
    ihmr_stage_t wts{};
    wts.w_joints_2d = w_joints_2d; wts.w_joints_3d = w_joints_3d;
    FrameLossArgs la = base_loss_args(B, B, params, tg, &wts, w);
    la.col_loss = w.col_loss;
    la.crit3 = criteria;
    k_frame_loss<<<B, FL_THREADS, 0, st>>>(la);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int opt_final(const ihmr_model* m, int B, const float* params, const ihmr_targets_t* tg, float* right_verts,
              float* left_verts, float* joints_3d, float* collision_loss, float* collision_origin,
              float* j3d_loss_p, void* ws, cudaStream_t st) {
    NvtxRange range("ihmr_opt_final");
    OptWs w;
    opt_ws_layout(ws, B, &w);
    int rc;
    if ((rc = forward_all(m, B, params, w, st))) return rc;
    SdfArgs sa;
    sa.verts = w.verts; sa.joints = w.joints; sa.params = params; sa.hand_type = tg->hand_type_array;
    sa.losses = collision_loss ? collision_loss : w.col_loss;
    sa.origin = collision_origin; sa.ws = w.sdf_ws; sa.bbox = w.bbox;
    if ((rc = launch_sdf(m, B, sa, st))) return rc;
    ihmr_stage_t dflt{};   // default_loss_weights (optimize_model.py:84-92)
    dflt.w_joints_2d = 10.f; dflt.w_joints_3d = 1000.f; dflt.w_trans = 100.f; dflt.w_shape_reg = 0.1f;
    dflt.w_collision = 1.f; dflt.w_finger_reg = 100000.f;
    FrameLossArgs la = base_loss_args(B, B, params, tg, &dflt, w);
    la.col_loss = sa.losses;
    la.joints_out = joints_3d;
    la.j3d_batch = j3d_loss_p;
    k_frame_loss<<<B, FL_THREADS, 0, st>>>(la);
    IHMR_LAUNCH_OK();
    if (right_verts && left_verts) {
        dim3 grid(B, (NV + 255) / 256);
        k_export_verts<<<grid, 256, 0, st>>>(B, w.verts, w.joints, params, right_verts, left_verts);
        IHMR_LAUNCH_OK();
    }
    return IHMR_OK;
}

}  // namespace ihmr
