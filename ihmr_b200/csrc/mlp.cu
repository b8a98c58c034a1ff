// IHMR-MLP inference pieces: the per-stage residual MLPs of
// /root/reference/src/models/networks.py:83-105 (InterHandSubNetwork: 1146 -> 512 -> 256 -> 128 -> update dim) and the
// parameter bookkeeping of /root/reference/src/models/mlp_model.py:458-472 (__update_params_single) and :592-637
// (select_better_params).  The matrix products run on the tcgen05 contraction of blend_tc.cu (3xTF32, fp32 accuracy);
// the MANO forward and the criteria come from the refinement path's own final-forward call.
#include "kernels.cuh"

namespace ihmr {

constexpr int MLP_FEAT = 1024;
constexpr int MLP_IN = 1152;        // 1024 image features + 122 parameters, padded to a multiple of 32

// x = [img_feat | cam 3 | pose 96 | shape 20 | hand_trans 3 | 0 ...]: the reference's final_params order
// (mlp_model.py:432-436) from the (B,122) layout of this library [cam | trans | pose | shape]
__global__ void k_mlp_input(int n, const float* __restrict__ feat, const float* __restrict__ params, float* __restrict__ x) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * MLP_IN) return;
    const int b = (int)(i / MLP_IN), c = (int)(i % MLP_IN);
    float v = 0.f;
    if (c < MLP_FEAT) v = feat[(size_t)b * MLP_FEAT + c];
    else if (c < MLP_FEAT + PD) {
        const int k = c - MLP_FEAT;                      // index into final_params
        const int src = k < 3 ? P_CAM + k : k < 99 ? P_POSE + (k - 3) : k < 119 ? P_SHAPE + (k - 99) : P_TRANS + (k - 119);
        v = params[(size_t)b * PD + src];
    }
    x[i] = v;
}

// y = act(y + bias) in place on the first out_dim columns of every row
__global__ void k_bias_act(int n, int out_dim, int ldy, const float* __restrict__ bias, int relu, float* __restrict__ y) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * out_dim) return;
    const int b = (int)(i / out_dim), c = (int)(i % out_dim);
    float v = y[(size_t)b * ldy + c] + (bias ? bias[c] : 0.f);
    y[(size_t)b * ldy + c] = relu ? fmaxf(v, 0.f) : v;
}

struct Segs {
    int n;
    int col[8], len[8];
};

// params_out = params_in, plus the residual on the listed column segments (residual columns in list order)
__global__ void k_mlp_apply(int n, const float* __restrict__ res, int ldr, Segs sg, const float* __restrict__ pin, float* __restrict__ pout) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * PD) return;
    const int b = (int)(i / PD), c = (int)(i % PD);
    float v = pin[i];
    int roff = 0;
    for (int s = 0; s < sg.n; ++s) {
        if (c >= sg.col[s] && c < sg.col[s] + sg.len[s]) v += res[(size_t)b * ldr + roff + (c - sg.col[s])];
        roff += sg.len[s];
    }
    pout[i] = v;
}

// select_better_params: the new parameters are kept where every filter criterion improved strictly within its
// margin (cur < prev * (1 + percent / 100)) and the select criterion did not get worse (cur <= prev); elsewhere the
// previous parameters and criteria stay.  crit: (n,3) = [joints_3d_loss_p, collision_loss, joints_2d_loss_p]
__global__ void k_select_better(int n, const float* __restrict__ cur, float* __restrict__ prev, int n_filters, int f0, int f1,
                                int f2, int f3, float p0, float p1, float p2, float p3, int sel, uint32_t mask,
                                const float* __restrict__ pnew, float* __restrict__ params, int* __restrict__ kept) {
    const int b = blockIdx.x, t = threadIdx.x;
    const int fl[4] = {f0, f1, f2, f3};
    const float ff[4] = {p0, p1, p2, p3};
    bool ok = true;
    for (int f = 0; f < n_filters; ++f) ok = ok && (cur[b * 3 + fl[f]] < prev[b * 3 + fl[f]] * ff[f]);
    ok = ok && (cur[b * 3 + sel] <= prev[b * 3 + sel]);
    __syncthreads();                       // every thread has read prev before thread 0 updates it
    if (ok) {
        for (int c = t; c < PD; c += blockDim.x) {
            const uint32_t g = c < 3 ? IHMR_P_CAM : c < 6 ? IHMR_P_TRANS : c < 9 ? IHMR_P_R_ORIENT : c < 54 ? IHMR_P_R_POSE
                             : c < 57 ? IHMR_P_L_ORIENT : c < 102 ? IHMR_P_L_POSE : c < 112 ? IHMR_P_R_SHAPE : IHMR_P_L_SHAPE;
            if (g & mask) params[(size_t)b * PD + c] = pnew[(size_t)b * PD + c];
        }
        if (t < 3) prev[b * 3 + t] = cur[b * 3 + t];
    }
    if (kept && t == 0) kept[b] = ok ? 1 : 0;
}

int launch_mlp_input(int n, const float* feat, const float* params, float* x, cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    const size_t tot = (size_t)n * MLP_IN;
    k_mlp_input<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, feat, params, x);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_linear(int n, int in_dim, int out_dim, const float* x, int ldx, const float* W, const float* bias, int relu,
                  float* y, int ldy, cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    // y (n x out_pad) = x (n x in_dim) . W^T with W (out_pad x in_dim) K-major: the nn.Linear weight layout
    const int out_pad = (out_dim + 3) & ~3;
    if (int rc = launch_gemm_tf32x3(n, out_pad, in_dim, x, ldx, W, in_dim, y, ldy, st)) return rc;
    const size_t tot = (size_t)n * out_dim;
    k_bias_act<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, out_dim, ldy, bias, relu, y);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_mlp_apply(int n, const float* res, int ldr, int n_seg, const int* col, const int* len, const float* pin, float* pout,
                     cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    Segs sg{};
    sg.n = n_seg;
    for (int i = 0; i < n_seg; ++i) { sg.col[i] = col[i]; sg.len[i] = len[i]; }
    const size_t tot = (size_t)n * PD;
    k_mlp_apply<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, res, ldr, sg, pin, pout);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

int launch_select_better(int n, const float* cur, float* prev, const ihmr_stage_t* stg, const float* pnew, float* params, int* kept,
                         cudaStream_t st) {
    if (n <= 0) return IHMR_OK;
    float ff[4] = {1.f, 1.f, 1.f, 1.f};
    int fl[4] = {0, 0, 0, 0};
    // idxs0 = cur_loss < prev_loss * (1 + float(percent) / 100)                      (mlp_model.py:601)
    for (int f = 0; f < stg->n_filters; ++f) { ff[f] = (float)(1.0 + (double)stg->filter_percent[f] / 100.0); fl[f] = stg->filter_loss[f]; }
    k_select_better<<<n, 128, 0, st>>>(n, cur, prev, stg->n_filters, fl[0], fl[1], fl[2], fl[3], ff[0], ff[1], ff[2], ff[3],
                                       stg->select_loss, stg->update_mask, pnew, params, kept);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

}  // namespace ihmr
