// Evaluator metrics on the device: the per-frame reductions the reference
// computes on the host after a 26 KB/frame device-to-host copy —
//   j3d_error                 utils/metric_utils.py:23-38  (get_single_joints_error), evaluator.py:74-81
//   pa_no_rot_inter_j3d_error utils/metric_utils.py:107-143 (calc_transform_no_rot + get_single_pa_inter_joints_error)
//   collision mean / max      utils/evaluator.py:163-181
// One warp per frame; the output is 6 floats per frame: [sum of joint errors, their count, sum of the
// no-rotation Procrustes errors, their count, mean and max of collision_loss_origin_scale].
#include "kernels.cuh"

namespace ihmr {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(128) k_eval_metrics(int B, const float* __restrict__ pred, const float* __restrict__ gt,
                                                     const float* __restrict__ origin, const float* __restrict__ scale,
                                                     float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* P = pred + (size_t)b * 126;
    const float* Gt = gt + (size_t)b * 168;
    const float sc = scale ? scale[b] : 1.0f;
    // each lane owns joints `lane` and `lane + 32` (< 42)
    float p[2][3], g[2][3], w[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int k = lane + 32 * t;
        const bool ok = k < 42;
#pragma unroll
        for (int c = 0; c < 3; ++c) { p[t][c] = ok ? P[k * 3 + c] : 0.f; g[t][c] = ok ? Gt[k * 4 + c] : 0.f; }
        w[t] = ok ? Gt[k * 4 + 3] : 0.f;
    }
    const float w0 = __shfl_sync(0xffffffffu, w[0], 0), w21 = __shfl_sync(0xffffffffu, w[0], 21);

    // ---- per-hand root-relative joint error (the reference subtracts in place, cumulatively)
    float q[2][3], h[2][3];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < 3; ++c) { q[t][c] = p[t][c]; h[t][c] = g[t][c]; }
    float esum = 0.f, ecnt = 0.f;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int root = pass ? 21 : 0;
        const float wr = pass ? w21 : w0;
        float rp[3], rg[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { rp[c] = __shfl_sync(0xffffffffu, q[0][c], root); rg[c] = __shfl_sync(0xffffffffu, h[0][c], root); }
        if (wr > 0.f) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { q[t][c] -= rp[c]; h[t][c] -= rg[c]; }
                const int k = lane + 32 * t;
                if (k >= root && k < root + 21 && w[t] > 0.f) {
                    const float dx = q[t][0] - h[t][0], dy = q[t][1] - h[t][1], dz = q[t][2] - h[t][2];
                    esum += sqrtf(dx * dx + dy * dy + dz * dz) / sc;
                    ecnt += 1.f;
                }
            }
        }
    }
    esum = wsum(esum); ecnt = wsum(ecnt);

    // ---- no-rotation Procrustes: per-axis mean / population std over the valid joints
    const float v0 = w[0] > 0.f ? 1.f : 0.f, v1 = (lane + 32 < 42 && w[1] > 0.f) ? 1.f : 0.f;
    const float nvalid = wsum(v0 + v1);
    const float wtot = wsum(w[0] + w[1]);                    // the reference tests sum(joints_valid) < 2
    float psum = 0.f, pcnt = 0.f;
    if (wtot >= 2.0f) {
        float d2[2] = {0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float m1 = wsum(v0 * p[0][c] + v1 * p[1][c]) / nvalid, m2 = wsum(v0 * g[0][c] + v1 * g[1][c]) / nvalid;
            const float a0 = p[0][c] - m1, a1 = p[1][c] - m1, b0 = g[0][c] - m2, b1 = g[1][c] - m2;
            const float s1 = sqrtf(wsum(v0 * a0 * a0 + v1 * a1 * a1) / nvalid), s2 = sqrtf(wsum(v0 * b0 * b0 + v1 * b1 * b1) / nvalid);
            const float t0 = a0 / s1 * s2 + m2 - g[0][c], t1 = a1 / s1 * s2 + m2 - g[1][c];
            d2[0] += t0 * t0; d2[1] += t1 * t1;
        }
        psum = wsum(v0 * sqrtf(d2[0]) / sc + v1 * sqrtf(d2[1]) / sc);
        pcnt = nvalid;
    }
    // ---- collision statistics over the 1556 per-vertex values
    float cs = 0.f, cm = -1e30f;
    const float* O = origin + (size_t)b * 1556;
    for (int i = lane; i < 1556; i += 32) { const float x = O[i]; cs += x; cm = fmaxf(cm, x); }
    cs = wsum(cs);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if (lane == 0) {
        float* r = out + (size_t)b * 6;
        r[0] = esum; r[1] = ecnt; r[2] = psum; r[3] = pcnt; r[4] = cs / 1556.0f; r[5] = cm;
    }
}

int launch_eval_metrics(int B, const float* pred, const float* gt, const float* origin, const float* scale, float* out,
                        cudaStream_t st) {
    if (B <= 0) return IHMR_OK;
    k_eval_metrics<<<(B + 3) / 4, 128, 0, st>>>(B, pred, gt, origin, scale, out);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

// ---- FP32 FMA peak: 8 independent chains per thread, nothing but FFMA in the loop (measurement aid)
__global__ void __launch_bounds__(256) k_ffma_peak(int iters, float* out) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    const float m = 1.0000001f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;       // keeps the chains alive
}

int measure_fp32_peak(int num_sms, float* tflops, float* scratch, cudaStream_t st) {
    const int iters = 4096, blocks = num_sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    IHMR_CUDA_OK(cudaEventCreate(&e0));
    IHMR_CUDA_OK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        k_ffma_peak<<<blocks, threads, 0, st>>>(iters, scratch);
        cudaEventRecord(e1, st);
        cudaError_t e = cudaStreamSynchronize(st);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            set_error("fp32 peak measurement failed: %s", cudaGetErrorString(e));
            return IHMR_E_CUDA;
        }
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    count_launch();
    *tflops = 2.0f * 8 * 16 * (float)iters * blocks * threads / (best * 1e-3f) / 1e12f;
    return IHMR_OK;
}

}  // namespace ihmr
