// Left/right interpenetration loss, forward + backward in one kernel (SURVEY.md §8 a10,
// Appendix B).  One CTA per frame.  For each direction (grid hand h, query hand o = 1-h):
//
//   1. bounding box of h -> centre c_h, scale s_h = 0.6 * max extent                     (A2)
//   2. mark the <= 8 voxel corners each query vertex touches                      (lazy grid)
//   3. inside/outside of the marked voxel columns: every face is rasterised onto the 32x32
//      (y,z) lattice of +x rays; a hit toggles the bits of all voxels left of the crossing (A4)
//   4. phi = min point-triangle distance for every voxel that is both marked and inside:
//      face-centric — each face enumerates the marked voxels within R of its bounding box,
//      (voxel, face) candidates are queued and the exact tests run densely, one per thread;
//      voxels whose nearest face is beyond R are finished by a warp-wide search over static
//      face clusters (nearest bounding box first)
//   5. trilinear sampling with grid_sample(align_corners=False, zeros) semantics, its
//      gradient w.r.t. the query vertex, per-frame loss = sum / 4                 (A3, A5, A6)
//
// The voxel values are exactly those of the brute-force 32^3 grid of the reference kernel
// (`sdf_cuda`, reached from /root/reference/src/models/loss_utils.py:181): integer crossing
// counts and `min` are order independent, and voxels that are not marked never contribute.
// The inside test uses the arithmetic contract of oracle/sdf_oracle.c (no FMA contraction,
// edge functions on (low id, high id) ordering) so both make the same decisions.
#include "kernels.cuh"

namespace ihmr {

constexpr int G = 32;
#ifndef SDF_NTHREADS
#define SDF_NTHREADS 512             // measured: 512 x 2 CTAs/SM beats 256 x 3 and 256 x 4
#endif
constexpr int SDF_THREADS = SDF_NTHREADS;
constexpr int SDF_WARPS = SDF_THREADS / 32;
constexpr int SDF_CPT = G * G / SDF_THREADS;   // (z,y) columns per thread in the scans
#ifndef SDF_MIN_CTAS
#define SDF_MIN_CTAS 2               // resident CTAs per SM the register allocation is held to
#endif
constexpr int SDF_SLOTS = (NV + SDF_THREADS - 1) / SDF_THREADS;   // query vertices per thread
constexpr int NCL = (NF + 31) / 32; // static face clusters of <= 32 faces (49)
#if SDF_MIN_CTAS >= 4               // <= 56 KB of shared memory per CTA
constexpr int PHI_CAP = 1024;       // voxels evaluated per pass (more voxels: further passes)
constexpr int Q_CAP = 2048;         // queued candidates (overflow is processed in place)
constexpr int P_CAP = 1024;         // queued (voxel, cluster) pairs of one chunk of voxels
constexpr int V_CHUNK = 224;        // voxels per (A) round: ~3-5 pairs each in the inner band, so P_CAP rarely overflows
constexpr int P_CHUNK = 160;        // pairs per (B) round: ~1/3 of their 32 faces pass, so Q_CAP rarely overflows
#elif SDF_MIN_CTAS == 3             // <= 75 KB
constexpr int PHI_CAP = 2048;
constexpr int Q_CAP = 4096;
constexpr int P_CAP = 1536;
constexpr int V_CHUNK = 320;
constexpr int P_CHUNK = 320;
#else                               // <= 112 KB
constexpr int PHI_CAP = 2048;
constexpr int Q_CAP = 6144;
constexpr int P_CAP = 3072;
constexpr int V_CHUNK = 640;
#ifndef SDF_P_CHUNK
#define SDF_P_CHUNK 480
#endif
constexpr int P_CHUNK = SDF_P_CHUNK;
#endif
#ifndef SDF_R_CELLS
#define SDF_R_CELLS 2.5f
#endif
// The nearest-face search works on boxes quantised to Q8 = 1/128 of the normalised cube (1/8 voxel):
// voxel centres are the integers 8i + 4, box distances are exact integers.
#ifndef SDF_SHELL
#define SDF_SHELL 0
#endif
#ifndef SDF_SPLIT
#define SDF_SPLIT 2                  // minimum number of (B)/(C) rounds of the inner band
#endif
#ifndef SDF_NBANDS
#define SDF_NBANDS 2                 // measured: 2 bands (clusters containing the voxel, then everything within R) beat 1 and 3
#endif
#ifndef SDF_T0
#define SDF_T0 8                     // squared radius of round 0 in Q8 units (0.35 voxel)
#endif
#ifndef SDF_T1
#define SDF_T1 64                    // round 1 (1 voxel); round 2 reaches R
#endif
constexpr int SDF_R_Q8 = (int)(SDF_R_CELLS * 8.0f);
constexpr int SDF_R2_Q8 = SDF_R_Q8 * SDF_R_Q8;
constexpr float Q8_TO_D2 = 1.0f / 16384.0f;         // Q8 units squared -> normalised units squared
constexpr float SDF_R2 = SDF_R2_Q8 * Q8_TO_D2;      // squared radius the rounds certify

struct __align__(16) SdfSmem {
    float V[2 * NV * 3];        // both hands as stored (one TMA bulk copy, 18,672 B)
    float U[NV * 3];            // normalised grid-hand vertices
    uint32_t needed[G * G];     // marked voxels per (z,y) column
    uint32_t work[G * G];       // parity bits, then marked & inside
    int region[4];              // lattice bounds of the marked columns: y min, y max, z min, z max
    uint16_t coloff[G * G];     // exclusive prefix of popc(work)
    uint32_t worklist[PHI_CAP]; // voxels of the current pass as packed Q8 centres: 8x+4 | (8y+4) << 8 | (8z+4) << 16
    uint16_t far_list[PHI_CAP]; // voxels whose nearest face is beyond SDF_R
    uint32_t best[PHI_CAP];     // bit pattern of the best squared distance (>= 0: orders like uint); then phi
    uint32_t queue[Q_CAP];      // (voxel index << 16) | slot of the face in the cluster table
    uint32_t pairs[P_CAP];      // (voxel index << 6) | cluster
    uint2 cl_box[NCL];          // union of the cluster's face boxes, same packing
    uint2 fbox[NCL * 32];       // per face (cluster-table order): box quantised outwards to Q8;
                                // .x = lo x | y << 8 | z << 16 | valid << 24, .y = hi x | y << 8 | z << 16
    float red[64];
    float box[2][2][3];         // [hand][lo/hi][xyz]
    int boxi[2][2][3];          // the same as order-preserving integers while it is being reduced
    float dirp[2][16];          // per grid hand: centre xyz, scale, tlo xyz, thi xyz, wlo xyz, whi xyz
    float shift[4];
    int scan_warp[SDF_WARPS];
    uint32_t qn[2];             // queue fill, double buffered by round
    uint32_t pn[3];             // pair fill, rotating by round (a round without pairs has no barrier of its own)
    int far_count;
    unsigned long long bar;     // mbarrier of the bulk copy
};

// (2i + 1 - G) / G; every intermediate is a small multiple of 1/G, so the fused form is exact too
__device__ __forceinline__ float voxel_center(int i) { return fmaf((float)i, 2.0f / G, (1.0f - G) / G); }

// ---- arithmetic contract shared with oracle/sdf_oracle.c (explicitly unfused) -----------
__device__ __forceinline__ bool edge_side(const float* P, int i0, int i1, float qy, float qz, float& w) {
    const bool fwd = i0 < i1;
    const float* lo = P + 3 * (fwd ? i0 : i1);
    const float* hi = P + 3 * (fwd ? i1 : i0);
    const float e = __fsub_rn(__fmul_rn(__fsub_rn(hi[1], lo[1]), __fsub_rn(qz, lo[2])),
                              __fmul_rn(__fsub_rn(hi[2], lo[2]), __fsub_rn(qy, lo[1])));
    w = fwd ? e : -e;
    return fwd ? (e >= 0.f) : (e < 0.f);
}

// true and x set when the +x ray through (qy,qz) pierces face (ia,ib,ic)
__device__ __forceinline__ bool ray_hit(const float* P, int ia, int ib, int ic, float qy, float qz, float& x) {
    float wa, wb, wc;
    const bool p0 = edge_side(P, ia, ib, qy, qz, wc);
    const bool p1 = edge_side(P, ib, ic, qy, qz, wa);
    const bool p2 = edge_side(P, ic, ia, qy, qz, wb);
    if (!((p0 && p1 && p2) || (!p0 && !p1 && !p2))) return false;
    const float sum = __fadd_rn(__fadd_rn(wa, wb), wc);
    if (sum == 0.f) return false;
    x = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, P[3 * ia]), __fmul_rn(wb, P[3 * ib])), __fmul_rn(wc, P[3 * ic])), sum);
    return true;
}

__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Squared distance point -> triangle, branch free (all lanes of a warp stay converged):
// min over the three edge segments and, when the projection falls inside, the plane distance.
// Same quantities as the closest-point-region formulation of oracle/sdf_oracle.c (d1..d6,
// va/vb/vc); agrees with it to rounding.  fminf/fmaxf drop NaNs, which makes zero-length edges
// and zero-area faces fall back to the remaining candidates.
__device__ __forceinline__ float pt_tri_dist2(const float* p, const float* a, const float* b, const float* c) {
    float ab[3], ac[3], ap[3], bp[3], cp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; bp[k] = p[k] - b[k]; cp[k] = p[k] - c[k]; }
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap), d3 = dot3(ab, bp), d4 = dot3(ac, bp), d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    // edge ab: t = d1 / |ab|^2, |ab|^2 = d1 - d3
    const float t1 = fminf(fmaxf(__fdividef(d1, d1 - d3), 0.f), 1.f);
    float e[3] = {ap[0] - t1 * ab[0], ap[1] - t1 * ab[1], ap[2] - t1 * ab[2]};
    float best = dot3(e, e);
    // edge ac: t = d2 / |ac|^2, |ac|^2 = d2 - d6
    const float t2 = fminf(fmaxf(__fdividef(d2, d2 - d6), 0.f), 1.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) e[k] = ap[k] - t2 * ac[k];
    best = fminf(best, dot3(e, e));
    // edge bc: t = (d4 - d3) / |bc|^2, |bc|^2 = (d4 - d3) + (d5 - d6)
    const float t3 = fminf(fmaxf(__fdividef(d4 - d3, (d4 - d3) + (d5 - d6)), 0.f), 1.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) e[k] = bp[k] - t3 * (ac[k] - ab[k]);
    best = fminf(best, dot3(e, e));
    // interior: barycentric coordinates all non-negative
    const float vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    const float rden = __fdividef(1.0f, va + vb + vc), v = vb * rden, w = vc * rden;
#pragma unroll
    for (int k = 0; k < 3; ++k) e[k] = ap[k] - v * ab[k] - w * ac[k];
    const float din = dot3(e, e);
    return (va >= 0.f && vb >= 0.f && vc >= 0.f) ? fminf(best, din) : best;
}

// floats <-> integers with the same ordering (an involution on the bit pattern)
__device__ __forceinline__ int float_ordered(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// ---- block primitives ---------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// deterministic block sum of up to 4 values per thread; result valid in every thread
__device__ __forceinline__ void block_sum4(float* v, int nval, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = 0; i < nval; ++i) {
        float s = warp_sum(v[i]);
        if (lane == 0) red[i * SDF_WARPS + warp] = s;
    }
    __syncthreads();
    for (int i = 0; i < nval; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < SDF_WARPS; ++w) s += red[i * SDF_WARPS + w];
        v[i] = s;
    }
    __syncthreads();
}

// exclusive scan of 1024 counts, SDF_CPT consecutive entries per thread; returns total
__device__ __forceinline__ int block_scan_1024(const int (&cnt)[SDF_CPT], int (&excl)[SDF_CPT], int* scan_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int local = 0;
#pragma unroll
    for (int i = 0; i < SDF_CPT; ++i) local += cnt[i];
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scan_warp[warp] = inc;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SDF_WARPS; ++w) {
        if (w < warp) base += scan_warp[w];
        total += scan_warp[w];
    }
    int run = base + inc - local;
#pragma unroll
    for (int i = 0; i < SDF_CPT; ++i) { excl[i] = run; run += cnt[i]; }
    __syncthreads();
    return total;
}

// Four times the squared distance, in Q8 units, from a packed voxel centre to a packed box, with byte-wise
// SIMD only: per axis 2 d = |q - lo| + |q - hi| - (hi - lo), and the sum of squares is expanded into
// dot products of the byte vectors (VABSDIFF4 + IDP.4A; no per-byte overflow).  Byte 3 is 0 in voxels and
// in valid boxes; boxes of empty table slots carry 255 there, which puts them beyond every radius.
__device__ __forceinline__ int qbox_4d2(uint2 bx, uint32_t q) {
    const uint32_t a = __vabsdiffu4(q, bx.x), b = __vabsdiffu4(q, bx.y), w = __vabsdiffu4(bx.y, bx.x);
    const uint32_t sq = __dp4a(w, w, __dp4a(b, b, __dp4a(a, a, 0u)));
    const uint32_t cr = __dp4a(b, w, __dp4a(a, w, 0u));
    return (int)(sq + 2u * (__dp4a(a, b, 0u) - cr));
}

__device__ __forceinline__ uint32_t pack_q8(int x, int y, int z) {
    return (uint32_t)(8 * x + 4) | ((uint32_t)(8 * y + 4) << 8) | ((uint32_t)(8 * z + 4) << 16);
}

// voxel centre (2i + 1 - G) / G = (8i + 4) / 128 - 1, exact either way
__device__ __forceinline__ void voxel_pos(uint32_t q8, float* q) {
    q[0] = fmaf((float)(q8 & 255u), 1.0f / 128.0f, -1.0f);
    q[1] = fmaf((float)((q8 >> 8) & 255u), 1.0f / 128.0f, -1.0f);
    q[2] = fmaf((float)((q8 >> 16) & 255u), 1.0f / 128.0f, -1.0f);
}

// one exact (voxel, face) test; the result lowers the voxel's best squared distance
__device__ __forceinline__ void pair_test(SdfSmem& s, const ushort4* __restrict__ cl_tri, int vi, int slot) {
    float q[3];
    voxel_pos(s.worklist[vi], q);
    const ushort4 id = cl_tri[slot];
    const float d2 = pt_tri_dist2(q, s.U + 3 * id.x, s.U + 3 * id.y, s.U + 3 * id.z);
    atomicMin(&s.best[vi], __float_as_uint(d2));
}

// Exact squared distance for a far voxel, one warp per voxel: the faces are grouped at model
// creation into NCL spatial clusters of <= 32 (cl_tri); clusters are visited nearest bounding
// box first, one face per lane, until the nearest unvisited box is farther than the best
// distance.  The open wrist makes some far-away voxels "inside" (odd crossing parity), and the
// reference's brute-force loop gives them their true distance, so they must be exact too.
__device__ __forceinline__ float eval_voxel_far(const SdfSmem& s, const ushort4* __restrict__ cl_tri, uint32_t q8,
                                                float best, int lane) {
    float q[3];
    voxel_pos(q8, q);
    float lb[2];                         // lower bounds: the quantised boxes contain the true ones
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int c = lane + 32 * t;
        lb[t] = 1e30f;
        if (c < NCL) lb[t] = (float)qbox_4d2(s.cl_box[c], q8) * (0.25f * Q8_TO_D2);
    }
    for (int it = 0; it < NCL; ++it) {
        float m = fminf(lb[0], lb[1]);
        int which = (lb[0] <= lb[1]) ? lane : lane + 32;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
            const int w2 = __shfl_xor_sync(0xffffffffu, which, o);
            if (m2 < m || (m2 == m && w2 < which)) { m = m2; which = w2; }
        }
        if (m > best * 1.00001f) break;
        const ushort4 id = cl_tri[which * 32 + lane];
        float d2 = 1e30f;
        if (id.w) d2 = pt_tri_dist2(q, s.U + 3 * id.x, s.U + 3 * id.y, s.U + 3 * id.z);
        best = fminf(best, warp_min(d2));
        if (which == lane) lb[0] = 1e30f;
        if (which == lane + 32) lb[1] = 1e30f;
    }
    return best;
}

#define SDF_STAT(i)                                                                       \
    if (a.stats && tid == 0) {                                                            \
        const long long t_now = clock64();                                                \
        a.stats[b * 32 + 8 + (i)] += (int)(t_now - t_prev);                               \
        t_prev = t_now;                                                                   \
    }

__global__ void __launch_bounds__(SDF_THREADS, SDF_MIN_CTAS)
k_sdf(int B, SdfArgs a, const uint16_t* __restrict__ faces_r, const uint16_t* __restrict__ faces_l,
      const ushort4* __restrict__ cl_r, const ushort4* __restrict__ cl_l) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SdfSmem& s = *reinterpret_cast<SdfSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const bool xform = (a.joints != nullptr);

    // Both hands of the frame are staged once by the TMA engine (one 1-D bulk copy completing on an
    // mbarrier); every later phase reads them from shared memory.
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s.bar);
    if (tid < 12) (&s.boxi[0][0][0])[tid] = ((tid / 3) & 1) ? (int)0x80000000 : 0x7fffffff;
    float mask = 1.0f;                   // both-hands flag; loaded here so that its latency hides behind the staging
    if (a.hand_type) mask = (a.hand_type[b * 2] + a.hand_type[b * 2 + 1] > 1.5f) ? 1.0f : 0.0f;
    if (tid == 0) {
        constexpr uint32_t V_BYTES = 2 * NV * 3 * sizeof(float);
        static_assert(V_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(V_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((uint32_t)__cvta_generic_to_shared(s.V)), "l"(a.verts + (size_t)b * (2 * NV * 3)), "r"(V_BYTES), "r"(bar) : "memory");
        float sh[3] = {0.f, 0.f, 0.f};
        if (xform) {
            const float* jr = a.joints + ((size_t)b * 2 + 0) * 48;
            const float* jl = a.joints + ((size_t)b * 2 + 1) * 48;
            const float* t = a.params + (size_t)b * PD + P_TRANS;
            sh[0] = t[0] + (jr[0] + jl[0]);      // - (-x)
            sh[1] = t[1] + (jr[1] - jl[1]);
            sh[2] = t[2] + (jr[2] - jl[2]);
        }
        s.shift[0] = sh[0]; s.shift[1] = sh[1]; s.shift[2] = sh[2];
    }
    __syncthreads();
    const float shx = s.shift[0], shy = s.shift[1], shz = s.shift[2];
    {
        uint32_t done;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        } while (!done);
    }
    auto load_vert = [&](int hand, int v, float* out) {
        const float* p = s.V + (hand * NV + v) * 3;
        out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
        if (xform && hand == 1) { out[0] = -out[0] + shx; out[1] += shy; out[2] += shz; }
    };
    long long t_prev = clock64();

    // ---- bounding boxes of both hands
    {
        float lo[2][3], hi[2][3];
#pragma unroll
        for (int hnd = 0; hnd < 2; ++hnd)
#pragma unroll
            for (int c = 0; c < 3; ++c) { lo[hnd][c] = 1e30f; hi[hnd][c] = -1e30f; }
        for (int v = tid; v < NV; v += SDF_THREADS) {
#pragma unroll
            for (int hnd = 0; hnd < 2; ++hnd) {
                float p[3];
                load_vert(hnd, v, p);
#pragma unroll
                for (int c = 0; c < 3; ++c) { lo[hnd][c] = fminf(lo[hnd][c], p[c]); hi[hnd][c] = fmaxf(hi[hnd][c], p[c]); }
            }
        }
        // floats as order-preserving integers: one REDUX per value and warp, one shared atomic per warp
#pragma unroll
        for (int hnd = 0; hnd < 2; ++hnd)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int l = __reduce_min_sync(0xffffffffu, float_ordered(lo[hnd][c]));
                const int hgh = __reduce_max_sync(0xffffffffu, float_ordered(hi[hnd][c]));
                if (lane == 0) { atomicMin(&s.boxi[hnd][0][c], l); atomicMax(&s.boxi[hnd][1][c], hgh); }
            }
        __syncthreads();
        // one thread per hand derives what both directions need (Appendix B: centre, scale = 0.6 * max extent)
        if (tid < 2) {
            float blo[3], bhi[3], cen[3], ext = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                blo[c] = ordered_float(s.boxi[tid][0][c]); bhi[c] = ordered_float(s.boxi[tid][1][c]);
                s.box[tid][0][c] = blo[c]; s.box[tid][1][c] = bhi[c];
                cen[c] = (blo[c] + bhi[c]) * 0.5f;
                ext = fmaxf(ext, bhi[c] - blo[c]);
            }
            const float scale = 0.6f * ext;      // (1 + 0.2) * 0.5 * max extent
            float* d = s.dirp[tid];
            d[3] = scale;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                d[c] = cen[c];
                d[4 + c] = (blo[c] - cen[c]) / scale - 1e-4f;      // normalised extent of the hand itself (+ rounding slack)
                d[7 + c] = (bhi[c] - cen[c]) / scale + 1e-4f;
                d[10 + c] = blo[c] - scale * (2.2f / G);           // the same box in world units, grown by one voxel (+ slack)
                d[13 + c] = bhi[c] + scale * (2.2f / G);
            }
        }
        __syncthreads();
    }
    SDF_STAT(1)

    float loss_part = 0.f;

    for (int h = 0; h < 2; ++h) {
        if ((a.skip_grid_mask >> h) & 1) continue;      // block-uniform: this direction is not needed
        const int o = 1 - h;
        const ushort4* f4 = reinterpret_cast<const ushort4*>(h ? faces_l : faces_r);
        const ushort4* cl_tri = h ? cl_l : cl_r;
        float cen[3], tlo[3], thi[3], wlo[3], whi[3];
        const float scale = s.dirp[h][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cen[c] = s.dirp[h][c]; tlo[c] = s.dirp[h][4 + c]; thi[c] = s.dirp[h][7 + c];
            wlo[c] = s.dirp[h][10 + c]; whi[c] = s.dirp[h][13 + c];
        }
        // block-uniform: can any query vertex pass the reject box at all?
        const bool may = s.box[o][0][0] <= whi[0] && s.box[o][1][1] >= wlo[1] && s.box[o][0][1] <= whi[1] &&
                         s.box[o][1][2] >= wlo[2] && s.box[o][0][2] <= whi[2];
        if (may) {
            for (int i = tid; i < G * G; i += SDF_THREADS) { s.needed[i] = 0u; s.work[i] = 0u; }
            if (tid < 4) s.region[tid] = (tid & 1) ? -1 : G;
            if (tid == 0) s.qn[0] = 0u;                    // queue of the parity rasterisation
            __syncthreads();
        }

        // ---- query vertices: normalised position, voxel corners, mark
        // (the cell of a query vertex is recomputed at sampling time rather than kept in registers)
        auto locate = [&](const float* p, float* fr, int* i0) -> bool {
            // only vertices within one voxel of the grid hand's own box can touch a voxel that may be inside
            // (no lower limit in x: behind the open wrist, voxels left of the mesh can have odd parity)
            bool in = p[0] <= whi[0] && p[1] >= wlo[1] && p[1] <= whi[1] && p[2] >= wlo[2] && p[2] <= whi[2];
            if (!in) return false;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float pn = (p[c] - cen[c]) / scale;
                const float ix = ((pn + 1.0f) * G - 1.0f) * 0.5f;
                const float fl = floorf(ix);
                fr[c] = ix - fl;
                // clamp before the int conversion: far-away vertices must not overflow
                i0[c] = (int)fminf(fmaxf(fl, -2.0f), (float)G);
                in = in && (i0[c] >= -1) && (i0[c] <= G - 1);
            }
            return in;
        };
        float acc[SDF_SLOTS][4];
        uint32_t act = 0u;
        bool any = false;
        int reg[4] = {G, -1, G, -1};         // this thread's marked columns: y min, y max, z min, z max
        float pq[SDF_SLOTS][3];              // all loads in flight before the first use
#pragma unroll
        for (int sl = 0; sl < SDF_SLOTS; ++sl) {
            const int v = tid + sl * SDF_THREADS;
            pq[sl][0] = 3e30f; pq[sl][1] = 0.f; pq[sl][2] = 0.f;         // beyond whi[0]: rejected
            if (may && v < NV) load_vert(o, v, pq[sl]);
        }
#pragma unroll
        for (int sl = 0; sl < SDF_SLOTS; ++sl) {
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[sl][c] = 0.f;
            if (may) {
                float fr[3];
                int i0[3];
                const bool in = locate(pq[sl], fr, i0);
                if (in) {
                    act |= 1u << sl;
                    // A voxel can only be inside (odd +x crossings) if its (y,z) lies within the
                    // mesh's (y,z) extent and its x is left of the mesh's largest x: other corners
                    // are certainly 0 and need not be marked.
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
                            const int zc = i0[2] + dz, yc = i0[1] + dy;
                            if (zc < 0 || zc >= G || yc < 0 || yc >= G) continue;
                            const float yv = voxel_center(yc), zv = voxel_center(zc);
                            if (yv < tlo[1] || yv > thi[1] || zv < tlo[2] || zv > thi[2]) continue;
                            uint32_t bits = 0u;
                            if (i0[0] >= 0 && voxel_center(i0[0]) <= thi[0]) bits |= 1u << i0[0];
                            if (i0[0] + 1 < G && voxel_center(i0[0] + 1) <= thi[0]) bits |= 1u << (i0[0] + 1);
                            if (bits) {
                                any = true;
                                atomicOr(&s.needed[zc * G + yc], bits);
                                reg[0] = min(reg[0], yc); reg[1] = max(reg[1], yc); reg[2] = min(reg[2], zc); reg[3] = max(reg[3], zc);
                            }
                        }
                }
            }
        }
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int r = reg[i];
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) {
                    const int t = __shfl_xor_sync(0xffffffffu, r, sft);
                    r = (i & 1) ? max(r, t) : min(r, t);
                }
                if (lane == 0) { if (i & 1) atomicMax(&s.region[i], r); else atomicMin(&s.region[i], r); }
            }
        }
        if (may) {
            // ---- normalised grid-hand vertices (same phase as the marking: the barrier below covers both)
            for (int v = tid; v < NV; v += SDF_THREADS) {
                float p[3];
                load_vert(h, v, p);
#pragma unroll
                for (int c = 0; c < 3; ++c) s.U[v * 3 + c] = (p[c] - cen[c]) / scale;
            }
        }
        const bool any_block = may && __syncthreads_or(any);
        if (a.stats) {
            int nact = 0;
            for (int sl = 0; sl < SDF_SLOTS; ++sl) nact += __syncthreads_count((act >> sl) & 1u);
            if (tid == 0) a.stats[b * 32 + 4 + h] = nact;
        }
        SDF_STAT(2)
        bool run = any_block;
        int total = 0;

        if (run) {
            SDF_STAT(3)
            // ---- parity of the marked columns, two steps so that the ray tests run with full warps:
            //      (1) thread per face: lattice points inside its (y,z) box whose column is marked
            //          -> (face, column) items; (2) thread per item: the exact ray test
            auto ray_item = [&](int f, int col) {
                const ushort4 id = f4[f];
                float x;
                if (!ray_hit(s.U, id.x, id.y, id.z, voxel_center(col & 31), voxel_center(col >> 5), x)) return;
                // voxels whose centre lies strictly left of the crossing
                int cnt = min(G, max(0, (int)ceilf((x * G + (G - 1)) * 0.5f)));
                while (cnt < G && x > voxel_center(cnt)) ++cnt;
                while (cnt > 0 && !(x > voxel_center(cnt - 1))) --cnt;
                if (cnt > 0) atomicXor(&s.work[col], cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u));
            };
            const float ry0 = voxel_center(s.region[0]) - 0.01f * (2.0f / G), ry1 = voxel_center(s.region[1]) + 0.01f * (2.0f / G);
            const float rz0 = voxel_center(s.region[2]) - 0.01f * (2.0f / G), rz1 = voxel_center(s.region[3]) + 0.01f * (2.0f / G);
            for (int f0 = 0; f0 < NF; f0 += SDF_THREADS) {
                const int f = f0 + tid;
                int j0 = 0, j1 = -1, k0 = 0, k1 = -1;
                if (f < NF) {
                    const ushort4 id = f4[f];
                    const float* A_ = s.U + 3 * id.x; const float* B_ = s.U + 3 * id.y; const float* C_ = s.U + 3 * id.z;
                    const float ymin = fminf(A_[1], fminf(B_[1], C_[1])), ymax = fmaxf(A_[1], fmaxf(B_[1], C_[1]));
                    const float zmin = fminf(A_[2], fminf(B_[2], C_[2])), zmax = fmaxf(A_[2], fmaxf(B_[2], C_[2]));
                    // faces that miss the (y,z) region of the marked columns (with 1/100 cell to spare) are done
                    if (!(ymax < ry0 || ymin > ry1 || zmax < rz0 || zmin > rz1)) {
                        // lattice points y_j = (2j+1-G)/G inside [ymin,ymax]
                        // (1e-3 of a cell absorbs the rounding of the index arithmetic; the ray test itself decides)
                        j0 = max(0, (int)ceilf((ymin * G + (G - 1)) * 0.5f - 1e-3f)); j1 = min(G - 1, (int)floorf((ymax * G + (G - 1)) * 0.5f + 1e-3f));
                        k0 = max(0, (int)ceilf((zmin * G + (G - 1)) * 0.5f - 1e-3f)); k1 = min(G - 1, (int)floorf((zmax * G + (G - 1)) * 0.5f + 1e-3f));
                    }
                }
                if (__ballot_sync(0xffffffffu, j1 >= j0 && k1 >= k0) == 0u) continue;      // no lane covers a lattice point
                // Most faces cover at most 2 x 2 lattice points: those are handled with the warp converged and
                // one queue reservation per warp and lattice slot; the lattice box of a larger face is spread
                // over the lanes of its warp.
                auto push = [&](bool ok, int face, int col) {          // warp-converged
                    const uint32_t m = __ballot_sync(0xffffffffu, ok);
                    if (m == 0u) return;
                    const int leader = __ffs(m) - 1;
                    uint32_t base = 0u;
                    if (lane == leader) base = atomicAdd(&s.qn[0], (uint32_t)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (ok) {
                        const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < Q_CAP) s.queue[pos] = ((uint32_t)face << 10) | (uint32_t)col;
                        else ray_item(face, col);
                    }
                };
                const bool small = (j1 - j0 <= 1) && (k1 - k0 <= 1);
#pragma unroll
                for (int dk = 0; dk < 2; ++dk)
#pragma unroll
                    for (int dj = 0; dj < 2; ++dj) {
                        const int k = k0 + dk, j = j0 + dj, col = k * G + j;
                        push(small && k <= k1 && j <= j1 && s.needed[col & (G * G - 1)] != 0u, f, col);
                    }
                uint32_t bigm = __ballot_sync(0xffffffffu, !small);
                while (bigm) {
                    const int src = __ffs(bigm) - 1;
                    bigm &= bigm - 1u;
                    const int fj0 = __shfl_sync(0xffffffffu, j0, src), fj1 = __shfl_sync(0xffffffffu, j1, src);
                    const int fk0 = __shfl_sync(0xffffffffu, k0, src), fk1 = __shfl_sync(0xffffffffu, k1, src);
                    const int ff = __shfl_sync(0xffffffffu, f, src);
                    const int nj = fj1 - fj0 + 1, npts = nj * (fk1 - fk0 + 1);
                    for (int t0 = 0; t0 < npts; t0 += 32) {
                        const int t = t0 + lane, dk = t / nj, col = (fk0 + dk) * G + fj0 + (t - dk * nj);
                        push(t < npts && s.needed[col & (G * G - 1)] != 0u, ff, col);
                    }
                }
            }
            __syncthreads();
            {
                const int nq = min((int)s.qn[0], Q_CAP);
                if (a.stats && tid == 0) a.stats[b * 32 + 21] += (int)s.qn[0];
                for (int p2 = tid; p2 < nq; p2 += SDF_THREADS) ray_item(s.queue[p2] >> 10, s.queue[p2] & 1023u);
            }
            __syncthreads();
            SDF_STAT(4)
            // ---- marked & inside, prefix offsets, per-row column masks
            int cnt[SDF_CPT], excl[SDF_CPT];
#pragma unroll
            for (int i = 0; i < SDF_CPT; ++i) {
                const int c = tid * SDF_CPT + i;
                const uint32_t wk = s.needed[c] & s.work[c];
                s.work[c] = wk;
                cnt[i] = __popc(wk);
            }
            if (a.stats) {
                int nm = 0;
                for (int i = 0; i < SDF_CPT; ++i) nm += __popc(s.needed[tid * SDF_CPT + i]);
                atomicAdd(&a.stats[b * 32 + 20], nm);
            }
            total = block_scan_1024(cnt, excl, s.scan_warp);
#pragma unroll
            for (int i = 0; i < SDF_CPT; ++i) s.coloff[tid * SDF_CPT + i] = (uint16_t)excl[i];
            run = total > 0;
            SDF_STAT(5)
        }
        if (run) {
            // ---- passes over the marked & inside voxels
            for (int pass0 = 0; pass0 < total; pass0 += PHI_CAP) {
                const int pass1 = min(total, pass0 + PHI_CAP);
                const int nvox = pass1 - pass0;
#pragma unroll
                for (int i = 0; i < SDF_CPT; ++i) {
                    const int c = tid * SDF_CPT + i;
                    uint32_t wk = s.work[c];
                    int idx = s.coloff[c];
                    while (wk) {
                        const int x = __ffs(wk) - 1;
                        wk &= wk - 1;
                        if (idx >= pass0 && idx < pass1) s.worklist[idx - pass0] = pack_q8(x, c & 31, c >> 5);
                        ++idx;
                    }
                }
                for (int i = tid; i < nvox; i += SDF_THREADS) s.best[i] = 0x7f7fffffu;
                if (tid == 0) { s.qn[0] = 0u; s.qn[1] = 0u; s.pn[0] = 0u; s.pn[1] = 0u; s.pn[2] = 0u; s.far_count = 0; }
                SDF_STAT(6)
                // ---- nearest face of every voxel, bulk-synchronous and balanced (work is indexed by voxel)
                //   (0) per face (static clusters of <= 32, spatially sorted): bounding box quantised outwards
                //       to Q8; per cluster: the union of its face boxes
                for (int c = warp; c < NCL; c += SDF_WARPS) {
                    const ushort4 id = cl_tri[c * 32 + lane];
                    int lo[3] = {255, 255, 255}, hi[3] = {0, 0, 0};
                    if (id.w) {
                        const float* A_ = s.U + 3 * id.x; const float* B_ = s.U + 3 * id.y; const float* C_ = s.U + 3 * id.z;
#pragma unroll
                        for (int ax = 0; ax < 3; ++ax) {
                            const float l = fminf(A_[ax], fminf(B_[ax], C_[ax])), hgh = fmaxf(A_[ax], fmaxf(B_[ax], C_[ax]));
                            lo[ax] = max(0, min(255, (int)floorf((l + 1.0f) * 128.0f - 1e-3f)));
                            hi[ax] = max(0, min(255, (int)ceilf((hgh + 1.0f) * 128.0f + 1e-3f)));
                        }
                    }
                    const uint32_t far = id.w ? 0u : 255u << 24;        // empty slot: beyond every radius
                    s.fbox[c * 32 + lane] = id.w ? make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 8) | ((uint32_t)lo[2] << 16),
                                                              (uint32_t)hi[0] | ((uint32_t)hi[1] << 8) | ((uint32_t)hi[2] << 16))
                                                 : make_uint2(far, far);
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
#pragma unroll
                        for (int sft = 16; sft >= 1; sft >>= 1) {
                            lo[ax] = min(lo[ax], __shfl_xor_sync(0xffffffffu, lo[ax], sft));
                            hi[ax] = max(hi[ax], __shfl_xor_sync(0xffffffffu, hi[ax], sft));
                        }
                    }
                    if (lane == 0)
                        s.cl_box[c] = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 8) | ((uint32_t)lo[2] << 16),
                                                 (uint32_t)hi[0] | ((uint32_t)hi[1] << 8) | ((uint32_t)hi[2] << 16));
                }
                __syncthreads();
                // Rounds of growing radius T_k (squared, Q8 units).  After round k every face whose quantised box
                // is closer than T_k has been tested against every voxel still open, or was rejected by that
                // voxel's best at the time (it cannot be nearer then): a voxel whose best is below T_k is final.
                //   (A) thread per (open voxel, cluster): cluster box closer than T_k and than the voxel's best
                //       -> (voxel, cluster) pairs.  The cluster box is the union of the face boxes, so it is
                //       never farther than any of them.
                //   (B) thread per (pair, face of the cluster): face box in the shell [T_{k-1}, T_k) and closer
                //       than the voxel's best -> (voxel, face) candidates (inner shells were done in earlier rounds)
                //   (C) thread per candidate: exact point-triangle test, atomicMin into the voxel
#if SDF_NBANDS == 3
                const int t2[4] = {0, SDF_T0, SDF_T1, SDF_R2_Q8};
#elif SDF_NBANDS == 2
                const int t2[3] = {0, SDF_T0, SDF_R2_Q8};
#else
                const int t2[2] = {0, SDF_R2_Q8};
#endif
                int pa = 0, qb = 0;                  // fill counters of the current (A) / (B) round
                for (int band = 0; band < SDF_NBANDS; ++band) {
                    const int t_lo = t2[band], t_hi = t2[band + 1];
                    const float b_lo = (float)t_lo * Q8_TO_D2;
                    for (int v0 = 0; v0 < nvox; v0 += V_CHUNK, pa = (pa == 2) ? 0 : pa + 1) {
                        uint32_t* pn = &s.pn[pa];
                        if (tid == 0) s.pn[(pa == 2) ? 0 : pa + 1] = 0u;   // read last two rounds ago, used next round
                        const int nvc = min(V_CHUNK, nvox - v0);
                        for (int i = tid; i < nvc * NCL; i += SDF_THREADS) {
                            const int vl = i / NCL, c = i - vl * NCL, v = v0 + vl;
                            const float bv = __uint_as_float(s.best[v]);
                            if (bv < b_lo) continue;
                            const int d2 = qbox_4d2(s.cl_box[c], s.worklist[v]);          // 4 x squared distance
                            if ((!SDF_SHELL && d2 < 4 * t_lo) || d2 >= 4 * t_hi || (float)d2 * (0.25f * Q8_TO_D2) >= bv) continue;
                            const uint32_t pos = atomicAdd(pn, 1u);
                            if (pos < P_CAP) s.pairs[pos] = ((uint32_t)v << 6) | (uint32_t)c;
                            else for (int l = 0; l < 32; ++l) if (cl_tri[c * 32 + l].w) pair_test(s, cl_tri, v, c * 32 + l);
                        }
                        __syncthreads();
                        const int np = min((int)*pn, P_CAP);
                        if (a.stats && tid == 0) { a.stats[b * 32 + 6] += (int)*pn; a.stats[b * 32 + 22 + 2 * band] += (int)*pn; }
                        // The pairs are dealt round-robin to the (B)/(C) rounds (at least SDF_SPLIT of them in the
                        // inner band): a voxel's clusters land in different rounds, so all but the first see its
                        // best distance and reject most faces.
                        const int nround = max((np + P_CHUNK - 1) / P_CHUNK, band == 0 ? min(SDF_SPLIT, np) : 1);
                        for (int r = 0; r < nround && np > 0; ++r, qb ^= 1) {
                            uint32_t* qn = &s.qn[qb];
                            const int npc = (np - r + nround - 1) / nround;      // pairs r, r + nround, ...
                            for (int jj = tid; jj < npc * 32; jj += SDF_THREADS) {
                                const uint32_t pr = s.pairs[r + (jj >> 5) * nround];
                                const int v = pr >> 6, slot = (pr & 63u) * 32 + (jj & 31);
                                const uint2 fb = s.fbox[slot];
                                const int d2 = qbox_4d2(fb, s.worklist[v]);                   // 4 x squared distance
                                if ((SDF_SHELL ? (d2 < 4 * t_lo || d2 >= 4 * t_hi) : d2 >= 4 * SDF_R2_Q8) || (float)d2 * (0.25f * Q8_TO_D2) >= __uint_as_float(s.best[v])) continue;
                                const uint32_t pos = atomicAdd(qn, 1u);
                                if (pos < Q_CAP) s.queue[pos] = ((uint32_t)v << 16) | (uint32_t)slot;
                                else pair_test(s, cl_tri, v, slot);          // queue full: test in place
                            }
                            __syncthreads();
                            const int nq = min((int)*qn, Q_CAP);
                            if (tid == 0) s.qn[qb ^ 1] = 0u;              // last read before this round's (B)
                            if (a.stats && tid == 0) { a.stats[b * 32 + 7] += (int)*qn; a.stats[b * 32 + 23 + 2 * band] += (int)*qn; }
                            for (int p2 = tid; p2 < nq; p2 += SDF_THREADS) pair_test(s, cl_tri, s.queue[p2] >> 16, s.queue[p2] & 0xffffu);
                            __syncthreads();
                        }
                    }
                }
                SDF_STAT(7)
                // ---- certified voxels get phi; the others (nearest face beyond R) go to the far search
                for (int i = tid; i < nvox; i += SDF_THREADS) {
                    const float d2 = __uint_as_float(s.best[i]);
                    if (d2 < SDF_R2 * 0.9999f) s.best[i] = __float_as_uint(sqrtf(d2));
                    else s.far_list[atomicAdd(&s.far_count, 1)] = (uint16_t)i;
                }
                __syncthreads();
                const int nfar = s.far_count;
                if (a.stats && tid == 0) { a.stats[b * 32 + 2 * h] += nvox; a.stats[b * 32 + 2 * h + 1] += nfar; }
                if (nfar > 0) {
                    for (int w = warp; w < nfar; w += SDF_WARPS) {
                        const int i = s.far_list[w];
                        const float d2 = eval_voxel_far(s, cl_tri, s.worklist[i], __uint_as_float(s.best[i]), lane);
                        if (lane == 0) s.best[i] = __float_as_uint(sqrtf(d2));
                    }
                    __syncthreads();
                }
                SDF_STAT(8)
                // ---- trilinear sample + gradient (grid_sampler_3d fwd/bwd, align_corners=False, zeros)
#pragma unroll
                for (int sl = 0; sl < SDF_SLOTS; ++sl) {
                    if (!((act >> sl) & 1u)) continue;
                    float fr[3], pv[3];
                    int i0[3];
                    load_vert(o, tid + sl * SDF_THREADS, pv);
                    locate(pv, fr, i0);
                    const float tx = fr[0], ty = fr[1], tz = fr[2];
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 2; ++dx) {
                                const int xc = i0[0] + dx, yc = i0[1] + dy, zc = i0[2] + dz;
                                if (xc < 0 || xc >= G || yc < 0 || yc >= G || zc < 0 || zc >= G) continue;
                                const int c = zc * G + yc;
                                const uint32_t wk = s.work[c];
                                if (!((wk >> xc) & 1u)) continue;
                                const int idx = s.coloff[c] + __popc(wk & ((1u << xc) - 1u));
                                if (idx < pass0 || idx >= pass1) continue;
                                const float val = __uint_as_float(s.best[idx - pass0]);
                                const float wx = dx ? tx : 1.0f - tx, wy = dy ? ty : 1.0f - ty, wz = dz ? tz : 1.0f - tz;
                                acc[sl][0] += val * wx * wy * wz;
                                acc[sl][1] += (dx ? val : -val) * wy * wz;
                                acc[sl][2] += (dy ? val : -val) * wx * wz;
                                acc[sl][3] += (dz ? val : -val) * wx * wy;
                            }
                }
                __syncthreads();
                SDF_STAT(9)
            }
        }

        // ---- per-vertex outputs of the query hand o
        float gsum[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int sl = 0; sl < SDF_SLOTS; ++sl) {
            const int v = tid + sl * SDF_THREADS;
            if (v >= NV) continue;
            const float psi = acc[sl][0];
            float rho = psi, drho = 1.0f;
            if (a.robustifier > 0.f) {
                const float t = psi / a.robustifier, frac = t * t;
                rho = frac / (frac + 1.0f);
                drho = 2.0f * t / a.robustifier / ((frac + 1.0f) * (frac + 1.0f));
            }
            loss_part += rho;
            const size_t ov = (size_t)b * (2 * NV) + o * NV + v;
            if (a.per_vert) a.per_vert[ov] = rho;
            if (a.origin) a.origin[ov] = psi * scale;
            if (a.gverts || a.gshift) {
                // d psi / d vertex = (G/2) * d psi / d(ix) / scale ; loss = sum(rho) / 4
                const float kk = mask * a.grad_scale * 0.25f * drho * (0.5f * G) / scale;
                float g[3] = {kk * acc[sl][1], kk * acc[sl][2], kk * acc[sl][3]};
                gsum[0] += g[0]; gsum[1] += g[1]; gsum[2] += g[2];
                if (a.gverts) {
                    if (xform && o == 1) g[0] = -g[0];
                    float* gp = a.gverts + ov * 3;
                    gp[0] = g[0]; gp[1] = g[1]; gp[2] = g[2];
                }
            }
        }
        if (a.gshift && o == 1) {
            block_sum4(gsum, 3, s.red);
            if (tid < 3) a.gshift[(size_t)b * 3 + tid] = gsum[tid];
        }
        __syncthreads();
        SDF_STAT(10)
    }
    float lp[1] = {loss_part};
    block_sum4(lp, 1, s.red);
    if (tid == 0) a.losses[b] = mask * lp[0] * 0.25f;
}

int launch_sdf(const ihmr_model* m, int B, const SdfArgs& a, cudaStream_t st) {
    if (B <= 0) return IHMR_OK;
    if (reinterpret_cast<uintptr_t>(a.verts) & 15u) { set_error("sdf: the vertex buffer must be 16-byte aligned"); return IHMR_E_INVALID; }
    static unsigned long long configured = 0ull;
    if (int rc = ensure_dynamic_smem(k_sdf, sizeof(SdfSmem), configured)) return rc;
    k_sdf<<<B, SDF_THREADS, sizeof(SdfSmem), st>>>(B, a, m->faces[0], m->faces[1],
                                                    reinterpret_cast<const ushort4*>(m->cl_tri[0]),
                                                    reinterpret_cast<const ushort4*>(m->cl_tri[1]));
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

}  // namespace ihmr
