// Left/right interpenetration loss, forward + backward.
//
// Two kernels per call:
//
//   k_sdf_prep   one warp per frame: bounding boxes of both hands (A2) -> centre, scale and the
//                reject boxes of both directions; decides per direction (grid hand h, query
//                hand o = 1-h) whether any query vertex can touch a voxel that may be inside.
//                Directions that cannot are finished here (zero outputs); the others are
//                appended to a compact work list.
//   k_sdf_dir    persistent CTAs draw (frame, direction) items from the work list:
//     1. mark the <= 8 voxel corners each query vertex touches                  (lazy grid)
//     2. per face (static clusters of <= 32 faces): bounding box quantised outwards to 1/8
//        voxel (Q8); per cluster the union.  Computed once, used by 3. and 4.
//     3. inside/outside of the marked voxel columns: every face whose box covers a marked
//        (y,z) lattice point is tested against that +x ray; a hit toggles the bits of all
//        voxels left of the crossing                                                   (A4)
//     4. phi = min point-triangle distance for every voxel that is both marked and inside.
//        A seed face per voxel (the nearest face of the previous iteration, else nearest
//        cluster box -> nearest face box) and its exact distance, one voxel per thread = a
//        tight upper bound; then, one warp per voxel, only the faces whose box is closer than
//        that bound are queued, and the exact tests run densely, one per thread.
//     5. trilinear sampling with grid_sample(align_corners=False, zeros) semantics, its
//        gradient w.r.t. the query vertex, per-direction loss part                (A3, A5, A6)
//
// The voxel values are exactly those of the brute-force 32^3 grid of the reference kernel
// (`sdf_cuda`, reached from /root/reference/src/models/loss_utils.py:181): integer crossing
// counts and `min` are order independent, voxels that are not marked never contribute, and a
// face is only ever skipped when a lower bound of its distance is not below an upper bound of
// the minimum.  The inside test uses the arithmetic contract of oracle/sdf_oracle.c (no FMA
// contraction, edge functions on (low id, high id) ordering) so both make the same decisions.
#include <algorithm>

#include "kernels.cuh"

namespace ihmr {

constexpr int G = 32;
#ifndef SDF_NTHREADS
#define SDF_NTHREADS 256
#endif
#ifndef SDF_ROLL
#define SDF_ROLL 1                  // 1: the per-thread vertex loops stay rolled (code size: the kernel is instruction-fetch sensitive)
#endif
#if SDF_ROLL
#define SDF_SLOT_LOOP _Pragma("unroll 1")
#else
#define SDF_SLOT_LOOP _Pragma("unroll")
#endif
constexpr int SDF_THREADS = SDF_NTHREADS;
constexpr int SDF_WARPS = SDF_THREADS / 32;
constexpr int SDF_CPT = G * G / SDF_THREADS;   // (z,y) columns per thread in the scans
#ifndef SDF_MIN_CTAS
#define SDF_MIN_CTAS 4               // resident CTAs per SM the register allocation is held to
#endif
constexpr int SDF_SLOTS = (NV + SDF_THREADS - 1) / SDF_THREADS;   // query vertices per thread
constexpr int NCL = (NF + 31) / 32; // static face clusters of <= 32 faces (49)
// Capacities (shared memory).  Everything beyond them is handled, not dropped: more voxels than
// PHI_CAP -> further passes (finished values parked in a global spill area), a candidate / ray queue
// segment that cannot take 32 more entries -> its entries are tested right away.  tests build a
// variant with tiny capacities.
#ifndef SDF_PHI_CAP
#define SDF_PHI_CAP 1024            // voxels evaluated per pass
#endif
#ifndef SDF_Q_CAP
#define SDF_Q_CAP 2304              // queued (voxel, face) candidates / (face, column) ray items
#endif
constexpr int PHI_CAP = SDF_PHI_CAP, Q_CAP = SDF_Q_CAP;
constexpr int QSEG = Q_CAP / SDF_WARPS;   // candidate queue segment of one warp
static_assert(PHI_CAP <= 65536, "queue entries pack the voxel index into 16 bits");
static_assert(QSEG >= 64, "a queue segment takes at least two rounds of 32 entries");
constexpr int SDF_SPILL = NV * 8;   // a direction evaluates at most 8 voxels per query vertex
constexpr int SDF_MAX_GRID = 160 * 8;
constexpr int SDF_HINTS = 2048;     // nearest-face hints per (frame, direction): one u16 per voxel of an 8 x 16 x 16 block (wraps)
constexpr uint32_t HINT_DONE = 0xffffffffu, HINT_NONE = 0xfffffffeu;   // states of a voxel in hintw (else: seed slot)
constexpr uint32_t SEED_NEW = 0x40000000u;                             // | seed slot: found in this call, distance not measured yet
constexpr int SDF_PCACHE = G * G + G;   // words per frame of the static-grid parity cache: 1024 columns + 32 words of "known" bits
constexpr int SDF_HDR = 40;         // floats per frame header (see k_sdf_prep)
constexpr float Q8_TO_D2 = 1.0f / 16384.0f;         // Q8 units squared -> normalised units squared
constexpr float Q4D2_TO_D2 = 0.25f * Q8_TO_D2;      // qbox_4d2 units -> normalised units squared

constexpr int SDF_BUCKETS = 4;      // work items are drawn heaviest bucket first (longest-processing-time-first: short tail)

struct SdfWs {
    uint32_t* counters;   // [0..3] items appended by k_sdf_prep per cost bucket, [4] ticket of k_sdf_dir
    float* hdr;           // (B, SDF_HDR)
    uint32_t* items;      // (SDF_BUCKETS, 2B) frame * 2 + grid hand
    float* parts;         // (B, 2) sum of rho over the query vertices of each direction
    float* spill;         // (SDF_MAX_GRID, SDF_SPILL)
};

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

size_t sdf_ws_bytes(int B) {
    return 256 + up256((size_t)B * SDF_HDR * 4) + up256((size_t)SDF_BUCKETS * B * 2 * 4) + up256((size_t)B * 2 * 4) +
           up256((size_t)SDF_MAX_GRID * SDF_SPILL * 4);
}

static SdfWs sdf_ws_carve(void* base, int B) {
    char* p = static_cast<char*>(base);
    SdfWs w;
    w.counters = reinterpret_cast<uint32_t*>(p); p += 256;
    w.hdr = reinterpret_cast<float*>(p); p += up256((size_t)B * SDF_HDR * 4);
    w.items = reinterpret_cast<uint32_t*>(p); p += up256((size_t)SDF_BUCKETS * B * 2 * 4);
    w.parts = reinterpret_cast<float*>(p); p += up256((size_t)B * 2 * 4);
    w.spill = reinterpret_cast<float*>(p);
    return w;
}

struct __align__(16) SdfSmem {
    float U[NV * 3];            // normalised grid-hand vertices
    uint32_t needed[G * G];     // marked voxels per (z,y) column
    uint32_t work[G * G];       // parity bits, then marked & inside
    uint16_t coloff[G * G];     // exclusive prefix of popc(work)
    uint32_t worklist[PHI_CAP]; // voxels of the current pass as packed Q8 centres: 8x+4 | (8y+4) << 8 | (8z+4) << 16
    uint32_t best[PHI_CAP];     // bit pattern of the best squared distance (>= 0: orders like uint); then phi
    uint32_t hintw[PHI_CAP];    // face slot behind best (to rounding): next iteration's seed
    uint32_t queue[Q_CAP];      // (voxel index << 16) | face slot;  parity: (face slot << 10) | column
    uint2 cl_box[NCL];          // union of the cluster's face boxes, same packing as fbox
    uint2 fbox[NCL * 32];       // per face (cluster-table order): box quantised outwards to Q8;
                                // .x = lo x | y << 8 | z << 16 | invalid << 24, .y = hi x | y << 8 | z << 16
    float red[4 * SDF_WARPS];
    int region[4];              // lattice bounds of the marked columns: y min, y max, z min, z max
    int scan_warp[SDF_WARPS];
    uint32_t item;              // code of the next work item (frame * 2 + grid hand), 0xffffffff: none left
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

// Ray axis (assumption A4): the kernels work in coordinates whose first component runs along the parity ray.
// world (w0,w1,w2) -> internal (w[axis], w[axis+1], w[axis+2]) (cyclic), and back for the gradients.
__device__ __forceinline__ void to_ray_frame(float* v, int axis) {
    if (axis == 0) return;
    const float a0 = v[0], a1 = v[1], a2 = v[2];
    if (axis == 1) { v[0] = a1; v[1] = a2; v[2] = a0; } else { v[0] = a2; v[1] = a0; v[2] = a1; }
}
__device__ __forceinline__ void from_ray_frame(float* v, int axis) {
    if (axis == 0) return;
    const float a0 = v[0], a1 = v[1], a2 = v[2];
    if (axis == 1) { v[0] = a2; v[1] = a0; v[2] = a1; } else { v[0] = a1; v[1] = a2; v[2] = a0; }
}

// (float_ordered / ordered_float: common.cuh)

// ---- block primitives ---------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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
// in valid boxes; boxes of empty table slots carry 255 there, which puts them beyond every real distance.
__device__ __forceinline__ uint32_t qbox_4d2(uint2 bx, uint32_t q) {
    const uint32_t a = __vabsdiffu4(q, bx.x), b = __vabsdiffu4(q, bx.y), w = __vabsdiffu4(bx.y, bx.x);
    const uint32_t sq = __dp4a(w, w, __dp4a(b, b, __dp4a(a, a, 0u)));
    const uint32_t cr = __dp4a(b, w, __dp4a(a, w, 0u));
    return sq + 2u * (__dp4a(a, b, 0u) - cr);
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

__device__ __forceinline__ float voxel_face_dist2(const SdfSmem& s, const ushort4* __restrict__ cl_tri, uint32_t q8, int slot) {
    float q[3];
    voxel_pos(q8, q);
    const ushort4 id = cl_tri[slot];
    return pt_tri_dist2(q, s.U + 3 * id.x, s.U + 3 * id.y, s.U + 3 * id.z);
}

// one exact (voxel, face) test; the result lowers the voxel's best squared distance.  hintw follows the
// improvements (atomic exchange, no ordering between concurrent improvers): it is a seed, not a result.
__device__ __forceinline__ void pair_test(SdfSmem& s, const ushort4* __restrict__ cl_tri, int vi, int slot) {
    const uint32_t bits = __float_as_uint(voxel_face_dist2(s, cl_tri, s.worklist[vi], slot));
    if (bits < atomicMin(&s.best[vi], bits)) atomicExch(&s.hintw[vi], (uint32_t)slot);
}

// A warp's queue segment of (voxel, face) candidates, tested one per lane.  Out of line for the rare flush in the
// middle of a voxel (segment nearly full); the flush at the end of a warp's voxels is inlined.
__device__ __forceinline__ void pair_flush(SdfSmem& s, const ushort4* __restrict__ cl_tri, const uint32_t* wqueue, int n) {
    __syncwarp();
    for (int p2 = (int)(threadIdx.x & 31); p2 < n; p2 += 32) {
        const uint32_t e = wqueue[p2];
        pair_test(s, cl_tri, (int)(e >> 16), (int)(e & 0xffffu));
    }
    __syncwarp();
}
__device__ __noinline__ void pair_flush_slow(SdfSmem& s, const ushort4* __restrict__ cl_tri, const uint32_t* wqueue, int n) {
    pair_flush(s, cl_tri, wqueue, n);
}

// one (face, column) item of the parity rasterisation: the exact ray test; a hit toggles the bits of all voxels whose
// centre lies strictly left of the crossing
__device__ __forceinline__ void ray_item(SdfSmem& s, const ushort4* __restrict__ cl_tri, int slot, int col) {
    const ushort4 id = cl_tri[slot];
    float x;
    if (!ray_hit(s.U, id.x, id.y, id.z, voxel_center(col & 31), voxel_center(col >> 5), x)) return;
    int cnt = min(G, max(0, (int)ceilf((x * G + (G - 1)) * 0.5f)));
    while (cnt < G && x > voxel_center(cnt)) ++cnt;
    while (cnt > 0 && !(x > voxel_center(cnt - 1))) --cnt;
    if (cnt > 0) atomicXor(&s.work[col], cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u));
}
__device__ __forceinline__ void ray_flush(SdfSmem& s, const ushort4* __restrict__ cl_tri, const uint32_t* rqueue, int n) {
    __syncwarp();
    for (int p2 = (int)(threadIdx.x & 31); p2 < n; p2 += 32) {
        const uint32_t e = rqueue[p2];
        ray_item(s, cl_tri, (int)(e >> 10), (int)(e & 1023u));
    }
    __syncwarp();
}
__device__ __noinline__ void ray_flush_slow(SdfSmem& s, const ushort4* __restrict__ cl_tri, const uint32_t* rqueue, int n) {
    ray_flush(s, cl_tri, rqueue, n);
}

// packed Q8 voxel centre -> index into the hint table: the low bits of (x, y, z), so that a block of
// 8 x 16 x 16 neighbouring voxels never collides
__device__ __forceinline__ int hint_index(uint32_t q8) {
    return (int)(((q8 >> 3) & 7u) | (((q8 >> 11) & 15u) << 3) | (((q8 >> 19) & 15u) << 7));
}
// the remaining bits of (x, y, z): an entry is (tag << 11) | (slot + 1), so a voxel only takes its own hint
__device__ __forceinline__ uint32_t hint_tag(uint32_t q8) {
    return ((q8 >> 6) & 3u) | (((q8 >> 15) & 1u) << 2) | (((q8 >> 23) & 1u) << 3);
}

// lattice indices j with 8j + 4 inside the byte range [lo, hi] (a superset of the lattice points inside the
// unquantised range, with at least the 1e-3 Q8 unit of the outward quantisation to spare)
__device__ __forceinline__ int lat_lo(uint32_t lo) { return (int)(lo + 3u) >> 3; }
__device__ __forceinline__ int lat_hi(uint32_t hi) { return ((int)hi - 4) >> 3; }

// ------------------------------------------------------------------------------------ prep
// Frame header (floats): per hand h at [16h]: centre xyz, scale, tlo xyz, thi xyz, wlo xyz, whi xyz
// (tlo/thi: normalised extent of the hand itself; wlo/whi: its box in world units grown by one voxel);
// [32..34] shift of the left hand, [35] both-hands mask.
constexpr int PREP_WARPS = 8;

__global__ void __launch_bounds__(PREP_WARPS * 32) k_sdf_prep(int B, SdfArgs a, SdfWs w) {
    __shared__ uint32_t s_cnt[SDF_BUCKETS], s_base[SDF_BUCKETS];
    __shared__ uint32_t s_items[SDF_BUCKETS][2 * PREP_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * PREP_WARPS + warp;
    if (threadIdx.x < SDF_BUCKETS) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    if (b < B) {
        const bool xform = (a.joints != nullptr);
        float sh[3] = {0.f, 0.f, 0.f};
        if (xform) {
            const float* jr = a.joints + ((size_t)b * 2 + 0) * 48;
            const float* jl = a.joints + ((size_t)b * 2 + 1) * 48;
            const float* t = a.params + (size_t)b * PD + P_TRANS;
            sh[0] = t[0] + (jr[0] + jl[0]);      // - (-x)
            sh[1] = t[1] + (jr[1] - jl[1]);
            sh[2] = t[2] + (jr[2] - jl[2]);
        }
        float mask = 1.0f;
        if (a.hand_type) mask = (a.hand_type[b * 2] + a.hand_type[b * 2 + 1] > 1.5f) ? 1.0f : 0.0f;
        float lo[2][3], hi[2][3];
        if (a.bbox) {                        // the kernels that wrote the vertices left their boxes (stored frame)
#pragma unroll
            for (int hnd = 0; hnd < 2; ++hnd)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    lo[hnd][c] = a.bbox[((size_t)b * 2 + hnd) * 6 + c];
                    hi[hnd][c] = a.bbox[((size_t)b * 2 + hnd) * 6 + 3 + c];
                }
        } else {
        const float* V = a.verts + (size_t)b * (2 * NV * 3);
#pragma unroll
        for (int hnd = 0; hnd < 2; ++hnd)
#pragma unroll
            for (int c = 0; c < 3; ++c) { lo[hnd][c] = 1e30f; hi[hnd][c] = -1e30f; }
#pragma unroll 5
        for (int v = lane; v < NV; v += 32) {
#pragma unroll
            for (int hnd = 0; hnd < 2; ++hnd) {
                const float* p = V + (hnd * NV + v) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) { const float x = p[c]; lo[hnd][c] = fminf(lo[hnd][c], x); hi[hnd][c] = fmaxf(hi[hnd][c], x); }
            }
        }
        // floats as order-preserving integers: one REDUX per value
#pragma unroll
        for (int hnd = 0; hnd < 2; ++hnd)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lo[hnd][c] = ordered_float(__reduce_min_sync(0xffffffffu, float_ordered(lo[hnd][c])));
                hi[hnd][c] = ordered_float(__reduce_max_sync(0xffffffffu, float_ordered(hi[hnd][c])));
            }
        }
        if (xform) {
            // world frame of the stored (mirrored) left hand: x -> -x + shift, y,z -> + shift.  Rounding is
            // monotone, so the box of the mapped vertices is the mapped box.
            const float l0 = -hi[1][0] + sh[0], h0 = -lo[1][0] + sh[0];
            lo[1][0] = l0; hi[1][0] = h0;
            lo[1][1] += sh[1]; hi[1][1] += sh[1];
            lo[1][2] += sh[2]; hi[1][2] += sh[2];
        }
        to_ray_frame(lo[0], a.ray_axis); to_ray_frame(hi[0], a.ray_axis);
        to_ray_frame(lo[1], a.ray_axis); to_ray_frame(hi[1], a.ray_axis);
        // Appendix B: centre = box midpoint, scale = (1 + scale_factor) * 0.5 * max extent   (A2; 0.6 by default)
        float dp[2][16];
#pragma unroll
        for (int hnd = 0; hnd < 2; ++hnd) {
            float cen[3], ext = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { cen[c] = (lo[hnd][c] + hi[hnd][c]) * 0.5f; ext = fmaxf(ext, hi[hnd][c] - lo[hnd][c]); }
            const float scale = a.box_scale * ext;
            dp[hnd][3] = scale;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                dp[hnd][c] = cen[c];
                dp[hnd][4 + c] = (lo[hnd][c] - cen[c]) / scale - 1e-4f;      // normalised extent of the hand itself (+ rounding slack)
                dp[hnd][7 + c] = (hi[hnd][c] - cen[c]) / scale + 1e-4f;
                dp[hnd][10 + c] = lo[hnd][c] - scale * (2.2f / G);           // the same box in world units, grown by one voxel (+ slack)
                dp[hnd][13 + c] = hi[hnd][c] + scale * (2.2f / G);
            }
        }
        float* hd = w.hdr + (size_t)b * SDF_HDR;
        if (lane < 2) {                      // lane l writes the 16 entries of hand l (static register indices)
            float4* h4 = reinterpret_cast<float4*>(hd + 16 * lane);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                h4[i] = lane ? make_float4(dp[1][4 * i], dp[1][4 * i + 1], dp[1][4 * i + 2], dp[1][4 * i + 3])
                             : make_float4(dp[0][4 * i], dp[0][4 * i + 1], dp[0][4 * i + 2], dp[0][4 * i + 3]);
        }
        if (lane < 3) hd[32 + lane] = (lane == 0) ? sh[0] : (lane == 1 ? sh[1] : sh[2]);
        if (lane == 3) hd[35] = mask;
        auto direction = [&](int h, const float (&dh)[16], const float (&olo)[3], const float (&ohi)[3]) {
            const int o = 1 - h;
            // can any query vertex pass the reject box at all?  (no lower limit in x: behind the open wrist,
            // voxels left of the mesh can have odd parity)
            const bool may = olo[0] <= dh[13] && ohi[1] >= dh[11] && olo[1] <= dh[14] && ohi[2] >= dh[12] && olo[2] <= dh[15];
            const bool skipped = (a.skip_grid_mask >> h) & 1;      // this direction is not wanted at all
            if (may && !skipped) {
                // cost proxy: volume (in voxels of grid hand h) of the query hand's box inside the grid hand's reject box
                float vol = 1.0f;
#pragma unroll
                for (int c = 0; c < 3; ++c) vol *= fmaxf(0.f, fminf(ohi[c], dh[13 + c]) - fmaxf(olo[c], dh[10 + c]));
                const float vs = dh[3] * (2.0f / G);
                vol /= vs * vs * vs;
                const int bucket = vol >= 1500.f ? 0 : vol >= 500.f ? 1 : vol >= 100.f ? 2 : 3;
                if (lane == 0) s_items[bucket][atomicAdd(&s_cnt[bucket], 1u)] = (uint32_t)(b * 2 + h);
                if (a.stats && lane == 0) a.stats[b * 32 + 14 + h] = (int)fminf(vol, 1e9f);
                return;
            }
            if (lane == 0) w.parts[b * 2 + h] = 0.f;
            if (skipped) return;
            // nothing of hand o can touch an inside voxel of hand h: its samples and gradients are exact zeros
            const size_t ov = (size_t)b * (2 * NV) + o * NV;
            if (a.per_vert) for (int v = lane; v < NV; v += 32) a.per_vert[ov + v] = 0.f;
            if (a.origin) for (int v = lane; v < NV; v += 32) a.origin[ov + v] = 0.f;
            if (a.gzero) { if (lane == 0) a.gzero[b * 2 + o] = 1; }      // consumers skip this hand's vertex gradients
            else if (a.gverts) { float* g = a.gverts + ov * 3; for (int i = lane; i < NV * 3; i += 32) g[i] = 0.f; }
            if (a.gshift && o == 1 && lane < 3) a.gshift[(size_t)b * 3 + lane] = 0.f;
            if (a.stats && lane == 0) a.stats[b * 32 + 16 + h] = 1;
        };
        direction(0, dp[0], lo[1], hi[1]);
        direction(1, dp[1], lo[0], hi[0]);
    }
    __syncthreads();
    if (threadIdx.x < SDF_BUCKETS && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&w.counters[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    for (int k = 0; k < SDF_BUCKETS; ++k)
        if (threadIdx.x < s_cnt[k]) w.items[(size_t)k * 2 * B + s_base[k] + threadIdx.x] = s_items[k][threadIdx.x];
}

// losses[b] = mask * (part_0 + part_1) / 4                                              (A6)
__global__ void k_sdf_finish(int B, const float* __restrict__ parts, const float* __restrict__ hand_type, float* losses) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float mask = 1.0f;
    if (hand_type) mask = (hand_type[b * 2] + hand_type[b * 2 + 1] > 1.5f) ? 1.0f : 0.0f;
    losses[b] = mask * (parts[b * 2] + parts[b * 2 + 1]) * 0.25f;
}

#define SDF_STAT(i)                                                                       \
    if (kStats && tid == 0) {                                                            \
        const long long t_now = clock64();                                                \
        atomicAdd(&a.stats[b * 32 + 18 + (i)], (int)(t_now - t_prev));                    \
        t_prev = t_now;                                                                   \
    }
// stats (B,32) int32, diagnostics only: [2h] voxels evaluated with grid hand h, [4+h] query vertices inside the grid
// box, [6] (voxel, cluster) pairs, [7] exact candidates, [8] marked voxels, [9] ray items, [10] passes, [11] ray queue
// segments flushed early (full), [12] candidate queue segments flushed early (full), [13] voxels answered by the
// static-grid cache, [14+h] cost proxy of the direction, [16+h] direction finished by k_sdf_prep, [19..27] cycles per
// phase (mark, face boxes, parity, scan, worklist + seeds, unused, candidate search + exact tests, finish, sample + outputs)

// kStatic = false compiles the static-grid cache out (stages in which both hands move)
// kStats = true only for ihmr_sdf_stats (tools, capacity tests): the counters cost ~300 instructions of code
template <bool kStatic, bool kStats>
__global__ void __launch_bounds__(SDF_THREADS, SDF_MIN_CTAS)
k_sdf_dir(int B, SdfArgs a, SdfWs w, const ushort4* __restrict__ cl_r, const ushort4* __restrict__ cl_l) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SdfSmem& s = *reinterpret_cast<SdfSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool xform = (a.joints != nullptr);
    int bucket_end[SDF_BUCKETS];                  // tickets [bucket_end[k-1], bucket_end[k]) belong to bucket k
    {
        int run = 0;
#pragma unroll
        for (int k = 0; k < SDF_BUCKETS; ++k) { run += (int)w.counters[k]; bucket_end[k] = run; }
    }
    const int n_items = bucket_end[SDF_BUCKETS - 1];
    float* spill = w.spill + (size_t)blockIdx.x * SDF_SPILL;

    // thread 0 draws the ticket of the NEXT item (and looks its code up) while the current one is processed, so the
    // round trips to the ticket counter and the item list are off the critical path
    auto draw = [&]() {
        const int t = (int)atomicAdd(&w.counters[SDF_BUCKETS], 1u);
        uint32_t code = 0xffffffffu;               // no more items
        if (t < n_items) {
            int k = 0, start = 0;
#pragma unroll
            for (int q = 0; q + 1 < SDF_BUCKETS; ++q) if (t >= bucket_end[q]) { k = q + 1; start = bucket_end[q]; }
            code = w.items[(size_t)k * 2 * B + (t - start)];
            // both hands of that frame (18,672 B, 16-byte aligned) on their way from DRAM to L2 while this item is processed
            if (a.l2_prefetch)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                             :: "l"(a.verts + (size_t)(code >> 1) * (2 * NV * 3)), "r"(2 * NV * 3 * 4) : "memory");
        }
        s.item = code;
    };
    if (tid == 0) draw();

    for (;;) {
        __syncthreads();                       // the previous item is completely finished; s.item is the next one
        const uint32_t code = s.item;
        if (code == 0xffffffffu) break;
        const int b = (int)(code >> 1), h = (int)(code & 1u), o = 1 - h;
        const ushort4* cl_tri = h ? cl_l : cl_r;
        uint16_t* hint = a.hints ? a.hints + ((size_t)b * 2 + h) * SDF_HINTS : nullptr;
        const float* hd = w.hdr + (size_t)b * SDF_HDR;
        float cen[3], tlo[3], thi[3], wlo[3], whi[3];
        const float scale = hd[16 * h + 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cen[c] = hd[16 * h + c]; tlo[c] = hd[16 * h + 4 + c]; thi[c] = hd[16 * h + 7 + c];
            wlo[c] = hd[16 * h + 10 + c]; whi[c] = hd[16 * h + 13 + c];
        }
        const float shx = hd[32], shy = hd[33], shz = hd[34], mask = hd[35];
        auto load_vert = [&](int hand, int v, float* out) {
            const float* p = a.verts + (((size_t)b * 2 + hand) * NV + v) * 3;
            out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
            if (xform && hand == 1) { out[0] = -out[0] + shx; out[1] += shy; out[2] += shz; }
            to_ray_frame(out, a.ray_axis);
        };
        long long t_prev = clock64();
        const uint32_t lt_mask = (1u << lane) - 1u;
        for (int i = tid; i < G * G; i += SDF_THREADS) { s.needed[i] = 0u; s.work[i] = 0u; }
        if (tid < 4) s.region[tid] = (tid & 1) ? -1 : G;
        __syncthreads();
        if (tid == 0) draw();                  // (every thread has read s.item of this iteration)

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
        bool any = false;
        int nact = 0;
        {
            int reg[4] = {G, -1, G, -1};         // this thread's marked columns: y min, y max, z min, z max
#if SDF_ROLL
            float nx[3] = {3e30f, 0.f, 0.f};     // the next vertex is in flight while this one is processed
            if (tid < NV) load_vert(o, tid, nx);
#else
            float pq[SDF_SLOTS][3];              // all loads in flight before the first use
#pragma unroll
            for (int sl = 0; sl < SDF_SLOTS; ++sl) {
                const int v = tid + sl * SDF_THREADS;
                pq[sl][0] = 3e30f; pq[sl][1] = 0.f; pq[sl][2] = 0.f;         // beyond whi[0]: rejected
                if (v < NV) load_vert(o, v, pq[sl]);
            }
#endif
            SDF_SLOT_LOOP
            for (int sl = 0; sl < SDF_SLOTS; ++sl) {
                float fr[3];
                int i0[3];
#if SDF_ROLL
                const float cur[3] = {nx[0], nx[1], nx[2]};
                nx[0] = 3e30f;                                                // beyond whi[0]: rejected
                if (tid + (sl + 1) * SDF_THREADS < NV) load_vert(o, tid + (sl + 1) * SDF_THREADS, nx);
                if (!locate(cur, fr, i0)) continue;
#else
                if (!locate(pq[sl], fr, i0)) continue;
#endif
                ++nact;
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
            if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = (i & 1) ? __reduce_max_sync(0xffffffffu, reg[i]) : __reduce_min_sync(0xffffffffu, reg[i]);
                    if (lane == 0) { if (i & 1) atomicMax(&s.region[i], r); else atomicMin(&s.region[i], r); }
                }
            }
        }
        const bool any_block = __syncthreads_or(any);
        if (kStats) {
            const int n = __reduce_add_sync(0xffffffffu, nact);
            if (lane == 0) atomicAdd(&a.stats[b * 32 + 4 + h], n);
        }
        SDF_STAT(1)
        int total = 0;
        // Static grid (a stage that moves only hand_trans: the right hand is bit-identical over the whole stage):
        // column parities and finished voxel distances of this frame are carried in HBM from one iteration to the
        // next, so only newly touched columns / voxels need the grid hand's geometry at all.
        const bool stat = kStatic && ((a.static_grid_mask >> h) & 1);
        uint32_t* pcw = stat ? a.pcache + (size_t)b * SDF_PCACHE : nullptr;      // [1024] parity words, [32] known bits
        float* phic = stat ? a.phic + (size_t)b * SDF_HINTS : nullptr;
        bool geom_ready = false;
        auto geometry = [&]() {          // block-uniform: U, face boxes, cluster boxes (ends with a barrier)
            // ---- normalised grid-hand vertices
            for (int v = tid; v < NV; v += SDF_THREADS) {
                float p[3];
                load_vert(h, v, p);
#pragma unroll
                for (int c = 0; c < 3; ++c) s.U[v * 3 + c] = (p[c] - cen[c]) / scale;
            }
            __syncthreads();
            // ---- per face (static clusters of <= 32, spatially sorted): bounding box quantised outwards to Q8;
            //      per cluster: the union of its face boxes
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
                const uint32_t far = 255u << 24;        // empty slot: beyond every real distance
                s.fbox[c * 32 + lane] = id.w ? make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 8) | ((uint32_t)lo[2] << 16),
                                                          (uint32_t)hi[0] | ((uint32_t)hi[1] << 8) | ((uint32_t)hi[2] << 16))
                                             : make_uint2(far, far);
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    lo[ax] = __reduce_min_sync(0xffffffffu, lo[ax]);
                    hi[ax] = __reduce_max_sync(0xffffffffu, hi[ax]);
                }
                if (lane == 0)
                    s.cl_box[c] = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 8) | ((uint32_t)lo[2] << 16),
                                             (uint32_t)hi[0] | ((uint32_t)hi[1] << 8) | ((uint32_t)hi[2] << 16));
            }
            __syncthreads();
            geom_ready = true;
        };
        if (any_block) {
            // which marked columns still need their parity: all of them, or (static grid) those not known yet
            bool any_new = true;
            if (stat) {
                bool mynew = false;
                for (int c = tid; c < G * G; c += SDF_THREADS) {
                    uint16_t flag = 0;
                    if (s.needed[c]) {
                        if ((pcw[G * G + (c >> 5)] >> (c & 31)) & 1u) s.work[c] = pcw[c];
                        else { flag = 1; mynew = true; }
                    }
                    s.coloff[c] = flag;      // (coloff is free until the scan)
                }
                any_new = __syncthreads_or(mynew);
            }
            if (any_new) {
            geometry();
            SDF_STAT(2)
            // ---- parity of the marked columns, two steps so that the ray tests run with full warps:
            //      (1) thread per face: lattice points inside its (y,z) box whose column is marked
            //          -> (face, column) items in the warp's queue segment; (2) lane per item: the exact ray test
            const int ry0 = s.region[0], ry1 = s.region[1], rz0 = s.region[2], rz1 = s.region[3];
            uint32_t* rqueue = s.queue + warp * QSEG;
            int rq = 0, nray = 0;                                // fill of this warp's segment (warp-uniform), items queued
            for (int c = warp; c < NCL; c += SDF_WARPS) {
                {   // clusters that miss the (y,z) region of the marked columns are done
                    const uint2 cb = s.cl_box[c];
                    if (lat_hi((cb.y >> 8) & 255u) < ry0 || lat_lo((cb.x >> 8) & 255u) > ry1 ||
                        lat_hi((cb.y >> 16) & 255u) < rz0 || lat_lo((cb.x >> 16) & 255u) > rz1) continue;
                }
                const int slot = c * 32 + lane;
                const uint2 fb = s.fbox[slot];
                int j0 = max(ry0, lat_lo((fb.x >> 8) & 255u)), j1 = min(ry1, lat_hi((fb.y >> 8) & 255u));
                int k0 = max(rz0, lat_lo((fb.x >> 16) & 255u)), k1 = min(rz1, lat_hi((fb.y >> 16) & 255u));
                if (fb.x >> 24) j1 = j0 - 1;                                   // empty table slot
                if (__ballot_sync(0xffffffffu, j1 >= j0 && k1 >= k0) == 0u) continue;      // no lane covers a lattice point
                // Most faces cover at most 2 x 2 lattice points: those are handled with the warp converged and
                // one queue reservation per warp and lattice slot; the lattice box of a larger face is spread
                // over the lanes of its warp.
                auto push = [&](bool ok, int face, int col) {          // warp-converged; the warp owns its queue segment
                    const uint32_t m = __ballot_sync(0xffffffffu, ok);
                    if (m == 0u) return;
                    if (rq > QSEG - 32) {                              // no room for 32 more: test what is queued now
                        ray_flush_slow(s, cl_tri, rqueue, rq);
                        rq = 0;
                        if (kStats && lane == 0) atomicAdd(&a.stats[b * 32 + 11], 1);
                    }
                    if (ok) rqueue[rq + __popc(m & lt_mask)] = ((uint32_t)face << 10) | (uint32_t)col;
                    rq += __popc(m);
                    if (kStats) nray += __popc(m);
                };
                const bool cover = (j1 >= j0) && (k1 >= k0);
                const bool small = cover && (j1 - j0 <= 1) && (k1 - k0 <= 1);
#pragma unroll 1
                for (int i = 0; i < 4; ++i) {          // (rolled: one reservation site)
                    const int k = k0 + (i >> 1), j = j0 + (i & 1), col = k * G + j;
                    push(small && k <= k1 && j <= j1 && (stat ? s.coloff[col & (G * G - 1)] != 0 : s.needed[col & (G * G - 1)] != 0u), slot, col);
                }
                uint32_t bigm = __ballot_sync(0xffffffffu, cover && !small);
                while (bigm) {
                    const int src = __ffs(bigm) - 1;
                    bigm &= bigm - 1u;
                    const int fj0 = __shfl_sync(0xffffffffu, j0, src), fj1 = __shfl_sync(0xffffffffu, j1, src);
                    const int fk0 = __shfl_sync(0xffffffffu, k0, src), fk1 = __shfl_sync(0xffffffffu, k1, src);
                    const int ff = c * 32 + src;
                    const int nj = fj1 - fj0 + 1, npts = nj * (fk1 - fk0 + 1);
                    // t / nj for t < 1024, nj <= 32 in floating point: (t + 0.5) / nj stays 1/64 away from every integer
                    const float rnj = __fdividef(1.0f, (float)nj);
#pragma unroll 1
                    for (int t0 = 0; t0 < npts; t0 += 32) {
                        const int t = t0 + lane, dk = (int)(((float)t + 0.5f) * rnj), col = (fk0 + dk) * G + fj0 + (t - dk * nj);
                        push(t < npts && (stat ? s.coloff[col & (G * G - 1)] != 0 : s.needed[col & (G * G - 1)] != 0u), ff, col);
                    }
                }
            }
            if (kStats && lane == 0) atomicAdd(&a.stats[b * 32 + 9], nray);
            ray_flush(s, cl_tri, rqueue, rq);     // the ray tests by the warp that queued them: no block barrier in between
            __syncthreads();
            if (stat) {                  // the new columns' parity words are known from now on
                for (int c = tid; c < G * G; c += SDF_THREADS)
                    if (s.coloff[c]) { pcw[c] = s.work[c]; atomicOr(&pcw[G * G + (c >> 5)], 1u << (c & 31)); }
                __syncthreads();         // (coloff is rewritten by the scan below)
            }
            }   // any_new
            SDF_STAT(3)
            // ---- marked & inside, prefix offsets
            int cnt[SDF_CPT], excl[SDF_CPT];
#pragma unroll
            for (int i = 0; i < SDF_CPT; ++i) {
                const int c = tid * SDF_CPT + i;
                const uint32_t wk = s.needed[c] & s.work[c];
                s.work[c] = wk;
                cnt[i] = __popc(wk);
            }
            if (kStats) {
                int nm = 0;
                for (int i = 0; i < SDF_CPT; ++i) nm += __popc(s.needed[tid * SDF_CPT + i]);
                nm = __reduce_add_sync(0xffffffffu, nm);
                if (lane == 0) atomicAdd(&a.stats[b * 32 + 8], nm);
            }
            total = block_scan_1024(cnt, excl, s.scan_warp);
#pragma unroll
            for (int i = 0; i < SDF_CPT; ++i) s.coloff[tid * SDF_CPT + i] = (uint16_t)excl[i];
            SDF_STAT(4)
        }
        const bool multi = total > PHI_CAP;      // more voxels than one pass holds: finished values go to the spill area
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
            __syncthreads();
            // ---- what is known about each voxel from the previous iteration (thread per voxel): the face that was nearest
            //      (a seed: its exact distance is an upper bound of the minimum), or — static grid — the finished distance
            bool any_todo = true;
            int nhit = 0;
            {
                bool mytodo = false, myblank = false;
                for (int i = tid; i < nvox; i += SDF_THREADS) {
                    uint32_t state = HINT_NONE;
                    const uint32_t q = s.worklist[i];
                    if (hint) {
                        const int hi_ = hint_index(q);
                        const uint32_t e = hint[hi_];
                        if (((e >> 11) & 15u) == hint_tag(q) && (e & 2047u)) {
                            if (stat && (e & 0x8000u)) { state = HINT_DONE; s.best[i] = __float_as_uint(phic[hi_]); ++nhit; }
                            else state = (e & 2047u) - 1u;
                        }
                    }
                    if (state < HINT_NONE && geom_ready) s.best[i] = __float_as_uint(voxel_face_dist2(s, cl_tri, q, (int)state));
                    s.hintw[i] = state;
                    mytodo = mytodo || state != HINT_DONE;
                    myblank = myblank || state == HINT_NONE;
                }
                // (without a static grid the geometry is always there and every voxel is to do)
                if (stat) {
                    any_todo = __syncthreads_or(mytodo);
                    if (any_todo && !geom_ready) {
                        // (the seeds above had no geometry to be measured against yet)
                        geometry();
                        for (int i = tid; i < nvox; i += SDF_THREADS) {
                            const uint32_t st_ = s.hintw[i];
                            if (st_ < HINT_NONE) s.best[i] = __float_as_uint(voxel_face_dist2(s, cl_tri, s.worklist[i], (int)st_));
                        }
                    }
                }
                // ---- voxels without a seed (first iteration, stateless calls, newly touched voxels).
                //   (S) warp per voxel: distances to the NCL cluster boxes -> nearest clusters -> their face boxes -> seed
                //   then thread per voxel: the seed's exact distance
                if (__syncthreads_or(myblank)) {
                    for (int v = warp; v < nvox; v += SDF_WARPS) {
                        if (s.hintw[v] != HINT_NONE) continue;
                        const uint32_t q = s.worklist[v];
                        const uint32_t d0 = qbox_4d2(s.cl_box[lane], q);
                        const uint32_t d1 = (lane + 32 < NCL) ? qbox_4d2(s.cl_box[lane + 32], q) : 0xffffffffu;
                        const uint32_t dmin = __reduce_min_sync(0xffffffffu, min(d0, d1));
                        // A voxel near the surface lies inside several cluster and face boxes (distance 0), so the seed
                        // is chosen among the faces of all nearest clusters (up to 4) by the distance to the FARTHEST
                        // corner of the face box: an upper bound of the distance to the face that ranks small faces
                        // like the distance to their centre.
                        uint32_t tie0 = __ballot_sync(0xffffffffu, d0 == dmin), tie1 = __ballot_sync(0xffffffffu, d1 == dmin);
                        uint32_t kf = 0xffffffffu;
                        for (int n = 0; n < 4 && (tie0 | tie1); ++n) {
                            int c;
                            if (tie0) { c = __ffs(tie0) - 1; tie0 &= tie0 - 1u; }
                            else { c = __ffs(tie1) + 31; tie1 &= tie1 - 1u; }
                            const int slot = c * 32 + lane;
                            const uint2 fb = s.fbox[slot];
                            const uint32_t mx = __vmaxu4(__vabsdiffu4(q, fb.x), __vabsdiffu4(q, fb.y));
                            const uint32_t kk = (__dp4a(mx, mx, 0u) << 11) | (uint32_t)slot;
                            kf = min(kf, (fb.x >> 24) ? 0xffffffffu : kk);          // empty table slots never win
                        }
                        kf = __reduce_min_sync(0xffffffffu, kf);
                        __syncwarp();
                        if (lane == 0) s.hintw[v] = SEED_NEW | (kf & 2047u);
                    }
                    __syncthreads();
                    for (int i = tid; i < nvox; i += SDF_THREADS) {
                        const uint32_t st_ = s.hintw[i];
                        if ((st_ & 0xc0000000u) == SEED_NEW) {
                            const uint32_t slot = st_ & 2047u;
                            s.best[i] = __float_as_uint(voxel_face_dist2(s, cl_tri, s.worklist[i], (int)slot));
                            s.hintw[i] = slot;
                        }
                    }
                    __syncthreads();
                }
            }
            if (kStats) {
                nhit = __reduce_add_sync(0xffffffffu, nhit);
                if (lane == 0 && nhit) atomicAdd(&a.stats[b * 32 + 13], nhit);
                if (tid == 0) { atomicAdd(&a.stats[b * 32 + 2 * h], nvox); atomicAdd(&a.stats[b * 32 + 10], 1); }
            }
            SDF_STAT(5)
            // ---- nearest face of every voxel, one warp per voxel (round-robin), no block barrier:
            //   (B) clusters whose box is closer than best (the cluster box is the union of the face boxes, so it is
            //       never farther than any of them) -> their faces whose box is closer than best -> (voxel, face)
            //       candidates in the warp's queue segment.  A face that fails either test cannot be nearer than best.
            //   (C) lane per candidate: exact point-triangle test, atomicMin into the voxel; when the warp has no
            //       voxels left, or earlier when the segment cannot take 32 more entries.
            if (any_todo) {
                uint32_t* wqueue = s.queue + warp * QSEG;      // candidates of this warp's voxels: no atomics
                int wq = 0, ncand = 0;                           // fill (warp-uniform), candidates found
                for (int v = warp; v < nvox; v += SDF_WARPS) {
                    const int seed = (int)s.hintw[v];
                    if (stat && seed == (int)HINT_DONE) continue;
                    const uint32_t q = s.worklist[v];
                    const float bv = __uint_as_float(s.best[v]);
                    // integer threshold: (float)d * Q4D2_TO_D2 < bv  <=>  d < bv / Q4D2_TO_D2 (a power of two: exact)
                    const uint32_t thr = (uint32_t)fminf(ceilf(bv * (1.0f / Q4D2_TO_D2)), 4.0e9f);
                    const uint32_t m0 = __ballot_sync(0xffffffffu, qbox_4d2(s.cl_box[lane], q) < thr);
                    const uint32_t m1 = __ballot_sync(0xffffffffu, lane + 32 < NCL && qbox_4d2(s.cl_box[min(lane + 32, NCL - 1)], q) < thr);
                    if (kStats && lane == 0) atomicAdd(&a.stats[b * 32 + 6], __popc(m0) + __popc(m1));
                    const uint32_t ventry = ((uint32_t)v << 16) | (uint32_t)lane;
                    auto cluster = [&](int c) {
                        if (wq > QSEG - 32) {                  // warp-uniform
                            pair_flush_slow(s, cl_tri, wqueue, wq);
                            wq = 0;
                            if (kStats && lane == 0) atomicAdd(&a.stats[b * 32 + 12], 1);
                        }
                        const int slot = c * 32 + lane;
                        const uint2 fb = s.fbox[slot];
                        const bool ok = qbox_4d2(fb, q) < thr && slot != seed && fb.x < (1u << 24);      // (not an empty table slot)
                        const uint32_t okm = __ballot_sync(0xffffffffu, ok);
                        const int pos = wq + __popc(okm & lt_mask);
                        if (ok) wqueue[pos] = ventry + ((uint32_t)c << 5);
                        wq += __popc(okm);
                        if (kStats) ncand += __popc(okm);
                    };
                    for (uint32_t mm = m0; mm; mm &= mm - 1u) cluster(__ffs(mm) - 1);
                    for (uint32_t mm = m1; mm; mm &= mm - 1u) cluster(__ffs(mm) + 31);
                }
                pair_flush(s, cl_tri, wqueue, wq);
                if (kStats && lane == 0) atomicAdd(&a.stats[b * 32 + 7], ncand);
            }
            __syncthreads();
            SDF_STAT(7)
            // ---- phi = distance
            for (int i = tid; i < nvox; i += SDF_THREADS) {
                float d = __uint_as_float(s.best[i]);
                if (!stat || s.hintw[i] != HINT_DONE) {              // (known voxels already hold the distance)
                    d = sqrtf(d);
                    s.best[i] = __float_as_uint(d);
                    if (hint) {
                        const uint32_t q = s.worklist[i];
                        const int hi_ = hint_index(q);
                        hint[hi_] = (uint16_t)((stat ? 0x8000u : 0u) | (hint_tag(q) << 11) | (s.hintw[i] + 1u));
                        if (stat) phic[hi_] = d;
                    }
                }
                if (multi) spill[pass0 + i] = d;
            }
            __syncthreads();
            SDF_STAT(8)
        }

        // ---- trilinear sample + gradient (grid_sampler_3d fwd/bwd, align_corners=False, zeros) and the
        //      per-vertex outputs of the query hand o
        const size_t ov0 = (size_t)b * (2 * NV) + o * NV;
        const bool want_shift = a.gshift && o == 1;
        if (total == 0) {          // no voxel of the grid hand is both touched and inside: exact zeros
            const float rho0 = 0.f;                     // rho(0) = 0 with and without the robustifier
            for (int v = tid; v < NV; v += SDF_THREADS) {
                if (a.per_vert) a.per_vert[ov0 + v] = rho0;
                if (a.origin) a.origin[ov0 + v] = 0.f;
            }
            if (a.gzero) { if (tid == 0) a.gzero[b * 2 + o] = 1; }
            else if (a.gverts) { float* gp = a.gverts + ov0 * 3; for (int i = tid; i < NV * 3; i += SDF_THREADS) gp[i] = 0.f; }
            if (tid == 0) w.parts[b * 2 + h] = 0.f;
            if (want_shift && tid < 3) a.gshift[(size_t)b * 3 + tid] = 0.f;
            SDF_STAT(9)
            continue;
        }
        // d psi / d vertex = (G/2) * d psi / d(ix) / scale ; loss = sum(rho) / 4
        const float kbase = mask * a.grad_scale * 0.25f * (0.5f * G) / scale;
        float sums[4] = {0.f, 0.f, 0.f, 0.f};      // sum of rho, gradient sum xyz
        SDF_SLOT_LOOP
        for (int sl = 0; sl < SDF_SLOTS; ++sl) {
            const int v = tid + sl * SDF_THREADS;
            if (v >= NV) continue;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            float fr[3], pv[3];
            int i0[3];
            load_vert(o, v, pv);
            if (locate(pv, fr, i0)) {
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
                            const float val = multi ? spill[idx] : __uint_as_float(s.best[idx]);
                            const float wx = dx ? tx : 1.0f - tx, wy = dy ? ty : 1.0f - ty, wz = dz ? tz : 1.0f - tz;
                            acc[0] += val * wx * wy * wz;
                            acc[1] += (dx ? val : -val) * wy * wz;
                            acc[2] += (dy ? val : -val) * wx * wz;
                            acc[3] += (dz ? val : -val) * wx * wy;
                        }
            }
            const float psi = acc[0];
            float rho = psi, drho = 1.0f;
            if (a.robustifier > 0.f) {
                const float t = psi / a.robustifier, frac = t * t;
                rho = frac / (frac + 1.0f);
                drho = 2.0f * t / a.robustifier / ((frac + 1.0f) * (frac + 1.0f));
            }
            sums[0] += rho;
            if (a.per_vert) a.per_vert[ov0 + v] = rho;
            if (a.origin) a.origin[ov0 + v] = psi * scale;
            if (a.gverts || a.gshift) {
                const float kk = kbase * drho;
                float g[3] = {kk * acc[1], kk * acc[2], kk * acc[3]};
                from_ray_frame(g, a.ray_axis);
                sums[1] += g[0]; sums[2] += g[1]; sums[3] += g[2];
                if (a.gverts) {
                    if (xform && o == 1) g[0] = -g[0];
                    float* gp = a.gverts + (ov0 + v) * 3;
                    gp[0] = g[0]; gp[1] = g[1]; gp[2] = g[2];
                }
            }
        }
        block_sum4(sums, want_shift ? 4 : 1, s.red);
        if (tid == 0) w.parts[b * 2 + h] = sums[0];
        if (a.gzero && tid == 0) a.gzero[b * 2 + o] = 0;
        if (want_shift && tid >= 1 && tid < 4) a.gshift[(size_t)b * 3 + (tid - 1)] = sums[tid];
        SDF_STAT(9)
    }
}

// =========================================================================================================
// EXACT (grid-free) penetration mode — SURVEY.md §8(f) rank 3.  NOT the reference's function: a separately named
// alternative to the 32^3 field.  For every query vertex p (normalised by the grid hand's box as in Appendix B):
//     psi = dist(p, mesh_h) if p is inside mesh_h (odd number of +x ray crossings) else 0
// i.e. the limit of the reference's field for an infinitely fine grid: no 7 mm voxel quantisation, and per direction
// at most 778 inside tests and distance searches instead of up to 8 voxels per vertex.  Same work list, boxes and
// loss conventions as the grid mode (loss = sum psi / 4, origin_scale = psi * scale, gradient to the query vertex only).
// One CTA per (frame, direction); a warp owns a vertex for the inside test and for the nearest-face search.
struct __align__(16) SdfExactSmem {
    float U[NV * 3];
    uint2 cl_box[NCL];
    uint2 fbox[NCL * 32];
    float4 act[NV];             // active query vertices: normalised position, vertex id (as bits)
    float4 outv[NV];            // per query vertex: psi, gradient direction d psi / d p
    float red[4 * SDF_WARPS];
    int nact, item;
};

// closest point of triangle (a,b,c) to p (Ericson, Real-Time Collision Detection 5.1.5)
__device__ __forceinline__ void pt_tri_closest(const float* p, const float* a, const float* b, const float* c, float* out) {
    float ab[3], ac[3], ap[3], bp[3], cp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; bp[k] = p[k] - b[k]; cp[k] = p[k] - c[k]; }
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap), d3 = dot3(ab, bp), d4 = dot3(ac, bp), d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    const float vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    float u = 0.f, v = 0.f;                      // closest = a + u ab + v ac
    if (d1 <= 0.f && d2 <= 0.f) { u = 0.f; v = 0.f; }
    else if (d3 >= 0.f && d4 <= d3) { u = 1.f; v = 0.f; }
    else if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { u = d1 / (d1 - d3); v = 0.f; }
    else if (d6 >= 0.f && d5 <= d6) { u = 0.f; v = 1.f; }
    else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { u = 0.f; v = d2 / (d2 - d6); }
    else if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) { v = (d4 - d3) / ((d4 - d3) + (d5 - d6)); u = 1.f - v; }
    else { const float den = 1.0f / (va + vb + vc); u = vb * den; v = vc * den; }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = a[k] + u * ab[k] + v * ac[k];
}

// squared distance (normalised units) from a point given in Q8 units (float) to a packed box: a lower bound of the
// distance to anything inside the box (the boxes are quantised outwards)
__device__ __forceinline__ float fbox_dist2(uint2 bx, const float* pq) {
    float s = 0.f;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float lo = (float)((bx.x >> (8 * ax)) & 255u), hi = (float)((bx.y >> (8 * ax)) & 255u);
        const float d = fmaxf(fmaxf(lo - pq[ax], pq[ax] - hi), 0.f);
        s += d * d;
    }
    return s * Q8_TO_D2;
}

__global__ void __launch_bounds__(SDF_THREADS, 4)
k_sdf_exact(int B, SdfArgs a, SdfWs w, const ushort4* __restrict__ cl_r, const ushort4* __restrict__ cl_l) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SdfExactSmem& s = *reinterpret_cast<SdfExactSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool xform = (a.joints != nullptr);
    int bucket_end[SDF_BUCKETS];
    {
        int run = 0;
#pragma unroll
        for (int k = 0; k < SDF_BUCKETS; ++k) { run += (int)w.counters[k]; bucket_end[k] = run; }
    }
    const int n_items = bucket_end[SDF_BUCKETS - 1];
    for (;;) {
        __syncthreads();
        if (tid == 0) { s.item = (int)atomicAdd(&w.counters[SDF_BUCKETS], 1u); s.nact = 0; }
        __syncthreads();
        if (s.item >= n_items) break;
        uint32_t code;
        {
            const int t = s.item;
            int k = 0, start = 0;
#pragma unroll
            for (int q = 0; q + 1 < SDF_BUCKETS; ++q) if (t >= bucket_end[q]) { k = q + 1; start = bucket_end[q]; }
            code = w.items[(size_t)k * 2 * B + (t - start)];
        }
        const int b = (int)(code >> 1), h = (int)(code & 1u), o = 1 - h;
        const ushort4* cl_tri = h ? cl_l : cl_r;
        const float* hd = w.hdr + (size_t)b * SDF_HDR;
        float cen[3], tlo[3], thi[3];
        const float scale = hd[16 * h + 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { cen[c] = hd[16 * h + c]; tlo[c] = hd[16 * h + 4 + c]; thi[c] = hd[16 * h + 7 + c]; }
        const float shx = hd[32], shy = hd[33], shz = hd[34], mask = hd[35];
        auto load_vert = [&](int hand, int v, float* out) {
            const float* p = a.verts + (((size_t)b * 2 + hand) * NV + v) * 3;
            out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
            if (xform && hand == 1) { out[0] = -out[0] + shx; out[1] += shy; out[2] += shz; }
            to_ray_frame(out, a.ray_axis);
        };
        // ---- query vertices that can be inside at all: within the mesh's (y,z) extent and not right of its largest x
        for (int v = tid; v < NV; v += SDF_THREADS) {
            float p[3], pn[3];
            load_vert(o, v, p);
#pragma unroll
            for (int c = 0; c < 3; ++c) pn[c] = (p[c] - cen[c]) / scale;
            s.outv[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pn[0] <= thi[0] && pn[1] >= tlo[1] && pn[1] <= thi[1] && pn[2] >= tlo[2] && pn[2] <= thi[2])
                s.act[atomicAdd(&s.nact, 1)] = make_float4(pn[0], pn[1], pn[2], __int_as_float(v));
        }
        __syncthreads();
        const int nact = s.nact;
        if (nact > 0) {
            // ---- grid-hand geometry in normalised coordinates: vertices, Q8 face boxes, cluster boxes (as in k_sdf_dir)
            for (int v = tid; v < NV; v += SDF_THREADS) {
                float p[3];
                load_vert(h, v, p);
#pragma unroll
                for (int c = 0; c < 3; ++c) s.U[v * 3 + c] = (p[c] - cen[c]) / scale;
            }
            __syncthreads();
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
                const uint32_t far = 255u << 24;
                s.fbox[c * 32 + lane] = id.w ? make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 8) | ((uint32_t)lo[2] << 16),
                                                          (uint32_t)hi[0] | ((uint32_t)hi[1] << 8) | ((uint32_t)hi[2] << 16))
                                             : make_uint2(far, far);
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) { lo[ax] = __reduce_min_sync(0xffffffffu, lo[ax]); hi[ax] = __reduce_max_sync(0xffffffffu, hi[ax]); }
                if (lane == 0)
                    s.cl_box[c] = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 8) | ((uint32_t)lo[2] << 16),
                                             (uint32_t)hi[0] | ((uint32_t)hi[1] << 8) | ((uint32_t)hi[2] << 16));
            }
            __syncthreads();
            // ---- one warp per active vertex: inside test, then (inside only) the exact nearest face
            for (int i = warp; i < nact; i += SDF_WARPS) {
                const float4 av = s.act[i];
                const float p[3] = {av.x, av.y, av.z};
                const int v = __float_as_int(av.w);
                const float pq[3] = {(p[0] + 1.0f) * 128.0f, (p[1] + 1.0f) * 128.0f, (p[2] + 1.0f) * 128.0f};     // Q8 units
                // clusters / faces whose (y,z) box contains the ray and that reach beyond p in x (boxes are quantised
                // outwards by at least 1e-3 Q8 units; 1e-2 more absorbs the rounding of pq)
                auto ray_box = [&](uint2 bx) {
                    return pq[1] >= (float)((bx.x >> 8) & 255u) - 1e-2f && pq[1] <= (float)((bx.y >> 8) & 255u) + 1e-2f &&
                           pq[2] >= (float)((bx.x >> 16) & 255u) - 1e-2f && pq[2] <= (float)((bx.y >> 16) & 255u) + 1e-2f &&
                           pq[0] <= (float)(bx.y & 255u) + 1e-2f && (bx.x >> 24) == 0u;
                };
                int crossings = 0;
                uint32_t m0 = __ballot_sync(0xffffffffu, ray_box(s.cl_box[lane]));
                uint32_t m1 = __ballot_sync(0xffffffffu, lane + 32 < NCL && ray_box(s.cl_box[min(lane + 32, NCL - 1)]));
                for (int half = 0; half < 2; ++half) {
                    for (uint32_t mm = half ? m1 : m0; mm; mm &= mm - 1u) {
                        const int slot = (__ffs(mm) - 1 + 32 * half) * 32 + lane;
                        if (ray_box(s.fbox[slot])) {
                            const ushort4 id = cl_tri[slot];
                            float x;
                            if (ray_hit(s.U, id.x, id.y, id.z, p[1], p[2], x) && x > p[0]) ++crossings;
                        }
                    }
                }
                crossings = __reduce_add_sync(0xffffffffu, crossings);
                if (!(crossings & 1)) continue;                    // warp-uniform: outside
                // nearest face: clusters nearest box first, one face per lane, until no unvisited box can be closer
                float lb[2];
                lb[0] = fbox_dist2(s.cl_box[lane], pq);
                lb[1] = (lane + 32 < NCL) ? fbox_dist2(s.cl_box[lane + 32], pq) : 3e30f;
                float best = 3e30f;
                int best_slot = 0;
                for (int it = 0; it < NCL; ++it) {
                    float mlb = fminf(lb[0], lb[1]);
                    int which = (lb[0] <= lb[1]) ? lane : lane + 32;
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1) {
                        const float m2 = __shfl_xor_sync(0xffffffffu, mlb, sft);
                        const int w2 = __shfl_xor_sync(0xffffffffu, which, sft);
                        if (m2 < mlb || (m2 == mlb && w2 < which)) { mlb = m2; which = w2; }
                    }
                    if (mlb >= best) break;
                    const int slot = which * 32 + lane;
                    const uint2 fb = s.fbox[slot];
                    float d2 = 3e30f;
                    if ((fb.x >> 24) == 0u && fbox_dist2(fb, pq) < best) {
                        const ushort4 id = cl_tri[slot];
                        d2 = pt_tri_dist2(p, s.U + 3 * id.x, s.U + 3 * id.y, s.U + 3 * id.z);
                    }
                    // warp argmin (ties: lowest slot), folded into the running best
                    float dm = d2;
                    int sm = slot;
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1) {
                        const float d3 = __shfl_xor_sync(0xffffffffu, dm, sft);
                        const int s3 = __shfl_xor_sync(0xffffffffu, sm, sft);
                        if (d3 < dm || (d3 == dm && s3 < sm)) { dm = d3; sm = s3; }
                    }
                    if (dm < best) { best = dm; best_slot = sm; }
                    if (which == lane) lb[0] = 3e30f;
                    if (which == lane + 32) lb[1] = 3e30f;
                }
                if (lane == 0) {
                    const ushort4 id = cl_tri[best_slot];
                    float cpt[3];
                    pt_tri_closest(p, s.U + 3 * id.x, s.U + 3 * id.y, s.U + 3 * id.z, cpt);
                    const float d[3] = {p[0] - cpt[0], p[1] - cpt[1], p[2] - cpt[2]};
                    const float psi = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    const float inv = psi > 0.f ? 1.0f / psi : 0.f;
                    s.outv[v] = make_float4(psi, d[0] * inv, d[1] * inv, d[2] * inv);
                }
            }
            __syncthreads();
        }
        // ---- per-vertex outputs of the query hand o
        const size_t ov0 = (size_t)b * (2 * NV) + o * NV;
        const bool want_shift = a.gshift && o == 1;
        const float kbase = mask * a.grad_scale * 0.25f / scale;       // d psi / d vertex = d psi / d p / scale ; loss = sum(psi) / 4
        float sums[4] = {0.f, 0.f, 0.f, 0.f};
        for (int v = tid; v < NV; v += SDF_THREADS) {
            const float4 r = s.outv[v];
            sums[0] += r.x;
            if (a.per_vert) a.per_vert[ov0 + v] = r.x;
            if (a.origin) a.origin[ov0 + v] = r.x * scale;
            float g[3] = {kbase * r.y, kbase * r.z, kbase * r.w};
            from_ray_frame(g, a.ray_axis);
            sums[1] += g[0]; sums[2] += g[1]; sums[3] += g[2];
            if (a.gverts) {
                if (xform && o == 1) g[0] = -g[0];
                float* gp = a.gverts + (ov0 + v) * 3;
                gp[0] = g[0]; gp[1] = g[1]; gp[2] = g[2];
            }
        }
        block_sum4(sums, want_shift ? 4 : 1, s.red);
        if (tid == 0) w.parts[b * 2 + h] = sums[0];
        if (want_shift && tid >= 1 && tid < 4) a.gshift[(size_t)b * 3 + (tid - 1)] = sums[tid];
    }
}

int launch_sdf_exact(const ihmr_model* m, int B, const SdfArgs& a, cudaStream_t st) {
    if (B <= 0) return IHMR_OK;
    if (!a.ws) { set_error("sdf_exact: no workspace"); return IHMR_E_INVALID; }
    static unsigned long long configured = 0ull;
    static int ctas_per_sm = 0;
    if (int rc = ensure_dynamic_smem(k_sdf_exact, sizeof(SdfExactSmem), configured)) return rc;
    if (ctas_per_sm == 0) {
        int n = 0;
        IHMR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_sdf_exact, SDF_THREADS, sizeof(SdfExactSmem)));
        if (n < 1) { set_error("sdf_exact: the kernel does not fit on an SM"); return IHMR_E_CUDA; }
        ctas_per_sm = n;
    }
    const SdfWs w = sdf_ws_carve(a.ws, B);
    IHMR_CUDA_OK(cudaMemsetAsync(w.counters, 0, (SDF_BUCKETS + 1) * 4, st));
    SdfArgs pa = a;
    pa.gzero = nullptr;                 // (the skipped directions get explicit zeros in this mode)
    pa.box_scale = m->sdf_box_scale; pa.ray_axis = m->sdf_ray_axis;
    k_sdf_prep<<<(B + PREP_WARPS - 1) / PREP_WARPS, PREP_WARPS * 32, 0, st>>>(B, pa, w);
    IHMR_LAUNCH_OK();
    const int grid = std::min(std::min(m->num_sms * ctas_per_sm, SDF_MAX_GRID), 2 * B);
    k_sdf_exact<<<grid, SDF_THREADS, sizeof(SdfExactSmem), st>>>(B, pa, w, reinterpret_cast<const ushort4*>(m->cl_tri[0]),
                                                               reinterpret_cast<const ushort4*>(m->cl_tri[1]));
    IHMR_LAUNCH_OK();
    if (a.losses) {
        k_sdf_finish<<<(B + 255) / 256, 256, 0, st>>>(B, w.parts, a.hand_type, a.losses);
        IHMR_LAUNCH_OK();
    }
    return IHMR_OK;
}

int launch_sdf(const ihmr_model* m, int B, const SdfArgs& a, cudaStream_t st) {
    if (B <= 0) return IHMR_OK;
    NvtxRange range("ihmr_sdf");
    if (!a.ws) { set_error("sdf: no workspace"); return IHMR_E_INVALID; }
    static unsigned long long configured[3] = {0ull, 0ull, 0ull};
    static int ctas_per_sm[3] = {0, 0, 0};
    const bool use_static = a.static_grid_mask && a.pcache && a.phic && a.hints;
    const int kv = a.stats ? 2 : (use_static ? 1 : 0);
    auto kernel = a.stats ? k_sdf_dir<false, true> : (use_static ? k_sdf_dir<true, false> : k_sdf_dir<false, false>);
    if (int rc = ensure_dynamic_smem(kernel, sizeof(SdfSmem), configured[kv])) return rc;
    if (ctas_per_sm[kv] == 0) {
        int n = 0;
        IHMR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, SDF_THREADS, sizeof(SdfSmem)));
        if (n < 1) { set_error("sdf: the kernel does not fit on an SM"); return IHMR_E_CUDA; }
        ctas_per_sm[kv] = n;
    }
    const SdfWs w = sdf_ws_carve(a.ws, B);
    IHMR_CUDA_OK(cudaMemsetAsync(w.counters, 0, (SDF_BUCKETS + 1) * 4, st));
    SdfArgs pa = a;
    pa.box_scale = m->sdf_box_scale; pa.ray_axis = m->sdf_ray_axis;       // conventions A2 / A4 of the model
    pa.l2_prefetch = (reinterpret_cast<uintptr_t>(a.verts) & 15u) == 0;   // the bulk prefetch wants 16-byte aligned frames
    k_sdf_prep<<<(B + PREP_WARPS - 1) / PREP_WARPS, PREP_WARPS * 32, 0, st>>>(B, pa, w);
    IHMR_LAUNCH_OK();
    const int grid = std::min(std::min(m->num_sms * ctas_per_sm[kv], SDF_MAX_GRID), 2 * B);
    kernel<<<grid, SDF_THREADS, sizeof(SdfSmem), st>>>(B, pa, w, reinterpret_cast<const ushort4*>(m->cl_tri[0]),
                                                       reinterpret_cast<const ushort4*>(m->cl_tri[1]));
    IHMR_LAUNCH_OK();
    if (a.losses) {
        k_sdf_finish<<<(B + 255) / 256, 256, 0, st>>>(B, w.parts, a.hand_type, a.losses);
        IHMR_LAUNCH_OK();
    }
    return IHMR_OK;
}

size_t sdf_pcache_bytes(int B) { return (size_t)B * SDF_PCACHE * sizeof(uint32_t); }
size_t sdf_phic_bytes(int B) { return (size_t)B * SDF_HINTS * sizeof(float); }
size_t sdf_hint_bytes(int B) { return (size_t)B * 2 * SDF_HINTS * sizeof(uint16_t); }

const float* sdf_ws_parts(void* ws, int B) { return sdf_ws_carve(ws, B).parts; }

}  // namespace ihmr
