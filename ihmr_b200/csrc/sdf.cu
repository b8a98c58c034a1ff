// Left/right interpenetration loss, forward + backward in one kernel (SURVEY.md §8 a10,
// Appendix B).  One CTA per frame.  For each direction (grid hand h, query hand o = 1-h):
//
//   1. bounding box of h -> centre c_h, scale s_h = 0.6 * max extent           (A2)
//   2. mark the <= 8 voxel corners each query vertex touches                    (lazy grid)
//   3. inside/outside of the marked voxel columns: every face is rasterised onto the 32x32
//      (y,z) lattice of +x rays; a hit toggles the bits of all voxels left of the crossing   (A4)
//   4. faces are binned by the (y,z) lattice cells their bounding box overlaps
//   5. phi = min point-triangle distance for every voxel that is both marked and inside,
//      searching lattice rings outwards until the ring's lower bound exceeds the best
//   6. trilinear sampling with grid_sample(align_corners=False, zeros) semantics, its
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
constexpr int SDF_THREADS = 256;
constexpr int SDF_SLOTS = 4;        // 4 x 256 >= 778 query vertices
constexpr int PHI_CAP = 4096;       // voxels evaluated per pass
constexpr int BIN_CAP = 6144;       // (face, lattice cell) pairs

struct __align__(16) SdfSmem {
    float U[NV * 3];
    uint32_t needed[G * G];     // marked voxels per (z,y) column; later bin counters/cursors
    uint32_t work[G * G];       // parity bits, then marked & inside
    uint16_t coloff[G * G];     // exclusive prefix of popc(work)
    uint16_t bin_start[G * G + 2];
    uint16_t bin_entries[BIN_CAP];
    uint16_t worklist[PHI_CAP];
    float phi[PHI_CAP];
    float red[64];
    float box[2][2][3];         // [hand][lo/hi][xyz]
    float shift[4];
    int scan_warp[8];
    int scalars[4];
};

__device__ __forceinline__ float voxel_center(int i) { return (2.0f * i + 1.0f - G) / G; }

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

// squared distance point -> triangle (closest-point regions)
__device__ __forceinline__ float pt_tri_dist2(const float* p, const float* a, const float* b, const float* c) {
    float ab[3], ac[3], ap[3], bp[3], cp[3], cl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return dot3(ap, ap);
#pragma unroll
    for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return dot3(bp, bp);
#pragma unroll
    for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return dot3(cp, cp);
    const float vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float t = d1 / (d1 - d3);
#pragma unroll
        for (int k = 0; k < 3; ++k) cl[k] = a[k] + t * ab[k];
    } else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float t = d2 / (d2 - d6);
#pragma unroll
        for (int k = 0; k < 3; ++k) cl[k] = a[k] + t * ac[k];
    } else if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
#pragma unroll
        for (int k = 0; k < 3; ++k) cl[k] = b[k] + t * (c[k] - b[k]);
    } else {
        const float den = va + vb + vc;
        if (den == 0.f) return fminf(dot3(ap, ap), fminf(dot3(bp, bp), dot3(cp, cp)));
        const float v = vb / den, w = vc / den;
#pragma unroll
        for (int k = 0; k < 3; ++k) cl[k] = a[k] + v * ab[k] + w * ac[k];
    }
    float d[3] = {p[0] - cl[0], p[1] - cl[1], p[2] - cl[2]};
    return dot3(d, d);
}

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
        if (lane == 0) red[i * 8 + warp] = s;
    }
    __syncthreads();
    for (int i = 0; i < nval; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < SDF_THREADS / 32; ++w) s += red[i * 8 + w];
        v[i] = s;
    }
    __syncthreads();
}

// exclusive scan of 1024 counts, 4 consecutive entries per thread; returns total
__device__ __forceinline__ int block_scan_1024(const int (&cnt)[4], int (&excl)[4], int* scan_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int local = cnt[0] + cnt[1] + cnt[2] + cnt[3];
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
    for (int w = 0; w < SDF_THREADS / 32; ++w) {
        if (w < warp) base += scan_warp[w];
        total += scan_warp[w];
    }
    int run = base + inc - local;
#pragma unroll
    for (int i = 0; i < 4; ++i) { excl[i] = run; run += cnt[i]; }
    __syncthreads();
    return total;
}

__device__ __forceinline__ int lattice_cell(float y) {   // cell of width 2/G centred on a voxel centre
    int c = (int)floorf((y + 1.0f) * (0.5f * G));
    return min(G - 1, max(0, c));
}

// phi of one voxel: min distance to the mesh, ring search over the lattice bins
__device__ float eval_voxel(const SdfSmem& s, const uint16_t* __restrict__ faces, int code, bool use_bins) {
    const int x = code & 31, col = code >> 5, j = col & 31, k = col >> 5;
    const float q[3] = {voxel_center(x), voxel_center(j), voxel_center(k)};
    float best = 1e30f;
    if (use_bins) {
        const float h = 2.0f / G;
        for (int ring = 0; ring < G; ++ring) {
            if (ring > 0) {
                const float lb = (ring - 0.5f) * h - 1e-5f;
                if (lb * lb >= best) break;
            }
            for (int kk = k - ring; kk <= k + ring; ++kk) {
                if (kk < 0 || kk >= G) continue;
                const bool edge_row = (kk == k - ring) || (kk == k + ring);
                const int step = edge_row ? 1 : max(1, 2 * ring);
                for (int jj = j - ring; jj <= j + ring; jj += step) {
                    if (jj < 0 || jj >= G) continue;
                    const int c = kk * G + jj;
                    for (int e = s.bin_start[c]; e < s.bin_start[c + 1]; ++e) {
                        const int f = s.bin_entries[e];
                        const ushort4 id = reinterpret_cast<const ushort4*>(faces)[f];
                        best = fminf(best, pt_tri_dist2(q, s.U + 3 * id.x, s.U + 3 * id.y, s.U + 3 * id.z));
                    }
                }
            }
        }
    } else {
        for (int f = 0; f < NF; ++f) {
            const ushort4 id = reinterpret_cast<const ushort4*>(faces)[f];
            best = fminf(best, pt_tri_dist2(q, s.U + 3 * id.x, s.U + 3 * id.y, s.U + 3 * id.z));
        }
    }
    return sqrtf(best);
}

__global__ void __launch_bounds__(SDF_THREADS)
k_sdf(int B, SdfArgs a, const uint16_t* __restrict__ faces_r, const uint16_t* __restrict__ faces_l) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SdfSmem& s = *reinterpret_cast<SdfSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const bool xform = (a.joints != nullptr);

    if (tid == 0) {
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
    auto load_vert = [&](int hand, int v, float* out) {
        const float* p = a.verts + (((size_t)b * 2 + hand) * NV + v) * 3;
        out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
        if (xform && hand == 1) { out[0] = -out[0] + shx; out[1] += shy; out[2] += shz; }
    };

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
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int hnd = 0; hnd < 2; ++hnd)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float l = lo[hnd][c], hgh = hi[hnd][c];
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
                    hgh = fmaxf(hgh, __shfl_xor_sync(0xffffffffu, hgh, o));
                }
                if (lane == 0) { s.phi[(hnd * 3 + c) * 16 + warp] = l; s.phi[(hnd * 3 + c) * 16 + 8 + warp] = hgh; }
            }
        __syncthreads();
        if (tid < 6) {
            float l = 1e30f, hgh = -1e30f;
            for (int w = 0; w < SDF_THREADS / 32; ++w) { l = fminf(l, s.phi[tid * 16 + w]); hgh = fmaxf(hgh, s.phi[tid * 16 + 8 + w]); }
            s.box[tid / 3][0][tid % 3] = l;
            s.box[tid / 3][1][tid % 3] = hgh;
        }
        __syncthreads();
    }

    float mask = 1.0f;
    if (a.hand_type) mask = (a.hand_type[b * 2] + a.hand_type[b * 2 + 1] > 1.5f) ? 1.0f : 0.0f;
    float loss_part = 0.f;

    for (int h = 0; h < 2; ++h) {
        const int o = 1 - h;
        const uint16_t* faces = h ? faces_l : faces_r;
        float cen[3], ext = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cen[c] = (s.box[h][0][c] + s.box[h][1][c]) * 0.5f;
            ext = fmaxf(ext, s.box[h][1][c] - s.box[h][0][c]);
        }
        const float scale = 0.6f * ext;      // (1 + 0.2) * 0.5 * max extent

        for (int i = tid; i < G * G; i += SDF_THREADS) { s.needed[i] = 0u; s.work[i] = 0u; }
        __syncthreads();

        // ---- query vertices: normalised position, voxel corners, mark
        float acc[SDF_SLOTS][4];
        float fr[SDF_SLOTS][3];
        int i0[SDF_SLOTS][3];
        bool act[SDF_SLOTS];
        bool any = false;
#pragma unroll
        for (int sl = 0; sl < SDF_SLOTS; ++sl) {
            const int v = tid + sl * SDF_THREADS;
            act[sl] = false;
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[sl][c] = 0.f;
            if (v < NV) {
                float p[3];
                load_vert(o, v, p);
                bool in = true;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float pn = (p[c] - cen[c]) / scale;
                    const float ix = ((pn + 1.0f) * G - 1.0f) * 0.5f;
                    const float fl = floorf(ix);
                    fr[sl][c] = ix - fl;
                    // clamp before the int conversion: far-away vertices must not overflow
                    i0[sl][c] = (int)fminf(fmaxf(fl, -2.0f), (float)G);
                    in = in && (i0[sl][c] >= -1) && (i0[sl][c] <= G - 1);
                }
                act[sl] = in;
                if (in) {
                    any = true;
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
                            const int zc = i0[sl][2] + dz, yc = i0[sl][1] + dy;
                            if (zc < 0 || zc >= G || yc < 0 || yc >= G) continue;
                            uint32_t bits = 0u;
                            if (i0[sl][0] >= 0) bits |= 1u << i0[sl][0];
                            if (i0[sl][0] + 1 < G) bits |= 1u << (i0[sl][0] + 1);
                            atomicOr(&s.needed[zc * G + yc], bits);
                        }
                }
            }
        }
        const bool any_block = __syncthreads_or(any);
        bool run = any_block;
        int total = 0;
        bool use_bins = true;

        if (run) {
            // ---- normalised grid-hand vertices
            for (int v = tid; v < NV; v += SDF_THREADS) {
                float p[3];
                load_vert(h, v, p);
#pragma unroll
                for (int c = 0; c < 3; ++c) s.U[v * 3 + c] = (p[c] - cen[c]) / scale;
            }
            __syncthreads();
            // ---- parity of the marked columns
            for (int f = tid; f < NF; f += SDF_THREADS) {
                const ushort4 id = reinterpret_cast<const ushort4*>(faces)[f];
                const float* A_ = s.U + 3 * id.x; const float* B_ = s.U + 3 * id.y; const float* C_ = s.U + 3 * id.z;
                const float ymin = fminf(A_[1], fminf(B_[1], C_[1])), ymax = fmaxf(A_[1], fmaxf(B_[1], C_[1]));
                const float zmin = fminf(A_[2], fminf(B_[2], C_[2])), zmax = fmaxf(A_[2], fmaxf(B_[2], C_[2]));
                // lattice points y_j = (2j+1-G)/G inside [ymin,ymax], one spare on each side
                const int j0 = max(0, (int)floorf((ymin * G + (G - 1)) * 0.5f)), j1 = min(G - 1, (int)ceilf((ymax * G + (G - 1)) * 0.5f));
                const int k0 = max(0, (int)floorf((zmin * G + (G - 1)) * 0.5f)), k1 = min(G - 1, (int)ceilf((zmax * G + (G - 1)) * 0.5f));
                for (int k = k0; k <= k1; ++k)
                    for (int j = j0; j <= j1; ++j) {
                        const int col = k * G + j;
                        if (s.needed[col] == 0u) continue;
                        float x;
                        if (!ray_hit(s.U, id.x, id.y, id.z, voxel_center(j), voxel_center(k), x)) continue;
                        // voxels whose centre lies strictly left of the crossing
                        int cnt = min(G, max(0, (int)ceilf((x * G + (G - 1)) * 0.5f)));
                        while (cnt < G && x > voxel_center(cnt)) ++cnt;
                        while (cnt > 0 && !(x > voxel_center(cnt - 1))) --cnt;
                        if (cnt > 0) atomicXor(&s.work[col], cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u));
                    }
            }
            __syncthreads();
            // ---- marked & inside, prefix offsets
            int cnt[4], excl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tid * 4 + i;
                const uint32_t wk = s.needed[c] & s.work[c];
                s.work[c] = wk;
                cnt[i] = __popc(wk);
            }
            total = block_scan_1024(cnt, excl, s.scan_warp);
#pragma unroll
            for (int i = 0; i < 4; ++i) s.coloff[tid * 4 + i] = (uint16_t)excl[i];
            run = total > 0;
        }
        if (run) {
            // ---- bin faces by lattice cell (needed[] is free now: counters, then cursors)
            for (int i = tid; i < G * G; i += SDF_THREADS) s.needed[i] = 0u;
            __syncthreads();
            for (int f = tid; f < NF; f += SDF_THREADS) {
                const ushort4 id = reinterpret_cast<const ushort4*>(faces)[f];
                const float* A_ = s.U + 3 * id.x; const float* B_ = s.U + 3 * id.y; const float* C_ = s.U + 3 * id.z;
                const int j0 = lattice_cell(fminf(A_[1], fminf(B_[1], C_[1]))), j1 = lattice_cell(fmaxf(A_[1], fmaxf(B_[1], C_[1])));
                const int k0 = lattice_cell(fminf(A_[2], fminf(B_[2], C_[2]))), k1 = lattice_cell(fmaxf(A_[2], fmaxf(B_[2], C_[2])));
                for (int k = k0; k <= k1; ++k)
                    for (int j = j0; j <= j1; ++j) atomicAdd(&s.needed[k * G + j], 1u);
            }
            __syncthreads();
            int cnt[4], excl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cnt[i] = (int)s.needed[tid * 4 + i];
            const int entries = block_scan_1024(cnt, excl, s.scan_warp);
            use_bins = entries <= BIN_CAP;
            if (use_bins) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { s.bin_start[tid * 4 + i] = (uint16_t)excl[i]; s.needed[tid * 4 + i] = 0u; }
                if (tid == 0) s.bin_start[G * G] = (uint16_t)entries;
                __syncthreads();
                for (int f = tid; f < NF; f += SDF_THREADS) {
                    const ushort4 id = reinterpret_cast<const ushort4*>(faces)[f];
                    const float* A_ = s.U + 3 * id.x; const float* B_ = s.U + 3 * id.y; const float* C_ = s.U + 3 * id.z;
                    const int j0 = lattice_cell(fminf(A_[1], fminf(B_[1], C_[1]))), j1 = lattice_cell(fmaxf(A_[1], fmaxf(B_[1], C_[1])));
                    const int k0 = lattice_cell(fminf(A_[2], fminf(B_[2], C_[2]))), k1 = lattice_cell(fmaxf(A_[2], fmaxf(B_[2], C_[2])));
                    for (int k = k0; k <= k1; ++k)
                        for (int j = j0; j <= j1; ++j) {
                            const int c = k * G + j;
                            const uint32_t pos = s.bin_start[c] + atomicAdd(&s.needed[c], 1u);
                            s.bin_entries[pos] = (uint16_t)f;
                        }
                }
            }
            __syncthreads();

            // ---- passes over the marked & inside voxels
            for (int pass0 = 0; pass0 < total; pass0 += PHI_CAP) {
                const int pass1 = min(total, pass0 + PHI_CAP);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = tid * 4 + i;
                    uint32_t wk = s.work[c];
                    int idx = s.coloff[c];
                    while (wk) {
                        const int x = __ffs(wk) - 1;
                        wk &= wk - 1;
                        if (idx >= pass0 && idx < pass1) s.worklist[idx - pass0] = (uint16_t)((c << 5) | x);
                        ++idx;
                    }
                }
                __syncthreads();
                for (int i = tid; i < pass1 - pass0; i += SDF_THREADS)
                    s.phi[i] = eval_voxel(s, faces, s.worklist[i], use_bins);
                __syncthreads();
                // ---- trilinear sample + gradient (grid_sampler_3d fwd/bwd, align_corners=False, zeros)
#pragma unroll
                for (int sl = 0; sl < SDF_SLOTS; ++sl) {
                    if (!act[sl]) continue;
                    const float tx = fr[sl][0], ty = fr[sl][1], tz = fr[sl][2];
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 2; ++dx) {
                                const int xc = i0[sl][0] + dx, yc = i0[sl][1] + dy, zc = i0[sl][2] + dz;
                                if (xc < 0 || xc >= G || yc < 0 || yc >= G || zc < 0 || zc >= G) continue;
                                const int c = zc * G + yc;
                                const uint32_t wk = s.work[c];
                                if (!((wk >> xc) & 1u)) continue;
                                const int idx = s.coloff[c] + __popc(wk & ((1u << xc) - 1u));
                                if (idx < pass0 || idx >= pass1) continue;
                                const float val = s.phi[idx - pass0];
                                const float wx = dx ? tx : 1.0f - tx, wy = dy ? ty : 1.0f - ty, wz = dz ? tz : 1.0f - tz;
                                acc[sl][0] += val * wx * wy * wz;
                                acc[sl][1] += (dx ? val : -val) * wy * wz;
                                acc[sl][2] += (dy ? val : -val) * wx * wz;
                                acc[sl][3] += (dz ? val : -val) * wx * wy;
                            }
                }
                __syncthreads();
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
            if (a.gverts) {
                // d psi / d vertex = (G/2) * d psi / d(ix) / scale ; loss = sum(rho) / 4
                const float k = mask * a.grad_scale * 0.25f * drho * (0.5f * G) / scale;
                float g[3] = {k * acc[sl][1], k * acc[sl][2], k * acc[sl][3]};
                gsum[0] += g[0]; gsum[1] += g[1]; gsum[2] += g[2];
                if (xform && o == 1) g[0] = -g[0];
                float* gp = a.gverts + ov * 3;
                gp[0] = g[0]; gp[1] = g[1]; gp[2] = g[2];
            }
        }
        if (a.gshift && o == 1) {
            block_sum4(gsum, 3, s.red);
            if (tid < 3) a.gshift[(size_t)b * 3 + tid] = gsum[tid];
        }
        __syncthreads();
    }
    float lp[1] = {loss_part};
    block_sum4(lp, 1, s.red);
    if (tid == 0) a.losses[b] = mask * lp[0] * 0.25f;
}

int launch_sdf(const ihmr_model* m, int B, const SdfArgs& a, cudaStream_t st) {
    if (B <= 0) return IHMR_OK;
    static bool configured = false;
    if (!configured) {
        IHMR_CUDA_OK(cudaFuncSetAttribute(k_sdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SdfSmem)));
        configured = true;
    }
    k_sdf<<<B, SDF_THREADS, sizeof(SdfSmem), st>>>(B, a, m->faces[0], m->faces[1]);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

}  // namespace ihmr
