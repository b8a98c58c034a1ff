// Blend-shape contraction on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
//   forward : off (N x 2336) = X (N x 160) . D (160 x 2336)        [posedirs ; shapedirs]
//   backward: dX  (N x 160)  = gposed (N x 2336) . D^T
//
// Both are C[M,Nc] = A[M,K] . B[Nc,K]^T with A and B "K-major" (row-major with K contiguous):
// forward B = D^T (2336 x 160), backward B = D (160 x 2336).  This is the dense
// "(B*2) x 135 by 135 x 2334" contraction BASELINE.json's north_star assigns to the tensor
// cores (smplx lbs: `torch.matmul(pose_feature, posedirs)`, reached from
// /root/reference/src/models/optimize_model.py:194), with the shape blend folded in.
//
// fp32 accuracy on a TF32 pipe: every operand is split in registers on its way to shared
// memory, x = hi + lo with hi = x rounded down to 10 mantissa bits (exact in TF32) and
// lo = x - hi (exact in fp32), and three MMAs are issued per K step: hi*hi + lo*hi + hi*lo.
// The dropped lo*lo term is ~2^-22 relative; accumulation is fp32 in TMEM.
//
// One CTA (8 warps) computes a 128 x BN tile: K is consumed in chunks of 32 (all threads load
// and split the chunk into the canonical no-swizzle K-major core-matrix layout, one elected
// thread issues 4 K-steps x 3 MMAs and commits to an mbarrier), then each warp drains its 32
// TMEM lanes with tcgen05.ld and writes its rows.  Two CTAs per SM overlap one CTA's loads with
// the other's MMAs.
#include <string.h>

#include "kernels.cuh"

namespace ihmr {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // K elements per chunk = 8 core-matrix columns of 16 bytes
constexpr int TC_THREADS = 256;            // 8 warps: all load; warp w drains TMEM lane quarter w % 4, column half w / 4
constexpr uint32_t TC_LBO = 128;          // bytes between core matrices adjacent in K
constexpr uint32_t TC_SBO = 1024;         // bytes between 8-row groups: 8 K-columns x 128 bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): 8 x 16-byte core matrices
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);            // start address, 16-byte units
    d |= (uint64_t)((TC_LBO >> 4) & 0x3fffu) << 16;     // leading byte offset
    d |= (uint64_t)((TC_SBO >> 4) & 0x3fffu) << 32;     // stride byte offset
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    return d;                                           // base offset 0, layout type 0 = no swizzle
}

// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// float4 of row-major src -> (hi, lo) 16-byte core-matrix rows
__device__ __forceinline__ void split_store(float4 v, float4* hi_dst, float4* lo_dst) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    *hi_dst = h;
    *lo_dst = l;
}

// rows x 32 chunk of a K-major matrix -> canonical layout [row/8][kcol 0..7][row%8][16 B], split hi/lo.
// Lane mapping: 8 consecutive rows x 4 consecutive K-columns per warp access = 512 contiguous bytes of
// shared memory.  The chunk is fetched into registers first (all loads in flight together, and in flight
// while the previous chunk's MMAs run) and split + stored afterwards.
template <int NITEMS>
struct ChunkRegs {
    float4 v[NITEMS];
};

__device__ __forceinline__ void item_coords(int it, int& r, int& kc, uint32_t& off) {
    const int r8 = it & 7, kc_lo = (it >> 3) & 3, blk = it >> 5;      // blk enumerates (row group, kcol high bit)
    kc = kc_lo + 4 * (blk & 1);
    const int rg = blk >> 1;
    r = rg * 8 + r8;
    off = rg * TC_SBO + kc * TC_LBO + r8 * 16;
}

template <int NITEMS>
__device__ __forceinline__ void fetch_chunk(ChunkRegs<NITEMS>& regs, const float* __restrict__ src, int ld, int row0,
                                            int rows_valid, int rows_tile, int k0, int tid) {
#pragma unroll
    for (int i = 0; i < NITEMS; ++i) {
        const int it = tid + i * TC_THREADS;
        int r, kc;
        uint32_t off;
        item_coords(it, r, kc, off);
        regs.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < rows_tile * 8 && r < rows_valid)
            regs.v[i] = *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld + k0 + kc * 4);
    }
}

// the same through a row list: rowsrc[i] is the source row of this thread's i-th item (fixed over the chunks), -1 = none
template <int NITEMS>
__device__ __forceinline__ void fetch_chunk_rows(ChunkRegs<NITEMS>& regs, const float* __restrict__ src, int ld,
                                                 const int (&rowsrc)[NITEMS], int k0, int tid) {
#pragma unroll
    for (int i = 0; i < NITEMS; ++i) {
        const int it = tid + i * TC_THREADS;
        int r, kc;
        uint32_t off;
        item_coords(it, r, kc, off);
        regs.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rowsrc[i] >= 0) regs.v[i] = *reinterpret_cast<const float4*>(src + (size_t)rowsrc[i] * ld + k0 + kc * 4);
    }
}

template <int NITEMS>
__device__ __forceinline__ void store_chunk(const ChunkRegs<NITEMS>& regs, int rows_tile, unsigned char* hi, unsigned char* lo, int tid) {
#pragma unroll
    for (int i = 0; i < NITEMS; ++i) {
        const int it = tid + i * TC_THREADS;
        int r, kc;
        uint32_t off;
        item_coords(it, r, kc, off);
        if (it < rows_tile * 8) split_store(regs.v[i], reinterpret_cast<float4*>(hi + off), reinterpret_cast<float4*>(lo + off));
    }
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS)
k_gemm_tf32x3(int M, int Nc, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
              float* __restrict__ C, int ldc, const int* __restrict__ rows, const int* __restrict__ nrows,
              const float* __restrict__ Bq, size_t part_stride) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + TC_BM * TC_BK * 4;
    unsigned char* b_hi = a_lo + TC_BM * TC_BK * 4;
    unsigned char* b_lo = b_hi + BN * TC_BK * 4;
    __shared__ __align__(8) uint64_t mbar, mbar_b;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    if (nrows) M = min(M, *nrows);                        // row list: only its first *nrows entries exist
    if (m0 >= M) return;
    const int rows_a = min(TC_BM, M - m0);
    const int rows_b = min(BN, Nc - n0);
    const int n_inst = (rows_b + 15) & ~15;               // MMA N: multiple of 16 covering the valid B rows
    constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar_b)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = umma_idesc(n_inst);

    uint32_t parity = 0;
    const int nchunks = K / TC_BK;
    // split K (gridDim.z > 1): this CTA takes a contiguous range of chunks and writes a partial product into its own
    // copy of C (part_stride floats apart); launch_gemm_tf32x3 sums the copies in fixed order afterwards
    const int ch_begin = (int)(((long long)nchunks * blockIdx.z) / gridDim.z), ch_end = (int)(((long long)nchunks * (blockIdx.z + 1)) / gridDim.z);
    C += (size_t)blockIdx.z * part_stride;
    // pre-split B: the (hi, lo) blocks of this N tile's chunk arrive by two bulk copies (first n_inst rows of each)
    auto issue_b = [&](int ch) {
        const float* src = Bq + ((size_t)blockIdx.x * nchunks + ch) * (2 * BN * TC_BK);
        const uint32_t bytes = (uint32_t)n_inst * (TC_BK * 4), bar = smem_u32(&mbar_b);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(b_hi)), "l"(src), "r"(bytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(b_lo)), "l"(src + BN * TC_BK), "r"(bytes), "r"(bar) : "memory");
    };
    constexpr int A_ITEMS = TC_BM * 8 / TC_THREADS;
    ChunkRegs<A_ITEMS> ra;
    ChunkRegs<BN * 8 / TC_THREADS> rb;
    int rowsrc[A_ITEMS];                                  // source row of this thread's A items (row list or identity)
#pragma unroll
    for (int i = 0; i < A_ITEMS; ++i) {
        int r, kc;
        uint32_t off;
        item_coords(tid + i * TC_THREADS, r, kc, off);
        rowsrc[i] = (r < rows_a) ? (rows ? rows[m0 + r] : m0 + r) : -1;
    }
    fetch_chunk_rows(ra, A, lda, rowsrc, ch_begin * TC_BK, tid);
    if (Bq) { if (tid == 0) issue_b(ch_begin); }
    else fetch_chunk(rb, B, ldb, n0, rows_b, n_inst, ch_begin * TC_BK, tid);
    for (int ch = ch_begin; ch < ch_end; ++ch) {
        store_chunk(ra, TC_BM, a_hi, a_lo, tid);
        if (!Bq) store_chunk(rb, n_inst, b_hi, b_lo, tid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy (tensor core)
        __syncthreads();
        if (tid == 0) {
            if (Bq) mbar_wait(smem_u32(&mbar_b), (uint32_t)((ch - ch_begin) & 1));    // this chunk's B blocks have landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
                const uint32_t koff = ks * 2 * TC_LBO;               // 8 tf32 = two 16-byte core-matrix columns
                const uint64_t ah = umma_desc(smem_u32(a_hi) + koff), al = umma_desc(smem_u32(a_lo) + koff);
                const uint64_t bh = umma_desc(smem_u32(b_hi) + koff), bl = umma_desc(smem_u32(b_lo) + koff);
                umma_tf32(tmem_d, ah, bh, idesc, ch > ch_begin || ks > 0);
                umma_tf32(tmem_d, al, bh, idesc, true);
                umma_tf32(tmem_d, ah, bl, idesc, true);
            }
            // arrives on the mbarrier when all MMAs issued so far have completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
        }
        if (ch + 1 < ch_end) {                                       // next chunk's global loads fly while the MMAs run
            fetch_chunk_rows(ra, A, lda, rowsrc, (ch + 1) * TC_BK, tid);
            if (!Bq) fetch_chunk(rb, B, ldb, n0, rows_b, n_inst, (ch + 1) * TC_BK, tid);
        }
        mbar_wait(smem_u32(&mbar), parity);                          // operands may be overwritten, accumulator is current
        parity ^= 1;
        if (Bq && tid == 0 && ch + 1 < ch_end) issue_b(ch + 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows m0 + 32w + lane.  tcgen05.ld hands every lane
    // 32 consecutive columns of ITS row; the 32 x 32 block goes through shared memory (the operand buffers
    // are free now) so that each store instruction of the warp writes four complete 128-byte row segments.
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    const int quarter = warp & 3;                          // a warp can only read TMEM lanes 32 (w % 4) .. +31
    for (int c0 = (warp >> 2) * 32; c0 < n_inst; c0 += 64) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stage + lane * 36 + q * 4) =
                make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
        __syncwarp();
        const int cq = (lane & 7) * 4;                       // this lane's 4 columns inside the 32-column block
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = (lane >> 3) + 4 * i;               // row inside the quarter's 32 rows
            const int row = m0 + quarter * 32 + r;
            if (row < M && n0 + c0 + cq < Nc) {
                const int drow = rows ? rows[row] : row;
                *reinterpret_cast<float4*>(C + (size_t)drow * ldc + n0 + c0 + cq) = *reinterpret_cast<const float4*>(stage + r * 36 + cq);
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------ skinning blend on the tensor cores
// T[v, (h,e)] = sum_j W[v,j] A[h][j][e] is a (778 x 16) . (16 x 12 H) contraction: 7 vertex tiles of 128
// rows (the constant operand, split hi/lo once per CTA and kept in shared memory for its lifetime) against the
// joint transforms of 16 hands (N = 192 columns, restaged per group), 2 K-steps x 3 MMAs per tile.  The
// accumulators of two tiles live in TMEM (2 x 192 of 512 columns) so that the MMAs of tile t+2 overlap the
// epilogue of tile t+1; the epilogue reads a vertex's twelve T entries per hand with tcgen05.ld, applies them to
// the posed vertex and writes the skinned vertex.  Persistent CTAs (one per SM) stride over the hand groups.
constexpr int SKT_THREADS = 512;          // 16 warps: lane quarter = warp % 4, hands 4 * (warp / 4) .. + 3
constexpr int SKT_HANDS = 16;             // hands per group = MMA N / 12
constexpr int SKT_N = SKT_HANDS * 12;     // 192
constexpr int SKT_TILES = (NV + TC_BM - 1) / TC_BM;       // 7
constexpr uint32_t SKT_SBO = 4 * TC_LBO;  // K = 16: four 16-byte K-columns per 8-row group
constexpr uint32_t SKT_W_TILE = TC_BM * NJ * 4;           // bytes of one 128 x 16 operand tile (8 KB)
constexpr uint32_t SKT_B_BYTES = SKT_N * NJ * 4;          // 12 KB

__device__ __forceinline__ uint64_t umma_desc_k16(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((TC_LBO >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((SKT_SBO >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// MODE 0: verts = T_v [v_posed; 1] (forward).  MODE 1: gposed = T_v^T g_v, the gradient of the posed vertex
// (vertex half of the skinning backward); `off` then holds the vertex gradients, `vtemp` the fingertip part.
template <int MODE>
__global__ void __launch_bounds__(SKT_THREADS, 1)
k_skin_fwd_tc(int n, const float* __restrict__ off, const float* __restrict__ A, const float* __restrict__ vtemp,
              const float* __restrict__ W4, float* __restrict__ verts, float* __restrict__ bbox) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ int s_box[SKT_HANDS][6];                           // MODE 0: boxes of the group's hands (for the penetration op)
    unsigned char* w_hi = smem;                                   // [7][128 x 16]
    unsigned char* w_lo = w_hi + SKT_TILES * SKT_W_TILE;
    unsigned char* b_hi = w_lo + SKT_TILES * SKT_W_TILE;          // [192 x 16]
    unsigned char* b_lo = b_hi + SKT_B_BYTES;
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the constant operand: W4[kc][v] is exactly the 16-byte K-column kc of row v
    for (int it = tid; it < SKT_TILES * TC_BM * 4; it += SKT_THREADS) {
        const int kc = it & 3, v = it >> 2, t = v >> 7, r = v & 127;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < NV) w = reinterpret_cast<const float4*>(W4)[kc * NV + v];
        const uint32_t o = t * SKT_W_TILE + (r >> 3) * SKT_SBO + kc * TC_LBO + (r & 7) * 16;
        split_store(w, reinterpret_cast<float4*>(w_hi + o), reinterpret_cast<float4*>(w_lo + o));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = umma_idesc(SKT_N);
    uint32_t phase = 0u;                                      // bit b: parity of the next completion of mbar[b]
    const int quarter = warp & 3, hsub = warp >> 2;

    auto issue_tile = [&](int t) {                            // one thread: 2 K-steps x 3 MMAs into buffer t & 1
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_d + (uint32_t)((t & 1) * 256);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t koff = ks * 2 * TC_LBO;
            const uint64_t ah = umma_desc_k16(smem_u32(w_hi) + t * SKT_W_TILE + koff), al = umma_desc_k16(smem_u32(w_lo) + t * SKT_W_TILE + koff);
            const uint64_t bh = umma_desc_k16(smem_u32(b_hi) + koff), bl = umma_desc_k16(smem_u32(b_lo) + koff);
            umma_tf32(d, ah, bh, idesc, ks > 0);
            umma_tf32(d, al, bh, idesc, true);
            umma_tf32(d, ah, bl, idesc, true);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar[t & 1])) : "memory");
    };

    const int ngroups = (n + SKT_HANDS - 1) / SKT_HANDS;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int h0 = grp * SKT_HANDS, nh = min(SKT_HANDS, n - h0);
        // B operand: row (hh, e), K = joint: A[h][j][e] gathered along j
        for (int it = tid; it < SKT_N * 4; it += SKT_THREADS) {
            const int kc = it & 3, row = it >> 2, hh = row / 12, e = row - hh * 12;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hh < nh) {
                const float* a = A + ((size_t)(h0 + hh) * NJ + kc * 4) * 12 + e;
                v = make_float4(a[0], a[12], a[24], a[36]);
            }
            const uint32_t o = (row >> 3) * SKT_SBO + kc * TC_LBO + (row & 7) * 16;
            split_store(v, reinterpret_cast<float4*>(b_hi + o), reinterpret_cast<float4*>(b_lo + o));
        }
        if (MODE == 0 && bbox && tid < SKT_HANDS * 6) box_slot_init(s_box[tid / 6], tid % 6);
        BoxAcc box[4];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) { issue_tile(0); issue_tile(1); }
        for (int t = 0; t < SKT_TILES; ++t) {
            const int buf = t & 1;
            const int v = t * TC_BM + quarter * 32 + lane;
            const bool vok = v < NV;
            // this thread's posed vertices of its four hands: in flight before the accumulator is awaited
            float vp[4][3];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int hh = hsub * 4 + q;
                if (MODE == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        vp[q][c] = (vok && hh < nh) ? vtemp[v * 3 + c] + off[(size_t)(h0 + hh) * LDN + v * 3 + c] : 0.f;
                } else {
                    const int tip = (v == 744) ? 0 : (v == 320) ? 1 : (v == 443) ? 2 : (v == 554) ? 3 : (v == 671) ? 4 : -1;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float g = (vok && hh < nh && off) ? off[((size_t)(h0 + hh) * NV + v) * 3 + c] : 0.f;
                        if (tip >= 0 && hh < nh && vtemp) g += vtemp[((size_t)(h0 + hh) * 5 + tip) * 3 + c];
                        vp[q][c] = g;
                    }
                }
            }
            mbar_wait(smem_u32(&mbar[buf]), (phase >> buf) & 1u);
            phase ^= 1u << buf;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int hh = hsub * 4 + q;
                uint32_t r[12];
                const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + hh * 12);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr + 4) : "memory");
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]) : "r"(taddr + 8) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (vok && hh < nh) {
                    if (MODE == 0) {
                        float* o = verts + ((size_t)(h0 + hh) * NV + v) * 3;
                        float x[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            x[c] = __uint_as_float(r[c * 4 + 0]) * vp[q][0] + __uint_as_float(r[c * 4 + 1]) * vp[q][1] +
                                   __uint_as_float(r[c * 4 + 2]) * vp[q][2] + __uint_as_float(r[c * 4 + 3]);
                            o[c] = x[c];
                        }
                        box[q].add(x[0], x[1], x[2]);
                    } else {
                        float* o = verts + (size_t)(h0 + hh) * LDN + v * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            o[c] = __uint_as_float(r[0 * 4 + c]) * vp[q][0] + __uint_as_float(r[1 * 4 + c]) * vp[q][1] +
                                   __uint_as_float(r[2 * 4 + c]) * vp[q][2];
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();                                  // every warp has drained accumulator buffer `buf`
            if (tid == 0 && t + 2 < SKT_TILES) issue_tile(t + 2);
        }
        // all seven commits were awaited: the B operand may be restaged
        if (MODE == 0 && bbox) {
#pragma unroll
            for (int q = 0; q < 4; ++q) box[q].commit(s_box[hsub * 4 + q]);
            __syncthreads();
            if (tid < SKT_HANDS * 6 && tid / 6 < nh) bbox[(size_t)(h0 + tid / 6) * 6 + tid % 6] = ordered_float(s_box[tid / 6][tid % 6]);
            // (the slots are re-initialised before the next group's first barrier, after this read: the restaging loop
            //  above runs first and thread tid re-initialises the very slot it has just read)
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(512) : "memory");
}

int launch_skin_fwd_tc(const ihmr_model* m, int n, const float* off, const float* A, float* verts, cudaStream_t st, float* bbox) {
    if (n <= 0) return IHMR_OK;
    const size_t smem = 2 * (size_t)SKT_TILES * SKT_W_TILE + 2 * SKT_B_BYTES;
    static unsigned long long configured = 0ull;
    if (int rc = ensure_dynamic_smem(k_skin_fwd_tc<0>, smem, configured)) return rc;
    const int ngroups = (n + SKT_HANDS - 1) / SKT_HANDS;
    k_skin_fwd_tc<0><<<min(ngroups, m->num_sms), SKT_THREADS, smem, st>>>(n, off, A, m->vtemp, m->W4, verts, bbox);
    IHMR_LAUNCH_OK();
    return IHMR_OK;
}

// C[row] = sum over the ksplit partial products, fixed order (row list as in the contraction)
__global__ void k_splitk_sum(int M, int Nc, int ksplit, const float* __restrict__ parts, size_t part_stride, float* __restrict__ C,
                             int ldc, const int* __restrict__ rows, const int* __restrict__ nrows) {
    if (nrows) M = min(M, *nrows);
    const int per_row = Nc / 4;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)M * per_row) return;
    const int r = (int)(i / per_row), c4 = (int)(i % per_row);
    const size_t o = (size_t)(rows ? rows[r] : r) * ldc + c4 * 4;
    float4 acc = *reinterpret_cast<const float4*>(parts + o);
    for (int z = 1; z < ksplit; ++z) {
        const float4 v = *reinterpret_cast<const float4*>(parts + (size_t)z * part_stride + o);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(C + o) = acc;
}

template <int BN>
static int launch_tc(int M, int Nc, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, cudaStream_t st,
                     const int* rows, const int* nrows, const float* Bq, int ksplit, float* parts, size_t part_stride) {
    const size_t smem = (size_t)(2 * TC_BM + 2 * BN) * TC_BK * 4;
    static unsigned long long configured = 0ull;
    if (int rc = ensure_dynamic_smem(k_gemm_tf32x3<BN>, smem, configured)) return rc;
    dim3 grid((Nc + BN - 1) / BN, (M + TC_BM - 1) / TC_BM, ksplit);
    k_gemm_tf32x3<BN><<<grid, TC_THREADS, smem, st>>>(M, Nc, K, A, lda, B, ldb, ksplit > 1 ? parts : C, ldc, rows, nrows, Bq,
                                                      ksplit > 1 ? part_stride : 0);
    IHMR_LAUNCH_OK();
    if (ksplit > 1) {
        const size_t total = (size_t)M * (Nc / 4);
        k_splitk_sum<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(M, Nc, ksplit, parts, part_stride, C, ldc, rows, nrows);
        IHMR_LAUNCH_OK();
    }
    return IHMR_OK;
}

// C[M,Nc] = A[M,K] . B[Nc,K]^T ; K % 32 == 0, Nc % 4 == 0, lda/ldb/ldc % 4 == 0
int launch_gemm_tf32x3(int M, int Nc, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       cudaStream_t st, const int* rows, const int* nrows, const float* Bq, int ksplit, float* parts) {
    if (M <= 0) return IHMR_OK;
    if (K % TC_BK || Nc % 4 || lda % 4 || ldb % 4 || ldc % 4) { set_error("gemm_tf32x3: unsupported shape"); return IHMR_E_INVALID; }
    if (ksplit < 1 || ksplit > K / TC_BK || (ksplit > 1 && !parts)) { set_error("gemm_tf32x3: bad K split"); return IHMR_E_INVALID; }
    const size_t part_stride = (size_t)M * ldc;          // one full copy of C per K slice (rows index the same space)
    if (Nc > 160) return launch_tc<256>(M, Nc, K, A, lda, B, ldb, C, ldc, st, rows, nrows, Bq, ksplit, parts, part_stride);
    return launch_tc<160>(M, Nc, K, A, lda, B, ldb, C, ldc, st, rows, nrows, Bq, ksplit, parts, part_stride);
}

// Host side of the pre-split operand: for every (N tile of BN rows, K chunk of 32) a block [hi | lo], each BN x 32 floats in
// the canonical no-swizzle K-major core-matrix layout the kernel's shared-memory descriptors expect
// ([row / 8][K column of 16 bytes][row % 8][4 floats]); rows beyond Nc are zero.  BN follows launch_gemm_tf32x3.
static int presplit_bn(int Nc) { return Nc > 160 ? 256 : 160; }
size_t gemm_presplit_floats(int Nc, int K) {
    const int BN = presplit_bn(Nc);
    return (size_t)((Nc + BN - 1) / BN) * (K / TC_BK) * 2 * BN * TC_BK;
}
void gemm_presplit_b(const float* B, int Nc, int K, int ldb, float* out) {
    const int BN = presplit_bn(Nc), ntiles = (Nc + BN - 1) / BN, nch = K / TC_BK;
    for (int t = 0; t < ntiles; ++t)
        for (int ch = 0; ch < nch; ++ch) {
            float* hi = out + ((size_t)t * nch + ch) * (2 * BN * TC_BK);
            float* lo = hi + BN * TC_BK;
            for (int r = 0; r < BN; ++r)
                for (int kc = 0; kc < 8; ++kc)
                    for (int j = 0; j < 4; ++j) {
                        const int row = t * BN + r;
                        const float v = row < Nc ? B[(size_t)row * ldb + ch * TC_BK + kc * 4 + j] : 0.f;
                        uint32_t bits;
                        memcpy(&bits, &v, 4);
                        bits &= 0xffffe000u;
                        float h;
                        memcpy(&h, &bits, 4);
                        const size_t o = (size_t)(r / 8) * (TC_SBO / 4) + kc * (TC_LBO / 4) + (r % 8) * 4 + j;
                        hi[o] = h;
                        lo[o] = v - h;
                    }
        }
}

}  // namespace ihmr
