// Kernel (a): query x reference distance stage (replaces apples/distance.py:718-745 jc69 and :681-715 scoredist
// as called from apples/Reference.py:138-152).
//
// Nucleotide: sequences are three bit-planes (lo, hi, valid) of 32-site words.  For a (query, reference) pair
//   valid    = popc(qv & rv)                                    distance.py:733-734
//   mismatch = popc(((qlo ^ rlo) | (qhi ^ rhi)) & qv & rv)      distance.py:737
// The dense kernel computes all pairs of a query block against all representatives with a GEMM-like tiling:
// operands are stored word-major ([plane][word][row]) so that one (plane, word) slice of a 64-row tile is 256
// contiguous bytes; a producer warp streams those slices into a 4-stage shared-memory ring with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx), 16 consumer warps hold a 4x4 register tile of pairs per thread (CTA tile 128
// queries x 64 representatives, one persistent CTA per SM) and do the LOP3/POPC work.  The binding resources are the
// integer pipes (XU POPC, ALU LOP3), not HBM ("rooflines").
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
#ifdef DT_PRODUCER_SLEEP
        __nanosleep(DT_PRODUCER_SLEEP);
#endif
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// three-input logic op with an explicit truth table (a = 0xf0, b = 0xcc, c = 0xaa)
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

// ---------------------------------------------------------------------------------------------------------------
// row-major [rows][3][W]  ->  word-major [3][Wp][rows_pad]   (padding must be pre-zeroed by the caller)
// ---------------------------------------------------------------------------------------------------------------
__global__ void transpose_nuc_kernel(const uint32_t* __restrict__ rm, int rows, int W, uint32_t* __restrict__ wm, int Wp,
                                     int rows_pad) {
    __shared__ uint32_t tile[32][33];
    const int plane = blockIdx.z;
    const int r0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, w = w0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && w < W) ? rm[((size_t)r * 3 + plane) * W + w] : 0u;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int w = w0 + i, r = r0 + threadIdx.x;
        if (w < Wp && r < rows_pad) wm[((size_t)plane * Wp + w) * rows_pad + r] = tile[threadIdx.x][i];
    }
}

void launch_transpose_nuc(const uint32_t* rm, int rows, int W, uint32_t* wm, int Wp, int rows_pad, cudaStream_t s) {
    dim3 grid((Wp + 31) / 32, (rows_pad + 31) / 32, 3), block(32, 8);
    transpose_nuc_kernel<<<grid, block, 0, s>>>(rm, rows, W, wm, Wp, rows_pad);
}

// ---------------------------------------------------------------------------------------------------------------
// dense nucleotide kernel
// ---------------------------------------------------------------------------------------------------------------
struct DenseNucArgs {
    const uint32_t* q_wm;  // [3][Wp][q_pad]
    const uint32_t* r_wm;  // [3][Wp][r_pad]
    int q_pad, r_pad, Wp;
    // keys epilogue
    uint32_t* keys;
    int64_t ldk;
    uint32_t k_one, k_two17;  // 1 and 1 << 17 (see the consumer loop)
    // full epilogue (parity export)
    int nq, n_ref, vmin;
    uint32_t* mism;
    uint32_t* valid;
    double* dist;
};

template <bool FULL>
__global__ void __launch_bounds__(DT_THREADS, DT_MINBLOCKS) dense_nuc_kernel(const DenseNucArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* stage_base = reinterpret_cast<uint32_t*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + DT_STAGES * DT_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + DT_STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_rt = a.r_pad / DT_TR;
    const int n_tiles = (a.q_pad / DT_TQ) * n_rt;
    const int n_chunks = a.Wp / DT_WC;

    if (tid == 0) {
        for (int s = 0; s < DT_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], DT_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one (plane, word) slice of a tile's rows = DT_TQ*4 / DT_TR*4 contiguous bytes per bulk copy
    auto issue_stage = [&](uint32_t git, int tile, int c) {
        const int s = git % DT_STAGES;
        const uint32_t ph = (git / DT_STAGES) & 1;
        const int qt = tile / n_rt, rt = tile % n_rt;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], DT_STAGE_BYTES);
        __syncwarp();
        uint32_t* sq = stage_base + (size_t)s * DT_STAGE_WORDS;
        uint32_t* sr = sq + 3 * DT_WC * DT_TQ;
        for (int i = lane; i < 3 * DT_WC; i += 32) {
            const int plane = i / DT_WC, w = i % DT_WC;
            const size_t grow = (size_t)plane * a.Wp + (size_t)c * DT_WC + w;
            tma_bulk_g2s(sq + i * DT_TQ, a.q_wm + grow * a.q_pad + (size_t)qt * DT_TQ, DT_TQ * 4, &full_bar[s]);
            tma_bulk_g2s(sr + i * DT_TR, a.r_wm + grow * a.r_pad + (size_t)rt * DT_TR, DT_TR * 4, &full_bar[s]);
        }
    };
#if DT_SELF_PRODUCE
    // EXPERIMENT (off by default, measured r01: 56 vs 78 Tcell-sites/s with the dedicated producer warp): no producer
    // warp, consumer warp 0 issues the copies of chunk it + STAGES - 1 before it computes chunk `it`.  It frees the
    // 17th warp's register granule (128 instead of 96 registers per thread) but couples warp 0 to the slowest warp.
    uint32_t p_it = 0;
    int p_tile = blockIdx.x, p_c = 0;
    auto produce_next = [&]() {
        if (p_tile < n_tiles) {
            issue_stage(p_it, p_tile, p_c);
            ++p_it;
            if (++p_c == n_chunks) { p_c = 0; p_tile += gridDim.x; }
        }
    };
    if (warp == 0)
        for (int k = 0; k < DT_STAGES - 1; ++k) produce_next();
#else
    if (warp == DT_CONSUMERS / 32) {
        // ===== producer warp =====
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int c = 0; c < n_chunks; ++c, ++it) {
                issue_stage(it, tile, c);
            }
        }
        return;
    }
#endif

    // ===== consumers: thread (tq, tr) owns queries 4*tq..+3 and representatives 4*tr..+3 of the tile =====
#ifdef DT_MAP_TR_FAST
    const int tr = tid % (DT_TR / 4), tq = tid / (DT_TR / 4);
#else
    const int tq = tid % (DT_TQ / 4), tr = tid / (DT_TQ / 4);
#endif
    uint32_t it = 0;
    const uint32_t k_one = a.k_one, k_two17 = a.k_two17;  // run-time multipliers: keeps the accumulations IMADs
#ifdef DT_CSA2
    const uint32_t k_four16 = a.k_two17 << 1;
#endif
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int qt = tile / n_rt, rt = tile % n_rt;
        // per pair: acc = mismatch count (low 16 bits) | valid count (high 16 bits); `ones` is the weight-1 plane of a
        // carry-save counter over the valid words: two valid words are folded with one full adder (2 LOP3) and only
        // the carry (weight 2) is popcounted, which takes a third of the POPC work off the XU pipe -- the kernel's
        // binding unit (DESIGN.md "rooflines")
        uint32_t acc[4][4], ones[4][4];
#ifdef DT_CSA2
        uint32_t twos[4][4];
#endif
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] = ones[i][j] = 0u;
#ifdef DT_CSA2
                twos[i][j] = 0u;
#endif
            }

        for (int c = 0; c < n_chunks; ++c, ++it) {
            const int s = it % DT_STAGES;
            const uint32_t ph = (it / DT_STAGES) & 1;
#if DT_SELF_PRODUCE
            if (warp == 0) produce_next();
#endif
            mbar_wait(&full_bar[s], ph);
            const uint32_t* sq = stage_base + (size_t)s * DT_STAGE_WORDS;
            const uint32_t* sr = sq + 3 * DT_WC * DT_TQ;
#ifdef DT_CSA2
            // EXPERIMENT (off by default, measured r01: spills at 96 registers, no gain): two-level carry-save counter
            // over the valid words: per four words three full adders (6 LOP3) and ONE popcount (of the weight-4
            // carry): 1.25 POPC + 5.5 LOP3 per word pair
#pragma unroll 1
            for (int w = 0; w < DT_WC; w += 4) {
                uint32_t ca[4][4];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t qlo[2][4], qhi[2][4], qva[2][4], rlo[2][4], rhi[2][4], rva[2][4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int ww = w + 2 * half + h;
                        const uint4 ql = *reinterpret_cast<const uint4*>(sq + (0 * DT_WC + ww) * DT_TQ + 4 * tq);
                        const uint4 qh = *reinterpret_cast<const uint4*>(sq + (1 * DT_WC + ww) * DT_TQ + 4 * tq);
                        const uint4 qv = *reinterpret_cast<const uint4*>(sq + (2 * DT_WC + ww) * DT_TQ + 4 * tq);
                        const uint4 rl = *reinterpret_cast<const uint4*>(sr + (0 * DT_WC + ww) * DT_TR + 4 * tr);
                        const uint4 rh = *reinterpret_cast<const uint4*>(sr + (1 * DT_WC + ww) * DT_TR + 4 * tr);
                        const uint4 rv = *reinterpret_cast<const uint4*>(sr + (2 * DT_WC + ww) * DT_TR + 4 * tr);
                        qlo[h][0] = ql.x; qlo[h][1] = ql.y; qlo[h][2] = ql.z; qlo[h][3] = ql.w;
                        qhi[h][0] = qh.x; qhi[h][1] = qh.y; qhi[h][2] = qh.z; qhi[h][3] = qh.w;
                        qva[h][0] = qv.x; qva[h][1] = qv.y; qva[h][2] = qv.z; qva[h][3] = qv.w;
                        rlo[h][0] = rl.x; rlo[h][1] = rl.y; rlo[h][2] = rl.z; rlo[h][3] = rl.w;
                        rhi[h][0] = rh.x; rhi[h][1] = rh.y; rhi[h][2] = rh.z; rhi[h][3] = rh.w;
                        rva[h][0] = rv.x; rva[h][1] = rv.y; rva[h][2] = rv.z; rva[h][3] = rv.w;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t v0 = lop3<0xc0>(qva[0][i], rva[0][j], 0u);
                            const uint32_t v1 = lop3<0xc0>(qva[1][i], rva[1][j], 0u);
                            const uint32_t x0 = lop3<0x3c>(qlo[0][i], rlo[0][j], 0u);
                            const uint32_t x1 = lop3<0x3c>(qlo[1][i], rlo[1][j], 0u);
                            const uint32_t t0 = lop3<0xbe>(qhi[0][i], rhi[0][j], x0);
                            const uint32_t t1 = lop3<0xbe>(qhi[1][i], rhi[1][j], x1);
                            const uint32_t m0 = lop3<0xc0>(t0, v0, 0u);
                            const uint32_t m1 = lop3<0xc0>(t1, v1, 0u);
                            const uint32_t o = ones[i][j];
                            const uint32_t c = lop3<0xe8>(o, v0, v1);
                            ones[i][j] = lop3<0x96>(o, v0, v1);
                            acc[i][j] = __popc(m0) * k_one + acc[i][j];
                            acc[i][j] = __popc(m1) * k_one + acc[i][j];
                            if (half == 0) {
                                ca[i][j] = c;
                            } else {
                                const uint32_t t2 = twos[i][j];
                                const uint32_t c4 = lop3<0xe8>(t2, ca[i][j], c);
                                twos[i][j] = lop3<0x96>(t2, ca[i][j], c);
                                acc[i][j] = __popc(c4) * k_four16 + acc[i][j];
                            }
                        }
                }
            }
#else
#pragma unroll 1
            for (int w = 0; w < DT_WC; w += 2) {
                uint32_t qlo[2][4], qhi[2][4], qva[2][4], rlo[2][4], rhi[2][4], rva[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4 ql = *reinterpret_cast<const uint4*>(sq + (0 * DT_WC + w + h) * DT_TQ + 4 * tq);
                    const uint4 qh = *reinterpret_cast<const uint4*>(sq + (1 * DT_WC + w + h) * DT_TQ + 4 * tq);
                    const uint4 qv = *reinterpret_cast<const uint4*>(sq + (2 * DT_WC + w + h) * DT_TQ + 4 * tq);
                    const uint4 rl = *reinterpret_cast<const uint4*>(sr + (0 * DT_WC + w + h) * DT_TR + 4 * tr);
                    const uint4 rh = *reinterpret_cast<const uint4*>(sr + (1 * DT_WC + w + h) * DT_TR + 4 * tr);
                    const uint4 rv = *reinterpret_cast<const uint4*>(sr + (2 * DT_WC + w + h) * DT_TR + 4 * tr);
                    qlo[h][0] = ql.x; qlo[h][1] = ql.y; qlo[h][2] = ql.z; qlo[h][3] = ql.w;
                    qhi[h][0] = qh.x; qhi[h][1] = qh.y; qhi[h][2] = qh.z; qhi[h][3] = qh.w;
                    qva[h][0] = qv.x; qva[h][1] = qv.y; qva[h][2] = qv.z; qva[h][3] = qv.w;
                    rlo[h][0] = rl.x; rlo[h][1] = rl.y; rlo[h][2] = rl.z; rlo[h][3] = rl.w;
                    rhi[h][0] = rh.x; rhi[h][1] = rh.y; rhi[h][2] = rh.z; rhi[h][3] = rh.w;
                    rva[h][0] = rv.x; rva[h][1] = rv.y; rva[h][2] = rv.z; rva[h][3] = rv.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // 10 LOP3 per pair per two words, pinned with explicit LUTs (left to itself ptxas re-fuses the
                        // AND into the adder and spends 12)
                        const uint32_t v0 = lop3<0xc0>(qva[0][i], rva[0][j], 0u);            // qv & rv
                        const uint32_t v1 = lop3<0xc0>(qva[1][i], rva[1][j], 0u);
                        const uint32_t x0 = lop3<0x3c>(qlo[0][i], rlo[0][j], 0u);            // qlo ^ rlo
                        const uint32_t x1 = lop3<0x3c>(qlo[1][i], rlo[1][j], 0u);
                        const uint32_t t0 = lop3<0xbe>(qhi[0][i], rhi[0][j], x0);            // (qhi ^ rhi) | x
                        const uint32_t t1 = lop3<0xbe>(qhi[1][i], rhi[1][j], x1);
                        const uint32_t m0 = lop3<0xc0>(t0, v0, 0u);                          // mismatching valid sites
                        const uint32_t m1 = lop3<0xc0>(t1, v1, 0u);
                        const uint32_t o = ones[i][j];
                        const uint32_t carry = lop3<0xe8>(o, v0, v1);                        // full adder: majority
                        ones[i][j] = lop3<0x96>(o, v0, v1);                                  //             parity
                        // the three accumulations go to the (idle) FMA pipe as IMADs: the multipliers are opaque
                        // registers so that ptxas cannot turn them back into ALU-pipe adds / shifts
                        acc[i][j] = __popc(m0) * k_one + acc[i][j];
                        acc[i][j] = __popc(m1) * k_one + acc[i][j];
                        acc[i][j] = __popc(carry) * k_two17 + acc[i][j];
                    }
            }
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        uint32_t accM[4][4], accV[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#ifdef DT_CSA2
                const uint32_t a2 = acc[i][j] + (__popc(ones[i][j]) << 16) + (__popc(twos[i][j]) << 17);
#else
                const uint32_t a2 = acc[i][j] + (__popc(ones[i][j]) << 16);
#endif
                accM[i][j] = a2 & 0xffffu;
                accV[i][j] = a2 >> 16;
            }

        // ---- epilogue ----
        const int q0 = qt * DT_TQ + 4 * tq, r0 = rt * DT_TR + 4 * tr;
        if (!FULL) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 o;
                o.x = accM[i][0] | (accV[i][0] << 16);
                o.y = accM[i][1] | (accV[i][1] << 16);
                o.z = accM[i][2] | (accV[i][2] << 16);
                o.w = accM[i][3] | (accV[i][3] << 16);
                *reinterpret_cast<uint4*>(a.keys + (size_t)(q0 + i) * a.ldk + r0) = o;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = q0 + i, r = r0 + j;
                    if (q < a.nq && r < a.n_ref) {
                        const size_t o = (size_t)q * a.n_ref + r;
                        a.mism[o] = accM[i][j];
                        a.valid[o] = accV[i][j];
                        a.dist[o] = jc69_from_counts(accM[i][j], accV[i][j], a.vmin);
                    }
                }
        }
    }
}

cudaError_t dense_nuc_configure() {
    cudaError_t e = cudaFuncSetAttribute(dense_nuc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(dense_nuc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
}

static int dense_grid(int q_pad, int r_pad, int num_sms) {
    int tiles = (q_pad / DT_TQ) * (r_pad / DT_TR);
    int g = DT_MINBLOCKS * num_sms;  // resident CTAs per SM, persistent over tiles
    return tiles < g ? tiles : g;
}

void launch_dense_nuc_keys(const uint32_t* q_wm, int q_pad, const uint32_t* r_wm, int r_pad, int Wp, uint32_t* keys,
                           int64_t ldk, int num_sms, cudaStream_t s) {
    DenseNucArgs a{};
    a.q_wm = q_wm; a.r_wm = r_wm; a.q_pad = q_pad; a.r_pad = r_pad; a.Wp = Wp; a.keys = keys; a.ldk = ldk;
    a.k_one = 1u; a.k_two17 = 1u << 17;
    dense_nuc_kernel<false><<<dense_grid(q_pad, r_pad, num_sms), DT_THREADS, DT_SMEM_BYTES, s>>>(a);
}

void launch_dense_nuc_full(const uint32_t* q_wm, int q_pad, int nq, const uint32_t* r_wm, int r_pad, int n_ref, int Wp,
                           int vmin, uint32_t* mism, uint32_t* valid, double* dist, int num_sms, cudaStream_t s) {
    DenseNucArgs a{};
    a.q_wm = q_wm; a.r_wm = r_wm; a.q_pad = q_pad; a.r_pad = r_pad; a.Wp = Wp;
    a.k_one = 1u; a.k_two17 = 1u << 17;
    a.nq = nq; a.n_ref = n_ref; a.vmin = vmin; a.mism = mism; a.valid = valid; a.dist = dist;
    dense_nuc_kernel<true><<<dense_grid(q_pad, r_pad, num_sms), DT_THREADS, DT_SMEM_BYTES, s>>>(a);
}

// ---------------------------------------------------------------------------------------------------------------
// dense amino-acid kernel: dist[q][r] = scoredist(q, r)  (distance.py:681-715)
// one thread per reference row, AA_TQ queries per block held in shared memory, BLOSUM45 (21x21, gap row/col = 0)
// in shared memory.  fp64 table sum per pair; the summation order differs from numpy's BLAS ddot (distance.py:706),
// parity target 1e-9 relative.
// ---------------------------------------------------------------------------------------------------------------
__constant__ double c_blosum45[441] = {
#include "blosum45.inc"
};

constexpr int AA_TQ = 4;
constexpr int AA_THREADS = 128;
constexpr int AA_CHUNK = 1024;  // sites of the query tile staged per pass

__global__ void __launch_bounds__(AA_THREADS) dense_aa_kernel(const uint8_t* __restrict__ q, int nq,
                                                               const uint8_t* __restrict__ r, int n_r, int Lp, int L,
                                                               double overlap, double* __restrict__ dist, int64_t ldd,
                                                               uint32_t* __restrict__ valid_out) {
    __shared__ double tab[441];
    __shared__ __align__(16) uint8_t qs[AA_TQ][AA_CHUNK];
    for (int i = threadIdx.x; i < 441; i += AA_THREADS) tab[i] = c_blosum45[i];
    const int q0 = blockIdx.y * AA_TQ;
    const int row = blockIdx.x * AA_THREADS + threadIdx.x;
    double sum[AA_TQ];
    uint32_t val[AA_TQ];
#pragma unroll
    for (int k = 0; k < AA_TQ; ++k) { sum[k] = 0.0; val[k] = 0u; }
    for (int c0 = 0; c0 < Lp; c0 += AA_CHUNK) {
        const int clen = min(AA_CHUNK, Lp - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < AA_TQ * (clen / 16); i += AA_THREADS) {
            const int k = i / (clen / 16), x = i % (clen / 16);
            uint4 v = make_uint4(0x14141414u, 0x14141414u, 0x14141414u, 0x14141414u);  // gaps
            if (q0 + k < nq) v = *reinterpret_cast<const uint4*>(q + (size_t)(q0 + k) * Lp + c0 + 16 * x);
            *reinterpret_cast<uint4*>(&qs[k][16 * x]) = v;
        }
        __syncthreads();
        if (row < n_r) {
            const uint8_t* rr = r + (size_t)row * Lp + c0;
            for (int x = 0; x < clen; x += 16) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rr + x);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const uint32_t rc = (rw[b >> 2] >> (8 * (b & 3))) & 0xffu;
#pragma unroll
                    for (int k = 0; k < AA_TQ; ++k) {
                        const uint32_t qc = qs[k][x + b];
                        sum[k] += tab[qc * 21 + rc];
                        val[k] += (qc < 20u && rc < 20u) ? 1u : 0u;
                    }
                }
            }
        }
    }
    if (row < n_r) {
#pragma unroll
        for (int k = 0; k < AA_TQ; ++k)
            if (q0 + k < nq) {
                dist[(size_t)(q0 + k) * ldd + row] = scoredist_from_sum(sum[k], val[k], L, overlap);
                if (valid_out) valid_out[(size_t)(q0 + k) * ldd + row] = val[k];
            }
    }
}

void launch_dense_aa(const uint8_t* q, int nq, const uint8_t* r, int n_r, int Lp, int L, double overlap, double* dist,
                     int64_t ldd, uint32_t* valid_out, cudaStream_t s) {
    dim3 grid((n_r + AA_THREADS - 1) / AA_THREADS, (nq + AA_TQ - 1) / AA_TQ);
    dense_aa_kernel<<<grid, AA_THREADS, 0, s>>>(q, nq, r, n_r, Lp, L, overlap, dist, ldd, valid_out);
}
