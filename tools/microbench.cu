// Integer-pipe microbenchmark behind the distance-kernel roofline: lane-ops per clock per SM of
// LOP3 (ALU pipe), POPC (XU pipe), IADD3 and IMAD (FMA pipe) alone, and of the two instruction mixes that bracket the
// dense kernel (plain counting: 3 LOP3 + 2 POPC + 2 IMAD per 32-site word pair; full carry-save: 5 LOP3 + 1 POPC +
// 1 IMAD).
//
// Every counted operation is ONE inline-PTX instruction in an `asm volatile` statement on one of ILP independent
// accumulator chains, so ptxas can neither merge two logic operations into one LOP3 nor move an addition to another pipe
// nor drop anything.  tools/microbench_sass.sh dumps the SASS of every kernel and counts the mnemonics inside the timed
// loop: profiles/microbench_sass_r02.txt shows one LOP3 / POPC / IADD3 / IMAD per counted operation.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 4;       // independent chains per thread (keeps the kernel under 32 registers: 2048 threads/SM resident)
constexpr int UNROLL = 4;    // chain steps per loop iteration
constexpr int ITERS = 2048;

#define OP_LOP3(x, y, z) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z))
#define OP_POPC(x) asm volatile("popc.b32 %0, %0;" : "+r"(x))
#define OP_IADD(x, y) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_IMAD(x, y, z) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x) : "r"(y), "r"(z))

enum { M_LOP3, M_POPC, M_IADD, M_IMAD, M_MIX_PLAIN, M_MIX_CSA, M_COUNT };

template <int MODE>
__global__ void __launch_bounds__(1024, 2) k(unsigned* out, unsigned seed, long long* cycles, unsigned long long* clk) {
    unsigned x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = seed + threadIdx.x * 7 + i * 13;
        y[i] = seed ^ (i * 0x9e3779b9u);
        z[i] = seed * (i + 3) + blockIdx.x;
    }
    __syncthreads();
    unsigned long long g0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == M_LOP3) {
                    OP_LOP3(x[i], y[i], z[i]);
                } else if (MODE == M_POPC) {
                    OP_POPC(x[i]);
                } else if (MODE == M_IADD) {
                    OP_IADD(x[i], y[i]);
                } else if (MODE == M_IMAD) {
                    OP_IMAD(x[i], y[i], z[i]);
                } else if (MODE == M_MIX_PLAIN) {   // one word pair: 3 LOP3 + 2 POPC + 2 IMAD
                    unsigned a = x[i], b = y[i], c = z[i], p, q;
                    OP_LOP3(a, b, c);
                    OP_LOP3(b, a, c);
                    OP_LOP3(c, a, b);
                    p = a; OP_POPC(p);
                    q = c; OP_POPC(q);
                    OP_IMAD(x[i], p, seed);
                    OP_IMAD(y[i], q, seed);
                } else {                            // one word pair with full carry-save: 5 LOP3 + 1 POPC + 1 IMAD
                    unsigned a = x[i], b = y[i], c = z[i], p;
                    OP_LOP3(a, b, c);
                    OP_LOP3(b, a, c);
                    OP_LOP3(c, a, b);
                    OP_LOP3(z[i], a, c);
                    OP_LOP3(y[i], b, c);
                    p = a; OP_POPC(p);
                    OP_IMAD(x[i], p, seed);
                }
            }
        }
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // SM clock during the run: clock64 ticks per globaltimer nanosecond
        unsigned long long g1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
        clk[0] = (unsigned long long)(t1 - t0);
        clk[1] = g1 - g0;
    }
}

template <int MODE>
void run(const char* name, int threads) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks_per_sm = 2048 / threads;
    const int nb = p.multiProcessorCount * blocks_per_sm;   // one full wave: every block is resident at once
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, threads, 0);
    unsigned* out;
    long long* cyc;
    unsigned long long* clk;
    cudaMalloc(&out, (size_t)nb * threads * 4);
    cudaMalloc(&cyc, nb * 8);
    cudaMalloc(&clk, 16);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<nb, threads>>>(out, 12345u, cyc, clk);
    cudaEventRecord(a);
    k<MODE><<<nb, threads>>>(out, 12345u, cyc, clk);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    unsigned long long hc[2];
    cudaMemcpy(hc, clk, 16, cudaMemcpyDeviceToHost);
    const double mhz = 1e3 * (double)hc[0] / (double)hc[1];
    long long* h = new long long[nb];
    cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < nb; ++i) mx = h[i] > mx ? h[i] : mx;
    // every block of the single wave runs the whole time, so the longest block's cycle count is the SM's busy time;
    // the event time x the in-kernel clock gives the same figure from the outside
    const double steps_per_sm = (double)blocks_per_sm * threads * ITERS * UNROLL * ILP;
    printf("%-48s threads/block %4d (resident blocks/SM %d of %d)  %6.2f steps/clk/SM by block cycles, %6.2f by event time x %.0f MHz "
           "(kernel %.3f ms)\n", name, threads, occ, blocks_per_sm, steps_per_sm / mx, steps_per_sm / (ms * 1e-3 * mhz * 1e6), mhz, ms);
    cudaFree(out);
    cudaFree(cyc);
    cudaFree(clk);
    delete[] h;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s  SMs %d  clock %d kHz   (one full wave of 2048 threads per SM, %d independent chains per thread)\n", p.name,
           p.multiProcessorCount, p.clockRate, ILP);
    printf("a step = one counted instruction (first four rows) or one 32-site word pair (last two rows) per lane\n");
    for (int th : {256, 1024}) {
        run<M_LOP3>("LOP3 (ALU pipe), 1 per step", th);
        run<M_POPC>("POPC (XU pipe), 1 per step", th);
        run<M_IADD>("add.u32 (ptxas splits it over IADD3 and IMAD), 1 per step", th);
        run<M_IMAD>("IMAD (FMA pipe), 1 per step", th);
        run<M_MIX_PLAIN>("word pair, plain: 3 LOP3 + 2 POPC + 2 IMAD", th);
        run<M_MIX_CSA>("word pair, carry-save: 5 LOP3 + 1 POPC + 1 IMAD", th);
    }
    printf("word-pair rows: steps = 32-site word pairs; x 32 = cell-sites/clk/SM.  Model inputs of the roofline: LOP3 64, POPC 16 "
           "lane-ops/clk/SM -> plain 16 / 2 = 8.0, carry-save 64 / 5 = 12.8, pipe-balanced optimum 13.71 word pairs/clk/SM.\n");
    return 0;
}
