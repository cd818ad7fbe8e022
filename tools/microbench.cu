// Integer-pipe microbenchmark for the distance-kernel roofline: POPC, LOP3, IADD3 lane-ops per clock per SM on the
// device it runs on, and the mix the dense kernel issues (2 POPC : 4 LOP3 : 2 IADD per 32-site word pair).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8;
constexpr int ITERS = 4096;

template <int MODE>
__global__ void k(unsigned* out, unsigned seed, long long* cycles) {
    unsigned x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = seed + threadIdx.x * 7 + i * 13; y[i] = seed ^ (i * 0x9e3779b9u); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) {            // POPC chain
                x[i] = __popc(x[i]) ^ y[i];
            } else if (MODE == 1) {     // LOP3 chain (3-input logic)
                x[i] = (x[i] & y[i]) ^ (y[i] | seed);
                asm volatile("" : "+r"(x[i]));
            } else if (MODE == 2) {     // IADD3
                x[i] = x[i] + y[i] + seed;
                asm volatile("" : "+r"(x[i]));
            } else {                    // dense-kernel mix for one word pair: 4 LOP3, 2 POPC, 2 IADD
                unsigned v = x[i] & y[i];
                unsigned m = ((x[i] ^ seed) | (y[i] ^ it)) & v;
                x[i] += __popc(m);
                y[i] += __popc(v);
            }
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_inner, int blocks_per_sm, int threads) {
    int dev = 0;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    int nb = p.multiProcessorCount * blocks_per_sm;
    unsigned* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)nb * threads * 4);
    cudaMalloc(&cyc, nb * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<nb, threads>>>(out, 12345u, cyc);
    cudaEventRecord(a);
    k<MODE><<<nb, threads>>>(out, 12345u, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    long long* h = new long long[nb];
    cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nb; ++i) avg += h[i];
    avg /= nb;
    double lane_ops_per_sm = (double)blocks_per_sm * threads * ITERS * ILP * ops_per_inner;
    printf("%-28s blocks/SM=%d threads=%d  %.1f lane-ops/clk/SM (per-block cycles %.0f)  kernel %.3f ms -> %.2f Tops/s chip\n",
           name, blocks_per_sm, threads, lane_ops_per_sm / avg, avg, ms,
           lane_ops_per_sm * p.multiProcessorCount / (ms * 1e-3) / 1e12);
    cudaFree(out);
    cudaFree(cyc);
    delete[] h;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s  SMs %d  clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int th : {256, 1024}) {
        run<0>("POPC (+1 LOP per op)", 1, 2048 / th, th);
        run<1>("LOP3 x2", 2, 2048 / th, th);
        run<2>("IADD3", 1, 2048 / th, th);
        run<3>("mix 4 LOP3 + 2 POPC + 2 IADD", 1, 2048 / th, th);
    }
    printf("mix row: lane-ops = 32-site word pairs; x32 = cell-sites/clk/SM\n");
    return 0;
}
