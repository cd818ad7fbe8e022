// the data formats either side of the hot path.
//
// (f1) packer: alignment bytes as fasta2dic leaves them (apples/fasta2dic.py:42-72: upper-case letters, '-' for gaps and
//      for every letter outside the alphabet) -> the device layout of DESIGN.md: nucleotide bit-planes (lo, hi, valid)
//      or one amino-acid code per site (a2i order, apples/distance.py:418-678).  One thread per 32-site word; HBM-bound
//      (1 byte read + 3/8 byte written per site).  Nucleotide bytes other than A,C,G,T,- cannot be expressed by the
//      2-bit code (the reference would treat them as a fifth symbol, distance.py:733-737) and raise an error flag.
// (f2) representative consensus: column-wise majority over a cluster's members in the reference's alphabet order with
//      first-maximum tie-break (apples/PoolRepresentativeWorker.py:30-85): [A,C,G,T,-] or the 20 amino acids + '-'.
#include "common.cuh"

__global__ void pack_nuc_kernel(const uint8_t* __restrict__ bytes, int64_t row_stride, int n, int L, int W,
                                uint32_t* __restrict__ out, int* __restrict__ bad, int* __restrict__ row_flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * W) return;
    const int row = (int)(t / W), w = (int)(t % W);
    const uint8_t* src = bytes + (size_t)row * row_stride + (size_t)w * 32;
    uint32_t lo = 0, hi = 0, va = 0;
    const int nsite = min(32, L - w * 32);
    for (int s = 0; s < nsite; ++s) {
        const uint32_t c = src[s];
        const uint32_t x = (c >> 1) & 3u;          // A 0, C 1, T 2, G 3
        const uint32_t code = x ^ (x >> 1);         // A 0, C 1, G 2, T 3
        const bool base = c == 'A' || c == 'C' || c == 'G' || c == 'T';
        if (base) {
            lo |= (code & 1u) << s;
            hi |= (code >> 1) << s;
            va |= 1u << s;
        } else if (c != '-') {
            // a byte the 2-bit code cannot carry: the row goes through the byte-compare fallback (distance.cu)
            *bad = 1;
            if (row_flag) row_flag[row] = 1;
        }
    }
    uint32_t* o = out + (size_t)row * 3 * W + w;
    o[0] = lo;
    o[W] = hi;
    o[2 * W] = va;
}

// byte -> amino-acid code, built at compile time so that the module's static initialiser puts it into the constant
// memory of EVERY device the library is used on (no per-call upload, no synchronisation)
struct AaCodeTable {
    uint8_t t[256];
    constexpr AaCodeTable() : t{} {
        const char order[21] = "ARNDCQEGHILKMFPSTWYV";
        for (int i = 0; i < 256; ++i) t[i] = 0;  // NA = 0: unknown bytes count as 'A' (distance.py:418)
        for (int i = 0; i < 20; ++i) {
            t[(unsigned char)order[i]] = (uint8_t)i;
            t[(unsigned char)(order[i] + 32)] = (uint8_t)i;
        }
        t[(unsigned char)'-'] = 20;
    }
};
__constant__ AaCodeTable c_aa_code = AaCodeTable();

__global__ void pack_aa_kernel(const uint8_t* __restrict__ bytes, int64_t row_stride, int n, int L, int Lp,
                               uint8_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * Lp) return;
    const int row = (int)(t / Lp), s = (int)(t % Lp);
    out[t] = s < L ? c_aa_code.t[bytes[(size_t)row * row_stride + s]] : (uint8_t)20;
}

cudaError_t launch_pack(int kind, const uint8_t* bytes, int64_t row_stride, int n, int L, void* out, int* bad,
                        cudaStream_t s, int* row_flag) {
    if (n <= 0) return cudaSuccess;
    if (kind == APPLES_NUC) {
        const int W = apples_words_per_row(L);
        const int64_t total = (int64_t)n * W;
        pack_nuc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(bytes, row_stride, n, L, W, (uint32_t*)out, bad, row_flag);
    } else {
        const int Lp = apples_aa_row_bytes(L);
        const int64_t total = (int64_t)n * Lp;
        pack_aa_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(bytes, row_stride, n, L, Lp, (uint8_t*)out);
    }
    return cudaGetLastError();
}

// one block per cluster, one thread per column (strided): counts over the members, first maximum in alphabet order
template <int NSYM>
__global__ void consensus_kernel(const uint8_t* __restrict__ bytes, int64_t row_stride, int L, const int* __restrict__ goff,
                                 const int* __restrict__ gmem, uint8_t* __restrict__ out, int64_t out_stride) {
    const int c = blockIdx.x;
    const int b = goff[c], e = goff[c + 1];
    const char* alpha = NSYM == 5 ? "ACGT-" : "ACDEFGHIKLMNPQRSTVWY-";   // PoolRepresentativeWorker.py:34-61
    for (int col = threadIdx.x; col < L; col += blockDim.x) {
        int cnt[NSYM];
#pragma unroll
        for (int k = 0; k < NSYM; ++k) cnt[k] = 0;
        for (int x = b; x < e; ++x) {
            const uint8_t ch = bytes[(size_t)gmem[x] * row_stride + col];
#pragma unroll
            for (int k = 0; k < NSYM; ++k) cnt[k] += (ch == (uint8_t)alpha[k]) ? 1 : 0;
        }
        int best = 0;
#pragma unroll
        for (int k = 1; k < NSYM; ++k)
            if (cnt[k] > cnt[best]) best = k;   // strict: the first maximum wins (np.argmax)
        out[(size_t)c * out_stride + col] = (uint8_t)alpha[best];
    }
}

cudaError_t launch_consensus(int kind, const uint8_t* bytes, int64_t row_stride, int L, int n_rep, const int* goff,
                             const int* gmem, uint8_t* out, int64_t out_stride, cudaStream_t s) {
    if (n_rep <= 0) return cudaSuccess;
    if (kind == APPLES_NUC)
        consensus_kernel<5><<<n_rep, 256, 0, s>>>(bytes, row_stride, L, goff, gmem, out, out_stride);
    else
        consensus_kernel<21><<<n_rep, 256, 0, s>>>(bytes, row_stride, L, goff, gmem, out, out_stride);
    return cudaGetLastError();
}

// gathers whole rows: dst[j] = src[idx[j]] (row_bytes a multiple of 16); used to collect the queries of a rerun
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ idx, uint4* __restrict__ dst,
                                   int n, int row_vec) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * row_vec) return;
    const int j = (int)(t / row_vec), c = (int)(t % row_vec);
    dst[t] = src[(size_t)idx[j] * row_vec + c];
}

cudaError_t launch_gather_rows(const void* src, const int* idx, void* dst, int n, size_t row_bytes, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int row_vec = (int)(row_bytes / 16);
    const int64_t total = (int64_t)n * row_vec;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const uint4*)src, idx, (uint4*)dst, n, row_vec);
    return cudaGetLastError();
}


// byte rows [n][src_stride] -> [n][Lp] with the tail padded by '-' (Lp a multiple of 16): the operand layout of the
// byte-compare fallback (whole 32-bit words, gaps do not count)
__global__ void repitch_bytes_kernel(const uint8_t* __restrict__ src, int64_t src_stride, int n, int L, int Lp,
                                     uint8_t* __restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * Lp) return;
    const int row = (int)(t / Lp), s = (int)(t % Lp);
    dst[t] = s < L ? src[(size_t)row * src_stride + s] : (uint8_t)'-';
}

cudaError_t launch_repitch_bytes(const uint8_t* src, int64_t src_stride, int n, int L, int Lp, uint8_t* dst, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = (int64_t)n * Lp;
    repitch_bytes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, src_stride, n, L, Lp, dst);
    return cudaGetLastError();
}

// bit-planes [n][3][W] -> bytes [n][Lp] ('A','C','G','T','-'; padding '-')
__global__ void unpack_nuc_kernel(const uint32_t* __restrict__ planes, int n, int L, int W, int Lp, uint8_t* __restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * Lp) return;
    const int row = (int)(t / Lp), s = (int)(t % Lp);
    uint8_t c = '-';
    if (s < L) {
        const uint32_t* p = planes + (size_t)row * 3 * W + s / 32;
        const uint32_t b = 1u << (s % 32);
        if (p[2 * W] & b) c = "ACGT"[((p[0] & b) ? 1 : 0) | ((p[W] & b) ? 2 : 0)];
    }
    dst[t] = c;
}

cudaError_t launch_unpack_nuc(const uint32_t* planes, int n, int L, int W, int Lp, uint8_t* dst, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = (int64_t)n * Lp;
    unpack_nuc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(planes, n, L, W, Lp, dst);
    return cudaGetLastError();
}
