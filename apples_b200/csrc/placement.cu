// Kernel (b): least-squares placement over every edge of the query's restricted backbone, one query per thread.
//
// Replaces, per query: Subtree (apples/Subtree.py:23-70), Algorithm.dp_frag (apples/Algorithm.py:16-21) with
// FM/OLS/BME/BE all_S_values + all_R_values (FM.py:6-76, OLS.py:12-80, BME.py:6-60, BE.py:6-57),
// placement_per_edge + util.solve2_2 (FM.py:79-93, util.py:6-54), error_per_edge (FM.py:97-124) and
// Algorithm.placement (Algorithm.py:62-101).
//
// Layout: node id = post-order rank (= edge_index).  The observed leaves arrive sorted by id, i.e. left to right.
// Walking leaf i upwards until the first ancestor that also contains leaf i+1 enumerates the valid nodes of the
// restricted subtree in post-order with O(1) state; a small stack of "pending attach nodes" links siblings in
// left-to-right order.  The S moments are accumulated during that same walk, the R moments in one reverse sweep,
// then every edge is solved in closed form and the criterion argmin is taken in post-order (first minimum wins,
// like Python's min()).
//
// One template for the four weightings: moment vector m[6] = sums over leaves of
//   [w, w d, w d^2, w D, w D d, w D^2]      d = path length, D = observed distance,
//   w = 1 (OLS, BME), 1/D^2 (FM), 1/D (BE); BME averages over valid children (BME.py:19,36-37).
// Mapping onto the reference's names is in DESIGN.md.  fp64 operations are written in the reference's order and
// this file is compiled with -fmad=false, so S/R moments, x_1, x_2 are bit-identical to the reference's for
// identical observed distances (the error differs only through Python's pow(x, 2), see DESIGN.md "parity").
#include "common.cuh"

__device__ __forceinline__ void leaf_moments(int method, double D, double* m) {
    m[1] = 0.0; m[2] = 0.0; m[4] = 0.0;
    if (method == APPLES_FM) {          // FM.py:21-26
        m[0] = 1.0 / (D * D); m[3] = 1.0 / D; m[5] = 1.0;
    } else if (method == APPLES_BE) {   // BE.py:10-15
        m[0] = 1.0 / D; m[3] = 1.0; m[5] = D;
    } else {                            // OLS.py:25-30, BME.py:10-15
        m[0] = 1.0; m[3] = D; m[5] = D * D;
    }
}

// moments of a child's leaf set seen from the parent end of the child's edge (FM.py:30-40, OLS.py:34-44, ...)
__device__ __forceinline__ void shifted(const double* m, double ln, double* o) {
    o[0] = m[0];
    o[1] = ln * m[0] + m[1];
    o[2] = m[0] * ln * ln + m[2] + 2.0 * ln * m[1];
    o[3] = m[3];
    o[4] = ln * m[3] + m[4];
    o[5] = m[5];
}

template <bool BME>
__device__ __forceinline__ void accumulate(double* acc, const double* m, double ln, double coef) {
    double sh[6];
    shifted(m, ln, sh);
#pragma unroll
    for (int t = 0; t < 6; ++t) acc[t] = acc[t] + (BME ? coef * sh[t] : sh[t]);
}

struct EdgeSol {
    double x1, x2, err;
    bool int0;
};

// placement_per_edge + solve2_2 + error_per_edge for one node
__device__ __forceinline__ EdgeSol solve_edge(const double* S, const double* R, double ln, int negative_branch) {
    const double a11 = R[0] + S[0];
    const double a12 = R[0] - S[0];
    const double a21 = a12, a22 = a11;
    const double c1 = R[3] + S[3] - ln * S[0] - R[1] - S[1];
    const double c2 = R[3] - S[3] + ln * S[0] - R[1] + S[1];
    const double det = 1.0 / (a11 * a22 - a12 * a21);
    const double x1n = (a22 * c1 - a12 * c2) * det;
    const double x2n = (-a21 * c1 + a11 * c2) * det;
    EdgeSol e;
    e.x1 = x1n;
    e.x2 = x2n;
    e.int0 = false;
    if (!negative_branch) {  // util.py:32-49, same case order and strict inequalities
        if (x1n < 0.0 && x2n < 0.0) {
            e.x1 = 0.0; e.x2 = 0.0; e.int0 = true;
        } else if (x1n > 0.0 && x2n < 0.0) {
            const double t = c1 / a11;
            if (0.0 > t) { e.x1 = 0.0; e.int0 = true; } else { e.x1 = t; }
            e.x2 = 0.0;
        } else if (x1n < 0.0 && 0.0 <= x2n && x2n <= ln) {
            e.x1 = 0.0; e.int0 = true;
            double t = c2 / a22;
            if (0.0 > t) t = 0.0;
            e.x2 = (ln < t) ? ln : t;
        } else if (x1n < 0.0 && x2n > ln) {
            e.x1 = 0.0; e.int0 = true; e.x2 = ln;
        } else if (x1n > 0.0 && x2n > ln) {
            const double t = (c1 - a12 * ln) / a11;
            if (0.0 > t) { e.x1 = 0.0; e.int0 = true; } else { e.x1 = t; }
            e.x2 = ln;
        }
    }
    const double u = e.x1 + e.x2;        // path growth on the R side
    const double v = ln + e.x1 - e.x2;   // path growth on the S side
    const double A = R[5] + S[5];
    const double B = 2.0 * u * R[1] + 2.0 * v * S[1];
    const double C = u * u * R[0] + v * v * S[0];   // (..)**2 in the reference: pow(x, 2), may differ by 1 ulp
    const double D = -2.0 * u * R[3] - 2.0 * v * S[3];
    const double E = -2.0 * R[4] - 2.0 * S[4];
    const double F = R[2] + S[2];
    e.err = A + B + C + D + E + F;
    return e;
}

template <int METHOD>
__global__ void __launch_bounds__(64) place_kernel(const PlaceArgs a) {
    constexpr bool BME = METHOD == APPLES_BME;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const int q = a.qlist ? a.qlist[t] : a.q_begin + t;
    const int slot = a.qlist ? t : q;
    const int st = a.status[q];
    if (st != ST_PLACE) {
        if (st == ST_OVERFLOW) return;  // handled by a later launch
        a.out_edge[q] = (st == ST_ZERO) ? a.zero_edge[q] : -1;
        a.out_error[q] = 0.0;
        a.out_distal[q] = 0.0;
        a.out_pendant[q] = 0.0;
        a.out_status[q] = (st == ST_ZERO) ? APPLES_ZERO_DIST_LEAF : APPLES_TOO_FEW_DISTANCES;
        return;
    }
    if (a.rec_off[t] < 0) return;  // belongs to the other pass (overflow rerun buffers)
    const int K = a.K[q];
    const int* __restrict__ onode = a.obs_node + (size_t)slot * a.cap;
    const double* __restrict__ odist = a.obs_dist + (size_t)slot * a.cap;
    NodeRec* __restrict__ rec = a.recs + a.rec_off[t];
    StackEnt* __restrict__ stk = a.stacks + a.stack_off[t];
    const int* __restrict__ parent = a.tree.parent;
    const int* __restrict__ first = a.tree.first;
    const double* __restrict__ elen = a.tree.elen;

    // ---------------- pass 1: enumerate valid nodes in post-order, link children, accumulate S ----------------
    int p = 0, sp = 0;
    const int leaf0 = onode[0];
    for (int i = 0; i < K; ++i) {
        int u = onode[i];
        const int nxt = (i + 1 < K) ? onode[i + 1] : -1;
        {
            NodeRec& r = rec[p];
            leaf_moments(METHOD, odist[i], r.S);
            r.len = elen[u];
            r.orig = u;
            r.fchild = -1;
            r.rsib = -1;
            r.nchild = 0;
        }
        int prev = p++;
        while (true) {
            const int par = parent[u];
            const bool top = (i + 1 < K) ? (par >= nxt) : (first[par] <= leaf0);
            const bool pending = sp > 0 && stk[sp - 1].A == par;
            if (top) {
                // `prev` is a non-last child of `par`, which a later chain (or the MRCA) owns
                if (pending) {
                    StackEnt& e = stk[sp - 1];
                    rec[e.last].rsib = prev;
                    e.last = prev;
                    e.n++;
                } else {
                    StackEnt& e = stk[sp++];
                    e.A = par; e.first = prev; e.last = prev; e.n = 1;
                }
                break;
            }
            int fc = prev, n = 1;
            if (pending) {
                const StackEnt e = stk[--sp];
                rec[e.last].rsib = prev;
                fc = e.first;
                n = e.n + 1;
            }
            NodeRec& r = rec[p];
            double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            const double coef = BME ? 1.0 / (double)n : 1.0;   // BME.py:19
            for (int c = fc; c >= 0; c = rec[c].rsib) accumulate<BME>(acc, rec[c].S, rec[c].len, coef);
#pragma unroll
            for (int k = 0; k < 6; ++k) r.S[k] = acc[k];
            r.len = elen[par];
            r.orig = par;
            r.fchild = fc;
            r.rsib = -1;
            r.nchild = n;
            prev = p++;
            u = par;
        }
    }
    const int Vn = p;
    // exactly one pending entry is left: the children of the MRCA (the subtree root, not a valid node)
    const int root_first = stk[0].first;
    const int root_n = stk[0].n;

    // ---------------- pass 2: R moments, parents before children (reverse post-order) ----------------
    // all_R_values: siblings in child order, then the parent's R shifted by the parent's edge unless the parent is
    // the subtree root (FM.py:53-76); BME: coefficient 1 / (nonroot + #valid siblings) (BME.py:36-37)
    for (int pp = Vn; pp >= 0; --pp) {
        int fc, n;
        const bool nonroot = pp < Vn;
        if (!nonroot) { fc = root_first; n = root_n; }
        else { fc = rec[pp].fchild; n = rec[pp].nchild; if (fc < 0) continue; }
        const double coef = BME ? 1.0 / (double)((nonroot ? 1 : 0) + n - 1) : 1.0;
        for (int c = fc; c >= 0; c = rec[c].rsib) {
            double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int s = fc; s >= 0; s = rec[s].rsib)
                if (s != c) accumulate<BME>(acc, rec[s].S, rec[s].len, coef);
            if (nonroot) accumulate<BME>(acc, rec[pp].R, rec[pp].len, coef);
#pragma unroll
            for (int k = 0; k < 6; ++k) rec[c].R[k] = acc[k];
        }
    }

    // ---------------- pass 3: per-edge closed-form solve + criterion selection in post-order ----------------
    const bool dbg = a.dbg_x1 != nullptr && q == a.dbg_query;
    int best = -1;
    EdgeSol bs;
    bs.x1 = bs.x2 = bs.err = 0.0;
    bs.int0 = false;
    // HYBRID: the floor(log2 V) smallest errors in (error, order) order (heapq.nsmallest, Algorithm.py:77-79)
    constexpr int HMAX = 32;
    double h_err[HMAX], h_x1[HMAX];
    int h_p[HMAX];
    int hn = 0;
    const int hcap = (a.criterion == APPLES_HYBRID) ? (31 - __clz(Vn)) : 0;
    for (int pp = 0; pp < Vn; ++pp) {
        const NodeRec& r = rec[pp];
        const EdgeSol e = solve_edge(r.S, r.R, r.len, a.negative_branch);
        if (dbg) {
            a.dbg_x1[r.orig] = e.x1;
            a.dbg_x2[r.orig] = e.x2;
            a.dbg_err[r.orig] = e.err;
            a.dbg_valid[r.orig] = 1;
        }
        if (a.criterion == APPLES_MLSE) {
            if (best < 0 || e.err < bs.err) { best = pp; bs = e; }
        } else if (a.criterion == APPLES_ME) {
            if (best < 0 || e.x1 < bs.x1) { best = pp; bs = e; }
        } else {
            if (hn < hcap || e.err < h_err[hn - 1]) {
                int pos = (hn < hcap) ? hn : hn - 1;
                while (pos > 0 && h_err[pos - 1] > e.err) {
                    h_err[pos] = h_err[pos - 1]; h_x1[pos] = h_x1[pos - 1]; h_p[pos] = h_p[pos - 1];
                    --pos;
                }
                h_err[pos] = e.err; h_x1[pos] = e.x1; h_p[pos] = pp;
                if (hn < hcap) ++hn;
            }
        }
    }
    if (a.criterion == APPLES_HYBRID) {
        int bi = 0;
        for (int i = 1; i < hn; ++i)
            if (h_x1[i] < h_x1[bi]) bi = i;
        best = h_p[bi];
        const NodeRec& r = rec[best];
        bs = solve_edge(r.S, r.R, r.len, a.negative_branch);
    }
    const NodeRec& rb = rec[best];
    // Algorithm.py:93-99
    const bool flag = bs.x1 == 0.0 && bs.err > 0.0 && (bs.x2 == 0.0 || bs.x2 == rb.len);
    a.out_edge[q] = rb.orig;
    a.out_error[q] = bs.err;
    a.out_distal[q] = rb.len - bs.x2;
    a.out_pendant[q] = bs.x1;
    a.out_status[q] = (flag ? APPLES_PLACED_MISPLACEMENT_FLAG : APPLES_PLACED) | (bs.int0 ? APPLES_FLAG_PENDANT_INT0 : 0);
}

void launch_place(int method, const PlaceArgs& a, cudaStream_t s) {
    if (a.n <= 0) return;
    const int threads = 64;
    dim3 grid((a.n + threads - 1) / threads), block(threads);
    switch (method) {
        case APPLES_FM: place_kernel<APPLES_FM><<<grid, block, 0, s>>>(a); break;
        case APPLES_BME: place_kernel<APPLES_BME><<<grid, block, 0, s>>>(a); break;
        case APPLES_BE: place_kernel<APPLES_BE><<<grid, block, 0, s>>>(a); break;
        default: place_kernel<APPLES_OLS><<<grid, block, 0, s>>>(a); break;
    }
}
