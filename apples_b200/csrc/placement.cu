// Kernel (b): least-squares placement over every edge of the query's restricted backbone, one WARP per query.
//
// Replaces, per query: Subtree (apples/Subtree.py:23-70), Algorithm.dp_frag (apples/Algorithm.py:16-21) with
// FM/OLS/BME/BE all_S_values + all_R_values (FM.py:6-76, OLS.py:12-80, BME.py:6-60, BE.py:6-57),
// placement_per_edge + util.solve2_2 (FM.py:79-93, util.py:6-54), error_per_edge (FM.py:97-124) and
// Algorithm.placement (Algorithm.py:62-101).
//
// Layout: node id = post-order (DFS) rank = edge_index.  The observed leaves arrive sorted by id, i.e. left to
// right.  "Chain i" = leaf i and its ancestors up to (excluding) the first ancestor that also contains leaf i+1
// (the last chain stops below the MRCA).  Concatenating the chains enumerates the valid nodes of the restricted
// subtree (Subtree.py:23-43) in post-order, so a prefix sum over the chain lengths gives every valid node a compact
// index, and the attach level of a chain (level of the node above its top) describes the whole topology:
//   * the node a chain hangs from lives on the next chain to the right with a strictly smaller attach level,
//   * chains hanging from the same node are consecutive "next smaller-or-equal" neighbours, in child order.
// Lanes work on chains / nodes in parallel: chain discovery and linking, then the S moments level by level from the
// deepest level up, the R moments level by level down, then every edge is solved in closed form and the criterion
// argmin is a warp-shuffle reduction (first minimum in post-order wins, like Python's min()).
//
// One template for the four weightings: moment vector m[6] = sums over leaves of
//   [w, w d, w d^2, w D, w D d, w D^2]      d = path length, D = observed distance,
//   w = 1 (OLS, BME), 1/D^2 (FM), 1/D (BE); BME averages over valid children (BME.py:19,36-37).
// Children are always summed in child order and fp64 operations are written in the reference's order; the file is
// compiled with -fmad=false, so S/R moments, x_1, x_2 are bit-identical to the reference's for identical observed
// distances (the error differs only through Python's pow(x, 2), see DESIGN.md "parity").
#include "common.cuh"

#define FULLMASK 0xffffffffu
#ifndef PL_HIST_V
#define PL_HIST_V 512
#endif
#ifndef PL_MINBLOCKS
#define PL_MINBLOCKS 1
#endif
constexpr int PL_HIST = PL_HIST_V;  // attach-level buckets per warp in shared memory

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
    bool deg;   // the reference raises here: 1 / 0.0 (ZeroDivisionError) or det == 0 (assert), util.py:26-27
};

// placement_per_edge + solve2_2 + error_per_edge for one node
__device__ __forceinline__ EdgeSol solve_edge(const double* S, const double* R, double ln, int negative_branch) {
    const double a11 = R[0] + S[0];
    const double a12 = R[0] - S[0];
    const double a21 = a12, a22 = a11;
    const double c1 = R[3] + S[3] - ln * S[0] - R[1] - S[1];
    const double c2 = R[3] - S[3] + ln * S[0] - R[1] + S[1];
    const double den = a11 * a22 - a12 * a21;
    const double det = 1.0 / den;
    const double x1n = (a22 * c1 - a12 * c2) * det;
    const double x2n = (-a21 * c1 + a11 * c2) * det;
    EdgeSol e;
    e.x1 = x1n;
    e.x2 = x2n;
    e.int0 = false;
    e.deg = den == 0.0 || det == 0.0;
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

// lexicographic (value, index) minimum over the warp; lanes without a candidate pass idx = INT_MAX
__device__ __forceinline__ void warp_argmin(double& val, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULLMASK, val, o);
        const int oi = __shfl_xor_sync(FULLMASK, idx, o);
        const bool take = (oi != 0x7fffffff) && (idx == 0x7fffffff || ov < val || (ov == val && oi < idx));
        if (take) { val = ov; idx = oi; }
    }
}

template <int METHOD>
__global__ void __launch_bounds__(128, PL_MINBLOCKS) place_kernel(const PlaceArgs a) {
    constexpr bool BME = METHOD == APPLES_BME;
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= a.n) return;
    const int q = a.qlist ? a.qlist[t] : a.q_begin + t;
    const int slot = a.qlist ? t : q;
    const int st = a.status[q];
    if (st != ST_PLACE) {
        if (st == ST_OVERFLOW || lane != 0) return;  // overflow: handled by a later launch
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
    const int* __restrict__ olen = a.obs_len + (size_t)slot * a.cap;
    NodeRec* rec = a.recs + a.rec_off[t];       // V valid nodes + 1 pseudo record for the subtree root (the MRCA)
    StackEnt* ch = a.stacks + a.stack_off[t];   // K chains: A = attach level, first = compact offset, last = leaf level
    const int* __restrict__ parent = a.tree.parent;
    const int* __restrict__ level = a.tree.level;
    const double* __restrict__ elen = a.tree.elen;

    // ---------------- chains: attach level, leaf level, compact offsets (prefix sum over chain lengths) ----------------
    int carry = 0, maxlev = 0;
    for (int i0 = 0; i0 < K; i0 += 32) {
        const int i = i0 + lane;
        int len = 0, ll = 0;
        if (i < K) {
            len = olen[i];
            ll = level[onode[i]];
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULLMASK, incl, o);
            if (lane >= o) incl += v;
        }
        if (i < K) {
            ch[i].A = ll - len;
            ch[i].first = carry + incl - len;
            ch[i].last = ll;
        }
        carry += __shfl_sync(FULLMASK, incl, 31);
        maxlev = max(maxlev, ll);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlev = max(maxlev, __shfl_xor_sync(FULLMASK, maxlev, o));
    const int V = carry;
    __syncwarp();
    const int rootlev = ch[K - 1].A;  // the last chain stops below the MRCA

    // ---------------- node records: one lane walks one chain ----------------
    for (int i = lane; i < K; i += 32) {
        int u = onode[i];
        const int off = ch[i].first, len = ch[i].last - ch[i].A;
        for (int s = 0; s < len; ++s) {
            NodeRec& r = rec[off + s];
            r.len = elen[u];
            r.orig = u;
            r.fchild = s ? off + s - 1 : -1;
            r.rsib = -1;
            r.nchild = s ? 1 : 0;
            r.par = off + s + 1;  // the top of the chain is re-linked below
            if (s == 0) leaf_moments(METHOD, odist[i], r.S);
            u = parent[u];
        }
    }
    if (lane == 0) {
        rec[V].fchild = -1;
        rec[V].nchild = 0;
    }
    __syncwarp();

    // ---------------- link every chain top to the node it hangs from ----------------
    for (int j = lane; j < K; j += 32) {
        const int alj = ch[j].A;
        const int top = ch[j].first + (ch[j].last - alj) - 1;
        // next chain to the right with attach level <= mine (next sibling or my owner), then < mine (my owner)
        int nse = j + 1;
        while (nse < K && ch[nse].A > alj) ++nse;
        int own = nse;
        while (own < K && ch[own].A >= alj) ++own;
        const int P = (own < K) ? ch[own].first + (ch[own].last - alj) : V;   // compact index of the node I hang from
        rec[top].par = P;
        if (nse < K && ch[nse].A == alj)
            rec[top].rsib = ch[nse].first + (ch[nse].last - alj) - 1;          // next attached chain of the same node
        else if (own < K)
            rec[top].rsib = P - 1;                                            // the owner chain's own child comes last
        // am I the first child?  (no chain to the left hangs from the same node)
        int pse = j - 1;
        while (pse >= 0 && ch[pse].A > alj) --pse;
        if (pse < 0 || ch[pse].A < alj) rec[P].fchild = top;
        atomicAdd(&rec[P].nchild, 1);
    }
    __syncwarp();

    // ---------------- S and R moments ----------------
    // A chain only depends on chains with a DEEPER attach level (the chains hanging from its nodes) for S, and on the
    // one chain with a shallower attach level that owns the node it hangs from for R.  So chains are bucketed by attach
    // level (counting sort in shared memory); S walks the buckets from the deepest level up, R from the shallowest
    // down, and inside a bucket every lane owns whole chains (sequential along the chain, which is what the recursion
    // is anyway).  Work is O(V + K) instead of O(levels x K).  Trees deeper than PL_HIST levels below the MRCA fall
    // back to a level-by-level sweep.
    auto s_node = [&](int p) {
        NodeRec& r = rec[p];
        const double coef = BME ? 1.0 / (double)r.nchild : 1.0;   // BME.py:19
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int c = r.fchild; c >= 0; c = rec[c].rsib) accumulate<BME>(acc, rec[c].S, rec[c].len, coef);
#pragma unroll
        for (int k = 0; k < 6; ++k) r.S[k] = acc[k];
    };
    // all_R_values: siblings in child order, then the parent's R shifted by the parent's edge unless the parent is
    // the subtree root (FM.py:53-76); BME: coefficient 1 / (nonroot + #valid siblings) (BME.py:36-37)
    auto r_node = [&](int p) {
        NodeRec& r = rec[p];
        const int P = r.par;
        const bool nonroot = P < V;
        const double coef = BME ? 1.0 / (double)((nonroot ? 1 : 0) + rec[P].nchild - 1) : 1.0;
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int sb = rec[P].fchild; sb >= 0; sb = rec[sb].rsib)
            if (sb != p) accumulate<BME>(acc, rec[sb].S, rec[sb].len, coef);
        if (nonroot) accumulate<BME>(acc, rec[P].R, rec[P].len, coef);
#pragma unroll
        for (int k = 0; k < 6; ++k) r.R[k] = acc[k];
    };
    const int range = maxlev - rootlev;  // attach levels lie in [rootlev, maxlev - 1]
    if (range <= PL_HIST) {
        __shared__ int s_hist[4][PL_HIST + 1];
        int* h = s_hist[threadIdx.x >> 5];
        for (int x = lane; x <= range; x += 32) h[x] = 0;
        __syncwarp();
        for (int i = lane; i < K; i += 32) atomicAdd(&h[ch[i].A - rootlev], 1);
        __syncwarp();
        {   // exclusive prefix sum over h[0 .. range): every lane owns a contiguous segment
            const int seg = (range + 31) / 32, b = min(range, lane * seg), e = min(range, b + seg);
            int sum = 0;
            for (int x = b; x < e; ++x) sum += h[x];
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULLMASK, incl, o);
                if (lane >= o) incl += v;
            }
            int run = incl - sum;
            for (int x = b; x < e; ++x) {
                const int c = h[x];
                h[x] = run;
                run += c;
            }
        }
        __syncwarp();
        for (int i = lane; i < K; i += 32) ch[atomicAdd(&h[ch[i].A - rootlev], 1)].n = i;  // h[x] becomes the bucket END
        __syncwarp();
        for (int x = range - 1; x >= 0; --x) {          // S: deepest attach level first
            const int b = x ? h[x - 1] : 0, e = h[x];
            for (int k = b + lane; k < e; k += 32) {
                const int i = ch[k].n;
                const int off = ch[i].first, len = ch[i].last - ch[i].A;
                for (int sdx = 1; sdx < len; ++sdx) s_node(off + sdx);
            }
            if (b != e) __syncwarp();
        }
        for (int x = 0; x < range; ++x) {               // R: shallowest attach level first, each chain top-down
            const int b = x ? h[x - 1] : 0, e = h[x];
            for (int k = b + lane; k < e; k += 32) {
                const int i = ch[k].n;
                const int off = ch[i].first, len = ch[i].last - ch[i].A;
                for (int sdx = len - 1; sdx >= 0; --sdx) r_node(off + sdx);
            }
            if (b != e) __syncwarp();
        }
    } else {
        for (int lv = maxlev - 1; lv > rootlev; --lv) {  // S: level by level upwards
            for (int i = lane; i < K; i += 32) {
                const int al = ch[i].A, ll = ch[i].last;
                if (al < lv && lv < ll) s_node(ch[i].first + (ll - lv));
            }
            __syncwarp();
        }
        for (int lv = rootlev + 1; lv <= maxlev; ++lv) {  // R: level by level downwards
            for (int i = lane; i < K; i += 32) {
                const int al = ch[i].A, ll = ch[i].last;
                if (al < lv && lv <= ll) r_node(ch[i].first + (ll - lv));
            }
            __syncwarp();
        }
    }

    // ---------------- per-edge closed-form solve + criterion selection (first minimum in post-order) ----------------
    const bool dbg = a.dbg_x1 != nullptr && q == a.dbg_query;
    double bval = 0.0;
    int bidx = 0x7fffffff;
    bool degenerate = false;
    for (int p = lane; p < V; p += 32) {
        NodeRec& r = rec[p];
        const EdgeSol e = solve_edge(r.S, r.R, r.len, a.negative_branch);
        degenerate |= e.deg;
        if (dbg) {
            a.dbg_x1[r.orig] = e.x1;
            a.dbg_x2[r.orig] = e.x2;
            a.dbg_err[r.orig] = e.err;
            a.dbg_valid[r.orig] = 1;
        }
        const double key = (a.criterion == APPLES_ME) ? e.x1 : e.err;
        if (bidx == 0x7fffffff || key < bval) { bval = key; bidx = p; }  // ascending p per lane: first minimum kept
    }
    warp_argmin(bval, bidx);
    degenerate = __any_sync(FULLMASK, degenerate);
    int best = bidx;
    if (a.criterion == APPLES_HYBRID) {
        // heapq.nsmallest(floor(log2 V), key=error) in (error, order) order, then the first minimum of x_1 among
        // them (Algorithm.py:77-82): floor(log2 V) rounds of "next smallest (error, index)"
        const int rounds = 31 - __clz(V);
        double last_e = 0.0, best_x1 = 0.0;
        int last_p = -1;
        best = -1;
        for (int rd = 0; rd < rounds; ++rd) {
            double v = 0.0;
            int ix = 0x7fffffff;
            for (int p = lane; p < V; p += 32) {
                const NodeRec& r = rec[p];
                const double e = solve_edge(r.S, r.R, r.len, a.negative_branch).err;
                const bool after = last_p < 0 || e > last_e || (e == last_e && p > last_p);
                if (after && (ix == 0x7fffffff || e < v)) { v = e; ix = p; }
            }
            warp_argmin(v, ix);
            if (ix == 0x7fffffff) break;
            const NodeRec& r = rec[ix];
            const double x1 = solve_edge(r.S, r.R, r.len, a.negative_branch).x1;
            if (best < 0 || x1 < best_x1) { best = ix; best_x1 = x1; }
            last_e = v;
            last_p = ix;
        }
    }
    if (lane == 0) {
        const NodeRec& rb = rec[best];
        const EdgeSol bs = solve_edge(rb.S, rb.R, rb.len, a.negative_branch);
        // Algorithm.py:93-99
        const bool flag = bs.x1 == 0.0 && bs.err > 0.0 && (bs.x2 == 0.0 || bs.x2 == rb.len);
        a.out_edge[q] = rb.orig;
        a.out_error[q] = bs.err;
        a.out_distal[q] = rb.len - bs.x2;
        a.out_pendant[q] = bs.x1;
        a.out_status[q] = (flag ? APPLES_PLACED_MISPLACEMENT_FLAG : APPLES_PLACED) | (bs.int0 ? APPLES_FLAG_PENDANT_INT0 : 0) |
                          (degenerate ? APPLES_FLAG_DEGENERATE : 0);
    }
}

void launch_place(int method, const PlaceArgs& a, cudaStream_t s) {
    if (a.n <= 0) return;
    const int warps = 4;
    dim3 grid((a.n + warps - 1) / warps), block(warps * 32);
    switch (method) {
        case APPLES_FM: place_kernel<APPLES_FM><<<grid, block, 0, s>>>(a); break;
        case APPLES_BME: place_kernel<APPLES_BME><<<grid, block, 0, s>>>(a); break;
        case APPLES_BE: place_kernel<APPLES_BE><<<grid, block, 0, s>>>(a); break;
        default: place_kernel<APPLES_OLS><<<grid, block, 0, s>>>(a); break;
    }
}
