"""CPU oracle for the APPLES placement hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module, and only as the checker or as the timed CPU arm.  The product (apples_b200/) never imports it and
fails loudly when its CUDA library is missing.

It is a numpy / plain-Python restatement of the reference's algorithm (balabanmetin/apples v2.0.11); every
function cites the reference file:line it follows.  Arithmetic is written in the reference's operation order so
that results are bit-identical to the reference on the same machine.

PINNING: oracle/gen_golden.py imports the UNMODIFIED reference from /root/reference (with oracle/treeswift_shim
standing in for the absent `treeswift` package and clusters supplied as an explicit TSV because `TreeCluster.py`
is absent), runs both on the reference's own data/ files and on seeded synthetic cases, asserts bit-equality and
writes tests/golden/*.json.  tests/test_oracle_golden.py re-checks the oracle against those committed vectors,
against data/dist.mat-derived known answers and against the SURVEY.md section 8(c) known answers.
"""
import heapq
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------------------------------
# tables (distance.py:12-415 BLOSUM45 as used by FastTree2, distance.py:418-678 a2i)
# ----------------------------------------------------------------------------------------------------------------
def _load_blosum45():
    path = os.path.join(_HERE, '..', 'apples_b200', 'data', 'blosum45_fasttree.txt')
    vals = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith('#'):
                continue
            vals.extend(float(x) for x in line.split())
    arr = np.array(vals, dtype=np.float64)
    assert arr.shape == (400,)
    return arr


BLOSUM45 = _load_blosum45()
_AA = 'ARNDCQEGHILKMFPSTWYV'
A2I = np.zeros(256, dtype=np.int64)  # NA = 0 (distance.py:418)
for _i, _c in enumerate(_AA):
    A2I[ord(_c)] = _i
    A2I[ord(_c.lower())] = _i


# ----------------------------------------------------------------------------------------------------------------
# (a1) jc69, (a2) scoredist
# ----------------------------------------------------------------------------------------------------------------
def nuc_counts(a, b):
    """(mismatch, valid) site counts of two 'S1' rows (distance.py:733-737)."""
    both = np.logical_and(a != b'-', b != b'-')
    valid = int(np.count_nonzero(both))
    mism = int(np.count_nonzero(np.logical_and(a != b, both)))
    return mism, valid


def jc69(a, b, overlap_frac):
    """distance.py:718-745."""
    both = np.logical_and(a != b'-', b != b'-')
    valid = np.count_nonzero(both)
    if not valid or valid / len(both) < overlap_frac:
        return -1.0
    p = np.count_nonzero(np.logical_and(a != b, both)) * 1.0 / valid
    if p - np.finfo(float).eps < 0:
        return 0.0
    loc = 1 - (4 * p / 3)
    if 0 >= loc:
        return -1.0
    return -0.75 * np.log(loc)


def jc69_from_counts(mism, valid, L, overlap_frac):
    """jc69 evaluated from the integer counts (same fp64 operations as distance.py:735-745)."""
    if not valid or valid / L < overlap_frac:
        return -1.0
    p = np.float64(mism * 1.0 / valid)
    if p - np.finfo(float).eps < 0:
        return 0.0
    loc = 1 - (4 * p / 3)
    if 0 >= loc:
        return -1.0
    return -0.75 * np.log(loc)


def scoredist(a, b, overlap_frac):
    """distance.py:681-715."""
    both = np.logical_and(a != b'-', b != b'-')
    valid = np.count_nonzero(both)
    if not valid or valid / len(both) < overlap_frac:
        return -1.0
    ia = A2I[a.view(np.uint8)]
    ib = A2I[b.view(np.uint8)]
    tot = np.sum(np.dot(both, BLOSUM45[20 * ia + ib]))
    if 0 >= 1 - tot / valid:
        return -1.0
    cd = -np.log(1 - tot / valid)
    return cd * 1.3


# ----------------------------------------------------------------------------------------------------------------
# tree (prepareTree.py:24-34, util.py:57-88) on top of the oracle's own newick reader
# ----------------------------------------------------------------------------------------------------------------
def load_tree(newick):
    """Returns (tree, name_to_node) with edge_index / level / valid set like prepareTree.py:24-34."""
    sys.path.insert(0, os.path.join(_HERE, 'treeswift_shim'))
    try:
        import treeswift as ts
    finally:
        sys.path.pop(0)
    tree = ts.read_tree(newick, schema='newick')
    k = 0
    for node in tree.traverse_postorder():  # util.py:64-69
        node.edge_index = k
        node.valid = False
        k += 1
    tree.root.level = 0  # util.py:72-88
    frontier = [tree.root]
    while frontier:
        nxt = []
        for n in frontier:
            for c in n.children:
                c.level = n.level + 1
            nxt.extend(n.children)
        frontier = nxt
    names = {}
    for leaf in tree.traverse_postorder(internal=False):
        names[leaf.label] = leaf
    tree.num_nodes = k
    return tree, names


# ----------------------------------------------------------------------------------------------------------------
# (a3) two-phase observed distances, (a4) distance-matrix filter
# ----------------------------------------------------------------------------------------------------------------
def observed_alignment(query, representatives, refs, dist_fn, threshold, baseobs, overlap_frac):
    """Reference.py:117-157.  representatives: list of (consensus 'S1' row, [member names])."""
    heap = []
    for i, (cons, _) in enumerate(representatives):
        d = dist_fn(query, cons, overlap_frac)
        if d >= 0:
            heap.append((d, i))
    heapq.heapify(heap)
    obs = {}
    n_obs = 0
    while heap:
        d, i = heapq.heappop(heap)
        if not (d <= threshold or n_obs < baseobs):
            break
        for name in representatives[i][1]:
            dm = dist_fn(query, refs[name], overlap_frac)
            if not dm < 0:
                obs[name] = dm
                n_obs += 1
    return obs


def observed_matrix(row, name_to_node, threshold, baseobs):
    """PoolQueryWorker.py:44-59.  row: {tag: float} in header order."""
    out = {}
    tx = 0
    for k, v in sorted(row.items(), key=lambda kv: kv[1]):
        if v < 0 or k not in name_to_node:
            continue
        tx += 1
        if tx > baseobs and v > threshold:
            break
        out[k] = v
    return out


# ----------------------------------------------------------------------------------------------------------------
# (a6) restricted subtree, (a7) traversals
# ----------------------------------------------------------------------------------------------------------------
def mark_subtree(obs, name_to_node):
    """Subtree.py:23-43: mark the union of leaf->MRCA paths (MRCA itself unmarked).  Returns (root, marked list).

    Restated as: repeatedly take the deepest unprocessed node, mark it, queue its parent once; the last node left
    is the MRCA.  A bucket-by-level queue replaces the heap; the marked set and the count are the same.
    """
    buckets = {}
    seen = set()
    n_in = 0
    for k in obs:
        if k in name_to_node:
            n = name_to_node[k]
            if id(n) not in seen:
                seen.add(id(n))
                buckets.setdefault(n.level, []).append(n)
                n_in += 1
    marked = []
    remaining = n_in
    lv = max(buckets) if buckets else 0
    while remaining > 1:
        while lv not in buckets or not buckets[lv]:
            lv -= 1
        x = buckets[lv].pop()
        remaining -= 1
        x.valid = True
        marked.append(x)
        p = x.parent
        if id(p) not in seen:
            seen.add(id(p))
            buckets.setdefault(p.level, []).append(p)
            remaining += 1
    while lv not in buckets or not buckets[lv]:
        lv -= 1
    root = buckets[lv].pop()
    return root, marked


def postorder_valid(root):
    """Subtree.py:56-70 restricted to valid nodes (children left to right, then the node); root excluded."""
    out = []
    s1 = [root]
    while s1:
        n = s1.pop()
        out.append(n)
        s1.extend(c for c in n.children if c.valid)
    out.reverse()
    return [n for n in out if n.valid]


def preorder_valid(root):
    """Subtree.py:45-54 restricted to valid nodes."""
    out = []
    s = [root]
    while s:
        n = s.pop()
        s.extend(c for c in n.children if c.valid)
        if n.valid:
            out.append(n)
    return out


# ----------------------------------------------------------------------------------------------------------------
# (a8)-(a11) moments, per-edge solve, error.  One template for the four weightings:
#   moment tuple m = [A0, A1, A2, B0, B1, C0] = sums over leaves of [w, w d, w d^2, w D, w D d, w D^2]
#   OLS w=1 (OLS.py:12-44)   FM w=1/D^2 (FM.py:6-40)   BE w=1/D (BE.py:6-30)   BME = OLS moments averaged over
#   valid children (BME.py:6-30).  The mapping onto the reference's attribute names is in DESIGN.md.
# ----------------------------------------------------------------------------------------------------------------
def leaf_moments(method, D):
    if method == 'FM':
        return [1.0 / (D * D), 0, 0, 1.0 / D, 0, 1]
    if method == 'BE':
        return [1.0 / D, 0, 0, 1, 0, D]
    return [1, 0, 0, D, 0, D * D]  # OLS, BME


def shifted(m, ln):
    """Moments of a child's leaf set seen from its parent end (path lengths grow by the child's edge length `ln`).
    Operation order follows e.g. FM.py:30-40 / OLS.py:34-44."""
    return [
        m[0],
        ln * m[0] + m[1],
        m[0] * ln * ln + m[2] + 2 * ln * m[1],
        m[3],
        ln * m[3] + m[4],
        m[5],
    ]


def compute_moments(method, root, obs, order_post, order_pre):
    """all_S_values then all_R_values (FM.py:6-76, OLS.py:12-80, BE.py:6-57, BME.py:6-60)."""
    bme = method == 'BME'
    for n in order_post:
        if n.is_leaf():
            n.mS = leaf_moments(method, obs[n.label])
        else:
            kids = [c for c in n.children if c.valid]
            coef = 1 / len(kids) if bme else None
            acc = [0, 0, 0, 0, 0, 0]
            for c in kids:
                sh = shifted(c.mS, c.edge_length)
                for t in range(6):
                    acc[t] = acc[t] + (coef * sh[t] if bme else sh[t])
            n.mS = acc
    for n in order_pre:
        par = n.parent
        sibs = [c for c in par.children if c.valid and c is not n]
        nonroot = par is not root and par.valid
        coef = 1 / ((1 if par is not root else 0) + len(sibs)) if bme else None
        acc = [0, 0, 0, 0, 0, 0]
        for c in sibs:
            sh = shifted(c.mS, c.edge_length)
            for t in range(6):
                acc[t] = acc[t] + (coef * sh[t] if bme else sh[t])
        if nonroot:
            sh = shifted(par.mR, par.edge_length)
            for t in range(6):
                acc[t] = acc[t] + (coef * sh[t] if bme else sh[t])
        n.mR = acc


def solve_edge(n, negative_branch):
    """placement_per_edge (FM.py:79-93 etc.) + util.solve2_2 (util.py:6-54)."""
    S, R, ln = n.mS, n.mR, n.edge_length
    a11 = R[0] + S[0]
    a12 = R[0] - S[0]
    a21 = a12
    a22 = a11
    c1 = R[3] + S[3] - ln * S[0] - R[1] - S[1]
    c2 = R[3] - S[3] + ln * S[0] - R[1] + S[1]
    det = 1 / (a11 * a22 - a12 * a21)   # ZeroDivisionError on a singular system, like util.py:26
    assert det != 0                      # util.py:27
    x1n = (a22 * c1 - a12 * c2) * det
    x2n = (-a21 * c1 + a11 * c2) * det
    x1, x2 = x1n, x2n
    if not negative_branch:
        if x1n < 0 and x2n < 0:
            x1, x2 = 0, 0
        elif x1n > 0 and x2n < 0:
            x1, x2 = max(c1 * 1.0 / a11, 0), 0
        elif x1n < 0 and 0 <= x2n and x2n <= ln:
            x1, x2 = 0, min(max(c2 * 1.0 / a22, 0), ln)
        elif x1n < 0 and x2n > ln:
            x1, x2 = 0, ln
        elif x1n > 0 and x2n > ln:
            x1, x2 = max((c1 * 1.0 - a12 * ln) / a11, 0), ln
    n.x1, n.x2, n.x1n, n.x2n = x1, x2, x1n, x2n


def edge_error(n):
    """error_per_edge (FM.py:97-124, OLS.py:101-128, BE.py:72-80, BME.py:75-83)."""
    S, R, ln, x1, x2 = n.mS, n.mR, n.edge_length, n.x1, n.x2
    A = R[5] + S[5]
    B = 2 * (x1 + x2) * R[1] + 2 * (ln + x1 - x2) * S[1]
    C = (x1 + x2) ** 2 * R[0] + (ln + x1 - x2) ** 2 * S[0]
    D = -2 * (x1 + x2) * R[3] - 2 * (ln + x1 - x2) * S[3]
    E = -2 * R[4] - 2 * S[4]
    F = R[2] + S[2]
    return A + B + C + D + E + F


def choose_edge(valids, criterion, num_nodes):
    """Algorithm.placement (Algorithm.py:62-101)."""
    if criterion == 'HYBRID':
        sm = heapq.nsmallest(math.floor(math.log2(num_nodes)), valids, key=edge_error)
        best = min(sm, key=lambda e: e.x1)
    elif criterion == 'ME':
        best = min(valids, key=lambda e: e.x1)
    else:
        best = min(valids, key=edge_error)
    err = edge_error(best)
    flag = 1 if (best.x1 == 0 and err > 0 and (best.x2 == 0 or best.x2 == best.edge_length)) else 0
    return [best.edge_index, err, 1, best.edge_length - best.x2, best.x1], flag


# ----------------------------------------------------------------------------------------------------------------
# (a5) per-query driver
# ----------------------------------------------------------------------------------------------------------------
PLACED, ZERO_DIST_LEAF, TOO_FEW_DISTANCES, PLACED_MISPLACEMENT_FLAG = 0, 1, 2, 3


def place_from_observed(query_name, obs, name_to_node, method='FM', criterion='MLSE', negative_branch=False,
                        exclude_intplace=False, detail=None):
    """PoolQueryWorker.runquery after the distance step (PoolQueryWorker.py:62-141).

    Returns (jplace dict, status).  With `detail` (a dict) also records per-edge x_1, x_2, error keyed by
    edge_index, the observed set and the number of valid nodes.
    """
    jp = {'placements': [{'p': [[0, 0, 1, 0, 0]], 'n': [query_name]}]}
    obs = dict(obs)
    if query_name in name_to_node:
        obs.pop(query_name, None)
        query_name = query_name + '-query'
        jp['placements'][0]['n'] = [query_name]
    if detail is not None:
        detail['observed'] = list(obs.items())
    for k, v in obs.items():
        if v == 0:
            jp['placements'][0]['p'][0][0] = name_to_node[k].edge_index
            return jp, ZERO_DIST_LEAF
    if len(obs) <= 2:
        jp['placements'][0]['p'][0][0] = -1
        return jp, TOO_FEW_DISTANCES
    root, marked = mark_subtree(obs, name_to_node)
    try:
        post = postorder_valid(root)
        pre = preorder_valid(root)
        compute_moments(method, root, obs, post, pre)
        for n in post:
            solve_edge(n, negative_branch)
        res, flag = choose_edge(post, criterion, len(marked))
        if detail is not None:
            detail['num_nodes'] = len(marked)
            detail['edges'] = {n.edge_index: (n.x1, n.x2, edge_error(n)) for n in post}
    finally:
        for n in marked:  # Subtree.unroll_changes (Subtree.py:72-76)
            n.valid = False
    jp['placements'][0]['p'] = [res]
    status = PLACED
    if flag == 1:
        status = PLACED_MISPLACEMENT_FLAG
        if exclude_intplace:
            jp['placements'][0]['p'][0][0] = -1
    return jp, status


class OracleContext:
    """Everything a worker needs (the analogue of PoolQueryWorker's class attributes, PoolQueryWorker.py:17-25)."""

    def __init__(self, tree, name_to_node, refs=None, representatives=None, protein=False, method='FM',
                 criterion='MLSE', negative_branch=False, filt_threshold=0.2, baseobs=25, overlap=0.001,
                 exclude_intplace=False):
        self.tree = tree
        self.name_to_node = name_to_node
        self.refs = refs
        self.representatives = representatives
        self.dist_fn = scoredist if protein else jc69
        self.method = method
        self.criterion = criterion
        self.negative_branch = negative_branch
        self.filt_threshold = filt_threshold
        self.baseobs = baseobs
        self.overlap = overlap
        self.exclude_intplace = exclude_intplace

    def runquery(self, query_name, query_seq, row, detail=None):
        if row:
            obs = observed_matrix(row, self.name_to_node, self.filt_threshold, self.baseobs)
        else:
            obs = observed_alignment(query_seq, self.representatives, self.refs, self.dist_fn, self.filt_threshold,
                                     self.baseobs, self.overlap)
        return place_from_observed(query_name, obs, self.name_to_node, self.method, self.criterion,
                                   self.negative_branch, self.exclude_intplace, detail)


_CTX = None


def _pool_run(args):
    return _CTX.runquery(*args)


def run_pool(ctx, queries, num_thread):
    """pool.starmap(queryworker.runquery, queries) (run_apples.py:94-102) with fork-inherited state."""
    global _CTX
    import multiprocessing as mp
    _CTX = ctx
    queries = list(queries)
    if num_thread <= 1:
        return [ctx.runquery(*q) for q in queries]
    mpctx = mp.get_context('fork')
    with mpctx.Pool(num_thread) as pool:
        return pool.map(_pool_run, queries, chunksize=max(1, len(queries) // (num_thread * 4)))


def _pool_run_detail(args):
    det = {}
    res, status = _CTX.runquery(*args, detail=det)
    return res, status, {'observed': det.get('observed'), 'num_nodes': det.get('num_nodes')}


def run_pool_detail(ctx, queries, num_thread):
    """run_pool that also returns, per query, the status code and the observed set / valid-node count the placement saw
    (for the parity tests at sizes where one process would take minutes)."""
    global _CTX
    import multiprocessing as mp
    _CTX = ctx
    queries = list(queries)
    if num_thread <= 1:
        return [_pool_run_detail(q) for q in queries]
    mpctx = mp.get_context('fork')
    with mpctx.Pool(num_thread) as pool:
        return pool.map(_pool_run_detail, queries, chunksize=1)


# ----------------------------------------------------------------------------------------------------------------
# representatives from an explicit cluster TSV (Reference.py:94-107, PoolRepresentativeWorker.py:16-103)
# ----------------------------------------------------------------------------------------------------------------
def consensus(rows, protein):
    """PoolRepresentativeWorker.py:30-85: column-wise majority over the alphabet (incl. '-'), first max wins."""
    if protein:
        alphabet = np.array(list('ACDEFGHIKLMNPQRSTVWY-'), dtype='S1')
    else:
        alphabet = np.array(list('ACGT-'), dtype='S1')
    mat = np.vstack(rows)
    freq = np.zeros((len(alphabet), mat.shape[1]))
    for i, ch in enumerate(alphabet):
        freq[i] = (mat == ch).sum(axis=0)
    return alphabet[np.argmax(freq, axis=0)]


def representatives_from_tsv(tsv_path, refs, protein):
    """Reference.py:94-107 with the TreeCluster output file given explicitly."""
    import itertools
    with open(tsv_path) as f:
        f.readline()
        lines = [x.strip().split('\t') for x in f.readlines()]
    lines_sorted = sorted(lines, key=lambda x: x[1])
    reps = []
    for key, grp in itertools.groupby(lines_sorted, lambda x: x[1]):
        names = [g[0] for g in grp]
        if key == '-1':
            reps.extend((refs[n], [n]) for n in names)
        else:
            reps.append((consensus([refs[n] for n in names], protein), names))
    return reps
