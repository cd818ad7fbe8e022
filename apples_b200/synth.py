"""Seeded synthetic inputs of the shapes named in BASELINE.json.

The reference ships no generator; this one is shared by tests, bench.py and oracle/gen_golden.py so that the
oracle and the device path always see identical inputs.  numpy only (deterministic for a given seed).

  random_tree      -- random topology by sequential random-edge insertion (Yule-like), 3-child root, leaves named
                      L%07d, edge lengths ~ Exp(mean) floored at 1e-6; optional polytomies / zero / negative edges
  evolve_alignment -- root sequence uniform, evolved down the tree under JC69 (nucleotide) or a uniform-replacement
                      model (amino acid); 3 % of cells set to '-', plus a leading/trailing gap run of U(0, 0.1 L)
  make_queries     -- each query = copy of a uniformly chosen leaf evolved by a further Exp(mean 0.05), fresh gaps
"""
import numpy as np

NUC_ALPHABET = np.frombuffer(b'ACGT', dtype=np.uint8)
AA_ALPHABET = np.frombuffer(b'ARNDCQEGHILKMFPSTWYV', dtype=np.uint8)


def random_tree(n_leaves, seed, mean_edge=0.02, polytomy_frac=0.0, zero_frac=0.0, neg_frac=0.0, prefix='L',
                model='yule'):
    """Returns a newick string.  model='yule': every insertion subdivides a uniformly chosen PENDANT edge (depth
    grows like log N, as in real backbones); model='uniform': a uniformly chosen edge (depth grows like sqrt N)."""
    rng = np.random.default_rng(seed)
    n = int(n_leaves)
    assert n >= 3
    # node arrays; node 0 = root with three leaf children 1,2,3
    cap = 2 * n + 2
    parent = np.full(cap, -1, dtype=np.int64)
    is_leaf = np.zeros(cap, dtype=bool)
    parent[1:4] = 0
    is_leaf[1:4] = True
    cnt = 4
    picks = rng.random(n)  # one uniform per insertion
    par = parent
    leaf_list = [1, 2, 3]
    for k in range(3, n):
        # choose a random non-root node e (= the edge above it), subdivide it and hang a new leaf
        if model == 'yule':
            e = leaf_list[int(picks[k] * len(leaf_list))]
            leaf_list.append(cnt + 1)
        else:
            e = 1 + int(picks[k] * (cnt - 1))
        mid = cnt
        leaf = cnt + 1
        cnt += 2
        par[mid] = par[e]
        par[e] = mid
        par[leaf] = mid
        is_leaf[leaf] = True
    m = cnt
    parent = parent[:m]
    is_leaf = is_leaf[:m]
    el = np.maximum(rng.exponential(mean_edge, m), 1e-6)
    el = np.round(el, 7)
    u = rng.random(m)
    if zero_frac > 0:
        el[u < zero_frac] = 0.0
    if neg_frac > 0:
        el[(u >= zero_frac) & (u < zero_frac + neg_frac)] *= -0.1
    # contract a fraction of internal edges to make polytomies
    if polytomy_frac > 0:
        v = rng.random(m)
        for x in range(1, m):
            if not is_leaf[x] and v[x] < polytomy_frac:
                parent[parent == x] = parent[x]
                parent[x] = -2  # removed
    kids = [[] for _ in range(m)]
    for x in range(1, m):
        if parent[x] >= 0:
            kids[parent[x]].append(x)
    # random child order so that leaf numbering is not correlated with topology
    names = {}
    leaf_ids = [x for x in range(m) if is_leaf[x] and parent[x] != -2]
    perm = rng.permutation(len(leaf_ids))
    for i, x in enumerate(leaf_ids):
        names[x] = '%s%07d' % (prefix, perm[i])
    # iterative newick
    out = {}
    stack = [(0, False)]
    while stack:
        x, done = stack.pop()
        if not done:
            if kids[x]:
                stack.append((x, True))
                for c in reversed(kids[x]):
                    stack.append((c, False))
            else:
                out[x] = '%s:%s' % (names[x], _fmt(el[x]))
        else:
            s = '(' + ','.join(out.pop(c) for c in kids[x]) + ')'
            out[x] = s if x == 0 else '%s:%s' % (s, _fmt(el[x]))
    return out[0] + ';'


def _fmt(x):
    return ('%.7f' % x).rstrip('0').rstrip('.') if x != 0 else '0'


def _sub_prob(t, n_states):
    # probability that a site shows a different state after branch length t (JC-type model with n states)
    a = n_states / (n_states - 1.0)
    return (1.0 / a) * (1.0 - np.exp(-a * np.maximum(t, 0.0)))


def _mutate(rng, seq, t, n_states):
    p = _sub_prob(t, n_states)
    hit = rng.random(seq.shape[-1]) < p
    k = int(hit.sum())
    if k:
        seq = seq.copy()
        seq[hit] = (seq[hit] + rng.integers(1, n_states, k)) % n_states
    return seq


def _gapify(rng, chars, L, gap_frac=0.03, edge_frac=0.1):
    g = rng.random(L) < gap_frac
    a = int(rng.integers(0, int(edge_frac * L) + 1))
    b = int(rng.integers(0, int(edge_frac * L) + 1))
    chars = chars.copy()
    chars[g] = ord('-')
    if a:
        chars[:a] = ord('-')
    if b:
        chars[L - b:] = ord('-')
    return chars


def evolve_alignment(tree, L, seed, protein=False, gap_frac=0.03, edge_frac=0.1):
    """tree: apples_b200.tree.BackboneTree.  Returns ({leaf name: 'S1' row}, state matrix of the leaves by node id)."""
    rng = np.random.default_rng(seed)
    ns = 20 if protein else 4
    alpha = AA_ALPHABET if protein else NUC_ALPHABET
    M = tree.num_nodes
    states = [None] * M
    states[M - 1] = rng.integers(0, ns, L)
    par = tree.parent.tolist()
    el = tree.edge_length.tolist()
    refs = {}
    leaf_states = {}
    for u in range(M - 2, -1, -1):  # parents have larger ids
        states[u] = _mutate(rng, states[par[u]], el[u], ns)
    for u in range(M):
        if tree.is_leaf[u]:
            leaf_states[u] = states[u]
        else:
            states[u] = None
    for u in sorted(leaf_states):
        refs[tree.label[u]] = _gapify(rng, alpha[leaf_states[u]], L, gap_frac, edge_frac).view('S1')
    return refs, leaf_states


def make_queries(tree, leaf_states, n_queries, seed, protein=False, mean_extra=0.05, gap_frac=0.03, edge_frac=0.1,
                 prefix='Q'):
    """Returns {query name: 'S1' row} and the list of source leaf ids."""
    rng = np.random.default_rng(seed)
    ns = 20 if protein else 4
    alpha = AA_ALPHABET if protein else NUC_ALPHABET
    leaves = sorted(leaf_states)
    src = rng.integers(0, len(leaves), n_queries)
    extra = rng.exponential(mean_extra, n_queries)
    out = {}
    L = len(leaf_states[leaves[0]])
    for i in range(n_queries):
        s = _mutate(rng, leaf_states[leaves[src[i]]], extra[i], ns)
        out['%s%07d' % (prefix, i)] = _gapify(rng, alpha[s], L, gap_frac, edge_frac).view('S1')
    return out, [leaves[j] for j in src]


def write_fasta(seqs, path):
    with open(path, 'w') as f:
        for name, row in seqs.items():
            f.write('>%s\n%s\n' % (name, row.tobytes().decode()))
