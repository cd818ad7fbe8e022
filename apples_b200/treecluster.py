"""Max-diameter clustering of the backbone leaves (build-time, host).

The reference shells out to the third-party `TreeCluster.py -m max -t 1.2*threshold` (apples/Reference.py:85-92)
and parses its TSV (`SequenceName<TAB>ClusterNumber`, header line, `-1` = singleton; Reference.py:94-100).  That
program is neither vendored in the reference nor installed here, and nothing in the reference pins its output, so
this is NOT a parity-checked restatement ("parity unpinned"): it is a greedy bottom-up clustering with
the same contract -- every cluster is a connected piece of the tree whose maximum pairwise leaf distance is at
most the threshold -- used when no cluster TSV is supplied and `TreeCluster.py` is not on PATH.  The placement
hot path takes clusters as an explicit input, so oracle and device always see identical clusters.
"""
import numpy as np


def max_diameter_clusters(tree, threshold):
    """Greedy post-order cut.  For every node keep the largest leaf depth of its still-attached part; whenever the
    two deepest attached children (depth + edge) sum to more than `threshold`, cut the deepest child off as a
    cluster and repeat.  Returns a list of clusters, each a list of leaf ids in left-to-right order; whatever is
    still attached to the root at the end forms the last cluster.
    """
    M = tree.num_nodes
    par = tree.parent.tolist()
    el = tree.edge_length.tolist()
    is_leaf = tree.is_leaf.tolist()
    depth = [0.0] * M          # max distance from node to an attached leaf below it
    attached = [True] * M      # node's own subtree part still attached to its parent
    kids = [[] for _ in range(M)]
    # still-attached leaves below each node as a singly linked list in left-to-right order (head, tail, next): joining
    # the lists of the attached children is O(children), cutting a child off hands over its list -- O(M) overall, also on
    # caterpillar backbones (a rescan of the id range [first[c], c] per cut is quadratic there)
    head = [-1] * M
    tail = [-1] * M
    nxt = [-1] * M
    clusters = []

    def take(c):
        out = []
        x = head[c]
        while x >= 0:
            out.append(x)
            x = nxt[x]
        return out

    for u in range(M):
        p = par[u]
        ch = kids[u]
        if ch:
            cand = [(depth[c] + max(el[c], 0.0), c) for c in ch if attached[c]]
            cand.sort()
            while len(cand) >= 2 and cand[-1][0] + cand[-2][0] > threshold:
                _, c = cand.pop()
                attached[c] = False
                leaves = take(c)
                if leaves:
                    clusters.append(leaves)
            if cand:
                depth[u] = cand[-1][0]
                for c in ch:                      # children in left-to-right order
                    if attached[c] and head[c] >= 0:
                        if head[u] < 0:
                            head[u] = head[c]
                        else:
                            nxt[tail[u]] = head[c]
                        tail[u] = tail[c]
            else:
                # every child was cut: u has no attached leaf; drop it from its parent as an empty piece
                attached[u] = False
        elif is_leaf[u]:
            head[u] = tail[u] = u
        kids[u] = None
        if p >= 0:
            kids[p].append(u)
    rest = take(M - 1)
    if rest:
        clusters.append(rest)
    return clusters


def write_cluster_tsv(tree, clusters, path):
    """TreeCluster's output format: header, then `name<TAB>cluster`, singletons as -1, clusters numbered from 1."""
    with open(path, 'w') as f:
        f.write('SequenceName\tClusterNumber\n')
        k = 0
        for cl in clusters:
            if len(cl) == 1:
                f.write('%s\t-1\n' % tree.label[cl[0]])
            else:
                k += 1
                for u in cl:
                    f.write('%s\t%d\n' % (tree.label[u], k))


def read_cluster_tsv(path):
    """Reference.py:94-100: sort lines by the cluster-id STRING (stable), group; returns [(key, [names])]."""
    import itertools
    with open(path) as f:
        f.readline()
        lines = [x.strip().split('\t') for x in f.readlines() if x.strip()]
    lines_sorted = sorted(lines, key=lambda x: x[1])
    return [(key, [i[0] for i in grp]) for key, grp in itertools.groupby(lines_sorted, lambda x: x[1])]
