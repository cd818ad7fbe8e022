"""Backbone tree as flat arrays in post-order (DFS) layout.

Host-side replacement for what the reference gets from treeswift + apples/prepareTree.py:9-40 +
apples/util.py:57-88 (index_edges, set_levels) + apples/jutil.py:22-96 (extended_newick).

Layout: the node id IS the reference's `edge_index`, i.e. the post-order rank over all nodes, root last
(util.py:64-69).  With that numbering
  * the children of a node, left to right, have increasing ids,
  * the subtree of node u is the contiguous id range [first[u], u],
  * the leaves of any restricted subtree, sorted by id, are in left-to-right order,
which is what the device kernels rely on ("data layout").
"""
import os
import re
import numpy as np

_TOKEN = re.compile(r"\s*(\(|\)|,|;|:|\[[^\]]*\]|'[^']*'|[^\s\(\),:;\[\]']+)")


class BackboneTree:
    """Flat arrays indexed by edge_index (post-order rank).

    parent[M] int32 (-1 for the root), edge_length[M] float64 (0.0 where the newick has none),
    has_length[M] bool, level[M] int32 (root 0, util.py:72-88), first[M] int32 (smallest id in the subtree),
    nchild[M] int32, label[M] list[str|None], is_rooted bool.
    """

    def __init__(self, parent, edge_length, has_length, label, is_rooted, length_text=None, level=None, first=None):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.edge_length = np.ascontiguousarray(edge_length, dtype=np.float64)
        self.has_length = np.ascontiguousarray(has_length, dtype=bool)
        self.label = list(label)
        self.is_rooted = bool(is_rooted)
        M = len(self.parent)
        self.num_nodes = M
        par = self.parent
        nchild = np.zeros(M, dtype=np.int32)
        np.add.at(nchild, par[par >= 0], 1)
        self.nchild = nchild
        if level is not None and first is not None:   # the native parser computes them with the arrays
            self.level = np.ascontiguousarray(level, dtype=np.int32)
            self.first = np.ascontiguousarray(first, dtype=np.int32)
        else:
            # level: parents have larger ids than children, so one descending sweep suffices
            pl = par.tolist()
            lv = [0] * M
            for u in range(M - 2, -1, -1):
                lv[u] = lv[pl[u]] + 1
            self.level = np.asarray(lv, dtype=np.int32)
            fl = list(range(M))
            for u in range(M - 1):
                p = pl[u]
                if fl[u] < fl[p]:
                    fl[p] = fl[u]
            self.first = np.asarray(fl, dtype=np.int32)
        self.is_leaf = nchild == 0
        self.leaf_ids = np.nonzero(self.is_leaf)[0].astype(np.int32)
        # name -> edge_index of the leaf (prepareTree.py:32-34; later duplicates overwrite, like a dict)
        self.name_to_node = {}
        for u in self.leaf_ids.tolist():
            self.name_to_node[self.label[u]] = u

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_newick(cls, newick, native=None):
        """Parse a newick string or file path (prepareTree.py:24).  The native parser of libapples_b200
        (apples_newick_parse, hostio.cpp) does the work when the library is there and the text is inside its language;
        this function's Python body is the definition of that language and the fallback.  native=True / False force one."""
        if isinstance(newick, (bytes, bytearray)):
            newick = newick.decode()
        if os.path.isfile(os.path.expanduser(newick)):
            with open(os.path.expanduser(newick)) as f:
                s = f.read()
        else:
            s = newick
        if native is not False:
            t = cls._from_newick_native(s)
            if t is not None:
                return t
            if native is True:
                raise ValueError('the native newick parser does not take this text')
        s = s.strip()
        is_rooted = s.startswith('[&R]')
        if s.startswith('['):
            s = s[s.index(']') + 1:]
        # creation (pre-order) arrays
        c_parent = [-1]
        c_label = [None]
        c_len = [None]
        cur = 0
        expect_len = False
        for m in _TOKEN.finditer(s):
            tok = m.group(1)
            c = tok[0]
            if c == '(':
                c_parent.append(cur)
                c_label.append(None)
                c_len.append(None)
                cur = len(c_parent) - 1
            elif c == ',':
                c_parent.append(c_parent[cur])
                c_label.append(None)
                c_len.append(None)
                cur = len(c_parent) - 1
            elif c == ')':
                cur = c_parent[cur]
            elif c == ':':
                expect_len = True
            elif c == ';':
                break
            elif c == '[':
                continue
            else:
                if c == "'":
                    tok = tok[1:-1]
                if expect_len:
                    c_len[cur] = float(tok)
                    expect_len = False
                else:
                    c_label[cur] = tok
        n = len(c_parent)
        # children lists in creation order == left-to-right order
        child_head = [[] for _ in range(n)]
        for v in range(1, n):
            child_head[c_parent[v]].append(v)
        # post-order rank, children left to right then the node (util.py:64-69 via treeswift's two-stack order)
        order = []
        s1 = [0]
        while s1:
            v = s1.pop()
            order.append(v)
            s1.extend(child_head[v])
        order.reverse()
        rank = [0] * n
        for r, v in enumerate(order):
            rank[v] = r
        parent = np.full(n, -1, dtype=np.int32)
        elen = np.zeros(n, dtype=np.float64)
        has = np.zeros(n, dtype=bool)
        label = [None] * n
        for v in range(n):
            r = rank[v]
            if c_parent[v] >= 0:
                parent[r] = rank[c_parent[v]]
            if c_len[v] is not None:
                elen[r] = c_len[v]
                has[r] = True
            label[r] = c_label[v]
        return cls(parent, elen, has, label, is_rooted)

    @classmethod
    def _from_newick_native(cls, s):
        """BackboneTree through apples_newick_parse, or None when the library is missing or declines the text."""
        import ctypes as C
        try:
            from . import _lib
            lib = _lib.load()
        except (RuntimeError, OSError):
            return None
        raw = s.encode('utf-8', 'surrogateescape') if isinstance(s, str) else bytes(s)
        h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = lib.apples_newick_parse(raw, len(raw), C.byref(h), err, 256)
        if rc == _lib.NEWICK_UNSUPPORTED:
            return None
        if rc != 0:
            raise ValueError(err.value.decode())
        try:
            n = int(lib.apples_newick_nodes(h))

            def arr(fn, dtype, count):
                p = getattr(lib, 'apples_newick_' + fn)(h)
                return np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(count,)).copy()
            parent, level, first = arr('parent', np.int32, n), arr('level', np.int32, n), arr('first', np.int32, n)
            elen, has = arr('edge_length', np.float64, n), arr('has_length', np.uint8, n).astype(bool)
            has_label = arr('has_label', np.uint8, n)
            off = arr('label_offsets', np.int64, n + 1)
            text = C.string_at(lib.apples_newick_labels(h), int(off[-1])).decode('utf-8', 'surrogateescape')
            if len(text) == int(off[-1]):      # ASCII labels: byte offsets are character offsets
                o = off.tolist()
                label = [text[o[i]:o[i + 1]] if f else None for i, f in enumerate(has_label.tolist())]
            else:
                b = text.encode('utf-8', 'surrogateescape')
                o = off.tolist()
                label = [b[o[i]:o[i + 1]].decode('utf-8', 'surrogateescape') if f else None for i, f in enumerate(has_label.tolist())]
            rooted = bool(lib.apples_newick_rooted(h))
        finally:
            lib.apples_newick_free(h)
        return cls(parent, elen, has, label, rooted, level=level, first=first)

    # ------------------------------------------------------------------ queries
    def children_of(self, u):
        """Children of u, left to right (increasing id)."""
        out = []
        c = u - 1
        f = int(self.first[u])
        while c >= f:
            out.append(c)
            c = int(self.first[c]) - 1
        out.reverse()
        return out

    def extended_newick(self, native=None):
        """Newick with `{edge_index}` after every non-root node (jutil.py:22-96), same number formatting.  Written by
        apples_newick_extended (hostio.cpp) when the library is there; the Python body below is the definition."""
        if native is not False:
            s = self._extended_newick_native()
            if s is not None:
                return s
            if native is True:
                raise ValueError('the native extended-newick writer declined this tree')
        M = self.num_nodes
        strs = [None] * M
        pend = [[] for _ in range(M)]
        par = self.parent.tolist()
        el = self.edge_length.tolist()
        has = self.has_length.tolist()
        for u in range(M):
            kids = pend[u]
            if not kids:
                s = '' if self.label[u] is None else str(self.label[u])
            else:
                s = '(' + ','.join(kids) + ')'
                if self.label[u] is not None:
                    s += str(self.label[u])
                pend[u] = None
            p = par[u]
            if p >= 0:
                if has[u]:
                    x = el[u]
                    l_str = str(int(x)) if x.is_integer() else str(x)
                    s += ':%s' % l_str
                s += '{%d}' % u
                pend[p].append(s)
            else:
                strs[u] = s
        root = strs[M - 1]
        if self.is_rooted:
            return '[&R] %s;' % root
        return '%s;' % root


def _extended_newick_native(self):
    import ctypes as C
    try:
        from . import _lib
        lib = _lib.load()
    except (RuntimeError, OSError):
        return None
    M = self.num_nodes
    enc = [None if l is None else str(l).encode('utf-8', 'surrogateescape') for l in self.label]
    has_label = np.fromiter((l is not None for l in enc), dtype=np.uint8, count=M)
    off = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.fromiter((0 if l is None else len(l) for l in enc), dtype=np.int64, count=M), out=off[1:])
    blob = b''.join(l for l in enc if l is not None)
    has_len = np.ascontiguousarray(self.has_length, dtype=np.uint8)
    out, n = C.c_void_p(), C.c_int64(0)
    err = C.create_string_buffer(256)
    rc = lib.apples_newick_extended(M, _lib.ptr(self.parent), _lib.ptr(self.edge_length), _lib.ptr(has_len), blob,
                                    _lib.ptr(off), _lib.ptr(has_label), 1 if self.is_rooted else 0, C.byref(out), C.byref(n),
                                    err, 256)
    if rc == _lib.NEWICK_UNSUPPORTED:
        return None
    if rc != 0:
        raise ValueError(err.value.decode())
    try:
        return C.string_at(out, n.value).decode('utf-8', 'surrogateescape')
    finally:
        lib.apples_free_text(out)


BackboneTree._extended_newick_native = _extended_newick_native


def prepare_tree(tree_fp):
    """Mirror of prepareTree (prepareTree.py:9-40) without backbone re-estimation (always `-D` here; the FastTree
    re-estimation step is out of the hot path).

    Returns (tree, name_to_node_map, extended_newick_string) with name_to_node_map: leaf label -> edge_index.
    """
    tree = BackboneTree.from_newick(tree_fp)
    return tree, tree.name_to_node, tree.extended_newick()
