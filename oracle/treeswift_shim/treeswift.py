"""Minimal stand-in for the third-party `treeswift` package (TEST INFRASTRUCTURE ONLY).

`treeswift` is an unpinned, un-vendored dependency of the reference (setup.py:20) and is not installed in
this image.  The reference touches only a tiny surface of it on the placement hot path
(apples/prepareTree.py:24-34, apples/util.py:57-88, apples/jutil.py:22-96, apples/Subtree.py,
apples/PrioritySet.py:21): `read_tree(path, schema='newick')`, `Tree.root`, `Tree.is_rooted`,
`Tree.traverse_postorder(leaves=, internal=)`, and on nodes `children`, `parent`, `edge_length`, `label`,
`is_leaf()`, `traverse_postorder()`, `__lt__`.

This shim provides exactly that surface so the UNMODIFIED reference under /root/reference can be imported and
run in this container by `oracle/gen_golden.py` to generate golden vectors.  Nothing in the product package
imports it.  Post-order is the two-stack order (children left to right, then the node), which is what the
reference's own Subtree.traverse_postorder uses (apples/Subtree.py:56-70).
"""
import os


class Node:
    __slots__ = ('children', 'parent', 'edge_length', 'label', '__dict__')

    def __init__(self, label=None, edge_length=None):
        self.children = []
        self.parent = None
        self.edge_length = edge_length
        self.label = label

    def is_leaf(self):
        return len(self.children) == 0

    def is_root(self):
        return self.parent is None

    def add_child(self, child):
        self.children.append(child)
        child.parent = self

    def __lt__(self, other):
        # only used to break priority ties inside a heap (apples/PrioritySet.py:21); the result does not
        # depend on the order
        a = '' if self.label is None else str(self.label)
        b = '' if other.label is None else str(other.label)
        return a < b

    def traverse_postorder(self, leaves=True, internal=True):
        s1 = [self]
        s2 = []
        while s1:
            n = s1.pop()
            s2.append(n)
            s1.extend(n.children)
        while s2:
            n = s2.pop()
            if (leaves and n.is_leaf()) or (internal and not n.is_leaf()):
                yield n

    def traverse_preorder(self, leaves=True, internal=True):
        s = [self]
        while s:
            n = s.pop()
            if (leaves and n.is_leaf()) or (internal and not n.is_leaf()):
                yield n
            s.extend(reversed(n.children))


class Tree:
    def __init__(self, is_rooted=True):
        self.root = Node()
        self.is_rooted = is_rooted

    def traverse_postorder(self, leaves=True, internal=True):
        return self.root.traverse_postorder(leaves=leaves, internal=internal)

    def traverse_preorder(self, leaves=True, internal=True):
        return self.root.traverse_preorder(leaves=leaves, internal=internal)


def _parse_newick(ts):
    ts = ts.strip()
    t = Tree()
    # treeswift marks a tree rooted only when the string carries the '[&R]' tag; data/prot/out.jplace (written
    # by the reference from an untagged newick) has no '[&R] ' prefix, which pins this behaviour
    t.is_rooted = ts.startswith('[&R]')
    if ts.startswith('['):
        ts = ts[ts.index(']') + 1:].strip()
    n = t.root
    i = 0
    ln = len(ts)
    parse_length = False
    while i < ln:
        c = ts[i]
        if c == ';':
            break
        if c == '(':
            child = Node()
            n.add_child(child)
            n = child
            i += 1
        elif c == ')':
            n = n.parent
            i += 1
        elif c == ',':
            sib = Node()
            n.parent.add_child(sib)
            n = sib
            i += 1
        elif c == ':':
            parse_length = True
            i += 1
        elif c == '[':
            # comment: skip
            j = ts.index(']', i)
            i = j + 1
        elif c in ' \t\r\n':
            i += 1
        else:
            if c == "'":
                j = ts.index("'", i + 1)
                tok = ts[i + 1:j]
                i = j + 1
            else:
                j = i
                while j < ln and ts[j] not in '(),:;[':
                    j += 1
                tok = ts[i:j].strip()
                i = j
            if parse_length:
                n.edge_length = float(tok)
                parse_length = False
            else:
                n.label = tok
    return t


def read_tree_newick(newick):
    if os.path.isfile(os.path.expanduser(newick)):
        with open(os.path.expanduser(newick)) as f:
            s = f.read()
    else:
        s = newick
    return _parse_newick(s)


def read_tree(input, schema):
    if schema.lower() != 'newick':
        raise ValueError('shim only reads newick')
    return read_tree_newick(input)
