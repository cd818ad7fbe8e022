"""Reduced reference: aligned backbone sequences, their clusters and one consensus representative per cluster.

Host-side mirror of apples/Reference.py:65-115 (`ReducedReference.__init__`, `set_baseobs`) and
apples/PoolRepresentativeWorker.py:16-103.  The per-query method `get_obs_dist` (Reference.py:117-157) is what the
device selection kernel replaces; it is intentionally absent here -- `apples_b200.placer.place_batch` is the
entry point and it fails loudly without the CUDA library.

Attributes kept from the reference: `refs` ({name: 'S1' row}), `prot_flag`, `threshold`, `baseobs`,
`representatives` ([(consensus 'S1' row, [member names])] in the reference's list order, which is the tie-break
key of the selection).
"""
import logging
import os
import shutil
import subprocess
import tempfile
import time

import numpy as np

from . import fasta as _fasta
from . import treecluster as _tc


def consensus_rows(mat, prot_flag):
    """Column-wise majority over the reference's alphabet order, gap included, first maximum wins
    (PoolRepresentativeWorker.py:30-85).  mat: uint8 [k, L] -> uint8 [L]."""
    alphabet = np.frombuffer((b'ACDEFGHIKLMNPQRSTVWY-' if prot_flag else b'ACGT-'), dtype=np.uint8)
    counts = np.stack([(mat == a).sum(axis=0) for a in alphabet])
    return alphabet[np.argmax(counts, axis=0)]


class ReducedReference:
    def __init__(self, ref_fp, prot_flag, tree_file, threshold, num_thread=0, cluster_tsv=None, tree=None,
                 refs=None):
        """Same positional arguments as Reference.py:65.  Extra keyword inputs:
        cluster_tsv -- a TreeCluster-format TSV to use instead of running a clustering;
        tree        -- an already parsed BackboneTree for `tree_file`;
        refs        -- an already loaded {name: 'S1' row} dict (skips reading ref_fp).
        """
        self._fm = None
        if refs is None:
            # native reader (hostio.cpp): one byte matrix for the whole alignment, `refs` holds views into it
            self._fm = _fasta.read_alignment(ref_fp, prot_flag, False, pinned=False)
            refs = self._fm.as_dict(copy=False)
        self.refs = refs
        self.prot_flag = bool(prot_flag)
        self.threshold = threshold
        self.baseobs = None
        start = time.time()
        clusters = None
        if cluster_tsv is None and shutil.which('TreeCluster.py') and tree_file and os.path.isfile(str(tree_file)):
            # the reference's own route (Reference.py:85-92)
            out = tempfile.NamedTemporaryFile(delete=False, mode='w+t').name
            try:
                with open(os.devnull, 'w') as nul:
                    subprocess.call(['TreeCluster.py', '-t', str(threshold * 1.2), '-i', tree_file, '-m', 'max',
                                     '-o', out], stdout=nul, stderr=nul)
                if os.path.getsize(out) > 0:
                    clusters = _tc.read_cluster_tsv(out)
            finally:
                os.unlink(out)
        if clusters is not None:
            pass
        elif cluster_tsv is not None:
            clusters = _tc.read_cluster_tsv(cluster_tsv)
        else:
            logging.warning('TreeCluster.py is not on PATH (or produced no output) and no --clusters file was given: '
                            'using the built-in max-diameter clustering, whose parity with TreeCluster is not pinned. '
                            'Representatives and observed sets can differ from an upstream run.')
            if tree is None:
                from .tree import BackboneTree
                tree = BackboneTree.from_newick(tree_file)
            cl = _tc.max_diameter_clusters(tree, threshold * 1.2)
            tmp = tempfile.NamedTemporaryFile(delete=False, mode='w+t', suffix='.tsv').name
            _tc.write_cluster_tsv(tree, cl, tmp)
            clusters = _tc.read_cluster_tsv(tmp)
            os.unlink(tmp)
        logging.info('[%s] Clustering is completed in %.3f seconds.' % (time.strftime('%H:%M:%S'), time.time() - start))
        # member lists in `representatives` order (Reference.py:101-107): singletons ('-1') are their own cluster
        self.groups = []
        for key, group in clusters:
            group = [g for g in group if g in self.refs]
            if not group:
                continue
            if key == '-1':
                self.groups.extend([g] for g in group)
            else:
                self.groups.append(group)
        self._representatives = None

    @property
    def representatives(self):
        """[(consensus 'S1' row, [member names])] like the reference's attribute.  Computed on the host on first use
        (tests, pickled databases); the placement path computes the consensus rows on the device instead
        (apples_set_reference_bytes)."""
        if self._representatives is None:
            reps = []
            for group in self.groups:
                if len(group) == 1:
                    reps.append((self.refs[group[0]], group))
                else:
                    mat = np.vstack([self.refs[g].view(np.uint8) for g in group])
                    reps.append((consensus_rows(mat, self.prot_flag).view('S1'), group))
            self._representatives = reps
        return self._representatives

    def set_baseobs(self, baseobs):
        self.baseobs = baseobs

    def __getstate__(self):
        # the pickled database owns its sequences (the reader's buffer is not part of it)
        d = dict(self.__dict__)
        if d.get('_fm') is not None:
            d['refs'] = {k: v.copy() for k, v in self.refs.items()}
            d['_fm'] = None
        return d

    def _byte_matrix(self, names, L):
        fm = getattr(self, '_fm', None)
        if fm is not None and fm.uniform and fm.n == len(names) and fm.names == names:
            return fm.matrix   # the reader's own matrix (rows padded to 16 bytes): the byte entry points take a row stride
        return _fasta.as_byte_matrix([self.refs[n] for n in names], L)

    def get_obs_dist(self, query_seq, query_tag, overlap_frac):
        raise RuntimeError('ReducedReference.get_obs_dist is computed on the GPU by apples_b200.placer.place_batch; '
                           'there is no CPU path in this package')

    # ---------------------------------------------------------------- device layout
    def device_arrays_bytes(self, name_to_node):
        """Inputs of apples_set_reference_bytes: raw alignment bytes + cluster CSR (packing and consensus on the device)."""
        names = list(self.refs.keys())
        row_of = {n: i for i, n in enumerate(names)}
        L = len(self.refs[names[0]]) if names else 0
        offs = np.zeros(len(self.groups) + 1, dtype=np.int32)
        members = []
        for i, group in enumerate(self.groups):
            members.extend(row_of[g] for g in group)
            offs[i + 1] = len(members)
        return dict(kind=_fasta.AA if self.prot_flag else _fasta.NUC, L=L, ref_names=names,
                    ref_bytes=self._byte_matrix(names, L),
                    ref_node=np.array([name_to_node.get(n, -1) for n in names], dtype=np.int32),
                    group_offsets=offs, group_members=np.asarray(members, dtype=np.int32))

    def device_arrays(self, name_to_node):
        """Packed arrays for apples_set_reference (include/apples_b200.h).

        Returns dict(kind, L, ref_names, packed_refs, ref_node, packed_reps, group_offsets, group_members).
        Reference rows follow `self.refs` insertion order; references whose name is not a leaf of the tree get
        ref_node = -1 (the reference silently skips them in Subtree.validate_edges, Subtree.py:31).
        """
        names = list(self.refs.keys())
        row_of = {n: i for i, n in enumerate(names)}
        kind = _fasta.AA if self.prot_flag else _fasta.NUC
        L = len(self.refs[names[0]]) if names else 0
        mat = self._byte_matrix(names, L)
        rep_mat = _fasta.as_byte_matrix([r[0] for r in self.representatives], L)
        offs = np.zeros(len(self.representatives) + 1, dtype=np.int32)
        members = []
        for i, (_, group) in enumerate(self.representatives):
            members.extend(row_of[g] for g in group)
            offs[i + 1] = len(members)
        return dict(
            kind=kind, L=L, ref_names=names,
            packed_refs=_fasta.pack(mat, kind),
            ref_node=np.array([name_to_node.get(n, -1) for n in names], dtype=np.int32),
            packed_reps=_fasta.pack(rep_mat, kind),
            group_offsets=offs,
            group_members=np.asarray(members, dtype=np.int32),
        )
