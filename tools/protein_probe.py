"""Quick throughput probe of the amino-acid path (config-3 shapes scaled up): not part of the bench contract."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from apples_b200 import synth, _lib, fasta
from apples_b200.tree import BackboneTree
from apples_b200.reference import ReducedReference
from apples_b200.placer import GpuPlacer

n_leaves, L, nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4000, 1638, int(sys.argv[2]) if len(sys.argv) > 2 else 8000
tree = BackboneTree.from_newick(synth.random_tree(n_leaves, seed=3))
refs, states = synth.evolve_alignment(tree, L, seed=4, protein=True)
q, _ = synth.make_queries(tree, states, 200, seed=5, protein=True)
ref = ReducedReference(None, True, None, 0.6, 1, tree=tree, refs=refs)
pl = GpuPlacer(tree, ref, tree.name_to_node, device=0)
mat = fasta.as_byte_matrix(list(q.values()), L)
mat = np.tile(mat, (nq // 200, 1))
params = _lib.make_params('FM', 'MLSE', filt_threshold=0.6)
pl.place_bytes(mat, None, params)
pl.timings(reset=True)
t0 = time.time()
out = pl.place_bytes(mat, None, params)
dt = time.time() - t0
t = pl.timings()
print('protein: %d queries x %d representatives (%d refs) x %d sites: %.1f ms -> %.0f queries/s; stages %s'
      % (len(mat), len(ref.groups), n_leaves, L, dt * 1e3, len(mat) / dt, {k: round(v, 2) for k, v in t.items() if k.endswith('_ms')}))
print('dense Gcell-sites/s: %.1f' % (len(mat) * len(ref.groups) * L / (t['rep_distance_ms'] * 1e-3) / 1e9))
