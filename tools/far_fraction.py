"""Tuning probe: how many queries of the benchmark workload need far representatives (distance > threshold) to reach
baseobs, and how many far units they take.  GPU only."""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from apples_b200 import _lib
from apples_b200.placer import GpuPlacer

args = bench.parse(sys.argv[1:])
args.queries_per_gpu = 4096
tree, arrays, packed_q, q_bytes, info, host = bench.build_workload(args, 'cuda:0', 0, False)
pl = GpuPlacer(tree, None, tree.name_to_node, device=0)
pl.set_reference_arrays(**arrays)
params = _lib.make_params(args.method, args.criterion)
pq = packed_q.cpu().numpy().view(np.uint32)
self_node = np.full(pq.shape[0], -1, np.int32)
count, node, dist = pl.observed_sets(params, packed=pq, self_node=self_node, cap=4096)
far = ((dist > 0.2) & (node >= 0)).sum(axis=1)
print('queries', len(count), 'mean observed', count.mean(), 'with far', (far > 0).mean(), 'mean far obs | far', far[far > 0].mean() if (far > 0).any() else 0)
print('far-count histogram', np.bincount(np.minimum(far, 30)))
print('observed histogram (bins of 25)', np.bincount(np.minimum(count // 25, 20)))
