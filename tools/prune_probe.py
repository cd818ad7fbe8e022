"""Probe (CPU, numpy/torch): on the benchmark workload, what fraction of (query, representative) pairs survives the
exact prefix bound  m_prefix * 2^16 >= P_hi * (v_prefix + min(nq_rest, nr_rest))  for several prefix lengths, and how
many queries need far representatives (obs_num < baseobs after the near clusters).  Decides the design of the two-pass
dense stage."""
import os, sys, time, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

args = bench.parse(sys.argv[1:])
args.queries_per_gpu = int(os.environ.get('PROBE_Q', '96'))
t0 = time.time()
tree, arrays, packed_q, q_bytes, info, host = bench.build_workload(args, 'cpu', 0, True)
print('workload', info, 'in %.0f s' % (time.time() - t0), flush=True)
reps = host['reps']                     # uint8 [R, L]
q = q_bytes.numpy()
R, L = reps.shape
offs = arrays['group_offsets']
sizes = np.diff(offs)
thr = 0.2
pstar = 0.75 * (1 - math.exp(-4 * thr / 3))
P_hi = math.ceil(pstar * (1 + 1e-9) * 65536) + 1
rv = reps != 45
print('R', R, 'L', L, 'pstar', pstar, flush=True)
fr = [0.25, 0.30, 0.35, 0.40, 0.45, 0.50, 0.60]
surv = {f: [] for f in fr}
need_far = 0
near_cnt = []
ps_all = []
for i in range(q.shape[0]):
    qi = q[i]
    qv = qi != 45
    both = rv & qv[None, :]
    mis = (reps != qi[None, :]) & both
    m = mis.sum(1); v = both.sum(1)
    ok = v > 0
    p = np.where(ok, m / np.maximum(v, 1), 9.0)
    ps_all.append(p)
    near = ok & (m * 65536 < P_hi * v)
    near_cnt.append(int(near.sum()))
    obs = int(sizes[near].sum())
    if obs < 25:
        need_far += 1
    for f in fr:
        n = int(round(f * L / 32)) * 32
        a = (L - n) // 2 // 32 * 32
        sl = slice(a, a + n)
        mp = mis[:, sl].sum(1); vp = both[:, sl].sum(1)
        nq_rest = int(qv.sum() - qv[sl].sum())
        nr_rest = rv.sum(1) - rv[:, sl].sum(1)
        vu = vp + np.minimum(nq_rest, nr_rest)
        pruned = mp * 65536 >= P_hi * vu
        surv[f].append(float((~pruned).mean()))
        assert not (pruned & near).any()
print('queries', q.shape[0], 'need far', need_far, 'mean near clusters', np.mean(near_cnt), 'max', np.max(near_cnt))
ps = np.concatenate(ps_all)
print('p quantiles (0.1,1,5,25,50 %)', np.quantile(ps, [0.001, 0.01, 0.05, 0.25, 0.5]))
for f in fr:
    s = np.array(surv[f])
    print('prefix %.2f: survivors mean %.4f  median %.4f  max %.4f  (per query of %d reps: mean %.0f)' % (f, s.mean(), np.median(s), s.max(), R, s.mean() * R))
