"""GPU (-m gpu): the CUDA hot path, called through the C ABI, against the golden vectors of the unmodified
reference and against the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): mismatch / overlap counts bit-exact; distances, branch lengths and LS scores
within 1e-9 relative in fp64 (scores with a tiny absolute floor because exact fits score ~1e-16 in the reference);
placement edges identical.  With identical input distances (distance-matrix mode) x_1, x_2, distal and pendant are
required to be BIT-IDENTICAL; the score may differ in the last bits only because the reference squares with
pow(x, 2) (DESIGN.md "parity").
"""
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

REL = 1e-9
ERR_FLOOR = 1e-12

ALL_CASES = util.golden_names()


def _placer(ci, tree, ref):
    from apples_b200.placer import GpuPlacer
    return GpuPlacer(tree, ref, tree.name_to_node, device=0)


def _check_p(case, qname, got, exp, exact_lengths, octx=None, query=None):
    assert [isinstance(x, int) for x in got] == [isinstance(x, int) for x in exp], (case, qname, got, exp)
    if got[0] != exp[0]:
        # documented exception: exact-score ties (BASELINE.json north_star).  Accept only if the oracle scores the two
        # edges within tolerance of each other.
        assert octx is not None and query is not None, (case, qname, got, exp)
        det = {}
        octx.runquery(*query, detail=det)
        ea, eb = det['edges'][got[0]][2], det['edges'][exp[0]][2]
        assert util.close(ea, eb, REL, ERR_FLOOR), ('edge differs and is not a score tie', case, qname, got, exp)
        return 'tie'
    assert util.close(got[1], exp[1], REL, ERR_FLOOR), (case, qname, got, exp)
    if exact_lengths:
        assert got[3] == exp[3] and got[4] == exp[4], (case, qname, got, exp)
    else:
        assert util.close(got[3], exp[3], REL, 1e-15) and util.close(got[4], exp[4], REL, 1e-15), (case, qname, got, exp)
    return 'ok'


@pytest.mark.parametrize('case', ALL_CASES)
def test_golden_placements(case, workdir):
    """place_batch (the drop-in for pool.starmap(runquery)) reproduces the reference's per-query results."""
    from apples_b200.placer import place_batch
    ci = util.CaseInputs(case, workdir)
    tree, ref = ci.product_state()
    res = place_batch(ref, ci.options, tree.name_to_node, ci.queries, tree=tree, device=0)
    assert len(res) == len(ci.queries)
    matrix = ci.refs is None
    octx = None
    ties = 0
    for query, rec, r in zip(ci.queries, ci.g['queries'], res):
        got = r['placements'][0]['p'][0]
        exp = [util.unhex(x) for x in rec['p']]
        assert r['placements'][0]['n'] == [rec['out_name']]
        if got[0] != exp[0] and octx is None:
            octx = ci.oracle_context()
        if _check_p(case, query[0], got, exp, matrix, octx, (query[0], query[1], dict(query[2]) if query[2] else None)) == 'tie':
            ties += 1
    assert ties <= max(1, len(res) // 10), 'too many tie exceptions: %d' % ties
    from apples_b200 import jplace
    import json
    joined = jplace.join_jplace(json.loads(json.dumps(res)))
    assert [p['n'][0] for p in joined['placements']] == ci.g['joined_names']


def test_counts_bit_exact_and_distances(workdir):
    """kernel (a): mismatch / valid counts of all 10 x 490 pairs of config 1 are bit-exact; jc69 within 1e-9 of the
    oracle and within 5e-9 (8-decimal rounding) of the reference's data/dist.mat."""
    from oracle import apples_oracle as orc
    ci = util.CaseInputs('c1_align_FM_MLSE', workdir)
    tree, ref = ci.product_state()
    pl = _placer(ci, tree, ref)
    packed = pl.pack_queries([q[1] for q in ci.queries])
    mism, valid, dist = pl.distance_counts(packed, 0.001)
    assert pl.ref_names == ci.g['ref_names']
    exp = np.array(ci.g['counts'], dtype=np.int64)
    assert (mism == exp[:, :, 0]).all() and (valid == exp[:, :, 1]).all()
    rows = util.read_dismat(util.gunzip_to('dist.mat', workdir))
    worst = 0.0
    for qi, (qname, qseq, _) in enumerate(ci.queries):
        for ri, rn in enumerate(pl.ref_names):
            d = orc.jc69(qseq, ci.refs[rn], 0.001)
            assert util.close(dist[qi, ri], d, REL, 0.0), (qname, rn, dist[qi, ri], d)
            worst = max(worst, abs(dist[qi, ri] - rows[qi][2][rn]))
    assert worst <= 5.1e-9
    # overlap gate (distance.py:735): with a 0.9 overlap requirement many pairs become missing data (-1.0)
    _, _, dist2 = pl.distance_counts(packed, 0.9)
    for qi, (qname, qseq, _) in enumerate(ci.queries[:3]):
        for ri, rn in enumerate(pl.ref_names):
            d = orc.jc69(qseq, ci.refs[rn], 0.9)
            assert util.close(dist2[qi, ri], d, REL, 0.0)
    pl.close()


def test_scoredist(workdir):
    """kernel (a), amino acids: scoredist with FastTree's BLOSUM45 within 1e-9 relative (BLAS summation order)."""
    from oracle import apples_oracle as orc
    ci = util.CaseInputs('c3_prot_FM_MLSE', workdir)
    tree, ref = ci.product_state()
    pl = _placer(ci, tree, ref)
    packed = pl.pack_queries([q[1] for q in ci.queries])
    _, valid, dist = pl.distance_counts(packed, 0.001)
    sd = ci.g['scoredist_q0']
    idx = {n: i for i, n in enumerate(pl.ref_names)}
    for n, d in zip(sd['refs'], sd['d']):
        assert util.close(dist[0, idx[n]], util.unhex(d), REL, 0.0), (n, dist[0, idx[n]], util.unhex(d))
    rng = np.random.default_rng(1)
    for qi in range(len(ci.queries)):
        for ri in rng.integers(0, len(pl.ref_names), 40):
            a, b = ci.queries[qi][1], ci.refs[pl.ref_names[ri]]
            assert util.close(dist[qi, ri], orc.scoredist(a, b, 0.001), REL, 0.0)
            assert valid[qi, ri] == np.count_nonzero((a != b'-') & (b != b'-'))
    pl.close()


@pytest.mark.parametrize('case', ['c1_align_FM_MLSE', 'c1_align_FM_MLSE_f045_b5', 'c1_align_special', 'c2_matrix_FM_MLSE',
                                  'c2_matrix_BME_ME_b5', 'c2_matrix_OLS_HYBRID_f100', 'syn300_FM_MLSE_pos',
                                  'c3_prot_FM_MLSE', 'small_FM_MLSE_pos'])
def test_observed_sets(case, workdir):
    """kernel (c): the observed set of every query equals the reference's obs_dist (after the own-entry removal of
    PoolQueryWorker.py:63-66): same leaves, distances exact (matrix) / 1e-9 (alignment)."""
    ci = util.CaseInputs(case, workdir)
    tree, ref = ci.product_state()
    pl = _placer(ci, tree, ref)
    params = pl.params_from_options(ci.options)
    names = [q[0] for q in ci.queries]
    self_node = pl.self_nodes(names)
    if ci.refs is None:
        tags = list(ci.queries[0][2].keys())
        pl.set_matrix_tags(tags)
        rows = np.array([[q[2][t] for t in tags] for q in ci.queries], dtype=np.float64)
        count, node, dist = pl.observed_sets(params, rows=rows, self_node=self_node, cap=2048)
    else:
        packed = pl.pack_queries([q[1] for q in ci.queries])
        count, node, dist = pl.observed_sets(params, packed=packed, self_node=self_node, cap=2048)
    for qi, rec in enumerate(ci.g['queries']):
        exp = {tree.name_to_node[k]: util.unhex(v) for k, v in rec['observed'] if k != rec['name']}
        k = int(count[qi])
        assert k == len(exp), (case, rec['name'], k, len(exp))
        if rec['status'] in (1, 2):
            continue  # zero-distance / too-few: the set is not sorted (no placement follows)
        got_nodes = node[qi, :k].tolist()
        assert got_nodes == sorted(exp.keys()), (case, rec['name'])
        for u, d in zip(got_nodes, dist[qi, :k].tolist()):
            if ci.refs is None:
                assert d == exp[u]
            else:
                assert util.close(d, exp[u], REL, 0.0)
    pl.close()


@pytest.mark.parametrize('case', [c for c in ALL_CASES if c.startswith(('small_', 'c2_', 'syn300_', 'c1_align_FM', 'c3_prot_FM'))])
def test_edge_solutions(case, workdir):
    """kernel (b): x_1, x_2 and the LS error on EVERY edge of the restricted subtree, and the valid-node set."""
    ci = util.CaseInputs(case, workdir)
    tree, ref = ci.product_state()
    pl = _placer(ci, tree, ref)
    params = pl.params_from_options(ci.options)
    matrix = ci.refs is None
    if matrix:
        tags = list(ci.queries[0][2].keys())
        pl.set_matrix_tags(tags)
    n = 0
    for (qname, qseq, row), rec in zip(ci.queries, ci.g['queries']):
        if 'edges' not in rec:
            continue
        n += 1
        sn = tree.name_to_node.get(qname, -1)
        if matrix:
            x1, x2, err, valid = pl.edge_solutions(params, row=np.array([[row[t] for t in tags]], dtype=np.float64),
                                                   self_node=sn)
        else:
            x1, x2, err, valid = pl.edge_solutions(params, packed_row=pl.pack_queries([qseq]), self_node=sn)
        exp = {int(k): [util.unhex(x) for x in v] for k, v in rec['edges'].items()}
        assert sorted(np.nonzero(valid)[0].tolist()) == sorted(exp.keys()), (case, qname)
        assert int(valid.sum()) == rec['num_nodes']
        for u, (e1, e2, ee) in exp.items():
            if matrix:
                assert x1[u] == e1 and x2[u] == e2, (case, qname, u, x1[u], e1, x2[u], e2)
            else:
                scale = max(abs(e1), abs(e2), tree.edge_length[u], 1e-3)
                assert abs(x1[u] - e1) <= 1e-9 * scale and abs(x2[u] - e2) <= 1e-9 * scale, (case, qname, u)
            assert util.close(err[u], ee, REL, ERR_FLOOR), (case, qname, u, err[u], ee)
    assert n > 0
    pl.close()


def _synthetic(n_leaves, L, n_queries, seed, workdir, protein=False, **tree_kw):
    from apples_b200 import synth
    from apples_b200.reference import ReducedReference
    from apples_b200.tree import BackboneTree
    nwk = synth.random_tree(n_leaves, seed=seed, **tree_kw)
    tree = BackboneTree.from_newick(nwk)
    refs, states = synth.evolve_alignment(tree, L, seed=seed + 1, protein=protein)
    queries, src = synth.make_queries(tree, states, n_queries, seed=seed + 2, protein=protein)
    ref = ReducedReference(None, protein, None, 0.2, 1, tree=tree, refs=refs)
    return nwk, tree, refs, ref, queries, src


@pytest.mark.parametrize('method,criterion', [('OLS', 'MLSE'), ('BME', 'MLSE'), ('FM', 'HYBRID'), ('BE', 'ME')])
def test_synthetic_against_oracle(method, criterion, workdir):
    """config-4-like shapes at a size the oracle finishes in seconds: 2000-leaf backbone, 1500 sites."""
    import os
    import types
    from oracle import apples_oracle as orc
    from apples_b200 import treecluster
    from apples_b200.placer import place_batch
    nwk, tree, refs, ref, queries, _ = _synthetic(2000, 1500, 48, 100, workdir)
    ref.set_baseobs(25)
    opt = types.SimpleNamespace(method_name=method, criterion_name=criterion, negative_branch=False,
                                base_observation_threshold=25, filt_threshold=0.2, minimum_alignment_overlap=0.001,
                                exclude_intplace=False)
    qlist = [(k, v, None) for k, v in queries.items()]
    res = place_batch(ref, opt, tree.name_to_node, qlist, tree=tree, device=0)
    tfp = os.path.join(workdir, 'syn2000.nwk')
    open(tfp, 'w').write(nwk)
    otree, onames = orc.load_tree(tfp)
    octx = orc.OracleContext(otree, onames, refs=refs, representatives=ref.representatives, method=method,
                             criterion=criterion)
    ties = 0
    for q, r in zip(qlist, res):
        exp, _ = octx.runquery(q[0], q[1], None)
        if _check_p('syn2000', q[0], r['placements'][0]['p'][0], exp['placements'][0]['p'][0], False, octx, q) == 'tie':
            ties += 1
    assert ties <= 2


def test_properties_at_scale(workdir):
    """Size-independent properties on a 20 000-leaf backbone with 5000 sites and 6000 queries (several sub-batches
    of the dense kernel): determinism, batch-composition independence, resident == host path, and copies of backbone
    leaves land on a zero-distance leaf whose sequence really is at distance 0."""
    from apples_b200 import _lib
    from apples_b200.placer import GpuPlacer
    nwk, tree, refs, ref, queries, src = _synthetic(20000, 5000, 6000, 200, workdir)
    pl = GpuPlacer(tree, ref, tree.name_to_node, device=0)
    pl.set_limits(max_subbatch=2048, scratch_bytes=32 << 20, slot_cap=32)  # force sub-batches, chunks and overflow reruns
    params = _lib.make_params('FM', 'MLSE')
    names = list(queries.keys())
    packed = pl.pack_queries([queries[n] for n in names])
    # exact copies of 50 leaves, appended
    leaf_names = [tree.label[u] for u in tree.leaf_ids[:50]]
    packed = np.concatenate([packed, pl.pack_queries([refs[n] for n in leaf_names])])
    self_node = np.full(packed.shape[0], -1, np.int32)
    a = pl.place_packed(packed, self_node, params)
    b = pl.place_packed(packed, self_node, params)
    for x, y in zip(a, b):
        assert (x == y).all()
    pl.set_limits(max_subbatch=32768, scratch_bytes=6 << 30, slot_cap=256)
    b2 = pl.place_packed(packed, self_node, params)
    for x, y in zip(a, b2):
        assert (x == y).all()
    sub = pl.place_packed(packed[1000:1100], self_node[1000:1100], params)
    for x, y in zip(a, sub):
        assert (x[1000:1100] == y).all()
    pl.upload_queries(packed)
    pl.place_resident(params)
    c = pl.download_results()
    for x, y in zip(a, c):
        assert (x == y).all()
    edge, error, distal, pendant, status = a
    code = status & 0xff
    assert (code[-50:] == _lib.ZERO_DIST_LEAF).all()
    mism, valid, dist = pl.distance_counts(packed[-50:], 0.001)
    row_of = {n: i for i, n in enumerate(pl.ref_names)}
    node_to_name = {tree.name_to_node[n]: n for n in pl.ref_names}
    for i in range(50):
        hit = node_to_name[int(edge[-50 + i])]
        assert mism[i, row_of[hit]] == 0 and dist[i, row_of[hit]] == 0.0
    placed = (code == _lib.PLACED) | (code == _lib.PLACED_MISPLACEMENT_FLAG)
    assert placed[:-50].mean() > 0.95
    el = tree.edge_length[edge[placed]]
    assert (distal[placed] >= -1e-12).all() and (distal[placed] <= np.maximum(el, 0) + 1e-12).all()
    assert (pendant[placed] >= 0).all()
    # the true source leaf of each query is usually inside the clade the query is placed into / next to
    t = pl.timings()
    assert t['rep_distance_launches'] >= 2
    pl.close()


def test_error_behaviour_and_ragged_inputs(workdir):
    """C-ABI error returns (no exceptions cross the boundary, no fallback) and ragged / degenerate inputs."""
    import ctypes as C
    from apples_b200 import _lib
    from apples_b200.placer import GpuPlacer
    from apples_b200.tree import BackboneTree
    lib = _lib.load()
    ci = util.CaseInputs('c1_align_FM_MLSE', workdir)
    tree, ref = ci.product_state()
    params = _lib.make_params()
    # placing before a reference is set is an error with a message
    pl = GpuPlacer(tree, None, tree.name_to_node, device=0)
    out = pl._outputs(1)
    dummy = np.zeros((1, 3, 52), np.uint32)
    rc = lib.apples_place_batch(pl.h, 1, _lib.ptr(dummy), None, C.byref(params), *[_lib.ptr(o) for o in out])
    assert rc != 0 and b'apples_set_reference' in lib.apples_last_error(pl.h)
    # unknown method / criterion
    bad = _lib.Params(9, 0, 0, 25, 0.2, 0.001)
    pl.set_reference(ref)
    rc = lib.apples_place_batch(pl.h, 1, _lib.ptr(dummy), None, C.byref(bad), *[_lib.ptr(o) for o in out])
    assert rc != 0 and b'method' in lib.apples_last_error(pl.h)
    # zero queries is fine and touches nothing
    assert lib.apples_place_batch(pl.h, 0, None, None, C.byref(params), None, None, None, None, None) == 0
    # a tree whose ids are not post-order ranks is rejected
    par = tree.parent.copy()
    par[0], par[1] = par[1], par[0] if par[0] != par[1] else 0
    par[3] = 1
    assert lib.apples_set_tree(pl.h, tree.num_nodes, _lib.ptr(par), _lib.ptr(tree.edge_length), _lib.ptr(tree.level),
                               _lib.ptr(tree.first)) != 0
    # ragged query alignment / wrong width / bytes outside the alphabet are host-side errors
    seqs = [q[1] for q in ci.queries]
    with pytest.raises(ValueError):
        pl.pack_queries([seqs[0], seqs[1][:-3]])
    with pytest.raises(ValueError):
        pl.pack_queries([s[:100] for s in seqs])
    bad_seq = seqs[0].copy()
    bad_seq[5] = b'.'
    with pytest.raises(ValueError):
        pl.pack_queries([bad_seq])
    # 1, 63, 64 and 65 queries (tile edges of the dense kernel) give the same answers as the full batch
    full = pl.place_packed(pl.pack_queries(seqs * 7), None, params)
    for n in (1, 63, 64, 65):
        part = pl.place_packed(pl.pack_queries((seqs * 7)[:n]), None, params)
        for x, y in zip(full, part):
            assert (x[:n] == y).all()
    pl.close()


def test_tiny_tree_and_all_singletons(workdir):
    """4-leaf backbone, every reference its own cluster (fewer representatives than one dense tile), -b larger than
    the reference: the selection takes everything, like the reference does."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200.placer import place_batch
    from apples_b200.reference import ReducedReference
    from apples_b200.tree import BackboneTree
    nwk = '((A:0.05,B:0.07):0.02,(C:0.04,D:0.03):0.05,E:0.1);'
    tfp = os.path.join(workdir, 'tiny.nwk')
    open(tfp, 'w').write(nwk)
    tree = BackboneTree.from_newick(tfp)
    rng = np.random.default_rng(7)
    base = rng.integers(0, 4, 300)
    alpha = np.frombuffer(b'ACGT', dtype=np.uint8)
    refs, qs = {}, []
    for n in 'ABCDE':
        s = base.copy()
        hit = rng.random(300) < 0.08
        s[hit] = (s[hit] + 1) % 4
        refs[n] = alpha[s].view('S1')
    for i in range(5):
        s = base.copy()
        hit = rng.random(300) < 0.1
        s[hit] = (s[hit] + 2) % 4
        qs.append(('q%d' % i, alpha[s].view('S1'), None))
    tsv = os.path.join(workdir, 'tiny.tsv')
    with open(tsv, 'w') as f:
        f.write('SequenceName\tClusterNumber\n' + ''.join('%s\t-1\n' % n for n in 'ABCDE'))
    # the reference object carries the expansion threshold (Reference.py:146), as upstream builds it from the run's -f
    ref = ReducedReference(None, False, tfp, 0.01, 1, cluster_tsv=tsv, tree=tree, refs=refs)
    for method, crit, b in [('FM', 'MLSE', 25), ('OLS', 'ME', 2), ('BME', 'HYBRID', 100)]:
        opt = types.SimpleNamespace(method_name=method, criterion_name=crit, negative_branch=False,
                                    base_observation_threshold=b, filt_threshold=0.01, minimum_alignment_overlap=0.001,
                                    exclude_intplace=False)
        res = place_batch(ref, opt, tree.name_to_node, qs, tree=tree, device=0)
        otree, onames = orc.load_tree(tfp)
        octx = orc.OracleContext(otree, onames, refs=refs, representatives=orc.representatives_from_tsv(tsv, refs, False),
                                 method=method, criterion=crit, filt_threshold=0.01, baseobs=b)
        for q, r in zip(qs, res):
            exp, _ = octx.runquery(q[0], q[1], None)
            _check_p('tiny', q[0], r['placements'][0]['p'][0], exp['placements'][0]['p'][0], False, octx, q)


@pytest.mark.parametrize('case', ['c1_align_FM_MLSE', 'syn300_OLS_MLSE_pos', 'c3_prot_FM_MLSE'])
def test_device_packer_and_consensus(case, workdir):
    """packing and consensus representatives on the device give the same placements, observed sets
    and counts as the host numpy packer / consensus (which the CPU suite checks against the reference's own)."""
    from apples_b200 import fasta
    from apples_b200.placer import GpuPlacer
    ci = util.CaseInputs(case, workdir)
    tree, ref = ci.product_state()
    params = GpuPlacer.params_from_options(ci.options)
    a = GpuPlacer(tree, None, tree.name_to_node, device=0)
    a.set_reference(ref, on_device=False)
    b = GpuPlacer(tree, None, tree.name_to_node, device=0)
    b.set_reference(ref, on_device=True)
    seqs = [q[1] for q in ci.queries]
    names = [q[0] for q in ci.queries]
    sn = a.self_nodes(names)
    ra = a.place_packed(a.pack_queries(seqs), sn, params)
    rb = b.place_bytes(fasta.as_byte_matrix(seqs, b.L), sn, params)
    for x, y in zip(ra, rb):
        assert (x == y).all()
    ca = a.distance_counts(a.pack_queries(seqs[:4]))
    cb = b.distance_counts(b.pack_queries(seqs[:4]))
    for x, y in zip(ca, cb):
        assert (x == y).all()
    oa = a.observed_sets(params, packed=a.pack_queries(seqs), self_node=sn, cap=2048)
    ob = b.observed_sets(params, packed=b.pack_queries(seqs), self_node=sn, cap=2048)
    for x, y in zip(oa, ob):
        assert (x == y).all()
    a.close()
    b.close()


@pytest.mark.parametrize('method', ['OLS', 'BME'])
def test_config4_shapes(method, workdir):
    """BASELINE.json config 4 at full size: 10 000-leaf backbone, 1500 sites, 100 000 queries, OLS and BME + MLSE.
    Full-size checks are the size-independent ones (determinism, every query accounted for, branch lengths inside the
    edge) plus oracle parity on a sample of the same batch."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200 import _lib, fasta
    from apples_b200.placer import GpuPlacer
    nwk, tree, refs, ref, queries, _ = _synthetic(10000, 1500, 100000, 300, workdir)
    pl = GpuPlacer(tree, ref, tree.name_to_node, device=0)
    params = _lib.make_params(method, 'MLSE')
    names = list(queries.keys())
    mat = fasta.as_byte_matrix([queries[n] for n in names], 1500)
    a = pl.place_bytes(mat, None, params)
    b = pl.place_bytes(mat, None, params)
    for x, y in zip(a, b):
        assert (x == y).all()
    edge, error, distal, pendant, status = a
    code = status & 0xff
    assert set(np.unique(code).tolist()) <= {0, 1, 2, 3}
    placed = (code == 0) | (code == 3)
    assert placed.mean() > 0.95
    el = tree.edge_length[edge[placed]]
    assert (distal[placed] >= -1e-12).all() and (distal[placed] <= el + 1e-12).all() and (pendant[placed] >= 0).all()
    tfp = os.path.join(workdir, 'c4.nwk')
    open(tfp, 'w').write(nwk)
    otree, onames = orc.load_tree(tfp)
    octx = orc.OracleContext(otree, onames, refs=refs, representatives=ref.representatives, method=method, criterion='MLSE')
    ties = 0
    for i in range(0, 100000, 2500):
        q = (names[i], queries[names[i]], None)
        exp, st = octx.runquery(*q)
        p = exp['placements'][0]['p'][0]
        if st in (0, 3):
            got = [int(edge[i]), float(error[i]), 1, float(distal[i]), 0 if status[i] & 0x100 else float(pendant[i])]
            if _check_p('c4', names[i], got, p, False, octx, q) == 'tie':
                ties += 1
        else:
            assert code[i] == st and (st != 1 or edge[i] == p[0])
    assert ties <= 1
    pl.close()


@pytest.mark.parametrize('method,criterion', [('FM', 'MLSE'), ('BME', 'HYBRID'), ('OLS', 'ME')])
def test_deep_caterpillar_all_leaves_observed(method, criterion, workdir):
    """A 1500-leaf caterpillar (1499 levels) placed from a distance matrix with the filter off: every leaf is observed
    (slot overflow reruns up to the full capacity), the restricted subtree is the whole tree and spans more levels
    than the shared-memory bucket table of the placement kernel (level-sweep fallback)."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200.placer import place_batch
    from apples_b200.tree import BackboneTree
    rng = np.random.default_rng(77)
    n = 1500
    s = 'L0:%.5f' % rng.uniform(0.001, 0.02)
    for i in range(1, n):
        s = '(%s,L%d:%.5f):%.5f' % (s, i, rng.uniform(0.001, 0.02), rng.uniform(0.0005, 0.01))
    s = s.rsplit(':', 1)[0] + ';'
    tfp = os.path.join(workdir, 'cat.nwk')
    open(tfp, 'w').write(s)
    tree = BackboneTree.from_newick(tfp)
    assert tree.level.max() >= 1400
    # path distances from 4 attachment points (sister to L100, L700, L1200, L1499) plus noise
    depth = np.zeros(tree.num_nodes)
    for u in range(tree.num_nodes - 2, -1, -1):
        depth[u] = depth[tree.parent[u]] + tree.edge_length[u]
    names = ['L%d' % i for i in range(n)]
    leaf = np.array([tree.name_to_node[x] for x in names])

    def pathdist(a, b):
        x, y = a, b
        while x != y:
            if x < y:
                x = tree.parent[x]
            else:
                y = tree.parent[y]
        return depth[a] + depth[b] - 2 * depth[x]
    queries = []
    for qi, anchor in enumerate([100, 700, 1200, 1499]):
        d = np.array([pathdist(leaf[anchor], leaf[j]) for j in range(n)]) + 0.013
        d = np.round(d * (1 + 0.02 * rng.standard_normal(n)), 6)
        d[d <= 0] = 0.001
        queries.append(('q%d' % qi, None, dict(zip(names, d.tolist()))))
    opt = types.SimpleNamespace(method_name=method, criterion_name=criterion, negative_branch=False,
                                base_observation_threshold=25, filt_threshold=100.0, minimum_alignment_overlap=0.001,
                                exclude_intplace=False)
    res = place_batch(None, opt, tree.name_to_node, queries, tree=tree, device=0)
    otree, onames = orc.load_tree(tfp)
    octx = orc.OracleContext(otree, onames, method=method, criterion=criterion, filt_threshold=100.0)
    for q, r in zip(queries, res):
        exp, _ = octx.runquery(q[0], None, dict(q[2]))
        g, e = r['placements'][0]['p'][0], exp['placements'][0]['p'][0]
        assert g[0] == e[0], (q[0], g, e)
        assert util.close(g[1], e[1], 1e-9, 1e-9) and g[3] == e[3] and g[4] == e[4], (q[0], g, e)


def _random_matrix_case(rng, workdir, tag):
    """random backbone (polytomies, zero / negative edges) + a random distance-matrix batch with ties, missing
    values (-1), tags that are not in the tree, zero distances and queries named like backbone leaves"""
    from apples_b200 import synth
    from apples_b200.tree import BackboneTree
    n = int(rng.integers(4, 90))
    nwk = synth.random_tree(n, seed=int(rng.integers(1, 1 << 30)), polytomy_frac=float(rng.choice([0.0, 0.2, 0.5])),
                            zero_frac=float(rng.choice([0.0, 0.1])), neg_frac=float(rng.choice([0.0, 0.05])),
                            model=str(rng.choice(['yule', 'uniform'])))
    tfp = os.path.join(workdir, 'rnd_%s.nwk' % tag)
    open(tfp, 'w').write(nwk)
    tree = BackboneTree.from_newick(tfp)
    leaves = [tree.label[u] for u in tree.leaf_ids.tolist()]
    tags = leaves + ['ghost1', 'ghost2']
    rng.shuffle(tags)
    depth = np.zeros(tree.num_nodes)
    for u in range(tree.num_nodes - 2, -1, -1):
        depth[u] = depth[tree.parent[u]] + abs(tree.edge_length[u])
    queries = []
    for qi in range(int(rng.integers(1, 9))):
        anchor = tree.name_to_node[leaves[int(rng.integers(0, n))]]
        row = {}
        for t in tags:
            if t in tree.name_to_node:
                a, b = anchor, tree.name_to_node[t]
                x, y = a, b
                while x != y:
                    if x < y:
                        x = tree.parent[x]
                    else:
                        y = tree.parent[y]
                d = depth[a] + depth[b] - 2 * depth[x] + 0.01
                d = round(float(d * (1 + 0.1 * rng.standard_normal())), int(rng.choice([2, 3, 6])))  # coarse rounding -> ties
                d = max(d, 0.001)
            else:
                d = float(rng.uniform(0, 0.5))
            u = rng.random()
            if u < 0.05:
                d = -1.0
            elif u < 0.07:
                d = 0.0
            row[t] = d
        name = leaves[int(rng.integers(0, n))] if rng.random() < 0.15 else 'rq%d' % qi
        queries.append((name, None, row))
    return tfp, tree, queries


def test_randomized_differential_matrix_mode(workdir):
    """300 seeded random cases (tree shape, polytomies, zero/negative edges, ties, missing values, zero distances,
    own-name queries, every method x criterion x -n, random -b / -f) against the oracle: edge, status and int-ness
    identical, branch lengths bit-identical, score within 1e-9."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200.placer import place_batch
    rng = np.random.default_rng(20240917)
    n_checked = n_degenerate = 0
    for case in range(300):
        tfp, tree, queries = _random_matrix_case(rng, workdir, case)
        method = str(rng.choice(['FM', 'OLS', 'BME', 'BE']))
        criterion = str(rng.choice(['MLSE', 'ME', 'HYBRID']))
        neg = bool(rng.random() < 0.25)
        b = int(rng.choice([1, 3, 5, 25, 1000]))
        f = float(rng.choice([0.0, 0.05, 0.2, 100.0]))
        opt = types.SimpleNamespace(method_name=method, criterion_name=criterion, negative_branch=neg,
                                    base_observation_threshold=b, filt_threshold=f, minimum_alignment_overlap=0.001,
                                    exclude_intplace=bool(rng.random() < 0.3))
        otree, onames = orc.load_tree(tfp)
        octx = orc.OracleContext(otree, onames, method=method, criterion=criterion, negative_branch=neg,
                                 filt_threshold=f, baseobs=b, exclude_intplace=opt.exclude_intplace)
        # singular 2x2 systems: the reference raises in util.solve2_2 (util.py:26-27) and the run dies; place_batch
        # raises ZeroDivisionError exactly for the batches in which the reference does (APPLES_FLAG_DEGENERATE)
        ref_raises = False
        for q in queries:
            try:
                octx.runquery(q[0], None, dict(q[2]))
            except (ZeroDivisionError, AssertionError):
                ref_raises = True
            except (FloatingPointError, OverflowError):
                pass
        try:
            res = place_batch(None, opt, tree.name_to_node, queries, tree=tree, device=0)
            assert not ref_raises, ('the reference raises on a singular system, place_batch did not', case)
        except ZeroDivisionError:
            assert ref_raises, ('place_batch raised on a system the reference solves', case)
            n_degenerate += 1
            continue
        for q, r in zip(queries, res):
            try:
                exp, st = octx.runquery(q[0], None, dict(q[2]))
            except (FloatingPointError, OverflowError):
                continue  # Python-level overflow in the reference: nothing to compare
            g, e = r['placements'][0]['p'][0], exp['placements'][0]['p'][0]
            ctx = (case, q[0], method, criterion, neg, b, f, g, e)
            assert r['placements'][0]['n'] == exp['placements'][0]['n'], ctx
            if any(isinstance(x, float) and (x != x or abs(x) == float('inf')) for x in e[1:]):
                continue  # degenerate systems (nan / inf) are outside the parity contract
            if g[0] != e[0]:
                det = {}
                octx.runquery(q[0], None, dict(q[2]), detail=det)
                assert g[0] in det['edges'] and util.close(det['edges'][g[0]][2], det['edges'][e[0]][2], 1e-9, 1e-12), ctx
                continue
            assert [isinstance(x, int) for x in g] == [isinstance(x, int) for x in e], ctx
            assert util.close(g[1], e[1], 1e-9, 1e-12) and g[3] == e[3] and g[4] == e[4], ctx
            n_checked += 1
    assert n_checked > 800


def test_randomized_differential_alignment_mode(workdir):
    """120 seeded random alignment cases (nucleotide and protein, heavy random gaps, random clusterings incl. singletons,
    random -V / -b / -f, copies of references, own-name queries) against the oracle: observed sets identical, edges
    identical (score ties excepted), values within 1e-9."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200 import synth
    from apples_b200.placer import GpuPlacer, results_to_jplace
    from apples_b200.reference import ReducedReference
    from apples_b200.tree import BackboneTree
    rng = np.random.default_rng(424242)
    checked = ties = 0
    for case in range(120):
        protein = bool(rng.random() < 0.3)
        n = int(rng.integers(5, 70))
        L = int(rng.integers(40, 700))
        nwk = synth.random_tree(n, seed=int(rng.integers(1, 1 << 30)), mean_edge=float(rng.choice([0.01, 0.03, 0.08])),
                                polytomy_frac=float(rng.choice([0.0, 0.3])))
        tfp = os.path.join(workdir, 'rnda_%d.nwk' % case)
        open(tfp, 'w').write(nwk)
        tree = BackboneTree.from_newick(tfp)
        refs, states = synth.evolve_alignment(tree, L, seed=int(rng.integers(1, 1 << 30)), protein=protein,
                                              gap_frac=float(rng.choice([0.0, 0.05, 0.4])), edge_frac=float(rng.choice([0.0, 0.3])))
        qd, _ = synth.make_queries(tree, states, int(rng.integers(1, 7)), seed=int(rng.integers(1, 1 << 30)), protein=protein,
                                   mean_extra=float(rng.choice([0.0, 0.03, 0.2])), gap_frac=float(rng.choice([0.0, 0.1, 0.6])))
        queries = [(k, v, None) for k, v in qd.items()]
        names = list(refs.keys())
        queries.append(('copy', refs[names[int(rng.integers(0, n))]], None))          # zero-distance shortcut
        own = names[int(rng.integers(0, n))]
        queries.append((own, refs[own], None))                                        # query named like a leaf
        # random clustering: random cluster ids, about a third singletons
        tsv = os.path.join(workdir, 'rnda_%d.tsv' % case)
        with open(tsv, 'w') as f:
            f.write('SequenceName\tClusterNumber\n')
            for nm in names:
                cid = -1 if rng.random() < 0.35 else int(rng.integers(1, max(2, n // 3)))
                f.write('%s\t%d\n' % (nm, cid))
        thr = float(rng.choice([0.05, 0.2, 0.6]))
        ref = ReducedReference(None, protein, tfp, thr, 1, cluster_tsv=tsv, tree=tree, refs=refs)
        opt = types.SimpleNamespace(method_name=str(rng.choice(['FM', 'OLS', 'BME', 'BE'])),
                                    criterion_name=str(rng.choice(['MLSE', 'ME', 'HYBRID'])),
                                    negative_branch=bool(rng.random() < 0.2), base_observation_threshold=int(rng.choice([2, 5, 25])),
                                    filt_threshold=thr, minimum_alignment_overlap=float(rng.choice([0.001, 0.2, 0.5])),
                                    exclude_intplace=False)
        pl = GpuPlacer(tree, ref, tree.name_to_node, device=0)
        params = pl.params_from_options(opt)
        qn = [q[0] for q in queries]
        sn = pl.self_nodes(qn)
        packed = pl.pack_queries([q[1] for q in queries])
        out = pl.place_packed(packed, sn, params)
        count, node, dist = pl.observed_sets(params, packed=packed, self_node=sn, cap=256)
        res = results_to_jplace(qn, [x in tree.name_to_node for x in qn], out, log=False, degenerate='keep')
        pl.close()
        otree, onames = orc.load_tree(tfp)
        octx = orc.OracleContext(otree, onames, refs=refs, representatives=orc.representatives_from_tsv(tsv, refs, protein),
                                 protein=protein, method=opt.method_name, criterion=opt.criterion_name,
                                 negative_branch=opt.negative_branch, filt_threshold=thr, baseobs=opt.base_observation_threshold,
                                 overlap=opt.minimum_alignment_overlap)
        for qi, (q, r) in enumerate(zip(queries, res)):
            det = {}
            ctx = (case, q[0], protein, vars(opt))
            try:
                exp, st = octx.runquery(q[0], q[1], None, detail=det)
            except (ZeroDivisionError, AssertionError):
                assert int(out[4][qi]) & 0x200, ('the reference raises on a singular system, no APPLES_FLAG_DEGENERATE', ctx)
                continue
            except (FloatingPointError, OverflowError):
                continue
            assert not (int(out[4][qi]) & 0x200), ('APPLES_FLAG_DEGENERATE on a system the reference solves', ctx)
            eobs = {tree.name_to_node[k]: v for k, v in det['observed']}
            assert int(count[qi]) == len(eobs), ctx
            if st in (0, 3):
                k = int(count[qi])
                assert node[qi, :k].tolist() == sorted(eobs), ctx
                for u, d in zip(node[qi, :k].tolist(), dist[qi, :k].tolist()):
                    assert util.close(d, eobs[u], 1e-9, 0.0), ctx
            g, e = r['placements'][0]['p'][0], exp['placements'][0]['p'][0]
            assert r['placements'][0]['n'] == exp['placements'][0]['n'], ctx
            if any(isinstance(x, float) and (x != x or abs(x) == float('inf')) for x in e[1:]):
                continue
            if _check_p('rnda', q[0], g, e, False, octx, q) == 'tie':
                ties += 1
            checked += 1
    assert checked > 500 and ties <= 5, (checked, ties)


def test_config5_parity_including_overflow_reruns(workdir):
    """BASELINE.json config 5 at full reference size (the bench workload: 200 000-leaf backbone, 5000 sites, 21 924
    representatives, FM + MLSE, -f 0.2 -b 25): 20 480 queries through apples_place_batch_bytes, then a seeded sample of
    64 ordinary queries plus 16 queries from the OVERFLOW set (observed set larger than the 256-entry slot: stash of the
    key row, gather, rerun selection with the larger slot, rerun placement) against the oracle: status, edge, error,
    distal, pendant, the observed set (leaves identical, distances 1e-9) and the number of valid nodes V
    (Reference.py:143-154, PoolQueryWorker.py:101-120)."""
    import bench
    from oracle import apples_oracle as orc
    from apples_b200 import _lib
    from apples_b200.placer import GpuPlacer
    args = bench.parse([])
    args.queries_per_gpu = 20480
    tree, arrays, packed_q, q_bytes, info, host = bench.build_workload(args, 'cuda:0', 0, True)
    nq = args.queries_per_gpu
    pl = GpuPlacer(tree, None, tree.name_to_node, device=0)
    pl.set_reference_arrays(**arrays)
    params = _lib.make_params('FM', 'MLSE')
    q_host = q_bytes.cpu().numpy()
    packed_host = packed_q.cpu().numpy().view(np.uint32)
    edge, error, distal, pendant, status = pl.place_bytes(q_host, None, params)
    K, V, over = pl.last_counts(nq)
    assert over.sum() >= 16, 'the workload is expected to hold ~1 %% overflow queries, saw %d' % over.sum()
    assert (K[over == 1] > 256).all() and (K[over == 0] <= 256).all()
    rng = np.random.default_rng(5)
    sample = np.concatenate([rng.choice(np.flatnonzero(over == 0), 64, replace=False),
                             rng.choice(np.flatnonzero(over == 1), 16, replace=False)])
    # the sample's observed sets through the same pipeline (slot capacity 256: the 16 are rerun here as well)
    count, node, dist = pl.observed_sets(params, packed=np.ascontiguousarray(packed_host[sample]), cap=4096)
    K2, V2, over2 = pl.last_counts(len(sample))
    assert (over2 == over[sample]).all() and (K2 == K[sample]).all() and (V2 == V[sample]).all() and (count == K2).all()
    octx = bench.cpu_context(args, tree, host)
    queries = [('Q%07d' % i, q_host[i].view('S1'), None) for i in sample.tolist()]
    exp = orc.run_pool_detail(octx, queries, os.cpu_count() or 1)
    ties = 0
    for j, (i, q, (res, st, det)) in enumerate(zip(sample.tolist(), queries, exp)):
        assert (int(status[i]) & 0xff) == st, (i, status[i], st)
        eobs = {tree.name_to_node[k]: v for k, v in det['observed']}
        assert int(K[i]) == len(eobs), (i, K[i], len(eobs))
        p = res['placements'][0]['p'][0]
        if st in (1, 2):
            assert st != 1 or int(edge[i]) == p[0]
            continue
        k = int(count[j])
        assert node[j, :k].tolist() == sorted(eobs), i
        for u, d in zip(node[j, :k].tolist(), dist[j, :k].tolist()):
            assert util.close(d, eobs[u], REL, 0.0), (i, u, d, eobs[u])
        assert int(V[i]) == det['num_nodes'], (i, V[i], det['num_nodes'])
        got = [int(edge[i]), float(error[i]), 1, float(distal[i]), 0 if status[i] & 0x100 else float(pendant[i])]
        if _check_p('c5', q[0], got, p, False, octx, q) == 'tie':
            ties += 1
    assert ties <= 2
    # the same batch in three sub-batches (a short first one: the copy stream runs ahead of the compute stream with the
    # device packer, api.cu run_macro), from bytes and from packed rows: byte-identical to the single-sub-batch result
    pl.set_limits(max_subbatch=10240)
    for other in (pl.place_bytes(q_host, None, params), pl.place_packed(packed_host, None, params)):
        for x, y in zip((edge, error, distal, pendant, status), other):
            assert x.tobytes() == y.tobytes()
    assert pl.timings()['rep_distance_launches'] >= 1 + 3 + 3
    pl.close()


@pytest.mark.parametrize('thr,baseobs', [(0.6, 25), (0.0005, 2500)])
def test_clusters_larger_than_the_member_list(thr, baseobs, workdir):
    """Four clusters of ~700 leaves each plus singletons: the members of one batch of pending clusters do not fit the
    selection kernel's shared-memory member list (352 entries in the first pass, 1888 in the rerun), so the list is
    filled and walked in rounds -- through the near phase (large threshold: everything is near) and through the far
    phase (tiny threshold, -b 2500: the clusters are taken in ascending order until 2500 leaves are observed).  Observed
    sets identical to the oracle's, placements within tolerance."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200 import synth
    from apples_b200.placer import GpuPlacer, results_to_jplace
    from apples_b200.reference import ReducedReference
    from apples_b200.tree import BackboneTree
    rng = np.random.default_rng(77)
    n, L = 3000, 400
    nwk = synth.random_tree(n, seed=5150, mean_edge=0.004)
    tfp = os.path.join(workdir, 'bigclusters.nwk')
    open(tfp, 'w').write(nwk)
    tree = BackboneTree.from_newick(tfp)
    refs, states = synth.evolve_alignment(tree, L, seed=5151, gap_frac=0.05)
    qd, _ = synth.make_queries(tree, states, 6, seed=5152, mean_extra=0.01)
    queries = [(k, v, None) for k, v in qd.items()]
    names = list(refs.keys())
    tsv = os.path.join(workdir, 'bigclusters_%s.tsv' % baseobs)
    with open(tsv, 'w') as f:
        f.write('SequenceName\tClusterNumber\n')
        for nm in names:
            f.write('%s\t%d\n' % (nm, -1 if rng.random() < 0.05 else int(rng.integers(1, 5))))
    ref = ReducedReference(None, False, tfp, thr, 1, cluster_tsv=tsv, tree=tree, refs=refs)
    opt = types.SimpleNamespace(method_name='FM', criterion_name='MLSE', negative_branch=False,
                                base_observation_threshold=baseobs, filt_threshold=thr, minimum_alignment_overlap=0.001,
                                exclude_intplace=False)
    ref.set_baseobs(baseobs)
    pl = GpuPlacer(tree, ref, tree.name_to_node, device=0)
    params = pl.params_from_options(opt)
    qn = [q[0] for q in queries]
    sn = pl.self_nodes(qn)
    packed = pl.pack_queries([q[1] for q in queries])
    out = pl.place_packed(packed, sn, params)
    count, node, dist = pl.observed_sets(params, packed=packed, self_node=sn, cap=4096)
    res = results_to_jplace(qn, [False] * len(qn), out, log=False, degenerate='keep')
    pl.close()
    otree, onames = orc.load_tree(tfp)
    octx = orc.OracleContext(otree, onames, refs=refs, representatives=orc.representatives_from_tsv(tsv, refs, False),
                             method='FM', criterion='MLSE', filt_threshold=thr, baseobs=baseobs, overlap=0.001)
    big = 0
    for qi, (q, r) in enumerate(zip(queries, res)):
        det = {}
        exp, st = octx.runquery(q[0], q[1], None, detail=det)
        eobs = {tree.name_to_node[k]: v for k, v in det['observed']}
        k = int(count[qi])
        assert k == len(eobs), (q[0], k, len(eobs))
        big += k > 1888
        if st in (0, 3):   # placed: sorted by node id, distances corrected (a zero-distance query keeps the raw list)
            assert node[qi, :k].tolist() == sorted(eobs), q[0]
            for u, d in zip(node[qi, :k].tolist(), dist[qi, :k].tolist()):
                assert util.close(d, eobs[u], 1e-9, 0.0), (q[0], u)
        else:
            assert sorted(node[qi, :k].tolist()) == sorted(eobs), q[0]
        _check_p('bigclusters', q[0], r['placements'][0]['p'][0], exp['placements'][0]['p'][0], False, octx, q)
    assert big >= 3, 'the case no longer overflows the rerun member list'


def _oracle_check(tag, tree, tfp, refs, reps, queries, res, opt, protein=False):
    from oracle import apples_oracle as orc
    otree, onames = orc.load_tree(tfp)
    octx = orc.OracleContext(otree, onames, refs=refs, representatives=reps, protein=protein, method=opt.method_name,
                             criterion=opt.criterion_name, negative_branch=opt.negative_branch,
                             filt_threshold=opt.filt_threshold, baseobs=opt.base_observation_threshold,
                             overlap=opt.minimum_alignment_overlap)
    ties = 0
    for q, r in zip(queries, res):
        exp, _ = octx.runquery(q[0], q[1], None)
        assert r['placements'][0]['n'] == exp['placements'][0]['n']
        if _check_p(tag, q[0], r['placements'][0]['p'][0], exp['placements'][0]['p'][0], False, octx, q) == 'tie':
            ties += 1
    return ties


def test_symbols_outside_acgt_count_as_characters(workdir):
    """Bytes other than A,C,G,T,- survive fasta2dic only as non-letters ('.', '*', '?', digits) and are ordinary
    characters for jc69 (distance.py:733-737: a site counts when neither byte is '-', and mismatches when the bytes
    differ).  The 2-bit planes cannot carry them: such QUERIES are recomputed by the byte-compare fallback inside the same
    batch; a REFERENCE that holds them puts the whole context into byte mode.  Both against the oracle."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200.placer import GpuPlacer, place_batch
    from apples_b200.reference import ReducedReference
    from apples_b200 import fasta
    ci = util.CaseInputs('c1_align_FM_MLSE', workdir)
    tree, _ = ci.product_state()
    rng = np.random.default_rng(99)
    names = list(ci.refs.keys())
    # ---- exotic queries, clean reference
    queries = []
    for i, (qn, qs, _) in enumerate(ci.queries):
        s = qs.copy()
        if i % 2 == 0:
            for pos in rng.integers(0, len(s), 1 + 3 * i):
                s[pos] = rng.choice([b'.', b'*', b'?', b'7'])
        queries.append((qn, s, None))
    twin = ci.refs[names[5]].copy()            # identical to a reference except for one exotic site: no zero shortcut
    twin[np.flatnonzero(twin != b'-')[10]] = b'?'
    queries.append(('twin', twin, None))
    ref = ReducedReference(None, False, ci.tree_fp, 0.2, 1, cluster_tsv=ci.tsv, tree=tree, refs=ci.refs)
    ref.set_baseobs(25)
    opt = types.SimpleNamespace(method_name='FM', criterion_name='MLSE', negative_branch=False, base_observation_threshold=25,
                                filt_threshold=0.2, minimum_alignment_overlap=0.001, exclude_intplace=False)
    pl = GpuPlacer(tree, ref, tree.name_to_node, device=0)
    res = place_batch(ref, opt, tree.name_to_node, queries, tree=tree, placer=pl)
    assert pl.timings()['fallback_queries'] == sum(1 for q in queries if not np.isin(q[1], [b'A', b'C', b'G', b'T', b'-']).all())
    reps = orc.representatives_from_tsv(ci.tsv, ci.refs, False)
    assert _oracle_check('exoticq', tree, ci.tree_fp, ci.refs, reps, queries, res, opt) <= 1
    # a small slot capacity sends exotic queries through the overflow escalation of the fallback as well
    pl.set_limits(slot_cap=8)
    res2 = place_batch(ref, opt, tree.name_to_node, queries, tree=tree, placer=pl)
    assert [r['placements'][0]['p'] for r in res2] == [r['placements'][0]['p'] for r in res]
    pl.close()
    # ---- exotic reference rows (members of clusters and singletons alike): the whole context runs in byte mode
    refs2 = {k: v.copy() for k, v in ci.refs.items()}
    for k in rng.choice(names, 40, replace=False):
        row = refs2[k]
        for pos in rng.integers(0, len(row), 5):
            row[pos] = rng.choice([b'?', b'.'])
    ref2 = ReducedReference(None, False, ci.tree_fp, 0.2, 1, cluster_tsv=ci.tsv, tree=tree, refs=refs2)
    ref2.set_baseobs(25)
    res3 = place_batch(ref2, opt, tree.name_to_node, queries, tree=tree, device=0)
    reps2 = orc.representatives_from_tsv(ci.tsv, refs2, False)
    assert _oracle_check('exoticr', tree, ci.tree_fp, refs2, reps2, queries, res3, opt) <= 1
    # counts export in byte mode: the reference's definition on the raw bytes
    pl2 = GpuPlacer(tree, ref2, tree.name_to_node, device=0)
    clean = [q[1] for q in ci.queries[:3]]
    mism, valid, dist = pl2.distance_counts(pl2.pack_queries(clean), 0.001)
    for qi, qs in enumerate(clean):
        for ri in range(0, len(pl2.ref_names), 11):
            m, v = orc.nuc_counts(qs, refs2[pl2.ref_names[ri]])
            assert (mism[qi, ri], valid[qi, ri]) == (m, v)
            assert util.close(dist[qi, ri], orc.jc69(qs, refs2[pl2.ref_names[ri]], 0.001), REL, 0.0)
    pl2.close()


def test_alignment_longer_than_65535_columns(workdir):
    """16-bit packed counts end at 65 535 columns; longer nucleotide alignments take the byte-compare path with 32-bit
    counts (the reference has no limit).  80 001 columns (not a multiple of 4 or 16), 40-leaf backbone, against the oracle,
    through the byte API and through the packed API."""
    import types
    from oracle import apples_oracle as orc
    from apples_b200 import synth, fasta, _lib
    from apples_b200.placer import GpuPlacer, place_batch, results_to_jplace
    from apples_b200.reference import ReducedReference
    from apples_b200.tree import BackboneTree
    L = 80001
    nwk = synth.random_tree(40, seed=5, mean_edge=0.03)
    tfp = os.path.join(workdir, 'long.nwk')
    open(tfp, 'w').write(nwk)
    tree = BackboneTree.from_newick(tfp)
    refs, states = synth.evolve_alignment(tree, L, seed=6, edge_frac=0.0)      # no end gaps: overlaps beyond 65 535 sites
    qd, _ = synth.make_queries(tree, states, 5, seed=7, edge_frac=0.0)
    queries = [(k, v, None) for k, v in qd.items()]
    ref = ReducedReference(None, False, None, 0.2, 1, tree=tree, refs=refs)
    ref.set_baseobs(25)
    opt = types.SimpleNamespace(method_name='OLS', criterion_name='MLSE', negative_branch=False, base_observation_threshold=25,
                                filt_threshold=0.2, minimum_alignment_overlap=0.001, exclude_intplace=False)
    res = place_batch(ref, opt, tree.name_to_node, queries, tree=tree, device=0)
    assert _oracle_check('long', tree, tfp, refs, ref.representatives, queries, res, opt) <= 1
    # packed API (host-packed planes in): same placements
    pl = GpuPlacer(tree, None, tree.name_to_node, device=0)
    pl.set_reference(ref, on_device=False)
    params = pl.params_from_options(opt, ref)
    names = [q[0] for q in queries]
    out = pl.place_packed(pl.pack_queries([q[1] for q in queries]), pl.self_nodes(names), params)
    res2 = results_to_jplace(names, [False] * len(names), out, log=False)
    assert [r['placements'][0]['p'] for r in res2] == [r['placements'][0]['p'] for r in res]
    mism, valid, dist = pl.distance_counts(pl.pack_queries([queries[0][1]]), 0.001)
    m, v = orc.nuc_counts(queries[0][1], refs[pl.ref_names[3]])
    assert (mism[0, 3], valid[0, 3]) == (m, v)
    assert valid.max() > 65535   # counts beyond 16 bits really occur
    pl.close()


def test_tensor_core_experiment_is_bit_identical(workdir):
    """dense_tc.cu (opt-in): the tcgen05 kind::i8 kernel computes the query x representative counts as one int8 dot product
    per pair (simplex embedding of the alphabet, S = 63503 * match - mismatch) and must give the SAME 32-bit keys as the
    integer-pipe kernel -- so every result array of a batch is byte-identical between the two modes."""
    from apples_b200 import _lib, fasta
    from apples_b200.placer import GpuPlacer
    nwk, tree, refs, ref, queries, _ = _synthetic(6000, 2500, 5000, 900, workdir)
    ref.set_baseobs(25)
    names = list(queries.keys())
    mat = fasta.as_byte_matrix([queries[k] for k in names], 2500)
    params = _lib.make_params('FM', 'MLSE')
    out = []
    for mode in (0, 1):
        pl = GpuPlacer(tree, None, tree.name_to_node, device=0)
        pl.set_dense_mode(mode)
        pl.set_reference(ref)
        pl.set_limits(max_subbatch=2048)     # several tensor-core launches incl. a ragged last one
        out.append(pl.place_bytes(mat, None, params))
        t = pl.timings()
        assert t['rep_distance_launches'] >= 3
        pl.close()
    for x, y in zip(*out):
        assert x.tobytes() == y.tobytes()
    assert ((out[0][4] & 0xff) == 0).mean() > 0.9
