"""GPU (-m gpu): the command-line scripts (same flags as the reference's run_apples.py / build_applesdtb.py) end to end:
jplace output equals the reference's for configs 1, 2, 3 and the APPLES-database round trip (README.md:55-75)."""
import json
import os
import pickle

import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _check_jplace(path, golden_case, workdir, tree_file):
    from apples_b200.tree import BackboneTree
    g = util.load_golden(golden_case)
    j = json.load(open(path))
    assert sorted(j.keys()) == ['fields', 'metadata', 'placements', 'tree', 'version']
    assert j['fields'] == ['edge_num', 'likelihood', 'like_weight_ratio', 'distal_length', 'pendant_length']
    assert j['version'] == 3
    assert j['tree'] == BackboneTree.from_newick(util.gunzip_to(tree_file, workdir)).extended_newick()
    assert [p['n'][0] for p in j['placements']] == g['joined_names']
    by_name = {q['out_name']: q for q in g['queries']}
    for pl in j['placements']:
        exp = [util.unhex(x) for x in by_name[pl['n'][0]]['p']]
        got = pl['p'][0]
        assert got[0] == exp[0] and got[2] == 1
        for a, b in zip(got[1:], exp[1:]):
            assert util.close(a, b, 1e-9, 1e-12)
    # the file is what the reference writes: sorted keys, indent 4, trailing newline (run_apples.py:116-117)
    txt = open(path).read()
    assert txt == json.dumps(j, sort_keys=True, indent=4) + '\n'


def test_cli_alignment_and_database(workdir):
    import run_apples
    import build_applesdtb
    ref, qry, tree = (util.gunzip_to(n, workdir) for n in ('ref.fa', 'query.fa', 'backbone.nwk'))
    tsv = os.path.join(util.GOLD, 'c1_clusters.tsv')
    out = os.path.join(workdir, 'c1.jplace')
    run_apples.main(['-s', ref, '-q', qry, '-t', tree, '-D', '--clusters', tsv, '-o', out, '-T', '2'])
    _check_jplace(out, 'c1_align_FM_MLSE', workdir, 'backbone.nwk')
    # -x: extended reference = reference + queries (run_apples.py:82-83)
    ext = os.path.join(workdir, 'extended.fa')
    with open(ext, 'w') as f:
        f.write(open(ref).read())
        f.write(open(qry).read())
    out2 = os.path.join(workdir, 'c1x.jplace')
    run_apples.main(['-s', ref, '-x', ext, '-t', tree, '-D', '--clusters', tsv, '-o', out2])
    _check_jplace(out2, 'c1_align_FM_MLSE', workdir, 'backbone.nwk')
    # database round trip (build_applesdtb.py, run_apples.py -a)
    dtb = os.path.join(workdir, 'apples.dtb')
    build_applesdtb.main(['-s', ref, '-t', tree, '-D', '--clusters', tsv, '-o', dtb])
    out3 = os.path.join(workdir, 'c1a.jplace')
    run_apples.main(['-a', dtb, '-q', qry, '-o', out3, '-m', 'OLS'])
    _check_jplace(out3, 'c1_align_OLS_MLSE', workdir, 'backbone.nwk')


def test_cli_matrix(workdir):
    import run_apples
    mat, tree = util.gunzip_to('dist.mat', workdir), util.gunzip_to('backbone.nwk', workdir)
    out = os.path.join(workdir, 'c2.jplace')
    run_apples.main(['-d', mat, '-t', tree, '-o', out])
    _check_jplace(out, 'c2_matrix_FM_MLSE', workdir, 'backbone.nwk')
    out = os.path.join(workdir, 'c2b.jplace')
    run_apples.main(['-d', mat, '-t', tree, '-o', out, '-m', 'BME', '-c', 'ME', '-b', '5', '-f', '0.05'])
    _check_jplace(out, 'c2_matrix_BME_ME_b5', workdir, 'backbone.nwk')
    out = os.path.join(workdir, 'c2c.jplace')
    run_apples.main(['-d', mat, '-t', tree, '-o', out, '-m', 'BE', '-n'])
    _check_jplace(out, 'c2_matrix_BE_MLSE_neg', workdir, 'backbone.nwk')


def test_cli_protein(workdir):
    import run_apples
    ref, qry, tree, tsv = (util.gunzip_to(n, workdir) for n in ('prot_ref.fa', 'prot_query.fa', 'prot_backbone.nwk', 'prot.tsv'))
    out = os.path.join(workdir, 'c3.jplace')
    run_apples.main(['-s', ref, '-q', qry, '-t', tree, '-p', '-f', '0.6', '-b', '25', '-D', '--clusters', tsv, '-o', out])
    _check_jplace(out, 'c3_prot_FM_MLSE', workdir, 'prot_backbone.nwk')


def test_database_threshold_comes_from_the_database(workdir):
    """with `-a database` the cluster-expansion threshold is the one stored in the reference object at build
    time (Reference.py:146 uses self.threshold), whatever -f the run is given: a database built with -f 0.45 and run
    with the default -f reproduces the reference's `-f 0.45` observed sets."""
    import run_apples
    import build_applesdtb
    ref, qry, tree = (util.gunzip_to(n, workdir) for n in ('ref.fa', 'query.fa', 'backbone.nwk'))
    tsv45 = os.path.join(util.GOLD, 'c1_clusters_f045.tsv')
    dtb = os.path.join(workdir, 'apples_f045.dtb')
    build_applesdtb.main(['-s', ref, '-t', tree, '-D', '-f', '0.45', '--clusters', tsv45, '-o', dtb])
    out = os.path.join(workdir, 'c1_f045.jplace')
    run_apples.main(['-a', dtb, '-q', qry, '-b', '5', '-o', out])          # run-time -f left at its default 0.2
    _check_jplace(out, 'c1_align_FM_MLSE_f045_b5', workdir, 'backbone.nwk')
    # a pickle that is not one of this build's databases is refused with a clear message
    bogus = os.path.join(workdir, 'bogus.dtb')
    with open(bogus, 'wb') as f:
        pickle.dump({'not': 'a tree'}, f)
    with pytest.raises(SystemExit):
        run_apples.main(['-a', bogus, '-q', qry, '-o', out])


def _n_gpus():
    from apples_b200 import _lib
    return _lib.device_count()


def test_multi_gpu_result_is_byte_identical(workdir):
    """queries sharded over N GPUs (one context and one host thread per GPU, in-process) give result
    arrays BYTE-IDENTICAL to the single-GPU arrays for the same queries; same for the jplace dicts of place_batch."""
    import types
    import numpy as np
    from apples_b200 import _lib, fasta
    from apples_b200.placer import GpuPlacer, MultiGpuPlacer, place_batch
    from tests.test_gpu_parity import _synthetic
    n = _n_gpus()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    nwk, tree, refs, ref, queries, _ = _synthetic(4000, 1500, 9001, 700, workdir)
    ref.set_baseobs(25)
    names = list(queries.keys())
    mat = fasta.as_byte_matrix([queries[k] for k in names], 1500)
    params = _lib.make_params('FM', 'MLSE')
    one = GpuPlacer(tree, ref, tree.name_to_node, device=0)
    a = one.place_bytes(mat, None, params)
    one.close()
    for g in sorted({2, n}):
        multi = MultiGpuPlacer(tree, ref, tree.name_to_node, devices=list(range(g)))
        b = multi.place_bytes(mat, None, params)
        multi.close()
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes(), g
    opt = types.SimpleNamespace(method_name='OLS', criterion_name='MLSE', negative_branch=False,
                                base_observation_threshold=25, filt_threshold=0.2, minimum_alignment_overlap=0.001,
                                exclude_intplace=False)
    qlist = [(k, queries[k], None) for k in names[:777]]
    r1 = place_batch(ref, opt, tree.name_to_node, qlist, tree=tree, device=0)
    r2 = place_batch(ref, opt, tree.name_to_node, qlist, tree=tree, devices=list(range(n)))
    assert json.dumps(r1) == json.dumps(r2)


def test_torchrun_cli_equals_single_process(workdir):
    """run_apples.py under torchrun (one process per GPU, NCCL all-gather of 32-byte records, rank 0 writes) produces
    the same jplace file as the single-process run."""
    import subprocess
    import sys
    if _n_gpus() < 2:
        pytest.skip('needs at least 2 GPUs')
    ref, qry, tree = (util.gunzip_to(n, workdir) for n in ('ref.fa', 'query.fa', 'backbone.nwk'))
    tsv = os.path.join(util.GOLD, 'c1_clusters.tsv')
    out1, out2 = os.path.join(workdir, 'tr1.jplace'), os.path.join(workdir, 'tr2.jplace')
    base = ['-s', ref, '-q', qry, '-t', tree, '-D', '--clusters', tsv]
    subprocess.run([sys.executable, os.path.join(util.ROOT, 'run_apples.py')] + base + ['-o', out1, '--gpus', '1'], check=True)
    subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                    '127.0.0.1', '--master-port', '29733', os.path.join(util.ROOT, 'run_apples.py')] + base + ['-o', out2],
                   check=True, timeout=600)
    a, b = json.load(open(out1)), json.load(open(out2))
    a.pop('metadata'), b.pop('metadata')      # the invocation line differs (output path, launcher)
    assert a == b
    _check_jplace(out2, 'c1_align_FM_MLSE', workdir, 'backbone.nwk')
