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
