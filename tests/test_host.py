"""CPU: host-side logic of the product (tree layout, FASTA input, packing, clusters, jplace assembly)."""
import hashlib
import io
import json
import os

import numpy as np
import pytest

from tests import util
from apples_b200 import fasta, jplace, synth, treecluster
from apples_b200.placer import results_to_jplace
from apples_b200.reference import ReducedReference, consensus_rows
from apples_b200.tree import BackboneTree


def test_tree_layout_small(workdir):
    t = BackboneTree.from_newick('((A:0.1,B:0.2):0.25,(C:0.3,(D:0.2,E:0.2):0.2):0.25);')
    # post-order ranks: A0 B1 (AB)2 C3 D4 E5 (DE)6 (C(DE))7 root8
    assert t.name_to_node == {'A': 0, 'B': 1, 'C': 3, 'D': 4, 'E': 5}
    assert t.parent.tolist() == [2, 2, 8, 7, 6, 6, 7, 8, -1]
    assert t.level.tolist() == [2, 2, 1, 2, 3, 3, 2, 1, 0]
    assert t.first.tolist() == [0, 1, 0, 3, 4, 5, 4, 3, 0]
    assert t.children_of(8) == [2, 7] and t.children_of(7) == [3, 6]
    assert t.extended_newick() == '((A:0.1{0},B:0.2{1}):0.25{2},(C:0.3{3},(D:0.2{4},E:0.2{5}):0.2{6}):0.25{7});'


@pytest.mark.parametrize('case,fn', [('c2_matrix_FM_MLSE', 'backbone.nwk'), ('c3_prot_FM_MLSE', 'prot_backbone.nwk'),
                                     ('syn300_FM_MLSE_pos', 'syn300.nwk')])
def test_extended_newick_matches_reference(case, fn, workdir):
    """the reference's extended newick (jutil.py:22-96) of the same file, by hash (oracle/gen_golden.py)"""
    t = BackboneTree.from_newick(util.gunzip_to(fn, workdir))
    assert hashlib.sha256(t.extended_newick().encode()).hexdigest() == util.load_golden(case)['extended_newick_sha']


def test_fasta2dic(tmp_path):
    p = tmp_path / 'x.fa'
    p.write_text('>s1 some description\nACGTnn-x\nRYacgt\n>s2\nAC.T*\n')
    d = fasta.fasta2dic(str(p), False, False)
    assert list(d) == ['s1', 's2']
    assert d['s1'].tobytes() == b'ACGT------ACGT'  # N, X, R, Y -> '-' (fasta2dic.py:61)
    assert d['s2'].tobytes() == b'AC.T*'           # non-letters survive, like the reference
    d = fasta.fasta2dic(str(p), False, True)
    assert d['s1'].tobytes() == b'ACGT' + b'-' * 10  # lower-case masked (fasta2dic.py:52-54); R, Y invalid
    d = fasta.fasta2dic(str(p), True, False)
    assert d['s1'].tobytes() == b'ACGTNN--RYACGT'  # protein alphabet: only B J O U X Z become '-' (fasta2dic.py:59)
    p2 = tmp_path / 'x.fq'
    p2.write_text('@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+\nIIII\n')
    d = fasta.fasta2dic(str(p2), False, False)
    assert {k: v.tobytes() for k, v in d.items()} == {'r1': b'ACGT', 'r2': b'GGCC'}


def test_pack_nucleotide_roundtrip():
    rng = np.random.default_rng(5)
    L = 1234
    mat = np.frombuffer(b'ACGT-', dtype=np.uint8)[rng.integers(0, 5, (7, L))]
    pk = fasta.pack_nucleotide(mat)
    W = fasta.words_per_row(L)
    assert pk.shape == (7, 3, W) and W % 4 == 0 and pk.dtype == np.uint32
    bits = np.unpackbits(pk.view(np.uint8).reshape(7, 3, W * 4), axis=2, bitorder='little')[:, :, :L]
    code = bits[:, 0] + 2 * bits[:, 1]
    back = np.where(bits[:, 2] == 1, np.frombuffer(b'ACGT', dtype=np.uint8)[code], ord('-'))
    assert (back == mat).all()
    assert (bits[:, 0][bits[:, 2] == 0] == 0).all() and (bits[:, 1][bits[:, 2] == 0] == 0).all()
    # counts from planes == counts from bytes (distance.py:733-737)
    a, b = pk[0], pk[1]
    v = a[2] & b[2]
    m = ((a[0] ^ b[0]) | (a[1] ^ b[1])) & v
    pop = lambda x: int(np.unpackbits(x.view(np.uint8)).sum())
    nd = (mat[0] != ord('-')) & (mat[1] != ord('-'))
    assert pop(v) == int(nd.sum()) and pop(m) == int(((mat[0] != mat[1]) & nd).sum())
    with pytest.raises(ValueError):
        fasta.pack_nucleotide(np.frombuffer(b'AC.T', dtype=np.uint8)[None, :])


def test_pack_protein():
    mat = np.frombuffer(b'ARNDCQEGHILKMFPSTWYV-*a', dtype=np.uint8)[None, :]
    pk = fasta.pack_protein(mat)
    assert pk.shape == (1, 32)
    assert pk[0, :23].tolist() == list(range(20)) + [20, 0, 0]
    assert (pk[0, 23:] == 20).all()


def test_clusters_and_consensus(workdir):
    from oracle import apples_oracle as orc
    ci = util.CaseInputs('syn300_FM_MLSE_pos', workdir)
    tree, ref = ci.product_state()
    reps = orc.representatives_from_tsv(ci.tsv, ci.refs, False)
    assert len(reps) == len(ref.representatives)
    for (c0, g0), (c1, g1) in zip(reps, ref.representatives):
        assert g0 == g1 and c0.tobytes() == c1.tobytes()
    # in-repo clustering: a partition of the leaves whose clusters respect the diameter bound
    cl = treecluster.max_diameter_clusters(tree, 0.24)
    leaves = sorted(u for c in cl for u in c)
    assert leaves == tree.leaf_ids.tolist()
    dev = ref.device_arrays(tree.name_to_node)
    assert dev['packed_refs'].shape[0] == 300 and dev['group_offsets'][-1] == 300
    assert sorted(dev['group_members'].tolist()) == list(range(300))


def test_cluster_diameter_bound():
    nwk = synth.random_tree(120, seed=3)
    t = BackboneTree.from_newick(nwk)
    cl = treecluster.max_diameter_clusters(t, 0.1)
    depth = np.zeros(t.num_nodes)
    for u in range(t.num_nodes - 2, -1, -1):
        depth[u] = depth[t.parent[u]] + max(t.edge_length[u], 0)

    def dist(a, b):
        x, y = a, b
        while x != y:
            if x < y:
                x = t.parent[x]
            else:
                y = t.parent[y]
        return depth[a] + depth[b] - 2 * depth[x]
    for c in cl:
        for i in range(len(c)):
            for j in range(i + 1, len(c)):
                assert dist(c[i], c[j]) <= 0.1 + 1e-12


def test_results_to_jplace_and_join():
    out = (np.array([5, 7, -1, 9], np.int32), np.array([0.5, 0, 0, 0.25]), np.array([0.1, 0, 0, 0.2]),
           np.array([0.3, 0, 0, 0.0]), np.array([0, 1, 2, 3 | 0x100], np.int32))
    res = results_to_jplace(['a', 'b', 'c', 'd'], [False, True, False, False], out, log=False)
    assert res[0] == {'placements': [{'p': [[5, 0.5, 1, 0.1, 0.3]], 'n': ['a']}]}
    assert res[1] == {'placements': [{'p': [[7, 0, 1, 0, 0]], 'n': ['b-query']}]}
    assert res[2]['placements'][0]['p'] == [[-1, 0, 1, 0, 0]]
    assert res[3]['placements'][0]['p'] == [[9, 0.25, 1, 0.2, 0]] and isinstance(res[3]['placements'][0]['p'][0][4], int)
    ex = results_to_jplace(['d'], [False], tuple(o[3:] for o in out), exclude_intplace=True, log=False)
    assert ex[0]['placements'][0]['p'][0][0] == -1
    joined = jplace.join_jplace(json.loads(json.dumps(res)))
    assert [p['n'][0] for p in joined['placements']] == ['a', 'b-query', 'd']  # jutil.py:11-19


def test_synth_determinism():
    a = synth.random_tree(50, seed=9)
    assert a == synth.random_tree(50, seed=9) and a != synth.random_tree(50, seed=10)
    t = BackboneTree.from_newick(a)
    assert len(t.leaf_ids) == 50 and t.nchild[t.num_nodes - 1] == 3


def test_option_validation():
    from apples_b200.options import options_config_run
    with pytest.raises(ValueError):
        options_config_run(['-d', 'x.mat', '-s', 'ref.fa', '-t', 't.nwk'])   # OptionsRun.py:90-91
    with pytest.raises(ValueError):
        options_config_run(['-q', 'q.fa'])                                     # OptionsRun.py:106-107
    with pytest.raises(ValueError):
        options_config_run(['-t', 't.nwk', '-q', 'q.fa', '-x', 'e.fa'])       # OptionsRun.py:108-109
    o, _ = options_config_run(['-t', 't.nwk', '-q', 'q.fa', '-s', 'r.fa'])
    assert (o.method_name, o.criterion_name, o.base_observation_threshold, o.filt_threshold,
            o.minimum_alignment_overlap) == ('FM', 'MLSE', 25, 0.2, 0.001)


def test_jplace_dumps_is_json_dumps_byte_for_byte():
    """jplace.dumps writes the placement records from a template; the text must equal
    json.dumps(result, sort_keys=True, indent=4) (run_apples.py:114) including special floats, escapes, the empty
    placement list of a single unplaceable query, and it must fall back for anything unexpected."""
    import json
    import random
    from apples_b200 import jplace
    rnd = random.Random(7)

    def num():
        r = rnd.random()
        if r < 0.1:
            return 0
        if r < 0.2:
            return 1
        if r < 0.25:
            return float('nan')
        if r < 0.3:
            return float('inf')
        if r < 0.35:
            return -float('inf')
        if r < 0.4:
            return -0.0
        if r < 0.5:
            return rnd.random() * 1e-300
        if r < 0.6:
            return rnd.random() * 1e22
        return rnd.random()

    def results(n):
        out = []
        for i in range(n):
            name = rnd.choice(['Q%d' % i, 'a"b\\c', 'ü-name', 'tab\tname', 'x' * 50])
            out.append({'placements': [{'p': [[rnd.choice([-1, rnd.randint(0, 10 ** 6)]), num(), 1, num(), num()]],
                                        'n': [name]}]})
        return out

    for n in (1, 2, 3, 50, 500):
        for _ in range(4):
            a = jplace.assemble(results(n), '((a:1,b:2):0.1,c);', argv=['run_apples.py', '-q', 'x y'])
            assert jplace.dumps(a) == json.dumps(a, sort_keys=True, indent=4)
    a = jplace.assemble([{'placements': [{'p': [[-1, 0, 1, 0, 0]], 'n': ['q']}]}], '(a,b);', argv=['x'])
    assert a['placements'] == [] and jplace.dumps(a) == json.dumps(a, sort_keys=True, indent=4)
    odd = {'placements': [{'p': [[1, 2, 3]], 'n': ['q']}], 'tree': 'x'}
    assert jplace.dumps(odd) == json.dumps(odd, sort_keys=True, indent=4)
    nested = {'placements': [{'p': [[1, 2.0, 1, 3.0, 4.0]], 'n': ['q'], 'extra': 1}], 'tree': 'x'}
    assert jplace.dumps(nested) == json.dumps(nested, sort_keys=True, indent=4)


def test_results_to_jplace_fast_path_equals_per_record_logic():
    """results_to_jplace builds the common records in one pass and runs the runquery control flow
    (PoolQueryWorker.py:63-130) only for the special ones; compare with the plain per-record restatement."""
    from apples_b200 import _lib
    from apples_b200.placer import results_to_jplace
    rng = np.random.default_rng(5)
    n = 400
    edge = rng.integers(0, 1000, n).astype(np.int32)
    error, distal, pendant = rng.random(n), rng.random(n), rng.random(n)
    status = rng.choice([_lib.PLACED, _lib.PLACED, _lib.PLACED, _lib.ZERO_DIST_LEAF, _lib.TOO_FEW_DISTANCES,
                         _lib.PLACED_MISPLACEMENT_FLAG, _lib.PLACED | _lib.FLAG_PENDANT_INT0,
                         _lib.PLACED_MISPLACEMENT_FLAG | _lib.FLAG_PENDANT_INT0], n).astype(np.int32)
    in_bb = (rng.random(n) < 0.1).tolist()
    names = ['q%d' % i for i in range(n)]
    for excl in (False, True):
        got = results_to_jplace(names, in_bb, (edge, error, distal, pendant, status), exclude_intplace=excl, log=False)
        assert len(got) == n
        for i in range(n):
            name = names[i] + ('-query' if in_bb[i] else '')
            code = int(status[i]) & _lib.STATUS_CODE_MASK
            if code == _lib.ZERO_DIST_LEAF:
                p = [int(edge[i]), 0, 1, 0, 0]
            elif code == _lib.TOO_FEW_DISTANCES:
                p = [-1, 0, 1, 0, 0]
            else:
                pend = 0 if int(status[i]) & _lib.FLAG_PENDANT_INT0 else float(pendant[i])
                p = [int(edge[i]), float(error[i]), 1, float(distal[i]), pend]
                if code == _lib.PLACED_MISPLACEMENT_FLAG and excl:
                    p[0] = -1
            exp = {'placements': [{'p': [p], 'n': [name]}]}
            assert got[i] == exp, i
            assert [type(x) for x in got[i]['placements'][0]['p'][0]] == [type(x) for x in p], i


def test_native_fasta_reader_equals_python_twin(tmp_path, workdir):
    """hostio.cpp apples_fasta_open (C++ twin of fasta2dic) gives the same names and bytes as the Python
    fasta2dic on the fixtures and on a file with multi-line records, FASTQ records, CRLF, blank lines, lower case, letters
    outside the alphabet and non-letter symbols; nucleotide / protein x mask on / off."""
    from apples_b200 import fasta
    files = [(util.gunzip_to('ref.fa', workdir), False), (util.gunzip_to('query.fa', workdir), False),
             (util.gunzip_to('prot_query.fa', workdir), True)]
    edge = tmp_path / 'edge.fa'
    with open(edge, 'w', newline='') as f:
        f.write('junk before\n>a desc here\nacgtNn-\r\nAC.GT\n\n>b\nACGTacgtXYZ\n@fq1 x\nACGTN\n+\nIIIII\n@fq2\nAC\nGT\n+fq2\nII\nII\n'
                '>dup\nAAAA\n>dup\nCCCC\n>last\nACGT*?')
    files += [(str(edge), False), (str(edge), True)]
    for fp, prot in files:
        for mask in (False, True):
            d = fasta.fasta2dic(fp, prot, mask)
            m = fasta.read_alignment(fp, prot, mask, pinned=False)
            d2 = m.as_dict()
            assert list(d) == list(d2), fp
            for k in d:
                assert d[k].tobytes() == d2[k].tobytes(), (fp, k, mask)
            assert m.n >= len(d) and m.stride % 16 == 0 and m.stride >= m.L
            if m.uniform and m.n:
                assert (m.full[:, m.L:] == ord('-')).all()
            m.close()
    with pytest.raises(OSError):
        fasta.read_alignment(str(tmp_path / 'missing.fa'))


def test_native_jplace_writer_equals_python_writer(tmp_path):
    """hostio.cpp apples_jplace_write writes, from the result arrays, the bytes that
    json.dumps(join_jplace(per-query dicts), sort_keys=True, indent=4) gives (run_apples.py:106-118, jutil.py:1-19):
    Python float repr (exponent thresholds, -0.0, denormals, nan/inf spellings), int 0 where the reference has ints,
    ASCII escapes of names incl. non-BMP characters, '-query' names, --exclude, the first-record quirk, empty lists."""
    import json
    from apples_b200 import jplace
    from apples_b200.placer import results_to_jplace
    vals = [0.0, -0.0, 1e-5, 1e-4, 0.0001234, 1e15, 1e16, 1.5e-7, 123456789.123, float('nan'), float('inf'), -float('inf'),
            5e-324, 1.7976931348623157e308, 0.1, 1 / 3, 2.0, 100.0, 1e22, 1e21, 123456789012345680.0, 9999999999999998.0]

    def case(n, first_bad, seed):
        rng = np.random.default_rng(seed)
        names = ['q%d' % i for i in range(n)]
        if n > 3:
            names[1], names[2], names[3] = 'na\u00efve "q"\\t\u2603', '\U0001F600x', 'tab\there\x7f'
        edge = rng.integers(0, 1000, n).astype(np.int32)

        def col():
            a = rng.random(n) * rng.choice([1e-18, 1e-9, 1e-3, 1.0, 1e6], n)
            k = min(n, len(vals))
            a[:k] = vals[:k]
            rng.shuffle(a)
            return a
        status = rng.choice([0, 0, 0, 0, 1, 2, 3, 0x100, 0x103], n).astype(np.int32)
        if first_bad:
            status[0] = 2
        return names, rng.random(n) < 0.1, (edge, col(), col(), col(), status)
    fp = str(tmp_path / 'o.jplace')
    for n, fb, ex, seed in [(1, False, False, 1), (1, True, False, 2), (2, True, True, 3), (40, False, False, 4),
                            (40, True, True, 5), (9000, False, True, 6), (3, True, False, 7)]:
        names, inb, out = case(n, fb, seed)
        res = results_to_jplace(names, inb.tolist(), out, ex, log=False, degenerate='keep')
        doc = jplace.assemble(res, '((a,b),c);', argv=['run_apples.py', '-x'])
        ref = json.dumps(doc, sort_keys=True, indent=4) + '\n'
        for th in (1, 3, 0):
            nw = jplace.write_arrays(fp, names, inb, out, '((a,b),c);', ex, argv=['run_apples.py', '-x'], threads=th)
            assert open(fp, encoding='utf-8').read() == ref, (n, fb, ex, th)
            assert nw == len(doc['placements'])


def _same_tree(a, b):
    assert (a.parent == b.parent).all() and (a.level == b.level).all() and (a.first == b.first).all()
    assert a.edge_length.tobytes() == b.edge_length.tobytes() and (a.has_length == b.has_length).all()
    assert a.label == b.label and a.is_rooted == b.is_rooted and a.name_to_node == b.name_to_node
    assert (a.nchild == b.nchild).all() and (a.leaf_ids == b.leaf_ids).all()


def test_native_newick_equals_python_twin(workdir):
    """apples_newick_parse / apples_newick_extended (hostio.cpp) against BackboneTree's Python parser and writer, which
    define the accepted language: the golden backbones, random trees with polytomies, and hand-written corner cases;
    text outside the native parser's language must come back as 'unsupported' and then parse through the Python twin."""
    from apples_b200.tree import BackboneTree
    texts = []
    for f in ('backbone.nwk', 'small_backbone.nwk', 'syn300.nwk', 'prot_backbone.nwk'):
        texts.append(open(util.gunzip_to(f, workdir)).read())
    for seed in range(5):
        texts.append(synth.random_tree(50 + 37 * seed, seed=seed, polytomy_frac=0.3 if seed % 2 else 0.0))
    texts += [
        '(a:1,b:2.5,(c:1e-3,d:0.0)x:3)root;',
        "[&R] ((a:1,b:2)90:0.5,('c d':1.25e+2,e:-0.5)'in ner':7);",
        '((a,b),(c,d));',                                   # no lengths
        '  (a:0.1[comment],b[&x=1]:2,(c:3,d:4)[c]:5);\n\n',  # comments, surrounding blanks
        '(a:1,\n b:2,\r\n\t(c:3 , d:4) : 5 ) ;',           # blanks everywhere
        '(a:10,b:3.0,(c:100000000000000000000,d:1e22):0.30000000000000004);',   # integral and shortest-repr lengths
        '(a:+1.5,b:.5,(c:5.,d:1E3):2)lab:9;',               # root with a length (not written: the root has no edge)
        'a;',                                               # a single node
        '(a:1,b:2)',                                        # no terminator
        "(a:1,'':2,(b:1,c:1)'':3);",                        # empty quoted labels are labels
        '(äö:1,中文:2,(x:1,y:1)ß:3);',      # non-ASCII labels
    ]
    n_native = 0
    for t in texts:
        py = BackboneTree.from_newick(t, native=False)
        auto = BackboneTree.from_newick(t)
        _same_tree(py, auto)
        nat = BackboneTree._from_newick_native(t)
        if nat is not None:
            n_native += 1
            _same_tree(py, nat)
        ext_py = py.extended_newick(native=False)
        assert auto.extended_newick() == ext_py
        ext_nat = auto._extended_newick_native()
        assert ext_nat is None or ext_nat == ext_py
    assert n_native >= len(texts) - 2     # the CJK label case and nothing much else may be declined
    # outside the native language: declined, and the public entry point still equals the Python twin (or raises like it)
    for t in ['(a:inf,b:nan);', '(a:1_0,b:2);', "(a:1,b:2]x);", "(a:1,'b:2);", '(a:0x10,b:1);', '(a:1,b:2));']:
        assert BackboneTree._from_newick_native(t) is None, t
        try:
            py = BackboneTree.from_newick(t, native=False)
        except Exception as e:   # the Python twin's own error (e.g. float('0x10')): the public entry point raises the same
            with pytest.raises(type(e)):
                BackboneTree.from_newick(t)
            continue
        _same_tree(py, BackboneTree.from_newick(t))
        assert BackboneTree.from_newick(t).extended_newick() == py.extended_newick(native=False)
    # integral lengths beyond 9e18 and non-finite lengths in the writer
    big = BackboneTree.from_newick('(a:1e30,b:2);', native=False)
    assert big._extended_newick_native() is None and big.extended_newick() == big.extended_newick(native=False)
    nf = BackboneTree.from_newick('(a:inf,b:nan,c:-inf);', native=False)
    assert nf.extended_newick() == nf.extended_newick(native=False) == '(a:inf{0},b:nan{1},c:-inf{2});'


def test_native_newick_fuzz_against_python_twin():
    """30 000 random strings over the newick punctuation: whatever the native parser accepts must give the arrays and
    the extended newick of the Python twin (it may decline anything; it must never differ or raise on its own)."""
    import random
    from apples_b200.tree import BackboneTree
    rnd = random.Random(20260117)
    alpha = "(((()))),,,,::;[]' ab1.e-+5 \n\t"
    accepted = 0
    for _ in range(30000):
        t = ''.join(rnd.choice(alpha) for _ in range(rnd.randint(1, 24)))
        nat = BackboneTree._from_newick_native(t)
        if nat is None:
            continue
        accepted += 1
        py = BackboneTree.from_newick(t, native=False)
        _same_tree(py, nat)
        ext = nat._extended_newick_native()
        assert ext is None or ext == py.extended_newick(native=False), t
    assert accepted > 3000
