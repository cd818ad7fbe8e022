"""Shared helpers for the tests: golden cases, their inputs, oracle and device contexts."""
import glob
import gzip
import json
import os
import re
import types

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
GOLD = os.path.join(ROOT, 'tests', 'golden')
DATA = os.path.join(GOLD, 'data')


def gunzip_to(name, workdir):
    """tests/golden/data/<name>.gz -> <workdir>/<name> (cached)."""
    dst = os.path.join(workdir, name)
    if not os.path.exists(dst):
        # written aside and renamed: the ranks of a multi-process test extract the same file at the same time, and a rank
        # must never see a half-written one
        tmp = '%s.%d.tmp' % (dst, os.getpid())
        with gzip.open(os.path.join(DATA, name + '.gz'), 'rb') as fi, open(tmp, 'wb') as fo:
            fo.write(fi.read())
        os.replace(tmp, dst)
    return dst


def unhex(v):
    return float.fromhex(v) if isinstance(v, str) else v


def golden_names(pattern='*'):
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLD, pattern + '.json')))


def load_golden(name):
    with open(os.path.join(GOLD, name + '.json')) as f:
        return json.load(f)


def options_of(g):
    o = types.SimpleNamespace(**g['options'])
    return o


def read_dismat(path):
    """same parsing as run_apples.py:43-54"""
    with open(path) as f:
        tags = list(re.split(r"\s+", f.readline().rstrip()))[1:]
        rows = []
        for line in f.readlines():
            d = list(re.split(r"\s+", line.strip()))
            rows.append((d[0], None, dict(zip(tags, map(float, d[1:])))))
    return rows


class CaseInputs:
    """Inputs of one golden case: tree file, queries [(name, seq|None, row|None)], refs, cluster tsv, protein flag."""

    def __init__(self, name, workdir):
        from apples_b200.fasta import fasta2dic
        self.name = name
        self.g = load_golden(name)
        self.options = options_of(self.g)
        self.protein = name.startswith('c3_prot')
        self.refs = None
        self.tsv = None
        fam = name.split('_')[0]
        if fam == 'small':
            self.tree_fp = gunzip_to('small_backbone.nwk', workdir)
            self.queries = read_dismat(gunzip_to('small_dist.mat', workdir))
        elif fam == 'c2':
            self.tree_fp = gunzip_to('backbone.nwk', workdir)
            self.queries = read_dismat(gunzip_to('dist.mat', workdir))
        elif fam == 'c1':
            self.tree_fp = gunzip_to('backbone.nwk', workdir)
            self.refs = fasta2dic(gunzip_to('ref.fa', workdir), False, False)
            q = fasta2dic(gunzip_to('query.fa', workdir), False, False)
            self.tsv = os.path.join(GOLD, 'c1_clusters_f045.tsv' if 'f045' in name else 'c1_clusters.tsv')
            qlist = [(k, v, None) for k, v in q.items()]
            if 'special' in name:
                names = list(self.refs.keys())
                some = names[17]
                gap = np.frombuffer(b'-' * len(self.refs[some]), dtype='S1')
                thin = gap.copy()
                thin[100:103] = self.refs[names[3]][100:103]
                special = [(some, self.refs[some], None), ('copy_of_' + names[40], self.refs[names[40]], None),
                           ('allgap', gap, None), ('thin', thin, None)]
                qlist = special + (qlist[:3] if 'exclude' in name else [])
            self.queries = qlist
        elif fam == 'syn300':
            self.tree_fp = gunzip_to('syn300.nwk', workdir)
            self.refs = fasta2dic(gunzip_to('syn300_ref.fa', workdir), False, False)
            q = fasta2dic(gunzip_to('syn300_query.fa', workdir), False, False)
            self.tsv = gunzip_to('syn300.tsv', workdir)
            self.queries = [(k, v, None) for k, v in q.items()]
        elif fam == 'c3':
            self.tree_fp = gunzip_to('prot_backbone.nwk', workdir)
            self.refs = fasta2dic(gunzip_to('prot_ref.fa', workdir), True, False)
            q = fasta2dic(gunzip_to('prot_query.fa', workdir), True, False)
            self.tsv = gunzip_to('prot.tsv', workdir)
            self.queries = [(k, v, None) for k, v in q.items()]
        else:
            raise KeyError(name)
        assert [q[0] for q in self.queries] == [r['name'] for r in self.g['queries']], name

    def oracle_context(self):
        from oracle import apples_oracle as orc
        tree, names = orc.load_tree(self.tree_fp)
        reps = None
        if self.refs is not None:
            reps = orc.representatives_from_tsv(self.tsv, self.refs, self.protein)
        o = self.options
        return orc.OracleContext(tree, names, refs=self.refs, representatives=reps, protein=self.protein,
                                 method=o.method_name, criterion=o.criterion_name, negative_branch=o.negative_branch,
                                 filt_threshold=o.filt_threshold, baseobs=o.base_observation_threshold,
                                 overlap=o.minimum_alignment_overlap, exclude_intplace=o.exclude_intplace)

    def product_state(self):
        """(BackboneTree, ReducedReference or None) built by the product's own host code."""
        from apples_b200.tree import BackboneTree
        from apples_b200.reference import ReducedReference
        tree = BackboneTree.from_newick(self.tree_fp)
        ref = None
        if self.refs is not None:
            ref = ReducedReference(None, self.protein, self.tree_fp, self.options.filt_threshold, 1,
                                   cluster_tsv=self.tsv, tree=tree, refs=self.refs)
            ref.set_baseobs(self.options.base_observation_threshold)
        return tree, ref


def close(a, b, rel=1e-9, floor=0.0):
    return abs(a - b) <= rel * max(abs(a), abs(b)) + floor
