#!/usr/bin/env python3
"""run_apples.py with the per-query worker pool replaced by the B200 hot path.

Same options, inputs and jplace output as the reference's run_apples.py; the one changed call is
`pool.starmap(queryworker.runquery, queries)` (reference run_apples.py:94-102) -> apples_b200.placer.place_batch.
"""
import logging
import pickle
import re
import sys
import time

from apples_b200 import jplace
from apples_b200.fasta import fasta2dic
from apples_b200.options import options_config_run
from apples_b200.placer import place_batch
from apples_b200.reference import ReducedReference
from apples_b200.tree import prepare_tree


def read_dismat(f):
    """run_apples.py:43-54"""
    tags = list(re.split(r"\s+", f.readline().rstrip()))[1:]
    for line in f.readlines():
        dists = list(re.split(r"\s+", line.strip()))
        yield (dists[0], None, dict(zip(tags, map(float, dists[1:]))))


def main(argv=None):
    startb = time.time()
    options, args = options_config_run(argv)
    logging.info('[%s] Options are parsed.' % time.strftime('%H:%M:%S'))
    tree = name_to_node_map = extended_newick_string = None
    up = fdtb = None
    if options.database_fp:
        start = time.time()
        fdtb = open(options.database_fp, 'rb')
        up = pickle.Unpickler(fdtb)
        tree = up.load()
        name_to_node_map = up.load()
        extended_newick_string = up.load()
        logging.info('[%s] Tree is loaded from APPLES database in %.3f seconds.' % (time.strftime('%H:%M:%S'),
                                                                                   time.time() - start))
    if options.tree_fp:
        tree, name_to_node_map, extended_newick_string = prepare_tree(options.tree_fp)

    if options.dist_fp:
        reference = None
        start = time.time()
        with open(options.dist_fp) as f:
            queries = list(read_dismat(f))
    else:
        start = time.time()
        if options.ref_fp:
            reference = ReducedReference(options.ref_fp, options.protein_seqs, options.tree_fp, options.filt_threshold,
                                         options.num_thread, cluster_tsv=options.cluster_fp, tree=tree)
            logging.info('[%s] Reduced reference is computed in %.3f seconds.' % (time.strftime('%H:%M:%S'),
                                                                                  time.time() - start))
        else:
            reference = up.load()
            fdtb.close()
            logging.info('[%s] Reduced reference is loaded from APPLES database in %.3f seconds.'
                         % (time.strftime('%H:%M:%S'), time.time() - start))
        reference.set_baseobs(options.base_observation_threshold)
        start = time.time()
        if options.query_fp:
            query_dict = fasta2dic(options.query_fp, options.protein_seqs, options.mask_lowconfidence)
        else:
            extended = fasta2dic(options.extended_ref_fp, options.protein_seqs, options.mask_lowconfidence)
            query_dict = {k: v for k, v in extended.items() if k not in reference.refs}
        queries = [(name, seq, None) for name, seq in query_dict.items()]
    logging.info('[%s] Query sequences are prepared in %.3f seconds.' % (time.strftime('%H:%M:%S'), time.time() - start))

    startq = time.time()
    results = place_batch(reference, options, name_to_node_map, queries, tree=tree, device=options.device)
    logging.info('[%s] Processed all queries in %.3f seconds.' % (time.strftime('%H:%M:%S'), time.time() - startq))

    jplace.write(jplace.assemble(results, extended_newick_string), options.output_fp)
    logging.warning('[%s] APPLES finished in %.3f seconds.' % (time.strftime('%H:%M:%S'), time.time() - startb))


if __name__ == '__main__':
    main()
