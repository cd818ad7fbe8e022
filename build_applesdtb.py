#!/usr/bin/env python3
"""build_applesdtb.py: packages tree + name map + extended newick + reduced reference into one pickle, the same
four sequential dumps as the reference (build_applesdtb.py:23-28), readable by run_apples.py -a."""
import logging
import pickle
import time

from apples_b200.options import options_config_build
from apples_b200.reference import ReducedReference
from apples_b200.tree import prepare_tree


def main(argv=None):
    startb = time.time()
    options, args = options_config_build(argv)
    tree, name_to_node_map, extended_newick_string = prepare_tree(options.tree_fp)
    start = time.time()
    reference = ReducedReference(options.ref_fp, options.protein_seqs, options.tree_fp, options.filt_threshold,
                                 options.num_thread, cluster_tsv=options.cluster_fp, tree=tree)
    logging.info('[%s] Reduced reference is computed in %.3f seconds.' % (time.strftime('%H:%M:%S'), time.time() - start))
    with open(options.output_fp, 'wb') as f:
        p = pickle.Pickler(f)
        p.dump(tree)
        p.dump(name_to_node_map)
        p.dump(extended_newick_string)
        p.dump(reference)
    logging.warning('[%s] APPLES database is built in %.3f seconds.' % (time.strftime('%H:%M:%S'), time.time() - startb))


if __name__ == '__main__':
    main()
