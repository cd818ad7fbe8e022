"""jplace assembly (host).  Mirrors apples/jutil.py:1-19 (join_jplace) and the tail of run_apples.py:106-118."""
import json
import sys


def join_jplace(lst):
    """Merge per-query dicts into one.  Same quirk as jutil.py:11-19: with a single result an unplaceable query
    (edge_num == -1) empties the list; with several, the FIRST result is kept even if its edge_num is -1 and later
    unplaceable ones are dropped."""
    result = lst[0]
    if len(lst) == 1:
        if result['placements'][0]['p'][0][0] == -1:
            result['placements'] = []
        return result
    merged = list(result['placements'])
    for item in lst[1:]:
        if item['placements'][0]['p'][0][0] != -1:
            merged.extend(item['placements'])
    result['placements'] = merged
    return result


def assemble(results, extended_newick_string, argv=None):
    """run_apples.py:106-110"""
    result = join_jplace(results)
    result['tree'] = extended_newick_string
    result['metadata'] = {'invocation': ' '.join(sys.argv if argv is None else argv)}
    result['fields'] = ['edge_num', 'likelihood', 'like_weight_ratio', 'distal_length', 'pendant_length']
    result['version'] = 3
    return result


def write(result, output_fp=None):
    """run_apples.py:112-118"""
    f = open(output_fp, 'w') if output_fp else sys.stdout
    f.write(json.dumps(result, sort_keys=True, indent=4))
    f.write('\n')
    if output_fp:
        f.close()
