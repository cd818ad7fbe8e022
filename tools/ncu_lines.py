#!/usr/bin/env python3
"""Attribute the warp-stall samples and executed instructions of an ncu report to CUDA source lines.

`ncu --page source --csv` lists per-SASS-instruction metrics but (in CLI mode) no source correlation; `nvdisasm -g` of
the same binary lists the SASS with `//## File "...", line N` markers.  The two listings have the same instruction
order, so they can be zipped (the script checks that every opcode matches).

    tools/ncu_lines.py gpurun_out/prof.ncu-rep apples_b200/libapples_b200.so select.sm_100a.cubin \
        '_Z13select_kernelILi0ELb0EEv10SelectArgs' [top_n]

The .so must be the build that was profiled (compiled with -lineinfo)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, cubin, func = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40   # argv[6]: optional substring of the demangled name (template args)
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    dis = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cubin)], capture_output=True, text=True, check=True).stdout
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip().startswith('.section') and '.text.' + func in l)
    end = next((i for i in range(start + 1, len(lines)) if lines[i].strip().startswith('.section')), len(lines))
    cur, seq = None, []
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            seq.append((m.group(2), cur))
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # a report with several captured launches lists them one after the other, each introduced by a "Kernel Name" row: take
    # the LAST launch whose demangled name starts like the function asked for (APPLES_NCU_PICK=n selects the n-th instead)
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    m0 = re.match(r'_Z(\d+)', func)
    want = func[m0.end():m0.end() + int(m0.group(1))] if m0 else func
    cand = [i for i in starts if want in rows[i][1]]
    if len(sys.argv) > 6:
        cand = [i for i in cand if sys.argv[6] in rows[i][1]]
    pick = int(os.environ.get('APPLES_NCU_PICK', '-1'))
    st = cand[pick]
    en = next((i for i in starts if i > st), len(rows))
    rows = rows[st:en]
    h = rows[1]
    si, ii, so_, ti = h.index('# Samples'), h.index('Instructions Executed'), h.index('Source'), h.index('Thread Instructions Executed')
    prof = [(r[so_].strip(), int(r[si]), int(r[ii]), int(r[ti])) for r in rows[2:] if len(r) > ii]
    n = min(len(seq), len(prof))
    bad = sum(1 for k in range(n) if seq[k][0].split()[0] != prof[k][0].split()[0])
    print('%s: %d SASS instructions in the binary, %d in the report, %d opcode mismatches' % (rows[0][1], len(seq), len(prof), bad))
    if bad or len(seq) != len(prof):
        print('WARNING: the binary is not the profiled build; line attribution is unreliable')
    samples, instr, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for k in range(n):
        key = seq[k][1] or ('?', 0)
        samples[key] += prof[k][1]
        instr[key] += prof[k][2]
        thr[key] += prof[k][3]
    tot = sum(samples.values()) or 1
    print('total samples %d, warp instructions %d' % (tot, sum(instr.values())))
    print('samples   %   warp-instr  lanes  file:line  source')
    cache = {}
    for key, s in samples.most_common(top):
        f, ln = key
        if f not in cache:
            try:
                cache[f] = open(f).read().splitlines()
            except OSError:
                cache[f] = []
        text = cache[f][ln - 1].strip()[:90] if 0 < ln <= len(cache[f]) else ''
        print('%7d %5.1f %11d %5.1f  %s:%d  %s' % (s, 100.0 * s / tot, instr[key], thr[key] / max(1, instr[key]),
                                                 os.path.basename(f), ln, text))


if __name__ == '__main__':
    main()
