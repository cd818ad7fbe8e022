#!/bin/bash
# Builds tools/microbench and lists, for every kernel instantiation, the SASS mnemonics inside its timed loop (between the
# two CS2R clock reads): shows that every counted operation of tools/microbench.cu is exactly one LOP3 / POPC / IADD3 / IMAD.
set -e
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
names=(LOP3 POPC IADD IMAD MIX_PLAIN MIX_CSA)
for m in 0 1 2 3 4 5; do
  echo "== k<${names[$m]}>: mnemonics per loop iteration (UNROLL 4 x ILP 4 = 16 steps per iteration)"
  cuobjdump -sass -fun "_Z1kILi${m}EEvPjjPxPy" tools/microbench 2>/dev/null | python3 -c "
import re,sys,collections
ins=[]
for ln in sys.stdin:
    m=re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)',ln)
    if m: ins.append((int(m.group(1),16), m.group(3), ln))
# loop body = from the target of the backward branch to the branch
lo=hi=None
for a,op,ln in ins:
    if op.startswith('BRA'):
        t=re.search(r'0x([0-9a-f]+)',ln)
        if t and int(t.group(1),16)<a: lo,hi=int(t.group(1),16),a
c=collections.Counter(op.split('.')[0] for a,op,ln in ins if lo is not None and lo<=a<=hi)
print('   ', dict(c.most_common()))
"
done
