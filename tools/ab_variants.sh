#!/bin/bash
# A/B of library variants on ONE box: tools/ab_variants.sh ROUNDS name1 name2 ...   ("base" = the in-tree library);
# BENCH_ARGS="--workload protein" adds arguments to the bench command
rounds=$1; shift
for i in $(seq $rounds); do
for v in "$@"; do
lib=""; [ "$v" != base ] && lib="APPLES_B200_LIB=build/variants/libapples_b200_$v.so"
env $lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cli --no-cpu-baseline --no-e2e $BENCH_ARGS > gpurun_out/ab.json 2>/dev/null
python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
print('%-10s'%sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v}, d['result_sha1_per_block'])
PY
done; done
