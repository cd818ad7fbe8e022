#!/bin/bash
# gpurun_out/ captures of tools/profile_round.sh -> the tracked summaries under profiles/ (run in the build container)
set -e
cd "$(dirname "$0")/.."
R=${1:-r02}
summarize() { ncu -i "$1" --page details 2>/dev/null | grep -v "^    OPT\|^          \|^    INF\|^    WRN" | grep -E "^\[|^  [a-z]|Duration|Throughput|Hit Rate|Registers Per|Issue Slots Busy|Ipc|Achieved Occupancy|Theoretical Occupancy|No Eligible|Active Warps|Shared Memory Config|Dynamic Shared|Warp Cycles Per Issued|Block Limit|Local|Elapsed Cycles|SM Frequency"; }
stalls() { ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    d=[(h[i],r[i]) for i in range(len(h)) if 'pcsamp_warps_issue_stalled' in h[i] and 'not_issued' not in h[i]]
    d=[(k,float(v.replace(',',''))) for k,v in d if v not in ('','n/a')]
    t=sum(v for k,v in d) or 1
    print(r[h.index('Kernel Name')], r[h.index('Grid Size')], ' '.join('%s %.1f'%(k.split('stalled_')[1],100*v/t) for k,v in sorted(d,key=lambda x:-x[1])[:7]))
"; }
raw() { ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
keys=sys.argv[1:]
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')], r[h.index('Grid Size')], ' | '.join('%s %s %s'%(k, r[h.index(k)], rows[1][h.index(k)]) for k in keys if k in h))
" "${@:2}"; }
{ echo "# ncu --set full --clock-control none -k regex:select_nuc_kernel -s 7 -c 7 (one step: six first-pass launches + the rerun launch <1>); tools/profile_round.sh"
  summarize gpurun_out/select.ncu-rep; echo; echo "# stall reasons (pc sampling, % of samples)"; stalls gpurun_out/select.ncu-rep
  echo; echo "# dram bytes per launch"; raw gpurun_out/select.ncu-rep dram__bytes_read.sum dram__bytes_write.sum; } > profiles/ncu_select_$R.txt
{ echo "# ncu --set full --clock-control none -k regex:place_ -s 6 -c 6 (one step: four shared-memory classes + the block-per-query launches); tools/profile_round.sh"
  summarize gpurun_out/place.ncu-rep; echo; echo "# stall reasons (pc sampling, % of samples)"; stalls gpurun_out/place.ncu-rep
  echo; echo "# dram bytes per launch"; raw gpurun_out/place.ncu-rep dram__bytes_read.sum dram__bytes_write.sum; } > profiles/ncu_place_$R.txt
{ echo "# ncu --set full --clock-control none -k regex:dense_tc_kernel -s 7 -c 1 (one 24 320-query launch); tools/profile_round.sh"
  summarize gpurun_out/dense_tc.ncu-rep; echo; echo "# raw metrics"
  raw gpurun_out/dense_tc.ncu-rep dram__bytes_read.sum dram__bytes_write.sum gpu__time_duration.sum lts__t_sector_hit_rate.pct l1tex__m_xbar2l1tex_read_bytes.sum sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed | tr '|' '\n'; } > profiles/ncu_dense_tc_$R.txt
python - "$R" <<'PY'
import csv, collections, sys
R = sys.argv[1]
rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 10]
h = rows[0]
agg = collections.OrderedDict(); lines = []
for r in rows[1:]:
    name = r[h.index('Kernel Name')]; v = float(r[h.index('Metric Value')].replace(',', '')); u = r[h.index('Metric Unit')]
    v = v / 1e6 if u == 'ns' else v / 1e3 if u in ('us', 'usecond') else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    lines.append('%-4s %-62s grid %-14s block %-12s %9.3f ms' % (r[0], name[:62], r[h.index('Grid Size')], r[h.index('Block Size')], v))
tot = sum(a[1] for a in agg.values())
out = ['# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:... python bench.py --steps 2 --warmup 1 --no-cli --no-cpu-baseline --no-e2e',
       '# (tools/profile_round.sh; 3 steps of 125 000 queries: 1 warm-up + 2 timed; per-launch times are cold-cache and serialised: shares, not absolutes)', '',
       '## per kernel', '%-66s %5s %10s %7s' % ('kernel', 'n', 'total ms', 'share')]
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append('%-66s %5d %10.3f %6.1f%%' % (k[:66], a[0], a[1], 100 * a[1] / tot))
out += ['%-66s %5s %10.3f' % ('all', '', tot), '', '## launch list'] + lines
open('profiles/launches_%s.txt' % R, 'w').write('\n'.join(out) + '\n')
PY
for f in bench_full:bench bench_ref:bench_ref bench_protein:bench_protein; do s=${f%%:*}; d=${f##*:}; [ -s gpurun_out/$s.json ] && cp gpurun_out/$s.json profiles/${d}_$R.json; done
echo done
