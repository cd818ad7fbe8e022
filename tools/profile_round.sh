#!/bin/bash
# Profile captures of one round (run on the GPU box through gpurun); outputs under gpurun_out/, summarised into profiles/ afterwards.
# Numbers printed by bench.py under ncu are never bench values.
B="python bench.py --steps 2 --warmup 1 --no-cli --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:dense_tc|tc_image|select_|place_|bin_classes|finalize|pack_|gather' --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 7 -c 1 -o gpurun_out/dense_tc -f $B > gpurun_out/ncu_tc.log 2>&1
# one whole step of selection launches (first pass x6 + rerun) and of placement launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_nuc_kernel -s 7 -c 7 -o gpurun_out/select -f $B > gpurun_out/ncu_select.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:place_ -s 6 -c 6 -o gpurun_out/place -f $B > gpurun_out/ncu_place.log 2>&1
timeout 600 python bench.py --workload protein --no-cli > gpurun_out/bench_protein.json 2> gpurun_out/bench_protein.err
ls -la gpurun_out/
