#!/bin/bash
# compute-sanitizer passes of one round on the GPU box (through gpurun); output gpurun_out/sanitizer.txt
O=gpurun_out/sanitizer.txt
S=/usr/local/cuda/bin/compute-sanitizer
K1="small_FM_MLSE_pos or c2_matrix_BME or syn300_BME_HYBRID or c1_align_special or tiny_tree or device_packer or counts or tensor_core or symbols_outside"
{
echo "# compute-sanitizer on the B200 box (r02 kernels: tensor-core count kernel, flat-list selection, shared-memory placement, side-stream reruns)"
echo "## memcheck: tests/test_gpu_parity.py -k '$K1'"
timeout 900 $S --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K1" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*error" | tail -6
echo "## memcheck: -k 'larger_than_the_member_list or properties_at_scale' (member list filled in rounds, overflow reruns on the side stream, sub-batches)"
timeout 1200 $S --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "larger_than_the_member_list or properties_at_scale" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | tail -6
echo "## racecheck (shared memory: selection queue / member list / sort keys, placement working set, tensor-core staging): -k 'syn300_FM_MLSE_pos and golden or larger_than_the_member_list'"
timeout 1200 $S --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "syn300_FM_MLSE_pos and golden or larger_than_the_member_list" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -6
echo "## synccheck: -k 'syn300_FM_MLSE_pos and golden'"
timeout 900 $S --tool synccheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "syn300_FM_MLSE_pos and golden" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -4
} > $O 2>&1
cat $O
