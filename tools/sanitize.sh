#!/bin/bash
# compute-sanitizer passes over small runs of every kernel variant (memcheck, racecheck, synccheck); logs in gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
timeout 400 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_drivers_gpu.py tests/test_next_rows_gpu.py -m gpu -x -q -k "teacher_forced or partial_ranges or peer_step_world1 or fused_step or ragged or metrics or multi_period or cluster_pairs or step_host" > gpurun_out/sanitizer_$tool.log 2>&1
echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Race|hazard" gpurun_out/sanitizer_$tool.log | tail -4
done
