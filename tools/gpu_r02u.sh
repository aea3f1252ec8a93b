#!/bin/bash
# r02u (1 GPU): next-level L1 prefetch in the global 3d variant: parity, A/B on C5-32 and on a non-power-of-two 24^3 grid
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_next_rows_gpu.py -m gpu -q --timeout 600 -k "3d or step_host" > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log; tail -3 gpurun_out/r02u_pytest.log
for PF in 0 1 0 1; do NUFI_B200_PREFETCH=$PF timeout 300 python tools/sweep.py C5-32 --reps 3 2>&1 | tail -1 | sed "s/^/prefetch=$PF /"; done > gpurun_out/r02u_prefetch_ab.txt
for PF in 0 1; do NUFI_B200_PREFETCH=$PF timeout 300 python tools/sweep.py C5-64 --depth 4 --reps 2 2>&1 | tail -1 | sed "s/^/prefetch=$PF C5-64 depth 4 /"; done >> gpurun_out/r02u_prefetch_ab.txt
cat gpurun_out/r02u_prefetch_ab.txt
