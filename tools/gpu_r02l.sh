#!/bin/bash
# r02l (1 GPU): full GPU suite on the self-validating-slot path, microbenchmarks (DFMA operand bandwidth), ring-shape sweeps, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest_gpu.log; tail -4 gpurun_out/r02l_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02l_smoke.log; tail -2 gpurun_out/r02l_smoke.log
./tools/build/microbench > gpurun_out/r02l_microbench.txt 2>&1; grep -i "DFMA" gpurun_out/r02l_microbench.txt
for W in C1 C2; do timeout 300 python tools/sweep.py $W --lc 0 8 4 2 --reps 10 2>&1 | tail -5; done > gpurun_out/r02l_sweep_1d.txt; cat gpurun_out/r02l_sweep_1d.txt
timeout 300 python tools/sweep.py C5-16 --lc 0 1 --reps 5 2>&1 | tail -3 > gpurun_out/r02l_sweep_3d.txt; cat gpurun_out/r02l_sweep_3d.txt
for T in 1024 512 256; do NUFI_B200_TAIL_THREADS=$T timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02l_bench_tail$T.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02l_bench_tail$T.json; done
python - <<'P'
import json
d=json.load(open('gpurun_out/r02l_bench_tail1024.json')); print(d['step_ms_rank0'])
P
