#!/bin/bash
# r02m (1 GPU): backtrace kernel launched programmatically behind the previous step's tail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest_gpu.log; tail -4 gpurun_out/r02m_pytest_gpu.log
for W in C2 C1 C4; do
timeout 300 python bench.py --workload $W --steps 50 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02m_bench_$W.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02m_bench_$W.json
NUFI_B200_PDL=0 timeout 300 python bench.py --workload $W --steps 50 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02m_bench_${W}_nopdl.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02m_bench_${W}_nopdl.json
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/r02m_bench_fullrun.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02m_bench_fullrun.json
