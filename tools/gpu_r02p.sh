#!/bin/bash
# r02p (1 GPU): direct-sum Stockham stages in the fused tail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02p_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest_gpu.log; tail -4 gpurun_out/r02p_pytest_gpu.log
for W in C2 C1 C4 C3; do
timeout 300 python bench.py --workload $W --steps 50 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02p_bench_$W.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02p_bench_$W.json
done
NUFI_B200_LIB=$PWD/numericalflowiteration_b200/lib_tt/libnufi_b200.so timeout 300 python tools/_tailtime.py > gpurun_out/r02p_tailtime.txt 2>&1; tail -12 gpurun_out/r02p_tailtime.txt | cut -c1-300
