#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r02w_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w_pytest_gpu.log; tail -3 gpurun_out/r02w_pytest_gpu.log
for W in C2 C1 C4 C3; do timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02w_bench_$W.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02w_bench_$W.json; done
NUFI_B200_LIB=$PWD/numericalflowiteration_b200/lib_tt/libnufi_b200.so timeout 200 python tools/_tailtime.py 2>&1 | grep -A3 "C2 threads 1024\|C4 threads 1024" | head; NUFI_B200_LIB=$PWD/numericalflowiteration_b200/lib_tt/libnufi_b200.so timeout 200 python tools/_tailtime.py > gpurun_out/r02w_tailtime.txt 2>&1; grep -B2 "threads 1024" gpurun_out/r02w_tailtime.txt | cut -c1-260
