#!/bin/bash
# one GPU-box visit: parity tests, smoke, bench, reference arm, ncu launch list, ncu full captures (evidence for profiles/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nproc >> gpurun_out/lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# skip the 800 history-building steps (2 launches each); list the launches of the warm-up and timed steps at depth 800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 200 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launch_bench.log 2>&1
for W in "C2 800" "C3 100" "C4 50" "C5-16 25"; do set -- $W
# skip the launches of the history build (n per workload) and of the first timed region; capture one steady-state backtrace launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backtrace_kernel -s $(($2+4)) -c 1 -f -o gpurun_out/prof_$1 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_full_$1.log 2>&1
# the reports are ~17 MB each and gpurun_out is capped at 64 MiB: keep the raw and source pages as gzipped CSV, drop the report
ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/raw_$1.csv.gz
ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/src_$1.csv.gz
rm -f gpurun_out/prof_$1.ncu-rep
done
[ -x tools/build/microbench ] || { mkdir -p tools/build; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/microbench tools/microbench.cu; }
./tools/build/microbench > gpurun_out/microbench.txt 2>&1
ls -la gpurun_out
