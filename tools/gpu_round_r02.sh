#!/bin/bash
# r02 evidence visit (1 GPU): parity tests, smoke, bench (+ reference arm), ncu launch list, ncu --set full captures of the five
# backtrace variants the bench runs, microbenchmarks.  Everything lands in gpurun_out/r02f_* ; tools/ncu_summary.py condenses it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02f_smi.txt 2>&1
lscpu | head -20 > gpurun_out/r02f_lscpu.txt; nproc >> gpurun_out/r02f_lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_gpu.log; tail -3 gpurun_out/r02f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02f_smoke.log; tail -2 gpurun_out/r02f_smoke.log
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r02f_bench.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-full-run > gpurun_out/r02f_bench_k20.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02f_bench_k20.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; cut -c1-200 gpurun_out/r02f_bench_ref.json
# skip the 800 history-building steps (2 launches each); list the launches of the warm-up and timed steps at depth 800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 200 --csv --log-file gpurun_out/r02f_launches_C2.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-full-run > gpurun_out/r02f_ncu_launch_bench.log 2>&1
for W in "C2 800" "C3 100" "C4 50" "C5-16 25" "C5-32 25"; do set -- $W
# skip the launches of the history build (n per workload) and the settle steps; capture one steady-state backtrace launch
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:backtrace_kernel -s $(($2+6)) -c 1 -f -o gpurun_out/prof_$1 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-extras --no-full-run > gpurun_out/r02f_ncu_full_$1.log 2>&1
ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02f_raw_$1.csv.gz
ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02f_src_$1.csv.gz
rm -f gpurun_out/prof_$1.ncu-rep
done
./tools/build/microbench > gpurun_out/r02f_microbench.txt 2>&1
for W in C2 C1; do timeout 300 python tools/sweep.py $W --ilp 1 2 --W 14 15 29 30 --reps 10 2>&1 | tail -9; done > gpurun_out/r02f_sweep_shape_1d.txt 2>&1
ls -la gpurun_out | grep r02f
