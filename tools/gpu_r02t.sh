#!/bin/bash
# r02t (1 GPU): final single-GPU check: full GPU suite, smoke, ncu capture + launch list of the C2 kernel as shipped (1 point per thread)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest_gpu.log; tail -4 gpurun_out/r02t_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02t_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02t_smoke.log; tail -2 gpurun_out/r02t_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 200 --csv --log-file gpurun_out/r02t_launches_C2.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-full-run > gpurun_out/r02t_ncu_launch_bench.log 2>&1
for W in "C2 800" "C1 800"; do set -- $W
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:backtrace_kernel -s $(($2+6)) -c 1 -f -o gpurun_out/prof_$1 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-extras --no-full-run > gpurun_out/r02t_ncu_full_$1.log 2>&1
ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02t_raw_$1.csv.gz
ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02t_src_$1.csv.gz
rm -f gpurun_out/prof_$1.ncu-rep
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-full-run > gpurun_out/r02t_bench_k20.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02t_bench_k20.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02t_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/r02t_bench_ref.json
