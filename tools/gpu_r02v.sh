#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/sweep.py C3 --ilp 1 2 --reps 3 2>&1 | tail -3
timeout 300 python tools/sweep.py C3 --ilp 1 --W 20 23 --reps 3 2>&1 | tail -2
timeout 300 python tools/sweep.py C4 --W 15 17 19 --reps 10 2>&1 | tail -3
timeout 300 python tools/sweep.py C5-16 --lc 1 2 --reps 5 2>&1 | tail -2
timeout 300 python tools/sweep.py C5-16 --lc 1 2 --reps 5 2>&1 | tail -2 ) > gpurun_out/r02v_sweeps.txt 2>&1; cat gpurun_out/r02v_sweeps.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02v_bench_k20.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02v_bench_k20.json; python -c "
import json; d=json.load(open('gpurun_out/r02v_bench_k20.json')); print(d['step_ms_rank0'], d['gpu_launches'])"
