#!/bin/bash
# r02q (1 GPU): 1d level layout [(p0,p1)] [p2]: LDS.128 + LDS.64 per point-step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "1d or C1 or C2 or full or defaults or tail or step" > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_pytest.log; tail -4 gpurun_out/r02q_pytest.log
for W in C2 C1; do
timeout 300 python bench.py --workload $W --steps 50 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02q_bench_$W.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02q_bench_$W.json
done
