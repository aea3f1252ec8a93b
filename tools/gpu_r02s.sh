#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02s_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest_gpu.log; tail -4 gpurun_out/r02s_pytest_gpu.log
timeout 900 python bench.py --steps 50 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r02s_bench.json
python - <<'P'
import json
d=json.load(open('gpurun_out/r02s_bench.json'))
print(d['e2e']); print(d['e2e_driver_sequence']['value']); print(d['other_workloads'][0]['parity'])
P
