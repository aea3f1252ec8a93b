#!/bin/bash
# r02o (1 GPU): cluster pairs sharing the history stream (multicast halves).  Every command under its own timeout: a protocol bug hangs.
mkdir -p gpurun_out
timeout 120 python -c "
import sys; sys.path.insert(0,'.')
from numericalflowiteration_b200 import Config1D, CudaScheduler, F0
s=CudaScheduler(Config1D(Nx=64,Nu=48,Nt=12),F0(1,0.01,0.5),device=0)
for m in range(12): s.step(m)
s.sync(); print('small ok', s.last_variant, s.download_energy(0,12)[-1])
" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 > gpurun_out/r02o_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r02o_pytest_parity.log
for W in C1 C2 C3 C4 C5-16; do
for C in 2 1; do
NUFI_B200_CLUSTER=$C timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02o_bench_${W}_c$C.json 2>gpurun_out/r02o_bench_${W}_c$C.err; python tools/show_bench.py gpurun_out/r02o_bench_${W}_c$C.json 2>&1 | tail -1; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_${W}_c$C.json')); print('   ', d['roofline']['kernel'])"
done
done
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest_gpu.log; tail -4 gpurun_out/r02o_pytest_gpu.log
