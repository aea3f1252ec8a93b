#!/bin/bash
# r02h (gpurun --gpus N): peer exchange v3 (self-validating words): multi-GPU parity tests, then C2 weak bench at N (no extras), then with extras
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_drivers_gpu.py tests/test_gpu_parity.py -m gpu -q -k "group or peer or torchrun" > gpurun_out/r02h_pytest_multi_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_pytest_multi_n$N.log; tail -5 gpurun_out/r02h_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/r02h_bench_n$N.err; python tools/show_bench.py gpurun_out/r02h_bench_n$N.json
timeout 300 python bench.py --steps 50 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; python tools/show_bench.py gpurun_out/r02h_bench_n1.json
