#!/bin/bash
# r02c (gpurun --gpus 2): multi-GPU parity tests, then the default bench under torchrun at N = 2 (C2 weak + 3d strong/weak extras)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt; nvidia-smi topo -m >> gpurun_out/smi_multi.txt 2>&1; nproc >> gpurun_out/smi_multi.txt
timeout 900 python -m pytest tests/test_drivers_gpu.py tests/test_gpu_parity.py -m gpu -q -k "group or peer or torchrun" > gpurun_out/r02_pytest_multi_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multi_n$N.log; tail -5 gpurun_out/r02_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err; cut -c1-300 gpurun_out/r02_bench_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_ref_n$N.json
