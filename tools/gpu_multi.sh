#!/bin/bash
# multi-GPU visit (gpurun --gpus N): one-process group + torchrun IPC tests, then bench at N with both rho exchanges
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt; nvidia-smi topo -m >> gpurun_out/smi_multi.txt 2>&1
timeout 900 python -m pytest tests/test_drivers_gpu.py -m gpu -x -q -k "group or peer or torchrun" > gpurun_out/pytest_multi.log 2>&1; tail -5 gpurun_out/pytest_multi.log
for W in ${WORKLOADS:-C2 C3 C5-16}; do
for X in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $W --exchange $X --steps ${STEPS:-50} --warmup 5 > gpurun_out/bench_${W}_n${N}_$X.json 2> gpurun_out/bench_${W}_n${N}_$X.err; tail -2 gpurun_out/bench_${W}_n${N}_$X.err; cut -c1-160 gpurun_out/bench_${W}_n${N}_$X.json
done
done
