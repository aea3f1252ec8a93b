#!/bin/bash
# multi-GPU visit (gpurun --gpus N): one-process group test + torchrun bench at N and N=1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt
timeout 600 python -m pytest tests/test_drivers_gpu.py -m gpu -x -q -k "group" > gpurun_out/pytest_group.log 2>&1; tail -3 gpurun_out/pytest_group.log
timeout 600 ./bin/build/test_nufi_gpu_3d --landau --steps 20 --fused --quiet --energy gpurun_out/e3d_fused_multi.txt > gpurun_out/driver3d_multi.log 2>&1; tail -2 gpurun_out/driver3d_multi.log
for W in C2 C3 C5-16; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $W --steps 30 --warmup 5 > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err; tail -2 gpurun_out/bench_${W}_n$N.err
timeout 900 python bench.py --gpus 1 --workload $W --steps 30 --warmup 5 --no-extras --no-cpu > gpurun_out/bench_${W}_n1.json 2> gpurun_out/bench_${W}_n1.err
done
