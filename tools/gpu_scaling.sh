#!/bin/bash
# scaling visit (gpurun --gpus N): bench.py under torchrun at N GPUs for the given workloads and exchanges; JSON lines in gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
for W in ${WORKLOADS:-C2 C3 C5-16}; do
for X in ${EXCHANGES:-peer}; do
if [ "$N" = "1" ]; then
timeout 600 python bench.py --gpus 1 --workload $W --steps ${STEPS:-50} --warmup 5 --no-cpu --no-extras > gpurun_out/scale_${W}_n1.json 2> gpurun_out/scale_${W}_n1.err
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $W --exchange $X --steps ${STEPS:-50} --warmup 5 > gpurun_out/scale_${W}_n${N}_$X.json 2> gpurun_out/scale_${W}_n${N}_$X.err
fi
done
done
ls gpurun_out | grep scale_
