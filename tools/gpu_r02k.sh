#!/bin/bash
# r02k (gpurun --gpus N): where does the N > 1 step time go -- rank alignment after the L2 flush / no flush at all
N=${1:-2}
mkdir -p gpurun_out
for V in "align:" "noalign:--no-align" "noflush:--no-flush --no-align"; do
T=${V%%:*}; F=${V#*:}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload C2 --steps 100 --warmup 5 --no-extras --no-cpu $F > gpurun_out/r02k_C2_n${N}_$T.json 2> gpurun_out/r02k_C2_n${N}_$T.err; echo "$T rc=$?"; python tools/show_bench.py gpurun_out/r02k_C2_n${N}_$T.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload C4 --steps 100 --warmup 5 --no-extras --no-cpu > gpurun_out/r02k_C4_n${N}_align.json 2> gpurun_out/r02k_C4_n${N}_align.err; python tools/show_bench.py gpurun_out/r02k_C4_n${N}_align.json
timeout 300 python bench.py --steps 100 --warmup 5 --no-extras --no-full-run --no-cpu --no-flush > gpurun_out/r02k_C2_n1_noflush.json 2> gpurun_out/r02k_C2_n1_noflush.err; python tools/show_bench.py gpurun_out/r02k_C2_n1_noflush.json
