#!/bin/bash
# r02j (gpurun --gpus N): peer exchange v5 (self-validating slots, the resident tail reduces + pushes + polls): parity tests,
# C2 / C4 weak bench at N (no extras), phase timers
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_drivers_gpu.py tests/test_gpu_parity.py -m gpu -q -k "group or peer or torchrun" > gpurun_out/r02j_pytest_multi_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_pytest_multi_n$N.log; tail -5 gpurun_out/r02j_pytest_multi_n$N.log
for W in C2 C4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $W --steps 50 --warmup 5 --no-extras > gpurun_out/r02j_bench_${W}_n$N.json 2> gpurun_out/r02j_bench_${W}_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/r02j_bench_${W}_n$N.err; python tools/show_bench.py gpurun_out/r02j_bench_${W}_n$N.json
done
NUFI_B200_LIB=$PWD/numericalflowiteration_b200/lib_tt/libnufi_b200.so timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/_peertime.py > gpurun_out/r02j_peertime_full.log 2>&1
grep -A12 "history of C2" gpurun_out/r02j_peertime_full.log | head -20 > gpurun_out/r02j_peertime_n$N.log; grep -A12 "history of C4" gpurun_out/r02j_peertime_full.log | head -20 >> gpurun_out/r02j_peertime_n$N.log
rm -f gpurun_out/r02j_peertime_full.log
cat gpurun_out/r02j_peertime_n$N.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-extras --no-full-run --no-cpu > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; python tools/show_bench.py gpurun_out/r02j_bench_n1.json
