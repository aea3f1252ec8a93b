#!/bin/bash
mkdir -p gpurun_out
NUFI_B200_LIB=$PWD/numericalflowiteration_b200/lib_tt/libnufi_b200.so timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/_peertime.py > gpurun_out/r02_peertime_full.log 2>&1
grep -A40 "history of C2" gpurun_out/r02_peertime_full.log | head -60 > gpurun_out/r02_peertime.log; grep -A30 "history of C4" gpurun_out/r02_peertime_full.log | head -40 >> gpurun_out/r02_peertime.log
rm -f gpurun_out/r02_peertime_full.log
cat gpurun_out/r02_peertime.log
