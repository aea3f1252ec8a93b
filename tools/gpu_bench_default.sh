#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r02x_bench.json; tail -3 gpurun_out/r02x_bench.err
