#!/bin/bash
# run the commands given as arguments on the GPU box (no tests/bench)
mkdir -p gpurun_out
for cmd in "$@"; do echo "### $cmd"; timeout 900 bash -c "$cmd"; done
