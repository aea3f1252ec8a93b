#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/_fillcheck.py > gpurun_out/r02n_fillcheck.txt 2>&1; cat gpurun_out/r02n_fillcheck.txt
