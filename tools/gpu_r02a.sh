#!/bin/bash
# r02a: evidence for the 3d large-grid (global) variant: ncu --set full of C5-32, interleave on/off sweep, reference CUDA kernel timed at C5-32
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for IL in 1 0; do
echo "### INTERLEAVE=$IL C5-32"; NUFI_B200_INTERLEAVE=$IL timeout 600 python tools/sweep.py C5-32 --W 16 12 8 --reps 2
echo "### INTERLEAVE=$IL C5-16 global"; NUFI_B200_INTERLEAVE=$IL timeout 300 python tools/sweep.py C5-16 --W 15 --reps 3
done > gpurun_out/r02a_sweep.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backtrace_kernel -s 29 -c 1 -f -o gpurun_out/prof_C5-32 python bench.py --workload C5-32 --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_full_C5-32.log 2>&1
ncu -i gpurun_out/prof_C5-32.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/raw_C5-32.csv.gz
ncu -i gpurun_out/prof_C5-32.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/src_C5-32.csv.gz
rm -f gpurun_out/prof_C5-32.ncu-rep
timeout 600 python - > gpurun_out/r02a_refcuda_C5-32.txt 2>&1 <<'PY'
import numpy as np, torch, json
from bench import make_workload, GpuRunner, free_run, reference_cuda_leg
from numericalflowiteration_b200 import stride_t
conf, f0, d, desc = make_workload("C5-32", 1)
torch.cuda.set_device(0)
r = GpuRunner(conf, f0, 0, 1, torch, None)
free_run(r, d)
st = stride_t(conf)
hist = np.zeros((d + 1) * st)
for l in range(d):
    hist[l*st:(l+1)*st] = r.s.download_phi(l)
print(json.dumps(reference_cuda_leg(conf, f0, d, hist, r.s, 0, reps=1)))
PY
ls -la gpurun_out | tail -20
