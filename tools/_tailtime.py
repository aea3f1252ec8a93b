# tail phase timings (needs the -DNUFI_TAIL_TIMING build: NUFI_B200_LIB=numericalflowiteration_b200/lib_tt/libnufi_b200.so)
import sys; sys.path.insert(0,'.')
import os
import torch
from bench import make_workload, GpuRunner, free_run
for w,n in (("C2",4),("C4",4),("C3",3)):
    for thr in (1024, 512, 256):
        os.environ["NUFI_B200_TAIL_THREADS"] = str(thr)
        conf,f0,d,desc=make_workload(w,1)
        r=GpuRunner(conf,f0,0,1,torch,None)
        free_run(r,n)
        r.s.sync()
        print("^^", w, "threads", thr, flush=True)
        r.s.close()
