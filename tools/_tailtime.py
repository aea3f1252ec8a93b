# tail phase timings (needs the -DNUFI_TAIL_TIMING build: NUFI_B200_LIB=numericalflowiteration_b200/lib_tt/libnufi_b200.so)
import sys; sys.path.insert(0,'.')
import torch
from bench import make_workload, GpuRunner, free_run
for w,n in (("C2",6),("C1",6),("C4",6),("C3",4)):
    conf,f0,d,desc=make_workload(w,1)
    r=GpuRunner(conf,f0,0,1,torch,None)
    print(desc, flush=True)
    free_run(r,n)
    r.s.sync()
