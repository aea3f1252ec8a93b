# phase timings of the multi-GPU step (needs the -DNUFI_TAIL_TIMING build: NUFI_B200_LIB=numericalflowiteration_b200/lib_tt/libnufi_b200.so)
#   torchrun --nproc-per-node 2 tools/_peertime.py
import os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from bench import make_workload, GpuRunner, free_run
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for w, n in (("C2", 800), ("C4", 50)):
    conf, f0, d, desc = make_workload(w, world)
    r = GpuRunner(conf, f0, rank, world, torch, dist)
    # build the history quietly: prints come from every step, keep only the last ones
    free_run(r, n)
    dist.barrier()
    if rank == 0:
        print("^^ history of", w, "built; 3 timed-shape steps follow", flush=True)
    for _ in range(3):
        r.step(n)
    r.stream.synchronize()
    dist.barrier()
    r.close()
dist.destroy_process_group()
