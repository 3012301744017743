import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from airpollution_b200 import crbe, workloads
from airpollution_b200.distributed import PartitionedCRBE
steps = 6
wl = workloads.unit_square(48, steps=steps, regime="P-T10", ny=96)
part = PartitionedCRBE(wl, device=dev)
for _ in range(steps): part.step()
sol = part.gather_solution()
if rank == 0:
    print("partitioned:", part.step_info)
    md = crbe.MeshData(wl.mesh(), wl.domain(), wl.nt, device=dev)
    for tma in (True, False):
        s = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False, tma=tma)
        ref = s.solve()[-1]
        print("single tma=%s:" % tma, s.step_info, "rel", np.linalg.norm(sol-ref)/np.linalg.norm(ref))
part.close()
dist.barrier(); dist.destroy_process_group()
