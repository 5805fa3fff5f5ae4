"""Multi-GPU check of the NVLS all-reduce kernel (torchrun): result vs NCCL, and time per all-reduce of the 57 MB bucket."""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200.distributed import FlatGradBucket
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 14_319_409
for ctas in (8, 16, 32):
    p = torch.nn.Parameter(torch.zeros(N, device=dev))
    g = torch.Generator(device=dev).manual_seed(rank)
    p.grad = torch.randn(N, generator=g, device=dev)
    ref = p.grad.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    b = FlatGradBucket([p], device=dev, nvls=True, nvls_ctas=ctas).adopt_grads()
    if not b.uses_nvls:
        if rank == 0:
            print("NVLS unavailable:", b.nvls_error)
        break
    b.start().finish()
    torch.cuda.synchronize()
    err = float((p.grad - ref).abs().max() / ref.abs().max())
    ev = lambda: torch.cuda.Event(enable_timing=True)
    a, c = ev(), ev()
    dist.barrier()
    a.record()
    for _ in range(20):
        b.start().finish()
    c.record()
    torch.cuda.synchronize()
    t_nvls = a.elapsed_time(c) / 20
    x = ref.clone()
    a, c = ev(), ev()
    dist.barrier()
    a.record()
    for _ in range(20):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    c.record()
    torch.cuda.synchronize()
    t_nccl = a.elapsed_time(c) / 20
    if rank == 0:
        print(f"world {world} ctas {ctas}: max rel err vs NCCL {err:.2e}; NVLS kernel + 2 barriers {t_nvls:.3f} ms, NCCL all_reduce {t_nccl:.3f} ms ({N * 4 / 1e6:.1f} MB)")
    del b, p
dist.destroy_process_group()
