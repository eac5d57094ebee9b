"""Device-time breakdown of the caption-sharded step (rank 0's view).  Launch with torch.distributed.run:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/profile_sharded.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200 import distributed
G.set_precision("bf16")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = B // world
gen = torch.Generator(device="cuda").manual_seed(rank)
t = [torch.randn(s, device="cuda", generator=gen, requires_grad=True) for s in ((n, 768, 19, 19), (n, 768, 97), (n, 768), (n, 768))]
lens = [97] * n


def step():
    for v in t:
        v.grad = None
    l = sum(distributed.sharded_loss(*t, lens))
    l.backward()


for _ in range(4):
    step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"N={world} B={B}: {e0.elapsed_time(e1) / 10:.3f} ms/step")
dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=32, max_name_column_width=60))
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    # timeline of the last step: kernel name, start offset, duration (us)
    t_end = ev[-1].time_range.end
    last = [e for e in ev if e.time_range.start > t_end - 16000]
    t0 = last[0].time_range.start
    for e in last:
        print(f"{e.time_range.start - t0:9.0f} {e.time_range.end - e.time_range.start:8.0f}  {e.name[:90]}")
dist.destroy_process_group()
