"""Timeline of bench.py's END-TO-END loop on the caption-sharded path (pinned host inputs copied in one step ahead, loss read
back asynchronously), rank 0's view, next to the same loop with resident inputs.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/profile_e2e_sharded.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200 import distributed, gloria_loss
G.set_precision("bf16")
B = 512
n = B // world
names = ("img_l", "txt_l", "img_g", "txt_g")
host = {k: torch.randn(s).pin_memory() for k, s in zip(names, ((n, 768, 19, 19), (n, 768, 97), (n, 768), (n, 768)))}
lens = gloria_loss.DeviceCapLens(torch.full((n,), 97, dtype=torch.int32, device=dev))
copy_stream = torch.cuda.Stream(device=dev)
buf = [torch.zeros(1).pin_memory() for _ in range(2)]
evs = [torch.cuda.Event(), torch.cuda.Event()]
state = {"n": 0}


def issue(frac=1.0):
    with torch.cuda.stream(copy_stream):
        if frac >= 1.0:
            t = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        else:      # copy a fraction only (the rest of the tensor is whatever the allocator hands out): bytes-dependence probe
            t = {}
            for k, v in host.items():
                d = torch.empty(v.shape, dtype=v.dtype, device=dev)
                m = max(1, int(v.shape[0] * frac))
                d[:m].copy_(v[:m], non_blocking=True)
                d[m:].normal_()
                t[k] = d
        ev = torch.cuda.Event(); ev.record(copy_stream)
    return t, ev


def step(cur, prefetch=True, late=False, readback=True, frac=1.0):
    t, ev = cur
    nxt = cur
    if prefetch and not late:
        nxt = issue(frac)
    torch.cuda.current_stream().wait_event(ev)
    for v in t.values():
        v.record_stream(torch.cuda.current_stream())
        v.requires_grad_(True)
        v.grad = None
    loss = sum(distributed.sharded_loss(t["img_l"], t["txt_l"], t["img_g"], t["txt_g"], lens))
    if prefetch and late == "mid":
        nxt = issue(frac)                       # after the forward has been enqueued, before the backward
    loss.backward()
    if prefetch and late is True:
        nxt = issue(frac)                       # after the whole step has been enqueued
    if readback:
        s = state["n"] & 1
        buf[s].copy_(loss.detach().reshape(1), non_blocking=True); evs[s].record(); state["n"] += 1
        if state["n"] > 1:
            evs[s ^ 1].synchronize()
    return nxt


def run(prefetch, steps=10, **kw):
    state["n"] = 0
    cur = issue()
    for _ in range(3):
        cur = step(cur, prefetch, **kw)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        cur = step(cur, prefetch, **kw)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) / steps * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), cur


for name, pf, kw in (("inputs resident, no copies", False, {}), ("prefetch at the start of the step", True, {}),
                     ("prefetch enqueued between forward and backward", True, {"late": "mid"}),
                     ("prefetch enqueued after the step", True, {"late": True}),
                     ("prefetch at the start, no loss read-back", True, {"readback": False}),
                     ("prefetch at the start, 1/8 of the bytes", True, {"frac": 0.125})):
    ms, cur = run(pf, **kw)
    if rank == 0:
        print(f"N={world}: {name}: {ms:.3f} ms/step", flush=True)
dist.barrier()
if "--timeline" not in sys.argv:
    dist.destroy_process_group()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        cur = step(cur, True)
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t_end = ev[-1].time_range.end
    last = [e for e in ev if e.time_range.start > t_end - 20000]
    t0 = last[0].time_range.start
    for e in last:
        d = e.time_range.end - e.time_range.start
        if d >= 20:
            print(f"{e.time_range.start - t0:9.0f} {d:8.0f}  {e.name[:100]}")
dist.destroy_process_group()
