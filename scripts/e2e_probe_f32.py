"""Development probe: end-to-end B = 48 fp32-mode step with and without the per-step pageable copies (found the 10 ms copy-engine stall fixed by gloria_b200_upload_ints)."""
import os, sys, time
sys.path.insert(0, ".")
import torch
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200 import gloria_loss
G.set_precision("fp32")
B, D, H, W, LW = 48, 768, 19, 19, 97
dev = torch.device("cuda", 0)
host = {k: torch.randn(s).pin_memory() for k, s in (("img_l", (B, D, H, W)), ("txt_l", (B, D, LW)), ("img_g", (B, D)), ("txt_g", (B, D)))}
lens = gloria_loss.DeviceCapLens(torch.full((B,), LW, dtype=torch.int32, device=dev))
copy_stream = torch.cuda.Stream(device=dev)
def issue():
    with torch.cuda.stream(copy_stream):
        t = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        ev = torch.cuda.Event(); ev.record(copy_stream)
    return t, ev
def step(cur, rs=True):
    t, ev = cur
    nxt = issue()
    torch.cuda.current_stream().wait_event(ev)
    for v in t.values():
        if rs: v.record_stream(torch.cuda.current_stream())
        v.requires_grad_(True)
    l0, l1, *_ = gloria_loss.local_loss(t["img_l"], t["txt_l"], lens)
    g0, g1 = gloria_loss.global_loss(t["img_g"], t["txt_g"])
    (l0 + l1 + g0 + g1).backward()
    return nxt
for rs in (True, False):
    cur = issue()
    for _ in range(3): cur = step(cur, rs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): cur = step(cur, rs)
    torch.cuda.synchronize()
    print("record_stream", rs, (time.perf_counter() - t0) * 100, "ms/step", "reserved GB", torch.cuda.memory_reserved() / 1e9, "mallocs", torch.cuda.memory_stats()["num_device_alloc"])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): cur = step(cur, True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=50))
