"""Where a small-batch step spends its time: torch.profiler over the public API (host ops + device kernels).
usage: python scripts/profile_step.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200 import gloria_loss
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
G.set_precision("bf16")
gen = torch.Generator(device="cuda").manual_seed(0)
img = torch.randn(B, 768, 19, 19, device="cuda", generator=gen, requires_grad=True)
txt = torch.randn(B, 768, 97, device="cuda", generator=gen, requires_grad=True)
ig = torch.randn(B, 768, device="cuda", generator=gen, requires_grad=True)
tg = torch.randn(B, 768, device="cuda", generator=gen, requires_grad=True)
lens = [97] * B


def step():
    for t in (img, txt, ig, tg):
        t.grad = None
    l0, l1, *_ = gloria_loss.local_loss(img, txt, lens)
    g0, g1 = gloria_loss.global_loss(ig, tg)
    (l0 + l1 + g0 + g1).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20):
    step()
t_host = (time.perf_counter() - t0) / 20
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 20
print(f"B={B}: host issue time {t_host * 1e3:.3f} ms/step, wall {t_all * 1e3:.3f} ms/step")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=70))
