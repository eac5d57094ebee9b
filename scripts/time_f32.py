"""fp32 mode: tensor-core path (tc_f32.cu) vs CUDA-core path (simt_f32.cu), cfg2 (B = 48 step) and cfg4 (zero-shot
10 000 x 25, forward only), plus the kernel table of one tensor-core step.
usage: python scripts/time_f32.py [B] [--profile] [--zs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200 import gloria_loss, ops

D, H, W, LW = 768, 19, 19, 97
dev = torch.device("cuda", 0)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 48
G.set_precision("fp32")


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


g = torch.Generator(device="cuda").manual_seed(0)
img = torch.randn(B, D, H, W, device=dev, generator=g).requires_grad_(True)
txt = torch.randn(B, D, LW, device=dev, generator=g).requires_grad_(True)
xg = torch.randn(B, D, device=dev, generator=g).requires_grad_(True)
yg = torch.randn(B, D, device=dev, generator=g).requires_grad_(True)
lens = [LW] * B


def step():
    for v in (img, txt, xg, yg):
        v.grad = None
    l0, l1, *_ = gloria_loss.local_loss(img, txt, lens)
    g0, g1 = gloria_loss.global_loss(xg, yg)
    (l0 + l1 + g0 + g1).backward()


def fwd_only():
    with torch.no_grad():
        gloria_loss.local_similarities(img, txt, lens)


out = {"B": B}
fl = 12.0 * H * W * D * B * sum(lens)
for name, flag in (("tensor_core", True), ("cuda_core", False)):
    ops._F32_TC = flag
    if not flag and B > 64:
        continue
    ms = timed(step, 10 if flag else 3, 3 if flag else 1)
    msf = timed(fwd_only, 10 if flag else 3, 2 if flag else 1)
    out[name] = {"ms_per_step": ms, "ms_forward_only": msf, "pairs_per_s": B / ms * 1e3, "tflops_fp32_equivalent": fl / ms / 1e9}
print(json.dumps(out), flush=True)

if "--profile" in sys.argv:
    ops._F32_TC = True
    from torch.profiler import profile, ProfilerActivity
    step(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))

if "--zs" in sys.argv:
    from gloria_nlp_project_b200.gloria_model import GLoRIALossMixin

    class Model(GLoRIALossMixin):
        temp1, temp2, temp3 = 4.0, 5.0, 10.0
    m = Model()
    N_IMG, N_TXT, LT = 10000, 25, 18
    gz = torch.Generator(device="cuda").manual_seed(3)
    zi = torch.randn(N_IMG, D, H, W, device=dev, generator=gz)
    zt = torch.randn(N_TXT, D, LT, device=dev, generator=gz)
    cl = torch.randint(4, 17, (N_TXT,), generator=torch.Generator().manual_seed(4)).tolist()
    res = {}
    for name, flag in (("tensor_core", True), ("cuda_core", False)):
        ops._F32_TC = flag
        res[name] = timed(lambda: m.get_local_similarities(zi, zt, cl), 3, 1)
    print(json.dumps({"cfg4_fp32_local_ms": res}), flush=True)
