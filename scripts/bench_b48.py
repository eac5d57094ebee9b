"""BASELINE.json configs[1] (B = 48 loss step, bf16 mode) three ways: host-list caption lengths (eager), device-side
caption lengths (eager, no host round trip), and the same step captured in a CUDA graph and replayed.
usage: python scripts/bench_b48.py [steps]   -> one JSON line"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200 import _lib, gloria_loss

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B, D, H, W, LW = 48, 768, 19, 19, 97
G.set_precision("bf16")
lib = _lib.lib()
gen = torch.Generator(device="cuda").manual_seed(0)
t = {k: torch.randn(s, device="cuda", generator=gen).requires_grad_(True)
     for k, s in (("img_l", (B, D, H, W)), ("txt_l", (B, D, LW)), ("img_g", (B, D)), ("txt_g", (B, D)))}
leaves = tuple(t.values())
host_lens = [LW] * B
dev_lens = gloria_loss.DeviceCapLens(torch.tensor(host_lens, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step(lens):
    l0, l1, _, _, _, _ = gloria_loss.local_loss(t["img_l"], t["txt_l"], lens)
    g0, g1 = gloria_loss.global_loss(t["img_g"], t["txt_g"])
    loss = l0 + l1 + g0 + g1
    return loss, torch.autograd.grad(loss, leaves)


def timed(fn):
    for i in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(steps):
        flush.fill_(i & 1)                      # inputs (53 MB) are smaller than L2: evict them between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / steps


out = {"workload": "b48: batch 48, 97 words, 361 regions, D=768, local+global loss fwd+bwd, bf16 mode", "steps": steps}
out["eager_host_lens_ms"] = timed(lambda: step(host_lens))
lib.gloria_b200_launch_count(1)
step(dev_lens)
out["own_launches_per_step"] = int(lib.gloria_b200_launch_count(1))
out["eager_device_lens_ms"] = timed(lambda: step(dev_lens))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        step(dev_lens)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    loss_g, grads_g = step(dev_lens)
out["cuda_graph_replay_ms"] = timed(graph.replay)
out["loss"] = float(loss_g)
f_step = 12.0 * 361 * D * B * B * LW
out["tflops_algorithmic_graph"] = f_step / out["cuda_graph_replay_ms"] / 1e9
print(json.dumps(out))
