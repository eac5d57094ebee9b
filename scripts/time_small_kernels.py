"""Scratch timing of the small kernels around the fused path (torch.profiler kernel times, not the bench): the training
backward at a rank's shard shape and the global-cosine backward.  usage: python scripts/time_small_kernels.py [B_img] [B_cap]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from gloria_nlp_project_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Bc = int(sys.argv[2]) if len(sys.argv) > 2 else 64
L = 97
lib = _lib.lib()
gen = torch.Generator(device="cuda").manual_seed(0)
ctx = torch.randn(B, 768, 361, device="cuda", generator=gen)
words = torch.randn(Bc, 768, 97, device="cuda", generator=gen)
lens = torch.full((Bc,), L, dtype=torch.int32, device="cuda")
pk = ops.tc_prepack(ctx, words, lens, L, 0)
sim = torch.empty(B, Bc, device="cuda")
st = torch.cuda.current_stream().cuda_stream
n = lib.gloria_b200_tc_train_workspace(B, Bc, 768, 361, L)
tws = torch.empty(n, dtype=torch.uint8, device="cuda")
dsim = torch.randn(B, Bc, device="cuda", generator=gen) * 0.01
d_ctx, d_words = torch.empty_like(ctx), torch.empty_like(words)
x = torch.randn(B, 768, device="cuda", generator=gen)
y = torch.randn(Bc, 768, device="cuda", generator=gen)
dcos = torch.randn(B, Bc, device="cuda", generator=gen)


def step():
    rc = lib.gloria_b200_tc_local_sim_fwd_train(pk.ctx_h.data_ptr(), pk.ctx_t.data_ptr(), pk.words_h.data_ptr(), pk.wnorm.data_ptr(),
                                                lens.data_ptr(), B, Bc, 768, 361, L, 4.0, 5.0, 0, 1e-8, sim.data_ptr(), tws.data_ptr(), n, st)
    assert rc == 0, lib.gloria_b200_last_error()
    rc = lib.gloria_b200_tc_local_sim_bwd_train(pk.ctx_t.data_ptr(), pk.words_t.data_ptr(), lens.data_ptr(), B, Bc, 768, 361, 97, L, 0,
                                                dsim.data_ptr(), d_ctx.data_ptr(), d_words.data_ptr(), tws.data_ptr(), n, st)
    assert rc == 0, lib.gloria_b200_last_error()
    cosm, xn, yn = ops.global_sim_fwd(x, y, 1e-8)
    ops.global_sim_bwd(x, y, xn, yn, dcos, 1e-8)


step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(f"# B_img={B} B_cap={Bc}: kernel times, average of 3 (us)")
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if ev.device_time_total > 0:
        print(f"{ev.device_time_total / ev.count:10.1f}  x{ev.count // 3}  {ev.key[:110]}")
