"""Scratch timing of the tensor-core backward pieces (not the bench).  usage: python scripts/time_bwd.py B [L]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gloria_nlp_project_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = int(sys.argv[2]) if len(sys.argv) > 2 else 97
n_it = int(os.environ.get("N_IT", "3"))
lib = _lib.lib()
gen = torch.Generator(device="cuda").manual_seed(0)
ctx = torch.randn(B, 768, 361, device="cuda", generator=gen)
words = torch.randn(B, 768, 97, device="cuda", generator=gen)
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
packed = ops.tc_prepack(ctx, words, lens, L, 0)
lpad = lib.gloria_b200_tc_lpad(L)
sim = torch.empty(B, B, device="cuda")
stats = torch.empty(B, B, 2, lpad, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rc = lib.gloria_b200_tc_local_sim_fwd(packed.ctx_h.data_ptr(), packed.ctx_n.data_ptr(), packed.words_h.data_ptr(), packed.wnorm.data_ptr(),
                                      lens.data_ptr(), B, B, 768, 361, L, 4.0, 5.0, 0, 1e-8, sim.data_ptr(), stats.data_ptr(), st)
assert rc == 0
dsim = torch.randn(B, B, device="cuda", generator=gen) * 0.01
nbytes = lib.gloria_b200_tc_bwd_workspace(B, B, 768, 361, L, 1, 0)
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_ctx, d_words = torch.empty_like(ctx), torch.empty_like(words)
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(3)]
for a, b in evs:
    a.record(); b.record()
lib.gloria_b200_set_timer_events(1, evs[1][0].cuda_event, evs[1][1].cuda_event)
lib.gloria_b200_set_timer_events(2, evs[2][0].cuda_event, evs[2][1].cuda_event)


def bwd():
    rc = lib.gloria_b200_tc_local_sim_bwd(packed.ctx_h.data_ptr(), packed.ctx_t.data_ptr(), packed.ctx_n.data_ptr(), packed.words_h.data_ptr(), packed.words_t.data_ptr(), packed.wnorm.data_ptr(),
                                          lens.data_ptr(), stats.data_ptr(), B, B, 768, 361, 97, L, 0, 4.0, 5.0, 0, 1e-8,
                                          dsim.data_ptr(), None, d_ctx.data_ptr(), d_words.data_ptr(), ws.data_ptr(), nbytes, st)
    assert rc == 0, lib.gloria_b200_last_error()


for _ in range(2):
    bwd()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(n_it):
    bwd()
e.record()
torch.cuda.synchronize()
t = s.elapsed_time(e) / n_it
flops = 8 * 361 * 768 * B * B * L
print(f"B={B} L={L}: tc bwd {t:.3f} ms (pair kernel {evs[1][0].elapsed_time(evs[1][1]):.3f} ms, gemms {evs[2][0].elapsed_time(evs[2][1]):.3f} ms)"
      f" -> {flops / t / 1e9:.1f} TFLOP/s algorithmic; workspace {nbytes / 1e9:.2f} GB")
