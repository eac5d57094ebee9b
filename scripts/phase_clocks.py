"""Development aid: per-CTA wait-time breakdown of the MMA issuer threads (library built with -DGLORIA_PHASE_CLOCKS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gloria_nlp_project_b200 import ops, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
L = 97
lib = _lib.lib()
gen = torch.Generator(device="cuda").manual_seed(0)
ctx = torch.randn(B, 768, 361, device="cuda", generator=gen)
words = torch.randn(B, 768, 97, device="cuda", generator=gen)
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
packed = ops.tc_prepack(ctx, words, lens, L, 0)
sim = torch.empty(B, B, device="cuda"); stats = torch.empty(B, B, 2, 112, device="cuda")
st = torch.cuda.current_stream().cuda_stream
dbg = torch.zeros(148, 32, dtype=torch.int64, device="cuda")
lib.gloria_b200_debug_phase_clocks(dbg.data_ptr())
def fwd():
    assert lib.gloria_b200_tc_local_sim_fwd(packed.ctx_h.data_ptr(), packed.ctx_n.data_ptr(), packed.words_h.data_ptr(), packed.wnorm.data_ptr(),
        lens.data_ptr(), B, B, 768, 361, L, 4.0, 5.0, 0, 1e-8, sim.data_ptr(), stats.data_ptr(), st) == 0
fwd(); fwd(); torch.cuda.synchronize()
d = dbg.cpu().double(); act = d[:, 5] > 0
m = d[act].mean(0)
print(f"fwd  B={B}: per pair cycles total {m[0]/m[5]:.0f}: wait full(TMA) {m[1]/m[5]:.0f}, D1 free {m[2]/m[5]:.0f}, E full {m[3]/m[5]:.0f}, D2 free {m[4]/m[5]:.0f}  (pairs/CTA {m[5]:.0f})")
dsim = torch.randn(B, B, device="cuda", generator=gen) * 0.01
nbytes = lib.gloria_b200_tc_bwd_workspace(B, B, 768, 361, L, 1, 0)
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_ctx, d_words = torch.empty_like(ctx), torch.empty_like(words)
dbg.zero_()
for _ in range(2):
    assert lib.gloria_b200_tc_local_sim_bwd(packed.ctx_h.data_ptr(), packed.ctx_t.data_ptr(), packed.ctx_n.data_ptr(), packed.words_h.data_ptr(), packed.words_t.data_ptr(), packed.wnorm.data_ptr(),
        lens.data_ptr(), stats.data_ptr(), B, B, 768, 361, 97, L, 0, 4.0, 5.0, 0, 1e-8, dsim.data_ptr(), None, d_ctx.data_ptr(),
        d_words.data_ptr(), ws.data_ptr(), nbytes, st) == 0
torch.cuda.synchronize()
d = dbg.cpu().double(); act = d[:, 5] > 0
m = d[act].mean(0)
n = m[5]
print(f"bwd  B={B}: per pair cycles: MMA issuer total {m[0]/n:.0f}: wait full(TMA) {m[1]/n:.0f}, S_ free {m[2]/n:.0f}, E full {m[3]/n:.0f}, T' free {m[4]/n:.0f}  (pairs/CTA {n:.0f})")
print(f"      producer total {m[8]/n:.0f}: wait empty slot {m[9]/n:.0f}")
for g in range(3):
    o = 16 + 5 * g
    print(f"      SIMT group {g} warp: total {m[o]/n:.0f}: wait S_ ready {m[o+1]/n:.0f}, T' ready {m[o+2]/n:.0f}, E free {m[o+3]/n:.0f}, bar.sync exchange {m[o+4]/n:.0f}")

# fused training kernel
n = lib.gloria_b200_tc_train_workspace(B, B, 768, 361, L)
tws = torch.empty(n, dtype=torch.uint8, device="cuda")
dbg.zero_()
for _ in range(2):
    assert lib.gloria_b200_tc_local_sim_fwd_train(packed.ctx_h.data_ptr(), packed.ctx_t.data_ptr(), packed.words_h.data_ptr(),
        packed.wnorm.data_ptr(), lens.data_ptr(), B, B, 768, 361, L, 4.0, 5.0, 0, 1e-8, sim.data_ptr(), tws.data_ptr(), n, st) == 0
torch.cuda.synchronize()
d = dbg.cpu().double(); act = d[:, 5] > 0
m = d[act].mean(0); n_ = m[5]
print(f"fused B={B}: per pair cycles: MMA issuer total {m[0]/n_:.0f}: wait full(TMA) {m[1]/n_:.0f}, S_ free {m[2]/n_:.0f}, E full {m[3]/n_:.0f}, T' free {m[4]/n_:.0f}  (pairs/CTA {n_:.0f})")
print(f"      producer total {m[8]/n_:.0f}: wait empty slot {m[9]/n_:.0f}")
for g in range(3):
    o = 16 + 5 * g
    print(f"      SIMT group {g} warp: total {m[o]/n_:.0f}: wait S_ ready {m[o+1]/n_:.0f}, T' ready {m[o+2]/n_:.0f}, E free {m[o+3]/n_:.0f}, bar.sync exchange {m[o+4]/n_:.0f}")
