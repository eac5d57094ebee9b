"""Scratch timing of the fused training path through the C ABI (not the bench).  usage: python scripts/time_train.py B_img [L] [B_cap]  (B_cap < B_img = the shard shape of one rank of a caption-sharded step)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gloria_nlp_project_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = int(sys.argv[2]) if len(sys.argv) > 2 else 97
Bc = int(sys.argv[3]) if len(sys.argv) > 3 else B
n_it = int(os.environ.get("N_IT", "3"))
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


def fwd():
    rc = lib.gloria_b200_tc_local_sim_fwd_train(pk.ctx_h.data_ptr(), pk.ctx_t.data_ptr(), pk.words_h.data_ptr(), pk.wnorm.data_ptr(),
                                                lens.data_ptr(), B, Bc, 768, 361, L, 4.0, 5.0, 0, 1e-8, sim.data_ptr(), tws.data_ptr(), n, st)
    assert rc == 0, lib.gloria_b200_last_error()


def bwd():
    rc = lib.gloria_b200_tc_local_sim_bwd_train(pk.ctx_t.data_ptr(), pk.words_t.data_ptr(), lens.data_ptr(), B, Bc, 768, 361, 97, L, 0,
                                                dsim.data_ptr(), d_ctx.data_ptr(), d_words.data_ptr(), tws.data_ptr(), n, st)
    assert rc == 0, lib.gloria_b200_last_error()


def timeit(fn):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n_it):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n_it


tf, tb = timeit(fwd), timeit(bwd)
flops = 12 * 361 * 768 * B * Bc * L
print(f"B_img={B} B_cap={Bc} L={L}: fused train fwd {tf:.3f} ms, bwd (scale + GEMMs) {tb:.3f} ms -> {flops / (tf + tb) / 1e9:.1f} TFLOP/s algorithmic; workspace {n / 1e9:.2f} GB")
