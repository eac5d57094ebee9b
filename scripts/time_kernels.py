"""Scratch timing of individual entry points with CUDA events (not the bench).

usage: python scripts/time_kernels.py [B:L ...]      e.g.  48:97 512:97 512:50
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gloria_nlp_project_b200 import ops, _lib


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(48, 97), (512, 97), (512, 50)]
n_it = int(os.environ.get("N_IT", "5"))
for B, L in cases:
    gen = torch.Generator(device="cuda").manual_seed(0)
    ctx = torch.randn(B, 768, 361, device="cuda", generator=gen)
    words = torch.randn(B, 768, 97, device="cuda", generator=gen)
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    packed = ops.tc_prepack(ctx, words, lens, L, 0)
    t_pack = timeit(lambda: ops.tc_prepack(ctx, words, lens, L, 0), n=n_it)
    lib = _lib.lib()
    sim = torch.empty(B, B, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def fwd():
        rc = lib.gloria_b200_tc_local_sim_fwd(packed.ctx_h.data_ptr(), packed.ctx_n.data_ptr(), packed.words_h.data_ptr(), packed.wnorm.data_ptr(), lens.data_ptr(), B, B, 768, 361, L, 4.0, 5.0, 0,
                                              1e-8, sim.data_ptr(), None, st)
        assert rc == 0, lib.gloria_b200_last_error()

    t = timeit(fwd, n=n_it)
    flops = 4 * 361 * 768 * B * B * L
    print(f"B={B} L={L}: prepack {t_pack:.3f} ms; tc fwd {t:.3f} ms -> {flops / t / 1e9:.1f} TFLOP/s algorithmic "
          f"({flops / t / 1e9 / 1656.1 * 100:.1f}% of measured bf16 burst)")
