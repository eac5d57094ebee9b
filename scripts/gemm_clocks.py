"""Development aid: where the warps of the CTA-pair GEMM wait (library built with -DGLORIA_PHASE_CLOCKS).
usage: GLORIA_B200_LIB=.../libgloria_b200_clk.so python scripts/gemm_clocks.py [B]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gloria_nlp_project_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sp, lp, D = 368, 104, 768
Mr, Kr = B * sp, 512 * lp
lib = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
gen = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((Mr, Kr), device="cuda", generator=gen, dtype=torch.bfloat16)
Wt = torch.randn((Kr, D), device="cuda", generator=gen, dtype=torch.bfloat16)
g = torch.randn((B, 512), device="cuda", generator=gen)
dR = torch.empty((Mr, D), device="cuda")
dbg = torch.zeros(148, 40, dtype=torch.int64, device="cuda")
lib.gloria_b200_debug_phase_clocks(C.c_void_p(dbg.data_ptr()))


def show(name, fn):
    fn(); torch.cuda.synchronize()
    dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    d = dbg.cpu().double()
    lead = d[0::2]
    n = lead[:, 4].mean()
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms, k-blocks/CTA {n:.0f}; cycles per k-block:")
    m = lead.mean(0)
    print(f"  MMA issuer (leader): total {m[0]/n:.0f}  wait B full {m[1]/n:.0f}  wait A ready {m[2]/n:.0f}  wait acc drained {m[3]/n:.0f}")
    for r, nm in ((0, "leader"), (1, "peer")):
        m = d[r::2].mean(0)
        print(f"  {nm}: producers: total {m[8]/n:.0f} wait A empty {m[9]/n:.0f} wait B empty {m[10]/n:.0f}")
        for gi in range(2):
            o = 16 + 8 * gi
            print(f"  {nm}: scale group {gi}: total {m[o]/n:.0f} wait A full {m[o+1]/n:.0f} wait TMEM free {m[o+2]/n:.0f} "
                  f"lds+math {m[o+3]/n:.0f} st+wait {m[o+4]/n:.0f} fence+arrive {m[o+6]/n:.0f} epilogue {m[o+5]/n:.0f}")


def own(A, Bm, Cm, M, N, K, ak, gp, g_sm, g_sk, m_div, k_div, force):
    return lambda: _lib.check(lib.gloria_b200_acc_gemm(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, ak, 1, 0,
                                                      gp.data_ptr() if gp is not None else None, g_sm, g_sk, m_div,
                                                      k_div, force, st), "acc_gemm")


Rt = torch.randn((Mr, D), device="cuda", generator=gen, dtype=torch.bfloat16)
dW = torch.empty((Kr, D), device="cuda")
show("dR plain", own(X, Wt, dR, Mr, D, Kr, 1, None, 0, 0, 1, 1, 0))
show("dR TMEM weight 1", own(X, Wt, dR, Mr, D, Kr, 1, None, 0, 0, 1, 1, 1))
show("dR TMEM scaled", own(X, Wt, dR, Mr, D, Kr, 1, g, 512, 1, sp, lp, 0))
show("dW plain", own(X, Rt, dW, Kr, D, Mr, 0, None, 0, 0, 1, 1, 0))
show("dW in-place weight 1", own(X, Rt, dW, Kr, D, Mr, 0, None, 0, 0, 1, 1, 1))
show("dW in-place scaled", own(X, Rt, dW, Kr, D, Mr, 0, g, 1, 512, lp, sp, 0))
