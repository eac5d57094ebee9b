"""One launch of each accumulation-GEMM variant (and cuBLAS on the same operands) for an ncu capture.
usage: python scripts/gemm_ncu.py [B_images]  (K = 512 captions x 104)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gloria_nlp_project_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sp, lp, D = 368, 104, 768
Mr, Kr = B * sp, 512 * lp
lib = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
gen = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((Mr, Kr), device="cuda", generator=gen, dtype=torch.bfloat16)
Wt = torch.randn((Kr, D), device="cuda", generator=gen, dtype=torch.bfloat16)
Rt = torch.randn((Mr, D), device="cuda", generator=gen, dtype=torch.bfloat16)
g = torch.randn((B, 512), device="cuda", generator=gen)
dR = torch.empty((Mr, D), device="cuda")
dW = torch.empty((Kr, D), device="cuda")


def own(A, Bm, Cm, M, N, K, ak, gp, g_sm, g_sk, m_div, k_div, force):
    _lib.check(lib.gloria_b200_acc_gemm(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, ak, 1, 0,
                                        gp.data_ptr() if gp is not None else None, g_sm, g_sk, m_div, k_div, force, st), "acc_gemm")


for _ in range(2):
    torch.matmul(X, Wt)
    own(X, Wt, dR, Mr, D, Kr, 1, None, 0, 0, 1, 1, 0)          # plain, A from smem
    own(X, Wt, dR, Mr, D, Kr, 1, g, 512, 1, sp, lp, 0)          # scaled through TMEM
    own(X, Rt, dW, Kr, D, Mr, 0, None, 0, 0, 1, 1, 0)          # plain, A^T from smem
    own(X, Rt, dW, Kr, D, Mr, 0, g, 1, 512, lp, sp, 0)          # scaled in place
torch.cuda.synchronize()
print("ok")
