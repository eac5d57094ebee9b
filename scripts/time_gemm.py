"""Time the CTA-pair accumulation GEMM (csrc/tc_gemm.cu) on the B = 512 backward shapes against torch.matmul (cuBLAS).
usage: python scripts/time_gemm.py [B]   (default 512)"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gloria_nlp_project_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sp, lp, D = 368, 104, 768
Mr, Kr = B * sp, B * lp
lib = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
gen = torch.Generator(device="cuda").manual_seed(0)
X = (torch.randn((Mr, Kr), device="cuda", generator=gen, dtype=torch.bfloat16))
Wt = torch.randn((Kr, D), device="cuda", generator=gen, dtype=torch.bfloat16)
Rt = torch.randn((Mr, D), device="cuda", generator=gen, dtype=torch.bfloat16)
g = torch.randn((B, B), device="cuda", generator=gen)
dR = torch.empty((Mr, D), device="cuda")
dW = torch.empty((Kr, D), device="cuda")


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, fn, flops):
    ms = timed(fn)
    print(f"{name:44s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


fl = 2.0 * Mr * Kr * D


def own(A, Bm, Cm, M, N, K, ak, ks, gp, g_sm, g_sk, m_div, k_div, force):
    return lambda: _lib.check(lib.gloria_b200_acc_gemm(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, ak, ks, 0,
                                                      gp.data_ptr() if gp is not None else None, g_sm, g_sk, m_div,
                                                      k_div, force, st), "acc_gemm")


run("cuBLAS dR  (X @ Wt, bf16 out)", lambda: torch.matmul(X, Wt), fl)
run("cuBLAS dW  (X^T @ Rt, bf16 out)", lambda: torch.matmul(X.t(), Rt), fl)
run("own dR plain (A from smem)", own(X, Wt, dR, Mr, D, Kr, 1, 1, None, 0, 0, 1, 1, 0), fl)
run("own dR through TMEM, weight 1", own(X, Wt, dR, Mr, D, Kr, 1, 1, None, 0, 0, 1, 1, 1), fl)
run("own dR through TMEM, scaled by g", own(X, Wt, dR, Mr, D, Kr, 1, 1, g, B, 1, sp, lp, 0), fl)
w = g.repeat_interleave(sp, 0)[:1024].repeat_interleave(lp, 1)
ref = ((X[:1024].float() * w).to(torch.bfloat16).float() @ Wt.float())
print("dR rows 0..1023 rel err", float((dR[:1024] - ref).abs().max() / ref.abs().max()))
run("own dW plain (A^T from smem)", own(X, Rt, dW, Kr, D, Mr, 0, 1, None, 0, 0, 1, 1, 0), fl)
run("own dW scaled in place, weight 1", own(X, Rt, dW, Kr, D, Mr, 0, 1, None, 0, 0, 1, 1, 1), fl)
run("own dW scaled in place by g", own(X, Rt, dW, Kr, D, Mr, 0, 1, g, 1, B, lp, sp, 0), fl)
wt = g.repeat_interleave(sp, 0)[:, :8].repeat_interleave(lp, 1)[:, :512]            # [Mr, 512] weights of columns 0..511
ref2 = (X[:, :512].float() * wt).to(torch.bfloat16).float().t() @ Rt.float()
print("dW rows 0..511 rel err", float((dW[:512] - ref2).abs().max() / ref2.abs().max()))
