"""How exact is fp32 accumulation in tensor memory?  (decides whether a split-bf16 fp32 mode can hold the 1e-5 gate)

C = A B with bf16 operands (products exact in fp32) through gloria_b200_acc_gemm, against fp64 of the same operands and
against an fp32 SIMT GEMM (torch.matmul, TF32 off).  Also: a 3-piece bf16 split of fp32 operands, six terms
concatenated along K, small terms first -- the form the fp32 mode would use.
"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gloria_nlp_project_b200 import _lib

torch.backends.cuda.matmul.allow_tf32 = False
lib = _lib.lib()


def gemm(A, B, M, N, K):
    out = torch.empty((M, N), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.gloria_b200_acc_gemm(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, 1, 1, 0, None, 0, 0, 1, 1, 0, C.c_void_p(st))
    _lib.check(rc, "acc_gemm")
    torch.cuda.synchronize()
    return out


def stats(name, out, ref64):
    err = (out.double() - ref64)
    scale = ref64.abs().mean()
    bias = (err * ref64.sign()).mean() / scale       # > 0: magnitudes too large, < 0: rounded toward zero
    print(f"{name:44s} max|err|/mean|C| {float(err.abs().max() / scale):.3e}  rms {float(err.pow(2).mean().sqrt() / scale):.3e}  signed bias {float(bias):+.3e}")


def split3(x):
    p1 = x.to(torch.bfloat16)
    r = x - p1.float()
    p2 = r.to(torch.bfloat16)
    p3 = (r - p2.float()).to(torch.bfloat16)
    return p1, p2, p3


g = torch.Generator(device="cuda").manual_seed(0)
M, N = 512, 512
for K in (768, 4608, 18432):
    A = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn((K, N), device="cuda", generator=g).to(torch.bfloat16)
    ref = A.double() @ B.double()
    stats(f"K={K} tcgen05 bf16 (exact products)", gemm(A, B, M, N, K), ref)
    stats(f"K={K} fp32 SIMT (torch, TF32 off)", A.float() @ B.float(), ref)
    # positive operands: the accumulator grows monotonically, rounding mode shows as bias
    Ap, Bp = A.abs(), B.abs()
    refp = Ap.double() @ Bp.double()
    stats(f"K={K} tcgen05, positive operands", gemm(Ap, Bp, M, N, K), refp)
    stats(f"K={K} fp32 SIMT, positive operands", Ap.float() @ Bp.float(), refp)

# split-bf16 emulation of an fp32 GEMM, K = 768 (the score GEMM), terms ordered small -> large
K = 768
A = torch.randn((M, K), device="cuda", generator=g)
B = torch.randn((K, N), device="cuda", generator=g)
ref = A.double() @ B.double()
a1, a2, a3 = split3(A)
b1, b2, b3 = split3(B)
for name, ta, tb in (("6 terms small->large", (a3, a2, a1, a2, a1, a1), (b1, b2, b3, b1, b2, b1)),
                     ("6 terms large->small", (a1, a1, a2, a1, a2, a3), (b1, b2, b1, b3, b2, b1)),
                     ("3 terms (hi/lo pairs only)", (a2, a1, a1), (b1, b2, b1))):
    Ac = torch.cat(ta, 1).contiguous()
    Bc = torch.cat(tb, 0).contiguous()
    stats(f"split-bf16 {name}", gemm(Ac, Bc, M, N, Ac.shape[1]), ref)
stats("fp32 SIMT (torch, TF32 off)", A @ B, ref)
# main term in the tensor core, corrections accumulated separately and added in fp32
main = gemm(a1.contiguous(), b1.contiguous(), M, N, K)
Ac = torch.cat((a3, a2, a1, a2, a1), 1).contiguous()
Bc = torch.cat((b1, b2, b3, b1, b2), 0).contiguous()
corr = gemm(Ac, Bc, M, N, Ac.shape[1])
stats("split-bf16 main + corrections in 2 accumulators", main + corr, ref)
# main term cut into K-chunks of 128, chunks summed in fp32 outside the tensor core
parts = corr.clone()
for k0 in range(0, K, 128):
    parts += gemm(a1[:, k0:k0 + 128].contiguous(), b1[k0:k0 + 128].contiguous(), M, N, 128)
stats("split-bf16 main term in K=128 chunks (fp32 adds)", parts, ref)
