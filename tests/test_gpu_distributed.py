"""Caption-sharded loss on real GPUs over NCCL (needs >= 2 GPUs; skipped on a single-GPU box).

Sharded result (2 ranks) == single-GPU result of the same kernels == oracle, losses and gradients.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gloria_oracle as O
from oracle.make_golden import gen_inputs

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    B = 8
    img_l, txt_l, img_g, txt_g, cl = gen_inputs(41, B, 768, 19, 19, 97, cap_lens=[97, 80, 64, 40, 33, 17, 9, 4], scale=0.05,
                                                dtype=np.float32)
    return B, img_l, txt_l, img_g, txt_g, cl


def _worker(rank, world, port, precision, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gloria_nlp_project_b200 as G
        from gloria_nlp_project_b200 import distributed as D
        G.set_precision(precision)
        B, img_l, txt_l, img_g, txt_g, cl = _inputs()
        n = B // world
        sl = slice(rank * n, (rank + 1) * n)
        leaves = [torch.tensor(a[sl], device="cuda").requires_grad_(True) for a in (img_l, txt_l, img_g, txt_g)]
        l0, l1, g0, g1 = D.sharded_loss(leaves[0], leaves[1], leaves[2], leaves[3], cl[sl])
        (l0 + 0.7 * l1 + 0.5 * g0 + 0.3 * g1).backward()
        torch.cuda.synchronize()
        q.put((rank, [float(v) for v in (l0, l1, g0, g1)], [t.grad.cpu().numpy() for t in leaves]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 1e-5, 1e-3), ("bf16", 2e-3, 1e-2)])
def test_sharded_two_gpus(precision, tol, gtol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, precision, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    B, img_l, txt_l, img_g, txt_g, cl = _inputs()
    i64, t64, ig64, tg64 = (a.astype(np.float64) for a in (img_l, txt_l, img_g, txt_g))
    o0, o1, *_ = O.local_loss(i64, t64, cl)
    q0, q1, _ = O.global_loss(ig64, tg64)
    d_img, d_txt = O.local_loss_bwd(i64, t64, cl, g0=1.0, g1=0.7)
    d_ig, d_tg = O.global_loss_bwd(ig64, tg64, g0=0.5, g1=0.3)
    want = [o0, o1, q0, q1]
    n = B // world

    def rel(a, b):
        return float(np.abs(a - b).max() / np.abs(b).max())
    for rank, losses, grads in res:
        for a, b in zip(losses, want):
            assert abs(a - b) / abs(b) < tol, (losses, want)
        sl = slice(rank * n, (rank + 1) * n)
        for got, ref in zip(grads, (d_img, d_txt, d_ig, d_tg)):
            assert rel(got.astype(np.float64), ref[sl]) < gtol
