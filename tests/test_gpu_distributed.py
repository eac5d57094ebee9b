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


def _calc_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gloria_nlp_project_b200 as G
        from gloria_nlp_project_b200 import distributed as D
        from tests.util import Holder
        G.set_precision("bf16")
        B, img_l, txt_l, img_g, txt_g, cl = _inputs()
        labels = (np.random.default_rng(5).random((B, 64, 64)) > 0.7).astype(np.float32)
        n = B // world
        sl = slice(rank * n, (rank + 1) * n)
        m = Holder(local_loss_weight=1.0, global_loss_weight=0.5, segmentation_loss_weight=2.0)
        leaves = [torch.tensor(a[sl], device="cuda").requires_grad_(True) for a in (img_l, txt_l, img_g, txt_g)]
        loss, maps = D.sharded_calc_loss(m, leaves[0], leaves[2], leaves[1], leaves[3], cl[sl],
                                         segmentation_labels=torch.tensor(labels[sl], device="cuda"))
        loss.backward()
        torch.cuda.synchronize()
        q.put((rank, float(loss), [t.grad.cpu().numpy() for t in leaves], [tuple(x.shape) for x in maps]))
    finally:
        dist.destroy_process_group()


def test_sharded_calc_loss_two_gpus():
    """calc_loss (contrastive terms + supervised attention, gloria_model.py:132-150) from two NCCL ranks == the
    single-GPU drop-in on the whole batch (which is pinned to the reference's goldens `ft_*` / `calc_*`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import gloria_nlp_project_b200 as G
    from gloria_nlp_project_b200.gloria_model import GLoRIALossMixin
    from tests.util import Holder

    class M(GLoRIALossMixin, Holder):
        pass

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_calc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    B, img_l, txt_l, img_g, txt_g, cl = _inputs()
    labels = (np.random.default_rng(5).random((B, 64, 64)) > 0.7).astype(np.float32)
    G.set_precision("bf16")
    try:
        m = M(local_loss_weight=1.0, global_loss_weight=0.5, segmentation_loss_weight=2.0)
        leaves = [torch.tensor(a, device="cuda").requires_grad_(True) for a in (img_l, txt_l, img_g, txt_g)]
        sents = [["[CLS]"] + ["w"] * (L - 1) + ["[SEP]"] for L in cl]
        want, _ = m.calc_loss(leaves[0], leaves[2], leaves[1], leaves[3], sents,
                              segmentation_labels=torch.tensor(labels, device="cuda"))
        want.backward()
    finally:
        G.set_precision("auto")
    n = B // world
    for rank, loss, grads, shapes in res:
        assert abs(loss - float(want)) < 2e-4 * abs(float(want))
        assert shapes == [(1, L, 19, 19) for L in cl[rank * n:(rank + 1) * n]]
        for got, leaf in zip(grads, leaves):
            ref = leaf.grad.cpu().numpy()[rank * n:(rank + 1) * n]
            assert float(np.abs(got - ref).max() / np.abs(ref).max()) < 2e-3


def test_part_pipelined_path_one_rank_packed_gradient():
    """The part-pipelined sharded path on ONE GPU (a world of one rank: every collective is a copy): its image-side
    gradient travels in the library's packed layout and only the rank's own images are unpacked
    (gloria_b200_tc_unpack_dctx).  Packed and unpacked routes must agree, and both with the oracle."""
    import gloria_nlp_project_b200 as G
    from gloria_nlp_project_b200 import distributed as D, gloria_loss, ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        B, img_l, txt_l, img_g, txt_g, cl = _inputs()
        dev_lens, _ = gloria_loss._cap_lens(cl, B, 0, txt_l.shape[2], torch.device("cuda", 0))
        gsim = torch.tensor(np.random.default_rng(3).standard_normal((B, B)).astype(np.float32), device="cuda")
        outs = []
        for packed in (True, False):
            D._PACKED_RS = packed
            leaves = [torch.tensor(a, device="cuda").requires_grad_(True) for a in (img_l, txt_l)]
            sim = D._ShardedLocalSimParts.apply(leaves[0], leaves[1], dev_lens, txt_l.shape[2], 4.0, 5.0, ops.AGG["sum"],
                                                1e-8, None, 2)
            (sim * gsim).sum().backward()
            torch.cuda.synchronize()
            outs.append((sim.detach().cpu().numpy(), leaves[0].grad.cpu().numpy(), leaves[1].grad.cpu().numpy()))
    finally:
        D._PACKED_RS = True
        dist.destroy_process_group()
    for a, b in zip(outs[0], outs[1]):      # (not bit for bit: the fused kernel's column sums use shared-memory atomics)
        assert float(np.abs(a - b).max() / np.abs(b).max()) < 1e-4
    i64, t64 = img_l.astype(np.float64), txt_l.astype(np.float64)
    want = O.local_similarities(i64, t64, cl)
    assert float(np.abs(outs[0][0] - want).max() / np.abs(want).max()) < 2e-3
    g64 = gsim.cpu().numpy().astype(np.float64)
    ctx = i64.reshape(B, 768, -1)
    d_img, d_txt = np.zeros_like(ctx), np.zeros_like(t64)
    for i in range(B):
        dc, dw = O.local_sim_pair_bwd(ctx, t64[i, :, :cl[i]], 4.0, 5.0, g64[:, i])
        d_img += dc
        d_txt[i, :, :cl[i]] = dw
    assert float(np.abs(outs[0][1].reshape(B, 768, -1) - d_img).max() / np.abs(d_img).max()) < 1e-2
    assert float(np.abs(outs[0][2] - d_txt).max() / np.abs(d_txt).max()) < 1e-2
