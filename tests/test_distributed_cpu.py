"""world_size-2 gloo test of the caption-sharded loss (host-side collective logic).

The block-similarity functions are the torch-CPU oracle here (the CUDA ops cannot run without a GPU); what is under
test is the sharding: all_gather of image features, per-rank logit column blocks, the logit all_gather, and the
reduce_scatter of the image-feature gradients.  Sharded result == unsharded oracle on the concatenated batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gloria_oracle_torch as T
from oracle.make_golden import gen_inputs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_fns():
    def local_sim(img_all, words_local, cap_lens_local, temp1, temp2, agg):
        return T.local_similarities(img_all, words_local, cap_lens_local, temp1, temp2, agg)[0]

    def global_cos(x, y, eps):
        return (x @ y.t()) / (x.norm(dim=1, keepdim=True) @ y.norm(dim=1, keepdim=True).t()).clamp(min=eps)

    def ce(m, scale):
        n = m.shape[0]
        lab = torch.arange(n)
        return (torch.nn.functional.cross_entropy(m * scale, lab), torch.nn.functional.cross_entropy(m.t() * scale, lab))

    return dict(local_sim_fn=local_sim, global_cos_fn=global_cos, ce_fn=ce)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gloria_nlp_project_b200 import distributed as D
        torch.set_num_threads(2)
        B, Dm, H, W, Lw = 4, 32, 3, 4, 9
        img_l, txt_l, img_g, txt_g, cl = gen_inputs(31, B, Dm, H, W, Lw, cap_lens=[9, 7, 4, 2])
        n = B // world
        sl = slice(rank * n, (rank + 1) * n)
        leaves = [torch.tensor(a[sl]).requires_grad_(True) for a in (img_l, txt_l, img_g, txt_g)]
        l0, l1, g0, g1 = D.sharded_loss(leaves[0], leaves[1], leaves[2], leaves[3], cl[sl], **_oracle_fns())
        (l0 + 0.7 * l1 + 0.5 * g0 + 0.3 * g1).backward()
        q.put((rank, [float(v) for v in (l0, l1, g0, g1)], [t.grad.numpy() for t in leaves]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_equals_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # unsharded oracle on the concatenated batch
    B, Dm, H, W, Lw = 4, 32, 3, 4, 9
    img_l, txt_l, img_g, txt_g, cl = gen_inputs(31, B, Dm, H, W, Lw, cap_lens=[9, 7, 4, 2])
    leaves = [torch.tensor(a).requires_grad_(True) for a in (img_l, txt_l, img_g, txt_g)]
    l0, l1, _, _ = T.local_loss(leaves[0], leaves[1], cl)
    g0, g1 = T.global_loss(leaves[2], leaves[3])
    (l0 + 0.7 * l1 + 0.5 * g0 + 0.3 * g1).backward()
    want = [float(v) for v in (l0, l1, g0, g1)]
    n = B // world
    for rank, losses, grads in res:
        np.testing.assert_allclose(losses, want, rtol=1e-12)
        for gsh, leaf in zip(grads, leaves):
            np.testing.assert_allclose(gsh, leaf.grad.numpy()[rank * n:(rank + 1) * n], rtol=1e-9, atol=1e-14)


def test_single_process_path_matches():
    """Without an initialised process group the function is the plain full-batch loss."""
    from gloria_nlp_project_b200 import distributed as D
    img_l, txt_l, img_g, txt_g, cl = gen_inputs(32, 3, 16, 2, 3, 6, cap_lens=[6, 3, 1])
    t = [torch.tensor(a) for a in (img_l, txt_l, img_g, txt_g)]
    l0, l1, g0, g1 = D.sharded_loss(*t, cl, **_oracle_fns())
    r0, r1, _, _ = T.local_loss(t[0], t[1], cl)
    q0, q1 = T.global_loss(t[2], t[3])
    np.testing.assert_allclose([float(l0), float(l1), float(g0), float(g1)], [float(r0), float(r1), float(q0), float(q1)],
                               rtol=1e-12)


def test_part_major_rows():
    """Index math of the part-pipelined sharded path: simulate the per-part all_gathers of image ids and check that
    `part_major_rows` finds every image where the gathers put it (and that reduce_scatter blocks are rank-major)."""
    import torch
    from gloria_nlp_project_b200.distributed import part_major_rows
    for world, n, parts in [(2, 4, 2), (8, 64, 2), (4, 6, 3), (3, 5, 1)]:
        m = n // parts
        own = [torch.arange(r * n, (r + 1) * n) for r in range(world)]          # image ids held by each rank
        gathered = torch.cat([torch.cat([own[r][p * m:(p + 1) * m] for r in range(world)]) for p in range(parts)])
        rows = part_major_rows(n, world, parts)
        assert sorted(rows.tolist()) == list(range(world * n))
        assert torch.equal(gathered[rows], torch.arange(world * n))
        # reduce_scatter of part p hands rank r the block [p*world*m + r*m, +m) = its own images p*m .. (p+1)*m
        for p in range(parts):
            for r in range(world):
                blk = gathered[p * world * m + r * m: p * world * m + (r + 1) * m]
                assert torch.equal(blk, own[r][p * m:(p + 1) * m])


def _agree_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gloria_nlp_project_b200 import distributed as D
        # rank 0: short captions, workspace fits;  rank 1: a 97-word caption and (say) no room for the workspace
        out = [D._agree(torch.device("cpu"), 50 if rank == 0 else 97, rank == 0, None),
               D._agree(torch.device("cpu"), 30 + rank, True, None)]
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_ranks_agree_on_the_collective_schedule():
    """The pipelined (P-part) and the single-gather paths issue different collectives, so the choice must not depend
    on rank-local state: one rank that cannot take the pipelined path moves every rank to the fallback, and the padded
    caption length is the maximum over ranks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_agree_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0] == res[1] == [(97, False), (31, True)]


class _Model:
    temp1, temp2, temp3 = 4.0, 5.0, 10.0
    local_loss_weight, global_loss_weight, segmentation_loss_weight = 1.0, 0.5, 2.0
    no_attn_vec = None


def _maps_fn(img, txt, cl):
    return T.local_loss(img, txt, cl)[2]             # att_maps of the (local) diagonal pairs, differentiable


def _calc_inputs():
    B, Dm, H, W, Lw = 4, 32, 3, 4, 9
    img_l, txt_l, img_g, txt_g, cl = gen_inputs(33, B, Dm, H, W, Lw, cap_lens=[9, 7, 4, 2])
    labels = (np.random.default_rng(3).random((B, 6, 8)) > 0.6).astype(np.float64)
    labels[:, 0, 0] = 1.0
    return B, img_l, txt_l, img_g, txt_g, cl, labels


def _calc_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gloria_nlp_project_b200 import distributed as D
        torch.set_num_threads(2)
        B, img_l, txt_l, img_g, txt_g, cl, labels = _calc_inputs()
        n = B // world
        sl = slice(rank * n, (rank + 1) * n)
        leaves = [torch.tensor(a[sl]).requires_grad_(True) for a in (img_l, txt_l, img_g, txt_g)]
        loss, maps = D.sharded_calc_loss(_Model(), leaves[0], leaves[2], leaves[1], leaves[3], cl[sl],
                                         segmentation_labels=torch.tensor(labels[sl]), attn_maps_fn=_maps_fn,
                                         **_oracle_fns())
        loss.backward()
        q.put((rank, float(loss), [t.grad.numpy() for t in leaves], [tuple(m.shape) for m in maps]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_calc_loss_equals_unsharded():
    """GLoRIA.calc_loss (contrastive terms + supervised attention on the diagonal pairs) from two shards == the same
    terms on the whole batch: loss value on every rank, gradients on each rank's shard."""
    from gloria_nlp_project_b200 import gloria_loss
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_calc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    B, img_l, txt_l, img_g, txt_g, cl, labels = _calc_inputs()
    leaves = [torch.tensor(a).requires_grad_(True) for a in (img_l, txt_l, img_g, txt_g)]
    l0, l1, maps, _ = T.local_loss(leaves[0], leaves[1], cl)
    g0, g1 = T.global_loss(leaves[2], leaves[3])
    m = _Model()
    want = (l0 + l1) * m.local_loss_weight + (g0 + g1) * m.global_loss_weight + \
        gloria_loss.supervised_attention_loss(maps, torch.tensor(labels)) * m.segmentation_loss_weight
    want.backward()
    n = B // world
    for rank, loss, grads, shapes in res:
        assert abs(loss - float(want)) < 1e-12 * abs(float(want))
        assert shapes == [(1, L, 3, 4) for L in cl[rank * n:(rank + 1) * n]]
        for gsh, leaf in zip(grads, leaves):
            np.testing.assert_allclose(gsh, leaf.grad.numpy()[rank * n:(rank + 1) * n], rtol=1e-9, atol=1e-14)
