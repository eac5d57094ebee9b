"""CPU-side checks of the boundary: the C-ABI library builds, loads, and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__
    __graft_entry__.build()
    from gloria_nlp_project_b200.build import LIB_PATH
    assert os.path.exists(LIB_PATH)
    return LIB_PATH


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gloria_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gloria_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(libpath):
    h = ctypes.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/gloria_b200.h but not exported"


def test_ctypes_prototypes_cover_header(libpath):
    from gloria_nlp_project_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    assert _lib.lib().gloria_b200_version() >= 100


def test_argument_validation_without_gpu(libpath):
    """Entry points validate arguments before touching the device: bad calls return an error code + message."""
    from gloria_nlp_project_b200 import _lib
    L = _lib.lib()
    rc = L.gloria_b200_local_sim_fwd_f32(None, None, None, 1, 1, 1, 1, 1, 1, 0, 4.0, 5.0, 0, 1e-8, None, None, None,
                                         None, 0, None)
    assert rc == 1 and b"null" in L.gloria_b200_last_error()
    rc = L.gloria_b200_ce_bidir_fwd(None, 4, 1.0, None, None, None, None)
    assert rc == 1
    assert L.gloria_b200_local_f32_workspace(48, 48, 768, 361, 97, 97, 0) > 0
    with pytest.raises(RuntimeError):
        _lib.check(rc, "ce_bidir_fwd")


def test_no_cpu_fallback():
    import torch
    from gloria_nlp_project_b200 import gloria_loss
    with pytest.raises(RuntimeError, match="CUDA"):
        gloria_loss.local_loss(torch.randn(2, 8, 2, 2), torch.randn(2, 8, 4), [3, 2])
    with pytest.raises(RuntimeError, match="CUDA"):
        gloria_loss.global_loss(torch.randn(2, 8), torch.randn(2, 8))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "gloria_nlp_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_supervised_attention_loss_per_cell_form():
    """gloria_model.py:143-147 (mean over words -> nearest upsample -> normalise -> -log of the labelled mass) against
    the per-cell form used by the drop-in (pure torch glue, so it runs on CPU tensors)."""
    import torch
    from gloria_nlp_project_b200.gloria_loss import supervised_attention_loss
    gen = torch.Generator().manual_seed(5)
    B, ih, iw = 4, 19, 19
    maps = []
    for L in (7, 3, 1, 11):
        a = torch.rand(1, L, ih * iw, generator=gen, dtype=torch.float64).softmax(-1).reshape(1, L, ih, iw)
        maps.append(a.requires_grad_())
    labels = torch.rand(B, 224, 224, generator=gen) > 0.7
    mean_maps = torch.cat([m.mean(1) for m in maps], 0)
    up = torch.nn.functional.interpolate(mean_maps.unsqueeze(1), size=labels.shape[1:]).squeeze(1)
    up = up / up.sum(-1, keepdims=True).sum(-2, keepdims=True)
    ref = -torch.log((labels * up).sum(-1).sum(-1)).mean()
    g_ref = torch.autograd.grad(ref, maps)
    out = supervised_attention_loss(maps, labels)
    g_out = torch.autograd.grad(out, maps)
    assert abs(float(out) - float(ref)) < 1e-12
    for a, b in zip(g_out, g_ref):
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-14)
    # non-square label resolution, as in the golden fixture (24 x 30)
    labels = torch.rand(B, 24, 30, generator=gen) > 0.5
    up = torch.nn.functional.interpolate(mean_maps.unsqueeze(1), size=labels.shape[1:]).squeeze(1)
    up = up / up.sum(-1, keepdims=True).sum(-2, keepdims=True)
    ref = -torch.log((labels * up).sum(-1).sum(-1)).mean()
    assert abs(float(supervised_attention_loss(maps, labels)) - float(ref)) < 1e-12


def test_length_bucket_plan():
    """Host logic of the length-bucketed bf16 path: groups partition the captions, each group's lcap is its longest
    caption, uniform lengths and small batches stay in one launch, a large ragged batch is split by padded length."""
    import torch
    from gloria_nlp_project_b200.gloria_loss import plan_length_buckets
    assert plan_length_buckets([97] * 512, 512) == [(list(range(512)), 97)]
    gen = torch.Generator().manual_seed(1)
    small = torch.randint(5, 98, (48,), generator=gen).tolist()
    assert len(plan_length_buckets(small, 48)) == 1                      # a launch costs more than the padding saves
    lens = sorted(torch.randint(5, 98, (512,), generator=gen).tolist(), reverse=True)
    plan = plan_length_buckets(lens, 512)
    assert len(plan) > 2
    seen = sorted(i for idx, _ in plan for i in idx)
    assert seen == list(range(512))
    for idx, lcap in plan:
        assert lcap == max(lens[i] for i in idx)
    pads = [(lcap + 15) // 16 * 16 for _, lcap in plan]
    assert pads == sorted(pads, reverse=True) and len(set(pads)) == len(pads)
    work = sum(len(idx) * p for (idx, _), p in zip(plan, pads))
    assert work < 0.65 * 512 * 112                                       # U{5..97}: ~52 % of the single-launch work
    # unsorted input: same partition by padded length
    shuffled = [lens[(7 * i) % 512] for i in range(512)]
    plan2 = plan_length_buckets(shuffled, 512)
    assert sorted(len(idx) for idx, _ in plan2) == sorted(len(idx) for idx, _ in plan)


def test_new_entry_points_validate_arguments(libpath):
    """Packed prompts, part-wise training calls and the word aggregation: host-side arithmetic and argument checks
    (no device work is reached)."""
    import torch
    from gloria_nlp_project_b200 import _lib, text_model, zero_shot
    L = _lib.lib()
    # 25 prompts -> 4 word tiles of 7 prompts (16 words each); 8 -> one full tile; 9 -> 2 tiles of 5
    assert (L.gloria_b200_tc_packed_groups(25), L.gloria_b200_tc_packed_per(25)) == (4, 7)
    assert (L.gloria_b200_tc_packed_groups(8), L.gloria_b200_tc_packed_per(8)) == (1, 8)
    assert (L.gloria_b200_tc_packed_groups(9), L.gloria_b200_tc_packed_per(9)) == (2, 5)
    assert L.gloria_b200_tc_packed_groups(0) == 0
    assert L.gloria_b200_tc_local_sim_fwd_packed(None, None, None, None, None, 4, 4, 768, 361, 4.0, 5.0, 2, 1e-8, None,
                                                 None) == 1
    # the image parts of the backward must divide the batch
    one = 1 << 20
    rc = L.gloria_b200_tc_local_sim_bwd_train_parts(one, one, one, 6, 2, 768, 361, 97, 97, 0, one, one, one, one, 1 << 30,
                                                    4, None, None)
    assert rc == 1 and b"n_parts" in L.gloria_b200_last_error()
    rc = L.gloria_b200_tc_local_sim_fwd_train_part(one, one, one, one, one, 8, 6, 4, 2, 768, 361, 97, 4.0, 5.0, 0, 1e-8,
                                                   one, one, 1 << 30, None)
    assert rc == 1 and b"image range" in L.gloria_b200_last_error()
    assert L.gloria_b200_aggregate_tokens_fwd(None, 0, None, 2, 4, 16, 24, None, None) == 1
    assert L.gloria_b200_word_ranges(None, None, 10, 2, 2, 16, None, None, None, None) == 1
    # no CPU fallback in the callers on either side of the path
    table = text_model.VocabTable({0: "[PAD]", 1: "[CLS]", 2: "[SEP]", 3: "a", 4: "##b"})
    assert table.sep_id == 2 and table.cont[4] and table.piece[4] == "b"
    with pytest.raises(RuntimeError, match="CUDA"):
        text_model.aggregate_tokens(torch.randn(1, 2, 4, 8), torch.tensor([[1, 3, 4, 2]]), table)
    with pytest.raises(RuntimeError):
        zero_shot.get_similarities(object(), torch.zeros(1), {"caption_ids": None}, similarity_type="cosine")


def test_lazy_containers_host_logic():
    """LazySentences / LazyAttnMaps / DeviceCapLens are plain host containers over tensors: their list behaviour is
    testable on CPU tensors (the kernels that fill them are covered by the `-m gpu` tests)."""
    import torch
    from gloria_nlp_project_b200 import gloria_loss, text_model
    vocab = ["[PAD]", "[CLS]", "[SEP]", "heart", "##s", "is", "[UNK]", "normal"]
    table = text_model.VocabTable(dict(enumerate(vocab)))
    assert table._brk_cpu.tolist() == [1, 1, 1, 0, 0, 0, 1, 0]
    ids = torch.tensor([[1, 3, 4, 5, 7, 2, 0, 0], [1, 6, 5, 2, 0, 0, 0, 0]])
    ref = text_model._sentences(ids.tolist(), table)
    lens = [len([w for w in s if not w.startswith("[")]) + 1 for s in ref]
    lazy = text_model.LazySentences(ids, table, torch.tensor(lens, dtype=torch.int32), torch.tensor([5, 4]))
    assert lazy._built is None and len(lazy) == 2
    assert text_model.cap_lens_from_sents(lazy) is lazy.cap_lens and lazy._built is None
    assert lazy == ref and ref == lazy and lazy[0][1] == "hearts" and [len(s) for s in lazy] == [8, 8]
    assert lazy.cap_lens.tolist() == lens == [4, 2]
    diag = torch.arange(2 * 6 * 5, dtype=torch.float32).reshape(2, 6, 5)
    maps = gloria_loss.LazyAttnMaps(diag, lazy.cap_lens, 1, 2, 2)
    assert len(maps) == 2 and maps[0].shape == (1, 4, 2, 2) and maps[-1].shape == (1, 2, 2, 2)
    assert torch.equal(maps[1], diag[1:2, :2, 1:].reshape(1, 2, 2, 2))
    assert [m.shape[1] for m in maps] == [4, 2]


def test_f32_tensor_core_host_side(libpath):
    """fp32 mode on the tensor cores (tc_f32.cu): shape coverage, workspace arithmetic and argument checks are host code."""
    from gloria_nlp_project_b200 import _lib
    L = _lib.lib()
    assert L.gloria_b200_f32tc_supported(768, 361, 97) == 0          # the path's own shape
    assert L.gloria_b200_f32tc_supported(768, 362, 128) == 0         # + the no-attention column, longest caption covered
    assert L.gloria_b200_f32tc_supported(48, 361, 97) != 0           # D % 64: stays on the CUDA-core kernels
    assert L.gloria_b200_f32tc_supported(768, 361, 129) != 0         # captions beyond 128 words
    assert L.gloria_b200_f32tc_supported(4096, 361, 97) != 0
    fwd = L.gloria_b200_local_f32tc_workspace(48, 48, 768, 361, 97, 97, 0, 0)
    bwd = L.gloria_b200_local_f32tc_workspace(48, 48, 768, 361, 97, 97, 0, 1)
    assert 0 < fwd < bwd < 8 << 30                                   # the backward layout holds more buffers; 3.9 GB at B = 48
    capped = L.gloria_b200_local_f32tc_workspace(512, 512, 768, 361, 97, 97, 4 << 30, 1)
    least = L.gloria_b200_local_f32tc_workspace(512, 1, 768, 361, 97, 97, 0, 1)
    assert capped == max(4 << 30, capped) and capped >= least > 0    # a budget below one caption's chunk is raised to it
    assert L.gloria_b200_local_f32tc_workspace(48, 48, 768, 361, 97, 97, 1 << 30, 0) <= max(1 << 30, fwd)
    assert L.gloria_b200_local_f32tc_workspace(0, 48, 768, 361, 97, 97, 0, 0) == 0
    rc = L.gloria_b200_local_sim_fwd_f32tc(None, None, None, 1, 1, 64, 1, 1, 1, 0, 4.0, 5.0, 0, 1e-8, None, None, None, None, 0,
                                           None)
    assert rc == 1 and b"null" in L.gloria_b200_last_error()
    # split-precision GEMM: bad term counts / shapes are refused before any launch
    one = ctypes.c_void_p(16)
    assert L.gloria_b200_acc_gemm_planes(one, one, one, 128, 128, 64, 1, 4, 1, 1, 0, None) == 1
    assert b"nterms" in L.gloria_b200_last_error()
    assert L.gloria_b200_acc_gemm_planes(one, one, one, 128, 100, 64, 1, 6, 1, 1, 0, None) == 1       # N % 16
    assert L.gloria_b200_acc_gemm_planes(one, one, one, 128, 128, 72, 0, 3, 2, 1, 0, None) == 1       # transposed A: K % 64
    assert L.gloria_b200_upload_ints(None, 4, None, None) == 1


def test_packed_image_gradient_entry_points(libpath):
    """Packed reduce_scatter of the sharded backward: where dRt sits in the training workspace, and the unpack's checks."""
    from gloria_nlp_project_b200 import _lib
    L = _lib.lib()
    total = L.gloria_b200_tc_train_workspace(512, 64, 768, 361, 97)
    off = L.gloria_b200_tc_train_drt_offset(512, 64, 768, 361, 97)
    sp = L.gloria_b200_tc_sp(361)
    assert off > 0 and off % 1024 == 0 and off + 512 * sp * 768 * 4 <= total     # dRt [Bi, sp, D] fp32 lies inside
    assert L.gloria_b200_tc_train_drt_offset(512, 64, 48, 361, 97) == 0          # unsupported shape
    assert L.gloria_b200_tc_train_drt_offset(0, 64, 768, 361, 97) == 0
    one = ctypes.c_void_p(16)
    assert L.gloria_b200_tc_unpack_dctx(None, one, 4, 768, 361, None) == 1 and b"null" in L.gloria_b200_last_error()
    assert L.gloria_b200_tc_unpack_dctx(one, one, 4, 100, 361, None) == 1         # D % 32
    assert L.gloria_b200_tc_unpack_dctx(one, one, 0, 768, 361, None) == 1
