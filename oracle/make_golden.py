"""Generate golden vectors by executing the REAL reference code (test infrastructure).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

Writes tests/golden/*.npz.  Every array in those files was produced by the
reference's own functions:
  * gloria/loss/gloria_loss.py loaded standalone with importlib (it imports only torch),
  * gloria/models/gloria_model.py::GLoRIA imported with stub modules for the
    dependencies that are absent here (SURVEY.md §8c recipe); the instance is built with
    GLoRIA.__new__ so no weights are downloaded, and the real calc_loss /
    get_local_similarities / get_global_similarities / get_attn_maps methods run.
Inputs come from numpy's PCG64 (`default_rng(seed)`); small cases store the inputs too, the
full-size case stores an input checksum so drift is detected instead of silently compared.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ---------------------------------------------------------------------------------------------
def load_reference_loss():
    spec = importlib.util.spec_from_file_location("ref_gloria_loss", f"{REF}/gloria/loss/gloria_loss.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_model_class():
    """Import the real GLoRIA class; modules that are absent in this image are stubbed on demand.

    `import gloria` pulls in the whole product (datasets, Lightning modules, UNet ...), none of which
    the loss path uses; every ModuleNotFoundError is answered with a permissive stub module and the
    import is retried, so exactly the missing names are stubbed and everything present stays real.
    """
    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any()

    def stub(name):
        m = types.ModuleType(name)
        m.__path__ = []

        def _getattr(k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any
        m.__getattr__ = _getattr
        sys.modules[name] = m
        return m

    if "numpy.lib.function_base" not in sys.modules:      # removed in NumPy 2 (vision_model.py:1)
        stub("numpy.lib.function_base").extract = np.extract
    sys.path.insert(0, REF)
    stubbed = []
    for _ in range(64):
        try:
            from gloria.models.gloria_model import GLoRIA  # noqa: E402
            break
        except ModuleNotFoundError as e:
            for k in [k for k in sys.modules if k == "gloria" or k.startswith("gloria.")]:
                del sys.modules[k]
            stub(e.name)
            stubbed.append(e.name)
    else:
        raise RuntimeError("could not import the reference GLoRIA class")
    print("stubbed modules:", stubbed)
    return GLoRIA


def make_model(GLoRIA, **kw):
    m = GLoRIA.__new__(GLoRIA)
    torch.nn.Module.__init__(m)
    ref_loss = sys.modules["gloria.loss.gloria_loss"]
    defaults = dict(local_loss=ref_loss.local_loss, global_loss=ref_loss.global_loss,
                    local_loss_weight=1.0, global_loss_weight=1.0, sparse_attn_loss_weight=None,
                    no_attn_loss_weight=None, attention_divergence_loss_weight=None,
                    attention_entropy_loss_weight=None, segmentation_loss_weight=None,
                    temp1=4.0, temp2=5.0, temp3=10.0, no_attn_vec=None)
    defaults.update(kw)
    for k, v in defaults.items():
        setattr(m, k, v)
    return m


# ---------------------------------------------------------------------------------------------
def gen_inputs(seed, B, D, H, W, Lmax, cap_lens=None, scale=1.0, dtype=np.float64, Bc=None):
    """Shared by make_golden.py and the tests: deterministic features, zero-padded words."""
    rng = np.random.default_rng(seed)
    Bc = B if Bc is None else Bc
    img_l = (rng.standard_normal((B, D, H, W)) * scale).astype(dtype)
    txt_l = (rng.standard_normal((Bc, D, Lmax)) * scale).astype(dtype)
    img_g = rng.standard_normal((B, D)).astype(dtype)
    txt_g = rng.standard_normal((Bc, D)).astype(dtype)
    if cap_lens is None:
        cap_lens = sorted((int(x) for x in rng.integers(2, Lmax + 1, size=Bc)), reverse=True)
    for i, L in enumerate(cap_lens):
        txt_l[i, :, L:] = 0
    return img_l, txt_l, img_g, txt_g, list(cap_lens)


def checksum(*arrs):
    return np.array([float(np.sum(np.asarray(a, dtype=np.float64) * np.cos(np.arange(a.size, dtype=np.float64)).reshape(a.shape)))
                     for a in arrs])


def t(x, grad=False):
    return torch.tensor(x, requires_grad=grad)


def run_local(ref, img_l, txt_l, cap_lens, grad_w=(1.0, 1.0), **kw):
    ti, tw = t(img_l, True), t(txt_l, True)
    nav = kw.pop("no_attn_vec", None)
    tv = t(nav, True) if nav is not None else None
    l0, l1, na, kl, ent, maps = ref.local_loss(ti, tw, cap_lens, no_attn_vec=tv, **kw)
    total = grad_w[0] * l0 + grad_w[1] * l1 + na + kl + ent
    total.backward()
    out = dict(loss0=l0.item(), loss1=l1.item(), no_attn_loss=float(na), kl_loss=float(kl), entropy_loss=float(ent),
               d_img=ti.grad.numpy(), d_txt=tw.grad.numpy())
    for i, m in enumerate(maps):
        out[f"att_{i}"] = m.detach().numpy()
    if tv is not None:
        out["d_nav"] = tv.grad.numpy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = load_reference_loss()
    GLoRIA = load_reference_model_class()

    # ---------------- small fp64 cases: full tensors stored -----------------------------------
    B, D, H, W, Lmax = 5, 48, 4, 5, 11
    img_l, txt_l, img_g, txt_g, cap_lens = gen_inputs(11, B, D, H, W, Lmax, cap_lens=[11, 9, 6, 2, 1])
    small = dict(img_l=img_l, txt_l=txt_l, img_g=img_g, txt_g=txt_g, cap_lens=np.array(cap_lens))

    # cosine_similarity + attention_fn (a1, a2)
    x1, x2 = t(img_g), t(txt_g)
    small["cos"] = ref.cosine_similarity(x1, x2).numpy()
    q = t(txt_l[1:2, :, :9]).repeat(B, 1, 1)
    wc, at = ref.attention_fn(q, t(img_l), 4.0)
    small["attn_wctx"], small["attn_map"] = wc.numpy(), at.numpy()
    nav = np.random.default_rng(5).standard_normal(D)
    wc, at = ref.attention_fn(q, t(img_l), 4.0, no_attn_vec=t(nav))
    small["nav"], small["attn_wctx_nav"], small["attn_map_nav"] = nav, wc.numpy(), at.numpy()

    # local_loss variants (a3) with autograd gradients
    for tag, kw in [("sum", dict()), ("mean", dict(agg="mean", temp1=3.0, temp2=6.0, temp3=7.0)),
                    ("reg", dict(no_attn_vec=nav, no_attn_loss_weight=0.3, attention_divergence_loss_weight=0.2,
                                 attention_entropy_loss_weight=0.1)),
                    ("ent_only", dict(attention_entropy_loss_weight=1.0, attention_divergence_loss_weight=0.5))]:
        res = run_local(ref, img_l, txt_l, cap_lens, grad_w=(1.0, 0.7), **kw)
        for k, v in res.items():
            small[f"local_{tag}_{k}"] = v

    # logits of the default case (recomputed through the reference functions, no autograd)
    sims = []
    for i in range(B):
        L = cap_lens[i]
        word = t(txt_l[i:i + 1, :, :L]).repeat(B, 1, 1)
        wc, _ = ref.attention_fn(word, t(img_l), 4.0)
        rs = ref.cosine_similarity(word.transpose(1, 2).reshape(B * L, -1), wc.transpose(1, 2).reshape(B * L, -1))
        rs = rs.view(B, L).mul(5.0).exp().sum(1, keepdim=True).log()
        sims.append(rs)
    small["local_sum_logits"] = (torch.cat(sims, 1) * 10.0).numpy()

    # global_loss (a4)
    tg, tt = t(img_g, True), t(txt_g, True)
    g0, g1 = ref.global_loss(tg, tt, temp3=10.0)
    (g0 + 0.7 * g1).backward()
    small.update(global_loss0=g0.item(), global_loss1=g1.item(), d_img_g=tg.grad.numpy(), d_txt_g=tt.grad.numpy())

    # zero word vector / zero global row edge case (eps clamp, gloria_loss.py:16,80)
    txt_z = txt_l.copy()
    txt_z[2, :, 3] = 0.0
    res = run_local(ref, img_l, txt_z, cap_lens)
    small.update(zero_word_loss0=res["loss0"], zero_word_loss1=res["loss1"], zero_word_d_img=res["d_img"],
                 zero_word_d_txt=res["d_txt"])

    # GLoRIA methods (a6-a10)
    sents = [["[CLS]"] + ["w"] * (L - 1) + ["[SEP]"] + ["[PAD]"] * (Lmax - L - 1) for L in cap_lens]
    seg = np.random.default_rng(2).random((B, 24, 30)) > 0.7
    model = make_model(GLoRIA, segmentation_loss_weight=0.5)
    ti, tw, tg, tt = t(img_l, True), t(txt_l, True), t(img_g, True), t(txt_g, True)
    loss, maps = model.calc_loss(ti, tg, tw, tt, sents, torch.tensor(seg))
    loss.backward()
    small.update(seg_labels=seg, calc_loss=loss.item(), calc_d_img_l=ti.grad.numpy(), calc_d_txt_l=tw.grad.numpy(),
                 calc_d_img_g=tg.grad.numpy(), calc_d_txt_g=tt.grad.numpy())
    model_ft = make_model(GLoRIA, local_loss_weight=0, global_loss_weight=0, segmentation_loss_weight=1.0)
    ti, tw = t(img_l, True), t(txt_l, True)
    loss, maps = model_ft.calc_loss(ti, t(img_g), tw, t(txt_g), sents, torch.tensor(seg))
    loss.backward()
    small.update(ft_loss=loss.item(), ft_d_img_l=ti.grad.numpy(), ft_d_txt_l=tw.grad.numpy())
    maps = model.get_attn_maps(t(img_l), t(txt_l), sents)
    for i, m in enumerate(maps):
        small[f"model_att_{i}"] = m.numpy()
    # rectangular zero-shot style: 5 images x 3 prompts, word slice [1:L+1]
    zl = [4, 7, 2]
    small["zs_cap_lens"] = np.array(zl)
    small["zs_local"] = model.get_local_similarities(t(img_l), t(txt_l[:3]), zl).numpy()
    small["zs_global"] = model.get_global_similarities(t(img_g), t(txt_g[:3])).numpy()
    np.savez_compressed(os.path.join(OUT, "small_fp64.npz"), **small)

    # ---------------- full-size dims (D=768, 19x19, <=97 words), B=3: fp64 + fp32 -------------
    B, D, H, W, Lmax = 3, 768, 19, 19, 97
    for scale, tag in [(1.0, "unit"), (0.05, "small")]:
        img_l, txt_l, img_g, txt_g, cap_lens = gen_inputs(7, B, D, H, W, Lmax, cap_lens=[97, 41, 5], scale=scale)
        full = dict(cap_lens=np.array(cap_lens), input_checksum=checksum(img_l, txt_l, img_g, txt_g))
        for dt, dtag in [(np.float64, "f64"), (np.float32, "f32")]:
            res = run_local(ref, img_l.astype(dt), txt_l.astype(dt), cap_lens)
            sims = []
            with torch.no_grad():
                for i in range(B):
                    L = cap_lens[i]
                    word = t(txt_l.astype(dt)[i:i + 1, :, :L]).repeat(B, 1, 1)
                    wc, _ = ref.attention_fn(word, t(img_l.astype(dt)), 4.0)
                    rs = ref.cosine_similarity(word.transpose(1, 2).reshape(B * L, -1),
                                               wc.transpose(1, 2).reshape(B * L, -1))
                    sims.append(rs.view(B, L).mul(5.0).exp().sum(1, keepdim=True).log())
            full[f"{dtag}_logits"] = (torch.cat(sims, 1) * 10.0).numpy()
            full[f"{dtag}_loss0"], full[f"{dtag}_loss1"] = res["loss0"], res["loss1"]
            full[f"{dtag}_att_1"] = res["att_1"]                       # [1, 41, 19, 19]
            full[f"{dtag}_d_img_sub"] = res["d_img"][:, ::16, ::3, ::3].copy()
            full[f"{dtag}_d_txt_sub"] = res["d_txt"][:, ::16, ::4].copy()
            full[f"{dtag}_d_img_absmax"] = np.abs(res["d_img"]).max()
            full[f"{dtag}_d_txt_absmax"] = np.abs(res["d_txt"]).max()
            full[f"{dtag}_d_img_cs"] = checksum(res["d_img"])
            full[f"{dtag}_d_txt_cs"] = checksum(res["d_txt"])
        model = make_model(GLoRIA)
        full["zs_local_f32"] = model.get_local_similarities(t(img_l.astype(np.float32)), t(txt_l.astype(np.float32)),
                                                            [c - 1 for c in cap_lens if c > 1] + [3]).numpy()
        np.savez_compressed(os.path.join(OUT, f"full_{tag}.npz"), **full)
    print("golden vectors written to", os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
