"""Timings of the BASELINE.json / SURVEY 8d configurations that bench.py's headline line does not cover.

    python scripts/bench_configs.py > profiles/rNN_configs.jsonl        (one JSON line per configuration, 1 GPU)

cfg2  B=48 pretraining step, bf16 and fp32 modes (local + global loss, fwd + bwd)
cfg3' B=512 with realistic caption lengths U{5..97} (sorted descending like the collate fn) and with features x0.05
cfg4  zero-shot scoring: get_local_similarities + get_global_similarities, 10 000 images x 25 prompts, forward only
cfg5  attention fine-tune: calc_loss with both contrastive weights 0 and segmentation labels, B=30 / 32, fwd + bwd
Device time by CUDA events after warm-up; between timed iterations a 256 MB buffer is written to flush L2 for the
configurations whose inputs fit in L2 (the flush is outside the timed events).
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200 import gloria_loss
from gloria_nlp_project_b200.gloria_model import GLoRIALossMixin

D, H, W, LW = 768, 19, 19, 97
S = H * W
dev = torch.device("cuda", 0)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters, warmup=3, flush=False):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def feats(B, seed=0, scale=1.0, lens=None):
    g = torch.Generator(device="cuda").manual_seed(seed)
    img_l = torch.randn(B, D, H, W, device=dev, generator=g) * scale
    txt_l = torch.randn(B, D, LW, device=dev, generator=g) * scale
    img_g = torch.randn(B, D, device=dev, generator=g) * scale
    txt_g = torch.randn(B, D, device=dev, generator=g) * scale
    lens = lens or [LW] * B
    for i, L in enumerate(lens):
        txt_l[i, :, L:] = 0
    return [t.requires_grad_(True) for t in (img_l, txt_l, img_g, txt_g)], lens


def train_step(t, lens):
    for v in t:
        v.grad = None
    l0, l1, *_ = gloria_loss.local_loss(t[0], t[1], lens)
    g0, g1 = gloria_loss.global_loss(t[2], t[3])
    (l0 + l1 + g0 + g1).backward()


def emit(**kw):
    print(json.dumps(kw), flush=True)


def flops(B_i, lens):
    return 12.0 * S * D * B_i * sum(lens)


class Model(GLoRIALossMixin):
    temp1, temp2, temp3 = 4.0, 5.0, 10.0
    local_loss_weight = global_loss_weight = 0.0
    segmentation_loss_weight = 1.0
    no_attn_vec = None
    no_attn_loss_weight = attention_divergence_loss_weight = attention_entropy_loss_weight = None


def main():
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    peaks = json.load(open(pk)) if os.path.exists(pk) else {}
    tf_peak, hbm_peak = peaks.get("bf16_tflops_sustained", 1400.0), peaks.get("hbm_gbs", 6650.0)
    # ---- cfg2
    for prec in ("bf16", "fp32"):
        G.set_precision(prec)
        t, lens = feats(48)
        ms = timed(lambda: train_step(t, lens), 20, 5, flush=True)
        emit(config="cfg2 chexpert_pretrain B=48 loss step fwd+bwd", precision=prec, ms_per_step=ms, pairs_per_s=48 / ms * 1e3,
             tflops_algorithmic=flops(48, lens) / ms / 1e9, l2="flushed between iterations")
    G.set_precision("bf16")
    # ---- cfg3 variants at B=512
    gl = torch.Generator().manual_seed(1)
    real = sorted(torch.randint(5, 98, (512,), generator=gl).tolist(), reverse=True)
    for name, scale, lens in (("B=512, 97 words, unit-variance features", 1.0, None),
                              ("B=512, 97 words, features x0.05 (scores O(1))", 0.05, None),
                              ("B=512, cap_lens U{5..97} sorted descending", 1.0, real)):
        t, lens = feats(512, scale=scale, lens=lens)
        ms = timed(lambda: train_step(t, lens), 5, 2)
        fl = flops(512, lens)
        emit(config="cfg3 " + name, precision="bf16", ms_per_step=ms, pairs_per_s=512 / ms * 1e3,
             tflops_algorithmic=fl / ms / 1e9, frac_of_sustained_bf16_peak=fl / ms / 1e9 / tf_peak,
             mean_cap_len=sum(lens) / len(lens))
        del t
    torch.cuda.empty_cache()
    # ---- cfg4 zero-shot
    m = Model()
    N_IMG, N_TXT, LT = 10000, 25, 18
    g = torch.Generator(device="cuda").manual_seed(3)
    img_l = torch.randn(N_IMG, D, H, W, device=dev, generator=g)
    img_g = torch.randn(N_IMG, D, device=dev, generator=g)
    txt_l = torch.randn(N_TXT, D, LT, device=dev, generator=g)
    txt_g = torch.randn(N_TXT, D, device=dev, generator=g)
    cl = torch.randint(4, 17, (N_TXT,), generator=torch.Generator().manual_seed(4)).tolist()
    for prec in ("bf16", "fp32"):
        G.set_precision(prec)
        out = {}

        def zs():
            out["l"] = m.get_local_similarities(img_l, txt_l, cl)
            out["g"] = m.get_global_similarities(img_g, txt_g)
        ms = timed(zs, 3, 1)
        in_bytes = img_l.numel() * 4
        emit(config="cfg4 zero-shot 10000 images x 25 prompts (get_local_similarities + get_global_similarities, CPU results)",
             precision=prec, ms=ms, images_per_s=N_IMG / ms * 1e3, region_feature_GBps=in_bytes / ms / 1e6,
             frac_of_hbm_peak=in_bytes / ms / 1e6 / hbm_peak, out_shape=list(out["l"].shape))
    del img_l
    torch.cuda.empty_cache()
    # ---- cfg5 attention fine-tune
    G.set_precision("bf16")
    for B in (30, 32):
        t, lens = feats(B, seed=5)
        sents = [["w"] * (L - 1) for L in lens]
        seg = torch.rand(B, 224, 224, device=dev, generator=torch.Generator(device="cuda").manual_seed(2)) > 0.7

        def ft():
            for v in t:
                v.grad = None
            loss, maps = m.calc_loss(t[0], t[2], t[1], t[3], sents, segmentation_labels=seg)
            loss.backward()
        ms = timed(ft, 20, 5, flush=True)
        emit(config=f"cfg5 imagenome_attn_finetune B={B}: calc_loss (contrastive weights 0, supervised attention) fwd+bwd",
             precision="fp32 diagonal kernels", ms_per_step=ms, pairs_touched=B, l2="flushed between iterations")


def text_side():
    """word-piece aggregation of the text encoder (B=48 / 512 captions x 4 layers x 97 tokens x 768)."""
    from gloria_nlp_project_b200 import text_model
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    hbm_peak = (json.load(open(pk)) if os.path.exists(pk) else {}).get("hbm_gbs", 6650.0)
    vocab = ["[PAD]", "[CLS]", "[SEP]"] + [f"w{i}" for i in range(1000)] + [f"##p{i}" for i in range(400)]
    table = text_model.VocabTable(dict(enumerate(vocab)))
    g = torch.Generator().manual_seed(7)
    for B in (48, 512):
        ids = torch.randint(3, len(vocab), (B, 97), generator=g)
        ids[:, 0] = 1
        for b in range(B):
            ids[b, int(torch.randint(5, 97, (1,), generator=g))] = 2
        emb = torch.randn(B, 4, 97, 768, device=dev)
        ids_dev = ids.to(dev)
        ms_all = timed(lambda: text_model.aggregate_tokens(emb, ids, table), 10, 3, flush=True)
        wr, tw, _ = text_model.ops.word_ranges(ids_dev, table.is_continuation(dev), table.sep_id)
        ms_k = timed(lambda: text_model.ops.aggregate_tokens(emb, wr, tw), 10, 3, flush=True)
        nbytes = 2 * emb.numel() * 4
        emit(config=f"8f-3 aggregate_tokens B={B} x 4 layers x 97 tokens x 768 fp32", ms_call_incl_host_sentences=ms_all,
             ms_kernel=ms_k, algorithmic_GB=nbytes / 1e9, kernel_GBps=nbytes / ms_k / 1e6,
             frac_of_hbm_peak=nbytes / ms_k / 1e6 / hbm_peak, l2="flushed between iterations")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "text":
        text_side()
    else:
        main()
        text_side()
