"""cfg4 (zero-shot 10 000 x 25, bf16) and cfg5 (attention fine-tune step, B = 30 / 32) alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200.gloria_model import GLoRIALossMixin

D, H, W, LW = 768, 19, 19, 97
dev = torch.device("cuda", 0)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


class Model(GLoRIALossMixin):
    temp1, temp2, temp3 = 4.0, 5.0, 10.0
    local_loss_weight = global_loss_weight = 0.0
    segmentation_loss_weight = 1.0
    no_attn_vec = None
    no_attn_loss_weight = attention_divergence_loss_weight = attention_entropy_loss_weight = None


def timed(fn, iters, warmup=3, flush=True):
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


m = Model()
G.set_precision("bf16")
if "cfg4" in sys.argv or len(sys.argv) == 1:
    g = torch.Generator(device="cuda").manual_seed(3)
    img_l = torch.randn(10000, D, H, W, device=dev, generator=g); img_g = torch.randn(10000, D, device=dev, generator=g)
    txt_l = torch.randn(25, D, 18, device=dev, generator=g); txt_g = torch.randn(25, D, device=dev, generator=g)
    cl = torch.randint(4, 17, (25,), generator=torch.Generator().manual_seed(4)).tolist()
    ms = timed(lambda: (m.get_local_similarities(img_l, txt_l, cl), m.get_global_similarities(img_g, txt_g)), 5, 2, flush=False)
    print(json.dumps({"cfg4_bf16_ms": ms}), flush=True)
    del img_l
if "cfg5" in sys.argv or len(sys.argv) == 1:
    for B in (30, 32):
        g = torch.Generator(device="cuda").manual_seed(5)
        t = [torch.randn(s, device=dev, generator=g).requires_grad_(True) for s in ((B, D, H, W), (B, D, LW), (B, D), (B, D))]
        lens = torch.randint(5, 98, (B,), generator=torch.Generator().manual_seed(1)).tolist()
        sents = [["w"] * (L - 1) for L in lens]
        seg = torch.rand(B, 224, 224, device=dev, generator=torch.Generator(device="cuda").manual_seed(2)) > 0.7

        def ft():
            for v in t:
                v.grad = None
            loss, maps = m.calc_loss(t[0], t[2], t[1], t[3], sents, segmentation_labels=seg)
            loss.backward()
        print(json.dumps({"cfg5_B": B, "ms_per_step": timed(ft, 20, 5)}), flush=True)
