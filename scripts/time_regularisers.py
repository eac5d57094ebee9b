"""Scratch timing of the regulariser training config (no_attn_vec + no-attn / KL / entropy terms, submit_job.sh:15 of
the reference) through the public API: lean forward kernel (sim + word-mean attention) + recompute backward.

usage: python scripts/time_regularisers.py [B]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gloria_nlp_project_b200 as g
from gloria_nlp_project_b200 import gloria_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_it = int(os.environ.get("N_IT", "3"))
g.set_precision("bf16")
gen = torch.Generator(device="cuda").manual_seed(0)
img = (torch.randn(B, 768, 19, 19, device="cuda", generator=gen) * 0.05).requires_grad_()
txt = (torch.randn(B, 768, 97, device="cuda", generator=gen) * 0.05).requires_grad_()
nav = (torch.randn(768, device="cuda", generator=gen) * 0.05).requires_grad_()
lens = [97] * B


def step():
    l0, l1, na, kl, ent, _ = gloria_loss.local_loss(img, txt, lens, no_attn_vec=nav, no_attn_loss_weight=1.0,
                                                    attention_divergence_loss_weight=1.0,
                                                    attention_entropy_loss_weight=1.0)
    (l0 + l1 + na + kl + ent).backward()
    img.grad = txt.grad = nav.grad = None


for _ in range(2):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(n_it):
    step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / n_it
print(f"regulariser config B={B}: {ms:.1f} ms/step -> {B / ms * 1e3:.0f} pairs/s "
      f"(peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)")
