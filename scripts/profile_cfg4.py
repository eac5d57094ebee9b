"""Kernel table of the zero-shot scoring call (cfg4: 10 000 images x 25 prompts, bf16 mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gloria_nlp_project_b200 as G
from gloria_nlp_project_b200.gloria_model import GLoRIALossMixin
from torch.profiler import profile, ProfilerActivity

D, H, W = 768, 19, 19
dev = torch.device("cuda", 0)


class Model(GLoRIALossMixin):
    temp1, temp2, temp3 = 4.0, 5.0, 10.0


m = Model()
N_IMG, N_TXT, LT = 10000, 25, 18
g = torch.Generator(device="cuda").manual_seed(3)
img_l = torch.randn(N_IMG, D, H, W, device=dev, generator=g)
img_g = torch.randn(N_IMG, D, device=dev, generator=g)
txt_l = torch.randn(N_TXT, D, LT, device=dev, generator=g)
txt_g = torch.randn(N_TXT, D, device=dev, generator=g)
cl = torch.randint(4, 17, (N_TXT,), generator=torch.Generator().manual_seed(4)).tolist()
G.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16")


def zs():
    return m.get_local_similarities(img_l, txt_l, cl), m.get_global_similarities(img_g, txt_g)


zs(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    zs()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
