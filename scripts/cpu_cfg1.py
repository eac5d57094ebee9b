"""SURVEY 8d cfg 1: the reference's CPU loss (oracle/gloria_oracle_torch.py: the same ATen op sequence and autograd as
gloria/loss/gloria_loss.py) at batch 16, fp32, local + global loss forward + backward, on all host cores; both
caption-length settings (97 words; U{5..97} sorted descending), median of 5.  One JSON line per setting.
Baseline infrastructure: this is the only script besides bench.py's CPU arm that executes oracle/ for timing."""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import gloria_oracle_torch as T

B, D, H, W, LW = 16, 768, 19, 19, 97
cores = os.cpu_count() or 1
torch.set_num_threads(cores)
g = torch.Generator().manual_seed(0)
img_l, txt_l = torch.randn(B, D, H, W, generator=g), torch.randn(B, D, LW, generator=g)
img_g, txt_g = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
real = sorted(torch.randint(5, 98, (B,), generator=torch.Generator().manual_seed(1)).tolist(), reverse=True)
for name, lens in (("97 words", [LW] * B), ("cap_lens U{5..97} sorted descending", real)):
    t = txt_l.clone()
    for i, L in enumerate(lens):
        t[i, :, L:] = 0
    T.loss_step(img_l, t, img_g, txt_g, lens)                      # warm-up
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        T.loss_step(img_l, t, img_g, txt_g, lens)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    print(json.dumps({"config": f"cfg1 reference CPU loss, B={B}, fp32, {name}", "cores": torch.get_num_threads(),
                      "ms_per_step_median_of_5": med * 1e3, "pairs_per_s": B / med, "torch": torch.__version__,
                      "flops_algorithmic": 12.0 * H * W * D * B * sum(lens),
                      "gflops": 12.0 * H * W * D * B * sum(lens) / med / 1e9}), flush=True)
