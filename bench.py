#!/usr/bin/env python
"""bench.py -- GLoRIA local+global loss forward+backward throughput on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload b512|b48] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic features: local loss (all B x B pairs, both cross
entropies) + global loss, forward and backward, through the package's public API (gloria_loss.local_loss /
global_loss, or distributed.sharded_loss for N > 1).  `value` = B / t_step (image-text pairs per second, whole job)
with inputs resident in HBM; `e2e` = the same with pinned HOST inputs copied in (one step ahead, on a copy stream, so
the copy overlaps the previous step's kernels) and the loss read back every step.
Workloads: b512 = BASELINE.json configs[2] at N GPUs (global batch 512 caption-sharded; N=1 is the north-star
single-GPU target), b48 = configs[1] (chexpert_pretrain_config batch).  All captions have 97 words.

`--impl reference` times the CPU restatement of the reference's loss (oracle/gloria_oracle_torch.py: same ATen op
sequence and autograd as gloria/loss/gloria_loss.py) on the host cores; /root/reference does not exist on the box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, H, W, LW = 768, 19, 19, 97
S = H * W
WORKLOADS = {"b512": 512, "b48": 48}
CPU_SAMPLE = 16          # the CPU arm times a CPU_SAMPLE x CPU_SAMPLE block of the B x B pair grid
# DRAM bytes per launch (read + write) of the dominant kernels at B=512, N=1, from the ncu captures under profiles/
TRAFFIC_B512 = {("tc_fwd_kernel", False): 2511333120 + 237689344,            # r01_traffic_B512_ncu.csv
                ("tc_bwd_pair_kernel", False): 5648215808 + 45244279296,
                ("tc_fwd_kernel", True): 3192498000 + 40386191000}            # r02_fused_train_B512_ncu_summary.txt


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="b512", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), burst=float(j["bf16_tflops"]),
                    hbm=float(j["hbm_gbs"]), src="measured (MEASURED_PEAKS.json; sustained bf16 figure: kernels are "
                    "timed inside a long step)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference loss; the checker, timed as the reported baseline)
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample_inputs(B_sample):
    import torch
    g = torch.Generator().manual_seed(0)
    img_l = torch.randn(B_sample, D, H, W, generator=g)
    txt_l = torch.randn(B_sample, D, LW, generator=g)
    img_g = torch.randn(B_sample, D, generator=g)
    txt_g = torch.randn(B_sample, D, generator=g)
    return img_l, txt_l, img_g, txt_g, [LW] * B_sample


def cpu_arm(B, steps, warmup):
    """Time `steps` passes of the reference's CPU loss at batch CPU_SAMPLE (= BASELINE.json configs[0]: batch 16, 97
    words, 361 regions, D=768, fwd+bwd) after `warmup` passes on the SAME shape.  That step is exactly a
    CPU_SAMPLE x CPU_SAMPLE block of the B x B pair grid, so the batch-B figure is the measured time x (B/16)^2
    (the reference cannot hold B = 512 at all: O(B^2 S L) autograd storage)."""
    import torch
    from oracle import gloria_oracle_torch as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp = cpu_sample_inputs(CPU_SAMPLE)
    for _ in range(warmup):
        T.loss_step(*inp)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        T.loss_step(*inp)
        times.append(time.perf_counter() - t0)
    times.sort()
    dt = times[len(times) // 2]                          # median pass
    pair_s = dt / (CPU_SAMPLE * CPU_SAMPLE)              # seconds per (image, caption) grid cell
    t_step = pair_s * B * B                              # extrapolated full step at batch B
    return dict(value=B / t_step, ms_per_step=t_step * 1e3, cores=torch.get_num_threads(), sample_s=dt,
                cfg1={"batch": CPU_SAMPLE, "ms_per_step": dt * 1e3, "pairs_per_s": CPU_SAMPLE / dt, "passes": steps,
                      "warmup_passes": warmup, "note": "measured, not extrapolated: BASELINE.json configs[0]"},
                extrapolation=f"x{(B * B) // (CPU_SAMPLE ** 2)} pair-grid cells (batch {B} = {(B // CPU_SAMPLE) ** 2} "
                              f"blocks of {CPU_SAMPLE}x{CPU_SAMPLE})",
                sample=f"{CPU_SAMPLE}x{CPU_SAMPLE} block of the {B}x{B} pair grid = one full batch-{CPU_SAMPLE} step "
                       f"(97 words, 361 regions, D=768), fwd+bwd, torch {torch.__version__} CPU, median of {steps} "
                       f"timed passes after {warmup} warm-up passes on the same shape; value extrapolated "
                       f"x{(B * B) // (CPU_SAMPLE ** 2)}")


def run_reference(args):
    B = WORKLOADS[args.workload]
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps = max(5, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    r = cpu_arm(B, steps, warm)
    line = {"impl": "reference", "metric": "image-text pairs/s, GLoRIA local+global loss fwd+bwd", "value": r["value"],
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: global batch {B}, 97 words, 361 regions, D=768 (CPU sample)"},
            "cpu_baseline": {"value": r["value"], "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"], "cfg1_measured": r["cfg1"], "extrapolation": r["extrapolation"]},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# pre-timing parity check (every N): the same public API the timed steps call, at batch 32, against the numpy oracle
# ------------------------------------------------------------------------------------------------------------------
PARITY_B = 32


def parity_check(world, rank, dev, loss_of_factory):
    """Loss and gradients of a batch-32 step (ragged captions) computed through the SAME code path the timed steps
    take at this N (single-GPU API, or distributed.sharded_loss over NCCL), compared with the numpy oracle (fp64, run
    on rank 0 only; the checker, never the thing measured).  Unit-variance features rounded to 16-bit-representable
    values so that kernel and oracle see identical operands.  Gates: loss 2e-3, gradients 1e-2."""
    import numpy as np
    import torch
    import torch.distributed as dist
    B = PARITY_B
    if B % world:
        return {"skipped": f"batch {B} not divisible by {world} ranks"}
    n = B // world
    rng = np.random.default_rng(4242)
    img_l = torch.from_numpy(rng.standard_normal((B, D, H, W), dtype=np.float32)).to(torch.bfloat16).float()
    txt_l = torch.from_numpy(rng.standard_normal((B, D, LW), dtype=np.float32)).to(torch.bfloat16).float()
    img_g = torch.from_numpy(rng.standard_normal((B, D), dtype=np.float32))
    txt_g = torch.from_numpy(rng.standard_normal((B, D), dtype=np.float32))
    lens = [int(v) for v in rng.integers(5, LW + 1, size=B)]
    for r in range(world):
        lens[r * n] = LW                                  # every shard holds one full-length caption
    for i, L in enumerate(lens):
        txt_l[i, :, L:] = 0
    sl = slice(rank * n, (rank + 1) * n)
    t = {k: v[sl].to(dev).requires_grad_(True) for k, v in
         (("img_l", img_l), ("txt_l", txt_l), ("img_g", img_g), ("txt_g", txt_g))}
    loss = loss_of_factory(lens[sl])(t)
    loss.backward()
    torch.cuda.synchronize()
    got = float(loss.detach())
    ref = {}
    if rank == 0:
        from oracle import gloria_oracle as O
        t0 = time.perf_counter()
        i64, w64, g64, y64 = (a.numpy().astype(np.float64) for a in (img_l, txt_l, img_g, txt_g))
        o = O.local_loss(i64, w64, lens)
        og = O.global_loss(g64, y64)
        ref["loss"] = float(o[0] + o[1] + og[0] + og[1])
        d_img, d_txt = O.local_loss_bwd(i64, w64, lens)
        d_ig, d_tg = O.global_loss_bwd(g64, y64)
        ref["grads"] = [torch.tensor(a, dtype=torch.float32) for a in (d_img, d_txt, d_ig, d_tg)]
        ref["oracle_s"] = time.perf_counter() - t0
    names = ("img_l", "txt_l", "img_g", "txt_g")
    shapes = ((B, D, H, W), (B, D, LW), (B, D), (B, D))
    errs = {}
    for k, (name, shp) in enumerate(zip(names, shapes)):
        full = ref["grads"][k].to(dev) if rank == 0 else torch.empty(shp, dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(full, 0)
        num = (t[name].grad.float() - full[sl]).abs().max().reshape(1)
        if world > 1:
            dist.all_reduce(num, op=dist.ReduceOp.MAX)
        errs["d_" + name] = float(num.item() / full.abs().max().item())
    if rank != 0:
        return None
    loss_err = abs(got - ref["loss"]) / abs(ref["loss"])
    ok = loss_err < 2e-3 and all(v < 1e-2 for v in errs.values())
    return {"batch": B, "n_ranks": world, "path": "distributed.sharded_loss (NCCL)" if world > 1 else
            "gloria_loss.local_loss + global_loss", "loss": got, "loss_oracle": ref["loss"], "loss_rel_err": loss_err,
            "grad_rel_err_maxnorm": errs, "tol": {"loss": 2e-3, "grad": 1e-2}, "ok": bool(ok),
            "oracle": "oracle/gloria_oracle.py (numpy fp64, rank 0), %.1f s" % ref["oracle_s"],
            "inputs": "seeded unit-variance features rounded to bf16-representable values, cap_lens ~ U{5..97}"}


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    import gloria_nlp_project_b200 as G
    from gloria_nlp_project_b200 import _lib, distributed, gloria_loss
    lib = _lib.lib()                                       # raises if the CUDA library is missing: no fallback
    G.set_precision(args.precision)

    def loss_factory(lens):
        def f(t):
            if world == 1:
                l0, l1, _, _, _, _ = gloria_loss.local_loss(t["img_l"], t["txt_l"], lens)
                g0, g1 = gloria_loss.global_loss(t["img_g"], t["txt_g"])
            else:
                l0, l1, g0, g1 = distributed.sharded_loss(t["img_l"], t["txt_l"], t["img_g"], t["txt_g"], lens)
            return l0 + l1 + g0 + g1
        return f

    parity = None
    if not args.no_parity_check and args.precision == "bf16":
        parity = parity_check(world, rank, dev, loss_factory)

    B = WORKLOADS[args.workload]
    if B % world:
        raise SystemExit(f"global batch {B} is not divisible by {world} ranks")
    n = B // world
    gen = torch.Generator().manual_seed(0)                 # same global tensors on every rank, sliced per rank
    host = {}
    for name, shape in (("img_l", (B, D, H, W)), ("txt_l", (B, D, LW)), ("img_g", (B, D)), ("txt_g", (B, D))):
        full = torch.randn(shape, generator=gen)
        host[name] = full[rank * n:(rank + 1) * n].contiguous().pin_memory()
        del full
    # caption lengths as the drop-in text encoder hands them over (text_model.aggregate_tokens -> sents.cap_lens): a device
    # tensor, so the step has no host round trip for them (gloria_loss.DeviceCapLens)
    cap_lens = gloria_loss.DeviceCapLens(torch.full((n,), LW, dtype=torch.int32, device=dev))
    names = ("img_l", "txt_l", "img_g", "txt_g")
    h2d = sum(host[k].numel() * host[k].element_size() for k in names)

    loss_of = loss_factory(cap_lens)

    resident = {k: host[k].to(dev).requires_grad_(True) for k in names}

    def step_resident():
        for v in resident.values():
            v.grad = None
        loss = loss_of(resident)
        loss.backward()
        return loss

    copy_stream = torch.cuda.Stream(device=dev)
    h2d_events = []                                         # (start, end) of every step's input copy on the copy stream

    def issue_h2d():
        """Issue this step's host->device copies on the copy stream (pinned memory, asynchronous)."""
        with torch.cuda.stream(copy_stream):
            e_a, ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_a.record(copy_stream)
            t = {k: host[k].to(dev, non_blocking=True) for k in names}
            ev.record(copy_stream)
        h2d_events.append((e_a, ev))
        return t, ev

    # Where the next step's input copy is enqueued.  Sharded over many ranks the step is short and starts with the gather
    # phase, under which a concurrent 90 MB copy per rank costs 2.9 ms per step (8 ranks; 0.5 ms between forward and
    # backward: scripts/profile_e2e_sharded.py); on one GPU the copy is 724 MB and is cheapest under the fused kernel at
    # the top of the step (0.75 ms against 2.0 ms under the HBM-bound backward).
    prefetch_mid = world >= 4
    readback = {"n": 0, "buf": [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)],
                "ev": [torch.cuda.Event(), torch.cuda.Event()]}

    def step_e2e(cur, prefetch_next):
        """One end-to-end step: inputs come from pinned host memory (their copy was issued one step earlier and
        overlaps the previous step's kernels), the loss value is read back to the host."""
        t, ev = cur
        nxt = issue_h2d() if (prefetch_next and not prefetch_mid) else None
        torch.cuda.current_stream().wait_event(ev)
        for v in t.values():
            v.record_stream(torch.cuda.current_stream())
            v.requires_grad_(True)
        loss = loss_of(t)
        if prefetch_next and prefetch_mid:
            nxt = issue_h2d()
        loss.backward()
        # device->host read of the step's result: an asynchronous copy into pinned memory, issued every step and consumed
        # one step later (a blocking float(loss) here would expose the host's launch latency of the NEXT step, which at
        # 8 ranks is a quarter of the 14 ms step)
        slot = readback["n"] & 1
        readback["buf"][slot].copy_(loss.detach().reshape(1), non_blocking=True)
        readback["ev"][slot].record()
        readback["n"] += 1
        val = None
        if readback["n"] > 1:
            readback["ev"][slot ^ 1].synchronize()
            val = float(readback["buf"][slot ^ 1][0])
        return val, nxt

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel timers: caller-owned events recorded by the library around its dominant kernels
    slots = {"tc_fwd_kernel": 0, "tc_bwd_pair_kernel": 1, "bwd_accum_gemms": 2}
    evs = {k: [] for k in slots}

    def arm_timers(i):
        for k, s in slots.items():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); b.record()                          # materialise the handles; re-recorded by the library
            evs[k].append((a, b))
            lib.gloria_b200_set_timer_events(s, a.cuda_event, b.cuda_event)

    # inputs smaller than L2 (b48): write a 256 MB buffer between the timed iterations; each step then has its own events.
    # Allocated (and used) before the warm-up so that the caching allocator is in its steady state when timing starts.
    small = B * D * S * 4 <= 126e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None
    for i in range(args.warmup):
        if small:
            flush_buf.fill_(i & 1)
        step_resident()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    lib.gloria_b200_launch_count(1)
    sync()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_step = []
    e0.record()
    for i in range(args.steps):
        if small:
            flush_buf.fill_(i & 1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        arm_timers(i)
        last = step_resident()
        if small:
            b.record()
            per_step.append((a, b))
    e1.record()
    sync()
    t_wall1 = time.time()
    launches = int(lib.gloria_b200_launch_count(1))
    for s in slots.values():
        lib.gloria_b200_set_timer_events(s, None, None)
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in per_step) / args.steps if small
                        else e0.elapsed_time(e1) / args.steps)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    loss_val = float(last.detach())

    kt = {}
    for k, lst in evs.items():
        v = [a.elapsed_time(b) for a, b in lst]
        v = [x for x in v if x > 1e-3]                      # slots that never fired keep their back-to-back records
        kt[k] = sum(v) / len(v) if v else 0.0

    # ---- end to end: pinned host inputs in, loss out, every step
    cur = issue_h2d()
    for _ in range(max(1, args.warmup // 2)):
        _, cur = step_e2e(cur, True)
    sync()
    e0.record()
    for i in range(args.steps):
        _, cur = step_e2e(cur, True)                         # every timed step issues one full H2D copy
    e1.record()
    sync()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    # how long the input copies themselves took (copy-stream events of the timed steps, max over ranks): when this is close
    # to the e2e step time, the end-to-end number is bound by the host's H2D bandwidth, not by the kernels
    cp = [a.elapsed_time(b) for a, b in h2d_events[-args.steps - 1:-1]] or [0.0]
    ms_h2d = max_over_ranks(sum(cp) / len(cp))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    pairs_grid = B * n                                      # (image, caption) cells per rank
    f_fwd = 4.0 * S * D * pairs_grid * LW                   # algorithmic FLOPs, SURVEY 8d (per rank)
    f_step = 3.0 * f_fwd
    # Algorithmic FLOPs (12 S D L per pair and step, no recompute credit) attributed per timed kernel.
    # Fused training path (timer slot 0 = tc_bwd_pair_kernel<FUSED>): forward 4SDL + the backward's dC^T R 2SDL in the
    # fused kernel, the remaining 6SDL (dW, dR, dC A) in the accumulation phase.  Recompute path: forward kernel 4SDL,
    # pair kernel + accumulation 8SDL.
    fused = kt["tc_bwd_pair_kernel"] < 0.05 * max(kt["tc_fwd_kernel"], 1e-9)
    timed = max(kt.values()) > 0.02                          # fp32 mode: the SIMT kernels carry no timer hooks
    if not timed:
        names_, kflops, kt_rep = {}, {}, {}
    elif fused:
        names_ = {"tc_fwd_kernel": "tc_fused_train_kernel (tc_bwd_pair_kernel<LPAD, FUSED=true>)",
                  "bwd_accum_gemms": "backward accumulation phase (scale by dsim + cuBLAS GEMMs)"}
        kflops = {"tc_fwd_kernel": 1.5 * f_fwd, "bwd_accum_gemms": 1.5 * f_fwd}
        kt_rep = {names_[k]: round(kt[k], 4) for k in names_}
    else:
        names_ = {"tc_fwd_kernel": "tc_fwd_kernel", "tc_bwd_pair_kernel": "tc_bwd_pair_kernel",
                  "bwd_accum_gemms": "backward accumulation phase"}
        kflops = {"tc_fwd_kernel": f_fwd, "tc_bwd_pair_kernel": 1.0 * f_fwd, "bwd_accum_gemms": 1.0 * f_fwd}
        kt_rep = {names_[k]: round(kt[k], 4) for k in names_}
    cand = [k for k in kflops if kt[k] > 0 and k != "bwd_accum_gemms"] or [k for k in kflops if kt[k] > 0]
    dom = max(cand, key=lambda k: kt[k]) if cand else None
    roofline = None
    if not timed:
        # fp32 mode (tensor-core path, tc_f32.cu): the whole step against the sustained bf16 tensor peak.  Every fp32 product
        # is executed as 6 (scores) or 3 (all other GEMMs) bf16 piece products: 9 piece GEMMs forward + 12 backward, each
        # 2 S D L per pair, against the 6 algorithmic GEMMs -- `frac` counts algorithmic FLOPs, `frac_executed` what ran
        f_exec = 21.0 * f_fwd / 2.0
        roofline = {"bound": "tensor",
                    "kernel": "fp32 mode: split-precision acc_gemm_kernel (6 / 3 bf16 piece products per fp32 product) + "
                              "streaming softmax / cosine kernels (whole step)",
                    "achieved": f_step / (ms * 1e-3) / 1e12, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": f_step / (ms * 1e-3) / 1e12 / pk["tflops"], "traffic": None, "peak_source": pk["src"],
                    "executed_piece_flops_per_step": f_exec,
                    "frac_executed": f_exec / (ms * 1e-3) / 1e12 / pk["tflops"]}
    elif dom:
        ach = kflops[dom] / (kt[dom] * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/ (B=512, N=1 captures), else null
        traffic = None
        if args.workload == "b512" and world == 1:
            traffic = TRAFFIC_B512.get((dom, fused))
        roofline = {"bound": "tensor", "kernel": names_[dom], "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": ach / pk["tflops"], "traffic": traffic, "peak_source": pk["src"],
                    "kernel_ms": kt_rep, "algorithmic_flops_per_launch": kflops[dom],
                    "step_achieved": f_step / (ms * 1e-3) / 1e12,
                    "step_frac": f_step / (ms * 1e-3) / 1e12 / pk["tflops"],
                    "step_frac_of_burst": f_step / (ms * 1e-3) / 1e12 / pk["burst"]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(B, 5, 1)
        cpu = {"value": r["value"], "unit": "pairs/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "cfg1_measured": r["cfg1"], "extrapolation": r["extrapolation"]}

    line = {"metric": "image-text pairs/s, GLoRIA local+global loss fwd+bwd", "value": B / (ms * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.workload}: global batch {B} ({n} captions x {B} images per rank), 97 words, "
                                   "361 regions (19x19), D=768, temps 4/5/10, local+global loss",
                       "parallelism": f"caption-sharded x{world}" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (region features %.0f MB per rank)" % (B * D * S * 4 / 1e6)
                             if B * D * S * 4 > 126e6 else "inputs smaller than L2: a 256 MB buffer is written between "
                                                           "timed iterations (outside the per-step events)",
                       "cap_lens": "device tensor (DeviceCapLens), all captions 97 words",
                       "loss": loss_val},
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                    "h2d_copy_ms_per_step": ms_h2d, "h2d_GBps_per_rank": h2d / max(ms_h2d, 1e-9) / 1e6,
                    "readback": "loss copied to pinned host memory every step (asynchronous), read by the host one step later",
                    "prefetch": "next step's inputs enqueued " + ("between forward and backward" if prefetch_mid
                                                                   else "at the top of the step") + " on a copy stream"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity_check": parity}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
