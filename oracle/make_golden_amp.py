"""How far the REAL reference's own mixed-precision mode is from its fp32 mode (test infrastructure; run in the build
container only -- needs /root/reference):

    python oracle/make_golden_amp.py      -> tests/golden/amp_unit.json

The reference trains under fp16 autocast (Lightning `precision: 16`), where torch runs the two `bmm`s of attention_fn
(gloria_loss.py:40,59) on fp16 operands.  On raw unit-variance 768-d features the word softmax amplifies that operand
rounding, so the reference's AMP gradients deviate from its fp32 gradients by 2-3 %.  The bf16 tensor-core mode of this
repo is gated against that number on the same seeded inputs (tests/test_gpu_bf16_parity.py): it must be at least as
close to the fp32 reference as the reference's own AMP run is.  Only scalars are stored (the gradients are 22 MB).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.make_golden import gen_inputs, load_reference_loss  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "amp_unit.json")
CASES = [dict(name="B3", B=3, seed=7, lens=[97, 41, 5]), dict(name="B16", B=16, seed=3, lens=None)]


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def main():
    ref = load_reference_loss()
    out = {"torch": torch.__version__, "reference": "gloria/loss/gloria_loss.py::local_loss, loss0 + 0.7 * loss1",
           "autocast": "torch.autocast('cpu', dtype=torch.float16)", "cases": {}}
    for c in CASES:
        img_l, txt_l, _, _, cl = gen_inputs(c["seed"], c["B"], 768, 19, 19, 97, cap_lens=c["lens"], dtype=np.float32)
        res = {}
        for tag, amp in (("fp32", None), ("fp16", torch.float16)):
            img = torch.tensor(img_l, requires_grad=True)
            txt = torch.tensor(txt_l, requires_grad=True)
            if amp is None:
                o = ref.local_loss(img, txt, cl)
            else:
                with torch.autocast("cpu", dtype=amp):
                    o = ref.local_loss(img, txt, cl)
            (o[0] + 0.7 * o[1]).float().backward()
            res[tag] = (float(o[0].detach()), float(o[1].detach()), img.grad.double().numpy(), txt.grad.double().numpy())
        a, b = res["fp16"], res["fp32"]
        out["cases"][c["name"]] = {
            "B": c["B"], "seed": c["seed"], "cap_lens": [int(v) for v in cl],
            "input_checksum": float(np.float64(img_l).sum() + np.float64(txt_l).sum()),
            "loss_fp32": [b[0], b[1]],
            "amp_loss_rel_err": [abs(a[0] - b[0]) / abs(b[0]), abs(a[1] - b[1]) / abs(b[1])],
            "amp_d_img_rel_err": rel(a[2], b[2]), "amp_d_txt_rel_err": rel(a[3], b[3]),
        }
        print(c["name"], out["cases"][c["name"]])
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
