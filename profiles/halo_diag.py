"""Diagnostic for the halo (row-offset descriptor) mode: one non-zero tap at a time."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heal_b200 import ops  # noqa: E402

torch.manual_seed(0)
x = torch.randn(1, 64, 6, 128)
for planes in (1,):
    for r in range(3):
        for s in range(3):
            conv = torch.nn.Conv2d(64, 64, 3, padding=1, bias=False)
            with torch.no_grad():
                w = torch.zeros_like(conv.weight)
                w[:, :, r, s] = torch.randn(64, 64) * 0.1
                conv.weight.copy_(w)
                y = conv(x)
            pc = ops.pack_conv_tc(conv, None, False, planes=planes).to("cuda")
            o, _ = ops.conv2d_tc(ops.convert(ops.to_act(x.cuda()), "split" if planes == 2 else "bf16"), pc)
            got = ops.act_to_nchw(o).cpu()
            err = (got - y).abs().max().item()
            # is the result a shifted version of the truth?  try horizontal shifts
            best = min(((got[..., :, max(0, d):128 + min(0, d)] - y[..., :, max(0, -d):128 + min(0, -d)]).abs().max().item(), d) for d in range(-3, 4))
            print(f"BO={os.environ.get('HEAL_TC_BO','1')} planes={planes} tap(r={r},s={s}): err={err:.3e} max|y|={y.abs().max().item():.2f} best_shift={best[1]} err_at_shift={best[0]:.3e}")
