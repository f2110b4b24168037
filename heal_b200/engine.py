"""Execution helpers shared by the mirror modules: parameter folding/packing cache and conv runners.

Parameters stay ordinary nn.Parameters with the reference's state-dict key names
(SURVEY.md 8b "Parameters / ownership"); what the kernels consume is a *derived* cache (BN folded in
fp64, weights re-laid out for the kernels) that is rebuilt whenever a source tensor's version counter
changes (load_state_dict / optimizer step) or the module moves device.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import ops

_CACHE = weakref.WeakKeyDictionary()


def _sig(*mods):
    sig = []
    for m in mods:
        if m is None:
            sig.append(None)
            continue
        for t in list(m.parameters(recurse=False)) + list(m.buffers(recurse=False)):
            sig.append((t.data_ptr(), t._version, str(t.device)))
    return tuple(sig)


def packed(conv: nn.Module, bn: Optional[nn.Module], relu: bool, extra_pad: int = 0) -> ops.PackedConv:
    """Folded + packed weights for `conv` (+ eval `bn`) on conv.weight's device, cached."""
    sig = (_sig(conv, bn), relu, extra_pad)
    hit = _CACHE.get(conv)
    if hit is not None and hit[0] == sig:
        return hit[1]
    if isinstance(conv, nn.ConvTranspose2d):
        pc = ops.pack_deconv(conv, bn, relu)
    else:
        pc = ops.pack_conv(conv, bn, relu, extra_pad)
    pc.to(conv.weight.device)
    _CACHE[conv] = (sig, pc)
    return pc


def require_eval(module: nn.Module):
    if module.training:
        raise NotImplementedError(
            f"{type(module).__name__}: the heal_b200 kernels implement the inference path "
            "(model.eval(), torch.no_grad()); autograd through them is a 'next' row (SURVEY.md 8f-4).")


def conv_bn_act(x, conv, bn=None, relu=False, residual=None, out=None, out_coffset=0, extra_pad=0):
    """x / residual / out: NHWC buffers.  conv(+bn)(+residual)(+relu) in one kernel."""
    return ops.conv2d(x, packed(conv, bn, relu, extra_pad), residual=residual, out=out, out_coffset=out_coffset)
