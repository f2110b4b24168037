"""CPU-only checks of the C-ABI boundary: the library loads and exports every symbol include/heal_b200.h
declares (no compute calls without a GPU), and the host-side packing logic."""
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "heal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(heal_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from heal_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/heal_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in heal_b200/_lib.py"
    assert _lib.lib.heal_abi_version() == 3
    assert _lib.lib.heal_launch_count() == 0


def test_product_path_has_no_cpu_fallback():
    from heal_b200 import ops
    import pytest
    with pytest.raises(RuntimeError):
        ops.to_act(torch.zeros(1, 4, 2, 2))
    with pytest.raises(RuntimeError):
        ops.mean_vfe(torch.zeros(3, 5, 4), torch.ones(3, dtype=torch.int32))


def test_bn_fold_and_packing():
    from heal_b200 import ops
    g = torch.Generator().manual_seed(0)
    conv = torch.nn.Conv2d(64, 20, 3, padding=1, bias=True)
    bn = torch.nn.BatchNorm2d(20, eps=1e-3).eval()
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(20, generator=g))
        bn.running_var.copy_(torch.rand(20, generator=g) + 0.5)
        bn.weight.copy_(torch.rand(20, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(20, generator=g))
    x = torch.randn(2, 64, 5, 7, generator=g)
    with torch.no_grad():
        y = bn(conv(x))
    pc = ops.pack_conv(conv, bn, relu=False)
    w = pc.weight[..., :20].permute(3, 2, 0, 1)                      # back to (Cout,Cin,kh,kw)
    y2 = torch.nn.functional.conv2d(x, w, pc.bias, padding=1)
    torch.testing.assert_close(y2, y, rtol=1e-4, atol=1e-4)
    tc = ops.pack_conv_tc(conv, bn, relu=False, planes=2)
    assert tc.coutp == 32 and tc.w.shape == (2, 9 * 32, 64) and tc.w.dtype == torch.bfloat16
    wm = ops.merge_bf16(tc.w).view(9, 32, 64)[:, :20].permute(1, 2, 0).reshape(20, 64, 3, 3)
    torch.testing.assert_close(wm, w.contiguous(), rtol=0, atol=2e-5)
    assert (ops.merge_bf16(tc.w).view(9, 32, 64)[:, 20:] == 0).all()


def test_split_bf16_roundtrip_precision():
    from heal_b200 import ops
    x = torch.randn(10000) * 100
    s = ops.split_bf16(x, 2)
    rel = ((ops.merge_bf16(s) - x).abs() / x.abs().clamp_min(1e-6)).max().item()
    assert rel < 2 ** -15
