"""Tensor-level wrappers over the C ABI (include/heal_b200.h).

PyTorch is used here only for device memory, streams and shapes: every wrapper passes raw device
pointers + the current CUDA stream to libheal_b200.so.  There is no eager / CPU fallback: a tensor
that is not on a CUDA device raises.

Layout convention: feature maps are exchanged as *logically* (N,C,H,W) torch tensors that are
*physically* channels-last, i.e. `x.permute(0,2,3,1)` is a contiguous (N,H,W,C) buffer.  This keeps
the reference's tensor shapes at every module boundary at zero cost.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence

import numpy as np
import torch

from ._lib import lib, check

_vp = ctypes.c_void_p


def _p(t: Optional[torch.Tensor]):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("heal_b200 ops need CUDA tensors (no CPU fallback in the product path)")


def _host_f32(vals: Sequence[float]):
    arr = (ctypes.c_float * len(vals))(*[float(np.float32(v)) for v in vals])
    return arr


def _host_i32(vals: Sequence[int]):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


_WS = {}

# bench.py instrumentation: when PROFILE is a list, every wrapper brackets its C-ABI call with CUDA events
# on the launching stream and appends (kernel family, algorithmic flops-or-bytes, start, end).
PROFILE = None


class _Prof:
    def __init__(self, name, work):
        self.name, self.work = name, work

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.append((self.name, float(self.work), self.a, self.b))
        return False


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.type, dev.index)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


# ------------------------------------------------------------------------------------------------
# layout helpers
# ------------------------------------------------------------------------------------------------
def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """(N,C,H,W) logical tensor -> contiguous (N,H,W,C) buffer (view if already channels-last)."""
    assert x.dim() == 4
    xp = x.permute(0, 2, 3, 1)
    return xp if xp.is_contiguous() else xp.contiguous()


def from_nhwc(buf: torch.Tensor) -> torch.Tensor:
    """contiguous (N,H,W,C) buffer -> logical (N,C,H,W) view (channels-last strides)."""
    return buf.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------
# voxelization
# ------------------------------------------------------------------------------------------------
def grid_size_of(lidar_range, voxel_size):
    """round((max-min)/vs) as sp_voxel_preprocessor.py:41-43 does (fp64 numpy)."""
    g = (np.array(lidar_range[3:6]) - np.array(lidar_range[0:3])) / np.array(voxel_size)
    return np.round(g).astype(np.int64)


def voxelize(points: torch.Tensor, agent_offsets: torch.Tensor, lidar_range, voxel_size,
             max_points_per_voxel: int, max_voxels: int, capacity: Optional[int] = None):
    """GPU SpVoxelPreprocessor.preprocess + collate for all agents of a scene.

    points (P,4) f32 cuda; agent_offsets (A+1) i32 cuda.  Returns (voxels (cap,T,4), coords (cap,4) i32
    [b,z,y,x], num_points (cap) i32, num_voxels (1+A) i32 device tensor).  Rows >= num_voxels[0] are
    undefined; use `trim_voxels` for the exact-size reference-shaped dict (one host sync)."""
    _need_cuda(points, agent_offsets)
    assert points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] == 4 and points.is_contiguous()
    assert agent_offsets.dtype == torch.int32
    A = agent_offsets.numel() - 1
    P = points.shape[0]
    grid = grid_size_of(lidar_range, voxel_size)
    cap = int(capacity) if capacity is not None else max(1, min(P, A * max_voxels))
    dev = points.device
    T = int(max_points_per_voxel)
    voxels = torch.empty((cap, T, 4), dtype=torch.float32, device=dev)
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    npts = torch.empty((cap,), dtype=torch.int32, device=dev)
    nvox = torch.zeros((1 + A,), dtype=torch.int32, device=dev)
    ws_bytes = lib.heal_voxelize_workspace(P, cap, A)
    ws = _workspace(dev, ws_bytes)
    with _Prof("voxelize", 0):
        rc = lib.heal_voxelize(_p(points), _p(agent_offsets), A, P,
                               _host_f32(lidar_range[0:3]), _host_f32(voxel_size), _host_i32(grid),
                               T, int(max_voxels), cap, _p(voxels), _p(coords), _p(npts), _p(nvox),
                               _p(ws), ws.numel(), _stream())
    check(rc, "heal_voxelize")
    return voxels, coords, npts, nvox


def trim_voxels(voxels, coords, npts, nvox):
    m = int(nvox[0].item())
    return {"voxel_features": voxels[:m], "voxel_coords": coords[:m], "voxel_num_points": npts[:m]}


def mean_vfe(voxels: torch.Tensor, num_points: torch.Tensor) -> torch.Tensor:
    _need_cuda(voxels, num_points)
    M, T, C = voxels.shape
    assert C == 4 and voxels.dtype == torch.float32 and voxels.is_contiguous()
    npts = num_points.to(torch.int32).contiguous()
    out = torch.empty((M, 4), dtype=torch.float32, device=voxels.device)
    check(lib.heal_mean_vfe(_p(voxels), _p(npts), M, T, _p(out), _stream()), "heal_mean_vfe")
    return out


# ------------------------------------------------------------------------------------------------
# PillarVFE + scatter
# ------------------------------------------------------------------------------------------------
def fold_linear_bn(weight: torch.Tensor, bn_w, bn_b, bn_mean, bn_var, eps: float):
    """Linear(no bias) followed by eval BatchNorm1d -> (W' (Cin,Cout) fp32, b' (Cout) fp32), folded in fp64."""
    w = weight.detach().double().cpu()                      # (Cout, Cin)
    scale = bn_w.detach().double().cpu() / torch.sqrt(bn_var.detach().double().cpu() + eps)
    shift = bn_b.detach().double().cpu() - bn_mean.detach().double().cpu() * scale
    wf = (w * scale[:, None]).t().contiguous()              # (Cin, Cout)
    return wf.float(), shift.float()


def pillar_vfe_scatter(voxel_features, voxel_num_points, voxel_coords, w_folded, b_folded,
                       voxel_size, lidar_range, nx: int, ny: int, batch_size: int,
                       want_pillar_features: bool = False, want_canvas: bool = True,
                       num_voxels_dev: Optional[torch.Tensor] = None):
    """Returns (pillar_features (M,64) | None, canvas logical (B,64,ny,nx) channels-last | None)."""
    _need_cuda(voxel_features, voxel_num_points, voxel_coords, w_folded, b_folded)
    M, T, C = voxel_features.shape
    assert C == 4
    dev = voxel_features.device
    vf = voxel_features.contiguous().float()
    npts = voxel_num_points.to(torch.int32).contiguous()
    coords = voxel_coords.to(torch.int32).contiguous()
    cout = w_folded.shape[1]
    pf = torch.empty((M, cout), dtype=torch.float32, device=dev) if want_pillar_features else None
    vs = [float(v) for v in voxel_size]
    off = [vs[i] / 2 + float(lidar_range[i]) for i in range(3)]
    with _Prof("pillar_vfe_scatter(+canvas memset)", 0):
        canvas = torch.zeros((batch_size, ny, nx, cout), dtype=torch.float32, device=dev) if want_canvas else None
        rc = lib.heal_pillar_vfe_scatter(_p(vf), _p(npts), _p(coords), _p(num_voxels_dev), M, T,
                                         _p(w_folded), _p(b_folded), w_folded.shape[0], cout,
                                         _host_f32(vs), _host_f32(off), int(nx), int(ny), _p(pf), _p(canvas), _stream())
    check(rc, "heal_pillar_vfe_scatter")
    return pf, (from_nhwc(canvas) if canvas is not None else None)


# ------------------------------------------------------------------------------------------------
# conv2d (fp32 CUDA-core path)
# ------------------------------------------------------------------------------------------------
class PackedConv:
    """Host-side folded + packed parameters of one Conv2d(+BN)(+ReLU) for the C ABI."""

    def __init__(self, weight, bias, kh, kw, stride, pad, groups, cin, cout, relu, w_cstride, deconv_up=1):
        self.weight, self.bias = weight, bias
        self.kh, self.kw, self.stride, self.pad, self.groups = kh, kw, stride, pad, groups
        self.cin, self.cout, self.relu, self.w_cstride, self.deconv_up = cin, cout, relu, w_cstride, deconv_up

    def to(self, device):
        self.weight = self.weight.to(device)
        self.bias = self.bias.to(device) if self.bias is not None else None
        return self


def _bn_scale_shift(bn, cout):
    if bn is None:
        return torch.ones(cout, dtype=torch.float64), torch.zeros(cout, dtype=torch.float64)
    scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
    shift = bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * scale
    return scale, shift


def pack_conv(conv: torch.nn.Conv2d, bn: Optional[torch.nn.Module], relu: bool, extra_pad: int = 0) -> PackedConv:
    """Fold eval-mode BN into the conv (fp64 on the host) and pack for heal_conv2d_nhwc_f32."""
    w = conv.weight.detach().double().cpu()            # (Cout, Cin/g, kh, kw)
    cout, cing, kh, kw = w.shape
    g = conv.groups
    scale, shift = _bn_scale_shift(bn, cout)
    b = shift.clone()
    if conv.bias is not None:
        b = b + conv.bias.detach().double().cpu() * scale
    w = w * scale[:, None, None, None]
    assert conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1] and conv.dilation == (1, 1)
    pad = conv.padding[0] + extra_pad
    if g == 1:
        cpad = (cout + 3) // 4 * 4
        wp = torch.zeros((kh, kw, cing, cpad), dtype=torch.float64)
        wp[..., :cout] = w.permute(2, 3, 1, 0)
        return PackedConv(wp.float().contiguous(), b.float().contiguous(), kh, kw, conv.stride[0], pad, 1,
                          cing, cout, relu, cpad)
    cg = cout // g
    assert cing == cg and kh == 3 and kw == 3
    # [tap][ci][co][G]
    wg = w.view(g, cg, cg, kh * kw)                     # (G, co, ci, tap)
    wp = wg.permute(3, 2, 1, 0).contiguous()            # (tap, ci, co, G)
    return PackedConv(wp.float(), b.float().contiguous(), kh, kw, conv.stride[0], pad, g, cing * g, cout, relu, 0)


def pack_deconv(deconv: torch.nn.ConvTranspose2d, bn, relu: bool) -> PackedConv:
    """ConvTranspose2d with kernel == stride (no overlap) -> up*up 1x1 weight planes [i][j][Cin][Cout]."""
    w = deconv.weight.detach().double().cpu()          # (Cin, Cout, k, k)
    cin, cout, k, k2 = w.shape
    assert k == k2 == deconv.stride[0] == deconv.stride[1] and deconv.padding == (0, 0) and deconv.groups == 1
    scale, shift = _bn_scale_shift(bn, cout)
    b = shift.clone()
    if deconv.bias is not None:
        b = b + deconv.bias.detach().double().cpu() * scale
    w = w * scale[None, :, None, None]
    cpad = (cout + 3) // 4 * 4
    wp = torch.zeros((k, k, cin, cpad), dtype=torch.float64)
    wp[..., :cout] = w.permute(2, 3, 0, 1)
    return PackedConv(wp.float().contiguous(), b.float().contiguous(), 1, 1, 1, 0, 1, cin, cout, relu, cpad, deconv_up=k)


def conv2d(x: torch.Tensor, pc: PackedConv, residual: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, out_coffset: int = 0, in_coffset: int = 0,
           cin: Optional[int] = None) -> torch.Tensor:
    """x, residual, out: contiguous (N,H,W,C*) NHWC buffers.  Returns the NHWC output buffer.
    `out`/`out_coffset` let a conv write a channel slice of a wider (concat) buffer; `in_coffset`/`cin`
    read a slice."""
    _need_cuda(x, pc.weight)
    N, H, W, Cs = x.shape
    cin = pc.cin if cin is None else cin
    assert x.is_contiguous() and x.dtype == torch.float32 and cin + in_coffset <= Cs
    up = pc.deconv_up
    Ho = (H + 2 * pc.pad - pc.kh) // pc.stride + 1
    Wo = (W + 2 * pc.pad - pc.kw) // pc.stride + 1
    if out is None:
        out = torch.empty((N, Ho * up, Wo * up, pc.cout), dtype=torch.float32, device=x.device)
    assert out.is_contiguous() and out.shape[0] == N and out.shape[1] == Ho * up and out.shape[2] == Wo * up
    ocs = out.shape[3]
    res_cs = 0
    if residual is not None:
        assert residual.is_contiguous() and residual.shape[:3] == out.shape[:3]
        res_cs = residual.shape[3]
    st = _stream()
    fam = ("conv_grouped3x3_f32" if pc.groups > 1 else f"conv_dense{pc.kh}x{pc.kw}_f32")
    flops = 2.0 * N * Ho * Wo * pc.cout * (cin // pc.groups) * pc.kh * pc.kw * up * up
    with _Prof(fam, flops):
      for i in range(up):
        for j in range(up):
            wptr = pc.weight if up == 1 else pc.weight[i, j]
            rc = lib.heal_conv2d_nhwc_f32(_p(x), N, H, W, cin, Cs, in_coffset, _p(wptr), pc.w_cstride, _p(pc.bias),
                                          pc.kh, pc.kw, pc.stride, pc.pad, pc.groups,
                                          _p(residual), res_cs, 0, _p(out), Ho, Wo, pc.cout, ocs, out_coffset,
                                          up, i, j, 1 if pc.relu else 0, st)
            check(rc, "heal_conv2d_nhwc_f32")
    return out


# ------------------------------------------------------------------------------------------------
# fusion
# ------------------------------------------------------------------------------------------------
def pyramid_fuse_level(feat: torch.Tensor, occ: torch.Tensor, theta: torch.Tensor, align_corners: bool,
                       crop_windows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                       out_coffset: int = 0) -> torch.Tensor:
    """feat (n,H,W,C) NHWC buffer, occ (n,H,W) f32, theta (n,2,3) f64 -> out (H,W,C) (NHWC, one scene)."""
    _need_cuda(feat, occ, theta)
    n, H, W, C = feat.shape
    assert feat.is_contiguous() and occ.is_contiguous() and occ.numel() == n * H * W
    th = theta.to(torch.float64).contiguous()
    assert th.shape == (n, 2, 3)
    if out is None:
        out = torch.empty((H, W, C), dtype=torch.float32, device=feat.device)
    ocs = out.shape[-1]
    cw = crop_windows.to(torch.int32).contiguous() if crop_windows is not None else None
    with _Prof("pyramid_fuse_level", 0):
        rc = lib.heal_pyramid_fuse_level(_p(feat), C, _p(occ), _p(th), _p(cw), n, H, W, C, 1 if align_corners else 0,
                                         _p(out), ocs, out_coffset, _stream())
    check(rc, "heal_pyramid_fuse_level")
    return out


def att_fuse(feat: torch.Tensor, theta: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """feat (n,H,W,C) NHWC buffer, theta (n,2,3) f64 -> (H,W,C)."""
    _need_cuda(feat, theta)
    n, H, W, C = feat.shape
    assert feat.is_contiguous()
    th = theta.to(torch.float64).contiguous()
    if out is None:
        out = torch.empty((H, W, C), dtype=torch.float32, device=feat.device)
    rc = lib.heal_att_fuse(_p(feat), C, _p(th), n, H, W, C, _p(out), out.shape[-1], 0, _stream())
    check(rc, "heal_att_fuse")
    return out


# ------------------------------------------------------------------------------------------------
# conv2d, tcgen05 tensor-core path (split-bf16 activations)
# ------------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor, planes: int = 2) -> torch.Tensor:
    """fp32 (...,) -> bf16 (planes, ...) with hi = bf16(x), lo = bf16(x - hi)  (host-side / test helper)."""
    hi = x.to(torch.bfloat16)
    if planes == 1:
        return hi.unsqueeze(0).contiguous()
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous()


def merge_bf16(s: torch.Tensor) -> torch.Tensor:
    return s.float().sum(0)


def _coutp(cout: int) -> int:
    for c in (16, 32, 64):
        if cout <= c:
            return c
    return (cout + 127) // 128 * 128


class PackedConvTC:
    def __init__(self, w, bias, kh, kw, pad, cin, cout, coutp, relu, up, planes):
        self.w, self.bias = w, bias
        self.kh, self.kw, self.pad, self.cin, self.cout, self.coutp = kh, kw, pad, cin, cout, coutp
        self.relu, self.up, self.planes = relu, up, planes

    def to(self, device):
        self.w = self.w.to(device)
        self.bias = self.bias.to(device)
        return self


def pack_conv_tc(conv, bn, relu: bool, planes: int = 2, extra_pad: int = 0) -> PackedConvTC:
    """Fold BN (fp64), lay weights out as [plane][tap*coutp + co][Cin] split-bf16 for heal_conv2d_tc."""
    scale_shift = _bn_scale_shift
    if isinstance(conv, torch.nn.ConvTranspose2d):
        w = conv.weight.detach().double().cpu()            # (Cin, Cout, k, k)
        cin, cout, k, _ = w.shape
        scale, shift = scale_shift(bn, cout)
        b = shift + (conv.bias.detach().double().cpu() * scale if conv.bias is not None else 0)
        w = w * scale[None, :, None, None]
        coutp = _coutp(cout)
        rows = torch.zeros((k * k, coutp, cin), dtype=torch.float64)
        rows[:, :cout, :] = w.permute(2, 3, 1, 0).reshape(k * k, cout, cin)
        wp = split_bf16(rows.reshape(k * k * coutp, cin).float(), planes)
        return PackedConvTC(wp, b.float().contiguous(), 1, 1, 0, cin, cout, coutp, relu, k, planes)
    w = conv.weight.detach().double().cpu()                # (Cout, Cin, kh, kw)
    cout, cin, kh, kw = w.shape
    assert conv.groups == 1 and conv.stride == (1, 1)
    scale, shift = scale_shift(bn, cout)
    b = shift + (conv.bias.detach().double().cpu() * scale if conv.bias is not None else 0)
    w = w * scale[:, None, None, None]
    coutp = _coutp(cout)
    rows = torch.zeros((kh * kw, coutp, cin), dtype=torch.float64)
    rows[:, :cout, :] = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    wp = split_bf16(rows.reshape(kh * kw * coutp, cin).float(), planes)
    return PackedConvTC(wp, b.float().contiguous(), kh, kw, conv.padding[0] + extra_pad, cin, cout, coutp, relu, 1, planes)


def conv2d_tc(x_split: torch.Tensor, pc: PackedConvTC, residual_split: Optional[torch.Tensor] = None,
              residual_f32: Optional[torch.Tensor] = None, want_split: bool = True, want_f32: bool = False,
              out_split: Optional[torch.Tensor] = None, out_coffset: int = 0, in_coffset: int = 0,
              out_f32: Optional[torch.Tensor] = None, out32_coffset: int = 0):
    """x_split: bf16 (planes,N,H,W,Cs).  Returns (out_split | None, out_f32 | None)."""
    _need_cuda(x_split, pc.w)
    P, N, H, W, Cs = x_split.shape
    assert P == pc.planes and x_split.dtype == torch.bfloat16 and x_split.is_contiguous()
    up = pc.up
    Ho, Wo = H + 2 * pc.pad - pc.kh + 1, W + 2 * pc.pad - pc.kw + 1
    dev = x_split.device
    if want_split and out_split is None:
        out_split = torch.empty((P, N, Ho * up, Wo * up, pc.cout), dtype=torch.bfloat16, device=dev)
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty((N, Ho * up, Wo * up, pc.cout), dtype=torch.float32, device=dev)
    res_cs = 0
    res_plane = 0
    if residual_split is not None:
        res_cs, res_plane = residual_split.shape[-1], residual_split[0].numel()
    elif residual_f32 is not None:
        res_cs = residual_f32.shape[-1]
    fam = f"conv_tc{pc.kh}x{pc.kw}" + ("_deconv" if up > 1 else "")
    flops = 2.0 * N * Ho * Wo * pc.cout * pc.cin * pc.kh * pc.kw * up * up
    with _Prof(fam, flops):
        rc = lib.heal_conv2d_tc(_p(x_split), x_split[0].numel(), N, H, W, pc.cin, Cs, in_coffset,
                                _p(pc.w), pc.w.shape[1], pc.coutp, _p(pc.bias), pc.kh, pc.kw, pc.pad, P,
                                _p(residual_split), res_plane, _p(residual_f32), res_cs, 0,
                                _p(out_split), out_split[0].numel() if out_split is not None else 0,
                                out_split.shape[-1] if out_split is not None else 0, out_coffset,
                                _p(out_f32), out_f32.shape[-1] if out_f32 is not None else 0, out32_coffset,
                                Ho, Wo, pc.cout, up, 1 if pc.relu else 0, _stream())
    check(rc, "heal_conv2d_tc")
    return out_split, out_f32
