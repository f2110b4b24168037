"""VoxelBackBone8x mirror (opencood/models/sub_modules/sparse_backbone_3d.py:11-152) without spconv.

SubMConv3d / SparseConv3d / SparseSequential are parameter containers with spconv-2.x weight layout
(Cout, kz, ky, kx, Cin) and the reference's module hierarchy (same state-dict keys, e.g.
`conv2.0.0.weight`, `conv2.0.1.running_mean`).  forward() builds output-stationary rulebooks on the GPU (one per
indice_key, reused by every conv sharing it) and runs each conv + BatchNorm1d + ReLU as one gather-GEMM kernel."""
from functools import partial

import torch
import torch.nn as nn

from ... import ops
from ...engine import _sig, require_eval


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


class _SpConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False, indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            raise NotImplementedError("the reference's sparse convs are all bias-free")
        self.bias = None


class SubMConv3d(_SpConvBase):
    subm = True


class SparseConv3d(_SpConvBase):
    subm = False


class SparseSequential(nn.Sequential):
    pass


def post_act_block(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0, conv_type='subm', norm_fn=None):
    if conv_type == 'subm':
        conv = SubMConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == 'spconv':
        conv = SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False, indice_key=indice_key)
    else:
        raise NotImplementedError(conv_type)
    return SparseSequential(conv, norm_fn(out_channels), nn.ReLU())


_FOLD = {}
_FOLD_TC = {}
TC_SHAPES = {(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128)}     # instantiations of heal_spconv_gather_gemm_tc


def folded_tc(conv, bn):
    """Packed split-bf16 weights for the tensor-core gather-GEMM (ops.pack_spconv_tc), cached like `folded`."""
    sig = _sig(conv, bn)
    hit = _FOLD_TC.get(id(conv))
    if hit is not None and hit[0] == sig:
        return hit[1]
    wp, _ = folded(conv, bn)
    pk = ops.pack_spconv_tc(wp)
    _FOLD_TC[id(conv)] = (sig, pk)
    return pk


def folded(conv, bn):
    """(W' (K,Cin,Cout), b' (Cout)) with eval BatchNorm1d folded in fp64, cached on parameter versions."""
    sig = _sig(conv, bn)
    hit = _FOLD.get(id(conv))
    if hit is not None and hit[0] == sig:
        return hit[1], hit[2]
    w = conv.weight.detach().double().cpu()                                   # (Cout, kz, ky, kx, Cin)
    scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
    shift = bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * scale
    K = conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2]
    wp = (w * scale[:, None, None, None, None]).reshape(conv.out_channels, K, conv.in_channels).permute(1, 2, 0).contiguous()
    dev = conv.weight.device
    wp, b = wp.float().to(dev), shift.float().to(dev)
    _FOLD[id(conv)] = (sig, wp, b)
    return wp, b


class VoxelBackBone8x(nn.Module):
    def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = [int(v) for v in (list(grid_size[::-1]))]
        self.sparse_shape[0] += 1                                             # grid_size[::-1] + [1, 0, 0]
        self.conv_input = SparseSequential(SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key='subm1'),
                                           norm_fn(16), nn.ReLU())
        block = post_act_block
        self.conv1 = SparseSequential(block(16, 16, 3, norm_fn=norm_fn, padding=1, indice_key='subm1'))
        self.conv2 = SparseSequential(
            block(16, 32, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv2', conv_type='spconv'),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key='subm2'),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key='subm2'))
        self.conv3 = SparseSequential(
            block(32, 64, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv3', conv_type='spconv'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm3'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm3'))
        self.conv4 = SparseSequential(
            block(64, 64, 3, norm_fn=norm_fn, stride=2, padding=(0, 1, 1), indice_key='spconv4', conv_type='spconv'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm4'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm4'))
        self.num_point_features = model_cfg.get('num_features_out', 128)
        self.conv_out = SparseSequential(
            SparseConv3d(64, self.num_point_features, (3, 1, 1), stride=(2, 1, 1), padding=0, bias=False, indice_key='spconv_down2'),
            norm_fn(self.num_point_features), nn.ReLU())
        self.backbone_channels = {'x_conv1': 16, 'x_conv2': 32, 'x_conv3': 64, 'x_conv4': 64}

    @staticmethod
    def _run(st, conv, bn, cache, want_f32=False):
        """one conv+BN+ReLU; returns the output SparseTensor (with feats).  Features are fp32 rows (cap, C) on the CUDA-core
        path and split-bf16 rows (cap, 2C) on the tensor-core path (every layer with >= 16 input channels unless the engine is
        in 'fp32' mode; conv_input's 4 input channels = K of 4 per offset stay on the fp32 kernel: 0.3 % of the FLOPs)."""
        from ... import engine
        w, b = folded(conv, bn)
        if conv.subm:
            key = conv.indice_key
            if key not in cache:
                cache[key] = ops.sp_subm_neighbors(st, conv.kernel_size)
            nbr, out = cache[key], st
        else:
            out, nbr = ops.sp_strided(st, conv.kernel_size, conv.stride, conv.padding)
            cache.setdefault("_strided", []).append(out)
        cin, cout = conv.in_channels, conv.out_channels
        if engine.PRECISION != "fp32" and (cin, cout) in TC_SHAPES:
            fin = st.feats if st.feats.dtype == torch.bfloat16 else ops.rows_to_split(st.feats, st.rows_dev)
            return out.with_feats(ops.sp_gather_gemm_tc(fin, nbr, out.rows_dev, folded_tc(conv, bn), b, True, cin, cout, want_f32=want_f32))
        assert st.feats.dtype == torch.float32, "fp32 gather-GEMM after a tensor-core layer is not a configured path"
        return out.with_feats(ops.sp_gather_gemm(st.feats, nbr, out.rows_dev, w, b, True))

    def forward_sparse(self, feats, coords, batch_size, rows_dev=None):
        require_eval(self)
        cache = {}
        st = ops.SparseTensor(feats.contiguous(), coords.to(torch.int32).contiguous(), rows_dev, self.sparse_shape, batch_size)
        x = self._run(st, self.conv_input[0], self.conv_input[1], cache)
        x_conv1 = self._run(x, self.conv1[0][0], self.conv1[0][1], cache)
        levels = [x_conv1]
        x = x_conv1
        for seq in (self.conv2, self.conv3, self.conv4):
            for blk in seq:
                x = self._run(x, blk[0], blk[1], cache)
            levels.append(x)
        out = self._run(x, self.conv_out[0], self.conv_out[1], cache, want_f32=True)
        if not torch.cuda.is_current_stream_capturing():
            ops.sp_check_overflow(cache.get("_strided", []))     # the encoder's only host sync (none inside a captured frame)
        return out, levels

    def forward(self, batch_dict):
        out, levels = self.forward_sparse(batch_dict['voxel_features'], batch_dict['voxel_coords'], int(batch_dict['batch_size']),
                                          batch_dict.get('num_voxels_dev'))
        batch_dict.update({'encoded_spconv_tensor': out, 'encoded_spconv_tensor_stride': 8,
                           'multi_scale_3d_features': dict(zip(('x_conv1', 'x_conv2', 'x_conv3', 'x_conv4'), levels)),
                           'multi_scale_3d_strides': {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8}})
        return batch_dict
