"""Deterministic, construction-order-independent parameters (workload data shared by bench, tests and the reference arm):
each tensor is drawn from a generator seeded by crc32(key)^seed, so the reference model (in the
build container), the oracle and the heal_b200 mirror modules (on the GPU box) all get identical
weights from just the {key: shape} table, without shipping a checkpoint."""
import zlib
from typing import Dict, Tuple

import torch


def make_tensor(key: str, shape: Tuple[int, ...], seed: int = 1234) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ seed) & 0x7FFFFFFF)
    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.long)
    if key.endswith("running_mean"):
        return torch.randn(shape, generator=g) * 0.1
    if key.endswith("running_var"):
        return torch.rand(shape, generator=g) + 0.5
    if len(shape) == 1:
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "weight":                      # BN / norm scale
            return torch.rand(shape, generator=g) + 0.5
        return torch.randn(shape, generator=g) * 0.1   # biases
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    if "deblocks" in key and len(shape) == 4:     # ConvTranspose2d weight (Cin, Cout, k, k): fan_in = Cin
        fan_in = shape[0]
    return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 1234) -> Dict[str, torch.Tensor]:
    return {k: make_tensor(k, tuple(v), seed) for k, v in shapes.items()}


def shapes_of(module) -> Dict[str, Tuple[int, ...]]:
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}
