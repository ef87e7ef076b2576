"""Bit-reproducible synthetic parameters for the golden fixtures and parity tests.

A splitmix64 hash of the element index gives 24 random bits -> an exactly representable fp32 in
[0, 1); no dependence on any library's RNG stream, so fixtures made in the build container and
tests run on the GPU box regenerate identical weights without storing them.
"""
import numpy as np
import torch


def det_uniform(shape, seed: int) -> np.ndarray:
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
    return u.reshape(shape)


def glorot(out_c: int, in_c: int, seed: int) -> torch.Tensor:
    """glorot-uniform U(-a, a), a = sqrt(6 / (in + out)) like PyG's Linear(weight_initializer='glorot')."""
    a = np.float32(np.sqrt(6.0 / (in_c + out_c)))
    return torch.from_numpy((det_uniform((out_c, in_c), seed) * np.float32(2.0) - np.float32(1.0)) * a)


def small_bias(out_c: int, seed: int) -> torch.Tensor:
    return torch.from_numpy((det_uniform((out_c,), seed) - np.float32(0.5)) * np.float32(0.2))


def features(shape, seed: int) -> torch.Tensor:
    """Roughly unit-variance features: sum of 4 uniforms, centred and scaled."""
    u = sum(det_uniform(shape, seed * 4 + i) for i in range(4))
    return torch.from_numpy((u - np.float32(2.0)) * np.float32(np.sqrt(3.0)))


def fill_model_(model: torch.nn.Module, seed: int = 23) -> None:
    """Deterministically (re)initialise every GCNConv-style parameter of a module tree."""
    with torch.no_grad():
        for i, (name, p) in enumerate(sorted(model.named_parameters())):
            if p.dim() == 2:
                p.copy_(glorot(p.shape[0], p.shape[1], seed * 1000 + i))
            else:
                p.copy_(small_bias(p.shape[0], seed * 1000 + i))
