"""``Adam``: the optimizer step of the reference training loop on the GPU in one launch.

Reference: ``optimizer = optim.Adam(model.parameters(), lr=config["lr"] * 10)`` (``src/gwen/train_gnn.py:111``),
stepped once per batch at ``src/gwen/models_gnn.py:373``.  Same constructor arguments, ``step()`` /
``zero_grad()`` / ``state_dict()`` behaviour and update rule as ``torch.optim.Adam`` (``amsgrad=False``,
``maximize=False``); the update of every parameter tensor runs in ONE kernel (``gwen_adam_step``) instead of
torch's ~10 multi-tensor launches per step.  fp32 parameters with fp32 gradients only (the numerics of BASELINE
config 5: bf16 activations, fp32 master weights); anything else raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib

__all__ = ["Adam"]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda:
                    raise RuntimeError("gwen_b200.optim.Adam runs on CUDA parameters only (no CPU fallback)")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise NotImplementedError("gwen_b200.optim.Adam: fp32 parameters with dense fp32 gradients only")
                if not p.is_contiguous():
                    raise NotImplementedError("gwen_b200.optim.Adam: contiguous parameters only")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            # torch steps every parameter's counter together; parameters that join later get their own count
            by_step = {}
            for p in ps:
                st = self.state[p]
                st["step"] += 1
                by_step.setdefault((st["step"], p.device), []).append(p)
            b1, b2 = group["betas"]
            for (step, dev), plist in by_step.items():
                n = len(plist)
                arr = C.c_void_p * n
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in plist]
                with torch.cuda.device(dev):
                    check(lib().gwen_adam_step(
                        n, arr(*[p.data_ptr() for p in plist]), arr(*[g.data_ptr() for g in grads]),
                        arr(*[self.state[p]["exp_avg"].data_ptr() for p in plist]),
                        arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in plist]),
                        (C.c_int64 * n)(*[p.numel() for p in plist]), step, float(group["lr"]), float(b1), float(b2),
                        float(group["eps"]), float(group["weight_decay"]),
                        torch.cuda.current_stream().cuda_stream), "gwen_adam_step")
                # the kernel wrote the parameters (and moments) through raw pointers: tell autograd, so that saved
                # tensors are invalidated and version-keyed caches (ops._cast_cached: the bf16 copy of an fp32
                # master weight) see the update
                torch.autograd.graph.increment_version(
                    plist + [self.state[p]["exp_avg"] for p in plist] + [self.state[p]["exp_avg_sq"] for p in plist])
        return loss
