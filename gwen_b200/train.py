"""Training-step glue around the layer stack: the inner loop of the
reference's ``train_with_configs`` (``src/gwen/models_gnn.py:359-375``) without its per-iteration
host round trips -- a fused, synchronisation-free masked L1 loss and a ``train_step`` that keeps
everything on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import check, lib
from .ops import dtype_code

__all__ = ["masked_l1_loss", "train_step", "eval_step", "gather_eval_results"]


class _MaskedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, mask, count=None):
        if not output.is_cuda:
            raise RuntimeError("gwen_b200.masked_l1_loss runs on CUDA tensors only (no CPU fallback)")
        y = output.contiguous()
        t = target.to(y.dtype).contiguous()
        m = mask.to(torch.bool).contiguous()
        n, f = y.shape[-2], y.shape[-1]
        b = y.numel() // max(1, n * f)
        if m.numel() != n or t.shape != y.shape:
            raise ValueError("mask must be [N] and target shaped like output")
        with torch.cuda.device(y.device):
            need = C.c_size_t()
            check(lib().gwen_masked_l1_workspace_bytes(n, C.byref(need)), "masked_l1 ws")
            ws = torch.empty(need.value, dtype=torch.uint8, device=y.device)
            ls = torch.empty(2, dtype=torch.float32, device=y.device)
            cnt = None if count is None else count.detach().to(device=y.device, dtype=torch.float32).reshape(1).contiguous()
            check(lib().gwen_masked_l1_fwd(y.data_ptr(), t.data_ptr(), m.data_ptr(), b, n, f, dtype_code(y.dtype),
                                           None if cnt is None else cnt.data_ptr(),
                                           ls.data_ptr(), ws.data_ptr(), need.value,
                                           torch.cuda.current_stream().cuda_stream), "gwen_masked_l1_fwd")
        ctx.save_for_backward(y, t, m, ls)
        return ls[0].clone()

    @staticmethod
    def backward(ctx, dloss):
        y, t, m, ls = ctx.saved_tensors
        n, f = y.shape[-2], y.shape[-1]
        b = y.numel() // max(1, n * f)
        g = dloss.to(torch.float32).contiguous()
        with torch.cuda.device(y.device):
            dy = torch.empty_like(y)
            check(lib().gwen_masked_l1_bwd(y.data_ptr(), t.data_ptr(), m.data_ptr(), ls.data_ptr(), g.data_ptr(),
                                           b, n, f, dtype_code(y.dtype), dy.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "gwen_masked_l1_bwd")
        return dy, None, None, None


def masked_l1_loss(output: torch.Tensor, target: torch.Tensor, target_mask: torch.Tensor,
                   count: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``loss_func(output, target, target_mask)`` of the reference (``models_gnn.py:261-265``) =
    ``L1Loss()(output[target_mask], target[target_mask])`` for ``output [..., N, C]`` and a boolean
    ``target_mask [N]``, computed without boolean-mask indexing (no host synchronisation).
    ``count`` (device scalar): divide by this many masked nodes instead of ``target_mask.sum()`` -- a
    rank of a partitioned mesh passes the global count and gets its share of the global loss."""
    return _MaskedL1.apply(output, target, target_mask, count)


def train_step(model, node_features: torch.Tensor, edge_index, target_mask: torch.Tensor,
               optimizer: Optional[torch.optim.Optimizer] = None, scheduler=None) -> torch.Tensor:
    """One iteration of the reference loop (``models_gnn.py:364-375``): zero_grad, forward, loss against
    the INPUT features on the masked nodes, backward, optimizer / scheduler step.  Returns the loss as
    a device scalar (the reference accumulates it without ``.item()`` as well)."""
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    else:
        for p in model.parameters():
            p.grad = None
    output = model(node_features, edge_index)
    loss = masked_l1_loss(output, node_features, target_mask)
    loss.backward()
    if optimizer is not None:
        optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()


@torch.no_grad()
def eval_step(model, node_features: torch.Tensor, edge_index, target_mask: torch.Tensor
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """One iteration of the reference evaluation loop (``models_gnn.py:440-450``): forward, the masked L1
    loss against the input features, and the prediction the reference keeps, ``output[1]``.  Device
    scalars / tensors, no host synchronisation."""
    output = model(node_features, edge_index)
    return masked_l1_loss(output, node_features, target_mask), output[1]


def gather_eval_results(avg_loss: torch.Tensor, y_preds: List[torch.Tensor], group=None
                        ) -> Optional[Tuple[float, torch.Tensor]]:
    """The collective tail of ``eval_gnn_with_configs`` (``models_gnn.py:452-489``): every rank
    contributes its average loss and its list of kept predictions; rank 0 returns
    ``(mean of the ranks' losses, predictions concatenated in rank order)``, the other ranks ``None``.
    Shapes follow the reference literally: ``torch.cat(y_preds)`` (``:465``) of the kept ``output[1]``
    rows -- 1-D ``[C]`` each for ``output [N, C]`` -- is the 1-D ``[T * C]`` tensor, and the rank-ordered
    ``torch.cat`` (``:482``) returns ``[world * T * C]``.
    The reference all_gathers the rank ids next to the predictions and sorts by them; ``all_gather``
    already returns rank order, so the sort is the identity and is not repeated here."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    y = torch.cat(list(y_preds)) if y_preds else avg_loss.new_zeros(0)
    ys = [torch.zeros_like(y) for _ in range(world)]
    dist.all_gather(ys, y.contiguous(), group=group)
    losses = [torch.zeros_like(avg_loss) for _ in range(world)]
    dist.all_gather(losses, avg_loss.detach().contiguous(), group=group)
    if rank != 0:
        return None
    return float(torch.stack(losses).mean()), torch.cat(ys).cpu()
