"""Test graphs that are not in the reference: an unstructured-grid stand-in for the ICON ``ncells`` axis."""
import torch


def tri_mesh_edges(h: int, w: int) -> torch.Tensor:
    """Cells = the 2 h w triangles of a regular triangulation of an h x w lattice of squares (node 2 (r w + c) + k,
    k = 0 the lower-left, 1 the upper-right triangle of square (r, c)); edges join triangles that share a side
    (three neighbours per cell, fewer on the rim), both directions.  int64 [2, E]."""
    r, c = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    sq = (r * w + c).reshape(-1)
    a, b = 2 * sq, 2 * sq + 1
    src, dst = [a], [b]                                   # the diagonal inside a square
    m = (c.reshape(-1) + 1 < w)
    src.append(b[m]); dst.append(a[m] + 2)                # right side: the lower-left triangle of square (r, c + 1)
    m = (r.reshape(-1) + 1 < h)
    src.append(b[m]); dst.append(a[m] + 2 * w)            # top side: the lower-left triangle of square (r + 1, c)
    s, d = torch.cat(src), torch.cat(dst)
    return torch.stack([torch.cat([s, d]), torch.cat([d, s])])
