"""Host-side logic that needs no GPU: tensor-view helpers, partition bookkeeping, the TF32 split
arithmetic the 3xTF32 kernels rely on (restated in numpy), ABI bookkeeping."""
import os
import re

import numpy as np
import pytest
import torch

from gwen_b200 import _lib, ops, partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rows_view_accepts_dense_rows_only():
    t = torch.zeros(3, 10, 8)
    assert ops._rows_view(t).shape == (3, 10, 8)
    assert ops._rows_view(t[0]).shape == (1, 10, 8)
    v = torch.zeros(3, 14, 8)[:, 2:12]                   # batch-strided, rows dense
    rv = ops._rows_view(v)
    assert rv is not None and rv.stride() == (14 * 8, 8, 1)
    assert ops._rows_view(torch.zeros(3, 10, 16)[:, :, :8]) is None      # row pitch != F
    assert ops._rows_view(torch.zeros(10, 8).t()) is None                # not unit stride
    four = torch.zeros(2, 3, 10, 8)
    assert ops._rows_view(four).shape == (6, 10, 8)


def test_copy_rows_matches_copy():
    src = torch.arange(3 * 10 * 4, dtype=torch.float32).reshape(3, 10, 4)
    big = torch.full((3, 16, 4), -1.0)
    ops.copy_rows_(big[:, 3:13], src)
    assert torch.equal(big[:, 3:13], src) and torch.all(big[:, :3] == -1) and torch.all(big[:, 13:] == -1)


@pytest.mark.parametrize("h,w,world", [(8, 6, 2), (41, 50, 3), (582, 390, 8), (7, 3, 7), (1158, 774, 8)])
def test_band_ranges_partition_the_mesh(h, w, world):
    ranges = partition.band_ranges(h, w, world)
    assert len(ranges) == world
    assert ranges[0].start == 0 and ranges[-1].stop == h * w
    rows = []
    for a, b in zip(ranges, ranges[1:]):
        assert a.stop == b.start
    for r in ranges:
        assert r.start % w == 0 and r.stop % w == 0      # whole grid rows
        rows.append(len(r) // w)
    assert max(rows) - min(rows) <= 1                    # balanced to one row


def _split_tf32_rn(x):
    """numpy restatement of split_tf32 (csrc/linear_tf32x3.cu): hi / lo rounded to nearest TF32."""
    u = x.view(np.uint32).astype(np.uint64)
    h = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)
    d = (x - h).astype(np.float32)
    l = ((d.view(np.uint32).astype(np.uint64) + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)
    return h, d, l


def test_tf32_split_is_exact_to_2e_minus_22():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    h, d, l = _split_tf32_rn(x)
    assert np.all((h.view(np.uint32) & 0x1FFF) == 0) and np.all((l.view(np.uint32) & 0x1FFF) == 0)   # TF32 values
    assert np.array_equal((x.astype(np.float64) - h.astype(np.float64)).astype(np.float32), d)       # x - hi is exact
    rel = np.abs(x.astype(np.float64) - h.astype(np.float64) - l.astype(np.float64)) / np.abs(x.astype(np.float64))
    assert rel.max() <= 2.0 ** -22
    assert np.abs(x.astype(np.float64) - h.astype(np.float64)).max() / 1.0 >= 0        # (hi alone is only 2^-11)
    rel_hi = np.abs(x.astype(np.float64) - h.astype(np.float64)) / np.abs(x.astype(np.float64))
    assert rel_hi.max() <= 2.0 ** -11 and rel_hi.max() > 2.0 ** -13


def test_three_term_product_reaches_fp32_accuracy():
    """lo_a hi_b + hi_a lo_b + hi_a hi_b (the three kind::tf32 products) vs the exact product."""
    rng = np.random.default_rng(1)
    a = rng.standard_normal(100000).astype(np.float32)
    b = rng.standard_normal(100000).astype(np.float32)
    ah, _, al = _split_tf32_rn(a)
    bh, _, bl = _split_tf32_rn(b)
    f = np.float64
    approx = al.astype(f) * bh.astype(f) + ah.astype(f) * bl.astype(f) + ah.astype(f) * bh.astype(f)
    exact = a.astype(f) * b.astype(f)
    rel = np.abs(approx - exact) / np.abs(exact)
    assert rel.max() <= 2.0 ** -20                       # dropped: lo*lo and the split remainders
    one = np.abs(ah.astype(f) * bh.astype(f) - exact) / np.abs(exact)
    assert one.max() > 2.0 ** -12                        # a single TF32 product would miss the 1e-5 bar


def test_every_extern_c_definition_is_declared_and_bound():
    """csrc/*.cu `extern "C"` definitions == include/gwen_b200.h declarations == ctypes PROTOTYPES."""
    defined = set()
    for f in os.listdir(_lib.CSRC):
        if f.endswith(".cu"):
            src = open(os.path.join(_lib.CSRC, f)).read()
            defined |= set(re.findall(r'extern "C"\s+[\w\s\*]+?\b(gwen_\w+)\s*\(', src))
    header = open(os.path.join(ROOT, "include", "gwen_b200.h")).read()
    declared = set(re.findall(r'\b(gwen_\w+)\s*\(', header)) - {"gwen_halo_peers", "gwen_tile_plan"}
    assert defined == declared, (sorted(defined - declared), sorted(declared - defined))
    assert set(_lib.PROTOTYPES) == declared, (sorted(set(_lib.PROTOTYPES) ^ declared))


def test_fused_choice_threshold():
    class G:
        is_plain_mesh = True
        grid_shape = (1158, 774)
    x = torch.empty(0)

    class X:      # a stand-in with the attributes gcn_fused_supported reads (no CUDA tensor on this box)
        dtype, is_cuda, shape = torch.bfloat16, True, (8, 1158 * 774, 64)

        def numel(self):
            return 8 * 1158 * 774 * 64
    w = torch.empty(1024, 64)
    assert ops.gcn_fused_supported(G, X(), w) and ops.gcn_fused_preferred(G, X(), w)
    G2 = type("G2", (), {"is_plain_mesh": True, "grid_shape": (582, 390)})
    X2 = type("X2", (), {"dtype": torch.bfloat16, "is_cuda": True, "shape": (1, 582 * 390, 64),
                         "numel": lambda self: 582 * 390 * 64})
    assert ops.gcn_fused_supported(G2, X2(), w) and not ops.gcn_fused_preferred(G2, X2(), w)   # small mesh
    assert ops.gcn_fused_supported(G, X(), torch.empty(1024, 512))                              # k_in = 512: one A buffer (round 2)
    assert not ops.gcn_fused_supported(G, X(), torch.empty(1024, 576))                          # k_in > 512
    assert not ops.gcn_fused_supported(G, X(), torch.empty(64, 1024))
