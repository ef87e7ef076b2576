"""Generates the committed golden fixtures under tests/golden/ (run in the build container).

Inputs:  the reference's only data fixture,
         /root/reference/tests/test_data/atmcirc-straka_93_-10.0_3000.0_2000.0_DOM01_ML_20080801T000000Z.nc
         (sha256 e0392ddc...c03495): theta_v(time=2, member=2, height=10, ncells=10) float32, stored
         contiguously at byte offset 0x1AAA; read without h5py.
Outputs: cfg1_x.npy            [2 time steps, 2 members, 100] fp32 -- GraphDataset.get layout
                               (reference src/gwen/utils.py:195-202: member rows, height*ncells cols)
         cfg1_out.npy          oracle forward of GNNModel(100 -> 1024 -> ... -> 100) for both time
                               steps on K_2 (edge_index [[0,1],[1,0]]); parameters regenerated
                               bit-exactly by tests/golden/weights.py (seed 23: reference
                               src/gwen/config.json:14), not stored
         grid_3x4_edges.npy    oracle grid(3, 4) edge_index (70 edges)
         grid_5x7_csr.npz      oracle dst-sorted CSR of grid(5, 7)
         layer_grid_6x5.npz    one GCNConv(12 -> 20) on grid(6, 5): x, W, b, out (fp32) + fp64 dense
The reference's PyG dependency cannot be imported here, so outputs come from oracle/gcn_oracle.py
(the restatement) and are cross-checked against its independent dense fp64 form before writing.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gcn_oracle as orc  # noqa: E402
from tests.golden import weights as wts  # noqa: E402

NC = ("/root/reference/tests/test_data/"
      "atmcirc-straka_93_-10.0_3000.0_2000.0_DOM01_ML_20080801T000000Z.nc")
SHA = "e0392ddc19312840c6094f0652f9273593720e3a5b83b73da5aeae6b6ac03495"


def main():
    buf = open(NC, "rb").read()
    assert hashlib.sha256(buf).hexdigest() == SHA
    theta = np.frombuffer(buf[0x1AAA:0x1AAA + 1600], "<f4").reshape(2, 2, 10, 10)  # t, member, h, cell
    x = theta.reshape(2, 2, 100).copy()
    np.save(os.path.join(HERE, "cfg1_x.npy"), x)

    model = orc.GNNModelOracle(100, 100, 1024)
    wts.fill_model_(model, 23)
    ei = torch.tensor([[0, 1], [1, 0]])
    assert torch.equal(ei, orc.complete_graph(2))
    with torch.no_grad():
        out = torch.stack([model(torch.from_numpy(x[t]), ei) for t in range(2)])
    # K_2: A_hat = [[.5,.5],[.5,.5]] at every layer -> both rows identical (Appendix C.1)
    assert torch.allclose(out[:, 0], out[:, 1], rtol=0, atol=1e-6)
    np.save(os.path.join(HERE, "cfg1_out.npy"), out.numpy())

    np.save(os.path.join(HERE, "grid_3x4_edges.npy"), orc.grid(3, 4).numpy())
    rowptr, src, perm, dis = orc.dst_sorted_csr(orc.grid(5, 7), 35)
    np.savez(os.path.join(HERE, "grid_5x7_csr.npz"), rowptr=rowptr, src=src, perm=perm, dis=dis)

    ei = orc.grid(6, 5)
    xl, wl, bl = wts.features((30, 12), 5), wts.glorot(20, 12, 6), wts.small_bias(20, 7)
    out = orc.gcn_conv_forward(xl, ei, wl, bl)
    dense = orc.dense_gcn_forward(xl, ei, wl, bl)
    assert (out.double() - dense).abs().max() / dense.abs().max() < 1e-6
    np.savez(os.path.join(HERE, "layer_grid_6x5.npz"), x=xl.numpy(), w=wl.numpy(), b=bl.numpy(),
             out=out.numpy(), dense=dense.numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
