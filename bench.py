#!/usr/bin/env python
"""bench.py -- headline benchmark of the GWEN GCN message-passing hot path on B200.

Workload (BASELINE.json configs[1]): synthetic COSMO-2E-sized grid, 582 x 390 = 226 980 nodes,
PyG ``grid`` 8-neighbour mesh (E' = 2 036 992 messages incl. self loops), F = 256, fp32.
One step = one single-layer message+aggregate pass (``GCNConv.propagate``: gather source rows,
scale by the symmetric GCN norm, segment-reduce into destinations, + bias) over the whole mesh.
Metric: message-passing edges/s = E' x steps / time (whole job, all ranks).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA path, C ABI)
  python bench.py --impl reference [--gpus N] ...                  # the reference's CPU path

N > 1 (launched with torch.distributed.run, one rank per GPU): WEAK scaling -- every rank owns a
582 x 390 row band of a (582 N) x 390 mesh.  The one-row halos are fetched INSIDE the aggregation
kernel from the neighbours' buffers over NVLink peer memory (gwen_grid_stencil_peer_fwd,
partition.PeerMeshBand); --halo nccl selects the NCCL send/recv exchange (partition.MeshBand).
Before anything is timed every rank checks its band of the partitioned result BITWISE against the
single-GPU stencil on the same global input (``partition_parity``).

Timing: W untimed steps, then exactly K steps bracketed by barrier + synchronize, CUDA events on
the launching stream, max over ranks.  The step rotates over three input/output buffer pairs
(1.4 GB per rank >> 126 MB L2): no line of a step's input or output survives in L2 until the
buffer comes round again.  When K < 200 a second region of 200 steps is timed as well
(``long_run``); the roofline uses the longer region.  ``e2e`` repeats the measurement through the
public API with pinned HOST buffers (H2D of x and D2H of the result inside the timed region).

``other_configs`` carries BASELINE's remaining configs, measured in the same invocation:
  N = 1: cfg 3 full six-layer forward (bf16, B = 8) next to the torch-op sequence torch_geometric 2.3.1
         runs for the same layers on the same GPU (``gpu_reference``) and an in-run parity check of a row
         band against the fp32 oracle; one cfg 5 member training step; cfg 4 (2048^2) forward and
         forward+backward on one GPU (the N = 1 point of the strong-scaling family); the cfg 2 aggregation on
         graphs that are not the plain mesh (node ids randomly permuted -> locality tiles; 10 % of the nodes
         cut out -> masked stencil), each next to the row kernel on the same graph.
  N > 1: cfg 4 strong scaling (the same fixed 2048^2 mesh split into row bands: forward, and
         forward + backward + gradient all-reduce); at N >= 4 the cfg 5 training step (B = 21).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, FEAT = 582, 390, 256
NBUF = 3
LONG_STEPS = 200
METRIC, UNIT = "gcn_message_passing_edges_per_s", "edges/s"
WORKLOAD = ("cfg2: synthetic COSMO-2E grid 582x390 (226980 nodes, E'=2036992 messages incl. self "
            "loops), F=256 fp32, single-layer message+aggregate (GCNConv.propagate)")
BF16_TOL = 2e-2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-model-probes", action="store_true",
                    help="skip other_configs (cfg 3 forward / cfg 4 strong scaling / cfg 5 training step)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo exchange inside the kernel over peer memory, or NCCL send/recv")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return (float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", 1407.0)),
                "measured (MEASURED_PEAKS.json: hbm_gbs copy peak, bf16_tflops_sustained)")
    except Exception:  # noqa: BLE001
        return 6650.0, 1407.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """(dram bytes per launch, source) of the headline kernel from the ncu capture committed for this round -- it
    is NOT measured in this run (ncu replays a kernel ~40 times; a number taken under it is never a bench value)."""
    rel = os.path.join("profiles", "r02_k_grid_stencil_cfg2.json")
    try:
        with open(os.path.join(ROOT, rel)) as f:
            v = json.load(f).get("dram_bytes_per_launch")
        return v, ("%s: dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture of "
                   "`bench.py --steps 20 --warmup 3` (committed, not measured in this run); ~44 MB of a launch's output "
                   "is still dirty in L2 when the kernel ends, so the counter reads below the 465.8 MB the kernel moves" % rel)
    except Exception:  # noqa: BLE001
        return None, "no committed ncu capture found"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the benchmark runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 8]
        window = "timed+e2e"
        if len(rows) < 3:
            rows = [r for _, r in self.rows if len(r) >= 8]
            window = "whole run (timed region shorter than the sampling period)"
        busy = [r for r in rows if r[3].isdigit() and int(r[3]) > 0] or rows
        try:
            sm = statistics.median(float(r[0]) for r in busy)
            mx = max(float(r[1]) for r in busy)
            pw = max(float(r[2]) for r in busy)
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "unparsable nvidia-smi output"}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in busy)]
        return {"sm_mhz": sm, "sm_max_mhz": mx, "power_w_max": pw, "reasons": reasons,
                "samples": len(busy), "window": window}


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle restatement of the PyG op sequence on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_propagate_rate(rows: int, steps: int, warmup: int = 1):
    """edges/s of index_select + scale + scatter_add_ + bias on a rows x W band of the mesh."""
    import torch
    from oracle import gcn_oracle as orc  # the checker, used here as the timed CPU reference
    torch.set_num_threads(os.cpu_count() or 1)
    n = rows * W
    ei = orc.grid(rows, W)
    ei2, ew, _ = orc.gcn_norm(ei, n)
    g = torch.Generator().manual_seed(23)
    x = torch.randn(n, FEAT, generator=g)
    bias = torch.randn(FEAT, generator=g) * 0.1
    for _ in range(warmup):
        orc.propagate(x, ei2, ew, n) + bias
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.propagate(x, ei2, ew, n) + bias
    dt = time.perf_counter() - t0
    return ei2.size(1) * steps / dt, dt / steps, ei2.size(1), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bound the run: one full-mesh step costs ~1 s on 8 cores; shrink to a row band if K is large
    budget_s = 120.0
    rate, per_step, _, cores = cpu_propagate_rate(64, 1, warmup=1)
    est_full = (H * W * 9) / rate
    rows = H if est_full * (args.steps + args.warmup) <= budget_s else \
        max(16, min(H, int(H * budget_s / (est_full * (args.steps + args.warmup)))))
    rate, per_step, edges, cores = cpu_propagate_rate(rows, args.steps, warmup=max(1, min(args.warmup, 3)))
    sample = "full 582x390 mesh per step" if rows == H else \
        "%dx390 row band of the mesh per step (%d messages), rate is per-edge so no scaling needed" % (rows, edges)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "arm": "oracle port of the PyG-2.3.1 "
                                        "GCNConv.propagate op sequence on host CPU (torch_geometric is not installable here)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
# "reference torch_geometric GPU path": the op sequence PyG 2.3.1 executes for GCNConv (SURVEY.md table
# 2.3: add_remaining_self_loops + gcn_norm per call, F.linear, index_select, multiply, scatter_add_,
# + bias, relu), restated with the same ATen ops on the GPU.  Baseline leg only: nothing of ours runs in it.
# ---------------------------------------------------------------------------------------------
def torch_gcn_norm(edge_index, n, dtype):
    import torch
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loop = torch.arange(n, device=edge_index.device)
    ei = torch.cat([edge_index[:, keep], torch.stack([loop, loop])], dim=1)
    ew = torch.ones(ei.size(1), dtype=dtype, device=ei.device)
    deg = torch.zeros(n, dtype=dtype, device=ei.device).scatter_add_(0, ei[1], ew)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return ei, dis[ei[0]] * ew * dis[ei[1]]


def torch_gcn_conv(x, edge_index, weight, bias, relu, cache=None):
    import torch
    n = x.size(-2)
    ei, ew = cache if cache is not None else torch_gcn_norm(edge_index, n, x.dtype)
    h = torch.nn.functional.linear(x, weight)
    msg = h.index_select(-2, ei[0]) * ew.view(-1, 1)
    out = torch.zeros_like(h).scatter_add_(-2, ei[1].view(-1, 1).expand_as(msg), msg)
    out = out + bias
    return torch.relu(out) if relu else out


def _timed(fn, warm, iters):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def _live_layers(model):
    d, u = model.conv_layers.down_conv_layers, model.conv_layers.up_conv_layers
    return (("conv1", d.conv1, True), ("conv2", d.conv2, True), ("conv3", d.conv3, True),
            ("upconv3", u.upconv3, True), ("upconv4", u.upconv4, True), ("upconv5", u.upconv5, False))


# ---------------------------------------------------------------------------------------------
# other_configs, N = 1
# ---------------------------------------------------------------------------------------------
def probe_full_forward_cfg3(dev):
    """BASELINE config 3: COSMO-1E grid, six layers, bf16, B = 8.  Ours, the PyG op sequence on the same
    GPU, and an in-run parity check of one member's row band against the fp32 oracle."""
    import torch
    import gwen_b200 as gw
    _, tf_peak, _ = measured_peaks()
    h, wd, c, b = 1158, 774, 64, 8
    n = h * wd
    torch.manual_seed(23)
    ei = gw.grid(h, wd, dev)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
    model32 = gw.GNNModel(cfg)                       # fp32 master copy (CPU) for the oracle
    with torch.no_grad():
        for p in model32.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)                    # exercise the bias path (PyG inits biases to 0)
    model = gw.GNNModel(cfg)
    model.load_state_dict(model32.state_dict())
    model = model.to(dev).to(torch.bfloat16)
    xg = torch.Generator().manual_seed(23)
    x_cpu = torch.randn(b, n, c, generator=xg)
    x = x_cpu.to(dev).to(torch.bfloat16)
    with torch.no_grad():
        _timed(lambda: model(x, ei), 2, 1)
        fs = ClockSampler(dev.index or 0)          # SM clock / power WHILE the forward runs (it sits on the power cap)
        t0 = time.time()
        ms = _timed(lambda: model(x, ei), 1, 12)
        fwd_clocks = fs.stop(t0, time.time())
        y = model(x, ei)
    e1 = gw.grid_edge_count(h, wd)
    flops = 2.0 * b * n * 1441792
    out = {
        "workload": "cfg3: COSMO-1E grid 1158x774 (896292 nodes, E'=%d), six GCN layers 64-1024-512-256-512-1024-64, bf16, "
                    "8 ensemble members per step, synthetic data, random-init weights" % e1,
        "ms_per_step": ms, "grid_steps_per_s": 1e3 / ms, "member_steps_per_s": b * 1e3 / ms,
        "edges_per_s": b * e1 * 6 / (ms * 1e-3), "flops_per_step": flops, "tflops": flops / (ms * 1e-3) / 1e12,
        "clocks_during_forward": fwd_clocks,
        "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": flops / (ms * 1e-3) / 1e12 / tf_peak,
                     "note": "whole forward; lower bounds: %.1f ms (flops at the sustained bf16 peak), 14.9 ms "
                             "(97 GB of unfused activation traffic at the copy peak)" % (flops / tf_peak / 1e9)}}
    # ---- in-run parity: member 0, rows [600, 608) vs the fp32 oracle on the 20-row band [594, 614) ----
    try:
        from oracle import gcn_oracle as orc  # checker only
        ref = orc.GNNModelOracle(c, c, 1024)
        ref.load_state_dict(model32.state_dict())
        r0, r1, halo = 600, 608, 6
        with torch.no_grad():
            band = x_cpu[0, (r0 - halo) * wd:(r1 + halo) * wd]
            yr = ref(band, orc.grid(r1 - r0 + 2 * halo, wd))[halo * wd:(halo + r1 - r0) * wd]
        yo = y[0, r0 * wd:r1 * wd].float().cpu()
        nmax = ((yo - yr).abs().max() / yr.abs().max()).item()
        out["parity"] = {"nmax_vs_fp32_oracle_band": nmax, "tol": BF16_TOL, "ok": bool(nmax <= BF16_TOL),
                         "what": "member 0, mesh rows 600..607 (6192 nodes x 64 channels) of the bf16 forward vs the "
                                 "fp32 CPU oracle evaluated on the 20-row band 594..613 (six layers reach six rows); "
                                 "max|y - y_ref| / max|y_ref|"}
    except Exception as e:  # noqa: BLE001
        out["parity"] = {"error": str(e)[:200]}
    # ---- the PyG op sequence on this GPU, one member at a time (its [E', 1024] message tensor is 16.5 GB) ----
    try:
        layers = _live_layers(model)

        def torch_forward(xm, cached):
            cache = torch_gcn_norm(ei, n, xm.dtype) if cached else None
            for _, conv, relu in layers:
                xm = torch_gcn_conv(xm, ei, conv.lin.weight, conv.bias, relu, cache)
            return xm
        with torch.no_grad():
            ms_re = _timed(lambda: torch_forward(x[0], False), 1, 3) * b
            ms_ca = _timed(lambda: torch_forward(x[0], True), 1, 3) * b
            yt = torch_forward(x[0], True)
        out["gpu_reference"] = {
            "what": "torch op sequence of torch_geometric 2.3.1 GCNConv (gcn_norm, F.linear, index_select, mul, "
                    "scatter_add_, +bias, relu) on the same GPU, bf16, one member at a time x %d" % b,
            "ms": ms_re, "norm": "recomputed", "speedup": ms_re / ms,
            "ms_norm_cached": ms_ca, "speedup_norm_cached": ms_ca / ms,
            "nmax_ours_vs_torch_bf16_sequence": ((y[0].float() - yt.float()).abs().max() / yt.float().abs().max()).item(),
            "note": "the torch sequence rounds every intermediate to bf16 and adds in atomics order, so the "
                    "comparison with it is not a parity statement; parity is the fp32-oracle band above"}
        del yt
    except Exception as e:  # noqa: BLE001
        out["gpu_reference"] = {"error": str(e)[:200]}
    del x, y, model
    gw.clear_graph_cache()
    torch.cuda.empty_cache()
    return out


def probe_train_member_cfg5(dev):
    import torch
    import gwen_b200 as gw
    h, wd, c = 1158, 774, 64
    n = h * wd
    torch.manual_seed(23)
    ei = gw.grid(h, wd, dev)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
    model = gw.GNNModel(cfg).to(dev)          # bf16 activations, fp32 master weights (cast once per weight version)
    x1 = torch.randn(1, n, c, device=dev).to(torch.bfloat16)
    mask = (torch.arange(n, device=dev) % 125) == 124
    ms_t = _timed(lambda: gw.train_step(model, x1, ei, mask), 2, 3)
    del model, x1
    gw.clear_graph_cache()
    torch.cuda.empty_cache()
    return {"workload": "cfg5 shape, one member: gwen_b200.train_step = forward + fused masked-L1 loss (target = input, "
                        "mask id%125==124) + backward through all six layers (tcgen05 dgrad/wgrad, stencil A^T), bf16 "
                        "activations, fp32 master weights and fp32 weight gradients, optimizer step excluded",
            "ms_per_member_step": ms_t, "member_steps_per_s": 1e3 / ms_t}


def probe_general_graphs_cfg2(dev):
    """The aggregation on graphs that are NOT the plain mesh, at cfg 2's size (582 x 390, F = 256, fp32), through the
    public call (``ops.aggregate(kernel="auto")``): the mesh with randomly permuted node ids (K0r locality tiles +
    the staged kernel's row-gather producer; the row kernel on the same graph beside it) and the mesh with 10 % of the
    nodes cut out (masked stencil).  Three rotating buffer pairs, 100-launch regions, result checks in-run."""
    import torch
    import gwen_b200 as gw
    from gwen_b200 import ops
    n = H * W
    peak = measured_peaks()[0]
    ei = gw.grid(H, W, dev)
    gen = torch.Generator(device="cpu").manual_seed(23)
    perm = torch.randperm(n, generator=gen).to(dev)
    g_perm = gw.build_graph(perm[ei].contiguous(), n)
    cut = (torch.rand(n, generator=gen) < 0.1).to(dev)
    g_mask = gw.build_graph(ei[:, ~(cut[ei[0]] | cut[ei[1]])].contiguous(), n)
    xs = [torch.randn(n, FEAT, device=dev) for _ in range(3)]
    outs = [torch.empty(n, FEAT, device=dev) for _ in range(3)]
    bias = torch.randn(FEAT, device=dev) * 0.1
    k = [0]

    def run(graph, kernel):
        def fn():
            for _ in range(100):
                i = k[0] % 3
                k[0] += 1
                ops.aggregate(graph, xs[i], bias, kernel=kernel, out=outs[i].unsqueeze(0))
        return fn

    def entry(graph, kernel, alg):
        us = _timed(run(graph, kernel), 1, 3) * 10.0       # ms per 100 launches -> us per launch
        return {"us_per_launch": us, "GBs": alg / us / 1e3, "frac_of_copy_peak": alg / us / 1e3 / peak}

    res = {"workload": "cfg2 mesh 582x390, F=256 fp32, single-layer message+aggregate on graphs that are not the plain "
                       "mesh; algorithmic bytes = SURVEY 8(d) with 8*E' for CSR graphs (src and w are read)"}
    alg = 2 * n * FEAT * 4 + 4 * (n + 1) + 8 * g_perm.num_messages + 4 * n
    plan = g_perm.locality_plan()
    ent = {"graph_kind": "no grid numbering detected" if g_perm.grid_shape is None else "grid",
           "locality_tiles": None if plan is None else {"tiles": plan.num_tiles, "max_tile_rows": plan.max_tile_rows,
                                                        "max_tile_sources": plan.max_tile_runs,
                                                        "staged_rows_per_destination_row": plan.amplification},
           "algorithmic_bytes": alg, "auto (locality tiles)": entry(g_perm, "auto", alg),
           "row kernel": entry(g_perm, "rows", alg)}
    ent["auto_bitwise_equals_row_kernel"] = bool(torch.equal(ops.aggregate(g_perm, xs[0], bias),
                                                             ops.aggregate(g_perm, xs[0], bias, kernel="rows")))
    res["permuted_node_ids"] = ent
    alg_m = 2 * n * FEAT * 4 + 4 * (n + 1) + 4 * g_mask.num_messages + 4 * n
    ent = {"mesh_kind": g_mask.mesh_kind, "cut_nodes": int(cut.sum()), "algorithmic_bytes": alg_m,
           "auto (masked stencil)": entry(g_mask, "auto", alg_m), "row kernel": entry(g_mask, "rows", alg_m)}
    a, b = ops.aggregate(g_mask, xs[0], bias), ops.aggregate(g_mask, xs[0], bias, kernel="rows")
    ent["nmax_vs_row_kernel"] = ((a - b).abs().max() / b.abs().max()).item()
    res["masked_mesh_10pct"] = ent
    del xs, outs
    gw.clear_graph_cache()
    torch.cuda.empty_cache()
    return res


def probe_cfg4_single(dev):
    """The N = 1 point of BASELINE config 4's strong-scaling family: the un-partitioned model on the whole
    2048 x 2048 mesh (B = 1, bf16)."""
    import torch
    import gwen_b200 as gw
    h = wd = 2048
    n, c = h * wd, 64
    torch.manual_seed(23)
    ei = gw.grid(h, wd, dev)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
    model = gw.GNNModel(cfg).to(dev).to(torch.bfloat16)
    x = torch.randn(1, n, c, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        ms_f = _timed(lambda: model(x, ei), 2, 5)

    def train():
        for p in model.parameters():
            p.grad = None
        model(x, ei).float().abs().mean().backward()
    ms_t = _timed(train, 1, 3)
    del model, x
    gw.clear_graph_cache()
    torch.cuda.empty_cache()
    return {"workload": "cfg4: 2048x2048 mesh (4194304 nodes), six layers, bf16, B=1, FIXED global size (strong scaling); "
                        "N=1: un-partitioned model", "n_gpus": 1, "fwd_ms": ms_f, "fwd_bwd_allreduce_ms": ms_t,
            "note": "N=1 has no all-reduce; the N>1 bench lines carry the same keys for the same global mesh"}


# ---------------------------------------------------------------------------------------------
# other_configs, N > 1 (collective: every rank calls these)
# ---------------------------------------------------------------------------------------------
def _timed_dist(fn, warm, iters, dev):
    import torch
    import torch.distributed as dist
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def _graphed(fn, dev):
    """fn captured in a CUDA graph (all ranks' launches queued ahead of the GPUs: the partitioned step is
    ~100 short kernels per rank and becomes launch-bound at N = 8).  Returns (callable, mode string)."""
    import torch
    try:
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream())
        gobj = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            with torch.cuda.graph(gobj, stream=cap):
                fn()
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        return gobj.replay, "CUDA graph replay"
    except Exception as e:  # noqa: BLE001
        print("bench: graph capture failed (%s); eager launches" % str(e)[:200], file=sys.stderr)
        torch.cuda.synchronize()
        return fn, "eager launches (graph capture failed)"


def probe_strong_cfg4(dev, world, rank):
    import torch
    import torch.distributed as dist
    import gwen_b200 as gw
    from gwen_b200 import partition
    h = wd = 2048
    n, c = h * wd, 64
    gw.clear_graph_cache()
    ei = gw.grid(h, wd, dev)
    g = gw.get_graph(ei, n)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
    torch.manual_seed(23)
    model = gw.GNNModel(cfg).to(dev).to(torch.bfloat16)
    band = partition.PeerMeshBand(h, wd, g.dis)
    net = partition.BandGNNModel(model, band)
    del ei, g
    gw.clear_graph_cache()
    xo = torch.randn(1, band.n_own, c, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        ms_f = _timed_dist(lambda: net(xo), 3, 5, dev)

    def train():
        for p in model.parameters():
            p.grad = None
        net(xo).float().abs().mean().backward()
        net.allreduce_grads()
    ms_t = _timed_dist(train, 2, 3, dev)
    res = {"workload": "cfg4: 2048x2048 mesh (4194304 nodes), six layers, bf16, B=1, FIXED global size (strong scaling), "
                       "row bands over %d GPUs, halo rows pulled over NVLink inside the aggregation kernels, weight "
                       "gradients written into one flat fp32 buffer during backward and all-reduced in place (%s)" % (
                           world, "one NCCL call after backward" if os.environ.get("GWEN_GRAD_BUCKETS", "flat") != "layer"
                           else "per layer, under the remaining backward"),
           "n_gpus": world, "fwd_ms": ms_f, "fwd_bwd_allreduce_ms": ms_t, "step_launch": "eager launches",
           "peer_error_word": band.error_word()}
    # the same step replayed from a CUDA graph (no host launch latency between the ~100 short kernels)
    try:
        with torch.no_grad():
            fwd_g, mode_f = _graphed(lambda: net(xo), dev)
            if mode_f.startswith("CUDA graph"):
                res["fwd_ms_graph"] = _timed_dist(fwd_g, 2, 5, dev)
    except Exception as e:  # noqa: BLE001
        res["graph_error"] = str(e)[:200]
    del model, net, band, xo
    torch.cuda.empty_cache()
    dist.barrier()
    return res


def probe_train_cfg5(dev, world, rank, batch=21):
    import torch
    import torch.distributed as dist
    import gwen_b200 as gw
    from gwen_b200 import partition
    h, wd, c = 1158, 774, 64
    n = h * wd
    gw.clear_graph_cache()
    ei = gw.grid(h, wd, dev)
    g = gw.get_graph(ei, n)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
    torch.manual_seed(23)
    model = gw.GNNModel(cfg).to(dev)                  # fp32 master weights, bf16 activations
    band = partition.PeerMeshBand(h, wd, g.dis)
    net = partition.BandGNNModel(model, band)
    del ei, g
    gw.clear_graph_cache()
    xo = torch.randn(batch, band.n_own, c, device=dev).to(torch.bfloat16)
    ids = torch.arange(band.r0 * wd, (band.r0 + band.rows) * wd, device=dev)
    mo = (ids % 125) == 124

    def train():
        for p in model.parameters():
            p.grad = None
        net.loss(net(xo), xo, mo).backward()
        net.allreduce_grads()
    ms_t = _timed_dist(train, 2, 3, dev)
    res = {"workload": "cfg5: COSMO-1E grid 1158x774, training step (forward + masked L1 vs the input + backward + "
                       "gradient all-reduce), %d-member ensemble batch, bf16 activations / fp32 master weights, mesh "
                       "partitioned into row bands over %d GPUs; optimizer step excluded" % (batch, world),
           "n_gpus": world, "train_step_ms": ms_t, "member_steps_per_s": batch * 1e3 / ms_t,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "peer_error_word": band.error_word()}
    del model, net, band, xo
    torch.cuda.empty_cache()
    dist.barrier()
    return res


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import gwen_b200 as gw
    from gwen_b200 import graph as gwgraph
    from gwen_b200 import ops, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: keep library banners (NCCL prints its version to
    # stdout) away from it by pointing fd 1 at stderr and writing the result to the saved fd.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    gen = torch.Generator().manual_seed(23 + rank)
    bias = (torch.randn(FEAT, generator=torch.Generator().manual_seed(23)) * 0.1).to(dev)
    # ---- graph: global mesh of (H * world) x W, this rank's band ---------------------------
    gh = H * world
    ei = gw.grid(gh, W, dev)
    g_global = gw.build_graph(ei, gh * W)
    launches_per_step = 1
    band = None
    halo_mode = None
    partition_parity = None
    if world == 1:
        graph, n_own = g_global, H * W
        msgs_local = graph.num_messages
        assert graph.is_plain_mesh
        x0 = torch.randn(n_own, FEAT, generator=gen).to(dev)
        xs = [x0] + [x0.clone() for _ in range(NBUF - 1)]
        x_own = x0
    else:
        halo_mode = args.halo
        if halo_mode == "peer":
            try:      # CUDA symmetric memory (peer mappings of the neighbours' buffers)
                band = partition.PeerMeshBand(gh, W, g_global.dis)
                xs = [band.alloc(1, FEAT, torch.float32, dev) for _ in range(NBUF)]
            except Exception as e:  # noqa: BLE001  (uniform across ranks: same driver / same box)
                print("bench: symmetric memory unavailable (%s); using the NCCL halo exchange" % str(e)[:200],
                      file=sys.stderr)
                halo_mode, band = "nccl", None
        if band is None:
            band = partition.MeshBand(gh, W, g_global.dis)
            xs = [band.alloc(1, FEAT, torch.float32, dev) for _ in range(NBUF)]
        n_own = band.n_own
        ranges = partition.band_ranges(gh, W, world)
        rp = g_global.rowptr
        msgs_local = int((rp[ranges[rank].stop] - rp[ranges[rank].start]).item())
        # peer: ONE launch (halo fetch inside); nccl: interior + first-row + last-row launches
        launches_per_step = 1 if halo_mode == "peer" else 3
        x_own = band.owned(xs[0][0])
        x_own.copy_(torch.randn(n_own, FEAT, generator=gen).to(dev))
        for t in xs[1:]:
            band.owned(t[0]).copy_(x_own)
        # ---- partition parity: this rank's band of the partitioned aggregation, BITWISE against the
        # single-GPU stencil on the same global input.  The neighbours' boundary rows travel by NCCL
        # send/recv here (not through the kernel under test); the single-GPU launch sees the rows
        # [r0 - 1, r0 + rows + 1) of the global mesh with the GLOBAL dis.
        torch.cuda.synchronize()
        dist.barrier()
        y_band = band.aggregate(xs[0], bias).clone()
        rows = band.rows
        chk = torch.zeros((rows + 2) * W, FEAT, device=dev)
        chk[W:(rows + 1) * W].copy_(x_own)
        p2p = []
        if band.up is not None:
            p2p += [dist.P2POp(dist.isend, x_own[:W].contiguous(), band.up), dist.P2POp(dist.irecv, chk[:W], band.up)]
        if band.down is not None:
            p2p += [dist.P2POp(dist.isend, x_own[(rows - 1) * W:].contiguous(), band.down),
                    dist.P2POp(dist.irecv, chk[(rows + 1) * W:], band.down)]
        for req in dist.batch_isend_irecv(p2p):
            req.wait()
        d2 = g_global.dis.view(gh, W)
        loc = torch.zeros((rows + 2, W), device=dev)
        lo, hi = max(band.r0 - 1, 0), min(band.r0 + rows + 1, gh)
        loc[lo - (band.r0 - 1):hi - (band.r0 - 1)] = d2[lo:hi]
        y_single = ops.mesh_stencil(chk.unsqueeze(0), gwgraph.bordered_dis(loc), rows + 2, rows, W, 1, bias=bias)
        okt = torch.tensor([1.0 if torch.equal(y_single.view(-1), y_band.view(-1)) else 0.0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        partition_parity = bool(okt.item() == 1.0)
        del chk, y_single, y_band, loc
        del g_global, ei
    outs = [torch.empty(n_own, FEAT, device=dev) for _ in range(NBUF)]

    def make_step(i):
        if band is not None:
            return lambda: band.aggregate(xs[i], bias, out=outs[i])
        return lambda: ops.aggregate(graph, xs[i], bias, kernel="stencil", out=outs[i])

    steps_fn = [make_step(i) for i in range(NBUF)]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        steps_fn[i % NBUF]()
    sync_all()
    step_mode = "eager launches"
    if band is not None:
        # Capture each buffer's partitioned step once in a CUDA graph and replay it, so that all ranks'
        # launches are queued ahead of the GPU (peer: the ranks' kernels wait on one another's flags;
        # nccl: the step is ~8 host-side operations for ~100 us of GPU work).
        graphed, modes = [], []
        for fn in steps_fn:
            g_fn, mode = _graphed(fn, dev)
            graphed.append(g_fn)
            modes.append(mode)
        if all(m.startswith("CUDA graph") for m in modes):
            steps_fn = graphed
            step_mode = "CUDA graph replay of the partitioned step (one graph per buffer pair)"
            for i in range(NBUF):
                steps_fn[i]()
        else:
            step_mode = "eager launches (graph capture failed)"
        sync_all()

    def timed_region(k):
        sync_all()
        if world > 1:
            # the ranks leave the barrier up to ~1 ms apart, and a partitioned step cannot finish before its
            # neighbours have started theirs: without this the first timed step absorbs that skew (a 20-step region
            # is 1.7 ms).  One more UNTIMED step lines the ranks up on the device (each waits for its neighbours'
            # epoch flags); the start event is recorded behind it on the stream.
            steps_fn[NBUF - 1]()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            steps_fn[i % NBUF]()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    t_wall0 = time.time()
    ms_total = timed_region(args.steps)
    tot = torch.tensor([float(msgs_local)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_msgs = tot.item()
    value = total_msgs * args.steps / (ms_total * 1e-3)
    long_run = None
    if args.steps < LONG_STEPS:
        ms_long = timed_region(LONG_STEPS)
        long_run = {"steps": LONG_STEPS, "ms_per_step": ms_long / LONG_STEPS,
                    "value": total_msgs * LONG_STEPS / (ms_long * 1e-3),
                    "note": "a second, longer timed region of the same step (the K-step region above is %.1f ms)" % ms_total}

    # ---- launch duration of the dominant kernel: a timed region is back-to-back launches of it on the
    # launching stream (one stencil launch per step at N = 1 and in peer mode), so its average duration is
    # the CUDA-event time of the region / steps; the longer region is used.  (Events around single launches
    # would add the host launch latency to every sample.)  nccl mode: 3 launches per step, timed
    # separately below on the whole band.
    if launches_per_step == 1:
        k_us = (long_run["ms_per_step"] if long_run else ms_total / args.steps) * 1e3
    else:
        per = []
        for _ in range(min(50, max(10, args.steps))):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.mesh_stencil(xs[0][0, :band.n_local], band.dis, band.rows + 2, band.rows, W, 1, bias=bias, out=outs[0])
            b.record()
            per.append((a, b))
        torch.cuda.synchronize()
        k_us = statistics.mean(a.elapsed_time(b) for a, b in per) * 1e3
    # 2 N F s (rows once in, once out) + 4 (N + 1) rowptr + 4 E' src + 4 N dis
    alg_bytes = 2 * n_own * FEAT * 4 + 4 * (n_own + 1) + 4 * msgs_local + 4 * n_own
    # what the mesh-stencil kernel itself has to move (it never reads rowptr / src): rows + dis + bias
    moved_bytes = 2 * n_own * FEAT * 4 + 4 * n_own + 4 * FEAT
    peak, _, peak_src = measured_peaks()
    achieved = alg_bytes / (k_us * 1e-6) / 1e9

    # ---- e2e: public API with pinned host buffers, H2D + D2H inside the timed region ----------
    if world > 1:
        x_host = gw.pinned_near_gpu((n_own, FEAT), torch.float32, local_rank)   # pages on the GPU's NUMA node
        out_host = gw.pinned_near_gpu((n_own, FEAT), torch.float32, local_rank)
    else:
        x_host = torch.empty(n_own, FEAT).pin_memory()
        out_host = torch.empty(n_own, FEAT).pin_memory()
    x_host.copy_(x_own)
    conv = gw.GCNConv(FEAT, FEAT).to(dev)
    with torch.no_grad():
        conv.bias.copy_(bias)
    e2e_api = None
    if band is None:
        # chunked H2D -> stencil -> D2H pipeline (gwen_b200/host_stream.py): the two PCIe
        # directions and the kernel overlap, within a call and across calls
        host_prop = gw.HostPropagator(graph, FEAT, torch.float32, chunks=8)
        e2e_api = "gwen_b200.HostPropagator(graph, F)(x_host, out_host, bias): pinned host x / out, 8 row chunks, H2D / aggregate / D2H overlapped"
    elif halo_mode == "peer":
        host_prop = gw.HostBandPropagator(band, FEAT, torch.float32, chunks=8)
        e2e_api = ("gwen_b200.HostBandPropagator(band, F)(x_host, out_host, bias): pinned host buffers on the GPU's NUMA node, "
                   "8 row chunks, H2D / sub-range stencil / D2H overlapped, first+last band row from one peer launch")
    else:
        host_prop = None
        e2e_api = "MeshBand.aggregate between a plain H2D and D2H"
    e2e_i = [0]

    def e2e_step():
        if host_prop is not None:
            host_prop(x_host, out_host, conv.bias)
        else:
            xb = xs[e2e_i[0] % NBUF]
            e2e_i[0] += 1
            band.owned(xb[0]).copy_(x_host, non_blocking=True)
            y = band.aggregate(xb, conv.bias)
            out_host.copy_(y.view(n_own, FEAT), non_blocking=True)

    for _ in range(3):
        e2e_step()
    sync_all()
    e2e_ok = None
    if True:  # the e2e result must be the device-path result (bitwise): checked once, outside the timed region
        ref_out = (band.aggregate(xs[0], conv.bias) if band is not None
                   else ops.aggregate(graph, xs[0], conv.bias, kernel="stencil")).view(n_own, FEAT)
        torch.cuda.synchronize()
        e2e_ok = bool(torch.equal(out_host, ref_out.cpu()))
        del ref_out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.e2e_steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    ms_e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e_value = total_msgs * args.e2e_steps / (ms_e.item() * 1e-3)
    t_wall1 = time.time()
    # where every rank's pinned buffers sit: (NUMA node of its GPU's PCIe root, CPUs of that node / CPUs allowed)
    from gwen_b200.host_stream import gpu_numa_cpus
    numa = gpu_numa_cpus(local_rank)
    numa_info = {"rank": rank, "gpu_numa_node": None if numa is None else numa[0],
                 "node_cpus": None if numa is None else len(numa[1]),
                 "allowed_cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
                 "e2e_ms_per_step": None}
    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0_.record()
    for _ in range(args.e2e_steps):
        e2e_step()
    e1_.record()
    torch.cuda.synchronize()
    numa_info["e2e_ms_per_step"] = e0_.elapsed_time(e1_) / args.e2e_steps      # this rank's own time (not the max)
    if world > 1:
        all_numa = [None] * world
        dist.all_gather_object(all_numa, numa_info)
    else:
        all_numa = [numa_info]
    del host_prop
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- other_configs (collective at N > 1) ---------------------------------------------------
    other = {}
    if not args.no_model_probes:
        del xs, outs, steps_fn
        if world == 1:
            del x0
        torch.cuda.empty_cache()
        probes = []
        if world == 1:
            probes = [("full_forward_cfg3", lambda: probe_full_forward_cfg3(dev)),
                      ("train_step_cfg5_member", lambda: probe_train_member_cfg5(dev)),
                      ("strong_cfg4", lambda: probe_cfg4_single(dev)),
                      ("general_graphs_cfg2", lambda: probe_general_graphs_cfg2(dev))]
        elif halo_mode == "peer":
            probes = [("strong_cfg4", lambda: probe_strong_cfg4(dev, world, rank))]
            if world >= 4:
                probes.append(("train_step_cfg5", lambda: probe_train_cfg5(dev, world, rank)))
        for name, fn in probes:
            try:
                other[name] = fn()
            except Exception as e:  # noqa: BLE001  (context only: never fail the headline line)
                other[name] = {"error": str(e)[:300]}
                if world > 1:
                    break           # a collective probe failed on this rank: do not start the next one

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD if world == 1 else WORKLOAD + "; weak scaling: one 582x390 "
                       "row band per rank of a %dx390 mesh, one-row halo exchange per step (%s)" % (gh, "inside the aggregation kernel: one warp per CTA pulls the neighbours' boundary rows over NVLink peer memory under the interior tiles, device-side flags" if halo_mode == "peer" else "NCCL send/recv on a side stream under the interior rows"),
                       "l2": "the step rotates over %d input/output buffer pairs (%.0f MB per rank > 126 MB L2): no line survives "
                             "in L2 between two uses of a buffer, no explicit flush" % (NBUF, NBUF * 2 * n_own * FEAT * 4 / 1e6),
                       "step_launch": step_mode,
                       "timing": "CUDA events on the launching stream around exactly K steps, max over ranks, barrier + "
                                 "synchronize on both sides" + ("; at N > 1 one untimed step behind the opening barrier "
                                 "lines the ranks up on the device, so the region does not include the barrier's exit skew"
                                 if world > 1 else ""),
                       "kernel": "k_grid_stencil (mesh fast path: 8x16 tiles, one 4-D TMA box per tile/slab, "
                                 "separable column sums in packed fp32x2 registers); exact CSR kernels "
                                 "k_agg_tiled / k_agg_rows remain for arbitrary graphs"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 * world,
                    "d2h_bytes_per_step": out_host.numel() * 4 * world, "steps": args.e2e_steps,
                    "gbs_per_direction_per_gpu": x_host.numel() * 4 * args.e2e_steps / (ms_e.item() * 1e-3) / 1e9,
                    "result_equals_device_path": e2e_ok, "api": e2e_api, "per_rank": all_numa,
                    "numa_binding": "off (GWEN_NO_NUMA)" if os.environ.get("GWEN_NO_NUMA") else "pinned buffers first-touched on the GPU's NUMA node"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"kernel": "k_grid_stencil<float,16>", "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "peak_source": peak_src,
                         "us_per_launch": k_us, "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_formula": "SURVEY 8(d): 2*N*F*4 + 4*(N+1) + 4*E' + 4*N",
                         "bytes_moved_by_kernel": moved_bytes,
                         "frac_bytes_moved_by_kernel": moved_bytes / (k_us * 1e-6) / 1e9 / peak,
                         "bytes_moved_formula": "2*N*F*4 + 4*N (dis) + 4*F (bias): the mesh kernel never reads rowptr/src",
                         "traffic": ncu_traffic()[0] if world == 1 else None,
                         "traffic_source": ncu_traffic()[1]},
        }
        if long_run is not None:
            line["long_run"] = long_run
        if partition_parity is not None:
            line["partition_parity"] = partition_parity
            line["partition_parity_what"] = ("every rank's band of the partitioned aggregation is bitwise equal to the "
                                             "single-GPU stencil on the same global input (neighbour rows sent by NCCL for the check)")
        if other:
            line["other_configs"] = other
        if world == 1 and not args.no_cpu_baseline:
            rate, per_step, edges, cores = cpu_propagate_rate(H, 3, warmup=1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "3 full-mesh steps (582x390, F=256) of the oracle port, "
                                              "%.2f s/step" % per_step}
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # captured NCCL work + process-group teardown can deadlock at interpreter exit: leave hard
        torch.cuda.synchronize()
        dist.barrier()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
