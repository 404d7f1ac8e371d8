#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its configs.

Headline (default): molecular graphs/sec, forward+backward(+Adam) of GLAM-GP on 4096-graph synthetic
MoleculeNet-shaped batches (BASELINE.json configs[1]; ~25 atoms, ~54 directed bonds, 9-dim atom /
3-dim bond features, C = 36, H = 3, 3 message steps, Set2Set readout, e_dim 1024).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (one process per GPU under torchrun)
  python bench.py --workload {gp,ddi,dti,screen}               which config is the headline of the line (default gp)
  python bench.py --impl reference ...                         the reference's CPU path (oracle port) on host cores

The default run also measures, in the same process and inside the same JSON line (`config.also_measured`,
`roofline.modules`): the other configs — GLAM-DDI pairs (configs[2]), GLAM-DTI ligand + ~500-residue protein
(configs[3]), screening inference streamed in 64k-graph batches up to >= 10 M molecules per job (configs[4]) — the CPU
config (configs[0]) as a cpu_baseline on its full 128-graph shape, the same training step in exact-fp32 math mode and
stepped through the reference's own per-step loop, and module-level rooflines (SURVEY.md §8d bytes / measured time)
of the hot-path modules.  One JSON line on stdout from rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "graphs/s"
DIMS = dict(node_dim=9, edge_dim=3)
C, H, DE = 36, 3, 3
GRAPHS = 4096
NODES_PER, EDGES_PER = 25, 54
TOTAL_NODES, TOTAL_EDGES = NODES_PER * GRAPHS, EDGES_PER * GRAPHS      # 102 400 / 221 184 (config 2)
MODEL_KW = dict(hid_dim_alpha=4, e_dim=1024, out_dim=1, mol_block="_TripletMessage", message_steps=3,
                mol_readout="Set2Set", pre_act="ReLU", graph_act="CELU", flat_act="ReLU")
NO_DROPOUT = dict(graph_do="_None()", flat_do="_None()", end_do="_None()")
OVERLAP_ALLREDUCE = False            # engine.TrainStep's opt-in early gradient bucket (measured slower on 8 GPUs: DESIGN.md §6)
N_RESIDENT = 16                      # distinct batches rotated through (16 x 10.7 MB inputs; each step also writes
                                     # ~0.7 GB of activations) -> nothing survives in the 126 MB L2 between steps
SCREEN_GRAPHS = 65536                # configs[4]: 64k-graph batches
SCREEN_TOTAL = 10_000_000            # molecules per job
DDI_PAIRS, DTI_PAIRS = 4096, 256
METRICS = {"gp": "molecular graphs/sec fwd+bwd (GLAM-GP training step)",
           "ddi": "drug pairs/sec fwd+bwd (GLAM-DDI training step)",
           "dti": "ligand-protein pairs/sec fwd+bwd (GLAM-DTI training step)",
           "screen": "molecular graphs/sec screening inference (GLAM-GP eval forward)"}
DTYPE = "tf32 operands / f32 accumulate (projections); f32 everywhere else"


def workload_config(world, which="gp"):
    par = (f"dp{world} (graphs sharded by molecule; one NCCL all-reduce of a flat fp32 grad bucket)" if world > 1 else "single GPU")
    base = {"gp": {"workload": "GLAM-GP training step, 4096-graph synthetic MoleculeNet-shaped batches per GPU "
                               "(BASELINE.json configs[1])",
                   "graphs_per_gpu_batch": GRAPHS, "nodes": TOTAL_NODES, "edges": TOTAL_EDGES, **DIMS, "hidden": C, "heads": H,
                   "message_steps": 3, "readout": "Set2Set", "e_dim": 1024, "loss": "mse", "optimizer": "Adam",
                   "graph_order": "tile-aware collation (glam_b200.synth.tile_order: the batch's graphs are laid out so that "
                                  "consecutive graphs fill the kernels' 128-row tiles; the random order is under also_measured)",
                   "l2": f"inputs rotate over {N_RESIDENT} distinct batches (171 MB) and every step streams ~0.6 GB of "
                         "activations: larger than the 126 MB L2"},
            "ddi": {"workload": "GLAM-DDI training step, 4096 synthetic DrugBank-shaped drug pairs per GPU (BASELINE.json configs[2])",
                    "pairs_per_gpu_batch": DDI_PAIRS, **DIMS, "hidden": C, "message_steps": 3, "readout": "Set2Set", "loss": "bce"},
            "dti": {"workload": "GLAM-DTI training step, 256 ligand + ~500-residue protein pairs per GPU (BASELINE.json configs[3])",
                    "pairs_per_gpu_batch": DTI_PAIRS, **DIMS, "protein_dim": 49, "protein_edge_dim": 8, "hidden": C,
                    "pro_block": "_GCNConv", "readout": "GlobalPool5", "loss": "cross_entropy"},
            "screen": {"workload": "LIT-PCBA-scale screening inference, 64k-graph batches sharded by molecule "
                                   "(BASELINE.json configs[4])",
                       "graphs_per_gpu_batch": SCREEN_GRAPHS, "molecules_per_job": SCREEN_TOTAL, **DIMS, "hidden": C}}[which]
    return {**base, "parallelism": par}


TILE_PACK = True     # batches in the tile-filling graph order (synth.tile_order): same molecules, order chosen at collation


def make_batches(n, rank, pin, graphs=GRAPHS, seed0=1234, targets="regression", tile_pack=None):
    from glam_b200.synth import make_molecule_batch
    out = []
    for i in range(n):
        b = make_molecule_batch(graphs, seed=seed0 + 1000 * rank + i, total_nodes=NODES_PER * graphs, total_edges=EDGES_PER * graphs,
                                targets=targets, tile_pack=TILE_PACK if tile_pack is None else tile_pack, **DIMS)
        out.append(b.pin_memory() if pin else b)
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.stamps = index, None, [], []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())
            self.stamps.append(time.time())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def wait_first_sample(self, timeout=5.0):
        """Block until nvidia-smi has produced its first line (NVML initialisation takes a driver-wide lock for a few
        hundred ms: it must be over before the timed region starts) or `timeout` seconds have passed."""
        if self.proc is None:
            return
        t0 = time.time()
        while not self.lines and time.time() - t0 < timeout and self.proc.poll() is None:
            time.sleep(0.02)

    def summary(self, t0=None, t1=None):
        """Median SM clock / reasons over the samples that arrived in [t0, t1] (the timed region); all samples if none did."""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = list(self.lines)
        if t0 is not None:
            inside = [ln for ln, ts in zip(lines, list(self.stamps)) if t0 <= ts <= t1 + 0.06]
            lines = inside or lines
        for ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def oracle_model(seed=0):
    from oracle import glam_oracle as O
    torch.manual_seed(seed)
    return O.ArchitectureGP(DIMS["node_dim"], DIMS["edge_dim"], **MODEL_KW)


def cpu_train_throughput(sample_graphs, steps, warmup):
    """The reference's unfused CPU path (oracle port): fwd + MSE + bwd + Adam on `sample_graphs`-graph batches."""
    from glam_b200.synth import make_molecule_batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = oracle_model().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    batches = [make_molecule_batch(sample_graphs, seed=77 + i, **DIMS) for i in range(2)]

    def one(b):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(b), b.y)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for i in range(warmup):
        one(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        one(batches[i % 2])
    dt = time.perf_counter() - t0
    return sample_graphs * steps / dt, dt / steps, threads


def cpu_screen_throughput(sample_graphs, steps, warmup):
    from glam_b200.synth import make_molecule_batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = oracle_model().eval()
    b = make_molecule_batch(sample_graphs, seed=78, **DIMS)
    with torch.no_grad():
        for _ in range(warmup):
            model(b)
        t0 = time.perf_counter()
        for _ in range(steps):
            model(b)
        dt = time.perf_counter() - t0
    return sample_graphs * steps / dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 512
    if args.workload == "screen":
        gps, threads = cpu_screen_throughput(2048, args.steps, max(args.warmup, 1))
        sec = 2048 / gps
        sample_txt = "2048-graph eval-mode forward passes of the oracle per step"
    else:
        gps, sec, threads = cpu_train_throughput(sample, args.steps, max(args.warmup, 1))
        sample_txt = (f"{sample}-graph slices of the 4096-graph workload per step (larger CPU batches were slower per graph)")
    cfg = workload_config(args.gpus, args.workload if args.workload in ("gp", "screen") else "gp")
    cfg["reference_arm_batch"] = sample_txt
    line = {"impl": "reference", "metric": METRICS["screen" if args.workload == "screen" else "gp"], "value": gps, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": gps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample_txt + " (oracle/glam_oracle.py: pure-PyTorch restatement of the reference's unfused "
                                                    "layer.py path; the reference itself needs torch_geometric, absent from this image)"},
            "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ------------------------------------------------------------------------------------------------ bytes models
def module_bytes(name, graphs, n=NODES_PER, e=EDGES_PER, Cc=C, De=DE):
    """SURVEY.md §8(d) / BASELINE.md §3 algorithmic bytes of one call on `graphs` molecule graphs (fp32 activations, int32
    CSR, parameters amortised to zero)."""
    per = {"TripletMessage fwd": 4 * (2 * n * Cc + e * De) + 8 * e + 4 * (n + 1),
           "TripletMessage bwd": 4 * (3 * n * Cc + e * De) + 8 * e + 4 * (n + 1),
           "MessageBlock fwd": 4 * (4 * n * Cc + e * De) + 8 * e + 4 * (n + 1),
           # by analogy with SURVEY 8(d)'s recompute-based backward rows: x, h in; g_x', g_h' in; g_x, g_h out (+ the graph)
           "MessageBlock bwd": 4 * (6 * n * Cc + e * De) + 8 * e + 4 * (n + 1),
           "GRU update": 12 * n * Cc,
           "Set2Set step": 4 * n * Cc + 4 * n + 16 * Cc,
           "GlobalAttention": 4 * n * Cc + 4 * n + 8 * Cc}[name]
    return per * graphs


def kernel_bytes_model(label, N, E, B, Cc=C, Hh=H, De=DE):
    """Bytes one launch of a per-op library call has to move (every operand read or written exactly once)."""
    import re
    HC, ld = Hh * Cc, (Hh * Cc + 2 * Hh + 3) // 4 * 4
    if label.startswith("glam_message_stack_fwd"):
        m = re.search(r"N=(\d+),S=(\d+),(\w+)", label)
        S, save = int(m.group(2)), m.group(3) == "save"
        out = 4 * N * Cc * (2 if not save else 0)
        # saved for backward per step and node: xpe, agg, tile-blocked gates (7C), m | h | 1 rows (2C + 4), x and h outputs
        sv = 4 * S * (N * (ld + HC + 7 * Cc + 2 * Cc + 4 + 2 * Cc) + E * Hh) + 8 * N * Cc if save else 0
        return 4 * N * Cc + 5 * E + 4 * N + out + sv
    if label.startswith("glam_message_stack_bwd"):   # reads the tile-blocked gate save, xpe, alpha; writes G4, G_PRE, G_XPE
        S = int(re.search(r"S=(\d+)", label).group(1))
        return 4 * S * (N * (7 * Cc + ld) + E * Hh + N * (4 * Cc + Cc + ld)) + 4 * 2 * N * Cc + 4 * (3 * E + 2 * N)
    if label.startswith("glam_triplet_edge_fwd"):
        return 4 * (N * ld + N * HC + E * De + E * Hh) + 4 * (E + N + 1)
    if label.startswith("glam_triplet_edge_bwd_dst"):
        return 4 * (N * ld + N * HC + E * De + 2 * E * Hh + N * Hh) + 4 * (E + N + 1)
    if label.startswith("glam_triplet_edge_bwd_src"):
        return 4 * (N * HC + N * ld + E * De + 2 * E * Hh) + 4 * (3 * E + N + 1)
    if label.startswith("glam_gemm_tn"):  # incl. glam_gemm_tn_ex
        m = re.search(r"M=(\d+),Ka=(\d+),Kb=(\d+)", label)
        M, Ka, Kb = map(int, m.groups())
        return 4 * (M * Ka + M * Kb + Ka * Kb)
    if label.startswith("glam_gemm_ex["):
        m = re.search(r"M=(\d+),N=(\d+),K=(\d+),\w+,epi=(\d)", label)
        M, Nn, K, epi = map(int, m.groups())
        extra = M * Nn if epi in (2, 3) else 0
        return 4 * (M * K + M * Nn + extra)
    if label.startswith("glam_colsum"):
        m = re.search(r"M=(\d+),N=(\d+)", label)
        M, Nn = map(int, m.groups())
        return 4 * (M * Nn + Nn)
    if label.startswith("glam_gru_fused_fwd"):       # in: m, h, identity; out: r|z|n (3C), gh_n, h_new, x_out
        return 4 * N * Cc * (3 + 3 + 1 + 2)
    if label.startswith("glam_gru_gates_fwd"):
        return 4 * N * Cc * (6 + 2 + 3 + 2)
    if label.startswith("glam_gru_gates_bwd"):
        return 4 * N * Cc * (3 + 1 + 1 + 1 + 2 + 6 + 2)
    return None


def kernel_family(label):
    """ncu kernel name behind a profiled library call (so that shape-suffixed labels do not split one kernel's share)."""
    for prefix, kern in (("glam_message_stack_fwd", "mp_fused_kernel"), ("glam_message_stack_bwd", "mp_fused_bwd_kernel"), ("glam_gemm_tn", "tc_gemm_tn_kernel"), ("glam_gemm_ex", "tc_gemm_kernel"),
                         ("glam_triplet_edge_bwd_dst", "edge_win_bwd_dst2_kernel"), ("glam_triplet_edge_bwd_src", "edge_win_bwd_src_kernel"),
                         ("glam_triplet_edge_fwd", "edge_win2_fwd_kernel"), ("glam_gru_gates_bwd", "gru_gates_bwd_vec"),
                         ("glam_gru_fused_fwd", "tc_gru_fwd_kernel"), ("glam_set2set_round_fwd", "set2set_round_fwd_rows_kernel"),
                         ("glam_set2set_round_bwd", "set2set_round_bwd_rows_kernel")):
        if label.startswith(prefix):
            return kern
    return label.split("[")[0]


def profile_kernels(ts, dev_batches, reps):
    """Eager (un-graphed) training steps with a CUDA-event pair around every C-ABI call on the launching stream.
    Each step starts with a ~15 ms device-side sleep so that the host has enqueued the whole step before the GPU
    reaches it: the event intervals are then pure device time, not host launch latency."""
    from glam_b200 import ops, graph as G, functional as Fn
    sink = []
    world, ts.world = ts.world, 1                    # rank-local measurement: no collective here
    overlap, Fn.OVERLAP_WGRAD = Fn.OVERLAP_WGRAD, False   # one stream: every interval is one kernel family, nothing concurrent
    try:
        for i in range(reps):
            ts.load(dev_batches[i % len(dev_batches)])
            G.clear_caches()
            torch.cuda.synchronize()
            torch.cuda._sleep(30_000_000)
            ops.set_profile(sink)
            ts._body(ts.statics[ts._slot])
            ops.set_profile(None)
        torch.cuda.synchronize()
    finally:
        ops.set_profile(None)
        ts.world = world
        Fn.OVERLAP_WGRAD = overlap
    agg = {}
    for name, e0, e1 in sink:
        t, n = agg.get(name, (0.0, 0))
        agg[name] = (t + e0.elapsed_time(e1), n + 1)
    return {k: (t / reps, n // reps, t / n) for k, (t, n) in agg.items()}     # ms per step, launches per step, ms per launch


class Harness:
    """Timing helpers shared by the workloads: device events, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms


def train_workload(hz, net, loss_fn, host, steps, warmup, units, capture=True, extra_warm=0):
    """(resident value, ms, e2e value, ms, h2d bytes, launches per step, final loss, TrainStep) for a list of pinned host batches."""
    from glam_b200 import _lib
    from glam_b200.engine import TrainStep, batch_nbytes
    resident = [tuple(g.to(hz.dev) for g in b) if isinstance(b, tuple) else b.to(hz.dev) for b in host]
    n0 = _lib.launch_count()
    ts = TrainStep(net, loss_fn, resident[0], lr=1e-3, device=hz.dev, world_size=hz.world, use_cuda_graph=capture, warmup=3,
                   double_buffer=True, overlap_allreduce=OVERLAP_ALLREDUCE)
    launches = (_lib.launch_count() - n0) // (6 if ts.overlap_allreduce else 5) if ts.graph is not None else None   # (probe +) 3 warm-ups + 2 captures
    nb = len(host)

    def step_resident(i):
        ts.load(resident[i % nb])
        ts.run_resident()

    for i in range(warmup + extra_warm):
        step_resident(i)
    ms = hz.timed(step_resident, steps) / steps
    losses = []

    def step_e2e(i):
        loss = ts.step(host[i % nb], prefetch=host[(i + 1) % nb])
        losses.append(loss.item())

    for i in range(warmup):
        step_e2e(i)
    ms_e2e = hz.timed(step_e2e, steps) / steps
    return dict(value=units * hz.world / (ms * 1e-3), ms=ms, e2e=units * hz.world / (ms_e2e * 1e-3), ms_e2e=ms_e2e,
                h2d=batch_nbytes(host[0]), launches=launches, loss=losses[-1] if losses else None, ts=ts, resident=resident,
                step_resident=step_resident)


def screening_workload(hz, net, rank, graphs=SCREEN_GRAPHS, total=SCREEN_TOTAL, n_batches=3, warmup=3):
    """BASELINE.json configs[4]: eval-mode forward on 64k-graph batches, graphs sharded by molecule over the ranks with no
    collective, >= `total` molecules streamed per job.  value = batches resident in HBM; e2e = pinned host batches through
    ScreenStep.step(batch, prefetch=next) with a checksum of the scores read back per batch."""
    from glam_b200.engine import ScreenStep
    host = make_batches(n_batches, rank, pin=True, graphs=graphs, seed0=5000)
    dev_b = [b.to(hz.dev) for b in host]
    ss = ScreenStep(net, dev_b[0], device=hz.dev, double_buffer=True)
    steps = max(-(-total // (graphs * hz.world)), 4)               # batches per rank so that the job covers >= total molecules

    def run(fn):
        for i in range(warmup):
            fn(i)
        return hz.timed(fn, steps) / steps

    ms = run(lambda i: ss.step(dev_b[i % n_batches]))
    sums = []
    ms_e2e = run(lambda i: sums.append(ss.step(host[i % n_batches], prefetch=host[(i + 1) % n_batches]).sum().item()))
    # the same molecules from the packed graph store (glam_b200/packed.py): prebuilt dst-sorted index in uint8/uint16, ~7x fewer
    # bytes over PCIe, no per-batch CSR build on the device
    from glam_b200 import packed
    pk_host = [packed.pack_batch(b).pin_memory() for b in host]
    pk_dev = [b.to(hz.dev) for b in pk_host]
    del ss
    sp = ScreenStep(net, pk_dev[0], device=hz.dev, double_buffer=True)
    ms_pk = run(lambda i: sp.step(pk_dev[i % n_batches]))
    ms_pk_e2e = run(lambda i: sums.append(sp.step(pk_host[i % n_batches], prefetch=pk_host[(i + 1) % n_batches]).sum().item()))
    del sp
    # the same job with the reference's default graph_norm (_PairNorm, src_1gp/run.py:28): applied inside the fused kernel
    import copy
    net_pn = copy.deepcopy(net)
    from glam_b200 import layer as _layer
    net_pn.mol_conv.norm = _layer._PairNorm(C)
    spn = ScreenStep(net_pn, dev_b[0], device=hz.dev, double_buffer=True)
    ms_pn = hz.timed(lambda i: spn.step(dev_b[i % n_batches]), 12) / 12
    del spn
    return {"value": graphs * hz.world / (ms * 1e-3), "unit": UNIT, "graphs_per_gpu_batch": graphs, "batches_per_gpu": steps,
            "molecules_per_job": steps * graphs * hz.world, "ms_per_batch": ms,
            "e2e": {"value": graphs * hz.world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_batch": ms_e2e,
                    "h2d_bytes_per_batch": host[0].nbytes(), "d2h_bytes_per_batch": 4},
            "packed_store": {"value": graphs * hz.world / (ms_pk * 1e-3), "unit": UNIT, "ms_per_batch": ms_pk,
                             "e2e": {"value": graphs * hz.world / (ms_pk_e2e * 1e-3), "unit": UNIT, "ms_per_batch": ms_pk_e2e,
                                     "h2d_bytes_per_batch": pk_host[0].nbytes(), "d2h_bytes_per_batch": 4},
                             "note": "inputs from glam_b200.packed.PackedBatch (uint8/uint16 prebuilt index, unpacked on the device "
                                     "inside the captured step); same molecules, bitwise the same scores"},
            "graph_norm=_PairNorm (the reference's default, inside the fused kernel)": {"value": graphs * hz.world / (ms_pn * 1e-3), "unit": UNIT,
                                                                                        "ms_per_batch": ms_pn},
            "note": "eval-mode forward, graphs sharded by molecule, no collective; value = inputs resident in HBM (3 distinct "
                    "171 MB batches rotate: larger than L2), e2e = pinned host batches through ScreenStep.step(batch, prefetch=next) "
                    "+ checksum read-back"}


def make_ddi_batches(n, rank, pairs):
    out = []
    for i in range(n):
        a = make_batches(1, rank, pin=False, graphs=pairs, seed0=9000 + 2 * i, targets="binary")[0]
        b = make_batches(1, rank, pin=False, graphs=pairs, seed0=9001 + 2 * i, tile_pack=False)[0]   # pair order is the first side's
        b.y = None
        out.append((a.pin_memory(), b.pin_memory()))
    return out


def make_dti_batches(n, rank, pairs):
    from glam_b200.synth import make_protein_batch
    out = []
    for i in range(n):
        lig = make_batches(1, rank, pin=False, graphs=pairs, seed0=9500 + i)[0]
        lig.y = torch.randint(0, 2, (pairs,), generator=torch.Generator().manual_seed(i))
        pro = make_protein_batch(pairs, seed=9600 + 1000 * rank + i)
        out.append((lig.pin_memory(), pro.pin_memory()))
    return out


# ------------------------------------------------------------------------------------------------ module rooflines
def module_rooflines(dev, peak):
    """Module-level rooflines at the bench shape: SURVEY.md §8(d) algorithmic bytes / CUDA-event time of the library call(s)
    that implement the module (L2 flushed between launches)."""
    from glam_b200 import graph as G, layer, ops
    from glam_b200.synth import make_molecule_batch, make_protein_batch
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def timeit(fn, reps=7, capture=True):
        # the module's launches are captured once and replayed: a 60 us kernel behind ~100 us of Python (layer dispatch, index
        # cache lookups, ctypes) would otherwise be timed as the host, not the device
        fn(); fn(); torch.cuda.synchronize()
        run = fn
        try:
            if not capture:
                raise RuntimeError("eager")
            cg = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.cuda.graph(cg, stream=side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            run = cg.replay
        except Exception:                                         # not capturable (host read-back inside): eager timing
            torch.cuda.synchronize()
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2] * 1e-3

    out = {}

    def put(name, nbytes, sec, units, unit_name="graphs", note=None):
        out[name] = {"algorithmic_MB": round(nbytes / 1e6, 2), "us": round(sec * 1e6, 1), "GBps": round(nbytes / sec / 1e9, 1),
                     "frac_of_hbm": round(nbytes / sec / 1e9 / peak, 4), f"M_{unit_name}_per_s": round(units / sec / 1e6, 2)}
        if note:
            out[name]["note"] = note

    b = make_molecule_batch(GRAPHS, seed=1234, total_nodes=TOTAL_NODES, total_edges=TOTAL_EDGES, tile_pack=TILE_PACK, **DIMS).to(dev)
    torch.manual_seed(0)
    blk = layer.MessageBlock(C, C, DE, norm="_None", dropout="_None()", conv="_TripletMessage", act="CELU", res=True).to(dev).eval()
    x0 = torch.randn(b.num_nodes, C, device=dev)
    kw = dict(batch=b.batch, num_graphs=b.num_graphs)
    with torch.no_grad():
        conv = blk.conv.conv
        put("TripletMessage fwd (fused kernel, conv only)", module_bytes("TripletMessage fwd", GRAPHS),
            timeit(lambda: conv(x0, b.edge_index, b.edge_attr, **kw)), GRAPHS)
        put("MessageBlock fwd (fused kernel, 1 step)", module_bytes("MessageBlock fwd", GRAPHS),
            timeit(lambda: blk(x0, b.edge_index, b.edge_attr, h=None, **kw)), GRAPHS)
        put("MessageBlock fwd x3 (fused kernel, one launch, eval)", 3 * module_bytes("MessageBlock fwd", GRAPHS),
            timeit(lambda: blk.run_steps(x0, b.edge_index, b.edge_attr, 3, keep="last", **kw)), GRAPHS)
        layer.USE_FUSED_STACK = False
        try:
            put("MessageBlock fwd x3 (per-op kernels, eval)", 3 * module_bytes("MessageBlock fwd", GRAPHS),
                timeit(lambda: blk.run_steps(x0, b.edge_index, b.edge_attr, 3, keep="last", **kw)), GRAPHS)
            put("TripletMessage fwd (per-op kernels)", module_bytes("TripletMessage fwd", GRAPHS),
                timeit(lambda: conv(x0, b.edge_index, b.edge_attr)), GRAPHS)
        finally:
            layer.USE_FUSED_STACK = True
    # backward of the 3-step stack: gradients of sum(x_3) w.r.t. x0 and all weights (weight-gradient contractions included)
    from glam_b200 import functional as Fn
    blk.train()
    xg = x0.clone().requires_grad_(True)

    go = torch.ones(b.num_nodes, C, device=dev)
    params = list(blk.parameters())

    def stack_bwd(fused):
        # captured forward+backward minus captured forward (training mode: the forward also writes what backward reads)
        Fn.USE_FUSED_BWD = fused
        try:
            def fwd():
                return blk.run_steps(xg, b.edge_index, b.edge_attr, 3, keep="last", **kw)[0][0]
            t_fb = timeit(lambda: torch.autograd.grad(fwd(), [xg] + params, go), reps=5)
            t_f = timeit(fwd, reps=5)
        finally:
            Fn.USE_FUSED_BWD = True
        return t_fb - t_f
    put("MessageBlock bwd x3 (one-launch backward + weight-gradient contractions)", 3 * module_bytes("MessageBlock bwd", GRAPHS), stack_bwd(True), GRAPHS)
    put("MessageBlock bwd x3 (per-op backward kernels + weight-gradient contractions)", 3 * module_bytes("MessageBlock bwd", GRAPHS), stack_bwd(False), GRAPHS)
    blk.eval()
    with torch.no_grad():
        s2s = layer.Set2Set(C, 3).to(dev)
        put("Set2Set readout (3 steps)", 3 * module_bytes("Set2Set step", GRAPHS), timeit(lambda: s2s(x0, b.batch, num_graphs=b.num_graphs)), GRAPHS)
        gla = layer.GlobalLAPool(C).to(dev)
        put("GlobalLAPool readout", module_bytes("GlobalAttention", GRAPHS), timeit(lambda: gla(x0, b.batch, num_graphs=b.num_graphs)), GRAPHS)
        p5 = layer.GlobalPool5()
        put("GlobalPool5 readout", 4 * b.num_nodes * C + 4 * GRAPHS * 5 * C, timeit(lambda: p5(x0, b.batch, num_graphs=b.num_graphs)), GRAPHS)
        # dot-pool: DDI 25 x 25 per pair, DTI 25 x ~500
        b2 = make_molecule_batch(GRAPHS, seed=99, total_nodes=TOTAL_NODES, total_edges=TOTAL_EDGES, **DIMS).to(dev)
        xb = torch.randn(b2.num_nodes, C, device=dev)
        put("dot_and_global_pool2, DDI 25x25 pairs", 4 * (b.num_nodes + b2.num_nodes) * C + 8 * GRAPHS,
            timeit(lambda: layer.dot_and_global_pool2(x0, xb, b.batch, b2.batch, num_graphs=GRAPHS)), GRAPHS, "pairs")
        lig = make_molecule_batch(DTI_PAIRS, seed=5, total_nodes=NODES_PER * DTI_PAIRS, total_edges=EDGES_PER * DTI_PAIRS, **DIMS).to(dev)
        pro = make_protein_batch(DTI_PAIRS, seed=6).to(dev)
        xl, xpr = torch.randn(lig.num_nodes, C, device=dev), torch.randn(pro.num_nodes, C, device=dev)
        put("dot_and_global_pool2, DTI 25x~500 pairs", 4 * (lig.num_nodes + pro.num_nodes) * C + 8 * DTI_PAIRS,
            timeit(lambda: layer.dot_and_global_pool2(xl, xpr, lig.batch, pro.batch, num_graphs=DTI_PAIRS)), DTI_PAIRS, "pairs")
        gcn = layer.GCNConv(C, C).to(dev)
        put("_GCNConv protein tower layer", 4 * (2 * pro.num_nodes * C) + 8 * pro.num_edges + 4 * pro.num_nodes,
            timeit(lambda: gcn(xpr, pro.edge_index)), DTI_PAIRS, "proteins", note=f"{pro.num_nodes} residues, {pro.num_edges} edges")
    G.clear_caches()
    return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from glam_b200 import _lib, layer, model as M
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    global OVERLAP_ALLREDUCE
    OVERLAP_ALLREDUCE = bool(args.allreduce_overlap)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    _lib.load()                                      # fail loudly if the CUDA library is missing
    # the LinearBlocks around the hot path are torch.nn.Linear: same operand format as the library's projections
    torch.backends.cuda.matmul.allow_tf32 = _lib.get_math_mode() == "tf32"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    hz = Harness(dev, world)
    mse, bce, xent = torch.nn.functional.mse_loss, torch.nn.functional.binary_cross_entropy_with_logits, torch.nn.functional.cross_entropy
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    def gp_net(seed=0):
        torch.manual_seed(seed)                      # identical replicas on every rank
        return M.ArchitectureGP(DIMS["node_dim"], DIMS["edge_dim"], **NO_DROPOUT, **MODEL_KW).train()

    # clocks are sampled on rank 0 only (one nvidia-smi poller per box: eight of them starting inside the timed region took
    # a driver-wide lock and doubled the 8-GPU step time), started BEFORE the warm-up so that NVML start-up is over
    clk = ClockSampler(local) if rank == 0 else None
    if clk is not None:
        clk.__enter__()
        clk.wait_first_sample()
    results = {}
    t_head = [None, None]

    def head_timer(fn):
        t_head[0] = time.time()
        r = fn()
        t_head[1] = time.time()
        return r

    # ---- configs[1]: GLAM-GP training step (the default headline)
    net = gp_net()
    host = make_batches(N_RESIDENT, rank, pin=True)
    if args.ncu_range:
        from glam_b200.engine import TrainStep
        resident = [b.to(dev) for b in host]
        ts = TrainStep(net, mse, resident[0], lr=1e-3, device=dev, world_size=world, use_cuda_graph=not args.no_cuda_graph, warmup=3)
        for i in range(args.warmup):
            ts.load(resident[i % N_RESIDENT]); ts.run_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(args.steps):
            ts.load(resident[i % N_RESIDENT]); ts.run_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if clk is not None:
            clk.__exit__(None, None, None)
        os._exit(0)
    gp = head_timer(lambda: train_workload(hz, net, mse, host, args.steps, args.warmup, GRAPHS, capture=not args.no_cuda_graph,
                                           extra_warm=10 if world > 1 else 0)) if args.workload == "gp" else \
        train_workload(hz, net, mse, host, args.steps, args.warmup, GRAPHS, capture=not args.no_cuda_graph, extra_warm=10 if world > 1 else 0)
    results["gp"] = gp

    # ---- configs[4]: screening, on the same weights
    scr = head_timer(lambda: screening_workload(hz, net, rank)) if args.workload == "screen" else (
        screening_workload(hz, net, rank) if not args.quick or args.workload == "screen" else None)

    # ---- configs[2], [3]: the two-tower models
    ddi = dti = None
    # (at N > 1 only the headline config and screening run by default: every extra captured step holds NCCL kernels)
    if (not args.quick and world == 1) or args.workload == "ddi":
        torch.manual_seed(1)
        ddi_net = M.ArchitectureDDI(DIMS["node_dim"], DIMS["edge_dim"], end_act="ReLU", **NO_DROPOUT, **MODEL_KW).train()
        run = lambda: train_workload(hz, ddi_net, bce, make_ddi_batches(4, rank, DDI_PAIRS), max(args.steps // 2, 5), args.warmup, DDI_PAIRS,
                                     extra_warm=10 if world > 1 else 0)
        ddi = head_timer(run) if args.workload == "ddi" else run()
    if (not args.quick and world == 1) or args.workload == "dti":
        torch.manual_seed(2)
        kw = dict(MODEL_KW); kw.update(out_dim=2, mol_readout="GlobalPool5")
        dti_net = M.ArchitectureDTI(DIMS["node_dim"], 49, DIMS["edge_dim"], 8, pro_block="_GCNConv", pro_readout="GlobalPool5",
                                    end_act="ReLU", **NO_DROPOUT, **kw).train()
        run = lambda: train_workload(hz, dti_net, xent, make_dti_batches(1, rank, DTI_PAIRS), max(args.steps // 2, 5), args.warmup, DTI_PAIRS,
                                     extra_warm=10 if world > 1 else 0)
        dti = head_timer(run) if args.workload == "dti" else run()
    dti_scr = None
    if dti is not None and world == 1 and not args.quick:
        # DTI virtual screening against ONE target (LIT-PCBA shape: src_2gi_dti_scr/dataset.py:297): the reference collates one
        # protein copy per pair; `pro_index` keeps the protein once (same scores, tests/test_gpu_parity.py)
        from glam_b200.synth import make_molecule_batch, make_protein_batch
        P = 1024
        dti_net.eval()
        lig = make_molecule_batch(P, seed=31, total_nodes=NODES_PER * P, total_edges=EDGES_PER * P, **DIMS).to(dev)
        one = make_protein_batch(1, seed=32, min_len=500, max_len=500).to(dev)
        rep = make_protein_batch(P, seed=32, min_len=500, max_len=500, same_protein=True).to(dev)
        idx = torch.zeros(P, dtype=torch.int32, device=dev)

        def timed(fn, reps=5):
            with torch.no_grad():
                fn(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        f_one, f_rep = (lambda: dti_net(lig, one, pro_index=idx)), (lambda: dti_net(lig, rep))
        timed(f_one, 2); timed(f_rep, 2)                         # first calls build the per-batch indices and size the allocator
        t_one, t_rep = min(timed(f_one), timed(f_one)), min(timed(f_rep), timed(f_rep))
        dti_scr = {"pairs_per_batch": P, "protein_residues": one.num_nodes,
                   "distinct protein once (pro_index)": {"value": P / (t_one * 1e-3), "unit": "pairs/s", "ms_per_batch": t_one},
                   "one protein copy per pair (the reference's collation)": {"value": P / (t_rep * 1e-3), "unit": "pairs/s", "ms_per_batch": t_rep},
                   "note": "eval-mode forward, eager (no CUDA graph), inputs resident"}
        dti_net.train()
    if clk is not None:
        clk.__exit__(None, None, None)

    def brief(r, unit):
        return None if r is None else {"value": r["value"], "unit": unit, "ms_per_step": r["ms"],
                                       "e2e": {"value": r["e2e"], "ms_per_step": r["ms_e2e"], "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4},
                                       "gpu_launches_per_step": r["launches"], "final_loss": r["loss"]}

    also = {"GLAM-GP training (configs[1])": brief(gp, UNIT), "GLAM-DDI training (configs[2])": brief(ddi, "pairs/s"),
            "GLAM-DTI training (configs[3])": brief(dti, "pairs/s"), "screening (configs[4])": scr}
    if dti_scr is not None:
        also["GLAM-DTI screening against one target (SURVEY 8f N3)"] = dti_scr
    if ddi is not None:
        also["GLAM-DDI training (configs[2])"]["shape"] = f"{DDI_PAIRS} pairs/GPU, two {NODES_PER}-atom towers, dot-pool2 x3, Set2Set"
    if dti is not None:
        also["GLAM-DTI training (configs[3])"]["shape"] = (f"{DTI_PAIRS} pairs/GPU, ligand _TripletMessage tower + protein _GCNConv tower "
                                                          "(300-700 residues), dot-pool2 x3, GlobalPool5; one batch (not rotated: the "
                                                          "protein batches differ in size)")

    # ---- the JSON line
    head = {"gp": gp, "ddi": ddi, "dti": dti}.get(args.workload)
    if args.workload == "screen":
        line = {"metric": METRICS["screen"], "value": scr["value"], "unit": UNIT, "ms_per_step": scr["ms_per_batch"], "steps": scr["batches_per_gpu"],
                "e2e": {"value": scr["e2e"]["value"], "unit": UNIT, "h2d_bytes_per_step": scr["e2e"]["h2d_bytes_per_batch"], "d2h_bytes_per_step": 4,
                        "ms_per_step": scr["e2e"]["ms_per_batch"]},
                "gpu_launches": None}
    else:
        unit = UNIT if args.workload == "gp" else "pairs/s"
        steps = args.steps if args.workload == "gp" else max(args.steps // 2, 5)
        line = {"metric": METRICS[args.workload], "value": head["value"], "unit": unit, "ms_per_step": head["ms"], "steps": steps,
                "e2e": {"value": head["e2e"], "unit": unit, "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": 4, "ms_per_step": head["ms_e2e"],
                        "note": "TrainStep.step(host_batch, prefetch=next_host_batch): every step's H2D copy is inside the timed region, "
                                "issued on a copy stream one step ahead into the other input buffer set"},
                "gpu_launches": int((head["launches"] or 0) * steps), "gpu_launches_per_step": head["launches"], "final_loss": head["loss"]}
    line.update({"n_gpus": world, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
                 "data": "synthetic", "cuda_graph": gp["ts"].graph is not None, "allreduce_overlapped_with_backward": bool(gp["ts"].overlap_allreduce), "warmup_internal_extra": 10 if world > 1 else 0,
                 "clocks": clk.summary(t_head[0], t_head[1]) if clk is not None else None})
    cfg = workload_config(world, args.workload)
    cfg["also_measured"] = also
    line["config"] = cfg
    # key order: the contract's keys first
    order = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"]
    line = {**{k: line[k] for k in order if k in line}, **{k: v for k, v in line.items() if k not in order}}

    if rank == 0 and world == 1 and not args.quick:
        # ---- the same GP step in other modes (each its own captured step)
        variants = {}
        ts = gp["ts"]
        net_loop = gp_net(); net_loop.stack_steps = False
        r = train_workload(hz, net_loop, mse, host[:4], 10, 3, GRAPHS)
        variants["reference loop (model.py steps MessageBlock.forward 3x)"] = {"value": r["value"], "ms_per_step": r["ms"], "gpu_launches_per_step": r["launches"]}
        host_plain = make_batches(4, rank, pin=True, tile_pack=False)
        r = train_workload(hz, gp_net(), mse, host_plain, 10, 3, GRAPHS)
        variants["batches in random graph order (no tile-aware collation: ~918 instead of ~815 tiles per batch)"] = {
            "value": r["value"], "ms_per_step": r["ms"], "gpu_launches_per_step": r["launches"]}
        del host_plain
        layer.USE_FUSED_STACK = False
        try:
            r = train_workload(hz, gp_net(), mse, host[:4], 10, 3, GRAPHS)
            variants["per-op kernels (fused message kernel off)"] = {"value": r["value"], "ms_per_step": r["ms"], "gpu_launches_per_step": r["launches"]}
        finally:
            layer.USE_FUSED_STACK = True
        # the reference CLI's own defaults for the block (src_1gp/run.py:28,31-37): PairNorm, Dropout(0.2), RReLU — train-mode
        # RReLU and dropout are torch's (philox), so this configuration runs the per-op stacked / looped path, not the fused kernels
        torch.manual_seed(0)
        kw_def = dict(MODEL_KW); kw_def.update(pre_act="RReLU", graph_act="RReLU", flat_act="RReLU")
        net_def = M.ArchitectureGP(DIMS["node_dim"], DIMS["edge_dim"], graph_norm="_PairNorm", graph_do="Dropout(0.2)",
                                   flat_do="_None()", end_do="Dropout(0.2)", **kw_def).train()
        r = train_workload(hz, net_def, mse, host[:4], 10, 3, GRAPHS)
        variants["reference CLI defaults (graph_norm=_PairNorm, graph_do=Dropout(0.2), RReLU everywhere)"] = {
            "value": r["value"], "ms_per_step": r["ms"], "gpu_launches_per_step": r["launches"]}
        _lib.set_math_mode("fp32"); torch.backends.cuda.matmul.allow_tf32 = False
        try:
            r = train_workload(hz, gp_net(), mse, host[:4], 10, 3, GRAPHS)
            variants["GLAM_B200_MATH=fp32 (exact fp32 projections on the CUDA cores)"] = {"value": r["value"], "ms_per_step": r["ms"],
                                                                                         "gpu_launches_per_step": r["launches"]}
        finally:
            _lib.set_math_mode("tf32"); torch.backends.cuda.matmul.allow_tf32 = True
        cfg["also_measured"]["GLAM-GP training variants"] = variants

        # ---- roofline: the dominant library call of the training step (CUDA events around every call, eager, same inputs), keyed on
        # the ncu kernel name, + module-level rooflines on SURVEY.md §8(d) bytes
        prof = profile_kernels(ts, gp["resident"], reps=3)
        fam = {}
        for k, (ms_step, n, ms_launch) in prof.items():
            f = fam.setdefault(kernel_family(k), [0.0, 0])
            f[0] += ms_step; f[1] += n
        total_ms = sum(v[0] for v in prof.values())
        top = max(prof.items(), key=lambda kv: kv[1][0])
        kbytes = kernel_bytes_model(top[0], TOTAL_NODES, TOTAL_EDGES, GRAPHS)
        if top[0].startswith("glam_message_stack_fwd"):
            abytes = 3 * module_bytes("MessageBlock fwd", GRAPHS)
            what = "3 x MessageBlock fwd (SURVEY.md 8d: 15 584 B/graph/step x 4096 graphs)"
        elif top[0].startswith("glam_message_stack_bwd"):
            abytes = 3 * module_bytes("MessageBlock bwd", GRAPHS)
            what = "3 x MessageBlock bwd (x, h, g_x', g_h' in; g_x, g_h out; edge_attr + CSR: 22 784 B/graph/step x 4096 graphs)"
        else:
            abytes, what = kbytes, "operands of the call read/written once"
        traffic = None
        try:                                             # dram__bytes_read+write per launch from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top[0].split("[")[0] + ("[save]" if "save" in top[0] else ""))
        except (OSError, ValueError):
            pass
        achieved = abytes / (top[1][2] * 1e-3) / 1e9 if abytes else None
        line["roofline"] = {"bound": "hbm", "kernel": kernel_family(top[0]), "call": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                            "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650",
                            "ms_per_launch": top[1][2], "launches_per_step": top[1][1], "share_of_library_time": top[1][0] / total_ms,
                            "algorithmic_bytes_per_launch": abytes, "algorithmic_bytes_are": what,
                            "kernel_bytes_per_launch": kbytes,
                            "frac_on_kernel_bytes": (kbytes / (top[1][2] * 1e-3) / 1e9 / peak) if kbytes else None,
                            "kernel_bytes_are": "what this launch must move incl. the tensors it saves for backward",
                            "kernel_share_of_step": {k: round(v[0] / total_ms, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])[:8]},
                            "kernel_launches_per_step": {k: v[1] for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])[:8]},
                            "modules": module_rooflines(dev, peak)}
        line["kernel_profile_ms_per_step"] = {k: round(v[0], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]}
        line["kernel_profile_total_ms"] = round(total_ms, 4)
        if not args.no_cpu_baseline:
            gps, sec, threads = cpu_train_throughput(512, 6, 2)
            g1, s1, _ = cpu_train_throughput(128, 12, 2)
            sg, _ = cpu_screen_throughput(2048, 3, 1)
            line["cpu_baseline"] = {"value": gps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "6 training steps on 512-graph slices of the same synthetic workload "
                                              "(oracle/glam_oracle.py, all host threads)",
                                    "config1_128_graphs_fwd_bwd": {"value": g1, "unit": UNIT, "ms_per_step": s1 * 1e3,
                                                                   "sample": "BASELINE.json configs[0]: 12 training steps on the full 128-graph ESOL-shaped batch"},
                                    "screening_eval_forward": {"value": sg, "unit": UNIT, "sample": "3 eval forwards on a 2048-graph batch"}}
    if rank == 0:
        _emit(line)
    if world > 1:
        # all ranks leave together; the captured graphs hold NCCL kernels, so skip the (occasionally hanging) communicator
        # teardown: the process is exiting anyway
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


_JSON_FD = None


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gp", choices=["gp", "ddi", "dti", "screen"])
    ap.add_argument("--quick", action="store_true", help="only the headline workload (no other configs, variants, rooflines)")
    ap.add_argument("--no-cuda-graph", action="store_true")
    ap.add_argument("--allreduce-overlap", action="store_true", help="N > 1: engine.TrainStep's early gradient bucket, reduced under the stack's backward (A/B; slower)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-range", action="store_true", help="profiler range around the timed resident steps, then exit")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # exactly ONE line on stdout: native libraries (NCCL prints its version banner there) get stderr instead
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
