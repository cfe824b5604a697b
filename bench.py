#!/usr/bin/env python
"""bench.py -- agent-steps/s of the GAT-ODE hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5] [--precision f32|bf16]
                    [--impl ours|reference|reference-gpu] [--scaling strong|weak]

One "step" = one pass of the hot path over one batch of synthetic agents: the whole trajectory
(T-1 solver intervals) for every agent of the batch, i.e. B*(T-1) agent-steps.
  value : whole-job agent-steps/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e   : the same metric through the public module API with HOST (pinned) inputs, H2D and D2H in the timed region
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md §Measurement
`--impl reference` times the reference's CPU implementation of the path (the oracle port: the reference's
own PyTorch ops + the restated torchdiffeq solver + the restated GATConv) on the host cores, same config/metric;
`--impl reference-gpu` runs the SAME reference-structured modules in eager PyTorch on one B200 (the
"beat this on the same box" bar).  With N > 1 the workload's agents are SHARDED over the ranks (`--scaling strong`,
BASELINE.json configs[3]); `--scaling weak` gives every rank the full agent count instead.
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
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "agent-steps/sec, GAT-ODE"
WORKLOADS = {
    # BASELINE.json configs[1]: 10k agents x 500-zone graph, 96 RK4 steps/day, single-head GAT, inference
    "c2": dict(B=10_000, Z=500, T=97, heads=1, mode="inference", method="rk4",
               name="configs[1]: 10k agents x 500 zones, 96 RK4 steps, 1-head GAT, inference"),
    # BASELINE.json configs[2]: 1M agents x 10k zones, 4-head GAT, dopri5 adaptive (rtol = atol = 1e-5, mode_sep/config.py:27-28),
    # fwd+bwd; dense output at the 97 grid points of the day.  `--solver rk4` runs the fixed-grid variant (96 steps).
    "c3": dict(B=1_000_000, Z=10_000, T=97, heads=4, mode="train", method="dopri5",
               name="configs[2]: 1M agents x 10k zones, 4-head GAT, dopri5 rtol=atol=1e-5, fwd+bwd (agent-chunked)"),
    # BASELINE.json configs[4]: 8M agents x 10k zones, adjoint backward (the odeint_adjoint seam), bf16 tensor-core projection.
    # Agents are independent, so the adjoint runs chunk by chunk: the saved steps of one chunk (not of 8M agents) live in HBM.
    "c5": dict(B=8_000_000, Z=10_000, T=97, heads=4, mode="train", method="dopri5", adjoint=True,
               name="configs[4]: 8M agents x 10k zones, 4-head GAT, dopri5 rtol=atol=1e-5, odeint_adjoint fwd+bwd (agent-chunked)"),
}
ALG_FLOP_EVAL = 188_928         # one drift evaluation per agent
ALG_FLOP_FWD = 755_712          # rk4, per agent-step: 4 evaluations (SURVEY.md §8(d) / BASELINE.md §3)
ALG_FLOP_FWDBWD = 3_022_848     # rk4: 4x forward (discrete adjoint with stage recompute)
ALG_FLOP_FWD_DOPRI5 = 6 * ALG_FLOP_EVAL          # 1,133,568: six new evaluations per accepted step (FSAL)
ALG_FLOP_FWDBWD_DOPRI5 = 4 * ALG_FLOP_FWD_DOPRI5  # 4,534,272: + recompute + dgrad + wgrad of the same six evaluations
ALG_FLOP_FWDBWD_DOPRI5_SAVED = 3 * ALG_FLOP_FWD_DOPRI5   # 3,400,704: saved_operands = all has no recompute (forward + dgrad + wgrad)
ALG_BYTES_FWD = 1_280
ALG_BYTES_FWDBWD = 3_840
# DRAM bytes per agent-evaluation of the three stage launches (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
# capture each, 250,112 agents x 6 evaluations per launch; profiles/r02_stage_kernels_ncu_summary.txt)
NCU_DRAM_PER_UNIT = {"attempt": (4.110471e9 + 4.204920e9) / 2 / 1_500_672,      # 2,770 B (read 1.05 GB + written 3.10 GB per launch)
                     "backward": (2.191809e9 + 3.250511e9) / 1_500_672,          # 3,627 B
                     "wgrad": (4.625270e9 + 0.058811e9) / 1_500_672}             # 3,121 B


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sust=float(d["bf16_tflops_sustained"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(cfg, seed=42, device="cpu"):
    """Synthetic inputs of SURVEY.md §8(d): one generator per tensor, seeds 42+k, made on the CPU."""
    B, Z, T = cfg["B"], cfg["Z"], cfg["T"]
    g = [torch.Generator().manual_seed(seed + k) for k in range(4)]
    home = torch.randint(0, Z, (B,), generator=g[0])
    work = torch.randint(0, Z, (B,), generator=g[1])
    traits = torch.stack([torch.rand(B, generator=g[2]) * 0.72 + 0.18, torch.rand(B, generator=g[3]) * 1.4 + 0.1], dim=-1)
    t = torch.linspace(0.0, 24.0, T)
    return home, work, traits, t


def chunk_bounds(B, max_chunk, slots, tile=128):
    """[(start, end)] covering B agents in ceil(B / max_chunk) parts whose sizes are whole waves of `slots` tiles (the last part
    takes the remainder)"""
    n_parts = max(1, -(-B // max_chunk))
    tiles = -(-B // tile)
    waves = -(-tiles // slots)
    if n_parts == 1 or waves < n_parts:
        per = -(-(-(-B // n_parts)) // tile) * tile
        return [(s, min(B, s + per)) for s in range(0, B, per)]
    out, s = [], 0
    for i in range(n_parts):
        w = waves // n_parts + (1 if i < waves % n_parts else 0)
        e = B if i == n_parts - 1 else min(B, s + w * slots * tile)
        if e > s:
            out.append((s, e))
        s = e
    return out


def build_model(cfg, precision, device, saved_operands="all", adjoint_mode="discrete"):
    """GAT-ODE: zone tables from the graph-attention layers, drift net + fused solver, random init (seed 42)."""
    import ananke_abm_b200 as ab
    from ananke_abm_b200.graph import synthetic_zone_graph
    torch.manual_seed(42)
    mc = ab.ModeSepConfig()
    mc.precision = precision
    mc.ode_method = cfg["method"]            # rtol = atol = 1e-5 (mode_sep/config.py:27-28) apply to dopri5 only
    mc.adjoint = bool(cfg.get("adjoint", False))
    # c5 = the odeint_adjoint seam.  "discrete" (default of the bench): the explicit discrete-adjoint opt-in on the tensor-core stage
    # path.  "continuous": torchdiffeq's own scheme -- forward without saved steps, backward = the augmented system [y, a_y, a_theta]
    # integrated per output interval with f and its vector-Jacobian products on the fp32 kernels (O(1) memory in solver steps)
    mc.adjoint_mode = "continuous" if adjoint_mode.startswith("continuous") else adjoint_mode
    mc.adjoint_options = {"norm": "seminorm"} if adjoint_mode == "continuous-seminorm" else None
    mc.step_size = cfg.get("step_size")      # rk4 only: torchdiffeq's fixed-grid option (continuous-rk4: 0.25 h over t = [0, 24])
    mc.error_norm = "global"                 # N > 1: one RMS error norm over all ranks' agents, as a single process would use
    mc.saved_operands = saved_operands
    model = ab.GATODEModel(7, mc, heads=cfg["heads"]).to(device)
    ei, feats = synthetic_zone_graph(cfg["Z"], k=6, seed=42)
    csr = ab.build_zone_csr(ei, cfg["Z"]).to(device)
    return model, feats.to(device), csr


def labels_from_path(model, y_path, class_table, t_chunk=8):
    """argmax zone label per (agent, time) without materialising [B,T,Z] at once (host-side glue)."""
    from ananke_abm_b200.inference import head_argmax
    T = y_path.shape[0]
    E = model.config.emb_dim
    out = []
    for s in range(0, T, t_chunk):
        pred_emb = model.decoder(y_path[s:s + t_chunk, :, :E].permute(1, 0, 2))      # decoder MLP: library GEMMs
        out.append(head_argmax(pred_emb, class_table, model.config.softmax_tau))   # fused cosine head + argmax (tcgen05)
    return torch.cat(out, dim=1)


class _TrajectoryLoss(torch.autograd.Function):
    """The bench's stand-in loss  mean(y_path^2)  as ONE reduction pass forward and ONE scalar-scaled copy backward (the plain
    torch expression costs ~8 full passes over the 20 GB trajectory of a chunk; the loss is harness, not hot path).  The
    reference arms compute the same quantity with plain torch ops."""

    @staticmethod
    def forward(ctx, y_path):
        ctx.save_for_backward(y_path)
        ctx.n = y_path.numel()
        return torch.linalg.vector_norm(y_path).pow(2) / ctx.n

    @staticmethod
    def backward(ctx, go):
        (y_path,) = ctx.saved_tensors
        return y_path * (float(go) * (2.0 / ctx.n))      # host scalar: the vectorised kernel (a 0-dim CUDA operand takes the strided one)


def _config_for(args):
    cfg = dict(WORKLOADS[args.workload])
    if args.agents:
        cfg["B"] = args.agents
    if cfg.get("adjoint") and args.adjoint_mode == "continuous-rk4":
        # configs[4] as named: ONE solve over all agents, torchdiffeq's continuous adjoint with the fixed-grid rk4 solver on the
        # tensor-core stage kernels (adjoint_tc.py).  t = [0, 24] with options['step_size'] = 0.25: 96 steps per day forward, 96
        # augmented steps backward, and only y(0), y(24) exist as rows -- memory does not grow with the step count.
        cfg["method"], cfg["T"], cfg["step_size"], cfg["grid_steps"] = "rk4", 2, 0.25, 96
        cfg["name"] = ("configs[4]: 8M agents x 10k zones, 4-head GAT, rk4 step_size=0.25 (96 steps/day), odeint_adjoint = continuous "
                       "adjoint on the tensor-core stage kernels, fwd+bwd, one solve (no saved steps)")
        return cfg
    if args.solver:
        cfg["method"] = args.solver
        cfg["name"] = cfg["name"].replace("96 RK4 steps", "dopri5 rtol=atol=1e-5") if args.solver == "dopri5" else \
            cfg["name"].replace("dopri5 rtol=atol=1e-5", "96 RK4 steps")
    return cfg


def run_ours(args):
    import torch.distributed as dist
    import ananke_abm_b200 as ab
    from ananke_abm_b200 import dist as abd
    import importlib
    from ananke_abm_b200 import _lib
    oi = importlib.import_module("ananke_abm_b200.odeint")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's banner / debug lines out of the JSON stream
        dist.init_process_group("nccl", device_id=dev)
    cfg = _config_for(args)
    train = cfg["mode"] == "train"
    T = cfg["T"]
    # N > 1: "strong" = the workload's agents are sharded over the ranks in contiguous blocks (BASELINE.json configs[3]: the same
    # 1M agents on 2/4/8 GPUs; every rank sees the slice [lo, hi) of the SAME synthetic inputs a single GPU would see);
    # "weak" = every rank runs the full agent count on inputs of its own seed.
    strong = args.scaling == "strong"
    B_total = cfg["B"] if strong else cfg["B"] * world
    if strong:
        lo, hi = abd.shard_bounds(cfg["B"], rank, world)
        home, work, traits, t = make_inputs(cfg, seed=42)
        home, work, traits = home[lo:hi].clone(), work[lo:hi].clone(), traits[lo:hi].clone()
    else:
        lo, hi = 0, cfg["B"]
        home, work, traits, t = make_inputs(cfg, seed=42 + rank)
    B = hi - lo
    # agents are processed in chunks of at most --chunk agents, cut at whole WAVES of the stage kernels (#SMs x 2 slots x 128
    # agents): 1 M agents = 7,813 tiles = 26.4 waves run as 7 + 7 + 7 + 6 waves (four equal parts would each pay a partial wave);
    # the saved steps of ONE chunk live in HBM (peak_mem_gb in the JSON line)
    chunk_cap = args.chunk
    cont_adj = bool(cfg.get("adjoint", False)) and args.adjoint_mode.startswith("continuous")
    cont_rk4 = cont_adj and args.adjoint_mode == "continuous-rk4"
    if cont_rk4:
        chunk_cap = max(chunk_cap, hi - lo)      # the point of the scheme: all agents in one solve
    if train and cfg["method"] == "dopri5" and args.precision == "bf16" and not cont_adj:
        # keep one chunk's saved steps inside the memory that is actually free (measured per agent of a chunk at ~30 accepted steps:
        # 471 KB with saved_operands = all, 288 KB inputs, 112 KB none; 15 % headroom for a longer step sequence)
        per_agent = {"all": 471e3, "inputs": 288e3, "none": 112e3}[args.saved_operands] * 1.15
        free_b, _ = torch.cuda.mem_get_info(dev)
        fit = int(free_b / per_agent) // (128 * 296) * (128 * 296)
        chunk_cap = max(128 * 296, min(chunk_cap, fit))
    bounds = chunk_bounds(B, chunk_cap, 2 * torch.cuda.get_device_properties(dev).multi_processor_count)
    chunk = max(e - s for s, e in bounds)
    model, zfeat, csr = build_model(cfg, args.precision, dev, args.saved_operands, args.adjoint_mode)
    pin = lambda x: x.pin_memory()   # noqa: E731
    h_home, h_work, h_traits, h_t = pin(home), pin(work), pin(traits), pin(t)
    d_home, d_work, d_traits, d_t = (x.to(dev) for x in (home, work, traits, t))
    params = [p for p in model.parameters()]

    adaptive = cfg["method"] == "dopri5"
    counter = {"agent_steps": 0, "accepted": 0, "rejected": 0, "solves": 0}
    snap_idx = torch.arange(4, T, 8, device=dev)[:12] if (train and T > 4) else None        # 12 snap rows of the day grid
    snap_target = (torch.randint(0, cfg["Z"], (12, B), generator=torch.Generator().manual_seed(99 + rank)).to(dev)
                   if (train and args.loss == "ce") else None)

    def count(nb):
        """agent-steps of the chunk just integrated: grid intervals for rk4, ACCEPTED steps for dopri5"""
        counter["solves"] += 1
        if adaptive:
            st = oi._LAST["solver"]
            counter["agent_steps"] += nb * st.n_accepted
            counter["accepted"] += st.n_accepted
            counter["rejected"] += st.n_rejected
        else:
            counter["agent_steps"] += nb * cfg.get("grid_steps", T - 1)

    def hot_step(hm, wk, tr, tt):
        """device-resident pass over all agents of this rank; returns a small device tensor"""
        if not train:
            with torch.no_grad():
                acc = None
                table, zemb = model.zone_tables(zfeat, csr)
                for s, e in bounds:
                    y0 = model.initial_state(table, zemb, hm[s:e], wk[s:e], tr[s:e])
                    y_path = model.integrate(y0, tt)
                    count(y0.shape[0])
                    acc = y_path[-1, :1, :1]
                    del y_path
                return acc
        for p in params:
            p.grad = None
        total = None
        for s, e in bounds:
            table, zemb = model.zone_tables(zfeat, csr)
            y0 = model.initial_state(table, zemb, hm[s:e], wk[s:e], tr[s:e])
            y_path = model.integrate(y0, tt)
            count(y0.shape[0])
            if args.loss == "ce":
                # the reference's training loss at the ground-truth snaps (ce_at_snaps, losses.py:14-22): decoder + fused
                # cross-entropy head at 12 of the 97 grid points (~12 GT snaps per agent-day)
                pred_emb = model.decoder(y_path[snap_idx, :, :model.config.emb_dim])
                rows = ab.head_ce_rows(pred_emb, table, snap_target[:, s:e], model.config.softmax_tau)
                loss = rows.sum() / (snap_idx.numel() * B_total)
            else:
                loss = _TrajectoryLoss.apply(y_path) * ((e - s) / B_total)
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
            del y_path, loss
        if world > 1:
            abd.allreduce_gradients(params)             # ONE NCCL sum all-reduce of the flat gradient buffer
        return total

    def e2e_step():
        hm = h_home.to(dev, non_blocking=True); wk = h_work.to(dev, non_blocking=True)
        tr = h_traits.to(dev, non_blocking=True); tt = h_t.to(dev, non_blocking=True)
        if not train:
            with torch.no_grad():
                outs = []
                table, zemb = model.zone_tables(zfeat, csr)
                for s, e in bounds:
                    y0 = model.initial_state(table, zemb, hm[s:e], wk[s:e], tr[s:e])
                    y_path = model.integrate(y0, tt)
                    count(y0.shape[0])
                    outs.append(labels_from_path(model, y_path, table).to(torch.int32))
                    del y_path
                res = torch.cat(outs, dim=0)
        else:
            res = hot_step(hm, wk, tr, tt).reshape(1)
        return res.to("cpu", non_blocking=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        counter.update(agent_steps=0, accepted=0, rejected=0, solves=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms

    def all_ranks(counted):
        """agent-steps over ALL ranks (they can differ by a few agents per rank when B % N != 0)"""
        if world == 1:
            return counted["agent_steps"]
        v = torch.tensor([float(counted["agent_steps"])], dtype=torch.float64, device=dev)
        dist.all_reduce(v)
        return float(v.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: hot_step(d_home, d_work, d_traits, d_t), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    steps_counted = dict(counter)
    total_agent_steps = all_ranks(steps_counted)

    # kernel-only time of the dominant kernel: raw C-ABI launches into preallocated buffers, CUDA events on the
    # launching stream, no allocation or host sync between launches
    tc_train = train and args.precision == "bf16" and (not cont_adj or cont_rk4)
    fp32_accepted = None
    with torch.no_grad():
        table, zemb = model.zone_tables(zfeat, csr)
        kchunk = min(chunk, 1_000_064) if cont_rk4 else chunk      # kernel timing of the 8M-agent solve: on 1M of its agents
        y0 = model.initial_state(table, zemb, d_home[:kchunk], d_work[:kchunk], d_traits[:kchunk]).contiguous()
        spec = ab.describe_drift(model.odefunc)
        wflat = spec.flat_params().detach().contiguous()
        if adaptive and rank == 0:
            # the step count the SAME solver takes when the drift is evaluated in strict fp32 (FFMA kernels) on a slice of
            # the same agents: what an agent-step of this line is worth in solver work (accepted_steps_vs_fp32)
            nsl = min(2048, y0.shape[0])
            ab.odeint(model.odefunc, y0[:nsl].clone(), d_t, method="dopri5", rtol=model.config.rtol, atol=model.config.atol,
                      options={"precision": "f32"})
            fp32_accepted = int(oi._LAST["solver"].n_accepted)
        reps = 5
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_kernels = None
        if tc_train and adaptive and args.saved_operands == "all":
            # the three tensor-core launches of one solver step, as the training step issues them (saved_operands = all): the
            # attempt (six fused evaluations that also save the backward pass' operands), the fused backward stages (+ gather
            # entry) and the weight-gradient pass over the step's blobs
            from ananke_abm_b200 import stage as st
            Bc = y0.shape[0]
            eng = st.TcEngine(spec, wflat)
            eng.fwd_format = st.fwd_format_code("fp16x2")
            yb = st.rows_block(y0)
            y_next = st.blocked_zeros(Bc, 160, dev)
            A = [st.rows_block(torch.randn(Bc, 64, device=dev) * 0.1) for _ in range(7)]
            sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
            xb = torch.empty(eng.xblob_bytes(Bc, 2) * 6 // 4, dtype=torch.float32, device=dev)
            Gb = st.rows_block(torch.randn(Bc, 64, device=dev) * 1e-3)
            GX = [st.blocked_zeros(Bc, 160, dev) for _ in range(7)]
            Gy0 = st.blocked_zeros(Bc, 160, dev)
            lam_a = st.blocked_zeros(Bc, 64, dev)
            dtk = 0.25
            times = [1.0 + st.DOPRI5.c[i] * dtk for i in range(7)]
            eng.backward_begin(Bc, 6)

            def k_attempt():
                eng.dopri5_attempt(yb, A, 1.0, dtk, Bc, y_next, sumsq, 1e-5, 1e-5, xb, 2)

            def k_backward():
                eng.used, eng.x_ring = 0, []
                st.stages_backward(eng, st.DOPRI5, Bc, yb, A, times, dtk, [Gb] * 7, GX, 1, 6, xb, 2, y0_accum=Gy0, fsal_out=lam_a)

            def k_wgrad():
                eng.used, eng.x_ring, eng.x_level = used0, list(ring0), 2
                eng.flush()

            def time_kernel(fn):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                k0.record()
                for _ in range(reps):
                    fn()
                k1.record()
                torch.cuda.synchronize()
                return k0.elapsed_time(k1) / reps
            ms_att = time_kernel(k_attempt)
            ms_bwd = time_kernel(k_backward)
            used0, ring0 = eng.used, list(eng.x_ring)
            ms_wg = time_kernel(k_wgrad)
            eng.check_status()
            units = Bc * 6                                   # agent-evaluations per launch
            att_per_step = (counter["accepted"] + counter["rejected"]) / max(1, counter["accepted"])
            # algorithmic bytes per agent-evaluation: what the launch must read / write once, with the tile's own
            # working set (y0, a_j, later stages' gx) counted once per launch; DRAM bytes per unit from `ncu --set full`
            # (profiles/r02_stage_kernels_ncu_summary.txt, 250,112 agents)
            stage_kernels = [
                dict(kernel="stage_fwd2_tc_kernel<save> (one attempted step: 6 fused evaluations, split-fp16 activations; saves stage "
                            "inputs, hidden activations and ReLU masks for the backward pass)",
                     kernel_ms=ms_att, launches_per_accepted_step=att_per_step, alg_flop_per_unit=ALG_FLOP_EVAL,
                     alg_bytes_per_unit=149 + 256 + 107 + 1712, traffic_per_unit=NCU_DRAM_PER_UNIT["attempt"]),
                dict(kernel="stage_bwd_tc_kernel (backward of the 6 stages of a step: dgrad only, gradient blobs, gather entry)",
                     kernel_ms=ms_bwd, launches_per_accepted_step=1.0, alg_flop_per_unit=ALG_FLOP_EVAL,
                     alg_bytes_per_unit=80 + 256 + 213 + 43 + 1408 + 640 + 43, traffic_per_unit=NCU_DRAM_PER_UNIT["backward"]),
                dict(kernel="wgrad_tc_kernel (weight gradients of a step from its activation / gradient blobs)",
                     kernel_ms=ms_wg, launches_per_accepted_step=1.0, alg_flop_per_unit=ALG_FLOP_EVAL,
                     alg_bytes_per_unit=3040, traffic_per_unit=NCU_DRAM_PER_UNIT["wgrad"]),
            ]
            for kd in stage_kernels:
                kd["units_per_launch"] = units
                kd["ms_per_accepted_step"] = kd["kernel_ms"] * kd["launches_per_accepted_step"]
            dom = max(stage_kernels, key=lambda kd: kd["ms_per_accepted_step"])
            kern_ms, kern_units, kern_name = dom["kernel_ms"], units, dom["kernel"]
            kern_flop_unit, kern_bytes_unit, kern_traffic_unit = dom["alg_flop_per_unit"], dom["alg_bytes_per_unit"], dom["traffic_per_unit"]
            del eng, yb, y_next, A, xb, Gb, GX, Gy0, lam_a
        elif tc_train:
            # backward stage kernel (stage_bwd_tc_kernel) in its recomputing form: one fused launch per solver step
            from ananke_abm_b200 import stage as st
            Bc = y0.shape[0]
            eng = st.TcEngine(spec, wflat)
            yb = st.rows_block(y0)
            A = [st.rows_block(torch.randn(Bc, 64, device=dev) * 0.1) for _ in range(7)]
            Gb = st.rows_block(torch.randn(Bc, 64, device=dev) * 1e-3)
            GX = [st.blocked_zeros(Bc, 160, dev) for _ in range(7)]
            dtk = 0.25
            tab, first, last = (st.DOPRI5, 1, 6) if adaptive else (st.RK38, 0, 3)      # the fused launch of one step's backward stages
            if cont_rk4 and chunk > 5_000_000:
                first = 3      # too many agents for a four-stage blob ring: the continuous adjoint issues ONE stage per launch
            times = [1.0 + tab.c[i] * dtk for i in range(last + 1)]
            n_fused = last - first + 1
            eng.backward_begin(Bc, n_fused)

            def one_step_bwd_stages():
                eng.used, eng.x_ring = 0, []
                st.stages_backward(eng, tab, Bc, yb, A[:last], times, dtk, [Gb] * 7, GX, first, last)
            for _ in range(3):
                one_step_bwd_stages()
            torch.cuda.synchronize()
            k0.record()
            for _ in range(reps):
                one_step_bwd_stages()
            k1.record()
            torch.cuda.synchronize()
            eng.check_status()
            kern_ms = k0.elapsed_time(k1) / reps
            kern_units = Bc * n_fused            # agent-stage evaluations per launch
            kern_name = "stage_bwd_tc_kernel (%d fused Runge-Kutta stages per launch: recompute + dgrad + blob spill)" % n_fused
            kern_flop_unit = 2 * ALG_FLOP_EVAL        # recompute + dgrad of one stage (wgrad runs in wgrad_tc_kernel)
            kern_bytes_unit = ALG_BYTES_FWDBWD / 4.0
            # DRAM bytes per agent-stage of this kernel from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum =
            # 5.746 + 7.312 GB for the fused 6-stage launch over 333,440 agents; profiles/r02_stage_kernels_ncu_summary.txt)
            kern_traffic_unit = None if cont_rk4 else (5.745611e9 + 7.312489e9) / (333440 * 6)
            del eng, yb, A, Gb, GX
        else:
            prec = {"f32": 0, "bf16": 1}[args.precision]
            ybuf = torch.empty((T, y0.shape[0], y0.shape[1]), dtype=torch.float32, device=dev)
            wsb = oi.rk4_workspace(spec, y0.shape[0], T, prec, dev)
            for _ in range(3):
                oi.rk4_forward_into(spec, wflat, y0, d_t, ybuf, wsb, prec)
            torch.cuda.synchronize()
            k0.record()
            for _ in range(reps):
                oi.rk4_forward_into(spec, wflat, y0, d_t, ybuf, wsb, prec)
            k1.record()
            torch.cuda.synchronize()
            kern_ms = k0.elapsed_time(k1) / reps
            kern_units = y0.shape[0] * (T - 1)   # agent-steps per launch
            kern_name = "rk4 fused trajectory (forward), %s" % ("rk4_tc_kernel" if prec == 1 else "rk4_f32_kernel")
            kern_flop_unit = ALG_FLOP_FWD
            kern_bytes_unit = ALG_BYTES_FWD
            kern_traffic_unit = None
            del ybuf
    _lib.LAUNCHES = 0
    hot_step(d_home, d_work, d_traits, d_t)
    torch.cuda.synchronize()
    launches_per_step = _lib.LAUNCHES
    e2e_steps = max(1, min(args.steps, 3))
    e2e_ms = timed(e2e_step, e2e_steps, max(1, min(args.warmup, 3)))     # same warm-up rule as the device-resident leg
    e2e_counted = dict(counter)
    e2e_agent_steps = all_ranks(e2e_counted)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = total_agent_steps / (ms * 1e-3)
    e2e_value = e2e_agent_steps / (e2e_ms * 1e-3)
    flops_kernel = kern_units * kern_flop_unit
    achieved_tf = flops_kernel / (kern_ms * 1e-3) / 1e12
    bytes_kernel = kern_units * kern_bytes_unit
    # which roof bounds the dominant kernel: its algorithmic intensity against the ridge of the measured peaks.  With the backward
    # pass' operands saved by the forward launch (saved_operands = all) every stage launch sits on the HBM side of the ridge.
    hbm_gbs = bytes_kernel / (kern_ms * 1e-3) / 1e9
    if stage_kernels and kern_flop_unit / kern_bytes_unit < pk["tf_burst"] * 1e12 / (pk["hbm"] * 1e9):
        roof_head = {"bound": "hbm", "achieved": hbm_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": hbm_gbs / pk["hbm"]}
    else:
        roof_head = {"bound": "tensor", "achieved": achieved_tf, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tf_burst"]}
    if stage_kernels:
        for kd in stage_kernels:
            tsec = kd["kernel_ms"] * 1e-3
            kd["alg_gbs"] = kd["units_per_launch"] * kd["alg_bytes_per_unit"] / tsec / 1e9
            kd["alg_tflops"] = kd["units_per_launch"] * kd["alg_flop_per_unit"] / tsec / 1e12
            kd["frac_hbm"] = kd["alg_gbs"] / pk["hbm"]
            kd["frac_tensor"] = kd["alg_tflops"] / pk["tf_burst"]
            kd["dram_gbs"] = (kd["units_per_launch"] * kd["traffic_per_unit"] / tsec / 1e9) if kd["traffic_per_unit"] else None
    h2d = sum(x.numel() * x.element_size() for x in (h_home, h_work, h_traits, h_t))
    d2h = (B * T * 4) if not train else 4
    acc_per = steps_counted["accepted"] / max(1, steps_counted["solves"])
    rej_per = steps_counted["rejected"] / max(1, steps_counted["solves"])
    if adaptive:
        fl_fwd, fl_fb = ALG_FLOP_FWD_DOPRI5, (ALG_FLOP_FWDBWD_DOPRI5_SAVED if stage_kernels else ALG_FLOP_FWDBWD_DOPRI5)
    else:
        fl_fwd, fl_fb = ALG_FLOP_FWD, ALG_FLOP_FWDBWD
        if cont_rk4:      # per grid step: 4 evaluations forward; backward 4 x (y stage + recompute inside the VJP launch + dgrad + wgrad)
            fl_fb = 20 * ALG_FLOP_EVAL
    out = {
        "metric": METRIC + (" fwd+bwd" if train else " fwd (inference)"), "value": value, "unit": "agent-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": ("strong" if strong else "weak"), "vs_baseline": None,
        "dtype": "f32" if args.precision == "f32" else "fp16 operands in the forward stages / bf16 operands in the vector-Jacobian products of the augmented adjoint system (tcgen05, fp32 accumulate; fp32 y, a_y, a_theta)" if cont_rk4 else "fp16 weights x split-fp16 (hi+lo) activations in the forward solve / fp32 FFMA kernels for the augmented adjoint system" if cont_adj else ("fp16 weights x split-fp16 (hi+lo) activations fwd / bf16 bwd operands (fp32 accumulate, fp32 state)" if (train and adaptive) else ("fp16 fwd / bf16 bwd operands (fp32 accumulate, fp32 state)" if train else "fp16 operands on tcgen05 (fp32 accumulate, fp32 state)")), "data": "synthetic",
        # equal-work figure: simulated agent-days (whole trajectories, forward + backward) per second, independent of how
        # many solver steps the adaptive controller needed
        "agent_days_per_s": B_total * args.steps / (ms * 1e-3),
        "config": {"workload": cfg["name"], "agents_total": B_total, "agents_per_gpu": B, "zones": cfg["Z"], "time_points": T,
                   "solver": cfg["method"], "agent_chunk": chunk, "agent_chunks": [e - s for s, e in bounds], "saved_operands": (args.saved_operands if (train and cfg["method"] == "dopri5" and args.precision == "bf16") else None),
                   "precision": args.precision, "loss": (args.loss if train else None),
                   "adjoint_mode": (args.adjoint_mode if cfg.get("adjoint") else None),
                   "solver_steps": ({"accepted_per_trajectory": acc_per, "rejected_per_trajectory": rej_per,
                                     "fp32_accepted_per_trajectory": fp32_accepted,
                                     "accepted_steps_vs_fp32": (acc_per / fp32_accepted if fp32_accepted else None),
                                     "fp32_note": "same solver, drift in strict fp32 (FFMA kernels), first %d agents of rank 0" % min(2048, chunk),
                                     "rtol": model.config.rtol, "atol": model.config.atol} if adaptive else {"grid_intervals": cfg.get("grid_steps", T - 1), "step_size": cfg.get("step_size")}),
                   "l2": "trajectory rows written per step (%.0f MB) exceed L2; weights are L2-resident by design" % (chunk * T * 640 / 1e6)},
        "roofline": {**roof_head,
                     "traffic": (kern_traffic_unit * kern_units if kern_traffic_unit else None),
                     "traffic_note": "DRAM bytes per launch scaled from one ncu --set full capture (profiles/r02_stage_kernels_ncu_summary.txt)"
                     if kern_traffic_unit else None,
                     "dram_achieved_gbs": (kern_traffic_unit * kern_units / (kern_ms * 1e-3) / 1e9 if kern_traffic_unit else None),
                     "dram_frac_of_hbm_peak": (kern_traffic_unit * kern_units / (kern_ms * 1e-3) / 1e9 / pk["hbm"] if kern_traffic_unit else None),
                     "peak_source": pk["src"],
                     "kernel": kern_name, "kernel_ms": kern_ms, "units_per_launch": kern_units,
                     "alg_flop_per_unit": kern_flop_unit, "alg_bytes_per_unit": kern_bytes_unit,
                     "alg_flop_per_agent_step": fl_fb if train else fl_fwd,
                     "alg_bytes_per_agent_step": ALG_BYTES_FWDBWD if train else ALG_BYTES_FWD,
                     "step_tflops_alg": value * (fl_fb if train else fl_fwd) / 1e12,
                     "hbm_achieved_gbs": bytes_kernel / (kern_ms * 1e-3) / 1e9, "hbm_peak_gbs": pk["hbm"],
                     "tensor_achieved_tflops": achieved_tf, "tensor_peak_tflops": pk["tf_burst"],
                     "arithmetic_intensity_flop_per_byte": kern_flop_unit / kern_bytes_unit,
                     "ridge_flop_per_byte": pk["tf_burst"] * 1e12 / (pk["hbm"] * 1e9),
                     "stage_kernels": stage_kernels,
                     # the three stage launches together: algorithmic-byte fraction of the HBM roof weighted by their time per accepted step
                     "stage_kernels_time_weighted_frac_hbm": (sum(kd["frac_hbm"] * kd["ms_per_accepted_step"] for kd in stage_kernels)
                                                              / sum(kd["ms_per_accepted_step"] for kd in stage_kernels)) if stage_kernels else None},
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / e2e_steps, "agent_days_per_s": B_total * e2e_steps / (e2e_ms * 1e-3)},
        "gpu_launches": args.steps * launches_per_step,
        "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1 and not cont_rk4:      # the CPU leg is timed on rank 0 at N = 1 only (torchrun pins OMP threads)
        out["cpu_baseline"] = reference_step_timer(cfg, train, "cpu", steps=1, warmup=0, budget_agents=args.ref_agents)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def reference_step_timer(cfg, train, device, steps=1, warmup=0, budget_agents=None, wall_budget_s=240.0):
    """The reference-structured GAT-ODE (oracle port: the reference's PyTorch ops for the drift / encoder / decoder, the
    restated torchdiffeq solver, the restated GATConv) on `device` ("cpu": all host cores; "cuda": eager PyTorch on one
    B200), on a bounded slice of the same workload: the first `budget_agents` agents, the full day, the full zone graph.
    One step = GAT zone tables + initial state + solve (+ loss backward when training).  `warmup` untimed steps, then up
    to `steps` timed ones (fewer only if the wall budget would be exceeded -- the returned `steps_run` is the real count);
    `seconds` is the mean timed step."""
    from oracle import models_oracle as mo
    from oracle import torchdiffeq_oracle as tdq
    from oracle import gat_oracle as go
    cores = os.cpu_count() or 1
    on_gpu = device != "cpu"
    if not on_gpu:
        torch.set_num_threads(cores)
    dev = torch.device(device)
    adaptive = cfg["method"] == "dopri5"
    Bs = min(cfg["B"], budget_agents or (32_768 if on_gpu else 10_000))
    Ts = cfg["T"]
    sub = dict(cfg, B=Bs)
    home, work, traits, t = make_inputs(sub)
    torch.manual_seed(42)
    m = mo.OracleGATODE(7, cfg["heads"]).to(dev)
    ei, feats = go.synthetic_zone_graph(cfg["Z"], k=6, seed=42)
    edges = go.symmetrise_with_self_loops(ei, cfg["Z"]).to(dev)
    feats, home, work, traits, t = (x.to(dev) for x in (feats, home, work, traits, t))
    kw = dict(method="dopri5", rtol=1e-5, atol=1e-5) if adaptive else dict(method="rk4")
    n_steps = Ts - 1

    def one():
        nonlocal n_steps
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if not train:
            with torch.no_grad():
                table, zemb = m.zone_tables(feats, edges)
                y0 = m.initial_state(table, zemb, home, work, traits)
                tdq.odeint(m.rhs, y0, t, **kw)
        else:
            m.zero_grad()
            table, zemb = m.zone_tables(feats, edges)
            y0 = m.initial_state(table, zemb, home, work, traits)
            yp = tdq.odeint(m.rhs, y0, t, **kw)
            (yp ** 2).mean().backward()
            del yp
        if on_gpu:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if adaptive:
            n_steps = sum(1 for (_, _, ok) in tdq._LAST_SOLVER["solver"].step_log if ok)
        return dt

    wall0 = time.perf_counter()
    for _ in range(warmup):
        one()
    times = []
    for it in range(steps):
        times.append(one())
        spent = time.perf_counter() - wall0
        if it + 1 < steps and spent + times[-1] > wall_budget_s:
            break
    mean = sum(times) / len(times)
    what = f"{n_steps} accepted dopri5 steps (rtol=atol=1e-5)" if adaptive else f"{Ts - 1} rk4 steps"
    where = (f"torch eager fp32 on {torch.cuda.get_device_name(dev)}" if on_gpu else f"torch CPU fp32, {cores} threads")
    return {"value": Bs * n_steps / mean, "unit": "agent-steps/s", "cores": (0 if on_gpu else cores), "kind": "port",
            "device": device, "agent_days_per_s": Bs / mean,
            "sample": f"{Bs} agents x {what} of the same workload ({'fwd+bwd' if train else 'fwd'}, {cfg['heads']}-head GAT over "
                      f"{cfg['Z']} zones included), {where}, mean of {len(times)} timed steps after {warmup} warm-up",
            "seconds": mean, "steps_run": len(times), "accepted_per_trajectory": (n_steps if adaptive else None)}


def run_reference(args, device="cpu"):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if device == "cpu" and "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("AB200_REF_CHILD") != "1":
        # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm must see all host cores as it does at N=1:
        # re-run this arm in a clean child process and relay its line
        env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
        env["AB200_REF_CHILD"] = "1"
        out = subprocess.run([sys.executable, str(REPO / "bench.py")] + sys.argv[1:], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if line:
            d = json.loads(line[-1])
            d["n_gpus"] = int(os.environ.get("WORLD_SIZE", "1"))
            print(json.dumps(d))
            return
    cfg = _config_for(args)
    train = cfg["mode"] == "train"
    t0 = time.perf_counter()
    if device != "cpu":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    cb = reference_step_timer(cfg, train, device, steps=max(1, args.steps), warmup=max(0, args.warmup), budget_agents=args.ref_agents)
    out = {"impl": "reference" if device == "cpu" else "reference-gpu",
           "metric": METRIC + (" fwd+bwd" if train else " fwd (inference)"), "value": cb["value"],
           "unit": "agent-steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": cb["steps_run"], "steps_requested": args.steps,
           "warmup": args.warmup, "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "agent_days_per_s": cb["agent_days_per_s"],
           "config": {"workload": cfg["name"], "sample": cb["sample"],
                      "same_config": "same model, zone graph, solver, tolerances and inputs as the GPU arm; a bounded slice of the agents "
                                     "(agents are independent: per-agent cost does not depend on the slice beyond CPU cache effects)"},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: 'strong' shards the workload's agents over the ranks (configs[3]); 'weak' gives every rank all of them")
    ap.add_argument("--ref-agents", type=int, default=0,
                    help="agents in the reference arms' / cpu_baseline's slice (default 10,000 on the CPU, 32,768 on the GPU)")
    ap.add_argument("--agents", type=int, default=0, help="override agents per GPU")
    ap.add_argument("--chunk", type=int, default=265_216,
                    help="upper bound on agents per launch sequence (the batch is cut into equal parts no larger than this); "
                         "265,216 = 7 x (148 SMs x 2 slots x 128 agents); dopri5 training with --saved-operands all keeps ~0.44 MB "
                         "per agent of the chunk (30 accepted steps x 10.3 KB + the 97-row trajectory and its gradient)")
    ap.add_argument("--saved-operands", default="all", choices=["all", "inputs", "none"],
                    help="dopri5 training: what an accepted attempt's forward launch keeps for the backward pass as operand images "
                         "(all: stage inputs + hidden activations + ReLU masks, the backward kernel recomputes nothing; inputs: "
                         "stage inputs only; none: the backward pass rebuilds everything from (y, a_j))")
    ap.add_argument("--adjoint-mode", default="discrete", choices=["discrete", "continuous", "continuous-seminorm", "continuous-rk4"],
                    help="--workload c5 (the odeint_adjoint seam): discrete adjoint of the accepted steps on the tensor-core stage path, or "
                         "torchdiffeq's continuous adjoint (no saved steps; augmented system on the fp32 kernels)")
    ap.add_argument("--solver", default="", choices=["", "rk4", "dopri5"], help="override the workload's solver")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--loss", default="traj", choices=["traj", "ce"],
                    help="training loss in the timed step: 'traj' = mean square of the trajectory (stand-in, default); "
                         "'ce' = decoder + fused cross-entropy head at 12 snap points per agent (tensor cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, "cpu")
    elif args.impl == "reference-gpu":
        run_reference(args, "cuda")
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
