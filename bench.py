#!/usr/bin/env python
"""bags/sec of one full adversarial G+D train step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|tf32|fp32] [--impl reference]

Workload (BASELINE.json configs[1]): AdvMIL-ABMIL, synthetic bags of 16,384 x 1024 fp32 features, 16 bags per optimiser
step and per GPU (bp_every_batch, config/cfg_nlst.yaml:71) = 1 GiB of features per step per GPU (>> 126 MB L2, so no
L2 flush is needed between iterations).  A "step" = _update_disc + _update_gen (model/model_handler.py:349-498): both
forward/backward passes and both Adam updates.  Weak scaling: every rank owns 16 bags; gradients all-reduced (NCCL).

Printed JSON line: value = device-timed bags/s with inputs resident in HBM; e2e = the same through the public API
(DeviceFeeder + AdvStep) including the pinned-host -> device copy of every step's bags and a device -> host read of
the losses; roofline = dominant kernel class from live CUDA events; cpu_baseline = the oracle port of the reference's
CPU path on this box's host cores (bounded sample).  `--impl reference` times only that CPU path.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "bags/sec per G+D train step (16k x 1024 bags)"
N_ROWS, C_IN, BAGS_PER_STEP = 16384, 1024, 16
FLOP_PER_ROW = {  # algorithmic FLOPs per instance row per launch
    "proj_fwd": 2 * 1024 * 384, "gate_fwd": 2 * 384 * 768 + 768, "embed_fwd": 2 * 1024 * 128,
    "proj_embed_fwd": 2 * 1024 * (384 + 128),      # K1 + K5 in one pass over x (stacked weights)
    "bwd_data": 2 * 768 * 384, "bwd_w_gate": 2 * 768 * 384, "bwd_w_proj": 2 * 384 * 1024, "bwd_w_embed": 2 * 128 * 1024,
}
def bytes_per_row(es):
    """Algorithmic HBM bytes per instance row per launch for the streaming kernels; es = bytes per activation element
    (4 in the fp32/tf32 modes, 2 in the bf16 mode).  pool: read h + s, write w; pool_gate_bwd: read h, ab, write dAB;
    ln_bwd: read y_pre, write dy (+ d_emb/16); colsum: read dh; dropout: read h_eval, write h."""
    return {"pool_fwd": 384 * es + 8, "pool_gate_bwd": 384 * es + 2 * 768 * es + 8, "ln_bwd": 2 * 128 * es + 32,
            "colsum": 384 * es, "dropout": 2 * 384 * es}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "src": "fallback"}


class NvmlSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML every 10 ms while the timed region runs (the recipe's clocks line,
    B200_PROFILING.md; nvidia-smi's own start-up is longer than a short timed region)."""

    def __init__(self, gpu_index=0, period=0.01):
        super().__init__(daemon=True)
        self.gpu, self.period, self.rows, self._stop_evt, self.ok = gpu_index, period, [], threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if gpu_index < len(ids) and ids[gpu_index].strip().isdigit():
                    idx = int(ids[gpu_index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((mhz, reasons, time.perf_counter()))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self, t0=None, t1=None):
        """t0, t1: perf_counter bounds of the timed region; only samples taken inside it are reported."""
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[2] <= t1]
            self.rows = inside if inside else self.rows[-3:]
        if not self.ok or not self.rows:
            return None
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}
        seen = set()
        for _, r, _t in self.rows:
            for name, b in bits.items():
                if r & b:
                    seen.add(name)
        return {"sm_mhz": float(np.median([m for m, _, _t in self.rows])), "sm_max_mhz": self.max_mhz, "reasons": sorted(seen),
                "samples": len(self.rows), "source": "nvml"}


class ClockSampler(threading.Thread):
    """Fallback: samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's PyTorch path; the reference is pure Python and cannot travel)
# ------------------------------------------------------------------------------------------------------
def cpu_reference_step_time(n_rows, bags_per_sample, steps, warmup, threads=None, backbone="abmil"):
    from oracle import advmil_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sdG = O.synth_state_dict(O.G_ESAT_SHAPES() if backbone == "patch" else O.G_SHAPES(), 42)
    sdD = O.synth_state_dict(O.D_SHAPES(), 43)
    tr = O.CpuTrainer(sdG, sdD, backbone=backbone)
    gen = torch.Generator().manual_seed(42)
    bags = [torch.randn(n_rows, C_IN, generator=gen) for _ in range(bags_per_sample)]
    ts, es = O.synth_labels(bags_per_sample, 42)
    es[0] = 1.0
    vis = [True] * bags_per_sample
    times = []
    for it in range(warmup + steps):
        nzd = [torch.rand(1, 192) for _ in bags]
        nzg = [torch.rand(1, 192) for _ in bags]
        gm = None if backbone == "patch" else [O.random_g_masks(n_rows, 384, 384, gen) for _ in bags]   # ESAT: dropout-free sample
        dmr = [O.random_d_masks(n_rows // 16, 128, gen) for _ in bags]
        dmf = [O.random_d_masks(n_rows // 16, 128, gen) for _ in bags]
        t0 = time.perf_counter()
        tr.step(bags, ts, es, vis, nzd, nzg, dmr, dmf, gm)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times)), float(np.min(times)), threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    mean_s, best_s, threads = cpu_reference_step_time(N_ROWS, 1, args.steps, args.warmup)
    val = 1.0 / mean_s
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "bags/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "AdvMIL-ABMIL G+D step, synthetic bags 16384x1024 fp32 (configs[1])",
                   "sample": "1 bag per step (the reference loops over bags one at a time, so bags/s is per-bag time)"},
        "cpu_baseline": {"value": val, "unit": "bags/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x 1 bag of 16384x1024 (D step + G step, Adam), torch CPU fp32"},
        "e2e": {"value": val, "unit": "bags/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_esat(args, rank, world, local_rank):
    """`--backbone patch`: the same metric for the ESAT generator (DualTrans_HS, bcb_mode patch; the reference's default
    backbone) with the RLIP discriminator, through step.ModuleAdvStep.  Same shape, timing rules and JSON keys as the ABMIL
    line; the roofline object is the self-attention forward kernel against the tf32 tensor peak (half the measured bf16
    figure: MEASURED_PEAKS.json holds no tf32 number)."""
    import contextlib
    from types import SimpleNamespace as NS

    import advmil_b200
    from advmil_b200 import _lib, ops
    from advmil_b200.dataset.packed import DeviceFeeder, synthetic_steps
    from advmil_b200.model.backbone import load_backbone
    from advmil_b200.model.GANSurv import Generator, PrjDiscriminator
    from advmil_b200.step import ModuleAdvStep

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    torch.manual_seed(42)
    with contextlib.redirect_stdout(sys.stderr):
        G = Generator(384, 1, load_backbone("patch", [C_IN, 384, 384]), NS(noise=[0, 1], hops=1, noise_dist="uniform"), False, 0.6,
                      "sigmoid").to(dev)
        D = PrjDiscriminator(NS(in_dim=C_IN, out_dim=128, ksize=1, backbone="avgpool", dropout=0.25),
                             NS(in_dim=1, hid_dims=[64, 128], norm=False, dropout=0.0), prj_path="x", inner_product="instance").to(dev)
    eng = ModuleAdvStep(G, D, precision=args.precision)
    feat_dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32
    steps = synthetic_steps(args.warmup + args.steps, args.bags, args.rows, C_IN, seed=42 + rank, distinct=2, dtype=feat_dtype)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    resident = [(ops.PackedBags(st.x.to(dev), st.lengths), st.t.to(dev), st.e.to(dev), st.visible.to(dev)) for st in steps[:2]]
    nz = torch.rand(args.bags, 192, device=dev)
    counts = [(float(((s.e == 1) & (s.visible != 0)).sum()) * world, float(args.bags * world), float(s.visible.sum()) * world) for s in steps[:2]]

    def one_step(i):
        b, t, e, v = resident[i % 2]
        return eng.step(b, t, e, v, noise_d=nz, noise_g=nz, global_counts=counts[i % 2])

    sampler = NvmlSampler(local_rank)
    if not sampler.ok:
        sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        one_step(i)
    barrier()
    lib.advmil_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    th0 = time.perf_counter()
    for i in range(args.steps):
        out = one_step(args.warmup + i)
    ev1.record()
    barrier()
    clocks = sampler.stop(th0, time.perf_counter()) if rank == 0 else None
    launches = int(lib.advmil_launch_count(0))
    tms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
    ms_dev = float(tms.item())
    value = args.bags * world * args.steps / (ms_dev / 1e3)
    # per-kernel-class events in a separate pass
    ntags = len(_lib.PROF_TAGS)
    pms, pcnt = (C.c_double * ntags)(), (C.c_int64 * ntags)()
    prof_steps = min(args.steps, 10)
    lib.advmil_profile_enable(1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(prof_steps):
        one_step(i)
    pe1.record()
    torch.cuda.synchronize()
    lib.advmil_profile_enable(0)
    lib.advmil_profile_read(pms, pcnt, ntags)
    prof_ms = pe0.elapsed_time(pe1)
    R = args.rows // 16
    pair_flop = 2.0 * 48 * 8 * args.bags * R * R                    # one [R, R] x 48 contraction over 8 heads and all bags
    flop = {"attn_fwd": 2 * pair_flop, "attn_bwd": 5 * pair_flop}   # algorithmic: S, PV / S, dP, dV, dK, dQ
    kern = {}
    for i, tag in enumerate(_lib.PROF_TAGS):
        if pcnt[i]:
            avg = pms[i] / pcnt[i]
            kern[tag] = {"ms_per_launch": avg, "launches_per_step": pcnt[i] / prof_steps, "share_of_step": pms[i] / prof_ms}
            if tag in flop:
                kern[tag]["tflops"] = flop[tag] / (avg * 1e-3) / 1e12
    peaks = load_peaks()
    roof = None
    if "attn_fwd" in kern and args.precision != "fp32":
        roof = {"kernel": "attn_fwd", "bound": "tensor", "achieved": kern["attn_fwd"]["tflops"], "peak": peaks["tflops"] / 2, "unit": "TFLOP/s",
                "frac": kern["attn_fwd"]["tflops"] / (peaks["tflops"] / 2), "traffic": None,
                "peak_source": peaks["src"] + " bf16 sustained / 2 (tf32 operands; warp-level mma.sync kernels)"}
    # end to end: pinned host -> device every step (12-bit transport in the bf16 mode)
    p12 = args.precision == "bf16" and args.transport == "p12"
    if p12:
        for st in steps[:2]:
            st.pack12()
    it = iter(DeviceFeeder(steps, device=dev, depth=2))
    h2d = steps[0].nbytes
    for i in range(args.warmup):
        s = next(it)
        float(eng.step(s.bags, s.t, s.e, s.visible, global_counts=s.counts if world == 1 else None)["dis_loss"])
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        s = next(it)
        o = eng.step(s.bags, s.t, s.e, s.visible, global_counts=s.counts if world == 1 else None)
        host = [float(o[k]) for k in ("dis_loss", "gen_loss", "t_reg_loss", "gen_total_loss")]     # device -> host read of the result
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ems = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ems, op=torch.distributed.ReduceOp.MAX)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mean_s, best_s, threads = cpu_reference_step_time(args.rows, 1, 3, 1, backbone="patch")
        cpu = {"value": 1.0 / mean_s, "unit": "bags/s", "cores": threads, "kind": "port",
               "sample": f"3 steps x 1 bag of {args.rows}x1024 after 1 warm-up (D step + G step + Adam), oracle port of the ESAT path, torch CPU fp32"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "bags/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
                "config": {"workload": f"AdvMIL-ESAT (bcb_mode patch: DualTrans_HS generator, RLIP discriminator) G+D step, {args.bags} "
                                       f"synthetic bags of {args.rows}x1024 per step per GPU; D step + G step + both Adam updates",
                           "rows_per_step_per_gpu": args.bags * args.rows, "precision_mode": args.precision,
                           "l2": "inputs (two alternating steps of 512 MiB or more) larger than the 126 MB L2; no flush",
                           "parallelism": f"dp{world} (bags sharded, NCCL all-reduce of flat G/D grads)"},
                "e2e": {"value": args.bags * world * args.steps / (float(ems.item()) / 1e3), "unit": "bags/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 16, "transport": "p12" if p12 else "raw"},
                "gpu_launches": launches, "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "clocks": clocks,
                "losses_last_step": dict(zip(("dis_loss", "gen_loss", "t_reg_loss", "gen_total_loss"), host))}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="advmil_b200")
    ap.add_argument("--precision", default=os.environ.get("ADVMIL_PRECISION", "bf16"), choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--bags", type=int, default=BAGS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="p12", choices=["p12", "raw"],
                    help="end-to-end leg, bf16 mode: copy the features in the packed loader's lossless 12-bit transport format "
                         "(decoded on the device) or as raw bf16")
    ap.add_argument("--backbone", default="abmil", choices=["abmil", "patch"],
                    help="generator backbone: abmil (the benchmark's workload, C-fused AdvStep) or patch (ESAT, ModuleAdvStep)")
    ap.add_argument("--ragged", action="store_true",
                    help="configs[2] instead of configs[1]: bag lengths log-uniform in [1024, 100000] (multiples of 16)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert args.warmup >= 3, "timing rule: at least 3 warm-up steps"
    if args.backbone == "patch":
        run_esat(args, rank, world, local_rank)
        return

    import advmil_b200
    from advmil_b200 import _lib, ops
    from advmil_b200.dataset.packed import DeviceFeeder, synthetic_steps
    from advmil_b200.step import AdvStep
    from tests.util import build_D, build_G

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    advmil_b200.set_precision(args.precision)
    torch.manual_seed(42)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):              # the modules print the reference's [info] lines
        G, D = build_G(device=dev), build_D(device=dev)
    from advmil_b200.model.model_utils import init_weights
    G.apply(init_weights)                                     # model_handler.py:81
    engine = AdvStep(G, D, precision=args.precision)

    # ---- synthetic bags in pinned host memory: 2 distinct steps (2 GiB) cycled ----
    n_total = args.warmup + args.steps
    feat_dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32     # the packed loader's storage format
    from advmil_b200.dataset.packed import loguniform_rows
    steps = synthetic_steps(n_total, args.bags, loguniform_rows() if args.ragged else args.rows, C_IN, seed=42 + rank,
                            distinct=2, dtype=feat_dtype)
    counts = None  # global pair counts come from a tiny all-reduce inside step()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ================= (A) device-resident timing: `value` =================
    resident = []
    for st in steps[:2]:
        resident.append((ops.PackedBags(st.x.to(dev), st.lengths), st.t.to(dev), st.e.to(dev), st.visible.to(dev)))
    nz = [torch.rand(args.bags, 192, device=dev) for _ in range(2)]
    hostcounts = [(float(((s.e == 1) & (s.visible != 0)).sum()) * world, float(args.bags * world), float(s.visible.sum()) * world)
                  for s in steps[:2]]

    def one_step(i):
        b, t, e, v = resident[i % 2]
        return engine.step(b, t, e, v, noise_d=nz[0], noise_g=nz[1], global_counts=hostcounts[i % 2])

    sampler = NvmlSampler(local_rank)
    if not sampler.ok:
        sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()             # started before the warm-up: nothing but the timed loop sits between the barriers
    for i in range(args.warmup):
        one_step(i)
    barrier()
    lib.advmil_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    th0 = time.perf_counter()
    for i in range(args.steps):
        out = one_step(args.warmup + i)
    host_ms = (time.perf_counter() - th0) * 1e3 / args.steps      # CPU time to ISSUE one step (no sync inside the loop)
    ev1.record()
    barrier()
    clocks = sampler.stop(th0, time.perf_counter()) if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = int(lib.advmil_launch_count(0))
    # per-kernel-class CUDA-event timing in a SEPARATE pass (the event records perturb the step, so they stay out of the
    # timed region above); shares are relative to this pass's own step time
    ntags = len(_lib.PROF_TAGS)
    pms, pcnt = (C.c_double * ntags)(), (C.c_int64 * ntags)()
    prof_steps = min(args.steps, 10)
    lib.advmil_profile_enable(1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(prof_steps):
        one_step(args.warmup + args.steps + i)
    pe1.record()
    torch.cuda.synchronize()
    lib.advmil_profile_enable(0)
    lib.advmil_profile_read(pms, pcnt, ntags)
    prof_ms = pe0.elapsed_time(pe1)
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
    ms_dev = float(tms.item())
    value = args.bags * world * args.steps / (ms_dev / 1e3)
    losses = engine.loss_dict(out)

    # ================= (B) end-to-end through the public API: pinned host -> device every step =================
    p12 = args.precision == "bf16" and args.transport == "p12"
    if p12:
        for st in steps[:2]:
            st.pack12()           # done once, at packing time (the list cycles the same two steps)
    feeder = DeviceFeeder(steps, device=dev, depth=2)
    it = iter(feeder)
    h2d = steps[0].nbytes
    for i in range(args.warmup):
        s = next(it)
        o = engine.step(s.bags, s.t, s.e, s.visible, global_counts=s.counts if world == 1 else None)
        engine.loss_dict(o)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2h = 0
    for i in range(args.steps):
        s = next(it)
        o = engine.step(s.bags, s.t, s.e, s.visible, global_counts=s.counts if world == 1 else None)
        host = o["losses"].tolist()      # device -> host read of the step's result
        d2h = len(host) * 4
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ems = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ems, op=torch.distributed.ReduceOp.MAX)
    e2e_value = args.bags * world * args.steps / (float(ems.item()) / 1e3)

    # ================= roofline of the dominant kernel class =================
    peaks = load_peaks()
    rows_per_launch = int(np.mean([sum(st.lengths) for st in steps[:2]]))
    BYTES_PER_ROW = bytes_per_row(2 if args.precision == "bf16" else 4)
    kern = {}
    for i, tag in enumerate(_lib.PROF_TAGS):
        if pcnt[i] == 0:
            continue
        avg_ms = pms[i] / pcnt[i]
        ent = {"ms_per_launch": avg_ms, "launches_per_step": pcnt[i] / prof_steps, "share_of_step": pms[i] / prof_ms}
        if tag in FLOP_PER_ROW:
            ent["tflops"] = FLOP_PER_ROW[tag] * rows_per_launch / (avg_ms * 1e-3) / 1e12
        if tag in BYTES_PER_ROW:
            ent["gbs"] = BYTES_PER_ROW[tag] * rows_per_launch / (avg_ms * 1e-3) / 1e9
        kern[tag] = ent
    rated = [k for k in kern if k in FLOP_PER_ROW or k in BYTES_PER_ROW]      # composites (latency-bound) have no roofline
    top = max(rated, key=lambda k: kern[k]["share_of_step"]) if rated else None
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if top is not None and os.path.exists(tpath):
        tj = json.load(open(tpath))
        tm = tj.get(args.precision, {})
        if tj.get("rows_per_launch") == rows_per_launch and top in tm:
            traffic = tm[top]["read"] + tm[top]["write"]   # bytes per launch from the committed ncu --set full capture
    if top is not None:
        if top in FLOP_PER_ROW:
            roof = {"kernel": top, "bound": "tensor", "achieved": kern[top]["tflops"], "peak": peaks["tflops"],
                    "unit": "TFLOP/s", "frac": kern[top]["tflops"] / peaks["tflops"], "traffic": traffic,
                    "peak_source": peaks["src"] + " bf16 sustained"}
        else:
            roof = {"kernel": top, "bound": "hbm", "achieved": kern[top]["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": kern[top]["gbs"] / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["src"]}

    # ================= CPU baseline (rank 0, N=1 only; bounded sample) =================
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mean_s, best_s, threads = cpu_reference_step_time(args.rows, 1, 3, 1)
        cpu = {"value": 1.0 / mean_s, "unit": "bags/s", "cores": threads, "kind": "port",
               "sample": f"3 steps x 1 bag of {args.rows}x1024 after 1 warm-up (D step + G step + Adam), oracle port, torch CPU fp32"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "bags/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": {"workload": (f"AdvMIL-ABMIL G+D step, {args.bags} synthetic bags of {args.rows}x1024 per step per GPU "
                                    "(configs[1]); D step + G step + both Adam updates") if not args.ragged else
                                   (f"AdvMIL-ABMIL G+D step, {args.bags} packed bags per step per GPU with lengths log-uniform in "
                                    f"[1024, 100000] (configs[2]; mean {rows_per_launch} rows per step)"),
                       "rows_per_step_per_gpu": rows_per_launch,
                       "precision_mode": args.precision,
                       "feature_format": ("bf16 (packed loader's bf16 storage: features rounded once at packing time, fp32 "
                                          "accumulation/statistics/parameters)" if args.precision == "bf16" else "fp32"),
                       "l2": f"inputs ({steps[0].x.numel() * steps[0].x.element_size() >> 20} MiB/step, two alternating "
                             "steps) larger than the 126 MB L2; no flush",
                       "parallelism": f"dp{world} (bags sharded, NCCL all-reduce of flat G/D grads)"},
            "e2e": {"value": e2e_value, "unit": "bags/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "h2d_gbs_per_gpu": h2d * args.steps / (float(ems.item()) / 1e3) / 1e9,
                    "transport": ("p12: the packed loader's lossless 12-bit form of the bf16 features (8 bits sign+mantissa, 4-bit "
                                  "exponent code, sparse escapes; encoded once at packing time like the bf16 rounding itself, held "
                                  "in pinned host memory), copied and decoded on the device inside the timed region"
                                  if p12 else "raw"),
                    "note": "pinned host -> device copy of every step's features overlapped with the previous step's compute "
                            "(DeviceFeeder); bound by the PCIe link when h2d_gbs_per_gpu is ~55 GB/s"},
            "gpu_launches": launches, "host_issue_ms_per_step": host_ms, "profiled_pass_ms_per_step": prof_ms / prof_steps,
            "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "clocks": clocks,
            "losses_last_step": losses,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
