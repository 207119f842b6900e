#!/usr/bin/env python
"""bench.py -- L1 cell-timesteps/s of the mHM hot path on N B200s (BASELINE.json metric).

Workload (config.workload): the per-GPU share of BASELINE config 5 -- a synthetic ~1M-cell
continental domain x 32 ensemble members per GPU (256 members on 8 GPUs), hourly forcing,
PET as input, Feddes soil moisture (nH = 2), level-scheduled Muskingum routing (case 1) on a
Scheidegger-type river network with L11 == L1.  Weak scaling: every rank owns 32 members of
the same domain, there is no data-path collective (members are independent evaluations,
SURVEY.md 8e); rank 0 gathers the members' gauge series in the e2e leg.

One "step" = one chunk of `block_hours` model steps for all cells and members of the rank:
fused cell kernel (meteo prologue + cascade) + routing + gauge extraction.
  value : forcing chunk resident in HBM (bound zero-copy), CUDA events on the library's
          stream, max over ranks.  Every step reads a different part of a forcing chunk that
          is larger than L2 (1M cells x 24 B x hours >> 126 MB).
  e2e   : the same steps through the C ABI with HOST buffers: H2D of the chunk's forcing from
          pinned memory, the run, D2H of the chunk's gauge series.
  --impl reference : the CPU restatement of the reference (oracle/, OpenMP over cells, serial
          routing -- the reference's own loop structure; the Fortran original cannot be built
          here) on a bounded sample of the same workload, all host cores.
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

# The one JSON line goes to the process's original stdout; everything else that writes to file
# descriptor 1 during the run (NCCL prints its version there) is sent to stderr.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj, flush=True):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=1180)
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--members", type=int, default=32, help="ensemble members per GPU")
    ap.add_argument("--block-hours", type=int, default=128)
    ap.add_argument("--mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-routing", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cells", type=int, default=100000)
    ap.add_argument("--cpu-hours", type=int, default=24)
    ap.add_argument("--mpr", action="store_true",
                    help="time mpr_eval (gamma -> all L1 effective parameters) on the per-GPU share of "
                         "BASELINE config 4: ~1.45e7 L0 cells (500 m) under a 1/16 deg L1 grid, factor 14")
    ap.add_argument("--shard", action="store_true",
                    help="strong scaling of ONE domain cut into sub-catchments (SURVEY 8e-3, BASELINE "
                         "config 4 shape): cut-link outflow series exchanged over NCCL once per time block")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line).
    NVML in-process (one sample every 10 ms; the timed region of the default run is shorter
    than one nvidia-smi invocation), nvidia-smi as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz, self.nvml, self.h = None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        nv = self.nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True,
                             text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 7 and r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
            self.max_mhz = float(r[1])
            for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"]):
                if r[3 + i].lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.01 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def member_params(base, rng, m):
    """member m's parameter set: the base set scaled inside the physical ranges (joint factors
    keep FC <= Sat, WP <= FC, k0 <= k1 <= k2 intact)"""
    if m == 0:
        return base
    P = {}
    f = lambda lo, hi: rng.uniform(lo, hi)
    soil = f(0.85, 1.15)
    rec = f(0.85, 1.15)
    for k, v in base.items():
        if k in ("L1_soilMoistFC", "L1_soilMoistSat", "L1_wiltingPoint"):
            P[k] = v * soil
        elif k in ("L1_kFastFlow", "L1_kSlowFlow", "L1_kBaseFlow"):
            P[k] = v * rec
        elif k in ("L1_fRoots", "L1_fSealed", "L1_karstLoss", "L1_jarvis_thresh_c1", "latitude"):
            P[k] = v
        elif k in ("L1_degDayNoPre", "L1_degDayMax"):
            P[k] = v  # keep ddmax >= ddnoprec
        else:
            P[k] = v * f(0.9, 1.1)
    return P


def build_problem(args, n_steps_total):
    from mhm_b200 import synth

    n_days = (n_steps_total + 23) // 24 + 1
    rng = np.random.default_rng(synth.SEED)
    t0 = time.time()
    from mhm_b200.interface import routing_order
    net = None
    if not args.no_routing:
        net = synth.make_network(rng, args.nx, args.ny, routing_order, 0.85, 1, 1, 2, 3, None)
        n = net["nCells1"]
    else:
        n = int(args.nx * args.ny * 0.85)
    prob = {"nH": 2, "nLAI": 12, "nLC": 2, "timestep_h": 1, "hourly": True, "soil_case": 1,
            "pet_case": -1, "rout_case": 1, "read_weights": False, "nCells": n, "net": net,
            "nTstepForcingDay": 24}
    prob["processMatrix"] = synth.process_matrix(1, -1, 0 if net is None else 1)
    prob["time"] = {"jul_start": synth.JUL_1990_01_01 + 150, "nTimeSteps": n_days * 24,
                    "warming_days": 0, "timeStep_LAI_input": 0, "lc_year_start": 1989,
                    "LCyearId": np.array([1, 1, 2, 2, 2, 2, 2, 2], dtype=np.int32)}
    prob["params"] = synth.make_params(rng, n, 2, 12, 2, -1)
    prob["horizon_depth"] = np.array([200.0, 400.0])
    prob["states0"] = synth.default_states(n, 2, prob["horizon_depth"])
    prob["inflowQ"] = np.zeros((0, n_days))
    prob["setup_s"] = time.time() - t0
    return prob, rng


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mhm_b200 import driver, ensemble, interface, synth

    rank = int(os.environ.get("RANK", 0))
    local = local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K, T, M = args.warmup, args.steps, args.block_hours, args.members
    n_total = (2 * (W + K) + 1) * T
    prob, rng = build_problem(args, n_total)
    n = prob["nCells"]
    ctx = interface.Context(local)
    ctx.set_math_mode(args.mode)
    mrng = np.random.default_rng(1000 + rank)
    dom = driver.setup_domain(ctx, 1, prob, nMembers=1, upload_forcing=False) if M == 1 else None
    if dom is None:
        # members are uploaded one at a time to bound host memory
        dom = ctx.register_domain(1, n, 2, 12, 2, prob["processMatrix"], timestep_h=1, nMembers=M)
        dom.set_meteo_config(-1, 24, True, False, synth.FNIGHT_PREC, synth.FNIGHT_PET,
                             synth.FNIGHT_TEMP, synth.EVAP_COEFF)
        dom.set_time(prob["time"])
        if prob["net"] is not None:
            dom.set_network(prob["net"])
        for m in range(M):
            P = member_params(prob["params"], mrng, m)
            for name, arr in P.items():
                dom.set_param(name, arr, member=m)
            for name, arr in prob["states0"].items():
                dom.set_state(name, arr, member=m)
            if prob["net"] is not None:
                net = prob["net"]
                dom.set_reg_rout(net["rout_param"] * mrng.uniform(0.95, 1.05, 5),
                                 net["L11_length"][: net["nNodes"] - 1],
                                 net["L11_slope"][: net["nNodes"] - 1], net["L11_nLinkFracFPimp"],
                                 member=m)
    # forcing chunk of T hours, generated on the device, mirrored into pinned host memory
    g = torch.Generator(device="cuda")
    g.manual_seed(synth.SEED + rank)
    hours = torch.arange(T, device="cuda", dtype=torch.float64)[:, None]
    wet = torch.rand((T, n), generator=g, device="cuda") < 0.2
    gam = torch.distributions.Gamma(torch.tensor(0.7, device="cuda", dtype=torch.float64),
                                    torch.tensor(1.0 / 1.6, device="cuda", dtype=torch.float64))
    pre = torch.where(wet, gam.sample((T, n)), torch.zeros((), device="cuda", dtype=torch.float64))
    temp = (12.0 + 4.0 * torch.sin(2 * np.pi * (hours % 24) / 24.0)
            + 2.0 * torch.randn((T, n), generator=g, device="cuda", dtype=torch.float64))
    pet = (torch.clamp(0.15 * torch.sin(np.pi * ((hours % 24) - 6.0) / 12.0), min=0.0)
           * (0.8 + 0.4 * torch.rand((T, n), generator=g, device="cuda", dtype=torch.float64)))
    dev = {"pre": pre.contiguous(), "temp": temp.contiguous(), "pet": pet.contiguous()}
    host = {k: torch.empty((T, n), dtype=torch.float64, pin_memory=True).copy_(v) for k, v in dev.items()}
    torch.cuda.synchronize()
    nG = dom.nGaugesTotal
    q_host = np.zeros((M, max(nG, 1), prob["time"]["nTimeSteps"]))
    gathered = {}

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step_value(i):
        first = i * T + 1
        for k, v in dev.items():
            dom.set_meteo_device(k, v.data_ptr(), first, T)
        dom.run_steps(first, T)

    def upload(i):
        for k, v in host.items():
            dom.set_meteo_host_ptr(k, v.data_ptr(), n, i * T + 1, T, async_copy=True)

    def step_e2e(i):
        # chunk i was uploaded while chunk i-1 was computing (double buffered in the library);
        # every step issues exactly one chunk upload (H2D) and one gauge-series download (D2H)
        first = i * T + 1
        dom.run_steps(first, T)
        upload(i + 1)
        if nG:
            for m in range(M):
                dom.get_runoff(first, T, member=m, out=q_host[m])
            # the ensemble's only exchange: every member's gauge series of the chunk on rank 0
            local = np.ascontiguousarray(q_host[:, :, first - 1: first - 1 + T])
            gathered["q"] = ensemble.gather_runoff(local, M * world, dist if world > 1 else None,
                                                   device=torch.device("cuda", local_rank))
        else:
            dom.get_state("L1_satSTW")

    def timed(fn, i0):
        for i in range(W):
            fn(i0 + i)
        barrier()
        ctx.kernel_stats_reset(True)
        sampler = ClockSampler(local)
        sampler.start()
        ctx.event_record(0)
        wall0 = time.time()
        for i in range(K):
            fn(i0 + W + i)
        ctx.event_record(1)
        barrier()
        wall = time.time() - wall0
        ms = ctx.event_elapsed_ms(0, 1)
        clocks = sampler.summary()
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, clocks

    ms_v, wall_v, clocks = timed(step_value, 0)
    cell_ms, cell_launches = ctx.kernel_stats(0)
    rout_ms, rout_launches = ctx.kernel_stats(1)
    upload(W + K)
    ms_e, wall_e, _ = timed(step_e2e, W + K)
    units_per_step = float(n) * M * T
    value = units_per_step * K * world / (ms_v * 1e-3)
    e2e_value = units_per_step * K * world / (wall_e)
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
        # algorithmic bytes per cell-step-member: forcing 24 B shared by M members
        # + 8 B total runoff written for the routing
        bytes_per_unit = 24.0 / M + 8.0
        launches_cell = max(1, cell_launches)
        cell_avg_ms = cell_ms / launches_cell
        units_per_launch = units_per_step * K / launches_cell
        achieved = units_per_launch * bytes_per_unit / (cell_avg_ms * 1e-3) / 1e9
        dfma = ctx.measure_dfma_peak()
        traffic = None
        try:  # DRAM bytes per unit of the same kernel from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_cell_traffic.json")))
            traffic = tj["dram_bytes_per_unit"] * units_per_launch / (cell_avg_ms * 1e-3) / 1e9
        except Exception:
            pass
        # useful fp64 lane operations per unit (DFMA + DADD + DMUL, predicated on), same capture
        fp64_ops_per_unit = 94.0
        out = {
            "metric": "L1 cell-timesteps/s", "value": value, "unit": "cell-timesteps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_v / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "BASELINE config 5 per-GPU share: synthetic %d-cell domain x %d members "
                            "per GPU, hourly forcing, %s" % (
                                n, M, "Muskingum routing case 1 (L11 == L1, %d nodes)" % n
                                if prob["net"] is not None else "no routing"),
                "cells": n, "members_per_gpu": M, "block_hours": T, "math_mode": args.mode,
                "nH": 2, "l2_policy": "inputs larger than L2 (forcing chunk %.1f GB, states+params "
                                      "%.1f GB per step)" % (3 * T * n * 8 / 1e9, 88 * 8 * n * M / 1e9),
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "cell_block_kernel_%s<2>" % args.mode,
                "kernel_ms_per_launch": cell_avg_ms, "units_per_launch": units_per_launch,
                "bytes_per_unit": bytes_per_unit,
                "kernel_share_of_step": cell_ms / ms_v,
            },
            "roofline_fp64": {
                "note": "the fused kernel is fp64-pipe bound (BASELINE.md 3); dfma_peak measured "
                        "live by a dependent-chain-free DFMA loop",
                "dfma_peak_per_s": dfma, "cell_kernel_units_per_s": units_per_launch / (cell_avg_ms * 1e-3),
                "fp64_lane_ops_per_unit": fp64_ops_per_unit,
                "achieved_lane_ops_per_s": fp64_ops_per_unit * units_per_launch / (cell_avg_ms * 1e-3),
                "frac": fp64_ops_per_unit * units_per_launch / (cell_avg_ms * 1e-3) / dfma if dfma else None,
            },
            "routing": {"ms": rout_ms, "kernel_launches": rout_launches, "share_of_step": rout_ms / ms_v},
            "e2e": {"value": e2e_value, "unit": "cell-timesteps/s",
                    "h2d_bytes_per_step": 3 * T * n * 8,
                    "d2h_bytes_per_step": M * max(nG, 1) * T * 8, "ms_per_step": wall_e * 1e3 / K},
            "gpu_launches": int(cell_launches + rout_launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args)
        emit(out)
    ctx.finalize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_mpr(args):
    """MPR on the device: every transfer function and all (13 + 8 nH) nLC + 12 upscalings"""
    from mhm_b200 import interface, synth_mpr

    nx0, ny0, f = (3500, 4130, 14) if (args.nx, args.ny) == (1180, 1000) else (args.nx, args.ny, 14)
    prob = synth_mpr.make_mpr_problem(nx0=nx0, ny0=ny0, factor=f, nLC=2, nLAI=12, nH=2, soil_case=1,
                                      pet_case=-1, n_soil=1475, n_geo=10, fill=0.85)
    ctx = interface.Context(0)
    out = {}
    for mode in ("strict", "fast"):
        ctx.set_math_mode(mode)
        dom = ctx.register_domain(1 if mode == "strict" else 2, prob["nL1"], prob["nH"], prob["nLAI"], prob["nLC"],
                                  prob["processMatrix"])
        synth_mpr.set_mpr_inputs(dom, prob)
        for _ in range(args.warmup):
            synth_mpr.mpr_eval(dom, prob["param"])
        ctx.synchronize()
        ctx.event_record(0)
        for _ in range(args.steps):
            synth_mpr.mpr_eval(dom, prob["param"])
        ctx.event_record(1)
        ctx.synchronize()
        out[mode] = ctx.event_elapsed_ms(0, 1) / args.steps
    nH, nLC = prob["nH"], prob["nLC"]
    fields = (13 + 8 * nH) * nLC + 12          # L0 fields upscaled per evaluation
    algo_bytes = fields * prob["nL0"] * 12.0   # 8 B value + 4 B cell index per L0 cell and field
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    ms = out["fast"]
    emit(({
        "metric": "MPR L0 cells/s (gamma -> all L1 effective parameters)", "value": prob["nL0"] / (ms * 1e-3),
        "unit": "L0 cells/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "per-GPU share of BASELINE config 4: %d L0 cells (%d x %d, 85 %% fill) -> %d L1 "
                               "cells, 1475 soil types, nH = 2, 2 land-cover scenes, 12 LAI steps" % (
                                   prob["nL0"], nx0, ny0, prob["nL1"]),
                   "fields_upscaled": fields},
        "ms_strict": out["strict"], "ms_fast": out["fast"],
        "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo_bytes / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "bytes_per_unit": fields * 12.0}}), flush=True)
    ctx.finalize()


def run_shard(args):
    """one ~500k-cell domain x `members` parameter sets, sub-catchment sharded over the ranks"""
    import torch
    import torch.distributed as dist

    from mhm_b200 import interface, shard, synth
    from mhm_b200.interface import routing_order

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K, T, M = args.warmup, args.steps, min(args.block_hours, 48), args.members
    if (args.nx, args.ny) == (1180, 1000):
        args.nx, args.ny = 1000, 590  # ~500k cells at 85 % fill
    n_days = ((W + K) * T + 23) // 24
    prob = synth.make_problem(nx=args.nx, ny=args.ny, n_days=n_days, hourly=True, start=(1990, 6, 1),
                              routing_order=routing_order)
    part = shard.partition(prob["net"], world)
    ctx = interface.Context(local)
    ctx.set_math_mode(args.mode)
    run = shard.ShardedRun(ctx, prob, part, rank, world, dist if world > 1 else None, nMembers=M)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for i in range(W):
        run.run_block(i * T + 1, T)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.kernel_stats_reset(True)
    t0 = time.time()
    for i in range(W, W + K):
        run.run_block(i * T + 1, T)
    barrier()
    wall = time.time() - t0
    clocks = sampler.summary()
    if world > 1:
        t = torch.tensor([wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t[0])
    cell_ms, cell_l = ctx.kernel_stats(0)
    rout_ms, rout_l = ctx.kernel_stats(1)
    if rank == 0:
        n = prob["nCells"]
        sh = run.sub["shard"]
        emit(({
            "metric": "L1 cell-timesteps/s", "value": float(n) * M * T * K / wall, "unit": "cell-timesteps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": wall * 1e3 / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 4 shape: ONE synthetic %d-cell domain x %d members, "
                                   "sub-catchment sharded over %d GPUs, %d cut links, routing case 1" % (
                                       n, M, world, sh["n_ghost"]),
                       "cells": n, "members": M, "block_hours": T, "math_mode": args.mode,
                       "cells_per_shard": np.bincount(part, minlength=world).tolist(),
                       "exchange_bytes_per_block": int(sh["n_ghost"]) * M * T * 8},
            "rank0": {"cell_ms": cell_ms, "routing_ms": rout_ms, "launches": int(cell_l + rout_l)},
            "gpu_launches": int(cell_l + rout_l), "clocks": clocks}), flush=True)
    ctx.finalize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_problem(args):
    """bounded sample of the bench workload for the CPU arm: the first ~cpu_cells cells of a
    domain of the same kind, one member, cpu_hours model steps"""
    from mhm_b200 import synth
    from mhm_b200.interface import routing_order

    nx = int(round((args.cpu_cells / 0.85) ** 0.5 * 1.086))
    ny = int(round(args.cpu_cells / 0.85 / nx))
    prob = synth.make_problem(nx=nx, ny=ny, n_days=max(1, (args.cpu_hours + 23) // 24), hourly=True,
                              routing=not args.no_routing, start=(1990, 6, 1),
                              routing_order=routing_order)
    return prob


def cpu_run(prob, hours, threads):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc_run

    o = orc_run.OracleRun(prob, num_threads=threads)
    t0 = time.time()
    o.run(1, hours)
    return time.time() - t0


def cpu_baseline(args):
    cores = os.cpu_count() or 1
    prob = cpu_problem(args)
    hours = min(args.cpu_hours, prob["time"]["nTimeSteps"])
    best = min(cpu_run(prob, hours, cores) for _ in range(2))
    return {"value": prob["nCells"] * hours / best, "unit": "cell-timesteps/s", "cores": cores,
            "kind": "port",
            "sample": "%d cells x 1 member x %d hourly steps, OpenMP static over cells + serial "
                      "routing (reference loop structure), best of 2" % (prob["nCells"], hours)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    prob = cpu_problem(args)
    hours = min(args.cpu_hours, prob["time"]["nTimeSteps"])
    for _ in range(args.warmup):
        cpu_run(prob, hours, cores)
    t = [cpu_run(prob, hours, cores) for _ in range(args.steps)]
    total = sum(t)
    v = prob["nCells"] * hours * args.steps / total
    sample = ("%d cells x 1 member x %d hourly steps per step; oracle/ restatement of the reference "
              "(Fortran original not buildable here)" % (prob["nCells"], hours))
    emit(({
        "impl": "reference", "metric": "L1 cell-timesteps/s", "value": v, "unit": "cell-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 5 per-GPU share (synthetic ~1M-cell domain x 32 members per "
                               "GPU, hourly forcing, Muskingum routing case 1), timed on a bounded sample: see "
                               "cpu_baseline.sample",
                   "cells": prob["nCells"], "members_per_gpu": 1, "block_hours": hours},
        "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.mpr:
        run_mpr(a)
    elif a.shard:
        run_shard(a)
    else:
        run_ours(a)
