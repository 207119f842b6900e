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
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed legs")
    ap.add_argument("--full-run", action="store_true",
                    help="the north-star run: --hours model hours (default one year) of the whole domain x members, "
                         "streamed forcing, default monthly outputs, checked against the oracle")
    ap.add_argument("--hours", type=int, default=8760)
    ap.add_argument("--no-outputs", action="store_true", help="--full-run without gridded outputs")
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


def kernel_profile():
    """profiles/cell_kernel_ncu.json: what only a profiler can count (fp64 instructions and useful
    lane operations per cell-step, DRAM bytes per cell-step), valid for the kernel sources it names"""
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import cell_profile

    path = os.path.join(ROOT, "profiles", "cell_kernel_ncu.json")
    try:
        prof = json.load(open(path))
    except Exception as e:  # noqa: BLE001
        return None, "profiles/cell_kernel_ncu.json unreadable: %s" % e
    now = cell_profile.kernel_hash()
    if prof.get("kernel_hash") != now:
        return None, "profiles/cell_kernel_ncu.json is stale: captured for kernel sources %s, built from %s" % (
            prof.get("kernel_hash"), now)
    return prof, None


class ParityChecker:
    """bench.py checks what it times: member 0 of the rank against the CPU oracle (test
    infrastructure, oracle/) -- a sample of cells through every step of both legs (states and the
    last step's fluxes, <= 1e-9 relative), and the gauge discharge of the first chunk from a
    full-domain oracle run (<= 1e-8 relative).  A mismatch raises: no number is printed."""

    def __init__(self, prob, host_forcing, T, n_chunks, n_sample=2048):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc_run
        import parity

        self.orc_run, self.parity, self.prob, self.T = orc_run, parity, prob, T
        n = prob["nCells"]
        rng = np.random.default_rng(7)
        self.cells = np.sort(rng.choice(n, size=min(n_sample, n), replace=False))
        sub = dict(prob)
        sub["nCells"] = len(self.cells)
        sub["net"] = None
        sub["params"] = {k: (np.ascontiguousarray(v[..., self.cells]) if k != "rout_param" else v)
                         for k, v in prob["params"].items()}
        sub["states0"] = {k: np.ascontiguousarray(v[..., self.cells]) for k, v in prob["states0"].items()}
        chunk = {k: np.ascontiguousarray(v.numpy()[:, self.cells]) for k, v in host_forcing.items()}
        sub["forcing"] = {k: np.ascontiguousarray(np.tile(v, (n_chunks, 1))) for k, v in chunk.items()}
        self.sample = orc_run.OracleRun(sub, num_threads=os.cpu_count() or 1)
        self.done = 0
        self.worst = {"state": 0.0, "flux": 0.0, "gauge": None}
        self.host_forcing = host_forcing

    def check_cells(self, dom, tt_last, what):
        o, P = self.sample, self.parity
        o.run(self.done + 1, tt_last)
        self.done = tt_last
        for name in ("L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"):
            got = dom.get_state(name, member=0)[..., self.cells]
            self.worst["state"] = max(self.worst["state"], P.assert_close(got, o.S[name], "%s: %s" % (what, name)))
        for name in ("L1_total_runoff", "L1_aETSoil", "L1_infilSoil", "L1_baseflow", "L1_slowRunoff", "L1_melt"):
            got = dom.get_flux(name, member=0)[..., self.cells]
            P.assert_close(got, o.F[name], "%s: %s" % (what, name))
            # reported figure: relative to max(|a|, |b|, 1e-6 mm) -- fluxes that are differences
            # of O(1) terms reach 1e-17 where a pure relative error means nothing (tests/parity.py)
            den = np.maximum(np.maximum(np.abs(got), np.abs(o.F[name])), 1e-6)
            self.worst["flux"] = max(self.worst["flux"], float((np.abs(got - o.F[name]) / den).max()))

    def check_gauges(self, dom):
        """gauge series of the first chunk: the whole domain through the oracle (cells + serial routing)"""
        prob, T = self.prob, self.T
        if prob["net"] is None:
            return
        full = dict(prob)
        full["forcing"] = {k: v.numpy() for k, v in self.host_forcing.items()}
        o = self.orc_run.OracleRun(full, num_threads=os.cpu_count() or 1)
        o.run(1, T)
        got = dom.get_runoff(1, T, member=0)[:, :T]
        self.worst["gauge"] = self.parity.assert_close(got, o.mRM_runoff[:, :T], "gauge discharge of the first chunk",
                                                       rtol=self.parity.RTOL_Q)
        assert np.abs(o.mRM_runoff[:, :T]).max() > 0.0


def run_ours(args):
    rank = int(os.environ.get("RANK", 0))
    local = local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # the CPU baseline runs before any process group exists (no other rank spins in NCCL beside it)
    cpu = cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    import torch
    import torch.distributed as dist

    from mhm_b200 import driver, ensemble, interface, synth

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K, T, M = args.warmup, args.steps, args.block_hours, args.members
    n_chunks = 2 * (W + K) + 1
    n_total = n_chunks * T
    prob, rng = build_problem(args, n_total)
    n = prob["nCells"]
    ctx = interface.Context(local)
    ctx.set_math_mode(args.mode)
    ctx.comm_init(dist if world > 1 else None)  # the library's own NCCL communicator (shared forcing)
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
                rp = net["rout_param"] if m == 0 else net["rout_param"] * mrng.uniform(0.95, 1.05, 5)
                dom.set_reg_rout(rp, net["L11_length"][: net["nNodes"] - 1],
                                 net["L11_slope"][: net["nNodes"] - 1], net["L11_nLinkFracFPimp"],
                                 member=m)
    # forcing chunk of T hours -- the SAME on every rank (the ranks run members of one domain) --
    # generated on the device, mirrored into pinned host memory
    g = torch.Generator(device="cuda")
    g.manual_seed(synth.SEED)
    torch.manual_seed(synth.SEED)
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
    if world > 1:  # bit-identical forcing on every rank, whatever the ranks' generators did
        for v in dev.values():
            dist.broadcast(v, src=0)
    host = {k: torch.empty((T, n), dtype=torch.float64, pin_memory=True).copy_(v) for k, v in dev.items()}
    torch.cuda.synchronize()
    nG = dom.nGaugesTotal
    q_host = np.zeros((M, max(nG, 1), prob["time"]["nTimeSteps"]))
    gathered = {}
    checker = ParityChecker(prob, host, T, n_chunks) if (rank == 0 and not args.no_parity) else None

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
        # every rank copies 1/world of the chunk from its pinned host memory; the ranks all-gather
        # the rest over NVLink on the library's upload stream (mhm_cuda_set_meteo_shared)
        for k, v in host.items():
            dom.set_meteo_shared(k, v.data_ptr(), n, i * T + 1, T)

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
        h2d0 = dom.meteo_h2d_bytes()
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
        h2d = (dom.meteo_h2d_bytes() - h2d0) / K
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, clocks, h2d

    ms_v, wall_v, clocks, _ = timed(step_value, 0)
    cell_ms, cell_launches = ctx.kernel_stats(0)
    rout_ms, rout_launches = ctx.kernel_stats(1)
    ctx.kernel_stats_reset(False)
    parity = None
    if checker is not None:
        checker.check_cells(dom, (W + K) * T, "value leg")
        checker.check_gauges(dom)
    barrier()
    upload(W + K)
    ms_e, wall_e, clocks_e, h2d_per_step = timed(step_e2e, W + K)
    if checker is not None:
        checker.check_cells(dom, 2 * (W + K) * T, "e2e leg")
        parity = {"checked": True, "member": 0, "cells_sampled": int(len(checker.cells)),
                  "steps_checked": 2 * (W + K) * T, "max_rel_state": checker.worst["state"],
                  "max_rel_flux": checker.worst["flux"], "gauge_max_rel_first_chunk": checker.worst["gauge"],
                  "tolerance": {"cells": 1e-9, "gauges": 1e-8}}
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
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # algorithmic bytes per cell-step-member: forcing 24 B shared by M members
        # + 8 B node runoff written for the routing
        bytes_per_unit = 24.0 / M + 8.0
        launches_cell = max(1, cell_launches)
        cell_avg_ms = cell_ms / launches_cell
        units_per_launch = units_per_step * K / launches_cell
        cell_rate = units_per_launch / (cell_avg_ms * 1e-3)  # cell-steps/s inside the kernel
        hbm_achieved = cell_rate * bytes_per_unit / 1e9
        dfma = ctx.measure_dfma_peak()
        prof, stale = kernel_profile()
        if stale:
            sys.stderr.write("bench.py: WARNING: %s -- roofline.frac withheld\n" % stale)
        sm_hz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        props = torch.cuda.get_device_properties(local)
        # the fp64 pipe of an SM sub-partition takes one warp instruction every two cycles (16 lanes)
        pipe_peak = props.multi_processor_count * 4 * sm_hz / 2.0
        roof = {"bound": "fp64", "unit": "fp64 warp-instructions/s", "peak": pipe_peak,
                "peak_source": "SMs x 4 sub-partitions x measured SM clock / 2 (consistent with the DFMA peak "
                               "measured live: %.3e lane-FMA/s)" % dfma,
                "kernel": "cell_block_kernel_%s<2, hourly, uniform, fused>" % args.mode,
                "kernel_ms_per_launch": cell_avg_ms, "units_per_launch": units_per_launch,
                "kernel_units_per_s": cell_rate, "kernel_share_of_step": cell_ms / ms_v,
                "achieved": None, "frac": None, "traffic": None}
        if prof is not None:
            fp64_rate = prof["fp64_warp_instructions_per_warp_step"] * cell_rate / 32.0
            roof.update({
                "achieved": fp64_rate, "frac": fp64_rate / pipe_peak,
                "frac_note": "fp64-pipe utilisation = fp64 warp instructions per warp-step (ncu capture of this "
                             "build, profiles/cell_kernel_ncu.json) x warp-steps/s (timed here) / peak",
                "ncu_fp64_pipe_frac": prof["fp64_pipe_pct"] / 100.0,
                "ncu_issue_active_frac": prof["issue_active_pct"] / 100.0,
                "ncu_lsu_data_pipe_frac": prof["lsu_data_pipe_pct"] / 100.0,
                "useful_frac": prof["fp64_lane_ops_per_unit"] * cell_rate / dfma if dfma else None,
                "fp64_lane_ops_per_unit": prof["fp64_lane_ops_per_unit"],
                "dfma_peak_per_s": dfma,
                "traffic": prof["dram_bytes_per_unit"] * cell_rate / 1e9, "traffic_unit": "GB/s",
                "profile_kernel_hash": prof["kernel_hash"]})
        else:
            roof["stale_profile"] = stale
        roof["hbm"] = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                       "bytes_per_unit": bytes_per_unit, "peak_source": peak_src}
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
            "roofline": roof,
            "routing": {"ms": rout_ms, "kernel_launches": rout_launches, "share_of_step": rout_ms / ms_v},
            "e2e": {"value": e2e_value, "unit": "cell-timesteps/s",
                    "h2d_bytes_per_step": int(h2d_per_step),
                    "h2d_note": "host bytes THIS rank copies per step: 1/%d of the shared forcing chunk; the rest "
                                "arrives by NCCL all-gather over NVLink (mhm_cuda_set_meteo_shared)" % world
                                if world > 1 else "the whole forcing chunk from pinned host memory",
                    "forcing_chunk_bytes": 3 * T * n * 8,
                    "d2h_bytes_per_step": M * max(nG, 1) * T * 8, "ms_per_step": wall_e * 1e3 / K,
                    "clocks": clocks_e},
            "gpu_launches": int(cell_launches + rout_launches),
            "clocks": clocks,
            "parity_checked": bool(parity),
            "parity": parity,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        emit(out)
    free_b, total_b = torch.cuda.mem_get_info()
    sys.stderr.write("bench.py: rank %d device memory in use at the end: %.1f of %.1f GB\n" % (
        rank, (total_b - free_b) / 1e9, total_b / 1e9))
    ctx.finalize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


DEFAULT_OUTPUTS = [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1]  # the reference's mhm_outputs.nml


def run_full(args):
    """The north-star run: the whole ~1M-cell domain x `members` per GPU through `--hours` model hours
    (default one year, hourly), forcing STREAMED from pinned host memory chunk by chunk (three distinct
    128-hour chunks -- warm, cold with snow, mild -- cycled, so consecutive chunks differ), across
    month / LAI / year / land-cover-scene boundaries, with the reference's default mhm_outputs.nml
    (monthly windows, 19 variables = 22 fields) accumulated on the device and every closed window of
    member 0 fetched.  Checked against the oracle: a sample of cells through ALL steps (final states
    and every monthly window) and the gauge discharge of the first 256 hours from a full-domain run."""
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    import torch
    import torch.distributed as dist

    from mhm_b200 import interface, synth

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T, M, H = args.block_hours, args.members, args.hours
    prob, rng = build_problem(args, H)
    n = prob["nCells"]
    ctx = interface.Context(local)
    ctx.set_math_mode(args.mode)
    ctx.comm_init(dist if world > 1 else None)
    mrng = np.random.default_rng(1000 + rank)
    dom = ctx.register_domain(1, n, 2, 12, 2, prob["processMatrix"], timestep_h=1, nMembers=M)
    dom.set_meteo_config(-1, 24, True, False, synth.FNIGHT_PREC, synth.FNIGHT_PET, synth.FNIGHT_TEMP,
                         synth.EVAP_COEFF)
    dom.set_time(prob["time"])
    net = prob["net"]
    dom.set_network(net)
    for m in range(M):
        for name, arr in member_params(prob["params"], mrng, m).items():
            dom.set_param(name, arr, member=m)
        for name, arr in prob["states0"].items():
            dom.set_state(name, arr, member=m)
        rp = net["rout_param"] if m == 0 else net["rout_param"] * mrng.uniform(0.95, 1.05, 5)
        dom.set_reg_rout(rp, net["L11_length"][: net["nNodes"] - 1], net["L11_slope"][: net["nNodes"] - 1],
                         net["L11_nLinkFracFPimp"], member=m)
    outputs = None if args.no_outputs else (DEFAULT_OUTPUTS, -2)
    if outputs:
        dom.set_outputs(*outputs)
    # three distinct forcing chunks in pinned host memory (the same on every rank)
    g = torch.Generator(device="cuda")
    g.manual_seed(synth.SEED)
    torch.manual_seed(synth.SEED)
    hours = torch.arange(T, device="cuda", dtype=torch.float64)[:, None]
    host = []
    for c, (t_mean, p_wet) in enumerate(((14.0, 0.2), (-3.0, 0.3), (6.0, 0.1))):
        wet = torch.rand((T, n), generator=g, device="cuda") < p_wet
        gam = torch.distributions.Gamma(torch.tensor(0.7, device="cuda", dtype=torch.float64),
                                        torch.tensor(1.0 / 1.6, device="cuda", dtype=torch.float64))
        pre = torch.where(wet, gam.sample((T, n)), torch.zeros((), device="cuda", dtype=torch.float64))
        temp = (t_mean + 4.0 * torch.sin(2 * np.pi * (hours % 24) / 24.0)
                + 2.0 * torch.randn((T, n), generator=g, device="cuda", dtype=torch.float64))
        pet = (torch.clamp(0.15 * torch.sin(np.pi * ((hours % 24) - 6.0) / 12.0), min=0.0)
               * (0.8 + 0.4 * torch.rand((T, n), generator=g, device="cuda", dtype=torch.float64)))
        ch = {"pre": pre, "temp": temp, "pet": pet}
        if world > 1:
            for v in ch.values():
                dist.broadcast(v, src=0)
        host.append({k: torch.empty((T, n), dtype=torch.float64, pin_memory=True).copy_(v) for k, v in ch.items()})
        del pre, temp, pet, wet, ch
    torch.cuda.empty_cache()
    firsts = list(range(1, H + 1, T))
    nG = dom.nGaugesTotal
    q_host = np.zeros((max(nG, 1), prob["time"]["nTimeSteps"]))
    windows = []  # (tt_end, {(var, horizon): member-0 field at the sample cells})
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    cells = np.sort(np.random.default_rng(7).choice(n, size=2048, replace=False))

    def upload(k):
        f, cnt = firsts[k], min(T, H - firsts[k] + 1)
        for v, t in host[k % 3].items():
            dom.set_meteo_shared(v, t.data_ptr(), n, f, cnt)

    ctx.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.time()
    upload(0)
    for k, f in enumerate(firsts):
        cnt = min(T, H - f + 1)
        dom.run_steps(f, cnt)
        if k + 1 < len(firsts):
            upload(k + 1)
        if nG:
            dom.get_runoff(f, cnt, member=0, out=q_host)
        if outputs:
            for w, tt_end in enumerate(dom.output_windows()):
                fld = {}
                for v in range(1, 22):
                    if not DEFAULT_OUTPUTS[v - 1]:
                        continue
                    for h in ((1, 2) if v in (3, 4, 17, 19) else (0,)):
                        fld[(v, h - 1 if h else -1)] = dom.get_output(w, v, h, member=0)[cells]
                windows.append((tt_end, fld))
    ctx.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.time() - t0
    clocks = sampler.summary()
    if world > 1:
        t = torch.tensor([wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t[0])
    parity = None
    if rank == 0 and not args.no_parity:
        import orc_run
        import parity as P

        sub = dict(prob)
        sub["nCells"], sub["net"] = len(cells), None
        sub["params"] = {k: (np.ascontiguousarray(v[..., cells]) if k != "rout_param" else v)
                         for k, v in prob["params"].items()}
        sub["states0"] = {k: np.ascontiguousarray(v[..., cells]) for k, v in prob["states0"].items()}
        seq = [k % 3 for k in range(len(firsts))]
        sub["forcing"] = {v: np.ascontiguousarray(np.concatenate([host[c][v].numpy()[:, cells] for c in seq])[:H])
                          for v in ("pre", "temp", "pet")}
        o = orc_run.OracleRun(sub, num_threads=os.cpu_count() or 1, outputs=outputs, max_windows=64)
        o.run(1, H)
        worst = {"state": 0.0, "output": 0.0}
        for name in ("L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"):
            got = dom.get_state(name, member=0)[..., cells]
            worst["state"] = max(worst["state"], P.assert_close(got, o.S[name], "full run: " + name))
        if outputs:
            ref_w = o.out_windows()
            assert [w[0] for w in ref_w] == [w[0] for w in windows], ([w[0] for w in ref_w], [w[0] for w in windows])
            for (tt_end, got), (_, ref) in zip(windows, ref_w):
                for key, val in got.items():
                    worst["output"] = max(worst["output"], P.assert_close(
                        val, ref[key], "full run: output %r of the window ending at step %d" % (key, tt_end)))
        # gauges: the first 256 hours of the whole domain through the oracle
        Hq = min(H, 256)
        full = dict(prob)
        full["forcing"] = {v: np.concatenate([host[c][v].numpy() for c in seq[: (Hq + T - 1) // T]])[:Hq]
                           for v in ("pre", "temp", "pet")}
        of = orc_run.OracleRun(full, num_threads=os.cpu_count() or 1)
        of.run(1, Hq)
        gq = P.assert_close(q_host[:, :Hq], of.mRM_runoff[:, :Hq], "full run: gauge discharge of the first hours",
                            rtol=P.RTOL_Q)
        parity = {"checked": True, "member": 0, "cells_sampled": int(len(cells)), "steps_checked": H,
                  "max_rel_state": worst["state"], "max_rel_output": worst["output"],
                  "output_windows_checked": len(windows), "gauge_hours_checked": Hq, "gauge_max_rel": gq,
                  "max_snowpack_mm": float(np.max(o.S["L1_snowPack"]))}
    if rank == 0:
        idx = interface.time_indices(prob["time"], 1, 24, 1, H)
        emit({
            "metric": "L1 cell-timesteps/s", "value": float(n) * M * H * world / wall, "unit": "cell-timesteps/s",
            "n_gpus": world, "steps": len(firsts), "warmup": 0, "ms_per_step": wall * 1e3 / len(firsts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "north-star run: synthetic %d-cell domain x %d members per GPU x %d hourly steps "
                                   "(%.1f days from 1990-05-31), forcing streamed from pinned host memory in %d-hour "
                                   "chunks, Muskingum routing case 1, %s" % (
                                       n, M, H, H / 24.0, T,
                                       "default mhm_outputs.nml (22 monthly fields) accumulated on the device"
                                       if outputs else "no gridded outputs"),
                       "cells": n, "members_per_gpu": M, "block_hours": T, "hours": H, "math_mode": args.mode,
                       "months_crossed": len({(s.year, s.month) for s in idx}) - 1,
                       "land_cover_scenes": sorted({s.yId for s in idx})},
            "wall_s": wall, "simulated_years_per_wall_hour": H / 8760.0 / (wall / 3600.0),
            "e2e": {"value": float(n) * M * H * world / wall, "unit": "cell-timesteps/s",
                    "h2d_bytes_per_step": int(dom.meteo_h2d_bytes() / len(firsts)),
                    "d2h_bytes_per_step": int(max(nG, 1) * T * 8 + (len(windows) * 22 * n * 8) / len(firsts))},
            "clocks": clocks, "parity_checked": bool(parity), "parity": parity})
    ctx.finalize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    W, K, T, M = args.warmup, args.steps, args.block_hours, args.members
    if (args.nx, args.ny) == (1180, 1000):
        args.nx, args.ny = 1000, 590  # ~500k cells at 85 % fill
    n_days = ((W + K) * T + 23) // 24
    prob = synth.make_problem(nx=args.nx, ny=args.ny, n_days=n_days, hourly=True, start=(1990, 6, 1),
                              routing_order=routing_order)
    part = shard.partition(prob["net"], world)
    ctx = interface.Context(local)
    ctx.set_math_mode(args.mode)
    ctx.comm_init(dist if world > 1 else None)  # the cut-link exchange runs inside the library (NCCL)
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
    run.finish()
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
    """The CPU arms' workload: the SAME synthetic domain as the GPU arm (all %d x %d x 0.85 cells, the
    same 1M-node river network and parameters) with one member and cpu_hours model steps.  Built
    without the product library: the routing order comes from the oracle's own linear-time
    restatement (oracle/mhm_oracle.c: orc_routing_order_linear)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc

    from mhm_b200 import synth  # numpy only; libmhm_cuda.so is not loaded

    prob = synth.make_problem(nx=args.nx, ny=args.ny, n_days=max(1, (args.cpu_hours + 23) // 24), hourly=True,
                              routing=not args.no_routing, start=(1990, 6, 1), routing_order=orc.routing_order)
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
            "sample": "the full %d-cell domain x 1 member x %d hourly steps, OpenMP static over cells + serial "
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
    sample = ("the full %d-cell domain x 1 member x %d hourly steps per step; oracle/ restatement of the reference "
              "(Fortran original not buildable here), %d OpenMP threads" % (prob["nCells"], hours, cores))
    assert "mhm_b200._lib" not in sys.modules or sys.modules["mhm_b200._lib"]._lib is None, \
        "the reference arm must not load libmhm_cuda.so"
    emit(({
        "impl": "reference", "metric": "L1 cell-timesteps/s", "value": v, "unit": "cell-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 5 per-GPU share (synthetic %d-cell domain, hourly forcing, Muskingum "
                               "routing case 1 on the same %d-node network); the CPU arm steps ONE member, the GPU arm "
                               "32 per GPU" % (prob["nCells"], prob["nCells"]),
                   "cells": prob["nCells"], "members_per_gpu": 1, "block_hours": hours},
        "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.full_run:
        run_full(a)
    elif a.mpr:
        run_mpr(a)
    elif a.shard:
        run_shard(a)
    else:
        run_ours(a)
