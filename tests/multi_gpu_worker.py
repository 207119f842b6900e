"""Worker of tests/test_gpu_multi.py: launched with torchrun, one rank per GPU, NCCL.
Checks the exchanges the library makes below the C ABI (csrc/comm.cu) against single-GPU runs.
Prints 'MULTI_GPU_OK <what>' lines; any failure raises."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import parity  # noqa: E402
from mhm_b200 import driver, interface, synth  # noqa: E402

STATES = ["L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"]


def shared_forcing(ctx, rank, world, is_f32):
    """every rank runs the same domain (other members); forcing chunks come through
    mhm_cuda_set_meteo_shared with only this rank's rows valid in host memory"""
    prob = synth.make_problem(nx=90, ny=60, n_days=6, hourly=True, start=(1990, 12, 28))
    nT, n = prob["time"]["nTimeSteps"], prob["nCells"]
    if is_f32:  # forcing as stored in a float32 NetCDF file
        prob["forcing"] = {k: v.astype(np.float32).astype(np.float64) for k, v in prob["forcing"].items()}
    rng = np.random.default_rng(100 + rank)
    P = dict(prob["params"], L1_kSlowFlow=prob["params"]["L1_kSlowFlow"] * rng.uniform(0.9, 1.1))
    dom = driver.setup_domain(ctx, 1, prob, member_params=[P])
    dom.run_steps(1, nT)
    want = {k: dom.get_state(k) for k in STATES}
    want["Q"] = dom.get_runoff()
    interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, 1))
    dom = driver.setup_domain(ctx, 1, prob, member_params=[P], upload_forcing=False)
    chunk = 37
    firsts = list(range(1, nT + 1, chunk))
    h2d = 0

    def upload(k):
        nonlocal h2d
        f, cnt = firsts[k], min(chunk, nT - firsts[k] + 1)
        lo, rows = interface.shared_rows(cnt, world, rank)
        for v in ("pre", "temp", "pet"):
            a = prob["forcing"][v][f - 1: f - 1 + cnt]
            host = np.full(a.shape, np.nan, dtype=np.float32 if is_f32 else np.float64)
            host[lo: lo + rows] = a[lo: lo + rows]      # rows of other ranks stay NaN: never read
            t = torch.from_numpy(host).pin_memory()
            keep.append(t)
            dom.set_meteo_shared(v, t.data_ptr(), n, f, cnt, is_f32=is_f32)
            h2d += rows * n * (4 if is_f32 else 8)

    keep = []
    upload(0)
    for k, f in enumerate(firsts):
        dom.run_steps(f, min(chunk, nT - f + 1))
        if k + 1 < len(firsts):
            upload(k + 1)
    ctx.synchronize()
    for k in STATES:
        parity.assert_bit_exact(dom.get_state(k), want[k], "shared forcing: " + k)
    parity.assert_bit_exact(dom.get_runoff(), want["Q"], "shared forcing: discharge")
    assert dom.meteo_h2d_bytes() == h2d, (dom.meteo_h2d_bytes(), h2d)
    full = 3 * nT * n * (4 if is_f32 else 8)
    assert h2d <= full // world + 3 * len(firsts) * n * 8
    interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, 1))
    if rank == 0:
        print("MULTI_GPU_OK shared forcing %s: %d ranks, %d of %d forcing bytes copied per rank" % (
            "f32" if is_f32 else "f64", world, h2d, full), flush=True)


def domain_per_gpu(ctx, rank, world):
    """BASELINE config 2: independent domains, one per GPU (the reference's MPI mode deals domains to
    ranks round robin, common/mo_common_read_config.F90:416-437), with a restart in the middle: every
    rank registers ITS domain of the module-global arrays (ld = nCellsTot, offset = s1 - 1), runs
    half of the period, reads the states back into the globals, registers the domain anew with
    read_states = 1 and finishes; the second half must equal the uninterrupted run bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_boundary import RSTATES, restart_problem, snapshot

    probs = [synth.make_problem(nx=26 + 7 * d, ny=15 + 3 * d, n_days=8, hourly=True, seed=11 + d,
                                start=(1990, 6, 1)) for d in range(world)]
    ncell = [p["nCells"] for p in probs]
    offs = np.concatenate([[0], np.cumsum(ncell)]).astype(int)
    ntot = int(offs[-1])
    pack = lambda key, sub: np.ascontiguousarray(np.concatenate([p[sub][key] for p in probs], axis=-1))
    G = {"params": {k: pack(k, "params") for k in probs[0]["params"] if k in interface.PARAM_NAMES},
         "states": {k: pack(k, "states0") for k in STATES},
         "forcing": {k: pack(k, "forcing") for k in ("pre", "temp", "pet")}}
    me, off = probs[rank], int(offs[rank])
    nT, k = me["time"]["nTimeSteps"], 4 * 24
    ctx.set_math_mode("strict")

    def register(iDomain, prob, states, read_states):
        d = ctx.register_domain(iDomain, prob["nCells"], prob["nH"], prob["nLAI"], prob["nLC"],
                                prob["processMatrix"], timestep_h=1, read_states=read_states)
        d.set_meteo_config(prob["pet_case"], 24, True, False, synth.FNIGHT_PREC, synth.FNIGHT_PET,
                           synth.FNIGHT_TEMP, synth.EVAP_COEFF)
        d.set_time(prob["time"])
        for name, a in G["params"].items():
            d.set_param(name, a, ld=ntot, offset=off)
        for name, a in states.items():
            d.set_state(name, a, ld=ntot, offset=off)
        net = prob["net"]
        d.set_network(net)
        d.set_reg_rout(net["rout_param"], net["L11_length"][: net["nNodes"] - 1],
                       net["L11_slope"][: net["nNodes"] - 1], net["L11_nLinkFracFPimp"])
        return d

    full = register(rank + 1, me, G["states"], False)
    for name, a in G["forcing"].items():
        full.set_meteo(name, a, ld=ntot, offset=off)
    full.run_steps(1, nT)
    want = snapshot(full)
    interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, rank + 1))
    first = register(rank + 1, me, G["states"], False)
    for name, a in G["forcing"].items():
        first.set_meteo(name, a, ld=ntot, offset=off)
    first.run_steps(1, k)
    # "write_restart_files": the states into this domain's section of the global arrays
    H = {name: np.array(a, copy=True) for name, a in G["states"].items()}
    for name in STATES:
        first.get_state(name, out=H[name], offset=off)
    rs = {name: first.get_routing_state(name) for name in RSTATES}
    interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, rank + 1))
    sub = restart_problem(me, k, {name: np.ascontiguousarray(H[name][..., off: off + me["nCells"]]) for name in STATES}, rs)
    second = register(rank + 1, sub, H, True)
    for name, a in G["forcing"].items():
        second.set_meteo(name, np.ascontiguousarray(a[k:]), ld=ntot, offset=off)
    for name in ("L11_qOUT", "L11_qTIN", "L11_qTR", "L11_qMod"):
        second.set_routing_state(name, rs[name])
    second.set_c1c2(rs["L11_C1"], rs["L11_C2"])
    second.run_steps(1, nT - k)
    got = snapshot(second)
    for name in STATES + RSTATES:
        parity.assert_bit_exact(got[name], want[name], "domain %d after the restart: %s" % (rank + 1, name))
    parity.assert_bit_exact(got["Q"], np.ascontiguousarray(want["Q"][:, k:]), "domain %d: gauge series" % (rank + 1))
    interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, rank + 1))
    del ctx.domains[rank + 1]
    done = torch.tensor([1], device="cuda")
    dist.all_reduce(done)
    if rank == 0:
        print("MULTI_GPU_OK %d domains (%s cells), one per GPU, restart after %d of %d steps" % (
            world, ncell, k, nT), flush=True)


def main():
    rank, local, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = interface.Context(local)
    assert ctx.comm_init(dist) == (world, rank)
    info = ctx.comm_info()
    assert info["nranks"] == world and info["rank"] == rank and info["nccl_version"] > 0, info
    for mode in ("strict", "fast"):
        ctx.set_math_mode(mode)
        shared_forcing(ctx, rank, world, is_f32=False)
    shared_forcing(ctx, rank, world, is_f32=True)
    domain_per_gpu(ctx, rank, world)
    which = sys.argv[1:] or ["all"]
    if ("all" in which or "shard" in which) and os.path.exists(os.path.join(ROOT, "tests", "multi_gpu_shard.py")):
        import multi_gpu_shard

        multi_gpu_shard.run(ctx, rank, world, dist)
    ctx.finalize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
