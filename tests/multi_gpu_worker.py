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
    which = sys.argv[1:] or ["all"]
    if ("all" in which or "shard" in which) and os.path.exists(os.path.join(ROOT, "tests", "multi_gpu_shard.py")):
        import multi_gpu_shard

        multi_gpu_shard.run(ctx, rank, world, dist)
    ctx.finalize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
