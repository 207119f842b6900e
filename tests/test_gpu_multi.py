"""Multi-GPU tests: the exchanges below the C ABI under real NCCL (csrc/comm.cu).  They need two
GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); on a one-GPU
box the two-rank test is skipped and the single-rank path of the same entry points is tested."""
import os
import subprocess
import sys

import numpy as np
import pytest

import parity
from mhm_b200 import driver, interface, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("is_f32", [False, True])
def test_shared_forcing_single_rank(is_f32):
    """nranks = 1: set_meteo_shared is an asynchronous upload (+ widening of float32); no NCCL needed"""
    import torch

    prob = synth.make_problem(nx=40, ny=30, n_days=3, hourly=True)
    if is_f32:
        prob["forcing"] = {k: v.astype(np.float32).astype(np.float64) for k, v in prob["forcing"].items()}
    nT, n = prob["time"]["nTimeSteps"], prob["nCells"]
    with interface.Context() as ctx:
        ctx.set_math_mode("fast")
        assert ctx.comm_init(None) == (1, 0)
        dom = driver.setup_domain(ctx, 1, prob)
        dom.run_steps(1, nT)
        want = dom.get_runoff(), dom.get_state("L1_soilMoist")
        dom = driver.setup_domain(ctx, 2, prob, upload_forcing=False)
        keep = []
        for f in range(1, nT + 1, 24):
            for v in ("pre", "temp", "pet"):
                a = prob["forcing"][v][f - 1: f + 23]
                t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32 if is_f32 else np.float64)).pin_memory()
                keep.append(t)
                dom.set_meteo_shared(v, t.data_ptr(), n, f, 24, is_f32=is_f32)
            dom.run_steps(f, 24)
        parity.assert_bit_exact(dom.get_runoff(), want[0], "discharge")
        parity.assert_bit_exact(dom.get_state("L1_soilMoist"), want[1], "soil moisture")
        assert dom.meteo_h2d_bytes() == 3 * nT * n * (4 if is_f32 else 8)


def test_exchanges_under_nccl_two_ranks():
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n_gpus(), 4)),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
    assert r.stdout.count("MULTI_GPU_OK") >= 6  # shared forcing x3, domain per GPU, sharded domain x2
