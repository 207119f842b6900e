"""time run_steps with and without gridded outputs / calibration aggregates (cells only)"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mhm_b200 import interface, driver, synth

prob = synth.make_problem(nx=640, ny=400, n_days=20, hourly=True, routing=False, start=(1990, 1, 1))
nT = prob["time"]["nTimeSteps"]
M = 8
print("cells", prob["nCells"], "members", M, "steps", nT)
with interface.Context(0) as ctx:
    ctx.set_math_mode("fast")
    for label, flags in (("warm", None), ("no outputs", None), ("5 outputs", [3, 9, 10, 11, 16]), ("all outputs", list(range(1, 18)) + [19, 20, 21]),
                         ("default mhm_outputs.nml", list(range(1, 17)) + [19, 20, 21])):
        for k in list(ctx.domains):
            interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, k)); del ctx.domains[k]
        dom = driver.setup_domain(ctx, 1, prob, nMembers=M, member_params=[prob["params"]] * M)
        if flags is not None:
            f = np.zeros(21, dtype=np.int32)
            for v in flags: f[v - 1] = 1
            dom.set_outputs(f, -2)   # monthly windows
        dom.run_steps(1, 160); torch.cuda.synchronize()     # same size as the timed calls: buffers exist
        dt = 1e9
        for first in (161, 321):
            t0 = time.perf_counter()
            dom.run_steps(first, 160); ctx.synchronize()
            dt = min(dt, time.perf_counter() - t0)
        print("%-12s %.3e cell-steps/s" % (label, prob["nCells"] * M * 160 / dt))
