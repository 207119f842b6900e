"""GPU tests of the arguments that exist BECAUSE of the Fortran driver, and of the forcing paths
bench.py times: asynchronous double-buffered uploads in chunks, zero-copy device forcing, many
members sharing one forcing, restart (read_states), several domains packed in the module-global
arrays (ld = nCellsTot, offset = s1 - 1), host coherence (bind_host_* + sync_to_host), and
re-evaluation of a routing-case-2 domain with another celerity.

Reference behaviour: DomainLoop mHM/mo_mhm_eval.f90:127, read_states mHM/mo_mhm.f90:448-450 and
mRM/mo_mrm_routing.f90:211, array sections mHM/mo_mhm_interface_run.f90:394-457, pybind getters
pybind/src/wrapper.f90:569-660, mrm_update_param mRM/mo_mrm_mpr.f90:241-329.
"""
import copy
import ctypes as C

import numpy as np
import pytest

import orc_run
import parity
from mhm_b200 import driver, interface, synth

pytestmark = pytest.mark.gpu

STATES = ["L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"]
FLUXES = orc_run.FLUX_ORDER + ["L1_aETSoil", "L1_infilSoil"]
RSTATES = ["L11_qOUT", "L11_qTIN", "L11_qTR", "L11_qMod", "L11_C1", "L11_C2"]


@pytest.fixture(scope="module")
def ctx():
    c = interface.Context()
    yield c
    c.finalize()


def clear(ctx):
    for k in list(ctx.domains):
        interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, k))
        del ctx.domains[k]


def snapshot(dom, routing=True):
    out = {k: dom.get_variable(k) for k in STATES + FLUXES}
    if routing:
        out.update({k: dom.get_routing_state(k) for k in RSTATES})
        out["Q"] = dom.get_runoff()
    return out


def assert_same(a, b, what):
    assert a.keys() == b.keys()
    for k in a:
        parity.assert_bit_exact(a[k], b[k], "%s: %s" % (what, k))


# --------------------------------------------------------------------------- forcing paths
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_async_chunks_and_device_forcing_equal_synchronous_upload(ctx, mode):
    """The same run with forcing (a) uploaded whole and synchronously (what every other test
    does), (b) uploaded in chunks through mhm_cuda_set_meteo_async with the upload of chunk k+1
    issued right after the run of chunk k (bench.py's e2e leg), (c) bound zero-copy chunk by chunk
    through mhm_cuda_set_meteo_device (bench.py's value leg): bit-identical, and equal to the oracle."""
    import torch

    prob = synth.make_problem(nx=180, ny=110, n_days=10, hourly=True, start=(1990, 12, 27))
    nT, n, M = prob["time"]["nTimeSteps"], prob["nCells"], 3
    rng = np.random.default_rng(5)
    mp = [prob["params"]] + [dict(prob["params"], L1_kSlowFlow=prob["params"]["L1_kSlowFlow"] * rng.uniform(0.9, 1.1))
                             for _ in range(M - 1)]
    ctx.set_math_mode(mode)
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, prob, nMembers=M, member_params=mp)
    dom.run_steps(1, nT)
    ref = [snapshot_member(dom, m) for m in range(M)]
    o = orc_run.OracleRun(prob)
    o.run(1, nT)
    parity.assert_close(ref[0]["Q"], o.mRM_runoff, "gauge discharge vs oracle", rtol=parity.RTOL_Q)
    parity.assert_close(ref[0]["L1_soilMoist"], o.S["L1_soilMoist"], "soil moisture vs oracle")

    chunk = 56  # neither a divisor of the run nor of a day: 5 chunks, the last one short
    firsts = list(range(1, nT + 1, chunk))
    assert len(firsts) >= 3
    vars_ = ["pre", "temp", "pet"]
    pinned = {v: torch.from_numpy(prob["forcing"][v]).pin_memory() for v in vars_}

    # (b) asynchronous, double buffered
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, prob, nMembers=M, member_params=mp, upload_forcing=False)

    def upload(k):
        f, cnt = firsts[k], min(chunk, nT - firsts[k] + 1)
        for v in vars_:
            dom.set_meteo_host_ptr(v, pinned[v][f - 1:].data_ptr(), n, f, cnt, async_copy=True)

    upload(0)
    for k, f in enumerate(firsts):
        dom.run_steps(f, min(chunk, nT - f + 1))
        if k + 1 < len(firsts):
            upload(k + 1)  # overlaps the kernels of chunk k
    ctx.synchronize()
    for m in range(M):
        assert_same(snapshot_member(dom, m), ref[m], "async chunks, member %d" % m)

    # (c) zero copy from a device-resident array
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, prob, nMembers=M, member_params=mp, upload_forcing=False)
    devf = {v: torch.from_numpy(prob["forcing"][v]).cuda() for v in vars_}
    torch.cuda.synchronize()
    for f in firsts:
        cnt = min(chunk, nT - f + 1)
        for v in vars_:
            dom.set_meteo_device(v, devf[v][f - 1:].data_ptr(), f, cnt)
        dom.run_steps(f, cnt)
    ctx.synchronize()
    for m in range(M):
        assert_same(snapshot_member(dom, m), ref[m], "device forcing, member %d" % m)


def snapshot_member(dom, m):
    out = {k: dom.get_variable(k, member=m) for k in STATES + FLUXES + RSTATES}
    out["Q"] = dom.get_runoff(member=m)
    return out


def test_async_upload_realloc_and_reuse(ctx):
    """growing chunk sizes force the double buffers to be reallocated while a run is queued"""
    import torch

    prob = synth.make_problem(nx=60, ny=40, n_days=6, hourly=True, routing=False)
    nT, n = prob["time"]["nTimeSteps"], prob["nCells"]
    ctx.set_math_mode("fast")
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, prob)
    dom.run_steps(1, nT)
    ref = snapshot(dom, routing=False)
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, prob, upload_forcing=False)
    pinned = {v: torch.from_numpy(prob["forcing"][v]).pin_memory() for v in ("pre", "temp", "pet")}
    sizes, f = [5, 11, 30, 7, 50, 41], 1
    for i, cnt in enumerate(sizes):
        cnt = min(cnt, nT - f + 1) if i + 1 < len(sizes) else nT - f + 1
        for v, t in pinned.items():
            dom.set_meteo_host_ptr(v, t[f - 1:].data_ptr(), n, f, cnt, async_copy=True)
        dom.run_steps(f, cnt)
        f += cnt
    assert f == nT + 1
    assert_same(snapshot(dom, routing=False), ref, "growing async chunks")


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_32_members_share_forcing(ctx, mode):
    """bench.py's layout: 32 members of one domain side by side (blockIdx = tile * M + member, one
    forcing row served to all members) equal 32 single-member runs bit for bit"""
    prob = synth.make_problem(nx=48, ny=30, n_days=5, hourly=True)
    nT, M = prob["time"]["nTimeSteps"], 32
    rng = np.random.default_rng(11)
    mp = []
    for m in range(M):
        P = dict(prob["params"])
        if m:
            P["L1_kSlowFlow"] = P["L1_kSlowFlow"] * rng.uniform(0.85, 1.15)
            P["L1_soilMoistExp"] = P["L1_soilMoistExp"] * rng.uniform(0.9, 1.1)
            P["L1_maxInter"] = P["L1_maxInter"] * rng.uniform(0.8, 1.2)
            P["rout_param"] = prob["net"]["rout_param"] * rng.uniform(0.95, 1.05, 5)
        mp.append(P)
    ctx.set_math_mode(mode)
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, prob, nMembers=M, member_params=mp)
    dom.run_steps(1, nT)
    got = {m: snapshot_member(dom, m) for m in (0, 1, 13, 31)}
    for m, g in got.items():
        clear(ctx)
        single = driver.setup_domain(ctx, 1, prob, nMembers=1, member_params=[mp[m]])
        single.run_steps(1, nT)
        assert_same(g, snapshot_member(single, 0), "member %d of 32" % m)
    o = orc_run.OracleRun(prob, params=mp[31])
    o.run(1, nT)
    parity.assert_close(got[31]["Q"], o.mRM_runoff, "member 31 vs oracle", rtol=parity.RTOL_Q)


# --------------------------------------------------------------------------- restart
def restart_problem(prob, k, states, rstates):
    """the problem a restart run sees: simulation period starting k steps later (k whole days),
    forcing from there on, states from the restart file"""
    assert k % 24 == 0
    sub = copy.copy(prob)
    sub["time"] = dict(prob["time"], jul_start=prob["time"]["jul_start"] + k // 24,
                       nTimeSteps=prob["time"]["nTimeSteps"] - k)
    per = 1 if prob["hourly"] else 24
    sub["forcing"] = {v: np.ascontiguousarray(a[k // per:]) for v, a in prob["forcing"].items()}
    sub["states0"] = {s: states[s] for s in STATES}
    if prob.get("net") is not None and prob["net"]["nInflowTotal"]:
        sub["inflowQ"] = np.ascontiguousarray(prob["inflowQ"][:, k // 24:])
    sub["restart_routing"] = rstates
    return sub


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("rout_case,hourly", [(1, True), (2, False)])
def test_restart_round_trip(ctx, mode, rout_case, hourly):
    """run k steps, read states and routing state back (what write_restart_files saves,
    mHM/mo_restart.f90:53-270, mRM/mo_mrm_restart.f90:56-428), register a fresh domain with
    read_states = 1 and continue: bit-identical to the uninterrupted run; the oracle with
    read_states = 1 agrees.  With read_states the 0.5 * FC start (mo_mhm.f90:448-450) and
    reg_rout (mo_mrm_routing.f90:211) are skipped although set_reg_rout is still called."""
    prob = synth.make_problem(nx=40, ny=26, n_days=9, hourly=hourly, start=(1990, 6, 1),
                              rout_case=rout_case)
    if rout_case != 1:
        orc_run.case23_params(prob["net"])
    nT, k = prob["time"]["nTimeSteps"], 4 * 24
    ctx.set_math_mode(mode)
    clear(ctx)
    full = driver.setup_domain(ctx, 1, prob)
    full.run_steps(1, nT)
    want = snapshot(full)
    clear(ctx)
    first = driver.setup_domain(ctx, 1, prob)
    first.run_steps(1, k)
    st = {s: first.get_state(s) for s in STATES}
    rs = {s: first.get_routing_state(s) for s in RSTATES}
    # a state the 0.5 * FC start would destroy: make sure the test can see a wrong restart
    assert np.abs(st["L1_soilMoist"] - 0.5 * prob["params"]["L1_soilMoistFC"][0]).max() > 1e-3
    sub = restart_problem(prob, k, st, rs)
    clear(ctx)
    dom = driver.setup_domain(ctx, 2, sub, read_states=True)
    for s in ("L11_qOUT", "L11_qTIN", "L11_qTR", "L11_qMod"):
        dom.set_routing_state(s, rs[s])
    if rout_case == 1:
        dom.set_c1c2(rs["L11_C1"], rs["L11_C2"])           # restart file's coefficients ...
        net = sub["net"]
        dom.set_reg_rout(net["rout_param"] * 1.7, net["L11_length"][: net["nNodes"] - 1],
                         net["L11_slope"][: net["nNodes"] - 1], net["L11_nLinkFracFPimp"])  # ... must survive this
    dom.run_steps(1, nT - k)
    got = snapshot(dom)
    for name in STATES + FLUXES + RSTATES:
        parity.assert_bit_exact(got[name], want[name], "restart: " + name)
    parity.assert_bit_exact(got["Q"], np.ascontiguousarray(want["Q"][:, k:]), "restart: gauge series")
    # oracle, restarted the same way
    o = orc_run.OracleRun(sub, read_states=True)
    for s in RSTATES:
        o.R[s][...] = rs[s]
    o.run(1, nT - k)
    parity.assert_close(got["Q"], o.mRM_runoff, "restart vs oracle: discharge", rtol=parity.RTOL_Q)
    for name in STATES:
        parity.assert_close(got[name], o.S[name], "restart vs oracle: " + name)
    # without read_states the restarted run must differ (the test would be blind otherwise)
    clear(ctx)
    dom = driver.setup_domain(ctx, 3, sub, read_states=False)
    dom.run_steps(1, 24)
    assert not np.array_equal(dom.get_runoff()[:, :24], want["Q"][:, k:k + 24])


# --------------------------------------------------------------------------- packed domains
def pack(arrs):
    return np.ascontiguousarray(np.concatenate(arrs, axis=-1))


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_two_domains_in_global_arrays(ctx, mode):
    """DomainLoop: two domains of different size whose parameters, states, forcing, fluxes and
    routing states live in the reference's module-global arrays (leading dimension nCellsTot /
    nNodesTot, domain 2 at offset s1 - 1 = n1).  Every transfer takes (base, ld, offset); results
    equal the same domains set up from dense arrays; bind_host_* + sync_to_host after every step
    of the per-step seam equal get_*, and leave the other domain's section untouched."""
    p1 = synth.make_problem(nx=24, ny=15, n_days=3, hourly=True, seed=1)
    p2 = synth.make_problem(nx=31, ny=11, n_days=3, hourly=True, seed=2, soil_case=1)
    probs, n1, n2 = [p1, p2], p1["nCells"], p2["nCells"]
    nT, ntot = p1["time"]["nTimeSteps"], n1 + n2
    offs = [0, n1]
    nn = [p["net"]["nNodes"] for p in probs]
    noffs, nntot = [0, nn[0]], nn[0] + nn[1]
    G = {
        "params": {k: pack([p["params"][k] for p in probs]) for k in p1["params"] if k in interface.PARAM_NAMES},
        "states": {k: pack([p["states0"][k] for p in probs]) for k in STATES},
        "forcing": {k: pack([p["forcing"][k] for p in probs]) for k in ("pre", "temp", "pet")},
    }
    ctx.set_math_mode(mode)
    # dense reference runs
    want = []
    for p in probs:
        clear(ctx)
        d = driver.setup_domain(ctx, 1, p)
        d.run_steps(1, nT)
        want.append(snapshot(d))

    def setup_packed():
        clear(ctx)
        doms = []
        for i, p in enumerate(probs):
            d = ctx.register_domain(i + 1, p["nCells"], p["nH"], p["nLAI"], p["nLC"], p["processMatrix"],
                                    timestep_h=1)
            d.set_meteo_config(p["pet_case"], 24, True, False, synth.FNIGHT_PREC, synth.FNIGHT_PET,
                               synth.FNIGHT_TEMP, synth.EVAP_COEFF)
            d.set_time(p["time"])
            for k, a in G["params"].items():
                d.set_param(k, a, ld=ntot, offset=offs[i])
            for k, a in G["states"].items():
                d.set_state(k, a, ld=ntot, offset=offs[i])
            for k, a in G["forcing"].items():
                d.set_meteo(k, a, first_step=1, ld=ntot, offset=offs[i])
            net = p["net"]
            d.set_network(net)
            d.set_reg_rout(net["rout_param"], net["L11_length"][: net["nNodes"] - 1],
                           net["L11_slope"][: net["nNodes"] - 1], net["L11_nLinkFracFPimp"])
            doms.append(d)
        return doms

    # (1) block seam, results fetched into the global arrays with (ld, offset)
    doms = setup_packed()
    for d in doms:
        d.run_steps(1, nT)
    hostS = {k: np.full((p1["nH"], ntot) if k == "L1_soilMoist" else ntot, np.nan) for k in STATES}
    hostF = {k: np.full((p1["nH"], ntot) if k in ("L1_aETSoil", "L1_infilSoil") else ntot, np.nan) for k in FLUXES}
    hostR = {k: np.full((2, nntot) if k in ("L11_qTIN", "L11_qTR") else nntot, np.nan) for k in RSTATES}
    for i, d in enumerate(doms):
        for k in STATES:
            d.get_state(k, out=hostS[k], offset=offs[i])
        for k in FLUXES:
            d.get_flux(k, out=hostF[k], offset=offs[i])
        for k in RSTATES:
            d.get_routing_state(k, out=hostR[k], offset=noffs[i])
    for i, p in enumerate(probs):
        sl = slice(offs[i], offs[i] + p["nCells"])
        for k in STATES:
            parity.assert_bit_exact(hostS[k][..., sl], want[i][k], "packed domain %d: %s" % (i + 1, k))
        for k in FLUXES:
            parity.assert_bit_exact(hostF[k][..., sl], want[i][k], "packed domain %d: %s" % (i + 1, k))
        nsl = slice(noffs[i], noffs[i] + nn[i])
        for k in RSTATES:
            parity.assert_bit_exact(hostR[k][..., nsl], want[i][k], "packed domain %d: %s" % (i + 1, k))
        parity.assert_bit_exact(doms[i].get_runoff(), want[i]["Q"], "packed domain %d: discharge" % (i + 1))

    # (2) per-step seam with bound host globals: sync_to_host == get_*, other sections untouched
    doms = setup_packed()
    hostS = {k: np.full((p1["nH"], ntot) if k == "L1_soilMoist" else ntot, -7.0) for k in STATES}
    hostF = {k: np.full((p1["nH"], ntot) if k in ("L1_aETSoil", "L1_infilSoil") else ntot, -7.0) for k in FLUXES}
    for i, d in enumerate(doms):
        for k in STATES:
            d.bind_host_state(k, hostS[k], offset=offs[i])
        for k in FLUXES:
            d.bind_host_flux(k, hostF[k], offset=offs[i])
    idx = [interface.time_indices(p["time"], 1, 24, 1, nT) for p in probs]
    for tt in range(1, 31):
        for i, d in enumerate(doms):
            d.do_time_step(tt, idx[i][tt - 1])
            d.route(tt, idx[i][tt - 1].yId, 1, 1.0)
        doms[0].sync_to_host()
        if tt == 1:  # domain 2 has not been synchronised yet: its section still holds the fill value
            assert (hostS["L1_satSTW"][n1:] == -7.0).all() and (hostF["L1_total_runoff"][n1:] == -7.0).all()
        doms[1].sync_to_host()
        for i, d in enumerate(doms):
            sl = slice(offs[i], offs[i] + probs[i]["nCells"])
            for k in STATES:
                parity.assert_bit_exact(hostS[k][..., sl], d.get_state(k), "sync_to_host tt=%d %s" % (tt, k))
            for k in FLUXES:
                parity.assert_bit_exact(hostF[k][..., sl], d.get_flux(k), "sync_to_host tt=%d %s" % (tt, k))
    # the per-step seam reproduces the block seam's first 30 steps
    # (strict: bit for bit; the fast block seam multiplies the node runoff by a hoisted 1000 / TST
    # where the per-step seam's unfused L11_runoff_acc divides -- 1 ulp)
    for i, d in enumerate(doms):
        if mode == "strict":
            parity.assert_bit_exact(d.get_runoff()[:, :30], want[i]["Q"][:, :30], "per-step seam discharge, domain %d" % (i + 1))
        else:
            parity.assert_close(d.get_runoff()[:, :30], want[i]["Q"][:, :30], "per-step seam discharge, domain %d" % (i + 1),
                                rtol=1e-13)


def test_bad_section_arguments_are_rejected(ctx):
    prob = synth.make_problem(nx=10, ny=8, n_days=1, routing=False)
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, prob)
    n = prob["nCells"]
    a = np.zeros(n)
    L, h = ctx.L, ctx.h
    pd = a.ctypes.data_as(C.POINTER(C.c_double))
    assert L.mhm_cuda_set_state(h, 1, 0, 0, pd, n - 1, 0) != 0       # ld < nCells
    assert L.mhm_cuda_set_state(h, 1, 0, 0, pd, n, -1) != 0          # negative offset
    assert L.mhm_cuda_get_flux(h, 1, 1, 0, pd, n, 0) != 0            # member out of range
    assert L.mhm_cuda_bind_host_state(h, 1, 99, pd, n, 0) != 0       # unknown state id
    assert b"bind_host_state" in L.mhm_cuda_last_error()
    assert L.mhm_cuda_sync_to_host(h, 7) != 0                        # unknown domain
    del dom


# --------------------------------------------------------------------------- re-evaluation
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_reevaluation_with_another_celerity_changes_tsrout(ctx, mode):
    """A calibration loop calls mrm_update_param at every evaluation (mRM/mo_mrm_mpr.f90:294-321):
    a new celerity gives new C1 / C2 and may give another adaptive routing step L11_TSrout.
    set_c1c2 must accept that on the same domain (round-1 defect: 'members must share
    L11_TSrout'); members that disagree are still rejected when a block is routed."""
    base = synth.make_problem(nx=30, ny=20, n_days=6, hourly=False, rout_case=2, start=(1990, 6, 1))
    orc_run.case23_params(base["net"])
    nT = base["time"]["nTimeSteps"]
    ctx.set_math_mode(mode)
    clear(ctx)
    dom = driver.setup_domain(ctx, 1, base, nMembers=2)
    seen = []
    for cel in (0.35, 2.5):
        prob = copy.copy(base)
        prob["net"] = dict(base["net"], celerity=cel)
        orc_run.case23_params(prob["net"])
        seen.append(prob["net"]["TSrout"])
        net = prob["net"]
        for m in range(2):  # a new evaluation: states back to the start, new routing parameters
            for k, a in prob["states0"].items():
                dom.set_state(k, a, member=m)
            for k in ("L11_qOUT", "L11_qMod"):
                dom.set_routing_state(k, np.zeros(net["nNodes"]), member=m)
            for k in ("L11_qTIN", "L11_qTR"):
                dom.set_routing_state(k, np.zeros((2, net["nNodes"])), member=m)
            dom.set_c1c2(net["C1"], net["C2"], net["TSrout"], member=m)
        dom.run_steps(1, nT)
        o = orc_run.OracleRun(prob)
        o.run(1, nT)
        for m in range(2):
            parity.assert_close(dom.get_runoff(member=m), o.mRM_runoff, "celerity %g, member %d" % (cel, m),
                                rtol=parity.RTOL_Q)
    assert seen[0] != seen[1], "the two celerities must give different routing steps: %r" % (seen,)
    # members that disagree about the routing step cannot be routed on one schedule
    dom.set_c1c2(net["C1"], net["C2"], seen[0], member=0)
    rc = ctx.L.mhm_cuda_run_steps(ctx.h, 1, 1, 24)
    assert rc != 0 and b"share L11_TSrout" in ctx.L.mhm_cuda_last_error()


# --------------------------------------------------------------------------- forcing through TMA
@pytest.mark.parametrize("nx,ny", [(52, 31), (45, 29), (64, 2)])
def test_forcing_through_tma_equals_per_lane_loads(ctx, monkeypatch, nx, ny):
    """Uniform-calendar launches with the forcing rows staged by the TMA unit (cp.async.bulk into a
    shared-memory ring, one mbarrier per warp and row; MHM_CUDA_FORCING_TMA=1) against the per-lane
    __ldg path: bit-identical states, fluxes, gauge series; domains whose last tile / last warp is
    ragged, a domain with an odd number of cells (rows are not 16-byte aligned: falls back), blocks
    that start inside the resident chunk, three members."""
    prob = synth.make_problem(nx=nx, ny=ny, n_days=7, hourly=True, start=(1990, 12, 28))
    nT, M = prob["time"]["nTimeSteps"], 3
    mp = [prob["params"]] + [dict(prob["params"], L1_kPerco=prob["params"]["L1_kPerco"] * f) for f in (0.9, 1.1)]
    ctx.set_math_mode("fast")
    res = {}
    for key in ("ldg", "tma"):
        monkeypatch.setenv("MHM_CUDA_FORCING_TMA", "1" if key == "tma" else "0")
        clear(ctx)
        dom = driver.setup_domain(ctx, 1, prob, nMembers=M, member_params=mp)
        for a, b in ((1, 50), (51, 3), (54, nT - 53)):
            dom.run_steps(a, b)
        res[key] = [snapshot_member(dom, m) for m in range(M)]
    monkeypatch.delenv("MHM_CUDA_FORCING_TMA")
    for m in range(M):
        assert_same(res["tma"][m], res["ldg"][m], "TMA forcing, member %d (%d cells)" % (m, prob["nCells"]))
    o = orc_run.OracleRun(prob)
    o.run(1, nT)
    parity.assert_close(res["tma"][0]["Q"], o.mRM_runoff, "TMA forcing vs oracle", rtol=parity.RTOL_Q)
    ctx.set_math_mode("strict")
