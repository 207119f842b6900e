"""Load a problem dict (mhm_b200.synth.make_problem layout) into a Context: the sequence of
C-ABI calls the reference's driver performs after mhm_initialize / mrm_init / mpr_eval."""
import numpy as np

from . import synth
from .interface import PARAM_NAMES

METEO_BY_CASE = {-1: ["pet"], 0: ["pet"], 1: ["tmin", "tmax"], 2: ["netrad"],
                 3: ["netrad", "absvappress", "windspeed"]}


def setup_domain(ctx, iDomain, prob, nMembers=1, member_params=None, upload_forcing=True,
                 read_states=False):
    n, nH, nLAI, nLC = prob["nCells"], prob["nH"], prob["nLAI"], prob["nLC"]
    dom = ctx.register_domain(iDomain, n, nH, nLAI, nLC, prob["processMatrix"],
                              timestep_h=prob["timestep_h"], read_states=read_states, nMembers=nMembers)
    dom.set_meteo_config(prob["pet_case"], prob["nTstepForcingDay"], prob["hourly"],
                         prob["read_weights"], synth.FNIGHT_PREC, synth.FNIGHT_PET,
                         synth.FNIGHT_TEMP, synth.EVAP_COEFF)
    dom.set_time(prob["time"])
    for m in range(nMembers):
        P = prob["params"] if member_params is None else member_params[m]
        for name, arr in P.items():
            if name in PARAM_NAMES:  # e.g. "rout_param" travels with the member but is not an L1 field
                dom.set_param(name, arr, member=m)
        for name, arr in prob["states0"].items():
            dom.set_state(name, arr, member=m)
    if upload_forcing:
        for var in ["pre", "temp"] + METEO_BY_CASE[prob["pet_case"]]:
            dom.set_meteo(var, prob["forcing"][var], first_step=1)
    if prob["read_weights"]:
        for var in ("pre", "temp", "pet"):
            dom.set_meteo_weights(var, prob["weights"][var])
    net = prob.get("net")
    if net is not None:
        dom.set_network(net)
        for m in range(nMembers):
            if net["processCase"] == 1:
                rp = net["rout_param"] if member_params is None else member_params[m].get(
                    "rout_param", net["rout_param"])
                dom.set_reg_rout(rp, net["L11_length"][: net["nNodes"] - 1],
                                 net["L11_slope"][: net["nNodes"] - 1], net["L11_nLinkFracFPimp"],
                                 member=m)
            else:
                dom.set_c1c2(net["C1"], net["C2"], net["TSrout"], member=m)
        if net["nInflowTotal"] > 0:
            dom.set_inflow(prob["inflowQ"])
    return dom
