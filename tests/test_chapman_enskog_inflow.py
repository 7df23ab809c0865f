"""uniGasChapmanEnskogFreeStreamInflowPatch (U/boundaries/derived/generalBoundaries/uniGasChapmanEnskogFreeStreamInflowPatch/
uniGasChapmanEnskogFreeStreamInflowPatch.C:40-150; count and sampler in U/boundaries/basic/uniGasGeneralBoundary/
uniGasGeneralBoundary.C:171-239, 880-940): free-stream insertion from the Chapman-Enskog distribution
f = f_M (1 + Gamma(C; q, tau)), the first-order correction carrying a prescribed heat flux q and shear stress tau.

  * q = 0, tau = 0: the count is Bird 4.22 and the inserted velocities are the inflowing half-range Maxwellian;
  * a heat flux along the inward normal: the number flux changes by the closed-form factor of the count formula, and the
    inserted parcels carry the energy flux excess q_n . A . dt (the defining moment of the distribution);
  * GPU: lockstep with the oracle (same counts each step; velocities equal to libm round-off).
"""
import copy
import math

import numpy as np
import pytest

from unigasfoam_b200 import cases, mesh as ugmesh
from unigasfoam_b200.cloud import UgfError

kB = cases.kB


def ce_channel(q=(0.0, 0.0, 0.0), stress=None, U=0.0, n_inf=4.247e20, T_inf=300.0, ppc=400, nx=6, ny=4, seed=3, vacuum=True):
    """A short channel whose xMin patch inserts the Chapman-Enskog stream (inward normal +x), deleting outlet at xMax."""
    sp = cases.ARGON_TUTORIAL
    kinds = {"xMin": ("inlet", "patch"), "xMax": ("outlet", "patch"), "yMin": ("bottom", "symmetryPlane"), "yMax": ("top", "symmetryPlane"),
             "zMin": ("back", "empty"), "zMax": ("front", "empty")}
    L = 0.05
    m = ugmesh.box_mesh(nx, ny, 1, L, 0.6 * L, 0.1 * L, kinds, solution_d=(1, 1, 0))
    m.meta_axis_aligned = True
    FN = n_inf * (L * 0.6 * L * 0.1 * L) / (ppc * nx * ny)
    props = cases._props("Ar", sp, FN, binary="noDSMCCollision", Tref=1000.0)
    inflow = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasChapmanEnskogFreeStreamInflowPatch",
              "uniGasChapmanEnskogFreeStreamInflowPatchProperties": {
                  "typeIds": ["Ar"], "numberDensities": {"Ar": n_inf}, "translationalTemperature": T_inf, "velocity": [U, 0.0, 0.0],
                  "heatFlux": list(q), "stress": list(np.zeros(9) if stress is None else np.asarray(stress, float).reshape(9))}}
    outflow = {"generalBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
               "uniGasFreeStreamInflowPatchProperties": {"typeIds": ["Ar"], "numberDensities": {"Ar": 0.0}, "translationalTemperature": T_inf,
                                                         "velocity": [0.0, 0.0, 0.0]}}
    cmp_ = cases.most_probable_speed(T_inf, sp["mass"])
    dt = 0.2 * (L / nx) / (abs(U) + cmp_)
    meta = dict(sp=sp, L=L, A=0.6 * L * 0.1 * L, FN=FN, cmp=cmp_, n=n_inf, T=T_inf, U=U, p=n_inf * kB * T_inf, dt=dt)
    meta["seed"] = seed
    return m, props, {"uniGasGeneralBoundaries": [inflow, outflow]}, dt, meta


def make(Cloud, *a, **kw):
    m, props, bd, dt, meta = ce_channel(*a, **kw)
    cl = Cloud(m, props, bd, dt, parcelCapacity=200_000, seed=meta["seed"])
    cl.setParcels(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, np.int32))
    return cl, meta


def inserted_stream(cl, steps):
    """The velocities of all parcels the inlet inserts in `steps` steps (insertion alone, before any move)."""
    out, count = [], 0
    for _ in range(steps):
        n0 = cl.size()
        cl.controlBeforeMove()
        q = cl.parcels()
        out.append(q["U"][n0:].copy())
        count += len(q["U"]) - n0
        cl.move(); cl.finishStep()
    return np.concatenate(out), count


def flux_factor(s):
    return (math.exp(-s * s) + math.sqrt(math.pi) * s * (1 + math.erf(s))) / (2 * math.sqrt(math.pi))


@pytest.mark.parametrize("U", [0.0, 150.0])
def test_oracle_zero_heat_flux_and_stress_is_the_maxwellian_stream(OracleCloud, U):
    cl, me = make(OracleCloud, U=U)
    vel, count = inserted_stream(cl, 60)
    expect = 60 * me["A"] * me["n"] * me["dt"] * me["cmp"] * flux_factor(U / me["cmp"]) / me["FN"]
    assert expect > 2000 and abs(count - expect) < 4 * math.sqrt(expect)
    assert (vel[:, 0] > 0).all()  # every inserted parcel moves into the domain
    m_ = me["sp"]["mass"]
    kTm = kB * me["T"] / m_
    # tangential components: the full Maxwellian; normal: the flux-weighted half range, <u_n^2> etc from Bird 4.22's integrals
    assert abs(vel[:, 1].mean()) < 4 * math.sqrt(kTm / count) and abs((vel[:, 1] ** 2).mean() / kTm - 1) < 0.05
    s = U / me["cmp"]
    # <u_n> of the inflowing stream = (number-flux-weighted) int u^2 f / int u f, evaluated numerically
    u = np.linspace(0, 8 * me["cmp"] + abs(U), 400_001)
    f = np.exp(-((u - U) / me["cmp"]) ** 2)
    un_mean = np.trapezoid(u * u * f, u) / np.trapezoid(u * f, u)
    assert abs(vel[:, 0].mean() / un_mean - 1) < 0.02, s


def test_oracle_heat_flux_changes_count_and_energy_flux(OracleCloud):
    """q along the inward normal, gas at rest: count factor 1 (s = 0 kills the q term), but the inserted stream carries
    the extra energy flux.  For the Chapman-Enskog distribution the one-sided energy flux through a plane at rest is
    the Maxwellian one plus q_n / 2 (the odd correction term contributes half of its full-range moment on each side)."""
    me0 = make(OracleCloud)[1]
    qn = 0.15 * me0["p"] * me0["cmp"]  # breakdown 0.3
    m_ = me0["sp"]["mass"]
    steps = 80
    res = {}
    for name, q in (("0", 0.0), ("+", qn), ("-", -qn)):
        cl, me = make(OracleCloud, q=(q, 0.0, 0.0), seed=11)
        vel, count = inserted_stream(cl, steps)
        res[name] = (count, 0.5 * m_ * (vel ** 2).sum() * me["FN"] / (steps * me["dt"] * me["A"]))  # W / m^2 carried in
    expect = steps * me0["A"] * me0["n"] * me0["dt"] * me0["cmp"] * flux_factor(0.0) / me0["FN"]
    for name in res:
        assert abs(res[name][0] - expect) < 4 * math.sqrt(expect), name
    e0 = me0["n"] * kB * me0["T"] * me0["cmp"] / math.sqrt(math.pi)  # Maxwellian one-sided energy flux n k T c_mp / sqrt(pi)
    assert abs(res["0"][1] / e0 - 1) < 0.03
    assert abs((res["+"][1] - res["-"][1]) / qn - 1.0) < 0.1  # (+q/2) - (-q/2)
    assert res["+"][1] > res["0"][1] > res["-"][1]


def test_oracle_normal_stress_changes_the_count(OracleCloud):
    me0 = make(OracleCloud)[1]
    tau = 0.2 * me0["p"]
    S = np.diag([tau, -0.5 * tau, -0.5 * tau])  # traceless, normal component along the inlet normal
    cl, me = make(OracleCloud, stress=S, U=100.0)
    _, count = inserted_stream(cl, 60)
    s = 100.0 / me["cmp"]
    fac = (math.exp(-s * s) * (1 - 0.5 * tau / me["p"]) + math.sqrt(math.pi) * s * (1 + math.erf(s))) / (2 * math.sqrt(math.pi))
    expect = 60 * me["A"] * me["n"] * me["dt"] * me["cmp"] * fac / me["FN"]
    assert abs(count - expect) < 4 * math.sqrt(expect)
    plain = 60 * me["A"] * me["n"] * me["dt"] * me["cmp"] * flux_factor(s) / me["FN"]
    assert abs(count - plain) > 4 * math.sqrt(plain)  # the correction is resolved by the sample


def test_shear_stress_tilts_the_inserted_stream(OracleCloud):
    """tau_xy > 0 in the correction -2 tau_xy C_x C_y / p: inflowing parcels (C_x > 0) are biased towards C_y < 0."""
    me0 = make(OracleCloud)[1]
    S = np.zeros((3, 3)); S[0, 1] = S[1, 0] = 0.25 * me0["p"]
    cl, me = make(OracleCloud, stress=S)
    vel, count = inserted_stream(cl, 60)
    sig = math.sqrt(kB * me["T"] / me["sp"]["mass"] / count)
    assert vel[:, 1].mean() < -8 * sig


def test_dictionary_needs_heat_flux_and_stress(OracleCloud):
    m, props, bd, dt, _ = ce_channel()
    bad = copy.deepcopy(bd)
    del bad["uniGasGeneralBoundaries"][0]["uniGasChapmanEnskogFreeStreamInflowPatchProperties"]["heatFlux"]
    with pytest.raises(KeyError):
        OracleCloud(m, props, bad, dt, parcelCapacity=1000)  # the capacity makes the constructor build the cloud at once
    bad = copy.deepcopy(bd)
    bad["uniGasGeneralBoundaries"][0]["uniGasChapmanEnskogFreeStreamInflowPatchProperties"]["stress"] = [0.0] * 6
    with pytest.raises((ValueError, UgfError)):
        OracleCloud(m, props, bad, dt, parcelCapacity=1000)


@pytest.mark.gpu
def test_gpu_chapman_enskog_inflow_in_lockstep_with_oracle(GpuCloud, OracleCloud):
    me0 = ce_channel()[4]
    S = np.zeros((3, 3)); S[0, 1] = S[1, 0] = 0.1 * me0["p"]; S[0, 0] = 0.1 * me0["p"]; S[1, 1] = -0.1 * me0["p"]
    q = (0.1 * me0["p"] * me0["cmp"], -0.05 * me0["p"] * me0["cmp"], 0.0)
    g, _ = make(GpuCloud, q=q, stress=S, U=120.0, seed=21)
    r, _ = make(OracleCloud, q=q, stress=S, U=120.0, seed=21)
    tot = 0
    for _ in range(12):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"] and cg["nParcels"] == cr["nParcels"]
        tot += cr["inserted"]
    assert tot > 500
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() > 0.99
    assert (np.abs(pg["position"] - pr["position"]) <= 1e-9 * np.abs(pr["position"]).max()).all(1).mean() > 0.99
