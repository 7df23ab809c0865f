"""The CPU oracle against closed-form kinetic theory.

The reference ships no tests or golden vectors for this path ("parity unpinned"), so the restatement is
pinned by known answers instead: conservation per collision, the equilibrium VHS collision rate, the NTC
candidate formula, BGK relaxation rate and conservation, the Bird 4.22 inflow flux, wall pressure / heat flux.
"""
import math

import numpy as np
import pytest

from unigasfoam_b200 import cases

kB = cases.kB


def cell_sums(p, m, nC, erot=False):
    mom = np.stack([np.bincount(p["cell"], m * p["U"][:, k], nC) for k in range(3)], axis=1)
    e = 0.5 * m * (p["U"] ** 2).sum(1) + (p["ERot"] if erot else 0.0)
    return mom, np.bincount(p["cell"], e, nC)


@pytest.mark.parametrize("binary,species", [
    ("variableHardSphere", ("Ar", cases.ARGON_GUIDE)),
    ("variableSoftSphere", ("Ar", dict(cases.ARGON_GUIDE, alpha=1.4))),
    ("LarsenBorgnakkeVariableHardSphere", ("N2", cases.NITROGEN)),
    ("LarsenBorgnakkeVariableSoftSphere", ("N2", dict(cases.NITROGEN, alpha=1.36))),
])
def test_binary_collisions_conserve_momentum_and_energy(OracleCloud, binary, species):
    case = cases.closed_box(n=6, parcels=15000, seed=21, binary=binary, species=species, dt_mct=1.0, Trot=200.0)
    cl = case.make_cloud(OracleCloud)
    m = species[1]["mass"]
    rot = species[1]["rotationalDegreesOfFreedom"] > 0
    cl.buildCellOccupancy(); cl.reorder()
    before = cl.parcels()
    cl.collide()
    after = cl.parcels()
    nC = case.mesh.n_cells
    pb, eb = cell_sums(before, m, nC, rot)
    pa, ea = cell_sums(after, m, nC, rot)
    scale = np.bincount(before["cell"], m * np.abs(before["U"]).sum(1), nC)[:, None]
    assert (np.abs(pa - pb) <= 1e-12 * scale).all()
    assert (np.abs(ea - eb) <= 1e-12 * eb).all()
    assert cl.counters()["collisions"] > 500
    if rot:
        assert (after["ERot"] != before["ERot"]).sum() > 100


def test_equilibrium_collision_rate_and_candidates(OracleCloud):
    """Accepted collisions per step = 1/2 N nu dt with the VHS equilibrium rate (Bird 4.64); NTC candidates of
    the first step = sum over cells of 1/2 N (N-1) F_N (sigma cR)max dt / V (noTimeCounter.C:184)."""
    case = cases.closed_box(n=8, parcels=40000, seed=22)
    cl = case.make_cloud(OracleCloud)
    N = case.n_parcels
    cnt = np.bincount(case.cell, minlength=case.mesh.n_cells)
    FN = case.uniGasProperties["nEquivalentParticles"]
    expect_cand = (0.5 * cnt * (cnt - 1) * FN * case.sigmaTcRMax * case.deltaT / case.mesh.cell_volumes).sum()
    cl.evolve(1)
    c = cl.counters()
    assert abs(c["collisionCandidates"] - expect_cand) < 4 * math.sqrt(case.mesh.n_cells * 0.25) + 1e-9 * expect_cand
    coll = []
    for _ in range(40):
        cl.evolve(1)
        coll.append(cl.counters()["collisions"])
    sp = case.meta["species"]
    nu = cases.vhs_collision_rate(case.meta["n"], case.meta["T0"], sp, case.meta["Tref"])
    expect = 0.5 * N * nu * case.deltaT
    # (per cell the rate goes as N(N-1), the formula as <N>^2: equal once the cell counts are Poisson, after warm-up)
    mean = np.mean(coll[10:])
    sigma = math.sqrt(expect / len(coll[10:]))
    assert abs(mean - expect) < 4 * sigma + 0.01 * expect, (mean, expect)


def test_subcells_pair_neighbours_within_the_cell(OracleCloud):
    """subCellLevels (2,2,2): a candidate's partner comes from its own virtual sub-cell whenever that holds another
    parcel (noTimeCounter.C:205-235), so colliding pairs sit closer together than with whole-cell selection;
    the candidate count (noTimeCounter.C:184) does not depend on the sub-cells."""
    res = {}
    for levels in (1, 2):
        case = cases.closed_box(n=5, parcels=12000, seed=31, dt_mct=0.1)
        case.subCellLevels = np.full((case.mesh.n_cells, 3), levels, np.int32)
        cl = case.make_cloud(OracleCloud)
        cl.buildCellOccupancy(); cl.reorder()
        before = cl.parcels()
        cl.collide()
        after = cl.parcels()
        c = cl.counters()
        hit = np.flatnonzero((after["U"] != before["U"]).any(axis=1))
        # partners share a cell: for every collided parcel, distance to the nearest other collided parcel of its cell
        lo, hi = case.mesh.cell_bb_min, case.mesh.cell_bb_max
        rel = (after["position"][hit] - lo[after["cell"][hit]]) / (hi - lo)[after["cell"][hit]]
        octant = (rel >= 0.5).astype(int) @ np.array([1, 2, 4])
        key = after["cell"][hit] * 8 + octant
        _, counts = np.unique(key, return_counts=True)
        res[levels] = (c["collisionCandidates"], c["collisions"], (counts % 2 == 0).mean())
        cl.close()
    assert res[1][0] == res[2][0]
    assert abs(res[1][1] - res[2][1]) < 0.1 * res[1][1]
    # with sub-cells nearly every (cell, octant) group holds an even number of collided parcels (whole pairs)
    assert res[2][2] > 0.8 > res[1][2]


def test_ntc_subcycled_keeps_the_equilibrium_collision_rate(OracleCloud):
    """noTimeCounterSubCycled with nSubCycles = 4: four passes at deltaT/4 give the same expected number of accepted
    collisions per step as one noTimeCounter pass (Bird 4.64: 1/2 N nu deltaT per cell)."""
    rates = {}
    for partner in ("noTimeCounter", "noTimeCounterSubCycled"):
        case = cases.closed_box(n=6, parcels=30000, seed=17, dt_mct=0.5, nSubCycles=4)
        case.uniGasProperties["dsmcCollisionPartnerModel"] = partner
        cl = case.make_cloud(OracleCloud)
        cl.evolve(10)  # let sigmaTcRMax settle
        n = 0
        for _ in range(10):
            cl.evolve(1)
            n += cl.counters()["collisions"]
        rates[partner] = n / 10
        cl.close()
    assert abs(rates["noTimeCounter"] - rates["noTimeCounterSubCycled"]) < 0.04 * rates["noTimeCounter"]


def test_specular_box_conserves_energy_and_maxwellian(OracleCloud):
    case = cases.closed_box(n=6, parcels=30000, seed=23)
    cl = case.make_cloud(OracleCloud)
    e0 = cl.counters()["linearKineticEnergy"]
    cl.evolve(30)
    c = cl.counters()
    assert abs(c["linearKineticEnergy"] - e0) <= 1e-12 * e0
    assert c["nParcels"] == case.n_parcels and c["stuck"] == 0
    U = cl.parcels()["U"]
    m = case.meta["species"]["mass"]
    T = m * (U ** 2).mean() / kB
    assert abs(T - case.meta["T0"]) < 0.02 * case.meta["T0"]
    kurt = ((U - U.mean(0)) ** 4).mean(0) / (U.var(0) ** 2)
    assert np.all(np.abs(kurt - 3.0) < 0.12)  # stays Gaussian
    f = cl.fields()
    assert abs(f["translationalT"].mean() - case.meta["T0"]) < 0.02 * case.meta["T0"]
    assert abs(f["rhoN"].mean() - case.meta["n"]) < 1e-3 * case.meta["n"]
    np.testing.assert_allclose(f["p"], f["rhoN"] * kB * f["translationalT"], rtol=1e-12)


def test_diffuse_wall_pressure_and_zero_net_heat_flux(OracleCloud):
    """Gas at rest at the wall temperature: wall pressure = n k T, net heat flux ~ 0 (uniGasPatchBoundary.C:213-302)."""
    case = cases.closed_box(n=4, parcels=40000, seed=24, wall="diffuse", binary="noDSMCCollision", dt_mct=0.5)
    cl = case.make_cloud(OracleCloud)
    cl.evolve(60)
    f = cl.fields()
    walls = f["wall_p"] != 0
    p_expect = case.meta["n"] * kB * case.meta["T0"]
    assert abs(f["wall_p"][walls].mean() - p_expect) < 0.03 * p_expect
    qscale = p_expect * cases.most_probable_speed(case.meta["T0"], case.meta["species"]["mass"])
    assert abs(f["surfaceHeatTransfer"][walls].mean()) < 0.02 * qscale
    assert f["surfaceShearStress"][walls].mean() < 0.1 * p_expect


def _with_cll_walls(case, alphaN, sigmaT, alphaR=1.0, T=None, U=(0.0, 0.0, 0.0)):
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        old = e.get(e["boundaryModel"] + "Properties", {})
        e["boundaryModel"] = "uniGasCLLWallPatch"
        e["uniGasCLLWallPatchProperties"] = {"temperature": old.get("temperature", T), "velocity": list(old.get("velocity", U)),
                                            "normalAccommCoeff": alphaN, "tangentialAccommCoeff": sigmaT, "rotEnergyAccommCoeff": alphaR}
    return case


def test_cll_wall_limits(OracleCloud):
    """uniGasCLLWallPatch (uniGasCLLWallPatch.C:80-254).  Zero accommodation is a specular wall: kinetic energy is
    conserved and |U.n| is mirrored.  Full accommodation keeps a gas at the wall temperature in equilibrium: wall
    pressure n k T, no net heat flux (the CLL kernel satisfies detailed balance)."""
    case = _with_cll_walls(cases.closed_box(n=4, parcels=20000, seed=26, wall="diffuse", binary="noDSMCCollision", dt_mct=0.5), 0.0, 0.0)
    cl = case.make_cloud(OracleCloud)
    e0 = cl.counters()["linearKineticEnergy"]
    cl.evolve(20)
    c = cl.counters()
    assert c["wallHits"] > 1000
    assert abs(c["linearKineticEnergy"] - e0) < 1e-9 * e0
    cl.close()
    case = _with_cll_walls(cases.closed_box(n=4, parcels=40000, seed=27, wall="diffuse", binary="noDSMCCollision", dt_mct=0.5), 1.0, 1.0)
    cl = case.make_cloud(OracleCloud)
    cl.evolve(60)
    f = cl.fields()
    walls = f["wall_p"] != 0
    p_expect = case.meta["n"] * kB * case.meta["T0"]
    assert abs(f["wall_p"][walls].mean() - p_expect) < 0.03 * p_expect
    qscale = p_expect * cases.most_probable_speed(case.meta["T0"], case.meta["species"]["mass"])
    assert abs(f["surfaceHeatTransfer"][walls].mean()) < 0.02 * qscale
    T = (2.0 / 3.0) * cl.counters()["linearKineticEnergy"] / (cl.size() * kB)
    assert abs(T - case.meta["T0"]) < 0.02 * case.meta["T0"]
    cl.close()


@pytest.mark.parametrize("bgk", ["stochasticParticleBGK", "stochasticParticleESBGK", "stochasticParticleSBGK", "unifiedStochasticParticleSBGK"])
def test_bgk_conserves_cell_momentum_and_energy(OracleCloud, bgk):
    case = cases.closed_box(n=5, parcels=12000, seed=25, mode="bgk", bgk=bgk, binary="noDSMCCollision", dt_mct=2.0,
                            velocity=(300.0, -80.0, 20.0))
    case.U[:, 1] *= 1.4
    cl = case.make_cloud(OracleCloud)
    m = case.meta["species"]["mass"]
    cl.buildCellOccupancy(); cl.reorder(); cl.calculateFields()
    before = cl.parcels()
    cl.relax()
    after = cl.parcels()
    nC = case.mesh.n_cells
    pb, eb = cell_sums(before, m, nC)
    pa, ea = cell_sums(after, m, nC)
    scale = np.bincount(before["cell"], m * np.abs(before["U"]).sum(1), nC)[:, None]
    assert (np.abs(pa - pb) <= 1e-11 * scale).all()
    assert (np.abs(ea - eb) <= 1e-11 * eb).all()
    assert cl.counters()["bgkRelaxations"] > 0.5 * case.n_parcels  # dt = 2 MCT: most parcels relax


def test_bgk_relaxes_anisotropy_at_the_model_rate(OracleCloud):
    """BGK: T_x - T decays as exp(-nu t), nu = p/mu with the VHS viscosity law (…BGK.C:545-575)."""
    sp = cases.ARGON_GUIDE
    case = cases.closed_box(n=3, parcels=54000, seed=26, mode="bgk", bgk="stochasticParticleBGK", binary="noDSMCCollision", dt_mct=0.1)
    case.U[:, 0] *= 1.5
    cl = case.make_cloud(OracleCloud)
    m = sp["mass"]
    U0 = case.U
    T = m * (U0 ** 2).mean() / kB
    Tx0 = m * (U0[:, 0] ** 2).mean() / kB
    Tref = case.meta["Tref"]
    mu_ref = 1.25 * 2 * 3 * math.sqrt(m * kB * Tref) / (1 * (5 - 2 * sp["omega"]) * (7 - 2 * sp["omega"]) * math.sqrt(math.pi) * sp["diameter"] ** 2)
    mu = mu_ref * (T / Tref) ** sp["omega"]
    nu = case.meta["n"] * kB * T / mu
    steps = 12
    cl.evolve(steps)
    U = cl.parcels()["U"]
    Tx = m * (U[:, 0] ** 2).mean() / kB
    expect = T + (Tx0 - T) * math.exp(-nu * steps * case.deltaT)
    assert abs(m * (U ** 2).mean() / kB - T) < 1e-9 * T  # energy conserved
    assert abs((Tx - T) - (expect - T)) < 0.08 * abs(Tx0 - T), (Tx, expect, T)


def test_inflow_flux_matches_bird_4_22(OracleCloud):
    """Mean inserted parcels per step = A n dt c_mp [exp(-s^2) + sqrt(pi) s (1 + erf s)] / (2 sqrt(pi) F_N)."""
    from unigasfoam_b200 import mesh as ugmesh
    sp = cases.ARGON_TUTORIAL
    kinds = {"xMin": ("inlet", "patch"), "xMax": ("outlet", "patch"), "yMin": ("bottom", "symmetryPlane"), "yMax": ("top", "symmetryPlane"),
             "zMin": ("back", "empty"), "zMax": ("front", "empty")}
    L = 0.05
    m = ugmesh.box_mesh(10, 6, 1, L, 0.6 * L, 0.1 * L, kinds, solution_d=(1, 1, 0))
    m.meta_axis_aligned = True
    n_inf, T_inf, Uinf = 4.247e20, 200.0, np.array([800.0, 0.0, 0.0])
    FN = n_inf * (L * 0.6 * L * 0.1 * L) / 6000
    props = cases._props("Ar", sp, FN, binary="noDSMCCollision", Tref=1000.0)
    inflow = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
              "uniGasFreeStreamInflowPatchProperties": {"typeIds": ["Ar"], "numberDensities": {"Ar": n_inf}, "translationalTemperature": T_inf,
                                                        "velocity": list(Uinf)}}
    outflow = dict(inflow, generalBoundaryProperties={"patch": "outlet"})
    dt = 0.3 * (L / 10) / (Uinf[0] + cases.most_probable_speed(T_inf, sp["mass"]))
    cl = OracleCloud(m, props, {"uniGasGeneralBoundaries": [inflow, outflow]}, dt, parcelCapacity=40000)
    rng = np.random.default_rng(5)
    pos, vel, cel, tid, _ = cases.mesh_fill(m, {"Ar": sp}, ["Ar"], {"Ar": n_inf}, T_inf, Uinf, FN, rng)
    cl.setParcels(pos, vel, cel)
    cl.setCellState(sigmaTcRMax=1e-16)
    cmp_ = cases.most_probable_speed(T_inf, sp["mass"])

    def flux(s):
        return (math.exp(-s * s) + math.sqrt(math.pi) * s * (1 + math.erf(s))) / (2 * math.sqrt(math.pi))

    A = 0.6 * L * 0.1 * L
    expect = A * n_inf * dt * cmp_ * (flux(Uinf[0] / cmp_) + flux(-Uinf[0] / cmp_)) / FN  # inlet + back-flux through the outlet
    ins, n_hist = [], []
    for _ in range(150):
        cl.evolve(1)
        c = cl.counters()
        ins.append(c["inserted"]); n_hist.append(c["nParcels"])
    mean = np.mean(ins)
    assert abs(mean - expect) < 4 * math.sqrt(expect / len(ins)) + 0.005 * expect, (mean, expect)
    # steady free stream: the population neither drains nor piles up, and the stream keeps its velocity
    assert abs(np.mean(n_hist[-50:]) - len(cel)) < 0.03 * len(cel)
    U = cl.parcels()["U"]
    assert abs(U[:, 0].mean() - Uinf[0]) < 0.02 * Uinf[0]
    assert cl.counters()["stuck"] == 0


def _hot_top_hybrid_couette(**kw):
    case = cases.couette(nx=16, ny=12, ppc=60, Kn=0.2, mode="hybrid", bgk="unifiedStochasticParticleSBGK", theta=0.1, **kw)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        if e["patchBoundaryProperties"]["patch"] == "top":
            e["uniGasDiffuseWallPatchProperties"]["temperature"] = 900.0
    return case


def test_local_knudsen_matches_the_formulas(OracleCloud):
    """localKnudsen::decompose with theta 1 and no smoothing against an independent numpy evaluation of
    localKnudsen.C:294-383 from the time-averaged fields of the same window: KnX = lambda max_nb |X_nb - X| / (d X),
    lambda from Bird 4.76/4.77, KnGLL = max of the three, mask = KnGLL > breakdownMax before the refinement."""
    case = _hot_top_hybrid_couette()
    cl = case.make_cloud(OracleCloud)
    cl.setHybridDecomposition({"decompositionModel": "localKnudsen", "timeProperties": {"decompositionInterval": 8},
                               "localKnudsenProperties": {"breakdownMax": 0.4, "theta": 1.0, "smoothingPasses": 0}})
    cl.evolve(8)
    d = cl.hybridDecomposition()
    f = cl.fields()  # uniGasVolFields averaged over the same 8 steps
    m, sp = case.mesh, case.meta["species"]
    nI = m.n_internal
    own, nei, cc = m.owner[:nI], m.neighbour, m.cell_centres
    dist = np.linalg.norm(cc[nei] - cc[own], axis=1)
    rho, T, magU = f["rhoM"], f["translationalT"], np.linalg.norm(f["UMean"], axis=1)
    def maxgrad(x):
        g = np.zeros(m.n_cells)
        d_ = np.abs(x[nei] - x[own]) / dist
        np.maximum.at(g, own, d_); np.maximum.at(g, nei, d_)
        return g
    lam = 1.0 / (math.pi * sp["diameter"] ** 2 * f["rhoN"] * (case.meta["Tref"] / T) ** (sp["omega"] - 0.5) * math.sqrt(2.0))
    u0 = np.sqrt(2.0 * kB * T / sp["mass"])
    knRho, knT, knU = lam * maxgrad(rho) / rho, lam * maxgrad(T) / T, lam * maxgrad(magU) / np.maximum(magU, u0)
    np.testing.assert_allclose(d["KnRho"], knRho, rtol=1e-9)
    np.testing.assert_allclose(d["KnT"], knT, rtol=1e-9)
    np.testing.assert_allclose(d["KnU"], knU, rtol=1e-9)
    np.testing.assert_allclose(d["KnGLL"], np.maximum(np.maximum(knRho, knT), knU), rtol=1e-9)
    assert d["KnT"].reshape(12, 16)[-1].mean() > 2 * d["KnT"].reshape(12, 16)[5].mean()  # the temperature jump sits at the hot wall
    raw = d["KnGLL"] > 0.4  # statistical scatter dominates at 60 parcels x 8 steps per cell: the raw mask is patchy
    ids = d["cellCollModelId"].reshape(12, 16)
    assert 0 < ids.sum() < m.n_cells
    assert (d["cellCollModelId"] == raw).mean() > 0.5  # the refinement sweeps tidy the raw mask ...
    iso = lambda a: ((a[1:-1, 1:-1] != a[:-2, 1:-1]) & (a[1:-1, 1:-1] != a[2:, 1:-1]) & (a[1:-1, 1:-1] != a[1:-1, :-2]) & (a[1:-1, 1:-1] != a[1:-1, 2:])).sum()
    assert iso(ids) <= iso(raw.reshape(12, 16).astype(int))  # ... they do not create isolated cells
    cl.close()


def test_local_knudsen_smoothing_and_blending(OracleCloud):
    """smoothingPasses reduce the cell-to-cell scatter of KnGLL; theta < 1 blends successive decompositions
    (localKnudsen.C:392-395): K_2 = theta K_inst + (1 - theta) K_1 lies between the two."""
    res = {}
    for passes in (0, 4):
        cl = _hot_top_hybrid_couette().make_cloud(OracleCloud)
        cl.setHybridDecomposition({"timeProperties": {"decompositionInterval": 6}, "localKnudsenProperties": {"smoothingPasses": passes, "theta": 0.5}})
        cl.evolve(6)
        k1 = cl.hybridDecomposition()["KnGLL"].copy()
        cl.evolve(6)
        k2 = cl.hybridDecomposition()["KnGLL"].copy()
        res[passes] = (k1, k2)
        cl.close()
    rough = lambda k: np.abs(np.diff(k.reshape(12, 16), axis=1)).mean() / k.mean()
    assert rough(res[4][0]) < 0.6 * rough(res[0][0])
    k1, k2 = res[0]
    inst = (k2 - 0.5 * k1) / 0.5  # the second window's instantaneous value implied by the blend
    assert (inst > 0).all() and np.isfinite(inst).all()
    # the fields start from zero: K_1 = 0.5 K_inst,1 and K_2 = 0.5 K_inst,2 + 0.25 K_inst,1, i.e. K_2 / K_1 ~ 1.5
    assert abs(np.median(k2 / k1) - 1.5) < 0.3


def test_mean_free_path_fields_match_kinetic_theory(OracleCloud):
    """measureMeanFreePath fields (uniGasVolFields.C:1124-1232) of a uniform gas at rest: MFP = VHS mean free path
    (Bird 4.65), MCR = equilibrium collision rate (Bird 4.64), dtMCT = deltaT x MCR, dxMFP = cell size / MFP."""
    case = cases.closed_box(n=5, parcels=60000, seed=33, dt_mct=0.3, lambda_per_dx=2.0, binary="noDSMCCollision")
    cl = case.make_cloud(OracleCloud)
    cl.evolve(10)
    f = cl.fields()
    lam = case.meta["lam"]
    nu = cases.vhs_collision_rate(case.meta["n"], case.meta["T0"], case.meta["species"], case.meta["Tref"])
    assert abs(f["MFP"].mean() - lam) < 0.02 * lam
    assert abs(f["MCR"].mean() - nu) < 0.02 * nu
    assert abs(f["dtMCT"].mean() - 0.3) < 0.02 * 0.3
    assert abs(f["dxMFP"].mean() - 0.5) < 0.02 * 0.5
    np.testing.assert_allclose(f["MCT"], 1.0 / f["MCR"], rtol=1e-12)
    # statistical error estimates (:1234-1254): 1 / sqrt(N per cell x samples), velocity error scaled by 1 / (Ma sqrt(gamma))
    nsamp = f["uniGasRhoNMean"] * 10
    np.testing.assert_allclose(f["densityError"], 1.0 / np.sqrt(nsamp), rtol=1e-12)
    np.testing.assert_allclose(f["velocityError"], f["densityError"] / (f["Ma"] * math.sqrt(5.0 / 3.0)), rtol=1e-9)
    cl.close()
