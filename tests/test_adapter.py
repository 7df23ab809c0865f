"""uniGasDynamicAdapter (U/dynamicAdaptation/uniGasDynamicAdapter.C:228-706), host side over the ABI: the operators it
is built from, its decisions against kinetic theory on an equilibrium gas, cell-weight adaptation on a graded mesh, and
GPU == oracle through adaptation steps (time step, sub-cell levels and weight factors all change under the run)."""
import math

import numpy as np
import pytest

from unigasfoam_b200 import cases
from unigasfoam_b200.adapter import FaceOperators, UniGasDynamicAdapter


def adaptive(case, **ap):
    case.uniGasProperties["adaptiveSimulation"] = True
    case.uniGasProperties["adaptiveProperties"] = ap
    return case


def test_average_interpolate_preserves_linear_fields_and_constants():
    m = cases.closed_box(n=6, parcels=100).mesh
    ops = FaceOperators(m)
    assert np.allclose(ops.average_interpolate(np.full(m.n_cells, 3.5)), 3.5, rtol=1e-14)
    x = m.cell_centres[:, 0]
    sm = ops.average_interpolate(x)
    inner = (x > x.min() + 1e-12) & (x < x.max() - 1e-12)
    assert np.allclose(sm[inner], x[inner], rtol=1e-12)       # uniform mesh: linear interpolation is exact inside
    assert (sm[~inner & (x < x.mean())] > x.min()).all()       # zero-gradient wall faces pull the edge cells inward
    v = np.column_stack([x, 2 * x, -x])
    assert np.allclose(ops.average_interpolate(v)[inner], v[inner], rtol=1e-12)


def test_fvc_smooth_is_the_least_field_with_bounded_neighbour_ratio():
    m = cases.closed_box(n=5, parcels=100).mesh
    ops = FaceOperators(m)
    rng = np.random.default_rng(3)
    f = np.exp(rng.normal(0, 1.5, m.n_cells))
    s = ops.smooth(f, 1.3)
    assert (s >= f).all() and ops.max_neighbour_ratio(s) <= 1.3 * (1 + 1e-12)
    assert s.max() == f.max()                                  # nothing is raised above what a neighbour forces
    raised = s > f
    # every raised cell sits exactly on the bound of one of its neighbours (minimality)
    lo = np.zeros(m.n_cells)
    np.maximum.at(lo, ops.own, s[ops.nei] / 1.3); np.maximum.at(lo, ops.nei, s[ops.own] / 1.3)
    assert np.allclose(s[raised], lo[raised], rtol=1e-14)
    assert np.array_equal(ops.smooth(s, 1.3), s)               # idempotent


def test_oracle_adapter_decisions_match_kinetic_theory(OracleCloud):
    """Equilibrium argon: the measured collision rate and mean free path are Bird 4.64 / 4.65, so the adapted time
    step is dt * min(0.2 / (dt nu), 0.5 / Co) and the sub-cell levels ceil((dx / lambda) / 0.5)."""
    case = adaptive(cases.closed_box(n=6, parcels=60000, seed=31, lambda_per_dx=0.4, dt_mct=0.5),
                    timeStepAdaptation=True, subCellAdaptation=True, adaptationInterval=10, smoothingPasses=5)
    cl = case.make_cloud(OracleCloud)
    ad = UniGasDynamicAdapter(cl, case.uniGasProperties)
    dt0 = case.deltaT
    assert ad.run(10) == 1
    m = case.meta
    nu = cases.vhs_collision_rate(m["n"], m["T0"], m["species"], m["Tref"])
    lam = cases.vhs_mean_free_path(m["n"], m["T0"], m["species"], m["Tref"])
    dx = m["L"] / 6
    assert np.allclose(ad.last["rhoN"], m["n"], rtol=0.15) and abs(ad.last["rhoN"].mean() / m["n"] - 1) < 0.01
    assert abs(ad.last["translationalT"].mean() / m["T0"] - 1) < 0.02
    assert abs(np.median(ad.last["timeStepMCTRatio"]) / (dt0 * nu) - 1) < 0.05
    assert abs(np.median(ad.last["cellSizeMFPRatio"]) / (dx / lam) - 1) < 0.05
    co = cases.most_probable_speed(m["T0"], m["species"]["mass"]) * dt0 / dx
    expect = dt0 * min(0.2 / ad.last["maxTimeStepMCTRatio"], 0.5 / ad.last["maxCourant"])
    assert cl.cfg.deltaT == pytest.approx(expect, rel=1e-12)
    assert abs(ad.last["maxCourant"] / co - 1) < 0.1 and cl.cfg.deltaT < dt0
    want = math.ceil((dx / lam) / 0.5)
    assert want >= 4 and (np.abs(ad.subCellLevels - want) <= 1).all() and np.median(ad.subCellLevels) == want
    cl.evolve(3)  # the new levels and time step are live: NTC now pairs inside sub-cells
    assert cl.counters()["collisions"] > 0


def closed_annulus(**kw):
    """The cylinder O-grid closed on all sides (every boundary a diffuse wall at the gas temperature): a gas at rest
    on a graded mesh, cell volumes spanning a factor 7."""
    case = cases.cylinder(U_inf=0.0, T_wall=200.0, binary="noDSMCCollision", **kw)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        if e["boundaryModel"] == "uniGasDeletionPatch":
            e["boundaryModel"] = "uniGasDiffuseWallPatch"
            e["uniGasDiffuseWallPatchProperties"] = {"velocity": [0, 0, 0], "temperature": 200.0}
    case.boundariesDict["uniGasGeneralBoundaries"] = []
    for p in case.mesh.patches:
        if p.kind == "patch":
            p.kind = "wall"
    return case


def test_oracle_cell_weight_adaptation_evens_out_parcels_per_cell(OracleCloud):
    """All factors 1 to start with: the adapter moves them towards n V / (particlesPerSubCell F_N) - as far as its own
    smoothing (neighbour ratio <= 1.05, capped by minParticlesPerSubCell) lets it - so the parcel counts even out
    while the real gas (sum of the carried factors) stays what it was."""
    case = adaptive(closed_annulus(nr=24, ntheta=24, ppc=30, seed=32, cellWeightFactor=1.0, grading=2.0),
                    cellWeightAdaptation=True, adaptationInterval=5, smoothingPasses=2)
    case.uniGasProperties["cellWeightedProperties"] = {"particlesPerSubCell": 30, "minParticlesPerSubCell": 10}
    cl = case.make_cloud(OracleCloud, parcelCapacity=6 * case.n_parcels)
    ad = UniGasDynamicAdapter(cl, case.uniGasProperties)
    cnt0 = np.bincount(case.cell, minlength=case.mesh.n_cells)
    cv0 = cnt0.std() / cnt0.mean()
    assert ad.run(60) == 12
    W = ad.cellWeightFactor
    target = case.meta["n"] * case.mesh.cell_volumes / (30 * cl.cfg.nParticle)
    assert np.corrcoef(np.log(W), np.log(target))[0, 1] > 0.85
    assert ad.ops.max_neighbour_ratio(ad.last["cellWeightTarget"]) <= 1.3
    cl.evolve(1)  # the weighting pass of this step applies the factors uploaded by the last adaptation
    p = cl.parcels()
    assert np.array_equal(p["cellWeight"], W[p["cell"]])
    cnt = np.bincount(p["cell"], minlength=case.mesh.n_cells)
    assert cnt.std() / cnt.mean() < 0.9 * cv0, (cv0, cnt.std() / cnt.mean())
    assert abs(p["cellWeight"].sum() / cnt0.sum() - 1) < 0.04
    assert cl.counters()["deleted"] == 0 and cl.counters()["stuck"] == 0


@pytest.mark.gpu
def test_gpu_adaptive_run_in_lockstep_with_oracle(GpuCloud, OracleCloud):
    """Collision-free so that the states stay bit-identical: then the accumulators, hence every adaptation decision
    (time step, sub-cell levels, factors), hence the cloned / deleted parcels must be the same on both sides."""
    def make():
        c = adaptive(cases.cylinder(nr=12, ntheta=20, ppc=25, seed=33, cellWeightFactor=1.0, binary="noDSMCCollision"),
                     timeStepAdaptation=True, subCellAdaptation=True, cellWeightAdaptation=True, adaptationInterval=4, smoothingPasses=3)
        c.uniGasProperties["cellWeightedProperties"] = {"particlesPerSubCell": 25}
        for e in c.boundariesDict["uniGasPatchBoundaries"]:
            if e["boundaryModel"] == "uniGasDiffuseWallPatch":
                e["boundaryModel"] = "uniGasSpecularWallPatch"
        return c
    case = make()
    g = case.make_cloud(GpuCloud, parcelCapacity=6 * case.n_parcels)
    r = case.make_cloud(OracleCloud, parcelCapacity=6 * case.n_parcels)
    ag, ar = UniGasDynamicAdapter(g, case.uniGasProperties), UniGasDynamicAdapter(r, case.uniGasProperties)
    for _ in range(3):
        assert ag.run(4) == 1 and ar.run(4) == 1
        assert g.cfg.deltaT == pytest.approx(r.cfg.deltaT, rel=1e-9)
        assert np.array_equal(ag.subCellLevels, ar.subCellLevels)
        assert np.allclose(ag.cellWeightFactor, ar.cellWeightFactor, rtol=1e-9)
    cg, cr = g.counters(), r.counters()
    for k in ("nParcels", "cloned", "weightDeleted", "inserted", "deleted"):
        assert cg[k] == cr[k], k
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert (np.abs(pg["position"] - pr["position"]) <= 1e-12 * np.abs(pr["position"]).max()).all(1).mean() > 0.999
