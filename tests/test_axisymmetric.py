"""axisymmetricSimulation true (U/clouds/uniGasCloud.C:420, 563-568, 1427-1570, uniGasCloudI.H:116-120): a wedge
about the x axis whose parcels carry a radial weighting factor RWF(x) = 1 + (maxRWF - 1) sqrt(y^2 + z^2) / radialExtent.

What the reference does with it, each checked here on the oracle (closed forms) and on the GPU against the oracle:
  * uniGasMeshFill and the inflow patches hand out RWF(centre of the parcel's cell); after every move axisymmetricWeighting /
    axisymmetricCellWeighting set RWF(position) and clone / delete with the ratio old / new of CWF * RWF;
  * the XnParticle cell sums weight every parcel with its own RWF (cellMeasurements.C:463-467), so density, velocity and
    temperature fields stay those of the gas although the parcels per unit volume fall towards the outer radius;
  * NTC candidates use the mean RWF(position) of the cell (noTimeCounter.C:168-184);
  * the inflow count divides by RWF(face centre) (uniGasGeneralBoundary.C:154-165);
  * wall sums and wall fields carry RWF(hit position) / RWF(face centre) (uniGasPatchBoundary.C:292-299, uniGasVolFields.C:1276-1278);
  * the BGK conservation step weights every parcel with CWF RWF nParticle (…USP.C:1017-1022);
  * the pressure inlets weight the cell's parcels with RWF(position) (…LiouFangPressureInletPatch.C:154)."""
import math

import numpy as np
import pytest

from unigasfoam_b200 import cases
from unigasfoam_b200.cloud import UgfError

kB = cases.kB


def rwf_of(case, pts):
    return cases.radial_weight(case.uniGasProperties, pts)


def make(Cloud, **kw):
    case = cases.axisymmetric_tube(**kw)
    return case, case.make_cloud(Cloud, parcelCapacity=3 * case.n_parcels + 4096)


def test_fill_hands_out_centre_weights_and_the_first_move_position_weights(OracleCloud):
    case, cl = make(OracleCloud, nx=8, nr=6, ppc=30, binary="noDSMCCollision")
    p = cl.parcels()
    assert np.array_equal(p["radialWeight"], rwf_of(case, case.mesh.cell_centres[p["cell"]]))  # uniGasMeshFill.C:260
    assert p["radialWeight"].min() >= 1.0 and p["radialWeight"].max() > 0.8 * 8.0
    cl.evolve(1)
    q = cl.parcels()
    assert np.array_equal(q["radialWeight"], rwf_of(case, q["position"]))  # uniGasCloud.C:1436-1438
    c = cl.counters()
    assert c["cloned"] > 0 and c["weightDeleted"] > 0
    # download -> upload round trip: position weights are accepted, anything else is refused
    cl2 = OracleCloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=4 * len(q["cell"]))
    cl2.setParcels(q["position"], q["U"], q["cell"], radialWeight=q["radialWeight"])
    assert np.array_equal(cl2.parcels()["radialWeight"], q["radialWeight"])
    with pytest.raises(UgfError, match="radialWeight"):
        cl2.setParcels(q["position"], q["U"], q["cell"], radialWeight=np.full(len(q["cell"]), 3.0))


def test_parcels_per_cell_follow_volume_over_rwf(OracleCloud):
    """n V / (F_N RWF) parcels per cell: with maxRWF = R-proportional weights the count per radial row is nearly flat instead
    of growing with the radius."""
    case = cases.axisymmetric_tube(nx=10, nr=12, ppc=60, max_rwf=50.0)
    cnt = np.bincount(case.cell, minlength=case.mesh.n_cells).reshape(12, 10).sum(1).astype(float)
    expect = (case.meta["n"] * case.mesh.cell_volumes / (case.uniGasProperties["nEquivalentParticles"] * case.meta["rwf_centre"])).reshape(12, 10).sum(1)
    assert np.all(np.abs(cnt - expect) < 5 * np.sqrt(expect) + 1)
    vol_rows = case.mesh.cell_volumes.reshape(12, 10).sum(1)
    assert vol_rows[-1] / vol_rows[1] > 7 and cnt[-1] / cnt[1] < 1.6


def test_weighting_keeps_the_gas_uniform(OracleCloud):
    """Gas at rest in a tube with specular wall and equilibrium reservoirs at both ends: the weighting pass clones / deletes so
    that the represented density stays n everywhere although RWF varies by a factor 8 across the radius."""
    case, cl = make(OracleCloud, nx=10, nr=10, ppc=60, U_inf=0.0, wall="uniGasSpecularWallPatch", seed=11)
    # reservoir at the outlet too (a second free-stream patch at rest)
    g = case.boundariesDict["uniGasGeneralBoundaries"]
    g.append({"generalBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
              "uniGasFreeStreamInflowPatchProperties": dict(g[0]["uniGasFreeStreamInflowPatchProperties"])})
    cl = case.make_cloud(OracleCloud, parcelCapacity=3 * case.n_parcels)
    n0 = case.n_parcels
    cloned = deleted = coll = 0
    steps = 60
    for _ in range(steps):
        cl.evolve(1)
        c = cl.counters()
        cloned += c["cloned"]; deleted += c["weightDeleted"]; coll += c["collisions"]
        assert c["stuck"] == 0
    assert cloned > 500 and deleted > 500
    assert abs(cloned - deleted) < 6 * math.sqrt(cloned + deleted)  # at rest as many move inwards as outwards
    assert abs(cl.size() - n0) < 0.05 * n0
    f = cl.fields()
    rows = (f["rhoN"] / case.meta["n"]).reshape(10, 10).mean(1)
    assert np.all(np.abs(rows[1:] - 1.0) < 0.04) and abs(rows[0] - 1.0) < 0.08, rows  # the axis row holds the fewest parcels
    T = f["translationalT"].reshape(10, 10).mean(1)
    assert np.all(np.abs(T / case.meta["T_inf"] - 1.0) < 0.06) and abs(T.mean() / case.meta["T_inf"] - 1.0) < 0.03, T  # 6000 parcels: the fill itself is 300 K +- 3 K
    # NTC with the mean RWF of the cell: the simulated collision count is 1/2 N nu dt for the parcels present
    nu = cases.vhs_collision_rate(case.meta["n"], case.meta["T_inf"], case.meta["species"], case.meta["Tref"])
    expect = 0.5 * n0 * nu * case.deltaT * steps
    assert abs(coll - expect) < 0.06 * expect, (coll, expect)


def test_inflow_count_divides_by_the_face_centre_weight(OracleCloud):
    case = cases.axisymmetric_tube(nx=6, nr=8, ppc=200, binary="noDSMCCollision", U_inf=300.0, max_rwf=20.0, seed=5)
    case.position, case.U, case.cell = case.position[:0], case.U[:0], case.cell[:0]
    case.ERot = None
    cl = case.make_cloud(OracleCloud, parcelCapacity=400_000)
    m = case.mesh
    pt = m.patches[m.patch_index("inlet")]
    S = m.face_areas[pt.start:pt.start + pt.size]
    A = np.linalg.norm(S, axis=1)
    fC = m.face_centres[pt.start:pt.start + pt.size]
    own = np.asarray(m.owner[pt.start:pt.start + pt.size])
    cmp_ = cases.most_probable_speed(case.meta["T_inf"], case.meta["species"]["mass"])
    s = case.meta["U_inf"] / cmp_
    flux = (math.exp(-s * s) + math.sqrt(math.pi) * s * (1 + math.erf(s))) / (2 * math.sqrt(math.pi))
    FN = case.uniGasProperties["nEquivalentParticles"]
    steps = 40
    expect = steps * A * case.meta["n"] * case.deltaT * cmp_ * flux / (FN * rwf_of(case, fC))
    counts = np.zeros(pt.size)
    face_of_cell = {int(c): i for i, c in enumerate(own)}
    for _ in range(steps):
        n0 = cl.size()
        cl.controlBeforeMove()
        q = cl.parcels()
        new = slice(n0, None)
        np.add.at(counts, [face_of_cell[int(c)] for c in q["cell"][new]], 1.0)
        assert np.array_equal(q["radialWeight"][new], rwf_of(case, m.cell_centres[q["cell"][new]]))  # uniGasGeneralBoundary.C:739
        cl.move(); cl.finishStep()
    assert expect.min() > 30
    assert np.all(np.abs(counts - expect) < 5 * np.sqrt(expect) + 1), (counts, expect)
    # without the division the outer faces would insert ~ RWF times more
    assert counts[-1] / counts[0] < 0.6 * (A[-1] / A[0])


def test_wall_pressure_of_a_gas_at_rest(OracleCloud):
    """Diffuse wall at the gas temperature, gas at rest: the sampled wall pressure is n k T - through RWF(hit position) in the
    momentum sums and nothing else."""
    case = cases.axisymmetric_tube(nx=12, nr=8, ppc=80, U_inf=0.0, T_wall=300.0, T_inf=300.0, binary="noDSMCCollision", seed=9)
    g = case.boundariesDict["uniGasGeneralBoundaries"]
    g.append({"generalBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
              "uniGasFreeStreamInflowPatchProperties": dict(g[0]["uniGasFreeStreamInflowPatchProperties"])})
    cl = case.make_cloud(OracleCloud, parcelCapacity=3 * case.n_parcels)
    cl.evolve(150)
    f = cl.fields()
    m = case.mesh
    pt = m.patches[m.patch_index("wall")]
    b0 = pt.start - m.n_internal
    p_wall = f["wall_p"][b0:b0 + pt.size]
    p = case.meta["n"] * kB * 300.0
    assert abs(p_wall.mean() / p - 1.0) < 0.03, (p_wall.mean(), p)


@pytest.mark.parametrize("bgk", ["stochasticParticleBGK", "unifiedStochasticParticleSBGK"])
def test_bgk_relaxation_conserves_weighted_momentum_and_energy(OracleCloud, bgk):
    case, cl = make(OracleCloud, nx=8, nr=8, ppc=60, mode="bgk", bgk=bgk, U_inf=150.0, seed=4)
    cl.evolve(3)
    cl.calculateFields()
    p0 = cl.parcels()
    cl.relax()
    p1 = cl.parcels()
    assert cl.counters()["bgkRelaxations"] > 100
    assert np.array_equal(p0["cell"], p1["cell"])
    w = p0["radialWeight"]
    nC = case.mesh.n_cells
    for k in range(3):
        a = np.bincount(p0["cell"], w * p0["U"][:, k], nC); b = np.bincount(p1["cell"], w * p1["U"][:, k], nC)
        assert np.allclose(a, b, rtol=0, atol=1e-9 * np.abs(a).max())
    e0 = np.bincount(p0["cell"], w * (p0["U"] ** 2).sum(1), nC); e1 = np.bincount(p1["cell"], w * (p1["U"] ** 2).sum(1), nC)
    assert np.allclose(e0, e1, rtol=1e-10)
    # unweighted sums are NOT conserved: the weights matter
    u0 = np.bincount(p0["cell"], p0["U"][:, 1], nC); u1 = np.bincount(p1["cell"], p1["U"][:, 1], nC)
    assert np.abs(u0 - u1).max() > 1e-6 * np.abs(p0["U"]).max()


def test_bad_axisymmetric_properties_fail(OracleCloud):
    case = cases.axisymmetric_tube(nx=4, nr=4, ppc=5)
    props = dict(case.uniGasProperties, axisymmetricProperties={"radialExtentOfDomain": 0.0, "maxRadialWeightingFactor": 5.0})
    with pytest.raises(UgfError, match="radialExtentOfDomain"):
        OracleCloud(case.mesh, props, case.boundariesDict, case.deltaT, parcelCapacity=1000)


def test_restart_through_time_directories(tmp_path, OracleCloud):
    """lagrangian/uniGas/radialWeight in the time directory: written after a step it holds RWF(position) and the restarted run
    continues bit for bit; written before the first step it holds the initialisation's RWF(cell centre) and the restart's first
    weighting pass clones / deletes exactly like the uninterrupted run's."""
    import os
    for steps_before in (3, 0):
        case = cases.axisymmetric_tube(nx=8, nr=6, ppc=30, cell_weighted=True, seed=17)
        kw = dict(parcelCapacity=4 * case.n_parcels)
        a = case.make_cloud(OracleCloud, **kw)
        if steps_before:
            a.evolve(steps_before)
        d = str(tmp_path / f"case{steps_before}")
        t = a.writeTime(d, "1e-06")
        assert os.path.exists(os.path.join(t, "lagrangian", "uniGas", "radialWeight"))
        pw = a.parcels()
        want = rwf_of(case, pw["position"]) if steps_before else rwf_of(case, case.mesh.cell_centres[pw["cell"]])
        assert np.array_equal(pw["radialWeight"], want)
        a.evolve(3)
        b = OracleCloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT, **kw)
        b.readTime(d, "1e-06")
        assert np.array_equal(b.parcels()["radialWeight"], want)
        b.evolve(3)
        pa, pb = a.parcels(), b.parcels()
        for k in ("cell", "position", "U", "radialWeight", "cellWeight"):
            assert np.array_equal(pa[k], pb[k]), (steps_before, k)
        ca, cb = a.counters(), b.counters()
        assert ca["cloned"] == cb["cloned"] and ca["weightDeleted"] == cb["weightDeleted"] and ca["collisions"] == cb["collisions"]


def test_adapter_cell_weights_carry_the_radial_factor(OracleCloud):
    """uniGasDynamicAdapter::calculateCellWeightFactor divides by RWF(cell centre) (uniGasDynamicAdapter.C:442): with the factor
    field it computes every cell of a uniform gas ends up with the requested parcels per cell, on the axis and at the wall."""
    from unigasfoam_b200.adapter import UniGasDynamicAdapter
    case = cases.axisymmetric_tube(nx=6, nr=8, ppc=25, cell_weighted=True, U_inf=0.0, wall="uniGasSpecularWallPatch", seed=19)
    props = dict(case.uniGasProperties, adaptiveSimulation=True,
                 adaptiveProperties={"timeStepAdaptation": False, "subCellAdaptation": False, "cellWeightAdaptation": True, "adaptationInterval": 5,
                                     "maxTimeStepMCTRatio": 0.2, "maxCourantNumber": 0.5, "maxSubCellSizeMFPRatio": 1.0})
    cl = OracleCloud(case.mesh, props, case.boundariesDict, case.deltaT, parcelCapacity=4 * case.n_parcels)
    cl.setCellState(cellWeightFactor=case.cellWeightFactor)
    cl.setParcels(case.position, case.U, case.cell)
    cl.setCellState(sigmaTcRMax=case.sigmaTcRMax)
    ad = UniGasDynamicAdapter(cl, props)
    rhoN = np.full(case.mesh.n_cells, case.meta["n"])
    levels = np.ones((case.mesh.n_cells, 3), np.int32)
    w = ad.calculate_cell_weight_factor(rhoN, levels)
    expect = case.meta["n"] * case.mesh.cell_volumes / (25 * props["nEquivalentParticles"] * case.meta["rwf_centre"])
    assert np.allclose(w, expect, rtol=1e-13)
    assert np.allclose(w, case.cellWeightFactor, rtol=1e-12)  # the same rule uniGasMeshFill applies (uniGasMeshFill.C:111-121)


def test_plume_impingement_dictionaries_run_as_written(tmp_path, OracleCloud):
    """The reference's axisymmetric tutorial plumeImpingement: its constant/uniGasProperties and system/*Dict copied unmodified
    (tests/golden/openfoam/plumeImpingement; one adaptation threshold overridden, see below) - axisymmetric + cell-weighted + adaptive + hybrid USP-SBGK / NTC with
    macroInterpolation true, Liou-Fang pressure inlet, two diffuse walls, deleting outlet, uniGasMeshFieldFill from the start time
    directory - on a wedge with the tutorial's patch names and radial extent (its four graded blocks with the curved nozzle are not
    rebuilt: one block, nozzle wall and outlet sharing the outer radius, impingement surface at the far end)."""
    import os
    import shutil
    from unigasfoam_b200 import foamdict, foamfile, mesh as ugmesh
    from unigasfoam_b200.adapter import UniGasDynamicAdapter
    gold = os.path.join(os.path.dirname(__file__), "golden", "openfoam", "plumeImpingement")
    case_dir = str(tmp_path / "plumeImpingement")
    shutil.copytree(gold, case_dir)
    nx, nr = 24, 10
    kinds = {"xMin": ("inlet", "patch"), "xMax": ("surface", "wall"), "yMin": ("axis", "symmetry"), "yMax": ("outer", "patch"),
             "zMin": ("backWedge", "symmetryPlane"), "zMax": ("frontWedge", "symmetryPlane")}
    m = ugmesh.structured_block(nx, nr, 1, ugmesh.wedge_map(nx, nr, 0.139, 7.5e-2, 0.5), kinds)
    ugmesh.split_patch(m, "outer", 8, "nozzle", "outlet", kind_a="wall", kind_b="patch")
    m.meta_axis_aligned = False
    assert [p.name for p in m.patches] == ["inlet", "surface", "axis", "nozzle", "outlet", "backWedge", "frontWedge"]
    # the start time directory uniGasMeshFieldFill reads: gas at rest at the inlet state (1000 Pa, 300 K), thinning out towards the
    # outlet half of the domain
    t0 = os.path.join(case_dir, "0")
    os.makedirs(t0)
    names = [p.name for p in m.patches]
    nC = m.n_cells
    n_in = 1000.0 / (kB * 300.0)
    n0 = n_in * np.where(m.cell_centres[:, 0] < 0.05, 1.0, 0.2)
    foamfile.write_vol_field(os.path.join(t0, "numberDensity_Ar"), "0", [0, -3, 0, 0, 0, 0, 0], n0, names)
    foamfile.write_vol_field(os.path.join(t0, "transT"), "0", [0, 0, 0, 1, 0, 0, 0], np.full(nC, 300.0), names)
    foamfile.write_vol_field(os.path.join(t0, "rotT"), "0", [0, 0, 0, 1, 0, 0, 0], np.full(nC, 300.0), names)
    foamfile.write_vol_field(os.path.join(t0, "U"), "0", [0, 1, -1, 0, 0, 0, 0], np.zeros((nC, 3)), names, vector=True)
    # one test-scale accommodation, as for the other tutorials (tests/test_foamdict.py): the cells here are ~1000 mean free paths wide,
    # so the sub-cell criterion (one sub-cell per mean free path, 20 parcels in each) is relaxed - else 4.7 M parcels in 240 cells
    case, ld = cases.from_case_dir(case_dir, m, seed=5, overrides={"adaptiveProperties": {"maxSubCellSizeMFPRatio": 1.0e4}})
    props = case.uniGasProperties
    assert props["axisymmetricSimulation"] is True and props["collisionProperties"]["macroInterpolation"] is True
    assert props["axisymmetricProperties"] == {"radialExtentOfDomain": 7.5e-2, "maxRadialWeightingFactor": 1000}
    # uniGasMeshFill's rule with both weights: particlesPerSubCell parcels in every sub-cell, on the axis and at the outer radius
    nSub = np.ones(nC) if case.subCellLevels is None else case.subCellLevels.prod(1)
    cnt = np.bincount(case.cell, minlength=nC) / nSub
    assert abs(cnt.mean() - 20) < 1.5 and cnt.reshape(nr, nx).mean(1).min() > 15
    cl = case.make_cloud(OracleCloud, parcelCapacity=8 * case.n_parcels, sampleInterval=foamdict.sample_interval(ld["fieldPropertiesDict"]))
    assert cl.axisymmetric and cl.cfg.macroInterpolation == 1
    cl.setHybridDecomposition(ld["hybridDecompositionDict"])
    ad = UniGasDynamicAdapter(cl, props)
    if case.subCellLevels is not None:
        ad.subCellLevels = case.subCellLevels.copy()
    n = cl.size()
    tally = dict(inserted=0, deleted=0, cloned=0, weightDeleted=0, collisions=0, bgkRelaxations=0, wallHits=0)
    for _ in range(3):
        ad.run(10)   # adaptationInterval 10 in the tutorial: three adaptations
        ad.cellCollModelId = cl.hybridDecomposition()["cellCollModelId"]
    c = cl.counters()
    assert c["stuck"] == 0 and c["step"] == 30
    p = cl.parcels()
    assert np.array_equal(p["radialWeight"], rwf_of(case, p["position"]))
    f = cl.fields()
    assert np.isfinite(f["rhoN"]).all() and np.isfinite(f["translationalT"]).all() and (f["rhoN"] > 0).all()
    assert abs(np.average(f["translationalT"], weights=m.cell_volumes) / 300.0 - 1.0) < 0.1
    v = cl.inletVelocity("inlet")
    assert np.isfinite(v).all()


# ---- GPU against the oracle ------------------------------------------------------------------------------------------------

def _lockstep(g, r, steps, fields=True, rtol=1e-9):
    for _ in range(steps):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        for k in ("nParcels", "inserted", "deleted", "cloned", "weightDeleted", "collisionCandidates", "collisions", "bgkRelaxations", "wallHits", "stuck"):
            assert cg[k] == cr[k], (k, cg[k], cr[k])
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert np.allclose(pg["radialWeight"], pr["radialWeight"], rtol=1e-12, atol=0)  # inserted parcels' positions go through libm on both sides
    if fields:
        fg, fr = g.fields(), r.fields()
        for k in ("rhoN", "rhoM", "translationalT", "p", "wall_p", "surfaceHeatTransfer"):
            assert np.allclose(fg[k], fr[k], rtol=rtol, atol=rtol * np.abs(fr[k]).max()), k
        assert np.allclose(fg["UMean"], fr["UMean"], rtol=rtol, atol=rtol * np.abs(fr["UMean"]).max())
    return pg, pr


@pytest.mark.gpu
def test_gpu_axisymmetric_collisionless_bit_exact(GpuCloud, OracleCloud):
    kw = dict(nx=10, nr=8, ppc=30, binary="noDSMCCollision", wall="uniGasSpecularWallPatch", seed=2)
    case, g = make(GpuCloud, **kw)
    _, r = make(OracleCloud, **kw)
    assert np.array_equal(g.parcels()["radialWeight"], r.parcels()["radialWeight"])
    for _ in range(10):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        for k in ("nParcels", "inserted", "deleted", "cloned", "weightDeleted", "wallHits"):
            assert cg[k] == cr[k], k
    assert cr["cloned"] > 0 and cr["weightDeleted"] > 0
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"]) and np.allclose(pg["radialWeight"], pr["radialWeight"], rtol=1e-12, atol=0)
    same = (pg["position"] == pr["position"]).all(1) & (pg["U"] == pr["U"]).all(1)
    assert same.mean() > 0.8  # every parcel of the initial fill and its clones: bit for bit; inserted ones go through libm on both sides
    assert np.array_equal(pg["radialWeight"][same], pr["radialWeight"][same])


@pytest.mark.gpu
@pytest.mark.parametrize("cell_weighted", [False, True])
def test_gpu_axisymmetric_dsmc_in_lockstep(GpuCloud, OracleCloud, cell_weighted):
    kw = dict(nx=12, nr=10, ppc=30, cell_weighted=cell_weighted, seed=6, species=("N2", cases.NITROGEN), binary="LarsenBorgnakkeVariableHardSphere")
    case, g = make(GpuCloud, **kw)
    _, r = make(OracleCloud, **kw)
    _lockstep(g, r, 12)
    assert r.counters()["collisions"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode,bgk", [("bgk", "unifiedStochasticParticleSBGK"), ("hybrid", "stochasticParticleESBGK")])
def test_gpu_axisymmetric_bgk_in_lockstep(GpuCloud, OracleCloud, mode, bgk):
    kw = dict(nx=10, nr=8, ppc=40, mode=mode, bgk=bgk, cell_weighted=True, seed=8)
    case = cases.axisymmetric_tube(**kw)
    if mode == "hybrid":
        case.cellCollModelId = (np.arange(case.mesh.n_cells) % 10 < 5).astype(np.int32)
    g = case.make_cloud(GpuCloud, parcelCapacity=3 * case.n_parcels)
    r = case.make_cloud(OracleCloud, parcelCapacity=3 * case.n_parcels)
    _lockstep(g, r, 10, rtol=1e-8)
    assert r.counters()["bgkRelaxations"] > 0


@pytest.mark.gpu
def test_gpu_axisymmetric_pressure_inlet_in_lockstep(GpuCloud, OracleCloud):
    kw = dict(nx=10, nr=8, ppc=40, inlet="uniGasLiouFangPressureInletPatch", cell_weighted=True, seed=12)
    case, g = make(GpuCloud, **kw)
    _, r = make(OracleCloud, **kw)
    _lockstep(g, r, 12)
    assert np.allclose(g.inletVelocity("inlet"), r.inletVelocity("inlet"), rtol=1e-10, atol=1e-9)
    assert np.abs(r.inletVelocity("inlet")).max() > 0


@pytest.mark.gpu
def test_gpu_axisymmetric_weighting_keeps_the_gas_uniform(GpuCloud):
    """The closed-form check of the oracle test above on the GPU path itself, at a size the GPU does in a blink."""
    case = cases.axisymmetric_tube(nx=40, nr=40, ppc=100, U_inf=0.0, wall="uniGasSpecularWallPatch", seed=13)
    gl = case.boundariesDict["uniGasGeneralBoundaries"]
    gl.append({"generalBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
               "uniGasFreeStreamInflowPatchProperties": dict(gl[0]["uniGasFreeStreamInflowPatchProperties"])})
    cl = case.make_cloud(GpuCloud, parcelCapacity=2 * case.n_parcels)
    n0 = case.n_parcels
    coll = 0
    steps = 100
    for _ in range(steps):
        cl.evolve(1)
        coll += cl.counters()["collisions"]
    assert abs(cl.size() - n0) < 0.03 * n0
    f = cl.fields()
    rows = (f["rhoN"] / case.meta["n"]).reshape(40, 40).mean(1)
    assert np.all(np.abs(rows - 1.0) < 0.02), rows
    nu = cases.vhs_collision_rate(case.meta["n"], case.meta["T_inf"], case.meta["species"], case.meta["Tref"])
    expect = 0.5 * n0 * nu * case.deltaT * steps
    assert abs(coll - expect) < 0.03 * expect
