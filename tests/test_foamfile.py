"""On-disk formats: the reader against field files shipped with the reference, write -> read round
trips that return the same bits, and checkpoint / restart of a run (the restarted cloud continues exactly like the
uninterrupted one because the step index - and with it every counter-based stream - is restored)."""
import os

import numpy as np
import pytest

from unigasfoam_b200 import cases, foamfile

GOLD = os.path.join(os.path.dirname(__file__), "golden", "openfoam")


def test_reads_reference_tutorial_fields():
    f = foamfile.read_vol_field(os.path.join(GOLD, "uspRhoNMean_Ar"))
    assert f["class"] == "volScalarField" and f["object"] == "uspRhoNMean_Ar"
    assert f["dimensions"] == [0, 0, 0, 0, 0, 0, 0] and f["uniform"] and f["internal"] == 1.0
    assert list(f["boundary"]) == ["inlet", "outlet", "nozzle", "surface", "axis", "frontWedge", "backWedge"]
    assert f["boundary"]["inlet"] == {"type": "calculated", "value": "uniform 1"}
    assert f["boundary"]["axis"]["type"] == "symmetry" and f["boundary"]["backWedge"]["type"] == "symmetryPlane"
    u = foamfile.read_vol_field(os.path.join(GOLD, "U"))
    assert u["class"] == "volVectorField" and u["dimensions"] == [0, 1, -1, 0, 0, 0, 0]
    assert u["uniform"] and np.array_equal(u["internal"], [0.0, 0.0, 0.0])
    assert np.array_equal(foamfile.expand_internal(u, 5), np.zeros((5, 3)))
    t = foamfile.read_vol_field(os.path.join(GOLD, "transT"))
    assert t["dimensions"] == [0, 0, 0, 1, 0, 0, 0] and t["uniform"]
    assert foamfile.expand_internal(t, 4).shape == (4,)


def test_vol_field_round_trip_bit_exact(tmp_path):
    rng = np.random.default_rng(0)
    s = rng.standard_normal(1000) * 1e-17 + rng.random(1000)
    v = rng.standard_normal((1000, 3)) * 1e3
    foamfile.write_vol_field(str(tmp_path / "uniGasSigmaTcRMax"), "0.001", [0, 3, -1, 0, 0, 0, 0], s, ["walls", "inlet"])
    foamfile.write_vol_field(str(tmp_path / "uniGasSubCellLevels"), "0.001", [0] * 7, v, {"walls": "zeroGradient", "sym": "symmetry"}, vector=True)
    a = foamfile.read_vol_field(str(tmp_path / "uniGasSigmaTcRMax"))
    b = foamfile.read_vol_field(str(tmp_path / "uniGasSubCellLevels"))
    assert not a["uniform"] and np.array_equal(a["internal"], s) and a["dimensions"] == [0, 3, -1, 0, 0, 0, 0]
    assert a["boundary"] == {"walls": {"type": "zeroGradient"}, "inlet": {"type": "zeroGradient"}}
    assert np.array_equal(b["internal"], v) and b["class"] == "volVectorField" and b["boundary"]["sym"]["type"] == "symmetry"
    foamfile.write_vol_field(str(tmp_path / "w"), "0", [0] * 7, np.full(7, 2.5), ["p"])
    w = foamfile.read_vol_field(str(tmp_path / "w"))
    assert w["uniform"] and w["internal"] == 2.5 and np.array_equal(foamfile.expand_internal(w, 7), np.full(7, 2.5))


def test_lagrangian_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(1)
    n = 257
    p = dict(position=rng.standard_normal((n, 3)), U=rng.standard_normal((n, 3)) * 300, cell=rng.integers(0, 99, n).astype(np.int32),
             typeId=rng.integers(0, 2, n).astype(np.int32), ERot=rng.random(n) * 1e-21, cellWeight=rng.random(n) + 0.5)
    d = foamfile.write_lagrangian(str(tmp_path), "0.002", p)
    assert sorted(os.listdir(d)) == sorted(["positions", "U", "cellWeight", "radialWeight", "ERot", "ELevel", "typeId", "newParcel", "vibLevel"])
    head = open(os.path.join(d, "positions")).read(1200)
    assert "class       Cloud<uniGasParcel>;" in head and 'location    "0.002/lagrangian/uniGas";' in head
    assert "class       vectorField;" in open(os.path.join(d, "U")).read(1200)
    q = foamfile.read_lagrangian(str(tmp_path), "0.002")
    for k in ("position", "U", "cell", "typeId", "ERot", "cellWeight"):
        assert np.array_equal(q[k], p[k]), k
    assert (q["radialWeight"] == 1).all() and (q["ELevel"] == 0).all() and q["vibLevel"] == [[]] * n
    # empty cloud
    foamfile.write_lagrangian(str(tmp_path), "0", dict(position=np.empty((0, 3)), U=np.empty((0, 3)), cell=np.empty(0, np.int32)))
    assert len(foamfile.read_lagrangian(str(tmp_path), "0")["cell"]) == 0
    # vibrational levels as the reference writes them: one label list per parcel
    p3 = dict(position=np.zeros((3, 3)), U=np.zeros((3, 3)), cell=np.zeros(3, np.int32), vibLevel=[[0], [2], [1]])
    foamfile.write_lagrangian(str(tmp_path), "1", p3)
    assert foamfile.read_lagrangian(str(tmp_path), "1")["vibLevel"] == [[0], [2], [1]]


def test_field_count_mismatch_is_an_error(tmp_path):
    p = dict(position=np.zeros((4, 3)), U=np.zeros((4, 3)), cell=np.zeros(4, np.int32))
    d = foamfile.write_lagrangian(str(tmp_path), "0", p)
    foamfile.write_io_field(os.path.join(d, "ERot"), "scalarField", "0/lagrangian/uniGas", np.zeros(3), "scalar")
    with pytest.raises(foamfile.FoamFormatError, match="ERot"):
        foamfile.read_lagrangian(str(tmp_path), "0")
    with open(os.path.join(d, "U"), "w") as f:
        f.write("FoamFile { version 2.0; format binary; class vectorField; object U; }\n4\n(")
    with pytest.raises(foamfile.FoamFormatError, match="ascii"):
        foamfile.read_io_field(os.path.join(d, "U"), "vector")


def _restart_equals_uninterrupted(Cloud, tmp_path, case, **kw):
    a = case.make_cloud(Cloud, **kw)
    a.evolve(3)
    t = a.writeTime(str(tmp_path), "3e-06")
    assert os.path.exists(os.path.join(t, "uniGasSigmaTcRMax")) and os.path.exists(os.path.join(t, "uniform", "time"))
    a.evolve(3)
    b = Cloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT, **kw)
    d = b.readTime(str(tmp_path), "3e-06")
    assert d["index"] == 3
    b.evolve(3)
    pa, pb = a.parcels(), b.parcels()
    for k in ("cell", "position", "U", "ERot", "cellWeight"):
        assert np.array_equal(pa[k], pb[k]), k
    ca, cb = a.counters(), b.counters()
    assert ca["step"] == cb["step"] == 6 and ca["collisions"] == cb["collisions"]
    assert np.array_equal(a.cellState()["sigmaTcRMax"], b.cellState()["sigmaTcRMax"])


def test_oracle_restart_continues_bit_exact(tmp_path, OracleCloud):
    case = cases.closed_box(n=6, parcels=15000, seed=21, wall="diffuse", species=("N2", cases.NITROGEN),
                            binary="LarsenBorgnakkeVariableHardSphere", Trot=280.0)
    _restart_equals_uninterrupted(OracleCloud, tmp_path, case)


def test_oracle_restart_cell_weighted(tmp_path, OracleCloud):
    def ramp(mesh):
        x = mesh.cell_centres[:, 0]
        return 0.6 + 1.2 * (x - x.min()) / (x.max() - x.min())
    case = cases.closed_box(n=6, parcels=15000, seed=22, cellWeightFactor=ramp)
    _restart_equals_uninterrupted(OracleCloud, tmp_path, case, parcelCapacity=40000)


@pytest.mark.gpu
def test_gpu_restart_continues_bit_exact(tmp_path, GpuCloud):
    def ramp(mesh):
        x = mesh.cell_centres[:, 0]
        return 0.6 + 1.2 * (x - x.min()) / (x.max() - x.min())
    case = cases.closed_box(n=6, parcels=15000, seed=23, wall="diffuse", cellWeightFactor=ramp)
    _restart_equals_uninterrupted(GpuCloud, tmp_path, case, parcelCapacity=40000)


@pytest.mark.gpu
def test_gpu_reads_what_the_oracle_wrote(tmp_path, GpuCloud, OracleCloud):
    """A time directory is the exchange format between implementations: the oracle writes, libugf restarts from it and
    both continue in lockstep."""
    case = cases.couette(nx=24, ny=12, ppc=20, binary="noDSMCCollision")
    for e in case.boundariesDict["uniGasPatchBoundaries"]:  # specular walls: no libm calls, so the continuation is bit-exact
        e["boundaryModel"] = "uniGasSpecularWallPatch"
    r = case.make_cloud(OracleCloud)
    r.evolve(2)
    r.writeTime(str(tmp_path), "2")
    g = GpuCloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT)
    g.readTime(str(tmp_path), "2")
    r.evolve(3); g.evolve(3)
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"]) and np.array_equal(pg["position"], pr["position"]) and np.array_equal(pg["U"], pr["U"])


def test_polymesh_round_trip_and_run(tmp_path, OracleCloud):
    """constant/polyMesh written in OpenFOAM's layout and read back: same topology, same derived geometry, and a run on
    the re-read mesh is bit-identical to the run on the original (cyclic partners, separations and the unsolved
    direction are recovered from the boundary file)."""
    for case in (cases.couette(nx=12, ny=6, ppc=10, binary="noDSMCCollision"), cases.cylinder(nr=8, ntheta=12, ppc=10, binary="noDSMCCollision")):
        m = case.mesh
        d = foamfile.write_polymesh(str(tmp_path / case.name), m)
        assert sorted(os.listdir(d)) == ["boundary", "faces", "neighbour", "owner", "points"]
        assert 'note        "nPoints:' in open(os.path.join(d, "owner")).read(1500)
        r = foamfile.read_polymesh(str(tmp_path / case.name))
        assert np.array_equal(r.points, m.points) and np.array_equal(r.owner, m.owner) and np.array_equal(r.neighbour, m.neighbour)
        assert np.array_equal(r.face_points, m.face_points) and np.array_equal(r.face_point_offsets, m.face_point_offsets)
        assert [(p.name, p.kind, p.start, p.size, p.partner if p.kind == "cyclic" else -1) for p in r.patches] == \
               [(p.name, p.kind, p.start, p.size, p.partner if p.kind == "cyclic" else -1) for p in m.patches]
        assert tuple(r.solution_d) == tuple(m.solution_d)
        for a in ("face_areas", "face_centres", "cell_volumes", "cell_centres", "cell_bb_min", "cell_bb_max"):
            assert np.array_equal(getattr(r, a), getattr(m, a)), a
        for p, q in zip(r.patches, m.patches):
            assert np.allclose(p.separation, q.separation, rtol=1e-12, atol=1e-15)
        a = case.make_cloud(OracleCloud)
        case.mesh = r
        b = case.make_cloud(OracleCloud)
        a.evolve(4); b.evolve(4)
        pa, pb = a.parcels(), b.parcels()
        assert np.array_equal(pa["cell"], pb["cell"]) and np.allclose(pa["position"], pb["position"], rtol=1e-13, atol=1e-15)


def _hybrid_adaptive_case():
    """Everything that carries state between steps: hybrid run with the local-Knudsen mask, USP-SBGK persistent fields,
    cell weighting with the adapter rewriting factors / levels / time step, time-averaged fields, a pressure inlet."""
    from unigasfoam_b200.adapter import UniGasDynamicAdapter
    case = cases.cylinder(nr=10, ntheta=16, ppc=25, seed=81, mode="hybrid", bgk="unifiedStochasticParticleSBGK", theta=0.3,
                          cellWeightFactor=("particlesPerSubCell", 25), U_inf=600.0)
    case.uniGasProperties["adaptiveSimulation"] = True
    case.uniGasProperties["adaptiveProperties"] = dict(timeStepAdaptation=True, subCellAdaptation=True, cellWeightAdaptation=True,
                                                       adaptationInterval=4, smoothingPasses=2, maxSubCellSizeMFPRatio=4.0)
    case.uniGasProperties["cellWeightedProperties"] = {"particlesPerSubCell": 25}
    n, T = case.meta["n"], case.meta["T_inf"]
    case.boundariesDict["uniGasGeneralBoundaries"] = [{
        "generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasLiouFangPressureInletPatch",
        "uniGasLiouFangPressureInletPatchProperties": {"typeIds": ["Ar"], "moleFractions": {"Ar": 1.0}, "theta": 0.4,
                                                       "inletPressure": 2 * n * cases.kB * T, "inletTemperature": T}}]
    hd = {"decompositionModel": "localKnudsen", "timeProperties": {"decompositionInterval": 3, "resetAtDecomposition": True},
          "localKnudsenProperties": {"breakdownMax": 0.05, "theta": 0.5, "smoothingPasses": 2}}
    return case, hd, UniGasDynamicAdapter


def _full_state_restart(Cloud, tmp_path):
    case, hd, Adapter = _hybrid_adaptive_case()
    kw = dict(parcelCapacity=6 * case.n_parcels)

    def start(restore=None):
        cl = case.make_cloud(Cloud, **kw) if restore is None else Cloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT, **kw)
        cl.setHybridDecomposition(hd)
        if restore is not None:
            cl.readTime(str(tmp_path), restore)
        return cl
    a = start()
    ada = Adapter(a, case.uniGasProperties)
    ada.run(8)                       # two adaptations; the second one has just uploaded factors the parcels do not carry yet
    a.writeTime(str(tmp_path), "8")
    assert os.path.exists(os.path.join(str(tmp_path), "8", "uniGasCellWeightFactorCarried"))
    a.evolve(5)
    b = start("8")
    assert b.counters()["step"] == 8 and b.cfg.deltaT == pytest.approx(ada.last["deltaT"], rel=1e-15)
    b.evolve(5)
    pa, pb = a.parcels(), b.parcels()
    for k in ("cell", "position", "U", "cellWeight"):
        assert np.array_equal(pa[k], pb[k]), k
    ca, cb = a.counters(), b.counters()
    for k in ("step", "nParcels", "collisions", "bgkRelaxations", "inserted", "cloned", "weightDeleted"):
        assert ca[k] == cb[k], k
    fa, fb = a.fields(), b.fields()
    for k in ("rhoN", "translationalT", "UMean"):
        assert np.array_equal(fa[k], fb[k]), k
    # wall tallies are floating-point atomics (device) / OpenMP atomics (oracle): two runs agree to summation order
    same = lambda x, y: np.allclose(x, y, rtol=1e-10, atol=1e-10 * np.abs(y).max())
    assert same(fa["surfaceHeatTransfer"], fb["surfaceHeatTransfer"]) and np.abs(fa["surfaceHeatTransfer"]).max() > 0
    assert np.array_equal(a.hybridDecomposition()["cellCollModelId"], b.hybridDecomposition()["cellCollModelId"])
    assert np.array_equal(a.inletVelocity("inlet"), b.inletVelocity("inlet"))
    sa, sb = a.cellState(), b.cellState()
    for k in ("sigmaTcRMax", "maxProb", "qPrev", "sPrev"):
        assert np.array_equal(sa[k], sb[k]), k
    assert same(a.state(), b.state())
    return a


def test_oracle_restart_restores_every_piece_of_state(tmp_path, OracleCloud):
    a = _full_state_restart(OracleCloud, tmp_path)
    bad = a.state()
    bad[2] += 1
    from unigasfoam_b200.cloud import UgfError
    with pytest.raises(UgfError, match="another mesh"):
        a.loadState(bad)
    with pytest.raises(UgfError, match="wrong size"):
        a.loadState(bad[:-1])


@pytest.mark.gpu
def test_gpu_restart_restores_every_piece_of_state(tmp_path, GpuCloud):
    _full_state_restart(GpuCloud, tmp_path)


@pytest.mark.gpu
def test_gpu_continues_a_hybrid_run_the_oracle_wrote(tmp_path, GpuCloud, OracleCloud):
    """The time directory plus the state array are implementation-neutral: the oracle runs and writes a hybrid, weighted
    case, libugf restarts from it, and the two continue in lockstep (same counters, velocities to round-off)."""
    def ramp(mesh):
        x = mesh.cell_centres[:, 0]
        return 0.7 + 0.9 * (x - x.min()) / (x.max() - x.min())
    case = cases.closed_box(n=6, parcels=30000, seed=83, mode="hybrid", bgk="unifiedStochasticParticleSBGK", number_density=4e20, theta=0.5,
                            cellWeightFactor=ramp)
    case.cellCollModelId = (np.arange(case.mesh.n_cells) % 2).astype(np.int32)
    r = case.make_cloud(OracleCloud)
    r.evolve(4)
    r.writeTime(str(tmp_path), "4")
    g = GpuCloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=int(1.5 * case.n_parcels) + 4096)
    g.readTime(str(tmp_path), "4")
    assert np.allclose(g.state(), r.state(), rtol=1e-12, atol=0)
    r.evolve(4); g.evolve(4)
    cg, cr = g.counters(), r.counters()
    for k in ("step", "nParcels", "cloned", "weightDeleted", "collisionCandidates"):
        assert cg[k] == cr[k], k
    assert abs(cg["collisions"] - cr["collisions"]) <= 2 and abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 2
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert (np.abs(pg["U"] - pr["U"]) <= 1e-8 * np.abs(pr["U"]).max()).all(1).mean() > 0.99
    fg, fr = g.fields(), r.fields()
    assert np.allclose(fg["rhoN"], fr["rhoN"], rtol=1e-9) and np.allclose(fg["translationalT"], fr["translationalT"], rtol=1e-6)
