"""The reference's own case dictionaries (tutorials/uniGasFoam/hypersonicCylinder, copied as fixtures under
tests/golden/openfoam/hypersonicCylinder) parsed and run: SURVEY §8c lists them as the fixtures that pin schemas and
constants.  The mesh is this repo's O-grid with the tutorial's patch names (the tutorial's blockMeshDict is not
read); `macroInterpolation true` of the tutorial is switched off (cell values), everything else is taken as written:
hybrid USP-SBGK / NTC-VHS, cell weighting, dynamic adaptation, local-Knudsen decomposition, free-stream inflow."""
import os

import numpy as np
import pytest

from unigasfoam_b200 import cases, foamdict, mesh as ugmesh
from unigasfoam_b200.adapter import UniGasDynamicAdapter

CASE = os.path.join(os.path.dirname(__file__), "golden", "openfoam", "hypersonicCylinder")


def test_parses_the_tutorial_dictionaries():
    ld = foamdict.load_case(CASE)
    p = ld["uniGasProperties"]
    assert p["nEquivalentParticles"] == 5.85e11 and p["cellWeightedSimulation"] is True and p["axisymmetricSimulation"] is False
    assert p["cellWeightedProperties"] == {"minParticlesPerSubCell": 20, "particlesPerSubCell": 20}
    assert p["adaptiveProperties"]["adaptationInterval"] == 100 and p["adaptiveProperties"]["subCellAdaptation"] is True
    assert p["collisionModel"] == "hybrid" and p["bgkCollisionModel"] == "unifiedStochasticParticleSBGK"
    assert p["dsmcCollisionModel"] == "variableHardSphere" and p["dsmcCollisionPartnerModel"] == "noTimeCounter"
    assert p["collisionProperties"] == {"Tref": 1000, "macroInterpolation": True, "theta": 0.1}
    assert p["typeIdList"] == ["Ar"]
    ar = p["moleculeProperties"]["Ar"]
    assert ar["mass"] == 66.3e-27 and ar["diameter"] == 3.595e-10 and ar["omega"] == 0.734 and ar["alpha"] == 1.0
    assert ar["characteristicVibrationalTemperature"] == [] and ar["electronicEnergyList"] == [0] and ar["degeneracyList"] == [1]
    assert ld["deltaT"] == 5e-8
    bd = ld["boundariesDict"]
    assert [e["boundaryModel"] for e in bd["uniGasPatchBoundaries"]] == ["uniGasDiffuseWallPatch", "uniGasDeletionPatch", "uniGasDeletionPatch"]
    assert bd["uniGasPatchBoundaries"][0]["patchBoundaryProperties"]["patch"] == "cylinder"
    assert bd["uniGasPatchBoundaries"][0]["uniGasDiffuseWallPatchProperties"] == {"velocity": [0, 0, 0], "temperature": 500}
    fs = bd["uniGasGeneralBoundaries"][0]["uniGasFreeStreamInflowPatchProperties"]
    assert fs["velocity"] == [2634.7, 0, 0] and fs["numberDensities"] == {"Ar": 4.247e20} and fs["typeIds"] == ["Ar"]
    assert bd["uniGasCyclicBoundaries"] == []
    hd = ld["hybridDecompositionDict"]
    assert hd["decompositionModel"] == "localKnudsen" and hd["localKnudsenProperties"] == {"breakdownMax": 0.05, "theta": 0.2, "smoothingPasses": 5}
    assert hd["timeProperties"]["resetAtDecomposition"] is True
    fp = ld["fieldPropertiesDict"]["uniGasFields"][0]
    assert fp["fieldModel"] == "uniGasVolFields" and fp["uniGasVolFieldsProperties"]["measureMeanFreePath"] is True
    assert foamdict.sample_interval(ld["fieldPropertiesDict"]) == 1
    ini = ld["uniGasInitialisationDict"]["configurations"][0]
    assert ini["type"] == "uniGasMeshFill" and ini["numberDensities"] == {"Ar": 4.247e20}


def test_parser_errors_and_syntax_corners():
    d = foamdict.parse('a 1; b (1 2 3); c { d on; e "x y"; } f{g 2.5e3;} h ((1 2) (3 4)); k; // tail\n/* block */ l word;')
    assert d == {"a": 1, "b": [1, 2, 3], "c": {"d": True, "e": "x y"}, "f": {"g": 2500.0}, "h": [[1, 2], [3, 4]], "k": None, "l": "word"}
    assert foamdict.parse("m ( n { o 1; } n { o 2; } );") == {"m": [{"o": 1}, {"o": 2}]}
    for bad in ("a 1", "a { b 1;", "a ( 1 2;", "} a 1;"):
        with pytest.raises(foamdict.FoamDictError):
            foamdict.parse(bad)


def tutorial_case(seed=61):
    m = ugmesh.half_annulus_mesh(20, 32, 0.5 * 0.3048, 2.0 * 0.3048, 0.1 * 0.3048, 5.0)
    m.meta_axis_aligned = False
    # test-scale accommodations: cell values instead of macroInterpolation; the mesh is ~40 times coarser than the tutorial's,
    # so the sub-cell criterion is relaxed by that much (else every cell gets the maximum of 100 sub-cells), the run is short
    case, ld = cases.from_case_dir(CASE, m, seed=seed, overrides={"collisionProperties": {"macroInterpolation": False},
                                                                 "adaptiveProperties": {"maxSubCellSizeMFPRatio": 4.0, "adaptationInterval": 20}})
    ld["hybridDecompositionDict"]["timeProperties"]["decompositionInterval"] = 20
    return case, ld


def run_tutorial(Cloud, steps, seed=61):
    case, ld = tutorial_case(seed)
    cl = case.make_cloud(Cloud, parcelCapacity=8 * case.n_parcels, sampleInterval=foamdict.sample_interval(ld["fieldPropertiesDict"]))
    cl.setHybridDecomposition(ld["hybridDecompositionDict"])
    ad = UniGasDynamicAdapter(cl, case.uniGasProperties)
    if case.subCellLevels is not None:
        ad.subCellLevels = case.subCellLevels.copy()
    tally = dict(inserted=0, deleted=0, wallHits=0, collisions=0, bgkRelaxations=0, cloned=0, weightDeleted=0)
    done = 0
    while done < steps:
        ad.run(20)
        done += 20
        ad.cellCollModelId = cl.hybridDecomposition()["cellCollModelId"]
        c = cl.counters()
        for k in tally:
            tally[k] += c[k]
    return case, cl, ad, tally


def test_oracle_runs_the_tutorial_case(OracleCloud):
    case, cl, ad, tally = run_tutorial(OracleCloud, 100)
    assert case.uniGasProperties["cellWeightedSimulation"] and case.deltaT > 5e-8   # setInitialConfiguration: Courant-limited on this mesh
    cnt0 = np.bincount(case.cell, minlength=case.mesh.n_cells)
    nSub = np.ones(case.mesh.n_cells) if case.subCellLevels is None else case.subCellLevels.prod(1)
    assert abs((cnt0 / nSub).mean() - 20) < 1  # uniGasMeshFill's weight rule: particlesPerSubCell parcels in every sub-cell
    c = cl.counters()
    assert c["stuck"] == 0 and c["step"] == 100
    assert all(tally[k] > 0 for k in ("inserted", "deleted", "wallHits", "cloned", "weightDeleted"))
    assert tally["collisions"] + tally["bgkRelaxations"] > 0
    assert cl.cfg.deltaT != case.deltaT  # timeStepAdaptation has acted
    f = cl.fields()
    assert np.isfinite(f["rhoN"]).all() and np.isfinite(f["translationalT"]).all()
    # free stream is still free stream upstream; gas piles up and heats in front of the cylinder
    x, y = case.mesh.cell_centres[:, 0], case.mesh.cell_centres[:, 1]
    r = np.hypot(x, y)
    up = (x < 0) & (r > 0.5)
    nose = (x < 0) & (np.abs(y) < 0.06) & (r < 0.19)
    assert abs(np.median(f["rhoN"][up]) / 4.247e20 - 1) < 0.15
    assert f["rhoN"][nose].mean() > 1.5 * 4.247e20 and f["translationalT"][nose].mean() > 2 * 200


@pytest.mark.gpu
def test_gpu_runs_the_tutorial_case_like_the_oracle(GpuCloud, OracleCloud):
    _, g, ag, tg = run_tutorial(GpuCloud, 60)
    case, r, ar, tr = run_tutorial(OracleCloud, 60)
    for k in ("inserted", "deleted", "wallHits"):
        assert abs(tg[k] - tr[k]) <= 0.1 * tr[k] + 10, (k, tg[k], tr[k])
    assert abs(g.counters()["nParcels"] / r.counters()["nParcels"] - 1) < 0.05
    assert g.cfg.deltaT == pytest.approx(r.cfg.deltaT, rel=0.05)
    fg, fr = g.fields(), r.fields()
    dens = lambda f: np.average(f["rhoN"], weights=case.mesh.cell_volumes)
    assert dens(fg) == pytest.approx(dens(fr), rel=0.03)
    assert g.counters()["stuck"] == 0


GOLD = os.path.dirname(CASE)


@pytest.mark.parametrize("name", ["hypersonicCylinder", "supersonicPlate", "plumeImpingement", "expansionInVacuum"])
def test_all_tutorial_dictionaries_parse(name):
    """Every case the reference ships: the dictionaries load, the run-time-selected words are known to the tables of
    unigasfoam_b200._capi, and what the B200 path does not cover yet is visible as such."""
    from unigasfoam_b200 import _capi
    ld = foamdict.load_case(os.path.join(GOLD, name))
    p = ld["uniGasProperties"]
    assert p["typeIdList"] == ["Ar"] and p["moleculeProperties"]["Ar"]["mass"] == 66.3e-27
    assert p["collisionModel"] in _capi.COLLISION_MODEL and p["bgkCollisionModel"] in _capi.BGK_MODEL
    assert p["dsmcCollisionModel"] in _capi.BINARY_MODEL and p["dsmcCollisionPartnerModel"] in _capi.PARTNER_MODEL
    assert p["cellWeightedSimulation"] is True and p["adaptiveSimulation"] is True
    assert ld["deltaT"] > 0
    known_general = {"uniGasFreeStreamInflowPatch", "uniGasLiouFangPressureInletPatch"}
    for e in ld["boundariesDict"]["uniGasPatchBoundaries"]:
        assert e["boundaryModel"] in _capi.WALL_MODEL, e["boundaryModel"]
        assert "patch" in e["patchBoundaryProperties"]
    for e in ld["boundariesDict"]["uniGasGeneralBoundaries"]:
        assert e["boundaryModel"] in known_general, e["boundaryModel"]
    assert ld["hybridDecompositionDict"]["decompositionModel"] == "localKnudsen"
    models = {f["fieldModel"] for f in ld["fieldPropertiesDict"]["uniGasFields"]}
    assert models <= {"uniGasVolFields", "uniGasMassFluxSurface", "uniGasForceSurface"}
    axi = name in ("plumeImpingement", "expansionInVacuum")
    assert p["axisymmetricSimulation"] is axi
    if axi:
        assert p["axisymmetricProperties"]["maxRadialWeightingFactor"] > 1


def test_unsupported_switches_fail_loudly(OracleCloud):
    """chemicalReactions (out of scope) says so instead of running something else; the axisymmetric tutorials'
    uniGasProperties construct a cloud as written (radial weighting: tests/test_axisymmetric.py)."""
    from unigasfoam_b200.cloud import UgfError
    m = cases.closed_box(n=3, parcels=100).mesh
    ld = foamdict.load_case(os.path.join(GOLD, "plumeImpingement"))
    cl = OracleCloud(m, ld["uniGasProperties"], {}, ld["deltaT"], parcelCapacity=1000)
    assert cl.axisymmetric and cl.cfg.maxRWF == 1000.0 and cl.cfg.radialExtent == 7.5e-2
    cl.close()
    with pytest.raises(UgfError, match="chemicalReactions"):
        OracleCloud(m, dict(ld["uniGasProperties"], chemicalReactions=True), {}, ld["deltaT"], parcelCapacity=1000)
    # macroInterpolation true (set by every tutorial) is covered: tests/test_macro_interpolation.py runs the tutorial's
    # collisionProperties as written
    ld = foamdict.load_case(CASE)
    assert ld["uniGasProperties"]["collisionProperties"]["macroInterpolation"] is True
    OracleCloud(m, ld["uniGasProperties"], {}, ld["deltaT"], parcelCapacity=1000).close()


# ---- the supersonicPlate tutorial on its own mesh layout; macroInterpolation (cell values) is the one override ------
PLATE = os.path.join(GOLD, "supersonicPlate")


def run_plate(Cloud, steps, seed=71):
    """tutorials/uniGasFoam/supersonicPlate (hybrid USP-SBGK / NTC-VHS, cell weighting, time-step / sub-cell /
    cell-weight adaptation every 10 steps, local-Knudsen decomposition, free-stream inflow at Mach 1.5, diffuse plate) on
    the tutorial's three-block mesh at a fifth of its resolution."""
    m = ugmesh.plate_mesh(nx=(10, 20, 15), ny=40)
    case, ld = cases.from_case_dir(PLATE, m, seed=seed, overrides={"collisionProperties": {"macroInterpolation": False}})
    cl = case.make_cloud(Cloud, parcelCapacity=3 * case.n_parcels, sampleInterval=foamdict.sample_interval(ld["fieldPropertiesDict"]))
    cl.setHybridDecomposition(ld["hybridDecompositionDict"])
    ad = UniGasDynamicAdapter(cl, case.uniGasProperties)
    if case.subCellLevels is not None:
        ad.subCellLevels = case.subCellLevels.copy()
    tally = dict(inserted=0, deleted=0, wallHits=0, collisions=0, bgkRelaxations=0, cloned=0, weightDeleted=0)
    for _ in range(steps // 10):
        ad.run(10)
        c = cl.counters()
        for k in tally:
            tally[k] += c[k]
    return case, ld, cl, ad, tally


def test_oracle_runs_supersonic_plate(OracleCloud):
    case, ld, cl, ad, tally = run_plate(OracleCloud, 30)
    p = case.uniGasProperties
    assert p["collisionModel"] == "hybrid" and p["adaptiveProperties"]["adaptationInterval"] == 10
    assert p["collisionProperties"]["theta"] == 0.1 and p["collisionProperties"]["Tref"] == 273
    assert ld["deltaT"] == 5e-9 and case.deltaT > 5 * ld["deltaT"]         # setInitialConfiguration raised the time step
    assert case.subCellLevels is not None and case.subCellLevels.max() > 1 and (case.subCellLevels[:, 2] == 1).all()
    cnt0 = np.bincount(case.cell, minlength=case.mesh.n_cells)
    nSub = case.subCellLevels.prod(1)
    assert abs((cnt0 / nSub).mean() - 20) < 1                              # particlesPerSubCell parcels in every sub-cell
    c = cl.counters()
    assert c["stuck"] == 0 and c["step"] == 30
    assert tally["inserted"] > 0 and tally["deleted"] > 0 and tally["wallHits"] > 0
    f = cl.fields()
    assert np.isfinite(f["rhoN"]).all()
    x, y = case.mesh.cell_centres[:, 0], case.mesh.cell_centres[:, 1]
    free = (y > 3e-3) & (x < 2e-3)
    assert abs(np.median(f["rhoN"][free]) / 3.416e22 - 1) < 0.05
    assert abs(np.median(f["UMean"][free, 0]) / 484.0 - 1) < 0.05
    near = (y < 1.5e-4) & (x > 0.7e-3) & (x < 1.4e-3)                      # the gas right above the plate has been slowed down
    assert f["UMean"][near, 0].mean() < 0.8 * 484.0
    assert np.abs(f["surfaceShearStress"]).max() > 0


@pytest.mark.gpu
def test_gpu_runs_supersonic_plate_like_the_oracle(GpuCloud, OracleCloud):
    case, _, g, ag, tg = run_plate(GpuCloud, 20)
    _, _, r, ar, tr = run_plate(OracleCloud, 20)
    for k in ("inserted", "deleted", "wallHits"):
        assert abs(tg[k] - tr[k]) <= 0.05 * tr[k] + 10, (k, tg[k], tr[k])
    assert abs(g.counters()["nParcels"] / r.counters()["nParcels"] - 1) < 0.02
    assert g.cfg.deltaT == pytest.approx(r.cfg.deltaT, rel=0.05)
    assert np.array_equal(ag.subCellLevels, ar.subCellLevels) or (np.abs(ag.subCellLevels - ar.subCellLevels) <= 1).all()
    fg, fr = g.fields(), r.fields()
    V = case.mesh.cell_volumes
    assert np.average(fg["rhoN"], weights=V) == pytest.approx(np.average(fr["rhoN"], weights=V), rel=0.01)
    assert np.average(fg["UMean"][:, 0], weights=V) == pytest.approx(np.average(fr["UMean"][:, 0], weights=V), rel=0.02)
    assert g.counters()["stuck"] == 0


def test_mesh_field_fill_from_a_time_directory(tmp_path, OracleCloud):
    """`type uniGasMeshFieldFill` (the initialisation of the plume / expansion tutorials): every cell is filled at the
    state the fields of the start time directory give it (numberDensity_<species>, transT, rotT, U)."""
    import shutil
    from unigasfoam_b200 import foamfile
    case_dir = tmp_path / "fieldFill"
    shutil.copytree(CASE, case_dir)
    (case_dir / "system" / "uniGasInitialisationDict").write_text(
        "FoamFile { version 2.0; format ascii; class dictionary; object uniGasInitialisationDict; }\n"
        "configurations ( configuration { type uniGasMeshFieldFill; typeIdList (Ar); } );\n")
    m = ugmesh.half_annulus_mesh(10, 16, 0.5 * 0.3048, 2.0 * 0.3048, 0.1 * 0.3048, 3.0)
    m.meta_axis_aligned = False
    x = m.cell_centres[:, 0]
    n = 4.247e20 * (1.0 + 0.8 * (x - x.min()) / (x.max() - x.min()))
    T = 200.0 + 300.0 * (x > 0)
    U = np.column_stack([1000.0 * (x < 0), np.zeros(m.n_cells), np.zeros(m.n_cells)])
    os.makedirs(case_dir / "0")
    patches = {p.name: ("empty" if p.kind == "empty" else "calculated") for p in m.patches}
    foamfile.write_vol_field(str(case_dir / "0" / "numberDensity_Ar"), "0", [0, -3, 0, 0, 0, 0, 0], n, patches)
    foamfile.write_vol_field(str(case_dir / "0" / "transT"), "0", [0, 0, 0, 1, 0, 0, 0], T, patches)
    foamfile.write_vol_field(str(case_dir / "0" / "rotT"), "0", [0, 0, 0, 1, 0, 0, 0], np.zeros(m.n_cells), patches)
    foamfile.write_vol_field(str(case_dir / "0" / "U"), "0", [0, 1, -1, 0, 0, 0, 0], U, patches, vector=True)
    FN = 4.247e20 * m.cell_volumes.sum() / 40000
    case, _ = cases.from_case_dir(str(case_dir), m, seed=5, overrides={
        "nEquivalentParticles": FN, "cellWeightedSimulation": False, "adaptiveSimulation": False, "collisionModel": "dsmc",
        "collisionProperties": {"macroInterpolation": False}})
    cnt = np.bincount(case.cell, minlength=m.n_cells)
    expect = n * m.cell_volumes / FN
    assert abs(cnt.sum() / expect.sum() - 1) < 0.01 and np.corrcoef(cnt, expect)[0, 1] > 0.97
    Tp = case.meta["species"]["mass"] * ((case.U - U[case.cell]) ** 2).sum(1) / (3 * cases.kB)
    assert abs(Tp[x[case.cell] > 0].mean() / 500.0 - 1) < 0.03 and abs(Tp[x[case.cell] < 0].mean() / 200.0 - 1) < 0.03
    assert abs(case.U[x[case.cell] < 0, 0].mean() / 1000.0 - 1) < 0.02 and abs(case.U[x[case.cell] > 0, 0].mean()) < 10.0
    cl = case.make_cloud(OracleCloud, parcelCapacity=3 * case.n_parcels)
    cl.evolve(3)
    assert cl.counters()["stuck"] == 0


def test_decompose_par_dict_to_partition():
    """system/decomposeParDict of the hypersonicCylinder tutorial (10 subdomains, scotch, weightField commented out) ->
    a cell-to-rank map; with the weightField switched on the slabs are cut at equal weight."""
    import numpy as np
    from unigasfoam_b200 import mesh as ugmesh
    d = foamdict.read(os.path.join(GOLD, "hypersonicCylinder", "system", "decomposeParDict"))
    assert d["numberOfSubdomains"] == 10 and d["method"] == "scotch" and "weightField" not in d
    m = ugmesh.half_annulus_mesh(40, 20, 0.1, 0.5, 0.01, grading=4.0)
    part, n = foamdict.partition_from_dict(m, d)
    assert n == 10 and set(part) == set(range(10)) and np.array_equal(part, ugmesh.slab_partition(m, 10, 0))
    w = np.exp(-8.0 * (np.hypot(m.cell_centres[:, 0], m.cell_centres[:, 1]) - 0.1))
    with pytest.raises(foamdict.FoamDictError, match="weightField"):
        foamdict.partition_from_dict(m, dict(d, weightField="uniGasRhoNMean_Ar"))
    wp, _ = foamdict.partition_from_dict(m, dict(d, weightField="uniGasRhoNMean_Ar"), weights=w)
    per = lambda p: np.bincount(p, weights=w, minlength=10)
    assert ugmesh.load_imbalance(per(wp)) < ugmesh.load_imbalance(per(part))
    simple = {"numberOfSubdomains": 4, "method": "simple", "simpleCoeffs": {"n": [1, 4, 1]}}
    sp, _ = foamdict.partition_from_dict(m, simple)
    assert np.array_equal(sp, ugmesh.slab_partition(m, 4, 1))
    bp, nb = foamdict.partition_from_dict(m, {"numberOfSubdomains": 4, "method": "simple", "simpleCoeffs": {"n": [2, 2, 1]}})
    assert nb == 4 and (np.bincount(bp) == m.n_cells // 4).all()
    blocks = ugmesh.decompose(m, bp, 4)
    assert sum(s.n_cells for s in blocks) == m.n_cells and all(sum(p.kind == "processor" for p in s.patches) >= 2 for s in blocks)
    with pytest.raises(foamdict.FoamDictError, match="one direction"):
        foamdict.partition_from_dict(m, {"numberOfSubdomains": 4, "method": "simple", "simpleCoeffs": {"n": [2, 2, 1]}, "weightField": "w"}, weights=w)
    with pytest.raises(foamdict.FoamDictError, match="multiply"):
        foamdict.partition_from_dict(m, {"numberOfSubdomains": 4, "method": "simple", "simpleCoeffs": {"n": [2, 3, 1]}})
    assert sum(s.n_cells for s in ugmesh.decompose(m, wp, 10)) == m.n_cells


def test_oracle_solver_loop_writes_what_the_reference_leaves(tmp_path, OracleCloud):
    """unigasfoam_b200.solver.run_case: the uniGasFoam time loop on the tutorial's own dictionaries - write times from
    controlDict, fields / accumulator dictionaries per fieldPropertiesDict entry, restart from latestTime."""
    from unigasfoam_b200 import foamfile, solver
    m = ugmesh.half_annulus_mesh(12, 16, 0.5 * 0.3048, 2.0 * 0.3048, 0.1 * 0.3048, 5.0)
    m.meta_axis_aligned = False
    ov = {"collisionProperties": {"macroInterpolation": False}, "adaptiveProperties": {"maxSubCellSizeMFPRatio": 8.0, "adaptationInterval": 4}}
    out = str(tmp_path)
    # endTime in seconds: ten steps of the (adapted) time step are not known beforehand, so drive by writeControl timeStep and stop by time
    case0, _ = cases.from_case_dir(CASE, m, overrides=ov, particles_per_cell=8)
    end = 10.5 * case0.deltaT
    lines = []
    r = solver.run_case(CASE, m, OracleCloud, out_dir=out, overrides=ov, particles_per_cell=8,
                        control={"writeControl": "timeStep", "writeInterval": 5, "endTime": end, "startFrom": "startTime"}, log=lines.append)
    # nTerminalOutputs 10 in the tutorial's controlDict: one block of info lines per ten steps
    assert r["steps"] >= 10 and len(r["written"]) >= 2
    assert sum(l.startswith("Time = ") for l in lines) == r["steps"] // 10 and any("Number of particles" in l for l in lines)
    first = os.path.join(out, r["written"][0])
    have = set(os.listdir(first))
    # cleanLagrangian: only the latest write keeps its parcels
    assert "lagrangian" not in have and os.path.isdir(os.path.join(out, r["written"][-1], "lagrangian"))
    assert {"uniform", "uniGasSigmaTcRMax", "uniGasCellWeightFactor", "uniGasSubCellLevels", "uniGasCollisionModelId",
            "rhoN_Ar", "p_Ar", "UMean_Ar", "translationalT_Ar", "surfaceHeatTransfer_Ar", "fD_Ar", "variableHardSphereMeanFreePath_Ar",
            "densityError_Ar"} <= have
    assert {"time", "volFieldsMethod_Ar"} <= set(os.listdir(os.path.join(first, "uniform")))
    rho = foamfile.read_vol_field(os.path.join(first, "rhoN_Ar"))
    assert rho["dimensions"] == [0, -3, 0, 0, 0, 0, 0] and np.median(foamfile.expand_internal(rho, m.n_cells)) > 1e20
    # resetAtOutput on (until 2e-3 s): the fields of a write average over the steps since the previous one, and the
    # accumulator dictionary, written after the reset (uniGasVolFields.C:1433-1507), starts again from zero
    from unigasfoam_b200 import volfields_io
    d2 = volfields_io.read_volfields_method(os.path.join(out, r["written"][1], "uniform", "volFieldsMethod_Ar"))
    assert d2["nTimeSteps"] == 0 and not d2["rhoNMean"].any()
    # startFrom latestTime: the run picks up the last time directory and goes on from its step index
    n_before = r["cloud"].counters()["step"]
    r2 = solver.run_case(CASE, m, OracleCloud, out_dir=out, overrides=ov, particles_per_cell=8,
                         control={"writeControl": "timeStep", "writeInterval": 5, "endTime": r["time"] + 3.5 * r["cloud"].cfg.deltaT,
                                  "startFrom": "latestTime"})
    assert r2["steps"] in (3, 4, 5) and r2["cloud"].counters()["step"] == n_before + r2["steps"]
    assert solver.latest_time(out)[1] == r2["written"][-1] and solver.time_name(0.0) == "0" and solver.time_name(5e-4) == "0.0005"


@pytest.mark.gpu
def test_gpu_solver_loop_on_the_tutorial_case(tmp_path, GpuCloud):
    """The same loop on libugf: the tutorial case runs from its dictionaries to a write time and restarts from it."""
    from unigasfoam_b200 import foamfile, solver
    m = ugmesh.half_annulus_mesh(12, 16, 0.5 * 0.3048, 2.0 * 0.3048, 0.1 * 0.3048, 5.0)
    m.meta_axis_aligned = False
    ov = {"collisionProperties": {"macroInterpolation": False}, "adaptiveProperties": {"maxSubCellSizeMFPRatio": 8.0, "adaptationInterval": 4}}
    case0, _ = cases.from_case_dir(CASE, m, overrides=ov, particles_per_cell=8)
    ctl = {"writeControl": "timeStep", "writeInterval": 5, "endTime": 10.5 * case0.deltaT, "startFrom": "startTime"}
    r = solver.run_case(CASE, m, GpuCloud, out_dir=str(tmp_path), overrides=ov, particles_per_cell=8, control=ctl)
    assert r["steps"] >= 10 and len(r["written"]) >= 2 and r["cloud"].counters()["stuck"] == 0
    last = os.path.join(str(tmp_path), r["written"][-1])
    rho = foamfile.expand_internal(foamfile.read_vol_field(os.path.join(last, "rhoN_Ar")), m.n_cells)
    assert np.isfinite(rho).all() and np.median(rho) > 1e20
    assert os.path.exists(os.path.join(last, "uniform", "volFieldsMethod_Ar")) and os.path.exists(os.path.join(last, "lagrangian", "uniGas", "positions"))
    n_before = r["cloud"].counters()["step"]
    r2 = solver.run_case(CASE, m, GpuCloud, out_dir=str(tmp_path), overrides=ov, particles_per_cell=8,
                         control=dict(ctl, startFrom="latestTime", endTime=r["time"] + 3.5 * r["cloud"].cfg.deltaT))
    assert r2["cloud"].counters()["step"] == n_before + r2["steps"] and r2["steps"] >= 3
