"""<time>/uniform/volFieldsMethod_<fieldName>: uniGasVolFields' accumulator dictionary under the
reference's entry names (writeOut / readIn, uniGasVolFields.C:549-669).  Checked: the entry list and list shapes are the
reference's, write -> read returns the same bits, the counted-list grammar OpenFOAM itself emits (N{v} uniform lists,
0(), white space) is read, and a run whose averages were carried over through the dictionary alone reports the same
fields as the uninterrupted one."""
import os

import numpy as np
import pytest

from unigasfoam_b200 import cases, volfields_io as vio

# the entries of uniGasVolFields::writeOut in its order (uniGasVolFields.C:625-662)
REFERENCE_ENTRIES = [
    "nTimeSteps", "timeCounter", "rhoNMean", "rhoNMeanXnParticle", "rhoNMeanInt", "molsElec", "rhoMMean", "rhoMMeanXnParticle",
    "linearKEMean", "linearKEMeanXnParticle", "rotationalEMean", "rotationalDofMean", "momentumMean", "momentumMeanXnParticle",
    "vibrationalETotal", "electronicETotal", "nParcels", "nParcelsXnParticle", "mccSpecies", "nGroundElectronicLevel",
    "nFirstElectronicLevel", "mfp", "mcr", "rhoNBF", "rhoMBF", "linearKEBF", "rotationalEBF", "rotationalDofBF", "qBF",
    "totalvDofBF", "speciesRhoNIntBF", "speciesRhoNElecBF", "momentumBF", "fDBF", "vibrationalEBF", "electronicEBF",
    "speciesRhoNBF", "mccSpeciesBF"]


def test_counted_list_grammar():
    d = vio.parse_volfields_method("""
        FoamFile { version 2.0; format ascii; class dictionary; location "1/uniform"; object volFieldsMethod_Ar; }
        nTimeSteps 12; timeCounter 1.5e-05; // comment
        a 4(1 2.5 -3e-2 4);
        u 5{0.25};
        v 2((1 2 3) (4 5 6));
        w 3{(1 0 -1)};
        bf 3(2(1 2) 0() 4{7});
        vbf 2(2((1 2 3)
                 (4 5 6)) 1((7 8 9)));
        deep 2(1(2(5 6)) 1(0()));
        empty 0();
    """)
    assert d["nTimeSteps"] == 12 and d["timeCounter"] == 1.5e-05
    assert np.array_equal(d["a"], [1, 2.5, -0.03, 4]) and np.array_equal(d["u"], np.full(5, 0.25))
    assert np.array_equal(d["v"], [[1, 2, 3], [4, 5, 6]]) and np.array_equal(d["w"], [[1, 0, -1]] * 3)
    assert [list(x) for x in d["bf"]] == [[1, 2], [], [7, 7, 7, 7]]
    assert np.array_equal(d["vbf"][0], [[1, 2, 3], [4, 5, 6]]) and np.array_equal(d["vbf"][1], [[7, 8, 9]])
    assert list(d["deep"][0][0]) == [5, 6] and len(d["deep"][1][0]) == 0 and len(d["empty"]) == 0
    with pytest.raises(vio.VolFieldsFormatError, match="announces 3"):
        vio.parse_volfields_method("FoamFile { format ascii; }\na 3(1 2);")
    with pytest.raises(vio.VolFieldsFormatError, match="missing ;"):
        vio.parse_volfields_method("FoamFile { format ascii; }\na 2(1 2)\nb 1;")
    with pytest.raises(Exception, match="ascii"):
        vio.parse_volfields_method("FoamFile { format binary; }\na 2(1 2);")


def test_field_names_from_the_tutorial_dictionary(tmp_path, OracleCloud):
    from unigasfoam_b200 import foamdict
    fp = foamdict.read(os.path.join(os.path.dirname(__file__), "golden", "openfoam", "hypersonicCylinder", "system", "fieldPropertiesDict"))
    assert foamdict.vol_field_names(fp) == ["Ar"] and foamdict.vol_field_names(fp, averaging_only=True) == ["Ar"]
    case = _couette()
    a = case.make_cloud(OracleCloud)
    a.evolve(2)
    t = a.writeTime(str(tmp_path), "2", fieldNames=foamdict.vol_field_names(fp))
    assert sorted(os.listdir(os.path.join(t, "uniform"))) == ["time", "ugfState.npy", "volFieldsMethod_Ar"]


def _couette(**kw):
    return cases.couette(nx=12, ny=8, ppc=30, seed=5, **kw)


def test_entries_shapes_and_bit_exact_round_trip(tmp_path, OracleCloud):
    case = _couette()
    a = case.make_cloud(OracleCloud)
    a.evolve(6)
    path = a.writeVolFieldsMethod(str(tmp_path), "6e-06", "Ar")
    assert path.endswith(os.path.join("6e-06", "uniform", "volFieldsMethod_Ar"))
    head = open(path).read(900)
    assert "class       dictionary;" in head and 'location    "6e-06/uniform";' in head and "object      volFieldsMethod_Ar;" in head
    d = vio.read_volfields_method(path)
    assert list(d) == REFERENCE_ENTRIES
    m = case.mesh
    nC, nP = m.n_cells, len(m.patches)
    assert d["nTimeSteps"] == 6 and d["timeCounter"] == pytest.approx(6 * case.deltaT, rel=1e-14)
    assert d["rhoNMean"].shape == (nC,) and d["momentumMean"].shape == (nC, 3)
    assert len(d["nParcels"]) == 1 and d["nParcels"][0].shape == (nC,) and len(d["vibrationalETotal"][0]) == 0
    assert len(d["rhoNBF"]) == nP and [len(x) for x in d["rhoNBF"]] == [p.size for p in m.patches]
    assert [np.shape(x) for x in d["fDBF"]] == [(p.size, 3) if p.size else (0,) for p in m.patches]
    assert len(d["speciesRhoNBF"]) == 1 and len(d["speciesRhoNBF"][0]) == nP
    # the numbers are the accumulators, bit for bit
    acc = a.accumulators()
    assert np.array_equal(d["rhoNMean"], acc["acc"][:, 0]) and np.array_equal(d["linearKEMeanXnParticle"], acc["acc"][:, 13])
    assert np.array_equal(d["momentumMeanXnParticle"], acc["acc"][:, 10:13]) and np.array_equal(d["nParcelsXnParticle"][0], acc["species"][:, 0])
    assert np.array_equal(d["nParcels"][0], d["rhoNMean"]) and np.array_equal(d["mccSpecies"][0], d["linearKEMean"])
    # every parcel of one species: mass x number, and the diffuse walls did tally something
    mAr = case.uniGasProperties["moleculeProperties"]["Ar"]["mass"]
    assert np.allclose(d["rhoMMean"], mAr * d["rhoNMean"], rtol=1e-14)
    walls = [i for i, p in enumerate(m.patches) if p.kind == "wall"]
    assert walls and all(d["rhoNBF"][i].sum() > 0 and np.abs(d["qBF"][i]).sum() > 0 for i in walls)
    assert all(np.array_equal(d["mccSpeciesBF"][0][i], 2.0 * d["linearKEBF"][i]) for i in walls)
    # back into a state array: the same bits, slot 15 rebuilt to round-off
    s0 = a.state()
    blank = s0.copy()
    v = vio.state_views(blank)
    v["acc"][:] = 0; v["accSpecies"][:] = 0; v["bacc"][:, :13] = 0; v["scalars"][1:3] = 0
    s1 = vio.apply_volfields_method(d, m, blank)
    v0, v1 = vio.state_views(s0), vio.state_views(s1)
    assert np.array_equal(v0["acc"][:, :15], v1["acc"][:, :15]) and np.allclose(v0["acc"][:, 15], v1["acc"][:, 15], rtol=1e-14)
    assert np.array_equal(v0["accSpecies"], v1["accSpecies"]) and np.array_equal(v0["bacc"][:, :13], v1["bacc"][:, :13])
    assert np.array_equal(v0["scalars"], v1["scalars"])
    with pytest.raises(vio.VolFieldsFormatError, match="patches"):
        vio.apply_volfields_method(dict(d, rhoNBF=d["rhoNBF"][:-1]), m, s0)
    with pytest.raises(vio.VolFieldsFormatError, match="cells"):
        vio.apply_volfields_method(dict(d, rhoNMean=d["rhoNMean"][:-1]), m, s0)


def _averages_continue_through_the_dictionary(Cloud, tmp_path, case):
    """averagingAcrossManyRuns: a restarted run (parcels + cell state from the time directory, averages from the
    dictionary only - ugfState.npy removed) reports the fields of the uninterrupted run."""
    a = case.make_cloud(Cloud)
    a.evolve(5)
    t = a.writeTime(str(tmp_path), "5")
    a.writeVolFieldsMethod(str(tmp_path), "5", "Ar")
    os.remove(os.path.join(t, "uniform", "ugfState.npy"))
    a.evolve(4)
    b = Cloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT)
    b.readTime(str(tmp_path), "5")
    assert b.accumulators()["nAvTimeSteps"] == 0  # nothing carried over yet
    assert b.readVolFieldsMethod(str(tmp_path), "5", "missing") is None
    d = b.readVolFieldsMethod(str(tmp_path), "5", "Ar")
    assert d["nTimeSteps"] == 5 and b.accumulators()["nAvTimeSteps"] == 5
    b.evolve(4)
    fa, fb = a.fields(), b.fields()
    for k in fa:  # atol: the device tallies wall hits with atomics, sums that cancel (wall momentum) differ in the last bits per run
        assert np.allclose(fa[k], fb[k], rtol=1e-9, atol=1e-9 * np.abs(fa[k]).max()), k
    assert b.accumulators()["nAvTimeSteps"] == 9
    return fa


def test_oracle_averages_continue_through_the_dictionary(tmp_path, OracleCloud):
    f = _averages_continue_through_the_dictionary(OracleCloud, tmp_path, _couette(binary="noDSMCCollision"))
    assert (f["rhoN"] > 0).all() and (f["wall_rhoN"] > 0).any()


def test_mixture_lists_per_species(tmp_path, OracleCloud):
    case = cases.mixture_box(n=4, parcels=6000, seed=3)
    a = case.make_cloud(OracleCloud)
    a.evolve(3)
    d = vio.read_volfields_method(a.writeVolFieldsMethod(str(tmp_path), "3", "mixture"))
    acc = a.accumulators()
    assert len(d["nParcelsXnParticle"]) == 2 and len(d["electronicETotal"]) == 2 and len(d["mccSpeciesBF"]) == 2
    for s in range(2):
        assert np.array_equal(d["nParcelsXnParticle"][s], acc["species"][:, s])
    assert np.allclose(d["nParcelsXnParticle"][0] + d["nParcelsXnParticle"][1], d["rhoNMeanXnParticle"], rtol=1e-13)
    s1 = vio.apply_volfields_method(d, case.mesh, a.state())
    assert np.array_equal(vio.state_views(s1)["bacc"], vio.state_views(a.state())["bacc"])  # slot 14 left alone for mixtures


@pytest.mark.gpu
def test_gpu_averages_continue_through_the_dictionary(tmp_path, GpuCloud):
    _averages_continue_through_the_dictionary(GpuCloud, tmp_path, _couette(binary="noDSMCCollision"))


@pytest.mark.gpu
def test_gpu_reads_the_dictionary_the_oracle_wrote(tmp_path, GpuCloud, OracleCloud):
    """The dictionary as exchange format: oracle accumulators -> file -> libugf state -> the same fields."""
    case = _couette()
    r = case.make_cloud(OracleCloud)
    r.evolve(4)
    r.writeVolFieldsMethod(str(tmp_path), "4", "Ar")
    g = case.make_cloud(GpuCloud)
    g.readVolFieldsMethod(str(tmp_path), "4", "Ar")
    fr, fg = r.fields(), g.fields()
    for k in fr:
        assert np.allclose(fr[k], fg[k], rtol=1e-12, atol=1e-12 * np.abs(fr[k]).max()), k


def test_output_fields_under_the_reference_names(tmp_path, OracleCloud):
    """uniGasVolFields at write time (uniGasVolFields.C:67-357, 1256-1428): rhoN_<f>, p_<f>, UMean_<f>, ... as volFields with
    the reference's dimensions; wall measurements on the wall patches, constraint patches by their type."""
    from unigasfoam_b200 import foamfile
    case = _couette()
    a = case.make_cloud(OracleCloud)
    a.evolve(6)
    files = a.writeFields(str(tmp_path), "6e-06", "Ar")
    names = sorted(os.path.basename(p) for p in files)
    want = ["uniGasRhoNMean", "rhoN", "rhoM", "p", "translationalT", "rotationalT", "vibrationalT", "electronicT", "overallT",
            "surfaceHeatTransfer", "surfaceShearStress", "Ma", "UMean", "fD", "variableHardSphereMeanFreePath", "subCellSizeMFPRatio",
            "meanCollisionRate", "meanCollisionTime", "timeStepMCTRatio", "densityError", "velocityError", "temperatureError", "pressureError"]
    assert names == sorted(n + "_Ar" for n in want)
    assert len(a.writeFields(str(tmp_path), "6e-06", "x", measureMeanFreePath=False, measureErrors=False)) == 14
    f, m = a.fields(), case.mesh
    nI = m.n_internal
    rd = lambda n: foamfile.read_vol_field(str(tmp_path / "6e-06" / (n + "_Ar")))
    p = rd("p")
    assert p["class"] == "volScalarField" and p["dimensions"] == [1, -1, -2, 0, 0, 0, 0] and np.array_equal(p["internal"], f["p"])
    u = rd("UMean")
    assert u["class"] == "volVectorField" and u["dimensions"] == [0, 1, -1, 0, 0, 0, 0] and np.array_equal(u["internal"], f["UMean"])
    assert rd("rhoM")["dimensions"] == [1, -3, 0, 0, 0, 0, 0] and rd("surfaceHeatTransfer")["dimensions"] == [1, 0, -3, 0, 0, 0, 0]
    assert rd("variableHardSphereMeanFreePath")["dimensions"] == [0, 1, 0, 0, 0, 0, 0] and rd("meanCollisionRate")["dimensions"] == [0, 0, -1, 0, 0, 0, 0]
    q, fd, rn = rd("surfaceHeatTransfer"), rd("fD"), rd("rhoN")
    assert q["uniform"] and q["internal"] == 0.0
    seen_wall = 0
    for pt in m.patches:
        sl = slice(pt.start - nI, pt.start - nI + pt.size)
        if pt.kind == "wall":
            seen_wall += 1
            assert p["boundary"][pt.name]["type"] == "calculated"
            assert np.array_equal(foamfile.boundary_values(p, pt.name, pt.size), f["wall_p"][sl])
            assert np.array_equal(foamfile.boundary_values(q, pt.name, pt.size), f["surfaceHeatTransfer"][sl])
            assert np.array_equal(foamfile.boundary_values(fd, pt.name, pt.size), f["fD"][sl])
            assert np.array_equal(foamfile.boundary_values(u, pt.name, pt.size), f["wall_UMean"][sl])
            own = np.asarray(m.owner[pt.start:pt.start + pt.size])
            assert np.array_equal(foamfile.boundary_values(rn, pt.name, pt.size), f["rhoN"][own])  # :1370-1372: the cell value
        else:
            assert p["boundary"][pt.name] == {"type": pt.kind}
    assert seen_wall == 2 and np.abs(f["wall_p"]).max() > 0
    # measureErrors (:1246-1250): pressureError = sqrt(gamma) densityError, argon gamma = 5/3
    ok = f["densityError"] > 0
    assert ok.any() and np.allclose(f["pressureError"][ok], np.sqrt(5.0 / 3.0) * f["densityError"][ok], rtol=1e-12)
    assert np.array_equal(rd("pressureError")["internal"], f["pressureError"])


def test_wall_rotational_and_overall_temperature(tmp_path, OracleCloud):
    """rotationalT / overallT on wall patches (uniGasVolFields.C:1299-1330) from the wall accumulators: nitrogen between
    diffuse 300 K walls - incident and re-emitted molecules both carry about k T per molecule of rotational energy."""
    from unigasfoam_b200 import foamfile
    case = cases.closed_box(n=5, parcels=40000, seed=13, wall="diffuse", species=("N2", cases.NITROGEN),
                            binary="LarsenBorgnakkeVariableHardSphere", Trot=300.0)
    a = case.make_cloud(OracleCloud)
    a.evolve(30)
    a.writeFields(str(tmp_path), "30", "N2")
    m = case.mesh
    rot, ov, tr = (foamfile.read_vol_field(str(tmp_path / "30" / (n + "_N2"))) for n in ("rotationalT", "overallT", "translationalT"))
    walls = [p for p in m.patches if p.kind == "wall"]
    assert walls
    R = np.concatenate([foamfile.boundary_values(rot, p.name, p.size) for p in walls])
    O = np.concatenate([foamfile.boundary_values(ov, p.name, p.size) for p in walls])
    Tt = np.concatenate([foamfile.boundary_values(tr, p.name, p.size) for p in walls])
    assert abs(R.mean() / 300.0 - 1) < 0.05 and (R > 150).all()
    assert np.allclose(O, (3 * Tt + 2 * R) / 5, rtol=1e-12)      # rotDoF 2
    # argon: no rotational energy at the wall, overallT = translationalT there
    c2 = _couette()
    b = c2.make_cloud(OracleCloud)
    b.evolve(4)
    b.writeFields(str(tmp_path), "4", "Ar")
    rot, ov, tr = (foamfile.read_vol_field(str(tmp_path / "4" / (n + "_Ar"))) for n in ("rotationalT", "overallT", "translationalT"))
    for p in c2.mesh.patches:
        if p.kind == "wall":
            assert not foamfile.boundary_values(rot, p.name, p.size).any()
            assert np.allclose(foamfile.boundary_values(ov, p.name, p.size), foamfile.boundary_values(tr, p.name, p.size), rtol=1e-14)
