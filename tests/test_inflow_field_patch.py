"""uniGasFreeStreamInflowFieldPatch (U/boundaries/derived/generalBoundaries/uniGasFreeStreamInflowFieldPatch/
uniGasFreeStreamInflowFieldPatch.C:50-228): the free-stream insertion with number density, temperatures and velocity per
face of the patch (the boundaryNumberDensity_<species>, boundaryTransT, boundaryRotT, boundaryU fields).  With uniform
fields it must be the free-stream patch bit for bit; with fields that vary along the patch every face inserts its own
Bird 4.22 flux at its own temperature."""
import copy
from math import erf

import numpy as np
import pytest

from unigasfoam_b200 import cases
from unigasfoam_b200.cloud import UgfError


def field_case(case, n=None, T=None, Trot=None, U=None):
    """The case with its free-stream patch replaced by the field variant; None keeps the free-stream value."""
    c = copy.deepcopy(case)
    e = c.boundariesDict["uniGasGeneralBoundaries"][0]
    assert e["boundaryModel"] == "uniGasFreeStreamInflowPatch"
    pr = e["uniGasFreeStreamInflowPatchProperties"]
    name = pr["typeIds"][0]
    c.boundariesDict["uniGasGeneralBoundaries"] = [{
        "generalBoundaryProperties": e["generalBoundaryProperties"], "boundaryModel": "uniGasFreeStreamInflowFieldPatch",
        "uniGasFreeStreamInflowFieldPatchProperties": {"typeIds": list(pr["typeIds"])},
        "boundaryNumberDensity": {name: pr["numberDensities"][name] if n is None else n},
        "boundaryTransT": pr["translationalTemperature"] if T is None else T,
        "boundaryRotT": pr.get("rotationalTemperature", 0.0) if Trot is None else Trot,
        "boundaryU": pr["velocity"] if U is None else U}]
    return c


def _inlet(case):
    m = case.mesh
    p = m.patches[m.patch_index(case.boundariesDict["uniGasGeneralBoundaries"][0]["generalBoundaryProperties"]["patch"])]
    S = m.face_areas[p.start:p.start + p.size]
    A = np.sqrt((S * S).sum(1))
    return p, A, -S / A[:, None], np.asarray(m.owner[p.start:p.start + p.size])


def test_oracle_uniform_fields_equal_the_free_stream_patch(OracleCloud):
    case = cases.cylinder(nr=10, ntheta=16, ppc=20, species=("N2", cases.NITROGEN), binary="LarsenBorgnakkeVariableHardSphere", seed=8)
    a = case.make_cloud(OracleCloud, parcelCapacity=3 * case.n_parcels)
    b = field_case(case).make_cloud(OracleCloud, parcelCapacity=3 * case.n_parcels)
    a.evolve(8); b.evolve(8)
    assert a.counters()["inserted"] == b.counters()["inserted"] > 10  # per step
    pa, pb = a.parcels(), b.parcels()
    for k in ("cell", "position", "U", "ERot"):
        assert np.array_equal(pa[k], pb[k]), k


def _varying(case):
    """Fields that vary along the patch: density and temperature ramps, a velocity that turns with the face normal."""
    p, A, nin, own = _inlet(case)
    s = np.linspace(0.0, 1.0, p.size)
    n = case.meta["n"] * (0.5 + 2.0 * s)
    T = case.meta["T_inf"] * (1.0 + s)
    cmp_ = np.sqrt(2 * cases.kB * T / case.meta["species"]["mass"])
    U = nin * (cmp_ * (1.5 * s - 0.5))[:, None] + np.array([0.0, 0.0, 0.0])  # speed ratio from -0.5 (outflow) to +1 along the patch
    return n, T, U


def test_oracle_each_face_inserts_its_own_flux(OracleCloud):
    base = cases.cylinder(nr=8, ntheta=24, ppc=400, binary="noDSMCCollision", seed=9)
    n, T, U = _varying(base)
    case = field_case(base, n=n, T=T, U=U)
    case.position, case.U, case.cell, case.typeId = case.position[:0], case.U[:0], case.cell[:0], case.typeId[:0]  # start from vacuum
    cl = case.make_cloud(OracleCloud, parcelCapacity=6 * base.n_parcels)
    p, A, nin, own = _inlet(case)
    m_ = base.meta["species"]["mass"]
    cmp_ = np.sqrt(2 * cases.kB * T / m_)
    sc = (U * nin).sum(1) / cmp_
    rate = A * n * cmp_ * (np.exp(-sc ** 2) + np.sqrt(np.pi) * sc * (1 + np.vectorize(erf)(sc))) / (2 * np.sqrt(np.pi) * cl.cfg.nParticle)
    counts, vn, c2 = np.zeros(p.size), np.zeros(p.size), np.zeros(p.size)
    face_of_cell = {int(c): i for i, c in enumerate(own)}
    assert len(face_of_cell) == p.size
    steps = 60
    for _ in range(steps):
        n0 = cl.size()
        cl.controlBeforeMove()          # the insertion alone: the new parcels sit on their faces, in the face's cell
        q = cl.parcels()
        new = slice(n0, None)
        f = np.array([face_of_cell[int(c)] for c in q["cell"][new]], int)
        np.add.at(counts, f, 1.0)
        np.add.at(vn, f, (q["U"][new] * nin[f]).sum(1))
        np.add.at(c2, f, ((q["U"][new] - U[f]) ** 2).sum(1) - ((q["U"][new] - U[f]) * nin[f]).sum(1) ** 2)
        cl.move(); cl.finishStep()
    expect = rate * case.deltaT * steps
    assert expect.min() > 20 and expect.max() / expect.min() > 5
    assert (np.abs(counts - expect) < 5 * np.sqrt(expect) + 1).all()
    assert abs(counts.sum() - expect.sum()) < 4 * np.sqrt(expect.sum())
    # tangential thermal energy of the inserted parcels: 2 x kT/m per parcel at the face's own temperature
    big = counts > 200
    assert big.sum() >= 4
    assert np.allclose(c2[big] / counts[big], 2 * cases.kB * T[big] / m_, rtol=0.15)
    assert (vn[big] > 0).all()  # inserted parcels move into the domain


def test_field_patch_rejects_bad_fields(OracleCloud):
    base = cases.cylinder(nr=8, ntheta=12, ppc=5, binary="noDSMCCollision")
    with pytest.raises(UgfError, match="positive temperature"):
        field_case(base, T=0.0).make_cloud(OracleCloud)
    with pytest.raises(ValueError):
        field_case(base, T=np.ones(3)).make_cloud(OracleCloud)  # not one value per face


@pytest.mark.gpu
def test_gpu_field_patch_in_lockstep_with_oracle(GpuCloud, OracleCloud):
    base = cases.cylinder(nr=10, ntheta=16, ppc=20, binary="noDSMCCollision", seed=10)
    for e in base.boundariesDict["uniGasPatchBoundaries"]:
        if e["boundaryModel"] == "uniGasDiffuseWallPatch":
            e["boundaryModel"] = "uniGasSpecularWallPatch"
    n, T, U = _varying(base)
    case = field_case(base, n=n, T=T, U=U)
    g = case.make_cloud(GpuCloud, parcelCapacity=4 * case.n_parcels)
    r = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels)
    for _ in range(10):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"] and cg["nParcels"] == cr["nParcels"]
    assert cr["inserted"] > 0
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])  # inserted velocities go through libm: equal to round-off, not bit for bit
    assert (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() > 0.99
    assert (np.abs(pg["position"] - pr["position"]) <= 1e-9 * np.abs(pr["position"]).max()).all(1).mean() > 0.99


@pytest.mark.gpu
def test_gpu_uniform_fields_equal_the_free_stream_patch(GpuCloud):
    case = cases.cylinder(nr=10, ntheta=16, ppc=20, species=("N2", cases.NITROGEN), binary="LarsenBorgnakkeVariableHardSphere", seed=8)
    a = case.make_cloud(GpuCloud, parcelCapacity=3 * case.n_parcels)
    b = field_case(case).make_cloud(GpuCloud, parcelCapacity=3 * case.n_parcels)
    a.evolve(8); b.evolve(8)
    assert a.counters()["inserted"] == b.counters()["inserted"] > 10  # per step
    pa, pb = a.parcels(), b.parcels()
    for k in ("cell", "position", "U", "ERot"):
        assert np.array_equal(pa[k], pb[k]), k


def test_field_patch_values_come_from_the_time_directory(tmp_path, OracleCloud):
    """The reference's *FieldPatch models read boundaryT / boundaryU / boundaryNumberDensity_<species> / boundaryTransT /
    boundaryRotT from the start time directory: written there as volFields, picked up by cases.field_patch_values, the
    run equals the one with the values given in the dictionary entry."""
    from unigasfoam_b200 import foamfile
    base = cases.cylinder(nr=8, ntheta=16, ppc=20, binary="noDSMCCollision", seed=12)
    n, T, U = _varying(base)
    direct = field_case(base, n=n, T=T, U=U)
    for e in direct.boundariesDict["uniGasPatchBoundaries"]:
        if e["boundaryModel"] == "uniGasDiffuseWallPatch":
            nW = base.mesh.patches[base.mesh.patch_index(e["patchBoundaryProperties"]["patch"])].size
            e["boundaryModel"] = "uniGasDiffuseWallFieldPatch"
            e["uniGasDiffuseWallFieldPatchProperties"] = {}
            e["boundaryT"] = np.linspace(300.0, 600.0, nW)
            e["boundaryU"] = np.zeros((nW, 3))
    m = base.mesh
    # the same values as volFields in <case>/0
    t0 = tmp_path / "0"
    t0.mkdir()
    g = direct.boundariesDict["uniGasGeneralBoundaries"][0]
    w = [e for e in direct.boundariesDict["uniGasPatchBoundaries"] if e["boundaryModel"] == "uniGasDiffuseWallFieldPatch"][0]
    inlet, wall = g["generalBoundaryProperties"]["patch"], w["patchBoundaryProperties"]["patch"]

    def write(name, dims, per_patch, vector=False):
        patches = {}
        for p in m.patches:
            if p.kind not in ("wall", "patch"):
                patches[p.name] = p.kind
            else:
                patches[p.name] = {"type": "calculated", "value": per_patch.get(p.name, np.zeros(3) if vector else 0.0)}
        zero = np.zeros((m.n_cells, 3)) if vector else np.zeros(m.n_cells)
        foamfile.write_vol_field(str(t0 / name), "0", dims, zero, patches, vector=vector)

    write("boundaryT", [0, 0, 0, 1, 0, 0, 0], {wall: w["boundaryT"]})
    write("boundaryU", [0, 1, -1, 0, 0, 0, 0], {wall: w["boundaryU"], inlet: g["boundaryU"]}, vector=True)
    write("boundaryNumberDensity_Ar", [0, -3, 0, 0, 0, 0, 0], {inlet: g["boundaryNumberDensity"]["Ar"]})
    write("boundaryTransT", [0, 0, 0, 1, 0, 0, 0], {inlet: g["boundaryTransT"]})
    write("boundaryRotT", [0, 0, 0, 1, 0, 0, 0], {inlet: np.broadcast_to(g["boundaryRotT"], T.shape)})
    from_files = copy.deepcopy(direct)
    for e in from_files.boundariesDict["uniGasPatchBoundaries"] + from_files.boundariesDict["uniGasGeneralBoundaries"]:
        for k in ("boundaryT", "boundaryU", "boundaryNumberDensity", "boundaryTransT", "boundaryRotT"):
            e.pop(k, None)
    with pytest.raises(FileNotFoundError):
        cases.field_patch_values(copy.deepcopy(from_files.boundariesDict), str(tmp_path), "1", m)
    cases.field_patch_values(from_files.boundariesDict, str(tmp_path), "0", m)
    g2 = from_files.boundariesDict["uniGasGeneralBoundaries"][0]
    assert np.array_equal(g2["boundaryNumberDensity"]["Ar"], n) and np.array_equal(g2["boundaryTransT"], T) and np.array_equal(g2["boundaryU"], U)
    a = direct.make_cloud(OracleCloud, parcelCapacity=4 * base.n_parcels)
    b = from_files.make_cloud(OracleCloud, parcelCapacity=4 * base.n_parcels)
    a.evolve(6); b.evolve(6)
    pa, pb = a.parcels(), b.parcels()
    assert a.counters()["wallHits"] > 0 and np.array_equal(pa["cell"], pb["cell"]) and np.array_equal(pa["U"], pb["U"])
