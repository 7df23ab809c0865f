"""collisionProperties.macroInterpolation true (set by all four reference tutorials): the BGK target fields are interpolated to the
relaxing parcel's position by OpenFOAM's interpolationCellPoint (U/bgkCollisions/derived/unifiedStochasticParticleSBGK/
unifiedStochasticParticleSBGK.C:893-947, same block in the other three models).  OpenFOAM is not in the reference tree: the geometry
(unigasfoam_b200.mesh.cell_point_data) and the interpolation are restated from its definition and pinned here to what that
definition implies - partition of unity, exactness for constants, the tets tiling every cell, coupled points seeing both sides,
symmetry points carrying no normal components - and the relaxation is pinned to conservation and to the cell-value model in the
limit where both must agree (a uniform gas)."""
import os

import numpy as np
import pytest

from unigasfoam_b200 import cases, foamdict, mesh as ugmesh

GOLD = os.path.join(os.path.dirname(__file__), "golden", "openfoam")


def _values_at_points(d, cell_values):
    nP = len(d["pointCellOffsets"]) - 1
    rows = np.repeat(np.arange(nP), np.diff(d["pointCellOffsets"]))
    return np.bincount(rows, weights=d["pointWeights"] * cell_values[d["pointCells"]], minlength=nP)


def test_cell_point_geometry_couette():
    case = cases.couette(nx=10, ny=6, ppc=2)
    m = case.mesh
    d = ugmesh.cell_point_data(m)
    nP = len(m.points)
    assert (np.diff(d["pointCellOffsets"]) > 0).all()                                   # every point is fed by some cell
    np.testing.assert_allclose(_values_at_points(d, np.ones(m.n_cells)), 1.0, rtol=1e-14)  # partition of unity
    assert (np.diff(d["tetOffsets"]) == 12).all()                                       # hex: 6 quads x 2 triangles
    cc = m.cell_centres[np.repeat(np.arange(m.n_cells), 12)]
    a, b, c = (m.points[d["tetPoints"][:, k]] for k in range(3))
    vol = np.abs(np.einsum("ij,ij->i", np.cross(a - cc, b - cc), c - cc)) / 6.0
    np.testing.assert_allclose(np.bincount(np.repeat(np.arange(m.n_cells), 12), weights=vol), m.cell_volumes, rtol=1e-12)
    # wall points are fed by the wall faces' owner cells only: the first cell row at the bottom wall
    y = m.points[:, 1]
    H = case.meta["H"]
    bottom = np.nonzero(y < 1e-12 * H)[0]
    rows = np.repeat(np.arange(nP), np.diff(d["pointCellOffsets"]))
    fed_by = d["pointCells"][np.isin(rows, bottom)]
    assert (fed_by // 10 == 0).all()
    # x is cyclic: a point on the left cyclic patch sees cells of the last column too
    left = np.nonzero((m.points[:, 0] < 1e-12 * case.meta["Lx"]) & (y > 0.3 * H) & (y < 0.7 * H))[0]
    for p in left[:4]:
        cols = d["pointCells"][d["pointCellOffsets"][p]:d["pointCellOffsets"][p + 1]] % 10
        assert set(cols) == {0, 9}
    assert not d["pointNormals"].any()                                                   # no symmetry patch here


def test_symmetry_points_carry_a_normal():
    m = ugmesh.half_annulus_mesh(6, 10, 0.15, 0.6, 0.03, 3.0)
    d = ugmesh.cell_point_data(m)
    on_axis = np.abs(m.points[:, 1]) < 1e-14
    n = d["pointNormals"]
    assert (np.abs(np.abs(n[on_axis, 1]) - 1.0) < 1e-12).all() and not n[~on_axis].any()
    np.testing.assert_allclose(_values_at_points(d, np.ones(m.n_cells)), 1.0, rtol=1e-14)


@pytest.mark.parametrize("bgk", ["stochasticParticleBGK", "stochasticParticleESBGK", "stochasticParticleSBGK", "unifiedStochasticParticleSBGK"])
def test_relaxation_with_interpolation_conserves(OracleCloud, bgk):
    case = cases.closed_box(n=5, parcels=16000, seed=51, mode="bgk", bgk=bgk, binary="noDSMCCollision", dt_mct=2.0, velocity=(150.0, 40.0, -20.0),
                            theta=0.5, macroInterpolation=True)
    case.U[:, 0] *= 1.4  # off equilibrium: heat flux and shear stress are not zero
    cl = case.make_cloud(OracleCloud)
    m = cases.ARGON_GUIDE["mass"]
    cl.buildCellOccupancy(); cl.reorder(); cl.calculateFields()
    before = cl.parcels()
    cl.relax()
    after = cl.parcels()
    nC = case.mesh.n_cells
    ok = np.bincount(before["cell"], minlength=nC) > 2
    for k in range(3):
        pb = np.bincount(before["cell"], m * before["U"][:, k], nC)
        pa = np.bincount(after["cell"], m * after["U"][:, k], nC)
        scale = np.bincount(before["cell"], m * np.abs(before["U"][:, k]), nC)
        assert (np.abs(pa - pb)[ok] <= 1e-11 * scale[ok]).all()
    eb = np.bincount(before["cell"], 0.5 * m * (before["U"] ** 2).sum(1), nC)
    ea = np.bincount(after["cell"], 0.5 * m * (after["U"] ** 2).sum(1), nC)
    assert (np.abs(ea - eb)[ok] <= 1e-11 * eb[ok]).all()
    assert cl.counters()["bgkRelaxations"] > 1000


def test_uniform_gas_relaxes_alike_with_and_without_interpolation(OracleCloud):
    """A uniform equilibrium gas: the interpolated target state differs from the cell's only by sampling noise, so the relaxation
    counts agree (exactly in the first step: they depend on the cell values alone) and the temperature stays put."""
    out = {}
    for interp in (False, True):
        case = cases.closed_box(n=5, parcels=40000, seed=52, mode="bgk", bgk="unifiedStochasticParticleSBGK", binary="noDSMCCollision",
                                dt_mct=1.0, theta=0.3, macroInterpolation=interp)
        cl = case.make_cloud(OracleCloud)
        cl.evolve(1)
        first = cl.counters()["bgkRelaxations"]
        cl.evolve(5)
        c = cl.counters()
        out[interp] = (c["bgkRelaxations"], cl.fields()["translationalT"].mean(), first)
    assert out[True][2] == out[False][2] > 1000                    # first step: same cell values, same counts
    assert abs(out[True][0] / out[False][0] - 1.0) < 0.01           # later the two clouds differ by sampling noise only
    assert abs(out[True][1] / 300.0 - 1.0) < 0.01 and abs(out[False][1] / 300.0 - 1.0) < 0.01


def test_tutorial_collision_properties_run_as_written(OracleCloud):
    """hypersonicCylinder/constant/uniGasProperties, copied unmodified under tests/golden/openfoam: collisionProperties
    { Tref 1000; macroInterpolation true; theta 0.1; } is taken as written."""
    m = ugmesh.half_annulus_mesh(12, 20, 0.5 * 0.3048, 2.0 * 0.3048, 0.1 * 0.3048, 5.0)
    m.meta_axis_aligned = False
    case, ld = cases.from_case_dir(os.path.join(GOLD, "hypersonicCylinder"), m, seed=7,
                                   overrides={"adaptiveProperties": {"maxSubCellSizeMFPRatio": 4.0, "adaptationInterval": 20}})
    assert case.uniGasProperties["collisionProperties"]["macroInterpolation"] is True
    cl = case.make_cloud(OracleCloud, parcelCapacity=8 * case.n_parcels)
    cl.setCellState(cellCollModelId=np.zeros(m.n_cells, np.int32))  # all cells relax (the hybrid mask starts as all-BGK)
    cl.evolve(5)
    c = cl.counters()
    assert c["bgkRelaxations"] > 100 and c["stuck"] == 0
    assert np.isfinite(cl.parcels()["U"]).all()


def frac_close(a, b, rtol=1e-8):
    scale = np.abs(b).max() + 1e-300
    return (np.abs(a - b) <= rtol * scale).all(axis=-1).mean()


@pytest.mark.gpu
@pytest.mark.parametrize("bgk", ["stochasticParticleBGK", "stochasticParticleESBGK", "stochasticParticleSBGK", "unifiedStochasticParticleSBGK"])
def test_gpu_relaxation_with_interpolation_tracks_the_oracle(GpuCloud, OracleCloud, bgk):
    """bgk_fields_kernel + bgk_points_kernel + the INTERP instantiation of bgk_kernel against the oracle: same streams, same relaxing
    parcels, same interpolated target states up to round-off; conservation per cell to 1e-11."""
    case = cases.closed_box(n=6, parcels=22000, seed=53, mode="bgk", bgk=bgk, binary="noDSMCCollision", dt_mct=2.0, velocity=(200.0, 50.0, -30.0),
                            theta=0.5, macroInterpolation=True)
    case.U[:, 0] *= 1.5
    g, r = case.make_cloud(GpuCloud), case.make_cloud(OracleCloud)
    m = cases.ARGON_GUIDE["mass"]
    for cl in (g, r):
        cl.buildCellOccupancy(); cl.reorder(); cl.calculateFields()
    before = g.parcels()
    for _ in range(2):
        for cl in (g, r):
            cl.relax(); cl.endStep(); cl.calculateFields()
    after, ref = g.parcels(), r.parcels()
    nC = case.mesh.n_cells
    ok = np.bincount(before["cell"], minlength=nC) > 2
    eb = np.bincount(before["cell"], 0.5 * m * (before["U"] ** 2).sum(1), nC)
    ea = np.bincount(after["cell"], 0.5 * m * (after["U"] ** 2).sum(1), nC)
    assert (np.abs(ea - eb)[ok] <= 1e-11 * eb[ok]).all()
    cg, cr = g.counters(), r.counters()
    assert cg["bgkRelaxations"] > 1000 and abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 2
    assert frac_close(after["U"], ref["U"]) > 0.99
    sg, sr = g.cellState(), r.cellState()
    np.testing.assert_allclose(sg["maxProb"], sr["maxProb"], rtol=1e-6)


@pytest.mark.gpu
def test_gpu_couette_with_walls_and_cyclics_interpolated(GpuCloud, OracleCloud):
    """2-D (empty direction, cyclic pair, wall boundary points): full steps in lockstep with the oracle."""
    case = cases.couette(nx=16, ny=12, ppc=50, Kn=0.05, mode="bgk", bgk="unifiedStochasticParticleSBGK", binary="noDSMCCollision", theta=0.5,
                         macroInterpolation=True)
    g, r = case.make_cloud(GpuCloud), case.make_cloud(OracleCloud)
    for _ in range(6):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 2 and cg["wallHits"] == cr["wallHits"]
    assert cg["bgkRelaxations"] > 500
    assert frac_close(g.parcels()["U"], r.parcels()["U"], 1e-7) > 0.98


@pytest.mark.gpu
def test_gpu_tutorial_collision_properties_as_written(GpuCloud, OracleCloud):
    m = ugmesh.half_annulus_mesh(12, 20, 0.5 * 0.3048, 2.0 * 0.3048, 0.1 * 0.3048, 5.0)
    m.meta_axis_aligned = False
    case, ld = cases.from_case_dir(os.path.join(GOLD, "hypersonicCylinder"), m, seed=7,
                                   overrides={"adaptiveProperties": {"maxSubCellSizeMFPRatio": 4.0, "adaptationInterval": 20}})
    clouds = []
    for Cloud in (GpuCloud, OracleCloud):
        cl = case.make_cloud(Cloud, parcelCapacity=8 * case.n_parcels)
        cl.setCellState(cellCollModelId=np.zeros(m.n_cells, np.int32))
        cl.evolve(5)
        clouds.append(cl)
    cg, cr = clouds[0].counters(), clouds[1].counters()
    for k in ("nParcels", "inserted", "deleted", "wallHits", "cloned", "weightDeleted"):
        assert cg[k] == cr[k], k
    assert abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 3 and cg["bgkRelaxations"] > 100
