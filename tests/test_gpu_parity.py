"""GPU (libugf, through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * deterministic work bit-exact: cell occupancy (offsets + ids) and collision-free tracking (cell index equal,
    positions identical - the north star allows 1e-12 relative, we require bit equality);
  * every collision conserves momentum and energy to 1e-12 relative;
  * stochastic work: same Philox streams on both sides, so states agree to round-off (libm ulps) for all but a
    vanishing fraction of parcels; sampled fields are additionally checked statistically in test_gpu_physics.py.
"""
import numpy as np
import pytest

from unigasfoam_b200 import cases, mesh as ugmesh

pytestmark = pytest.mark.gpu


def both(case, GpuCloud, OracleCloud, **kw):
    return case.make_cloud(GpuCloud, **kw), case.make_cloud(OracleCloud, **kw)


def assert_parcels_identical(pg, pr):
    assert np.array_equal(pg["cell"], pr["cell"])
    assert np.array_equal(pg["position"], pr["position"])
    assert np.array_equal(pg["U"], pr["U"])


def frac_close(a, b, rtol=1e-9):
    scale = np.abs(b).max() + 1e-300
    return (np.abs(a - b) <= rtol * scale).all(axis=-1).mean()


# ---------------------------------------------------------------------------------------------------------
def test_sort_bit_exact_random_order(GpuCloud, OracleCloud):
    """buildCellOccupancy on a shuffled cloud: CSR offsets and per-cell id order equal the oracle's stable sort."""
    case = cases.closed_box(n=12, parcels=60000, seed=3)
    rng = np.random.default_rng(0)
    perm = rng.permutation(case.n_parcels)
    case.position, case.U, case.cell, case.typeId = case.position[perm], case.U[perm], case.cell[perm], case.typeId[perm]
    g, r = both(case, GpuCloud, OracleCloud)
    g.buildCellOccupancy(); r.buildCellOccupancy()
    og, ig = g.cellOccupancy(); orf, irf = r.cellOccupancy()
    assert np.array_equal(og, orf)
    assert np.array_equal(ig, irf)
    # ids ascending inside every cell (stability) and a permutation overall
    assert np.array_equal(np.sort(ig), np.arange(case.n_parcels))
    seg = np.repeat(np.arange(len(og) - 1), np.diff(og))
    same = seg[1:] == seg[:-1]
    assert (np.diff(ig)[same] > 0).all()
    g.reorder(); r.reorder()
    assert_parcels_identical(g.parcels(), r.parcels())
    assert (np.diff(g.parcels()["cell"]) >= 0).all()


def test_sort_ragged_and_empty(GpuCloud, OracleCloud):
    """Empty cells, one giant cell (> shared-memory segment sort) and single-parcel cells."""
    case = cases.closed_box(n=6, parcels=3000, seed=4)
    n = 9000
    rng = np.random.default_rng(1)
    cells = np.concatenate([np.full(5000, 17), rng.integers(100, 216, n - 5000)]).astype(np.int32)  # cells 0..99 mostly empty
    rng.shuffle(cells)
    lo, hi = case.mesh.cell_bb_min[cells], case.mesh.cell_bb_max[cells]
    case.position = lo + rng.random((n, 3)) * (hi - lo)
    case.U = rng.standard_normal((n, 3)) * 300
    case.cell = cells
    case.typeId = np.zeros(n, np.int32)
    g, r = both(case, GpuCloud, OracleCloud)
    g.buildCellOccupancy(); r.buildCellOccupancy()
    og, ig = g.cellOccupancy(); orf, irf = r.cellOccupancy()
    assert np.array_equal(og, orf) and np.array_equal(ig, irf)
    assert (np.diff(og) == 0).any() and np.diff(og).max() >= 5000


@pytest.mark.parametrize("courant", [0.4, 3.7])
def test_ballistic_tracking_bit_exact_box(GpuCloud, OracleCloud, courant):
    """Collision-free move in a closed specular box; large Courant numbers force multi-face crossings and
    corner hits.  Cell index and position must be bit-identical, energy exactly conserved."""
    case = cases.closed_box(n=10, parcels=40000, seed=5, binary="noDSMCCollision")
    dx = case.meta["L"] / 10
    case.deltaT = courant * dx / cases.most_probable_speed(300.0, cases.ARGON_GUIDE["mass"])
    g, r = both(case, GpuCloud, OracleCloud)
    for _ in range(4):
        g.evolve(1); r.evolve(1)
        assert_parcels_identical(g.parcels(), r.parcels())
    cg, cr = g.counters(), r.counters()
    assert cg["stuck"] == cr["stuck"] == 0
    assert cg["wallHits"] == cr["wallHits"] > 0
    assert cg["nParcels"] == case.n_parcels
    # the parcel really is inside the cell it claims
    p = g.parcels()
    lo, hi = case.mesh.cell_bb_min[p["cell"]], case.mesh.cell_bb_max[p["cell"]]
    tol = 1e-12 * case.meta["L"]
    assert ((p["position"] >= lo - tol) & (p["position"] <= hi + tol)).all()


def test_ballistic_tracking_cyclic_2d(GpuCloud, OracleCloud):
    """2-D (empty z) channel, cyclic in x, specular walls: cyclic jumps + direction constraint, bit-exact."""
    case = cases.couette(nx=40, ny=20, ppc=25, binary="noDSMCCollision")
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        e["boundaryModel"] = "uniGasSpecularWallPatch"
    case.deltaT *= 5.0
    z0 = case.position[:, 2].copy()
    g, r = both(case, GpuCloud, OracleCloud)
    for _ in range(5):
        g.evolve(1); r.evolve(1)
    pg = g.parcels()
    assert_parcels_identical(pg, r.parcels())
    assert np.array_equal(np.sort(pg["position"][:, 2]), np.sort(z0))  # never moved in the empty direction
    assert (pg["position"][:, 0] >= -1e-12).all() and (pg["position"][:, 0] <= case.meta["Lx"] * (1 + 1e-12)).all()


def test_moments_match(GpuCloud, OracleCloud):
    case = cases.closed_box(n=8, parcels=30000, seed=6, velocity=(120.0, -40.0, 15.0))
    g, r = both(case, GpuCloud, OracleCloud)
    g.calculateFields(); r.calculateFields()
    mg, mr = g.cellMoments(), r.cellMoments()
    assert mg.shape == mr.shape
    for k in range(mg.shape[-1]):  # sums of signed terms: tolerance relative to the slot's magnitude over the mesh
        scale = np.abs(mr[..., k]).max()
        assert np.abs(mg[..., k] - mr[..., k]).max() <= 1e-11 * scale, k
    # count slot is exact
    assert np.array_equal(mg[:, 0, 0], np.bincount(case.cell, minlength=case.mesh.n_cells))


@pytest.mark.parametrize("binary", ["variableHardSphere", "variableSoftSphere"])
def test_collide_conserves_and_tracks_oracle(GpuCloud, OracleCloud, binary):
    sp = dict(cases.ARGON_GUIDE, alpha=1.4) if binary == "variableSoftSphere" else cases.ARGON_GUIDE
    case = cases.closed_box(n=8, parcels=40000, seed=8, binary=binary, species=("Ar", sp), dt_mct=1.0)
    g, r = both(case, GpuCloud, OracleCloud)
    m = sp["mass"]
    for cl in (g, r):
        cl.buildCellOccupancy(); cl.reorder()
    before = g.parcels()
    for cl in (g, r):
        cl.collide()
    after, ref = g.parcels(), r.parcels()
    assert np.array_equal(after["cell"], before["cell"]) and np.array_equal(after["position"], before["position"])
    cg, cr = g.counters(), r.counters()
    assert cg["collisionCandidates"] == cr["collisionCandidates"] > 1000
    assert cg["collisions"] == cr["collisions"] > 300
    # per-cell momentum and energy conserved by the collisions (pairs never leave their cell)
    nC = case.mesh.n_cells
    for k in range(3):
        pb = np.bincount(before["cell"], m * before["U"][:, k], nC)
        pa = np.bincount(after["cell"], m * after["U"][:, k], nC)
        scale = np.bincount(before["cell"], m * np.abs(before["U"][:, k]), nC)
        assert (np.abs(pa - pb) <= 1e-12 * scale).all()
    eb = np.bincount(before["cell"], 0.5 * m * (before["U"] ** 2).sum(1), nC)
    ea = np.bincount(after["cell"], 0.5 * m * (after["U"] ** 2).sum(1), nC)
    assert (np.abs(ea - eb) <= 1e-12 * eb).all()
    changed = (after["U"] != before["U"]).any(axis=1).sum()
    assert changed > 0 and changed <= 2 * cg["collisions"]
    # same streams => same collisions as the oracle's sequential loop, up to libm ulps
    assert frac_close(after["U"], ref["U"]) > 0.999
    sg, sr = g.cellState()["sigmaTcRMax"], r.cellState()["sigmaTcRMax"]
    np.testing.assert_allclose(sg, sr, rtol=1e-12)


@pytest.mark.parametrize("levels", [(2, 2, 2), (3, 1, 2)])
def test_subcell_partner_selection_tracks_oracle(GpuCloud, OracleCloud, levels):
    """subCellLevels > 1 (noTimeCounter.C:96-235): same candidates, same sub-cell partners, same collisions as the
    oracle's sequential loop over three full steps (move + sort + collide), cells with mixed levels."""
    case = cases.closed_box(n=6, parcels=13000, seed=12, dt_mct=0.8)
    lv = np.ones((case.mesh.n_cells, 3), np.int32)
    lv[::2] = levels  # every other cell keeps (1,1,1)
    case.subCellLevels = lv
    g, r = both(case, GpuCloud, OracleCloud)
    for cl in (g, r):
        cl.evolve(3)
    cg, cr = g.counters(), r.counters()
    assert cg["collisionCandidates"] == cr["collisionCandidates"] > 100
    assert abs(cg["collisions"] - cr["collisions"]) <= 2
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert frac_close(pg["U"], pr["U"]) > 0.995
    np.testing.assert_allclose(g.cellState()["sigmaTcRMax"], r.cellState()["sigmaTcRMax"], rtol=1e-12)


def test_ntc_subcycled_tracks_oracle(GpuCloud, OracleCloud):
    """dsmcCollisionPartnerModel noTimeCounterSubCycled (nSubCycles 3): three NTC passes with deltaT/3 each
    (noTimeCounterSubCycled.C:86-190) - same candidates and collisions as the oracle, and about as many candidates in
    total as one noTimeCounter pass."""
    counts = {}
    for partner in ("noTimeCounter", "noTimeCounterSubCycled"):
        case = cases.closed_box(n=6, parcels=20000, seed=14, dt_mct=0.9, nSubCycles=3)
        case.uniGasProperties["dsmcCollisionPartnerModel"] = partner
        g, r = both(case, GpuCloud, OracleCloud)
        for cl in (g, r):
            cl.evolve(2)
        cg, cr = g.counters(), r.counters()
        assert cg["collisionCandidates"] == cr["collisionCandidates"] > 500
        assert abs(cg["collisions"] - cr["collisions"]) <= 2
        assert frac_close(g.parcels()["U"], r.parcels()["U"]) > 0.995
        counts[partner] = cg["collisionCandidates"]
    assert abs(counts["noTimeCounter"] - counts["noTimeCounterSubCycled"]) < 0.15 * counts["noTimeCounter"]


def test_larsen_borgnakke_conserves_total_energy(GpuCloud, OracleCloud):
    case = cases.closed_box(n=6, parcels=20000, seed=9, binary="LarsenBorgnakkeVariableHardSphere", species=("N2", cases.NITROGEN),
                            dt_mct=1.0, Trot=150.0, rotationalRelaxationCollisionNumber=5.0, electronicRelaxationCollisionNumber=500.0)
    g, r = both(case, GpuCloud, OracleCloud)
    m = cases.NITROGEN["mass"]
    for cl in (g, r):
        cl.buildCellOccupancy(); cl.reorder()
    before = g.parcels()
    for cl in (g, r):
        cl.collide()
    after, ref = g.parcels(), r.parcels()
    nC = case.mesh.n_cells
    eb = np.bincount(before["cell"], 0.5 * m * (before["U"] ** 2).sum(1) + before["ERot"], nC)
    ea = np.bincount(after["cell"], 0.5 * m * (after["U"] ** 2).sum(1) + after["ERot"], nC)
    assert (np.abs(ea - eb) <= 1e-12 * eb).all()
    assert (after["ERot"] != before["ERot"]).sum() > 50  # rotational exchange happened
    assert (after["ERot"] >= 0).all()
    assert g.counters()["collisions"] == r.counters()["collisions"]
    assert frac_close(after["U"], ref["U"]) > 0.999
    assert frac_close(after["ERot"][:, None], ref["ERot"][:, None]) > 0.999


def test_full_loop_lockstep_closed_box(GpuCloud, OracleCloud):
    """evolve() x 10 on both: counters equal, state equal to round-off for (almost) every parcel."""
    case = cases.closed_box(n=8, parcels=30000, seed=10)
    g, r = both(case, GpuCloud, OracleCloud)
    e0 = g.counters()["linearKineticEnergy"]
    for _ in range(10):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["collisionCandidates"] == cr["collisionCandidates"]
        assert abs(cg["collisions"] - cr["collisions"]) <= 1
    pg, pr = g.parcels(), r.parcels()
    assert (pg["cell"] == pr["cell"]).mean() > 0.999
    assert frac_close(pg["U"], pr["U"]) > 0.995
    assert abs(g.counters()["linearKineticEnergy"] - e0) <= 1e-11 * e0  # specular box: energy constant
    fg, fr = g.fields(), r.fields()
    np.testing.assert_allclose(fg["rhoN"], fr["rhoN"], rtol=1e-12)
    np.testing.assert_allclose(fg["translationalT"], fr["translationalT"], rtol=1e-6)


def test_diffuse_walls_couette_matches_oracle(GpuCloud, OracleCloud):
    case = cases.couette(nx=24, ny=16, ppc=30, Kn=0.5)
    g, r = both(case, GpuCloud, OracleCloud)
    for cl in (g, r):
        cl.move()
    bg, br = g.boundaryMeasurements(), r.boundaryMeasurements()
    assert (bg[:, 15] == br[:, 15]).all() and br[:, 15].sum() > 100  # same wall hits on the same faces
    np.testing.assert_allclose(bg, br, rtol=1e-9, atol=1e-12 * np.abs(br).max())
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert frac_close(pg["position"], pr["position"], 1e-12) > 0.9999
    assert frac_close(pg["U"], pr["U"]) > 0.9999
    for cl in (g, r):
        cl.buildCellOccupancy(); cl.collide(); cl.accumulateFields(); cl.endStep()
    g.evolve(5); r.evolve(5)
    fg, fr = g.fields(), r.fields()
    nI = case.mesh.n_internal
    wall = np.zeros(case.mesh.n_boundary_faces, bool)
    for name in ("bottom", "top"):
        p = case.mesh.patches[case.mesh.patch_index(name)]
        wall[p.start - nI : p.start - nI + p.size] = True
    assert np.abs(fr["surfaceHeatTransfer"][wall]).max() > 0
    np.testing.assert_allclose(fg["fD"][wall], fr["fD"][wall], rtol=1e-6, atol=1e-9 * np.abs(fr["fD"]).max())
    np.testing.assert_allclose(fg["rhoN"], fr["rhoN"], rtol=1e-12)


def test_cll_walls_match_oracle(GpuCloud, OracleCloud):
    """uniGasCLLWallPatch on the Couette walls (partial accommodation, moving walls) and, with nitrogen, Lord's
    rotational-energy extension: same hits on the same faces, same reflected state to round-off, same measurements."""
    for species in (("Ar", cases.ARGON_GUIDE), ("N2", cases.NITROGEN)):
        case = cases.couette(nx=24, ny=16, ppc=30, Kn=0.5, species=species)
        for e in case.boundariesDict["uniGasPatchBoundaries"]:
            old = e["uniGasDiffuseWallPatchProperties"]
            e["boundaryModel"] = "uniGasCLLWallPatch"
            e["uniGasCLLWallPatchProperties"] = dict(old, normalAccommCoeff=0.8, tangentialAccommCoeff=0.9, rotEnergyAccommCoeff=0.7)
        g, r = both(case, GpuCloud, OracleCloud)
        for cl in (g, r):
            cl.evolve(3)
        bg, br = g.fields(), r.fields()
        pg, pr = g.parcels(), r.parcels()
        assert g.counters()["wallHits"] == r.counters()["wallHits"] > 50
        assert np.array_equal(pg["cell"], pr["cell"])
        assert frac_close(pg["U"], pr["U"]) > 0.999
        if species[0] == "N2":
            assert frac_close(pg["ERot"][:, None], pr["ERot"][:, None]) > 0.999
        np.testing.assert_allclose(bg["fD"], br["fD"], rtol=1e-6, atol=1e-8 * np.abs(br["fD"]).max())


def test_wall_field_patch_matches_oracle_and_follows_the_field(GpuCloud, OracleCloud):
    """uniGasDiffuseWallFieldPatch: wall temperature / velocity per face (boundaryT, boundaryU).  A bottom wall that
    is hot on its right half: GPU == oracle, and the parcels leaving the hot faces carry the hot temperature."""
    case = cases.couette(nx=24, ny=16, ppc=40, Kn=0.5, binary="noDSMCCollision")
    nx = 24
    T = np.where(np.arange(nx) < nx // 2, 273.0, 1200.0)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        old = e.pop("uniGasDiffuseWallPatchProperties")
        e["boundaryModel"] = "uniGasDiffuseWallFieldPatch"
        e["uniGasDiffuseWallFieldPatchProperties"] = {}
        e["boundaryT"] = T if e["patchBoundaryProperties"]["patch"] == "bottom" else old["temperature"]
        e["boundaryU"] = np.tile(np.asarray(old["velocity"], float), (nx, 1))
    g, r = both(case, GpuCloud, OracleCloud)
    for cl in (g, r):
        cl.move()
    bg, br = g.boundaryMeasurements(), r.boundaryMeasurements()
    assert (bg[:, 15] == br[:, 15]).all() and br[:, 15].sum() > 100
    np.testing.assert_allclose(bg, br, rtol=1e-9, atol=1e-12 * np.abs(br).max())
    assert frac_close(g.parcels()["U"], r.parcels()["U"]) > 0.9999
    for cl in (g, r):
        cl.buildCellOccupancy(); cl.collide(); cl.accumulateFields(); cl.endStep()
    g.evolve(30)
    f = g.fields()
    p = case.mesh.patches[case.mesh.patch_index("bottom")]
    b0 = p.start - case.mesh.n_internal
    q = f["surfaceHeatTransfer"][b0:b0 + p.size]
    assert q[nx // 2:].mean() < 5 * q[:nx // 2].mean() - 1e-12 and q[nx // 2:].mean() < 0  # the hot half heats the gas


@pytest.mark.parametrize("bgk", ["stochasticParticleBGK", "stochasticParticleESBGK", "stochasticParticleSBGK", "unifiedStochasticParticleSBGK"])
def test_bgk_family_conserves_and_tracks_oracle(GpuCloud, OracleCloud, bgk):
    case = cases.closed_box(n=6, parcels=22000, seed=11, mode="bgk", bgk=bgk, binary="noDSMCCollision", dt_mct=2.0,
                            velocity=(200.0, 50.0, -30.0), theta=0.5)
    # a non-equilibrium start so that q and sigma are non-trivial
    case.U[:, 0] *= 1.5
    g, r = both(case, GpuCloud, OracleCloud)
    m = cases.ARGON_GUIDE["mass"]
    for cl in (g, r):
        cl.buildCellOccupancy(); cl.reorder(); cl.calculateFields()
    before = g.parcels()
    for step in range(2):
        for cl in (g, r):
            cl.relax(); cl.endStep()
            cl.calculateFields()
    after, ref = g.parcels(), r.parcels()
    nC = case.mesh.n_cells
    cnt = np.bincount(before["cell"], minlength=nC)
    ok = cnt > 2
    for k in range(3):
        pb = np.bincount(before["cell"], m * before["U"][:, k], nC)
        pa = np.bincount(after["cell"], m * after["U"][:, k], nC)
        scale = np.bincount(before["cell"], m * np.abs(before["U"][:, k]), nC)
        assert (np.abs(pa - pb)[ok] <= 1e-11 * scale[ok]).all()
    eb = np.bincount(before["cell"], 0.5 * m * (before["U"] ** 2).sum(1), nC)
    ea = np.bincount(after["cell"], 0.5 * m * (after["U"] ** 2).sum(1), nC)
    assert (np.abs(ea - eb)[ok] <= 1e-11 * eb[ok]).all()
    cg, cr = g.counters(), r.counters()
    assert cg["bgkRelaxations"] > 1000
    assert abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 2
    assert frac_close(after["U"], ref["U"], 1e-8) > 0.99
    sg, sr = g.cellState(), r.cellState()
    np.testing.assert_allclose(sg["maxProb"], sr["maxProb"], rtol=1e-6)
    np.testing.assert_allclose(sg["qPrev"], sr["qPrev"], rtol=1e-6, atol=1e-9 * np.abs(sr["qPrev"]).max() + 1e-300)


def test_hybrid_mask_splits_cells(GpuCloud, OracleCloud):
    case = cases.closed_box(n=6, parcels=20000, seed=12, mode="hybrid", bgk="unifiedStochasticParticleSBGK", dt_mct=1.0)
    mask = (np.arange(case.mesh.n_cells) % 2).astype(np.int32)  # 1 = dsmc, 0 = bgk
    case.cellCollModelId = mask
    g, r = both(case, GpuCloud, OracleCloud)
    g.evolve(3); r.evolve(3)
    cg, cr = g.counters(), r.counters()
    assert cg["collisions"] > 0 and cg["bgkRelaxations"] > 0
    assert cg["collisionCandidates"] == cr["collisionCandidates"]
    assert abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 2
    assert frac_close(g.parcels()["U"], r.parcels()["U"], 1e-8) > 0.99


def test_error_paths_fail_loudly(GpuCloud):
    from unigasfoam_b200 import UgfError
    case = cases.closed_box(n=4, parcels=500, seed=13)
    case.boundariesDict["uniGasPatchBoundaries"] = case.boundariesDict["uniGasPatchBoundaries"][:-1]
    cl = case.make_cloud(GpuCloud)
    with pytest.raises(UgfError, match="wall patch without a boundary model"):
        cl.evolve(1)
    case2 = cases.closed_box(n=4, parcels=500, seed=13)
    case2.cell = case2.cell.copy(); case2.cell[0] = 10 ** 6
    with pytest.raises(UgfError, match="cell out of range"):
        case2.make_cloud(GpuCloud)
    with pytest.raises(UgfError, match="Unknown dsmcCollisionModel"):
        cases.closed_box(n=4, parcels=500, binary="hardSphere").make_cloud(GpuCloud)


def test_cylinder_inflow_outflow_matches_oracle(GpuCloud, OracleCloud):
    """Config 3 geometry (half O-grid, non-axis-aligned hexes): free-stream insertion (same streams => identical new
    parcels), deleting outflow, symmetry axis, diffuse cylinder; collision-free first, then with NTC/VHS."""
    case = cases.cylinder(nr=20, ntheta=40, ppc=25, binary="noDSMCCollision")
    g, r = both(case, GpuCloud, OracleCloud, parcelCapacity=3 * case.n_parcels)
    ins = dele = 0
    for _ in range(12):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        for k in ("nParcels", "inserted", "deleted", "wallHits", "stuck"):
            assert cg[k] == cr[k], k
        ins += cg["inserted"]; dele += cg["deleted"]
    assert ins > 300 and dele > 300 and cg["stuck"] == 0
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    # only parcels that met the diffuse wall may differ, and then only by libm round-off
    assert frac_close(pg["position"], pr["position"], 1e-12) == 1.0
    assert (pg["position"] == pr["position"]).all(axis=1).mean() > 0.98
    assert frac_close(pg["U"], pr["U"]) == 1.0
    # the same case with collisions: lockstep counters, near-identical state
    case = cases.cylinder(nr=20, ntheta=40, ppc=25)
    g, r = both(case, GpuCloud, OracleCloud, parcelCapacity=3 * case.n_parcels)
    for _ in range(8):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["collisionCandidates"] == cr["collisionCandidates"] and cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"]
    assert cg["collisions"] > 10
    assert (g.parcels()["cell"] == r.parcels()["cell"]).mean() > 0.999
    fg, fr = g.fields(), r.fields()
    np.testing.assert_allclose(fg["rhoN"], fr["rhoN"], rtol=1e-9)
    wall = slice(0, 40)  # the cylinder patch comes first among the boundary faces
    np.testing.assert_allclose(fg["surfaceHeatTransfer"][wall], fr["surfaceHeatTransfer"][wall], rtol=1e-6, atol=1e-9 * np.abs(fr["surfaceHeatTransfer"]).max())


def test_local_knudsen_decomposition_matches_oracle(GpuCloud, OracleCloud):
    """decompositionModel localKnudsen in a hybrid run (hot top wall): the time averages, smoothed fields, Knudsen fields
    and the refined DSMC / BGK mask follow the oracle through three decompositions; the parcels stay in lockstep, which
    they only do while both sides relax / collide the same cells."""
    case = cases.couette(nx=16, ny=12, ppc=60, Kn=0.2, mode="hybrid", bgk="unifiedStochasticParticleSBGK", theta=0.1)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        if e["patchBoundaryProperties"]["patch"] == "top":
            e["uniGasDiffuseWallPatchProperties"]["temperature"] = 900.0
    g, r = both(case, GpuCloud, OracleCloud)
    hd = {"decompositionModel": "localKnudsen", "timeProperties": {"decompositionInterval": 4, "resetAtDecomposition": True},
          "localKnudsenProperties": {"breakdownMax": 0.14, "theta": 0.5, "smoothingPasses": 2}}
    for cl in (g, r):
        cl.setHybridDecomposition(hd)
    split = False
    for k in range(3):
        for cl in (g, r):
            cl.evolve(4)
        dg, dr = g.hybridDecomposition(), r.hybridDecomposition()
        for name in ("KnRho", "KnT", "KnU", "KnGLL"):
            np.testing.assert_allclose(dg[name], dr[name], rtol=1e-7, err_msg=f"{name} after decomposition {k + 1}")
        assert np.array_equal(dg["cellCollModelId"], dr["cellCollModelId"])
        split = split or 0 < dr["cellCollModelId"].sum() < case.mesh.n_cells
    assert split  # at least one of the masks mixes DSMC and BGK cells
    cg, cr = g.counters(), r.counters()
    assert cg["bgkRelaxations"] == cr["bgkRelaxations"] > 0 and abs(cg["collisions"] - cr["collisions"]) <= 2
    assert frac_close(g.parcels()["U"], r.parcels()["U"], 1e-7) > 0.99


def test_gas_mixture_lockstep_and_fields(GpuCloud, OracleCloud):
    """typeIdList (Ar N2): the multi-species paths - typeId per parcel, per-species cell sums, cross-species NTC pairs
    with Larsen-Borgnakke exchange for the nitrogen partner only, diffuse walls, per-species accumulators of the
    mean-free-path fields - in lockstep with the oracle."""
    case = cases.mixture_box(rotationalRelaxationCollisionNumber=5.0, electronicRelaxationCollisionNumber=500.0)
    g, r = both(case, GpuCloud, OracleCloud)
    for cl in (g, r):
        cl.calculateFields()
    mg, mr = g.cellMoments(), r.cellMoments()
    assert mg.shape == mr.shape == (case.mesh.n_cells, 2, 32)
    for k in range(32):
        assert np.abs(mg[..., k] - mr[..., k]).max() <= 1e-11 * max(np.abs(mr[..., k]).max(), 1e-300), k
    for cl in (g, r):
        cl.evolve(6)
    pg, pr = g.parcels(), r.parcels()
    cg, cr = g.counters(), r.counters()
    assert cg["collisionCandidates"] == cr["collisionCandidates"] > 1000 and abs(cg["collisions"] - cr["collisions"]) <= 2
    assert cg["wallHits"] == cr["wallHits"] > 500
    assert np.array_equal(pg["cell"], pr["cell"]) and np.array_equal(pg["typeId"], pr["typeId"])
    assert set(np.unique(pg["typeId"])) == {0, 1}
    assert frac_close(pg["U"], pr["U"]) > 0.995
    assert frac_close(pg["ERot"][:, None], pr["ERot"][:, None]) > 0.995
    assert (pg["ERot"][pg["typeId"] == 0] == 0).all()  # argon carries no rotational energy
    fg, fr = g.fields(), r.fields()
    for name in ("rhoN", "rhoM", "translationalT", "rotationalT", "overallT", "p", "MFP", "MCR", "dtMCT", "dxMFP", "densityError", "temperatureError"):
        np.testing.assert_allclose(fg[name], fr[name], rtol=1e-6, err_msg=name)


def test_empty_cloud_filled_by_the_inflow(GpuCloud, OracleCloud):
    """Start from vacuum (no parcels at all, as the reference's expansionInVacuum tutorial does): every kernel of the
    step must cope with an empty array, and the free-stream patch fills the domain identically on both sides."""
    case = cases.cylinder(nr=10, ntheta=16, ppc=20, binary="noDSMCCollision")
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        if e["boundaryModel"] == "uniGasDiffuseWallPatch":
            e["boundaryModel"] = "uniGasSpecularWallPatch"
    n0 = case.n_parcels
    case.position, case.U = case.position[:0], case.U[:0]
    case.cell, case.typeId = case.cell[:0], case.typeId[:0]
    g, r = both(case, GpuCloud, OracleCloud, parcelCapacity=2 * n0)
    assert g.size() == 0
    g.evolve(1); r.evolve(1)
    for _ in range(15):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["inserted"] == cr["inserted"] and cg["nParcels"] == cr["nParcels"]
    assert cg["nParcels"] > 200
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])  # inserted velocities go through libm: equal to round-off, not bit for bit
    assert frac_close(pg["position"], pr["position"], rtol=1e-12) > 0.99 and frac_close(pg["U"], pr["U"]) > 0.99
    f = g.fields()
    assert np.isfinite(f["rhoN"]).all() and (f["rhoN"] == 0).any() and (f["rhoN"] > 0).any()
