"""Cell weighting (cellWeightedSimulation true): uniGasCloud::weighting() / cellWeighting()
(U/clouds/uniGasCloud.C:203-220, 1353-1424) and the places the cell weight factor enters the loop - NTC candidate count
(noTimeCounter.C:168-184), weighted cell sums (cellMeasurements.C:463-467), BGK state, inflow count
(uniGasGeneralBoundary.C:154-165), wall heat flux / force (uniGasPatchBoundary.C:292-299), wall fields
(uniGasVolFields.C:1276-1278).

CPU part: the oracle against what the scheme must deliver (uniform *weighted* density from non-uniform parcel counts,
the equilibrium collision rate per real molecule, exact weights on the parcels).  GPU part: libugf against the oracle
on the same seeded inputs - same streams, so the clone / delete decisions, the order of the clones inside their cells
and every parcel's state must agree.
"""
import numpy as np
import pytest

from unigasfoam_b200 import cases
from unigasfoam_b200.cloud import UgfError


def x_ramp(lo=0.5, hi=2.0):
    def f(mesh):
        x = mesh.cell_centres[:, 0]
        return lo + (hi - lo) * (x - x.min()) / (x.max() - x.min())
    return f


# ---------------------------------------------------------------------------------------------------------
# CPU: oracle
# ---------------------------------------------------------------------------------------------------------
def test_oracle_weighted_box_keeps_real_density_uniform(OracleCloud):
    case = cases.closed_box(n=8, parcels=60000, seed=3, cellWeightFactor=x_ramp())
    W = case.cellWeightFactor
    cnt0 = np.bincount(case.cell, minlength=case.mesh.n_cells)
    # uniGasMeshFill: N = n V / (F_N CWF): light cells hold more parcels
    assert cnt0[W < 0.7].mean() > 2.0 * cnt0[W > 1.8].mean()
    cl = case.make_cloud(OracleCloud)
    cloned = deleted = coll = 0
    steps = 60
    for _ in range(steps):
        cl.evolve(1)
        c = cl.counters()
        cloned += c["cloned"]; deleted += c["weightDeleted"]; coll += c["collisions"]
    p = cl.parcels()
    assert np.array_equal(p["cellWeight"], W[p["cell"]])  # every parcel carries its cell's factor
    cnt = np.bincount(p["cell"], minlength=case.mesh.n_cells)
    real = (cnt * W).reshape(8, 8, 8)
    slab = real.mean(axis=(0, 1))  # x is the fastest index
    assert np.abs(slab / slab.mean() - 1).max() < 0.04, slab
    # parcels per cell follow 1/W
    target = cnt0.sum() * (1 / W) / (1 / W).sum()
    assert abs(np.corrcoef(cnt, target)[0, 1]) > 0.9
    assert cloned > 0 and deleted > 0
    assert abs(cloned - deleted) < 0.1 * cloned  # stationary: as many made as removed
    # equilibrium collision rate: accepted collisions per step = sum_c 1/2 N_c nu dt, nu from the REAL density
    # (Bird 4.64); the F_N CWF factor in the candidate count is what makes this hold
    m = case.meta
    nu = cases.vhs_collision_rate(m["n"], m["T0"], m["species"], m["Tref"])
    expect = 0.5 * cnt.sum() * nu * case.deltaT
    assert abs(coll / steps / expect - 1) < 0.05, (coll / steps, expect)
    # temperature unchanged by cloning / deleting (info() sums are unweighted: compare per parcel)
    T = case.meta["species"]["mass"] * (p["U"] ** 2).sum() / (3 * cases.kB * len(p["U"]))
    assert abs(T / m["T0"] - 1) < 0.02


def test_oracle_uniform_factor_changes_nothing_but_the_scale(OracleCloud):
    """CWF = 2 everywhere with F_N halved is the same simulation: no clones, no deletions, same collisions."""
    a = cases.closed_box(n=6, parcels=20000, seed=9)
    b = cases.closed_box(n=6, parcels=20000, seed=9)
    b.uniGasProperties["nEquivalentParticles"] = a.uniGasProperties["nEquivalentParticles"] / 2
    b.uniGasProperties["cellWeightedSimulation"] = True
    b.cellWeightFactor = np.full(b.mesh.n_cells, 2.0)
    ca, cb = a.make_cloud(OracleCloud), b.make_cloud(OracleCloud, parcelCapacity=int(1.25 * a.n_parcels) + 1024)
    ca.evolve(5); cb.evolve(5)
    x, y = ca.counters(), cb.counters()
    assert y["cloned"] == y["weightDeleted"] == 0
    assert x["collisions"] == y["collisions"] and x["collisionCandidates"] == y["collisionCandidates"]
    pa, pb = ca.parcels(), cb.parcels()
    assert np.array_equal(pa["position"], pb["position"]) and np.array_equal(pa["U"], pb["U"])


def test_oracle_weighted_inflow_count(OracleCloud):
    """Inserted parcels per step scale with 1 / CWF of the inlet cells (uniGasGeneralBoundary.C:154-165)."""
    base = cases.cylinder(nr=12, ntheta=24, ppc=10, seed=4)
    w = cases.cylinder(nr=12, ntheta=24, ppc=10, seed=4, cellWeightFactor=4.0)
    cb, cw = base.make_cloud(OracleCloud, parcelCapacity=4 * base.n_parcels), w.make_cloud(OracleCloud)
    nb = nw = 0
    for _ in range(40):
        cb.evolve(1); cw.evolve(1)
        nb += cb.counters()["inserted"]; nw += cw.counters()["inserted"]
    assert abs(nw / nb - 0.25) < 0.03, (nb, nw)


def test_oracle_clone_and_delete_probabilities(OracleCloud):
    """cellWeighting() parcel by parcel (U/clouds/uniGasCloud.C:1366-1420): crossing into a cell whose factor is r times
    smaller yields floor(1/r - 1) clones plus one more with the remaining probability; crossing the other way the
    parcel survives with probability r.  One step across a sharp factor step, counted per crossing."""
    lo, hi = 0.4, 1.0   # hi -> lo: 1.5 clones on average; lo -> hi: survives with probability 0.4
    def step(mesh):
        mid = 0.5 * (mesh.points[:, 0].min() + mesh.points[:, 0].max())
        return np.where(mesh.cell_centres[:, 0] < mid, lo, hi)
    case = cases.closed_box(n=8, parcels=120000, seed=19, binary="noDSMCCollision", cellWeightFactor=step)
    W = case.cellWeightFactor
    cl = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels)
    w0 = W[case.cell]
    cl.move(); cl.buildCellOccupancy()
    c = cl.counters()
    # replay the move without weighting to know who crossed: same seed, factor field all equal
    ref = cases.closed_box(n=8, parcels=120000, seed=19, binary="noDSMCCollision", cellWeightFactor=1.0)
    rc = ref.make_cloud(OracleCloud, parcelCapacity=4 * ref.n_parcels)
    # same initial parcels? the fill divides by the factor, so rebuild the weighted cloud's parcels in the reference cloud
    rc.setParcels(case.position, case.U, case.cell)
    rc.move()
    w1 = W[rc.parcels()["cell"]]
    down = (w0 == hi) & (w1 == lo)
    up = (w0 == lo) & (w1 == hi)
    nd, nu = int(down.sum()), int(up.sum())
    assert nd > 1000 and nu > 1000
    exp_clones, var_clones = nd * (hi / lo - 1.0), nd * 0.25   # 1 + Bernoulli(0.5) per crossing
    assert abs(c["cloned"] - exp_clones) < 4 * np.sqrt(var_clones), (c["cloned"], exp_clones)
    p_del = 1.0 - lo / hi
    assert abs(c["weightDeleted"] - nu * p_del) < 4 * np.sqrt(nu * p_del * (1 - p_del)), (c["weightDeleted"], nu * p_del)
    assert c["nParcels"] == case.n_parcels + c["cloned"] - c["weightDeleted"]


def test_cell_weight_needs_the_switch(OracleCloud):
    case = cases.closed_box(n=4, parcels=2000, seed=1)
    cl = case.make_cloud(OracleCloud)
    with pytest.raises(UgfError, match="cellWeightedSimulation"):
        cl.setCellState(cellWeightFactor=2.0)


# ---------------------------------------------------------------------------------------------------------
# GPU: libugf vs oracle
# ---------------------------------------------------------------------------------------------------------
def both(case, GpuCloud, OracleCloud, **kw):
    return case.make_cloud(GpuCloud, **kw), case.make_cloud(OracleCloud, **kw)


def frac_close(a, b, rtol=1e-9):
    scale = np.abs(b).max() + 1e-300
    return (np.abs(a - b) <= rtol * scale).all(axis=-1).mean()


def assert_lockstep(g, r, keys=("cloned", "weightDeleted", "nParcels", "collisionCandidates"), exact=True):
    """exact: collision-free runs are bit-identical; with collisions the velocities agree to libm round-off, so later
    positions differ in the last bits for a small share of the parcels while cells, counts and decisions stay equal."""
    cg, cr = g.counters(), r.counters()
    for k in keys:
        assert cg[k] == cr[k], (k, cg[k], cr[k])
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    if exact:
        assert np.array_equal(pg["position"], pr["position"])
    else:
        assert frac_close(pg["position"], pr["position"], rtol=1e-12) > 0.995
    assert np.array_equal(pg["cellWeight"], pr["cellWeight"])
    return pg, pr, cg, cr


@pytest.mark.gpu
def test_gpu_weighting_collisionless_bit_exact(GpuCloud, OracleCloud):
    """Move + weighting + occupancy without collisions: clones, deletions and the cell-major order are deterministic
    given the streams, so the whole parcel state must be bit-identical to the oracle, step after step."""
    case = cases.closed_box(n=8, parcels=50000, seed=11, binary="noDSMCCollision", cellWeightFactor=x_ramp(0.4, 2.5))
    g, r = both(case, GpuCloud, OracleCloud)
    tot = 0
    for _ in range(6):
        g.evolve(1); r.evolve(1)
        pg, pr, cg, _ = assert_lockstep(g, r)
        assert np.array_equal(pg["U"], pr["U"])
        tot += cg["cloned"]
    assert tot > 1000
    assert (np.diff(pg["cell"]) >= 0).all()


@pytest.mark.gpu
def test_gpu_weighting_large_ratio_multiple_clones(GpuCloud, OracleCloud):
    """A factor step of 5.5 between the two halves of the box: parcels crossing it are cloned 4 or 5 times."""
    def step(mesh):
        return np.where(mesh.cell_centres[:, 0] < 0.5 * (mesh.points[:, 0].min() + mesh.points[:, 0].max()), 0.4, 2.2)
    case = cases.closed_box(n=8, parcels=30000, seed=12, binary="noDSMCCollision", cellWeightFactor=step)
    g, r = both(case, GpuCloud, OracleCloud, parcelCapacity=4 * case.n_parcels)
    for _ in range(4):
        g.evolve(1); r.evolve(1)
        pg, pr, cg, _ = assert_lockstep(g, r)
        assert np.array_equal(pg["U"], pr["U"])
    assert cg["cloned"] > 300 and cg["weightDeleted"] > 300
    # clones are exact copies: some state appears five or six times (4-5 clones + the source)
    _, mult = np.unique(pg["position"], axis=0, return_counts=True)
    assert mult.max() >= 5


@pytest.mark.gpu
def test_gpu_weighting_in_cells_larger_than_the_staging_buffer(GpuCloud, OracleCloud):
    """~310 parcels per cell: the cell kernel's slice-by-slice path and the shared / global-memory segment sorts see
    the clone entries too."""
    case = cases.closed_box(n=4, parcels=20000, seed=18, binary="noDSMCCollision", cellWeightFactor=x_ramp(0.5, 2.0))
    g, r = both(case, GpuCloud, OracleCloud)
    for _ in range(4):
        g.evolve(1); r.evolve(1)
        pg, pr, cg, _ = assert_lockstep(g, r)
        assert np.array_equal(pg["U"], pr["U"])
    assert cg["cloned"] > 100 and np.bincount(pg["cell"]).max() > 2 * 128


@pytest.mark.gpu
def test_gpu_weighting_with_ntc_collisions(GpuCloud, OracleCloud):
    case = cases.closed_box(n=8, parcels=50000, seed=13, cellWeightFactor=x_ramp())
    g, r = both(case, GpuCloud, OracleCloud)
    for _ in range(8):
        g.evolve(1); r.evolve(1)
    pg, pr, cg, cr = assert_lockstep(g, r, exact=False)
    assert abs(cg["collisions"] - cr["collisions"]) <= 2
    assert frac_close(pg["U"], pr["U"]) > 0.995
    assert abs(cg["linearKineticEnergy"] / cr["linearKineticEnergy"] - 1) < 1e-9


@pytest.mark.gpu
def test_gpu_weighting_bgk_relaxation(GpuCloud, OracleCloud):
    case = cases.closed_box(n=6, parcels=30000, seed=14, mode="bgk", binary="noDSMCCollision", bgk="unifiedStochasticParticleSBGK",
                            number_density=1e21, cellWeightFactor=x_ramp(0.6, 1.8))
    g, r = both(case, GpuCloud, OracleCloud)
    for _ in range(5):
        g.evolve(1); r.evolve(1)
    pg, pr, cg, cr = assert_lockstep(g, r, keys=("cloned", "weightDeleted", "nParcels"), exact=False)
    assert abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 2
    assert frac_close(pg["U"], pr["U"], rtol=1e-8) > 0.99


@pytest.mark.gpu
def test_gpu_weighting_hybrid_mask_and_sub_cells(GpuCloud, OracleCloud):
    """Everything the reference tutorials switch on at once: hybrid run (every other cell USP-SBGK, the rest NTC + VHS with
    2 x 2 x 1 sub-cells) under a factor ramp."""
    case = cases.closed_box(n=6, parcels=40000, seed=24, mode="hybrid", bgk="unifiedStochasticParticleSBGK", number_density=4e20,
                            cellWeightFactor=x_ramp(0.6, 1.7), theta=0.5)
    nC = case.mesh.n_cells
    case.cellCollModelId = (np.arange(nC) % 2).astype(np.int32)
    case.subCellLevels = np.tile(np.array([2, 2, 1], np.int32), (nC, 1))
    g, r = both(case, GpuCloud, OracleCloud)
    for _ in range(6):
        g.evolve(1); r.evolve(1)
    pg, pr, cg, cr = assert_lockstep(g, r, keys=("cloned", "weightDeleted", "nParcels", "collisionCandidates"), exact=False)
    assert cg["collisions"] > 0 and cg["bgkRelaxations"] > 0
    assert abs(cg["collisions"] - cr["collisions"]) <= 2 and abs(cg["bgkRelaxations"] - cr["bgkRelaxations"]) <= 2
    assert frac_close(pg["U"], pr["U"], rtol=1e-8) > 0.99


@pytest.mark.gpu
def test_gpu_weighting_mixture_with_rotation_and_tracker(GpuCloud, OracleCloud):
    """Argon / nitrogen mixture (typeId and ERot travel with the clones through the gather), Larsen-Borgnakke collisions,
    diffuse walls, per-species face tallies - the multi-species and rotational template paths under a factor ramp."""
    case = cases.mixture_box(n=6, parcels=30000, seed=25, cellWeightFactor=x_ramp(0.6, 1.8))
    m = case.mesh
    nI = m.n_internal
    mid = 0.5 * (m.points[:, 0].min() + m.points[:, 0].max())
    zone = np.nonzero((np.abs(m.face_areas[:nI, 0]) > 0) & (np.abs(m.face_centres[:nI, 0] - mid) < 1e-9 * mid))[0].astype(np.int32)
    assert len(zone) == 36
    g, r = both(case, GpuCloud, OracleCloud)
    g.setFaceTracker(zone); r.setFaceTracker(zone)
    for _ in range(6):
        g.evolve(1); r.evolve(1)
    pg, pr, cg, cr = assert_lockstep(g, r, exact=False)
    assert np.array_equal(pg["typeId"], pr["typeId"]) and set(np.unique(pg["typeId"])) == {0, 1}
    assert frac_close(pg["U"], pr["U"], rtol=1e-8) > 0.99
    assert (np.abs(pg["ERot"] - pr["ERot"]) <= 1e-8 * np.abs(pr["ERot"]).max()).mean() > 0.99
    assert (pg["ERot"][pg["typeId"] == 0] == 0).all() and (pg["ERot"][pg["typeId"] == 1] > 0).mean() > 0.99
    assert abs(cg["collisions"] - cr["collisions"]) <= 2 and cg["cloned"] > 100
    tg, tr = g.faceTracker(), r.faceTracker()
    assert tr.shape == (36, 2, 6) and np.abs(tr[:, 1, 0]).sum() > 0
    assert np.allclose(tg, tr, rtol=1e-6, atol=1e-6 * np.abs(tr).max(axis=(0, 1), keepdims=True))


@pytest.mark.gpu
def test_gpu_weighted_cylinder_inflow_walls_fields(GpuCloud, OracleCloud):
    """Graded O-grid with uniGasMeshFill's rule (CWF proportional to the cell volume: the same number of parcels in
    every cell), free-stream inflow, deleting outflow, diffuse wall: insertion counts, parcels, wall and volume fields."""
    case = cases.cylinder(nr=24, ntheta=40, ppc=12, seed=15, cellWeightFactor=("particlesPerSubCell", 12))
    W = case.cellWeightFactor
    assert W.max() / W.min() > 3
    cnt0 = np.bincount(case.cell, minlength=case.mesh.n_cells)
    assert abs(cnt0.mean() - 12) < 0.5 and cnt0.std() < 4.5
    g, r = both(case, GpuCloud, OracleCloud)
    ins = 0
    for _ in range(10):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"]
        ins += cg["inserted"]
    assert ins > 0
    pg, pr, cg, cr = assert_lockstep(g, r, exact=False)
    assert cg["wallHits"] == cr["wallHits"]
    assert frac_close(pg["U"], pr["U"]) > 0.995
    fg, fr = g.fields(), r.fields()
    for key in ("rhoN", "UMean", "translationalT", "p", "wall_rhoN", "surfaceHeatTransfer", "fD", "wall_p", "surfaceShearStress"):
        a, b = np.asarray(fg[key], float), np.asarray(fr[key], float)
        ok = np.abs(a - b) <= 1e-6 * np.abs(b) + 1e-6 * np.abs(b).max()
        assert ok.mean() > 0.99, key
    assert np.abs(fr["surfaceHeatTransfer"]).max() > 0 and fr["rhoN"].max() > 0


@pytest.mark.gpu
def test_gpu_factor_update_between_steps(GpuCloud, OracleCloud):
    """uniGasDynamicAdapter rewrites the factor field while parcels exist (uniGasDynamicAdapter.C:660-677): the next
    weighting pass clones / deletes by (factor the parcel carries) / (new factor of its cell), on both sides alike."""
    case = cases.closed_box(n=6, parcels=20000, seed=16, binary="noDSMCCollision", cellWeightFactor=x_ramp(0.8, 1.25))
    g, r = both(case, GpuCloud, OracleCloud, parcelCapacity=3 * case.n_parcels)
    g.evolve(2); r.evolve(2)
    new = x_ramp(1.6, 0.5)(case.mesh)
    g.setCellState(cellWeightFactor=new); r.setCellState(cellWeightFactor=new)
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cellWeight"], pr["cellWeight"])  # still the old factors
    g.evolve(1); r.evolve(1)
    pg, pr, cg, _ = assert_lockstep(g, r)
    assert np.array_equal(pg["cellWeight"], new[pg["cell"]])
    assert cg["cloned"] > 2000 and cg["weightDeleted"] > 2000
    g.evolve(2); r.evolve(2)
    assert_lockstep(g, r)


@pytest.mark.gpu
def test_gpu_clones_beyond_capacity_are_reported(GpuCloud):
    def step(mesh):
        return np.where(mesh.cell_centres[:, 0] < 0.5 * (mesh.points[:, 0].min() + mesh.points[:, 0].max()), 0.05, 2.0)
    case = cases.closed_box(n=6, parcels=20000, seed=17, binary="noDSMCCollision", cellWeightFactor=step)
    g = case.make_cloud(GpuCloud, parcelCapacity=case.n_parcels + 256)
    with pytest.raises(UgfError, match="capacity"):
        for _ in range(10):
            g.evolve(1)
            g.counters()
