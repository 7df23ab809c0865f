"""Statistical parity of sampled fields, GPU vs CPU oracle with INDEPENDENT seeds (BASELINE.json north_star):
n, U, T, p and wall shear / heat flux agree within 3 sigma of the statistical error and the domain means within 1 %
over equal sample counts.  sigma is measured, not modelled: every run is cut into K consecutive sampling windows (fields
written with resetAtOutput after each), and the standard error of a time-averaged quantity is the scatter of its K window
means / sqrt(K) - that sees every fluctuation mode, including the ones uniform along the homogeneous direction that a
scatter-along-x estimate misses.  Two cases: the collision path (NTC + VHS, Kn 0.2) and the relaxation path (USP-SBGK, Kn 0.05)."""
import numpy as np
import pytest

from unigasfoam_b200 import cases

pytestmark = pytest.mark.gpu

NX, NY = 48, 32
K, WINDOW = 10, 80


def run(cloud_cls, seed, warm=160, **kw):
    case = cases.couette(nx=NX, ny=NY, ppc=50, Uw=300.0, courant=0.4, **kw)
    cl = case.make_cloud(cloud_cls, seed=seed)
    cl.evolve(warm)
    cl.fields(resetAtOutput=True)
    windows = []
    for _ in range(K):
        cl.evolve(WINDOW)
        windows.append(cl.fields(resetAtOutput=True))
    return case, windows, cl.counters()


def rows(windows, key, comp=None, shape=(NY, NX)):
    """[K, rows]: per window the profile over the slow mesh index (mean along the fast one: the homogeneous direction of the
    Couette cases, the circumference of the cylinder case)."""
    a = np.stack([w[key] if comp is None else w[key][:, comp] for w in windows])
    return a.reshape(K, shape[0], shape[1]).mean(2)


def gate(name, wg, wr, key, comp=None, rel=0.01, shape=(NY, NX)):
    """The north-star gate for one sampled field: the domain means agree within 3 sigma of their standard error and within `rel`
    (antisymmetric fields, rel None: 3 sigma only); row by row the two profiles agree within the standard error of the rows - with
    2 (K - 1) = 18 degrees of freedom behind every sigma, 0.8 % of identical samplers' rows fall outside 3 sigma, so the bound on
    single rows among the 32 is 4.5 sigma, with at most one row in ten beyond 3."""
    rg, rr = rows(wg, key, comp, shape), rows(wr, key, comp, shape)
    mg, mr = rg.mean(0), rr.mean(0)
    sigma = np.sqrt(rg.var(0, ddof=1) / K + rr.var(0, ddof=1) / K)
    z = np.abs(mg - mr) / sigma
    assert z.max() < 4.5 and (z > 3).mean() < 0.1, (name, z.max(), (z > 3).mean())
    dg, dr = rg.mean(1), rr.mean(1)  # domain mean per window
    sig_mean = np.sqrt(dg.var(ddof=1) / K + dr.var(ddof=1) / K)
    assert abs(dg.mean() - dr.mean()) < 3.0 * sig_mean, (name, dg.mean(), dr.mean(), sig_mean)
    if rel is not None:
        assert abs(dg.mean() - dr.mean()) < rel * abs(dr.mean()), name


def wall_gate(case, wg, wr, keys):
    nI = case.mesh.n_internal
    for wall in ("bottom", "top"):
        p = case.mesh.patches[case.mesh.patch_index(wall)]
        sl = slice(p.start - nI, p.start - nI + p.size)
        for key, tol in keys:
            a = np.array([w[key][sl].mean() for w in wg]); b = np.array([w[key][sl].mean() for w in wr])
            sigma = np.sqrt(a.var(ddof=1) / K + b.var(ddof=1) / K)
            assert abs(a.mean() - b.mean()) < 3.0 * sigma, (wall, key, a.mean(), b.mean(), sigma)
            if tol:
                assert abs(a.mean() - b.mean()) < tol * abs(b.mean()), (wall, key)


def all_gates(case, wg, wr, heat_tol):
    gate("rhoN", wg, wr, "rhoN")
    gate("Ux", wg, wr, "UMean", 0, rel=None)  # antisymmetric: the domain mean is zero
    gate("translationalT", wg, wr, "translationalT")
    gate("p", wg, wr, "p")
    wall_gate(case, wg, wr, (("surfaceShearStress", 0.03), ("wall_p", 0.01), ("surfaceHeatTransfer", heat_tol)))


def test_couette_fields_within_3_sigma_and_1_percent(GpuCloud, OracleCloud):
    case, wg, cg = run(GpuCloud, seed=101, Kn=0.2)
    _, wr, cr = run(OracleCloud, seed=202, Kn=0.2)
    assert cg["nParcels"] == cr["nParcels"] == case.n_parcels and cg["stuck"] == 0
    all_gates(case, wg, wr, None)  # the wall heat flux is the small difference of two large energy fluxes: its standard error is as large
                                   # as the flux itself at this sample size, so the 3 sigma bound is the whole statement
    # antisymmetric Couette profile with velocity slip at Kn = 0.2
    ux = rows(wg, "UMean", 0).mean(0)
    assert ux[0] < -150 and ux[-1] > 150 and abs(ux[0] + ux[-1]) < 15
    assert abs(ux[0]) < 300  # slip


def test_bgk_couette_fields_within_3_sigma_and_1_percent(GpuCloud, OracleCloud):
    """The same gate on the relaxation path: near-continuum Couette flow (Kn 0.05) with the unified stochastic-particle S-BGK model
    in every cell, independent seeds on the two sides."""
    kw = dict(Kn=0.05, mode="bgk", bgk="unifiedStochasticParticleSBGK", binary="noDSMCCollision", theta=0.2)
    case, wg, cg = run(GpuCloud, seed=303, **kw)
    _, wr, cr = run(OracleCloud, seed=404, **kw)
    assert cg["bgkRelaxations"] > 1000 and cr["bgkRelaxations"] > 1000 and cg["stuck"] == 0
    all_gates(case, wg, wr, None)


def test_nitrogen_larsen_borgnakke_couette_fields(GpuCloud, OracleCloud):
    """The gate on the Larsen-Borgnakke path: nitrogen Couette flow, rotational temperature included."""
    kw = dict(Kn=0.2, species=("N2", cases.NITROGEN), binary="LarsenBorgnakkeVariableHardSphere")
    case, wg, cg = run(GpuCloud, seed=505, **kw)
    _, wr, cr = run(OracleCloud, seed=606, **kw)
    assert cg["collisions"] > 100 and cg["stuck"] == 0
    all_gates(case, wg, wr, None)
    gate("rotationalT", wg, wr, "rotationalT")
    gate("overallT", wg, wr, "overallT")
    # viscous heating reaches the rotational mode through the collisions: T_rot in the core well above the wall temperature
    assert rows(wg, "rotationalT").mean(0)[NY // 2] > 1.05 * case.meta["Tw"]


def test_cylinder_inflow_outflow_fields(GpuCloud, OracleCloud):
    """The gate on an open flow: Mach-10 argon past the cylinder (free-stream inflow, deleting outflow, diffuse body) while the
    bow shock forms, sector profiles (mean over the radius per azimuthal index) of n, U, T, p and the body's wall sums.  The flow
    is not stationary over the windows: their scatter then contains the drift as well, which only widens sigma."""
    nr, nth = 24, 48

    def go(cloud_cls, seed):
        case = cases.cylinder(nr=nr, ntheta=nth, ppc=40, seed=3)
        cl = case.make_cloud(cloud_cls, seed=seed, parcelCapacity=4 * case.n_parcels)
        cl.evolve(60)
        cl.fields(resetAtOutput=True)
        w = []
        for _ in range(K):
            cl.evolve(40)
            w.append(cl.fields(resetAtOutput=True))
        return case, w, cl.counters()

    case, wg, cg = go(GpuCloud, 707)
    _, wr, cr = go(OracleCloud, 808)
    assert cg["stuck"] == 0 and cg["inserted"] > 0 and cg["deleted"] > 0
    assert case.mesh.shape[:2] == (nr, nth)
    shape = (nth, nr)  # cell index = i_r + nr * i_theta: rows = theta sectors
    for name, key, comp, rel in (("rhoN", "rhoN", None, 0.01), ("Ux", "UMean", 0, 0.01), ("Uy", "UMean", 1, None), ("translationalT", "translationalT", None, 0.01),
                                 ("p", "p", None, 0.01)):
        gate(name, wg, wr, key, comp, rel=rel, shape=shape)
    nI = case.mesh.n_internal
    p = case.mesh.patches[case.mesh.patch_index("cylinder")]
    sl = slice(p.start - nI, p.start - nI + p.size)
    for key in ("wall_p", "surfaceHeatTransfer", "surfaceShearStress"):
        a = np.array([w[key][sl].mean() for w in wg]); b = np.array([w[key][sl].mean() for w in wr])
        sigma = np.sqrt(a.var(ddof=1) / K + b.var(ddof=1) / K)
        assert abs(a.mean() - b.mean()) < 3.0 * sigma, (key, a.mean(), b.mean(), sigma)
    assert np.mean([w["surfaceHeatTransfer"][sl].mean() for w in wg]) > 0  # the Mach-10 stream heats the body


def test_equilibrium_collision_rate_on_gpu(GpuCloud):
    import math
    case = cases.closed_box(n=16, parcels=200000, seed=41)
    cl = case.make_cloud(GpuCloud)
    cl.evolve(15)
    coll = []
    for _ in range(30):
        cl.evolve(1)
        coll.append(cl.counters()["collisions"])
    cnt = np.bincount(case.cell, minlength=case.mesh.n_cells).astype(float)
    nu = cases.vhs_collision_rate(case.meta["n"], case.meta["T0"], case.meta["species"], case.meta["Tref"])
    # per cell the rate goes as N(N-1) and the uniform-gas formula as <N>^2: equal once the cell counts are Poisson,
    # which they are after the warm-up steps
    expect = 0.5 * case.n_parcels * nu * case.deltaT
    assert abs(np.mean(coll) - expect) < 4 * math.sqrt(expect / len(coll)) + 0.01 * expect
