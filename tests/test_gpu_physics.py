"""Statistical parity of sampled fields, GPU vs CPU oracle with INDEPENDENT seeds (BASELINE.json north_star):
n, U, T, p and wall shear / heat flux agree within 3 sigma of the statistical error and the domain means within 1 %
over equal sample counts.  sigma is measured from the scatter along the homogeneous (x) direction of each run.  Two cases: the
collision path (NTC + VHS, Kn 0.2) and the relaxation path (USP-SBGK, Kn 0.05)."""
import numpy as np
import pytest

from unigasfoam_b200 import cases

pytestmark = pytest.mark.gpu

NX, NY = 48, 32


def run(cloud_cls, seed, warm=150, steps=400, **kw):
    case = cases.couette(nx=NX, ny=NY, ppc=50, Uw=300.0, courant=0.4, **kw)
    cl = case.make_cloud(cloud_cls, seed=seed)
    cl.evolve(warm)
    cl.fields(resetAtOutput=True)
    cl.evolve(steps)
    return case, cl.fields(), cl.counters()


def profile(a):
    a = a.reshape(NY, NX)
    return a.mean(1), a.std(1, ddof=1) / np.sqrt(NX)


def gate(name, a, b, rel=0.01):
    """The north-star gate for one sampled field (profiles over y): the domain means agree within 3 sigma of their statistical
    error and within `rel`; row by row the two profiles agree within the scatter.  sigma comes from the scatter along the
    homogeneous direction, which misses fluctuation modes that are uniform in x (they shift a whole row of the time average and
    correlate neighbouring rows), so the per-row bound allows for that: no row beyond 4.5 sigma, at most one in ten beyond 3."""
    mg, sg = profile(a)
    mr, sr = profile(b)
    sigma = np.sqrt(sg ** 2 + sr ** 2)
    z = np.abs(mg - mr) / sigma
    assert z.max() < 4.5 and (z > 3).mean() < 0.1, (name, z.max())
    sig_mean = np.sqrt((sg ** 2).sum() + (sr ** 2).sum()) / NY
    if rel is not None:
        assert abs(mg.mean() - mr.mean()) < max(3.0 * sig_mean, 1e-3 * abs(mr.mean())), (name, mg.mean(), mr.mean(), sig_mean)
        assert abs(mg.mean() - mr.mean()) < rel * abs(mr.mean()), name
    else:
        assert abs(mg.mean() - mr.mean()) < 4.5 * sig_mean, (name, mg.mean(), mr.mean(), sig_mean)


def wall_gate(case, fg, fr, keys):
    nI = case.mesh.n_internal
    for wall in ("bottom", "top"):
        p = case.mesh.patches[case.mesh.patch_index(wall)]
        sl = slice(p.start - nI, p.start - nI + p.size)
        for key, tol in keys:
            a, b = fg[key][sl], fr[key][sl]
            sigma = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
            assert abs(a.mean() - b.mean()) < 3.0 * sigma, (wall, key, a.mean(), b.mean(), sigma)
            if tol:
                assert abs(a.mean() - b.mean()) < tol * abs(b.mean()), (wall, key)


def test_couette_fields_within_3_sigma_and_1_percent(GpuCloud, OracleCloud):
    case, fg, cg = run(GpuCloud, seed=101, Kn=0.2)
    _, fr, cr = run(OracleCloud, seed=202, Kn=0.2)
    assert cg["nParcels"] == cr["nParcels"] == case.n_parcels and cg["stuck"] == 0
    gate("rhoN", fg["rhoN"], fr["rhoN"])
    gate("Ux", fg["UMean"][:, 0], fr["UMean"][:, 0], rel=None)  # antisymmetric: the domain mean is zero
    gate("translationalT", fg["translationalT"], fr["translationalT"])
    gate("p", fg["p"], fr["p"])
    # antisymmetric Couette profile with velocity slip at Kn = 0.2
    ux, _ = profile(fg["UMean"][:, 0])
    assert ux[0] < -150 and ux[-1] > 150 and abs(ux[0] + ux[-1]) < 15
    assert abs(ux[0]) < 300  # slip
    wall_gate(case, fg, fr, (("surfaceShearStress", 0.03), ("wall_p", 0.01), ("surfaceHeatTransfer", 0.05)))
    nI = case.mesh.n_internal
    p = case.mesh.patches[case.mesh.patch_index("bottom")]
    assert fg["surfaceHeatTransfer"][p.start - nI:p.start - nI + p.size].mean() < 0  # viscous heating leaves through the walls


def test_bgk_couette_fields_within_3_sigma_and_1_percent(GpuCloud, OracleCloud):
    """The same gate on the relaxation path: near-continuum Couette flow (Kn 0.05) with the unified stochastic-particle S-BGK model
    in every cell, independent seeds on the two sides."""
    kw = dict(Kn=0.05, mode="bgk", bgk="unifiedStochasticParticleSBGK", binary="noDSMCCollision", theta=0.2)
    case, fg, cg = run(GpuCloud, seed=303, **kw)
    _, fr, cr = run(OracleCloud, seed=404, **kw)
    assert cg["bgkRelaxations"] > 1000 and cr["bgkRelaxations"] > 1000 and cg["stuck"] == 0
    gate("rhoN", fg["rhoN"], fr["rhoN"])
    gate("Ux", fg["UMean"][:, 0], fr["UMean"][:, 0], rel=None)
    gate("translationalT", fg["translationalT"], fr["translationalT"])
    gate("p", fg["p"], fr["p"])
    wall_gate(case, fg, fr, (("surfaceShearStress", 0.03), ("wall_p", 0.01), ("surfaceHeatTransfer", None)))


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
