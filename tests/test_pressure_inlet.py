"""uniGasLiouFangPressureInletPatch (U/boundaries/derived/generalBoundaries/uniGasLiouFangPressureInletPatch/
uniGasLiouFangPressureInletPatch.C:54-174): a reservoir at (p, T) feeding a channel.  Number density p / (k T); the
inflow velocity of each inlet face follows the gas in its cell."""
import numpy as np
import pytest

from unigasfoam_b200 import cases
from unigasfoam_b200.cloud import UgfError


def reservoir_case(theta=0.2, p_in=None, seed=51, **kw):
    """The cylinder O-grid with its outer upstream arc as pressure inlet at the conditions of the gas inside, the
    downstream arc deleting: gas at rest starts to drain towards the outlet and the inlet replaces it."""
    case = cases.cylinder(nr=12, ntheta=24, ppc=25, U_inf=0.0, T_wall=200.0, seed=seed, **kw)
    n, T = case.meta["n"], case.meta["T_inf"]
    inlet = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasLiouFangPressureInletPatch",
             "uniGasLiouFangPressureInletPatchProperties": {"typeIds": ["Ar"], "moleFractions": {"Ar": 1.0}, "theta": theta,
                                                            "inletPressure": (p_in if p_in is not None else n * cases.kB * T), "inletTemperature": T}}
    case.boundariesDict["uniGasGeneralBoundaries"] = [inlet]
    return case


def test_oracle_pressure_inlet_feeds_at_reservoir_density(OracleCloud):
    """Reservoir at four times the pressure of the gas inside: the first step inserts the effusion flux of the
    reservoir gas, n c_mp / (2 sqrt(pi)) per area and time with n = p / (k T); then the inlet cells fill with inward
    moving gas, the face velocities turn inward and the insertion rate grows with the speed ratio (Bird 4.22)."""
    case0 = reservoir_case(binary="noDSMCCollision")
    p0 = case0.meta["n"] * cases.kB * case0.meta["T_inf"]
    case = reservoir_case(binary="noDSMCCollision", p_in=4 * p0, theta=0.5)
    cl = case.make_cloud(OracleCloud, parcelCapacity=6 * case.n_parcels)
    m = case.mesh
    p = m.patches[m.patch_index("inlet")]
    S = m.face_areas[p.start:p.start + p.size]
    area = np.sqrt((S * S).sum(1)).sum()
    cmp_ = cases.most_probable_speed(case.meta["T_inf"], case.meta["species"]["mass"])
    expect = 4 * case.meta["n"] * cmp_ / (2 * np.sqrt(np.pi)) * area * case.deltaT / cl.cfg.nParticle
    ins = []
    for _ in range(80):
        cl.evolve(1); ins.append(cl.counters()["inserted"])
    assert expect > 50 and abs(ins[0] / expect - 1) < 4 / np.sqrt(expect)   # step 1: inlet velocity still zero
    U = cl.inletVelocity("inlet")
    nin = -S / np.sqrt((S * S).sum(1))[:, None]
    un = (U * nin).sum(1)
    assert un.mean() > 0.1 * cmp_ and (un > 0).mean() > 0.8
    assert np.mean(ins[-20:]) > 1.15 * expect
    assert cl.counters()["stuck"] == 0


def test_pressure_inlet_theta_is_checked(OracleCloud):
    case = reservoir_case(theta=1.5)
    with pytest.raises(UgfError, match="Theta"):
        case.make_cloud(OracleCloud)


@pytest.mark.gpu
def test_gpu_pressure_inlet_in_lockstep_with_oracle(GpuCloud, OracleCloud):
    """Collision-free, specular cylinder: identical parcel states, hence identical cell mean velocities, inlet
    velocities and insertion counts, step after step."""
    case = reservoir_case(binary="noDSMCCollision", theta=0.3, p_in=None)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        if e["boundaryModel"] == "uniGasDiffuseWallPatch":
            e["boundaryModel"] = "uniGasSpecularWallPatch"
    g = case.make_cloud(GpuCloud, parcelCapacity=4 * case.n_parcels)
    r = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels)
    for _ in range(12):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"] and cg["nParcels"] == cr["nParcels"]
        assert np.allclose(g.inletVelocity("inlet"), r.inletVelocity("inlet"), rtol=1e-12, atol=1e-9)
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() > 0.999
    assert np.abs(r.inletVelocity("inlet")).max() > 1.0


@pytest.mark.gpu
def test_gpu_pressure_inlet_with_collisions_and_weights(GpuCloud, OracleCloud):
    case = reservoir_case(theta=0.2, cellWeightFactor=("particlesPerSubCell", 25))
    g = case.make_cloud(GpuCloud, parcelCapacity=4 * case.n_parcels)
    r = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels)
    ig = ir = 0
    for _ in range(20):
        g.evolve(1); r.evolve(1)
        ig += g.counters()["inserted"]; ir += r.counters()["inserted"]
    assert abs(ig - ir) <= 0.02 * ir + 5
    assert np.allclose(g.inletVelocity("inlet"), r.inletVelocity("inlet"), rtol=0.2, atol=0.05 * np.abs(r.inletVelocity("inlet")).max())


# ---- uniGasWangPressureInletPatch (…/uniGasWangPressureInletPatch/uniGasWangPressureInletPatch.C:54-281) ------------------
def wang_case(p_in=None, **kw):
    case = reservoir_case(p_in=p_in, **kw)
    e = case.boundariesDict["uniGasGeneralBoundaries"][0]
    pr = e.pop("uniGasLiouFangPressureInletPatchProperties")
    pr.pop("theta")
    e["boundaryModel"] = "uniGasWangPressureInletPatch"
    e["uniGasWangPressureInletPatchProperties"] = pr
    return case


def test_oracle_wang_inlet_velocity_is_the_running_mean_then_pressure_corrected(OracleCloud):
    """Collision-free gas at rest behind an inlet at the same pressure: up to step 100 the face velocity is the running
    mass-weighted mean velocity of the parcels seen in the face's cell (checked against sums kept by the test); from step
    101 the characteristic correction (p_cell - p_in) / (rho a) along the outward normal is added (:258-266)."""
    case = wang_case(binary="noDSMCCollision")
    cl = case.make_cloud(OracleCloud, parcelCapacity=6 * case.n_parcels)
    m = case.mesh
    pt = m.patches[m.patch_index("inlet")]
    own = np.asarray(m.owner[pt.start:pt.start + pt.size])
    S = m.face_areas[pt.start:pt.start + pt.size]
    nout = S / np.sqrt((S * S).sum(1))[:, None]
    mAr = case.meta["species"]["mass"]
    n0, T0 = case.meta["n"], case.meta["T_inf"]
    p_in = n0 * cases.kB * T0
    expect0 = n0 * cases.most_probable_speed(T0, mAr) / (2 * np.sqrt(np.pi)) * np.sqrt((S * S).sum(1)).sum() * case.deltaT / cl.cfg.nParticle
    N, SU, SQ = np.zeros(pt.size), np.zeros((pt.size, 3)), np.zeros((pt.size, 3))
    for step in range(1, 106):
        cl.evolve(1)
        if step == 1:
            ins = cl.counters()["inserted"]
            assert abs(ins / expect0 - 1) < 4 / np.sqrt(expect0)        # n = p / (k T), velocity still zero (:107)
        q = cl.parcels()
        for i, c in enumerate(own):
            sel = q["cell"] == c
            N[i] += sel.sum(); SU[i] += q["U"][sel].sum(0); SQ[i] += (q["U"][sel] ** 2).sum(0)
        v = cl.inletVelocity("inlet")
        mean = SU / N[:, None]                                          # one species, one weight: momentum / mass = mean velocity
        if step <= 100:
            assert np.allclose(v, mean, rtol=1e-9, atol=1e-9 * np.abs(mean).max()), step
        else:
            V = m.cell_volumes[own]
            rho = N * cl.cfg.nParticle * mAr / (V * step)
            T = mAr / (3 * cases.kB) * ((SQ / N[:, None]).sum(1) - ((SU / N[:, None]) ** 2).sum(1))
            a = np.sqrt(5.0 / 3.0 * cases.kB / mAr * T)
            corr = (rho / mAr * cases.kB * T - p_in) / (rho * a)
            assert np.allclose(v, mean + corr[:, None] * nout, rtol=1e-8, atol=1e-8 * np.abs(mean).max()), step
            assert np.abs(corr).max() > 0
    assert cl.counters()["stuck"] == 0


def test_oracle_wang_inlet_state_survives_a_restart(tmp_path, OracleCloud):
    case = wang_case(binary="noDSMCCollision")
    a = case.make_cloud(OracleCloud, parcelCapacity=6 * case.n_parcels)
    a.evolve(4)
    a.writeTime(str(tmp_path), "4")
    a.evolve(4)
    b = OracleCloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=6 * case.n_parcels)
    b.readTime(str(tmp_path), "4")
    b.evolve(4)
    assert np.array_equal(a.inletVelocity("inlet"), b.inletVelocity("inlet")) and np.abs(a.inletVelocity("inlet")).max() > 0
    assert np.array_equal(a.parcels()["U"], b.parcels()["U"])


@pytest.mark.gpu
def test_gpu_wang_inlet_in_lockstep_with_oracle(GpuCloud, OracleCloud):
    case = wang_case(binary="noDSMCCollision", p_in=None)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        if e["boundaryModel"] == "uniGasDiffuseWallPatch":
            e["boundaryModel"] = "uniGasSpecularWallPatch"
    g = case.make_cloud(GpuCloud, parcelCapacity=4 * case.n_parcels)
    r = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels)
    for step in range(104):                                             # through the switch-on of the pressure correction
        g.evolve(1); r.evolve(1)
        if step % 8 == 0 or step > 98:
            cg, cr = g.counters(), r.counters()
            assert cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"] and cg["nParcels"] == cr["nParcels"], step
            assert np.allclose(g.inletVelocity("inlet"), r.inletVelocity("inlet"), rtol=1e-10, atol=1e-8), step
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() > 0.999
    sg, sr = g.state(), r.state()                                       # running sums and step count in the shared state layout
    assert len(sg) == len(sr) and np.allclose(sg, sr, rtol=1e-9, atol=1e-9 * np.abs(sr).max())


# ---- uniGasLiouFangPressureOutletPatch (…/uniGasLiouFangPressureOutletPatch/uniGasLiouFangPressureOutletPatch.C:50-322) ----
def outlet_case(p_out_factor=1.0, T0_factor=1.0, **kw):
    """Gas at rest in the O-grid; the downstream arc is a pressure outlet at p_out_factor x the pressure inside, the upstream
    arc keeps its pressure inlet at the inside pressure."""
    case = reservoir_case(**kw)
    n, T = case.meta["n"], case.meta["T_inf"]
    case.boundariesDict["uniGasGeneralBoundaries"].append(
        {"generalBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasLiouFangPressureOutletPatch",
         "uniGasLiouFangPressureOutletPatchProperties": {"typeIds": ["Ar"], "moleFractions": {"Ar": 1.0},
                                                         "outletPressure": p_out_factor * n * cases.kB * T, "initialOutletTemperature": T0_factor * T}})
    return case


def _outlet_state(cl, nF):
    """faceVel [nF,3], sums [nF,11], steps, faceN [nF], faceT [nF,2] of the last pressure patch from the tail of the state array."""
    s = cl.state()
    k = 3 * nF + 11 * nF + 1 + nF + 2 * nF
    t = s[len(s) - k:]
    return dict(U=t[:3 * nF].reshape(nF, 3), sums=t[3 * nF:14 * nF].reshape(nF, 11), steps=t[14 * nF], n=t[14 * nF + 1:15 * nF + 1],
                T=t[15 * nF + 1:].reshape(nF, 2))


def test_oracle_pressure_outlet_state_follows_liou_fang(OracleCloud):
    case = outlet_case(binary="noDSMCCollision", p_out_factor=0.8)
    cl = case.make_cloud(OracleCloud, parcelCapacity=6 * case.n_parcels)
    m = case.mesh
    pt = m.patches[m.patch_index("outlet")]
    own = np.asarray(m.owner[pt.start:pt.start + pt.size])
    S = m.face_areas[pt.start:pt.start + pt.size]
    nout = S / np.sqrt((S * S).sum(1))[:, None]
    mAr = case.meta["species"]["mass"]
    p_out = 0.8 * case.meta["n"] * cases.kB * case.meta["T_inf"]
    N, SU, SQ = np.zeros(pt.size), np.zeros((pt.size, 3)), np.zeros((pt.size, 3))
    ins_outlet = 0
    for step in range(1, 31):
        n0 = cl.size()
        cl.controlBeforeMove()
        q = cl.parcels()
        new_cells = q["cell"][n0:]
        ins_outlet += np.isin(new_cells, own).sum()
        if step == 1:
            assert np.isin(new_cells, own).sum() == 0          # outletNumberDensity_ starts at zero (:70)
        cl.move(); cl.finishStep()
        q = cl.parcels()
        for i, c in enumerate(own):
            sel = q["cell"] == c
            N[i] += sel.sum(); SU[i] += q["U"][sel].sum(0); SQ[i] += (q["U"][sel] ** 2).sum(0)
        st = _outlet_state(cl, pt.size)
        assert st["steps"] == step and np.array_equal(st["sums"][:, 0], N)
        rho = N * cl.cfg.nParticle * mAr / (m.cell_volumes[own] * step)
        T = mAr / (3 * cases.kB) * ((SQ / N[:, None]).sum(1) - ((SU / N[:, None]) ** 2).sum(1))
        p = rho / mAr * cases.kB * T
        a = np.sqrt(5.0 / 3.0 * cases.kB / mAr * T)
        rhoE = rho + (p_out - p) / a ** 2
        assert (rhoE > 0).all()
        assert np.allclose(st["n"], rhoE / mAr, rtol=1e-9) and np.allclose(st["T"][:, 0], p_out / (cases.kB / mAr * rhoE), rtol=1e-9)
        assert np.array_equal(st["T"][:, 0], st["T"][:, 1])
        want_U = SU / N[:, None] + ((p - p_out) / (rho * a))[:, None] * nout
        assert np.allclose(st["U"], want_U, rtol=1e-8, atol=1e-8 * np.abs(want_U).max())
    # lower pressure outside: the outlet velocity points out of the domain, and the face still feeds some gas back
    un = (st["U"] * nout).sum(1)
    assert un.mean() > 0 and (un > 0).mean() > 0.75 and ins_outlet > 0
    assert cl.counters()["stuck"] == 0


def test_oracle_pressure_outlet_restart_and_bound(tmp_path, OracleCloud):
    case = outlet_case(binary="noDSMCCollision", p_out_factor=0.9)
    a = case.make_cloud(OracleCloud, parcelCapacity=6 * case.n_parcels)
    a.evolve(4)
    a.writeTime(str(tmp_path), "4")
    a.evolve(4)
    b = OracleCloud(case.mesh, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=6 * case.n_parcels)
    b.readTime(str(tmp_path), "4")
    b.evolve(4)
    assert np.array_equal(a.state(), b.state()) and np.array_equal(a.parcels()["U"], b.parcels()["U"])
    # a bound set from an absurd initial outlet temperature (p_e / (k T_0) a millionth of the gas inside) is exceeded by the
    # first evaluated state: a loud error, not a clamp
    bad = outlet_case(binary="noDSMCCollision", T0_factor=1e6)
    c = bad.make_cloud(OracleCloud, parcelCapacity=6 * bad.n_parcels)
    with pytest.raises(UgfError, match="insertion bound"):
        for _ in range(5):
            c.evolve(1); c.counters()


@pytest.mark.gpu
def test_gpu_pressure_outlet_in_lockstep_with_oracle(GpuCloud, OracleCloud):
    case = outlet_case(binary="noDSMCCollision", p_out_factor=0.8)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        if e["boundaryModel"] == "uniGasDiffuseWallPatch":
            e["boundaryModel"] = "uniGasSpecularWallPatch"
    g = case.make_cloud(GpuCloud, parcelCapacity=4 * case.n_parcels)
    r = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels)
    nF = case.mesh.patches[case.mesh.patch_index("outlet")].size
    for step in range(24):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        assert cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"] and cg["nParcels"] == cr["nParcels"], step
    sg, sr = _outlet_state(g, nF), _outlet_state(r, nF)
    assert sg["steps"] == sr["steps"] == 24 and np.array_equal(sg["sums"][:, 0], sr["sums"][:, 0])
    for k in ("U", "sums", "n", "T"):
        assert np.allclose(sg[k], sr[k], rtol=1e-9, atol=1e-9 * np.abs(sr[k]).max()), k
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    assert (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() > 0.999
    bad = outlet_case(binary="noDSMCCollision", T0_factor=1e6)
    c = bad.make_cloud(GpuCloud, parcelCapacity=6 * bad.n_parcels)
    with pytest.raises(UgfError, match="insertion bound"):
        for _ in range(5):
            c.evolve(1); c.counters()
