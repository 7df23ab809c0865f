"""uniGasMassFlowRateInletPatch (U/boundaries/derived/generalBoundaries/uniGasMassFlowRateInletPatch/
uniGasMassFlowRateInletPatch.C:53-302): after the collisions of every step the patch sets the number density and inlet velocity
of each face so that the next step's insertion count adds up to massFlowRate dt / (m F_N) plus the parcels that left through the
patch this step (the face tracker's parcelIdFlux on the patch faces).

  * a channel closed at the far end fills at exactly the prescribed rate: N(t) grows by massFlowRate dt / (m F_N) per step;
  * the inlet velocity relaxes with theta and never points out of the domain; parcels are inserted from a gas at rest (as the
    reference does, :139-149);
  * restart through ugf_state_save / load continues bit for bit; an inlet whose cells are empty is an error (0 / 0 in the reference);
  * GPU: lockstep with the oracle."""
import math

import numpy as np
import pytest

from unigasfoam_b200 import cases, mesh as ugmesh
from unigasfoam_b200.cloud import UgfError

kB = cases.kB


def channel(mdot_factor=1.0, theta=0.5, nx=12, ny=6, ppc=40, n0=1e20, T=300.0, closed=True, binary="variableHardSphere", seed=3, fill=True,
            species=("Ar", cases.ARGON_GUIDE), initialVelocity=(0.0, 0.0, 0.0)):
    name, sp = species
    kinds = {"xMin": ("inlet", "patch"), "xMax": ("end", "wall" if closed else "patch"), "yMin": ("bottom", "symmetryPlane"),
             "yMax": ("top", "symmetryPlane"), "zMin": ("back", "empty"), "zMax": ("front", "empty")}
    lam = cases.vhs_mean_free_path(n0, T, sp, 273.0)
    dx = lam / 2.0
    L, H, W = nx * dx, ny * dx, dx
    m = ugmesh.box_mesh(nx, ny, 1, L, H, W, kinds, solution_d=(1, 1, 0))
    m.meta_axis_aligned = True
    FN = n0 * L * H * W / (ppc * nx * ny)
    props = cases._props(name, sp, FN, binary=binary, Tref=273.0)
    cmp_ = cases.most_probable_speed(T, sp["mass"])
    dt = 0.3 * dx / cmp_
    # a flow rate comparable with the one-sided thermal flux through the inlet area
    mdot = mdot_factor * sp["mass"] * n0 * cmp_ / (2 * math.sqrt(math.pi)) * H * W
    inlet = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasMassFlowRateInletPatch",
             "uniGasMassFlowRateInletPatchProperties": {"typeIds": [name], "moleFractions": {name: 1.0}, "inletTemperature": T,
                                                        "massFlowRate": mdot, "theta": theta, "initialVelocity": list(initialVelocity)}}
    patch = [{"patchBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasDeletionPatch"}]
    if closed:
        patch.append({"patchBoundaryProperties": {"patch": "end"}, "boundaryModel": "uniGasSpecularWallPatch"})
    else:
        patch.append({"patchBoundaryProperties": {"patch": "end"}, "boundaryModel": "uniGasDeletionPatch"})
    bd = {"uniGasPatchBoundaries": patch, "uniGasGeneralBoundaries": [inlet]}
    rng = np.random.default_rng(seed)
    if fill:
        pos, vel, cel, tid, erot = cases.mesh_fill(m, {name: sp}, [name], {name: n0}, T, (0.0, 0.0, 0.0), FN, rng)
    else:
        pos, vel, cel = np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, np.int32)
    sig0 = math.pi * sp["diameter"] ** 2 * cmp_
    meta = dict(FN=FN, mdot=mdot, dt=dt, mass=sp["mass"], n0=n0, T=T, cmp=cmp_, H=H, W=W)
    return m, props, bd, dt, pos, vel, cel, sig0, meta


def make(Cloud, seed_cloud=7, **kw):
    m, props, bd, dt, pos, vel, cel, sig0, meta = channel(**kw)
    cl = Cloud(m, props, bd, dt, parcelCapacity=6 * max(len(cel), 2000), seed=seed_cloud)
    cl.setParcels(pos, vel, cel)
    cl.setCellState(sigmaTcRMax=sig0)
    return cl, m, meta


def test_closed_channel_fills_at_the_prescribed_rate(OracleCloud):
    cl, m, me = make(OracleCloud, mdot_factor=1.0)
    per_step = me["mdot"] * me["dt"] / (me["mass"] * me["FN"])
    cl.evolve(3)  # the first step inserts nothing (number densities start at zero, :84-89)
    n1 = cl.size()
    steps = 120
    ins = dele = 0
    for _ in range(steps):
        cl.evolve(1)
        c = cl.counters()
        ins += c["inserted"]; dele += c["deleted"]
    gain = cl.size() - n1
    assert gain == ins - dele
    expect = steps * per_step
    assert expect > 2000
    # the count of every step is the target plus what left the step before: the gain telescopes to the target up to the rounding
    # draws (one Bernoulli per slot and step) and the difference of the first and last steps' back-flux
    assert abs(gain - expect) < 4 * math.sqrt(steps * 6 * 0.25) + 0.02 * expect, (gain, expect)
    assert dele > 0.3 * ins  # a large share of the inserted parcels leaves again through the inlet: the patch compensates for it
    v = cl.inletVelocity("inlet")
    assert (v[:, 0] >= 0).all()  # never out of the domain (:228-231)


def test_first_step_inserts_nothing_and_empty_inlet_cells_are_an_error(OracleCloud):
    cl, m, me = make(OracleCloud)
    cl.evolve(1)
    assert cl.counters()["inserted"] == 0
    cl.evolve(1)
    assert cl.counters()["inserted"] > 0
    cl2, _, _ = make(OracleCloud, fill=False)
    with pytest.raises(UgfError, match="no parcels in the cells of the inlet patch"):
        cl2.evolve(1)


def test_inserted_parcels_come_from_a_gas_at_rest_and_velocity_relaxes(OracleCloud):
    cl, m, me = make(OracleCloud, mdot_factor=3.0, theta=0.25, binary="noDSMCCollision", ppc=200, closed=False)
    cl.evolve(2)
    vn, cnt = 0.0, 0
    hist = []
    for _ in range(60):
        n0 = cl.size()
        cl.controlBeforeMove()
        q = cl.parcels()
        new = q["U"][n0:]
        vn += new[:, 0].sum(); cnt += len(new)
        cl.move(); cl.finishStep()
        hist.append(cl.inletVelocity("inlet")[:, 0].mean())
    # half-range Maxwellian of a gas at rest: <u_n> = sqrt(pi) / 2 c_mp for the flux-weighted inflowing molecules
    assert abs(vn / cnt / (0.5 * math.sqrt(math.pi) * me["cmp"]) - 1) < 0.03
    assert hist[-1] > 0 and hist[0] < 0.6 * np.mean(hist[-10:])  # relaxing upwards with theta = 0.25 from zero


def test_restart_is_bit_exact(OracleCloud):
    a, m, me = make(OracleCloud)
    a.evolve(10)
    st, p = a.state(), a.parcels()
    b, _, _ = make(OracleCloud)
    b.setParcels(p["position"], p["U"], p["cell"])
    b.loadState(st)
    a.evolve(6); b.evolve(6)
    pa, pb = a.parcels(), b.parcels()
    assert a.counters()["inserted"] == b.counters()["inserted"] > 0
    assert np.array_equal(pa["cell"], pb["cell"]) and np.array_equal(pa["U"], pb["U"]) and np.array_equal(pa["position"], pb["position"])


def test_mixture_needs_all_species_in_order(OracleCloud):
    case = cases.mixture_box(n=3, parcels=500)
    m, props = case.mesh, case.uniGasProperties
    entry = {"generalBoundaryProperties": {"patch": m.patches[0].name}, "boundaryModel": "uniGasMassFlowRateInletPatch",
             "uniGasMassFlowRateInletPatchProperties": {"typeIds": ["N2"], "moleFractions": {"N2": 1.0}, "inletTemperature": 300.0, "massFlowRate": 1e-9}}
    with pytest.raises(UgfError, match="typeIdList order"):
        type(case.make_cloud(OracleCloud))(m, props, {"uniGasGeneralBoundaries": [entry]}, case.deltaT, parcelCapacity=1000)


@pytest.mark.gpu
def test_gpu_mass_flow_inlet_in_lockstep_with_oracle(GpuCloud, OracleCloud):
    g, m, me = make(GpuCloud, mdot_factor=1.5, theta=0.5)
    r, _, _ = make(OracleCloud, mdot_factor=1.5, theta=0.5)
    tot = 0
    for _ in range(25):
        g.evolve(1); r.evolve(1)
        cg, cr = g.counters(), r.counters()
        for k in ("nParcels", "inserted", "deleted", "collisions", "collisionCandidates"):
            assert cg[k] == cr[k], (k, cg[k], cr[k])
        tot += cr["inserted"]
    assert tot > 500
    assert np.allclose(g.inletVelocity("inlet"), r.inletVelocity("inlet"), rtol=1e-10, atol=1e-9)
    pg, pr = g.parcels(), r.parcels()
    assert np.array_equal(pg["cell"], pr["cell"])
    # restart on the GPU continues like the uninterrupted run
    st, p = g.state(), g.parcels()
    h, _, _ = make(GpuCloud, mdot_factor=1.5, theta=0.5)
    h.setParcels(p["position"], p["U"], p["cell"])
    h.loadState(st)
    g.evolve(5); h.evolve(5)
    assert g.counters()["inserted"] == h.counters()["inserted"] and g.size() == h.size()
    assert np.array_equal(g.parcels()["U"], h.parcels()["U"])


@pytest.mark.gpu
def test_gpu_closed_channel_fills_at_the_prescribed_rate(GpuCloud):
    cl, m, me = make(GpuCloud, mdot_factor=1.0, nx=40, ny=20, ppc=60)
    per_step = me["mdot"] * me["dt"] / (me["mass"] * me["FN"])
    cl.evolve(3)
    n1 = cl.size()
    steps = 200
    cl.evolve(steps)
    gain = cl.size() - n1
    expect = steps * per_step
    assert abs(gain - expect) < 4 * math.sqrt(steps * 20 * 0.25) + 0.02 * expect, (gain, expect)
