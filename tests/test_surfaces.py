"""uniGasFaceTracker (U/faceTracker/uniGasFaceTracker.C:90-152) and its consumers uniGasMassFluxSurface /
uniGasForceSurface: a drifting gas carries rho u through any cross-section; the device tallies equal the oracle's."""
import numpy as np
import pytest

from unigasfoam_b200 import cases
from unigasfoam_b200.surfaces import UniGasForceSurface, UniGasMassFluxSurface

U_DRIFT = 400.0


def drifting_channel(**kw):
    case = cases.couette(nx=20, ny=10, ppc=40, binary="noDSMCCollision", Uw=0.0, **kw)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        e["boundaryModel"] = "uniGasSpecularWallPatch"
    case.U = case.U + np.array([U_DRIFT, 0.0, 0.0])
    return case


def plane_faces(mesh, x):
    """Internal faces of the cross-section x = const (a face zone)."""
    nI = mesh.n_internal
    S, C = mesh.face_areas[:nI], mesh.face_centres[:nI]
    dx = (mesh.cell_bb_max - mesh.cell_bb_min)[0, 0]
    return np.nonzero((np.abs(S[:, 0]) > 0) & (np.abs(C[:, 0] - x) < 1e-6 * dx))[0].astype(np.int32)


def test_oracle_mass_flux_through_a_cross_section(OracleCloud):
    case = drifting_channel()
    m = case.mesh
    dx = (m.cell_bb_max - m.cell_bb_min)[0, 0]
    zone = plane_faces(m, 7 * dx)
    assert len(zone) == 10
    cl = case.make_cloud(OracleCloud)
    surf = UniGasMassFluxSurface(cl, {"field": "Ar", "faceZone": zone, "fluxDirection": [2.0, 0.0, 0.0], "typeIds": ["Ar"]})
    assert surf.zoneSurfaceArea == pytest.approx(case.meta["H"] * dx, rel=1e-12)
    surf.begin()
    cl.evolve(40)
    f = surf.calculateField()
    rho = case.meta["n"] * case.meta["species"]["mass"]
    assert f["massFlux"] == pytest.approx(rho * U_DRIFT, rel=0.03)
    assert f["molFlux"] == pytest.approx(case.meta["n"] * U_DRIFT, rel=0.03)
    # momentum flux along x through the plane: sum over crossings of m u_x - unsigned in the reference, so both directions add:
    # n m (u^2 + kT/m) for a drifting Maxwellian
    kT_m = cases.kB * case.meta["Tw"] / case.meta["species"]["mass"]
    assert f["momentumFlux"] == pytest.approx(rho * (U_DRIFT ** 2 + kT_m), rel=0.05)
    # energy flux of a drifting monatomic Maxwellian: rho u (u^2/2 + 5/2 kT/m)
    assert f["energyFlux"] == pytest.approx(rho * U_DRIFT * (0.5 * U_DRIFT ** 2 + 2.5 * kT_m), rel=0.05)
    # reset: the next window starts from zero
    surf.calculateField(resetAtOutput=True)
    cl.evolve(5)
    assert surf.calculateField()["massFlux"] == pytest.approx(rho * U_DRIFT, rel=0.1)
    # the opposite direction flips the signed fluxes
    back = UniGasMassFluxSurface(cl, {"field": "Ar", "faceZone": zone, "fluxDirection": [-1.0, 0.0, 0.0]})
    back.begin(); cl.evolve(10)
    assert back.calculateField()["massFlux"] == pytest.approx(-rho * U_DRIFT, rel=0.06)


def test_tracker_rejects_bad_face_lists(OracleCloud):
    from unigasfoam_b200.cloud import UgfError
    case = drifting_channel()
    cl = case.make_cloud(OracleCloud)
    with pytest.raises(UgfError, match="out of range"):
        cl.setFaceTracker([case.mesh.n_faces])
    with pytest.raises(UgfError, match="twice"):
        cl.setFaceTracker([3, 3])


@pytest.mark.gpu
def test_gpu_face_tallies_equal_the_oracle(GpuCloud, OracleCloud):
    """Internal cross-section, the cyclic faces (booked on the partner face) and the specular wall faces (booked after
    the reflection): bit-identical trajectories, so the tallies agree to summation order."""
    case = drifting_channel()
    m = case.mesh
    dx = (m.cell_bb_max - m.cell_bb_min)[0, 0]
    zone = np.concatenate([plane_faces(m, 7 * dx), plane_faces(m, 13 * dx)] +
                          [np.arange(p.start, p.start + p.size) for p in m.patches if p.kind in ("cyclic", "wall")]).astype(np.int32)
    g, r = case.make_cloud(GpuCloud), case.make_cloud(OracleCloud)
    g.setFaceTracker(zone); r.setFaceTracker(zone)
    g.evolve(12); r.evolve(12)
    tg, tr = g.faceTracker(), r.faceTracker()
    assert np.abs(tr[:, 0, 0]).sum() > 1000 and (np.abs(tr[20:, 0, 0]).sum() > 0)
    assert np.allclose(tg, tr, rtol=1e-11, atol=1e-11 * np.abs(tr).max(axis=(0, 1), keepdims=True))
    assert np.array_equal(tg[:, :, 0], tr[:, :, 0])  # parcel counts are integers: exact
    tg2 = g.faceTracker(reset=True)
    assert np.array_equal(tg2, tg) and not g.faceTracker().any()


@pytest.mark.gpu
def test_gpu_weighted_tallies_and_wall_force(GpuCloud, OracleCloud):
    """Cell-weighted Couette flow: the tallies carry the parcels' weight factors; the force on the moving walls
    (uniGasForceSurface) is equal and opposite and matches the oracle."""
    def ramp(mesh):
        y = mesh.cell_centres[:, 1]
        return 0.7 + 0.8 * (y - y.min()) / (y.max() - y.min())
    case = cases.couette(nx=24, ny=12, ppc=30, Kn=0.5, cellWeightFactor=ramp)
    m = case.mesh
    dx = (m.cell_bb_max - m.cell_bb_min)[0, 0]
    zone = plane_faces(m, 9 * dx)
    g, r = case.make_cloud(GpuCloud), case.make_cloud(OracleCloud)
    g.setFaceTracker(zone); r.setFaceTracker(zone)
    g.evolve(30); r.evolve(30)
    tg, tr = g.faceTracker(), r.faceTracker()
    assert np.allclose(tg[:, :, 0], tr[:, :, 0], rtol=1e-9, atol=1e-9)           # weighted counts: sums of the factors
    assert np.abs(tr[:, 0, 0] - np.rint(tr[:, 0, 0])).max() > 1e-3              # ... which are not integers
    assert np.allclose(tg, tr, rtol=1e-6, atol=1e-6 * np.abs(tr).max(axis=(0, 1), keepdims=True))
    fg = [UniGasForceSurface(g, {"field": "Ar", "patch": p}).calculateField() for p in ("bottom", "top")]
    fr = [UniGasForceSurface(r, {"field": "Ar", "patch": p}).calculateField() for p in ("bottom", "top")]
    for a, b in zip(fg, fr):
        assert np.allclose(a, b, rtol=1e-6, atol=1e-9 * np.abs(b).max())
    # gas pressure pushes the walls apart; the shear drags each wall against its own motion (bottom moves -x, top +x)
    assert fr[0][1] < 0 < fr[1][1] and fr[0][0] > 0 > fr[1][0]
    assert abs(fr[0][1] + fr[1][1]) < 0.1 * abs(fr[1][1])
