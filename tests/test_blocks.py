"""Per-rank block builders of the full-size BASELINE cases (cases.cylinder_block, cases.blunt_body_block), the sphere-cone grid,
the torch parcel fill and the in-process subdomain driver of bench.py's reference arm - CPU only, the oracle as the engine."""
import numpy as np

from unigasfoam_b200 import cases, mesh as M
from unigasfoam_b200.exchange import LocalSubdomains


def test_sphere_cone_faces_planar_and_cells_convex():
    pm = M.sphere_cone_map(10, 18, 6)
    kinds = {"xMin": ("body", "wall"), "xMax": ("inlet", "patch"), "yMin": ("axis", "symmetry"), "yMax": ("outlet", "patch"),
             "zMin": ("symA", "symmetryPlane"), "zMax": ("symB", "symmetryPlane")}
    m = M.structured_block(10, 18, 6, pm, kinds)
    assert (m.cell_volumes > 0).all() and np.isfinite(m.face_centres).all() and np.isfinite(m.cell_centres).all()
    A = np.linalg.norm(m.face_areas, axis=1)
    P = m.points[m.face_points.reshape(m.n_faces, 4)]
    n = m.face_areas / np.where(A > 0, A, 1.0)[:, None]
    dev = np.abs(np.einsum("fkj,fj->fk", P - m.face_centres[:, None, :], n)).max()
    assert dev < 1e-14  # isosceles trapezoids and meridian-plane quads: planar
    ax = m.patches[m.patch_index("axis")]
    assert (A[ax.start:ax.start + ax.size] == 0).all()  # the j = 0 side collapses onto the axis exactly
    hc = M.hex_corners(m)
    worst = 0.0
    for c in range(m.n_cells):
        for f in m.cell_faces[m.cell_face_offsets[c]:m.cell_face_offsets[c + 1]]:
            S = m.face_areas[f] * (1.0 if m.owner[f] == c else -1.0)
            worst = max(worst, ((m.points[hc[c]] - m.face_centres[f]) @ S / max(A[f], 1e-300)).max())
    assert worst < 1e-14  # every corner of a cell lies inside all of its face planes


def test_cylinder_blocks_tile_the_single_block_mesh():
    nr, nt = 12, 30
    whole = cases.cylinder_block(0, 1, nr=nr, ntheta=nt, parcels=5000, device="cpu")
    parts = [cases.cylinder_block(r, 3, nr=nr, ntheta=nt, parcels=5000, device="cpu") for r in range(3)]
    assert sum(p.mesh.n_cells for p in parts) == whole.mesh.n_cells == nr * nt
    vol = np.concatenate([p.mesh.cell_volumes for p in parts])
    np.testing.assert_allclose(vol, whole.mesh.cell_volumes, rtol=1e-13)  # same point map; sums run in another face order next to a cut
    assert len({p.deltaT for p in parts} | {whole.deltaT}) == 1
    assert len({p.uniGasProperties["nEquivalentParticles"] for p in parts}) == 1
    for r, p in enumerate(parts):
        for q in p.mesh.patches:
            if q.kind != "processor":
                continue
            peer = parts[q.partner].mesh
            back = [k for k in peer.patches if k.kind == "processor" and k.partner == r]
            assert len(back) == 1 and back[0].size == q.size
            a = p.mesh.face_centres[q.start:q.start + q.size]
            b = peer.face_centres[back[0].start:back[0].start + back[0].size]
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-15)
            assert (p.mesh.face_areas[q.start:q.start + q.size, 2] == 0).all()  # in-plane faces: z component snapped to exactly zero
            np.testing.assert_allclose(p.mesh.face_areas[q.start:q.start + q.size], -peer.face_areas[back[0].start:back[0].start + back[0].size],
                                       rtol=1e-13, atol=1e-19)
    # every rank knows every patch (zero faces where it does not touch it), so one boundariesDict serves all ranks
    for p in parts:
        for name in ("cylinder", "inlet", "outlet", "axisDown", "axisUp"):
            p.mesh.patch_index(name)
    assert parts[0].mesh.patches[parts[0].mesh.patch_index("inlet")].size == 0
    assert parts[2].mesh.patches[parts[2].mesh.patch_index("outlet")].size == 0


def test_fill_parcels_torch_statistics():
    c = cases.cylinder_block(0, 1, nr=10, ntheta=24, parcels=200000, device="cpu")
    m = c.mesh
    cnt = np.bincount(c.cell, minlength=m.n_cells)
    FN = c.uniGasProperties["nEquivalentParticles"]
    expect = c.meta["n"] * m.cell_volumes / FN
    assert np.abs(cnt - expect).max() <= 1.0 + 1e-9  # stochastic rounding of n V / F_N
    assert (np.diff(c.cell) >= 0).all()              # cell-major
    lo, hi = m.cell_bb_min[c.cell], m.cell_bb_max[c.cell]
    assert ((c.position >= lo - 1e-12) & (c.position <= hi + 1e-12)).all()
    assert np.ptp(c.position[:, 2]) == 0.0           # empty direction: on the mid-plane
    # inside the cell itself, not only its bounding box: inside all four in-plane face planes
    for f in range(m.n_faces):
        S, Cf = m.face_areas[f], m.face_centres[f]
        if S[2] != 0.0:
            continue
        own = c.cell == m.owner[f]
        if own.any():
            assert (((c.position[own] - Cf) @ S) <= 1e-12 * np.linalg.norm(S)).all()
    sig2 = cases.kB * c.meta["T_inf"] / c.meta["species"]["mass"]
    assert abs(c.U[:, 0].mean() - c.meta["U_inf"]) < 5 * np.sqrt(sig2 / c.n_parcels)
    assert abs(c.U[:, 1].var() / sig2 - 1.0) < 0.02


def test_blunt_body_eight_blocks_conserve_parcels(OracleCloud):
    """The 8-block decomposition of configs[4] (4 along the body x 2 in azimuth) on a tiny grid: inflow, outflow, diffuse body,
    symmetry planes, wedge cells on the axis, cell weighting with clones carried across processor patches, corner crossings -
    nothing is lost, nothing gets stuck, every parcel ends inside the cell it claims."""
    N = 8
    cs = [cases.blunt_body_block(rank=r, n_ranks=N, n_eta=8, n_s=16, n_phi=8, ppc=10, device="cpu") for r in range(N)]
    assert len({c.deltaT for c in cs}) == 1 and len({c.uniGasProperties["nEquivalentParticles"] for c in cs}) == 1
    cl = [c.make_cloud(OracleCloud, rank=r, nRanks=N, parcelCapacity=4 * c.n_parcels) for r, c in enumerate(cs)]
    L = LocalSubdomains(cl, [c.mesh for c in cs])
    n0 = L.size()
    ins = dele = cloned = wdel = mig = 0
    for _ in range(12):
        L.evolve(1, inflow=True)
        for c in cl:
            k = c.counters()
            ins += k["inserted"]; dele += k["deleted"]; cloned += k["cloned"]; wdel += k["weightDeleted"]; mig += k["migrated"]
            assert k["stuck"] == 0
    assert L.size() == n0 + ins - dele + cloned - wdel
    assert ins > 0 and dele > 0 and cloned > 0 and mig > 100 and L.rounds >= 12
    for c, case in zip(cl, cs):
        p = c.parcels()
        lo, hi = case.mesh.cell_bb_min[p["cell"]], case.mesh.cell_bb_max[p["cell"]]
        assert ((p["position"] >= lo - 1e-9) & (p["position"] <= hi + 1e-9)).all()


def test_local_subdomains_equal_the_single_domain_run(OracleCloud):
    """bench.py's reference arm at N ranks = N oracle subdomains in one process: on the periodic Couette channel the two-slab
    run ends where the single-domain run of the same channel does (collision-free: bit for bit)."""
    one = cases.couette(nx=24, ny=10, ppc=12, binary="noDSMCCollision")
    for e in one.boundariesDict["uniGasPatchBoundaries"]:  # wall draws are keyed by the parcel's array index, which differs
        e["boundaryModel"] = "uniGasSpecularWallPatch"     # between the two runs: specular walls keep the comparison exact
    from unigasfoam_b200.mesh import decompose, slab_partition
    subs = decompose(one.mesh, slab_partition(one.mesh, 2, axis=0), 2)
    clouds = []
    for r, sm in enumerate(subs):
        inv = {g: l for l, g in enumerate(sm.cell_map)}
        sel = np.nonzero(np.isin(one.cell, sm.cell_map))[0]
        c = cases.Case("part", sm, one.uniGasProperties, one.boundariesDict, one.deltaT, one.position[sel], one.U[sel],
                       np.array([inv[g] for g in one.cell[sel]], np.int32), one.typeId[sel], None, one.sigmaTcRMax)
        clouds.append(c.make_cloud(OracleCloud, rank=r, nRanks=2, parcelCapacity=2 * one.n_parcels))
    L = LocalSubdomains(clouds, subs)
    ref = one.make_cloud(OracleCloud)
    L.evolve(6)
    ref.evolve(6)
    got = np.concatenate([np.column_stack([c.parcels()["position"], c.parcels()["U"]]) for c in clouds])
    want = np.column_stack([ref.parcels()["position"], ref.parcels()["U"]])
    assert got.shape == want.shape
    key = lambda a: a[np.lexsort(a.T[::-1])]
    np.testing.assert_array_equal(key(got), key(want))


def test_mixed_wall_shear_scales_with_the_diffuse_fraction_on_the_oracle(OracleCloud):
    """uniGasMixedDiffuseSpecularWallPatch.C:77-97 pinned to its closed-form consequence on the CPU restatement (the GPU test of
    the same name in test_gpu_paths.py holds the CUDA path to it): wall shear = diffuseFraction x the fully diffuse shear."""
    tau = {}
    for f in (1.0, 0.5, 0.0):
        case = cases.couette(nx=96, ny=16, ppc=60, Kn=0.5, binary="noDSMCCollision")  # ~10 k wall hits per wall: 2-3 % noise
        for e in case.boundariesDict["uniGasPatchBoundaries"]:
            old = e.pop("uniGasDiffuseWallPatchProperties")
            e["boundaryModel"] = "uniGasMixedDiffuseSpecularWallPatch"
            e["uniGasMixedDiffuseSpecularWallPatchProperties"] = dict(old, diffuseFraction=f)
        cl = case.make_cloud(OracleCloud)
        cl.evolve(6)
        fd = cl.fields()["fD"]
        nI = case.mesh.n_internal
        p = case.mesh.patches[case.mesh.patch_index("bottom")]
        tau[f] = fd[p.start - nI:p.start - nI + p.size, 0].mean()
    assert abs(tau[0.0]) < 0.02 * abs(tau[1.0])
    assert abs(tau[0.5] / tau[1.0] - 0.5) < 0.05
