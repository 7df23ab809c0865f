"""Host-side logic on CPU: mesh generator, decomposePar stand-in, run-time selection words, and that the C-ABI
library loads and exports every symbol include/ugf.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from unigasfoam_b200 import _capi, cases, mesh as ugmesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_box_mesh_geometry():
    m = ugmesh.box_mesh(4, 3, 2, 2.0, 1.5, 1.0)
    assert m.n_cells == 24 and m.n_internal == 3 * 3 * 2 + 4 * 2 * 2 + 4 * 3 * 1
    assert m.n_faces == m.n_internal + 2 * (3 * 2 + 4 * 2 + 4 * 3)
    np.testing.assert_allclose(m.cell_volumes, 0.5 * 0.5 * 0.5)
    # owner < neighbour, upper-triangular order (OpenFOAM convention)
    assert (m.owner[: m.n_internal] < m.neighbour).all()
    key = m.owner[: m.n_internal].astype(np.int64) * m.n_cells + m.neighbour
    assert (np.diff(key) > 0).all()
    # closed cells: outward area vectors sum to zero
    S = np.zeros((m.n_cells, 3))
    np.add.at(S, m.owner, m.face_areas)
    np.subtract.at(S, m.neighbour, m.face_areas[: m.n_internal])
    assert np.abs(S).max() < 1e-14
    # face area vectors point from owner to neighbour / out of the domain
    d = m.cell_centres[m.neighbour] - m.cell_centres[m.owner[: m.n_internal]]
    assert ((d * m.face_areas[: m.n_internal]).sum(1) > 0).all()
    bf = np.arange(m.n_internal, m.n_faces)
    assert (((m.face_centres[bf] - m.cell_centres[m.owner[bf]]) * m.face_areas[bf]).sum(1) > 0).all()
    assert np.diff(m.cell_face_offsets).tolist() == [6] * 24
    np.testing.assert_allclose(m.cell_bb_max - m.cell_bb_min, 0.5)


def test_cyclic_pairing_and_patch_cover():
    c = cases.couette(nx=6, ny=4, ppc=3)
    m = c.mesh
    left, right = m.patches[m.patch_index("left")], m.patches[m.patch_index("right")]
    assert left.kind == right.kind == "cyclic" and left.partner == m.patch_index("right")
    np.testing.assert_allclose(np.array(left.separation) + np.array(right.separation), 0, atol=1e-18)
    np.testing.assert_allclose(left.separation[0], c.meta["Lx"])
    assert (m.boundary_face_patch() >= 0).all()
    assert m.solution_d == (1, 1, 0)
    assert np.allclose(c.position[:, 2], c.position[0, 2])  # parcels sit on the mid-plane of the empty direction


def test_decompose_matches_faces_across_ranks():
    m = ugmesh.box_mesh(6, 4, 2, 3.0, 2.0, 1.0, {"xMin": ("l", "cyclic"), "xMax": ("r", "cyclic"), "yMin": ("b", "wall"), "yMax": ("t", "wall"),
                                                   "zMin": ("k", "wall"), "zMax": ("f", "wall")}, cyclic_pairs=[("xMin", "xMax")])
    subs = ugmesh.decompose(m, ugmesh.slab_partition(m, 3), 3)
    assert sum(s.n_cells for s in subs) == m.n_cells
    for r, s in enumerate(subs):
        assert (s.boundary_face_patch() >= 0).all()
        for p in s.patches:
            if p.kind != "processor":
                continue
            q = subs[p.partner].patches[p.peer_patch]
            assert q.partner == r and q.size == p.size
            a = s.face_centres[p.start : p.start + p.size] + np.array(p.separation)
            b = subs[p.partner].face_centres[q.start : q.start + q.size]
            np.testing.assert_allclose(a, b, atol=1e-12)  # same faces, same order, shifted by the separation
            np.testing.assert_allclose(s.face_areas[p.start : p.start + p.size], -subs[p.partner].face_areas[q.start : q.start + q.size])
        S = np.zeros((s.n_cells, 3))
        np.add.at(S, s.owner, s.face_areas)
        np.subtract.at(S, s.neighbour, s.face_areas[: s.n_internal])
        assert np.abs(S).max() < 1e-14


def test_decomposed_oracle_run_equals_single_domain(OracleCloud):
    """Collision-free tracking on a 2-way decomposition (in-process exchange through the migrate_* entry points)
    reproduces the undecomposed run bit for bit - faces inherit the parent geometry."""
    from unigasfoam_b200.cloud import UniGasCloud  # noqa: F401
    case = cases.closed_box(n=6, parcels=4000, seed=31, binary="noDSMCCollision")
    dx = case.meta["L"] / 6
    case.deltaT = 1.7 * dx / cases.most_probable_speed(300.0, cases.ARGON_GUIDE["mass"])
    single = case.make_cloud(OracleCloud)
    rank = ugmesh.slab_partition(case.mesh, 2)
    subs = ugmesh.decompose(case.mesh, rank, 2)
    clouds = []
    for r, sm in enumerate(subs):
        g2l = np.full(case.mesh.n_cells, -1)
        g2l[sm.cell_map] = np.arange(sm.n_cells)
        sel = rank[case.cell] == r
        cl = OracleCloud(sm, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=8000, seed=20261017 - r)
        cl.setParcels(case.position[sel], case.U[sel], g2l[case.cell[sel]])
        cl.setCellState(sigmaTcRMax=case.sigmaTcRMax)
        clouds.append(cl)
    for _ in range(4):
        single.evolve(1)
        for cl in clouds:
            cl.move()
        for _round in range(20):
            moved = 0
            packs = []
            for r, cl in enumerate(clouds):
                for pi, p in enumerate(subs[r].patches):
                    if p.kind == "processor":
                        buf, n = cl.migratePack(pi)
                        rec = np.ctypeslib.as_array(buf, shape=(n * _capi.UGF_MIGRATE_STRIDE,)).copy() if n else np.empty(0)
                        packs.append((p.partner, p.peer_patch, rec, n))
                        moved += n
            if moved == 0:
                break
            for dst, patch, rec, n in packs:
                if n:
                    clouds[dst].migrateUnpack(patch, rec.ctypes.data_as(C.POINTER(C.c_double)), n)
            for cl in clouds:
                cl.moveReceived()
        for cl in clouds:
            cl.buildCellOccupancy(); cl.collide(); cl.accumulateFields(); cl.endStep()
    ps = single.parcels()
    got = []
    for r, cl in enumerate(clouds):
        p = cl.parcels()
        live = p["cell"] >= 0
        got.append(np.column_stack([p["position"][live], p["U"][live], subs[r].cell_map[p["cell"][live]]]))
    got = np.concatenate(got)
    ref = np.column_stack([ps["position"], ps["U"], ps["cell"]])
    assert got.shape == ref.shape
    key = lambda a: a[np.lexsort(a.T[::-1])]
    assert np.array_equal(key(got), key(ref))


def test_runtime_selection_words():
    from unigasfoam_b200 import UgfError
    from unigasfoam_b200.cloud import _lookup
    assert _capi.BINARY_MODEL["LarsenBorgnakkeVariableHardSphere"] == 3
    assert _capi.BGK_MODEL["unifiedStochasticParticleSBGK"] == 4
    assert _capi.WALL_MODEL["uniGasMixedDiffuseSpecularWallPatch"] == 3
    with pytest.raises(UgfError, match="Unknown bgkCollisionModel"):
        _lookup(_capi.BGK_MODEL, "ellipsoidalBGK", "bgkCollisionModel")


def test_header_and_bindings_agree():
    """Every function include/ugf.h declares is bound in _capi.SIGNATURES and vice versa; constants match."""
    hdr = open(os.path.join(ROOT, "include", "ugf.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+ugf_(\w+)\(", hdr, re.M))
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    for name in ("UGF_ABI_VERSION", "UGF_NMOM", "UGF_NBM", "UGF_NFIELD", "UGF_NWALLFIELD", "UGF_MIGRATE_STRIDE", "UGF_MAX_SPECIES"):
        assert int(re.search(rf"#define {name} (\d+)", hdr).group(1)) == getattr(_capi, name)


def test_libugf_loads_exports_every_symbol_and_fails_loudly_without_gpu():
    """The product library must be in-tree, export the whole C ABI, and refuse to run without CUDA (no CPU fallback)."""
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    api = _capi.Api(g.LIB, "ugf_")  # resolves all 36 entry points or raises
    assert api.abi_version() == _capi.UGF_ABI_VERSION
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device error path cannot be exercised")
    cfg = _capi.Config()
    cfg.abiVersion = _capi.UGF_ABI_VERSION
    cfg.parcelCapacity = 1024
    cfg.partnerModel = 1
    h = _capi.H()
    assert api.create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in api.last_error(None)
    from unigasfoam_b200 import UgfError
    from unigasfoam_b200.cloud import UniGasCloud
    with pytest.raises(UgfError, match="no CUDA device"):
        cases.closed_box(n=3, parcels=100).make_cloud(UniGasCloud)


def test_oracle_sort_is_stable_counting_sort(OracleCloud):
    case = cases.closed_box(n=5, parcels=3000, seed=32)
    rng = np.random.default_rng(3)
    perm = rng.permutation(case.n_parcels)
    case.position, case.U, case.cell = case.position[perm], case.U[perm], case.cell[perm]
    cl = case.make_cloud(OracleCloud)
    cl.buildCellOccupancy()
    off, ids = cl.cellOccupancy()
    assert np.array_equal(ids, np.argsort(case.cell, kind="stable"))
    assert np.array_equal(np.diff(off), np.bincount(case.cell, minlength=case.mesh.n_cells))


def test_weighted_partition_balances_parcels_not_cells():
    """decomposeParDict `weightField uniGasRhoNMean_Ar` (hypersonicCylinder tutorial): on a cell-weighted-free cylinder case
    the parcels pile up in the small cells near the body, equal-cell slabs are unbalanced, weighted slabs are not."""
    from unigasfoam_b200 import cases, mesh as ugmesh
    case = cases.cylinder(nr=96, ntheta=8, ppc=2)
    m = case.mesh
    x = m.cell_centres
    w = np.exp(-5.0 * (np.hypot(x[:, 0], x[:, 1]) - case.meta["r0"]) / (case.meta["r1"] - case.meta["r0"]))  # a shock-layer-like pile-up
    for ranks in (2, 4, 8):
        plain = ugmesh.slab_partition(m, ranks, axis=0)
        wtd = ugmesh.weighted_slab_partition(m, ranks, w, axis=0)
        per = lambda part: np.bincount(part, weights=w, minlength=ranks)
        assert set(wtd) == set(range(ranks)) and (np.diff(wtd.reshape(m.shape[::-1])[0, 0]) >= 0).all()   # contiguous slabs, none empty
        assert ugmesh.load_imbalance(per(wtd)) < 0.5 * ugmesh.load_imbalance(per(plain))
        assert ugmesh.load_imbalance(per(wtd)) < 25.0
        subs = ugmesh.decompose(m, wtd, ranks)
        assert sum(s.n_cells for s in subs) == m.n_cells
    assert np.array_equal(ugmesh.weighted_slab_partition(m, 4, np.ones(m.n_cells), axis=0), ugmesh.slab_partition(m, 4, axis=0))
    assert np.array_equal(ugmesh.weighted_slab_partition(m, 4, np.zeros(m.n_cells), axis=0), ugmesh.slab_partition(m, 4, axis=0))
    assert ugmesh.load_imbalance([100, 100, 100, 100]) == 0.0 and ugmesh.load_imbalance([150, 50]) == 50.0
    with pytest.raises(ValueError):
        ugmesh.weighted_slab_partition(m, 4, np.ones(3))
    # all the weight in one layer: every rank still gets cells
    spike = np.zeros(m.n_cells); spike[:m.shape[0]:m.shape[0]] = 1.0
    assert set(ugmesh.weighted_slab_partition(m, 8, spike, axis=0)) == set(range(8))
