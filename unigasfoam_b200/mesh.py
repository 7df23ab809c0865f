"""polyMesh stand-ins for blockMesh and decomposePar (host side, numpy).

`PolyMesh` holds exactly what OpenFOAM's polyMesh exposes to uniGasFoam's particle loop
(points, faces, owner, neighbour, patches) plus the derived geometry primitiveMesh
computes (faceAreas, faceCentres, cellVolumes, cellCentres), laid out as include/ugf.h's
`ugf_mesh` wants it.  `structured_block` is the blockMesh stand-in (one hex block with an
arbitrary point map); `decompose` is the decomposePar stand-in (any cell->rank map,
processor patches on the cuts, cyclic pairs split into processorCyclic patches).

Reference: tutorials/uniGasFoam/*/system/blockMeshDict, decomposeParDict; face ordering
follows OpenFOAM's (internal faces upper-triangular by owner/neighbour, then patches).
"""
from dataclasses import dataclass, field
import ctypes as C

import numpy as np

from . import _capi


@dataclass
class Patch:
    name: str
    kind: str  # wall | symmetry | symmetryPlane | cyclic | empty | processor | patch
    start: int
    size: int
    partner: int = -1  # cyclic: partner patch index; processor: peer rank
    separation: tuple = (0.0, 0.0, 0.0)
    peer_patch: int = -1  # processor: index of the matching patch on the peer rank
    tag: tuple = ()


@dataclass
class PolyMesh:
    points: np.ndarray  # [nPoints,3]
    face_point_offsets: np.ndarray  # [nFaces+1]
    face_points: np.ndarray
    owner: np.ndarray  # [nFaces]
    neighbour: np.ndarray  # [nInternal]
    patches: list
    solution_d: tuple = (1, 1, 1)
    # derived
    face_areas: np.ndarray = None
    face_centres: np.ndarray = None
    cell_volumes: np.ndarray = None
    cell_centres: np.ndarray = None
    cell_face_offsets: np.ndarray = None
    cell_faces: np.ndarray = None
    cell_bb_min: np.ndarray = None
    cell_bb_max: np.ndarray = None
    shape: tuple = None  # (nx, ny, nz) for structured blocks
    cell_map: np.ndarray = None  # decompose: local cell -> global cell
    meta_axis_aligned: bool = False  # every cell is an axis-aligned box
    cell_quads: np.ndarray = None  # extruded 2-D meshes: [nCells,4,2] xy corners, counter-clockwise
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_cells(self):
        return int(self.owner.max()) + 1 if self.cell_volumes is None else len(self.cell_volumes)

    @property
    def n_faces(self):
        return len(self.owner)

    @property
    def n_internal(self):
        return len(self.neighbour)

    @property
    def n_boundary_faces(self):
        return self.n_faces - self.n_internal

    def patch_index(self, name):
        for i, p in enumerate(self.patches):
            if p.name == name:
                return i
        raise KeyError(name)

    def boundary_face_patch(self):
        out = np.full(self.n_boundary_faces, -1, np.int32)
        for i, p in enumerate(self.patches):
            out[p.start - self.n_internal : p.start - self.n_internal + p.size] = i
        return out

    # -- primitiveMesh geometry ------------------------------------------------
    def compute_geometry(self, n_cells=None, snap=False):
        pts = self.points
        off = self.face_point_offsets
        nF = len(self.owner)
        nv = np.diff(off)
        Sf = np.zeros((nF, 3))
        Cf = np.zeros((nF, 3))
        fmin = np.zeros((nF, 3))
        fmax = np.zeros((nF, 3))
        kinds = np.unique(nv) if nF == 0 or nv.min() != nv.max() else nv[:1]
        for k in kinds:  # faces grouped by vertex count (all quads for hex blocks)
            all_k = len(kinds) == 1
            sel = slice(None) if all_k else np.nonzero(nv == k)[0]
            if all_k:
                P = pts[self.face_points.reshape(nF, k)]  # [n,k,3]
            else:
                idx = off[sel][:, None] + np.arange(k)[None, :]
                P = pts[self.face_points[idx]]
            fmin[sel] = P.min(axis=1)
            fmax[sel] = P.max(axis=1)
            if k == 3:
                n2 = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
                Sf[sel] = 0.5 * n2
                Cf[sel] = P.mean(axis=1)
                continue
            fc = P.mean(axis=1)
            sumN = np.zeros((len(fc), 3))
            sumA = np.zeros(len(fc))
            sumAc = np.zeros((len(fc), 3))
            for t in range(k):
                a, b = P[:, t], P[:, (t + 1) % k]
                n = np.cross(b - a, fc - a)
                c = a + b + fc
                mag = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2])
                sumN += n
                sumA += mag
                sumAc += mag[:, None] * c
            Sf[sel] = 0.5 * sumN
            flat = sumA == 0.0  # collapsed faces (all vertices on one line, e.g. on the axis of a revolved block): never crossed
            if flat.any():
                sumA = np.where(flat, 1.0, sumA)
                sumAc = np.where(flat[:, None], 3.0 * fc, sumAc)
            Cf[sel] = sumAc / (3.0 * sumA[:, None])
        if snap:  # round-off residue of a component that is zero by construction (the z component of the side faces of an
            # extruded 2-D block, 1e-20 of the area): exactly zero, so that such meshes qualify for the packed 2-D cell record
            A = np.sqrt((Sf * Sf).sum(1))
            Sf[np.abs(Sf) < 1e-14 * A[:, None]] = 0.0
        self.face_areas, self.face_centres = Sf, Cf
        nI = len(self.neighbour)
        nC = n_cells if n_cells is not None else int(self.owner.max()) + 1
        # cell -> faces CSR, faces ascending within a cell
        if nI == 0 or (np.diff(self.owner[:nI]) >= 0).all():
            # upper-triangular order: the faces a cell is neighbour of precede the faces it owns
            cells = np.concatenate([self.neighbour, self.owner])
            faces = np.concatenate([np.arange(nI), np.arange(nF)])
            order = np.argsort(cells, kind="stable")
            cf = faces[order]
            # boundary faces of a cell follow its internal ones, but a cell's owned internal faces may interleave with
            # nothing else: only the (neighbour-of, owner-of) split needs the upper-triangular property
            chk = cells[order]
            asc = (np.diff(cf) > 0) | (np.diff(chk) != 0)
            if not asc.all():
                order = np.lexsort((faces, cells))
                cf = faces[order]
            self.cell_faces = cf.astype(np.int32)
        else:
            cells = np.concatenate([self.owner, self.neighbour])
            faces = np.concatenate([np.arange(nF), np.arange(nI)])
            order = np.lexsort((faces, cells))
            self.cell_faces = faces[order].astype(np.int32)
        counts = np.bincount(cells, minlength=nC)
        self.cell_face_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        # cell centres / volumes by pyramid decomposition (scatter-adds in face order: owner part, then neighbour part)
        on = np.concatenate([self.owner, self.neighbour])

        def scatter3(a_own, a_nei):
            w = np.concatenate([a_own, a_nei])
            return np.stack([np.bincount(on, weights=w[:, d], minlength=nC) for d in range(3)], axis=1)

        cEst = scatter3(Cf, Cf[:nI])
        cEst /= counts[:, None]
        pyrO = np.einsum("ij,ij->i", Sf, Cf - cEst[self.owner])
        pcO = 0.75 * Cf + 0.25 * cEst[self.owner]
        pyrN = np.einsum("ij,ij->i", Sf[:nI], cEst[self.neighbour] - Cf[:nI])
        pcN = 0.75 * Cf[:nI] + 0.25 * cEst[self.neighbour]
        vol3 = np.bincount(on, weights=np.concatenate([pyrO, pyrN]), minlength=nC)
        cc = scatter3(pyrO[:, None] * pcO, pyrN[:, None] * pcN)
        self.cell_centres = cc / vol3[:, None]
        self.cell_volumes = vol3 / 3.0
        # bounding box over cellPoints (noTimeCounter.C:112-127): min / max over the vertices of the cell's faces
        starts = self.cell_face_offsets[:-1].astype(np.int64)
        self.cell_bb_min = np.minimum.reduceat(fmin[self.cell_faces], starts, axis=0)
        self.cell_bb_max = np.maximum.reduceat(fmax[self.cell_faces], starts, axis=0)
        return self

    # -- C view -------------------------------------------------------------------
    def as_c(self):
        """ugf_mesh struct over contiguous copies (kept alive on self)."""
        def f64(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_double))

        def i32(a):
            a = np.ascontiguousarray(a, dtype=np.int32)
            self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int32))

        self._keep.clear()
        m = _capi.Mesh()
        m.nCells, m.nFaces, m.nInternalFaces = self.n_cells, self.n_faces, self.n_internal
        m.nPatches, m.nPoints = len(self.patches), len(self.points)
        m.owner, m.neighbour = i32(self.owner), i32(self.neighbour)
        m.faceAreas, m.faceCentres = f64(self.face_areas), f64(self.face_centres)
        m.cellFaceOffsets, m.cellFaces = i32(self.cell_face_offsets), i32(self.cell_faces)
        m.cellVolumes, m.cellCentres = f64(self.cell_volumes), f64(self.cell_centres)
        m.cellBbMin, m.cellBbMax = f64(self.cell_bb_min), f64(self.cell_bb_max)
        m.patchStart = i32([p.start for p in self.patches])
        m.patchSize = i32([p.size for p in self.patches])
        m.patchKind = i32([_capi.PATCH_KIND[p.kind] for p in self.patches])
        m.patchPartner = i32([p.partner for p in self.patches])
        m.patchSeparation = f64(np.array([p.separation for p in self.patches], dtype=np.float64).reshape(-1, 3))
        m.points = f64(self.points)
        m.facePointOffsets, m.facePoints = i32(self.face_point_offsets), i32(self.face_points)
        return m


def structured_block(nx, ny, nz, point_map, patch_kinds, cyclic_pairs=(), solution_d=(1, 1, 1), snap=False):
    """One hex block of nx*ny*nz cells.

    point_map(I, J, K) -> (x, y, z) arrays for integer vertex indices.
    patch_kinds: dict side -> (name, kind) for sides xMin,xMax,yMin,yMax,zMin,zMax.
    cyclic_pairs: iterable of (sideA, sideB) made cyclic partners (translational).
    """
    K, J, I = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    X, Y, Z = point_map(I.astype(np.float64), J.astype(np.float64), K.astype(np.float64))
    points = np.stack([np.broadcast_to(X, I.shape), np.broadcast_to(Y, I.shape), np.broadcast_to(Z, I.shape)], axis=-1).reshape(-1, 3)

    def pid(i, j, k):
        return i + (nx + 1) * (j + (ny + 1) * k)

    def cid(i, j, k):
        return i + nx * (j + ny * k)

    def grid(ni, nj, nk):
        k, j, i = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
        return i.ravel(), j.ravel(), k.ravel()

    def xface(i, j, k):  # normal +x
        return np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], axis=1)

    def yface(i, j, k):  # normal +y
        return np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], axis=1)

    def zface(i, j, k):  # normal +z
        return np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], axis=1)

    own, nei, fpts = [], [], []
    i, j, k = grid(nx - 1, ny, nz)
    own.append(cid(i, j, k)); nei.append(cid(i + 1, j, k)); fpts.append(xface(i + 1, j, k))
    i, j, k = grid(nx, ny - 1, nz)
    own.append(cid(i, j, k)); nei.append(cid(i, j + 1, k)); fpts.append(yface(i, j + 1, k))
    i, j, k = grid(nx, ny, nz - 1)
    own.append(cid(i, j, k)); nei.append(cid(i, j, k + 1)); fpts.append(zface(i, j, k + 1))
    own = np.concatenate(own); nei = np.concatenate(nei); fpts = np.concatenate(fpts)
    order = np.lexsort((nei, own))  # upper-triangular ordering
    own, nei, fpts = own[order], nei[order], fpts[order]
    n_internal = len(own)

    sides = {}
    j, k = grid(1, ny, nz)[1:]
    z0 = np.zeros_like(j)
    sides["xMin"] = (cid(z0, j, k), xface(z0, j, k)[:, ::-1])
    sides["xMax"] = (cid(z0 + nx - 1, j, k), xface(z0 + nx, j, k))
    i, _, k = grid(nx, 1, nz)
    z0 = np.zeros_like(i)
    sides["yMin"] = (cid(i, z0, k), yface(i, z0, k)[:, ::-1])
    sides["yMax"] = (cid(i, z0 + ny - 1, k), yface(i, z0 + ny, k))
    i, j, _ = grid(nx, ny, 1)
    z0 = np.zeros_like(i)
    sides["zMin"] = (cid(i, j, z0), zface(i, j, z0)[:, ::-1])
    sides["zMax"] = (cid(i, j, z0 + nz - 1), zface(i, j, z0 + nz))

    patches, b_own, b_fpts = [], [], []
    start = n_internal
    side_patch = {}
    for side in ("xMin", "xMax", "yMin", "yMax", "zMin", "zMax"):
        name, kind = patch_kinds[side]
        o, f = sides[side]
        side_patch[side] = len(patches)
        patches.append(Patch(name=name, kind=kind, start=start, size=len(o)))
        b_own.append(o); b_fpts.append(f)
        start += len(o)
    owner = np.concatenate([own] + b_own).astype(np.int32)
    fpts = np.concatenate([fpts] + b_fpts).astype(np.int32)
    mesh = PolyMesh(
        points=points,
        face_point_offsets=(4 * np.arange(len(owner) + 1)).astype(np.int32),
        face_points=fpts.ravel(),
        owner=owner,
        neighbour=nei.astype(np.int32),
        patches=patches,
        solution_d=tuple(solution_d),
        shape=(nx, ny, nz),
    )
    mesh.compute_geometry(n_cells=nx * ny * nz, snap=snap)
    for a, b in cyclic_pairs:
        pa, pb = side_patch[a], side_patch[b]
        A, B = patches[pa], patches[pb]
        assert A.kind == "cyclic" and B.kind == "cyclic" and A.size == B.size
        sep = mesh.face_centres[B.start : B.start + B.size] - mesh.face_centres[A.start : A.start + A.size]
        sepm = sep.mean(axis=0)
        assert np.allclose(sep, sepm, rtol=0, atol=1e-9 * (np.abs(sepm).max() + 1e-300)), "cyclic pair is not translational"
        A.partner, B.partner = pb, pa
        A.separation, B.separation = tuple(sepm), tuple(-sepm)
    return mesh


def split_patch(mesh, name, n_first, name_a, name_b, kind_a=None, kind_b=None):
    """Split patch `name` into its first n_first faces and the rest (blockMesh lists several face sets per side)."""
    pi = mesh.patch_index(name)
    p = mesh.patches[pi]
    assert 0 < n_first < p.size and p.kind != "cyclic"
    a = Patch(name_a, kind_a or p.kind, p.start, n_first)
    b = Patch(name_b, kind_b or p.kind, p.start + n_first, p.size - n_first)
    mesh.patches[pi : pi + 1] = [a, b]
    for q in mesh.patches:  # cyclic partners are stored as patch indices
        if q.kind == "cyclic" and q.partner > pi:
            q.partner += 1
    return mesh


def merge_patches(mesh, name_a, name_b, new_name=None):
    """Make the faces of patch `name_b` part of patch `name_a` (blockMesh lists several face sets under one patch, e.g. the
    left and the top side of a block as `inlet`): the boundary faces are reordered so that b's follow a's, before the
    geometry is computed.  Neither patch may be cyclic."""
    ia, ib = mesh.patch_index(name_a), mesh.patch_index(name_b)
    pa, pb = mesh.patches[ia], mesh.patches[ib]
    assert pa.kind != "cyclic" and pb.kind != "cyclic" and pa.kind == pb.kind and ia != ib
    nF = mesh.n_faces
    order = list(range(mesh.n_internal))
    new_patches = []
    for k, p in enumerate(mesh.patches):
        if k == ib:
            continue
        start = len(order)
        order.extend(range(p.start, p.start + p.size))
        size = p.size
        if k == ia:
            order.extend(range(pb.start, pb.start + pb.size))
            size += pb.size
        new_patches.append(Patch(new_name if (k == ia and new_name) else p.name, p.kind, start, size, p.partner, p.separation))
    order = np.asarray(order)
    assert len(order) == nF
    off = mesh.face_point_offsets
    sizes = np.diff(off)[order]
    new_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    new_fp = np.concatenate([mesh.face_points[off[f]:off[f + 1]] for f in order]).astype(np.int32)
    mesh.face_point_offsets, mesh.face_points = new_off, new_fp
    mesh.owner = mesh.owner[order]
    # cyclic partners are patch indices: shift those above the removed patch
    for q in new_patches:
        if q.kind == "cyclic" and q.partner > ib:
            q.partner -= 1
    mesh.patches = new_patches
    if mesh.face_areas is not None:
        n_cells = mesh.n_cells
        mesh.compute_geometry(n_cells)
    return mesh


def _graded(t, g):
    """blockMesh simpleGrading: position fraction of vertex fraction t when last / first cell size = g."""
    g = float(g)
    return t if abs(g - 1.0) < 1e-12 else (g ** t - 1.0) / (g - 1.0)


def plate_mesh(nx=(50, 100, 75), ny=200, x_breaks=(0.0, 0.5e-3, 1.5e-3, 3.0e-3), height=5.0e-3, lz=0.02e-3, gx=(1.0, 1.0, 4.0), gy=5.0):
    """The mesh of tutorials/uniGasFoam/supersonicPlate (system/blockMeshDict): three hex blocks side by side along x
    (cell counts nx, gradings gx), ny cells in y with grading gy, one cell in z.  Patches as in the tutorial: inlet
    (left side and top, type patch), outlet (right, patch), symmetryFront / plate (wall) / symmetryBack along the
    bottom, emptyBoundaries."""
    nxt = int(sum(nx))
    edges = np.concatenate([[0], np.cumsum(nx)])

    def xmap(I):
        x = np.zeros_like(I)
        for b in range(len(nx)):
            sel = (I >= edges[b]) & (I <= edges[b + 1])
            t = (I - edges[b]) / nx[b]
            frac = t if abs(gx[b] - 1.0) < 1e-12 else (gx[b] ** t - 1.0) / (gx[b] - 1.0)
            x = np.where(sel, x_breaks[b] + (x_breaks[b + 1] - x_breaks[b]) * frac, x)
        return x

    def pm(I, J, K):
        t = J / ny
        fy = t if abs(gy - 1.0) < 1e-12 else (gy ** t - 1.0) / (gy - 1.0)
        return xmap(I), height * fy, lz * (K - 0.5)

    kinds = {"xMin": ("inlet", "patch"), "xMax": ("outlet", "patch"), "yMin": ("bottom", "symmetry"), "yMax": ("inletTop", "patch"),
             "zMin": ("emptyBack", "empty"), "zMax": ("emptyBoundaries", "empty")}
    m = structured_block(nxt, ny, 1, pm, kinds, solution_d=(1, 1, 0))
    split_patch(m, "bottom", int(nx[0]), "symmetryFront", "rest")
    split_patch(m, "rest", int(nx[1]), "plate", "symmetryBack", kind_a="wall")
    merge_patches(m, "inlet", "inletTop")
    merge_patches(m, "emptyBoundaries", "emptyBack")
    m.meta_axis_aligned = True
    return m


def half_annulus_mesh(nr, ntheta, r0, r1, lz, grading=5.0):
    """Half O-grid around a cylinder on the symmetry axis: the topology of tutorials/uniGasFoam/hypersonicCylinder
    (system/blockMeshDict), as one block.  i runs radially outwards (last/first cell size = grading), j in theta
    from 0 (downstream axis) to pi (upstream axis), one cell in z.  Patches: cylinder (wall), outlet / inlet (outer
    arc, type patch), axisDown / axisUp (symmetryPlane), back / front (empty)."""
    g = float(grading)

    def pm(I, J, K):
        t = I / nr
        frac = t if abs(g - 1.0) < 1e-12 else (g ** t - 1.0) / (g - 1.0)
        r = r0 + (r1 - r0) * frac
        th = np.pi * J / ntheta
        return r * np.cos(th), r * np.sin(th), lz * (K - 0.5)

    kinds = {"xMin": ("cylinder", "wall"), "xMax": ("outer", "patch"), "yMin": ("axisDown", "symmetryPlane"), "yMax": ("axisUp", "symmetryPlane"),
             "zMin": ("back", "empty"), "zMax": ("front", "empty")}
    m = structured_block(nr, ntheta, 1, pm, kinds, solution_d=(1, 1, 0))
    split_patch(m, "outer", ntheta // 2, "outlet", "inlet")
    # exact axis: the theta = 0 and theta = pi point rows must sit on y = 0 (sin(pi) is 1.2e-16, not 0)
    c = np.arange(nr * ntheta)
    ci, cj = c % nr, c // nr
    def corner(i, j):
        p = m.points[i + (nr + 1) * j]  # k = 0 layer
        return p[:, :2]
    m.cell_quads = np.stack([corner(ci, cj), corner(ci + 1, cj), corner(ci + 1, cj + 1), corner(ci, cj + 1)], axis=1)
    return m


def box_mesh(nx, ny, nz, lx, ly, lz, patch_kinds=None, cyclic_pairs=(), solution_d=(1, 1, 1), origin=(0.0, 0.0, 0.0)):
    """Uniform Cartesian block (blockMesh with one hex and simpleGrading 1)."""
    if patch_kinds is None:
        patch_kinds = {s: (s, "wall") for s in ("xMin", "xMax", "yMin", "yMax", "zMin", "zMax")}

    def pm(I, J, K):
        return origin[0] + lx * I / nx, origin[1] + ly * J / ny, origin[2] + lz * K / nz

    return structured_block(nx, ny, nz, pm, patch_kinds, cyclic_pairs, solution_d)


def slab_partition(mesh, n_ranks, axis=0):
    """Block partition of a structured mesh into n_ranks slabs along `axis` (decomposePar simple)."""
    nx, ny, nz = mesh.shape
    c = np.arange(nx * ny * nz)
    ijk = (c % nx, (c // nx) % ny, c // (nx * ny))
    n = mesh.shape[axis]
    return np.minimum((ijk[axis] * n_ranks) // n, n_ranks - 1).astype(np.int32)


def block_partition(mesh, parts):
    """decomposePar `method simple` with `n (px py pz)`: px x py x pz blocks of a structured mesh, rank = bx + px (by + py bz)."""
    nx, ny, nz = mesh.shape
    px, py, pz = (int(v) for v in parts)
    if px < 1 or py < 1 or pz < 1 or px > nx or py > ny or pz > nz:
        raise ValueError(f"cannot cut a {mesh.shape} mesh into {parts} blocks")
    c = np.arange(nx * ny * nz)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    b = lambda idx, n, p: np.minimum((idx * p) // n, p - 1)
    return (b(i, nx, px) + px * (b(j, ny, py) + py * b(k, nz, pz))).astype(np.int32)


def weighted_slab_partition(mesh, n_ranks, weights, axis=0):
    """Slabs along `axis` of a structured mesh cut at equal cumulative *weight* instead of equal cell count - the
    decomposeParDict `weightField uniGasRhoNMean_<species>` the tutorials hint at (hypersonicCylinder/system/
    decomposeParDict): with the time-averaged parcel count per cell as weight every rank gets the same number of parcels
    (the 1 -> 8 target needs the partitioner to balance parcels, not cells).  Every rank keeps at least one
    layer of cells."""
    nx, ny, nz = mesh.shape
    c = np.arange(nx * ny * nz)
    ijk = (c % nx, (c // nx) % ny, c // (nx * ny))
    n = mesh.shape[axis]
    if n < n_ranks:
        raise ValueError(f"{n} layers along axis {axis} cannot feed {n_ranks} ranks")
    w = np.asarray(weights, float)
    if w.shape != (mesh.n_cells,) or (w < 0).any():
        raise ValueError("weights: one non-negative value per cell")
    layer = np.bincount(ijk[axis], weights=w, minlength=n)
    if layer.sum() <= 0:
        return slab_partition(mesh, n_ranks, axis)
    cum = np.cumsum(layer)
    # layer i goes to the rank whose share of the total its mid-point falls into, then every rank is made non-empty
    r = np.minimum(((cum - 0.5 * layer) / cum[-1] * n_ranks).astype(int), n_ranks - 1)
    r = np.maximum.accumulate(r)
    for k in range(n):                      # no rank skipped from the left ...
        r[k] = min(r[k], (r[k - 1] + 1) if k else 0)
    for k in range(n - 1, -1, -1):          # ... and none left empty at the right end
        r[k] = max(r[k], n_ranks - (n - k))
    return r[ijk[axis]].astype(np.int32)


def load_imbalance(parcels_per_rank):
    """uniGasDynamicLoadBalancing::calculate (U/dynamicLoadBalancing/uniGasDynamicLoadBalancing.C:48-68): the number the
    reference prints as `Maximum imbalance` - max |N_rank - N/nRanks| / (N/nRanks) in per cent."""
    n = np.asarray(parcels_per_rank, float)
    ideal = n.sum() / len(n)
    return 100.0 * np.abs(n - ideal).max() / ideal if ideal > 0 else 0.0


def decompose(mesh, cell_rank, n_ranks):
    """Split `mesh` into n_ranks sub-meshes with processor patches (decomposePar stand-in).

    Returns a list of PolyMesh; each has `cell_map` (local -> global cell) and processor
    patches whose `partner` is the peer rank and `peer_patch` the matching patch index there.
    Faces of matching processor patches are in the same order on both sides.
    """
    nI = mesh.n_internal
    own, nei = mesh.owner, mesh.neighbour
    off = mesh.face_point_offsets
    subs = []
    for r in range(n_ranks):
        local_cells = np.nonzero(cell_rank == r)[0]
        g2l = np.full(mesh.n_cells, -1, np.int64)
        g2l[local_cells] = np.arange(len(local_cells))
        ro, rn = cell_rank[own[:nI]], cell_rank[nei]
        # faces: (global face, flip) lists per local patch
        f_int = np.nonzero((ro == r) & (rn == r))[0]
        face_ids = [f_int]
        flips = [np.zeros(len(f_int), bool)]
        l_own = [g2l[own[f_int]]]
        l_nei = g2l[nei[f_int]]
        patches = []
        start = len(f_int)
        proc_groups = {}  # (peer, tag) -> (faces, flip, sep)
        for pi, p in enumerate(mesh.patches):
            f = np.arange(p.start, p.start + p.size)
            mine = cell_rank[own[f]] == r
            if p.kind == "cyclic":
                q = mesh.patches[p.partner]
                peer = cell_rank[own[q.start : q.start + q.size]]
                stay = mine & (peer == r)
                cross = mine & (peer != r)
                for b in np.unique(peer[cross]):
                    s = cross & (peer == b)
                    proc_groups[(int(b), ("c", pi, p.partner))] = (f[s], np.zeros(s.sum(), bool), p.separation)
                mine = stay
            fl = f[mine]
            patches.append(Patch(p.name, p.kind, start, len(fl), p.partner, p.separation))
            face_ids.append(fl); flips.append(np.zeros(len(fl), bool)); l_own.append(g2l[own[fl]])
            start += len(fl)
        cutO = np.nonzero((ro == r) & (rn != r))[0]
        cutN = np.nonzero((ro != r) & (rn == r))[0]
        for b in np.unique(np.concatenate([rn[cutO], ro[cutN]])):
            fo = cutO[rn[cutO] == b]
            fn = cutN[ro[cutN] == b]
            f = np.concatenate([fo, fn])
            fl = np.concatenate([np.zeros(len(fo), bool), np.ones(len(fn), bool)])
            o = np.argsort(f, kind="stable")
            proc_groups[(int(b), ("i",))] = (f[o], fl[o], (0.0, 0.0, 0.0))
        for (peer, tag) in sorted(proc_groups, key=lambda k: (k[0], k[1])):
            f, fl, sep = proc_groups[(peer, tag)]
            patches.append(Patch(f"procBoundary{r}to{peer}" + ("" if tag[0] == "i" else f"through{mesh.patches[tag[1]].name}"),
                                 "processor", start, len(f), peer, tuple(sep), tag=tag))
            face_ids.append(f); flips.append(fl)
            l_own.append(np.where(fl, g2l[nei[np.minimum(f, nI - 1)]] if nI else -1, g2l[own[f]]))
            start += len(f)
        gf = np.concatenate(face_ids)
        flip = np.concatenate(flips)
        lo = np.concatenate(l_own).astype(np.int32)
        # face vertex lists (reverse flipped faces so that the normal follows the new owner)
        nv = np.diff(off)[gf]
        assert (nv == nv[0]).all(), "decompose assumes a uniform vertex count per face"
        k = int(nv[0])
        verts = mesh.face_points[off[gf][:, None] + np.arange(k)[None, :]]
        verts = np.where(flip[:, None], verts[:, ::-1], verts)
        sub = PolyMesh(
            points=mesh.points,
            face_point_offsets=(k * np.arange(len(gf) + 1)).astype(np.int32),
            face_points=verts.ravel().astype(np.int32),
            owner=lo,
            neighbour=l_nei.astype(np.int32),
            patches=patches,
            solution_d=mesh.solution_d,
            cell_map=local_cells,
        )
        # geometry is inherited, not recomputed, so that every rank sees bit-identical planes
        sgn = np.where(flip, -1.0, 1.0)[:, None]
        sub.face_areas = mesh.face_areas[gf] * sgn
        sub.face_centres = mesh.face_centres[gf]
        nC = len(local_cells)
        sub.cell_volumes = mesh.cell_volumes[local_cells]
        sub.cell_centres = mesh.cell_centres[local_cells]
        sub.cell_bb_min = mesh.cell_bb_min[local_cells]
        sub.cell_bb_max = mesh.cell_bb_max[local_cells]
        cells = np.concatenate([sub.owner, sub.neighbour])
        faces = np.concatenate([np.arange(len(sub.owner)), np.arange(len(sub.neighbour))])
        order = np.lexsort((faces, cells))
        sub.cell_faces = faces[order].astype(np.int32)
        sub.cell_face_offsets = np.concatenate([[0], np.cumsum(np.bincount(cells, minlength=nC))]).astype(np.int32)
        subs.append(sub)
    # match processor patches across ranks
    for r, sub in enumerate(subs):
        for p in sub.patches:
            if p.kind != "processor":
                continue
            want = ("i",) if p.tag[0] == "i" else ("c", p.tag[2], p.tag[1])
            peer = subs[p.partner]
            for qi, q in enumerate(peer.patches):
                if q.kind == "processor" and q.partner == r and q.tag == want:
                    p.peer_patch = qi
                    assert q.size == p.size
                    break
            assert p.peer_patch >= 0
    return subs


# ---- per-rank blocks of a structured mesh, built directly (no global mesh in memory) ---------------------------------

def block_ranges(n, parts):
    """The index ranges `method simple` gives `parts` blocks along an axis of n cells (same rule as block_partition)."""
    idx = np.arange(n)
    b = np.minimum((idx * parts) // n, parts - 1)
    return [(int(np.nonzero(b == k)[0][0]), int(np.nonzero(b == k)[0][-1]) + 1) for k in range(parts)]


def hex_corners(mesh):
    """[nCells, 8] point labels of a structured block's cells: corner (di, dj, dk) at column di + 2 dj + 4 dk."""
    nx, ny, nz = mesh.shape
    c = np.arange(nx * ny * nz)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    cols = []
    for dk in (0, 1):
        for dj in (0, 1):
            for di in (0, 1):
                cols.append((i + di) + (nx + 1) * ((j + dj) + (ny + 1) * (k + dk)))
    return np.stack(cols, axis=1).astype(np.int32)


def structured_subblock(global_shape, parts, rank, point_map, patch_kinds, solution_d=(1, 1, 1)):
    """Block `rank` (= bx + px (by + py bz), the numbering of block_partition / decomposePar `simple`) of a structured mesh
    of global_shape cells cut into parts = (px, py, pz) blocks, built on its own: what decomposePar would hand this rank,
    without ever holding the global mesh.  point_map takes *global* vertex indices; patch_kinds names the sides of the
    global block; sides that face another block become processor patches (`procBoundary<r>to<peer>`, same face order on
    both sides).  -> (mesh, (i0, j0, k0)) with the block's first global cell index."""
    px, py, pz = (int(v) for v in parts)
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
    rx, ry, rz = block_ranges(global_shape[0], px)[bx], block_ranges(global_shape[1], py)[by], block_ranges(global_shape[2], pz)[bz]
    i0, j0, k0 = rx[0], ry[0], rz[0]

    def pm(I, J, K):
        return point_map(I + i0, J + j0, K + k0)

    kinds, peers = {}, {}
    for side, (b, p, step) in {"xMin": (bx, px, -1), "xMax": (bx, px, 1), "yMin": (by, py, -px), "yMax": (by, py, px),
                               "zMin": (bz, pz, -px * py), "zMax": (bz, pz, px * py)}.items():
        at_edge = (b == 0) if side.endswith("Min") else (b == p - 1)
        if at_edge:
            kinds[side] = patch_kinds[side]
        else:
            peer = rank + step
            kinds[side] = (f"procBoundary{rank}to{peer}", "processor")
            peers[side] = peer
    m = structured_block(rx[1] - rx[0], ry[1] - ry[0], rz[1] - rz[0], pm, kinds, solution_d=solution_d, snap=True)
    for pi, side in enumerate(("xMin", "xMax", "yMin", "yMax", "zMin", "zMax")):
        if side in peers:
            q = m.patches[pi]
            q.partner, q.tag = peers[side], ("i",)
    have = {q.name for q in m.patches}
    for side in ("xMin", "xMax", "yMin", "yMax", "zMin", "zMax"):  # decomposePar keeps every patch on every rank, empty where
        name, kind = patch_kinds[side]                             # the rank does not touch it
        if name not in have and kind != "cyclic":
            m.patches.append(Patch(name, kind, m.n_faces, 0))
            have.add(name)
    return m, (i0, j0, k0)


def sphere_cone_map(n_eta, n_s, n_phi, nose_radius=0.05, cone_half_angle_deg=30.0, cone_length=0.2, standoff=(0.02, 0.10),
                    phi_max=0.5 * np.pi, grading=4.0):
    """Point map of a blunted cone (sphere-cone fore-body) in a body-fitted grid: i runs along the body normal from the wall
    (last / first cell size = grading) to the outer boundary, j along the body from the stagnation line, k in azimuth about the
    x axis.  Every face is planar (isosceles trapezoids or meridian-plane quads) and the cells touching the axis (j = 0) are
    wedges whose j = 0 side collapses onto the axis."""
    a = np.deg2rad(cone_half_angle_deg)
    th_t = 0.5 * np.pi - a                      # polar angle of the sphere / cone tangency point
    s_sph = nose_radius * th_t
    S = s_sph + cone_length
    g = float(grading)

    def pm(I, J, K):
        s = S * J / n_s
        on_sphere = s <= s_sph
        th = np.where(on_sphere, s / nose_radius, th_t)
        along = np.where(on_sphere, 0.0, s - s_sph)
        bx = nose_radius * (1.0 - np.cos(th)) + along * np.cos(a)
        br = nose_radius * np.sin(th) + along * np.sin(a)
        nxm, nr = -np.cos(th), np.sin(th)       # outward normal in the meridian plane
        t = I / n_eta
        frac = t if abs(g - 1.0) < 1e-12 else (g ** t - 1.0) / (g - 1.0)
        d = (standoff[0] + (standoff[1] - standoff[0]) * (s / S)) * frac
        x = bx + d * nxm
        r = np.where(J == 0, 0.0, br + d * nr)  # the stagnation line is the axis exactly
        phi = -phi_max * K / n_phi              # (normal, along the body, azimuth) right-handed: revolve towards -z
        return x, r * np.cos(phi), r * np.sin(phi)

    return pm


def wedge_map(nx, nr, length, radius, half_angle_deg=0.5, r_inner=0.0, gx=1.0, gr=1.0):
    """Point map of an axisymmetric wedge about the x axis, one cell thick in azimuth between the planes phi = -a and phi = +a
    (the tutorials' blockMesh wedges, e.g. plumeImpingement/system/blockMeshDict: half angle 0.5 degrees, side patches typed
    symmetryPlane as uniGasBoundaries.C:438-445 demands): i along x, j along r, k = 0 / 1 the two side planes.  With
    r_inner = 0 the j = 0 faces collapse onto the axis (zero area; patch type symmetry as in the tutorials)."""
    a = np.deg2rad(half_angle_deg)

    def frac(t, g):
        return t if abs(g - 1.0) < 1e-12 else (g ** t - 1.0) / (g - 1.0)

    def pm(I, J, K):
        x = length * frac(I / nx, gx)
        r = np.where((J == 0) & (r_inner == 0.0), 0.0, r_inner + (radius - r_inner) * frac(J / nr, gr))
        phi = -a + 2.0 * a * K
        return x, r * np.cos(phi), r * np.sin(phi)

    return pm


# ---- interpolationCellPoint support (collisionProperties.macroInterpolation) ------------------------------------------

def cell_point_data(mesh):
    """What OpenFOAM's interpolationCellPoint needs, restated from its definition (volPointInterpolation + cellPointWeight;
    OpenFOAM is not in /root/reference, call sites: U/bgkCollisions/derived/*/…C `interpolationCellPoint<T>::New("cellPoint", psi)`):

    * point values = inverse-distance weighted mean of the values of the cells around the point (weight 1 / |p - C_c|); on the
      points of boundary faces of non-empty, non-coupled patches instead the inverse-distance weighted mean (1 / |p - C_f|) of the
      values on those faces, which for the zeroGradient fields of the BGK models are the owner-cell values; points of a cyclic
      pair see the cells on both sides; at points of symmetry patches vectors / tensors lose their normal components;
    * a cell is split into tets (cell centre, face point 0, face point k, face point k + 1) over its faces in cell-face order;
      the value at a position is linear in the tet that contains it.

    -> dict(tetOffsets [nCells+1], tetPoints [nTets,3], pointCellOffsets [nPoints+1], pointCells, pointWeights (normalised),
            pointNormals [nPoints,3]).  Processor patches are not handled (the point sums would need a halo)."""
    if any(p.kind == "processor" and p.size for p in mesh.patches):
        raise NotImplementedError("macroInterpolation on a decomposed mesh: point values across processor patches need a halo")
    nP, nC, nF, nI = len(mesh.points), mesh.n_cells, mesh.n_faces, mesh.n_internal
    off, fp = mesh.face_point_offsets, mesh.face_points
    nv = np.diff(off)
    # tets: per cell, over its faces in cell-face order, fan about the face's first point
    cf_cell = np.repeat(np.arange(nC), np.diff(mesh.cell_face_offsets))
    cf_face = mesh.cell_faces
    ntri = nv[cf_face] - 2
    tet_cell = np.repeat(cf_cell, ntri)
    tet_face = np.repeat(cf_face, ntri)
    k = np.arange(len(tet_face)) - np.repeat(np.concatenate([[0], np.cumsum(ntri)[:-1]]), ntri) + 1
    base = fp[off[tet_face]]
    pa = fp[off[tet_face] + k]
    pb = fp[off[tet_face] + k + 1]
    tet_points = np.stack([base, pa, pb], axis=1).astype(np.int32)
    tet_offsets = np.concatenate([[0], np.cumsum(np.bincount(tet_cell, minlength=nC))]).astype(np.int32)
    # point classes across cyclic pairs (translational): representative = smallest label
    rep = np.arange(nP)
    for pi, p in enumerate(mesh.patches):
        if p.kind != "cyclic" or p.partner < pi:
            continue
        q = mesh.patches[p.partner]
        sep = np.asarray(p.separation, float)
        pa_pts = np.unique(np.concatenate([fp[off[f]:off[f + 1]] for f in range(p.start, p.start + p.size)]))
        pb_pts = np.unique(np.concatenate([fp[off[f]:off[f + 1]] for f in range(q.start, q.start + q.size)]))
        scale = np.abs(mesh.points).max() + 1e-300
        key = lambda x: tuple(np.round(x / (1e-9 * scale)).astype(np.int64))
        lut = {key(mesh.points[b]): b for b in pb_pts}
        for a in pa_pts:
            b = lut.get(key(mesh.points[a] + sep))
            if b is not None:
                ra, rb = rep[a], rep[b]
                lo, hi = min(ra, rb), max(ra, rb)
                rep[rep == hi] = lo
    # boundary points: points of faces of non-empty, non-coupled patches
    bface = []
    sym_normals = np.zeros((nP, 3))
    for p in mesh.patches:
        if p.kind in ("empty", "cyclic", "processor") or p.size == 0:
            continue
        f = np.arange(p.start, p.start + p.size)
        A = np.linalg.norm(mesh.face_areas[f], axis=1)
        f = f[A > 0]  # collapsed faces carry nothing
        bface.append(f)
        if p.kind in ("symmetry", "symmetryPlane"):
            for ff in f:
                n = mesh.face_areas[ff] / np.linalg.norm(mesh.face_areas[ff])
                sym_normals[fp[off[ff]:off[ff + 1]]] = n
    bface = np.concatenate(bface) if bface else np.zeros(0, np.int64)
    is_bpoint = np.zeros(nP, bool)
    rows_p, rows_c, rows_w = [], [], []
    if len(bface):
        bf_pts = np.concatenate([fp[off[f]:off[f + 1]] for f in bface])
        bf_face = np.repeat(bface, nv[bface])
        is_bpoint[bf_pts] = True
        d = np.linalg.norm(mesh.points[bf_pts] - mesh.face_centres[bf_face], axis=1)
        rows_p.append(bf_pts); rows_c.append(mesh.owner[bf_face]); rows_w.append(1.0 / d)
    # internal points: cells around the point (through the faces), each cell once
    f_pts = fp
    f_face = np.repeat(np.arange(nF), nv)
    pc = np.concatenate([np.stack([f_pts, mesh.owner[f_face]], 1),
                         np.stack([f_pts[f_face < nI], mesh.neighbour[f_face[f_face < nI]]], 1)])
    pc = np.unique(pc, axis=0)
    pc = pc[~is_bpoint[pc[:, 0]]]
    d = np.linalg.norm(mesh.points[pc[:, 0]] - mesh.cell_centres[pc[:, 1]], axis=1)
    rows_p.append(pc[:, 0]); rows_c.append(pc[:, 1]); rows_w.append(1.0 / d)
    P_ = np.concatenate(rows_p); C_ = np.concatenate(rows_c); W_ = np.concatenate(rows_w)
    # merge cyclic classes: every member gets the union of the members' contributions (coupled points are never boundary
    # points in OpenFOAM; a class with a boundary member keeps that member's boundary rule for itself)
    cls = rep[P_]
    if (rep != np.arange(nP)).any():
        members = {}
        for pt in np.nonzero(rep != np.arange(nP))[0]:
            members.setdefault(rep[pt], [rep[pt]]).append(pt)
        extraP, extraC, extraW = [], [], []
        for r, mem in members.items():
            sel = np.isin(P_, mem) & ~is_bpoint[P_]
            for pt in mem:
                if is_bpoint[pt]:
                    continue
                other = sel & (P_ != pt)
                extraP.append(np.full(other.sum(), pt)); extraC.append(C_[other]); extraW.append(W_[other])
        if extraP:
            P_ = np.concatenate([P_] + extraP); C_ = np.concatenate([C_] + extraC); W_ = np.concatenate([W_] + extraW)
    order = np.lexsort((C_, P_))
    P_, C_, W_ = P_[order], C_[order], W_[order]
    wsum = np.bincount(P_, weights=W_, minlength=nP)
    W_ = W_ / wsum[P_]
    pc_off = np.concatenate([[0], np.cumsum(np.bincount(P_, minlength=nP))]).astype(np.int32)
    return dict(tetOffsets=tet_offsets, tetPoints=tet_points, pointCellOffsets=pc_off, pointCells=C_.astype(np.int32),
                pointWeights=np.ascontiguousarray(W_), pointNormals=np.ascontiguousarray(sym_normals))
