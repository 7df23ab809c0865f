"""BASELINE.json's configurations as concrete synthetic cases.

Each builder returns a `Case`: the mesh, the uniGasProperties / boundariesDict
dictionaries (reference key names), deltaT, the initial parcels (uniGasMeshFill
semantics, U/uniGasInitialisation/derived/uniGasMeshFill/uniGasMeshFill.C:174-296) and
the initial cell state.  Host-side numpy only: initialisation runs once and is outside
the per-step hot path.
"""
from dataclasses import dataclass, field
import math
import os

import numpy as np

from . import mesh as _mesh

kB = 1.38065e-23  # OpenFOAM physicoChemical::k default

# moleculeProperties presets
ARGON_GUIDE = dict(mass=66.3e-27, diameter=4.17e-10, omega=0.81, alpha=1.0, rotationalDegreesOfFreedom=0,
                   vibrationalModes=0, charge=0, numberOfElectronicLevels=1, electronicEnergyList=[0.0], degeneracyList=[1])
# tutorials/uniGasFoam/hypersonicCylinder/constant/uniGasProperties:88-106
ARGON_TUTORIAL = dict(ARGON_GUIDE, diameter=3.595e-10, omega=0.734)
# Bird (1994) Appendix A; not in the reference repo
NITROGEN = dict(mass=46.5e-27, diameter=4.17e-10, omega=0.74, alpha=1.0, rotationalDegreesOfFreedom=2,
                vibrationalModes=0, charge=0, numberOfElectronicLevels=1, electronicEnergyList=[0.0], degeneracyList=[1])


def vhs_mean_free_path(n, T, sp, Tref):
    """Bird eq 4.65."""
    return 1.0 / (math.sqrt(2.0) * math.pi * sp["diameter"] ** 2 * n * (Tref / T) ** (sp["omega"] - 0.5))


def vhs_collision_rate(n, T, sp, Tref):
    """Equilibrium collision rate per molecule, Bird eq 4.64 (coded at uniGasVolFields.C:1150-1151)."""
    return 4.0 * sp["diameter"] ** 2 * n * math.sqrt(math.pi * kB * Tref / sp["mass"]) * (T / Tref) ** (1.0 - sp["omega"])


def most_probable_speed(T, m):
    return math.sqrt(2.0 * kB * T / m)


@dataclass
class Case:
    name: str
    mesh: object
    uniGasProperties: dict
    boundariesDict: dict
    deltaT: float
    position: np.ndarray
    U: np.ndarray
    cell: np.ndarray
    typeId: np.ndarray = None
    ERot: np.ndarray = None
    sigmaTcRMax: float = 0.0
    cellCollModelId: np.ndarray = None
    subCellLevels: np.ndarray = None
    cellWeightFactor: np.ndarray = None   # uniGasCellWeightFactor (cellWeightedSimulation true)
    meta: dict = field(default_factory=dict)
    vibLevel: np.ndarray = None           # [n, nModes] vibrational quantum levels
    ELevel: np.ndarray = None             # [n] electronic levels

    @property
    def n_parcels(self):
        return len(self.cell)

    def make_cloud(self, cloud_cls=None, **kw):
        if cloud_cls is None:
            from .cloud import UniGasCloud as cloud_cls
        if self.cellWeightFactor is not None:  # handle must exist before the parcels: fix the capacity now (clones need room)
            kw.setdefault("parcelCapacity", int(1.5 * self.n_parcels) + 4096)
        cl = cloud_cls(self.mesh, self.uniGasProperties, self.boundariesDict, self.deltaT, **kw)
        if self.cellWeightFactor is not None:  # the parcels take the factor of their cell: the field goes in first
            cl.setCellState(cellWeightFactor=self.cellWeightFactor)
        cl.setParcels(self.position, self.U, self.cell, self.typeId, self.ERot, vibLevel=self.vibLevel, ELevel=self.ELevel)
        cl.setCellState(sigmaTcRMax=self.sigmaTcRMax, cellCollModelId=self.cellCollModelId, subCellLevels=self.subCellLevels)
        return cl


def _uniform_in_cells(mesh, cells, rng):
    """Uniform points in hex cells.  Axis-aligned boxes: uniform in the bounding box.  Extruded
    2-D cells: triangle split of the zMin quad, uniform height (tet.randomPoint analogue)."""
    lo, hi = mesh.cell_bb_min[cells], mesh.cell_bb_max[cells]
    if mesh.meta_axis_aligned:
        return lo + rng.random((len(cells), 3)) * (hi - lo)
    if mesh.cell_quads is None:
        # general hex cells of a structured block (wedges, bodies of revolution): six tets around the 0-7 diagonal, one chosen by
        # volume, uniform point in it (tetPointRef::randomPoint, as uniGasMeshFill does per tet, uniGasMeshFill.C:174-278)
        corners = mesh.points[_mesh.hex_corners(mesh)[cells]]            # [n,8,3]
        tets = np.asarray(_HEX_TETS)
        a = corners[:, tets[:, 0]]
        e = np.stack([corners[:, tets[:, 1]] - a, corners[:, tets[:, 2]] - a, corners[:, tets[:, 3]] - a], axis=-2)  # [n,6,3,3]
        cum = np.cumsum(np.abs(np.linalg.det(e)), axis=1)
        pick = (rng.random(len(cells))[:, None] * cum[:, -1:] >= cum).sum(1).clip(0, 5)
        idx = np.arange(len(cells))
        a, e = a[idx, pick], e[idx, pick]
        s, t, u = rng.random(len(cells)), rng.random(len(cells)), rng.random(len(cells))
        # fold the unit cube into the unit tetrahedron (Rocchini & Cignoni)
        f = s + t > 1.0
        s, t = np.where(f, 1.0 - s, s), np.where(f, 1.0 - t, t)
        g1 = t + u > 1.0
        t2 = np.where(g1, 1.0 - u, t); u2 = np.where(g1, 1.0 - s - t, u)
        g2 = ~g1 & (s + t + u > 1.0)
        s2 = np.where(g2, 1.0 - t - u, s); u2 = np.where(g2, s + t + u - 1.0, u2)
        s, t, u = s2, t2, u2
        return a + s[:, None] * e[:, 0] + t[:, None] * e[:, 1] + u[:, None] * e[:, 2]
    # general extruded quads: use the cell's zMin face (patch order guarantees it exists for nz = 1)
    quad = mesh.cell_quads[cells]  # [n,4,2] xy of the 4 corners, counter-clockwise
    a, b, c, d = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    cross2 = lambda p, q: p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0]
    A1 = 0.5 * np.abs(cross2(b - a, c - a))
    A2 = 0.5 * np.abs(cross2(c - a, d - a))
    pick = rng.random(len(cells)) * (A1 + A2) < A1
    s, t = rng.random(len(cells)), rng.random(len(cells))
    fl = s + t > 1
    s = np.where(fl, 1 - s, s); t = np.where(fl, 1 - t, t)
    p1 = a + s[:, None] * (b - a) + t[:, None] * (c - a)
    p2 = a + s[:, None] * (c - a) + t[:, None] * (d - a)
    xy = np.where(pick[:, None], p1, p2)
    z = lo[:, 2] + rng.random(len(cells)) * (hi[:, 2] - lo[:, 2])
    return np.column_stack([xy, z])


def mesh_fill(mesh, species, names, number_densities, T, velocity, nParticle, rng, Trot=None, cells=None, cell_weight=None):
    """uniGasMeshFill::setInitialConfiguration (…/uniGasMeshFill.C:174-278), per cell instead of per tet:
    N = n V / (F_N CWF) with stochastic rounding (:196-199), uniform position, Maxwellian + drift, equipartition ERot.
    number_densities[name], T, Trot may be scalars or per-cell arrays and velocity a vector or [nCells,3] - the latter is
    uniGasMeshFieldFill (…/uniGasMeshFieldFill/uniGasMeshFieldFill.C:60-330: every cell filled at its own state)."""
    pos, vel, cel, tid, erot = [], [], [], [], []
    all_cells = np.arange(mesh.n_cells) if cells is None else np.asarray(cells)
    for ti, name in enumerate(names):
        sp = species[name]
        nd = np.broadcast_to(np.asarray(number_densities[name], float), (mesh.n_cells,))
        req = nd[all_cells] / nParticle * mesh.cell_volumes[all_cells]
        if cell_weight is not None:
            req = req / np.asarray(cell_weight, float)[all_cells]
        cnt = np.floor(req).astype(np.int64)
        cnt += (req - cnt) > rng.random(len(all_cells))
        cl = np.repeat(all_cells, cnt)
        n = len(cl)
        pos.append(_uniform_in_cells(mesh, cl, rng))
        Tc = np.broadcast_to(np.asarray(T, float), (mesh.n_cells,))[cl]
        Uc = np.broadcast_to(np.asarray(velocity, float), (mesh.n_cells, 3))[cl]
        vel.append(np.sqrt(kB * Tc / sp["mass"])[:, None] * rng.standard_normal((n, 3)) + Uc)
        cel.append(cl); tid.append(np.full(n, ti, np.int32))
        rd = sp.get("rotationalDegreesOfFreedom", 0)
        tr = np.broadcast_to(np.asarray(T if Trot is None else Trot, float), (mesh.n_cells,))[cl]
        if rd == 0:
            erot.append(np.zeros(n))
        elif rd == 2:
            erot.append(-np.log(1.0 - rng.random(n)) * kB * tr)
        else:
            erot.append(rng.gamma(0.5 * rd, 1.0, n) * kB * tr)
    pos = np.concatenate(pos); vel = np.concatenate(vel); cel = np.concatenate(cel)
    tid = np.concatenate(tid); erot = np.concatenate(erot)
    # the cloud is a linked list in insertion order: cell-major here
    order = np.argsort(cel, kind="stable")
    pos, vel, cel, tid, erot = pos[order], vel[order], cel[order], tid[order], erot[order]
    for d in range(3):  # empty directions: particles sit on the mesh mid-plane (deviationFromMeshCentre)
        if not mesh.solution_d[d]:
            pos[:, d] = 0.5 * (mesh.points[:, d].min() + mesh.points[:, d].max())
    return pos, vel, cel.astype(np.int32), tid, erot


def cell_weight_factor(mesh, spec, number_density, nParticle):
    """cellWeightFactor field of a case.  spec: None | array [nCells] | callable(mesh) -> array |
    ("particlesPerSubCell", k): uniGasMeshFill's rule CWF = n V / (k nSubCells F_N) with one sub-cell per cell
    and without its smoothing passes (U/uniGasInitialisation/derived/uniGasMeshFill/uniGasMeshFill.C:111-121)."""
    if spec is None:
        return None
    if callable(spec):
        w = spec(mesh)
    elif isinstance(spec, tuple) and spec[0] == "particlesPerSubCell":
        w = number_density * mesh.cell_volumes / (float(spec[1]) * nParticle)
    else:
        w = np.broadcast_to(np.asarray(spec, float), (mesh.n_cells,))
    return np.ascontiguousarray(w, dtype=np.float64)


def _weighted(case, w):
    if w is not None:
        case.cellWeightFactor = w
        case.uniGasProperties["cellWeightedSimulation"] = True
    return case


def _props(species_name, sp, nParticle, mode="dsmc", binary="variableHardSphere", bgk="noBGKCollision", Tref=273.0, **cp):
    return {
        "nEquivalentParticles": nParticle,
        "chemicalReactions": False, "cellWeightedSimulation": False, "axisymmetricSimulation": False,
        "adaptiveSimulation": False,
        "collisionModel": mode, "bgkCollisionModel": bgk,
        "dsmcCollisionPartnerModel": "noTimeCounter", "dsmcCollisionModel": binary,
        "collisionProperties": dict(Tref=Tref, **cp),
        "typeIdList": [species_name], "moleculeProperties": {species_name: sp},
    }


def closed_box(n=32, parcels=1_000_000, wall="specular", T0=300.0, number_density=1e20, species=("Ar", ARGON_GUIDE),
               Tref=273.0, binary="variableHardSphere", mode="dsmc", bgk="noBGKCollision", seed=1, dt_mct=0.2,
               lambda_per_dx=2.0, velocity=(0.0, 0.0, 0.0), Trot=None, cellWeightFactor=None, **cp):
    """Config 1: 3-D closed box of gas at equilibrium, n^3 cells, dx = lambda/2, dt = 0.2 MCT."""
    name, sp = species
    lam = vhs_mean_free_path(number_density, T0, sp, Tref)
    dx = lam / lambda_per_dx
    L = n * dx
    m = _mesh.box_mesh(n, n, n, L, L, L)
    m.meta_axis_aligned = True
    nParticle = number_density * L ** 3 / parcels
    rng = np.random.default_rng(seed)
    cwf = cell_weight_factor(m, cellWeightFactor, number_density, nParticle)
    pos, vel, cel, tid, erot = mesh_fill(m, {name: sp}, [name], {name: number_density}, T0, velocity, nParticle, rng, Trot=Trot, cell_weight=cwf)
    dt = dt_mct / vhs_collision_rate(number_density, T0, sp, Tref)
    if wall == "specular":
        model = lambda p: {"patchBoundaryProperties": {"patch": p}, "boundaryModel": "uniGasSpecularWallPatch"}
    else:
        model = lambda p: {"patchBoundaryProperties": {"patch": p}, "boundaryModel": "uniGasDiffuseWallPatch",
                           "uniGasDiffuseWallPatchProperties": {"temperature": T0, "velocity": [0, 0, 0]}}
    bd = {"uniGasPatchBoundaries": [model(p.name) for p in m.patches]}
    sig0 = math.pi * sp["diameter"] ** 2 * most_probable_speed(T0, sp["mass"])  # uniGasMeshFill.C:284-296
    return _weighted(Case("closed_box", m, _props(name, sp, nParticle, mode, binary, bgk, Tref, **cp), bd, dt, pos, vel, cel, tid,
                          erot if sp.get("rotationalDegreesOfFreedom", 0) else None, sig0,
                          meta=dict(n=number_density, T0=T0, lam=lam, L=L, Tref=Tref, species=sp)), cwf)


def mixture_box(n=6, parcels=20000, fractions=(("Ar", None, 0.6), ("N2", None, 0.4)), T0=300.0, number_density=1e20, Tref=273.0,
                binary="LarsenBorgnakkeVariableHardSphere", mode="dsmc", bgk="noBGKCollision", wall="diffuse", seed=5, dt_mct=0.5,
                lambda_per_dx=1.0, Trot=None, cellWeightFactor=None, **cp):
    """Closed 3-D box holding a gas mixture (typeIdList with several species: the multi-species code paths -
    typeId per parcel, per-species cell sums, cross-species collision pairs)."""
    table = {"Ar": ARGON_GUIDE, "N2": NITROGEN}
    names = [f[0] for f in fractions]
    sps = {f[0]: (f[1] or table[f[0]]) for f in fractions}
    dens = {f[0]: number_density * f[2] for f in fractions}
    sp0 = sps[names[0]]
    lam = vhs_mean_free_path(number_density, T0, sp0, Tref)
    L = n * lam / lambda_per_dx
    m = _mesh.box_mesh(n, n, n, L, L, L)
    m.meta_axis_aligned = True
    nParticle = number_density * L ** 3 / parcels
    rng = np.random.default_rng(seed)
    cwf = cell_weight_factor(m, cellWeightFactor, number_density, nParticle)
    pos, vel, cel, tid, erot = mesh_fill(m, sps, names, dens, T0, (0.0, 0.0, 0.0), nParticle, rng, Trot=Trot, cell_weight=cwf)
    dt = dt_mct / vhs_collision_rate(number_density, T0, sp0, Tref)
    if wall == "specular":
        model = lambda p: {"patchBoundaryProperties": {"patch": p}, "boundaryModel": "uniGasSpecularWallPatch"}
    else:
        model = lambda p: {"patchBoundaryProperties": {"patch": p}, "boundaryModel": "uniGasDiffuseWallPatch",
                           "uniGasDiffuseWallPatchProperties": {"temperature": T0, "velocity": [0, 0, 0]}}
    bd = {"uniGasPatchBoundaries": [model(p.name) for p in m.patches]}
    props = _props(names[0], sp0, nParticle, mode, binary, bgk, Tref, **cp)
    props["typeIdList"] = names
    props["moleculeProperties"] = sps
    sig0 = math.pi * sp0["diameter"] ** 2 * most_probable_speed(T0, sp0["mass"])
    return _weighted(Case("mixture_box", m, props, bd, dt, pos, vel, cel, tid, erot, sig0,
                          meta=dict(n=number_density, T0=T0, lam=lam, L=L, Tref=Tref, species=sps, fractions=dens)), cwf)


def couette(nx=1000, ny=500, ppc=20, Kn=0.1, Tw=273.0, Uw=150.0, number_density=1e20, species=("Ar", ARGON_GUIDE),
            Tref=273.0, courant=0.5, seed=2, rank=0, n_ranks=1, binary="variableHardSphere", mode="dsmc", bgk="noBGKCollision", cellWeightFactor=None, **cp):
    """Config 2: 2-D Couette flow, x cyclic, y walls diffuse at Tw moving at -+Uw, z empty; H = lambda/Kn.

    n_ranks > 1 builds rank `rank`'s slab of a channel n_ranks*nx cells long (weak scaling: nx x ny cells per
    rank): the x ends become processor patches to the neighbouring slabs, the periodic wrap a processorCyclic
    pair with a +-n_ranks*Lx separation - what decomposePar makes of the cyclic channel."""
    name, sp = species
    lam = vhs_mean_free_path(number_density, Tw, sp, Tref)
    H = lam / Kn
    dy = H / ny
    dx = dy
    Lx = nx * dx
    if n_ranks == 1:
        kinds = {"xMin": ("left", "cyclic"), "xMax": ("right", "cyclic"), "yMin": ("bottom", "wall"), "yMax": ("top", "wall"),
                 "zMin": ("back", "empty"), "zMax": ("front", "empty")}
        m = _mesh.box_mesh(nx, ny, 1, Lx, H, dx, kinds, cyclic_pairs=[("xMin", "xMax")], solution_d=(1, 1, 0))
    else:
        lo, hi = (rank - 1) % n_ranks, (rank + 1) % n_ranks
        kinds = {"xMin": (f"procBoundary{rank}to{lo}lo", "processor"), "xMax": (f"procBoundary{rank}to{hi}hi", "processor"),
                 "yMin": ("bottom", "wall"), "yMax": ("top", "wall"), "zMin": ("back", "empty"), "zMax": ("front", "empty")}
        m = _mesh.box_mesh(nx, ny, 1, Lx, H, dx, kinds, solution_d=(1, 1, 0), origin=(rank * Lx, 0.0, 0.0))
        pl, ph = m.patches[0], m.patches[1]
        pl.partner, ph.partner = lo, hi
        pl.tag, ph.tag = ("c", 0, 1), ("c", 1, 0)  # my xMin matches the peer's xMax and vice versa
        pl.peer_patch, ph.peer_patch = 1, 0
        if rank == 0:
            pl.separation = (n_ranks * Lx, 0.0, 0.0)
        if rank == n_ranks - 1:
            ph.separation = (-n_ranks * Lx, 0.0, 0.0)
        seed = seed + 7919 * rank
    m.meta_axis_aligned = True
    nParticle = number_density * dx * dy * dx / ppc
    rng = np.random.default_rng(seed)
    cwf = cell_weight_factor(m, cellWeightFactor, number_density, nParticle)
    pos, vel, cel, tid, erot = mesh_fill(m, {name: sp}, [name], {name: number_density}, Tw, (0, 0, 0), nParticle, rng, cell_weight=cwf)
    dt = courant * dx / most_probable_speed(Tw, sp["mass"])
    wallp = lambda p, u: {"patchBoundaryProperties": {"patch": p}, "boundaryModel": "uniGasDiffuseWallPatch",
                          "uniGasDiffuseWallPatchProperties": {"temperature": Tw, "velocity": [u, 0, 0]}}
    bd = {"uniGasPatchBoundaries": [wallp("bottom", -Uw), wallp("top", Uw)]}
    sig0 = math.pi * sp["diameter"] ** 2 * most_probable_speed(Tw, sp["mass"])
    return _weighted(Case("couette", m, _props(name, sp, nParticle, mode, binary, bgk, Tref, **cp), bd, dt, pos, vel, cel, tid,
                          None, sig0, meta=dict(n=number_density, Tw=Tw, Uw=Uw, lam=lam, H=H, Lx=Lx, Tref=Tref, species=sp)), cwf)


def cylinder(nr=100, ntheta=200, ppc=20, n_inf=4.247e20, T_inf=200.0, U_inf=2634.7, T_wall=500.0, r0=0.5 * 0.3048, r1=2.0 * 0.3048,
             lz=0.1 * 0.3048, grading=5.0, species=("Ar", ARGON_TUTORIAL), Tref=1000.0, courant=0.3, seed=3,
             binary="variableHardSphere", mode="dsmc", bgk="noBGKCollision", cellWeightFactor=None, **cp):
    """Config 3: 2-D hypersonic (Mach 10) argon flow over a cylinder - the geometry, free stream and wall of
    tutorials/uniGasFoam/hypersonicCylinder (system/blockMeshDict, system/boundariesDict) without its cell
    weighting / adaptation: free-stream inflow on the upstream half of the outer arc, deleting outflow on the
    downstream half, diffuse isothermal cylinder, symmetry axis."""
    name, sp = species
    m = _mesh.half_annulus_mesh(nr, ntheta, r0, r1, lz, grading)
    m.meta_axis_aligned = False
    n_parcels = ppc * m.n_cells
    nParticle = n_inf * m.cell_volumes.sum() / n_parcels
    rng = np.random.default_rng(seed)
    cwf = cell_weight_factor(m, cellWeightFactor, n_inf, nParticle)
    pos, vel, cel, tid, erot = mesh_fill(m, {name: sp}, [name], {name: n_inf}, T_inf, (U_inf, 0.0, 0.0), nParticle, rng, cell_weight=cwf)
    dr_min = (m.cell_bb_max - m.cell_bb_min)[:, :2].min()
    dt = courant * dr_min / (U_inf + most_probable_speed(T_inf, sp["mass"]))
    inflow = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
              "uniGasFreeStreamInflowPatchProperties": {"typeIds": [name], "numberDensities": {name: n_inf}, "translationalTemperature": T_inf,
                                                        "rotationalTemperature": T_inf, "vibrationalTemperature": T_inf,
                                                        "electronicTemperature": T_inf, "velocity": [U_inf, 0.0, 0.0]}}
    bd = {
        "uniGasPatchBoundaries": [
            {"patchBoundaryProperties": {"patch": "cylinder"}, "boundaryModel": "uniGasDiffuseWallPatch",
             "uniGasDiffuseWallPatchProperties": {"velocity": [0, 0, 0], "temperature": T_wall}},
            {"patchBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasDeletionPatch"},
            {"patchBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasDeletionPatch"},
        ],
        "uniGasGeneralBoundaries": [inflow],
    }
    sig0 = math.pi * sp["diameter"] ** 2 * most_probable_speed(T_inf, sp["mass"])
    return _weighted(Case("cylinder", m, _props(name, sp, nParticle, mode, binary, bgk, Tref, **cp), bd, dt, pos, vel, cel, tid,
                          erot if sp.get("rotationalDegreesOfFreedom", 0) else None, sig0,
                          meta=dict(n=n_inf, T_inf=T_inf, U_inf=U_inf, T_wall=T_wall, r0=r0, r1=r1, Tref=Tref, species=sp)), cwf)


def field_patch_values(boundariesDict, case_dir, start_time, mesh):
    """The *FieldPatch boundary models read volFields from the start time directory and use their values on the model's
    patch: boundaryT / boundaryU for the wall field patches (e.g. …/uniGasDiffuseWallFieldPatch/uniGasDiffuseWallFieldPatch.C:
    56-121), boundaryNumberDensity_<species> / boundaryTransT / boundaryRotT / boundaryU for uniGasFreeStreamInflowFieldPatch
    (…/uniGasFreeStreamInflowFieldPatch.C:65-183).  UniGasCloud takes those values with the dictionary entry; this fills them
    in from the files for every entry that does not carry them already."""
    from . import foamfile
    t0 = os.path.join(case_dir, start_time)
    cache = {}

    def on_patch(fname, patch_name):
        if fname not in cache:
            path = os.path.join(t0, fname)
            if not os.path.exists(path):
                raise FileNotFoundError(f"{patch_name}: a field patch model needs {path}")
            cache[fname] = foamfile.read_vol_field(path)
        p = mesh.patches[mesh.patch_index(patch_name)]
        v = foamfile.boundary_values(cache[fname], patch_name, p.size)
        if v is None:
            raise ValueError(f"{fname}: patch {patch_name} has no value entry")
        return v

    for e in boundariesDict.get("uniGasPatchBoundaries", []):
        if e["boundaryModel"].endswith("FieldPatch"):
            name = e["patchBoundaryProperties"]["patch"]
            for k in ("boundaryT", "boundaryU"):
                if k not in e:
                    e[k] = on_patch(k, name)
    for e in boundariesDict.get("uniGasGeneralBoundaries", []):
        if e["boundaryModel"] == "uniGasFreeStreamInflowFieldPatch":
            name = e["generalBoundaryProperties"]["patch"]
            ids = e[e["boundaryModel"] + "Properties"]["typeIds"]
            if "boundaryNumberDensity" not in e:
                e["boundaryNumberDensity"] = {s: on_patch("boundaryNumberDensity_" + s, name) for s in ids}
            for k in ("boundaryTransT", "boundaryRotT", "boundaryU"):
                if k not in e:
                    e[k] = on_patch(k, name)


def from_case_dir(case_dir, mesh, seed=7, overrides=None, particles_per_cell=None, start_time="0"):
    """A uniGasFoam case directory (constant/uniGasProperties, system/{controlDict, boundariesDict,
    uniGasInitialisationDict, ...}) on a given mesh -> Case: the dictionaries are used as they are
    (unigasfoam_b200.foamdict.load_case), the initial parcels come from the `uniGasMeshFill` configuration
    (U/uniGasInitialisation/derived/uniGasMeshFill/uniGasMeshFill.C:60-296) including its cell-weight rule when
    cellWeightedSimulation is on.  particles_per_cell overrides cellWeightedProperties.particlesPerSubCell (coarse test
    meshes).  Returns (case, loaded) where loaded holds the remaining dictionaries (field properties, hybrid
    decomposition, controlDict)."""
    from . import foamdict
    ld = foamdict.load_case(case_dir, overrides)
    props = ld["uniGasProperties"]
    field_patch_values(ld["boundariesDict"], case_dir, start_time, mesh)
    init = (ld["uniGasInitialisationDict"] or {}).get("configurations", [])
    if len(init) != 1 or init[0].get("type") not in ("uniGasMeshFill", "uniGasMeshFieldFill"):
        raise ValueError("from_case_dir handles a single uniGasMeshFill / uniGasMeshFieldFill configuration")
    cfgi = init[0]
    names = props["typeIdList"]
    FN = float(props["nEquivalentParticles"])
    if cfgi["type"] == "uniGasMeshFieldFill":  # per-cell state from the fields of the start time directory
        from . import foamfile
        t0 = os.path.join(case_dir, start_time)
        fld = lambda f: foamfile.expand_internal(foamfile.read_vol_field(os.path.join(t0, f)), mesh.n_cells)
        dens = {n: fld("numberDensity_" + n) for n in names}
        T, Trot, vel = fld("transT"), fld("rotT"), fld("U")
        ntot = sum(dens.values())
    else:
        dens = {n: float(cfgi["numberDensities"][n]) for n in names}
        ntot = sum(dens.values())
        T = float(cfgi["translationalTemperature"])
        Trot = float(cfgi.get("rotationalTemperature", T))
        vel = [float(v) for v in cfgi["velocity"]]
    deltaT, levels = ld["deltaT"], None
    if props.get("adaptiveSimulation", False):  # uniGasMeshFill.C:101-109: first time step and sub-cell levels from the initial state
        from types import SimpleNamespace
        from .adapter import UniGasDynamicAdapter
        stand_in = SimpleNamespace(mesh=mesh, cellWeighted=bool(props.get("cellWeightedSimulation", False)), _subCellLevels=None,
                                   _cellWeightFactor=None, cfg=SimpleNamespace(deltaT=deltaT, nParticle=FN))
        ad0 = UniGasDynamicAdapter(stand_in, props)
        deltaT, levels = ad0.set_initial_configuration([dens[n] for n in names], T, vel)
        if not ad0.subCellAdaptation:
            levels = None
    cwf = None
    rwf_c = radial_weight(props, mesh.cell_centres)  # axisymmetricSimulation: RWF of the cell centres (uniGasMeshFill.C:117, 204, 260); else 1
    if props.get("cellWeightedSimulation", False):
        pps = particles_per_cell or int(props["cellWeightedProperties"]["particlesPerSubCell"])
        nSub = 1.0 if levels is None else levels.prod(1).astype(float)
        cwf = cell_weight_factor(mesh, ("particlesPerSubCell", pps), ntot, FN) / nSub / rwf_c   # uniGasMeshFill.C:111-121
    rng = np.random.default_rng(seed)
    sps = props["moleculeProperties"]
    fill_w = cwf if not props.get("axisymmetricSimulation", False) else (rwf_c if cwf is None else cwf * rwf_c)
    pos, velp, cel, tid, erot = mesh_fill(mesh, sps, names, dens, T, vel, FN, rng, Trot=Trot, cell_weight=fill_w)
    sp0 = sps[names[int(np.argmax([np.mean(dens[n]) for n in names]))]]
    sig0 = math.pi * sp0["diameter"] ** 2 * most_probable_speed(float(np.mean(T)), sp0["mass"])  # uniGasMeshFill.C:284-296
    any_rot = any(sps[n].get("rotationalDegreesOfFreedom", 0) for n in names)
    case = Case(os.path.basename(os.path.normpath(case_dir)), mesh, props, ld["boundariesDict"], deltaT, pos, velp, cel, tid,
                erot if any_rot else None, sig0, meta=dict(n=float(np.mean(ntot)), T_inf=float(np.mean(T)), U_inf=np.asarray(vel, float).reshape(-1, 3).mean(0), species=sp0, Tref=float(props.get("collisionProperties", {}).get("Tref", 273.0))))
    if cwf is not None:
        case.cellWeightFactor = cwf
    if levels is not None and (levels != 1).any():
        case.subCellLevels = levels
    return case, ld


# ---- BASELINE configs 3 / 4 / 5 at their stated size: per-rank blocks, parcels generated with torch ---------------------

_HEX_TETS = ((0, 1, 3, 7), (0, 3, 2, 7), (0, 2, 6, 7), (0, 6, 4, 7), (0, 4, 5, 7), (0, 5, 1, 7))  # six tets around the 0-7 diagonal


def fill_parcels_torch(mesh, sp, number_density, T, velocity, nParticle, cell_weight=None, Trot=None, seed=1, device=None, chunk=8_000_000):
    """uniGasMeshFill for the large bench / full-size test cases (…/uniGasMeshFill.C:174-278): same rule as mesh_fill -
    N = n V / (F_N CWF) per cell with stochastic rounding, uniform position in the cell, Maxwellian + drift, equipartition
    ERot - but vectorised with torch on `device` (the GPU when there is one: synthetic input generation, not the product
    path) and without mesh_fill's fixed draw order.  Hex cells of structured blocks are split into six tets, a tet is
    chosen by volume and the point is uniform in it (tetPointRef::randomPoint).  One species.
    -> position [n,3], U [n,3], cell [n] (cell-major), ERot [n] or None; numpy, host."""
    import torch
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    f64 = dict(dtype=torch.float64, device=device)
    nC = mesh.n_cells
    vol = torch.as_tensor(mesh.cell_volumes, **f64)
    req = float(number_density) / float(nParticle) * vol
    if cell_weight is not None:
        req = req / torch.as_tensor(np.asarray(cell_weight, float), **f64)
    cnt = torch.floor(req)
    cnt = (cnt + ((req - cnt) > torch.rand(nC, generator=gen, **f64)).to(torch.float64)).to(torch.int64)
    off = torch.zeros(nC + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(cnt, 0)
    n = int(off[-1].item())
    pos = np.empty((n, 3)); vel = np.empty((n, 3)); cel = np.empty(n, np.int32)
    rd = int(sp.get("rotationalDegreesOfFreedom", 0))
    erot = np.empty(n) if rd else None
    axis_aligned = bool(mesh.meta_axis_aligned)
    if axis_aligned:
        lo_all, hi_all = torch.as_tensor(mesh.cell_bb_min, **f64), torch.as_tensor(mesh.cell_bb_max, **f64)
    else:
        corners = torch.as_tensor(mesh.points, **f64)[torch.as_tensor(_mesh.hex_corners(mesh).astype(np.int64), device=device)]  # [nC,8,3]
        tets = torch.as_tensor(_HEX_TETS, dtype=torch.int64, device=device)
        a = corners[:, tets[:, 0]]
        tv = torch.linalg.det(torch.stack([corners[:, tets[:, 1]] - a, corners[:, tets[:, 2]] - a, corners[:, tets[:, 3]] - a], dim=-2)).abs()  # [nC,6]
        cum = torch.cumsum(tv, 1)
    Uinf = torch.as_tensor(np.asarray(velocity, float), **f64)
    sig = math.sqrt(kB * float(T) / sp["mass"])
    mids = [0.5 * (mesh.points[:, d].min() + mesh.points[:, d].max()) for d in range(3)]
    c0 = 0
    offh = off.cpu().numpy()
    while c0 < nC:  # chunks of whole cells holding <= `chunk` parcels
        c1 = int(np.searchsorted(offh, offh[c0] + chunk, side="right")) - 1
        c1 = min(max(c1, c0 + 1), nC)
        b, e = int(offh[c0]), int(offh[c1])
        m = e - b
        if m:
            cl = torch.repeat_interleave(torch.arange(c0, c1, device=device), cnt[c0:c1])
            if axis_aligned:
                p = lo_all[cl] + torch.rand(m, 3, generator=gen, **f64) * (hi_all[cl] - lo_all[cl])
            else:
                u = torch.rand(m, generator=gen, **f64) * cum[cl, 5]
                ti = (u[:, None] >= cum[cl, :5]).sum(1)                      # tet by volume
                v = corners[cl[:, None], tets[ti]]                           # [m,4,3]
                w = -torch.log(1.0 - torch.rand(m, 4, generator=gen, **f64))  # uniform barycentric coordinates
                w = w / w.sum(1, keepdim=True)
                p = (w[:, :, None] * v).sum(1)
            for d in range(3):
                if not mesh.solution_d[d]:
                    p[:, d] = mids[d]
            pos[b:e] = p.cpu().numpy()
            vel[b:e] = (sig * torch.randn(m, 3, generator=gen, **f64) + Uinf).cpu().numpy()
            cel[b:e] = cl.to(torch.int32).cpu().numpy()
            if rd == 2:
                tr = float(T if Trot is None else Trot)
                erot[b:e] = (-torch.log(1.0 - torch.rand(m, generator=gen, **f64)) * (kB * tr)).cpu().numpy()
            elif rd:
                tr = float(T if Trot is None else Trot)
                g = torch.distributions.Gamma(torch.tensor(0.5 * rd, **f64), torch.tensor(1.0, **f64))
                erot[b:e] = (g.sample((m,)) * (kB * tr)).cpu().numpy()
        c0 = c1
    return pos, vel, cel, erot


def _empty_patch_split(m, name, n_first, name_a, name_b, kind_a=None, kind_b=None):
    """split_patch that also accepts n_first = 0 or = size: decomposePar keeps every patch on every rank, with zero faces where
    the rank does not touch it."""
    pi = m.patch_index(name)
    p = m.patches[pi]
    n_first = int(min(max(n_first, 0), p.size))
    a = _mesh.Patch(name_a, kind_a or p.kind, p.start, n_first)
    b = _mesh.Patch(name_b, kind_b or p.kind, p.start + n_first, p.size - n_first)
    m.patches[pi:pi + 1] = [a, b]
    return m


def cylinder_block(rank=0, n_ranks=1, nr=1000, ntheta=2500, parcels=50_000_000, n_inf=4.247e20, T_inf=200.0, U_inf=2634.7, T_wall=500.0,
                   r0=0.5 * 0.3048, r1=2.0 * 0.3048, lz=0.1 * 0.3048, grading=5.0, species=("Ar", ARGON_TUTORIAL), Tref=1000.0, courant=0.3,
                   seed=3, hybrid=False, theta_blend=0.1, device=None):
    """BASELINE configs[2] (hybrid=False: Mach-10 argon cylinder, DSMC NTC + VHS, 50 M parcels, 2.5 M cells) and configs[3]
    (hybrid=True: the same topology at 10 n_inf, USP-SBGK in the upstream half - the compressed fore-body side - and NTC / VHS
    in the wake half, 100 M parcels), SURVEY 8d rows 3 and 4, as rank `rank` of n_ranks blocks along theta (`method simple`,
    n (1 n_ranks 1)): the geometry, free stream and wall of tutorials/uniGasFoam/hypersonicCylinder."""
    name, sp = species
    g = float(grading)

    def pm(I, J, K):
        t = I / nr
        frac = t if abs(g - 1.0) < 1e-12 else (g ** t - 1.0) / (g - 1.0)
        r = r0 + (r1 - r0) * frac
        th = np.pi * J / ntheta
        y = np.where((J == 0) | (J == ntheta), 0.0, r * np.sin(th))  # the axis rows sit on y = 0 exactly
        return r * np.cos(th), y, lz * (K - 0.5)

    kinds = {"xMin": ("cylinder", "wall"), "xMax": ("outer", "patch"), "yMin": ("axisDown", "symmetryPlane"), "yMax": ("axisUp", "symmetryPlane"),
             "zMin": ("back", "empty"), "zMax": ("front", "empty")}
    m, (i0, j0, k0) = _mesh.structured_subblock((nr, ntheta, 1), (1, n_ranks, 1), rank, pm, kinds, solution_d=(1, 1, 0))
    nyl = m.shape[1]
    _empty_patch_split(m, "outer", ntheta // 2 - j0, "outlet", "inlet")
    m.meta_axis_aligned = False
    if hybrid:
        n_inf = 10.0 * n_inf
    volume = 0.5 * math.pi * (r1 * r1 - r0 * r0) * lz  # the polygonal mesh differs from the annulus by O(1/ntheta^2): irrelevant for F_N
    nParticle = n_inf * volume / parcels
    pos, vel, cel, erot = fill_parcels_torch(m, sp, n_inf, T_inf, (U_inf, 0.0, 0.0), nParticle, seed=seed + 7919 * rank, device=device)
    dr_min = (r1 - r0) * ((g ** (1.0 / nr) - 1.0) / (g - 1.0) if abs(g - 1.0) > 1e-12 else 1.0 / nr)
    dt = courant * min(dr_min, math.pi * r0 / ntheta) / (U_inf + most_probable_speed(T_inf, sp["mass"]))
    dens = n_inf
    inflow = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
              "uniGasFreeStreamInflowPatchProperties": {"typeIds": [name], "numberDensities": {name: dens}, "translationalTemperature": T_inf,
                                                        "rotationalTemperature": T_inf, "vibrationalTemperature": T_inf,
                                                        "electronicTemperature": T_inf, "velocity": [U_inf, 0.0, 0.0]}}
    bd = {
        "uniGasPatchBoundaries": [
            {"patchBoundaryProperties": {"patch": "cylinder"}, "boundaryModel": "uniGasDiffuseWallPatch",
             "uniGasDiffuseWallPatchProperties": {"velocity": [0, 0, 0], "temperature": T_wall}},
            {"patchBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasDeletionPatch"},
            {"patchBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasDeletionPatch"},
        ],
        "uniGasGeneralBoundaries": [inflow],
    }
    sig0 = math.pi * sp["diameter"] ** 2 * most_probable_speed(T_inf, sp["mass"])
    if hybrid:
        props = _props(name, sp, nParticle, "hybrid", "variableHardSphere", "unifiedStochasticParticleSBGK", Tref, theta=theta_blend)
    else:
        props = _props(name, sp, nParticle, "dsmc", "variableHardSphere", "noBGKCollision", Tref)
    case = Case("cylinder_hybrid" if hybrid else "cylinder", m, props, bd, dt, pos, vel, cel, None, None, sig0,
                meta=dict(n=n_inf, T_inf=T_inf, U_inf=U_inf, T_wall=T_wall, r0=r0, r1=r1, Tref=Tref, species=sp, block=(i0, j0, k0)))
    if hybrid:
        c = np.arange(m.n_cells)
        jg = (c // m.shape[0]) % nyl + j0
        case.cellCollModelId = (jg < ntheta // 2).astype(np.int32)  # 1 = dsmc (wake half, theta < pi/2), 0 = bgk (upstream half)
    return case


def radial_weight(props, points):
    """uniGasCloud::axiRWF (U/clouds/uniGasCloudI.H:116-120) at points [n,3]; 1 without axisymmetricSimulation."""
    pts = np.asarray(points, float)
    if not props.get("axisymmetricSimulation", False):
        return np.ones(len(pts))
    ap = props["axisymmetricProperties"]
    return 1.0 + (float(ap["maxRadialWeightingFactor"]) - 1.0) * np.sqrt(pts[:, 1] * pts[:, 1] + pts[:, 2] * pts[:, 2]) / float(ap["radialExtentOfDomain"])


def axisymmetric_tube(nx=24, nr=16, length=None, radius=None, ppc=30, n_inf=1e20, T_inf=300.0, U_inf=200.0, T_wall=400.0, max_rwf=8.0,
                      species=("Ar", ARGON_GUIDE), Tref=273.0, binary="variableHardSphere", mode="dsmc", bgk="noBGKCollision",
                      cell_weighted=False, wall="uniGasDiffuseWallPatch", inlet="uniGasFreeStreamInflowPatch", half_angle_deg=0.5,
                      courant=0.3, lambda_per_dx=2.0, seed=3, gr=1.0, device="cpu", **cp):
    """Axisymmetric flow through a tube, the set-up of the reference's axisymmetric tutorials (plumeImpingement,
    expansionInVacuum) in small: a wedge of +-half_angle about the x axis, one cell thick, side patches symmetryPlane, the
    axis a collapsed symmetry patch; free-stream (or Liou-Fang pressure) inlet at x = 0, deletion outlet at x = length, wall at
    r = radius; axisymmetricSimulation true with radialExtentOfDomain = radius.  Filled as uniGasMeshFill does it: N = n V /
    (F_N CWF RWF(cell centre)) parcels per cell, which then carry RWF(cell centre) (uniGasMeshFill.C:196-260).  With
    cell_weighted the factor field is uniGasMeshFill's n V / (ppc F_N RWF) (:111-121), so every cell starts with ppc parcels."""
    name, sp = species
    lam = vhs_mean_free_path(n_inf, T_inf, sp, Tref)
    dx = lam / lambda_per_dx
    length = nx * dx if length is None else length
    radius = nr * dx if radius is None else radius
    kinds = {"xMin": ("inlet", "patch"), "xMax": ("outlet", "patch"), "yMin": ("axis", "symmetry"), "yMax": ("wall", "wall"),
             "zMin": ("backWedge", "symmetryPlane"), "zMax": ("frontWedge", "symmetryPlane")}
    m = _mesh.structured_block(nx, nr, 1, _mesh.wedge_map(nx, nr, length, radius, half_angle_deg, gr=gr), kinds)
    m.meta_axis_aligned = False
    volume = m.cell_volumes.sum()
    mode_bgk = mode != "dsmc"
    props = _props(name, sp, 1.0, mode, binary if mode != "bgk" else "noDSMCCollision", bgk if mode_bgk else "noBGKCollision", Tref, **cp)
    props["axisymmetricSimulation"] = True
    props["axisymmetricProperties"] = {"radialExtentOfDomain": radius, "maxRadialWeightingFactor": max_rwf}
    rwf_c = radial_weight(props, m.cell_centres)
    if cell_weighted:
        nParticle = n_inf * np.median(m.cell_volumes / rwf_c) / ppc   # the median cell carries factor 1
        cwf = n_inf * m.cell_volumes / (ppc * nParticle * rwf_c)
        props["cellWeightedSimulation"] = True
        props["cellWeightedProperties"] = {"particlesPerSubCell": ppc, "minParticlesPerSubCell": ppc}
    else:
        nParticle = n_inf * (m.cell_volumes / rwf_c).sum() / (ppc * m.n_cells)
        cwf = None
    props["nEquivalentParticles"] = nParticle
    pos, vel, cel, erot = fill_parcels_torch(m, sp, n_inf, T_inf, (U_inf, 0.0, 0.0), nParticle, cell_weight=rwf_c if cwf is None else cwf * rwf_c,
                                             Trot=T_inf, seed=seed, device=device)
    dt = courant * dx / (abs(U_inf) + most_probable_speed(T_inf, sp["mass"]))
    if inlet == "uniGasFreeStreamInflowPatch":
        general = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": inlet,
                   inlet + "Properties": {"typeIds": [name], "numberDensities": {name: n_inf}, "translationalTemperature": T_inf,
                                          "rotationalTemperature": T_inf, "velocity": [U_inf, 0.0, 0.0]}}
    else:  # the pressure inlets: uniGasLiouFangPressureInletPatch / uniGasWangPressureInletPatch
        general = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": inlet,
                   inlet + "Properties": {"typeIds": [name], "moleFractions": {name: 1.0}, "inletPressure": n_inf * kB * T_inf,
                                          "inletTemperature": T_inf, "theta": 0.2}}
    wall_entry = {"patchBoundaryProperties": {"patch": "wall"}, "boundaryModel": wall}
    if wall != "uniGasSpecularWallPatch":
        wall_entry[wall + "Properties"] = {"velocity": [0, 0, 0], "temperature": T_wall}
    bd = {"uniGasPatchBoundaries": [wall_entry,
                                    {"patchBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasDeletionPatch"},
                                    {"patchBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasDeletionPatch"}],
          "uniGasGeneralBoundaries": [general]}
    sig0 = math.pi * sp["diameter"] ** 2 * most_probable_speed(T_inf, sp["mass"])
    case = Case("axisymmetric_tube", m, props, bd, dt, pos, vel, cel, None, erot, sig0,
                meta=dict(n=n_inf, T_inf=T_inf, U_inf=U_inf, T_wall=T_wall, Tref=Tref, species=sp, radius=radius, length=length, lam=lam,
                          rwf_centre=rwf_c, volume=volume))
    if cwf is not None:
        case.cellWeightFactor = np.ascontiguousarray(cwf)
    return case


def blunt_body_block(rank=0, n_ranks=8, n_eta=200, n_s=500, n_phi=250, ppc=20, n_inf=5e21, T_inf=200.0, mach=10.0, T_wall=500.0,
                     species=("N2", NITROGEN), Tref=273.0, courant=0.3, seed=5, zrot=5.0, zelec=50.0, device=None, **geom):
    """BASELINE configs[4]: 3-D nitrogen Mach-10 flow over a blunted cone (sphere-cone fore-body, body-fitted
    grid revolved about the axis), Larsen-Borgnakke rotational relaxation, cell-weighted as every reference tutorial is
    (ppc parcels in every cell at the free-stream state), 62.5 M parcels and 3.1 M cells per rank at the default size.
    The full case is 8 blocks - 4 along the body x 2 in azimuth over a 90 degree sector between two symmetry planes; with
    n_ranks < 8 the case is the part of it those ranks own: 4 ranks = the 45 degree sector (by symmetry the same flow), 2 / 1
    ranks = its first two / first body blocks with the downstream cut as outflow.  `method simple`, n (1 4 2)."""
    name, sp = species
    if n_ranks not in (1, 2, 4, 8):
        raise ValueError("blunt_body_block: 1, 2, 4 or 8 ranks")
    ps, pp = (4, 2)
    ns_blocks = min(n_ranks, 4)
    np_blocks = 2 if n_ranks == 8 else 1
    rs = _mesh.block_ranges(n_s, ps)
    rp = _mesh.block_ranges(n_phi, pp)
    gs = rs[ns_blocks - 1][1]                      # body cells of the part that is run
    gp = rp[np_blocks - 1][1]
    pm = _mesh.sphere_cone_map(n_eta, n_s, n_phi, **geom)
    kinds = {"xMin": ("body", "wall"), "xMax": ("inlet", "patch"), "yMin": ("axis", "symmetry"), "yMax": ("outlet", "patch"),
             "zMin": ("symmetryA", "symmetryPlane"), "zMax": ("symmetryB", "symmetryPlane")}
    m, blk = _mesh.structured_subblock((n_eta, gs, gp), (1, ns_blocks, np_blocks), rank, pm, kinds)
    m.meta_axis_aligned = False
    a = math.sqrt(1.4 * kB * T_inf / sp["mass"])
    U_inf = mach * a
    target = ppc
    # cellWeightedSimulation, uniGasMeshFill's rule (uniGasMeshFill.C:111-121): CWF = n V / (particlesPerSubCell F_N); F_N such
    # that the median cell carries factor 1
    # the same on every rank: tied to the cell in the middle of the global grid, which then carries factor 1
    nParticle = n_inf * _blunt_reference_volume(n_eta, n_s, n_phi, geom) / target
    cwf = n_inf * m.cell_volumes / (target * nParticle)
    pos, vel, cel, erot = fill_parcels_torch(m, sp, n_inf, T_inf, (U_inf, 0.0, 0.0), nParticle, cell_weight=cwf, Trot=T_inf,
                                             seed=seed + 7919 * rank, device=device)
    dt = _blunt_reference_dt(n_eta, n_s, n_phi, geom, courant, U_inf + most_probable_speed(T_inf, sp["mass"]))
    inflow = {"generalBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasFreeStreamInflowPatch",
              "uniGasFreeStreamInflowPatchProperties": {"typeIds": [name], "numberDensities": {name: n_inf}, "translationalTemperature": T_inf,
                                                        "rotationalTemperature": T_inf, "vibrationalTemperature": T_inf,
                                                        "electronicTemperature": T_inf, "velocity": [U_inf, 0.0, 0.0]}}
    bd = {
        "uniGasPatchBoundaries": [
            {"patchBoundaryProperties": {"patch": "body"}, "boundaryModel": "uniGasDiffuseWallPatch",
             "uniGasDiffuseWallPatchProperties": {"velocity": [0, 0, 0], "temperature": T_wall}},
            {"patchBoundaryProperties": {"patch": "inlet"}, "boundaryModel": "uniGasDeletionPatch"},
            {"patchBoundaryProperties": {"patch": "outlet"}, "boundaryModel": "uniGasDeletionPatch"},
        ],
        "uniGasGeneralBoundaries": [inflow],
    }
    props = _props(name, sp, nParticle, "dsmc", "LarsenBorgnakkeVariableHardSphere", "noBGKCollision", Tref,
                   rotationalRelaxationCollisionNumber=zrot, electronicRelaxationCollisionNumber=zelec)
    props["cellWeightedSimulation"] = True
    props["cellWeightedProperties"] = {"particlesPerSubCell": target}
    sig0 = math.pi * sp["diameter"] ** 2 * most_probable_speed(T_inf, sp["mass"])
    case = Case("blunt_body", m, props, bd, dt, pos, vel, cel, None, erot, sig0,
                meta=dict(n=n_inf, T_inf=T_inf, U_inf=U_inf, T_wall=T_wall, Tref=Tref, species=sp, block=blk))
    case.cellWeightFactor = np.ascontiguousarray(cwf)
    return case


def _blunt_reference_cell(n_eta, n_s, n_phi, geom):
    """Extents (normal, along the body, azimuthal) and volume of the cell in the middle of the global blunt-body grid: the
    rank-independent reference F_N and deltaT are tied to."""
    pm = _mesh.sphere_cone_map(n_eta, n_s, n_phi, **geom)
    i, j, k = n_eta // 2, n_s // 2, n_phi // 2
    I, J, K = np.meshgrid(np.array([i, i + 1.0]), np.array([j, j + 1.0]), np.array([k, k + 1.0]), indexing="ij")
    x, y, z = pm(I, J, K)
    P = np.stack([np.broadcast_to(x, I.shape), np.broadcast_to(y, I.shape), np.broadcast_to(z, I.shape)], -1)
    d_eta = np.linalg.norm(P[1, 0, 0] - P[0, 0, 0])
    d_s = np.linalg.norm(P[0, 1, 0] - P[0, 0, 0])
    d_phi = np.linalg.norm(P[0, 0, 1] - P[0, 0, 0])
    return d_eta, d_s, d_phi


def _blunt_reference_volume(n_eta, n_s, n_phi, geom):
    a, b, c = _blunt_reference_cell(n_eta, n_s, n_phi, geom)
    return a * b * c


def _blunt_reference_dt(n_eta, n_s, n_phi, geom, courant, speed):
    return courant * min(_blunt_reference_cell(n_eta, n_s, n_phi, geom)) / speed


# ---- internal energy modes beyond rotation (a9: vibrational quantum levels, electronic levels) ---------------------------

# a diatomic with one vibrational mode and three electronic levels, oxygen-like (Bird 1994 App. A for the VHS constants; level
# energies / degeneracies of the O2 X, a, b states).  Zref is set low on purpose where a test wants fast vibrational relaxation.
OXYGEN_VIB = dict(mass=53.12e-27, diameter=4.07e-10, omega=0.77, alpha=1.0, rotationalDegreesOfFreedom=2, vibrationalModes=1,
                  characteristicVibrationalTemperature=[2256.0], dissociationTemperature=[59500.0], Zref=[17900.0], referenceTempForZref=[2256.0],
                  charge=0, numberOfElectronicLevels=3, electronicEnergyList=[0.0, 1.5727e-19, 2.6203e-19], degeneracyList=[3, 2, 1])


def equilibrium_levels(sp, T, n, rng):
    """Boltzmann-distributed vibrational quantum levels [n, nModes] (harmonic oscillator: geometric in the level, which is what
    equipartitionVibrationalEnergyLevel's int(-ln(R) T / thetaV) draws, U/clouds/uniGasCloud.C:1020-1050) and electronic levels
    [n] (g_j exp(-E_j / k T), sampled exactly) of one species at temperature T."""
    nm = int(sp.get("vibrationalModes", 0))
    vib = np.zeros((n, max(nm, 1)), np.int32)
    for m in range(nm):
        vib[:, m] = np.floor(-np.log(1.0 - rng.random(n)) * T / sp["characteristicVibrationalTemperature"][m]).astype(np.int32)
    E = np.asarray(sp.get("electronicEnergyList", [0.0]), float)
    g = np.asarray(sp.get("degeneracyList", [1]), float)
    w = g * np.exp(-E / (kB * T))
    elev = rng.choice(len(E), size=n, p=w / w.sum()).astype(np.int32)
    return vib[:, :nm] if nm else None, elev


def with_internal_modes(case, Tvib=None, Tel=None, seed=11):
    """Give the parcels of a single-species case vibrational / electronic levels at the given temperatures (default: the case's
    temperature)."""
    name = case.uniGasProperties["typeIdList"][0]
    sp = case.uniGasProperties["moleculeProperties"][name]
    T0 = case.meta.get("T0", case.meta.get("Tw", case.meta.get("T_inf")))
    rng = np.random.default_rng(seed)
    vib, _ = equilibrium_levels(sp, T0 if Tvib is None else Tvib, case.n_parcels, rng)
    _, elev = equilibrium_levels(sp, T0 if Tel is None else Tel, case.n_parcels, rng)
    case.vibLevel, case.ELevel = vib, elev
    return case
