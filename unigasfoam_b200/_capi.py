"""ctypes view of include/ugf.h.

The same struct layouts serve libugf.so (prefix ``ugf_``) and, in tests only, the CPU
oracle (prefix ``ugfo_``): `Api(lib, prefix)` resolves every entry point the header
declares and fails loudly if one is missing.
"""
import ctypes as C
import os

UGF_ABI_VERSION = 4
UGF_MAX_SPECIES = 8
UGF_MAX_VIB_MODES = 4
UGF_MAX_ELEC_LEVELS = 16
UGF_NMOM = 32
UGF_NBM = 16
UGF_NPHASE = 7
UGF_NFIELD = 21
UGF_NINT = 8
UGF_NWALLFIELD = 12
UGF_NFT = 6
UGF_MIGRATE_STRIDE = 10

# run-time selection tables: dictionary word -> enum  (include/ugf.h cites the reference tables)
COLLISION_MODEL = {"dsmc": 1, "bgk": 2, "hybrid": 3}
PARTNER_MODEL = {"noTimeCounter": 1, "noTimeCounterSubCycled": 2}
BINARY_MODEL = {
    "noDSMCCollision": 0,
    "variableHardSphere": 1,
    "variableSoftSphere": 2,
    "LarsenBorgnakkeVariableHardSphere": 3,
    "LarsenBorgnakkeVariableSoftSphere": 4,
}
BGK_MODEL = {
    "noBGKCollision": 0,
    "stochasticParticleBGK": 1,
    "stochasticParticleESBGK": 2,
    "stochasticParticleSBGK": 3,
    "unifiedStochasticParticleSBGK": 4,
}
PATCH_KIND = {"wall": 1, "symmetry": 2, "symmetryPlane": 2, "cyclic": 3, "empty": 4, "processor": 5, "patch": 6}
WALL_MODEL = {
    "uniGasDiffuseWallPatch": 1,
    "uniGasSpecularWallPatch": 2,
    "uniGasMixedDiffuseSpecularWallPatch": 3,
    "uniGasDeletionPatch": 4,
    "uniGasCLLWallPatch": 5,
    # *FieldPatch variants: same kernels, wall T / U per face (boundaryT / boundaryU)
    "uniGasDiffuseWallFieldPatch": 1,
    "uniGasMixedDiffuseSpecularWallFieldPatch": 3,
    "uniGasCLLWallFieldPatch": 5,
}

i32, i64, u64, f64 = C.c_int32, C.c_int64, C.c_uint64, C.c_double
P = C.POINTER


class Config(C.Structure):
    _fields_ = [
        ("abiVersion", i32), ("device", i32), ("seed", u64), ("nParticle", f64), ("deltaT", f64),
        ("solutionD", i32 * 3), ("collisionModel", i32), ("partnerModel", i32), ("binaryModel", i32),
        ("bgkModel", i32), ("nSubCycles", i32), ("macroInterpolation", i32), ("Tref", f64), ("theta", f64),
        ("rotationalRelaxationCollisionNumber", f64), ("electronicRelaxationCollisionNumber", f64),
        ("parcelCapacity", i64), ("sampleInterval", i32), ("measureWalls", i32), ("rank", i32), ("nRanks", i32),
        ("axisymmetric", i32), ("radialExtent", f64), ("maxRWF", f64),
    ]


class Species(C.Structure):
    _fields_ = [
        ("mass", f64), ("d", f64), ("omega", f64), ("alpha", f64), ("rotationalDoF", i32), ("vibrationalDoF", i32),
        ("thetaV", f64 * UGF_MAX_VIB_MODES), ("thetaD", f64 * UGF_MAX_VIB_MODES), ("Zref", f64 * UGF_MAX_VIB_MODES),
        ("TrefZv", f64 * UGF_MAX_VIB_MODES), ("charge", i32), ("nElectronicLevels", i32),
        ("electronicEnergy", f64 * UGF_MAX_ELEC_LEVELS), ("degeneracy", i32 * UGF_MAX_ELEC_LEVELS),
    ]


class Mesh(C.Structure):
    _fields_ = [
        ("nCells", i32), ("nFaces", i32), ("nInternalFaces", i32), ("nPatches", i32), ("nPoints", i32),
        ("owner", P(i32)), ("neighbour", P(i32)), ("faceAreas", P(f64)), ("faceCentres", P(f64)),
        ("cellFaceOffsets", P(i32)), ("cellFaces", P(i32)), ("cellVolumes", P(f64)), ("cellCentres", P(f64)),
        ("cellBbMin", P(f64)), ("cellBbMax", P(f64)), ("patchStart", P(i32)), ("patchSize", P(i32)),
        ("patchKind", P(i32)), ("patchPartner", P(i32)), ("patchSeparation", P(f64)),
        ("points", P(f64)), ("facePointOffsets", P(i32)), ("facePoints", P(i32)),
    ]


class CellPoint(C.Structure):
    _fields_ = [
        ("nPoints", i32), ("points", P(f64)), ("tetOffsets", P(i32)), ("tetPoints", P(i32)), ("pointCellOffsets", P(i32)),
        ("pointCells", P(i32)), ("pointWeights", P(f64)), ("pointNormals", P(f64)),
    ]


class Inflow(C.Structure):
    _fields_ = [
        ("nTypeIds", i32), ("typeIds", i32 * UGF_MAX_SPECIES), ("numberDensities", f64 * UGF_MAX_SPECIES),
        ("translationalTemperature", f64), ("rotationalTemperature", f64), ("vibrationalTemperature", f64),
        ("electronicTemperature", f64), ("velocity", f64 * 3),
    ]


class PressureInlet(C.Structure):
    _fields_ = [
        ("nTypeIds", i32), ("typeIds", i32 * UGF_MAX_SPECIES), ("moleFractions", f64 * UGF_MAX_SPECIES),
        ("inletPressure", f64), ("inletTemperature", f64), ("theta", f64),
    ]


class Parcels(C.Structure):
    _fields_ = [
        ("n", i64), ("x", P(f64)), ("y", P(f64)), ("z", P(f64)), ("Ux", P(f64)), ("Uy", P(f64)), ("Uz", P(f64)),
        ("cell", P(i32)), ("typeId", P(i32)), ("ERot", P(f64)), ("newParcel", P(i32)), ("cellWeight", P(f64)),
        ("vibLevel", P(i32)), ("ELevel", P(i32)), ("radialWeight", P(f64)),
    ]


class Decomposition(C.Structure):
    _fields_ = [
        ("decompositionInterval", i32), ("resetAtDecomposition", i32), ("resetAtDecompositionUntilTime", f64),
        ("breakdownMax", f64), ("theta", f64), ("smoothingPasses", i32), ("refinementPasses", i32), ("neighborLevels", i32),
        ("maxNeighborFraction", f64),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("step", i64), ("nParcels", i64), ("collisionCandidates", i64), ("collisions", i64), ("bgkRelaxations", i64),
        ("inserted", i64), ("deleted", i64), ("migrated", i64), ("wallHits", i64), ("stuck", i64),
        ("linearKineticEnergy", f64), ("rotationalEnergy", f64), ("momentum", f64 * 3), ("cloned", i64), ("weightDeleted", i64),
        ("vibrationalEnergy", f64), ("electronicEnergy", f64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "momentum"}
        d["momentum"] = list(self.momentum)
        return d


H = C.c_void_p
PF, PI32, PI64 = P(f64), P(i32), P(i64)

# name -> (restype, argtypes); names are given without the library prefix
SIGNATURES = {
    "create": (C.c_int, [P(Config), P(H)]),
    "destroy": (C.c_int, [H]),
    "last_error": (C.c_char_p, [H]),
    "abi_version": (C.c_int, []),
    "set_species": (C.c_int, [H, i32, P(Species)]),
    "set_mesh": (C.c_int, [H, P(Mesh)]),
    "set_patch_model": (C.c_int, [H, i32, i32, PF, i32]),
    "set_patch_wall_fields": (C.c_int, [H, i32, PF, PF]),
    "set_inflow": (C.c_int, [H, i32, P(Inflow)]),
    "set_chapman_enskog_inflow": (C.c_int, [H, i32, P(Inflow), PF, PF]),
    "set_inflow_fields": (C.c_int, [H, i32, i32, PI32, PF, PF, PF, PF]),
    "set_pressure_inlet": (C.c_int, [H, i32, P(PressureInlet)]),
    "set_wang_pressure_inlet": (C.c_int, [H, i32, P(PressureInlet)]),
    "set_pressure_outlet": (C.c_int, [H, i32, P(PressureInlet)]),
    "set_mass_flow_inlet": (C.c_int, [H, i32, P(PressureInlet), f64, P(f64)]),
    "download_inlet_velocity": (C.c_int, [H, i32, PF]),
    "upload_parcels": (C.c_int, [H, P(Parcels)]),
    "upload_cell_state": (C.c_int, [H, PF, PI32, PI32, PF]),
    "set_deltaT": (C.c_int, [H, f64]),
    "set_time_index": (C.c_int, [H, i64]),
    "state_size": (C.c_int, [H, PI64]),
    "state_save": (C.c_int, [H, PF, i64]),
    "state_load": (C.c_int, [H, PF, i64]),
    "download_accumulators": (C.c_int, [H, PF, PF, PF, PI64]),
    "download_internal_accumulators": (C.c_int, [H, PF]),
    "set_face_tracker": (C.c_int, [H, i32, PI32]),
    "download_face_tracker": (C.c_int, [H, PF, i32]),
    "step": (C.c_int, [H, i32]),
    "control_before_move": (C.c_int, [H]),
    "move": (C.c_int, [H]),
    "sort": (C.c_int, [H]),
    "reorder": (C.c_int, [H]),
    "sample": (C.c_int, [H]),
    "collide": (C.c_int, [H]),
    "relax": (C.c_int, [H]),
    "accumulate_fields": (C.c_int, [H]),
    "set_decomposition": (C.c_int, [H, P(Decomposition)]),
    "set_macro_interpolation": (C.c_int, [H, P(CellPoint)]),
    "decompose": (C.c_int, [H]),
    "download_decomposition": (C.c_int, [H, PI32, PF]),
    "end_step": (C.c_int, [H]),
    "finish_step": (C.c_int, [H]),
    "migrate_counts": (C.c_int, [H, PI64]),
    "migrate_pack": (C.c_int, [H, i32, P(PF), PI64]),
    "migrate_unpack": (C.c_int, [H, i32, PF, i64]),
    "move_received": (C.c_int, [H]),
    "migrate_pack_slots": (C.c_int, [H, PF, i64]),
    "migrate_unpack_slots": (C.c_int, [H, PF, i64]),
    "peer_alloc": (C.c_int, [H, i64, P(C.c_void_p), C.c_char_p]),
    "peer_open": (C.c_int, [H, C.c_char_p, P(C.c_void_p)]),
    "migrate_pack_peer": (C.c_int, [H, P(C.c_void_p), P(C.c_void_p), i64, C.c_uint64]),
    "migrate_unpack_peer": (C.c_int, [H, C.c_void_p, C.c_void_p, i64, C.c_uint64]),
    "migrate_inflight": (C.c_int, [H, P(PI64)]),
    "stream": (C.c_int, [H, P(C.c_void_p)]),
    "counters_get": (C.c_int, [H, P(Counters)]),
    "num_parcels": (C.c_int, [H, PI64]),
    "download_parcels": (C.c_int, [H, P(Parcels)]),
    "download_cell_occupancy": (C.c_int, [H, PI32, PI32]),
    "download_cell_moments": (C.c_int, [H, PF]),
    "download_cell_state": (C.c_int, [H, PF, PF, PF, PF]),
    "download_fields": (C.c_int, [H, PF, PF, i32]),
    "download_boundary_meas": (C.c_int, [H, PF]),
    "phase_times": (C.c_int, [H, PF]),
    "launch_count": (C.c_int, [H, PI64]),
    "transfer_bytes": (C.c_int, [H, PI64, PI64, PI64]),
    "host_alloc": (C.c_int, [H, i64, P(C.c_void_p)]),
    "host_free": (C.c_int, [H, C.c_void_p]),
}


class UgfError(RuntimeError):
    pass


class Api:
    """Resolved entry points of one shared library."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise UgfError(f"shared library not found: {path}")
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, prefix + name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)
        if self.abi_version() != UGF_ABI_VERSION:
            raise UgfError(f"{path}: ABI version {self.abi_version()} != {UGF_ABI_VERSION}")


_HERE = os.path.dirname(os.path.abspath(__file__))
LIBUGF_PATH = os.path.join(_HERE, "libugf.so")
_libugf = None


def libugf():
    """The CUDA library.  There is no CPU fallback: a missing library is an error."""
    global _libugf
    if _libugf is None:
        if not os.path.exists(LIBUGF_PATH):
            raise UgfError(
                f"{LIBUGF_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). unigasfoam_b200 has no CPU fallback."
            )
        _libugf = Api(LIBUGF_PATH, "ugf_")
    return _libugf
