"""Host-side mirror of uniGasCloud (U/clouds/uniGasCloud.{H,C}) over the libugf C ABI.

The constructor takes the same dictionaries a uniGasFoam case holds
(constant/uniGasProperties, system/boundariesDict, controlDict's deltaT) as Python
dicts with the reference's key names, resolves the run-time-selected model names
through the same words the reference's selection tables use, and forwards every
per-step call to the CUDA library.  Method names follow the reference
(evolve, info, cellOccupancy, buildCellOccupancy, ...).

There is no CPU path here: `api` defaults to libugf.so and construction fails if it is
missing.  Tests pass the oracle's Api object to drive the CPU restatement through the
same code (oracle/oracle_cloud.py).
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import UgfError


def _species_struct(d):
    """moleculeProperties.<name> -> ugf_species (U/parcels/uniGasParcelI.H:40-141)."""
    s = _capi.Species()
    s.mass = float(d["mass"])
    s.d = float(d["diameter"])
    s.omega = float(d["omega"])
    s.alpha = float(d.get("alpha", 1.0))
    s.rotationalDoF = int(d.get("rotationalDegreesOfFreedom", 0))
    s.vibrationalDoF = int(d.get("vibrationalModes", 0))
    for i, v in enumerate(d.get("characteristicVibrationalTemperature", [])):
        s.thetaV[i] = float(v)
    for i, v in enumerate(d.get("dissociationTemperature", [])):
        s.thetaD[i] = float(v)
    for i, v in enumerate(d.get("Zref", [])):
        s.Zref[i] = float(v)
    for i, v in enumerate(d.get("referenceTempForZref", [])):
        s.TrefZv[i] = float(v)
    s.charge = int(d.get("charge", 0))
    elist = list(d.get("electronicEnergyList", [0.0]))
    glist = list(d.get("degeneracyList", [1]))
    s.nElectronicLevels = int(d.get("numberOfElectronicLevels", len(elist)))
    for i, v in enumerate(elist):
        s.electronicEnergy[i] = float(v)
    for i, v in enumerate(glist):
        s.degeneracy[i] = int(v)
    return s


def _lookup(table, word, key):
    if word not in table:
        # FatalIOErrorInLookup (e.g. U/dsmcCollisions/basic/dsmcCollisionModel/dsmcCollisionModel.C:71-80)
        raise UgfError(f"Unknown {key} type {word!r}. Valid types: {sorted(table)}")
    return table[word]


class UniGasCloud:
    migrate_buffers = "device"   # ugf_migrate_pack hands out device pointers (the oracle's are host pointers)

    def __init__(self, mesh, uniGasProperties, boundariesDict=None, deltaT=None, *, api=None, device=0,
                 seed=20261017, parcelCapacity=None, sampleInterval=1, measureWalls=True, rank=0, nRanks=1):
        self.api = api if api is not None else _capi.libugf()
        self.mesh = mesh
        props = uniGasProperties
        self.typeIdList = list(props["typeIdList"])
        cp = props.get("collisionProperties", {})
        cfg = _capi.Config()
        cfg.abiVersion = _capi.UGF_ABI_VERSION
        cfg.device = device
        cfg.seed = seed + rank
        cfg.nParticle = float(props["nEquivalentParticles"])
        cfg.deltaT = float(deltaT)
        for k in range(3):
            cfg.solutionD[k] = int(mesh.solution_d[k])
        mode = props.get("collisionModel", "dsmc")
        cfg.collisionModel = _lookup(_capi.COLLISION_MODEL, mode, "collisionModel")
        cfg.partnerModel = _lookup(_capi.PARTNER_MODEL, props.get("dsmcCollisionPartnerModel", "noTimeCounter"), "dsmcCollisionPartnerModel")
        cfg.binaryModel = _lookup(_capi.BINARY_MODEL, props.get("dsmcCollisionModel", "noDSMCCollision"), "dsmcCollisionModel")
        cfg.bgkModel = _lookup(_capi.BGK_MODEL, props.get("bgkCollisionModel", "noBGKCollision"), "bgkCollisionModel")
        cfg.nSubCycles = int(cp.get("nSubCycles", 1))
        cfg.macroInterpolation = int(bool(cp.get("macroInterpolation", False)))
        cfg.Tref = float(cp.get("Tref", 273.0))
        cfg.theta = float(cp.get("theta", 1.0))
        cfg.rotationalRelaxationCollisionNumber = float(cp.get("rotationalRelaxationCollisionNumber", 5.0))
        cfg.electronicRelaxationCollisionNumber = float(cp.get("electronicRelaxationCollisionNumber", 500.0))
        cfg.parcelCapacity = int(parcelCapacity) if parcelCapacity else 0
        cfg.sampleInterval = int(sampleInterval)
        cfg.measureWalls = int(bool(measureWalls))
        cfg.rank, cfg.nRanks = rank, nRanks
        # cellWeightedSimulation (U/clouds/uniGasCloud.C:417): the cellWeightFactor field itself comes through
        # setCellState(cellWeightFactor=...) before the parcels, as the reference reads uniGasCellWeightFactor
        self.cellWeighted = bool(props.get("cellWeightedSimulation", False))
        # adaptiveSimulation: unigasfoam_b200.adapter.UniGasDynamicAdapter drives the cloud (host side, every adaptationInterval steps)
        self.adaptive = bool(props.get("adaptiveSimulation", False))
        # axisymmetricSimulation (U/clouds/uniGasCloud.C:420, 563-568): radial weighting on a wedge mesh about the x axis
        self.axisymmetric = bool(props.get("axisymmetricSimulation", False))
        if self.axisymmetric:
            ap = props["axisymmetricProperties"]
            cfg.axisymmetric = 1
            cfg.radialExtent = float(ap["radialExtentOfDomain"])
            cfg.maxRWF = float(ap["maxRadialWeightingFactor"])
        if props.get("chemicalReactions", False):
            raise UgfError("chemicalReactions true is not supported by the B200 path (SURVEY §2: out of scope)")
        self.cfg = cfg
        self._h = _capi.H()
        self._cellCollModelId = self._subCellLevels = self._cellWeightFactor = None  # host copies for writeTime
        self._adapter = None
        self._cwfCarried = None    # the factor field the parcels still carry while a newer one waits for the next weighting pass
        self._nParcelsSet = False
        self._pending_capacity = cfg.parcelCapacity == 0
        self._created = False
        self._species = (_capi.Species * len(self.typeIdList))(
            *[_species_struct(props["moleculeProperties"][name]) for name in self.typeIdList])
        self._boundariesDict = boundariesDict or {}
        if not self._pending_capacity:
            self._create()

    # -- construction -----------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self.api.last_error(self._h if self._created else None)
            raise UgfError(msg.decode() if msg else f"libugf status {rc}")

    def _create(self):
        rc = self.api.create(C.byref(self.cfg), C.byref(self._h))
        if rc != 0:
            msg = self.api.last_error(None)
            raise UgfError(msg.decode() if msg else "ugf_create failed")
        self._created = True
        self._check(self.api.set_species(self._h, len(self.typeIdList), self._species))
        cm = self.mesh.as_c()
        self._check(self.api.set_mesh(self._h, C.byref(cm)))
        bd = self._boundariesDict
        for entry in bd.get("uniGasPatchBoundaries", []):
            patch = self.mesh.patch_index(entry["patchBoundaryProperties"]["patch"])
            word = entry["boundaryModel"]
            model = _lookup(_capi.WALL_MODEL, word, "boundaryModel")
            pr = entry.get(word + "Properties", {})
            params = []
            field = word.endswith("FieldPatch")
            if field:
                # the reference reads the volFields boundaryT / boundaryU from the time directory
                # (…WallFieldPatch.C:56-80); here their values on this patch come with the dictionary entry
                nF = self.mesh.patches[patch].size
                bT = self._f64(np.broadcast_to(np.asarray(entry["boundaryT"], float), (nF,)))
                bU = self._f64(np.broadcast_to(np.asarray(entry["boundaryU"], float), (nF, 3)))
                pr = dict(pr, temperature=float(bT.mean()), velocity=[float(v) for v in bU.mean(0)])
            if model in (1, 3, 5):
                params = [float(pr["temperature"])] + [float(v) for v in pr["velocity"]]
                if model == 3:
                    params.append(float(pr["diffuseFraction"]))
                if model == 5:  # uniGasCLLWallPatch.C:49-55
                    params += [float(pr["normalAccommCoeff"]), float(pr["tangentialAccommCoeff"]), float(pr["rotEnergyAccommCoeff"])]
            if self.mesh.patches[patch].kind != "wall":
                # uniGasBoundaries.C:448-488 wants a model on every non-constraint patch, but only wall
                # patches ever reach controlParticle: accept and ignore.
                continue
            arr = (C.c_double * max(len(params), 1))(*params)
            self._check(self.api.set_patch_model(self._h, patch, model, arr, len(params)))
            if field:
                PD = C.POINTER(C.c_double)
                self._check(self.api.set_patch_wall_fields(self._h, patch, bT.ctypes.data_as(PD), bU.ctypes.data_as(PD)))
        if self.cfg.macroInterpolation:  # interpolationCellPoint geometry (mesh.cell_point_data restates what OpenFOAM derives)
            from .mesh import cell_point_data
            d = cell_point_data(self.mesh)
            cp = _capi.CellPoint()
            keep = [np.ascontiguousarray(self.mesh.points, np.float64), d["tetOffsets"], np.ascontiguousarray(d["tetPoints"]), d["pointCellOffsets"],
                    d["pointCells"], d["pointWeights"], d["pointNormals"]]
            PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)
            cp.nPoints = len(self.mesh.points)
            cp.points = keep[0].ctypes.data_as(PD)
            cp.tetOffsets, cp.tetPoints = keep[1].ctypes.data_as(PI), keep[2].ctypes.data_as(PI)
            cp.pointCellOffsets, cp.pointCells = keep[3].ctypes.data_as(PI), keep[4].ctypes.data_as(PI)
            cp.pointWeights, cp.pointNormals = keep[5].ctypes.data_as(PD), keep[6].ctypes.data_as(PD)
            self._check(self.api.set_macro_interpolation(self._h, C.byref(cp)))
        for entry in bd.get("uniGasGeneralBoundaries", []):
            word = entry["boundaryModel"]
            known = ("uniGasFreeStreamInflowPatch", "uniGasChapmanEnskogFreeStreamInflowPatch", "uniGasFreeStreamInflowFieldPatch",
                     "uniGasLiouFangPressureInletPatch", "uniGasWangPressureInletPatch", "uniGasLiouFangPressureOutletPatch",
                     "uniGasMassFlowRateInletPatch")
            if word not in known:
                raise UgfError(f"general boundary model {word!r} is not supported ({', '.join(known)})")
            patch = self.mesh.patch_index(entry["generalBoundaryProperties"]["patch"])
            if self.mesh.patches[patch].size == 0:
                continue  # a decomposed case: this rank holds no face of the patch
            pr = entry[word + "Properties"]
            if word == "uniGasFreeStreamInflowFieldPatch":
                # the reference reads the volFields boundaryNumberDensity_<species>, boundaryTransT, boundaryRotT, boundaryU from
                # the time directory (…/uniGasFreeStreamInflowFieldPatch.C:65-183); here their values on this patch come with
                # the dictionary entry (scalars / one vector are broadcast over the faces), as for the wall field patches
                nF = self.mesh.patches[patch].size
                ids = np.array([self.typeIdList.index(n) for n in pr["typeIds"]], np.int32)
                bn = self._f64(np.stack([np.broadcast_to(np.asarray(entry["boundaryNumberDensity"][self.typeIdList[t]], float), (nF,))
                                         for t in ids]))
                bT = self._f64(np.broadcast_to(np.asarray(entry["boundaryTransT"], float), (nF,)))
                bR = self._f64(np.broadcast_to(np.asarray(entry.get("boundaryRotT", 0.0), float), (nF,)))
                bU = self._f64(np.broadcast_to(np.asarray(entry["boundaryU"], float), (nF, 3)))
                PD = C.POINTER(C.c_double)
                self._check(self.api.set_inflow_fields(self._h, patch, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)), bn.ctypes.data_as(PD),
                                                       bT.ctypes.data_as(PD), bR.ctypes.data_as(PD), bU.ctypes.data_as(PD)))
                continue
            if word == "uniGasLiouFangPressureOutletPatch":  # …/uniGasLiouFangPressureOutletPatch.C:50-118
                pout = _capi.PressureInlet()
                ids = [self.typeIdList.index(n) for n in pr["typeIds"]]
                pout.nTypeIds = len(ids)
                for i, t in enumerate(ids):
                    pout.typeIds[i] = t
                    pout.moleFractions[i] = float(pr["moleFractions"][self.typeIdList[t]])
                pout.inletPressure = float(pr["outletPressure"])
                pout.inletTemperature = float(pr.get("initialOutletTemperature", 300.0))  # outletTemperature_ starts at 300 K (:74)
                pout.theta = 1.0
                self._check(self.api.set_pressure_outlet(self._h, patch, C.byref(pout)))
                continue
            if word == "uniGasMassFlowRateInletPatch":  # …/uniGasMassFlowRateInletPatch.C:53-125
                pin = _capi.PressureInlet()
                ids = [self.typeIdList.index(n) for n in pr["typeIds"]]
                pin.nTypeIds = len(ids)
                for i, t in enumerate(ids):
                    pin.typeIds[i] = t
                    pin.moleFractions[i] = float(pr["moleFractions"][self.typeIdList[t]])
                pin.inletTemperature = float(pr["inletTemperature"])
                pin.theta = float(pr.get("theta", 1.0))
                v0 = self._f64(pr.get("initialVelocity", [0.0, 0.0, 0.0])).reshape(3)
                self._check(self.api.set_mass_flow_inlet(self._h, patch, C.byref(pin), float(pr["massFlowRate"]), v0.ctypes.data_as(C.POINTER(C.c_double))))
                continue
            if word in ("uniGasLiouFangPressureInletPatch", "uniGasWangPressureInletPatch"):  # …/uniGasLiouFangPressureInletPatch.C:54-103
                pin = _capi.PressureInlet()
                ids = [self.typeIdList.index(n) for n in pr["typeIds"]]
                pin.nTypeIds = len(ids)
                for i, t in enumerate(ids):
                    pin.typeIds[i] = t
                    pin.moleFractions[i] = float(pr["moleFractions"][self.typeIdList[t]])
                pin.inletPressure = float(pr["inletPressure"])
                pin.inletTemperature = float(pr["inletTemperature"])
                pin.theta = float(pr.get("theta", 1.0))
                if word == "uniGasWangPressureInletPatch":  # …/uniGasWangPressureInletPatch.C:54-116 (no theta)
                    pin.theta = 1.0
                    self._check(self.api.set_wang_pressure_inlet(self._h, patch, C.byref(pin)))
                else:
                    self._check(self.api.set_pressure_inlet(self._h, patch, C.byref(pin)))
                continue
            inf = _capi.Inflow()
            ids = [self.typeIdList.index(n) for n in pr["typeIds"]]
            inf.nTypeIds = len(ids)
            for i, t in enumerate(ids):
                inf.typeIds[i] = t
                inf.numberDensities[i] = float(pr["numberDensities"][self.typeIdList[t]])
            inf.translationalTemperature = float(pr["translationalTemperature"])
            inf.rotationalTemperature = float(pr.get("rotationalTemperature", 0.0))
            inf.vibrationalTemperature = float(pr.get("vibrationalTemperature", 0.0))
            inf.electronicTemperature = float(pr.get("electronicTemperature", 0.0))
            for k in range(3):
                inf.velocity[k] = float(pr["velocity"][k])
            if word == "uniGasChapmanEnskogFreeStreamInflowPatch":  # …/uniGasChapmanEnskogFreeStreamInflowPatch.C:66-67
                q = self._f64(pr["heatFlux"]).reshape(3)
                st = self._f64(pr["stress"]).reshape(9)
                PD = C.POINTER(C.c_double)
                self._check(self.api.set_chapman_enskog_inflow(self._h, patch, C.byref(inf), q.ctypes.data_as(PD), st.ctypes.data_as(PD)))
                continue
            self._check(self.api.set_inflow(self._h, patch, C.byref(inf)))

    def close(self):
        if self._created:
            self.api.destroy(self._h)
            self._created = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state upload -------------------------------------------------------------
    @staticmethod
    def _f64(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    @staticmethod
    def _i32(a):
        return np.ascontiguousarray(a, dtype=np.int32)

    def setParcels(self, position, U, cell, typeId=None, ERot=None, newParcel=None, cellWeight=None, vibLevel=None, ELevel=None, radialWeight=None):
        """addNewParcel for a whole configuration (U/clouds/uniGasCloud.C:260-290)."""
        n = len(cell)
        if self._pending_capacity:
            self.cfg.parcelCapacity = max(int(n * 1.25) + 1024, 4096)
            self._pending_capacity = False
            self._create()
        position = np.asarray(position, dtype=np.float64)
        U = np.asarray(U, dtype=np.float64)
        cols = [self._f64(position[:, k]) for k in range(3)] + [self._f64(U[:, k]) for k in range(3)]
        cl = self._i32(cell)
        p = _capi.Parcels()
        p.n = n
        PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        p.x, p.y, p.z, p.Ux, p.Uy, p.Uz = [c.ctypes.data_as(PD) for c in cols]
        p.cell = cl.ctypes.data_as(PI)
        keep = [cols, cl]
        if typeId is not None:
            t = self._i32(typeId); keep.append(t); p.typeId = t.ctypes.data_as(PI)
        if ERot is not None:
            e = self._f64(ERot); keep.append(e); p.ERot = e.ctypes.data_as(PD)
        if newParcel is not None:
            q = self._i32(newParcel); keep.append(q); p.newParcel = q.ctypes.data_as(PI)
        if cellWeight is not None:
            w = self._f64(cellWeight); keep.append(w); p.cellWeight = w.ctypes.data_as(PD)
        if vibLevel is not None:  # [n, nModes <= UGF_MAX_VIB_MODES] quantum levels; padded to the ABI's row length
            v = np.zeros((n, _capi.UGF_MAX_VIB_MODES), np.int32)
            vl = np.asarray(vibLevel, np.int32).reshape(n, -1)
            v[:, :vl.shape[1]] = vl
            keep.append(v); p.vibLevel = v.ctypes.data_as(PI)
        if ELevel is not None:
            e = self._i32(ELevel); keep.append(e); p.ELevel = e.ctypes.data_as(PI)
        if radialWeight is not None:
            rw = self._f64(radialWeight); keep.append(rw); p.radialWeight = rw.ctypes.data_as(PD)
        self._check(self.api.upload_parcels(self._h, C.byref(p)))
        self._nParcelsSet = True
        self._cwfCarried = None

    def setCellState(self, sigmaTcRMax=None, cellCollModelId=None, subCellLevels=None, cellWeightFactor=None):
        PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        nC = self.mesh.n_cells
        a = self._f64(np.broadcast_to(sigmaTcRMax, (nC,))) if sigmaTcRMax is not None else None
        b = self._i32(np.broadcast_to(cellCollModelId, (nC,))) if cellCollModelId is not None else None
        c = self._i32(np.broadcast_to(subCellLevels, (nC, 3))) if subCellLevels is not None else None
        d = self._f64(np.broadcast_to(cellWeightFactor, (nC,))) if cellWeightFactor is not None else None
        if b is not None:
            self._cellCollModelId = b.astype(float)
        if c is not None:
            self._subCellLevels = c.astype(float)
        if d is not None:
            if self._created and self._nParcelsSet and self._cwfCarried is None:
                self._cwfCarried = self._cellWeightFactor.copy() if self._cellWeightFactor is not None else np.ones(nC)
            self._cellWeightFactor = d.copy()
        if d is not None and not self.cellWeighted:
            raise UgfError("cellWeightFactor given but cellWeightedSimulation is not true in uniGasProperties")
        self._check(self.api.upload_cell_state(
            self._h,
            a.ctypes.data_as(PD) if a is not None else None,
            b.ctypes.data_as(PI) if b is not None else None,
            c.ctypes.data_as(PI) if c is not None else None,
            d.ctypes.data_as(PD) if d is not None else None))

    def setHybridDecomposition(self, hybridDecompositionDict):
        """system/hybridDecompositionDict: `decompositionModel localKnudsen` with its timeProperties and
        localKnudsenProperties sub-dictionaries (same keys and defaults as the reference:
        uniGasHybridDecomposition.C:66-73, localKnudsen.C:52-55)."""
        d = hybridDecompositionDict
        _lookup({"localKnudsen": 1}, d.get("decompositionModel", "localKnudsen"), "decompositionModel")
        tp, kp = d.get("timeProperties", {}), d.get("localKnudsenProperties", {})
        dec = _capi.Decomposition()
        dec.decompositionInterval = int(tp.get("decompositionInterval", 100))
        dec.resetAtDecomposition = int(bool(tp.get("resetAtDecomposition", True)))
        dec.resetAtDecompositionUntilTime = float(tp.get("resetAtDecompositionUntilTime", 1e300))
        dec.breakdownMax = float(kp.get("breakdownMax", 0.05))
        dec.theta = float(kp.get("theta", 1.0))
        dec.smoothingPasses = int(kp.get("smoothingPasses", 0))
        dec.refinementPasses, dec.neighborLevels, dec.maxNeighborFraction = 3, 3, 0.4
        self._check(self.api.set_decomposition(self._h, C.byref(dec)))

    def decomposition(self):
        """uniGasCloud::decomposition(), for phase-wise drivers (evolve() calls it itself)."""
        self._check(self.api.decompose(self._h))

    def hybridDecomposition(self):
        """cellCollModelId (0 = bgk, 1 = dsmc) and the blended fields KnRho, KnT, KnU, KnGLL per cell."""
        nC = self.mesh.n_cells
        ids = np.empty(nC, np.int32)
        kn = np.empty((nC, 4))
        self._check(self.api.download_decomposition(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), kn.ctypes.data_as(C.POINTER(C.c_double))))
        return dict(cellCollModelId=ids, KnRho=kn[:, 0], KnT=kn[:, 1], KnU=kn[:, 2], KnGLL=kn[:, 3])

    # -- write / restart (OpenFOAM time directories, unigasfoam_b200/foamfile.py) ----------------------
    def state(self):
        """ugf_state_save: everything carried between steps besides parcels and cell-state fields, as a float64 array."""
        n = C.c_int64()
        self._check(self.api.state_size(self._h, C.byref(n)))
        buf = np.empty(n.value)
        self._check(self.api.state_save(self._h, buf.ctypes.data_as(C.POINTER(C.c_double)), n.value))
        return buf

    def loadState(self, buf):
        buf = self._f64(buf)
        self._check(self.api.state_load(self._h, buf.ctypes.data_as(C.POINTER(C.c_double)), len(buf)))

    def writeTime(self, case_dir, time_name, fieldNames=()):
        """What uniGasFoam leaves in <time>/ for the cloud: lagrangian/uniGas/*, uniGas{SigmaTcRMax, CellWeightFactor,
        SubCellLevels, CollisionModelId} and uniform/time (U/parcels/uniGasParcelIO.C:141-181, U/clouds/uniGasCloud.C:433-488);
        for every name in fieldNames (the `field` words of the uniGasVolFields entries of fieldPropertiesDict) also
        uniform/volFieldsMethod_<name> (uniGasVolFields.C:609-669)."""
        from . import foamfile
        p = self.parcels()
        st = self.cellState()
        mode = self.cfg.collisionModel
        ids = self._cellCollModelId
        if ids is None:
            ids = np.full(self.mesh.n_cells, 1.0 if mode == _capi.COLLISION_MODEL["dsmc"] else 0.0)
        c = self.counters()
        t = foamfile.write_cloud_time(case_dir, time_name, self.mesh, p, st["sigmaTcRMax"], self._cellWeightFactor, self._subCellLevels,
                                      ids, deltaT=self.cfg.deltaT, index=c["step"])
        if self._cwfCarried is not None:  # a factor field uploaded since the last step: the parcels still carry the previous one
            foamfile.write_vol_field(os.path.join(t, "uniGasCellWeightFactorCarried"), time_name, [0] * 7, self._cwfCarried,
                                     [q.name for q in self.mesh.patches])
        st = self.state()
        np.save(os.path.join(t, "uniform", "ugfState.npy"), st)  # accumulators, BGK / decomposition / inlet state
        if fieldNames:
            from . import volfields_io
            for name in fieldNames:
                volfields_io.write_volfields_method(case_dir, time_name, name, self.mesh, st)
        return t

    def readTime(self, case_dir, time_name):
        """Restart from a time directory written by writeTime or by the reference solver (ASCII): cell state first, then
        the parcels, then the step count."""
        from . import foamfile
        d = foamfile.read_cloud_time(case_dir, time_name, self.mesh.n_cells)
        p = d["parcels"]
        if not self.axisymmetric and (p["radialWeight"] != 1.0).any():
            raise UgfError("radialWeight != 1 in the parcel files but axisymmetricSimulation is not true in uniGasProperties")
        vib = None
        if any(len(v) for v in p["vibLevel"]):
            nm = max(len(v) for v in p["vibLevel"])
            vib = np.zeros((len(p["cell"]), nm), np.int32)
            for i, v in enumerate(p["vibLevel"]):
                vib[i, :len(v)] = v
        kw = {}
        if d["sigmaTcRMax"] is not None:
            kw["sigmaTcRMax"] = d["sigmaTcRMax"]
        if d["collisionModelId"] is not None and self.cfg.collisionModel == _capi.COLLISION_MODEL["hybrid"]:
            kw["cellCollModelId"] = np.rint(d["collisionModelId"]).astype(np.int32)
        if d["subCellLevels"] is not None and (d["subCellLevels"] != 1).any():
            kw["subCellLevels"] = np.rint(d["subCellLevels"]).astype(np.int32)
        if self.cellWeighted and d["cellWeightFactor"] is not None:
            kw["cellWeightFactor"] = d["cellWeightFactor"]
        if self._pending_capacity:
            self.cfg.parcelCapacity = max(int(len(p["cell"]) * 1.25) + 1024, 4096)
            self._pending_capacity = False
            self._create()
        carried = os.path.join(case_dir, time_name, "uniGasCellWeightFactorCarried")
        new_cwf = None
        if self.cellWeighted and os.path.exists(carried):  # restore "parcels carry the previous field, the new one is pending"
            new_cwf = kw.get("cellWeightFactor")
            kw["cellWeightFactor"] = foamfile.expand_internal(foamfile.read_vol_field(carried), self.mesh.n_cells)
        if kw:
            self.setCellState(**kw)
        self.setParcels(p["position"], p["U"], p["cell"], p["typeId"], p["ERot"], cellWeight=p["cellWeight"] if self.cellWeighted else None,
                        vibLevel=vib, ELevel=p["ELevel"] if (p["ELevel"] != 0).any() else None,
                        radialWeight=p["radialWeight"] if self.axisymmetric else None)
        if new_cwf is not None:
            self.setCellState(cellWeightFactor=new_cwf)
        if d.get("deltaT") is not None:
            self.setDeltaT(d["deltaT"])
        self._check(self.api.set_time_index(self._h, int(d.get("index", 0))))
        sp = os.path.join(case_dir, time_name, "uniform", "ugfState.npy")
        if os.path.exists(sp):  # written by writeTime: makes the restart exact for BGK / hybrid runs and the field averages
            self.loadState(np.load(sp))
        return d

    # name, key of fields() (None: zero field), dimensions [kg m s K mol A cd], boundary rule, wall key, vector
    # (uniGasVolFields.C:67-357).  Boundary rules follow :1256-1394: "cell" - every wall / generic patch face takes the value
    # of its cell; "wall" - wall faces take the wall measurement, generic patch faces the cell value; "surface" - wall faces
    # take the wall measurement, everything else is zero; "zero".  Constraint patches (empty, cyclic, symmetry, processor) are
    # written by their type.  Wall rotationalT / overallT (:1299-1330) are derived in writeFields from the wall accumulators.
    _OUTPUT_FIELDS = (
        ("uniGasRhoNMean", "uniGasRhoNMean", [0, -3, 0, 0, 0, 0, 0], "cell", None, False),
        ("rhoN", "rhoN", [0, -3, 0, 0, 0, 0, 0], "cell", None, False),
        ("rhoM", "rhoM", [1, -3, 0, 0, 0, 0, 0], "cell", None, False),
        ("p", "p", [1, -1, -2, 0, 0, 0, 0], "wall", "wall_p", False),
        ("translationalT", "translationalT", [0, 0, 0, 1, 0, 0, 0], "wall", "wall_translationalT", False),
        ("rotationalT", "rotationalT", [0, 0, 0, 1, 0, 0, 0], "wall", "wall_rotationalT", False),
        ("vibrationalT", "vibrationalT", [0, 0, 0, 1, 0, 0, 0], "cell", None, False),
        ("electronicT", "electronicT", [0, 0, 0, 1, 0, 0, 0], "cell", None, False),
        ("overallT", "overallT", [0, 0, 0, 1, 0, 0, 0], "wall", "wall_overallT", False),
        ("surfaceHeatTransfer", None, [1, 0, -3, 0, 0, 0, 0], "surface", "surfaceHeatTransfer", False),
        ("surfaceShearStress", None, [1, -1, -2, 0, 0, 0, 0], "surface", "surfaceShearStress", False),
        ("Ma", "Ma", [0] * 7, "cell", None, False),
        ("UMean", "UMean", [0, 1, -1, 0, 0, 0, 0], "wall", "wall_UMean", True),
        ("fD", None, [1, -1, -2, 0, 0, 0, 0], "surface", "fD", True),
    )
    _MFP_FIELDS = (("variableHardSphereMeanFreePath", "MFP", [0, 1, 0, 0, 0, 0, 0]), ("subCellSizeMFPRatio", "dxMFP", [0] * 7),
                   ("meanCollisionRate", "MCR", [0, 0, -1, 0, 0, 0, 0]), ("meanCollisionTime", "MCT", [0, 0, 1, 0, 0, 0, 0]),
                   ("timeStepMCTRatio", "dtMCT", [0] * 7))
    _ERROR_FIELDS = (("densityError", "densityError", [0] * 7), ("velocityError", "velocityError", [0] * 7),
                     ("temperatureError", "temperatureError", [0] * 7), ("pressureError", "pressureError", [0] * 7))

    def writeFields(self, case_dir, time_name, fieldName, resetAtOutput=False, measureMeanFreePath=True, measureErrors=True):
        """The output volFields of a uniGasVolFields entry at write time, under the reference's names and dimensions
        (`rhoN_<field>`, `p_<field>`, `UMean_<field>`, `surfaceHeatTransfer_<field>` ..., uniGasVolFields.C:67-357,
        1397-1428): cell values as internalField, boundary values by the rules above.  -> the list of files written."""
        from . import foamfile, volfields_io
        bacc = volfields_io.state_views(self.state())["bacc"].copy()  # before a reset: rotationalEBF 6, rotationalDofBF 7, rhoNBF 0
        f = self.fields(resetAtOutput=resetAtOutput)
        with np.errstate(divide="ignore", invalid="ignore"):
            f["wall_rotationalT"] = np.where(bacc[:, 7] > 1e-300, (2.0 / 1.38065e-23) * bacc[:, 6] / bacc[:, 7], 0.0)       # :1299-1308
            nRot = np.where(bacc[:, 0] > 1e-300, bacc[:, 7] / bacc[:, 0], 0.0)
            f["wall_overallT"] = (3.0 * f["wall_translationalT"] + nRot * f["wall_rotationalT"]) / (3.0 + nRot)             # :1318-1330
        m = self.mesh
        nI = m.n_internal
        t = os.path.join(case_dir, time_name)
        os.makedirs(t, exist_ok=True)
        todo = list(self._OUTPUT_FIELDS)
        if measureMeanFreePath:
            todo += [(n, k, d, "cell", None, False) for n, k, d in self._MFP_FIELDS]
        if measureErrors:
            todo += [(n, k, d, "zero", None, False) for n, k, d in self._ERROR_FIELDS]
        out = []
        for name, key, dims, rule, wkey, vector in todo:
            internal = f[key] if key is not None else (np.zeros((m.n_cells, 3)) if vector else np.zeros(m.n_cells))
            patches = {}
            for p in m.patches:
                if p.kind not in ("wall", "patch"):
                    patches[p.name] = p.kind
                    continue
                own = np.asarray(m.owner[p.start:p.start + p.size])
                if p.kind == "wall" and rule in ("wall", "surface"):
                    val = f[wkey][p.start - nI:p.start - nI + p.size]
                elif rule in ("cell", "wall"):
                    val = internal[own]
                else:
                    val = np.zeros(3) if vector else 0.0
                patches[p.name] = {"type": "calculated", "value": val}
            path = os.path.join(t, f"{name}_{fieldName}")
            foamfile.write_vol_field(path, time_name, dims, internal, patches, vector=vector)
            out.append(path)
        return out

    def writeVolFieldsMethod(self, case_dir, time_name, fieldName):
        """uniGasVolFields::writeOut (uniGasVolFields.C:609-669): the accumulators of the field object `fieldName` as
        <time>/uniform/volFieldsMethod_<fieldName>, under the reference's entry names (unigasfoam_b200/volfields_io.py)."""
        from . import volfields_io
        return volfields_io.write_volfields_method(case_dir, time_name, fieldName, self.mesh, self.state())

    def readVolFieldsMethod(self, case_dir, time_name, fieldName):
        """uniGasVolFields::readIn (uniGasVolFields.C:549-605, `averagingAcrossManyRuns`): continue the time averages of a
        dictionary written by writeVolFieldsMethod or by the reference solver on the same mesh.  -> the entries read."""
        from . import volfields_io
        path = os.path.join(case_dir, time_name, "uniform", "volFieldsMethod_" + fieldName)
        if not os.path.exists(path):  # READ_IF_PRESENT
            return None
        d = volfields_io.read_volfields_method(path)
        self.loadState(volfields_io.apply_volfields_method(d, self.mesh, self.state()))
        return d

    def setDeltaT(self, dt):
        self._check(self.api.set_deltaT(self._h, float(dt)))
        self.cfg.deltaT = float(dt)

    # -- the loop -------------------------------------------------------------------
    def evolve(self, nSteps=1):
        """uniGasCloud::evolve (U/clouds/uniGasCloud.C:821-869)."""
        self._check(self.api.step(self._h, int(nSteps)))
        if nSteps > 0:
            self._cwfCarried = None

    def controlBeforeMove(self):
        self._check(self.api.control_before_move(self._h))

    def move(self):
        self._check(self.api.move(self._h))
        self._cwfCarried = None  # the weighting pass rides on the move

    def buildCellOccupancy(self):
        self._check(self.api.sort(self._h))

    def reorder(self):
        self._check(self.api.reorder(self._h))

    def calculateFields(self):
        """cellMeas_.calculateFields() (U/cellMeasurements/cellMeasurements.C:408-513)."""
        self._check(self.api.sample(self._h))

    def collide(self):
        self._check(self.api.collide(self._h))

    def relax(self):
        self._check(self.api.relax(self._h))

    def accumulateFields(self):
        self._check(self.api.accumulate_fields(self._h))

    def endStep(self):
        self._check(self.api.end_step(self._h))

    def finishStep(self):
        """buildCellOccupancy ... cellMeas_.clean() of evolve() in one call (U/clouds/uniGasCloud.C:839-866)."""
        self._check(self.api.finish_step(self._h))

    # -- migration --------------------------------------------------------------------
    def migrateCounts(self):
        out = (C.c_int64 * len(self.mesh.patches))()
        self._check(self.api.migrate_counts(self._h, out))
        return np.array(out[:], dtype=np.int64)

    def migratePack(self, patch):
        buf = C.POINTER(C.c_double)()
        n = C.c_int64()
        self._check(self.api.migrate_pack(self._h, patch, C.byref(buf), C.byref(n)))
        return buf, n.value

    def migrateUnpack(self, patch, buf_ptr, n):
        self._check(self.api.migrate_unpack(self._h, patch, buf_ptr, int(n)))

    def migratePackSlots(self, send_ptr, slot_capacity):
        self._check(self.api.migrate_pack_slots(self._h, send_ptr, int(slot_capacity)))

    def migrateUnpackSlots(self, recv_ptr, slot_capacity):
        self._check(self.api.migrate_unpack_slots(self._h, recv_ptr, int(slot_capacity)))

    # NVLink peer-memory transfer (ugf_peer_* / ugf_migrate_*_peer): raw device addresses in, nothing copied by the host
    def peerAlloc(self, nbytes):
        """Receive region in this GPU's memory; returns (device address, 64-byte cudaIpc handle)."""
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self.api.peer_alloc(self._h, int(nbytes), C.byref(ptr), handle))
        return ptr.value, handle.raw

    def peerOpen(self, handle):
        ptr = C.c_void_p()
        self._check(self.api.peer_open(self._h, C.create_string_buffer(bytes(handle), 64), C.byref(ptr)))
        return ptr.value

    def migratePackPeer(self, dst_slots, dst_flags, slot_capacity, epoch):
        n = len(dst_slots)
        a = (C.c_void_p * max(n, 1))(*dst_slots)
        b = (C.c_void_p * max(n, 1))(*dst_flags)
        self._check(self.api.migrate_pack_peer(self._h, a, b, int(slot_capacity), int(epoch)))

    def migrateUnpackPeer(self, recv_addr, flags_addr, slot_capacity, epoch):
        self._check(self.api.migrate_unpack_peer(self._h, C.c_void_p(recv_addr), C.c_void_p(flags_addr), int(slot_capacity), int(epoch)))

    def migrateInflightPtr(self):
        p = C.POINTER(C.c_int64)()
        self._check(self.api.migrate_inflight(self._h, C.byref(p)))
        return C.cast(p, C.c_void_p).value

    def moveReceived(self):
        self._check(self.api.move_received(self._h))

    def stream(self):
        s = C.c_void_p()
        self._check(self.api.stream(self._h, C.byref(s)))
        return s.value

    # -- results ------------------------------------------------------------------------
    def counters(self):
        c = _capi.Counters()
        self._check(self.api.counters_get(self._h, C.byref(c)))
        return c.as_dict()

    def info(self):
        """uniGasCloud::info (U/clouds/uniGasCloud.C:878-920) as a dict."""
        c = self.counters()
        n = c["nParcels"]
        nMol = n * self.cfg.nParticle
        out = {"nParticles": n, "deltaT": self.cfg.deltaT}
        if n:
            out["avgLinearKE"] = c["linearKineticEnergy"] * self.cfg.nParticle / nMol
            out["avgRotationalE"] = c["rotationalEnergy"] * self.cfg.nParticle / nMol
            out["avgVibrationalE"] = c["vibrationalEnergy"] * self.cfg.nParticle / nMol
            out["avgElectronicE"] = c["electronicEnergy"] * self.cfg.nParticle / nMol
            out["totalEnergy"] = (c["linearKineticEnergy"] + c["rotationalEnergy"] + c["vibrationalEnergy"] + c["electronicEnergy"]) * self.cfg.nParticle
        return out

    def size(self):
        n = C.c_int64()
        self._check(self.api.num_parcels(self._h, C.byref(n)))
        return n.value

    def parcels(self):
        cap = int(self.cfg.parcelCapacity)
        PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        f = [np.empty(cap, np.float64) for _ in range(9)]
        ii = [np.empty(cap, np.int32) for _ in range(3)]
        vib = np.empty((cap, _capi.UGF_MAX_VIB_MODES), np.int32)
        p = _capi.Parcels()
        p.n = cap
        p.x, p.y, p.z, p.Ux, p.Uy, p.Uz, p.ERot, p.cellWeight, p.radialWeight = [a.ctypes.data_as(PD) for a in f]
        p.cell, p.typeId, p.ELevel = [a.ctypes.data_as(PI) for a in ii]
        p.vibLevel = vib.ctypes.data_as(PI)
        self._check(self.api.download_parcels(self._h, C.byref(p)))
        n = p.n
        return {
            "ELevel": ii[2][:n].copy(), "vibLevel": vib[:n].copy(),
            "position": np.stack([f[0][:n], f[1][:n], f[2][:n]], axis=1),
            "U": np.stack([f[3][:n], f[4][:n], f[5][:n]], axis=1),
            "ERot": f[6][:n].copy(), "cell": ii[0][:n].copy(), "typeId": ii[1][:n].copy(), "cellWeight": f[7][:n].copy(),
            "radialWeight": f[8][:n].copy(),
        }

    def cellOccupancy(self):
        """(offsets [nCells+1], ids [n]) - CloudWithModels::cellOccupancy as CSR."""
        nC = self.mesh.n_cells
        off = np.empty(nC + 1, np.int32)
        PI = C.POINTER(C.c_int32)
        self._check(self.api.download_cell_occupancy(self._h, off.ctypes.data_as(PI), None))
        ids = np.empty(int(off[-1]), np.int32)
        self._check(self.api.download_cell_occupancy(self._h, off.ctypes.data_as(PI), ids.ctypes.data_as(PI)))
        return off, ids

    def cellMoments(self):
        a = np.empty((self.mesh.n_cells, len(self.typeIdList), _capi.UGF_NMOM), np.float64)
        self._check(self.api.download_cell_moments(self._h, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a

    def cellState(self):
        nC = self.mesh.n_cells
        PD = C.POINTER(C.c_double)
        s, mp, q, sp = np.empty(nC), np.empty(nC), np.empty((nC, 3)), np.empty((nC, 6))
        self._check(self.api.download_cell_state(self._h, s.ctypes.data_as(PD), mp.ctypes.data_as(PD),
                                                 q.ctypes.data_as(PD), sp.ctypes.data_as(PD)))
        return {"sigmaTcRMax": s, "maxProb": mp, "qPrev": q, "sPrev": sp}

    def fields(self, resetAtOutput=False):
        """uniGasVolFields at write time: dict of per-cell and per-boundary-face arrays."""
        nC, nB = self.mesh.n_cells, self.mesh.n_boundary_faces
        PD = C.POINTER(C.c_double)
        if resetAtOutput and self._adapter is not None:
            self._adapter.before_reset()  # its interval sums are differences of the accumulators about to be zeroed
        cf = np.empty((nC, _capi.UGF_NFIELD))
        wf = np.empty((max(nB, 1), _capi.UGF_NWALLFIELD))
        self._check(self.api.download_fields(self._h, cf.ctypes.data_as(PD), wf.ctypes.data_as(PD), int(resetAtOutput)))
        wf = wf[:nB]
        # pressureError = sqrt(gamma) densityError (uniGasVolFields.C:1250) with gamma recovered from velocityError =
        # densityError / (Ma sqrt(gamma)) (:1248): no extra device field
        with np.errstate(divide="ignore", invalid="ignore"):
            pErr = np.where((cf[:, 17] > 0) & (cf[:, 10] > 0), cf[:, 11] * cf[:, 11] / (cf[:, 17] * cf[:, 10]), 0.0)
        return {
            "pressureError": pErr,
            "uniGasRhoNMean": cf[:, 0], "rhoN": cf[:, 1], "rhoM": cf[:, 2], "UMean": cf[:, 3:6],
            "translationalT": cf[:, 6], "rotationalT": cf[:, 7], "overallT": cf[:, 8], "p": cf[:, 9],
            "Ma": cf[:, 10], "densityError": cf[:, 11],
            "MFP": cf[:, 12], "dxMFP": cf[:, 13], "MCR": cf[:, 14], "MCT": cf[:, 15], "dtMCT": cf[:, 16],
            "velocityError": cf[:, 17], "temperatureError": cf[:, 18], "vibrationalT": cf[:, 19], "electronicT": cf[:, 20],
            "wall_rhoN": wf[:, 0], "wall_rhoM": wf[:, 1], "wall_UMean": wf[:, 2:5], "wall_translationalT": wf[:, 5],
            "surfaceHeatTransfer": wf[:, 6], "fD": wf[:, 7:10], "wall_p": wf[:, 10], "surfaceShearStress": wf[:, 11],
        }

    def accumulators(self):
        """Raw time-weighted sums of uniGasVolFields (see ugf_download_accumulators): dict acc [nCells,16],
        species [nCells,nSpecies], timeAvCounter, nAvTimeSteps."""
        nC, nS = self.mesh.n_cells, len(self.typeIdList)
        PD = C.POINTER(C.c_double)
        acc, sp = np.empty((nC, 16)), np.empty((nC, nS))
        t, n = C.c_double(), C.c_int64()
        self._check(self.api.download_accumulators(self._h, acc.ctypes.data_as(PD), sp.ctypes.data_as(PD), C.byref(t), C.byref(n)))
        return {"acc": acc, "species": sp, "timeAvCounter": t.value, "nAvTimeSteps": n.value}

    def internalAccumulators(self):
        """[nCells, nSpecies, UGF_NINT] time-weighted sums behind vibrationalT / electronicT (ugf_download_internal_accumulators)."""
        a = np.empty((self.mesh.n_cells, len(self.typeIdList), _capi.UGF_NINT))
        self._check(self.api.download_internal_accumulators(self._h, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a

    def setFaceTracker(self, faces):
        """Faces whose crossings uniGasFaceTracker tallies (the union of the face zones the surface models read)."""
        f = self._i32(faces)
        self._trackedFaces = f.copy()
        self._check(self.api.set_face_tracker(self._h, len(f), f.ctypes.data_as(C.POINTER(C.c_int32)) if len(f) else None))

    def faceTracker(self, reset=False):
        """[nTracked, nSpecies, 6]: parcels, mass, momentum (3), energy carried through each tracked face since the last reset."""
        n, nS = len(self._trackedFaces), len(self.typeIdList)
        out = np.empty((n, nS, _capi.UGF_NFT))
        self._check(self.api.download_face_tracker(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), int(reset)))
        return out

    def inletVelocity(self, patch_name):
        """Inflow velocity per face of a uniGasLiouFangPressureInletPatch."""
        p = self.mesh.patch_index(patch_name)
        U = np.empty((self.mesh.patches[p].size, 3))
        self._check(self.api.download_inlet_velocity(self._h, p, U.ctypes.data_as(C.POINTER(C.c_double))))
        return U

    def boundaryMeasurements(self):
        nB = self.mesh.n_boundary_faces
        a = np.empty((max(nB, 1), _capi.UGF_NBM))
        self._check(self.api.download_boundary_meas(self._h, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a[:nB]

    def phaseTimes(self):
        a = (C.c_double * _capi.UGF_NPHASE)()
        self._check(self.api.phase_times(self._h, a))
        return dict(zip(("inflow", "move", "sort", "cell", "collide", "relax", "fields"), a[:]))

    def transferBytes(self):
        """(host -> device, device -> host, kernel-argument) bytes moved by this cloud since construction (ugf_transfer_bytes)."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.api.transfer_bytes(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def hostArray(self, n, dtype=np.float64):
        """A page-locked numpy array of n items owned by the library (ugf_host_alloc): parcels handed over in such arrays
        cross PCIe as asynchronous DMA.  Valid until the cloud is closed."""
        dt = np.dtype(dtype)
        ptr = C.c_void_p()
        self._check(self.api.host_alloc(self._h, int(n) * dt.itemsize, C.byref(ptr)))
        buf = (C.c_byte * (int(n) * dt.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dt, count=int(n))

    def setParcelsSoA(self, n, x, y, z, Ux, Uy, Uz, cell, typeId=None, ERot=None):
        """setParcels from structure-of-arrays columns that are already contiguous float64 / int32 (e.g. hostArray buffers):
        no host-side copies."""
        PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        p = _capi.Parcels()
        p.n = int(n)
        p.x, p.y, p.z, p.Ux, p.Uy, p.Uz = [a.ctypes.data_as(PD) for a in (x, y, z, Ux, Uy, Uz)]
        p.cell = cell.ctypes.data_as(PI)
        if typeId is not None:
            p.typeId = typeId.ctypes.data_as(PI)
        if ERot is not None:
            p.ERot = ERot.ctypes.data_as(PD)
        self._check(self.api.upload_parcels(self._h, C.byref(p)))
        self._nParcelsSet = True
        self._cwfCarried = None

    def parcelsInto(self, x, y, z, Ux, Uy, Uz, cell, ERot=None, typeId=None):
        """Download the parcels into caller-owned columns (capacity = len(cell)); returns the count."""
        PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        p = _capi.Parcels()
        p.n = len(cell)
        p.x, p.y, p.z, p.Ux, p.Uy, p.Uz = [a.ctypes.data_as(PD) for a in (x, y, z, Ux, Uy, Uz)]
        p.cell = cell.ctypes.data_as(PI)
        if ERot is not None:
            p.ERot = ERot.ctypes.data_as(PD)
        if typeId is not None:
            p.typeId = typeId.ctypes.data_as(PI)
        self._check(self.api.download_parcels(self._h, C.byref(p)))
        return int(p.n)

    def launchCount(self):
        n = C.c_int64()
        self._check(self.api.launch_count(self._h, C.byref(n)))
        return n.value
