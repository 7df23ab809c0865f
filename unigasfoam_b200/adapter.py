"""uniGasDynamicAdapter - dynamic adaptation of the time step, the sub-cell levels and the cell weight factors
(U/dynamicAdaptation/uniGasDynamicAdapter.C:228-706), host side.

The adapter is not part of the per-step particle loop: every `adaptationInterval` steps it turns time-averaged cell
sums into three pieces of cell state and hands them back to the cloud.  Here the sums are the ones the device already
keeps for uniGasVolFields (`ugf_download_accumulators`; two downloads are differenced, so no second set of
accumulators is carried through the hot kernels), the arithmetic is numpy on the host mesh, and the results go back
through the same ABI calls the OpenFOAM shim uses: `ugf_set_deltaT`, `ugf_upload_cell_state(subCellLevels,
cellWeightFactor)`.  A rewritten factor field is applied by the next step's weighting pass (include/ugf.h).

Restated from the reference:
  adapt()                         :503-706  accumulate -> rhoN, T, U -> per-cell ratios -> smooth -> dt, levels, weights
  calculateAdaptationQuantities   :228-333  Bird 4.74 / 4.76 / 4.77 / 1.38 collision rate and mean free path
  calculateTimeStep               :335-388
  calculateSubCellLevels          :390-431
  calculateCellWeightFactor       :433-446
  smoothCellWeightFactor          :448-500
OpenFOAM pieces restated from their definition (not in /root/reference): fvc::average(fvc::interpolate(f)) = face-area
weighted mean of the linearly interpolated face values (the same operator as the localKnudsen smoothing, oracle
smoothFields); fvc::smooth(f, r) = the least field >= f in which no cell is more than a factor r below a neighbour
(FaceCellWave / smoothData propagate exactly that bound).  Requires sampleInterval 1 (the accumulators then advance
every step, as the adapter's own sums do in the reference).
"""
import math

import numpy as np

from ._capi import COLLISION_MODEL

kB = 1.38065e-23
VSMALL, SMALL, GREAT, VGREAT = 1e-300, 1e-15, 1e15, 1e300


class FaceOperators:
    """fvc::average(fvc::interpolate(.)) and fvc::smooth on a PolyMesh (zero-gradient / symmetry / cyclic boundary faces,
    empty faces take no part).  Processor faces are coupled like in OpenFOAM when a `halo` is given - a callable that takes
    this rank's values on its processor faces ([nProcFaces, k], patch order) and returns the neighbour ranks' values on the
    same faces (`exchange.ProcessorHalo`); `reduce_max` then makes the smoothing wave stop on all ranks together.  Without
    a halo processor faces are zero-gradient."""

    def __init__(self, mesh, halo=None, reduce_max=None):
        m = mesh
        self.halo = halo
        self.reduce_max = reduce_max or (lambda v: v)
        nI = m.n_internal
        self.nC, self.nI = m.n_cells, nI
        self.own, self.nei = np.asarray(m.owner[:nI]), np.asarray(m.neighbour[:nI])
        S, Cf = m.face_areas, m.face_centres
        self.A = np.sqrt((S * S).sum(1))
        dO = Cf[:nI] - m.cell_centres[self.own]
        dN = m.cell_centres[self.nei] - Cf[:nI]
        sO, sN = np.abs((S[:nI] * dO).sum(1)), np.abs((S[:nI] * dN).sum(1))
        self.w = sN / (sO + sN)  # surfaceInterpolation::makeWeights
        # boundary faces that take part: (face, owner, partner owner or -1, weight of the owner side, unit normal, symmetry?)
        bf, bo, bq, bw, sym, proc = [], [], [], [], [], []
        for p in m.patches:
            if p.kind == "empty" or p.size == 0:
                continue
            f = np.arange(p.start, p.start + p.size)
            o = np.asarray(m.owner[f])
            q = np.full(p.size, -1)
            w = np.ones(p.size)
            if p.kind == "cyclic":
                pp = m.patches[p.partner]
                nf = np.arange(pp.start, pp.start + pp.size)
                q = np.asarray(m.owner[nf])
                di = ((Cf[f] - m.cell_centres[o]) * S[f]).sum(1) / self.A[f]
                dni = ((Cf[nf] - m.cell_centres[q]) * S[nf]).sum(1) / self.A[nf]
                w = dni / (di + dni)
            bf.append(f); bo.append(o); bq.append(q); bw.append(w)
            sym.append(np.full(p.size, p.kind in ("symmetry", "symmetryPlane")))
            proc.append(np.full(p.size, p.kind == "processor"))
        cat = lambda L, dt: np.concatenate(L).astype(dt) if L else np.empty(0, dt)
        self.bf, self.bo, self.bq, self.bw, self.bsym = cat(bf, int), cat(bo, int), cat(bq, int), cat(bw, float), cat(sym, bool)
        self.bproc = cat(proc, bool) if halo is not None else np.zeros(len(self.bf), bool)
        self.po = self.bo[self.bproc]  # owner cells of the processor faces, in the order the halo exchanges them
        if halo is not None:
            # weights of a coupled face (processorFvPatch::makeWeights): the neighbour's own normal distance is my dni
            f = self.bf[self.bproc]
            di = ((Cf[f] - m.cell_centres[self.po]) * S[f]).sum(1) / self.A[f]
            dni = halo(di[:, None])[:, 0]
            self.bw[self.bproc] = dni / (di + dni)
        # unit normals of the boundary faces; a face collapsed onto an axis has no area and no normal (its weight below is zero too)
        self.bn = S[self.bf] / np.where(self.A[self.bf] > 0, self.A[self.bf], 1.0)[:, None] if len(self.bf) else np.empty((0, 3))
        den = np.zeros(self.nC)
        np.add.at(den, self.own, self.A[:nI]); np.add.at(den, self.nei, self.A[:nI]); np.add.at(den, self.bo, self.A[self.bf])
        self.den = den

    def average_interpolate(self, f, vector=False):
        f = np.asarray(f, float)
        two = f.ndim == 2
        F = f if two else f[:, None]
        wv = self.w[:, None]
        face = wv * F[self.own] + (1.0 - wv) * F[self.nei]
        num = np.zeros_like(F)
        Ai = self.A[:self.nI, None]
        np.add.at(num, self.own, Ai * face)
        np.add.at(num, self.nei, Ai * face)
        if len(self.bf):
            val = F[self.bo].copy()
            cyc = self.bq >= 0
            if cyc.any():
                val[cyc] = self.bw[cyc, None] * F[self.bo[cyc]] + (1.0 - self.bw[cyc, None]) * F[self.bq[cyc]]
            if self.halo is not None:  # also with no processor faces of its own a rank takes part in the exchange
                pr = self.bproc
                val[pr] = self.bw[pr, None] * F[self.po] + (1.0 - self.bw[pr, None]) * self.halo(F[self.po])
            if vector and self.bsym.any():  # symmetry planes mirror a vector: the face value has no normal component
                s = self.bsym
                vn = (F[self.bo[s]] * self.bn[s]).sum(1)
                val[s] = F[self.bo[s]] - vn[:, None] * self.bn[s]
            np.add.at(num, self.bo, self.A[self.bf][:, None] * val)
        out = num / self.den[:, None]
        return out if two else out[:, 0]

    def smooth(self, f, ratio):
        """fvc::smooth(f, ratio): raise values until every cell is within a factor `ratio` of each neighbour across an
        internal, cyclic or (with a halo) processor face - the least such field above f (what the smoothData wave
        converges to; FaceCellWave carries it through cyclic and processor patches)."""
        v = np.array(f, float)
        cyc = self.bq >= 0
        for _ in range(10 * (self.nC + 1) if self.halo is None else 10 ** 9):
            lo_o = v[self.nei] / ratio
            lo_n = v[self.own] / ratio
            new = v.copy()
            np.maximum.at(new, self.own, lo_o)
            np.maximum.at(new, self.nei, lo_n)
            if cyc.any():
                np.maximum.at(new, self.bo[cyc], v[self.bq[cyc]] / ratio)
            if self.halo is not None:  # the wave crosses processor faces
                np.maximum.at(new, self.po, self.halo(v[self.po][:, None])[:, 0] / ratio)
            if not self.reduce_max(0.0 if np.array_equal(new, v) else 1.0):
                break
            v = new
        return v

    def max_neighbour_ratio(self, f):
        a, b = f[self.own], f[self.nei]
        r = np.maximum(a / b, b / a)
        return float(r.max()) if len(r) else VSMALL


class UniGasDynamicAdapter:
    def __init__(self, cloud, uniGasProperties, reduce_min=None, reduce_max=None, halo=None):
        """reduce_min / reduce_max: callables float -> float that reduce over the ranks of a decomposed run (the reference's
        reduce(deltaT, minOp) at :382-385 and reduce(maxCellWeightRatio, maxOp) at :493-496); identity on one rank.
        halo: exchange of cell values across processor faces (FaceOperators; `exchange.ProcessorHalo`) - with it the
        smoothing of the ratios and of the cell weight factors sees the neighbour subdomains, as the reference's does;
        without it processor faces are zero-gradient."""
        self.reduce_min = reduce_min or (lambda v: v)
        self.reduce_max = reduce_max or (lambda v: v)
        self.cloud, self.mesh = cloud, cloud.mesh
        props = uniGasProperties
        ap = props.get("adaptiveProperties", {})
        self.timeStepAdaptation = bool(ap.get("timeStepAdaptation", False))      # :199-213
        self.subCellAdaptation = bool(ap.get("subCellAdaptation", False))
        self.cellWeightAdaptation = bool(ap.get("cellWeightAdaptation", False))
        self.adaptationInterval = int(ap.get("adaptationInterval", 20))
        self.smoothingPasses = int(ap.get("smoothingPasses", 25))
        self.maxTimeStepMCTRatio = float(ap.get("maxTimeStepMCTRatio", 0.2))
        self.maxCourantNumber = float(ap.get("maxCourantNumber", 0.5))
        self.maxSubCellSizeMFPRatio = float(ap.get("maxSubCellSizeMFPRatio", 0.5))
        self.minSubCellLevels, self.maxSubCellLevels, self.theta = 1, 10, 0.2   # :48-58
        cw = props.get("cellWeightedProperties", {})
        self.particlesPerSubCell = int(cw.get("particlesPerSubCell", 20))
        self.minParticlesPerSubCell = int(cw.get("minParticlesPerSubCell", self.particlesPerSubCell))
        self.maxCellWeightRatio, self.maxSmoothingPasses = 0.05, 500              # uniGasCloud.C:418-419
        if self.cellWeightAdaptation and not cloud.cellWeighted:
            raise ValueError("cellWeightAdaptation needs cellWeightedSimulation true")
        self.species = [props["moleculeProperties"][n] for n in props["typeIdList"]]
        self.Tref = float(props.get("collisionProperties", {}).get("Tref", 273.0))
        self.bgkName = props.get("bgkCollisionModel", "noBGKCollision")
        self.ops = FaceOperators(self.mesh, halo, reduce_max if halo is not None else None)
        nC = self.mesh.n_cells
        self.prevCellSizeMFPRatio = np.zeros((nC, 3))
        self.timeSteps = 0
        self._snap = None           # accumulators at the start of the current interval
        self._carry = None          # sums folded in before the accumulators were reset (resetAtOutput)
        self.subCellLevels = np.ones((nC, 3), np.int32) if cloud._subCellLevels is None else np.rint(cloud._subCellLevels).astype(np.int32)
        self.cellWeightFactor = np.ones(nC) if cloud._cellWeightFactor is None else cloud._cellWeightFactor.copy()
        self.cellCollModelId = None  # hybrid runs: set from the cloud's decomposition before adapt()
        self.last = {}
        cloud._adapter = self

    # ---- sums over the interval ------------------------------------------------------------------------------
    @staticmethod
    def _pack(a):
        acc, sp = a["acc"], a["species"]
        return np.column_stack([acc[:, 0], acc[:, 8], acc[:, 9], acc[:, 13], acc[:, 10:13], sp]), a["timeAvCounter"]

    def _begin_interval(self):
        self._snap = self._pack(self.cloud.accumulators())
        self._carry = None

    def before_reset(self):
        """The cloud is about to zero the accumulators (fields(resetAtOutput=True)): fold what this interval has so far."""
        if self._snap is None:
            return
        now, t = self._pack(self.cloud.accumulators())
        d = (now - self._snap[0], t - self._snap[1])
        self._carry = d if self._carry is None else (self._carry[0] + d[0], self._carry[1] + d[1])
        self._snap = (np.zeros_like(now), 0.0)

    def _interval_sums(self):
        now, t = self._pack(self.cloud.accumulators())
        s, tt = now - self._snap[0], t - self._snap[1]
        if self._carry is not None:
            s, tt = s + self._carry[0], tt + self._carry[1]
        return s, tt

    # ---- the reference's pieces ------------------------------------------------------------------------------
    def adaptation_quantities(self, rhoN, transT, U, speciesRhoN, deltaT):
        """calculateAdaptationQuantities for all cells at once (:228-333)."""
        nC, nS = len(rhoN), len(self.species)
        ok = transT > SMALL
        T = np.where(ok, transT, 1.0)
        MCRs, MFPs = np.zeros((nC, nS)), np.zeros((nC, nS))
        u0 = np.zeros(nC)
        for i, a in enumerate(self.species):
            for q, b in enumerate(self.species):
                dPQ = 0.5 * (a["diameter"] + b["diameter"])
                omegaPQ = 0.5 * (a["omega"] + b["omega"])
                massRatio = a["mass"] / b["mass"]
                mr = a["mass"] * b["mass"] / (a["mass"] + b["mass"])
                on = (speciesRhoN[:, q] > VSMALL) & (transT > VSMALL)
                MCRs[:, i] += np.where(on, 2.0 * math.sqrt(math.pi) * dPQ * dPQ * speciesRhoN[:, q] * (T / self.Tref) ** (1.0 - omegaPQ)
                                       * math.sqrt(2.0 * kB * self.Tref / mr), 0.0)                                    # Bird 4.74
                MFPs[:, i] += np.where(on, math.pi * dPQ * dPQ * speciesRhoN[:, q] * (self.Tref / T) ** (omegaPQ - 0.5)
                                       * math.sqrt(1.0 + massRatio), 0.0)                                               # Bird 4.76
            u0 = np.where(speciesRhoN[:, i] > VSMALL, np.maximum(u0, np.sqrt(2.0 * kB / a["mass"] * T)), u0)
        present = speciesRhoN > VSMALL
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.where(present, 1.0 / MFPs, 0.0)
            frac = np.where(present, speciesRhoN / rhoN[:, None], 0.0)
        MCR = (MCRs * frac).sum(1)   # Bird 1.38
        MFP = (inv * frac).sum(1)    # Bird 4.77
        size = self.mesh.cell_bb_max - self.mesh.cell_bb_min
        with np.errstate(divide="ignore", invalid="ignore"):
            tsr = np.where(ok, deltaT * MCR, 0.0)
            cou = np.where(ok[:, None], np.maximum(u0[:, None], U) * deltaT / size, 0.0)
            csr = np.where(ok[:, None], size / MFP[:, None], self.prevCellSizeMFPRatio)
        return tsr, cou, csr

    def calculate_time_step(self, tsr, cou, deltaT, collId):
        """calculateTimeStep (:335-388)."""
        dsmcCell = np.ones(len(tsr), bool) if collId is None else (np.asarray(collId) != 0)
        mct = 0.0
        if self.bgkName != "unifiedStochasticParticleSBGK" and dsmcCell.any():
            mct = float(max(0.0, tsr[dsmcCell].max()))
        solved = [d for d in range(3) if self.mesh.solution_d[d]]
        co = float(max(0.0, cou[:, solved].max())) if solved else 0.0
        if mct > VSMALL and co < VSMALL:
            deltaT *= self.maxTimeStepMCTRatio / mct
        elif mct < VSMALL and co > VSMALL:
            deltaT *= self.maxCourantNumber / co
        elif mct > VSMALL and co > VSMALL:
            deltaT *= min(self.maxTimeStepMCTRatio / mct, self.maxCourantNumber / co)
        return float(self.reduce_min(deltaT)), mct, co

    def calculate_sub_cell_levels(self, csr, collId):
        """calculateSubCellLevels (:390-431)."""
        lv = np.minimum(self.maxSubCellLevels, np.maximum(self.minSubCellLevels, np.ceil(csr / self.maxSubCellSizeMFPRatio))).astype(np.int32)
        if collId is not None:
            lv[np.asarray(collId) == 0] = self.minSubCellLevels  # BGK cells
        for d in range(3):
            if not self.mesh.solution_d[d]:
                lv[:, d] = 1
        return lv

    def _cell_rwf(self):
        """uniGasCloud::axiRWF at the cell centres (uniGasCloudI.H:116-120); 1 without axisymmetricSimulation."""
        cfg = self.cloud.cfg
        if not cfg.axisymmetric:
            return 1.0
        cc = self.mesh.cell_centres
        return 1.0 + (cfg.maxRWF - 1.0) * np.sqrt(cc[:, 1] * cc[:, 1] + cc[:, 2] * cc[:, 2]) / cfg.radialExtent

    def calculate_cell_weight_factor(self, rhoN, levels):
        """calculateCellWeightFactor (:433-446); RWF of the cell centre with axisymmetricSimulation (:442)."""
        nSub = levels.prod(1).astype(float)
        return rhoN * self.mesh.cell_volumes / (self.particlesPerSubCell * nSub * self.cloud.cfg.nParticle * self._cell_rwf())

    def smooth_cell_weight_factor(self, rhoN, cwf, levels):
        """smoothCellWeightFactor (:448-500)."""
        nSub = levels.prod(1).astype(float)
        cap = rhoN * self.mesh.cell_volumes / (self.minParticlesPerSubCell * nSub * self.cloud.cfg.nParticle * self._cell_rwf())  # :473
        cwf = self.ops.smooth(cwf, 1.3)
        passes = 0
        while True:
            passes += 1
            cwf = self.ops.average_interpolate(cwf)
            cwf = np.maximum(np.minimum(cap, cwf), SMALL)
            if not (self.reduce_max(self.ops.max_neighbour_ratio(cwf)) > 1.0 + self.maxCellWeightRatio and passes < self.maxSmoothingPasses):
                break
        return cwf, passes

    # ---- driving ---------------------------------------------------------------------------------------------
    def adapt(self):
        """One call per step, after evolve (uniGasCloud::adaptation, U/clouds/uniGasCloud.C:223-234).  Returns True when
        this call was an adaptation step."""
        if self._snap is None:
            raise RuntimeError("call begin() before the first step of an interval (run() does)")
        self.timeSteps += 1
        if self.timeSteps != self.adaptationInterval:
            return False
        s, tAv = self._interval_sums()
        V = self.mesh.cell_volumes
        nC, nS = self.mesh.n_cells, len(self.species)
        have = s[:, 0] > VSMALL
        with np.errstate(divide="ignore", invalid="ignore"):
            rhoN = np.where(have, s[:, 1] / (V * tAv), 0.0)
            rhoM = s[:, 2] / (V * tAv)
            U = np.where(have[:, None], s[:, 4:7] / (rhoM * V * tAv)[:, None], 0.0)
            T = np.where(have, 2.0 / (3.0 * kB * rhoN) * (0.5 * s[:, 3] / (V * tAv) - 0.5 * rhoM * (U * U).sum(1)), 0.0)
        good = have & (T > VSMALL)
        minRhoN = rhoN[good].min() if good.any() else VGREAT
        minT = T[good].min() if good.any() else VGREAT
        spRho = s[:, 7:7 + nS] / (V * tAv)[:, None]
        empty = rhoN == 0.0   # :583-594: empty cells take the smallest non-zero density / temperature
        rhoN = np.where(empty, minRhoN, rhoN)
        spRho = np.where(empty[:, None], (rhoN / nS)[:, None], spRho)
        T = np.where(T == 0.0, minT, T)
        deltaT = self.cloud.cfg.deltaT
        tsr, cou, csr = self.adaptation_quantities(rhoN, T, U, spRho, deltaT)
        for _ in range(self.smoothingPasses):   # :614-619: only the size / mean-free-path ratio is smoothed here
            csr = self.ops.average_interpolate(csr)
        collId = self.cellCollModelId
        self.last = dict(rhoN=rhoN, translationalT=T, UMean=U, timeStepMCTRatio=tsr, courantNumber=cou, cellSizeMFPRatio=csr, timeAv=tAv)
        if self.timeStepAdaptation:
            newDt, mct, co = self.calculate_time_step(tsr, cou, deltaT, collId)
            self.last.update(maxTimeStepMCTRatio=mct, maxCourant=co, deltaT=newDt)
            self.cloud.setDeltaT(newDt)
        kw = {}
        if self.subCellAdaptation:
            self.subCellLevels = self.calculate_sub_cell_levels(csr, collId)
            kw["subCellLevels"] = self.subCellLevels
        if self.cellWeightAdaptation:
            target = self.calculate_cell_weight_factor(rhoN, self.subCellLevels)
            target, passes = self.smooth_cell_weight_factor(rhoN, target, self.subCellLevels)
            self.cellWeightFactor = self.theta * target + (1.0 - self.theta) * self.cellWeightFactor   # :672-677
            kw["cellWeightFactor"] = self.cellWeightFactor
            self.last.update(cellWeightTarget=target, cellWeightSmoothingPasses=passes)
        if kw:
            self.cloud.setCellState(**kw)
        self.prevCellSizeMFPRatio = csr
        self.timeSteps = 0
        self._begin_interval()
        return True

    def set_initial_configuration(self, speciesRhoN, transT, U):
        """setInitialConfiguration (:708-775), called by uniGasMeshFill before the parcels are created: the adaptation
        quantities of the uniform initial state, smoothed, give the first time step and sub-cell levels.  Returns
        (deltaT, subCellLevels); the cloud's time step is set when a cloud is attached."""
        nC, nS = self.mesh.n_cells, len(self.species)
        sp = np.stack([np.broadcast_to(np.asarray(v, float), (nC,)) for v in speciesRhoN], axis=1)  # scalars or per-cell fields
        rhoN = sp.sum(1)
        deltaT = self.cloud.cfg.deltaT
        tsr, cou, csr = self.adaptation_quantities(rhoN, np.broadcast_to(np.asarray(transT, float), (nC,)).copy(),
                                                   np.broadcast_to(np.asarray(U, float), (nC, 3)).copy(), sp, deltaT)
        for _ in range(self.smoothingPasses):
            tsr = self.ops.average_interpolate(tsr)
            cou = self.ops.average_interpolate(cou)
            csr = self.ops.average_interpolate(csr)
        self.prevCellSizeMFPRatio = csr
        if self.timeStepAdaptation:
            deltaT, _, _ = self.calculate_time_step(tsr, cou, deltaT, self.cellCollModelId)
            if hasattr(self.cloud, "setDeltaT"):
                self.cloud.setDeltaT(deltaT)
            else:
                self.cloud.cfg.deltaT = deltaT
        if self.subCellAdaptation:
            self.subCellLevels = self.calculate_sub_cell_levels(csr, self.cellCollModelId)
        return deltaT, self.subCellLevels

    def begin(self):
        self._begin_interval()

    def run(self, nSteps):
        """evolve + adaptation for nSteps steps; the steps between two adaptations go to the device in one call."""
        if self._snap is None:
            self.begin()
        done, adapted = 0, 0
        while done < nSteps:
            k = min(self.adaptationInterval - self.timeSteps, nSteps - done)
            self.cloud.evolve(k)
            done += k
            self.timeSteps += k - 1
            adapted += bool(self.adapt())
        return adapted
