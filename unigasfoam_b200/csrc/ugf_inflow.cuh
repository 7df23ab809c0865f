// ugf_inflow.cuh — free-stream insertion and multi-rank parcel migration (pack / unpack).
//
// Inflow replaces uniGasFreeStreamInflowPatch::controlParcelsBeforeMove
// (U/boundaries/derived/generalBoundaries/uniGasFreeStreamInflowPatch/uniGasFreeStreamInflowPatch.C:111-128) =
// uniGasGeneralBoundary::computeParcelsToInsert (…/uniGasGeneralBoundary.C:115-169, Bird eq 4.22) +
// insertParcels (:537-761, Bird eq 12.5).  Per-face triangle fans are flattened once on the host; the per-step
// work is count -> single-block scan -> insert, so new parcels land at deterministic indices (face-major,
// species, k), the same order the oracle appends them in.
//
// Migration replaces the transfer loop of OpenFOAM's Cloud::move: parcels stopped on a processor
// face (cell = -2 - boundaryFace) are compacted, in index order, into a send buffer of UGF_MIGRATE_STRIDE
// doubles per parcel; received records are appended and resume tracking from their stepFraction.
#pragma once
#include "ugf_common.cuh"
#include "ugf_move.cuh"
#include "ugf_sort.cuh"

namespace ugf {

constexpr int INFLOW_GEOM = 13;  // fA, n[3] (into the domain), t1[3], t2[3], p0[3]
constexpr int INFLOW_TRI = 7;    // a[3], b[3], cumulative area fraction

struct InflowDev {
    int nFaces, nTypeIds;
    int typeIds[UGF_MAX_SPECIES];
    double numDen[UGF_MAX_SPECIES];
    double Ttr, Trot, Tvib, Tel;
    double vel[3];
    int ce;                           // uniGasChapmanEnskogFreeStreamInflowPatch: heat flux, stress (row-major), mixture pressure n k T
    double ceQ[3], ceS[9], cePressure;
    double molFrac[UGF_MAX_SPECIES];  // 1 for free-stream patches (numDen is per species there)
    double* faceVel;                  // pressure inlets and field patches: inflow velocity per face [nFaces*3], else null (vel everywhere)
    double theta;                     // pressure inlets: relaxation of faceVel towards the cell mean velocity
    int pressure;                     // pressure inlet (speed ratio of the count formula capped at 5)
    // uniGasWangPressureInletPatch: running sums per face [nFaces*WANG_NSUM] (else null), inlet pressure, mixture molecular
    // mass, gamma * R of the mixture
    double* wangSums;
    double wangP, wangM, wangGammaR;
    double* faceN;                    // field patch / pressure outlet: number density per slot [nFaces*nTypeIds], else null
    double* faceT;                    // field patch / pressure outlet: translational, rotational temperature per face [nFaces*2], else null
    // uniGasLiouFangPressureOutletPatch: faceN / faceT / faceVel follow the flow (outlet_state_kernel); the count of a slot is
    // capped at the count a gas at (capN, capT) with a speed ratio of 5 would give - the host's insertion bound - and a hit
    // raises device error 6
    int outlet;
    double capN, capT;
    int* err;
    // uniGasMassFlowRateInletPatch: faceN / faceVel follow the flow (mass_flow_face_kernel, mass_flow_scale_kernel); parcels are
    // inserted from a gas at rest (…MassFlowRateInletPatch.C:139-149).  mfWork [nFaces*(nTypeIds+1)]: per face the parcelsIn
    // contribution of every species and the parcelsToInsert contribution; mfTotal [2]: expected insertions of the next step, ok flag
    int massFlow;
    double massFlowRate, patchArea, mfMolFrac[UGF_MAX_SPECIES];
    double* outFlux;
    double* mfWork;
    double* mfTotal;
    const int* faceBfi;
    const int* faceCell;
    const double* geom;
    const int* triOff;
    const double* tri;
    int* nIns;
    int* insOff;
};

__global__ void __launch_bounds__(256) inflow_count_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InflowDev f, uint32_t step) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= f.nFaces * f.nTypeIds) return;
    const int face = slot / f.nTypeIds, iD = slot % f.nTypeIds;
    const double* g = f.geom + (size_t)face * INFLOW_GEOM;
    const DevSpecies& s = prm.sp[f.typeIds[iD]];
    const double fA = g[0];
    const double Ttr = f.faceT ? f.faceT[2 * (size_t)face] : f.Ttr;
    const double numDen = f.faceN ? f.faceN[slot] : f.numDen[iD];
    const double cmp = sqrt(2.0 * kB * Ttr / s.mass);
    const double* vel = f.faceVel ? f.faceVel + 3 * (size_t)face : f.vel;
    double sCos = (vel[0] * g[1] + vel[1] * g[2] + vel[2] * g[3]) / cmp;
    if (f.pressure && sCos > 5.0) sCos = 5.0;  // the host's insertion bound assumes speed ratios <= 5 on pressure inlets
    const double sqrtPi = sqrt(PI);
    // nParticle * CWF(face cell) * RWF(face centre) (uniGasGeneralBoundary.C:154-155)
    const double fnFace = prm.axi ? cell_fn(prm, f.faceCell[face]) * __ldg(&prm.bfRwf[f.faceBfi[face]]) : cell_fn(prm, f.faceCell[face]);
    double accum = f.molFrac[iD] * (fA * numDen * prm.deltaT * cmp * (exp(-(sCos * sCos)) + sqrtPi * sCos * (1 + erf(sCos))))
                   / (2.0 * sqrtPi * fnFace);  // uniGasGeneralBoundary.C:154-165
    if (f.ce) {  // :171-239: the normal stress and heat flux correct the Maxwellian flux
        const double n[3] = {g[1], g[2], g[3]};
        const double qn = f.ceQ[0] * n[0] + f.ceQ[1] * n[1] + f.ceQ[2] * n[2];
        double snn = 0.0;
        for (int k = 0; k < 3; ++k) snn += (f.ceS[3 * k] * n[0] + f.ceS[3 * k + 1] * n[1] + f.ceS[3 * k + 2] * n[2]) * n[k];
        accum = (fA * numDen * prm.deltaT * cmp
                 * (exp(-(sCos * sCos)) * (1.0 - 0.5 * snn / f.cePressure - 0.4 * qn * sCos / f.cePressure / cmp) + sqrtPi * sCos * (1 + erf(sCos))))
                / (2.0 * sqrtPi * fnFace);
    }
    if (f.outlet) {
        const double cmpCap = sqrt(2.0 * kB * f.capT / s.mass);
        const double cap = f.molFrac[iD] * (fA * f.capN * prm.deltaT * cmpCap * (exp(-25.0) + sqrtPi * 5.0 * (1 + erf(5.0))))
                           / (2.0 * sqrtPi * fnFace);
        if (!(accum <= cap)) { accum = accum > cap ? cap : 0.0; atomicExch(f.err, 6); }
    }
    Stream rc(prm.seed, KIND_INFLOW, (uint32_t)iD, step, (uint32_t)f.faceBfi[face], 0);
    int nIns = max((int)accum, 0);
    if ((accum - nIns) > rc.u01()) ++nIns;
    f.nIns[slot] = nIns;
}

// single block: insOff = exclusive scan of nIns; insOff[nSlots] = old array length (base); *dN += total
__global__ void __launch_bounds__(SCAN_THREADS) inflow_scan_kernel(const int* __restrict__ nIns, int nSlots, int* __restrict__ insOff,
                                                                  long long* dN, long long capacity, DevCounters* cnt, int* errFlag) {
    __shared__ int sm[33];
    int carry = 0;
    for (int base = 0; base < nSlots; base += SCAN_THREADS) {
        const int idx = base + threadIdx.x;
        const int v = idx < nSlots ? nIns[idx] : 0;
        int t;
        const int ex = block_exclusive_scan(v, &t, sm);
        if (idx < nSlots) insOff[idx] = carry + ex;
        carry += t;
    }
    if (threadIdx.x == 0) {
        const long long base = *dN;
        insOff[nSlots] = (int)base;
        if (base + carry > capacity) { *errFlag = 1; insOff[nSlots] = -1; }
        else { *dN = base + carry; atomicAdd(&cnt->inserted, (unsigned long long)carry); }
    }
}

template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(128) inflow_insert_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InflowDev f,
                                                            ParcelBuf P, uint32_t step) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int nSlots = f.nFaces * f.nTypeIds;
    if (slot >= nSlots) return;
    const int base = f.insOff[nSlots];
    if (base < 0) return;
    const int nIns = f.nIns[slot];
    if (nIns == 0) return;
    const int face = slot / f.nTypeIds, iD = slot % f.nTypeIds;
    const double* g = f.geom + (size_t)face * INFLOW_GEOM;
    const double n[3] = {g[1], g[2], g[3]}, t1[3] = {g[4], g[5], g[6]}, t2[3] = {g[7], g[8], g[9]}, p0[3] = {g[10], g[11], g[12]};
    const int typeId = f.typeIds[iD];
    const DevSpecies& s = prm.sp[typeId];
    const double Ttr = f.faceT ? f.faceT[2 * (size_t)face] : f.Ttr;
    const double Trot = f.faceT ? f.faceT[2 * (size_t)face + 1] : f.Trot;
    const double cmp = sqrt(2.0 * kB * Ttr / s.mass);
    const double atRest[3] = {0.0, 0.0, 0.0};
    const double* vel = f.massFlow ? atRest : (f.faceVel ? f.faceVel + 3 * (size_t)face : f.vel);
    const double vn = vel[0] * n[0] + vel[1] * n[1] + vel[2] * n[2];
    const double sCos = vn / cmp;
    const int tb = f.triOff[face], te = f.triOff[face + 1];
    const int cellI = f.faceCell[face];
    const uint32_t bfi = (uint32_t)f.faceBfi[face];
    for (int i = 0; i < nIns; ++i) {
        Stream r(prm.seed, KIND_INFLOW, (uint32_t)iD, step, bfi, (uint32_t)(i + 1));
        const double triSel = r.u01();
        int sel = tb;
        for (int t = tb; t < te; ++t) { sel = t; if (f.tri[(size_t)t * INFLOW_TRI + 6] >= triSel) break; }
        const double* tr = f.tri + (size_t)sel * INFLOW_TRI;
        double bs = r.u01(), bt = r.u01();
        if (bs + bt > 1) { bs = 1 - bs; bt = 1 - bt; }
        double x[3];
        for (int k = 0; k < 3; ++k) x[k] = (1 - bs - bt) * p0[k] + bs * tr[k] + bt * tr[3 + k];
        double U[3];
        if (f.ce) {  // Chapman-Enskog velocity (uniGasGeneralBoundary.C:880-940), same draw order as the oracle
            double maxQ = -1.0, maxS = -1.0;
            for (int k = 0; k < 3; ++k) maxQ = fmax(maxQ, fabs(f.ceQ[k]));
            for (int k = 0; k < 9; ++k) maxS = fmax(maxS, fabs(f.ceS[k]));
            const double breakdown = fmax(2.0 * maxQ / (f.cePressure * cmp), maxS / f.cePressure);
            const double amplitude = 1.0 + 60.0 * breakdown;
            const double lower = fmin(sCos - 4.0, -5.0), upper = fmin(sCos, 5.0);
            const double uMax = 0.5 * (sCos - sqrt(sCos * sCos + 2.0));
            double Uc[3], gamma;
            do {
                double uN;
                if (fabs(vn) > VSMALL) {
                    do { uN = lower + r.u01() * (upper - lower); }
                    while ((sCos - uN) / (sCos - uMax) * exp(uMax * uMax - uN * uN) < r.u01());
                } else {
                    uN = -sqrt(-log(1.0 - r.u01()));
                }
                double g1c, g2c;
                r.gauss2(g1c, g2c);
                for (int k = 0; k < 3; ++k) Uc[k] = g1c / sqrt(2.0) * t1[k] + g2c / sqrt(2.0) * t2[k] - uN * n[k];
                const double* S9 = f.ceS;
                const double qU = f.ceQ[0] * Uc[0] + f.ceQ[1] * Uc[1] + f.ceQ[2] * Uc[2];
                const double UU = Uc[0] * Uc[0] + Uc[1] * Uc[1] + Uc[2] * Uc[2];
                gamma = 1.0 + (2.0 / cmp * qU * (0.4 * UU - 1.0) - 2.0 * (S9[1] * Uc[0] * Uc[1] + S9[2] * Uc[0] * Uc[2] + S9[5] * Uc[1] * Uc[2])
                               - S9[0] * (Uc[0] * Uc[0] - Uc[2] * Uc[2]) - S9[4] * (Uc[1] * Uc[1] - Uc[2] * Uc[2])) / f.cePressure;
            } while (amplitude * r.u01() > gamma);
            for (int k = 0; k < 3; ++k) U[k] = cmp * Uc[k] + vel[k];
        } else {
        const double A = sCos + sqrt(sCos * sCos + 2.0);
        const double B = 0.5 * (1.0 + sCos * (sCos - sqrt(sCos * sCos + 2.0)));
        double scaling = 3.0;
        if (sCos < -3) scaling = fabs(sCos) + 1;
        double Pp = -1, uNormal, uNormalThermal;
        if (fabs(vn) > VSMALL) {
            do {  // Bird eq 12.5
                uNormalThermal = scaling * (2.0 * r.u01() - 1);
                uNormal = uNormalThermal + sCos;
                if (uNormal < 0.0) Pp = -1;
                else Pp = 2.0 * uNormal / A * exp(B - uNormalThermal * uNormalThermal);
            } while (Pp < r.u01());
        } else {
            uNormal = sqrt(-log(1.0 - r.u01()));
        }
        double g1, g2;
        r.gauss2(g1, g2);
        const double cth = sqrt(kB * Ttr / s.mass);
        const double vt1 = t1[0] * vel[0] + t1[1] * vel[1] + t1[2] * vel[2];
        const double vt2 = t2[0] * vel[0] + t2[1] * vel[1] + t2[2] * vel[2];
        for (int k = 0; k < 3; ++k) U[k] = cth * (g1 * t1[k] + g2 * t2[k]) + vt1 * t1[k] + vt2 * t2[k] + cmp * uNormal * n[k];
        }
        const double erot = equipartition_rotational_energy(r, Trot, s.rotDoF);
        const long long dst = (long long)base + f.insOff[slot] + i;
        P.x[dst] = x[0]; P.y[dst] = x[1]; P.z[dst] = x[2];
        P.ux[dst] = U[0]; P.uy[dst] = U[1]; P.uz[dst] = U[2];
        P.cell[dst] = cellI;
        if (HAS_ROT) P.erot[dst] = erot;
        if (MULTI) P.type[dst] = (uint8_t)typeId;
        if (prm.spi) {  // uniGasGeneralBoundary.C:721-735: vibrational and electronic levels at the patch's temperatures
            const DevSpeciesInt& S = prm.spi[typeId];
            if (P.vib) P.vib[dst] = s.vibDoF > 0 ? equipartition_vib_levels(r, f.Tvib, S, s.vibDoF) : 0ull;
            if (P.elev) P.elev[dst] = s.nElec > 1 ? (uint8_t)equipartition_elec_level(r, f.Tel, S, s.nElec) : (uint8_t)0;
        }
    }
}

// uniGasMassFlowRateInletPatch::controlParcelsAfterCollisions (…/uniGasMassFlowRateInletPatch.C:155-302), first half: per face
// (one thread, cell-list order = the oracle's) the relaxed inlet velocity, the number density of every species from the parcels
// in the cell, and the face's contributions to parcelsIn (per species) and parcelsToInsert.
__global__ void __launch_bounds__(128) mass_flow_face_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InflowDev f, ParcelBuf P,
                                                             const int* __restrict__ off, const double* __restrict__ vol, bool multi) {
    const int face = blockIdx.x * blockDim.x + threadIdx.x;
    if (face >= f.nFaces) return;
    const int nT = f.nTypeIds;
    const int c = f.faceCell[face];
    const double* g = f.geom + (size_t)face * INFLOW_GEOM;
    const double fA = g[0];
    const double nIn[3] = {g[1], g[2], g[3]};
    const double nParticle = prm.nParticle, dt = prm.deltaT;
    const double CWF = prm.cwf ? __ldg(&prm.cwf[c]) : 1.0;
    const double RWFf = prm.axi ? __ldg(&prm.bfRwf[f.faceBfi[face]]) : 1.0;
    double totalMass = 0.0;
    for (int i = 0; i < nT; ++i) totalMass += prm.sp[f.typeIds[i]].mass * f.mfMolFrac[i];
    double* work = f.mfWork + (size_t)face * (nT + 1);
    for (int i = 0; i < nT; ++i) {
        const double moleFlowRate = f.mfMolFrac[i] * (f.massFlowRate / totalMass);
        double* fl = f.outFlux + (size_t)face * prm.nSpecies + i;
        work[i] = moleFlowRate * dt * (fA / f.patchArea) / (nParticle * CWF * RWFf) + *fl / (CWF * RWFf);
        *fl = 0.0;  // uniGasFaceTracker::clean at the end of the step (uniGasCloud.C:864)
    }
    double mom[3] = {0, 0, 0}, mass = 0.0;
    double* nD = f.faceN + (size_t)face * nT;
    for (int i = 0; i < nT; ++i) nD[i] = 0.0;
    for (int j = off[c]; j < off[c + 1]; ++j) {
        const int t = multi ? P.type[j] : 0;
        const double pMass = nParticle * prm.sp[t].mass;
        const double RWF = prm.axi ? axi_rwf(prm, P.y[j], P.z[j]) : 1.0;
        nD[t] += 1.0;
        mom[0] += pMass * CWF * RWF * P.ux[j]; mom[1] += pMass * CWF * RWF * P.uy[j]; mom[2] += pMass * CWF * RWF * P.uz[j];
        mass += pMass * CWF * RWF;
    }
    double* v = f.faceVel + 3 * (size_t)face;
    const double prev[3] = {v[0], v[1], v[2]};
    double nv[3] = {0, 0, 0};
    if (mass > VSMALL) for (int k = 0; k < 3; ++k) nv[k] = mom[k] / mass;
    double nw[3];
    for (int k = 0; k < 3; ++k) nw[k] = f.theta * nv[k] + (1.0 - f.theta) * prev[k];
    if (nw[0] * nIn[0] + nw[1] * nIn[1] + nw[2] * nIn[2] < 0.0) for (int k = 0; k < 3; ++k) nw[k] = prev[k];
    for (int k = 0; k < 3; ++k) v[k] = nw[k];
    for (int i = 0; i < nT; ++i) nD[i] = nD[i] * nParticle * CWF * RWFf / vol[c];
    double pti = 0.0;
    const double sqrtPi = sqrt(PI);
    for (int i = 0; i < nT; ++i) {
        const double cmp = sqrt(2.0 * kB * f.Ttr / prm.sp[f.typeIds[i]].mass);
        const double sCos = (nw[0] * nIn[0] + nw[1] * nIn[1] + nw[2] * nIn[2]) / cmp;
        pti += (fA * nD[i] * dt * cmp * (exp(-(sCos * sCos)) + sqrtPi * sCos * (1 + erf(sCos)))) / (2.0 * sqrtPi * nParticle * CWF * RWFf);
    }
    work[nT] = pti;
}

// second half: the patch-wide sums in face order (one thread: the oracle's order, identical bits) and the scaling of the number
// densities by parcelsIn / parcelsToInsert; mfTotal[0] = the insertions the next step will make, mfTotal[1] = 0 when the
// reference's ratio is 0 / 0 (no parcel in any inlet cell)
__global__ void mass_flow_scale_kernel(const __grid_constant__ InflowDev f) {
    const int nT = f.nTypeIds;
    if (blockIdx.x != 0) return;
    __shared__ double sIn[UGF_MAX_SPECIES];
    __shared__ double sTo;
    if (threadIdx.x == 0) {
        double pIn[UGF_MAX_SPECIES], pTo = 0.0;
        for (int i = 0; i < nT; ++i) pIn[i] = 0.0;
        for (int face = 0; face < f.nFaces; ++face) {
            const double* work = f.mfWork + (size_t)face * (nT + 1);
            for (int i = 0; i < nT; ++i) pIn[i] += work[i];
        }
        // parcelsToInsert_ += <slot count> adds to every species' total (scalarField += scalar): one total, species inside faces
        for (int face = 0; face < f.nFaces; ++face) pTo += f.mfWork[(size_t)face * (nT + 1) + nT];
        double tot = 0.0;
        for (int i = 0; i < nT; ++i) { sIn[i] = pIn[i]; tot += pIn[i]; }
        sTo = pTo;
        f.mfTotal[0] = tot;
        f.mfTotal[1] = pTo > 0.0 ? 1.0 : 0.0;
    }
    __syncthreads();
    if (!(sTo > 0.0)) return;
    for (int s = threadIdx.x; s < f.nFaces * nT; s += blockDim.x) f.faceN[s] = f.faceN[s] * (sIn[s % nT] / sTo);
}

// uniGasLiouFangPressureInletPatch::controlParcelsAfterCollisions (…/uniGasLiouFangPressureInletPatch.C:126-174): the
// inflow velocity of every inlet face moves towards the mass-weighted mean velocity of the parcels now in its cell.  One
// thread per face, summing in cell-list order (the oracle's order: identical bits); the array is cell-major here.
template <bool MULTI>
__global__ void __launch_bounds__(128) inlet_velocity_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InflowDev f, ParcelBuf P,
                                                             const int* __restrict__ off) {
    const int face = blockIdx.x * blockDim.x + threadIdx.x;
    if (face >= f.nFaces) return;
    const int c = f.faceCell[face];
    const double w = cell_fn(prm, c);  // nParticle * CWF: the same for every parcel of the cell
    double mom[3] = {0, 0, 0}, mass = 0;
    for (int j = off[c]; j < off[c + 1]; ++j) {
        const double m = (prm.axi ? w * axi_rwf(prm, P.y[j], P.z[j]) : w) * (MULTI ? prm.sp[P.type[j]].mass : prm.sp[0].mass);  // nParticle*CWF*RWF(position)*mass
        mom[0] += m * P.ux[j]; mom[1] += m * P.uy[j]; mom[2] += m * P.uz[j];
        mass += m;
    }
    double* v = f.faceVel + 3 * (size_t)face;
    for (int k = 0; k < 3; ++k) {
        const double nv = mass > 0 ? mom[k] / mass : 0.0;
        v[k] = f.theta * nv + (1.0 - f.theta) * v[k];
    }
}

// uniGasWangPressureInletPatch::controlParcelsAfterCollisions (…/uniGasWangPressureInletPatch.C:131-281): running sums of the
// parcels seen in each inlet face's cell since the start -> inlet velocity = mean momentum / mean mass, after 100 steps
// corrected with the characteristic relation (p_cell - p_in) / (rho a) along the outward normal.  Sums per face: 0 parcels,
// 1 mass, 2-4 momentum, 5-7 sum U^2 per component, 8-10 sum U per component (the last two only advance once more than one
// parcel has been seen, :206-209).  One thread per face, cell-list order (= the oracle's: identical bits).
constexpr int WANG_NSUM = 11;
template <bool MULTI>
__global__ void __launch_bounds__(128) wang_inlet_velocity_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InflowDev f, ParcelBuf P,
                                                                  const int* __restrict__ off, const double* __restrict__ vol, double nTimeSteps) {
    const int face = blockIdx.x * blockDim.x + threadIdx.x;
    if (face >= f.nFaces) return;
    const int c = f.faceCell[face];
    const double w = cell_fn(prm, c);
    double mom[3] = {0, 0, 0}, mass = 0, nP = 0, sq[3] = {0, 0, 0}, su[3] = {0, 0, 0};
    for (int j = off[c]; j < off[c + 1]; ++j) {
        const int t = MULTI ? P.type[j] : 0;
        bool mine = false;
        for (int i = 0; i < f.nTypeIds; ++i) mine = mine || f.typeIds[i] == t;
        if (!mine) continue;
        const double m = (prm.axi ? w * axi_rwf(prm, P.y[j], P.z[j]) : w) * prm.sp[t].mass;  // nParticle*CWF*RWF(position)*mass
        const double U[3] = {P.ux[j], P.uy[j], P.uz[j]};
        for (int k = 0; k < 3; ++k) { mom[k] += m * U[k]; sq[k] += U[k] * U[k]; su[k] += U[k]; }
        mass += m;
        nP += 1.0;
    }
    double* S = f.wangSums + (size_t)face * WANG_NSUM;
    S[0] += nP; S[1] += mass;
    for (int k = 0; k < 3; ++k) S[2 + k] += mom[k];
    if (S[0] > 1) {
        for (int k = 0; k < 3; ++k) { S[5 + k] += sq[k]; S[8 + k] += su[k]; }
        const double massDensity = S[1] / (vol[c] * nTimeSteps);
        const double numberDensity = massDensity / f.wangM;
        double m2 = 0, mm = 0;
        for (int k = 0; k < 3; ++k) { m2 += S[5 + k] / S[0]; const double a = S[8 + k] / S[0]; mm += a * a; }
        double T = (0.5 * f.wangM) * (2.0 / (3.0 * kB)) * (m2 - mm);
        if (T < VSMALL) T = 300.0;
        const double pressure = numberDensity * kB * T;
        const double sound = sqrt(f.wangGammaR * T);
        const double* g = f.geom + (size_t)face * INFLOW_GEOM;  // g[1..3]: unit normal into the domain
        double* v = f.faceVel + 3 * (size_t)face;
        for (int k = 0; k < 3; ++k) v[k] = S[2 + k] / S[1];
        if (nTimeSteps > 100) {
            const double corr = (pressure - f.wangP) / (massDensity * sound);
            for (int k = 0; k < 3; ++k) v[k] += corr * -g[1 + k];
        }
    }
}

// uniGasLiouFangPressureOutletPatch::controlParcelsAfterCollisions (…/uniGasLiouFangPressureOutletPatch.C:144-322): the same
// running sums as the Wang inlet, except that the parcel count and the velocity moments run over all species (:196-207);
// from them the cell's density, temperature and pressure, then (Liou & Fang 2000, eq 26) the outlet state
//   rho_e = rho + (p_e - p) / a^2,  n_e = rho_e / m,  T_e = p_e / (R rho_e),  u_e = <m u> / <m> + (p - p_e) / (rho a) n_out.
// A non-positive rho_e (the reference would divide by it) switches the face's insertion off for the step.
template <bool MULTI>
__global__ void __launch_bounds__(128) outlet_state_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InflowDev f, ParcelBuf P,
                                                           const int* __restrict__ off, const double* __restrict__ vol, double nTimeSteps) {
    const int face = blockIdx.x * blockDim.x + threadIdx.x;
    if (face >= f.nFaces) return;
    const int c = f.faceCell[face];
    const double w = cell_fn(prm, c);
    double mom[3] = {0, 0, 0}, mass = 0, nP = 0, sq[3] = {0, 0, 0}, su[3] = {0, 0, 0};
    for (int j = off[c]; j < off[c + 1]; ++j) {
        const int t = MULTI ? P.type[j] : 0;
        bool mine = false;
        for (int i = 0; i < f.nTypeIds; ++i) mine = mine || f.typeIds[i] == t;
        const double U[3] = {P.ux[j], P.uy[j], P.uz[j]};
        if (mine) {
            const double m = (prm.axi ? w * axi_rwf(prm, P.y[j], P.z[j]) : w) * prm.sp[t].mass;  // nParticle*CWF*RWF(position)*mass
            for (int k = 0; k < 3; ++k) mom[k] += m * U[k];
            mass += m;
        }
        for (int k = 0; k < 3; ++k) { sq[k] += U[k] * U[k]; su[k] += U[k]; }
        nP += 1.0;
    }
    double* S = f.wangSums + (size_t)face * WANG_NSUM;
    S[0] += nP; S[1] += mass;
    for (int k = 0; k < 3; ++k) S[2 + k] += mom[k];
    if (S[0] > 1) {
        for (int k = 0; k < 3; ++k) { S[5 + k] += sq[k]; S[8 + k] += su[k]; }
        const double massDensity = S[1] / (vol[c] * nTimeSteps);
        const double numberDensity = massDensity / f.wangM;
        double m2 = 0, mm = 0;
        for (int k = 0; k < 3; ++k) { m2 += S[5 + k] / S[0]; const double a = S[8 + k] / S[0]; mm += a * a; }
        double T = (0.5 * f.wangM) * (2.0 / (3.0 * kB)) * (m2 - mm);
        if (T < VSMALL) T = 300.0;
        const double pressure = numberDensity * kB * T;
        const double sound = sqrt(f.wangGammaR * T);
        const double rhoE = massDensity + (f.wangP - pressure) / (sound * sound);
        const double* g = f.geom + (size_t)face * INFLOW_GEOM;  // g[1..3]: unit normal into the domain
        double* v = f.faceVel + 3 * (size_t)face;
        for (int k = 0; k < 3; ++k) v[k] = S[1] > 0 ? S[2 + k] / S[1] : 0.0;
        if (massDensity > 0) {
            const double corr = (pressure - f.wangP) / (massDensity * sound);
            for (int k = 0; k < 3; ++k) v[k] += corr * -g[1 + k];
        }
        const double nE = rhoE > 0 ? rhoE / f.wangM : 0.0;
        for (int i = 0; i < f.nTypeIds; ++i) f.faceN[(size_t)face * f.nTypeIds + i] = nE;
        if (rhoE > 0) {
            const double TE = f.wangP / ((kB / f.wangM) * rhoE);
            f.faceT[2 * (size_t)face] = TE; f.faceT[2 * (size_t)face + 1] = TE;
        }
    }
}

// ---- migration ------------------------------------------------------------------------------------------

__device__ __forceinline__ bool on_patch(const MeshDev& mesh, int cell, int patch) {
    return cell <= -2 && mesh.bfPatch[-2 - cell] == patch;
}

// record slot 8 = local face index + MIG_TYPE_SHIFT * typeId (both exact in a double); slot 9 = the cell weight factor the
// parcel carries (cell weighting; 1 otherwise)
constexpr double MIG_TYPE_SHIFT = 4294967296.0;

__global__ void __launch_bounds__(1024) mig_count_kernel(MeshDev mesh, const int* __restrict__ cell, const long long* dN, int patch,
                                                         int* __restrict__ blockCounts) {
    __shared__ int sm[33];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int flag = (i < *dN) && on_patch(mesh, cell[i], patch);
    int total;
    block_exclusive_scan(flag, &total, sm);
    if (threadIdx.x == 0) blockCounts[blockIdx.x] = total;
}

template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(1024) mig_pack_kernel(MeshDev mesh, ParcelBuf P, const double* __restrict__ sf, const double* __restrict__ wq, const long long* dN, int patch,
                                                        const int* __restrict__ blockOffsets, double* __restrict__ buf) {
    __shared__ int sm[33];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int c = -1;
    if (i < *dN) c = P.cell[i];
    const int flag = (i < *dN) && on_patch(mesh, c, patch);
    int total;
    const int ex = block_exclusive_scan(flag, &total, sm);
    if (flag) {
        double* r = buf + (size_t)(blockOffsets[blockIdx.x] + ex) * UGF_MIGRATE_STRIDE;
        const int bfi = -2 - c;
        r[0] = P.x[i]; r[1] = P.y[i]; r[2] = P.z[i];
        r[3] = P.ux[i]; r[4] = P.uy[i]; r[5] = P.uz[i];
        r[6] = HAS_ROT ? P.erot[i] : 0.0;
        r[7] = sf[i];
        r[8] = (double)(bfi - mesh.patches[patch].startBfi) + (MULTI ? MIG_TYPE_SHIFT * (double)P.type[i] : 0.0);
        r[9] = wq ? wq[i] : 1.0;
        P.cell[i] = -1;
    }
}

template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(256) mig_unpack_kernel(MeshDev mesh, ParcelBuf P, double* __restrict__ sf, double* __restrict__ wq, long long base, long long n, int patch,
                                                         const double* __restrict__ buf, int* errFlag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = buf + (size_t)i * UGF_MIGRATE_STRIDE;
    const long long dst = base + i;
    const int type = (int)(r[8] * (1.0 / MIG_TYPE_SHIFT));
    const int lf = (int)(r[8] - MIG_TYPE_SHIFT * type);
    const DevPatch& pt = mesh.patches[patch];
    if (lf < 0 || lf >= pt.size) { *errFlag = 2; P.cell[dst] = -1; return; }
    P.x[dst] = r[0]; P.y[dst] = r[1]; P.z[dst] = r[2];
    P.ux[dst] = r[3]; P.uy[dst] = r[4]; P.uz[dst] = r[5];
    if (HAS_ROT) P.erot[dst] = r[6];
    sf[dst] = r[7];
    P.cell[dst] = mesh.bfOwner[pt.startBfi + lf];
    if (MULTI) P.type[dst] = (uint8_t)type;
    if (wq) wq[dst] = r[9];
}

// ---- fixed-slot migration: all processor patches in two passes over the parcels, no host round trip ------------
// where slot k of a pack goes: the local send buffer (NCCL path) or, over NVLink peer memory, directly the matching
// receive slot of the neighbouring rank
struct MigDst {
    double* slot[MIG_MAXP];
};

__device__ __forceinline__ int mig_slot_of(const MeshDev& mesh, const MigSlots& ms, int cell) {
    if (cell > -2) return -1;
    const int p = mesh.bfPatch[-2 - cell];
    int slot = -1;
#pragma unroll
    for (int k = 0; k < MIG_MAXP; ++k) if (k < ms.nProc && ms.patch[k] == p) slot = k;
    return slot;
}

constexpr int MIG_THREADS = 256;
constexpr int MIG_PER_THREAD = 8;
constexpr int MIG_TILE = MIG_THREADS * MIG_PER_THREAD;  // parcels per block in the slot pack passes

// pass 1: per (slot, block) number of waiting parcels; 8 consecutive cell ids per thread
// dStart (optional): only parcels with index >= *dStart can be waiting (second and later transfer rounds of a step:
// only what was received in this step); tiles entirely below it are skipped.
__global__ void __launch_bounds__(MIG_THREADS) mig_count_all_kernel(MeshDev mesh, MigSlots ms, const int* __restrict__ cell, const long long* dN,
                                                                    const long long* dStart, int* __restrict__ blockCounts, int nBlocks) {
    __shared__ int cnt[MIG_MAXP];
    if (dStart && (long long)(blockIdx.x + 1) * MIG_TILE <= *dStart) {
        if (threadIdx.x < ms.nProc) blockCounts[threadIdx.x * nBlocks + blockIdx.x] = 0;
        return;
    }
    if (threadIdx.x < MIG_MAXP) cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long n = *dN;
    const long long base = (long long)blockIdx.x * MIG_TILE + (long long)threadIdx.x * MIG_PER_THREAD;
    int c[MIG_PER_THREAD];
    if (base + MIG_PER_THREAD <= n) {
        const int4 v0 = __ldg(reinterpret_cast<const int4*>(cell + base));
        const int4 v1 = __ldg(reinterpret_cast<const int4*>(cell + base) + 1);
        c[0] = v0.x; c[1] = v0.y; c[2] = v0.z; c[3] = v0.w; c[4] = v1.x; c[5] = v1.y; c[6] = v1.z; c[7] = v1.w;
    } else {
#pragma unroll
        for (int e = 0; e < MIG_PER_THREAD; ++e) c[e] = (base + e < n) ? cell[base + e] : -1;
    }
    int mn = c[0];
#pragma unroll
    for (int e = 1; e < MIG_PER_THREAD; ++e) mn = min(mn, c[e]);
    if (mn <= -2) {  // rare: some of this thread's parcels wait on a processor patch
#pragma unroll
        for (int e = 0; e < MIG_PER_THREAD; ++e) {
            const int slot = mig_slot_of(mesh, ms, c[e]);
            if (slot >= 0) atomicAdd(&cnt[slot], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < ms.nProc) blockCounts[threadIdx.x * nBlocks + blockIdx.x] = cnt[threadIdx.x];
}

// one block per slot: exclusive scan of that slot's per-block counts -> blockOffsets, total to totals[slot]
__global__ void __launch_bounds__(SCAN_THREADS) mig_scan_kernel(const int* __restrict__ blockCounts, int* __restrict__ blockOffsets, int nBlocks, int* totals) {
    __shared__ int sm[33];
    const int* b = blockCounts + (size_t)blockIdx.x * nBlocks;
    int* o = blockOffsets + (size_t)blockIdx.x * nBlocks;
    int carry = 0;
    for (int base = 0; base < nBlocks; base += SCAN_THREADS) {
        const int idx = base + threadIdx.x;
        const int v = idx < nBlocks ? b[idx] : 0;
        int t;
        const int ex = block_exclusive_scan(v, &t, sm);
        if (idx < nBlocks) o[idx] = carry + ex;
        carry += t;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// pass 2: blocks whose tile holds no waiting parcel return after reading their counts; the others rank their
// waiting parcels in index order per slot and write the records
template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(MIG_THREADS) mig_pack_all_kernel(MeshDev mesh, MigSlots ms, ParcelBuf P, const double* __restrict__ sf, const double* __restrict__ wq, const long long* dN,
                                                                   const int* __restrict__ blockCounts, const int* __restrict__ blockOffsets,
                                                                   const int* __restrict__ totals, int nBlocks, const MigDst dst,
                                                                   long long slotCapacity, int* errFlag) {
    __shared__ int sm[33];
    if (blockIdx.x == 0 && threadIdx.x < ms.nProc) {  // headers
        double* hdr = dst.slot[threadIdx.x];
        const int tot = totals[threadIdx.x];
        hdr[0] = (double)(tot <= slotCapacity ? tot : slotCapacity);
        for (int j = 1; j < UGF_MIGRATE_STRIDE; ++j) hdr[j] = 0.0;
        if (tot > slotCapacity) *errFlag = 3;
    }
    int any = 0;
    for (int k = 0; k < ms.nProc; ++k) any |= blockCounts[k * nBlocks + blockIdx.x];
    if (any == 0) return;
    const long long n = *dN;
    const long long base = (long long)blockIdx.x * MIG_TILE + (long long)threadIdx.x * MIG_PER_THREAD;
    int slot[MIG_PER_THREAD], c[MIG_PER_THREAD];
#pragma unroll
    for (int e = 0; e < MIG_PER_THREAD; ++e) {
        c[e] = (base + e < n) ? P.cell[base + e] : -1;
        slot[e] = mig_slot_of(mesh, ms, c[e]);
    }
    for (int k = 0; k < ms.nProc; ++k) {
        if (blockCounts[k * nBlocks + blockIdx.x] == 0) continue;  // uniform over the block
        int mine = 0;
#pragma unroll
        for (int e = 0; e < MIG_PER_THREAD; ++e) mine += (slot[e] == k);
        int total;
        int pos = block_exclusive_scan(mine, &total, sm) + blockOffsets[k * nBlocks + blockIdx.x];
#pragma unroll
        for (int e = 0; e < MIG_PER_THREAD; ++e) {
            if (slot[e] != k) continue;
            const long long i = base + e;
            if (pos < slotCapacity) {
                double* r = dst.slot[k] + (1 + (long long)pos) * UGF_MIGRATE_STRIDE;
                const int bfi = -2 - c[e];
                r[0] = P.x[i]; r[1] = P.y[i]; r[2] = P.z[i];
                r[3] = P.ux[i]; r[4] = P.uy[i]; r[5] = P.uz[i];
                r[6] = HAS_ROT ? P.erot[i] : 0.0;
                r[7] = sf[i];
                r[8] = (double)(bfi - mesh.patches[ms.patch[k]].startBfi) + (MULTI ? MIG_TYPE_SHIFT * (double)P.type[i] : 0.0);
                r[9] = wq ? wq[i] : 1.0;
            }
            P.cell[i] = -1;
            ++pos;
        }
    }
}

// Unpack every slot in one launch (blockIdx.y = slot): slot k appends after the parcels of slots < k.
template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(256) mig_unpack_all_kernel(MeshDev mesh, MigSlots ms, ParcelBuf P, double* __restrict__ sf, double* __restrict__ wq, const long long* dN,
                                                             long long capacity, const double* __restrict__ recv, long long slotCapacity, int* errFlag) {
    const int k = blockIdx.y;
    const long long slotStride = (slotCapacity + 1) * UGF_MIGRATE_STRIDE;
    const double* slot = recv + k * slotStride;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)slot[0];
    if (n < 0 || n > slotCapacity) { if (i == 0) *errFlag = 4; return; }
    if (i >= n) return;
    long long base = *dN;
    for (int j = 0; j < k; ++j) base += (long long)recv[j * slotStride];
    const long long dst = base + i;
    if (dst >= capacity) { *errFlag = 1; return; }
    const double* r = slot + (1 + i) * UGF_MIGRATE_STRIDE;
    const int type = (int)(r[8] * (1.0 / MIG_TYPE_SHIFT));
    const int lf = (int)(r[8] - MIG_TYPE_SHIFT * type);
    const DevPatch& pt = mesh.patches[ms.patch[k]];
    if (lf < 0 || lf >= pt.size) { *errFlag = 2; P.cell[dst] = -1; return; }
    P.x[dst] = r[0]; P.y[dst] = r[1]; P.z[dst] = r[2];
    P.ux[dst] = r[3]; P.uy[dst] = r[4]; P.uz[dst] = r[5];
    if (HAS_ROT) P.erot[dst] = r[6];
    sf[dst] = r[7];
    P.cell[dst] = mesh.bfOwner[pt.startBfi + lf];
    if (MULTI) P.type[dst] = (uint8_t)type;
    if (wq) wq[dst] = r[9];
}

// after the unpack: remember where the received parcels start, extend the array, reset the in-flight tally
__global__ void mig_commit_kernel(long long* dN, long long* dRecvStart, unsigned long long* inflight, const double* recv, int nProc,
                                  long long slotCapacity, long long capacity, int* errFlag) {
    const long long slotStride = (slotCapacity + 1) * UGF_MIGRATE_STRIDE;
    long long add = 0;
    for (int k = 0; k < nProc; ++k) {
        const long long n = (long long)recv[k * slotStride];
        if (n > 0 && n <= slotCapacity) add += n;
    }
    *dRecvStart = *dN;
    if (*dN + add <= capacity) *dN += add; else *errFlag = 1;
    *inflight = 0ull;
}

// Pack from the migrant lists the move kernel filled (one block per processor patch): the indices are sorted
// ascending in shared memory (bitonic; the records must leave in index order, like the search-based pack above)
// and the records written straight to the slot.  No pass over the cloud at all.
constexpr int MIG_LIST_CAP = 8192;  // migrants per patch and round the list path handles (32 KB of shared memory)

struct MigFlags {
    unsigned long long* flag[MIG_MAXP];
};

template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(1024) mig_pack_list_kernel(MeshDev mesh, MigSlots ms, ParcelBuf P, const double* __restrict__ sf, const double* __restrict__ wq, int* __restrict__ migCount,
                                                             const int* __restrict__ migList, int listCap, const MigDst dst, long long slotCapacity,
                                                             int* errFlag, const MigFlags fl, unsigned long long epoch, unsigned long long* inflight) {
    __shared__ int s[MIG_LIST_CAP];
    const int k = blockIdx.x;
    const int patch = ms.patch[k];
    int n = migCount[patch];
    if (n > listCap || n > slotCapacity) {
        if (threadIdx.x == 0) *errFlag = 3;
        n = (int)min((long long)min(n, listCap), slotCapacity);
    }
    int m = 1;
    while (m < n) m <<= 1;
    for (int j = threadIdx.x; j < m; j += blockDim.x) s[j] = j < n ? migList[(size_t)k * listCap + j] : 0x7fffffff;
    __syncthreads();
    for (int kk = 2; kk <= m; kk <<= 1)
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            for (int t = threadIdx.x; t < m; t += blockDim.x) {
                const int p = t ^ jj;
                if (p > t) {
                    const int x = s[t], y = s[p];
                    const bool up = (t & kk) == 0;
                    if ((x > y) == up) { s[t] = y; s[p] = x; }
                }
            }
            __syncthreads();
        }
    double* slot = dst.slot[k];
    if (threadIdx.x == 0) {
        slot[0] = (double)n;
        for (int j = 1; j < UGF_MIGRATE_STRIDE; ++j) slot[j] = 0.0;
    }
    const int startBfi = mesh.patches[patch].startBfi;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int i = s[j];
        double* r = slot + (1 + (long long)j) * UGF_MIGRATE_STRIDE;
        const int bfi = -2 - P.cell[i];
        r[0] = P.x[i]; r[1] = P.y[i]; r[2] = P.z[i];
        r[3] = P.ux[i]; r[4] = P.uy[i]; r[5] = P.uz[i];
        r[6] = HAS_ROT ? P.erot[i] : 0.0;
        r[7] = sf[i];
        r[8] = (double)(bfi - startBfi) + (MULTI ? MIG_TYPE_SHIFT * (double)P.type[i] : 0.0);
        r[9] = wq ? wq[i] : 1.0;
        P.cell[i] = -1;
    }
    // peer-memory rounds (epoch != 0): the records went straight into the neighbour's receive slot - publish the round's epoch in its
    // flag word behind them (system-scope release), one kernel instead of pack + signal
    if (epoch) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        migCount[patch] = 0;
        if (epoch) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl.flag[k]), "l"(epoch) : "memory");
        if (k == 0 && inflight) *inflight = 0ull;  // everything that waited on a processor patch leaves with this round
    }
}

// ---- NVLink peer-memory transfer: the pack kernel above has written straight into the neighbours' receive slots ----
// signal: one thread per processor patch publishes the round's epoch in the receiver's flag word, after a system-scope
// fence that orders it behind the records written by the (completed) pack kernel.
__global__ void mig_signal_kernel(const MigFlags f, int nProc, unsigned long long epoch) {
    __threadfence_system();
    if ((int)threadIdx.x < nProc) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.flag[threadIdx.x]), "l"(epoch) : "memory");
    }
}

// wait: spin until every local flag has reached the epoch (the neighbours' records for this round are complete);
// gives up after ~4 s of GPU clock with the handle's error flag raised rather than hanging the device.
__global__ void mig_wait_kernel(const unsigned long long* flags, int nProc, unsigned long long epoch, int* errFlag) {
    if ((int)threadIdx.x < nProc) {
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > 8000000000LL) { *errFlag = 5; break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// ---- fused receive: wait for the neighbours' flags, append their records and continue the tracks, all in one launch --------------
// Replaces mig_wait_kernel + mig_unpack_all_kernel + mig_commit_kernel + move_kernel (four dependent launches per transfer round,
// each a few microseconds of launch latency on an otherwise idle GPU).  Grid: x covers a slot, y = slot.  Every block first spins on
// all local flags (the block needs every slot's count for its append position), then each thread places one record and tracks it to
// the end of the step (or to the next processor patch).  The last block to finish publishes the new array length.
struct RecvArgs {
    MigSlots ms;
    const double* recv;
    const unsigned long long* flags;
    unsigned long long epoch;  // 0: no flag wait (slots filled by a completed NCCL transfer)
    long long slotCapacity, capacity;
    long long* dN;
    long long* dRecvStart;
    unsigned int* done;        // zero on entry, zero again on exit
    int* errFlag;
};

template <bool HAS_ROT, bool MULTI, int NF>
__global__ void __launch_bounds__(256, 4) mig_recv_move_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ MoveArgs a,
                                                              const __grid_constant__ RecvArgs ra) {
    if (ra.epoch && (int)threadIdx.x < ra.ms.nProc) {
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ra.flags + threadIdx.x) : "memory");
            if (v >= ra.epoch) break;
            if (clock64() - t0 > 8000000000LL) { *ra.errFlag = 5; break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    const int k = blockIdx.y;
    const long long slotStride = (ra.slotCapacity + 1) * UGF_MIGRATE_STRIDE;
    const double* slot = ra.recv + k * slotStride;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)slot[0];
    if (n < 0 || n > ra.slotCapacity) { if (i == 0) *ra.errFlag = 4; n = 0; }
    long long base = *ra.dN;
    long long add = 0;
    for (int j = 0; j < ra.ms.nProc; ++j) {
        long long nj = (long long)ra.recv[j * slotStride];
        if (nj < 0 || nj > ra.slotCapacity) nj = 0;
        if (j < k) base += nj;
        add += nj;
    }
    bool valid = i < n;
    const long long dst = base + i;
    if (valid && dst >= ra.capacity) { *ra.errFlag = 1; valid = false; }
    int cell = -1;
    double x0 = 0, x1 = 0, x2 = 0, U0 = 0, U1 = 0, U2 = 0;
    if (valid) {
        const double* r = slot + (1 + i) * UGF_MIGRATE_STRIDE;
        const int type = (int)(r[8] * (1.0 / MIG_TYPE_SHIFT));
        const int lf = (int)(r[8] - MIG_TYPE_SHIFT * type);
        const DevPatch& pt = a.mesh.patches[ra.ms.patch[k]];
        x0 = r[0]; x1 = r[1]; x2 = r[2]; U0 = r[3]; U1 = r[4]; U2 = r[5];
        a.P.x[dst] = x0; a.P.y[dst] = x1; a.P.z[dst] = x2;
        a.P.ux[dst] = U0; a.P.uy[dst] = U1; a.P.uz[dst] = U2;
        if (HAS_ROT) a.P.erot[dst] = r[6];
        a.sf[dst] = r[7];
        if (MULTI) a.P.type[dst] = (uint8_t)type;
        if (a.wq) a.wq[dst] = r[9];
        if (lf < 0 || lf >= pt.size) { *ra.errFlag = 2; a.P.cell[dst] = -1; }
        else { cell = a.mesh.bfOwner[pt.startBfi + lf]; a.P.cell[dst] = cell; }
    }
    track_parcel<HAS_ROT, MULTI, NF>(prm, a, dst, valid, cell, x0, x1, x2, U0, U1, U2);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(ra.done, 1u) == total - 1) {  // every block has read the old length and placed its records
            *ra.dRecvStart = *ra.dN;
            if (*ra.dN + add <= ra.capacity) *ra.dN += add; else *ra.errFlag = 1;
            *ra.done = 0u;
        }
    }
}

}  // namespace ugf
