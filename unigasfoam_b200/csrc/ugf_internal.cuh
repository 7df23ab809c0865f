// ugf_internal.cuh — internal energy beyond rotation: vibrational quantum levels (quantum-kinetic exchange) and electronic levels.
//
// Replaces uniGasCloud::equipartitionVibrationalEnergyLevel / equipartitionElectronicLevel / postCollisionVibrationalEnergyLevel /
// postCollisionElectronicEnergyLevel (U/clouds/uniGasCloud.C:1020-1126, 1192-1326), the vibrational and electronic blocks of
// LarsenBorgnakkeVariableHardSphere::collide (…/LarsenBorgnakkeVariableHardSphere.C:125-416), the vibrational / electronic sums of
// cellMeasurements::calculateFields (U/cellMeasurements/cellMeasurements.C:436-510) and their accumulation
// (…/uniGasVolFields.C:775-793).
//
// Storage: one 64-bit word per parcel holding the quantum level of up to four modes (16 bits each) and one byte holding the
// electronic level; both arrays exist only when some species has vibrational modes / more than one electronic level, every
// consumer tests the pointer.  The per-species tables live in one small device array (DevParams::spi).  None of this is on the
// path of a gas without such species.  Draw order is shared with the oracle.
#pragma once
#include "ugf_common.cuh"
#include "ugf_rng.cuh"

namespace ugf {

__device__ __forceinline__ int vib_level(unsigned long long v, int m) { return (int)((v >> (16 * m)) & 0xFFFFull); }
__device__ __forceinline__ unsigned long long vib_set(unsigned long long v, int m, int level) {
    const unsigned long long l = (unsigned long long)(level < 0 ? 0 : (level > 65535 ? 65535 : level));
    return (v & ~(0xFFFFull << (16 * m))) | (l << (16 * m));
}

__device__ inline double vib_energy(const DevSpeciesInt& S, int vibDoF, unsigned long long v) {
    double e = 0.0;
    for (int m = 0; m < vibDoF; ++m) e += vib_level(v, m) * S.thetaV[m] * kB;
    return e;
}

// equipartitionVibrationalEnergyLevel (uniGasCloud.C:1020-1050)
__device__ inline unsigned long long equipartition_vib_levels(Stream& r, double T, const DevSpeciesInt& S, int vibDoF) {
    unsigned long long v = 0ull;
    for (int m = 0; m < vibDoF; ++m) v = vib_set(v, m, (int)(-log(1.0 - r.u01()) * T / S.thetaV[m]));
    return v;
}

// equipartitionElectronicLevel (uniGasCloud.C:1053-1126)
__device__ inline int equipartition_elec_level(Stream& r, double T, const DevSpeciesInt& S, int jMax) {
    if (jMax == 1) return 0;
    if (T < VSMALL) return 0;
    const double EMax = kB * T;
    double expSum = 0.0;
    for (int i = 0; i < jMax; ++i) expSum += S.g[i] * exp(-S.elecE[i] / EMax);
    double boltzMax = 0.0;
    int jSelect = 0;
    for (int i = 0; i < jMax; ++i) {
        const double boltz = S.g[i] * exp(-S.elecE[i] / EMax) / expSum;
        if (boltzMax < boltz) { boltzMax = boltz; jSelect = i; }
    }
    const double expMax = S.g[jSelect] * exp(-S.elecE[jSelect] / EMax);
    const double eps = r.u01();
    int jDash;
    double func;
    do {
        jDash = r.position(jMax);
        func = S.g[jDash] * exp(-S.elecE[jDash] / EMax) / expMax;
    } while (!(func > eps));
    return jDash;
}

// postCollisionVibrationalEnergyLevel, postReaction = false (uniGasCloud.C:1192-1264)
__device__ inline int post_collision_vib_level(Stream& r, int vibLevel, int iMax, double thetaV, double thetaD, double refTempZv, double omega,
                                               double Zref, double Ec) {
    int iDash = vibLevel;
    const double TColl = (iMax * thetaV) / (3.5 - omega);
    const double pow1 = pow(thetaD / TColl, 0.33333) - 1.0;
    const double pow2 = pow(thetaD / refTempZv, 0.33333) - 1.0;
    const double ZvP1 = pow(thetaD / TColl, omega);
    const double ZvP2 = pow(Zref * pow(thetaD / refTempZv, -omega), pow1 / pow2);
    const double Zv = ZvP1 * ZvP2;
    const double inverseVibrationalCollisionNumber = 1.0 / (5.0 * Zv);
    if (inverseVibrationalCollisionNumber > r.u01()) {
        double func, EVib;
        do {
            const int i = (int)(r.u01() * (iMax + 1));  // Random::position<label>(0, iMax)
            iDash = i < iMax ? i : iMax;
            EVib = iDash * kB * thetaV;
            func = pow(1.0 - EVib / Ec, 1.5 - omega);
        } while (!(func > r.u01()));
    }
    return iDash;
}

// postCollisionElectronicEnergyLevel (uniGasCloud.C:1267-1326), any number of levels
__device__ inline int post_collision_elec_level(Stream& r, double Ec, int jMax, double omega, const DevSpeciesInt& S) {
    int nPossStates = 0;
    if (jMax == 1) nPossStates = S.g[0];
    else for (int n = 0; n < jMax; ++n) if (Ec > S.elecE[n]) nPossStates += S.g[n];
    for (;;) {
        const int nState = (int)ceil(r.u01() * nPossStates);
        int nAvail = 0, nLevel = -1;
        for (int n = 0; n < jMax; ++n) {
            nAvail += S.g[n];
            if (nState <= nAvail && nLevel < 0) nLevel = n;
        }
        if (nLevel < 0) nLevel = 0;
        if (Ec > S.elecE[nLevel]) {
            const double prob = pow(1.0 - S.elecE[nLevel] / Ec, 1.5 - omega);
            if (prob > r.u01()) return nLevel;
        }
    }
}

// ---- per-step sums of the internal modes, per (cell, species) ----------------------------------------------------------------------
// One warp per cell over the cell's parcels (through the occupancy permutation when the array is not cell-major yet).  Writes the
// per-step block momI [cell][species][UGF_NINT] = {count, sum E_el, parcels in the ground level, in the first level, sum E_vib per
// mode}, adds accDt x it to the time accumulators and - when the moment blocks are kept - fills moment slots 22 (sum E_vib), 23-25
// (sum E_vib U), 26 (sum E_el), 27-30 (sum E_vib per mode).
struct InternalArgs {
    int nCells;
    const int* off;
    const int* perm;   // null: identity
    ParcelBuf in;
    double* momI;
    double* accI;
    double* mom;       // null: moment blocks are not kept this step
    double accDt;
};

__global__ void __launch_bounds__(256) internal_modes_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InternalArgs a) {
    const int lane = threadIdx.x & 31;
    const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= a.nCells) return;
    const int b = a.off[cell], e = a.off[cell + 1];
    const int nS = prm.nSpecies;
    for (int s = 0; s < nS; ++s) {
        const DevSpecies& sp = prm.sp[s];
        const DevSpeciesInt& S = prm.spi[s];
        double cnt = 0, eel = 0, nG = 0, nF = 0, ev[UGF_MAX_VIB_MODES] = {0, 0, 0, 0}, evt = 0, evu = 0, evv = 0, evw = 0;
        for (int j = b + lane; j < e; j += 32) {
            const int src = a.perm ? (a.perm[j] & CLONE_MASK) : j;
            if (nS > 1 && a.in.type[src] != s) continue;
            cnt += 1.0;
            const int lev = a.in.elev ? a.in.elev[src] : 0;
            eel += S.elecE[lev];
            if (sp.nElec > 1) { nG += (lev == 0); nF += (lev == 1); }
            if (sp.vibDoF > 0) {
                const unsigned long long v = a.in.vib[src];
                double tot = 0;
                for (int m = 0; m < sp.vibDoF; ++m) { const double x = vib_level(v, m) * kB * S.thetaV[m]; ev[m] += x; tot += x; }
                evt += tot;
                if (a.mom) { evu += tot * a.in.ux[src]; evv += tot * a.in.uy[src]; evw += tot * a.in.uz[src]; }
            }
        }
        cnt = warp_sum(cnt); eel = warp_sum(eel); nG = warp_sum(nG); nF = warp_sum(nF); evt = warp_sum(evt);
        for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) ev[m] = warp_sum(ev[m]);
        if (a.mom) { evu = warp_sum(evu); evv = warp_sum(evv); evw = warp_sum(evw); }
        if (lane == 0) {
            const double v[UGF_NINT] = {cnt, eel, nG, nF, ev[0], ev[1], ev[2], ev[3]};
            double* I = a.momI + ((size_t)cell * nS + s) * UGF_NINT;
            double* A = a.accI + ((size_t)cell * nS + s) * UGF_NINT;
            for (int k = 0; k < UGF_NINT; ++k) {
                I[k] = v[k];
                if (a.accDt != 0.0) A[k] += a.accDt * v[k];
            }
            if (a.mom) {
                double* M = a.mom + ((size_t)cell * nS + s) * UGF_NMOM;
                M[22] = evt; M[23] = evu; M[24] = evv; M[25] = evw; M[26] = eel;
                for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) M[27 + m] = ev[m];
            }
        }
    }
}

__global__ void __launch_bounds__(256) accumulate_internal_kernel(long long n, double dt, const double* __restrict__ momI, double* __restrict__ accI) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) accI[i] += dt * momI[i];
}

// info(): sum of the vibrational and of the electronic energy over the live parcels -> tot[0], tot[1]
__global__ void __launch_bounds__(256) internal_totals_kernel(const __grid_constant__ DevParams prm, ParcelBuf P, const long long* dN, double* tot) {
    const long long n = *dN;
    double ev = 0, ee = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (P.cell[i] < 0) continue;
        const int t = P.type ? P.type[i] : 0;
        if (P.vib) ev += vib_energy(prm.spi[t], prm.sp[t].vibDoF, P.vib[i]);
        ee += prm.spi[t].elecE[P.elev ? P.elev[i] : 0];
    }
    ev = warp_sum(ev); ee = warp_sum(ee);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&tot[0], ev); atomicAdd(&tot[1], ee); }
}

// vibrationalT, electronicT and the internal degrees of freedom that enter overallT (uniGasVolFields.C:930-1079) for one cell
__device__ inline void derive_internal(const DevParams& prm, const double* accI /* this cell: [nSpecies][UGF_NINT] */, double rhoNMeanInt, double rhoNMean,
                                       double& vibT, double& elecT, double& totalvDof, double& totalEDof) {
    vibT = 0; elecT = 0; totalvDof = 0; totalEDof = 0;
    const int nS = prm.nSpecies;
    double molsElec = 0;
    for (int s = 0; s < nS; ++s) if (prm.sp[s].nElec > 1) molsElec += accI[(size_t)s * UGF_NINT];
    for (int s = 0; s < nS; ++s) {
        const DevSpecies& sp = prm.sp[s];
        const DevSpeciesInt& S = prm.spi[s];
        const double* I = accI + (size_t)s * UGF_NINT;
        double dofSpecies = 0, vibTID = 0, dofMode[UGF_MAX_VIB_MODES] = {0, 0, 0, 0}, vibTMode[UGF_MAX_VIB_MODES] = {0, 0, 0, 0};
        for (int v = 0; v < sp.vibDoF; ++v) {
            if (I[4 + v] > VSMALL && I[0] > VSMALL) {
                const double iMean = (I[4 + v] / I[0]) / (kB * S.thetaV[v]);
                vibTMode[v] = S.thetaV[v] / log(1.0 + 1.0 / iMean);
                dofMode[v] = (2.0 * S.thetaV[v] / vibTMode[v]) / (exp(S.thetaV[v] / vibTMode[v]) - 1.0);
            }
            dofSpecies += dofMode[v];
        }
        for (int v = 0; v < sp.vibDoF; ++v)
            if (dofSpecies > VSMALL) vibTID += vibTMode[v] * dofMode[v] / dofSpecies;
        totalvDof += dofSpecies;
        if (rhoNMeanInt > VSMALL && rhoNMean > VSMALL && I[0] > VSMALL) vibT += vibTID * I[0] / rhoNMeanInt;
        if (sp.nElec > 1 && I[2] > VSMALL && I[3] > VSMALL && I[3] * S.g[0] != I[2] * S.g[1]) {
            const double elecTID = (S.elecE[1] - S.elecE[0]) / (kB * log((I[2] * S.g[1]) / (I[3] * S.g[0])));
            const double fraction = I[0] / molsElec;
            if (elecTID > VSMALL) elecT += fraction * elecTID;
            totalEDof += fraction * ((2.0 * (I[1] / I[0])) / (kB * elecTID));
        }
    }
}

}  // namespace ugf
