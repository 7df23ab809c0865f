// ugf_fields.cuh — time-averaged field accumulation and write-time derivation, plus the info() totals.
//
// Replaces uniGasVolFields::calculateField (U/macroscopicProperties/derived/volumetric/uniGasVolFields/
// uniGasVolFields.C:723-837 accumulation, :839-1254 cell fields, :1256-1352 wall fields) and the energy sums
// of uniGasCloud::info (U/clouds/uniGasCloud.C:878-920).  One thread per cell / boundary face; each thread
// reads its own 256-byte moment block, so every fetched sector is consumed.
#pragma once
#include "ugf_common.cuh"
#include "ugf_internal.cuh"

namespace ugf {

// accS (multi-species only, else null): nParcelsXnParticle per species for the mean-free-path fields; with one
// species it equals accumulator slot 8.
__global__ void __launch_bounds__(256) accumulate_cells_kernel(const __grid_constant__ DevParams prm, int nCells,
                                                               const double* __restrict__ mom, double* __restrict__ acc, double* __restrict__ accS) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const double dt = prm.deltaT, FN = cell_fn(prm, c);
    double A[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) A[k] = acc[(size_t)c * NACC + k];
    for (int s = 0; s < prm.nSpecies; ++s) {
        const double* a = mom + ((size_t)c * prm.nSpecies + s) * UGF_NMOM;
        const DevSpecies& S = prm.sp[s];
        const double ms = S.mass;
        const double a0 = a[0], a1 = a[1];
        A[0] += dt * a0;
        A[1] += dt * (ms * a0);
        A[2] += dt * (ms * (a[8] + a[11] + a[13]));
        A[3] += dt * (ms * a[2]); A[4] += dt * (ms * a[3]); A[5] += dt * (ms * a[4]);
        A[6] += dt * a[18];
        A[7] += dt * (S.rotDoF * a0);
        A[8] += dt * (a1 * FN);
        A[9] += dt * (ms * a1 * FN);
        A[10] += dt * (ms * a[5] * FN); A[11] += dt * (ms * a[6] * FN); A[12] += dt * (ms * a[7] * FN);
        A[13] += dt * (ms * a[14] * FN);
        A[14] += dt * (S.rotDoF > 0 ? a0 : 0.0);
        A[15] += dt * ((5.0 + S.rotDoF) * a0);
        if (accS) accS[(size_t)c * prm.nSpecies + s] += dt * (a1 * FN);
    }
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[(size_t)c * NACC + k] = A[k];
}

// bacc += dt * bm ; bm = 0  (boundaryMeas_.clean, U/clouds/uniGasCloud.C:865)
__global__ void __launch_bounds__(256) accumulate_walls_kernel(double dt, int accumulate, long long n, double* __restrict__ bm, double* __restrict__ bacc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = bm[i];
    if (accumulate && v != 0.0) bacc[i] += dt * v;
    if (v != 0.0) bm[i] = 0.0;
}

__global__ void __launch_bounds__(256) derive_cells_kernel(const __grid_constant__ DevParams prm, int nCells, const double* __restrict__ acc,
                                                           const double* __restrict__ accS, const double* __restrict__ vol,
                                                           const double* __restrict__ bbMin, const double* __restrict__ bbMax,
                                                           const int* __restrict__ subLevels, double t, double nAvSteps, double* __restrict__ out,
                                                           const double* __restrict__ accI) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const double* A = acc + (size_t)c * NACC;
    double F[UGF_NFIELD];
#pragma unroll
    for (int k = 0; k < UGF_NFIELD; ++k) F[k] = 0;
    const double V = vol[c];
    if (A[0] > VSMALL) {
        F[0] = A[0] / t;
        F[1] = A[8] / (t * V);
        F[2] = A[9] / (t * V);
        const double rhoMMean = A[9] / (V * t);
        for (int k = 0; k < 3; ++k) F[3 + k] = A[10 + k] / (rhoMMean * V * t);
        const double linearKEMean = 0.5 * A[13] / (V * t);
        const double rhoNMean = A[8] / (V * t);
        F[6] = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * (F[3] * F[3] + F[4] * F[4] + F[5] * F[5]));
        F[9] = F[1] * kB * F[6];
    } else {
        F[0] = 0.001;  // uniGasVolFields.C:907-908
    }
    if (A[7] > VSMALL && t > VSMALL) F[7] = (2.0 / kB) * ((A[6] / t) / (A[7] / t));
    double nRotDof = 0;
    if (A[0] > VSMALL) nRotDof = A[7] / A[0];
    double totalvDof = 0, totalEDof = 0;
    if (accI) derive_internal(prm, accI + (size_t)c * prm.nSpecies * UGF_NINT, A[14], A[0], F[19], F[20], totalvDof, totalEDof);
    F[8] = (3.0 * F[6] + nRotDof * F[7] + totalvDof * F[19] + totalEDof * F[20]) / (3.0 + nRotDof + totalvDof + totalEDof);
    double gamma = 0, Cv_p = 0;
    if (A[0] > VSMALL) {
        const double molecularMass = A[1] / A[0];
        const double Cp = A[15] / A[0], Cv = Cp - 2.0;
        Cv_p = Cv / NAvo;
        gamma = Cp / Cv;
        if (F[6] > VSMALL && molecularMass > VSMALL) {
            const double cs = sqrt(gamma * (kB / molecularMass) * F[6]);
            F[10] = sqrt(F[3] * F[3] + F[4] * F[4] + F[5] * F[5]) / cs;
        }
    }
    if (F[0] > VSMALL && F[10] > VSMALL && gamma > VSMALL && Cv_p > VSMALL) {
        F[11] = 1.0 / sqrt(F[0] * nAvSteps);                                         // densityError (:1248)
        F[17] = (1.0 / sqrt(F[0] * nAvSteps)) * (1.0 / (F[10] * sqrt(gamma)));       // velocityError
        F[18] = (1.0 / sqrt(F[0] * nAvSteps)) * sqrt(kB / Cv_p);                     // temperatureError
    }
    // mean free path / collision rate fields (uniGasVolFields.C:1124-1232)
    {
        const int nS = prm.nSpecies;
        double MFP = 0, MCR = 0;
        for (int i = 0; i < nS; ++i) {
            double mfpI = 0, mcrI = 0;
            for (int q = 0; q < nS; ++q) {
                const double aq = accS ? accS[(size_t)c * nS + q] : A[8];
                const double dPQ = 0.5 * (prm.sp[i].d + prm.sp[q].d), omegaPQ = 0.5 * (prm.sp[i].omega + prm.sp[q].omega);
                const double massRatio = prm.sp[i].mass / prm.sp[q].mass;
                if (aq > VSMALL && F[6] > VSMALL) {
                    const double nDensQ = aq / (V * t);
                    const double reducedMass = prm.sp[i].mass * prm.sp[q].mass / (prm.sp[i].mass + prm.sp[q].mass);
                    mfpI += PI * dPQ * dPQ * nDensQ * pow(prm.Tref / F[6], omegaPQ - 0.5) * sqrt(1.0 + massRatio);
                    mcrI += 2.0 * sqrt(PI) * dPQ * dPQ * nDensQ * pow(F[6] / prm.Tref, 1.0 - omegaPQ) * sqrt(2.0 * kB * prm.Tref / reducedMass);
                }
            }
            if (mfpI > VSMALL) mfpI = 1.0 / mfpI;
            if (F[1] > VSMALL) {
                const double nDensP = (accS ? accS[(size_t)c * nS + i] : A[8]) / (V * t);
                MFP += mfpI * nDensP / F[1];
                MCR += mcrI * nDensP / F[1];
            }
        }
        if (MFP < VSMALL) MFP = GREAT;
        F[12] = MFP;
        F[14] = MCR;
        if (MCR > VSMALL) { F[15] = 1.0 / MCR; F[16] = prm.deltaT / F[15]; } else { F[15] = GREAT; F[16] = GREAT; }
        double largest = 0.0;
        for (int d = 0; d < 3; ++d) {
            const double dim = (bbMax[3 * (size_t)c + d] - bbMin[3 * (size_t)c + d]) / (subLevels ? subLevels[3 * (size_t)c + d] : 1);
            if (prm.solD[d] && largest < dim) largest = dim;
        }
        F[13] = largest / MFP;
    }
#pragma unroll
    for (int k = 0; k < UGF_NFIELD; ++k) out[(size_t)c * UGF_NFIELD + k] = F[k];
}

__global__ void __launch_bounds__(256) derive_walls_kernel(const __grid_constant__ DevParams prm, MeshDev mesh, const double* __restrict__ bacc,
                                                           const double* __restrict__ bfS /* [nBFaces*3] */, double t, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mesh.nBFaces) return;
    double F[UGF_NWALLFIELD];
#pragma unroll
    for (int k = 0; k < UGF_NWALLFIELD; ++k) F[k] = 0;
    const int patch = mesh.bfPatch[b];
    if (mesh.patches[patch].kind == UGF_PATCH_WALL) {
        const double* B = bacc + (size_t)b * UGF_NBM;
        double nPart = cell_fn(prm, mesh.bfOwner[b]);  // uniGasVolFields.C:1276-1278: CWF of the boundary cell, RWF of the face centre
        if (prm.axi) nPart = nPart * __ldg(&prm.bfRwf[b]);
        if (B[0] > VSMALL) {
            F[0] = B[0] * nPart / t;
            F[1] = B[1] * nPart / t;
            for (int k = 0; k < 3; ++k) F[2 + k] = B[3 + k] * nPart / (F[1] * t);
            const double rhoMMean = B[1] * nPart / t, linearKEMean = B[2] * nPart / t, rhoNMean = B[0] * nPart / t;
            F[5] = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * (F[2] * F[2] + F[3] * F[3] + F[4] * F[4]));
        }
        F[6] = B[8] / t;
        for (int k = 0; k < 3; ++k) F[7 + k] = B[9 + k] / t;
        const double sx = bfS[3 * (size_t)b], sy = bfS[3 * (size_t)b + 1], sz = bfS[3 * (size_t)b + 2];
        const double fA = sqrt(sx * sx + sy * sy + sz * sz);
        const double nw[3] = {sx / fA, sy / fA, sz / fA};
        F[10] = F[7] * nw[0] + F[8] * nw[1] + F[9] * nw[2];
        const double ft[3] = {F[7] - F[10] * nw[0], F[8] - F[10] * nw[1], F[9] - F[10] * nw[2]};
        F[11] = sqrt(ft[0] * ft[0] + ft[1] * ft[1] + ft[2] * ft[2]);
    }
#pragma unroll
    for (int k = 0; k < UGF_NWALLFIELD; ++k) out[(size_t)b * UGF_NWALLFIELD + k] = F[k];
}

// info(): sum 0.5 m |U|^2, sum ERot, sum m U, live count -> tot[0..5]
template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(256) totals_kernel(const __grid_constant__ DevParams prm, ParcelBuf P, const long long* dN, double* tot) {
    __shared__ double sm[8][6];
    const long long n = *dN;
    double ke = 0, er = 0, mx = 0, my = 0, mz = 0, cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (P.cell[i] < 0) continue;
        const double m = prm.sp[MULTI ? P.type[i] : 0].mass;
        const double u = P.ux[i], v = P.uy[i], w = P.uz[i];
        ke += 0.5 * m * (u * u + v * v + w * w);
        if (HAS_ROT) er += P.erot[i];
        mx += m * u; my += m * v; mz += m * w;
        cnt += 1.0;
    }
    ke = warp_sum(ke); er = warp_sum(er); mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz); cnt = warp_sum(cnt);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sm[wid][0] = ke; sm[wid][1] = er; sm[wid][2] = mx; sm[wid][3] = my; sm[wid][4] = mz; sm[wid][5] = cnt; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
        atomicAdd(&tot[threadIdx.x], s);
    }
}

}  // namespace ugf
