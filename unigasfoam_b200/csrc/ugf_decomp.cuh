// ugf_decomp.cuh — hybrid DSMC / BGK decomposition by the gradient-length local Knudsen number.
//
// Replaces localKnudsen::decompose (U/hybridDecomposition/derived/localKnudsen/localKnudsen.C:216-582): time averages
// of the cell sums over the decomposition interval, macroscopic fields, `smoothingPasses` applications of
// fvc::average(fvc::interpolate(.)) with zero-gradient / symmetry / cyclic boundary faces, maximum neighbour
// gradients, Bird's mean free path (eqs 4.76/4.77), the theta-blended Knudsen fields and the breakdown threshold.
// Everything here is one thread per cell over arrays of doubles (HBM-bound, run once per decompositionInterval
// except the accumulation).  The three sequential, in-place refinement sweeps of the reference (:428-541) are
// order-dependent by construction and run on the host (ugf_api.cu: refine_mask) on the downloaded mask.
#pragma once
#include "ugf_common.cuh"

namespace ugf {

constexpr int KN_NACC = 7;   // 0 N, 1 rhoN X, 2 rhoM X, 3 linearKE X, 4-6 momentum X; then nParcels X per species
constexpr int KN_NF = 7;     // derived fields per cell: 0 rhoN, 1 rhoM, 2 p, 3 T, 4-6 U
enum { DF_OWNER = 0, DF_NEIGHBOUR = 1, DF_SYMMETRY = 2, DF_BOUNDARY = 3 };  // internal face seen from its owner / neighbour; boundary faces

// cell -> face slots of the smoothing operator (empty faces left out, cell-face order kept)
struct DecompGeom {
    int nCells;
    const int* off;      // [nCells+1]
    const int* nb;       // [slots] cell on the other side (internal, cyclic) or the cell itself (zero gradient, symmetry)
    const int* info;     // [slots] DF_* | boundaryFace << 2 (boundary kinds)
    const double* w;     // [slots] linear weight of the owner side (boundary faces: of this side; 1 = zero gradient)
    const double* A;     // [slots] |Sf|
    const double* bfS;   // [nBFaces*3] boundary face area vectors (symmetry mirror)
    const double* cc;    // [nCells*3] cell centres
    const double* vol;
};

__global__ void __launch_bounds__(256) kn_accumulate_kernel(const __grid_constant__ DevParams prm, int nCells, const double* __restrict__ mom,
                                                            double* __restrict__ acc) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const int nS = prm.nSpecies, W = KN_NACC + nS;
    const double dt = prm.deltaT, FN = cell_fn(prm, c);
    double* a = acc + (size_t)c * W;
    double v[KN_NACC];
#pragma unroll
    for (int k = 0; k < KN_NACC; ++k) v[k] = a[k];
    for (int s = 0; s < nS; ++s) {
        const double* m = mom + ((size_t)c * nS + s) * UGF_NMOM;
        const double ms = prm.sp[s].mass;
        const double m1 = m[1];
        v[0] += dt * m[0];
        v[1] += dt * (m1 * FN);
        v[2] += dt * (ms * m1 * FN);
        v[3] += dt * (ms * m[prm.axi ? 31 : 14] * FN);
#pragma unroll
        for (int k = 0; k < 3; ++k) v[4 + k] += dt * (ms * m[5 + k] * FN);
        a[KN_NACC + s] += dt * (m1 * FN);
    }
#pragma unroll
    for (int k = 0; k < KN_NACC; ++k) a[k] = v[k];
}

__global__ void __launch_bounds__(256) kn_derive_kernel(int nCells, int W, const double* __restrict__ acc, const double* __restrict__ vol, double tAv,
                                                        double* __restrict__ F) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const double* a = acc + (size_t)c * W;
    double f[KN_NF] = {0, 0, 0, 0, 0, 0, 0};
    if (a[0] > VSMALL) {
        const double V = vol[c];
        f[0] = a[1] / (tAv * V);
        f[1] = a[2] / (tAv * V);
        const double rhoMMean = a[2] / (V * tAv);
        for (int k = 0; k < 3; ++k) f[4 + k] = a[4 + k] / (rhoMMean * V * tAv);
        const double linearKEMean = 0.5 * a[3] / (V * tAv);
        const double rhoNMean = a[1] / (V * tAv);
        f[3] = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * (f[4] * f[4] + f[5] * f[5] + f[6] * f[6]));
        f[2] = f[0] * kB * f[3];
    }
    for (int k = 0; k < KN_NF; ++k) F[(size_t)c * KN_NF + k] = f[k];
}

// out = fvc::average(fvc::interpolate(in)) for components [first, NC) of an [nCells][NC] array; components below
// `first` are copied.  VEC >= 0: components VEC..VEC+2 form a vector (mirrored on symmetry planes).
template <int NC, int VEC>
__global__ void __launch_bounds__(256) kn_smooth_kernel(const DecompGeom g, int first, const double* __restrict__ in, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.nCells) return;
    double fc[NC], num[NC], den = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) { fc[k] = in[(size_t)c * NC + k]; num[k] = 0; }
    for (int j = g.off[c]; j < g.off[c + 1]; ++j) {
        const int q = g.nb[j], info = g.info[j];
        const double w = g.w[j], A = g.A[j];
        double val[NC];
        if ((info & 3) == DF_NEIGHBOUR) {
#pragma unroll
            for (int k = 0; k < NC; ++k) val[k] = w * in[(size_t)q * NC + k] + (1.0 - w) * fc[k];
        } else {
#pragma unroll
            for (int k = 0; k < NC; ++k) val[k] = w * fc[k] + (1.0 - w) * in[(size_t)q * NC + k];
            if (VEC >= 0 && (info & 3) == DF_SYMMETRY) {
                const double* S = g.bfS + 3 * (size_t)(info >> 2);
                const double n[3] = {S[0] / A, S[1] / A, S[2] / A};
                const double vn = fc[VEC >= 0 ? VEC : 0] * n[0] + fc[VEC >= 0 ? VEC + 1 : 0] * n[1] + fc[VEC >= 0 ? VEC + 2 : 0] * n[2];
#pragma unroll
                for (int k = 0; k < 3; ++k) val[(VEC >= 0 ? VEC : 0) + k] = fc[(VEC >= 0 ? VEC : 0) + k] - vn * n[k];
            }
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) num[k] += A * val[k];
        den += A;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) out[(size_t)c * NC + k] = (k < first) ? fc[k] : num[k] / den;
}

// maximum neighbour gradients, mean free path, instantaneous Knudsen numbers, theta blend (localKnudsen.C:294-397)
__global__ void __launch_bounds__(256) kn_kernel(const __grid_constant__ DevParams prm, const DecompGeom g, int W, const double* __restrict__ acc,
                                                 const double* __restrict__ F, double breakdownMax, double theta, double* __restrict__ K) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.nCells) return;
    const int nS = prm.nSpecies;
    double f[KN_NF];
#pragma unroll
    for (int k = 0; k < KN_NF; ++k) f[k] = F[(size_t)c * KN_NF + k];
    const double* a = acc + (size_t)c * W;
    double gRho = 0, gT = 0, gU = 0;
    const double magU = sqrt(f[4] * f[4] + f[5] * f[5] + f[6] * f[6]);
    const double c0 = g.cc[3 * (size_t)c], c1 = g.cc[3 * (size_t)c + 1], c2 = g.cc[3 * (size_t)c + 2];
    for (int j = g.off[c]; j < g.off[c + 1]; ++j) {
        if ((g.info[j] & 3) > DF_NEIGHBOUR) continue;  // mesh.cellCells(): internal faces only
        const int q = g.nb[j];
        const double* h = F + (size_t)q * KN_NF;
        const double d0 = g.cc[3 * (size_t)q] - c0, d1 = g.cc[3 * (size_t)q + 1] - c1, d2 = g.cc[3 * (size_t)q + 2] - c2;
        const double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        gRho = fmax(gRho, fabs(h[1] - f[1]) / dist);
        gT = fmax(gT, fabs(h[3] - f[3]) / dist);
        gU = fmax(gU, fabs(sqrt(h[4] * h[4] + h[5] * h[5] + h[6] * h[6]) - magU) / dist);
    }
    double knRho, knT, knU, knG;
    if (a[0] > VSMALL && f[3] > VSMALL) {
        const double V = g.vol[c];
        double mfp = 0;
        for (int i = 0; i < nS; ++i) {
            double inv = 0;
            for (int q = 0; q < nS; ++q) {
                const double dPQ = 0.5 * (prm.sp[i].d + prm.sp[q].d), omegaPQ = 0.5 * (prm.sp[i].omega + prm.sp[q].omega);
                const double massRatio = prm.sp[i].mass / prm.sp[q].mass;
                if (a[KN_NACC + q] > VSMALL) {
                    const double nDensQ = a[KN_NACC + q] / V;
                    inv += PI * dPQ * dPQ * nDensQ * pow(prm.Tref / f[3], omegaPQ - 0.5) * sqrt(1.0 + massRatio);  // Bird 4.76
                }
            }
            if (a[KN_NACC + i] > VSMALL) mfp += (1.0 / inv) * a[KN_NACC + i] / (f[0] * V);  // Bird 4.77
        }
        const double u0 = sqrt(2.0 * kB / (f[1] / f[0]) * f[3]);
        knRho = mfp * gRho / f[1];
        knT = mfp * gT / f[3];
        knU = mfp * gU / fmax(magU, u0);
        knG = fmax(fmax(knRho, knT), knU);
    } else {
        knRho = knT = knU = knG = 2.0 * breakdownMax;
    }
    double* k4 = K + (size_t)c * 4;
    k4[0] = theta * knRho + (1.0 - theta) * k4[0];
    k4[1] = theta * knT + (1.0 - theta) * k4[1];
    k4[2] = theta * knU + (1.0 - theta) * k4[2];
    k4[3] = theta * knG + (1.0 - theta) * k4[3];
}

__global__ void __launch_bounds__(256) kn_threshold_kernel(int nCells, const double* __restrict__ K, double breakdownMax, int* __restrict__ collModelId) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nCells) collModelId[c] = K[(size_t)c * 4 + 3] > breakdownMax ? 1 : 0;
}

}  // namespace ugf
