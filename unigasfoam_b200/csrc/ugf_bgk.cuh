// ugf_bgk.cuh — stochastic-particle BGK-family relaxation.
//
// Replaces bgkCollisionModel::collide for
//   stochasticParticleBGK          (U/bgkCollisions/derived/stochasticParticleBGK/stochasticParticleBGK.C:628-796)
//   stochasticParticleESBGK        (…/stochasticParticleESBGK/stochasticParticleESBGK.C:887-906, nu = Pr p/mu :641)
//   stochasticParticleSBGK         (…/stochasticParticleSBGK/stochasticParticleSBGK.C:1009-1046)
//   unifiedStochasticParticleSBGK  (…/unifiedStochasticParticleSBGK/unifiedStochasticParticleSBGK.C:379-1096)
// i.e. calculateProperties (cell macroscopic state from the pre-collision moments), selection of the relaxing
// parcels, sampling of the target distribution, conserveMomentumAndEnergy and resetProperties.
//
// One warp per chunk of BGK_CHUNK consecutive cells, in place on the cell-major buffer:
//   A  lane c computes cell c's macroscopic state (no redundant work) and its number of relaxing parcels;
//   B  the chunk's velocities are staged in shared memory (one coalesced pass) and every parcel draws its
//      selection key from its own Philox stream (step, cell, slot);
//   C  the nRel smallest keys of each cell are selected (rank among the cell's keys) and compacted into one list
//      for the whole chunk, so that
//   D  the expensive part - Philox + Box-Muller + acceptance-rejection per relaxing parcel - runs with all 32 lanes
//      busy whatever the per-cell counts are.  The acceptance-rejection envelope (maxProb) follows the reference's
//      sequential semantics: lanes sample speculatively; within a cell, everything after the first lane that raised
//      the envelope is re-sampled against the raised value;
//   E  momentum / energy sums by 4 lanes per cell, F the conservation rescale and one coalesced write-back.
// Cells larger than the staging capacity take a whole-warp path working in global memory.
//
// HBM traffic: 24 B read + 24 B written per parcel of a BGK cell, 256 B x species + 100 B per cell.
//
// Deviation from the reference, statistically equivalent: the relaxing subset is the nRel smallest of one
// uniform key per parcel instead of "shuffle the cell list five times and take the first nRel" — both are a
// uniformly random nRel-subset.
#pragma once
#include "ugf_common.cuh"
#include "ugf_rng.cuh"
#include "ugf_sort.cuh"

namespace ugf {

constexpr int BGK_THREADS = 128;
constexpr int BGK_WARPS = BGK_THREADS / 32;
constexpr int BGK_CHUNK = 8;             // cells per warp chunk
constexpr int BGK_LPC = 32 / BGK_CHUNK;  // lanes per cell in the conservation sums
constexpr int BGK_CAP = 256;             // parcels staged per run of cells

// ---- macroInterpolation (collisionProperties.macroInterpolation true): interpolationCellPoint restated ----------------------------
// calculateProperties runs for every cell in its own launch (bgk_fields_kernel: the complete Macro per cell, kept for the relaxation
// kernel, and the NIF target values), bgk_points_kernel takes them to the mesh points (inverse-distance weights, ugf_cell_point), and
// the relaxation kernel evaluates the target state of every relaxing parcel at its position: linear in the tet (cell centre + three
// face points) that contains it (…USP.C:893-947; OpenFOAM volPointInterpolation / cellPointWeight).
constexpr int NIF = 22;  // 0 Pr, 1 nu, 2 p, 3 T, 4-6 U, 7-9 q, 10-15 shear stress, 16-21 pressure tensor

struct InterpDev {
    const double* points;    // [nPoints*3]
    const int* tetOff;       // [nCells+1]
    const int* tetPts;       // [nTets*3]
    const int* pcOff;        // [nPoints+1]
    const int* pc;
    const double* pw;
    const double* pnormal;   // [nPoints*3]
    const double* cc;        // [nCells*3] cell centres
    double* cellF;           // [nCells*NIF]
    double* pointF;          // [nPoints*NIF]
    int nPoints;
};

struct BgkArgs {
    int nCells;
    const int* off;
    ParcelBuf P;  // cell-major, in place
    const double* mom;
    const double* vol;
    const int* collModelId;
    double* maxProb;
    double* qPrev;
    double* sPrev;
    double* keyScratch;  // [capacity] selection keys for cells larger than the staging capacity
    uint32_t step;
    DevCounters* cnt;
    struct Macro* macroCell;  // macroInterpolation: the cells' macroscopic state computed by bgk_fields_kernel, else null
    InterpDev ip;
};

struct alignas(16) Macro {
    bool perform;
    double N, rhoN, p, T, U[3], q[3], s[6], P[6], Pr, nu, rhoNX, rhoMX;
};

// calculateProperties for one cell (…USP.C:379-800), by one thread; it also stores the blended heat flux /
// shear stress for the next step (…USP.C:777-783).
__device__ inline void bgk_macro(const DevParams& prm, double FN, const double* __restrict__ momCell, double V, double* qPrevCell, double* sPrevCell,
                                 Macro& m) {
    const int nS = prm.nSpecies;
    const int model = prm.bgkModel;
    double N = 0, rhoM = 0, rhoNX = 0, rhoMX = 0, momX[3] = {0, 0, 0}, keX = 0;
    double muu[6] = {0, 0, 0, 0, 0, 0}, mcc = 0, mccu[3] = {0, 0, 0}, eInt = 0, eIntU[3] = {0, 0, 0};
    double visc = 0, Pr = 0;
    for (int s = 0; s < nS; ++s) {
        const double* mv = momCell + (size_t)s * UGF_NMOM;
        const double ms = prm.sp[s].mass;
        const double a0 = mv[0], a1 = mv[1];
        N += a0; rhoM += ms * a0;
        rhoNX += a1 * FN; rhoMX += ms * a1 * FN;
#pragma unroll
        for (int k = 0; k < 3; ++k) momX[k] += ms * mv[5 + k] * FN;
        keX += ms * mv[prm.axi ? 31 : 14] * FN;  // axisymmetric: the RWF-weighted sum (axi_moments_kernel)
        double uu[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { uu[k] = mv[8 + k]; muu[k] += ms * uu[k]; }
        mcc += ms * (uu[0] + uu[3] + uu[5]);
#pragma unroll
        for (int k = 0; k < 3; ++k) mccu[k] += ms * mv[15 + k];
        eInt += mv[18] + mv[22];
#pragma unroll
        for (int k = 0; k < 3; ++k) eIntU[k] += mv[19 + k] + mv[23 + k];
    }
    m.perform = true;
    m.N = N; m.rhoNX = rhoNX; m.rhoMX = rhoMX;
#pragma unroll
    for (int k = 0; k < 3; ++k) { m.U[k] = 0; m.q[k] = 0; }
#pragma unroll
    for (int k = 0; k < 6; ++k) { m.s[k] = 0; m.P[k] = 0; }
    m.rhoN = m.p = m.T = m.Pr = m.nu = 0;
    if (N > VSMALL) {
        m.rhoN = rhoNX / V;
        const double rhoMMean = rhoMX / V;
#pragma unroll
        for (int k = 0; k < 3; ++k) m.U[k] = momX[k] / (rhoMMean * V);
        const double linearKEMean = 0.5 * keX / V;
        const double rhoNMean = rhoNX / V;
        const double UU = m.U[0] * m.U[0] + m.U[1] * m.U[1] + m.U[2] * m.U[2];
        m.T = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * UU);
        m.p = m.rhoN * kB * m.T;
        const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
        for (int k = 0; k < 6; ++k) m.P[k] = m.rhoN * (muu[k] / N - (rhoM / N) * m.U[ia[k]] * m.U[ib[k]]);
        const double sp = (m.P[0] + m.P[3] + m.P[5]) / 3.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) m.s[k] = m.P[k];
        m.s[0] -= sp; m.s[3] -= sp; m.s[5] -= sp;
        const double Pf[3][3] = {{m.P[0], m.P[1], m.P[2]}, {m.P[1], m.P[3], m.P[4]}, {m.P[2], m.P[4], m.P[5]}};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            m.q[k] = m.rhoN * (0.5 * (mccu[k] / N) - 0.5 * (mcc / N) * m.U[k] + eIntU[k] / N - (eInt / N) * m.U[k])
                     - Pf[k][0] * m.U[0] - Pf[k][1] * m.U[1] - Pf[k][2] * m.U[2];
        const bool third = (model == UGF_BGK_SBGK || model == UGF_BGK_USP_SBGK);
        if (third ? (N > 2.0) : (N > 1.0)) {
            const double f1 = N / (N - 1.0);
            m.p = f1 * m.p;
            m.T = f1 * m.T;
            if (model == UGF_BGK_ESBGK) for (int k = 0; k < 6; ++k) m.P[k] = f1 * m.P[k];
            if (third) { const double f3 = (N * N) / (N - 1.0) / (N - 2.0); for (int k = 0; k < 3; ++k) m.q[k] = f3 * m.q[k]; }
            if (model == UGF_BGK_USP_SBGK) for (int k = 0; k < 6; ++k) m.s[k] = f1 * m.s[k];
        } else {
            m.perform = false;
        }
    } else {
        m.perform = false;
    }
    if (m.T > VSMALL) {
        for (int s = 0; s < nS; ++s) {
            const DevSpecies& S = prm.sp[s];
            const double nSp = momCell[(size_t)s * UGF_NMOM];
            const double al = S.alpha;
            const double viscRef = 1.25 * (1.0 + al) * (2.0 + al) * sqrt(S.mass * kB * prm.Tref)
                                   / (al * (5.0 - 2.0 * S.omega) * (7.0 - 2.0 * S.omega) * sqrt(PI) * (S.d * S.d));
            visc += nSp * viscRef * pow(m.T / prm.Tref, S.omega);
            Pr += nSp * (5.0 + S.rotDoF) / (7.5 + S.rotDoF);
        }
        visc /= N; Pr /= N;
        m.Pr = Pr;
        m.nu = (model == UGF_BGK_ESBGK ? Pr : 1.0) * m.p / visc;
    } else {
        m.perform = false;
        m.Pr = 0; m.nu = 0;
    }
    const double th = prm.theta, dt = prm.deltaT;
    if (model == UGF_BGK_SBGK || model == UGF_BGK_USP_SBGK) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            m.q[k] = th * m.q[k] / (1.0 + 0.5 * m.Pr * m.nu * dt) + (1.0 - th) * qPrevCell[k];
            qPrevCell[k] = m.q[k];
        }
    }
    if (model == UGF_BGK_USP_SBGK) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            m.s[k] = th * m.s[k] / (1.0 + 0.5 * m.nu * dt) + (1.0 - th) * sPrevCell[k];
            sPrevCell[k] = m.s[k];
        }
    }
}

// samplePostCollisionVelocity for the selected model; returns v (in units of u0) and whether the
// acceptance-rejection envelope was raised (S-BGK / USP only).
__device__ inline bool bgk_sample(const DevParams& prm, Stream& r, const Macro& m, double u0, double& E, double v[3]) {
    const int model = prm.bgkModel;
    const double isq2 = sqrt(2.0);
    if (model == UGF_BGK_BGK) {
        double g0, g1, g2;
        r.gauss3(g0, g1, g2);
        v[0] = g0 / isq2; v[1] = g1 / isq2; v[2] = g2 / isq2;
        return false;
    }
    if (model == UGF_BGK_ESBGK) {
        double g[3];
        r.gauss3(g[0], g[1], g[2]);
        for (int k = 0; k < 3; ++k) g[k] = g[k] / isq2;
        const double f = 0.5 * (1 - m.Pr) / m.Pr;
        const double S[3][3] = {{1 - f * (m.P[0] / m.p - 1), -f * (m.P[1] / m.p), -f * (m.P[2] / m.p)},
                                {-f * (m.P[1] / m.p), 1 - f * (m.P[3] / m.p - 1), -f * (m.P[4] / m.p)},
                                {-f * (m.P[2] / m.p), -f * (m.P[4] / m.p), 1 - f * (m.P[5] / m.p - 1)}};
        for (int k = 0; k < 3; ++k) v[k] = S[k][0] * g[0] + S[k][1] * g[1] + S[k][2] * g[2];
        return false;
    }
    double coeffQ, coeffS = 0;
    if (model == UGF_BGK_SBGK) {
        coeffQ = 2.0 * (1.0 - m.Pr);
    } else {
        const double tau = 0.5 * m.nu * prm.deltaT;
        const double e = 1.0 + 2.0 / (exp(2.0 * tau) - 1.0);
        coeffQ = 2.0 * (1.0 - m.Pr * tau * e);
        coeffS = (1.0 - tau * e);
    }
    for (;;) {
        double g0, g1, g2;
        r.gauss3(g0, g1, g2);
        v[0] = g0 / isq2; v[1] = g1 / isq2; v[2] = g2 / isq2;
        const double vSq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        const double vTr = vSq / 3.0;
        double prob = 1.0 + coeffQ / (m.p * u0) * (m.q[0] * v[0] + m.q[1] * v[1] + m.q[2] * v[2]) * (vSq / 2.5 - 1.0);
        if (model == UGF_BGK_USP_SBGK)
            prob += coeffS / m.p * (m.s[0] * (v[0] * v[0] - vTr) + m.s[3] * (v[1] * v[1] - vTr) + m.s[5] * (v[2] * v[2] - vTr)
                                    + 2.0 * m.s[1] * v[0] * v[1] + 2.0 * m.s[2] * v[0] * v[2] + 2.0 * m.s[4] * v[1] * v[2]);
        if (prob > E && prob < 10.0) { E = prob; return true; }
        if (r.u01() < prob / E) return false;
    }
}

// number of relaxing parcels of a cell (…USP.C:917-923): stochastic rounding of N (1 - exp(-nu dt))
__device__ __forceinline__ int bgk_count(const DevParams& prm, uint32_t step, int cell, const Macro& m, int n) {
    const double pc = m.N * (1.0 - exp(-m.nu * prm.deltaT));
    int nRel = (int)pc;
    Stream rc(prm.seed, KIND_BGK, 0, step, (uint32_t)cell, 0xFFFFFFFFu);
    if (rc.u01() < (pc - nRel)) nRel++;
    return min(nRel, n);
}

__device__ __forceinline__ void macro_to_fields(const Macro& m, double* f) {
    f[0] = m.Pr; f[1] = m.nu; f[2] = m.p; f[3] = m.T;
    for (int k = 0; k < 3; ++k) { f[4 + k] = m.U[k]; f[7 + k] = m.q[k]; }
    for (int k = 0; k < 6; ++k) { f[10 + k] = m.s[k]; f[16 + k] = m.P[k]; }
}

__device__ inline void project_sym(double* t, const double* n) {  // symmetric tensor (xx,xy,xz,yy,yz,zz) -> P T P, P = I - n n
    const double T[3][3] = {{t[0], t[1], t[2]}, {t[1], t[3], t[4]}, {t[2], t[4], t[5]}};
    double Pm[3][3], A[3][3], B[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Pm[i][j] = (i == j ? 1.0 : 0.0) - n[i] * n[j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = 0; for (int k = 0; k < 3; ++k) A[i][j] += Pm[i][k] * T[k][j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { B[i][j] = 0; for (int k = 0; k < 3; ++k) B[i][j] += A[i][k] * Pm[k][j]; }
    t[0] = B[0][0]; t[1] = B[0][1]; t[2] = B[0][2]; t[3] = B[1][1]; t[4] = B[1][2]; t[5] = B[2][2];
}

// calculateProperties for every cell (one thread per cell): the Macro is kept for bgk_kernel, which must not evaluate it a second
// time (the theta-blend of q / sigma advances the stored previous values)
__global__ void __launch_bounds__(128) bgk_fields_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ BgkArgs a) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.nCells) return;
    Macro m;
    bgk_macro(prm, cell_fn(prm, cell), a.mom + (size_t)cell * prm.nSpecies * UGF_NMOM, a.vol[cell], a.qPrev + 3 * (size_t)cell, a.sPrev + 6 * (size_t)cell, m);
    a.macroCell[cell] = m;
    if (a.ip.cellF) {  // macroInterpolation: the fields the cell -> point interpolation takes
        double f[NIF];
        macro_to_fields(m, f);
        for (int k = 0; k < NIF; ++k) a.ip.cellF[(size_t)cell * NIF + k] = f[k];
    }
}

__global__ void __launch_bounds__(128) bgk_points_kernel(const __grid_constant__ InterpDev ip) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ip.nPoints) return;
    double f[NIF];
    for (int k = 0; k < NIF; ++k) f[k] = 0.0;
    for (int j = ip.pcOff[p]; j < ip.pcOff[p + 1]; ++j) {
        const double w = ip.pw[j];
        const double* cf = ip.cellF + (size_t)ip.pc[j] * NIF;
        for (int k = 0; k < NIF; ++k) f[k] += w * cf[k];
    }
    const double n[3] = {ip.pnormal[3 * (size_t)p], ip.pnormal[3 * (size_t)p + 1], ip.pnormal[3 * (size_t)p + 2]};
    if (n[0] != 0.0 || n[1] != 0.0 || n[2] != 0.0) {
        for (int v = 4; v <= 7; v += 3) {
            const double vn = f[v] * n[0] + f[v + 1] * n[1] + f[v + 2] * n[2];
            for (int k = 0; k < 3; ++k) f[v + k] -= vn * n[k];
        }
        project_sym(f + 10, n);
        project_sym(f + 16, n);
    }
    for (int k = 0; k < NIF; ++k) ip.pointF[(size_t)p * NIF + k] = f[k];
}

// target state at position x of cell `cell` (cellPointWeight::findTetrahedron + interpolationCellPoint::interpolate); out of line:
// only the macroInterpolation instantiation of the relaxation kernel calls it
__device__ __noinline__ void interpolate_macro(const InterpDev& ip, const double* __restrict__ vol, int cell, double x0, double x1, double x2, Macro& m) {
    const double cc[3] = {ip.cc[3 * (size_t)cell], ip.cc[3 * (size_t)cell + 1], ip.cc[3 * (size_t)cell + 2]};
    const double r[3] = {x0 - cc[0], x1 - cc[1], x2 - cc[2]};
    const double V = vol[cell];
    int best = -1;
    double bestMin = -1e300, bw0 = 1, bw1 = 0, bw2 = 0, bw3 = 0;
    for (int t = ip.tetOff[cell]; t < ip.tetOff[cell + 1]; ++t) {
        const int* tp = ip.tetPts + 3 * (size_t)t;
        double e[3][3];
        for (int q = 0; q < 3; ++q)
            for (int k = 0; k < 3; ++k) e[q][k] = ip.points[3 * (size_t)tp[q] + k] - cc[k];
        const double c12[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2], e[1][0] * e[2][1] - e[1][1] * e[2][0]};
        const double det = e[0][0] * c12[0] + e[0][1] * c12[1] + e[0][2] * c12[2];
        if (!(fabs(det / V) > SMALL)) continue;
        const double c20[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2], e[2][0] * e[0][1] - e[2][1] * e[0][0]};
        const double c01[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2], e[0][0] * e[1][1] - e[0][1] * e[1][0]};
        const double l1 = (r[0] * c12[0] + r[1] * c12[1] + r[2] * c12[2]) / det;
        const double l2 = (r[0] * c20[0] + r[1] * c20[1] + r[2] * c20[2]) / det;
        const double l3 = (r[0] * c01[0] + r[1] * c01[1] + r[2] * c01[2]) / det;
        const double l0 = 1.0 - l1 - l2 - l3;
        const double mn = fmin(fmin(l0, l1), fmin(l2, l3));
        if (mn > bestMin) { bestMin = mn; best = t; bw0 = l0; bw1 = l1; bw2 = l2; bw3 = l3; }
        if (mn + SMALL > 0) break;
    }
    const double* cf = ip.cellF + (size_t)cell * NIF;
    double f[NIF];
    if (best < 0) {
        for (int k = 0; k < NIF; ++k) f[k] = cf[k];
    } else {
        const int* tp = ip.tetPts + 3 * (size_t)best;
        const double* p0 = ip.pointF + (size_t)tp[0] * NIF;
        const double* p1 = ip.pointF + (size_t)tp[1] * NIF;
        const double* p2 = ip.pointF + (size_t)tp[2] * NIF;
        for (int k = 0; k < NIF; ++k) f[k] = bw0 * cf[k] + bw1 * p0[k] + bw2 * p1[k] + bw3 * p2[k];
    }
    m.Pr = f[0]; m.nu = f[1]; m.p = f[2]; m.T = f[3];
    for (int k = 0; k < 3; ++k) { m.U[k] = f[4 + k]; m.q[k] = f[7 + k]; }
    for (int k = 0; k < 6; ++k) { m.s[k] = f[10 + k]; m.P[k] = f[16 + k]; }
}

struct alignas(16) BgkWarpSmem {
    Macro macBuf[2][BGK_CHUNK];  // the chunk's macroscopic states; the next chunk's are copied in while this one is processed
    double u[3][BGK_CAP];
    unsigned long long key[BGK_CAP];  // selection key: the uniform's 53 random bits << 8 | slot in the run's cell-local order
    double E[BGK_CHUNK];
    double pU[BGK_CHUNK][3];
    double fscale[BGK_CHUNK];  // < 0: no rescale
    unsigned long long thr[BGK_CHUNK];  // parcels of the cell with key < thr relax
    int nRel[BGK_CHUNK];
    int raised[BGK_CHUNK];
    int active[BGK_CHUNK];
    int cb[BGK_CHUNK + 1];     // cell begins of the current run, relative to its first parcel
    unsigned short sel[BGK_CAP];
    unsigned char cellOf[BGK_CAP];
    unsigned char type[BGK_CAP];
};

// A cell larger than the staging capacity: the whole warp works on it in global memory (keys in keyScratch).
template <bool MULTI, bool INTERP>
__device__ __noinline__ void bgk_giant_cell(const DevParams& prm, const BgkArgs& a, int cell, int beg, int n, const Macro& m, int nRel, double& E,
                                            int& raisedOut, int& myRel, int lane) {
    double* pu0 = a.P.ux + beg; double* pu1 = a.P.uy + beg; double* pu2 = a.P.uz + beg; double* pk = a.keyScratch + beg;
    const uint8_t* pt = MULTI ? a.P.type + beg : nullptr;
    for (int j = lane; j < n; j += 32) {
        Stream r(prm.seed, KIND_BGK, 0, a.step, (uint32_t)cell, (uint32_t)j);
        pk[j] = r.u01();
    }
    __syncwarp();
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        bool sel = false;
        double mass = prm.sp[0].mass;
        if (j < n) {
            const double kj = pk[j];
            int rank = 0;
            for (int i = 0; i < n; ++i) {
                const double ki = pk[i];
                rank += (ki < kj) || (ki == kj && i < j);
            }
            sel = rank < nRel;
            if (MULTI) mass = prm.sp[pt[j]].mass;
        }
        const double u0 = sqrt(2.0 * kB * m.T / mass);
        unsigned pending = __ballot_sync(0xffffffffu, sel);
        double v[3] = {0, 0, 0};
        Macro mi;
        double u0i = u0;
        while (pending) {
            const bool mine = (pending >> lane) & 1u;
            bool rz = false;
            double newE = E;
            if (mine) {
                Stream r(prm.seed, KIND_BGK, 0, a.step, (uint32_t)cell, (uint32_t)j);
                (void)r.u01();  // the selection key
                if (INTERP) {
                    mi = m;
                    interpolate_macro(a.ip, a.vol, cell, a.P.x[beg + j], a.P.y[beg + j], a.P.z[beg + j], mi);
                    u0i = sqrt(2.0 * kB * mi.T / mass);
                    rz = bgk_sample(prm, r, mi, u0i, newE, v);
                } else
                rz = bgk_sample(prm, r, m, u0, newE, v);
            }
            const unsigned raisedMask = __ballot_sync(0xffffffffu, mine && rz);
            unsigned commitMask;
            if (!raisedMask) {
                commitMask = pending;
                pending = 0;
            } else {
                const int first = __ffs(raisedMask) - 1;
                const unsigned upto = (first == 31) ? 0xffffffffu : ((2u << first) - 1u);
                commitMask = pending & upto;
                pending &= ~upto;
                E = __shfl_sync(0xffffffffu, newE, first);
                raisedOut = 1;
            }
            if ((commitMask >> lane) & 1u) {
                pu0[j] = (INTERP ? mi.U[0] : m.U[0]) + u0i * v[0];
                pu1[j] = (INTERP ? mi.U[1] : m.U[1]) + u0i * v[1];
                pu2[j] = (INTERP ? mi.U[2] : m.U[2]) + u0i * v[2];
                myRel++;
            }
        }
    }
    __syncwarp();
    // conserveMomentumAndEnergy (…USP.C:996-1045)
    const double FN = cell_fn(prm, cell);
    double keX = 0, mx = 0, my = 0, mz = 0;
    for (int j = lane; j < n; j += 32) {
        const double mass = MULTI ? prm.sp[pt[j]].mass : prm.sp[0].mass;
        const double u = pu0[j], vv = pu1[j], w = pu2[j];
        const double wF = prm.axi ? FN * parcel_rwf(prm, cell, a.P.y[beg + j], a.P.z[beg + j]) : FN;  // CWF*RWF*nParticle (…USP.C:1017-1022)
        keX += mass * (u * u + vv * vv + w * w) * wF;
        mx += mass * u * wF; my += mass * vv * wF; mz += mass * w * wF;
    }
    keX = warp_sum(keX); mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz);
    const double pU[3] = {mx / m.rhoMX, my / m.rhoMX, mz / m.rhoMX};
    const double postT = m.N / (3.0 * (m.N - 1.0) * kB * m.rhoNX) * (keX - m.rhoMX * (pU[0] * pU[0] + pU[1] * pU[1] + pU[2] * pU[2]));
    if (postT > VSMALL) {
        const double f = sqrt(m.T / postT);
        for (int j = lane; j < n; j += 32) {
            pu0[j] = m.U[0] + (pu0[j] - pU[0]) * f;
            pu1[j] = m.U[1] + (pu1[j] - pU[1]) * f;
            pu2[j] = m.U[2] + (pu2[j] - pU[2]) * f;
        }
    }
    __syncwarp();
}

template <bool MULTI, bool INTERP = false>
__global__ void __launch_bounds__(BGK_THREADS) bgk_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ BgkArgs a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    BgkWarpSmem& S = reinterpret_cast<BgkWarpSmem*>(smemRaw)[wib];
    const int warpsTotal = gridDim.x * BGK_WARPS;
    const int model = prm.bgkModel;
    const bool envelope = (model == UGF_BGK_SBGK || model == UGF_BGK_USP_SBGK);
    const int nChunks = (a.nCells + BGK_CHUNK - 1) / BGK_CHUNK;
    int myRel = 0;

    // The per-chunk inputs (CSR offsets, model id, envelope, macroscopic states) of the NEXT chunk are requested while the current
    // one is processed: offsets / id / envelope into registers, the states by cp.async into the other half of macBuf.
    auto request = [&](int chunkP, int buf, int& offR, int& idR, double& eR) {
        offR = 0x7fffffff; idR = 1; eR = 1.0;
        if (chunkP >= nChunks) return;
        const int c0p = chunkP * BGK_CHUNK;
        const int ncp = min(BGK_CHUNK, a.nCells - c0p);
        if (lane <= ncp) offR = a.off[c0p + lane];
        if (lane < ncp) {
            idR = a.collModelId[c0p + lane];
            if (envelope) eR = a.maxProb[c0p + lane];
        }
        constexpr int PER = (int)(sizeof(Macro) / 16);
        const char* src = reinterpret_cast<const char*>(a.macroCell + c0p);
        char* dst = reinterpret_cast<char*>(&S.macBuf[buf][0]);
        for (int t = lane; t < ncp * PER; t += 32) cp_async16(dst + 16 * t, src + 16 * t);
    };
    int offR, idR;
    double eR;
    int cur = 0;
    request(blockIdx.x * BGK_WARPS + wib, cur, offR, idR, eR);

    for (int chunk = blockIdx.x * BGK_WARPS + wib; chunk < nChunks; chunk += warpsTotal, cur ^= 1) {
        const int c0 = chunk * BGK_CHUNK;
        const int nc = min(BGK_CHUNK, a.nCells - c0);
        const int offv = offR;
        const int myId = idR;
        double Eold = eR;
        cp_async_wait_all();
        __syncwarp();
        Macro* const mac = S.macBuf[cur];
        request(chunk + warpsTotal, cur ^ 1, offR, idR, eR);
        const int offNext = __shfl_down_sync(0xffffffffu, offv, 1);
        // ---- A: relaxing count, one cell per lane (calculateProperties ran in bgk_fields_kernel, one thread per cell) -----------
        if (lane < nc) {
            const int cell = c0 + lane;
            const Macro& m = mac[lane];
            const bool act = myId == 0 && m.perform;
            S.E[lane] = Eold;
            S.raised[lane] = 0;
            S.active[lane] = act ? 1 : 0;
            S.nRel[lane] = act ? bgk_count(prm, a.step, cell, m, offNext - offv) : 0;
        }
        __syncwarp();
        const unsigned activeMask = __ballot_sync(0xffffffffu, lane < nc && S.active[lane & (BGK_CHUNK - 1)] != 0);
        int done = 0;
        while (done < nc) {
            const int b0 = __shfl_sync(0xffffffffu, offv, done);
            const unsigned fit = __ballot_sync(0xffffffffu, lane > done && lane <= nc && (offv - b0) <= BGK_CAP);
            const int k = __popc(fit);
            if (k == 0) {  // a single cell larger than the staging buffer
                const int n = __shfl_sync(0xffffffffu, offv, done + 1) - b0;
                if ((activeMask >> done) & 1u) {
                    double E = S.E[done];
                    int rs = 0;
                    bgk_giant_cell<MULTI, INTERP>(prm, a, c0 + done, b0, n, mac[done], S.nRel[done], E, rs, myRel, lane);
                    __syncwarp();
                    if (lane == 0) { S.E[done] = E; if (rs) S.raised[done] = 1; }
                    __syncwarp();
                }
                done += 1;
                continue;
            }
            const unsigned runMask = ((k >= 32 ? 0u : (1u << k)) - 1u) << done;
            if (!(activeMask & runMask)) { done += k; continue; }
            const int ntot = __shfl_sync(0xffffffffu, offv, done + k) - b0;
            const int cbv = __shfl_sync(0xffffffffu, offv, (done + lane) & 31) - b0;
            if (lane <= k) S.cb[lane] = cbv;
            __syncwarp();
            // ---- B: stage velocities, cell slot of every parcel, selection keys ----------------------------------------
            for (int f = lane; f < ntot; f += 32) {  // all copies of the run in flight while the keys below are drawn
                cp_async8(&S.u[0][f], &a.P.ux[b0 + f]);
                cp_async8(&S.u[1][f], &a.P.uy[b0 + f]);
                cp_async8(&S.u[2][f], &a.P.uz[b0 + f]);
            }
            for (int f = lane; f < ntot; f += 32) {
                int g = 0;
#pragma unroll
                for (int t = 1; t < BGK_CHUNK; ++t) g += (t < k && S.cb[t] <= f) ? 1 : 0;
                S.cellOf[f] = (unsigned char)g;
                if (MULTI) S.type[f] = a.P.type[b0 + f];
                if (S.nRel[done + g] > 0) {
                    // ordering by (u01, index) == ordering by the uniform's 53-bit integer with the index appended
                    // (u01 = integer * 2^-53 is monotone; f - cb < BGK_CAP = 2^8): one integer compare per pair
                    Stream r(prm.seed, KIND_BGK, 0, a.step, (uint32_t)(c0 + done + g), (uint32_t)(f - S.cb[g]));
                    S.key[f] = (r.u53() << 8) | (unsigned long long)(f - S.cb[g]);
                }
            }
            cp_async_wait_all();
            __syncwarp();
            // ---- C: the nRel smallest keys of every cell, compacted over the run ---------------------------------------
            // BGK_LPC lanes per cell find the cell's threshold key by extracting minima (nRel of them) or maxima (N - nRel of
            // them), whichever is fewer: nRel N / BGK_LPC compares per cell instead of the N^2 of a rank count.  Keys are unique
            // within a cell (the slot index is part of the key).
            {
                const int g = lane / BGK_LPC, q = lane % BGK_LPC;
                const bool on = g < k;
                const int nRel = on ? S.nRel[done + g] : 0;
                const int cb = on ? S.cb[g] : 0, ce = on ? S.cb[g + 1] : 0;
                const int N = ce - cb;
                const bool fromBelow = nRel <= N - nRel;
                int rounds = (nRel <= 0 || nRel >= N) ? 0 : (fromBelow ? nRel : N - nRel);
                int maxRounds = rounds;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) maxRounds = max(maxRounds, __shfl_xor_sync(0xffffffffu, maxRounds, o));
                unsigned long long prev = 0ull;
                for (int rd = 0; rd < maxRounds; ++rd) {
                    unsigned long long best = fromBelow ? ~0ull : 0ull;
                    if (rd < rounds) {
                        for (int i = cb + q; i < ce; i += BGK_LPC) {
                            const unsigned long long kk = S.key[i];
                            if (fromBelow) { if ((rd == 0 || kk > prev) && kk < best) best = kk; }
                            else { if ((rd == 0 || kk < prev) && kk > best) best = kk; }
                        }
                    }
#pragma unroll
                    for (int o = 1; o < BGK_LPC; o <<= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                        best = fromBelow ? (other < best ? other : best) : (other > best ? other : best);
                    }
                    if (rd < rounds) prev = best;
                }
                if (on && q == 0) {
                    unsigned long long thr = 0ull;               // nRel == 0: nobody
                    if (nRel >= N && nRel > 0) thr = ~0ull;      // everybody
                    else if (rounds > 0) thr = fromBelow ? prev + 1ull : prev;  // key <= nRel-th smallest  /  key < (N - nRel)-th largest
                    S.thr[g] = thr;
                }
            }
            __syncwarp();
            int nSel = 0;
            for (int base = 0; base < ntot; base += 32) {
                const int f = base + lane;
                bool sel = false;
                if (f < ntot) {
                    const int g = S.cellOf[f];
                    if (S.nRel[done + g] > 0) sel = S.key[f] < S.thr[g];
                }
                const unsigned sm = __ballot_sync(0xffffffffu, sel);
                if (sel) S.sel[nSel + __popc(sm & ((1u << lane) - 1u))] = (unsigned short)f;
                nSel += __popc(sm);
            }
            __syncwarp();
            // ---- D: sample the target distribution for the selected parcels, all cells of the run together -----------
            for (int base = 0; base < nSel; base += 32) {
                const int t = base + lane;
                const bool mine0 = t < nSel;
                const int f = mine0 ? S.sel[t] : 0;
                const int g = S.cellOf[f];
                const int cl = done + g;
                const Macro& mc = mac[cl];
                const double mass = MULTI ? prm.sp[S.type[f]].mass : prm.sp[0].mass;
                Macro mi;
                if (INTERP && mine0) {  // target state at the parcel's position instead of the cell's (…USP.C:936-947)
                    mi = mc;
                    interpolate_macro(a.ip, a.vol, c0 + cl, a.P.x[b0 + f], a.P.y[b0 + f], a.P.z[b0 + f], mi);
                }
                const Macro& m = (INTERP && mine0) ? mi : mc;
                const double u0 = sqrt(2.0 * kB * m.T / mass);
                int head, cnt, rnk;
                warp_runs(mine0 ? g : -1, lane, head, cnt, rnk);  // the list is cell-major: a cell's lanes are adjacent
                const unsigned cellMask = (cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u)) << head;
                unsigned pending = __ballot_sync(0xffffffffu, mine0);
                double v[3] = {0, 0, 0};
                while (pending) {
                    const bool mine = (pending >> lane) & 1u;
                    bool rz = false;
                    double E = 1.0;
                    if (mine) {
                        Stream r(prm.seed, KIND_BGK, 0, a.step, (uint32_t)(c0 + cl), (uint32_t)(f - S.cb[g]));
                        (void)r.u01();  // the selection key
                        E = S.E[cl];
                        rz = bgk_sample(prm, r, m, u0, E, v);
                    }
                    const unsigned raisedMask = __ballot_sync(0xffffffffu, mine && rz) & cellMask;
                    const int firstR = raisedMask ? __ffs(raisedMask) - 1 : 32;
                    const bool ok = mine && lane <= firstR;  // nothing earlier in this cell raised the envelope
                    __syncwarp();
                    if (ok) {
                        if (rz) { S.E[cl] = E; S.raised[cl] = 1; }
                        S.u[0][f] = m.U[0] + u0 * v[0];
                        S.u[1][f] = m.U[1] + u0 * v[1];
                        S.u[2][f] = m.U[2] + u0 * v[2];
                        myRel++;
                    }
                    pending &= ~__ballot_sync(0xffffffffu, ok);
                    __syncwarp();
                }
            }
            __syncwarp();
            // ---- E: conserveMomentumAndEnergy (…USP.C:996-1045), BGK_LPC lanes per cell --------------------------------
            {
                const int g = lane / BGK_LPC, q = lane % BGK_LPC;
                const int cl = done + g;
                const bool cellOn = g < k && S.active[cl & (BGK_CHUNK - 1)] != 0;
                double keX = 0, mx = 0, my = 0, mz = 0;
                if (cellOn) {
                    const double FN = cell_fn(prm, c0 + cl);
                    const int cb = S.cb[g], ce = S.cb[g + 1];
                    for (int i = cb + q; i < ce; i += BGK_LPC) {
                        const double mass = MULTI ? prm.sp[S.type[i]].mass : prm.sp[0].mass;
                        const double u = S.u[0][i], vv = S.u[1][i], w = S.u[2][i];
                        const double wF = prm.axi ? FN * parcel_rwf(prm, c0 + cl, a.P.y[b0 + i], a.P.z[b0 + i]) : FN;
                        keX += mass * (u * u + vv * vv + w * w) * wF;
                        mx += mass * u * wF; my += mass * vv * wF; mz += mass * w * wF;
                    }
                }
#pragma unroll
                for (int o = 1; o < BGK_LPC; o <<= 1) {
                    keX += __shfl_xor_sync(0xffffffffu, keX, o);
                    mx += __shfl_xor_sync(0xffffffffu, mx, o);
                    my += __shfl_xor_sync(0xffffffffu, my, o);
                    mz += __shfl_xor_sync(0xffffffffu, mz, o);
                }
                if (cellOn && q == 0) {
                    const Macro& m = mac[cl];
                    const double pU[3] = {mx / m.rhoMX, my / m.rhoMX, mz / m.rhoMX};
                    const double postT = m.N / (3.0 * (m.N - 1.0) * kB * m.rhoNX) * (keX - m.rhoMX * (pU[0] * pU[0] + pU[1] * pU[1] + pU[2] * pU[2]));
                    S.pU[g][0] = pU[0]; S.pU[g][1] = pU[1]; S.pU[g][2] = pU[2];
                    S.fscale[g] = (postT > VSMALL) ? sqrt(m.T / postT) : -1.0;
                }
            }
            __syncwarp();
            // ---- F: rescale and write back (coalesced) -----------------------------------------------------------------
            for (int f = lane; f < ntot; f += 32) {
                const int g = S.cellOf[f];
                const int cl = done + g;
                if (!S.active[cl]) continue;
                double u = S.u[0][f], vv = S.u[1][f], w = S.u[2][f];
                const double fs = S.fscale[g];
                if (fs >= 0.0) {
                    const Macro& m = mac[cl];
                    u = m.U[0] + (u - S.pU[g][0]) * fs;
                    vv = m.U[1] + (vv - S.pU[g][1]) * fs;
                    w = m.U[2] + (w - S.pU[g][2]) * fs;
                }
                a.P.ux[b0 + f] = u; a.P.uy[b0 + f] = vv; a.P.uz[b0 + f] = w;
            }
            __syncwarp();
            done += k;
        }
        // resetProperties: envelope decay (…USP.C:859-863)
        if (envelope && lane < nc) a.maxProb[c0 + lane] = S.raised[lane] ? S.E[lane] : Eold * (model == UGF_BGK_USP_SBGK ? 0.999 : 0.9999);
        __syncwarp();
    }
    const int wr = warp_sum_int(myRel);
    if (lane == 0 && wr) atomicAdd(&a.cnt->bgk, (unsigned long long)wr);
}

}  // namespace ugf
