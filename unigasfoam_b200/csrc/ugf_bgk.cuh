// ugf_bgk.cuh — stochastic-particle BGK-family relaxation, one warp per cell.
//
// Replaces bgkCollisionModel::collide for
//   stochasticParticleBGK          (U/bgkCollisions/derived/stochasticParticleBGK/stochasticParticleBGK.C:628-796)
//   stochasticParticleESBGK        (…/stochasticParticleESBGK/stochasticParticleESBGK.C:887-906, nu = Pr p/mu :641)
//   stochasticParticleSBGK         (…/stochasticParticleSBGK/stochasticParticleSBGK.C:1009-1046)
//   unifiedStochasticParticleSBGK  (…/unifiedStochasticParticleSBGK/unifiedStochasticParticleSBGK.C:379-1096)
// i.e. calculateProperties (cell macroscopic state from the pre-collision moments), selection of the relaxing
// parcels, sampling of the target distribution, conserveMomentumAndEnergy and resetProperties.
//
// The cell's velocities are staged in shared memory; the macroscopic state is computed redundantly by all lanes
// from the 256-byte moment block; the acceptance-rejection envelope (maxProb) follows the reference's
// sequential semantics: lanes sample speculatively, and everything after the first lane that raised the
// envelope is re-sampled against the raised value.
//
// Deviation from the reference, statistically equivalent: the relaxing subset is the nRel smallest of one
// uniform key per parcel instead of "shuffle the cell list five times and take the first nRel" — both are a
// uniformly random nRel-subset.
#pragma once
#include "ugf_common.cuh"
#include "ugf_rng.cuh"

namespace ugf {

constexpr int BGK_THREADS = 256;
constexpr int BGK_WARPS = BGK_THREADS / 32;

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
    int cap;
};

struct Macro {
    bool perform;
    double N, rhoN, p, T, U[3], q[3], s[6], P[6], Pr, nu, rhoNX, rhoMX;
};

// calculateProperties for one cell (…USP.C:379-800), all lanes redundantly; lane 0 stores the blended
// heat flux / shear stress for the next step (…USP.C:777-783).
__device__ inline void bgk_macro(const DevParams& prm, const double* __restrict__ momCell, double V, int lane,
                                 double* qPrevCell, double* sPrevCell, Macro& m) {
    const int nS = prm.nSpecies;
    const int model = prm.bgkModel;
    const double FN = prm.nParticle;
    double N = 0, rhoM = 0, rhoNX = 0, rhoMX = 0, momX[3] = {0, 0, 0}, keX = 0;
    double muu[6] = {0, 0, 0, 0, 0, 0}, mcc = 0, mccu[3] = {0, 0, 0}, eInt = 0, eIntU[3] = {0, 0, 0};
    double nSp[UGF_MAX_SPECIES];
#pragma unroll
    for (int s = 0; s < UGF_MAX_SPECIES; ++s) nSp[s] = 0;
    for (int s = 0; s < nS; ++s) {
        const double mv = momCell[(size_t)s * UGF_NMOM + lane];
        const double ms = prm.sp[s].mass;
        const double a0 = __shfl_sync(0xffffffffu, mv, 0), a1 = __shfl_sync(0xffffffffu, mv, 1);
        N += a0; rhoM += ms * a0;
        nSp[s] = a0;
        rhoNX += a1 * FN; rhoMX += ms * a1 * FN;
#pragma unroll
        for (int k = 0; k < 3; ++k) momX[k] += ms * __shfl_sync(0xffffffffu, mv, 5 + k) * FN;
        keX += ms * __shfl_sync(0xffffffffu, mv, 14) * FN;
        double uu[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { uu[k] = __shfl_sync(0xffffffffu, mv, 8 + k); muu[k] += ms * uu[k]; }
        mcc += ms * (uu[0] + uu[3] + uu[5]);
#pragma unroll
        for (int k = 0; k < 3; ++k) mccu[k] += ms * __shfl_sync(0xffffffffu, mv, 15 + k);
        eInt += __shfl_sync(0xffffffffu, mv, 18) + __shfl_sync(0xffffffffu, mv, 22);
#pragma unroll
        for (int k = 0; k < 3; ++k) eIntU[k] += __shfl_sync(0xffffffffu, mv, 19 + k) + __shfl_sync(0xffffffffu, mv, 23 + k);
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
        double visc = 0, Pr = 0;
        for (int s = 0; s < nS; ++s) {
            const DevSpecies& S = prm.sp[s];
            const double al = S.alpha;
            const double viscRef = 1.25 * (1.0 + al) * (2.0 + al) * sqrt(S.mass * kB * prm.Tref)
                                   / (al * (5.0 - 2.0 * S.omega) * (7.0 - 2.0 * S.omega) * sqrt(PI) * (S.d * S.d));
            visc += nSp[s] * viscRef * pow(m.T / prm.Tref, S.omega);
            Pr += nSp[s] * (5.0 + S.rotDoF) / (7.5 + S.rotDoF);
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
        }
        __syncwarp();
        if (lane == 0) for (int k = 0; k < 3; ++k) qPrevCell[k] = m.q[k];
    }
    if (model == UGF_BGK_USP_SBGK) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            m.s[k] = th * m.s[k] / (1.0 + 0.5 * m.nu * dt) + (1.0 - th) * sPrevCell[k];
        }
        __syncwarp();
        if (lane == 0) for (int k = 0; k < 6; ++k) sPrevCell[k] = m.s[k];
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

template <bool MULTI>
__global__ void __launch_bounds__(BGK_THREADS) bgk_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ BgkArgs a) {
    extern __shared__ double smemD[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int cap = a.cap;
    const int perWarp = cap * 4;
    double* sU0 = smemD + (size_t)wib * perWarp;
    double* sU1 = sU0 + cap;
    double* sU2 = sU1 + cap;
    double* sK = sU2 + cap;
    uint8_t* sT = reinterpret_cast<uint8_t*>(smemD + (size_t)BGK_WARPS * perWarp) + (size_t)wib * cap;
    const int warpsTotal = gridDim.x * BGK_WARPS;
    const int nS = prm.nSpecies;
    const int model = prm.bgkModel;
    const bool envelope = (model == UGF_BGK_SBGK || model == UGF_BGK_USP_SBGK);
    int myRel = 0;

    for (int cell = blockIdx.x * BGK_WARPS + wib; cell < a.nCells; cell += warpsTotal) {
        const int beg = a.off[cell];
        const int n = a.off[cell + 1] - beg;
        Macro m;
        bgk_macro(prm, a.mom + (size_t)cell * nS * UGF_NMOM, a.vol[cell], lane, a.qPrev + 3 * (size_t)cell, a.sPrev + 6 * (size_t)cell, m);
        bool raised = false;
        double E = envelope ? a.maxProb[cell] : 1.0;
        if (a.collModelId[cell] == 0 && m.perform) {
            const bool useSmem = n <= cap;
            double *pu0, *pu1, *pu2, *pk;
            uint8_t* pt;
            if (useSmem) {
                pu0 = sU0; pu1 = sU1; pu2 = sU2; pk = sK; pt = sT;
                for (int j = lane; j < n; j += 32) {
                    sU0[j] = a.P.ux[beg + j]; sU1[j] = a.P.uy[beg + j]; sU2[j] = a.P.uz[beg + j];
                    if (MULTI) sT[j] = a.P.type[beg + j];
                }
            } else {
                pu0 = a.P.ux + beg; pu1 = a.P.uy + beg; pu2 = a.P.uz + beg; pk = a.keyScratch + beg;
                pt = MULTI ? a.P.type + beg : nullptr;
            }
            // number of relaxing parcels (…USP.C:917-923)
            const double dt = prm.deltaT;
            const double pc = m.N * (1.0 - exp(-m.nu * dt));
            int nRel = (int)pc;
            {
                Stream rc(prm.seed, KIND_BGK, 0, a.step, (uint32_t)cell, 0xFFFFFFFFu);
                if (rc.u01() < (pc - nRel)) nRel++;
            }
            nRel = min(nRel, n);
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
                while (pending) {
                    const bool mine = (pending >> lane) & 1u;
                    bool rz = false;
                    double newE = E;
                    if (mine) {
                        Stream r(prm.seed, KIND_BGK, 0, a.step, (uint32_t)cell, (uint32_t)j);
                        (void)r.u01();  // the selection key
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
                        raised = true;
                    }
                    if ((commitMask >> lane) & 1u) {
                        pu0[j] = m.U[0] + u0 * v[0];
                        pu1[j] = m.U[1] + u0 * v[1];
                        pu2[j] = m.U[2] + u0 * v[2];
                        myRel++;
                    }
                }
            }
            __syncwarp();
            // conserveMomentumAndEnergy (…USP.C:996-1045)
            const double FN = prm.nParticle;
            double keX = 0, mx = 0, my = 0, mz = 0;
            for (int j = lane; j < n; j += 32) {
                const double mass = MULTI ? prm.sp[pt[j]].mass : prm.sp[0].mass;
                const double u = pu0[j], vv = pu1[j], w = pu2[j];
                keX += mass * (u * u + vv * vv + w * w) * FN;
                mx += mass * u * FN; my += mass * vv * FN; mz += mass * w * FN;
            }
            keX = warp_sum(keX); mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz);
            const double pU[3] = {mx / m.rhoMX, my / m.rhoMX, mz / m.rhoMX};
            const double postT = m.N / (3.0 * (m.N - 1.0) * kB * m.rhoNX) * (keX - m.rhoMX * (pU[0] * pU[0] + pU[1] * pU[1] + pU[2] * pU[2]));
            const bool rescale = postT > VSMALL;
            const double f = rescale ? sqrt(m.T / postT) : 1.0;
            if (rescale || useSmem) {
                for (int j = lane; j < n; j += 32) {
                    double u = pu0[j], vv = pu1[j], w = pu2[j];
                    if (rescale) {
                        u = m.U[0] + (u - pU[0]) * f;
                        vv = m.U[1] + (vv - pU[1]) * f;
                        w = m.U[2] + (w - pU[2]) * f;
                    }
                    a.P.ux[beg + j] = u; a.P.uy[beg + j] = vv; a.P.uz[beg + j] = w;
                }
            }
            __syncwarp();
        }
        // resetProperties: envelope decay (…USP.C:859-863)
        if (envelope && lane == 0) a.maxProb[cell] = raised ? E : E * (model == UGF_BGK_USP_SBGK ? 0.999 : 0.9999);
    }
    const int wr = warp_sum_int(myRel);
    if (lane == 0 && wr) atomicAdd(&a.cnt->bgk, (unsigned long long)wr);
}

}  // namespace ugf
