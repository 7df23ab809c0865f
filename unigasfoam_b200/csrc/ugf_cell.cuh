// ugf_cell.cuh — the per-cell kernel: payload gather through the occupancy permutation (the sort's scatter),
// cell-moment sampling and NTC collisions in one pass over each cell's parcels.
//
// Replaces, fused:  cellMeasurements::calculateFields (U/cellMeasurements/cellMeasurements.C:408-513),
//                   noTimeCounter::collide (U/dsmcCollisionPartner/derived/noTimeCounter/noTimeCounter.C:66-343),
//                   variableHardSphere / variableSoftSphere / LarsenBorgnakke* ::{sigmaTcR, collide}
//                   (U/dsmcCollisions/derived/*), and the kinetic samplers they call
//                   (U/clouds/uniGasCloud.C:1129-1189, 1267-1326).
//
// Two kernels, both one warp per run of consecutive cells (persistent CTAs; cell_kernel claims tasks of CELL_TASK
// cells from a global counter, ntc_kernel strides over chunks of 32 cells):
//   cell_kernel  streaming: the run's parcels are pulled through the permutation into shared memory with cp.async
//                (every load of the run in flight at once), moments are reduced by 4 lanes per cell and written
//                as one 256-byte block per (cell, species), and the run is written cell-major, fully coalesced.
//                Sampling sees the pre-collision state, as in the reference (uniGasCloud.C:846-850).
//   ntc_kernel   collisions, in place on the cell-major buffer: candidate counts for 8 cells at a time (one cell
//                per lane); only runs that drew candidates are staged.  The candidates of all cells of a run are
//                spread over the 32 lanes, each with its own Philox stream (step, cell, candidate#); a candidate
//                commits as soon as no earlier uncommitted candidate shares a parcel with it, which reproduces
//                the reference's sequential candidate loop exactly (same results as the oracle's serial loop).
// The split keeps the transcendental-heavy collision code (pow, sincos, Philox) out of the streaming kernel, whose
// register budget and shared-memory footprint decide how much HBM traffic can be kept in flight.
//
// HBM traffic: cell_kernel per parcel 4 (perm) + 56/64 (gather incl. cell id) read, 52/60 written, per cell
// 4 (offsets) + 256 x species (moments); ntc_kernel per cell 4 + 8 + 8 + 4, plus 24-32 B/parcel read and
// written for the runs that collide.
#pragma once
#include "ugf_common.cuh"
#include "ugf_rng.cuh"
#include "ugf_internal.cuh"

namespace ugf {

constexpr int CELL_THREADS = 256;
constexpr int CELL_WARPS = CELL_THREADS / 32;

struct CellArgs {
    int nCells;
    const int* off;
    const int* perm;  // null: identity (array already cell-major)
    ParcelBuf in, out;
    int gather;       // 1: write the cell-major copy into `out`; 0: sample only
    int doSample;
    int writeMom;     // 0: the moment blocks of staged cells are not stored (nothing downstream reads them: pure DSMC step)
    double* mom;
    double* acc;      // time-averaged accumulators [nCells][NACC]; updated in the same pass when accDt != 0
    double* accS;     // multi-species only: nParcelsXnParticle per species [nCells][nSpecies], else null
    double accDt;
    int* taskCounter; // zero at launch: next unclaimed task of taskCells cells
    int* taskReset;   // the counter of the next launch: zeroed here, so no memset sits between the kernels of a step
    long long* dN;    // gather: the array length becomes the live count *total (else null)
    const int* total;
    int taskCells;    // <= CELL_TASK_MAX
    int flags;        // tuning: 1 = positions pass through registers (L2 loads, no staging); 2 = velocities loaded with L2 loads
};

// postCollisionRotationalEnergy (U/clouds/uniGasCloud.C:1129-1189)
__device__ inline double post_collision_rotational_energy(Stream& r, int rotDoF, double ChiB) {
    double energyRatio = 0.0;
    if (rotDoF == 2) {
        energyRatio = 1.0 - pow(r.u01(), 1.0 / ChiB);
    } else {
        const double ChiA = 0.5 * rotDoF;
        const double A1 = ChiA - 1, B1 = ChiB - 1;
        if (A1 < SMALL && B1 < SMALL) return r.u01();
        double Pp;
        const double eps = r.u01();
        do {
            energyRatio = r.u01();
            if (A1 < SMALL) Pp = pow(1.0 - energyRatio, B1);
            else if (B1 < SMALL) Pp = pow(1.0 - energyRatio, A1);
            else Pp = pow((A1 + B1) * energyRatio / A1, A1) * pow((A1 + B1) * (1 - energyRatio) / B1, B1);
        } while (Pp < eps);
    }
    return energyRatio;
}

// postCollisionElectronicEnergyLevel for a single-level species (uniGasCloud.C:1267-1326): consumes the same
// draws as the reference loop and always returns level 0; Ec - E0 is the translational energy left.
__device__ inline void post_collision_electronic_single(Stream& r, double Ec, double E0, double omega) {
    for (;;) {
        (void)r.u01();  // nState draw
        if (Ec > E0) {
            const double prob = pow(1.0 - E0 / Ec, 1.5 - omega);
            if (prob > r.u01()) return;
        }
    }
}

__device__ __forceinline__ void scatter_vhs(Stream& r, double cR, double rel[3]) {
    const double cosTheta = 2.0 * r.u01() - 1.0;
    const double sinTheta = sqrt(1.0 - cosTheta * cosTheta);
    const double phi = TWO_PI * r.u01();
    double s, c;
    sincos(phi, &s, &c);
    rel[0] = cR * cosTheta;
    rel[1] = cR * (sinTheta * c);
    rel[2] = cR * (sinTheta * s);
}

// Bird eq 2.22 (variableSoftSphere.C:147-170)
__device__ __forceinline__ void scatter_vss(Stream& r, const double cRc[3], double alphaPQ, double scale, double rel[3]) {
    const double cR = sqrt(cRc[0] * cRc[0] + cRc[1] * cRc[1] + cRc[2] * cRc[2]);
    const double cosTheta = 2.0 * pow(r.u01(), 1.0 / alphaPQ) - 1.0;
    const double sinTheta = sqrt(1.0 - cosTheta * cosTheta);
    const double phi = TWO_PI * r.u01();
    const double D = sqrt(cRc[1] * cRc[1] + cRc[2] * cRc[2]);
    double sp, cp;
    sincos(phi, &sp, &cp);
    rel[0] = scale * (cosTheta * cRc[0] + sinTheta * sp * D);
    rel[1] = scale * (cosTheta * cRc[1] + sinTheta * (cR * cRc[2] * cp - cRc[0] * cRc[1] * sp) / D);
    rel[2] = scale * (cosTheta * cRc[2] - sinTheta * (cR * cRc[1] * cp + cRc[0] * cRc[2] * sp) / D);
}

// dsmcCollisionModel::collide for the selected model; UP/UQ/ERot are updated in place.
__device__ inline void collide_pair(const DevParams& prm, Stream& r, const DevSpecies& a, const DevSpecies& b,
                                    double UP[3], double UQ[3], double& erotP, double& erotQ) {
    const double mP = a.mass, mQ = b.mass, mS = mP + mQ;
    double Ucm[3], cRc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Ucm[k] = (mP * UP[k] + mQ * UQ[k]) / mS;
        cRc[k] = UP[k] - UQ[k];
    }
    const double cRsqr = cRc[0] * cRc[0] + cRc[1] * cRc[1] + cRc[2] * cRc[2];
    double rel[3];
    const int model = prm.binaryModel;
    if (model == UGF_BINARY_VHS) {
        scatter_vhs(r, sqrt(cRsqr), rel);
    } else if (model == UGF_BINARY_VSS) {
        scatter_vss(r, cRc, 0.5 * (a.alpha + b.alpha), 1.0, rel);
    } else {
        // Larsen-Borgnakke, serial application P then Q (LarsenBorgnakkeVariableHardSphere.C:125-416)
        const double omegaPQ = 0.5 * (a.omega + b.omega);
        const double mR = mP * mQ / mS;
        double Etr = 0.5 * mR * cRsqr;
        const double ChiB = 2.5 - omegaPQ;
        const double preERotP = erotP, preERotQ = erotQ;
        if (prm.invZel > r.u01()) {
            const double Ec = Etr + a.E0;
            post_collision_electronic_single(r, Ec, a.E0, omegaPQ);
            Etr = Ec - a.E0;
        }
        if (a.rotDoF > 0) {
            if (prm.invZrot > r.u01()) {
                const double Ec = Etr + preERotP;
                const double ratio = post_collision_rotational_energy(r, a.rotDoF, ChiB);
                erotP = ratio * Ec;
                Etr = Ec - erotP;
            }
        }
        if (prm.invZel > r.u01()) {
            const double Ec = Etr + b.E0;
            post_collision_electronic_single(r, Ec, b.E0, omegaPQ);
            Etr = Ec - b.E0;
        }
        if (b.rotDoF > 0) {
            if (prm.invZrot > r.u01()) {
                const double Ec = Etr + preERotQ;
                const double ratio = post_collision_rotational_energy(r, b.rotDoF, ChiB);
                erotQ = ratio * Ec;
                Etr = Ec - erotQ;
            }
        }
        const double cRnew = sqrt((2.0 * Etr) / mR);
        if (model == UGF_BINARY_LB_VHS) scatter_vhs(r, cRnew, rel);
        else scatter_vss(r, cRc, 0.5 * (a.alpha + b.alpha), cRnew / sqrt(cRsqr), rel);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        UP[k] = Ucm[k] + rel[k] * mQ / mS;
        UQ[k] = Ucm[k] - rel[k] * mP / mS;
    }
}

// The energy-exchange part of LarsenBorgnakke*::collide for one partner with every internal mode (serial application: electronic,
// vibrational mode by mode, rotational; …/LarsenBorgnakkeVariableHardSphere.C:232-323); Etr is the relative translational energy
// carried from exchange to exchange.
__device__ inline void lb_exchange_partner(const DevParams& prm, Stream& r, const DevSpecies& a, const DevSpeciesInt& A, double omegaPQ, double ChiB,
                                           double& Etr, double& erot, unsigned long long& vib, int& elev) {
    const double preERot = erot;
    double preEVib[UGF_MAX_VIB_MODES];
    for (int m = 0; m < a.vibDoF; ++m) preEVib[m] = vib_level(vib, m) * A.thetaV[m] * kB;
    const double preEEle = A.elecE[elev];
    if (prm.invZel > r.u01()) {
        const double Ec = Etr + preEEle;
        elev = post_collision_elec_level(r, Ec, a.nElec, omegaPQ, A);
        Etr = Ec - A.elecE[elev];
    }
    for (int m = 0; m < a.vibDoF; ++m) {
        const double Ec = Etr + preEVib[m];
        const int iMax = (int)(Ec / (kB * A.thetaV[m]));
        if (iMax > 0) {
            const int lv = post_collision_vib_level(r, vib_level(vib, m), iMax, A.thetaV[m], A.thetaD[m], A.TrefZv[m], omegaPQ, A.Zref[m], Ec);
            vib = vib_set(vib, m, lv);
            Etr = Ec - lv * A.thetaV[m] * kB;
        }
    }
    if (a.rotDoF > 0) {
        if (prm.invZrot > r.u01()) {
            const double Ec = Etr + preERot;
            const double ratio = post_collision_rotational_energy(r, a.rotDoF, ChiB);
            erot = ratio * Ec;
            Etr = Ec - erot;
        }
    }
}

// Larsen-Borgnakke collision of a pair that carries vibrational / electronic levels (DevParams::spi set): out of line, the gases of
// the BASELINE configurations never come here.
__device__ __noinline__ void collide_pair_internal(const DevParams& prm, Stream& r, int tP, int tQ, double UP[3], double UQ[3], double& erotP,
                                                   double& erotQ, unsigned long long& vibP, unsigned long long& vibQ, int& elevP, int& elevQ) {
    const DevSpecies& a = prm.sp[tP];
    const DevSpecies& b = prm.sp[tQ];
    const double mP = a.mass, mQ = b.mass, mS = mP + mQ;
    double Ucm[3], cRc[3];
    for (int k = 0; k < 3; ++k) {
        Ucm[k] = (mP * UP[k] + mQ * UQ[k]) / mS;
        cRc[k] = UP[k] - UQ[k];
    }
    const double cRsqr = cRc[0] * cRc[0] + cRc[1] * cRc[1] + cRc[2] * cRc[2];
    const double omegaPQ = 0.5 * (a.omega + b.omega);
    const double mR = mP * mQ / mS;
    double Etr = 0.5 * mR * cRsqr;
    const double ChiB = 2.5 - omegaPQ;
    lb_exchange_partner(prm, r, a, prm.spi[tP], omegaPQ, ChiB, Etr, erotP, vibP, elevP);
    lb_exchange_partner(prm, r, b, prm.spi[tQ], omegaPQ, ChiB, Etr, erotQ, vibQ, elevQ);
    const double cRnew = sqrt((2.0 * Etr) / mR);
    double rel[3];
    if (prm.binaryModel == UGF_BINARY_LB_VHS) scatter_vhs(r, cRnew, rel);
    else scatter_vss(r, cRc, 0.5 * (a.alpha + b.alpha), cRnew / sqrt(cRsqr), rel);
    for (int k = 0; k < 3; ++k) {
        UP[k] = Ucm[k] + rel[k] * mQ / mS;
        UQ[k] = Ucm[k] - rel[k] * mP / mS;
    }
}

// axisymmetricSimulation: mean of RWF(position) over the cell's parcels in occupancy order (noTimeCounter.C:170-182); out of
// line, one lane per cell
__device__ __noinline__ double ntc_mean_rwf(const DevParams& prm, const ParcelBuf& P, int cell, int beg, int n) {
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += axi_rwf(prm, P.y[beg + i], P.z[beg + i]);
    return sum / (double)n;
}

// NTC candidates for one cell: n_sel = 1/2 N (N-1) F_N (sigma_T c_r)max dt / V with stochastic rounding
// (noTimeCounter.C:184-191); the rounding draw comes from the cell's own stream.
__device__ __forceinline__ int ntc_candidates(const DevParams& prm, uint32_t step, uint32_t sub, double dtSub, int cell, int n, double sMaxOld,
                                              double vol, double rwfMean = 1.0) {
    const double selectedPairs = 0.5 * n * (n - 1) * (cell_fn(prm, cell) * rwfMean) * sMaxOld * dtSub / vol;  // noTimeCounter.C:168-184
    int nCand = (int)selectedPairs;
    Stream rc(prm.seed, KIND_NTC, sub, step, (uint32_t)cell, 0xFFFFFFFFu);
    if (rc.u01() < (selectedPairs - nCand)) nCand++;
    return nCand;
}

// sigma_T c_r of a pair (variableHardSphere.C:72-115; the VSS and LB models code the same expression)
__device__ __forceinline__ double sigma_tcr(const DevParams& prm, const DevSpecies& A, const DevSpecies& B, int tP, int tQ, double cR2) {
    if (cR2 < VSMALL) return 0.0;
    const double dPQ = 0.5 * (A.d + B.d);
    const double omegaPQ = 0.5 * (A.omega + B.omega);
    const double mR = A.mass * B.mass / (A.mass + B.mass);
    const double sigmaTPQ = PI * dPQ * dPQ * pow(2.0 * kB * prm.Tref / (mR * cR2), omegaPQ - 0.5) * prm.pairInvGamma[tP * UGF_MAX_SPECIES + tQ];
    return sigmaTPQ * sqrt(cR2);
}

constexpr int CELL_CHUNK = 8;    // most cells staged together as one run (moment phase: 32 / CELL_CHUNK lanes per cell)
constexpr int CELL_TASK_MAX = 24; // most consecutive cells a warp claims at a time (their CSR offsets sit in lanes 0..taskCells)
constexpr int CELL_CAP = 128;    // parcels staged per warp; larger cells take the single-cell paths
constexpr int CELL_ITERS = CELL_CAP / 32;
constexpr int CELL_LPC = 32 / CELL_CHUNK;  // lanes cooperating on one cell in the moment phase (4)

__device__ __forceinline__ double select4(int q, double v0, double v1, double v2, double v3) {
    const double lo = (q & 1) ? v1 : v0;
    const double hi = (q & 1) ? v3 : v2;
    return (q & 2) ? hi : lo;
}

// ---- streaming kernel -----------------------------------------------------------------------------------------
// shared memory per warp: three double arrays (+ERot) that hold the run's velocities for the moment phase and
// are then reused for its positions, int cell ids, typeId bytes.  Keeping the carve-out this small matters: the
// cp.async.ca fills in flight need L1 lines, so a large shared-memory carve-out (small L1) caps the memory-level
// parallelism (measured: profiles/README.md).
__host__ __device__ constexpr int cell_doubles_per_parcel(bool hasRot) { return hasRot ? 4 : 3; }
__host__ __device__ constexpr size_t cell_smem_bytes(bool hasRot) {
    return (size_t)CELL_WARPS * CELL_CAP * (cell_doubles_per_parcel(hasRot) * sizeof(double) + sizeof(int) + 1);
}

// Contribution of one species' cell sums to accumulator slot k of uniGasVolFields::calculateField
// (uniGasVolFields.C:761-797; slot list in DESIGN.md section fields).  FN = nParticle * CWF of the cell (cell_fn), RWF = 1.
__device__ __forceinline__ double acc_term(const DevParams& prm, double FN, int s, int k, double cnt, double su, double sv, double sw, double scc, double se) {
    const DevSpecies& S = prm.sp[s];
    const double m = S.mass;
    switch (k) {
        case 0: return cnt;
        case 1: return m * cnt;
        case 2: return m * scc;
        case 3: return m * su;
        case 4: return m * sv;
        case 5: return m * sw;
        case 6: return se;
        case 7: return S.rotDoF * cnt;
        // the XnParticle sums: with axisymmetricSimulation every parcel counts with its own RWF - axi_moments_kernel adds them
        case 8: return prm.axi ? 0.0 : cnt * FN;
        case 9: return prm.axi ? 0.0 : m * cnt * FN;
        case 10: return prm.axi ? 0.0 : m * su * FN;
        case 11: return prm.axi ? 0.0 : m * sv * FN;
        case 12: return prm.axi ? 0.0 : m * sw * FN;
        case 13: return prm.axi ? 0.0 : m * scc * FN;
        case 14: return S.rotDoF > 0 ? cnt : 0.0;
        default: return (5.0 + S.rotDoF) * cnt;
    }
}

// Moments of one cell for every species from a randomly accessible view (pu0.. indexed 0..n-1), written as
// one coalesced 256-byte block per (cell, species): lane k stores slot k (DESIGN.md section moments).  Used for
// cells that do not fit a staged run; `accumulate` adds to the block instead of overwriting it.
template <bool HAS_ROT, bool MULTI>
__device__ __forceinline__ void cell_moments(const DevParams& prm, double* __restrict__ mom, int cell, int n, const double* pu0,
                                             const double* pu1, const double* pu2, const double* pe, const uint8_t* pt, int lane, bool accumulate) {
    const int nS = prm.nSpecies;
    for (int s = 0; s < nS; ++s) {
        // value index: 0-2 U, 3-8 uu uv uw vv vw ww, 9 cc, 10-12 cc*U, 13 count, 14-17 ERot, ERot*U
        constexpr int LOG = HAS_ROT ? 5 : 4;
        double acc[1 << LOG];
#pragma unroll
        for (int k = 0; k < (1 << LOG); ++k) acc[k] = 0.0;
        for (int j = lane; j < n; j += 32) {
            if (MULTI && pt[j] != s) continue;
            const double u = pu0[j], v = pu1[j], w = pu2[j];
            const double cc = u * u + v * v + w * w;
            acc[0] += u; acc[1] += v; acc[2] += w;
            acc[3] += u * u; acc[4] += u * v; acc[5] += u * w; acc[6] += v * v; acc[7] += v * w; acc[8] += w * w;
            acc[9] += cc;
            acc[10] += cc * u; acc[11] += cc * v; acc[12] += cc * w;
            acc[13] += 1.0;
            if (HAS_ROT) { const double e = pe[j]; acc[14] += e; acc[15] += e * u; acc[16] += e * v; acc[17] += e * w; }
        }
        const double tot = warp_reduce_transpose<LOG>(acc, lane);  // lane l: total of value l
        int src = -1;
        if (lane < 2) src = 13;
        else if (lane < 5) src = lane - 2;
        else if (lane < 8) src = lane - 5;
        else if (lane < 18) src = lane - 5;
        else if (lane < 22) src = HAS_ROT ? lane - 4 : -1;
        else if (lane == 26) src = 13;
        double o = __shfl_sync(0xffffffffu, tot, src & 31);
        if (src < 0) o = 0.0;
        if (lane == 26) o = o * prm.sp[s].E0;
        double* at = mom + ((size_t)cell * nS + s) * UGF_NMOM + lane;
        *at = accumulate ? *at + o : o;
    }
}

// A cell larger than the staging buffer: slices of CELL_CAP parcels go through shared memory one after another.
template <bool HAS_ROT, bool MULTI>
__device__ __noinline__ void stream_giant_cell(const DevParams& prm, const CellArgs& a, int cell, int beg, int n, double* sU0, double* sU1,
                                               double* sU2, double* sE, uint8_t* sT, int lane) {
    for (int b = 0; b < n; b += CELL_CAP) {
        const int m = min(CELL_CAP, n - b);
        __syncwarp();
        for (int j = lane; j < m; j += 32) {
            const int dst = beg + b + j;
            const int src = a.perm ? (a.perm[dst] & CLONE_MASK) : dst;
            const double ux = a.in.ux[src], uy = a.in.uy[src], uz = a.in.uz[src];
            sU0[j] = ux; sU1[j] = uy; sU2[j] = uz;
            if (HAS_ROT) sE[j] = a.in.erot[src];
            if (MULTI) sT[j] = a.in.type[src];
            if (a.gather) {
                a.out.x[dst] = a.in.x[src]; a.out.y[dst] = a.in.y[src]; a.out.z[dst] = a.in.z[src];
                a.out.ux[dst] = ux; a.out.uy[dst] = uy; a.out.uz[dst] = uz;
                a.out.cell[dst] = cell;
                if (HAS_ROT) a.out.erot[dst] = sE[j];
                if (MULTI) a.out.type[dst] = sT[j];
                if (a.in.vib) a.out.vib[dst] = a.in.vib[src];
                if (a.in.elev) a.out.elev[dst] = a.in.elev[src];
            }
        }
        __syncwarp();
        if (a.doSample) cell_moments<HAS_ROT, MULTI>(prm, a.mom, cell, m, sU0, sU1, sU2, sE, sT, lane, b > 0);
    }
    __syncwarp();
    if (a.doSample && a.accDt != 0.0 && lane < NACC) {  // uniGasVolFields accumulation from the finished block
        double add = 0.0;
        const double FN = cell_fn(prm, cell);
        for (int s = 0; s < prm.nSpecies; ++s) {
            const double* mm = a.mom + ((size_t)cell * prm.nSpecies + s) * UGF_NMOM;
            add += acc_term(prm, FN, s, lane, mm[0], mm[2], mm[3], mm[4], mm[14], mm[18]);
            if (MULTI && a.accS && lane == 0 && !prm.axi) a.accS[(size_t)cell * prm.nSpecies + s] += a.accDt * (mm[1] * FN);
        }
        a.acc[(size_t)cell * NACC + lane] += a.accDt * add;
    }
    __syncwarp();
}

template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(CELL_THREADS, 4) cell_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ CellArgs a) {
    extern __shared__ double smemD[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    constexpr int cap = CELL_CAP;
    constexpr int perWarp = cap * cell_doubles_per_parcel(HAS_ROT);
    double* const sU0 = smemD + (size_t)wib * perWarp;
    double* const sU1 = sU0 + cap;
    double* const sU2 = sU1 + cap;
    double* const sE = sU2 + cap;  // valid only if HAS_ROT
    int* const sC = reinterpret_cast<int*>(smemD + (size_t)CELL_WARPS * perWarp) + (size_t)wib * cap;
    uint8_t* const sT = reinterpret_cast<uint8_t*>(reinterpret_cast<int*>(smemD + (size_t)CELL_WARPS * perWarp) + (size_t)CELL_WARPS * cap) + (size_t)wib * cap;
    const int CELL_TASK = a.taskCells;
    const int nTasks = (a.nCells + CELL_TASK - 1) / CELL_TASK;
    const int nS = prm.nSpecies;

    // Tasks of CELL_TASK consecutive cells are claimed from a global counter (zeroed by the host before the launch):
    // warps that draw sparse cells simply claim more tasks.  The id of the following task and its CSR offsets are
    // requested while the current one is processed, so neither round trip is exposed.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *a.taskReset = 0;
        if (a.dN) *a.dN = *a.total;
    }
    int task = 0, offNext = 0x7fffffff;
    if (lane == 0) task = atomicAdd(a.taskCounter, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task < nTasks && lane <= min(CELL_TASK, a.nCells - task * CELL_TASK)) offNext = a.off[task * CELL_TASK + lane];
    while (task < nTasks) {
        const int c0 = task * CELL_TASK;
        const int nc = min(CELL_TASK, a.nCells - c0);
        const int offv = offNext;  // lanes 0..nc hold the task's CSR offsets
        int taskNext = 0;
        if (lane == 0) taskNext = atomicAdd(a.taskCounter, 1);
        bool nextRequested = false;
        int done = 0;
        while (done < nc) {
            // largest run of cells [done, done+k), k <= CELL_CHUNK, whose parcels fit the staging buffer together
            const int b0 = __shfl_sync(0xffffffffu, offv, done);
            const unsigned fit = __ballot_sync(0xffffffffu, lane > done && lane <= nc && lane <= done + CELL_CHUNK && (offv - b0) <= cap);
            const int k = __popc(fit);
            if (k == 0) {  // a single cell larger than the buffer
                const int e0 = __shfl_sync(0xffffffffu, offv, done + 1);
                stream_giant_cell<HAS_ROT, MULTI>(prm, a, c0 + done, b0, e0 - b0, sU0, sU1, sU2, sE, sT, lane);
                done += 1;
                continue;
            }
            const int ntot = __shfl_sync(0xffffffffu, offv, done + k) - b0;
            // ---- phase 1: velocities (+ERot, cell id) of the run into shared memory, all copies in flight at once ----
            int srcs[CELL_ITERS];
#pragma unroll
            for (int it = 0; it < CELL_ITERS; ++it) {
                const int j = it * 32 + lane;
                srcs[it] = (j < ntot) ? (a.perm ? (a.perm[b0 + j] & CLONE_MASK) : b0 + j) : -1;
            }
#pragma unroll
            for (int it = 0; it < CELL_ITERS; ++it) {
                const int j = it * 32 + lane;
                const int src = srcs[it];
                if (src >= 0 && !(a.flags & 2)) {
                    cp_async8(&sU0[j], &a.in.ux[src]);
                    cp_async8(&sU1[j], &a.in.uy[src]);
                    cp_async8(&sU2[j], &a.in.uz[src]);
                    if (a.gather) cp_async4(&sC[j], &a.in.cell[src]);
                    if (HAS_ROT) cp_async8(&sE[j], &a.in.erot[src]);
                    if (MULTI) sT[j] = a.in.type[src];
                }
            }
            if (a.flags & 2) {
                double u[CELL_ITERS], v[CELL_ITERS], w[CELL_ITERS];
#pragma unroll
                for (int it = 0; it < CELL_ITERS; ++it) {
                    const int src = srcs[it];
                    if (src >= 0) { u[it] = __ldcg(&a.in.ux[src]); v[it] = __ldcg(&a.in.uy[src]); w[it] = __ldcg(&a.in.uz[src]); }
                }
#pragma unroll
                for (int it = 0; it < CELL_ITERS; ++it) {
                    const int j = it * 32 + lane;
                    const int src = srcs[it];
                    if (src >= 0) {
                        sU0[j] = u[it]; sU1[j] = v[it]; sU2[j] = w[it];
                        if (a.gather) sC[j] = __ldcg(&a.in.cell[src]);
                        if (HAS_ROT) sE[j] = __ldcg(&a.in.erot[src]);
                        if (MULTI) sT[j] = a.in.type[src];
                    }
                }
            }
            if (!nextRequested) {  // the first copies of this task are in flight: fetch the next task's offsets behind them
                nextRequested = true;
                taskNext = __shfl_sync(0xffffffffu, taskNext, 0);
                offNext = 0x7fffffff;
                if (taskNext < nTasks && lane <= min(CELL_TASK, a.nCells - taskNext * CELL_TASK)) offNext = a.off[taskNext * CELL_TASK + lane];
            }
            cp_async_wait_all();
            __syncwarp();
            // ---- moments: CELL_LPC lanes per cell, all cells of the run at once -------------------------------------
            if (a.doSample) {
                const int ci = done + (lane / CELL_LPC);  // lane holding this cell's offset
                const int q = lane % CELL_LPC;
                const bool cellValid = (lane / CELL_LPC) < k;
                const int cb = __shfl_sync(0xffffffffu, offv, ci & 31);
                const int ce = __shfl_sync(0xffffffffu, offv, (ci + 1) & 31);
                const int s0 = cb - b0;
                const int n = cellValid ? ce - cb : 0;
                const bool doAcc = a.accDt != 0.0;
                double* arow = a.acc + (size_t)(c0 + ci) * NACC + q;  // this lane's 4 accumulator slots: q, q+4, q+8, q+12
                double ac0 = 0, ac1 = 0, ac2 = 0, ac3 = 0;
                if (doAcc && cellValid) { ac0 = arow[0]; ac1 = arow[4]; ac2 = arow[8]; ac3 = arow[12]; }
                for (int s = 0; s < nS; ++s) {
                    double su = 0, sv = 0, sw = 0, suu = 0, suv = 0, suw = 0, svv = 0, svw = 0, sww = 0, scc = 0, scu = 0, scv = 0, scw = 0, cnt = 0;
                    double se = 0, seu = 0, sev = 0, sew = 0;
                    for (int i = q; i < n; i += CELL_LPC) {
                        if (MULTI && sT[s0 + i] != s) continue;
                        const double u = sU0[s0 + i], v = sU1[s0 + i], w = sU2[s0 + i];
                        const double cc = u * u + v * v + w * w;
                        su += u; sv += v; sw += w;
                        suu += u * u; suv += u * v; suw += u * w; svv += v * v; svw += v * w; sww += w * w;
                        scc += cc;
                        scu += cc * u; scv += cc * v; scw += cc * w;
                        cnt += 1.0;
                        if (HAS_ROT) { const double e = sE[s0 + i]; se += e; seu += e * u; sev += e * v; sew += e * w; }
                    }
#pragma unroll
                    for (int m = 1; m < CELL_LPC; m <<= 1) {
                        su += __shfl_xor_sync(0xffffffffu, su, m); sv += __shfl_xor_sync(0xffffffffu, sv, m); sw += __shfl_xor_sync(0xffffffffu, sw, m);
                        suu += __shfl_xor_sync(0xffffffffu, suu, m); suv += __shfl_xor_sync(0xffffffffu, suv, m); suw += __shfl_xor_sync(0xffffffffu, suw, m);
                        svv += __shfl_xor_sync(0xffffffffu, svv, m); svw += __shfl_xor_sync(0xffffffffu, svw, m); sww += __shfl_xor_sync(0xffffffffu, sww, m);
                        scc += __shfl_xor_sync(0xffffffffu, scc, m);
                        scu += __shfl_xor_sync(0xffffffffu, scu, m); scv += __shfl_xor_sync(0xffffffffu, scv, m); scw += __shfl_xor_sync(0xffffffffu, scw, m);
                        cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
                        if (HAS_ROT) {
                            se += __shfl_xor_sync(0xffffffffu, se, m); seu += __shfl_xor_sync(0xffffffffu, seu, m);
                            sev += __shfl_xor_sync(0xffffffffu, sev, m); sew += __shfl_xor_sync(0xffffffffu, sew, m);
                        }
                    }
                    if (cellValid && a.writeMom) {
                        // the cell's 4 lanes write its 32 slots, 4 consecutive slots (one 32-byte sector) per store
                        double* mrow = a.mom + ((size_t)(c0 + ci) * nS + s) * UGF_NMOM + q;
                        mrow[0] = select4(q, cnt, cnt, su, sv);
                        mrow[4] = select4(q, sw, su, sv, sw);
                        mrow[8] = select4(q, suu, suv, suw, svv);
                        mrow[12] = select4(q, svw, sww, scc, scu);
                        mrow[16] = select4(q, scv, scw, se, seu);
                        mrow[20] = select4(q, sev, sew, 0.0, 0.0);
                        mrow[24] = select4(q, 0.0, 0.0, cnt * prm.sp[s].E0, 0.0);
                        mrow[28] = 0.0;
                    }
                    if (doAcc) {  // uniGasVolFields accumulation fused in: slot 4t + q
                        const double FN = (prm.cwf && cellValid) ? prm.nParticle * __ldg(&prm.cwf[c0 + ci]) : prm.nParticle;
                        if (MULTI && a.accS && cellValid && q == 0 && !prm.axi) a.accS[(size_t)(c0 + ci) * nS + s] += a.accDt * (cnt * FN);
                        ac0 += a.accDt * acc_term(prm, FN, s, q, cnt, su, sv, sw, scc, se);
                        ac1 += a.accDt * acc_term(prm, FN, s, q + 4, cnt, su, sv, sw, scc, se);
                        ac2 += a.accDt * acc_term(prm, FN, s, q + 8, cnt, su, sv, sw, scc, se);
                        ac3 += a.accDt * acc_term(prm, FN, s, q + 12, cnt, su, sv, sw, scc, se);
                    }
                }
                if (doAcc && cellValid) { arow[0] = ac0; arow[4] = ac1; arow[8] = ac2; arow[12] = ac3; }
            }
            if (a.gather) {
                // ---- write velocities cell-major (coalesced), then reuse the buffers for the positions (phase 2) ----
                __syncwarp();
#pragma unroll
                for (int it = 0; it < CELL_ITERS; ++it) {
                    const int j = it * 32 + lane;
                    if (j < ntot) {
                        const double u = sU0[j], v = sU1[j], w = sU2[j];
                        a.out.ux[b0 + j] = u; a.out.uy[b0 + j] = v; a.out.uz[b0 + j] = w;
                        a.out.cell[b0 + j] = sC[j];
                        if (HAS_ROT) a.out.erot[b0 + j] = sE[j];
                        if (MULTI) a.out.type[b0 + j] = sT[j];
                        if (a.in.vib) a.out.vib[b0 + j] = a.in.vib[srcs[it]];
                        if (a.in.elev) a.out.elev[b0 + j] = a.in.elev[srcs[it]];
                        if (!(a.flags & 1)) {
                            cp_async8(&sU0[j], &a.in.x[srcs[it]]);  // same lane, same slot: no cross-lane hazard
                            cp_async8(&sU1[j], &a.in.y[srcs[it]]);
                            cp_async8(&sU2[j], &a.in.z[srcs[it]]);
                        }
                    }
                }
                if (a.flags & 1) {
                    double px[CELL_ITERS], py[CELL_ITERS], pz[CELL_ITERS];
#pragma unroll
                    for (int it = 0; it < CELL_ITERS; ++it)
                        if (srcs[it] >= 0) { px[it] = __ldcg(&a.in.x[srcs[it]]); py[it] = __ldcg(&a.in.y[srcs[it]]); pz[it] = __ldcg(&a.in.z[srcs[it]]); }
#pragma unroll
                    for (int it = 0; it < CELL_ITERS; ++it) {
                        const int j = it * 32 + lane;
                        if (srcs[it] >= 0) { a.out.x[b0 + j] = px[it]; a.out.y[b0 + j] = py[it]; a.out.z[b0 + j] = pz[it]; }
                    }
                } else {
                    cp_async_wait_all();
#pragma unroll
                    for (int it = 0; it < CELL_ITERS; ++it) {
                        const int j = it * 32 + lane;
                        if (j < ntot) { a.out.x[b0 + j] = sU0[j]; a.out.y[b0 + j] = sU1[j]; a.out.z[b0 + j] = sU2[j]; }
                    }
                }
            }
            __syncwarp();
            done += k;
        }
        if (!nextRequested) {  // a task of giant cells only
            taskNext = __shfl_sync(0xffffffffu, taskNext, 0);
            offNext = 0x7fffffff;
            if (taskNext < nTasks && lane <= min(CELL_TASK, a.nCells - taskNext * CELL_TASK)) offNext = a.off[taskNext * CELL_TASK + lane];
        }
        task = taskNext;
    }
}

// ---- axisymmetricSimulation: the RWF-weighted cell sums -------------------------------------------------------------------
// cellMeasurements.C:463-467 weights the XnParticle sums of every parcel with its own RWF.  One warp per cell over the
// cell-major array, after cell_kernel: slots 1 (sum RWF), 5-7 (sum RWF U) and 31 (sum RWF |U|^2) of the cell's moment block
// and the accumulators that take them (uniGasVolFields.C:789-797: slots 8-13, per-species nParcelsXnParticle).
struct AxiMomArgs {
    int nCells;
    const int* off;
    ParcelBuf P;  // cell-major, or reached through perm
    const int* perm;
    double* mom;  // or null
    double* acc;
    double* accS;
    double accDt;
};

template <bool MULTI>
__global__ void __launch_bounds__(256) axi_moments_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ AxiMomArgs a) {
    const int lane = threadIdx.x & 31;
    const int cell = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (cell >= a.nCells) return;
    const int b = a.off[cell], e = a.off[cell + 1];
    const double FN = cell_fn(prm, cell);
    double add[6] = {0, 0, 0, 0, 0, 0};
    for (int s = 0; s < prm.nSpecies; ++s) {
        double sr = 0, su = 0, sv = 0, sw = 0, scc = 0;
        for (int jj = b + lane; jj < e; jj += 32) {
            const int j = a.perm ? (a.perm[jj] & CLONE_MASK) : jj;
            if (MULTI && a.P.type[j] != s) continue;
            const double rw = parcel_rwf(prm, cell, a.P.y[j], a.P.z[j]);
            const double u = a.P.ux[j], v = a.P.uy[j], w = a.P.uz[j];
            sr += rw; su += rw * u; sv += rw * v; sw += rw * w; scc += rw * (u * u + v * v + w * w);
        }
        sr = warp_sum(sr); su = warp_sum(su); sv = warp_sum(sv); sw = warp_sum(sw); scc = warp_sum(scc);
        if (lane == 0) {
            if (a.mom) {
                double* m = a.mom + ((size_t)cell * prm.nSpecies + s) * UGF_NMOM;
                m[1] = sr; m[5] = su; m[6] = sv; m[7] = sw; m[31] = scc;
            }
            if (a.accDt != 0.0) {
                const double ms = prm.sp[s].mass;
                add[0] += sr * FN; add[1] += ms * sr * FN; add[2] += ms * su * FN; add[3] += ms * sv * FN; add[4] += ms * sw * FN;
                add[5] += ms * scc * FN;
                if (a.accS) a.accS[(size_t)cell * prm.nSpecies + s] += a.accDt * (sr * FN);
            }
        }
    }
    if (lane == 0 && a.accDt != 0.0)
        for (int k = 0; k < 6; ++k) a.acc[(size_t)cell * NACC + 8 + k] += a.accDt * add[k];
}

// ---- NTC kernel -----------------------------------------------------------------------------------------------------
// In place on the cell-major buffer.  One warp takes 32 consecutive cells: lane l computes cell l's candidate
// count (noTimeCounter.C:184-191), then the candidates of all 32 cells are laid out over the lanes (candidate g of
// the warp = candidate kk of cell slot cs), so sparse per-cell counts still fill the warp.  Each candidate draws
// from its own Philox stream (step, cell, kk) and touches only its two parcels, directly in global memory - no
// staging, no limit on the cell size.  A candidate commits when no earlier uncommitted candidate shares a parcel
// with it (owner marks = lowest lane touching a parcel), which reproduces the reference's sequential per-cell
// candidate loop (noTimeCounter.C:195-312) exactly.
struct NtcArgs {
    int nCells;
    const int* off;
    ParcelBuf P;  // cell-major, collided in place
    const double* vol;
    double* sigmaTcRMax;
    const int* collModelId;
    int* owner;   // [capacity] conflict marks, 0x7f7f7f7f when idle
    const unsigned short* sub;  // [capacity] virtual sub-cell of every parcel (SUBCELLS only)
    uint32_t step;
    uint32_t sub_cycle;  // noTimeCounterSubCycled: pass index (stream aux), 0 for noTimeCounter
    double dtSub;        // deltaT / nSubCycles (noTimeCounterSubCycled.C:190), deltaT for noTimeCounter
    DevCounters* cnt;
};

// Virtual Cartesian sub-cell of every parcel (noTimeCounter.C:96-161): index = sum over solved directions of
// int(L_d (x_d - min_d) / (max_d - min_d)) * weight_d, with the cell's point bounding box; the index is clamped to
// L_d - 1 (the reference overflows its list for a parcel exactly on the max face).
struct SubcellArgs {
    ParcelBuf P;
    const long long* dN;
    const double* bbMin;
    const double* bbMax;
    const int* levels;  // [nCells*3]
    unsigned short* sub;
};

__global__ void __launch_bounds__(256) subcell_index_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ SubcellArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *a.dN) return;
    const int c = a.P.cell[i];
    if (c < 0) return;
    const double x[3] = {a.P.x[i], a.P.y[i], a.P.z[i]};
    int sc = 0, w = 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (!prm.solD[d]) continue;
        const int L = __ldg(&a.levels[3 * (size_t)c + d]);
        const double mn = __ldg(&a.bbMin[3 * (size_t)c + d]), mx = __ldg(&a.bbMax[3 * (size_t)c + d]);
        int k = (int)(L * (x[d] - mn) / (mx - mn));
        k = k < 0 ? 0 : (k > L - 1 ? L - 1 : k);
        sc += k * w;
        w *= L;
    }
    a.sub[i] = (unsigned short)sc;
}

constexpr int NTC_THREADS = 256;
constexpr int NTC_WARPS = NTC_THREADS / 32;
constexpr int NTC_IDLE = 0x7f7f7f7f;
constexpr int NTC_MARKS = 1024;  // parcels per 32-cell chunk whose conflict marks live in shared memory (global marks beyond)

// INTERNAL: some species carries vibrational modes / several electronic levels (its own instantiation: the Larsen-Borgnakke exchange
// with every mode is a large out-of-line function whose call site alone costs the common kernels registers)
template <bool HAS_ROT, bool MULTI, bool SUBCELLS, bool INTERNAL = false>
__global__ void __launch_bounds__(NTC_THREADS, 2) ntc_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ NtcArgs a) {
    __shared__ double sMaxW[NTC_WARPS][32];
    __shared__ int sMarkW[NTC_WARPS][NTC_MARKS];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    double* const sM = sMaxW[wib];
    int* const sMark = sMarkW[wib];  // conflict marks of the chunk's parcels (slot = parcel - first parcel of the chunk)
    for (int j = lane; j < NTC_MARKS; j += 32) sMark[j] = NTC_IDLE;
    __syncwarp();
    const int warpsTotal = gridDim.x * NTC_WARPS;
    const int nChunks = (a.nCells + 31) / 32;
    unsigned long long wCand = 0;
    int myColl = 0;

    for (int chunk = blockIdx.x * NTC_WARPS + wib; chunk < nChunks; chunk += warpsTotal) {
        const int c0 = chunk * 32;
        const int myCell = c0 + lane;
        const bool valid = myCell < a.nCells;
        // all per-cell inputs are requested together: one memory round trip before the candidate counts
        const int myBeg = valid ? a.off[myCell] : 0;
        const int myEnd = valid ? a.off[myCell + 1] : 0;
        const int myModel = valid ? a.collModelId[myCell] : 0;
        const double sMaxIn = valid ? a.sigmaTcRMax[myCell] : 0.0;
        const double myVol = valid ? a.vol[myCell] : 1.0;
        const int myN = myEnd - myBeg;
        int myCand = 0;
        double mySMax = 0.0;
        if (myN > 1 && myModel == 1) {
            mySMax = sMaxIn;
            double rwfMean = 1.0;
            if (prm.axi) rwfMean = ntc_mean_rwf(prm, a.P, myCell, myBeg, myN);
            myCand = ntc_candidates(prm, a.step, a.sub_cycle, a.dtSub, myCell, myN, mySMax, myVol, rwfMean);
        }
        if (!__any_sync(0xffffffffu, myCand > 0)) continue;
        int incl = myCand;  // inclusive prefix of the candidate counts over the 32 cells
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - myCand;
        const int T = __shfl_sync(0xffffffffu, incl, 31);
        const int chunkBeg = __shfl_sync(0xffffffffu, myBeg, 0);
        const int chunkEnd = __shfl_sync(0xffffffffu, myEnd, min(31, a.nCells - 1 - c0));
        const bool marksInSmem = (chunkEnd - chunkBeg) <= NTC_MARKS;
        if (lane == 0) wCand += (unsigned long long)T;
        sM[lane] = mySMax;
        __syncwarp();
        for (int base = 0; base < T; base += 32) {
            const int g = base + lane;
            bool act = g < T;
            // cell slot of candidate g: number of cells whose inclusive prefix is <= g
            int cs = 0;
#pragma unroll
            for (int stp = 16; stp > 0; stp >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, cs + stp - 1);
                if (v <= g) cs += stp;
            }
            cs &= 31;
            const int kk = g - __shfl_sync(0xffffffffu, excl, cs);
            const int beg = __shfl_sync(0xffffffffu, myBeg, cs);
            const int n = __shfl_sync(0xffffffffu, myN, cs);
            const double sMaxOld = __shfl_sync(0xffffffffu, mySMax, cs);
            Stream r(prm.seed, KIND_NTC, a.sub_cycle, a.step, (uint32_t)(c0 + cs), (uint32_t)kk);
            int gP = 0, gQ = 0;
            int tP = 0, tQ = 0;
            if (act) {
                const int cP = r.position(n);
                int cQ;
                int nSC = 0;
                unsigned short scP = 0;
                if (SUBCELLS) {  // partner from P's sub-cell if it holds another parcel (noTimeCounter.C:205-235)
                    scP = a.sub[beg + cP];
                    for (int i = 0; i < n; ++i) nSC += (a.sub[beg + i] == scP);
                }
                if (SUBCELLS && nSC > 1) {
                    do {
                        int t = r.position(nSC);
                        cQ = 0;
                        for (int i = 0; i < n; ++i) {
                            if (a.sub[beg + i] == scP) {
                                if (t == 0) { cQ = i; break; }
                                --t;
                            }
                        }
                    } while (cP == cQ);
                } else {
                    do { cQ = r.position(n); } while (cP == cQ);
                }
                gP = beg + cP; gQ = beg + cQ;
                if (MULTI) { tP = a.P.type[gP]; tQ = a.P.type[gQ]; }
                if (prm.sp[tP].charge == -1 && prm.sp[tQ].charge == -1) act = false;  // noTimeCounter.C:245-247
            }
            double lmax = 0.0;
            unsigned pending = __ballot_sync(0xffffffffu, act);
            while (pending) {
                const bool mine = (pending >> lane) & 1u;
                bool ready;
                if (marksInSmem) {  // the usual case: the whole chunk fits, no global round trips for the marks
                    volatile int* mk = sMark;
                    if (mine) { atomicMin(&sMark[gP - chunkBeg], lane); atomicMin(&sMark[gQ - chunkBeg], lane); }
                    __syncwarp();
                    ready = mine && mk[gP - chunkBeg] == lane && mk[gQ - chunkBeg] == lane;
                    __syncwarp();
                    if (mine) { mk[gP - chunkBeg] = NTC_IDLE; mk[gQ - chunkBeg] = NTC_IDLE; }
                } else {
                    if (mine) { atomicMin(&a.owner[gP], lane); atomicMin(&a.owner[gQ], lane); }
                    __syncwarp();
                    ready = mine && __ldcg(&a.owner[gP]) == lane && __ldcg(&a.owner[gQ]) == lane;
                    __syncwarp();
                    if (mine) { __stcg(&a.owner[gP], NTC_IDLE); __stcg(&a.owner[gQ], NTC_IDLE); }
                }
                if (ready) {
                    const DevSpecies& A = prm.sp[tP];
                    const DevSpecies& B = prm.sp[tQ];
                    double UP[3] = {__ldcg(&a.P.ux[gP]), __ldcg(&a.P.uy[gP]), __ldcg(&a.P.uz[gP])};
                    double UQ[3] = {__ldcg(&a.P.ux[gQ]), __ldcg(&a.P.uy[gQ]), __ldcg(&a.P.uz[gQ])};
                    const double d0 = UP[0] - UQ[0], d1 = UP[1] - UQ[1], d2 = UP[2] - UQ[2];
                    const double sig = sigma_tcr(prm, A, B, tP, tQ, d0 * d0 + d1 * d1 + d2 * d2);
                    if (sig > lmax) lmax = sig;
                    if ((sig / sMaxOld) > r.u01()) {
                        double eP = 0.0, eQ = 0.0;
                        if (HAS_ROT) { eP = __ldcg(&a.P.erot[gP]); eQ = __ldcg(&a.P.erot[gQ]); }
                        if (INTERNAL && prm.binaryModel >= UGF_BINARY_LB_VHS) {  // species with vibrational modes / several electronic levels
                            unsigned long long vP = a.P.vib ? __ldcg(&a.P.vib[gP]) : 0ull, vQ = a.P.vib ? __ldcg(&a.P.vib[gQ]) : 0ull;
                            int lP = a.P.elev ? (int)__ldcg(&a.P.elev[gP]) : 0, lQ = a.P.elev ? (int)__ldcg(&a.P.elev[gQ]) : 0;
                            collide_pair_internal(prm, r, tP, tQ, UP, UQ, eP, eQ, vP, vQ, lP, lQ);
                            if (a.P.vib) { __stcg(&a.P.vib[gP], vP); __stcg(&a.P.vib[gQ], vQ); }
                            if (a.P.elev) { __stcg(&a.P.elev[gP], (uint8_t)lP); __stcg(&a.P.elev[gQ], (uint8_t)lQ); }
                        } else
                        collide_pair(prm, r, A, B, UP, UQ, eP, eQ);
                        __stcg(&a.P.ux[gP], UP[0]); __stcg(&a.P.uy[gP], UP[1]); __stcg(&a.P.uz[gP], UP[2]);
                        __stcg(&a.P.ux[gQ], UQ[0]); __stcg(&a.P.uy[gQ], UQ[1]); __stcg(&a.P.uz[gQ], UQ[2]);
                        if (HAS_ROT) { __stcg(&a.P.erot[gP], eP); __stcg(&a.P.erot[gQ], eQ); }
                        myColl++;
                    }
                }
                __syncwarp();
                pending &= ~__ballot_sync(0xffffffffu, ready);
            }
            // running per-cell maximum of sigma_T c_r (positive doubles order like their bit patterns)
            if (lmax > sMaxOld) atomicMax(reinterpret_cast<unsigned long long*>(&sM[cs]), (unsigned long long)__double_as_longlong(lmax));
        }
        __syncwarp();
        const double m = sM[lane];
        if (m > mySMax) a.sigmaTcRMax[myCell] = m;
        __syncwarp();
    }
    const int wc = warp_sum_int(myColl);
    if (lane == 0) {
        if (wCand) atomicAdd(&a.cnt->cand, wCand);
        if (wc) atomicAdd(&a.cnt->coll, (unsigned long long)wc);
    }
}

}  // namespace ugf
