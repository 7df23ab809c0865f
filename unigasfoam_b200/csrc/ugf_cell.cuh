// ugf_cell.cuh — the per-cell kernel: payload gather through the occupancy permutation (the sort's scatter),
// cell-moment sampling and NTC collisions in one pass over each cell's parcels.
//
// Replaces, fused:  cellMeasurements::calculateFields (U/cellMeasurements/cellMeasurements.C:408-513),
//                   noTimeCounter::collide (U/dsmcCollisionPartner/derived/noTimeCounter/noTimeCounter.C:66-343),
//                   variableHardSphere / variableSoftSphere / LarsenBorgnakke* ::{sigmaTcR, collide}
//                   (U/dsmcCollisions/derived/*), and the kinetic samplers they call
//                   (U/clouds/uniGasCloud.C:1129-1189, 1267-1326).
//
// One warp per cell (grid-stride over cells, persistent CTAs).  The cell's velocities (and ERot, typeId) are
// staged in shared memory, moments are reduced with warp shuffles and written as one coalesced 256-byte block
// per (cell, species), and the collided velocities are written straight into the cell-major output buffer.
// Sampling uses the pre-collision state, as the reference does (uniGasCloud.C:846-850).
//
// NTC candidates of a cell are processed 32 at a time, one per lane, each from its own Philox stream
// (step, cell, candidate#).  A candidate commits as soon as no earlier uncommitted candidate shares a parcel
// with it, which reproduces the reference's sequential candidate loop exactly (same results as the oracle's
// serial loop) while independent pairs collide in parallel.
//
// HBM-bound: per parcel 4 (perm) + 52/60 (gather) read, 52/60 written; per cell 8 (offsets) + 16 (sigmaTcRMax)
// + 256 x species (moments).
#pragma once
#include "ugf_common.cuh"
#include "ugf_rng.cuh"

namespace ugf {

constexpr int CELL_THREADS = 256;
constexpr int CELL_WARPS = CELL_THREADS / 32;

struct CellArgs {
    int nCells;
    const int* off;
    const int* perm;  // null: identity (array already cell-major)
    ParcelBuf in, out;
    int gather;       // 1: write the cell-major copy into `out`; 0: operate in place on `in`
    int doSample, doCollide;
    double* mom;
    const double* vol;
    double* sigmaTcRMax;
    const int* collModelId;
    uint32_t step;
    DevCounters* cnt;
    int cap;  // staging capacity per warp, parcels
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

// NTC candidates for one cell: n_sel = 1/2 N (N-1) F_N (sigma_T c_r)max dt / V with stochastic rounding
// (noTimeCounter.C:184-191); the rounding draw comes from the cell's own stream.
__device__ __forceinline__ int ntc_candidates(const DevParams& prm, uint32_t step, int cell, int n, double sMaxOld, double vol) {
    const double selectedPairs = 0.5 * n * (n - 1) * prm.nParticle * sMaxOld * prm.deltaT / vol;
    int nCand = (int)selectedPairs;
    Stream rc(prm.seed, KIND_NTC, 0, step, (uint32_t)cell, 0xFFFFFFFFu);
    if (rc.u01() < (selectedPairs - nCand)) nCand++;
    return nCand;
}

// NTC candidate loop for one cell (noTimeCounter.C:195-312); the whole warp calls this, out of line so that
// the streaming part of the kernel stays lean.  pu0..pt give random access to the cell's staged velocities /
// ERot / typeId (shared or global memory).
template <bool HAS_ROT, bool MULTI>
__device__ __noinline__ int ntc_collide(const DevParams& prm, const CellArgs& a, int cell, int n, int nCand, double sMaxOld,
                                        double* pu0, double* pu1, double* pu2, double* pe, const uint8_t* pt, int lane) {
    int myColl = 0;
    double localMax = sMaxOld;
    for (int base = 0; base < nCand; base += 32) {
        const int k = base + lane;
        bool act = k < nCand;
        Stream r(prm.seed, KIND_NTC, 0, a.step, (uint32_t)cell, (uint32_t)k);
        int cP = -1, cQ = -2;
        int tP = 0, tQ = 0;
        if (act) {
            cP = r.position(n);
            do { cQ = r.position(n); } while (cP == cQ);
            if (MULTI) { tP = pt[cP]; tQ = pt[cQ]; }
            // electron-electron pairs are skipped (noTimeCounter.C:245-247)
            if (prm.sp[tP].charge == -1 && prm.sp[tQ].charge == -1) act = false;
        }
        unsigned pending = __ballot_sync(0xffffffffu, act);
        while (pending) {
            const bool mine = (pending >> lane) & 1u;
            bool blocked = false;
            for (unsigned mm = pending; mm; mm &= mm - 1) {
                const int j = __ffs(mm) - 1;
                const int pj = __shfl_sync(0xffffffffu, cP, j);
                const int qj = __shfl_sync(0xffffffffu, cQ, j);
                if (j < lane && (pj == cP || pj == cQ || qj == cP || qj == cQ)) blocked = true;
            }
            const bool ready = mine && !blocked;
            if (ready) {
                const DevSpecies& A = prm.sp[tP];
                const DevSpecies& B = prm.sp[tQ];
                double UP[3] = {pu0[cP], pu1[cP], pu2[cP]};
                double UQ[3] = {pu0[cQ], pu1[cQ], pu2[cQ]};
                const double d0 = UP[0] - UQ[0], d1 = UP[1] - UQ[1], d2 = UP[2] - UQ[2];
                const double cR2 = d0 * d0 + d1 * d1 + d2 * d2;
                double sig = 0.0;
                if (!(cR2 < VSMALL)) {  // variableHardSphere.C:72-115
                    const double dPQ = 0.5 * (A.d + B.d);
                    const double omegaPQ = 0.5 * (A.omega + B.omega);
                    const double mR = A.mass * B.mass / (A.mass + B.mass);
                    const double sigmaTPQ = PI * dPQ * dPQ * pow(2.0 * kB * prm.Tref / (mR * cR2), omegaPQ - 0.5)
                                            * prm.pairInvGamma[tP * UGF_MAX_SPECIES + tQ];
                    sig = sigmaTPQ * sqrt(cR2);
                }
                if (sig > localMax) localMax = sig;
                if ((sig / sMaxOld) > r.u01()) {
                    double eP = 0.0, eQ = 0.0;
                    if (HAS_ROT) { eP = pe[cP]; eQ = pe[cQ]; }
                    collide_pair(prm, r, A, B, UP, UQ, eP, eQ);
                    pu0[cP] = UP[0]; pu1[cP] = UP[1]; pu2[cP] = UP[2];
                    pu0[cQ] = UQ[0]; pu1[cQ] = UQ[1]; pu2[cQ] = UQ[2];
                    if (HAS_ROT) { pe[cP] = eP; pe[cQ] = eQ; }
                    myColl++;
                }
            }
            __syncwarp();
            pending &= ~__ballot_sync(0xffffffffu, ready);
        }
    }
    localMax = warp_max(localMax);
    if (lane == 0 && localMax > sMaxOld) a.sigmaTcRMax[cell] = localMax;
    return myColl;
}

// Moments of one cell for every species from a randomly accessible view (pu0.. indexed 0..n-1), written as
// one coalesced 256-byte block per (cell, species): lane k stores slot k (DESIGN.md section moments).
template <bool HAS_ROT, bool MULTI>
__device__ __forceinline__ void cell_moments(const DevParams& prm, double* __restrict__ mom, int cell, int n, const double* pu0,
                                             const double* pu1, const double* pu2, const double* pe, const uint8_t* pt, int lane) {
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
        mom[((size_t)cell * nS + s) * UGF_NMOM + lane] = o;
    }
}

// One cell whose parcels do not fit the staging buffer together with its neighbours (or not at all): the
// general path.  Stages in shared memory when n <= cap, otherwise works on the output arrays in global memory.
template <bool HAS_ROT, bool MULTI>
__device__ __noinline__ void cell_single(const DevParams& prm, const CellArgs& a, int cell, int beg, int n, double* sU0, double* sU1,
                                         double* sU2, double* sE, uint8_t* sT, int lane) {
    const int cap = a.cap;  // == CELL_CAP
    const bool collideHere = a.doCollide && n > 1 && a.collModelId[cell] == 1;
    const bool useSmem = n <= cap;
    const ParcelBuf& fin = a.gather ? a.out : a.in;
    if (a.gather || useSmem) {
        for (int j = lane; j < n; j += 32) {
            const int src = a.perm ? a.perm[beg + j] : beg + j;
            const double ux = a.in.ux[src], uy = a.in.uy[src], uz = a.in.uz[src];
            double e = 0.0;
            if (HAS_ROT) e = a.in.erot[src];
            uint8_t t = 0;
            if (MULTI) t = a.in.type[src];
            if (a.gather) {
                a.out.x[beg + j] = a.in.x[src];
                a.out.y[beg + j] = a.in.y[src];
                a.out.z[beg + j] = a.in.z[src];
                a.out.cell[beg + j] = cell;
                if (MULTI) a.out.type[beg + j] = t;
            }
            if (useSmem) {
                sU0[j] = ux; sU1[j] = uy; sU2[j] = uz;
                if (HAS_ROT) sE[j] = e;
                if (MULTI) sT[j] = t;
            } else {
                a.out.ux[beg + j] = ux; a.out.uy[beg + j] = uy; a.out.uz[beg + j] = uz;
                if (HAS_ROT) a.out.erot[beg + j] = e;
            }
        }
        __syncwarp();
    }
    // randomly accessible view: shared memory, or the final arrays (gathered copy / in place).  A cell larger than
    // the staging buffer that is sampled through a permutation without gathering has no such view: the host never
    // asks for that combination with collisions, and sampling falls back to a strided pass below.
    double *pu0, *pu1, *pu2, *pe;
    uint8_t* pt;
    const bool direct = useSmem || a.gather || a.perm == nullptr;
    if (useSmem) { pu0 = sU0; pu1 = sU1; pu2 = sU2; pe = sE; pt = sT; }
    else { pu0 = fin.ux + beg; pu1 = fin.uy + beg; pu2 = fin.uz + beg; pe = HAS_ROT ? fin.erot + beg : nullptr; pt = MULTI ? fin.type + beg : nullptr; }
    if (a.doSample) {
        if (direct) {
            cell_moments<HAS_ROT, MULTI>(prm, a.mom, cell, n, pu0, pu1, pu2, pe, pt, lane);
        } else {
            // giant cell, sample-only through the permutation: stage slices of `cap` parcels and add up
            const int nS = prm.nSpecies;
            for (int s = 0; s < nS; ++s) a.mom[((size_t)cell * nS + s) * UGF_NMOM + lane] = 0.0;
            for (int b = 0; b < n; b += cap) {
                const int m = min(cap, n - b);
                __syncwarp();
                for (int j = lane; j < m; j += 32) {
                    const int src = a.perm[beg + b + j];
                    sU0[j] = a.in.ux[src]; sU1[j] = a.in.uy[src]; sU2[j] = a.in.uz[src];
                    if (HAS_ROT) sE[j] = a.in.erot[src];
                    if (MULTI) sT[j] = a.in.type[src];
                }
                __syncwarp();
                for (int s = 0; s < nS; ++s) {
                    const size_t at = ((size_t)cell * nS + s) * UGF_NMOM + lane;
                    const double prev = a.mom[at];
                    cell_moments<HAS_ROT, MULTI>(prm, a.mom, cell, m, sU0, sU1, sU2, sE, sT, lane);
                    a.mom[at] += prev;
                }
            }
        }
    }
    if (collideHere) {
        const double sMaxOld = a.sigmaTcRMax[cell];
        const int nCand = ntc_candidates(prm, a.step, cell, n, sMaxOld, a.vol[cell]);
        if (lane == 0 && nCand > 0) atomicAdd(&a.cnt->cand, (unsigned long long)nCand);
        if (nCand > 0) {
            const int nc = warp_sum_int(ntc_collide<HAS_ROT, MULTI>(prm, a, cell, n, nCand, sMaxOld, pu0, pu1, pu2, pe, pt, lane));
            if (lane == 0 && nc > 0) atomicAdd(&a.cnt->coll, (unsigned long long)nc);
        }
    }
    if (useSmem && (a.gather || collideHere)) {
        for (int j = lane; j < n; j += 32) {
            fin.ux[beg + j] = sU0[j]; fin.uy[beg + j] = sU1[j]; fin.uz[beg + j] = sU2[j];
            if (HAS_ROT) fin.erot[beg + j] = sE[j];
        }
    }
    __syncwarp();
}

constexpr int CELL_CHUNK = 8;    // consecutive cells examined by one warp per iteration
constexpr int CELL_CAP = 128;    // parcels staged per warp (fast path); larger cells take cell_single
constexpr int CELL_ITERS = CELL_CAP / 32;
constexpr int CELL_LPC = 32 / CELL_CHUNK;  // lanes cooperating on one cell in the moment phase (4)

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory doubles per staged parcel: U, x (gather only) (+ERot); then int cell ids and typeId bytes
__host__ __device__ constexpr int cell_doubles_per_parcel(bool hasRot) { return hasRot ? 7 : 6; }
__host__ __device__ constexpr size_t cell_smem_bytes(bool hasRot) {
    return (size_t)CELL_WARPS * CELL_CAP * (cell_doubles_per_parcel(hasRot) * sizeof(double) + sizeof(int) + 1);
}

__device__ __forceinline__ double select4(int q, double v0, double v1, double v2, double v3) {
    const double lo = (q & 1) ? v1 : v0;
    const double hi = (q & 1) ? v3 : v2;
    return (q & 2) ? hi : lo;
}

template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(CELL_THREADS, 3) cell_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ CellArgs a) {
    extern __shared__ double smemD[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    constexpr int cap = CELL_CAP;
    constexpr int perWarp = cap * cell_doubles_per_parcel(HAS_ROT);
    double* sU0 = smemD + (size_t)wib * perWarp;
    double* sU1 = sU0 + cap;
    double* sU2 = sU1 + cap;
    double* sX0 = sU2 + cap;
    double* sX1 = sX0 + cap;
    double* sX2 = sX1 + cap;
    double* sE = sX2 + cap;  // valid only if HAS_ROT
    int* sC = reinterpret_cast<int*>(smemD + (size_t)CELL_WARPS * perWarp) + (size_t)wib * cap;
    uint8_t* sT = reinterpret_cast<uint8_t*>(reinterpret_cast<int*>(smemD + (size_t)CELL_WARPS * perWarp) + (size_t)CELL_WARPS * cap) + (size_t)wib * cap;
    const int warpsTotal = gridDim.x * CELL_WARPS;
    const int nChunks = (a.nCells + CELL_CHUNK - 1) / CELL_CHUNK;
    const ParcelBuf& fin = a.gather ? a.out : a.in;
    const int nS = prm.nSpecies;
    unsigned long long wCand = 0;
    int myColl = 0;

    for (int chunk = blockIdx.x * CELL_WARPS + wib; chunk < nChunks; chunk += warpsTotal) {
        const int c0 = chunk * CELL_CHUNK;
        const int nc = min(CELL_CHUNK, a.nCells - c0);
        const int offv = (lane <= nc) ? a.off[c0 + lane] : 0x7fffffff;  // lanes 0..nc hold the chunk's CSR offsets
        int done = 0;
        while (done < nc) {
            // largest run of cells [done, done+k) whose parcels fit the staging buffer together
            const int b0 = __shfl_sync(0xffffffffu, offv, done);
            const unsigned fit = __ballot_sync(0xffffffffu, lane > done && lane <= nc && (offv - b0) <= cap);
            const int k = __popc(fit);
            if (k == 0) {  // a single cell larger than the buffer
                const int e0 = __shfl_sync(0xffffffffu, offv, done + 1);
                cell_single<HAS_ROT, MULTI>(prm, a, c0 + done, b0, e0 - b0, sU0, sU1, sU2, sE, sT, lane);
                done += 1;
                continue;
            }
            const int ntot = __shfl_sync(0xffffffffu, offv, done + k) - b0;
            // ---- gather the run into shared memory with cp.async: every load of the run is in flight at once ----
            int srcs[CELL_ITERS];
#pragma unroll
            for (int it = 0; it < CELL_ITERS; ++it) {
                const int j = it * 32 + lane;
                srcs[it] = (j < ntot) ? (a.perm ? a.perm[b0 + j] : b0 + j) : -1;
            }
#pragma unroll
            for (int it = 0; it < CELL_ITERS; ++it) {
                const int j = it * 32 + lane;
                const int src = srcs[it];
                if (src >= 0) {
                    cp_async8(&sU0[j], &a.in.ux[src]);
                    cp_async8(&sU1[j], &a.in.uy[src]);
                    cp_async8(&sU2[j], &a.in.uz[src]);
                    if (a.gather) {
                        cp_async8(&sX0[j], &a.in.x[src]);
                        cp_async8(&sX1[j], &a.in.y[src]);
                        cp_async8(&sX2[j], &a.in.z[src]);
                        cp_async4(&sC[j], &a.in.cell[src]);
                    }
                    if (HAS_ROT) cp_async8(&sE[j], &a.in.erot[src]);
                    if (MULTI) sT[j] = a.in.type[src];
                }
            }
            // ---- per-cell scalars while the copies fly, one cell per lane: size, NTC candidate count ----------------
            const int myEnd = __shfl_down_sync(0xffffffffu, offv, 1);
            const int myCell = c0 + lane;
            const bool mineInRun = lane >= done && lane < done + k;
            const int myN = mineInRun ? myEnd - offv : 0;
            int myCand = 0;
            double mySMax = 0.0;
            if (a.doCollide && myN > 1 && a.collModelId[myCell] == 1) {
                mySMax = a.sigmaTcRMax[myCell];
                myCand = ntc_candidates(prm, a.step, myCell, myN, mySMax, a.vol[myCell]);
            }
            const unsigned candMask = __ballot_sync(0xffffffffu, myCand > 0);
            if (candMask) {
                const int candSum = warp_sum_int(myCand);
                if (lane == 0) wCand += (unsigned long long)candSum;
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
                    if (cellValid) {
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
                }
            }
            // ---- NTC collisions for the cells of the run that drew candidates --------------------------------------
            for (unsigned cm = candMask; cm; cm &= cm - 1) {
                const int c = __ffs(cm) - 1;
                const int s0 = __shfl_sync(0xffffffffu, offv, c) - b0;
                const int n = __shfl_sync(0xffffffffu, offv, c + 1) - b0 - s0;
                const int nCand = __shfl_sync(0xffffffffu, myCand, c);
                const double sMaxOld = __shfl_sync(0xffffffffu, mySMax, c);
                myColl += ntc_collide<HAS_ROT, MULTI>(prm, a, c0 + c, n, nCand, sMaxOld, sU0 + s0, sU1 + s0, sU2 + s0, sE + s0, sT + s0, lane);
            }
            __syncwarp();
            // ---- write the run to its final place, coalesced ----------------------------------------------------------
            if (a.gather || candMask) {
#pragma unroll
                for (int it = 0; it < CELL_ITERS; ++it) {
                    const int j = it * 32 + lane;
                    if (j < ntot) {
                        fin.ux[b0 + j] = sU0[j]; fin.uy[b0 + j] = sU1[j]; fin.uz[b0 + j] = sU2[j];
                        if (a.gather) {
                            fin.x[b0 + j] = sX0[j]; fin.y[b0 + j] = sX1[j]; fin.z[b0 + j] = sX2[j];
                            fin.cell[b0 + j] = sC[j];
                            if (MULTI) fin.type[b0 + j] = sT[j];
                        }
                        if (HAS_ROT) fin.erot[b0 + j] = sE[j];
                    }
                }
            }
            __syncwarp();
            done += k;
        }
    }
    const int wc = warp_sum_int(myColl);
    if (lane == 0) {
        if (wCand) atomicAdd(&a.cnt->cand, wCand);
        if (wc) atomicAdd(&a.cnt->coll, (unsigned long long)wc);
    }
}

}  // namespace ugf
