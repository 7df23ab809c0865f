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

// NTC candidate selection and collisions for one cell (noTimeCounter.C:164-312); the whole warp calls this.
// pu0..pt give random access to the cell's staged velocities / ERot / typeId (shared or global memory).
template <bool HAS_ROT, bool MULTI>
__device__ __noinline__ void ntc_collide(const DevParams& prm, const CellArgs& a, int cell, int n, double* pu0, double* pu1, double* pu2,
                                         double* pe, const uint8_t* pt, int lane, unsigned long long& wCand, int& myColl) {
    const double sMaxOld = a.sigmaTcRMax[cell];
    const double selectedPairs = 0.5 * n * (n - 1) * prm.nParticle * sMaxOld * prm.deltaT / a.vol[cell];
    int nCand = (int)selectedPairs;
    {
        Stream rc(prm.seed, KIND_NTC, 0, a.step, (uint32_t)cell, 0xFFFFFFFFu);
        if (rc.u01() < (selectedPairs - nCand)) nCand++;
    }
    if (nCand == 0) return;
    if (lane == 0) wCand += (unsigned long long)nCand;
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
}

template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(CELL_THREADS, 4) cell_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ CellArgs a) {
    extern __shared__ double smemD[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int cap = a.cap;
    const int perWarp = cap * (HAS_ROT ? 4 : 3);
    double* sU0 = smemD + (size_t)wib * perWarp;
    double* sU1 = sU0 + cap;
    double* sU2 = sU1 + cap;
    double* sE = sU2 + cap;  // valid only if HAS_ROT
    uint8_t* sT = reinterpret_cast<uint8_t*>(smemD + (size_t)CELL_WARPS * perWarp) + (size_t)wib * cap;
    const int warpsTotal = gridDim.x * CELL_WARPS;
    const int nS = prm.nSpecies;
    unsigned long long wCand = 0;
    int myColl = 0;

    for (int cell = blockIdx.x * CELL_WARPS + wib; cell < a.nCells; cell += warpsTotal) {
        const int beg = a.off[cell];
        const int n = a.off[cell + 1] - beg;
        const bool collideHere = a.doCollide && n > 1 && a.collModelId[cell] == 1;
        const bool useSmem = n <= cap;
        const ParcelBuf& fin = a.gather ? a.out : a.in;  // where the cell's final data lives

        // staged (randomly accessible) view of the cell's velocities
        double *pu0 = nullptr, *pu1 = nullptr, *pu2 = nullptr, *pe = nullptr;
        uint8_t* pt = nullptr;
        if (collideHere) {
            if (useSmem) { pu0 = sU0; pu1 = sU1; pu2 = sU2; pe = sE; pt = sT; }
            else { pu0 = fin.ux + beg; pu1 = fin.uy + beg; pu2 = fin.uz + beg; pe = HAS_ROT ? fin.erot + beg : nullptr; pt = MULTI ? fin.type + beg : nullptr; }
        }

        // ---- phase A: gather ---------------------------------------------------------------------
        if (a.gather || (collideHere && useSmem)) {
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
                if (collideHere && useSmem) {
                    sU0[j] = ux; sU1[j] = uy; sU2[j] = uz;
                    if (HAS_ROT) sE[j] = e;
                    if (MULTI) sT[j] = t;
                } else if (a.gather) {
                    a.out.ux[beg + j] = ux; a.out.uy[beg + j] = uy; a.out.uz[beg + j] = uz;
                    if (HAS_ROT) a.out.erot[beg + j] = e;
                }
            }
            __syncwarp();
        }

        // ---- phase B: cell moments (pre-collision state) ----------------------------------------------
        if (a.doSample) {
            for (int s = 0; s < nS; ++s) {
                // value index: 0-2 U, 3-8 uu uv uw vv vw ww, 9 cc, 10-12 cc*U, 13 count, 14-17 ERot, ERot*U
                constexpr int LOG = HAS_ROT ? 5 : 4;
                double acc[1 << LOG];
#pragma unroll
                for (int k = 0; k < (1 << LOG); ++k) acc[k] = 0.0;
                for (int j = lane; j < n; j += 32) {
                    double u, v, w, e = 0.0;
                    int t = 0;
                    if (collideHere) {
                        u = pu0[j]; v = pu1[j]; w = pu2[j];
                        if (HAS_ROT) e = pe[j];
                        if (MULTI) t = pt[j];
                    } else {
                        const int idx = a.gather ? beg + j : (a.perm ? a.perm[beg + j] : beg + j);
                        u = fin.ux[idx]; v = fin.uy[idx]; w = fin.uz[idx];
                        if (HAS_ROT) e = fin.erot[idx];
                        if (MULTI) t = fin.type[idx];
                    }
                    if (MULTI && t != s) continue;
                    const double cc = u * u + v * v + w * w;
                    acc[0] += u; acc[1] += v; acc[2] += w;
                    acc[3] += u * u; acc[4] += u * v; acc[5] += u * w; acc[6] += v * v; acc[7] += v * w; acc[8] += w * w;
                    acc[9] += cc;
                    acc[10] += cc * u; acc[11] += cc * v; acc[12] += cc * w;
                    acc[13] += 1.0;
                    if (HAS_ROT) { acc[14] += e; acc[15] += e * u; acc[16] += e * v; acc[17] += e * w; }
                }
                const double tot = warp_reduce_transpose<LOG>(acc, lane);  // lane l: total of value l
                // lane k writes moment slot k (DESIGN.md section moments): pull the value that slot needs
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
                a.mom[((size_t)cell * nS + s) * UGF_NMOM + lane] = o;
            }
        }

        // ---- phase C: NTC collisions (noTimeCounter.C:164-312), out of line to keep the streaming part lean ----
        if (collideHere) {
            ntc_collide<HAS_ROT, MULTI>(prm, a, cell, n, pu0, pu1, pu2, pe, pt, lane, wCand, myColl);

            // ---- phase D: write the collided velocities to their final place -------------------------------
            if (useSmem) {
                for (int j = lane; j < n; j += 32) {
                    fin.ux[beg + j] = sU0[j]; fin.uy[beg + j] = sU1[j]; fin.uz[beg + j] = sU2[j];
                    if (HAS_ROT) fin.erot[beg + j] = sE[j];
                }
            }
            __syncwarp();
        }
    }
    const int wc = warp_sum_int(myColl);
    if (lane == 0) {
        if (wCand) atomicAdd(&a.cnt->cand, wCand);
        if (wc) atomicAdd(&a.cnt->coll, (unsigned long long)wc);
    }
}

}  // namespace ugf
