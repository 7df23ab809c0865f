// ugf_move.cuh — parcel move: ballistic advance with face-crossing tracking, patch interactions and the
// per-cell histogram of destination cells fused in.
//
// Replaces Cloud<uniGasParcel>::move + uniGasParcel::move (U/parcels/uniGasParcel.C:35-108), hitWallPatch and
// the wall models of U/boundaries/basic/uniGasPatchBoundary/uniGasPatchBoundary.C:130-403.
// One thread per parcel; parcels are cell-major from the previous step so neighbouring lanes read the same
// face planes (L1 broadcast) and the histogram atomics are aggregated per warp with match.any.
//
// HBM-bound: algorithmic traffic 80 B/parcel (read x,U,cell = 52; write x,cell = 28); U is only rewritten on
// wall/symmetry hits.  All tracking arithmetic is IEEE mul/add/div plus explicit fma() in the same order as the
// oracle (the library is compiled with -fmad=false, so nothing else is contracted), hence final positions and
// cells match it bit for bit.
#pragma once
#include "ugf_common.cuh"
#include "ugf_rng.cuh"
#include "ugf_sort.cuh"
#include "ugf_internal.cuh"

namespace ugf {

// One 256-bit load through the read-only path (LDG.E.256.CONSTANT, sm_100+): a lane that sits alone in its cell pays
// one L1 wavefront per load instruction whatever its width, so the plane records are fetched in as few as possible.
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// 32-byte plane record, stored as {Sx, Sy, S.Cf, Sz}; returned as {x, y, z, w} = {Sx, Sy, Sz, S.Cf}
__device__ __forceinline__ double4 load_plane(const double4* p) {
    double4 r;
    ldg256(reinterpret_cast<const double*>(p), r.x, r.y, r.w, r.z);
    return r;
}

__device__ __forceinline__ void unit_normal(const double4& pl, double nw[3], double& fA) {
    fA = sqrt(pl.x * pl.x + pl.y * pl.y + pl.z * pl.z);
    nw[0] = pl.x / fA; nw[1] = pl.y / fA; nw[2] = pl.z / fA;
}

// equipartitionRotationalEnergy (U/clouds/uniGasCloud.C:975-1017)
__device__ inline double equipartition_rotational_energy(Stream& r, double T, int rotDoF) {
    if (rotDoF < 1) return 0.0;
    if (rotDoF == 2) return -log(1.0 - r.u01()) * kB * T;
    const double a = 0.5 * rotDoF - 1;
    double energyRatio, Pp;
    const double eps = r.u01();
    do {
        energyRatio = 10 * r.u01();
        Pp = pow(energyRatio / a, a) * exp(a - energyRatio);
    } while (Pp < eps);
    return energyRatio * kB * T;
}

struct WallPre { double IE; double mom[3]; };

// measurePropertiesBeforeControl / AfterControl (uniGasPatchBoundary.C:130-302): slots in DESIGN.md §walls
__device__ inline void measure_wall(const DevParams& prm, double* bm, int bfi, int cell, const DevSpecies& s, const double U[3],
                                    double erot, const double nw[3], double fA, WallPre& pre, bool after, double evib = 0.0, double eelIn = -1.0,
                                    double hitY = 0.0, double hitZ = 0.0) {
    const double eel = eelIn >= 0.0 ? eelIn : s.E0;  // electronic energy of the parcel's level (ground level when the species has one)
    const double m = s.mass;
    const double Un = dot3(U[0], U[1], U[2], nw[0], nw[1], nw[2]);
    const double inv = 1.0 / fmax(fabs(Un) * fA, VSMALL);
    const double UU = dot3(U[0], U[1], U[2], U[0], U[1], U[2]);
    double* b = bm + (size_t)bfi * UGF_NBM;
    atomicAdd(&b[0], inv);
    atomicAdd(&b[1], m * inv);
    atomicAdd(&b[2], 0.5 * m * UU * inv);
    atomicAdd(&b[3], m * U[0] * inv);
    atomicAdd(&b[4], m * U[1] * inv);
    atomicAdd(&b[5], m * U[2] * inv);
    if (s.rotDoF > 0) {
        atomicAdd(&b[6], erot * inv);
        atomicAdd(&b[7], s.rotDoF * inv);
        atomicAdd(&b[12], inv);
    }
    if (evib != 0.0) atomicAdd(&b[13], evib * inv);
    if (eel != 0.0) atomicAdd(&b[14], eel * inv);
    const double IE = 0.5 * m * UU + erot + eel + evib;
    if (!after) {
        pre.IE = IE;
        pre.mom[0] = m * U[0]; pre.mom[1] = m * U[1]; pre.mom[2] = m * U[2];
        atomicAdd(&b[15], 1.0);
    } else {
        double nPart = cell_fn(prm, cell);  // nParticle * CWF of the wall cell * RWF(hit position) (uniGasPatchBoundary.C:292-294)
        if (prm.axi) nPart = nPart * axi_rwf(prm, hitY, hitZ);
        const double dq = nPart * (pre.IE - IE) / (prm.deltaT * fA);
        if (dq != 0.0) atomicAdd(&b[8], dq);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double d = nPart * (pre.mom[k] - m * U[k]) / (prm.deltaT * fA);
            if (d != 0.0) atomicAdd(&b[9 + k], d);
        }
    }
}

// diffuseReflection (uniGasPatchBoundary.C:305-388)
__device__ inline void diffuse_reflection(Stream& r, const DevSpecies& s, double U[3], double& erot, const double nw[3],
                                          double T, const double Uw[3]) {
    double Un = dot3(U[0], U[1], U[2], nw[0], nw[1], nw[2]);
    double Ut[3] = {U[0] - Un * nw[0], U[1] - Un * nw[1], U[2] - Un * nw[2]};
    double magUt = sqrt(dot3(Ut[0], Ut[1], Ut[2], Ut[0], Ut[1], Ut[2]));
    while (magUt < SMALL) {
        U[0] = U[0] * (0.8 + 0.2 * r.u01());
        U[1] = U[1] * (0.8 + 0.2 * r.u01());
        U[2] = U[2] * (0.8 + 0.2 * r.u01());
        Un = dot3(U[0], U[1], U[2], nw[0], nw[1], nw[2]);
        for (int k = 0; k < 3; ++k) Ut[k] = U[k] - Un * nw[k];
        magUt = sqrt(dot3(Ut[0], Ut[1], Ut[2], Ut[0], Ut[1], Ut[2]));
        if (dot3(U[0], U[1], U[2], U[0], U[1], U[2]) == 0.0) {  // reference would spin forever on U == 0
            const double a0 = fabs(nw[0]), a1 = fabs(nw[1]), a2 = fabs(nw[2]);
            const int kmin = a0 <= a1 ? (a0 <= a2 ? 0 : 2) : (a1 <= a2 ? 1 : 2);
            double e[3] = {0, 0, 0};
            e[kmin] = 1.0;
            const double en = dot3(e[0], e[1], e[2], nw[0], nw[1], nw[2]);
            for (int k = 0; k < 3; ++k) Ut[k] = e[k] - en * nw[k];
            magUt = sqrt(dot3(Ut[0], Ut[1], Ut[2], Ut[0], Ut[1], Ut[2]));
        }
    }
    const double tw1[3] = {Ut[0] / magUt, Ut[1] / magUt, Ut[2] / magUt};
    const double tw2[3] = {nw[1] * tw1[2] - nw[2] * tw1[1], nw[2] * tw1[0] - nw[0] * tw1[2], nw[0] * tw1[1] - nw[1] * tw1[0]};
    double g1, g2;
    r.gauss2(g1, g2);
    const double gn = sqrt(-2.0 * log(fmax(1 - r.u01(), VSMALL)));
    const double c = sqrt(kB * T / s.mass);
    for (int k = 0; k < 3; ++k) U[k] = c * (g1 * tw1[k] + g2 * tw2[k] - gn * nw[k]);
    erot = equipartition_rotational_energy(r, T, s.rotDoF);
    for (int k = 0; k < 3; ++k) U[k] += Uw[k];
}

// uniGasCLLWallPatch::controlParticle (uniGasCLLWallPatch.C:80-254), same restatement as the oracle's cllReflection
__device__ inline void cll_reflection(Stream& r, const DevSpecies& s, double U[3], double& erot, const double nw[3], const DevPatch& pt) {
    bool degenerate = false;
    double Un = dot3(U[0], U[1], U[2], nw[0], nw[1], nw[2]);
    double Ut[3] = {U[0] - Un * nw[0], U[1] - Un * nw[1], U[2] - Un * nw[2]};
    double magUt = sqrt(dot3(Ut[0], Ut[1], Ut[2], Ut[0], Ut[1], Ut[2]));
    while (magUt < SMALL) {
        U[0] = U[0] * (0.8 + 0.2 * r.u01());
        U[1] = U[1] * (0.8 + 0.2 * r.u01());
        U[2] = U[2] * (0.8 + 0.2 * r.u01());
        Un = dot3(U[0], U[1], U[2], nw[0], nw[1], nw[2]);
        for (int k = 0; k < 3; ++k) Ut[k] = U[k] - Un * nw[k];
        magUt = sqrt(dot3(Ut[0], Ut[1], Ut[2], Ut[0], Ut[1], Ut[2]));
        if (dot3(U[0], U[1], U[2], U[0], U[1], U[2]) == 0.0) {
            const double a0 = fabs(nw[0]), a1 = fabs(nw[1]), a2 = fabs(nw[2]);
            const int kmin = a0 <= a1 ? (a0 <= a2 ? 0 : 2) : (a1 <= a2 ? 1 : 2);
            double e[3] = {0, 0, 0};
            e[kmin] = 1.0;
            const double en = dot3(e[0], e[1], e[2], nw[0], nw[1], nw[2]);
            for (int k = 0; k < 3; ++k) Ut[k] = e[k] - en * nw[k];
            magUt = sqrt(dot3(Ut[0], Ut[1], Ut[2], Ut[0], Ut[1], Ut[2]));
            degenerate = true;
            break;
        }
    }
    const double tw1[3] = {Ut[0] / magUt, Ut[1] / magUt, Ut[2] / magUt};
    const double tw2[3] = {nw[1] * tw1[2] - nw[2] * tw1[1], nw[2] * tw1[0] - nw[0] * tw1[2], nw[0] * tw1[1] - nw[1] * tw1[0]};
    const double T = pt.T;
    const double alphaT = pt.sigmaT * (2.0 - pt.sigmaT), alphaN = pt.alphaN, alphaR = pt.alphaR;
    const double cmp = sqrt(2.0 * kB * T / s.mass);
    const double utN = degenerate ? 0.0 : magUt / cmp;
    const double unN = Un / cmp;
    const double thetaNormal = TWO_PI * r.u01();
    const double rNormal = sqrt(-alphaN * log(fmax(1 - r.u01(), VSMALL)));
    const double thetaTangential = TWO_PI * r.u01();
    const double rTangential = sqrt(-alphaT * log(fmax(1 - r.u01(), VSMALL)));
    const double um = sqrt(1.0 - alphaN) * unN;
    const double vN = sqrt(rNormal * rNormal + um * um + 2.0 * rNormal * um * cos(thetaNormal));
    const double vT1 = sqrt(1.0 - alphaT) * utN + rTangential * cos(thetaTangential);
    const double vT2 = rTangential * sin(thetaTangential);
    double V[3];
    for (int k = 0; k < 3; ++k) V[k] = cmp * (vT1 * tw1[k] + vT2 * tw2[k] - vN * nw[k]);
    const double wN = dot3(pt.Uw[0], pt.Uw[1], pt.Uw[2], nw[0], nw[1], nw[2]);
    const double w1 = dot3(pt.Uw[0], pt.Uw[1], pt.Uw[2], tw1[0], tw1[1], tw1[2]);
    const double w2 = dot3(pt.Uw[0], pt.Uw[1], pt.Uw[2], tw2[0], tw2[1], tw2[2]);
    const double uN = dot3(V[0], V[1], V[2], nw[0], nw[1], nw[2]);
    const double u1 = dot3(V[0], V[1], V[2], tw1[0], tw1[1], tw1[2]);
    const double u2 = dot3(V[0], V[1], V[2], tw2[0], tw2[1], tw2[2]);
    for (int k = 0; k < 3; ++k)
        U[k] = (uN * nw[k] + wN * nw[k] * alphaN) + (u1 * tw1[k] + w1 * tw1[k] * alphaT) + (u2 * tw2[k] + w2 * tw2[k] * alphaT);
    if (s.rotDoF == 2) {
        const double om = sqrt(erot * (1.0 - alphaR) / (kB * T));
        const double rRot = sqrt(-alphaR * log(fmax(1.0 - r.u01(), VSMALL)));
        const double thetaRot = TWO_PI * r.u01();
        erot = kB * T * (rRot * rRot + om * om + 2.0 * rRot * om * cos(thetaRot));
    } else if (s.rotDoF == 3) {
        double X, A;
        do {
            X = 4.0 * r.u01();
            A = 2.7182818 * X * X * exp(-(X * X));
        } while (A < r.u01());
        const double om = sqrt(erot * (1.0 - alphaR) / (kB * T));
        const double rRot = sqrt(alphaR) * X;
        const double thetaRot = 2.0 * r.u01() - 1.0;
        erot = kB * T * (rRot * rRot + om * om + 2.0 * rRot * om * cos(thetaRot));
    }
}

struct MoveArgs {
    MeshDev mesh;
    ParcelBuf P;
    const long long* dN;   // device: current array length
    long long begin;       // first parcel index of this launch
    long long newFrom;     // parcels with index >= newFrom were inserted this step (newParcel == 1)
    double* sf;            // per-parcel step fraction: in for received parcels, out for migrants (may be null)
    int useSfIn;           // 1: start from sf[i] (received parcels)
    uint32_t step;
    uint32_t aux;          // stream aux: 0 for the step's first move, 1 for received parcels
    int* cellCount;        // [nCells] histogram of destination cells
    int* migCount;         // [nPatches]
    unsigned long long* inflight;  // parcels waiting on processor patches, all patches
    const long long* dBegin;       // if set: the launch covers the parcels from index *dBegin on (received parcels)
    MigSlots ms;                   // processor patches in slot order
    int* migList;                  // [MIG_MAXP][migListCap] indices of the parcels waiting on each processor patch (or null)
    int migListCap;
    double* bm;            // [nBFaces][UGF_NBM]
    DevCounters* cnt;
    double* wq;            // cell weighting + processor patches: the factor a parcel in flight carries (out for migrants, in for received)
    const int* slotTrack;  // face tracker: per face slot, +-(k+1) if the face is entry k of the tracked list (+: the slot's cell owns the
                           // face), 0 otherwise; null = no tracker
    const int* bfTrack;    // per boundary face: k+1 of the face the crossing is booked on (cyclic: the partner face), 0 = not tracked
    double* ft;            // [nTracked][nSpecies][UGF_NFT] running tallies
    uint8_t* nclone;       // cell weighting: clones each parcel gets from cellWeighting() (written for every parcel), else null
    double* queueD;        // move_stream2_kernel: per-warp queues of parcels in mid-track (7 * MQ_CAP doubles, 3 * MQ_CAP ints per warp)
    int* queueI;
};

// The weight (CWF, times RWF with axisymmetricSimulation) a parcel carries into this step's weighting pass: the factor of the
// cell it started the step in (still in P.cell; the previous field while a new one is pending) and RWF of the position it
// started from (still in P.y / P.z) - or RWF of that cell's centre for parcels the inflow inserted this step and for uploaded
// parcels that said so (DevParams::rwfCentre).  Must be called before the move writes the parcel back.
__device__ __noinline__ double carried_weight(const DevParams& prm, const MoveArgs& a, long long i) {
    const int c0 = a.P.cell[i];
    double w = (prm.cwfDirty && i < a.newFrom) ? __ldg(&prm.cwfPrev[c0]) : __ldg(&prm.cwf[c0]);
    if (prm.axi) w = w * ((i >= a.newFrom || prm.rwfCentre) ? __ldg(&prm.cellRwf[c0]) : axi_rwf(prm, a.P.y[i], a.P.z[i]));
    return w;
}

// uniGasFaceTracker::updateFields (U/faceTracker/uniGasFaceTracker.C:90-152) for one crossing of a tracked face: number,
// mass, momentum and energy carried through it, weighted with the parcel's cell weight factor and signed with the
// direction of travel relative to the face area vector (momentum unsigned, as in the reference).  Rare path (only the
// faces of the registered face zones), out of line.
__device__ __noinline__ void face_tally(const DevParams& prm, const MoveArgs& a, long long i, int trk, double U0, double U1, double U2, int type,
                                        double erot, bool haveErot, bool hasRot, double posY = 0.0, double posZ = 0.0) {
    const int k = (trk > 0 ? trk : -trk) - 1;
    const double sgn = trk > 0 ? 1.0 : -1.0;
    double w = 1.0;
    if (prm.cwf) {
        if (a.useSfIn) w = a.wq[i];
        else {
            const int c0 = a.P.cell[i];
            w = (prm.cwfDirty && i < a.newFrom) ? __ldg(&prm.cwfPrev[c0]) : __ldg(&prm.cwf[c0]);
        }
        if (prm.axi) w = w * axi_rwf(prm, posY, posZ);  // the carried CWF times RWF(position at the crossing) (uniGasFaceTracker.C:98-99)
    }
    const DevSpecies& s = prm.sp[type];
    if (hasRot && !haveErot) erot = a.P.erot[i];
    double eInt = s.E0;
    if (prm.spi) {
        eInt = prm.spi[type].elecE[a.P.elev ? a.P.elev[i] : 0];
        if (a.P.vib) eInt += vib_energy(prm.spi[type], s.vibDoF, a.P.vib[i]);
    }
    const double e = 0.5 * s.mass * (U0 * U0 + U1 * U1 + U2 * U2) + (hasRot ? erot : 0.0) + eInt;
    double* t = a.ft + ((size_t)k * prm.nSpecies + type) * UGF_NFT;
    atomicAdd(&t[0], sgn * w);
    atomicAdd(&t[1], sgn * s.mass * w);
    atomicAdd(&t[2], s.mass * U0 * w);
    atomicAdd(&t[3], s.mass * U1 * w);
    atomicAdd(&t[4], s.mass * U2 * w);
    atomicAdd(&t[5], sgn * e * w);
}

// diffuseReflection's tail (uniGasPatchBoundary.C:373-383): vibrational and electronic levels redrawn at the wall temperature, from
// the same stream right after the rotational energy; the new levels go straight to the parcel arrays
__device__ __noinline__ void wall_internal_levels(const DevParams& prm, const MoveArgs& a, Stream& r, int type, double T, long long i, double& evib,
                                                  double& eel) {
    const DevSpecies& sp = prm.sp[type];
    const DevSpeciesInt& S = prm.spi[type];
    if (sp.vibDoF > 0) {
        const unsigned long long v = equipartition_vib_levels(r, T, S, sp.vibDoF);
        a.P.vib[i] = v;
        evib = vib_energy(S, sp.vibDoF, v);
    }
    if (sp.nElec > 1) {
        const int lev = equipartition_elec_level(r, T, S, sp.nElec);
        a.P.elev[i] = (uint8_t)lev;
        eel = S.elecE[lev];
    }
}

enum { HIT_CHANGED_U = 1, HIT_DELETED = 2, HIT_STUCK = 8, HIT_MIGRATED = 16, HIT_WDELETED = 32 };

struct HitState {
    double x[3], U[3], erot, sf;
    int cell, nDraws, flags, nWall;
};

// Boundary-face interaction: the rare path of the tracking loop, kept out of line so that the hot loop
// (internal-face hops) stays small in registers.  The parcel's Philox stream is rebuilt here at draw
// position nDraws (counter-based: no state has to live across the hot loop).
template <bool HAS_ROT, bool MULTI>
__device__ __forceinline__ void boundary_interaction(const DevParams& prm, const MoveArgs& a, int bfi, int hitSlot, long long i, int type, HitState& st) {
    const int patch = __ldg(&a.mesh.bfPatch[bfi]);
    const DevPatch& pt = a.mesh.patches[patch];
    const DevSpecies& sp = prm.sp[type];
    if (pt.kind == UGF_PATCH_WALL) {
        st.nWall++;
        if (pt.wallModel == UGF_WALL_DELETION) {
            st.cell = -1; st.flags |= HIT_DELETED;
            return;
        }
        const double4 pl = load_plane(&a.mesh.plane[hitSlot]);
        double nw[3], fA;
        unit_normal(pl, nw, fA);
        WallPre pre;
        double evib = 0.0, eel = -1.0;  // internal modes beyond rotation (rare gases: DevParams::spi), read from / written to global memory here
        if (prm.spi) {
            if (a.P.vib) evib = vib_energy(prm.spi[type], sp.vibDoF, a.P.vib[i]);
            eel = prm.spi[type].elecE[a.P.elev ? a.P.elev[i] : 0];
        }
        if (prm.measureWalls) measure_wall(prm, a.bm, bfi, st.cell, sp, st.U, st.erot, nw, fA, pre, false, evib, eel, st.x[1], st.x[2]);
        bool diffuse = (pt.wallModel == UGF_WALL_DIFFUSE);
        const bool cll = (pt.wallModel == UGF_WALL_CLL);
        if (pt.wallModel != UGF_WALL_SPECULAR) {
            Stream r(prm.seed, KIND_MOVE, a.aux, a.step, (uint32_t)i, 0);
            r.c3 = (uint32_t)(st.nDraws >> 1);
            if (st.nDraws & 1) { r.block(); r.have = 1; }
            if (pt.wallModel == UGF_WALL_MIXED) diffuse = (pt.diffuseFraction > r.u01());
            if (pt.faceT) {  // …WallFieldPatch.C:108-114: this face's boundaryT / boundaryU
                DevPatch pf = pt;
                const int lf = bfi - pt.startBfi;
                pf.T = __ldg(&pt.faceT[lf]);
                for (int k = 0; k < 3; ++k) pf.Uw[k] = __ldg(&pt.faceU[3 * (size_t)lf + k]);
                if (cll) cll_reflection(r, sp, st.U, st.erot, nw, pf);
                else if (diffuse) diffuse_reflection(r, sp, st.U, st.erot, nw, pf.T, pf.Uw);
                if (diffuse && !cll && prm.spi) wall_internal_levels(prm, a, r, type, pf.T, i, evib, eel);
            } else if (cll) cll_reflection(r, sp, st.U, st.erot, nw, pt);
            else if (diffuse) {
                diffuse_reflection(r, sp, st.U, st.erot, nw, pt.T, pt.Uw);
                if (prm.spi) wall_internal_levels(prm, a, r, type, pt.T, i, evib, eel);
            }
            st.nDraws = 2 * (int)r.c3 - r.have;
        }
        if (!diffuse && !cll) {
            const double Un = dot3(st.U[0], st.U[1], st.U[2], nw[0], nw[1], nw[2]);
            if (Un > 0.0) for (int k = 0; k < 3; ++k) st.U[k] = st.U[k] - 2.0 * Un * nw[k];
        }
        st.flags |= HIT_CHANGED_U;
        if (prm.measureWalls) measure_wall(prm, a.bm, bfi, st.cell, sp, st.U, st.erot, nw, fA, pre, true, evib, eel, st.x[1], st.x[2]);
    } else if (pt.kind == UGF_PATCH_SYMMETRY) {
        const double4 pl = load_plane(&a.mesh.plane[hitSlot]);
        double nw[3], fA;
        unit_normal(pl, nw, fA);
        const double Un = dot3(st.U[0], st.U[1], st.U[2], nw[0], nw[1], nw[2]);
        for (int k = 0; k < 3; ++k) st.U[k] = st.U[k] - 2.0 * Un * nw[k];
        st.flags |= HIT_CHANGED_U;
    } else if (pt.kind == UGF_PATCH_CYCLIC) {
        st.cell = __ldg(&a.mesh.bfOwner[pt.partnerStartBfi + (bfi - pt.startBfi)]);
        for (int k = 0; k < 3; ++k) st.x[k] = st.x[k] + pt.sep[k];
    } else if (pt.kind == UGF_PATCH_PROCESSOR) {
        for (int k = 0; k < 3; ++k) st.x[k] = st.x[k] + pt.sep[k];
        st.cell = -2 - bfi;
        st.flags |= HIT_MIGRATED;
        if (a.sf) a.sf[i] = st.sf;
        if (prm.cwf && a.wq && !a.useSfIn) a.wq[i] = carried_weight(prm, a, i);  // first hop: the weight the parcel started the step with
        const int pos = atomicAdd(&a.migCount[patch], 1);
        if (a.migList) {  // remember who waits here: the pack kernel then never has to search the whole cloud
            int slot = -1;
#pragma unroll
            for (int k = 0; k < MIG_MAXP; ++k) if (k < a.ms.nProc && a.ms.patch[k] == patch) slot = k;
            if (slot >= 0 && pos < a.migListCap) a.migList[(size_t)slot * a.migListCap + pos] = (int)i;
        }
        atomicAdd(a.inflight, 1ull);
    } else if (pt.kind == UGF_PATCH_GENERIC) {
        if (pt.outFlux) {  // uniGasFaceTracker::updateFields on the faces of a mass-flow-rate inlet (uniGasFaceTracker.C:98-141): sgn CWF RWF
            double w = 1.0;
            if (prm.cwf) {
                if (a.useSfIn) w = a.wq[i];
                else {
                    const int c0 = a.P.cell[i];
                    w = (prm.cwfDirty && i < a.newFrom) ? __ldg(&prm.cwfPrev[c0]) : __ldg(&prm.cwf[c0]);
                }
                if (prm.axi) w = w * axi_rwf(prm, st.x[1], st.x[2]);
            }
            const double4 pl = load_plane(&a.mesh.plane[hitSlot]);
            const double un = st.U[0] * pl.x + st.U[1] * pl.y + st.U[2] * pl.z;
            atomicAdd(&pt.outFlux[(size_t)(bfi - pt.startBfi) * prm.nSpecies + type], (un >= 0.0 ? 1.0 : -1.0) * w);
        }
        st.cell = -1; st.flags |= HIT_DELETED;
    } else {
        st.cell = -1; st.flags |= HIT_STUCK;
    }
}

template <bool HAS_ROT, bool MULTI>
__device__ __noinline__ void boundary_hit(const DevParams& prm, const MoveArgs& a, int bfi, int hitSlot, long long i, int type, HitState& st) {
    boundary_interaction<HAS_ROT, MULTI>(prm, a, bfi, hitSlot, i, type, st);
    if (a.slotTrack) {  // face tracker: boundary faces are booked after the patch interaction, with the sign of the velocity
        const int trk = __ldg(&a.bfTrack[bfi]);  // the parcel has now (uniGasFaceTracker.C:107-135); cyclic: on the partner face
        if (trk) {
            const double4 pl = load_plane(&a.mesh.plane[hitSlot]);
            const double un = st.U[0] * pl.x + st.U[1] * pl.y + st.U[2] * pl.z;
            face_tally(prm, a, i, un >= 0.0 ? trk : -trk, st.U[0], st.U[1], st.U[2], type, st.erot, true, HAS_ROT, st.x[1], st.x[2]);
        }
    }
}

// uniGasCloud::cellWeighting for one parcel (U/clouds/uniGasCloud.C:1360-1421): the parcel carried wOld and now sits
// in a cell whose factor is wNew.  Returns the number of clones (>= 0) or -1 = delete.  The rare path of the move
// (only parcels whose factor changes), out of line; one uniform from the parcel's own stream.
__device__ __noinline__ int weighting_decision(uint64_t seed, uint32_t step, uint32_t i, double wOld, double wNew) {
    Stream r(seed, KIND_WEIGHT, 0, step, i, 0);
    if (wOld > wNew) {
        double prob = wOld / wNew - 1.0;
        int k = 0;
        while (prob > 1.0 && k < 254) { ++k; prob -= 1.0; }
        if (prob > r.u01()) ++k;
        return k;
    }
    return (wOld / wNew < r.u01()) ? -1 : 0;
}

// ---- bulk-copy (TMA) staging helpers ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Warp-collective variants: every lane calls with the same arguments and one elected lane issues (elect.sync), so the call
// site stays convergent and ptxas needs no per-lane loop around the uniform-datapath bulk copy.
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* bar, unsigned bytes) {
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n @p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_elect(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n @p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

// One parcel's track (thread i), its stores and the warp-aggregated histogram / counter updates.  Every lane of
// the warp must call it (valid = false for lanes without a parcel).
// NF > 0: every cell has exactly NF face slots (hex meshes: 6, or 4 after the never-hit faces of an empty
// direction were pruned at set-up): no offset loads and a fully unrolled face loop, so all plane loads of a hop
// and the ids of the cells behind them are issued together (one dependent memory level per hop).
// NF == 0: general polyhedra through the CSR offsets.
// NFT == NF_REC2D: 4 slots per cell read from the packed 2-D record (mesh.rec2d).
constexpr int NF_REC2D = 14;
template <bool HAS_ROT, bool MULTI, int NFT>
__device__ __forceinline__ void track_parcel(const DevParams& prm, const MoveArgs& a, const long long i, const bool valid, int cell, double x0,
                                             double x1, double x2, double U0, double U1, double U2) {
    constexpr bool REC2D = (NFT == NF_REC2D);
    constexpr int NF = REC2D ? 4 : NFT;
    int flags = 0, nWall = 0, nClone = 0;
    if (valid && cell >= 0) {
        int nDraws = 0;
        double sf = 0.0;
        if (a.useSfIn) sf = a.sf[i];
        else if (i >= a.newFrom) {  // U/parcels/uniGasParcel.C:47-53
            Stream r(prm.seed, KIND_MOVE, a.aux, a.step, (uint32_t)i, 0);
            sf = r.u01();
            nDraws = 1;
        }
        const double dt = prm.deltaT;
        const bool s0 = prm.solD[0] != 0, s1 = prm.solD[1] != 0, s2 = prm.solD[2] != 0;
        int iters = 0;
        double erot = 0.0;
        bool erotLoaded = false;
        while (cell >= 0 && sf < 1) {
            const double rem = 1 - sf;
            const double s = rem * dt;
            const double d0 = s0 ? s * U0 : 0.0, d1 = s1 ? s * U1 : 0.0, d2 = s2 ? s * U2 : 0.0;  // constrainDirection
            // first face crossed: min of num/nd over faces with nd > 0, compared by cross-multiplication
            // (one division per hop, branch-free loop body); (bnum, bnd) = (1, 1) encodes "end of step"
            double bnum = 1.0, bnd = 1.0;
            int hit = -1, nb = 0;
            if (NF > 0) {
                const int jb = cell * NF;
                double4 pl[NF > 0 ? NF : 1];
                int nbs[NF > 0 ? NF : 1];
                if (REC2D) {
                    // straight 2-D mesh: the cell's whole record {Sx,Sy}x4, {S.Cf}x4, nbr x4 is one 128-byte line,
                    // fetched with three 256-bit loads and one 128-bit load
                    const double* rp = a.mesh.rec2d + (size_t)cell * 16;
                    ldg256(rp, pl[0].x, pl[0].y, pl[1].x, pl[1].y);
                    ldg256(rp + 4, pl[2].x, pl[2].y, pl[3].x, pl[3].y);
                    ldg256(rp + 8, pl[0].w, pl[1].w, pl[2].w, pl[3].w);
                    const int4 v = __ldg(reinterpret_cast<const int4*>(rp + 12));
                    nbs[0] = v.x; nbs[1] = v.y; nbs[2] = v.z; nbs[3] = v.w;
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) pl[f] = load_plane(&a.mesh.plane[jb + f]);
                }
                if (REC2D) {
                } else if (NF == 4) {
                    const int4 v = __ldg(reinterpret_cast<const int4*>(a.mesh.nbr + jb));
                    nbs[0] = v.x; nbs[1] = v.y; nbs[2] = v.z; nbs[3] = v.w;
                } else {
#pragma unroll
                    for (int f = 0; f < NF; f += 2) {
                        const int2 v = __ldg(reinterpret_cast<const int2*>(a.mesh.nbr + jb + f));
                        nbs[f] = v.x; nbs[f + 1] = v.y;
                    }
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    // REC2D: Sz == 0 exactly, and fma(0, finite, t) == t, so the z terms are dropped (same bits)
                    const double nd = REC2D ? fma(pl[f].y, d1, pl[f].x * d0) : fma(pl[f].z, d2, fma(pl[f].y, d1, pl[f].x * d0));
                    double num = pl[f].w - (REC2D ? fma(pl[f].y, x1, pl[f].x * x0) : fma(pl[f].z, x2, fma(pl[f].y, x1, pl[f].x * x0)));
                    num = num < 0 ? 0.0 : num;
                    const bool better = (nd > 0) && (num * bnd < bnum * nd);
                    bnum = better ? num : bnum;
                    bnd = better ? nd : bnd;
                    hit = better ? jb + f : hit;
                    nb = better ? nbs[f] : nb;
                }
            } else {
                const int jb = __ldg(&a.mesh.cfOff[cell]), je = __ldg(&a.mesh.cfOff[cell + 1]);
                for (int j = jb; j < je; ++j) {
                    const double4 pl = load_plane(&a.mesh.plane[j]);
                    const int nbj = __ldg(&a.mesh.nbr[j]);
                    const double nd = fma(pl.z, d2, fma(pl.y, d1, pl.x * d0));
                    double num = pl.w - fma(pl.z, x2, fma(pl.y, x1, pl.x * x0));
                    num = num < 0 ? 0.0 : num;
                    const bool better = (nd > 0) && (num * bnd < bnum * nd);
                    bnum = better ? num : bnum;
                    bnd = better ? nd : bnd;
                    hit = better ? j : hit;
                    nb = better ? nbj : nb;
                }
            }
            if (hit < 0) {
                x0 = x0 + d0; x1 = x1 + d1; x2 = x2 + d2;
                sf = 1;
                break;
            }
            const double lamMin = bnum / bnd;
            x0 = fma(lamMin, d0, x0); x1 = fma(lamMin, d1, x1); x2 = fma(lamMin, d2, x2);
            sf = fma(rem, lamMin, sf);
            if (nb >= 0) {
                if (a.slotTrack) {  // face tracker on: is this one of the registered faces?
                    const int trk = __ldg(&a.slotTrack[hit]);
                    if (trk) {
                        int type = 0;
                        if (MULTI) type = a.P.type[i];
                        face_tally(prm, a, i, trk, U0, U1, U2, type, erot, erotLoaded, HAS_ROT, x1, x2);
                    }
                }
                cell = nb;
            } else {
                HitState st;
                st.x[0] = x0; st.x[1] = x1; st.x[2] = x2;
                st.U[0] = U0; st.U[1] = U1; st.U[2] = U2;
                if (HAS_ROT && !erotLoaded) { erot = a.P.erot[i]; erotLoaded = true; }
                st.erot = erot; st.sf = sf; st.cell = cell; st.nDraws = nDraws; st.flags = flags; st.nWall = nWall;
                int type = 0;
                if (MULTI) type = a.P.type[i];
                boundary_hit<HAS_ROT, MULTI>(prm, a, -nb - 1, hit, i, type, st);
                x0 = st.x[0]; x1 = st.x[1]; x2 = st.x[2];
                U0 = st.U[0]; U1 = st.U[1]; U2 = st.U[2];
                erot = st.erot; cell = st.cell; nDraws = st.nDraws; flags = st.flags; nWall = st.nWall;
            }
            if (++iters > MAX_TRACK_ITERS) { cell = -1; flags |= HIT_STUCK; break; }
        }
        if (prm.cwf) {  // weighting() right after the move (U/clouds/uniGasCloud.C:839-842): old factor = the one of the
                        // cell the parcel started the step in (re-read before it is overwritten), new = where it stopped
            int k = 0;
            if (cell >= 0) {
                double wNew = __ldg(&prm.cwf[cell]);
                double wOld;
                if (a.useSfIn) {
                    wOld = a.wq[i];  // received parcel: the factor travelled with it
                } else if (prm.axi) {
                    wOld = carried_weight(prm, a, i);
                } else {
                    const int cell0 = a.P.cell[i];
                    wOld = (prm.cwfDirty && i < a.newFrom) ? __ldg(&prm.cwfPrev[cell0]) : __ldg(&prm.cwf[cell0]);
                }
                if (prm.axi) wNew = wNew * axi_rwf(prm, x1, x2);  // axisymmetric(Cell)Weighting (uniGasCloud.C:1427-1570): RWF of where it stopped
                if (wOld != wNew) k = weighting_decision(prm.seed, a.step, (uint32_t)i, wOld, wNew);
                if (k < 0) { cell = -1; flags |= HIT_WDELETED; k = 0; }
            }
            a.nclone[i] = (uint8_t)k;
            nClone = k;
        }
        a.P.x[i] = x0; a.P.y[i] = x1; a.P.z[i] = x2;
        a.P.cell[i] = cell;
        if (flags & HIT_CHANGED_U) {
            a.P.ux[i] = U0; a.P.uy[i] = U1; a.P.uz[i] = U2;
            if (HAS_ROT) a.P.erot[i] = erot;
        }
    }
    // histogram of destination cells, aggregated over lanes that landed in the same cell
    const int lane = threadIdx.x & 31;
    const bool live = valid && cell >= 0;
    {
        int head, cnt, rank;
        warp_runs(live ? cell : -1, lane, head, cnt, rank);
        if (live && rank == 0) atomicAdd(&a.cellCount[cell], cnt);
        if (nClone) atomicAdd(&a.cellCount[cell], nClone);  // clones take slots of the same cell
    }
    if (prm.cwf && __any_sync(0xffffffffu, (nClone | (flags & HIT_WDELETED)) != 0)) {
        const int sc = warp_sum_int(nClone);
        const int swd = __popc(__ballot_sync(0xffffffffu, (flags & HIT_WDELETED) != 0));
        if (lane == 0) {
            if (sc) atomicAdd(&a.cnt->cloned, (unsigned long long)sc);
            if (swd) atomicAdd(&a.cnt->wdeleted, (unsigned long long)swd);
        }
    }
    if (__any_sync(0xffffffffu, (flags | nWall) != 0)) {
        const int sd = __popc(__ballot_sync(0xffffffffu, (flags & HIT_DELETED) != 0));
        const int ss = __popc(__ballot_sync(0xffffffffu, (flags & HIT_STUCK) != 0));
        const int sm = __popc(__ballot_sync(0xffffffffu, (flags & HIT_MIGRATED) != 0));
        const int sw = warp_sum_int(nWall);
        if (lane == 0) {
            if (sd) atomicAdd(&a.cnt->deleted, (unsigned long long)sd);
            if (ss) atomicAdd(&a.cnt->stuck, (unsigned long long)ss);
            if (sm) atomicAdd(&a.cnt->migrated, (unsigned long long)sm);
            if (sw) atomicAdd(&a.cnt->wallHits, (unsigned long long)sw);
        }
    }
}

// Direct kernel, one thread per parcel of [begin, n): used for the resumed tracks of received parcels (small,
// unaligned ranges).
template <bool HAS_ROT, bool MULTI, int NF>
__global__ void __launch_bounds__(256, 4) move_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ MoveArgs a) {
    const long long i = (a.dBegin ? *a.dBegin : a.begin) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = *a.dN;
    if (i - threadIdx.x >= n) return;  // whole block beyond the array
    const bool valid = i < n;
    int cell = -1;
    double x0 = 0, x1 = 0, x2 = 0, U0 = 0, U1 = 0, U2 = 0;
    if (valid) cell = a.P.cell[i];
    if (valid && cell >= 0) {
        x0 = a.P.x[i]; x1 = a.P.y[i]; x2 = a.P.z[i];
        U0 = a.P.ux[i]; U1 = a.P.uy[i]; U2 = a.P.uz[i];
    }
    track_parcel<HAS_ROT, MULTI, NF>(prm, a, i, valid, cell, x0, x1, x2, U0, U1, U2);
}

// The step's main move: persistent CTAs (a multiple of the SM count) whose warps each stream tiles of 32 parcels
// through a private MOVE_STAGES-deep ring in shared memory.  The seven per-parcel arrays of a tile (x, y, z, Ux, Uy,
// Uz, cell: 1664 B) are brought in by cp.async.bulk (1-D TMA) copies that complete on the ring slot's mbarrier,
// MOVE_STAGES tiles ahead of the one being tracked, so the HBM stream stays in flight while the lanes sit in the
// dependent cell -> planes -> next-cell load chain of the tracking loop.  Warps never wait for each other (no
// block barrier in the loop).  Tiles are always copied whole (the SoA arrays are allocated in multiples of
// MOVE_TILE), begin must be 0.
constexpr int MOVE_TILE = 256;   // allocation granule of the SoA arrays (parcels)
constexpr int MOVE_WARPS = 8;
constexpr int MOVE_STAGES = 2;

template <bool HAS_ROT, bool MULTI, int NF, int BPS>
__global__ void __launch_bounds__(MOVE_WARPS * 32, BPS) move_stream_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ MoveArgs a) {
    __shared__ __align__(128) double sD[MOVE_WARPS][MOVE_STAGES][6][32];
    __shared__ __align__(16) int sC[MOVE_WARPS][MOVE_STAGES][32];
    __shared__ __align__(8) uint64_t bar[MOVE_WARPS][MOVE_STAGES];
    const long long n = *a.dN;
    const long long nTiles = (n + 31) / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long stride = (long long)gridDim.x * MOVE_WARPS;
    auto issue = [&](long long tile, int s) {
        const long long b = tile * 32;
        constexpr unsigned bytesD = 32 * sizeof(double), bytesC = 32 * sizeof(int);
        uint64_t* br = &bar[w][s];
        mbar_expect_tx(br, 6 * bytesD + bytesC);
        bulk_g2s(sD[w][s][0], a.P.x + b, bytesD, br);
        bulk_g2s(sD[w][s][1], a.P.y + b, bytesD, br);
        bulk_g2s(sD[w][s][2], a.P.z + b, bytesD, br);
        bulk_g2s(sD[w][s][3], a.P.ux + b, bytesD, br);
        bulk_g2s(sD[w][s][4], a.P.uy + b, bytesD, br);
        bulk_g2s(sD[w][s][5], a.P.uz + b, bytesD, br);
        bulk_g2s(sC[w][s], a.P.cell + b, bytesC, br);
    };
    const long long first = (long long)blockIdx.x * MOVE_WARPS + w;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < MOVE_STAGES; ++s) mbar_init(&bar[w][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int s = 0; s < MOVE_STAGES; ++s) {
            const long long t = first + (long long)s * stride;
            if (t < nTiles) issue(t, s);
        }
    }
    __syncwarp();
    int k = 0;
    for (long long tile = first; tile < nTiles; tile += stride, ++k) {
        const int s = k % MOVE_STAGES;
        mbar_wait(&bar[w][s], (unsigned)(k / MOVE_STAGES) & 1u);
        const long long i = tile * 32 + lane;
        const bool valid = i < n;
        const int cell = valid ? sC[w][s][lane] : -1;
        const double x0 = sD[w][s][0][lane], x1 = sD[w][s][1][lane], x2 = sD[w][s][2][lane];
        const double U0 = sD[w][s][3][lane], U1 = sD[w][s][4][lane], U2 = sD[w][s][5][lane];
        __syncwarp();  // the slot is in registers: refill it with the tile MOVE_STAGES ahead
        if (lane == 0) {
            const long long t = tile + (long long)MOVE_STAGES * stride;
            if (t < nTiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t, s);
            }
        }
        track_parcel<HAS_ROT, MULTI, NF>(prm, a, i, valid, cell, x0, x1, x2, U0, U1, U2);
    }
}


// ---- streamed move with hop compaction ---------------------------------------------------------------------------------
// The tracking loop of track_parcel keeps a warp in lockstep over its 32 parcels: the second and third face crossing of a
// tile run with half and a tenth of the lanes (16 of 32 active on average at Courant 0.5, profiles/ r1).  Here a pass of
// the warp does at most MOVE_HOPS hops for 32 parcels.  A parcel whose track is over is finalised at once (its stores, the weighting
// decision, the histogram); one that goes on is pushed - position, velocity, step fraction, cell, index, draw / iteration
// counters - onto the warp's own queue (global memory, L2-resident).  Whenever the queue holds 32 parcels the next pass takes them instead
// of a fresh tile, so second and later hops run on full warps, while the first hop of a tile stays tile-coherent (two or
// three neighbouring cells per warp: broadcast plane loads, coalesced stores, run-aggregated histogram).  No block barrier,
// the queue is private to the warp; results do not depend on the processing order (per-parcel Philox streams).
// On straight 2-D meshes (packed record, z empty) z and Uz are neither staged nor queued nor written: the tracking does not
// touch them (the rare boundary interaction reads them from global memory), which takes 24 of the 80 B per parcel off the bus.
constexpr int MQ_CAP = 64;  // a pass pops 32 when the queue holds >= 32 and pushes <= 32
constexpr int MOVE_HOPS = 2; // hops a pass runs in lockstep before the unfinished parcels are queued

template <bool FLAT>
struct MoveSmem {
    static constexpr int ND = FLAT ? 4 : 6;  // staged double arrays per tile: x, y, (z), Ux, Uy, (Uz)
    static constexpr int NQ = FLAT ? 5 : 7;  // queued doubles per parcel: the same + step fraction
    static constexpr size_t ringBytes = (size_t)MOVE_STAGES * (ND * 32 * sizeof(double) + 32 * sizeof(int));
    static constexpr size_t perWarp = ringBytes;
    static constexpr size_t total = MOVE_WARPS * perWarp + MOVE_WARPS * MOVE_STAGES * sizeof(uint64_t);
    // the queue of a warp lives in global memory (L2-resident: a few MB for the whole grid, touched by one parcel in ten):
    // shared memory would take the L1 lines the plane loads live on (measured: L1 hit rate 68 % -> 38 % with a shared queue)
    static constexpr size_t queueDoubles = (size_t)7 * MQ_CAP;  // per warp, sized for the 3-D layout
    static constexpr size_t queueInts = (size_t)3 * MQ_CAP;
};

constexpr int MISC_DRAW_SHIFT = 13;  // misc = tracking iterations (13 bits) | Philox draws consumed so far << 13

// One hop of one parcel.  Returns true when the track is over (end of step, deleted, waiting on a processor patch, stuck).
template <bool HAS_ROT, bool MULTI, int NFT, bool FLAT>
__device__ __forceinline__ bool hop_once(const DevParams& prm, const MoveArgs& a, const int i, int& cell, double& x0, double& x1, double& x2,
                                         double& U0, double& U1, double& U2, double& sf, int& misc) {
    constexpr bool REC2D = (NFT == NF_REC2D);
    constexpr int NF = REC2D ? 4 : NFT;
    const double rem = 1 - sf;
    const double s = rem * prm.deltaT;
    const double d0 = prm.solD[0] ? s * U0 : 0.0, d1 = prm.solD[1] ? s * U1 : 0.0;
    const double d2 = FLAT ? 0.0 : (prm.solD[2] ? s * U2 : 0.0);  // constrainDirection
    double bnum = 1.0, bnd = 1.0;
    int hit = -1, nb = 0;
    if (NF > 0) {
        const int jb = cell * NF;
        double4 pl[NF > 0 ? NF : 1];
        int nbs[NF > 0 ? NF : 1];
        if (REC2D) {
            const double* rp = a.mesh.rec2d + (size_t)cell * 16;
            ldg256(rp, pl[0].x, pl[0].y, pl[1].x, pl[1].y);
            ldg256(rp + 4, pl[2].x, pl[2].y, pl[3].x, pl[3].y);
            ldg256(rp + 8, pl[0].w, pl[1].w, pl[2].w, pl[3].w);
            const int4 v = __ldg(reinterpret_cast<const int4*>(rp + 12));
            nbs[0] = v.x; nbs[1] = v.y; nbs[2] = v.z; nbs[3] = v.w;
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) pl[f] = load_plane(&a.mesh.plane[jb + f]);
            if (NF == 4) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(a.mesh.nbr + jb));
                nbs[0] = v.x; nbs[1] = v.y; nbs[2] = v.z; nbs[3] = v.w;
            } else {
#pragma unroll
                for (int f = 0; f < NF; f += 2) {
                    const int2 v = __ldg(reinterpret_cast<const int2*>(a.mesh.nbr + jb + f));
                    nbs[f] = v.x; nbs[f + 1] = v.y;
                }
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            // REC2D: Sz == 0 exactly, and fma(0, finite, t) == t, so the z terms are dropped (same bits)
            const double nd = REC2D ? fma(pl[f].y, d1, pl[f].x * d0) : fma(pl[f].z, d2, fma(pl[f].y, d1, pl[f].x * d0));
            double num = pl[f].w - (REC2D ? fma(pl[f].y, x1, pl[f].x * x0) : fma(pl[f].z, x2, fma(pl[f].y, x1, pl[f].x * x0)));
            num = num < 0 ? 0.0 : num;
            const bool better = (nd > 0) && (num * bnd < bnum * nd);
            bnum = better ? num : bnum;
            bnd = better ? nd : bnd;
            hit = better ? jb + f : hit;
            nb = better ? nbs[f] : nb;
        }
    } else {
        const int jb = __ldg(&a.mesh.cfOff[cell]), je = __ldg(&a.mesh.cfOff[cell + 1]);
        for (int j = jb; j < je; ++j) {
            const double4 pl = load_plane(&a.mesh.plane[j]);
            const int nbj = __ldg(&a.mesh.nbr[j]);
            const double nd = fma(pl.z, d2, fma(pl.y, d1, pl.x * d0));
            double num = pl.w - fma(pl.z, x2, fma(pl.y, x1, pl.x * x0));
            num = num < 0 ? 0.0 : num;
            const bool better = (nd > 0) && (num * bnd < bnum * nd);
            bnum = better ? num : bnum;
            bnd = better ? nd : bnd;
            hit = better ? j : hit;
            nb = better ? nbj : nb;
        }
    }
    if (hit < 0) {
        x0 = x0 + d0; x1 = x1 + d1;
        if (!FLAT) x2 = x2 + d2;
        sf = 1;
        return true;
    }
    const double lamMin = bnum / bnd;
    x0 = fma(lamMin, d0, x0); x1 = fma(lamMin, d1, x1);
    if (!FLAT) x2 = fma(lamMin, d2, x2);
    sf = fma(rem, lamMin, sf);
    if (nb >= 0) {
        if (a.slotTrack) {  // face tracker on: is this one of the registered faces?
            const int trk = __ldg(&a.slotTrack[hit]);
            if (trk) {
                int type = 0;
                if (MULTI) type = a.P.type[i];
                face_tally(prm, a, i, trk, U0, U1, FLAT ? a.P.uz[i] : U2, type, 0.0, false, HAS_ROT, x1, FLAT ? 0.0 : x2);
            }
        }
        cell = nb;
    } else {
        HitState st;
        const double zin = FLAT ? a.P.z[i] : x2;
        st.x[0] = x0; st.x[1] = x1; st.x[2] = zin;
        st.U[0] = U0; st.U[1] = U1; st.U[2] = FLAT ? a.P.uz[i] : U2;
        st.erot = 0.0;
        if (HAS_ROT) st.erot = a.P.erot[i];  // an earlier wall hit of this track has stored its result already
        st.sf = sf; st.cell = cell; st.nDraws = misc >> MISC_DRAW_SHIFT; st.flags = 0; st.nWall = 0;
        int type = 0;
        if (MULTI) type = a.P.type[i];
        boundary_hit<HAS_ROT, MULTI>(prm, a, -nb - 1, hit, i, type, st);
        x0 = st.x[0]; x1 = st.x[1];
        if (!FLAT) x2 = st.x[2];
        else if (st.x[2] != zin) a.P.z[i] = st.x[2];
        U0 = st.U[0]; U1 = st.U[1];
        if (!FLAT) U2 = st.U[2];
        if (st.flags & HIT_CHANGED_U) {  // stored now: the queue does not carry what the tracking does not need
            a.P.ux[i] = st.U[0]; a.P.uy[i] = st.U[1]; a.P.uz[i] = st.U[2];
            if (HAS_ROT) a.P.erot[i] = st.erot;
        }
        cell = st.cell;
        misc = (misc & ((1 << MISC_DRAW_SHIFT) - 1)) | (st.nDraws << MISC_DRAW_SHIFT);
        // rare path (a boundary face): the step's counters are bumped right here instead of riding along in registers
        if (st.nWall) atomicAdd(&a.cnt->wallHits, (unsigned long long)st.nWall);
        if (st.flags & HIT_DELETED) atomicAdd(&a.cnt->deleted, 1ull);
        if (st.flags & HIT_STUCK) atomicAdd(&a.cnt->stuck, 1ull);
        if (st.flags & HIT_MIGRATED) atomicAdd(&a.cnt->migrated, 1ull);
    }
    misc += 1;
    if ((misc & ((1 << MISC_DRAW_SHIFT) - 1)) > MAX_TRACK_ITERS) { cell = -1; atomicAdd(&a.cnt->stuck, 1ull); return true; }
    return !(cell >= 0 && sf < 1);
}

template <bool HAS_ROT, bool MULTI, int NFT, int BPS>
__global__ void __launch_bounds__(MOVE_WARPS * 32, BPS) move_stream2_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ MoveArgs a) {
    constexpr bool FLAT = (NFT == NF_REC2D);  // the host selects the packed 2-D record only when z is an empty direction
    using L = MoveSmem<FLAT>;
    constexpr int ND = L::ND;
    extern __shared__ __align__(128) unsigned char moveSmem[];
    // the warp index through a shuffle: the compiler then knows it (and the tile counters, queue length and ring addresses derived
    // from it) to be warp-uniform and keeps them in uniform registers - the bulk copies need no per-lane loop, the loop control no
    // vector registers
    const int lane = threadIdx.x & 31, w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    unsigned char* const base = moveSmem + (size_t)w * L::perWarp;
    double* const ringD = reinterpret_cast<double*>(base);                                                  // [STAGES][ND][32]
    int* const ringC = reinterpret_cast<int*>(base + (size_t)MOVE_STAGES * ND * 32 * sizeof(double));        // [STAGES][32]
    const size_t warpId = (size_t)blockIdx.x * MOVE_WARPS + w;
    double* const qD = a.queueD + warpId * L::queueDoubles;  // [NQ][MQ_CAP]
    int* const qI = a.queueI + warpId * L::queueInts;        // [3][MQ_CAP]
    uint64_t* const bar = reinterpret_cast<uint64_t*>(moveSmem + (size_t)MOVE_WARPS * L::perWarp) + w * MOVE_STAGES;
    const int n = (int)*a.dN;  // parcel capacity < 2^31
    const int nTiles = (n + 31) / 32;
    const int stride = (int)gridDim.x * MOVE_WARPS;
    auto issue = [&](int tile, int s) {
        const size_t b = (size_t)tile * 32;
        constexpr unsigned bytesD = 32 * sizeof(double), bytesC = 32 * sizeof(int);
        uint64_t* br = &bar[s];
        double* d = ringD + (size_t)s * ND * 32;
        mbar_expect_tx_elect(br, ND * bytesD + bytesC);
        bulk_g2s_elect(d, a.P.x + b, bytesD, br);
        bulk_g2s_elect(d + 32, a.P.y + b, bytesD, br);
        if (FLAT) {
            bulk_g2s_elect(d + 64, a.P.ux + b, bytesD, br);
            bulk_g2s_elect(d + 96, a.P.uy + b, bytesD, br);
        } else {
            bulk_g2s_elect(d + 64, a.P.z + b, bytesD, br);
            bulk_g2s_elect(d + 96, a.P.ux + b, bytesD, br);
            bulk_g2s_elect(d + 128, a.P.uy + b, bytesD, br);
            bulk_g2s_elect(d + 160, a.P.uz + b, bytesD, br);
        }
        bulk_g2s_elect(ringC + s * 32, a.P.cell + b, bytesC, br);
    };
    const int first = (int)blockIdx.x * MOVE_WARPS + w;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < MOVE_STAGES; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < MOVE_STAGES; ++s) {
        const int t = first + s * stride;
        if (t < nTiles) issue(t, s);  // warp-collective: one elected lane issues
    }
    int qn = 0, k = 0;
    int tile = first;
    for (;;) {
        bool active = false;
        int i = 0;
        int cell = -1, misc = 0;
        double x0 = 0, x1 = 0, x2 = 0, U0 = 0, U1 = 0, U2 = 0, sf = 0;
        if (qn >= 32 || (tile >= nTiles && qn > 0)) {  // a warp's worth of parcels in mid-track (or the last ones): continue those
            const int take = qn < 32 ? qn : 32;
            const int slot = qn - take + lane;
            active = lane < take;
            if (active) {
                x0 = __ldcg(&qD[slot]); x1 = __ldcg(&qD[MQ_CAP + slot]);
                if (FLAT) { U0 = __ldcg(&qD[2 * MQ_CAP + slot]); U1 = __ldcg(&qD[3 * MQ_CAP + slot]); sf = __ldcg(&qD[4 * MQ_CAP + slot]); }
                else {
                    x2 = __ldcg(&qD[2 * MQ_CAP + slot]); U0 = __ldcg(&qD[3 * MQ_CAP + slot]); U1 = __ldcg(&qD[4 * MQ_CAP + slot]);
                    U2 = __ldcg(&qD[5 * MQ_CAP + slot]); sf = __ldcg(&qD[6 * MQ_CAP + slot]);
                }
                cell = __ldcg(&qI[slot]); i = __ldcg(&qI[MQ_CAP + slot]); misc = __ldcg(&qI[2 * MQ_CAP + slot]);
            }
            qn -= take;
            __syncwarp();  // all reads done before this pass pushes into the same slots
        } else if (tile < nTiles) {
            const int s = k % MOVE_STAGES;
            mbar_wait(&bar[s], (unsigned)(k / MOVE_STAGES) & 1u);
            const double* d = ringD + (size_t)s * ND * 32;
            i = tile * 32 + lane;
            cell = (i < n) ? ringC[s * 32 + lane] : -1;
            x0 = d[lane]; x1 = d[32 + lane];
            if (FLAT) { U0 = d[64 + lane]; U1 = d[96 + lane]; }
            else { x2 = d[64 + lane]; U0 = d[96 + lane]; U1 = d[128 + lane]; U2 = d[160 + lane]; }
            __syncwarp();  // the slot is in registers: refill it with the tile MOVE_STAGES ahead
            {
                const int t = tile + MOVE_STAGES * stride;
                if (t < nTiles) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(t, s);
                }
            }
            active = cell >= 0;
            if (active && i >= a.newFrom) {  // U/parcels/uniGasParcel.C:47-53
                Stream r(prm.seed, KIND_MOVE, a.aux, a.step, (uint32_t)i, 0);
                sf = r.u01();
                misc = 1 << MISC_DRAW_SHIFT;
                if (!(sf < 1)) active = true;
            }
            tile += stride;
            ++k;
        } else {
            break;
        }
        // up to MOVE_HOPS hops in lockstep: the second hop still runs with close to half of the warp; whoever needs more
        // (one parcel in ten at Courant 0.5) waits in the queue for a full warp
        bool done = !active || !(sf < 1);
#pragma unroll 1
        for (int hop = 0; hop < MOVE_HOPS; ++hop) {
            if (!done) done = hop_once<HAS_ROT, MULTI, NFT, FLAT>(prm, a, i, cell, x0, x1, x2, U0, U1, U2, sf, misc);
            if (__all_sync(0xffffffffu, done)) break;
        }
        // ---- finalise the parcels whose track is over ----
        const bool fin = active && done;
        int nClone = 0;
        if (fin) {
            if (prm.cwf) {  // weighting() right after the move (U/clouds/uniGasCloud.C:839-842), see track_parcel
                int kc = 0;
                if (cell >= 0) {
                    double wNew = __ldg(&prm.cwf[cell]);
                    double wOld;
                    if (prm.axi) {
                        wOld = carried_weight(prm, a, i);
                        wNew = wNew * axi_rwf(prm, x1, FLAT ? 0.0 : x2);
                    } else {
                        const int cell0 = a.P.cell[i];
                        wOld = (prm.cwfDirty && i < a.newFrom) ? __ldg(&prm.cwfPrev[cell0]) : __ldg(&prm.cwf[cell0]);
                    }
                    if (wOld != wNew) kc = weighting_decision(prm.seed, a.step, (uint32_t)i, wOld, wNew);
                    if (kc < 0) { cell = -1; atomicAdd(&a.cnt->wdeleted, 1ull); kc = 0; }
                }
                a.nclone[i] = (uint8_t)kc;
                nClone = kc;
                if (kc) atomicAdd(&a.cnt->cloned, (unsigned long long)kc);
            }
            a.P.x[i] = x0; a.P.y[i] = x1;
            if (!FLAT) a.P.z[i] = x2;
            a.P.cell[i] = cell;
        }
        {
            const bool live = fin && cell >= 0;
            int head, cnt, rank;
            warp_runs(live ? cell : -1, lane, head, cnt, rank);
            if (live && rank == 0) atomicAdd(&a.cellCount[cell], cnt);
            if (nClone) atomicAdd(&a.cellCount[cell], nClone);  // clones take slots of the same cell
        }
        // ---- the others wait in the queue for a full warp ----
        const bool cont = active && !done;
        const unsigned cm = __ballot_sync(0xffffffffu, cont);
        if (cont) {
            const int slot = qn + __popc(cm & ((1u << lane) - 1u));
            __stcg(&qD[slot], x0); __stcg(&qD[MQ_CAP + slot], x1);
            if (FLAT) { __stcg(&qD[2 * MQ_CAP + slot], U0); __stcg(&qD[3 * MQ_CAP + slot], U1); __stcg(&qD[4 * MQ_CAP + slot], sf); }
            else {
                __stcg(&qD[2 * MQ_CAP + slot], x2); __stcg(&qD[3 * MQ_CAP + slot], U0); __stcg(&qD[4 * MQ_CAP + slot], U1);
                __stcg(&qD[5 * MQ_CAP + slot], U2); __stcg(&qD[6 * MQ_CAP + slot], sf);
            }
            __stcg(&qI[slot], cell); __stcg(&qI[MQ_CAP + slot], i); __stcg(&qI[2 * MQ_CAP + slot], misc);
        }
        qn += __popc(cm);
        __syncwarp();
    }
}

}  // namespace ugf
