// ugf_oracle.cpp — CPU restatement of uniGasFoam's per-timestep particle loop.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (unigasfoam_b200/, libugf.so)
// may include, link or call this file.  It is used by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs, as the checker and as the timed
// CPU baseline.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
// path, and cannot be built here (needs OpenFOAM >= v2212 + MPI).  The
// restatement is pinned instead by closed-form kinetic-theory answers (tests/) and by
// Philox known-answer vectors for the RNG.
//
// What is restated (paths relative to the uniGasFoam root; U/ = src/lagrangian/uniGas/,
// CWM/ = src/lagrangian/CloudWithModels/):
//   evolve() phase order                U/clouds/uniGasCloud.C:821-869
//   parcel move loop                    U/parcels/uniGasParcel.C:35-108
//   wall models / measurements          U/boundaries/basic/uniGasPatchBoundary/uniGasPatchBoundary.C:130-403
//   free-stream inflow                  U/boundaries/basic/uniGasGeneralBoundary/uniGasGeneralBoundary.C:115-169,537-761
//   cell occupancy ("the sort")         CWM/CloudWithModels/CloudWithModels.C:110-138
//   cell moments                        U/cellMeasurements/cellMeasurements.C:408-513
//   NTC partner selection               U/dsmcCollisionPartner/derived/noTimeCounter/noTimeCounter.C:66-343
//   VHS / VSS / Larsen-Borgnakke        U/dsmcCollisions/derived/*/ *.C
//   kinetic samplers                    U/clouds/uniGasCloud.C:927-1326
//   BGK / ES-BGK / S-BGK / USP-SBGK     U/bgkCollisions/derived/*/ *.C
//   volume + wall fields                U/macroscopicProperties/derived/volumetric/uniGasVolFields/uniGasVolFields.C:723-1352
//
// What is NOT in the reference tree and is restated from OpenFOAM's documented
// behaviour: Cloud::move / particle::trackToAndHitFace are
// replaced by a face-plane walker (identical on planar-faced convex cells), and
// Foam::Random by counter-based Philox4x32-10 streams (the reference seeds from the wall
// clock, so no stream is reproducible upstream anyway).  The stream layout is part of
// the contract with the CUDA path and is documented in DESIGN.md §RNG.
//
// Layout follows the reference's style on purpose (array-of-structs parcels, one pass
// per phase, per-cell index lists): this is also the timed CPU baseline.

#include "../include/ugf.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// OpenFOAM DimensionedConstants defaults (SURVEY §8c "constants").
constexpr double kB = 1.38065e-23;         // physicoChemical::k
constexpr double NA = 6.02214e+23;         // physicoChemical::NA
constexpr double PI = 3.14159265358979323846;
constexpr double TWO_PI = 6.28318530717958647692;
constexpr double GREAT = 1e15;
constexpr double VSMALL = 1e-300;
constexpr double SMALL = 1e-15;

// ---------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), restated from the published algorithm.
// ---------------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { KIND_MOVE = 1, KIND_NTC = 2, KIND_BGK = 3, KIND_INFLOW = 4, KIND_WEIGHT = 5 };

// One stream = (seed, kind, aux) key + (a, b, c) counter prefix; the 4th counter word
// counts Philox blocks.  Each block yields two uniforms in [0,1) with 53 random bits.
struct Stream {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t out[4];
    int have;
    Stream(uint64_t seed, uint32_t kind, uint32_t aux, uint32_t a, uint32_t b, uint32_t c) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32) ^ (kind << 24) ^ aux;
        ctr[0] = a; ctr[1] = b; ctr[2] = c; ctr[3] = 0;
        have = 0;
    }
    inline double u01() {  // Random::sample01<scalar>()
        if (!have) { philox4x32_10(ctr, key, out); ctr[3]++; have = 2; }
        const uint32_t hi = out[(2 - have) * 2], lo = out[(2 - have) * 2 + 1];
        have--;
        return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
    }
    // Two independent N(0,1) from two uniforms (Box-Muller).  Foam::Random::GaussNormal is
    // a cached polar method; the distribution is the same, the draw count is fixed here.
    inline void gauss2(double& g1, double& g2) {
        const double u1 = u01(), u2 = u01();
        const double r = std::sqrt(-2.0 * std::log(1.0 - u1));
        const double th = TWO_PI * u2;
        g1 = r * std::cos(th);
        g2 = r * std::sin(th);
    }
    inline void gauss3(double g[3]) {
        double d;
        gauss2(g[0], g[1]);
        gauss2(g[2], d);
    }
    // Random::position<label>(0, n-1)
    inline int position(int n) {
        int i = (int)(u01() * n);
        return i < n - 1 ? i : n - 1;
    }
};

struct Parcel {  // uniGasParcel (U/parcels/uniGasParcel.H:217-239) + particle position/cell
    double x[3];
    double U[3];
    double ERot;
    double CWF;      // cell weight factor carried by the parcel (U/parcels/uniGasParcel.H:226)
    double RWF = 1.0;// radial weight factor carried by the parcel (uniGasParcel.H:229); 1 unless axisymmetricSimulation
    double sf;       // stepFraction
    int32_t cell;    // >=0 live; -1 deleted; <= -2 waiting on a processor face (-2 - boundaryFaceIndex)
    int32_t typeId;
    int32_t newParcel;
    int32_t ELevel = 0;                                   // electronic level (U/parcels/uniGasParcel.H:232)
    int32_t vib[UGF_MAX_VIB_MODES] = {0, 0, 0, 0};        // vibrational quantum level per mode (uniGasParcel.H:238)
};

inline double vibEnergy(const ugf_species& s, const Parcel& p) {  // sum over modes of level * k * thetaV
    double e = 0;
    for (int m = 0; m < s.vibrationalDoF; ++m) e += p.vib[m] * s.thetaV[m] * kB;
    return e;
}
inline double elecEnergy(const ugf_species& s, const Parcel& p) { return s.electronicEnergy[p.ELevel]; }

struct WallModel {
    int model = UGF_WALL_UNSET;
    double T = 0, Uw[3] = {0, 0, 0}, diffuseFraction = 1, alphaN = 1, sigmaT = 1, alphaR = 1;
    std::vector<double> faceT, faceU;  // *FieldPatch variants: boundaryT / boundaryU per face of the patch
};

struct InflowPatch {
    int patch; ugf_inflow in;
    // uniGasLiouFangPressureInletPatch: mole fractions, relaxation factor, inflow velocity per face
    bool pressure = false;
    double molFrac[UGF_MAX_SPECIES] = {1, 1, 1, 1, 1, 1, 1, 1};
    double theta = 1.0;
    std::vector<double> faceVel;
    // uniGasFreeStreamInflowFieldPatch (…/uniGasFreeStreamInflowFieldPatch.C:50-228): values per face of the patch;
    // faceVel holds boundaryU then.  faceN [nTypeIds][nFaces], faceTtr / faceTrot [nFaces]
    bool fields = false;
    std::vector<double> faceN, faceTtr, faceTrot;
    // uniGasWangPressureInletPatch (…/uniGasWangPressureInletPatch.C:54-281): running sums per face (0 parcels, 1 mass, 2-4
    // momentum, 5-7 sum U^2, 8-10 sum U), step count, inlet pressure, mixture molecular mass, gamma * R
    bool wang = false;
    // uniGasLiouFangPressureOutletPatch (…/uniGasLiouFangPressureOutletPatch.C:50-322): the same sums; faceN [nFaces*nTypeIds]
    // (face-major here), faceTtr = faceTrot and faceVel follow the flow; capN / capT: the device library's insertion bound
    bool outlet = false;
    double capN = 0.0, capT = 0.0;
    std::vector<double> wangSums;
    double wangSteps = 0.0, wangP = 0.0, wangM = 0.0, wangGammaR = 0.0;
    bool ce = false;               // uniGasChapmanEnskogFreeStreamInflowPatch
    // uniGasMassFlowRateInletPatch (…/uniGasMassFlowRateInletPatch.C:53-302): faceN [nFaces*nTypeIds] and faceVel follow the flow so
    // that the patch lets in massFlowRate; outFlux [nFaces*nSpecies] = this step's parcelIdFlux of the face tracker on the patch faces
    bool massFlow = false;
    double massFlowRate = 0.0, patchArea = 0.0;
    std::vector<double> outFlux, mfMolFrac;
    double ceQ[3] = {0, 0, 0}, ceS[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};
constexpr int WANG_NSUM = 11;

constexpr int NACC = 16;

}  // namespace

struct ugfo_handle {
    ugf_config cfg;
    std::string err;
    int nSpecies = 0;
    ugf_species sp[UGF_MAX_SPECIES];

    // mesh
    int nCells = 0, nFaces = 0, nInternal = 0, nPatches = 0, nBFaces = 0;
    std::vector<int32_t> owner, neighbour, cfOff, cf, pStart, pSize, pKind, pPartner;
    std::vector<double> Sf, Cf, vol, cc, bbMin, bbMax, pSep, points;
    std::vector<int32_t> fpOff, fp;
    std::vector<int32_t> facePatch;  // [nBFaces]
    std::vector<WallModel> wall;     // [nPatches]
    std::vector<InflowPatch> inflows;

    // cloud
    std::vector<Parcel> P, Pswap;
    int64_t nBeforeInsert = 0;
    int64_t receivedStart = -1;
    std::vector<int32_t> occOff, occIds;  // cell occupancy CSR
    bool occValid = false, occIdentity = false;
    bool outletBoundHit = false;  // a pressure-outlet face asked for more parcels than the device library's bound
    bool weightPending = false;   // a move happened since the last weighting() pass
    std::vector<int32_t> faceTrack;   // [nFaces] k + 1 of a tracked face, 0 otherwise (uniGasFaceTracker)
    std::vector<int> patchFlux;       // [nPatches] index in inflows of the mass-flow inlet on that patch, or -1; empty without one
    std::vector<double> ft;           // [nTracked][nSpecies][UGF_NFT]
    int nTracked = 0;

    // cell state (U/clouds/uniGasCloud.H:189-201)
    std::vector<double> sigmaTcRMax;
    std::vector<int32_t> collModelId, subLevels;
    std::vector<double> cellWF;   // cellWeightFactor_ (U/clouds/uniGasCloud.C:433); all 1 until uploaded
    bool cellWeighted = false;    // cellWeightedSimulation: a cellWeightFactor field was uploaded
    // BGK persistent state
    std::vector<double> maxProb, qPrev, sPrev;
    // per-step measurements
    std::vector<double> mom;   // [nCells][nSpecies][UGF_NMOM]
    bool momValid = false;
    std::vector<double> bm;    // [nBFaces][UGF_NBM]
    // time-averaged accumulators (uniGasVolFields)
    std::vector<double> acc;   // [nCells][NACC]
    std::vector<double> accS;  // [nCells][nSpecies] nParcelsXnParticle per species (mean free path fields)
    std::vector<double> bacc;  // [nBFaces][UGF_NBM]
    std::vector<double> momE;  // [nCells][nSpecies][2] parcels in the ground / first electronic level at the last sample (species with > 1 level)
    std::vector<double> accI;  // [nCells][nSpecies][UGF_NINT] internal-mode accumulators, only when a species has vibrational modes or > 1 electronic level
    bool internalModes = false;
    double timeAvCounter = 0;
    int64_t nAvTimeSteps = 0;
    int sampleCounter = 0;

    // localKnudsen hybrid decomposition (U/hybridDecomposition)
    bool decompOn = false;
    ugf_decomposition dec{};
    int decTimeSteps = 0;
    double decTimeAv = 0;
    std::vector<double> knAcc;          // [nCells][KN_NACC + nSpecies]
    std::vector<double> knFields;       // [nCells][4] KnRho, KnT, KnU, KnGLL
    std::vector<double> faceW, faceMagSf;   // linear-interpolation weight of the owner side, |Sf|
    std::vector<int32_t> ccOff, ccIds;  // mesh.cellCells(): neighbours across internal faces

    int64_t step = 0;
    bool stepOpen = false;
    ugf_counters cnt;
    std::vector<std::vector<double>> packBuf;
    int64_t inflight = 0;
    // interpolationCellPoint geometry (collisionProperties.macroInterpolation), see ugf_cell_point
    bool cpSet = false, relaxFailed = false;
    std::vector<double> cpPoints, cpW, cpNormals;
    std::vector<int32_t> cpTetOff, cpTetPts, cpPcOff, cpPc;
    std::vector<int64_t> migIdx;  // indices of the parcels waiting on processor patches (ascending), filled by moveRange
};

namespace {

inline int fail(ugfo_handle* h, const std::string& m) { if (h) h->err = m; return 1; }

// nParticle * CWF of a cell; the radial factor comes on top where the reference applies it
inline double FNc(const ugfo_handle& h, int c) { return h.cfg.nParticle * h.cellWF[c]; }
// uniGasCloud::axiRWF (U/clouds/uniGasCloudI.H:116-120); 1 without axisymmetricSimulation
inline double axiRWF(const ugfo_handle& h, const double* x) {
    if (!h.cfg.axisymmetric) return 1.0;
    const double radius = std::sqrt(x[1] * x[1] + x[2] * x[2]);
    return 1.0 + (h.cfg.maxRWF - 1.0) * radius / h.cfg.radialExtent;
}

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ---------------------------------------------------------------------------------
// kinetic samplers  (U/clouds/uniGasCloud.C:975-1017, 1129-1189, 1267-1326)
// ---------------------------------------------------------------------------------
double equipartitionRotationalEnergy(Stream& r, double T, int rotDoF) {
    if (rotDoF < 1) return 0.0;
    if (rotDoF == 2) {
        // reference: -log(sample01)*k*T; 1-u is used so that u = 0 cannot give inf.
        return -std::log(1.0 - r.u01()) * kB * T;
    }
    const double a = 0.5 * rotDoF - 1;
    double energyRatio, Pp;
    const double eps = r.u01();
    do {
        energyRatio = 10 * r.u01();
        Pp = std::pow(energyRatio / a, a) * std::exp(a - energyRatio);
    } while (Pp < eps);
    return energyRatio * kB * T;
}

// equipartitionVibrationalEnergyLevel (U/clouds/uniGasCloud.C:1020-1050): level = int(-ln(R) T / thetaV) per mode; 1 - u is
// used like for the rotational energy so that u = 0 cannot give inf (the reference redraws a zero instead)
void equipartitionVibrationalEnergyLevel(Stream& r, double T, const ugf_species& s, int32_t* vib) {
    for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) vib[m] = 0;
    for (int m = 0; m < s.vibrationalDoF; ++m) vib[m] = (int32_t)(-std::log(1.0 - r.u01()) * T / s.thetaV[m]);
}

// equipartitionElectronicLevel (U/clouds/uniGasCloud.C:1053-1126): acceptance-rejection against the most populated level
// (Liechty eqs 3.1.1 / 3.1.2); the threshold is drawn once per call, as the reference does
int equipartitionElectronicLevel(Stream& r, double T, const ugf_species& s) {
    const int jMax = s.nElectronicLevels;
    if (jMax == 1) return 0;
    if (T < VSMALL) return 0;
    const double EMax = kB * T;
    double expSum = 0.0;
    for (int i = 0; i < jMax; ++i) expSum += s.degeneracy[i] * std::exp(-s.electronicEnergy[i] / EMax);
    double boltzMax = 0.0;
    int jSelect = 0;
    for (int i = 0; i < jMax; ++i) {
        const double boltz = s.degeneracy[i] * std::exp(-s.electronicEnergy[i] / EMax) / expSum;
        if (boltzMax < boltz) { boltzMax = boltz; jSelect = i; }
    }
    const double expMax = s.degeneracy[jSelect] * std::exp(-s.electronicEnergy[jSelect] / EMax);
    const double eps = r.u01();
    int jDash;
    double func;
    do {
        jDash = r.position(jMax);
        func = s.degeneracy[jDash] * std::exp(-s.electronicEnergy[jDash] / EMax) / expMax;
    } while (!(func > eps));
    return jDash;
}

// postCollisionVibrationalEnergyLevel, postReaction = false (U/clouds/uniGasCloud.C:1192-1264): quantum-kinetic exchange with
// the collision-temperature dependent vibrational collision number (Bird 2010 eqs 2, 3; Bird 5.42, 5.61)
int postCollisionVibrationalEnergyLevel(Stream& r, int vibLevel, int iMax, double thetaV, double thetaD, double refTempZv, double omega,
                                        double Zref, double Ec) {
    int iDash = vibLevel;
    const double TColl = (iMax * thetaV) / (3.5 - omega);
    const double pow1 = std::pow(thetaD / TColl, 0.33333) - 1.0;
    const double pow2 = std::pow(thetaD / refTempZv, 0.33333) - 1.0;
    const double ZvP1 = std::pow(thetaD / TColl, omega);
    const double ZvP2 = std::pow(Zref * std::pow(thetaD / refTempZv, -omega), pow1 / pow2);
    const double Zv = ZvP1 * ZvP2;
    const double inverseVibrationalCollisionNumber = 1.0 / (5.0 * Zv);
    if (inverseVibrationalCollisionNumber > r.u01()) {
        double func, EVib;
        do {
            const int i = (int)(r.u01() * (iMax + 1));  // Random::position<label>(0, iMax)
            iDash = i < iMax ? i : iMax;
            EVib = iDash * kB * thetaV;
            func = std::pow(1.0 - EVib / Ec, 1.5 - omega);
        } while (!(func > r.u01()));
    }
    return iDash;
}

double postCollisionRotationalEnergy(Stream& r, int rotDoF, double ChiB) {
    double energyRatio = 0.0;
    if (rotDoF == 2) {
        energyRatio = 1.0 - std::pow(r.u01(), 1.0 / ChiB);
    } else {
        const double ChiA = 0.5 * rotDoF;
        const double A1 = ChiA - 1, B1 = ChiB - 1;
        if (A1 < SMALL && B1 < SMALL) return r.u01();
        double Pp;
        const double eps = r.u01();
        do {
            energyRatio = r.u01();
            if (A1 < SMALL) Pp = std::pow(1.0 - energyRatio, B1);
            else if (B1 < SMALL) Pp = std::pow(1.0 - energyRatio, A1);
            else Pp = std::pow((A1 + B1) * energyRatio / A1, A1) * std::pow((A1 + B1) * (1 - energyRatio) / B1, B1);
        } while (Pp < eps);
    }
    return energyRatio;
}

int postCollisionElectronicEnergyLevel(Stream& r, double Ec, int jMax, double omega, const double* EE, const int32_t* g) {
    int nPossStates = 0;
    if (jMax == 1) nPossStates = g[0];
    else for (int n = 0; n < jMax; ++n) if (Ec > EE[n]) nPossStates += g[n];
    int ELevel = -1;
    for (;;) {
        const int nState = (int)std::ceil(r.u01() * nPossStates);
        int nAvail = 0, nLevel = -1;
        for (int n = 0; n < jMax; ++n) {
            nAvail += g[n];
            if (nState <= nAvail && nLevel < 0) nLevel = n;
        }
        if (nLevel < 0) nLevel = 0;
        if (Ec > EE[nLevel]) {
            const double prob = std::pow(1.0 - EE[nLevel] / Ec, 1.5 - omega);
            if (prob > r.u01()) { ELevel = nLevel; break; }
        }
    }
    return ELevel;
}

// ---------------------------------------------------------------------------------
// binary collision models
// ---------------------------------------------------------------------------------
// variableHardSphere::sigmaTcR (…/variableHardSphere/variableHardSphere.C:72-115); the VSS and
// both LB models code the same expression.
double sigmaTcR(const ugfo_handle& h, const Parcel& p, const Parcel& q) {
    const double d0 = p.U[0] - q.U[0], d1 = p.U[1] - q.U[1], d2 = p.U[2] - q.U[2];
    const double cR2 = d0 * d0 + d1 * d1 + d2 * d2;
    if (cR2 < VSMALL) return 0;
    const ugf_species& a = h.sp[p.typeId];
    const ugf_species& b = h.sp[q.typeId];
    const double dPQ = 0.5 * (a.d + b.d);
    const double omegaPQ = 0.5 * (a.omega + b.omega);
    const double mR = a.mass * b.mass / (a.mass + b.mass);
    const double sigmaTPQ = PI * dPQ * dPQ * std::pow(2.0 * kB * h.cfg.Tref / (mR * cR2), omegaPQ - 0.5)
                            / std::exp(std::lgamma(2.5 - omegaPQ));
    return sigmaTPQ * std::sqrt(cR2);
}

// isotropic (VHS) scattering of a relative velocity of magnitude cR (variableHardSphere.C:137-160)
inline void scatterVHS(Stream& r, double cR, double out[3]) {
    const double cosTheta = 2.0 * r.u01() - 1.0;
    const double sinTheta = std::sqrt(1.0 - cosTheta * cosTheta);
    const double phi = TWO_PI * r.u01();
    out[0] = cR * cosTheta;
    out[1] = cR * (sinTheta * std::cos(phi));
    out[2] = cR * (sinTheta * std::sin(phi));
}

// VSS scattering, Bird eq 2.22 (variableSoftSphere.C:147-170).  cRc = pre-collision relative
// velocity components; scale = |c_r'|/|c_r| (1 for plain VSS, sqrt(Etr'/Etr) after LB exchange).
inline void scatterVSS(Stream& r, const double cRc[3], double alphaPQ, double scale, double out[3]) {
    const double cR = std::sqrt(cRc[0] * cRc[0] + cRc[1] * cRc[1] + cRc[2] * cRc[2]);
    const double cosTheta = 2.0 * std::pow(r.u01(), 1.0 / alphaPQ) - 1.0;
    const double sinTheta = std::sqrt(1.0 - cosTheta * cosTheta);
    const double phi = TWO_PI * r.u01();
    const double D = std::sqrt(cRc[1] * cRc[1] + cRc[2] * cRc[2]);
    const double sp = std::sin(phi), cp = std::cos(phi);
    out[0] = scale * (cosTheta * cRc[0] + sinTheta * sp * D);
    out[1] = scale * (cosTheta * cRc[1] + sinTheta * (cR * cRc[2] * cp - cRc[0] * cRc[1] * sp) / D);
    out[2] = scale * (cosTheta * cRc[2] - sinTheta * (cR * cRc[1] * cp + cRc[0] * cRc[2] * sp) / D);
}

void collidePair(const ugfo_handle& h, Stream& r, Parcel& p, Parcel& q) {
    const ugf_species& a = h.sp[p.typeId];
    const ugf_species& b = h.sp[q.typeId];
    const double mP = a.mass, mQ = b.mass, mS = mP + mQ;
    double Ucm[3], cRc[3];
    for (int k = 0; k < 3; ++k) {
        Ucm[k] = (mP * p.U[k] + mQ * q.U[k]) / mS;
        cRc[k] = p.U[k] - q.U[k];
    }
    const double cRsqr = cRc[0] * cRc[0] + cRc[1] * cRc[1] + cRc[2] * cRc[2];
    double rel[3];
    const int model = h.cfg.binaryModel;
    if (model == UGF_BINARY_VHS) {
        scatterVHS(r, std::sqrt(cRsqr), rel);
    } else if (model == UGF_BINARY_VSS) {
        scatterVSS(r, cRc, 0.5 * (a.alpha + b.alpha), 1.0, rel);
    } else {
        // Larsen-Borgnakke, serial application, P then Q
        // (…/LarsenBorgnakkeVariableHardSphere.C:125-416)
        const double omegaPQ = 0.5 * (a.omega + b.omega);
        const double mR = mP * mQ / mS;
        double Etr = 0.5 * mR * cRsqr;
        const double ChiB = 2.5 - omegaPQ;
        const double invZrot = 1.0 / h.cfg.rotationalRelaxationCollisionNumber;
        const double invZel = 1.0 / h.cfg.electronicRelaxationCollisionNumber;
        const double preERotP = p.ERot, preERotQ = q.ERot;
        double preEVibP[UGF_MAX_VIB_MODES], preEVibQ[UGF_MAX_VIB_MODES];
        for (int m = 0; m < a.vibrationalDoF; ++m) preEVibP[m] = p.vib[m] * a.thetaV[m] * kB;
        for (int m = 0; m < b.vibrationalDoF; ++m) preEVibQ[m] = q.vib[m] * b.thetaV[m] * kB;
        const double preEEleP = a.electronicEnergy[p.ELevel], preEEleQ = b.electronicEnergy[q.ELevel];
        // P: electronic, vibrational (quantum-kinetic, mode by mode), rotational
        if (invZel > r.u01()) {
            const double Ec = Etr + preEEleP;
            p.ELevel = postCollisionElectronicEnergyLevel(r, Ec, a.nElectronicLevels, omegaPQ, a.electronicEnergy, a.degeneracy);
            Etr = Ec - a.electronicEnergy[p.ELevel];
        }
        for (int m = 0; m < a.vibrationalDoF; ++m) {  // LarsenBorgnakkeVariableHardSphere.C:261-300
            const double Ec = Etr + preEVibP[m];
            const int iMax = (int)(Ec / (kB * a.thetaV[m]));
            if (iMax > 0) {
                p.vib[m] = postCollisionVibrationalEnergyLevel(r, p.vib[m], iMax, a.thetaV[m], a.thetaD[m], a.TrefZv[m], omegaPQ, a.Zref[m], Ec);
                Etr = Ec - p.vib[m] * a.thetaV[m] * kB;
            }
        }
        if (a.rotationalDoF > 0) {
            if (invZrot > r.u01()) {
                const double Ec = Etr + preERotP;
                const double ratio = postCollisionRotationalEnergy(r, a.rotationalDoF, ChiB);
                p.ERot = ratio * Ec;
                Etr = Ec - p.ERot;
            }
        }
        if (invZel > r.u01()) {
            const double Ec = Etr + preEEleQ;
            q.ELevel = postCollisionElectronicEnergyLevel(r, Ec, b.nElectronicLevels, omegaPQ, b.electronicEnergy, b.degeneracy);
            Etr = Ec - b.electronicEnergy[q.ELevel];
        }
        for (int m = 0; m < b.vibrationalDoF; ++m) {  // :337-376
            const double Ec = Etr + preEVibQ[m];
            const int iMax = (int)(Ec / (kB * b.thetaV[m]));
            if (iMax > 0) {
                q.vib[m] = postCollisionVibrationalEnergyLevel(r, q.vib[m], iMax, b.thetaV[m], b.thetaD[m], b.TrefZv[m], omegaPQ, b.Zref[m], Ec);
                Etr = Ec - q.vib[m] * b.thetaV[m] * kB;
            }
        }
        if (b.rotationalDoF > 0) {
            if (invZrot > r.u01()) {
                const double Ec = Etr + preERotQ;
                const double ratio = postCollisionRotationalEnergy(r, b.rotationalDoF, ChiB);
                q.ERot = ratio * Ec;
                Etr = Ec - q.ERot;
            }
        }
        const double cRnew = std::sqrt((2.0 * Etr) / mR);
        if (model == UGF_BINARY_LB_VHS) scatterVHS(r, cRnew, rel);
        else scatterVSS(r, cRc, 0.5 * (a.alpha + b.alpha), cRnew / std::sqrt(cRsqr), rel);
    }
    for (int k = 0; k < 3; ++k) {
        p.U[k] = Ucm[k] + rel[k] * mQ / mS;
        q.U[k] = Ucm[k] - rel[k] * mP / mS;
    }
}

// ---------------------------------------------------------------------------------
// wall interaction  (uniGasPatchBoundary.C:130-403)
// ---------------------------------------------------------------------------------
inline void unitNormal(const double* S, double nw[3], double& fA) {
    fA = std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
    nw[0] = S[0] / fA; nw[1] = S[1] / fA; nw[2] = S[2] / fA;
}

void measureWall(ugfo_handle& h, const Parcel& p, int bfi, const double nw[3], double fA, double* preIE, double* preIMom, bool after) {
    const ugf_species& s = h.sp[p.typeId];
    const double m = s.mass;
    const double Un = dot3(p.U, nw);
    const double inv = 1.0 / std::max(std::fabs(Un) * fA, VSMALL);
    const double UU = dot3(p.U, p.U);
    double* b = &h.bm[(size_t)bfi * UGF_NBM];
    const double add[UGF_NBM] = {
        inv, m * inv, 0.5 * m * UU * inv, m * p.U[0] * inv, m * p.U[1] * inv, m * p.U[2] * inv,
        p.ERot * inv, s.rotationalDoF * inv, 0, 0, 0, 0, (s.rotationalDoF > 0 ? inv : 0.0), vibEnergy(s, p) * inv, elecEnergy(s, p) * inv,
        after ? 0.0 : 1.0};
    const double IE = 0.5 * m * UU + p.ERot + elecEnergy(s, p) + vibEnergy(s, p);
    double dq = 0, dfd[3] = {0, 0, 0};
    if (!after) {
        *preIE = IE;
        for (int k = 0; k < 3; ++k) preIMom[k] = m * p.U[k];
    } else {
        const double nPart = FNc(h, p.cell) * axiRWF(h, p.x);  // nParticle*CWF of the wall cell*RWF(hit position) (uniGasPatchBoundary.C:292-294)
        dq = nPart * (*preIE - IE) / (h.cfg.deltaT * fA);
        for (int k = 0; k < 3; ++k) dfd[k] = nPart * (preIMom[k] - m * p.U[k]) / (h.cfg.deltaT * fA);
    }
    for (int k = 0; k < UGF_NBM; ++k) {
        double v = add[k];
        if (k == 8) v = dq;
        if (k >= 9 && k <= 11) v = dfd[k - 9];
        if (v != 0.0) {
#pragma omp atomic
            b[k] += v;
        }
    }
}

void diffuseReflection(const ugfo_handle& h, Stream& r, Parcel& p, const double nw[3], double T, const double* Uw) {
    double Un = dot3(p.U, nw);
    double Ut[3] = {p.U[0] - Un * nw[0], p.U[1] - Un * nw[1], p.U[2] - Un * nw[2]};
    double magUt = std::sqrt(dot3(Ut, Ut));
    while (magUt < SMALL) {
        p.U[0] = p.U[0] * (0.8 + 0.2 * r.u01());
        p.U[1] = p.U[1] * (0.8 + 0.2 * r.u01());
        p.U[2] = p.U[2] * (0.8 + 0.2 * r.u01());
        Un = dot3(p.U, nw);
        for (int k = 0; k < 3; ++k) Ut[k] = p.U[k] - Un * nw[k];
        magUt = std::sqrt(dot3(Ut, Ut));
        if (dot3(p.U, p.U) == 0.0) {  // reference would spin forever on U == 0; pick a tangent
            const int kmin = std::fabs(nw[0]) <= std::fabs(nw[1]) ? (std::fabs(nw[0]) <= std::fabs(nw[2]) ? 0 : 2)
                                                                 : (std::fabs(nw[1]) <= std::fabs(nw[2]) ? 1 : 2);
            double e[3] = {0, 0, 0};
            e[kmin] = 1.0;
            const double en = dot3(e, nw);
            for (int k = 0; k < 3; ++k) Ut[k] = e[k] - en * nw[k];
            magUt = std::sqrt(dot3(Ut, Ut));
        }
    }
    const double tw1[3] = {Ut[0] / magUt, Ut[1] / magUt, Ut[2] / magUt};
    const double tw2[3] = {nw[1] * tw1[2] - nw[2] * tw1[1], nw[2] * tw1[0] - nw[0] * tw1[2], nw[0] * tw1[1] - nw[1] * tw1[0]};
    const ugf_species& s = h.sp[p.typeId];
    double g1, g2;
    r.gauss2(g1, g2);
    const double gn = std::sqrt(-2.0 * std::log(std::max(1 - r.u01(), VSMALL)));
    const double c = std::sqrt(kB * T / s.mass);
    for (int k = 0; k < 3; ++k) p.U[k] = c * (g1 * tw1[k] + g2 * tw2[k] - gn * nw[k]);
    p.ERot = equipartitionRotationalEnergy(r, T, s.rotationalDoF);
    if (s.vibrationalDoF > 0) equipartitionVibrationalEnergyLevel(r, T, s, p.vib);        // uniGasPatchBoundary.C:373-374
    if (s.nElectronicLevels > 1) p.ELevel = equipartitionElectronicLevel(r, T, s);         // :376-383
    for (int k = 0; k < 3; ++k) p.U[k] += Uw[k];
}

// uniGasCLLWallPatch::controlParticle (uniGasCLLWallPatch.C:80-254): Cercignani-Lampis-Lord kernel with normal /
// tangential / rotational accommodation.  Draw order as in the reference; -log(1-u) for its -log(u);
// rotDoF 3 uses sqrt(alphaR) where the reference writes sqrt(-alphaR) (NaN for any positive coefficient).
void cllReflection(const ugfo_handle& h, Stream& r, Parcel& p, const double nw[3], const WallModel& w) {
    bool degenerate = false;  // U == 0: no incident tangential speed, tangent chosen arbitrarily
    double Un = dot3(p.U, nw);
    double Ut[3] = {p.U[0] - Un * nw[0], p.U[1] - Un * nw[1], p.U[2] - Un * nw[2]};
    double magUt = std::sqrt(dot3(Ut, Ut));
    while (magUt < SMALL) {
        p.U[0] = p.U[0] * (0.8 + 0.2 * r.u01());
        p.U[1] = p.U[1] * (0.8 + 0.2 * r.u01());
        p.U[2] = p.U[2] * (0.8 + 0.2 * r.u01());
        Un = dot3(p.U, nw);
        for (int k = 0; k < 3; ++k) Ut[k] = p.U[k] - Un * nw[k];
        magUt = std::sqrt(dot3(Ut, Ut));
        if (dot3(p.U, p.U) == 0.0) {  // reference would spin forever on U == 0; pick a tangent
            const int kmin = std::fabs(nw[0]) <= std::fabs(nw[1]) ? (std::fabs(nw[0]) <= std::fabs(nw[2]) ? 0 : 2)
                                                                 : (std::fabs(nw[1]) <= std::fabs(nw[2]) ? 1 : 2);
            double e[3] = {0, 0, 0};
            e[kmin] = 1.0;
            const double en = dot3(e, nw);
            for (int k = 0; k < 3; ++k) Ut[k] = e[k] - en * nw[k];
            magUt = std::sqrt(dot3(Ut, Ut));
            degenerate = true;
            break;
        }
    }
    const double tw1[3] = {Ut[0] / magUt, Ut[1] / magUt, Ut[2] / magUt};
    const double tw2[3] = {nw[1] * tw1[2] - nw[2] * tw1[1], nw[2] * tw1[0] - nw[0] * tw1[2], nw[0] * tw1[1] - nw[1] * tw1[0]};
    const ugf_species& s = h.sp[p.typeId];
    const double T = w.T;
    const double alphaT = w.sigmaT * (2.0 - w.sigmaT), alphaN = w.alphaN, alphaR = w.alphaR;
    const double cmp = std::sqrt(2.0 * kB * T / s.mass);
    const double utN = degenerate ? 0.0 : magUt / cmp;
    const double unN = Un / cmp;
    const double thetaNormal = TWO_PI * r.u01();
    const double rNormal = std::sqrt(-alphaN * std::log(std::max(1 - r.u01(), VSMALL)));
    const double thetaTangential = TWO_PI * r.u01();
    const double rTangential = std::sqrt(-alphaT * std::log(std::max(1 - r.u01(), VSMALL)));
    const double um = std::sqrt(1.0 - alphaN) * unN;
    const double vN = std::sqrt(rNormal * rNormal + um * um + 2.0 * rNormal * um * std::cos(thetaNormal));
    const double vT1 = std::sqrt(1.0 - alphaT) * utN + rTangential * std::cos(thetaTangential);
    const double vT2 = rTangential * std::sin(thetaTangential);
    double U[3];
    for (int k = 0; k < 3; ++k) U[k] = cmp * (vT1 * tw1[k] + vT2 * tw2[k] - vN * nw[k]);
    const double wN = dot3(w.Uw, nw), w1 = dot3(w.Uw, tw1), w2 = dot3(w.Uw, tw2);
    const double uN = dot3(U, nw), u1 = dot3(U, tw1), u2 = dot3(U, tw2);
    for (int k = 0; k < 3; ++k)
        p.U[k] = (uN * nw[k] + wN * nw[k] * alphaN) + (u1 * tw1[k] + w1 * tw1[k] * alphaT) + (u2 * tw2[k] + w2 * tw2[k] * alphaT);
    if (s.rotationalDoF == 2) {
        const double om = std::sqrt(p.ERot * (1.0 - alphaR) / (kB * T));
        const double rRot = std::sqrt(-alphaR * std::log(std::max(1.0 - r.u01(), VSMALL)));
        const double thetaRot = TWO_PI * r.u01();
        p.ERot = kB * T * (rRot * rRot + om * om + 2.0 * rRot * om * std::cos(thetaRot));
    } else if (s.rotationalDoF == 3) {
        double X, A;
        do {
            X = 4.0 * r.u01();
            A = 2.7182818 * X * X * std::exp(-(X * X));
        } while (A < r.u01());
        const double om = std::sqrt(p.ERot * (1.0 - alphaR) / (kB * T));
        const double rRot = std::sqrt(alphaR) * X;
        const double thetaRot = 2.0 * r.u01() - 1.0;
        p.ERot = kB * T * (rRot * rRot + om * om + 2.0 * rRot * om * std::cos(thetaRot));
    }
}

// ---------------------------------------------------------------------------------
// move: face-plane walker standing in for particle::trackToAndHitFace
// ---------------------------------------------------------------------------------
constexpr int MAX_TRACK_ITERS = 4096;

struct MoveTally { int64_t deleted = 0, wallHits = 0, stuck = 0, migrated = 0; };

void moveParcel(ugfo_handle& h, Parcel& p, int64_t idx, MoveTally& t, bool freshStream) {
    Stream r(h.cfg.seed, KIND_MOVE, freshStream ? 0u : 1u, (uint32_t)h.step, (uint32_t)idx, 0);
    const double dt = h.cfg.deltaT;
    if (p.newParcel == 1) {  // U/parcels/uniGasParcel.C:47-53
        p.sf = r.u01();
        p.newParcel = 0;
    }
    int iters = 0;
    while (p.cell >= 0 && p.sf < 1) {
        const double rem = 1 - p.sf;
        const double s = rem * dt;
        double disp[3];
        for (int k = 0; k < 3; ++k) disp[k] = h.cfg.solutionD[k] ? s * p.U[k] : 0.0;  // constrainDirection
        // First face crossed: minimise lambda = num/nd over faces with nd > 0 (lambda clamped at 0).  Fractions are
        // compared by cross-multiplication, so there is one division per hop; (bnum, bnd) = (1, 1) encodes "end of step".
        double bnum = 1.0, bnd = 1.0;
        int hit = -1;
        bool hitFlip = false;
        const int c = p.cell;
        for (int j = h.cfOff[c]; j < h.cfOff[c + 1]; ++j) {
            const int f = h.cf[j];
            const bool own = (h.owner[f] == c);
            const double* S = &h.Sf[3 * (size_t)f];
            const double* C = &h.Cf[3 * (size_t)f];
            double nd = std::fma(S[2], disp[2], std::fma(S[1], disp[1], S[0] * disp[0]));
            double num = (S[0] * C[0] + S[1] * C[1] + S[2] * C[2]) - std::fma(S[2], p.x[2], std::fma(S[1], p.x[1], S[0] * p.x[0]));
            if (!own) { nd = -nd; num = -num; }
            if (num < 0) num = 0;
            if (nd > 0 && num * bnd < bnum * nd) { bnum = num; bnd = nd; hit = f; hitFlip = !own; }
        }
        if (hit < 0) {
            for (int k = 0; k < 3; ++k) p.x[k] = p.x[k] + disp[k];
            p.sf = 1;
            break;
        }
        const double lamMin = bnum / bnd;
        for (int k = 0; k < 3; ++k) p.x[k] = std::fma(lamMin, disp[k], p.x[k]);
        p.sf = std::fma(rem, lamMin, p.sf);
        if (hit < h.nInternal) {
            p.cell = hitFlip ? h.owner[hit] : h.neighbour[hit];
        } else {
            const int bfi = hit - h.nInternal;
            const int patch = h.facePatch[bfi];
            const int kind = h.pKind[patch];
            const double* S = &h.Sf[3 * (size_t)hit];
            if (kind == UGF_PATCH_WALL) {
                WallModel wl = h.wall[patch];
                if (!h.wall[patch].faceT.empty()) {  // …WallFieldPatch.C:108-114: this face's boundaryT / boundaryU
                    const int lf = hit - h.pStart[patch];
                    wl.T = h.wall[patch].faceT[lf];
                    for (int k = 0; k < 3; ++k) wl.Uw[k] = h.wall[patch].faceU[3 * (size_t)lf + k];
                    wl.faceT.clear(); wl.faceU.clear();
                }
                const WallModel& w = wl;
                t.wallHits++;
                if (w.model == UGF_WALL_DELETION) {
                    p.cell = -1; t.deleted++;
                } else {
                    double nw[3], fA, preIE = 0, preIMom[3] = {0, 0, 0};
                    unitNormal(S, nw, fA);
                    if (h.cfg.measureWalls) measureWall(h, p, bfi, nw, fA, &preIE, preIMom, false);
                    bool diffuse = (w.model == UGF_WALL_DIFFUSE);
                    if (w.model == UGF_WALL_MIXED) diffuse = (w.diffuseFraction > r.u01());
                    if (w.model == UGF_WALL_CLL) {
                        cllReflection(h, r, p, nw, w);
                    } else if (diffuse) {
                        diffuseReflection(h, r, p, nw, w.T, w.Uw);
                    } else {
                        const double Un = dot3(p.U, nw);
                        if (Un > 0.0) for (int k = 0; k < 3; ++k) p.U[k] = p.U[k] - 2.0 * Un * nw[k];
                    }
                    if (h.cfg.measureWalls) measureWall(h, p, bfi, nw, fA, &preIE, preIMom, true);
                }
            } else if (kind == UGF_PATCH_SYMMETRY) {
                double nw[3], fA;
                unitNormal(S, nw, fA);
                const double Un = dot3(p.U, nw);
                for (int k = 0; k < 3; ++k) p.U[k] = p.U[k] - 2.0 * Un * nw[k];
            } else if (kind == UGF_PATCH_CYCLIC) {
                const int q = h.pPartner[patch];
                const int nf = h.pStart[q] + (hit - h.pStart[patch]);
                p.cell = h.owner[nf];
                for (int k = 0; k < 3; ++k) p.x[k] = p.x[k] + h.pSep[3 * patch + k];
            } else if (kind == UGF_PATCH_PROCESSOR) {
                for (int k = 0; k < 3; ++k) p.x[k] = p.x[k] + h.pSep[3 * patch + k];
                p.cell = -2 - bfi;
                t.migrated++;
            } else if (kind == UGF_PATCH_GENERIC) {
                if (!h.patchFlux.empty() && h.patchFlux[patch] >= 0) {  // uniGasFaceTracker::updateFields on the faces of a mass-flow inlet (:98-141)
                    double* fl = h.inflows[h.patchFlux[patch]].outFlux.data() + (size_t)(hit - h.pStart[patch]) * h.nSpecies + p.typeId;
                    const double add = (dot3(p.U, S) >= 0.0 ? 1.0 : -1.0) * (p.CWF * axiRWF(h, p.x));
#pragma omp atomic
                    *fl += add;
                }
                p.cell = -1; t.deleted++;
            } else {  // empty patch hit: mesh/solutionD mismatch
                p.cell = -1; t.stuck++;
            }
        }
        if (h.nTracked) {  // uniGasFaceTracker::updateFields (U/faceTracker/uniGasFaceTracker.C:90-152), after the patch interaction
            int face = hit;
            if (hit >= h.nInternal) {
                const int patch = h.facePatch[hit - h.nInternal];
                if (h.pKind[patch] == UGF_PATCH_CYCLIC) face = h.pStart[h.pPartner[patch]] + (hit - h.pStart[patch]);
            }
            const int k = h.faceTrack[face] - 1;
            if (k >= 0) {
                const double* S = &h.Sf[3 * (size_t)hit];
                const double sgn = (hit < h.nInternal) ? (hitFlip ? -1.0 : 1.0) : (dot3(p.U, S) >= 0.0 ? 1.0 : -1.0);
                const ugf_species& s = h.sp[p.typeId];
                const double w = p.CWF * axiRWF(h, p.x);  // uniGasFaceTracker.C:98-99
                const double e = 0.5 * s.mass * dot3(p.U, p.U) + p.ERot + elecEnergy(s, p) + vibEnergy(s, p);
                double* tt = &h.ft[((size_t)k * h.nSpecies + p.typeId) * UGF_NFT];
                const double add[UGF_NFT] = {sgn * w, sgn * s.mass * w, s.mass * p.U[0] * w, s.mass * p.U[1] * w, s.mass * p.U[2] * w, sgn * e * w};
                for (int q = 0; q < UGF_NFT; ++q) {
#pragma omp atomic
                    tt[q] += add[q];
                }
            }
        }
        if (++iters > MAX_TRACK_ITERS) { p.cell = -1; t.stuck++; break; }
    }
}

void moveRange(ugfo_handle& h, int64_t begin, int64_t end, bool fresh) {
    int64_t deleted = 0, wallHits = 0, stuck = 0, migrated = 0;
    // parcels that stop on a processor patch are remembered (index order) so that the transfer loop costs what it moves,
    // not a pass over the cloud per call - as a rank's transfer lists do in Cloud::move
    if (fresh) h.migIdx.clear();
    int nT = 1;
#ifdef _OPENMP
    nT = omp_get_max_threads();
#endif
    std::vector<std::vector<int64_t>> local((size_t)nT);
#pragma omp parallel reduction(+ : deleted, wallHits, stuck, migrated)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        std::vector<int64_t>& mine = local[(size_t)tid];
#pragma omp for schedule(static)
        for (int64_t i = begin; i < end; ++i) {
            MoveTally t;
            moveParcel(h, h.P[i], i, t, fresh);
            deleted += t.deleted; wallHits += t.wallHits; stuck += t.stuck; migrated += t.migrated;
            if (h.P[i].cell <= -2) mine.push_back(i);
        }
    }
    for (const auto& v : local) h.migIdx.insert(h.migIdx.end(), v.begin(), v.end());  // static schedule: already ascending
    h.cnt.deleted += deleted; h.cnt.wallHits += wallHits; h.cnt.stuck += stuck; h.cnt.migrated += migrated;
    h.occValid = false;
    h.momValid = false;
}

// ---------------------------------------------------------------------------------
// inflow  (uniGasGeneralBoundary.C:115-169, 537-761)
// ---------------------------------------------------------------------------------
void doInflow(ugfo_handle& h) {
    h.nBeforeInsert = (int64_t)h.P.size();
    const double dt = h.cfg.deltaT;
    const double sqrtPi = std::sqrt(PI);
    int64_t inserted = 0;
    for (const InflowPatch& ip : h.inflows) {
        const int patch = ip.patch;
        for (int lf = 0; lf < h.pSize[patch]; ++lf) {
            const int f = h.pStart[patch] + lf;
            const int bfi = f - h.nInternal;
            const int cellI = h.owner[f];
            const double* S = &h.Sf[3 * (size_t)f];
            const double* fC = &h.Cf[3 * (size_t)f];
            const double fA = std::sqrt(dot3(S, S));
            // triangle fan about the face's first point
            const int np = h.fpOff[f + 1] - h.fpOff[f];
            const int32_t* fpts = &h.fp[h.fpOff[f]];
            const double* p0 = &h.points[3 * (size_t)fpts[0]];
            std::vector<double> cTri(np - 2);
            double cum = 0;
            for (int t = 0; t < np - 2; ++t) {
                const double* a = &h.points[3 * (size_t)fpts[t + 1]];
                const double* b = &h.points[3 * (size_t)fpts[t + 2]];
                const double e1[3] = {a[0] - p0[0], a[1] - p0[1], a[2] - p0[2]};
                const double e2[3] = {b[0] - p0[0], b[1] - p0[1], b[2] - p0[2]};
                const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
                cum += 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz) / fA;
                cTri[t] = cum;
            }
            cTri[np - 3] = 1.0;
            const double n[3] = {S[0] / -fA, S[1] / -fA, S[2] / -fA};  // into the domain
            double t1[3] = {fC[0] - p0[0], fC[1] - p0[1], fC[2] - p0[2]};
            const double m1 = std::sqrt(dot3(t1, t1));
            for (int k = 0; k < 3; ++k) t1[k] /= m1;
            double t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
            const double m2 = std::sqrt(dot3(t2, t2));
            for (int k = 0; k < 3; ++k) t2[k] /= m2;
            const double* velCount = (ip.pressure || ip.fields) ? &ip.faceVel[3 * (size_t)lf] : ip.in.velocity;
            // the mass-flow inlet counts with its inlet velocity but inserts from a gas at rest: controlParcelsBeforeMove zeroes
            // inletVelocity_ before insertParcels (…MassFlowRateInletPatch.C:139-149)
            static const double atRest[3] = {0.0, 0.0, 0.0};
            const double* vel = ip.massFlow ? atRest : velCount;
            const bool perFace = ip.fields || ip.outlet;
            const double fnFace = FNc(h, cellI) * axiRWF(h, fC);  // nParticle * CWF(face cell) * RWF(face centre) (uniGasGeneralBoundary.C:154-155)
            const double Ttr = perFace ? ip.faceTtr[lf] : ip.in.translationalTemperature;
            const double Trot = perFace ? ip.faceTrot[lf] : ip.in.rotationalTemperature;
            for (int iD = 0; iD < ip.in.nTypeIds; ++iD) {
                const int typeId = ip.in.typeIds[iD];
                const ugf_species& s = h.sp[typeId];
                const double numDen = (ip.outlet || ip.massFlow) ? ip.faceN[(size_t)lf * ip.in.nTypeIds + iD]
                                      : ip.fields ? ip.faceN[(size_t)iD * h.pSize[patch] + lf] : ip.in.numberDensities[iD];
                const double cmp = std::sqrt(2.0 * kB * Ttr / s.mass);
                const double sCosFull = dot3(vel, n) / cmp;
                const double sCosCount = dot3(velCount, n) / cmp;
                const double sCos = (ip.pressure && sCosCount > 5.0) ? 5.0 : sCosCount;  // count only: the device library's insertion bound
                // Bird eq 4.22 (uniGasGeneralBoundary.C:154-165); CWF of the face's cell, RWF = 1
                double accum = ip.molFrac[iD] * (fA * numDen * dt * cmp
                                * (std::exp(-(sCos * sCos)) + sqrtPi * sCos * (1 + std::erf(sCos))))
                               / (2.0 * sqrtPi * fnFace);
                double cePressure = 0.0;
                if (ip.ce) {  // uniGasGeneralBoundary.C:171-239: normal stress and heat flux correct the Maxwellian flux
                    for (int j = 0; j < ip.in.nTypeIds; ++j) cePressure += ip.in.numberDensities[j];
                    cePressure *= kB * Ttr;
                    const double qn = dot3(ip.ceQ, n);
                    double sn[3];
                    for (int k = 0; k < 3; ++k) sn[k] = ip.ceS[3 * k] * n[0] + ip.ceS[3 * k + 1] * n[1] + ip.ceS[3 * k + 2] * n[2];
                    const double snn = dot3(sn, n);
                    accum = (fA * numDen * dt * cmp
                             * (std::exp(-(sCos * sCos)) * (1.0 - 0.5 * snn / cePressure - 0.4 * qn * sCos / cePressure / cmp)
                                + sqrtPi * sCos * (1 + std::erf(sCos))))
                            / (2.0 * sqrtPi * fnFace);
                }
                if (ip.outlet) {  // the device library's insertion bound; reaching it is an error there
                    const double cmpCap = std::sqrt(2.0 * kB * ip.capT / s.mass);
                    const double cap = ip.molFrac[iD] * (fA * ip.capN * dt * cmpCap * (std::exp(-25.0) + sqrtPi * 5.0 * (1 + std::erf(5.0))))
                                       / (2.0 * sqrtPi * fnFace);
                    if (!(accum <= cap)) { accum = accum > cap ? cap : 0.0; h.outletBoundHit = true; }
                }
                Stream rc(h.cfg.seed, KIND_INFLOW, (uint32_t)iD, (uint32_t)h.step, (uint32_t)bfi, 0);
                int nIns = std::max((int)accum, 0);
                if ((accum - nIns) > rc.u01()) ++nIns;
                for (int i = 0; i < nIns; ++i) {
                    Stream r(h.cfg.seed, KIND_INFLOW, (uint32_t)iD, (uint32_t)h.step, (uint32_t)bfi, (uint32_t)(i + 1));
                    const double triSel = r.u01();
                    int sel = 0;
                    for (int t = 0; t < np - 2; ++t) { sel = t; if (cTri[t] >= triSel) break; }
                    const double* a = &h.points[3 * (size_t)fpts[sel + 1]];
                    const double* b = &h.points[3 * (size_t)fpts[sel + 2]];
                    double bs = r.u01(), bt = r.u01();
                    if (bs + bt > 1) { bs = 1 - bs; bt = 1 - bt; }
                    Parcel np_;
                    for (int k = 0; k < 3; ++k) np_.x[k] = (1 - bs - bt) * p0[k] + bs * a[k] + bt * b[k];
                    if (ip.ce) {  // Chapman-Enskog velocity (uniGasGeneralBoundary.C:880-940; Garcia & Alder 1998, Garcia 2006)
                        double maxQ = -1.0, maxS = -1.0;
                        for (int k = 0; k < 3; ++k) maxQ = std::max(maxQ, std::fabs(ip.ceQ[k]));
                        for (int k = 0; k < 9; ++k) maxS = std::max(maxS, std::fabs(ip.ceS[k]));
                        const double breakdown = std::max(2.0 * maxQ / (cePressure * cmp), maxS / cePressure);
                        const double amplitude = 1.0 + 60.0 * breakdown;
                        const double sC = sCosFull;
                        const double lower = std::min(sC - 4.0, -5.0), upper = std::min(sC, 5.0);
                        const double uMax = 0.5 * (sC - std::sqrt(sC * sC + 2.0));
                        double Uc[3], gamma;
                        do {
                            double uN;
                            if (std::fabs(dot3(vel, n)) > VSMALL) {
                                do { uN = lower + r.u01() * (upper - lower); }
                                while ((sC - uN) / (sC - uMax) * std::exp(uMax * uMax - uN * uN) < r.u01());
                            } else {
                                uN = -std::sqrt(-std::log(1.0 - r.u01()));
                            }
                            double g1c, g2c;
                            r.gauss2(g1c, g2c);
                            for (int k = 0; k < 3; ++k) Uc[k] = g1c / std::sqrt(2.0) * t1[k] + g2c / std::sqrt(2.0) * t2[k] - uN * n[k];
                            const double* S9 = ip.ceS;
                            gamma = 1.0 + (2.0 / cmp * dot3(ip.ceQ, Uc) * (0.4 * dot3(Uc, Uc) - 1.0)
                                           - 2.0 * (S9[1] * Uc[0] * Uc[1] + S9[2] * Uc[0] * Uc[2] + S9[5] * Uc[1] * Uc[2])
                                           - S9[0] * (Uc[0] * Uc[0] - Uc[2] * Uc[2]) - S9[4] * (Uc[1] * Uc[1] - Uc[2] * Uc[2])) / cePressure;
                        } while (amplitude * r.u01() > gamma);
                        for (int k = 0; k < 3; ++k) np_.U[k] = cmp * Uc[k] + vel[k];
                    } else {
                    const double A = sCosFull + std::sqrt(sCosFull * sCosFull + 2.0);
                    const double B = 0.5 * (1.0 + sCosFull * (sCosFull - std::sqrt(sCosFull * sCosFull + 2.0)));
                    double scaling = 3.0;
                    if (sCosFull < -3) scaling = std::fabs(sCosFull) + 1;
                    double Pp = -1, uNormal, uNormalThermal;
                    if (std::fabs(dot3(vel, n)) > VSMALL) {
                        do {  // Bird eq 12.5
                            uNormalThermal = scaling * (2.0 * r.u01() - 1);
                            uNormal = uNormalThermal + sCosFull;
                            if (uNormal < 0.0) Pp = -1;
                            else Pp = 2.0 * uNormal / A * std::exp(B - uNormalThermal * uNormalThermal);
                        } while (Pp < r.u01());
                    } else {
                        uNormal = std::sqrt(-std::log(1.0 - r.u01()));
                    }
                    double g1, g2;
                    r.gauss2(g1, g2);
                    const double cth = std::sqrt(kB * Ttr / s.mass);
                    const double vt1 = dot3(t1, vel), vt2 = dot3(t2, vel);
                    for (int k = 0; k < 3; ++k)
                        np_.U[k] = cth * (g1 * t1[k] + g2 * t2[k]) + vt1 * t1[k] + vt2 * t2[k] + cmp * uNormal * n[k];
                    }
                    np_.ERot = equipartitionRotationalEnergy(r, Trot, s.rotationalDoF);
                    np_.ELevel = 0;
                    for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) np_.vib[m] = 0;
                    if (s.vibrationalDoF > 0) equipartitionVibrationalEnergyLevel(r, ip.in.vibrationalTemperature, s, np_.vib);  // uniGasGeneralBoundary.C:721-726
                    if (s.nElectronicLevels > 1) np_.ELevel = equipartitionElectronicLevel(r, ip.in.electronicTemperature, s);   // :728-735
                    np_.CWF = h.cellWF[cellI];  // uniGasGeneralBoundary.C:738
                    np_.RWF = axiRWF(h, &h.cc[3 * (size_t)cellI]);  // :739 the factor of the cell centre
                    np_.sf = 0;
                    np_.cell = cellI;
                    np_.typeId = typeId;
                    np_.newParcel = 1;
                    h.P.push_back(np_);
                    ++inserted;
                }
            }
        }
    }
    h.cnt.inserted += inserted;
    h.occValid = false;
    h.momValid = false;
}

// ---------------------------------------------------------------------------------
// cell occupancy: stable counting sort w.r.t. current array order; deleted parcels and
// parcels waiting on a processor patch are dropped.  (CloudWithModels.C:110-138)
// ---------------------------------------------------------------------------------
void buildOccupancyRaw(ugfo_handle& h) {
    const int nC = h.nCells;
    const int64_t n = (int64_t)h.P.size();
    h.occOff.assign(nC + 1, 0);
#ifdef _OPENMP
    const int nT = std::max(1, std::min(omp_get_max_threads(), (int)(n / 65536) + 1));
#else
    const int nT = 1;
#endif
    // stable parallel counting sort: thread t owns the t-th contiguous chunk of the array
    std::vector<std::vector<int32_t>> hist(nT);
#pragma omp parallel num_threads(nT)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        std::vector<int32_t>& H = hist[t];
        H.assign(nC, 0);
        const int64_t b = n * t / nT, e = n * (t + 1) / nT;
        for (int64_t i = b; i < e; ++i) if (h.P[i].cell >= 0) H[h.P[i].cell]++;
#pragma omp barrier
#pragma omp for schedule(static)
        for (int c = 0; c < nC; ++c) {
            int32_t run = 0;
            for (int k = 0; k < nT; ++k) { const int32_t v = hist[k][c]; hist[k][c] = run; run += v; }
            h.occOff[c + 1] = run;  // per-cell count, prefix-summed below
        }
#pragma omp single
        {
            for (int c = 0; c < nC; ++c) h.occOff[c + 1] += h.occOff[c];
            h.occIds.resize(h.occOff[nC]);
        }
        for (int64_t i = b; i < e; ++i) {
            const int c = h.P[i].cell;
            if (c >= 0) h.occIds[h.occOff[c] + H[c]++] = (int32_t)i;
        }
    }
    h.occValid = true;
    h.occIdentity = false;
    h.cnt.nParcels = h.occOff[nC];
}

// uniGasCloud::cellWeighting (U/clouds/uniGasCloud.C:1353-1424), called by weighting() (:203-220) right after the
// first buildCellOccupancy of evolve() (:839-842): every parcel takes the weight of the cell it now sits in; a parcel
// that came from a heavier cell is cloned floor(old/new - 1) times plus once more with the remaining probability, a
// parcel that came from a lighter cell survives with probability old/new.  Clones are appended to the cloud in
// (cell, occupancy) order, so after the second buildCellOccupancy they follow the cell's own parcels in source order.
// One uniform per parcel from its own stream (KIND_WEIGHT, step, array index) replaces the shared generator.
void cellWeighting(ugfo_handle& h) {
    int64_t cloned = 0, wdel = 0;
    const int64_t nBefore = (int64_t)h.P.size();
    for (int c = 0; c < h.nCells; ++c) {
        for (int j = h.occOff[c]; j < h.occOff[c + 1]; ++j) {
            const int32_t idx = h.occIds[j];
            // axisymmetricWeighting / axisymmetricCellWeighting (:1427-1570): the same pass on RWF(position) or on the product
            // of both factors; without axisymmetricSimulation RWF is 1 on both sides
            const double newR = axiRWF(h, h.P[idx].x);
            const double oldW = h.P[idx].CWF * h.P[idx].RWF;
            const double newW = h.cellWF[c] * newR;
            h.P[idx].CWF = h.cellWF[c];
            h.P[idx].RWF = newR;
            if (oldW == newW) continue;
            Stream r(h.cfg.seed, KIND_WEIGHT, 0, (uint32_t)h.step, (uint32_t)idx, 0);
            if (oldW > newW) {
                double prob = oldW / newW - 1.0;
                while (prob > 1.0) { h.P.push_back(h.P[idx]); ++cloned; prob -= 1.0; }
                if (prob > r.u01()) { h.P.push_back(h.P[idx]); ++cloned; }
            } else if (oldW / newW < r.u01()) {
                h.P[idx].cell = -1;
                ++wdel;
            }
        }
    }
    (void)nBefore;
    h.cnt.cloned += cloned;
    h.cnt.weightDeleted += wdel;
}

// buildCellOccupancy; the first one after a move is followed by weighting() and a second buildCellOccupancy
// (U/clouds/uniGasCloud.C:839-842, 203-220)
void buildOccupancy(ugfo_handle& h) {
    buildOccupancyRaw(h);
    if (h.weightPending) {
        h.weightPending = false;
        if (h.cellWeighted || h.cfg.axisymmetric) {
            cellWeighting(h);
            buildOccupancyRaw(h);
        }
    }
}

void reorder(ugfo_handle& h) {
    if (!h.occValid) buildOccupancy(h);
    if (h.occIdentity) return;
    const int64_t n = (int64_t)h.occIds.size();
    std::vector<Parcel>& Q = h.Pswap;  // persistent second buffer (ping-pong), like the device path
    Q.resize(n);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) { Q[j] = h.P[h.occIds[j]]; h.occIds[j] = (int32_t)j; }
    h.P.swap(Q);
    h.migIdx.clear();
    h.occIdentity = true;
}

// ---------------------------------------------------------------------------------
// cell moments  (cellMeasurements.C:408-513); slot list in DESIGN.md
// ---------------------------------------------------------------------------------
void sampleCell(ugfo_handle& h, int c) {
    const int nS = h.nSpecies;
    double* M = &h.mom[(size_t)c * nS * UGF_NMOM];
    std::fill(M, M + (size_t)nS * UGF_NMOM, 0.0);
    if (h.internalModes) std::fill(&h.momE[(size_t)c * nS * 2], &h.momE[(size_t)c * nS * 2] + (size_t)nS * 2, 0.0);
    for (int j = h.occOff[c]; j < h.occOff[c + 1]; ++j) {
        const Parcel& p = h.P[h.occIds[j]];
        double* m = M + (size_t)p.typeId * UGF_NMOM;
        const double u = p.U[0], v = p.U[1], w = p.U[2];
        const double cc = u * u + v * v + w * w;
        // slots 1, 5-7 (and 31 with axisymmetricSimulation) are the sums the XnParticle fields take: weighted by the parcel's
        // RWF (cellMeasurements.C:463-467); the cell's nParticle * CWF is applied by the consumers
        const double rw = p.RWF;
        m[0] += 1.0; m[1] += rw;
        m[2] += u; m[3] += v; m[4] += w;
        m[5] += rw * u; m[6] += rw * v; m[7] += rw * w;
        if (h.cfg.axisymmetric) m[31] += rw * cc;
        m[8] += u * u; m[9] += u * v; m[10] += u * w; m[11] += v * v; m[12] += v * w; m[13] += w * w;
        m[14] += cc;
        m[15] += cc * u; m[16] += cc * v; m[17] += cc * w;
        m[18] += p.ERot;
        m[19] += p.ERot * u; m[20] += p.ERot * v; m[21] += p.ERot * w;
        const ugf_species& S = h.sp[p.typeId];
        m[26] += elecEnergy(S, p);
        if (S.nElectronicLevels > 1 && p.ELevel < 2) h.momE[((size_t)c * nS + p.typeId) * 2 + p.ELevel] += 1.0;  // cellMeasurements.C:499-510
        if (S.vibrationalDoF > 0) {  // cellMeasurements.C:436-451, 489-492
            const double ev = vibEnergy(S, p);
            m[22] += ev;
            m[23] += ev * u; m[24] += ev * v; m[25] += ev * w;
            for (int k = 0; k < S.vibrationalDoF; ++k) m[27 + k] += p.vib[k] * kB * S.thetaV[k];
        }
    }
}

void sampleAll(ugfo_handle& h) {
    if (!h.occValid) buildOccupancy(h);
    h.mom.resize((size_t)h.nCells * h.nSpecies * UGF_NMOM);
#pragma omp parallel for schedule(dynamic, 256)
    for (int c = 0; c < h.nCells; ++c) sampleCell(h, c);
    h.momValid = true;
}

// ---------------------------------------------------------------------------------
// NTC  (noTimeCounter.C:66-343)
// ---------------------------------------------------------------------------------
// subCycle: sub-cycle index, dtSub = deltaT / nSubCycles (noTimeCounterSubCycled.C:86,190; 0 and deltaT for noTimeCounter)
void collideCell(ugfo_handle& h, int c, int subCycle, double dtSub, int64_t& cand, int64_t& coll) {
    if (h.collModelId[c] != 1) return;
    const int beg = h.occOff[c];
    const int nC = h.occOff[c + 1] - beg;
    if (nC <= 1) return;
    const int32_t* ids = &h.occIds[beg];
    // sub-cells (:96-161)
    const int32_t* L = &h.subLevels[3 * (size_t)c];
    int dimW[3] = {0, 0, 0};
    int prod = 1;
    for (int d = 0; d < 3; ++d) if (h.cfg.solutionD[d]) { dimW[d] = prod; prod *= L[d]; }
    const int nSub = L[0] * L[1] * L[2];
    std::vector<int> which;
    std::vector<std::vector<int>> sub;
    if (nSub > 1) {
        which.resize(nC);
        sub.resize(nSub);
        const double* mn = &h.bbMin[3 * (size_t)c];
        const double* mx = &h.bbMax[3 * (size_t)c];
        for (int i = 0; i < nC; ++i) {
            const Parcel& p = h.P[ids[i]];
            int sc = 0;
            for (int d = 0; d < 3; ++d) if (h.cfg.solutionD[d]) {
                int k = (int)(L[d] * (p.x[d] - mn[d]) / (mx[d] - mn[d]));
                k = k < 0 ? 0 : (k > L[d] - 1 ? L[d] - 1 : k);  // reference overflows on the max face (Appendix D.8)
                sc += k * dimW[d];
            }
            which[i] = sc;
            sub[sc].push_back(i);
        }
    }
    const double sMaxOld = h.sigmaTcRMax[c];
    // :168-184  CWF of the cell; RWF = mean over the cell's parcels of axiRWF(position)
    double rwfMean = 1.0;
    if (h.cfg.axisymmetric) {
        double sum = 0.0;
        for (int i = 0; i < nC; ++i) sum += axiRWF(h, h.P[h.occIds[beg + i]].x);
        rwfMean = sum / (double)nC;
    }
    const double selectedPairs = 0.5 * nC * (nC - 1) * (FNc(h, c) * rwfMean) * sMaxOld * dtSub / h.vol[c];
    int nCand = (int)selectedPairs;
    {
        Stream rc(h.cfg.seed, KIND_NTC, (uint32_t)subCycle, (uint32_t)h.step, (uint32_t)c, 0xFFFFFFFFu);
        if (rc.u01() < (selectedPairs - nCand)) nCand++;
    }
    cand += nCand;
    double sMax = sMaxOld;
    for (int k = 0; k < nCand; ++k) {
        Stream r(h.cfg.seed, KIND_NTC, (uint32_t)subCycle, (uint32_t)h.step, (uint32_t)c, (uint32_t)k);
        const int cP = r.position(nC);
        int cQ = -1;
        if (nSub > 1 && (int)sub[which[cP]].size() > 1) {
            const std::vector<int>& s = sub[which[cP]];
            do { cQ = s[r.position((int)s.size())]; } while (cP == cQ);
        } else {
            do { cQ = r.position(nC); } while (cP == cQ);
        }
        Parcel& pP = h.P[ids[cP]];
        Parcel& pQ = h.P[ids[cQ]];
        if (h.sp[pP.typeId].charge == -1 && h.sp[pQ.typeId].charge == -1) continue;
        const double s = sigmaTcR(h, pP, pQ);
        if (s > sMax) sMax = s;
        if ((s / sMaxOld) > r.u01()) {
            collidePair(h, r, pP, pQ);
            coll++;
        }
    }
    h.sigmaTcRMax[c] = sMax;
}

void collideAll(ugfo_handle& h) {
    if (h.cfg.binaryModel == UGF_BINARY_NONE) return;
    if (!(h.cfg.collisionModel == UGF_COLL_DSMC || h.cfg.collisionModel == UGF_COLL_HYBRID)) return;
    int64_t cand = 0, coll = 0;
    // noTimeCounterSubCycled repeats the whole pass nSubCycles times with deltaT / nSubCycles (…SubCycled.C:86-190);
    // cells are independent, so the sub-cycles of a cell can run back to back
    const int nSub = h.cfg.partnerModel == UGF_PARTNER_NTC_SUBCYCLED ? std::max(1, h.cfg.nSubCycles) : 1;
    const double dtSub = nSub > 1 ? h.cfg.deltaT / nSub : h.cfg.deltaT;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : cand, coll)
    for (int c = 0; c < h.nCells; ++c)
        for (int sub = 0; sub < nSub; ++sub) collideCell(h, c, sub, dtSub, cand, coll);
    h.cnt.collisionCandidates += cand;
    h.cnt.collisions += coll;
}

// ---------------------------------------------------------------------------------
// BGK family  (stochasticParticleBGK.C, …ESBGK.C, …SBGK.C, unifiedStochasticParticleSBGK.C)
// ---------------------------------------------------------------------------------
struct Macro {
    bool perform;
    double N, rhoN, p, T, U[3], q[3], s[6] /* shear stress xx,xy,xz,yy,yz,zz */, P[6] /* pressure tensor */, Pr, nu;
    double rhoNX, rhoMX;  // weighted sums (x F_N)
};

// calculateProperties for one cell (…USP.C:379-800).  Blends q/sigma with the previous step and
// stores them (…USP.C:777-783).
void bgkMacro(ugfo_handle& h, int c, Macro& m) {
    const int nS = h.nSpecies;
    const int model = h.cfg.bgkModel;
    const double* M = &h.mom[(size_t)c * nS * UGF_NMOM];
    const double FN = FNc(h, c);  // every parcel of the cell carries CWF = cellWF[c] after weighting()
    double N = 0, rhoM = 0, rhoNX = 0, rhoMX = 0, momX[3] = {0, 0, 0}, keX = 0;
    double muu[6] = {0, 0, 0, 0, 0, 0}, mcc = 0, mccu[3] = {0, 0, 0}, eInt = 0, eIntU[3] = {0, 0, 0};
    for (int s = 0; s < nS; ++s) {
        const double* a = M + (size_t)s * UGF_NMOM;
        const double ms = h.sp[s].mass;
        N += a[0]; rhoM += ms * a[0];
        rhoNX += a[1] * FN; rhoMX += ms * a[1] * FN;
        for (int k = 0; k < 3; ++k) momX[k] += ms * a[5 + k] * FN;
        keX += ms * a[h.cfg.axisymmetric ? 31 : 14] * FN;
        for (int k = 0; k < 6; ++k) muu[k] += ms * a[8 + k];
        mcc += ms * (a[8] + a[11] + a[13]);
        for (int k = 0; k < 3; ++k) mccu[k] += ms * a[15 + k];
        eInt += a[18] + a[22];
        for (int k = 0; k < 3; ++k) eIntU[k] += a[19 + k] + a[23 + k];
    }
    m.perform = true;
    m.N = N; m.rhoNX = rhoNX; m.rhoMX = rhoMX;
    for (int k = 0; k < 3; ++k) { m.U[k] = 0; m.q[k] = 0; }
    for (int k = 0; k < 6; ++k) { m.s[k] = 0; m.P[k] = 0; }
    m.rhoN = m.p = m.T = m.Pr = m.nu = 0;
    const double V = h.vol[c];
    if (N > VSMALL) {
        m.rhoN = rhoNX / V;
        const double rhoMMean = rhoMX / V;
        for (int k = 0; k < 3; ++k) m.U[k] = momX[k] / (rhoMMean * V);
        const double linearKEMean = 0.5 * keX / V;
        const double rhoNMean = rhoNX / V;
        m.T = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * dot3(m.U, m.U));
        m.p = m.rhoN * kB * m.T;
        static const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
        for (int k = 0; k < 6; ++k) m.P[k] = m.rhoN * (muu[k] / N - (rhoM / N) * m.U[ia[k]] * m.U[ib[k]]);
        const double sp = (m.P[0] + m.P[3] + m.P[5]) / 3.0;
        for (int k = 0; k < 6; ++k) m.s[k] = m.P[k];
        m.s[0] -= sp; m.s[3] -= sp; m.s[5] -= sp;
        // heat flux vector (…USP.C:518-549)
        const double Pfull[3][3] = {{m.P[0], m.P[1], m.P[2]}, {m.P[1], m.P[3], m.P[4]}, {m.P[2], m.P[4], m.P[5]}};
        for (int k = 0; k < 3; ++k)
            m.q[k] = m.rhoN * (0.5 * (mccu[k] / N) - 0.5 * (mcc / N) * m.U[k] + eIntU[k] / N - (eInt / N) * m.U[k])
                     - Pfull[k][0] * m.U[0] - Pfull[k][1] * m.U[1] - Pfull[k][2] * m.U[2];
        // small-sample debiasing (BGK :355, ESBGK :434, SBGK :530-534, USP :551-563)
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
            const ugf_species& S = h.sp[s];
            const double a = S.alpha;
            const double viscRef = 1.25 * (1.0 + a) * (2.0 + a) * std::sqrt(S.mass * kB * h.cfg.Tref)
                                   / (a * (5.0 - 2.0 * S.omega) * (7.0 - 2.0 * S.omega) * std::sqrt(PI) * (S.d * S.d));
            const double ns = M[(size_t)s * UGF_NMOM];
            visc += ns * viscRef * std::pow(m.T / h.cfg.Tref, S.omega);
            Pr += ns * (5.0 + S.rotationalDoF) / (7.5 + S.rotationalDoF);
        }
        visc /= N; Pr /= N;
        m.Pr = Pr;
        m.nu = (model == UGF_BGK_ESBGK ? Pr : 1.0) * m.p / visc;
    } else {
        m.perform = false;
        m.Pr = 0; m.nu = 0;
    }
    // time blending, all cells (SBGK :749-755, USP :774-783)
    const double th = h.cfg.theta, dt = h.cfg.deltaT;
    if (model == UGF_BGK_SBGK || model == UGF_BGK_USP_SBGK) {
        double* qp = &h.qPrev[3 * (size_t)c];
        for (int k = 0; k < 3; ++k) { m.q[k] = th * m.q[k] / (1.0 + 0.5 * m.Pr * m.nu * dt) + (1.0 - th) * qp[k]; qp[k] = m.q[k]; }
    }
    if (model == UGF_BGK_USP_SBGK) {
        double* spv = &h.sPrev[6 * (size_t)c];
        for (int k = 0; k < 6; ++k) { m.s[k] = th * m.s[k] / (1.0 + 0.5 * m.nu * dt) + (1.0 - th) * spv[k]; spv[k] = m.s[k]; }
    }
}

// ---- macroInterpolation: interpolationCellPoint restated (…USP.C:893-947; OpenFOAM volPointInterpolation + cellPointWeight) ----
constexpr int NIF = 22;  // interpolated values: 0 Pr, 1 nu, 2 p, 3 T, 4-6 U, 7-9 q, 10-15 shear stress, 16-21 pressure tensor

inline void macroToFields(const Macro& m, double* f) {
    f[0] = m.Pr; f[1] = m.nu; f[2] = m.p; f[3] = m.T;
    for (int k = 0; k < 3; ++k) { f[4 + k] = m.U[k]; f[7 + k] = m.q[k]; }
    for (int k = 0; k < 6; ++k) { f[10 + k] = m.s[k]; f[16 + k] = m.P[k]; }
}
inline void fieldsToMacro(const double* f, Macro& m) {
    m.Pr = f[0]; m.nu = f[1]; m.p = f[2]; m.T = f[3];
    for (int k = 0; k < 3; ++k) { m.U[k] = f[4 + k]; m.q[k] = f[7 + k]; }
    for (int k = 0; k < 6; ++k) { m.s[k] = f[10 + k]; m.P[k] = f[16 + k]; }
}
// symmetric tensor (xx,xy,xz,yy,yz,zz) -> P T P with P = I - n n
inline void projectSym(double* t, const double* n) {
    const double T[3][3] = {{t[0], t[1], t[2]}, {t[1], t[3], t[4]}, {t[2], t[4], t[5]}};
    double Pm[3][3], A[3][3], B[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Pm[i][j] = (i == j ? 1.0 : 0.0) - n[i] * n[j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = 0; for (int k = 0; k < 3; ++k) A[i][j] += Pm[i][k] * T[k][j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { B[i][j] = 0; for (int k = 0; k < 3; ++k) B[i][j] += A[i][k] * Pm[k][j]; }
    t[0] = B[0][0]; t[1] = B[0][1]; t[2] = B[0][2]; t[3] = B[1][1]; t[4] = B[1][2]; t[5] = B[2][2];
}
// cell values -> point values
void pointFields(const ugfo_handle& h, const std::vector<double>& cellF, std::vector<double>& pointF) {
    const int nP = (int)(h.cpPcOff.size() - 1);
    pointF.assign((size_t)nP * NIF, 0.0);
#pragma omp parallel for schedule(static)
    for (int p = 0; p < nP; ++p) {
        double* f = &pointF[(size_t)p * NIF];
        for (int j = h.cpPcOff[p]; j < h.cpPcOff[p + 1]; ++j) {
            const double w = h.cpW[j];
            const double* cf = &cellF[(size_t)h.cpPc[j] * NIF];
            for (int k = 0; k < NIF; ++k) f[k] += w * cf[k];
        }
        const double* n = &h.cpNormals[3 * (size_t)p];
        if (n[0] != 0.0 || n[1] != 0.0 || n[2] != 0.0) {  // symmetry patch: vectors / tensors keep their in-plane part
            for (int v = 4; v <= 7; v += 3) {
                const double vn = f[v] * n[0] + f[v + 1] * n[1] + f[v + 2] * n[2];
                for (int k = 0; k < 3; ++k) f[v + k] -= vn * n[k];
            }
            projectSym(f + 10, n);
            projectSym(f + 16, n);
        }
    }
}
// value at x in cell c: linear in the tet (cell centre, 3 face points) that contains x (cellPointWeight::findTetrahedron); if
// round-off leaves x outside every tet, the tet it is least outside of
void interpolateAt(const ugfo_handle& h, int c, const double* x, const std::vector<double>& cellF, const std::vector<double>& pointF, double* out) {
    const double* cc = &h.cc[3 * (size_t)c];
    const double tol = SMALL;
    int best = -1;
    double bestMin = -1e300, bw[4] = {1, 0, 0, 0};
    for (int t = h.cpTetOff[c]; t < h.cpTetOff[c + 1]; ++t) {
        const int32_t* tp = &h.cpTetPts[3 * (size_t)t];
        double e[3][3], r[3];
        for (int k = 0; k < 3; ++k) {
            for (int q = 0; q < 3; ++q) e[q][k] = h.cpPoints[3 * (size_t)tp[q] + k] - cc[k];
            r[k] = x[k] - cc[k];
        }
        const double c12[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2], e[1][0] * e[2][1] - e[1][1] * e[2][0]};
        const double det = e[0][0] * c12[0] + e[0][1] * c12[1] + e[0][2] * c12[2];
        if (!(std::fabs(det / h.vol[c]) > tol)) continue;
        const double c20[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2], e[2][0] * e[0][1] - e[2][1] * e[0][0]};
        const double c01[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2], e[0][0] * e[1][1] - e[0][1] * e[1][0]};
        const double l1 = (r[0] * c12[0] + r[1] * c12[1] + r[2] * c12[2]) / det;
        const double l2 = (r[0] * c20[0] + r[1] * c20[1] + r[2] * c20[2]) / det;
        const double l3 = (r[0] * c01[0] + r[1] * c01[1] + r[2] * c01[2]) / det;
        const double l0 = 1.0 - l1 - l2 - l3;
        const double mn = std::min(std::min(l0, l1), std::min(l2, l3));
        if (mn > bestMin) { bestMin = mn; best = t; bw[0] = l0; bw[1] = l1; bw[2] = l2; bw[3] = l3; }
        if (mn + tol > 0) break;  // inside: the first such tet, as the reference's search
    }
    const double* cf = &cellF[(size_t)c * NIF];
    if (best < 0) { for (int k = 0; k < NIF; ++k) out[k] = cf[k]; return; }
    const int32_t* tp = &h.cpTetPts[3 * (size_t)best];
    for (int k = 0; k < NIF; ++k)
        out[k] = bw[0] * cf[k] + bw[1] * pointF[(size_t)tp[0] * NIF + k] + bw[2] * pointF[(size_t)tp[1] * NIF + k] + bw[3] * pointF[(size_t)tp[2] * NIF + k];
}

struct InterpCtx { const std::vector<Macro>* mac; const std::vector<double>* cellF; const std::vector<double>* pointF; };

void relaxCell(ugfo_handle& h, int c, int64_t& nrel, const InterpCtx* ic = nullptr) {
    const int model = h.cfg.bgkModel;
    Macro mCell;
    if (ic) mCell = (*ic->mac)[c]; else bgkMacro(h, c, mCell);
    const Macro& m = mCell;
    const bool envelope = (model == UGF_BGK_SBGK || model == UGF_BGK_USP_SBGK);
    bool raised = false;
    if (h.collModelId[c] == 0 && m.perform) {
        const int beg = h.occOff[c];
        const int N = h.occOff[c + 1] - beg;
        const int32_t* ids = &h.occIds[beg];
        const double dt = h.cfg.deltaT;
        // number of relaxing parcels (…USP.C:917-923)
        const double pc = m.N * (1.0 - std::exp(-m.nu * dt));
        int nRel = (int)pc;
        {
            Stream rc(h.cfg.seed, KIND_BGK, 0, (uint32_t)h.step, (uint32_t)c, 0xFFFFFFFFu);
            if (rc.u01() < (pc - nRel)) nRel++;
        }
        nRel = std::min(nRel, N);
        // uniform nRel-subset: the reference shuffles the cell list 5 times and takes the first
        // nRel (…USP.C:912-915,925); here each parcel draws one key from its own stream and the
        // nRel smallest keys are taken (ties by index) - the same distribution over subsets.
        std::vector<double> key(N);
        for (int i = 0; i < N; ++i) {
            Stream r(h.cfg.seed, KIND_BGK, 0, (uint32_t)h.step, (uint32_t)c, (uint32_t)i);
            key[i] = r.u01();
        }
        std::vector<char> selected(N, 0);
        for (int i = 0; i < N; ++i) {
            int rank = 0;
            for (int j = 0; j < N; ++j) rank += (key[j] < key[i]) || (key[j] == key[i] && j < i);
            selected[i] = rank < nRel;
        }
        double& maxProb = h.maxProb[c];
        for (int i = 0; i < N; ++i) {
            if (!selected[i]) continue;
            Parcel& p = h.P[ids[i]];
            const double mass = h.sp[p.typeId].mass;
            Stream r(h.cfg.seed, KIND_BGK, 0, (uint32_t)h.step, (uint32_t)c, (uint32_t)i);
            (void)r.u01();  // the selection key
            Macro mi = mCell;  // target state: the cell's, or interpolated to the parcel's position (…USP.C:936-947)
            if (ic) {
                double f[NIF];
                interpolateAt(h, c, p.x, *ic->cellF, *ic->pointF, f);
                fieldsToMacro(f, mi);
            }
            const Macro& m = mi;
            const double u0 = std::sqrt(2.0 * kB * m.T / mass);
            double v[3];
            if (model == UGF_BGK_BGK) {  // …BGK.C:782-796
                double g[3]; r.gauss3(g);
                for (int k = 0; k < 3; ++k) v[k] = g[k] / std::sqrt(2.0);
            } else if (model == UGF_BGK_ESBGK) {  // …ESBGK.C:887-906
                double g[3]; r.gauss3(g);
                for (int k = 0; k < 3; ++k) g[k] = g[k] / std::sqrt(2.0);
                const double f = 0.5 * (1 - m.Pr) / m.Pr;
                const double S[3][3] = {{1 - f * (m.P[0] / m.p - 1), -f * (m.P[1] / m.p), -f * (m.P[2] / m.p)},
                                        {-f * (m.P[1] / m.p), 1 - f * (m.P[3] / m.p - 1), -f * (m.P[4] / m.p)},
                                        {-f * (m.P[2] / m.p), -f * (m.P[4] / m.p), 1 - f * (m.P[5] / m.p - 1)}};
                for (int k = 0; k < 3; ++k) v[k] = S[k][0] * g[0] + S[k][1] * g[1] + S[k][2] * g[2];
            } else {  // S-BGK (…SBGK.C:1009-1046) and USP (…USP.C:1047-1096)
                double coeffQ, coeffS = 0;
                if (model == UGF_BGK_SBGK) {
                    coeffQ = 2.0 * (1.0 - m.Pr);
                } else {
                    const double tau = 0.5 * m.nu * dt;
                    const double e = 1.0 + 2.0 / (std::exp(2.0 * tau) - 1.0);
                    coeffQ = 2.0 * (1.0 - m.Pr * tau * e);
                    coeffS = (1.0 - tau * e);
                }
                for (;;) {
                    double g[3]; r.gauss3(g);
                    for (int k = 0; k < 3; ++k) v[k] = g[k] / std::sqrt(2.0);
                    const double vSq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
                    const double vTr = vSq / 3.0;
                    double prob = 1.0 + coeffQ / (m.p * u0) * (m.q[0] * v[0] + m.q[1] * v[1] + m.q[2] * v[2]) * (vSq / 2.5 - 1.0);
                    if (model == UGF_BGK_USP_SBGK)
                        prob += coeffS / m.p * (m.s[0] * (v[0] * v[0] - vTr) + m.s[3] * (v[1] * v[1] - vTr) + m.s[5] * (v[2] * v[2] - vTr)
                                                + 2.0 * m.s[1] * v[0] * v[1] + 2.0 * m.s[2] * v[0] * v[2] + 2.0 * m.s[4] * v[1] * v[2]);
                    if (prob > maxProb && prob < 10.0) { maxProb = prob; raised = true; break; }
                    if (r.u01() < prob / maxProb) break;
                }
            }
            for (int k = 0; k < 3; ++k) p.U[k] = m.U[k] + u0 * v[k];
            nrel++;
        }
        // conserveMomentumAndEnergy (…USP.C:996-1045)
        const double FN = FNc(h, c);
        double keX = 0, momX[3] = {0, 0, 0};
        for (int i = 0; i < N; ++i) {
            const Parcel& p = h.P[ids[i]];
            const double mass = h.sp[p.typeId].mass;
            const double wFN = FN * p.RWF;  // CWF*RWF*nParticle of the parcel (…USP.C:1017-1022)
            keX += mass * dot3(p.U, p.U) * wFN;
            for (int k = 0; k < 3; ++k) momX[k] += mass * p.U[k] * wFN;
        }
        const double postU[3] = {momX[0] / m.rhoMX, momX[1] / m.rhoMX, momX[2] / m.rhoMX};
        const double postT = m.N / (3.0 * (m.N - 1.0) * kB * m.rhoNX) * (keX - m.rhoMX * dot3(postU, postU));
        if (postT > VSMALL) {
            const double f = std::sqrt(m.T / postT);
            for (int i = 0; i < N; ++i) {
                Parcel& p = h.P[ids[i]];
                for (int k = 0; k < 3; ++k) p.U[k] = m.U[k] + (p.U[k] - postU[k]) * f;
            }
        }
    }
    // resetProperties: envelope decay for every cell (…USP.C:859-863)
    if (envelope && !raised) h.maxProb[c] *= (model == UGF_BGK_USP_SBGK ? 0.999 : 0.9999);
}

void relaxAll(ugfo_handle& h) {
    if (h.cfg.bgkModel == UGF_BGK_NONE) return;
    if (h.cfg.macroInterpolation && !h.cpSet) { h.err = "macroInterpolation true needs ugf_set_macro_interpolation"; h.relaxFailed = true; return; }
    if (!(h.cfg.collisionModel == UGF_COLL_BGK || h.cfg.collisionModel == UGF_COLL_HYBRID)) return;
    int64_t nrel = 0;
    if (h.cfg.macroInterpolation) {
        // calculateProperties for every cell first, then the cell -> point interpolation of the target fields, then the cells
        std::vector<Macro> mac((size_t)h.nCells);
        std::vector<double> cellF((size_t)h.nCells * NIF), pointF;
#pragma omp parallel for schedule(dynamic, 256)
        for (int c = 0; c < h.nCells; ++c) { bgkMacro(h, c, mac[c]); macroToFields(mac[c], &cellF[(size_t)c * NIF]); }
        pointFields(h, cellF, pointF);
        const InterpCtx ic{&mac, &cellF, &pointF};
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nrel)
        for (int c = 0; c < h.nCells; ++c) relaxCell(h, c, nrel, &ic);
        h.cnt.bgkRelaxations += nrel;
        return;
    }
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nrel)
    for (int c = 0; c < h.nCells; ++c) relaxCell(h, c, nrel);
    h.cnt.bgkRelaxations += nrel;
}

// ---------------------------------------------------------------------------------
// hybrid decomposition  (localKnudsen.C:216-582, uniGasHybridDecomposition.C:127-174)
// ---------------------------------------------------------------------------------
constexpr int KN_NACC = 7;  // 0 N, 1 rhoN X, 2 rhoM X, 3 linearKE X, 4-6 momentum X; then nParcels X per species

void decompGeometry(ugfo_handle& h) {
    h.faceW.assign(h.nFaces, 1.0);
    h.faceMagSf.assign(h.nFaces, 0.0);
    std::vector<std::vector<int32_t>> cc(h.nCells);
    for (int f = 0; f < h.nFaces; ++f) {
        const double* S = &h.Sf[3 * (size_t)f];
        const double* C = &h.Cf[3 * (size_t)f];
        h.faceMagSf[f] = std::sqrt(dot3(S, S));
        if (f < h.nInternal) {  // surfaceInterpolation::makeWeights
            const double* cP = &h.cc[3 * (size_t)h.owner[f]];
            const double* cN = &h.cc[3 * (size_t)h.neighbour[f]];
            const double dO[3] = {C[0] - cP[0], C[1] - cP[1], C[2] - cP[2]}, dN[3] = {cN[0] - C[0], cN[1] - C[1], cN[2] - C[2]};
            const double sO = std::fabs(dot3(S, dO)), sN = std::fabs(dot3(S, dN));
            h.faceW[f] = sN / (sO + sN);
        }
    }
    for (int c = 0; c < h.nCells; ++c)  // cellCells in the order of the cell's faces
        for (int j = h.cfOff[c]; j < h.cfOff[c + 1]; ++j) {
            const int f = h.cf[j];
            if (f < h.nInternal) cc[c].push_back(h.owner[f] == c ? h.neighbour[f] : h.owner[f]);
        }
    h.ccOff.assign(h.nCells + 1, 0);
    h.ccIds.clear();
    for (int c = 0; c < h.nCells; ++c) {
        h.ccIds.insert(h.ccIds.end(), cc[c].begin(), cc[c].end());
        h.ccOff[c + 1] = (int32_t)h.ccIds.size();
    }
}

// fvc::average(fvc::interpolate(f)) followed by correctBoundaryConditions, for nComp scalars per cell (AoS).
// vec3At >= 0: components vec3At..vec3At+2 form a vector (symmetry planes mirror it).  Boundary faces: zeroGradient
// on wall / generic / processor patches, symmetry, cyclic (coupled, linear weights), empty faces take no part.
void smoothFields(const ugfo_handle& h, std::vector<double>& f, int nComp, int vec3At) {
    std::vector<double> out(f.size());
#pragma omp parallel for schedule(static)
    for (int c = 0; c < h.nCells; ++c) {
        double num[16], den = 0;
        for (int k = 0; k < nComp; ++k) num[k] = 0;
        const double* fc = &f[(size_t)c * nComp];
        for (int j = h.cfOff[c]; j < h.cfOff[c + 1]; ++j) {
            const int face = h.cf[j];
            const double A = h.faceMagSf[face];
            double val[16];
            if (face < h.nInternal) {
                const double w = h.faceW[face];
                const double* fo = &f[(size_t)h.owner[face] * nComp];
                const double* fn = &f[(size_t)h.neighbour[face] * nComp];
                for (int k = 0; k < nComp; ++k) val[k] = w * fo[k] + (1.0 - w) * fn[k];
            } else {
                const int patch = h.facePatch[face - h.nInternal];
                const int kind = h.pKind[patch];
                if (kind == UGF_PATCH_EMPTY) continue;
                for (int k = 0; k < nComp; ++k) val[k] = fc[k];
                if (kind == UGF_PATCH_SYMMETRY && vec3At >= 0) {
                    const double* S = &h.Sf[3 * (size_t)face];
                    const double n[3] = {S[0] / A, S[1] / A, S[2] / A};
                    const double vn = fc[vec3At] * n[0] + fc[vec3At + 1] * n[1] + fc[vec3At + 2] * n[2];
                    for (int k = 0; k < 3; ++k) val[vec3At + k] = fc[vec3At + k] - vn * n[k];
                } else if (kind == UGF_PATCH_CYCLIC) {
                    const int nf = h.pStart[h.pPartner[patch]] + (face - h.pStart[patch]);
                    const int q = h.owner[nf];
                    const double* S = &h.Sf[3 * (size_t)face];
                    const double* Sn = &h.Sf[3 * (size_t)nf];
                    const double An = h.faceMagSf[nf];
                    const double* C = &h.Cf[3 * (size_t)face];
                    const double* Cn = &h.Cf[3 * (size_t)nf];
                    const double* cP = &h.cc[3 * (size_t)c];
                    const double* cQ = &h.cc[3 * (size_t)q];
                    const double di = ((C[0] - cP[0]) * S[0] + (C[1] - cP[1]) * S[1] + (C[2] - cP[2]) * S[2]) / A;
                    const double dni = ((Cn[0] - cQ[0]) * Sn[0] + (Cn[1] - cQ[1]) * Sn[1] + (Cn[2] - cQ[2]) * Sn[2]) / An;
                    const double w = dni / (di + dni);
                    const double* fq = &f[(size_t)q * nComp];
                    for (int k = 0; k < nComp; ++k) val[k] = w * fc[k] + (1.0 - w) * fq[k];
                }
            }
            for (int k = 0; k < nComp; ++k) num[k] += A * val[k];
            den += A;
        }
        for (int k = 0; k < nComp; ++k) out[(size_t)c * nComp + k] = num[k] / den;
    }
    f.swap(out);
}

void neighbourhood(const ugfo_handle& h, int cell, int nLevels, std::vector<int>& nb, std::vector<char>& mark) {
    nb.clear();
    nb.push_back(cell);
    mark[cell] = 1;
    size_t first = 0;
    for (int level = 1; level <= nLevels; ++level) {
        const size_t last = nb.size();
        for (size_t i = first; i < last; ++i)
            for (int j = h.ccOff[nb[i]]; j < h.ccOff[nb[i] + 1]; ++j) {
                const int q = h.ccIds[j];
                if (!mark[q]) { mark[q] = 1; nb.push_back(q); }
            }
        first = last;
    }
    for (int q : nb) mark[q] = 0;
}

// the sequential, in-place refinement sweeps (localKnudsen.C:428-541)
void refineMask(const ugfo_handle& h, std::vector<int32_t>& id) {
    std::vector<int> nb;
    std::vector<char> mark(h.nCells, 0);
    for (int pass = 1; pass <= h.dec.refinementPasses; ++pass) {
        for (int which = 1; which >= 0; --which) {  // first the dsmc cells (1), then the bgk cells (0)
            for (int c = 0; c < h.nCells; ++c) {
                if (id[c] != which) continue;
                int same = 0, other = 0;
                for (int j = h.ccOff[c]; j < h.ccOff[c + 1]; ++j) (id[h.ccIds[j]] == which ? same : other)++;
                if (same == 0 || (same == 1 && other > 1)) { id[c] = 1 - which; continue; }
                neighbourhood(h, c, h.dec.neighborLevels, nb, mark);
                int nSame = 0;
                for (int q : nb) nSame += (id[q] == which);
                if (nSame < h.dec.maxNeighborFraction * nb.size()) { id[c] = 1 - which; continue; }
            }
        }
    }
}

void decompose(ugfo_handle& h) {
    if (!h.decompOn || h.cfg.collisionModel != UGF_COLL_HYBRID) return;
    if (!h.momValid) sampleAll(h);
    const int nS = h.nSpecies, W = KN_NACC + nS;
    const double dt = h.cfg.deltaT;
    h.decTimeSteps++;
    h.decTimeAv += dt;
    for (int c = 0; c < h.nCells; ++c) {
        const double FN = FNc(h, c);
        double* a = &h.knAcc[(size_t)c * W];
        for (int s = 0; s < nS; ++s) {
            const double* m = &h.mom[((size_t)c * nS + s) * UGF_NMOM];
            const double ms = h.sp[s].mass;
            a[0] += dt * m[0];
            a[1] += dt * (m[1] * FN);
            a[2] += dt * (ms * m[1] * FN);
            a[3] += dt * (ms * m[h.cfg.axisymmetric ? 31 : 14] * FN);
            for (int k = 0; k < 3; ++k) a[4 + k] += dt * (ms * m[5 + k] * FN);
            a[KN_NACC + s] += dt * (m[1] * FN);
        }
    }
    if (h.decTimeSteps != h.dec.decompositionInterval) return;
    const double tAv = h.decTimeAv;
    // 0 rhoN, 1 rhoM, 2 p, 3 T, 4-6 U; rhoN is not smoothed (localKnudsen.C:283-292)
    std::vector<double> F((size_t)h.nCells * 7, 0.0);
    for (int c = 0; c < h.nCells; ++c) {
        const double* a = &h.knAcc[(size_t)c * W];
        double* f = &F[(size_t)c * 7];
        if (a[0] > VSMALL) {
            const double V = h.vol[c];
            f[0] = a[1] / (tAv * V);
            f[1] = a[2] / (tAv * V);
            const double rhoMMean = a[2] / (V * tAv);
            for (int k = 0; k < 3; ++k) f[4 + k] = a[4 + k] / (rhoMMean * V * tAv);
            const double linearKEMean = 0.5 * a[3] / (V * tAv);
            const double rhoNMean = a[1] / (V * tAv);
            f[3] = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * (f[4] * f[4] + f[5] * f[5] + f[6] * f[6]));
            f[2] = f[0] * kB * f[3];
        }
    }
    {
        std::vector<double> G((size_t)h.nCells * 6);  // rhoM, p, T, U
        for (int c = 0; c < h.nCells; ++c) for (int k = 0; k < 6; ++k) G[(size_t)c * 6 + k] = F[(size_t)c * 7 + 1 + k];
        for (int pass = 1; pass <= h.dec.smoothingPasses; ++pass) smoothFields(h, G, 6, 3);
        for (int c = 0; c < h.nCells; ++c) for (int k = 0; k < 6; ++k) F[(size_t)c * 7 + 1 + k] = G[(size_t)c * 6 + k];
    }
    for (int c = 0; c < h.nCells; ++c) {
        const double* f = &F[(size_t)c * 7];
        const double* a = &h.knAcc[(size_t)c * W];
        double gRho = 0, gT = 0, gU = 0;
        const double magU = std::sqrt(f[4] * f[4] + f[5] * f[5] + f[6] * f[6]);
        for (int j = h.ccOff[c]; j < h.ccOff[c + 1]; ++j) {
            const int q = h.ccIds[j];
            const double* g = &F[(size_t)q * 7];
            const double d[3] = {h.cc[3 * (size_t)q] - h.cc[3 * (size_t)c], h.cc[3 * (size_t)q + 1] - h.cc[3 * (size_t)c + 1],
                                 h.cc[3 * (size_t)q + 2] - h.cc[3 * (size_t)c + 2]};
            const double dist = std::sqrt(dot3(d, d));
            gRho = std::max(gRho, std::fabs(g[1] - f[1]) / dist);
            gT = std::max(gT, std::fabs(g[3] - f[3]) / dist);
            gU = std::max(gU, std::fabs(std::sqrt(g[4] * g[4] + g[5] * g[5] + g[6] * g[6]) - magU) / dist);
        }
        double knRho, knT, knU, knG;
        if (a[0] > VSMALL && f[3] > VSMALL) {
            const double V = h.vol[c];
            double mfp = 0;
            for (int i = 0; i < nS; ++i) {
                double inv = 0;
                for (int q = 0; q < nS; ++q) {
                    const double dPQ = 0.5 * (h.sp[i].d + h.sp[q].d), omegaPQ = 0.5 * (h.sp[i].omega + h.sp[q].omega);
                    const double massRatio = h.sp[i].mass / h.sp[q].mass;
                    if (a[KN_NACC + q] > VSMALL) {
                        const double nDensQ = a[KN_NACC + q] / V;
                        inv += PI * dPQ * dPQ * nDensQ * std::pow(h.cfg.Tref / f[3], omegaPQ - 0.5) * std::sqrt(1.0 + massRatio);  // Bird 4.76
                    }
                }
                if (a[KN_NACC + i] > VSMALL) mfp += (1.0 / inv) * a[KN_NACC + i] / (f[0] * V);  // Bird 4.77
            }
            const double u0 = std::sqrt(2.0 * kB / (f[1] / f[0]) * f[3]);
            knRho = mfp * gRho / f[1];
            knT = mfp * gT / f[3];
            knU = mfp * gU / std::max(magU, u0);
            knG = std::max(std::max(knRho, knT), knU);
        } else {
            knRho = knT = knU = knG = 2.0 * h.dec.breakdownMax;
        }
        double* K = &h.knFields[(size_t)c * 4];
        const double th = h.dec.theta;
        K[0] = th * knRho + (1.0 - th) * K[0];
        K[1] = th * knT + (1.0 - th) * K[1];
        K[2] = th * knU + (1.0 - th) * K[2];
        K[3] = th * knG + (1.0 - th) * K[3];
    }
    for (int pass = 1; pass <= h.dec.smoothingPasses; ++pass) smoothFields(h, h.knFields, 4, -1);
    for (int c = 0; c < h.nCells; ++c) h.collModelId[c] = h.knFields[(size_t)c * 4 + 3] > h.dec.breakdownMax ? 1 : 0;
    refineMask(h, h.collModelId);
    h.decTimeSteps = 0;
    const double timeNow = (double)(h.step + 1) * dt;
    if (h.dec.resetAtDecomposition && timeNow < h.dec.resetAtDecompositionUntilTime + 0.5 * dt) {
        h.decTimeAv = 0.0;
        std::fill(h.knAcc.begin(), h.knAcc.end(), 0.0);
    }
}

// ---------------------------------------------------------------------------------
// time-averaged fields  (uniGasVolFields.C:723-1352)
// uniGasMassFlowRateInletPatch::controlParcelsAfterCollisions (…/uniGasMassFlowRateInletPatch.C:155-302): inflow velocity per
// face relaxed towards the mean velocity of the cell (kept when it would point out of the domain), number density per face and
// species from the parcels in the cell, scaled by parcelsIn / parcelsToInsert so that the next step's insertion count adds up to
// the parcels the mass flow rate brings in plus the ones that left through the patch this step (the tracker's parcelIdFlux).
// As written there: parcelsToInsert_ += adds every slot's count to EVERY species' total (a scalarField += scalar), and the
// tracker's flux is read with the patch-local species index; the second is avoided by requiring typeIds = 0..n-1 (checked).
void updateMassFlowInlet(ugfo_handle& h, InflowPatch& ip) {
    const int nT = ip.in.nTypeIds, nF = h.pSize[ip.patch];
    const double dt = h.cfg.deltaT, nParticle = h.cfg.nParticle;
    double totalMass = 0.0;
    for (int i = 0; i < nT; ++i) totalMass += h.sp[ip.in.typeIds[i]].mass * ip.mfMolFrac[i];
    std::vector<double> parcelsIn(nT, 0.0);
    double parcelsToInsert = 0.0;
    const double T = ip.in.translationalTemperature;
    for (int lf = 0; lf < nF; ++lf) {
        const int f = h.pStart[ip.patch] + lf;
        const int c = h.owner[f];
        const double* Sf = &h.Sf[3 * (size_t)f];
        const double fA = std::sqrt(dot3(Sf, Sf));
        const double CWF = h.cellWF[c], RWFf = axiRWF(h, &h.Cf[3 * (size_t)f]);
        for (int i = 0; i < nT; ++i) {
            const double moleFlowRate = ip.mfMolFrac[i] * (ip.massFlowRate / totalMass);
            parcelsIn[i] += moleFlowRate * dt * (fA / ip.patchArea) / (nParticle * CWF * RWFf) + ip.outFlux[(size_t)lf * h.nSpecies + i] / (CWF * RWFf);
        }
        double mom[3] = {0, 0, 0}, mass = 0.0;
        double* nD = &ip.faceN[(size_t)lf * nT];
        for (int i = 0; i < nT; ++i) nD[i] = 0.0;
        for (int j = h.occOff[c]; j < h.occOff[c + 1]; ++j) {
            const Parcel& p = h.P[h.occIds[j]];
            const double pMass = nParticle * h.sp[p.typeId].mass;
            const double RWF = axiRWF(h, p.x);
            nD[p.typeId] += 1.0;
            for (int k = 0; k < 3; ++k) mom[k] += pMass * CWF * RWF * p.U[k];
            mass += pMass * CWF * RWF;
        }
        double* v = &ip.faceVel[3 * (size_t)lf];
        const double prev[3] = {v[0], v[1], v[2]};
        double nv[3] = {0, 0, 0};
        if (mass > VSMALL) for (int k = 0; k < 3; ++k) nv[k] = mom[k] / mass;
        for (int k = 0; k < 3; ++k) v[k] = ip.theta * nv[k] + (1.0 - ip.theta) * prev[k];
        const double nIn[3] = {Sf[0] / -fA, Sf[1] / -fA, Sf[2] / -fA};
        if (dot3(v, nIn) < 0.0) for (int k = 0; k < 3; ++k) v[k] = prev[k];
        for (int i = 0; i < nT; ++i) nD[i] = nD[i] * nParticle * CWF * RWFf / h.vol[c];
        double pti = 0.0;  // the face's slots first, then the faces: the order of the device kernels
        for (int i = 0; i < nT; ++i) {
            const double cmp = std::sqrt(2.0 * kB * T / h.sp[ip.in.typeIds[i]].mass);
            const double sCos = dot3(v, nIn) / cmp;
            pti += (fA * nD[i] * dt * cmp * (std::exp(-(sCos * sCos)) + std::sqrt(PI) * sCos * (1 + std::erf(sCos))))
                   / (2.0 * std::sqrt(PI) * nParticle * CWF * RWFf);
        }
        parcelsToInsert += pti;
    }
    if (!(parcelsToInsert > 0.0)) {  // the reference divides by zero here (no parcel in any inlet cell): say so
        h.err = "mass-flow-rate inlet: no parcels in the cells of the inlet patch (the reference's parcelsIn / parcelsToInsert is 0 / 0 here)";
        h.relaxFailed = true;
    } else {
        for (int lf = 0; lf < nF; ++lf)
            for (int i = 0; i < nT; ++i) ip.faceN[(size_t)lf * nT + i] = ip.faceN[(size_t)lf * nT + i] * (parcelsIn[i] / parcelsToInsert);
    }
    std::fill(ip.outFlux.begin(), ip.outFlux.end(), 0.0);  // uniGasFaceTracker::clean at the end of the step (uniGasCloud.C:864)
}

// ---------------------------------------------------------------------------------
// uniGasLiouFangPressureInletPatch::controlParcelsAfterCollisions (…/uniGasLiouFangPressureInletPatch.C:126-174)
void updateInletVelocities(ugfo_handle& h) {
    bool any = false;
    for (const InflowPatch& ip : h.inflows) any = any || ip.pressure;
    if (!any) return;
    if (!h.occValid) buildOccupancy(h);
    for (InflowPatch& ip : h.inflows) {
        if (!ip.pressure) continue;
        if (ip.massFlow) { updateMassFlowInlet(h, ip); continue; }
        if (ip.outlet) {  // …/uniGasLiouFangPressureOutletPatch.C:144-322
            ip.wangSteps += 1.0;
            for (int lf = 0; lf < h.pSize[ip.patch]; ++lf) {
                const int f = h.pStart[ip.patch] + lf;
                const int c = h.owner[f];
                const double w = FNc(h, c);
                double mom[3] = {0, 0, 0}, mass = 0, nP = 0, sq[3] = {0, 0, 0}, su[3] = {0, 0, 0};
                for (int j = h.occOff[c]; j < h.occOff[c + 1]; ++j) {
                    const Parcel& p = h.P[h.occIds[j]];
                    bool mine = false;
                    for (int i = 0; i < ip.in.nTypeIds; ++i) mine = mine || ip.in.typeIds[i] == p.typeId;
                    if (mine) {
                        const double m = w * axiRWF(h, p.x) * h.sp[p.typeId].mass;  // nParticle*CWF*RWF(position)*mass
                        for (int k = 0; k < 3; ++k) mom[k] += m * p.U[k];
                        mass += m;
                    }
                    for (int k = 0; k < 3; ++k) { sq[k] += p.U[k] * p.U[k]; su[k] += p.U[k]; }
                    nP += 1.0;
                }
                double* S = &ip.wangSums[(size_t)lf * WANG_NSUM];
                S[0] += nP; S[1] += mass;
                for (int k = 0; k < 3; ++k) S[2 + k] += mom[k];
                if (S[0] > 1) {
                    for (int k = 0; k < 3; ++k) { S[5 + k] += sq[k]; S[8 + k] += su[k]; }
                    const double massDensity = S[1] / (h.vol[c] * ip.wangSteps);
                    const double numberDensity = massDensity / ip.wangM;
                    double m2 = 0, mm = 0;
                    for (int k = 0; k < 3; ++k) { m2 += S[5 + k] / S[0]; const double a = S[8 + k] / S[0]; mm += a * a; }
                    double T = (0.5 * ip.wangM) * (2.0 / (3.0 * kB)) * (m2 - mm);
                    if (T < VSMALL) T = 300.0;
                    const double pressure = numberDensity * kB * T;
                    const double sound = std::sqrt(ip.wangGammaR * T);
                    const double rhoE = massDensity + (ip.wangP - pressure) / (sound * sound);   // Liou & Fang 2000, eq 26
                    const double* Sf = &h.Sf[3 * (size_t)f];
                    const double fA = std::sqrt(dot3(Sf, Sf));
                    double* v = &ip.faceVel[3 * (size_t)lf];
                    for (int k = 0; k < 3; ++k) v[k] = S[1] > 0 ? S[2 + k] / S[1] : 0.0;
                    if (massDensity > 0) {
                        const double corr = (pressure - ip.wangP) / (massDensity * sound);
                        for (int k = 0; k < 3; ++k) v[k] += corr * -(Sf[k] / -fA);
                    }
                    const double nE = rhoE > 0 ? rhoE / ip.wangM : 0.0;
                    for (int i = 0; i < ip.in.nTypeIds; ++i) ip.faceN[(size_t)lf * ip.in.nTypeIds + i] = nE;
                    if (rhoE > 0) { const double TE = ip.wangP / ((kB / ip.wangM) * rhoE); ip.faceTtr[lf] = TE; ip.faceTrot[lf] = TE; }
                }
            }
            continue;
        }
        if (ip.wang) {  // …/uniGasWangPressureInletPatch.C:131-281
            ip.wangSteps += 1.0;
            for (int lf = 0; lf < h.pSize[ip.patch]; ++lf) {
                const int f = h.pStart[ip.patch] + lf;
                const int c = h.owner[f];
                const double w = FNc(h, c);
                double mom[3] = {0, 0, 0}, mass = 0, nP = 0, sq[3] = {0, 0, 0}, su[3] = {0, 0, 0};
                for (int j = h.occOff[c]; j < h.occOff[c + 1]; ++j) {
                    const Parcel& p = h.P[h.occIds[j]];
                    bool mine = false;
                    for (int i = 0; i < ip.in.nTypeIds; ++i) mine = mine || ip.in.typeIds[i] == p.typeId;
                    if (!mine) continue;
                    const double m = w * axiRWF(h, p.x) * h.sp[p.typeId].mass;  // nParticle*CWF*RWF(position)*mass
                    for (int k = 0; k < 3; ++k) { mom[k] += m * p.U[k]; sq[k] += p.U[k] * p.U[k]; su[k] += p.U[k]; }
                    mass += m;
                    nP += 1.0;
                }
                double* S = &ip.wangSums[(size_t)lf * WANG_NSUM];
                S[0] += nP; S[1] += mass;
                for (int k = 0; k < 3; ++k) S[2 + k] += mom[k];
                if (S[0] > 1) {
                    for (int k = 0; k < 3; ++k) { S[5 + k] += sq[k]; S[8 + k] += su[k]; }
                    const double massDensity = S[1] / (h.vol[c] * ip.wangSteps);
                    const double numberDensity = massDensity / ip.wangM;
                    double m2 = 0, mm = 0;
                    for (int k = 0; k < 3; ++k) { m2 += S[5 + k] / S[0]; const double a = S[8 + k] / S[0]; mm += a * a; }
                    double T = (0.5 * ip.wangM) * (2.0 / (3.0 * kB)) * (m2 - mm);
                    if (T < VSMALL) T = 300.0;
                    const double pressure = numberDensity * kB * T;
                    const double sound = std::sqrt(ip.wangGammaR * T);
                    const double* Sf = &h.Sf[3 * (size_t)f];
                    const double fA = std::sqrt(dot3(Sf, Sf));
                    double* v = &ip.faceVel[3 * (size_t)lf];
                    for (int k = 0; k < 3; ++k) v[k] = S[2 + k] / S[1];
                    if (ip.wangSteps > 100) {
                        const double corr = (pressure - ip.wangP) / (massDensity * sound);
                        for (int k = 0; k < 3; ++k) v[k] += corr * -(Sf[k] / -fA);  // the same unit normal the device table holds
                    }
                }
            }
            continue;
        }
        for (int lf = 0; lf < h.pSize[ip.patch]; ++lf) {
            const int c = h.owner[h.pStart[ip.patch] + lf];
            const double w = FNc(h, c);
            double mom[3] = {0, 0, 0}, mass = 0;
            for (int j = h.occOff[c]; j < h.occOff[c + 1]; ++j) {
                const Parcel& p = h.P[h.occIds[j]];
                const double m = w * axiRWF(h, p.x) * h.sp[p.typeId].mass;  // nParticle*CWF*RWF(position)*mass
                for (int k = 0; k < 3; ++k) mom[k] += m * p.U[k];
                mass += m;
            }
            for (int k = 0; k < 3; ++k) {
                const double nv = mass > 0 ? mom[k] / mass : 0.0;
                ip.faceVel[3 * (size_t)lf + k] = ip.theta * nv + (1.0 - ip.theta) * ip.faceVel[3 * (size_t)lf + k];
            }
        }
    }
}

void accumulateFields(ugfo_handle& h) {
    h.sampleCounter++;
    const double dt = h.cfg.deltaT;
    const int interval = h.cfg.sampleInterval > 0 ? h.cfg.sampleInterval : 1;
    if (interval <= h.sampleCounter) {
        h.nAvTimeSteps++;
        h.timeAvCounter += dt;
        const int nS = h.nSpecies;
#pragma omp parallel for schedule(static)
        for (int c = 0; c < h.nCells; ++c) {
            const double FN = FNc(h, c);
            double* A = &h.acc[(size_t)c * NACC];
            for (int s = 0; s < nS; ++s) {
                const double* a = &h.mom[((size_t)c * nS + s) * UGF_NMOM];
                const ugf_species& S = h.sp[s];
                const double ms = S.mass;
                A[0] += dt * a[0];
                A[1] += dt * (ms * a[0]);
                A[2] += dt * (ms * (a[8] + a[11] + a[13]));
                A[3] += dt * (ms * a[2]); A[4] += dt * (ms * a[3]); A[5] += dt * (ms * a[4]);
                A[6] += dt * a[18];
                A[7] += dt * (S.rotationalDoF * a[0]);
                A[8] += dt * (a[1] * FN);
                A[9] += dt * (ms * a[1] * FN);
                A[10] += dt * (ms * a[5] * FN); A[11] += dt * (ms * a[6] * FN); A[12] += dt * (ms * a[7] * FN);
                A[13] += dt * (ms * a[h.cfg.axisymmetric ? 31 : 14] * FN);
                A[14] += dt * (S.rotationalDoF > 0 ? a[0] : 0.0);
                A[15] += dt * ((5.0 + S.rotationalDoF) * a[0]);
                h.accS[(size_t)c * nS + s] += dt * (a[1] * FN);
                if (h.internalModes) {  // uniGasVolFields.C:775-793
                    double* I = &h.accI[((size_t)c * nS + s) * UGF_NINT];
                    I[0] += dt * a[0];
                    I[1] += dt * a[26];
                    I[2] += dt * h.momE[((size_t)c * nS + s) * 2];
                    I[3] += dt * h.momE[((size_t)c * nS + s) * 2 + 1];
                    for (int k = 0; k < UGF_MAX_VIB_MODES; ++k) I[4 + k] += dt * a[27 + k];
                }
            }
        }
        for (size_t i = 0; i < h.bm.size(); ++i) h.bacc[i] += dt * h.bm[i];
        h.sampleCounter = 0;
    }
    // boundaryMeas_.clean(); cellMeas_.clean()  (uniGasCloud.C:864-866)
    std::fill(h.bm.begin(), h.bm.end(), 0.0);
}

void deriveFields(ugfo_handle& h, double* cellF, double* wallF) {
    const double t = h.timeAvCounter;
    if (cellF) {
        for (int c = 0; c < h.nCells; ++c) {
            const double* A = &h.acc[(size_t)c * NACC];
            double* F = &cellF[(size_t)c * UGF_NFIELD];
            for (int k = 0; k < UGF_NFIELD; ++k) F[k] = 0;
            const double V = h.vol[c];
            if (A[0] > VSMALL) {
                F[0] = A[0] / t;
                F[1] = A[8] / (t * V);
                F[2] = A[9] / (t * V);
                const double rhoMMean = A[9] / (V * t);
                for (int k = 0; k < 3; ++k) F[3 + k] = A[10 + k] / (rhoMMean * V * t);
                const double linearKEMean = 0.5 * A[13] / (V * t);
                const double rhoNMean = A[8] / (V * t);
                F[6] = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * dot3(&F[3], &F[3]));
                F[9] = F[1] * kB * F[6];
            } else {
                F[0] = 0.001;  // uniGasVolFields.C:907-908
            }
            if (A[7] > VSMALL && t > VSMALL) F[7] = (2.0 / kB) * ((A[6] / t) / (A[7] / t));
            double nRotDof = 0;
            if (A[0] > VSMALL) nRotDof = A[7] / A[0];
            double totalvDof = 0, totalEDof = 0;
            if (h.internalModes) {
                // vibrational temperature (uniGasVolFields.C:930-1010): per species and mode from the mean quantum level, modes
                // weighted by their effective degrees of freedom, species by their share of the molecules with internal structure
                double vibT = 0;
                double molsElec = 0;
                for (int s2 = 0; s2 < h.nSpecies; ++s2)
                    if (h.sp[s2].nElectronicLevels > 1) molsElec += h.accI[((size_t)c * h.nSpecies + s2) * UGF_NINT];
                double elecT = 0;
                for (int s2 = 0; s2 < h.nSpecies; ++s2) {
                    const ugf_species& S = h.sp[s2];
                    const double* I = &h.accI[((size_t)c * h.nSpecies + s2) * UGF_NINT];
                    double dofSpecies = 0, vibTID = 0, dofMode[UGF_MAX_VIB_MODES] = {0, 0, 0, 0}, vibTMode[UGF_MAX_VIB_MODES] = {0, 0, 0, 0};
                    for (int v = 0; v < S.vibrationalDoF; ++v) {
                        if (I[4 + v] > VSMALL && I[0] > VSMALL) {
                            const double iMean = (I[4 + v] / I[0]) / (kB * S.thetaV[v]);
                            vibTMode[v] = S.thetaV[v] / std::log(1.0 + 1.0 / iMean);
                            dofMode[v] = (2.0 * S.thetaV[v] / vibTMode[v]) / (std::exp(S.thetaV[v] / vibTMode[v]) - 1.0);
                        }
                        dofSpecies += dofMode[v];
                    }
                    for (int v = 0; v < S.vibrationalDoF; ++v)
                        if (dofSpecies > VSMALL) vibTID += vibTMode[v] * dofMode[v] / dofSpecies;
                    totalvDof += dofSpecies;
                    if (A[14] > VSMALL && A[0] > VSMALL && I[0] > VSMALL) vibT += vibTID * I[0] / A[14];
                    // electronic temperature (:1012-1062): two-level Boltzmann ratio of the ground and first level populations
                    if (S.nElectronicLevels > 1 && I[2] > VSMALL && I[3] > VSMALL && I[3] * S.degeneracy[0] != I[2] * S.degeneracy[1]) {
                        const double elecTID = (S.electronicEnergy[1] - S.electronicEnergy[0]) /
                                               (kB * std::log((I[2] * S.degeneracy[1]) / (I[3] * S.degeneracy[0])));
                        const double fraction = I[0] / molsElec;
                        if (elecTID > VSMALL) elecT += fraction * elecTID;
                        totalEDof += fraction * ((2.0 * (I[1] / I[0])) / (kB * elecTID));
                    }
                }
                F[19] = vibT;
                F[20] = elecT;
            }
            F[8] = (3.0 * F[6] + nRotDof * F[7] + totalvDof * F[19] + totalEDof * F[20]) / (3.0 + nRotDof + totalvDof + totalEDof);
            // Mach number (:1078-1121)
            double gamma = 0, Cv_p = 0;
            if (A[0] > VSMALL) {
                const double molecularMass = A[1] / A[0];
                const double Cp = A[15] / A[0], Cv = Cp - 2.0;
                Cv_p = Cv / NA;
                gamma = Cp / Cv;
                if (F[6] > VSMALL && molecularMass > VSMALL) {
                    const double a = std::sqrt(gamma * (kB / molecularMass) * F[6]);
                    F[10] = std::sqrt(dot3(&F[3], &F[3])) / a;
                }
            }
            if (F[0] > VSMALL && F[10] > VSMALL && gamma > VSMALL && Cv_p > VSMALL) {
                F[11] = 1.0 / std::sqrt(F[0] * (double)h.nAvTimeSteps);  // densityError (:1248)
                F[17] = (1.0 / std::sqrt(F[0] * (double)h.nAvTimeSteps)) * (1.0 / (F[10] * std::sqrt(gamma)));  // velocityError (:1249)
                F[18] = (1.0 / std::sqrt(F[0] * (double)h.nAvTimeSteps)) * std::sqrt(kB / Cv_p);               // temperatureError (:1250)
            }
            // mean free path / collision rate fields (:1124-1232)
            {
                const int nS = h.nSpecies;
                const double* aS = &h.accS[(size_t)c * nS];
                double MFP = 0, MCR = 0;
                for (int i = 0; i < nS; ++i) {
                    double mfpI = 0, mcrI = 0;
                    for (int q = 0; q < nS; ++q) {
                        const double dPQ = 0.5 * (h.sp[i].d + h.sp[q].d), omegaPQ = 0.5 * (h.sp[i].omega + h.sp[q].omega);
                        const double massRatio = h.sp[i].mass / h.sp[q].mass;
                        if (aS[q] > VSMALL && F[6] > VSMALL) {
                            const double nDensQ = aS[q] / (V * t);
                            const double reducedMass = h.sp[i].mass * h.sp[q].mass / (h.sp[i].mass + h.sp[q].mass);
                            mfpI += PI * dPQ * dPQ * nDensQ * std::pow(h.cfg.Tref / F[6], omegaPQ - 0.5) * std::sqrt(1.0 + massRatio);
                            mcrI += 2.0 * std::sqrt(PI) * dPQ * dPQ * nDensQ * std::pow(F[6] / h.cfg.Tref, 1.0 - omegaPQ) * std::sqrt(2.0 * kB * h.cfg.Tref / reducedMass);
                        }
                    }
                    if (mfpI > VSMALL) mfpI = 1.0 / mfpI;
                    if (F[1] > VSMALL) {
                        const double nDensP = aS[i] / (V * t);
                        MFP += mfpI * nDensP / F[1];
                        MCR += mcrI * nDensP / F[1];
                    }
                }
                if (MFP < VSMALL) MFP = GREAT;
                F[12] = MFP;
                F[14] = MCR;
                if (MCR > VSMALL) { F[15] = 1.0 / MCR; F[16] = h.cfg.deltaT / F[15]; } else { F[15] = GREAT; F[16] = GREAT; }
                double largest = 0.0;
                for (int d = 0; d < 3; ++d) {
                    const double dim = (h.bbMax[3 * (size_t)c + d] - h.bbMin[3 * (size_t)c + d]) / h.subLevels[3 * (size_t)c + d];
                    if (h.cfg.solutionD[d] && largest < dim) largest = dim;
                }
                F[13] = largest / MFP;  // MFP > VSMALL always holds here (GREAT otherwise)
            }
        }
    }
    if (wallF) {
        for (int b = 0; b < h.nBFaces; ++b) {
            double* F = &wallF[(size_t)b * UGF_NWALLFIELD];
            for (int k = 0; k < UGF_NWALLFIELD; ++k) F[k] = 0;
            const int patch = h.facePatch[b];
            if (h.pKind[patch] != UGF_PATCH_WALL) continue;
            const double* B = &h.bacc[(size_t)b * UGF_NBM];
            const double nPart = FNc(h, h.owner[b + h.nInternal]) * axiRWF(h, &h.Cf[3 * (size_t)(b + h.nInternal)]);  // :1276-1278 CWF of the boundary cell, RWF of the face centre
            if (B[0] > VSMALL) {  // :1274-1301
                F[0] = B[0] * nPart / t;
                F[1] = B[1] * nPart / t;
                for (int k = 0; k < 3; ++k) F[2 + k] = B[3 + k] * nPart / (F[1] * t);
                const double rhoMMean = B[1] * nPart / t, linearKEMean = B[2] * nPart / t, rhoNMean = B[0] * nPart / t;
                F[5] = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * dot3(&F[2], &F[2]));
            }
            F[6] = B[8] / t;
            for (int k = 0; k < 3; ++k) F[7 + k] = B[9 + k] / t;
            const double* S = &h.Sf[3 * (size_t)(b + h.nInternal)];
            double nw[3], fA;
            unitNormal(S, nw, fA);
            F[10] = dot3(&F[7], nw);  // p = fD & n
            const double ft[3] = {F[7] - F[10] * nw[0], F[8] - F[10] * nw[1], F[9] - F[10] * nw[2]};
            F[11] = std::sqrt(dot3(ft, ft));  // tau = sqrt((fD&t1)^2 + (fD&t2)^2)
        }
    }
}

void energyTotals(ugfo_handle& h) {
    double ke = 0, er = 0, mx = 0, my = 0, mz = 0, ev = 0, ee = 0;
    int64_t n = 0;
    for (const Parcel& p : h.P) {
        if (p.cell < 0) continue;
        const double m = h.sp[p.typeId].mass;
        ke += 0.5 * m * dot3(p.U, p.U);
        er += p.ERot;
        ev += vibEnergy(h.sp[p.typeId], p);
        ee += elecEnergy(h.sp[p.typeId], p);
        mx += m * p.U[0]; my += m * p.U[1]; mz += m * p.U[2];
        ++n;
    }
    h.cnt.linearKineticEnergy = ke; h.cnt.rotationalEnergy = er;
    h.cnt.vibrationalEnergy = ev; h.cnt.electronicEnergy = ee;
    h.cnt.momentum[0] = mx; h.cnt.momentum[1] = my; h.cnt.momentum[2] = mz;
    h.cnt.nParcels = n;
}

void resetStepCounters(ugfo_handle& h) {
    h.cnt.collisionCandidates = h.cnt.collisions = h.cnt.bgkRelaxations = 0;
    h.cnt.inserted = h.cnt.deleted = h.cnt.migrated = h.cnt.wallHits = 0;
    h.cnt.cloned = h.cnt.weightDeleted = 0;
}

}  // namespace

// =====================================================================================
// C API — mirrors include/ugf.h with the prefix ugfo_ so that the same test driver can
// talk to either library.
// =====================================================================================
extern "C" {

static void countInflight(ugfo_handle* h);
int ugfo_migrate_pack(ugfo_handle* h, int32_t patch, double** buf, int64_t* n);
int ugfo_migrate_unpack(ugfo_handle* h, int32_t patch, const double* buf, int64_t n);

int ugfo_abi_version(void) { return UGF_ABI_VERSION; }

static std::string g_createErr;

const char* ugfo_last_error(const ugfo_handle* h) { return h ? h->err.c_str() : g_createErr.c_str(); }

int ugfo_create(const ugf_config* cfg, ugfo_handle** out) {
    if (!cfg || !out) { g_createErr = "null argument"; return 1; }
    if (cfg->abiVersion != UGF_ABI_VERSION) { g_createErr = "ABI version mismatch"; return 1; }
    if (cfg->axisymmetric && !(cfg->radialExtent > 0.0 && cfg->maxRWF >= 1.0)) {
        g_createErr = "axisymmetricSimulation needs radialExtentOfDomain > 0 and maxRadialWeightingFactor >= 1";
        return 1;
    }
    ugfo_handle* h = new ugfo_handle();
    h->cfg = *cfg;
    std::memset(&h->cnt, 0, sizeof(h->cnt));
    *out = h;
    return 0;
}

int ugfo_destroy(ugfo_handle* h) { delete h; return 0; }

int ugfo_set_species(ugfo_handle* h, int32_t n, const ugf_species* sp) {
    if (n < 1 || n > UGF_MAX_SPECIES) return fail(h, "species count out of range");
    for (int i = 0; i < n; ++i) {
        if (sp[i].vibrationalDoF < 0 || sp[i].vibrationalDoF > UGF_MAX_VIB_MODES) return fail(h, "bad vibrationalDoF");
        if (sp[i].nElectronicLevels < 1 || sp[i].nElectronicLevels > UGF_MAX_ELEC_LEVELS) return fail(h, "bad nElectronicLevels");
        for (int m = 0; m < sp[i].vibrationalDoF; ++m)
            if (!(sp[i].thetaV[m] > 0) || !(sp[i].thetaD[m] > 0) || !(sp[i].Zref[m] > 0) || !(sp[i].TrefZv[m] > 0))
                return fail(h, "vibrational mode needs positive characteristicVibrationalTemperature, dissociationTemperature, Zref and referenceTempForZref");
        h->sp[i] = sp[i];
        if (sp[i].vibrationalDoF > 0 || sp[i].nElectronicLevels > 1) h->internalModes = true;
    }
    h->nSpecies = n;
    return 0;
}

int ugfo_set_mesh(ugfo_handle* h, const ugf_mesh* m) {
    h->nCells = m->nCells; h->nFaces = m->nFaces; h->nInternal = m->nInternalFaces; h->nPatches = m->nPatches;
    h->nBFaces = m->nFaces - m->nInternalFaces;
    h->owner.assign(m->owner, m->owner + m->nFaces);
    h->neighbour.assign(m->neighbour, m->neighbour + m->nInternalFaces);
    h->Sf.assign(m->faceAreas, m->faceAreas + 3 * (size_t)m->nFaces);
    h->Cf.assign(m->faceCentres, m->faceCentres + 3 * (size_t)m->nFaces);
    h->cfOff.assign(m->cellFaceOffsets, m->cellFaceOffsets + m->nCells + 1);
    h->cf.assign(m->cellFaces, m->cellFaces + h->cfOff[m->nCells]);
    h->vol.assign(m->cellVolumes, m->cellVolumes + m->nCells);
    h->cc.assign(m->cellCentres, m->cellCentres + 3 * (size_t)m->nCells);
    h->bbMin.assign(m->cellBbMin, m->cellBbMin + 3 * (size_t)m->nCells);
    h->bbMax.assign(m->cellBbMax, m->cellBbMax + 3 * (size_t)m->nCells);
    h->pStart.assign(m->patchStart, m->patchStart + m->nPatches);
    h->pSize.assign(m->patchSize, m->patchSize + m->nPatches);
    h->pKind.assign(m->patchKind, m->patchKind + m->nPatches);
    h->pPartner.assign(m->patchPartner, m->patchPartner + m->nPatches);
    h->pSep.assign(m->patchSeparation, m->patchSeparation + 3 * (size_t)m->nPatches);
    if (m->points && m->facePointOffsets && m->facePoints) {
        h->points.assign(m->points, m->points + 3 * (size_t)m->nPoints);
        h->fpOff.assign(m->facePointOffsets, m->facePointOffsets + m->nFaces + 1);
        h->fp.assign(m->facePoints, m->facePoints + h->fpOff[m->nFaces]);
    }
    h->facePatch.assign(h->nBFaces, -1);
    for (int p = 0; p < h->nPatches; ++p)
        for (int k = 0; k < h->pSize[p]; ++k) h->facePatch[h->pStart[p] + k - h->nInternal] = p;
    for (int b = 0; b < h->nBFaces; ++b) if (h->facePatch[b] < 0) return fail(h, "boundary face not covered by a patch");
    h->wall.assign(h->nPatches, WallModel());
    h->sigmaTcRMax.assign(h->nCells, 0.0);
    h->collModelId.assign(h->nCells, h->cfg.collisionModel == UGF_COLL_DSMC ? 1 : 0);  // uniGasCloud.C:713,723,731
    h->subLevels.assign(3 * (size_t)h->nCells, 1);
    h->cellWF.assign(h->nCells, 1.0);
    h->cellWeighted = false;
    h->maxProb.assign(h->nCells, 1.0);
    h->qPrev.assign(3 * (size_t)h->nCells, 0.0);
    h->sPrev.assign(6 * (size_t)h->nCells, 0.0);
    h->bm.assign((size_t)h->nBFaces * UGF_NBM, 0.0);
    h->bacc.assign((size_t)h->nBFaces * UGF_NBM, 0.0);
    h->acc.assign((size_t)h->nCells * NACC, 0.0);
    h->accS.assign((size_t)h->nCells * h->nSpecies, 0.0);
    if (h->internalModes) { h->accI.assign((size_t)h->nCells * h->nSpecies * UGF_NINT, 0.0); h->momE.assign((size_t)h->nCells * h->nSpecies * 2, 0.0); }
    h->packBuf.resize(h->nPatches);
    return 0;
}

int ugfo_set_patch_model(ugfo_handle* h, int32_t patch, int32_t model, const double* prm, int32_t n) {
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->pKind[patch] != UGF_PATCH_WALL) return fail(h, "patch models apply to wall patches only");
    WallModel w;
    w.model = model;
    if (model == UGF_WALL_DIFFUSE || model == UGF_WALL_MIXED || model == UGF_WALL_CLL) {
        if (n < 4) return fail(h, "diffuse wall needs T, Ux, Uy, Uz");
        w.T = prm[0]; w.Uw[0] = prm[1]; w.Uw[1] = prm[2]; w.Uw[2] = prm[3];
        if (model == UGF_WALL_MIXED) { if (n < 5) return fail(h, "mixed wall needs diffuseFraction"); w.diffuseFraction = prm[4]; }
        if (model == UGF_WALL_CLL) {
            if (n < 7) return fail(h, "CLL wall needs normalAccommCoeff, tangentialAccommCoeff, rotEnergyAccommCoeff");
            w.alphaN = prm[4]; w.sigmaT = prm[5]; w.alphaR = prm[6];
        }
    } else if (model != UGF_WALL_SPECULAR && model != UGF_WALL_DELETION) {
        return fail(h, "unknown wall model");
    }
    h->wall[patch] = w;
    return 0;
}

int ugfo_set_patch_wall_fields(ugfo_handle* h, int32_t patch, const double* T, const double* U) {
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->pKind[patch] != UGF_PATCH_WALL || h->wall[patch].model == UGF_WALL_UNSET) return fail(h, "wall fields need a wall patch with a model");
    const size_t n = (size_t)h->pSize[patch];
    h->wall[patch].faceT.assign(T, T + n);
    h->wall[patch].faceU.assign(U, U + 3 * n);
    return 0;
}

int ugfo_set_inflow(ugfo_handle* h, int32_t patch, const ugf_inflow* in) {
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->points.empty()) return fail(h, "inflow needs mesh points/facePoints");
    InflowPatch ip; ip.patch = patch; ip.in = *in;
    h->inflows.push_back(ip);
    return 0;
}

int ugfo_set_chapman_enskog_inflow(ugfo_handle* h, int32_t patch, const ugf_inflow* in, const double* heatFlux, const double* stress) {
    if (int rc = ugfo_set_inflow(h, patch, in)) return rc;
    InflowPatch& ip = h->inflows.back();
    ip.ce = true;
    for (int k = 0; k < 3; ++k) ip.ceQ[k] = heatFlux[k];
    for (int k = 0; k < 9; ++k) ip.ceS[k] = stress[k];
    return 0;
}

int ugfo_set_inflow_fields(ugfo_handle* h, int32_t patch, int32_t nTypeIds, const int32_t* typeIds, const double* numberDensity,
                           const double* transT, const double* rotT, const double* U) {
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->points.empty()) return fail(h, "inflow needs mesh points/facePoints");
    if (!typeIds || !numberDensity || !transT || !U) return fail(h, "null inflow field");
    if (nTypeIds < 1 || nTypeIds > UGF_MAX_SPECIES) return fail(h, "inflow typeIds out of range");
    InflowPatch ip; ip.patch = patch;
    std::memset(&ip.in, 0, sizeof(ip.in));
    ip.in.nTypeIds = nTypeIds;
    for (int i = 0; i < nTypeIds; ++i) ip.in.typeIds[i] = typeIds[i];
    const size_t nF = (size_t)h->pSize[patch];
    for (size_t f = 0; f < nF; ++f) {
        if (!(transT[f] > 0.0)) return fail(h, "inflow needs a positive temperature and a non-negative number density on every face");
        for (int i = 0; i < nTypeIds; ++i)
            if (!(numberDensity[(size_t)i * nF + f] >= 0.0)) return fail(h, "inflow needs a positive temperature and a non-negative number density on every face");
    }
    ip.fields = true;
    ip.faceN.assign(numberDensity, numberDensity + (size_t)nTypeIds * nF);
    ip.faceTtr.assign(transT, transT + nF);
    if (rotT) ip.faceTrot.assign(rotT, rotT + nF); else ip.faceTrot.assign(nF, 0.0);
    ip.faceVel.assign(U, U + 3 * nF);
    h->inflows.push_back(ip);
    return 0;
}

int ugfo_set_pressure_inlet(ugfo_handle* h, int32_t patch, const ugf_pressure_inlet* pin) {
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->points.empty()) return fail(h, "inflow needs mesh points/facePoints");
    if (!(pin->theta >= 0.0 && pin->theta <= 1.0)) return fail(h, "Theta must be a value between 0 and 1");
    InflowPatch ip; ip.patch = patch;
    std::memset(&ip.in, 0, sizeof(ip.in));
    ip.in.nTypeIds = pin->nTypeIds;
    const double n = pin->inletPressure / (kB * pin->inletTemperature);  // …LiouFangPressureInletPatch.C:102
    for (int i = 0; i < pin->nTypeIds; ++i) { ip.in.typeIds[i] = pin->typeIds[i]; ip.in.numberDensities[i] = n; ip.molFrac[i] = pin->moleFractions[i]; }
    ip.in.translationalTemperature = ip.in.rotationalTemperature = ip.in.vibrationalTemperature = ip.in.electronicTemperature = pin->inletTemperature;
    ip.pressure = true;
    ip.theta = pin->theta;
    ip.faceVel.assign(3 * (size_t)h->pSize[patch], 0.0);
    h->inflows.push_back(ip);
    return 0;
}

int ugfo_set_mass_flow_inlet(ugfo_handle* h, int32_t patch, const ugf_pressure_inlet* pin, double massFlowRate, const double* initialVelocity) {
    if (!(massFlowRate > 0.0) || !(pin->inletTemperature > 0.0)) return fail(h, "mass-flow-rate inlet needs a positive massFlowRate and inletTemperature");
    if (pin->nTypeIds != h->nSpecies) return fail(h, "mass-flow-rate inlet: typeIds must list every species in typeIdList order");
    for (int i = 0; i < pin->nTypeIds; ++i) if (pin->typeIds[i] != i) return fail(h, "mass-flow-rate inlet: typeIds must list every species in typeIdList order");
    ugf_pressure_inlet q = *pin;
    q.inletPressure = 0.0;  // no pressure in this model: the number densities come from the cells
    if (int rc = ugfo_set_pressure_inlet(h, patch, &q)) return rc;
    InflowPatch& ip = h->inflows.back();
    const size_t nF = (size_t)h->pSize[patch];
    ip.massFlow = true;
    ip.massFlowRate = massFlowRate;
    for (int i = 0; i < pin->nTypeIds; ++i) ip.molFrac[i] = 1.0;  // the count takes the per-species number densities as they are
    double totalMass = 0.0;
    for (int i = 0; i < pin->nTypeIds; ++i) totalMass += h->sp[i].mass * pin->moleFractions[i];
    if (!(totalMass > 0.0)) { h->inflows.pop_back(); return fail(h, "mole fractions of the mass-flow-rate inlet sum to zero"); }
    ip.patchArea = 0.0;
    for (size_t lf = 0; lf < nF; ++lf) { const double* S = &h->Sf[3 * ((size_t)h->pStart[patch] + lf)]; ip.patchArea += std::sqrt(dot3(S, S)); }
    ip.faceN.assign(nF * pin->nTypeIds, 0.0);
    for (size_t lf = 0; lf < nF; ++lf) for (int k = 0; k < 3; ++k) ip.faceVel[3 * lf + k] = initialVelocity ? initialVelocity[k] : 0.0;
    ip.outFlux.assign(nF * h->nSpecies, 0.0);
    ip.mfMolFrac.assign(pin->moleFractions, pin->moleFractions + pin->nTypeIds);
    h->patchFlux.assign(h->nPatches, -1);
    for (size_t k = 0; k < h->inflows.size(); ++k) if (h->inflows[k].massFlow) h->patchFlux[h->inflows[k].patch] = (int)k;
    return 0;
}

int ugfo_set_wang_pressure_inlet(ugfo_handle* h, int32_t patch, const ugf_pressure_inlet* pin) {
    const int rc = ugfo_set_pressure_inlet(h, patch, pin);
    if (rc) return rc;
    InflowPatch& ip = h->inflows.back();
    double M = 0, cp = 0, cv = 0;
    for (int i = 0; i < pin->nTypeIds; ++i) {
        const ugf_species& sp = h->sp[pin->typeIds[i]];
        M += sp.mass * pin->moleFractions[i];
        cp += (5.0 + sp.rotationalDoF) * pin->moleFractions[i];
        cv += (3.0 + sp.rotationalDoF) * pin->moleFractions[i];
    }
    if (!(M > 0.0)) { h->inflows.pop_back(); return fail(h, "mole fractions of the pressure inlet sum to zero"); }
    ip.wang = true;
    ip.wangSums.assign((size_t)WANG_NSUM * h->pSize[patch], 0.0);
    ip.wangP = pin->inletPressure; ip.wangM = M; ip.wangGammaR = (cp / cv) * (kB / M);
    return 0;
}

int ugfo_set_pressure_outlet(ugfo_handle* h, int32_t patch, const ugf_pressure_inlet* pout) {
    if (!(pout->inletPressure > 0.0) || !(pout->inletTemperature > 0.0)) return fail(h, "pressure outlet needs a positive pressure and initial temperature");
    ugf_pressure_inlet q = *pout;
    q.theta = 1.0;
    const int rc = ugfo_set_wang_pressure_inlet(h, patch, &q);
    if (rc) return rc;
    InflowPatch& ip = h->inflows.back();
    const size_t nF = (size_t)h->pSize[patch];
    ip.outlet = true;
    ip.capN = 2.0 * pout->inletPressure / (kB * pout->inletTemperature);
    ip.capT = pout->inletTemperature;
    ip.faceN.assign(nF * pout->nTypeIds, 0.0);
    ip.faceTtr.assign(nF, pout->inletTemperature);
    ip.faceTrot.assign(nF, pout->inletTemperature);
    return 0;
}

int ugfo_download_inlet_velocity(ugfo_handle* h, int32_t patch, double* U) {
    for (const InflowPatch& ip : h->inflows)
        if (ip.patch == patch && ip.pressure) { std::copy(ip.faceVel.begin(), ip.faceVel.end(), U); return 0; }
    return fail(h, "no pressure inlet on this patch");
}

int ugfo_upload_parcels(ugfo_handle* h, const ugf_parcels* p) {
    if (p->n > h->cfg.parcelCapacity) return fail(h, "parcel count exceeds parcelCapacity");
    h->P.resize(p->n);
    h->migIdx.clear();
    for (int64_t i = 0; i < p->n; ++i) {
        Parcel& q = h->P[i];
        q.x[0] = p->x[i]; q.x[1] = p->y[i]; q.x[2] = p->z[i];
        q.U[0] = p->Ux[i]; q.U[1] = p->Uy[i]; q.U[2] = p->Uz[i];
        q.cell = p->cell[i];
        q.typeId = p->typeId ? p->typeId[i] : 0;
        q.ERot = p->ERot ? p->ERot[i] : 0.0;
        q.newParcel = p->newParcel ? p->newParcel[i] : 0;
        q.ELevel = p->ELevel ? p->ELevel[i] : 0;
        for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) q.vib[m] = p->vibLevel ? p->vibLevel[(size_t)i * UGF_MAX_VIB_MODES + m] : 0;
        q.sf = 0;
        if (q.cell < 0 || q.cell >= h->nCells) return fail(h, "parcel cell out of range");
        q.CWF = h->cellWF[q.cell];  // implicit weights: a parcel carries its cell's factor (true after any weighting pass)
        if (p->cellWeight && p->cellWeight[i] != q.CWF) return fail(h, "parcel cellWeight differs from the cellWeightFactor of its cell");
        if (q.typeId < 0 || q.typeId >= h->nSpecies) return fail(h, "parcel typeId out of range");
        if (q.ELevel < 0 || q.ELevel >= h->sp[q.typeId].nElectronicLevels) return fail(h, "parcel ELevel out of range");
        for (int m = 0; m < UGF_MAX_VIB_MODES; ++m)
            if (q.vib[m] < 0 || q.vib[m] > 65535 || (m >= h->sp[q.typeId].vibrationalDoF && q.vib[m] != 0)) return fail(h, "parcel vibLevel out of range");
    }
    if (h->cfg.axisymmetric) {
        // radialWeight: RWF(cell centre) as the initialisation models hand it out (default), or RWF(position) as every parcel
        // carries it after a weighting pass - the same rule as the device library (include/ugf.h, ugf_parcels)
        bool asPos = p->radialWeight != nullptr, asCentre = true;
        if (p->radialWeight)
            for (int64_t i = 0; i < p->n; ++i) {
                const Parcel& q = h->P[i];
                const double rp = axiRWF(*h, q.x), rc = axiRWF(*h, &h->cc[3 * (size_t)q.cell]), r = p->radialWeight[i];
                asPos = asPos && std::fabs(r - rp) <= 1e-6 * rp;
                asCentre = asCentre && std::fabs(r - rc) <= 1e-6 * rc;
            }
        if (!asPos && !asCentre) return fail(h, "radialWeight must be RWF(position) for every parcel or RWF(cell centre) for every parcel");
        for (int64_t i = 0; i < p->n; ++i) {
            Parcel& q = h->P[i];
            q.RWF = asPos ? axiRWF(*h, q.x) : axiRWF(*h, &h->cc[3 * (size_t)q.cell]);
        }
    }
    h->occValid = false; h->momValid = false;
    h->nBeforeInsert = p->n;
    h->cnt.nParcels = p->n;
    return 0;
}

int ugfo_upload_cell_state(ugfo_handle* h, const double* s, const int32_t* id, const int32_t* lv, const double* cwf) {
    if (s) h->sigmaTcRMax.assign(s, s + h->nCells);
    if (id) h->collModelId.assign(id, id + h->nCells);
    if (lv) {
        h->subLevels.assign(lv, lv + 3 * (size_t)h->nCells);
        for (int32_t v : h->subLevels) if (v < 1) return fail(h, "subCellLevels must be >= 1");
    }
    if (cwf) {
        for (int c = 0; c < h->nCells; ++c) if (!(cwf[c] > 0.0)) return fail(h, "cellWeightFactor must be positive");
        h->cellWF.assign(cwf, cwf + h->nCells);
        h->cellWeighted = true;
    }
    return 0;
}

int ugfo_set_deltaT(ugfo_handle* h, double dt) { h->cfg.deltaT = dt; return 0; }
int ugfo_set_time_index(ugfo_handle* h, int64_t index) { h->step = index; h->cnt.step = index; return 0; }

static void openStep(ugfo_handle* h) {
    if (!h->stepOpen) { resetStepCounters(*h); h->stepOpen = true; }
}

int ugfo_control_before_move(ugfo_handle* h) { openStep(h); doInflow(*h); return 0; }

int ugfo_move(ugfo_handle* h) {
    for (int p = 0; p < h->nPatches; ++p)
        if (h->pKind[p] == UGF_PATCH_WALL && h->wall[p].model == UGF_WALL_UNSET)
            return fail(h, "wall patch without a boundary model");  // uniGasBoundaries.C:448-488
    openStep(h);
    // parcels that were not inserted this step start the step at stepFraction 0
    for (int64_t i = 0; i < (int64_t)h->P.size(); ++i) if (!h->P[i].newParcel) h->P[i].sf = 0;
    moveRange(*h, 0, (int64_t)h->P.size(), true);
    h->weightPending = true;
    h->receivedStart = (int64_t)h->P.size();
    countInflight(h);
    return 0;
}

int ugfo_sort(ugfo_handle* h) { buildOccupancy(*h); return 0; }
int ugfo_reorder(ugfo_handle* h) { reorder(*h); return 0; }
int ugfo_sample(ugfo_handle* h) { sampleAll(*h); return 0; }

int ugfo_collide(ugfo_handle* h) {
    reorder(*h);
    if (!h->momValid) sampleAll(*h);  // sampling precedes collisions (uniGasCloud.C:846-850)
    collideAll(*h);
    return 0;
}

int ugfo_relax(ugfo_handle* h) {
    reorder(*h);
    if (!h->momValid) sampleAll(*h);
    relaxAll(*h);
    return h->relaxFailed ? 1 : 0;
}

int ugfo_accumulate_fields(ugfo_handle* h) {
    if (!h->momValid) sampleAll(*h);
    updateInletVelocities(*h);
    accumulateFields(*h);
    return 0;
}

int ugfo_set_macro_interpolation(ugfo_handle* h, const ugf_cell_point* cp) {
    if (!h->nCells) return fail(h, "mesh not set");
    const size_t nP = (size_t)cp->nPoints, nT = (size_t)cp->tetOffsets[h->nCells], nW = (size_t)cp->pointCellOffsets[nP];
    h->cpPoints.assign(cp->points, cp->points + 3 * nP);
    h->cpTetOff.assign(cp->tetOffsets, cp->tetOffsets + h->nCells + 1);
    h->cpTetPts.assign(cp->tetPoints, cp->tetPoints + 3 * nT);
    h->cpPcOff.assign(cp->pointCellOffsets, cp->pointCellOffsets + nP + 1);
    h->cpPc.assign(cp->pointCells, cp->pointCells + nW);
    h->cpW.assign(cp->pointWeights, cp->pointWeights + nW);
    h->cpNormals.assign(cp->pointNormals, cp->pointNormals + 3 * nP);
    h->cpSet = true;
    return 0;
}

int ugfo_set_decomposition(ugfo_handle* h, const ugf_decomposition* d) {
    if (h->cfg.collisionModel != UGF_COLL_HYBRID) return fail(h, "a decomposition model needs collisionModel hybrid");
    if (h->nCells == 0) return fail(h, "mesh not set");
    if (d->decompositionInterval < 1 || d->smoothingPasses < 0 || d->refinementPasses < 0 || d->neighborLevels < 0) return fail(h, "bad decomposition properties");
    h->dec = *d;
    h->decompOn = true;
    h->decTimeSteps = 0;
    h->decTimeAv = 0;
    h->knAcc.assign((size_t)h->nCells * (KN_NACC + h->nSpecies), 0.0);
    h->knFields.assign((size_t)h->nCells * 4, 0.0);
    decompGeometry(*h);
    return 0;
}

int ugfo_decompose(ugfo_handle* h) { decompose(*h); return 0; }

int ugfo_download_decomposition(ugfo_handle* h, int32_t* id, double* kn) {
    if (!h->decompOn) return fail(h, "no decomposition model set");
    if (id) std::copy(h->collModelId.begin(), h->collModelId.end(), id);
    if (kn) std::copy(h->knFields.begin(), h->knFields.end(), kn);
    return 0;
}

// ---- state checkpoint: the layout documented at ugf_state_* in unigasfoam_b200/csrc/ugf_api.cu ----------------------
namespace {
constexpr double STATE_MAGIC = 1431783237.0;
long long inletVelocityDoubles(const ugfo_handle* h) {
    long long n = 0;
    for (const InflowPatch& ip : h->inflows) if (ip.pressure) n += (long long)ip.faceVel.size() + (ip.wang ? (long long)ip.wangSums.size() + 1 : 0) + (ip.outlet ? (long long)ip.faceN.size() + 2LL * (long long)ip.faceTtr.size() : 0) + (ip.massFlow ? (long long)ip.faceN.size() : 0);
    return n;
}
long long stateDoubles(const ugfo_handle* h) {
    const long long nC = h->nCells, nS = h->nSpecies, nB = h->nBFaces;
    long long n = 8 + 6 + nC * (1 + 1 + 1 + 3 + 6 + NACC + nS) + nB * UGF_NBM;
    if (h->decompOn) n += nC * (KN_NACC + nS) + nC * 4;
    return n + inletVelocityDoubles(h) + (h->internalModes ? nC * nS * UGF_NINT : 0);
}
}  // namespace

int ugfo_state_size(ugfo_handle* h, int64_t* n) { *n = stateDoubles(h); return 0; }

int ugfo_state_save(ugfo_handle* h, double* buf, int64_t nDoubles) {
    if (nDoubles != stateDoubles(h)) return fail(h, "state buffer has the wrong size");
    double* p = buf;
    const double hdr[8] = {STATE_MAGIC, 1.0, (double)h->nCells, (double)h->nSpecies, (double)h->nBFaces, h->decompOn ? 1.0 : 0.0,
                           (double)inletVelocityDoubles(h), h->internalModes ? 1.0 : 0.0};
    p = std::copy(hdr, hdr + 8, p);
    const double sc[6] = {(double)h->step, h->timeAvCounter, (double)h->nAvTimeSteps, (double)h->sampleCounter, (double)h->decTimeSteps, h->decTimeAv};
    p = std::copy(sc, sc + 6, p);
    p = std::copy(h->sigmaTcRMax.begin(), h->sigmaTcRMax.end(), p);
    for (int32_t v : h->collModelId) *p++ = v;
    p = std::copy(h->maxProb.begin(), h->maxProb.end(), p);
    p = std::copy(h->qPrev.begin(), h->qPrev.end(), p);
    p = std::copy(h->sPrev.begin(), h->sPrev.end(), p);
    p = std::copy(h->acc.begin(), h->acc.end(), p);
    p = std::copy(h->accS.begin(), h->accS.end(), p);
    p = std::copy(h->bacc.begin(), h->bacc.end(), p);
    if (h->decompOn) { p = std::copy(h->knAcc.begin(), h->knAcc.end(), p); p = std::copy(h->knFields.begin(), h->knFields.end(), p); }
    for (const InflowPatch& ip : h->inflows) {
        if (!ip.pressure) continue;
        p = std::copy(ip.faceVel.begin(), ip.faceVel.end(), p);
        if (ip.massFlow) p = std::copy(ip.faceN.begin(), ip.faceN.end(), p);
        if (ip.wang) { p = std::copy(ip.wangSums.begin(), ip.wangSums.end(), p); *p++ = ip.wangSteps; }
        if (ip.outlet) {
            p = std::copy(ip.faceN.begin(), ip.faceN.end(), p);
            for (size_t f = 0; f < ip.faceTtr.size(); ++f) { *p++ = ip.faceTtr[f]; *p++ = ip.faceTrot[f]; }
        }
    }
    if (h->internalModes) p = std::copy(h->accI.begin(), h->accI.end(), p);  // last block, announced by header word 7
    return (p - buf) == nDoubles ? 0 : fail(h, "internal: state size mismatch");
}

int ugfo_state_load(ugfo_handle* h, const double* buf, int64_t nDoubles) {
    if (nDoubles != stateDoubles(h) || nDoubles < 14) return fail(h, "state buffer has the wrong size for this set-up");
    if (buf[0] != STATE_MAGIC || buf[1] != 1.0) return fail(h, "not a ugf state buffer (magic / version)");
    if (buf[2] != (double)h->nCells || buf[3] != (double)h->nSpecies || buf[4] != (double)h->nBFaces || buf[5] != (h->decompOn ? 1.0 : 0.0) ||
        buf[6] != (double)inletVelocityDoubles(h) || buf[7] != (h->internalModes ? 1.0 : 0.0))
        return fail(h, "state buffer was written for another mesh / species / model set-up");
    const double* p = buf + 8;
    h->step = (int64_t)p[0]; h->cnt.step = h->step; h->timeAvCounter = p[1]; h->nAvTimeSteps = (int64_t)p[2]; h->sampleCounter = (int)p[3];
    h->decTimeSteps = (int)p[4]; h->decTimeAv = p[5];
    p += 6;
    auto take = [&](std::vector<double>& v) { std::copy(p, p + v.size(), v.begin()); p += v.size(); };
    take(h->sigmaTcRMax);
    for (int32_t& v : h->collModelId) v = (int32_t)*p++;
    take(h->maxProb); take(h->qPrev); take(h->sPrev); take(h->acc); take(h->accS); take(h->bacc);
    if (h->decompOn) { take(h->knAcc); take(h->knFields); }
    for (InflowPatch& ip : h->inflows) {
        if (!ip.pressure) continue;
        take(ip.faceVel);
        if (ip.massFlow) take(ip.faceN);
        if (ip.wang) { take(ip.wangSums); ip.wangSteps = *p++; }
        if (ip.outlet) {
            take(ip.faceN);
            for (size_t f = 0; f < ip.faceTtr.size(); ++f) { ip.faceTtr[f] = *p++; ip.faceTrot[f] = *p++; }
        }
    }
    if (h->internalModes) take(h->accI);
    h->momValid = false;
    return 0;
}

int ugfo_end_step(ugfo_handle* h) { h->step++; h->cnt.step = h->step; h->stepOpen = false; return 0; }

int ugfo_finish_step(ugfo_handle* h) {
    buildOccupancy(*h);
    reorder(*h);
    sampleAll(*h);
    collideAll(*h);
    relaxAll(*h);
    if (h->relaxFailed) return 1;
    updateInletVelocities(*h);
    if (h->relaxFailed) return 1;
    accumulateFields(*h);
    decompose(*h);
    return ugfo_end_step(h);
}

int ugfo_step(ugfo_handle* h, int32_t nSteps) {
    for (int s = 0; s < nSteps; ++s) {
        resetStepCounters(*h);
        h->stepOpen = true;
        int rc;
        if (!h->inflows.empty()) doInflow(*h); else h->nBeforeInsert = (int64_t)h->P.size();
        if ((rc = ugfo_move(h))) return rc;
        if (h->cnt.migrated) return fail(h, "ugf_step is single-rank: a parcel reached a processor patch");
        buildOccupancy(*h);
        reorder(*h);
        sampleAll(*h);
        collideAll(*h);
        relaxAll(*h);
        if (h->relaxFailed) return 1;
        updateInletVelocities(*h);
        if (h->relaxFailed) return 1;
        accumulateFields(*h);
        decompose(*h);
        ugfo_end_step(h);
    }
    return 0;
}

int ugfo_migrate_counts(ugfo_handle* h, int64_t* counts) {
    for (int p = 0; p < h->nPatches; ++p) counts[p] = 0;
    for (const int64_t i : h->migIdx) { const Parcel& q = h->P[(size_t)i]; if (q.cell <= -2) counts[h->facePatch[-2 - q.cell]]++; }
    return 0;
}

int ugfo_migrate_pack(ugfo_handle* h, int32_t patch, double** buf, int64_t* n) {
    if (h->internalModes) return fail(h, "vibrational / electronic levels do not travel across processor patches yet");
    std::vector<double>& b = h->packBuf[patch];
    b.clear();
    for (const int64_t i : h->migIdx) {
        Parcel& q = h->P[(size_t)i];
        if (q.cell > -2 || h->facePatch[-2 - q.cell] != patch) continue;
        const int lface = (-2 - q.cell) + h->nInternal - h->pStart[patch];
        const double rec[UGF_MIGRATE_STRIDE] = {q.x[0], q.x[1], q.x[2], q.U[0], q.U[1], q.U[2], q.ERot, q.sf, (double)lface + 4294967296.0 * (double)q.typeId, q.CWF * q.RWF};
        b.insert(b.end(), rec, rec + UGF_MIGRATE_STRIDE);
        q.cell = -1;
    }
    *buf = b.data();
    *n = (int64_t)(b.size() / UGF_MIGRATE_STRIDE);
    return 0;
}

int ugfo_migrate_unpack(ugfo_handle* h, int32_t patch, const double* buf, int64_t n) {
    if (h->pKind[patch] != UGF_PATCH_PROCESSOR) return fail(h, "unpack on a non-processor patch");
    if (h->P.capacity() < h->P.size() + (size_t)n) h->P.reserve(h->P.size() + (size_t)n + h->P.size() / 8 + 1024);
    for (int64_t i = 0; i < n; ++i) {
        const double* r = buf + i * UGF_MIGRATE_STRIDE;
        Parcel q;
        for (int k = 0; k < 3; ++k) { q.x[k] = r[k]; q.U[k] = r[3 + k]; }
        q.ERot = r[6]; q.sf = r[7];
        const int type = (int)(r[8] * (1.0 / 4294967296.0));
        const int lf = (int)(r[8] - 4294967296.0 * type);
        if (lf < 0 || lf >= h->pSize[patch]) return fail(h, "received face index out of range");
        q.cell = h->owner[h->pStart[patch] + lf];
        q.typeId = type;
        q.CWF = r[9];  // the weight factor travels with the parcel (the product CWF * RWF: the next weighting pass only needs that)
        q.RWF = 1.0;
        q.newParcel = 0;
        h->P.push_back(q);
    }
    h->occValid = false; h->momValid = false;
    return 0;
}

int ugfo_move_received(ugfo_handle* h) {
    if (h->receivedStart < 0) return fail(h, "ugf_move_received before ugf_move");
    moveRange(*h, h->receivedStart, (int64_t)h->P.size(), false);
    h->receivedStart = (int64_t)h->P.size();
    countInflight(h);
    return 0;
}

static void countInflight(ugfo_handle* h) {
    int64_t n = 0;
    for (const int64_t i : h->migIdx) n += (h->P[(size_t)i].cell <= -2);
    h->inflight = n;
}

int ugfo_migrate_pack_slots(ugfo_handle* h, double* send, int64_t slotCapacity) {
    int k = 0;
    for (int p = 0; p < h->nPatches; ++p) {
        if (h->pKind[p] != UGF_PATCH_PROCESSOR) continue;
        double* slot = send + (size_t)k * (slotCapacity + 1) * UGF_MIGRATE_STRIDE;
        double* buf; int64_t n;
        ugfo_migrate_pack(h, p, &buf, &n);
        if (n > slotCapacity) return fail(h, "migration slot overflow (raise the slot capacity)");
        for (int j = 0; j < UGF_MIGRATE_STRIDE; ++j) slot[j] = 0.0;
        slot[0] = (double)n;
        std::copy(buf, buf + n * UGF_MIGRATE_STRIDE, slot + UGF_MIGRATE_STRIDE);
        ++k;
    }
    return 0;
}

int ugfo_migrate_unpack_slots(ugfo_handle* h, const double* recv, int64_t slotCapacity) {
    int k = 0;
    for (int p = 0; p < h->nPatches; ++p) {
        if (h->pKind[p] != UGF_PATCH_PROCESSOR) continue;
        const double* slot = recv + (size_t)k * (slotCapacity + 1) * UGF_MIGRATE_STRIDE;
        const int64_t n = (int64_t)slot[0];
        if (n < 0 || n > slotCapacity) return fail(h, "corrupt migration slot header");
        if (int rc = ugfo_migrate_unpack(h, p, slot + UGF_MIGRATE_STRIDE, n)) return rc;
        ++k;
    }
    return 0;
}

// NVLink peer-memory transfer is a device-library feature; the oracle exports the symbols so that both libraries
// bind the same table, and reports that it has no such path.
int ugfo_peer_alloc(ugfo_handle* h, int64_t, void**, unsigned char*) { return fail(h, "peer-memory transfer is not available in the CPU oracle"); }
int ugfo_peer_open(ugfo_handle* h, const unsigned char*, void**) { return fail(h, "peer-memory transfer is not available in the CPU oracle"); }
int ugfo_migrate_pack_peer(ugfo_handle* h, double* const*, uint64_t* const*, int64_t, uint64_t) { return fail(h, "peer-memory transfer is not available in the CPU oracle"); }
int ugfo_migrate_unpack_peer(ugfo_handle* h, const double*, const uint64_t*, int64_t, uint64_t) { return fail(h, "peer-memory transfer is not available in the CPU oracle"); }

int ugfo_migrate_inflight(ugfo_handle* h, int64_t** p) { *p = &h->inflight; return 0; }

int ugfo_stream(ugfo_handle*, void** s) { *s = nullptr; return 0; }

int ugfo_counters_get(ugfo_handle* h, ugf_counters* out) {
    if (h->outletBoundHit)
        return fail(h, "pressure outlet: a face asked for more parcels than the insertion bound (outlet pressure at the initial outlet temperature, speed ratio 5) allows");
    energyTotals(*h); *out = h->cnt; return 0;
}

int ugfo_num_parcels(ugfo_handle* h, int64_t* n) {
    int64_t k = 0;
    for (const Parcel& q : h->P) k += (q.cell >= 0);
    *n = k;
    return 0;
}

int ugfo_download_parcels(ugfo_handle* h, ugf_parcels* p) {
    const int64_t n = (int64_t)h->P.size();
    if (p->n < n) return fail(h, "download buffer too small");
    for (int64_t i = 0; i < n; ++i) {
        const Parcel& q = h->P[i];
        if (p->x) p->x[i] = q.x[0];
        if (p->y) p->y[i] = q.x[1];
        if (p->z) p->z[i] = q.x[2];
        if (p->Ux) p->Ux[i] = q.U[0];
        if (p->Uy) p->Uy[i] = q.U[1];
        if (p->Uz) p->Uz[i] = q.U[2];
        if (p->cell) p->cell[i] = q.cell;
        if (p->typeId) p->typeId[i] = q.typeId;
        if (p->ERot) p->ERot[i] = q.ERot;
        if (p->newParcel) p->newParcel[i] = q.newParcel;
        if (p->cellWeight) p->cellWeight[i] = q.CWF;
        if (p->radialWeight) p->radialWeight[i] = q.RWF;
        if (p->ELevel) p->ELevel[i] = q.ELevel;
        if (p->vibLevel) for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) p->vibLevel[(size_t)i * UGF_MAX_VIB_MODES + m] = q.vib[m];
    }
    p->n = n;
    return 0;
}

int ugfo_download_cell_occupancy(ugfo_handle* h, int32_t* off, int32_t* ids) {
    if (!h->occValid) return fail(h, "cell occupancy not built (call ugf_sort)");
    std::copy(h->occOff.begin(), h->occOff.end(), off);
    if (ids) std::copy(h->occIds.begin(), h->occIds.end(), ids);
    return 0;
}

int ugfo_download_cell_moments(ugfo_handle* h, double* m) {
    if (!h->momValid) return fail(h, "cell moments not sampled");
    std::copy(h->mom.begin(), h->mom.end(), m);
    return 0;
}

int ugfo_download_cell_state(ugfo_handle* h, double* s, double* mp, double* q, double* sp) {
    if (s) std::copy(h->sigmaTcRMax.begin(), h->sigmaTcRMax.end(), s);
    if (mp) std::copy(h->maxProb.begin(), h->maxProb.end(), mp);
    if (q) std::copy(h->qPrev.begin(), h->qPrev.end(), q);
    if (sp) std::copy(h->sPrev.begin(), h->sPrev.end(), sp);
    return 0;
}

int ugfo_download_fields(ugfo_handle* h, double* cellF, double* wallF, int32_t reset) {
    deriveFields(*h, cellF, wallF);
    if (reset) {
        std::fill(h->acc.begin(), h->acc.end(), 0.0);
        std::fill(h->accS.begin(), h->accS.end(), 0.0);
        std::fill(h->bacc.begin(), h->bacc.end(), 0.0);
        std::fill(h->accI.begin(), h->accI.end(), 0.0);
        h->timeAvCounter = 0; h->nAvTimeSteps = 0;
    }
    return 0;
}

int ugfo_download_internal_accumulators(ugfo_handle* h, double* accInt) {
    if (h->internalModes) std::copy(h->accI.begin(), h->accI.end(), accInt);
    else std::fill(accInt, accInt + (size_t)h->nCells * h->nSpecies * UGF_NINT, 0.0);
    return 0;
}
int ugfo_download_accumulators(ugfo_handle* h, double* acc, double* accS, double* timeAv, int64_t* nAv) {
    if (acc) std::copy(h->acc.begin(), h->acc.end(), acc);
    if (accS) std::copy(h->accS.begin(), h->accS.end(), accS);
    if (timeAv) *timeAv = h->timeAvCounter;
    if (nAv) *nAv = h->nAvTimeSteps;
    return 0;
}

int ugfo_set_face_tracker(ugfo_handle* h, int32_t n, const int32_t* faces) {
    h->faceTrack.assign(h->nFaces, 0);
    h->nTracked = 0;
    h->ft.clear();
    if (n <= 0 || !faces) return 0;
    if (h->cfg.axisymmetric)  // the same refusal as the device library: the weight a migrating parcel carries is CWF x RWF
        for (int p = 0; p < h->nPatches; ++p)
            if (h->pKind[p] == UGF_PATCH_PROCESSOR && h->pSize[p] > 0) return fail(h, "face tracker on a decomposed axisymmetric case is not supported");
    for (int k = 0; k < n; ++k) {
        if (faces[k] < 0 || faces[k] >= h->nFaces) return fail(h, "tracked face out of range");
        if (h->faceTrack[faces[k]]) return fail(h, "face listed twice in the face tracker");
        h->faceTrack[faces[k]] = k + 1;
    }
    h->nTracked = n;
    h->ft.assign((size_t)n * h->nSpecies * UGF_NFT, 0.0);
    return 0;
}

int ugfo_download_face_tracker(ugfo_handle* h, double* out, int32_t reset) {
    if (!h->nTracked) return fail(h, "no face tracker set");
    if (out) std::copy(h->ft.begin(), h->ft.end(), out);
    if (reset) std::fill(h->ft.begin(), h->ft.end(), 0.0);
    return 0;
}

int ugfo_download_boundary_meas(ugfo_handle* h, double* bm) { std::copy(h->bm.begin(), h->bm.end(), bm); return 0; }

int ugfo_phase_times(ugfo_handle*, double* ms) { for (int i = 0; i < UGF_NPHASE; ++i) ms[i] = 0; return 0; }
int ugfo_launch_count(ugfo_handle*, int64_t* n) { *n = 0; return 0; }
int ugfo_transfer_bytes(ugfo_handle*, int64_t* a, int64_t* b, int64_t* c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return 0; }
int ugfo_host_alloc(ugfo_handle*, int64_t bytes, void** ptr) { *ptr = std::malloc((size_t)std::max<int64_t>(bytes, 1)); return *ptr ? 0 : 1; }
int ugfo_host_free(ugfo_handle*, void* ptr) { std::free(ptr); return 0; }

// Known-answer hook for the RNG: one Philox4x32-10 block.
void ugfo_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
// First n uniforms of a stream, for cross-checking the CUDA generator.
void ugfo_stream_u01(uint64_t seed, uint32_t kind, uint32_t aux, uint32_t a, uint32_t b, uint32_t c, int32_t n, double* out) {
    Stream s(seed, kind, aux, a, b, c);
    for (int i = 0; i < n; ++i) out[i] = s.u01();
}
int ugfo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
// Thread count of the timed CPU baseline: bench.py sets it to the host's core count explicitly (a launcher such as torchrun
// exports OMP_NUM_THREADS=1 to its ranks).
void ugfo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"
