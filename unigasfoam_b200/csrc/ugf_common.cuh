// ugf_common.cuh — device-side types shared by the sm_100a kernels of libugf.
//
// Data layout in HBM:
//   parcels   SoA, fp64 x,y,z,Ux,Uy,Uz (+ERot), int32 cell (+uint8 typeId); two buffers (ping-pong) so that the
//             cell kernel can gather through the occupancy permutation and write cell-major order.
//   mesh      CSR cell->face slots; per slot an outward-oriented face plane {Sx,Sy,S.Cf,Sz} (32 B) and the id of
//             the cell behind it (>=0) or -(boundaryFace+1).
//   cells     offsets[nCells+1], perm[n], moments [cell][species][UGF_NMOM], accumulators [cell][NACC].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ugf.h"

namespace ugf {

// OpenFOAM DimensionedConstants defaults.
constexpr double kB = 1.38065e-23;
constexpr double NAvo = 6.02214e+23;
constexpr double PI = 3.14159265358979323846;
constexpr double TWO_PI = 6.28318530717958647692;
constexpr double VSMALL = 1e-300;
constexpr double SMALL = 1e-15;
constexpr double GREAT = 1e15;

constexpr int NACC = 16;             // time-averaged accumulators per cell
constexpr int MAX_TRACK_ITERS = 4096;
constexpr int NUM_SMS = 148;

enum { KIND_MOVE = 1, KIND_NTC = 2, KIND_BGK = 3, KIND_INFLOW = 4, KIND_WEIGHT = 5 };

// Occupancy entries >= CLONE_FLAG stand for a clone of parcel (entry - CLONE_FLAG) made by cellWeighting(): they sort
// behind the cell's own parcels, in source order, and the gather materialises them.  Parcel capacity < 2^30.
constexpr int CLONE_FLAG = 0x40000000;
constexpr int CLONE_MASK = CLONE_FLAG - 1;

struct DevSpecies {
    double mass, d, omega, alpha, E0;  // E0 = electronicEnergy[0]
    int rotDoF, charge, nElec, g0;
    int vibDoF, pad;                   // number of vibrational modes
};

// Tables of a species' vibrational modes and electronic levels (moleculeProperties: characteristicVibrationalTemperature,
// dissociationTemperature, Zref, referenceTempForZref, electronicEnergyList, degeneracyList); one device array for all species,
// reached through DevParams::spi, which is null when no species has vibrational modes or more than one electronic level.
struct DevSpeciesInt {
    double thetaV[UGF_MAX_VIB_MODES], thetaD[UGF_MAX_VIB_MODES], Zref[UGF_MAX_VIB_MODES], TrefZv[UGF_MAX_VIB_MODES];
    double elecE[UGF_MAX_ELEC_LEVELS];
    int g[UGF_MAX_ELEC_LEVELS];
};

struct DevPatch {
    int kind;             // UGF_PATCH_*
    int startBfi;         // first boundary-face index of the patch
    int size;
    int partnerStartBfi;  // cyclic: startBfi of the partner patch
    int wallModel;        // UGF_WALL_*
    int pad;
    double sep[3];
    double T;
    double Uw[3];
    double diffuseFraction;
    double alphaN, sigmaT, alphaR;  // CLL: normalAccommCoeff, tangentialAccommCoeff, rotEnergyAccommCoeff
    const double* faceT;            // *FieldPatch variants: boundaryT [size] / boundaryU [size*3] on this patch, else null
    const double* faceU;
    double* outFlux;                // generic patch carrying a mass-flow-rate inlet: this step's parcelIdFlux [size*nSpecies], else null
};

struct DevParams {
    uint64_t seed;
    double nParticle, deltaT, Tref, theta, invZrot, invZel;
    int solD[3];
    int collisionModel, binaryModel, bgkModel, nSpecies, measureWalls;
    DevSpecies sp[UGF_MAX_SPECIES];
    double pairInvGamma[UGF_MAX_SPECIES * UGF_MAX_SPECIES];  // 1/Gamma(5/2 - omega_pq)
    // cell weighting (cellWeightedSimulation): cellWeightFactor per cell, null = all 1.  cwfPrev = the factors the
    // parcels still carry when the field was replaced since the last weighting pass (cwfDirty), else = cwf.
    const double* cwf;
    const double* cwfPrev;
    int cwfDirty;
    const DevSpeciesInt* spi;  // [nSpecies] or null (ugf_internal.cuh)
    // axisymmetricSimulation: RWF(x) = 1 + rwfMaxM1 * sqrt(y^2 + z^2) / radialExtent (uniGasCloudI.H:116-120).  A parcel's RWF is
    // implicit like its CWF: RWF(position) after any weighting pass; RWF(centre of its cell) - cellRwf - for parcels the inflow
    // inserted this step and, while rwfCentre is set, for the uploaded parcels before their first move (include/ugf.h, ugf_parcels).
    // cwf is never null in this mode (a field of ones without cellWeightedSimulation).
    int axi, rwfCentre;
    double rwfMaxM1, radialExtent;
    const double* cellRwf;     // [nCells] RWF(cell centre)
    const double* bfRwf;       // [nBFaces] RWF(boundary face centre)
};

struct DevCounters {
    unsigned long long cand, coll, bgk, inserted, deleted, migrated, wallHits, stuck, cloned, wdeleted;
};

struct ParcelBuf {
    double *x, *y, *z, *ux, *uy, *uz, *erot;
    int* cell;
    uint8_t* type;
    unsigned long long* vib;  // vibrational quantum levels, 16 bits per mode, or null
    uint8_t* elev;            // electronic level, or null
};

constexpr int MIG_MAXP = 8;  // processor patches per rank the slot path handles

struct MigSlots {
    int nProc;
    int patch[MIG_MAXP];
};

struct MeshDev {
    int nCells, nBFaces, nPatches;
    const int* cfOff;         // [nCells+1]
    int planeNoZ;             // 1: every stored plane has Sz == 0 exactly (straight-extruded 2-D mesh)
    const double4* plane;     // [slots] outward plane stored as {Sx,Sy,S.Cf,Sz}
    const int* nbr;           // [slots] cell behind the face, or -(bfi+1)
    const double* rec2d;      // straight 2-D meshes with 4 slots per cell: [nCells][16] = {Sx,Sy}x4, {S.Cf}x4, nbr x4 (as ints), pad; else null
    const int* bfPatch;       // [nBFaces]
    const int* bfOwner;       // [nBFaces]
    const DevPatch* patches;  // [nPatches]
    const double* vol;        // [nCells]
    const double* bbMin;      // [nCells*3]
    const double* bbMax;      // [nCells*3]
};

__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return ax * bx + ay * by + az * bz;
}

// nParticle * CWF of a cell: every parcel of a cell carries the cell's factor after weighting(); the radial factor comes on top
// where the reference applies it
__device__ __forceinline__ double cell_fn(const DevParams& prm, int cell) {
    return prm.cwf ? prm.nParticle * __ldg(&prm.cwf[cell]) : prm.nParticle;
}
// uniGasCloud::axiRWF (U/clouds/uniGasCloudI.H:116-120); callers test prm.axi first
__device__ __forceinline__ double axi_rwf(const DevParams& prm, double y, double z) {
    const double radius = sqrt(y * y + z * z);
    return 1.0 + prm.rwfMaxM1 * radius / prm.radialExtent;
}
// the RWF a parcel of `cell` at (y, z) carries outside the move: RWF(position), or RWF(cell centre) before the first move of an
// upload that said so
__device__ __forceinline__ double parcel_rwf(const DevParams& prm, int cell, double y, double z) {
    return prm.rwfCentre ? __ldg(&prm.cellRwf[cell]) : axi_rwf(prm, y, z);
}

// asynchronous global -> shared copies (LDGSTS): no registers, every copy of a staging loop in flight at once
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ int warp_sum_int(int v) { return __reduce_add_sync(0xffffffffu, v); }

// Transposing warp reduction: every lane brings 2^LOG partial sums; on return lane l holds the warp total of
// value (l mod 2^LOG).  Costs 2^LOG - 1 (+1 if LOG < 5) 64-bit shuffles instead of 5 * 2^LOG for one butterfly
// per value: at each step a lane keeps half of its values and trades the other half with its partner.
template <int LOG>
__device__ __forceinline__ double warp_reduce_transpose(double (&v)[1 << LOG], int lane) {
#pragma unroll
    for (int s = 0; s < LOG; ++s) {
        const int m = 1 << s;
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int k = 0; k < ((1 << LOG) >> (s + 1)); ++k) {
            const double a = v[2 * k], b = v[2 * k + 1];
            const double send = up ? a : b;
            const double keep = up ? b : a;
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
    double r = v[0];
#pragma unroll
    for (int m = 1 << LOG; m < 32; m <<= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
    return r;
}

}  // namespace ugf
