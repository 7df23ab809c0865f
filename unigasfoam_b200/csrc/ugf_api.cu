// ugf_api.cu — libugf: handle, device memory and the extern "C" entry points of include/ugf.h.
//
// Host side of the B200 particle loop.  One handle owns one GPU's resident state (two SoA parcel buffers,
// the flattened mesh, per-cell arrays) and one CUDA stream; every call enqueues kernels on that stream and
// only the entry points that return host data synchronise.  There is no CPU fallback: any CUDA failure is
// reported through the status code / ugf_last_error.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off (see
// __graft_entry__.build).  -fmad=false keeps the tracking arithmetic bit-identical to the CPU oracle.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/ugf.h"
#include "ugf_bgk.cuh"
#include "ugf_cell.cuh"
#include "ugf_common.cuh"
#include "ugf_decomp.cuh"
#include "ugf_fields.cuh"
#include "ugf_inflow.cuh"
#include "ugf_internal.cuh"
#include "ugf_move.cuh"
#include "ugf_sort.cuh"

using namespace ugf;

namespace {

struct InflowHost {
    int patch;
    ugf_inflow in;
    InflowDev dev;
    int nSlots;
    long long maxInsert;
    bool pressureInlet = false;
    bool wang = false;            // uniGasWangPressureInletPatch / pressure outlet: running sums per face, step count
    bool outlet = false;          // uniGasLiouFangPressureOutletPatch: number density and temperature per face follow the flow
    bool massFlow = false;        // uniGasMassFlowRateInletPatch: number density and velocity per face follow the flow
    double mfExpected = 0.0;      //   insertions the next step will make (device-computed, read back after every update)
    double wangSteps = 0.0;
    std::vector<double> accum1;   // per (face, species) slot: expected insertions per second at F_N = 1, CWF = 1
    std::vector<int> slotCell;    // owner cell of the slot's face
    std::vector<void*> owned;
};

__global__ void set_n_kernel(const int* total, long long* dN) { *dN = *total; }

__global__ void __launch_bounds__(256) hist_kernel(const int* __restrict__ cell, const long long* dN, int* __restrict__ cellCount,
                                                   const uint8_t* __restrict__ nclone) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int c = -1;
    if (i < *dN) c = cell[i];
    const bool live = c >= 0;
    if (live && nclone && nclone[i]) atomicAdd(&cellCount[c], (int)nclone[i]);  // clones of cellWeighting()
    const unsigned liveMask = __ballot_sync(0xffffffffu, live);
    if (live) {
        const unsigned peers = __match_any_sync(liveMask, c);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cellCount[c], __popc(peers));
    }
}

}  // namespace

struct ugf_handle {
    ugf_config cfg;
    std::string err;
    cudaStream_t stream = nullptr;
    int numSMs = NUM_SMS;
    long long launches = 0;
    long long h2dBytes = 0, d2hBytes = 0, argBytes = 0;  // explicit copies and kernel-argument blocks since create

    int nSpecies = 0;
    ugf_species spHost[UGF_MAX_SPECIES];
    DevParams prm;
    bool hasRot = false, multi = false;
    bool hasVib = false, hasElec = false;   // some species has vibrational modes / more than one electronic level
    DevSpeciesInt* dSpi = nullptr;          // per-species tables of those modes (DevParams::spi)
    InterpDev interp{};                     // macroInterpolation geometry + work arrays (ugf_set_macro_interpolation)
    Macro* dMacro = nullptr;   // macroscopic state per cell (bgk_fields_kernel), allocated with the first relaxation
    bool interpSet = false;    // ugf_set_macro_interpolation done
    std::vector<void*> interpOwned;
    double* dMomI = nullptr; double* dAccI = nullptr;  // [nCells][nSpecies][UGF_NINT] per-step sums / time accumulators of the internal modes

    // mesh
    bool meshSet = false;
    int nCells = 0, nFaces = 0, nInternal = 0, nBFaces = 0, nPatches = 0, nSlots = 0;
    MeshDev mesh{};
    std::vector<DevPatch> patchesHost;
    std::vector<int> patchKind, patchStart, patchSize;
    std::vector<int32_t> ownerHost, neighbourHost, cfOffHost, cfHost, patchPartnerHost;
    std::vector<double> SfHost, CfHost, pointsHost, ccHost;
    std::vector<int32_t> fpOffHost, fpHost;
    double* dRec2d = nullptr;
    double* dMoveQD = nullptr; int* dMoveQI = nullptr;  // per-warp queues of the hop-compacting move
    int* dCfOff = nullptr; double4* dPlane = nullptr; int* dNbr = nullptr; int* dBfPatch = nullptr; int* dBfOwner = nullptr;
    DevPatch* dPatches = nullptr; double* dVol = nullptr; double* dBbMin = nullptr; double* dBbMax = nullptr; double* dBfS = nullptr;
    bool hasProcessor = false;
    int moveNF = 0;  // uniform face-slot count per cell (4 or 6), 0 = general CSR
    int moveBps = 0;          // resident CTAs per SM of the streamed move kernel (0 = the variant's default; UGF_MOVE_BPS)
    int cellTask = 0, cellFlags = 0;  // cellTask 0: chosen per launch from the mean cell occupancy
    bool moveDirect = false;  // tuning: UGF_MOVE_DIRECT=1 runs the step's move with the one-thread-per-parcel kernel
    bool moveV2 = true;       // hop-compacting streamed move (UGF_MOVE_V2=0: the lockstep kernel of round 1)

    // parcels
    long long capacity = 0;
    ParcelBuf buf[2]{};
    int cur = 0;
    double* dSf = nullptr;
    double* dWq = nullptr;          // cell weighting on a decomposed case: factor carried by each parcel in flight
    long long* dN = nullptr;
    long long nUpper = 0;           // host upper bound of *dN
    long long* pinN = nullptr;      // pinned readback of *dN
    cudaEvent_t evN = nullptr;
    bool nPending = false;
    long long appendBound = 0;      // upper bound of the parcels appended since the pending read-back was requested
    long long newFrom = 0;          // parcels >= newFrom were inserted this step
    bool inflowDone = false;        // control_before_move ran since the last move
    bool stepOpen = false;          // phase-wise driving: counters were reset for the current step
    long long recvStart = -1;

    // cells
    int* dCellCount = nullptr; int* dOff = nullptr; int* dPerm = nullptr; int* dBlockSums = nullptr; int* dTotal = nullptr;
    int* dMigCount = nullptr; int* dMigBlock = nullptr; int* dMigTotals = nullptr; int* dMigList = nullptr;
    long long lastSlotCapacity = 0;  // slot capacity of the latest ugf_migrate_unpack_slots
    bool migSearchPack = false;      // tuning: UGF_MIG_SEARCH_PACK=1 packs by searching the cell ids instead of the migrant lists
    unsigned long long* dInflight = nullptr; long long* dRecvStart = nullptr;
    long long nAtMove = 0;          // exact array length when the step's move was launched
    bool slotRound = false;         // received parcels of this round came through the slot path
    bool migFused = true;           // peer-memory rounds in two launches (pack + signal, wait + unpack + move); UGF_MIG_FUSED=0: six
    bool recvMoved = false;         // the fused receive kernel has already continued the tracks of this round's parcels
    unsigned int* dMigDone = nullptr;  // block counter of the fused receive kernel
    bool nExact = true;             // nUpper is the exact array length (needed by the exact-count unpack)
    MigSlots migSlots{};
    double* dAccS = nullptr;  // multi-species: per-species nParcelsXnParticle accumulators
    double* dMom = nullptr; double* dAcc = nullptr; double* dBm = nullptr; double* dBacc = nullptr;
    double* dSigma = nullptr; int* dCollId = nullptr; double* dMaxProb = nullptr; double* dQPrev = nullptr; double* dSPrev = nullptr;
    double* dKeyScratch = nullptr;
    int* dOwner = nullptr;           // NTC conflict marks, one per parcel slot
    int* dSubLevels = nullptr;       // [nCells*3] sub-cell levels (allocated when some level > 1)
    unsigned short* dSub = nullptr;  // [capacity] sub-cell index per parcel
    DevCounters* dCnt = nullptr;
    int* dErr = nullptr;
    int* dTask = nullptr;            // cell_kernel task counters (two, used alternately: each launch zeroes the other one)
    int taskSel = 0;
    bool cellCountZero = false;      // dCellCount is all zero (left so by the segment sort)
    double* dTot = nullptr;
    bool histValid = false, occValid = false, occIdentity = false, momValid = false;
    // cell weighting (cellWeightedSimulation)
    // axisymmetricSimulation: RWF of cell centres / boundary face centres; rwfCentre = the uploaded parcels carry RWF(cell centre)
    // until their first move (DevParams)
    double* dCellRwf = nullptr;
    double* dBfRwf = nullptr;
    std::vector<double> cellRwfHost, bfRwfHost;
    bool rwfCentre = false;
    double* dCwf[2] = {nullptr, nullptr};  // [cwfCur]: current cellWeightFactor; the other: the factors the parcels carry while cwfDirty
    int cwfCur = 0;
    bool cwfDirty = false;
    std::vector<double> cwfHost, cwfHostPrev;
    uint8_t* dNclone = nullptr;            // [capacity] clones per parcel decided by the move
    // face tracker
    int nTracked = 0;
    int* dSlotTrack = nullptr; int* dBfTrack = nullptr; double* dFt = nullptr;
    std::vector<int> slotFaceHost;         // face label of every stored face slot (set_mesh), -1 for a padding slot
    std::vector<int> slotOffHost;          // [nCells+1] first slot of every cell
    bool cloneValid = false;               // dNclone belongs to the current (not yet gathered) array
    bool subLevelsAllOne = true;

    // localKnudsen hybrid decomposition
    bool decompOn = false;
    ugf_decomposition dec{};
    int decTimeSteps = 0;
    double decTimeAv = 0;
    DecompGeom dgeom{};
    double* dKnAcc = nullptr; double* dKnF[2] = {nullptr, nullptr}; double* dKnK[2] = {nullptr, nullptr};
    int knCur = 0;                       // which dKnK buffer holds the Knudsen fields
    std::vector<void*> decompOwned;
    std::vector<int32_t> ccOffHost, ccIdsHost;  // mesh.cellCells() for the refinement sweeps

    std::vector<InflowHost> inflows;
    std::vector<double*> wallFieldOwned;      // boundaryT / boundaryU of *FieldPatch walls
    std::vector<void*> hostOwned;              // pinned host buffers handed out by ugf_host_alloc
    std::vector<void*> peerOwned, peerOpened;  // NVLink peer-memory transfer buffers (cudaIpc)
    std::vector<double*> packBuf;
    std::vector<long long> packCap;

    // fields
    double timeAvCounter = 0;
    long long nAvTimeSteps = 0;
    int sampleCounter = 0;

    long long step = 0;
    int cellCap = CELL_CAP;
    int cellBlocks = 0, ntcBlocks = 0, bgkBlocks = 0;
    size_t cellSmem = 0, ntcSmem = 0, bgkSmem = 0;
    cudaEvent_t ev[UGF_NPHASE + 1]{};
    
    bool timingValid = false;
};

namespace {

std::string g_createErr;

// every explicit host <-> device copy goes through these two so that ugf_transfer_bytes can report what a step really moved
inline void count_copy(ugf_handle* h, size_t bytes, cudaMemcpyKind kind) {
    if (kind == cudaMemcpyHostToDevice) h->h2dBytes += (long long)bytes;
    else if (kind == cudaMemcpyDeviceToHost) h->d2hBytes += (long long)bytes;
}
inline cudaError_t countedMemcpyAsync(ugf_handle* h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    count_copy(h, bytes, kind);
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
}
inline cudaError_t countedMemcpy(ugf_handle* h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    count_copy(h, bytes, kind);
    return cudaMemcpy(dst, src, bytes, kind);
}

int fail(ugf_handle* h, const std::string& m) {
    if (h) h->err = m; else g_createErr = m;
    return 1;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define LAUNCHED()                                                                                 \
    do {                                                                                           \
        h->launches++;                                                                             \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, std::string("kernel launch: ") + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

template <class F>
void dispatch(const ugf_handle* h, F&& f) {
    if (h->hasRot) {
        if (h->multi) f(std::true_type{}, std::true_type{}); else f(std::true_type{}, std::false_type{});
    } else {
        if (h->multi) f(std::false_type{}, std::true_type{}); else f(std::false_type{}, std::false_type{});
    }
}

template <class T>
int dalloc(ugf_handle* h, T** p, size_t n) {
    CU(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

template <class T>
int upload(ugf_handle* h, T* dst, const T* src, size_t n) {
    if (n) CU(countedMemcpyAsync(h, dst, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

// bytes of a kernel's argument block (what the launch sends to the device besides the grid), summed over its arguments
template <class... A>
constexpr long long arg_bytes(const A&...) { return (long long)(0 + ... + sizeof(A)); }

inline unsigned grid_for(long long n, int block) { return (unsigned)std::max<long long>(1, (n + block - 1) / block); }

int check_device_error(ugf_handle* h) {
    int e = 0;
    CU(countedMemcpyAsync(h, &e, h->dErr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (e == 1) return fail(h, "parcel capacity exceeded while inserting or cloning parcels (raise parcelCapacity)");
    if (e == 2) return fail(h, "received parcel with a face index outside the processor patch");
    if (e == 3) return fail(h, "migration slot overflow: more parcels crossed a processor patch than slotCapacity");
    if (e == 4) return fail(h, "corrupt migration slot header");
    if (e == 5) return fail(h, "timed out waiting for a neighbour's migration slot (peer-memory transfer)");
    if (e == 6) return fail(h, "pressure outlet: a face asked for more parcels than the insertion bound (outlet pressure at the initial outlet temperature, speed ratio 5) allows");
    if (e) return fail(h, "device error flag " + std::to_string(e));
    return 0;
}

// Wait for the pending read-back of the array length, if any.
int refresh_n(ugf_handle* h) {
    if (h->nPending) {
        CU(cudaEventSynchronize(h->evN));
        h->nUpper = std::min<long long>(h->capacity, *h->pinN + h->appendBound);
        h->nPending = false;
    }
    return 0;
}

// Non-blocking variant for drivers that keep the GPU queue full: take the exact length if its read-back has landed,
// otherwise fall back to the capacity as launch bound (kernels compare against the device-resident length anyway).
int refresh_n_lazy(ugf_handle* h, bool* exact) {
    *exact = true;
    if (h->nPending) {
        const cudaError_t q = cudaEventQuery(h->evN);
        if (q == cudaSuccess) {
            h->nUpper = std::min<long long>(h->capacity, *h->pinN + h->appendBound);
            h->nPending = false;
            *exact = (h->appendBound == 0);
        } else if (q == cudaErrorNotReady) {
            h->nUpper = h->capacity;
            *exact = false;
        } else {
            return fail(h, std::string("cudaEventQuery: ") + cudaGetErrorString(q));
        }
    }
    return 0;
}

int request_n(ugf_handle* h) {
    CU(countedMemcpyAsync(h, h->pinN, h->dN, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->evN, h->stream));
    h->nPending = true;
    h->appendBound = 0;  // the copy is ordered after everything appended so far
    return 0;
}

void build_params(ugf_handle* h) {
    DevParams& p = h->prm;
    std::memset(&p, 0, sizeof(p));
    p.seed = h->cfg.seed;
    p.nParticle = h->cfg.nParticle;
    p.deltaT = h->cfg.deltaT;
    p.Tref = h->cfg.Tref;
    p.theta = h->cfg.theta;
    p.invZrot = 1.0 / h->cfg.rotationalRelaxationCollisionNumber;
    p.invZel = 1.0 / h->cfg.electronicRelaxationCollisionNumber;
    for (int k = 0; k < 3; ++k) p.solD[k] = h->cfg.solutionD[k];
    p.collisionModel = h->cfg.collisionModel;
    p.binaryModel = h->cfg.binaryModel;
    p.bgkModel = h->cfg.bgkModel;
    p.nSpecies = h->nSpecies;
    p.measureWalls = h->cfg.measureWalls;
    p.cwf = h->dCwf[0] ? h->dCwf[h->cwfCur] : nullptr;
    p.cwfPrev = h->dCwf[0] ? h->dCwf[h->cwfCur ^ 1] : nullptr;
    p.cwfDirty = h->cwfDirty ? 1 : 0;
    p.spi = h->dSpi;
    p.axi = h->cfg.axisymmetric ? 1 : 0;
    p.rwfCentre = h->rwfCentre ? 1 : 0;
    p.rwfMaxM1 = h->cfg.maxRWF - 1.0;
    p.radialExtent = h->cfg.radialExtent;
    p.cellRwf = h->dCellRwf;
    p.bfRwf = h->dBfRwf;
    for (int i = 0; i < h->nSpecies; ++i) {
        const ugf_species& s = h->spHost[i];
        DevSpecies& d = p.sp[i];
        d.mass = s.mass; d.d = s.d; d.omega = s.omega; d.alpha = s.alpha; d.E0 = s.electronicEnergy[0];
        d.rotDoF = s.rotationalDoF; d.charge = s.charge; d.nElec = s.nElectronicLevels; d.g0 = s.degeneracy[0];
        d.vibDoF = s.vibrationalDoF; d.pad = 0;
    }
    for (int i = 0; i < h->nSpecies; ++i)
        for (int j = 0; j < h->nSpecies; ++j) {
            const double om = 0.5 * (h->spHost[i].omega + h->spHost[j].omega);
            p.pairInvGamma[i * UGF_MAX_SPECIES + j] = 1.0 / std::exp(std::lgamma(2.5 - om));
        }
}

int alloc_parcels(ugf_handle* h) {
    if (h->buf[0].x) return 0;
    const size_t cap = (size_t)h->capacity;
    const size_t capPad = (cap + MOVE_TILE - 1) / MOVE_TILE * MOVE_TILE;  // move_stream_kernel copies whole tiles
    for (int b = 0; b < 2; ++b) {
        ParcelBuf& P = h->buf[b];
        if (dalloc(h, &P.x, capPad) || dalloc(h, &P.y, capPad) || dalloc(h, &P.z, capPad) || dalloc(h, &P.ux, capPad) ||
            dalloc(h, &P.uy, capPad) || dalloc(h, &P.uz, capPad) || dalloc(h, &P.cell, capPad))
            return 1;
        if (h->hasRot && dalloc(h, &P.erot, capPad)) return 1;
        if (h->multi && dalloc(h, &P.type, capPad)) return 1;
        if (h->hasVib && dalloc(h, &P.vib, capPad)) return 1;
        if (h->hasElec && dalloc(h, &P.elev, capPad)) return 1;
    }
    if (dalloc(h, &h->dPerm, cap)) return 1;
    {
        const size_t warps = (size_t)h->numSMs * 4 * MOVE_WARPS;  // the largest grid the streamed move launches
        if (dalloc(h, &h->dMoveQD, warps * MoveSmem<false>::queueDoubles) || dalloc(h, &h->dMoveQI, warps * MoveSmem<false>::queueInts)) return 1;
    }
    if (dalloc(h, &h->dOwner, cap)) return 1;
    CU(cudaMemsetAsync(h->dOwner, 0x7f, cap * sizeof(int), h->stream));
    if (dalloc(h, &h->dMigBlock, (size_t)2 * MIG_MAXP * (cap / 1024 + 2))) return 1;  // per-(slot, block) counts and offsets
    return 0;
}

// cell occupancy: (histogram if needed) -> scan -> index scatter -> per-cell segment sort
int do_sort(ugf_handle* h) {
    if (refresh_n(h)) return 1;
    const int nC = h->nCells;
    if (!h->histValid) {
        if (!h->cellCountZero) CU(cudaMemsetAsync(h->dCellCount, 0, sizeof(int) * nC, h->stream));
        h->cellCountZero = false;
        hist_kernel<<<grid_for(h->nUpper, 256), 256, 0, h->stream>>>(h->buf[h->cur].cell, h->dN, h->dCellCount, h->cloneValid ? h->dNclone : nullptr);
        LAUNCHED();
    }
    const int nb = (nC + SCAN_TILE - 1) / SCAN_TILE;
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, h->stream>>>(h->dCellCount, nC, h->dBlockSums);
    LAUNCHED();
    h->argBytes += 20 + 56 + 56 + 32;  // the four occupancy kernels' argument blocks
    scan_final_self_kernel<<<nb, SCAN_THREADS, 0, h->stream>>>(h->dCellCount, nC, h->dBlockSums, nb, h->dTotal, h->dOff, h->capacity, h->dErr);
    LAUNCHED();
    scatter_index_kernel<<<grid_for(h->nUpper, 256 * SCAT_ROWS), 256, 0, h->stream>>>(h->buf[h->cur].cell, h->dN, h->dOff, h->dCellCount, h->dPerm,
                                                                                                    h->cloneValid ? h->dNclone : nullptr, h->capacity);
    LAUNCHED();
    segment_sort_kernel<<<grid_for(((long long)nC + SEG_CHUNK - 1) / SEG_CHUNK, SEG_THREADS / 32), SEG_THREADS, 0, h->stream>>>(h->dOff, nC, h->dPerm, h->dCellCount);
    LAUNCHED();
    h->cellCountZero = true;
    h->histValid = false;
    h->occValid = true;
    h->occIdentity = false;

    return 0;
}

// after a gather the array is cell-major and its length is the live count
int after_gather(ugf_handle* h, bool setN = true) {
    h->cur ^= 1;
    if (h->cloneValid) h->nUpper = h->capacity;  // the gather may have materialised clones: the array can be longer than the
                                                 // old bound until the length read-back below has landed
    h->cloneValid = false;  // the clones are parcels of their own now
    if (setN) {  // the cell kernel writes the new length itself
        set_n_kernel<<<1, 1, 0, h->stream>>>(h->dTotal, h->dN);
        LAUNCHED();
    }
    h->occIdentity = true;
    return request_n(h);
}

// streaming kernel: gather through the occupancy permutation (optional) + cell moments (optional)
// keepMoments = false (fused step without a BGK model): the 256-byte moment block per (cell, species) is consumed in
// registers by the field accumulation and not written to HBM
int run_cell_kernel(ugf_handle* h, bool gather, bool doSample, bool accumulate = false, bool keepMoments = true) {
    if (!gather && !doSample) return 0;
    CellArgs a{};
    a.nCells = h->nCells;
    a.off = h->dOff;
    a.perm = h->occIdentity ? nullptr : h->dPerm;
    a.in = h->buf[h->cur];
    a.out = gather ? h->buf[h->cur ^ 1] : h->buf[h->cur];
    a.gather = gather ? 1 : 0;
    a.doSample = doSample ? 1 : 0;
    a.writeMom = keepMoments ? 1 : 0;
    a.mom = h->dMom;
    a.acc = h->dAcc;
    a.accS = h->dAccS;
    a.accDt = (accumulate && doSample) ? h->cfg.deltaT : 0.0;
    a.taskCounter = h->dTask + h->taskSel;
    a.taskReset = h->dTask + (h->taskSel ^ 1);
    h->taskSel ^= 1;
    a.dN = gather ? h->dN : nullptr;
    a.total = h->dTotal;
    // a task should fill the staging buffer once: cells per task from the mean occupancy (UGF_CELL_TASK overrides)
    a.taskCells = h->cellTask > 0 ? h->cellTask
                                  : std::max(2, std::min(CELL_CHUNK, (int)((double)CELL_CAP * h->nCells / (double)std::max<long long>(h->nUpper, 1))));
    a.flags = h->cellFlags;
    const DevParams prm = h->prm;
    dispatch(h, [&](auto R, auto M) {
        cell_kernel<decltype(R)::value, decltype(M)::value><<<h->cellBlocks, CELL_THREADS, h->cellSmem, h->stream>>>(prm, a);
    });
    LAUNCHED();
    h->argBytes += arg_bytes(prm, a);
    if (doSample && prm.axi) {  // axisymmetricSimulation: the XnParticle sums weighted with every parcel's own RWF
        AxiMomArgs xa{};
        xa.nCells = h->nCells; xa.off = h->dOff;
        xa.P = gather ? a.out : a.in;
        xa.perm = gather ? nullptr : a.perm;
        xa.mom = keepMoments ? h->dMom : nullptr; xa.acc = h->dAcc; xa.accS = h->dAccS; xa.accDt = a.accDt;
        if (h->multi) axi_moments_kernel<true><<<grid_for(h->nCells, 8), 256, 0, h->stream>>>(prm, xa);
        else axi_moments_kernel<false><<<grid_for(h->nCells, 8), 256, 0, h->stream>>>(prm, xa);
        LAUNCHED();
    }
    if (doSample && h->dAccI) {  // vibrational / electronic sums of the same (pre-collision) state
        InternalArgs ia{};
        ia.nCells = h->nCells; ia.off = h->dOff; ia.perm = a.perm; ia.in = a.in;
        ia.momI = h->dMomI; ia.accI = h->dAccI; ia.mom = keepMoments ? h->dMom : nullptr; ia.accDt = a.accDt;
        internal_modes_kernel<<<grid_for(h->nCells, 8), 256, 0, h->stream>>>(prm, ia);
        LAUNCHED();
    }
    if (doSample) h->momValid = keepMoments;
    if (gather) return after_gather(h, false);
    return 0;
}

// NTC collisions in place on the cell-major buffer
int run_ntc_kernel(ugf_handle* h) {
    NtcArgs a{};
    a.nCells = h->nCells;
    a.off = h->dOff;
    a.P = h->buf[h->cur];
    a.vol = h->dVol;
    a.sigmaTcRMax = h->dSigma;
    a.collModelId = h->dCollId;
    a.owner = h->dOwner;
    a.sub = h->dSub;
    a.step = (uint32_t)h->step;
    a.cnt = h->dCnt;
    const DevParams prm = h->prm;
    // noTimeCounterSubCycled: the whole pass nSubCycles times with deltaT / nSubCycles (…SubCycled.C:86-190)
    const int nSub = h->cfg.partnerModel == UGF_PARTNER_NTC_SUBCYCLED ? std::max(1, h->cfg.nSubCycles) : 1;
    a.dtSub = nSub > 1 ? h->cfg.deltaT / nSub : h->cfg.deltaT;
    if (!h->subLevelsAllOne) {
        SubcellArgs sa{};
        sa.P = h->buf[h->cur]; sa.dN = h->dN; sa.bbMin = h->dBbMin; sa.bbMax = h->dBbMax; sa.levels = h->dSubLevels; sa.sub = h->dSub;
        subcell_index_kernel<<<grid_for(h->nUpper, 256), 256, 0, h->stream>>>(prm, sa);
        LAUNCHED();
    }
    for (int sub = 0; sub < nSub; ++sub) {
        a.sub_cycle = (uint32_t)sub;
        const bool subc = !h->subLevelsAllOne, internal = h->dSpi != nullptr;
        dispatch(h, [&](auto R, auto M) {
            constexpr bool r = decltype(R)::value, mm = decltype(M)::value;
            if (internal) {
                if (subc) ntc_kernel<r, mm, true, true><<<h->ntcBlocks, NTC_THREADS, 0, h->stream>>>(prm, a);
                else ntc_kernel<r, mm, false, true><<<h->ntcBlocks, NTC_THREADS, 0, h->stream>>>(prm, a);
            } else {
                if (subc) ntc_kernel<r, mm, true><<<h->ntcBlocks, NTC_THREADS, 0, h->stream>>>(prm, a);
                else ntc_kernel<r, mm, false><<<h->ntcBlocks, NTC_THREADS, 0, h->stream>>>(prm, a);
            }
        });
        LAUNCHED();
        h->argBytes += arg_bytes(prm, a);
    }
    return 0;
}

bool dsmc_active(const ugf_handle* h) {
    return (h->cfg.collisionModel == UGF_COLL_DSMC || h->cfg.collisionModel == UGF_COLL_HYBRID) && h->cfg.binaryModel != UGF_BINARY_NONE;
}
bool bgk_active(const ugf_handle* h) {
    return (h->cfg.collisionModel == UGF_COLL_BGK || h->cfg.collisionModel == UGF_COLL_HYBRID) && h->cfg.bgkModel != UGF_BGK_NONE;
}

int run_bgk_kernel(ugf_handle* h) {
    BgkArgs a{};
    a.nCells = h->nCells;
    a.off = h->dOff;
    a.P = h->buf[h->cur];
    a.mom = h->dMom;
    a.vol = h->dVol;
    a.collModelId = h->dCollId;
    a.maxProb = h->dMaxProb;
    a.qPrev = h->dQPrev;
    a.sPrev = h->dSPrev;
    a.keyScratch = h->dKeyScratch;
    a.step = (uint32_t)h->step;
    a.cnt = h->dCnt;
    const DevParams prm = h->prm;
    if (h->cfg.macroInterpolation && !h->interpSet) return fail(h, "macroInterpolation true needs ugf_set_macro_interpolation");
    if (!h->dMacro && dalloc(h, &h->dMacro, (size_t)h->nCells)) return 1;
    a.macroCell = h->dMacro;
    if (h->cfg.macroInterpolation) {
        a.ip = h->interp;
        bgk_fields_kernel<<<grid_for(h->nCells, 128), 128, 0, h->stream>>>(prm, a);
        LAUNCHED();
        bgk_points_kernel<<<grid_for(h->interp.nPoints, 128), 128, 0, h->stream>>>(h->interp);
        LAUNCHED();
        if (h->multi) bgk_kernel<true, true><<<h->bgkBlocks, BGK_THREADS, h->bgkSmem, h->stream>>>(prm, a);
        else bgk_kernel<false, true><<<h->bgkBlocks, BGK_THREADS, h->bgkSmem, h->stream>>>(prm, a);
        LAUNCHED();
        h->argBytes += 3 * arg_bytes(prm, a);
        return 0;
    }
    bgk_fields_kernel<<<grid_for(h->nCells, 128), 128, 0, h->stream>>>(prm, a);  // a.ip.cellF == null: the cells' state only
    LAUNCHED();
    if (h->multi) bgk_kernel<true><<<h->bgkBlocks, BGK_THREADS, h->bgkSmem, h->stream>>>(prm, a);
    else bgk_kernel<false><<<h->bgkBlocks, BGK_THREADS, h->bgkSmem, h->stream>>>(prm, a);
    LAUNCHED();
    h->argBytes += 2 * arg_bytes(prm, a);
    return 0;
}

int do_inflow(ugf_handle* h) {
    if (refresh_n(h)) return 1;
    h->newFrom = h->nUpper;  // exact: refresh_n just ran or no gather happened since
    const DevParams prm = h->prm;
    for (InflowHost& f : h->inflows) {
        if (h->nUpper + f.maxInsert > h->capacity) return fail(h, "parcelCapacity too small for the inflow patches");
        inflow_count_kernel<<<grid_for(f.nSlots, 256), 256, 0, h->stream>>>(prm, f.dev, (uint32_t)h->step);
        LAUNCHED();
        inflow_scan_kernel<<<1, SCAN_THREADS, 0, h->stream>>>(f.dev.nIns, f.nSlots, f.dev.insOff, h->dN, h->capacity, h->dCnt, h->dErr);
        LAUNCHED();
        const InflowDev dev = f.dev;
        ParcelBuf P = h->buf[h->cur];
        const uint32_t step = (uint32_t)h->step;
        dispatch(h, [&](auto R, auto M) {
            inflow_insert_kernel<decltype(R)::value, decltype(M)::value><<<grid_for(f.nSlots, 128), 128, 0, h->stream>>>(prm, dev, P, step);
        });
        LAUNCHED();
        h->argBytes += 3 * arg_bytes(prm, dev) + 64;
        h->nUpper += f.maxInsert;
        h->appendBound += f.maxInsert;
    }
    h->inflowDone = true;
    h->histValid = false; h->occValid = false; h->momValid = false;
    return 0;
}

MoveArgs make_move_args(ugf_handle* h, long long begin, bool received);
int launch_move(ugf_handle* h, const MoveArgs& a, long long count, long long begin, bool received);

int do_move(ugf_handle* h, long long begin, bool received) {
    for (int p = 0; p < h->nPatches; ++p)
        if (h->patchKind[p] == UGF_PATCH_WALL && h->patchesHost[p].wallModel == UGF_WALL_UNSET)
            return fail(h, "wall patch without a boundary model");  // uniGasBoundaries.C:448-488
    if (!received) {
        if (!h->cellCountZero) CU(cudaMemsetAsync(h->dCellCount, 0, sizeof(int) * h->nCells, h->stream));
        if (h->hasProcessor) CU(cudaMemsetAsync(h->dMigCount, 0, sizeof(int) * std::max(h->nPatches, 1), h->stream));
    }
    h->cellCountZero = false;
    if (h->hasProcessor && !(received && h->slotRound)) CU(cudaMemsetAsync(h->dInflight, 0, sizeof(unsigned long long), h->stream));
    MoveArgs a = make_move_args(h, begin, received);
    long long count = h->nUpper - begin;
    if (a.dBegin) count = std::min<long long>(count, (long long)h->migSlots.nProc * h->lastSlotCapacity);  // what one unpack can append
    return launch_move(h, a, count, begin, received);
}

MoveArgs make_move_args(ugf_handle* h, long long begin, bool received) {
    MoveArgs a{};
    a.mesh = h->mesh;
    a.P = h->buf[h->cur];
    a.dN = h->dN;
    a.begin = begin;
    a.newFrom = received ? (1LL << 62) : h->newFrom;
    a.sf = h->dSf;
    a.wq = h->dWq;
    a.useSfIn = received ? 1 : 0;
    a.step = (uint32_t)h->step;
    a.aux = received ? 1u : 0u;
    a.cellCount = h->dCellCount;
    a.migCount = h->dMigCount;
    a.inflight = h->dInflight;
    a.dBegin = (received && h->slotRound) ? h->dRecvStart : nullptr;
    a.ms = h->migSlots;
    a.migList = (h->migSlots.nProc <= MIG_MAXP) ? h->dMigList : nullptr;
    a.migListCap = MIG_LIST_CAP;
    a.bm = h->dBm;
    a.cnt = h->dCnt;
    a.nclone = h->prm.cwf ? h->dNclone : nullptr;
    a.queueD = h->dMoveQD;
    a.queueI = h->dMoveQI;
    a.slotTrack = h->nTracked ? h->dSlotTrack : nullptr;
    a.bfTrack = h->dBfTrack;
    a.ft = h->dFt;
    return a;
}

int launch_move(ugf_handle* h, const MoveArgs& a, long long count, long long begin, bool received) {
    if (count > 0) {
        const DevParams prm = h->prm;
        const bool streamed = !received && begin == 0 && !h->moveDirect;
        const unsigned grid = streamed ? (unsigned)std::min<long long>(grid_for(count, MOVE_WARPS * 32), (long long)h->numSMs * 4)
                                       : grid_for(count, 256);
        dispatch(h, [&](auto R, auto M) {
            constexpr bool r = decltype(R)::value, mm = decltype(M)::value;
            if (streamed && h->moveV2) {
                // hop-compacting kernel: the packed 2-D record only where z is an empty direction (z, Uz are then not staged)
                const bool flat = h->moveNF == 4 && h->mesh.rec2d && !h->cfg.solutionD[2];
#define UGF_MOVE_STREAM2(NF_, BPS_, FLAT_)                                                                                   \
    do {                                                                                                                      \
        auto kfn = move_stream2_kernel<r, mm, NF_, BPS_>;                                                                      \
        static bool attrSet = false;                                                                                          \
        if (!attrSet) { cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MoveSmem<FLAT_>::total); attrSet = true; } \
        const unsigned g2 = (unsigned)std::min<long long>(grid_for(count, MOVE_WARPS * 32), (long long)h->numSMs * BPS_);        \
        kfn<<<g2, MOVE_WARPS * 32, MoveSmem<FLAT_>::total, h->stream>>>(prm, a);                                               \
    } while (0)
                if (flat && h->moveBps == 3) UGF_MOVE_STREAM2(NF_REC2D, 3, true);
                else if (flat) UGF_MOVE_STREAM2(NF_REC2D, 4, true);
                else if (h->moveNF == 6 && h->moveBps != 3) UGF_MOVE_STREAM2(6, 2, false);  // 126 registers, no spills: measured 10 % faster than 3 CTAs / SM at 80
                else if (h->moveNF == 6) UGF_MOVE_STREAM2(6, 3, false);
                else if (h->moveNF == 4) UGF_MOVE_STREAM2(4, 3, false);
                else UGF_MOVE_STREAM2(0, 3, false);
#undef UGF_MOVE_STREAM2
            } else if (streamed) {
#define UGF_MOVE_STREAM(NF_)                                                                                      \
    do {                                                                                                          \
        move_stream_kernel<r, mm, NF_, 4><<<grid, MOVE_WARPS * 32, 0, h->stream>>>(prm, a);                        \
    } while (0)
                if (h->moveNF == 6) UGF_MOVE_STREAM(6);
                else if (h->moveNF == 4 && h->mesh.rec2d) UGF_MOVE_STREAM(NF_REC2D);
                else if (h->moveNF == 4) UGF_MOVE_STREAM(4);
                else UGF_MOVE_STREAM(0);
#undef UGF_MOVE_STREAM
            } else {
                if (h->moveNF == 6) move_kernel<r, mm, 6><<<grid, 256, 0, h->stream>>>(prm, a);
                else if (h->moveNF == 4 && h->mesh.rec2d) move_kernel<r, mm, NF_REC2D><<<grid, 256, 0, h->stream>>>(prm, a);
                else if (h->moveNF == 4) move_kernel<r, mm, 4><<<grid, 256, 0, h->stream>>>(prm, a);
                else move_kernel<r, mm, 0><<<grid, 256, 0, h->stream>>>(prm, a);
            }
        });
        LAUNCHED();
        h->argBytes += arg_bytes(prm, a);
    }
    h->histValid = true;
    h->occValid = false;
    h->momValid = false;
    if (h->rwfCentre) { h->rwfCentre = false; h->prm.rwfCentre = 0; }  // the weighting pass that rode on this move left RWF(position) on every parcel
    if (h->prm.cwf) {
        h->cloneValid = true;
        if (h->cwfDirty) {  // every parcel now carries the new factors
            CU(countedMemcpyAsync(h, h->dCwf[h->cwfCur ^ 1], h->dCwf[h->cwfCur], sizeof(double) * h->nCells, cudaMemcpyDeviceToDevice, h->stream));
            h->cwfHostPrev = h->cwfHost;
            h->cwfDirty = false;
            h->prm.cwfDirty = 0;
        }
    }
    return 0;
}

// true if the next do_accumulate() call falls on a sampling step (fieldPropertiesDict sampleInterval)
bool next_step_samples(const ugf_handle* h) {
    const int interval = h->cfg.sampleInterval > 0 ? h->cfg.sampleInterval : 1;
    return interval <= h->sampleCounter + 1;
}

// cellsDone: the streaming kernel already added this step's cell sums to the accumulators
int do_accumulate(ugf_handle* h, bool cellsDone = false) {
    h->sampleCounter++;
    const int interval = h->cfg.sampleInterval > 0 ? h->cfg.sampleInterval : 1;
    int accumulate = 0;
    if (interval <= h->sampleCounter) {
        h->nAvTimeSteps++;
        h->timeAvCounter += h->cfg.deltaT;
        accumulate = 1;
        if (!cellsDone) {
            accumulate_cells_kernel<<<grid_for(h->nCells, 256), 256, 0, h->stream>>>(h->prm, h->nCells, h->dMom, h->dAcc, h->dAccS);
            LAUNCHED();
            if (h->dAccI) {
                const long long nI = (long long)h->nCells * h->nSpecies * UGF_NINT;
                accumulate_internal_kernel<<<grid_for(nI, 256), 256, 0, h->stream>>>(nI, h->cfg.deltaT, h->dMomI, h->dAccI);
                LAUNCHED();
            }
        }
        h->sampleCounter = 0;
    }
    if (h->cfg.measureWalls) {
        // only wall patches carry boundary measurements; merge adjacent wall patches into one launch
        int p = 0;
        while (p < h->nPatches) {
            if (h->patchKind[p] != UGF_PATCH_WALL || h->patchSize[p] == 0) { ++p; continue; }
            const long long first = h->patchesHost[p].startBfi;
            long long last = first + h->patchSize[p];
            int q = p + 1;
            while (q < h->nPatches && h->patchKind[q] == UGF_PATCH_WALL && h->patchesHost[q].startBfi == last) { last += h->patchSize[q]; ++q; }
            const long long cnt = (last - first) * UGF_NBM;
            accumulate_walls_kernel<<<grid_for(cnt, 256), 256, 0, h->stream>>>(h->cfg.deltaT, accumulate, cnt, h->dBm + first * UGF_NBM, h->dBacc + first * UGF_NBM);
            LAUNCHED();
            h->argBytes += 40;
            p = q;
        }
    }
    return 0;
}

// device-side pack of every processor patch into the given slot addresses: from the migrant lists of the move
// kernel when they can hold a full slot, else by two passes over the cell ids
int pack_slots_to(ugf_handle* h, const MigDst& dst, long long slotCapacity, const MigFlags* fl = nullptr, unsigned long long epoch = 0, bool* signalled = nullptr) {
    if (signalled) *signalled = false;
    if (h->dMigList && slotCapacity <= MIG_LIST_CAP && !h->migSearchPack) {
        ParcelBuf P = h->buf[h->cur];
        const MigSlots ms = h->migSlots;
        dispatch(h, [&](auto R, auto M) {
            mig_pack_list_kernel<decltype(R)::value, decltype(M)::value><<<ms.nProc, 1024, 0, h->stream>>>(h->mesh, ms, P, h->dSf, h->dWq, h->dMigCount, h->dMigList, MIG_LIST_CAP,
                                                                                                          dst, slotCapacity, h->dErr, fl ? *fl : MigFlags{}, fl ? epoch : 0ull,
                                                                                                          fl ? h->dInflight : nullptr);
        });
        LAUNCHED();
        if (signalled && fl) *signalled = true;
        return 0;
    }
    const int nb = (int)grid_for(h->nUpper, MIG_TILE);
    ParcelBuf P = h->buf[h->cur];
    const MigSlots ms = h->migSlots;
    int* counts = h->dMigBlock;
    int* offsets = h->dMigBlock + (size_t)MIG_MAXP * (h->capacity / 1024 + 2);
    mig_count_all_kernel<<<nb, MIG_THREADS, 0, h->stream>>>(h->mesh, ms, P.cell, h->dN, h->slotRound ? h->dRecvStart : nullptr, counts, nb);
    LAUNCHED();
    mig_scan_kernel<<<ms.nProc, SCAN_THREADS, 0, h->stream>>>(counts, offsets, nb, h->dMigTotals);
    LAUNCHED();
    dispatch(h, [&](auto R, auto M) {
        mig_pack_all_kernel<decltype(R)::value, decltype(M)::value><<<nb, MIG_THREADS, 0, h->stream>>>(h->mesh, ms, P, h->dSf, h->dWq, h->dN, counts, offsets, h->dMigTotals, nb,
                                                                                                        dst, slotCapacity, h->dErr);
    });
    LAUNCHED();
    CU(cudaMemsetAsync(h->dMigCount, 0, sizeof(int) * std::max(h->nPatches, 1), h->stream));
    return 0;
}

// fetchCellNeighborhood (uniGasHybridDecomposition.C:127-174): the cell and everything within nLevels face hops
void cell_neighbourhood(const ugf_handle* h, int cell, int nLevels, std::vector<int>& nb, std::vector<char>& mark) {
    nb.clear();
    nb.push_back(cell);
    mark[cell] = 1;
    size_t first = 0;
    for (int level = 1; level <= nLevels; ++level) {
        const size_t last = nb.size();
        for (size_t i = first; i < last; ++i)
            for (int j = h->ccOffHost[nb[i]]; j < h->ccOffHost[nb[i] + 1]; ++j) {
                const int q = h->ccIdsHost[j];
                if (!mark[q]) { mark[q] = 1; nb.push_back(q); }
            }
        first = last;
    }
    for (int q : nb) mark[q] = 0;
}

// The refinement sweeps of localKnudsen::decompose (localKnudsen.C:428-541) update the mask in place while walking
// the cells in index order, so their result depends on that order: host code, on the downloaded mask.
void refine_mask(const ugf_handle* h, std::vector<int32_t>& id) {
    std::vector<int> nb;
    std::vector<char> mark((size_t)h->nCells, 0);
    for (int pass = 1; pass <= h->dec.refinementPasses; ++pass)
        for (int which = 1; which >= 0; --which)  // dsmc cells first, then bgk cells
            for (int c = 0; c < h->nCells; ++c) {
                if (id[c] != which) continue;
                int same = 0, other = 0;
                for (int j = h->ccOffHost[c]; j < h->ccOffHost[c + 1]; ++j) (id[h->ccIdsHost[j]] == which ? same : other)++;
                if (same == 0 || (same == 1 && other > 1)) { id[c] = 1 - which; continue; }
                cell_neighbourhood(h, c, h->dec.neighborLevels, nb, mark);
                int nSame = 0;
                for (int q : nb) nSame += (id[q] == which);
                if (nSame < h->dec.maxNeighborFraction * nb.size()) id[c] = 1 - which;
            }
}

// uniGasCloud::decomposition (U/clouds/uniGasCloud.C:250-256): accumulate every step, re-decompose every interval
int do_decompose(ugf_handle* h) {
    if (!h->decompOn) return 0;
    if (!h->momValid) return fail(h, "ugf_decompose needs the step's cell moments (ugf_sample / ugf_collide / ugf_relax first)");
    const int nC = h->nCells, W = KN_NACC + h->nSpecies;
    const unsigned grid = grid_for(nC, 256);
    const DevParams prm = h->prm;
    kn_accumulate_kernel<<<grid, 256, 0, h->stream>>>(prm, nC, h->dMom, h->dKnAcc);
    LAUNCHED();
    h->decTimeSteps++;
    h->decTimeAv += h->cfg.deltaT;
    if (h->decTimeSteps != h->dec.decompositionInterval) return 0;
    const DecompGeom g = h->dgeom;
    int cur = 0;
    kn_derive_kernel<<<grid, 256, 0, h->stream>>>(nC, W, h->dKnAcc, h->dVol, h->decTimeAv, h->dKnF[cur]);
    LAUNCHED();
    for (int pass = 1; pass <= h->dec.smoothingPasses; ++pass) {  // rhoM, p, T, U; rhoN is not smoothed (:283-292)
        kn_smooth_kernel<KN_NF, 4><<<grid, 256, 0, h->stream>>>(g, 1, h->dKnF[cur], h->dKnF[cur ^ 1]);
        LAUNCHED();
        cur ^= 1;
    }
    kn_kernel<<<grid, 256, 0, h->stream>>>(prm, g, W, h->dKnAcc, h->dKnF[cur], h->dec.breakdownMax, h->dec.theta, h->dKnK[h->knCur]);
    LAUNCHED();
    for (int pass = 1; pass <= h->dec.smoothingPasses; ++pass) {
        kn_smooth_kernel<4, -1><<<grid, 256, 0, h->stream>>>(g, 0, h->dKnK[h->knCur], h->dKnK[h->knCur ^ 1]);
        LAUNCHED();
        h->knCur ^= 1;
    }
    kn_threshold_kernel<<<grid, 256, 0, h->stream>>>(nC, h->dKnK[h->knCur], h->dec.breakdownMax, h->dCollId);
    LAUNCHED();
    if (h->dec.refinementPasses > 0) {
        std::vector<int32_t> id((size_t)nC);
        CU(countedMemcpyAsync(h, id.data(), h->dCollId, sizeof(int32_t) * (size_t)nC, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        refine_mask(h, id);
        if (upload(h, h->dCollId, id.data(), (size_t)nC)) return 1;
        CU(cudaStreamSynchronize(h->stream));
    }
    h->decTimeSteps = 0;
    const double timeNow = (double)(h->step + 1) * h->cfg.deltaT;
    if (h->dec.resetAtDecomposition && timeNow < h->dec.resetAtDecompositionUntilTime + 0.5 * h->cfg.deltaT) {
        h->decTimeAv = 0.0;
        CU(cudaMemsetAsync(h->dKnAcc, 0, sizeof(double) * (size_t)nC * W, h->stream));
    }
    return 0;
}

int zero_step_counters(ugf_handle* h) {
    CU(cudaMemsetAsync(h->dCnt, 0, sizeof(DevCounters), h->stream));
    return 0;
}

}  // namespace

// =========================================================================================================
extern "C" {

int ugf_abi_version(void) { return UGF_ABI_VERSION; }

const char* ugf_last_error(const ugf_handle* h) { return h ? h->err.c_str() : g_createErr.c_str(); }

int ugf_create(const ugf_config* cfg, ugf_handle** out) {
    ugf_handle* h = nullptr;
    if (!cfg || !out) return fail(nullptr, "null argument");
    if (cfg->abiVersion != UGF_ABI_VERSION) return fail(nullptr, "ABI version mismatch");
    if (cfg->partnerModel != UGF_PARTNER_NTC && cfg->partnerModel != UGF_PARTNER_NTC_SUBCYCLED) return fail(nullptr, "unknown dsmcCollisionPartnerModel");
    if (cfg->partnerModel == UGF_PARTNER_NTC_SUBCYCLED && cfg->nSubCycles < 1) return fail(nullptr, "noTimeCounterSubCycled needs nSubCycles >= 1");
    if (cfg->parcelCapacity <= 0 || cfg->parcelCapacity > 2000000000LL) return fail(nullptr, "parcelCapacity out of range");
    if (cfg->axisymmetric && !(cfg->radialExtent > 0.0 && cfg->maxRWF >= 1.0))
        return fail(nullptr, "axisymmetricSimulation needs radialExtentOfDomain > 0 and maxRadialWeightingFactor >= 1");
    if (cfg->axisymmetric && !(cfg->solutionD[0] && cfg->solutionD[1] && cfg->solutionD[2]))
        return fail(nullptr, "axisymmetricSimulation runs on a wedge mesh with symmetryPlane sides: all three directions are solved");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libugf has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "device ordinal out of range");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    h = new ugf_handle();
    h->cfg = *cfg;
    h->capacity = cfg->parcelCapacity;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->numSMs = prop.multiProcessorCount;
    auto bail = [&](const char* what, cudaError_t err) {
        g_createErr = std::string(what) + ": " + cudaGetErrorString(err);
        delete h;
        return 1;
    };
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->evN, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i < UGF_NPHASE + 1; ++i)
        if ((e = cudaEventCreate(&h->ev[i])) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMallocHost((void**)&h->pinN, sizeof(long long))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMalloc((void**)&h->dN, sizeof(long long))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dCnt, sizeof(DevCounters))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dErr, sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dTask, 2 * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemsetAsync(h->dTask, 0, 2 * sizeof(int), h->stream);
    if ((e = cudaMalloc((void**)&h->dTot, 8 * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dTotal, sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dMigTotals, MIG_MAXP * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dInflight, sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dRecvStart, sizeof(long long))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->dMigDone, sizeof(unsigned int))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemsetAsync(h->dMigDone, 0, sizeof(unsigned int), h->stream);
    if (const char* ev = std::getenv("UGF_MIG_FUSED")) h->migFused = std::atoi(ev) != 0;
    cudaMemsetAsync(h->dInflight, 0, sizeof(unsigned long long), h->stream);
    cudaMemsetAsync(h->dRecvStart, 0, sizeof(long long), h->stream);
    cudaMemsetAsync(h->dN, 0, sizeof(long long), h->stream);
    cudaMemsetAsync(h->dCnt, 0, sizeof(DevCounters), h->stream);
    cudaMemsetAsync(h->dErr, 0, sizeof(int), h->stream);
    cudaMemsetAsync(h->dTotal, 0, sizeof(int), h->stream);
    *out = h;
    return 0;
}

int ugf_destroy(ugf_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    for (int b = 0; b < 2; ++b) {
        ParcelBuf& P = h->buf[b];
        cudaFree(P.x); cudaFree(P.y); cudaFree(P.z); cudaFree(P.ux); cudaFree(P.uy); cudaFree(P.uz);
        cudaFree(P.erot); cudaFree(P.cell); cudaFree(P.type); cudaFree(P.vib); cudaFree(P.elev);
    }
    void* ptrs[] = {h->dCfOff, h->dPlane, h->dNbr, h->dBfPatch, h->dBfOwner, h->dPatches, h->dVol, h->dBbMin, h->dBbMax, h->dBfS,
                    h->dSf, h->dN, h->dCellCount, h->dOff, h->dPerm, h->dBlockSums, h->dTotal, h->dMigCount, h->dMigBlock, h->dMigTotals, h->dMigList, h->dInflight, h->dRecvStart, h->dMigDone,
                    h->dMom, h->dAcc, h->dAccS, h->dBm, h->dBacc, h->dSigma, h->dCollId, h->dMaxProb, h->dQPrev, h->dSPrev, h->dKeyScratch, h->dOwner, h->dSubLevels, h->dSub,
                    h->dCnt, h->dErr, h->dTot, h->dTask, h->dCwf[0], h->dCwf[1], h->dNclone, h->dWq, h->dRec2d, h->dMoveQD, h->dMoveQI, h->dSpi, h->dMomI, h->dAccI, h->dSlotTrack, h->dBfTrack, h->dFt, h->dMacro, h->dCellRwf, h->dBfRwf};
    for (void* p : ptrs) cudaFree(p);
    for (InflowHost& f : h->inflows) for (void* p : f.owned) cudaFree(p);
    for (double* p : h->packBuf) cudaFree(p);
    for (double* p : h->wallFieldOwned) cudaFree(p);
    for (void* p : h->decompOwned) cudaFree(p);
    for (void* p : h->interpOwned) cudaFree(p);
    for (void* p : h->peerOpened) cudaIpcCloseMemHandle(p);
    for (void* p : h->peerOwned) cudaFree(p);
    for (void* p : h->hostOwned) cudaFreeHost(p);
    if (h->pinN) cudaFreeHost(h->pinN);
    if (h->evN) cudaEventDestroy(h->evN);
    for (int i = 0; i < UGF_NPHASE + 1; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int ugf_set_species(ugf_handle* h, int32_t n, const ugf_species* sp) {
    if (!h) return 1;
    if (n < 1 || n > UGF_MAX_SPECIES) return fail(h, "species count out of range");
    if (h->buf[0].x) return fail(h, "ugf_set_species must precede ugf_upload_parcels");
    h->hasRot = false;
    for (int i = 0; i < n; ++i) {
        if (sp[i].vibrationalDoF < 0 || sp[i].vibrationalDoF > UGF_MAX_VIB_MODES) return fail(h, "bad vibrationalDoF");
        if (sp[i].nElectronicLevels < 1 || sp[i].nElectronicLevels > UGF_MAX_ELEC_LEVELS) return fail(h, "bad nElectronicLevels");
        for (int m = 0; m < sp[i].vibrationalDoF; ++m)
            if (!(sp[i].thetaV[m] > 0) || !(sp[i].thetaD[m] > 0) || !(sp[i].Zref[m] > 0) || !(sp[i].TrefZv[m] > 0))
                return fail(h, "vibrational mode needs positive characteristicVibrationalTemperature, dissociationTemperature, Zref and referenceTempForZref");
        if (!(sp[i].mass > 0) || !(sp[i].d > 0)) return fail(h, "species mass and diameter must be positive");
        h->spHost[i] = sp[i];
        if (sp[i].rotationalDoF > 0) h->hasRot = true;
        if (sp[i].vibrationalDoF > 0) h->hasVib = true;
        if (sp[i].nElectronicLevels > 1) h->hasElec = true;
    }
    h->nSpecies = n;
    h->multi = n > 1;
    if (h->hasVib || h->hasElec) {  // tables of the internal modes on the device (ugf_internal.cuh)
        CU(cudaSetDevice(h->cfg.device));
        std::vector<DevSpeciesInt> tab((size_t)n);
        for (int i = 0; i < n; ++i) {
            std::memset(&tab[i], 0, sizeof(DevSpeciesInt));
            for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) {
                tab[i].thetaV[m] = sp[i].thetaV[m]; tab[i].thetaD[m] = sp[i].thetaD[m]; tab[i].Zref[m] = sp[i].Zref[m]; tab[i].TrefZv[m] = sp[i].TrefZv[m];
            }
            for (int k = 0; k < UGF_MAX_ELEC_LEVELS; ++k) { tab[i].elecE[k] = sp[i].electronicEnergy[k]; tab[i].g[k] = sp[i].degeneracy[k]; }
        }
        if (!h->dSpi && dalloc(h, &h->dSpi, (size_t)UGF_MAX_SPECIES)) return 1;
        if (upload(h, h->dSpi, tab.data(), tab.size())) return 1;
        CU(cudaStreamSynchronize(h->stream));
    }
    build_params(h);
    return 0;
}

int ugf_set_mesh(ugf_handle* h, const ugf_mesh* m) {
    if (!h || !m) return 1;
    if (h->nSpecies == 0) return fail(h, "ugf_set_species must precede ugf_set_mesh");
    if (h->meshSet) return fail(h, "mesh already set");
    CU(cudaSetDevice(h->cfg.device));
    h->nCells = m->nCells; h->nFaces = m->nFaces; h->nInternal = m->nInternalFaces; h->nPatches = m->nPatches;
    h->nBFaces = m->nFaces - m->nInternalFaces;
    const int nC = h->nCells, nB = h->nBFaces, nI = h->nInternal;
    if (nC <= 0 || nB < 0) return fail(h, "bad mesh sizes");
    // flatten: per (cell, local face) an outward plane and the cell behind it.  Faces whose area vector has no
    // component in any solved direction (the front/back faces of a 2-D case) can never be crossed - the tracking
    // displacement is zero along empty directions (U/parcels/uniGasParcel.C:64-71) - and are left out.
    std::vector<double4> plane;
    h->slotFaceHost.clear();
    std::vector<int> nbr, cfOffDev((size_t)nC + 1, 0);
    plane.reserve(m->cellFaceOffsets[nC]);
    nbr.reserve(m->cellFaceOffsets[nC]);
    int uniformNF = -1, planeNoZ = 1;
    for (int c = 0; c < nC; ++c) {
        for (int j = m->cellFaceOffsets[c]; j < m->cellFaceOffsets[c + 1]; ++j) {
            const int f = m->cellFaces[j];
            if (f < 0 || f >= m->nFaces) return fail(h, "cellFaces entry out of range");
            const bool own = m->owner[f] == c;
            if (!own && !(f < nI && m->neighbour[f] == c)) return fail(h, "cellFaces lists a face that does not touch the cell");
            const double* S = m->faceAreas + 3 * (size_t)f;
            const double* C = m->faceCentres + 3 * (size_t)f;
            bool reachable = false;
            for (int k = 0; k < 3; ++k) if (h->cfg.solutionD[k] && S[k] != 0.0) reachable = true;
            if (!reachable) continue;
            const double d = S[0] * C[0] + S[1] * C[1] + S[2] * C[2];
            double4 pl;
            // storage order {Sx, Sy, S.Cf, Sz}: see load_plane / load_plane_noz
            if (own) { pl.x = S[0]; pl.y = S[1]; pl.z = d; pl.w = S[2]; }
            else { pl.x = -S[0]; pl.y = -S[1]; pl.z = -d; pl.w = -S[2]; }
            if (S[2] != 0.0) planeNoZ = 0;
            plane.push_back(pl);
            h->slotFaceHost.push_back(f);
            nbr.push_back(f < nI ? (own ? m->neighbour[f] : m->owner[f]) : -(f - nI + 1));
        }
        cfOffDev[c + 1] = (int)plane.size();
        const int cnt = cfOffDev[c + 1] - cfOffDev[c];
        if (c == 0) uniformNF = cnt; else if (cnt != uniformNF) uniformNF = 0;
    }
    // Cells with fewer reachable faces than their neighbours (wedges on the axis of a revolved block, prisms, cells that lost
    // a collapsed face) are padded with null planes {0,0,0,0}: nd = 0 for every displacement, so such a slot is never
    // selected, and the mesh keeps the unrolled 4- / 6-slot tracking loop.  UGF_MOVE_NF0=1 forces the general CSR walk.
    const bool forceCsr = std::getenv("UGF_MOVE_NF0") && std::atoi(std::getenv("UGF_MOVE_NF0")) != 0;
    int maxCnt = 0;
    for (int c = 0; c < nC; ++c) maxCnt = std::max(maxCnt, cfOffDev[c + 1] - cfOffDev[c]);
    if (!forceCsr && uniformNF != 4 && uniformNF != 6 && maxCnt <= 6 && maxCnt >= 1) {
        const int target = maxCnt <= 4 ? 4 : 6;
        std::vector<double4> plane2((size_t)nC * target, double4{0.0, 0.0, 0.0, 0.0});
        std::vector<int> nbr2((size_t)nC * target), face2((size_t)nC * target, -1);
        for (int c = 0; c < nC; ++c) {
            const int b = cfOffDev[c], n = cfOffDev[c + 1] - b;
            for (int k = 0; k < target; ++k) {
                const size_t d = (size_t)c * target + k;
                if (k < n) { plane2[d] = plane[b + k]; nbr2[d] = nbr[b + k]; face2[d] = h->slotFaceHost[b + k]; }
                else nbr2[d] = c;
            }
        }
        for (int c = 0; c <= nC; ++c) cfOffDev[c] = c * target;
        plane.swap(plane2); nbr.swap(nbr2); h->slotFaceHost.swap(face2);
        uniformNF = target;
    }
    if (forceCsr) uniformNF = 0;
    h->slotOffHost = cfOffDev;
    h->nSlots = (int)plane.size();
    h->moveNF = (uniformNF == 4 || uniformNF == 6) ? uniformNF : 0;
    std::vector<int> bfPatch(std::max(nB, 1), -1), bfOwner(std::max(nB, 1), 0);
    std::vector<double> bfS(3 * (size_t)std::max(nB, 1), 0.0);
    h->patchesHost.assign(h->nPatches, DevPatch{});
    h->patchKind.assign(m->patchKind, m->patchKind + h->nPatches);
    h->patchStart.assign(m->patchStart, m->patchStart + h->nPatches);
    h->patchSize.assign(m->patchSize, m->patchSize + h->nPatches);
    for (int p = 0; p < h->nPatches; ++p) {
        DevPatch& d = h->patchesHost[p];
        d.kind = m->patchKind[p];
        d.startBfi = m->patchStart[p] - nI;
        d.size = m->patchSize[p];
        d.partnerStartBfi = -1;
        d.wallModel = UGF_WALL_UNSET;
        for (int k = 0; k < 3; ++k) d.sep[k] = m->patchSeparation[3 * p + k];
        d.diffuseFraction = 1.0;
        if (d.startBfi < 0 || d.startBfi + d.size > nB) return fail(h, "patch range outside the boundary faces");
        if (d.kind == UGF_PATCH_CYCLIC) {
            const int q = m->patchPartner[p];
            if (q < 0 || q >= h->nPatches || m->patchKind[q] != UGF_PATCH_CYCLIC || m->patchSize[q] != d.size)
                return fail(h, "cyclic patch without a matching partner");
            d.partnerStartBfi = m->patchStart[q] - nI;
        }
        if (d.kind == UGF_PATCH_PROCESSOR) {
            h->hasProcessor = true;
            if (h->migSlots.nProc < MIG_MAXP) h->migSlots.patch[h->migSlots.nProc] = p;
            h->migSlots.nProc++;
        }
        if (d.kind < UGF_PATCH_WALL || d.kind > UGF_PATCH_GENERIC) return fail(h, "unknown patch kind");
        for (int k = 0; k < d.size; ++k) bfPatch[d.startBfi + k] = p;
    }
    for (int b = 0; b < nB; ++b) {
        if (bfPatch[b] < 0) return fail(h, "boundary face not covered by a patch");
        bfOwner[b] = m->owner[nI + b];
        for (int k = 0; k < 3; ++k) bfS[3 * (size_t)b + k] = m->faceAreas[3 * (size_t)(nI + b) + k];
    }
    // host copies needed later (inflow geometry)
    h->ownerHost.assign(m->owner, m->owner + m->nFaces);
    h->neighbourHost.assign(m->neighbour, m->neighbour + nI);
    h->cfOffHost.assign(m->cellFaceOffsets, m->cellFaceOffsets + nC + 1);
    h->cfHost.assign(m->cellFaces, m->cellFaces + m->cellFaceOffsets[nC]);
    h->patchPartnerHost.assign(m->patchPartner, m->patchPartner + h->nPatches);
    h->ccHost.assign(m->cellCentres, m->cellCentres + 3 * (size_t)nC);
    h->SfHost.assign(m->faceAreas, m->faceAreas + 3 * (size_t)m->nFaces);
    h->CfHost.assign(m->faceCentres, m->faceCentres + 3 * (size_t)m->nFaces);
    if (m->points && m->facePointOffsets && m->facePoints) {
        h->pointsHost.assign(m->points, m->points + 3 * (size_t)m->nPoints);
        h->fpOffHost.assign(m->facePointOffsets, m->facePointOffsets + m->nFaces + 1);
        h->fpHost.assign(m->facePoints, m->facePoints + h->fpOffHost[m->nFaces]);
    }
    if (dalloc(h, &h->dCfOff, (size_t)nC + 1) || dalloc(h, &h->dPlane, (size_t)h->nSlots) || dalloc(h, &h->dNbr, (size_t)h->nSlots) ||
        dalloc(h, &h->dBfPatch, (size_t)nB) || dalloc(h, &h->dBfOwner, (size_t)nB) || dalloc(h, &h->dPatches, (size_t)h->nPatches) ||
        dalloc(h, &h->dVol, (size_t)nC) || dalloc(h, &h->dBbMin, 3 * (size_t)nC) || dalloc(h, &h->dBbMax, 3 * (size_t)nC) ||
        dalloc(h, &h->dBfS, 3 * (size_t)nB))
        return 1;
    if (upload(h, h->dCfOff, cfOffDev.data(), (size_t)nC + 1) || upload(h, h->dPlane, plane.data(), plane.size()) ||
        upload(h, h->dNbr, nbr.data(), nbr.size()) || upload(h, h->dBfPatch, bfPatch.data(), (size_t)nB) ||
        upload(h, h->dBfOwner, bfOwner.data(), (size_t)nB) || upload(h, h->dPatches, h->patchesHost.data(), (size_t)h->nPatches) ||
        upload(h, h->dVol, m->cellVolumes, (size_t)nC) || upload(h, h->dBbMin, m->cellBbMin, 3 * (size_t)nC) ||
        upload(h, h->dBbMax, m->cellBbMax, 3 * (size_t)nC) || upload(h, h->dBfS, bfS.data(), 3 * (size_t)nB))
        return 1;
    CU(cudaStreamSynchronize(h->stream));  // host staging vectors go out of scope
    h->mesh.nCells = nC; h->mesh.nBFaces = nB; h->mesh.nPatches = h->nPatches; h->mesh.planeNoZ = planeNoZ;
    h->mesh.cfOff = h->dCfOff; h->mesh.plane = h->dPlane; h->mesh.nbr = h->dNbr; h->mesh.bfPatch = h->dBfPatch;
    if (h->moveNF == 4 && planeNoZ && !(std::getenv("UGF_MOVE_REC") && std::atoi(std::getenv("UGF_MOVE_REC")) == 0)) {
        // packed per-cell record of the move kernel (one 128-byte line per cell), see track_parcel
        std::vector<double> rec((size_t)nC * 16, 0.0);
        for (int c = 0; c < nC; ++c) {
            double* r = &rec[(size_t)c * 16];
            int* ri = reinterpret_cast<int*>(r + 12);
            for (int f = 0; f < 4; ++f) {
                const double4& pl = plane[(size_t)c * 4 + f];
                r[2 * f] = pl.x; r[2 * f + 1] = pl.y; r[8 + f] = pl.z;
                ri[f] = nbr[(size_t)c * 4 + f];
            }
        }
        if (dalloc(h, &h->dRec2d, rec.size()) || upload(h, h->dRec2d, rec.data(), rec.size())) return 1;
        CU(cudaStreamSynchronize(h->stream));
        h->mesh.rec2d = h->dRec2d;
    }
    h->mesh.bfOwner = h->dBfOwner; h->mesh.patches = h->dPatches; h->mesh.vol = h->dVol; h->mesh.bbMin = h->dBbMin; h->mesh.bbMax = h->dBbMax;

    // per-cell arrays
    const size_t nS = (size_t)h->nSpecies;
    if (dalloc(h, &h->dCellCount, (size_t)nC) || dalloc(h, &h->dOff, (size_t)nC + 1) ||
        dalloc(h, &h->dBlockSums, (size_t)std::max((nC + SCAN_TILE - 1) / SCAN_TILE, (int)(h->capacity / 1024 + 2)) + 2) ||
        dalloc(h, &h->dMigCount, (size_t)h->nPatches) || dalloc(h, &h->dMom, (size_t)nC * nS * UGF_NMOM) ||
        dalloc(h, &h->dAcc, (size_t)nC * NACC) || dalloc(h, &h->dBm, (size_t)nB * UGF_NBM) || dalloc(h, &h->dBacc, (size_t)nB * UGF_NBM) ||
        dalloc(h, &h->dSigma, (size_t)nC) || dalloc(h, &h->dCollId, (size_t)nC) || dalloc(h, &h->dMaxProb, (size_t)nC) ||
        dalloc(h, &h->dQPrev, 3 * (size_t)nC) || dalloc(h, &h->dSPrev, 6 * (size_t)nC))
        return 1;
    CU(cudaMemsetAsync(h->dMigCount, 0, sizeof(int) * std::max(h->nPatches, 1), h->stream));
    CU(cudaMemsetAsync(h->dOff, 0, sizeof(int) * ((size_t)nC + 1), h->stream));
    CU(cudaMemsetAsync(h->dCellCount, 0, sizeof(int) * (size_t)nC, h->stream));
    CU(cudaMemsetAsync(h->dMom, 0, sizeof(double) * (size_t)nC * nS * UGF_NMOM, h->stream));
    CU(cudaMemsetAsync(h->dAcc, 0, sizeof(double) * (size_t)nC * NACC, h->stream));
    if (h->hasVib || h->hasElec) {
        if (h->hasProcessor) return fail(h, "vibrational / electronic levels do not travel across processor patches yet");
        if (dalloc(h, &h->dMomI, (size_t)nC * nS * UGF_NINT) || dalloc(h, &h->dAccI, (size_t)nC * nS * UGF_NINT)) return 1;
        CU(cudaMemsetAsync(h->dMomI, 0, sizeof(double) * (size_t)nC * nS * UGF_NINT, h->stream));
        CU(cudaMemsetAsync(h->dAccI, 0, sizeof(double) * (size_t)nC * nS * UGF_NINT, h->stream));
    }
    if (h->multi) {
        if (dalloc(h, &h->dAccS, (size_t)nC * nS)) return 1;
        CU(cudaMemsetAsync(h->dAccS, 0, sizeof(double) * (size_t)nC * nS, h->stream));
    }
    CU(cudaMemsetAsync(h->dBm, 0, sizeof(double) * std::max<size_t>((size_t)nB * UGF_NBM, 1), h->stream));
    CU(cudaMemsetAsync(h->dBacc, 0, sizeof(double) * std::max<size_t>((size_t)nB * UGF_NBM, 1), h->stream));
    CU(cudaMemsetAsync(h->dSigma, 0, sizeof(double) * (size_t)nC, h->stream));
    CU(cudaMemsetAsync(h->dQPrev, 0, sizeof(double) * 3 * (size_t)nC, h->stream));
    CU(cudaMemsetAsync(h->dSPrev, 0, sizeof(double) * 6 * (size_t)nC, h->stream));
    {
        std::vector<int> ids((size_t)nC, h->cfg.collisionModel == UGF_COLL_DSMC ? 1 : 0);  // uniGasCloud.C:713,723,731
        std::vector<double> ones((size_t)nC, 1.0);
        if (upload(h, h->dCollId, ids.data(), ids.size()) || upload(h, h->dMaxProb, ones.data(), ones.size())) return 1;
        CU(cudaStreamSynchronize(h->stream));
    }
    if (h->hasProcessor && dalloc(h, &h->dSf, (size_t)h->capacity)) return 1;
    if (h->hasProcessor && dalloc(h, &h->dMigList, (size_t)MIG_MAXP * MIG_LIST_CAP)) return 1;
    if (const char* e = std::getenv("UGF_MIG_SEARCH_PACK")) h->migSearchPack = std::atoi(e) != 0;
    if (bgk_active(h) && dalloc(h, &h->dKeyScratch, (size_t)h->capacity)) return 1;
    h->packBuf.assign(h->nPatches, nullptr);
    h->packCap.assign(h->nPatches, 0);

    // launch geometry: persistent grids sized to the SM count
    h->cellSmem = cell_smem_bytes(h->hasRot);
    h->bgkSmem = (size_t)BGK_WARPS * sizeof(BgkWarpSmem);
    int occCell = 1, occNtc = 1, occBgk = 1;
    cudaError_t e1 = cudaSuccess;
    dispatch(h, [&](auto R, auto M) {
        auto k = cell_kernel<decltype(R)::value, decltype(M)::value>;
        e1 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cellSmem);
        if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occCell, k, CELL_THREADS, h->cellSmem);
        auto kn = ntc_kernel<decltype(R)::value, decltype(M)::value, false>;
        if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occNtc, kn, NTC_THREADS, 0);
    });
    CU(e1);
    if (h->multi) {
        CU(cudaFuncSetAttribute(bgk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->bgkSmem));
        CU(cudaFuncSetAttribute(bgk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->bgkSmem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occBgk, bgk_kernel<true>, BGK_THREADS, h->bgkSmem));
    } else {
        CU(cudaFuncSetAttribute(bgk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->bgkSmem));
        CU(cudaFuncSetAttribute(bgk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->bgkSmem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occBgk, bgk_kernel<false>, BGK_THREADS, h->bgkSmem));
    }
    auto persistent = [&](int occ, int warpsPerBlock) {
        const int need = (nC + warpsPerBlock - 1) / warpsPerBlock;
        return std::max(1, std::min(need, h->numSMs * std::max(occ, 1)));
    };
    // The streaming kernel's cp.async fills need L1 lines while in flight: fewer resident blocks leave a larger
    // L1 next to the shared-memory carve-out and raise the achievable memory-level parallelism (profiles/).
    int cellBps = 4, cellCarve = -1;
    if (const char* e = std::getenv("UGF_MOVE_DIRECT")) h->moveDirect = std::atoi(e) != 0;
    if (const char* e = std::getenv("UGF_MOVE_V2")) h->moveV2 = std::atoi(e) != 0;
    if (const char* e = std::getenv("UGF_MOVE_BPS")) h->moveBps = std::max(0, std::min(4, std::atoi(e)));
    if (const char* e = std::getenv("UGF_CELL_TASK")) h->cellTask = std::max(1, std::min(CELL_TASK_MAX, std::atoi(e)));
    if (const char* e = std::getenv("UGF_CELL_FLAGS")) h->cellFlags = std::atoi(e);
    if (const char* e = std::getenv("UGF_CELL_BPS")) cellBps = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("UGF_CELL_CARVEOUT")) cellCarve = std::atoi(e);
    occCell = std::min(occCell, cellBps);
    if (cellCarve < 0) cellCarve = (int)std::min<size_t>(100, (occCell * (h->cellSmem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    dispatch(h, [&](auto R, auto M) {
        e1 = cudaFuncSetAttribute(cell_kernel<decltype(R)::value, decltype(M)::value>, cudaFuncAttributePreferredSharedMemoryCarveout, cellCarve);
    });
    CU(e1);
    h->cellBlocks = persistent(occCell, CELL_WARPS * 2);
    h->ntcBlocks = persistent(occNtc, NTC_WARPS * 32);
    h->bgkBlocks = persistent(occBgk, BGK_WARPS * BGK_CHUNK);
    h->meshSet = true;
    if (h->cfg.axisymmetric) {
        // RWF of every cell centre and boundary face centre, evaluated on the host with the expression of uniGasCloudI.H:116-120
        auto rwf = [&](const double* x) { return 1.0 + (h->cfg.maxRWF - 1.0) * std::sqrt(x[1] * x[1] + x[2] * x[2]) / h->cfg.radialExtent; };
        h->cellRwfHost.resize(nC);
        for (int c = 0; c < nC; ++c) h->cellRwfHost[c] = rwf(&h->ccHost[3 * (size_t)c]);
        h->bfRwfHost.resize(std::max(nB, 1));
        for (int b = 0; b < nB; ++b) h->bfRwfHost[b] = rwf(&h->CfHost[3 * (size_t)(nI + b)]);
        if (dalloc(h, &h->dCellRwf, nC) || dalloc(h, &h->dBfRwf, (size_t)std::max(nB, 1))) return 1;
        if (upload(h, h->dCellRwf, h->cellRwfHost.data(), nC) || upload(h, h->dBfRwf, h->bfRwfHost.data(), (size_t)std::max(nB, 1))) return 1;
        h->prm.cellRwf = h->dCellRwf;
        h->prm.bfRwf = h->dBfRwf;
        // the weighting pass works on CWF * RWF: without cellWeightedSimulation the cell factors are a field of ones
        std::vector<double> ones(nC, 1.0);
        if (ugf_upload_cell_state(h, nullptr, nullptr, nullptr, ones.data())) return 1;
    }
    return 0;
}

int ugf_set_patch_model(ugf_handle* h, int32_t patch, int32_t model, const double* prm, int32_t n) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->patchKind[patch] != UGF_PATCH_WALL) return fail(h, "patch models apply to wall patches only");
    DevPatch& d = h->patchesHost[patch];
    if (model == UGF_WALL_DIFFUSE || model == UGF_WALL_MIXED || model == UGF_WALL_CLL) {
        if (n < 4 || !prm) return fail(h, "diffuse wall needs T, Ux, Uy, Uz");
        d.T = prm[0]; d.Uw[0] = prm[1]; d.Uw[1] = prm[2]; d.Uw[2] = prm[3];
        if (model == UGF_WALL_MIXED) { if (n < 5) return fail(h, "mixed wall needs diffuseFraction"); d.diffuseFraction = prm[4]; }
        if (model == UGF_WALL_CLL) {
            if (n < 7) return fail(h, "CLL wall needs normalAccommCoeff, tangentialAccommCoeff, rotEnergyAccommCoeff");
            d.alphaN = prm[4]; d.sigmaT = prm[5]; d.alphaR = prm[6];
        }
    } else if (model != UGF_WALL_SPECULAR && model != UGF_WALL_DELETION) {
        return fail(h, "unknown wall model");
    }
    d.wallModel = model;
    CU(countedMemcpyAsync(h, h->dPatches + patch, &d, sizeof(DevPatch), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ugf_set_patch_wall_fields(ugf_handle* h, int32_t patch, const double* T, const double* U) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->patchKind[patch] != UGF_PATCH_WALL || h->patchesHost[patch].wallModel == UGF_WALL_UNSET)
        return fail(h, "wall fields need a wall patch with a model");
    if (!T || !U) return fail(h, "null wall field");
    DevPatch& d = h->patchesHost[patch];
    const size_t n = (size_t)d.size;
    double *dT = nullptr, *dU = nullptr;
    if (dalloc(h, &dT, n) || dalloc(h, &dU, 3 * n)) return 1;
    h->wallFieldOwned.push_back(dT);
    h->wallFieldOwned.push_back(dU);
    if (upload(h, dT, T, n) || upload(h, dU, U, 3 * n)) return 1;
    d.faceT = dT; d.faceU = dU;
    CU(countedMemcpyAsync(h, h->dPatches + patch, &d, sizeof(DevPatch), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// upper bound of the parcels one step can insert (sizes the launches and the capacity check of do_inflow); depends on
// deltaT and, with cell weighting, on the factor of each inlet face's cell (uniGasGeneralBoundary.C:154-165)
static void recompute_inflow_bounds(ugf_handle* h) {
    for (InflowHost& f : h->inflows) {
        if (f.massFlow) continue;  // bounded by the device-computed expectation instead (update_inlet_velocities)
        f.maxInsert = 0;
        for (size_t k = 0; k < f.accum1.size(); ++k) {
            const double w = h->cwfHost.empty() ? 1.0 : h->cwfHost[f.slotCell[k]];
            const double accum = f.accum1[k] * h->cfg.deltaT / (h->cfg.nParticle * w);
            f.maxInsert += (long long)std::max(accum, 0.0) + 2;
        }
    }
}

// common part of the free-stream and the pressure inlet: geometry of the patch faces, insertion bound, device tables.
// pin != null: pressure inlet (mole fractions, velocity per face, relaxation factor); fld != null: number density,
// temperatures and velocity per face (uniGasFreeStreamInflowFieldPatch)
struct InflowFields { const double *numDen, *transT, *rotT, *U; };  // numDen [nTypeIds][nFaces]
static int set_inflow_common(ugf_handle* h, int32_t patch, const ugf_inflow* in, const ugf_pressure_inlet* pin, const InflowFields* fld = nullptr,
                             bool wang = false, bool outlet = false, const double* ceQ = nullptr, const double* ceS = nullptr) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (patch < 0 || patch >= h->nPatches) return fail(h, "patch out of range");
    if (h->pointsHost.empty()) return fail(h, "inflow needs mesh points/facePoints");
    if (in->nTypeIds < 1 || in->nTypeIds > h->nSpecies) return fail(h, "inflow typeIds out of range");
    InflowHost f;
    f.patch = patch;
    f.in = *in;
    const int nF = h->patchSize[patch];
    f.nSlots = nF * in->nTypeIds;
    std::vector<int> faceBfi(nF), faceCell(nF), triOff(nF + 1, 0);
    std::vector<double> geom((size_t)nF * INFLOW_GEOM), tri;
    const double sqrtPi = std::sqrt(PI);
    f.maxInsert = 0;
    for (int lf = 0; lf < nF; ++lf) {
        const int face = h->patchStart[patch] + lf;
        faceBfi[lf] = face - h->nInternal;
        faceCell[lf] = h->ownerHost[face];
        const double* S = &h->SfHost[3 * (size_t)face];
        const double* fC = &h->CfHost[3 * (size_t)face];
        const double fA = std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
        const int np = h->fpOffHost[face + 1] - h->fpOffHost[face];
        if (np < 3) return fail(h, "inflow face with fewer than 3 points");
        const int32_t* fp = &h->fpHost[h->fpOffHost[face]];
        const double* p0 = &h->pointsHost[3 * (size_t)fp[0]];
        double cum = 0;
        for (int t = 0; t < np - 2; ++t) {
            const double* a = &h->pointsHost[3 * (size_t)fp[t + 1]];
            const double* b = &h->pointsHost[3 * (size_t)fp[t + 2]];
            const double e1[3] = {a[0] - p0[0], a[1] - p0[1], a[2] - p0[2]};
            const double e2[3] = {b[0] - p0[0], b[1] - p0[1], b[2] - p0[2]};
            const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
            cum += 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz) / fA;
            const double rec[INFLOW_TRI] = {a[0], a[1], a[2], b[0], b[1], b[2], (t == np - 3) ? 1.0 : cum};
            tri.insert(tri.end(), rec, rec + INFLOW_TRI);
        }
        triOff[lf + 1] = triOff[lf] + (np - 2);
        const double n[3] = {S[0] / -fA, S[1] / -fA, S[2] / -fA};
        double t1[3] = {fC[0] - p0[0], fC[1] - p0[1], fC[2] - p0[2]};
        const double m1 = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
        for (int k = 0; k < 3; ++k) t1[k] /= m1;
        double t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
        const double m2 = std::sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
        for (int k = 0; k < 3; ++k) t2[k] /= m2;
        double* g = &geom[(size_t)lf * INFLOW_GEOM];
        g[0] = fA;
        for (int k = 0; k < 3; ++k) { g[1 + k] = n[k]; g[4 + k] = t1[k]; g[7 + k] = t2[k]; g[10 + k] = p0[k]; }
        for (int iD = 0; iD < in->nTypeIds; ++iD) {
            const int t = in->typeIds[iD];
            if (t < 0 || t >= h->nSpecies) return fail(h, "inflow typeId out of range");
            const double Ttr = fld ? fld->transT[lf] : in->translationalTemperature;
            const double* vel = fld ? fld->U + 3 * (size_t)lf : in->velocity;
            const double numDen = fld ? fld->numDen[(size_t)iD * nF + lf] : in->numberDensities[iD];
            if (!(Ttr > 0.0) || !(numDen >= 0.0)) return fail(h, "inflow needs a positive temperature and a non-negative number density on every face");
            const double cmp = std::sqrt(2.0 * kB * Ttr / h->spHost[t].mass);
            double sCos = (vel[0] * n[0] + vel[1] * n[1] + vel[2] * n[2]) / cmp;
            if (pin) sCos = 5.0;  // the face velocities follow the flow: bound the insertions with a speed ratio of 5
            double accum = (pin ? pin->moleFractions[iD] : 1.0) * (fA * numDen * h->cfg.deltaT * cmp * (std::exp(-(sCos * sCos)) + sqrtPi * sCos * (1 + std::erf(sCos))))
                           / (2.0 * sqrtPi * h->cfg.nParticle);
            if (ceQ) {  // Chapman-Enskog count (uniGasGeneralBoundary.C:171-239) for the capacity bound
                double pr = 0.0;
                for (int j = 0; j < in->nTypeIds; ++j) pr += in->numberDensities[j];
                pr *= kB * Ttr;
                const double qn = ceQ[0] * n[0] + ceQ[1] * n[1] + ceQ[2] * n[2];
                double snn = 0.0;
                for (int k = 0; k < 3; ++k) snn += (ceS[3 * k] * n[0] + ceS[3 * k + 1] * n[1] + ceS[3 * k + 2] * n[2]) * n[k];
                accum = (fA * numDen * h->cfg.deltaT * cmp * (std::exp(-(sCos * sCos)) * (1.0 - 0.5 * snn / pr - 0.4 * qn * sCos / pr / cmp) + sqrtPi * sCos * (1 + std::erf(sCos))))
                        / (2.0 * sqrtPi * h->cfg.nParticle);
            }
            f.maxInsert += (long long)std::max(accum, 0.0) + 2;
            f.accum1.push_back(accum * h->cfg.nParticle / h->cfg.deltaT);
            f.slotCell.push_back(faceCell[lf]);
        }
    }
    InflowDev& d = f.dev;
    std::memset(&d, 0, sizeof(d));
    d.nFaces = nF; d.nTypeIds = in->nTypeIds;
    for (int i = 0; i < in->nTypeIds; ++i) { d.typeIds[i] = in->typeIds[i]; d.numDen[i] = in->numberDensities[i]; }
    d.Ttr = in->translationalTemperature; d.Trot = in->rotationalTemperature;
    d.Tvib = in->vibrationalTemperature; d.Tel = in->electronicTemperature;
    if (ceQ) {
        d.ce = 1;
        for (int k = 0; k < 3; ++k) d.ceQ[k] = ceQ[k];
        for (int k = 0; k < 9; ++k) d.ceS[k] = ceS[k];
        d.cePressure = 0.0;
        for (int j = 0; j < in->nTypeIds; ++j) d.cePressure += in->numberDensities[j];
        d.cePressure *= kB * in->translationalTemperature;
    }
    for (int k = 0; k < 3; ++k) d.vel[k] = in->velocity[k];
    for (int i = 0; i < UGF_MAX_SPECIES; ++i) d.molFrac[i] = (pin && i < in->nTypeIds) ? pin->moleFractions[i] : 1.0;
    d.theta = pin ? pin->theta : 1.0;
    d.pressure = pin ? 1 : 0;
    int *dBfi, *dCell, *dTriOff, *dNIns, *dInsOff;
    double *dGeom, *dTri;
    if (dalloc(h, &dBfi, (size_t)nF) || dalloc(h, &dCell, (size_t)nF) || dalloc(h, &dTriOff, (size_t)nF + 1) || dalloc(h, &dGeom, geom.size()) ||
        dalloc(h, &dTri, tri.size()) || dalloc(h, &dNIns, (size_t)f.nSlots) || dalloc(h, &dInsOff, (size_t)f.nSlots + 1))
        return 1;
    f.owned = {dBfi, dCell, dTriOff, dGeom, dTri, dNIns, dInsOff};
    if (upload(h, dBfi, faceBfi.data(), faceBfi.size()) || upload(h, dCell, faceCell.data(), faceCell.size()) ||
        upload(h, dTriOff, triOff.data(), triOff.size()) || upload(h, dGeom, geom.data(), geom.size()) || upload(h, dTri, tri.data(), tri.size()))
        return 1;
    CU(cudaStreamSynchronize(h->stream));
    d.faceBfi = dBfi; d.faceCell = dCell; d.triOff = dTriOff; d.geom = dGeom; d.tri = dTri; d.nIns = dNIns; d.insOff = dInsOff;
    if (pin) {  // inletVelocity_ starts at zero (…LiouFangPressureInletPatch.C:63)
        double* dVel;
        if (dalloc(h, &dVel, 3 * (size_t)nF)) return 1;
        CU(cudaMemsetAsync(dVel, 0, sizeof(double) * 3 * (size_t)nF, h->stream));
        f.owned.push_back(dVel);
        d.faceVel = dVel;
        f.pressureInlet = true;
        if (wang) {  // …/uniGasWangPressureInletPatch.C:133-157: mixture molecular mass, gamma, R = k / m
            double M = 0, cp = 0, cv = 0;
            for (int i = 0; i < in->nTypeIds; ++i) {
                const ugf_species& sp = h->spHost[in->typeIds[i]];
                M += sp.mass * pin->moleFractions[i];
                cp += (5.0 + sp.rotationalDoF) * pin->moleFractions[i];
                cv += (3.0 + sp.rotationalDoF) * pin->moleFractions[i];
            }
            if (!(M > 0.0)) return fail(h, "mole fractions of the pressure inlet sum to zero");
            double* dS;
            if (dalloc(h, &dS, (size_t)WANG_NSUM * nF)) return 1;
            CU(cudaMemsetAsync(dS, 0, sizeof(double) * WANG_NSUM * (size_t)nF, h->stream));
            f.owned.push_back(dS);
            d.wangSums = dS; d.wangP = pin->inletPressure; d.wangM = M; d.wangGammaR = (cp / cv) * (kB / M);
            f.wang = true;
        }
        if (outlet) {  // outletNumberDensity_ 0, outletTemperature_ 300 K, outletVelocity_ 0 at the start (…PressureOutletPatch.C:61-74)
            std::vector<double> fT(2 * (size_t)nF, pin->inletTemperature);
            double *dN, *dT;
            if (dalloc(h, &dN, (size_t)f.nSlots) || dalloc(h, &dT, fT.size())) return 1;
            CU(cudaMemsetAsync(dN, 0, sizeof(double) * (size_t)f.nSlots, h->stream));
            f.owned.push_back(dN); f.owned.push_back(dT);
            if (upload(h, dT, fT.data(), fT.size())) return 1;
            CU(cudaStreamSynchronize(h->stream));
            d.faceN = dN; d.faceT = dT; d.outlet = 1; d.capN = in->numberDensities[0]; d.capT = pin->inletTemperature; d.err = h->dErr;
            f.outlet = true;
        }
    }
    if (fld) {  // per-face tables: velocity [nF*3], number density per slot [nF*nTypeIds], (Ttr, Trot) [nF*2]
        std::vector<double> fN((size_t)f.nSlots), fT(2 * (size_t)nF);
        for (int lf = 0; lf < nF; ++lf) {
            for (int iD = 0; iD < in->nTypeIds; ++iD) fN[(size_t)lf * in->nTypeIds + iD] = fld->numDen[(size_t)iD * nF + lf];
            fT[2 * (size_t)lf] = fld->transT[lf];
            fT[2 * (size_t)lf + 1] = fld->rotT ? fld->rotT[lf] : 0.0;
        }
        double *dVel, *dN, *dT;
        if (dalloc(h, &dVel, 3 * (size_t)nF) || dalloc(h, &dN, fN.size()) || dalloc(h, &dT, fT.size())) return 1;
        f.owned.push_back(dVel); f.owned.push_back(dN); f.owned.push_back(dT);
        if (upload(h, dVel, fld->U, 3 * (size_t)nF) || upload(h, dN, fN.data(), fN.size()) || upload(h, dT, fT.data(), fT.size())) return 1;
        CU(cudaStreamSynchronize(h->stream));
        d.faceVel = dVel; d.faceN = dN; d.faceT = dT;
    }
    h->inflows.push_back(f);
    return 0;
}

int ugf_set_inflow(ugf_handle* h, int32_t patch, const ugf_inflow* in) { return set_inflow_common(h, patch, in, nullptr); }

int ugf_set_chapman_enskog_inflow(ugf_handle* h, int32_t patch, const ugf_inflow* in, const double* heatFlux, const double* stress) {
    if (!heatFlux || !stress) return fail(h, "Chapman-Enskog inflow needs heatFlux [3] and stress [9]");
    return set_inflow_common(h, patch, in, nullptr, nullptr, false, false, heatFlux, stress);
}

int ugf_set_inflow_fields(ugf_handle* h, int32_t patch, int32_t nTypeIds, const int32_t* typeIds, const double* numberDensity,
                          const double* transT, const double* rotT, const double* U) {
    if (!h) return 1;
    if (!typeIds || !numberDensity || !transT || !U) return fail(h, "null inflow field");
    if (nTypeIds < 1 || nTypeIds > UGF_MAX_SPECIES) return fail(h, "inflow typeIds out of range");
    ugf_inflow in{};
    in.nTypeIds = nTypeIds;
    for (int i = 0; i < nTypeIds; ++i) in.typeIds[i] = typeIds[i];
    const InflowFields fld{numberDensity, transT, rotT, U};
    return set_inflow_common(h, patch, &in, nullptr, &fld);
}

int ugf_set_pressure_inlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* pin) {
    if (!h || !pin) return 1;
    if (!(pin->theta >= 0.0 && pin->theta <= 1.0)) return fail(h, "Theta must be a value between 0 and 1");  // :69-75
    if (pin->nTypeIds < 1 || pin->nTypeIds > UGF_MAX_SPECIES) return fail(h, "inflow typeIds out of range");
    ugf_inflow in{};
    in.nTypeIds = pin->nTypeIds;
    const double n = pin->inletPressure / (kB * pin->inletTemperature);  // :102
    for (int i = 0; i < pin->nTypeIds; ++i) { in.typeIds[i] = pin->typeIds[i]; in.numberDensities[i] = n; }
    in.translationalTemperature = in.rotationalTemperature = in.vibrationalTemperature = in.electronicTemperature = pin->inletTemperature;
    return set_inflow_common(h, patch, &in, pin);
}

int ugf_set_mass_flow_inlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* pin, double massFlowRate, const double* initialVelocity) {
    if (!h || !pin) return 1;
    if (!(pin->theta >= 0.0 && pin->theta <= 1.0)) return fail(h, "Theta must be a value between 0 and 1");
    if (!(massFlowRate > 0.0) || !(pin->inletTemperature > 0.0)) return fail(h, "mass-flow-rate inlet needs a positive massFlowRate and inletTemperature");
    if (pin->nTypeIds != h->nSpecies) return fail(h, "mass-flow-rate inlet: typeIds must list every species in typeIdList order");
    for (int i = 0; i < pin->nTypeIds; ++i) if (pin->typeIds[i] != i) return fail(h, "mass-flow-rate inlet: typeIds must list every species in typeIdList order");
    if (h->hasProcessor) return fail(h, "mass-flow-rate inlet on a decomposed mesh is not supported (its patch-wide sums would need a reduction over the ranks)");
    if (patch < 0 || patch >= h->nPatches || h->patchKind[patch] != UGF_PATCH_GENERIC) return fail(h, "mass-flow-rate inlet needs a patch of type patch");
    double totalMass = 0.0;
    for (int i = 0; i < pin->nTypeIds; ++i) totalMass += h->spHost[i].mass * pin->moleFractions[i];
    if (!(totalMass > 0.0)) return fail(h, "mole fractions of the mass-flow-rate inlet sum to zero");
    ugf_inflow in{};
    in.nTypeIds = pin->nTypeIds;
    for (int i = 0; i < pin->nTypeIds; ++i) { in.typeIds[i] = i; in.numberDensities[i] = 0.0; }  // inletNumberDensity_ starts at zero (:84-89)
    in.translationalTemperature = in.rotationalTemperature = in.vibrationalTemperature = in.electronicTemperature = pin->inletTemperature;
    ugf_pressure_inlet q = *pin;
    for (int i = 0; i < pin->nTypeIds; ++i) q.moleFractions[i] = 1.0;  // the count takes the per-species number densities as they are
    if (int rc = set_inflow_common(h, patch, &in, &q)) return rc;
    InflowHost& f = h->inflows.back();
    InflowDev& d = f.dev;
    const size_t nF = (size_t)d.nFaces, nT = (size_t)d.nTypeIds;
    double *dN, *dFlux, *dWork, *dTot;
    if (dalloc(h, &dN, nF * nT) || dalloc(h, &dFlux, nF * (size_t)h->nSpecies) || dalloc(h, &dWork, nF * (nT + 1)) || dalloc(h, &dTot, 2)) return 1;
    f.owned.push_back(dN); f.owned.push_back(dFlux); f.owned.push_back(dWork); f.owned.push_back(dTot);
    CU(cudaMemsetAsync(dN, 0, sizeof(double) * nF * nT, h->stream));
    CU(cudaMemsetAsync(dFlux, 0, sizeof(double) * nF * (size_t)h->nSpecies, h->stream));
    CU(cudaMemsetAsync(dTot, 0, sizeof(double) * 2, h->stream));
    std::vector<double> v0(3 * nF, 0.0);
    if (initialVelocity) for (size_t lf = 0; lf < nF; ++lf) for (int k = 0; k < 3; ++k) v0[3 * lf + k] = initialVelocity[k];
    if (upload(h, d.faceVel, v0.data(), v0.size())) return 1;
    d.faceN = dN; d.outFlux = dFlux; d.mfWork = dWork; d.mfTotal = dTot;
    d.massFlow = 1; d.massFlowRate = massFlowRate;
    for (int i = 0; i < pin->nTypeIds; ++i) d.mfMolFrac[i] = pin->moleFractions[i];
    d.patchArea = 0.0;
    for (size_t lf = 0; lf < nF; ++lf) {
        const double* S = &h->SfHost[3 * ((size_t)h->patchStart[patch] + lf)];
        d.patchArea += std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
    }
    f.massFlow = true;
    f.mfExpected = 0.0;
    f.maxInsert = (long long)f.nSlots + 2;
    DevPatch& dp = h->patchesHost[patch];
    dp.outFlux = dFlux;
    CU(countedMemcpyAsync(h, h->dPatches + patch, &dp, sizeof(DevPatch), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ugf_set_wang_pressure_inlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* pin) {
    if (!h || !pin) return 1;
    if (pin->nTypeIds < 1 || pin->nTypeIds > UGF_MAX_SPECIES) return fail(h, "inflow typeIds out of range");
    ugf_inflow in{};
    in.nTypeIds = pin->nTypeIds;
    const double n = pin->inletPressure / (kB * pin->inletTemperature);  // …/uniGasWangPressureInletPatch.C:107
    for (int i = 0; i < pin->nTypeIds; ++i) { in.typeIds[i] = pin->typeIds[i]; in.numberDensities[i] = n; }
    in.translationalTemperature = in.rotationalTemperature = in.vibrationalTemperature = in.electronicTemperature = pin->inletTemperature;
    return set_inflow_common(h, patch, &in, pin, nullptr, true);
}

int ugf_set_pressure_outlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* pout) {
    if (!h || !pout) return 1;
    if (pout->nTypeIds < 1 || pout->nTypeIds > UGF_MAX_SPECIES) return fail(h, "inflow typeIds out of range");
    if (!(pout->inletPressure > 0.0) || !(pout->inletTemperature > 0.0)) return fail(h, "pressure outlet needs a positive pressure and initial temperature");
    ugf_inflow in{};
    in.nTypeIds = pout->nTypeIds;
    const double n = 2.0 * pout->inletPressure / (kB * pout->inletTemperature);  // the bound's number density: twice p_e / (k T_0)
    for (int i = 0; i < pout->nTypeIds; ++i) { in.typeIds[i] = pout->typeIds[i]; in.numberDensities[i] = n; }
    in.translationalTemperature = in.rotationalTemperature = in.vibrationalTemperature = in.electronicTemperature = pout->inletTemperature;
    return set_inflow_common(h, patch, &in, pout, nullptr, true, true);
}

int ugf_download_inlet_velocity(ugf_handle* h, int32_t patch, double* U) {
    if (!h) return 1;
    for (InflowHost& f : h->inflows)
        if (f.patch == patch && f.pressureInlet) {
            CU(countedMemcpyAsync(h, U, f.dev.faceVel, sizeof(double) * 3 * (size_t)f.dev.nFaces, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            return 0;
        }
    return fail(h, "no pressure inlet on this patch");
}

// controlParcelsAfterCollisions of the pressure inlets: needs the cell-major array and its offsets (after the gather)
static int update_inlet_velocities(ugf_handle* h) {
    for (InflowHost& f : h->inflows) {
        if (!f.pressureInlet) continue;
        const DevParams prm = h->prm;
        const InflowDev dev = f.dev;
        ParcelBuf P = h->buf[h->cur];
        if (f.massFlow) {
            mass_flow_face_kernel<<<grid_for(dev.nFaces, 128), 128, 0, h->stream>>>(prm, dev, P, h->dOff, h->dVol, h->multi);
            LAUNCHED();
            mass_flow_scale_kernel<<<1, 256, 0, h->stream>>>(dev);
            LAUNCHED();
            // the next step's insertions are bounded by what the patch asks for: 16 bytes back to the host, once per step
            double tot[2] = {0.0, 0.0};
            CU(countedMemcpyAsync(h, tot, dev.mfTotal, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            if (tot[1] == 0.0)
                return fail(h, "mass-flow-rate inlet: no parcels in the cells of the inlet patch (the reference's parcelsIn / parcelsToInsert is 0 / 0 here)");
            f.mfExpected = tot[0];
            f.maxInsert = (long long)std::ceil(std::max(tot[0], 0.0)) + f.nSlots + 2;
            continue;
        }
        if (f.outlet) {
            f.wangSteps += 1.0;  // nTimeSteps_ (…PressureOutletPatch.C:146)
            if (h->multi) outlet_state_kernel<true><<<grid_for(dev.nFaces, 128), 128, 0, h->stream>>>(prm, dev, P, h->dOff, h->dVol, f.wangSteps);
            else outlet_state_kernel<false><<<grid_for(dev.nFaces, 128), 128, 0, h->stream>>>(prm, dev, P, h->dOff, h->dVol, f.wangSteps);
            LAUNCHED();
            continue;
        }
        if (f.wang) {
            f.wangSteps += 1.0;  // nTimeSteps_ (:133)
            if (h->multi) wang_inlet_velocity_kernel<true><<<grid_for(dev.nFaces, 128), 128, 0, h->stream>>>(prm, dev, P, h->dOff, h->dVol, f.wangSteps);
            else wang_inlet_velocity_kernel<false><<<grid_for(dev.nFaces, 128), 128, 0, h->stream>>>(prm, dev, P, h->dOff, h->dVol, f.wangSteps);
            LAUNCHED();
            continue;
        }
        if (h->multi) inlet_velocity_kernel<true><<<grid_for(dev.nFaces, 128), 128, 0, h->stream>>>(prm, dev, P, h->dOff);
        else inlet_velocity_kernel<false><<<grid_for(dev.nFaces, 128), 128, 0, h->stream>>>(prm, dev, P, h->dOff);
        LAUNCHED();
    }
    return 0;
}

int ugf_upload_parcels(ugf_handle* h, const ugf_parcels* p) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (p->n > h->capacity) return fail(h, "parcel count exceeds parcelCapacity");
    if (p->n >= 2147483647LL) return fail(h, "parcel count exceeds int32 indexing");
    CU(cudaSetDevice(h->cfg.device));
    if (alloc_parcels(h)) return 1;
    const size_t n = (size_t)p->n;
    for (size_t i = 0; i < n; ++i) {
        if (p->cell[i] < 0 || p->cell[i] >= h->nCells) return fail(h, "parcel cell out of range");
        if (p->typeId && (p->typeId[i] < 0 || p->typeId[i] >= h->nSpecies)) return fail(h, "parcel typeId out of range");
        if (p->newParcel && p->newParcel[i]) return fail(h, "uploaded parcels must have newParcel == 0");
        if (p->cellWeight && p->cellWeight[i] != (h->cwfHost.empty() ? 1.0 : h->cwfHost[p->cell[i]]))
            return fail(h, "parcel cellWeight differs from the cellWeightFactor of its cell");
    }
    ParcelBuf& P = h->buf[h->cur];
    if (upload(h, P.x, p->x, n) || upload(h, P.y, p->y, n) || upload(h, P.z, p->z, n) || upload(h, P.ux, p->Ux, n) ||
        upload(h, P.uy, p->Uy, n) || upload(h, P.uz, p->Uz, n) || upload(h, P.cell, p->cell, n))
        return 1;
    std::vector<uint8_t> types, elevs;
    std::vector<unsigned long long> vibs;
    if (h->hasRot) {
        if (p->ERot) { if (upload(h, P.erot, p->ERot, n)) return 1; }
        else CU(cudaMemsetAsync(P.erot, 0, n * sizeof(double), h->stream));
    }
    if (p->vibLevel && !h->hasVib) {
        for (size_t i = 0; i < n * UGF_MAX_VIB_MODES; ++i) if (p->vibLevel[i] != 0) return fail(h, "vibLevel given but no species has vibrational modes");
    }
    if (p->ELevel && !h->hasElec) {
        for (size_t i = 0; i < n; ++i) if (p->ELevel[i] != 0) return fail(h, "ELevel given but no species has more than one electronic level");
    }
    if (h->hasVib) {  // 16 bits per mode in one word per parcel
        vibs.assign(n, 0ull);
        if (p->vibLevel)
            for (size_t i = 0; i < n; ++i) {
                const int t = p->typeId ? p->typeId[i] : 0;
                unsigned long long v = 0ull;
                for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) {
                    const int32_t l = p->vibLevel[i * UGF_MAX_VIB_MODES + m];
                    if (l < 0 || l > 65535 || (m >= h->spHost[t].vibrationalDoF && l != 0)) return fail(h, "parcel vibLevel out of range");
                    v |= (unsigned long long)l << (16 * m);
                }
                vibs[i] = v;
            }
        if (upload(h, P.vib, vibs.data(), n)) return 1;
    }
    if (h->hasElec) {
        elevs.assign(n, 0);
        if (p->ELevel)
            for (size_t i = 0; i < n; ++i) {
                const int t = p->typeId ? p->typeId[i] : 0;
                if (p->ELevel[i] < 0 || p->ELevel[i] >= h->spHost[t].nElectronicLevels) return fail(h, "parcel ELevel out of range");
                elevs[i] = (uint8_t)p->ELevel[i];
            }
        if (upload(h, P.elev, elevs.data(), n)) return 1;
    }
    if (h->multi) {
        types.assign(n, 0);
        if (p->typeId) for (size_t i = 0; i < n; ++i) types[i] = (uint8_t)p->typeId[i];
        if (upload(h, P.type, types.data(), n)) return 1;
    }
    if (h->cfg.axisymmetric) {  // radialWeight: see include/ugf.h (ugf_parcels)
        bool asPos = p->radialWeight != nullptr, asCentre = true;
        if (p->radialWeight)
            for (size_t i = 0; i < n; ++i) {
                const double rp = 1.0 + (h->cfg.maxRWF - 1.0) * std::sqrt(p->y[i] * p->y[i] + p->z[i] * p->z[i]) / h->cfg.radialExtent;
                const double rc = h->cellRwfHost[p->cell[i]], r = p->radialWeight[i];
                asPos = asPos && std::fabs(r - rp) <= 1e-6 * rp;
                asCentre = asCentre && std::fabs(r - rc) <= 1e-6 * rc;
            }
        if (!asPos && !asCentre) return fail(h, "radialWeight must be RWF(position) for every parcel or RWF(cell centre) for every parcel");
        h->rwfCentre = !asPos;
        h->prm.rwfCentre = h->rwfCentre ? 1 : 0;
    }
    const long long nn = p->n;
    CU(countedMemcpyAsync(h, h->dN, &nn, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->nUpper = nn; h->nPending = false; h->newFrom = nn; h->recvStart = -1;
    h->cloneValid = false;
    if (h->dCwf[0] && h->cwfDirty) {  // a fresh cloud carries the current factors
        CU(countedMemcpy(h, h->dCwf[h->cwfCur ^ 1], h->dCwf[h->cwfCur], sizeof(double) * h->nCells, cudaMemcpyDeviceToDevice));
        h->cwfHostPrev = h->cwfHost;
        h->cwfDirty = false;
        h->prm.cwfDirty = 0;
    }
    h->histValid = h->occValid = h->occIdentity = h->momValid = false;
    return 0;
}

int ugf_upload_cell_state(ugf_handle* h, const double* s, const int32_t* id, const int32_t* lv, const double* cwf) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    const size_t nC = (size_t)h->nCells;
    if (lv) {
        bool allOne = true;
        for (size_t c = 0; c < nC; ++c) {
            long long prod = 1;
            for (int d = 0; d < 3; ++d) {
                if (lv[3 * c + d] < 1) return fail(h, "subCellLevels must be >= 1");
                prod *= lv[3 * c + d];
                if (lv[3 * c + d] != 1) allOne = false;
            }
            if (prod > 65535) return fail(h, "more than 65535 sub-cells in a cell");
        }
        if (!allOne) {
            if (!h->dSubLevels && dalloc(h, &h->dSubLevels, 3 * nC)) return 1;
            if (!h->dSub && dalloc(h, &h->dSub, (size_t)h->capacity)) return 1;
            if (upload(h, h->dSubLevels, lv, 3 * nC)) return 1;
        }
        h->subLevelsAllOne = allOne;
    }
    if (cwf) {  // cellWeightedSimulation: see include/ugf.h
        for (size_t c = 0; c < nC; ++c) if (!(cwf[c] > 0.0)) return fail(h, "cellWeightFactor must be positive");
        if (h->capacity >= (long long)CLONE_FLAG) return fail(h, "cell weighting needs parcelCapacity < 2^30");
        const bool haveParcels = h->buf[0].x && h->nUpper > 0;
        if (!h->dCwf[0]) {
            if (dalloc(h, &h->dCwf[0], nC) || dalloc(h, &h->dCwf[1], nC)) return 1;
            if (dalloc(h, &h->dNclone, (size_t)h->capacity + MOVE_TILE)) return 1;
            if (h->hasProcessor && dalloc(h, &h->dWq, (size_t)h->capacity)) return 1;
            h->cwfHostPrev.assign(nC, 1.0);  // parcels uploaded so far carry factor 1
            if (upload(h, h->dCwf[1], h->cwfHostPrev.data(), nC)) return 1;
            h->cwfCur = 0;
        } else if (!h->cwfDirty) {
            h->cwfCur ^= 1;  // the previous field stays behind as the factors the parcels carry
            h->cwfHostPrev = h->cwfHost;
        }
        h->cwfHost.assign(cwf, cwf + nC);
        if (upload(h, h->dCwf[h->cwfCur], cwf, nC)) return 1;
        h->cwfDirty = haveParcels;
        if (!haveParcels) {
            h->cwfHostPrev = h->cwfHost;
            if (upload(h, h->dCwf[h->cwfCur ^ 1], cwf, nC)) return 1;
        }
        h->prm.cwf = h->dCwf[h->cwfCur];
        h->prm.cwfPrev = h->dCwf[h->cwfCur ^ 1];
        h->prm.cwfDirty = h->cwfDirty ? 1 : 0;
        recompute_inflow_bounds(h);
    }
    if (s && upload(h, h->dSigma, s, nC)) return 1;
    if (id && upload(h, h->dCollId, id, nC)) return 1;
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ugf_set_deltaT(ugf_handle* h, double dt) {
    if (!h) return 1;
    h->cfg.deltaT = dt;
    h->prm.deltaT = dt;
    recompute_inflow_bounds(h);
    return 0;
}

// ---- phases ---------------------------------------------------------------------------------------------

static int open_step(ugf_handle* h) {
    if (!h->stepOpen) {
        if (zero_step_counters(h)) return 1;
        h->stepOpen = true;
    }
    return 0;
}

int ugf_control_before_move(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (open_step(h)) return 1;
    return do_inflow(h);
}

int ugf_move(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    h->nExact = true;
    if (h->hasProcessor && h->inflows.empty() && !h->inflowDone) {
        if (refresh_n_lazy(h, &h->nExact)) return 1;  // multi-rank, no insertion: do not stall the host on the previous step
    } else if (refresh_n(h)) {
        return 1;
    }
    if (open_step(h)) return 1;
    if (!h->inflowDone) h->newFrom = h->nExact ? h->nUpper : (1LL << 62);
    if (do_move(h, 0, false)) return 1;
    if (h->inflowDone) {  // the insert count is only known on the device: make the host mirror exact again
        if (request_n(h) || refresh_n(h)) return 1;
    }
    h->inflowDone = false;
    h->recvStart = h->nUpper;
    h->nAtMove = h->nExact ? h->nUpper : 0;
    h->slotRound = false;
    return 0;
}

int ugf_sort(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    return do_sort(h);
}

int ugf_reorder(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (!h->occValid && do_sort(h)) return 1;
    if (h->occIdentity) return 0;
    ParcelBuf in = h->buf[h->cur], out = h->buf[h->cur ^ 1];
    dispatch(h, [&](auto R, auto M) {
        reorder_kernel<decltype(R)::value, decltype(M)::value><<<grid_for(h->cloneValid ? h->capacity : h->nUpper, 256), 256, 0, h->stream>>>(in, out, h->dPerm, h->dTotal);
    });
    LAUNCHED();
    return after_gather(h);
}

int ugf_sample(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (!h->occValid && do_sort(h)) return 1;
    return run_cell_kernel(h, false, true);
}

int ugf_collide(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (!h->occValid && do_sort(h)) return 1;
    if (run_cell_kernel(h, !h->occIdentity, !h->momValid)) return 1;  // sampling precedes collisions
    if (!dsmc_active(h)) return 0;
    return run_ntc_kernel(h);
}

int ugf_relax(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (!h->occValid && do_sort(h)) return 1;
    if (run_cell_kernel(h, !h->occIdentity, !h->momValid)) return 1;
    if (!bgk_active(h)) return 0;
    return run_bgk_kernel(h);
}

int ugf_accumulate_fields(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (!h->momValid) {
        if (!h->occValid && do_sort(h)) return 1;
        if (run_cell_kernel(h, false, true)) return 1;
    }
    if (h->occValid && h->occIdentity && update_inlet_velocities(h)) return 1;  // boundaries_.controlAfterCollisions()
    return do_accumulate(h);
}

int ugf_set_decomposition(ugf_handle* h, const ugf_decomposition* d) {
    if (!h || !h->meshSet || !d) return fail(h, "mesh not set");
    if (h->cfg.collisionModel != UGF_COLL_HYBRID) return fail(h, "a decomposition model needs collisionModel hybrid");
    if (d->decompositionInterval < 1 || d->smoothingPasses < 0 || d->refinementPasses < 0 || d->neighborLevels < 0)
        return fail(h, "bad decomposition properties");
    if (h->decompOn) return fail(h, "decomposition already set");
    CU(cudaSetDevice(h->cfg.device));
    const int nC = h->nCells, nI = h->nInternal;
    // smoothing-operator slots per cell, in cell-face order, empty faces left out; mesh.cellCells() alongside
    std::vector<int> off((size_t)nC + 1, 0), nb, info;
    std::vector<double> w, A;
    h->ccOffHost.assign((size_t)nC + 1, 0);
    h->ccIdsHost.clear();
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    for (int c = 0; c < nC; ++c) {
        for (int j = h->cfOffHost[c]; j < h->cfOffHost[c + 1]; ++j) {
            const int f = h->cfHost[j];
            const double* S = &h->SfHost[3 * (size_t)f];
            const double* C = &h->CfHost[3 * (size_t)f];
            const double magS = std::sqrt(dot(S, S));
            if (f < nI) {  // surfaceInterpolation::makeWeights
                const int o = h->ownerHost[f], n = h->neighbourHost[f];
                const double* cP = &h->ccHost[3 * (size_t)o];
                const double* cN = &h->ccHost[3 * (size_t)n];
                const double dO[3] = {C[0] - cP[0], C[1] - cP[1], C[2] - cP[2]}, dN[3] = {cN[0] - C[0], cN[1] - C[1], cN[2] - C[2]};
                const double sO = std::fabs(dot(S, dO)), sN = std::fabs(dot(S, dN));
                nb.push_back(o == c ? n : o);
                info.push_back(o == c ? DF_OWNER : DF_NEIGHBOUR);
                w.push_back(sN / (sO + sN));
                A.push_back(magS);
                h->ccIdsHost.push_back(o == c ? n : o);
                continue;
            }
            const int bfi = f - nI;
            int patch = -1;
            for (int p = 0; p < h->nPatches; ++p)
                if (f >= h->patchStart[p] && f < h->patchStart[p] + h->patchSize[p]) patch = p;
            const int kind = h->patchKind[patch];
            if (kind == UGF_PATCH_EMPTY) continue;
            if (kind == UGF_PATCH_CYCLIC) {  // coupled patch, linear weights from both sides' face-normal distances
                const int nf = h->patchStart[h->patchPartnerHost[patch]] + (f - h->patchStart[patch]);
                const int q = h->ownerHost[nf];
                const double* Sn = &h->SfHost[3 * (size_t)nf];
                const double* Cn = &h->CfHost[3 * (size_t)nf];
                const double An = std::sqrt(dot(Sn, Sn));
                const double* cP = &h->ccHost[3 * (size_t)c];
                const double* cQ = &h->ccHost[3 * (size_t)q];
                const double di = ((C[0] - cP[0]) * S[0] + (C[1] - cP[1]) * S[1] + (C[2] - cP[2]) * S[2]) / magS;
                const double dni = ((Cn[0] - cQ[0]) * Sn[0] + (Cn[1] - cQ[1]) * Sn[1] + (Cn[2] - cQ[2]) * Sn[2]) / An;
                nb.push_back(q);
                info.push_back(DF_BOUNDARY | (bfi << 2));
                w.push_back(dni / (di + dni));
            } else {  // zero gradient (wall, patch, processor) or symmetry
                nb.push_back(c);
                info.push_back((kind == UGF_PATCH_SYMMETRY ? DF_SYMMETRY : DF_BOUNDARY) | (bfi << 2));
                w.push_back(1.0);
            }
            A.push_back(magS);
        }
        off[c + 1] = (int)nb.size();
        h->ccOffHost[c + 1] = (int32_t)h->ccIdsHost.size();
    }
    const int W = KN_NACC + h->nSpecies;
    int *dOff, *dNb, *dInfo;
    double *dW, *dA, *dCc;
    if (dalloc(h, &dOff, off.size()) || dalloc(h, &dNb, nb.size()) || dalloc(h, &dInfo, info.size()) || dalloc(h, &dW, w.size()) ||
        dalloc(h, &dA, A.size()) || dalloc(h, &dCc, 3 * (size_t)nC) || dalloc(h, &h->dKnAcc, (size_t)nC * W) ||
        dalloc(h, &h->dKnF[0], (size_t)nC * KN_NF) || dalloc(h, &h->dKnF[1], (size_t)nC * KN_NF) || dalloc(h, &h->dKnK[0], (size_t)nC * 4) ||
        dalloc(h, &h->dKnK[1], (size_t)nC * 4))
        return 1;
    h->decompOwned = {dOff, dNb, dInfo, dW, dA, dCc, h->dKnAcc, h->dKnF[0], h->dKnF[1], h->dKnK[0], h->dKnK[1]};
    if (upload(h, dOff, off.data(), off.size()) || upload(h, dNb, nb.data(), nb.size()) || upload(h, dInfo, info.data(), info.size()) ||
        upload(h, dW, w.data(), w.size()) || upload(h, dA, A.data(), A.size()) || upload(h, dCc, h->ccHost.data(), 3 * (size_t)nC))
        return 1;
    CU(cudaMemsetAsync(h->dKnAcc, 0, sizeof(double) * (size_t)nC * W, h->stream));
    CU(cudaMemsetAsync(h->dKnK[0], 0, sizeof(double) * (size_t)nC * 4, h->stream));
    CU(cudaMemsetAsync(h->dKnK[1], 0, sizeof(double) * (size_t)nC * 4, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->dgeom.nCells = nC; h->dgeom.off = dOff; h->dgeom.nb = dNb; h->dgeom.info = dInfo; h->dgeom.w = dW; h->dgeom.A = dA;
    h->dgeom.bfS = h->dBfS; h->dgeom.cc = dCc; h->dgeom.vol = h->dVol;
    h->dec = *d;
    h->decompOn = true;
    h->decTimeSteps = 0;
    h->decTimeAv = 0;
    h->knCur = 0;
    return 0;
}

int ugf_set_macro_interpolation(ugf_handle* h, const ugf_cell_point* cp) {
    if (!h || !h->meshSet || !cp) return fail(h, "mesh not set");
    if (h->interpSet) return fail(h, "macro interpolation already set");
    if (h->hasProcessor) return fail(h, "macroInterpolation on a decomposed mesh is not supported (point values across processor patches need a halo)");
    CU(cudaSetDevice(h->cfg.device));
    const size_t nC = (size_t)h->nCells, nP = (size_t)cp->nPoints, nT = (size_t)cp->tetOffsets[nC], nW = (size_t)cp->pointCellOffsets[nP];
    double *dPts = nullptr, *dW = nullptr, *dNrm = nullptr, *dCc = nullptr, *dCellF = nullptr, *dPointF = nullptr;
    int *dTo = nullptr, *dTp = nullptr, *dPo = nullptr, *dPc = nullptr;
    if (dalloc(h, &dPts, 3 * nP) || dalloc(h, &dW, nW) || dalloc(h, &dNrm, 3 * nP) || dalloc(h, &dCc, 3 * nC) || dalloc(h, &dCellF, nC * NIF) ||
        dalloc(h, &dPointF, nP * NIF) || dalloc(h, &dTo, nC + 1) || dalloc(h, &dTp, 3 * nT) || dalloc(h, &dPo, nP + 1) || dalloc(h, &dPc, nW))
        return 1;
    for (void* p : {(void*)dPts, (void*)dW, (void*)dNrm, (void*)dCc, (void*)dCellF, (void*)dPointF, (void*)dTo, (void*)dTp, (void*)dPo, (void*)dPc})
        h->interpOwned.push_back(p);
    h->interpSet = true;
    if (upload(h, dPts, cp->points, 3 * nP) || upload(h, dW, cp->pointWeights, nW) || upload(h, dNrm, cp->pointNormals, 3 * nP) ||
        upload(h, dCc, h->ccHost.data(), 3 * nC) || upload(h, dTo, cp->tetOffsets, nC + 1) || upload(h, dTp, cp->tetPoints, 3 * nT) ||
        upload(h, dPo, cp->pointCellOffsets, nP + 1) || upload(h, dPc, cp->pointCells, nW))
        return 1;
    CU(cudaStreamSynchronize(h->stream));
    h->interp.points = dPts; h->interp.tetOff = dTo; h->interp.tetPts = dTp; h->interp.pcOff = dPo; h->interp.pc = dPc; h->interp.pw = dW;
    h->interp.pnormal = dNrm; h->interp.cc = dCc; h->interp.cellF = dCellF; h->interp.pointF = dPointF; h->interp.nPoints = (int)nP;
    return 0;
}

int ugf_decompose(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (!h->decompOn) return fail(h, "no decomposition model set");
    return do_decompose(h);
}

int ugf_download_decomposition(ugf_handle* h, int32_t* id, double* kn) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (!h->decompOn) return fail(h, "no decomposition model set");
    if (id) CU(countedMemcpyAsync(h, id, h->dCollId, sizeof(int32_t) * (size_t)h->nCells, cudaMemcpyDeviceToHost, h->stream));
    if (kn) CU(countedMemcpyAsync(h, kn, h->dKnK[h->knCur], sizeof(double) * (size_t)h->nCells * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ugf_set_time_index(ugf_handle* h, int64_t index) {
    if (!h) return 1;
    if (index < 0) return fail(h, "time index must be >= 0");
    h->step = index;
    return 0;
}

// ---- state checkpoint (ugf_state_*): flat array of doubles, same layout as the oracle's ---------------------------------
// header [8]: magic, version, nCells, nSpecies, nBFaces, decomposition on, inlet velocity doubles, reserved
// scalars [6]: step, timeAvCounter, nAvTimeSteps, sampleCounter, decTimeSteps, decTimeAv
// sigmaTcRMax [nC], collModelId [nC], maxProb [nC], qPrev [3 nC], sPrev [6 nC], acc [16 nC], accSpecies [nS nC],
// bacc [16 nB], then if decomposition: knAcc [(7 + nS) nC], knFields [4 nC]; then the pressure inlets' face velocities
namespace {
constexpr double STATE_MAGIC = 1431783237.0;  // "UGFS"
long long inlet_velocity_doubles(const ugf_handle* h) {
    long long n = 0;
    for (const InflowHost& f : h->inflows) if (f.pressureInlet) n += 3LL * f.dev.nFaces + (f.wang ? (long long)WANG_NSUM * f.dev.nFaces + 1 : 0) + (f.outlet ? (long long)f.nSlots + 2LL * f.dev.nFaces : 0) + (f.massFlow ? (long long)f.nSlots : 0);
    return n;
}
long long state_doubles(const ugf_handle* h) {
    const long long nC = h->nCells, nS = h->nSpecies, nB = h->nBFaces;
    long long n = 8 + 6 + nC * (1 + 1 + 1 + 3 + 6 + NACC + nS) + nB * UGF_NBM;
    if (h->decompOn) n += nC * (KN_NACC + nS) + nC * 4;
    return n + inlet_velocity_doubles(h) + (h->dAccI ? nC * nS * UGF_NINT : 0);  // internal-mode accumulators: last block, header word 7
}
}  // namespace

int ugf_state_size(ugf_handle* h, int64_t* n) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    *n = state_doubles(h);
    return 0;
}

int ugf_state_save(ugf_handle* h, double* buf, int64_t nDoubles) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (nDoubles != state_doubles(h)) return fail(h, "state buffer has the wrong size");
    CU(cudaSetDevice(h->cfg.device));
    const size_t nC = (size_t)h->nCells, nS = (size_t)h->nSpecies, nB = (size_t)h->nBFaces;
    double* p = buf;
    const double hdr[8] = {STATE_MAGIC, 1.0, (double)nC, (double)nS, (double)nB, h->decompOn ? 1.0 : 0.0, (double)inlet_velocity_doubles(h), h->dAccI ? 1.0 : 0.0};
    std::copy(hdr, hdr + 8, p); p += 8;
    const double sc[6] = {(double)h->step, h->timeAvCounter, (double)h->nAvTimeSteps, (double)h->sampleCounter, (double)h->decTimeSteps, h->decTimeAv};
    std::copy(sc, sc + 6, p); p += 6;
    auto d2h = [&](const double* src, size_t n) -> int {
        if (n) CU(countedMemcpyAsync(h, p, src, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
        p += n;
        return 0;
    };
    if (d2h(h->dSigma, nC)) return 1;
    std::vector<int> ids(nC);
    CU(countedMemcpyAsync(h, ids.data(), h->dCollId, sizeof(int) * nC, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (size_t c = 0; c < nC; ++c) p[c] = ids[c];
    p += nC;
    if (d2h(h->dMaxProb, nC) || d2h(h->dQPrev, 3 * nC) || d2h(h->dSPrev, 6 * nC)) return 1;
    double* accAt = p;
    if (d2h(h->dAcc, NACC * nC)) return 1;
    double* accSAt = p;
    if (h->dAccS) { if (d2h(h->dAccS, nS * nC)) return 1; } else p += nS * nC;
    if (d2h(h->dBacc, UGF_NBM * nB)) return 1;
    if (h->decompOn) { if (d2h(h->dKnAcc, (KN_NACC + nS) * nC) || d2h(h->dKnK[h->knCur], 4 * nC)) return 1; }
    for (const InflowHost& f : h->inflows) {
        if (!f.pressureInlet) continue;
        if (d2h(f.dev.faceVel, 3 * (size_t)f.dev.nFaces)) return 1;
        if (f.massFlow && d2h(f.dev.faceN, (size_t)f.nSlots)) return 1;
        if (f.wang) { if (d2h(f.dev.wangSums, (size_t)WANG_NSUM * f.dev.nFaces)) return 1; *p++ = f.wangSteps; }
        if (f.outlet && (d2h(f.dev.faceN, (size_t)f.nSlots) || d2h(f.dev.faceT, 2 * (size_t)f.dev.nFaces))) return 1;
    }
    if (h->dAccI && d2h(h->dAccI, (size_t)UGF_NINT * nS * nC)) return 1;
    CU(cudaStreamSynchronize(h->stream));
    if (!h->dAccS) for (size_t c = 0; c < nC; ++c) accSAt[c] = accAt[c * NACC + 8];  // one species: nParcelsXnParticle = slot 8
    return 0;
}

int ugf_state_load(ugf_handle* h, const double* buf, int64_t nDoubles) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (nDoubles != state_doubles(h) || nDoubles < 14) return fail(h, "state buffer has the wrong size for this set-up");
    const size_t nC = (size_t)h->nCells, nS = (size_t)h->nSpecies, nB = (size_t)h->nBFaces;
    if (buf[0] != STATE_MAGIC || buf[1] != 1.0) return fail(h, "not a ugf state buffer (magic / version)");
    if (buf[2] != (double)nC || buf[3] != (double)nS || buf[4] != (double)nB || buf[5] != (h->decompOn ? 1.0 : 0.0) ||
        buf[6] != (double)inlet_velocity_doubles(h) || buf[7] != (h->dAccI ? 1.0 : 0.0))
        return fail(h, "state buffer was written for another mesh / species / model set-up");
    CU(cudaSetDevice(h->cfg.device));
    const double* p = buf + 8;
    h->step = (long long)p[0]; h->timeAvCounter = p[1]; h->nAvTimeSteps = (long long)p[2]; h->sampleCounter = (int)p[3];
    h->decTimeSteps = (int)p[4]; h->decTimeAv = p[5];
    p += 6;
    auto h2d = [&](double* dst, size_t n) -> int {
        if (n && dst) CU(countedMemcpyAsync(h, dst, p, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
        p += n;
        return 0;
    };
    if (h2d(h->dSigma, nC)) return 1;
    std::vector<int> ids(nC);
    for (size_t c = 0; c < nC; ++c) ids[c] = (int)p[c];
    p += nC;
    CU(countedMemcpyAsync(h, h->dCollId, ids.data(), sizeof(int) * nC, cudaMemcpyHostToDevice, h->stream));
    if (h2d(h->dMaxProb, nC) || h2d(h->dQPrev, 3 * nC) || h2d(h->dSPrev, 6 * nC) || h2d(h->dAcc, NACC * nC) || h2d(h->dAccS, nS * nC) ||
        h2d(h->dBacc, UGF_NBM * nB))
        return 1;
    if (h->decompOn) { if (h2d(h->dKnAcc, (KN_NACC + nS) * nC) || h2d(h->dKnK[h->knCur], 4 * nC)) return 1; }
    for (InflowHost& f : h->inflows) {
        if (!f.pressureInlet) continue;
        const double* vL = p;
        if (h2d(f.dev.faceVel, 3 * (size_t)f.dev.nFaces)) return 1;
        if (f.massFlow) {  // the insertion bound follows from the velocities and number densities just loaded: the slots' counts
            const double* nD = p;
            if (h2d(f.dev.faceN, (size_t)f.nSlots)) return 1;
            double bound = 0.0;
            for (int lf = 0; lf < f.dev.nFaces; ++lf)
                for (int i = 0; i < f.dev.nTypeIds; ++i) {
                    const double* S = &h->SfHost[3 * ((size_t)h->patchStart[f.patch] + lf)];
                    const double fA = std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
                    const double cmp = std::sqrt(2.0 * kB * f.dev.Ttr / h->spHost[i].mass);
                    const double w = h->cwfHost.empty() ? 1.0 : h->cwfHost[f.slotCell[(size_t)lf * f.dev.nTypeIds + i]];
                    const double sCos = std::min(5.0, -(vL[3 * lf] * S[0] + vL[3 * lf + 1] * S[1] + vL[3 * lf + 2] * S[2]) / fA / cmp);
                    bound += 1.000001 * fA * nD[(size_t)lf * f.dev.nTypeIds + i] * h->cfg.deltaT * cmp * (std::exp(-(sCos * sCos)) + std::sqrt(PI) * sCos * (1 + std::erf(sCos)))
                             / (2.0 * std::sqrt(PI) * h->cfg.nParticle * w);
                }
            f.maxInsert = (long long)std::ceil(bound) + f.nSlots + 2;
        }
        if (f.wang) { if (h2d(f.dev.wangSums, (size_t)WANG_NSUM * f.dev.nFaces)) return 1; f.wangSteps = *p++; }
        if (f.outlet && (h2d(f.dev.faceN, (size_t)f.nSlots) || h2d(f.dev.faceT, 2 * (size_t)f.dev.nFaces))) return 1;
    }
    if (h->dAccI && h2d(h->dAccI, (size_t)UGF_NINT * nS * nC)) return 1;
    CU(cudaStreamSynchronize(h->stream));  // ids and the caller's buffer may go away
    h->momValid = false;
    return 0;
}

int ugf_end_step(ugf_handle* h) {
    if (!h) return 1;
    h->step++;
    h->stepOpen = false;
    return 0;
}

int ugf_finish_step(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (do_sort(h)) return 1;
    const bool fuseAcc = next_step_samples(h);
    if (run_cell_kernel(h, true, true, fuseAcc, bgk_active(h))) return 1;
    if (dsmc_active(h) && run_ntc_kernel(h)) return 1;
    if (bgk_active(h) && run_bgk_kernel(h)) return 1;
    if (update_inlet_velocities(h)) return 1;
    if (do_accumulate(h, fuseAcc)) return 1;
    if (do_decompose(h)) return 1;
    return ugf_end_step(h);
}

int ugf_step(ugf_handle* h, int32_t nSteps) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (h->hasProcessor) return fail(h, "ugf_step is single-rank: the mesh has processor patches (drive the phases and ugf_migrate_* instead)");
    CU(cudaSetDevice(h->cfg.device));
    for (int s = 0; s < nSteps; ++s) {
        if (refresh_n(h)) return 1;
        if (zero_step_counters(h)) return 1;
        const bool last = (s == nSteps - 1);
        if (last) CU(cudaEventRecord(h->ev[0], h->stream));
        if (!h->inflows.empty()) { if (do_inflow(h)) return 1; } else h->newFrom = h->nUpper;
        if (last) CU(cudaEventRecord(h->ev[1], h->stream));
        if (do_move(h, 0, false)) return 1;
        h->inflowDone = false;
        if (last) CU(cudaEventRecord(h->ev[2], h->stream));
        if (do_sort(h)) return 1;
        if (last) CU(cudaEventRecord(h->ev[3], h->stream));
        const bool fuseAcc = next_step_samples(h);
        if (run_cell_kernel(h, true, true, fuseAcc, bgk_active(h))) return 1;
        if (last) CU(cudaEventRecord(h->ev[4], h->stream));
        if (dsmc_active(h) && run_ntc_kernel(h)) return 1;
        if (last) CU(cudaEventRecord(h->ev[5], h->stream));
        if (bgk_active(h) && run_bgk_kernel(h)) return 1;
        if (update_inlet_velocities(h)) return 1;
        if (last) CU(cudaEventRecord(h->ev[6], h->stream));
        if (do_accumulate(h, fuseAcc)) return 1;
        if (do_decompose(h)) return 1;
        if (last) CU(cudaEventRecord(h->ev[7], h->stream));
        h->step++;
    }
    h->timingValid = true;
    return 0;
}

// ---- migration --------------------------------------------------------------------------------------------

int ugf_migrate_counts(ugf_handle* h, int64_t* counts) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    std::vector<int> c(std::max(h->nPatches, 1), 0);
    CU(countedMemcpyAsync(h, c.data(), h->dMigCount, sizeof(int) * h->nPatches, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int p = 0; p < h->nPatches; ++p) counts[p] = c[p];
    return 0;
}

int ugf_migrate_pack(ugf_handle* h, int32_t patch, double** devBuf, int64_t* nPacked) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (patch < 0 || patch >= h->nPatches || h->patchKind[patch] != UGF_PATCH_PROCESSOR) return fail(h, "pack on a non-processor patch");
    int count = 0;
    CU(countedMemcpyAsync(h, &count, h->dMigCount + patch, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (count > h->packCap[patch]) {
        if (h->packBuf[patch]) CU(cudaFree(h->packBuf[patch]));
        h->packCap[patch] = std::max<long long>(2LL * count, 1024);
        if (dalloc(h, &h->packBuf[patch], (size_t)h->packCap[patch] * UGF_MIGRATE_STRIDE)) return 1;
    }
    *devBuf = h->packBuf[patch];
    *nPacked = count;
    if (count == 0) return 0;
    const unsigned nb = grid_for(h->nUpper, 1024);
    ParcelBuf P = h->buf[h->cur];
    mig_count_kernel<<<nb, 1024, 0, h->stream>>>(h->mesh, P.cell, h->dN, patch, h->dMigBlock);
    LAUNCHED();
    scan_top_kernel<<<1, SCAN_THREADS, 0, h->stream>>>(h->dMigBlock, (int)nb, h->dBlockSums, nullptr, 0x7fffffffLL, nullptr);
    LAUNCHED();
    double* buf = h->packBuf[patch];
    dispatch(h, [&](auto R, auto M) {
        mig_pack_kernel<decltype(R)::value, decltype(M)::value><<<nb, 1024, 0, h->stream>>>(h->mesh, P, h->dSf, h->dWq, h->dN, patch, h->dMigBlock, buf);
    });
    LAUNCHED();
    CU(cudaMemsetAsync(h->dMigCount + patch, 0, sizeof(int), h->stream));
    return 0;
}

int ugf_migrate_unpack(ugf_handle* h, int32_t patch, const double* devBuf, int64_t n) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (patch < 0 || patch >= h->nPatches || h->patchKind[patch] != UGF_PATCH_PROCESSOR) return fail(h, "unpack on a non-processor patch");
    if (n <= 0) return 0;
    if (h->nPending && h->nExact) return fail(h, "ugf_migrate_unpack must follow ugf_move");
    if (!h->nExact) {  // the exact-count path appends at a host-known index: wait for the length now
        if (refresh_n(h)) return 1;
        h->nExact = true;
        h->recvStart = h->nUpper;
    }
    if (h->nUpper + n > h->capacity) return fail(h, "parcelCapacity too small for the received parcels");
    ParcelBuf P = h->buf[h->cur];
    const long long base = h->nUpper;
    dispatch(h, [&](auto R, auto M) {
        mig_unpack_kernel<decltype(R)::value, decltype(M)::value><<<grid_for(n, 256), 256, 0, h->stream>>>(h->mesh, P, h->dSf, h->dWq, base, n, patch, devBuf, h->dErr);
    });
    LAUNCHED();
    h->nUpper += n;
    h->appendBound += n;
    const long long nn = h->nUpper;
    // the host mirror is exact here (no insertion since the last refresh), so the device length follows it
    CU(countedMemcpyAsync(h, h->dN, &nn, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->occValid = false; h->momValid = false;
    return 0;
}

int ugf_move_received(ugf_handle* h) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (h->recvStart < 0) return fail(h, "ugf_move_received before ugf_move");
    if (h->recvMoved) { h->recvMoved = false; return 0; }  // the fused receive kernel has tracked them already
    if (h->slotRound) {
        // received through slots: the exact start lives on the device (dRecvStart); launch over everything appended
        // since the step's move and let the kernel skip what precedes it
        if (do_move(h, h->nAtMove, true)) return 1;
        return 0;
    }
    if (do_move(h, h->recvStart, true)) return 1;
    h->recvStart = h->nUpper;
    return 0;
}

int ugf_migrate_pack_slots(ugf_handle* h, double* devSend, int64_t slotCapacity) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (h->migSlots.nProc > MIG_MAXP) return fail(h, "too many processor patches for the slot path");
    if (h->migSlots.nProc == 0) return 0;
    if (slotCapacity < 1) return fail(h, "slotCapacity must be positive");
    MigDst dst{};
    const MigSlots& ms = h->migSlots;
    for (int k = 0; k < ms.nProc; ++k) dst.slot[k] = devSend + (size_t)k * (size_t)(slotCapacity + 1) * UGF_MIGRATE_STRIDE;
    return pack_slots_to(h, dst, slotCapacity);
}

int ugf_peer_alloc(ugf_handle* h, int64_t bytes, void** devPtr, unsigned char* ipcHandle64) {
    if (!h || !devPtr || !ipcHandle64 || bytes <= 0) return fail(h, "ugf_peer_alloc: bad argument");
    CU(cudaSetDevice(h->cfg.device));
    void* p = nullptr;
    CU(cudaMalloc(&p, (size_t)bytes));
    CU(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t hd;
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(cudaIpcGetMemHandle(&hd, p));
    std::memcpy(ipcHandle64, &hd, 64);
    h->peerOwned.push_back(p);
    *devPtr = p;
    return 0;
}

int ugf_peer_open(ugf_handle* h, const unsigned char* ipcHandle64, void** devPtr) {
    if (!h || !devPtr || !ipcHandle64) return fail(h, "ugf_peer_open: bad argument");
    CU(cudaSetDevice(h->cfg.device));
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, ipcHandle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peerOpened.push_back(p);
    *devPtr = p;
    return 0;
}

int ugf_migrate_pack_peer(ugf_handle* h, double* const* dstSlots, uint64_t* const* dstFlags, int64_t slotCapacity, uint64_t epoch) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (h->migSlots.nProc > MIG_MAXP) return fail(h, "too many processor patches for the slot path");
    if (h->migSlots.nProc == 0) return 0;
    if (slotCapacity < 1) return fail(h, "slotCapacity must be positive");
    MigDst dst{};
    MigFlags fl{};
    for (int k = 0; k < h->migSlots.nProc; ++k) {
        dst.slot[k] = dstSlots[k];
        fl.flag[k] = reinterpret_cast<unsigned long long*>(dstFlags[k]);
    }
    bool signalled = false;
    if (pack_slots_to(h, dst, slotCapacity, &fl, (unsigned long long)epoch, &signalled)) return 1;
    if (!signalled) {  // the search-based pack (slots larger than the migrant lists): flags from their own launch
        mig_signal_kernel<<<1, 32, 0, h->stream>>>(fl, h->migSlots.nProc, (unsigned long long)epoch);
        LAUNCHED();
        if (h->migFused) CU(cudaMemsetAsync(h->dInflight, 0, sizeof(unsigned long long), h->stream));  // the list pack does this itself
    }
    return 0;
}

int ugf_migrate_unpack_peer(ugf_handle* h, const double* devRecv, const uint64_t* devFlags, int64_t slotCapacity, uint64_t epoch) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (h->migSlots.nProc == 0) return 0;
    if (!h->migFused) {
        mig_wait_kernel<<<1, 32, 0, h->stream>>>(reinterpret_cast<const unsigned long long*>(devFlags), h->migSlots.nProc, (unsigned long long)epoch, h->dErr);
        LAUNCHED();
        return ugf_migrate_unpack_slots(h, devRecv, slotCapacity);
    }
    // one launch: wait for the flags, append the records, continue their tracks, publish the new length
    if (h->migSlots.nProc > MIG_MAXP) return fail(h, "too many processor patches for the slot path");
    if (h->nExact && h->nUpper + (long long)h->migSlots.nProc * slotCapacity > h->capacity)
        return fail(h, "parcelCapacity too small for the migration slots (needs room for nProcPatches x slotCapacity parcels)");
    h->slotRound = true;
    MoveArgs a = make_move_args(h, 0, true);
    a.dBegin = nullptr;
    RecvArgs ra{};
    ra.ms = h->migSlots; ra.recv = devRecv; ra.flags = reinterpret_cast<const unsigned long long*>(devFlags); ra.epoch = (unsigned long long)epoch;
    ra.slotCapacity = slotCapacity; ra.capacity = h->capacity; ra.dN = h->dN; ra.dRecvStart = h->dRecvStart; ra.done = h->dMigDone; ra.errFlag = h->dErr;
    const DevParams prm = h->prm;
    const dim3 grid(grid_for(slotCapacity, 256), h->migSlots.nProc);
    dispatch(h, [&](auto R, auto M) {
        constexpr bool r = decltype(R)::value, mm = decltype(M)::value;
        if (h->moveNF == 6) mig_recv_move_kernel<r, mm, 6><<<grid, 256, 0, h->stream>>>(prm, a, ra);
        else if (h->moveNF == 4 && h->mesh.rec2d) mig_recv_move_kernel<r, mm, NF_REC2D><<<grid, 256, 0, h->stream>>>(prm, a, ra);
        else if (h->moveNF == 4) mig_recv_move_kernel<r, mm, 4><<<grid, 256, 0, h->stream>>>(prm, a, ra);
        else mig_recv_move_kernel<r, mm, 0><<<grid, 256, 0, h->stream>>>(prm, a, ra);
    });
    LAUNCHED();
    h->nUpper = std::min<long long>(h->capacity, h->nUpper + (long long)h->migSlots.nProc * slotCapacity);
    h->appendBound += (long long)h->migSlots.nProc * slotCapacity;
    h->lastSlotCapacity = slotCapacity;
    h->recvMoved = true;  // ugf_move_received has nothing left to do for this round
    h->occValid = false; h->momValid = false;
    return 0;
}

int ugf_migrate_unpack_slots(ugf_handle* h, const double* devRecv, int64_t slotCapacity) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    if (h->migSlots.nProc > MIG_MAXP) return fail(h, "too many processor patches for the slot path");
    if (h->migSlots.nProc == 0) return 0;
    if (h->nExact && h->nUpper + (long long)h->migSlots.nProc * slotCapacity > h->capacity)
        return fail(h, "parcelCapacity too small for the migration slots (needs room for nProcPatches x slotCapacity parcels)");
    ParcelBuf P = h->buf[h->cur];
    const MigSlots ms = h->migSlots;
    dispatch(h, [&](auto R, auto M) {
        mig_unpack_all_kernel<decltype(R)::value, decltype(M)::value><<<dim3(grid_for(slotCapacity, 256), ms.nProc), 256, 0, h->stream>>>(
            h->mesh, ms, P, h->dSf, h->dWq, h->dN, h->capacity, devRecv, (long long)slotCapacity, h->dErr);
    });
    LAUNCHED();
    mig_commit_kernel<<<1, 1, 0, h->stream>>>(h->dN, h->dRecvStart, h->dInflight, devRecv, ms.nProc, (long long)slotCapacity, h->capacity, h->dErr);
    LAUNCHED();
    h->nUpper = std::min<long long>(h->capacity, h->nUpper + (long long)ms.nProc * slotCapacity);  // upper bound; overflow raises the device flag
    h->appendBound += (long long)ms.nProc * slotCapacity;
    h->slotRound = true;
    h->lastSlotCapacity = slotCapacity;
    h->occValid = false; h->momValid = false;
    return 0;
}

int ugf_migrate_inflight(ugf_handle* h, int64_t** dev) {
    if (!h) return 1;
    *dev = reinterpret_cast<int64_t*>(h->dInflight);
    return 0;
}

int ugf_stream(ugf_handle* h, void** s) {
    if (!h) return 1;
    *s = (void*)h->stream;
    return 0;
}

// ---- results ------------------------------------------------------------------------------------------------

int ugf_num_parcels(ugf_handle* h, int64_t* n) {
    if (!h) return 1;
    ugf_counters c;
    if (ugf_counters_get(h, &c)) return 1;
    *n = c.nParcels;
    return 0;
}

int ugf_counters_get(ugf_handle* h, ugf_counters* out) {
    if (!h) return 1;
    std::memset(out, 0, sizeof(*out));
    CU(cudaSetDevice(h->cfg.device));
    if (refresh_n(h)) return 1;
    DevCounters c;
    double tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h->buf[0].x) {
        CU(cudaMemsetAsync(h->dTot, 0, 8 * sizeof(double), h->stream));
        ParcelBuf P = h->buf[h->cur];
        const DevParams prm = h->prm;
        dispatch(h, [&](auto R, auto M) {
            totals_kernel<decltype(R)::value, decltype(M)::value><<<h->numSMs * 4, 256, 0, h->stream>>>(prm, P, h->dN, h->dTot);
        });
        LAUNCHED();
        if (h->dSpi) {
            internal_totals_kernel<<<h->numSMs * 4, 256, 0, h->stream>>>(prm, P, h->dN, h->dTot + 6);
            LAUNCHED();
        }
        CU(countedMemcpyAsync(h, tot, h->dTot, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(countedMemcpyAsync(h, &c, h->dCnt, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    int devErr = 0;
    CU(countedMemcpyAsync(h, &devErr, h->dErr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));  // same round trip as the counters
    CU(cudaStreamSynchronize(h->stream));
    if (devErr && check_device_error(h)) return 1;
    out->step = h->step;
    out->nParcels = (int64_t)tot[5];
    out->collisionCandidates = (int64_t)c.cand;
    out->collisions = (int64_t)c.coll;
    out->bgkRelaxations = (int64_t)c.bgk;
    out->inserted = (int64_t)c.inserted;
    out->deleted = (int64_t)c.deleted;
    out->migrated = (int64_t)c.migrated;
    out->wallHits = (int64_t)c.wallHits;
    out->stuck = (int64_t)c.stuck;
    out->cloned = (int64_t)c.cloned;
    out->weightDeleted = (int64_t)c.wdeleted;
    out->linearKineticEnergy = tot[0];
    out->rotationalEnergy = tot[1];
    out->vibrationalEnergy = tot[6];
    out->electronicEnergy = tot[7];
    out->momentum[0] = tot[2]; out->momentum[1] = tot[3]; out->momentum[2] = tot[4];
    return 0;
}

int ugf_download_parcels(ugf_handle* h, ugf_parcels* p) {
    if (!h || !h->buf[0].x) return fail(h, "no parcels uploaded");
    CU(cudaSetDevice(h->cfg.device));
    long long n = 0;
    CU(countedMemcpyAsync(h, &n, h->dN, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (p->n < n) return fail(h, "download buffer too small");
    const ParcelBuf& P = h->buf[h->cur];
    const size_t nb = (size_t)n;
    auto dl = [&](double* dst, const double* src) -> int {
        if (dst && nb) CU(countedMemcpyAsync(h, dst, src, nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        return 0;
    };
    if (dl(p->x, P.x) || dl(p->y, P.y) || dl(p->z, P.z) || dl(p->Ux, P.ux) || dl(p->Uy, P.uy) || dl(p->Uz, P.uz)) return 1;
    if (p->cell && nb) CU(countedMemcpyAsync(h, p->cell, P.cell, nb * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (p->ERot) {
        if (h->hasRot) { if (dl(p->ERot, P.erot)) return 1; }
        else std::fill(p->ERot, p->ERot + nb, 0.0);
    }
    std::vector<uint8_t> types;
    if (p->typeId) {
        if (h->multi && nb) {
            types.resize(nb);
            CU(countedMemcpyAsync(h, types.data(), P.type, nb, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    if (p->typeId) for (size_t i = 0; i < nb; ++i) p->typeId[i] = h->multi ? types[i] : 0;
    if (p->newParcel) std::fill(p->newParcel, p->newParcel + nb, 0);
    if (p->vibLevel) {
        if (h->hasVib && nb) {
            std::vector<unsigned long long> v(nb);
            CU(countedMemcpy(h, v.data(), P.vib, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nb; ++i)
                for (int m = 0; m < UGF_MAX_VIB_MODES; ++m) p->vibLevel[i * UGF_MAX_VIB_MODES + m] = (int32_t)((v[i] >> (16 * m)) & 0xFFFFull);
        } else std::fill(p->vibLevel, p->vibLevel + nb * UGF_MAX_VIB_MODES, 0);
    }
    if (p->ELevel) {
        if (h->hasElec && nb) {
            std::vector<uint8_t> e(nb);
            CU(countedMemcpy(h, e.data(), P.elev, nb, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nb; ++i) p->ELevel[i] = e[i];
        } else std::fill(p->ELevel, p->ELevel + nb, 0);
    }
    if (p->cellWeight) {  // implicit on the device: the factor of the parcel's cell (the previous field while an update is pending)
        std::vector<int> cells;
        const int* cp = p->cell;
        if (!cp && nb) {
            cells.resize(nb);
            CU(countedMemcpy(h, cells.data(), P.cell, nb * sizeof(int), cudaMemcpyDeviceToHost));
            cp = cells.data();
        }
        const std::vector<double>& w = h->cwfDirty ? h->cwfHostPrev : h->cwfHost;
        for (size_t i = 0; i < nb; ++i) p->cellWeight[i] = (w.empty() || cp[i] < 0) ? 1.0 : w[cp[i]];
    }
    if (p->radialWeight) {  // implicit on the device: RWF(position), or RWF(cell centre) before the first move of such an upload
        if (!h->cfg.axisymmetric) std::fill(p->radialWeight, p->radialWeight + nb, 1.0);
        else {
            std::vector<double> y, z;
            std::vector<int> cells;
            const double *yp = p->y, *zp = p->z;
            const int* cp = p->cell;
            if (!yp && nb) { y.resize(nb); CU(countedMemcpy(h, y.data(), P.y, nb * sizeof(double), cudaMemcpyDeviceToHost)); yp = y.data(); }
            if (!zp && nb) { z.resize(nb); CU(countedMemcpy(h, z.data(), P.z, nb * sizeof(double), cudaMemcpyDeviceToHost)); zp = z.data(); }
            if (!cp && nb) { cells.resize(nb); CU(countedMemcpy(h, cells.data(), P.cell, nb * sizeof(int), cudaMemcpyDeviceToHost)); cp = cells.data(); }
            for (size_t i = 0; i < nb; ++i)
                p->radialWeight[i] = (h->rwfCentre && cp[i] >= 0) ? h->cellRwfHost[cp[i]]
                                                                  : 1.0 + (h->cfg.maxRWF - 1.0) * std::sqrt(yp[i] * yp[i] + zp[i] * zp[i]) / h->cfg.radialExtent;
        }
    }
    p->n = n;
    return 0;
}

int ugf_download_cell_occupancy(ugf_handle* h, int32_t* off, int32_t* ids) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (!h->occValid) return fail(h, "cell occupancy not built (call ugf_sort)");
    CU(countedMemcpyAsync(h, off, h->dOff, sizeof(int) * ((size_t)h->nCells + 1), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (ids) {
        const int n = off[h->nCells];
        if (h->occIdentity) { for (int i = 0; i < n; ++i) ids[i] = i; }
        else if (n) {
            CU(countedMemcpyAsync(h, ids, h->dPerm, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        }
    }
    return 0;
}

int ugf_download_cell_moments(ugf_handle* h, double* m) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (!h->momValid) return fail(h, "cell moments not sampled");
    CU(countedMemcpyAsync(h, m, h->dMom, sizeof(double) * (size_t)h->nCells * h->nSpecies * UGF_NMOM, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ugf_download_cell_state(ugf_handle* h, double* s, double* mp, double* q, double* sp) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    const size_t nC = (size_t)h->nCells;
    if (s) CU(countedMemcpyAsync(h, s, h->dSigma, sizeof(double) * nC, cudaMemcpyDeviceToHost, h->stream));
    if (mp) CU(countedMemcpyAsync(h, mp, h->dMaxProb, sizeof(double) * nC, cudaMemcpyDeviceToHost, h->stream));
    if (q) CU(countedMemcpyAsync(h, q, h->dQPrev, sizeof(double) * 3 * nC, cudaMemcpyDeviceToHost, h->stream));
    if (sp) CU(countedMemcpyAsync(h, sp, h->dSPrev, sizeof(double) * 6 * nC, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ugf_download_fields(ugf_handle* h, double* cellF, double* wallF, int32_t reset) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    CU(cudaSetDevice(h->cfg.device));
    const int nC = h->nCells, nB = h->nBFaces;
    double* tmp = nullptr;
    const size_t need = std::max((size_t)nC * UGF_NFIELD, (size_t)std::max(nB, 1) * UGF_NWALLFIELD);
    if (dalloc(h, &tmp, need)) return 1;
    const double t = h->timeAvCounter;
    if (cellF) {
        derive_cells_kernel<<<grid_for(nC, 256), 256, 0, h->stream>>>(h->prm, nC, h->dAcc, h->dAccS, h->dVol, h->dBbMin, h->dBbMax,
                                                                      h->subLevelsAllOne ? nullptr : h->dSubLevels, t, (double)h->nAvTimeSteps, tmp, h->dAccI);
        LAUNCHED();
        CU(countedMemcpyAsync(h, cellF, tmp, sizeof(double) * (size_t)nC * UGF_NFIELD, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (wallF && nB > 0) {
        derive_walls_kernel<<<grid_for(nB, 256), 256, 0, h->stream>>>(h->prm, h->mesh, h->dBacc, h->dBfS, t, tmp);
        LAUNCHED();
        CU(countedMemcpyAsync(h, wallF, tmp, sizeof(double) * (size_t)nB * UGF_NWALLFIELD, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(cudaFree(tmp));
    if (reset) {
        CU(cudaMemsetAsync(h->dAcc, 0, sizeof(double) * (size_t)nC * NACC, h->stream));
        if (h->dAccS) CU(cudaMemsetAsync(h->dAccS, 0, sizeof(double) * (size_t)nC * h->nSpecies, h->stream));
        CU(cudaMemsetAsync(h->dBacc, 0, sizeof(double) * std::max<size_t>((size_t)nB * UGF_NBM, 1), h->stream));
        if (h->dAccI) CU(cudaMemsetAsync(h->dAccI, 0, sizeof(double) * (size_t)nC * h->nSpecies * UGF_NINT, h->stream));
        h->timeAvCounter = 0; h->nAvTimeSteps = 0;
    }
    return 0;
}

int ugf_download_internal_accumulators(ugf_handle* h, double* accInt) {
    if (!h || !h->meshSet || !accInt) return fail(h, "mesh not set");
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->nCells * h->nSpecies * UGF_NINT;
    if (h->dAccI) {
        CU(countedMemcpyAsync(h, accInt, h->dAccI, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    } else {
        std::fill(accInt, accInt + n, 0.0);
    }
    return 0;
}

int ugf_download_accumulators(ugf_handle* h, double* acc, double* accS, double* timeAv, int64_t* nAv) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    CU(cudaSetDevice(h->cfg.device));
    const size_t nC = (size_t)h->nCells, nS = (size_t)h->nSpecies;
    std::vector<double> tmp;
    if (accS && !h->dAccS && !acc) tmp.resize(nC * NACC);
    double* a = acc ? acc : (tmp.empty() ? nullptr : tmp.data());
    if (a) CU(countedMemcpyAsync(h, a, h->dAcc, sizeof(double) * nC * NACC, cudaMemcpyDeviceToHost, h->stream));
    if (accS && h->dAccS) CU(countedMemcpyAsync(h, accS, h->dAccS, sizeof(double) * nC * nS, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (accS && !h->dAccS) for (size_t c = 0; c < nC; ++c) accS[c] = a[c * NACC + 8];  // one species: slot 8
    if (timeAv) *timeAv = h->timeAvCounter;
    if (nAv) *nAv = h->nAvTimeSteps;
    return 0;
}

int ugf_set_face_tracker(ugf_handle* h, int32_t n, const int32_t* faces) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->dSlotTrack); cudaFree(h->dBfTrack); cudaFree(h->dFt);
    h->dSlotTrack = nullptr; h->dBfTrack = nullptr; h->dFt = nullptr;
    h->nTracked = 0;
    if (n <= 0 || !faces) return 0;
    if (h->cfg.axisymmetric && h->hasProcessor)  // the weight a migrating parcel carries is CWF x RWF; the tracker needs CWF alone
        return fail(h, "face tracker on a decomposed axisymmetric case is not supported");
    std::vector<int> faceIdx((size_t)h->nFaces, 0);  // k + 1 of a tracked face
    for (int k = 0; k < n; ++k) {
        if (faces[k] < 0 || faces[k] >= h->nFaces) return fail(h, "tracked face out of range");
        if (faceIdx[faces[k]]) return fail(h, "face listed twice in the face tracker");
        faceIdx[faces[k]] = k + 1;
    }
    std::vector<int> slotTrack(h->slotFaceHost.size(), 0), bfTrack((size_t)std::max(h->nBFaces, 1), 0);
    for (int c = 0; c < h->nCells; ++c)
        for (int slot = h->slotOffHost[c]; slot < h->slotOffHost[c + 1]; ++slot) {
            const int f = h->slotFaceHost[slot];
            if (f >= 0 && f < h->nInternal && faceIdx[f]) slotTrack[slot] = (h->ownerHost[f] == c) ? faceIdx[f] : -faceIdx[f];
        }
    for (int p = 0; p < h->nPatches; ++p)
        for (int k = 0; k < h->patchSize[p]; ++k) {
            int f = h->patchStart[p] + k;
            const int b = f - h->nInternal;
            if (h->patchKind[p] == UGF_PATCH_CYCLIC) f = h->patchStart[h->patchPartnerHost[p]] + k;  // booked on the partner face
            bfTrack[b] = faceIdx[f];
        }
    const size_t nv = (size_t)n * h->nSpecies * UGF_NFT;
    if (dalloc(h, &h->dSlotTrack, slotTrack.size()) || dalloc(h, &h->dBfTrack, bfTrack.size()) || dalloc(h, &h->dFt, nv)) return 1;
    if (upload(h, h->dSlotTrack, slotTrack.data(), slotTrack.size()) || upload(h, h->dBfTrack, bfTrack.data(), bfTrack.size())) return 1;
    CU(cudaMemsetAsync(h->dFt, 0, sizeof(double) * nv, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->nTracked = n;
    return 0;
}

int ugf_download_face_tracker(ugf_handle* h, double* out, int32_t reset) {
    if (!h || !h->nTracked) return fail(h, "no face tracker set");
    CU(cudaSetDevice(h->cfg.device));
    const size_t nv = (size_t)h->nTracked * h->nSpecies * UGF_NFT;
    if (out) CU(countedMemcpyAsync(h, out, h->dFt, sizeof(double) * nv, cudaMemcpyDeviceToHost, h->stream));
    if (reset) CU(cudaMemsetAsync(h->dFt, 0, sizeof(double) * nv, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ugf_download_boundary_meas(ugf_handle* h, double* bm) {
    if (!h || !h->meshSet) return fail(h, "mesh not set");
    if (h->nBFaces > 0) {
        CU(countedMemcpyAsync(h, bm, h->dBm, sizeof(double) * (size_t)h->nBFaces * UGF_NBM, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int ugf_transfer_bytes(ugf_handle* h, int64_t* h2d, int64_t* d2h, int64_t* kernelArgs) {
    if (!h) return 1;
    if (h2d) *h2d = h->h2dBytes;
    if (d2h) *d2h = h->d2hBytes;
    if (kernelArgs) *kernelArgs = h->argBytes;
    return 0;
}

int ugf_host_alloc(ugf_handle* h, int64_t bytes, void** ptr) {
    if (!h || !ptr || bytes < 0) return fail(h, "ugf_host_alloc: bad argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 1)));
    h->hostOwned.push_back(*ptr);
    return 0;
}

int ugf_host_free(ugf_handle* h, void* ptr) {
    if (!h) return 1;
    auto it = std::find(h->hostOwned.begin(), h->hostOwned.end(), ptr);
    if (it == h->hostOwned.end()) return fail(h, "ugf_host_free: not a ugf_host_alloc pointer of this handle");
    h->hostOwned.erase(it);
    CU(cudaFreeHost(ptr));
    return 0;
}

int ugf_phase_times(ugf_handle* h, double* ms) {
    if (!h) return 1;
    for (int i = 0; i < UGF_NPHASE; ++i) ms[i] = 0;
    if (!h->timingValid) return 0;
    CU(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < UGF_NPHASE; ++i) {
        float t = 0;
        CU(cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
        ms[i] = t;
    }
    return 0;
}

int ugf_launch_count(ugf_handle* h, int64_t* n) {
    if (!h) return 1;
    *n = h->launches;
    return 0;
}

}  // extern "C"
