/* ugf.h — C ABI of libugf, the B200 (sm_100a) implementation of uniGasFoam's
 * per-timestep particle loop (uniGasCloud::evolve()).
 *
 * Every entry point is extern "C", takes plain pointers and sizes, returns an
 * int status (0 = ok, non-zero = error; ugf_last_error() gives the message) and
 * never throws.  One handle = one GPU = one decomposePar subdomain (rank).
 * Calls are stream-ordered on the handle's stream; entry points that hand data
 * back to the host synchronise that stream.
 *
 * Each declaration cites the reference interface it replaces; paths are relative
 * to the uniGasFoam repository root, with
 *   U/   = src/lagrangian/uniGas/
 *   CWM/ = src/lagrangian/CloudWithModels/
 *
 * The reference has no C ABI or FFI of its own: its plugin API is OpenFOAM's
 * run-time selection tables (declareRunTimeSelectionTable).  The enums below are
 * the TypeName strings of those tables, one enumerator per selectable model.
 * INTEGRATION.md shows the shim classes a uniGasFoam maintainer would register
 * under the same dictionary keys to forward into these calls.
 */
#ifndef UGF_H
#define UGF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGF_ABI_VERSION 4
#define UGF_MAX_SPECIES 8
#define UGF_MAX_VIB_MODES 4
#define UGF_MAX_ELEC_LEVELS 16

/* ---- run-time selection tables as enums ---------------------------------- */

/* constant/uniGasProperties: collisionModel  (U/clouds/uniGasCloud.C:236-248,656,708) */
enum { UGF_COLL_DSMC = 1, UGF_COLL_BGK = 2, UGF_COLL_HYBRID = 3 };

/* dsmcCollisionPartnerModel  (U/dsmcCollisionPartner/basic/dsmcCollisionPartner.H:78-89) */
enum { UGF_PARTNER_NTC = 1, UGF_PARTNER_NTC_SUBCYCLED = 2 };

/* dsmcCollisionModel  (U/dsmcCollisions/basic/dsmcCollisionModel/dsmcCollisionModel.H:76-86) */
enum {
    UGF_BINARY_NONE = 0,   /* noDSMCCollision */
    UGF_BINARY_VHS = 1,    /* variableHardSphere */
    UGF_BINARY_VSS = 2,    /* variableSoftSphere */
    UGF_BINARY_LB_VHS = 3, /* LarsenBorgnakkeVariableHardSphere */
    UGF_BINARY_LB_VSS = 4  /* LarsenBorgnakkeVariableSoftSphere */
};

/* bgkCollisionModel  (U/bgkCollisions/basic/bgkCollisionModel/bgkCollisionModel.H:82-93) */
enum {
    UGF_BGK_NONE = 0,     /* noBGKCollision */
    UGF_BGK_BGK = 1,      /* stochasticParticleBGK */
    UGF_BGK_ESBGK = 2,    /* stochasticParticleESBGK */
    UGF_BGK_SBGK = 3,     /* stochasticParticleSBGK */
    UGF_BGK_USP_SBGK = 4  /* unifiedStochasticParticleSBGK */
};

/* polyPatch types as seen by particle::trackToAndHitFace (OpenFOAM; call site
 * U/parcels/uniGasParcel.C:74) */
enum {
    UGF_PATCH_WALL = 1,      /* -> hitWallPatch -> patch boundary model (U/parcels/uniGasParcel.C:131-145) */
    UGF_PATCH_SYMMETRY = 2,  /* symmetry / symmetryPlane: reflect (U/parcels/uniGasParcel.C:148-155) */
    UGF_PATCH_CYCLIC = 3,    /* jump to partner patch face (same local index) + separation */
    UGF_PATCH_EMPTY = 4,     /* never hit: tracking velocity is constrained (U/parcels/uniGasParcel.C:64-71) */
    UGF_PATCH_PROCESSOR = 5, /* hitProcessorPatch: migrate (U/parcels/uniGasParcel.C:121-128) */
    UGF_PATCH_GENERIC = 6    /* "type patch": OpenFOAM default, td.keepParticle = false */
};

/* uniGasPatchBoundary models  (U/boundaries/basic/uniGasPatchBoundary/uniGasPatchBoundary.H:90-101) */
enum {
    UGF_WALL_UNSET = 0,
    UGF_WALL_DIFFUSE = 1,  /* uniGasDiffuseWallPatch   params: T, Ux, Uy, Uz */
    UGF_WALL_SPECULAR = 2, /* uniGasSpecularWallPatch  params: none */
    UGF_WALL_MIXED = 3,    /* uniGasMixedDiffuseSpecularWallPatch params: T, Ux, Uy, Uz, diffuseFraction */
    UGF_WALL_DELETION = 4, /* uniGasDeletionPatch      params: none */
    UGF_WALL_CLL = 5       /* uniGasCLLWallPatch (Cercignani-Lampis-Lord) params: T, Ux, Uy, Uz, normalAccommCoeff,
                              tangentialAccommCoeff, rotEnergyAccommCoeff
                              (U/boundaries/derived/patchBoundaries/uniGasCLLWallPatch/uniGasCLLWallPatch.C:80-254) */
};

/* ---- plain-old-data inputs ------------------------------------------------ */

/* constant/uniGasProperties + system/controlDict subset. */
typedef struct ugf_config {
    int32_t abiVersion;        /* must be UGF_ABI_VERSION */
    int32_t device;            /* CUDA device ordinal */
    uint64_t seed;             /* counter-based Philox key; reference seeds from the wall clock
                                  (U/clouds/uniGasCloud.C:490) */
    double nParticle;          /* nEquivalentParticles, F_N  (U/clouds/uniGasCloud.C:414) */
    double deltaT;             /* controlDict deltaT */
    int32_t solutionD[3];      /* 1 = solved direction, 0 = empty  (U/clouds/uniGasCloud.C:513-532) */
    int32_t collisionModel;    /* UGF_COLL_* */
    int32_t partnerModel;      /* UGF_PARTNER_* */
    int32_t binaryModel;       /* UGF_BINARY_* */
    int32_t bgkModel;          /* UGF_BGK_* */
    int32_t nSubCycles;        /* noTimeCounterSubCycled only */
    int32_t macroInterpolation;/* collisionProperties.macroInterpolation: 1 = the BGK target fields are interpolated to the parcel's
                                  position (interpolationCellPoint); needs ugf_set_macro_interpolation */
    double Tref;               /* collisionProperties.Tref */
    double theta;              /* collisionProperties.theta (default 1) */
    double rotationalRelaxationCollisionNumber; /* LB: Z_rot  (…LBVHS.C:60-63) */
    double electronicRelaxationCollisionNumber; /* LB: Z_elec (…LBVHS.C:64-67) */
    int64_t parcelCapacity;    /* SoA capacity in parcels (resident state is 2 buffers of this size) */
    int32_t sampleInterval;    /* fieldPropertiesDict timeProperties.sampleInterval (default 1) */
    int32_t measureWalls;      /* 1 = accumulate boundaryMeasurements on wall faces */
    int32_t rank;              /* subdomain index (Pstream::myProcNo) */
    int32_t nRanks;
    int32_t axisymmetric;      /* axisymmetricSimulation (U/clouds/uniGasCloud.C:420, 563-568): radial weighting factor
                                  RWF(x) = 1 + (maxRWF - 1) sqrt(y^2 + z^2) / radialExtent (uniGasCloudI.H:116-120) on a wedge
                                  mesh about the x axis whose side patches are symmetryPlane (uniGasBoundaries.C:438-445) */
    double radialExtent;       /* axisymmetricProperties.radialExtentOfDomain */
    double maxRWF;             /* axisymmetricProperties.maxRadialWeightingFactor */
} ugf_config;

/* moleculeProperties.<species>  (U/parcels/uniGasParcel.H:68-120, uniGasParcelI.H:40-141) */
typedef struct ugf_species {
    double mass;
    double d;              /* diameter */
    double omega;
    double alpha;
    int32_t rotationalDoF;
    int32_t vibrationalDoF;            /* number of vibrational modes */
    double thetaV[UGF_MAX_VIB_MODES];
    double thetaD[UGF_MAX_VIB_MODES];
    double Zref[UGF_MAX_VIB_MODES];
    double TrefZv[UGF_MAX_VIB_MODES];
    int32_t charge;
    int32_t nElectronicLevels;
    double electronicEnergy[UGF_MAX_ELEC_LEVELS];
    int32_t degeneracy[UGF_MAX_ELEC_LEVELS];
} ugf_species;

/* polyMesh flattened (constant/polyMesh/{points,faces,owner,neighbour,boundary} plus
 * the derived geometry OpenFOAM computes: faceAreas, faceCentres, cellVolumes,
 * cellCentres, cellPoints bounding box).  All vectors are AoS xyz, fp64; all
 * labels int32 (WM_LABEL_SIZE=32).  Boundary faces of patch p are the contiguous
 * range [patchStart[p], patchStart[p]+patchSize[p]).  Face area vectors point
 * from owner to neighbour (out of the domain on boundary faces).
 * points/facePointOffsets/facePoints are only needed for inflow patches
 * (triangle fan for insertion, U/boundaries/basic/uniGasGeneralBoundary/uniGasGeneralBoundary.C:558-578)
 * and may be NULL otherwise. */
typedef struct ugf_mesh {
    int32_t nCells, nFaces, nInternalFaces, nPatches, nPoints;
    const int32_t* owner;           /* [nFaces] */
    const int32_t* neighbour;       /* [nInternalFaces] */
    const double* faceAreas;        /* [nFaces*3]  Sf */
    const double* faceCentres;      /* [nFaces*3]  Cf */
    const int32_t* cellFaceOffsets; /* [nCells+1]  CSR cell -> faces */
    const int32_t* cellFaces;       /* [cellFaceOffsets[nCells]] */
    const double* cellVolumes;      /* [nCells] */
    const double* cellCentres;      /* [nCells*3] */
    const double* cellBbMin;        /* [nCells*3] min over cellPoints (noTimeCounter.C:112-127) */
    const double* cellBbMax;        /* [nCells*3] */
    const int32_t* patchStart;      /* [nPatches] */
    const int32_t* patchSize;       /* [nPatches] */
    const int32_t* patchKind;       /* [nPatches] UGF_PATCH_* */
    const int32_t* patchPartner;    /* [nPatches] cyclic: partner patch; processor: peer rank; else -1 */
    const double* patchSeparation;  /* [nPatches*3] added to the position when crossing (cyclic, processorCyclic) */
    const double* points;           /* [nPoints*3] or NULL */
    const int32_t* facePointOffsets;/* [nFaces+1] or NULL */
    const int32_t* facePoints;      /* CSR or NULL */
} ugf_mesh;

/* boundariesDict: uniGasFreeStreamInflowPatchProperties
 * (U/boundaries/derived/generalBoundaries/uniGasFreeStreamInflowPatch/uniGasFreeStreamInflowPatch.C:61-96) */
typedef struct ugf_inflow {
    int32_t nTypeIds;
    int32_t typeIds[UGF_MAX_SPECIES];
    double numberDensities[UGF_MAX_SPECIES];
    double translationalTemperature;
    double rotationalTemperature;
    double vibrationalTemperature;
    double electronicTemperature;
    double velocity[3];
} ugf_inflow;

/* boundariesDict: uniGasLiouFangPressureInletPatchProperties
 * (U/boundaries/derived/generalBoundaries/uniGasLiouFangPressureInletPatch/uniGasLiouFangPressureInletPatch.C:54-103):
 * subsonic pressure inlet - number density p / (k T), inflow velocity per face relaxed towards the mean velocity of the
 * parcels in the face's cell after the collisions of every step (:126-174). */
typedef struct ugf_pressure_inlet {
    int32_t nTypeIds;
    int32_t typeIds[UGF_MAX_SPECIES];
    double moleFractions[UGF_MAX_SPECIES];
    double inletPressure;
    double inletTemperature;
    double theta;              /* default 1 */
} ugf_pressure_inlet;

/* Parcels as host SoA.  Mandatory: x,y,z,Ux,Uy,Uz,cell.  Optional (NULL = default):
 * typeId (0), ERot (0), stepFraction/newParcel (0).  (U/parcels/uniGasParcel.H:217-239)
 * cellWeight (the parcel's CWF, lagrangian/uniGas/cellWeight on disk) is implicit on the device: a parcel carries the
 * cellWeightFactor of the cell it was in at the last weighting pass, which is what the reference guarantees after
 * uniGasCloud::weighting() (U/clouds/uniGasCloud.C:203-220, 1353-1424).  On upload it is optional and must equal
 * cellWeightFactor[cell] (checked); on download it is filled in when a pointer is given.
 * radialWeight (the parcel's RWF, lagrangian/uniGas/radialWeight) is implicit on the device in the same way: after
 * axisymmetricWeighting() (U/clouds/uniGasCloud.C:1427-1570) every parcel carries RWF(its position); the initialisation
 * models and the inflow patches hand out RWF(centre of the parcel's cell) instead (uniGasMeshFill.C:260,
 * uniGasGeneralBoundary.C:739).  On upload it is optional: NULL = RWF(cell centre) as the initialisation models set it;
 * given, all values must be RWF(position) or all RWF(cell centre) (checked to 1e-6, then recomputed); the
 * first move weights against them.  On download it is filled in when a pointer is given. */
typedef struct ugf_parcels {
    int64_t n;
    double* x; double* y; double* z;
    double* Ux; double* Uy; double* Uz;
    int32_t* cell;
    int32_t* typeId;
    double* ERot;
    int32_t* newParcel;
    double* cellWeight;
    int32_t* vibLevel;   /* [n][UGF_MAX_VIB_MODES] vibrational quantum level per mode (uniGasParcel.H:238); NULL = 0 */
    int32_t* ELevel;     /* [n] electronic level (uniGasParcel.H:232); NULL = 0 */
    double* radialWeight;/* [n] RWF (uniGasParcel.H:229); axisymmetric simulations only, see above */
} ugf_parcels;

/* Per-step log quantities (noTimeCounter.C:318-342, …USP.C:976-992, uniGasCloud.C:878-920). */
typedef struct ugf_counters {
    int64_t step;                /* steps taken so far */
    int64_t nParcels;            /* live parcels after the last phase */
    int64_t collisionCandidates; /* NTC candidates, last step */
    int64_t collisions;          /* accepted DSMC collisions, last step */
    int64_t bgkRelaxations;      /* parcels relaxed by the BGK model, last step */
    int64_t inserted;            /* parcels inserted by inflow patches, last step */
    int64_t deleted;             /* parcels deleted at patches, last step */
    int64_t migrated;            /* parcels sent to other ranks, last step */
    int64_t wallHits;            /* wall-patch interactions, last step */
    int64_t stuck;               /* parcels dropped by the tracking iteration guard (must stay 0) */
    double linearKineticEnergy;  /* sum 0.5 m |U|^2 over parcels (x F_N = info()) */
    double rotationalEnergy;     /* sum ERot */
    double momentum[3];          /* sum m U */
    int64_t cloned;              /* parcels added by cellWeighting(), last step (U/clouds/uniGasCloud.C:1366-1407) */
    int64_t weightDeleted;       /* parcels removed by cellWeighting(), last step (:1409-1420) */
    double vibrationalEnergy;    /* sum over parcels and modes of level * k * thetaV (info(): U/clouds/uniGasCloud.C:878-920) */
    double electronicEnergy;     /* sum of electronicEnergyList[ELevel] */
} ugf_counters;

/* system/hybridDecompositionDict: uniGasHybridDecomposition (U/hybridDecomposition/basic/uniGasHybridDecomposition.C:48-76)
 * + localKnudsen (U/hybridDecomposition/derived/localKnudsen/localKnudsen.C:52-55). */
typedef struct ugf_decomposition {
    int32_t decompositionInterval;         /* timeProperties.decompositionInterval (default 100) */
    int32_t resetAtDecomposition;          /* timeProperties.resetAtDecomposition (default 1) */
    double resetAtDecompositionUntilTime;  /* timeProperties.resetAtDecompositionUntilTime (default 1e300) */
    double breakdownMax;                   /* localKnudsenProperties.breakdownMax (default 0.05) */
    double theta;                          /* localKnudsenProperties.theta (default 1) */
    int32_t smoothingPasses;               /* localKnudsenProperties.smoothingPasses (default 0) */
    int32_t refinementPasses;              /* 3, hard-wired in the reference (uniGasHybridDecomposition.C:71) */
    int32_t neighborLevels;                /* 3 (:72) */
    double maxNeighborFraction;            /* 0.4 (:73) */
} ugf_decomposition;

/* interpolationCellPoint for collisionProperties.macroInterpolation true (U/bgkCollisions/derived/unifiedStochasticParticleSBGK/
 * unifiedStochasticParticleSBGK.C:893-947 and the same block in the other three BGK models): the geometry OpenFOAM derives from the
 * polyMesh, handed over once.  Point value = sum of pointWeights x cell value over pointCells (inverse-distance weights of
 * volPointInterpolation, boundary points fed by their boundary faces' owner cells, cyclic partners merged; normalised per point);
 * at points with a non-zero pointNormals entry (symmetry patches) vectors / tensors lose their normal components.  A cell is split
 * into the tets (cell centre, tetPoints[t][0..2]); the value at a position is linear in the tet that contains it (cellPointWeight). */
typedef struct ugf_cell_point {
    int32_t nPoints;
    const double* points;             /* [nPoints*3] */
    const int32_t* tetOffsets;        /* [nCells+1] */
    const int32_t* tetPoints;         /* [nTets*3] */
    const int32_t* pointCellOffsets;  /* [nPoints+1] */
    const int32_t* pointCells;
    const double* pointWeights;
    const double* pointNormals;       /* [nPoints*3] */
} ugf_cell_point;

/* Number of fp64 values per (cell, species) in the cell-moment block; see DESIGN.md
 * for the slot list.  (U/cellMeasurements/cellMeasurements.H:73-145) */
#define UGF_NMOM 32
/* Internal-mode accumulators per (cell, species), kept only when a species has vibrational modes or more than one
 * electronic level (uniGasVolFields.C:775-793): time-weighted sums of 0 nParcels, 1 electronicETotal, 2 nGroundElectronicLevel,
 * 3 nFirstElectronicLevel, 4..7 vibrationalETotal per mode. */
#define UGF_NINT 8

/* Number of fp64 values per wall face in the boundary-measurement block
 * (U/boundaryMeasurements/boundaryMeasurements.C:70-121). */
#define UGF_NBM 16

/* Number of fp64 values per (tracked face, species) of the face tracker: 0 parcels, 1 mass, 2-4 momentum, 5 energy, each
 * weighted with the parcel's cell weight factor (U/faceTracker/uniGasFaceTracker.C:90-152). */
#define UGF_NFT 6

/* Number of fp64 output fields per cell from ugf_download_fields
 * (U/macroscopicProperties/derived/volumetric/uniGasVolFields/uniGasVolFields.C:839-1254):
 * 0 uniGasRhoNMean, 1 rhoN, 2 rhoM, 3-5 UMean, 6 translationalT, 7 rotationalT,
 * 8 overallT, 9 p, 10 Ma, 11 densityError (0 if undefined); measureMeanFreePath (:1124-1232, Bird eqs 4.76, 4.77,
 * 4.74, 1.38; Tref = collisionProperties.Tref): 12 MFP, 13 dxMFP (largest sub-cell dimension / MFP), 14 MCR, 15 MCT,
 * 16 dtMCT (deltaT / MCT); measureErrors (:1234-1254): 17 velocityError, 18 temperatureError; 19 vibrationalT (:930-1010),
 * 20 electronicT (:1012-1062).  overallT (8) weights all four modes (:1071-1079). */
#define UGF_NFIELD 21
/* per wall face: 0 rhoN, 1 rhoM, 2-4 UMean, 5 translationalT, 6 q (surfaceHeatTransfer),
 * 7-9 fD, 10 p, 11 tau. */
#define UGF_NWALLFIELD 12

typedef struct ugf_handle ugf_handle;

/* ---- lifetime -------------------------------------------------------------- */

/* uniGasCloud constructor (U/clouds/uniGasCloud.C:400-780). */
int ugf_create(const ugf_config* cfg, ugf_handle** out);
int ugf_destroy(ugf_handle* h);
/* Message for the last non-zero status on this handle (h may be NULL for create errors).
 * Replaces FatalErrorInFunction ... exit(FatalError). */
const char* ugf_last_error(const ugf_handle* h);
int ugf_abi_version(void);

/* ---- case set-up (construction-time in the reference) ---------------------- */

/* buildConstProps (U/clouds/uniGasCloud.C:42-62). */
int ugf_set_species(ugf_handle* h, int32_t n, const ugf_species* sp);
/* polyMesh reference held by the cloud; flattened to CSR + face planes on the device. */
int ugf_set_mesh(ugf_handle* h, const ugf_mesh* mesh);
/* uniGasBoundaries: patchToModelId_ (U/boundaries/basic/uniGasBoundaries/uniGasBoundaries.C:420-490).
 * params: see UGF_WALL_*.  Only wall-kind patches take a model. */
int ugf_set_patch_model(ugf_handle* h, int32_t patch, int32_t wallModel, const double* params, int32_t nParams);
/* The *FieldPatch variants (uniGasDiffuseWallFieldPatch, uniGasMixedDiffuseSpecularWallFieldPatch,
 * uniGasCLLWallFieldPatch: …/uniGasDiffuseWallFieldPatch/uniGasDiffuseWallFieldPatch.C:104-121): wall temperature and
 * velocity per boundary face instead of per patch, i.e. the boundaryT / boundaryU fields on this patch.
 * T [patchSize], U [patchSize*3] (host); call after ugf_set_patch_model on the same patch. */
int ugf_set_patch_wall_fields(ugf_handle* h, int32_t patch, const double* T, const double* U);
/* uniGasFreeStreamInflowPatch on a patch (any kind). */
int ugf_set_inflow(ugf_handle* h, int32_t patch, const ugf_inflow* inflow);
/* uniGasFreeStreamInflowFieldPatch on a patch (U/boundaries/derived/generalBoundaries/uniGasFreeStreamInflowFieldPatch/
 * uniGasFreeStreamInflowFieldPatch.C:50-228): the free-stream insertion with number density, temperatures and velocity
 * per face of the patch, i.e. the values of the volFields boundaryNumberDensity_<species>, boundaryTransT, boundaryRotT
 * and boundaryU on this patch (computeParcelsToInsert / insertParcels with per-face lists, uniGasGeneralBoundary.C:
 * 115-169, 537-761).  numberDensity [nTypeIds][patchSize], transT / rotT [patchSize] (rotT may be NULL), U [patchSize*3]
 * (host). */
int ugf_set_inflow_fields(ugf_handle* h, int32_t patch, int32_t nTypeIds, const int32_t* typeIds, const double* numberDensity,
                          const double* transT, const double* rotT, const double* U);
/* uniGasChapmanEnskogFreeStreamInflowPatch on a patch (U/boundaries/derived/generalBoundaries/uniGasChapmanEnskogFreeStreamInflowPatch/
 * uniGasChapmanEnskogFreeStreamInflowPatch.C:52-127): free-stream insertion from a Chapman-Enskog distribution with the given heat flux
 * vector [3] and stress tensor [9, row-major] - the count of Bird 4.22 corrected by the normal stress and heat flux
 * (uniGasGeneralBoundary.C:171-239), velocities by Garcia & Alder's acceptance-rejection (:763-1001). */
int ugf_set_chapman_enskog_inflow(ugf_handle* h, int32_t patch, const ugf_inflow* inflow, const double* heatFlux, const double* stress);
/* uniGasLiouFangPressureInletPatch on a patch.  The insertion itself is the free-stream one with a velocity per face
 * (uniGasGeneralBoundary.C:369-425, 1003-1230); the count formula is evaluated with speed ratios up to 5. */
int ugf_set_pressure_inlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* inlet);
/* uniGasWangPressureInletPatch on a patch (U/boundaries/derived/generalBoundaries/uniGasWangPressureInletPatch/
 * uniGasWangPressureInletPatch.C:54-281): number density p / (k T) as above; the inflow velocity of a face is the mean
 * momentum over the mean mass of all parcels seen in its cell since the start and, after 100 steps, is corrected by
 * (p_cell - p_in) / (rho a) along the outward normal (:258-266).  `theta` of the struct is not used.  The running sums
 * and the step count travel in ugf_state_save. */
int ugf_set_wang_pressure_inlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* inlet);
/* uniGasLiouFangPressureOutletPatch on a patch (U/boundaries/derived/generalBoundaries/uniGasLiouFangPressureOutletPatch/
 * uniGasLiouFangPressureOutletPatch.C:50-322).  The struct is read as: inletPressure = outletPressure, inletTemperature =
 * the outlet temperature before the first evaluation (300 K in the reference, :74), theta unused.  After the collisions of
 * every step the running sums of each outlet face's cell give density, temperature and pressure there and, with the
 * characteristic relations of Liou & Fang (2000, eq 26), the number density, temperature and velocity the face inserts
 * with at the next step.  The device caps a slot's insertions at the count of a gas at twice p_e / (k T_0), T_0 and a speed
 * ratio of 5 (the bound the array capacity is checked against); reaching it is an error, not a silent clamp. */
int ugf_set_pressure_outlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* outlet);
/* uniGasMassFlowRateInletPatch on a patch of type patch (U/boundaries/derived/generalBoundaries/uniGasMassFlowRateInletPatch/
 * uniGasMassFlowRateInletPatch.C:53-302).  The struct is read as: moleFractions, inletTemperature, theta; inletPressure unused.
 * After the collisions of every step the inlet velocity of each face relaxes (theta) towards the mean velocity of its cell (kept
 * when it would point out of the domain), the number density of each species follows the parcels in the cell, and all are scaled
 * by parcelsIn / parcelsToInsert so that the next step's count adds up to massFlowRate dt / (mean molecular mass F_N) plus the
 * parcels that left through the patch this step (the face tracker's parcelIdFlux on the patch faces, kept by the move).  As in
 * the reference the parcels themselves are inserted from a gas at rest at inletTemperature (:139-149) and the patch-wide
 * parcelsToInsert sums over all species (scalarField += scalar, :262-271).  typeIds must be 0..nSpecies-1 (the reference indexes
 * the tracker and the densities with the patch-local index); a patch whose cells are all empty is an error (0 / 0 there);
 * decomposed meshes are refused (patch-wide sums).  initialVelocity [3] or NULL (zero). */
int ugf_set_mass_flow_inlet(ugf_handle* h, int32_t patch, const ugf_pressure_inlet* inlet, double massFlowRate, const double* initialVelocity);
/* Inlet velocity per face of a pressure inlet [patchSize*3] (diagnostic / restart). */
int ugf_download_inlet_velocity(ugf_handle* h, int32_t patch, double* U);
/* addNewParcel over a whole initial configuration (U/clouds/uniGasCloud.C:260-290). Replaces the cloud. */
int ugf_upload_parcels(ugf_handle* h, const ugf_parcels* p);
/* Cell state carried between steps (U/clouds/uniGasCloud.H:189-201): sigmaTcRMax [nCells],
 * collModelId [nCells] (0 = bgk, 1 = dsmc), subCellLevels [nCells*3], cellWeightFactor [nCells] (> 0; all 1 until
 * uploaded).  NULL keeps the current value.
 * Uploading cellWeightFactor switches cell weighting on (cellWeightedSimulation, U/clouds/uniGasCloud.C:417): every
 * step the move is followed by cellWeighting() (:1353-1424) - a parcel that arrives in a cell with a smaller factor
 * is cloned, one that arrives in a cell with a larger factor survives with probability old/new - and the factor
 * scales F_N in the NTC candidate count (noTimeCounter.C:168-184), the weighted cell sums (cellMeasurements.C:463-467),
 * the BGK macroscopic state, the inflow count (uniGasGeneralBoundary.C:154-165), the wall heat flux / force
 * (uniGasPatchBoundary.C:292-299) and the wall fields (uniGasVolFields.C:1276-1278).  Upload it before the parcels.
 * A factor changed while parcels exist (uniGasDynamicAdapter.C:660-677) is applied by the next step's weighting pass
 * (old factor of the parcel's previous cell / new factor of its current cell).  Across processor patches the factor
 * travels with the parcel (migration record slot 9). */
int ugf_upload_cell_state(ugf_handle* h, const double* sigmaTcRMax, const int32_t* collModelId,
                          const int32_t* subCellLevels, const double* cellWeightFactor);
int ugf_set_deltaT(ugf_handle* h, double deltaT);
/* Restart: continue the step count (and with it the counter-based random streams) from Time::timeIndex as stored in
 * <time>/uniform/time (U/clouds/uniGasCloud.C:581-593 reads deltaT from the same dictionary). */
int ugf_set_time_index(ugf_handle* h, int64_t index);
/* Checkpoint of everything the loop carries between steps besides the parcels and the cell-state fields of
 * ugf_upload_cell_state: step count, uniGasVolFields accumulators and counters (the reference stores them in
 * <time>/uniform/volFieldsMethod_*, uniGasVolFields.C:550-669), BGK persistent state (maxProb, previous heat flux / shear
 * stress), sigmaTcRMax, collision-model mask, local-Knudsen accumulators and blended Knudsen fields, pressure-inlet face
 * velocities.  A flat array of doubles with a self-describing header; written by one implementation it can be read by
 * another on the same mesh / species / model set-up (checked).  ugf_state_size gives the length in doubles. */
int ugf_state_size(ugf_handle* h, int64_t* nDoubles);
int ugf_state_save(ugf_handle* h, double* buf, int64_t nDoubles);
int ugf_state_load(ugf_handle* h, const double* buf, int64_t nDoubles);

/* ---- the hot path ----------------------------------------------------------- */

/* uniGasCloud::evolve() nSteps times (U/clouds/uniGasCloud.C:821-869):
 * inflow -> move(+patches) -> occupancy -> sample -> collide/relax -> field accumulation.
 * Single-rank only; multi-rank callers drive the phases below with the migrate calls. */
int ugf_step(ugf_handle* h, int32_t nSteps);

/* Phase-wise entry points, for parity tests, profiling and multi-rank drivers. */
/* boundaries_.controlBeforeMove(): free-stream insertion (uniGasGeneralBoundary.C:115-169,537-761). */
int ugf_control_before_move(ugf_handle* h);
/* Cloud<uniGasParcel>::move (U/clouds/uniGasCloud.C:836, U/parcels/uniGasParcel.C:35-108). Tracks every
 * parcel to the end of the step or to a processor face; also counts parcels per destination cell. */
int ugf_move(ugf_handle* h);
/* buildCellOccupancy (CWM/CloudWithModels/CloudWithModels.C:110-138): CSR offsets + parcel ids, stable. */
int ugf_sort(ugf_handle* h);
/* Physically reorder the SoA into cell-major order (no reference analogue: the reference keeps pointers). */
int ugf_reorder(ugf_handle* h);
/* cellMeasurements::calculateFields (U/cellMeasurements/cellMeasurements.C:408-513). */
int ugf_sample(ugf_handle* h);
/* dsmcCollisionPartner::collide (…/noTimeCounter/noTimeCounter.C:66-343) with the selected binary model. */
int ugf_collide(ugf_handle* h);
/* bgkCollisionModel::collide (e.g. …/unifiedStochasticParticleSBGK.C:871-994). */
int ugf_relax(ugf_handle* h);
/* uniGasVolFields::calculateField accumulation part (uniGasVolFields.C:723-837) +
 * cellMeas_/boundaryMeas_ clean (U/clouds/uniGasCloud.C:864-866). */
int ugf_accumulate_fields(ugf_handle* h);
/* decompositionModel localKnudsen (collisionModel hybrid only).  Once set, every step adds the step's cell sums to the
 * decomposition's own time averages and every decompositionInterval-th step recomputes the DSMC / BGK mask
 * (collModelId) from the gradient-length local Knudsen number: localKnudsen::decompose
 * (U/hybridDecomposition/derived/localKnudsen/localKnudsen.C:216-582).  ugf_step / ugf_finish_step call it after the
 * collisions like uniGasCloud::evolve does (U/clouds/uniGasCloud.C:862); phase-wise drivers call ugf_decompose after
 * ugf_collide / ugf_relax.  Processor faces are treated as zero-gradient by the smoothing operator. */
int ugf_set_decomposition(ugf_handle* h, const ugf_decomposition* d);
/* Geometry for macroInterpolation (see ugf_cell_point); call after ugf_set_mesh when ugf_config.macroInterpolation is 1. */
int ugf_set_macro_interpolation(ugf_handle* h, const ugf_cell_point* cp);
int ugf_decompose(ugf_handle* h);
/* collModelId [nCells] (0 = bgk, 1 = dsmc) and the time-blended Knudsen fields kn [nCells][4] = KnRho, KnT, KnU,
 * KnGLL (either may be NULL). */
int ugf_download_decomposition(ugf_handle* h, int32_t* collModelId, double* kn);
/* End-of-step bookkeeping when phases are driven one by one: step counter ++. */
int ugf_end_step(ugf_handle* h);
/* Everything of evolve() that follows the move and the parcel transfers, in one call and with the same kernel
 * fusion as ugf_step: occupancy -> reorder + sample (+ field accumulation) -> collide -> relax -> wall fields ->
 * end of step (U/clouds/uniGasCloud.C:839-866).  Multi-rank drivers call it once the transfer loop has settled. */
int ugf_finish_step(ugf_handle* h);

/* ---- multi-rank parcel migration (Cloud::move transfer loop, OpenFOAM; §2.1 of SURVEY) ---- */

/* After ugf_move: number of parcels waiting on each processor patch [nPatches] (0 for other kinds). */
int ugf_migrate_counts(ugf_handle* h, int64_t* sendCounts);
/* Pack the parcels waiting on `patch` into a device buffer of UGF_MIGRATE_STRIDE doubles per parcel
 * (x,y,z,Ux,Uy,Uz,ERot,stepFraction, localFace + 2^32 typeId, cellWeight; all doubles) and remove them from the cloud.
 * *devBuf is owned by the handle and valid until the next pack on the same patch. */
#define UGF_MIGRATE_STRIDE 10
int ugf_migrate_pack(ugf_handle* h, int32_t patch, double** devBuf, int64_t* nPacked);
/* Append n received parcels (same record layout, device pointer) arriving through `patch`; they are
 * placed in the owner cell of the matching local face.  Follow with ugf_move_received. */
int ugf_migrate_unpack(ugf_handle* h, int32_t patch, const double* devBuf, int64_t n);
/* Continue tracking the parcels appended since the last ugf_move / ugf_move_received. */
int ugf_move_received(ugf_handle* h);
/* Fixed-slot variant of pack/unpack for drivers that must not synchronise with the host: every processor patch
 * (in ascending patch order, k = 0,1,...) owns slot k of a caller-provided device buffer.  A slot is
 * (slotCapacity + 1) records of UGF_MIGRATE_STRIDE doubles: record 0 is a header whose first double is the
 * number of parcels that follow.  ugf_migrate_pack_slots fills the slots from the parcels waiting on the
 * processor patches (index order, stable) and removes them from the cloud; ugf_migrate_unpack_slots appends
 * the parcels found in received slots (slot k = what the peer packed for the patch matching patch k).  Neither
 * call blocks; a slot overflow raises the handle's device error flag (reported by ugf_counters_get). */
int ugf_migrate_pack_slots(ugf_handle* h, double* devSend, int64_t slotCapacity);
int ugf_migrate_unpack_slots(ugf_handle* h, const double* devRecv, int64_t slotCapacity);
/* NVLink peer-memory variant of the slot transfer: no send buffer and no NCCL call on the data path.  Every rank
 * allocates its receive region with ugf_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; the 64-byte handle is passed
 * to the neighbours by the caller), maps the neighbours' regions with ugf_peer_open, and then per transfer round
 *   ugf_migrate_pack_peer   packs every processor patch k straight into dstSlots[k] - the matching receive slot in
 *                           the neighbour's memory - and then stores `epoch` to dstFlags[k] (system-scope release);
 *   ugf_migrate_unpack_peer waits on the device until its own flags [nProcPatches] have reached `epoch`, then unpacks
 *                           devRecv like ugf_migrate_unpack_slots.
 * Neither call blocks the host.  Slot layout as above; epochs must increase from round to round and callers must
 * alternate between two receive buffers from one round to the next (a slot is rewritten only after the round in
 * between has been acknowledged by both sides).  A wait that sees no signal for several seconds raises the handle's
 * device error flag instead of hanging. */
int ugf_peer_alloc(ugf_handle* h, int64_t bytes, void** devPtr, unsigned char* ipcHandle64);
int ugf_peer_open(ugf_handle* h, const unsigned char* ipcHandle64, void** devPtr);
int ugf_migrate_pack_peer(ugf_handle* h, double* const* dstSlots, uint64_t* const* dstFlags, int64_t slotCapacity, uint64_t epoch);
int ugf_migrate_unpack_peer(ugf_handle* h, const double* devRecv, const uint64_t* devFlags, int64_t slotCapacity, uint64_t epoch);
/* Device address of an int64 holding the number of parcels currently waiting on processor patches (valid after
 * ugf_move / ugf_move_received): all-reduce it to decide whether another transfer round is needed. */
int ugf_migrate_inflight(ugf_handle* h, int64_t** devCounter);
/* Raw stream (cudaStream_t) the handle launches on, for interop with NCCL / torch. */
int ugf_stream(ugf_handle* h, void** stream);

/* ---- results ----------------------------------------------------------------- */

int ugf_counters_get(ugf_handle* h, ugf_counters* out);
int ugf_num_parcels(ugf_handle* h, int64_t* n);
/* Download parcels in current device order; any pointer in p may be NULL. p->n in: capacity, out: count. */
int ugf_download_parcels(ugf_handle* h, ugf_parcels* p);
/* cellOccupancy() (CWM/CloudWithModels/CloudWithModelsI.H:65-75) as CSR: offsets [nCells+1], ids [n]. */
int ugf_download_cell_occupancy(ugf_handle* h, int32_t* offsets, int32_t* ids);
/* Moments of the last ugf_sample / ugf_collide / ugf_relax: [nCells][nSpecies][UGF_NMOM].  ugf_step and
 * ugf_finish_step keep them only when a BGK model is active (it reads them); a pure-DSMC step folds them into the
 * field accumulators without storing them - call ugf_sample to get them. */
int ugf_download_cell_moments(ugf_handle* h, double* moments);
/* sigmaTcRMax [nCells]; BGK persistent state maxProb [nCells], qPrev [nCells*3], sPrev [nCells*6] (NULL skips). */
int ugf_download_cell_state(ugf_handle* h, double* sigmaTcRMax, double* maxProb, double* qPrev, double* sPrev);
/* Time-averaged fields, derived as at write time: cells [nCells][UGF_NFIELD]; wall faces
 * [nBoundaryFaces][UGF_NWALLFIELD] (zero on non-wall faces). resetAtOutput clears the accumulators. */
int ugf_download_fields(ugf_handle* h, double* cellFields, double* wallFields, int32_t resetAtOutput);
/* Raw per-step boundary measurements of the last move: [nBoundaryFaces][UGF_NBM]. */
int ugf_download_boundary_meas(ugf_handle* h, double* bm);
/* uniGasFaceTracker (U/faceTracker/uniGasFaceTracker.C:90-152; called from uniGasParcel::move at every face hit,
 * U/parcels/uniGasParcel.C:85): number, mass, momentum and energy carried through faces, per species.  The reference
 * tallies every face of the mesh every step and its consumers (uniGasMassFluxSurface) read the faces of their face zone;
 * here only the registered faces are tallied (global face labels, internal or boundary; NULL / 0 switches it off) and
 * the tallies run on until downloaded with reset != 0 - the consumer's sums over sampling steps.  Internal faces: signed
 * with the direction of travel relative to the face area vector (owner -> neighbour); boundary faces: booked after the
 * patch interaction with the sign of the velocity the parcel then has; cyclic faces: booked on the partner face. */
int ugf_set_face_tracker(ugf_handle* h, int32_t nFaces, const int32_t* faces);
/* out [nFaces][nSpecies][UGF_NFT] in the order of the registered list. */
int ugf_download_face_tracker(ugf_handle* h, double* out, int32_t reset);
/* The raw time-weighted sums behind ugf_download_fields: acc [nCells][16] (slot list in DESIGN.md: 0 sum dt N, 8 .. 13
 * the XnParticle sums of number, mass, momentum, kinetic energy), accSpecies [nCells][nSpecies] (nParcelsXnParticle per
 * species) and the averaging time / step count.  uniGasDynamicAdapter::adapt accumulates exactly these sums over its
 * own interval (U/dynamicAdaptation/uniGasDynamicAdapter.C:510-527); the host-side adapter differences two downloads
 * instead of keeping a second set of accumulators on the device.  Any pointer may be NULL. */
int ugf_download_accumulators(ugf_handle* h, double* acc, double* accSpecies, double* timeAvCounter, int64_t* nAvTimeSteps);
/* The internal-mode accumulators [nCells][nSpecies][UGF_NINT] (zeros when no species carries vibrational modes or several
 * electronic levels). */
int ugf_download_internal_accumulators(ugf_handle* h, double* accInt);
/* Per-phase device time of the last ugf_step in ms (UGF_NPHASE values): inflow, move, sort, cell (gather + sample +
 * field accumulation), collide (NTC), relax (BGK family), fields (wall accumulation). */
#define UGF_NPHASE 7
int ugf_phase_times(ugf_handle* h, double* ms);
/* Number of kernel launches issued by this handle so far. */
int ugf_launch_count(ugf_handle* h, int64_t* n);
/* Bytes this handle has moved between host and device since ugf_create: every explicit host -> device and device -> host
 * copy, and the argument blocks of the per-step kernels (what a launch sends besides its grid).  Measurement only: bench.py
 * differences two calls around its end-to-end loop.  Any pointer may be NULL.  (No reference analogue.) */
int ugf_transfer_bytes(ugf_handle* h, int64_t* h2d, int64_t* d2h, int64_t* kernelArgs);
/* Page-locked host memory for callers that hand the parcels over every step (the solver owns the parcel list, as
 * uniGasCloud does: U/clouds/uniGasCloudI.H:375-378): ugf_upload_parcels / ugf_download_parcels from / into such
 * buffers run as asynchronous DMA at PCIe rate instead of through the driver's pageable staging.  Freed by ugf_host_free
 * or with the handle. */
int ugf_host_alloc(ugf_handle* h, int64_t bytes, void** ptr);
int ugf_host_free(ugf_handle* h, void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* UGF_H */
