// box_driver.cpp - a C++ host program on nothing but include/ugf.h: the call sequence a uniGasFoam-side shim makes
// (INTEGRATION.md), on a closed box of argon.  It flattens a hex mesh the way the shim flattens a polyMesh (owner /
// neighbour, face areas / centres, cell -> faces CSR, cell volumes / centres / bounding boxes, patch table), attaches
// specular wall models, uploads a Maxwellian cloud, runs uniGasCloud::evolve() steps through ugf_step and checks what
// must hold: parcel count and kinetic energy conserved between specular walls (collisions conserve energy pair by pair),
// momentum components reversed only by walls, collision count near 1/2 N nu dt (Bird 4.64).  Exit code 0 = all checks pass.
//
// Build:  g++ -O2 -std=c++17 -I include examples/box_driver.cpp -L unigasfoam_b200 -lugf -Wl,-rpath,'$ORIGIN/../unigasfoam_b200' -o examples/box_driver
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "ugf.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        if ((call) != 0) {                                                               \
            std::fprintf(stderr, "%s failed: %s\n", #call, ugf_last_error(h));           \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 12;           // cells per direction
    const long long nParcels = argc > 2 ? std::atoll(argv[2]) : 60000;
    const int steps = argc > 3 ? std::atoi(argv[3]) : 40;
    const double kB = 1.38065e-23, PI = 3.14159265358979323846;
    const double mass = 66.3e-27, d = 4.17e-10, omega = 0.81, Tref = 273.0, T0 = 300.0, nDen = 1e20;
    const double lambda = 1.0 / (std::sqrt(2.0) * PI * d * d * nDen * std::pow(Tref / T0, omega - 0.5));  // Bird 4.65
    const double dx = 0.5 * lambda, L = n * dx;
    const double nu = 4.0 * d * d * nDen * std::sqrt(PI * kB * Tref / mass) * std::pow(T0 / Tref, 1.0 - omega);  // Bird 4.64
    const double dt = 0.2 / nu;
    const double FN = nDen * L * L * L / (double)nParcels;

    // ---- hex mesh, OpenFOAM ordering: internal faces first, then the six patches --------------------------------
    const int nC = n * n * n;
    auto cid = [&](int i, int j, int k) { return i + n * (j + n * k); };
    struct Face { int own, nei; double S[3], C[3]; };
    std::vector<Face> internal, bnd[6];
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i)
                for (int dir = 0; dir < 3; ++dir) {
                    const int idx[3] = {i, j, k};
                    double c[3] = {(i + 0.5) * dx, (j + 0.5) * dx, (k + 0.5) * dx};
                    // the face on the + side of direction dir
                    Face f{};
                    f.own = cid(i, j, k);
                    f.S[dir] = dx * dx;
                    for (int q = 0; q < 3; ++q) f.C[q] = c[q];
                    f.C[dir] += 0.5 * dx;
                    if (idx[dir] + 1 < n) {
                        int nb[3] = {i, j, k};
                        nb[dir]++;
                        f.nei = cid(nb[0], nb[1], nb[2]);
                        internal.push_back(f);
                    } else {
                        f.nei = -1;
                        bnd[2 * dir + 1].push_back(f);
                    }
                    if (idx[dir] == 0) {  // and the boundary face on the - side
                        Face g{};
                        g.own = cid(i, j, k);
                        g.nei = -1;
                        g.S[dir] = -dx * dx;
                        for (int q = 0; q < 3; ++q) g.C[q] = c[q];
                        g.C[dir] -= 0.5 * dx;
                        bnd[2 * dir].push_back(g);
                    }
                }
    std::vector<Face> faces = internal;
    std::vector<int32_t> patchStart(6), patchSize(6), patchKind(6, UGF_PATCH_WALL), patchPartner(6, -1);
    std::vector<double> patchSep(18, 0.0);
    for (int p = 0; p < 6; ++p) {
        patchStart[p] = (int32_t)faces.size();
        patchSize[p] = (int32_t)bnd[p].size();
        faces.insert(faces.end(), bnd[p].begin(), bnd[p].end());
    }
    const int nF = (int)faces.size(), nI = (int)internal.size();
    std::vector<int32_t> owner(nF), neighbour(nI);
    std::vector<double> Sf(3 * (size_t)nF), Cf(3 * (size_t)nF);
    std::vector<std::vector<int32_t>> cf(nC);
    for (int f = 0; f < nF; ++f) {
        owner[f] = faces[f].own;
        cf[faces[f].own].push_back(f);
        if (f < nI) { neighbour[f] = faces[f].nei; cf[faces[f].nei].push_back(f); }
        for (int q = 0; q < 3; ++q) { Sf[3 * (size_t)f + q] = faces[f].S[q]; Cf[3 * (size_t)f + q] = faces[f].C[q]; }
    }
    std::vector<int32_t> cfOff(nC + 1, 0), cfFlat;
    std::vector<double> vol(nC, dx * dx * dx), cc(3 * (size_t)nC), bbMin(3 * (size_t)nC), bbMax(3 * (size_t)nC);
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                const int c = cid(i, j, k);
                const int idx[3] = {i, j, k};
                for (int q = 0; q < 3; ++q) {
                    cc[3 * (size_t)c + q] = (idx[q] + 0.5) * dx;
                    bbMin[3 * (size_t)c + q] = idx[q] * dx;
                    bbMax[3 * (size_t)c + q] = (idx[q] + 1) * dx;
                }
            }
    for (int c = 0; c < nC; ++c) {
        cfFlat.insert(cfFlat.end(), cf[c].begin(), cf[c].end());
        cfOff[c + 1] = (int32_t)cfFlat.size();
    }

    // ---- the ABI ------------------------------------------------------------------------------------------------
    ugf_handle* h = nullptr;
    ugf_config cfg{};
    cfg.abiVersion = UGF_ABI_VERSION;
    cfg.device = 0;
    cfg.seed = 20261017;
    cfg.nParticle = FN;
    cfg.deltaT = dt;
    cfg.solutionD[0] = cfg.solutionD[1] = cfg.solutionD[2] = 1;
    cfg.collisionModel = UGF_COLL_DSMC;
    cfg.partnerModel = UGF_PARTNER_NTC;
    cfg.binaryModel = UGF_BINARY_VHS;
    cfg.bgkModel = UGF_BGK_NONE;
    cfg.nSubCycles = 1;
    cfg.Tref = Tref;
    cfg.theta = 1.0;
    cfg.rotationalRelaxationCollisionNumber = 5.0;
    cfg.electronicRelaxationCollisionNumber = 500.0;
    cfg.parcelCapacity = nParcels + nParcels / 4 + 1024;
    cfg.sampleInterval = 1;
    cfg.measureWalls = 1;
    cfg.rank = 0;
    cfg.nRanks = 1;
    if (ugf_create(&cfg, &h) != 0) {
        std::fprintf(stderr, "ugf_create failed: %s\n", ugf_last_error(nullptr));
        return 2;
    }
    ugf_species ar{};
    ar.mass = mass; ar.d = d; ar.omega = omega; ar.alpha = 1.0;
    ar.nElectronicLevels = 1; ar.degeneracy[0] = 1;
    CHECK(ugf_set_species(h, 1, &ar));
    ugf_mesh m{};
    m.nCells = nC; m.nFaces = nF; m.nInternalFaces = nI; m.nPatches = 6; m.nPoints = 0;
    m.owner = owner.data(); m.neighbour = neighbour.data(); m.faceAreas = Sf.data(); m.faceCentres = Cf.data();
    m.cellFaceOffsets = cfOff.data(); m.cellFaces = cfFlat.data(); m.cellVolumes = vol.data(); m.cellCentres = cc.data();
    m.cellBbMin = bbMin.data(); m.cellBbMax = bbMax.data();
    m.patchStart = patchStart.data(); m.patchSize = patchSize.data(); m.patchKind = patchKind.data();
    m.patchPartner = patchPartner.data(); m.patchSeparation = patchSep.data();
    CHECK(ugf_set_mesh(h, &m));
    for (int p = 0; p < 6; ++p) CHECK(ugf_set_patch_model(h, p, UGF_WALL_SPECULAR, nullptr, 0));

    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    std::normal_distribution<double> gauss(0.0, std::sqrt(kB * T0 / mass));
    std::vector<double> x(nParcels), y(nParcels), z(nParcels), ux(nParcels), uy(nParcels), uz(nParcels);
    std::vector<int32_t> cell(nParcels);
    for (long long i = 0; i < nParcels; ++i) {  // cell-major, like the reference's mesh fill
        const int c = (int)(i * nC / nParcels);
        cell[i] = c;
        x[i] = bbMin[3 * (size_t)c] + uni(rng) * dx; y[i] = bbMin[3 * (size_t)c + 1] + uni(rng) * dx; z[i] = bbMin[3 * (size_t)c + 2] + uni(rng) * dx;
        ux[i] = gauss(rng); uy[i] = gauss(rng); uz[i] = gauss(rng);
    }
    ugf_parcels P{};
    P.n = nParcels; P.x = x.data(); P.y = y.data(); P.z = z.data(); P.Ux = ux.data(); P.Uy = uy.data(); P.Uz = uz.data(); P.cell = cell.data();
    CHECK(ugf_upload_parcels(h, &P));
    std::vector<double> sig(nC, PI * d * d * std::sqrt(2.0 * kB * T0 / mass));  // uniGasMeshFill.C:284-296
    CHECK(ugf_upload_cell_state(h, sig.data(), nullptr, nullptr, nullptr));

    ugf_counters c0{}, c{};
    CHECK(ugf_counters_get(h, &c0));
    long long collisions = 0;
    for (int s = 0; s < steps; ++s) {
        CHECK(ugf_step(h, 1));
        CHECK(ugf_counters_get(h, &c));
        collisions += c.collisions;
    }
    const double eDrift = std::fabs(c.linearKineticEnergy / c0.linearKineticEnergy - 1.0);
    const double expectColl = 0.5 * (double)nParcels * nu * dt * steps;
    const double T = 2.0 * c.linearKineticEnergy / (3.0 * kB * (double)c.nParcels);
    std::printf("steps %lld parcels %lld collisions %lld (kinetic theory %.0f) T %.2f K energy drift %.2e wall hits/step %lld stuck %lld\n",
                (long long)c.step, (long long)c.nParcels, collisions, expectColl, T, eDrift, (long long)c.wallHits, (long long)c.stuck);
    int rc = 0;
    if (c.nParcels != nParcels || c.stuck != 0) { std::fprintf(stderr, "parcels lost\n"); rc = 1; }
    if (eDrift > 1e-10) { std::fprintf(stderr, "energy not conserved\n"); rc = 1; }
    if (std::fabs((double)collisions / expectColl - 1.0) > 0.08) { std::fprintf(stderr, "collision rate off\n"); rc = 1; }
    std::vector<double> fields((size_t)nC * UGF_NFIELD);
    CHECK(ugf_download_fields(h, fields.data(), nullptr, 0));
    double nMean = 0;
    for (int cI = 0; cI < nC; ++cI) nMean += fields[(size_t)cI * UGF_NFIELD + 1] / nC;
    if (std::fabs(nMean / nDen - 1.0) > 0.01) { std::fprintf(stderr, "rhoN field off: %g\n", nMean); rc = 1; }
    CHECK(ugf_destroy(h));
    std::printf(rc == 0 ? "box_driver: ok\n" : "box_driver: FAILED\n");
    return rc;
}
