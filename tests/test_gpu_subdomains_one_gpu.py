"""Decomposed runs on ONE GPU: every subdomain of a decomposed case gets its own libugf handle on the same device and
`exchange.LocalSubdomains` carries the parcels between them (device buffers, `ugf_migrate_pack` -> `ugf_migrate_unpack` ->
`ugf_move_received`, exact termination rule).  This runs the multi-rank code of the product - processor patches in the move
kernel, migrant lists, pack / unpack kernels, resumed tracks with the stored step fraction, cell weights travelling with the
parcels - on a single-GPU lease, against the oracle doing the same decomposition (and, collision-free, against the undecomposed
run).  The two-GPU transports (NCCL, NVLink peer memory) are tests/test_gpu_multirank.py."""
import numpy as np
import pytest

from unigasfoam_b200 import cases
from unigasfoam_b200.exchange import LocalSubdomains
from unigasfoam_b200.mesh import decompose, slab_partition

pytestmark = pytest.mark.gpu


def split(one, n_ranks, Cloud, axis=0, **kw):
    subs = decompose(one.mesh, slab_partition(one.mesh, n_ranks, axis=axis), n_ranks)
    clouds = []
    for r, sm in enumerate(subs):
        inv = np.full(one.mesh.n_cells, -1, np.int64)
        inv[np.asarray(sm.cell_map)] = np.arange(len(sm.cell_map))
        sel = np.nonzero(inv[one.cell] >= 0)[0]
        c = cases.Case("part", sm, one.uniGasProperties, one.boundariesDict, one.deltaT, one.position[sel], one.U[sel],
                       inv[one.cell[sel]].astype(np.int32), None if one.typeId is None else one.typeId[sel],
                       None if one.ERot is None else one.ERot[sel], one.sigmaTcRMax)
        if one.cellWeightFactor is not None:
            c.cellWeightFactor = np.ascontiguousarray(one.cellWeightFactor[np.asarray(sm.cell_map)])
        clouds.append(c.make_cloud(Cloud, rank=r, nRanks=n_ranks, parcelCapacity=3 * one.n_parcels, **kw))
    return LocalSubdomains(clouds, subs), clouds, subs


def sorted_state(clouds):
    a = np.concatenate([np.column_stack([c.parcels()["position"], c.parcels()["U"]]) for c in clouds])
    return a[np.lexsort(a.T[::-1])]


def test_three_slabs_on_one_gpu_equal_the_single_domain_run(GpuCloud):
    """Collision-free periodic channel in three slabs (processor and processorCyclic neighbours): bit for bit the parcels of the
    undecomposed GPU run."""
    one = cases.couette(nx=36, ny=12, ppc=15, binary="noDSMCCollision", Kn=0.5)
    for e in one.boundariesDict["uniGasPatchBoundaries"]:  # wall draws are keyed by the parcel's array index, which differs between the runs
        e["boundaryModel"] = "uniGasSpecularWallPatch"
    one.deltaT *= 4.0  # Courant > 1: several cells per step, parcels cross a whole slab corner now and then
    L, clouds, _ = split(one, 3, GpuCloud)
    ref = one.make_cloud(GpuCloud)
    L.evolve(8)
    ref.evolve(8)
    assert L.rounds >= 8
    assert sum(c.counters()["stuck"] for c in clouds) == 0
    np.testing.assert_array_equal(sorted_state(clouds), sorted_state([ref]))


@pytest.mark.parametrize("binary", ["variableHardSphere", "LarsenBorgnakkeVariableHardSphere"])
def test_decomposed_collisions_on_one_gpu_in_lockstep_with_the_oracle(GpuCloud, OracleCloud, binary):
    sp = ("N2", cases.NITROGEN) if binary.startswith("Larsen") else ("Ar", cases.ARGON_GUIDE)
    one = cases.couette(nx=32, ny=12, ppc=20, binary=binary, Kn=0.3, species=sp)
    Lg, cg, _ = split(one, 2, GpuCloud)
    Lr, cr, _ = split(one, 2, OracleCloud)
    for _ in range(6):
        Lg.evolve(1); Lr.evolve(1)
        for g, r in zip(cg, cr):
            a, b = g.counters(), r.counters()
            for k in ("nParcels", "collisionCandidates", "collisions", "wallHits", "migrated", "stuck"):
                assert a[k] == b[k], (k, a[k], b[k])
    assert sum(r.counters()["collisions"] for r in cr) > 0 and Lr.rounds >= 6
    for g, r in zip(cg, cr):
        pg, pr = g.parcels(), r.parcels()
        assert np.array_equal(pg["cell"], pr["cell"])
        assert (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() > 0.99


def test_decomposed_weighted_cylinder_with_inflow_on_one_gpu(GpuCloud, OracleCloud):
    """Cell weighting across processor patches (the carried factor travels in the migration record), inflow and outflow patches
    cut by the decomposition, diffuse body: four slabs along the body on one GPU in lockstep with the oracle."""
    one = cases.cylinder(nr=12, ntheta=32, ppc=20, seed=4, cellWeightFactor=("particlesPerSubCell", 20))
    Lg, cg, _ = split(one, 4, GpuCloud, axis=1)
    Lr, cr, _ = split(one, 4, OracleCloud, axis=1)
    tot = dict(cloned=0, inserted=0, migrated=0)
    for _ in range(8):
        Lg.evolve(1, inflow=True); Lr.evolve(1, inflow=True)
        for g, r in zip(cg, cr):
            a, b = g.counters(), r.counters()
            for k in ("nParcels", "inserted", "deleted", "cloned", "weightDeleted", "collisions", "wallHits", "migrated", "stuck"):
                assert a[k] == b[k], (k, a[k], b[k])
            for k in tot:
                tot[k] += b[k]
    assert tot["cloned"] > 0 and tot["inserted"] > 0 and tot["migrated"] > 0
    for g, r in zip(cg, cr):
        assert np.array_equal(g.parcels()["cell"], r.parcels()["cell"])
