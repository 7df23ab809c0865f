"""World-size-2 run of the multi-rank path on CPU (gloo): slab decomposition with processor / processorCyclic
patches, pack -> all_to_all -> unpack -> resume tracking, iterate until nothing is in flight.  The CPU oracle stands
in for the device library behind the same Exchanger; the result must equal the single-domain run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unigasfoam_b200 import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, steps, out, slots, tail=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_cloud import OracleCloud
    from unigasfoam_b200.exchange import Exchanger, SlotExchanger, evolve_distributed
    case = cases.couette(nx=12, ny=8, ppc=12, rank=rank, n_ranks=world, binary="noDSMCCollision")
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        e["boundaryModel"] = "uniGasSpecularWallPatch"
    case.deltaT *= 40.0 if tail else 6.0  # several cells per step: parcels cross slabs, some wrap around the periodic end (tail: a whole slab and more)
    cl = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels, rank=rank, nRanks=world)
    ex = SlotExchanger(cl, case.mesh, rank, world, slot_capacity=600, cuda=False) if slots else Exchanger(cl, case.mesh, rank, world, cuda=False)
    n0 = torch.tensor([cl.size()])
    dist.all_reduce(n0)
    if tail:  # one unsynchronised round (not enough here: some parcels wrap through two patches), then the exact tail
        evolve_distributed(cl, ex, steps, fixed_rounds=1, exact_tail=True)
        assert ex.tail_rounds > 0
    else:
        evolve_distributed(cl, ex, steps, fixed_rounds=2 if slots else None)
    if slots and not tail:
        ex.check_settled()
    p = cl.parcels()
    n1 = torch.tensor([cl.size()])
    dist.all_reduce(n1)
    res = dict(rank=rank, n0=int(n0), n1=int(n1), pos=p["position"], U=p["U"], cell=p["cell"], rounds=ex.rounds, sent=getattr(ex, 'sent', 10 ** 6),
               stuck=cl.counters()["stuck"], x0=rank * case.meta["Lx"], Lx=case.meta["Lx"],
               init=(case.position.copy(), case.U.copy()))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("slots", [False, True, "tail"], ids=["exact-count", "fixed-slot", "fixed-slot-exact-tail"])
def test_two_rank_migration_matches_single_domain(tmp_path, OracleCloud, slots):
    world, steps = 2, 5
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), steps, out, bool(slots), slots == "tail"), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    assert res[0]["n0"] == res[0]["n1"] == sum(len(r["cell"]) for r in res)  # parcels conserved across ranks
    assert all(r["stuck"] == 0 for r in res)
    assert sum(r["sent"] for r in res) > 50 and all(r["rounds"] >= steps for r in res)
    for r in res:  # every parcel ended inside its rank's slab
        assert (r["pos"][:, 0] >= r["x0"] - 1e-12).all() and (r["pos"][:, 0] <= r["x0"] + r["Lx"] * (1 + 1e-12)).all()
    # single-domain reference: the same parcels in one periodic channel of twice the length
    one = cases.couette(nx=24, ny=8, ppc=12, binary="noDSMCCollision")
    for e in one.boundariesDict["uniGasPatchBoundaries"]:
        e["boundaryModel"] = "uniGasSpecularWallPatch"
    one.deltaT *= 40.0 if slots == "tail" else 6.0
    pos = np.concatenate([r["init"][0] for r in res])
    U = np.concatenate([r["init"][1] for r in res])
    dx = one.meta["Lx"] / 24
    dy = one.meta["H"] / 8
    ci = np.minimum((pos[:, 0] / dx).astype(int), 23) + 24 * np.minimum((pos[:, 1] / dy).astype(int), 7)
    one.position, one.U, one.cell, one.typeId = pos, U, ci.astype(np.int32), np.zeros(len(ci), np.int32)
    ref = one.make_cloud(OracleCloud, parcelCapacity=4 * len(ci))
    ref.evolve(steps)
    pr = ref.parcels()
    got = np.concatenate([np.column_stack([r["pos"], r["U"]]) for r in res])
    want = np.column_stack([pr["position"], pr["U"]])
    key = lambda a: a[np.lexsort(np.round(a / (np.abs(a).max(0) + 1e-300), 9).T[::-1])]
    g, w = key(got), key(want)
    assert g.shape == w.shape
    err = np.abs(g - w) / (np.abs(w).max(0) + 1e-300)
    assert err.max() < 1e-11, (err.max(), np.unravel_index(err.argmax(), err.shape))


def _worker_weighted(rank, world, port, steps, out):
    """Decomposed, cell-weighted cylinder (the reference tutorial's setting) on the oracle: a parcel that crosses the
    cut carries its weight factor in the migration record and is cloned / deleted against the factor of the cell it
    lands in on the other rank."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_cloud import OracleCloud
    from unigasfoam_b200 import mesh as ugmesh
    from unigasfoam_b200.exchange import SlotExchanger, evolve_distributed
    case = cases.cylinder(nr=12, ntheta=24, ppc=20, cellWeightFactor=("particlesPerSubCell", 20))
    part = ugmesh.slab_partition(case.mesh, world, axis=1)
    sub = ugmesh.decompose(case.mesh, part, world)[rank]
    g2l = np.full(case.mesh.n_cells, -1)
    g2l[sub.cell_map] = np.arange(sub.n_cells)
    sel = part[case.cell] == rank
    W = case.cellWeightFactor[sub.cell_map]
    cl = OracleCloud(sub, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=4 * int(sel.sum()) + 4096, rank=rank, nRanks=world)
    cl.setCellState(cellWeightFactor=W)
    cl.setParcels(case.position[sel], case.U[sel], g2l[case.cell[sel]])
    cl.setCellState(sigmaTcRMax=case.sigmaTcRMax)
    ex = SlotExchanger(cl, sub, rank, world, slot_capacity=4000, cuda=False)
    tally = np.zeros(3)
    for _ in range(steps):
        evolve_distributed(cl, ex, 1, inflow=True)
        c = cl.counters()
        tally += [c["cloned"], c["weightDeleted"], c["migrated"]]
    p = cl.parcels()
    res = dict(ok=bool(np.array_equal(p["cellWeight"], W[p["cell"]])), tally=tally, n=len(p["cell"]), stuck=cl.counters()["stuck"],
               real=float(W[p["cell"]].sum()), real0=float(case.cellWeightFactor[case.cell[sel]].sum()))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_cell_weighted_cylinder(tmp_path, OracleCloud):
    world, steps = 2, 12
    out = str(tmp_path / "resw.pt")
    mp.spawn(_worker_weighted, args=(world, _free_port(), steps, out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    assert all(r["ok"] and r["stuck"] == 0 for r in res), res
    tot = sum(r["tally"] for r in res)
    assert (tot > 0).all(), tot  # clones, deletions and migrations all happened
    real, real0 = sum(r["real"] for r in res), sum(r["real0"] for r in res)
    assert abs(real / real0 - 1) < 0.05  # real molecules (sum of weights) stay put over a few steps of a uniform free stream


def _worker_halo(rank, world, port, out):
    """fvc::average(fvc::interpolate) and fvc::smooth on a decomposed mesh with the processor-face halo (SURVEY 8e:
    'hybrid-mask and adaptation smoothing need a 1-cell halo of cell fields')."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unigasfoam_b200 import mesh as ugmesh
    from unigasfoam_b200.adapter import FaceOperators
    from unigasfoam_b200.exchange import ProcessorHalo, group_reducers
    res = {}
    for name, (full, part) in _halo_meshes(world).items():
        sub = ugmesh.decompose(full, part, world)[rank]
        halo = ProcessorHalo(sub, rank, world)
        _, rmax = group_reducers()
        ops = FaceOperators(sub, halo, rmax)
        f, v = _halo_fields(full.n_cells)
        fs, vs = f[sub.cell_map], v[sub.cell_map]
        a1 = ops.average_interpolate(fs)
        a2 = ops.average_interpolate(ops.average_interpolate(fs))  # the halo carries the updated values of the second pass
        av = ops.average_interpolate(vs, vector=True)
        sm = ops.smooth(fs, 1.3)
        zg = FaceOperators(sub).average_interpolate(fs)  # no halo: zero-gradient processor faces
        res[name] = dict(cells=sub.cell_map, a1=a1, a2=a2, av=av, sm=sm, zg=zg, calls=halo.calls, nproc=halo.n)
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


def _halo_meshes(world):
    from unigasfoam_b200 import mesh as ugmesh
    c = cases.couette(nx=12, ny=8, ppc=1, binary="noDSMCCollision").mesh  # x cyclic: the cut gives processor + processorCyclic patches
    a = ugmesh.half_annulus_mesh(8, 16, 0.1, 0.5, 0.01, grading=4.0)  # graded: interpolation weights != 1/2, symmetry planes
    return {"couette": (c, ugmesh.slab_partition(c, world, axis=0)), "annulus": (a, ugmesh.slab_partition(a, world, axis=1))}


def _halo_fields(n):
    rng = np.random.default_rng(11)
    return np.exp(3.0 * rng.standard_normal(n)), rng.standard_normal((n, 3))


@pytest.mark.timeout(300)
def test_two_rank_smoothing_operators_match_single_domain(tmp_path):
    from unigasfoam_b200.adapter import FaceOperators
    world = 2
    out = str(tmp_path / "halo.pt")
    mp.spawn(_worker_halo, args=(world, _free_port(), out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    for name, (full, _) in _halo_meshes(world).items():
        ops = FaceOperators(full)
        f, v = _halo_fields(full.n_cells)
        want = dict(a1=ops.average_interpolate(f), a2=ops.average_interpolate(ops.average_interpolate(f)),
                    av=ops.average_interpolate(v, vector=True), sm=ops.smooth(f, 1.3))
        assert all(r[name]["nproc"] > 0 for r in res)
        assert len({r[name]["calls"] for r in res}) == 1  # the smoothing wave stopped on both ranks in the same sweep
        for k, w in want.items():
            got = np.empty_like(w)
            for r in res:
                got[r[name]["cells"]] = r[name][k]
            if k == "sm":
                assert np.array_equal(got, w), (name, k)  # maxima of the same quotients: bit-identical
                assert (got >= f).all() and (got > f).any()
            else:
                assert np.allclose(got, w, rtol=1e-13, atol=1e-13 * np.abs(w).max()), (name, k)
        zg = np.empty(full.n_cells)
        for r in res:
            zg[r[name]["cells"]] = r[name]["zg"]
        assert not np.allclose(zg, want["a1"], rtol=1e-6)  # without the halo the cut is visible in the result


def _adaptive_props(full):
    case = cases.cylinder(nr=8, ntheta=16, ppc=5, binary="noDSMCCollision")
    props = dict(case.uniGasProperties)
    props["adaptiveSimulation"] = True
    props["adaptiveProperties"] = dict(timeStepAdaptation=True, subCellAdaptation=True, adaptationInterval=10, smoothingPasses=6)
    return props, case.uniGasProperties["nEquivalentParticles"] if "nEquivalentParticles" in case.uniGasProperties else 1.0


def _initial_state(mesh_cells_xyz):
    """A non-uniform initial state (density over two decades, temperature ramp) so that the smoothing passes matter."""
    x, y = mesh_cells_xyz[:, 0], mesh_cells_xyz[:, 1]
    r = np.hypot(x, y)
    n = 1e21 * np.exp(-6.0 * (r - r.min()) / (r.max() - r.min())) * (1.0 + 0.5 * np.sin(2.3 * np.arctan2(y, x) + 0.4))  # not symmetric about the cut
    T = 200.0 + 300.0 * (r - r.min()) / (r.max() - r.min())
    U = np.column_stack([300.0 * np.cos(np.arctan2(y, x)), 100.0 * np.ones_like(x), np.zeros_like(x)])
    return n, T, U


def _stand_in(mesh, dt):
    from types import SimpleNamespace
    return SimpleNamespace(mesh=mesh, cellWeighted=False, _subCellLevels=None, _cellWeightFactor=None, _adapter=None,
                           cfg=SimpleNamespace(deltaT=dt, nParticle=1e10))


def _worker_adapter(rank, world, port, out):
    """uniGasDynamicAdapter::setInitialConfiguration (:708-775) on a decomposed mesh: ratios smoothed through the halo,
    time step reduced over the ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unigasfoam_b200 import mesh as ugmesh
    from unigasfoam_b200.adapter import UniGasDynamicAdapter
    from unigasfoam_b200.exchange import ProcessorHalo, group_reducers
    full = _halo_meshes(world)["annulus"][0]
    sub = ugmesh.decompose(full, ugmesh.slab_partition(full, world, axis=1), world)[rank]
    props, _ = _adaptive_props(full)
    rmin, rmax = group_reducers()
    n, T, U = _initial_state(full.cell_centres)
    res = {}
    for label, halo in (("halo", ProcessorHalo(sub, rank, world)), ("cut", None)):
        ad = UniGasDynamicAdapter(_stand_in(sub, 1e-6), props, reduce_min=rmin, reduce_max=rmax, halo=halo)
        dt, levels = ad.set_initial_configuration([n[sub.cell_map]], T[sub.cell_map], U[sub.cell_map])
        res[label] = dict(dt=dt, levels=levels, csr=ad.prevCellSizeMFPRatio)
    res["cells"] = sub.cell_map
    from unigasfoam_b200.exchange import max_imbalance
    res["imbalance"] = max_imbalance(100 + 50 * rank)   # 100 and 150 parcels: 20 % off the mean of 125
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_adapter_initial_configuration_matches_single_domain(tmp_path):
    from unigasfoam_b200.adapter import UniGasDynamicAdapter
    world = 2
    out = str(tmp_path / "adapt.pt")
    mp.spawn(_worker_adapter, args=(world, _free_port(), out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    full = _halo_meshes(world)["annulus"][0]
    props, _ = _adaptive_props(full)
    n, T, U = _initial_state(full.cell_centres)
    ad = UniGasDynamicAdapter(_stand_in(full, 1e-6), props)
    dt, levels = ad.set_initial_configuration([n], T, U)
    assert dt != 1e-6 and levels.max() > levels.min()          # the state does drive both decisions
    assert all(r["halo"]["dt"] == res[0]["halo"]["dt"] for r in res)
    assert res[0]["halo"]["dt"] == pytest.approx(dt, rel=1e-12)
    got_l, got_c, cut_c = np.empty_like(levels), np.empty_like(ad.prevCellSizeMFPRatio), np.empty_like(ad.prevCellSizeMFPRatio)
    for r in res:
        got_l[r["cells"]] = r["halo"]["levels"]
        got_c[r["cells"]] = r["halo"]["csr"]
        cut_c[r["cells"]] = r["cut"]["csr"]
    assert np.allclose(got_c, ad.prevCellSizeMFPRatio, rtol=1e-12)
    assert np.array_equal(got_l, levels)
    assert not np.allclose(cut_c, ad.prevCellSizeMFPRatio, rtol=1e-3)   # zero-gradient processor faces leave a seam
    from unigasfoam_b200 import mesh as ugmesh
    assert all(r["imbalance"] == pytest.approx(20.0) for r in res) and ugmesh.load_imbalance([100, 150]) == pytest.approx(20.0)


def _worker_blocks(rank, world, port, steps, out):
    """2 x 2 block decomposition (decomposePar simple n (2 2 1)) of the periodic channel: every rank has a processor
    neighbour in x (also through the cyclic pair) and one in y; parcels that leave through a corner need two transfers."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_cloud import OracleCloud
    from unigasfoam_b200 import mesh as ugmesh
    from unigasfoam_b200.exchange import Exchanger, evolve_distributed, max_imbalance
    case = _block_case()
    part = ugmesh.block_partition(case.mesh, (2, 2, 1))
    sub = ugmesh.decompose(case.mesh, part, world)[rank]
    g2l = np.full(case.mesh.n_cells, -1)
    g2l[sub.cell_map] = np.arange(sub.n_cells)
    sel = part[case.cell] == rank
    cl = OracleCloud(sub, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=4 * int(sel.sum()) + 4096, rank=rank, nRanks=world)
    cl.setParcels(case.position[sel], case.U[sel], g2l[case.cell[sel]])
    ex = Exchanger(cl, sub, rank, world, cuda=False)
    evolve_distributed(cl, ex, steps)
    p = cl.parcels()
    res = dict(pos=p["position"], U=p["U"], cells=sub.cell_map[p["cell"]], rounds=ex.rounds, stuck=cl.counters()["stuck"],
               nproc=sum(q.kind == "processor" for q in sub.patches), imbalance=max_imbalance(cl.size()))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


def _block_case():
    case = cases.couette(nx=12, ny=8, ppc=10, binary="noDSMCCollision", seed=17)
    for e in case.boundariesDict["uniGasPatchBoundaries"]:
        e["boundaryModel"] = "uniGasSpecularWallPatch"
    case.deltaT *= 5.0
    return case


@pytest.mark.timeout(300)
def test_four_rank_block_decomposition_matches_single_domain(tmp_path, OracleCloud):
    world, steps = 4, 4
    out = str(tmp_path / "blocks.pt")
    mp.spawn(_worker_blocks, args=(world, _free_port(), steps, out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    case = _block_case()
    ref = case.make_cloud(OracleCloud, parcelCapacity=4 * case.n_parcels)
    ref.evolve(steps)
    pr = ref.parcels()
    assert all(r["stuck"] == 0 for r in res) and all(r["nproc"] >= 2 for r in res) and max(r["rounds"] for r in res) >= steps
    assert sum(len(r["cells"]) for r in res) == len(pr["cell"])
    got = np.concatenate([np.column_stack([r["pos"], r["U"], r["cells"]]) for r in res])
    want = np.column_stack([pr["position"], pr["U"], pr["cell"]])
    key = lambda a: a[np.lexsort(np.round(a[:, :6] / (np.abs(a[:, :6]).max(0) + 1e-300), 9).T[::-1])]
    g, w = key(got), key(want)
    assert np.array_equal(g[:, 6], w[:, 6])                                  # every parcel in the same (global) cell
    err = np.abs(g[:, :6] - w[:, :6]) / (np.abs(w[:, :6]).max(0) + 1e-300)
    assert err.max() < 1e-11
    assert res[0]["imbalance"] < 15.0
