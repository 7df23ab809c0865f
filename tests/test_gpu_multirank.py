"""Two GPUs, one process each: slab decomposition, device-side pack/unpack, all_to_all over NCCL.  Each rank also
runs the CPU oracle through the same Exchanger (gloo) and the two must agree bit for bit on collision-free
tracking.  Skipped on single-GPU boxes (the CPU twin of this test is tests/test_multirank_gloo.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unigasfoam_b200 import cases

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    meta = dist.new_group(backend="gloo")
    from oracle.oracle_cloud import OracleCloud
    from unigasfoam_b200.cloud import UniGasCloud
    from unigasfoam_b200.exchange import Exchanger, PeerExchanger, SlotExchanger, evolve_distributed
    ok = True
    msgs = []
    for binary, steps, slots in (("noDSMCCollision", 6, False), ("variableHardSphere", 4, False), ("noDSMCCollision", 6, True), ("variableHardSphere", 4, True),
                                 ("noDSMCCollision", 6, "peer"), ("variableHardSphere", 4, "peer")):
        case = cases.couette(nx=32, ny=16, ppc=20, rank=rank, n_ranks=world, binary=binary, Kn=0.5)
        if binary == "noDSMCCollision":
            for e in case.boundariesDict["uniGasPatchBoundaries"]:
                e["boundaryModel"] = "uniGasSpecularWallPatch"
            case.deltaT *= 5.0
        kw = dict(parcelCapacity=4 * case.n_parcels, rank=rank, nRanks=world)
        g = case.make_cloud(UniGasCloud, device=rank, **kw)
        r = case.make_cloud(OracleCloud, **kw)
        if slots == "peer":  # NVLink peer-memory transfer on the device path
            exg = PeerExchanger(g, case.mesh, rank, world, slot_capacity=2000, group=None, meta_group=meta)
            exr = SlotExchanger(r, case.mesh, rank, world, slot_capacity=2000, group=meta, cuda=False)
        elif slots:
            exg = SlotExchanger(g, case.mesh, rank, world, slot_capacity=2000, group=None, cuda=True)
            exr = SlotExchanger(r, case.mesh, rank, world, slot_capacity=2000, group=meta, cuda=False)
        else:
            exg = Exchanger(g, case.mesh, rank, world, data_group=None, meta_group=meta, cuda=True)
            exr = Exchanger(r, case.mesh, rank, world, data_group=meta, meta_group=meta, cuda=False)
        evolve_distributed(g, exg, steps, fixed_rounds=2 if slots else None)  # fixed-round mode on the device path
        if slots:
            exg.check_settled()
        evolve_distributed(r, exr, steps)
        pg, pr = g.parcels(), r.parcels()
        same_cells = np.array_equal(pg["cell"], pr["cell"])
        if binary == "noDSMCCollision":
            good = same_cells and np.array_equal(pg["position"], pr["position"]) and np.array_equal(pg["U"], pr["U"])
        else:
            close = (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() if same_cells else 0.0
            good = same_cells and close > 0.999 and g.counters()["collisions"] == r.counters()["collisions"]
        msgs.append((binary, slots, good, exg.rounds, exr.rounds, g.size(), r.size()))
        ok = ok and good and (slots or exg.rounds == exr.rounds) and g.counters()["migrated"] > 0
    res = [None] * world
    dist.all_gather_object(res, (ok, msgs), group=meta)
    if rank == 0:
        torch.save(res, out)
    dist.barrier(group=meta)
    dist.destroy_process_group()


def _worker_3d(rank, world, port, out):
    """Config-5 shape: 3-D box of nitrogen with Larsen-Borgnakke collisions, cut in two by mesh.decompose
    (decomposePar stand-in, processor patches with inherited geometry), NVLink peer-memory transfer on the GPUs,
    gloo slots for the oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    meta = dist.new_group(backend="gloo")
    from oracle.oracle_cloud import OracleCloud
    from unigasfoam_b200 import mesh as ugmesh
    from unigasfoam_b200.cloud import UniGasCloud
    from unigasfoam_b200.exchange import PeerExchanger, SlotExchanger, evolve_distributed
    case = cases.closed_box(n=8, parcels=40000, seed=41, wall="diffuse", binary="LarsenBorgnakkeVariableHardSphere",
                            species=("N2", cases.NITROGEN), dt_mct=0.5, lambda_per_dx=0.6, Trot=250.0,
                            rotationalRelaxationCollisionNumber=5.0, electronicRelaxationCollisionNumber=500.0)
    part = ugmesh.slab_partition(case.mesh, world)
    sub = ugmesh.decompose(case.mesh, part, world)[rank]
    g2l = np.full(case.mesh.n_cells, -1)
    g2l[sub.cell_map] = np.arange(sub.n_cells)
    sel = part[case.cell] == rank
    res = {}
    for name, cls in (("gpu", UniGasCloud), ("oracle", OracleCloud)):
        kw = dict(device=rank) if name == "gpu" else {}
        cl = cls(sub, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=3 * int(sel.sum()) + 1024, rank=rank, nRanks=world, **kw)
        cl.setParcels(case.position[sel], case.U[sel], g2l[case.cell[sel]], None, case.ERot[sel])
        cl.setCellState(sigmaTcRMax=case.sigmaTcRMax)
        if name == "gpu":
            ex = PeerExchanger(cl, sub, rank, world, slot_capacity=4000, group=None, meta_group=meta)
            evolve_distributed(cl, ex, 5, fixed_rounds=3)
            ex.check_settled()
        else:
            ex = SlotExchanger(cl, sub, rank, world, slot_capacity=4000, group=meta, cuda=False)
            evolve_distributed(cl, ex, 5)
        res[name] = (cl.parcels(), cl.counters())
    (pg, cg), (pr, cr) = res["gpu"], res["oracle"]
    same = np.array_equal(pg["cell"], pr["cell"])
    closeU = (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() if same else 0.0
    closeE = (np.abs(pg["ERot"] - pr["ERot"]) <= 1e-9 * np.abs(pr["ERot"]).max()).mean() if same else 0.0
    ok = bool(same and closeU > 0.995 and closeE > 0.995 and cg["migrated"] > 0 and abs(cg["collisions"] - cr["collisions"]) <= 2 and cg["stuck"] == 0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, dict(same=same, closeU=float(closeU), closeE=float(closeE), n=len(pg["cell"]), cg=cg["collisions"], cr=cr["collisions"])), group=meta)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier(group=meta)
    dist.destroy_process_group()


def _worker_cyl(rank, world, port, out, weighted=False):
    """Config-3 shape: Mach-10 cylinder O-grid with free-stream inflow, deleting outflow, diffuse wall and symmetry
    axis, cut along the wake axis by mesh.decompose: insertion on one rank, migration across the cut, outflow on the
    other, every step."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    meta = dist.new_group(backend="gloo")
    from oracle.oracle_cloud import OracleCloud
    from unigasfoam_b200 import mesh as ugmesh
    from unigasfoam_b200.cloud import UniGasCloud
    from unigasfoam_b200.exchange import PeerExchanger, SlotExchanger, evolve_distributed
    # weighted: the reference tutorial's own setting (hypersonicCylinder: cellWeightedSimulation true) - the factor
    # follows the cell volume, parcels are cloned / deleted as they cross cells and carry their factor across the cut
    case = cases.cylinder(nr=16, ntheta=32, ppc=25, cellWeightFactor=("particlesPerSubCell", 25) if weighted else None)
    part = ugmesh.slab_partition(case.mesh, world, axis=1)
    sub = ugmesh.decompose(case.mesh, part, world)[rank]
    g2l = np.full(case.mesh.n_cells, -1)
    g2l[sub.cell_map] = np.arange(sub.n_cells)
    sel = part[case.cell] == rank
    res = {}
    for name, cls in (("gpu", UniGasCloud), ("oracle", OracleCloud)):
        kw = dict(device=rank) if name == "gpu" else {}
        cl = cls(sub, case.uniGasProperties, case.boundariesDict, case.deltaT, parcelCapacity=4 * int(sel.sum()) + 4096, rank=rank, nRanks=world, **kw)
        if weighted:
            cl.setCellState(cellWeightFactor=case.cellWeightFactor[sub.cell_map])
        cl.setParcels(case.position[sel], case.U[sel], g2l[case.cell[sel]])
        cl.setCellState(sigmaTcRMax=case.sigmaTcRMax)
        if name == "gpu":
            ex = PeerExchanger(cl, sub, rank, world, slot_capacity=4000, group=None, meta_group=meta)
            evolve_distributed(cl, ex, 8, inflow=True, fixed_rounds=3)
            ex.check_settled()
        else:
            ex = SlotExchanger(cl, sub, rank, world, slot_capacity=4000, group=meta, cuda=False)
            evolve_distributed(cl, ex, 8, inflow=True)
        res[name] = (cl.parcels(), cl.counters())
    (pg, cg), (pr, cr) = res["gpu"], res["oracle"]
    same = len(pg["cell"]) == len(pr["cell"]) and np.array_equal(pg["cell"], pr["cell"])
    closeU = (np.abs(pg["U"] - pr["U"]) <= 1e-9 * np.abs(pr["U"]).max()).all(1).mean() if same else 0.0
    closeX = (np.abs(pg["position"] - pr["position"]) <= 1e-12 * np.abs(pr["position"]).max()).all(1).mean() if same else 0.0
    tot = torch.tensor([cg["inserted"], cg["deleted"], cg["migrated"]] + ([cg["cloned"], cg["weightDeleted"]] if weighted else []))
    if weighted:
        same = same and cg["cloned"] == cr["cloned"] and cg["weightDeleted"] == cr["weightDeleted"] and np.array_equal(pg["cellWeight"], pr["cellWeight"]) \
            and np.array_equal(pg["cellWeight"], case.cellWeightFactor[sub.cell_map][pg["cell"]])
    dist.all_reduce(tot, group=meta)
    ok = bool(same and closeU > 0.99 and closeX > 0.99 and cg["stuck"] == 0 and cg["inserted"] == cr["inserted"] and cg["deleted"] == cr["deleted"]
              and abs(cg["collisions"] - cr["collisions"]) <= 2 and (tot > 0).all())
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, dict(same=bool(same), closeU=float(closeU), closeX=float(closeX), n=(len(pg["cell"]), len(pr["cell"])),
                                               ins=(cg["inserted"], cr["inserted"]), dele=(cg["deleted"], cr["deleted"]), tot=tot.tolist())), group=meta)
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier(group=meta)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_decomposed_cylinder_with_inflow_matches_oracle(tmp_path, GpuCloud):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "rescyl.pt")
    mp.spawn(_worker_cyl, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert all(ok for ok, _ in res), res


@pytest.mark.timeout(600)
def test_two_gpu_decomposed_weighted_cylinder_matches_oracle(tmp_path, GpuCloud):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "rescylw.pt")
    mp.spawn(_worker_cyl, args=(2, _free_port(), out, True), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert all(ok for ok, _ in res), res


@pytest.mark.timeout(600)
def test_two_gpu_decomposed_3d_nitrogen_lb_matches_oracle(tmp_path, GpuCloud):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "res3d.pt")
    mp.spawn(_worker_3d, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert all(ok for ok, _ in res), res


@pytest.mark.timeout(600)
def test_two_gpu_migration_matches_oracle(tmp_path, GpuCloud):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert all(ok for ok, _ in res), res
