#!/usr/bin/env python
"""bench.py - particle-steps/s of the uniGasCloud::evolve loop (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Headline workload (configs[1] of BASELINE.json): 2-D Couette flow, argon, diffuse isothermal
moving walls, Kn 0.1, 1000 x 500 cells, 20 parcels/cell = 10 M parcels per GPU, pure DSMC (NTC + VHS), full
loop move + sort + sample + collide + field accumulation.  N > 1: weak scaling, one 1000 x 500 slab per rank
joined by processor patches; migrating parcels are packed by a kernel straight into the neighbour's HBM over NVLink
peer memory (PeerExchanger; UGF_EXCHANGE=nccl selects grouped NCCL send/recv instead).

One JSON line on rank 0.  `value` = parcels processed by all ranks x K / device time of the K steps (CUDA
events on the library's stream, max over ranks).  Inputs (10 M parcels, 520 MB) are far larger than L2 between
steps.  `other_configs` carries BASELINE configs[2], [3] and [4] at their stated size (50 M-parcel cylinder, 100 M-parcel
hybrid USP-SBGK / NTC, nitrogen Larsen-Borgnakke blunt body at 62.5 M parcels per GPU), decomposed over the N ranks,
each with value, ms/step, step-level roofline fraction and collisions per step.  `--impl reference` runs the CPU
restatement of the same N-subdomain run (same partition, same parcels) on all host cores.
See DESIGN.md §measurement for roofline / e2e / cpu_baseline definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_PARCEL = 236.0  # algorithmic bytes per parcel-step, argon (SURVEY §8d / BASELINE.md §3)
B_ALG_PARCEL_ROT = 276.0  # with ERot (nitrogen, Larsen-Borgnakke)
B_ALG_CELL = 380.0
# per-kernel algorithmic bytes
MOVE_B_PARCEL, MOVE_B_CELL = 80.0, 4 * 36.0 + 4.0 + 4.0   # 4 face slots per cell on the 2-D bench mesh
CELL_B_PARCEL, CELL_B_CELL = 108.0, 4.0 + 2 * 128.0  # offsets, accumulator read + write (moment blocks are not stored in pure-DSMC steps)
SORT_B_PARCEL, SORT_B_CELL = 16.0, 16.0
METRIC = "particle-steps/sec (move+sort+collide+sample)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ugf", choices=["ugf", "reference"])
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--ppc", type=int, default=20)
    ap.add_argument("--case", default="couette", choices=["couette", "box", "cylinder"], help="box = config 1 style dense-collision case (tuning only)")
    ap.add_argument("--collision", default="dsmc", choices=["dsmc", "bgk", "hybrid"], help="box case only (tuning)")
    ap.add_argument("--gas", default="argon", choices=["argon", "n2lb"], help="box case: n2lb = nitrogen with Larsen-Borgnakke (config 5 per-GPU shape)")
    ap.add_argument("--box-n", type=int, default=64)
    ap.add_argument("--box-parcels", type=int, default=8_000_000)
    ap.add_argument("--cpu-steps", type=int, default=2, help="steps of the bounded cpu_baseline sample")
    ap.add_argument("--weighted", action="store_true", help="cylinder case: cell-weighted simulation (tuning / overhead measurement)")
    ap.add_argument("--settle", type=int, default=40, help="extra untimed steps before the W warm-up steps so that (sigma_T c_r)max and with it the collision rate are at their steady value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-state", action="store_true")
    ap.add_argument("--other", default="all", help="comma list of config3,config4,config5 | all | none: the other BASELINE configs run after the headline one")
    ap.add_argument("--other-scale", type=float, default=1.0, help="linear mesh scale of the other configs (1 = BASELINE size; parcels scale with the cell count)")
    ap.add_argument("--other-timeout", type=float, default=420.0, help="seconds after which rank 0 prints the line without waiting for the other configs")
    ap.add_argument("--ref-steps", type=int, default=0, help="reference arm: steps per timed sample (0 = --steps)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_case(args, rank, world):
    from unigasfoam_b200 import cases
    if args.case == "box":
        kw = {}
        if args.collision != "dsmc":
            kw = dict(mode=args.collision, bgk="unifiedStochasticParticleSBGK", theta=0.1)
        if args.gas == "n2lb":
            kw.update(binary="LarsenBorgnakkeVariableHardSphere", species=("N2", cases.NITROGEN), Trot=300.0,
                      rotationalRelaxationCollisionNumber=5.0, electronicRelaxationCollisionNumber=500.0)
        c = cases.closed_box(n=args.box_n, parcels=args.box_parcels, wall="diffuse", **kw)
        if args.collision == "hybrid":
            import numpy as np
            c.cellCollModelId = (np.arange(c.mesh.n_cells) % 2).astype(np.int32)
        return c
    if args.case == "cylinder":
        # --weighted: the reference tutorial's setting (cellWeightedSimulation, factor from uniGasMeshFill's rule)
        return cases.cylinder(nr=500, ntheta=1000, ppc=20, cellWeightFactor=("particlesPerSubCell", 20) if args.weighted else None)
    return cases.couette(nx=args.nx, ny=args.ny, ppc=args.ppc, rank=rank, n_ranks=world)


def workload_name(args):
    if args.case == "cylinder":
        return "cylinder2d_mach10_argon_500x1000cells_20ppc_inflow_outflow_dsmc_ntc_vhs" + ("_cellweighted" if args.weighted else "")
    if args.case == "box":
        return f"closedbox3d_{args.gas}_{args.box_n}^3cells_{args.box_parcels}parcels_{args.collision}_dt0.2mct"
    return f"couette2d_argon_kn0.1_{args.nx}x{args.ny}cells_{args.ppc}ppc_dsmc_ntc_vhs"


def config_dict(args, world, parcels_per_gpu, cells_per_gpu):
    """The `config` object of the JSON line - identical in the ugf and the reference arm for the same command line."""
    return {"workload": workload_name(args), "parcels_per_gpu": int(parcels_per_gpu), "cells_per_gpu": int(cells_per_gpu),
            "l2": "inputs (52 B x 10 M parcels per buffer) exceed the 126 MB L2 between steps",
            "parallelism": f"domain-decomposition x{world}" if world > 1 else "single subdomain"}


def expected_collisions_per_step(case, n_parcels):
    """1/2 N nu dt with the equilibrium VHS collision rate (Bird 4.64, the expression the reference codes at
    uniGasVolFields.C:1150-1151) at the case's initial state; None where the case has no uniform equilibrium state."""
    from unigasfoam_b200 import cases
    m = case.meta
    if "n" not in m or "species" not in m or not isinstance(m["species"], dict):
        return None
    T = m.get("Tw", m.get("T0", m.get("T_inf")))
    if T is None:
        return None
    nu = cases.vhs_collision_rate(m["n"], T, m["species"], m.get("Tref", 273.0))
    return 0.5 * n_parcels * nu * case.deltaT


# ======================================================================================================================
# reference arm: the CPU restatement of the same N-subdomain run on the host's cores
def run_reference(args, rank, world):
    """The reference's algorithm restated on the CPU (oracle/), all host cores.  The OpenFOAM binary cannot be built in this
    image (no OpenFOAM, no MPI), so this is the reference arm (cpu_baseline.kind = port).  With N ranks it runs the same N
    subdomains the GPU arm runs - same partition, same parcels, the Cloud::move transfer loop between them - inside rank 0's
    process: the subdomains take turns on all the host's cores (a process per rank would only split the same cores N ways).
    The other ranks exit without work."""
    if rank != 0:
        return
    from oracle.oracle_cloud import OracleCloud, set_num_threads
    from unigasfoam_b200.exchange import LocalSubdomains
    cores = set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: override it
    steps = args.ref_steps or args.steps
    t_build = time.perf_counter()
    cs = [build_case(args, r, world) for r in range(world)]
    clouds = [c.make_cloud(OracleCloud, measureWalls=True, rank=r, nRanks=world) for r, c in enumerate(cs)]
    t_build = time.perf_counter() - t_build
    n0 = sum(c.size() for c in clouds)
    if world == 1:
        cl = clouds[0]
        step = cl.evolve
    else:
        sub = LocalSubdomains(clouds, [c.mesh for c in cs])
        step = sub.evolve
    step(max(args.warmup, 1))
    t0 = time.perf_counter()
    step(steps)
    dt = time.perf_counter() - t0
    n = sum(c.size() for c in clouds)
    coll = sum(c.counters()["collisions"] for c in clouds)
    v = n * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world, n / world, cs[0].mesh.n_cells),
        "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                         "sample": f"full workload: {world} subdomain(s), {n} parcels, {steps} steps after {max(args.warmup, 1)} warm-up; "
                                   "OpenMP over parcels / cells on all host cores, subdomains in turn, transfer loop between them"},
        "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "checks": {"parcels_before": int(n0), "parcels_after": int(n), "collisions_last_step": int(coll),
                   "transfer_rounds": (sub.rounds if world > 1 else 0)},
        "case_build_s": t_build,
    }), flush=True)


# ======================================================================================================================
class Runner:
    """One decomposed (or single) case on this rank's GPU: cloud, exchanger, stepping, device timing."""

    def __init__(self, case, args, rank, world, local, meta, inflow, capacity_factor=1.5, exact_rounds=False):
        import torch
        import torch.distributed as dist
        from unigasfoam_b200.cloud import UniGasCloud
        from unigasfoam_b200.exchange import Exchanger, PeerExchanger, SlotExchanger
        self.torch, self.dist = torch, dist
        self.case, self.rank, self.world, self.inflow = case, rank, world, inflow
        self.cloud = case.make_cloud(UniGasCloud, device=local, measureWalls=True, seed=20261017, rank=rank, nRanks=world,
                                     parcelCapacity=int(capacity_factor * case.n_parcels) + 65536)
        cloud = self.cloud
        self.stream = torch.cuda.ExternalStream(cloud.stream())
        self.ex = None
        self.fixed_rounds = None
        if world > 1:
            # slot capacity from one exact-count round: 4x the busiest processor patch of a warm-up step
            probe = Exchanger(cloud, case.mesh, rank, world, data_group=None, meta_group=meta, cuda=True)
            if inflow:
                cloud.controlBeforeMove()
            cloud.move()
            worst = torch.tensor([int(cloud.migrateCounts().max())], device="cuda")
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            while probe.exchange() > 0:
                pass
            cloud.finishStep()
            slot_cap = max(4096, 4 * int(worst.item()))
            if os.environ.get("UGF_EXCHANGE", "peer") == "nccl":
                self.ex = SlotExchanger(cloud, case.mesh, rank, world, slot_capacity=slot_cap, group=None, cuda=True)
            else:
                self.ex = PeerExchanger(cloud, case.mesh, rank, world, slot_capacity=slot_cap, group=None, meta_group=meta)
            # rounds this decomposition needs per step, measured with the exact termination rule over a few steps
            need = 1
            for _ in range(5):
                if inflow:
                    cloud.controlBeforeMove()
                cloud.move()
                self.ex.begin_step()
                r = 1
                while self.ex.exchange() > 0:
                    r += 1
                need = max(need, r)
                cloud.finishStep()
            # a block decomposition with corners (more than two processor patches per rank) now and then needs one round more than
            # three probe steps show (a parcel crossing an edge of the block late in its track): one round of margin there
            n_proc = sum(1 for p in case.mesh.patches if p.kind == "processor")
            self.fixed_rounds = need + (1 if n_proc > 2 else 0)
            self.probe_rounds = need
            if exact_rounds:
                self.fixed_rounds = None

    def step(self, n):
        from unigasfoam_b200.exchange import evolve_distributed
        if self.world == 1:
            self.cloud.evolve(n)
        else:
            evolve_distributed(self.cloud, self.ex, n, inflow=self.inflow, fixed_rounds=self.fixed_rounds, max_rounds=512)

    def barrier(self):
        self.stream.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, vals, op="sum"):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def timed(self, steps):
        """steps steps bracketed by barrier + synchronize; -> ms (max over ranks), launches."""
        torch = self.torch
        l0 = self.cloud.launchCount()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        self.step(steps)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.ex is not None:
            self.ex.check_settled()  # the lagged quiescence check of the last timed step
        return self.reduce([ms], "max")[0], self.cloud.launchCount() - l0

    def totals(self):
        """Global sums of the last step's counters and of the cloud's size / energy."""
        c = self.cloud.counters()
        keys = ("nParcels", "collisionCandidates", "collisions", "bgkRelaxations", "inserted", "deleted", "cloned", "weightDeleted",
                "migrated", "wallHits", "stuck", "linearKineticEnergy", "rotationalEnergy")
        return dict(zip(keys, self.reduce([c[k] for k in keys])))

    def close(self):
        self.ex = None
        self.cloud.close()
        self.torch.cuda.empty_cache()


def phase_profile(cloud, steps):
    phases = {k: 0.0 for k in ("inflow", "move", "sort", "cell", "collide", "relax", "fields")}
    for _ in range(steps):
        cloud.evolve(1)
        for k, v in cloud.phaseTimes().items():
            phases[k] += v
    return {k: v / steps for k, v in phases.items()}


def other_config(name, args, rank, world, local, meta, peak):
    """One of BASELINE configs[2..4] at its stated size (x other_scale), decomposed over the ranks of this run."""
    from unigasfoam_b200 import cases
    s = args.other_scale
    t0 = time.perf_counter()
    if name == "config3":
        if world > 4:
            return {"skipped": "BASELINE names 2 / 4 GPUs for this config"}
        case = cases.cylinder_block(rank, world, nr=max(8, int(1000 * s)), ntheta=max(8 * world, int(2500 * s)), parcels=int(50e6 * s * s))
        label = "configs[2]: 2-D Mach-10 argon cylinder, inflow / outflow, DSMC NTC + VHS, 50 M parcels, 2.5 M cells"
        bpp = B_ALG_PARCEL
    elif name == "config4":
        case = cases.cylinder_block(rank, world, nr=max(8, int(1000 * s)), ntheta=max(8 * world, int(2500 * s)), parcels=int(100e6 * s * s), hybrid=True)
        label = ("configs[3]: hybrid USP-SBGK (upstream half) / NTC + VHS (wake half) on the cylinder topology at 10 n_inf, 100 M parcels, "
                 "2.5 M cells, macroInterpolation false, frozen mask")
        bpp = B_ALG_PARCEL
    else:
        case = cases.blunt_body_block(rank, world, n_eta=max(4, int(200 * s)), n_s=max(8, int(500 * s)), n_phi=max(4, int(250 * s)))
        label = ("configs[4]: 3-D nitrogen Mach-10 blunted cone, Larsen-Borgnakke, cell-weighted, 62.5 M parcels and 3.1 M cells per GPU "
                 f"({world} of the 8 blocks of the 500 M-parcel case: 4 along the body x 2 in azimuth)")
        bpp = B_ALG_PARCEL_ROT
    t_case = time.perf_counter() - t0
    # These runs use the reference's own termination rule for the transfer loop - rounds until no rank has a parcel in flight, one
    # small all-reduce per round - instead of the headline's fixed round count: with 500 M parcels a few dozen parcels per step need
    # five or more transfers (next to the axis of the 90-degree sector a parcel is mirrored between the two symmetry planes and
    # crosses the azimuthal cut between the blocks every time, the closer to the axis the more often within one step), and a
    # step of ~10 ms hides the synchronisation.
    run = Runner(case, args, rank, world, local, meta, inflow=True, capacity_factor=1.6, exact_rounds=True)
    try:
        nC = case.mesh.n_cells
        before = run.totals()
        run.step(max(args.warmup, 3))
        ms, launches = run.timed(args.steps)
        after = run.totals()
        n_tot = after["nParcels"]
        cells_tot = run.reduce([nC])[0]
        value = 0.5 * (before["nParcels"] + n_tot) * args.steps / (ms * 1e-3)  # parcels grow / shrink through the patches: mean of the ends
        step_bytes = (bpp * n_tot + B_ALG_CELL * cells_tot) / world
        gbps = step_bytes / (ms / args.steps * 1e-3) / 1e9
        out = {"what": label, "n_gpus": world, "value": value, "unit": "particle-steps/s", "ms_per_step": ms / args.steps, "steps": args.steps,
               "parcels": int(n_tot), "cells": int(cells_tot),
               "roofline_step": {"alg_bytes_per_gpu": step_bytes, "GBps_per_gpu": gbps, "frac": gbps / peak},
               "per_step": {k: int(after[k]) for k in ("collisionCandidates", "collisions", "bgkRelaxations", "inserted", "deleted", "cloned",
                                                       "weightDeleted", "migrated", "wallHits", "stuck")},
               "rounds_per_step": getattr(run, "probe_rounds", None), "transfer_loop": "exact termination rule (all-reduce per round)" if world > 1 else None,
               "gpu_launches": int(launches), "case_build_s": t_case, "deltaT": case.deltaT}
        if world == 1:
            out["phase_ms"] = phase_profile(run.cloud, min(args.steps, 5))
        return out
    finally:
        run.close()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libugf has no CPU fallback")
    torch.cuda.set_device(local)
    meta = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        meta = dist.new_group(backend="gloo")

    t_case = time.perf_counter()
    case = build_case(args, rank, world)
    run = Runner(case, args, rank, world, local, meta, inflow=(args.case == "cylinder"))
    cloud, ex = run.cloud, run.ex
    nC = case.mesh.n_cells
    t_case = time.perf_counter() - t_case

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms before its first line: start it ahead of the warm-up
    start = run.totals()
    warm = max(args.warmup, 3)
    run.step(args.settle + warm)  # settle: (sigma_T c_r)max reaches its steady value, then the W warm-up steps
    run.barrier()
    if rank == 0:
        sampler.lines.clear()  # keep only what is sampled from here on
    before = run.totals()
    ms, launches = run.timed(args.steps)
    after = run.totals()
    # The timed region is tens of milliseconds, the sampler ticks every 100 ms: keep the identical load running
    # (untimed, all ranks) until a few samples exist, so the clocks line always describes this workload under load.
    t_load = time.perf_counter()
    while True:
        enough = torch.tensor([1 if (rank != 0 or len(sampler.lines) >= 3 or time.perf_counter() - t_load > 3.0) else 0], device="cuda")
        if world > 1:
            dist.broadcast(enough, 0)
        if int(enough.item()):
            break
        run.step(10)
        run.barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region plus the same step loop continued until >= 3 samples (nvidia-smi -lms 100)"
    n_parcels = cloud.size()
    total_parcels = after["nParcels"]
    value = total_parcels * args.steps / (ms * 1e-3)
    expect = expected_collisions_per_step(case, total_parcels)
    checks = {
        "parcels_at_start": int(start["nParcels"]), "parcels_before_timed": int(before["nParcels"]), "parcels_after_timed": int(after["nParcels"]),
        "parcel_balance_ok": bool(args.case == "cylinder" or before["nParcels"] == after["nParcels"] == start["nParcels"]),
        "stuck": int(after["stuck"]),
        "kinetic_energy_rel_change_timed": (after["linearKineticEnergy"] - before["linearKineticEnergy"]) / before["linearKineticEnergy"],
        "collisions_per_step": int(after["collisions"]), "candidates_per_step": int(after["collisionCandidates"]),
        "expected_collisions_per_step_equilibrium": expect,
        "collisions_vs_expected": (after["collisions"] / expect) if expect else None,
        "wall_hits_per_step": int(after["wallHits"]), "migrated_per_step": int(after["migrated"]),
        "note": "closed system (cyclic / processor patches + walls): the global parcel count must not change; diffuse moving walls do work on the gas, so the energy drifts slowly upwards",
    }

    # ---- per-kernel device times over a second pass of K steps (events inside the library, per phase) ----
    phases = {k: 0.0 for k in ("inflow", "move", "sort", "cell", "collide", "relax", "fields")}
    if world == 1:
        phases = phase_profile(cloud, args.steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    roofline = None
    traffic = {}
    try:  # DRAM bytes per launch from the committed `ncu --set full` capture of this workload (profiles/)
        for name in ("ncu_traffic_r2.json", "ncu_traffic_r1.json"):
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                traffic = json.load(open(path))
                break
    except Exception:
        pass
    default_workload = args.case == "couette" and (args.nx, args.ny, args.ppc) == (1000, 500, 20)
    if world > 1:
        step_bytes = B_ALG_PARCEL * n_parcels + B_ALG_CELL * nC  # per GPU
        gb = step_bytes / (ms / args.steps * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "whole step, per GPU (per-kernel split is reported at N=1)", "achieved": gb, "peak": peak, "unit": "GB/s",
                    "frac": gb / peak, "traffic": None, "peak_source": peak_src}
    if world == 1 and phases["cell"] > 0:
        kb = {
            "move_kernel": (MOVE_B_PARCEL * n_parcels + MOVE_B_CELL * nC, phases["move"]),
            "cell_kernel": (CELL_B_PARCEL * n_parcels + CELL_B_CELL * nC, phases["cell"]),
            "sort(scan+scatter+segment)": (SORT_B_PARCEL * n_parcels + SORT_B_CELL * nC, phases["sort"]),
        }
        dom = max(kb, key=lambda k: kb[k][1])
        ach = kb[dom][0] / (kb[dom][1] * 1e-3) / 1e9
        step_bytes = B_ALG_PARCEL * n_parcels + B_ALG_CELL * nC
        tkeys = {"move_kernel": ("move_stream2_kernel", "move_stream_kernel"), "cell_kernel": ("cell_kernel",)}.get(dom, ())
        tval = next((traffic[k] for k in tkeys if k in traffic), None)
        roofline = {
            "bound": "hbm", "kernel": {"move_kernel": "move_stream2_kernel (phase: move)"}.get(dom, dom), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": tval if default_workload else None, "traffic_source": traffic.get("source") if default_workload else None,
            "peak_source": peak_src,
            "per_kernel": {k: {"ms": v[1], "alg_bytes": v[0], "GBps": v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else None} for k, v in kb.items()},
            "phase_ms": phases,
            "step": {"alg_bytes": step_bytes, "GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                     "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
        }

    # ---- e2e: the call a user makes - evolve() + info() per step through the host API ---------------------
    # State stays resident (it is the simulation state, like model weights); per step the host sends the
    # step's control block (kernel argument blocks incl. deltaT) and reads back the step's log quantities.
    # Bytes are measured by the library (ugf_transfer_bytes), not estimated.
    e2e = None
    host_state = None
    tb0 = cloud.transferBytes()
    run.barrier()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        cloud.setDeltaT(case.deltaT)
        run.step(1)
        done += cloud.counters()["nParcels"]  # D2H: counters + energy / momentum totals, synchronises
    run.barrier()
    dt = time.perf_counter() - t0
    tb1 = cloud.transferBytes()
    tsum = run.reduce([done])
    tmax = run.reduce([dt], "max")
    e2e = {"value": tsum[0] / tmax[0], "unit": "particle-steps/s",
           "h2d_bytes_per_step": int(((tb1[0] - tb0[0]) + (tb1[2] - tb0[2])) / args.steps), "d2h_bytes_per_step": int((tb1[1] - tb0[1]) / args.steps),
           "bytes": "measured by libugf per rank (ugf_transfer_bytes): explicit copies + kernel-argument blocks",
           "what": ("UniGasCloud.evolve(1)+counters() per step" if world == 1 else
                    "per rank and step: move + transfer rounds + finishStep + counters() through the host API; max over ranks") + ", state resident in HBM"}
    if world == 1 and not args.no_host_state:
        # plugin-level integration (parcels owned by the host solver, as uniGasCloud's IDLList is): upload + step + download
        # every step, through page-locked SoA columns handed out by the library (ugf_host_alloc)
        import numpy as np
        cap = int(cloud.cfg.parcelCapacity)
        cols = [cloud.hostArray(cap) for _ in range(6)]
        cellc = cloud.hostArray(cap, np.int32)
        n = cloud.parcelsInto(*cols, cellc)
        tb0 = cloud.transferBytes()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            cloud.setParcelsSoA(n, *cols, cellc)
            cloud.evolve(1)
            n = cloud.parcelsInto(*cols, cellc)
        dt = time.perf_counter() - t0
        tb1 = cloud.transferBytes()
        host_state = {"value": n * reps / dt, "unit": "particle-steps/s",
                      "h2d_bytes_per_step": int(((tb1[0] - tb0[0]) + (tb1[2] - tb0[2])) / reps), "d2h_bytes_per_step": int((tb1[1] - tb0[1]) / reps),
                      "what": "every step: page-locked host SoA (ugf_host_alloc) uploaded, one step, parcels downloaded into it"}

    # ---- cpu baseline: the oracle on this host's cores, bounded sample --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # N = 1 only: under torchrun the ranks share the host cores
        from oracle.oracle_cloud import OracleCloud, set_num_threads
        cores = set_num_threads(os.cpu_count() or 1)
        oc = case.make_cloud(OracleCloud)
        oc.evolve(1)
        t0 = time.perf_counter()
        oc.evolve(args.cpu_steps)
        dt = time.perf_counter() - t0
        cpu = {"value": oc.size() * args.cpu_steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
               "sample": f"same workload ({oc.size()} parcels), {args.cpu_steps} steps after 1 warm-up; CPU restatement of the reference loop "
                         "(oracle/), OpenMP over parcels/cells - not the OpenFOAM binary"}
        oc.close()

    rounds = (ex.rounds, ex.cap, run.fixed_rounds, type(ex).__name__) if ex is not None else None
    run.close()

    # ---- the headline line is complete here; make sure it gets out whatever happens to the other configs ----------------
    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "settle_steps": args.settle, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world, n_parcels, nC),
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "e2e_host_state": host_state, "checks": checks, "other_configs": {}, "case_build_s": t_case,
        }
        if rounds is not None:
            peer = rounds[3] == "PeerExchanger"
            out["migration"] = {"rounds": rounds[0], "slot_capacity": rounds[1], "rounds_per_step": rounds[2],
                                "transport": "NVLink peer memory (pack kernel writes the neighbour's receive slot, device-side flag wait)" if peer
                                else "NCCL grouped send/recv between neighbours",
                                "protocol": "fixed slots per processor patch, rounds per step measured with the exact termination rule during warm-up, quiescence verified by a lagged all-reduce"}
    emitted = threading.Lock()

    def emit(note=None):
        if out is None or not emitted.acquire(blocking=False):
            return
        if note:
            out["other_configs_note"] = note
        print(json.dumps(out), flush=True)

    if rank == 0:
        import signal

        def on_term(signum, frame):  # another rank died inside one of the other configs: torchrun is taking the job down
            emit("terminated while running the other configs (another rank failed); headline numbers are complete")
            os._exit(1)

        signal.signal(signal.SIGTERM, on_term)

        def watchdog():
            emit(f"the other configs did not finish within {args.other_timeout} s; headline numbers are complete")
            os._exit(0)

        wd = threading.Timer(args.other_timeout, watchdog)
        wd.daemon = True
        wd.start()

    # ---- the other BASELINE configs at their stated size -------------------------------------------------------
    others = {}
    want = [] if args.other == "none" else (["config3", "config4", "config5"] if args.other == "all" else args.other.split(","))
    if args.case != "couette":
        want = []
    for name in want:
        try:
            others[name] = other_config(name, args, rank, world, local, meta, peak)
        except Exception as e:  # the headline line must survive a failure here; the failure is reported, not hidden
            others[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            if world > 1:  # the ranks are no longer in step: finish here (rank 0's line goes out through emit)
                if rank == 0:
                    out["other_configs"] = others
                    emit("a rank failed in " + name)
                raise

    if rank == 0:
        out["other_configs"] = others
        emit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
