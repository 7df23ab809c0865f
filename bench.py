#!/usr/bin/env python
"""bench.py - particle-steps/s of the uniGasCloud::evolve loop (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (configs[1] of BASELINE.json): 2-D Couette flow, argon, diffuse isothermal
moving walls, Kn 0.1, 1000 x 500 cells, 20 parcels/cell = 10 M parcels per GPU, pure DSMC (NTC + VHS), full
loop move + sort + sample + collide + field accumulation.  N > 1: weak scaling, one 1000 x 500 slab per rank
joined by processor patches, parcels migrated by all_to_all over NCCL.

One JSON line on rank 0.  `value` = parcels processed by all ranks x K / device time of the K steps (CUDA
events on the library's stream, max over ranks).  Inputs (10 M parcels, 520 MB) are far larger than L2 between
steps.  See DESIGN.md §measurement for roofline / e2e / cpu_baseline definitions.
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
B_ALG_CELL = 380.0
# per-kernel algorithmic bytes
MOVE_B_PARCEL, MOVE_B_CELL = 80.0, 4 * 36.0 + 4.0 + 4.0   # 4 face slots per cell on the 2-D bench mesh
CELL_B_PARCEL, CELL_B_CELL = 108.0, 4.0 + 256.0 + 2 * 128.0  # offsets, moment block, accumulator read + write
SORT_B_PARCEL, SORT_B_CELL = 16.0, 16.0


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-state", action="store_true")
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


def run_reference(args, rank, world):
    """The reference's algorithm restated on the CPU (oracle/), all host threads.  The OpenFOAM binary cannot
    be built in this image (no OpenFOAM, no MPI), so this is the reference arm (cpu_baseline.kind = port)."""
    if rank != 0:
        return
    from oracle.oracle_cloud import OracleCloud, num_threads
    case = build_case(args, 0, 1)
    cl = case.make_cloud(OracleCloud, measureWalls=True)
    cl.evolve(args.warmup)
    t0 = time.perf_counter()
    cl.evolve(args.steps)
    dt = time.perf_counter() - t0
    n = cl.size()
    v = n * args.steps / dt
    cores = num_threads()
    print(json.dumps({
        "impl": "reference", "metric": "particle-steps/sec (move+sort+collide+sample)", "value": v, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "parcels": n, "cells": case.mesh.n_cells},
        "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                         "sample": f"full workload, {args.steps} steps after {args.warmup} warm-up, OpenMP over parcels/cells"},
        "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


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
    from unigasfoam_b200.cloud import UniGasCloud
    from unigasfoam_b200.exchange import Exchanger, PeerExchanger, SlotExchanger, evolve_distributed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libugf has no CPU fallback")
    torch.cuda.set_device(local)
    meta = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        meta = dist.new_group(backend="gloo")

    t_case = time.perf_counter()
    case = build_case(args, rank, world)
    cloud = case.make_cloud(UniGasCloud, device=local, measureWalls=True, seed=20261017, parcelCapacity=int(1.5 * case.n_parcels) + 4096)
    nC = case.mesh.n_cells
    t_case = time.perf_counter() - t_case
    stream = torch.cuda.ExternalStream(cloud.stream())
    ex = None
    if world > 1:
        # slot capacity from one exact-count round: 4x the busiest processor patch of a warm-up step
        probe = Exchanger(cloud, case.mesh, rank, world, data_group=None, meta_group=meta, cuda=True)
        cloud.move()
        worst = torch.tensor([int(cloud.migrateCounts().max())], device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        while probe.exchange() > 0:
            pass
        cloud.buildCellOccupancy(); cloud.collide(); cloud.relax(); cloud.accumulateFields(); cloud.endStep()
        slot_cap = max(4096, 4 * int(worst.item()))
        if os.environ.get("UGF_EXCHANGE", "peer") == "nccl":
            ex = SlotExchanger(cloud, case.mesh, rank, world, slot_capacity=slot_cap, group=None, cuda=True)
        else:
            ex = PeerExchanger(cloud, case.mesh, rank, world, slot_capacity=slot_cap, group=None, meta_group=meta)
        # rounds this decomposition needs per step, measured with the exact termination rule over a few steps
        need = 1
        for _ in range(3):
            cloud.move()
            ex.begin_step()
            r = 1
            while ex.exchange() > 0:
                r += 1
            need = max(need, r)
            cloud.finishStep()
        fixed_rounds = need

    def step(n):
        if world == 1:
            cloud.evolve(n)
        else:
            evolve_distributed(cloud, ex, n, fixed_rounds=fixed_rounds)

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms before its first line: start it ahead of the warm-up
    step(max(args.warmup, 3))
    barrier()
    if rank == 0:
        sampler.lines.clear()  # keep only what is sampled from here on
    l0 = cloud.launchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    step(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if ex is not None:
        ex.check_settled()  # the lagged quiescence check of the last timed step
    launches = cloud.launchCount() - l0
    # The timed region is tens of milliseconds, the sampler ticks every 100 ms: keep the identical load running
    # (untimed, all ranks) until a few samples exist, so the clocks line always describes this workload under load.
    t_load = time.perf_counter()
    while True:
        enough = torch.tensor([1 if (rank != 0 or len(sampler.lines) >= 3 or time.perf_counter() - t_load > 3.0) else 0], device="cuda")
        if world > 1:
            dist.broadcast(enough, 0)
        if int(enough.item()):
            break
        step(10)
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region plus the same step loop continued until >= 3 samples (nvidia-smi -lms 100)"
    n_parcels = cloud.size()
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_parcels)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(tms.item())
    total_parcels = float(tot.item())
    value = total_parcels * args.steps / (ms * 1e-3)

    # ---- per-kernel device times over a second pass of K steps (events inside the library, per phase) ----
    phases = {k: 0.0 for k in ("inflow", "move", "sort", "cell", "collide", "relax", "fields")}
    if world == 1:
        for _ in range(args.steps):
            cloud.evolve(1)
            for k, v in cloud.phaseTimes().items():
                phases[k] += v
        for k in phases:
            phases[k] /= args.steps
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
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")))
    except Exception:
        pass
    default_workload = args.case == "couette" and (args.nx, args.ny, args.ppc) == (1000, 500, 20)
    if world > 1:
        step_bytes = B_ALG_PARCEL * n_parcels + B_ALG_CELL * nC  # per GPU
        g = step_bytes / (ms / args.steps * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "whole step, per GPU (per-kernel split is reported at N=1)", "achieved": g, "peak": peak, "unit": "GB/s",
                    "frac": g / peak, "traffic": None, "peak_source": peak_src}
    if world == 1 and phases["cell"] > 0:
        kb = {
            "move_kernel": (MOVE_B_PARCEL * n_parcels + MOVE_B_CELL * nC, phases["move"]),
            "cell_kernel": (CELL_B_PARCEL * n_parcels + CELL_B_CELL * nC, phases["cell"]),
            "sort(scan+scatter+segment)": (SORT_B_PARCEL * n_parcels + SORT_B_CELL * nC, phases["sort"]),
        }
        dom = max(kb, key=lambda k: kb[k][1])
        ach = kb[dom][0] / (kb[dom][1] * 1e-3) / 1e9
        step_bytes = B_ALG_PARCEL * n_parcels + B_ALG_CELL * nC
        tkey = {"move_kernel": "move_stream_kernel", "cell_kernel": "cell_kernel"}.get(dom)
        roofline = {
            "bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic.get(tkey) if default_workload else None, "traffic_source": traffic.get("source") if default_workload else None,
            "peak_source": peak_src,
            "per_kernel": {k: {"ms": v[1], "alg_bytes": v[0], "GBps": v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else None} for k, v in kb.items()},
            "phase_ms": phases,
            "step": {"alg_bytes": step_bytes, "GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                     "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
        }

    # ---- e2e: the call a user makes - evolve() + info() per step through the host API ---------------------
    # State stays resident (it is the simulation state, like model weights); per step the host sends the
    # step's control block (kernel parameter blocks incl. deltaT) and reads back the step's log quantities.
    e2e = None
    host_state = None
    if world > 1:
        barrier()
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            cloud.setDeltaT(case.deltaT)
            evolve_distributed(cloud, ex, 1, fixed_rounds=fixed_rounds)
            done += cloud.counters()["nParcels"]
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, float(done)], dtype=torch.float64, device="cuda")
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        e2e = {"value": float(tt[1].item()) / float(tmax[0].item()), "unit": "particle-steps/s",
               "h2d_bytes_per_step": int(launches / args.steps * 1400), "d2h_bytes_per_step": 64 + 48 + 8 + 8,
               "what": "per rank and step: move + transfer rounds + finishStep + counters() through the host API, state resident in HBM; max over ranks"}
    if world == 1:
        barrier()
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            cloud.setDeltaT(case.deltaT)
            cloud.evolve(1)
            c = cloud.counters()  # D2H: counters + energy/momentum totals, synchronises
            done += c["nParcels"]
        dt = time.perf_counter() - t0
        per_step_launches = launches / args.steps
        e2e = {"value": done / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": int(per_step_launches * 1400),
               "d2h_bytes_per_step": 64 + 48 + 8, "what": "UniGasCloud.evolve(1)+counters() per step, state resident in HBM"}
        if not args.no_host_state:
            # plugin-level integration (parcels owned by the host solver): upload + step + download every step
            P = cloud.parcels()
            t0 = time.perf_counter()
            reps = max(2, min(args.steps, 4))
            for _ in range(reps):
                cloud.setParcels(P["position"], P["U"], P["cell"])
                cloud.evolve(1)
                P = cloud.parcels()
            dt = time.perf_counter() - t0
            host_state = {"value": len(P["cell"]) * reps / dt, "unit": "particle-steps/s",
                          "h2d_bytes_per_step": 52 * len(P["cell"]), "d2h_bytes_per_step": 64 * len(P["cell"]),
                          "what": "pageable host SoA uploaded, one step, parcels downloaded - every step"}

    # ---- cpu baseline: the oracle on this host's cores, bounded sample --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # N = 1 only: under torchrun the ranks share the host cores
        from oracle.oracle_cloud import OracleCloud, num_threads
        ocase = case if world == 1 else build_case(args, 0, 1)
        oc = ocase.make_cloud(OracleCloud)
        oc.evolve(1)
        t0 = time.perf_counter()
        oc.evolve(args.cpu_steps)
        dt = time.perf_counter() - t0
        cpu = {"value": oc.size() * args.cpu_steps / dt, "unit": "particle-steps/s", "cores": num_threads(), "kind": "port",
               "sample": f"same workload ({oc.size()} parcels), {args.cpu_steps} steps after 1 warm-up; CPU restatement of the reference loop "
                         "(oracle/), OpenMP over parcels/cells - not the OpenFOAM binary"}
        oc.close()

    if rank == 0:
        out = {
            "metric": "particle-steps/sec (move+sort+collide+sample)", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "parcels_per_gpu": n_parcels, "cells_per_gpu": nC,
                       "l2": "inputs (52 B x 10 M parcels per buffer) exceed the 126 MB L2 between steps",
                       "parallelism": f"domain-decomposition x{world}" if world > 1 else "single subdomain"},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "e2e_host_state": host_state, "case_build_s": t_case,
        }
        if world > 1:
            out["migration"] = {"rounds": ex.rounds, "slot_capacity": ex.cap, "rounds_per_step": fixed_rounds,
                                "transport": "NVLink peer memory (pack kernel writes the neighbour's receive slot, device-side flag wait)" if isinstance(ex, PeerExchanger)
                                else "NCCL grouped send/recv between neighbours",
                                "protocol": "fixed slots per processor patch, rounds per step measured with the exact termination rule during warm-up, quiescence verified by a lagged all-reduce"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
