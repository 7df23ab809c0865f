"""The time loop of the uniGasFoam solver around the cloud (applications/solvers/discreteMethods/uniGasFoam/uniGasFoam.C:
74-111), host side: `while (runTime.loop()) { uniGas.evolve(); uniGas.info(); runTime.write(); }` with the write times,
averaging resets and restart behaviour the case's own system/controlDict and system/fieldPropertiesDict ask for.

Not part of the hot path (the solver executable stays OpenFOAM's); it exists so that a case directory written
for the reference can be run and leaves the files the reference would leave:

    <time>/lagrangian/uniGas/*, <time>/uniGas{SigmaTcRMax, CellWeightFactor, SubCellLevels, CollisionModelId},
    <time>/uniform/time, <time>/uniform/volFieldsMethod_<field>, <time>/{rhoN,p,UMean,...}_<field>

controlDict keys honoured: startFrom (latestTime / startTime), startTime, endTime, deltaT, writeControl (timeStep /
runTime / adjustableRunTime), writeInterval, timePrecision, nTerminalOutputs (every that many steps the `Time =` line,
uniGasCloud::info and the clock line go to `log`, uniGasFoam.C:76-104).  fieldPropertiesDict: per uniGasVolFields entry `field`,
timeProperties.{sampleInterval, resetAtOutput, resetAtOutputUntilTime}, averagingAcrossManyRuns,
measureMeanFreePath, measureErrors.
"""
import os
import shutil
import time

from . import _capi, cases, foamdict
from .adapter import UniGasDynamicAdapter


def time_name(t, precision=6):
    """Time::timeName with `timeFormat general`."""
    s = f"{t:.{int(precision)}g}"
    return "0" if float(s) == 0.0 else s


def latest_time(case_dir):
    """The largest numeric directory name holding a cloud, or None."""
    best = None
    for d in os.listdir(case_dir):
        try:
            v = float(d)
        except ValueError:
            continue
        if os.path.isdir(os.path.join(case_dir, d, "lagrangian")) and (best is None or v > best[0]):
            best = (v, d)
    return best


def _vol_fields(fp):
    out = []
    for f in (fp or {}).get("uniGasFields", []):
        if f.get("fieldModel") != "uniGasVolFields":
            continue
        pr, tp = f.get("uniGasVolFieldsProperties", {}), f.get("timeProperties", {})
        out.append(dict(name=str(pr["field"]), reset=bool(tp.get("resetAtOutput", False)),
                        until=float(tp.get("resetAtOutputUntilTime", float("inf"))), carry=bool(pr.get("averagingAcrossManyRuns", False)),
                        mfp=bool(pr.get("measureMeanFreePath", False)), err=bool(pr.get("measureErrors", False))))
    return out


def run_case(case_dir, mesh, Cloud, out_dir=None, overrides=None, control=None, seed=7, particles_per_cell=None,
             capacity_factor=8, log=None, keep_lagrangian=False):
    """Runs the case of `case_dir` on `mesh` with the cloud class `Cloud` (UniGasCloud, or the oracle's in tests) from its
    start time to endTime and writes the time directories into `out_dir` (default: the case directory).  `control`
    overrides controlDict entries (e.g. {"endTime": 1e-6}).  Unless keep_lagrangian (the solver's -keep-lagrangian option),
    the parcel files of the previous write time are removed after every write (uniGasCloud::cleanLagrangian,
    U/clouds/uniGasCloud.C:1591-1618).  -> dict(cloud, adapter, written=[time names], time, steps)."""
    out_dir = out_dir or case_dir
    case, ld = cases.from_case_dir(case_dir, mesh, seed=seed, overrides=overrides, particles_per_cell=particles_per_cell)
    ctl = dict(ld["controlDict"], **(control or {}))
    fields = _vol_fields(ld["fieldPropertiesDict"])
    cloud = case.make_cloud(Cloud, parcelCapacity=capacity_factor * max(case.n_parcels, 1024),
                            sampleInterval=foamdict.sample_interval(ld["fieldPropertiesDict"]))
    hybrid = bool(ld.get("hybridDecompositionDict")) and cloud.cfg.collisionModel == _capi.COLLISION_MODEL["hybrid"]
    if hybrid:
        cloud.setHybridDecomposition(ld["hybridDecompositionDict"])
    adapter = None
    if case.uniGasProperties.get("adaptiveSimulation", False):
        adapter = UniGasDynamicAdapter(cloud, case.uniGasProperties)
        if case.subCellLevels is not None:
            adapter.subCellLevels = case.subCellLevels.copy()
    t = float(ctl.get("startTime", 0.0))
    if ctl.get("startFrom", "startTime") == "latestTime":
        lt = latest_time(out_dir)
        if lt is not None and lt[0] > 0.0:  # restart: parcels, cell state, step index, and the averages that are to be carried on
            cloud.readTime(out_dir, lt[1])
            for f in fields:
                if f["carry"] and not os.path.exists(os.path.join(out_dir, lt[1], "uniform", "ugfState.npy")):
                    cloud.readVolFieldsMethod(out_dir, lt[1], f["name"])
            t = lt[0]
    end = float(ctl["endTime"])
    by_step = ctl.get("writeControl", "timeStep") == "timeStep"
    interval = float(ctl.get("writeInterval", 1))
    prec = int(ctl.get("timePrecision", 6))
    next_write = (int(t / interval + 1e-9) + 1) * interval if not by_step else None
    written, steps = [], 0
    n_out, info_counter, t_wall = int(ctl.get("nTerminalOutputs", 1)), 0, time.perf_counter()
    while t < end * (1.0 - 1e-12):
        info_counter += 1
        dt = cloud.cfg.deltaT  # the step runs with the time step set before it; the adapter may change it afterwards
        if adapter is not None:
            if hybrid and adapter.timeSteps == adapter.adaptationInterval - 1:
                adapter.cellCollModelId = cloud.hybridDecomposition()["cellCollModelId"]
            adapter.run(1)
        else:
            cloud.evolve(1)
        t += dt
        steps += 1
        due = (steps % int(interval) == 0) if by_step else (t >= next_write * (1.0 - 1e-9))
        if due or t >= end * (1.0 - 1e-12):
            name = time_name(t, prec)
            # uniGasVolFields::calculateField at a write time (:839-1507): fields, then the reset, then the accumulator dictionary
            for f in fields:
                cloud.writeFields(out_dir, name, f["name"], resetAtOutput=f["reset"] and t < f["until"] + 0.5 * dt,
                                  measureMeanFreePath=f["mfp"], measureErrors=f["err"])
            cloud.writeTime(out_dir, name, fieldNames=[f["name"] for f in fields if f["carry"]])
            written.append(name)
            if not keep_lagrangian and len(written) >= 2:
                shutil.rmtree(os.path.join(out_dir, written[-2], "lagrangian"), ignore_errors=True)
            if not by_step:
                while next_write * (1.0 - 1e-9) <= t:
                    next_write += interval
        if log and info_counter >= n_out:  # uniGasFoam.C:78-104, uniGasCloud::info (uniGasCloud.C:878-920)
            i, c = cloud.info(), cloud.counters()
            log(f"Time = {time_name(t, prec)}")
            log(f"    Number of particles             = {i['nParticles']}")
            if i["nParticles"]:
                log(f"    Average linear kinetic energy   = {i['avgLinearKE']}")
                log(f"    Average rotational energy       = {i['avgRotationalE']}")
                log(f"    Total energy                    = {i['totalEnergy']}")
            log(f"    DSMC Collisions                 = {c['collisions']}")
            log(f"    BGK relaxations                 = {c['bgkRelaxations']}")
            log(f"ClockTime = {time.perf_counter() - t_wall:.3f} s")
            info_counter = 0
    return dict(cloud=cloud, adapter=adapter, written=written, time=t, steps=steps, case=case)
