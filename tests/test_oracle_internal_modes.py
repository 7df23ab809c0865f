"""Vibrational (quantum-kinetic) and multi-level electronic energy exchange - SURVEY 8 row a9 - on the CPU restatement, pinned to
closed forms: energy conservation per collision including all four modes, relaxation of an excited / cold vibrational mode to the
common temperature (detailed balance of the QK selection, Bird 5.61), vibrationalT / electronicT of a Boltzmann population, wall
re-equilibration.  Reference: U/clouds/uniGasCloud.C:1020-1126, 1192-1326; …/LarsenBorgnakkeVariableHardSphere.C:125-416;
U/cellMeasurements/cellMeasurements.C:436-510; …/uniGasVolFields.C:930-1079."""
import numpy as np
import pytest

from unigasfoam_b200 import cases

kB = cases.kB


def _box(T0, wall="specular", parcels=24000, n=5, Zref=None, zelec=3.0, zrot=3.0, dt_mct=1.0, seed=31, **kw):
    sp = dict(cases.OXYGEN_VIB)
    if Zref is not None:
        sp["Zref"] = [Zref]
    case = cases.closed_box(n=n, parcels=parcels, seed=seed, wall=wall, T0=T0, binary="LarsenBorgnakkeVariableHardSphere", species=("O2", sp),
                            dt_mct=dt_mct, Trot=T0, rotationalRelaxationCollisionNumber=zrot, electronicRelaxationCollisionNumber=zelec, **kw)
    return case, sp


def _energies(p, sp):
    m = sp["mass"]
    etr = 0.5 * m * (p["U"] ** 2).sum(1)
    evib = p["vibLevel"][:, 0] * kB * sp["characteristicVibrationalTemperature"][0]
    eel = np.asarray(sp["electronicEnergyList"])[p["ELevel"]]
    return etr, p["ERot"], evib, eel


def test_collisions_conserve_energy_over_all_modes(OracleCloud):
    case, sp = _box(6000.0, Zref=2.0)
    cases.with_internal_modes(case)
    cl = case.make_cloud(OracleCloud)
    cl.buildCellOccupancy(); cl.reorder()
    before = cl.parcels()
    cl.collide()
    after = cl.parcels()
    nC = case.mesh.n_cells
    eb = np.bincount(before["cell"], sum(_energies(before, sp)), nC)
    ea = np.bincount(after["cell"], sum(_energies(after, sp)), nC)
    assert (np.abs(ea - eb) <= 1e-12 * eb).all()  # every collision conserves the sum of all modes (pairs never leave their cell)
    m = sp["mass"]
    for k in range(3):
        pb = np.bincount(before["cell"], m * before["U"][:, k], nC)
        pa = np.bincount(after["cell"], m * after["U"][:, k], nC)
        assert (np.abs(pa - pb) <= 1e-12 * np.bincount(before["cell"], m * np.abs(before["U"][:, k]), nC)).all()
    assert cl.counters()["collisions"] > 500
    assert (after["vibLevel"] != before["vibLevel"]).sum() > 50 and (after["ELevel"] != before["ELevel"]).sum() > 50
    assert after["vibLevel"].min() >= 0 and after["ELevel"].min() >= 0 and after["ELevel"].max() <= 2
    c = cl.counters()
    et, er, ev, ee = _energies(after, sp)
    assert abs(c["vibrationalEnergy"] - ev.sum()) <= 1e-12 * ev.sum() and abs(c["electronicEnergy"] - ee.sum()) <= 1e-12 * ee.sum()


def test_cold_vibration_relaxes_to_the_common_temperature(OracleCloud):
    """Adiabatic box, vibration and electronic levels start in the ground state at T_tr = T_rot = 6000 K: the total energy stays put
    and the modes meet at one temperature - the mean vibrational level and the level populations are then Boltzmann at it."""
    case, sp = _box(6000.0, Zref=1.0, parcels=30000, n=4, dt_mct=2.0)
    case.vibLevel = np.zeros((case.n_parcels, 1), np.int32)
    case.ELevel = np.zeros(case.n_parcels, np.int32)
    cl = case.make_cloud(OracleCloud)
    p0 = cl.parcels()
    e0 = sum(e.sum() for e in _energies(p0, sp))
    cl.evolve(60)
    p = cl.parcels()
    et, er, ev, ee = _energies(p, sp)
    assert abs((et.sum() + er.sum() + ev.sum() + ee.sum()) / e0 - 1.0) < 1e-10
    n = case.n_parcels
    Ttr = et.mean() / (1.5 * kB)
    Trot = er.mean() / kB
    th = sp["characteristicVibrationalTemperature"][0]
    Tvib = th / np.log(1.0 + 1.0 / p["vibLevel"][:, 0].mean())
    assert Ttr < 5600.0                       # energy went into the cold modes
    assert abs(Trot / Ttr - 1.0) < 0.04 and abs(Tvib / Ttr - 1.0) < 0.06
    E, g = np.asarray(sp["electronicEnergyList"]), np.asarray(sp["degeneracyList"], float)
    pop = np.bincount(p["ELevel"], minlength=3) / n
    w = g * np.exp(-E / (kB * Ttr)); w /= w.sum()
    assert np.abs(pop - w).max() < 0.03


def test_vibrational_and_electronic_temperature_fields(OracleCloud):
    """A Boltzmann population at 4000 K, no collisions: vibrationalT and electronicT of uniGasVolFields return 4000 K, overallT the
    dof-weighted mean of equal temperatures."""
    case, sp = _box(4000.0, parcels=120000, n=3)
    case.uniGasProperties["dsmcCollisionModel"] = "noDSMCCollision"
    cases.with_internal_modes(case)
    cl = case.make_cloud(OracleCloud)
    cl.evolve(4)
    f = cl.fields()
    assert np.abs(f["vibrationalT"] / 4000.0 - 1.0).max() < 0.06
    assert np.abs(f["electronicT"] / 4000.0 - 1.0).max() < 0.12
    assert np.abs(f["overallT"] / 4000.0 - 1.0).max() < 0.04
    I = cl.internalAccumulators()
    acc = cl.accumulators()
    np.testing.assert_allclose(I[:, 0, 0], acc["acc"][:, 0], rtol=1e-13)  # nParcels of the single species = the cell count sum
    mom = None
    cl.calculateFields()
    mom = cl.cellMoments()
    p = cl.parcels()
    ev = p["vibLevel"][:, 0] * kB * sp["characteristicVibrationalTemperature"][0]
    np.testing.assert_allclose(mom[:, 0, 22], np.bincount(p["cell"], ev, case.mesh.n_cells), rtol=1e-12)
    np.testing.assert_allclose(mom[:, 0, 27], mom[:, 0, 22], rtol=1e-13)  # one mode: per-mode sum = total
    np.testing.assert_allclose(mom[:, 0, 23], np.bincount(p["cell"], ev * p["U"][:, 0], case.mesh.n_cells), rtol=1e-10, atol=1e-30)
    np.testing.assert_allclose(mom[:, 0, 26], np.bincount(p["cell"], np.asarray(sp["electronicEnergyList"])[p["ELevel"]], case.mesh.n_cells), rtol=1e-12)


def test_diffuse_walls_reset_the_levels_to_the_wall_temperature(OracleCloud):
    """diffuseReflection redraws vibLevel and ELevel at the wall temperature (uniGasPatchBoundary.C:373-383): a collision-free gas
    with highly excited levels between 500 K walls loses the excitation wall hit by wall hit."""
    case, sp = _box(500.0, wall="diffuse", parcels=30000, n=4, dt_mct=3.0)
    case.uniGasProperties["dsmcCollisionModel"] = "noDSMCCollision"
    cases.with_internal_modes(case, Tvib=9000.0, Tel=9000.0)
    cl = case.make_cloud(OracleCloud)
    v0 = case.vibLevel[:, 0].mean()
    for _ in range(60):
        cl.evolve(1)
    p = cl.parcels()
    th = sp["characteristicVibrationalTemperature"][0]
    expect = 1.0 / (np.exp(th / 500.0) - 1.0)          # ~0.011
    assert v0 > 3.0 and p["vibLevel"][:, 0].mean() < 0.1 * v0
    hit = p["vibLevel"][:, 0] == 0
    assert hit.mean() > 0.9 and abs(p["vibLevel"][:, 0].mean() - expect) < 0.2 + 0.1 * v0 * (1 - hit.mean())
    assert (p["ELevel"] == 0).mean() > 0.9              # 500 K: the excited electronic levels are empty
    f = cl.fields()
    assert np.isfinite(f["surfaceHeatTransfer"]).all()
    assert f["surfaceHeatTransfer"].sum() > 0           # the de-excitation energy went into the walls
