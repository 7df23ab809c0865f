#!/usr/bin/env python
"""Generates tests/golden/*.npz: small seeded cases run through the CPU oracle (oracle/ugf_oracle.cpp).

The reference ships no golden vectors for this path and cannot be built here (no OpenFOAM), so these fixtures
pin the *restatement*: they freeze what the oracle - itself pinned by closed-form kinetic theory in
tests/test_oracle_physics.py - produced when they were generated, so that neither the oracle nor the CUDA path can
drift unnoticed, and so that the GPU tests have a reference that does not need the oracle at run time.

    python tests/golden/make_golden.py        # rewrites the fixtures (only after a deliberate change of the contract)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from unigasfoam_b200 import cases  # noqa: E402


def golden_cases():
    """name -> (case, steps).  Small enough to commit (a few hundred KB in total)."""
    out = {}
    out["closed_box_vhs"] = (cases.closed_box(n=4, parcels=3000, seed=51, dt_mct=0.7), 6)
    out["closed_box_lb_n2"] = (cases.closed_box(n=4, parcels=3000, seed=52, dt_mct=0.7, binary="LarsenBorgnakkeVariableHardSphere",
                                                species=("N2", cases.NITROGEN), Trot=220.0, wall="diffuse",
                                                rotationalRelaxationCollisionNumber=5.0, electronicRelaxationCollisionNumber=500.0), 6)
    out["couette_diffuse"] = (cases.couette(nx=12, ny=8, ppc=30, Kn=0.5), 6)
    out["closed_box_usp_sbgk"] = (cases.closed_box(n=4, parcels=3000, seed=53, dt_mct=1.5, mode="bgk", bgk="unifiedStochasticParticleSBGK",
                                                   binary="noDSMCCollision", theta=0.3, velocity=(40.0, -10.0, 5.0)), 5)
    out["cylinder_inflow"] = (cases.cylinder(nr=8, ntheta=16, ppc=12), 5)
    return out


def run(case, steps, cloud_cls, **kw):
    cl = case.make_cloud(cloud_cls, seed=777, **kw)
    cl.evolve(steps)
    p, c = cl.parcels(), cl.counters()
    cl.close()
    return p, c


def main():
    from oracle.oracle_cloud import OracleCloud
    for name, (case, steps) in golden_cases().items():
        p, c = run(case, steps, OracleCloud)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), position=p["position"], U=p["U"], cell=p["cell"], ERot=p["ERot"],
                            counters=np.array([c[k] for k in ("nParcels", "collisionCandidates", "collisions", "bgkRelaxations", "inserted", "deleted", "wallHits")], np.int64),
                            energy=np.array([c["linearKineticEnergy"], c["rotationalEnergy"]]))
        print(name, len(p["cell"]), c["collisions"], c["bgkRelaxations"], c["inserted"], c["deleted"], c["wallHits"])


if __name__ == "__main__":
    main()
