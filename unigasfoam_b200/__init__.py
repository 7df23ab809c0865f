"""unigasfoam_b200 - B200-native particle loop for uniGasFoam-style DSMC / stochastic-particle BGK.

The package holds only what the hot path (uniGasCloud::evolve) needs: the CUDA kernels
and C ABI (csrc/, built into libugf.so), the host-side mirror of the reference's cloud
interface (cloud.py), and the host-side case/mesh generators that stand in for
blockMesh / decomposePar / uniGasInitialisation (mesh.py, cases.py).
"""
from ._capi import UgfError, libugf  # noqa: F401
from .cloud import UniGasCloud  # noqa: F401
from . import mesh, cases  # noqa: F401
