"""The oracle's Philox4x32-10 against the Random123 known-answer vectors (Salmon et al., SC'11 kat_vectors)."""
import ctypes as C

import numpy as np

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers(oracle_api):
    f = oracle_api.lib.ugfo_philox
    f.restype = None
    for ctr, key, want in KAT:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        f(c, k, o)
        assert tuple(o) == want


def test_stream_uniforms(oracle_api):
    f = oracle_api.lib.ugfo_stream_u01
    f.restype = None
    n = 200000
    out = np.empty(n)
    f(C.c_uint64(20261017), C.c_uint32(2), C.c_uint32(0), C.c_uint32(3), C.c_uint32(5), C.c_uint32(7), C.c_int32(n),
      out.ctypes.data_as(C.POINTER(C.c_double)))
    assert out.min() >= 0.0 and out.max() < 1.0
    assert abs(out.mean() - 0.5) < 4 / np.sqrt(12 * n)
    assert abs(out.var() - 1 / 12) < 5e-4
    # distinct stream addresses give distinct sequences
    out2 = np.empty(16)
    f(C.c_uint64(20261017), C.c_uint32(2), C.c_uint32(0), C.c_uint32(3), C.c_uint32(5), C.c_uint32(8), C.c_int32(16),
      out2.ctypes.data_as(C.POINTER(C.c_double)))
    assert not np.array_equal(out[:16], out2)
