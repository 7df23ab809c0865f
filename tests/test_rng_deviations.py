"""The declared deviations of the random-number contract from the reference (DESIGN.md section RNG), each tested for the
distributional equivalence it claims:

  * BGK relaxing subset: the reference shuffles the cell's parcel list five times (Fisher-Yates, Foam::Random::shuffle) and
    relaxes the first nRel (…/unifiedStochasticParticleSBGK.C:912-925); here every parcel draws one uniform key from its own
    Philox stream and the nRel smallest keys are taken.  Both must be a uniformly random nRel-subset.
  * Gaussians: Foam::Random::GaussNormal is the polar (Marsaglia) method with a cached second deviate; here Box-Muller pairs.
  * exponentials: -log(1 - u) with u in [0, 1) where the reference takes -log(u).
"""
import ctypes as C
import math

import numpy as np


def stream(api, kind, aux, a, b, c, n, seed=20261017):
    f = api.lib.ugfo_stream_u01
    f.restype = None
    out = np.empty(n)
    f(C.c_uint64(seed), C.c_uint32(kind), C.c_uint32(aux), C.c_uint32(a), C.c_uint32(b), C.c_uint32(c), C.c_int32(n), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def ks_uniformity(p_values):
    """Kolmogorov-Smirnov distance of a sample of p-values / probabilities from the uniform distribution, times sqrt(n)."""
    x = np.sort(np.asarray(p_values))
    n = len(x)
    d = max(np.abs(x - np.arange(1, n + 1) / n).max(), np.abs(x - np.arange(0, n) / n).max())
    return d * math.sqrt(n)


def test_key_rank_subset_is_a_uniform_subset_like_the_shuffle(oracle_api):
    N, nRel, trials = 12, 4, 6000
    # (a) the product's rule with the product's streams: key of parcel i in cell c at step s = first uniform of stream (3; s, c, i)
    incl = np.zeros(N)
    pair = np.zeros((N, N))
    for t in range(trials):
        keys = np.array([stream(oracle_api, 3, 0, t, 7, i, 1)[0] for i in range(N)])
        sel = np.argsort(keys, kind="stable")[:nRel]
        incl[sel] += 1
        pair[np.ix_(sel, sel)] += 1
    # (b) the reference's rule: five Fisher-Yates passes from the top index down, then the first nRel
    rng = np.random.default_rng(5)
    incl_ref = np.zeros(N)
    pair_ref = np.zeros((N, N))
    for t in range(trials):
        lst = np.arange(N)
        for _ in range(5):
            for i in range(N - 1, 0, -1):
                j = int(rng.random() * (i + 1))
                lst[i], lst[j] = lst[j], lst[i]
        sel = lst[:nRel]
        incl_ref[sel] += 1
        pair_ref[np.ix_(sel, sel)] += 1
    # every parcel is included with probability nRel / N, every pair with nRel (nRel - 1) / (N (N - 1)): chi-square against the
    # exact expectation, for both rules, and the two rules against each other
    p1 = nRel / N
    p2 = nRel * (nRel - 1) / (N * (N - 1))
    for inc, pr in ((incl, pair), (incl_ref, pair_ref)):
        chi_single = ((inc - trials * p1) ** 2 / (trials * p1 * (1 - p1))).sum()
        assert chi_single < 35.0, chi_single  # 12 cells (11 dof after the fixed total): 99.9 % quantile 31.3; margin for the constraint
        iu = np.triu_indices(N, 1)
        chi_pair = ((pr[iu] - trials * p2) ** 2 / (trials * p2 * (1 - p2))).sum()
        assert chi_pair < 110.0, chi_pair     # 66 pairs: 99.9 % quantile of chi2(66) = 107
    assert np.abs(incl / trials - incl_ref / trials).max() < 5 * math.sqrt(2 * p1 * (1 - p1) / trials)


def test_box_muller_and_polar_gaussians_are_both_standard_normal(oracle_api):
    n = 200000
    u = stream(oracle_api, 1, 0, 11, 13, 17, 2 * n)
    g_bm = np.sqrt(-2.0 * np.log(1.0 - u[0::2])) * np.cos(2.0 * np.pi * u[1::2])  # the product's pairing (first deviate)
    # the reference's polar method on the same uniform source
    v = 2.0 * stream(oracle_api, 1, 0, 11, 13, 18, 4 * n).reshape(-1, 2) - 1.0
    r2 = (v ** 2).sum(1)
    ok = (r2 < 1.0) & (r2 > 0.0)
    g_polar = (v[ok, 0] * np.sqrt(-2.0 * np.log(r2[ok]) / r2[ok]))[:n]
    from math import erf
    cdf = np.vectorize(lambda x: 0.5 * (1.0 + erf(x / math.sqrt(2.0))))
    for g in (g_bm, g_polar):
        assert ks_uniformity(cdf(g[:50000])) < 1.95  # 99.9 % point of the Kolmogorov distribution
        assert abs(g.mean()) < 4 / math.sqrt(len(g)) and abs(g.var() - 1.0) < 0.02
        assert abs((g ** 4).mean() - 3.0) < 0.1
    # and against each other (two-sample KS)
    a, b = np.sort(g_bm[:50000]), np.sort(g_polar[:50000])
    grid = np.concatenate([a, b])
    d = np.abs(np.searchsorted(a, grid, side="right") / len(a) - np.searchsorted(b, grid, side="right") / len(b)).max()
    assert d * math.sqrt(len(a) * len(b) / (len(a) + len(b))) < 1.95


def test_exponential_from_one_minus_u(oracle_api):
    u = stream(oracle_api, 1, 0, 3, 5, 9, 200000)
    for e in (-np.log(1.0 - u), -np.log(np.where(u > 0, u, 0.5))):
        assert ks_uniformity(1.0 - np.exp(-e[:50000])) < 1.95
        assert abs(e.mean() - 1.0) < 0.01 and abs(e.var() - 1.0) < 0.03
