// ugf_rng.cuh — counter-based Philox4x32-10 streams (Salmon et al., SC'11).
//
// Replaces the single shared Foam::Random of the reference (U/clouds/uniGasCloud.H:207,
// seeded from the wall clock at uniGasCloud.C:490).  A stream is addressed by
//   key = (seed.lo, seed.hi ^ kind<<24 ^ aux),  counter = (a, b, c, block)
// so that every parcel / cell / collision candidate owns its own reproducible sequence
// regardless of which thread runs it.  Stream addresses per phase are listed in DESIGN.md §RNG.
#pragma once
#include "ugf_common.cuh"

namespace ugf {

struct Stream {
    uint32_t k0, k1, c0, c1, c2, c3;
    uint32_t o0, o1, o2, o3;
    int have;

    __device__ __forceinline__ Stream(uint64_t seed, uint32_t kind, uint32_t aux, uint32_t a, uint32_t b, uint32_t c) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32) ^ (kind << 24) ^ aux;
        c0 = a; c1 = b; c2 = c; c3 = 0;
        have = 0;
        o0 = o1 = o2 = o3 = 0;
    }

    __device__ __forceinline__ void block() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t n0 = hi1 ^ x1 ^ a, n2 = hi0 ^ x3 ^ b;
            x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        o0 = x0; o1 = x1; o2 = x2; o3 = x3;
        c3++;
    }

    // uniform in [0,1), 53 random bits: Random::sample01<scalar>()
    __device__ __forceinline__ double u01() {
        if (!have) { block(); have = 2; }
        const uint32_t hi = (have == 2) ? o0 : o2;
        const uint32_t lo = (have == 2) ? o1 : o3;
        have--;
        return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
    }

    // two independent N(0,1) (Box-Muller); fixed draw count, unlike Foam::Random's cached polar method
    __device__ __forceinline__ void gauss2(double& g1, double& g2) {
        const double u1 = u01(), u2 = u01();
        const double r = sqrt(-2.0 * log(1.0 - u1));
        double s, c;
        sincos(TWO_PI * u2, &s, &c);
        g1 = r * c;
        g2 = r * s;
    }

    __device__ __forceinline__ void gauss3(double& g0, double& g1, double& g2) {
        double d;
        gauss2(g0, g1);
        gauss2(g2, d);
    }

    // Random::position<label>(0, n-1)
    __device__ __forceinline__ int position(int n) {
        const int i = (int)(u01() * n);
        return i < n - 1 ? i : n - 1;
    }
};

}  // namespace ugf
