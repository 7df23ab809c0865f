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

// One Philox4x32-10 block.  Deliberately not inlined: a stream is consumed at dozens of call sites (every u01() may
// start a new block) and 10 unrolled rounds per site made the collision / relaxation kernels instruction-cache
// bound (ncu: "no instruction" was the top stall of bgk_kernel).  Arguments and result travel in registers.
__device__ __noinline__ uint4 philox4x32_10(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t a, uint32_t b) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const uint32_t n0 = hi1 ^ x1 ^ a, n2 = hi0 ^ x3 ^ b;
        x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
}

// Box-Muller pair from two uniforms, out of line for the same reason (log + sqrt + sincos in fp64).
__device__ __noinline__ double2 box_muller(double u1, double u2) {
    const double r = sqrt(-2.0 * log(1.0 - u1));
    double s, c;
    sincos(TWO_PI * u2, &s, &c);
    return make_double2(r * c, r * s);
}

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
        const uint4 o = philox4x32_10(c0, c1, c2, c3, k0, k1);
        o0 = o.x; o1 = o.y; o2 = o.z; o3 = o.w;
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

    // the 53 random bits of the next uniform as an integer: u01() == u53() * 2^-53
    __device__ __forceinline__ unsigned long long u53() {
        if (!have) { block(); have = 2; }
        const uint32_t hi = (have == 2) ? o0 : o2;
        const uint32_t lo = (have == 2) ? o1 : o3;
        have--;
        return ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    }

    // two independent N(0,1) (Box-Muller); fixed draw count, unlike Foam::Random's cached polar method
    __device__ __forceinline__ void gauss2(double& g1, double& g2) {
        const double u1 = u01(), u2 = u01();
        const double2 g = box_muller(u1, u2);
        g1 = g.x;
        g2 = g.y;
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
