// Counter-based random numbers shared by the forward-model kernels (rip_fwd.cu, rip_sim.cu).
#pragma once
#include <math.h>
#include <stdint.h>

namespace rip {

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011); one stream per (pixel, purpose)
// ---------------------------------------------------------------------------------------------------------
struct Philox {
    uint32_t c[4], k[2];
    uint32_t out[4];
    int have;
    __device__ __forceinline__ void init(uint64_t seed, uint64_t idx, uint32_t stream) {
        k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
        c[0] = 0u; c[1] = stream; c[2] = (uint32_t)idx; c[3] = (uint32_t)(idx >> 32);
        have = 0;
    }
    __device__ __forceinline__ void round_(uint32_t (&x)[4], uint32_t k0, uint32_t k1) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        const uint32_t hi0 = __umulhi(M0, x[0]), lo0 = M0 * x[0];
        const uint32_t hi1 = __umulhi(M1, x[2]), lo1 = M1 * x[2];
        const uint32_t y0 = hi1 ^ x[1] ^ k0, y1 = lo1, y2 = hi0 ^ x[3] ^ k1, y3 = lo0;
        x[0] = y0; x[1] = y1; x[2] = y2; x[3] = y3;
    }
    __device__ __forceinline__ void gen() {
        uint32_t x[4] = {c[0], c[1], c[2], c[3]};
        uint32_t k0 = k[0], k1 = k[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round_(x, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
        ++c[0];
        have = 4;
    }
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) gen();
        return out[--have];
    }
    // uniform strictly inside (0,1), 23-bit: (x >> 9) + 0.5 is exact in float32 (24 significant bits), so the result lies
    // in [2^-24, 1 - 2^-24]; the earlier 24-bit form rounded its top value to exactly 1.0f.
    // Stream ids in use (one purpose each, disjoint): 1 apportioning, 2/3 scene Poisson, 8 reset noise of the reference
    // pixels, 16+g forward read noise, 32+g / 48+g reference-pixel and white noise per group, 96+k noise-layer draws,
    // 128 Poisson re-sampling, 160 Pearson draws, 200 / 201 cosmic-ray counts / events, 1024+ 1/f frames.
    __device__ __forceinline__ float uniform() { return ((float)(next() >> 9) + 0.5f) * (1.0f / 8388608.0f); }
    // uniform in (0,1), 53-bit
    __device__ __forceinline__ double uniform53() {
        const uint64_t a = next() >> 5, b = next() >> 6;
        return ((double)a * 67108864.0 + (double)b + 0.5) * (1.0 / 9007199254740992.0);
    }
    // two independent normals from one Box-Muller pair
    __device__ __forceinline__ void normal2(float& a, float& b) {
        const float u1 = uniform(), u2 = uniform();
        const float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        a = r * cs;
        b = r * sn;
    }
    __device__ __forceinline__ float normal() {
        const float u1 = uniform(), u2 = uniform();
        return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
};

// ---------------------------------------------------------------------------------------------------------
// Poisson(lam): multiplication method below 10, Hormann's PTRS transformed rejection above (the algorithm of
// NumPy's legacy generator).  Result clipped to int32.
// ---------------------------------------------------------------------------------------------------------
__device__ inline long poisson_draw(Philox& rng, double lam) {
    if (!(lam > 0.0)) return 0;
    if (lam < 10.0) {
        const double enlam = exp(-lam);
        long k = 0;
        double prod = 1.0;
        for (;;) {
            prod *= rng.uniform53();
            if (prod > enlam) ++k;
            else return k;
        }
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        const double U = rng.uniform53() - 0.5, V = rng.uniform53();
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return kf > 2147483647.0 ? 2147483647L : (long)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + kf * loglam - lgamma(kf + 1.0)))
            return kf > 2147483647.0 ? 2147483647L : (long)kf;
    }
}

}  // namespace rip
