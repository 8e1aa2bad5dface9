// Counter-based RNG: Philox4x32 (Salmon et al. 2011), uniform / normal / exponential transforms.
// PHILOX_ROUNDS = 7: the smallest round count the Random123 authors report as passing BigCrush
// ("Philox4x32-7"); 10 is their conservative default.  The step kernel is instruction-issue bound and
// Philox is ~15 % of its instructions, so the three extra rounds are not free.
// Counter layout (rank- and launch-geometry-invariant, so sharding items over GPUs reproduces the
// single-GPU stream):  c0 = index (particle quad, particle, or tile), c1 = step | stream<<16 | sub<<20,
// c2 = global item id, c3 = call offset;  key = 64-bit seed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fastlog.cuh"

namespace sgm {

enum : uint32_t { STREAM_NORMAL = 0, STREAM_UNIFORM = 1, STREAM_GAMMA = 2, STREAM_PARIS = 3, STREAM_EXACT = 4, STREAM_PRED = 5 };

struct RngKey { uint32_t k0, k1, item, offset; };

constexpr int PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += W0; k1 += W1;
    }
    return c;
}
__device__ __forceinline__ uint4 rng_raw(const RngKey& k, uint32_t index, uint32_t step, uint32_t stream, uint32_t sub) {
    return philox4x32(make_uint4(index, (step & 0xffffu) | (stream << 16) | (sub << 20), k.item, k.offset), k.k0, k.k1);
}
// uniforms: f32 in (0, 1] (x * 2^-32 + 2^-33, rounds to 1 with probability 2^-25; every consumer tolerates 1), f64 in (0, 1)
__device__ __forceinline__ float u01f(uint32_t x) { return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }  // (0,1]: I2FP + FFMA
__device__ __forceinline__ double u01d(uint32_t a, uint32_t b) {                                                         // (0,1), 52 bits
    // (2 k + 1) 2^-53 for the 52-bit k = (top 20 bits of a, b): built as a double in [1, 2) and shifted down (exact), which
    // avoids the 64-bit integer -> double conversion
    return __hiloint2double((int)(0x3ff00000u | (a >> 12)), (int)b) - 0.99999999999999989;
}

// 4 uniforms in (0,1) for particle quad q
__device__ __forceinline__ void rng_uniform4(const RngKey& k, uint32_t q, uint32_t step, uint32_t stream, float* u) {
    const uint4 r = rng_raw(k, q, step, stream, 0);
    u[0] = u01f(r.x); u[1] = u01f(r.y); u[2] = u01f(r.z); u[3] = u01f(r.w);
}
__device__ __forceinline__ void rng_uniform4(const RngKey& k, uint32_t q, uint32_t step, uint32_t stream, double* u) {
    const uint4 r0 = rng_raw(k, q, step, stream, 0), r1 = rng_raw(k, q, step, stream, 1);
    u[0] = u01d(r0.x, r0.y); u[1] = u01d(r0.z, r0.w); u[2] = u01d(r1.x, r1.y); u[3] = u01d(r1.z, r1.w);
}
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// 4 standard normals for particle quad q (Box-Muller)
__device__ __forceinline__ void rng_normal4(const RngKey& k, uint32_t q, uint32_t step, float* z) {
    const uint4 r = rng_raw(k, q, step, STREAM_NORMAL, 0);
    // sqrt(-2 ln u) = sqrt(log2(u) * (-2 ln 2)): one multiply between the two MUFU ops
    const float r0 = sqrt_approx(__log2f(u01f(r.x)) * -1.3862943611198906f), r1 = sqrt_approx(__log2f(u01f(r.z)) * -1.3862943611198906f);
    float s0, c0, s1, c1;
    __sincosf(6.28318530717958647692f * u01f(r.y), &s0, &c0);
    __sincosf(6.28318530717958647692f * u01f(r.w), &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}
__device__ __forceinline__ void rng_normal4(const RngKey& k, uint32_t q, uint32_t step, double* z) {
    const uint4 a = rng_raw(k, q, step, STREAM_NORMAL, 0), b = rng_raw(k, q, step, STREAM_NORMAL, 1);
    const double r0 = sqrt(-2.0 * fast_log(u01d(a.x, a.y))), r1 = sqrt(-2.0 * fast_log(u01d(b.x, b.y)));
    double s0, c0, s1, c1;
    fast_sincos2pi(a.z, a.w, &s0, &c0);
    fast_sincos2pi(b.z, b.w, &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}
// one standard normal for (index, step, stream, sub)
__device__ __forceinline__ void rng_normal1(const RngKey& k, uint32_t index, uint32_t step, uint32_t stream, uint32_t sub, float& z) {
    const uint4 r = rng_raw(k, index, step, stream, sub);
    float s, c;
    __sincosf(6.28318530717958647692f * u01f(r.y), &s, &c);
    z = sqrtf(-2.0f * __logf(u01f(r.x))) * c;
}
__device__ __forceinline__ void rng_normal1(const RngKey& k, uint32_t index, uint32_t step, uint32_t stream, uint32_t sub, double& z) {
    const uint4 r = rng_raw(k, index, step, stream, sub);
    double s, c;
    sincospi(2.0 * u01d(r.z, r.w), &s, &c);
    z = sqrt(-2.0 * log(u01d(r.x, r.y))) * c;
}
// Gamma(shape, 1), shape >= 1 (Marsaglia & Tsang 2000) with their squeeze test (no logs ~98 % of the time);
// the variate is generated in f32 (relative resolution 6e-8, far below its Monte-Carlo spread) and
// returned as f64.  Own sub-stream per attempt.
__device__ inline double rng_gamma(const RngKey& k, uint32_t index, uint32_t step, double shape) {
    const float d = (float)(shape - 1.0 / 3.0), c = rsqrtf(9.0f * d);
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {
        const uint4 r = rng_raw(k, index, step, STREAM_GAMMA, attempt);
        float s, co;
        __sincosf(6.28318530717958647692f * u01f(r.y), &s, &co);
        const float x = sqrtf(-2.0f * __logf(u01f(r.x))) * co;
        const float u = u01f(r.z);
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        const float x2 = x * x;
        if (u < 1.0f - 0.0331f * x2 * x2) return (double)(d * v);
        if (__logf(u) < 0.5f * x2 + d * (1.0f - v + __logf(v))) return (double)(d * v);
    }
    return (double)d;   // unreachable in practice (acceptance > 95 % per attempt)
}

}  // namespace sgm
