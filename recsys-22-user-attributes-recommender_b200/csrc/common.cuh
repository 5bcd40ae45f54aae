// Shared device/host helpers for the asme_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/asme_b200.h"

// ---------------------------------------------------------------------------------------------
// error reporting: every extern "C" entry point returns 0 or a negative code and records a message
// ---------------------------------------------------------------------------------------------
void asme_set_error(const char* fmt, ...);

#define ASME_REQUIRE(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            asme_set_error(__VA_ARGS__);         \
            return ASME_ERR_INVALID;             \
        }                                        \
    } while (0)

#define ASME_CUDA_OK(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            asme_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ASME_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

// Opt a kernel in to the full 227 KB of dynamic shared memory ONCE per process (the attribute call is not something to repeat
// on every launch, and it must not happen while a stream is being captured into a CUDA graph).
int asme_ensure_max_smem(const void* kernel);
#define ASME_MAX_DYN_SMEM 232448
void asme_count_launch();   // every kernel launch of this library is counted (bench.py reports it)
#define ASME_LAUNCH_OK()                  \
    do {                                  \
        asme_count_launch();              \
        ASME_CUDA_OK(cudaGetLastError()); \
    } while (0)

__host__ __device__ static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#define ASME_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// Device-side row counts.  Buffers of a row selection (asme_b200_select_rows) are allocated for the CAPACITY (every position
// could be selected); how many rows are live is only known on the device.  Kernels on that path take the capacity as their row
// count plus ``n_live`` (device pointer, may be NULL = all rows live) and touch the live rows only: one CUDA graph serves
// every batch of a shape, whatever its number of selected rows, and the host never waits for the count.
__device__ __forceinline__ int asme_live_rows(int capacity, const int32_t* __restrict__ n_live) {
    if (n_live == nullptr) return capacity;
    const int n = __ldg(n_live);
    return n < capacity ? n : capacity;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// reduction over an aligned sub-group of LANES lanes (LANES power of two <= 32)
// The member mask names ONLY the group's lanes: groups of one warp may run different trip counts of a
// grid-stride loop, so a full-warp mask would wait for lanes that already left the loop (deadlock).
template <int LANES>
__device__ __forceinline__ unsigned group_mask() {
    if (LANES >= 32) return 0xffffffffu;
    return ((1u << (LANES & 31)) - 1u) << (((threadIdx.x & 31) / LANES) * LANES);
}
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
    const unsigned mask = group_mask<LANES>();
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx gelu(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Tensor-core (bf16 policy) epilogues: the GELU evaluations are what the FFN GEMM epilogues spend their issue slots on (erff is
// ~25 instructions, erff + expf ~45; an Abramowitz & Stegun 7.1.26 erf with MUFU.EX2 / MUFU.RCP was 17).  These epilogues store
// bf16 (2^-9 relative), so the erf form x * Phi(x) is evaluated as x * sigmoid(2u), u = x (a + b x^2 + c x^4): the classic
// tanh form with one more term, coefficients from a minimax fit against 0.5 x (1 + erf(x / sqrt 2)) on [-8, 8]:
// |error| <= 2.6e-5 absolute for every x (1/60 of a bf16 ulp at 1), derivative within 1.1e-4.  Seven instructions:
// x^2 (clamped at 64, beyond which the sigmoid is saturated and the quartic would turn around), two FMAs for
// w = -2 log2(e) (a + b x^2 + c x^4), x w, MUFU.EX2, 1 + e, MUFU.RCP, x * r.  Used where the epilogue stores bf16 only; an epilogue that also
// writes an fp32 copy uses the A&S erf below (3e-7), and the fp32 strict-parity path keeps erff / expf.
__device__ __forceinline__ float exp2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {      // one MUFU.RCP (__frcp_rn is a Newton step plus a slow-path call)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// erf form for the epilogues that also hand out an fp32 copy: Abramowitz & Stegun 7.1.26, |erf error| <= 3e-7 (17 instructions)
__device__ __forceinline__ float erf_abs_as(float ax, float e) {   // erf(|x|) given e = exp(-x*x)
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    return fmaf(-p * t, e, 1.0f);
}
__device__ __forceinline__ float gelu_erf_as(float x) {
    const float ax = fabsf(x) * 0.70710678118654752440f;
    const float y = erf_abs_as(ax, exp2_approx(-ax * ax * 1.4426950408889634f));
    return 0.5f * x * (1.0f + copysignf(y, x));
}
__device__ __forceinline__ float gelu_erf_grad_as(float x) {
    const float ax = fabsf(x) * 0.70710678118654752440f;
    const float e = exp2_approx(-ax * ax * 1.4426950408889634f);   // exp(-x^2 / 2)
    const float cdf = 0.5f * (1.0f + copysignf(erf_abs_as(ax, e), x));
    return fmaf(x * 0.39894228040143267794f, e, cdf);
}
#define GELU_FIT_A 7.97507884e-01f
#define GELU_FIT_B 3.70056460e-02f
#define GELU_FIT_C -3.51516788e-04f
#define GELU_M2LOG2E -2.8853900817779268f               // -2 log2(e)
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float x2 = fminf(x * x, 64.0f);
    float w = fmaf(x2, GELU_M2LOG2E * GELU_FIT_C, GELU_M2LOG2E * GELU_FIT_B);
    w = fmaf(x2, w, GELU_M2LOG2E * GELU_FIT_A);
    const float e = exp2_approx(x * w);                  // exp(-2u); +inf for very negative x -> rcp -> 0 -> -0
    return x * rcp_approx(1.0f + e);
}
// d/dx [x sigmoid(v)], v = 2u: sigmoid + x sigmoid (1 - sigmoid) v', v' = 2 (a + 3 b x^2 + 5 c x^4)
__device__ __forceinline__ float gelu_erf_grad_fast(float x) {
    const float x2 = fminf(x * x, 64.0f);
    float w = fmaf(x2, GELU_M2LOG2E * GELU_FIT_C, GELU_M2LOG2E * GELU_FIT_B);
    w = fmaf(x2, w, GELU_M2LOG2E * GELU_FIT_A);
    const float sg = rcp_approx(1.0f + exp2_approx(x * w));
    float dv = fmaf(x2, 10.0f * GELU_FIT_C, 6.0f * GELU_FIT_B);
    dv = fmaf(x2, dv, 2.0f * GELU_FIT_A);
    return fmaf(x * sg * (1.0f - sg), dv, sg);
}

// Counter-based dropout generator: the mask is a pure function of (seed, site, element index), so the backward pass recomputes
// it instead of storing it, and every kernel that touches a site (fused GEMM epilogues, the row-wise kernels, fp32 and
// tensor-core flavours alike) sees the same stream.  Four elements share two 32-bit hashes (lowbias32 finaliser, one 16-bit
// keep decision per element): ~28 instructions per four elements.  (The first version ran Philox4x32-7, ~83 instructions per
// four elements, which made the dropout the largest single item of the FFN GEMM epilogues.)
__device__ __forceinline__ uint32_t asme_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
    return x;
}
// two words = four 16-bit uniform lanes for elements 4*idx4 .. 4*idx4+3 of `site`
__device__ __forceinline__ uint2 dropout_bits4(uint64_t seed, uint32_t site, uint64_t idx4) {
    const uint32_t key = asme_mix32((uint32_t)seed ^ (site * 0x9E3779B9u)) ^ (uint32_t)(seed >> 32);
    const uint32_t a = asme_mix32(((uint32_t)idx4 ^ key) + (uint32_t)(idx4 >> 32) * 0x85EBCA6Bu);
    const uint32_t b = asme_mix32(a + 0x9E3779B9u);
    return make_uint2(a, b);
}
__device__ __forceinline__ uint32_t dropout_thr16(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
// A dropout seed argument with bit 63 set is a DEVICE POINTER (low 48 bits) to the seed: a training step captured in a CUDA
// graph re-reads the seed that a tiny "advance" kernel bumps at the start of every replay (asme_b200_step_state_advance),
// so the by-value launch arguments baked into the graph never change.  Real seeds always have bit 63 clear.
#define ASME_SEED_INDIRECT (1ull << 63)
__device__ __forceinline__ uint64_t asme_seed(uint64_t s) {
    return (s & ASME_SEED_INDIRECT) ? *reinterpret_cast<const uint64_t*>(s & ~ASME_SEED_INDIRECT) : s;
}

// keep-probability scale for element `idx` of dropout site `site`: 0 if dropped, 1/(1-p) if kept
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t site, uint64_t idx, float p, float inv_keep) {
    const uint2 r = dropout_bits4(seed, site, idx >> 2);
    const uint32_t lane = (uint32_t)idx & 3u;
    const uint32_t w = (lane & 2u) ? r.y : r.x;
    const uint32_t u = (lane & 1u) ? (w >> 16) : (w & 0xffffu);
    return u < dropout_thr16(p) ? 0.0f : inv_keep;
}
// the same for four consecutive elements starting at element 4*idx4
__device__ __forceinline__ void dropout_scales4(uint64_t seed, uint32_t site, uint64_t idx4, float p, float inv_keep, float (&s)[4]) {
    const uint2 r = dropout_bits4(seed, site, idx4);
    const uint32_t thr = dropout_thr16(p);
    s[0] = (r.x & 0xffffu) < thr ? 0.f : inv_keep;
    s[1] = (r.x >> 16) < thr ? 0.f : inv_keep;
    s[2] = (r.y & 0xffffu) < thr ? 0.f : inv_keep;
    s[3] = (r.y >> 16) < thr ? 0.f : inv_keep;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void add4(float4& a, const float4& b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
