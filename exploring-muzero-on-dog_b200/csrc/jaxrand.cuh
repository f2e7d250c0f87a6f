// jaxrand.cuh — on-device restatement of jax.random (jax 0.8.1, threefry2x32, partitionable
// mode) as the reference's hot path uses it:
//   MADN/deterministic_madn.py:60-62, MADN/classic_madn.py:70-72,238-240, DOG/dog.py:102-104,246-247,
//   MuZero_det_MADN/game_agent.py:60,187-188, MuZero_det_MADN/evaluate_agent.py:337,741.
// Algorithms: Random123 Threefry-2x32-20; jax/_src/prng.py (_threefry_split_foldlike,
// _threefry_random_bits_partitionable); jax/_src/random.py (uniform, randint, choice, gumbel).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dogstep {

struct Key2 {
  uint32_t a, b;
};

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

// x0 + x1 of a Threefry round.  On the device it is written as  x1 * 1 + x0  with the 1 read from the constant bank, so that
// ptxas keeps it an IMAD on the fma pipe: shift and xor of a round already fill the integer-alu pipe (half rate), which is what
// bounds the play kernels.  Measured (scripts/microbench/threefry_variants.cu, V3): 349 -> 381 G Threefry/s at 16 warps per SM.
#ifdef __CUDACC__
static __constant__ uint32_t c_dogstep_one = 1u;
#endif
__host__ __device__ __forceinline__ uint32_t tf_add(uint32_t x0, uint32_t x1) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x1), "r"(c_dogstep_one), "r"(x0));
  return r;
#else
  return x0 + x1;
#endif
}

// a + k + c of a key injection (c = 1..5).  With both adds forced onto the fma pipe (DOGSTEP_TF_INJECT_IMAD) the config-2 play
// kernel was 3 % SLOWER (1.80 against 1.75 ms): k + c is an integer add the compiler folds with its neighbours.
__host__ __device__ __forceinline__ uint32_t tf_inject(uint32_t a, uint32_t k, uint32_t c) {
#if defined(__CUDA_ARCH__) && defined(DOGSTEP_TF_INJECT_IMAD)
  uint32_t t, r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(k), "r"(c_dogstep_one), "r"(c));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(t), "r"(c_dogstep_one), "r"(a));
  return r;
#else
  return tf_add(a, k + c);
#endif
}

// x1 = rotl(x1, r) ^ x0 of a Threefry round.  DOGSTEP_TF_MULROT (an experiment switch, default 0) selects the rounds (bit j of the
// mask = round j of a group of four) that rotate by a 32 x 32 -> 64 multiply with 2^r on the fma pipe followed by ONE three-input
// xor (lo ^ hi ^ x0), instead of funnel shift + xor (two alu-pipe instructions); the multiplier is read from the constant bank so
// that ptxas does not turn it back into a shift.  Measured in the config-2 play kernel (ALU pipe ~90 % busy, fma pipe half
// idle): one round in four 1.75 -> 1.89 ms, two in four 2.01 ms, all 2.24 ms — the wide multiply costs more than it frees.
#ifndef DOGSTEP_TF_MULROT
#define DOGSTEP_TF_MULROT 0
#endif
#ifdef __CUDACC__
static __constant__ uint32_t c_dogstep_pow2[8] = {1u << 13, 1u << 15, 1u << 26, 1u << 6, 1u << 17, 1u << 29, 1u << 16, 1u << 24};
#endif
template <int J, int R>
__host__ __device__ __forceinline__ uint32_t tf_rotx(uint32_t x1, uint32_t x0) {
#ifdef __CUDA_ARCH__
  if ((DOGSTEP_TF_MULROT >> (J & 3)) & 1) {
    uint32_t lo, hi;
    asm("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x1), "r"(c_dogstep_pow2[J]));
    return lo ^ hi ^ x0;
  }
#endif
  return rotl32(x1, R) ^ x0;
}

// Threefry-2x32, 20 rounds.  Fully unrolled: 20 x (IMAD, SHF, LOP) + 5 key injections.
__host__ __device__ __forceinline__ Key2 threefry2x32(Key2 k, uint32_t c0, uint32_t c1) {
  const uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
  uint32_t x0 = c0 + ks0, x1 = c1 + ks1;
#define DOGSTEP_TF_ROUND(j, r) \
  x0 = tf_add(x0, x1);         \
  x1 = tf_rotx<j, r>(x1, x0);
#define DOGSTEP_TF_GROUP_A DOGSTEP_TF_ROUND(0, 13) DOGSTEP_TF_ROUND(1, 15) DOGSTEP_TF_ROUND(2, 26) DOGSTEP_TF_ROUND(3, 6)
#define DOGSTEP_TF_GROUP_B DOGSTEP_TF_ROUND(4, 17) DOGSTEP_TF_ROUND(5, 29) DOGSTEP_TF_ROUND(6, 16) DOGSTEP_TF_ROUND(7, 24)
  DOGSTEP_TF_GROUP_A
  x0 = tf_add(x0, ks1); x1 = tf_inject(x1, ks2, 1u);
  DOGSTEP_TF_GROUP_B
  x0 = tf_add(x0, ks2); x1 = tf_inject(x1, ks0, 2u);
  DOGSTEP_TF_GROUP_A
  x0 = tf_add(x0, ks0); x1 = tf_inject(x1, ks1, 3u);
  DOGSTEP_TF_GROUP_B
  x0 = tf_add(x0, ks1); x1 = tf_inject(x1, ks2, 4u);
  DOGSTEP_TF_GROUP_A
  x0 = tf_add(x0, ks2); x1 = tf_inject(x1, ks0, 5u);
#undef DOGSTEP_TF_ROUND
#undef DOGSTEP_TF_GROUP_A
#undef DOGSTEP_TF_GROUP_B
  return Key2{x0, x1};
}

// jax.random.PRNGKey(int32 seed) -> raw key [0, seed]
__host__ __device__ __forceinline__ Key2 prng_key(int32_t seed) { return Key2{0u, (uint32_t)seed}; }

// jax.random.split(key, n)[i]
__host__ __device__ __forceinline__ Key2 split_i(Key2 k, uint32_t i) { return threefry2x32(k, 0u, i); }

// jax.random.bits(key, (n,), uint32)[i]
__host__ __device__ __forceinline__ uint32_t bits_i(Key2 k, uint32_t i) {
  Key2 o = threefry2x32(k, 0u, i);
  return o.a ^ o.b;
}

__host__ __device__ __forceinline__ float bits_to_unit_float(uint32_t bits) {
  uint32_t u = (bits >> 9) | 0x3F800000u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u) - 1.0f;
#else
  float f;
  memcpy(&f, &u, 4);
  return f - 1.0f;
#endif
}

// jax.random.uniform(key, shape, f32, minval, maxval)[i]; explicit mul/add so no FMA contraction
__device__ __forceinline__ float uniform_i(Key2 k, uint32_t i, float minval, float maxval) {
  float f = bits_to_unit_float(bits_i(k, i));
  float v = __fadd_rn(__fmul_rn(f, __fsub_rn(maxval, minval)), minval);
  return fmaxf(minval, v);
}

// jax.random.randint(key, shape, lo, hi)[i] (uint32 arithmetic with wraparound)
__host__ __device__ __forceinline__ int32_t randint_i(Key2 k, uint32_t i, int32_t lo, int32_t hi) {
  Key2 k1 = split_i(k, 0), k2 = split_i(k, 1);
  uint32_t hb = bits_i(k1, i), lb = bits_i(k2, i);
  uint32_t span = (hi <= lo) ? 1u : (uint32_t)(hi - lo);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  uint32_t off = ((hb % span) * mult + (lb % span)) % span;
  return (int32_t)((uint32_t)lo + off);
}

// jax.random.gumbel "low" mode sample i.  logf is CUDA's (<= 1 ulp); see DESIGN.md on float parity.
__device__ __forceinline__ float gumbel_i(Key2 k, uint32_t i) {
  float u = uniform_i(k, i, 1.17549435e-38f, 1.0f);
  return -logf(-logf(u));
}

// jax.random.choice(key, 6 items, p) -> index; cumsum uses XLA:CPU's association for n = 6.
__device__ __forceinline__ int choice6(Key2 k, const float p[6]) {
  float s01 = __fadd_rn(p[0], p[1]), s23 = __fadd_rn(p[2], p[3]), s45 = __fadd_rn(p[4], p[5]);
  float c[6];
  c[0] = p[0];
  c[1] = s01;
  c[2] = __fadd_rn(s01, p[2]);
  c[3] = __fadd_rn(s01, s23);
  c[4] = __fadd_rn(c[3], p[4]);
  c[5] = __fadd_rn(c[3], s45);
  float u = uniform_i(k, 0, 0.0f, 1.0f);
  float r = __fmul_rn(c[5], __fsub_rn(1.0f, u));
  int idx = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) idx += (c[j] < r);
  return idx > 5 ? 5 : idx;
}

}  // namespace dogstep
