/* jaxrand_oracle.h — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the pieces of jax.random (jax 0.8.1, threefry2x32,
 * jax_threefry_partitionable=True) that the reference's hot path calls:
 *   MADN/deterministic_madn.py:60-62, MADN/classic_madn.py:70-72,238-240,
 *   DOG/dog.py:102-104,246-247, MuZero_det_MADN/game_agent.py:60,187-188,
 *   MuZero_det_MADN/evaluate_agent.py:322,330,337,350.
 * jax itself is a third-party dependency that is absent from /root/reference and from
 * this image (uv.lock pins jax 0.8.1); this follows its published algorithm
 * (Random123 Threefry-2x32-20 + jax/_src/prng.py, jax/_src/random.py).
 *
 * PARITY STATUS (tests/test_oracle_threefry.py): threefry2x32 is pinned by the three Random123 known-answer vectors;
 * split / uniform by the values printed in JAX's documentation for PRNGKey(0) and PRNGKey(42); split + choice(p) + randint by
 * outputs the reference itself recorded (MADN/jupyter_code/test_functions.ipynb cell 1: die throws 4, 1, 6 for
 * split(PRNGKey(1|2|3))[1] under the uniform distribution, and randint(…, 0, 4) == 2 on the next split of PRNGKey(3)).
 * gumbel / categorical follow the published jax source (argmax of logits - log(-log(uniform))) and have no recorded
 * output to compare with: "parity unpinned" for those two versus a live jax 0.8.1.
 */
#ifndef JAXRAND_ORACLE_H
#define JAXRAND_ORACLE_H
#include <stdint.h>
#include <string.h>
#include <math.h>

static inline uint32_t orc_rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* Threefry-2x32, 20 rounds (Random123); jax/_src/prng.py threefry2x32 lowering. */
static inline void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                    uint32_t *o0, uint32_t *o1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int g = 1; g <= 5; ++g) {
    const int *rot = R[(g - 1) & 1];
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = orc_rotl32(x1, rot[r]);
      x1 ^= x0;
    }
    x0 += ks[g % 3];
    x1 += ks[(g + 1) % 3] + (uint32_t)g;
  }
  *o0 = x0;
  *o1 = x1;
}

/* jax.random.PRNGKey(seed): raw uint32[2] = [hi32(seed), lo32(seed)]; int32 seeds -> hi = 0
 * for non-negative seeds (negative int32 seeds sign-extend to 0xFFFFFFFF in the high word
 * only when x64 is on; with x64 off the seed is converted through int32 -> hi = 0... jax
 * shifts a 32-bit value right by 32 which XLA defines as 0). */
static inline void orc_prngkey(int32_t seed, uint32_t key[2]) {
  key[0] = 0u;
  key[1] = (uint32_t)seed;
}

/* jax.random.split(key, n)[i]  (fold-like split, partitionable mode) */
static inline void orc_split_i(const uint32_t key[2], uint32_t i, uint32_t out[2]) {
  orc_threefry2x32(key[0], key[1], 0u, i, &out[0], &out[1]);
}

/* random_bits(key, 32, shape)[i] for a shape with < 2^32 elements */
static inline uint32_t orc_bits_i(const uint32_t key[2], uint32_t i) {
  uint32_t a, b;
  orc_threefry2x32(key[0], key[1], 0u, i, &a, &b);
  return a ^ b;
}

static inline float orc_bits_to_unit_float(uint32_t bits) {
  uint32_t u = (bits >> 9) | 0x3F800000u;
  float f;
  memcpy(&f, &u, 4);
  return f - 1.0f;
}

/* jax.random.uniform(key, shape, float32, minval, maxval)[i] */
static inline float orc_uniform_i(const uint32_t key[2], uint32_t i, float minval, float maxval) {
  float f = orc_bits_to_unit_float(orc_bits_i(key, i));
  float v = f * (maxval - minval) + minval;
  return v > minval ? v : minval; /* lax.max(minval, ...) */
}

/* jax.random.randint(key, shape, lo, hi, int32)[i] — all arithmetic in uint32 with wraparound */
static inline int32_t orc_randint_i(const uint32_t key[2], uint32_t i, int32_t lo, int32_t hi) {
  uint32_t k1[2], k2[2];
  orc_split_i(key, 0, k1);
  orc_split_i(key, 1, k2);
  uint32_t hb = orc_bits_i(k1, i), lb = orc_bits_i(k2, i);
  uint32_t span = (uint32_t)(hi - lo);
  if (hi <= lo) span = 1u;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  uint32_t off = ((hb % span) * mult + (lb % span)) % span;
  return (int32_t)((uint32_t)lo + off);
}

/* jax.random.gumbel(key, shape)[i], "low" mode:  -log(-log(uniform(minval=tiny, maxval=1))) */
static inline float orc_gumbel_i(const uint32_t key[2], uint32_t i) {
  const float tiny = 1.17549435e-38f;
  float u = orc_uniform_i(key, i, tiny, 1.0f);
  return -logf(-logf(u));
}

/* CPU association of jnp.cumsum for 6 elements */
static inline void orc_cumsum6(const float p[6], float c[6]) {
  float s01 = p[0] + p[1];
  float s23 = p[2] + p[3];
  float s45 = p[4] + p[5];
  c[0] = p[0];
  c[1] = s01;
  c[2] = s01 + p[2];
  c[3] = s01 + s23;
  c[4] = c[3] + p[4];
  c[5] = c[3] + s45;
}

/* jax.random.choice(key, a(6), p=p) scalar draw -> index 0..5 */
static inline int orc_choice6(const uint32_t key[2], const float p[6]) {
  float c[6];
  orc_cumsum6(p, c);
  float u = orc_uniform_i(key, 0, 0.0f, 1.0f);
  float r = c[5] * (1.0f - u);
  int idx = 0; /* searchsorted(c, r, side='left') = #elements < r */
  for (int k = 0; k < 6; ++k) idx += (c[k] < r);
  if (idx > 5) idx = 5; /* a[ind] gather clamps */
  return idx;
}

#endif
