#include <cstdio>
#include <cstdint>
namespace dogstep {}
#include "../../exploring-muzero-on-dog_b200/csrc/ttt_core.cuh"
#define main dogstep_unused_main
#include "../../exploring-muzero-on-dog_b200/csrc/mcts_kernels.cu"
#undef main
using namespace dogstep;
__global__ void k(unsigned long long* bad) {
  // every uniform the rollouts draw: bits_to_unit_float(b) for all 2^23 mantissas, clamped like the rollout; and the log of each
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (1u << 23)) return;
  const float tiny = 1.17549435e-38f;
  const float f = __uint_as_float(m | 0x3F800000u) - 1.0f;
  const float u = fmaxf(tiny, __fadd_rn(__fmul_rn(f, __fsub_rn(1.0f, tiny)), tiny));
  const float a = t_log(u), b = t_log_p(u);
  const float c = t_log(-a), d = t_log_p(-b);
  if (__float_as_uint(a) != __float_as_uint(b) || __float_as_uint(c) != __float_as_uint(d)) atomicAdd(bad, 1ull);
  // and the doubles themselves
  if (__double_as_longlong(log((double)u)) != __double_as_longlong(t_log_pos((double)u))) atomicAdd(bad + 1, 1ull);
}
// f_exp against (float)exp((double)x) for EVERY float bit pattern (NaNs compare as NaN == NaN)
__global__ void ke(unsigned long long* bad) {
  for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < (1ull << 32); b += (uint64_t)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((uint32_t)b);
    const float a = (float)exp((double)x), c = f_exp(x);
    const bool same = (a != a && c != c) || __float_as_uint(a) == __float_as_uint(c);
    if (!same) atomicAdd(bad + 2, 1ull);
  }
}
int main() {
  unsigned long long* bad; cudaMallocManaged(&bad, 32); bad[0] = bad[1] = bad[2] = 0;
  k<<<(1 << 23) / 256, 256>>>(bad); cudaDeviceSynchronize();
  ke<<<148 * 8, 256>>>(bad); cudaDeviceSynchronize();
  printf("log: float mismatches %llu  double mismatches %llu;  exp over all 2^32 floats: mismatches %llu\n", bad[0], bad[1], bad[2]);
  return bad[0] || bad[1] || bad[2];
}
