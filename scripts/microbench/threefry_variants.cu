// Microbenchmark (profiling helper, not part of the product): Threefry-2x32-20 throughput on sm_100a for
// three ways of writing the rotate-xor  x1 = rotl(x1, r) ^ x0:
//   V0  SHF.L.W (alu pipe) + LOP3 (alu pipe)                       — what nvcc emits for __funnelshift_l
//   V1  IMAD.WIDE.U32 by 2^r (fma pipe) + 3-input LOP3 lo^hi^x0    — multiplier read from the constant bank so that
//       ptxas cannot strength-reduce it back into a shift
//   V2  V0 and V1 alternating round by round (balances the two pipes)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a threefry_variants.cu -o tf && ./tf
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

struct Mults { uint32_t m[8]; uint32_t one; };  // 2^13, 2^15, 2^26, 2^6, 2^17, 2^29, 2^16, 2^24; one = 1 (from the constant bank)

__device__ __forceinline__ uint32_t rotx_shf(uint32_t x1, uint32_t x0, int r) { return __funnelshift_l(x1, x1, r) ^ x0; }
__device__ __forceinline__ uint32_t rotx_mul(uint32_t x1, uint32_t x0, uint32_t mult) {
  uint32_t lo, hi;
  asm("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x1), "r"(mult));
  return lo ^ hi ^ x0;
}

// V3: x0 += x1 as  mad.lo.u32 x0 = x1 * 1 + x0  with the 1 read from the constant bank (fma pipe) — the adds are a third of a round
__device__ __forceinline__ uint32_t add_mad(uint32_t a, uint32_t b, uint32_t one) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(one), "r"(a));
  return r;
}

template <int V>
__device__ __forceinline__ uint2 threefry(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, const Mults& M) {
  const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  uint32_t x0 = c0 + ks0, x1 = c1 + ks1;
  const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
#pragma unroll
  for (int g = 0; g < 5; ++g) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ri = (g & 1) * 4 + j;
      x0 = (V >= 3) ? add_mad(x0, x1, M.one) : x0 + x1;
      const bool mul = (V == 1) || (V == 4) || ((V == 2 || V == 5) && (j & 1)) || (V == 6 && j == 3);  // V3 keeps the shf rotate; V4/V5/V6 = V3 + mul rotates (all / every second / every fourth round)
      x1 = mul ? rotx_mul(x1, x0, M.m[ri]) : rotx_shf(x1, x0, R[ri]);
    }
    const uint32_t ka = (g % 3 == 0) ? ks1 : (g % 3 == 1) ? ks2 : ks0;
    const uint32_t kb = (g % 3 == 0) ? ks2 : (g % 3 == 1) ? ks0 : ks1;
    x0 += ka;
    x1 += kb + (uint32_t)(g + 1);
  }
  return make_uint2(x0, x1);
}

template <int V, int ILP>
__global__ void __launch_bounds__(256) k(const __grid_constant__ Mults M, uint32_t* out, int iters) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint2 s[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) s[i] = make_uint2(tid, i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) s[i] = threefry<V>(s[i].x, s[i].y, 0u, (uint32_t)it, M);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc ^= s[i].x ^ s[i].y;
  out[tid] = acc;
}

template <int V, int ILP>
void run(const char* name, int blocks_per_sm) {
  Mults M{{1u << 13, 1u << 15, 1u << 26, 1u << 6, 1u << 17, 1u << 29, 1u << 16, 1u << 24}, 1u};
  int sms = 148, iters = 2000, blocks = sms * blocks_per_sm;
  uint32_t* out;
  cudaMalloc(&out, (size_t)blocks * 256 * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<V, ILP><<<blocks, 256>>>(M, out, 10);
  cudaEventRecord(a);
  k<V, ILP><<<blocks, 256>>>(M, out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double n = (double)blocks * 256 * iters * ILP;
  uint32_t h; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM %3d ILP %d : %8.1f G threefry/s  (%.3f ms) chk %08x\n", name, blocks_per_sm * 8, ILP, n / ms / 1e6, ms, h);
  cudaFree(out);
}

int main() {
  // reference values: all variants must agree
  for (int bps : {1, 2, 8}) {
    run<0, 1>("V0 shf", bps); run<1, 1>("V1 mul.wide", bps); run<2, 1>("V2 alternate", bps);
    run<0, 2>("V0 shf", bps); run<1, 2>("V1 mul.wide", bps); run<2, 2>("V2 alternate", bps);
    run<0, 4>("V0 shf", bps); run<2, 4>("V2 alternate", bps);
    run<3, 1>("V3 add via mad", bps); run<3, 2>("V3 add via mad", bps); run<3, 4>("V3 add via mad", bps);
    run<4, 2>("V4 mad + mul all", bps); run<5, 2>("V5 mad + mul alternate", bps); run<6, 2>("V6 mad + mul every 4th", bps);
    run<5, 4>("V5 mad + mul alternate", bps); run<6, 4>("V6 mad + mul every 4th", bps);
  }
  return 0;
}
