// madn_kernels.cu — CUDA kernels (sm_100a) + C-ABI for deterministic and classic MADN.
//
// Layout in HBM = the batched leaves of the reference pytrees (structure of arrays, game axis
// leading): board int8[n,56], current_player int8[n], pins int8[n,4,4], reward int8[n],
// done u8[n], action_set int8[n,4,6] / die int8[n], key u32[n,2].  One game per thread; a thread
// pulls its game into registers (bitboards + packed bytes, madn_core.cuh), applies the rules and
// writes the leaves back.  Kernels are HBM/launch bound (99 mutable bytes per game and step); the
// persistent play_random kernel keeps the game in registers across all lockstep iterations.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../include/dogstep.h"
#include "common.cuh"
#include "jaxrand.cuh"
#include "madn_core.cuh"
#include "madn_fast.cuh"
#include "madn_track.cuh"

namespace dogstep {

constexpr int kThreads = 128;

struct MadnPtrs {
  int8_t* board;
  int8_t* cur;
  int8_t* pins;
  int8_t* reward;
  uint8_t* done;
  int8_t* aset;  // det only
  int8_t* die;   // classic only
  uint32_t* key;
};

// ---- board bytes <-> bitboards ---------------------------------------------------------------
__device__ __forceinline__ void board_word_to_occ(uint32_t w, int c, uint64_t occ[4]) {
  const uint32_t L = 0x01010101u;
  uint32_t occm = (~w >> 7) & L, b0 = w & L, b1 = (w >> 1) & L;
  uint32_t e0 = occm & ~b0 & ~b1, e1 = occm & b0 & ~b1, e2 = occm & ~b0 & b1, e3 = occm & b0 & b1;
  occ[0] |= (uint64_t)((e0 * 0x01020408u) >> 24) << c;
  occ[1] |= (uint64_t)((e1 * 0x01020408u) >> 24) << c;
  occ[2] |= (uint64_t)((e2 * 0x01020408u) >> 24) << c;
  occ[3] |= (uint64_t)((e3 * 0x01020408u) >> 24) << c;
}

__device__ __forceinline__ uint32_t occ_to_board_word(const uint64_t occ[4], int c) {
  const uint32_t L = 0x01010101u, S = 0x00204081u;
  uint32_t e0 = (((uint32_t)(occ[0] >> c) & 0xFu) * S) & L;
  uint32_t e1 = (((uint32_t)(occ[1] >> c) & 0xFu) * S) & L;
  uint32_t e2 = (((uint32_t)(occ[2] >> c) & 0xFu) * S) & L;
  uint32_t e3 = (((uint32_t)(occ[3] >> c) & 0xFu) * S) & L;
  uint32_t any = e0 | e1 | e2 | e3;
  uint32_t val = (e1 | e3) | ((e2 | e3) << 1);
  return val | ((any ^ L) * 0xFFu);
}

template <bool DET>
__device__ __forceinline__ void load_state(const MadnGeom& g, const MadnPtrs& p, int64_t i, MadnRegs& s) {
  const uint32_t* bw = reinterpret_cast<const uint32_t*>(p.board + i * g.total);
  s.occ[0] = s.occ[1] = s.occ[2] = s.occ[3] = 0ull;
  const int nw = g.total >> 2;
  for (int w = 0; w < nw; ++w) board_word_to_occ(__ldg(bw + w), 4 * w, s.occ);
  if (g.n == 4) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p.pins) + i);
    s.pins[0] = v.x; s.pins[1] = v.y; s.pins[2] = v.z; s.pins[3] = v.w;
  } else {
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(p.pins) + i * g.n;
#pragma unroll
    for (int q = 0; q < 4; ++q) s.pins[q] = (q < g.n) ? __ldg(pw + q) : 0xFFFFFFFFu;
  }
  if (DET) {
    if (g.n == 4) {
      const uint64_t* aw = reinterpret_cast<const uint64_t*>(p.aset) + i * 3;
      uint64_t q0 = __ldg(aw), q1 = __ldg(aw + 1), q2 = __ldg(aw + 2);
      const uint64_t M = 0xFFFFFFFFFFFFull;
      s.as[0] = q0 & M;
      s.as[1] = ((q0 >> 48) | (q1 << 16)) & M;
      s.as[2] = ((q1 >> 32) | (q2 << 32)) & M;
      s.as[3] = q2 >> 16;
    } else {
      const uint16_t* aw = reinterpret_cast<const uint16_t*>(p.aset) + i * g.n * 3;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint64_t v = 0;
        if (q < g.n) v = (uint64_t)__ldg(aw + q * 3) | ((uint64_t)__ldg(aw + q * 3 + 1) << 16) | ((uint64_t)__ldg(aw + q * 3 + 2) << 32);
        s.as[q] = v;
      }
    }
    s.die = 0;
  } else {
    s.as[0] = s.as[1] = s.as[2] = s.as[3] = 0ull;
    s.die = (int)p.die[i];
  }
  s.cur = (int)p.cur[i];
  s.done = p.done[i] != 0;
  s.reward = (int)p.reward[i];
}

__device__ __forceinline__ void store_board(const MadnGeom& g, int8_t* board, int64_t i, const MadnRegs& s) {
  uint32_t* bw = reinterpret_cast<uint32_t*>(board + i * g.total);
  const int nw = g.total >> 2;
  for (int w = 0; w < nw; ++w) bw[w] = occ_to_board_word(s.occ, 4 * w);
}

__device__ __forceinline__ void store_pins(const MadnGeom& g, int8_t* pins, int64_t i, const MadnRegs& s) {
  if (g.n == 4) {
    reinterpret_cast<uint4*>(pins)[i] = make_uint4(s.pins[0], s.pins[1], s.pins[2], s.pins[3]);
  } else {
    uint32_t* pw = reinterpret_cast<uint32_t*>(pins) + i * g.n;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < g.n) pw[q] = s.pins[q];
  }
}

__device__ __forceinline__ void store_aset(const MadnGeom& g, int8_t* aset, int64_t i, const MadnRegs& s) {
  if (g.n == 4) {
    uint64_t* aw = reinterpret_cast<uint64_t*>(aset) + i * 3;
    aw[0] = s.as[0] | (s.as[1] << 48);
    aw[1] = (s.as[1] >> 16) | (s.as[2] << 32);
    aw[2] = (s.as[2] >> 32) | (s.as[3] << 16);
  } else {
    uint16_t* aw = reinterpret_cast<uint16_t*>(aset) + i * g.n * 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < g.n) {
        aw[q * 3] = (uint16_t)s.as[q];
        aw[q * 3 + 1] = (uint16_t)(s.as[q] >> 16);
        aw[q * 3 + 2] = (uint16_t)(s.as[q] >> 32);
      }
  }
}

// ---- kernels ----------------------------------------------------------------------------------
template <bool DET>
__global__ void __launch_bounds__(kThreads) k_madn_reset(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                         const int32_t* __restrict__ seeds, int starting_player) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  Key2 k0 = prng_key(seeds[i]);
  Key2 knew = split_i(k0, 0), sub = split_i(k0, 1);
  int sp = starting_player;
  if (sp < 0 || sp >= g.n) sp = randint_i(sub, 0, 0, g.n);
  MadnRegs s;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w = 0xFFFFFFFFu;
    if (q < g.n && DS_RULE(g, DOGSTEP_RULE_INITIAL_FREE_PIN)) w = 0xFFFFFF00u | (uint32_t)g.start[q];
    s.pins[q] = w;
    s.as[q] = 0x040404040404ull;
  }
  rebuild_occ(g, s);
  store_board(g, p.board, i, s);
  store_pins(g, p.pins, i, s);
  if (DET) store_aset(g, p.aset, i, s);
  else p.die[i] = 0;
  p.cur[i] = (int8_t)sp;
  p.reward[i] = 0;
  p.done[i] = 0;
  p.key[2 * i] = knew.a;
  p.key[2 * i + 1] = knew.b;
}

__global__ void __launch_bounds__(kThreads) k_madn_set_pins_on_board(const __grid_constant__ MadnGeom g,
                                                                     const int8_t* __restrict__ pins,
                                                                     int8_t* __restrict__ board, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  MadnRegs s;
  const uint32_t* pw = reinterpret_cast<const uint32_t*>(pins) + i * g.n;
#pragma unroll
  for (int q = 0; q < 4; ++q) s.pins[q] = (q < g.n) ? __ldg(pw + q) : 0xFFFFFFFFu;
  rebuild_occ(g, s);
  store_board(g, board, i, s);
}

template <bool DET>
__global__ void __launch_bounds__(kThreads) k_madn_valid_action(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                uint8_t* __restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  MadnRegs s;
  load_state<DET>(g, p, i, s);
  if (DET) {
    uint32_t m = madn_det_valid_mask(g, s);
    uint64_t* out = reinterpret_cast<uint64_t*>(mask) + i * 3;  // 24 bytes, 8-aligned
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      uint32_t bits = (m >> (8 * w)) & 0xFFu;
      // spread 8 bits to 8 bytes
      uint64_t lo = ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
      uint64_t hi = (((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
      out[w] = lo | (hi << 32);
    }
  } else {
    uint32_t m = madn_cls_valid_mask(g, s);
    reinterpret_cast<uint32_t*>(mask)[i] = ((m & 0xFu) * 0x00204081u) & 0x01010101u;
  }
}

template <bool DET>
__global__ void __launch_bounds__(kThreads) k_madn_step(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                        const int8_t* __restrict__ action, int8_t* __restrict__ reward,
                                                        uint8_t* __restrict__ done) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  MadnRegs s;
  load_state<DET>(g, p, i, s);
  if (DET) {
    uint32_t m = madn_det_valid_mask(g, s);
    madn_det_step(g, s, (int)action[2 * i], (int)action[2 * i + 1], m);
    store_aset(g, p.aset, i, s);
  } else {
    uint32_t m = madn_cls_valid_mask(g, s);
    madn_cls_step(g, s, (int)action[i], m);
  }
  store_board(g, p.board, i, s);
  store_pins(g, p.pins, i, s);
  p.cur[i] = (int8_t)s.cur;
  p.reward[i] = (int8_t)s.reward;
  p.done[i] = (uint8_t)s.done;
  if (reward) reward[i] = (int8_t)s.reward;
  if (done) done[i] = (uint8_t)s.done;
}

template <bool DET>
__global__ void __launch_bounds__(kThreads) k_madn_no_step(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                           int8_t* __restrict__ reward, uint8_t* __restrict__ done) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  int pid = (int)p.cur[i];
  if (DET && pid >= -g.n && pid < g.n) {
    int row = gidx(pid, g.n);
    uint16_t* aw = reinterpret_cast<uint16_t*>(p.aset) + (i * g.n + row) * 3;
    aw[0] = aw[1] = aw[2] = 0x0404u;
  }
  p.cur[i] = (int8_t)floormod(pid + 1, g.n);
  if (reward) reward[i] = 0;
  if (done) done[i] = p.done[i];
}

// encode_board: one thread per 4 output cells.
template <bool DET>
__global__ void __launch_bounds__(256) k_madn_encode_board(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                           int8_t* __restrict__ obs) {
  const int C = DET ? 8 * g.n + 2 : 2 * g.n + 3;
  const int wpr = g.total >> 2;  // words per channel row
  const int64_t wpg = (int64_t)C * wpr;
  int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n * wpg) return;
  int64_t i = e / wpg;
  int r = (int)(e - i * wpg);
  int c = r / wpr, k0 = (r - c * wpr) * 4;
  const int8_t* board = p.board + i * g.total;
  const int cur = (int)p.cur[i];
  uint32_t out = 0;
  if (c < g.n + 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + j;
      int src = (k < g.bs) ? floormod(k + g.d * cur, g.bs) : g.bs + floormod(k - g.bs + 4 * cur, 16);
      int v = (int)board[src];
      int rel = (v < 0) ? -1 : floormod(v - cur, g.n);  // channel index of the owner in the mover's frame
      int val;
      if (c < g.n) val = rel == c;
      else if (DS_RULE(g, DOGSTEP_RULE_TEAMS)) val = (c == g.n) ? (rel >= 0 && (rel & 1) == 0) : (rel >= 0 && (rel & 1) == 1);
      else val = (c == g.n) ? (rel == 0) : (rel >= 1);
      out |= (uint32_t)val << (8 * j);
    }
  } else if (c < 2 * g.n + 2) {
    int src = floormod(c - g.n - 2 + cur, g.n);
    uint32_t pw = reinterpret_cast<const uint32_t*>(p.pins)[i * g.n + src];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) cnt += (byte_s(pw, j) == -1);
    out = (uint32_t)cnt * 0x01010101u;
  } else if (DET) {
    int a = c - 2 * g.n - 2;
    int src = floormod(a / 6 + cur, g.n);
    int v = (int)p.aset[(i * g.n + src) * 6 + a % 6];
    out = (uint32_t)(v & 0xFF) * 0x01010101u;
  } else {
    out = (uint32_t)((int)p.die[i] & 0xFF) * 0x01010101u;
  }
  reinterpret_cast<uint32_t*>(obs)[e] = out;
}

// categorical(key, where(mask, 0, -1e9)) == first valid action with the largest 23-bit uniform
// mantissa (gumbel = -log(-log(u)) is strictly increasing in u; see DESIGN.md "random policy").
// Four legal actions per pass: the four Threefry evaluations are independent instruction streams, so a thread that owns a whole
// game (the per-call kernels: 14 warps per SM, bound by the latency of the longest chain of any warp) waits for ceil(legal / 4)
// evaluations instead of `legal` of them.  Actions are visited in ascending order, a later one wins only with a strictly larger
// mantissa: the first maximum, as jnp.argmax.
__device__ __forceinline__ int categorical_masked(Key2 key, uint32_t mask) {
  int best = -1;
  uint32_t bm = 0;
  for (uint32_t m = mask; m;) {
    int a[4];
    bool v[4];
    uint32_t mant[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = m != 0u;
      a[k] = v[k] ? __ffs(m) - 1 : 0;
      m &= m - 1;  // 0 stays 0
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) mant[k] = bits_i(key, (uint32_t)a[k]) >> 9;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (v[k] && (best < 0 || mant[k] > bm)) { best = a[k]; bm = mant[k]; }
  }
  return best;
}

__device__ __forceinline__ void madn_det_random_turn(const MadnGeom& g, MadnRegs& s, Key2 key) {
  uint32_t m = madn_det_valid_mask(g, s);
  if (m) {
    int a = categorical_masked(key, m);
    madn_det_step(g, s, a / 6, a % 6 + 1, m);  // map_action (deterministic_madn.py:469-479)
  } else {
    madn_det_no_step(g, s);
  }
}

__device__ __forceinline__ void store_det_all(const MadnGeom& g, const MadnPtrs& p, int64_t i, const MadnRegs& s) {
  store_board(g, p.board, i, s);
  store_pins(g, p.pins, i, s);
  store_aset(g, p.aset, i, s);
  p.cur[i] = (int8_t)s.cur;
  p.reward[i] = (int8_t)s.reward;
  p.done[i] = (uint8_t)s.done;
}

// One lockstep iteration per launch.  A thread owns a game (load, legal mask, move, store); the Threefry draws — one per legal
// action, 40 % of the call's instructions when every lane draws for its own game, because a warp then runs as many passes as its
// busiest lane has legal actions — are pooled per warp: the (game, action) pairs go to a shared list and are dealt out evenly to
// the 32 lanes, the winner of a game is a shared-memory atomicMax on (mantissa << 5 | 23 - action): largest 23-bit mantissa, lowest
// action on ties (jax.random.categorical's choice, as in the persistent kernel).
__global__ void __launch_bounds__(kThreads) k_madn_det_random_step(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                   Key2 rng, int64_t game_offset,
                                                                   unsigned long long* __restrict__ active_count) {
  __shared__ uint16_t s_items[kThreads / 32][32 * 24];
  __shared__ uint32_t s_best[kThreads / 32][32];
  const uint32_t FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const bool alive = i < n && p.done[i] == 0;
  MadnRegs s;
  Key2 key{0u, 0u};
  uint32_t m = 0u;
  int cp = 0;
  bool fast = false;
  const RuleSet<kRulesRuntime> R{g.rules};
  if (alive) {
    load_state<true>(g, p, i, s);
    key = split_i(rng, (uint32_t)(game_offset + i + 1));
    // 4 players, distance 10, board consistent with the pins: the branch-free bit rows of madn_fast.cuh (same results as the
    // generic rules on every such state, tests/test_madn_fast_core.py)
    fast = g.n == 4 && g.d == 10 && is_canonical4(s, s.occ);
    m = fast ? det_valid_mask4(R, g, s, cp) : madn_det_valid_mask(g, s);
  }
  s_best[warp][lane] = 0u;
  const int cnt = __popc(m);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(FULL, incl, 31);
  int off = incl - cnt;
  for (uint32_t mm = m; mm; mm &= mm - 1) s_items[warp][off++] = (uint16_t)((lane << 8) | (__ffs(mm) - 1));
  __syncwarp();
  for (int base = 0; base < total; base += 64) {  // two independent Threefry chains per lane and pass
    const int j0 = base + lane, j1 = j0 + 32;
    const int it0 = (j0 < total) ? (int)s_items[warp][j0] : 0, it1 = (j1 < total) ? (int)s_items[warp][j1] : 0;
    const Key2 k0{__shfl_sync(FULL, key.a, it0 >> 8), __shfl_sync(FULL, key.b, it0 >> 8)};
    const Key2 k1{__shfl_sync(FULL, key.a, it1 >> 8), __shfl_sync(FULL, key.b, it1 >> 8)};
    const uint32_t v0 = bits_i(k0, (uint32_t)(it0 & 0xFF)) >> 9, v1 = bits_i(k1, (uint32_t)(it1 & 0xFF)) >> 9;
    if (j0 < total) atomicMax(&s_best[warp][it0 >> 8], (v0 << 5) | (uint32_t)(23 - (it0 & 0xFF)));
    if (j1 < total) atomicMax(&s_best[warp][it1 >> 8], (v1 << 5) | (uint32_t)(23 - (it1 & 0xFF)));
  }
  __syncwarp();
  if (alive) {
    if (m) {
      const int a = 23 - (int)(s_best[warp][lane] & 31u);
      if (fast) det_step4(R, s, cp, a);
      else madn_det_step(g, s, a / 6, a % 6 + 1, m);  // map_action (deterministic_madn.py:469-479)
    } else if (fast) {
      det_no_step4(s);
    } else {
      madn_det_no_step(g, s);
    }
    store_det_all(g, p, i, s);
  }
  if (active_count) {
    const unsigned b = __ballot_sync(FULL, alive);
    if (lane == 0 && b) atomicAdd(active_count, (unsigned long long)__popc(b));
  }
}

// ---- evaluation loop: one lockstep iteration of play_eval_loop_jitted (MuZero_det_MADN/evaluate_agent.py:733-930) -------
// Every seat is played by an agent type (params['type']): 3 = random legal policy, 2 = the rule-based scorer, anything
// else = tree search (the caller's search supplies the action).  Thread per game.
__device__ __forceinline__ float eval_log_f(float x) { return (float)log((double)x); }

// do_rule_based (:780-878), literally — including that it scores env.pins[env.current_player] (NOT the team proxy the
// legal mask was computed for), that the candidate distances are arange(6) = 0..5 (one less than the move an action
// plays), and that base_score = repeat(action_abundance, 4) indexes the six abundances by a // 4.
__device__ int madn_rule_based_action(const MadnGeom& g, const MadnRegs& s, uint32_t m, Key2 key) {
  const int cur = s.cur;
  const uint32_t pw = pick4(s.pins, cur);
  const int start = g.start[cur], target = g.target[cur], goal0 = g.goal0[cur];
  const int mts = DS_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START) ? 1 : 0;
  const int mate = DS_RULE(g, DOGSTEP_RULE_TEAMS) ? ((cur + 2) & 3) : -1;
  int pins_in_home = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) pins_in_home += byte_s(pw, p) < 0;
  const float out_w = pins_in_home >= 2 ? 3.0f : 2.0f;
  const float denom = fmaxf((float)__popc(m), 1.0f);
  float abundance[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int c = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) c += (m >> (6 * p + k)) & 1u;
    abundance[k] = __fdiv_rn((float)c, denom);
  }
  float best = 0.0f;
  int best_a = -1;
  // Only the legal actions are scored: an illegal one has logit -inf, so its gumbel + logit is -inf and it can only be the
  // argmax if nothing is legal — and the caller comes here with m != 0.  (A Threefry pass and two double-precision logs per
  // action: the loop is the latency of the evaluation step.)
#pragma unroll 1
  for (uint32_t mm = m; mm; mm &= mm - 1) {
    const int a = __ffs(mm) - 1;
    const int p = a / 6, k = a - 6 * p;
    const int cur_pos = byte_s(pw, p);
    const int moved = cur_pos + k, fitted = floormod(moved, g.bs);
    const int x = moved - target - mts;
    int new_pos = fitted;
    if (x <= 4 && x > 0 && cur_pos <= target) new_pos = goal0 + x - 1;
    if (cur_pos >= g.bs) new_pos = moved;
    if (cur_pos < 0) new_pos = start;
    const bool into_goal = (unsigned)(new_pos - goal0) <= 3u && cur_pos < g.bs;
    const bool leaves_home = cur_pos < 0 && new_pos == start;
    bool hits = false;
    for (int q = 0; q < g.n; ++q) {
      if (q == cur || q == mate) continue;
      const uint32_t ow = pick4(s.pins, q);
#pragma unroll
      for (int j = 0; j < 4; ++j) hits = hits || byte_s(ow, j) == new_pos;
    }
    hits = hits && new_pos != cur_pos;
    float score = abundance[a >> 2];
    score = __fadd_rn(score, into_goal ? 5.0f : 0.0f);
    score = __fadd_rn(score, leaves_home ? out_w : 0.0f);
    score = __fadd_rn(score, hits ? 2.0f : 0.0f);
    const float logit = __fdiv_rn(score, 0.25f);
    const float u = uniform_i(key, (uint32_t)a, 1.17549435e-38f, 1.0f);
    const float v = __fadd_rn(-eval_log_f(-eval_log_f(u)), logit);  // jax.random.categorical: argmax(gumbel + logits)
    if (best_a < 0 || v > best) { best = v; best_a = a; }
  }
  return best_a;
}

// do_rule_based of the dice game (MuZero_Classic_MADN/evaluate_agent_stochastic.py:782-872): four actions (the pin to move by
// env.die), scores = goal bonus 5 + leaving-home bonus 3 / 2 + hit bonus 2.5, categorical over score / 0.25
__device__ int madn_cls_rule_based_action(const MadnGeom& g, const MadnRegs& s, uint32_t m, Key2 key) {
  const int cur = s.cur;
  const uint32_t pw = pick4(s.pins, cur);
  const int start = g.start[cur], target = g.target[cur], goal0 = g.goal0[cur];
  const int mts = DS_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START) ? 1 : 0;
  const int mate = DS_RULE(g, DOGSTEP_RULE_TEAMS) ? ((cur + 2) & 3) : -1;
  int pins_in_home = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) pins_in_home += byte_s(pw, p) < 0;
  const float out_w = pins_in_home >= 2 ? 3.0f : 2.0f;
  float best = 0.0f;
  int best_a = -1;
#pragma unroll 1
  for (int a = 0; a < 4; ++a) {
    const int cur_pos = byte_s(pw, a);
    const int moved = cur_pos + s.die, fitted = floormod(moved, g.bs);
    const int x = moved - target - mts;
    int new_pos = fitted;
    if (x <= 4 && x > 0 && cur_pos <= target) new_pos = goal0 + x - 1;
    if (cur_pos >= g.bs) new_pos = moved;
    if (cur_pos < 0) new_pos = start;
    const bool into_goal = (unsigned)(new_pos - goal0) <= 3u && cur_pos < g.bs;
    const bool leaves_home = cur_pos < 0 && new_pos == start;
    bool hits = false;
    for (int q = 0; q < g.n; ++q) {
      if (q == cur || q == mate) continue;
      const uint32_t ow = pick4(s.pins, q);
#pragma unroll
      for (int j = 0; j < 4; ++j) hits = hits || byte_s(ow, j) == new_pos;
    }
    hits = hits && new_pos != cur_pos;
    float score = __fadd_rn(0.0f, into_goal ? 5.0f : 0.0f);
    score = __fadd_rn(score, leaves_home ? out_w : 0.0f);
    score = __fadd_rn(score, hits ? 2.5f : 0.0f);
    const float logit = ((m >> a) & 1u) ? __fdiv_rn(score, 0.25f) : __int_as_float(0xFF800000);
    const float u = uniform_i(key, (uint32_t)a, 1.17549435e-38f, 1.0f);
    const float v = __fadd_rn(-eval_log_f(-eval_log_f(u)), logit);
    if (best_a < 0 || v > best) { best = v; best_a = a; }
  }
  return best_a;
}

template <bool DET>
__global__ void __launch_bounds__(kThreads) k_madn_eval_step(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                 int4 agent_type, const int32_t* __restrict__ search_action,
                                                                 Key2 rng, int64_t game_offset, int32_t* __restrict__ winners,
                                                                 unsigned long long* __restrict__ active_count) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int active = 0;
  if (i < n && p.done[i] == 0) {
    MadnRegs s;
    load_state<DET>(g, p, i, s);
    const Key2 key = split_i(rng, (uint32_t)(game_offset + i + 1));  // rng_key, *step_keys = split(rng_key, num_envs + 1)
    const uint32_t m = DET ? madn_det_valid_mask(g, s) : madn_cls_valid_mask(g, s);
    if (m) {
      const int cur = s.cur;
      const int type = cur == 0 ? agent_type.x : cur == 1 ? agent_type.y : cur == 2 ? agent_type.z : agent_type.w;
      int a;
      if (type == 3) a = categorical_masked(key, m);
      else if (type == 2) a = DET ? madn_rule_based_action(g, s, m, key) : madn_cls_rule_based_action(g, s, m, key);
      else a = search_action[i];
      if (DET) madn_det_step(g, s, a / 6, a % 6 + 1, m);  // map_action
      else madn_cls_step(g, s, a, m);
    } else {
      if (DET) madn_det_no_step(g, s);
      else s.cur = (int)(int8_t)floormod(s.cur + 1, g.n);  // no_step of the dice game (classic_madn.py:353-365): the turn passes
    }
    if (DET) {
      store_det_all(g, p, i, s);
    } else {
      store_board(g, p.board, i, s);
      store_pins(g, p.pins, i, s);
      p.cur[i] = (int8_t)s.cur;
      p.reward[i] = (int8_t)s.reward;
      p.done[i] = (uint8_t)s.done;
    }
    if (s.done && winners) {  // manual_get_winner (:16-45) on the new board
      const uint64_t any = s.occ[0] | s.occ[1] | s.occ[2] | s.occ[3];
      int w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = player_done(g, any, q);
      if (DS_RULE(g, DOGSTEP_RULE_TEAMS)) {
        const int t0 = w[0] & w[2], t1 = w[1] & w[3];
        const int none = (t0 & t1) | !(t0 | t1);
        w[0] = w[2] = !none && t0;
        w[1] = w[3] = !none && !t0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) winners[4 * i + q] += w[q];
    }
    active = 1;
  }
  if (active_count) {
    const unsigned b = __ballot_sync(0xFFFFFFFFu, active);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(active_count, (unsigned long long)__popc(b));
  }
}

// Persistent lockstep loop.  One game per lane, but the expensive part of a turn — one
// Threefry-2x32-20 per LEGAL action for the categorical draw — is pooled per warp: the (game, action)
// pairs of all 32 games are compacted into a shared list and dealt out evenly to the lanes, so a game
// with 14 legal actions does not stall 31 lanes that have 3.  Results come back through a
// [32 games x 24 actions] mantissa table in shared memory.
constexpr int kPlayWarps = kThreads / 32;

// One warp, one game per lane, to the end (generic rules: any player count / distance / state).  `items` [32 * 24] and
// `mant` [32 * 24] are this warp's shared work arrays.  Also the fallback of the specialised kernel below for CTAs that hold
// a state outside the specialised rules' domain — a separate function with its own register allocation.
__device__ __noinline__ void play_random_warp(const MadnGeom& g, const MadnPtrs& p, int64_t i, bool have, MadnRegs& s, Key2 rng,
                                              uint32_t my, int max_steps, int32_t* __restrict__ game_len,
                                              unsigned long long* __restrict__ total_steps, uint16_t* items, uint32_t* mant) {
  const int lane = threadIdx.x & 31;
  const uint32_t FULL = 0xFFFFFFFFu;
  int len = 0;
  bool alive = have && !s.done;
  for (int t = 0; t < max_steps; ++t) {
    if (!__any_sync(FULL, alive)) break;
    const Key2 key = split_i(rng, my);  // split(rng, N+1)[j+1]
    rng = split_i(rng, 0u);             // split(rng, N+1)[0]
    const uint32_t m = alive ? madn_det_valid_mask(g, s) : 0u;
    // warp-wide compaction of (lane, action) pairs
    const int cnt = __popc(m);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    int off = incl - cnt;
    for (uint32_t mm = m; mm; mm &= mm - 1) items[off++] = (uint16_t)((lane << 8) | (__ffs(mm) - 1));
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
      const int j = base + lane;
      const int it = (j < total) ? (int)items[j] : 0;
      const int o = it >> 8, a = it & 0xFF;
      Key2 k{__shfl_sync(FULL, key.a, o), __shfl_sync(FULL, key.b, o)};
      const uint32_t v = bits_i(k, (uint32_t)a) >> 9;
      if (j < total) mant[o * 24 + a] = v;
    }
    __syncwarp();
    if (alive) {
      if (m) {
        int best = -1;
        uint32_t bm = 0;
        for (uint32_t mm = m; mm; mm &= mm - 1) {
          const int a = __ffs(mm) - 1;
          const uint32_t v = mant[lane * 24 + a];
          if (best < 0 || v > bm) { best = a; bm = v; }
        }
        madn_det_step(g, s, best / 6, best % 6 + 1, m);  // map_action (deterministic_madn.py:469-479)
      } else {
        madn_det_no_step(g, s);
      }
      ++len;
      alive = !s.done;
    }
    __syncwarp();
  }
  if (have) {
    store_det_all(g, p, i, s);
    if (game_len) game_len[i] = len;
  }
  if (total_steps) {
    unsigned v = (unsigned)len;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0 && v) atomicAdd(total_steps, (unsigned long long)v);
  }
}

__global__ void __launch_bounds__(kThreads) k_madn_det_play_random(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                   Key2 rng0, int64_t game_offset, int max_steps,
                                                                   int32_t* __restrict__ game_len,
                                                                   unsigned long long* __restrict__ total_steps) {
  __shared__ uint16_t s_items[kPlayWarps][32 * 24];
  __shared__ uint32_t s_mant[kPlayWarps][32 * 24];
  const int warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  MadnRegs s;
  if (i < n) load_state<true>(g, p, i, s);
  play_random_warp(g, p, i, i < n, s, rng0, (uint32_t)(game_offset + i + 1), max_steps, game_len, total_steps, s_items[warp],
                   s_mant[warp]);
}

// Persistent lockstep loop for 4 players, distance 10 — every configuration of the reference (the generic kernel above
// covers other geometries).  One CTA per SM (up to 512 games), one game per lane held in registers (PlayState: the track
// state of madn_track.cuh for the training rule dict, the bitboard registers of madn_fast.cuh for a run-time rule mask).
// Per iteration:
//   1. each lane takes its step key and derives the 24-bit legal mask (branch-free bit rows);
//   2. the (game, action) pairs of the warp are compacted into a shared list and dealt out evenly to the 32 lanes,
//      two per lane and pass, so the Threefry calls — one per LEGAL action, half of all instructions — are
//      load-balanced; the first pass also derives the NEXT iteration's step key (a third independent chain);
//   3. the argmax of the categorical draw is a shared-memory atomicMax per game on (mantissa << 5 | 23 - action):
//      largest 23-bit mantissa first, lowest action index on ties — jax.random.categorical's choice;
//   4. each lane applies its game's move; a finished game waits in its lane for the next compaction point, where the finished
//      games of a warp are written back together.
// Games end at different plies (mean 403, max ~850 in config 2): every kPlayRound iterations the live games of the CTA are
// packed through shared memory (PlayState::kWords + 3 words per game) into the fewest warps that give all four warp schedulers
// the same number of games (PlayShare), and the emptied warps only wait at the barriers.  The loop key chain
// (rng <- split(rng, N + 1)[0], the same for every game) is produced a round ahead by one extra warp per CTA into a
// double-buffered shared ring.  Once a CTA is down to kPEnter games it switches to the draw-ahead mode (see the kernel).
// A CTA that loaded a game outside the specialised rules' domain (madn_fast.cuh / madn_track.cuh) plays all its games with the
// generic rules of madn_core.cuh (play_random_warp).
#ifdef DOGSTEP_TRACE
__device__ unsigned long long g_play_trace[128];
__device__ int g_play_trace_block;
#endif
constexpr int kPlayRound = 48;  // 16 / 24 / 32 / 40 / 48 / 64: 1.81 / 1.78 / 1.77 / 1.76 / 1.75 / 1.76 ms on config 2 (finished games wait for the next compaction point)
constexpr int kPlayMaxThreads = 512;  // game threads per CTA (+ 32: the producer warp)
constexpr int kXWords = 23;  // cur|reward, len, game index + PlayState::kWords (at most 20: occ 8, pins 4, action set 8)

constexpr int kRingMax = 64;  // longest round the key ring holds
// draw-ahead mode (see the kernel): at most kPSlots live games on four warps, the other warps draw kPChunk iterations ahead
constexpr int kPEnter = 64;   // live games per CTA at which the mode is entered (<= kPSlots)
constexpr int kPSlots = 128, kPChunk = 4, kPStride = 28;  // 24 draws per (iteration, game) row, padded: conflict-free 16-byte reads
// `threads` = game threads (the CTA has one more warp, the key-chain producer)
static size_t play_smem_bytes(int threads) {
  return (size_t)(threads / 32) * (24 * 32 * 2 + 32 * 4 + 32 * 8) + (size_t)kXWords * threads * 4 + 2 * 34 * 4 +
         2 * (kRingMax + 2) * 8 + (size_t)2 * kPChunk * kPSlots * (kPStride * 4 + 8) + kPSlots * 4;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// The four warp schedulers of an SM take the warps of a CTA round robin (warp w -> scheduler w % 4) and every one of them is
// bound by its own integer pipe, so a round lasts as long as the scheduler with the most games needs: L live games go to the
// fewest warps that can hold them with the SAME number on every scheduler (4 * ceil(L / 128) warps), spread evenly — 443
// games are 16 warps of 27-28 instead of 13 of 32 and one of 27 (schedulers 0 and 1 then carried 128 games against 96 and
// 91).  The odd games go to the highest warps: scheduler 0 also runs the key-chain producer.
struct PlayShare {
  int wa, base, first_big;  // warps in use, games per warp, first warp that holds base + 1
  __device__ PlayShare(int L, int W) {
    wa = min(4 * ((L + 127) >> 7), W);
    base = L / wa;
    first_big = wa - (L - base * wa);
  }
  __device__ int count(int w) const { return w < wa ? base + (w >= first_big) : 0; }
  __device__ int start(int w) const { return w * base + max(0, w - first_big); }
};

// What the persistent kernel keeps of a game, per rule program.  Run-time rule mask: the bitboard registers and branch-free
// rules of madn_fast.cuh (23 words).  The training rule dict: the track state of madn_track.cuh (8 words, no board).
template <uint32_t CT>
struct PlayState {
  static constexpr int kWords = 20;
  MadnRegs s;
  __device__ bool from_regs(const MadnRegs& r) { s = r; return is_canonical4(r, r.occ); }
  __device__ void store(const MadnGeom& g, const MadnPtrs& p, int64_t i) const { store_det_all(g, p, i, s); }
  __device__ bool done() const { return s.done != 0; }
  __device__ uint32_t mask(const RuleSet<CT> R, const MadnGeom& g, int& cp) const { return det_valid_mask4(R, g, s, cp); }
  __device__ void step(const RuleSet<CT> R, int cp, int a) { det_step4(R, s, cp, a); }
  __device__ void no_step() { det_no_step4(s); }
  __device__ void pack(uint32_t* x, int T) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      x[(2 * q) * T] = (uint32_t)s.occ[q];
      x[(2 * q + 1) * T] = (uint32_t)(s.occ[q] >> 32);
      x[(8 + q) * T] = s.pins[q];
      x[(12 + 2 * q) * T] = (uint32_t)s.as[q];
      x[(13 + 2 * q) * T] = (uint32_t)(s.as[q] >> 32);
    }
  }
  __device__ void unpack(const uint32_t* x, int T) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s.occ[q] = (uint64_t)x[(2 * q) * T] | ((uint64_t)x[(2 * q + 1) * T] << 32);
      s.pins[q] = x[(8 + q) * T];
      s.as[q] = (uint64_t)x[(12 + 2 * q) * T] | ((uint64_t)x[(13 + 2 * q) * T] << 32);
    }
    s.done = 0;
    s.die = 0;
  }
  __device__ int& cur() { return s.cur; }
  __device__ int& reward() { return s.reward; }
};

template <>
struct PlayState<kTrainRules> {
  static constexpr int kWords = 8;
  Track4 s;
  __device__ bool from_regs(const MadnRegs& r) { return track_from_regs(r, s); }
  // every leaf straight from the track state: board = set_pins_on_board(pins) (deterministic_madn.py:259-271) is -1 everywhere,
  // then one byte per pin (a thread's stores to one address keep their order; canonical: no cell is shared by two players)
  __device__ void store(const MadnGeom&, const MadnPtrs& p, int64_t i) const {
    int8_t* board = p.board + i * 56;
    uint2* bw = reinterpret_cast<uint2*>(board);
#pragma unroll
    for (int w = 0; w < 7; ++w) bw[w] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    uint32_t pw[4];
    uint64_t as[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int pos = track_to_cell(q, byte_s(s.pins[q], k));
        if (pos >= 0) board[pos] = (int8_t)q;
        w |= (uint32_t)(pos & 0xFF) << (8 * k);
      }
      pw[q] = w;
      as[q] = track_as_bytes(s.as[q]);
    }
    reinterpret_cast<uint4*>(p.pins)[i] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
    uint64_t* aw = reinterpret_cast<uint64_t*>(p.aset) + i * 3;
    aw[0] = as[0] | (as[1] << 48);
    aw[1] = (as[1] >> 16) | (as[2] << 32);
    aw[2] = (as[2] >> 32) | (as[3] << 16);
    p.cur[i] = (int8_t)s.cur;
    p.reward[i] = (int8_t)s.reward;
    p.done[i] = (uint8_t)s.done;
  }
  __device__ bool done() const { return s.done != 0; }
  __device__ uint32_t mask(const RuleSet<kTrainRules>, const MadnGeom&, int& cp) const { return track_valid_mask(s, cp); }
  __device__ void step(const RuleSet<kTrainRules>, int cp, int a) { track_step(s, cp, a); }
  __device__ void no_step() { track_no_step(s); }
  __device__ void pack(uint32_t* x, int T) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      x[q * T] = s.pins[q];
      x[(4 + q) * T] = s.as[q];
    }
  }
  __device__ void unpack(const uint32_t* x, int T) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s.pins[q] = x[q * T];
      s.as[q] = x[(4 + q) * T];
    }
    s.done = 0;
  }
  __device__ int& cur() { return s.cur; }
  __device__ int& reward() { return s.reward; }
};

template <uint32_t CT>
__global__ void __launch_bounds__(kPlayMaxThreads + 32) k_madn_det_play_cta(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                      Key2 rng0, int64_t game_offset, int max_steps,
                                                                      int32_t* __restrict__ game_len,
                                                                      unsigned long long* __restrict__ total_steps, int round_len,
                                                                      int games_per_cta, int p_enter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x - 32, W = T >> 5;  // game threads / warps; the CTA's last warp produces the key chain
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool producer = warp == W;
  const int pw = producer ? 0 : warp;  // the producer never touches the per-warp arrays
  uint2* s_key = reinterpret_cast<uint2*>(smem_raw) + pw * 32;                                  // [W][32]
  uint32_t* s_best = reinterpret_cast<uint32_t*>(smem_raw + (size_t)W * 256) + pw * 32;         // [W][32]
  uint16_t* s_items = reinterpret_cast<uint16_t*>(smem_raw + (size_t)W * 384) + pw * (24 * 32);  // [W][768]
  uint32_t* s_x = reinterpret_cast<uint32_t*>(smem_raw + (size_t)W * 1920);                     // [kXWords][T]
  uint32_t* s_cnt = s_x + (size_t)kXWords * T;                                                  // 2 x [34] counts
  uint2* s_ring = reinterpret_cast<uint2*>(s_cnt + 2 * 34);                                     // 2 x [kRingMax + 2] keys
  uint32_t* s_draw = reinterpret_cast<uint32_t*>(s_ring + 2 * (kRingMax + 2));                  // [2][kPChunk][kPSlots][kPStride]
  uint2* s_pkey = reinterpret_cast<uint2*>(s_draw + 2 * kPChunk * kPSlots * kPStride);          // [2][kPSlots * kPChunk] step keys
  uint32_t* s_pmy = reinterpret_cast<uint32_t*>(s_pkey + 2 * kPChunk * kPSlots);                // [kPSlots] split index of the game
  const uint32_t FULL = 0xFFFFFFFFu;
  const int64_t cta_base = (int64_t)blockIdx.x * games_per_cta;  // games_per_cta <= T: the games are spread over ALL SMs
  const int games_here = (int)max((int64_t)0, min((int64_t)games_per_cta, n - cta_base));
  int gi = 0;  // game held by this lane, relative to cta_base
  const RuleSet<CT> R{g.rules};
  int len = 0;
  unsigned steps_done = 0;
  PlayState<CT> s;
  MadnRegs r0;  // as loaded (only live until the rule program is chosen)
  bool alive = false, canon = true, have = false;
  bool pending = false;  // finished, not written back yet: the finished games of a warp are stored together at the next
                         // compaction point (one lane at a time the store path was 3 % of all issued instructions)
  if (producer) {
    // The loop key chain rng_{t+1} = split(rng_t, N + 1)[0] (game_agent.py:60) is the same for every game: one warp per CTA
    // computes it, a round ahead, into a double-buffered ring (ring[i] = rng_{t0 + i}, i = 0..round_len + 1) instead of
    // every lane of every warp repeating it each iteration (it was one Threefry in seven, and sat on the critical path
    // of the next step key).
    Key2 r = rng0;
    if (lane == 0) s_ring[0] = make_uint2(r.a, r.b);
    const int need = min(round_len, max_steps);  // a launch of a few iterations (random_steps) does not pay for a whole round
    for (int i = 1; i <= need; ++i) {
      r = split_i(r, 0u);
      if (lane == 0) s_ring[i] = make_uint2(r.a, r.b);
    }
  } else if (games_here > 0 && lane < PlayShare(games_here, W).count(warp)) {
    gi = PlayShare(games_here, W).start(warp) + lane;
    have = true;
    load_state<true>(g, p, cta_base + gi, r0);
    alive = !r0.done;
    canon = s.from_regs(r0) || !alive;  // a finished game is left as it is, whatever it holds
  }
  if (__syncthreads_or(!canon)) {
    // some game of this CTA is outside the specialised rules' domain (see madn_fast.cuh / madn_track.cuh): the whole CTA plays
    // with the generic rules, warp by warp; same keys, same results
    if (!producer) {
      MadnRegs tmp = r0;  // the callee takes it by reference: a copy, so that r0 itself can stay in registers
      play_random_warp(g, p, cta_base + gi, have, tmp, rng0, (uint32_t)(game_offset + cta_base + gi + 1), max_steps, game_len,
                       total_steps, s_items, s_draw + (size_t)warp * (32 * 24));
    }
    return;
  }
  if (have && !alive && game_len) game_len[cta_base + gi] = 0;
  const uint32_t lane_hi = (uint32_t)lane << 8;
  if (!producer) s_best[lane] = 0u;
  int t = 0, round = 0;
#ifdef DOGSTEP_TRACE
  long long tr_last = 0;
#endif
  bool pmode = false;  // CTA-uniform: draw-ahead mode (see below)
  while (true) {
    // ---- compaction point
    if (pending) {
      s.store(g, p, cta_base + gi);
      if (game_len) game_len[cta_base + gi] = len;
      pending = false;
    }
    // (counts and ring are double-buffered by round parity: an empty warp can reach the next point while others still read)
    const uint32_t ab = __ballot_sync(FULL, alive);
    const int par = round & 1;
    uint32_t* cnt = s_cnt + par * 34;
    // (bit 16: the live lanes are not the lowest ones — the packed order of the draw-ahead mode needs them to be)
    if (lane == 0 && !producer) cnt[warp] = (uint32_t)__popc(ab) | ((ab & (ab + 1u)) ? 0x10000u : 0u);
    ++round;
    __syncthreads();
    int before = 0, live = 0, nonempty = 0;
    for (int w = 0; w < W; ++w) {
      const int c = (int)(cnt[w] & 0xFFFFu);
      before += (w < warp) ? c : 0;
      live += c;
      nonempty += (c > 0);
    }
    const PlayShare share(max(live, 1), W);
    bool uneven = false;  // CTA-uniform: some warp does not hold its even share (in its lowest lanes)
    for (int w = 0; w < W; ++w) uneven = uneven || (int)cnt[w] != share.count(w);
#ifdef DOGSTEP_TRACE
    if ((int)blockIdx.x == g_play_trace_block && threadIdx.x == 0 && round <= 64) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      g_play_trace[2 * (round - 1)] = now;
      g_play_trace[2 * (round - 1) + 1] = (unsigned long long)live | ((unsigned long long)nonempty << 32) | ((unsigned long long)t << 48);
    }
#endif
    if (live == 0 || t >= max_steps) break;
    // Draw-ahead mode.  Below ~100 live games the SM is no longer short of issue slots: an iteration lasts as long as ONE
    // warp's chain mask -> item list -> Threefry passes -> argmax -> move (~3 k cycles).  The draws do not depend on the
    // state — bits(split(rng_t, N+1)[j+1], a) is a function of (iteration, game, action) — so from here on the live games sit
    // on four warps (one per scheduler, one game per lane) that only run mask -> masked max -> move, and every other warp
    // computes, kPChunk iterations ahead, the draws of ALL 24 actions of every live game into a double-buffered table
    // (named barriers: full[b] helpers -> players, empty[b] players -> helpers).
    const bool enter_p = !pmode && live <= p_enter && W >= 8;
    if (enter_p || uneven) {  // CTA-uniform
      if (alive) {
        const int slot = before + __popc(ab & ((1u << lane) - 1u));
        uint32_t* x = s_x + slot;
        s.pack(x + 3 * T, T);
        x[0] = (uint32_t)(s.cur() & 0xFF) | ((uint32_t)(s.reward() & 0xFF) << 8);
        x[T] = (uint32_t)len;
        x[2 * T] = (uint32_t)gi;
      }
      __syncthreads();
      alive = lane < share.count(warp) && !producer;
      if (alive) {
        const uint32_t* x = s_x + share.start(warp) + lane;
        s.unpack(x + 3 * T, T);
        const uint32_t misc = x[0];
        s.cur() = (int)(int8_t)(misc & 0xFFu);
        s.reward() = (int)(int8_t)((misc >> 8) & 0xFFu);
        len = (int)x[T];
        gi = (int)x[2 * T];
      }
      // the next write to s_x happens after the next round's first barrier, i.e. after every read above
    }
    pmode = pmode || enter_p;
    const int q = share.start(warp) + lane;  // slot of this lane's game in the packed order (draw-ahead mode: warps 0..3)
    if (pmode) {
      if (alive) s_pmy[q] = (uint32_t)(game_offset + cta_base + gi + 1);
      __syncthreads();
    }
    // ---- kPlayRound lockstep iterations
    const int tend = min(t + round_len, max_steps);
    const uint2* rk = s_ring + par * (kRingMax + 2);  // rk[i] = the loop key of iteration t + i
    if (producer) {  // next round's keys (none after the last round)
      const int need = min(round_len, max_steps - tend);
      if (need > 0) {
        uint2* nk = s_ring + (par ^ 1) * (kRingMax + 2);
        const uint2 r0 = rk[round_len];
        Key2 r{r0.x, r0.y};
        if (lane == 0) nk[0] = r0;
        for (int i = 1; i <= need; ++i) {
          r = split_i(r, 0u);
          if (lane == 0) nk[i] = make_uint2(r.a, r.b);
        }
      }
      t = tend;
      continue;
    }
    if (pmode) {
      const int n_it = tend - t, nc = (n_it + kPChunk - 1) / kPChunk;
      if (warp >= 4) {  // ---- helpers: the draws of chunk c (iterations c * kPChunk ..) into buffer c & 1
        const int h = (int)threadIdx.x - 128, H = T - 128;
        const int pairs = live * kPChunk;  // pair = slot * kPChunk + iteration in the chunk
        for (int c = 0; c < nc + 2; ++c) {
          const int b = c & 1;
          if (c >= 2) named_bar_sync(3 + b, T);  // the players have consumed chunk c - 2
          if (c >= nc) continue;
          uint2* pk = s_pkey + b * (kPSlots * kPChunk);
          for (int j = h; j < pairs; j += H) {  // step keys: split(rng_t, N+1)[game + 1]
            const int it = c * kPChunk + (j & (kPChunk - 1));
            const uint2 r = rk[min(it, round_len)];
            const Key2 k = split_i(Key2{r.x, r.y}, s_pmy[j / kPChunk]);
            pk[j] = make_uint2(k.a, k.b);
          }
          named_bar_sync(5, H);
          uint32_t* dr = s_draw + b * (kPChunk * kPSlots * kPStride);
          const int items = pairs * 24;
          for (int j0 = h; j0 < items; j0 += 2 * H) {  // two independent Threefry chains per thread
            const int j1 = j0 + H;
            const bool h1 = j1 < items;
            const int p0 = j0 / 24, a0 = j0 - 24 * p0, p1 = h1 ? j1 / 24 : 0, a1 = h1 ? j1 - 24 * p1 : 0;
            const uint2 k0 = pk[p0], k1 = pk[p1];
            const uint32_t v0 = bits_i(Key2{k0.x, k0.y}, (uint32_t)a0) >> 9, v1 = bits_i(Key2{k1.x, k1.y}, (uint32_t)a1) >> 9;
            dr[((p0 & (kPChunk - 1)) * kPSlots + p0 / kPChunk) * kPStride + a0] = (v0 << 5) | (uint32_t)(23 - a0);
            if (h1) dr[((p1 & (kPChunk - 1)) * kPSlots + p1 / kPChunk) * kPStride + a1] = (v1 << 5) | (uint32_t)(23 - a1);
          }
          named_bar_arrive(1 + b, T);
        }
      } else {  // ---- players
        for (int c = 0; c < nc; ++c) {
          const int b = c & 1;
          named_bar_sync(1 + b, T);
          const uint32_t dr = (uint32_t)__cvta_generic_to_shared(s_draw + (b * kPChunk * kPSlots + q) * kPStride);
          uint4 v[6];  // this iteration's 24 draws; the next row is fetched while the rules run (the loads are pinned: left to
                       // itself the compiler sinks them behind the legal mask, 6 x 30 cycles on the critical path)
#pragma unroll
          for (int k = 0; k < 6; ++k) v[k] = lds128(dr + 16 * k);
#pragma unroll
          for (int ii = 0; ii < kPChunk; ++ii) {
            uint4 nv[6];
            if (ii + 1 < kPChunk) {
#pragma unroll
              for (int k = 0; k < 6; ++k) nv[k] = lds128(dr + (ii + 1) * (kPSlots * kPStride * 4) + 16 * k);
            }
            if (c * kPChunk + ii < n_it && alive) {
#ifdef DOGSTEP_TRACE
              const long long c0 = clock64();
#endif
              int cp = 0;
              const uint32_t m = s.mask(R, g, cp);
#ifdef DOGSTEP_TRACE
              const long long c1 = clock64();
              long long c2 = c1;
#endif
              if (m) {
                // largest 23-bit mantissa among the legal actions, lowest action index on ties; six independent chains
                uint32_t pm[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                  const uint32_t mk = m >> (4 * k);
                  pm[k] = max(max((mk & 1u) ? v[k].x : 0u, (mk & 2u) ? v[k].y : 0u), max((mk & 4u) ? v[k].z : 0u, (mk & 8u) ? v[k].w : 0u));
                }
                const uint32_t best = max(max(pm[0], pm[1]), max(max(pm[2], pm[3]), max(pm[4], pm[5])));
#ifdef DOGSTEP_TRACE
                c2 = clock64() + (best & 0u);
#endif
                s.step(R, cp, 23 - (int)(best & 31u));
              } else {
                s.no_step();
              }
              ++len;
              ++steps_done;
#ifdef DOGSTEP_TRACE
              if ((int)blockIdx.x == g_play_trace_block && live == 1) {
                const long long c3 = clock64() + (s.cur() & 0);
                atomicAdd(&g_play_trace[120], (unsigned long long)(c1 - c0));
                atomicAdd(&g_play_trace[121], (unsigned long long)(c2 - c1));
                atomicAdd(&g_play_trace[122], (unsigned long long)(c3 - c2));
                atomicAdd(&g_play_trace[123], 1ull);
                if (tr_last) atomicAdd(&g_play_trace[124], (unsigned long long)(c0 - tr_last));
                tr_last = c3;
              }
#endif
              if (s.done()) {
                alive = false;
                pending = true;
              }
            }
            if (ii + 1 < kPChunk) {
#pragma unroll
              for (int k = 0; k < 6; ++k) v[k] = nv[k];
            }
          }
          named_bar_arrive(3 + b, T);
        }
      }
      t = tend;
      continue;
    }
    const uint32_t my = (uint32_t)(game_offset + cta_base + gi + 1);
    Key2 key = split_i(Key2{rk[0].x, rk[0].y}, my);  // split(rng, N+1)[j+1]   (game_agent.py:60 / evaluate_agent.py:336)
    int ri = 0;
#pragma unroll 1
    for (; t < tend; ++t) {
      if (!__any_sync(FULL, alive)) break;
      ++ri;
      const uint2 rn = rk[ri];  // split(rng, N+1)[0] of this iteration = the loop key of the next
      int cp = 0;
      uint32_t m = 0u;
      if (alive) m = s.mask(R, g, cp);
      s_key[lane] = make_uint2(key.a, key.b);
      const int cnt = __popc(m);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(FULL, incl, 31);
      {  // predicated stores, no find-first-set chain: 24 x (store, bump) under the mask bits
        uint16_t* dst = s_items + (incl - cnt);
#pragma unroll
        for (int a = 0; a < 24; ++a) {
          if ((m >> a) & 1u) *dst++ = (uint16_t)(lane_hi | (uint32_t)a);
        }
      }
      __syncwarp();
      Key2 key_next;
      {  // first pass, peeled: items lane and lane + 32 plus NEXT iteration's step key — three independent Threefry
         // chains per lane, and the key derivation is off the critical path of the following iteration
        const bool h0 = lane < total, h1 = lane + 32 < total;
        const uint32_t it0 = h0 ? (uint32_t)s_items[lane] : 0u, it1 = h1 ? (uint32_t)s_items[lane + 32] : 0u;
        const uint32_t o0 = it0 >> 8, a0 = it0 & 0xFFu, o1 = it1 >> 8, a1 = it1 & 0xFFu;
        const uint2 k0 = s_key[o0], k1 = s_key[o1];
        const uint32_t v0 = bits_i(Key2{k0.x, k0.y}, a0) >> 9, v1 = bits_i(Key2{k1.x, k1.y}, a1) >> 9;
        key_next = split_i(Key2{rn.x, rn.y}, my);
        if (h0) atomicMax(&s_best[o0], (v0 << 5) | (23u - a0));
        if (h1) atomicMax(&s_best[o1], (v1 << 5) | (23u - a1));
      }
      for (int base = 64; base < total; base += 64) {
        const int j0 = base + lane, j1 = j0 + 32;
        if (base + 32 < total) {  // warp-uniform: two independent Threefry chains per lane
          const uint32_t it0 = s_items[j0], it1 = (j1 < total) ? (uint32_t)s_items[j1] : 0u;
          const uint32_t o0 = it0 >> 8, a0 = it0 & 0xFFu, o1 = it1 >> 8, a1 = it1 & 0xFFu;
          const uint2 k0 = s_key[o0], k1 = s_key[o1];
          const uint32_t v0 = bits_i(Key2{k0.x, k0.y}, a0) >> 9, v1 = bits_i(Key2{k1.x, k1.y}, a1) >> 9;
          atomicMax(&s_best[o0], (v0 << 5) | (23u - a0));
          if (j1 < total) atomicMax(&s_best[o1], (v1 << 5) | (23u - a1));
        } else if (j0 < total) {
          const uint32_t it0 = s_items[j0];
          const uint32_t o0 = it0 >> 8, a0 = it0 & 0xFFu;
          const uint2 k0 = s_key[o0];
          const uint32_t v0 = bits_i(Key2{k0.x, k0.y}, a0) >> 9;
          atomicMax(&s_best[o0], (v0 << 5) | (23u - a0));
        }
      }
      __syncwarp();
      if (alive) {
        if (m) {
          const int a = 23 - (int)(s_best[lane] & 31u);
          s.step(R, cp, a);  // map_action (deterministic_madn.py:469-479) is inside
        } else {
          s.no_step();
        }
        ++len;
        ++steps_done;
        if (s.done()) {  // finished: the lane is free from the next compaction point on
          alive = false;
          pending = true;
        }
      }
      s_best[lane] = 0u;
      key = key_next;
      __syncwarp();
    }
    t = tend;
  }
  if (alive || pending) {  // max_steps reached with the game still running
    s.store(g, p, cta_base + gi);
    if (game_len) game_len[cta_base + gi] = len;
  }
  if (total_steps) {
    unsigned v = steps_done;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0 && v) atomicAdd(total_steps, (unsigned long long)v);
  }
}

// MuZero_det_MADN/game_agent.py:12-22 — the rule dict of every training / benchmark configuration gets its own program

// dice_probabilities (classic_madn.py:14-18,208-228): float32 literals rounded from doubles
__device__ __forceinline__ void madn_dice_probabilities(const MadnGeom& g, const MadnRegs& s, float pr[6]) {
  const int locked = madn_soft_locked(g, s) && DS_RULE(g, DOGSTEP_RULE_DICE_RETHROW);
  if (locked && DS_RULE(g, DOGSTEP_RULE_START_ON_1)) {
    pr[0] = pr[5] = (float)(76.0 / 216);
    pr[1] = pr[2] = pr[3] = pr[4] = (float)(16.0 / 216);
  } else if (locked) {
    pr[0] = pr[1] = pr[2] = pr[3] = pr[4] = (float)(25.0 / 216);
    pr[5] = (float)(91.0 / 216);
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) pr[k] = (float)(1.0 / 6);
  }
}

// throw_die (:230-242): key, rng_key = split(env.key); die = choice(rng_key, 1..6, p = dice_probabilities(env))
__device__ __forceinline__ void madn_throw_die(const MadnGeom& g, MadnRegs& s, Key2& key) {
  float pr[6];
  madn_dice_probabilities(g, s, pr);
  const Key2 knew = split_i(key, 0), sub = split_i(key, 1);
  s.die = choice6(sub, pr) + 1;
  key = knew;
}

__global__ void __launch_bounds__(kThreads) k_madn_cls_throw_die(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                 float* __restrict__ probs, int write_die, int only_active) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  if (only_active && p.done[i] != 0) return;
  MadnRegs s;
  load_state<false>(g, p, i, s);
  if (probs) {
    float pr[6];
    madn_dice_probabilities(g, s, pr);
#pragma unroll
    for (int k = 0; k < 6; ++k) probs[i * 6 + k] = pr[k];
  }
  if (write_die) {
    Key2 key{p.key[2 * i], p.key[2 * i + 1]};
    madn_throw_die(g, s, key);
    p.die[i] = (int8_t)s.die;
    p.key[2 * i] = key.a;
    p.key[2 * i + 1] = key.b;
  }
}

// ---- true-env mctx callbacks of the deterministic game (MADN/deterministic_madn.py:480-590) -------------------------------
// winning_action / policy_function / rollout / value_function / root_fn / recurrent_fn, the callbacks MADN/simulate_deterministicMADN.py
// hands to mctx.gumbel_muzero_policy.  One warp per game; every lane holds the game in registers (generic rules: any geometry,
// any rule set), lane a < 24 owns action a = pin * 6 + move - 1: its legality, the hypothetical env_step that tells whether it
// wins (reward == 1), its Gumbel draw.  The env embedding is the state as floats: board[total], current_player, pins[4 n],
// reward, done, action_set[6 n].  The reference's rollout value is a float32[4] of four equal entries (its `winner == -1` test
// compares a bool array): +1 if the root player('s team) has won when the rollout stops, -1 otherwise, also at the 300-step cap;
// the scalar is returned.
constexpr int kTrueEnvWarps = 4;

__device__ __forceinline__ float madn_policy_lane(const MadnGeom& g, const MadnRegs& s, uint32_t m, int a) {
  MadnRegs t = s;
  madn_det_step(g, t, a / 6, a % 6 + 1, m);  // winning_action (:480-493): env_step on a copy
  return __fadd_rn(((m >> a) & 1u) ? 100.0f : 0.0f, t.reward == 1 ? 200.0f : 0.0f);
}

__device__ float madn_rollout_warp(const MadnGeom& g, const MadnRegs& s0, Key2 key, int lane) {
  const uint32_t FULL = 0xFFFFFFFFu;
  MadnRegs e = s0;
  const int a = min(lane, 23);
  for (int steps = 0; !e.done && steps < 300; ++steps) {
    const Key2 nk = split_i(key, 0u), sub = split_i(key, 1u);  // key, subkey = split(key)
    key = nk;
    const uint32_t m = madn_det_valid_mask(g, e);
    if (m == 0u) {
      madn_det_no_step(g, e);
      continue;
    }
    const float lg = madn_policy_lane(g, e, m, a);
    const float u = uniform_i(sub, (uint32_t)a, 1.17549435e-38f, 1.0f);
    const float v = __fadd_rn(-eval_log_f(-eval_log_f(u)), lg);  // jax.random.categorical: first maximum of gumbel + logits
    uint32_t ord = __float_as_uint(__fadd_rn(v, 0.0f));
    ord = (ord & 0x80000000u) ? ~ord : (ord | 0x80000000u);
    ord = lane < 24 ? ord : 0u;
    const uint32_t best = __reduce_max_sync(FULL, ord);
    const int act = __ffs(__ballot_sync(FULL, ord == best && lane < 24)) - 1;
    madn_det_step(g, e, act / 6, act % 6 + 1, m);  // map_action (:469-479)
  }
  return ((madn_winner_mask(g, e) >> gidx(s0.cur, 4)) & 1u) ? 1.0f : -1.0f;
}

__device__ void madn_from_emb(const MadnGeom& g, const float* __restrict__ f, MadnRegs& s) {
#pragma unroll
  for (int p = 0; p < 4; ++p) { s.occ[p] = 0ull; s.pins[p] = 0xFFFFFFFFu; s.as[p] = 0ull; }
  for (int c = 0; c < g.total; ++c) {
    const int v = (int)(int8_t)f[c];
#pragma unroll
    for (int p = 0; p < 4; ++p) s.occ[p] |= (v == p) ? (1ull << c) : 0ull;
  }
  const float* q = f + g.total;
  s.cur = (int)(int8_t)q[0];
  for (int p = 0; p < g.n; ++p) {
    uint32_t w = 0;
    for (int k = 0; k < 4; ++k) w |= (uint32_t)((int)(int8_t)q[1 + 4 * p + k] & 0xFF) << (8 * k);
    s.pins[p] = w;
  }
  q += 1 + 4 * g.n;
  s.reward = (int)(int8_t)q[0];
  s.done = ((int)q[1] & 0xFF) != 0;
  for (int p = 0; p < g.n; ++p) {
    uint64_t w = 0;
    for (int k = 0; k < 6; ++k) w |= (uint64_t)((int)(int8_t)q[2 + 6 * p + k] & 0xFF) << (8 * k);
    s.as[p] = w;
  }
  s.die = 0;
}

// every lane holds the same state; the lanes write the row together
__device__ void madn_to_emb(const MadnGeom& g, const MadnRegs& s, float* __restrict__ f, int lane) {
  const int E = g.total + 10 * g.n + 3;
  for (int k = lane; k < E; k += 32) {
    int v;
    if (k < g.total) {
      v = -1;
#pragma unroll
      for (int p = 0; p < 4; ++p) v = ((s.occ[p] >> k) & 1ull) ? p : v;  // later players win, as set_pins_on_board writes them
    } else if (k == g.total) {
      v = s.cur;
    } else if (k < g.total + 1 + 4 * g.n) {
      const int j = k - g.total - 1;
      v = byte_s(pick4(s.pins, j >> 2), j & 3);
    } else if (k == g.total + 1 + 4 * g.n) {
      v = s.reward;
    } else if (k == g.total + 2 + 4 * g.n) {
      v = s.done;
    } else {
      const int j = k - (g.total + 3 + 4 * g.n);
      v = byte_s64(pick4(s.as, j / 6), j % 6);
    }
    f[k] = (float)v;
  }
}

__global__ void __launch_bounds__(kTrueEnvWarps * 32) k_madn_det_policy_function(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                                float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kTrueEnvWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  MadnRegs s;
  load_state<true>(g, p, i, s);
  const uint32_t m = madn_det_valid_mask(g, s);
  const float lg = madn_policy_lane(g, s, m, min(lane, 23));
  if (lane < 24) logits[i * 24 + lane] = lg;
}

__global__ void __launch_bounds__(kTrueEnvWarps * 32) k_madn_det_root_fn(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                        const uint32_t* __restrict__ keys, float* __restrict__ prior,
                                                                        float* __restrict__ value, float* __restrict__ emb) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kTrueEnvWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  MadnRegs s;
  load_state<true>(g, p, i, s);
  const uint32_t m = madn_det_valid_mask(g, s);
  const float lg = madn_policy_lane(g, s, m, min(lane, 23));
  if (lane < 24) prior[i * 24 + lane] = lg;
  const float v = madn_rollout_warp(g, s, Key2{keys[2 * i], keys[2 * i + 1]}, lane);
  if (lane == 0) value[i] = v;
  madn_to_emb(g, s, emb + i * (g.total + 10 * g.n + 3), lane);
}

__global__ void __launch_bounds__(kTrueEnvWarps * 32) k_madn_det_recurrent_fn(const __grid_constant__ MadnGeom g, int64_t n,
                                                                             const uint32_t* __restrict__ keys,
                                                                             const int32_t* __restrict__ action,
                                                                             const float* __restrict__ emb_in, float* __restrict__ prior,
                                                                             float* __restrict__ value, float* __restrict__ reward,
                                                                             float* __restrict__ discount, float* __restrict__ emb_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kTrueEnvWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  const int E = g.total + 10 * g.n + 3;
  MadnRegs s;
  madn_from_emb(g, emb_in + i * E, s);
  const int a = action[i];
  madn_det_step(g, s, (int)(int8_t)floordiv(a, 6), (int)(int8_t)(floormod(a, 6) + 1), madn_det_valid_mask(g, s));  // map_action: both int8
  const uint32_t m = madn_det_valid_mask(g, s);
  const float lg = madn_policy_lane(g, s, m, min(lane, 23));
  if (lane < 24) prior[i * 24 + lane] = lg;
  const float v = s.done ? 0.0f : madn_rollout_warp(g, s, Key2{keys[2 * i], keys[2 * i + 1]}, lane);
  if (lane == 0) {
    value[i] = v;
    reward[i] = (float)s.reward;
    discount[i] = s.done ? 0.0f : -1.0f;
  }
  __syncwarp();
  madn_to_emb(g, s, emb_out + i * E, lane);  // emb_out may alias emb_in: every lane has read its input above
}

// ---- true-env mctx callbacks of the dice game (MADN/classic_madn.py:541-714) ------------------------------------
// winning_action / policy_function / rollout / root_fn / recurrent_fn (decision node -> afterstate) / recurrent_chance_fn
// (afterstate + die -> state), for mctx.stochastic_muzero_policy on the true env.  The reference as it stands raises before any
// of them returns (winning_action builds its scratch copy without the dataclass's `key` field, :551-565); what is restated is
// what its function bodies compute once that constructor call goes through (pinned by tests/golden/madn_cls_reference_trueenv.npz).
// One warp per game, every lane holds the game, lane a < 4 owns pin a.  Embedding: board[total], current_player, pins[4 n],
// reward, done, die, key as four 16-bit halves (a float32 cannot hold a uint32).  The rollout throws the die with the ENV's key
// chain (throw_die, :230-242) and draws the move with the rollout's.
__device__ __forceinline__ float madn_cls_policy_lane(const MadnGeom& g, const MadnRegs& s, uint32_t m, int a) {
  MadnRegs t = s;
  madn_cls_step(g, t, a, m);  // winning_action (:543-569)
  return __fadd_rn(((m >> a) & 1u) ? 100.0f : 0.0f, t.reward == 1 ? 200.0f : 0.0f);
}

__device__ float madn_cls_rollout_warp(const MadnGeom& g, const MadnRegs& s0, Key2 env_key, Key2 key, int lane) {
  const uint32_t FULL = 0xFFFFFFFFu;
  MadnRegs e = s0;
  const int a = min(lane, 3);
  for (int steps = 0; !e.done && steps < 300; ++steps) {
    const Key2 nk = split_i(key, 0u), sub = split_i(key, 1u);
    key = nk;
    madn_throw_die(g, e, env_key);
    const uint32_t m = madn_cls_valid_mask(g, e);
    if (m == 0u) {
      e.cur = (int)(int8_t)floormod(e.cur + 1, g.n);  // no_step (:353-365)
      continue;
    }
    const float lg = madn_cls_policy_lane(g, e, m, a);
    const float u = uniform_i(sub, (uint32_t)a, 1.17549435e-38f, 1.0f);
    const float v = __fadd_rn(-eval_log_f(-eval_log_f(u)), lg);
    uint32_t ord = __float_as_uint(__fadd_rn(v, 0.0f));
    ord = (ord & 0x80000000u) ? ~ord : (ord | 0x80000000u);
    ord = lane < 4 ? ord : 0u;
    const uint32_t best = __reduce_max_sync(FULL, ord);
    const int act = __ffs(__ballot_sync(FULL, ord == best && lane < 4)) - 1;
    madn_cls_step(g, e, act, m);
  }
  return ((madn_winner_mask(g, e) >> gidx(s0.cur, 4)) & 1u) ? 1.0f : -1.0f;
}

__device__ void madn_cls_from_emb(const MadnGeom& g, const float* __restrict__ f, MadnRegs& s, Key2& key) {
#pragma unroll
  for (int p = 0; p < 4; ++p) { s.occ[p] = 0ull; s.pins[p] = 0xFFFFFFFFu; s.as[p] = 0ull; }
  for (int c = 0; c < g.total; ++c) {
    const int v = (int)(int8_t)f[c];
#pragma unroll
    for (int p = 0; p < 4; ++p) s.occ[p] |= (v == p) ? (1ull << c) : 0ull;
  }
  const float* q = f + g.total;
  s.cur = (int)(int8_t)q[0];
  for (int p = 0; p < g.n; ++p) {
    uint32_t w = 0;
    for (int k = 0; k < 4; ++k) w |= (uint32_t)((int)(int8_t)q[1 + 4 * p + k] & 0xFF) << (8 * k);
    s.pins[p] = w;
  }
  q += 1 + 4 * g.n;
  s.reward = (int)(int8_t)q[0];
  s.done = ((int)q[1] & 0xFF) != 0;
  s.die = (int)(int8_t)q[2];
  key = Key2{(uint32_t)q[3] | ((uint32_t)q[4] << 16), (uint32_t)q[5] | ((uint32_t)q[6] << 16)};
}

__device__ void madn_cls_to_emb(const MadnGeom& g, const MadnRegs& s, Key2 key, float* __restrict__ f, int lane) {
  const int E = g.total + 4 * g.n + 8, tail = g.total + 1 + 4 * g.n;
  for (int k = lane; k < E; k += 32) {
    float v;
    if (k < g.total) {
      int b = -1;
#pragma unroll
      for (int p = 0; p < 4; ++p) b = ((s.occ[p] >> k) & 1ull) ? p : b;
      v = (float)b;
    } else if (k == g.total) {
      v = (float)s.cur;
    } else if (k < tail) {
      const int j = k - g.total - 1;
      v = (float)byte_s(pick4(s.pins, j >> 2), j & 3);
    } else {
      const int j = k - tail;
      v = j == 0 ? (float)s.reward : j == 1 ? (float)s.done : j == 2 ? (float)s.die
          : j == 3 ? (float)(key.a & 0xFFFFu) : j == 4 ? (float)(key.a >> 16) : j == 5 ? (float)(key.b & 0xFFFFu) : (float)(key.b >> 16);
    }
    f[k] = v;
  }
}

// mode 0: policy_function only (logits), 1: root_fn
__global__ void __launch_bounds__(kTrueEnvWarps * 32) k_madn_cls_root_fn(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                        const uint32_t* __restrict__ keys, float* __restrict__ prior,
                                                                        float* __restrict__ value, float* __restrict__ emb, int mode) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kTrueEnvWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  MadnRegs s;
  load_state<false>(g, p, i, s);
  const uint32_t m = madn_cls_valid_mask(g, s);
  const float lg = madn_cls_policy_lane(g, s, m, min(lane, 3));
  if (lane < 4) prior[i * 4 + lane] = lg;
  if (mode == 0) return;
  const Key2 ek{p.key[2 * i], p.key[2 * i + 1]};
  const float v = madn_cls_rollout_warp(g, s, ek, Key2{keys[2 * i], keys[2 * i + 1]}, lane);
  if (lane == 0) value[i] = v;
  madn_cls_to_emb(g, s, ek, emb + i * (g.total + 4 * g.n + 8), lane);
}

// recurrent_fn (:657-688), the decision node
__global__ void __launch_bounds__(kTrueEnvWarps * 32) k_madn_cls_decision_fn(const __grid_constant__ MadnGeom g, int64_t n,
                                                                            const uint32_t* __restrict__ keys,
                                                                            const int32_t* __restrict__ action,
                                                                            const float* __restrict__ emb_in,
                                                                            float* __restrict__ chance_logits,
                                                                            float* __restrict__ afterstate_value,
                                                                            float* __restrict__ emb_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kTrueEnvWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  const int E = g.total + 4 * g.n + 8;
  MadnRegs s;
  Key2 ek;
  madn_cls_from_emb(g, emb_in + i * E, s, ek);
  const uint32_t m = madn_cls_valid_mask(g, s);
  if (m == 0u) s.cur = (int)(int8_t)floormod(s.cur + 1, g.n);  // no_step
  else madn_cls_step(g, s, action[i], m);
  if (lane < 6) chance_logits[i * 6 + lane] = eval_log_f((float)(1.0 / 6.0));  // jnp.ones(6) * jnp.log(1.0 / 6.0)
  const float v = madn_cls_rollout_warp(g, s, ek, Key2{keys[2 * i], keys[2 * i + 1]}, lane);
  if (lane == 0) afterstate_value[i] = v;
  __syncwarp();
  madn_cls_to_emb(g, s, ek, emb_out + i * E, lane);
}

// recurrent_chance_fn (:624-655)
__global__ void __launch_bounds__(kTrueEnvWarps * 32) k_madn_cls_chance_fn(const __grid_constant__ MadnGeom g, int64_t n,
                                                                          const uint32_t* __restrict__ keys,
                                                                          const int32_t* __restrict__ outcome,
                                                                          const float* __restrict__ emb_in, float* __restrict__ action_logits,
                                                                          float* __restrict__ value, float* __restrict__ reward,
                                                                          float* __restrict__ discount, float* __restrict__ emb_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kTrueEnvWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  const int E = g.total + 4 * g.n + 8;
  MadnRegs s;
  Key2 ek;
  madn_cls_from_emb(g, emb_in + i * E, s, ek);
  s.die = (int)(int8_t)(outcome[i] + 1);  // set_die(afterstate, chance_outcome + 1)
  const uint32_t m = madn_cls_valid_mask(g, s);
  if (lane < 4) action_logits[i * 4 + lane] = ((m >> lane) & 1u) ? 1.0f : 0.0f;
  const float v = madn_cls_rollout_warp(g, s, ek, Key2{keys[2 * i], keys[2 * i + 1]}, lane);
  if (lane == 0) {
    value[i] = v;
    reward[i] = (float)s.reward;
    discount[i] = s.done ? 0.0f : 1.0f;
  }
  __syncwarp();
  madn_cls_to_emb(g, s, ek, emb_out + i * E, lane);
}

// ---- self-play bookkeeping: one lockstep iteration of play_batch_of_games_jitted after the search ---------------------
// MuZero_det_MADN/game_agent.py:64-148 (do_active_step) and MuZero_Classic_MADN/game_agent_stochastic.py:86-204.
// CTAs of 256 threads take 32 games: warp 0 steps them thread per game (coalesced leaf loads, env_step / no_step, targets,
// the scalar trajectory entries), then all eight warps move the observation / policy rows (16-byte vectors).  The first
// version gave every game a warp whose lane 0 ran the env step alone: 6.9 of 32 lanes active, 43 us per iteration.
constexpr int kAgentGames = 32;
template <bool DET>
__global__ void __launch_bounds__(256) k_madn_agent_step(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                         const int32_t* __restrict__ action, const float* __restrict__ root_value,
                                                         const float* __restrict__ weights, const int8_t* __restrict__ obs,
                                                         dogstep_replay_arrays tr) {
  __shared__ int sh_idx[kAgentGames], sh_valid[kAgentGames];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kAgentGames;
  const int A = DET ? 24 : 4;
  if (warp == 0) {
    const int64_t i = base + lane;
    int out_idx = -1, has_valid = 0;
    if (i < n && p.done[i] == 0) {  // do_skip_step: finished games are left untouched
      MadnRegs s;
      load_state<DET>(g, p, i, s);
      const uint32_t m = DET ? madn_det_valid_mask(g, s) : madn_cls_valid_mask(g, s);
      const int teams = DS_RULE(g, DOGSTEP_RULE_TEAMS);
      const int pid = s.cur;
      const int team_before = teams ? floormod(pid, 2) : -1;
      const int idx = tr.episode_lengths[i];
      const int dice = s.die;
      has_valid = m != 0u;
      int act = -1, rew_t = 1, disc_t = 1;
      if (has_valid) {
        act = action[i];
        if (DET) madn_det_step(g, s, act / 6, act % 6 + 1, m);  // map_action
        else madn_cls_step(g, s, act, m);
        const int next_team = teams ? floormod(s.cur, 2) : -1;
        rew_t = (s.done && s.reward > 0) ? 2 : ((s.done && s.reward < 0) ? 0 : 1);
        disc_t = s.done ? 1 : (teams ? (team_before == next_team ? 2 : 0) : (pid == s.cur ? 2 : 0));
        if (DET) store_aset(g, p.aset, i, s);
        store_board(g, p.board, i, s);
        store_pins(g, p.pins, i, s);
        p.reward[i] = (int8_t)s.reward;
        p.done[i] = (uint8_t)s.done;
      } else {
        if (DET) { madn_det_no_step(g, s); store_aset(g, p.aset, i, s); }
        else s.cur = (int)(int8_t)floormod(s.cur + 1, g.n);
      }
      p.cur[i] = (int8_t)s.cur;
      if (idx >= 0 && idx < tr.max_episode_length) {  // .at[idx].set drops out-of-range rows
        const int64_t r = i * tr.max_episode_length + idx;
        tr.actions[r] = act;
        tr.rewards[r] = rew_t;
        tr.root_values[r] = has_valid ? root_value[i] : 0.0f;
        tr.masks[r] = has_valid ? 1.0f : 0.0f;
        tr.players[r] = pid;
        tr.teams[r] = team_before;
        tr.discounts[r] = disc_t;
        if (!DET) {
          tr.dice_outcomes[r] = dice;
          // dice_probabilities(next_env) (game_agent_stochastic.py:160)
          const int locked = madn_soft_locked(g, s) && DS_RULE(g, DOGSTEP_RULE_DICE_RETHROW);
          float pr[6];
          if (locked && DS_RULE(g, DOGSTEP_RULE_START_ON_1)) { pr[0] = pr[5] = (float)(76.0 / 216); pr[1] = pr[2] = pr[3] = pr[4] = (float)(16.0 / 216); }
          else if (locked) { pr[0] = pr[1] = pr[2] = pr[3] = pr[4] = (float)(25.0 / 216); pr[5] = (float)(91.0 / 216); }
          else { for (int k = 0; k < 6; ++k) pr[k] = (float)(1.0 / 6); }
          for (int k = 0; k < 6; ++k) tr.dice_distributions[r * 6 + k] = pr[k];
        }
        out_idx = idx;
      }
      tr.episode_lengths[i] = idx + 1;
    }
    sh_idx[lane] = out_idx;
    sh_valid[lane] = has_valid;
  }
  __syncthreads();
  for (int q = warp; q < kAgentGames; q += 8) {
    const int idx = sh_idx[q];
    if (idx < 0) continue;
    const int64_t i = base + q;
    const int has_valid = sh_valid[q];
    const int64_t r = i * tr.max_episode_length + idx;
    if (has_valid) coop_copy_bytes(tr.child_visits + r * A, weights + i * A, (int64_t)A * 4, lane, 32);
    else coop_zero_bytes(tr.child_visits + r * A, (int64_t)A * 4, lane, 32);
    const int8_t* src = obs + i * tr.obs_size;
    if (tr.obs_is_int8) {
      int8_t* d = (int8_t*)tr.observations + r * tr.obs_size;
      if (has_valid) coop_copy_bytes(d, src, tr.obs_size, lane, 32);
      else coop_zero_bytes(d, tr.obs_size, lane, 32);
    } else {
      float* d = (float*)tr.observations + r * tr.obs_size;
      if (has_valid) coop_widen_i8_f32(d, src, tr.obs_size, lane, 32);
      else coop_zero_bytes(d, (int64_t)tr.obs_size * 4, lane, 32);
    }
  }
}

// keys[i] -> split(keys[i], m)[index] for every game (game_agent.py:80 hands step_keys[i] to run_muzero_mcts, which
// keeps split(key)[1], muzero_deterministic_madn.py:665)
__global__ void k_random_split_each(const uint32_t* __restrict__ keys, int64_t n, uint32_t index, uint32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Key2 k = split_i(Key2{keys[2 * i], keys[2 * i + 1]}, index);
  out[2 * i] = k.a;
  out[2 * i + 1] = k.b;
}

// `rng_key, *step_keys = jax.random.split(rng_key, n + 1)` with the loop key RESIDENT ON THE DEVICE (game_agent.py:60): the
// self-play loop then needs no host round trip per lockstep iteration and one iteration can be replayed as a CUDA graph.
__global__ void k_random_split_chain_keys(const uint32_t* __restrict__ key, int64_t n, uint32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Key2 k = split_i(Key2{key[0], key[1]}, (uint32_t)(i + 1));
  reinterpret_cast<uint2*>(out)[i] = make_uint2(k.a, k.b);
}
__global__ void k_random_split_chain_carry(uint32_t* __restrict__ key) {
  Key2 k = split_i(Key2{key[0], key[1]}, 0u);
  key[0] = k.a;
  key[1] = k.b;
}

// ---- jax.random helpers -------------------------------------------------------------------------
__global__ void k_random_split(Key2 key, int64_t n, uint32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Key2 k = split_i(key, (uint32_t)i);
  reinterpret_cast<uint2*>(out)[i] = make_uint2(k.a, k.b);
}
__global__ void k_random_randint(Key2 key, int64_t n, int32_t lo, int32_t hi, int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = randint_i(key, (uint32_t)i, lo, hi);
}
__global__ void k_random_uniform(Key2 key, int64_t n, float lo, float hi, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = uniform_i(key, (uint32_t)i, lo, hi);
}
__global__ void k_random_bits(Key2 key, int64_t n, uint32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = bits_i(key, (uint32_t)i);
}

// ---- host-side argument plumbing -------------------------------------------------------------
static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

static int det_ptrs(const dogstep_madn_det_state* s, MadnPtrs* p) {
  if (!s || !s->board || !s->current_player || !s->pins || !s->reward || !s->done || !s->action_set)
    return DOGSTEP_ERR_INVALID_ARG;
  *p = MadnPtrs{s->board, s->current_player, s->pins, s->reward, s->done, s->action_set, nullptr, s->key};
  return DOGSTEP_OK;
}
static int cls_ptrs(const dogstep_madn_cls_state* s, MadnPtrs* p) {
  if (!s || !s->board || !s->current_player || !s->pins || !s->reward || !s->done || !s->die)
    return DOGSTEP_ERR_INVALID_ARG;
  *p = MadnPtrs{s->board, s->current_player, s->pins, s->reward, s->done, nullptr, s->die, s->key};
  return DOGSTEP_OK;
}

#define DS_PROLOGUE(PTRFN)                                   \
  MadnGeom g;                                                \
  MadnPtrs p;                                                \
  if (n < 0) return DOGSTEP_ERR_INVALID_ARG;                 \
  if (int rc = madn_make_geom(cfg, &g)) return rc;           \
  if (int rc = PTRFN(s, &p)) return rc;                      \
  if (n == 0) return DOGSTEP_OK;                             \
  cudaStream_t st = (cudaStream_t)stream;

}  // namespace dogstep

using namespace dogstep;

extern "C" {

int dogstep_madn_det_reset(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* seeds,
                           int32_t starting_player, void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!seeds || !p.key) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_reset<true><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, seeds, starting_player);
  return check_launch();
}

int dogstep_madn_cls_reset(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* seeds,
                           int32_t starting_player, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!seeds || !p.key) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_reset<false><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, seeds, starting_player);
  return check_launch();
}

int dogstep_madn_set_pins_on_board(const int8_t* pins, int8_t* board, int64_t n, const dogstep_madn_cfg* cfg, void* stream) {
  MadnGeom g;
  if (n < 0 || !pins || !board) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = madn_make_geom(cfg, &g)) return rc;
  if (n == 0) return DOGSTEP_OK;
  k_madn_set_pins_on_board<<<blocks_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(g, pins, board, n);
  return check_launch();
}

int dogstep_madn_det_valid_action(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, uint8_t* mask,
                                  void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!mask) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_valid_action<true><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, mask);
  return check_launch();
}

int dogstep_madn_cls_valid_action(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, uint8_t* mask,
                                  void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!mask) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_valid_action<false><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, mask);
  return check_launch();
}

int dogstep_madn_det_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int8_t* action,
                          int8_t* reward, uint8_t* done, void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!action) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_step<true><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, action, reward, done);
  return check_launch();
}

int dogstep_madn_cls_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int8_t* action,
                          int8_t* reward, uint8_t* done, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!action) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_step<false><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, action, reward, done);
  return check_launch();
}

int dogstep_madn_det_no_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, int8_t* reward,
                             uint8_t* done, void* stream) {
  DS_PROLOGUE(det_ptrs)
  k_madn_no_step<true><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, reward, done);
  return check_launch();
}

int dogstep_madn_cls_no_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, int8_t* reward,
                             uint8_t* done, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  k_madn_no_step<false><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, reward, done);
  return check_launch();
}

int dogstep_madn_det_encode_board(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, int8_t* obs,
                                  void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!obs) return DOGSTEP_ERR_INVALID_ARG;
  int64_t words = n * (8 * g.n + 2) * (g.total >> 2);
  k_madn_encode_board<true><<<blocks_for(words, 256), 256, 0, st>>>(g, p, n, obs);
  return check_launch();
}

int dogstep_madn_cls_encode_board(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, int8_t* obs,
                                  void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!obs) return DOGSTEP_ERR_INVALID_ARG;
  int64_t words = n * (2 * g.n + 3) * (g.total >> 2);
  k_madn_encode_board<false><<<blocks_for(words, 256), 256, 0, st>>>(g, p, n, obs);
  return check_launch();
}

int dogstep_madn_det_embed_dim(const dogstep_madn_cfg* cfg) {
  MadnGeom g;
  if (madn_make_geom(cfg, &g)) return -1;
  return g.total + 10 * g.n + 3;
}

int dogstep_madn_det_policy_function(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, float* logits,
                                     void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!logits) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_det_policy_function<<<blocks_for(n, kTrueEnvWarps), kTrueEnvWarps * 32, 0, st>>>(g, p, n, logits);
  return check_launch();
}

int dogstep_madn_det_root_fn(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys,
                             float* prior_logits, float* value, float* embedding, void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!keys || !prior_logits || !value || !embedding) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_det_root_fn<<<blocks_for(n, kTrueEnvWarps), kTrueEnvWarps * 32, 0, st>>>(g, p, n, keys, prior_logits, value, embedding);
  return check_launch();
}

int dogstep_madn_det_recurrent_fn(int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys, const int32_t* action,
                                  const float* embedding_in, float* prior_logits, float* value, float* reward, float* discount,
                                  float* embedding_out, void* stream) {
  MadnGeom g;
  if (n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = madn_make_geom(cfg, &g)) return rc;
  if (!keys || !action || !embedding_in || !prior_logits || !value || !reward || !discount || !embedding_out) return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_madn_det_recurrent_fn<<<blocks_for(n, kTrueEnvWarps), kTrueEnvWarps * 32, 0, (cudaStream_t)stream>>>(
      g, n, keys, action, embedding_in, prior_logits, value, reward, discount, embedding_out);
  return check_launch();
}

int dogstep_madn_cls_embed_dim(const dogstep_madn_cfg* cfg) {
  MadnGeom g;
  if (madn_make_geom(cfg, &g)) return -1;
  return g.total + 4 * g.n + 8;
}

int dogstep_madn_cls_policy_function(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, float* logits,
                                     void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!logits) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_cls_root_fn<<<blocks_for(n, kTrueEnvWarps), kTrueEnvWarps * 32, 0, st>>>(g, p, n, nullptr, logits, nullptr, nullptr, 0);
  return check_launch();
}

int dogstep_madn_cls_root_fn(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys,
                             float* prior_logits, float* value, float* embedding, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!keys || !prior_logits || !value || !embedding || !p.key) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_cls_root_fn<<<blocks_for(n, kTrueEnvWarps), kTrueEnvWarps * 32, 0, st>>>(g, p, n, keys, prior_logits, value, embedding, 1);
  return check_launch();
}

int dogstep_madn_cls_decision_recurrent_fn(int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys, const int32_t* action,
                                           const float* embedding_in, float* chance_logits, float* afterstate_value,
                                           float* embedding_out, void* stream) {
  MadnGeom g;
  if (n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = madn_make_geom(cfg, &g)) return rc;
  if (!keys || !action || !embedding_in || !chance_logits || !afterstate_value || !embedding_out) return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_madn_cls_decision_fn<<<blocks_for(n, kTrueEnvWarps), kTrueEnvWarps * 32, 0, (cudaStream_t)stream>>>(
      g, n, keys, action, embedding_in, chance_logits, afterstate_value, embedding_out);
  return check_launch();
}

int dogstep_madn_cls_chance_recurrent_fn(int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys, const int32_t* chance_outcome,
                                         const float* embedding_in, float* action_logits, float* value, float* reward,
                                         float* discount, float* embedding_out, void* stream) {
  MadnGeom g;
  if (n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = madn_make_geom(cfg, &g)) return rc;
  if (!keys || !chance_outcome || !embedding_in || !action_logits || !value || !reward || !discount || !embedding_out)
    return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_madn_cls_chance_fn<<<blocks_for(n, kTrueEnvWarps), kTrueEnvWarps * 32, 0, (cudaStream_t)stream>>>(
      g, n, keys, chance_outcome, embedding_in, action_logits, value, reward, discount, embedding_out);
  return check_launch();
}

int dogstep_madn_det_random_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                 const uint32_t* host_rng_key, int64_t game_offset, unsigned long long* active_count,
                                 void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!host_rng_key) return DOGSTEP_ERR_INVALID_ARG;
  Key2 rng{host_rng_key[0], host_rng_key[1]};
  k_madn_det_random_step<<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, rng, game_offset, active_count);
  return check_launch();
}

int dogstep_madn_det_play_random(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                 const uint32_t* host_rng_key, int64_t game_offset, int32_t max_steps, int32_t* game_len,
                                 unsigned long long* total_steps, void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!host_rng_key || max_steps < 0) return DOGSTEP_ERR_INVALID_ARG;
  Key2 rng{host_rng_key[0], host_rng_key[1]};
#ifdef DOGSTEP_TRACE
  if (const char* tb = getenv("DOGSTEP_PLAY_TRACE")) {
    const int b = atoi(tb);
    unsigned long long z[128] = {0};
    cudaMemcpyToSymbol(g_play_trace_block, &b, sizeof(b));
    cudaMemcpyToSymbol(g_play_trace, z, sizeof(z));
  }
#endif
  if (g.n == 4 && g.d == 10) {
    // the games are spread evenly over the SMs (config 2: 65,536 games -> 148 CTAs of 443), at most 512 per CTA
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    int64_t per = (n + sms - 1) / sms;
    if (per > kPlayMaxThreads) per = kPlayMaxThreads;
    const int gpc = (int)per;                          // games per CTA (config 2: 443 on each of the 148 SMs)
    // game warps: a multiple of the four schedulers (see PlayShare), at least eight (four players + helpers of the draw-ahead mode)
    const int threads = gpc > 128 ? ((gpc + 127) / 128) * 128 : 256;
    int p_enter = kPEnter;
#ifdef DOGSTEP_TRACE
    if (const char* e = getenv("DOGSTEP_PENTER")) p_enter = atoi(e);
#endif
    const size_t smem = play_smem_bytes(threads);
    const unsigned blocks = blocks_for(n, gpc);
    int round_len = kPlayRound;  // <= kRingMax
#ifdef DOGSTEP_TRACE
    if (const char* e = getenv("DOGSTEP_ROUND")) round_len = atoi(e);
#endif
    if (g.rules == kTrainRules) {
      cudaFuncSetAttribute(k_madn_det_play_cta<kTrainRules>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_madn_det_play_cta<kTrainRules><<<blocks, threads + 32, smem, st>>>(g, p, n, rng, game_offset, max_steps, game_len, total_steps, round_len, gpc, p_enter);
    } else {
      cudaFuncSetAttribute(k_madn_det_play_cta<kRulesRuntime>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_madn_det_play_cta<kRulesRuntime><<<blocks, threads + 32, smem, st>>>(g, p, n, rng, game_offset, max_steps, game_len, total_steps, round_len, gpc, p_enter);
    }
  } else {
    k_madn_det_play_random<<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, rng, game_offset, max_steps, game_len,
                                                                         total_steps);
  }
#ifdef DOGSTEP_TRACE
  if (getenv("DOGSTEP_PLAY_TRACE")) {
    cudaStreamSynchronize(st);
    unsigned long long h[128];
    cudaMemcpyFromSymbol(h, g_play_trace, sizeof(h));
    if (h[123])
      fprintf(stderr, "draw-ahead player, one live game: mask %.0f  max %.0f  step %.0f  between iterations %.0f cycles (%llu iterations)\n",
              (double)h[120] / h[123], (double)h[121] / h[123], (double)h[122] / h[123], (double)h[124] / h[123], h[123]);
    for (int r = 0; r < 60 && h[2 * r]; ++r)
      fprintf(stderr, "round %2d  t %4llu  +%8.1f us  live %5llu  warps %3llu\n", r, h[2 * r + 1] >> 48, (h[2 * r] - h[0]) / 1e3,
              h[2 * r + 1] & 0xFFFFFFFFull, (h[2 * r + 1] >> 32) & 0xFFFFull);
  }
#endif
  return check_launch();
}

int dogstep_madn_cls_throw_die(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!p.key) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_cls_throw_die<<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, nullptr, 1, 0);
  return check_launch();
}

int dogstep_madn_cls_dice_probabilities(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, float* probs,
                                        void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!probs) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_cls_throw_die<<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, probs, 0, 0);
  return check_launch();
}

int dogstep_random_split(const uint32_t* host_key, int64_t n, uint32_t* out, void* stream) {
  if (!host_key || !out || n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_random_split<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Key2{host_key[0], host_key[1]}, n, out);
  return check_launch();
}
int dogstep_random_randint(const uint32_t* host_key, int64_t n, int32_t lo, int32_t hi, int32_t* out, void* stream) {
  if (!host_key || !out || n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_random_randint<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Key2{host_key[0], host_key[1]}, n, lo, hi, out);
  return check_launch();
}
int dogstep_random_uniform(const uint32_t* host_key, int64_t n, float lo, float hi, float* out, void* stream) {
  if (!host_key || !out || n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_random_uniform<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Key2{host_key[0], host_key[1]}, n, lo, hi, out);
  return check_launch();
}
int dogstep_random_bits(const uint32_t* host_key, int64_t n, uint32_t* out, void* stream) {
  if (!host_key || !out || n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_random_bits<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Key2{host_key[0], host_key[1]}, n, out);
  return check_launch();
}

int dogstep_madn_cls_throw_die_active(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!p.key) return DOGSTEP_ERR_INVALID_ARG;
  k_madn_cls_throw_die<<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, nullptr, 1, 1);
  return check_launch();
}

static int traj_check(const dogstep_replay_arrays* tr, int64_t n, int obs_size, int action_dim, int stochastic) {
  if (!tr || tr->capacity < n || tr->obs_size != obs_size || tr->action_dim != action_dim || tr->stochastic != stochastic)
    return DOGSTEP_ERR_INVALID_ARG;
  if (!tr->observations || !tr->actions || !tr->rewards || !tr->root_values || !tr->child_visits || !tr->masks || !tr->players ||
      !tr->teams || !tr->discounts || !tr->episode_lengths)
    return DOGSTEP_ERR_INVALID_ARG;
  if (stochastic && (!tr->dice_outcomes || !tr->dice_distributions)) return DOGSTEP_ERR_INVALID_ARG;
  return DOGSTEP_OK;
}

int dogstep_madn_det_agent_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* action,
                                const float* root_value, const float* action_weights, const int8_t* obs,
                                const dogstep_replay_arrays* traj, void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!action || !root_value || !action_weights || !obs) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = traj_check(traj, n, (8 * g.n + 2) * g.total, 24, 0)) return rc;
  k_madn_agent_step<true><<<blocks_for(n, kAgentGames), 256, 0, st>>>(g, p, n, action, root_value, action_weights, obs, *traj);
  return check_launch();
}

int dogstep_madn_cls_agent_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* action,
                                const float* root_value, const float* action_weights, const int8_t* obs,
                                const dogstep_replay_arrays* traj, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!action || !root_value || !action_weights || !obs) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = traj_check(traj, n, (2 * g.n + 3) * g.total, 4, 1)) return rc;
  k_madn_agent_step<false><<<blocks_for(n, kAgentGames), 256, 0, st>>>(g, p, n, action, root_value, action_weights, obs, *traj);
  return check_launch();
}

int dogstep_madn_det_eval_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* agent_type,
                               const int32_t* search_action, const uint32_t* host_rng_key, int64_t game_offset, int32_t* winners,
                               unsigned long long* active_count, void* stream) {
  DS_PROLOGUE(det_ptrs)
  if (!agent_type || !host_rng_key) return DOGSTEP_ERR_INVALID_ARG;
  for (int q = 0; q < g.n; ++q)
    if (agent_type[q] != 2 && agent_type[q] != 3 && !search_action) return DOGSTEP_ERR_INVALID_ARG;  // a search seat needs actions
  const int4 at = make_int4(agent_type[0], agent_type[1], agent_type[2], agent_type[3]);
  k_madn_eval_step<true><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, at, search_action, Key2{host_rng_key[0], host_rng_key[1]},
                                                                       game_offset, winners, active_count);
  return check_launch();
}

int dogstep_madn_cls_eval_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* agent_type,
                               const int32_t* search_action, const uint32_t* host_rng_key, int64_t game_offset, int32_t* winners,
                               unsigned long long* active_count, void* stream) {
  DS_PROLOGUE(cls_ptrs)
  if (!agent_type || !host_rng_key) return DOGSTEP_ERR_INVALID_ARG;
  for (int q = 0; q < g.n; ++q)
    if (agent_type[q] != 2 && agent_type[q] != 3 && !search_action) return DOGSTEP_ERR_INVALID_ARG;
  const int4 at = make_int4(agent_type[0], agent_type[1], agent_type[2], agent_type[3]);
  k_madn_eval_step<false><<<blocks_for(n, kThreads), kThreads, 0, st>>>(g, p, n, at, search_action, Key2{host_rng_key[0], host_rng_key[1]},
                                                                        game_offset, winners, active_count);
  return check_launch();
}

// host-side scalar key arithmetic (no device work): jax.random.split(key, num) for small num, and the loop-key chain
// rng <- split(rng, N + 1)[0] applied `steps` times (element 0 of a split does not depend on N) — what a host driver that
// launches k lockstep iterations at a time needs between launches
int dogstep_host_split(const uint32_t* key, int32_t num, uint32_t* out) {
  if (!key || !out || num < 0) return DOGSTEP_ERR_INVALID_ARG;
  const Key2 k{key[0], key[1]};
  for (int32_t i = 0; i < num; ++i) {
    const Key2 o = split_i(k, (uint32_t)i);
    out[2 * i] = o.a;
    out[2 * i + 1] = o.b;
  }
  return DOGSTEP_OK;
}
int dogstep_host_key_chain(const uint32_t* key, int32_t steps, uint32_t* out) {
  if (!key || !out || steps < 0) return DOGSTEP_ERR_INVALID_ARG;
  Key2 k{key[0], key[1]};
  for (int32_t i = 0; i < steps; ++i) k = split_i(k, 0u);
  out[0] = k.a;
  out[1] = k.b;
  return DOGSTEP_OK;
}

int dogstep_random_split_chain(uint32_t* key, int64_t n, uint32_t* step_keys, void* stream) {
  if (!key || !step_keys || n < 0 || n >= 0xFFFFFFFFll) return DOGSTEP_ERR_INVALID_ARG;
  if (n) k_random_split_chain_keys<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(key, n, step_keys);
  k_random_split_chain_carry<<<1, 1, 0, (cudaStream_t)stream>>>(key);  // after every reader of the old key (stream order)
  return check_launch();
}

int dogstep_random_split_each(const uint32_t* keys, int64_t n, uint32_t index, uint32_t* out, void* stream) {
  if (!keys || !out || n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_random_split_each<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(keys, n, index, out);
  return check_launch();
}

}  // extern "C"
