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
__device__ __forceinline__ int categorical_masked(Key2 key, uint32_t mask) {
  int best = -1;
  uint32_t bm = 0;
  for (uint32_t m = mask; m; m &= m - 1) {
    int a = __ffs(m) - 1;
    uint32_t mant = bits_i(key, (uint32_t)a) >> 9;
    if (best < 0 || mant > bm) { best = a; bm = mant; }
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

__global__ void __launch_bounds__(kThreads) k_madn_det_random_step(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                   Key2 rng, int64_t game_offset,
                                                                   unsigned long long* __restrict__ active_count) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int active = 0;
  if (i < n && p.done[i] == 0) {
    MadnRegs s;
    load_state<true>(g, p, i, s);
    madn_det_random_turn(g, s, split_i(rng, (uint32_t)(game_offset + i + 1)));
    store_det_all(g, p, i, s);
    active = 1;
  }
  if (active_count) {
    unsigned b = __ballot_sync(0xFFFFFFFFu, active);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(active_count, (unsigned long long)__popc(b));
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
#pragma unroll 1
  for (int a = 0; a < 24; ++a) {
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
    const float logit = ((m >> a) & 1u) ? __fdiv_rn(score, 0.25f) : __int_as_float(0xFF800000);
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

__global__ void __launch_bounds__(kThreads) k_madn_det_play_random(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                   Key2 rng0, int64_t game_offset, int max_steps,
                                                                   int32_t* __restrict__ game_len,
                                                                   unsigned long long* __restrict__ total_steps) {
  __shared__ uint16_t s_items[kPlayWarps][32 * 24];
  __shared__ uint32_t s_mant[kPlayWarps][32 * 24];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t FULL = 0xFFFFFFFFu;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int len = 0;
  MadnRegs s;
  bool alive = false;
  if (i < n) {
    load_state<true>(g, p, i, s);
    alive = !s.done;
  }
  Key2 rng = rng0;
  const uint32_t my = (uint32_t)(game_offset + i + 1);
  uint16_t* items = s_items[warp];
  uint32_t* mant = s_mant[warp];
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
  if (i < n) {
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

// Persistent lockstep loop for 4 players, distance 10 — every configuration of the reference (the generic kernel above
// covers other geometries).  One CTA per SM (up to 512 games), one game per lane held in registers.  Per iteration:
//   1. each lane takes its step key and derives the 24-bit legal mask (madn_fast.cuh: branch-free bit rows);
//   2. the (game, action) pairs of the warp are compacted into a shared list and dealt out evenly to the 32 lanes,
//      two per lane and pass, so the Threefry calls — one per LEGAL action, half of all instructions — are
//      load-balanced; the first pass also derives the NEXT iteration's step key (a third independent chain);
//   3. the argmax of the categorical draw is a shared-memory atomicMax per game on (mantissa << 5 | 23 - action):
//      largest 23-bit mantissa first, lowest action index on ties — jax.random.categorical's choice;
//   4. each lane applies its game's move (incremental bitboard update); a finished game is written back at once.
// Games end at different plies (mean 403, max ~850 in config 2), so a warp that kept its 32 games to the end would idle
// 30 % of its lane-iterations: every kPlayRound iterations the live games of the CTA are packed into the lowest warps
// through shared memory (23 words per game) and the emptied warps only wait at the barriers.  The loop key chain
// (rng <- split(rng, N + 1)[0], the same for every game) is produced a round ahead by one extra warp per CTA into a
// double-buffered shared ring, so a warp that ran empty and is refilled finds the current value there.
// A CTA that loaded a non-canonical game (see madn_fast.cuh) runs the generic rules of madn_core.cuh instead.
#ifdef DOGSTEP_TRACE
__device__ unsigned long long g_play_trace[128];
#endif
constexpr int kPlayRound = 64;  // 24 / 32 / 48 / 64: 2.21 / 2.19 / 2.18 / 2.175 ms on config 2 (with key ring and tail mode)
constexpr int kPlayMaxThreads = 512;  // game threads per CTA (+ 32: the producer warp)
constexpr int kXWords = 23;  // occ 8, pins 4, action set 8, cur|reward, len, game index

constexpr int kRingMax = 64;  // longest round the key ring holds
// `threads` = game threads (the CTA has one more warp, the key-chain producer)
static size_t play_smem_bytes(int threads) {
  return (size_t)(threads / 32) * (24 * 32 * 2 + 32 * 4 + 32 * 8) + (size_t)kXWords * threads * 4 + 2 * 34 * 4 +
         2 * (kRingMax + 2) * 8;
}

template <uint32_t CT>
__global__ void __launch_bounds__(kPlayMaxThreads + 32) k_madn_det_play_cta(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                      Key2 rng0, int64_t game_offset, int max_steps,
                                                                      int32_t* __restrict__ game_len,
                                                                      unsigned long long* __restrict__ total_steps, int round_len,
                                                                      int games_per_cta) {
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
  const uint32_t FULL = 0xFFFFFFFFu;
  const int64_t cta_base = (int64_t)blockIdx.x * games_per_cta;  // games_per_cta <= T: the games are spread over ALL SMs
  int gi = threadIdx.x;  // game held by this lane, relative to cta_base
  const RuleSet<CT> R{g.rules};
  int len = 0;
  unsigned steps_done = 0;
  MadnRegs s;
  bool alive = false, canon = true;
  if (producer) {
    // The loop key chain rng_{t+1} = split(rng_t, N + 1)[0] (game_agent.py:60) is the same for every game: one warp per CTA
    // computes it, a round ahead, into a double-buffered ring (ring[i] = rng_{t0 + i}, i = 0..round_len + 1) instead of
    // every lane of every warp repeating it each iteration (it was one Threefry in seven, and sat on the critical path
    // of the next step key).
    Key2 r = rng0;
    if (lane == 0) s_ring[0] = make_uint2(r.a, r.b);
    for (int i = 1; i <= round_len + 1; ++i) {
      r = split_i(r, 0u);
      if (lane == 0) s_ring[i] = make_uint2(r.a, r.b);
    }
  } else if (gi < games_per_cta && cta_base + gi < n) {
    load_state<true>(g, p, cta_base + gi, s);
    canon = is_canonical4(s, s.occ);
    alive = !s.done;
    if (!alive && game_len) game_len[cta_base + gi] = 0;
  }
  const bool fast = !__syncthreads_or(!canon);
  const uint32_t lane_hi = (uint32_t)lane << 8;
  if (!producer) s_best[lane] = 0u;
  int t = 0, round = 0;
  bool tail = false;  // CTA-uniform: one game per WARP from here on (see below)
  while (true) {
    // ---- compaction point
    // (counts and ring are double-buffered by round parity: an empty warp can reach the next point while others still read)
    const uint32_t ab = __ballot_sync(FULL, alive);
    const int par = round & 1;
    uint32_t* cnt = s_cnt + par * 34;
    if (lane == 0 && !producer) cnt[warp] = tail ? (ab != 0u) : (uint32_t)__popc(ab);
    ++round;
    __syncthreads();
    int before = 0, live = 0, nonempty = 0;
    for (int w = 0; w < W; ++w) {
      const int c = (int)cnt[w];
      before += (w < warp) ? c : 0;
      live += c;
      nonempty += (c > 0);
    }
#ifdef DOGSTEP_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0 && round <= 64) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      g_play_trace[2 * (round - 1)] = now;
      g_play_trace[2 * (round - 1) + 1] = (unsigned long long)live | ((unsigned long long)nonempty << 32);
    }
#endif
    if (live == 0 || t >= max_steps) break;
    // Tail: once the CTA is down to one game per warp or fewer, the lockstep iteration is bound by the latency of ONE warp's
    // instruction stream (~1.2 us), not by issue slots.  From then on every warp holds ONE game, replicated in all lanes:
    // lane a draws action a's Threefry bits one iteration ahead (they do not depend on the state), the legal mask and
    // the move are computed redundantly by all lanes, and the argmax is one redux — no scan, no item list, no atomics.
    const bool enter_tail = fast && !tail && live <= W;
    if (enter_tail || (!tail && ((live + 31) >> 5) < nonempty)) {  // CTA-uniform: packing frees at least one warp
      if (alive) {
        const int slot = before + __popc(ab & ((1u << lane) - 1u));
        uint32_t* x = s_x + slot;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          x[(2 * q) * T] = (uint32_t)s.occ[q];
          x[(2 * q + 1) * T] = (uint32_t)(s.occ[q] >> 32);
          x[(8 + q) * T] = s.pins[q];
          x[(12 + 2 * q) * T] = (uint32_t)s.as[q];
          x[(13 + 2 * q) * T] = (uint32_t)(s.as[q] >> 32);
        }
        x[20 * T] = (uint32_t)(s.cur & 0xFF) | ((uint32_t)(s.reward & 0xFF) << 8);
        x[21 * T] = (uint32_t)len;
        x[22 * T] = (uint32_t)gi;
      }
      __syncthreads();
      tail = tail || enter_tail;
      alive = (tail ? warp : (int)threadIdx.x) < live && !producer;
      if (alive) {
        const uint32_t* x = s_x + (tail ? warp : (int)threadIdx.x);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          s.occ[q] = (uint64_t)x[(2 * q) * T] | ((uint64_t)x[(2 * q + 1) * T] << 32);
          s.pins[q] = x[(8 + q) * T];
          s.as[q] = (uint64_t)x[(12 + 2 * q) * T] | ((uint64_t)x[(13 + 2 * q) * T] << 32);
        }
        const uint32_t misc = x[20 * T];
        s.cur = (int)(int8_t)(misc & 0xFFu);
        s.reward = (int)(int8_t)((misc >> 8) & 0xFFu);
        s.done = 0;
        s.die = 0;
        len = (int)x[21 * T];
        gi = (int)x[22 * T];
      }
      // the next write to s_x happens after the next round's first barrier, i.e. after every read above
    }
    // ---- kPlayRound lockstep iterations
    const int tend = min(t + round_len, max_steps);
    const uint2* rk = s_ring + par * (kRingMax + 2);  // rk[i] = the loop key of iteration t + i
    if (producer) {  // next round's keys
      uint2* nk = s_ring + (par ^ 1) * (kRingMax + 2);
      const uint2 r0 = rk[round_len];
      Key2 r{r0.x, r0.y};
      if (lane == 0) nk[0] = r0;
      for (int i = 1; i <= round_len + 1; ++i) {
        r = split_i(r, 0u);
        if (lane == 0) nk[i] = make_uint2(r.a, r.b);
      }
      t = tend;
      continue;
    }
    const uint32_t my = (uint32_t)(game_offset + cta_base + gi + 1);
    Key2 key = split_i(Key2{rk[0].x, rk[0].y}, my);  // split(rng, N+1)[j+1]   (game_agent.py:60 / evaluate_agent.py:336)
    int ri = 0;
    if (tail) {
      uint32_t mant = bits_i(key, (uint32_t)lane) >> 9;  // lane a: the draw of action a (lanes >= 24 unused)
      Key2 key1 = split_i(Key2{rk[1].x, rk[1].y}, my);   // step key of the NEXT iteration
      const uint32_t tag = lane < 24 ? 23u - (uint32_t)lane : 0u;
#pragma unroll 1
      for (; t < tend && alive; ++t) {
        ++ri;
        // two independent Threefry chains per iteration, both with inputs known at the top of the loop: the draws of
        // iteration t + 1 (from its key) and the key of iteration t + 2 — they overlap each other and the state-dependent
        // chain below (mask -> argmax -> move)
        const uint2 rn = rk[ri + 1];
        const uint32_t mant_next = bits_i(key1, (uint32_t)lane) >> 9;
        const Key2 key2 = split_i(Key2{rn.x, rn.y}, my);
        int cp = 0;
        const uint32_t m = det_valid_mask4(R, g, s, cp);
        if (m) {
          const uint32_t v = ((m >> lane) & 1u) ? ((mant << 5) | tag) : 0u;  // m has 24 bits: lanes >= 24 contribute 0
          const int a = 23 - (int)(__reduce_max_sync(FULL, v) & 31u);
          det_step4(R, s, cp, a);
        } else {
          det_no_step4(s);
        }
        ++len;
        steps_done += (lane == 0);
        if (s.done) {
          alive = false;
          if (lane == 0) {
            store_det_all(g, p, cta_base + gi, s);
            if (game_len) game_len[cta_base + gi] = len;
          }
        }
        mant = mant_next;
        key1 = key2;
      }
      t = tend;
      continue;
    }
#pragma unroll 1
    for (; t < tend; ++t) {
      if (!__any_sync(FULL, alive)) break;
      ++ri;
      const uint2 rn = rk[ri];  // split(rng, N+1)[0] of this iteration = the loop key of the next
      int cp = 0;
      uint32_t m = 0u;
      if (alive) m = fast ? det_valid_mask4(R, g, s, cp) : madn_det_valid_mask(g, s);
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
          if (fast) det_step4(R, s, cp, a);
          else madn_det_step(g, s, a / 6, a % 6 + 1, m);  // map_action (deterministic_madn.py:469-479)
        } else {
          if (fast) det_no_step4(s);
          else madn_det_no_step(g, s);
        }
        ++len;
        ++steps_done;
        if (s.done) {  // finished: write the game back now, the lane is free from here on
          alive = false;
          store_det_all(g, p, cta_base + gi, s);
          if (game_len) game_len[cta_base + gi] = len;
        }
      }
      s_best[lane] = 0u;
      key = key_next;
      __syncwarp();
    }
    t = tend;
  }
  if (alive && (!tail || lane == 0)) {  // max_steps reached with the game still running
    store_det_all(g, p, cta_base + gi, s);
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
constexpr uint32_t kTrainRules = DOGSTEP_RULE_TEAMS | DOGSTEP_RULE_INITIAL_FREE_PIN | DOGSTEP_RULE_JUMP_IN_GOAL |
                                 DOGSTEP_RULE_START_ON_1 | DOGSTEP_RULE_BONUS_TURN_ON_6;

__global__ void __launch_bounds__(kThreads) k_madn_cls_throw_die(const __grid_constant__ MadnGeom g, MadnPtrs p, int64_t n,
                                                                 float* __restrict__ probs, int write_die, int only_active) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  if (only_active && p.done[i] != 0) return;
  MadnRegs s;
  load_state<false>(g, p, i, s);
  // dice_probabilities (classic_madn.py:14-18,208-228): float32 literals rounded from doubles
  float pr[6];
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
  if (probs) {
#pragma unroll
    for (int k = 0; k < 6; ++k) probs[i * 6 + k] = pr[k];
  }
  if (write_die) {
    Key2 key{p.key[2 * i], p.key[2 * i + 1]};
    Key2 knew = split_i(key, 0), sub = split_i(key, 1);
    p.die[i] = (int8_t)(choice6(sub, pr) + 1);
    p.key[2 * i] = knew.a;
    p.key[2 * i + 1] = knew.b;
  }
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
  if (g.n == 4 && g.d == 10) {
    // the games are spread evenly over the SMs (config 2: 65,536 games -> 148 CTAs of 443), at most 512 per CTA
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    int64_t per = (n + sms - 1) / sms;
    if (per > kPlayMaxThreads) per = kPlayMaxThreads;
    const int gpc = (int)per;                          // games per CTA (config 2: 443 on each of the 148 SMs)
    const int threads = ((gpc + 31) / 32) * 32;
    const size_t smem = play_smem_bytes(threads);
    const unsigned blocks = blocks_for(n, gpc);
    const int round_len = kPlayRound;  // <= kRingMax
    if (g.rules == kTrainRules) {
      cudaFuncSetAttribute(k_madn_det_play_cta<kTrainRules>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_madn_det_play_cta<kTrainRules><<<blocks, threads + 32, smem, st>>>(g, p, n, rng, game_offset, max_steps, game_len, total_steps, round_len, gpc);
    } else {
      cudaFuncSetAttribute(k_madn_det_play_cta<kRulesRuntime>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_madn_det_play_cta<kRulesRuntime><<<blocks, threads + 32, smem, st>>>(g, p, n, rng, game_offset, max_steps, game_len, total_steps, round_len, gpc);
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
    for (int r = 0; r < 64 && h[2 * r]; ++r)
      fprintf(stderr, "round %2d (%d)  +%8.1f us  live %5llu  warps %3llu\n", r, r, (h[2 * r] - h[0]) / 1e3,
              h[2 * r + 1] & 0xFFFFFFFFull, h[2 * r + 1] >> 32);
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
