// dog_kernels.cu — CUDA kernels (sm_100a) + C-ABI for the DOG environment.  One warp per game,
// four games per 128-thread CTA; see dog_core.cuh for the per-warp shared record and the rules.
// HBM layout = batched leaves of the reference `DOG` pytree (DOG/dog.py:31-56), game axis leading.
#include <atomic>
#include <cstdint>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../include/dogstep.h"
#include "common.cuh"
#include "jaxrand.cuh"
#include "dog_core.cuh"
#include "dog_fast.cuh"

namespace dogstep {

constexpr int kDogThreads = 128;
constexpr int kDogWarps = kDogThreads / 32;

struct DogPtrs {
  int8_t* board;
  int8_t* cur;
  int32_t* pins;
  int8_t* reward;
  uint8_t* done;
  int8_t* deck;
  int8_t* hands;
  int8_t* swap_choices;
  int8_t* round_starter;
  int8_t* phase;
  uint32_t* key;
  int8_t* hand_size;
};

__device__ __forceinline__ void dog_load(const DogGeom& g, const DogPtrs& p, int64_t i, DogS& s, int lane) {
  for (int k = lane; k < 64; k += 32) s.board[k] = (k < g.total) ? p.board[i * g.total + k] : (int8_t)-1;
  if (lane < 16) s.pins[lane >> 2][lane & 3] = (lane < g.n * 4) ? p.pins[i * g.n * 4 + lane] : -1;
  for (int k = lane; k < 64; k += 32) {
    int q = k >> 4, c = k & 15;
    s.hands[q][c] = (q < g.n && c < kNCard) ? p.hands[(i * g.n + q) * kNCard + c] : (int8_t)0;
  }
  if (lane < 16) s.deck[lane] = (lane < kNCard) ? p.deck[i * kNCard + lane] : (int8_t)0;
  if (lane < 4) s.swap_choices[lane] = p.swap_choices[i * 4 + lane];
  if (lane < 2) s.key[lane] = p.key[i * 2 + lane];
  if (lane == 0) {
    s.cur = p.cur[i];
    s.reward = p.reward[i];
    s.done = p.done[i] != 0;
    s.round_starter = p.round_starter[i];
    s.phase = p.phase[i];
    s.hand_size = p.hand_size[i];
  }
  __syncwarp();
  // canonical 4-player / distance-10 records take the bitboard rules of dog_fast.cuh (warp-uniform flag)
  const bool fast = g.n == 4 && g.d == 10 && dog4_canonical_warp(s, lane);
  if (lane == 0) { s.scratch[7] = fast; s.scratch[6] = 0; }
  __syncwarp();
}

__device__ __forceinline__ void dog_build_mask_any(const DogGeom& g, DogS& s, int lane) {
  if (s.scratch[7]) dog4_build_mask(dg4_rules(g.rules), s, lane);
  else dog_build_mask(g, s, lane);
}

template <bool TRUSTED = false, bool LAZY_BOARD = false>
__device__ __forceinline__ void dog_env_step_any(const DogGeom& g, const Dog4Rules& R4, DogS& s, int lane, int action, int& r, int& d) {
  if (s.scratch[7]) dog4_env_step<TRUSTED, LAZY_BOARD>(R4, g, s, lane, action, r, d);
  else dog_env_step(g, s, lane, action, r, d);
}

__device__ __forceinline__ void dog_store(const DogGeom& g, const DogPtrs& p, int64_t i, const DogS& s, int lane) {
  __syncwarp();
  for (int k = lane; k < g.total; k += 32) p.board[i * g.total + k] = s.board[k];
  if (lane < g.n * 4) p.pins[i * g.n * 4 + lane] = s.pins[lane >> 2][lane & 3];
  for (int k = lane; k < 64; k += 32) {
    int q = k >> 4, c = k & 15;
    if (q < g.n && c < kNCard) p.hands[(i * g.n + q) * kNCard + c] = s.hands[q][c];
  }
  if (lane < kNCard) p.deck[i * kNCard + lane] = s.deck[lane];
  if (lane < 4) p.swap_choices[i * 4 + lane] = s.swap_choices[lane];
  if (lane < 2) p.key[i * 2 + lane] = s.key[lane];
  if (lane == 0) {
    p.cur[i] = (int8_t)s.cur;
    p.reward[i] = (int8_t)s.reward;
    p.done[i] = (uint8_t)s.done;
    p.round_starter[i] = (int8_t)s.round_starter;
    p.phase[i] = (int8_t)s.phase;
    p.hand_size[i] = (int8_t)s.hand_size;
  }
}

// categorical(key, where(mask, 0, -1e9)) over the legal mask in s.mask: legal actions are compacted into
// s.items, the Threefry draws are dealt evenly to the lanes, the winner is the FIRST action with the largest
// 23-bit uniform mantissa (== argmax of logits + gumbel).  Returns -1 if nothing is legal.
__device__ __forceinline__ int dog_categorical(const DogGeom& g, DogS& s, int lane, Key2 key) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const int nwords = (g.num_actions + 31) >> 5;
  uint32_t w = (lane < nwords) ? s.mask[lane] : 0u;
  int cnt = __popc(w), incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(FULL, incl, 31);
  if (total == 0) return -1;
  int off = incl - cnt;
  for (uint32_t m = w; m; m &= m - 1) s.items[off++] = (uint16_t)(lane * 32 + __ffs(m) - 1);
  __syncwarp();
  uint32_t best_m = 0;
  int best_a = 0x7FFFFFFF;
  for (int j = lane; j < total; j += 32) {
    int a = s.items[j];
    uint32_t m = bits_i(key, (uint32_t)a) >> 9;
    if (best_a == 0x7FFFFFFF || m > best_m) { best_m = m; best_a = a; }  // items ascend, so ties keep the first
  }
  // largest mantissa, lowest action among equals: two redux instead of a five-step shuffle butterfly
  const uint32_t top = __reduce_max_sync(FULL, best_a == 0x7FFFFFFF ? 0u : best_m);
  const int first = (int)__reduce_min_sync(FULL, (best_a != 0x7FFFFFFF && best_m == top) ? (uint32_t)best_a : 0x7FFFFFFFu);
  __syncwarp();
  return first;
}

// dog_categorical for the persistent play kernel, with the NEXT turn's two key derivations riding in the idle lanes of the
// last item pass (a game has 10-40 legal actions, so the last pass of 32 usually has room): `key` is this turn's step key,
// `rng1` the loop key of the next turn; on return key = split(rng1, N + 1)[g + 1] (the next step key) and
// rng1 = split(rng1, N + 1)[0] (the loop key after that).  Saves the separate Threefry pass per turn.
__device__ __forceinline__ int dog_categorical_pipelined(const DogGeom& g, DogS& s, int lane, Key2& key, Key2& rng1, uint32_t my) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const int nwords = (g.num_actions + 31) >> 5;
  const uint32_t w = (lane < nwords) ? s.mask[lane] : 0u;
  const int cnt = __popc(w);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(FULL, incl, 31);
  int off = incl - cnt;
  for (uint32_t m = w; m; m &= m - 1) s.items[off++] = (uint16_t)(lane * 32 + __ffs(m) - 1);
  __syncwarp();
  // slots total and total + 1 of the padded item list carry the two key derivations
  const int slots = total + 2;
  uint32_t best_m = 0;
  int best_a = 0x7FFFFFFF;
  Key2 nkey{0u, 0u}, nrng{0u, 0u};
  for (int base = 0; base < slots; base += 32) {
    const int j = base + lane;
    const bool item = j < total, kslot = j == total, rslot = j == total + 1;
    const int a = item ? (int)s.items[j] : 0;
    const Key2 k = item ? key : rng1;
    const uint32_t c = item ? (uint32_t)a : (kslot ? my : 0u);
    const Key2 o = threefry2x32(k, 0u, c);
    if (item) {
      const uint32_t m = (o.a ^ o.b) >> 9;
      if (best_a == 0x7FFFFFFF || m > best_m) { best_m = m; best_a = a; }  // items ascend, so ties keep the first
    }
    const uint32_t kb = __ballot_sync(FULL, kslot), rb = __ballot_sync(FULL, rslot);
    if (kb) { const int src = __ffs(kb) - 1; nkey = Key2{__shfl_sync(FULL, o.a, src), __shfl_sync(FULL, o.b, src)}; }
    if (rb) { const int src = __ffs(rb) - 1; nrng = Key2{__shfl_sync(FULL, o.a, src), __shfl_sync(FULL, o.b, src)}; }
  }
  key = nkey;
  rng1 = nrng;
  if (total == 0) return -1;
  const uint32_t top = __reduce_max_sync(FULL, best_a == 0x7FFFFFFF ? 0u : best_m);
  const int first = (int)__reduce_min_sync(FULL, (best_a != 0x7FFFFFFF && best_m == top) ? (uint32_t)best_a : 0x7FFFFFFFu);
  __syncwarp();
  return first;
}

#define DOG_KERNEL_PROLOGUE                                            \
  __shared__ DogS sh[kDogWarps];                                       \
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;          \
  const int64_t i = (int64_t)blockIdx.x * kDogWarps + warp;            \
  DogS& s = sh[warp];

__global__ void __launch_bounds__(kDogThreads) k_dog_reset(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                           const int32_t* __restrict__ seeds, int starting_player) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  if (lane == 0) {  // env_reset (dog.py:83-181)
    Key2 k0 = prng_key(seeds[i]);
    Key2 knew = split_i(k0, 0), sub = split_i(k0, 1);
    int sp = starting_player;
    if (sp < 0 || sp >= g.n) sp = randint_i(sub, 0, 0, g.n);
    for (int q = 0; q < 4; ++q)
      for (int k = 0; k < 4; ++k) s.pins[q][k] = -1;
    if (DG_RULE(g, DOGSTEP_RULE_INITIAL_FREE_PIN))
      for (int q = 0; q < g.n; ++q) s.pins[q][0] = g.start[q];
    dog_set_pins_on_board(g, s.pins, s.board);
    for (int q = 0; q < 4; ++q)
      for (int k = 0; k < 16; ++k) s.hands[q][k] = 0;
    for (int k = 0; k < 16; ++k) s.deck[k] = (k < kNCard) ? 8 : 0;
    s.deck[0] = 6;
    for (int q = 0; q < 4; ++q) s.swap_choices[q] = -1;
    s.cur = sp;
    s.reward = 0;
    s.done = 0;
    s.round_starter = -1;
    s.phase = 0;
    s.hand_size = 6;
    s.key[0] = knew.a;
    s.key[1] = knew.b;
  }
  __syncwarp();
  dog_distribute_cards(g, s, lane);
  dog_store(g, p, i, s, lane);
}

__global__ void __launch_bounds__(kDogThreads) k_dog_valid_actions(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                                   uint8_t* __restrict__ mask) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  dog_load(g, p, i, s, lane);
  dog_build_mask_any(g, s, lane);
  uint8_t* out = mask + i * g.num_actions;
  for (int a = lane; a < g.num_actions; a += 32) out[a] = (uint8_t)((s.mask[a >> 5] >> (a & 31)) & 1u);
}

// DOG observation encoder — DESIGN WORK, no reference counterpart: DOG/dog.py:1264-1272 leaves encode_board as a TODO and the
// DOG networks are stubs (MuZero_DOG/muzero_dog.py:85-99).  The layout continues the MADN encoders the reference does have
// (deterministic_madn.py:395-438: everything rolled into the mover's frame, scalar facts broadcast over a plane), restricted
// to what the seat to move may know (its own hand, not the others'; hand SIZES are public).  int8 [8 + 3n + 14, total]:
//   [0, n)        occupancy of seat (cur + r) % n, ring rolled so that the mover's start is cell 0, goal lanes rotated alike
//   n, n + 1      own side / other side sums of those planes (teams: seats r even / odd; else mover / everyone else)
//   [n+2, 2n+2)   pins at home of seat (cur + r) % n
//   [2n+2, 2n+16) the mover's own hand: count of card c (0 joker, 1 swap, 2..13)
//   [2n+16, 3n+16) cards in hand of seat (cur + r) % n (a count: public)
//   3n+16 ..      phase (1 = partner swap), hand_size of the NEXT deal, (round_starter - cur) mod n or -1, the card the mover has
//                 put aside for its partner + 1 (0 = none), and two planes of zeros reserved for the discard pile
__global__ void __launch_bounds__(kDogThreads) k_dog_encode_board(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                                  int8_t* __restrict__ obs) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  dog_load(g, p, i, s, lane);
  const int N = g.n, T = g.total, bs = g.bs;
  const int cur = d_gidx(s.cur, N);
  const int planes = 8 + 3 * N + kNCard;
  int8_t* out = obs + i * (int64_t)planes * T;
  // per-seat scalars
  int home[4], hsz[4];
  for (int r = 0; r < N; ++r) {
    const int q = (cur + r) % N;
    int h = 0, c = 0;
    for (int k = 0; k < 4; ++k) h += s.pins[q][k] == -1;
    for (int k = 0; k < kNCard; ++k) c += s.hands[q][k];
    home[r] = h;
    hsz[r] = c;
  }
  const bool teams = DG_RULE(g, DOGSTEP_RULE_TEAMS);
  const int rs_rel = s.round_starter < 0 ? -1 : d_fmod(s.round_starter - cur, N);
  for (int e = lane; e < planes * T; e += 32) {
    const int plane = e / T, k = e - plane * T;
    const int src = k < bs ? (k + g.d * cur) % bs : bs + (k - bs + 4 * cur) % 16;   // jnp.roll(x, -shift)[k] = x[(k + shift) % len]
    const int owner = s.board[src];                                                   // -1 or seat
    const int rel = owner < 0 ? -1 : d_fmod(owner - cur, N);
    int v;
    if (plane < N) v = rel == plane;
    else if (plane == N) v = rel >= 0 && (teams ? (rel % 2 == 0) : rel == 0);
    else if (plane == N + 1) v = rel >= 0 && (teams ? (rel % 2 == 1) : rel != 0);
    else if (plane < 2 * N + 2) v = home[plane - N - 2];
    else if (plane < 2 * N + 2 + kNCard) v = s.hands[cur][plane - 2 * N - 2];
    else if (plane < 3 * N + 2 + kNCard) v = hsz[plane - 2 * N - 2 - kNCard];
    else {
      const int m = plane - (3 * N + 2 + kNCard);
      v = m == 0 ? s.phase : m == 1 ? s.hand_size : m == 2 ? rs_rel : m == 3 ? s.swap_choices[cur] + 1 : 0;
    }
    out[e] = (int8_t)v;
  }
}

__global__ void __launch_bounds__(kDogThreads) k_dog_step(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                          const int32_t* __restrict__ action, int8_t* __restrict__ reward,
                                                          uint8_t* __restrict__ done) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  dog_load(g, p, i, s, lane);
  int r, d;
  dog_env_step_any(g, dg4_rules(g.rules), s, lane, action[i], r, d);
  dog_store(g, p, i, s, lane);
  if (lane == 0) {
    if (reward) reward[i] = (int8_t)r;
    if (done) done[i] = (uint8_t)d;
  }
}

__global__ void __launch_bounds__(kDogThreads) k_dog_no_step(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                             int8_t* __restrict__ reward, uint8_t* __restrict__ done) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  dog_load(g, p, i, s, lane);
  dog_no_step(g, s, lane);
  dog_store(g, p, i, s, lane);
  if (lane == 0) {
    if (reward) reward[i] = 0;
    if (done) done[i] = (uint8_t)s.done;
  }
}

__global__ void __launch_bounds__(kDogThreads) k_dog_distribute_cards(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  dog_load(g, p, i, s, lane);
  dog_distribute_cards(g, s, lane);
  dog_store(g, p, i, s, lane);
}

// the reference's module-level sub-steps (dog.py:755, 790, 861, 913); env left untouched
__global__ void __launch_bounds__(kDogThreads) k_dog_substep(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                             const int32_t* __restrict__ kind, const int32_t* __restrict__ args,
                                                             int8_t* __restrict__ board_out, int32_t* __restrict__ pins_out,
                                                             int8_t* __restrict__ reward, uint8_t* __restrict__ done) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  dog_load(g, p, i, s, lane);
  if (lane == 0) {
    const int32_t* a = args + i * 4;
    int r = 0, d = 0;
    int k = kind[i];
    if (k == 0) dog_step_normal(g, s, a[0], a[1], r, d);
    else if (k == 1) dog_step_neg(g, s, a[0], a[1], r, d);
    else if (k == 2) dog_step_swap(g, s, a[0], a[1], r, d);
    else {
      int dist[4] = {a[0], a[1], a[2], a[3]};
      dog_step_hot7(g, s, dist, r, d);
    }
    reward[i] = (int8_t)r;
    done[i] = (uint8_t)d;
  }
  __syncwarp();
  for (int k = lane; k < g.total; k += 32) board_out[i * g.total + k] = s.board[k];
  if (lane < g.n * 4) pins_out[i * g.n * 4 + lane] = s.pins[lane >> 2][lane & 3];
}

// ---- self-play bookkeeping for DOG (BASELINE config 5): one lockstep iteration after the search ------------------------
// The reference has no DOG self-play loop (MuZero_DOG/muzero_dog.py:85-99 are stubs), so this is the det-MADN loop
// (MuZero_det_MADN/game_agent.py:64-148, do_active_step) applied to the DOG env: env_step(action) if a legal action
// exists else no_step, reward / discount class targets and the trajectory row at traj.episode_lengths[g].  One warp per
// game; the 806-wide policy row and the observation row are copied by all lanes.
__global__ void __launch_bounds__(kDogThreads) k_dog_agent_step(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                                const int32_t* __restrict__ action,
                                                                const float* __restrict__ root_value,
                                                                const float* __restrict__ weights, const int8_t* __restrict__ obs,
                                                                dogstep_replay_arrays tr) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  if (p.done[i] != 0) return;  // do_skip_step: finished games are left untouched (warp-uniform)
  dog_load(g, p, i, s, lane);
  dog_build_mask_any(g, s, lane);
  __syncwarp();
  uint32_t anyw = 0u;
  for (int w = lane; w < kDogMaskWords; w += 32) anyw |= s.mask[w];
  const int has_valid = __any_sync(0xFFFFFFFFu, anyw != 0u);
  const int teams = DG_RULE(g, DOGSTEP_RULE_TEAMS);
  const int pid = s.cur;
  const int team_before = teams ? (((pid % 2) + 2) % 2) : -1;
  const int idx = tr.episode_lengths[i];
  __syncwarp();
  int act = -1, rew_t = 1, disc_t = 1;
  if (has_valid) {
    int r, d;
    act = action[i];
    dog_env_step_any(g, dg4_rules(g.rules), s, lane, act, r, d);
    __syncwarp();
    const int next = s.cur;
    const int next_team = teams ? (((next % 2) + 2) % 2) : -1;
    rew_t = (d && r > 0) ? 2 : ((d && r < 0) ? 0 : 1);
    disc_t = d ? 1 : (teams ? (team_before == next_team ? 2 : 0) : (pid == next ? 2 : 0));
  } else {
    dog_no_step(g, s, lane);
  }
  dog_store(g, p, i, s, lane);
  if (lane == 0) tr.episode_lengths[i] = idx + 1;
  if (idx < 0 || idx >= tr.max_episode_length) return;  // .at[idx].set drops out-of-range rows
  const int64_t r = i * tr.max_episode_length + idx;
  if (lane == 0) {
    tr.actions[r] = act;
    tr.rewards[r] = rew_t;
    tr.root_values[r] = has_valid ? root_value[i] : 0.0f;
    tr.masks[r] = has_valid ? 1.0f : 0.0f;
    tr.players[r] = pid;
    tr.teams[r] = team_before;
    tr.discounts[r] = disc_t;
  }
  const int A = tr.action_dim;
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

__device__ __forceinline__ void dog_random_turn(const DogGeom& g, DogS& s, int lane, Key2 key) {
  dog_build_mask_any(g, s, lane);
  int a = dog_categorical(g, s, lane, key);
  if (a >= 0) {
    int r, d;
    dog_env_step_any<true>(g, dg4_rules(g.rules), s, lane, a, r, d);  // drawn from the mask just built: no second validation
  } else {
    dog_no_step(g, s, lane);
  }
}

__global__ void __launch_bounds__(kDogThreads) k_dog_random_step(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                                 Key2 rng, int64_t game_offset,
                                                                 unsigned long long* __restrict__ active_count) {
  DOG_KERNEL_PROLOGUE
  if (i >= n) return;
  if (p.done[i] != 0) return;  // warp-uniform
  dog_load(g, p, i, s, lane);
  dog_random_turn(g, s, lane, split_i(rng, (uint32_t)(game_offset + i + 1)));
  dog_store(g, p, i, s, lane);
  if (active_count && lane == 0) atomicAdd(active_count, 1ull);
}

// Persistent random-policy play, phase-synchronous: one CTA of 32 warps per SM, every warp owns one game at a time
// (game index strided over all warps of the grid, a finished game is written back and the next one loaded), and the
// two halves of a turn — legal mask | categorical draw + state transition — are separated by CTA barriers.
// Why: with free-running warps this kernel is bound by instruction fetch (ncu r1k: the SM instruction cache hits 53 %,
// the GPC-level instruction cache runs at 92 % of its request rate) because 32 resident warps walk 6 k distinct hot
// SASS instructions at 32 different places.  With the phases aligned the SM fetches each phase's code once per turn
// instead of once per warp (ncu r1o: 85 % hits, no_instruction stalls 18 -> 0.7 per issue); what remains is the
// imbalance between warps inside a phase (barrier stalls).
constexpr int kSyncWarps = 32;

// Games are handed out dynamically: the first `first` games go to the warps by index, every further one to whichever warp
// finishes first (one atomic per game).  DOG games last 222..2,000 plies (mean 822): with static striding the warp with the
// longest three or four games ran 4,999 plies against a mean of 2,843 (57 % utilisation), with the queue 4,066 (70 %).  A
// game's randomness depends on its index only, so the assignment does not change any result.
__device__ __forceinline__ int64_t dog_next_game(unsigned int* queue, int64_t first, int lane) {
  unsigned int k = 0;
  if (lane == 0) k = atomicAdd(queue, 1u);
  k = __shfl_sync(0xFFFFFFFFu, k, 0);
  return first + (int64_t)k;
}

#ifdef DOGSTEP_TRACE
__device__ unsigned long long g_dog_trace[1024];  // CTA 0, every 64th turn: time, live warps
__device__ unsigned long long g_dog_solo[8];      // CTA 0 with ONE live game: cycles in flags / tasks / draw / transition / turn, turns
#endif
__global__ void __launch_bounds__(kSyncWarps * 32, 1) k_dog_play_random(const __grid_constant__ DogGeom g, DogPtrs p, int64_t n,
                                                                      Key2 rng0, int64_t game_offset, int max_steps,
                                                                      int32_t* __restrict__ game_len,
                                                                      unsigned long long* __restrict__ total_steps,
                                                                      unsigned int* __restrict__ queue) {
  extern __shared__ __align__(16) unsigned char dog_smem_raw[];
  DogS* sh = reinterpret_cast<DogS*>(dog_smem_raw);
  // hot-seven task queues, double-buffered by turn parity: [head, tail, 128 x (slot * 4 + chunk)]
  int* s_q = reinterpret_cast<int*>(dog_smem_raw + sizeof(DogS) * kSyncWarps);
  const Dog4Rules R4 = dg4_rules(g.rules);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DogS& s = sh[warp];
  int turn = 0;
  if (threadIdx.x < 4) s_q[(threadIdx.x >> 1) * 130 + (threadIdx.x & 1)] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * kSyncWarps;
  int64_t i = (int64_t)blockIdx.x * kSyncWarps + warp;
  bool have = false, dealing = false;  // dealing: the game sits this turn out and is dealt (see below)
  int len = 0;
  unsigned long long steps = 0;
  Key2 rng = rng0, key = rng0;  // rng: loop key of the NEXT turn, key: step key of this turn (see dog_categorical_pipelined)
  while (true) {
    while (!have && i < n) {  // next game of this warp (games that are already over cost nothing)
      dog_load(g, p, i, s, lane);
      if (s.done || max_steps <= 0) {
        if (lane == 0 && game_len) game_len[i] = 0;
        __syncwarp();
        i = dog_next_game(queue, stride, lane);
      } else {
        have = true;
        len = 0;
        // first step key and the loop key of the second turn, in one Threefry pass (lane 0: split(rng0, N + 1)[0])
        const Key2 both = split_i(rng0, lane == 0 ? 0u : (uint32_t)(game_offset + i + 1));
        key = Key2{__shfl_sync(0xFFFFFFFFu, both.a, 1), __shfl_sync(0xFFFFFFFFu, both.b, 1)};
        rng = Key2{__shfl_sync(0xFFFFFFFFu, both.a, 0), __shfl_sync(0xFFFFFFFFu, both.b, 0)};
      }
    }
    // phase A, legal masks.  A hand with a seven or a joker needs the 120 hot-seven splits on top of the moves every hand has
    // (11 k cycles for the slowest warp of a turn against a mean of 4.7 k): its four 32-split chunks go to a CTA queue and
    // are taken by whichever warp has finished its own mask.
    int* q = s_q + (turn & 1) * 130;
    int flags = 0;
    bool shared_mask = false;
#ifdef DOGSTEP_TRACE
    const long long tc0 = clock64();
#endif
    if (have && dealing) {
      // The previous transition ended a round.  A deal is 120 Threefry draws plus a 24-step selection — 11 k cycles against
      // 5 k for a plain transition, and with 26 live games one of them deals in four turns out of five — so it is not done
      // inline: the game sits THIS turn out, its four 32-slot draw passes go to the task queue next to the hot-seven
      // chunks, and its owner selects and deals while the other games draw and move.
      dog_deal_begin(g, s, lane);
      if (lane == 0) {
        const int pos = atomicAdd(&q[1], 4);
        for (int k = 0; k < 4; ++k) q[2 + pos + k] = 0x100 | (warp * 4 + k);
      }
    } else if (have) {
      shared_mask = s.phase == 0 && s.scratch[7] != 0;
      if (shared_mask) {
        for (int w = lane; w < kDogMaskWords; w += 32) s.mask[w] = 0u;
        flags = dog4_mask_flags(R4, s, lane);
        if (lane == 0 && (flags & 2)) {
          const int pos = atomicAdd(&q[1], 4);
          for (int k = 0; k < 4; ++k) q[2 + pos + k] = warp * 4 + k;
        }
      }
    }
#ifdef DOGSTEP_TRACE
    int live_tr = 0;
    {
      const int live = __syncthreads_count(have);
      live_tr = live / 32;
      if (blockIdx.x == 0 && threadIdx.x == 0 && (turn & 63) == 0 && (turn >> 6) < 511) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        g_dog_trace[2 * (turn >> 6)] = now;
        g_dog_trace[2 * (turn >> 6) + 1] = (unsigned long long)(live / 32);
      }
    }
#endif
    if (!__syncthreads_or(have)) break;
    if (threadIdx.x == 0) { int* qn = s_q + ((turn + 1) & 1) * 130; qn[0] = 0; qn[1] = 0; }  // next turn's queue (see parity note)
#ifdef DOGSTEP_TRACE
    const long long tc1 = clock64();
#endif
    if (have && !dealing) {
      if (shared_mask) {
        dog4_mask_task(R4, s, 4, lane);
        dog4_mask_task(R4, s, 5, lane);
        if (flags & 1) dog4_mask_task(R4, s, 6, lane);
      } else {
        dog_build_mask_any(g, s, lane);  // swap phase (14 card bits) or a non-canonical record: the owner alone
      }
    }
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&q[0], 1);
      t = __shfl_sync(0xFFFFFFFFu, t, 0);
      if (t >= *(volatile int*)&q[1]) break;
      const int task = q[2 + t];
      if (task & 0x100) dog_deal_draw(sh[(task & 0xFF) >> 2], task & 3, lane);
      else dog4_mask_task(R4, sh[task >> 2], task & 3, lane);
    }
    ++turn;
    __syncthreads();
#ifdef DOGSTEP_TRACE
    const long long tc2 = clock64();
    long long tc3 = tc2;
#endif
    if (have && dealing) {                                          // the draws are in the record: select, deal, next round
      dog_deal_finish(g, s, lane);
      dealing = false;
    } else if (have) {
      const int a = dog_categorical_pipelined(g, s, lane, key, rng, (uint32_t)(game_offset + i + 1));  // phase B
#ifdef DOGSTEP_TRACE
      tc3 = clock64();
#endif
      if (lane == 0) s.scratch[6] = 1;                              // phase C: transition, a deal that falls due is held back
      __syncwarp();
      if (a >= 0) {
        int r, d;
        dog_env_step_any<true, true>(g, R4, s, lane, a, r, d);  // drawn from the mask just built: no second validation
      } else {
        dog_no_step(g, s, lane);
      }
      ++len;
      ++steps;
      __syncwarp();
      const bool need_deal = s.scratch[0] != 0;
      const bool finished = s.done || len >= max_steps;
      __syncwarp();
      if (lane == 0) s.scratch[6] = 0;
      __syncwarp();
      if (finished) {
        if (need_deal) dog_distribute_cards(g, s, lane);            // the state written back is the reference's: dealt
        if (s.scratch[7]) dog4_rebuild_board_warp(s, lane);         // the board bytes were left stale by the lazy transitions
        dog_store(g, p, i, s, lane);
        if (lane == 0 && game_len) game_len[i] = len;
        __syncwarp();
        have = false;
        i = dog_next_game(queue, stride, lane);
      } else {
        dealing = need_deal;
      }
    }
#ifdef DOGSTEP_TRACE
    if (blockIdx.x == 0 && live_tr == 1 && have && lane == 0) {
      const long long tc4 = clock64();
      atomicAdd(&g_dog_solo[0], (unsigned long long)(tc1 - tc0));
      atomicAdd(&g_dog_solo[1], (unsigned long long)(tc2 - tc1));
      atomicAdd(&g_dog_solo[2], (unsigned long long)(tc3 - tc2));
      atomicAdd(&g_dog_solo[3], (unsigned long long)(tc4 - tc3));
      atomicAdd(&g_dog_solo[4], 1ull);
    }
#endif
  }
  if (lane == 0 && total_steps && steps) atomicAdd(total_steps, steps);
}

static inline unsigned dog_blocks(int64_t n) { return (unsigned)((n + kDogWarps - 1) / kDogWarps); }

static int dog_ptrs(const dogstep_dog_state* s, DogPtrs* p) {
  if (!s || !s->board || !s->current_player || !s->pins || !s->reward || !s->done || !s->deck || !s->hands ||
      !s->swap_choices || !s->round_starter || !s->phase || !s->key || !s->hand_size)
    return DOGSTEP_ERR_INVALID_ARG;
  *p = DogPtrs{s->board, s->current_player, s->pins, s->reward, s->done, s->deck, s->hands, s->swap_choices,
               s->round_starter, s->phase, s->key, s->hand_size};
  return DOGSTEP_OK;
}

#define DOG_PROLOGUE                                 \
  DogGeom g;                                         \
  DogPtrs p;                                         \
  if (n < 0) return DOGSTEP_ERR_INVALID_ARG;         \
  if (int rc = dog_make_geom(cfg, &g)) return rc;    \
  if (int rc = dog_ptrs(s, &p)) return rc;           \
  if (n == 0) return DOGSTEP_OK;                     \
  cudaStream_t st = (cudaStream_t)stream;

}  // namespace dogstep

using namespace dogstep;

extern "C" {

int dogstep_dog_num_actions(const dogstep_dog_cfg* cfg) {
  DogGeom g;
  if (int rc = dog_make_geom(cfg, &g)) return rc;
  return g.num_actions;
}

int dogstep_dog_reset(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* seeds,
                      int32_t starting_player, void* stream) {
  DOG_PROLOGUE
  if (!seeds) return DOGSTEP_ERR_INVALID_ARG;
  k_dog_reset<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, seeds, starting_player);
  return check_launch();
}

int dogstep_dog_valid_actions(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, uint8_t* mask, void* stream) {
  DOG_PROLOGUE
  if (!mask) return DOGSTEP_ERR_INVALID_ARG;
  k_dog_valid_actions<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, mask);
  return check_launch();
}

int dogstep_dog_obs_planes(const dogstep_dog_cfg* cfg) {
  DogGeom g;
  if (int rc = dog_make_geom(cfg, &g)) return rc;
  return 8 + 3 * g.n + kNCard;
}

int dogstep_dog_encode_board(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, int8_t* obs, void* stream) {
  DOG_PROLOGUE
  if (!obs) return DOGSTEP_ERR_INVALID_ARG;
  k_dog_encode_board<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, obs);
  return check_launch();
}

int dogstep_dog_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* action, int8_t* reward,
                     uint8_t* done, void* stream) {
  DOG_PROLOGUE
  if (!action) return DOGSTEP_ERR_INVALID_ARG;
  k_dog_step<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, action, reward, done);
  return check_launch();
}

int dogstep_dog_no_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, int8_t* reward, uint8_t* done,
                        void* stream) {
  DOG_PROLOGUE
  k_dog_no_step<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, reward, done);
  return check_launch();
}

int dogstep_dog_distribute_cards(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, void* stream) {
  DOG_PROLOGUE
  k_dog_distribute_cards<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n);
  return check_launch();
}

int dogstep_dog_substep(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* kind,
                        const int32_t* args, int8_t* board_out, int32_t* pins_out, int8_t* reward, uint8_t* done, void* stream) {
  DOG_PROLOGUE
  if (!kind || !args || !board_out || !pins_out || !reward || !done) return DOGSTEP_ERR_INVALID_ARG;
  k_dog_substep<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, kind, args, board_out, pins_out, reward, done);
  return check_launch();
}

int dogstep_dog_random_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const uint32_t* host_rng_key,
                            int64_t game_offset, unsigned long long* active_count, void* stream) {
  DOG_PROLOGUE
  if (!host_rng_key) return DOGSTEP_ERR_INVALID_ARG;
  k_dog_random_step<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, Key2{host_rng_key[0], host_rng_key[1]}, game_offset,
                                                           active_count);
  return check_launch();
}

int dogstep_dog_play_random(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const uint32_t* host_rng_key,
                            int64_t game_offset, int32_t max_steps, int32_t* game_len, unsigned long long* total_steps,
                            void* stream) {
  DOG_PROLOGUE
  if (!host_rng_key || max_steps < 0) return DOGSTEP_ERR_INVALID_ARG;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int64_t ctas_needed = (n + kSyncWarps - 1) / kSyncWarps;
  const unsigned grid = (unsigned)(ctas_needed < sms ? ctas_needed : sms);  // one persistent CTA per SM
  const size_t smem = sizeof(DogS) * kSyncWarps + 2 * 130 * sizeof(int);
  cudaFuncSetAttribute(k_dog_play_random, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // per-launch game queue: a rotating pool of counters so that launches in flight on different streams do not share one
  static unsigned int* pools[64] = {};  // one pool per device ordinal
  static std::atomic<unsigned> ticket{0};
  static std::mutex pool_mu;
  if (dev < 0 || dev >= 64) return DOGSTEP_ERR_UNSUPPORTED;
  {
    std::lock_guard<std::mutex> lock(pool_mu);
    if (!pools[dev] && cudaMalloc(&pools[dev], 64 * sizeof(unsigned int)) != cudaSuccess) return check_launch();
  }
  unsigned int* queue = pools[dev] + (ticket.fetch_add(1) & 63u);
  cudaMemsetAsync(queue, 0, sizeof(unsigned int), st);
#ifdef DOGSTEP_TRACE
  if (getenv("DOGSTEP_DOG_TRACE")) {
    static unsigned long long z[1024];
    cudaMemcpyToSymbol(g_dog_trace, z, sizeof(z));
  }
#endif
  k_dog_play_random<<<grid, kSyncWarps * 32, smem, st>>>(g, p, n, Key2{host_rng_key[0], host_rng_key[1]}, game_offset, max_steps,
                                                        game_len, total_steps, queue);
#ifdef DOGSTEP_TRACE
  if (getenv("DOGSTEP_DOG_TRACE")) {
    cudaStreamSynchronize(st);
    static unsigned long long h[1024];
    cudaMemcpyFromSymbol(h, g_dog_trace, sizeof(h));
    for (int r = 0; r < 511 && h[2 * r]; ++r)
      fprintf(stderr, "turn %5d  +%9.1f us  live warps %2llu%s", 64 * r, (h[2 * r] - h[0]) / 1e3, h[2 * r + 1], (r % 4 == 3) ? "\n" : "   ");
    fprintf(stderr, "\n");
    unsigned long long q[8], zz[8] = {0};
    cudaMemcpyFromSymbol(q, g_dog_solo, sizeof(q));
    cudaMemcpyToSymbol(g_dog_solo, zz, sizeof(zz));
    if (q[4])
      fprintf(stderr, "one live game (%llu turns): flags + barrier %.0f  mask tasks + barrier %.0f  draw %.0f  transition (or deal) %.0f cycles\n", q[4],
              (double)q[0] / q[4], (double)q[1] / q[4], (double)q[2] / q[4], (double)q[3] / q[4]);
  }
#endif
  return check_launch();
}

int dogstep_dog_agent_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* action,
                           const float* root_value, const float* action_weights, const int8_t* obs,
                           const dogstep_replay_arrays* traj, void* stream) {
  DOG_PROLOGUE
  if (!action || !root_value || !action_weights || !traj) return DOGSTEP_ERR_INVALID_ARG;
  if (traj->capacity < n || traj->action_dim != g.num_actions || traj->stochastic != 0 || traj->obs_size < 0 ||
      (traj->obs_size > 0 && !obs))
    return DOGSTEP_ERR_INVALID_ARG;
  if (!traj->observations || !traj->actions || !traj->rewards || !traj->root_values || !traj->child_visits || !traj->masks ||
      !traj->players || !traj->teams || !traj->discounts || !traj->episode_lengths)
    return DOGSTEP_ERR_INVALID_ARG;
  k_dog_agent_step<<<dog_blocks(n), kDogThreads, 0, st>>>(g, p, n, action, root_value, action_weights, obs, *traj);
  return check_launch();
}

}  // extern "C"
