// ttt_kernels.cu — TicTacToe / TicTacToeV2 envs and their true-env mctx callbacks (root_fn, recurrent_fn with a random
// rollout to the end of the game) for BASELINE config 1.  One game per thread; a game is 18 bytes.
// Restates TicTacToe/TicTacToe.py:19-117 (variant 0) and TicTacToe/TicTacToeV2.py:22-140 (variant 1, including the two
// operator-precedence quirks of env_step :66 and :70).
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/dogstep.h"
#include "common.cuh"
#include "jaxrand.cuh"
#include "ttt_core.cuh"

namespace dogstep {

struct TttPtrs { int8_t* board; int8_t* cur; int8_t* reward; uint8_t* done; int8_t* memory; };

__device__ __forceinline__ void ttt_load(const TttPtrs& p, int64_t g, Ttt& e) {
  for (int k = 0; k < 9; ++k) e.board[k] = p.board[9 * g + k];
  e.cur = p.cur[g]; e.reward = p.reward[g]; e.done = p.done[g] != 0;
  for (int k = 0; k < 6; ++k) e.memory[k] = p.memory[6 * g + k];
}
__device__ __forceinline__ void ttt_store(const TttPtrs& p, int64_t g, const Ttt& e) {
  for (int k = 0; k < 9; ++k) p.board[9 * g + k] = e.board[k];
  p.cur[g] = e.cur; p.reward[g] = e.reward; p.done[g] = (uint8_t)e.done;
  for (int k = 0; k < 6; ++k) p.memory[6 * g + k] = e.memory[k];
}
__global__ void k_ttt_reset(TttPtrs p, int64_t n) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  Ttt e;
  for (int k = 0; k < 9; ++k) e.board[k] = 0;
  e.cur = 1; e.reward = 0; e.done = 0;
  for (int k = 0; k < 6; ++k) e.memory[k] = -1;
  ttt_store(p, g, e);
}
__global__ void k_ttt_step(TttPtrs p, int64_t n, int variant, const int8_t* __restrict__ action, int8_t* __restrict__ reward,
                           uint8_t* __restrict__ done) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  Ttt e;
  ttt_load(p, g, e);
  ttt_step(variant, e, action[g]);
  ttt_store(p, g, e);
  if (reward) reward[g] = e.reward;
  if (done) done[g] = (uint8_t)e.done;
}
__global__ void k_ttt_policy(TttPtrs p, int64_t n, int variant, float* __restrict__ logits, uint8_t* __restrict__ valid) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  Ttt e;
  ttt_load(p, g, e);
  if (logits) { float lg[9]; ttt_policy(variant, e, lg); for (int a = 0; a < 9; ++a) logits[9 * g + a] = lg[a]; }
  if (valid) for (int a = 0; a < 9; ++a) valid[9 * g + a] = (uint8_t)(!e.done && e.board[a] == 0);  // valid_action_mask
}
// 16 lanes per game (see ttt_rollout_group)
__global__ void __launch_bounds__(128) k_ttt_root_fn(TttPtrs p, int64_t n, int variant, const uint32_t* __restrict__ keys,
                                                     float* __restrict__ prior, float* __restrict__ value, float* __restrict__ emb) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g = t >> 4;
  const int sub = (int)(t & 15);
  if (g >= n) return;
  const uint32_t gmask = 0xFFFFu << (16 * ((threadIdx.x >> 4) & 1));
  Ttt e;
  ttt_load(p, g, e);
  if (sub < 9) prior[9 * g + sub] = ttt_policy_a(variant, e, sub);
  const float v = ttt_rollout_group(variant, e, Key2{keys[2 * g], keys[2 * g + 1]}, sub, gmask);
  if (sub == 0) {
    value[g] = v;
    ttt_to_emb(e, emb + 18 * g);
  }
}
__global__ void __launch_bounds__(128) k_ttt_recurrent_fn(int64_t n, int variant, const uint32_t* __restrict__ keys,
                                                          const int32_t* __restrict__ action, const float* __restrict__ emb_in,
                                                          float* __restrict__ prior, float* __restrict__ value, float* __restrict__ reward,
                                                          float* __restrict__ discount, float* __restrict__ emb_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g = t >> 4;
  const int sub = (int)(t & 15);
  if (g >= n) return;
  const uint32_t gmask = 0xFFFFu << (16 * ((threadIdx.x >> 4) & 1));
  Ttt e;
  ttt_from_emb(e, emb_in + 18 * g);
  ttt_step(variant, e, (int)(int8_t)action[g]);
  if (sub < 9) prior[9 * g + sub] = ttt_policy_a(variant, e, sub);
  const float v = e.done ? 0.0f : ttt_rollout_group(variant, e, Key2{keys[2 * g], keys[2 * g + 1]}, sub, gmask);
  __syncwarp(gmask);  // emb_out may alias emb_in: every lane of the group has read the embedding before lane 0 writes
  if (sub == 0) {
    reward[g] = (float)e.reward;
    discount[g] = e.done ? 0.0f : -1.0f;
    value[g] = v;
    ttt_to_emb(e, emb_out + 18 * g);
  }
}

static int ttt_ptrs(const dogstep_ttt_state* s, TttPtrs* p) {
  if (!s || !s->board || !s->current_player || !s->reward || !s->done || !s->memory) return DOGSTEP_ERR_INVALID_ARG;
  *p = TttPtrs{s->board, s->current_player, s->reward, s->done, s->memory};
  return DOGSTEP_OK;
}
static inline unsigned tb(int64_t n) { return (unsigned)((n + 127) / 128); }
// One ply of the reference's match loops (TicTacToe/eval.py:97-125 play_match, :151-176 play_mcts_match): a game that is not
// done plays get_mcts_action (:28-34) — the first largest action weight among the empty cells, argmax(where(board == 0,
// action_weights, -inf)) — through env_step; a finished game is left alone (the reference's `while not env.done`).
__global__ void k_ttt_play_move(TttPtrs p, int64_t n, int variant, const float* __restrict__ weights, int8_t* __restrict__ action,
                                int32_t* __restrict__ plies) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  Ttt e;
  ttt_load(p, g, e);
  if (e.done) {
    if (action) action[g] = -1;
    return;
  }
  int a = 0;
  float best = -INFINITY;
  for (int c = 0; c < 9; ++c) {
    const float x = e.board[c] == 0 ? weights[9 * g + c] : -INFINITY;
    if (x > best) { best = x; a = c; }
  }
  ttt_step(variant, e, (int8_t)a);
  ttt_store(p, g, e);
  if (action) action[g] = (int8_t)a;
  if (plies) plies[g] += 1;
}
static inline unsigned tb16(int64_t n) { return (unsigned)((16 * n + 127) / 128); }  // 16 lanes per game

}  // namespace dogstep

using namespace dogstep;

extern "C" {

int dogstep_ttt_reset(const dogstep_ttt_state* s, int64_t n, void* stream) {
  TttPtrs p;
  if (n < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = ttt_ptrs(s, &p)) return rc;
  if (n == 0) return DOGSTEP_OK;
  k_ttt_reset<<<tb(n), 128, 0, (cudaStream_t)stream>>>(p, n);
  return check_launch();
}
int dogstep_ttt_step(const dogstep_ttt_state* s, int64_t n, int32_t variant, const int8_t* action, int8_t* reward, uint8_t* done,
                     void* stream) {
  TttPtrs p;
  if (n < 0 || !action || variant < 0 || variant > 1) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = ttt_ptrs(s, &p)) return rc;
  if (n == 0) return DOGSTEP_OK;
  k_ttt_step<<<tb(n), 128, 0, (cudaStream_t)stream>>>(p, n, variant, action, reward, done);
  return check_launch();
}
int dogstep_ttt_play_move(const dogstep_ttt_state* s, int64_t n, int32_t variant, const float* action_weights, int8_t* action,
                          int32_t* plies, void* stream) {
  TttPtrs p;
  if (n < 0 || !action_weights || variant < 0 || variant > 1) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = ttt_ptrs(s, &p)) return rc;
  if (n == 0) return DOGSTEP_OK;
  k_ttt_play_move<<<tb(n), 128, 0, (cudaStream_t)stream>>>(p, n, variant, action_weights, action, plies);
  return check_launch();
}
int dogstep_ttt_policy_function(const dogstep_ttt_state* s, int64_t n, int32_t variant, float* logits, uint8_t* valid_mask,
                                void* stream) {
  TttPtrs p;
  if (n < 0 || variant < 0 || variant > 1 || (!logits && !valid_mask)) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = ttt_ptrs(s, &p)) return rc;
  if (n == 0) return DOGSTEP_OK;
  k_ttt_policy<<<tb(n), 128, 0, (cudaStream_t)stream>>>(p, n, variant, logits, valid_mask);
  return check_launch();
}
int dogstep_ttt_root_fn(const dogstep_ttt_state* s, int64_t n, int32_t variant, const uint32_t* keys, float* prior_logits,
                        float* value, float* embedding, void* stream) {
  TttPtrs p;
  if (n < 0 || variant < 0 || variant > 1 || !keys || !prior_logits || !value || !embedding) return DOGSTEP_ERR_INVALID_ARG;
  if (int rc = ttt_ptrs(s, &p)) return rc;
  if (n == 0) return DOGSTEP_OK;
  k_ttt_root_fn<<<tb16(n), 128, 0, (cudaStream_t)stream>>>(p, n, variant, keys, prior_logits, value, embedding);
  return check_launch();
}
int dogstep_ttt_recurrent_fn(int64_t n, int32_t variant, const uint32_t* keys, const int32_t* action, const float* embedding_in,
                             float* prior_logits, float* value, float* reward, float* discount, float* embedding_out,
                             void* stream) {
  if (n < 0 || variant < 0 || variant > 1 || !keys || !action || !embedding_in || !prior_logits || !value || !reward ||
      !discount || !embedding_out)
    return DOGSTEP_ERR_INVALID_ARG;
  if (n == 0) return DOGSTEP_OK;
  k_ttt_recurrent_fn<<<tb16(n), 128, 0, (cudaStream_t)stream>>>(n, variant, keys, action, embedding_in, prior_logits, value, reward,
                                                              discount, embedding_out);
  return check_launch();
}

}  // extern "C"
