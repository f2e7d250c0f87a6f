// replay_kernels.cu — device-resident replay ring buffer: episode save, sampling plan, K-step window gather with
// value targets.  Restates MuZero_det_MADN/vec_replay_buffer.py:36-264 (+ the stochastic deltas of
// MuZero_Classic_MADN/vec_replay_buffer_stochastic.py).  All three kernels are pure HBM copies / gathers:
//   save    one CTA per (game, chunk of plies): coalesced row copies trajectory -> ring slot
//   gather  one CTA per sample: root observation copy (obs_size floats) + K small rows; the target arithmetic is a few
//           float64 multiplies per (sample, k)
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/dogstep.h"
#include "common.cuh"
#include "jaxrand.cuh"

namespace dogstep {

__global__ void __launch_bounds__(256) k_replay_save(dogstep_replay_arrays buf, dogstep_replay_arrays traj, int64_t n_games,
                                                     const int32_t* __restrict__ slot) {
  const int64_t game = blockIdx.y;
  if (game >= n_games) return;
  const int pos = slot[game];
  if (pos < 0) return;
  const int length = min(traj.episode_lengths[game], buf.max_episode_length);
  const int Tb = buf.max_episode_length, Tt = traj.max_episode_length;
  const int t0 = blockIdx.x * 8, t1 = min(t0 + 8, length);  // 8 plies per CTA
  const int tid = threadIdx.x;
  for (int t = t0; t < t1; ++t) {
    const int64_t src = game * Tt + t, dst = (int64_t)pos * Tb + t;
    if (buf.obs_is_int8) {
      const int8_t* s = (const int8_t*)traj.observations + src * buf.obs_size;
      int8_t* d = (int8_t*)buf.observations + dst * buf.obs_size;
      if (traj.obs_is_int8) for (int k = tid; k < buf.obs_size; k += 256) d[k] = s[k];
      else { const float* sf = (const float*)traj.observations + src * buf.obs_size; for (int k = tid; k < buf.obs_size; k += 256) d[k] = (int8_t)sf[k]; }
    } else {
      float* d = (float*)buf.observations + dst * buf.obs_size;
      if (traj.obs_is_int8) { const int8_t* s = (const int8_t*)traj.observations + src * buf.obs_size; for (int k = tid; k < buf.obs_size; k += 256) d[k] = (float)s[k]; }
      else { const float* s = (const float*)traj.observations + src * buf.obs_size; for (int k = tid; k < buf.obs_size; k += 256) d[k] = s[k]; }
    }
    for (int k = tid; k < buf.action_dim; k += 256) buf.child_visits[dst * buf.action_dim + k] = traj.child_visits[src * buf.action_dim + k];
    if (buf.stochastic && tid < 6) buf.dice_distributions[dst * 6 + tid] = traj.dice_distributions[src * 6 + tid];
    if (tid == 0) {
      buf.actions[dst] = traj.actions[src];
      buf.rewards[dst] = traj.rewards[src];
      buf.root_values[dst] = traj.root_values[src];
      buf.masks[dst] = traj.masks[src];
      buf.players[dst] = traj.players[src];
      buf.teams[dst] = traj.teams[src];
      buf.discounts[dst] = traj.discounts[src];
      if (buf.stochastic) buf.dice_outcomes[dst] = traj.dice_outcomes[src];
    }
  }
  if (blockIdx.x == 0 && tid == 0) buf.episode_lengths[pos] = length;
}

__global__ void k_replay_plan(const int32_t* __restrict__ episode_lengths, int size, int B, int unroll_steps, int n_normal,
                              Key2 key, int32_t* __restrict__ ep_indices, int32_t* __restrict__ t_starts) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t r0 = bits_i(key, 2u * b), r1 = bits_i(key, 2u * b + 1u);
  const int ep = (int)(r0 % (uint32_t)max(size, 1));
  const int len = max(episode_lengths[ep], 1);
  int t;
  if (b < n_normal) {
    t = (int)(r1 % (uint32_t)len);  // t_start uniform in [0, len-1]
  } else {
    const int max_k = min(unroll_steps - 1, len - 1);
    const int k = (int)(r1 % (uint32_t)(max(max_k, 0) + 1));
    t = max(len - 1 - k, 0);
  }
  ep_indices[b] = ep;
  t_starts[b] = t;
}

__global__ void __launch_bounds__(128) k_replay_gather(dogstep_replay_arrays buf, int B, int unroll_steps, int TD, int bootstrap_flag,
                                                       const double* __restrict__ gamma_pow, const int32_t* __restrict__ ep_indices,
                                                       const int32_t* __restrict__ t_starts, dogstep_replay_batch out) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const int K = unroll_steps + 1, A = buf.action_dim, T = buf.max_episode_length, tid = threadIdx.x;
  const int ep = ep_indices[b], t0 = t_starts[b];
  const int len = buf.episode_lengths[ep];
  const int64_t base = (int64_t)ep * T;
  // root observation (:104)
  {
    float* d = out.observations + (int64_t)b * buf.obs_size;
    const int64_t o = (base + t0) * buf.obs_size;
    if (buf.obs_is_int8) { const int8_t* s = (const int8_t*)buf.observations + o; for (int k = tid; k < buf.obs_size; k += 128) d[k] = (float)s[k]; }
    else { const float* s = (const float*)buf.observations + o; for (int k = tid; k < buf.obs_size; k += 128) d[k] = s[k]; }
  }
  // policies (:146, :246)
  for (int e = tid; e < K * A; e += 128) {
    const int k = e / A, a = e - k * A;
    const int idx = t0 + k, valid = idx < len, ci = min(idx, len - 1);
    out.policies[((int64_t)b * K + k) * A + a] = valid ? buf.child_visits[(base + ci) * A + a] : 0.0f;
  }
  if (buf.stochastic)
    for (int e = tid; e < (K - 1) * 6; e += 128) {
      const int k = e / 6, j = e - k * 6;
      const int idx = t0 + k, valid = idx < len, ci = min(idx, len - 1);
      out.dice_probs[((int64_t)b * (K - 1) + k) * 6 + j] = valid ? buf.dice_distributions[(base + ci) * 6 + j] : (float)(1.0 / 6.0);
    }
  if (tid < K) {
    const int k = tid;
    const int idx = t0 + k, valid = idx < len, ci = min(idx, len - 1);
    const int fin = len - 1;
    const int final_reward = buf.rewards[base + fin], final_player = buf.players[base + fin], final_team = buf.teams[base + fin];
    const int player = buf.players[base + ci], team = buf.teams[base + ci];
    // z from the perspective of THIS timestep (:170-185)
    const int won = buf.stochastic ? (final_reward > 0) : (final_reward == 2);
    double z = 0.0;
    if (won) z = (team == -1) ? ((final_player == player) ? 1.0 : -1.0) : ((final_team == team) ? 1.0 : -1.0);
    const int steps_until_end = len - 1 - idx;
    const int bootstrap_from_value = steps_until_end >= TD;
    const int bi = min(idx + TD, len - 1);
    float bv = buf.root_values[base + bi];
    const int same = (team != -1) ? (team == buf.teams[base + bi]) : (player == buf.players[base + bi]);
    if (!same) bv = -bv;
    z = z * gamma_pow[max(steps_until_end, 0)];
    double target;
    if (z == 0.0 || (bootstrap_from_value && bootstrap_flag)) target = (double)bv * gamma_pow[max(min(TD, steps_until_end), 0)];
    else target = z;
    target = fmin(fmax(target, -1.0), 1.0);
    out.values[b * K + k] = valid ? buf.root_values[base + ci] : 0.0f;
    out.masks[b * K + k] = valid ? buf.masks[base + ci] : 0.0f;
    out.target_values[b * K + k] = valid ? (float)target : 0.0f;
    if (k < K - 1) {
      out.actions[b * (K - 1) + k] = valid ? buf.actions[base + ci] : 0;
      out.rewards[b * (K - 1) + k] = valid ? buf.rewards[base + ci] : 1;
      out.discount_targets[b * (K - 1) + k] = valid ? buf.discounts[base + ci] : 1;
      if (buf.stochastic) out.dice_outcomes[b * (K - 1) + k] = max((valid ? buf.dice_outcomes[base + ci] : 0) - 1, 0);
    }
  }
}

static int replay_check(const dogstep_replay_arrays* a) {
  if (!a || a->capacity < 1 || a->max_episode_length < 1 || a->obs_size < 1 || a->action_dim < 1) return DOGSTEP_ERR_INVALID_ARG;
  if (!a->observations || !a->actions || !a->rewards || !a->root_values || !a->child_visits || !a->masks || !a->players ||
      !a->teams || !a->discounts || !a->episode_lengths)
    return DOGSTEP_ERR_INVALID_ARG;
  if (a->stochastic && (!a->dice_outcomes || !a->dice_distributions)) return DOGSTEP_ERR_INVALID_ARG;
  return DOGSTEP_OK;
}

}  // namespace dogstep

using namespace dogstep;

extern "C" {

int dogstep_replay_save(const dogstep_replay_arrays* buf, const dogstep_replay_arrays* traj, int64_t n_games, const int32_t* slot,
                        void* stream) {
  if (int rc = replay_check(buf)) return rc;
  if (int rc = replay_check(traj)) return rc;
  if (!slot || n_games < 0 || n_games > 65535) return DOGSTEP_ERR_INVALID_ARG;
  if (buf->obs_size != traj->obs_size || buf->action_dim != traj->action_dim || buf->stochastic != traj->stochastic)
    return DOGSTEP_ERR_INVALID_ARG;
  if (n_games == 0) return DOGSTEP_OK;
  dim3 grid((unsigned)((traj->max_episode_length + 7) / 8), (unsigned)n_games);
  k_replay_save<<<grid, 256, 0, (cudaStream_t)stream>>>(*buf, *traj, n_games, slot);
  return check_launch();
}

int dogstep_replay_plan(const dogstep_replay_arrays* buf, int32_t size, int32_t batch_size, int32_t unroll_steps, float terminal_ratio,
                        const uint32_t* host_key, int32_t* ep_indices, int32_t* t_starts, void* stream) {
  if (int rc = replay_check(buf)) return rc;
  if (!host_key || !ep_indices || !t_starts || size < 1 || size > buf->capacity || batch_size < 1 || unroll_steps < 1)
    return DOGSTEP_ERR_INVALID_ARG;
  const int n_terminal = (int)((double)batch_size * (double)terminal_ratio);  // int(batch_size * TERMINAL_RATIO) (:73)
  k_replay_plan<<<(batch_size + 127) / 128, 128, 0, (cudaStream_t)stream>>>(buf->episode_lengths, size, batch_size, unroll_steps,
                                                                           batch_size - n_terminal, Key2{host_key[0], host_key[1]},
                                                                           ep_indices, t_starts);
  return check_launch();
}

int dogstep_replay_gather(const dogstep_replay_arrays* buf, int32_t batch_size, int32_t unroll_steps, int32_t td_steps,
                          int32_t bootstrap_value_target, const double* gamma_pow, const int32_t* ep_indices, const int32_t* t_starts,
                          const dogstep_replay_batch* out, void* stream) {
  if (int rc = replay_check(buf)) return rc;
  if (!gamma_pow || !ep_indices || !t_starts || !out || batch_size < 1 || unroll_steps < 1 || unroll_steps + 1 > 128 || td_steps < 0)
    return DOGSTEP_ERR_INVALID_ARG;
  if (!out->observations || !out->actions || !out->rewards || !out->policies || !out->values || !out->masks ||
      !out->target_values || !out->discount_targets)
    return DOGSTEP_ERR_INVALID_ARG;
  if (buf->stochastic && (!out->dice_outcomes || !out->dice_probs)) return DOGSTEP_ERR_INVALID_ARG;
  k_replay_gather<<<batch_size, 128, 0, (cudaStream_t)stream>>>(*buf, batch_size, unroll_steps, td_steps, bootstrap_value_target,
                                                                gamma_pow, ep_indices, t_starts, *out);
  return check_launch();
}

}  // extern "C"
