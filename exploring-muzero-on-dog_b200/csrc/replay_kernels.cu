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
  const int64_t game = blockIdx.x;  // games on grid.x (2^31 - 1 of them), ply chunks on grid.y
  if (game >= n_games) return;
  const int pos = slot[game];
  if (pos < 0) return;
  const int Tb = buf.max_episode_length, Tt = traj.max_episode_length;
  // an episode can be no longer than the rows that exist on either side: idx > T' (a trajectory cut short) must not make the
  // stored length exceed the plies actually copied
  const int length = min(traj.episode_lengths[game], min(Tb, Tt));
  const int t0 = blockIdx.y * 8, t1 = min(t0 + 8, length);  // 8 plies per CTA
  const int tid = threadIdx.x;
  if (t1 > t0) {
    // the plies t0..t1-1 of one episode are contiguous in the trajectory and in the ring slot: one vectorised chunk per leaf
    const int np = t1 - t0;
    const int64_t src = game * Tt + t0, dst = (int64_t)pos * Tb + t0;
    const int64_t ne = (int64_t)np * buf.obs_size;
    if (buf.obs_is_int8 == traj.obs_is_int8) {
      const int es = buf.obs_is_int8 ? 1 : 4;
      coop_copy_bytes((char*)buf.observations + dst * buf.obs_size * es, (const char*)traj.observations + src * buf.obs_size * es, ne * es, tid, 256);
    } else if (buf.obs_is_int8) {
      const float* sf = (const float*)traj.observations + src * buf.obs_size;
      int8_t* d = (int8_t*)buf.observations + dst * buf.obs_size;
      for (int64_t k = tid; k < ne; k += 256) d[k] = (int8_t)sf[k];
    } else {
      coop_widen_i8_f32((float*)buf.observations + dst * buf.obs_size, (const int8_t*)traj.observations + src * buf.obs_size, ne, tid, 256);
    }
    coop_copy_bytes(buf.child_visits + dst * buf.action_dim, traj.child_visits + src * buf.action_dim, (int64_t)np * buf.action_dim * 4, tid, 256);
    if (buf.stochastic && tid < 6 * np) buf.dice_distributions[dst * 6 + tid] = traj.dice_distributions[src * 6 + tid];
    if (tid < np) {
      const int64_t a = src + tid, b = dst + tid;
      buf.actions[b] = traj.actions[a];
      buf.rewards[b] = traj.rewards[a];
      buf.root_values[b] = traj.root_values[a];
      buf.masks[b] = traj.masks[a];
      buf.players[b] = traj.players[a];
      buf.teams[b] = traj.teams[a];
      buf.discounts[b] = traj.discounts[a];
      if (buf.stochastic) buf.dice_outcomes[b] = traj.dice_outcomes[a];
    }
  }
  if (blockIdx.y == 0 && tid == 0) buf.episode_lengths[pos] = length;
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
  // a caller-supplied plan is clamped into the stored arrays (an empty slot reads as one all-zero ply)
  const int ep = min(max(ep_indices[b], 0), buf.capacity - 1);
  const int len = min(max(buf.episode_lengths[ep], 1), T);
  const int t0 = min(max(t_starts[b], 0), len - 1);
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

// ---- prioritised sampling (an extension beyond the reference, which samples uniformly with a terminal quota) --------
// Proportional prioritisation P(e, t) = p[e, t] / sum p over a two-level sum structure: per-(episode, ply) priorities and
// per-episode row sums.  Priorities are FIXED POINT (uint32, 2^-20 units) and sums uint64, so every sum is exact and
// independent of the order it is taken in: the parallel scans below and the NumPy oracle agree bit for bit.
constexpr int kPrioShift = 20;

// a stored priority is never 0 (floor: one fixed-point unit): a ply whose error reached 0 stays drawable, and the total can
// only be 0 for an empty buffer
__device__ __forceinline__ uint32_t prio_to_fixed(float p) {
  if (!(p > 0.0f)) return 1u;
  const double v = (double)p * (double)(1u << kPrioShift);
  return v >= 4294967295.0 ? 0xFFFFFFFFu : max((uint32_t)(v + 0.5), 1u);
}

// rows[i] < 0 is skipped; row = value for t < episode_lengths[row], 0 beyond; row sum recomputed (warp per row)
__global__ void __launch_bounds__(128) k_replay_prio_fill(uint32_t* __restrict__ prio, unsigned long long* __restrict__ row_sum,
                                                          const int32_t* __restrict__ episode_lengths, int T,
                                                          const int32_t* __restrict__ rows, int n_rows, float value) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= n_rows) return;
  const int row = rows[i];
  if (row < 0) return;
  const int len = min(episode_lengths[row], T);
  const uint32_t v = prio_to_fixed(value);
  for (int t = lane; t < T; t += 32) prio[(int64_t)row * T + t] = t < len ? v : 0u;
  if (lane == 0) row_sum[row] = (unsigned long long)v * (unsigned long long)max(len, 0);
}

// prio[ep[b], t[b]] = value[b]; the row sum follows through atomics, so duplicates in the batch stay consistent.  Pairs outside
// the stored episodes (ep outside [0, capacity), t outside [0, episode length)) are ignored.
__global__ void k_replay_prio_update(uint32_t* __restrict__ prio, unsigned long long* __restrict__ row_sum,
                                     const int32_t* __restrict__ episode_lengths, int capacity, int T, int B,
                                     const int32_t* __restrict__ ep, const int32_t* __restrict__ ts, const float* __restrict__ value) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int e = ep[b], t = ts[b];
  if (e < 0 || e >= capacity || t < 0 || t >= min(episode_lengths[e], T)) return;
  const uint32_t nv = prio_to_fixed(value[b]);
  const uint32_t old = atomicExch(&prio[(int64_t)e * T + t], nv);
  atomicAdd(&row_sum[e], (unsigned long long)nv - (unsigned long long)old);  // wraps modulo 2^64: exact
}

// inclusive scan of row_sum[0..size) -> cdf (one CTA of 1024 threads, chunks of 1024 with a running carry)
__global__ void __launch_bounds__(1024) k_replay_prio_scan(const unsigned long long* __restrict__ row_sum, int size,
                                                           unsigned long long* __restrict__ cdf) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0ull;
  __syncthreads();
  for (int base = 0; base < size; base += 1024) {
    const int i = base + tid;
    unsigned long long v = i < size ? row_sum[i] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, w, o);
        if (lane >= o) w += u;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const unsigned long long before = (warp ? warp_tot[warp - 1] : 0ull) + carry_s;
    if (i < size) cdf[i] = v + before;
    __syncthreads();
    if (tid == 1023) carry_s = v + before;
    __syncthreads();
  }
}

// sample b: target = floor(bits64 * total / 2^64) -> episode by binary search in cdf, ply by a row scan
__global__ void k_replay_plan_prio(const uint32_t* __restrict__ prio, const unsigned long long* __restrict__ cdf, int size, int T, int B,
                                   Key2 key, int32_t* __restrict__ ep_indices, int32_t* __restrict__ t_starts,
                                   double* __restrict__ prob) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const unsigned long long total = cdf[size - 1];
  const unsigned long long bits = ((unsigned long long)bits_i(key, 2u * b) << 32) | (unsigned long long)bits_i(key, 2u * b + 1u);
  const unsigned long long target = __umul64hi(bits, total);  // uniform in [0, total)
  int lo = 0, hi = size - 1;                                   // first episode with cdf > target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] > target) hi = mid;
    else lo = mid + 1;
  }
  const int ep = lo;
  unsigned long long r = target - (ep ? cdf[ep - 1] : 0ull), acc = 0ull;
  int t = 0;
  uint32_t pv = 0u;
  for (int k = 0; k < T; ++k) {
    pv = prio[(int64_t)ep * T + k];
    acc += pv;
    t = k;
    if (acc > r) break;
  }
  ep_indices[b] = ep;
  t_starts[b] = t;
  prob[b] = total ? (double)pv / (double)total : 0.0;
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
  if (!slot || n_games < 0 || n_games > 0x7FFFFFFF || (traj->max_episode_length + 7) / 8 > 65535) return DOGSTEP_ERR_INVALID_ARG;
  if (buf->obs_size != traj->obs_size || buf->action_dim != traj->action_dim || buf->stochastic != traj->stochastic)
    return DOGSTEP_ERR_INVALID_ARG;
  if (n_games == 0) return DOGSTEP_OK;
  dim3 grid((unsigned)n_games, (unsigned)((traj->max_episode_length + 7) / 8));
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

int dogstep_replay_prio_fill(uint32_t* prio, unsigned long long* row_sum, const int32_t* episode_lengths, int32_t max_episode_length,
                             const int32_t* rows, int32_t n_rows, float value, void* stream) {
  if (!prio || !row_sum || !episode_lengths || !rows || max_episode_length < 1 || n_rows < 0) return DOGSTEP_ERR_INVALID_ARG;
  if (n_rows == 0) return DOGSTEP_OK;
  k_replay_prio_fill<<<(n_rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(prio, row_sum, episode_lengths, max_episode_length, rows, n_rows, value);
  return check_launch();
}

int dogstep_replay_prio_update(uint32_t* prio, unsigned long long* row_sum, const int32_t* episode_lengths, int32_t capacity,
                               int32_t max_episode_length, int32_t batch_size, const int32_t* ep_indices, const int32_t* t_starts,
                               const float* value, void* stream) {
  if (!prio || !row_sum || !episode_lengths || !ep_indices || !t_starts || !value || capacity < 1 || max_episode_length < 1 ||
      batch_size < 0)
    return DOGSTEP_ERR_INVALID_ARG;
  if (batch_size == 0) return DOGSTEP_OK;
  k_replay_prio_update<<<(batch_size + 127) / 128, 128, 0, (cudaStream_t)stream>>>(prio, row_sum, episode_lengths, capacity,
                                                                                   max_episode_length, batch_size, ep_indices,
                                                                                   t_starts, value);
  return check_launch();
}

int dogstep_replay_plan_prioritized(const uint32_t* prio, const unsigned long long* row_sum, unsigned long long* cdf_work, int32_t size,
                                    int32_t max_episode_length, int32_t batch_size, const uint32_t* host_key, int32_t* ep_indices,
                                    int32_t* t_starts, double* prob, void* stream) {
  if (!prio || !row_sum || !cdf_work || !host_key || !ep_indices || !t_starts || !prob || size < 1 || max_episode_length < 1 ||
      batch_size < 1)
    return DOGSTEP_ERR_INVALID_ARG;
  k_replay_prio_scan<<<1, 1024, 0, (cudaStream_t)stream>>>(row_sum, size, cdf_work);
  k_replay_plan_prio<<<(batch_size + 127) / 128, 128, 0, (cudaStream_t)stream>>>(prio, cdf_work, size, max_episode_length, batch_size,
                                                                                 Key2{host_key[0], host_key[1]}, ep_indices, t_starts, prob);
  return check_launch();
}

}  // extern "C"
