// madn_core.cuh — register-resident rules of the two MADN environments.
//
// One game lives in one thread's registers:
//   occ[p]  uint64 bitboard of the cells whose board value is p   (board  int8[total<=64])
//   pins[p] four int8 pin positions packed in one uint32          (pins   int8[P,4])
//   as[p]   six int8 card counts packed in the low 48 bits        (action_set int8[P,6], det only)
// so `board[c] != cp` is one bit test and a capture is a byte-wise compare on a packed word.
// Reference semantics restated here (bit-exact, including the quirks of SURVEY Appendix A):
//   valid_action  MADN/deterministic_madn.py:299-393, MADN/classic_madn.py:367-461
//   env_step      MADN/deterministic_madn.py:170-257, MADN/classic_madn.py:257-337
//   no_step       MADN/deterministic_madn.py:283-297, MADN/classic_madn.py:353-365
//   get_winner    MADN/deterministic_madn.py:122-168
//   goal-lane path check  utils/utility_funcs.py:142-184
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/dogstep.h"
#include "hostdev.cuh"

namespace dogstep {

struct MadnGeom {
  int n, d, bs, total;
  int start[4], target[4], goal0[4];
  uint32_t rules;
};

// geometry of env_reset (deterministic_madn.py:62-78); returns DOGSTEP_* code
inline int madn_make_geom(const dogstep_madn_cfg* cfg, MadnGeom* g) {
  if (!cfg) return DOGSTEP_ERR_INVALID_ARG;
  if (cfg->num_players < 2 || cfg->num_players > 4) return DOGSTEP_ERR_INVALID_ARG;
  if (cfg->distance < 1) return DOGSTEP_ERR_INVALID_ARG;
  if (cfg->distance > 12) return DOGSTEP_ERR_UNSUPPORTED;  // bitboard holds 64 cells
  g->n = cfg->num_players;
  g->d = cfg->distance;
  g->bs = 4 * cfg->distance;
  g->total = g->bs + 16;
  uint32_t rules = cfg->rules;
  if (g->n != 4) rules &= ~DOGSTEP_RULE_TEAMS;
  g->rules = rules;
  int layout = cfg->layout_mask & 0xF, cnt = 0;
  for (int i = 0; i < 4; ++i) cnt += (layout >> i) & 1;
  if (cnt != g->n || (layout == 0xF && g->n < 4)) layout = (1 << g->n) - 1;
  int r = 0;
  for (int i = 0; i < 4; ++i) {
    if (!((layout >> i) & 1)) continue;
    g->start[r] = i * g->d;
    g->target[r] = (g->start[r] - 1 + g->bs) % g->bs;
    g->goal0[r] = g->bs + 4 * i;
    ++r;
  }
  for (; r < 4; ++r) g->start[r] = g->target[r] = g->goal0[r] = 0;
  return DOGSTEP_OK;
}

struct MadnRegs {
  uint64_t occ[4];
  uint32_t pins[4];
  uint64_t as[4];
  int cur, done, reward, die;
};

#define DS_RULE(g, bit) (((g).rules & (bit)) != 0u)

template <typename T>
DS_FN T pick4(const T (&a)[4], int i) {
  T r = a[0];
  r = (i == 1) ? a[1] : r;
  r = (i == 2) ? a[2] : r;
  r = (i == 3) ? a[3] : r;
  return r;
}
DS_FN int byte_s(uint32_t w, int i) { return (int)(int8_t)(w >> (8 * i)); }
DS_FN int byte_s64(uint64_t w, int i) { return (int)(int8_t)(w >> (8 * i)); }
DS_FN int bit64(uint64_t w, int i) { return (int)((w >> i) & 1ull); }
DS_FN int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
DS_FN int floormod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }
// JAX gather index: negative wraps once, then clamp
DS_FN int gidx(int i, int size) {
  i = (i < 0) ? i + size : i;
  return min(max(i, 0), size - 1);
}

// cells of player p's four pins as a bitboard (-1 / out-of-range pins dropped)
DS_FN uint64_t pins_to_bits(uint32_t w, int total) {
  uint64_t b = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int pos = byte_s(w, i);
    b |= (pos >= 0 && pos < total) ? (1ull << pos) : 0ull;
  }
  return b;
}

// set_pins_on_board (deterministic_madn.py:259-271): later (higher player / pin) writes win
DS_FN void rebuild_occ(const MadnGeom& g, MadnRegs& s) {
  uint64_t later = 0;
#pragma unroll
  for (int p = 3; p >= 0; --p) {
    uint64_t b = (p < g.n) ? pins_to_bits(s.pins[p], g.total) : 0ull;
    s.occ[p] = b & ~later;
    later |= b;
  }
}

// is_player_done (deterministic_madn.py:122-137): the four goal cells are occupied by ANYONE
DS_FN int player_done(const MadnGeom& g, uint64_t anyocc, int p) {
  return (p >= 0 && p < g.n) ? (int)(((anyocc >> g.goal0[p]) & 0xFull) == 0xFull) : 0;
}

// team proxy (deterministic_madn.py:184,310)
DS_FN int mover_of(const MadnGeom& g, const MadnRegs& s, int pid) {
  uint64_t any = s.occ[0] | s.occ[1] | s.occ[2] | s.occ[3];
  int cp = (DS_RULE(g, DOGSTEP_RULE_TEAMS) && player_done(g, any, pid)) ? ((pid + 2) & 3) : pid;
  return gidx(cp, g.n);
}

// goal-lane path check: no own pin on lane cell k with s < k < e (lane = own bits of the 4 goal cells)
DS_FN int lane_clear(uint32_t lane, int s, int e) {
  int lo = max(s + 1, 0), hi = min(e - 1, 3);
  uint32_t m = (lo > hi) ? 0u : (((2u << hi) - 1u) & ~((1u << lo) - 1u));
  return (lane & m) == 0u;
}

// 4-bit mask, bit q = pins_on_start[q] = (board[start[q]] == q)
DS_FN uint32_t pins_on_start_mask(const MadnGeom& g, const MadnRegs& s) {
  uint32_t m = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) m |= (q < g.n) ? ((uint32_t)bit64(s.occ[q], g.start[q]) << q) : 0u;
  return m;
}

// One (pin position, move) cell of valid_action before home-pin / action-set handling.
DS_FN int move_ok(const MadnGeom& g, uint64_t own, uint32_t posmask, int cp, int pos, int m) {
  const int mts = DS_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START);
  const int circ = DS_RULE(g, DOGSTEP_RULE_CIRCULAR_BOARD);
  const int jump = DS_RULE(g, DOGSTEP_RULE_JUMP_IN_GOAL);
  const int target = g.target[cp], goal0 = g.goal0[cp];
  int moved = pos + m;
  int fitted = floormod(moved, g.bs);
  int x = moved - target - mts;
  int result = !bit64(own, fitted) || DS_RULE(g, DOGSTEP_RULE_FRIENDLY_FIRE);
  if (DS_RULE(g, DOGSTEP_RULE_START_BLOCKING)) {
    int nsb = floormod(floordiv(pos, g.d) + 1, g.n);
    int nsa = min(fitted / g.d, g.n - 1);
    int trav = g.start[nsb] == g.start[nsa];
    int blocked = (posmask >> nsa) & 1;
    if (trav) result = result && (!blocked || pos == g.start[cp]);
    if (mts && trav && blocked) x = 0;
  }
  if (!circ && pos <= target && (x > 4 || (x == 0 && mts))) result = 0;
  uint32_t lane = (uint32_t)(own >> goal0) & 0xFu;
  if (x >= 1 && x <= 4 && pos <= target) {
    int B = !((lane >> (x - 1)) & 1u);
    int C = jump || lane_clear(lane, -1, x);
    result = (circ && result) || (B && C);
  }
  if (pos >= goal0 && pos <= goal0 + 3) {
    int k1 = moved - goal0;
    int ok = moved <= goal0 + 3 && k1 >= 0;
    int landing_free = ok ? !((lane >> k1) & 1u) : 0;
    int D = jump || lane_clear(lane, pos - goal0, k1 + 1);
    result = ok && landing_free && D;
  }
  return result;
}

// Bit-parallel form of move_ok for the six moves m = 1..6 of one pin (bit m-1), valid when
// start blocking is off (then x is never overridden and the start lookups drop out).  Same
// reference lines as move_ok; cross-checked against it over random rule sets in tests/.
DS_FN uint32_t move_row_fast(const MadnGeom& g, uint64_t own, int cp, int pos) {
  const int mts = DS_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START);
  const int circ = DS_RULE(g, DOGSTEP_RULE_CIRCULAR_BOARD);
  const int jump = DS_RULE(g, DOGSTEP_RULE_JUMP_IN_GOAL);
  const int target = g.target[cp], goal0 = g.goal0[cp];
  const uint32_t lane = (uint32_t)(own >> goal0) & 0xFu;
  if (pos >= goal0 && pos <= goal0 + 3) {  // pin in its goal lane (:376-381)
    const int k0 = pos - goal0;
    const uint32_t above = lane >> (k0 + 1);                        // bit m-1 = own pin on lane cell k0+m
    const uint32_t clear = above ? ((above & (0u - above)) - 1u) : 0xFu;  // cells below the first own pin
    return (jump ? ~above : clear) & ((1u << (3 - k0)) - 1u);
  }
  const uint64_t ring = own & ((1ull << g.bs) - 1ull);
  const uint32_t landing_own = (uint32_t)(((ring | (ring << g.bs)) >> (pos + 1)) & 0x3Full);
  uint32_t row = DS_RULE(g, DOGSTEP_RULE_FRIENDLY_FIRE) ? 0x3Fu : (~landing_own & 0x3Fu);
  if (pos <= target) {
    const int t = target + mts - pos;  // x = m - t, t >= 0 here
    if (!circ) {
      // (:349-357) x > 4  <=>  m > t + 4 ;  x == 0 && mts  <=>  m == t
      const uint32_t beyond = (t + 4 < 6) ? (0x3Fu & ~((1u << (t + 4)) - 1u)) : 0u;
      const uint32_t at_target = (mts && t >= 1 && t <= 6) ? (1u << (t - 1)) : 0u;
      row &= ~(beyond | at_target);
    }
    if (t < 6) {  // goal entry window 1 <= x <= 4  <=>  t+1 <= m <= t+4  (:360-372)
      const uint32_t clear = lane ? ((lane & (0u - lane)) - 1u) : 0xFu;  // bit x-1: lane cells 0..x-1 free of own pins
      const uint32_t goal_ok = (jump ? ~lane : clear) & 0xFu;            // B & C indexed by x-1
      const uint32_t win = (0xFu << t) & 0x3Fu;
      row = (row & ~win) | (win & ((circ ? row : 0u) | (goal_ok << t)));
    }
  }
  return row;
}

// valid_action, deterministic variant -> 24-bit mask, bit pin*6 + (move-1)
DS_FN uint32_t madn_det_valid_mask(const MadnGeom& g, const MadnRegs& s) {
  const int pid = s.cur;
  const int cp = mover_of(g, s, pid);
  const uint64_t own = pick4(s.occ, cp);
  const uint32_t pw = pick4(s.pins, cp);
  const uint64_t asr = pick4(s.as, cp);
  const uint32_t posmask = pins_on_start_mask(g, s);
  // home pins: board[start[cp]] != env.current_player  (UN-proxied id, :390)
  const int start_free = !bit64(pick4(s.occ, gidx(pid, g.n)), g.start[cp]) || (pid < 0 || pid >= g.n);
  uint32_t avail = 0;  // action_set[cp][m-1] > 0
#pragma unroll
  for (int m = 1; m <= 6; ++m) avail |= (byte_s64(asr, m - 1) > 0) ? (1u << (m - 1)) : 0u;
  const uint32_t home_moves = DS_RULE(g, DOGSTEP_RULE_START_ON_1) ? 0x21u : 0x20u;
  uint32_t mask = 0;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    int pos = byte_s(pw, i);
    uint32_t row = 0;
    if (pos == -1) {
      row = start_free ? home_moves : 0u;
    } else {
      if (!DS_RULE(g, DOGSTEP_RULE_START_BLOCKING)) {
        row = move_row_fast(g, own, cp, pos);
      } else {
#pragma unroll
        for (int m = 1; m <= 6; ++m) row |= move_ok(g, own, posmask, cp, pos, m) ? (1u << (m - 1)) : 0u;
      }
    }
    mask |= (row & avail) << (6 * i);
  }
  return mask;
}

// valid_action, classic variant -> 4-bit mask
DS_FN uint32_t madn_cls_valid_mask(const MadnGeom& g, const MadnRegs& s) {
  const int pid = s.cur, die = s.die;
  const int cp = mover_of(g, s, pid);
  const uint64_t own = pick4(s.occ, cp);
  const uint32_t pw = pick4(s.pins, cp);
  const uint32_t posmask = pins_on_start_mask(g, s);
  const int home_ok = (die == 6 || (DS_RULE(g, DOGSTEP_RULE_START_ON_1) ? die == 1 : die == -1)) &&
                      !((posmask >> cp) & 1u);
  uint32_t mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int pos = byte_s(pw, i);
    int ok = (pos == -1) ? home_ok : move_ok(g, own, posmask, cp, pos, die);
    mask |= ok ? (1u << i) : 0u;
  }
  return mask;
}

// exact zero-byte detector: 0x80 in every byte of v that is zero
DS_FN uint32_t zero_bytes(uint32_t v) {
  return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu);
}

// Shared move application (deterministic_madn.py:188-230 / classic_madn.py:278-321).
DS_FN void madn_apply_move(const MadnGeom& g, MadnRegs& s, int cp, int pin, int move, int invalid) {
  const int mts = DS_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START);
  const int jump = DS_RULE(g, DOGSTEP_RULE_JUMP_IN_GOAL);
  const uint64_t own = pick4(s.occ, cp);
  const int target = g.target[cp], goal0 = g.goal0[cp];
  const int pos = byte_s(pick4(s.pins, cp), pin);
  const int moved = (int)(int8_t)(pos + move);
  const int fitted = floormod(moved, g.bs);
  const int x = (int)(int8_t)(moved - target - mts);
  const int in_goal = pos >= goal0 && pos <= goal0 + 3;
  const uint32_t lane = (uint32_t)(own >> goal0) & 0xFu;
  const int a = in_goal ? lane_clear(lane, pos - goal0, moved - goal0 + 1) : lane_clear(lane, -1, x);
  const int gxi = gidx(x - 1, 4);
  const int gx = goal0 + gxi;
  const int A = !((lane >> gxi) & 1u) && (jump || a);
  int new_pos;
  if (pos == -1) new_pos = g.start[cp];
  else if (in_goal) new_pos = moved;
  else if (x >= 1 && x <= 4 && A && pos <= target) new_pos = gx;
  else new_pos = fitted;
  if (invalid) return;
  const int cell = gidx(new_pos, g.total);
  int pin_at_pos = -1;
#pragma unroll
  for (int p = 0; p < 4; ++p) pin_at_pos = bit64(s.occ[p], cell) ? p : pin_at_pos;
  const int capture = pin_at_pos != -1 && (pin_at_pos != cp || DS_RULE(g, DOGSTEP_RULE_FRIENDLY_FIRE));
  const uint32_t np4 = (uint32_t)(new_pos & 0xFF) * 0x01010101u;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint32_t w = s.pins[p];
    if (capture && p == pin_at_pos) {
      uint32_t z = zero_bytes(w ^ np4);  // 0x80 where pin == new_pos
      w |= (z >> 7) * 0xFFu;             // -> -1 (home)
    }
    if (p == cp) w = (w & ~(0xFFu << (8 * pin))) | ((uint32_t)(new_pos & 0xFF) << (8 * pin));
    s.pins[p] = w;
  }
  rebuild_occ(g, s);
}

// get_winner (deterministic_madn.py:139-168) -> 4-bit mask
DS_FN uint32_t madn_winner_mask(const MadnGeom& g, const MadnRegs& s) {
  uint64_t any = s.occ[0] | s.occ[1] | s.occ[2] | s.occ[3];
  uint32_t pd = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) pd |= (uint32_t)player_done(g, any, p) << p;
  if (DS_RULE(g, DOGSTEP_RULE_TEAMS)) {
    int t0 = (pd & 5u) == 5u, t1 = (pd & 10u) == 10u;
    if (t0 == t1) return 0u;  // both or none
    return t0 ? 5u : 10u;
  }
  return pd;
}

DS_FN void madn_finish_step(const MadnGeom& g, MadnRegs& s, int pid, int cp, int invalid, int move) {
  uint32_t win = madn_winner_mask(g, s);
  int reward = s.done ? 0 : (invalid ? -1 : (int)((win >> cp) & 1u));
  int done = s.done || (win != 0u);
  int bonus = DS_RULE(g, DOGSTEP_RULE_BONUS_TURN_ON_6) && ((int)(int8_t)move == 6);
  s.cur = (int)(int8_t)((done || bonus) ? pid : floormod(pid + 1, g.n));
  s.reward = reward;
  s.done = done;
}

// env_step, deterministic (deterministic_madn.py:170-257). `valid_bit` = valid_action[pin, move-1].
DS_FN void madn_det_step(const MadnGeom& g, MadnRegs& s, int pin_in, int move_in, uint32_t vmask) {
  const int pid = s.cur;
  const int cp = mover_of(g, s, pid);
  const int pin = gidx(pin_in, 4), mi = gidx(move_in - 1, 6);
  const int invalid = !((vmask >> (pin * 6 + mi)) & 1u);
  uint64_t old_as[4] = {s.as[0], s.as[1], s.as[2], s.as[3]};
  madn_apply_move(g, s, cp, pin, move_in, invalid);
  // action-set bookkeeping with the pre-decrement refill quirk (:232-240, :273-281)
  const int in_range = (move_in - 1 >= -6) && (move_in - 1 < 6);
  uint64_t row = pick4(old_as, cp);
  const int curr = byte_s64(row, mi);
  if (in_range && !(invalid || curr == 0))
    row = (row & ~(0xFFull << (8 * mi))) | ((uint64_t)((curr - 1) & 0xFF) << (8 * mi));
  const int all_zero = (row & 0xFFFFFFFFFFFFull) == 0ull;
  const int refill_ok = pid >= -g.n && pid < g.n;
  const int refill_row = gidx(pid, g.n);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint64_t v = old_as[p];
    if (!all_zero && p == cp) v = row;
    if (all_zero && refill_ok && p == refill_row) v = 0x040404040404ull;
    s.as[p] = v;
  }
  madn_finish_step(g, s, pid, cp, invalid, move_in);
}

// no_step, deterministic (deterministic_madn.py:283-297)
DS_FN void madn_det_no_step(const MadnGeom& g, MadnRegs& s) {
  const int pid = s.cur;
  if (pid >= -g.n && pid < g.n) {
    const int row = gidx(pid, g.n);
#pragma unroll
    for (int p = 0; p < 4; ++p) s.as[p] = (p == row) ? 0x040404040404ull : s.as[p];
  }
  s.cur = (int)(int8_t)floormod(pid + 1, g.n);
}

// env_step, classic (classic_madn.py:257-337)
DS_FN void madn_cls_step(const MadnGeom& g, MadnRegs& s, int pin_in, uint32_t vmask) {
  const int pid = s.cur;
  const int cp = mover_of(g, s, pid);
  const int pin = gidx(pin_in, 4);
  const int invalid = !((vmask >> pin) & 1u);
  madn_apply_move(g, s, cp, pin, s.die, invalid);
  madn_finish_step(g, s, pid, cp, invalid, s.die);
}

// is_soft_locked (classic_madn.py:180-206); uses the UN-proxied current player
DS_FN int madn_soft_locked(const MadnGeom& g, const MadnRegs& s) {
  const int p = gidx(s.cur, g.n);
  const uint32_t pw = pick4(s.pins, p);
  int not_home = 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) not_home -= (byte_s(pw, i) == -1);
  if (not_home <= 0) return 1;
  const uint32_t lane = (uint32_t)(pick4(s.occ, p) >> g.goal0[p]) & 0xFu;
  const uint32_t relevant = (0xFu << (4 - not_home)) & 0xFu;
  return (lane & relevant) == relevant;
}

}  // namespace dogstep
