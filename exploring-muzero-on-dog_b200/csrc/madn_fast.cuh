// madn_fast.cuh — deterministic MADN rules specialised for the geometry every training / benchmark
// configuration of the reference uses (4 players, distance 10: ring 0..39, goal lanes 40..55;
// MuZero_det_MADN/game_agent.py:24-41) and for CANONICAL states, i.e. states the reference's own
// env_reset / env_step can produce:
//   current_player in 0..3, pins in -1..55, goal-lane pins only in their owner's lane, no cell shared
//   by two players, board == set_pins_on_board(pins).  (Two pins of ONE player may share a cell: a
//   team-proxied home exit onto the mover's own pin is legal in the reference because the start test
//   looks at the un-proxied player, deterministic_madn.py:390 — about 7 % of random games do it.)
// Under those conditions the generic restatement in madn_core.cuh (which also carries JAX's
// out-of-range gather/scatter semantics for arbitrary inputs) collapses to branch-free bit logic:
// no runtime divisions, the board is updated incrementally instead of being rebuilt from the pins,
// and the six moves of a pin are one 6-bit row.  is_canonical4() is the gate; non-canonical games
// keep using madn_core.cuh.  Same reference lines:
//   valid_action  MADN/deterministic_madn.py:299-393     env_step  :170-257
//   get_winner    :122-168     _refill_action_set quirk  :232-240, :273-281
// Every function is __host__ __device__ so tests/ can run it on the CPU against the oracle.
#pragma once
#include "madn_core.cuh"

namespace dogstep {

constexpr uint32_t kRulesRuntime = 0x80000000u;  // template value: read the rule mask at run time

// compile-time rule mask (one program per rule dict, like XLA) or a warp-uniform run-time mask
template <uint32_t CT>
struct RuleSet {
  uint32_t rt;
  DS_FN bool on(uint32_t bit) const { return CT == kRulesRuntime ? (rt & bit) != 0u : (CT & bit) != 0u; }
};

DS_FN int start4(int p) { return 10 * p; }
DS_FN int target4(int p) { return p ? 10 * p - 1 : 39; }  // (start - 1) mod 40
DS_FN int goal4(int p) { return 40 + 4 * p; }

// gate for the fast path (see header)
DS_FN bool is_canonical4(const MadnRegs& s, const uint64_t board_occ[4]) {
  bool ok = s.cur >= 0 && s.cur <= 3;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint64_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pos = byte_s(s.pins[p], i);
      ok = ok && pos >= -1 && pos <= 55 && (pos < 40 || (unsigned)(pos - goal4(p)) <= 3u);
      if (pos >= 0) bits |= 1ull << (pos & 63);
    }
    ok = ok && bits == board_occ[p];  // also rules out a cell shared by two players (the board holds one of them)
  }
  return ok;
}

// positive bytes of a 48-bit action-set row -> 6-bit mask (action_set[cp][m-1] > 0)
DS_FN uint32_t avail6(uint64_t row) {
  const uint32_t lo = (uint32_t)row, hi = (uint32_t)(row >> 32) & 0xFFFFu;
  const uint32_t nzl = (((lo & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | lo) & ~lo & 0x80808080u;
  const uint32_t nzh = (((hi & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | hi) & ~hi & 0x00008080u;
  return (((nzl >> 7) * 0x01020408u) >> 24) | ((nzh >> 3) & 0x10u) | ((nzh >> 10) & 0x20u);
}

// valid_action -> 24-bit mask (bit pin*6 + move-1); also returns the team-proxied mover
template <uint32_t CT>
DS_FN uint32_t det_valid_mask4(const RuleSet<CT> R, const MadnGeom& g, const MadnRegs& s, int& cp_out) {
  const int pid = s.cur;
  const uint64_t any = s.occ[0] | s.occ[1] | s.occ[2] | s.occ[3];
  const bool pid_done = (((uint32_t)(any >> 40) >> (4 * pid)) & 0xFu) == 0xFu;  // is_player_done: occupied by ANYONE (:136)
  const int cp = (R.on(DOGSTEP_RULE_TEAMS) && pid_done) ? (pid ^ 2) : pid;      // (:184,310)
  cp_out = cp;
  const uint64_t own = pick4(s.occ, cp);
  const uint32_t pw = pick4(s.pins, cp);
  const uint32_t avail = avail6(pick4(s.as, cp));
  const int target = target4(cp), goal0 = goal4(cp);
  const bool mts = R.on(DOGSTEP_RULE_MUST_TRAVERSE_START), circ = R.on(DOGSTEP_RULE_CIRCULAR_BOARD);
  const bool jump = R.on(DOGSTEP_RULE_JUMP_IN_GOAL);
  // home pins: board[start[cp]] != env.current_player — the UN-proxied id (:390)
  const bool start_free = !((pick4(s.occ, pid) >> start4(cp)) & 1ull);
  const uint32_t home_row = start_free ? (R.on(DOGSTEP_RULE_START_ON_1) ? 0x21u : 0x20u) : 0u;
  const uint32_t lane = (uint32_t)(own >> goal0) & 0xFu;
  const uint64_t ring = own & 0xFFFFFFFFFFull;
  const uint64_t ring2 = ring | (ring << 40);  // ring cell (c mod 40) at bit c, c < 64
  // bit x-1: goal cell x-1 free of own pins and (jump or lane cells 0..x-1 all free) (:360-372)
  const uint32_t lowclear = lane ? ((lane & (0u - lane)) - 1u) : 0xFu;
  const uint32_t goal_ok = (jump ? ~lane : lowclear) & 0xFu;
  const uint32_t posmask = R.on(DOGSTEP_RULE_START_BLOCKING) ? pins_on_start_mask(g, s) : 0u;
  uint32_t mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = byte_s(pw, i);
    uint32_t row;
    if (R.on(DOGSTEP_RULE_START_BLOCKING)) {  // rare rule set: per-move generic check
      row = 0;
      if (pos >= 0) {
#pragma unroll
        for (int m = 1; m <= 6; ++m) row |= move_ok(g, own, posmask, cp, pos, m) ? (1u << (m - 1)) : 0u;
      }
    } else {
      // ring pin: landing cell not own (:329)
      const uint32_t landing_own = (uint32_t)(ring2 >> ((pos + 1) & 63)) & 0x3Fu;
      row = R.on(DOGSTEP_RULE_FRIENDLY_FIRE) ? 0x3Fu : (~landing_own & 0x3Fu);
      // x = m - t with t = target + mts - pos  (t >= 0 whenever pos <= target)
      const uint32_t t = ds_min_u32((uint32_t)(target + (int)mts - pos), 8u);
      uint32_t row_t = row;
      if (!circ) {  // overshoot x > 4  <=>  m > t+4 ; x == 0 && mts  <=>  m == t  (:349-357)
        const uint32_t beyond = (0xFFFFFFFFu << (t + 4u)) & 0x3Fu;
        const uint32_t at_target = mts ? ((1u << t) >> 1) : 0u;
        row_t &= ~(beyond | at_target);
      }
      const uint32_t win = (0xFu << t) & 0x3Fu;  // goal-entry window 1 <= x <= 4
      row_t = (row_t & ~win) | (win & ((circ ? row_t : 0u) | (goal_ok << t)));
      row = (pos <= target) ? row_t : row;
      // pin inside its goal lane (:376-381)
      const uint32_t k0 = (uint32_t)(pos - goal0);
      const uint32_t above = lane >> ((k0 + 1u) & 7u);
      const uint32_t upclear = above ? ((above & (0u - above)) - 1u) : 0xFu;
      const uint32_t row_goal = (jump ? ~above : upclear) & ((1u << ((3u - k0) & 3u)) - 1u);
      row = (k0 <= 3u) ? row_goal : row;
    }
    row = (pos < 0) ? home_row : row;
    mask |= (row & avail) << (6 * i);
  }
  return mask;
}

// env_step for a VALID action a = pin*6 + move-1 of a live canonical game (:170-257); the result is canonical.
template <uint32_t CT>
DS_FN void det_step4(const RuleSet<CT> R, MadnRegs& s, int cp, int a) {
  const int pid = s.cur;
  const int pin = (a * 43) >> 8;  // a / 6 for a < 24   (map_action :469-479)
  const int mi = a - 6 * pin, move = mi + 1;
  const bool mts = R.on(DOGSTEP_RULE_MUST_TRAVERSE_START), jump = R.on(DOGSTEP_RULE_JUMP_IN_GOAL);
  const bool ff = R.on(DOGSTEP_RULE_FRIENDLY_FIRE);
  const int target = target4(cp), goal0 = goal4(cp);
  const uint64_t own = pick4(s.occ, cp);
  const uint32_t pw = pick4(s.pins, cp);
  const uint32_t lane = (uint32_t)(own >> goal0) & 0xFu;
  const int pos = byte_s(pw, pin);
  const int moved = pos + move;
  const int fitted = moved >= 40 ? moved - 40 : moved;
  const int x = moved - target - (int)mts;
  const bool in_goal = (unsigned)(pos - goal0) <= 3u;
  const uint32_t xi = (uint32_t)(x - 1) & 3u;
  const bool A = jump ? !((lane >> xi) & 1u) : ((lane & ((2u << xi) - 1u)) == 0u);
  int new_pos = fitted;
  new_pos = (x >= 1 && x <= 4 && A && pos <= target) ? goal0 + x - 1 : new_pos;
  new_pos = in_goal ? moved : new_pos;
  new_pos = (pos < 0) ? start4(cp) : new_pos;
  // capture (:205-216): whoever stands on the landing cell goes home; an own pin only under friendly fire
  const uint64_t nbit = 1ull << new_pos;
  const uint64_t obit = (pos < 0) ? 0ull : (1ull << (pos & 63));
  const uint32_t np4 = (uint32_t)new_pos * 0x01010101u;
  const uint32_t op4 = (uint32_t)(pos & 0xFF) * 0x01010101u;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const bool hit = (s.occ[p] & nbit) != 0ull && (p != cp || ff);
    uint32_t w = s.pins[p];
    const uint32_t z = zero_bytes(w ^ np4);  // 0x80 where pin == new_pos
    w = hit ? (w | ((z >> 7) * 0xFFu)) : w;  // -> -1 (home)
    uint64_t o = s.occ[p] & ~nbit;
    if (p == cp) {
      w = (w & ~(0xFFu << (8 * pin))) | ((uint32_t)new_pos << (8 * pin));
      const bool vacated = zero_bytes(w ^ op4) == 0u;  // no other own pin stacked on the old cell
      o = (vacated ? (o & ~obit) : o) | nbit;
    }
    s.pins[p] = w;
    s.occ[p] = o;
  }
  // action set with the pre-decrement refill quirk (:232-240, :273-281)
  const uint64_t row = pick4(s.as, cp) - (1ull << (8 * mi));  // valid => count > 0
  const bool all_zero = (row & 0xFFFFFFFFFFFFull) == 0ull;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint64_t v = s.as[p];
    v = (!all_zero && p == cp) ? row : v;
    v = (all_zero && p == pid) ? 0x040404040404ull : v;
    s.as[p] = v;
  }
  // get_winner (:139-168) on the new board
  const uint32_t lanes = (uint32_t)((s.occ[0] | s.occ[1] | s.occ[2] | s.occ[3]) >> 40) & 0xFFFFu;
  const uint32_t full = lanes & (lanes >> 1) & (lanes >> 2) & (lanes >> 3) & 0x1111u;  // bit 4p: player p done
  uint32_t win;
  if (R.on(DOGSTEP_RULE_TEAMS)) {
    const bool t0 = (full & 0x0101u) == 0x0101u, t1 = (full & 0x1010u) == 0x1010u;
    win = (t0 == t1) ? 0u : (t0 ? 0x0101u : 0x1010u);
  } else {
    win = full;
  }
  const int done = win != 0u;
  s.reward = (int)((win >> (4 * cp)) & 1u);
  s.done = done;
  const bool bonus = R.on(DOGSTEP_RULE_BONUS_TURN_ON_6) && move == 6;
  s.cur = (done || bonus) ? pid : ((pid + 1) & 3);
}

// no_step (:283-297)
DS_FN void det_no_step4(MadnRegs& s) {
  const int pid = s.cur;
#pragma unroll
  for (int p = 0; p < 4; ++p) s.as[p] = (p == pid) ? 0x040404040404ull : s.as[p];
  s.cur = (pid + 1) & 3;
}

}  // namespace dogstep
