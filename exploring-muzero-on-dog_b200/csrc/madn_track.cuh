// madn_track.cuh — deterministic MADN under the rule dict of every training / benchmark configuration
// (MuZero_det_MADN/game_agent.py:12-22: teams, initial free pin, jump in goal area, start on 1, bonus turn on 6; no circular
// board, no friendly fire, no start blocking, no must-traverse-start), 4 players, distance 10, canonical states (madn_fast.cuh).
//
// Under that rule set a pin never leaves its owner's TRACK: the 40 ring cells counted from the owner's start cell, followed by
// the owner's four goal cells — 44 cells, -1 = home.  In track coordinates the reference's case analysis
//   ring pin before / after its target cell, goal entry window 1 <= x <= 4, overshoot, pin inside the goal lane
//   (MADN/deterministic_madn.py:299-393 valid_action, :170-257 env_step)
// collapses to "u + move <= 43 and no own pin on u + move", and env_step's new position to u + move.  What is left per ply:
// the mover's 44-bit occupancy (four shifts), one 6-bit window per pin, and — for captures — the landing cell expressed on
// each other player's track (+10 cells per seat).  The board is not kept at all: it is set_pins_on_board(pins) (:259-271)
// and rebuilt when a game is stored.  State per game: eight 32-bit words (pins as bytes, action-set counts as nibbles).
// Same results as madn_fast.cuh / madn_core.cuh on every canonical live state (tests/test_madn_fast_core.py runs this file on
// the CPU against the oracle); track_from_regs() is the gate.
#pragma once
#include "madn_fast.cuh"

namespace dogstep {

constexpr uint32_t kTrainRules = DOGSTEP_RULE_TEAMS | DOGSTEP_RULE_INITIAL_FREE_PIN | DOGSTEP_RULE_JUMP_IN_GOAL |
                                 DOGSTEP_RULE_START_ON_1 | DOGSTEP_RULE_BONUS_TURN_ON_6;

struct Track4 {
  uint32_t pins[4];  // byte i of word p: pin i of player p on p's track (0..39 ring from p's start, 40..43 goal), -1 = home
  uint32_t as[4];    // nibble m-1 of word p: copies of move m left in player p's action set (0..7)
  int cur, reward, done;
};

// 0x80 in every byte that holds 40..43 (bytes are -1 or 0..43)
DS_FN uint32_t track_in_goal(uint32_t w) { return ((w & 0x7F7F7F7Fu) + 0x58585858u) & ~w & 0x80808080u; }
DS_FN bool track_player_done(uint32_t w) { return track_in_goal(w) == 0x80808080u; }  // is_player_done (:122-137)

// nonzero nibbles 0..5 -> 6-bit mask (action_set[cp][m-1] > 0)
DS_FN uint32_t track_avail6(uint32_t v) {
  uint32_t t = (v | (v >> 1) | (v >> 2)) & 0x111111u;
  t = (t | (t >> 3)) & 0x030303u;
  return (t | (t >> 6) | (t >> 12)) & 0x3Fu;
}

// track position of player p -> board cell (-1 = home)
DS_FN int track_to_cell(int p, int u) {
  int pos = u + start4(p);
  pos = pos >= 40 ? pos - 40 : pos;
  pos = u >= 40 ? u - 40 + goal4(p) : pos;
  return u < 0 ? -1 : pos;
}

// nibble k -> byte k (one action-set row as the reference stores it)
DS_FN uint64_t track_as_bytes(uint32_t v) {
  uint32_t lo = v & 0xFFFFu, hi = (v >> 16) & 0xFFu;
  lo = (lo | (lo << 8)) & 0x00FF00FFu;
  lo = (lo | (lo << 4)) & 0x0F0F0F0Fu;
  hi = (hi | (hi << 4)) & 0x0F0Fu;
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

// MadnRegs (absolute cells) -> track state; false if the state is outside what this file covers (the caller then keeps the
// generic rules): not canonical, a count outside 0..7, or a team already complete (the reference would report that winner)
DS_FN bool track_from_regs(const MadnRegs& r, Track4& s) {
  bool ok = is_canonical4(r, r.occ);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint32_t w = 0, a = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pos = byte_s(r.pins[p], i);
      int u = pos - start4(p);
      u = u < 0 ? u + 40 : u;
      u = pos >= 40 ? pos - goal4(p) + 40 : u;
      u = pos < 0 ? -1 : u;
      w |= (uint32_t)(u & 0xFF) << (8 * i);
    }
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int c = byte_s64(r.as[p], m);
      ok = ok && c >= 0 && c <= 7;
      a |= (uint32_t)(c & 0xF) << (4 * m);
    }
    s.pins[p] = w;
    s.as[p] = a;
  }
  s.cur = r.cur;
  s.reward = r.reward;
  s.done = r.done;
  const bool t0 = track_player_done(s.pins[0]) && track_player_done(s.pins[2]);
  const bool t1 = track_player_done(s.pins[1]) && track_player_done(s.pins[3]);
  return ok && !t0 && !t1;
}

DS_FN void track_to_regs(const MadnGeom& g, const Track4& s, MadnRegs& r) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) w |= (uint32_t)(track_to_cell(p, byte_s(s.pins[p], i)) & 0xFF) << (8 * i);
    r.pins[p] = w;
    r.as[p] = track_as_bytes(s.as[p]);
  }
  r.cur = s.cur;
  r.reward = s.reward;
  r.done = s.done;
  r.die = 0;
  rebuild_occ(g, r);  // board = set_pins_on_board(pins) (:259-271)
}

// valid_action (:299-393) -> 24-bit mask (bit pin*6 + move-1); also returns the team-proxied mover (:184,310)
DS_FN uint32_t track_valid_mask(const Track4& s, int& cp_out) {
  const int pid = s.cur;
  const bool pid_done = track_player_done(pick4(s.pins, pid));
  const int cp = pid_done ? (pid ^ 2) : pid;
  cp_out = cp;
  const uint32_t pw = pick4(s.pins, cp);
  const uint32_t avail = track_avail6(pick4(s.as, cp));
  uint64_t own = 0;  // the mover's pins on its track; a home pin sets bit 63, which no window below reaches
#pragma unroll
  for (int i = 0; i < 4; ++i) own |= 1ull << ((pw >> (8 * i)) & 63u);
  // home pins leave on 1 or 6 if board[start[cp]] != env.current_player — the UN-proxied player (:390), whose pins are all in
  // its goal when it moves for its partner
  const uint32_t home_row = (pid_done || !(own & 1ull)) ? 0x21u : 0u;
  uint32_t mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int u = byte_s(pw, i);
    const uint32_t landing_own = (uint32_t)(own >> ((u + 1) & 63)) & 0x3Fu;  // own pin on u + m (:329, :360-381 with jump)
    const uint32_t reach = 0x3Fu >> (uint32_t)max(u - 37, 0);                // u + m <= 43: no overshoot (:349-357)
    uint32_t row = ~landing_own & reach;
    row = (u < 0) ? home_row : row;
    mask |= (row & avail) << (6 * i);
  }
  return mask;
}

// env_step for a VALID action a = pin*6 + move-1 of a live state (:170-257)
DS_FN void track_step(Track4& s, int cp, int a) {
  const int pid = s.cur;
  const int pin = (a * 43) >> 8;  // a / 6 for a < 24   (map_action :469-479)
  const int mi = a - 6 * pin, move = mi + 1;
  const uint32_t pw = pick4(s.pins, cp);
  const int u = byte_s(pw, pin);
  const int nu = (u < 0) ? 0 : u + move;
  const uint32_t pwn = (pw & ~(0xFFu << (8 * pin))) | ((uint32_t)nu << (8 * pin));
  // capture (:205-216): the other players' pins on the landing cell go home (own pins only under friendly fire).  Ring cell
  // `cell` (absolute) is (cell - 10 q) mod 40 on player q's track; goal cells belong to one player only.
  int cell = nu + start4(cp);
  cell = cell >= 40 ? cell - 40 : cell;
  const bool on_ring = nu < 40;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int t = cell - start4(q);
    t = t < 0 ? t + 40 : t;
    t = (on_ring && q != cp) ? t : 0x7F;  // no byte holds 0x7F
    uint32_t w = s.pins[q];
    const uint32_t z = zero_bytes(w ^ ((uint32_t)t * 0x01010101u));  // 0x80 where pin == landing cell
    w |= (z >> 7) * 0xFFu;                                           // -> -1 (home)
    s.pins[q] = (q == cp) ? pwn : w;
  }
  // action set with the pre-decrement refill quirk (:232-240, :273-281)
  const uint32_t row = pick4(s.as, cp) - (1u << (4 * mi));  // valid => count > 0
  const bool all_zero = row == 0u;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint32_t v = s.as[p];
    v = (!all_zero && p == cp) ? row : v;
    v = (all_zero && p == pid) ? 0x444444u : v;
    s.as[p] = v;
  }
  // get_winner (:139-168): the state was live, so only the mover's team can have completed
  const int done = track_player_done(pwn) && track_player_done(pick4(s.pins, cp ^ 2));
  s.reward = done;
  s.done = done;
  s.cur = (done || move == 6) ? pid : ((pid + 1) & 3);  // bonus turn on 6 (:242-250)
}

// no_step (:283-297)
DS_FN void track_no_step(Track4& s) {
  const int pid = s.cur;
#pragma unroll
  for (int p = 0; p < 4; ++p) s.as[p] = (p == pid) ? 0x444444u : s.as[p];
  s.cur = (pid + 1) & 3;
}

}  // namespace dogstep
