// dog_fast.cuh — DOG rules specialised for the geometry every configuration of the reference uses (4 players,
// distance 10: ring 0..39, goal lanes 40..55; MuZero_DOG/game_agent.py:12-23) and for CANONICAL states:
//   current_player in 0..3, pins in -1..55, goal-area pins only in their owner's lane, board == set_pins_on_board(pins).
// The board is four 64-bit bitboards derived from the pins (later player wins a shared cell, like the scatter of
// set_pins_on_board), so there is no runtime division, no per-cell loop and no goal-lane scan:
//   `board[c] != cp` is a bit test, the goal-lane path checks are 4-bit masks, the hot-seven path matrix is four
//   64-bit range masks.  The generic restatement (dog_core.cuh, JAX out-of-range semantics included) stays the
//   path for everything else.  Same reference lines as dog_core.cuh:
//   val_swap dog.py:361-390, val_action_7 :393-481, val_action_normal_move :483-566, val_neg_move :568-614,
//   step_swap :755, step_normal_move :790, step_neg_move :861, step_hot_7 :913-984, env_step_play_phase :987-1062,
//   utils/utility_funcs.py:165-319.
// Everything here is __host__ __device__ scalar code: tests/host_core runs it on the CPU against the oracle.
#pragma once
#include "dog_core.cuh"
#include "hostdev.cuh"

namespace dogstep {

DS_FN int dg4_start(int p) { return 10 * p; }
DS_FN int dg4_target(int p) { return p ? 10 * p - 1 : 39; }
DS_FN int dg4_goal(int p) { return 40 + 4 * p; }
DS_FN int dg4_div10(int pos) { return (pos * 26) >> 8; }  // pos / 10 for 0 <= pos <= 68
DS_FN int dg4_mod40(int v) { return v >= 40 ? v - 40 : (v < 0 ? v + 40 : v); }  // v mod 40 for -40 <= v < 80
DS_FN int dg4_bit(uint64_t w, int i) { return (int)((w >> i) & 1ull); }
// cells lo..hi inclusive (0 <= lo, hi <= 63); empty when lo > hi
DS_FN uint64_t dg4_range(int lo, int hi) { return lo > hi ? 0ull : (((hi >= 63) ? ~0ull : ((2ull << hi) - 1ull)) & ~((1ull << lo) - 1ull)); }

struct Dog4Rules {
  bool teams, circ, sb, jump, ff, mts;
};
DS_FN Dog4Rules dg4_rules(uint32_t r) {
  return Dog4Rules{(r & DOGSTEP_RULE_TEAMS) != 0u, (r & DOGSTEP_RULE_CIRCULAR_BOARD) != 0u, (r & DOGSTEP_RULE_START_BLOCKING) != 0u,
                   (r & DOGSTEP_RULE_JUMP_IN_GOAL) != 0u, (r & DOGSTEP_RULE_FRIENDLY_FIRE) != 0u,
                   (r & DOGSTEP_RULE_MUST_TRAVERSE_START) != 0u};
}

// struct Dog4View: dog_core.cuh (it is part of the per-warp record)

DS_FN uint64_t dg4_pin_bits(const int32_t* pins4) {
  uint64_t b = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = pins4[i];
    b |= (pos >= 0 && pos < 56) ? (1ull << pos) : 0ull;
  }
  return b;
}

DS_FN bool dg4_canonical(const int32_t (*pins)[4], const int8_t* board, int cur) {
  bool ok = cur >= 0 && cur <= 3;
  uint64_t bits[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    bits[p] = dg4_pin_bits(pins[p]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pos = pins[p][i];
      ok = ok && pos >= -1 && pos <= 55 && (pos < 40 || (unsigned)(pos - dg4_goal(p)) <= 3u);
    }
  }
  // board == set_pins_on_board(pins): later players overwrite earlier ones
  uint64_t later = 0;
  for (int p = 3; p >= 0; --p) {
    const uint64_t mine = bits[p] & ~later;
    later |= bits[p];
    for (int c = 0; c < 56; ++c) {
      if (((mine >> c) & 1ull) && board[c] != p) ok = false;
    }
  }
  for (int c = 0; c < 56; ++c)
    if (!((later >> c) & 1ull) && board[c] != -1) ok = false;
  return ok;
}

// view from the per-player pin bitboards (bits[p] = cells holding a pin of p) and the mover's pin row
DS_FN void dg4_view_bits(const Dog4Rules& R, const uint64_t* bits, const int32_t (*pins)[4], int pid, Dog4View& v) {
  const uint64_t b0 = bits[0], b1 = bits[1], b2 = bits[2], b3 = bits[3];
  v.occ[3] = b3;
  v.occ[2] = b2 & ~b3;
  v.occ[1] = b1 & ~(b2 | b3);
  v.occ[0] = b0 & ~(b1 | b2 | b3);
  v.any = b0 | b1 | b2 | b3;
  v.pid = pid;
  const bool pid_done = (((uint32_t)(v.any >> 40) >> (4 * pid)) & 0xFu) == 0xFu;  // is_player_done: ANY occupant
  const int cp = (R.teams && pid_done) ? (pid ^ 2) : pid;
  v.cp = cp;
  v.later = (cp < 1 ? b1 : 0ull) | (cp < 2 ? b2 : 0ull) | (cp < 3 ? b3 : 0ull);
  v.posmask = (uint32_t)dg4_bit(v.occ[0], 0) | ((uint32_t)dg4_bit(v.occ[1], 10) << 1) | ((uint32_t)dg4_bit(v.occ[2], 20) << 2) |
              ((uint32_t)dg4_bit(v.occ[3], 30) << 3);
#pragma unroll
  for (int i = 0; i < 4; ++i) v.cur[i] = pins[cp][i];
  const uint64_t own = cp == 0 ? v.occ[0] : cp == 1 ? v.occ[1] : cp == 2 ? v.occ[2] : v.occ[3];
  v.lane = (uint32_t)(own >> dg4_goal(cp)) & 0xFu;
}

DS_FN void dg4_view(const Dog4Rules& R, const int32_t (*pins)[4], int pid, Dog4View& v) {
  uint64_t bits[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) bits[p] = dg4_pin_bits(pins[p]);
  dg4_view_bits(R, bits, pins, pid, v);
}

DS_FN uint64_t dg4_own(const Dog4View& v) { return v.cp == 0 ? v.occ[0] : v.cp == 1 ? v.occ[1] : v.cp == 2 ? v.occ[2] : v.occ[3]; }
DS_FN bool dg4_in_goal(int cp, int pos) { return (unsigned)(pos - dg4_goal(cp)) <= 3u; }
// no own pin on lane cells k with s < k < e  (check_goal_path_for_pin, utility_funcs.py:165-184)
DS_FN bool dg4_lane_clear(uint32_t lane, int s, int e) {
  const int lo = s + 1 < 0 ? 0 : s + 1, hi = e - 1 > 3 ? 3 : e - 1;
  const uint32_t m = (lo > hi) ? 0u : (((2u << hi) - 1u) & ~((1u << lo) - 1u));
  return (lane & m) == 0u;
}

// val_swap (:361-390): result[i][cell] = pin_ok bit i & cell_ok bit cell
DS_FN void dg4_val_swap(const Dog4Rules& R, const Dog4View& v, uint32_t& pin_ok, uint64_t& cell_ok) {
  const int cp = v.cp;
  uint64_t ok = v.any & ~dg4_own(v);
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // start cells: occupied and not held by a blocking owner
    const int cell = dg4_start(q);
    const bool val = !(dg4_bit(v.occ[q], cell) && R.sb) && dg4_bit(v.any, cell);
    ok = (ok & ~(1ull << cell)) | ((uint64_t)val << cell);
  }
  uint32_t po = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = v.cur[i];
    if (pos >= 0) ok &= ~(1ull << pos);
    const bool bad = (pos == -1) || dg4_in_goal(cp, pos) || (R.sb && pos == dg4_start(cp));
    po |= (uint32_t)(!bad) << i;
  }
  ok &= 0xFFFFFFFFFFull;  // no goal cells
  pin_ok = po;
  cell_ok = ok;
}

// val_action_normal_move for one pin (:483-566)
DS_FN bool dg4_val_normal(const Dog4Rules& R, const Dog4View& v, int i, int move) {
  const int cp = v.cp, target = dg4_target(cp), g0 = dg4_goal(cp);
  const int pos = v.cur[i];
  if (move <= 0) return false;
  if (pos == -1) return (move == 1 || move == 11 || move == 13) && !((v.posmask >> cp) & 1u);
  const uint64_t own = dg4_own(v);
  const int moved = pos + move;
  if (dg4_in_goal(cp, pos)) {
    const bool D = R.jump || dg4_lane_clear(v.lane, pos - g0, moved - g0 + 1);
    return moved <= g0 + 3 && !dg4_bit(own, moved & 63) && D;
  }
  const int fitted = dg4_mod40(moved);
  int x = moved - target - (int)R.mts;
  bool result = !dg4_bit(own, fitted) || R.ff;
  const int nsb = (dg4_div10(pos) + 1) & 3, nsa = dg4_div10(fitted);
  const bool trav = nsb == nsa, blocked = (v.posmask >> nsa) & 1u;
  if (R.sb && trav) result = (!blocked || pos == dg4_start(cp)) && result;
  if (R.mts && R.sb && trav && blocked) x = 0;
  if (!R.circ && pos <= target && (x > 4 || (x == 0 && R.mts))) result = false;
  if (x >= 1 && x <= 4 && pos <= target) {
    const bool B = !((v.lane >> (x - 1)) & 1u);
    const bool C = R.jump || dg4_lane_clear(v.lane, -1, x);
    result = (R.circ && result) || (B && C);
  }
  return result;
}

// val_neg_move for one pin (:568-614)
DS_FN bool dg4_val_neg(const Dog4Rules& R, const Dog4View& v, int i, int move) {
  const int cp = v.cp;
  const int pos = v.cur[i];
  if (pos == -1 || dg4_in_goal(cp, pos)) return false;
  const int moved = pos + move, fitted = dg4_mod40(moved);
  bool result = !dg4_bit(dg4_own(v), fitted) || R.ff;
  const int nsb = dg4_div10(pos), nsa = (dg4_div10(fitted) + 1) & 3;
  if (R.sb && nsb == nsa) result = (!((v.posmask >> nsa) & 1u) || pos == dg4_start(cp)) && result;
  return result && (R.circ || moved >= dg4_start(cp));
}

DS_FN int dg4_sgn(int a) { return (a > 0) - (a < 0); }

// own bits of cp's goal lane on the board with cp's in-goal pins already at their moved cells (:454-455, :934-935)
DS_FN uint32_t dg4_tmp_lane(const Dog4View& v, const int moved[4]) {
  uint64_t b = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = dg4_in_goal(v.cp, v.cur[i]) ? moved[i] : v.cur[i];
    b |= (pos >= 0 && pos < 56) ? (1ull << pos) : 0ull;
  }
  return (uint32_t)((b & ~v.later) >> dg4_goal(v.cp)) & 0xFu;
}

// val_action_7 -> scalar (:393-481)
DS_FN bool dg4_val_7(const Dog4Rules& R, const Dog4View& v, const int dist[4]) {
  const int cp = v.cp, target = dg4_target(cp), g0 = dg4_goal(cp), start = dg4_start(cp);
  int moved[4];
  bool own_start_stays = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    moved[i] = v.cur[i] + dist[i];
    own_start_stays = own_start_stays || (v.cur[i] == start && dist[i] == 0);
  }
  const uint32_t tmp_lane = R.jump ? 0u : dg4_tmp_lane(v, moved);
  bool all = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cur = v.cur[i], mv = moved[i];
    bool result;
    if (cur == -1) {
      result = dist[i] == 0;  // a home pin must not move (:480-481); its other tests are vacuous
    } else if (dg4_in_goal(cp, cur)) {
      bool order_ok = true;   // check_relative_order_preserved (utility_funcs.py:186-234)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (v.cur[j] >= 40 && dg4_sgn(cur - v.cur[j]) != dg4_sgn(mv - moved[j])) order_ok = false;
      result = mv <= g0 + 3 && (R.jump || order_ok);
    } else {
      const int fitted = dg4_mod40(mv);
      int x = mv - target - (int)R.mts;
      result = R.circ ? true : !((cur <= target) && ((mv > target + 4) || (x == 0 && R.mts)));
      if (R.sb) {
        const int nsb = (dg4_div10(cur) + 1) & 3, nsa = dg4_div10(fitted);
        const bool trav = nsb == nsa;
        const bool blocked = (nsa == cp) ? own_start_stays : (bool)((v.posmask >> nsa) & 1u);
        if (trav) result = !blocked && result;
        if (R.mts && trav && blocked) x = 0;
      }
      if (x >= 1 && x <= 4 && cur <= target) {
        const bool C = R.jump || dg4_lane_clear(tmp_lane, -1, x);
        result = (R.circ && result) || C;
      }
    }
    all = all && result;
  }
  return all;
}

// the 120 splits of 7 in lexicographic (a,b,c) order (utility_funcs.py:4-21), d = 7-a-b-c
DS_FN void dg4_dist_of(int k, int d[4]) {
  // first index of each a-block: T(8-a) = (8-a)(9-a)/2 entries
  int a = 0, rem = k;
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int cnt = (8 - t) * (9 - t) / 2;
    const bool go = rem >= cnt && a == t;
    rem -= go ? cnt : 0;
    a += go ? 1 : 0;
  }
  int b = 0;
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int cnt = 8 - a - t;
    const bool go = rem >= cnt && b == t && cnt > 0;
    rem -= go ? cnt : 0;
    b += go ? 1 : 0;
  }
  d[0] = a; d[1] = b; d[2] = rem; d[3] = 7 - a - b - rem;
}

// card that pays for base action b (valid_step_actions :618-691); category: 0 swap, 1 hot seven, 2 normal, 3 -4
DS_FN int dg4_card_of_base(int b) {
  if (b < 224) return 1;
  if (b < 344) return 7;
  if (b < 392) {
    const int k = (b - 344) % 12;
    int move = k + 1;
    move += (move >= 7);
    return k == 0 ? 11 : move;
  }
  return 4;
}

// validity of base action b in [0, 396) ignoring the hand
DS_FN bool dg4_base_valid(const Dog4Rules& R, const Dog4View& v, int b, uint32_t pin_ok, uint64_t cell_ok) {
  if (b < 224) {
    const int i = (b * 1171) >> 16;  // b / 56 for b < 224
    const int c = b - 56 * i;
    return ((pin_ok >> i) & 1u) && ((cell_ok >> c) & 1ull);
  }
  if (b < 344) {
    int d[4];
    dg4_dist_of(b - 224, d);
    return dg4_val_7(R, v, d);
  }
  if (b < 392) {
    const int na = b - 344;
    const int i = (na * 43) >> 9;  // na / 12 for na < 48
    int move = na - 12 * i + 1;
    move += (move >= 7);
    return dg4_val_normal(R, v, i, move);
  }
  return dg4_val_neg(R, v, b - 392, -4);
}

// ---- state transition of the play phase (env_step_play_phase :987-1062 up to the turn hand-over) on the pins ------
// Returns reward; `done_out` = env.done | any winner.  Pins are updated in place; the caller rebuilds the board.
struct Dog4StepOut {
  int reward, done, cp;
};

DS_FN uint32_t dg4_winner_mask(const Dog4Rules& R, uint64_t any) {
  const uint32_t lanes = (uint32_t)(any >> 40) & 0xFFFFu;
  const uint32_t full = lanes & (lanes >> 1) & (lanes >> 2) & (lanes >> 3) & 0x1111u;
  const uint32_t pd = (full & 1u) | ((full >> 3) & 2u) | ((full >> 6) & 4u) | ((full >> 9) & 8u);
  if (R.teams) {
    const bool t0 = (pd & 5u) == 5u, t1 = (pd & 10u) == 10u;
    if (t0 == t1) return 0u;
    return t0 ? 5u : 10u;
  }
  return pd;
}

DS_FN uint64_t dg4_any_bits(const int32_t (*pins)[4]) {
  return dg4_pin_bits(pins[0]) | dg4_pin_bits(pins[1]) | dg4_pin_bits(pins[2]) | dg4_pin_bits(pins[3]);
}

// capture whoever owns the landing cell (own pins only under friendly fire) and place the pin (:327-336 of the oracle,
// dog.py step_normal_move / step_neg_move)
DS_FN void dg4_capture_and_place(const Dog4Rules& R, const Dog4View& v, int32_t (*pins)[4], int pin, int new_pos) {
  int owner = -1;
#pragma unroll
  for (int p = 0; p < 4; ++p) owner = dg4_bit(v.occ[p], new_pos) ? p : owner;
  if (owner != -1 && (owner != v.cp || R.ff)) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (pins[owner][k] == new_pos) pins[owner][k] = -1;
  }
  pins[v.cp][pin] = new_pos;
}

// one play-phase action on the pins: returns invalid flag (pins untouched when invalid)
// TRUSTED: the action was drawn from this position's own legal mask (the random-policy drivers), so the validity tests — the
// same dg4_val_* calls that built the mask — cannot fail and are skipped; an arbitrary caller-supplied action is never trusted.
template <bool TRUSTED = false>
DS_FN bool dg4_apply_play_action(const Dog4Rules& R, const Dog4View& v, int32_t (*pins)[4], const int mv[6]) {
  const int cp = v.cp, target = dg4_target(cp), g0 = dg4_goal(cp);
  const int* d = mv + 2;
  if (mv[1] == 1) {  // step_swap (:755-788)
    const int pi = (d[0] >= 0) ? 0 : (d[1] >= 0) ? 1 : (d[2] >= 0) ? 2 : (d[3] >= 0) ? 3 : 0;
    const int sp = d[pi] < 0 ? d[pi] + 56 : d[pi];  // gather index: negative wraps once (only -1 can occur)
    if (!TRUSTED) {
      uint32_t pin_ok;
      uint64_t cell_ok;
      dg4_val_swap(R, v, pin_ok, cell_ok);
      if (!(((pin_ok >> pi) & 1u) && ((cell_ok >> sp) & 1ull))) return true;
    }
    int swapped = -1;
#pragma unroll
    for (int p = 0; p < 4; ++p) swapped = dg4_bit(v.occ[p], sp) ? p : swapped;
    const int pin_pos = v.cur[pi];
    pins[cp][pi] = d[pi];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (pins[swapped][k] == d[pi]) pins[swapped][k] = pin_pos;
    return false;
  }
  if (d[0] + d[1] + d[2] + d[3] == 7) {  // step_hot_7 (:913-984)
    if (!TRUSTED && !dg4_val_7(R, v, d)) return true;
    int moved[4], nw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) moved[i] = v.cur[i] + d[i];
    const uint32_t tmp_lane = dg4_tmp_lane(v, moved);
    bool any_diff = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cur = v.cur[i];
      const int x = moved[i] - target - (int)R.mts;
      const bool ig = dg4_in_goal(cp, cur);
      const bool A = R.jump || ig || dg4_lane_clear(tmp_lane, -1, x);
      int np;
      if (cur == -1) np = -1;
      else if (ig) np = moved[i];
      else if (x >= 1 && x <= 4 && A && cur <= target) np = g0 + x - 1;
      else np = dg4_mod40(moved[i]);
      nw[i] = np;
      any_diff = any_diff || (ig != dg4_in_goal(cp, np));
    }
    // get_path_matrix rows (utility_funcs.py:237-303) as range masks
    uint64_t M[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cur = v.cur[i], np = nw[i];
      const bool A = dg4_in_goal(cp, cur), B = dg4_in_goal(cp, np);
      uint64_t row = 0;
      if (A == B) {
        if (cur != -1 && np != -1 && cur != np) {
          row = (cur <= np) ? dg4_range(cur, np) : (dg4_range(cur, 63) | dg4_range(0, np));
          row &= 0xFFFFFFFFFFull;  // only ring columns
        }
      } else {
        uint64_t ring = 0;
        if (cur != -1) ring = (cur <= target) ? dg4_range(cur, target) : (dg4_range(cur, 63) | dg4_range(0, target));
        ring &= 0xFFFFFFFFFFull;
        uint64_t goal = 0;
        if (np != -1) goal = (g0 <= np) ? dg4_range(g0, np) : (dg4_range(g0, 63) | dg4_range(0, np));
        goal &= 0xFFFFFFFFFFFFFFull;  // 56 cells
        row = ring | goal;
      }
      if (any_diff) row |= 1ull << dg4_start(cp);
      M[i] = row;
    }
    const uint64_t anyrow = M[0] | M[1] | M[2] | M[3];
    uint32_t hit = 0;  // bit p*4+i; a home pin reads cell 55 (gather index -1 wraps) and stays home either way
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pos = pins[p][i];
        hit |= (uint32_t)dg4_bit(anyrow, pos < 0 ? 55 : pos) << (p * 4 + i);
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // check_moving_pins_hit (:310-319)
      const uint64_t others = (i == 0 ? 0ull : M[0]) | (i == 1 ? 0ull : M[1]) | (i == 2 ? 0ull : M[2]) | (i == 3 ? 0ull : M[3]);
      const int h = dg4_bit(others, v.cur[i] < 0 ? 55 : v.cur[i]) & dg4_bit(others, nw[i] < 0 ? 55 : nw[i]);
      hit = (hit & ~(1u << (cp * 4 + i))) | ((uint32_t)h << (cp * 4 + i));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) pins[cp][i] = nw[i];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if ((hit >> (p * 4 + i)) & 1u) pins[p][i] = -1;
    return false;
  }
  const int pi = (d[0] != 0) ? 0 : (d[1] != 0) ? 1 : (d[2] != 0) ? 2 : (d[3] != 0) ? 3 : 0;
  const int move = d[pi], pos = v.cur[pi];
  if (move < 0) {  // step_neg_move (:861-911)
    if (!TRUSTED && !dg4_val_neg(R, v, pi, move)) return true;
    dg4_capture_and_place(R, v, pins, pi, dg4_mod40(pos + move));
    return false;
  }
  if (!TRUSTED && !dg4_val_normal(R, v, pi, move)) return true;  // step_normal_move (:790-859)
  const int moved = pos + move;
  const int x = moved - target - (int)R.mts;
  const bool ig = dg4_in_goal(cp, pos);
  const bool a = ig ? dg4_lane_clear(v.lane, pos - g0, moved - g0 + 1) : dg4_lane_clear(v.lane, -1, x);
  const int xi = (x - 1) < 0 ? ((x - 1 + 4) < 0 ? 0 : x - 1 + 4) : ((x - 1) > 3 ? 3 : x - 1);  // gather index into goal[cp]
  const bool A = !((v.lane >> xi) & 1u) && (R.jump || a);
  int new_pos;
  if (pos == -1) new_pos = dg4_start(cp);
  else if (ig) new_pos = moved;
  else if (x >= 1 && x <= 4 && A && pos <= target) new_pos = g0 + xi;
  else new_pos = dg4_mod40(moved);
  dg4_capture_and_place(R, v, pins, pi, new_pos);
  return false;
}

// map_action_to_move (:1134-1196) for half = 396, total = 56
DS_FN void dg4_map_action_to_move(int action, int mv[6]) {
  const int is_joker = (action - 396) < 0;
  int act = action % 396;
  act = act < 0 ? act + 396 : act;
  const int is_swap = act < 224;
  int d[4] = {0, 0, 0, 0};
  if (is_swap) {
    d[0] = d[1] = d[2] = d[3] = -1;
    const int k = (act * 1171) >> 16;  // act / 56
    const int c = act - 56 * k;
#pragma unroll
    for (int q = 0; q < 4; ++q) d[q] = (q == k) ? c : d[q];
  } else if (act < 344) {
    dg4_dist_of(act - 224, d);
  } else if (act < 392) {
    const int na = act - 344;
    const int i = (na * 43) >> 9;
    int move = na - 12 * i + 1;
    move += (move >= 7);
#pragma unroll
    for (int q = 0; q < 4; ++q) d[q] = (q == i) ? move : 0;
  } else {
    const int k = act - 392;
#pragma unroll
    for (int q = 0; q < 4; ++q) d[q] = (q == k) ? -4 : 0;
  }
  mv[0] = is_joker; mv[1] = is_swap;
  mv[2] = d[0]; mv[3] = d[1]; mv[4] = d[2]; mv[5] = d[3];
}

// sum of a 16-byte hand row (two zero pad bytes)
DS_FN int dg4_hand_sum(const int8_t* row) {
#ifdef __CUDA_ARCH__
  const int4 w = *reinterpret_cast<const int4*>(row);  // 16-byte row, pad bytes are zero
  return __dp4a(w.x, 0x01010101, __dp4a(w.y, 0x01010101, __dp4a(w.z, 0x01010101, __dp4a(w.w, 0x01010101, 0))));
#else
  int sum = 0;
  for (int k = 0; k < kNCard; ++k) sum += row[k];
  return sum;
#endif
}

// first seat after `from` that still holds cards (dog.py:1043-1047, :733-741)
DS_FN int dg4_next_with_cards(const DogS& s, int from, int& all_empty, int& any_left) {
  int sums[4];
  int nz = 0, pos = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    sums[q] = dg4_hand_sum(s.hands[q]);
    nz |= (sums[q] != 0);
    pos |= (sums[q] > 0);
  }
  int next = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cand = (from + i + 1) & 3;
    const int sc = cand == 0 ? sums[0] : cand == 1 ? sums[1] : cand == 2 ? sums[2] : sums[3];
    if (next == -1 && sc > 0) next = cand;
  }
  all_empty = !nz;
  any_left = pos;
  return next;
}

// env_step_play_phase (:987-1062) for a canonical state and action in [0, 792): pins, hands, cur, reward, done are
// updated; s.board is NOT (the caller rebuilds it from the pins before anything reads it).  Returns "deal next".
template <bool TRUSTED = false>
DS_FN int dg4_play_phase(const Dog4Rules& R, DogS& s, int action, int& reward_out, int& done_out) {
  Dog4View local;
  if (!TRUSTED) {
#ifdef __CUDA_ARCH__
    dg4_view_bits(R, s.pbits, s.pins, s.cur, local);  // kept current by dog4_rebuild_board_warp
#else
    dg4_view(R, s.pins, s.cur, local);
#endif
  }
  const Dog4View& v = TRUSTED ? s.view : local;  // TRUSTED: the mask of this very position was just built, and with it s.view
  const int pid = s.cur, cp = v.cp;
  int mv[6];
  dg4_map_action_to_move(action, mv);
  const int card = dog_map_action_to_card(mv);
  const int ci = card < 0 ? (card + kNCard < 0 ? 0 : card + kNCard) : (card > kNCard - 1 ? kNCard - 1 : card);
  int reward, done;
  if (s.hands[cp][ci] > 0) {
    const bool invalid = dg4_apply_play_action<TRUSTED>(R, v, s.pins, mv);
    const uint32_t w = dg4_winner_mask(R, dg4_any_bits(s.pins));
    done = s.done || (w != 0u);
    reward = s.done ? 0 : (invalid ? -1 : (int)((w >> cp) & 1u));
  } else {
    reward = -1;
    done = s.done;
  }
  if (card >= 0 && card < kNCard) s.hands[cp][card] = (int8_t)(s.hands[cp][card] + (reward == -1 ? 0 : -1));
  int all_empty, any_left;
  const int next = dg4_next_with_cards(s, pid, all_empty, any_left);
  s.cur = done ? cp : next;  // the PROXIED id is kept when the game ends (:1048)
  s.reward = reward;
  s.done = done;
  reward_out = reward;
  done_out = done;
  return (all_empty || next == -1) && !done;
}

// ---------------------------------------------------------------------------------------------------------------
// warp-collective pieces (device only): one warp per game, state in the per-warp shared record DogS
#ifdef __CUDACC__

// the 120 splits of 7 (utility_funcs.py:4-21), packed a | b<<3 | c<<6 | d<<9; same order as dg4_dist_of
__device__ const uint16_t g_dog_splits7[120] = {3584, 3136, 2688, 2240, 1792, 1344, 896, 448, 3080, 2632, 2184, 1736, 1288, 840, 392, 2576, 2128, 1680, 1232, 784, 336, 2072, 1624, 1176, 728, 280, 1568, 1120, 672, 224, 1064, 616, 168, 560, 112, 56, 3073, 2625, 2177, 1729, 1281, 833, 385, 2569, 2121, 1673, 1225, 777, 329, 2065, 1617, 1169, 721, 273, 1561, 1113, 665, 217, 1057, 609, 161, 553, 105, 49, 2562, 2114, 1666, 1218, 770, 322, 2058, 1610, 1162, 714, 266, 1554, 1106, 658, 210, 1050, 602, 154, 546, 98, 42, 2051, 1603, 1155, 707, 259, 1547, 1099, 651, 203, 1043, 595, 147, 539, 91, 35, 1540, 1092, 644, 196, 1036, 588, 140, 532, 84, 28, 1029, 581, 133, 525, 77, 21, 518, 70, 14, 7};

// is the staged record canonical (see header)?  Also publishes the pin bitboards s.pbits.  All 32 lanes must call;
// warp-uniform result.
__device__ __forceinline__ bool dog4_canonical_warp(DogS& s, int lane) {
  const uint32_t FULL = 0xFFFFFFFFu;
  uint64_t bits[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) bits[p] = dg4_pin_bits(s.pins[p]);
  const uint64_t o3 = bits[3], o2 = bits[2] & ~bits[3], o1 = bits[1] & ~(bits[2] | bits[3]), o0 = bits[0] & ~(bits[1] | bits[2] | bits[3]);
  bool ok = s.cur >= 0 && s.cur <= 3;
  if (lane < 16) {
    const int p = lane >> 2, pos = s.pins[p][lane & 3];
    ok = ok && pos >= -1 && pos <= 55 && (pos < 40 || (unsigned)(pos - dg4_goal(p)) <= 3u);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int c = lane + 32 * r;
    if (c < 56) {
      const int want = dg4_bit(o3, c) ? 3 : dg4_bit(o2, c) ? 2 : dg4_bit(o1, c) ? 1 : dg4_bit(o0, c) ? 0 : -1;
      ok = ok && (int)s.board[c] == want;
    }
  }
  if (lane < 4) s.pbits[lane] = bits[lane == 0 ? 0 : lane == 1 ? 1 : lane == 2 ? 2 : 3];
  const bool all_ok = __all_sync(FULL, ok);
  __syncwarp();
  return all_ok;
}

// board bytes from the pins (set_pins_on_board: later players overwrite earlier ones).  All lanes call.
__device__ __forceinline__ void dog4_rebuild_board_warp(DogS& s, int lane) {
  __syncwarp();
  if (lane < 16) reinterpret_cast<uint32_t*>(s.board)[lane] = 0xFFFFFFFFu;
  else if (lane < 20) s.pbits[lane - 16] = dg4_pin_bits(s.pins[lane - 16]);
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (lane < 4) {
      const int pos = s.pins[p][lane];
      if (pos >= 0 && pos < 56) s.board[pos] = (int8_t)p;
    }
    __syncwarp();
  }
}

// the bitboards alone (what the dog_fast.cuh rules read); the board BYTES are then stale until dog4_rebuild_board_warp
__device__ __forceinline__ void dog4_refresh_pbits(DogS& s, int lane) {
  __syncwarp();
  if (lane < 4) s.pbits[lane] = dg4_pin_bits(s.pins[lane]);
  __syncwarp();
}

__device__ __forceinline__ void dog4_set_pair(DogS& s, int b, bool joker, bool card) {
  if (joker) atomicOr(&s.mask[b >> 5], 1u << (b & 31));
  if (card) {
    const int a = 396 + b;
    atomicOr(&s.mask[a >> 5], 1u << (a & 31));
  }
}

// valid_actions (dog.py:693-711) into s.mask for a canonical record.  Only the categories the hand can pay for are
// evaluated (a base action contributes a bit only if the joker or its own card is held), the 120 hot-seven splits
// and the 48 normal moves are dealt round-robin to the lanes.  All 32 lanes must call it.
__device__ __forceinline__ void dog4_build_mask(const Dog4Rules& R, DogS& s, int lane) {
  for (int w = lane; w < kDogMaskWords; w += 32) s.mask[w] = 0u;
  __syncwarp();
  if (s.phase != 0) {
    if (lane < kNCard && s.hands[s.cur][lane] > 0) {
      const int a = 792 + lane;
      atomicOr(&s.mask[a >> 5], 1u << (a & 31));
    }
    __syncwarp();
    return;
  }
  Dog4View v;
  dg4_view_bits(R, s.pbits, s.pins, s.cur, v);
  if (lane == 0) s.view = v;  // a TRUSTED transition of this position reads it (dg4_play_phase)
  const int8_t* hand = s.hands[v.cp];
  const bool joker = hand[0] > 0;
  if (joker || hand[1] > 0) {  // swaps: pin_ok x cell_ok
    uint32_t pin_ok;
    uint64_t cell_ok;
    dg4_val_swap(R, v, pin_ok, cell_ok);
    const bool card = hand[1] > 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int c = lane + 32 * r;
      if (c < 40 && ((cell_ok >> c) & 1ull)) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((pin_ok >> i) & 1u) dog4_set_pair(s, 56 * i + c, joker, card);
      }
    }
  }
  if (joker || hand[7] > 0) {  // hot seven: 120 splits
    const bool card = hand[7] > 0;
    for (int it = lane; it < 120; it += 32) {
      const uint32_t e = g_dog_splits7[it];
      const int d[4] = {(int)(e & 7u), (int)((e >> 3) & 7u), (int)((e >> 6) & 7u), (int)((e >> 9) & 7u)};
      if (dg4_val_7(R, v, d)) dog4_set_pair(s, 224 + it, joker, card);
    }
  }
  for (int it = lane; it < 48; it += 32) {  // normal moves: pin x {1..6, 8..13}
    const int i = (it * 43) >> 9, k = it - 12 * i;
    int move = k + 1;
    move += (move >= 7);
    const int cid = (k == 0) ? 11 : move;  // move 1 is the low face of card 11 (dog.py:660-670)
    const bool card = hand[cid] > 0;
    if ((joker || card) && dg4_val_normal(R, v, i, move)) dog4_set_pair(s, 344 + it, joker, card);
  }
  if (lane < 4 && (joker || hand[4] > 0) && dg4_val_neg(R, v, lane, -4)) dog4_set_pair(s, 392 + lane, joker, hand[4] > 0);
  __syncwarp();
}

// The same mask cut into tasks that ANY warp of the CTA can run on a record (bits are OR-ed into s.mask atomically), so that
// the four 32-split chunks of the hot seven — the heavy part, needed only when the hand holds a seven or a joker — can be
// shared out to warps whose own game has a light hand (k_dog_play_random).  sub 0..3 = the hot-seven chunks, 4..5 = the normal
// moves (the second chunk also carries the four -4 moves), 6 = swaps.  `flags` (dog4_mask_flags): bit 0 swaps, bit 1 hot
// seven, bit 2 moves.  s must be a canonical record in the play phase with s.mask zeroed.
__device__ __forceinline__ int dog4_mask_flags(const Dog4Rules& R, DogS& s, int lane) {
  Dog4View v;
  dg4_view_bits(R, s.pbits, s.pins, s.cur, v);
  if (lane == 0) s.view = v;  // read by dog4_mask_task, whichever warp runs it (the caller's barrier publishes it)
  __syncwarp();
  const int8_t* hand = s.hands[v.cp];
  const bool joker = hand[0] > 0;
  return ((joker || hand[1] > 0) ? 1 : 0) | ((joker || hand[7] > 0) ? 2 : 0) | 4;
}

__device__ __forceinline__ void dog4_mask_task(const Dog4Rules& R, DogS& s, int sub, int lane) {
  const Dog4View& v = s.view;
  const int8_t* hand = s.hands[v.cp];
  const bool joker = hand[0] > 0;
  if (sub < 4) {  // hot seven: splits 32 sub .. 32 sub + 31
    const int it = 32 * sub + lane;
    if (it < 120) {
      const uint32_t e = g_dog_splits7[it];
      const int d[4] = {(int)(e & 7u), (int)((e >> 3) & 7u), (int)((e >> 6) & 7u), (int)((e >> 9) & 7u)};
      if (dg4_val_7(R, v, d)) dog4_set_pair(s, 224 + it, joker, hand[7] > 0);
    }
  } else if (sub < 6) {  // normal moves: pin x {1..6, 8..13}; then the -4 move of each pin
    const int it = 32 * (sub - 4) + lane;
    if (it < 48) {
      const int i = (it * 43) >> 9, k = it - 12 * i;
      int move = k + 1;
      move += (move >= 7);
      const int cid = (k == 0) ? 11 : move;  // move 1 is the low face of card 11 (dog.py:660-670)
      const bool card = hand[cid] > 0;
      if ((joker || card) && dg4_val_normal(R, v, i, move)) dog4_set_pair(s, 344 + it, joker, card);
    } else if (it < 52) {
      const int i = it - 48;
      if ((joker || hand[4] > 0) && dg4_val_neg(R, v, i, -4)) dog4_set_pair(s, 392 + i, joker, hand[4] > 0);
    }
  } else {  // swaps: pin_ok x cell_ok
    uint32_t pin_ok;
    uint64_t cell_ok;
    dg4_val_swap(R, v, pin_ok, cell_ok);
    const bool card = hand[1] > 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int c = lane + 32 * r;
      if (c < 40 && ((cell_ok >> c) & 1ull)) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((pin_ok >> i) & 1u) dog4_set_pair(s, 56 * i + c, joker, card);
      }
    }
  }
}

// env_step (dog.py:1117-1131) for a canonical record: the swap phase and the deal are the generic code (they do not
// touch the board), the play phase runs on the bitboards; the board bytes are rebuilt from the pins afterwards.
// All lanes call; lane 0 applies the move, all lanes deal if needed.
// LAZY_BOARD (the persistent play kernel): only the bitboards follow the move; the caller rebuilds the board bytes before it
// stores the game (nothing on the canonical path reads them in between).
template <bool TRUSTED = false, bool LAZY_BOARD = false>
__device__ __forceinline__ void dog4_env_step(const Dog4Rules& R, const DogGeom& g, DogS& s, int lane, int action, int& reward_out,
                                              int& done_out) {
  __syncwarp();
  const bool play = s.phase != 1;
  if (play && (action < 0 || action >= 792)) {  // a swap-phase index in the play phase: generic path (warp-uniform)
    dog_env_step(g, s, lane, action, reward_out, done_out);
    dog4_rebuild_board_warp(s, lane);  // refreshes s.pbits
    return;
  }
  if (lane == 0) {
    if (play) {
      int reward, done;
      s.scratch[0] = dg4_play_phase<TRUSTED>(R, s, action, reward, done);
      s.scratch[1] = reward;
      s.scratch[2] = done;
    } else {
      dog_swap_phase(g, s, action);
      s.scratch[0] = 0;
      s.scratch[1] = 0;
      s.scratch[2] = s.done;
    }
  }
  __syncwarp();
  if (!play) {
    reward_out = s.scratch[1];
    done_out = s.scratch[2];
    __syncwarp();
    return;
  }
  if (LAZY_BOARD) dog4_refresh_pbits(s, lane);
  else dog4_rebuild_board_warp(s, lane);
  if (s.scratch[0] && !s.scratch[6]) dog_distribute_cards(g, s, lane);  // scratch[6]: the caller deals (k_dog_play_random)
  reward_out = s.scratch[1];
  done_out = s.scratch[2];
  __syncwarp();
}

#endif  // __CUDACC__

}  // namespace dogstep
