// dog_core.cuh — the DOG environment (2v2 card game) with one WARP per game.
//
// A game's mutable state (208 B) is staged once into a per-warp shared-memory record (DogS) and
// stays there; the 32 lanes split the wide work — the 396+396+14 legal-action mask (120 hot-seven
// splits dominate), the Threefry draws of the categorical policy and the 120-slot deck shuffle —
// while the short, branchy state transition of the chosen action runs on lane 0.
// Reference semantics restated here (bit-exact, quirks included — SURVEY Appendix A.6/A.6b):
//   DOG/dog.py: distribute_cards :201-298, val_swap :361-390, val_action_7 :393-481,
//   val_action_normal_move :483-566, val_neg_move :568-614, valid_step_actions :618-691,
//   valid_actions :693-711, no_step :713-752, step_swap :755, step_normal_move :790,
//   step_neg_move :861, step_hot_7 :913-984, env_step_play_phase :987-1062,
//   env_step_swap_phase :1078-1114, map_action_to_move :1134-1196, map_action_to_card :1241-1262
//   utils/utility_funcs.py: all_pin_distributions :4-21, check_goal_path_for_pin :165-184,
//   check_relative_order_preserved :186-234, get_path_matrix :237-303, check_moving_pins_hit :310-319
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/dogstep.h"
#include "jaxrand.cuh"
#include "hostdev.cuh"

namespace dogstep {

constexpr int kNCard = 14;
constexpr int kDogMaskWords = 28;  // >= ceil((2*(4*(13+64)+120)+14)/32)

struct DogGeom {
  int n, d, bs, total;
  int start[4], target[4], goal0[4];
  uint32_t rules;
  int play_actions, half, num_actions;
};

inline int dog_make_geom(const dogstep_madn_cfg* cfg, DogGeom* g) {
  if (!cfg) return DOGSTEP_ERR_INVALID_ARG;
  if (cfg->num_players < 2 || cfg->num_players > 4 || cfg->distance < 1) return DOGSTEP_ERR_INVALID_ARG;
  if (cfg->distance > 12) return DOGSTEP_ERR_UNSUPPORTED;
  // the reference's disable_* switches shrink num_cards but keep hard-coded card ids (dog.py:139, 660-666);
  // only the all-enabled deck is self-consistent and only that one is implemented.
  if (cfg->rules & (DOGSTEP_RULE_DISABLE_SWAPPING | DOGSTEP_RULE_DISABLE_HOT_SEVEN | DOGSTEP_RULE_DISABLE_JOKER))
    return DOGSTEP_ERR_UNSUPPORTED;
  g->n = cfg->num_players;
  g->d = cfg->distance;
  g->bs = 4 * cfg->distance;
  g->total = g->bs + 16;
  uint32_t rules = cfg->rules;
  if (g->n != 4) rules &= ~DOGSTEP_RULE_TEAMS;
  g->rules = rules;
  g->play_actions = 2 * (4 * (13 + g->total) + 120);
  g->half = g->play_actions / 2;
  g->num_actions = g->play_actions + kNCard;
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

// what the rules read of a position, for the (team-proxied) mover
struct Dog4View {
  int pid, cp;
  int cur[4];        // the mover's pins
  uint64_t occ[4];   // occ[p] bit c  <=>  board[c] == p
  uint64_t any;      // board[c] != -1
  uint64_t later;    // pins of players > cp (they win a shared cell on any rebuilt board)
  uint32_t posmask;  // bit q: board[start[q]] == q
  uint32_t lane;     // bit k: board[goal[cp][k]] == cp
};

// per-warp shared record
struct alignas(16) DogS {
  int32_t pins[4][4];
  int8_t board[64];
  int8_t hands[4][16];
  int8_t deck[16];
  int8_t swap_choices[4];
  uint32_t key[2];
  int cur, reward, done, round_starter, phase, hand_size;
  uint32_t mask[kDogMaskWords];
  uint16_t items[2 * (4 * (13 + 64) + 120) + kNCard + 2];
  int scratch[8];
  uint64_t pbits[4];  // bit c of pbits[p]: a pin of player p stands on cell c (kept by the dog_fast.cuh path)
  Dog4View view;      // the mover's view of this turn (dog4_mask_flags): the mask tasks of ANY warp read it instead of rebuilding it
};

#define DG_RULE(g, bit) (((g).rules & (bit)) != 0u)

DS_FN int d_fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
DS_FN int d_fmod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }
DS_FN int d_gidx(int i, int size) { i = (i < 0) ? i + size : i; return min(max(i, 0), size - 1); }
DS_FN int d_sidx(int i, int size) { i = (i < 0) ? i + size : i; return (i < 0 || i >= size) ? -1 : i; }

// lexicographic (a,b,c) enumeration of the 120 splits of 7 (utility_funcs.py:4-21) -> split k
DS_FN void dog_dist_of(int k, int d[4]) {
  // number of (b,c) with b+c <= 7-a is T(8-a) = (8-a)(9-a)/2; cumulative over a
  int a = 0, rem = k;
  while (true) {
    int cnt = (8 - a) * (9 - a) / 2;
    if (rem < cnt) break;
    rem -= cnt;
    ++a;
  }
  int b = 0;
  while (true) {
    int cnt = 8 - a - b;
    if (rem < cnt) break;
    rem -= cnt;
    ++b;
  }
  d[0] = a; d[1] = b; d[2] = rem; d[3] = 7 - a - b - rem;
}

DS_FN void dog_set_pins_on_board(const DogGeom& g, const int32_t pins[4][4], int8_t* board) {
  for (int k = 0; k < g.total; ++k) board[k] = -1;
  for (int p = 0; p < g.n; ++p)
    for (int i = 0; i < 4; ++i) {
      int pos = pins[p][i];
      if (pos >= 0 && pos < g.total) board[pos] = (int8_t)p;
    }
}

DS_FN int dog_player_done(const DogGeom& g, const int8_t* board, int player) {
  if (player >= g.n) return 0;
  int p = d_gidx(player, g.n);
  for (int k = 0; k < 4; ++k)
    if (board[g.goal0[p] + k] < 0) return 0;
  return 1;
}

DS_FN uint32_t dog_winner_mask(const DogGeom& g, const int8_t* board) {
  uint32_t pd = 0;
  for (int p = 0; p < 4; ++p) pd |= (uint32_t)dog_player_done(g, board, p) << p;
  if (DG_RULE(g, DOGSTEP_RULE_TEAMS)) {
    int t0 = (pd & 5u) == 5u, t1 = (pd & 10u) == 10u;
    if (t0 == t1) return 0u;
    return t0 ? 5u : 10u;
  }
  return pd;
}

DS_FN int dog_mover(const DogGeom& g, const DogS& s) {
  int pid = s.cur;
  int cp = (DG_RULE(g, DOGSTEP_RULE_TEAMS) && dog_player_done(g, s.board, pid)) ? ((pid + 2) & 3) : pid;
  return d_gidx(cp, g.n);
}

DS_FN int dog_in_goal(const DogGeom& g, int cp, int pos) { return pos >= g.goal0[cp] && pos <= g.goal0[cp] + 3; }

DS_FN int dog_path_clear(const DogGeom& g, const int8_t* board, int cp, int s, int e) {
  for (int k = 0; k < 4; ++k)
    if (s < k && k < e && board[g.goal0[cp] + k] == cp) return 0;
  return 1;
}

DS_FN int dog_pins_on_start(const DogGeom& g, const int8_t* board, int q) {
  q = d_gidx(q, g.n);
  return board[g.start[q]] == q;
}

// val_swap: result[i][cell] = pin_ok bit i & cell_ok bit cell
DS_FN void dog_val_swap(const DogGeom& g, const DogS& s, int cp, uint32_t& pin_ok, uint64_t& cell_ok) {
  const int sb = DG_RULE(g, DOGSTEP_RULE_START_BLOCKING);
  uint64_t ok = 0;
  for (int k = 0; k < g.total; ++k) ok |= (uint64_t)(s.board[k] != -1 && s.board[k] != cp) << k;
  for (int q = 0; q < g.n; ++q) {
    int cell = g.start[q];
    uint64_t v = (uint64_t)(!((s.board[cell] == q) && sb) && (s.board[cell] != -1));
    ok = (ok & ~(1ull << cell)) | (v << cell);
  }
  for (int i = 0; i < 4; ++i) {
    int col = d_sidx(s.pins[cp][i], g.total);
    if (col >= 0) ok &= ~(1ull << col);
  }
  for (int q = 0; q < g.n; ++q) ok &= ~(0xFull << g.goal0[q]);
  uint32_t po = 0;
  for (int i = 0; i < 4; ++i) {
    int pos = s.pins[cp][i];
    int bad = (pos == -1) || dog_in_goal(g, cp, pos) || (sb && pos == g.start[cp]);
    po |= (uint32_t)(!bad) << i;
  }
  pin_ok = po;
  cell_ok = ok;
}

// val_action_normal_move for ONE pin
DS_FN int dog_val_normal(const DogGeom& g, const DogS& s, int cp, int i, int move) {
  const int8_t* board = s.board;
  const int mts = DG_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START), sb = DG_RULE(g, DOGSTEP_RULE_START_BLOCKING);
  const int circ = DG_RULE(g, DOGSTEP_RULE_CIRCULAR_BOARD), jump = DG_RULE(g, DOGSTEP_RULE_JUMP_IN_GOAL);
  const int target = g.target[cp], g0 = g.goal0[cp];
  int pos = s.pins[cp][i];
  int moved = pos + move, fitted = d_fmod(moved, g.bs);
  int x = moved - target - mts;
  int result = (board[fitted] != cp) || DG_RULE(g, DOGSTEP_RULE_FRIENDLY_FIRE);
  int nsb = d_fmod(d_fdiv(pos, g.d) + 1, g.n), nsa = fitted / g.d;
  int trav = g.start[d_gidx(nsb, g.n)] == g.start[d_gidx(nsa, g.n)];
  int blocked = dog_pins_on_start(g, board, nsa);
  if (sb && trav) result = (!blocked || pos == g.start[cp]) && result;
  if (mts && sb && trav && blocked) x = 0;
  if (!circ && pos <= target && (x > 4 || (x == 0 && mts))) result = 0;
  if (4 >= x && x > 0 && pos <= target) {
    int B = board[g0 + x - 1] != cp;
    int C = jump || dog_path_clear(g, board, cp, -1, x);
    result = (circ && result) || (B && C);
  }
  if (dog_in_goal(g, cp, pos)) {
    int D = jump || dog_path_clear(g, board, cp, pos - g0, moved - g0 + 1);
    result = (moved <= g0 + 3) && (board[d_gidx(moved, g.total)] != cp) && D;
  }
  if (pos == -1) result = (move == 1 || move == 11 || move == 13) && !dog_pins_on_start(g, board, cp);
  return result && move > 0;
}

// val_neg_move for ONE pin
DS_FN int dog_val_neg(const DogGeom& g, const DogS& s, int cp, int i, int move) {
  const int8_t* board = s.board;
  int pos = s.pins[cp][i];
  int moved = pos + move, fitted = d_fmod(moved, g.bs);
  int result = (board[fitted] != cp) || DG_RULE(g, DOGSTEP_RULE_FRIENDLY_FIRE);
  int nsb = d_fdiv(pos, g.d), nsa = d_fmod(fitted / g.d + 1, g.n);
  int cond = g.start[d_gidx(nsb, g.n)] == g.start[d_gidx(nsa, g.n)];
  if (DG_RULE(g, DOGSTEP_RULE_START_BLOCKING) && cond)
    result = (!dog_pins_on_start(g, board, nsa) || pos == g.start[cp]) && result;
  result = result && (DG_RULE(g, DOGSTEP_RULE_CIRCULAR_BOARD) || moved >= g.start[cp]);
  if (pos == -1 || dog_in_goal(g, cp, pos)) result = 0;
  return result;
}

DS_FN int d_sgn(int v) { return (v > 0) - (v < 0); }

// own-goal-lane occupancy of tmp_board (dog.py:454-455, 934-935): the board with cp's in-goal pins already
// at their moved cells.  Only cp's lane cells matter for the path check, so 4 bits suffice.
DS_FN uint32_t dog_hot7_tmp_lane(const DogGeom& g, const DogS& s, int cp, const int moved[4]) {
  // set_pins_on_board: later players overwrite earlier ones on the same cell
  const int g0 = g.goal0[cp];
  uint32_t lane = 0;
  for (int k = 0; k < 4; ++k) {
    int owner = -1;
    for (int p = 0; p < g.n; ++p)
      for (int i = 0; i < 4; ++i) {
        int pos = s.pins[p][i];
        if (p == cp && dog_in_goal(g, cp, pos)) pos = moved[i];
        if (pos == g0 + k) owner = p;
      }
    lane |= (uint32_t)(owner == cp) << k;
  }
  return lane;
}

DS_FN int d_lane_clear(uint32_t lane, int s, int e) {
  for (int k = 0; k < 4; ++k)
    if (s < k && k < e && ((lane >> k) & 1u)) return 0;
  return 1;
}

// val_action_7 -> scalar
DS_FN int dog_val_7(const DogGeom& g, const DogS& s, int cp, const int dist[4]) {
  const int8_t* board = s.board;
  const int mts = DG_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START), sb = DG_RULE(g, DOGSTEP_RULE_START_BLOCKING);
  const int circ = DG_RULE(g, DOGSTEP_RULE_CIRCULAR_BOARD), jump = DG_RULE(g, DOGSTEP_RULE_JUMP_IN_GOAL);
  const int target = g.target[cp], g0 = g.goal0[cp];
  int cur[4], moved[4];
  int own_start_stays = 0;
  for (int i = 0; i < 4; ++i) {
    cur[i] = s.pins[cp][i];
    moved[i] = cur[i] + dist[i];
    if (cur[i] == g.start[cp] && moved[i] == g.start[cp]) own_start_stays = 1;
  }
  const uint32_t tmp_lane = jump ? 0u : dog_hot7_tmp_lane(g, s, cp, moved);
  int all = 1;
  for (int i = 0; i < 4; ++i) {
    int fitted = d_fmod(moved[i], g.bs);
    int x = moved[i] - target - mts;
    int result = circ ? 1 : !((cur[i] <= target) && ((moved[i] > target + 4) || (x == 0 && mts)));
    if (sb) {
      int nsb = d_fmod(d_fdiv(cur[i], g.d) + 1, g.n), nsa = d_gidx(fitted / g.d, g.n);
      int trav = g.start[d_gidx(nsb, g.n)] == g.start[nsa];
      int blocked = (nsa == cp) ? own_start_stays : (board[g.start[nsa]] == nsa);
      if (trav) result = !blocked && result;
      if (mts && trav && blocked) x = 0;
    }
    if (4 >= x && x > 0 && cur[i] <= target) {
      int C = jump || d_lane_clear(tmp_lane, -1, x);
      result = (circ && result) || C;
    }
    if (dog_in_goal(g, cp, cur[i])) {
      int order_ok = 1;  // check_relative_order_preserved
      for (int j = 0; j < 4; ++j)
        if (cur[j] >= g.bs && d_sgn(cur[i] - cur[j]) != d_sgn(moved[i] - moved[j])) order_ok = 0;
      result = (moved[i] <= g0 + 3) && (jump || order_ok);
    }
    int board_mover = (cur[i] == -1) ? (moved[i] == -1) : 1;
    all = all && result && board_mover;
  }
  return all;
}

// ---- legal mask, lane-parallel: base action b in [0, half) -> validity ignoring the card in hand ----
DS_FN int dog_base_action_valid(const DogGeom& g, const DogS& s, int cp, int b, uint32_t pin_ok,
                                                      uint64_t cell_ok, int& card) {
  const int pxb = 4 * g.total;
  if (b < pxb) {
    card = 1;
    int i = b / g.total, c = b - i * g.total;
    return (int)((pin_ok >> i) & 1u) & (int)((cell_ok >> c) & 1ull);
  }
  if (b < pxb + 120) {
    card = 7;
    int d[4];
    dog_dist_of(b - pxb, d);
    return dog_val_7(g, s, cp, d);
  }
  if (b < g.half - 4) {
    int na = b - pxb - 120;
    int i = na / 12, k = na - i * 12;
    int move = k + 1;
    move += (move >= 7);
    card = (k == 0) ? 11 : move;  // move 1 is the low face of card 11 (dog.py:660-670)
    return dog_val_normal(g, s, cp, i, move);
  }
  card = 4;
  return dog_val_neg(g, s, cp, b - (g.half - 4), -4);
}

// valid_actions (dog.py:693-711) into s.mask (bit a = action a legal).  All 32 lanes must call it.
__device__ __noinline__ void dog_build_mask(const DogGeom& g, DogS& s, int lane) {
  const uint32_t FULL = 0xFFFFFFFFu;
  for (int w = lane; w < kDogMaskWords; w += 32) s.mask[w] = 0u;
  __syncwarp();
  if (s.phase != 0) {
    if (lane < kNCard) {
      int row = d_gidx(s.cur, g.n);
      if (s.hands[row][lane] > 0) {
        int a = g.play_actions + lane;
        atomicOr(&s.mask[a >> 5], 1u << (a & 31));
      }
    }
    __syncwarp();
    return;
  }
  const int cp = dog_mover(g, s);
  uint32_t pin_ok;
  uint64_t cell_ok;
  dog_val_swap(g, s, cp, pin_ok, cell_ok);
  const int8_t* hand = s.hands[cp];
  const int has_joker = hand[0] > 0;
  // hot-seven splits are ~10x the cost of the other actions: deal them round-robin first, then the rest
  const int pxb = 4 * g.total;
  for (int it = lane; it < g.half; it += 32) {
    // permute so that each lane gets an equal share of the 120 expensive splits
    int b = (it < 120) ? pxb + it : (it < 120 + pxb ? it - 120 : it);
    int card;
    int v = dog_base_action_valid(g, s, cp, b, pin_ok, cell_ok, card);
    if (v) {
      if (has_joker) atomicOr(&s.mask[b >> 5], 1u << (b & 31));
      if (hand[card] > 0) {
        int a = g.half + b;
        atomicOr(&s.mask[a >> 5], 1u << (a & 31));
      }
    }
  }
  __syncwarp(FULL);
}

// ---- state transition of one action: lane 0 only -----------------------------------------------------
DS_FN void dog_finish_substep(const DogGeom& g, const DogS& s, int cp, int invalid, int& reward, int& done) {
  uint32_t w = dog_winner_mask(g, s.board);
  done = s.done || (w != 0u);
  reward = s.done ? 0 : (invalid ? -1 : (int)((w >> cp) & 1u));
}

DS_FN void dog_capture_and_place(const DogGeom& g, DogS& s, int cp, int pin, int new_pos) {
  int pin_at_pos = s.board[d_gidx(new_pos, g.total)];
  if (pin_at_pos != -1 && (pin_at_pos != cp || DG_RULE(g, DOGSTEP_RULE_FRIENDLY_FIRE))) {
    int q = d_gidx(pin_at_pos, g.n);
    for (int k = 0; k < 4; ++k)
      if (s.pins[q][k] == new_pos) s.pins[q][k] = -1;
  }
  s.pins[cp][pin] = new_pos;
  dog_set_pins_on_board(g, s.pins, s.board);
}

__host__ __device__ inline void dog_step_swap(const DogGeom& g, DogS& s, int pin_idx, int swap_pos, int& reward, int& done) {
  int cp = dog_mover(g, s);
  uint32_t pin_ok;
  uint64_t cell_ok;
  dog_val_swap(g, s, cp, pin_ok, cell_ok);
  int pi = d_gidx(pin_idx, 4), sp = d_gidx(swap_pos, g.total);
  int invalid = !(((pin_ok >> pi) & 1u) && ((cell_ok >> sp) & 1ull));
  if (!invalid) {
    int swapped = s.board[sp];
    int pin_pos = s.pins[cp][pi];
    s.board[sp] = (int8_t)cp;
    s.board[pin_pos] = (int8_t)swapped;
    s.pins[cp][pi] = swap_pos;
    for (int k = 0; k < 4; ++k)
      if (s.pins[swapped][k] == swap_pos) s.pins[swapped][k] = pin_pos;
  }
  dog_finish_substep(g, s, cp, invalid, reward, done);
}

__host__ __device__ inline void dog_step_normal(const DogGeom& g, DogS& s, int pin_in, int move, int& reward, int& done) {
  int cp = dog_mover(g, s);
  int pin = d_gidx(pin_in, 4);
  int invalid = !dog_val_normal(g, s, cp, pin, move);
  if (!invalid) {
    const int mts = DG_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START);
    int pos = s.pins[cp][pin];
    int moved = pos + move, fitted = d_fmod(moved, g.bs);
    int x = moved - g.target[cp] - mts;
    int g0 = g.goal0[cp];
    int a = dog_in_goal(g, cp, pos) ? dog_path_clear(g, s.board, cp, pos - g0, moved - g0 + 1)
                                    : dog_path_clear(g, s.board, cp, -1, x);
    int gx = g0 + d_gidx(x - 1, 4);
    int A = (s.board[gx] != cp) && (DG_RULE(g, DOGSTEP_RULE_JUMP_IN_GOAL) || a);
    int new_pos;
    if (pos == -1) new_pos = g.start[cp];
    else if (dog_in_goal(g, cp, pos)) new_pos = moved;
    else if (4 >= x && x > 0 && A && pos <= g.target[cp]) new_pos = gx;
    else new_pos = fitted;
    dog_capture_and_place(g, s, cp, pin, new_pos);
  }
  dog_finish_substep(g, s, cp, invalid, reward, done);
}

__host__ __device__ inline void dog_step_neg(const DogGeom& g, DogS& s, int pin_in, int move, int& reward, int& done) {
  int cp = dog_mover(g, s);
  int pin = d_gidx(pin_in, 4);
  int invalid = !dog_val_neg(g, s, cp, pin, move);
  if (!invalid) dog_capture_and_place(g, s, cp, pin, d_fmod(s.pins[cp][pin] + move, g.bs));
  dog_finish_substep(g, s, cp, invalid, reward, done);
}

// get_path_matrix membership (utility_funcs.py:256-277)
DS_FN int d_giv(int si, int ei, int idx, int same_area) {
  if (si == -1 || ei == -1 || (same_area && si == ei)) return 0;
  if (si <= ei) return idx >= si && idx <= ei;
  return idx >= si || idx <= ei;
}

__host__ __device__ inline void dog_step_hot7(const DogGeom& g, DogS& s, const int dist[4], int& reward, int& done) {
  int cp = dog_mover(g, s);
  int invalid = !dog_val_7(g, s, cp, dist);
  if (!invalid) {
    const int mts = DG_RULE(g, DOGSTEP_RULE_MUST_TRAVERSE_START), jump = DG_RULE(g, DOGSTEP_RULE_JUMP_IN_GOAL);
    const int target = g.target[cp], g0 = g.goal0[cp];
    int cur[4], moved[4], nw[4];
    for (int i = 0; i < 4; ++i) { cur[i] = s.pins[cp][i]; moved[i] = cur[i] + dist[i]; }
    const uint32_t tmp_lane = dog_hot7_tmp_lane(g, s, cp, moved);
    int any_diff = 0;
    for (int i = 0; i < 4; ++i) {
      int fitted = d_fmod(moved[i], g.bs), x = moved[i] - target - mts;
      int a = dog_in_goal(g, cp, cur[i]) ? 1 : d_lane_clear(tmp_lane, -1, x);
      int A = jump || a;
      if (cur[i] == -1) nw[i] = -1;
      else if (dog_in_goal(g, cp, cur[i])) nw[i] = moved[i];
      else if (4 >= x && x > 0 && A && cur[i] <= target) nw[i] = g0 + d_gidx(x - 1, 4);
      else nw[i] = fitted;
      if (dog_in_goal(g, cp, cur[i]) != dog_in_goal(g, cp, nw[i])) any_diff = 1;
    }
    // path rows as 64-bit masks (get_path_matrix with traversal_over_start=True)
    uint64_t M[4];
    for (int i = 0; i < 4; ++i) {
      int A = dog_in_goal(g, cp, cur[i]), B = dog_in_goal(g, cp, nw[i]);
      uint64_t row = 0;
      for (int k = 0; k < g.total; ++k) {
        int v;
        if (A == B) v = (k < g.bs) ? d_giv(cur[i], nw[i], k, 1) : 0;
        else v = ((k < g.bs) ? d_giv(cur[i], target, k, 0) : 0) | d_giv(g0, nw[i], k, 0);
        row |= (uint64_t)v << k;
      }
      if (any_diff) row |= 1ull << g.start[cp];
      M[i] = row;
    }
    const uint64_t anyrow = M[0] | M[1] | M[2] | M[3];
    uint32_t hit = 0;  // bit p*4+i
    for (int p = 0; p < g.n; ++p)
      for (int i = 0; i < 4; ++i) hit |= (uint32_t)((anyrow >> d_gidx(s.pins[p][i], g.total)) & 1ull) << (p * 4 + i);
    for (int i = 0; i < 4; ++i) {  // check_moving_pins_hit
      uint64_t others = 0;
      for (int r = 0; r < 4; ++r)
        if (r != i) others |= M[r];
      int h = (int)((others >> d_gidx(cur[i], g.total)) & 1ull) & (int)((others >> d_gidx(nw[i], g.total)) & 1ull);
      hit = (hit & ~(1u << (cp * 4 + i))) | ((uint32_t)h << (cp * 4 + i));
    }
    for (int i = 0; i < 4; ++i) s.pins[cp][i] = nw[i];
    for (int p = 0; p < g.n; ++p)
      for (int i = 0; i < 4; ++i)
        if ((hit >> (p * 4 + i)) & 1u) s.pins[p][i] = -1;
    dog_set_pins_on_board(g, s.pins, s.board);
  }
  dog_finish_substep(g, s, cp, invalid, reward, done);
}

// first seat after `from` that still holds cards; all_empty = all(hand_cards == 0)
DS_FN int dog_next_with_cards(const DogGeom& g, const DogS& s, int from, int& all_empty, int& any_left) {
  int sums[4] = {0, 0, 0, 0};  // fixed trip counts and a select chain: the array stays in registers (no local memory)
  int nz = 0, pos = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int sum = 0;
    if (q < g.n)
      for (int k = 0; k < kNCard; ++k) sum += s.hands[q][k];
    sums[q] = sum;
    nz |= (sum != 0);
    pos |= (sum > 0);
  }
  int next = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < g.n) {
      const int cand = d_fmod(from + i + 1, g.n);
      const int sc = cand == 0 ? sums[0] : cand == 1 ? sums[1] : cand == 2 ? sums[2] : sums[3];
      if (next == -1 && sc > 0) next = cand;
    }
  }
  all_empty = !nz;
  any_left = pos;
  return next;
}

// map_action_to_move (dog.py:1134-1196)
DS_FN void dog_map_action_to_move(const DogGeom& g, int action, int mv[6]) {
  const int pxb = 4 * g.total;
  int is_joker = (action - g.half) < 0;
  int act = d_fmod(action, g.half);
  int is_swap = act < pxb;
  int d[4] = {0, 0, 0, 0};
  if (is_swap) {
    d[0] = d[1] = d[2] = d[3] = -1;
    int k = d_sidx(act / g.total, 4);
    if (k >= 0) d[k] = act % g.total;
  } else if (act < pxb + 120) {
    dog_dist_of(act - pxb, d);
  } else if (act < g.half - 4) {
    int na = act - (pxb + 120);
    int move = na % 12 + 1;
    move += (move >= 7);
    d[na / 12] = move;
  } else {
    int k = d_sidx(act - (g.half - 4), 4);
    if (k >= 0) d[k] = -4;
  }
  mv[0] = is_joker; mv[1] = is_swap;
  mv[2] = d[0]; mv[3] = d[1]; mv[4] = d[2]; mv[5] = d[3];
}

DS_FN int dog_map_action_to_card(const int mv[6]) {
  int sum = mv[2] + mv[3] + mv[4] + mv[5];
  if (mv[0] == 1) return 0;
  if (mv[1] == 1) return 1;
  if (sum == -4) return 4;
  return sum == 1 ? 11 : sum;
}

// distribute_cards (dog.py:201-298).  All 32 lanes must call it.
__device__ __noinline__ void dog_distribute_cards(const DogGeom& g, DogS& s, int lane) {
  const int n = g.n, quantity = s.hand_size;
  __syncwarp();
  if (lane == 0) {
    int deck_sum = 0;
    for (int k = 0; k < kNCard; ++k) deck_sum += s.deck[k];
    if (deck_sum < (int)(int8_t)(quantity * n)) {  // reset_deck (:183-186) with the joker enabled
      for (int k = 0; k < kNCard; ++k) s.deck[k] = 8;
    }
  }
  __syncwarp();
  const Key2 key{s.key[0], s.key[1]};
  // key, sub = split(env.key): one Threefry pass, even lanes take element 0 and odd lanes element 1
  const Key2 both = split_i(key, (uint32_t)(lane & 1));
  const Key2 knew{__shfl_sync(0xFFFFFFFFu, both.a, 0), __shfl_sync(0xFFFFFFFFu, both.b, 0)};
  const Key2 sub{__shfl_sync(0xFFFFFFFFu, both.a, 1), __shfl_sync(0xFFFFFFFFu, both.b, 1)};
  // slot j of the expanded pool holds card type c iff cum[c] <= j < cum[c+1]; dummies after the real cards.
  // argsort(uniform) is only needed for its first n*quantity entries: every lane keeps its four (uniform, index) keys and
  // card types in registers and the warp extracts the minimum `need` times with redux.sync (stable: the index is in the key)
  int deck_total = 0;
  for (int k = 0; k < kNCard; ++k) deck_total += s.deck[k];
  uint32_t sk[4];
  int ct[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = lane + 32 * r;
    sk[r] = 0xFFFFFFFFu;
    ct[r] = 0;
    if (j < 120 && j < deck_total) {
      sk[r] = ((bits_i(sub, (uint32_t)j) >> 9) << 7) | (uint32_t)j;
      int c = 0, acc = s.deck[0];
      while (j >= acc) { ++c; acc += s.deck[c]; }
      ct[r] = c;
    }
  }
  const int need = n * quantity;
  for (int rank = 0; rank < need; ++rank) {
    const uint32_t mine = min(min(sk[0], sk[1]), min(sk[2], sk[3]));
    const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, mine);
    if (m == 0xFFFFFFFFu) break;  // fewer real cards than seats x hand size
    if (mine == m) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (sk[r] == m) {
          if (rank < (int)(sizeof(s.items) / sizeof(s.items[0]))) s.items[rank] = (uint16_t)ct[r];
          sk[r] = 0xFFFFFFFFu;
        }
    }
  }
  __syncwarp();
  // deal: seat p receives the cards of ranks p * quantity .. p * quantity + quantity - 1.  Lanes own (seat, card type) pairs
  // and count their cards among the seat's ranks (the scalar loop over the 24 ranks on lane 0 was the longest serial
  // stretch of a turn); lanes 0..13 then take the dealt cards off the deck.
  {
    const int q6 = quantity < 6 ? quantity : 6;
    for (int pc = lane; pc < n * kNCard; pc += 32) {
      const int p = pc / kNCard, c = pc - p * kNCard;
      int cnt = 0;
      for (int slot = 0; slot < q6; ++slot) {
        const int idx = p * quantity + slot;
        cnt += (idx < deck_total && (int)s.items[idx] == c) ? 1 : 0;
      }
      if (cnt) s.hands[p][c] = (int8_t)(s.hands[p][c] + cnt);
    }
    if (lane < kNCard) {
      int cnt = 0;
      for (int p = 0; p < n; ++p)
        for (int slot = 0; slot < q6; ++slot) {
          const int idx = p * quantity + slot;
          cnt += (idx < deck_total && (int)s.items[idx] == lane) ? 1 : 0;
        }
      if (cnt) s.deck[lane] = (int8_t)(s.deck[lane] - cnt);
    }
  }
  __syncwarp();
  if (lane == 0) {
    int rs = (s.round_starter == -1) ? s.cur : d_fmod(s.round_starter + 1, n);
    s.cur = rs;
    s.round_starter = rs;
    for (int q = 0; q < 4; ++q) s.swap_choices[q] = -1;
    s.phase = (DG_RULE(g, DOGSTEP_RULE_TEAMS) && n == 4) ? 1 : 0;
    s.key[0] = knew.a;
    s.key[1] = knew.b;
    s.hand_size = (quantity == 2) ? 6 : quantity - 1;
  }
  __syncwarp();
}

// distribute_cards in three parts for the persistent play kernel, which shares the 120 Threefry draws of a deal among the
// warps of its CTA: begin (owner: deck reset, key split), draw (ANY warp: 32 pool slots per call), finish (owner: selection
// of the first n * hand_size cards, dealing, round bookkeeping).  Same arithmetic as dog_distribute_cards above.  The
// slot keys / card types travel through the record's item list (unused while a game deals): bytes 64.. of s.items.
__device__ __forceinline__ uint32_t* dog_deal_keys(DogS& s) { return reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s.items) + 64); }
__device__ __forceinline__ uint8_t* dog_deal_types(DogS& s) { return reinterpret_cast<unsigned char*>(s.items) + 64 + 512; }

__device__ __forceinline__ void dog_deal_begin(const DogGeom& g, DogS& s, int lane) {
  const int n = g.n, quantity = s.hand_size;
  __syncwarp();
  if (lane == 0) {
    int deck_sum = 0;
    for (int k = 0; k < kNCard; ++k) deck_sum += s.deck[k];
    if (deck_sum < (int)(int8_t)(quantity * n)) {  // reset_deck (:183-186) with the joker enabled
      for (int k = 0; k < kNCard; ++k) s.deck[k] = 8;
    }
  }
  __syncwarp();
  const Key2 key{s.key[0], s.key[1]};
  const Key2 both = split_i(key, (uint32_t)(lane & 1));  // key, sub = split(env.key)
  int deck_total = 0;
  for (int k = 0; k < kNCard; ++k) deck_total += s.deck[k];
  if (lane < 2) {
    s.scratch[1 + 2 * lane] = (int)both.a;  // [1],[2] = new key, [3],[4] = sub
    s.scratch[2 + 2 * lane] = (int)both.b;
  }
  if (lane == 0) s.scratch[5] = deck_total;
  __syncwarp();
}

__device__ __forceinline__ void dog_deal_draw(DogS& s, int r, int lane) {
  const int j = lane + 32 * r, deck_total = s.scratch[5];
  const Key2 sub{(uint32_t)s.scratch[3], (uint32_t)s.scratch[4]};
  uint32_t k = 0xFFFFFFFFu;
  int c = 0;
  if (j < 120 && j < deck_total) {
    k = ((bits_i(sub, (uint32_t)j) >> 9) << 7) | (uint32_t)j;
    int acc = s.deck[0];
    while (j >= acc) { ++c; acc += s.deck[c]; }
  }
  dog_deal_keys(s)[j] = k;
  dog_deal_types(s)[j] = (uint8_t)c;
}

__device__ __forceinline__ void dog_deal_finish(const DogGeom& g, DogS& s, int lane) {
  const int n = g.n, quantity = s.hand_size, deck_total = s.scratch[5];
  __syncwarp();
  uint32_t sk[4];
  int ct[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    sk[r] = dog_deal_keys(s)[lane + 32 * r];
    ct[r] = dog_deal_types(s)[lane + 32 * r];
  }
  const int need = n * quantity;
  for (int rank = 0; rank < need; ++rank) {
    const uint32_t mine = min(min(sk[0], sk[1]), min(sk[2], sk[3]));
    const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, mine);
    if (m == 0xFFFFFFFFu) break;  // fewer real cards than seats x hand size
    if (mine == m) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (sk[r] == m) {
          if (rank < 32) s.items[rank] = (uint16_t)ct[r];
          sk[r] = 0xFFFFFFFFu;
        }
    }
  }
  __syncwarp();
  {
    const int q6 = quantity < 6 ? quantity : 6;
    for (int pc = lane; pc < n * kNCard; pc += 32) {
      const int p = pc / kNCard, c = pc - p * kNCard;
      int cnt = 0;
      for (int slot = 0; slot < q6; ++slot) {
        const int idx = p * quantity + slot;
        cnt += (idx < deck_total && (int)s.items[idx] == c) ? 1 : 0;
      }
      if (cnt) s.hands[p][c] = (int8_t)(s.hands[p][c] + cnt);
    }
    if (lane < kNCard) {
      int cnt = 0;
      for (int p = 0; p < n; ++p)
        for (int slot = 0; slot < q6; ++slot) {
          const int idx = p * quantity + slot;
          cnt += (idx < deck_total && (int)s.items[idx] == lane) ? 1 : 0;
        }
      if (cnt) s.deck[lane] = (int8_t)(s.deck[lane] - cnt);
    }
  }
  __syncwarp();
  if (lane == 0) {
    int rs = (s.round_starter == -1) ? s.cur : d_fmod(s.round_starter + 1, n);
    s.cur = rs;
    s.round_starter = rs;
    for (int q = 0; q < 4; ++q) s.swap_choices[q] = -1;
    s.phase = (DG_RULE(g, DOGSTEP_RULE_TEAMS) && n == 4) ? 1 : 0;
    s.key[0] = (uint32_t)s.scratch[1];
    s.key[1] = (uint32_t)s.scratch[2];
    s.hand_size = (quantity == 2) ? 6 : quantity - 1;
  }
  __syncwarp();
}

// env_step_swap_phase (dog.py:1078-1114) on the staged record: scalar, no board access
DS_FN void dog_swap_phase(const DogGeom& g, DogS& s, int action) {
  int card_idx = action - g.play_actions;
  int cur = s.cur;
  int cs = d_sidx(card_idx, kNCard), row = d_sidx(cur, g.n);
  if (cs >= 0 && row >= 0) s.hands[row][cs] = (int8_t)(s.hands[row][cs] - 1);
  int sc = d_sidx(cur, 4);
  if (sc >= 0) s.swap_choices[sc] = (int8_t)card_idx;
  int next = d_fmod(cur + 1, g.n);
  if (next == s.round_starter) {
    const int partners[4] = {2, 3, 0, 1};
    for (int q = 0; q < g.n; ++q) {  // execute_team_swap (:1065-1075)
      int rc = s.swap_choices[partners[q]];
      if (rc >= 0 && rc < kNCard) s.hands[q][rc] = (int8_t)(s.hands[q][rc] + 1);
    }
    s.phase = 0;
    s.cur = s.round_starter;
    for (int q = 0; q < 4; ++q) s.swap_choices[q] = -1;
  } else {
    s.cur = next;
  }
  s.reward = 0;
}

// env_step (dog.py:1117-1131).  All lanes call; lane 0 applies the move, all lanes deal if needed.
__device__ __noinline__ void dog_env_step(const DogGeom& g, DogS& s, int lane, int action, int& reward_out, int& done_out) {
  __syncwarp();
  if (lane == 0) {
    int deal = 0, reward = 0, done = s.done;
    if (s.phase == 1) {
      dog_swap_phase(g, s, action);
      reward = 0;
      done = s.done;
    } else {  // env_step_play_phase (:987-1062)
      int pid = s.cur;
      int cp = dog_mover(g, s);
      int mv[6];
      dog_map_action_to_move(g, action, mv);
      int card = dog_map_action_to_card(mv);
      int valid_card = s.hands[cp][d_gidx(card, kNCard)] > 0;
      if (valid_card) {
        int* d = mv + 2;
        if (mv[1] == 1) {
          int pi = (d[0] >= 0) ? 0 : (d[1] >= 0) ? 1 : (d[2] >= 0) ? 2 : (d[3] >= 0) ? 3 : 0;
          dog_step_swap(g, s, pi, d[pi], reward, done);
        } else if (d[0] + d[1] + d[2] + d[3] == 7) {
          dog_step_hot7(g, s, d, reward, done);
        } else {
          int pi = (d[0] != 0) ? 0 : (d[1] != 0) ? 1 : (d[2] != 0) ? 2 : (d[3] != 0) ? 3 : 0;
          if (d[pi] < 0) dog_step_neg(g, s, pi, d[pi], reward, done);
          else dog_step_normal(g, s, pi, d[pi], reward, done);
        }
      } else {
        reward = -1;
        done = s.done;
      }
      int cs = d_sidx(card, kNCard);
      if (cs >= 0) s.hands[cp][cs] = (int8_t)(s.hands[cp][cs] + (reward == -1 ? 0 : -1));
      int all_empty, any_left;
      int next = dog_next_with_cards(g, s, pid, all_empty, any_left);
      s.cur = done ? cp : next;  // the PROXIED id is kept when the game ends (:1048)
      s.reward = reward;
      s.done = done;
      deal = (all_empty || next == -1) && !done;
    }
    s.scratch[0] = deal;
    s.scratch[1] = reward;
    s.scratch[2] = done;
  }
  __syncwarp();
  if (s.scratch[0] && !s.scratch[6]) dog_distribute_cards(g, s, lane);  // scratch[6]: the caller deals (k_dog_play_random)
  reward_out = s.scratch[1];
  done_out = s.scratch[2];
  __syncwarp();
}

// no_step (dog.py:713-752)
__device__ __noinline__ void dog_no_step(const DogGeom& g, DogS& s, int lane) {
  __syncwarp();
  if (lane == 0) {
    int row = d_sidx(s.cur, g.n);
    if (row >= 0)
      for (int k = 0; k < kNCard; ++k) s.hands[row][k] = 0;
    int all_empty, any_left;
    int next = dog_next_with_cards(g, s, s.cur, all_empty, any_left);
    int cont = any_left && next != -1;
    if (cont) s.cur = next;
    s.scratch[0] = !cont;
  }
  __syncwarp();
  if (s.scratch[0] && !s.scratch[6]) dog_distribute_cards(g, s, lane);  // scratch[6]: the caller deals (k_dog_play_random)
  __syncwarp();
}

}  // namespace dogstep
