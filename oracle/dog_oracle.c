/* dog_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Scalar restatement of the reference's DOG environment, /root/reference/DOG/dog.py, plus the
 * hot-7 helpers of /root/reference/utils/utility_funcs.py (:186-234, :237-303, :310-319).
 * Every function cites the reference lines it follows; JAX gather (wrap-once-then-clamp) and
 * scatter (wrap-once-then-drop) index rules are reproduced literally (gidx / sidx below).
 *
 * PARITY STATUS: pinned by the reference's own 112 DOG pytest cases (DOG/test.py, replayed through
 * env_step by tests/test_oracle_dog.py; the one case that contradicts the reference code is xfail
 * with the code's answer) and by trajectories the reference itself produced on oracle/jaxshim
 * (tests/golden/dog_reference_trajectories.npz).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "../include/dogstep_rules.h"
#include "jaxrand_oracle.h"

#define NCARD 14
#define PLAY_ACTIONS 792 /* get_play_action_size: 2*(4*(12+1+56)+120) for total_board_size 56 (:58-59) */

typedef struct {
  int n, d, bs, total;
  int start[4], target[4], goal[4][4];
  uint32_t rules;
  int play_actions, half; /* 2*(4*(13+total)+120) and its half */
} dog_cfg;

typedef struct {
  int8_t board[64];
  int8_t cur;
  int32_t pins[4][4];
  int8_t reward;
  uint8_t done;
  int8_t deck[NCARD];
  int8_t hands[4][NCARD];
  int8_t swap_choices[4];
  int8_t round_starter, phase, hand_size;
  uint32_t key[2];
} dog_state;

#define RULE(c, bit) (((c)->rules & (bit)) != 0)

static int fdiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
static int fmod_(int a, int b) { int r = a % b; if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }
static int gidx(int i, int size) { if (i < 0) i += size; if (i < 0) i = 0; if (i > size - 1) i = size - 1; return i; }
/* scatter index: negative wraps once; returns -1 when the update is dropped */
static int sidx(int i, int size) { if (i < 0) i += size; return (i < 0 || i >= size) ? -1 : i; }

/* 120 splits of 7 over 4 pins, lexicographic in (a,b,c) — utility_funcs.py:4-21, dog.py:11 */
static int8_t DISTS[120][4];
static int dists_ready = 0;
static void init_dists(void) {
  if (dists_ready) return;
  int k = 0;
  for (int a = 0; a <= 7; ++a)
    for (int b = 0; b <= 7; ++b)
      for (int c = 0; c <= 7; ++c) {
        int dd = 7 - a - b - c;
        if (dd >= 0) { DISTS[k][0] = (int8_t)a; DISTS[k][1] = (int8_t)b; DISTS[k][2] = (int8_t)c; DISTS[k][3] = (int8_t)dd; ++k; }
      }
  dists_ready = 1;
}

/* geometry — dog.py:107-122 */
static int make_cfg(dog_cfg *c, int num_players, int layout_mask, int distance, uint32_t rules) {
  if (num_players < 1 || num_players > 4 || distance < 1) return -1;
  init_dists();
  c->n = num_players; c->d = distance; c->bs = 4 * distance; c->total = c->bs + 16;
  if (num_players != 4) rules &= ~DOGSTEP_RULE_TEAMS; /* :111 */
  c->rules = rules;
  c->play_actions = 2 * (4 * (12 + 1 + c->total) + 120);
  c->half = c->play_actions / 2;
  int cnt = 0;
  for (int i = 0; i < 4; ++i) cnt += (layout_mask >> i) & 1;
  if (cnt != num_players || (layout_mask == 0xF && num_players < 4)) layout_mask = (1 << num_players) - 1;
  int r = 0;
  for (int i = 0; i < 4; ++i) {
    if (!((layout_mask >> i) & 1)) continue;
    c->start[r] = i * distance;
    c->target[r] = fmod_(c->start[r] - 1, c->bs);
    for (int k = 0; k < 4; ++k) c->goal[r][k] = c->bs + 4 * i + k;
    ++r;
  }
  for (; r < 4; ++r) { c->start[r] = c->target[r] = 0; memset(c->goal[r], 0, sizeof c->goal[r]); }
  return 0;
}

/* set_pins_on_board — dog.py:346-358 */
static void set_pins_on_board(const dog_cfg *c, int32_t pins[4][4], int8_t *board) {
  for (int k = 0; k < c->total; ++k) board[k] = -1;
  for (int p = 0; p < c->n; ++p)
    for (int i = 0; i < 4; ++i) {
      int pos = pins[p][i];
      if (pos >= 0 && pos < c->total) board[pos] = (int8_t)p;
    }
}

/* is_player_done — dog.py:300-315 */
static int is_player_done(const dog_cfg *c, const int8_t *board, int player) {
  if (player >= c->n) return 0;
  int p = gidx(player, c->n);
  for (int k = 0; k < 4; ++k) if (board[gidx(c->goal[p][k], c->total)] < 0) return 0;
  return 1;
}

/* get_winner — dog.py:317-344 */
static void get_winner(const dog_cfg *c, const int8_t *board, int w[4]) {
  int pd[4];
  for (int p = 0; p < 4; ++p) pd[p] = is_player_done(c, board, p);
  if (RULE(c, DOGSTEP_RULE_TEAMS)) {
    int t0 = pd[0] && pd[2], t1 = pd[1] && pd[3];
    if ((t0 && t1) || !(t0 || t1)) { w[0] = w[1] = w[2] = w[3] = 0; }
    else if (t0) { w[0] = 1; w[1] = 0; w[2] = 1; w[3] = 0; }
    else { w[0] = 0; w[1] = 1; w[2] = 0; w[3] = 1; }
  } else for (int p = 0; p < 4; ++p) w[p] = pd[p];
}

static int mover_of(const dog_cfg *c, const dog_state *s) {
  int pid = s->cur;
  int cp = (RULE(c, DOGSTEP_RULE_TEAMS) && is_player_done(c, s->board, pid)) ? fmod_(pid + 2, 4) : pid;
  return gidx(cp, c->n);
}

static int in_goal_of(const dog_cfg *c, int cp, int pos) {
  for (int k = 0; k < 4; ++k) if (pos == c->goal[cp][k]) return 1;
  return 0;
}

/* check_goal_path_for_pin — utility_funcs.py:165-184 */
static int goal_path_clear(const dog_cfg *c, const int8_t *board, int cp, int s, int e) {
  for (int k = 0; k < 4; ++k)
    if (s < k && k < e && board[gidx(c->goal[cp][k], c->total)] == cp) return 0;
  return 1;
}

static int pins_on_start(const dog_cfg *c, const int8_t *board, int q) {
  q = gidx(q, c->n);
  return board[gidx(c->start[q], c->total)] == q;
}

/* val_swap — dog.py:361-390: result[i][cell] = pin_ok[i] & cell_ok[cell] */
static void val_swap(const dog_cfg *c, const dog_state *s, uint8_t pin_ok[4], uint8_t cell_ok[64]) {
  int cp = mover_of(c, s);
  int sb = RULE(c, DOGSTEP_RULE_START_BLOCKING);
  for (int k = 0; k < c->total; ++k) cell_ok[k] = (s->board[k] != -1 && s->board[k] != cp);
  for (int q = 0; q < c->n; ++q) { /* :380 columns at the start cells */
    int cell = c->start[q];
    cell_ok[cell] = (uint8_t)(!((s->board[cell] == q) && sb) && (s->board[cell] != -1));
  }
  for (int i = 0; i < 4; ++i) { /* :381 own pins (a home pin, -1, wraps to the last column) */
    int col = sidx(s->pins[cp][i], c->total);
    if (col >= 0) cell_ok[col] = 0;
  }
  for (int q = 0; q < c->n; ++q) for (int k = 0; k < 4; ++k) cell_ok[c->goal[q][k]] = 0; /* :382 */
  for (int i = 0; i < 4; ++i) { /* :384-389 */
    int pos = s->pins[cp][i];
    int bad = (pos == -1) || in_goal_of(c, cp, pos) || (sb && pos == c->start[cp]);
    pin_ok[i] = (uint8_t)!bad;
  }
}

/* val_action_normal_move — dog.py:483-566 */
static void val_normal(const dog_cfg *c, const dog_state *s, int move, uint8_t out[4]) {
  int cp = mover_of(c, s);
  const int8_t *board = s->board;
  int mts = RULE(c, DOGSTEP_RULE_MUST_TRAVERSE_START), sb = RULE(c, DOGSTEP_RULE_START_BLOCKING);
  int circ = RULE(c, DOGSTEP_RULE_CIRCULAR_BOARD), jump = RULE(c, DOGSTEP_RULE_JUMP_IN_GOAL);
  int target = c->target[cp], g0 = c->goal[cp][0];
  for (int i = 0; i < 4; ++i) {
    int pos = s->pins[cp][i];
    int moved = pos + move, fitted = fmod_(moved, c->bs);
    int x = moved - target - mts;
    int result = (board[gidx(fitted, c->total)] != cp) || RULE(c, DOGSTEP_RULE_FRIENDLY_FIRE);
    int nsb = fmod_(fdiv(pos, c->d) + 1, c->n), nsa = fdiv(fitted, c->d);
    int trav = c->start[gidx(nsb, c->n)] == c->start[gidx(nsa, c->n)];
    int blocked = pins_on_start(c, board, nsa);
    if (sb && trav) result = (!blocked || pos == c->start[cp]) && result;
    if (mts && sb && trav && blocked) x = 0;
    if (!circ && pos <= target && (x > 4 || (x == 0 && mts))) result = 0;
    int A = circ && result;
    int B = board[gidx(c->goal[cp][gidx(x - 1, 4)], c->total)] != cp;
    int C = jump || goal_path_clear(c, board, cp, -1, x);
    if (4 >= x && x > 0 && pos <= target) result = A || (B && C);
    int D = jump || goal_path_clear(c, board, cp, pos - g0, moved - g0 + 1);
    if (in_goal_of(c, cp, pos)) result = (moved <= c->goal[cp][3]) && (board[gidx(moved, c->total)] != cp) && D;
    if (pos == -1) result = (move == 1 || move == 11 || move == 13) && !pins_on_start(c, board, cp);
    out[i] = (uint8_t)(result && move > 0);
  }
}

/* val_neg_move — dog.py:568-614 */
static void val_neg(const dog_cfg *c, const dog_state *s, int move, uint8_t out[4]) {
  int cp = mover_of(c, s);
  const int8_t *board = s->board;
  for (int i = 0; i < 4; ++i) {
    int pos = s->pins[cp][i];
    int moved = pos + move, fitted = fmod_(moved, c->bs);
    int result = (board[gidx(fitted, c->total)] != cp) || RULE(c, DOGSTEP_RULE_FRIENDLY_FIRE);
    int nsb = fdiv(pos, c->d), nsa = fmod_(fdiv(fitted, c->d) + 1, c->n);
    int cond = c->start[gidx(nsb, c->n)] == c->start[gidx(nsa, c->n)];
    if (RULE(c, DOGSTEP_RULE_START_BLOCKING) && cond)
      result = (!pins_on_start(c, board, nsa) || pos == c->start[cp]) && result;
    result = result && (RULE(c, DOGSTEP_RULE_CIRCULAR_BOARD) || moved >= c->start[cp]);
    if (pos == -1 || in_goal_of(c, cp, pos)) result = 0;
    out[i] = (uint8_t)result;
  }
}

/* check_relative_order_preserved — utility_funcs.py:186-234 */
static int sgn(int v) { return (v > 0) - (v < 0); }
static void rel_order(const int old[4], const int nw[4], int bs, int out[4]) {
  for (int i = 0; i < 4; ++i) {
    int ok = 1;
    for (int j = 0; j < 4; ++j)
      if (old[i] >= bs && old[j] >= bs && sgn(old[i] - old[j]) != sgn(nw[i] - nw[j])) ok = 0;
    out[i] = (old[i] < bs) || ok;
  }
}

/* tmp_board of val_action_7 / step_hot_7 (dog.py:454-455, 934-935) */
static void hot7_tmp_board(const dog_cfg *c, const dog_state *s, int cp, const int moved[4], int8_t *tmp) {
  int32_t tp[4][4];
  memcpy(tp, s->pins, sizeof tp);
  for (int i = 0; i < 4; ++i) if (in_goal_of(c, cp, s->pins[cp][i])) tp[cp][i] = moved[i];
  set_pins_on_board(c, tp, tmp);
}

/* val_action_7 — dog.py:393-481 → scalar */
static int val_7(const dog_cfg *c, const dog_state *s, const int dist[4]) {
  int cp = mover_of(c, s);
  const int8_t *board = s->board;
  int mts = RULE(c, DOGSTEP_RULE_MUST_TRAVERSE_START), sb = RULE(c, DOGSTEP_RULE_START_BLOCKING);
  int circ = RULE(c, DOGSTEP_RULE_CIRCULAR_BOARD), jump = RULE(c, DOGSTEP_RULE_JUMP_IN_GOAL);
  int target = c->target[cp];
  int cur[4], moved[4], fitted[4], x[4];
  int pos_[4] = {0, 0, 0, 0};
  for (int q = 0; q < c->n; ++q) pos_[q] = pins_on_start(c, board, q);
  int own_start_stays = 0;
  for (int i = 0; i < 4; ++i) {
    cur[i] = s->pins[cp][i];
    moved[i] = cur[i] + dist[i];
    fitted[i] = fmod_(moved[i], c->bs);
    x[i] = moved[i] - target - mts;
    if (cur[i] == c->start[cp] && moved[i] == c->start[cp]) own_start_stays = 1;
  }
  pos_[cp] = own_start_stays; /* :426 */
  int8_t tmp[64];
  hot7_tmp_board(c, s, cp, moved, tmp);
  int order[4];
  rel_order(cur, moved, c->bs, order);
  int all = 1;
  for (int i = 0; i < 4; ++i) {
    int result = circ ? 1 : !((cur[i] <= target) && ((moved[i] > target + 4) || (x[i] == 0 && mts)));
    int nsb = fmod_(fdiv(cur[i], c->d) + 1, c->n), nsa = fdiv(fitted[i], c->d);
    int trav = c->start[gidx(nsb, c->n)] == c->start[gidx(nsa, c->n)];
    int blocked = pos_[gidx(nsa, c->n)];
    if (sb && trav) result = !blocked && result;
    int xi = x[i];
    if (mts && sb && trav && blocked) xi = 0;
    int A = circ && result;
    int C = jump || goal_path_clear(c, tmp, cp, -1, xi);
    if (4 >= xi && xi > 0 && cur[i] <= target) result = A || C;
    int D = jump || order[i];
    if (in_goal_of(c, cp, cur[i])) result = (moved[i] <= c->goal[cp][3]) && D;
    int board_mover = (cur[i] == -1) ? (moved[i] == -1) : 1;
    all = all && result && board_mover;
  }
  return all;
}

/* get_path_matrix — utility_funcs.py:237-303 (traversal_over_start=True) */
static int giv(int si, int ei, int idx, int same_area) {
  if (si == -1 || ei == -1 || (same_area && si == ei)) return 0;
  if (si <= ei) return idx >= si && idx <= ei;
  return idx >= si || idx <= ei;
}
static void path_matrix(const dog_cfg *c, int cp, const int cur[4], const int nw[4], uint8_t M[4][64]) {
  int A[4], B[4], any_diff = 0;
  int g0 = c->goal[cp][0], target = c->target[cp];
  memset(M, 0, 4 * 64);
  for (int i = 0; i < 4; ++i) {
    A[i] = in_goal_of(c, cp, cur[i]);
    B[i] = in_goal_of(c, cp, nw[i]);
    if (A[i] != B[i]) any_diff = 1;
    for (int k = 0; k < c->total; ++k) {
      int v;
      if (A[i] == B[i]) v = (k < c->bs) ? giv(cur[i], nw[i], k, 1) : 0;
      else v = ((k < c->bs) ? giv(cur[i], target, k, 0) : 0) | giv(g0, nw[i], k, 0);
      M[i][k] = (uint8_t)v;
    }
  }
  if (any_diff) for (int i = 0; i < 4; ++i) M[i][c->start[cp]] = 1;
}

/* step results: board/pins written back into s only when the sub-step is valid */
typedef struct { int reward; int done; } step_out;

static step_out finish_substep(const dog_cfg *c, dog_state *s, int cp, int invalid) {
  int w[4];
  get_winner(c, s->board, w);
  step_out o;
  o.done = s->done || w[0] || w[1] || w[2] || w[3];
  o.reward = s->done ? 0 : (invalid ? -1 : w[gidx(cp, 4)]);
  return o;
}

/* step_swap — dog.py:755-787 */
static step_out step_swap(const dog_cfg *c, dog_state *s, int pin_idx, int swap_pos) {
  int cp = mover_of(c, s);
  uint8_t pin_ok[4], cell_ok[64];
  val_swap(c, s, pin_ok, cell_ok);
  int pi = gidx(pin_idx, 4), sp = gidx(swap_pos, c->total);
  int invalid = !(pin_ok[pi] && cell_ok[sp]);
  if (!invalid) {
    int swapped = s->board[sp];
    int pin_pos = s->pins[cp][pi];
    s->board[sp] = (int8_t)cp;
    s->board[pin_pos] = (int8_t)swapped;
    s->pins[cp][pi] = swap_pos;
    for (int k = 0; k < 4; ++k) if (s->pins[swapped][k] == swap_pos) s->pins[swapped][k] = pin_pos;
  }
  return finish_substep(c, s, cp, invalid);
}

static void capture_and_place(const dog_cfg *c, dog_state *s, int cp, int pin, int new_pos) {
  int pin_at_pos = s->board[gidx(new_pos, c->total)];
  if (pin_at_pos != -1 && (pin_at_pos != cp || RULE(c, DOGSTEP_RULE_FRIENDLY_FIRE))) {
    int q = gidx(pin_at_pos, c->n);
    for (int k = 0; k < 4; ++k) if (s->pins[q][k] == new_pos) s->pins[q][k] = -1;
  }
  s->pins[cp][pin] = new_pos;
  set_pins_on_board(c, s->pins, s->board);
}

/* step_normal_move — dog.py:790-858 */
static step_out step_normal(const dog_cfg *c, dog_state *s, int pin_in, int move) {
  int cp = mover_of(c, s);
  uint8_t v[4];
  val_normal(c, s, move, v);
  int pin = gidx(pin_in, 4);
  int invalid = !v[pin];
  int mts = RULE(c, DOGSTEP_RULE_MUST_TRAVERSE_START);
  int pos = s->pins[cp][pin];
  int moved = pos + move, fitted = fmod_(moved, c->bs);
  int x = moved - c->target[cp] - mts;
  int g0 = c->goal[cp][0];
  int a = in_goal_of(c, cp, pos) ? goal_path_clear(c, s->board, cp, pos - g0, moved - g0 + 1)
                                 : goal_path_clear(c, s->board, cp, -1, x);
  int gx = c->goal[cp][gidx(x - 1, 4)];
  int A = (s->board[gidx(gx, c->total)] != cp) && (RULE(c, DOGSTEP_RULE_JUMP_IN_GOAL) || a);
  int new_pos;
  if (pos == -1) new_pos = c->start[cp];
  else if (in_goal_of(c, cp, pos)) new_pos = moved;
  else if (4 >= x && x > 0 && A && pos <= c->target[cp]) new_pos = gx;
  else new_pos = fitted;
  if (!invalid) capture_and_place(c, s, cp, pin, new_pos);
  return finish_substep(c, s, cp, invalid);
}

/* step_neg_move — dog.py:861-910 */
static step_out step_neg(const dog_cfg *c, dog_state *s, int pin_in, int move) {
  int cp = mover_of(c, s);
  uint8_t v[4];
  val_neg(c, s, move, v);
  int pin = gidx(pin_in, 4);
  int invalid = !v[pin];
  int new_pos = fmod_(s->pins[cp][pin] + move, c->bs);
  if (!invalid) capture_and_place(c, s, cp, pin, new_pos);
  return finish_substep(c, s, cp, invalid);
}

/* step_hot_7 — dog.py:913-984 */
static step_out step_hot7(const dog_cfg *c, dog_state *s, const int dist[4]) {
  int cp = mover_of(c, s);
  int invalid = !val_7(c, s, dist);
  if (!invalid) {
    int mts = RULE(c, DOGSTEP_RULE_MUST_TRAVERSE_START), jump = RULE(c, DOGSTEP_RULE_JUMP_IN_GOAL);
    int target = c->target[cp];
    int cur[4], moved[4], nw[4];
    for (int i = 0; i < 4; ++i) { cur[i] = s->pins[cp][i]; moved[i] = cur[i] + dist[i]; }
    int8_t tmp[64];
    hot7_tmp_board(c, s, cp, moved, tmp);
    for (int i = 0; i < 4; ++i) {
      int fitted = fmod_(moved[i], c->bs), x = moved[i] - target - mts;
      int a = in_goal_of(c, cp, cur[i]) ? 1 : goal_path_clear(c, tmp, cp, -1, x);
      int A = jump || a;
      if (cur[i] == -1) nw[i] = -1;
      else if (in_goal_of(c, cp, cur[i])) nw[i] = moved[i];
      else if (4 >= x && x > 0 && A && cur[i] <= target) nw[i] = c->goal[cp][gidx(x - 1, 4)];
      else nw[i] = fitted;
    }
    uint8_t M[4][64];
    path_matrix(c, cp, cur, nw, M);
    uint8_t anyrow[64];
    for (int k = 0; k < c->total; ++k) anyrow[k] = M[0][k] | M[1][k] | M[2][k] | M[3][k];
    int hit[4][4];
    for (int p = 0; p < c->n; ++p)
      for (int i = 0; i < 4; ++i) hit[p][i] = anyrow[gidx(s->pins[p][i], c->total)]; /* :963 (-1 wraps to the last cell) */
    for (int i = 0; i < 4; ++i) { /* check_moving_pins_hit — utility_funcs.py:310-319 */
      int sh = 0, eh = 0;
      for (int r = 0; r < 4; ++r) {
        if (r == i) continue;
        sh |= M[r][gidx(cur[i], c->total)];
        eh |= M[r][gidx(nw[i], c->total)];
      }
      hit[cp][i] = sh && eh;
    }
    for (int i = 0; i < 4; ++i) s->pins[cp][i] = nw[i];
    for (int p = 0; p < c->n; ++p)
      for (int i = 0; i < 4; ++i) if (hit[p][i]) s->pins[p][i] = -1;
    set_pins_on_board(c, s->pins, s->board);
  }
  return finish_substep(c, s, cp, invalid);
}

/* reset_deck — dog.py:183-186 */
static void reset_deck(const dog_cfg *c, int8_t deck[NCARD]) {
  for (int k = 0; k < NCARD; ++k) deck[k] = 8;
  deck[0] = (int8_t)(6 + 2 * (RULE(c, DOGSTEP_RULE_DISABLE_JOKER) ? 0 : 1));
}

/* distribute_cards — dog.py:201-298 */
static void distribute_cards(const dog_cfg *c, dog_state *s) {
  const int n = c->n, quantity = s->hand_size;
  int deck_sum = 0;
  for (int k = 0; k < NCARD; ++k) deck_sum += s->deck[k];
  int8_t deck[NCARD];
  if (deck_sum < (int8_t)(quantity * n)) reset_deck(c, deck);
  else memcpy(deck, s->deck, NCARD);
  /* expanded pool of 120 slots: card types in order, then dummies (:230-242) */
  int pool[120], np_ = 0;
  for (int k = 0; k < NCARD; ++k) for (int r = 0; r < deck[k] && np_ < 120; ++r) pool[np_++] = k;
  while (np_ < 120) pool[np_++] = NCARD;
  uint32_t knew[2], sub[2];
  orc_split_i(s->key, 0, knew);
  orc_split_i(s->key, 1, sub);
  float pri[120];
  for (int j = 0; j < 120; ++j) pri[j] = (pool[j] == NCARD) ? 2.0f : orc_uniform_i(sub, (uint32_t)j, 0.0f, 1.0f);
  int order[120]; /* stable argsort (:253) */
  for (int j = 0; j < 120; ++j) {
    int k = j;
    while (k > 0 && pri[order[k - 1]] > pri[j]) { order[k] = order[k - 1]; --k; }
    order[k] = j;
  }
  int8_t add[4][NCARD];
  memset(add, 0, sizeof add);
  for (int p = 0; p < n; ++p)
    for (int slot = 0; slot < 6; ++slot) {
      if (slot >= quantity) continue;
      int card = pool[order[gidx(p * quantity + slot, 120)]];
      if (card < NCARD) add[p][card]++;
    }
  for (int k = 0; k < NCARD; ++k) {
    int tot = 0;
    for (int p = 0; p < n; ++p) { s->hands[p][k] = (int8_t)(s->hands[p][k] + add[p][k]); tot += add[p][k]; }
    s->deck[k] = (int8_t)(deck[k] - tot);
  }
  int rs = (s->round_starter == -1) ? s->cur : fmod_(s->round_starter + 1, n);
  s->cur = (int8_t)rs;
  s->round_starter = (int8_t)rs;
  for (int q = 0; q < 4; ++q) s->swap_choices[q] = -1;
  s->phase = (int8_t)((RULE(c, DOGSTEP_RULE_TEAMS) && n == 4) ? 1 : 0);
  s->key[0] = knew[0];
  s->key[1] = knew[1];
  s->hand_size = (int8_t)(quantity == 2 ? 6 : quantity - 1);
}

/* first seat after env.current_player that still holds cards (dog.py:726-731, 1041-1046) */
static int next_with_cards(const dog_cfg *c, const dog_state *s, int from, int *all_empty) {
  int next = -1, any = 0;
  for (int q = 0; q < c->n; ++q) {
    int sum = 0;
    for (int k = 0; k < NCARD; ++k) sum += s->hands[q][k];
    if (sum != 0) any = 1; /* jnp.all(hand_cards == 0) */
  }
  for (int i = 0; i < c->n; ++i) {
    int cand = fmod_(from + i + 1, c->n);
    int sum = 0;
    for (int k = 0; k < NCARD; ++k) sum += s->hands[cand][k];
    if (next == -1 && sum > 0) next = cand;
  }
  *all_empty = !any;
  return next;
}

/* map_action_to_move — dog.py:1134-1196 → [is_joker, is_swap, d0..d3] */
void orc_dog_map_action_to_move_one(const dog_cfg *c, int action, int out[6]) {
  int half = c->half, pxb = 4 * c->total;
  int is_joker = (action - half) < 0;
  int act = fmod_(action, half);
  int is_swap = act < pxb;
  int dist[4] = {0, 0, 0, 0};
  if (is_swap) {
    for (int i = 0; i < 4; ++i) dist[i] = -1;
    int k = sidx(act / c->total, 4);
    if (k >= 0) dist[k] = act % c->total;
  } else if (act < pxb + 120) {
    for (int i = 0; i < 4; ++i) dist[i] = DISTS[act - pxb][i];
  } else if (act < half - 4) {
    int na = act - (pxb + 120);
    int move = na % 12 + 1;
    move += (move >= 7);
    dist[na / 12] = move;
  } else {
    int k = sidx(act - (half - 4), 4);
    if (k >= 0) dist[k] = -4;
  }
  out[0] = is_joker; out[1] = is_swap;
  for (int i = 0; i < 4; ++i) out[2 + i] = dist[i];
}

/* map_action_to_card — dog.py:1241-1262 */
static int map_action_to_card(const int mv[6]) {
  int sum = mv[2] + mv[3] + mv[4] + mv[5];
  if (mv[0] == 1) return 0;
  if (mv[1] == 1) return 1;
  if (sum == -4) return 4;
  return sum == 1 ? 11 : sum;
}

static int argmax_first(const int v[4]) { for (int i = 0; i < 4; ++i) if (v[i]) return i; return 0; }

/* env_step_play_phase — dog.py:987-1062 */
static void step_play(const dog_cfg *c, dog_state *s, int action, int8_t *reward, uint8_t *done) {
  int pid = s->cur;
  int cp = mover_of(c, s);
  int mv[6];
  orc_dog_map_action_to_move_one(c, action, mv);
  int card = map_action_to_card(mv);
  int ci = gidx(card, NCARD);
  int valid_card = s->hands[cp][ci] > 0;
  step_out o;
  if (valid_card) {
    int *d = mv + 2;
    if (mv[1] == 1) {
      int ge[4] = {d[0] >= 0, d[1] >= 0, d[2] >= 0, d[3] >= 0};
      int pi = argmax_first(ge);
      o = step_swap(c, s, pi, d[pi]);
    } else if (d[0] + d[1] + d[2] + d[3] == 7) {
      o = step_hot7(c, s, d);
    } else {
      int nz[4] = {d[0] != 0, d[1] != 0, d[2] != 0, d[3] != 0};
      int pi = argmax_first(nz);
      o = (d[pi] < 0) ? step_neg(c, s, pi, d[pi]) : step_normal(c, s, pi, d[pi]);
    }
  } else {
    o.reward = -1;
    o.done = s->done;
  }
  int cs = sidx(card, NCARD);
  if (cs >= 0) s->hands[cp][cs] = (int8_t)(s->hands[cp][cs] + (o.reward == -1 ? 0 : -1)); /* :1039 */
  int all_empty;
  int next = next_with_cards(c, s, pid, &all_empty);
  s->cur = (int8_t)(o.done ? cp : next); /* :1048 — the PROXIED id is kept when the game ends */
  s->reward = (int8_t)o.reward;
  s->done = (uint8_t)o.done;
  if ((all_empty || next == -1) && !o.done) distribute_cards(c, s);
  *reward = (int8_t)o.reward;
  *done = (uint8_t)o.done;
}

/* env_step_swap_phase — dog.py:1078-1114 (+ execute_team_swap :1065-1075) */
static void step_swap_phase(const dog_cfg *c, dog_state *s, int card_idx, int8_t *reward, uint8_t *done) {
  int cur = s->cur;
  int cs = sidx(card_idx, NCARD), row = sidx(cur, c->n);
  if (cs >= 0 && row >= 0) s->hands[row][cs] = (int8_t)(s->hands[row][cs] - 1);
  int sc = sidx(cur, 4);
  if (sc >= 0) s->swap_choices[sc] = (int8_t)card_idx;
  int next = fmod_(cur + 1, c->n);
  int cycle = next == s->round_starter;
  if (cycle) {
    static const int partners[4] = {2, 3, 0, 1};
    for (int q = 0; q < c->n; ++q) {
      int rc = s->swap_choices[partners[q]];
      if (rc >= 0 && rc < NCARD) s->hands[q][rc] = (int8_t)(s->hands[q][rc] + 1);
    }
    s->phase = 0;
    s->cur = s->round_starter;
    for (int q = 0; q < 4; ++q) s->swap_choices[q] = -1;
  } else {
    s->cur = (int8_t)next;
  }
  s->reward = 0;
  *reward = 0;
  *done = s->done;
}

/* env_step — dog.py:1117-1131 */
void orc_dog_step_one(const dog_cfg *c, dog_state *s, int action, int8_t *reward, uint8_t *done) {
  if (s->phase == 1) step_swap_phase(c, s, action - c->play_actions, reward, done);
  else step_play(c, s, action, reward, done);
}

/* no_step — dog.py:713-752 */
void orc_dog_no_step_one(const dog_cfg *c, dog_state *s, int8_t *reward, uint8_t *done) {
  int row = sidx(s->cur, c->n);
  if (row >= 0) memset(s->hands[row], 0, NCARD);
  int all_empty;
  int next = next_with_cards(c, s, s->cur, &all_empty);
  int any_left = 0;
  for (int q = 0; q < c->n; ++q) { int sum = 0; for (int k = 0; k < NCARD; ++k) sum += s->hands[q][k]; if (sum > 0) any_left = 1; }
  if (any_left && next != -1) s->cur = (int8_t)next;
  else distribute_cards(c, s);
  *reward = 0;
  *done = s->done;
}

/* valid_step_actions + valid_actions — dog.py:618-711 → uint8[play_actions + 14] */
void orc_dog_valid_actions_one(const dog_cfg *c, const dog_state *s, uint8_t *mask) {
  const int half = c->half, pxb = 4 * c->total, nact = c->play_actions + NCARD;
  memset(mask, 0, (size_t)nact);
  if (s->phase != 0) { /* swap phase: the UN-proxied seat's cards (:701,710) */
    int row = gidx(s->cur, c->n);
    for (int k = 0; k < NCARD; ++k) mask[c->play_actions + k] = s->hands[row][k] > 0;
    return;
  }
  int cp = mover_of(c, s);
  const int8_t *hand = s->hands[cp];
  uint8_t *joker = mask, *all = mask + half;
  uint8_t pin_ok[4], cell_ok[64];
  val_swap(c, s, pin_ok, cell_ok);
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < c->total; ++k) {
      uint8_t v = pin_ok[i] && cell_ok[k];
      joker[i * c->total + k] = v;
      all[i * c->total + k] = (hand[1] > 0) ? v : 0;
    }
  for (int k = 0; k < 120; ++k) {
    int d[4] = {DISTS[k][0], DISTS[k][1], DISTS[k][2], DISTS[k][3]};
    uint8_t v = (uint8_t)val_7(c, s, d);
    joker[pxb + k] = v;
    all[pxb + k] = (hand[7] > 0) ? v : 0;
  }
  static const int moves[12] = {1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13};
  for (int k = 0; k < 12; ++k) {
    uint8_t v[4];
    val_normal(c, s, moves[k], v);
    int card = (k == 0) ? 11 : moves[k]; /* :660-670: move 1 needs card 11, the others their own card */
    for (int i = 0; i < 4; ++i) {
      joker[pxb + 120 + i * 12 + k] = v[i];
      all[pxb + 120 + i * 12 + k] = (hand[card] > 0) ? v[i] : 0;
    }
  }
  uint8_t vn[4];
  val_neg(c, s, -4, vn);
  for (int i = 0; i < 4; ++i) { joker[half - 4 + i] = vn[i]; all[half - 4 + i] = (hand[4] > 0) ? vn[i] : 0; }
  if (!(hand[0] > 0)) memset(joker, 0, (size_t)half);
}

/* env_reset — dog.py:83-181 */
void orc_dog_reset_one(const dog_cfg *c, dog_state *s, int32_t seed, int starting_player) {
  uint32_t k0[2], knew[2], sub[2];
  orc_prngkey(seed, k0);
  orc_split_i(k0, 0, knew);
  orc_split_i(k0, 1, sub);
  int sp = starting_player;
  if (sp < 0 || sp >= c->n) sp = orc_randint_i(sub, 0, 0, c->n);
  memset(s, 0, sizeof *s);
  for (int p = 0; p < 4; ++p) for (int i = 0; i < 4; ++i) s->pins[p][i] = -1;
  if (RULE(c, DOGSTEP_RULE_INITIAL_FREE_PIN)) for (int p = 0; p < c->n; ++p) s->pins[p][0] = c->start[p];
  set_pins_on_board(c, s->pins, s->board);
  s->cur = (int8_t)sp;
  for (int k = 0; k < NCARD; ++k) s->deck[k] = 8;
  s->deck[0] = (int8_t)(6 + 2 * RULE(c, DOGSTEP_RULE_DISABLE_JOKER));
  for (int q = 0; q < 4; ++q) s->swap_choices[q] = -1;
  s->round_starter = -1;
  s->phase = 0;
  s->hand_size = 6;
  s->key[0] = knew[0];
  s->key[1] = knew[1];
  distribute_cards(c, s);
}

/* ------------------------------------------------------------------------------------------
 * Batched SoA entry points (leaves of the vmapped pytree): board i8[n,T], current_player i8[n],
 * pins i32[n,P,4], reward i8[n], done u8[n], deck i8[n,14], hands i8[n,P,14], swap_choices i8[n,4],
 * round_starter i8[n], phase i8[n], key u32[n,2], hand_size i8[n].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int8_t *board, *cur; int32_t *pins; int8_t *reward; uint8_t *done; int8_t *deck, *hands, *swap_choices,
      *round_starter, *phase; uint32_t *key; int8_t *hand_size;
} dog_soa;

static void soa_load(const dog_cfg *c, const dog_soa *a, int64_t g, dog_state *s) {
  memset(s, 0, sizeof *s);
  memcpy(s->board, a->board + g * c->total, (size_t)c->total);
  s->cur = a->cur[g];
  for (int p = 0; p < 4; ++p) for (int i = 0; i < 4; ++i) s->pins[p][i] = (p < c->n) ? a->pins[(g * c->n + p) * 4 + i] : -1;
  s->reward = a->reward[g];
  s->done = a->done[g];
  memcpy(s->deck, a->deck + g * NCARD, NCARD);
  for (int p = 0; p < c->n; ++p) memcpy(s->hands[p], a->hands + (g * c->n + p) * NCARD, NCARD);
  memcpy(s->swap_choices, a->swap_choices + g * 4, 4);
  s->round_starter = a->round_starter[g];
  s->phase = a->phase[g];
  s->key[0] = a->key[2 * g];
  s->key[1] = a->key[2 * g + 1];
  s->hand_size = a->hand_size[g];
}

static void soa_store(const dog_cfg *c, const dog_soa *a, int64_t g, const dog_state *s) {
  memcpy(a->board + g * c->total, s->board, (size_t)c->total);
  a->cur[g] = s->cur;
  for (int p = 0; p < c->n; ++p) for (int i = 0; i < 4; ++i) a->pins[(g * c->n + p) * 4 + i] = s->pins[p][i];
  a->reward[g] = s->reward;
  a->done[g] = s->done;
  memcpy(a->deck + g * NCARD, s->deck, NCARD);
  for (int p = 0; p < c->n; ++p) memcpy(a->hands + (g * c->n + p) * NCARD, s->hands[p], NCARD);
  memcpy(a->swap_choices + g * 4, s->swap_choices, 4);
  a->round_starter[g] = s->round_starter;
  a->phase[g] = s->phase;
  a->key[2 * g] = s->key[0];
  a->key[2 * g + 1] = s->key[1];
  a->hand_size[g] = s->hand_size;
}

#define CFG_ARGS int num_players, int layout_mask, int distance, uint32_t rules
#define MAKE_CFG dog_cfg cfg; if (make_cfg(&cfg, num_players, layout_mask, distance, rules)) return -1

int orc_dog_num_actions(CFG_ARGS) { MAKE_CFG; return cfg.play_actions + NCARD; }

int orc_dog_reset(CFG_ARGS, int64_t n, const int32_t *seeds, int starting_player, const dog_soa *a) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) { dog_state s; orc_dog_reset_one(&cfg, &s, seeds[g], starting_player); soa_store(&cfg, a, g, &s); }
  return 0;
}

int orc_dog_valid_actions(CFG_ARGS, int64_t n, const dog_soa *a, uint8_t *mask) {
  MAKE_CFG;
  const int na = cfg.play_actions + NCARD;
  for (int64_t g = 0; g < n; ++g) { dog_state s; soa_load(&cfg, a, g, &s); orc_dog_valid_actions_one(&cfg, &s, mask + g * na); }
  return 0;
}

int orc_dog_step(CFG_ARGS, int64_t n, const dog_soa *a, const int32_t *action, int8_t *reward, uint8_t *done) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) {
    dog_state s;
    soa_load(&cfg, a, g, &s);
    orc_dog_step_one(&cfg, &s, action[g], reward + g, done + g);
    soa_store(&cfg, a, g, &s);
  }
  return 0;
}

int orc_dog_no_step(CFG_ARGS, int64_t n, const dog_soa *a, int8_t *reward, uint8_t *done) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) {
    dog_state s;
    soa_load(&cfg, a, g, &s);
    orc_dog_no_step_one(&cfg, &s, reward + g, done + g);
    soa_store(&cfg, a, g, &s);
  }
  return 0;
}

int orc_dog_distribute_cards(CFG_ARGS, int64_t n, const dog_soa *a) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) { dog_state s; soa_load(&cfg, a, g, &s); distribute_cards(&cfg, &s); soa_store(&cfg, a, g, &s); }
  return 0;
}

int orc_dog_map_action_to_move(CFG_ARGS, int64_t n, const int32_t *action, int32_t *out /*[n,6]*/) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) { int mv[6]; orc_dog_map_action_to_move_one(&cfg, action[g], mv); for (int k = 0; k < 6; ++k) out[g * 6 + k] = mv[k]; }
  return 0;
}

/* Random-legal-policy lockstep driver for DOG (config 4).  The reference's DOG eval loop
 * (MuZero_DOG/evaluate_agent.py:333-527) is an unfinished copy of the MADN one; the semantics used
 * here are the MADN driver's (MuZero_det_MADN/evaluate_agent.py:733-930, do_random) with DOG's
 * functions: key_j = split(rng, N+1)[j+1]; mask = valid_actions(env) (806);
 * any(mask) ? env_step(env, categorical(key_j, where(mask, 0, -1e9))) : no_step(env). */
typedef struct {
  dog_cfg cfg; const dog_soa *a; int64_t n, game_offset; uint32_t rng0[2]; int max_steps, float_gumbel;
  int32_t *game_len; int nthreads, tid; int64_t total; int iters;
} dog_job;

static int categorical_masked(const uint32_t key[2], const uint8_t *mask, int na, int float_gumbel) {
  int best = -1;
  if (float_gumbel) {
    float bv = 0.f;
    for (int a = 0; a < na; ++a) {
      float v = (mask[a] ? 0.0f : -1e9f) + orc_gumbel_i(key, (uint32_t)a);
      if (best < 0 || v > bv) { best = a; bv = v; }
    }
  } else {
    uint32_t bm = 0;
    for (int a = 0; a < na; ++a) {
      if (!mask[a]) continue;
      uint32_t m = orc_bits_i(key, (uint32_t)a) >> 9;
      if (best < 0 || m > bm) { best = a; bm = m; }
    }
  }
  return best;
}

static void *dog_worker(void *arg) {
  dog_job *j = (dog_job *)arg;
  const dog_cfg *c = &j->cfg;
  const int na = c->play_actions + NCARD;
  uint8_t *mask = (uint8_t *)malloc((size_t)na);
  int64_t total = 0;
  int iters = 0;
  for (int64_t g = j->tid; g < j->n; g += j->nthreads) {
    dog_state s;
    soa_load(c, j->a, g, &s);
    uint32_t rng[2] = {j->rng0[0], j->rng0[1]};
    int len = 0;
    for (int t = 0; t < j->max_steps && !s.done; ++t) {
      uint32_t key[2], nxt[2];
      orc_split_i(rng, (uint32_t)(j->game_offset + g + 1), key);
      orc_split_i(rng, 0, nxt);
      rng[0] = nxt[0]; rng[1] = nxt[1];
      orc_dog_valid_actions_one(c, &s, mask);
      int any = 0;
      for (int k = 0; k < na; ++k) any |= mask[k];
      int8_t r; uint8_t d;
      if (any) orc_dog_step_one(c, &s, categorical_masked(key, mask, na, j->float_gumbel), &r, &d);
      else orc_dog_no_step_one(c, &s, &r, &d);
      ++len;
    }
    soa_store(c, j->a, g, &s);
    if (j->game_len) j->game_len[g] = len;
    total += len;
    if (len > iters) iters = len;
  }
  free(mask);
  j->total = total;
  j->iters = iters;
  return NULL;
}

int orc_dog_play_random(CFG_ARGS, int64_t n, int64_t game_offset, const dog_soa *a, uint32_t *rng_key, int max_steps,
                        int float_gumbel, int nthreads, int32_t *game_len, int64_t *total_steps) {
  MAKE_CFG;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  dog_job jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    dog_job *j = &jobs[t];
    j->cfg = cfg; j->a = a; j->n = n; j->game_offset = game_offset; j->rng0[0] = rng_key[0]; j->rng0[1] = rng_key[1];
    j->max_steps = max_steps; j->float_gumbel = float_gumbel; j->game_len = game_len; j->nthreads = nthreads; j->tid = t;
    j->total = 0; j->iters = 0;
    if (nthreads > 1) pthread_create(&th[t], NULL, dog_worker, j); else dog_worker(j);
  }
  int64_t total = 0;
  int iters = 0;
  for (int t = 0; t < nthreads; ++t) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    total += jobs[t].total;
    if (jobs[t].iters > iters) iters = jobs[t].iters;
  }
  uint32_t k[2] = {rng_key[0], rng_key[1]};
  for (int t = 0; t < iters; ++t) { uint32_t nx[2]; orc_split_i(k, 0, nx); k[0] = nx[0]; k[1] = nx[1]; }
  rng_key[0] = k[0]; rng_key[1] = k[1];
  if (total_steps) *total_steps = total;
  return 0;
}

/* The reference's module-level sub-steps, as its tests call them (DOG/test.py:376-832):
 * kind 0 step_normal_move(env, pin, move)  1 step_neg_move(env, pin, move)  2 step_swap(env, pin, pos)
 * 3 step_hot_7(env, dist[4]).  args int32[n,4]; returns (board, pins, reward, done) without touching env. */
int orc_dog_substep(CFG_ARGS, int64_t n, const dog_soa *a, const int32_t *kind, const int32_t *args, int8_t *board_out,
                    int32_t *pins_out, int8_t *reward, uint8_t *done) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) {
    dog_state s;
    soa_load(&cfg, a, g, &s);
    const int32_t *ar = args + g * 4;
    step_out o;
    if (kind[g] == 0) o = step_normal(&cfg, &s, ar[0], ar[1]);
    else if (kind[g] == 1) o = step_neg(&cfg, &s, ar[0], ar[1]);
    else if (kind[g] == 2) o = step_swap(&cfg, &s, ar[0], ar[1]);
    else { int d[4] = {ar[0], ar[1], ar[2], ar[3]}; o = step_hot7(&cfg, &s, d); }
    memcpy(board_out + g * cfg.total, s.board, (size_t)cfg.total);
    for (int p = 0; p < cfg.n; ++p) for (int i = 0; i < 4; ++i) pins_out[(g * cfg.n + p) * 4 + i] = s.pins[p][i];
    reward[g] = (int8_t)o.reward;
    done[g] = (uint8_t)o.done;
  }
  return 0;
}
