/* madn_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Scalar, clarity-first restatement of the reference's two "Mensch aergere dich nicht"
 * environments.  Every function names the reference lines it follows:
 *   deterministic variant  /root/reference/MADN/deterministic_madn.py
 *   classic (dice) variant /root/reference/MADN/classic_madn.py
 *   goal-lane helper       /root/reference/utils/utility_funcs.py:142-184
 *   random lockstep driver /root/reference/MuZero_det_MADN/evaluate_agent.py:733-930 (do_random)
 * JAX indexing rules are reproduced literally: a gather wraps a
 * negative index once and then clamps; a scatter drops out-of-range updates.
 *
 * PARITY STATUS: pinned by the reference's own 64+64 pytest cases (MADN/test.py), which
 * tests/test_oracle_madn.py replays from tests/golden/madn_reference_cases.json, and by the
 * goldens the reference itself produced under oracle/jaxshim (tests/golden/ npz files).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may call this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <math.h>
#include "../include/dogstep_rules.h"
#include "jaxrand_oracle.h"

typedef struct {
  int n;           /* num_players */
  int d;           /* distance between starts */
  int board_size;  /* 4*d */
  int total;       /* 4*d + 16 */
  int start[4], target[4], goal[4][4]; /* rows [0,n) valid */
  uint32_t rules;
} madn_cfg;

#define RULE(c, bit) (((c)->rules & (bit)) != 0)

/* python floor division / modulo */
static int fdiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
static int fmod_(int a, int b) { int r = a % b; if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }

/* JAX gather index: wrap a negative index once, then clamp into [0,size) */
static int gidx(int i, int size) {
  if (i < 0) i += size;
  if (i < 0) i = 0;
  if (i > size - 1) i = size - 1;
  return i;
}

/* deterministic_madn.py:62-78 / classic_madn.py:72-88 — geometry from (num_players, layout, distance) */
int orc_madn_make_cfg(madn_cfg *c, int num_players, int layout_mask, int distance, uint32_t rules) {
  if (num_players < 1 || num_players > 4 || distance < 1) return -1;
  c->n = num_players;
  c->d = distance;
  c->board_size = 4 * distance;
  c->total = c->board_size + 16;
  /* enable_teams = enable_teams & (num_players == 4)  (:67) */
  if (num_players != 4) rules &= ~DOGSTEP_RULE_TEAMS;
  c->rules = rules;
  int cnt = 0;
  for (int i = 0; i < 4; ++i) cnt += (layout_mask >> i) & 1;
  /* (:70-74) layout falls back to the first n seats */
  if (cnt != num_players || (layout_mask == 0xF && num_players < 4)) layout_mask = (1 << num_players) - 1;
  int r = 0;
  for (int i = 0; i < 4; ++i) {
    if (!((layout_mask >> i) & 1)) continue;
    c->start[r] = i * distance;
    c->target[r] = fmod_(c->start[r] - 1, c->board_size);
    for (int k = 0; k < 4; ++k) c->goal[r][k] = c->board_size + 4 * i + k;
    ++r;
  }
  for (; r < 4; ++r) { c->start[r] = 0; c->target[r] = 0; memset(c->goal[r], 0, sizeof c->goal[r]); }
  return 0;
}

/* set_pins_on_board — deterministic_madn.py:259-271: -1 pins dropped, later writes win */
static void set_pins_on_board(const madn_cfg *c, const int8_t *pins /*[n][4]*/, int8_t *board) {
  for (int k = 0; k < c->total; ++k) board[k] = -1;
  for (int p = 0; p < c->n; ++p)
    for (int i = 0; i < 4; ++i) {
      int pos = pins[p * 4 + i];
      if (pos >= 0 && pos < c->total) board[pos] = (int8_t)p;
    }
}

/* is_player_done — deterministic_madn.py:122-137 */
static int is_player_done(const madn_cfg *c, const int8_t *board, int player) {
  if (player >= c->n) return 0;
  int p = gidx(player, c->n);
  for (int k = 0; k < 4; ++k)
    if (board[gidx(c->goal[p][k], c->total)] < 0) return 0;
  return 1;
}

/* get_winner — deterministic_madn.py:139-168 */
static void get_winner(const madn_cfg *c, const int8_t *board, int winner[4]) {
  int pd[4];
  for (int p = 0; p < 4; ++p) pd[p] = is_player_done(c, board, p);
  if (RULE(c, DOGSTEP_RULE_TEAMS)) {
    int t0 = pd[0] && pd[2], t1 = pd[1] && pd[3];
    int both = t0 && t1, none = !(t0 || t1);
    if (both || none) { winner[0] = winner[1] = winner[2] = winner[3] = 0; }
    else if (t0) { winner[0] = 1; winner[1] = 0; winner[2] = 1; winner[3] = 0; }
    else { winner[0] = 0; winner[1] = 1; winner[2] = 0; winner[3] = 1; }
  } else {
    for (int p = 0; p < 4; ++p) winner[p] = pd[p];
  }
}

/* check_goal_path_for_pin{,2} — utility_funcs.py:142-184: no own pin on lane cells k, s < k < e */
static int goal_path_clear(const madn_cfg *c, const int8_t *board, int cp, int s, int e) {
  for (int k = 0; k < 4; ++k)
    if (s < k && k < e && board[gidx(c->goal[cp][k], c->total)] == cp) return 0;
  return 1;
}

static int in_goal_of(const madn_cfg *c, int cp, int pos) {
  for (int k = 0; k < 4; ++k) if (pos == c->goal[cp][k]) return 1;
  return 0;
}

/* team proxy — deterministic_madn.py:184,310 / classic_madn.py:273,379 */
static int mover_of(const madn_cfg *c, const int8_t *board, int player_id) {
  if (RULE(c, DOGSTEP_RULE_TEAMS) && is_player_done(c, board, player_id)) return fmod_(player_id + 2, 4);
  return player_id;
}

/* One (pin, move) cell of valid_action before the action-set / home-pin handling:
 * deterministic_madn.py:323-381, classic_madn.py:392-445 (same geometry). */
static int board_move_ok(const madn_cfg *c, const int8_t *board, int cp, int pos, int m) {
  const int n = c->n, bs = c->board_size;
  int mts = RULE(c, DOGSTEP_RULE_MUST_TRAVERSE_START);
  int moved = pos + m;
  int fitted = fmod_(moved, bs);
  int target = c->target[cp];
  int x = moved - target - mts;
  /* :329 own pin on the landing cell */
  int result = (board[gidx(fitted, c->total)] != cp) || RULE(c, DOGSTEP_RULE_FRIENDLY_FIRE);
  /* :332-340 start blocking */
  int nsb = fmod_(fdiv(pos, c->d) + 1, n);
  int nsa = fdiv(fitted, c->d);
  int trav = c->start[gidx(nsb, n)] == c->start[gidx(nsa, n)];
  int q = gidx(nsa, n);
  int pins_on_start_after = board[gidx(c->start[q], c->total)] == q;
  if (RULE(c, DOGSTEP_RULE_START_BLOCKING) && trav)
    result = (!pins_on_start_after || pos == c->start[cp]) && result;
  /* :343-347 */
  if (mts && RULE(c, DOGSTEP_RULE_START_BLOCKING) && trav && pins_on_start_after) x = 0;
  /* :349-357 */
  if (!RULE(c, DOGSTEP_RULE_CIRCULAR_BOARD))
    if (pos <= target && (x > 4 || (x == 0 && mts))) result = 0;
  /* :360-372 goal entry window */
  int A = RULE(c, DOGSTEP_RULE_CIRCULAR_BOARD) && result;
  int B = board[gidx(c->goal[cp][gidx(x - 1, 4)], c->total)] != cp;
  int C = RULE(c, DOGSTEP_RULE_JUMP_IN_GOAL) || goal_path_clear(c, board, cp, -1, x);
  if (4 >= x && x > 0 && pos <= target) result = A || (B && C);
  /* :376-381 pin already in its goal lane */
  int g0 = c->goal[cp][0];
  int D = RULE(c, DOGSTEP_RULE_JUMP_IN_GOAL) || goal_path_clear(c, board, cp, pos - g0, moved - g0 + 1);
  if (in_goal_of(c, cp, pos))
    result = (moved <= c->goal[cp][3]) && (board[gidx(moved, c->total)] != cp) && D;
  return result;
}

/* valid_action — deterministic_madn.py:299-393 → bool[4][6] */
void orc_madn_det_valid_action_one(const madn_cfg *c, const int8_t *board, int player_id,
                                   const int8_t *pins, const int8_t *action_set, uint8_t *mask /*[24]*/) {
  int cp = mover_of(c, board, player_id);
  cp = gidx(cp, c->n);
  for (int i = 0; i < 4; ++i) {
    int pos = pins[cp * 4 + i];
    for (int m = 1; m <= 6; ++m) {
      int result = board_move_ok(c, board, cp, pos, m);
      if (pos == -1) { /* :383-392 — note: compares with the UN-proxied player id */
        int is_start_move = RULE(c, DOGSTEP_RULE_START_ON_1) ? (m == 1 || m == 6) : (m == 6);
        result = is_start_move && (board[gidx(c->start[cp], c->total)] != player_id);
      }
      mask[i * 6 + (m - 1)] = (uint8_t)(result && action_set[cp * 6 + (m - 1)] > 0);
    }
  }
}

/* valid_action — classic_madn.py:367-461 → bool[4] */
void orc_madn_cls_valid_action_one(const madn_cfg *c, const int8_t *board, int player_id,
                                   const int8_t *pins, int die, uint8_t *mask /*[4]*/) {
  int cp = gidx(mover_of(c, board, player_id), c->n);
  int pins_on_own_start = board[gidx(c->start[cp], c->total)] == cp;
  for (int i = 0; i < 4; ++i) {
    int pos = pins[cp * 4 + i];
    int result = board_move_ok(c, board, cp, pos, die);
    if (pos == -1) { /* :448-459 */
      int is_start_move = RULE(c, DOGSTEP_RULE_START_ON_1) ? (die == 1 || die == 6) : (die == -1 || die == 6);
      result = is_start_move && !pins_on_own_start;
    }
    mask[i] = (uint8_t)result;
  }
}

/* shared move application: deterministic_madn.py:188-230 / classic_madn.py:278-321.
 * Writes pins/board in place; returns nothing; `invalid` decided by the caller's mask. */
static void apply_move(const madn_cfg *c, int8_t *board, int8_t *pins, int cp, int pin, int move, int invalid) {
  const int bs = c->board_size;
  int mts = RULE(c, DOGSTEP_RULE_MUST_TRAVERSE_START);
  int pos = pins[cp * 4 + pin];
  int moved = (int8_t)(pos + move);
  int fitted = fmod_(moved, bs);
  int x = (int8_t)(moved - c->target[cp] - mts);
  int g0 = c->goal[cp][0];
  int a = in_goal_of(c, cp, pos) ? goal_path_clear(c, board, cp, pos - g0, moved - g0 + 1)
                                 : goal_path_clear(c, board, cp, -1, x);
  int gx = c->goal[cp][gidx(x - 1, 4)];
  int A = (board[gidx(gx, c->total)] != cp) && (RULE(c, DOGSTEP_RULE_JUMP_IN_GOAL) || a);
  int new_pos;
  if (pos == -1) new_pos = c->start[cp];
  else if (in_goal_of(c, cp, pos)) new_pos = moved;
  else if (4 >= x && x > 0 && A && pos <= c->target[cp]) new_pos = gx;
  else new_pos = fitted;
  int pin_at_pos = board[gidx(new_pos, c->total)];
  if (pin_at_pos != -1 && (pin_at_pos != cp || RULE(c, DOGSTEP_RULE_FRIENDLY_FIRE)) && !invalid) {
    int q = gidx(pin_at_pos, c->n);
    for (int k = 0; k < 4; ++k)
      if (pins[q * 4 + k] == new_pos) pins[q * 4 + k] = -1;
  }
  if (!invalid) {
    pins[cp * 4 + pin] = (int8_t)new_pos;
    set_pins_on_board(c, pins, board);
  }
}

/* env_step — deterministic_madn.py:170-257.  action = [pin, move]. */
void orc_madn_det_step_one(const madn_cfg *c, int8_t *board, int8_t *current_player, int8_t *pins,
                           int8_t *reward, uint8_t *done, int8_t *action_set, int pin_in, int move_in) {
  int player_id = *current_player;
  int cp = gidx(mover_of(c, board, player_id), c->n);
  uint8_t mask[24];
  orc_madn_det_valid_action_one(c, board, player_id, pins, action_set, mask);
  int pin = gidx(pin_in, 4), mi = gidx(move_in - 1, 6);
  int invalid = !mask[pin * 6 + mi];
  int8_t old_as[24];
  memcpy(old_as, action_set, 24);
  apply_move(c, board, pins, cp, pin, move_in, invalid);
  /* :232-240 action-set bookkeeping incl. the pre-decrement refill quirk */
  int in_range = (move_in - 1 >= -6 && move_in - 1 < 6); /* scatter drops OOB; negative wraps once */
  int curr = old_as[cp * 6 + mi];
  if (in_range) action_set[cp * 6 + mi] = (int8_t)((invalid || curr == 0) ? curr : curr - 1);
  int all_zero = 1;
  for (int k = 0; k < 6; ++k) all_zero &= (action_set[cp * 6 + k] == 0);
  if (all_zero) {
    memcpy(action_set, old_as, 24);
    int row = gidx(player_id, c->n); /* scatter on env.current_player; in range for valid states */
    if (player_id >= -c->n && player_id < c->n)
      for (int k = 0; k < 6; ++k) action_set[row * 6 + k] = 4;
  }
  int winner[4];
  get_winner(c, board, winner);
  int8_t rew = (int8_t)(*done ? 0 : (invalid ? -1 : winner[gidx(cp, 4)]));
  int any = winner[0] | winner[1] | winner[2] | winner[3];
  uint8_t dn = (uint8_t)(*done || any);
  int bonus = RULE(c, DOGSTEP_RULE_BONUS_TURN_ON_6) && ((int8_t)move_in == 6);
  *current_player = (int8_t)((dn || bonus) ? player_id : fmod_(player_id + 1, c->n));
  *reward = rew;
  *done = dn;
}

/* no_step — deterministic_madn.py:283-297 */
void orc_madn_det_no_step_one(const madn_cfg *c, int8_t *current_player, int8_t *action_set) {
  int pid = *current_player;
  if (pid >= -c->n && pid < c->n) {
    int row = gidx(pid, c->n);
    for (int k = 0; k < 6; ++k) action_set[row * 6 + k] = 4;
  }
  *current_player = (int8_t)fmod_(pid + 1, c->n);
}

/* env_step — classic_madn.py:257-337.  action = pin; move = env.die */
void orc_madn_cls_step_one(const madn_cfg *c, int8_t *board, int8_t *current_player, int8_t *pins,
                           int8_t *reward, uint8_t *done, int die, int pin_in) {
  int player_id = *current_player;
  int cp = gidx(mover_of(c, board, player_id), c->n);
  uint8_t mask[4];
  orc_madn_cls_valid_action_one(c, board, player_id, pins, die, mask);
  int pin = gidx(pin_in, 4);
  int invalid = !mask[pin];
  apply_move(c, board, pins, cp, pin, die, invalid);
  int winner[4];
  get_winner(c, board, winner);
  int8_t rew = (int8_t)(*done ? 0 : (invalid ? -1 : winner[gidx(cp, 4)]));
  int any = winner[0] | winner[1] | winner[2] | winner[3];
  uint8_t dn = (uint8_t)(*done || any);
  int bonus = RULE(c, DOGSTEP_RULE_BONUS_TURN_ON_6) && (die == 6);
  *current_player = (int8_t)((dn || bonus) ? player_id : fmod_(player_id + 1, c->n));
  *reward = rew;
  *done = dn;
}

/* is_soft_locked — classic_madn.py:180-206 (uses the UN-proxied current player) */
static int is_soft_locked(const madn_cfg *c, const int8_t *board, int player_id, const int8_t *pins) {
  int p = gidx(player_id, c->n);
  int not_home = 4;
  for (int i = 0; i < 4; ++i) not_home -= (pins[p * 4 + i] == -1);
  if (not_home <= 0) return 1;
  for (int k = 0; k < 4; ++k) {
    int relevant = k >= 4 - not_home;
    int occupied = board[gidx(c->goal[p][k], c->total)] == player_id;
    if (!(occupied || !relevant)) return 0;
  }
  return 1;
}

/* dice_probabilities — classic_madn.py:14-18,208-228 (f32 literals rounded from doubles) */
void orc_madn_cls_dice_probabilities_one(const madn_cfg *c, const int8_t *board, int player_id,
                                         const int8_t *pins, float p[6]) {
  int locked = is_soft_locked(c, board, player_id, pins);
  if (locked && RULE(c, DOGSTEP_RULE_DICE_RETHROW)) {
    if (RULE(c, DOGSTEP_RULE_START_ON_1)) {
      const double v[6] = {76.0 / 216, 16.0 / 216, 16.0 / 216, 16.0 / 216, 16.0 / 216, 76.0 / 216};
      for (int k = 0; k < 6; ++k) p[k] = (float)v[k];
    } else {
      const double v[6] = {25.0 / 216, 25.0 / 216, 25.0 / 216, 25.0 / 216, 25.0 / 216, 91.0 / 216};
      for (int k = 0; k < 6; ++k) p[k] = (float)v[k];
    }
  } else {
    for (int k = 0; k < 6; ++k) p[k] = (float)(1.0 / 6);
  }
}

/* throw_die — classic_madn.py:230-242 */
void orc_madn_cls_throw_die_one(const madn_cfg *c, const int8_t *board, int player_id, const int8_t *pins,
                                uint32_t key[2], int8_t *die) {
  uint32_t knew[2], sub[2];
  orc_split_i(key, 0, knew);
  orc_split_i(key, 1, sub);
  float p[6];
  orc_madn_cls_dice_probabilities_one(c, board, player_id, pins, p);
  *die = (int8_t)(orc_choice6(sub, p) + 1);
  key[0] = knew[0];
  key[1] = knew[1];
}

/* encode_board — deterministic_madn.py:395-438 → int8[34][total]; classic_madn.py:463-497 → int8[11][total] */
static void encode_common(const madn_cfg *c, const int8_t *board, int cur, const int8_t *pins, int8_t *obs) {
  const int n = c->n, T = c->total, bs = c->board_size;
  int8_t rb[64];
  /* jnp.roll(x, -s)[i] = x[(i+s) % len] */
  for (int i = 0; i < bs; ++i) rb[i] = board[fmod_(i + c->d * cur, bs)];
  for (int i = 0; i < 16; ++i) rb[bs + i] = board[bs + fmod_(i + 4 * cur, 16)];
  int rolled[4];
  for (int r = 0; r < n; ++r) rolled[r] = fmod_(r + cur, n);
  for (int r = 0; r < n; ++r)
    for (int k = 0; k < T; ++k) obs[r * T + k] = (int8_t)(rb[k] == rolled[r]);
  for (int k = 0; k < T; ++k) {
    int team = 0, opp = 0;
    if (RULE(c, DOGSTEP_RULE_TEAMS)) {
      for (int r = 0; r < n; r += 2) team += obs[r * T + k];
      for (int r = 1; r < n; r += 2) opp += obs[r * T + k];
    } else {
      team = obs[k];
      for (int r = 1; r < n; ++r) opp += obs[r * T + k];
    }
    obs[n * T + k] = (int8_t)team;
    obs[(n + 1) * T + k] = (int8_t)opp;
  }
  for (int r = 0; r < n; ++r) {
    int cnt = 0;
    for (int i = 0; i < 4; ++i) cnt += (pins[rolled[r] * 4 + i] == -1);
    for (int k = 0; k < T; ++k) obs[(n + 2 + r) * T + k] = (int8_t)cnt;
  }
}

void orc_madn_det_encode_board_one(const madn_cfg *c, const int8_t *board, int cur, const int8_t *pins,
                                   const int8_t *action_set, int8_t *obs) {
  const int n = c->n, T = c->total;
  encode_common(c, board, cur, pins, obs);
  int base = 2 * n + 2;
  for (int r = 0; r < n; ++r) {
    int src = fmod_(r + cur, n);
    for (int a = 0; a < 6; ++a)
      for (int k = 0; k < T; ++k) obs[(base + r * 6 + a) * T + k] = action_set[src * 6 + a];
  }
}

void orc_madn_cls_encode_board_one(const madn_cfg *c, const int8_t *board, int cur, const int8_t *pins,
                                   int die, int8_t *obs) {
  const int n = c->n, T = c->total;
  encode_common(c, board, cur, pins, obs);
  for (int k = 0; k < T; ++k) obs[(2 * n + 2) * T + k] = (int8_t)die;
}

/* ------------------------------------------------------------------------------------------
 * Batched entry points over SoA leaves (the layout a vmapped reference pytree has).
 * ------------------------------------------------------------------------------------------ */
#define CFG_ARGS int num_players, int layout_mask, int distance, uint32_t rules
#define MAKE_CFG madn_cfg cfg; if (orc_madn_make_cfg(&cfg, num_players, layout_mask, distance, rules)) return -1; \
                 const int T = cfg.total, NP = cfg.n; (void)T; (void)NP

int orc_madn_geometry(CFG_ARGS, int32_t *start, int32_t *target, int32_t *goal) {
  MAKE_CFG;
  for (int p = 0; p < NP; ++p) {
    start[p] = cfg.start[p];
    target[p] = cfg.target[p];
    for (int k = 0; k < 4; ++k) goal[p * 4 + k] = cfg.goal[p][k];
  }
  return 0;
}

/* env_reset — deterministic_madn.py:42-120 / classic_madn.py:51-131 (action_set/die may be NULL) */
int orc_madn_reset(CFG_ARGS, int64_t n, const int32_t *seeds, int starting_player, int8_t *board,
                   int8_t *current_player, int8_t *pins, int8_t *reward, uint8_t *done,
                   int8_t *action_set, int8_t *die, uint32_t *key) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) {
    uint32_t k0[2], knew[2], sub[2];
    orc_prngkey(seeds[g], k0);
    orc_split_i(k0, 0, knew);
    orc_split_i(k0, 1, sub);
    int sp = starting_player;
    if (sp < 0 || sp >= NP) sp = orc_randint_i(sub, 0, 0, NP);
    int8_t *pg = pins + g * NP * 4;
    for (int i = 0; i < NP * 4; ++i) pg[i] = -1;
    if (RULE(&cfg, DOGSTEP_RULE_INITIAL_FREE_PIN))
      for (int p = 0; p < NP; ++p) pg[p * 4] = (int8_t)cfg.start[p];
    set_pins_on_board(&cfg, pg, board + g * T);
    current_player[g] = (int8_t)sp;
    reward[g] = 0;
    done[g] = 0;
    if (action_set) for (int i = 0; i < NP * 6; ++i) action_set[g * NP * 6 + i] = 4;
    if (die) die[g] = 0;
    key[g * 2] = knew[0];
    key[g * 2 + 1] = knew[1];
  }
  return 0;
}

int orc_madn_set_pins_on_board(CFG_ARGS, int64_t n, const int8_t *pins, int8_t *board) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) set_pins_on_board(&cfg, pins + g * NP * 4, board + g * T);
  return 0;
}

int orc_madn_det_valid_action(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player,
                              const int8_t *pins, const int8_t *action_set, uint8_t *mask) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_det_valid_action_one(&cfg, board + g * T, current_player[g], pins + g * NP * 4,
                                  action_set + g * NP * 6, mask + g * 24);
  return 0;
}

int orc_madn_det_step(CFG_ARGS, int64_t n, int8_t *board, int8_t *current_player, int8_t *pins,
                      int8_t *reward, uint8_t *done, int8_t *action_set, const int8_t *action) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_det_step_one(&cfg, board + g * T, current_player + g, pins + g * NP * 4, reward + g, done + g,
                          action_set + g * NP * 6, action[g * 2], action[g * 2 + 1]);
  return 0;
}

int orc_madn_det_no_step(CFG_ARGS, int64_t n, int8_t *current_player, int8_t *reward, const uint8_t *done,
                         int8_t *action_set) {
  MAKE_CFG;
  (void)done;
  for (int64_t g = 0; g < n; ++g) {
    orc_madn_det_no_step_one(&cfg, current_player + g, action_set + g * NP * 6);
    (void)reward; /* env.reward is NOT touched by no_step; the returned reward is the constant 0 */
  }
  return 0;
}

int orc_madn_det_encode_board(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player,
                              const int8_t *pins, const int8_t *action_set, int8_t *obs) {
  MAKE_CFG;
  const int C = 8 * NP + 2;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_det_encode_board_one(&cfg, board + g * T, current_player[g], pins + g * NP * 4,
                                  action_set + g * NP * 6, obs + g * C * T);
  return 0;
}

int orc_madn_cls_valid_action(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player,
                              const int8_t *pins, const int8_t *die, uint8_t *mask) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_cls_valid_action_one(&cfg, board + g * T, current_player[g], pins + g * NP * 4, die[g], mask + g * 4);
  return 0;
}

int orc_madn_cls_step(CFG_ARGS, int64_t n, int8_t *board, int8_t *current_player, int8_t *pins,
                      int8_t *reward, uint8_t *done, const int8_t *die, const int8_t *action) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_cls_step_one(&cfg, board + g * T, current_player + g, pins + g * NP * 4, reward + g, done + g,
                          die[g], action[g]);
  return 0;
}

int orc_madn_cls_no_step(CFG_ARGS, int64_t n, int8_t *current_player) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g) current_player[g] = (int8_t)fmod_(current_player[g] + 1, NP);
  return 0;
}

int orc_madn_cls_throw_die(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player,
                           const int8_t *pins, uint32_t *key, int8_t *die) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_cls_throw_die_one(&cfg, board + g * T, current_player[g], pins + g * NP * 4, key + g * 2, die + g);
  return 0;
}

int orc_madn_cls_dice_probabilities(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player,
                                    const int8_t *pins, float *p) {
  MAKE_CFG;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_cls_dice_probabilities_one(&cfg, board + g * T, current_player[g], pins + g * NP * 4, p + g * 6);
  return 0;
}

int orc_madn_cls_encode_board(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player,
                              const int8_t *pins, const int8_t *die, int8_t *obs) {
  MAKE_CFG;
  const int C = 2 * NP + 3;
  for (int64_t g = 0; g < n; ++g)
    orc_madn_cls_encode_board_one(&cfg, board + g * T, current_player[g], pins + g * NP * 4, die[g],
                                  obs + g * C * T);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Random-legal-policy lockstep driver (config 2): evaluate_agent.py:733-930 with every seat
 * type==3 (do_random :772-776), loop cap :918.  Per lockstep iteration t:
 *     rng, *step_keys = split(rng, N+1)        (:741)  -> carry = index 0, game j = index j+1
 *     not done: mask = valid_action(env); any(mask) ? env_step(categorical(key, where(mask,0,-1e9)))
 *                                                    : no_step(env)
 * categorical = argmax(logits + gumbel(key,(24,))).  float_gumbel=1 evaluates that literally with
 * libm logf; float_gumbel=0 uses the exact integer equivalent (first valid action with the largest
 * 23-bit uniform mantissa; -log(-log(u)) is strictly increasing), which is what the CUDA path does.
 * ------------------------------------------------------------------------------------------ */
static int categorical_valid(const uint32_t key[2], const uint8_t *mask, int na, int float_gumbel) {
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
      uint32_t mant = orc_bits_i(key, (uint32_t)a) >> 9;
      if (best < 0 || mant > bm) { best = a; bm = mant; }
    }
  }
  return best;
}

typedef struct {
  madn_cfg cfg;
  int64_t n, game_offset;
  int8_t *board, *current_player, *pins, *reward, *action_set;
  uint8_t *done;
  uint32_t rng0[2];
  int max_steps, float_gumbel;
  int32_t *game_len;
  int nthreads, tid;
  int64_t total;
  int iters_needed;
} play_job;

static void *play_random_worker(void *arg) {
  play_job *j = (play_job *)arg;
  const madn_cfg *cfg = &j->cfg;
  const int T = cfg->total, NP = cfg->n;
  int64_t total = 0;
  int iters_needed = 0;
  /* interleaved blocks of 64 games per thread */
  for (int64_t blk = j->tid; blk * 64 < j->n; blk += j->nthreads) {
    int64_t hi = (blk + 1) * 64 < j->n ? (blk + 1) * 64 : j->n;
    for (int64_t g = blk * 64; g < hi; ++g) {
      uint32_t rng[2] = {j->rng0[0], j->rng0[1]};
      int len = 0;
      for (int t = 0; t < j->max_steps && !j->done[g]; ++t) {
        uint32_t key[2], nxt[2];
        orc_split_i(rng, (uint32_t)(j->game_offset + g + 1), key);
        orc_split_i(rng, 0, nxt);
        rng[0] = nxt[0];
        rng[1] = nxt[1];
        uint8_t mask[24];
        orc_madn_det_valid_action_one(cfg, j->board + g * T, j->current_player[g], j->pins + g * NP * 4,
                                      j->action_set + g * NP * 6, mask);
        int any = 0;
        for (int k = 0; k < 24; ++k) any |= mask[k];
        if (any) {
          int a = categorical_valid(key, mask, 24, j->float_gumbel);
          /* map_action — deterministic_madn.py:469-479 */
          orc_madn_det_step_one(cfg, j->board + g * T, j->current_player + g, j->pins + g * NP * 4,
                                j->reward + g, j->done + g, j->action_set + g * NP * 6, a / 6, a % 6 + 1);
        } else {
          orc_madn_det_no_step_one(cfg, j->current_player + g, j->action_set + g * NP * 6);
        }
        ++len;
      }
      if (j->game_len) j->game_len[g] = len;
      total += len;
      if (len > iters_needed) iters_needed = len;
    }
  }
  j->total = total;
  j->iters_needed = iters_needed;
  return NULL;
}

int orc_madn_det_play_random(CFG_ARGS, int64_t n, int64_t game_offset, int8_t *board, int8_t *current_player,
                             int8_t *pins, int8_t *reward, uint8_t *done, int8_t *action_set,
                             uint32_t *rng_key /*[2] in/out*/, int max_steps, int float_gumbel, int nthreads,
                             int32_t *game_len /*[n] active iterations*/, int64_t *total_steps) {
  MAKE_CFG;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  play_job jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    play_job *j = &jobs[t];
    j->cfg = cfg; j->n = n; j->game_offset = game_offset;
    j->board = board; j->current_player = current_player; j->pins = pins; j->reward = reward;
    j->action_set = action_set; j->done = done;
    j->rng0[0] = rng_key[0]; j->rng0[1] = rng_key[1];
    j->max_steps = max_steps; j->float_gumbel = float_gumbel; j->game_len = game_len;
    j->nthreads = nthreads; j->tid = t; j->total = 0; j->iters_needed = 0;
    if (nthreads > 1) pthread_create(&th[t], NULL, play_random_worker, j);
    else play_random_worker(j);
  }
  int64_t total = 0;
  int iters_needed = 0;
  for (int t = 0; t < nthreads; ++t) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    total += jobs[t].total;
    if (jobs[t].iters_needed > iters_needed) iters_needed = jobs[t].iters_needed;
  }
  /* the carried key advances once per lockstep iteration the while_loop actually ran */
  uint32_t final_key[2] = {rng_key[0], rng_key[1]};
  for (int t = 0; t < iters_needed; ++t) {
    uint32_t nxt[2];
    orc_split_i(final_key, 0, nxt);
    final_key[0] = nxt[0];
    final_key[1] = nxt[1];
  }
  rng_key[0] = final_key[0];
  rng_key[1] = final_key[1];
  if (total_steps) *total_steps = total;
  return 0;
}

/* exported thin wrappers of the jax.random restatement, for tests */
void orc_threefry2x32_v(uint32_t k0, uint32_t k1, int64_t n, const uint32_t *c0, const uint32_t *c1,
                        uint32_t *o0, uint32_t *o1) {
  for (int64_t i = 0; i < n; ++i) orc_threefry2x32(k0, k1, c0[i], c1[i], o0 + i, o1 + i);
}
void orc_split(const uint32_t *key, int64_t n, uint32_t *out /*[n,2]*/) {
  for (int64_t i = 0; i < n; ++i) orc_split_i(key, (uint32_t)i, out + 2 * i);
}
void orc_random_bits(const uint32_t *key, int64_t n, uint32_t *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = orc_bits_i(key, (uint32_t)i);
}
void orc_uniform(const uint32_t *key, int64_t n, float lo, float hi, float *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = orc_uniform_i(key, (uint32_t)i, lo, hi);
}
void orc_randint(const uint32_t *key, int64_t n, int32_t lo, int32_t hi, int32_t *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = orc_randint_i(key, (uint32_t)i, lo, hi);
}
void orc_gumbel(const uint32_t *key, int64_t n, float *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = orc_gumbel_i(key, (uint32_t)i);
}
int orc_choice6_v(const uint32_t *key, const float *p) { return orc_choice6(key, p); }
int orc_categorical(const uint32_t *key, const uint8_t *mask, int na, int float_gumbel) {
  return categorical_valid(key, mask, na, float_gumbel);
}


/* ------------------------------------------------------------------------------------------
 * True-env mctx callbacks of the deterministic game — MADN/deterministic_madn.py:480-590
 * (winning_action, policy_function, rollout, value_function, root_fn, recurrent_fn).
 * A state here is the six leaves of one game; `emb` is the same state as floats:
 * board[total], current_player, pins[4 n], reward, done, action_set[6 n].
 * The reference's rollout returns jnp.where(winner == -1, 0.0, jnp.where(winner[root_player], 1.0, -1.0)) with `winner`
 * the BOOL array of get_winner: `winner == -1` is never true, so the value is a float32[4] of four equal entries,
 * +1 if the root player('s team) has won when the rollout stops, else -1 (also when the 300-step cap ends it).  The scalar
 * is returned.  Pinned by tests/golden/madn_det_reference_trueenv.npz (the reference's own functions on the shim).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int8_t board[64];
  int8_t cur, reward;
  uint8_t done;
  int8_t pins[16];
  int8_t aset[24];
} det_env;

static float f_log_d(float x) { return (float)log((double)x); } /* float contract of DESIGN 5 */

static void det_policy_one(const madn_cfg *c, const det_env *e, float logits[24]) {
  uint8_t mask[24];
  orc_madn_det_valid_action_one(c, e->board, e->cur, e->pins, e->aset, mask);
  for (int a = 0; a < 24; ++a) {  /* winning_action (:480-493): env_step on a copy, reward == 1 */
    det_env t = *e;
    orc_madn_det_step_one(c, t.board, &t.cur, t.pins, &t.reward, &t.done, t.aset, a / 6, a % 6 + 1);
    logits[a] = (mask[a] ? 100.0f : 0.0f) + (t.reward == 1 ? 200.0f : 0.0f);
  }
}

static float det_rollout_one(const madn_cfg *c, const det_env *e0, const uint32_t key_in[2]) {
  det_env e = *e0;
  uint32_t key[2] = {key_in[0], key_in[1]};
  for (int steps = 0; !e.done && steps < 300; ++steps) {
    uint32_t nk[2], sub[2];
    orc_split_i(key, 0, nk); orc_split_i(key, 1, sub);
    key[0] = nk[0]; key[1] = nk[1];
    uint8_t mask[24];
    orc_madn_det_valid_action_one(c, e.board, e.cur, e.pins, e.aset, mask);
    int any = 0;
    for (int a = 0; a < 24; ++a) any |= mask[a];
    if (!any) { orc_madn_det_no_step_one(c, &e.cur, e.aset); continue; }
    float lg[24];
    det_policy_one(c, &e, lg);
    int best = 0; float bv = 0.f;
    for (int a = 0; a < 24; ++a) {  /* jax.random.categorical: first maximum of gumbel + logits */
      float u = orc_uniform_i(sub, (uint32_t)a, 1.17549435e-38f, 1.0f);
      float v = -f_log_d(-f_log_d(u)) + lg[a];
      if (a == 0 || v > bv) { bv = v; best = a; }
    }
    orc_madn_det_step_one(c, e.board, &e.cur, e.pins, &e.reward, &e.done, e.aset, best / 6, best % 6 + 1);
  }
  int winner[4];
  get_winner(c, e.board, winner);
  return winner[gidx(e0->cur, 4)] ? 1.0f : -1.0f;
}

static void det_ld(const madn_cfg *c, det_env *e, int64_t g, const int8_t *board, const int8_t *cur, const int8_t *pins,
                   const int8_t *reward, const uint8_t *done, const int8_t *aset) {
  memset(e, 0, sizeof(*e));
  memcpy(e->board, board + g * c->total, c->total);
  e->cur = cur[g]; e->reward = reward[g]; e->done = done[g];
  memcpy(e->pins, pins + g * c->n * 4, c->n * 4);
  memcpy(e->aset, aset + g * c->n * 6, c->n * 6);
}
static void det_to_emb(const madn_cfg *c, const det_env *e, float *f) {
  int k = 0;
  for (int i = 0; i < c->total; ++i) f[k++] = e->board[i];
  f[k++] = e->cur;
  for (int i = 0; i < c->n * 4; ++i) f[k++] = e->pins[i];
  f[k++] = e->reward; f[k++] = e->done;
  for (int i = 0; i < c->n * 6; ++i) f[k++] = e->aset[i];
}
static void det_from_emb(const madn_cfg *c, det_env *e, const float *f) {
  int k = 0;
  memset(e, 0, sizeof(*e));
  for (int i = 0; i < c->total; ++i) e->board[i] = (int8_t)f[k++];
  e->cur = (int8_t)f[k++];
  for (int i = 0; i < c->n * 4; ++i) e->pins[i] = (int8_t)f[k++];
  e->reward = (int8_t)f[k++]; e->done = (uint8_t)f[k++];
  for (int i = 0; i < c->n * 6; ++i) e->aset[i] = (int8_t)f[k++];
}

int orc_madn_det_embed_dim(CFG_ARGS) { MAKE_CFG; return cfg.total > 64 ? -1 : cfg.total + 10 * cfg.n + 3; }

/* policy_function (:495-507) -> f32 [n, 24] */
int orc_madn_det_policy_function(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player, const int8_t *pins,
                                 const int8_t *reward, const uint8_t *done, const int8_t *action_set, float *logits) {
  MAKE_CFG;
  if (cfg.total > 64) return -1;
  for (int64_t g = 0; g < n; ++g) {
    det_env e; det_ld(&cfg, &e, g, board, current_player, pins, reward, done, action_set);
    det_policy_one(&cfg, &e, logits + 24 * g);
  }
  return 0;
}

/* root_fn (:551-566): prior = policy_function(env), value = rollout(env, key), embedding = env.  keys u32 [n, 2] */
int orc_madn_det_root_fn(CFG_ARGS, int64_t n, const int8_t *board, const int8_t *current_player, const int8_t *pins,
                         const int8_t *reward, const uint8_t *done, const int8_t *action_set, const uint32_t *keys,
                         float *prior, float *value, float *emb) {
  MAKE_CFG;
  if (cfg.total > 64) return -1;
  const int E = cfg.total + 10 * cfg.n + 3;
  for (int64_t g = 0; g < n; ++g) {
    det_env e; det_ld(&cfg, &e, g, board, current_player, pins, reward, done, action_set);
    det_policy_one(&cfg, &e, prior + 24 * g);
    value[g] = det_rollout_one(&cfg, &e, keys + 2 * g);
    det_to_emb(&cfg, &e, emb + (int64_t)E * g);
  }
  return 0;
}

/* recurrent_fn (:568-590): env_step(embedding, map_action(action)); reward, discount = done ? 0 : -1, prior = policy_function,
 * value = done ? 0 : rollout */
int orc_madn_det_recurrent_fn(CFG_ARGS, int64_t n, const uint32_t *keys, const int32_t *action, const float *emb_in,
                              float *prior, float *value, float *reward, float *discount, float *emb_out) {
  MAKE_CFG;
  if (cfg.total > 64) return -1;
  const int E = cfg.total + 10 * cfg.n + 3;
  for (int64_t g = 0; g < n; ++g) {
    det_env e; det_from_emb(&cfg, &e, emb_in + (int64_t)E * g);
    const int a = action[g];
    orc_madn_det_step_one(&cfg, e.board, &e.cur, e.pins, &e.reward, &e.done, e.aset, (int8_t)fdiv(a, 6), (int8_t)(fmod_(a, 6) + 1)); /* map_action (:469-479): both int8 */
    reward[g] = (float)e.reward;
    discount[g] = e.done ? 0.0f : -1.0f;
    det_policy_one(&cfg, &e, prior + 24 * g);
    value[g] = e.done ? 0.0f : det_rollout_one(&cfg, &e, keys + 2 * g);
    det_to_emb(&cfg, &e, emb_out + (int64_t)E * g);
  }
  return 0;
}


/* ------------------------------------------------------------------------------------------
 * True-env mctx callbacks of the dice game — MADN/classic_madn.py:541-714 (winning_action, policy_function, rollout,
 * root_fn, recurrent_fn = decision node -> afterstate, recurrent_chance_fn = afterstate + die -> state).  SURVEY 8 row b4.
 * The reference as it stands raises before any of them returns: winning_action (:551-565) builds its scratch copy without the
 * dataclass's `key` field.  env_step never reads that field; tests/golden/gen_madn_cls_trueenv_goldens.py runs the
 * reference's own function bodies with that one constructor call made to go through, and this restatement is pinned to
 * its outputs (tests/golden/madn_cls_reference_trueenv.npz).
 * emb: board[total], current_player, pins[4 n], reward, done, die, key as four 16-bit halves (k0 lo, k0 hi, k1 lo, k1 hi —
 * a float32 cannot hold a uint32).  The rollout value is +-1 exactly as in the deterministic game (see above).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int8_t board[64];
  int8_t cur, reward, die;
  uint8_t done;
  int8_t pins[16];
  uint32_t key[2];
} cls_env;

static void cls_policy_one(const madn_cfg *c, const cls_env *e, float logits[4]) {
  uint8_t mask[4];
  orc_madn_cls_valid_action_one(c, e->board, e->cur, e->pins, e->die, mask);
  for (int a = 0; a < 4; ++a) {  /* winning_action (:543-569) */
    cls_env t = *e;
    orc_madn_cls_step_one(c, t.board, &t.cur, t.pins, &t.reward, &t.done, t.die, a);
    logits[a] = (mask[a] ? 100.0f : 0.0f) + (t.reward == 1 ? 200.0f : 0.0f);
  }
}

static float cls_rollout_one(const madn_cfg *c, const cls_env *e0, const uint32_t key_in[2]) {
  cls_env e = *e0;
  uint32_t key[2] = {key_in[0], key_in[1]};
  for (int steps = 0; !e.done && steps < 300; ++steps) {
    uint32_t nk[2], sub[2];
    orc_split_i(key, 0, nk); orc_split_i(key, 1, sub);
    key[0] = nk[0]; key[1] = nk[1];
    orc_madn_cls_throw_die_one(c, e.board, e.cur, e.pins, e.key, &e.die);  /* env = throw_die(env): the ENV's key chain */
    uint8_t mask[4];
    orc_madn_cls_valid_action_one(c, e.board, e.cur, e.pins, e.die, mask);
    if (!(mask[0] | mask[1] | mask[2] | mask[3])) { e.cur = (int8_t)fmod_(e.cur + 1, c->n); continue; }  /* no_step (:353-365) */
    float lg[4];
    cls_policy_one(c, &e, lg);
    int best = 0; float bv = 0.f;
    for (int a = 0; a < 4; ++a) {
      float u = orc_uniform_i(sub, (uint32_t)a, 1.17549435e-38f, 1.0f);
      float v = -f_log_d(-f_log_d(u)) + lg[a];
      if (a == 0 || v > bv) { bv = v; best = a; }
    }
    orc_madn_cls_step_one(c, e.board, &e.cur, e.pins, &e.reward, &e.done, e.die, best);
  }
  int winner[4];
  get_winner(c, e.board, winner);
  return winner[gidx(e0->cur, 4)] ? 1.0f : -1.0f;
}

static void cls_to_emb(const madn_cfg *c, const cls_env *e, float *f) {
  int k = 0;
  for (int i = 0; i < c->total; ++i) f[k++] = e->board[i];
  f[k++] = e->cur;
  for (int i = 0; i < c->n * 4; ++i) f[k++] = e->pins[i];
  f[k++] = e->reward; f[k++] = e->done; f[k++] = e->die;
  f[k++] = (float)(e->key[0] & 0xFFFFu); f[k++] = (float)(e->key[0] >> 16);
  f[k++] = (float)(e->key[1] & 0xFFFFu); f[k++] = (float)(e->key[1] >> 16);
}
static void cls_from_emb(const madn_cfg *c, cls_env *e, const float *f) {
  int k = 0;
  memset(e, 0, sizeof(*e));
  for (int i = 0; i < c->total; ++i) e->board[i] = (int8_t)f[k++];
  e->cur = (int8_t)f[k++];
  for (int i = 0; i < c->n * 4; ++i) e->pins[i] = (int8_t)f[k++];
  e->reward = (int8_t)f[k++]; e->done = (uint8_t)f[k++]; e->die = (int8_t)f[k++];
  e->key[0] = (uint32_t)f[k] | ((uint32_t)f[k + 1] << 16);
  e->key[1] = (uint32_t)f[k + 2] | ((uint32_t)f[k + 3] << 16);
}
static void cls_ld(const madn_cfg *c, cls_env *e, int64_t g, const int8_t *board, const int8_t *cur, const int8_t *pins,
                   const int8_t *reward, const uint8_t *done, const int8_t *die, const uint32_t *key) {
  memset(e, 0, sizeof(*e));
  memcpy(e->board, board + g * c->total, c->total);
  e->cur = cur[g]; e->reward = reward[g]; e->done = done[g]; e->die = die[g];
  memcpy(e->pins, pins + g * c->n * 4, c->n * 4);
  e->key[0] = key[2 * g]; e->key[1] = key[2 * g + 1];
}

int orc_madn_cls_embed_dim(CFG_ARGS) { MAKE_CFG; return cfg.total > 64 ? -1 : cfg.total + 4 * cfg.n + 8; }

#define CLS_LEAVES const int8_t *board, const int8_t *current_player, const int8_t *pins, const int8_t *reward, const uint8_t *done, \
                   const int8_t *die, const uint32_t *key

/* policy_function (:571-583) -> f32 [n, 4] */
int orc_madn_cls_policy_function(CFG_ARGS, int64_t n, CLS_LEAVES, float *logits) {
  MAKE_CFG;
  if (cfg.total > 64) return -1;
  for (int64_t g = 0; g < n; ++g) {
    cls_env e; cls_ld(&cfg, &e, g, board, current_player, pins, reward, done, die, key);
    cls_policy_one(&cfg, &e, logits + 4 * g);
  }
  return 0;
}

/* root_fn (:690-714) */
int orc_madn_cls_root_fn(CFG_ARGS, int64_t n, CLS_LEAVES, const uint32_t *keys, float *prior, float *value, float *emb) {
  MAKE_CFG;
  if (cfg.total > 64) return -1;
  const int E = cfg.total + 4 * cfg.n + 8;
  for (int64_t g = 0; g < n; ++g) {
    cls_env e; cls_ld(&cfg, &e, g, board, current_player, pins, reward, done, die, key);
    cls_policy_one(&cfg, &e, prior + 4 * g);
    value[g] = cls_rollout_one(&cfg, &e, keys + 2 * g);
    cls_to_emb(&cfg, &e, emb + (int64_t)E * g);
  }
  return 0;
}

/* recurrent_fn (:657-688), the decision node: no_step if nothing is legal, else env_step(action); chance_logits = log(1/6) x 6,
 * afterstate_value = rollout(afterstate) */
int orc_madn_cls_decision_recurrent_fn(CFG_ARGS, int64_t n, const uint32_t *keys, const int32_t *action, const float *emb_in,
                                       float *chance_logits, float *afterstate_value, float *emb_out) {
  MAKE_CFG;
  if (cfg.total > 64) return -1;
  const int E = cfg.total + 4 * cfg.n + 8;
  const float l6 = f_log_d((float)(1.0 / 6.0));
  for (int64_t g = 0; g < n; ++g) {
    cls_env e; cls_from_emb(&cfg, &e, emb_in + (int64_t)E * g);
    uint8_t mask[4];
    orc_madn_cls_valid_action_one(&cfg, e.board, e.cur, e.pins, e.die, mask);
    if (!(mask[0] | mask[1] | mask[2] | mask[3])) e.cur = (int8_t)fmod_(e.cur + 1, cfg.n);
    else orc_madn_cls_step_one(&cfg, e.board, &e.cur, e.pins, &e.reward, &e.done, e.die, action[g]);
    for (int k = 0; k < 6; ++k) chance_logits[6 * g + k] = l6;
    afterstate_value[g] = cls_rollout_one(&cfg, &e, keys + 2 * g);
    cls_to_emb(&cfg, &e, emb_out + (int64_t)E * g);
  }
  return 0;
}

/* recurrent_chance_fn (:624-655): set_die(afterstate, outcome + 1); action_logits = valid_action as 0 / 1, value = rollout,
 * reward = env.reward, discount = done ? 0 : 1 */
int orc_madn_cls_chance_recurrent_fn(CFG_ARGS, int64_t n, const uint32_t *keys, const int32_t *outcome, const float *emb_in,
                                     float *action_logits, float *value, float *reward, float *discount, float *emb_out) {
  MAKE_CFG;
  if (cfg.total > 64) return -1;
  const int E = cfg.total + 4 * cfg.n + 8;
  for (int64_t g = 0; g < n; ++g) {
    cls_env e; cls_from_emb(&cfg, &e, emb_in + (int64_t)E * g);
    e.die = (int8_t)(outcome[g] + 1);
    uint8_t mask[4];
    orc_madn_cls_valid_action_one(&cfg, e.board, e.cur, e.pins, e.die, mask);
    for (int a = 0; a < 4; ++a) action_logits[4 * g + a] = mask[a] ? 1.0f : 0.0f;
    value[g] = cls_rollout_one(&cfg, &e, keys + 2 * g);
    reward[g] = (float)e.reward;
    discount[g] = e.done ? 0.0f : 1.0f;
    cls_to_emb(&cfg, &e, emb_out + (int64_t)E * g);
  }
  return 0;
}
