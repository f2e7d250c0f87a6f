/* ttt_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Scalar restatement of the reference's two TicTacToe environments and their true-env mctx callbacks:
 *   /root/reference/TicTacToe/TicTacToe.py:19-117   (variant 0: classic)
 *   /root/reference/TicTacToe/TicTacToeV2.py:22-140 (variant 1: only each player's last 3 moves persist), including the two
 *   operator-precedence quirks of its env_step (:66, :70 — SURVEY Appendix A.7), reproduced literally.
 * categorical() is argmax(logits + gumbel) in float32 with log rounded from double (the float contract of DESIGN.md).
 *
 * PARITY STATUS: pinned by trajectories / callback outputs the reference's own code produced on oracle/jaxshim
 * (tests/golden/ttt_reference.npz); the reference has no TicTacToe test of its own.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "jaxrand_oracle.h"

typedef struct { int8_t board[9]; int8_t cur, reward; uint8_t done; int8_t memory[6]; } ttt;

static float f_log(float x) { return (float)log((double)x); }

/* get_winner — TicTacToeV2.py:22-36 */
static int get_winner(const int8_t *b) {
  static const int L[8][3] = {{0,1,2},{3,4,5},{6,7,8},{0,3,6},{1,4,7},{2,5,8},{0,4,8},{2,4,6}};
  int w = 0, neg = 0;
  for (int i = 0; i < 8; ++i) { int s = b[L[i][0]] + b[L[i][1]] + b[L[i][2]]; if (s == 3) w = 1; if (s == -3) neg = 1; }
  return neg ? -1 : w;
}
static int pyfloordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
static int pymod(int a, int b) { int r = a % b; if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }
static int wrapidx(int i, int n) { if (i < 0) i += n; if (i < 0) i = 0; if (i > n - 1) i = n - 1; return i; }

/* env_step — TicTacToe.py:42-59 (variant 0), TicTacToeV2.py:46-79 (variant 1) */
void orc_ttt_step_one(int variant, ttt *e, int action_in) {
  int action = (int8_t)action_in;
  int row = wrapidx(pyfloordiv(action, 3), 3), col = wrapidx(pymod(action, 3), 3);
  int cell = row * 3 + col;
  int invalid = e->board[cell] != 0;
  int8_t board[9];
  memcpy(board, e->board, 9);
  int keep = e->done || invalid;
  if (!keep) board[cell] = e->cur;
  if (variant == 1) {
    int player = e->cur < 0;
    int8_t *m = e->memory + 3 * player;
    int removed = m[0]; /* roll(-1) then read the last slot */
    /* (:66) `env.done | invalid_move | removed_action == -1` parses as ((done | invalid) | removed) == -1 in int8 */
    int8_t lhs = (int8_t)((int8_t)keep | (int8_t)removed);
    int keep_removed = lhs == -1;
    int rr = wrapidx(pyfloordiv(removed, 3), 3), rc = wrapidx(pymod(removed, 3), 3);
    if (!keep_removed) board[rr * 3 + rc] = 0;
    if (!keep) { m[0] = m[1]; m[1] = m[2]; m[2] = (int8_t)action; }
  }
  int reward = e->done ? 0 : (invalid ? -1 : get_winner(board) * e->cur);
  int full = 1;
  for (int k = 0; k < 9; ++k) full &= board[k] != 0;
  int done;
  if (variant == 1) {
    /* (:70) `env.done | reward != 0 | invalid_move | all(board != 0)` parses as (done | reward) != (0 | invalid | full) */
    int8_t lhs = (int8_t)((int8_t)e->done | (int8_t)reward);
    int rhs = 0 | invalid | full;
    done = lhs != rhs;
  } else {
    done = e->done || reward != 0 || invalid || full;
  }
  memcpy(e->board, board, 9);
  e->cur = (int8_t)(done ? e->cur : -e->cur);
  e->done = (uint8_t)done;
  e->reward = (int8_t)reward;
}

/* policy_function — TicTacToeV2.py:96-102: 100 legal + 200 opponent-winning + 300 own-winning */
void orc_ttt_policy_one(int variant, const ttt *e, float logits[9]) {
  for (int a = 0; a < 9; ++a) {
    float v = (!e->done && e->board[a] == 0) ? 100.0f : 0.0f;
    for (int side = 0; side < 2; ++side) {
      ttt t = *e;
      t.cur = (int8_t)(side == 0 ? -e->cur : e->cur);
      orc_ttt_step_one(variant, &t, a);
      if (t.reward == 1) v = v + (side == 0 ? 200.0f : 300.0f);
    }
    logits[a] = v;
  }
}

static int categorical9(const uint32_t key[2], const float logits[9]) {
  int best = 0; float bv = 0;
  for (int a = 0; a < 9; ++a) {
    float u = orc_uniform_i(key, (uint32_t)a, 1.17549435e-38f, 1.0f);
    float v = -f_log(-f_log(u)) + logits[a];
    if (a == 0 || v > bv) { bv = v; best = a; }
  }
  return best;
}

/* rollout — TicTacToeV2.py:104-116 */
float orc_ttt_rollout_one(int variant, const ttt *e0, const uint32_t key_in[2]) {
  ttt e = *e0;
  uint32_t key[2] = {key_in[0], key_in[1]};
  for (int it = 0; it < 100000 && !e.done; ++it) {
    uint32_t nk[2], sub[2];
    orc_split_i(key, 0, nk); orc_split_i(key, 1, sub);
    key[0] = nk[0]; key[1] = nk[1];
    float lg[9];
    orc_ttt_policy_one(variant, &e, lg);
    orc_ttt_step_one(variant, &e, categorical9(sub, lg));
  }
  return (float)(int8_t)(e.reward * e.cur * e0->cur);
}

/* batched SoA: board i8[n,9], cur i8[n], reward i8[n], done u8[n], memory i8[n,6] */
static void ld(ttt *e, int64_t g, const int8_t *board, const int8_t *cur, const int8_t *reward, const uint8_t *done, const int8_t *mem) {
  memcpy(e->board, board + 9 * g, 9); e->cur = cur[g]; e->reward = reward[g]; e->done = done[g]; memcpy(e->memory, mem + 6 * g, 6);
}
static void st(const ttt *e, int64_t g, int8_t *board, int8_t *cur, int8_t *reward, uint8_t *done, int8_t *mem) {
  memcpy(board + 9 * g, e->board, 9); cur[g] = e->cur; reward[g] = e->reward; done[g] = e->done; memcpy(mem + 6 * g, e->memory, 6);
}
static void to_emb(const ttt *e, float *f) { for (int k = 0; k < 9; ++k) f[k] = e->board[k]; f[9] = e->cur; f[10] = e->reward; f[11] = e->done; for (int k = 0; k < 6; ++k) f[12 + k] = e->memory[k]; }
static void from_emb(ttt *e, const float *f) { for (int k = 0; k < 9; ++k) e->board[k] = (int8_t)f[k]; e->cur = (int8_t)f[9]; e->reward = (int8_t)f[10]; e->done = (uint8_t)f[11]; for (int k = 0; k < 6; ++k) e->memory[k] = (int8_t)f[12 + k]; }

void orc_ttt_reset(int64_t n, int8_t *board, int8_t *cur, int8_t *reward, uint8_t *done, int8_t *mem) {
  for (int64_t g = 0; g < n; ++g) { memset(board + 9 * g, 0, 9); cur[g] = 1; reward[g] = 0; done[g] = 0; memset(mem + 6 * g, -1, 6); }
}
void orc_ttt_step(int variant, int64_t n, int8_t *board, int8_t *cur, int8_t *reward, uint8_t *done, int8_t *mem, const int8_t *action) {
  for (int64_t g = 0; g < n; ++g) { ttt e; ld(&e, g, board, cur, reward, done, mem); orc_ttt_step_one(variant, &e, action[g]); st(&e, g, board, cur, reward, done, mem); }
}
void orc_ttt_policy(int variant, int64_t n, const int8_t *board, const int8_t *cur, const int8_t *reward, const uint8_t *done, const int8_t *mem, float *logits) {
  for (int64_t g = 0; g < n; ++g) { ttt e; ld(&e, g, board, cur, reward, done, mem); orc_ttt_policy_one(variant, &e, logits + 9 * g); }
}
/* root_fn — TicTacToeV2.py:121-126 */
void orc_ttt_root_fn(int variant, int64_t n, const int8_t *board, const int8_t *cur, const int8_t *reward, const uint8_t *done, const int8_t *mem,
                     const uint32_t *keys, float *prior, float *value, float *emb) {
  for (int64_t g = 0; g < n; ++g) {
    ttt e; ld(&e, g, board, cur, reward, done, mem);
    orc_ttt_policy_one(variant, &e, prior + 9 * g);
    value[g] = orc_ttt_rollout_one(variant, &e, keys + 2 * g);
    to_emb(&e, emb + 18 * g);
  }
}
/* recurrent_fn — TicTacToeV2.py:128-140 */
void orc_ttt_recurrent_fn(int variant, int64_t n, const uint32_t *keys, const int32_t *action, const float *emb_in, float *prior, float *value,
                          float *reward, float *discount, float *emb_out) {
  for (int64_t g = 0; g < n; ++g) {
    ttt e; from_emb(&e, emb_in + 18 * g);
    orc_ttt_step_one(variant, &e, (int8_t)action[g]);
    reward[g] = (float)e.reward;
    discount[g] = e.done ? 0.0f : -1.0f;
    orc_ttt_policy_one(variant, &e, prior + 9 * g);
    value[g] = e.done ? 0.0f : orc_ttt_rollout_one(variant, &e, keys + 2 * g);
    to_emb(&e, emb_out + 18 * g);
  }
}
