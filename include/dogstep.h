/* dogstep.h — C ABI of libdogstep.so: the B200-native self-play hot path behind the
 * Exploring-MuZero-on-DOG Python API.
 *
 * The reference has no FFI of its own: its boundary is a set of pure Python/JAX functions on
 * flax.struct pytrees that are vmapped over games.  Each entry point
 * below replaces one of those functions *after vmap*, i.e. it sees the batched leaves of the
 * pytree as structure-of-arrays device buffers with a leading game axis `n`.  Every entry point
 * cites the reference function it stands in for.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says `host_`; the caller owns and
 *     pre-allocates every buffer; state is updated in place (jax.ffi: input_output_aliases)
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised
 *   - return 0 on success, DOGSTEP_ERR_* otherwise; an invalid *game* action is not an error
 *     (the reference answers reward -1 and still passes the turn)
 *   - rule dicts become one uint32 bitmask (dogstep_rules.h)
 *   - re-entrant, no global mutable state
 */
#ifndef DOGSTEP_H
#define DOGSTEP_H
#include <stdint.h>
#include "dogstep_rules.h"

#ifdef __cplusplus
extern "C" {
#endif

int dogstep_version(void);
/* last CUDA error string seen by this thread ("" if none) */
const char* dogstep_last_error(void);

/* ---------------------------------------------------------------- board geometry / rule set
 * Static (pytree_node=False) fields of the reference envs: num_players, board_size, rules
 * (MADN/deterministic_madn.py:32,38-40), plus the `layout` argument of env_reset (:45,70-78).
 * board_size = 4*distance, total_board_size = board_size + 16. */
typedef struct {
  int32_t num_players; /* 2..4 */
  int32_t layout_mask; /* bit i = seat i present; replaced by the first n seats exactly like :70-74 */
  int32_t distance;    /* cells between two starts; 1..12 (total_board_size <= 64) */
  uint32_t rules;      /* DOGSTEP_RULE_* */
} dogstep_madn_cfg;

/* ---------------------------------------------------------------- deterministic MADN state
 * Batched leaves of `deterministic_MADN` (MADN/deterministic_madn.py:24-40).
 * start/target/goal are pure functions of the cfg and are not stored per game. */
typedef struct {
  int8_t* board;          /* [n, total_board_size]  -1 empty, else owning player */
  int8_t* current_player; /* [n] */
  int8_t* pins;           /* [n, num_players, 4]    -1 home, 0..bs-1 ring, bs.. goal lanes */
  int8_t* reward;         /* [n] */
  uint8_t* done;          /* [n] bool */
  int8_t* action_set;     /* [n, num_players, 6]    remaining copies of move cards 1..6 */
  uint32_t* key;          /* [n, 2] raw threefry key */
} dogstep_madn_det_state;

/* env_reset vmapped over seeds — MADN/deterministic_madn.py:42-120 (game_agent.py:24-44).
 * starting_player outside [0, num_players) draws randint(subkey, (), 0, num_players) (:62). */
int dogstep_madn_det_reset(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                           const int32_t* seeds /*[n]*/, int32_t starting_player, void* stream);

/* valid_action — MADN/deterministic_madn.py:299-393.  mask: uint8 [n, 4, 6] (0/1). */
int dogstep_madn_det_valid_action(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                  uint8_t* mask, void* stream);

/* env_step — MADN/deterministic_madn.py:170-257.  action: int8 [n, 2] = [pin, move].
 * reward/done outputs may be NULL (they are also written into the state leaves). */
int dogstep_madn_det_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                          const int8_t* action, int8_t* reward, uint8_t* done, void* stream);

/* no_step — MADN/deterministic_madn.py:283-297 (returned reward is the constant 0). */
int dogstep_madn_det_no_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                             int8_t* reward, uint8_t* done, void* stream);

/* set_pins_on_board — MADN/deterministic_madn.py:259-271.  pins [n, P, 4] -> board [n, total]. */
int dogstep_madn_set_pins_on_board(const int8_t* pins, int8_t* board, int64_t n, const dogstep_madn_cfg* cfg,
                                   void* stream);

/* True-env mctx callbacks of the deterministic game — MADN/deterministic_madn.py:480-590, the functions
 * MADN/simulate_deterministicMADN.py:12-35 hands to mctx.gumbel_muzero_policy.  The env embedding is the state as floats,
 * E = total + 10 * num_players + 3: board[total], current_player, pins[4 P], reward, done, action_set[6 P]
 * (dogstep_madn_det_embed_dim).  Actions are indices 0..23 (map_action :469-479: pin = a // 6, move = a % 6 + 1, both int8).
 *   policy_function (:495-507): logits f32 [n,24] = 100 * valid_action + 200 * winning_action, winning_action (:480-493) =
 *     "env_step on a copy returns reward 1";
 *   root_fn (:551-566): prior = policy_function(env), value = rollout(env, key) (:509-541: at most 300 steps of
 *     key, subkey = split(key); no_step if nothing is legal, else categorical(subkey, policy_function) -> env_step), embedding = env.
 *     The reference's value is a float32[4] of four equal entries (its `winner == -1` test compares a bool array): +1 if the
 *     root player('s team) has won when the rollout stops, -1 otherwise — also when the cap ends it; the scalar is written;
 *   recurrent_fn (:568-590): env_step(embedding, map_action(action)); reward, discount = done ? 0 : -1, prior of the successor,
 *     value = done ? 0 : rollout.  embedding_out may alias embedding_in.
 * keys u32 [n,2].  One warp per game, generic rules (any geometry / rule set). */
int dogstep_madn_det_embed_dim(const dogstep_madn_cfg* cfg);
int dogstep_madn_det_policy_function(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, float* logits,
                                     void* stream);
int dogstep_madn_det_root_fn(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys,
                             float* prior_logits, float* value, float* embedding, void* stream);
int dogstep_madn_det_recurrent_fn(int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys, const int32_t* action,
                                  const float* embedding_in, float* prior_logits, float* value, float* reward, float* discount,
                                  float* embedding_out, void* stream);

/* encode_board — MADN/deterministic_madn.py:395-438.  obs: int8 [n, 8*P+2, total]. */
int dogstep_madn_det_encode_board(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                  int8_t* obs, void* stream);

/* One lockstep iteration of the random-legal-policy driver, fused:
 * MuZero_det_MADN/evaluate_agent.py:733-930 with do_random (:772-776):
 *   game j not done: mask = valid_action; any(mask) ? env_step(map_action(categorical(key_j, where(mask,0,-1e9))))
 *                                                  : no_step
 *   key_j = split(rng, N+1)[game_offset + j + 1]  (:741);  the carried key is split(rng, N+1)[0].
 * host_rng_key: uint32[2] on the HOST (the loop-carried key); active_count (device int64, may be
 * NULL) is incremented by the number of games that were not done. */
int dogstep_madn_det_random_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                 const uint32_t* host_rng_key, int64_t game_offset,
                                 unsigned long long* active_count, void* stream);

/* The whole while_loop of that driver as ONE persistent launch: every game is advanced until it is
 * done or `max_steps` lockstep iterations have passed (cap 2000 at evaluate_agent.py:918).
 * game_len: int32 [n] iterations in which game j was still active (may be NULL);
 * total_steps: device int64 accumulator, += sum(game_len) (may be NULL). */
int dogstep_madn_det_play_random(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                 const uint32_t* host_rng_key, int64_t game_offset, int32_t max_steps,
                                 int32_t* game_len, unsigned long long* total_steps, void* stream);

/* ---------------------------------------------------------------- classic (dice) MADN state
 * Batched leaves of `classic_MADN` (MADN/classic_madn.py:33-49). */
typedef struct {
  int8_t* board;          /* [n, total_board_size] */
  int8_t* current_player; /* [n] */
  int8_t* pins;           /* [n, num_players, 4] */
  int8_t* reward;         /* [n] */
  uint8_t* done;          /* [n] */
  int8_t* die;            /* [n] */
  uint32_t* key;          /* [n, 2] */
} dogstep_madn_cls_state;

/* env_reset — MADN/classic_madn.py:51-131 */
int dogstep_madn_cls_reset(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                           const int32_t* seeds, int32_t starting_player, void* stream);
/* throw_die — MADN/classic_madn.py:230-242 (split env.key, choice with dice_probabilities) */
int dogstep_madn_cls_throw_die(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, void* stream);
/* dice_probabilities — MADN/classic_madn.py:208-228.  p: float32 [n, 6] */
int dogstep_madn_cls_dice_probabilities(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                        float* p, void* stream);
/* valid_action — MADN/classic_madn.py:367-461.  mask: uint8 [n, 4] */
int dogstep_madn_cls_valid_action(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                  uint8_t* mask, void* stream);
/* env_step — MADN/classic_madn.py:257-337.  action: int8 [n] pin index */
int dogstep_madn_cls_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                          const int8_t* action, int8_t* reward, uint8_t* done, void* stream);
/* no_step — MADN/classic_madn.py:353-365 */
int dogstep_madn_cls_no_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                             int8_t* reward, uint8_t* done, void* stream);
/* encode_board — MADN/classic_madn.py:463-497.  obs: int8 [n, 2*P+3, total] */
int dogstep_madn_cls_encode_board(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg,
                                  int8_t* obs, void* stream);

/* True-env mctx callbacks of the dice game — MADN/classic_madn.py:541-714, for mctx.stochastic_muzero_policy
 * on the true env.  The reference as it stands raises before any of them returns: winning_action (:551-565) builds its scratch
 * copy without the dataclass's `key` field; these entry points compute what the reference's function bodies compute once that
 * constructor call goes through (pinned by tests/golden/madn_cls_reference_trueenv.npz, see its generator).
 * Embedding E = total + 4 * num_players + 8: board[total], current_player, pins[4 P], reward, done, die, env.key as four 16-bit
 * halves (k0 lo, k0 hi, k1 lo, k1 hi).  Actions are pins 0..3, chance outcomes 0..5 (die - 1).  keys u32 [n,2].
 *   policy_function (:571-583): logits f32 [n,4] = 100 * valid_action + 200 * winning_action;
 *   root_fn (:690-714): prior = policy_function(env), value = rollout(env, key) (:585-616: at most 300 steps of key, subkey =
 *     split(key); env = throw_die(env) on the ENV's key chain; no_step if nothing is legal, else categorical(subkey,
 *     policy_function) -> env_step; value +-1 as in the deterministic game), embedding = env;
 *   recurrent_fn (:657-688, decision node): afterstate = no_step(env) if nothing is legal else env_step(env, action);
 *     chance_logits f32 [n,6] = log(1/6), afterstate_value = rollout(afterstate, key);
 *   recurrent_chance_fn (:624-655): env = set_die(afterstate, outcome + 1); action_logits f32 [n,4] = valid_action as 0 / 1,
 *     value = rollout(env, key), reward = env.reward, discount = done ? 0 : 1.
 * embedding_out may alias embedding_in. */
int dogstep_madn_cls_embed_dim(const dogstep_madn_cfg* cfg);
int dogstep_madn_cls_policy_function(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, float* logits,
                                     void* stream);
int dogstep_madn_cls_root_fn(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys,
                             float* prior_logits, float* value, float* embedding, void* stream);
int dogstep_madn_cls_decision_recurrent_fn(int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys, const int32_t* action,
                                           const float* embedding_in, float* chance_logits, float* afterstate_value,
                                           float* embedding_out, void* stream);
int dogstep_madn_cls_chance_recurrent_fn(int64_t n, const dogstep_madn_cfg* cfg, const uint32_t* keys, const int32_t* chance_outcome,
                                         const float* embedding_in, float* action_logits, float* value, float* reward,
                                         float* discount, float* embedding_out, void* stream);


/* ---------------------------------------------------------------- DOG (2v2 card game)
 * Batched leaves of the `DOG` dataclass (DOG/dog.py:31-56).  num_cards is 14: the disable_* rule bits
 * are rejected with DOGSTEP_ERR_UNSUPPORTED (the reference itself is only self-consistent with all cards).
 * Action space: [0,396) joker copies, [396,792) real cards, [792,806) swap-phase card choice, for
 * total_board_size 56 (get_play_action_size, DOG/dog.py:58-59). */
typedef dogstep_madn_cfg dogstep_dog_cfg;
typedef struct {
  int8_t* board;          /* [n, total_board_size] */
  int8_t* current_player; /* [n] */
  int32_t* pins;          /* [n, num_players, 4]  (int32 in the reference) */
  int8_t* reward;         /* [n] */
  uint8_t* done;          /* [n] */
  int8_t* deck;           /* [n, 14] remaining copies per card type */
  int8_t* hands;          /* [n, num_players, 14] */
  int8_t* swap_choices;   /* [n, 4] */
  int8_t* round_starter;  /* [n] */
  int8_t* phase;          /* [n] 0 play, 1 partner card swap */
  uint32_t* key;          /* [n, 2] */
  int8_t* hand_size;      /* [n] size of the NEXT deal (6,5,4,3,2,6,...) */
} dogstep_dog_state;

/* get_play_action_size(env) + 14 — DOG/dog.py:58-59, 693-711 (negative = error code) */
int dogstep_dog_num_actions(const dogstep_dog_cfg* cfg);
/* env_reset (+ first distribute_cards) — DOG/dog.py:83-181 */
int dogstep_dog_reset(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* seeds,
                      int32_t starting_player, void* stream);
/* valid_actions — DOG/dog.py:693-711 (valid_step_actions :618-691).  mask: uint8 [n, num_actions] */
int dogstep_dog_valid_actions(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, uint8_t* mask, void* stream);
/* Observation encoder for the DOG env — NOT a reference function: DOG/dog.py:1264-1272 is a TODO and the DOG networks are stubs
 * (MuZero_DOG/muzero_dog.py:85-99); SURVEY 8(f).4 calls it design work.  Same conventions as the MADN encoders the reference has
 * (MADN/deterministic_madn.py:395-438): mover's frame, scalar facts broadcast over planes, only what the seat to move may know.
 * obs int8 [n, dogstep_dog_obs_planes(cfg), total_board_size]; planes = 8 + 3 * num_players + 14 (34 for four players), laid out as
 * documented at k_dog_encode_board (csrc/dog_kernels.cu). */
int dogstep_dog_obs_planes(const dogstep_dog_cfg* cfg);
int dogstep_dog_encode_board(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, int8_t* obs, void* stream);
/* env_step — DOG/dog.py:1117-1131.  action: int32 [n] */
int dogstep_dog_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* action, int8_t* reward,
                     uint8_t* done, void* stream);
/* no_step — DOG/dog.py:713-752 */
int dogstep_dog_no_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, int8_t* reward, uint8_t* done,
                        void* stream);
/* distribute_cards — DOG/dog.py:201-298 */
int dogstep_dog_distribute_cards(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, void* stream);
/* step_normal_move / step_neg_move / step_swap / step_hot_7 — DOG/dog.py:790, 861, 755, 913 as the reference's
 * tests call them: kind int32[n] (0,1,2,3), args int32[n,4] = (pin, move) | (pin, move) | (pin, pos) | dist[4];
 * returns (board int8[n,total], pins int32[n,P,4], reward, done) and leaves the env untouched. */
int dogstep_dog_substep(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* kind,
                        const int32_t* args, int8_t* board_out, int32_t* pins_out, int8_t* reward, uint8_t* done, void* stream);
/* one fused lockstep iteration of the random-legal-policy driver over DOG's 806 actions (same key
 * derivation as dogstep_madn_det_random_step) */
int dogstep_dog_random_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const uint32_t* host_rng_key,
                            int64_t game_offset, unsigned long long* active_count, void* stream);
/* the whole random-policy loop as one persistent launch (cap MuZero_DOG/evaluate_agent.py:518).  Games are handed to the
 * warps through a device-side queue counter; the library keeps one 256-byte scratch allocation for these counters (made at
 * the first call, never freed) — its only allocation. */
int dogstep_dog_play_random(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const uint32_t* host_rng_key,
                            int64_t game_offset, int32_t max_steps, int32_t* game_len, unsigned long long* total_steps,
                            void* stream);

/* ---------------------------------------------------------------- MCTS (mctx 0.0.6 search, per-game trees)
 * Stand-in for the search calls the reference makes through the third-party `mctx` package:
 *   mctx.gumbel_muzero_policy      MuZero_det_MADN/muzero_deterministic_madn.py:673-684 (run_muzero_mcts :663-704)
 *   mctx.stochastic_muzero_policy  MuZero_Classic_MADN/muzero_classic_madn.py:488-501 (run_stochastic_muzero_mcts :464-517)
 *   mctx.muzero_policy / gumbel    TicTacToe/mcts.py:13-22, 29-37
 * The networks (root_fn / recurrent_fn) stay with the caller: one simulation is
 *   dogstep_mcts_select  ->  caller evaluates recurrent_fn on the gathered parent embeddings  ->  dogstep_mcts_expand
 * Tree buffers have mctx's `Tree` layout (batch, node, action) so `policy_output.search_tree` can alias them.
 * Float contract: IEEE add/mul/div/sqrt without FMA contraction, exp/log correctly rounded from double, sums in the
 * fixed order "lane l accumulates a = l, l+32, ...; then butterfly 16,8,4,2,1" (MCTS numerics). */
#define DOGSTEP_MCTS_MUZERO      0 /* mctx.muzero_policy: PUCT + tie-break noise, sample from visit counts */
#define DOGSTEP_MCTS_GUMBEL      1 /* mctx.gumbel_muzero_policy: sequential halving at the root */
#define DOGSTEP_MCTS_STOCHASTIC  2 /* mctx.stochastic_muzero_policy: decision / chance levels alternate */
#define DOGSTEP_Q_BY_MIN_MAX             0 /* qtransform_by_min_max(min_value, max_value) */
#define DOGSTEP_Q_BY_PARENT_AND_SIBLINGS 1 /* qtransform_by_parent_and_siblings(epsilon) */
#define DOGSTEP_Q_COMPLETED_BY_MIX_VALUE 2 /* qtransform_completed_by_mix_value(value_scale, maxvisit_init, epsilon) */

typedef struct {
  int32_t policy, qtransform;
  int32_t num_simulations, max_depth;
  int32_t num_actions;  /* A: decision actions */
  int32_t num_chance;   /* C: chance outcomes (stochastic only, else 0); children arrays are A' = A + C wide */
  int32_t embed_dim;    /* floats per stored node embedding */
  int32_t max_num_considered_actions; /* gumbel */
  float q_min, q_max;                 /* by_min_max */
  float value_scale, maxvisit_init;   /* completed_by_mix_value */
  float epsilon;                      /* 1e-8 in mctx */
  float pb_c_init, pb_c_base;         /* 1.25, 19652 */
  float dirichlet_fraction;           /* muzero / stochastic root noise mix; the Dirichlet SAMPLE is an input */
  float temperature;                  /* final sampling temperature (muzero / stochastic) */
  float gumbel_scale;                 /* gumbel */
  int32_t state_embed_dim;            /* stochastic: floats per row of the `embedding` argument of expand (0 = embed_dim) */
  int32_t afterstate_embed_dim;       /* stochastic: floats per row of `afterstate_embedding` (0 = embed_dim).  mctx pads the
                                       * narrower of the two embeddings to the wider one in its tree; with these widths the
                                       * caller hands over the unpadded rows and expand zero-fills, instead of a pad copy per
                                       * simulation on the caller's side */
} dogstep_mcts_cfg;

typedef struct {
  int32_t* node_visits;           /* [n, N]      N = num_simulations + 1 */
  float* raw_values;              /* [n, N] */
  float* node_values;             /* [n, N] */
  int32_t* parents;               /* [n, N]      -1 = no parent */
  int32_t* action_from_parent;    /* [n, N] */
  int32_t* children_index;        /* [n, N, A']  -1 = unvisited */
  float* children_prior_logits;   /* [n, N, A'] */
  int32_t* children_visits;       /* [n, N, A'] */
  float* children_rewards;        /* [n, N, A'] */
  float* children_discounts;      /* [n, N, A'] */
  float* children_values;         /* [n, N, A'] */
  float* embeddings;              /* [n, N, E] */
  uint8_t* is_decision;           /* [n, N]  stochastic only (may be NULL otherwise) */
  uint8_t* root_invalid_actions;  /* [n, A'] */
  float* root_gumbel;             /* [n, A'] gumbel only (may be NULL otherwise) */
  uint32_t* search_key;           /* [n, 2]  key chain of search(): rng, sim, exp = split(rng, 3) per simulation */
  uint32_t* policy_key;           /* [n, 2]  key of the final categorical draw (muzero / stochastic) */
  int32_t* path;                  /* [n, 65] optional scratch (NULL allowed): select records the descent (edge count, then
                                   * (node, action) for the first 32 edges) and expand uses it to fetch every level of the
                                   * backup at once instead of walking parent pointers; results are identical */
  uint32_t* select_aux;           /* [n, N + 1, 52] optional scratch (NULL allowed): per-node select cache for wide Gumbel trees
                                   * (32 < A' <= 832, completed_by_mix_value; DOG's 806 actions) — bitmap of the children with
                                   * visits, their visit sum / maximum, the node's prior softmax statistics and its eight
                                   * largest prior logits with their indices — so that a level reads the few visited children
                                   * (and the prior row only when its decision cannot be proven without it) instead of five
                                   * dense rows.
                                   * Written by init / expand, read by select; results are identical with and without it.
                                   * If given, every call on this tree (init, select, expand, expand_select) must get it. */
  int32_t* select_action_decision; /* [n] optional output of select (stochastic): min(action, A - 1), the index the
                                   * decision_recurrent_fn is evaluated with */
  int32_t* select_action_chance;  /* [n] optional output of select (stochastic): clamp(action - A, 0, C - 1), the chance
                                   * outcome the chance_recurrent_fn is evaluated with (mctx evaluates both callbacks on every
                                   * simulation and picks by node type) */
} dogstep_mcts_tree;

/* policy prologue + instantiate_tree_from_root.  keys: uint32 [n,2] = the rng_key each game hands to the mctx policy.
 * root_prior_logits f32 [n,A], root_value f32 [n], root_embedding f32 [n,E], invalid_actions u8 [n,A] (NULL = none),
 * dirichlet_noise f32 [n,A] (NULL = no root noise; ignored by gumbel). */
int dogstep_mcts_init(const dogstep_mcts_tree* t, int64_t n, const dogstep_mcts_cfg* cfg, const uint32_t* keys,
                      const float* root_prior_logits, const float* root_value, const float* root_embedding,
                      const uint8_t* invalid_actions, const float* dirichlet_noise, void* stream);
/* Wide Gumbel trees that carry a select cache (select_aux != NULL, 32 < A' <= 832, qtransform_completed_by_mix_value) are SPARSE:
 * mcts_init writes only the root rows of the six [n, N, A'] child arrays (mctx's search.py instantiate_tree_from_root zero-fills
 * all of them: 16.9 GB per search at BASELINE config 5); every kernel of that path reads a child entry only where the cache's
 * bitmap marks a child with visits.  dogstep_mcts_is_sparse tells (1 / 0) whether a (tree, cfg) pair runs that way;
 * dogstep_mcts_materialize writes the defaults (index -1, zero visits / reward / discount / value; zero prior row and embedding
 * of nodes that were never created) so that the buffers equal the dense mctx.Tree — only needed by callers that read non-root
 * rows; a no-op for every other tree.  Idempotent. */
int dogstep_mcts_is_sparse(const dogstep_mcts_tree* t, const dogstep_mcts_cfg* cfg);
int dogstep_mcts_materialize(const dogstep_mcts_tree* t, int64_t n, const dogstep_mcts_cfg* cfg, void* stream);

/* simulate(): descend from the root to (parent, action) for simulation `sim`; gathers the parent's embedding.
 * parent_out i32 [n], action_out i32 [n], embedding_out f32 [n,E], is_decision_out u8 [n] (may be NULL),
 * expand_key_out u32 [n,2] (may be NULL): the `expand_key` mctx hands to recurrent_fn for this simulation. */
int dogstep_mcts_select(const dogstep_mcts_tree* t, int64_t n, const dogstep_mcts_cfg* cfg, int32_t sim, int32_t* parent_out,
                        int32_t* action_out, float* embedding_out, uint8_t* is_decision_out, uint32_t* expand_key_out,
                        void* stream);
/* expand() + backward() with the caller's recurrent_fn outputs for the (parent, action) pairs of `sim`:
 * prior_logits f32 [n,A], value/reward/discount f32 [n], embedding f32 [n,E].
 * Stochastic: the same five arrays are the chance_recurrent_fn outputs (action_logits, value, reward, discount,
 * state embedding) and chance_logits f32 [n,C], afterstate_value f32 [n], afterstate_embedding f32 [n,E] are the
 * decision_recurrent_fn outputs; the kernel picks per game by the parent's node type like mctx's where(). */
int dogstep_mcts_expand(const dogstep_mcts_tree* t, int64_t n, const dogstep_mcts_cfg* cfg, int32_t sim, const int32_t* parent,
                        const int32_t* action, const float* prior_logits, const float* value, const float* reward,
                        const float* discount, const float* embedding, const float* chance_logits,
                        const float* afterstate_value, const float* afterstate_embedding, void* stream);
/* expand + backup of simulation `sim` followed by the descent of simulation sim + 1 in ONE launch (the search loop always
 * issues them back to back; the caller's network sits between a select and the next expand).  `parent` / `action` are read
 * (the leaf edge chosen by the previous select) and then overwritten with the next leaf edge; embedding_out must not alias
 * the expand inputs.  Requires sim + 1 < num_simulations; results identical to expand(sim) then select(sim + 1). */
int dogstep_mcts_expand_select(const dogstep_mcts_tree* t, int64_t n, const dogstep_mcts_cfg* cfg, int32_t sim, int32_t* parent,
                               int32_t* action, const float* prior_logits, const float* value, const float* reward,
                               const float* discount, const float* embedding, const float* chance_logits,
                               const float* afterstate_value, const float* afterstate_embedding, float* embedding_out,
                               uint8_t* is_decision_out, uint32_t* expand_key_out, void* stream);
/* The exp of the float contract above, out[i] = (float)exp((double)x[i]), as the tree kernels evaluate it.  Exposed so
 * that the parity tests can check the device function itself against the oracle's libm expression (the two double
 * libraries may differ in the last bit; the float results differ only when that straddles a float rounding boundary,
 * about one argument in 2^29). */
int dogstep_exp_f32(const float* x, int64_t n, float* out, void* stream);
/* policy epilogue: action i32 [n], action_weights f32 [n,A], root_value f32 [n] (= search_tree.summary().value) */
int dogstep_mcts_policy_output(const dogstep_mcts_tree* t, int64_t n, const dogstep_mcts_cfg* cfg, int32_t* action,
                               float* action_weights, float* root_value, void* stream);

/* ---------------------------------------------------------------- replay ring buffer (per-GPU shard)
 * Stand-in for VectorizedReplayBuffer (MuZero_det_MADN/vec_replay_buffer.py:9-264) and
 * VectorizedReplayBufferStochastic (MuZero_Classic_MADN/vec_replay_buffer_stochastic.py), whose arrays live in host
 * NumPy in the reference; here they live in HBM next to the trajectory buffers self-play writes. */
typedef struct {
  int32_t capacity, max_episode_length, obs_size, action_dim;
  int32_t obs_is_int8;   /* 0: observations float32 (reference dtype), 1: int8 (4x smaller, same values) */
  int32_t stochastic;    /* 1: dice_outcomes / dice_distributions present, game_won = final_reward > 0 */
  void* observations;        /* [capacity, T, obs_size] */
  int32_t* actions;          /* [capacity, T] */
  int32_t* rewards;          /* [capacity, T]   class index 0/1/2 */
  float* root_values;        /* [capacity, T] */
  float* child_visits;       /* [capacity, T, action_dim] */
  float* masks;              /* [capacity, T] */
  int32_t* players;          /* [capacity, T] */
  int32_t* teams;            /* [capacity, T] */
  int32_t* discounts;        /* [capacity, T]   class index 0/1/2 */
  int32_t* episode_lengths;  /* [capacity] */
  int32_t* dice_outcomes;    /* [capacity, T]    stochastic only */
  float* dice_distributions; /* [capacity, T, 6] stochastic only */
} dogstep_replay_arrays;

/* save_games_from_buffers (:36-61): trajectories `traj` (same struct, capacity = number of games, episode_lengths = the
 * 'idx' leaf) are copied to ring slots slot[i] (device int32 [n_games], -1 = skip: length 0); only the first
 * length rows of a slot are overwritten, as in the reference.  The slot order is host logic (position/size). */
int dogstep_replay_save(const dogstep_replay_arrays* buf, const dogstep_replay_arrays* traj, int64_t n_games,
                        const int32_t* slot, void* stream);

/* The random part of sample_batch (:73-97): 75 % uniform (episode, t), 25 % windows aligned so that the episode's last
 * step falls at a random unroll position.  Drawn with threefry from host_key (the reference uses unseeded np.random,
 * so only the distribution is defined).  Outputs int32 [batch_size] each. */
int dogstep_replay_plan(const dogstep_replay_arrays* buf, int32_t size, int32_t batch_size, int32_t unroll_steps,
                        float terminal_ratio, const uint32_t* host_key, int32_t* ep_indices, int32_t* t_starts, void* stream);

typedef struct {
  float* observations;       /* [B, obs_size] */
  int32_t* actions;          /* [B, K-1] */
  int32_t* rewards;          /* [B, K-1] */
  float* policies;           /* [B, K, action_dim] */
  float* values;             /* [B, K] */
  float* masks;              /* [B, K] */
  float* target_values;      /* [B, K] */
  int32_t* discount_targets; /* [B, K-1] */
  int32_t* dice_outcomes;    /* [B, K-1]    stochastic only */
  float* dice_probs;         /* [B, K-1, 6] stochastic only */
} dogstep_replay_batch;

/* The deterministic part of sample_batch (:104-264): K = unroll_steps+1 window gathers, z / n-step bootstrap value
 * targets with the team / player perspective flip, clip, padding.  gamma_pow: device float64 [T+1] = GAMMA ** k
 * (host-computed so the powers are the reference's own libm values); arithmetic in float64, rounded to float32 once. */
int dogstep_replay_gather(const dogstep_replay_arrays* buf, int32_t batch_size, int32_t unroll_steps, int32_t td_steps,
                          int32_t bootstrap_value_target, const double* gamma_pow, const int32_t* ep_indices,
                          const int32_t* t_starts, const dogstep_replay_batch* out, void* stream);

/* Prioritised sampling — an EXTENSION: the reference samples uniformly with a terminal quota (vec_replay_buffer.py:73-97);
 * BASELINE's north star asks for prioritised sampling on the per-GPU shard.  Proportional prioritisation
 * P(e, t) = p[e, t] / sum(p) over fixed-point priorities: prio uint32 [capacity, T] in 2^-20 units, row_sum uint64 [capacity];
 * every sum is an exact integer, so the result does not depend on summation order.
 *   prio_fill    rows[i] (int32, -1 = skip) get `value` for t < episode_lengths[row], 0 beyond; row sums recomputed
 *   prio_update  prio[ep[b], t[b]] = value[b] (float32 [B]); row sums follow atomically; pairs outside the stored episodes
 *                (ep outside [0, capacity), t outside [0, episode_lengths[ep])) are ignored
 * A stored priority is at least one fixed-point unit, so no stored ply ever becomes undrawable.
 *   plan_prioritized  batch_size draws: target = floor(bits64 * total / 2^64) (threefry bits 2b, 2b+1 of host_key), episode
 *                     by binary search in the scanned row sums (cdf_work: uint64 [capacity] scratch), ply by a row scan;
 *                     prob[b] = P(e, t) as float64 (for importance weights) */
int dogstep_replay_prio_fill(uint32_t* prio, unsigned long long* row_sum, const int32_t* episode_lengths, int32_t max_episode_length,
                             const int32_t* rows, int32_t n_rows, float value, void* stream);
int dogstep_replay_prio_update(uint32_t* prio, unsigned long long* row_sum, const int32_t* episode_lengths, int32_t capacity,
                               int32_t max_episode_length, int32_t batch_size, const int32_t* ep_indices, const int32_t* t_starts,
                               const float* value, void* stream);
int dogstep_replay_plan_prioritized(const uint32_t* prio, const unsigned long long* row_sum, unsigned long long* cdf_work, int32_t size,
                                    int32_t max_episode_length, int32_t batch_size, const uint32_t* host_key, int32_t* ep_indices,
                                    int32_t* t_starts, double* prob, void* stream);

/* ---------------------------------------------------------------- self-play bookkeeping
 * One lockstep iteration of play_batch_of_games_jitted AFTER the search (MuZero_det_MADN/game_agent.py:64-148,
 * MuZero_Classic_MADN/game_agent_stochastic.py:86-204): for every game that is not done, env_step(map_action(action)) if a
 * legal action exists else no_step, reward / discount class targets, and the trajectory row at traj.episode_lengths[g]
 * ('idx'): obs, act, rew, val, pol, mask, player, team, discount (+ dice, dice_dist).  `traj` uses the replay struct with
 * capacity = n and max_episode_length = max_steps.  obs: int8 [n, C, total] = encode_board(env) taken BEFORE the step. */
int dogstep_madn_det_agent_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* action,
                                const float* root_value, const float* action_weights, const int8_t* obs,
                                const dogstep_replay_arrays* traj, void* stream);
int dogstep_madn_cls_agent_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* action,
                                const float* root_value, const float* action_weights, const int8_t* obs,
                                const dogstep_replay_arrays* traj, void* stream);
/* throw_die for the games that are not done (game_agent_stochastic.py:90 runs inside do_active_step) */
int dogstep_madn_cls_throw_die_active(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, void* stream);
/* The same iteration on the DOG env (BASELINE config 5).  The reference has no DOG self-play loop — MuZero_DOG/muzero_dog.py:85-99
 * are stubs and DOG/dog.py:1264-1272 has no observation encoder — so this applies do_active_step of
 * MuZero_det_MADN/game_agent.py:64-148 to DOG/dog.py's env_step (:1117) / no_step (:713) / valid_actions (:693):
 * traj.action_dim = dogstep_dog_num_actions(cfg) (806), obs: int8 [n, traj.obs_size] supplied by the caller (may be NULL
 * when traj.obs_size == 0). */
int dogstep_dog_agent_step(const dogstep_dog_state* s, int64_t n, const dogstep_dog_cfg* cfg, const int32_t* action,
                           const float* root_value, const float* action_weights, const int8_t* obs,
                           const dogstep_replay_arrays* traj, void* stream);

/* ---------------------------------------------------------------- evaluation loop
 * One lockstep iteration of play_eval_loop_jitted (MuZero_det_MADN/evaluate_agent.py:733-930) for every game that is not
 * done: the seat to move is played by agent_type[current_player] (host int32[4], the reference's params['type']):
 * 3 = random legal policy (do_random :774-778), 2 = the rule-based scorer (do_rule_based :780-878, categorical over
 * score / 0.25), anything else = tree search, whose action for this game is read from search_action[g] (device int32 [n];
 * may be NULL when no seat searches).  No legal action -> no_step.  Per-game step key = split(rng_key, n + 1)[g + 1] with
 * the global game index g = game_offset + local index, as in the other lockstep drivers.  winners int32 [n,4] (may be NULL)
 * accumulates manual_get_winner (:16-45) when a game ends.  Float contract of the scorer: IEEE float32 add / div in the
 * reference's order, gumbel = -log(-log(u)) with each log rounded once from double. */
int dogstep_madn_det_eval_step(const dogstep_madn_det_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* agent_type,
                               const int32_t* search_action, const uint32_t* host_rng_key, int64_t game_offset, int32_t* winners,
                               unsigned long long* active_count, void* stream);
/* The dice game's loop body (MuZero_Classic_MADN/evaluate_agent_stochastic.py:738-905) AFTER its throw_die (call
 * dogstep_madn_cls_throw_die_active first): four actions = the pin to move by env.die; rule-based scorer :782-872. */
int dogstep_madn_cls_eval_step(const dogstep_madn_cls_state* s, int64_t n, const dogstep_madn_cfg* cfg, const int32_t* agent_type,
                               const int32_t* search_action, const uint32_t* host_rng_key, int64_t game_offset, int32_t* winners,
                               unsigned long long* active_count, void* stream);

/* ---------------------------------------------------------------- TicTacToe (BASELINE config 1)
 * Batched leaves of `TicTacToe` / `TicTacToeV2` (TicTacToe/TicTacToe.py:12-17, TicTacToeV2.py:14-20).
 * variant 0 = TicTacToe.py, 1 = TicTacToeV2.py (last three moves per player persist; `memory`). */
typedef struct {
  int8_t* board;          /* [n, 3, 3]  0 empty, +1 / -1 */
  int8_t* current_player; /* [n]  +1 / -1 */
  int8_t* reward;         /* [n] */
  uint8_t* done;          /* [n] */
  int8_t* memory;         /* [n, 2, 3]  V2 only (still required; -1 filled) */
} dogstep_ttt_state;
/* env_reset — TicTacToeV2.py:38-45 */
int dogstep_ttt_reset(const dogstep_ttt_state* s, int64_t n, void* stream);
/* env_step — TicTacToe.py:42-59 / TicTacToeV2.py:46-79.  action int8 [n] */
int dogstep_ttt_step(const dogstep_ttt_state* s, int64_t n, int32_t variant, const int8_t* action, int8_t* reward, uint8_t* done,
                     void* stream);
/* One ply of play_match / play_mcts_match — TicTacToe/eval.py:97-125, :151-176 with get_mcts_action :28-34: every game that
 * is not done plays argmax(where(board == 0, action_weights, -inf)) (first maximum) through env_step, in place; finished
 * games are left untouched.  action_weights f32 [n,9] (PolicyOutput.action_weights); action i8 [n] or NULL receives the move
 * (-1 for a finished game); plies i32 [n] or NULL is incremented for every game that moved */
int dogstep_ttt_play_move(const dogstep_ttt_state* s, int64_t n, int32_t variant, const float* action_weights, int8_t* action,
                          int32_t* plies, void* stream);
/* policy_function (:96-102) -> logits f32 [n,9] and/or valid_action_mask (:81-82) -> u8 [n,9]; either may be NULL */
int dogstep_ttt_policy_function(const dogstep_ttt_state* s, int64_t n, int32_t variant, float* logits, uint8_t* valid_mask,
                                void* stream);
/* root_fn (:121-126): prior = policy_function(env), value = rollout(env, key), embedding = the env as 18 floats
 * (board 9, current_player, reward, done, memory 6).  keys u32 [n,2] */
int dogstep_ttt_root_fn(const dogstep_ttt_state* s, int64_t n, int32_t variant, const uint32_t* keys, float* prior_logits,
                        float* value, float* embedding, void* stream);
/* recurrent_fn (:128-140): env_step(embedding, action) then policy_function / rollout value on the successor */
int dogstep_ttt_recurrent_fn(int64_t n, int32_t variant, const uint32_t* keys, const int32_t* action, const float* embedding_in,
                             float* prior_logits, float* value, float* reward, float* discount, float* embedding_out,
                             void* stream);

/* run_mcts / run_gumbel of TicTacToe/mcts.py:9-38 as ONE launch per move: root_fn, tree init, num_simulations x (select ->
 * recurrent_fn on the true env, incl. its random rollout -> expand + backup) and the policy output, one game per warp with no
 * synchronisation between games (a simulation driven from the host waits for the longest rollout of the whole batch).  Same
 * device code, keys and order of operations as dogstep_ttt_root_fn / dogstep_mcts_init / _select / dogstep_ttt_recurrent_fn /
 * _expand / _policy_output called in sequence: bit-identical results.  search_keys u32 [n,2] = the rng_key handed to
 * mctx.muzero_policy (mcts.py:12 key1), root_keys u32 [n,2] = the key root_fn's rollout gets (split(key2, 1)[0], :15);
 * the scratch rows are per-game work buffers the caller allocates once; outputs as dogstep_mcts_policy_output.
 * cfg: num_actions 9, num_chance 0, embed_dim 18, dirichlet_fraction 0 (else DOGSTEP_ERR_UNSUPPORTED). */
typedef struct {
  int32_t* parent;          /* [n] */
  int32_t* action;          /* [n] */
  float* embedding;         /* [n, 18] parent embedding of the current simulation */
  uint32_t* expand_key;     /* [n, 2] */
  float* prior_logits;      /* [n, 9] */
  float* value;             /* [n] */
  float* reward;            /* [n] */
  float* discount;          /* [n] */
  float* next_embedding;    /* [n, 18] */
  float* root_prior_logits; /* [n, 9] */
  float* root_value;        /* [n] */
  float* root_embedding;    /* [n, 18] */
} dogstep_ttt_search_scratch;
int dogstep_ttt_search(const dogstep_ttt_state* s, int64_t n, int32_t variant, const dogstep_mcts_tree* t, const dogstep_mcts_cfg* cfg,
                       const uint32_t* search_keys, const uint32_t* root_keys, const dogstep_ttt_search_scratch* scratch,
                       int32_t* action_out, float* action_weights_out, float* root_value_out, void* stream);

/* ---------------------------------------------------------------- jax.random on device
 * Stand-ins for the jax.random calls the self-play drivers make around the env functions
 * (game_agent.py:60,187-188).  keys are raw uint32[2]. */
/* jax.random.split(key, n) -> out [n, 2] */
int dogstep_random_split(const uint32_t* host_key, int64_t n, uint32_t* out, void* stream);
/* jax.random.randint(key, (n,), lo, hi) int32 */
int dogstep_random_randint(const uint32_t* host_key, int64_t n, int32_t lo, int32_t hi, int32_t* out, void* stream);
/* jax.random.uniform(key, (n,), float32, lo, hi) */
int dogstep_random_uniform(const uint32_t* host_key, int64_t n, float lo, float hi, float* out, void* stream);
/* HOST-side scalar key arithmetic (plain CPU code, no stream): out uint32 [num,2] = jax.random.split(key, num); and the loop-key
 * chain rng <- split(rng, N + 1)[0] applied `steps` times (out uint32 [2]) — what a driver that launches k lockstep iterations
 * per call (play_random with max_steps = k) advances its key by between launches. */
int dogstep_host_split(const uint32_t* key, int32_t num, uint32_t* out);
int dogstep_host_key_chain(const uint32_t* key, int32_t steps, uint32_t* out);
/* rng_key, *step_keys = jax.random.split(rng_key, n + 1) (MuZero_det_MADN/game_agent.py:60) with the loop key uint32 [2] ON
 * THE DEVICE: step_keys uint32 [n,2] = split(key, n + 1)[1:], then key <- split(key, n + 1)[0] in place.  No host value is
 * baked into the launch, so a lockstep iteration that starts with this call can be captured once and replayed as a CUDA graph. */
int dogstep_random_split_chain(uint32_t* key, int64_t n, uint32_t* step_keys, void* stream);
/* out[i] = jax.random.split(keys[i], m)[index] for every i: keys, out uint32 [n,2] on the device */
int dogstep_random_split_each(const uint32_t* keys, int64_t n, uint32_t index, uint32_t* out, void* stream);
/* jax.random.bits(key, (n,), uint32) */
int dogstep_random_bits(const uint32_t* host_key, int64_t n, uint32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DOGSTEP_H */
