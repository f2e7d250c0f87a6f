/* mcts_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Scalar restatement of the tree search the reference runs through the third-party package
 * mctx 0.0.6 (pinned in /root/reference/uv.lock:655-656; its source is NOT under /root/reference
 * and the package is not installed in this image).  Call sites in the reference:
 *   mctx.gumbel_muzero_policy      MuZero_det_MADN/muzero_deterministic_madn.py:673-684
 *   mctx.stochastic_muzero_policy  MuZero_Classic_MADN/muzero_classic_madn.py:488-501
 *   mctx.muzero_policy / gumbel    TicTacToe/mcts.py:13-22, 29-37
 * Restated from mctx's published algorithm (mctx/_src/{search,policies,action_selection,qtransforms,
 * seq_halving,tree}.py): search() = num_simulations x { simulate -> expand -> backward };
 * muzero_action_selection (PUCT + 1e-7 uniform tie-break), gumbel root (sequential halving) / interior
 * selection, stochastic decision/chance wrapper, the three qtransforms, and the policy epilogues.
 *
 * PARITY STATUS: **bit-level parity unpinned** — the reference holds no test or golden vector for any MCTS result and
 * mctx cannot be run here.  This file is an independent scalar statement of the same algorithm that the CUDA
 * kernels are compared against bit-for-bit (visit counts exact, values bit-equal) on identical network outputs.
 * STATISTICALLY pinned by what the reference measured with mctx itself (TicTacToe/results.md:12-15, :53-68): with the
 * true-env callbacks of ttt_oracle.c, muzero_policy and gumbel_muzero_policy here reproduce the recorded win / loss / tie
 * rates against a random bot (5 / 10 / 30 / 100 simulations) and the recorded first-player / second-player / draw shares of
 * search-vs-search play (e.g. 100 simulations: 64.1 / 20.3 / 15.6 % recorded, 65.6 / 20.9 / 13.5 % here, 1,000 games each)
 * within sampling error — tests/test_ttt.py.
 *
 * Float contract shared with the kernels: no FMA contraction (-ffp-contract=off), exp/log rounded from double,
 * sums in the fixed order  partial[l] = x[l] + x[l+32] + ... ; then butterfly over l^16, l^8, l^4, l^2, l^1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include "../include/dogstep.h"
#include "jaxrand_oracle.h"

static float f_exp(float x) { return (float)exp((double)x); }
static float f_log(float x) { return (float)log((double)x); }

/* the fixed summation order of the float contract */
static float sum_tree(const float *x, int n) {
  float part[32];
  for (int l = 0; l < 32; ++l) {
    float acc = 0.0f;
    for (int a = l; a < n; a += 32) acc = acc + x[a];
    part[l] = acc;
  }
  for (int o = 16; o; o >>= 1) {
    float nxt[32];
    for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ o];
    memcpy(part, nxt, sizeof part);
  }
  return part[0];
}

static void softmax(const float *x, int n, float *p) {
  float m = -INFINITY;
  for (int a = 0; a < n; ++a) if (x[a] > m) m = x[a];
  float e[1024];
  for (int a = 0; a < n; ++a) e[a] = f_exp(x[a] - m);
  float s = sum_tree(e, n);
  for (int a = 0; a < n; ++a) p[a] = e[a] / s;
}

static int argmax_first(const float *x, int n) {
  int best = 0;
  for (int a = 1; a < n; ++a) if (x[a] > x[best]) best = a;
  return best;
}

/* one game's view of the tree */
typedef struct {
  int N, A, E; /* A = A' (children width) */
  int32_t *node_visits; float *raw_values, *node_values; int32_t *parents, *action_from_parent;
  int32_t *children_index; float *children_prior_logits; int32_t *children_visits;
  float *children_rewards, *children_discounts, *children_values, *embeddings;
  uint8_t *is_decision, *root_invalid; float *root_gumbel; uint32_t *search_key, *policy_key;
} gtree;

static gtree view(const dogstep_mcts_tree *t, const dogstep_mcts_cfg *c, int64_t g) {
  gtree v;
  v.N = c->num_simulations + 1; v.A = c->num_actions + c->num_chance; v.E = c->embed_dim;
  int64_t nn = (int64_t)v.N, na = nn * v.A;
  v.node_visits = t->node_visits + g * nn; v.raw_values = t->raw_values + g * nn; v.node_values = t->node_values + g * nn;
  v.parents = t->parents + g * nn; v.action_from_parent = t->action_from_parent + g * nn;
  v.children_index = t->children_index + g * na; v.children_prior_logits = t->children_prior_logits + g * na;
  v.children_visits = t->children_visits + g * na; v.children_rewards = t->children_rewards + g * na;
  v.children_discounts = t->children_discounts + g * na; v.children_values = t->children_values + g * na;
  v.embeddings = t->embeddings + g * nn * v.E;
  v.is_decision = t->is_decision ? t->is_decision + g * nn : NULL;
  v.root_invalid = t->root_invalid_actions + g * v.A;
  v.root_gumbel = t->root_gumbel ? t->root_gumbel + g * v.A : NULL;
  v.search_key = t->search_key + 2 * g; v.policy_key = t->policy_key + 2 * g;
  return v;
}

/* tree.qvalues(node) */
static void qvalues(const gtree *t, int node, float *q) {
  for (int a = 0; a < t->A; ++a) {
    int k = node * t->A + a;
    q[a] = t->children_rewards[k] + t->children_discounts[k] * t->children_values[k];
  }
}

/* qtransforms.py */
static void qtransform(const gtree *t, const dogstep_mcts_cfg *c, int node, float *out) {
  const int A = t->A;
  float q[1024];
  qvalues(t, node, q);
  const int32_t *vc = t->children_visits + node * A;
  if (c->qtransform == DOGSTEP_Q_BY_MIN_MAX) {
    for (int a = 0; a < A; ++a) out[a] = ((vc[a] > 0 ? q[a] : c->q_min) - c->q_min) / (c->q_max - c->q_min);
  } else if (c->qtransform == DOGSTEP_Q_BY_PARENT_AND_SIBLINGS) {
    float nv = t->node_values[node], mn = nv, mx = nv;
    for (int a = 0; a < A; ++a) { float s = vc[a] > 0 ? q[a] : nv; if (s < mn) mn = s; if (s > mx) mx = s; }
    float den = mx - mn; if (!(den > c->epsilon)) den = c->epsilon; /* jnp.maximum(max - min, eps) */
    for (int a = 0; a < A; ++a) out[a] = ((vc[a] > 0 ? q[a] : mn) - mn) / den;
  } else { /* completed_by_mix_value */
    float p[1024], tmp[1024];
    softmax(t->children_prior_logits + node * A, A, p);
    int sum_vc = 0, maxvisit = 0;
    for (int a = 0; a < A; ++a) { sum_vc += vc[a]; if (vc[a] > maxvisit) maxvisit = vc[a]; }
    for (int a = 0; a < A; ++a) { if (p[a] < FLT_MIN) p[a] = FLT_MIN; tmp[a] = vc[a] > 0 ? p[a] : 0.0f; }
    float sum_probs = sum_tree(tmp, A);
    for (int a = 0; a < A; ++a) tmp[a] = vc[a] > 0 ? (p[a] * q[a]) / sum_probs : 0.0f;
    float weighted_q = sum_tree(tmp, A);
    float value = (t->raw_values[node] + (float)sum_vc * weighted_q) / (float)(sum_vc + 1);
    float mn = INFINITY, mx = -INFINITY;
    for (int a = 0; a < A; ++a) { tmp[a] = vc[a] > 0 ? q[a] : value; if (tmp[a] < mn) mn = tmp[a]; if (tmp[a] > mx) mx = tmp[a]; }
    float den = mx - mn; if (!(den > c->epsilon)) den = c->epsilon;
    float scale = (c->maxvisit_init + (float)maxvisit) * c->value_scale;
    for (int a = 0; a < A; ++a) out[a] = scale * ((tmp[a] - mn) / den);
  }
}

/* seq_halving.get_sequence_of_considered_visits(m, S)[i] without the table */
static int considered_visit(int m, int S, int i) {
  if (i > S - 1) i = S - 1; /* table gather clamps */
  if (i < 0) i = 0;
  if (m <= 1) return i;
  int log2max = 0; while ((1 << log2max) < m) ++log2max;
  int k = m, pos = 0, base = 0;
  for (;;) {
    int extra = S / (log2max * k); if (extra < 1) extra = 1;
    int len = k * extra;
    if (i < pos + len) return base + (i - pos) / k;
    pos += len; base += extra; k = k / 2 < 2 ? 2 : k / 2;
  }
}

/* seq_halving.score_considered */
static void score_considered(int cv, const float *gumbel, const float *logits, const float *nq, const int32_t *vc, int A, float *out) {
  float mx = -INFINITY;
  for (int a = 0; a < A; ++a) if (logits[a] > mx) mx = logits[a];
  for (int a = 0; a < A; ++a) {
    float v = (gumbel[a] + (logits[a] - mx)) + nq[a];
    if (!(v > -1e9f)) v = -1e9f;
    out[a] = v + (vc[a] == cv ? 0.0f : -INFINITY);
  }
}

/* action_selection.muzero_action_selection */
static int select_puct(const gtree *t, const dogstep_mcts_cfg *c, int node, int depth, const uint32_t key[2]) {
  const int A = t->A;
  const int32_t *vc = t->children_visits + node * A;
  float nvis = (float)t->node_visits[node];
  float pb_c = c->pb_c_init + f_log(((nvis + c->pb_c_base) + 1.0f) / c->pb_c_base);
  float p[1024], vs[1024], score[1024];
  softmax(t->children_prior_logits + node * A, A, p);
  qtransform(t, c, node, vs);
  float sq = sqrtf(nvis);
  for (int a = 0; a < A; ++a) {
    float policy = ((sq * pb_c) * p[a]) / (float)(vc[a] + 1);
    float noise = 1e-7f * orc_uniform_i(key, (uint32_t)a, 0.0f, 1.0f);
    score[a] = (vs[a] + policy) + noise;
    if (depth == 0 && t->root_invalid[a]) score[a] = -INFINITY;
  }
  return argmax_first(score, A);
}

static int select_gumbel_root(const gtree *t, const dogstep_mcts_cfg *c, int node) {
  const int A = t->A;
  const int32_t *vc = t->children_visits + node * A;
  float cq[1024], score[1024];
  qtransform(t, c, node, cq);
  int num_valid = 0, sim_index = 0;
  for (int a = 0; a < A; ++a) { num_valid += 1 - (t->root_invalid[a] != 0); sim_index += vc[a]; }
  int m = c->max_num_considered_actions < num_valid ? c->max_num_considered_actions : num_valid;
  int cv = considered_visit(m, c->num_simulations, sim_index);
  score_considered(cv, t->root_gumbel, t->children_prior_logits + node * A, cq, vc, A, score);
  for (int a = 0; a < A; ++a) if (t->root_invalid[a]) score[a] = -INFINITY;
  return argmax_first(score, A);
}

static int select_gumbel_interior(const gtree *t, const dogstep_mcts_cfg *c, int node) {
  const int A = t->A;
  const int32_t *vc = t->children_visits + node * A;
  float cq[1024], x[1024], p[1024];
  qtransform(t, c, node, cq);
  int sum_vc = 0;
  for (int a = 0; a < A; ++a) { x[a] = t->children_prior_logits[node * A + a] + cq[a]; sum_vc += vc[a]; }
  softmax(x, A, p);
  for (int a = 0; a < A; ++a) x[a] = p[a] - (float)vc[a] / (float)(1 + sum_vc);
  return argmax_first(x, A);
}

static int select_chance(const gtree *t, int node) {
  const int A = t->A;
  float p[1024];
  softmax(t->children_prior_logits + node * A, A, p);
  for (int a = 0; a < A; ++a) p[a] = p[a] / (float)(t->children_visits[node * A + a] + 1);
  return argmax_first(p, A);
}

static int select_action(const gtree *t, const dogstep_mcts_cfg *c, int node, int depth, const uint32_t key[2]) {
  if (c->policy == DOGSTEP_MCTS_GUMBEL) return depth == 0 ? select_gumbel_root(t, c, node) : select_gumbel_interior(t, c, node);
  if (c->policy == DOGSTEP_MCTS_STOCHASTIC && !t->is_decision[node]) return select_chance(t, node);
  return select_puct(t, c, node, depth, key);
}

/* _mask_invalid_actions (policies.py) */
static void mask_invalid(float *logits, const uint8_t *invalid, int A) {
  float mx = -INFINITY;
  for (int a = 0; a < A; ++a) if (logits[a] > mx) mx = logits[a];
  for (int a = 0; a < A; ++a) logits[a] = (invalid && invalid[a]) ? -FLT_MAX : logits[a] - mx;
}

int orc_mcts_init(const dogstep_mcts_tree *tr, int64_t n, const dogstep_mcts_cfg *c, const uint32_t *keys,
                  const float *root_prior_logits, const float *root_value, const float *root_embedding,
                  const uint8_t *invalid_actions, const float *dirichlet_noise) {
  const int A0 = c->num_actions;
  for (int64_t g = 0; g < n; ++g) {
    gtree t = view(tr, c, g);
    const int A = t.A;
    for (int k = 0; k < t.N; ++k) { t.node_visits[k] = 0; t.raw_values[k] = 0; t.node_values[k] = 0; t.parents[k] = -1; t.action_from_parent[k] = -1; }
    for (int k = 0; k < t.N * A; ++k) {
      t.children_index[k] = -1; t.children_prior_logits[k] = 0; t.children_visits[k] = 0;
      t.children_rewards[k] = 0; t.children_discounts[k] = 0; t.children_values[k] = 0;
    }
    memset(t.embeddings, 0, sizeof(float) * (size_t)t.N * t.E);
    if (t.is_decision) { memset(t.is_decision, 0, (size_t)t.N); t.is_decision[0] = 1; }
    const uint32_t *key = keys + 2 * g;
    const uint8_t *inv = invalid_actions ? invalid_actions + g * A0 : NULL;
    float logits[1024];
    memcpy(logits, root_prior_logits + g * A0, sizeof(float) * A0);
    if (c->policy == DOGSTEP_MCTS_GUMBEL) {
      /* root prior = mask(prior); rng, gumbel_rng = split(key); gumbel = scale * gumbel(gumbel_rng, (A,)) */
      mask_invalid(logits, inv, A0);
      uint32_t k0[2], k1[2];
      orc_split_i(key, 0, k0); orc_split_i(key, 1, k1);
      for (int a = 0; a < A0; ++a) {
        float u = orc_uniform_i(k1, (uint32_t)a, FLT_MIN, 1.0f);
        t.root_gumbel[a] = c->gumbel_scale * (-f_log(-f_log(u)));
      }
      t.search_key[0] = k0[0]; t.search_key[1] = k0[1];
      t.policy_key[0] = t.policy_key[1] = 0;
    } else {
      /* rng, dirichlet_rng, search_rng = split(key, 3); noisy = log((1-f)*softmax(prior) + f*noise) masked */
      uint32_t k0[2], k2[2];
      orc_split_i(key, 0, k0); orc_split_i(key, 2, k2);
      float p[1024];
      softmax(logits, A0, p);
      for (int a = 0; a < A0; ++a) {
        float pr = p[a];
        if (dirichlet_noise) pr = (1.0f - c->dirichlet_fraction) * pr + c->dirichlet_fraction * dirichlet_noise[g * A0 + a];
        float lg = f_log(pr);
        logits[a] = lg > -FLT_MAX ? lg : -FLT_MAX; /* jnp.maximum(log(p), finfo.min) */
      }
      mask_invalid(logits, inv, A0);
      t.search_key[0] = k2[0]; t.search_key[1] = k2[1];
      t.policy_key[0] = k0[0]; t.policy_key[1] = k0[1];
    }
    for (int a = 0; a < A; ++a) {
      t.children_prior_logits[a] = a < A0 ? logits[a] : -INFINITY; /* chance slots padded with -inf */
      t.root_invalid[a] = a < A0 ? (inv ? inv[a] : 0) : 1;
    }
    t.raw_values[0] = t.node_values[0] = root_value[g];
    t.node_visits[0] = 1;
    memcpy(t.embeddings, root_embedding + g * t.E, sizeof(float) * t.E);
  }
  return 0;
}

/* search.simulate */
int orc_mcts_select(const dogstep_mcts_tree *tr, int64_t n, const dogstep_mcts_cfg *c, int32_t sim, int32_t *parent_out,
                    int32_t *action_out, float *embedding_out, uint8_t *is_decision_out, uint32_t *expand_key_out) {
  (void)sim;
  for (int64_t g = 0; g < n; ++g) {
    gtree t = view(tr, c, g);
    /* rng, simulate_key, expand_key = split(rng, 3); simulate_keys = split(simulate_key, 1) */
    uint32_t k0[2], k1[2], r[2];
    orc_split_i(t.search_key, 0, k0); orc_split_i(t.search_key, 1, k1);
    if (expand_key_out) orc_split_i(t.search_key, 2, expand_key_out + 2 * g);
    t.search_key[0] = k0[0]; t.search_key[1] = k0[1];
    orc_split_i(k1, 0, r);
    int node = 0, depth = 0, action = 0, parent = 0;
    for (;;) {
      uint32_t nr[2], ak[2];
      orc_split_i(r, 0, nr); orc_split_i(r, 1, ak);
      r[0] = nr[0]; r[1] = nr[1];
      action = select_action(&t, c, node, depth, ak);
      parent = node;
      int next = t.children_index[node * t.A + action];
      ++depth;
      if (next == -1 || depth >= c->max_depth) break;
      node = next;
    }
    parent_out[g] = parent; action_out[g] = action;
    memcpy(embedding_out + g * t.E, t.embeddings + (int64_t)parent * t.E, sizeof(float) * t.E);
    if (is_decision_out) is_decision_out[g] = t.is_decision ? t.is_decision[parent] : 1;
  }
  return 0;
}

/* search.expand + search.backward */
int orc_mcts_expand(const dogstep_mcts_tree *tr, int64_t n, const dogstep_mcts_cfg *c, int32_t sim, const int32_t *parent_in,
                    const int32_t *action_in, const float *prior_logits, const float *value, const float *reward,
                    const float *discount, const float *embedding, const float *chance_logits, const float *afterstate_value,
                    const float *afterstate_embedding) {
  const int A0 = c->num_actions, C = c->num_chance;
  for (int64_t g = 0; g < n; ++g) {
    gtree t = view(tr, c, g);
    const int A = t.A;
    int parent = parent_in[g], action = action_in[g];
    int node = t.children_index[parent * A + action];
    if (node == -1) node = sim + 1;
    float lg[1024], v, rw, dc;
    const float *emb;
    int parent_is_decision = t.is_decision ? t.is_decision[parent] : 0;
    if (c->policy == DOGSTEP_MCTS_STOCHASTIC && parent_is_decision) {
      for (int a = 0; a < A; ++a) lg[a] = a < A0 ? -INFINITY : chance_logits[g * C + (a - A0)];
      v = afterstate_value[g]; rw = 0.0f; dc = 1.0f;
      emb = afterstate_embedding + g * t.E;
    } else {
      for (int a = 0; a < A; ++a) lg[a] = a < A0 ? prior_logits[g * A0 + a] : -INFINITY;
      v = value[g]; rw = reward[g]; dc = discount[g];
      emb = embedding + g * t.E;
    }
    /* update_tree_node */
    memcpy(t.children_prior_logits + node * A, lg, sizeof(float) * A);
    t.raw_values[node] = v; t.node_values[node] = v; t.node_visits[node] += 1;
    memcpy(t.embeddings + (int64_t)node * t.E, emb, sizeof(float) * t.E);
    if (t.is_decision) t.is_decision[node] = (uint8_t)!parent_is_decision;
    t.children_index[parent * A + action] = node;
    t.children_rewards[parent * A + action] = rw;
    t.children_discounts[parent * A + action] = dc;
    t.parents[node] = parent; t.action_from_parent[node] = action;
    /* backward */
    float leaf = t.node_values[node];
    int idx = node;
    while (idx != 0) {
      int p = t.parents[idx], a = t.action_from_parent[idx];
      float cnt = (float)t.node_visits[p];
      leaf = t.children_rewards[p * A + a] + t.children_discounts[p * A + a] * leaf;
      float pv = (t.node_values[p] * cnt + leaf) / (cnt + 1.0f);
      t.node_values[p] = pv;
      t.node_visits[p] += 1;
      t.children_values[p * A + a] = t.node_values[idx];
      t.children_visits[p * A + a] += 1;
      idx = p;
    }
  }
  return 0;
}

/* policy epilogues (policies.py) */
int orc_mcts_policy_output(const dogstep_mcts_tree *tr, int64_t n, const dogstep_mcts_cfg *c, int32_t *action_out,
                           float *action_weights, float *root_value) {
  const int A0 = c->num_actions;
  for (int64_t g = 0; g < n; ++g) {
    gtree t = view(tr, c, g);
    const int A = t.A;
    const int32_t *vc = t.children_visits; /* root row */
    root_value[g] = t.node_values[0];
    if (c->policy == DOGSTEP_MCTS_GUMBEL) {
      int cv = 0;
      for (int a = 0; a < A; ++a) if (vc[a] > cv) cv = vc[a];
      float cq[1024], score[1024], lg[1024];
      qtransform(&t, c, 0, cq);
      score_considered(cv, t.root_gumbel, t.children_prior_logits, cq, vc, A, score);
      for (int a = 0; a < A; ++a) if (t.root_invalid[a]) score[a] = -INFINITY;
      action_out[g] = argmax_first(score, A);
      for (int a = 0; a < A; ++a) lg[a] = t.children_prior_logits[a] + cq[a];
      mask_invalid(lg, t.root_invalid, A);
      softmax(lg, A, action_weights + g * A0);
    } else {
      /* summary(): visit_probs over the DECISION actions; action ~ categorical(rng, log(probs)/T) */
      int tot = 0;
      for (int a = 0; a < A0; ++a) tot += vc[a];
      float lg[1024], mx = -INFINITY;
      for (int a = 0; a < A0; ++a) {
        float pr = tot > 0 ? (float)vc[a] / (float)(tot > 1 ? tot : 1) : 1.0f / (float)A0;
        action_weights[g * A0 + a] = pr;
        float l = f_log(pr);
        lg[a] = l > -FLT_MAX ? l : -FLT_MAX;
        if (lg[a] > mx) mx = lg[a];
      }
      float temp = c->temperature > FLT_MIN ? c->temperature : FLT_MIN;
      float score[1024];
      for (int a = 0; a < A0; ++a) {
        float u = orc_uniform_i(t.policy_key, (uint32_t)a, FLT_MIN, 1.0f);
        score[a] = -f_log(-f_log(u)) + (lg[a] - mx) / temp;
      }
      action_out[g] = argmax_first(score, A0);
    }
  }
  return 0;
}

int orc_considered_visit(int m, int S, int i) { return considered_visit(m, S, i); }
