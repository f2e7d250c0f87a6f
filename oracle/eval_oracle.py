"""eval_oracle.py — TEST INFRASTRUCTURE ONLY.

NumPy restatement of play_eval_loop_jitted (MuZero_det_MADN/evaluate_agent.py:733-930) on top of the C env oracle: per
lockstep iteration and live game, the seat to move is played by its agent type (params['type']: 3 random :774-778,
2 rule based :780-878, else tree search — supplied by the caller), no legal action -> no_step, winners accumulate
manual_get_winner (:16-45) when a game ends.

PARITY STATUS: pinned since round 2 by the reference's own play_eval_loop_jitted run on oracle/jaxshim
(tests/golden/loops_reference.npz, tests/test_golden_loops.py: winners and every final leaf of random and rule-based seats,
20 games per run to termination); the env functions underneath are pinned by madn_oracle.c's goldens.  do_rule_based is
restated array expression by array expression (float32, same association).  One caveat: jax.random.categorical's float
Gumbel uses -log(-log(u)) with each log rounded once from a double log — this repo's float contract, the same on
the shim, here and on the GPU; XLA's own float32 log may differ by 1 ulp, which matters only on an exact near-tie.

Note for maintainers: do_rule_based builds its candidate cells from `actions = jnp.arange(6)` (distances 0..5) while action
(pin, k) plays move k + 1 (map_action), so every bonus is evaluated one cell short.  Restated literally — the code is the
spec — the scorer loses to the random policy about 15 : 1 (tests/test_eval_loop.py); with distances 1..6 it wins every game.
"""
import math

import numpy as np

import oracle as O

TINY = np.float32(1.17549435e-38)


def _rule(cfg, bit):
    return bool(cfg.rules & bit)


RULE_TEAMS, RULE_MTS = 1 << 0, 1 << 8  # include/dogstep_rules.h


def gumbel24(key):
    u = O.uniform(key, 24, float(TINY), 1.0)
    l1 = np.array([math.log(float(v)) for v in u], np.float64).astype(np.float32)
    return -np.array([math.log(float(-v)) for v in l1], np.float64).astype(np.float32)


def rule_based_action(cfg, geometry, pins, current_player, valid_mask, key):
    """do_rule_based (:780-878) for one game.  pins int8 [P,4], valid_mask bool [24]."""
    start, target, goal = geometry
    board_size = 4 * cfg.distance
    cp = int(current_player)
    current_goal = goal[cp]                                                   # (4,)
    current_positions = pins[cp].astype(np.int8)[:, None]                     # (4, 1)
    actions = np.arange(6, dtype=np.int32)
    moved_positions = current_positions.astype(np.int32) + actions            # (4, 6)
    fitted_positions = moved_positions % board_size
    x = moved_positions - int(target[cp]) - int(_rule(cfg, RULE_MTS))
    goal_pick = np.asarray(current_goal)[np.clip(np.where(x - 1 < 0, x - 1 + 4, x - 1), 0, 3)]  # gather: wrap once, then clamp
    new_positions = np.where(current_positions < 0, int(start[cp]),
                             np.where(current_positions >= board_size, moved_positions,
                                      np.where((4 >= x) & (x > 0) & (current_positions <= int(target[cp])), goal_pick,
                                               fitted_positions))).astype(np.int32)
    opp = np.ones_like(pins, dtype=np.int32)
    opp[cp] = 0
    if _rule(cfg, RULE_TEAMS):
        opp[(cp + 2) % 4] = 0
    opponent_pins = np.where(opp == 1, pins, -1).flatten()
    pins_in_home = int(np.sum(pins[cp] < 0))
    vm = valid_mask.reshape(4, 6)
    action_counts = vm.sum(0).astype(np.int32)
    action_abundance = action_counts.astype(np.float32) / np.float32(max(float(action_counts.sum()), 1.0))
    base_score = np.repeat(action_abundance, 4)                               # (24,): abundance[a // 4] — the reference's indexing
    goal_bonus = np.where(np.isin(new_positions, current_goal) & (current_positions < board_size), np.float32(5.0), np.float32(0.0)).flatten()
    out_w = np.float32(3.0 if pins_in_home >= 2 else 2.0)
    out_bonus = np.where((current_positions < 0) & (new_positions == int(start[cp])), out_w, np.float32(0.0)).flatten()
    hit_bonus = np.where((new_positions != current_positions) & np.isin(new_positions, opponent_pins), np.float32(2.0), np.float32(0.0)).flatten()
    scores = ((base_score.astype(np.float32) + goal_bonus.astype(np.float32)).astype(np.float32) + out_bonus.astype(np.float32)).astype(np.float32)
    scores = (scores + hit_bonus.astype(np.float32)).astype(np.float32)
    scores = np.where(valid_mask, scores, np.float32(-np.inf)).astype(np.float32)
    logits = (scores / np.float32(0.25)).astype(np.float32)
    return int(np.argmax((gumbel24(key) + logits).astype(np.float32)))


def manual_get_winner(cfg, geometry, board):
    _, _, goal = geometry
    done = np.array([bool((board[goal[p]] >= 0).all()) if p < cfg.num_players else False for p in range(4)])
    if _rule(cfg, RULE_TEAMS):
        t0, t1 = done[0] and done[2], done[1] and done[3]
        if (t0 and t1) or not (t0 or t1):
            return np.zeros(4, bool)
        return np.array([True, False, True, False]) if t0 else np.array([False, True, False, True])
    return done


def play_eval_loop(state, agent_types, rng_key, search_fn=None, max_steps=2000):
    """state: O.MadnState (det), stepped in place.  search_fn(step_keys [n,2], valid [n,24]) -> action int [n] for search seats.
    Returns winners int32 [n,4]."""
    cfg, n = state.cfg, state.n
    geo = cfg.geometry()
    winners = np.zeros((n, 4), np.int32)
    key = np.asarray(rng_key, np.uint32)
    step = 0
    while step < max_steps and not state.done.all():
        keys = O.split(key, n + 1)
        key, step_keys = keys[0], keys[1:]
        live = state.done == 0
        valid = O.madn_det_valid_action(state).reshape(n, 24)
        action = np.zeros(n, np.int64)
        searched = search_fn(step_keys, valid) if search_fn is not None else None
        for g in range(n):
            if not live[g] or not valid[g].any():
                continue
            t = agent_types[int(state.current_player[g])]
            if t == 3:
                action[g] = O.categorical_masked(step_keys[g], valid[g].astype(np.uint8))
            elif t == 2:
                action[g] = rule_based_action(cfg, geo, state.pins[g], state.current_player[g], valid[g], step_keys[g])
            else:
                action[g] = int(searched[g])
        stepped, skipped = state.copy(), state.copy()
        _, d = O.madn_det_step(stepped, np.stack([action // 6, action % 6 + 1], 1).astype(np.int8))
        O.madn_det_no_step(skipped)
        for g in range(n):
            if not live[g]:
                continue
            src = stepped if valid[g].any() else skipped
            for f, v in src.fields().items():
                getattr(state, f)[g] = v[g]
            if state.done[g]:
                winners[g] += manual_get_winner(cfg, geo, state.board[g]).astype(np.int32)
        step += 1
    return winners


# --------------------------------------------------------------------------- the dice game
def gumbel_n(key, n):
    u = O.uniform(key, n, float(TINY), 1.0)
    l1 = np.array([math.log(float(v)) for v in u], np.float64).astype(np.float32)
    return -np.array([math.log(float(-v)) for v in l1], np.float64).astype(np.float32)


def cls_rule_based_action(cfg, geometry, pins, current_player, die, valid_mask, key):
    """do_rule_based of MuZero_Classic_MADN/evaluate_agent_stochastic.py:782-872 for one game.  pins int8 [P,4], valid bool [4]."""
    start, target, goal = geometry
    board_size = 4 * cfg.distance
    cp = int(current_player)
    current_positions = pins[cp].astype(np.int32)
    current_goal = np.asarray(goal[cp])
    moved_positions = current_positions + int(die)
    fitted_positions = moved_positions % board_size
    x = moved_positions - int(target[cp]) - int(_rule(cfg, RULE_MTS))
    goal_pick = current_goal[np.clip(np.where(x - 1 < 0, x - 1 + 4, x - 1), 0, 3)]
    new_positions = np.where(current_positions < 0, int(start[cp]),
                             np.where(current_positions >= board_size, moved_positions,
                                      np.where((4 >= x) & (x > 0) & (current_positions <= int(target[cp])), goal_pick, fitted_positions)))
    opp = np.ones_like(pins, dtype=np.int32)
    opp[cp] = 0
    if _rule(cfg, RULE_TEAMS):
        opp[(cp + 2) % 4] = 0
    opponent_pins = np.where(opp == 1, pins, -1).flatten()
    pins_in_home = int(np.sum(current_positions < 0))
    base_score = np.zeros(4, np.float32)
    goal_bonus = np.where(np.isin(new_positions, current_goal) & (current_positions < board_size), np.float32(5.0), np.float32(0.0))
    out_bonus = np.where((current_positions < 0) & (new_positions == int(start[cp])), np.float32(3.0 if pins_in_home >= 2 else 2.0), np.float32(0.0))
    hit_bonus = np.where((new_positions != current_positions) & np.isin(new_positions, opponent_pins), np.float32(2.5), np.float32(0.0))
    scores = ((base_score + goal_bonus.astype(np.float32)).astype(np.float32) + out_bonus.astype(np.float32)).astype(np.float32)
    scores = (scores + hit_bonus.astype(np.float32)).astype(np.float32)
    scores = np.where(valid_mask, scores, np.float32(-np.inf)).astype(np.float32)
    logits = (scores / np.float32(0.25)).astype(np.float32)
    return int(np.argmax((gumbel_n(key, 4) + logits).astype(np.float32)))


def play_eval_loop_classic(state, agent_types, rng_key, search_fn=None, max_steps=2000):
    """play_eval_loop_jitted of MuZero_Classic_MADN/evaluate_agent_stochastic.py:738-905 on the C env oracle.
    state: O.MadnState (det=False), stepped in place.  Returns winners int32 [n,4]."""
    cfg, n = state.cfg, state.n
    geo = cfg.geometry()
    winners = np.zeros((n, 4), np.int32)
    key = np.asarray(rng_key, np.uint32)
    step = 0
    while step < max_steps and not state.done.all():
        keys = O.split(key, n + 1)
        key, step_keys = keys[0], keys[1:]
        live = state.done == 0
        thrown = state.copy()                      # env = throw_die(env) inside do_step: live games only
        O.madn_cls_throw_die(thrown)
        state.die[live] = thrown.die[live]
        state.key[live] = thrown.key[live]
        valid = O.madn_cls_valid_action(state).reshape(n, 4)
        action = np.zeros(n, np.int64)
        searched = search_fn(step_keys, valid) if search_fn is not None else None
        for g in range(n):
            if not live[g] or not valid[g].any():
                continue
            t = agent_types[int(state.current_player[g])]
            if t == 3:
                action[g] = O.categorical_masked(step_keys[g], valid[g].astype(np.uint8))
            elif t == 2:
                action[g] = cls_rule_based_action(cfg, geo, state.pins[g], state.current_player[g], state.die[g], valid[g], step_keys[g])
            else:
                action[g] = int(searched[g])
        stepped, skipped = state.copy(), state.copy()
        O.madn_cls_step(stepped, action.astype(np.int8))
        O.madn_cls_no_step(skipped)
        for g in range(n):
            if not live[g]:
                continue
            src = stepped if valid[g].any() else skipped
            for f, v in src.fields().items():
                getattr(state, f)[g] = v[g]
            if state.done[g]:
                winners[g] += manual_get_winner(cfg, geo, state.board[g]).astype(np.int32)
        step += 1
    return winners
