"""selfplay_oracle.py — TEST INFRASTRUCTURE ONLY.

NumPy restatement of the bookkeeping of play_batch_of_games_jitted (MuZero_det_MADN/game_agent.py:50-183 and
MuZero_Classic_MADN/game_agent_stochastic.py:52-218) on top of the C env oracle: per lockstep iteration and live game,
encode / mask / (search result supplied by the caller) / env_step or no_step / reward+discount class targets / the
trajectory row.  The search itself is an input (`search_fn`), exactly as in the CUDA mirror.

PARITY STATUS: pinned since round 2 by the reference's own play_batch_of_games_jitted (both game agents) run on
oracle/jaxshim with the search replaced by a deterministic stand-in (tests/golden/loops_reference.npz,
tests/test_golden_loops.py: every trajectory buffer of 2 x 2 runs, incl. games that end inside the recording); the env
functions underneath are pinned by madn_oracle.c's goldens.
"""
import numpy as np

import oracle as O


def play_batch_of_games(state, max_steps, rng_key, search_fn, teams):
    """state: O.MadnState (det or classic).  search_fn(step_keys [n,2], obs, invalid [n,A]) -> (action, weights, value)."""
    n, det = state.n, state.det
    A = 24 if det else 4
    C_, T = (8 * state.cfg.num_players + 2, state.cfg.total) if det else (2 * state.cfg.num_players + 3, state.cfg.total)
    buf = dict(obs=np.zeros((n, max_steps, C_, T), np.float32), act=np.zeros((n, max_steps), np.int32),
               rew=np.zeros((n, max_steps), np.int32), val=np.zeros((n, max_steps), np.float32), pol=np.zeros((n, max_steps, A), np.float32),
               mask=np.zeros((n, max_steps), np.float32), player=np.zeros((n, max_steps), np.int32),
               team=np.full((n, max_steps), -1, np.int32), discount=np.zeros((n, max_steps), np.int32), idx=np.zeros(n, np.int32))
    if not det:
        buf["dice"] = np.zeros((n, max_steps), np.int32)
        buf["dice_dist"] = np.zeros((n, max_steps, 6), np.float32)
    key = np.asarray(rng_key, np.uint32)
    step = 0
    while step < max_steps and not state.done.all():
        keys = O.split(key, n + 1)
        key, step_keys = keys[0], keys[1:]
        live = state.done == 0
        if not det:  # throw_die only inside do_active_step
            thrown = state.copy()
            O.madn_cls_throw_die(thrown)
            state.die[live] = thrown.die[live]
            state.key[live] = thrown.key[live]
        obs = O.madn_det_encode_board(state) if det else O.madn_cls_encode_board(state)
        valid = (O.madn_det_valid_action(state) if det else O.madn_cls_valid_action(state)).reshape(n, -1)
        action, weights, value = search_fn(step_keys, obs, ~valid)
        stepped, skipped = state.copy(), state.copy()
        if det:
            r, d = O.madn_det_step(stepped, np.stack([action // 6, action % 6 + 1], 1).astype(np.int8))
            O.madn_det_no_step(skipped)
        else:
            r, d = O.madn_cls_step(stepped, action.astype(np.int8))
            O.madn_cls_no_step(skipped)
        if not det:
            dist_step, dist_skip = O.madn_cls_dice_probabilities(stepped), O.madn_cls_dice_probabilities(skipped)
        for g in range(n):
            if not live[g]:
                continue
            i = buf["idx"][g]
            has_valid = valid[g].any()
            pid = int(state.current_player[g])
            team = pid % 2 if teams else -1
            src = stepped if has_valid else skipped
            if has_valid:
                nxt, nd, rw = int(stepped.current_player[g]), bool(d[g]), int(r[g])
                rew_t = 2 if (nd and rw > 0) else (0 if (nd and rw < 0) else 1)
                same = (team == nxt % 2) if teams else (pid == nxt)
                disc_t = 1 if nd else (2 if same else 0)
                row = (obs[g], int(action[g]), rew_t, value[g], weights[g], 1.0, disc_t)
            else:
                row = (np.zeros_like(obs[g]), -1, 1, 0.0, np.zeros(A, np.float32), 0.0, 1)
            if i < max_steps:
                buf["obs"][g, i], buf["act"][g, i], buf["rew"][g, i], buf["val"][g, i] = row[0], row[1], row[2], row[3]
                buf["pol"][g, i], buf["mask"][g, i], buf["discount"][g, i] = row[4], row[5], row[6]
                buf["player"][g, i], buf["team"][g, i] = pid, team
                if not det:
                    buf["dice"][g, i] = state.die[g]
                    buf["dice_dist"][g, i] = (dist_step if has_valid else dist_skip)[g]
            buf["idx"][g] = i + 1
            for f, v in src.fields().items():
                getattr(state, f)[g] = v[g]
        step += 1
    return buf


def dog_raw_observation(state):
    """NumPy twin of DOG.dog.raw_observation (not a reference function: the reference has no DOG encoder)"""
    n, P = state.n, state.cfg.num_players
    own = state.hands[np.arange(n), state.current_player.astype(np.int64) % P]
    misc = np.stack([state.phase, state.hand_size, state.current_player, state.round_starter], 1)
    return np.concatenate([state.board, own, misc], 1).astype(np.int8)


def dog_encode_board(state):
    """NumPy twin of DOG.dog.encode_board (this repo's design, see include/dogstep.h: the reference has no DOG encoder) — a
    checker for the CUDA kernel, written from the plane list, not from the kernel"""
    cfg = state.cfg
    n, N, T, bs, d = state.n, cfg.num_players, cfg.total, 4 * cfg.distance, cfg.distance
    teams = bool(cfg.rules & 1)
    P = 8 + 3 * N + 14
    obs = np.zeros((n, P, T), np.int8)
    for g in range(n):
        cur = int(np.clip(state.current_player[g] + (N if state.current_player[g] < 0 else 0), 0, N - 1))
        board = state.board[g].astype(np.int64)
        rolled = np.concatenate([np.roll(board[:bs], -d * cur), np.roll(board[bs:bs + 16], -4 * cur)])
        rel = np.where(rolled < 0, -1, (rolled - cur) % N)
        for r in range(N):
            obs[g, r] = rel == r
        own = (rel >= 0) & ((rel % 2 == 0) if teams else (rel == 0))
        oth = (rel >= 0) & ((rel % 2 == 1) if teams else (rel != 0))
        obs[g, N], obs[g, N + 1] = own, oth
        for r in range(N):
            q = (cur + r) % N
            obs[g, N + 2 + r] = int((state.pins[g, q] == -1).sum())
            obs[g, 2 * N + 16 + r] = int(state.hands[g, q].sum())
        for c in range(14):
            obs[g, 2 * N + 2 + c] = state.hands[g, cur, c]
        m = 3 * N + 16
        rs = int(state.round_starter[g])
        obs[g, m], obs[g, m + 1] = state.phase[g], state.hand_size[g]
        obs[g, m + 2] = -1 if rs < 0 else (rs - cur) % N
        obs[g, m + 3] = int(state.swap_choices[g, cur]) + 1
    return obs


def play_batch_of_games_dog(state, max_steps, rng_key, search_fn, teams):
    """do_active_step of MuZero_det_MADN/game_agent.py:64-148 applied to the DOG env oracle (BASELINE config 5; the reference
    has no DOG self-play loop, MuZero_DOG/muzero_dog.py:85-99).  state: O.DogState."""
    n = state.n
    A = state.cfg.num_actions
    buf = dict(obs=np.zeros((n, max_steps, 74), np.float32), act=np.zeros((n, max_steps), np.int32),
               rew=np.zeros((n, max_steps), np.int32), val=np.zeros((n, max_steps), np.float32), pol=np.zeros((n, max_steps, A), np.float32),
               mask=np.zeros((n, max_steps), np.float32), player=np.zeros((n, max_steps), np.int32),
               team=np.full((n, max_steps), -1, np.int32), discount=np.zeros((n, max_steps), np.int32), idx=np.zeros(n, np.int32))
    key = np.asarray(rng_key, np.uint32)
    step = 0
    while step < max_steps and not state.done.all():
        keys = O.split(key, n + 1)
        key, step_keys = keys[0], keys[1:]
        live = state.done == 0
        obs = dog_raw_observation(state)
        valid = O.dog_valid_actions(state)
        action, weights, value = search_fn(step_keys, obs, ~valid)
        stepped, skipped = state.copy(), state.copy()
        r, d = O.dog_step(stepped, np.where(valid.any(1), action, 0))
        O.dog_no_step(skipped)
        for g in range(n):
            if not live[g]:
                continue
            i = buf["idx"][g]
            has_valid = valid[g].any()
            pid = int(state.current_player[g])
            team = pid % 2 if teams else -1
            src = stepped if has_valid else skipped
            if i < max_steps:
                buf["player"][g, i], buf["team"][g, i] = pid, team
                if has_valid:
                    nxt, nd, rw = int(stepped.current_player[g]), bool(d[g]), int(r[g])
                    same = (team == nxt % 2) if teams else (pid == nxt)
                    buf["obs"][g, i], buf["act"][g, i], buf["val"][g, i], buf["pol"][g, i] = obs[g], int(action[g]), value[g], weights[g]
                    buf["rew"][g, i] = 2 if (nd and rw > 0) else (0 if (nd and rw < 0) else 1)
                    buf["mask"][g, i], buf["discount"][g, i] = 1.0, (1 if nd else (2 if same else 0))
                else:
                    buf["act"][g, i], buf["rew"][g, i], buf["discount"][g, i] = -1, 1, 1
            buf["idx"][g] = i + 1
            for f, v in src.fields().items():
                getattr(state, f)[g] = v[g]
        step += 1
    return buf
