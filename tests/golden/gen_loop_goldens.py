"""Goldens for the LOOPS around the env, produced by the reference's own loop code (build container only):

    python tests/golden/gen_loop_goldens.py     ->  tests/golden/loops_reference.npz

  eval_det / eval_cls    play_n_games_for_eval_jitted + play_eval_loop_jitted of MuZero_det_MADN/evaluate_agent.py:715-927 and
                         MuZero_Classic_MADN/evaluate_agent_stochastic.py:719-932 with seats played by the random policy
                         (params['type'] == 3) and the rule-based scorer (type 2): do_random, do_rule_based, do_no_step,
                         manual_get_winner, the key chain, the winners bookkeeping
  selfplay_det / _cls    play_n_games_v3 + play_batch_of_games_jitted of MuZero_det_MADN/game_agent.py:50-192 and
                         MuZero_Classic_MADN/game_agent_stochastic.py:52-244: everything AROUND the search — key chain, throw_die,
                         encode_board, legal mask, env_step / no_step, reward / discount class targets, skipped turns, the
                         trajectory buffers — with run_muzero_mcts / run_stochastic_muzero_mcts replaced by a deterministic
                         stand-in (an integer hash of the key the loop hands to the search), because mctx and the Flax networks
                         are not installable here.  tests/ replays the same stand-in through oracle/selfplay_oracle.py and
                         through the CUDA loop.

The reference modules cannot be imported as they are: both evaluate_agent files run a tournament at import time and all four
import the Flax network module.  So the network module is replaced in sys.modules by a stub that exports the stand-in search,
and for the evaluate_agent files only the imports, function definitions and the RULES / batch_* assignments are executed
(selected from the parsed source; nothing is copied into this repo).  Executed on oracle/jaxshim, whose jax.random.categorical
uses the float contract of DESIGN.md for the two float32 logs of the Gumbel noise (XLA's own logf polynomial is not
reproducible here; a faithful logf differs by at most 1 ulp, which can only matter on an exact near-tie of two scores).
"""
import ast
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
sys.path.insert(0, "/root/reference")

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402

KEEP_ASSIGN = {"RULES", "batch_reset", "batch_valid_action", "batch_encode", "batch_env_step", "batch_map_action", "batch_throw_die"}


# ---- the stand-in search: same arithmetic as tests/test_selfplay_gpu.py::host_fn ------------------------------------------
class _PolicyOutput:
    def __init__(self, action, action_weights):
        self.action, self.action_weights, self.search_tree = action, action_weights, None


def _hash_search(key, invalid, A):
    k = np.asarray(key).astype(np.int64).reshape(2)
    inv = np.asarray(invalid).astype(bool).reshape(A)
    h = int((k[0] * 7 + k[1]) % 1000003)
    score = (h + np.arange(A, dtype=np.int64) * 40503) % 1009
    score[inv] = -1
    action = np.int32(score.argmax())
    w = np.where(inv, 0.0, 1.0).astype(np.float32)
    w = (w / np.float32(max(w.sum(), 1))).astype(np.float32)
    value = np.float32((h % 2001) - 1000) / np.float32(1024.0)
    return action, w, value


def run_muzero_mcts(params, rng_key, observations, invalid_actions, num_simulations, max_depth, temperature):
    a, w, v = _hash_search(rng_key, invalid_actions, 24)
    return _PolicyOutput(jnp.array([a]), jnp.array(w[None])), jnp.array([v])


def run_stochastic_muzero_mcts(params, rng_key, observations, invalid_actions, num_simulations, max_depth, temperature):
    a, w, v = _hash_search(rng_key, invalid_actions, 4)
    return _PolicyOutput(jnp.array([a]), jnp.array(w[None])), jnp.array([v])


def _stub_network_modules():
    for name, fn in (("MuZero_det_MADN.muzero_deterministic_madn", "run_muzero_mcts"),
                     ("MuZero_Classic_MADN.muzero_classic_madn", "run_stochastic_muzero_mcts")):
        m = types.ModuleType(name)
        setattr(m, fn, globals()[fn])
        # the loop files rely on names the star import of the network module leaks (functools, jax, jnp, ...)
        import functools
        import chex
        m.functools, m.jax, m.jnp, m.np, m.chex = functools, jax, jnp, np, chex
        sys.modules[name] = m


def _functions_of(path):
    """execute the imports, defs and the KEEP_ASSIGN assignments of a reference file (not its script part) -> namespace"""
    tree = ast.parse(open(path).read())
    body, seen = [], set()
    for n in tree.body:
        if isinstance(n, (ast.Import, ast.ImportFrom, ast.FunctionDef, ast.ClassDef)):
            body.append(n)
        elif isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name) and n.targets[0].id in KEEP_ASSIGN \
                and n.targets[0].id not in seen:
            seen.add(n.targets[0].id)
            body.append(n)
    body.sort(key=lambda n: (0 if isinstance(n, (ast.Import, ast.ImportFrom)) else 1 if (isinstance(n, ast.Assign) and n.targets[0].id == "RULES") else 2, n.lineno))
    ns = {"__name__": "reference_functions", "__file__": path}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def _leaves(envs, names):
    return {k: np.asarray(getattr(envs, k)) for k in names}


DET_LEAVES = ("board", "current_player", "pins", "reward", "done", "action_set", "key")
CLS_LEAVES = ("board", "current_player", "pins", "reward", "done", "die", "key")


def gen_eval(out):
    for tag, path, leaves in (("eval_det", "/root/reference/MuZero_det_MADN/evaluate_agent.py", DET_LEAVES),
                              ("eval_cls", "/root/reference/MuZero_Classic_MADN/evaluate_agent_stochastic.py", CLS_LEAVES)):
        ns = _functions_of(path)
        out[f"{tag}_rules"] = np.frombuffer(json.dumps({k: bool(v) for k, v in ns["RULES"].items()}).encode(), dtype=np.uint8)
        for ci, types_ in enumerate(((2, 3, 2, 3), (3, 2, 3, 3))):
            params = [{"type": jnp.int32(t)} for t in types_]
            rng_key = jax.random.PRNGKey(100 + ci)
            num_envs = 5
            # play_n_games_for_eval_jitted (:715-731), kept in pieces so that the final envs can be recorded too
            rng_key, subkey = jax.random.split(rng_key)
            seeds = jax.random.randint(subkey, (num_envs * 4,), 0, 1000000)
            envs = ns["batch_reset"](seeds, jnp.repeat(jnp.arange(4), num_envs))
            final_envs, winners = ns["play_eval_loop_jitted"](envs, tuple(params), subkey, num_envs * 4)
            out[f"{tag}_{ci}_types"] = np.array(types_, np.int32)
            out[f"{tag}_{ci}_seeds"] = np.asarray(seeds).astype(np.int32)
            out[f"{tag}_{ci}_key"] = np.asarray(subkey).astype(np.uint32)
            out[f"{tag}_{ci}_winners"] = np.asarray(winners).astype(np.int32)
            for k, v in _leaves(final_envs, leaves).items():
                out[f"{tag}_{ci}_final_{k}"] = v
            print(tag, ci, "winners", np.asarray(winners).sum(0).tolist(), "done", int(np.asarray(final_envs.done).sum()), flush=True)


def gen_selfplay(out):
    import importlib
    for tag, mod, leaves, shape in (("selfplay_det", "MuZero_det_MADN.game_agent", DET_LEAVES, (34, 56)),
                                    ("selfplay_cls", "MuZero_Classic_MADN.game_agent_stochastic", CLS_LEAVES, (11, 56))):
        ga = importlib.import_module(mod)
        out[f"{tag}_rules"] = np.frombuffer(json.dumps({k: bool(v) for k, v in ga.RULES.items()}).encode(), dtype=np.uint8)
        for ci, (num_envs, max_steps) in enumerate(((12, 150), (6, 700))):
            rng_key = jax.random.PRNGKey(200 + ci)
            # play_n_games_v3 (:185-192), kept in pieces so that the final envs can be recorded too
            rng_key, subkey = jax.random.split(rng_key)
            seeds = jax.random.randint(subkey, (num_envs,), 0, 1000000)
            envs = ga.batch_reset(seeds)
            # the jitted function does not return the envs: run it on a copy for the buffers, and once more, unrolled by hand
            # below, nothing else is needed — the final envs follow from the recorded actions (the tests recompute them)
            buf = ga.play_batch_of_games_jitted(envs, num_envs, shape, None, subkey, 16, 8, max_steps, 1.0)
            out[f"{tag}_{ci}_seeds"] = np.asarray(seeds).astype(np.int32)
            out[f"{tag}_{ci}_key"] = np.asarray(subkey).astype(np.uint32)
            out[f"{tag}_{ci}_max_steps"] = np.int32(max_steps)
            for k, v in buf.items():
                a = np.asarray(v)
                out[f"{tag}_{ci}_buf_{k}"] = a.astype(np.int8) if k == "obs" else a
            print(tag, ci, "idx", np.asarray(buf["idx"]).tolist(), flush=True)


def main():
    _stub_network_modules()
    out = {}
    gen_selfplay(out)
    gen_eval(out)
    np.savez_compressed(os.path.join(HERE, "loops_reference.npz"), **out)
    print("saved", os.path.getsize(os.path.join(HERE, "loops_reference.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
