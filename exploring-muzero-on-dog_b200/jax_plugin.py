"""jax.ffi registration of libdogstep.so — the layer that lets the reference's jitted code call the CUDA path unchanged.

    from exploring_muzero_on_dog_b200 import jax_plugin
    jax_plugin.load("libdogstep_ffi.so")             # registers one XLA FFI target per entry point (platform "CUDA")
    env, reward, done = jax_plugin.det.env_step(env, action)        # same signature as MADN/deterministic_madn.py:170

The handlers are generated from include/dogstep.h (scripts/gen_ffi.py -> csrc/ffi/dogstep_ffi.cc, ffi_table.json); this module
is the Python half: `call()` turns a handler's table entry into a `jax.ffi.ffi_call` (operand order, result shapes,
input_output_aliases for every in/out buffer, attributes with their C types, vmap_method="broadcast_all" because the kernels are
natively batched over games), and `det` / `cls` / `dog` wrap it in functions with the reference's names and return tuples.  They
serve the call sites MuZero_det_MADN/game_agent.py:66-84,114 and evaluate_agent.py:331-350 (encode_board, valid_action, env_step,
no_step) and the dice / DOG twins.  A call made on un-vmapped leaves is a batch of one.

jax is not installable in the build container of this repo (no wheel, no network), so this module imports without it and
tests/test_ffi.py drives it against a recording stand-in for jax.ffi; on a machine with jax + jaxlib it works as is.
"""
import ctypes
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TABLE = {h["name"]: h for h in json.load(open(os.path.join(_HERE, "ffi_table.json")))["handlers"]}

RULE_BITS = {"enable_teams": 1 << 0, "enable_initial_free_pin": 1 << 1, "enable_circular_board": 1 << 2, "enable_start_blocking": 1 << 3,
             "enable_jump_in_goal_area": 1 << 4, "enable_friendly_fire": 1 << 5, "enable_start_on_1": 1 << 6,
             "enable_bonus_turn_on_6": 1 << 7, "must_traverse_start": 1 << 8, "enable_dice_rethrow": 1 << 9}  # include/dogstep_rules.h
_NP = {"int32_t": np.int32, "int64_t": np.int64, "uint32_t": np.uint32, "float": np.float32}

try:  # pragma: no cover - jax is absent in the build container
    import jax
    import jax.numpy as jnp
    _ffi = jax.ffi
except Exception:  # noqa: BLE001
    jax = jnp = _ffi = None

_loaded = None


def available():
    return _ffi is not None


def load(path=None):
    """dlopen libdogstep_ffi.so (built from csrc/ffi/dogstep_ffi.cc where jaxlib's headers are) and register every handler"""
    global _loaded
    if _ffi is None:
        raise RuntimeError("jax is not importable: jax_plugin needs jax >= 0.4.38 (jax.ffi) and jaxlib's XLA FFI headers")
    if _loaded is None:
        lib = ctypes.cdll.LoadLibrary(path or os.path.join(_HERE, "libdogstep_ffi.so"))
        for h in TABLE.values():
            _ffi.register_ffi_target(h["name"], _ffi.pycapsule(getattr(lib, h["symbol"])), platform="CUDA")
        _loaded = lib
    return _loaded


def rules_to_mask(rules):
    """the reference's static rules dict (MADN/deterministic_madn.py:38-40) -> the uint32 of include/dogstep_rules.h"""
    return sum(bit for k, bit in RULE_BITS.items() if rules.get(k, False))


def call(name, operands, attrs, xp=None):
    """One custom call.  operands: name -> array for every operand of the handler (None = absent optional leaf: an empty
    buffer); attrs: name -> Python scalar.  Returns name -> new array for every in/out operand."""
    h = TABLE[name]
    xp = xp or jnp
    ops, out_types, aliases, out_names = [], [], {}, []
    for i, o in enumerate(h["operands"]):
        a = operands.get(o["name"])
        if a is None:
            a = xp.zeros((0,), dtype=xp.uint8)
        ops.append(a)
        if o["mutable"]:
            aliases[i] = len(out_types)
            out_types.append(_shape_dtype(a))
            out_names.append(o["name"])
    missing = [a["name"] for a in h["attrs"] if a["name"] not in attrs]
    if missing:
        raise TypeError(f"{name}: missing attributes {missing}")
    typed = {a["name"]: _NP[a["type"]](attrs[a["name"]]) for a in h["attrs"]}
    outs = _ffi.ffi_call(name, out_types, input_output_aliases=aliases, vmap_method="broadcast_all")(*ops, **typed)
    return dict(zip(out_names, outs))


def _shape_dtype(a):
    return jax.ShapeDtypeStruct(a.shape, a.dtype) if jax is not None else (tuple(a.shape), a.dtype)


def _batched(x, unbatched_ndim):
    """(array with a leading game axis, was it added?)"""
    return (x, False) if x.ndim > unbatched_ndim else (x[None], True)


class _Env:
    """reference-shaped functions of one env module on top of call()"""

    def __init__(self, prefix, leaves, leaf_ndim, mask_shape, obs_rows):
        self.prefix, self.leaves, self.leaf_ndim, self.mask_shape, self.obs_rows = prefix, leaves, leaf_ndim, mask_shape, obs_rows

    def _state(self, env):
        ops, added = {}, False
        for k in self.leaves:
            v, a = _batched(getattr(env, k), self.leaf_ndim[k])
            if v.dtype == jnp.bool_:
                v = v.astype(jnp.uint8)
            ops["s_" + k], added = v, added or a
        n = ops["s_board"].shape[0]
        attrs = dict(n=n, cfg_num_players=int(env.num_players), cfg_layout_mask=0xF, cfg_distance=int(env.board_size) // 4,
                     cfg_rules=rules_to_mask(env.rules))
        return ops, attrs, n, added

    def _new_env(self, env, out, squeeze):
        upd = {}
        for k in self.leaves:
            v = out["s_" + k]
            if getattr(env, k).dtype == jnp.bool_:
                v = v.astype(jnp.bool_)
            upd[k] = v[0] if squeeze else v
        return env.replace(**upd)

    def valid_action(self, env):
        """valid_action (MADN/deterministic_madn.py:299, classic_madn.py:367) / valid_actions (DOG/dog.py:693)"""
        ops, attrs, n, sq = self._state(env)
        ops["mask"] = jnp.zeros((n,) + self.mask_shape(env), jnp.uint8)
        fn = self.prefix + ("_valid_actions" if self.prefix.endswith("dog") else "_valid_action")
        m = call(fn, ops, attrs)["mask"].astype(jnp.bool_)
        return m[0] if sq else m

    def env_step(self, env, action):
        """env_step -> (env, reward int8, done bool)  (deterministic_madn.py:170, classic_madn.py:257, dog.py:1117)"""
        ops, attrs, n, sq = self._state(env)
        det = self.prefix.endswith("det")
        act = jnp.asarray(action, jnp.int32 if self.prefix.endswith("dog") else jnp.int8)
        ops["action"] = act.reshape((n, 2) if det else (n,))
        ops["reward"], ops["done"] = jnp.zeros((n,), jnp.int8), jnp.zeros((n,), jnp.uint8)
        out = call(self.prefix + "_step", ops, attrs)
        r, d = out["reward"], out["done"].astype(jnp.bool_)
        return self._new_env(env, out, sq), (r[0] if sq else r), (d[0] if sq else d)

    def no_step(self, env):
        """no_step -> (env, 0, done)  (deterministic_madn.py:283, classic_madn.py:353, dog.py:713)"""
        ops, attrs, n, sq = self._state(env)
        ops["reward"], ops["done"] = jnp.zeros((n,), jnp.int8), jnp.zeros((n,), jnp.uint8)
        out = call(self.prefix + "_no_step", ops, attrs)
        r, d = out["reward"], out["done"].astype(jnp.bool_)
        return self._new_env(env, out, sq), (r[0] if sq else r), (d[0] if sq else d)

    def encode_board(self, env):
        """encode_board -> int8 [rows, total_board_size]  (deterministic_madn.py:395, classic_madn.py:463)"""
        ops, attrs, n, sq = self._state(env)
        ops["obs"] = jnp.zeros((n, self.obs_rows(env), int(env.total_board_size)), jnp.int8)
        o = call(self.prefix + "_encode_board", ops, attrs)["obs"]
        return o[0] if sq else o


_MADN_NDIM = {"board": 1, "current_player": 0, "pins": 2, "reward": 0, "done": 0, "action_set": 2, "die": 0, "key": 1}
det = _Env("dogstep_madn_det", ("board", "current_player", "pins", "reward", "done", "action_set", "key"), _MADN_NDIM,
           lambda env: (4, 6), lambda env: 8 * int(env.num_players) + 2)
cls = _Env("dogstep_madn_cls", ("board", "current_player", "pins", "reward", "done", "die", "key"), _MADN_NDIM,
           lambda env: (4,), lambda env: 2 * int(env.num_players) + 3)
dog = _Env("dogstep_dog", ("board", "current_player", "pins", "reward", "done", "deck", "hands", "swap_choices", "round_starter", "phase",
                           "key", "hand_size"),
           {"board": 1, "current_player": 0, "pins": 2, "reward": 0, "done": 0, "deck": 1, "hands": 2, "swap_choices": 1,
            "round_starter": 0, "phase": 0, "key": 1, "hand_size": 0},
           lambda env: (2 * (4 * (12 + 1 + int(env.total_board_size)) + 120) + 14,), None)
