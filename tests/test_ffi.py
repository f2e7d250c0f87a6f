"""The jax.ffi registration layer (VERDICT r1 "missing" #2): csrc/ffi/dogstep_ffi.cc + ffi_table.json are generated from
include/dogstep.h, compile against a stand-in for jaxlib's xla/ffi/api/ffi.h (every handler's signature is statically checked
against its binding), can be driven through fake call frames, and jax_plugin.py builds the ffi_call a jitted caller needs.
jax itself is not installable here, so the Python half runs against a recording stand-in for jax.ffi."""
import dataclasses
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "exploring-muzero-on-dog_b200")
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_generated_files_are_up_to_date_and_cover_the_header():
    import gen_ffi
    cc, js, handlers, protos = gen_ffi.generate()
    assert open(gen_ffi.OUT_CC).read() == cc, "run python scripts/gen_ffi.py"
    assert open(gen_ffi.OUT_JSON).read() == js, "run python scripts/gen_ffi.py"
    with_stream = [n for n, params in protos if any(p == ("void*", "stream") for p in params)]
    assert sorted(h["name"] for h in handlers) == sorted(with_stream) and len(handlers) >= 50
    host_only = sorted(set(n for n, _ in protos) - set(with_stream))
    assert host_only == ["dogstep_dog_num_actions", "dogstep_dog_obs_planes", "dogstep_host_key_chain", "dogstep_host_split", "dogstep_madn_cls_embed_dim", "dogstep_madn_det_embed_dim",
                         "dogstep_mcts_is_sparse", "dogstep_version"]


def test_shim_compiles_against_the_stand_in_header_and_handlers_run(tmp_path):
    import __graft_entry__ as g
    g.build()
    exe = str(tmp_path / "ffi_driver")
    cmd = ["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "tests", "ffi_stub"), "-I" + os.path.join(ROOT, "include"),
           "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "ffi_stub", "driver.cc"), "-L" + PKG, "-ldogstep",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + PKG, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout + r.stderr


def test_shim_refuses_to_build_without_the_xla_header(tmp_path):
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                        os.path.join(PKG, "csrc", "ffi", "dogstep_ffi.cc")], capture_output=True, text=True)
    assert r.returncode != 0 and "xla/ffi/api/ffi.h not found" in r.stderr


class _FakeFfi:
    """records what jax_plugin asks jax.ffi for"""

    def __init__(self):
        self.calls = []

    def ffi_call(self, name, out_types, input_output_aliases=None, vmap_method=None):
        def run(*ops, **attrs):
            self.calls.append(dict(name=name, ops=ops, attrs=attrs, aliases=dict(input_output_aliases), out_types=out_types, vmap=vmap_method))
            return [np.array(ops[i]) for i, _ in sorted(input_output_aliases.items(), key=lambda kv: kv[1])]
        return run


def _fake_jax(plugin):
    fake = _FakeFfi()
    plugin._ffi = fake
    plugin.jnp = types.SimpleNamespace(zeros=np.zeros, asarray=np.asarray, uint8=np.uint8, int8=np.int8, int32=np.int32, bool_=np.bool_)
    plugin.jax = None
    return fake


@dataclasses.dataclass(frozen=True)
class _DetEnv:  # the leaves and static fields of deterministic_MADN (MADN/deterministic_madn.py:24-40)
    board: np.ndarray
    current_player: np.ndarray
    pins: np.ndarray
    reward: np.ndarray
    done: np.ndarray
    action_set: np.ndarray
    key: np.ndarray
    num_players: int = 4
    board_size: int = 40
    total_board_size: int = 56
    rules: dict = dataclasses.field(default_factory=dict)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


def _env(batch=None):
    b = () if batch is None else (batch,)
    return _DetEnv(board=np.full(b + (56,), -1, np.int8), current_player=np.zeros(b, np.int8), pins=np.full(b + (4, 4), -1, np.int8),
                   reward=np.zeros(b, np.int8), done=np.zeros(b, np.bool_), action_set=np.full(b + (4, 6), 4, np.int8),
                   key=np.zeros(b + (2,), np.uint32), rules={"enable_teams": True, "enable_bonus_turn_on_6": True})


def test_plugin_builds_the_custom_call_of_env_step():
    from exploring_muzero_on_dog_b200 import jax_plugin
    fake = _fake_jax(jax_plugin)
    env = _env(batch=5)
    new_env, reward, done = jax_plugin.det.env_step(env, np.zeros((5, 2), np.int8))
    c = fake.calls[-1]
    h = jax_plugin.TABLE["dogstep_madn_det_step"]
    assert c["name"] == "dogstep_madn_det_step" and c["vmap"] == "broadcast_all"
    assert len(c["ops"]) == len(h["operands"]) == 10                     # 7 leaves, action, reward, done
    assert c["aliases"] == {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 8: 7, 9: 8}   # everything but the read-only action is updated in place
    assert c["attrs"]["n"] == 5 and c["attrs"]["cfg_distance"] == 10 and c["attrs"]["cfg_rules"] == (1 << 0) | (1 << 7)
    assert type(c["attrs"]["cfg_rules"]) is np.uint32 and type(c["attrs"]["n"]) is np.int64
    assert new_env.board.shape == (5, 56) and new_env.done.dtype == np.bool_ and reward.shape == (5,) and done.dtype == np.bool_
    # an un-vmapped env is a batch of one and comes back without the game axis
    env1, r1, d1 = jax_plugin.det.env_step(_env(), np.array([0, 1], np.int8))
    assert fake.calls[-1]["attrs"]["n"] == 1 and env1.pins.shape == (4, 4) and r1.shape == ()
    m = jax_plugin.det.valid_action(_env(batch=3))
    assert m.shape == (3, 4, 6) and fake.calls[-1]["name"] == "dogstep_madn_det_valid_action"
    o = jax_plugin.det.encode_board(_env(batch=3))
    assert o.shape == (3, 34, 56)


def test_plugin_table_matches_the_ctypes_mirror():
    """the same entry points the shipped ctypes mirror calls exist as FFI targets"""
    import re
    from exploring_muzero_on_dog_b200 import jax_plugin
    used = set()
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py") and f != "jax_plugin.py":
                used |= set(re.findall(r"\.(dogstep_[a-z0-9_]+)\b", open(os.path.join(dirpath, f)).read()))
    used -= {"dogstep_last_error", "dogstep_host_split", "dogstep_host_key_chain", "dogstep_mcts_is_sparse", "dogstep_dog_num_actions", "dogstep_dog_obs_planes",
             "dogstep_madn_det_embed_dim", "dogstep_madn_cls_embed_dim"}
    assert used and used <= set(jax_plugin.TABLE), sorted(used - set(jax_plugin.TABLE))
