"""TicTacToe goldens from the REFERENCE's own sources (/root/reference/TicTacToe/TicTacToe{,V2}.py on the jaxshim):
env_step trajectories (legal, illegal and post-terminal moves), policy_function, and the true-env mctx callbacks
root_fn / recurrent_fn (rollout values for given keys).      python tests/golden/gen_ttt_goldens.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
sys.path.insert(0, "/root/reference/TicTacToe")

import chex  # noqa: E402
import dataclasses  # noqa: E402

chex.dataclass = lambda cls: dataclasses.dataclass(cls)  # chex.dataclass -> plain dataclass under the shim
import jax  # noqa: E402
import jax.numpy as jnp  # noqa: E402
import TicTacToe as v1  # noqa: E402
import TicTacToeV2 as v2  # noqa: E402


def leaves(env, variant):
    mem = np.asarray(env.memory) if variant == 1 else np.full((2, 3), -1, np.int8)
    return np.asarray(env.board).copy(), int(env.current_player), int(env.reward), bool(env.done), mem.copy()


def main():
    rng = np.random.default_rng(5)
    out = {}
    for variant, mod in ((0, v1), (1, v2)):
        boards, curs, rews, dones, mems, acts, pols = [], [], [], [], [], [], []
        rkeys, rvals, cact, cprior, cval, crew, cdisc, cb2, cc2, cr2, cd2, cm2 = ([] for _ in range(12))
        for game in range(40):
            env = mod.env_reset(0)
            for t in range(14):
                b, c, r, d, m = leaves(env, variant)
                boards.append(b); curs.append(c); rews.append(r); dones.append(d); mems.append(m)
                pols.append(np.asarray(mod.policy_function(env)).astype(np.float32))
                if game % 4 == 0 and t < 6 and not d:
                    key = jax.random.split(jax.random.PRNGKey(int(rng.integers(1 << 30))))[0]
                    root = mod.root_fn(env, key)
                    a = int(rng.integers(9))
                    ro, env2 = mod.recurrent_fn(None, key, jnp.int8(a), env)
                    rkeys.append(np.asarray(key)); rvals.append(float(root.value)); cact.append(a)
                    cprior.append(np.asarray(ro.prior_logits).astype(np.float32)); cval.append(float(ro.value))
                    crew.append(float(ro.reward)); cdisc.append(float(ro.discount))
                    b2, c2, r2, d2, m2 = leaves(env2, variant)
                    cb2.append(b2); cc2.append(c2); cr2.append(r2); cd2.append(d2); cm2.append(m2)
                else:
                    rkeys.append(np.zeros(2, np.uint32)); rvals.append(np.nan); cact.append(-1)
                    cprior.append(np.zeros(9, np.float32)); cval.append(np.nan); crew.append(np.nan); cdisc.append(np.nan)
                    cb2.append(b); cc2.append(c); cr2.append(r); cd2.append(d); cm2.append(m)
                free = np.flatnonzero(b.reshape(-1) == 0)
                a = int(rng.integers(9)) if (rng.random() < 0.25 or free.size == 0) else int(rng.choice(free))
                acts.append(a)
                env, _, _ = mod.env_step(env, jnp.int8(a))
            b, c, r, d, m = leaves(env, variant)
            boards.append(b); curs.append(c); rews.append(r); dones.append(d); mems.append(m)
            pols.append(np.asarray(mod.policy_function(env)).astype(np.float32)); acts.append(-1)
            rkeys.append(np.zeros(2, np.uint32)); rvals.append(np.nan); cact.append(-1); cprior.append(np.zeros(9, np.float32))
            cval.append(np.nan); crew.append(np.nan); cdisc.append(np.nan); cb2.append(b); cc2.append(c); cr2.append(r); cd2.append(d); cm2.append(m)
        tag = f"v{variant}"
        out.update({f"{tag}_board": np.array(boards, np.int8), f"{tag}_cur": np.array(curs, np.int8), f"{tag}_reward": np.array(rews, np.int8),
                    f"{tag}_done": np.array(dones), f"{tag}_memory": np.array(mems, np.int8), f"{tag}_action": np.array(acts, np.int8),
                    f"{tag}_policy": np.array(pols, np.float32), f"{tag}_key": np.array(rkeys, np.uint32), f"{tag}_root_value": np.array(rvals, np.float32),
                    f"{tag}_rec_action": np.array(cact, np.int32), f"{tag}_rec_prior": np.array(cprior, np.float32), f"{tag}_rec_value": np.array(cval, np.float32),
                    f"{tag}_rec_reward": np.array(crew, np.float32), f"{tag}_rec_discount": np.array(cdisc, np.float32),
                    f"{tag}_rec_board": np.array(cb2, np.int8), f"{tag}_rec_cur": np.array(cc2, np.int8), f"{tag}_rec_rew": np.array(cr2, np.int8),
                    f"{tag}_rec_done": np.array(cd2), f"{tag}_rec_memory": np.array(cm2, np.int8)})
    np.savez_compressed(os.path.join(HERE, "ttt_reference.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_board")})


if __name__ == "__main__":
    main()
