"""Pins the oracle's jax.random restatement: Random123 KATs + the split/uniform values printed in
JAX's own documentation for partitionable threefry."""
import numpy as np

import oracle as O


def test_threefry_random123_kats():
    for k0, k1, c0, c1, e0, e1 in [
        (0, 0, 0, 0, 0x6B200159, 0x99BA4EFE),
        (0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0x1CB996FC, 0xBB002BE7),
        (0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3, 0xC4923A9C, 0x483DF7A0),
    ]:
        o0, o1 = O.threefry2x32(k0, k1, [c0], [c1])
        assert (int(o0[0]), int(o1[0])) == (e0, e1)


def test_split_matches_documented_values():
    assert O.split(O.prng_key(0)).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]
    assert O.split(O.prng_key(42)).tolist() == [[1832780943, 270669613], [64467757, 2916123636]]


def test_uniform_documented_values():
    assert abs(float(O.uniform(O.prng_key(0), 1)[0]) - 0.947667) < 1e-6
    assert abs(float(O.uniform(O.prng_key(42), 1)[0]) - 0.48870957) < 1e-7


def test_randint_range_and_determinism():
    k = O.split(O.prng_key(0))[1]
    a = O.randint(k, 4096, 0, 1_000_000)
    assert a.min() >= 0 and a.max() < 1_000_000 and len(np.unique(a)) > 4000
    assert np.array_equal(a, O.randint(k, 4096, 0, 1_000_000))


def test_integer_categorical_equals_float_gumbel_argmax():
    """The CUDA path picks the first valid action with the largest 23-bit mantissa; this must equal
    argmax(logits + gumbel) evaluated literally in float (libm logf) — checked on 20k draws."""
    rng = np.random.default_rng(0)
    for t in range(20000):
        key = rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
        mask = rng.integers(0, 2, 24).astype(np.uint8)
        if not mask.any():
            continue
        assert O.categorical_masked(key, mask, False) == O.categorical_masked(key, mask, True)


def test_choice6_distribution():
    p = np.full(6, 1 / 6, np.float32)
    rng = np.random.default_rng(1)
    cnt = np.zeros(6)
    for _ in range(6000):
        cnt[O.choice6(rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32), p)] += 1
    assert cnt.min() > 800


def test_reference_notebook_known_answers():
    """Known answers recorded by the reference itself: MADN/jupyter_code/test_functions.ipynb cell 1 prints, for step(env, s)
    with s = 1, 2, 3 on states whose dice_probabilities are uniform, "Die throw: 4 / 1 / 6" and, for s = 3 with four valid
    actions, "Chosen action: 2".  step (MADN/simulate_classicMADN.py:112-141) draws
        rng_key, sub = split(PRNGKey(s));  die = choice(sub, [1..6], p)          (classic_madn.py:238-242)
        rng_key, sub = split(rng_key);     idx = randint(sub, (), 0, N)           (:136-137)
    so these four values pin split + choice(p) + randint of the restatement (and of the shim the goldens were made on)
    to real jax.random output."""
    p = np.full(6, 1 / 6, np.float32)
    recorded = {1: 4, 2: 1, 3: 6}
    for s, die in recorded.items():
        k1, sub = O.split(O.prng_key(s))
        assert O.choice6(sub, p) + 1 == die
        _, sub2 = O.split(k1)
        assert int(O.randint(sub2, 1, 0, 1)[0]) == 0  # N = 1 rows of the same printout
    _, sub2 = O.split(O.split(O.prng_key(3))[0])
    assert int(O.randint(sub2, 1, 0, 4)[0]) == 2

    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "jaxshim"))
    try:
        for m in [m for m in sys.modules if m == "jax" or m.startswith("jax.")]:
            del sys.modules[m]
        import jax
        import jax.numpy as jnp

        faces = jnp.array([1, 2, 3, 4, 5, 6], dtype=jnp.int8)
        for s, die in recorded.items():
            k1, sub = jax.random.split(jax.random.PRNGKey(s))
            assert int(jax.random.choice(sub, faces, p=jnp.ones(6) / 6)) == die
        _, sub2 = jax.random.split(jax.random.split(jax.random.PRNGKey(3))[0])
        assert int(jax.random.randint(sub2, (), 0, 4)) == 2
    finally:
        sys.path.pop(0)
        for m in [m for m in sys.modules if m == "jax" or m.startswith("jax.")]:
            del sys.modules[m]
