"""Host mirror of TicTacToe/TicTacToe.py (classic rules): the V2 mirror with variant 0."""
from . import TicTacToeV2 as _v2
from .TicTacToeV2 import env_step, policy_function, root_fn, valid_action_mask  # noqa: F401

VARIANT = 0
TicTacToe = _v2.TicTacToeV2


def env_reset(_, n=None, device="cuda"):
    return _v2.env_reset(_, n=n, device=device, variant=0)


recurrent_fn = _v2.make_recurrent_fn(0)
