"""dogstep — B200-native self-play hot path behind the Exploring-MuZero-on-DOG Python API.

Sub-modules mirror the reference's module names for the hot path:
    MADN.deterministic_madn, MADN.classic_madn, DOG.dog, TicTacToe, mcts, vec_replay_buffer
Everything computes in libdogstep.so (hand-written CUDA, sm_100a) through the C-ABI in
include/dogstep.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
