"""ctypes binding of libdogstep.so (the C-ABI declared in include/dogstep.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, this module
raises.  PyTorch is used only as the owner of device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DOGSTEP_LIB") or os.path.join(_HERE, "libdogstep.so")  # DOGSTEP_LIB: an alternative build (profiling)
_lib = None

ERRORS = {-1: "invalid argument", -2: "unsupported configuration", -3: "CUDA failure"}


class DogstepError(RuntimeError):
    pass


class MadnCfg(C.Structure):
    _fields_ = [("num_players", C.c_int32), ("layout_mask", C.c_int32), ("distance", C.c_int32), ("rules", C.c_uint32)]


class MadnDetState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("board", "current_player", "pins", "reward", "done", "action_set", "key")]


class MadnClsState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("board", "current_player", "pins", "reward", "done", "die", "key")]


def lib():
    """Load libdogstep.so; raise loudly when it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DogstepError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.dogstep_last_error.restype = C.c_char_p
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dogstep_last_error().decode() if rc == -3 else ""
        raise DogstepError(f"{what} failed: {ERRORS.get(rc, rc)} {msg}")


def ptr(t):
    """device pointer of a CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise DogstepError("dogstep kernels need CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise DogstepError("dogstep kernels need contiguous tensors")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_key(key):
    """uint32[2] host array from a tensor / ndarray / sequence"""
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().tolist()
    k = [int(x) & 0xFFFFFFFF for x in list(key)]
    return (C.c_uint32 * 2)(k[0], k[1])


class DogState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("board", "current_player", "pins", "reward", "done", "deck", "hands",
                                          "swap_choices", "round_starter", "phase", "key", "hand_size")]


class MctsCfg(C.Structure):
    _fields_ = [("policy", C.c_int32), ("qtransform", C.c_int32), ("num_simulations", C.c_int32), ("max_depth", C.c_int32),
                ("num_actions", C.c_int32), ("num_chance", C.c_int32), ("embed_dim", C.c_int32),
                ("max_num_considered_actions", C.c_int32), ("q_min", C.c_float), ("q_max", C.c_float),
                ("value_scale", C.c_float), ("maxvisit_init", C.c_float), ("epsilon", C.c_float), ("pb_c_init", C.c_float),
                ("pb_c_base", C.c_float), ("dirichlet_fraction", C.c_float), ("temperature", C.c_float),
                ("gumbel_scale", C.c_float)]


MCTS_TREE_FIELDS = ("node_visits", "raw_values", "node_values", "parents", "action_from_parent", "children_index",
                    "children_prior_logits", "children_visits", "children_rewards", "children_discounts", "children_values",
                    "embeddings", "is_decision", "root_invalid_actions", "root_gumbel", "search_key", "policy_key", "path", "select_aux")


class MctsTree(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in MCTS_TREE_FIELDS]


class ReplayArrays(C.Structure):
    _fields_ = [("capacity", C.c_int32), ("max_episode_length", C.c_int32), ("obs_size", C.c_int32), ("action_dim", C.c_int32),
                ("obs_is_int8", C.c_int32), ("stochastic", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("observations", "actions", "rewards", "root_values", "child_visits", "masks", "players",
                                          "teams", "discounts", "episode_lengths", "dice_outcomes", "dice_distributions")]


class ReplayBatch(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("observations", "actions", "rewards", "policies", "values", "masks", "target_values",
                                          "discount_targets", "dice_outcomes", "dice_probs")]


class TttState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("board", "current_player", "reward", "done", "memory")]
