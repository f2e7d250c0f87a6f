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


class _DevPtr(C.c_void_p):
    """a device pointer that remembers which GPU it lives on (attribute `dev`)"""


class _Stream:
    """placeholder argument: replaced at call time by the current torch stream of the device the pointers live on"""


_STREAM = _Stream()


class _Entry:
    """One extern "C" entry point.  The kernels launch on the CUDA device that is current at call time, so the call is made
    with the device of its pointer arguments current and on torch's current stream OF THAT DEVICE: an env, tree or replay
    shard created with device='cuda:1' works whatever torch.cuda.current_device() is, and stays ordered with the torch ops
    on its tensors.  Pointers of two different GPUs in one call raise."""
    __slots__ = ("fn", "name")

    def __init__(self, fn, name):
        self.fn, self.name = fn, name

    def __call__(self, *args):
        dev = None
        for a in args:
            d = getattr(a, "dev", None)
            if d is None:
                o = getattr(a, "_obj", None)  # byref(struct): the struct carries the device of its leaves (tag())
                if o is not None:
                    d = getattr(o, "dev", None)
            if d is not None:
                if dev is None:
                    dev = d
                elif d != dev:
                    raise DogstepError(f"{self.name}: arguments live on different GPUs (cuda:{dev} and cuda:{d})")
        if dev is None and not any(a is _STREAM for a in args):
            return self.fn(*args)  # no device pointer and no stream placeholder: argument validation only (works without a GPU)
        if dev is None or dev == torch.cuda.current_device():
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            return self.fn(*[s if a is _STREAM else a for a in args])
        with torch.cuda.device(dev):
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            return self.fn(*[s if a is _STREAM else a for a in args])


class Prepared:
    """An entry point with its argument list built ONCE (structs, sizes, pointers): a call that is repeated every lockstep
    iteration on the same buffers then costs the C call plus one current-stream lookup instead of rebuilding ~8 ctypes
    objects and re-deriving the device.  `args` may be edited between calls (e.g. the host key slot)."""
    __slots__ = ("fn", "args", "stream_slots", "dev", "name")

    def __init__(self, entry, args):
        self.fn, self.name, self.args = entry.fn, entry.name, list(args)
        self.stream_slots = [i for i, a in enumerate(args) if a is _STREAM]
        dev = None
        for a in args:
            d = getattr(a, "dev", None)
            if d is None:
                o = getattr(a, "_obj", None)
                if o is not None:
                    d = getattr(o, "dev", None)
            if d is not None:
                if dev is not None and d != dev:
                    raise DogstepError(f"{self.name}: arguments live on different GPUs (cuda:{dev} and cuda:{d})")
                dev = d
        self.dev = dev

    def __call__(self):
        if self.dev is not None and self.dev != torch.cuda.current_device():
            with torch.cuda.device(self.dev):
                return self._go()
        return self._go()

    def _go(self):
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in self.stream_slots:
            self.args[i] = s
        return self.fn(*self.args)


class _Lib:
    def __init__(self, cdll):
        self._cdll = cdll
        cdll.dogstep_last_error.restype = C.c_char_p
        self.dogstep_last_error = cdll.dogstep_last_error

    def __getattr__(self, name):
        e = _Entry(getattr(self._cdll, name), name)  # AttributeError for a symbol the library does not export
        setattr(self, name, e)
        return e


def lib():
    """Load libdogstep.so; raise loudly when it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DogstepError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = _Lib(C.CDLL(LIB_PATH))
    return _lib


def tag(struct, device):
    """mark a ctypes struct of device pointers with the GPU its leaves live on (see _Entry)"""
    struct.dev = device.index if device.index is not None else torch.cuda.current_device()
    return struct


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
    p = _DevPtr(t.data_ptr())
    p.dev = t.device.index
    return p


def stream():
    """the `stream` argument of an entry point: torch's current stream on the device of the call's pointers"""
    return _STREAM


_U32x2 = C.c_uint32 * 2


def host_key(key):
    """uint32[2] host array from a tensor / ndarray / sequence (a jaxrand.KeyChain or a ctypes array is used as it is)"""
    if isinstance(key, _U32x2):
        return key
    buf = getattr(key, "buf", None)
    if isinstance(buf, _U32x2):
        return buf
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
                ("gumbel_scale", C.c_float), ("state_embed_dim", C.c_int32), ("afterstate_embed_dim", C.c_int32)]


MCTS_TREE_FIELDS = ("node_visits", "raw_values", "node_values", "parents", "action_from_parent", "children_index",
                    "children_prior_logits", "children_visits", "children_rewards", "children_discounts", "children_values",
                    "embeddings", "is_decision", "root_invalid_actions", "root_gumbel", "search_key", "policy_key", "path", "select_aux",
                    "select_action_decision", "select_action_chance")


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


class TttSearchScratch(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("parent", "action", "embedding", "expand_key", "prior_logits", "value", "reward", "discount",
                                          "next_embedding", "root_prior_logits", "root_value", "root_embedding")]


class TttState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("board", "current_player", "reward", "done", "memory")]
