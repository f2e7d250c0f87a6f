"""mctx-shaped search API on libdogstep.so's per-game tree kernels.

Stands in for the three mctx 0.0.6 policies the reference calls (MuZero_det_MADN/muzero_deterministic_madn.py:673-684,
MuZero_Classic_MADN/muzero_classic_madn.py:488-501, TicTacToe/mcts.py:13-22,29-37) with the same names, keyword
arguments and PolicyOutput fields.  The networks stay the caller's: `recurrent_fn(params, rng_key, action, embedding)`
is any callable on CUDA tensors (a Flax function behind DLPack in the reference's setting, a torch module in the
tests) returning an object with .prior_logits/.value/.reward/.discount and the next embedding.  One simulation is
select kernel -> recurrent_fn on [games, E] -> expand+backup kernel.

Differences from mctx that are deliberate and documented: `rng_key` is a uint32 [games, 2] tensor of
per-game keys (the reference vmaps a B=1 search per game, game_agent.py:80), and the Dirichlet root-noise SAMPLE of
muzero_policy / stochastic_muzero_policy is an input (`dirichlet_noise`) instead of jax.random.dirichlet.
"""
import ctypes as C
import functools
from collections import namedtuple

import torch

from . import _lib

PolicyOutput = namedtuple("PolicyOutput", ["action", "action_weights", "search_tree"])
RootFnOutput = namedtuple("RootFnOutput", ["prior_logits", "value", "embedding"])
RecurrentFnOutput = namedtuple("RecurrentFnOutput", ["reward", "discount", "prior_logits", "value"])
DecisionRecurrentFnOutput = namedtuple("DecisionRecurrentFnOutput", ["chance_logits", "afterstate_value"])
ChanceRecurrentFnOutput = namedtuple("ChanceRecurrentFnOutput", ["action_logits", "value", "reward", "discount"])
SearchSummary = namedtuple("SearchSummary", ["visit_counts", "visit_probs", "value", "qvalues"])

MUZERO, GUMBEL, STOCHASTIC = 0, 1, 2
AUX_WORDS = 52  # words per node of dogstep_mcts_tree.select_aux (include/dogstep.h)


class _QT:
    def __init__(self, kind, **kw):
        self.kind, self.kw = kind, kw

    def __call__(self, **kw):  # functools.partial-style refinement: qtransform_by_min_max(min_value=-1, max_value=1)
        return _QT(self.kind, **{**self.kw, **kw})


qtransform_by_min_max = _QT(0, min_value=0.0, max_value=1.0)
qtransform_by_parent_and_siblings = _QT(1, epsilon=1e-8)
qtransform_completed_by_mix_value = _QT(2, value_scale=0.1, maxvisit_init=50.0, epsilon=1e-8)


def _resolve_qt(qt):
    if isinstance(qt, functools.partial):
        return qt.func(**qt.keywords)
    return qt


class Tree:
    """mctx.Tree-shaped view of the device buffers ([games, nodes, actions]).

    Wide Gumbel trees (DOG's 806 actions) are SPARSE on the device: the search initialises and reads only the root rows and the
    children that have visits (include/dogstep.h, dogstep_mcts_materialize).  Reading one of the dense per-child arrays or
    `embeddings` through this object first fills in the defaults, so what the caller sees always equals the dense mctx.Tree;
    the root-only accessors (summary(), qvalues(0)) and the search itself never pay for that."""

    ROOT_INDEX, NO_PARENT, UNVISITED = 0, -1, -1
    _DENSE = ("children_index", "children_prior_logits", "children_visits", "children_rewards", "children_discounts",
              "children_values", "embeddings")

    def __init__(self, n, num_simulations, num_actions, num_chance, embed_dim, policy, device):
        N, A = num_simulations + 1, num_actions + num_chance
        i32, f32 = torch.int32, torch.float32
        e = functools.partial(torch.empty, device=device)
        b = self._buf = {}
        self.node_visits, self.raw_values, self.node_values = e((n, N), dtype=i32), e((n, N), dtype=f32), e((n, N), dtype=f32)
        self.parents, self.action_from_parent = e((n, N), dtype=i32), e((n, N), dtype=i32)
        b["children_index"], b["children_visits"] = e((n, N, A), dtype=i32), e((n, N, A), dtype=i32)
        b["children_prior_logits"], b["children_rewards"] = e((n, N, A), dtype=f32), e((n, N, A), dtype=f32)
        b["children_discounts"], b["children_values"] = e((n, N, A), dtype=f32), e((n, N, A), dtype=f32)
        b["embeddings"] = e((n, N, embed_dim), dtype=f32)
        self.is_decision = e((n, N), dtype=torch.uint8) if policy == STOCHASTIC else None
        self.root_invalid_actions = e((n, A), dtype=torch.uint8)
        self.root_gumbel = e((n, A), dtype=f32) if policy == GUMBEL else None
        self.search_key, self.policy_key = e((n, 2), dtype=torch.uint32), e((n, 2), dtype=torch.uint32)
        self.path = torch.zeros((n, 65), dtype=i32, device=device)  # descent scratch for the parallel backup
        # per-node select cache of the wide Gumbel path (DOG's 806 actions): 144 B per node
        self.select_aux = (torch.zeros((n, N + 1, AUX_WORDS), dtype=torch.uint32, device=device)
                           if policy == GUMBEL and num_chance == 0 and 32 < A <= 832 else None)
        # stochastic: the two action indices the callbacks are evaluated with, written by select (no clamp launches per simulation)
        self.select_action_decision = e((n,), dtype=i32) if policy == STOCHASTIC else None
        self.select_action_chance = e((n,), dtype=i32) if policy == STOCHASTIC else None
        self.num_actions, self.num_chance, self.n = num_actions, num_chance, n
        self._cfg = None  # set by Search: the configuration the kernels run this tree with

    def __getattr__(self, name):
        # only reached for names that are not instance attributes: the dense per-child arrays
        if name in Tree._DENSE:
            buf = self.__dict__["_buf"]
            self.materialize()
            return buf[name]
        raise AttributeError(name)

    def raw(self, name):
        """the device buffer as the kernels see it (sparse for wide Gumbel trees: only root rows and visited children defined)"""
        return self._buf[name] if name in self._buf else getattr(self, name)

    def materialize(self):
        """fill in the rows / entries a sparse search never wrote (no-op for dense trees)"""
        cfg = self.__dict__.get("_cfg")
        if cfg is None or self.select_aux is None:
            return
        ct = self.cstruct()
        _lib.check(_lib.lib().dogstep_mcts_materialize(C.byref(ct), C.c_int64(self.n), C.byref(cfg), _lib.stream()), "mcts_materialize")

    def cstruct(self):
        return _lib.tag(_lib.MctsTree(*[None if self.raw(k) is None else C.c_void_p(self.raw(k).data_ptr())
                                        for k in _lib.MCTS_TREE_FIELDS]), self.node_visits.device)

    def qvalues(self, indices=0):
        src = self._buf if isinstance(indices, int) and indices == 0 else {k: getattr(self, k) for k in ("children_rewards", "children_discounts", "children_values")}
        return src["children_rewards"][:, indices] + src["children_discounts"][:, indices] * src["children_values"][:, indices]

    def summary(self):
        """mctx Tree.summary() restricted to the decision actions (root rows only: always defined)"""
        A = self.num_actions
        value = self.node_values[:, 0]
        visits = self._buf["children_visits"][:, 0, :A]
        vc = visits.to(torch.float32)
        tot = vc.sum(-1, keepdim=True)
        probs = torch.where(tot > 0, vc / tot.clamp(min=1), torch.full_like(vc, 1.0 / A))
        return SearchSummary(visit_counts=visits, visit_probs=probs, value=value, qvalues=self.qvalues(0)[:, :A])


def _cfg(policy, qtransform, num_simulations, max_depth, A, Cn, E, **kw):
    qt = _resolve_qt(qtransform)
    return _lib.MctsCfg(policy=policy, qtransform=qt.kind, num_simulations=num_simulations,
                        max_depth=num_simulations if max_depth is None else max_depth, num_actions=A, num_chance=Cn, embed_dim=E,
                        max_num_considered_actions=kw.get("max_num_considered_actions", 16),
                        q_min=qt.kw.get("min_value", 0.0), q_max=qt.kw.get("max_value", 1.0),
                        value_scale=qt.kw.get("value_scale", 0.1), maxvisit_init=qt.kw.get("maxvisit_init", 50.0),
                        epsilon=qt.kw.get("epsilon", 1e-8), pb_c_init=kw.get("pb_c_init", 1.25), pb_c_base=kw.get("pb_c_base", 19652.0),
                        dirichlet_fraction=kw.get("dirichlet_fraction", 0.25), temperature=kw.get("temperature", 1.0),
                        gumbel_scale=kw.get("gumbel_scale", 1.0))


class Search:
    """The explicit select / expand stepping API (what a jax.ffi caller drives once per simulation)."""

    def __init__(self, cfg, n, device="cuda"):
        self.cfg, self.n, self.device = cfg, n, torch.device(device)
        self.tree = Tree(n, cfg.num_simulations, cfg.num_actions, cfg.num_chance, cfg.embed_dim, cfg.policy, self.device)
        self.tree._cfg = cfg
        self._ct = self.tree.cstruct()
        self.parent = torch.empty(n, dtype=torch.int32, device=self.device)
        self.action = torch.empty(n, dtype=torch.int32, device=self.device)
        self.embedding = torch.empty((n, cfg.embed_dim), dtype=torch.float32, device=self.device)
        self.is_decision = torch.empty(n, dtype=torch.uint8, device=self.device)
        self.expand_key = torch.empty((n, 2), dtype=torch.uint32, device=self.device)

    def init(self, keys, root, invalid_actions=None, dirichlet_noise=None):
        f = lambda x: None if x is None else x.contiguous()
        inv = None if invalid_actions is None else invalid_actions.reshape(self.n, -1).to(torch.uint8).contiguous()
        _lib.check(_lib.lib().dogstep_mcts_init(C.byref(self._ct), C.c_int64(self.n), C.byref(self.cfg), _lib.ptr(keys.contiguous()),
                                               _lib.ptr(f(root.prior_logits.float())), _lib.ptr(f(root.value.float())),
                                               _lib.ptr(f(root.embedding.float().reshape(self.n, -1))), _lib.ptr(inv),
                                               _lib.ptr(f(dirichlet_noise)), _lib.stream()), "mcts_init")

    def select(self, sim):
        _lib.check(_lib.lib().dogstep_mcts_select(C.byref(self._ct), C.c_int64(self.n), C.byref(self.cfg), C.c_int32(sim),
                                                 _lib.ptr(self.parent), _lib.ptr(self.action), _lib.ptr(self.embedding),
                                                 _lib.ptr(self.is_decision), _lib.ptr(self.expand_key), _lib.stream()), "mcts_select")
        return self.parent, self.action, self.embedding, self.is_decision

    def expand(self, sim, prior_logits, value, reward, discount, embedding, chance_logits=None, afterstate_value=None,
               afterstate_embedding=None):
        f = lambda x: None if x is None else x.float().contiguous()
        _lib.check(_lib.lib().dogstep_mcts_expand(C.byref(self._ct), C.c_int64(self.n), C.byref(self.cfg), C.c_int32(sim),
                                                 _lib.ptr(self.parent), _lib.ptr(self.action), _lib.ptr(f(prior_logits)),
                                                 _lib.ptr(f(value)), _lib.ptr(f(reward)), _lib.ptr(f(discount)),
                                                 _lib.ptr(f(embedding)), _lib.ptr(f(chance_logits)), _lib.ptr(f(afterstate_value)),
                                                 _lib.ptr(f(afterstate_embedding)), _lib.stream()), "mcts_expand")

    def expand_select(self, sim, prior_logits, value, reward, discount, embedding, chance_logits=None, afterstate_value=None,
                      afterstate_embedding=None):
        """expand(sim) followed by select(sim + 1) in one launch; refreshes self.parent / action / embedding / is_decision /
        expand_key in place and returns them like select()"""
        f = lambda x: None if x is None else x.float().contiguous()
        emb, aemb = f(embedding), f(afterstate_embedding)
        if emb.data_ptr() == self.embedding.data_ptr():
            emb = emb.clone()
        if aemb is not None and aemb.data_ptr() == self.embedding.data_ptr():
            aemb = aemb.clone()
        _lib.check(_lib.lib().dogstep_mcts_expand_select(C.byref(self._ct), C.c_int64(self.n), C.byref(self.cfg), C.c_int32(sim),
                                                        _lib.ptr(self.parent), _lib.ptr(self.action), _lib.ptr(f(prior_logits)),
                                                        _lib.ptr(f(value)), _lib.ptr(f(reward)), _lib.ptr(f(discount)), _lib.ptr(emb),
                                                        _lib.ptr(f(chance_logits)), _lib.ptr(f(afterstate_value)), _lib.ptr(aemb),
                                                        _lib.ptr(self.embedding), _lib.ptr(self.is_decision), _lib.ptr(self.expand_key),
                                                        _lib.stream()), "mcts_expand_select")
        return self.parent, self.action, self.embedding, self.is_decision

    def policy_output(self):
        A = self.cfg.num_actions
        action = torch.empty(self.n, dtype=torch.int32, device=self.device)
        weights = torch.empty((self.n, A), dtype=torch.float32, device=self.device)
        value = torch.empty(self.n, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().dogstep_mcts_policy_output(C.byref(self._ct), C.c_int64(self.n), C.byref(self.cfg), _lib.ptr(action),
                                                        _lib.ptr(weights), _lib.ptr(value), _lib.stream()), "mcts_policy_output")
        return PolicyOutput(action=action, action_weights=weights, search_tree=self.tree), value


def _run(search, params, root, recurrent_fn, invalid_actions, keys, dirichlet_noise=None):
    search.init(keys, root, invalid_actions, dirichlet_noise)
    S = search.cfg.num_simulations
    _, action, emb, _ = search.select(0)
    for sim in range(S):
        out, nxt = recurrent_fn(params, search.expand_key, action, emb)  # action: int32 [games] (a valid torch index dtype)
        step = search.expand_select if sim + 1 < S else search.expand  # expand(sim) + select(sim + 1) fused into one launch
        step(sim, out.prior_logits, out.value, out.reward, out.discount, nxt.reshape(search.n, -1))
    return search.policy_output()[0]


def _tree_copy_(dst, src):
    """copy the tensors of a nested dict / list / tuple `src` into the same-shaped structure `dst` in place; False if the
    structures do not match (different keys, shapes or dtypes)"""
    if isinstance(dst, torch.Tensor):
        if not isinstance(src, torch.Tensor) or dst.shape != src.shape or dst.dtype != src.dtype:
            return False
        dst.copy_(src)
        return True
    if isinstance(dst, dict):
        return isinstance(src, dict) and dst.keys() == src.keys() and all(_tree_copy_(dst[k], src[k]) for k in dst)
    if isinstance(dst, (list, tuple)):
        return isinstance(src, (list, tuple)) and len(dst) == len(src) and all(_tree_copy_(a, b) for a, b in zip(dst, src))
    return dst is src or dst == src


class GraphCache:
    """Caller-owned, bounded cache of searches captured as CUDA graphs.

    A search is 2-3 small launches per simulation (select, the caller's recurrent function, expand); at the batch sizes of
    the reference's configurations that loop is launch-bound.  Passing the same `GraphCache()` to a policy function on
    every move makes the first call capture `init -> num_simulations x (select, recurrent_fn, expand) -> policy_output`
    into one CUDA graph and every later call with the same shapes replay it: inputs are copied into the captured buffers,
    the returned PolicyOutput tensors are the captured ones (consume them before the next call).  The recurrent function
    must be capturable (CUDA work on the current stream only, no host synchronisation).

    `params`: the captured kernels read the parameter tensors that were passed at capture time.  A later call with the SAME
    object (torch modules / tensors updated in place by the optimiser) replays as is.  A later call with a NEW object of the
    same structure (immutable pytrees replaced at every optimiser step, as Flax does) has its tensors copied into the captured
    ones before the replay — no re-capture, nothing leaks.  A structure mismatch re-captures.  At most `max_entries` graphs
    (with their trees) are kept, least recently used first out; clear() drops all."""

    def __init__(self, max_entries=4):
        from collections import OrderedDict
        self.entries, self.max_entries = OrderedDict(), max(1, int(max_entries))

    def clear(self):
        self.entries.clear()

    def run(self, key, inputs, fn, params=None):
        """inputs: dict name -> tensor or None; fn(static_inputs) -> PolicyOutput"""
        ent = self.entries.get(key)
        if ent is not None and ent[4] is not params and not _tree_copy_(ent[4], params):
            del self.entries[key]
            ent = None
        if ent is None:
            static = {k: (None if v is None else v.detach().clone().contiguous()) for k, v in inputs.items()}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture: library handles, allocator pools
                fn(static)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = fn(static)
            # `fn` is kept: its closure owns the Search whose work buffers (parent, action, embedding, keys) the captured
            # kernels write on every replay; `params` is kept because the captured kernels read its tensors
            ent = self.entries[key] = (static, graph, out, fn, params)
            while len(self.entries) > self.max_entries:
                self.entries.popitem(last=False)
        self.entries.move_to_end(key)
        static, graph, out = ent[0], ent[1], ent[2]
        for k, v in inputs.items():
            if v is not None:
                static[k].copy_(v)
        graph.replay()
        return out


def _cfg_key(cfg):
    return tuple(getattr(cfg, f) for f, _ in cfg._fields_)


def _run_cached(graph_cache, cfg, n, dev, params, root, recurrent_fn, invalid_actions, keys, dirichlet_noise):
    if graph_cache is None:
        return _run(Search(cfg, n, dev), params, root, recurrent_fn, invalid_actions, keys, dirichlet_noise)
    inputs = dict(keys=keys, prior=root.prior_logits.float(), value=root.value.float(), emb=root.embedding.float().reshape(n, -1),
                  invalid=None if invalid_actions is None else invalid_actions.reshape(n, -1).to(torch.uint8), noise=dirichlet_noise)
    key = (_cfg_key(cfg), n, str(dev), id(recurrent_fn), invalid_actions is None, dirichlet_noise is None)
    search = Search(cfg, n, dev) if key not in graph_cache.entries else None

    def fn(st):
        return _run(search, params, RootFnOutput(st["prior"], st["value"], st["emb"]), recurrent_fn, st["invalid"], st["keys"], st["noise"])

    return graph_cache.run(key, inputs, fn, params)


def muzero_policy(params, rng_key, root, recurrent_fn, num_simulations, invalid_actions=None, max_depth=None, *,
                  qtransform=qtransform_by_parent_and_siblings, dirichlet_fraction=0.25, dirichlet_alpha=0.3, pb_c_init=1.25,
                  pb_c_base=19652, temperature=1.0, dirichlet_noise=None, graph_cache=None):
    """mctx.muzero_policy.  rng_key: uint32 [games, 2].  graph_cache: optional GraphCache (CUDA-graph replay)."""
    n, A = root.prior_logits.shape
    E = root.embedding.reshape(n, -1).shape[1]
    if dirichlet_noise is None and dirichlet_fraction > 0:
        dirichlet_noise = torch.distributions.Dirichlet(torch.full((A,), float(dirichlet_alpha), device=root.value.device)).sample((n,))
    cfg = _cfg(MUZERO, qtransform, num_simulations, max_depth, A, 0, E, pb_c_init=pb_c_init, pb_c_base=pb_c_base,
               dirichlet_fraction=dirichlet_fraction, temperature=temperature)
    return _run_cached(graph_cache, cfg, n, root.value.device, params, root, recurrent_fn, invalid_actions, rng_key,
                       dirichlet_noise if dirichlet_fraction > 0 else None)


def gumbel_muzero_policy(params, rng_key, root, recurrent_fn, num_simulations, invalid_actions=None, max_depth=None, *,
                         qtransform=qtransform_completed_by_mix_value, max_num_considered_actions=16, gumbel_scale=1.0,
                         graph_cache=None):
    """mctx.gumbel_muzero_policy.  rng_key: uint32 [games, 2].  graph_cache: optional GraphCache (CUDA-graph replay)."""
    n, A = root.prior_logits.shape
    E = root.embedding.reshape(n, -1).shape[1]
    cfg = _cfg(GUMBEL, qtransform, num_simulations, max_depth, A, 0, E, max_num_considered_actions=max_num_considered_actions,
               gumbel_scale=gumbel_scale)
    return _run_cached(graph_cache, cfg, n, root.value.device, params, root, recurrent_fn, invalid_actions, rng_key, None)


def stochastic_muzero_policy(params, rng_key, root, decision_recurrent_fn, chance_recurrent_fn, num_simulations,
                             invalid_actions=None, max_depth=None, *, qtransform=qtransform_by_parent_and_siblings,
                             dirichlet_fraction=0.25, dirichlet_alpha=0.3, pb_c_init=1.25, pb_c_base=19652, temperature=1.0,
                             dirichlet_noise=None, num_chance_outcomes=None, graph_cache=None):
    """mctx.stochastic_muzero_policy: decision nodes (A actions) alternate with chance nodes (C outcomes).
    Both callbacks are evaluated every simulation (as mctx does) and the kernel picks per game by node type; the stored
    embedding is padded to the wider of the state / afterstate embeddings."""
    n, A = root.prior_logits.shape
    dev = root.value.device
    state_emb = root.embedding.reshape(n, -1).float()
    dummy_out, dummy_after = decision_recurrent_fn(params, None, torch.zeros(n, dtype=torch.long, device=dev), state_emb)
    Cn = dummy_out.chance_logits.shape[-1] if num_chance_outcomes is None else num_chance_outcomes
    Es, Ea = state_emb.shape[1], dummy_after.reshape(n, -1).shape[1]
    E = max(Es, Ea)
    pad = lambda x: torch.nn.functional.pad(x.reshape(n, -1).float(), (0, E - x.reshape(n, -1).shape[1]))
    if dirichlet_noise is None and dirichlet_fraction > 0:
        dirichlet_noise = torch.distributions.Dirichlet(torch.full((A,), float(dirichlet_alpha), device=dev)).sample((n,))
    cfg = _cfg(STOCHASTIC, qtransform, num_simulations, max_depth, A, Cn, E, pb_c_init=pb_c_init, pb_c_base=pb_c_base,
               dirichlet_fraction=dirichlet_fraction, temperature=temperature)
    cfg.state_embed_dim, cfg.afterstate_embed_dim = Es, Ea
    noise = dirichlet_noise if dirichlet_fraction > 0 else None

    def search_loop(s, st):
        s.init(st["keys"], RootFnOutput(st["prior"], st["value"], st["emb"]), st["invalid"], st["noise"])
        _, action, emb, is_dec = s.select(0)
        # the select kernel also writes the two indices the callbacks are evaluated with (decision action / chance outcome of the
        # selected edge, clamped like mctx's where()), and expand takes the state / afterstate embeddings at their own widths and
        # zero-fills the stored row: no clamp or pad launch per simulation on this side
        a_dec, a_ch = s.tree.select_action_decision, s.tree.select_action_chance
        for sim in range(num_simulations):
            dec, after = decision_recurrent_fn(params, None, a_dec, emb[:, :Es])
            ch, nxt = chance_recurrent_fn(params, None, a_ch, emb[:, :Ea])
            step = s.expand_select if sim + 1 < num_simulations else s.expand
            step(sim, ch.action_logits, ch.value, ch.reward, ch.discount, nxt.reshape(n, -1), dec.chance_logits, dec.afterstate_value,
                 after.reshape(n, -1))
        return s.policy_output()[0]

    inputs = dict(keys=rng_key, prior=root.prior_logits.float(), value=root.value.float(), emb=pad(state_emb),
                  invalid=None if invalid_actions is None else invalid_actions.reshape(n, -1).to(torch.uint8), noise=noise)
    if graph_cache is None:
        return search_loop(Search(cfg, n, dev), inputs)
    key = (_cfg_key(cfg), n, str(dev), id(decision_recurrent_fn), id(chance_recurrent_fn), invalid_actions is None, noise is None)
    s = Search(cfg, n, dev) if key not in graph_cache.entries else None
    return graph_cache.run(key, inputs, lambda st: search_loop(s, st), params)
