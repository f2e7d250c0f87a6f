"""`mctx` stand-in: only the callback output containers the env modules construct.  The search policies
themselves are NOT provided (mctx is a pinned third-party dependency whose algorithm oracle/mcts_oracle.c restates)."""
import collections

RootFnOutput = collections.namedtuple("RootFnOutput", ["prior_logits", "value", "embedding"])
RecurrentFnOutput = collections.namedtuple("RecurrentFnOutput", ["reward", "discount", "prior_logits", "value"])
DecisionRecurrentFnOutput = collections.namedtuple("DecisionRecurrentFnOutput", ["chance_logits", "afterstate_value"])
ChanceRecurrentFnOutput = collections.namedtuple("ChanceRecurrentFnOutput", ["action_logits", "value", "reward", "discount"])


def _unavailable(*a, **k):
    raise NotImplementedError("jaxshim does not implement mctx search policies")


muzero_policy = gumbel_muzero_policy = stochastic_muzero_policy = _unavailable


class qtransforms:
    qtransform_by_min_max = qtransform_by_parent_and_siblings = qtransform_completed_by_mix_value = staticmethod(_unavailable)
