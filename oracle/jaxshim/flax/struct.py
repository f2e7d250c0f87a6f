"""`flax.struct` stand-in: frozen dataclass with .replace(); `pytree_node=False` kept as field metadata."""
import dataclasses


def field(pytree_node=True, **kw):
    md = dict(kw.pop("metadata", {}) or {})
    md["pytree_node"] = pytree_node
    return dataclasses.field(metadata=md, **kw)


def dataclass(cls=None, **kw):
    def wrap(c):
        c = dataclasses.dataclass(frozen=True, eq=False)(c)

        def replace(self, **updates):
            return dataclasses.replace(self, **updates)

        c.replace = replace
        return c

    return wrap(cls) if cls is not None else wrap


class PyTreeNode:
    pass
