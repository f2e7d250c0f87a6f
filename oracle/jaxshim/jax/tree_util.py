from ._core import tree_map, tree_leaves  # noqa: F401
