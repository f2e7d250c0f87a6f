from typing import Any

Array = Any
PRNGKey = Any
Numeric = Any
ArrayTree = Any
Scalar = Any
