"""Importable alias of the package directory `exploring-muzero-on-dog_b200/`.

The directory name mirrors the upstream repo name and contains a hyphen, which Python cannot
import directly; this module turns itself into that package (sets __path__ and runs its
__init__) so `import exploring_muzero_on_dog_b200.MADN.deterministic_madn` works.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "exploring-muzero-on-dog_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
