"""Import alias: the package directory is `ppo-radiotherapy_b200/` (a hyphen is not legal in an
import statement), so `import ppo_radiotherapy_b200` resolves to that package."""
import importlib
import sys

_REAL = "ppo-radiotherapy_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
