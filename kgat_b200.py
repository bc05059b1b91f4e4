"""Importable alias of the package directory ``problem-recommender-system-using-kgat-in-codeforces_b200/``
(its name follows the reference repository and is not a valid Python identifier).

    import kgat_b200
    from kgat_b200.model import KGAT, KGATArgs, KGATMode      # drop-in for src.model.KGAT.model
"""

import importlib as _importlib
import sys as _sys

_REAL = "problem-recommender-system-using-kgat-in-codeforces_b200"
_pkg = _importlib.import_module(_REAL)
for _name, _mod in list(_sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        _sys.modules[__name__ + _name[len(_REAL):]] = _mod
