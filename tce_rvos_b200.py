"""Import shim: the product package lives in the directory ``tce-rvos_b200/`` (hyphenated, as the
project layout requires), which Python cannot name in an ``import`` statement.  ``import
tce_rvos_b200`` loads that package and aliases it under this name."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("tce-rvos_b200")
sys.modules[__name__] = _pkg
