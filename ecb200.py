"""Importable alias of the package directory `rustcrypto-elliptic-curves_b200/` (the hyphenated name the
layout contract asks for cannot be written in an `import` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("rustcrypto-elliptic-curves_b200")
sys.modules[__name__] = _pkg
