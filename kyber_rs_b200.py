"""Import alias: the package directory is literally `kyber-rs_b200/` (not a valid identifier)."""
import importlib
import sys

_pkg = importlib.import_module("kyber-rs_b200")
sys.modules[__name__] = _pkg
