"""Importable alias of the package directory ``sharkshark-4k_b200`` (a hyphen cannot appear in an
``import`` statement):  ``import ss4k_b200``  ==  ``importlib.import_module("sharkshark-4k_b200")``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("sharkshark-4k_b200")
sys.modules[__name__] = _pkg
