"""Registers the package directory `the-algorithm_b200/` (not a valid Python identifier) as the importable
module `the_algorithm_b200`.  Used by tests/, bench.py and __graft_entry__.py:

    import _pkg; _pkg.load()
    from the_algorithm_b200.ann.brute_force import BruteForceIndex
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

NAME = "the_algorithm_b200"
ROOT = Path(__file__).resolve().parent / "the-algorithm_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, ROOT / "__init__.py", submodule_search_locations=[str(ROOT)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
