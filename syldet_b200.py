"""Import alias: `import syldet_b200` loads the package directory `syllable-detector-swift_b200/` (whose name is not a
valid Python identifier)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("syllable-detector-swift_b200")
sys.modules[__name__] = _pkg
