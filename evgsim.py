"""Import alias: the package directory is named ``everglades-ai-wargame_b200`` (not a valid Python
identifier), so ``import evgsim`` loads it from that directory under the name ``evgsim``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "everglades-ai-wargame_b200")
_spec = importlib.util.spec_from_file_location("evgsim", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["evgsim"] = _mod
_spec.loader.exec_module(_mod)
