"""Import alias: the package directory is `operator-builder_b200/` (hyphen, as the brief names it);
this module loads it under the importable name `operator_builder_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "operator-builder_b200")
_spec = importlib.util.spec_from_file_location("operator_builder_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["operator_builder_b200"] = _mod
_spec.loader.exec_module(_mod)
