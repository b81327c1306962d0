"""Import shim: the package directory is `kafka-assigner_b200/` (hyphen, as the project is named);
this module makes it importable as `kafka_assigner_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kafka-assigner_b200")
_spec = importlib.util.spec_from_file_location("kafka_assigner_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["kafka_assigner_b200"] = _mod
_spec.loader.exec_module(_mod)
