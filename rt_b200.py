"""Import shim: the package directory is named `raytracing-course-hw-public_b200` (after the reference
repository), which is not a Python identifier.  `import rt_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raytracing-course-hw-public_b200")
_NAME = "rt_b200"

if not (_NAME in sys.modules and getattr(sys.modules[_NAME], "__path__", None)):
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
