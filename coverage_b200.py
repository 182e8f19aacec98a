"""Import alias for the package directory `maximumareacoverageoptimization.jl_b200/` (its name
contains a dot, which Python's import statement cannot spell).  `import coverage_b200` gives the
package object; submodules are reachable as attributes and as `coverage_b200.<name>` imports."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maximumareacoverageoptimization.jl_b200")
_NAME = "coverage_b200"

_spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod
_spec.loader.exec_module(_mod)
