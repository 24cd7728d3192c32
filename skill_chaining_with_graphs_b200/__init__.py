"""Import alias for the package directory `skill-chaining-with-graphs_b200/` (a hyphenated name is
not a Python identifier).  `import skill_chaining_with_graphs_b200 as scg` loads that directory's
__init__.py; submodules resolve there too."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "skill-chaining-with-graphs_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
