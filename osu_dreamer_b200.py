"""Import shim: the package lives in the directory `osu-dreamer_b200/` (hyphenated as the project name is),
which is not a valid Python identifier; importing `osu_dreamer_b200` loads it from there."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'osu-dreamer_b200')
_spec = _ilu.spec_from_file_location('osu_dreamer_b200', _os.path.join(_dir, '__init__.py'),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules['osu_dreamer_b200'] = _mod
_spec.loader.exec_module(_mod)
