"""Import shim: makes the in-tree directory `climaocean.jl_b200/` importable as the dotted module
`climaocean.jl_b200` (a directory whose name contains a dot cannot be found by the default
path finder, so a one-entry meta-path finder maps the dotted name onto it)."""
import importlib.abc
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_pkg_dir = os.path.join(_root, "climaocean.jl_b200")
_NAME = "climaocean.jl_b200"


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != _NAME:
            return None
        return importlib.util.spec_from_file_location(_NAME, os.path.join(_pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[_pkg_dir])


if not any(isinstance(f, _Finder) for f in sys.meta_path):
    sys.meta_path.insert(0, _Finder())
