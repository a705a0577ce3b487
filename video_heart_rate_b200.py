"""Importable alias for the package directory ``video-heart-rate_b200/`` (a hyphen is not a
valid Python identifier): ``import video_heart_rate_b200`` loads that directory as a package
under this name, sub-modules included."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "video-heart-rate_b200")
_spec = importlib.util.spec_from_file_location(
    "video_heart_rate_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["video_heart_rate_b200"] = _mod
_spec.loader.exec_module(_mod)
