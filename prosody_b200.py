"""Import shim: the product package lives in ``prosody-control-french-tts_b200/`` (not an importable name);
``import prosody_b200`` loads it under this name."""
import importlib.util
import pathlib
import sys

_pkg = pathlib.Path(__file__).resolve().parent / "prosody-control-french-tts_b200"
_spec = importlib.util.spec_from_file_location("prosody_b200", _pkg / "__init__.py", submodule_search_locations=[str(_pkg)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["prosody_b200"] = _mod
_spec.loader.exec_module(_mod)
