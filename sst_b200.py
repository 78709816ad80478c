"""Import shim: the package directory carries the repository's (hyphenated, un-importable) name
`emg-based-speech-recognition-with-heterogenous-data_b200/`; this module loads it as `sst_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "emg-based-speech-recognition-with-heterogenous-data_b200")


def _load():
    if "sst_b200" in sys.modules and getattr(sys.modules["sst_b200"], "__path__", None):
        return sys.modules["sst_b200"]
    spec = importlib.util.spec_from_file_location(
        "sst_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sst_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
