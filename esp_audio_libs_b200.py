"""Import alias: the package directory is `esp-audio-libs_b200/` (not a valid Python
identifier), so `import esp_audio_libs_b200` loads it through importlib."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("esp-audio-libs_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
