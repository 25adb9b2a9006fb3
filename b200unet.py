"""Import alias: `import b200unet` loads the package whose directory name (mandated by the build
contract) is not a valid Python identifier."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
PACKAGE = "one-stop-for-covid-19-infection-and-lung-segmentation-plus-classification_b200"
_pkg = importlib.import_module(PACKAGE)
for _sub in ("layers", "graphs", "plan", "losses"):
    importlib.import_module(PACKAGE + "." + _sub)


def load(sub):
    """b200unet.load('model') -> the package submodule (imports torch lazily for the heavy ones)."""
    return importlib.import_module(PACKAGE + "." + sub)


sys.modules[__name__].__dict__.update({k: v for k, v in _pkg.__dict__.items() if not k.startswith("__")})
