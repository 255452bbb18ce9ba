"""The repo root carries one-line `models.*` shims at the reference's import paths (drop-in entry points); the fixture
generators need the REFERENCE's own `models` package instead.  It has no __init__.py (a namespace package), so a regular
package of the same name anywhere on sys.path would win: bind the name explicitly."""
import sys
import types

REF = "/root/reference"


def use_reference_models():
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    m = types.ModuleType("models")
    m.__path__ = [REF + "/models"]
    sys.modules["models"] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
