"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (/root/reference) through
the inert shims in oracle/shims so that its own functions can be executed in THIS
container to (a) pin the restatements in oracle/ and (b) generate tests/golden/*.npz.

/root/reference does not exist on the GPU box: nothing under tests -m gpu, smoke() or
bench.py imports this module.  The product (radarslampy_b200/) never imports oracle/.
"""
import contextlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("RADARFE_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "parseData.py"))


def activate():
    """Put shims + reference on sys.path (shims first).  Idempotent."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _SHIMS)


@contextlib.contextmanager
def quiet():
    """The reference prints on every call; silence it while timing / generating."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def ref_module(name):
    """Import one reference module by its flat name (e.g. 'parseData')."""
    activate()
    import importlib
    with quiet():
        return importlib.import_module(name)
