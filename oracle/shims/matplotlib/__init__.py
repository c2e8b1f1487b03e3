"""Import shim (test infrastructure): lets the read-only reference import in a headless
container that has no matplotlib.  Every attribute is an inert callable/namespace."""
import sys, types


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: _Inert()  # PEP 562
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


pyplot = _mod("matplotlib.pyplot")
transforms = _mod("matplotlib.transforms", Bbox=_Inert)
ft2font = _mod("matplotlib.ft2font", BOLD=0)


def __getattr__(name):
    return _Inert()
