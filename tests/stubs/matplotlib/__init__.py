"""Test stand-in for matplotlib (not installed in this image): absorbs every call the reference's callers and the
product's plotting layer make, draws nothing.  Only tests/test_reference_callers.py puts this directory on sys.path."""
import sys
import types


class _Anything:
    """Callable, subscriptable, iterable-as-empty object whose every attribute is another one of itself."""

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _any_attribute(attr):
    """PEP 562 module __getattr__: any public name resolves; dunder probes (__file__, __path__, ... as the import
    machinery and inspect make them on every entry of sys.modules) fail like on a plain module."""
    if attr.startswith("__") and attr.endswith("__"):
        raise AttributeError(attr)
    return _Anything()


def _module(name):
    m = types.ModuleType(name)
    m.__getattr__ = _any_attribute
    return m


def use(*a, **k):
    return None


__getattr__ = _any_attribute


for _sub in ("pyplot", "tri", "colors", "cm"):
    _m = _module(f"matplotlib.{_sub}")
    sys.modules[f"matplotlib.{_sub}"] = _m
    globals()[_sub] = _m
# plt.subplots(...) is unpacked into (fig, axes) by callers
sys.modules["matplotlib.pyplot"].subplots = lambda *a, **k: (_Anything(), _Anything())
