"""No-op stand-in for matplotlib (absent from the image): lets the reference's validation drivers run
headless in tools/run_reference_drivers.sh.  Every attribute is a callable that returns the same
object, so `plt.figure(); plt.plot(...); plt.savefig(...)` do nothing.  Tooling only."""
import sys as _sys
import types as _types


class _Nop:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter((self, self))

    def __getitem__(self, i):
        return self


_nop = _Nop()


class _Mod(_types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _nop


for _n in ("pyplot", "animation", "colors", "widgets", "cm", "patches"):
    _m = _Mod(f"matplotlib.{_n}")
    _sys.modules[f"matplotlib.{_n}"] = _m
    globals()[_n] = _m


def use(*a, **k):
    return None
