"""Stand-ins for plotting packages the reference drivers import but this image lacks (matplotlib): every attribute is a
callable that returns another stub, `subplots()` returns a (figure, axes) pair.  Test harness only."""
import sys
import types


class _Anything:
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


class _Pyplot(types.ModuleType):
    def __init__(self):
        super().__init__("matplotlib.pyplot")

    def subplots(self, *a, **k):
        return _Anything(), _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):          # inspect / importlib probe modules for __file__, __path__, ...
            raise AttributeError(name)
        return _Anything()


def install_matplotlib_stub():
    """Returns the list of module names inserted (so the caller can remove them again)."""
    try:
        import matplotlib.pyplot  # noqa: F401
        return []
    except ImportError:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.pyplot = _Pyplot()
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    return ["matplotlib", "matplotlib.pyplot"]
