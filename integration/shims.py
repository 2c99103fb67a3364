"""Stand-ins for the six pip packages the reference's ``tools/*.py`` / ``lib/core`` / ``lib/dataset`` / ``lib/utils/vis.py``
import but this image lacks (SURVEY.md section 8b "What else the tools need to start"): ``easydict``, ``vedo``,
``tensorboardX``, ``matplotlib``, ``json_tricks``, ``prettytable``.  TEST / DEMO HARNESS ONLY -- the backend itself
imports none of them.  Each stand-in implements only what the unmodified tools touch on the path exercised here
(logging scalars, an ASCII table, JSON load/dump); plotting calls are accepted and dropped (``DEBUG.DEBUG: false``).
"""
import json
import sys
import types


class EasyDict(dict):
    """Attribute-style dict (the behaviour ``lib/core/config.py`` relies on)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) else x for x in v)
        dict.__setitem__(self, k, v)
        object.__setattr__(self, k, v)

    __setitem__ = __setattr__


class _Anything:
    """Accepts any attribute access / call (plotting objects whose output nobody reads here)."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step=None):
        self.scalars.append((tag, float(value), step))

    def close(self):
        pass


class PrettyTable:
    def __init__(self):
        self.field_names, self.rows = [], []

    def add_row(self, row):
        self.rows.append(list(row))

    def __str__(self):
        cells = [[str(c) for c in self.field_names]] + [[str(c) for c in r] for r in self.rows]
        width = [max(len(r[i]) for r in cells if i < len(r)) for i in range(max(len(r) for r in cells))]
        return "\n".join(" | ".join(c.ljust(width[i]) for i, c in enumerate(r)) for r in cells)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    def fallback(attr):      # PEP 562: any other public name resolves to a do-nothing object
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = fallback
    return m


def install():
    """Register the stand-ins in ``sys.modules`` (only for packages that are really missing)."""
    def missing(name):
        if name in sys.modules:
            return False
        try:
            __import__(name)
            return False
        except ImportError:
            return True

    if missing("easydict"):
        sys.modules["easydict"] = _module("easydict", EasyDict=EasyDict)
    if missing("vedo"):
        sys.modules["vedo"] = _module("vedo")
    if missing("tensorboardX"):
        sys.modules["tensorboardX"] = _module("tensorboardX", SummaryWriter=SummaryWriter)
    if missing("prettytable"):
        sys.modules["prettytable"] = _module("prettytable", PrettyTable=PrettyTable)
    if missing("json_tricks"):
        sys.modules["json_tricks"] = _module("json_tricks", load=json.load, loads=json.loads, dump=json.dump,
                                             dumps=json.dumps)
    if missing("matplotlib"):
        mpl = _module("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _module("matplotlib.pyplot")
        mpl.colors = _module("matplotlib.colors")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot
        sys.modules["matplotlib.colors"] = mpl.colors
        tk = _module("mpl_toolkits")
        tk.mplot3d = _module("mpl_toolkits.mplot3d", Axes3D=_Anything)
        sys.modules["mpl_toolkits"] = tk
        sys.modules["mpl_toolkits.mplot3d"] = tk.mplot3d
