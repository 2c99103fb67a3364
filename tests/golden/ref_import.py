"""Import the reference's ``lib/`` packages in THIS container (golden generation only).

``/root/reference`` is not available on the GPU box, so nothing under ``tests/``
that runs there imports this file; it is used by ``make_golden.py`` alone.
Two import-time-only dependencies are missing from the image and are shimmed:
``easydict`` (``lib/core/config.py:15``) and ``vedo``
(``lib/models/cuboid_proposal_net_soft.py:14``).
"""
import sys
import types

REF_ROOT = "/root/reference"


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) else x for x in v)
        dict.__setitem__(self, k, v)
        object.__setattr__(self, k, v)

    __setitem__ = __setattr__


def install():
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    if "vedo" not in sys.modules:
        m = types.ModuleType("vedo")
        m.Volume = m.show = None
        sys.modules["vedo"] = m
    lib = REF_ROOT + "/lib"
    if lib not in sys.path:
        sys.path.insert(0, lib)
