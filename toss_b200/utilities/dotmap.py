"""Minimal stand-in for `dotmap.DotMap` (not installed here): nested attribute-style dict.

The reference builds `args = DotMap(cfg, _dynamic=False)` (toss/utilities/load_default_cfg.py:12-14),
then reads and writes `args.section.key`.  If the real package is importable it is used instead.
"""
try:                                    # pragma: no cover - depends on the environment
    from dotmap import DotMap           # noqa: F401
except ImportError:

    class DotMap(dict):
        def __init__(self, *args, _dynamic=True, **kwargs):
            super().__init__()
            object.__setattr__(self, "_dynamic", _dynamic)
            for k, v in dict(*args, **kwargs).items():
                self[k] = self._wrap(v)

        @staticmethod
        def _wrap(v):
            if isinstance(v, dict) and not isinstance(v, DotMap):
                return DotMap(v, _dynamic=False)
            return v

        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            try:
                return self[k]
            except KeyError:
                if object.__getattribute__(self, "_dynamic"):
                    self[k] = DotMap()
                    return self[k]
                raise AttributeError(k) from None

        def __setattr__(self, k, v):
            self[k] = self._wrap(v)

        def __delattr__(self, k):
            del self[k]

        def toDict(self):
            return {k: (v.toDict() if isinstance(v, DotMap) else v) for k, v in self.items()}

        def __getstate__(self):
            return dict(self), object.__getattribute__(self, "_dynamic")

        def __setstate__(self, st):
            d, dyn = st
            object.__setattr__(self, "_dynamic", dyn)
            self.update(d)

        def __reduce__(self):
            return (DotMap, (), self.__getstate__(), None, None)

        def __deepcopy__(self, memo):
            import copy
            out = DotMap(_dynamic=object.__getattribute__(self, "_dynamic"))
            memo[id(self)] = out
            for k, v in self.items():
                dict.__setitem__(out, k, copy.deepcopy(v, memo))
            return out
