"""Only `isinstance(y, nap.TsdFrame)` is evaluated on the path (core.py:459, :601)."""


class TsdFrame:
    def __init__(self, d=None, t=None, **kw):
        self.d, self.t = d, t


class Tsd:
    def __init__(self, d=None, t=None, **kw):
        self.d, self.t = d, t


class IntervalSet:
    pass
