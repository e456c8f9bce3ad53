"""Passive containers standing in for gymnasium.spaces (see package docstring)."""


class Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Dict(dict):
    def __init__(self, spaces=None):
        super().__init__(spaces or {})
