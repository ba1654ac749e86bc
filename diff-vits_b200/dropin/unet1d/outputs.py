"""Minimal output container with attribute + index access (reference: unet1d/outputs.py:36)."""
from collections import OrderedDict


class BaseOutput(OrderedDict):
    def __init__(self, **kwargs):
        super().__init__()
        for k, v in kwargs.items():
            if v is not None:
                self[k] = v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return tuple(self.values())[k]

    def to_tuple(self):
        return tuple(self.values())
