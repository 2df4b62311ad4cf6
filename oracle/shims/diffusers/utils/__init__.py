"""diffusers.utils names used by the reference (restated; test infrastructure)."""
import dataclasses
import logging as _pylogging
from collections import OrderedDict

WEIGHTS_NAME = "diffusion_pytorch_model.bin"


class BaseOutput(OrderedDict):
    """Dict + attribute access.  Works both for @dataclass subclasses (UNet3DConditionOutput) and for the
    reference's plain annotated subclasses constructed with keyword arguments (stablemtl_pipeline.py:32-109)."""

    def __post_init__(self):
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if v is not None:
                super().__setitem__(f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        if name in self.keys() and value is not None:
            super().__setitem__(name, value)
        super().__setattr__(name, value)

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        super().__setattr__(key, value)

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name=None):
        return _pylogging.getLogger(name)


logging = _Logging()
from . import import_utils  # noqa: E402,F401
