"""ConfigMixin / register_to_config restated: constructor kwargs are recorded in `self.config`."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        if not hasattr(self, "_internal_dict"):
            self._internal_dict = FrozenDict()
        self._internal_dict.update(kwargs)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        init.update(kwargs)
        return cls(**init)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = list(sig.parameters.items())[1:]
        cfg = {name: p.default for name, p in params if p.default is not inspect.Parameter.empty}
        for (name, _), a in zip(params, args):
            cfg[name] = a
        cfg.update(kwargs)
        if not isinstance(self, ConfigMixin):
            raise RuntimeError("register_to_config needs ConfigMixin")
        ConfigMixin.register_to_config(self, **cfg)
        init(self, *args, **kwargs)

    return inner
