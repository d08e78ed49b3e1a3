"""YAML -> attribute dict, same call as the reference (utils/cfg.py:5-9).

`easydict` is not a dependency here; `AttrDict` gives the attribute access the
model code needs (config.model.volume_size ...).
"""
import yaml


class AttrDict(dict):
    def __init__(self, mapping=None, **kwargs):
        super().__init__()
        for k, v in dict(mapping or {}, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, list):
            return [AttrDict._wrap(x) for x in v]
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def load_config(path):
    with open(path) as fin:
        return AttrDict(yaml.safe_load(fin))


def default_config(batch_size=4, volume_size=64):
    """The reference's experiments/sceneego/test/sceneego.yaml (shipped as sceneego_b200/data/sceneego.yaml) with
    `opt.batch_size` and `model.volume_size` set -- what bench.py, smoke() and the tests construct the module from."""
    from .. import DEFAULT_CONFIG
    c = load_config(DEFAULT_CONFIG)
    c.opt.batch_size = batch_size
    c.model.volume_size = volume_size
    return c
