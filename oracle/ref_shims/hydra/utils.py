"""TEST INFRASTRUCTURE ONLY -- `hydra.utils.instantiate` as the reference uses it (models/encoder_processor_decoder.py:69-106):
import the class named by `_target_`, call it with the remaining config keys merged with the call's keyword arguments
(nested configs that carry their own `_target_` are instantiated first)."""
import importlib


def _build(value):
    if isinstance(value, dict) and "_target_" in value:
        return instantiate(value)
    if isinstance(value, dict):
        return {k: _build(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return type(value)(_build(v) for v in value)
    return value


def instantiate(config, *args, **kwargs):
    cfg = dict(config)
    target = cfg.pop("_target_")
    module_name, _, attr = target.rpartition(".")
    cls = getattr(importlib.import_module(module_name), attr)
    params = {k: _build(v) for k, v in cfg.items() if not (k.startswith("_") and k.endswith("_"))}
    params.update(kwargs)
    return cls(*args, **params)
