"""Parameter access with the reference's surface (`/root/reference/src/2d/params.py`).

Same four entry points and the same observable behaviour:

* ``load_params(config_file=None)``  flat dict of every key of every section; the FIRST
  successful load is cached process-wide and later calls return it whatever
  ``config_file`` says (reference params.py:11,26-27,52).
* ``get_param(name, default=None)``  (params.py:56-68)
* ``get_param_group(group)``  re-reads the DEFAULT file on every call and ignores any
  file given to ``load_params``; ``{}`` for ``None`` / unknown groups (params.py:70-98).
* ``get_all_params()``  (params.py:100-107)

Errors: ``FileNotFoundError("Configuration file not found: ...")`` and
``ValueError("Invalid JSON in configuration file: ...")`` (params.py:40-43,92-95).

Additions (not in the reference): ``reset_cache()`` and ``load_config(path)`` which
returns the sectioned dict without touching the cache.
"""
from __future__ import annotations

import json
import os

_CACHE = None


def default_config_file():
    # <repo>/config/faghri.json -- the reference resolves it as three dirnames up from
    # src/2d/params.py (params.py:31-34); this package sits one level below the repo root.
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(os.path.dirname(here), 'config', 'faghri.json')


def _read(path):
    try:
        with open(path, 'r') as fh:
            return json.load(fh)
    except FileNotFoundError:
        raise FileNotFoundError(f"Configuration file not found: {path}")
    except json.JSONDecodeError:
        raise ValueError(f"Invalid JSON in configuration file: {path}")


def load_config(config_file=None):
    """Sectioned configuration (no caching)."""
    return _read(config_file or default_config_file())


def load_params(config_file=None):
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    sections = _read(default_config_file() if config_file is None else config_file)
    flat = {}
    for body in sections.values():
        flat.update(body)
    _CACHE = flat
    return flat


def get_param(name, default=None):
    return load_params().get(name, default)


def get_param_group(group_name):
    if group_name is None:
        return {}
    return _read(default_config_file()).get(group_name, {})


def get_all_params():
    return load_params()


def reset_cache():
    global _CACHE
    _CACHE = None
