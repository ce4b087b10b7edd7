"""Shared constants and the singleton metaclass of the facade.

The reference's ptina/common.py is mostly Taichi DSL glue (out of scope); what the hot path needs from it are the
constants `eps`/`inf` (common.py:32-33) and `Singleton` (common.py:407-413), which gives every pool class its
"construct once, fetch anywhere" behaviour (`ModelPool()` always returns the same object).
"""
import numpy as np  # noqa: F401  (re-exported: reference modules do `from ptina.common import *`)

eps = 1e-6
inf = 1e6


class Singleton(type):
    """`Cls()` constructs on first call and returns that instance afterwards (arguments of later calls are ignored)."""
    _instance = None

    def __call__(cls, *args, **kwargs):
        if cls._instance is None:
            cls._instance = super().__call__(*args, **kwargs)
        return cls._instance

    def _forget(cls):
        cls._instance = None
