"""The `sql.*` configuration keys the hot path honours (defaults as in dask_sql/sql.yaml:1-36).

dask.config is not available here; this module offers the same get / set-as-context-manager
behaviour for the keys the plugins read (context.py:519, join.py:228, aggregate.py:321).
"""
import contextlib
import os
import threading

_DEFAULTS = {
    "sql.aggregate.split_out": 1,
    "sql.aggregate.split_every": None,
    "sql.identifier.case_sensitive": True,
    "sql.join.broadcast": None,
    "sql.optimize": True,
    "sql.predicate_pushdown": True,
    "sql.dynamic_partition_pruning": True,
    "sql.optimizer.verbose": False,
}
_local = threading.local()


def _stack():
    if not hasattr(_local, "stack"):
        _local.stack = [dict(_DEFAULTS)]
        for k in list(_DEFAULTS):
            env = "DASK_" + k.upper().replace(".", "__")
            if env in os.environ:
                v = os.environ[env]
                _local.stack[0][k] = {"true": True, "false": False, "none": None}.get(v.lower(), v)
    return _local.stack


def get(key, default=None):
    cur = _stack()[-1]
    if key in cur:
        return cur[key]
    # prefix lookup: get("sql.aggregate") -> {"split_out":..., "split_every":...}
    pre = key + "."
    sub = {k[len(pre):]: v for k, v in cur.items() if k.startswith(pre)}
    if sub:
        return sub
    return default


@contextlib.contextmanager
def set(options=None, **kwargs):  # noqa: A001  (mirrors dask.config.set)
    new = dict(_stack()[-1])
    for k, v in {**(options or {}), **kwargs}.items():
        new[k] = v
    _stack().append(new)
    try:
        yield
    finally:
        _stack().pop()
