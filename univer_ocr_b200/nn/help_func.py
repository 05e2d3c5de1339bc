"""Argument helpers with the reference's semantics (nn/help_func.py:4-31)."""
from collections.abc import Iterable


def make_list_if_not(var):
    return var if isinstance(var, list) else [var]


def tuplize(name, var, length):
    """int -> (var,)*length; iterable of `length` ints -> tuple.  Negative -> ValueError,
    anything else -> TypeError (same exceptions as the reference)."""
    if isinstance(var, bool):
        values = None
    elif isinstance(var, int):
        values = (var,) * length
    elif isinstance(var, Iterable):
        values = tuple(var)
        if len(values) != length or not all(isinstance(v, int) and not isinstance(v, bool) for v in values):
            values = None
    else:
        values = None
    if values is None:
        raise TypeError(f'{name} must be either int or iterable of ints of length {length}, '
                        f'found {type(var).__name__}')
    if any(v < 0 for v in values):
        raise ValueError(f'{name} cannot be negative, found: {var}')
    return values
