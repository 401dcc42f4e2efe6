"""Minimal stand-in for the `nvtx` package (not installed in this image).

The reference hard-imports `nvtx` (reference src/SPART/SPART.py:23) and only uses
`nvtx.annotate(...)` as a context manager / decorator.  This stub lets the fixture
generators under tools/ import the unmodified reference; it is never imported by the
product or by the tests.
"""
import contextlib
import functools


class annotate(contextlib.ContextDecorator):
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def push_range(*a, **k):
    return None


def pop_range(*a, **k):
    return None
