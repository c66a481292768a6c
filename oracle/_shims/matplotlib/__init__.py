"""Stub standing in for matplotlib (absent in this image) so that the reference package
imports.  TEST INFRASTRUCTURE ONLY -- used by oracle/refload.py when generating golden
vectors from /root/reference in the build container.  Swallows any attribute / call."""


class _Anything(object):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __getitem__(self, item):
        return _Anything()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Anything()
