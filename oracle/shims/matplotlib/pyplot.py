"""Stub of matplotlib.pyplot: every attribute is a no-op callable (TEST INFRASTRUCTURE ONLY)."""


class _Nop:
    def __call__(self, *a, **k):
        return _Nop()

    def __getattr__(self, name):
        return _Nop()

    def __iter__(self):
        return iter((_Nop(), _Nop()))


def __getattr__(name):
    return _Nop()
