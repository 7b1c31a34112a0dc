"""TEST INFRASTRUCTURE: empty stand-in (see matplotlib/__init__.py)."""


def __getattr__(name):
    def _noop(*a, **k):
        raise RuntimeError("matplotlib is not available here; plotting is not part of the recorded path")
    return _noop
