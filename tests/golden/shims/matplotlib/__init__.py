"""TEST INFRASTRUCTURE: import stand-in for matplotlib (absent from this image).  The reference's environment modules
import it at module level for their plotting helpers; nothing on the control path calls into it."""
import types


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        return _Anything(name)

    def __call__(self, *a, **k):
        return _Anything("call")


def use(*a, **k):
    pass
