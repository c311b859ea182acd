"""Import stub; every attribute is a no-op callable."""


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
