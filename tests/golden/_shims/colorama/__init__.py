"""Minimal stand-in for `colorama` so the reference imports in the golden-vector
generator (tests/golden/make_golden.py).  Test infrastructure only."""


class _Codes:
    def __getattr__(self, name):
        return ""


Fore = _Codes()
Back = _Codes()
Style = _Codes()


def init(*args, **kwargs):
    return None
