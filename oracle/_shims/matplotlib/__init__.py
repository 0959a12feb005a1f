"""Import-time stand-in for matplotlib (absent in this image).

Only used by oracle/gen_golden.py so that `/root/reference/baler` can be imported:
the reference imports its plotting module at load time (helper.py:31), but the
hot path never draws anything.
"""


def use(*_a, **_k):
    return None


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def __getattr__(name):
    return _Anything()
