"""Import-only stubs for nitorch.io (unires/_util.py:4); I/O is out of scope."""


def map(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('nitorch.io is out of scope')


def savef(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('nitorch.io is out of scope')
