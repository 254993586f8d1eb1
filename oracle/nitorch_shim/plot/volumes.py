"""Import-only stub for nitorch.plot.volumes.show_slices (unires/_update.py:10)."""


def show_slices(*args, **kwargs):  # pragma: no cover
    return None
