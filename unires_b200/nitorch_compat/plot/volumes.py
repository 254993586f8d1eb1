"""nitorch.plot.volumes.show_slices (unires/_update.py:10, unires/run.py:8): plotting is not
part of the path; a no-op keeps the verbose branches importable."""


def show_slices(*args, **kwargs):
    return None
