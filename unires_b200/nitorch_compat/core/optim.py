"""nitorch.core.optim names used by UniRes (unires/_update.py:9, unires/run.py:7)."""
from ...optim import cg, get_gain  # noqa: F401


def plot_convergence(*args, **kwargs):
    """Live matplotlib plot of the objective (unires/run.py:90-99, sett.plot_conv): no-op."""
    return None
