"""nitorch.core.constants (unires/_core.py:16)."""
inf = float('inf')
