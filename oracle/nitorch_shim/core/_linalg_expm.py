"""Import-only stub: _expm belongs to the rigid update (out of scope)."""


def _expm(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('rigid update is out of scope (SURVEY.md 8f #2)')
