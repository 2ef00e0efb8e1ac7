from . import nps  # noqa: F401
