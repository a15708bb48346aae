from . import assertion, exceptions  # noqa: F401
