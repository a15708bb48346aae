"""Validation errors raised before any device work starts.

Same names and base class as the reference (decomp/utils/exceptions.py:1-10) so that
``except ValueError`` / ``except ShapeMismatchError`` in user code keeps working.
"""


class ShapeMismatchError(ValueError):
    """Two arguments disagree on an axis they must share."""


class DimInvalidError(ValueError):
    """An argument has the wrong number of dimensions."""


class DtypeMismatchError(ValueError):
    """Arguments mix dtypes, or use a dtype kind the solver does not accept."""
