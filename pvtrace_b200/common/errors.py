"""Exception types (names kept from pvtrace/common/errors.py:1-13)."""


class AppError(Exception):
    """Misuse of the public API (e.g. emitting from a node without a light)."""


class TraceError(AppError):
    """A ray could not be traced."""


class GeometryError(AppError):
    """A geometric query was made with an invalid argument (e.g. a point off the surface)."""
