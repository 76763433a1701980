from . import filters, util  # noqa: F401
