from . import type_check  # noqa: F401
