from . import modules, initialization  # noqa: F401
