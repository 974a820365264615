from .api import Runner, get_runner  # noqa: F401
