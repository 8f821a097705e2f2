"""reference module of the same name: it only re-exports src.containers."""
from src.containers import *  # noqa: F401,F403
from src.containers import __all__  # noqa: F401
