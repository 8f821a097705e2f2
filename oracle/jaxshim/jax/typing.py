"""jax.typing stand-in (annotations only)."""
from typing import Any
ArrayLike = Any
