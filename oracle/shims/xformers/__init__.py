"""Stand-in for xformers==0.0.27 (test infrastructure; see oracle/shims/README.md)."""
from . import ops  # noqa: F401
