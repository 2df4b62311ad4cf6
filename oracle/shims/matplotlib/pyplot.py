"""Import-only stand-in for matplotlib.pyplot."""
