"""Import-only stand-in (the reference imports matplotlib at module scope; the hot path never calls it)."""
colormaps = {}
