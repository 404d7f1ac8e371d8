"""Shim of torch-geometric==1.7.2 (test infrastructure; see ../README.md)."""
__version__ = "1.7.2+shim"
