"""Fake ``diffusers`` namespace for importing the reference pipelines offline (see ../README.md)."""
__version__ = "0.0.refshim"
