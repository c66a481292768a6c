"""Stub standing in for getdist (absent in this image); only nnest/ensemble.py imports it."""
