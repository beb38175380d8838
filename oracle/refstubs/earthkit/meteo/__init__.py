"""Stub package (see oracle/refstubs/README.md)."""
