"""Inert stand-in for `matplotlib` (absent in this image); the reference imports pyplot at module top."""
