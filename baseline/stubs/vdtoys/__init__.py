"""Minimal stand-in for the `vdtoys` package (absent from this image, no network) so that the
UNMODIFIED reference under baseline/_ref can be imported for golden-vector generation and for the
reference-extension timing.  Only the two helpers the reference uses are provided."""
