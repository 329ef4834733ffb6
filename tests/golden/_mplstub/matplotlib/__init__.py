"""No-op stand-in for matplotlib (absent from the image) so the reference's modules import."""
