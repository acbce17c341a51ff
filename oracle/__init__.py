"""CPU oracle for the PnP-AdaNet hot path (test infrastructure only; parity unpinned -- see tf14_numpy.py)."""
