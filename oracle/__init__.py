"""CPU oracle for the PnP-AdaNet hot path (test infrastructure only; TF-kernel numerics are an unpinned restatement, the
rest is pinned to executed reference code -- see tf14_numpy.py and DESIGN.md section 2)."""
