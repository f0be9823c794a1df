"""Dead import in the reference (gp_kernel.py:4); nothing on the path uses it."""
