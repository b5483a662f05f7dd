"""Drop-in mirrors of the reference's `mesh` package for the hot path (same class names, constructor arguments,
forward signatures and return structures; /root/reference/mesh/*.py).  The arithmetic is in libspherehand_b200.so."""
