"""Stand-in for the reference's f2py module ``search`` (engine/search.f90:1-16).

TEST INFRASTRUCTURE.  The prebuilt engine/search.cpython-36m/37m .so files cannot load
in CPython 3.12 and there is no gfortran here, so the harness supplies the same
semantics: cast the haystack to INTEGER(4), return the first 0-based index whose
element equals ``needle``, or -1.
"""
import numpy as np


def find_first(needle, haystack):
    hay = np.asarray(haystack).astype(np.int32, copy=False).ravel()
    hits = np.flatnonzero(hay == np.int32(needle))
    return int(hits[0]) if hits.size else -1
