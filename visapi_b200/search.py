"""Drop-in for the reference's f2py module ``search`` (engine/search.f90:1-16, imported at
engine/solution_checker.py:10): ``find_first(needle, haystack)`` on the device."""
from . import runtime


def find_first(needle, haystack):
    """First 0-based index k with int32(haystack[k]) == needle, else -1."""
    return runtime.get_engine().find_first(needle, haystack)
