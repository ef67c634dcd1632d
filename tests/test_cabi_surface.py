"""The C-ABI shared library loads without a GPU and exports exactly what
include/rectpack_b200.h declares; the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rectpack_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z_0-9]+)\s*\(", text)))


def test_build_entry_point_compiles_everything():
    import __graft_entry__
    __graft_entry__.build()
    assert os.path.exists(os.path.join(ROOT, "visapi_b200", "librectpack_b200.so"))
    assert os.path.exists(os.path.join(ROOT, "oracle", "_build", "librectpack_oracle.so"))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, "visapi_b200", "librectpack_b200.so"))
    for n in names:
        assert hasattr(lib, n), "declared in the header but not exported: %s" % n


def test_binding_table_matches_header():
    from visapi_b200 import _native
    bound = sorted(n for n, _, _ in _native.SYMBOLS)
    assert bound == declared_functions()
    _native.load_library()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import visapi_b200
    with pytest.raises(visapi_b200.EngineUnavailable):
        visapi_b200.Engine(0)
    from visapi_b200.solution_checker import SolutionChecker
    import numpy as np
    with pytest.raises(visapi_b200.EngineUnavailable):
        SolutionChecker.get_next_lfb_on_grid(np.zeros((2, 2)))
    from visapi_b200 import runtime
    with pytest.raises(ValueError):
        runtime.parse_device("cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "visapi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "rectpack_oracle" not in src, f


def test_search_modes_match_the_header():
    """bp_search_mode in include/rectpack_b200.h and the names Search.set_mode accepts."""
    import inspect
    import re
    from visapi_b200 import _native
    header = open(os.path.join(ROOT, "include", "rectpack_b200.h")).read()
    enum = re.search(r"typedef enum \{([^}]*)\} bp_search_mode;", header).group(1)
    values = {k.strip(): int(v) for k, v in (item.split("=") for item in enum.split(","))}
    assert values == {"BP_SEARCH_AUTO": 0, "BP_SEARCH_STEPPED": 1, "BP_SEARCH_RESIDENT": 2, "BP_SEARCH_HYBRID": 3}
    src = inspect.getsource(_native.Search.set_mode)
    for name, value in (("auto", 0), ("stepped", 1), ("resident", 2), ("hybrid", 3)):
        assert '"%s": %d' % (name, value) in src
