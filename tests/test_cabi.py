"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports exactly
the entry points include/rz_b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os

from radzero_b200 import _lib


def test_library_builds_and_loads():
    lib = _lib.load()
    assert os.path.exists(_lib.LIB_PATH)
    assert lib.rz_version() >= 1
    assert lib.rz_strerror(0) == b"ok"
    assert b"invalid" in lib.rz_strerror(-1)


def test_every_declared_symbol_is_exported_and_bound():
    declared = _lib.declared_symbols()
    assert len(declared) >= 8
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in rz_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "python binding table out of sync with the header"


def test_argument_validation_without_gpu():
    lib = _lib.load()
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.rz_prep_rows(None, 0, None, None, 4, 4, 4, None, None, None, 1, None) == -1
    assert lib.rz_upsample_maps(None, 0, 1, 37, 8, 8, 8, 8, 0, 0, 0.0, 0, 0.5, None, None) == -1
    assert lib.rz_mpnce_partials(None, 0, 1, 1, 1, None, 0, 1.0, None, 1e-8, 0, None, None, None, None, None) == -1
    assert lib.rz_sim_fwd_tokens(None, 0, None, None, 1, 1, 1370, None, 14, 1.0, None, None, None, 0, 0, 1,
                                 None, 0, 0, 1.0, None, 0, None, None, 0, None) == -1
    assert lib.rz_sim_fwd_tokens_workspace_bytes(0, 14) == 0
    assert lib.rz_sim_fwd_large_workspace_bytes(2, 100, 1408) > 2 * 100 * 1408 * 2


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under radzero_b200/ may import it (only tests/,
    __graft_entry__.smoke() and bench.py's CPU legs do)."""
    import ast
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "radzero_b200")
    bad = []
    for dirpath, _, files in os.walk(root):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module:
                    names = [node.module]
                if any(n == "oracle" or n.startswith("oracle.") for n in names):
                    bad.append(f)
    assert not bad, f"product modules importing the oracle: {bad}"
