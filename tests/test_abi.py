"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/gkquad_b200.h declares, answers the host-only queries, and refuses to compute
without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (ROOT / "include" / "gkquad_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gkq_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_lists():
    from gkquad_b200 import _ffi
    assert _header_symbols() == sorted(_ffi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from gkquad_b200 import _ffi
    L = _ffi.lib()
    for s in _header_symbols():
        assert hasattr(L, s), f"libgkquad_b200.so does not export {s}"
    assert L.gkq_abi_version() == 2


def test_struct_layouts_match_header():
    from gkquad_b200 import _ffi
    assert C.sizeof(_ffi.gkq_tolerance) == 24
    assert C.sizeof(_ffi.gkq_config) == 8 + 24 + 8 + 8 + 8 + 8
    assert C.sizeof(_ffi.gkq_stats) == 11 * 8
    cfg = _ffi.gkq_config()
    assert _ffi.lib().gkq_config_default(1, C.byref(cfg)) == 0
    assert (cfg.algo, cfg.tol.kind, cfg.tol.abs_tol, cfg.tol.rel_tol, cfg.max_evals) == (0, 2, 1.49e-8, 1.49e-8, 2000)
    assert _ffi.lib().gkq_config_default(2, C.byref(cfg)) == 0 and cfg.max_evals == 100000
    assert _ffi.lib().gkq_config_default(3, C.byref(cfg)) < 0


def test_registry_matches_oracle_registry():
    """Same ids, names and parameter counts on both sides of the parity tests."""
    from gkquad_b200 import _ffi
    from oracle import oracle as O
    for dim in (1, 2, 0):
        ours = [(i, n, p) for i, n, p, _ in _ffi.functor_table(dim)]
        # (the oracle may carry extra probe functors behind the shared ids, e.g. `expcos_p_ulp`:
        # tools/libm_sensitivity.py)
        theirs = [row for row in O.functor_table(dim) if not row[1].endswith("_ulp")]
        assert ours == theirs and theirs == O.functor_table(dim)[:len(theirs)], dim
    L = _ffi.lib()
    assert L.gkq_functor_lookup(1, b"expcos_p") == 3
    assert L.gkq_functor_lookup(1, b"no_such_functor") < 0
    assert L.gkq_functor_count(7) < 0
    # every transcendental functor carries a measured flop count (profiles/functor_flops_r01.md)
    assert dict((n, f) for _, n, _, f in _ffi.functor_table(1))["expcos_p"] == 61.0


def test_no_device_means_no_compute():
    """Without a CUDA device gkq_create must fail loudly; there is no CPU path in the product."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gkquad_b200 import Engine, GkqError
    with pytest.raises(GkqError) as e:
        Engine(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    L = __import__("gkquad_b200")._ffi.lib()
    assert L.gkq_last_error(None)  # message of the failed create is retrievable
    # NULL handle is rejected, not dereferenced
    assert L.gkq_sync(None, None) < 0 and L.gkq_destroy(None) == 0


def test_product_does_not_reach_into_the_oracle():
    """The oracle is test infrastructure: nothing under gkquad-rs_b200/ may import, include or
    link it."""
    pkg = ROOT / "gkquad-rs_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + [pkg / "Makefile"]:
        if "build" in p.parts:
            continue
        text = p.read_text()
        assert "gkq_oracle" not in text and "from oracle" not in text and "import oracle" not in text, p
        assert "libgkq_oracle" not in text, p
