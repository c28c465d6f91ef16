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
    with pytest.raises(GkqError) as e:  # the priority variant (gkq_create_prio) refuses the same way
        Engine(0, stream_priority=-1)
    assert e.value.code == -3
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


def test_shard_geometry_matches_the_host_mirror():
    """gkq_shard_size / gkq_shard_index (pure host functions of the multi-GPU ABI) describe the same
    block-cyclic partition as the Python mirror: every integral has exactly one
    owner, ragged last blocks included."""
    import numpy as np
    from gkquad_b200 import _ffi
    from gkquad_b200.workloads import block_cyclic_shard
    L = _ffi.lib()
    for n, w, b in [(0, 2, 4096), (1, 3, 4), (100, 3, 8), (1 << 20, 8, 4096), (100000, 8, 4096), (100000, 2, 4096),
                    (12345, 5, 7), (4096 * 8 + 1, 8, 4096)]:
        seen = np.zeros(n, dtype=np.int32)
        for r in range(w):
            idx = block_cyclic_shard(n, r, w, b)
            assert L.gkq_shard_size(n, w, r, b) == len(idx)
            got = np.array([L.gkq_shard_index(n, w, r, b, j) for j in range(len(idx))] if n <= 20000 else
                           [L.gkq_shard_index(n, w, r, b, j) for j in (0, len(idx) // 2, len(idx) - 1) if len(idx)])
            want = idx if n <= 20000 else idx[[0, len(idx) // 2, len(idx) - 1]] if len(idx) else idx
            assert np.array_equal(got, want)
            seen[idx] += 1
            assert L.gkq_shard_size(n, w, 0, b) >= len(idx)  # rank 0 owns the largest shard (gkq_gather's common range)
        assert np.all(seen == 1)
    assert L.gkq_shard_size(10, 0, 0, 4) < 0 and L.gkq_shard_index(10, 2, 0, 4, 99) < 0


def test_comm_entries_refuse_misuse_without_a_device():
    from gkquad_b200 import _ffi
    L = _ffi.lib()
    assert L.gkq_comm_init(None, 2, 0, None) < 0 and L.gkq_gather(None, 0, 1, 0, *([None] * 9)) < 0
    assert L.gkq_comm_destroy(None) < 0 and L.gkq_comm_get_unique_id(None) < 0
