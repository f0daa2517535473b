"""The C-ABI library: loads, exports every symbol include/*.h declares, host-side tables are right,
and it fails loudly without a GPU.  No compute calls (CPU only)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_libs as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    inc = os.path.join(ROOT, "include")
    for f in sorted(os.listdir(inc)):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            names += re.findall(r"SRSLTE_(?:B200_)?API\s+[\w\s\*]+?\b(srslte_\w+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    syms = _declared_symbols()
    assert len(syms) >= 18
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(pkg.EXPORTS) <= set(syms)


def test_host_tables_match_oracle(pkg):
    L, P = pkg.lib(), ol.port()
    for i, K in enumerate(ol.ALL_K):
        assert L.srslte_b200_cb_size(i) == K
        assert L.srslte_b200_cb_index(K) == i == L.srslte_b200_cb_index(K - 1)
        assert L.srslte_b200_nof_windows(K) == P.port_nof_subblocks(K)
        assert L.srslte_b200_working_len(K) == (3 * (K + 32) + 12 if K > 400 else 3 * K + 12)
    assert L.srslte_b200_cb_size(188) == -1 and L.srslte_b200_cb_index(6145) == -1


def test_rm_rx_table_matches_golden_and_oracle(pkg, golden):
    P = ol.port()
    for key, want in golden["rm_tables"].items():
        K, rv, sb = (int(x) for x in key.replace("K", "").replace("rv", "").replace("sb", "").split("_"))
        assert np.array_equal(pkg.rm_rx_table(K, rv, bool(sb)), want), key
    for K in ol.ALL_K[::5]:
        for rv in range(4):
            t = np.zeros(3 * K + 12, np.uint16)
            P.port_rm_rx_table(K, rv, 1, t)
            assert np.array_equal(pkg.rm_rx_table(K, rv, True), t), (K, rv)
    t = np.zeros(3 * 40 + 12, np.uint16)
    assert pkg.lib().srslte_b200_rm_rx_table(41, 0, 1, t.ctypes.data) == pkg.ERROR_INVALID_INPUTS
    assert pkg.lib().srslte_b200_rm_rx_table(40, 4, 1, t.ctypes.data) == pkg.ERROR_INVALID_INPUTS


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the context cannot be created; with one it can.  Either way NULL is rejected."""
    import torch
    L = pkg.lib()
    assert L.srslte_b200_ctx_create(None, 0) == pkg.ERROR_INVALID_INPUTS
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert L.srslte_b200_ctx_create(C.byref(h), 0) == pkg.ERROR
        assert not h.value
        with pytest.raises(pkg.B200Error):
            pkg.Context(0)


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pk = os.path.join(ROOT, "srslte-emane_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_libs" not in src and "tdec_port" not in src and "libsrslte_ref" not in src, f


def test_vectors_roundtrip_tables(pkg, vec):
    """the TX mirror used to make inputs is consistent with the library's receive tables."""
    K, rv = 1024, 2
    rng = np.random.default_rng(0)
    coded = rng.integers(0, 2, (1, 3 * K + 12), dtype=np.uint8)
    N = 3 * K + 12
    e = vec.rate_match(coded, N, rv)[0]
    tab = pkg.rm_rx_table(K, rv, sb_layout=False)
    back = np.zeros(N, np.uint8)
    back[tab] = e
    assert np.array_equal(back, coded[0])
