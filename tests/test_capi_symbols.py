"""CPU: the C-ABI shared library loads and exports every symbol that
include/sdfcontact_b200.h declares; no compute entry succeeds without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sdfcontact_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SDFC_API\s+[\w\s\*]+?\b(sdfc_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("sdfc_sdf_create", "sdfc_points_create", "sdfc_pair_create", "sdfc_pair_query", "sdfc_pair_force",
                 "sdfc_pair_force_deriv", "sdfc_batch_create", "sdfc_batch_eval", "sdfc_batch_eval_device",
                 "sdfc_last_error"):
        assert must in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol():
    from assemble_them_all_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with python assemble-them-all_b200/build.py"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"not exported: {missing}"
    # the python binding table covers the header too
    assert set(_lib.SIGNATURES) >= set(n for n in declared_symbols() if not n.startswith("sdfc_step")), \
        set(declared_symbols()) - set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    """Without a CUDA device, creating a context must fail loudly."""
    from assemble_them_all_b200 import _lib, contact
    lib = _lib.load()
    assert lib.sdfc_version() >= 100
    if lib.sdfc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(contact.SdfContactError) as ei:
        contact.Context(0)
    assert "no CUDA device" in str(ei.value)


def test_host_scatter_helpers_match_reference_block_layout():
    """sdfc_scatter_* are pure host index arithmetic (ForceGeneralSDFContact.cpp:60-61,88-99)."""
    import numpy as np
    from assemble_them_all_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    fm_sub, Km_sub, Dm_sub = rng.normal(size=12), rng.normal(size=(12, 12)), rng.normal(size=(12, 12))
    nm, ia, ib = 18, 12, 0
    fm = np.zeros(nm)
    Km = np.zeros((nm, nm), order="F")
    Dm = np.zeros((nm, nm), order="F")
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
    kf, df = np.asfortranarray(Km_sub), np.asfortranarray(Dm_sub)
    lib.sdfc_scatter_force(ia, ib, dp(fm_sub), dp(fm))
    lib.sdfc_scatter_deriv(nm, ia, ib, dp(kf), dp(df), dp(Km), dp(Dm))
    ref_f = np.zeros(nm); ref_f[ia:ia + 6] += fm_sub[:6]; ref_f[ib:ib + 6] += fm_sub[6:]
    ref_K = np.zeros((nm, nm))
    for (r0, rs) in ((ia, 0), (ib, 6)):
        for (c0, cs) in ((ia, 0), (ib, 6)):
            ref_K[r0:r0 + 6, c0:c0 + 6] += Km_sub[rs:rs + 6, cs:cs + 6]
    assert np.array_equal(fm, ref_f) and np.array_equal(Km, ref_K)
