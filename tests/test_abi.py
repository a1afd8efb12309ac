"""The C-ABI library loads here (no GPU), exports every symbol include/bang_b200.h declares, and fails loudly
without a device — there is no CPU fallback."""
import ctypes
import os
import re

import pytest

from bang_b200 import api, build

from conftest import ROOT, has_gpu


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "bang_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bang_b200_\w+|bang_\w+_c)\s*\(", hdr)))


def test_header_symbols_all_exported():
    build.build_cuda()
    lib = ctypes.CDLL(build.LIB_CUDA)
    declared = _declared_symbols()
    assert len(declared) >= 24
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/bang_b200.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == declared


def test_cpp_shim_symbols_exported():
    # BANGSearch<float|uint8_t|int8_t> with the reference's method set (bang.h:36-87)
    import subprocess
    out = subprocess.run(["nm", "-DC", build.LIB_CUDA], capture_output=True, text=True).stdout
    for t in ("float", "unsigned char", "signed char"):
        for m in ("bang_load(char*)", "bang_alloc(int)", "bang_init(int)", "bang_set_searchparams(int, int, _DistFunc)",
                  "bang_free()", "bang_unload()"):
            assert f"BANGSearch<{t}>::{m}" in out, (t, m)
        assert f"BANGSearch<{t}>::bang_query(" in out


def test_kernels_are_sm_100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_CUDA], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(has_gpu(), reason="this checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(api.BangError) as e:
        api.BANGSearch("uint8", "base")
    assert e.value.code == -4  # BANG_E_CUDA


def test_reference_driver_links_against_this_library():
    """Drop-in proof: the reference's own test_driver.cpp compiled against include/bang.h + libbang_b200.so
    (built by oracle/build_ref.sh when /root/reference is mounted)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "bang_search_dropin")
    if not os.path.exists(exe):
        pytest.skip("reference not mounted at build time")
    import subprocess
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libbang_b200.so" in out and "libbang.so" not in out


def test_filter_slots_match_the_reference_hashes():
    """The load-time form of the visited-filter slots (rows with the slot block) and the per-hop form in the kernel share one
    host/device function pair; its host side must give the reference's hashFn1_d / hashFn2_d (bang_search.cu:1168-1189, as
    restated in oracle/bang_oracle.c) for any id, and the documented word format.  Host arithmetic only: no device needed."""
    import numpy as np
    import oracle as O
    from bang_b200 import api
    rng = np.random.default_rng(5)
    ids = [0, 1, 254, 255, 256, 65535, 65536, 399886, 399887, 2**24 - 1, 2**24, 2**31 - 1, 2**31, 0xFFFFFFFE, 0xFFFFFFFF]
    ids += [int(x) for x in rng.integers(0, 2**32, size=2000, dtype=np.uint64)]
    for i in ids:
        pos, words = api.filter_slots(i)
        assert pos == (O.hash1(i), O.hash2(i)), i
        for p, w in zip(pos, words):
            assert p < 399887
            assert w == ((p // 255 * 8) | ((p % 255) << 24))
            assert (w & 0xFFFF) < 1569 * 8 and (w >> 24) < 255 and (w & 0x00FF0007) == 0
