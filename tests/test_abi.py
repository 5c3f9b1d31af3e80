"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares.
No compute calls are made here (there is no GPU on the build machine)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_functions(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return sorted(set(re.findall(r"\b(mm2b_[a-z0-9_]+|mm_chain_dp)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    bld = pkg("build")
    lib = bld.build_all()
    assert os.path.exists(lib)
    L = ctypes.CDLL(lib)
    names = sorted(set(_declared_functions(os.path.join(ROOT, "include", "mm2chain_b200.h")) + _declared_functions(os.path.join(ROOT, "include", "mm2seed_b200.h"))))
    assert "mm_chain_dp" in names and "mm2b_chain_batch" in names and "mm2b_map_batch" in names and len(names) >= 31
    for n in names:
        assert hasattr(L, n), "C ABI symbol missing from the library: " + n
    binding = pkg("binding")
    assert sorted(binding.EXPORTS) == names, "binding.EXPORTS out of sync with the header"
    assert binding.load().mm2b_abi_version() == 4


def test_params_struct_layout_matches_oracle(pkg, oracle):
    binding = pkg("binding")
    assert ctypes.sizeof(binding.Params) == ctypes.sizeof(oracle.Params) == 40
    assert [f[0] for f in binding.Params._fields_] == [f[0] for f in oracle.Params._fields_]


def test_no_cpu_fallback_without_gpu(pkg):
    """On a machine without a usable CUDA device the product path must fail loudly, not compute on the CPU."""
    binding = pkg("binding")
    L = binding.load()
    if L.mm2b_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    import numpy as np
    off = np.array([0, 3], np.int64)
    a = np.zeros(3, binding.ANCHOR)
    with pytest.raises(binding.Mm2bError):
        binding.chain_batch(binding.Params(), off, a)
    with pytest.raises(binding.Mm2bError):
        binding.init()


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under minimap2-fpga_b200/ or include/ may reference it."""
    bad = []
    for base in ("minimap2-fpga_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    if re.search(r"import oracle|from oracle|oracle_py|oracle/|mm2o_|liboracle|libmm2ref|chain_oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def _unpack_model(lo, xr, yr):
    """numpy model of unpack_kernel (csrc/chain_kernels.cu): every anchor takes the high words of the last run starting at or before it."""
    import numpy as np
    n = len(lo)
    i = np.arange(n, dtype=np.uint32)
    xh = xr[np.searchsorted(xr[:, 0], i, side="right") - 1, 1].astype(np.uint64)
    yh = yr[np.searchsorted(yr[:, 0], i, side="right") - 1, 1].astype(np.uint64)
    return (xh << np.uint64(32)) | lo[:, 0], (yh << np.uint64(32)) | lo[:, 1]


def test_host_packer_round_trips(pkg):
    """mm2b_pack_anchors is pure host code: 8-byte words + runs of the high words restore every anchor bit for bit."""
    import numpy as np
    binding = pkg("binding")
    wl = pkg("workload")
    rng = np.random.default_rng(3)
    _, a = wl.synth_anchor_batch(40, seed=11)
    cases = [a, a[:1], a[:2]]
    wild = np.zeros(5000, binding.ANCHOR)                      # high words changing at random places, flags and segment ids in y
    wild["x"] = (rng.integers(0, 3, 5000).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 2**32, 5000).astype(np.uint64)
    wild["x"] |= rng.integers(0, 2, 5000).astype(np.uint64) << np.uint64(63)
    wild["y"] = (rng.integers(0, 4, 5000).astype(np.uint64) << np.uint64(48)) | (rng.integers(10, 30, 5000).astype(np.uint64) << np.uint64(32)) \
        | rng.integers(0, 2**32, 5000).astype(np.uint64)
    cases.append(wild)
    for arr in cases:
        lo, xr, yr = binding.pack_anchors(arr)
        assert xr[0, 0] == 0 and yr[0, 0] == 0 and np.all(np.diff(xr[:, 0].astype(np.int64)) > 0) and np.all(np.diff(yr[:, 0].astype(np.int64)) > 0)
        x, y = _unpack_model(lo, xr, yr)
        assert np.array_equal(x, arr["x"]) and np.array_equal(y, arr["y"])
    lo, xr, yr = binding.pack_anchors(a)
    assert len(yr) == 1 and len(xr) < len(a) / 50              # the format's premise on real-shaped input
    with pytest.raises(binding.Mm2bError):                    # high words too varied for the run budget: reported, not truncated
        binding.pack_anchors(wild, cap_runs=16)
