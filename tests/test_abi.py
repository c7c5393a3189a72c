"""CPU tests of the drop-in boundary: the C-ABI library and the MEX plugins load, export every
declared symbol, and fail loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import CHOL_PARS, ROOT, dense_L, full_pattern, gpu, random_spd
from sedumi_b200 import device
from sedumi_b200.mx import MexError

PLUGINS = ["blkchol", "fwblkslv", "bwblkslv", "getada1", "getada2", "getada3", "invcholfac", "psdscale",
           "ddot", "qblkmul", "quadadd", "psdframeit", "psdinvjmul", "urotorder", "givensrot",
           "dpr1fact", "fwdpr1", "bwdpr1", "adendotd", "adenscale",
           "vecsym", "sqrtinv", "qrK", "psdjmul", "triumtriu", "psdfactor", "psdinvscale", "psdeig", "minpsdeig"]


def test_library_exports_every_declared_symbol():
    L = device.lib()
    hdr = open(os.path.join(ROOT, "include", "sedumi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sb200_\w+)\s*\(", hdr)))
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(set(device.EXPORTS) - set(declared)) == []


@pytest.mark.parametrize("name", PLUGINS)
def test_mex_plugin_exports_mexFunction(name):
    so = os.path.join(ROOT, "sedumi_b200", "mex", name + ".so")
    assert os.path.exists(so), "run __graft_entry__.build()"
    lib = C.CDLL(so)
    assert hasattr(lib, "mexFunction")


def test_every_plugin_source_has_a_binary():
    srcs = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "sedumi_b200", "mex", "*.cpp")))
    assert srcs == sorted(PLUGINS)


@pytest.mark.skipif(device.lib().sb200_device_count() > 0, reason="a GPU is present")
def test_no_gpu_means_loud_failure_not_fallback():
    m = 5
    X = random_spd(m, 1)
    with pytest.raises(MexError, match="no CUDA device"):
        gpu.blkchol(dense_L(m), full_pattern(X), CHOL_PARS, np.diag(X).copy(), nlhs=4)
    with pytest.raises(MexError, match="no CUDA device"):
        gpu.quadadd(np.ones(3), np.zeros(3), np.ones(3), nlhs=2)


def test_argument_validation_like_the_reference():
    """Bad inputs raise MATLAB-style errors before any device work (fwblkslv.c:213-248 messages)."""
    m = 4
    L = dense_L(m)
    with pytest.raises(MexError, match="Size mismatch b"):
        gpu.fwblkslv(L, np.ones((m + 1, 1)))
    with pytest.raises(MexError, match="Missing field L.perm"):
        gpu.fwblkslv({"L": L["L"], "xsuper": L["xsuper"]}, np.ones((m, 1)))
    with pytest.raises(MexError, match="P must be square"):
        gpu.blkchol(L, sp.csc_matrix(np.ones((m, m + 1))), CHOL_PARS, nlhs=1)


def test_content_hash_does_not_depend_on_the_thread_count():
    """The 128-bit content key of the plan / mirror caches (context.cu: chunk hashes combined in chunk order): same value with
    1 and with 8 hashing threads, across the single-/multi-threaded size boundary, and sensitive to one flipped bit."""
    import subprocess
    import sys
    code = (
        "import ctypes as C, numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "from sedumi_b200 import device\n"
        "L = device.lib()\n"
        "rng = np.random.default_rng(5)\n"
        "out = []\n"
        "for nbytes in (0, 17, 512 * 1024, 512 * 1024 + 8, 3 * 1024 * 1024 + 40):\n"
        "    buf = rng.integers(0, 256, nbytes, dtype=np.uint8)\n"
        "    h = (C.c_uint64 * 2)()\n"
        "    L.sb200_content_hash(buf.ctypes.data_as(C.c_void_p), C.c_int64(nbytes), h)\n"
        "    out.append((h[0], h[1]))\n"
        "    if nbytes:\n"
        "        buf[nbytes // 2] ^= 1\n"
        "        g = (C.c_uint64 * 2)()\n"
        "        L.sb200_content_hash(buf.ctypes.data_as(C.c_void_p), C.c_int64(nbytes), g)\n"
        "        assert (g[0], g[1]) != (h[0], h[1])\n"
        "print(out)\n" % ROOT)
    res = []
    for nthreads in ("1", "8"):
        env = dict(os.environ, SB200_HASH_THREADS=nthreads)
        res.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout)
    assert res[0] == res[1] and len(res[0]) > 20
