"""the direct step of wrapPcg on the device (sb200_wrappcg_dev) against the restated wrapPcg.m:42-97
driving the reference's own fwblkslv / bwblkslv / vecsym (oracle/refpath.py::wrappcg_direct).  Gate: search direction
and residual 1e-8 relative."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, relerr
from sedumi_b200.host import cones, problems, setup

sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def _run(raw, perm=None, seed=5, use_rb=True):
    import torch
    import refpath
    from sedumi_b200 import device
    At, b, c, K = cones.pretransfo(*raw)[:4]
    S = setup.build_setup(At, b, c, K, perm=perm)
    assert len(K["q"]) == 0 and not len(S.dense.cols)
    d = problems.scaling(K, "S1", seed=seed)
    rng = np.random.default_rng(seed)
    N = S.At.shape[0]
    rv = rng.standard_normal(N)
    rb = rng.standard_normal(S.m) if use_rb else None
    hp = device.HotPath(S)
    with torch.cuda.stream(hp.stream()):
        hp.set_scaling(d)
        hp.set_rhs(rng.standard_normal((S.m, 1)))
        hp.invcholfac(); hp.getada(); hp.blkchol()
        hp.sync()
    got = hp.wrappcg(rv, rb)
    R = refpath.RefHotPath(S)
    udsqr, ADA, absd = R.assemble(d)
    L = R.factor(ADA, absd)
    ref = refpath.wrappcg_direct(R, L, d, rv, rb)
    for k in ("y", "dx"):
        assert np.linalg.norm(got[k] - ref[k]) <= 1e-8 * np.linalg.norm(ref[k]), (k, np.linalg.norm(got[k] - ref[k]))
    for k in ("ssqrNew", "ssqrdx", "alpha"):
        assert abs(got[k] - ref[k]) <= 1e-9 * abs(ref[k]), (k, got[k], ref[k])
    # after the direct step the residual is rounding noise on both sides: compare on the scale of the data
    scale = np.linalg.norm(rv) + (np.linalg.norm(rb) if rb is not None else 0.0)
    assert np.linalg.norm(got["r"] - ref["r"]) <= 1e-8 * scale and abs(got["normr"] - ref["normr"]) <= 1e-8 * scale
    assert got["normr"] <= 1e-8 * scale

def test_wrappcg_small_sdp():
    _run(problems.synth_small_mixed(seed=7, m=30, l=4, q=(), s=(9, 6), density=0.25))


def test_wrappcg_control07():
    _run(problems.load_fixture("control07"), use_rb=False)


def test_wrappcg_blockdiag_multisupernode():
    raw = problems.synth_blockdiag_sdp(nblk=6, n=30, m=200, nlink=20, density=0.04, seed=8)
    _run(raw, perm=np.arange(raw[0].shape[1]))
