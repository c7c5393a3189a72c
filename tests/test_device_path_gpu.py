"""The device-resident pipeline (sedumi_b200.device.HotPath, what bench.py's `value` times) against
the reference call sequence (oracle/refpath.py).  Gates: ADA/absd/factor 1e-10, search direction 1e-8."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, relerr
from sedumi_b200.host import cones, problems, setup

sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def _run(raw, perm=None, seed=3):
    import torch
    import refpath
    from sedumi_b200 import device
    At, b, c, K = cones.pretransfo(*raw)[:4]
    S = setup.build_setup(At, b, c, K, perm=perm)
    d = problems.scaling(K, "S1", seed=seed)
    rng = np.random.default_rng(seed)
    rhs = rng.standard_normal((S.m, 2))
    psd_x = rng.standard_normal(int((np.asarray(K["s"]) ** 2).sum()))
    hp = device.HotPath(S)
    st = hp.stream()
    with torch.cuda.stream(st):
        hp.set_scaling(d)
        hp.set_rhs(rhs)
        hp.psd_x[:psd_x.size].copy_(torch.from_numpy(psd_x))
        st.synchronize()
        hp.iteration(1, 2)
        hp.sync()
        out = dict(ADA=hp.ADA.cpu().numpy()[:S.ADA.nnz], absd=hp.absd.cpu().numpy()[:S.m],
                   d=hp.dvec.cpu().numpy()[:S.m], y=hp.y.cpu().numpy().T, psd=hp.psd_y.cpu().numpy()[:psd_x.size],
                   udsqr=hp.udsqr.cpu().numpy()[:psd_x.size])
    ref = refpath.RefHotPath(S).iteration(d, rhs, psd_x, 1, 2)
    assert relerr(out["udsqr"], ref["udsqr"].ravel()) <= 1e-10
    assert relerr(out["ADA"], ref["ADA"].data) <= 1e-10
    assert relerr(out["absd"], ref["absd"].ravel()) <= 1e-10
    assert relerr(out["d"], ref["L"]["d"]) <= 1e-10
    assert relerr(out["y"], ref["y"]) <= 1e-8
    assert relerr(out["psd"], ref["psd"]) <= 1e-10
    return S


def test_control07():
    _run(problems.load_fixture("control07"))


def test_arch0():
    _run(problems.load_fixture("arch0"))


def test_small_sdp():
    _run(problems.synth_small_mixed(seed=7, m=30, l=0, q=(), s=(9, 6), density=0.25))


def test_blockdiag_arrow_multisupernode():
    raw = problems.synth_blockdiag_sdp(nblk=6, n=30, m=200, nlink=20, density=0.04, seed=8)
    m = raw[0].shape[1]
    S = _run(raw, perm=np.arange(m))
    assert len(S.L["xsuper"]) - 1 > 1
