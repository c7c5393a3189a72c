"""Subtree-sharded factor and solves on ONE GPU: two HotPath objects play rank 0 and rank 1 of a
world of 2, the all-reduces are emulated by adding their buffers; the result must equal the unsharded path."""
import numpy as np
import pytest

from helpers import relerr
from sedumi_b200.host import cones, problems, setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nblk,n,m,nlink", [(6, 12, 90, 8), (9, 10, 140, 20)])
def test_subtree_sharded_factor_and_solve_match_unsharded(nblk, n, m, nlink):
    import torch
    from sedumi_b200 import device
    raw = problems.synth_blockdiag_sdp(nblk=nblk, n=n, m=m, nlink=nlink, density=0.12, seed=21 + nblk)
    At, b, c, K = cones.pretransfo(*raw)[:4]
    S = setup.build_setup(At, b, c, K, perm=np.arange(At.shape[1]))
    assert len(S.L["xsuper"].ravel()) - 1 > 2
    d = problems.scaling(K, "S1", seed=4)
    rng = np.random.default_rng(1)
    rhs = rng.standard_normal((S.m, 2))
    hps = [device.HotPath(S) for _ in range(3)]          # [unsharded, rank 0, rank 1]
    st = hps[0].stream()
    with torch.cuda.stream(st):
        for hp in hps:
            hp.set_scaling(d)
            hp.set_rhs(rhs)
            hp.invcholfac()
            hp.getada()
        ref = hps[0]
        ref.blkchol()
        ref.solve()
        r0, r1 = hps[1], hps[2]
        sh0, sh1 = r0.shard_factor_setup(2, 0), r1.shard_factor_setup(2, 1)
        assert sh0["t0"] == sh1["t0"] and 0 < sh0["t0"] < len(S.L["xsuper"].ravel()) - 1
        for hp in (r0, r1):
            hp.blkchol_shard_local()
        tot = r0.top_panels() + r1.top_panels()           # the all-reduce
        r0.top_panels().copy_(tot); r1.top_panels().copy_(tot)
        for hp in (r0, r1):
            hp.blkchol_shard_top()
        st.synchronize()
        # pivots: every column is answered for by exactly one rank (its trees) or by the replicated top
        col0 = sh0["col0"]
        dref = ref.dvec.cpu().numpy()[:S.m]
        d0, d1 = r0.dvec.cpu().numpy()[:S.m], r1.dvec.cpu().numpy()[:S.m]
        assert relerr(d0[col0:], dref[col0:]) <= 1e-10 and relerr(d1[col0:], dref[col0:]) <= 1e-10
        own = (d0[:col0] != 0).astype(int) + (d1[:col0] != 0).astype(int)
        assert np.all(own == 1)
        assert relerr(d0[:col0] + d1[:col0], dref[:col0]) <= 1e-10
        # top panels of the factor
        assert relerr(r0.top_panels().cpu().numpy(), ref.Lrect[sh0["top_off"]:sh0["top_off"] + sh0["top_len"]].cpu().numpy()) <= 1e-10
        # solves
        for hp in (r0, r1):
            hp.solve_shard_local()
        seg = r0.w[:, col0:] + r1.w[:, col0:]
        r0.w[:, col0:].copy_(seg); r1.w[:, col0:].copy_(seg)
        for hp in (r0, r1):
            hp.solve_shard_top()
        z = r0.w + r1.w
        r0.w.copy_(z); r1.w.copy_(z)
        for hp in (r0, r1):
            hp.solve_shard_finish()
        st.synchronize()
        yref = ref.y.cpu().numpy()
        assert relerr(r0.y.cpu().numpy(), yref) <= 1e-9
        assert relerr(r1.y.cpu().numpy(), yref) <= 1e-9
