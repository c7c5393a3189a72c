"""Hermitian PSD blocks (K.s entries after K.rsdpN, stored [vec Re; vec Im]): plugins against the reference's
complex branches (invcholfac.c:131-141, psdframeit.c:86-97, ...) and the numpy restatement of psdscale.m."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, gpu, ref, relerr
from sedumi_b200.host import cones

sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def _K(sreal, sherm):
    K = {"l": 2.0, "q": np.array([3.0]), "s": np.array(list(sreal) + list(sherm), dtype=float), "rsdpN": len(sreal)}
    return cones.finish_K(K)


def _factor(K, rng, junk_lower=True):
    """d.u: per block an upper-triangular factor with positive real diagonal; the strictly lower part holds
    junk that triu() must ignore.  Hermitian blocks: [vec Re; vec Im]."""
    out = []
    nr = K["rsdpN"]
    for i, n in enumerate(K["s"].astype(int)):
        U = np.triu(rng.standard_normal((n, n))) + (0 if i < nr else 1j * np.triu(rng.standard_normal((n, n)), 1))
        U[np.diag_indices(n)] = np.abs(U[np.diag_indices(n)].real) + 0.5
        J = np.tril(rng.standard_normal((n, n)), -1) * (1.0 if junk_lower else 0.0)
        out.append((U.real + J).ravel(order="F"))
        if i >= nr:
            out.append((U.imag + J).ravel(order="F"))
    return np.concatenate(out)


def _herm_vec(K, rng):
    out = []
    nr = K["rsdpN"]
    for i, n in enumerate(K["s"].astype(int)):
        X = rng.standard_normal((n, n)) + (0 if i < nr else 1j * rng.standard_normal((n, n)))
        X = X + X.conj().T
        out.append(X.real.ravel(order="F"))
        if i >= nr:
            out.append(X.imag.ravel(order="F"))
    return np.concatenate(out)


def _perm(K, rng):
    return np.concatenate([rng.permutation(int(n)) + 1.0 for n in K["s"]]).reshape(-1, 1)


CASES = [((), (3,)), ((4,), (5,)), ((), (1, 2)), ((7, 2), (9, 33)), ((), (70,))]


@pytest.mark.parametrize("sreal,sherm", CASES)
@pytest.mark.parametrize("withperm", [False, True])
def test_invcholfac_hermitian(sreal, sherm, withperm):
    K = _K(sreal, sherm)
    Km = cones.K_for_mex(K)
    rng = np.random.default_rng(len(sherm) + 10 * sum(sherm))
    u = _factor(K, rng)
    args = (u, Km) + ((_perm(K, rng),) if withperm else ())
    yr = ref.invcholfac(*args)
    yg = gpu.invcholfac(*args)
    assert yg.shape == yr.shape
    assert relerr(yg, yr) <= 1e-10


@pytest.mark.parametrize("sreal,sherm", CASES)
@pytest.mark.parametrize("transp", [0.0, 1.0])
@pytest.mark.parametrize("withperm", [False, True])
def test_psdscale_hermitian(sreal, sherm, transp, withperm):
    import restate
    K = _K(sreal, sherm)
    Km = cones.K_for_mex(K)
    rng = np.random.default_rng(3 + len(sreal) + 10 * sum(sherm))
    u = _factor(K, rng, junk_lower=False)
    u = u + 0.0          # psdscale uses tril(U) for transp=0: give the lower triangle real content as well
    x = _herm_vec(K, rng)
    ud = {"u": _factor(K, rng), "perm": _perm(K, rng) if withperm else np.zeros((0, 0))}
    xfull = np.r_[rng.standard_normal(4), x]           # psdscale reads the PSD part from the tail
    yr = restate.psdscale(ud, xfull, K, bool(transp))
    yg = gpu.psdscale(ud, xfull, Km, transp)
    assert relerr(yg.ravel(), yr) <= 1e-10


def _frames(K, rng):
    """vfrm.s from the reference's own qrK on a random (complex for the Hermitian blocks) matrix per block."""
    x = []
    nr = K["rsdpN"]
    for i, n in enumerate(K["s"].astype(int)):
        x.append(rng.standard_normal(n * n))
        if i >= nr:
            x.append(rng.standard_normal(n * n))
    return ref.qrK(np.concatenate(x), cones.K_for_mex(K))


FCASES = [((), (2,)), ((), (3,)), ((4,), (5,)), ((), (1, 6)), ((7, 2), (9, 33)), ((), (70,)), ((100,), (60,))]


@pytest.mark.parametrize("sreal,sherm", FCASES)
def test_psdframeit_hermitian(sreal, sherm):
    K = _K(sreal, sherm)
    Km = cones.K_for_mex(K)
    rng = np.random.default_rng(7 + len(sreal) + 10 * sum(sherm))
    frms = _frames(K, rng)
    lab = rng.uniform(0.1, 3.0, int(K["s"].sum()))
    xr = ref.psdframeit(lab, frms, Km)
    xg = gpu.psdframeit(lab, frms, Km)
    assert xg.shape == xr.shape
    assert relerr(xg, xr) <= 1e-10


@pytest.mark.parametrize("sreal,sherm", FCASES)
def test_psdinvjmul_hermitian(sreal, sherm):
    K = _K(sreal, sherm)
    Km = cones.K_for_mex(K)
    rng = np.random.default_rng(11 + len(sreal) + 10 * sum(sherm))
    frms = _frames(K, rng)
    xlab = rng.uniform(0.5, 2.0, int(K["s"].sum()))
    y = _herm_vec(K, rng)
    zr = ref.psdinvjmul(xlab, frms, y, Km)
    zg = gpu.psdinvjmul(xlab, frms, y, Km)
    assert relerr(zg, zr) <= 1e-10


@pytest.mark.parametrize("l,sreal,sherm,m,dens", [(3, (4,), (3,), 6, 0.45), (2, (), (5,), 9, 0.5), (4, (6, 3), (4, 7), 25, 0.3),
                                                  (2, (), (20,), 40, 0.1), (2, (12,), (16, 2), 60, 0.15)])
def test_getada3_hermitian(l, sreal, sherm, m, dens):
    """[ADA,absd] = getada3(...) with Hermitian blocks against the reference's spcpxdxd path (spscale.c:332-435)."""
    import refpath
    from sedumi_b200.host import problems, setup
    At, b, c, K = problems.synth_hermitian_mixed(l=l, sreal=sreal, sherm=sherm, m=m, density=dens, seed=17 + m)
    S = setup.build_setup(At, b, c, K)
    d = problems.scaling_hermitian(K, 3 + m)
    R = refpath.RefHotPath(S)
    Km = R.Km
    udsqr = ref.invcholfac(d["u"], Km, d["perm"])
    assert relerr(gpu.invcholfac(d["u"], Km, d["perm"]), udsqr) <= 1e-10
    A1 = ref.getada1(R.ADA0, S.At, S.Ablkjc[:, 2], S.Aord["lqperm"], {"l": d["l"], "det": d["det"]}, S.K["qblkstart"].reshape(1, -1))
    A2 = ref.getada2(A1, {"q": R.DAtq(d)}, S.Aord, Km)
    Ar, absr = ref.getada3(A2, S.At, S.Ablkjc[:, 2], S.Aord, udsqr, Km, nlhs=2)
    Ag, absg = gpu.getada3(A2, S.At, S.Ablkjc[:, 2], S.Aord, udsqr, Km, nlhs=2)
    assert np.array_equal(Ag.indices, Ar.indices) and np.array_equal(Ag.indptr, Ar.indptr)
    assert relerr(Ag.data, Ar.data) <= 1e-10
    assert relerr(absg, absr) <= 1e-10
    assert np.array_equal(Ag.toarray(), Ag.toarray().T)


def _bad_factor_h(K, rng):
    """Upper-triangular (complex for the Hermitian blocks, real positive diagonal) factors whose leading columns
    are tiny: forces pivoting in (prpi)rotorder.  Stored Hermitian-mirrored like d.u."""
    out = []
    nr = K["rsdpN"]
    for i, n in enumerate(K["s"].astype(int)):
        scale = 10.0 ** rng.uniform(-3, 0, n)
        scale[: max(n // 3, 1)] *= 1e-3
        U = np.triu(rng.standard_normal((n, n))) + (0 if i < nr else 1j * np.triu(rng.standard_normal((n, n)), 1))
        U[np.diag_indices(n)] = np.abs(U[np.diag_indices(n)].real) + 0.5
        U = U * scale[None, :]
        full = U + np.triu(U, 1).conj().T
        out.append(full.real.ravel(order="F"))
        if i >= nr:
            out.append(full.imag.ravel(order="F"))
    return np.concatenate(out)


@pytest.mark.parametrize("sreal,sherm,maxu", [((), (2,), 1.1), ((), (6,), 1.1), ((5,), (7, 3), 1.1), ((20,), (33,), 1.1), ((), (50,), 3.0)])
def test_urotorder_givensrot_hermitian_bit_exact(sreal, sherm, maxu):
    K = _K(sreal, sherm)
    Km = cones.K_for_mex(K)
    rng = np.random.default_rng(sum(sherm) + 3)
    u = _bad_factor_h(K, rng)
    permin = _perm(K, rng)
    for extra in ((), (permin,)):
        ur, pr, gjr, gr = ref.urotorder(u, Km, maxu, *extra, nlhs=4)
        ug, pg, gjg, gg = gpu.urotorder(u, Km, maxu, *extra, nlhs=4)
        assert np.array_equal(pg, pr) and np.array_equal(gjg, gjr)
        assert gg.shape == gr.shape and np.array_equal(gg, gr)
        assert np.array_equal(ug, ur)
    assert gr.size > 0, "test input does not rotate"
    x = _herm_vec(K, rng)
    yr = ref.givensrot(gjr, gr, x, Km)
    yg = gpu.givensrot(gjr, gr, x, Km)
    assert np.array_equal(yg, yr)
