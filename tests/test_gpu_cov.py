"""f3: the lagged covariance blocks of the reference's VAMP-2 scan (CAVIAR-1.0.py:396-421) on the tensor cores
(csrc/cv_cov.cu: tcgen05 kind::tf32, three TF32 products per block) against a float64 restatement of those lines, with
the float32 cuBLAS product of the same inputs as the yardstick for "float32-class accuracy"."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ds():
    import __graft_entry__ as g
    g.build()
    from caviar_b200 import downstream
    return downstream


def _rel_err(c, ref, scale=None):
    """max |c - ref| relative to the largest variance involved (the scale of a covariance entry)."""
    return float(np.abs(c - ref).max() / (np.abs(np.diag(ref)).max() if scale is None else scale))


def _features(rng, n, K):
    """Correlated, un-centred distance-like columns (means 5..40 A, slow drifts + noise)."""
    t = np.linspace(0, 1, n)[:, None]
    base = rng.uniform(5.0, 40.0, size=(1, K))
    drift = np.sin(2 * np.pi * (t * rng.uniform(0.5, 3.0, size=(1, K)) + rng.random((1, K)))) * rng.uniform(0.2, 2.0, size=(1, K))
    return (base + drift + rng.normal(0, 0.3, size=(n, K))).astype(np.float32)


@pytest.mark.parametrize("n,K", [(1000, 300), (257, 128), (4099, 515), (31, 40)])
def test_cov_blocks_have_float32_accuracy(ds, n, K):
    import torch
    rng = np.random.default_rng(n + K)
    x = _features(rng, n + 7, K)
    mu = x.mean(0, dtype=np.float64).astype(np.float32)
    xc = x - mu                                                     # float32, as the reference centres (:381-382)
    x0, xt = xc[:-7], xc[7:]
    d0, dt = torch.from_numpy(x0).cuda(), torch.from_numpy(xt).cuda()
    got = [c.cpu().numpy() for c in ds.cov_blocks(d0, dt)]
    lib = [c.cpu().numpy() for c in ds.cov_blocks_cublas(d0, dt)]
    a64, b64 = x0.astype(np.float64), xt.astype(np.float64)
    want = [a64.T @ a64 / n, b64.T @ b64 / n, a64.T @ b64 / n]       # _cov_blocks (:396-402) in float64
    scale = max(np.diag(want[0]).max(), np.diag(want[1]).max())
    for c, l, w in zip(got, lib, want):
        assert c.shape == (K, K) and np.isfinite(c).all()
        e_ours, e_lib = _rel_err(c, w, scale), _rel_err(l, w, scale)
        assert e_ours <= 2e-6, (e_ours, e_lib)                      # a single TF32 pass sits at 1e-4 .. 5e-4
        assert e_ours <= 4 * e_lib + 2e-7, (e_ours, e_lib)          # the class of the float32 library product


def test_vamp_blocks_follow_the_reference_stacking(ds):
    """_stack_even_odd + _cov_blocks (:396-421): lagged pairs, even / odd selection, two ensembles stacked, one 1/n."""
    import torch
    rng = np.random.default_rng(3)
    K, tau = 200, 5
    xa, xb, xc = _features(rng, 700, K), _features(rng, 433, K), _features(rng, 4, K)      # the last one is too short
    mu = ((xa.mean(0) + xb.mean(0)) / 2.0).astype(np.float32)                               # mu_pool_sel (:379)
    ops = [ds.split_operand(torch.from_numpy(m).cuda(), torch.from_numpy(mu).cuda()) for m in (xa, xb, xc)]
    for parity in (None, 0, 1):
        x0s, xts = [], []
        for m in (xa, xb, xc):
            loc = (m - mu).astype(np.float64)
            if loc.shape[0] <= tau:
                continue
            x0, xt = loc[:-tau], loc[tau:]
            sel = np.ones(len(x0), bool) if parity is None else (np.arange(len(x0)) % 2 == parity)
            x0s.append(x0[sel]); xts.append(xt[sel])
        X0, Xt = np.vstack(x0s), np.vstack(xts)
        n = max(1, X0.shape[0])
        want = [X0.T @ X0 / n, Xt.T @ Xt / n, X0.T @ Xt / n]
        got = ds.vamp_cov_blocks(ops, tau, parity)
        scale = max(np.diag(want[0]).max(), np.diag(want[1]).max())
        for c, w in zip(got, want):
            assert _rel_err(c.cpu().numpy(), w, scale) <= 2e-6
    assert ds.vamp_cov_blocks(ops[2:], tau) is None


def test_split_operand_is_exact(ds):
    import torch
    rng = np.random.default_rng(9)
    x = torch.from_numpy((rng.normal(size=(50, 70)) * 10 ** rng.uniform(-3, 3, size=(50, 70))).astype(np.float32)).cuda()
    mu = x.mean(0)
    s = ds.split_operand(x, mu)
    assert s.hi.shape == (50, 128) and s.K == 70
    ref = x - mu
    assert torch.equal((s.hi[:, :70].double() + s.lo[:, :70].double()).float(), ref)        # hi + lo == x - mu, exactly
    assert (s.hi.view(torch.int32) & 0x1FFF).abs().max() == 0                               # hi is a TF32 number
    assert (s.hi[:, 70:] == 0).all() and (s.lo[:, 70:] == 0).all()
    assert (s.lo.abs() <= s.hi.abs() * 2.0 ** -11 + 1e-38).all()


def test_vamp_blocks_against_the_reference_own_functions(ds, golden_dir):
    """tests/golden/cov_blocks.npz was produced by the reference's own _stack_even_odd + _cov_blocks (cut out of
    CAVIAR-1.0.py:396-421 with ast, unmodified, numpy float32): the tensor-core path must agree to float32 accuracy."""
    import os
    import torch
    g = np.load(os.path.join(golden_dir, "cov_blocks.npz"))
    mu = torch.from_numpy(g["mu"]).cuda()
    ops = [ds.split_operand(torch.from_numpy(g[k]).cuda(), mu) for k in ("dist_a", "dist_b")]
    tau = int(g["tau"])
    for parity in (0, 1):
        got = ds.vamp_cov_blocks(ops, tau, parity)
        scale = max(np.diag(g[f"p{parity}_C00"]).max(), np.diag(g[f"p{parity}_Ctt"]).max())
        for c, name in zip(got, ("C00", "Ctt", "C0t")):
            want = g[f"p{parity}_{name}"]
            assert c.shape == want.shape
            assert np.abs(c.cpu().numpy() - want).max() <= 3e-6 * scale, (parity, name)      # two float32 products of the same data
