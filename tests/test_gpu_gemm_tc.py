"""GPU parity for the tensor-core prefill GEMM (tcgen05 + TMEM + TMA): Y = X * W^T with f16 weights and
f32 activations split into hi + lo f16 terms, against the oracle's cmdbuf_mat_mul restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import token_hawk_b200 as t
    d = t.Device(0)
    yield d
    d.close()


@pytest.mark.parametrize("M,N,K", [(128, 4096, 4096), (128, 11008, 4096), (128, 4096, 11008), (100, 512, 512), (200, 1536, 512), (8, 32, 64)])
def test_gemm_f16_tc_matches_oracle(dev, oracle, M, N, K):
    import token_hawk_b200 as t
    Kl = t.kernels()
    r = np.random.default_rng(M + N + K)
    X = (r.standard_normal((M, K)) * 1.5).astype(np.float32)
    W = (r.standard_normal((N, K)) * 0.02).astype(np.float16)
    dX, dW, dY = dev.array(X), dev.array(W), dev.array(np.full((M, N), np.nan, np.float32))
    assert Kl.thk_gemm_f16_tc(dev.h, dX.ptr, dW.ptr, dY.ptr, M, N, K) == 0, Kl.thk_last_error()
    assert Kl.thk_gemm_check(dev.h) == 0, Kl.thk_last_error()
    Y = dY.numpy()
    assert np.isfinite(Y).all()
    # oracle on a sample of rows (the full product is large): cmdbuf_mat_mul, transposeB=1, f16 B
    rows = sorted(set([0, 1, M // 2, M - 1] + r.integers(0, M, 4).tolist()))
    ref = oracle.mat_mul(X[rows][None], W[None], True)[0]
    bound = np.abs(X[rows]).astype(np.float64) @ np.abs(W.astype(np.float64)).T
    err = np.abs(Y[rows] - ref) / bound
    assert err.max() < 2e-6, err.max()                      # hi/lo split keeps ~22 bits of X
    assert np.abs(Y[rows] - ref).max() / np.abs(ref).max() < 1e-4                 # budget on logits is 1e-3
    # and the whole matrix against float64 numpy
    full = X.astype(np.float64) @ W.astype(np.float64).T
    assert np.abs(Y - full).max() / np.abs(full).max() < 1e-4


def test_gemm_rejects_bad_shapes(dev):
    import token_hawk_b200 as t
    Kl = t.kernels()
    d = dev.empty(16)
    assert Kl.thk_gemm_f16_tc(dev.h, d.ptr, d.ptr, d.ptr, 8, 48, 64) == -1      # N % 32
    assert Kl.thk_gemm_f16_tc(dev.h, d.ptr, d.ptr, d.ptr, 8, 32, 96) == -1      # K % 64
    assert b"multiple of" in Kl.thk_last_error()


def test_two_contexts_on_one_device_do_not_share_operands(oracle):
    """the hi/lo split workspace used to be process-global per device, so two contexts on different
    streams overwrote each other's operands in flight.  Two contexts, different shapes, launches interleaved without any
    host synchronisation in between; both results must be right."""
    import token_hawk_b200 as t
    Kl = t.kernels()
    a, b = t.Device(0), t.Device(0)                       # two thk_ctx on device 0, each with its own non-blocking stream
    assert Kl.thk_gemm_reserve(a.h, 128, 4096) == 0 and Kl.thk_gemm_reserve(b.h, 64, 2048) == 0
    r = np.random.default_rng(11)
    Xa = r.standard_normal((128, 4096)).astype(np.float32); Wa = (r.standard_normal((1024, 4096)) * 0.02).astype(np.float16)
    Xb = r.standard_normal((64, 2048)).astype(np.float32); Wb = (r.standard_normal((512, 2048)) * 0.02).astype(np.float16)
    dXa, dWa, dYa = a.array(Xa), a.array(Wa), a.array(np.zeros((128, 1024), np.float32))
    dXb, dWb, dYb = b.array(Xb), b.array(Wb), b.array(np.zeros((64, 512), np.float32))
    for _ in range(20):
        assert Kl.thk_gemm_f16_tc(a.h, dXa.ptr, dWa.ptr, dYa.ptr, 128, 1024, 4096) == 0, Kl.thk_last_error()
        assert Kl.thk_gemm_f16_tc(b.h, dXb.ptr, dWb.ptr, dYb.ptr, 64, 512, 2048) == 0, Kl.thk_last_error()
    assert Kl.thk_gemm_check(a.h) == 0 and Kl.thk_gemm_check(b.h) == 0
    fa = Xa.astype(np.float64) @ Wa.astype(np.float64).T
    fb = Xb.astype(np.float64) @ Wb.astype(np.float64).T
    assert np.abs(dYa.numpy() - fa).max() / np.abs(fa).max() < 1e-4
    assert np.abs(dYb.numpy() - fb).max() / np.abs(fb).max() < 1e-4
    del dXa, dWa, dYa, dXb, dWb, dYb
    a.close(); b.close()
