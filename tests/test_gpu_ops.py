"""GPU parity, op level: every thk_* op kernel (one per reference cmdbuf_* op) against the oracle on
the same seeded inputs, called through the C ABI.  Tolerances: elementwise ops bit-exact or 1 ulp-ish;
reductions 1e-5 relative to the row's absolute-value dot product (only summation order differs)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import token_hawk_b200 as t
    d = t.Device(0)
    yield d
    d.close()


@pytest.fixture(scope="module")
def K():
    import token_hawk_b200 as t
    return t.kernels()


def _ok(rc, K):
    assert rc == 0, K.thk_last_error().decode()


def rng(seed):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("R,C", [(4096, 4096), (11008, 4096), (4096, 11008), (32000, 4096), (40, 264), (7, 8), (512, 1376)])
def test_vector_mat_mul_trans_f16(dev, K, oracle, R, C):
    r = rng(R + C)
    W = (r.standard_normal((R, C)) * 0.02).astype(np.float16)
    x = r.standard_normal(C).astype(np.float32)
    dW, dx, dy = dev.array(W), dev.array(x), dev.empty(R)
    _ok(K.thk_vector_mat_mul_trans(dev.h, dx.ptr, 0, dW.ptr, dy.ptr, R, C, 1, 1), K)
    y = dy.numpy()
    oracle.lib().tho_set_strict_order(1)
    ref = oracle.matvec_f16(x, W)
    bound = np.abs(W.astype(np.float32)) @ np.abs(x)
    assert np.max(np.abs(y - ref) / bound) < 2e-6
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-5


def test_vector_mat_mul_trans_f32_batch_offset(dev, K, oracle):
    r = rng(5)
    R, Cc, B = 96, 512, 3
    W = r.standard_normal((B, R, Cc)).astype(np.float32)
    x = r.standard_normal((B + 1, Cc)).astype(np.float32)
    dW, dx, dy = dev.array(W), dev.array(x), dev.empty((B, R))
    _ok(K.thk_vector_mat_mul_trans(dev.h, dx.ptr, Cc * 4, dW.ptr, dy.ptr, R, Cc, B, 0), K)   # aOffset skips row 0
    y = dy.numpy()
    for b in range(B):
        ref = oracle.matvec_f32(x[b + 1], W[b])
        assert np.abs(y[b] - ref).max() < 1e-4 * np.abs(ref).max()


def test_vector_mat_mul_trans_rejects_bad_args(dev, K):
    dx, dW, dy = dev.empty(12), dev.empty((4, 12), np.float16), dev.empty(4)
    assert K.thk_vector_mat_mul_trans(dev.h, dx.ptr, 0, dW.ptr, dy.ptr, 4, 12, 1, 1) == -1   # C % 8 != 0
    assert b"C % 8" in K.thk_last_error()
    assert K.thk_vector_mat_mul_trans(dev.h, None, 0, dW.ptr, dy.ptr, 4, 16, 1, 1) == -1
    assert K.thk_vector_mat_mul_trans(dev.h, dx.ptr, 4, dW.ptr, dy.ptr, 4, 8, 1, 1) == -1      # misaligned aOffset


def test_split_matvec_computes_full_logits(dev, K, oracle):
    """Reference F3 defect: 1280 logits miss the second half.  The CUDA path returns the full sum."""
    r = rng(6)
    V, E = 32000, 512
    W = (r.standard_normal((V, E)) * 0.05).astype(np.float16)
    x = r.standard_normal(E).astype(np.float32)
    h1, h2 = np.ascontiguousarray(W[:, :E // 2]), np.ascontiguousarray(W[:, E // 2:])
    d1, d2, dx, dc, ds = dev.array(h1), dev.array(h2), dev.array(x), dev.empty(V), dev.empty(V)
    arr = (C.c_void_p * 2)(d1.ptr, d2.ptr)
    _ok(K.thk_vector_multi_mat_mul_split_trans(dev.h, dx.ptr, 0, arr, 2, dc.ptr, ds.ptr, V, E, 1), K)
    full = oracle.matvec_f16(x, W)
    got = dc.numpy()
    assert np.abs(got - full).max() < 1e-5 * np.abs(full).max() + 1e-6
    bug_ids = np.nonzero(np.arange(V) % 4000 >= 3840)[0]
    assert len(bug_ids) == 1280 and np.abs(got[bug_ids] - full[bug_ids]).max() < 1e-4


def test_rms_norm_gain(dev, K, oracle):
    r = rng(7)
    x = (r.standard_normal((3, 4096)) * 2).astype(np.float32)
    g = oracle.fill_gain(1, 5, 4096)
    dx, dg = dev.array(x), dev.array(g)
    _ok(K.thk_rms_norm(dev.h, dx.ptr, 3, 4096), K)
    y = dx.numpy()
    ref = oracle.rms_norm(x)
    assert np.abs(y - ref).max() < 2e-6 * np.abs(ref).max()
    _ok(K.thk_row_element_multiply(dev.h, dx.ptr, dg.ptr, 3, 4096), K)
    assert np.array_equal(dx.numpy(), y * g)


@pytest.mark.parametrize("n_past", [0, 1, 37, 511, 2047])
def test_rope(dev, K, oracle, n_past):
    import token_hawk_b200 as t
    x = rng(8).standard_normal((2, 32, 128)).astype(np.float32)
    dx = dev.array(x)
    u = t.device_struct(dev, t.NetworkUniforms(n_past=n_past, n_tokens=2))
    _ok(K.thk_rope(dev.h, dx.ptr, 2, 32, 128, u.ptr), K)
    ref = oracle.rope(x, n_past)
    # libm differences (powf/sinf/cosf): angles up to ~2e3 rad, 1 ulp of the angle is ~1e-4 absolute
    assert np.abs(dx.numpy() - ref).max() < 2e-4 * max(1.0, n_past / 512)


def test_transpose_both_kinds_and_uniform_dims(dev, K, oracle):
    import token_hawk_b200 as t
    a = rng(9).standard_normal((5, 8, 64)).astype(np.float32)
    da, dzy, dyx = dev.array(a), dev.empty((8, 5, 64)), dev.empty((5, 64, 8))
    _ok(K.thk_transpose(dev.h, da.ptr, dzy.ptr, 5, 8, 64, 1, None), K)
    _ok(K.thk_transpose(dev.h, da.ptr, dyx.ptr, 5, 8, 64, 0, None), K)
    assert np.array_equal(dzy.numpy(), oracle.transpose(a, True))
    assert np.array_equal(dyx.numpy(), oracle.transpose(a, False))
    # dims from the uniform block: transpose only the first 3 "positions" of a cache-like buffer
    u = t.device_struct(dev, t.DimsUniforms(A_B=3, A_M=8, A_N=64))
    out = dev.array(np.zeros((8, 3, 64), np.float32))
    _ok(K.thk_transpose(dev.h, da.ptr, out.ptr, 5, 8, 64, 1, u.ptr), K)
    assert np.array_equal(out.numpy(), a[:3].transpose(1, 0, 2))
    assert K.thk_transpose(dev.h, da.ptr, da.ptr, 5, 8, 64, 1, None) == -1


def test_attention_ops_chain(dev, K, oracle):
    """QK^T*scale -> row softmax -> PV exactly as th-llama.cpp:365-380 issues them (uniform dims)."""
    import token_hawk_b200 as t
    r = rng(10)
    H, D, N = 32, 128, 200
    q = r.standard_normal((H, 1, D)).astype(np.float32)
    Kc = r.standard_normal((H, N, D)).astype(np.float32)
    Vc = r.standard_normal((H, N, D)).astype(np.float32)
    dq, dk, dv, ds, do = dev.array(q), dev.array(Kc), dev.array(Vc), dev.empty((H, 1, N)), dev.empty((H, 1, D))
    scale = np.float32(1.0) / np.sqrt(np.float32(D))
    u2 = t.device_struct(dev, t.DimsUniforms(A_B=H, A_M=1, A_N=D, scale=float(scale), B_B=H, B_M=D, B_N=N))
    u3 = t.device_struct(dev, t.DimsUniforms(A_B=H, A_M=1, A_N=N))
    u4 = t.device_struct(dev, t.DimsUniforms(A_B=H, A_M=1, A_N=N, scale=1.0, B_B=H, B_M=N, B_N=D))
    _ok(K.thk_mat_mul(dev.h, dq.ptr, dk.ptr, ds.ptr, H, 1, D, N, 1, 0, u2.ptr), K)
    s_ref = oracle.mat_mul(q, Kc, True, scale=float(scale))
    assert np.abs(ds.numpy() - s_ref).max() < 1e-5
    _ok(K.thk_row_softmax(dev.h, ds.ptr, H, 1, N, u3.ptr), K)
    p_ref = oracle.row_softmax(s_ref)
    assert np.abs(ds.numpy() - p_ref).max() < 1e-6
    _ok(K.thk_mat_mul(dev.h, ds.ptr, dv.ptr, do.ptr, H, 1, N, D, 0, 0, u4.ptr), K)
    o_ref = oracle.mat_mul(p_ref, Vc, False, scale=1.0)
    assert np.abs(do.numpy() - o_ref).max() < 1e-5


@pytest.mark.parametrize("B,M,Kd,N", [(3, 20, 33, 70), (32, 128, 128, 128), (2, 64, 16, 64), (1, 8, 200, 5), (4, 130, 64, 129)])
@pytest.mark.parametrize("transposeB", [1, 0])
@pytest.mark.parametrize("use_uniforms", [False, True])
def test_mat_mul_f32_tiled_batch_path(dev, K, oracle, B, M, Kd, N, transposeB, use_uniforms):
    """the attention matmuls of a batched prompt pass (M >= 8, f32 B): 64 x 64 tiled kernel, ragged edges in every dimension"""
    import token_hawk_b200 as t
    r = rng(B * 1000 + M + N + Kd)
    A = r.standard_normal((B, M, Kd)).astype(np.float32)
    Bm = r.standard_normal((B, N, Kd) if transposeB else (B, Kd, N)).astype(np.float32)
    dA, dB = dev.array(A), dev.array(Bm)
    dC = dev.array(np.full((B, M, N), np.nan, np.float32))
    scale = 0.125 if use_uniforms else None
    u = t.device_struct(dev, t.DimsUniforms(A_B=B, A_M=M, A_N=Kd, scale=0.125, B_B=B, B_M=Kd, B_N=N)) if use_uniforms else None
    _ok(K.thk_mat_mul(dev.h, dA.ptr, dB.ptr, dC.ptr, B, M, Kd, N, transposeB, 0, u.ptr if u else None), K)
    ref = oracle.mat_mul(A, Bm, bool(transposeB), scale=scale)
    got = dC.numpy()
    bound = np.abs(A).astype(np.float64) @ (np.abs(Bm).astype(np.float64).transpose(0, 2, 1) if transposeB else np.abs(Bm).astype(np.float64))
    assert np.isfinite(got).all()
    assert (np.abs(got - ref) <= 2e-6 * bound * (0.125 if use_uniforms else 1.0) + 1e-7).all()


def test_mat_mul_f16_weights_batch_path(dev, K, oracle):
    r = rng(11)
    M, Kd, N = 8, 512, 96
    A = r.standard_normal((1, M, Kd)).astype(np.float32)
    W = (r.standard_normal((1, N, Kd)) * 0.05).astype(np.float16)
    dA, dW, dC = dev.array(A), dev.array(W), dev.empty((1, M, N))
    _ok(K.thk_mat_mul(dev.h, dA.ptr, dW.ptr, dC.ptr, 1, M, Kd, N, 1, 1, None), K)
    ref = oracle.mat_mul(A, W, True)
    assert np.abs(dC.numpy() - ref).max() < 1e-5 * np.abs(ref).max() + 1e-6


def test_masked_softmax_is_causal_with_n_past(dev, K, oracle):
    a = rng(12).standard_normal((4, 8, 13)).astype(np.float32)   # M=8 queries, N=13 -> n_past=5
    da = dev.array(a)
    _ok(K.thk_masked_softmax(dev.h, da.ptr, 4, 8, 13, None), K)
    ref = oracle.causal_softmax(a, 5)
    assert np.abs(da.numpy() - ref).max() < 1e-6


def test_elementwise(dev, K, oracle):
    r = rng(13)
    n = 11008
    a = (r.standard_normal(n) * 4).astype(np.float32)
    b = r.standard_normal(n).astype(np.float32)
    da, db, dc = dev.array(a), dev.array(b), dev.empty(n)
    _ok(K.thk_addition(dev.h, da.ptr, db.ptr, dc.ptr, n), K)
    assert np.array_equal(dc.numpy(), a + b)
    _ok(K.thk_silu(dev.h, da.ptr, n), K)
    s_ref = oracle.silu(a)
    assert np.abs(da.numpy() - s_ref).max() < 1e-6 * np.abs(s_ref).max()
    s = da.numpy()
    _ok(K.thk_element_mult_in_place(dev.h, da.ptr, db.ptr, n), K)
    assert np.array_equal(da.numpy(), s * b)
    _ok(K.thk_vector_reduce(dev.h, dc.ptr, db.ptr, n), K)
    assert np.array_equal(dc.numpy(), (a + b) + b)


def test_f16_f32_conversion_all_codes(dev, K, oracle):
    codes = np.arange(65536, dtype=np.uint16)
    din, dout = dev.array(codes), dev.empty(65536)
    _ok(K.thk_f16_f32_conversion(dev.h, dout.ptr, 0, din.ptr, 0, 65536), K)
    got = dout.numpy().view(np.uint32)
    ref = oracle.fp16_to_fp32_table().view(np.uint32)
    nan = np.isnan(ref.view(np.float32))
    assert np.array_equal(got[~nan], ref[~nan])       # cvt == the reference's bit trick, every code
    assert np.isnan(got.view(np.float32)[nan]).all()
    # row offsets (bytes), as the embedding gather uses them
    _ok(K.thk_f16_f32_conversion(dev.h, dout.ptr, 16, din.ptr, 2 * 1024, 8), K)
    assert np.array_equal(dout.numpy()[4:12].view(np.uint32), ref[1024:1032])


def test_synthetic_fill_matches_oracle(dev, K, oracle):
    d = dev.empty((64, 256), np.uint16)
    _ok(K.thk_fill_f16(dev.h, d.ptr, 123, 9, 64, 256, 32, 512, 4096), K)
    assert np.array_equal(d.numpy(), oracle.fill_f16(123, 9, 64, 256, row0=32, col0=512, full_cols=4096).view(np.uint16))
    g = dev.empty(4096)
    _ok(K.thk_fill_gain(dev.h, g.ptr, 123, 4, 4096), K)
    assert np.array_equal(g.numpy(), oracle.fill_gain(123, 4, 4096))
    H, D, n_ctx, npos = 8, 64, 32, 5
    kv = dev.array(np.zeros((H, n_ctx, D), np.float32))
    _ok(K.thk_fill_kv(dev.h, kv.ptr, 99, 1000, npos, n_ctx, H, 0, H, D), K)
    ref = oracle.fill_kv(99, 1000, npos * H * D).reshape(npos, H, D)
    got = kv.numpy()
    assert np.array_equal(got[:, :npos], ref.transpose(1, 0, 2)) and np.all(got[:, npos:] == 0)
