"""GPU parity for the f16 KV-cache option of the fused decode path (LlamaModel(kv_f16=True) /
thk_llama_dims.kv_f16).  Not a reference feature: the reference's cache is f32 (th-llama-loader.cpp:335).  K (after RoPE)
and V are rounded to f16, nearest even, when they are appended; everything else computes as before.

Parity budget, written here: against the oracle in the SAME mode (oracle.Model.set_kv_f16: rounds at append) the logits
agree like the f32 path does (<= 5e-5 relative, greedy ids identical); against the reference-exact f32 oracle the option
costs ~4e-4 relative on the logits (measured on the CPU oracle: 3.9e-4 tiny model, 4.2e-4 at 7B width), inside north_star's
1e-3 -- asserted below at 1e-3."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SAME_MODE_TOL = 5e-5
F32_ORACLE_TOL = 1e-3


@pytest.fixture(scope="module")
def th():
    import token_hawk_b200 as t
    return t


@pytest.fixture(scope="module")
def dev(th):
    d = th.Device(0)
    yield d
    d.close()


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def make_pair(th, dev, oracle, cfg, seed=0x7B5EED):
    g = th.LlamaModel.synthetic(dev, cfg.n_vocab, cfg.n_embd, cfg.n_mult, cfg.n_head, cfg.n_layer, cfg.n_ctx, seed, kv_f16=True)
    o16 = oracle.Model.synthetic(cfg, seed)
    o16.set_kv_f16(True)
    o32 = oracle.Model.synthetic(cfg, seed)
    return g, o16, o32


def test_tiny_model_f16_kv_steps(th, dev, oracle):
    cfg = oracle.TINY
    g, o16, o32 = make_pair(th, dev, oracle, cfg)
    toks = [1, 17, 400, 33, 2, 99, 257, 5, 5, 311, 48, 7, 500, 128, 64, 3, 77, 78, 79, 80]
    worst32 = 0.0
    for i, t in enumerate(toks):
        tok, logits = g.eval([t], i)
        ref = o16.eval([t], i)
        assert rel(logits, ref) < SAME_MODE_TOL, (i, rel(logits, ref))
        assert tok == oracle.greedy(ref) == oracle.greedy(logits)
        worst32 = max(worst32, rel(logits, o32.eval([t], i)))
    assert 1e-5 < worst32 < F32_ORACLE_TOL, worst32          # the option is really on, and inside the budget
    # the cache itself: f16 [head][pos][dim].  The rows are rounded from values that differ from the oracle's by the f32
    # path's own error (relative to the row's scale, not per element), so a small element can land a few f16 steps away:
    # almost all elements are bit-equal, none is further off than one f16 step of the largest element.
    H, D = cfg.n_head, cfg.n_embd // cfg.n_head
    for name, idx in (("key_cache_hpd", 0), ("value_cache_hpd", 1)):
        dev_kv = g.tensor(f"layers.1.{name}")
        assert dev_kv.dtype == np.float16
        dev_kv = dev_kv.reshape(H, cfg.n_ctx, D)[:, :len(toks)].astype(np.float32)
        ref_kv = o16.kv_cache(1)[idx][:len(toks)].transpose(1, 0, 2)        # [pos][head][dim] -> [head][pos][dim]
        assert np.abs(dev_kv - ref_kv).max() <= np.abs(ref_kv).max() * 2.0 ** -10
        assert (dev_kv == ref_kv).mean() > 0.97, (dev_kv == ref_kv).mean()
    g.close()


@pytest.mark.parametrize("n_ctx,n_past", [(512, 511), (1024, 700)])
def test_7b_width_f16_kv_long_context(th, dev, oracle, n_ctx, n_past):
    """attention over full and ragged 128-position f16 tiles, 4-8 KV splits per head, synthetic cache (rounded like an append)"""
    cfg = oracle.Config(n_vocab=32000, n_embd=4096, n_mult=256, n_head=32, n_layer=2, n_ctx=n_ctx)
    g, o16, o32 = make_pair(th, dev, oracle, cfg)
    g.fill_kv(n_past)
    o16.fill_kv_synthetic(n_past)
    o32.fill_kv_synthetic(n_past)
    tok, logits = g.eval([77], n_past)
    ref = o16.eval([77], n_past)
    assert rel(logits, ref) < SAME_MODE_TOL, rel(logits, ref)
    assert tok == oracle.greedy(ref)
    assert rel(logits, o32.eval([77], n_past)) < F32_ORACLE_TOL
    g.close()


def test_f16_kv_prefill_then_decode_then_opgraph(th, dev, oracle):
    """Rows cross between the op graph's f32 [pos][head][dim] cache and the fused f16 cache in both directions
    (thk_kv_to_hpd_f16 / thk_kv_from_hpd_f16).  The batched pass attends over its own unrounded f32 rows, so against the
    oracle (which rounds every row at append) the agreement is the option's own budget, not the same-mode one."""
    cfg = oracle.TINY
    g, o16, _ = make_pair(th, dev, oracle, cfg)
    prompt = [3, 9, 27, 81, 243, 11, 5, 6]
    tok, logits = g.eval(prompt, 0)                             # batched prefill through the op graph, then rows -> f16
    ref = o16.eval(prompt, 0)
    assert rel(logits, ref) < F32_ORACLE_TOL and tok == oracle.greedy(ref)
    n = len(prompt)
    for i in range(4):                                          # fused decode on top of the converted rows
        nxt = oracle.greedy(ref)
        tok, logits = g.eval([nxt], n + i)
        ref = o16.eval([nxt], n + i)
        assert rel(logits, ref) < F32_ORACLE_TOL and tok == oracle.greedy(ref), (i, rel(logits, ref))
    g.set_eval_path(th.EVAL_OPGRAPH)                            # the op graph continues: f16 rows widened back
    for i in range(4, 6):
        nxt = oracle.greedy(ref)
        tok, logits = g.eval([nxt], n + i)
        ref = o16.eval([nxt], n + i)
        assert rel(logits, ref) < F32_ORACLE_TOL and tok == oracle.greedy(ref), ("opgraph", i, rel(logits, ref))
    g.close()


def test_kv_copy_kernels_f16(th, dev):
    K = th.kernels()
    H, D, n_ctx, pos0, npos = 4, 64, 32, 5, 9
    r = np.random.default_rng(0)
    phd = r.standard_normal((n_ctx, H, D)).astype(np.float32)
    d_phd = dev.array(phd)
    d_hpd = dev.array(np.zeros((H, n_ctx, D), np.float16))
    assert K.thk_kv_to_hpd_f16(dev.h, d_phd.ptr, d_hpd.ptr, pos0, npos, n_ctx, H, D) == 0, K.thk_last_error()
    hpd = d_hpd.numpy()
    want = np.zeros((H, n_ctx, D), np.float16)
    want[:, pos0:pos0 + npos] = phd[pos0:pos0 + npos].transpose(1, 0, 2).astype(np.float16)      # numpy rounds to nearest even
    assert np.array_equal(hpd.view(np.uint16), want.view(np.uint16))
    d_back = dev.array(np.zeros((n_ctx, H, D), np.float32))
    assert K.thk_kv_from_hpd_f16(dev.h, d_hpd.ptr, d_back.ptr, pos0, npos, n_ctx, H, D) == 0, K.thk_last_error()
    back = d_back.numpy()
    assert np.array_equal(back[pos0:pos0 + npos], want[:, pos0:pos0 + npos].astype(np.float32).transpose(1, 0, 2))
    assert not back[:pos0].any() and not back[pos0 + npos:].any()
    d16 = dev.array(np.zeros(1000, np.float16))
    src = (r.standard_normal(1000) * 100).astype(np.float32)
    d32 = dev.array(src)
    assert K.thk_f32_to_f16(dev.h, d32.ptr, d16.ptr, 1000) == 0
    assert np.array_equal(d16.numpy().view(np.uint16), src.astype(np.float16).view(np.uint16))
