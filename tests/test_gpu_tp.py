"""GPU parity for the tensor-parallel path (BASELINE configs[4]): column/row-sharded matvecs with the
in-kernel one-shot all-reduce over peer memory, tp = 2/4/8, against the oracle and against tp = 1.
Needs >= 2 GPUs on one box (gpurun --gpus N); skipped otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("tp", [2, 4, 8])
def test_tp_matches_oracle_tiny(oracle, tp):
    if n_gpus() < tp:
        pytest.skip(f"needs {tp} GPUs")
    from token_hawk_b200 import tp as tpmod
    cfg = oracle.TINY                      # 8 heads, n_ff 1536, vocab 512: divisible by 2/4/8
    g = tpmod.LocalGroup(tp, cfg.n_vocab, cfg.n_embd, cfg.n_mult, cfg.n_head, cfg.n_layer, cfg.n_ctx)
    o = oracle.Model.synthetic(cfg)
    toks = [1, 17, 400, 33, 2, 99, 257, 5, 5, 311, 48, 7]
    for i, t in enumerate(toks):
        tok, logits = g.eval([t], i)
        ref = o.eval([t], i)
        assert logits.shape == ref.shape
        assert rel(logits, ref) < 1e-3 and rel(logits, ref) < 5e-5, (tp, i, rel(logits, ref))
        assert tok == oracle.greedy(ref)
    # greedy continuation under TP: every rank embeds the token the ranks agreed on inside the kernel
    o2 = oracle.Model.synthetic(cfg)
    ref_ids = o2.greedy_decode([1, 42, 300, 7], 20)
    g2 = tpmod.LocalGroup(tp, cfg.n_vocab, cfg.n_embd, cfg.n_mult, cfg.n_head, cfg.n_layer, cfg.n_ctx)
    for i, t in enumerate([1, 42, 300]):
        g2.eval([t], i)
    ids, tok = [], 7
    for i in range(20):
        tok, _ = g2.eval([tok], 3 + i)
        ids.append(tok)
    g2.close()
    assert ids == ref_ids
    g.close()


def test_tp2_7b_shapes_vs_single_gpu(oracle):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import token_hawk_b200 as th
    from token_hawk_b200 import tp as tpmod
    cfg = oracle.Config(n_layer=2, n_ctx=64)
    g = tpmod.LocalGroup(2, cfg.n_vocab, cfg.n_embd, cfg.n_mult, cfg.n_head, cfg.n_layer, cfg.n_ctx)
    single = th.LlamaModel.synthetic(g.devs[0], cfg.n_vocab, cfg.n_embd, cfg.n_mult, cfg.n_head, cfg.n_layer, cfg.n_ctx)
    o = oracle.Model.synthetic(cfg)
    for i, t in enumerate([1, 3000, 31999, 15]):
        tok, logits = g.eval([t], i)
        tok1, logits1 = single.eval([t], i)
        ref = o.eval([t], i)
        assert rel(logits, ref) < 5e-5 and rel(logits, logits1) < 1e-5 and tok == tok1 == oracle.greedy(ref)
    single.close()
    g.close()


def test_tp_loader_slices_match_oracle_shards(oracle, tmp_path):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from token_hawk_b200 import tp as tpmod
    cfg = oracle.TINY
    o = oracle.Model.synthetic(cfg, 77)
    path = str(tmp_path / "tiny.ggjt")
    o.write_ggjt(path)
    g = tpmod.LocalGroup(2, 0, 0, 0, 0, 0, cfg.n_ctx, path=path)
    for r, m in enumerate(g.models):
        for name in ("layers.0.attention.wq.weight", "layers.1.attention.wo.weight", "layers.0.feed_forward.w2.weight",
                     "layers.1.feed_forward.w3.weight", "output.weight", "norm.weight"):
            want = tpmod.shard(o.tensor(name), name, r, 2)
            assert np.array_equal(m.tensor(name).view(np.uint8), want.view(np.uint8)), (r, name)
    tok, logits = g.eval([5], 0)
    assert rel(logits, o.eval([5], 0)) < 5e-5
    g.close()
