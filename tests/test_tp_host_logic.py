"""CPU tests for the tensor-parallel host logic: shard algebra against the oracle's ops, the cross-rank
argmax merge, and the rank wiring over a real 2-process gloo group (no GPU involved)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_kinds_and_partial_sums_reproduce_the_full_matvec(oracle):
    from token_hawk_b200 import tp
    assert tp.shard_kind("layers.0.attention.wq.weight") == "rows"
    assert tp.shard_kind("layers.31.attention.wo.weight") == "cols"
    assert tp.shard_kind("layers.5.feed_forward.w2.weight") == "cols"
    assert tp.shard_kind("layers.5.feed_forward.w1.weight") == "rows"
    assert tp.shard_kind("output.weight") == "rows"
    assert tp.shard_kind("layers.5.ffn_norm.weight") == tp.shard_kind("tok_embeddings.weight") == "replicated"
    r = np.random.default_rng(0)
    E, F = 512, 1536
    x = r.standard_normal(E).astype(np.float32)
    Wo = (r.standard_normal((E, E)) * 0.05).astype(np.float16)
    W1 = (r.standard_normal((F, E)) * 0.05).astype(np.float16)
    full_o = oracle.matvec_f16(x, Wo)
    full_1 = oracle.matvec_f16(x, W1)
    for g in (2, 4, 8):
        # column shards: every rank multiplies its slice of x; the all-reduce sums the partial vectors
        parts = [oracle.matvec_f16(x[k * E // g:(k + 1) * E // g], tp.shard(Wo, "layers.0.attention.wo.weight", k, g)) for k in range(g)]
        s = parts[0].copy()
        for p in parts[1:]:
            s = s + p                     # rank order, as the kernel's prologue adds them
        assert np.abs(s - full_o).max() < 1e-5 * np.abs(full_o).max()
        # row shards: concatenation, bit-exact
        rows = np.concatenate([oracle.matvec_f16(x, tp.shard(W1, "layers.0.feed_forward.w1.weight", k, g)) for k in range(g)])
        assert np.array_equal(rows, full_1)


def test_merge_argmax_semantics(oracle):
    from token_hawk_b200 import tp
    r = np.random.default_rng(1)
    for _ in range(50):
        logits = r.standard_normal(512).astype(np.float32)
        if _ % 5 == 0:
            i, j = sorted(r.choice(512, 2, replace=False).tolist())
            logits[i] = logits[j] = logits.max() + 1
        g = int(r.choice([2, 4, 8]))
        n = 512 // g
        cands = []
        for k in range(g):
            sl = logits[k * n:(k + 1) * n]
            cands.append((float(sl.max()), k * n + int(sl.argmax())))
        assert tp.merge_argmax(cands) == oracle.greedy(logits)
    assert tp.merge_argmax([(0.0, -1), (-3.0, 9)]) == 9


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    handles = [None] * world
    dist.all_gather_object(handles, bytes([rank]) * 64)         # stands in for the 64-byte CUDA IPC handle
    # each rank evaluates "its vocabulary slice" and the ranks agree on the global greedy token
    sys.path.insert(0, ROOT)
    from token_hawk_b200 import tp
    logits = np.random.default_rng(7).standard_normal(64).astype(np.float32)
    n = 64 // world
    sl = logits[rank * n:(rank + 1) * n]
    cands = [None] * world
    dist.all_gather_object(cands, (float(sl.max()), rank * n + int(sl.argmax())))
    out.put((rank, [h[0] for h in handles], tp.merge_argmax(cands), int(logits.argmax())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_wiring_over_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, seen, merged, want in res:
        assert seen == [0, 1]            # every rank holds every rank's handle, in rank order
        assert merged == want
