"""Tensor-parallel wiring for the fused decoder: Megatron-style shards, one GPU per rank.

  Wq/Wk/Wv, W1/W3, output.weight : rows of the local heads / ff units / vocabulary slice
  Wo, W2                         : the matching COLUMNS (input dimension split)
  everything else                : replicated

The all-reduce after Wo and W2 is not an NCCL call: the decode kernel's epilogue writes each rank's
partial vector straight into every peer's exchange region over NVLink and the next phase sums them
(thk_decoder_set_peers).  This module only moves the region addresses between ranks:
  * same process, several devices : `LocalGroup`
  * one process per GPU (torchrun): `wire_distributed` (CUDA IPC handles through torch.distributed)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import token_hawk_b200 as th

ROW_SHARDED = ("attention.wq.weight", "attention.wk.weight", "attention.wv.weight", "feed_forward.w1.weight",
               "feed_forward.w3.weight", "output.weight")
COL_SHARDED = ("attention.wo.weight", "feed_forward.w2.weight")


def shard_kind(name: str) -> str:
    """'rows' | 'cols' | 'replicated' for a ggjt tensor name."""
    if any(name.endswith(s) for s in ROW_SHARDED):
        return "rows"
    if any(name.endswith(s) for s in COL_SHARDED):
        return "cols"
    return "replicated"


def shard(arr: np.ndarray, name: str, rank: int, size: int) -> np.ndarray:
    """The slice of a full tensor that rank `rank` of `size` holds (what the loader keeps on that GPU)."""
    kind = shard_kind(name)
    if kind == "rows":
        n = arr.shape[0] // size
        return np.ascontiguousarray(arr[rank * n:(rank + 1) * n])
    if kind == "cols":
        n = arr.shape[1] // size
        return np.ascontiguousarray(arr[:, rank * n:(rank + 1) * n])
    return arr


def merge_argmax(cands):
    """Cross-rank greedy argmax as the kernel does it: cands = [(value, global_id)] in rank order;
    highest value wins, lowest id on ties (th-llama.cpp:826-838 semantics over the whole vocabulary)."""
    best_v, best_i = None, -1
    for v, i in cands:
        if i < 0:
            continue
        if best_i < 0 or v > best_v or (v == best_v and i < best_i):
            best_v, best_i = v, i
    return best_i


def export_handle(model: "th.LlamaModel") -> bytes:
    ptr, _ = model.exchange_info()
    h = (C.c_ubyte * 64)()
    th._check(th.kernels().thk_ipc_export(model.dev.h, C.c_void_p(ptr), h))
    return bytes(h)


def import_handle(dev: "th.Device", handle: bytes) -> int:
    out = C.c_void_p()
    th._check(th.kernels().thk_ipc_import(dev.h, (C.c_ubyte * 64).from_buffer_copy(handle), C.byref(out)))
    return out.value


def wire_distributed(model: "th.LlamaModel", rank: int, world: int) -> None:
    """torchrun path: all-gather the CUDA IPC handles of the exchange regions and map the peers'."""
    import torch.distributed as dist
    handles = [None] * world
    dist.all_gather_object(handles, export_handle(model))
    ptrs = []
    for r in range(world):
        ptrs.append(model.exchange_info()[0] if r == rank else import_handle(model.dev, handles[r]))
    model.set_peers(ptrs)
    dist.barrier()


class LocalGroup:
    """N tensor-parallel ranks driven by one host thread (tests, single-process runs)."""

    def __init__(self, n: int, n_vocab, n_embd, n_mult, n_head, n_layer, n_ctx, seed=0x7B5EED, path: str | None = None):
        K = th.kernels()
        self.n = n
        self.devs = [th.Device(i) for i in range(n)]
        for a in self.devs:
            for b in self.devs:
                if a is not b:
                    th._check(K.thk_enable_peer_access(a.h, b.h))
        if path is None:
            self.models = [th.LlamaModel.synthetic(self.devs[r], n_vocab, n_embd, n_mult, n_head, n_layer, n_ctx, seed, tp_rank=r, tp_size=n)
                           for r in range(n)]
        else:
            self.models = [th.LlamaModel.load(self.devs[r], path, n_ctx, tp_rank=r, tp_size=n) for r in range(n)]
        ptrs = [m.exchange_info()[0] for m in self.models]
        for m in self.models:
            m.set_peers(ptrs)

    def eval(self, tokens, n_past: int):
        for m in self.models:                     # every rank is launched before any is waited for
            m.eval_launch(tokens, n_past)
        outs = [m.eval_finish() for m in self.models]
        toks = {t for t, _ in outs}
        assert len(toks) == 1, f"ranks disagree on the greedy token: {toks}"
        return outs[0][0], np.concatenate([l for _, l in outs])

    def fill_kv(self, n_positions: int, seed: int = 99):
        for m in self.models:
            m.fill_kv(n_positions, seed)

    def close(self):
        for m in self.models:
            m.close()
        for d in self.devs:
            d.close()
