"""CPU test of the persistent decode kernel's static tile schedule (thk_decoder_plan runs, on the host, the same RowIt /
PhaseDesc code the kernel runs): for every phase, shape, grid size and tensor-parallel split every weight row is covered
exactly once, row groups never straddle a segment or a CTA boundary, shares are even-aligned (RoPE pairs and W1/W3 pairs
stay in one lane pair), and the K tiling covers all columns in <= 32 KB tiles."""
import ctypes as C

import numpy as np
import pytest

import token_hawk_b200 as th


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_vocab", "n_embd", "n_head", "n_layer", "n_ff", "n_ctx", "tp_rank", "tp_size", "kv_f16")]


def plan(dims, phase, grid, cta):
    K = th.kernels()
    out = (C.c_int32 * 4096)()
    rc = K.thk_decoder_plan(C.byref(dims), phase, grid, cta, out, 4096)
    assert rc == 0, K.thk_last_error()
    n = out[4]
    groups = [(out[5 + 3 * i], out[6 + 3 * i], out[7 + 3 * i]) for i in range(n)]
    return dict(C=out[0], KT=out[1], CT=out[2], paired=out[3]), groups


SHAPES = [
    dict(n_vocab=32000, n_embd=4096, n_head=32, n_ff=11008),     # LLaMA-7B
    dict(n_vocab=32000, n_embd=5120, n_head=40, n_ff=13824),     # 13B
    dict(n_vocab=512, n_embd=512, n_head=8, n_ff=1536),          # tiny test model
    dict(n_vocab=1000, n_embd=256, n_head=4, n_ff=688),          # fewer rows than CTAs in some phases
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("tp", [1, 2, 4, 8])
@pytest.mark.parametrize("grid", [148, 132, 7])
def test_every_row_once(shape, tp, grid):
    if shape["n_head"] % tp or shape["n_ff"] % tp or shape["n_vocab"] % tp or (shape["n_ff"] // tp) % 8:
        pytest.skip("shape does not split this way")
    th.kernels().thk_decoder_plan.argtypes = [C.POINTER(Dims), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]
    d = Dims(n_layer=2, n_ctx=64, tp_rank=0, tp_size=tp, **shape)
    E, Eh, Fh, Vl = shape["n_embd"], shape["n_embd"] // tp, shape["n_ff"] // tp, shape["n_vocab"] // tp
    want = {0: ([Eh, Eh, Eh], E), 1: ([E], Eh), 2: ([Fh, Fh], E), 3: ([E], Fh), 4: ([Vl], E)}
    for phase, (seg_rows, Ccols) in want.items():
        seen = [np.zeros(r, dtype=np.int32) for r in seg_rows]
        for cta in range(grid):
            meta, groups = plan(d, phase, grid, cta)
            assert meta["C"] == Ccols
            assert meta["KT"] * meta["CT"] >= Ccols > (meta["KT"] - 1) * meta["CT"]       # K tiles cover the columns, none is empty
            assert meta["CT"] % 256 == 0 and 8 * meta["CT"] * 2 <= 32 * 1024              # 8 rows of a tile fit one 32 KB slot
            first = True
            for si, row0, nrows in groups:
                assert 1 <= nrows <= 8 and 0 <= si < len(seg_rows)
                assert row0 + nrows <= seg_rows[si]                                       # never past the segment
                if first and not (meta["paired"] and si):
                    assert row0 % 2 == 0 or len(seg_rows) > 1                              # even-aligned share
                first = False
                if meta["paired"]:                                                         # W1 / W3: the same rows of both
                    assert si == 0
                    seen[0][row0:row0 + nrows] += 1
                    seen[1][row0:row0 + nrows] += 1
                else:
                    seen[si][row0:row0 + nrows] += 1
        for s_ in seen:
            assert (s_ == 1).all(), (phase, int(s_.min()), int(s_.max()))


def test_plan_rejects_bad_arguments():
    K = th.kernels()
    K.thk_decoder_plan.argtypes = [C.POINTER(Dims), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]
    d = Dims(n_vocab=512, n_embd=512, n_head=8, n_layer=2, n_ff=1536, n_ctx=64, tp_rank=0, tp_size=1)
    out = (C.c_int32 * 64)()
    assert K.thk_decoder_plan(C.byref(d), 5, 148, 0, out, 64) != 0
    assert K.thk_decoder_plan(C.byref(d), 0, 148, 148, out, 64) != 0
    assert K.thk_decoder_plan(C.byref(d), 0, 1, 0, out, 64) != 0          # 768 rows in one CTA: 96 groups do not fit 64 ints
    assert K.thk_decoder_plan(None, 0, 148, 0, out, 64) != 0
