"""ctypes binding of oracle/libth_oracle.so (the CPU restatement of the reference's decode path).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (token_hawk_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libth_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libthref_host.so")


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "th_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "libth_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/th.cpp"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def _cap_threads():
    """libgomp with one thread per logical CPU is pathological on large hosts (128 logical CPUs on the GPU
    box: 20x slower than 64 threads, measured).  Cap the default; OMP_NUM_THREADS still overrides."""
    if "OMP_NUM_THREADS" not in os.environ:
        n = os.cpu_count() or 1
        try:
            n = len(os.sched_getaffinity(0))
        except Exception:
            pass
        os.environ["OMP_NUM_THREADS"] = str(max(1, min(64, n // 2 if n > 16 else n)))
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _cap_threads()
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        f32p, u16p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_uint16), C.POINTER(C.c_int32)
        i64 = C.c_int64
        L.tho_fp16_to_fp32.restype = C.c_float
        L.tho_fp16_to_fp32.argtypes = [C.c_uint16]
        L.tho_fp32_to_fp16.restype = C.c_uint16
        L.tho_fp32_to_fp16.argtypes = [C.c_float]
        L.tho_vector_mat_mul_trans_f16.argtypes = [f32p, u16p, f32p, i64, i64, i64]
        L.tho_vector_mat_mul_trans_f32.argtypes = [f32p, f32p, f32p, i64, i64, i64]
        L.tho_rms_norm.argtypes = [f32p, i64, i64]
        L.tho_row_element_multiply.argtypes = [f32p, f32p, i64, i64]
        L.tho_rope.argtypes = [f32p, i64, i64, i64, C.c_uint32]
        L.tho_transpose.argtypes = [f32p, f32p, i64, i64, i64, C.c_int]
        L.tho_mat_mul.argtypes = [f32p, C.c_void_p, f32p, i64, i64, i64, i64, C.c_int, C.c_int,
                                  C.c_float, C.c_int]
        L.tho_row_softmax.argtypes = [f32p, i64, i64, i64]
        L.tho_causal_softmax.argtypes = [f32p, i64, i64, i64, i64]
        L.tho_addition.argtypes = [f32p, f32p, f32p, i64]
        L.tho_silu.argtypes = [f32p, i64]
        L.tho_element_mult_in_place.argtypes = [f32p, f32p, i64]
        L.tho_vector_reduce.argtypes = [f32p, f32p, i64, C.c_int, C.c_int]
        L.tho_f16_f32_conversion.argtypes = [f32p, u16p, i64]
        L.tho_greedy.restype = C.c_int32
        L.tho_greedy.argtypes = [f32p, C.c_int32]
        L.tho_hash.restype = C.c_uint64
        L.tho_hash.argtypes = [C.c_uint64] * 3
        L.tho_fill_f16.argtypes = [u16p, C.c_uint64, C.c_uint64, i64, i64, i64, i64, i64]
        L.tho_fill_gain.argtypes = [f32p, C.c_uint64, C.c_uint64, i64]
        L.tho_fill_kv.argtypes = [f32p, C.c_uint64, C.c_uint64, i64]
        L.tho_n_ff.restype = C.c_int32
        L.tho_n_ff.argtypes = [C.c_void_p]
        L.tho_model_create.restype = C.c_void_p
        L.tho_model_create.argtypes = [C.c_void_p]
        L.tho_model_free.argtypes = [C.c_void_p]
        L.tho_model_set_tensor.restype = C.c_int
        L.tho_model_set_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, i64, i64]
        L.tho_model_get_tensor.restype = C.c_void_p
        L.tho_model_get_tensor.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int),
                                           C.POINTER(i64), C.POINTER(i64)]
        L.tho_model_ready.restype = C.c_int
        L.tho_model_ready.argtypes = [C.c_void_p]
        L.tho_model_fill_synthetic.argtypes = [C.c_void_p, C.c_uint64]
        L.tho_model_fill_kv_synthetic.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.tho_model_reset.argtypes = [C.c_void_p]
        L.tho_model_key_cache.restype = f32p
        L.tho_model_key_cache.argtypes = [C.c_void_p, C.c_int]
        L.tho_model_value_cache.restype = f32p
        L.tho_model_value_cache.argtypes = [C.c_void_p, C.c_int]
        L.tho_eval.restype = C.c_int
        L.tho_eval.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, f32p, f32p]
        L.tho_load_ggjt.restype = C.c_void_p
        L.tho_load_ggjt.argtypes = [C.c_char_p, C.c_int32]
        L.tho_write_ggjt.restype = C.c_int
        L.tho_write_ggjt.argtypes = [C.c_void_p, C.c_char_p]
        L.tho_model_hparams.restype = C.c_void_p
        L.tho_model_hparams.argtypes = [C.c_void_p]
        L.tho_num_threads.restype = C.c_int
        L.tho_set_num_threads.argtypes = [C.c_int]
        L.tho_set_strict_order.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u16(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


class HParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n_vocab", "n_embd", "n_mult", "n_head", "n_layer", "n_rot", "f16", "n_ctx")]


@dataclass
class Config:
    n_vocab: int = 32000
    n_embd: int = 4096
    n_mult: int = 256
    n_head: int = 32
    n_layer: int = 32
    n_ctx: int = 512

    @property
    def head_dim(self):
        return self.n_embd // self.n_head

    @property
    def n_ff(self):
        return ((2 * (4 * self.n_embd) // 3 + self.n_mult - 1) // self.n_mult) * self.n_mult

    def hparams(self) -> HParams:
        return HParams(self.n_vocab, self.n_embd, self.n_mult, self.n_head, self.n_layer,
                       self.head_dim, 1, self.n_ctx)


TINY = Config(n_vocab=512, n_embd=512, n_mult=256, n_head=8, n_layer=2, n_ctx=64)
LLAMA_7B = Config()

LAYER_TENSORS = ("attention_norm.weight", "attention.wq.weight", "attention.wk.weight",
                 "attention.wv.weight", "attention.wo.weight", "ffn_norm.weight",
                 "feed_forward.w1.weight", "feed_forward.w2.weight", "feed_forward.w3.weight")


def tensor_names(n_layer: int):
    names = ["tok_embeddings.weight", "norm.weight", "output.weight"]
    for l in range(n_layer):
        names += [f"layers.{l}.{s}" for s in LAYER_TENSORS]
    return names


def tensor_id(name: str) -> int:
    """Synthetic-weight tensor id (th_oracle.c: 0 emb, 1 norm, 2 output, 3 + 9*l + k)."""
    if name == "tok_embeddings.weight":
        return 0
    if name == "norm.weight":
        return 1
    if name == "output.weight":
        return 2
    _, l, rest = name.split(".", 2)
    return 3 + 9 * int(l) + LAYER_TENSORS.index(rest)


class Model:
    """Owns a tho_model."""

    def __init__(self, cfg: Config | None = None, handle=None):
        L = lib()
        if handle is not None:
            self.h = handle
            hp = HParams.from_address(L.tho_model_hparams(handle))
            self.cfg = Config(hp.n_vocab, hp.n_embd, hp.n_mult, hp.n_head, hp.n_layer, hp.n_ctx)
        else:
            self.cfg = cfg
            hp = cfg.hparams()
            self.h = L.tho_model_create(C.byref(hp))

    def __del__(self):
        try:
            if self.h:
                lib().tho_model_free(self.h)
                self.h = None
        except Exception:
            pass

    @classmethod
    def synthetic(cls, cfg: Config, seed: int = 0x7B5EED) -> "Model":
        m = cls(cfg)
        lib().tho_model_fill_synthetic(m.h, seed)
        return m

    @classmethod
    def load(cls, path: str, n_ctx: int = 512) -> "Model":
        h = lib().tho_load_ggjt(path.encode(), n_ctx)
        if not h:
            raise IOError(f"tho_load_ggjt failed: {path}")
        return cls(handle=h)

    def write_ggjt(self, path: str) -> None:
        rc = lib().tho_write_ggjt(self.h, path.encode())
        if rc:
            raise IOError(f"tho_write_ggjt rc={rc}")

    def set_tensor(self, name: str, arr: np.ndarray) -> None:
        arr = np.ascontiguousarray(arr)
        ftype = {np.dtype(np.float32): 0, np.dtype(np.float16): 1, np.dtype(np.uint16): 1}[arr.dtype]
        rows, cols = (1, arr.shape[0]) if arr.ndim == 1 else arr.shape
        rc = lib().tho_model_set_tensor(self.h, name.encode(), arr.ctypes.data_as(C.c_void_p), ftype,
                                        rows, cols)
        if rc:
            raise ValueError(f"set_tensor({name}) rc={rc}")

    def tensor(self, name: str) -> np.ndarray:
        ft, r, c = C.c_int(), C.c_int64(), C.c_int64()
        p = lib().tho_model_get_tensor(self.h, name.encode(), C.byref(ft), C.byref(r), C.byref(c))
        if not p:
            raise KeyError(name)
        dt = np.float16 if ft.value == 1 else np.float32
        n = r.value * c.value
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt).reshape(r.value, c.value)

    def set_kv_f16(self, on: bool = True):
        """product option (not the reference's): K/V rounded to f16 when appended; twin of LlamaModel(kv_f16=True)"""
        L = lib()
        L.tho_model_set_kv_f16.restype = None
        L.tho_model_set_kv_f16.argtypes = [C.c_void_p, C.c_int]
        L.tho_model_set_kv_f16(self.h, 1 if on else 0)

    def reset(self):
        lib().tho_model_reset(self.h)

    def fill_kv_synthetic(self, n_positions: int, seed: int = 99):
        lib().tho_model_fill_kv_synthetic(self.h, seed, n_positions)

    def kv_cache(self, layer: int):
        n = self.cfg.n_ctx * self.cfg.n_embd
        shape = (self.cfg.n_ctx, self.cfg.n_head, self.cfg.head_dim)
        k = np.ctypeslib.as_array(lib().tho_model_key_cache(self.h, layer), shape=(n,)).reshape(shape)
        v = np.ctypeslib.as_array(lib().tho_model_value_cache(self.h, layer), shape=(n,)).reshape(shape)
        return k, v

    def eval(self, tokens, n_past: int, want_hidden: bool = False):
        toks = np.ascontiguousarray(np.asarray(tokens, dtype=np.int32).reshape(-1))
        logits = np.empty(self.cfg.n_vocab, dtype=np.float32)
        hidden = np.empty((self.cfg.n_layer + 1, self.cfg.n_embd), dtype=np.float32) if want_hidden else None
        rc = lib().tho_eval(self.h, toks.ctypes.data_as(C.POINTER(C.c_int32)), len(toks), n_past,
                            _f32(logits), _f32(hidden) if want_hidden else None)
        if rc:
            raise RuntimeError(f"tho_eval rc={rc}")
        return (logits, hidden) if want_hidden else logits

    def greedy_decode(self, prompt, n_new: int, n_past: int = 0):
        """Feeds prompt tokens one at a time, then n_new greedy steps; returns generated ids."""
        out, tok = [], None
        for t in prompt:
            logits = self.eval([t], n_past)
            n_past += 1
        tok = greedy(logits)
        for _ in range(n_new):
            out.append(tok)
            logits = self.eval([tok], n_past)
            n_past += 1
            tok = greedy(logits)
        return out


# ---- op wrappers (numpy in / numpy out) ----
def fp16_to_fp32_table() -> np.ndarray:
    L = lib()
    return np.array([L.tho_fp16_to_fp32(i) for i in range(65536)], dtype=np.float32)


def matvec_f16(x, W):
    x = np.ascontiguousarray(x, np.float32)
    W = np.ascontiguousarray(W).view(np.uint16)
    R, Cc = W.shape[-2], W.shape[-1]
    batch = 1 if W.ndim == 2 else W.shape[0]
    y = np.empty((batch, R) if W.ndim == 3 else (R,), np.float32)
    lib().tho_vector_mat_mul_trans_f16(_f32(x), _u16(W), _f32(y), R, Cc, batch)
    return y


def matvec_f32(x, W):
    x = np.ascontiguousarray(x, np.float32)
    W = np.ascontiguousarray(W, np.float32)
    R, Cc = W.shape
    y = np.empty(R, np.float32)
    lib().tho_vector_mat_mul_trans_f32(_f32(x), _f32(W), _f32(y), R, Cc, 1)
    return y


def rms_norm(x):
    x = np.array(x, np.float32, copy=True, order="C")
    rows, N = (1, x.shape[0]) if x.ndim == 1 else x.shape
    lib().tho_rms_norm(_f32(x), rows, N)
    return x


def row_element_multiply(x, g):
    x = np.array(x, np.float32, copy=True, order="C")
    g = np.ascontiguousarray(g, np.float32)
    rows, N = (1, x.shape[0]) if x.ndim == 1 else x.shape
    lib().tho_row_element_multiply(_f32(x), _f32(g), rows, N)
    return x


def rope(x, n_past: int):
    """x: [n_tokens, n_head, head_dim]"""
    x = np.array(x, np.float32, copy=True, order="C")
    t, h, d = x.shape
    lib().tho_rope(_f32(x), t, h, d, n_past)
    return x


def transpose(a, zy: bool):
    a = np.ascontiguousarray(a, np.float32)
    B, M, N = a.shape
    c = np.empty((M, B, N) if zy else (B, N, M), np.float32)
    lib().tho_transpose(_f32(a), _f32(c), B, M, N, int(zy))
    return c


def mat_mul(A, Bm, transposeB: bool, scale=None):
    A = np.ascontiguousarray(A, np.float32)
    is_f16 = Bm.dtype in (np.float16, np.uint16)
    Bm = np.ascontiguousarray(Bm)
    batch, M, K = A.shape
    N = Bm.shape[1] if transposeB else Bm.shape[2]
    out = np.empty((batch, M, N), np.float32)
    lib().tho_mat_mul(_f32(A), Bm.ctypes.data_as(C.c_void_p), _f32(out), batch, M, K, N,
                      int(transposeB), int(scale is not None), float(scale or 1.0), int(is_f16))
    return out


def row_softmax(a):
    a = np.array(a, np.float32, copy=True, order="C")
    b, m, n = a.shape
    lib().tho_row_softmax(_f32(a), b, m, n)
    return a


def causal_softmax(a, n_past: int):
    a = np.array(a, np.float32, copy=True, order="C")
    b, m, n = a.shape
    lib().tho_causal_softmax(_f32(a), b, m, n, n_past)
    return a


def addition(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    c = np.empty_like(a)
    lib().tho_addition(_f32(a), _f32(b), _f32(c), a.size)
    return c


def silu(a):
    a = np.array(a, np.float32, copy=True, order="C")
    lib().tho_silu(_f32(a), a.size)
    return a


def element_mult(a, b):
    a = np.array(a, np.float32, copy=True, order="C")
    b = np.ascontiguousarray(b, np.float32)
    lib().tho_element_mult_in_place(_f32(a), _f32(b), a.size)
    return a


def vector_reduce(a, b, num_splits=8, refbug=False):
    a = np.array(a, np.float32, copy=True, order="C")
    b = np.ascontiguousarray(b, np.float32)
    lib().tho_vector_reduce(_f32(a), _f32(b), a.size, num_splits, int(refbug))
    return a


def eval_trace(model: "Model", token: int, n_past: int):
    """The commands tho_eval issues for one token, as lists: ["op", label, [operands], [uniform words]] or
    ["copy", src, dst, dst_off, size] -- reference labels and buffer names (th_oracle.c, TR lines)."""
    L = lib()
    L.tho_trace_begin.restype = None
    L.tho_trace_end.restype = C.c_char_p
    L.tho_trace_begin()
    model.eval([token], n_past)
    out = []
    for ln in L.tho_trace_end().decode().splitlines():
        f = ln.split("|")
        if f[0] == "op":
            out.append(["op", f[1], f[2].split(",") if f[2] else [], [int(x) for x in f[3].split(",")] if f[3] else []])
        else:
            out.append(["copy", f[1], f[2], int(f[3]), int(f[4])])
    return out


def greedy(logits) -> int:
    logits = np.ascontiguousarray(logits, np.float32)
    return int(lib().tho_greedy(_f32(logits), logits.size))


def fill_f16(seed, tid, rows, cols, row0=0, col0=0, full_cols=None):
    out = np.empty((rows, cols), np.uint16)
    lib().tho_fill_f16(_u16(out), seed, tid, rows, cols, row0, col0, full_cols or cols)
    return out.view(np.float16)


def fill_gain(seed, tid, n):
    out = np.empty(n, np.float32)
    lib().tho_fill_gain(_f32(out), seed, tid, n)
    return out


def fill_kv(seed, tid, n):
    out = np.empty(n, np.float32)
    lib().tho_fill_kv(_f32(out), seed, tid, n)
    return out


def num_threads() -> int:
    return lib().tho_num_threads()


def use_all_cores() -> int:
    """Run the OpenMP loops on every core this process may use, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to every rank).  Returns the thread count in effect."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().tho_set_num_threads(int(n))
    return num_threads()


# ---- oracle/_ref: the reference's own host code (present only where it was built) ----
_ref = None


def ref_lib():
    """libthref_host.so (real reference host code on the WebGPU stub) or None."""
    global _ref
    if _ref is None and os.path.exists(_REF_PATH):
        R = C.CDLL(_REF_PATH)
        R.ref_fp16_to_fp32.restype = C.c_float
        R.ref_fp16_to_fp32.argtypes = [C.c_uint16]
        R.ref_fp32_to_fp16.restype = C.c_uint16
        R.ref_fp32_to_fp16.argtypes = [C.c_float]
        R.ref_greedy.restype = C.c_int
        R.ref_greedy.argtypes = [C.POINTER(C.c_float), C.c_int]
        R.ref_load.restype = C.c_void_p
        R.ref_load.argtypes = [C.c_char_p]
        R.ref_free.argtypes = [C.c_void_p]
        R.ref_hparams.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        R.ref_tensor.restype = C.c_void_p
        R.ref_tensor.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64)]
        R.ref_tokenize.restype = C.c_int
        R.ref_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.c_int]
        R.ref_vocab_model.restype = C.c_void_p
        R.ref_vocab_model.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_int]
        R.ref_sample.restype = C.c_int
        R.ref_sample.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_float, C.c_float,
                                 C.c_float, C.c_uint32]
        R.ref_dispatch_count.restype = C.c_uint64
        _ref = R
    return _ref


# ---------------------------------------------------------------------------------------------
# Host-side restatements: tokenizer and sampler.  Pure Python, small inputs only.
# Pinned against the reference's own code through tests/golden/{tokenizer,sampler}.json.
# ---------------------------------------------------------------------------------------------
def tokenize(tokens, scores, text: bytes, add_bos: bool):
    """TkLlamaTokenizer (th-llama.cpp:909-1041): cut into UTF-8 characters, repeatedly merge the adjacent pair whose
    concatenation is the best-scoring vocabulary entry (ties: leftmost), byte-fallback (id = byte + 3) for the rest.
    tokens: list of bytes (id -> piece), later duplicates win the lookup like the loader (th-llama-loader.cpp:560)."""
    import heapq
    out = []
    if len(text) == 0:
        return out
    if add_bos:
        out.append(1)
    tid = {}
    for i, t in enumerate(tokens):
        tid[bytes(t)] = i
    lens = [1] * 12 + [2, 2, 3, 4]
    off, piece = 0, []          # piece: [prev, next, off, len]
    while off < len(text):
        n = min(len(text) - off, lens[text[off] >> 4])
        piece.append([len(piece) - 1, 0, off, n])
        off += n
        piece[-1][1] = -1 if off == len(text) else len(piece)
    heap = []

    def propose(l, r):
        if l < 0 or r < 0:
            return
        s_ = text[piece[l][2]:piece[l][2] + piece[l][3] + piece[r][3]]
        i = tid.get(s_)
        if i is None or i >= len(tokens):
            return
        heapq.heappush(heap, (-float(np.float32(scores[i])), l, r, len(s_)))   # highest score first, then smallest left

    for i in range(1, len(piece)):
        propose(i - 1, i)
    while heap:
        _, l, r, size = heapq.heappop(heap)
        if piece[l][3] == 0 or piece[r][3] == 0 or piece[l][3] + piece[r][3] != size:
            continue
        piece[l][3] += piece[r][3]
        piece[r][3] = 0
        piece[l][1] = piece[r][1]
        if piece[r][1] >= 0:
            piece[piece[r][1]][0] = l
        propose(piece[l][0], l)
        propose(l, piece[l][1])
    i = 0
    while i != -1:
        _, nxt, o_, n = piece[i]
        t = tid.get(text[o_:o_ + n])
        if t is None:
            out.extend(b + 3 for b in text[o_:o_ + n])
        else:
            out.append(t)
        i = nxt
    return out


class MT19937:
    """std::mt19937 (32-bit Mersenne twister), as seeded by seed(uint32)."""
    def __init__(self, seed):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def __call__(self):
        if self.idx >= 624:
            mt = self.mt
            for i in range(624):
                y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def sample_top_p_top_k(logits, last_n, top_k, top_p, temp, repeat_penalty, seed):
    """llama_sample_top_p_top_k (th-llama.cpp:814-907) with libstdc++'s std::discrete_distribution<int> over an
    mt19937 seeded with `seed` (generate_canonical<double, 53>: two 32-bit draws; index = lower_bound of the draw in
    the cumulative probabilities).  float32 arithmetic where the reference uses float, double where it uses double."""
    f32 = np.float32
    lg = np.asarray(logits, dtype=np.float32)
    n = lg.size
    if temp <= 0:
        return int(np.argmax(lg))                                   # first maximum = lowest id
    scale = f32(1.0) / f32(temp)
    rp = f32(repeat_penalty)
    seen = set(int(t) for t in last_n)
    cand = []
    for i in range(n):
        v = f32(lg[i] * scale)
        if i in seen:
            v = f32(v * rp) if lg[i] < 0 else f32(v / rp)
        cand.append((v, i))
    if 0 < top_k < n:
        # std::partial_sort is not stable; the fixtures avoid ties among the top_k + 1 largest scores
        cand = sorted(cand, key=lambda c: -float(c[0]))[:top_k]
    maxl = max(c[0] for c in cand)
    probs, total = [], 0.0
    for v, _ in cand:
        pr = f32(np.exp(np.float64(f32(v - maxl))))                   # expf: correctly rounded float exp (glibc)
        probs.append(pr)
        total += float(pr)
    probs = [f32(float(pr) / total) for pr in probs]
    if top_p < 1.0:
        cum = 0.0
        for i, pr in enumerate(probs):
            cum += float(pr)
            if cum >= float(f32(top_p)):
                probs = probs[:i + 1]
                cand = cand[:i + 1]
                break
        inv = 1.0 / cum
        probs = [f32(float(pr) * inv) for pr in probs]
    # std::discrete_distribution: normalise in double, cumulative sums, last = 1.0 (libstdc++ random.tcc)
    w = [float(pr) for pr in probs]
    if len(w) < 2:
        return cand[0][1]
    sw = sum(w)
    w = [x / sw for x in w]
    cp, acc = [], 0.0
    for x in w:
        acc += x
        cp.append(acc)
    cp[-1] = 1.0
    rng = MT19937(seed)
    a, b = rng(), rng()
    u = (a + b * 4294967296.0) / 18446744073709551616.0
    if u >= 1.0:
        u = float(np.nextafter(1.0, 0.0))
    import bisect
    return cand[bisect.bisect_left(cp, u)][1]
