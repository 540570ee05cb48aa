"""token_hawk_b200 -- B200-native (sm_100a) single-token LLaMA decode engine behind TokenHawk's op API.

This Python module is only a ctypes harness over the two native libraries (tests, bench.py):

  lib/libthk_sm100a.so   hand-written CUDA kernels + C ABI   (include/thk_cabi.h)
  lib/libth_b200.so      host C++: th:: op surface, LLaMA graph, ggjt loader, capi_* exports

There is no CPU fallback anywhere: loading fails loudly if the libraries are missing, and every
device entry point fails if no CUDA device is present.  The CPU oracle lives in oracle/ and is never
imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_PKG, os.environ.get("THK_LIBDIR", "lib"))   # lib8/ etc. hold kernel variants for A/B runs
KERNEL_LIB = os.path.join(_LIBDIR, "libthk_sm100a.so")
HOST_LIB = os.path.join(_LIBDIR, "libth_b200.so")

THK_OK, THK_E_INVALID, THK_E_CUDA, THK_E_UNSUPPORTED, THK_E_NCCL, THK_E_TIMEOUT = 0, -1, -2, -3, -4, -5
EVAL_FUSED, EVAL_OPGRAPH = 0, 1


class ThkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"thk error {code}: {msg}")
        self.code = code


_k = None
_h = None

i64, u64, vp, f32p, i32p, u16p = C.c_int64, C.c_uint64, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint16)


def kernels() -> C.CDLL:
    """libthk_sm100a.so with argtypes set.  Raises if the library has not been built."""
    global _k
    if _k is None:
        if not os.path.exists(KERNEL_LIB):
            raise ImportError(f"{KERNEL_LIB} missing: run `python -m token_hawk_b200.build` (no CPU fallback exists)")
        K = C.CDLL(KERNEL_LIB, mode=C.RTLD_GLOBAL)
        K.thk_last_error.restype = C.c_char_p
        K.thk_version.restype = C.c_char_p
        K.thk_stream.restype = vp
        sig = {
            "thk_init": [C.c_int, C.POINTER(vp)],
            "thk_init_on_stream": [C.c_int, vp, C.POINTER(vp)],
            "thk_destroy": [vp],
            "thk_device_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)],
            "thk_stream": [vp],
            "thk_malloc": [vp, C.c_size_t, C.POINTER(vp)],
            "thk_free": [vp, vp],
            "thk_memset": [vp, vp, C.c_int, C.c_size_t],
            "thk_upload": [vp, vp, C.c_size_t, vp, C.c_size_t],
            "thk_download": [vp, vp, vp, C.c_size_t, C.c_size_t],
            "thk_copy": [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t],
            "thk_sync": [vp],
            "thk_host_alloc": [vp, C.c_size_t, C.POINTER(vp)],
            "thk_host_free": [vp, vp],
            "thk_vector_mat_mul_trans": [vp, vp, C.c_size_t, vp, vp, i64, i64, i64, C.c_int],
            "thk_vector_multi_mat_mul_split_trans": [vp, vp, C.c_size_t, C.POINTER(vp), C.c_int, vp, vp, i64, i64, C.c_int],
            "thk_vector_reduce": [vp, vp, vp, i64],
            "thk_rms_norm": [vp, vp, i64, i64],
            "thk_row_element_multiply": [vp, vp, vp, i64, i64],
            "thk_rope": [vp, vp, i64, i64, i64, vp],
            "thk_transpose": [vp, vp, vp, i64, i64, i64, C.c_int, vp],
            "thk_mat_mul": [vp, vp, vp, vp, i64, i64, i64, i64, C.c_int, C.c_int, vp],
            "thk_row_softmax": [vp, vp, i64, i64, i64, vp],
            "thk_masked_softmax": [vp, vp, i64, i64, i64, vp],
            "thk_addition": [vp, vp, vp, vp, i64],
            "thk_silu": [vp, vp, i64],
            "thk_element_mult_in_place": [vp, vp, vp, i64],
            "thk_f16_f32_conversion": [vp, vp, C.c_size_t, vp, C.c_size_t, i64],
            "thk_kv_to_hpd": [vp, vp, vp, i64, i64, i64, i64, i64],
            "thk_kv_to_hpd_f16": [vp, vp, vp, i64, i64, i64, i64, i64],
            "thk_kv_from_hpd_f16": [vp, vp, vp, i64, i64, i64, i64, i64],
            "thk_f32_to_f16": [vp, vp, vp, i64],
            "thk_fill_f16": [vp, vp, u64, u64, i64, i64, i64, i64, i64],
            "thk_fill_gain": [vp, vp, u64, u64, i64],
            "thk_fill_kv": [vp, vp, u64, u64, i64, i64, i64, i64, i64, i64],
            "thk_decoder_create": [vp, vp, vp, vp, vp, vp, C.POINTER(vp)],
            "thk_decoder_destroy": [vp],
            "thk_decoder_step": [vp, vp, C.c_int32, vp, vp, vp],
            "thk_decoder_generate": [vp, vp, C.c_int32, C.c_int32, vp, vp],
            "thk_decoder_hidden": [vp, C.POINTER(vp)],
            "thk_decoder_last_launches": [vp],
            "thk_decoder_check": [vp],
            "thk_decoder_profile": [vp, C.c_int, C.POINTER(C.c_uint64), C.c_int],
            "thk_decoder_exchange_info": [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(C.c_size_t)],
            "thk_decoder_set_peers": [vp, C.POINTER(vp), C.POINTER(vp), C.c_int],
            "thk_gemm_f16_tc": [vp, vp, vp, vp, i64, i64, i64],
            "thk_gemm_check": [vp],
            "thk_gemm_reserve": [vp, i64, i64],
            "thk_ipc_export": [vp, vp, C.POINTER(C.c_ubyte)],
            "thk_ipc_import": [vp, C.POINTER(C.c_ubyte), C.POINTER(vp)],
            "thk_ipc_close": [vp, vp],
            "thk_enable_peer_access": [vp, vp],
        }
        for name, args in sig.items():
            fn = getattr(K, name)
            fn.argtypes = args
            if name != "thk_stream":
                fn.restype = C.c_int
        _k = K
    return _k


def trace_commands(fn):
    """Test hook: run fn() and return the labels of the commands the host library issued through the reference's op
    surface (cmdbuf_* names, "copy" for build_layer_cmdbuf's buffer-to-buffer copies)."""
    H = host()
    H.capi_trace_begin()
    buf = C.create_string_buffer(1 << 20)
    try:
        fn()
    finally:
        n = H.capi_trace_end(buf, len(buf))
    labels = buf.value.decode().splitlines()
    assert n == len(labels), (n, len(labels))
    return labels


def host() -> C.CDLL:
    """libth_b200.so (host C++ layer)."""
    global _h
    if _h is None:
        kernels()
        if not os.path.exists(HOST_LIB):
            raise ImportError(f"{HOST_LIB} missing: run `python -m token_hawk_b200.build`")
        H = C.CDLL(HOST_LIB)
        H.capi_last_error.restype = C.c_char_p
        H.capi_device_create.restype = vp
        H.capi_device_create.argtypes = [C.c_int]
        H.capi_device_destroy.argtypes = [vp]
        H.capi_model_synthetic.restype = vp
        H.capi_model_synthetic.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u64]
        H.capi_model_synthetic_tp.restype = vp
        H.capi_model_synthetic_tp.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u64, C.c_int, C.c_int]
        H.capi_model_load_tp.restype = vp
        H.capi_model_load_tp.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int]
        H.capi_model_tp.argtypes = [vp, i32p]
        H.capi_exchange_info.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        H.capi_set_peers.argtypes = [vp, C.POINTER(vp), C.c_int]
        H.capi_eval_launch.argtypes = [vp, i32p, C.c_int, C.c_int]
        H.capi_eval_finish.argtypes = [vp, f32p]
        H.capi_model_load.restype = vp
        H.capi_model_load.argtypes = [vp, C.c_char_p, C.c_int]
        H.capi_model_free.argtypes = [vp]
        H.capi_model_dims.argtypes = [vp, i32p]
        H.capi_set_eval_path.argtypes = [vp, C.c_int]
        H.capi_reset.argtypes = [vp]
        H.capi_set_batch_prefill.argtypes = [vp, C.c_int]
        H.capi_eval.argtypes = [vp, i32p, C.c_int, C.c_int, f32p]
        H.capi_last_launches.restype = i64
        H.capi_last_launches.argtypes = [vp]
        H.capi_generate.argtypes = [vp, i32p, C.c_int, C.c_int, i32p]
        H.capi_generate_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32p, f32p]
        H.capi_step_async.argtypes = [vp, C.c_int]
        H.capi_set_token.argtypes = [vp, C.c_int]
        H.capi_sync.argtypes = [vp]
        H.capi_check.argtypes = [vp]
        H.capi_stream.restype = vp
        H.capi_stream.argtypes = [vp]
        H.capi_fill_kv.argtypes = [vp, u64, C.c_int]
        H.capi_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.c_int]
        H.capi_tune.argtypes = [vp, C.c_char_p, C.c_int]
        H.capi_hidden.argtypes = [vp, f32p]
        H.capi_set_default_kv_f16.restype = None
        H.capi_set_default_kv_f16.argtypes = [C.c_int]
        H.capi_trace_begin.restype = None
        H.capi_trace_begin.argtypes = []
        H.capi_trace_end.argtypes = [C.c_char_p, C.c_int]
        H.capi_tensor_info.restype = i64
        H.capi_tensor_info.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
        H.capi_tensor_download.argtypes = [vp, C.c_char_p, vp, i64]
        H.capi_vocab_size.argtypes = [vp]
        H.capi_sample.argtypes = [C.c_int, f32p, i32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint32]
        H.capi_vocab_create.restype = vp
        H.capi_vocab_create.argtypes = [C.POINTER(C.c_char_p), i32p, f32p, C.c_int]
        H.capi_vocab_free.argtypes = [vp]
        H.capi_vocab_tokenize.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, i32p, C.c_int]
        H.capi_tokenize.argtypes = [vp, C.c_char_p, C.c_int, i32p, C.c_int]
        H.capi_token_str.restype = C.c_char_p
        H.capi_token_str.argtypes = [vp, C.c_int]
        H.capi_set_sampler.argtypes = [vp, C.c_float, C.c_uint32]
        H.capi_inference.argtypes = [vp, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        _h = H
    return _h


def _check(rc):
    if rc != 0:
        raise ThkError(rc, kernels().thk_last_error().decode())


def sample(logits: np.ndarray, last_n=(), top_k=40, top_p=0.95, temp=0.8, repeat_penalty=1.1, seed=0) -> int:
    """th::llama_sample_logits (the reference's llama_sample_top_p_top_k, th-llama.cpp:814-907) with an mt19937
    seeded with `seed`.  Host-only: needs no GPU."""
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    last = np.ascontiguousarray(last_n, dtype=np.int32)
    return int(host().capi_sample(lg.size, lg.ctypes.data_as(f32p), last.ctypes.data_as(i32p) if last.size else None, int(last.size),
                                  int(top_k), float(top_p), float(temp), float(repeat_penalty), int(seed)))


class Vocab:
    """A bare th::LlamaVocab for the tokenizer (th-llama.cpp:909-1041).  Host-only: needs no GPU."""

    def __init__(self, tokens, scores):
        toks = [bytes(t) for t in tokens]
        arr = (C.c_char_p * len(toks))(*toks)
        lens = np.array([len(t) for t in toks], dtype=np.int32)
        sc = np.ascontiguousarray(scores, dtype=np.float32)
        self.h = host().capi_vocab_create(arr, lens.ctypes.data_as(i32p), sc.ctypes.data_as(f32p), len(toks))

    def tokenize(self, text: bytes, add_bos: bool = True):
        cap = 4 * len(text) + 8
        out = np.zeros(cap, dtype=np.int32)
        n = host().capi_vocab_tokenize(self.h, text, len(text), int(add_bos), out.ctypes.data_as(i32p), cap)
        return [int(x) for x in out[:n]]

    def __del__(self):
        try:
            if self.h:
                host().capi_vocab_free(self.h)
                self.h = None
        except Exception:
            pass


class Device:
    """thk_ctx: one CUDA device + stream (the WGPUDevice/WGPUQueue pair of the reference)."""

    def __init__(self, ordinal: int = 0, stream: int | None = None):
        K = kernels()
        h = vp()
        if stream is None:
            _check(K.thk_init(ordinal, C.byref(h)))
        else:
            _check(K.thk_init_on_stream(ordinal, vp(stream), C.byref(h)))
        self.h = h
        sm, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        _check(K.thk_device_info(h, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
        self.sm_count, self.cc, self.total_mem = sm.value, (maj.value, mnr.value), mem.value

    def close(self):
        if self.h:
            kernels().thk_destroy(self.h)
            self.h = None

    def sync(self):
        _check(kernels().thk_sync(self.h))

    # ---- buffers ----
    def empty(self, shape, dtype=np.float32) -> "DeviceArray":
        return DeviceArray(self, shape, dtype)

    def array(self, a: np.ndarray) -> "DeviceArray":
        a = np.ascontiguousarray(a)
        d = DeviceArray(self, a.shape, a.dtype)
        d.upload(a)
        return d


class DeviceArray:
    """A device allocation with a numpy shape/dtype, for the op-level tests."""

    def __init__(self, dev: Device, shape, dtype):
        self.dev, self.shape, self.dtype = dev, tuple(int(s) for s in np.atleast_1d(shape)), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = vp()
        _check(kernels().thk_malloc(dev.h, self.nbytes, C.byref(p)))
        self.ptr = p

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes
        _check(kernels().thk_upload(self.dev.h, self.ptr, 0, a.ctypes.data_as(vp), self.nbytes))
        self.dev.sync()

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        _check(kernels().thk_download(self.dev.h, out.ctypes.data_as(vp), self.ptr, 0, self.nbytes))
        return out

    def __del__(self):
        try:
            if self.ptr and self.dev.h:
                kernels().thk_free(self.dev.h, self.ptr)
        except Exception:
            pass


class NetworkUniforms(C.Structure):   # th-llama.hpp:181-192
    _fields_ = [("n_past", C.c_uint32), ("n_tokens", C.c_uint32), ("pad2", C.c_float), ("pad3", C.c_float), ("pad4", C.c_uint32 * 4)]


class DimsUniforms(C.Structure):      # th-llama.hpp:195-204
    _fields_ = [("A_B", C.c_uint32), ("A_M", C.c_uint32), ("A_N", C.c_uint32), ("scale", C.c_float),
                ("B_B", C.c_uint32), ("B_M", C.c_uint32), ("B_N", C.c_uint32), ("offset", C.c_float)]


def device_struct(dev: Device, s: C.Structure) -> DeviceArray:
    raw = np.frombuffer(bytes(s), dtype=np.uint8)
    return dev.array(raw)


class LlamaModel:
    """th::LlamaModel through the capi_* exports."""

    def __init__(self, dev: Device, handle):
        if not handle:
            raise ThkError(-1, host().capi_last_error().decode())
        self.dev, self.h = dev, handle
        d = (C.c_int32 * 9)()
        host().capi_model_dims(handle, d)
        (self.n_vocab, self.n_embd, self.n_mult, self.n_head, self.n_layer, self.n_ctx, self.n_ff,
         self.has_fused, self.eval_path) = list(d)
        t = (C.c_int32 * 2)()
        host().capi_model_tp(handle, t)
        self.tp_rank, self.tp_size = t[0], t[1]
        self.n_vocab_local = self.n_vocab // self.tp_size

    @classmethod
    def synthetic(cls, dev: Device, n_vocab=32000, n_embd=4096, n_mult=256, n_head=32, n_layer=32, n_ctx=512, seed=0x7B5EED,
                  tp_rank=0, tp_size=1, kv_f16=False):
        """kv_f16: the fused path keeps its KV cache in f16 (K after RoPE and V rounded to nearest-even at append): half the
        KV bytes per token; parity is then against the oracle's kv_f16 mode (tests/test_gpu_kv_f16.py)."""
        host().capi_set_default_kv_f16(1 if kv_f16 else 0)
        try:
            return cls(dev, host().capi_model_synthetic_tp(dev.h, n_vocab, n_embd, n_mult, n_head, n_layer, n_ctx, seed, tp_rank, tp_size))
        finally:
            host().capi_set_default_kv_f16(0)

    @classmethod
    def load(cls, dev: Device, path: str, n_ctx: int = 512, tp_rank=0, tp_size=1, kv_f16=False):
        host().capi_set_default_kv_f16(1 if kv_f16 else 0)
        try:
            return cls(dev, host().capi_model_load_tp(dev.h, path.encode(), n_ctx, tp_rank, tp_size))
        finally:
            host().capi_set_default_kv_f16(0)

    # ---- tensor-parallel wiring (token_hawk_b200.tp) ----
    def exchange_info(self):
        p, n = vp(), u64()
        if host().capi_exchange_info(self.h, C.byref(p), C.byref(n)):
            raise ThkError(-1, host().capi_last_error().decode())
        return p.value, n.value

    def set_peers(self, ptrs):
        arr = (vp * len(ptrs))(*[vp(x) for x in ptrs])
        if host().capi_set_peers(self.h, arr, len(ptrs)):
            raise ThkError(-1, host().capi_last_error().decode())

    def eval_launch(self, tokens, n_past: int):
        toks = np.ascontiguousarray(np.asarray(tokens, np.int32).reshape(-1))
        if host().capi_eval_launch(self.h, toks.ctypes.data_as(i32p), len(toks), n_past):
            raise ThkError(-1, host().capi_last_error().decode())

    def eval_finish(self):
        logits = np.empty(self.n_vocab_local, np.float32)
        tok = host().capi_eval_finish(self.h, logits.ctypes.data_as(f32p))
        if tok < 0:
            raise ThkError(tok, host().capi_last_error().decode())
        return tok, logits

    def close(self):
        if self.h:
            host().capi_model_free(self.h)
            self.h = None

    def set_eval_path(self, path: int):
        if host().capi_set_eval_path(self.h, path):
            raise ThkError(-3, host().capi_last_error().decode())
        self.eval_path = path

    def eval(self, tokens, n_past: int, want_logits: bool = True):
        """th_eval_gpu: host token ids in, (greedy token, host logits) out."""
        toks = np.ascontiguousarray(np.asarray(tokens, np.int32).reshape(-1))
        logits = np.empty(self.n_vocab_local, np.float32) if want_logits else None
        tok = host().capi_eval(self.h, toks.ctypes.data_as(i32p), len(toks), n_past,
                               logits.ctypes.data_as(f32p) if want_logits else None)
        if tok < 0:
            raise ThkError(tok, host().capi_last_error().decode())
        return tok, logits

    @property
    def last_launches(self) -> int:
        return int(host().capi_last_launches(self.h))

    def generate(self, prompt, n_new: int):
        p = np.ascontiguousarray(np.asarray(prompt, np.int32))
        out = np.empty(n_new, np.int32)
        n = host().capi_generate(self.h, p.ctypes.data_as(i32p), len(p), n_new, out.ctypes.data_as(i32p))
        return out[:n].tolist()

    def reset(self):
        host().capi_reset(self.h)

    def set_batch_prefill(self, on: bool):
        """n_tokens > 1: one batched pass on the tensor cores (default) or n single-token steps."""
        host().capi_set_batch_prefill(self.h, int(on))

    def generate_device(self, first_token: int, n_past: int, n_steps: int, want_logits=False):
        out = np.empty(n_steps, np.int32)
        logits = np.empty(self.n_vocab_local, np.float32) if want_logits else None
        rc = host().capi_generate_device(self.h, first_token, n_past, n_steps, out.ctypes.data_as(i32p),
                                         logits.ctypes.data_as(f32p) if want_logits else None)
        if rc:
            raise ThkError(rc, host().capi_last_error().decode())
        return (out.tolist(), logits) if want_logits else out.tolist()

    def fill_kv(self, n_positions: int, seed: int = 99):
        if host().capi_fill_kv(self.h, seed, n_positions):
            raise ThkError(-1, host().capi_last_error().decode())

    def hidden(self) -> np.ndarray:
        out = np.empty(self.n_embd, np.float32)
        if host().capi_hidden(self.h, out.ctypes.data_as(f32p)):
            raise ThkError(-1, host().capi_last_error().decode())
        return out

    def tensor(self, name: str) -> np.ndarray:
        info = (i64 * 3)()
        nbytes = host().capi_tensor_info(self.h, name.encode(), info)
        if nbytes < 0:
            raise KeyError(name)
        dt = np.float16 if info[0] else np.float32
        out = np.empty((info[1], info[2]), dt)
        if host().capi_tensor_download(self.h, name.encode(), out.ctypes.data_as(vp), nbytes):
            raise ThkError(-1, host().capi_last_error().decode())
        return out

    def profile(self, enable: bool, fetch: bool = False):
        """In-kernel timeline of the last fused launch.  Returns (marks[cta][phase<256][8] ns, producer[cta][4])
        when fetch is set; slots: 0 phase start, 1 prologue end, 2 first tile ready, 3 last tile done,
        4 barrier arrive, 5 after fence, 6 consumer ring-wait cycles."""
        grid = self.dev.sm_count
        n0, n1, nt, nw = grid * 256 * 8, grid * 4, grid * 64, grid * 12 * 4
        n = n0 + n1 + 2 * nt + nw if fetch else 0
        buf = np.zeros(max(n, 1), dtype=np.uint64)
        if host().capi_profile(self.h, int(enable), buf.ctypes.data_as(C.POINTER(C.c_uint64)) if fetch else None, n):
            raise ThkError(-1, host().capi_last_error().decode())
        if not fetch:
            return None
        # per-tile times of the phase selected with tune("prof_phase", k): [cta][64] retired, [cta][64] issued
        self.last_tile_times = (buf[n0 + n1:n0 + n1 + nt].reshape(grid, 64).astype(np.int64),
                                buf[n0 + n1 + nt:n0 + n1 + 2 * nt].reshape(grid, 64).astype(np.int64))
        # cycles every warp spent waiting: [cta][warp: 16 math, producer, epilogue, reducer][ring, hand-off, poll, lifetime]
        self.last_wait_cycles = buf[n0 + n1 + 2 * nt:].reshape(grid, 12, 4).astype(np.int64)
        return buf[:n0].reshape(grid, 256, 8).astype(np.int64), buf[n0:n0 + n1].reshape(grid, 4).astype(np.int64)

    def tokenize(self, text: str, add_bos: bool = True):
        cap = 4 * len(text.encode()) + 8
        out = np.zeros(cap, dtype=np.int32)
        n = host().capi_tokenize(self.h, text.encode(), int(add_bos), out.ctypes.data_as(i32p), cap)
        return [int(x) for x in out[:n]]

    def token_str(self, token: int) -> bytes:
        return host().capi_token_str(self.h, int(token)) or b""

    def set_sampler(self, temp: float, seed: int = 0):
        host().capi_set_sampler(self.h, float(temp), int(seed))

    def inference(self, prompt: str, max_new_tokens: int = 32) -> bytes:
        """th::do_inference (th-llama.cpp:111-168): tokenise, feed the prompt, generate; returns the generated bytes."""
        cap = 1 << 16
        buf = C.create_string_buffer(cap)
        n = host().capi_inference(self.h, prompt.encode(), int(max_new_tokens), buf, cap)
        if n < 0:
            raise ThkError(-1, host().capi_last_error().decode())
        return buf.raw[:min(n, cap - 1)]

    def tune(self, key: str, value: int):
        """Persistent-kernel tuning knob (thk_decoder_tune): 'l2_ahead_kb', 'poll_depth'."""
        if host().capi_tune(self.h, key.encode(), int(value)):
            raise ThkError(-1, host().capi_last_error().decode())

    # enqueue-only calls for stream timing
    def set_token(self, tok: int):
        _check(host().capi_set_token(self.h, tok))

    def step_async(self, n_past: int):
        rc = host().capi_step_async(self.h, n_past)
        if rc:
            raise ThkError(rc, host().capi_last_error().decode())

    def check(self):
        rc = host().capi_check(self.h)
        if rc:
            raise ThkError(rc, host().capi_last_error().decode())
