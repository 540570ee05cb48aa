"""Generates tests/golden/*.npz|json from the REAL reference host code.

Run in the build container (needs /root/reference): `python tests/golden/make_golden.py`.
It builds oracle/_ref/libthref_host.so (reference th.cpp / th-llama.cpp / th-llama-loader.cpp
compiled unmodified against oracle/webgpu_stub) and records what the reference itself computes:

  fp16_table.npz   ggml_compute_fp16_to_fp32 for all 65536 codes (th.cpp:312) and
                   ggml_compute_fp32_to_fp16 on 20000 seeded floats (th.cpp:335)
  greedy.json      llama_sample_top_p_top_k(temp=0) on seeded logit vectors incl. ties
                   (th-llama.cpp:826-838)
  loader_tiny.json load_llama_file on a tiny ggjt file written by the oracle's writer
                   (th-llama-loader.cpp:485): hparams + per tensor dtype / shape / sha256

The GPU box has no /root/reference; tests there read only these committed files.
"""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TINY_SEED = 0x7B5EED


def greedy_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for n in (2, 7, 512, 4000, 32000):
        for k in range(4):
            v = rng.standard_normal(n).astype(np.float32)
            if k == 1:   # exact tie at the maximum: lowest index must win
                i, j = sorted(rng.choice(n, 2, replace=False).tolist())
                v[i] = v[j] = np.float32(v.max() + 1.0)
            if k == 2:   # all equal
                v[:] = np.float32(0.25)
            if k == 3:   # maximum at the last position, with -inf entries
                v[rng.integers(0, n - 1)] = -np.inf
                v[n - 1] = np.float32(v[np.isfinite(v)].max() + 0.5)
            cases.append(v)
    return cases


CORPUS = ("the quick brown fox jumps over the lazy dog . the theory of the thing is that there is nothing then . "
          "hello world , hello there , hello again . tokens and tokenizers tokenize text into tokens . "
          "caf\u00e9 na\u00efve r\u00e9sum\u00e9 \u65e5\u672c\u8a9e \u65e5\u672c \u8a9e\u5b66 \U0001f600 \U0001f600\U0001f600 "
          "aaaa aaa aa a abababab abcabcabc 1234567890 12 123 1234 .").encode("utf-8")


def build_vocab(n_merges=220):
    """Deterministic toy SentencePiece-like vocabulary: 3 specials, 256 byte pieces, every UTF-8 character of CORPUS,
    then `n_merges` BPE merges learnt on CORPUS (score = -rank; characters score -1000 - i).  Two entries are given EQUAL
    scores on purpose (tie -> leftmost merge first) and one piece is listed twice (the later id wins the lookup)."""
    text = CORPUS.decode("utf-8")
    toks = [b"<unk>", b"<s>", b"</s>"] + [("<0x%02X>" % b).encode() for b in range(256)]
    scores = [0.0, 0.0, 0.0] + [0.0] * 256
    chars = sorted(set(text))
    for i, ch in enumerate(chars):
        toks.append(ch.encode("utf-8"))
        scores.append(-1000.0 - i)
    words = [list(w) for w in text.split(" ")]
    words = [[" "] + w for w in words]                 # pieces may start with the separating space
    merges = []
    for rank in range(n_merges):
        cnt = {}
        for w in words:
            for a, b in zip(w, w[1:]):
                cnt[(a, b)] = cnt.get((a, b), 0) + 1
        if not cnt:
            break
        best = sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0]))[0][0]
        merges.append(best)
        for w in words:
            i = 0
            while i + 1 < len(w):
                if (w[i], w[i + 1]) == best:
                    w[i:i + 2] = [w[i] + w[i + 1]]
                else:
                    i += 1
        toks.append((best[0] + best[1]).encode("utf-8"))
        scores.append(-float(rank))
    # equal scores: "ab" and "ba" style neighbours
    for piece in (b"ab", b"ba"):
        if piece in toks:
            scores[toks.index(piece)] = -5.5
        else:
            toks.append(piece)
            scores.append(-5.5)
    toks.append(b" the")                                # duplicate piece: later id wins
    scores.append(-0.25)
    return toks, scores


TOKENIZER_TEXTS = [
    " the quick brown fox", "the theory of nothing", " hello world , hello again .", "tokenizers tokenize tokens",
    " caf\u00e9 na\u00efve r\u00e9sum\u00e9", "\u65e5\u672c\u8a9e\u5b66", " \U0001f600\U0001f600\U0001f600 ok", "abababababa", "bababab", "aaaaaaa",
    "abcabcabcabc", " 12345678901234", "x", " ", "  double  spaces  ", "unseen: \u00df\u2603 ~^", "Zebra QUIZ", "\n\ttabs\n",
    "th", "the", " the", "thethethe", "\u00e9\u00e9\u00e9", "a\u65e5b\u672cc", "end.",
]


def sampler_cases():
    rng = np.random.default_rng(TINY_SEED + 7)
    cases = []
    for i in range(48):
        n = int(rng.integers(40, 600))
        # distinct scores (no ties among the largest: std::partial_sort is not stable)
        lg = (rng.permutation(n).astype(np.float32) * np.float32(0.037) - np.float32(n * 0.0185)).astype(np.float32)
        lg += rng.standard_normal(n).astype(np.float32) * np.float32(0.001)
        last = rng.integers(0, n, size=int(rng.integers(0, 24))).astype(np.int32)
        cases.append({"logits": lg, "last_n": last, "top_k": int(rng.choice([0, 1, 5, 40, n + 5])),
                      "top_p": float(rng.choice([1.0, 0.95, 0.9, 0.5, 0.05])), "temp": float(rng.choice([0.8, 1.0, 0.3, 2.0, 0.0])),
                      "repeat_penalty": float(rng.choice([1.0, 1.1, 1.3])), "seed": int(rng.integers(0, 2 ** 31))})
    return cases


def host_side_fixtures(R):
    """tests/golden/tokenizer.json and sampler.json from the REAL reference tokenizer / sampler (oracle/_ref)."""
    toks, scores = build_vocab()
    arr = (C.c_char_p * len(toks))(*toks)
    lens = (C.c_int32 * len(toks))(*[len(t) for t in toks])
    sc = (C.c_float * len(toks))(*scores)
    h = R.ref_vocab_model(arr, lens, sc, len(toks))
    out = {"vocab_hex": [t.hex() for t in toks], "scores": [float(np.float32(x)) for x in scores], "cases": []}
    buf = (C.c_int32 * 4096)()
    for t in TOKENIZER_TEXTS:
        for bos in (0, 1):
            raw = t.encode("utf-8")
            n = R.ref_tokenize(h, raw, bos, buf, 4096)
            out["cases"].append({"text_hex": raw.hex(), "bos": bos, "ids": list(buf[:n])})
    R.ref_free(h)
    json.dump(out, open(os.path.join(HERE, "tokenizer.json"), "w"), indent=0)
    cases = []
    for c in sampler_cases():
        lg, last = c["logits"], c["last_n"]
        tok = R.ref_sample(lg.size, lg.ctypes.data_as(C.POINTER(C.c_float)), last.ctypes.data_as(C.POINTER(C.c_int32)), last.size,
                           c["top_k"], c["top_p"], c["temp"], c["repeat_penalty"], c["seed"])
        cases.append({"logits_hex": lg.tobytes().hex(), "last_n": [int(x) for x in last], "top_k": c["top_k"], "top_p": c["top_p"],
                      "temp": c["temp"], "repeat_penalty": c["repeat_penalty"], "seed": c["seed"], "token": int(tok)})
    json.dump({"cases": cases}, open(os.path.join(HERE, "sampler.json"), "w"), indent=0)


def main():
    o.build()
    R = o.ref_lib()
    assert R is not None, "oracle/_ref not built (no /root/reference?)"

    table = np.array([R.ref_fp16_to_fp32(i) for i in range(65536)], dtype=np.float32)
    rng = np.random.default_rng(7)
    f = np.concatenate([
        rng.standard_normal(8000).astype(np.float32),
        (rng.standard_normal(4000) * 1e-6).astype(np.float32),
        (rng.standard_normal(4000) * 7e4).astype(np.float32),
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0, 65520.0, 6e-8, 5.96e-8, 2.98e-8],
                 dtype=np.float32),
        (rng.standard_normal(3990) * 0.02).astype(np.float32)])
    h = np.array([R.ref_fp32_to_fp16(float(x)) for x in f], dtype=np.uint16)
    np.savez_compressed(os.path.join(HERE, "fp16_table.npz"), f16_to_f32_bits=table.view(np.uint32),
                        f32_in_bits=f.view(np.uint32), f16_out=h)

    g = []
    for v in greedy_cases():
        vv = np.ascontiguousarray(v)
        g.append(int(R.ref_greedy(vv.ctypes.data_as(C.POINTER(C.c_float)), vv.size)))
    json.dump({"seed": 2024, "expected": g}, open(os.path.join(HERE, "greedy.json"), "w"))

    m = o.Model.synthetic(o.TINY, TINY_SEED)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "tiny.ggjt")
        m.write_ggjt(path)
        file_sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
        h_ = R.ref_load(path.encode())
        assert h_, "reference loader rejected the file"
        hp = (C.c_int32 * 9)()
        R.ref_hparams(h_, hp)
        tensors = {}
        names = o.tensor_names(o.TINY.n_layer) + ["output.weight-split1", "output.weight-split2",
                                                  "layers.0.key_cache"]
        for name in names:
            ty, shp, nb = C.c_int(), (C.c_int64 * 4)(), C.c_int64()
            p = R.ref_tensor(h_, name.encode(), C.byref(ty), shp, C.byref(nb))
            assert p, name
            raw = C.string_at(p, nb.value)
            tensors[name] = {"f16": ty.value, "shape_lbrc": list(shp), "nbytes": nb.value,
                             "sha256": hashlib.sha256(raw).hexdigest()}
        toks = (C.c_int32 * 64)()
        n = R.ref_tokenize(h_, b"<5><17>", 1, toks, 64)
        R.ref_free(h_)
    json.dump({"seed": TINY_SEED, "file_sha256": file_sha, "hparams": list(hp),
               "dispatches_during_load": int(R.ref_dispatch_count()),
               "tokenize_<5><17>": list(toks[:n]), "tensors": tensors},
              open(os.path.join(HERE, "loader_tiny.json"), "w"), indent=1)
    host_side_fixtures(R)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
