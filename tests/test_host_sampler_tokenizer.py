"""Sampler, tokenizer: host C++ (th::llama_sample_logits, th::tk_llama_tokenize) and the oracle's Python
restatements against fixtures recorded from the REFERENCE's own code (tests/golden/make_golden.py drives
th-llama.cpp:814-1041 through oracle/_ref), plus live comparisons with the reference where oracle/_ref is present.
CPU only: the host library needs no device for these entry points."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import token_hawk_b200 as th
from oracle import oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
TOK = json.load(open(os.path.join(HERE, "golden", "tokenizer.json")))
SAMP = json.load(open(os.path.join(HERE, "golden", "sampler.json")))
TOKENS = [bytes.fromhex(h) for h in TOK["vocab_hex"]]
SCORES = np.array(TOK["scores"], dtype=np.float32)


def test_fixture_shape():
    assert len(TOKENS) == len(SCORES) > 300 and len(TOK["cases"]) >= 40 and len(SAMP["cases"]) >= 40
    assert any(c["temp"] == 0.0 for c in SAMP["cases"]) and any(c["top_p"] < 1.0 for c in SAMP["cases"])
    # the duplicate piece and the equal-score pair the generator plants are really there
    assert TOKENS.count(b" the") == 2
    assert SCORES[TOKENS.index(b"ab")] == SCORES[TOKENS.index(b"ba")]


def test_oracle_tokenizer_matches_reference_fixtures():
    for c in TOK["cases"]:
        assert o.tokenize(TOKENS, SCORES, bytes.fromhex(c["text_hex"]), bool(c["bos"])) == c["ids"], c["text_hex"]


def test_host_tokenizer_matches_reference_fixtures_and_oracle():
    v = th.Vocab(TOKENS, SCORES)
    for c in TOK["cases"]:
        text = bytes.fromhex(c["text_hex"])
        got = v.tokenize(text, bool(c["bos"]))
        assert got == c["ids"], text
        assert got == o.tokenize(TOKENS, SCORES, text, bool(c["bos"]))
    assert v.tokenize(b"", True) == []                       # empty text: no BOS either (th-llama.cpp:1046-1048)


def test_tokenizer_byte_fallback_and_roundtrip():
    v = th.Vocab(TOKENS, SCORES)
    ids = v.tokenize("unseen: ß".encode(), False)
    assert all(0 <= i < len(TOKENS) + 256 for i in ids)
    # bytes that are not in the vocabulary come out as byte + 3
    assert [b + 3 for b in "ß".encode()] == ids[-2:]
    # pieces concatenate back to the text when every piece is a vocabulary entry
    text = b" the quick brown fox"
    assert b"".join(TOKENS[i] for i in v.tokenize(text, False)) == text


def _case_arrays(c):
    return np.frombuffer(bytes.fromhex(c["logits_hex"]), dtype=np.float32), np.array(c["last_n"], dtype=np.int32)


def test_oracle_sampler_matches_reference_fixtures():
    for c in SAMP["cases"]:
        lg, last = _case_arrays(c)
        assert o.sample_top_p_top_k(lg, last, c["top_k"], c["top_p"], c["temp"], c["repeat_penalty"], c["seed"]) == c["token"]


def test_host_sampler_matches_reference_fixtures():
    for c in SAMP["cases"]:
        lg, last = _case_arrays(c)
        assert th.sample(lg, last, c["top_k"], c["top_p"], c["temp"], c["repeat_penalty"], c["seed"]) == c["token"]


def test_sampler_properties():
    rng = np.random.default_rng(5)
    lg = rng.standard_normal(500).astype(np.float32)
    # temp <= 0 is the greedy branch: lowest index wins ties
    lg2 = lg.copy(); lg2[[7, 300]] = lg.max() + 1
    assert th.sample(lg2, temp=0.0) == 7
    # top_k = 1 always returns the arg max whatever the seed
    assert {th.sample(lg, top_k=1, temp=0.8, seed=s) for s in range(8)} == {int(np.argmax(lg))}
    # a tiny nucleus keeps only the most likely token -- when top-k sorted the candidates first; with top_k = 0 the
    # reference cuts the nucleus in id order (th-llama.cpp:881-890), which the product reproduces
    assert {th.sample(lg * 20, top_k=40, top_p=0.01, temp=1.0, seed=s) for s in range(8)} == {int(np.argmax(lg))}
    assert {th.sample(lg * 20, top_k=0, top_p=0.01, temp=1.0, seed=s) for s in range(8)} == \
           {o.sample_top_p_top_k(lg * 20, [], 0, 0.01, 1.0, 1.1, s) for s in range(8)}
    # the draw is a function of the seed only
    assert th.sample(lg, seed=123) == th.sample(lg, seed=123)
    # the repetition penalty pushes a dominant token out when it was seen before
    hot = np.zeros(50, np.float32); hot[3] = 2.0; hot[9] = 1.9
    assert th.sample(hot, last_n=[3], top_k=1, temp=1.0, repeat_penalty=1.5) == 9


@pytest.mark.skipif(o.ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_live_against_reference_random():
    R = o.ref_lib()
    rng = np.random.default_rng(11)
    arr = (C.c_char_p * len(TOKENS))(*TOKENS)
    lens = (C.c_int32 * len(TOKENS))(*[len(t) for t in TOKENS])
    sc = (C.c_float * len(TOKENS))(*[float(x) for x in SCORES])
    h = R.ref_vocab_model(arr, lens, sc, len(TOKENS))
    v = th.Vocab(TOKENS, SCORES)
    alphabet = list(" thequickbrownfxjmpsvlazydg.,abc123é日本\U0001f600Z~")
    buf = (C.c_int32 * 4096)()
    for _ in range(200):
        text = "".join(rng.choice(alphabet, size=int(rng.integers(1, 60)))).encode("utf-8")
        n = R.ref_tokenize(h, text, 1, buf, 4096)
        assert v.tokenize(text, True) == list(buf[:n]), text
    R.ref_free(h)
    for _ in range(60):
        n = int(rng.integers(30, 500))
        lg = rng.permutation(n).astype(np.float32) * np.float32(0.05) + rng.standard_normal(n).astype(np.float32) * np.float32(1e-3)
        last = rng.integers(0, n, size=int(rng.integers(0, 16))).astype(np.int32)
        a = dict(top_k=int(rng.choice([0, 3, 40])), top_p=float(rng.choice([1.0, 0.9, 0.3])), temp=float(rng.choice([0.7, 1.0, 1.5])),
                 repeat_penalty=float(rng.choice([1.0, 1.1])), seed=int(rng.integers(0, 2 ** 31)))
        ref = R.ref_sample(n, lg.ctypes.data_as(C.POINTER(C.c_float)), last.ctypes.data_as(C.POINTER(C.c_int32)), last.size,
                           a["top_k"], a["top_p"], a["temp"], a["repeat_penalty"], a["seed"])
        assert th.sample(lg, last, **a) == ref
