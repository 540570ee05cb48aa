// th_host_test.cpp -- the device-free part of the th:: host surface from C++, the way a C++ front end (cli/main.cpp in
// the reference) would use it: vocabulary + tokenizer (th-llama.cpp:909-1107) and the sampler (th-llama.cpp:814-907).
// Runs on the CPU; the expectations are properties plus two hand-checked cases (the reference-recorded fixtures are
// exercised from Python, tests/test_host_sampler_tokenizer.py).
#include <stdio.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include "th/th-llama.hpp"

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

static void add(th::LlamaVocab& v, const std::string& tok, float score) {
    v.token_to_id[tok] = (int32_t)v.id_to_token.size();
    v.id_to_token.push_back({tok, score});
}

int main() {
    th::LlamaVocab v;
    add(v, "<unk>", 0.f); add(v, "<s>", 0.f); add(v, "</s>", 0.f);
    for (int b = 0; b < 256; ++b) { char buf[8]; snprintf(buf, sizeof buf, "<0x%02X>", b); add(v, buf, 0.f); }
    const int first = (int)v.id_to_token.size();
    add(v, " ", -10.f); add(v, "h", -11.f); add(v, "e", -12.f); add(v, "l", -13.f); add(v, "o", -14.f);   // first .. first+4
    add(v, "he", -1.f); add(v, "ll", -2.f); add(v, "hell", -0.5f); add(v, "hello", -0.25f); add(v, " hello", -0.1f);
    add(v, "lo", -3.f);

    // " hello": he, ll merge first (best scores among the initial pairs), then hell, hello, " hello"
    std::vector<th::tk_llama_token> t = th::tk_llama_tokenize(v, " hello", true);
    CHECK(t.size() == 2 && t[0] == th::tk_llama_token_bos() && t[1] == v.token_to_id[" hello"]);
    // "hellol": ... "hello" then the trailing "l" stays a single piece ("lo" lost its left half to "hello")
    t = th::tk_llama_tokenize(v, "hellol", false);
    CHECK(t.size() == 2 && t[0] == v.token_to_id["hello"] && t[1] == first + 3);
    // characters outside the vocabulary come out as byte tokens (byte + 3), multi-byte UTF-8 as several of them
    t = th::tk_llama_tokenize(v, "h\xC3\xA9", false);
    CHECK(t.size() == 3 && t[0] == first + 1 && t[1] == 0xC3 + 3 && t[2] == 0xA9 + 3);
    CHECK(th::tk_llama_tokenize(v, "", true).empty());
    CHECK(th::tk_llama_token_eos() == 2);

    // sampler: greedy branch, top-k = 1, determinism per seed, repetition penalty
    std::vector<float> logits(64);
    for (int i = 0; i < 64; ++i) logits[i] = 0.01f * (float)((i * 37) % 64);
    int best = 0;
    for (int i = 1; i < 64; ++i) if (logits[i] > logits[best]) best = i;
    std::mt19937 rng(5);
    CHECK(th::llama_sample_logits(rng, logits.data(), 64, {}, 40, 0.95f, 0.0f, 1.1f) == best);
    for (unsigned s = 0; s < 8; ++s) {
        std::mt19937 r1(s), r2(s);
        CHECK(th::llama_sample_logits(r1, logits.data(), 64, {}, 1, 1.0f, 0.8f, 1.0f) == best);
        std::mt19937 r3(s);
        const int a = th::llama_sample_logits(r2, logits.data(), 64, {}, 40, 0.95f, 0.8f, 1.1f);
        const int b = th::llama_sample_logits(r3, logits.data(), 64, {}, 40, 0.95f, 0.8f, 1.1f);
        CHECK(a == b && a >= 0 && a < 64);
    }
    std::vector<float> hot(16, 0.f);
    hot[3] = 2.0f; hot[9] = 1.9f;
    std::mt19937 r4(1);
    CHECK(th::llama_sample_logits(r4, hot.data(), 16, {3}, 1, 1.0f, 1.0f, 1.5f) == 9);
    printf("PASS th_host_test\n");
    return 0;
}
