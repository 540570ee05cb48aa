// th_llama.cpp -- LLaMA graph on the th:: op surface (mirrors kayvr/token-hawk th-llama.cpp).
//
// Two ways to evaluate a token behind the same th_eval_gpu entry point:
//   * EvalPath_OpGraph: the reference's graph, one op per launch, in the reference's order and with
//     its buffer roles (build_layer_cmdbuf th-llama.cpp:270-452, build_final_compute_cmdbuf :240-268).
//     It exists to show the op surface is a drop-in and as a second, independent GPU path for parity.
//   * EvalPath_Fused: thk_decoder_step -- one persistent sm_100a kernel per token (default).
#include "th/th-llama.hpp"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <queue>
#include <utility>

namespace th {

extern int64_t g_launch_count;   // th.cpp: launches issued through emit()
// rows [pos0, pos0 + npos) between the op graph's f32 [pos][head][dim] cache and the fused kernel's [head][pos][dim] cache
// (f32, or f16 when LlamaModel::kvF16: rounded on the way in, widened on the way out)
static int kv_rows_to_hpd(const LlamaModel& m, const TensorBuffer& phd, const TensorBuffer& hpd, int64_t pos0, int64_t npos) {
    const int64_t H = m.n_head, D = m.n_embd / m.n_head;
    return m.kvF16 ? thk_kv_to_hpd_f16(m.device, (const float*)phd.gpu, (uint16_t*)hpd.gpu, pos0, npos, m.n_ctx, H, D)
                   : thk_kv_to_hpd(m.device, (const float*)phd.gpu, (float*)hpd.gpu, pos0, npos, m.n_ctx, H, D);
}
static int kv_rows_from_hpd(const LlamaModel& m, const TensorBuffer& hpd, const TensorBuffer& phd, int64_t pos0, int64_t npos) {
    const int64_t H = m.n_head, D = m.n_embd / m.n_head;
    return m.kvF16 ? thk_kv_from_hpd_f16(m.device, (const uint16_t*)hpd.gpu, (float*)phd.gpu, pos0, npos, m.n_ctx, H, D)
                   : thk_kv_from_hpd(m.device, (const float*)hpd.gpu, (float*)phd.gpu, pos0, npos, m.n_ctx, H, D);
}
void trace_command(const char* label);   // th.cpp: test hook
// wgpuCommandEncoderCopyBufferToBuffer of build_layer_cmdbuf (th-llama.cpp:337-338, :412, :450)
static inline void copy_command(WGPUDevice device, void* dst, size_t dst_off, const void* src, size_t src_off, size_t bytes) {
    trace_command("copy");
    thk_copy(device, dst, dst_off, src, src_off, bytes);
}

static EncoderTag* const kEncoder = reinterpret_cast<EncoderTag*>(1);   // "a submission is open"

LlamaModel::~LlamaModel() {
    if (decoder) thk_decoder_destroy(decoder);
    if (device) {
        if (d_token) thk_free(device, d_token);
        if (d_next) thk_free(device, d_next);
        if (networkUniforms) thk_free(device, networkUniforms);
        for (auto& u : dimsUniforms) if (u) thk_free(device, u);
        if (pinnedLogits) thk_host_free(device, pinnedLogits);
    }
}

void reset_layer_tensors(LlamaLayer& l) {   // th-llama.cpp:78-93
    l.attention_norm.reset_shape();
    l.wq.reset_shape(); l.wk.reset_shape(); l.wv.reset_shape(); l.wo.reset_shape();
    l.ffn_norm.reset_shape();
    l.w1.reset_shape(); l.w2.reset_shape(); l.w3.reset_shape();
    l.key_cache.reset_shape(); l.value_cache.reset_shape();
}

void reset_working_memory_tensors(LlamaModel& m) {   // th-llama.cpp:95-106
    m.working_key_cache.reset_shape();
    m.working_val_cache.reset_shape();
    for (int i = 0; i < LlamaModel::nInpBuffers; ++i) m.inp[i].reset_shape();
    m.ffWorking[0].reset_shape();
    m.ffWorking[1].reset_shape();
}

// Dry run with a null encoder: every op site only records its shape contract (th-llama.cpp:66-76).
// Only the single-token sets are built; the 8-token sets of the reference feed a dormant path.
void build_pipelines_llama(WGPUDevice device, WGPUQueue, std::shared_ptr<LlamaModel> m) {
    LlamaLayer& l = m->layers[0];
    build_layer_cmdbuf(device, nullptr, m, l, m->ps, 1, 0);
    build_final_compute_cmdbuf(device, nullptr, m, m->pfs, 1);
    reset_working_memory_tensors(*m);
    reset_layer_tensors(l);
}

// th-llama.cpp:240-268.  No output split and no reduce: one [n_vocab, n_embd] matvec.
void build_final_compute_cmdbuf(WGPUDevice device, WGPUCommandEncoder encoder, std::shared_ptr<LlamaModel> m,
                                LlamaFinalComputePipeline& p, int n_tokens) {
    reset_working_memory_tensors(*m);
    m->inp[0].shape = TensorShape{0, 0, n_tokens, m->n_embd};
    cmdbuf_rms_norm(device, encoder, nullptr, &p.p01, m->inp[0]);
    cmdbuf_row_element_multiply(device, encoder, nullptr, &p.p02, m->inp[0], m->norm);
    // last row only (aOffset), th-llama.cpp:265-266
    const int64_t aOffset = ((m->inp[0].shape.r - 1) * m->inp[0].shape.c) * (int64_t)get_TensorType_size(m->inp[0].type);
    TensorBuffer& x = m->inp[0];
    const TensorShape keep = x.shape;
    x.shape = TensorShape{0, 0, 1, m->n_embd};   // the matvec sees one row; the offset selects which
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p03, x, m->outputMat, m->out, aOffset);
    x.shape = keep;
}

// th-llama.cpp:270-452, n_tokens > 1 branch (the reference's `pb` path, generalised): any batch size up to
// n_batch, a causal mask that honours n_past (the reference's masked softmax ignores it), and the
// weight matmuls on the tensor cores (cmdbuf_mat_mul -> thk_gemm_f16_tc).  Pipelines are not cached here:
// the shapes change with the batch size.
static void build_layer_cmdbuf_batch(WGPUDevice device, WGPUCommandEncoder encoder, std::shared_ptr<LlamaModel> m, LlamaLayer& l,
                                     int n_tokens, int n_past) {
    if (!encoder) return;                         // nothing to pre-build
    TensorBuffer& queryBuf = m->inp[1];
    TensorBuffer& keyBuf = m->inp[2];
    TensorBuffer& valueBuf = m->inp[3];
    TensorBuffer& queryTranspose = m->inp[4];
    const int n_embd = m->n_embd, n_head = m->n_head, hd = n_embd / n_head, T = n_tokens, N = n_past + n_tokens;
    const TensorShape rows{0, 0, T, n_embd};
    for (TensorBuffer* t : {&m->inp[0], &queryBuf, &keyBuf, &valueBuf, &queryTranspose, &m->inp[6]}) t->shape = rows;

    cmdbuf_rms_norm(device, encoder, nullptr, nullptr, m->inp[0]);
    cmdbuf_row_element_multiply(device, encoder, nullptr, nullptr, m->inp[0], l.attention_norm);
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, m->inp[0], l.wq, queryBuf, 1);             // :308-310
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, m->inp[0], l.wk, keyBuf, 1);
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, m->inp[0], l.wv, valueBuf, 1);

    queryBuf.shape = TensorShape{0, T, n_head, hd};
    keyBuf.shape = TensorShape{0, T, n_head, hd};
    cmdbuf_RoPE(device, encoder, nullptr, nullptr, queryBuf, m->networkUniforms);
    cmdbuf_RoPE(device, encoder, nullptr, nullptr, keyBuf, m->networkUniforms);
    const size_t off = (size_t)n_past * n_embd * sizeof(float), bytes = (size_t)T * n_embd * sizeof(float);
    copy_command(device, l.key_cache.gpu, off, keyBuf.gpu, 0, bytes);                                // :337-338
    copy_command(device, l.value_cache.gpu, off, valueBuf.gpu, 0, bytes);

    l.key_cache.shape.b = N;
    l.value_cache.shape.b = N;
    m->working_key_cache.shape = TensorShape{0, n_head, N, hd};
    m->working_val_cache.shape = TensorShape{0, n_head, N, hd};
    queryTranspose.shape = TensorShape{0, n_head, T, hd};
    cmdbuf_transpose(device, encoder, nullptr, nullptr, l.key_cache, m->working_key_cache, true, m->dimsUniforms[0]);
    cmdbuf_transpose(device, encoder, nullptr, nullptr, l.value_cache, m->working_val_cache, true, m->dimsUniforms[0]);
    cmdbuf_transpose(device, encoder, nullptr, nullptr, queryBuf, queryTranspose, true, m->dimsUniforms[1]);

    std::swap(m->working_key_cache.shape.r, m->working_key_cache.shape.c);                       // logical [H, D, N] for A.c == B.r
    m->inp[5].shape = TensorShape{0, n_head, T, N};
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, queryTranspose, m->working_key_cache, m->inp[5], 1, m->dimsUniforms[2]);
    cmdbuf_masked_softmax(device, encoder, nullptr, nullptr, m->inp[5], m->dimsUniforms[3]);     // causal, row i sees <= n_past + i
    keyBuf.shape = TensorShape{0, n_head, T, hd};
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, m->inp[5], m->working_val_cache, keyBuf, 0, m->dimsUniforms[4]);
    valueBuf.shape = TensorShape{0, T, n_head, hd};
    cmdbuf_transpose(device, encoder, nullptr, nullptr, keyBuf, valueBuf, true, nullptr);        // [H,T,D] -> [T,H,D]

    valueBuf.shape = rows;
    m->inp[1].shape = rows;
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, valueBuf, l.wo, m->inp[1], 1);             // :404
    m->inp[6].shape = rows; m->inp[2].shape = rows; m->inp[3].shape = rows;
    cmdbuf_addition(device, encoder, nullptr, nullptr, m->inp[1], m->inp[6], m->inp[2]);
    copy_command(device, m->inp[3].gpu, 0, m->inp[2].gpu, 0, bytes);
    cmdbuf_rms_norm(device, encoder, nullptr, nullptr, m->inp[2]);
    cmdbuf_row_element_multiply(device, encoder, nullptr, nullptr, m->inp[2], l.ffn_norm);

    m->ffWorking[0].shape.r = T;
    m->ffWorking[1].shape.r = T;
    std::swap(l.w1.shape.r, l.w1.shape.c);                                                       // :427-428: B described as [K, N]
    std::swap(l.w3.shape.r, l.w3.shape.c);
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, m->inp[2], l.w1, m->ffWorking[0], 1);
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, m->inp[2], l.w3, m->ffWorking[1], 1);
    cmdbuf_silu(device, encoder, nullptr, nullptr, m->ffWorking[0]);
    cmdbuf_element_mult_in_place(device, encoder, nullptr, nullptr, m->ffWorking[0], m->ffWorking[1]);
    std::swap(l.w2.shape.r, l.w2.shape.c);                                                       // :444
    cmdbuf_mat_mul(device, encoder, nullptr, nullptr, m->ffWorking[0], l.w2, m->inp[2], 1);
    m->inp[0].shape = rows;
    cmdbuf_addition(device, encoder, nullptr, nullptr, m->inp[3], m->inp[2], m->inp[0]);
    copy_command(device, m->inp[6].gpu, 0, m->inp[0].gpu, 0, bytes);
    reset_layer_tensors(l);
}

// th-llama.cpp:270-452, n_tokens == 1 branch.  Buffer roles as in the reference: inp0 = x (normed in
// place), inp6 = residual copy, inp1..3 = Q/K/V, inp4 = Q^T, inp5 = scores.
void build_layer_cmdbuf(WGPUDevice device, WGPUCommandEncoder encoder, std::shared_ptr<LlamaModel> m, LlamaLayer& l,
                        LlamaLayerComputePipeline& p, int n_tokens, int n_past) {
    reset_working_memory_tensors(*m);
    reset_layer_tensors(l);
    if (n_tokens != 1) { build_layer_cmdbuf_batch(device, encoder, m, l, n_tokens, n_past); return; }

    TensorBuffer& queryBuf = m->inp[1];
    TensorBuffer& keyBuf = m->inp[2];
    TensorBuffer& valueBuf = m->inp[3];
    TensorBuffer& queryTranspose = m->inp[4];
    const int n_embd = m->n_embd, n_head = m->n_head, hd = n_embd / n_head;

    for (TensorBuffer* t : {&m->inp[0], &queryBuf, &keyBuf, &valueBuf, &queryTranspose, &m->inp[6]}) t->shape = TensorShape{0, 0, n_tokens, n_embd};

    cmdbuf_rms_norm(device, encoder, nullptr, &p.p01, m->inp[0]);                                   // :299
    cmdbuf_row_element_multiply(device, encoder, nullptr, &p.p02, m->inp[0], l.attention_norm);     // :300
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p03_mm, m->inp[0], l.wq, queryBuf, 0); // :304
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p03_mm, m->inp[0], l.wk, keyBuf, 0);   // :305
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p03_mm, m->inp[0], l.wv, valueBuf, 0); // :306

    queryBuf.shape = TensorShape{0, n_tokens, n_head, hd};                                          // :317-319
    keyBuf.shape = TensorShape{0, n_tokens, n_head, hd};
    valueBuf.shape = TensorShape{0, n_head, n_tokens, hd};
    cmdbuf_RoPE(device, encoder, nullptr, &p.p04_rope, queryBuf, m->networkUniforms);               // :321
    cmdbuf_RoPE(device, encoder, nullptr, &p.p04_rope, keyBuf, m->networkUniforms);                 // :322

    if (encoder) {   // KV append, :332-339
        const size_t off = (size_t)n_past * n_embd * sizeof(float);
        copy_command(device, l.key_cache.gpu, off, keyBuf.gpu, 0, keyBuf.get_size_bytes());
        copy_command(device, l.value_cache.gpu, off, valueBuf.gpu, 0, valueBuf.get_size_bytes());
    }

    l.key_cache.shape.b = n_past + n_tokens;                                                        // :341-350
    m->working_key_cache.shape = l.key_cache.shape;
    std::swap(m->working_key_cache.shape.b, m->working_key_cache.shape.r);
    l.value_cache.shape.b = n_past + n_tokens;
    m->working_val_cache.shape = l.value_cache.shape;
    std::swap(m->working_val_cache.shape.b, m->working_val_cache.shape.r);
    queryTranspose.shape = queryBuf.shape;
    std::swap(queryTranspose.shape.b, queryTranspose.shape.r);

    cmdbuf_transpose(device, encoder, nullptr, &p.p05_trans, l.key_cache, m->working_key_cache, true, m->dimsUniforms[0]);   // :353
    cmdbuf_transpose(device, encoder, nullptr, &p.p05_trans, l.value_cache, m->working_val_cache, true, m->dimsUniforms[0]); // :354
    cmdbuf_transpose(device, encoder, nullptr, &p.p05_trans, queryBuf, queryTranspose, true, m->dimsUniforms[1]);            // :355

    std::swap(m->working_key_cache.shape.r, m->working_key_cache.shape.c);                          // :361
    m->inp[5].shape = TensorShape{0, n_head, n_tokens, n_tokens + n_past};
    cmdbuf_mat_mul(device, encoder, nullptr, &p.p06_mm, queryTranspose, m->working_key_cache, m->inp[5], 1, m->dimsUniforms[2]); // :365
    cmdbuf_row_softmax(device, encoder, nullptr, &p.p07_softmax, m->inp[5], m->dimsUniforms[3]);    // :373

    keyBuf.shape = TensorShape{0, n_head, n_tokens, hd};
    cmdbuf_mat_mul(device, encoder, nullptr, &p.p08_mm, m->inp[5], m->working_val_cache, keyBuf, 0, m->dimsUniforms[4]);     // :380

    std::swap(valueBuf.shape.b, valueBuf.shape.r);                                                  // :396
    cmdbuf_transpose(device, encoder, nullptr, &p.p09_t, keyBuf, valueBuf, true, nullptr);          // :397

    valueBuf.shape = TensorShape{0, 0, n_tokens, n_embd};
    m->inp[1].shape = TensorShape{0, 0, n_tokens, n_embd};
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p10_mm, valueBuf, l.wo, m->inp[1], 0); // :402

    m->inp[6].shape = m->inp[1].shape;
    m->inp[2].shape = m->inp[1].shape;
    cmdbuf_addition(device, encoder, nullptr, &p.p11_add, m->inp[1], m->inp[6], m->inp[2]);         // :409
    if (encoder) copy_command(device, m->inp[3].gpu, 0, m->inp[2].gpu, 0, m->inp[2].get_size_bytes());  // :412

    cmdbuf_rms_norm(device, encoder, nullptr, &p.p12_rms, m->inp[2]);                               // :415
    cmdbuf_row_element_multiply(device, encoder, nullptr, &p.p13_norm, m->inp[2], l.ffn_norm);      // :416

    m->ffWorking[0].shape.r = n_tokens;
    m->ffWorking[1].shape.r = n_tokens;
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p14_mm, m->inp[2], l.w1, m->ffWorking[0], 0);  // :423
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p14_mm, m->inp[2], l.w3, m->ffWorking[1], 0);  // :424
    cmdbuf_silu(device, encoder, nullptr, &p.p15_silu, m->ffWorking[0]);                            // :436
    cmdbuf_element_mult_in_place(device, encoder, nullptr, &p.p16_hadamard, m->ffWorking[0], m->ffWorking[1]); // :438
    cmdbuf_vector_mat_mul_trans(device, encoder, nullptr, &p.p17_mm, m->ffWorking[0], l.w2, m->inp[2], 0);  // :441

    m->inp[3].shape = m->inp[2].shape;
    m->inp[0].shape = m->inp[2].shape;
    cmdbuf_addition(device, encoder, nullptr, &p.p18_add, m->inp[3], m->inp[2], m->inp[0]);         // :447
    if (encoder) copy_command(device, m->inp[6].gpu, 0, m->inp[0].gpu, 0, m->inp[0].get_size_bytes());  // :450
}

// greedy branch of th-llama.cpp:814-838; other temperatures are outside the hot path
// th-llama.cpp:814-907.  temp <= 0: the token with the highest logit (lowest id wins ties).  Otherwise: logits / temp,
// CTRL repetition penalty on the ids in last_n_tokens (negative scores are multiplied, positive ones divided), top-k by
// partial sort, softmax over the survivors with a double accumulator, nucleus cut at cumulative probability top_p and
// renormalisation, then one draw from std::discrete_distribution with the model's mt19937.  Same library calls in the
// same order as the reference, so with the same seed the draw is the same token (tests/golden/sampler.json).
tk_llama_token llama_sample_top_p_top_k(std::shared_ptr<LlamaModel> m, const std::vector<tk_llama_token>& last_n_tokens, int top_k,
                                        float top_p, float temp, float repeat_penalty, std::vector<float>& logits) {
    const int n_logits = m->n_vocab;
    if ((int)logits.size() < n_logits || n_logits <= 0) return -1;
    return llama_sample_logits(m->rng, logits.data() + logits.size() - n_logits, n_logits, last_n_tokens, top_k, top_p, temp, repeat_penalty);
}
tk_llama_token llama_sample_logits(std::mt19937& rng, const float* pl, int n_logits, const std::vector<tk_llama_token>& last_n_tokens, int top_k,
                                   float top_p, float temp, float repeat_penalty) {
    if (temp <= 0) {
        float max_logit = pl[0];
        tk_llama_token max_id = 0;
        for (int i = 1; i < n_logits; ++i)
            if (pl[i] > max_logit) { max_logit = pl[i]; max_id = i; }
        return max_id;
    }
    typedef std::pair<float, tk_llama_token> Cand;
    std::vector<Cand> cand;
    cand.reserve(n_logits);
    const float scale = 1.0f / temp;
    for (int i = 0; i < n_logits; ++i) {
        const bool seen = std::find(last_n_tokens.begin(), last_n_tokens.end(), i) != last_n_tokens.end();
        float v = pl[i] * scale;
        if (seen) v = (pl[i] < 0.0f) ? v * repeat_penalty : v / repeat_penalty;
        cand.push_back(Cand(v, i));
    }
    if (top_k > 0 && top_k < n_logits) {
        std::partial_sort(cand.begin(), cand.begin() + top_k, cand.end(), [](const Cand& a, const Cand& b) { return a.first > b.first; });
        cand.resize(top_k);
    }
    float maxl = -std::numeric_limits<float>::infinity();
    for (const Cand& c : cand) maxl = std::max(maxl, c.first);
    std::vector<float> probs;
    probs.reserve(cand.size());
    double sum = 0.0;
    for (const Cand& c : cand) {
        const float pr = expf(c.first - maxl);
        probs.push_back(pr);
        sum += pr;
    }
    for (float& pr : probs) pr /= sum;
    if (top_p < 1.0) {
        double cumsum = 0.0;
        for (int i = 0; i < (int)probs.size(); ++i) {
            cumsum += probs[i];
            if (cumsum >= top_p) {
                probs.resize(i + 1);
                cand.resize(i + 1);
                break;
            }
        }
        cumsum = 1.0 / cumsum;
        for (float& pr : probs) pr *= cumsum;
    }
    std::discrete_distribution<> dist(probs.begin(), probs.end());
    return cand[dist(rng)].second;
}

// ---- tokenizer (TkLlamaTokenizer, th-llama.cpp:909-1041): SentencePiece-style greedy merging ----
// The text is cut into UTF-8 characters; every adjacent pair whose concatenation is a vocabulary entry is a merge
// candidate scored by that entry; the best candidate (highest score, then leftmost) is applied until none is left.
// Pieces that are not in the vocabulary come out as byte tokens (byte value + 3).
namespace {
struct Piece { int prev, next; size_t off, len; };
struct Merge { int left, right; float score; size_t size; };
struct MergeLess {   // priority_queue keeps the LARGEST on top: higher score first, then the smaller left index
    bool operator()(const Merge& a, const Merge& b) const { return a.score < b.score || (a.score == b.score && a.left > b.left); }
};
size_t utf8_char_len(unsigned char lead) {
    static const unsigned char len_by_high_nibble[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4};
    return len_by_high_nibble[lead >> 4];
}
}  // namespace

std::vector<tk_llama_token> tk_llama_tokenize(const LlamaVocab& vocab, const std::string& text, bool add_bos) {
    std::vector<tk_llama_token> out;
    if (text.empty()) return out;
    if (add_bos) out.push_back(tk_llama_token_bos());
    std::vector<Piece> pieces;
    for (size_t off = 0; off < text.size();) {
        const size_t len = std::min(text.size() - off, utf8_char_len((unsigned char)text[off]));
        Piece pc;
        pc.prev = (int)pieces.size() - 1;
        pc.off = off;
        pc.len = len;
        off += len;
        pc.next = off == text.size() ? -1 : (int)pieces.size() + 1;
        pieces.push_back(pc);
    }
    std::priority_queue<Merge, std::vector<Merge>, MergeLess> queue;
    auto propose = [&](int left, int right) {
        if (left < 0 || right < 0) return;
        const std::string joined = text.substr(pieces[left].off, pieces[left].len + pieces[right].len);
        auto it = vocab.token_to_id.find(joined);
        if (it == vocab.token_to_id.end() || (size_t)it->second >= vocab.id_to_token.size()) return;
        queue.push(Merge{left, right, vocab.id_to_token[it->second].score, joined.size()});
    };
    for (size_t i = 1; i < pieces.size(); ++i) propose((int)i - 1, (int)i);
    while (!queue.empty()) {
        const Merge mg = queue.top();
        queue.pop();
        Piece& l = pieces[mg.left];
        Piece& r = pieces[mg.right];
        if (l.len == 0 || r.len == 0 || l.len + r.len != mg.size) continue;      // stale: one side was merged meanwhile
        l.len += r.len;
        r.len = 0;
        l.next = r.next;
        if (r.next >= 0) pieces[r.next].prev = mg.left;
        propose(l.prev, mg.left);
        propose(mg.left, l.next);
    }
    for (int i = 0; i != -1; i = pieces[i].next) {
        const Piece& pc = pieces[i];
        auto it = vocab.token_to_id.find(text.substr(pc.off, pc.len));
        if (it != vocab.token_to_id.end()) { out.push_back(it->second); continue; }
        for (size_t j = 0; j < pc.len; ++j) out.push_back((tk_llama_token)(unsigned char)text[pc.off + j] + 3);
    }
    return out;
}
std::vector<tk_llama_token> tk_llama_tokenize(std::shared_ptr<LlamaModel> m, const std::string& text, bool add_bos) {
    return tk_llama_tokenize(m->vocab, text, add_bos);
}
int tk_llama_tokenize(std::shared_ptr<LlamaModel> m, const char* text, tk_llama_token* tokens, int n_max_tokens, bool add_bos) {
    const std::vector<tk_llama_token> res = tk_llama_tokenize(m->vocab, text, add_bos);
    if (n_max_tokens < (int)res.size()) {
        fprintf(stderr, "%s: too many tokens\n", __func__);
        return -(int)res.size();
    }
    for (size_t i = 0; i < res.size(); ++i) tokens[i] = res[i];
    return (int)res.size();
}
const char* tk_llama_token_to_str(std::shared_ptr<LlamaModel> m, tk_llama_token token) {      // th-llama.cpp:1094-1100
    if (token < 0 || (size_t)token >= m->vocab.id_to_token.size()) return nullptr;
    return m->vocab.id_to_token[token].tok.c_str();
}
tk_llama_token tk_llama_token_bos() { return 1; }
tk_llama_token tk_llama_token_eos() { return 2; }

static bool write_uniforms(WGPUQueue queue, std::shared_ptr<LlamaModel> m, int n_tokens, int n_past) {
    // th-llama.cpp:479-550
    const uint32_t H = m->n_head, D = m->n_embd / m->n_head, Np = (uint32_t)(n_past + n_tokens);
    LlamaNetworkUniforms nu{};
    nu.n_past = (uint32_t)n_past; nu.n_tokens = (uint32_t)n_tokens;
    LlamaTensorDimsUniforms d[5]{};
    d[0].A_B = Np ? Np : 1; d[0].A_M = H; d[0].A_N = D;                                                   // cache transpose
    d[1].A_B = (uint32_t)n_tokens; d[1].A_M = H; d[1].A_N = D;                                          // q transpose
    d[2].A_B = H; d[2].A_M = (uint32_t)n_tokens; d[2].A_N = D; d[2].scale = 1.0f / sqrtf((float)D);
    d[2].B_B = H; d[2].B_M = D; d[2].B_N = Np;                                                           // QK^T
    d[3].A_B = H; d[3].A_M = (uint32_t)n_tokens; d[3].A_N = Np;                                          // softmax
    d[4].A_B = H; d[4].A_M = (uint32_t)n_tokens; d[4].A_N = Np; d[4].scale = 1.0f;
    d[4].B_B = H; d[4].B_M = Np; d[4].B_N = D;                                                           // PV
    if (thk_upload(queue, m->networkUniforms, 0, &nu, kLlamaUniformsSize)) return false;
    for (int i = 0; i < 5; ++i)
        if (thk_upload(queue, m->dimsUniforms[i], 0, &d[i], kLlamaUniformsSize)) return false;
    return true;
}

static bool eval_one_opgraph(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, tk_llama_token token, int n_past) {
    if (!write_uniforms(queue, m, 1, n_past)) return false;
    // embedding row -> inp0 and inp6 on the device (the kUseGpuEmbeddingSelection branch, th-llama.cpp:552-575)
    reset_working_memory_tensors(*m);
    m->inp[0].shape = TensorShape{0, 0, 1, m->n_embd};
    m->inp[6].shape = m->inp[0].shape;
    const int64_t stride = m->tok_embeddings.shape.c * (int64_t)get_TensorType_size(m->tok_embeddings.type);
    TensorBuffer& emb = m->tok_embeddings;
    const TensorShape keep = emb.shape;
    emb.shape = TensorShape{0, 0, 1, m->n_embd};
    CommandBuffer cb = cmdbuf_f16_f32_conversion(device, kEncoder, nullptr, nullptr, m->inp[0], emb, 4, 0, (int)(stride * token));
    emb.shape = keep;
    if (!cb.is_valid()) return false;
    copy_command(device, m->inp[6].gpu, 0, m->inp[0].gpu, 0, m->inp[0].get_size_bytes());    // :573
    for (auto& l : m->layers) build_layer_cmdbuf(device, kEncoder, m, l, m->ps, 1, n_past);   // :596-618
    reset_working_memory_tensors(*m);
    build_final_compute_cmdbuf(device, kEncoder, m, m->pfs, 1);                               // :632
    return true;
}

// n_tokens > 1 in one pass (the reference's 8-token `pb` path, th-llama.cpp:600-608, for any size <= n_batch)
static bool eval_batch_opgraph(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, const tk_llama_token* tokens, int T,
                               int n_past) {
    if (!write_uniforms(queue, m, T, n_past)) return false;
    reset_working_memory_tensors(*m);
    m->inp[0].shape = TensorShape{0, 0, 1, m->n_embd};
    const int64_t stride = (int64_t)m->n_embd * (int64_t)get_TensorType_size(m->tok_embeddings.type);
    TensorBuffer& emb = m->tok_embeddings;
    const TensorShape keep = emb.shape;
    emb.shape = TensorShape{0, 0, 1, m->n_embd};
    bool ok = true;
    for (int i = 0; i < T && ok; ++i) {             // embedding rows -> inp0 (th-llama.cpp:552-575)
        if (tokens[i] < 0 || tokens[i] >= m->n_vocab) { ok = false; break; }
        ok = cmdbuf_f16_f32_conversion(device, kEncoder, nullptr, nullptr, m->inp[0], emb, 4, (int)(i * m->n_embd * 4),
                                       (int)(stride * tokens[i])).is_valid();
    }
    emb.shape = keep;
    if (!ok) return false;
    copy_command(device, m->inp[6].gpu, 0, m->inp[0].gpu, 0, (size_t)T * m->n_embd * sizeof(float));   // :573
    for (auto& l : m->layers) {
        build_layer_cmdbuf(device, kEncoder, m, l, m->pb, T, n_past);
        if (kv_rows_to_hpd(*m, l.key_cache, l.key_cache_hpd, n_past, T)) return false;
        if (kv_rows_to_hpd(*m, l.value_cache, l.value_cache_hpd, n_past, T)) return false;
    }
    reset_working_memory_tensors(*m);
    LlamaFinalComputePipeline pf;                                     // uncached: shapes depend on the batch size
    pf.p01.buildPipelineFlag = pf.p02.buildPipelineFlag = pf.p03.buildPipelineFlag = pf.p03_reduce.buildPipelineFlag = false;
    build_final_compute_cmdbuf(device, kEncoder, m, pf, T);          // logits of the last row (aOffset)
    return thk_gemm_check(device) == THK_OK;
}

// One authoritative KV row per position: before a path evaluates at n_past, the rows [0, n_past) it has not written
// itself are copied over from the other path's layout; after it wrote [n_past, n_past + T) the other layout is stale
// from n_past on.  (Tensor parallel models only run the fused path.)
static bool kv_sync(std::shared_ptr<LlamaModel> m, bool want_hpd, int n_past, int T) {
    if (m->tp_size == 1) {
        int32_t& have = want_hpd ? m->kvValidHpd : m->kvValidPhd;
        const int32_t other = want_hpd ? m->kvValidPhd : m->kvValidHpd;
        if (have < n_past) {
            if (other < n_past) { fprintf(stderr, "th_eval_gpu: n_past %d but only %d positions were evaluated\n", n_past, std::max(have, other)); return false; }
            for (auto& l : m->layers) {
                const int rc = want_hpd
                    ? (kv_rows_to_hpd(*m, l.key_cache, l.key_cache_hpd, have, n_past - have) ||
                       kv_rows_to_hpd(*m, l.value_cache, l.value_cache_hpd, have, n_past - have))
                    : (kv_rows_from_hpd(*m, l.key_cache_hpd, l.key_cache, have, n_past - have) ||
                       kv_rows_from_hpd(*m, l.value_cache_hpd, l.value_cache, have, n_past - have));
                if (rc) { fprintf(stderr, "th_eval_gpu: %s\n", thk_last_error()); return false; }
            }
        }
        have = n_past + T;
        int32_t& stale = want_hpd ? m->kvValidPhd : m->kvValidHpd;
        stale = std::min<int32_t>(stale, n_past);
    }
    return true;
}

tk_llama_token th_eval_gpu(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, const tk_llama_token* tokens,
                           int n_tokens, int n_past) {
    if (th_eval_gpu_launch(device, queue, m, tokens, n_tokens, n_past) != 0) return -1;
    return th_eval_gpu_finish(device, queue, m);
}

// First half of th_eval_gpu: upload the token, enqueue the evaluation (no host wait).  Returns 0 on success.
int th_eval_gpu_launch(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, const tk_llama_token* tokens, int n_tokens,
                       int n_past) {
    if (!m || !tokens || n_tokens < 1) { fprintf(stderr, "th_eval_gpu: bad arguments\n"); return -1; }
    if (n_past < 0 || n_past + n_tokens > m->n_ctx) {
        fprintf(stderr, "th_eval_gpu: n_past %d + n_tokens %d exceeds the context (%d)\n", n_past, n_tokens, m->n_ctx);
        if (m->onError) m->onError("context full");
        return -1;
    }
    const int64_t launches0 = g_launch_count;
    int64_t fused_launches = 0;
    if (n_tokens > 1 && m->tp_size == 1 && m->batchPrefill) {
        // batched prompt: chunks of up to n_batch tokens through the op graph with tensor-core matmuls, then the
        // new KV rows are handed to the fused decoder's cache layout
        for (int t0 = 0; t0 < n_tokens; t0 += m->n_batch) {
            const int T = std::min<int>(m->n_batch, n_tokens - t0);
            if (!kv_sync(m, false, n_past + t0, T)) return -1;
            if (m->kvValidHpd >= n_past + t0) m->kvValidHpd = n_past + t0 + T;      // eval_batch_opgraph hands its rows to the fused layout too
            if (!eval_batch_opgraph(device, queue, m, tokens + t0, T, n_past + t0)) { fprintf(stderr, "th_eval_gpu: batched prefill failed\n"); return -1; }
        }
        m->gpuLaunches = g_launch_count - launches0;
        return 0;
    }
    for (int i = 0; i < n_tokens; ++i) {
        const tk_llama_token tok = tokens[i];
        if (tok < 0 || tok >= m->n_vocab) { fprintf(stderr, "th_eval_gpu: token %d outside the vocabulary\n", tok); return -1; }
        if (m->evalPath == EvalPath_Fused) {
            if (!m->decoder) { fprintf(stderr, "th_eval_gpu: fused decoder missing\n"); return -1; }
            if (!kv_sync(m, true, n_past + i, 1)) return -1;
            if (thk_upload(queue, m->d_token, 0, &tok, sizeof(int32_t))) { fprintf(stderr, "th_eval_gpu: %s\n", thk_last_error()); return -1; }
            if (thk_decoder_step(m->decoder, m->d_token, n_past + i, (float*)m->out.gpu, m->d_next, nullptr)) {
                fprintf(stderr, "th_eval_gpu: %s\n", thk_last_error());
                return -1;
            }
            fused_launches += 1;
        } else {
            if (!kv_sync(m, false, n_past + i, 1)) return -1;
            if (!eval_one_opgraph(device, queue, m, tok, n_past + i)) { fprintf(stderr, "th_eval_gpu: op graph failed\n"); return -1; }
        }
    }
    m->gpuLaunches = (g_launch_count - launches0) + fused_launches;
    return 0;
}

// Second half of th_eval_gpu: logits -> pinned host buffer (resultBuffer + map, th-llama.cpp:646-706), then
// sample on the host.  Split out so that several tensor-parallel ranks driven by one thread can all be
// launched before any of them is waited for.
tk_llama_token th_eval_gpu_finish(WGPUDevice, WGPUQueue queue, std::shared_ptr<LlamaModel> m) {
    const int64_t nlocal = m->n_vocab / m->tp_size;
    if (thk_download(queue, m->pinnedLogits, m->out.gpu, 0, (size_t)nlocal * sizeof(float))) {
        fprintf(stderr, "th_eval_gpu: %s\n", thk_last_error());
        return -1;
    }
    if (m->evalPath == EvalPath_Fused && thk_decoder_check(m->decoder)) { fprintf(stderr, "th_eval_gpu: %s\n", thk_last_error()); return -1; }
    m->lastLogits.assign(m->pinnedLogits, m->pinnedLogits + nlocal);
    if (m->tp_size > 1) {
        // each rank holds a vocabulary slice; the greedy id over the whole vocabulary was agreed on inside the kernel
        int32_t tok = -1;
        if (thk_download(queue, &tok, m->d_next, 0, sizeof tok)) { fprintf(stderr, "th_eval_gpu: %s\n", thk_last_error()); return -1; }
        if (m->samplerTemp > 0) { fprintf(stderr, "th_eval_gpu: tensor-parallel evaluation samples greedily\n"); return -1; }
        return tok;
    }
    return llama_sample_top_p_top_k(m, {}, 40, 0.95f, m->samplerTemp, 1.10f, m->lastLogits);
}

std::vector<tk_llama_token> generate_greedy(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m,
                                            const std::vector<tk_llama_token>& prompt, int n_new) {
    std::vector<tk_llama_token> out;
    tk_llama_token tok = -1;
    for (tk_llama_token t : prompt) {            // prompt is consumed one token per evaluation (kAllowedSubsequentBatchSize = 1)
        tok = th_eval_gpu(device, queue, m, &t, 1, m->n_past);
        if (tok < 0) return out;
        m->n_past += 1;
    }
    for (int i = 0; i < n_new && tok >= 0; ++i) {
        out.push_back(tok);
        m->lastGeneratedToken = tok;
        if (m->n_past >= m->n_ctx) break;
        tok = th_eval_gpu(device, queue, m, &tok, 1, m->n_past);
        m->n_past += 1;
    }
    return out;
}

// do_inference, synchronous flavour (th-llama.cpp:111-168, 199-238): a leading space is added to the first prompt of a
// context, the prompt is tokenised with BOS and consumed one token per evaluation (kAllowedSubsequentBatchSize = 1,
// th-llama.cpp:15) while last_n_tokens slides, then tokens are generated until EOS, max_new_tokens, or the context is
// full; every generated piece goes to onNewToken(piece, message so far) and the whole message is returned.
// Differences from the reference: the limit is a parameter (reference: kMaxOutputTokens = 500 over a 512 context), the
// sampler temperature is m->samplerTemp (reference: 0.8 hard-coded), nothing is printed.
std::string do_inference(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, std::string prompt, int max_new_tokens) {
    if (m->n_past >= m->n_ctx) {
        const std::string error = "Maximum context reached (" + std::to_string(m->n_ctx) + "). Please reset context.";
        if (m->onError) m->onError(error);
        return std::string();
    }
    if (m->n_past == 0) prompt.insert(0, 1, ' ');
    m->embd_inp = tk_llama_tokenize(m, prompt, true);
    m->n_consumed = 0;
    if ((int)m->embd_inp.size() > m->n_ctx - 4 - m->n_past) {
        fprintf(stderr, "%s: error: prompt is too long (%d tokens, max %d)\n", __func__, (int)m->embd_inp.size(), m->n_ctx - 4 - m->n_past);
        if (m->onError) m->onError("prompt is too long");
        return std::string();
    }
    if (m->last_n_tokens.empty()) m->last_n_tokens.assign((size_t)m->n_ctx, 0);
    m->generatedMessage.clear();
    int produced = 0;
    while (m->n_past < m->n_ctx) {
        tk_llama_token in;
        const bool from_prompt = (int)m->embd_inp.size() > m->n_consumed;
        if (from_prompt) {
            in = m->embd_inp[m->n_consumed++];
            m->last_n_tokens.erase(m->last_n_tokens.begin());
            m->last_n_tokens.push_back(in);
        } else {
            in = m->lastGeneratedToken;
            if (in == tk_llama_token_eos() || produced >= max_new_tokens) break;
            const char* piece = tk_llama_token_to_str(m, in);
            m->generatedMessage += piece ? piece : "";
            ++produced;
            if (m->onNewToken) m->onNewToken(piece ? piece : "", m->generatedMessage);
        }
        const tk_llama_token next = th_eval_gpu(device, queue, m, &in, 1, m->n_past);
        if (next < 0) { if (m->onError) m->onError("evaluation failed"); break; }
        m->n_past += 1;
        m->lastGeneratedToken = next;
    }
    if (m->onInferenceComplete) m->onInferenceComplete(m->generatedMessage);
    return m->generatedMessage;
}

}  // namespace th
