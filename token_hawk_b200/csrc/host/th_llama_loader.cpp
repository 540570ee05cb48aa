// th_llama_loader.cpp -- GGML `ggjt` v1 loader + model initialisation (mirrors kayvr/token-hawk
// th-llama-loader.cpp).  Same file format, tensor names and entry points; differences:
//   * tensors are uploaded straight to device memory; output.weight is NOT split into halves
//     (the 256 MB WebGPU cap, th-llama-loader.cpp:197-242, does not exist here);
//   * tok_embeddings stays f16 on the device (the reference keeps an f32 CPU copy, :180-196);
//   * n_ff is derived, not asserted to be 11008 (:349-350); n_ctx is a parameter.
#include "th/th-llama-loader.hpp"

#include <string.h>
#include <algorithm>

#include <fstream>

namespace th {
bool g_default_kv_f16 = false;   // models created from now on keep the fused path's KV cache in f16 (LlamaModel::kvF16)


static const int32_t kftype_f32 = 0, kftype_f16 = 1;
static const uint32_t ggmlMagicUnversioned = 0x67676d6c, ggmlMagicValue = 0x67676a74, llamaFileVersion = 1;

namespace {
struct Reader {
    const char* p; size_t n, off = 0; bool ok = true;
    template <typename T> T get() {
        T v{};
        if (off + sizeof(T) > n) { ok = false; return v; }
        memcpy(&v, p + off, sizeof(T));
        off += sizeof(T);
        return v;
    }
    bool raw(void* dst, size_t k) {
        if (off + k > n) { ok = false; return false; }
        memcpy(dst, p + off, k);
        off += k;
        return true;
    }
};
}  // namespace

// th-llama-loader.cpp:47-119
bool load_header(LlamaModel* m, void* data, int64_t dataSize, int64_t) {
    Reader r{(const char*)data, (size_t)dataSize};
    const uint32_t magic = r.get<uint32_t>();
    if (magic == ggmlMagicUnversioned) { printf("ERROR: load_header: Old version of magic.\n"); return false; }
    if (magic != ggmlMagicValue) { printf("ERROR: load_header: Invalid magic value.\n"); return false; }
    if (r.get<uint32_t>() != llamaFileVersion) { printf("ERROR: load_header: Invalid file version.\n"); return false; }
    m->n_vocab = r.get<int32_t>(); m->n_embd = r.get<int32_t>(); m->n_mult = r.get<int32_t>(); m->n_head = r.get<int32_t>();
    m->n_layer = r.get<int32_t>(); m->n_rot = r.get<int32_t>(); m->f16 = r.get<int32_t>();
    if (!r.ok || m->n_vocab <= 0 || m->n_vocab > (1 << 24)) { printf("ERROR: load_header: truncated or implausible header.\n"); return false; }
    m->vocab.id_to_token.resize(m->n_vocab);
    std::string word;
    for (int i = 0; i < m->n_vocab; ++i) {
        const uint32_t len = r.get<uint32_t>();
        if (len > 8096) { printf("ERROR: load_header: Vocabulary element should not be larger than 8096.\n"); return false; }
        word.resize(len);
        if (len && !r.raw(&word[0], len)) return false;
        const float score = r.get<float>();
        if (!r.ok) return false;
        m->vocab.token_to_id[word] = i;
        m->vocab.id_to_token[i].tok = word;
        m->vocab.id_to_token[i].score = score;
    }
    return r.ok;
}

// th-llama-loader.cpp:121-265
bool load_weights(LlamaModel* m, WGPUDevice device, WGPUQueue queue, void* data, int64_t dataSize, int64_t numElementsInFile,
                  int64_t originalFileOffset) {
    Reader r{(const char*)data, (size_t)dataSize};
    for (int64_t i = 0; i < numElementsInFile; ++i) {
        const int32_t ndims = r.get<int32_t>(), nameLen = r.get<int32_t>(), ftype = r.get<int32_t>();
        if (!r.ok || ndims < 1 || ndims > 3 || nameLen < 0 || nameLen > 512 || ftype < 0) { printf("Detected an error\n"); return false; }
        TensorShape shape{};
        if (ndims >= 1) shape.c = r.get<int32_t>();
        if (ndims >= 2) shape.r = r.get<int32_t>();
        if (ndims >= 3) shape.b = r.get<int32_t>();
        TensorType type = TensorType_Unknown;
        if (ftype == kftype_f32) type = TensorType_F32;
        else if (ftype == kftype_f16) type = TensorType_F16;
        else { printf("ERROR: Quantized formats not supported yet\n"); return false; }
        std::string name((size_t)nameLen, '\0');
        if (nameLen && !r.raw(&name[0], (size_t)nameLen)) return false;
        const int64_t abs = originalFileOffset + (int64_t)r.off;          // data is padded to a 32-byte FILE offset
        r.off += (size_t)(((abs + 31) & -32) - abs);
        shape.canonicalize();
        const size_t bytes = (size_t)shape.get_total_num_elements() * get_TensorType_size(type);
        if (r.off + bytes > r.n) { printf("ERROR: tensor %s truncated\n", name.c_str()); return false; }
        TensorBuffer t(r.p + r.off, shape, type, false, device, queue);
        t.name = name;
        if (!t.gpu) { printf("ERROR: Failed to upload to GPU.\n"); return false; }
        m->loadedMapping[name] = std::move(t);
        r.off += bytes;
    }
    return true;
}

// Shard kinds (Megatron-style tensor parallel): rows of the local heads / ff units / vocab
// entries, or the matching columns for the matrices whose INPUT dimension is split (wo, w2).
enum ShardKind { Shard_None, Shard_Rows, Shard_Cols };

static TensorBuffer take(LlamaModel* m, const std::string& name, TensorType type, int64_t rows, int64_t cols, bool& ok,
                         ShardKind shard = Shard_None) {
    auto it = m->loadedMapping.find(name);
    if (it == m->loadedMapping.end() || !it->second.gpu) { fprintf(stderr, "model is missing tensor %s\n", name.c_str()); ok = false; return {}; }
    TensorBuffer t = std::move(it->second);
    m->loadedMapping.erase(it);
    const int64_t tp = m->tp_size, rk = m->tp_rank;
    if (tp > 1 && shard != Shard_None && t.shape.r == rows * (shard == Shard_Rows ? tp : 1) && t.shape.c == cols * (shard == Shard_Cols ? tp : 1)) {
        // the file (or a full synthetic tensor) holds the whole matrix: keep this rank's slice
        const size_t es = get_TensorType_size(t.type);
        TensorBuffer sl(TensorShape{0, 0, rows, cols}, t.type, m->device);
        sl.name = t.name;
        if (!sl.gpu) { ok = false; return {}; }
        if (shard == Shard_Rows) {
            thk_copy(m->device, sl.gpu, 0, t.gpu, (size_t)(rk * rows) * cols * es, (size_t)rows * cols * es);
        } else {
            const int64_t full = cols * tp;
            for (int64_t r = 0; r < rows; ++r)
                thk_copy(m->device, sl.gpu, (size_t)r * cols * es, t.gpu, ((size_t)r * full + rk * cols) * es, (size_t)cols * es);
        }
        thk_sync(m->device);
        t = std::move(sl);
    }
    if (t.type != type || t.shape.r != rows || t.shape.c != cols) {
        fprintf(stderr, "tensor %s: expected %s [%lld,%lld], file has %s %s\n", name.c_str(), get_TensorType_name(type).c_str(),
                (long long)rows, (long long)cols, get_TensorType_name(t.type).c_str(), t.shape.to_string().c_str());
        ok = false;
    }
    return t;
}

// th-llama-loader.cpp:330-435
void post_load_init_model(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m) {
    m->device = device;
    m->rng = std::mt19937(780658349);                                       // :332
    m->n_ff = ((2 * (4 * m->n_embd) / 3 + m->n_mult - 1) / m->n_mult) * m->n_mult;   // :349
    const int64_t E = m->n_embd, H = m->n_head, D = E / H, F = m->n_ff, V = m->n_vocab;
    const int64_t tp = m->tp_size > 0 ? m->tp_size : 1;
    bool ok = (E > 0 && H > 0 && E % H == 0 && m->n_layer > 0 && m->n_ctx > 0);
    if (ok && (H % tp || F % tp || V % tp || m->tp_rank < 0 || m->tp_rank >= tp)) { fprintf(stderr, "post_load_init_model: tp_size %lld must divide n_head, n_ff and n_vocab\n", (long long)tp); ok = false; }
    if (!ok) { fprintf(stderr, "post_load_init_model: bad hyper-parameters\n"); m->loadFailed = true; return; }
    const int64_t Eh = E / tp, Fh = F / tp, Vl = V / tp, Hl = H / tp;

    m->tok_embeddings = take(m.get(), "tok_embeddings.weight", TensorType_F16, V, E, ok);
    m->norm = take(m.get(), "norm.weight", TensorType_F32, 1, E, ok);
    m->outputMat = take(m.get(), "output.weight", TensorType_F16, Vl, E, ok, Shard_Rows);
    const TensorShape kvShape{0, m->n_ctx, Hl, D};         // [pos][head][dim], :335
    const TensorShape kvShapeHpd{0, Hl, m->n_ctx, D};      // [head][pos][dim] for the fused kernel
    for (int i = 0; i < m->n_layer && ok; ++i) {
        LlamaLayer l{};
        l.index = i;
        const std::string p = "layers." + std::to_string(i) + ".";
        l.attention_norm = take(m.get(), p + "attention_norm.weight", TensorType_F32, 1, E, ok);
        l.wq = take(m.get(), p + "attention.wq.weight", TensorType_F16, Eh, E, ok, Shard_Rows);
        l.wk = take(m.get(), p + "attention.wk.weight", TensorType_F16, Eh, E, ok, Shard_Rows);
        l.wv = take(m.get(), p + "attention.wv.weight", TensorType_F16, Eh, E, ok, Shard_Rows);
        l.wo = take(m.get(), p + "attention.wo.weight", TensorType_F16, E, Eh, ok, Shard_Cols);
        l.ffn_norm = take(m.get(), p + "ffn_norm.weight", TensorType_F32, 1, E, ok);
        l.w1 = take(m.get(), p + "feed_forward.w1.weight", TensorType_F16, Fh, E, ok, Shard_Rows);
        l.w2 = take(m.get(), p + "feed_forward.w2.weight", TensorType_F16, E, Fh, ok, Shard_Cols);
        l.w3 = take(m.get(), p + "feed_forward.w3.weight", TensorType_F16, Fh, E, ok, Shard_Rows);
        l.key_cache = TensorBuffer(kvShape, TensorType_F32, device);
        l.value_cache = TensorBuffer(kvShape, TensorType_F32, device);
        l.key_cache_hpd = TensorBuffer(kvShapeHpd, m->kvF16 ? TensorType_F16 : TensorType_F32, device);
        l.value_cache_hpd = TensorBuffer(kvShapeHpd, m->kvF16 ? TensorType_F16 : TensorType_F32, device);
        for (TensorBuffer* t : {&l.key_cache, &l.value_cache, &l.key_cache_hpd, &l.value_cache_hpd}) {
            if (!t->gpu) { ok = false; break; }
            thk_memset(device, t->gpu, 0, t->get_size_bytes());
        }
        m->layers.push_back(std::move(l));
    }
    if (!ok) { m->loadFailed = true; return; }

    m->working_key_cache = TensorBuffer(kvShape, TensorType_F32, device);   // :338-339
    m->working_val_cache = TensorBuffer(kvShape, TensorType_F32, device);
    // inp[5] holds the scores [n_head][1][n_ctx]; the reference's [8, n_embd] only covers ctx 512 (:341-344)
    if (m->n_batch > m->n_ctx) m->n_batch = m->n_ctx;
    int64_t inpCols = E;                       // per row; scores need n_head * n_batch * n_ctx floats in total
    if (Hl * (int64_t)m->n_ctx > inpCols) inpCols = Hl * (int64_t)m->n_ctx;
    for (int i = 0; i < LlamaModel::nInpBuffers; ++i) {
        m->inp[i] = TensorBuffer(TensorShape{0, 0, m->n_batch, inpCols}, TensorType_F32, device);
        m->inp[i].shape = m->inp[i].originalShape = TensorShape{0, 0, m->n_batch, E};
    }
    m->ffWorking[0] = TensorBuffer(TensorShape{0, 0, m->n_batch, F}, TensorType_F32, device);   // :351-352
    m->ffWorking[1] = TensorBuffer(TensorShape{0, 0, m->n_batch, F}, TensorType_F32, device);
    // workspace of the batched-prompt GEMM (f16 hi/lo split of up to n_batch activation rows), sized once here like every
    // other working buffer so that the evaluation path never allocates
    if (m->tp_size == 1 && thk_gemm_reserve(device, m->n_batch, std::max<int64_t>(E, F)) != THK_OK) fprintf(stderr, "post_load_init_model: %s\n", thk_last_error());
    m->out = TensorBuffer(TensorShape{0, 0, 1, Vl}, TensorType_F32, device);                    // :360
    m->outScratch = TensorBuffer(TensorShape{0, 0, 1, Vl}, TensorType_F32, device);
    void* p = nullptr;
    if (thk_host_alloc(device, (size_t)V * sizeof(float), &p) == THK_OK) m->pinnedLogits = (float*)p;     // resultBuffer, :363
    LlamaNetworkUniforms zero{};
    if (thk_malloc(device, kLlamaUniformsSize, &m->networkUniforms) == THK_OK) thk_upload(queue, m->networkUniforms, 0, &zero, kLlamaUniformsSize);
    for (auto& u : m->dimsUniforms)
        if (thk_malloc(device, kLlamaUniformsSize, &u) == THK_OK) thk_upload(queue, u, 0, &zero, kLlamaUniformsSize);
    void* t0 = nullptr; void* t1 = nullptr;
    thk_malloc(device, 16, &t0); thk_malloc(device, 16, &t1);
    m->d_token = (int32_t*)t0; m->d_next = (int32_t*)t1;
    if (!m->pinnedLogits || !m->networkUniforms || !m->d_token || !m->d_next || !m->out.gpu) { m->loadFailed = true; return; }

    if (tp == 1) build_pipelines_llama(device, queue, m);                                       // :434 (op graph: unsharded only)

    // fused decoder over the same weights
    thk_llama_dims dims{};
    dims.n_vocab = m->n_vocab; dims.n_embd = m->n_embd; dims.n_head = m->n_head; dims.n_layer = m->n_layer;
    dims.n_ff = m->n_ff; dims.n_ctx = m->n_ctx; dims.tp_rank = m->tp_rank; dims.tp_size = (int32_t)tp;
    dims.kv_f16 = m->kvF16 ? 1 : 0;
    std::vector<thk_llama_layer> L(m->n_layer);
    for (int i = 0; i < m->n_layer; ++i) {
        const LlamaLayer& l = m->layers[i];
        L[i].attention_norm = (const float*)l.attention_norm.gpu;
        L[i].wq = (const uint16_t*)l.wq.gpu; L[i].wk = (const uint16_t*)l.wk.gpu; L[i].wv = (const uint16_t*)l.wv.gpu;
        L[i].wo = (const uint16_t*)l.wo.gpu; L[i].ffn_norm = (const float*)l.ffn_norm.gpu;
        L[i].w1 = (const uint16_t*)l.w1.gpu; L[i].w2 = (const uint16_t*)l.w2.gpu; L[i].w3 = (const uint16_t*)l.w3.gpu;
        L[i].key_cache = (float*)l.key_cache_hpd.gpu; L[i].value_cache = (float*)l.value_cache_hpd.gpu;
    }
    if (thk_decoder_create(device, &dims, L.data(), (const uint16_t*)m->tok_embeddings.gpu, (const float*)m->norm.gpu,
                           (const uint16_t*)m->outputMat.gpu, &m->decoder) != THK_OK) {
        fprintf(stderr, "post_load_init_model: fused decoder unavailable (%s); op graph only\n", thk_last_error());
        m->decoder = nullptr;
        m->evalPath = EvalPath_OpGraph;
        if (tp > 1) { m->loadFailed = true; return; }
    }
    thk_sync(queue);
}

// th-llama-loader.cpp:485-635
std::shared_ptr<LlamaModel> load_llama_file(WGPUDevice device, WGPUQueue queue, const std::string& filename, int32_t n_ctx, int32_t tp_rank,
                                            int32_t tp_size) {
    std::ifstream fin(filename, std::ios::binary | std::ios::ate);
    if (!fin) { fprintf(stderr, "Unable to open file: %s\n", filename.c_str()); return {}; }
    const int64_t fileSize = (int64_t)fin.tellg();
    fin.seekg(0, std::ios::beg);
    auto m = std::make_shared<LlamaModel>();
    m->kvF16 = g_default_kv_f16;
    m->n_ctx = n_ctx;
    m->tp_rank = tp_rank; m->tp_size = tp_size > 0 ? tp_size : 1;
    m->device = device;

    // header + vocab: read a bounded prefix, parse with load_header, then find where it ended
    {
        uint32_t hdr[9];
        fin.read((char*)hdr, sizeof hdr);
        if (!fin) { fprintf(stderr, "%s: file too short\n", __func__); return {}; }
        if (hdr[0] == ggmlMagicUnversioned) { fprintf(stderr, "%s: invalid model file '%s' (too old)\n", __func__, filename.c_str()); return {}; }
        if (hdr[0] != ggmlMagicValue) { printf("%s: Bad magic\n", __func__); return {}; }
        if (hdr[1] != llamaFileVersion) { fprintf(stderr, "%s: unsupported format version %u\n", __func__, hdr[1]); return {}; }
        const int32_t n_vocab = (int32_t)hdr[2];
        if (n_vocab <= 0 || n_vocab > (1 << 24)) { fprintf(stderr, "%s: implausible n_vocab\n", __func__); return {}; }
        // walk the vocab to find the end of the header
        int64_t off = sizeof hdr;
        for (int i = 0; i < n_vocab; ++i) {
            uint32_t len = 0;
            fin.seekg(off, std::ios::beg);
            fin.read((char*)&len, 4);
            if (!fin || len > 8096) { fprintf(stderr, "%s: bad vocabulary entry %d\n", __func__, i); return {}; }
            off += 4 + (int64_t)len + 4;
        }
        if (off > fileSize) { fprintf(stderr, "%s: truncated vocabulary\n", __func__); return {}; }
        std::vector<char> head((size_t)off);
        fin.seekg(0, std::ios::beg);
        fin.read(head.data(), off);
        if (!load_header(m.get(), head.data(), off, 0)) return {};
    }

    std::vector<char> tensorData;
    while (true) {                                                           // :572-621
        const int64_t weightsBegin = (int64_t)fin.tellg();
        if (weightsBegin >= fileSize) break;
        int32_t n_dims = 0, length = 0, ftype = 0;
        fin.read((char*)&n_dims, 4); fin.read((char*)&length, 4); fin.read((char*)&ftype, 4);
        if (fin.eof() || !fin) break;
        if (n_dims < 1 || n_dims > 2 || length < 0 || length > 512) { fprintf(stderr, "%s: corrupt tensor header at %lld\n", __func__, (long long)weightsBegin); return {}; }
        int64_t nelements = 1;
        int32_t ne[2] = {1, 1};
        for (int i = 0; i < n_dims; ++i) { fin.read((char*)&ne[i], 4); nelements *= ne[i]; }
        if (ftype != kftype_f32 && ftype != kftype_f16) { printf("ERROR: Quantized formats not supported yet\n"); return {}; }
        const int64_t bytes = nelements * (ftype == kftype_f16 ? 2 : 4);
        int64_t cur = weightsBegin + 12 + 4 * n_dims + length;
        cur = ((cur + 31) & -32) + bytes;
        if (cur > fileSize) { fprintf(stderr, "%s: tensor data runs past the end of the file\n", __func__); return {}; }
        tensorData.resize((size_t)(cur - weightsBegin));
        fin.seekg(weightsBegin, std::ios::beg);
        fin.read(tensorData.data(), (std::streamsize)tensorData.size());
        if (!load_weights(m.get(), device, queue, tensorData.data(), (int64_t)tensorData.size(), 1, weightsBegin)) return {};
    }
    fin.close();
    post_load_init_model(device, queue, m);
    if (m->loadFailed) return {};
    m->loadedMapping.clear();
    return m;
}

static const char* kLayerTensorNames[9] = {"attention_norm.weight", "attention.wq.weight", "attention.wk.weight", "attention.wv.weight",
                                           "attention.wo.weight", "ffn_norm.weight", "feed_forward.w1.weight", "feed_forward.w2.weight",
                                           "feed_forward.w3.weight"};

std::shared_ptr<LlamaModel> create_synthetic_llama(WGPUDevice device, WGPUQueue queue, int32_t n_vocab, int32_t n_embd, int32_t n_mult,
                                                   int32_t n_head, int32_t n_layer, int32_t n_ctx, uint64_t seed, int32_t tp_rank,
                                                   int32_t tp_size) {
    auto m = std::make_shared<LlamaModel>();
    m->kvF16 = g_default_kv_f16;
    m->tp_rank = tp_rank; m->tp_size = tp_size > 0 ? tp_size : 1;
    m->device = device;
    if (n_head % m->tp_size || n_vocab % m->tp_size) { fprintf(stderr, "create_synthetic_llama: tp_size must divide n_head and n_vocab\n"); return {}; }
    m->n_vocab = n_vocab; m->n_embd = n_embd; m->n_mult = n_mult; m->n_head = n_head; m->n_layer = n_layer;
    m->n_rot = n_head ? n_embd / n_head : 0; m->n_ctx = n_ctx; m->f16 = 1;
    const int64_t E = n_embd, V = n_vocab;
    const int64_t F = ((2 * (4 * E) / 3 + n_mult - 1) / n_mult) * n_mult;
    bool ok = true;
    const int64_t tp = m->tp_size, rk = m->tp_rank;
    if (F % tp) { fprintf(stderr, "create_synthetic_llama: tp_size must divide n_ff\n"); return {}; }
    // shard: 0 none, 1 rows, 2 cols -- each rank generates only its slice of the full [R, C] tensor
    auto mat = [&](const std::string& name, uint64_t id, int64_t R, int64_t C, int shard = 0) {
        const int64_t r = shard == 1 ? R / tp : R, c = shard == 2 ? C / tp : C;
        TensorBuffer t(TensorShape{0, 0, r, c}, TensorType_F16, device);
        t.name = name;
        if (!t.gpu || thk_fill_f16(device, (uint16_t*)t.gpu, seed, id, r, c, shard == 1 ? rk * r : 0, shard == 2 ? rk * c : 0, C)) {
            fprintf(stderr, "synthetic %s: %s\n", name.c_str(), thk_last_error());
            ok = false;
        }
        m->loadedMapping[name] = std::move(t);
    };
    auto gain = [&](const std::string& name, uint64_t id) {
        TensorBuffer t(TensorShape{0, 0, 1, E}, TensorType_F32, device);
        t.name = name;
        if (!t.gpu || thk_fill_gain(device, (float*)t.gpu, seed, id, E)) { fprintf(stderr, "synthetic %s: %s\n", name.c_str(), thk_last_error()); ok = false; }
        m->loadedMapping[name] = std::move(t);
    };
    mat("tok_embeddings.weight", 0, V, E);
    gain("norm.weight", 1);
    mat("output.weight", 2, V, E, 1);
    for (int l = 0; l < n_layer && ok; ++l)
        for (int k = 0; k < 9; ++k) {
            const std::string name = "layers." + std::to_string(l) + "." + kLayerTensorNames[k];
            const uint64_t id = 3 + 9ull * l + k;
            if (k == 0 || k == 5) gain(name, id);
            else if (k == 6 || k == 8) mat(name, id, F, E, 1);
            else if (k == 7) mat(name, id, E, F, 2);
            else if (k == 4) mat(name, id, E, E, 2);
            else mat(name, id, E, E, 1);
        }
    if (!ok) return {};
    thk_sync(queue);
    post_load_init_model(device, queue, m);
    if (m->loadFailed) return {};
    m->loadedMapping.clear();
    return m;
}

bool fill_kv_synthetic(std::shared_ptr<LlamaModel> m, uint64_t seed, int n_positions) {
    if (!m || n_positions < 1 || n_positions > m->n_ctx) return false;
    const int64_t H = m->n_head, D = m->n_embd / m->n_head, Hl = H / m->tp_size, h0 = m->tp_rank * Hl;
    float* tmp = nullptr;                        // f16 KV: the synthetic values are generated in f32 and rounded like an append would
    if (m->kvF16) {
        void* t = nullptr;
        if (thk_malloc(m->device, (size_t)Hl * m->n_ctx * D * sizeof(float), &t) != THK_OK) return false;
        tmp = (float*)t;
    }
    bool ok = true;
    for (int l = 0; l < m->n_layer && ok; ++l) {
        LlamaLayer& L = m->layers[l];
        for (int kv = 0; kv < 2 && ok; ++kv) {
            TensorBuffer& hpd = kv == 0 ? L.key_cache_hpd : L.value_cache_hpd;
            TensorBuffer& phd = kv == 0 ? L.key_cache : L.value_cache;
            float* f32_hpd = m->kvF16 ? tmp : (float*)hpd.gpu;
            // fused layout [head][pos][dim] (local heads only under tensor parallelism)
            ok = ok && !thk_fill_kv(m->device, f32_hpd, seed, 1000 + 2ull * l + kv, n_positions, m->n_ctx, H, h0, Hl, D);
            if (m->kvF16) {
                const int64_t n = Hl * m->n_ctx * D;
                ok = ok && !thk_f32_to_f16(m->device, f32_hpd, (uint16_t*)hpd.gpu, n);
                ok = ok && !thk_f16_f32_conversion(m->device, f32_hpd, 0, (const uint16_t*)hpd.gpu, 0, n);   // the rounded values
            }
            // reference layout [pos][head][dim] = transpose of the above (zy on [H][n_ctx][D])
            ok = ok && !thk_transpose(m->device, f32_hpd, (float*)phd.gpu, Hl, m->n_ctx, D, 1, nullptr);
        }
    }
    if (tmp) { thk_sync(m->device); thk_free(m->device, tmp); }
    if (!ok) return false;
    m->kvValidPhd = m->tp_size == 1 ? n_positions : 0;
    m->kvValidHpd = n_positions;
    return thk_sync(m->device) == THK_OK;
}

}  // namespace th
