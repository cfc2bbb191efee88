// esmdiff_b200 -- VQ-VAE structure encoder context and its C ABI (include/esmdiff_b200.h, "structure encoder").
// Second translation unit of libesmdiff_b200.so; kernels in encoder.cuh / geom.cuh.
#include "../../include/esmdiff_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "encoder.cuh"
#include "geom.cuh"

using namespace esmdiff;

static std::string g_enc_create_error;

struct EncLayerW {
    float *s_norm = nullptr, *proj = nullptr, *out_proj = nullptr, *w_dist = nullptr, *w_rot = nullptr;
    float *ln_w = nullptr, *ln_b = nullptr, *w1 = nullptr, *w2 = nullptr;
};

struct esmdiff_encoder {
    esmdiff_encoder_cfg cfg;
    int device = 0;
    std::string err;
    std::vector<EncLayerW> layers;
    float *norm_w = nullptr, *vq_w = nullptr, *vq_b = nullptr, *codebook = nullptr, *code_sq = nullptr, *relpos = nullptr;
    std::set<std::string> loaded;
    bool finalized = false;
    std::vector<void*> owned;
    // workspace (grows with the largest B * L seen)
    long long ws_res = 0;
    float *rot = nullptr, *trans = nullptr, *x = nullptr, *ns = nullptr, *p = nullptr, *att = nullptr, *u = nullptr,
          *hb = nullptr, *zq = nullptr, *zout = nullptr, *dots = nullptr;
    unsigned char* mask = nullptr;
    int *edges = nullptr, *frame_idx = nullptr;
    std::vector<void*> ws_owned;

    int fail(const std::string& m) {
        err = m;
        return 1;
    }
    template <typename T>
    int alloc(T** q, size_t n, std::vector<void*>& pool) {
        void* r = nullptr;
        const cudaError_t e = cudaMalloc(&r, (n ? n : 1) * sizeof(T));
        if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        pool.push_back(r);
        *q = reinterpret_cast<T*>(r);
        return 0;
    }
};

#define ECK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return c->fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

struct EncSlot {
    float** dst;
    long long rows, cols;
};

bool enc_resolve(esmdiff_encoder* c, const std::string& key, EncSlot* s, bool* dropped) {
    const esmdiff_encoder_cfg& g = c->cfg;
    const long long D = g.d_model, H = g.v_heads, F = g.ffn_hidden;
    *dropped = false;
    if (key.rfind("codebook.", 0) == 0 && key != "codebook.embeddings") {       // EMA bookkeeping buffers
        *dropped = true;
        return true;
    }
    if (key == "transformer.norm.weight") { *s = {&c->norm_w, D, 1}; return true; }
    if (key == "pre_vq_proj.weight") { *s = {&c->vq_w, g.d_out, D}; return true; }
    if (key == "pre_vq_proj.bias") { *s = {&c->vq_b, g.d_out, 1}; return true; }
    if (key == "codebook.embeddings") { *s = {&c->codebook, g.n_codes, g.d_out}; return true; }
    if (key == "relative_positional_embedding.embedding.weight") { *s = {&c->relpos, 2 * g.rel_bins + 2, D}; return true; }
    const std::string pre = "transformer.blocks.";
    if (key.rfind(pre, 0) != 0) return false;
    const size_t dot = key.find('.', pre.size());
    if (dot == std::string::npos) return false;
    const int l = atoi(key.substr(pre.size(), dot - pre.size()).c_str());
    if (l < 0 || l >= g.n_layers) return false;
    EncLayerW& w = c->layers[l];
    const std::string rest = key.substr(dot + 1);
    if (rest == "geom_attn.s_norm.weight") { *s = {&w.s_norm, D, 1}; return true; }
    if (rest == "geom_attn.proj.weight") { *s = {&w.proj, 15 * H, D}; return true; }
    if (rest == "geom_attn.out_proj.weight") { *s = {&w.out_proj, D, 3 * H}; return true; }
    if (rest == "geom_attn.distance_scale_per_head") { *s = {&w.w_dist, H, 1}; return true; }
    if (rest == "geom_attn.rotation_scale_per_head") { *s = {&w.w_rot, H, 1}; return true; }
    if (rest == "ffn.0.weight") { *s = {&w.ln_w, D, 1}; return true; }
    if (rest == "ffn.0.bias") { *s = {&w.ln_b, D, 1}; return true; }
    if (rest == "ffn.1.weight") { *s = {&w.w1, 2 * F, D}; return true; }
    if (rest == "ffn.3.weight") { *s = {&w.w2, D, F}; return true; }
    return false;
}

std::vector<std::string> enc_required(const esmdiff_encoder* c) {
    std::vector<std::string> k = {"transformer.norm.weight", "pre_vq_proj.weight", "pre_vq_proj.bias", "codebook.embeddings",
                                  "relative_positional_embedding.embedding.weight"};
    for (int l = 0; l < c->cfg.n_layers; ++l) {
        const std::string p = "transformer.blocks." + std::to_string(l) + ".";
        for (const char* r : {"geom_attn.s_norm.weight", "geom_attn.proj.weight", "geom_attn.out_proj.weight",
                              "geom_attn.distance_scale_per_head", "geom_attn.rotation_scale_per_head", "ffn.0.weight",
                              "ffn.0.bias", "ffn.1.weight", "ffn.3.weight"})
            k.push_back(p + r);
    }
    return k;
}

int enc_sgemm(esmdiff_encoder* c, int epi, const float* A, long long lda, const float* W, float* C, long long ldc,
              const float* bias, long long M, int N, int K, float scale, cudaStream_t st) {
    if (K % 16 != 0 || lda % 4 != 0) return c->fail("encoder sgemm: K must be a multiple of 16 and lda of 4");
    if (M <= 0) return 0;
    const dim3 grid((N + 127) / 128, static_cast<unsigned>((M + 127) / 128));
    if (epi == 0) enc::sgemm_tn_kernel<0><<<grid, 256, 0, st>>>(A, W, C, bias, (int)M, N, K, lda, ldc, scale);
    else enc::sgemm_tn_kernel<1><<<grid, 256, 0, st>>>(A, W, C, bias, (int)M, N, K, lda, ldc, scale);
    ECK(cudaGetLastError());
    return 0;
}

template <typename TOut>
int launch_geom_attention(const float* r, const float* rot, const unsigned char* mask, const int* frame_idx,
                          const float* w_rot, const float* w_dist, TOut* out, int ldo, long long G, int S, int H,
                          int zero_frameless, cudaStream_t st) {
    // queries per thread: all of a 16-key neighbourhood in 4 passes, 8 at a time for long sequences
    if (S <= 16) {
        const dim3 grid((S + 3) / 4, static_cast<unsigned>(G));
        geom::attention_kernel<4, TOut><<<grid, H, 0, st>>>(r, rot, mask, frame_idx, w_rot, w_dist, out, ldo, S, H, zero_frameless);
    } else {
        const dim3 grid((S + 7) / 8, static_cast<unsigned>(G));
        geom::attention_kernel<8, TOut><<<grid, H, 0, st>>>(r, rot, mask, frame_idx, w_rot, w_dist, out, ldo, S, H, zero_frameless);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace

// Shared with esmdiff_b200.cu (block 0's live geometric attention): frames of a coordinate batch, and the
// rotate + attention pair on a bf16 projection.
int esmdiff_geom_frames(const float* coords, int B, int L, float* rot, float* trans, unsigned char* mask, cudaStream_t st) {
    geom::frames_kernel<<<B, 256, 0, st>>>(coords, rot, trans, mask, L);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int esmdiff_geom_attention_bf16(const __nv_bfloat16* proj, float* work, const float* rot, const float* trans,
                                const unsigned char* mask, const float* w_rot, const float* w_dist, __nv_bfloat16* out,
                                int ldo, int B, int T, int H, cudaStream_t st) {
    const long long M = static_cast<long long>(B) * T, nv = M * 5 * H;
    geom::rotate_kernel<__nv_bfloat16><<<static_cast<unsigned>((nv + 255) / 256), 256, 0, st>>>(proj, work, rot, trans, nullptr, M, H);
    if (cudaGetLastError() != cudaSuccess) return 1;
    return launch_geom_attention<__nv_bfloat16>(work, rot, mask, nullptr, w_rot, w_dist, out, ldo, B, T, H, 1, st);
}

extern "C" {

const char* esmdiff_encoder_last_error(const esmdiff_encoder* c) { return c ? c->err.c_str() : g_enc_create_error.c_str(); }

int esmdiff_encoder_create(const esmdiff_encoder_cfg* cfg, int device, esmdiff_encoder** out) {
    if (!cfg || !out) { g_enc_create_error = "encoder_create: null argument"; return 1; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        g_enc_create_error = "encoder_create: no CUDA device (esmdiff_b200 has no CPU fallback)";
        cudaGetLastError();
        return 1;
    }
    if (device < 0 || device >= n) { g_enc_create_error = "encoder_create: bad device index"; return 1; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) { g_enc_create_error = "encoder_create: this library is built for sm_100a (B200) only"; return 1; }
    if (cfg->d_model <= 0 || cfg->d_model % 16 != 0 || cfg->v_heads <= 0 || cfg->v_heads > 256 || (3 * cfg->v_heads) % 16 != 0 ||
        cfg->ffn_hidden <= 0 || cfg->ffn_hidden % 16 != 0 || cfg->d_out <= 0 || cfg->d_out % 16 != 0 || cfg->n_codes <= 0 ||
        cfg->n_layers <= 0 || cfg->knn <= 0 || cfg->rel_bins <= 0) {
        g_enc_create_error = "encoder_create: d_model, 3 v_heads, ffn_hidden, d_out must be positive multiples of 16, v_heads <= 256";
        return 1;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_enc_create_error = "encoder_create: cudaSetDevice failed"; return 1; }
    esmdiff_encoder* c = new esmdiff_encoder();
    c->cfg = *cfg;
    c->device = device;
    c->layers.resize(cfg->n_layers);
    *out = c;
    return 0;
}

int esmdiff_encoder_destroy(esmdiff_encoder* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    for (void* p : c->owned) if (p) cudaFree(p);
    for (void* p : c->ws_owned) if (p) cudaFree(p);
    delete c;
    return 0;
}

int esmdiff_encoder_set_weight(esmdiff_encoder* c, const char* key, const void* data, int on_device, int dtype,
                               const int64_t* shape, int ndim) {
    if (!c || !key || !data) return 1;
    if (dtype != ESMDIFF_F32) return c->fail(std::string("encoder set_weight: fp32 only (") + key + ")");
    EncSlot s;
    bool dropped = false;
    if (!enc_resolve(c, key, &s, &dropped)) return c->fail(std::string("encoder set_weight: unknown key ") + key);
    if (dropped) return 0;
    long long n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    if (n != s.rows * s.cols) return c->fail(std::string("encoder set_weight: shape mismatch for ") + key);
    ECK(cudaSetDevice(c->device));
    if (!*s.dst && c->alloc(s.dst, static_cast<size_t>(n), c->owned)) return 1;
    ECK(cudaMemcpy(*s.dst, data, static_cast<size_t>(n) * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    c->loaded.insert(key);
    c->finalized = false;
    return 0;
}

int esmdiff_encoder_finalize(esmdiff_encoder* c) {
    if (!c) return 1;
    for (const std::string& k : enc_required(c))
        if (!c->loaded.count(k)) return c->fail("encoder finalize: missing weight " + k);
    ECK(cudaSetDevice(c->device));
    if (!c->code_sq && c->alloc(&c->code_sq, c->cfg.n_codes, c->owned)) return 1;
    enc::rowsumsq_kernel<<<(c->cfg.n_codes + 7) / 8, 256>>>(c->codebook, c->code_sq, c->cfg.n_codes, c->cfg.d_out);
    ECK(cudaGetLastError());
    ECK(cudaDeviceSynchronize());
    c->finalized = true;
    return 0;
}

int esmdiff_encode_structure(esmdiff_encoder* c, const float* coords, const int64_t* residue_index, int B, int L,
                             int64_t* codes_out, float* z_out, int32_t* edges_out, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("encode_structure: esmdiff_encoder_finalize has not succeeded");
    if (!coords || !codes_out || B <= 0 || L <= 0) return c->fail("encode_structure: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const esmdiff_encoder_cfg& g = c->cfg;
    const int D = g.d_model, H = g.v_heads, F = g.ffn_hidden, E = g.knn < L ? g.knn : L;
    const long long R = static_cast<long long>(B) * L, M = R * E;
    if (M > (1ll << 22)) return c->fail("encode_structure: B * L * knn too large for one call (split the batch)");
    ECK(cudaSetDevice(c->device));
    if (R > c->ws_res) {
        ECK(cudaStreamSynchronize(st));
        for (void* p : c->ws_owned) if (p) cudaFree(p);
        c->ws_owned.clear();
        c->ws_res = 0;
        const long long Mc = R * g.knn;
        if (c->alloc(&c->rot, R * 9, c->ws_owned) || c->alloc(&c->trans, R * 3, c->ws_owned) || c->alloc(&c->mask, R, c->ws_owned) ||
            c->alloc(&c->edges, Mc, c->ws_owned) || c->alloc(&c->frame_idx, Mc, c->ws_owned) || c->alloc(&c->x, Mc * D, c->ws_owned) ||
            c->alloc(&c->ns, Mc * D, c->ws_owned) || c->alloc(&c->p, Mc * 15 * H, c->ws_owned) || c->alloc(&c->att, Mc * 3 * H, c->ws_owned) ||
            c->alloc(&c->u, Mc * 2 * F, c->ws_owned) || c->alloc(&c->hb, Mc * F, c->ws_owned) || c->alloc(&c->zq, R * D, c->ws_owned) ||
            c->alloc(&c->zout, R * g.d_out, c->ws_owned) || c->alloc(&c->dots, R * g.n_codes, c->ws_owned))
            return 1;
        c->ws_res = R;
    }
    const float rs = sqrtf(static_cast<float>(g.n_layers) / 36.0f);           // TransformerStack(scale_residue=True)
    geom::frames_kernel<<<B, 256, 0, st>>>(coords, c->rot, c->trans, c->mask, L);
    enc::knn_kernel<<<dim3(L, B), 128, 0, st>>>(coords, c->mask, c->edges, L, E);
    enc::relpos_gather_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, st>>>(
        c->edges, reinterpret_cast<const long long*>(residue_index), c->relpos, c->x, c->frame_idx, M, L, E, D, g.rel_bins);
    ECK(cudaGetLastError());
    const unsigned rows_grid = static_cast<unsigned>((M + 7) / 8);
    for (int l = 0; l < g.n_layers; ++l) {
        const EncLayerW& w = c->layers[l];
        enc::layernorm_f32_kernel<<<rows_grid, 256, 0, st>>>(c->x, D, w.s_norm, nullptr, c->ns, M, D, nullptr);
        if (enc_sgemm(c, 0, c->ns, D, w.proj, c->p, 15 * H, nullptr, M, 15 * H, D, 1.f, st)) return 1;
        const long long nv = M * 5 * H;
        geom::rotate_kernel<float><<<static_cast<unsigned>((nv + 255) / 256), 256, 0, st>>>(c->p, c->p, c->rot, c->trans, c->frame_idx, M, H);
        if (launch_geom_attention<float>(c->p, c->rot, c->mask, c->frame_idx, w.w_rot, w.w_dist, c->att, 3 * H, R, E, H, 0, st))
            return c->fail("encode_structure: geometric attention launch failed");
        if (enc_sgemm(c, 1, c->att, 3 * H, w.out_proj, c->x, D, nullptr, M, D, 3 * H, 1.0f / rs, st)) return 1;
        enc::layernorm_f32_kernel<<<rows_grid, 256, 0, st>>>(c->x, D, w.ln_w, w.ln_b, c->ns, M, D, nullptr);
        if (enc_sgemm(c, 0, c->ns, D, w.w1, c->u, 2 * F, nullptr, M, 2 * F, D, 1.f, st)) return 1;
        enc::swiglu_kernel<<<static_cast<unsigned>((M * F + 255) / 256), 256, 0, st>>>(c->u, c->hb, M, F);
        if (enc_sgemm(c, 1, c->hb, F, w.w2, c->x, D, nullptr, M, D, F, 1.0f / rs, st)) return 1;
        ECK(cudaGetLastError());
    }
    // the query node is neighbour 0 of its own neighbourhood (distance 0 sorts first): rows r * E
    enc::layernorm_f32_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, st>>>(c->x, static_cast<long long>(E) * D, c->norm_w,
                                                                               nullptr, c->zq, R, D, c->mask);
    if (enc_sgemm(c, 0, c->zq, D, c->vq_w, c->zout, g.d_out, c->vq_b, R, g.d_out, D, 1.f, st)) return 1;
    if (enc_sgemm(c, 0, c->zout, g.d_out, c->codebook, c->dots, g.n_codes, nullptr, R, g.n_codes, g.d_out, 1.f, st)) return 1;
    enc::codebook_argmin_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, st>>>(
        c->zout, c->dots, c->code_sq, reinterpret_cast<long long*>(codes_out), R, g.n_codes, g.d_out);
    ECK(cudaGetLastError());
    if (z_out) ECK(cudaMemcpyAsync(z_out, c->zout, static_cast<size_t>(R) * g.d_out * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (edges_out) ECK(cudaMemcpyAsync(edges_out, c->edges, static_cast<size_t>(M) * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int esmdiff_op_backbone_frames(const float* coords, int B, int L, float* rot_out, float* trans_out, uint8_t* mask_out,
                               void* stream) {
    if (!coords || !rot_out || !trans_out || !mask_out || B <= 0 || L <= 0) return 1;
    return esmdiff_geom_frames(coords, B, L, rot_out, trans_out, mask_out, reinterpret_cast<cudaStream_t>(stream));
}

int esmdiff_op_geometric_attention(const float* proj, const float* rot, const float* trans, const uint8_t* mask,
                                   const int32_t* frame_idx, const float* rot_scale, const float* dist_scale, int G, int S,
                                   int H, int zero_frameless, float* work, float* out, void* stream) {
    if (!proj || !rot || !trans || !mask || !rot_scale || !dist_scale || !work || !out || G <= 0 || S <= 0 || H <= 0 || H > 256)
        return 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long M = static_cast<long long>(G) * S, nv = M * 5 * H;
    geom::rotate_kernel<float><<<static_cast<unsigned>((nv + 255) / 256), 256, 0, st>>>(proj, work, rot, trans, frame_idx, M, H);
    if (cudaGetLastError() != cudaSuccess) return 1;
    return launch_geom_attention<float>(work, rot, mask, frame_idx, rot_scale, dist_scale, out, 3 * H, G, S, H, zero_frameless, st);
}

}  // extern "C"
