// esmdiff_b200: context, weights, forward orchestration and the C ABI (include/esmdiff_b200.h).
// Unity build: all kernels are included here so the library is one translation unit.
#include "../../include/esmdiff_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "attention.cuh"
#include "attention_resident.cuh"
#include "decoder.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "ptx.cuh"
#include "sampler.cuh"
#include "gibbs.cuh"

using namespace esmdiff;
typedef __nv_bfloat16 bf16;

// geom.cuh kernels, compiled in encoder.cu (second translation unit of the library)
int esmdiff_geom_frames(const float* coords, int B, int L, float* rot, float* trans, unsigned char* mask, cudaStream_t st);
int esmdiff_geom_attention_bf16(const __nv_bfloat16* proj, float* work, const float* rot, const float* trans,
                                const unsigned char* mask, const float* w_rot, const float* w_dist, __nv_bfloat16* out,
                                int ldo, int B, int T, int H, cudaStream_t st);

static std::string g_create_error;

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to (function, device) -- not to the process (a
// second device never saw the call) and not to a context (two contexts on one device share it: the
// one asking for less must not lower it under the other).  Largest value set so far per pair.
static std::map<std::pair<int, const void*>, int> g_smem_attr;
template <typename K>
static cudaError_t ensure_dynamic_smem(int device, K kernel, int bytes) {
    int& have = g_smem_attr[std::make_pair(device, reinterpret_cast<const void*>(kernel))];
    if (bytes <= have) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            return c->fail(std::string(#call) + ": " + cudaGetErrorString(e_));                \
        }                                                                                      \
    } while (0)

// Launch with programmatic stream serialization (ptx.cuh): only for kernels that call pdl_wait().
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct LayerW {
    float *ln1_w = nullptr, *ln1_b = nullptr, *qln_w = nullptr, *kln_w = nullptr;
    float *ln2_w = nullptr, *ln2_b = nullptr;
    bf16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
    // LayerNorm folded through the QKV / W1 GEMMs (gemm.cuh): fp32 masters kept from set_weight
    // until finalize folds gamma into the bf16 weights; c = colsum(gamma.W), b = beta W^T
    float *wqkv_f32 = nullptr, *w1_f32 = nullptr;
    float *cqkv = nullptr, *bqkv = nullptr, *c1 = nullptr, *b1 = nullptr;
    float* qk_gamma = nullptr;     // [2 D] q_ln.weight | k_ln.weight (QKV epilogue with q/k-LN + RoPE folded in)
    // block 0's geometric attention (esm GeometricReasoningOriginalImpl; live only with structure coordinates)
    float *g_snorm = nullptr, *g_wdist = nullptr, *g_wrot = nullptr;
    bf16 *g_proj = nullptr, *g_out = nullptr;
    bool dirty = true;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct esmdiff_ctx {
    esmdiff_cfg cfg;
    int device = 0;
    int num_sms = 148;
    std::string err;
    int64_t launches = 0;
    int attn_variant = 0;      // 0 = resident K/V where it fits (attention_resident.cuh), 1 = always the streaming kernel
                               // (ESMDIFF_ATTN=stream), 2 = resident without the CUDA-core leftover rows (tiles)
    bool ln_fold = true;       // block pre-LayerNorms folded through the GEMMs; ESMDIFF_LN=separate -> stand-alone kernel
    int* gibbs_cand = nullptr;         // per token row: candidate id of the current gibbs step (-1: not masked)
    float* gibbs_ent = nullptr;        // ... and the entropy of its filtered distribution
    int64_t gibbs_rows = 0;
    int attn_fold = 1;         // resident attention: fold a <= 4-key tail tile into the previous step (ESMDIFF_ATTN_FOLD=0: off)
    int attn_qsplit = 1;       // query-range split of the resident attention: 1 = off (default: measured no gain at 13 samples --
                               // the 16 CTAs of the second round run alone on their SMs and finish in half the time anyway),
                               // 0 = decide per launch, 2 = always; ESMDIFF_ATTN_QSPLIT
    int resid_bn = 0;          // tile width of the residual GEMMs: 0 = per launch (launch_gemm), 192 / 256 forced; ESMDIFF_RESID_BN
    const void* stats_ptr = nullptr;   // buffer the span below describes (single-kernel entry points)
    int stats_span = 128;      // columns per partial LayerNorm statistic currently in `stats` (128: embedding kernel and
                               // 256-wide residual tiles; 96: 192-wide residual tiles)
    int qkv_run = 0;           // tile schedule of the QKV GEMM with the RoPE epilogue (gemm.cuh TileSchedule; measured r2f at
                               // B=100, T=258: strided 234 us, runs of 2 / 3 / 6 column tiles 243 / 255 / 290 us, contiguous ranges
                               // 263 us -- re-reading the rotary table row per tile is cheaper than any loss of L2 locality
                               // or balance); ESMDIFF_QKV_RUN overrides
    bool skip_denoise = false; // ESMDIFF_SKIP_DENOISE_FORWARD=1: esmdiff_ddpm_sample leaves out the noise-removal forward when no
                               // MASK is left after the last step (exact: unmasked rows return themselves, model.py:575-579;
                               // SURVEY.md section 7 step 6).  Costs one stream synchronisation; off by default -- the
                               // benchmark runs the reference's 26 forwards
    int* dev_count = nullptr;
    bool pdl = true;           // programmatic dependent launch between the kernels of a forward; ESMDIFF_PDL=0 -> off
    bool qk_fused = true;      // q_ln / k_ln + RoPE folded into the QKV epilogue and the attention kernel
                               // (needs ln_fold); ESMDIFF_QK=separate -> stand-alone ew::qk_layernorm_rope_kernel
    EncodeTiledFn encode = nullptr;
    // esmdiff_set_structure_coords: backbone frames of the batch the next forwards run on (net.py:433-441);
    // geom_on == false is the ddpm path's NaN coordinates (geometric attention contributes exactly 0)
    bool geom_on = false;
    int geom_B = 0, geom_T = 0;
    int64_t geom_rows = 0;
    float *g_rot = nullptr, *g_trans = nullptr, *g_work = nullptr;
    unsigned char* g_mask = nullptr;
    bf16* g_projout = nullptr;

    std::vector<LayerW> layers;
    float *seq_embed = nullptr, *struct_embed = nullptr, *plddt_w = nullptr, *plddt_b = nullptr;
    float *res_w = nullptr, *res_b = nullptr, *ss8 = nullptr, *sasa = nullptr, *const_vec = nullptr;
    float *norm_w = nullptr, *h0_b = nullptr, *h2_w = nullptr, *h2_b = nullptr, *h3_b = nullptr;
    bf16 *h0_w = nullptr, *h3_w = nullptr;
    float *te_w0 = nullptr, *te_b0 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr;
    // model_kind 1 (VQ-VAE structure token decoder): token embedding and the pLDDT regression head;
    // h0/h2/h3 above then hold Dim6RotStructureHead's ffn1 / norm / proj
    float *dec_embed = nullptr, *p0_b = nullptr, *p2_w = nullptr, *p2_b = nullptr, *p3_b = nullptr;
    bf16 *p0_w = nullptr, *p3_w = nullptr;
    float* aux_ws = nullptr;                           // [rows][n_aux_out] pLDDT logits
    std::set<std::string> loaded;
    bool finalized = false;
    std::vector<void*> owned;

    // Two workspace sets.  The members below are the ACTIVE set (what the launchers read); activate(i)
    // swaps them with the parked one.  Launches are enqueued by one host thread and capture their
    // pointers at enqueue time, so switching between enqueues is safe.  Set 1 exists for the
    // two-stream sampling loop (esmdiff_ddpm_sample): small batches are split into two independent
    // halves that run on two streams and fill each other's partial waves.
    struct WorkspaceSet {
        int64_t ws_rows = 0;
        float *x = nullptr, *headh = nullptr, *logits_ws = nullptr, *qk_sumsq = nullptr, *aux_ws = nullptr;
        float *cond = nullptr, *te_hidden = nullptr;
        float2* stats = nullptr;
        bf16 *xn = nullptr, *qkv = nullptr, *att = nullptr, *hbuf = nullptr;
    };
    WorkspaceSet parked;
    int active_ws = 0;
    void swap_ws() {
        WorkspaceSet cur;
        cur.ws_rows = ws_rows; cur.x = x; cur.headh = headh; cur.logits_ws = logits_ws; cur.qk_sumsq = qk_sumsq;
        cur.aux_ws = aux_ws; cur.cond = cond; cur.te_hidden = te_hidden; cur.stats = stats; cur.xn = xn; cur.qkv = qkv;
        cur.att = att; cur.hbuf = hbuf;
        ws_rows = parked.ws_rows; x = parked.x; headh = parked.headh; logits_ws = parked.logits_ws;
        qk_sumsq = parked.qk_sumsq; aux_ws = parked.aux_ws; cond = parked.cond; te_hidden = parked.te_hidden;
        stats = parked.stats; xn = parked.xn; qkv = parked.qkv; att = parked.att; hbuf = parked.hbuf;
        parked = cur;
        active_ws ^= 1;
    }
    int activate(int i) {
        if (i != active_ws) swap_ws();
        if (!cond && (alloc(&cond, cfg.d_model) || alloc(&te_hidden, cfg.d_model))) return 1;
        return 0;
    }
    int64_t split_rows = 0;                    // ESMDIFF_SPLIT_ROWS: batches of at most this many token rows are sampled as two halves
                                               // (off by default: measured +1.2 % at 13 samples with 256-wide residual tiles, -2 % against
                                               // the 192-wide ones -- the chip is power-capped even there, so filling idle SMs buys clocks down)
    cudaStream_t sstream[2] = {nullptr, nullptr};
    cudaEvent_t ev_sfork = nullptr, ev_sjoin[2] = {nullptr, nullptr};
    // workspace (grows with the largest B*T seen)
    int64_t ws_rows = 0;
    float *x = nullptr, *headh = nullptr, *logits_ws = nullptr;
    float2* stats = nullptr;                           // [rows][d_model / 128] partial (mean, M2) of x
    bf16 *xn = nullptr, *qkv = nullptr, *att = nullptr, *hbuf = nullptr;
    float *cond = nullptr, *te_hidden = nullptr, *inv_freq = nullptr;
    float *cos_t = nullptr, *sin_t = nullptr, *rope_rows = nullptr, *colmean = nullptr;
    float* qk_sumsq = nullptr;                         // [rows][2 * d_model / 128] (gemm.cuh EPI_QKV_ROPE_LN)
    int rope_T = 0;
    int* dev_err = nullptr;
    int64_t *x_tok = nullptr, *seq_tok = nullptr;     // staging for the *_host entry point
    int64_t tok_rows = 0;

    std::map<std::tuple<const void*, uint64_t, uint64_t, uint32_t, uint32_t>, CUtensorMap> tmaps;

    // CUDA graphs of one forward for small batches (esmdiff_ddpm_sample): at B*T of a few hundred
    // rows the ~290 kernels of a forward are launch-latency bound (sweep: 24 us per launch against
    // 5-10 us of work).  Keyed by the shape and the token pointers the captured kernels read.
    int graph_mode = -1;                       // ESMDIFF_GRAPH: 0 never, 1 always, -1 when B*T <= graph_max_rows
    int64_t graph_max_rows = 8192;
    cudaStream_t gstream = nullptr;            // capture is illegal on the legacy default stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    struct GraphRec { cudaGraphExec_t exec = nullptr; int64_t kernels = 0; bool warmed = false; };
    std::map<std::tuple<int, int, const void*, const void*, int>, GraphRec> graphs;
    void drop_graphs() {
        for (auto& kv : graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        graphs.clear();
    }

    // optional per-launch device timing (esmdiff_profile_*): CUDA events on the launching stream
    struct ProfRec { int kind; double work; cudaEvent_t a, b; };
    bool prof = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    cudaEvent_t next_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }

    int fail(const std::string& m) {
        err = m;
        return 1;
    }
    template <typename T>
    int alloc(T** p, size_t n) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, n * sizeof(T));
        if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        owned.push_back(q);
        *p = reinterpret_cast<T*>(q);
        return 0;
    }
};

// Records an event pair around one kernel launch while profiling is on.
struct ProfScope {
    esmdiff_ctx* c;
    cudaStream_t st;
    size_t idx = 0;
    bool on;
    ProfScope(esmdiff_ctx* c_, int kind, double work, cudaStream_t st_) : c(c_), st(st_), on(c_->prof) {
        if (!on) return;
        esmdiff_ctx::ProfRec r;
        r.kind = kind; r.work = work;
        r.a = c->next_event(); r.b = c->next_event();
        cudaEventRecord(r.a, st);
        idx = c->prof_recs.size();
        c->prof_recs.push_back(r);
    }
    ~ProfScope() {
        if (on) cudaEventRecord(c->prof_recs[idx].b, st);
    }
};

// ------------------------------------------------------------------------------------------------
// TMA descriptors: row-major [rows, cols] (cols contiguous), box [box_rows][128 bytes], 128B swizzle
// ------------------------------------------------------------------------------------------------
static int get_tmap(esmdiff_ctx* c, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                    uint32_t box_rows, CUtensorMap* out, bool f32 = false) {
    const uint32_t esz = f32 ? 4 : 2;
    const uint32_t box_cols = 128 / esz;
    auto key = std::make_tuple(base, rows, cols * 1000003ull + ld_elems, box_rows, box_cols);
    auto it = c->tmaps.find(key);
    if (it != c->tmaps.end()) {
        *out = it->second;
        return 0;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * esz};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = c->encode(&tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[200];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu esz=%u base=%p", (int)r,
                 (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, esz, base);
        return c->fail(buf);
    }
    if (c->tmaps.size() > 4096) c->tmaps.clear();
    c->tmaps[key] = tm;
    *out = tm;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel launchers
// ------------------------------------------------------------------------------------------------
struct GemmLN {                // operands of the LayerNorm-folded epilogues (gemm.cuh)
    const float2* stats_in = nullptr;
    const float* colsum = nullptr;
    float2* stats_out = nullptr;
    bf16* xb_out = nullptr;
    int stats_span = 128;          // columns per partial statistic behind stats_in
    // EPI_QKV_ROPE_LN
    const float* rope = nullptr;
    const float* qk_gamma = nullptr;
    float* qk_sumsq = nullptr;
    int T = 0, n_rope = 0;
};

static int launch_gemm(esmdiff_ctx* c, int epi, const bf16* A, const bf16* W, int M, int N, int K,
                       void* out, int64_t ldo, const float* bias, float scale, cudaStream_t st,
                       const GemmLN& ln = GemmLN()) {
    if (K % gemm::BK != 0 || K <= 0) return c->fail("gemm: K must be a positive multiple of 64");
    const int max_clusters = c->num_sms / 2;
    const int m_tiles = (M + gemm::BM - 1) / gemm::BM;
    const bool is_store = epi == gemm::EPI_STORE_BF16 || epi == gemm::EPI_STORE_BF16_LN || epi == gemm::EPI_QKV_ROPE_LN;
    const bool is_swiglu = epi == gemm::EPI_SWIGLU_BF16 || epi == gemm::EPI_SWIGLU_BF16_LN;
    const bool is_resid = epi == gemm::EPI_RESID_F32 || epi == gemm::EPI_RESID_F32_LN;
    // Tile width of the residual GEMMs: 256, or 192 when that quantises better on the 74 CTA pairs.  A
    // 192-wide tile does 3/4 of the work of a 256-wide one at slightly lower efficiency (28 instead of 32
    // KiB of operands per 384 instead of 512 tensor clocks), so it pays when the wave count does not grow
    // by 4/3: N = 1536 at 13 samples is 84 tiles = 2 waves against 112 tiles = 2 waves of 3/4 the length.
    int BN = 256;
    if (is_resid && N % 192 == 0 && c->resid_bn != 256) {
        if (c->resid_bn == 192) {
            BN = 192;
        } else {
            const double w256 = ceil((double)m_tiles * (N / 256) / max_clusters), w192 = ceil((double)m_tiles * (N / 192) / max_clusters);
            if (N % 256 != 0 || w192 * 0.75 * 1.06 < w256) BN = 192;
        }
    }
    if ((is_swiglu || is_resid || epi == gemm::EPI_STORE_BF16_LN || epi == gemm::EPI_QKV_ROPE_LN) && N % BN != 0)
        return c->fail("gemm: SwiGLU / residual / LayerNorm-folded epilogues need N % 256 == 0");
    if (epi == gemm::EPI_QKV_ROPE_LN) {
        if (!ln.rope || !ln.qk_gamma || !ln.qk_sumsq || ln.T <= 0 || ln.n_rope <= 0 || ln.n_rope % BN != 0 || ln.n_rope > N)
            return c->fail("gemm: q/k-LayerNorm + RoPE epilogue needs the rotary table, gamma, a statistics buffer, T and n_rope % 256 == 0");
    }
    if (epi == gemm::EPI_STORE_BF16_LN || epi == gemm::EPI_SWIGLU_BF16_LN || epi == gemm::EPI_QKV_ROPE_LN) {
        if (!ln.stats_in || !ln.colsum || !bias || K % 256 != 0 || K > 1536)
            return c->fail("gemm: LayerNorm-folded epilogue needs statistics, colsum, bias and K % 256 == 0 <= 1536");
    }
    if (epi == gemm::EPI_RESID_F32_LN) {
        if (!ln.stats_out || !ln.xb_out || ldo != N)
            return c->fail("gemm: residual+statistics epilogue needs stats_out, xb_out and a dense [M, N] stream");
    }
    CUtensorMap ta, tb, tc;
    if (get_tmap(c, A, M, K, K, gemm::BM_CTA, &ta)) return 1;
    if (get_tmap(c, W, N, K, K, BN / 2, &tb)) return 1;
    tc = ta;                                     // unused by the direct-store epilogues
    if (is_store || is_swiglu) {
        const int out_cols = is_swiglu ? N / 2 : N;
        if (ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15))
            return c->fail("gemm: bf16 output needs a 16-byte aligned base and row stride");
        if (get_tmap(c, out, M, out_cols, ldo, 32, &tc)) return 1;
    } else if (is_resid) {
        if (ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 31))
            return c->fail("gemm: fp32 residual needs a 32-byte aligned base and row stride");
    }
    gemm::Params p;
    p.M = M; p.N = N; p.K = K;
    p.m_tiles = m_tiles;
    p.n_tiles = (N + BN - 1) / BN;
    p.out = out; p.ldo = ldo; p.bias = bias; p.scale = scale;
    p.stats_in = ln.stats_in; p.colsum = ln.colsum; p.stats_out = ln.stats_out; p.xb_out = ln.xb_out;
    p.ln_eps = 1e-5f;
    p.stats_span = ln.stats_span;
    p.rope = ln.rope; p.qk_gamma = ln.qk_gamma; p.qk_sumsq = ln.qk_sumsq; p.T = ln.T; p.n_rope = ln.n_rope;
    p.run = c->qkv_run > 0 && p.n_tiles % c->qkv_run != 0 ? -1 : c->qkv_run;
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = 2 * (tiles < max_clusters ? tiles : max_clusters);
    // profile kinds: the LayerNorm-folded variants are booked under their plain counterparts
    const int kind = (epi == gemm::EPI_STORE_BF16_LN || epi == gemm::EPI_QKV_ROPE_LN) ? gemm::EPI_STORE_BF16
                   : epi == gemm::EPI_RESID_F32_LN ? gemm::EPI_RESID_F32
                   : epi == gemm::EPI_SWIGLU_BF16_LN ? gemm::EPI_SWIGLU_BF16 : epi;
    ProfScope prof(c, kind, 2.0 * M * (double)N * K, st);
#define LAUNCH_GEMM(E, BNV)                                                                    \
    {                                                                                          \
        CK(ensure_dynamic_smem(c->device, gemm::gemm_bf16_tn_kernel<E, BNV>, gemm::Cfg<E, BNV>::SMEM_BYTES));       \
        CK(launch_pdl(c->pdl, gemm::gemm_bf16_tn_kernel<E, BNV>, dim3(grid), dim3(gemm::THREADS),                   \
                      gemm::Cfg<E, BNV>::SMEM_BYTES, st, ta, tb, tc, p));                                           \
    }
    switch (epi) {
        case gemm::EPI_STORE_BF16: LAUNCH_GEMM(gemm::EPI_STORE_BF16, 256) break;
        case gemm::EPI_RESID_F32:
            if (BN == 192) LAUNCH_GEMM(gemm::EPI_RESID_F32, 192) else LAUNCH_GEMM(gemm::EPI_RESID_F32, 256)
            break;
        case gemm::EPI_SWIGLU_BF16: LAUNCH_GEMM(gemm::EPI_SWIGLU_BF16, 256) break;
        case gemm::EPI_BIAS_GELU_F32: LAUNCH_GEMM(gemm::EPI_BIAS_GELU_F32, 256) break;
        case gemm::EPI_BIAS_F32: LAUNCH_GEMM(gemm::EPI_BIAS_F32, 256) break;
        case gemm::EPI_STORE_BF16_LN: LAUNCH_GEMM(gemm::EPI_STORE_BF16_LN, 256) break;
        case gemm::EPI_RESID_F32_LN:
            if (BN == 192) LAUNCH_GEMM(gemm::EPI_RESID_F32_LN, 192) else LAUNCH_GEMM(gemm::EPI_RESID_F32_LN, 256)
            c->stats_span = BN / 2;            // what the next LayerNorm-folded consumer will find in stats_out
            c->stats_ptr = ln.stats_out;
            break;
        case gemm::EPI_SWIGLU_BF16_LN: LAUNCH_GEMM(gemm::EPI_SWIGLU_BF16_LN, 256) break;
        case gemm::EPI_QKV_ROPE_LN: LAUNCH_GEMM(gemm::EPI_QKV_ROPE_LN, 256) break;
        default: return c->fail("gemm: unknown epilogue");
    }
#undef LAUNCH_GEMM
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_layernorm(esmdiff_ctx* c, const float* x, const float* w, const float* b, bf16* y,
                            int M, int D, cudaStream_t st) {
    if (D % 128 != 0 || D > 128 * 12) return c->fail("layernorm: D must be a multiple of 128, <= 1536");
    const int grid = (M + ew::ROWS_PER_BLOCK - 1) / ew::ROWS_PER_BLOCK;
    ProfScope prof(c, ESMDIFF_PROF_LAYERNORM, 6.0 * M * D, st);
    CK(launch_pdl(c->pdl, ew::layernorm_f32_to_bf16_kernel<12>, dim3(grid), dim3(ew::ROWS_PER_BLOCK * 32), 0, st, x, w, b, y, M, D, 1e-5f));
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int ensure_rope(esmdiff_ctx* c, int T, cudaStream_t st) {
    if (T <= c->rope_T) return 0;
    if (!c->inv_freq) {
        float h[32];
        for (int i = 0; i < 32; ++i) h[i] = 1.0f / powf(10000.0f, static_cast<float>(2 * i) / 64.0f);
        if (c->alloc(&c->inv_freq, 32)) return 1;
        CK(cudaMemcpy(c->inv_freq, h, sizeof h, cudaMemcpyHostToDevice));
    }
    int cap = 1056;
    while (cap < T) cap *= 2;
    if (c->alloc(&c->cos_t, (size_t)cap * 32)) return 1;     // old tables stay owned until destroy
    if (c->alloc(&c->sin_t, (size_t)cap * 32)) return 1;
    if (c->alloc(&c->rope_rows, (size_t)cap * 64)) return 1;
    ew::rope_table_kernel<<<(cap * 32 + 255) / 256, 256, 0, st>>>(c->inv_freq, c->cos_t, c->sin_t, cap);
    ew::rope_rows_kernel<<<(cap * 32 + 255) / 256, 256, 0, st>>>(c->inv_freq, c->rope_rows, cap);
    c->launches += 2;
    c->drop_graphs();                          // captured kernels point at the old tables
    CK(cudaGetLastError());
    c->rope_T = cap;
    return 0;
}

static int launch_qk_norm_rope(esmdiff_ctx* c, bf16* qkv, const float* qw, const float* kw, int M, int T,
                               int D, cudaStream_t st) {
    if (D % 256 != 0 || D > 256 * 6) return c->fail("qk_norm_rope: D must be a multiple of 256, <= 1536");
    if (ensure_rope(c, T, st)) return 1;
    const int grid = (M + ew::ROWS_PER_BLOCK - 1) / ew::ROWS_PER_BLOCK;
    ProfScope prof(c, ESMDIFF_PROF_QK_NORM_ROPE, 8.0 * M * D, st);
    CK(launch_pdl(c->pdl, ew::qk_layernorm_rope_kernel<6>, dim3(grid), dim3(ew::ROWS_PER_BLOCK * 32), 0, st, qkv, qw, kw,
                  c->cos_t, c->sin_t, M, D, T, 1e-5f));
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_attention_streaming(esmdiff_ctx* c, const bf16* qkv, bf16* out, int B, int T, int H,
                                      const float* qk_sumsq, cudaStream_t st) {
    const int D = H * attn::DH;
    const int64_t M = (int64_t)B * T;
    CUtensorMap tq, tkv;
    if (get_tmap(c, qkv, M, 3 * D, 3 * D, attn::BQ, &tq)) return 1;
    if (get_tmap(c, qkv, M, 3 * D, 3 * D, attn::BKV, &tkv)) return 1;
    attn::Params p;
    p.B = B; p.T = T; p.H = H;
    p.q_tiles = (T + attn::BQ - 1) / attn::BQ;
    p.ctx = out;
    p.scale_log2 = 0.125f * 1.4426950408889634f;
    p.qk_sumsq = qk_sumsq; p.nspan = D / 128; p.ln_eps = 1e-5f;
    const int smem = attn::smem_bytes(T);
    if (smem > 227 * 1024) return c->fail("attention: sequence too long for the shared-memory rstd_k table");
    CK(ensure_dynamic_smem(c->device, attn::attention_fwd_kernel, smem));
    const int grid = B * H * p.q_tiles;
    ProfScope prof(c, ESMDIFF_PROF_ATTENTION, 4.0 * B * H * (double)T * T * attn::DH, st);
    CK(launch_pdl(c->pdl, attn::attention_fwd_kernel, dim3(grid), dim3(attn::THREADS), smem, st, tq, tkv, p));
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// K/V of one (sample, head) resident in shared memory: every T whose K and V fit beside two query
// buffers (T <= 766); longer sequences stream K/V tiles (attention.cuh).
static int launch_attention(esmdiff_ctx* c, const bf16* qkv, bf16* out, int B, int T, int H, const float* qk_sumsq,
                            cudaStream_t st) {
    const int nkv = (T + attn2::BKV - 1) / attn2::BKV;
    const int tail_cols = ((T - (nkv - 1) * attn2::BKV) + 15) / 16 * 16;
    const int smem = attn2::smem_bytes(nkv, tail_cols);
    if (c->attn_variant == 1 || nkv > attn2::MAX_KV_TILES || smem > 227 * 1024)
        return launch_attention_streaming(c, qkv, out, B, T, H, qk_sumsq, st);
    const int D = H * attn2::DH;
    const int64_t M = (int64_t)B * T;
    CUtensorMap tq, tkv, tkvt;
    if (get_tmap(c, qkv, M, 3 * D, 3 * D, attn2::BQ, &tq)) return 1;
    if (get_tmap(c, qkv, M, 3 * D, 3 * D, attn2::BKV, &tkv)) return 1;
    if (get_tmap(c, qkv, M, 3 * D, 3 * D, tail_cols, &tkvt)) return 1;
    attn2::Params p;
    p.B = B; p.T = T; p.H = H;
    p.nq = (T + attn2::BQ - 1) / attn2::BQ;
    p.n_left = 0;
    if (T > attn2::BQ && T % attn2::BQ != 0 && T % attn2::BQ <= attn2::MAX_LEFT && c->attn_variant != 2) {
        p.n_left = T % attn2::BQ;                  // T = 128 k + 2 (BOS/EOS): no tensor tile for two rows
        p.nq = T / attn2::BQ;
    }
    p.qkv = qkv;
    p.nkv = nkv;
    p.tail_cols = tail_cols;
    p.ctx = out;
    p.scale_log2 = 0.125f * 1.4426950408889634f;
    p.qk_sumsq = qk_sumsq; p.nspan = D / 128; p.ln_eps = 1e-5f;
    // T = 64 k + e, e <= 4: the e tail keys ride on the last full step (attention_resident.cuh "Folded tail")
    p.fold = (nkv >= 2 && T - (nkv - 1) * attn2::BKV <= attn2::FOLD_MAX && c->attn_fold) ? 1 : 0;
    CK(ensure_dynamic_smem(c->device, attn2::attention_resident_kernel, smem));
    // query-range split (attention_resident.cuh Params::q_splits): two half-length CTAs per (sample, head)
    // when that needs fewer rounds of the 2 x num_sms resident CTAs (each half reloads K/V: ~10 % extra)
    p.q_splits = 1;
    if (p.nq >= 2 && c->attn_qsplit != 1) {
        const double slots = 2.0 * c->num_sms, items = (double)B * H;
        const double r1 = ceil(items / slots), r2 = ceil(2.0 * items / slots) * 0.55;
        if (c->attn_qsplit == 2 || r2 < r1 * 0.97) p.q_splits = 2;
    }
    ProfScope prof(c, ESMDIFF_PROF_ATTENTION, 4.0 * B * H * (double)T * T * attn2::DH, st);
    CK(launch_pdl(c->pdl, attn2::attention_resident_kernel, dim3(B * H * p.q_splits), dim3(attn2::THREADS), smem, st, tq,
                  tkv, tkvt, p));
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_time_embed(esmdiff_ctx* c, float sigma, float* cond, cudaStream_t st) {
    const int D = c->cfg.d_model;
    if (!c->cfg.time_conditioning) sigma = 0.f;           // model.py:538-539
    CK(launch_pdl(c->pdl, ew::time_embed_hidden_kernel, dim3((D + 7) / 8), dim3(256), 0, st, sigma, c->te_w0, c->te_b0,
                  c->te_hidden, D, c->cfg.time_freq_dim));
    CK(launch_pdl(c->pdl, ew::time_embed_out_kernel, dim3((D + 7) / 8), dim3(256), 0, st, c->te_hidden, c->te_w2, c->te_b2,
                  cond, D));
    c->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

static int ensure_workspace(esmdiff_ctx* c, int64_t M) {
    if (M <= c->ws_rows) return 0;
    // free the previous workspace buffers
    void* olds[] = {c->x, c->headh, c->logits_ws, c->xn, c->qkv, c->att, c->hbuf, c->stats, c->qk_sumsq, c->aux_ws};
    for (void* o : olds)
        if (o) {
            cudaFree(o);
            for (auto& q : c->owned)
                if (q == o) q = nullptr;
        }
    c->tmaps.clear();
    c->drop_graphs();                          // captured kernels point into the old workspace
    const int64_t D = c->cfg.d_model, F = c->cfg.ffn_hidden, V = c->cfg.n_structure_heads;
    c->x = nullptr; c->headh = nullptr; c->logits_ws = nullptr;
    c->xn = nullptr; c->qkv = nullptr; c->att = nullptr; c->hbuf = nullptr; c->stats = nullptr; c->qk_sumsq = nullptr;
    if (c->alloc(&c->stats, M * ((D + 95) / 96))) return 1;      // per-row partials: D/128 or D/96 spans
    if (c->alloc(&c->qk_sumsq, M * 2 * (D / 128))) return 1;
    c->aux_ws = nullptr;
    if (c->cfg.n_aux_out > 0 && c->alloc(&c->aux_ws, M * c->cfg.n_aux_out)) return 1;
    if (c->alloc(&c->x, M * D)) return 1;
    if (c->alloc(&c->headh, M * D)) return 1;
    if (c->alloc(&c->logits_ws, M * V)) return 1;
    if (c->alloc(&c->xn, M * D)) return 1;
    if (c->alloc(&c->qkv, M * 3 * D)) return 1;
    if (c->alloc(&c->att, M * D)) return 1;
    if (c->alloc(&c->hbuf, M * F)) return 1;
    c->ws_rows = M;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// transformer blocks (esm UnifiedTransformerBlock without geometric attention), shared by the
// sampling network (48 x d=1536, residual / sqrt(48/36)) and the structure token decoder
// (30 x d=1280, scale_residue=False).  In: c->x (fp32 stream) and, with the LayerNorms folded,
// c->xn (bf16 copy) + c->stats.  Out: c->x.
// ------------------------------------------------------------------------------------------------
// Block 0's geometric attention with live frames (esmdiff_set_structure_coords): s_norm (weight-only LayerNorm)
// -> proj on the tensor cores (bf16 [M, 15 v_heads]) -> rotate into the global frame + attention over the T keys
// of every sample in fp32 on the CUDA cores (geom.cuh; 3-vectors, translations of tens of Angstrom: not bf16
// work) -> out_proj with the residual epilogue.  c->att serves as the LayerNorm output and then as the
// attention output (ld = 3 v_heads).
static int run_geom_attention(esmdiff_ctx* c, const LayerW& w, int B, int T, float rs, int resid_epi, const GemmLN& make,
                              cudaStream_t st) {
    const int M = B * T, D = c->cfg.d_model, H = c->cfg.v_heads;
    if (B != c->geom_B || T != c->geom_T)
        return c->fail("forward: the batch shape differs from the one esmdiff_set_structure_coords was given");
    if (!w.g_snorm || !w.g_proj || !w.g_out || !w.g_wdist || !w.g_wrot)
        return c->fail("forward: structure coordinates are set but the geom_attn.* weights of block 0 were never loaded");
    if (launch_layernorm(c, c->x, w.g_snorm, nullptr, c->att, M, D, st)) return 1;
    if (launch_gemm(c, gemm::EPI_STORE_BF16, c->att, w.g_proj, M, 15 * H, D, c->g_projout, 15 * H, nullptr, 1.f, st)) return 1;
    if (esmdiff_geom_attention_bf16(c->g_projout, c->g_work, c->g_rot, c->g_trans, c->g_mask, w.g_wrot, w.g_wdist, c->att,
                                    3 * H, B, T, H, st))
        return c->fail("forward: geometric attention launch failed");
    c->launches += 2;
    return launch_gemm(c, resid_epi, c->att, w.g_out, M, D, 3 * H, c->x, D, nullptr, rs, st, make);
}

static int run_blocks(esmdiff_ctx* c, int B, int T, cudaStream_t st) {
    const int M = B * T;
    const int D = c->cfg.d_model, F = c->cfg.ffn_hidden, H = c->cfg.n_heads;
    const float rs = c->cfg.model_kind == 1 ? 1.0f : sqrtf((float)c->cfg.n_layers / 36.0f);
    for (int l = 0; l < c->cfg.n_layers; ++l) {
        const LayerW& w = c->layers[l];
        if (c->ln_fold) {
            // pre-LayerNorms folded through the GEMMs: xn is the bf16 copy of the RAW stream and
            // stats its per-row partial statistics, both written by the producer of x (embedding
            // kernel / residual epilogue); gamma is folded into wqkv / w1 at finalize
            GemmLN use, make;
            use.stats_in = c->stats;
            make.stats_out = c->stats;
            make.xb_out = c->xn;
            const bool last = l == c->cfg.n_layers - 1;           // the final norm + head use the LN kernel
            use.colsum = w.cqkv;
            use.stats_span = c->stats_span;                       // as left by the producer of x (embedding: 128)
            if (c->qk_fused) {
                // q_ln / k_ln + RoPE inside the QKV epilogue; 1/std of the q and k rows inside attention
                GemmLN qk = use;
                qk.rope = c->rope_rows; qk.qk_gamma = w.qk_gamma; qk.qk_sumsq = c->qk_sumsq; qk.T = T; qk.n_rope = 2 * D;
                if (launch_gemm(c, gemm::EPI_QKV_ROPE_LN, c->xn, w.wqkv, M, 3 * D, D, c->qkv, 3 * D, w.bqkv, 1.f, st, qk)) return 1;
                if (launch_attention(c, c->qkv, c->att, B, T, H, c->qk_sumsq, st)) return 1;
            } else {
                if (launch_gemm(c, gemm::EPI_STORE_BF16_LN, c->xn, w.wqkv, M, 3 * D, D, c->qkv, 3 * D, w.bqkv, 1.f, st, use)) return 1;
                if (launch_qk_norm_rope(c, c->qkv, w.qln_w, w.kln_w, M, T, D, st)) return 1;
                if (launch_attention(c, c->qkv, c->att, B, T, H, nullptr, st)) return 1;
            }
            if (launch_gemm(c, gemm::EPI_RESID_F32_LN, c->att, w.wo, M, D, D, c->x, D, nullptr, rs, st, make)) return 1;
            // block 0's geometric attention contributes exactly 0 without coordinates
            if (l == 0 && c->geom_on && run_geom_attention(c, w, B, T, rs, gemm::EPI_RESID_F32_LN, make, st)) return 1;
            use.colsum = w.c1;
            use.stats_span = c->stats_span;
            if (launch_gemm(c, gemm::EPI_SWIGLU_BF16_LN, c->xn, w.w1, M, 2 * F, D, c->hbuf, F, w.b1, 1.f, st, use)) return 1;
            if (launch_gemm(c, last ? gemm::EPI_RESID_F32 : gemm::EPI_RESID_F32_LN, c->hbuf, w.w2, M, D, F, c->x, D,
                            nullptr, rs, st, make)) return 1;
            continue;
        }
        if (launch_layernorm(c, c->x, w.ln1_w, w.ln1_b, c->xn, M, D, st)) return 1;
        if (launch_gemm(c, gemm::EPI_STORE_BF16, c->xn, w.wqkv, M, 3 * D, D, c->qkv, 3 * D, nullptr, 1.f, st)) return 1;
        if (launch_qk_norm_rope(c, c->qkv, w.qln_w, w.kln_w, M, T, D, st)) return 1;
        if (launch_attention(c, c->qkv, c->att, B, T, H, nullptr, st)) return 1;
        if (launch_gemm(c, gemm::EPI_RESID_F32, c->att, w.wo, M, D, D, c->x, D, nullptr, rs, st)) return 1;
        // block 0's geometric attention contributes exactly 0 without coordinates
        if (l == 0 && c->geom_on && run_geom_attention(c, w, B, T, rs, gemm::EPI_RESID_F32, GemmLN(), st)) return 1;
        if (launch_layernorm(c, c->x, w.ln2_w, w.ln2_b, c->xn, M, D, st)) return 1;
        if (launch_gemm(c, gemm::EPI_SWIGLU_BF16, c->xn, w.w1, M, 2 * F, D, c->hbuf, F, nullptr, 1.f, st)) return 1;
        if (launch_gemm(c, gemm::EPI_RESID_F32, c->hbuf, w.w2, M, D, F, c->x, D, nullptr, rs, st)) return 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
static int forward_impl(esmdiff_ctx* c, const int64_t* seq, const int64_t* xt, int B, int T,
                        const float* aux, int64_t aux_stride, float* logits, float* emb_out,
                        cudaStream_t st) {
    if (!c->finalized) return c->fail("forward: esmdiff_finalize_weights has not succeeded");
    if (B <= 0 || T <= 0) return c->fail("forward: B and T must be positive");
    const int64_t M64 = (int64_t)B * T;
    if (M64 > (1ll << 30)) return c->fail("forward: B*T too large");
    const int M = (int)M64;
    const int D = c->cfg.d_model, V = c->cfg.n_structure_heads;
    if (ensure_workspace(c, M)) return 1;
    if (c->qk_fused && ensure_rope(c, T, st)) return 1;
    if (c->cfg.model_kind != 0) return c->fail("forward: this context holds a structure token decoder (esmdiff_decode_structure)");

    const int rgrid = (M + ew::ROWS_PER_BLOCK - 1) / ew::ROWS_PER_BLOCK;
    {
    ProfScope prof(c, ESMDIFF_PROF_EMBED, 12.0 * M * D, st);
    ew::embed_kernel<<<rgrid, 256, 0, st>>>(reinterpret_cast<const long long*>(seq),
                                            reinterpret_cast<const long long*>(xt), c->seq_embed,
                                            c->struct_embed, c->const_vec, aux, aux_stride, c->x, M, D,
                                            c->cfg.seq_vocab, c->cfg.struct_vocab, c->dev_err,
                                            c->ln_fold ? c->xn : nullptr, c->stats);
    c->stats_span = 128;
    c->stats_ptr = c->stats;
    }
    c->launches++;
    CK(cudaGetLastError());

    if (run_blocks(c, B, T, st)) return 1;
    if (emb_out) CK(cudaMemcpyAsync(emb_out, c->x, (size_t)M * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (launch_layernorm(c, c->x, c->norm_w, nullptr, c->xn, M, D, st)) return 1;
    if (launch_gemm(c, gemm::EPI_BIAS_GELU_F32, c->xn, c->h0_w, M, D, D, c->headh, D, c->h0_b, 1.f, st)) return 1;
    if (launch_layernorm(c, c->headh, c->h2_w, c->h2_b, c->xn, M, D, st)) return 1;
    if (launch_gemm(c, gemm::EPI_BIAS_F32, c->xn, c->h3_w, M, V, D, logits, V, c->h3_b, 1.f, st)) return 1;
    return 0;
}

// One forward of the sampling loop, replayed from a CUDA graph when the batch is small enough to be
// launch bound.  The first call for a key runs eagerly (allocations, function attributes, TMA
// descriptors), the second captures, later ones replay.  Falls back to eager launches whenever
// capture is not possible (profiling on, capture failure).
static int forward_step(esmdiff_ctx* c, const int64_t* seq, const int64_t* xt, int B, int T, float* logits,
                        cudaStream_t st) {
    const int64_t M = (int64_t)B * T;
    const bool want = !c->prof && (c->graph_mode == 1 || (c->graph_mode == -1 && M <= c->graph_max_rows));
    if (!want) return forward_impl(c, seq, xt, B, T, c->cond, 0, logits, nullptr, st);
    // the library's own (non-default, non-blocking) streams can be captured directly; a caller's stream
    // may be the legacy default stream, where capture is illegal -> fork to gstream
    const bool own = st != nullptr && (st == c->sstream[0] || st == c->sstream[1]);
    auto key = std::make_tuple(B, T, (const void*)seq, (const void*)xt, c->active_ws);
    if (c->graphs.size() > 64 && !c->graphs.count(key)) c->drop_graphs();     // callers that never reuse buffers
    esmdiff_ctx::GraphRec& g = c->graphs[key];
    if (logits != c->logits_ws) return forward_impl(c, seq, xt, B, T, c->cond, 0, logits, nullptr, st);
    if (!g.exec) {
        if (!g.warmed) {
            g.warmed = true;
            return forward_impl(c, seq, xt, B, T, c->cond, 0, logits, nullptr, st);
        }
        if (!own && !c->gstream) {
            CK(cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        }
        const int64_t l0 = c->launches;
        cudaGraph_t graph = nullptr;
        cudaStream_t cs = own ? st : c->gstream;
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        const int rc = forward_impl(c, seq, xt, B, T, c->cond, 0, logits, nullptr, cs);
        const cudaError_t e = cudaStreamEndCapture(cs, &graph);
        g.kernels = c->launches - l0;
        c->launches = l0;                                  // nothing ran yet
        if (rc != 0 || e != cudaSuccess || graph == nullptr) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            c->graph_mode = 0;                             // do not try again
            return forward_impl(c, seq, xt, B, T, c->cond, 0, logits, nullptr, st);
        }
        const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            g.exec = nullptr;
            cudaGetLastError();
            c->graph_mode = 0;
            return forward_impl(c, seq, xt, B, T, c->cond, 0, logits, nullptr, st);
        }
    }
    if (own) {
        CK(cudaGraphLaunch(g.exec, st));
    } else {
        CK(cudaEventRecord(c->ev_fork, st));
        CK(cudaStreamWaitEvent(c->gstream, c->ev_fork, 0));
        CK(cudaGraphLaunch(g.exec, c->gstream));
        CK(cudaEventRecord(c->ev_join, c->gstream));
        CK(cudaStreamWaitEvent(st, c->ev_join, 0));
    }
    c->launches += g.kernels;
    return 0;
}

template <int MODE>
static int launch_sampler(esmdiff_ctx* c, const float* logits, const float* u, int64_t* x, float* logp,
                          int M, float mc_t, float mc_s, uint64_t seed, uint32_t step, cudaStream_t st,
                          uint32_t row_offset = 0) {
    const int V = c->cfg.n_structure_heads;
    if (V > sampler::THREADS * sampler::MAX_PER_THREAD) return c->fail("sampler: vocabulary too large");
    ProfScope prof(c, ESMDIFF_PROF_SAMPLER, (u ? 8.0 : 4.0) * M * V, st);
    sampler::sample_rows_kernel<MODE><<<M, sampler::THREADS, 0, st>>>(
        logits, (long long)V, u, reinterpret_cast<long long*>(x), logp, M, V,
        ESMDIFF_STRUCTURE_MASK_TOKEN, mc_t, mc_s, (unsigned long long)seed, step, row_offset);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void fill_i64_kernel(long long* p, long long v, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// number of entries equal to `v` (is a MASK left before the noise-removal forward?)
__global__ void count_equal_kernel(const long long* p, long long v, long long n, int* count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = i < n && p[i] == v;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}
__global__ void bf16_to_f32_kernel(const bf16* s, float* d, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __bfloat162float(s[i]);
}
enum Kind { K_F32, K_BF16, K_BF16_SWIGLU, K_SKIP };
struct Slot {
    Kind kind;
    void** dst;
    std::vector<int64_t> shape;
    float** master = nullptr;      // LayerNorm-folded Linears: the fp32 copy is kept until finalize
    int layer = -1;                // block whose folded weights this key invalidates
};
}  // namespace

static bool resolve_key(esmdiff_ctx* c, const std::string& key, Slot* s) {
    const int64_t D = c->cfg.d_model, F = c->cfg.ffn_hidden, V = c->cfg.n_structure_heads;
    auto f32 = [&](float** p, std::vector<int64_t> shp) { *s = {K_F32, (void**)p, shp}; return true; };
    auto b16 = [&](bf16** p, std::vector<int64_t> shp) { *s = {K_BF16, (void**)p, shp}; return true; };
    auto skip = [&]() { *s = {K_SKIP, nullptr, {}}; return true; };
    std::string k;
    if (c->cfg.model_kind == 1) {
        // StructureTokenDecoder.state_dict() names
        const int64_t A = c->cfg.n_aux_out;
        if (key == "embed.weight") return f32(&c->dec_embed, {c->cfg.struct_vocab, D});
        if (key == "decoder_stack.norm.weight") return f32(&c->norm_w, {D});
        if (key.rfind("affine_output_projection.", 0) == 0) {
            const std::string r = key.substr(25);
            if (r == "ffn1.weight") return b16(&c->h0_w, {D, D});
            if (r == "ffn1.bias") return f32(&c->h0_b, {D});
            if (r == "norm.weight") return f32(&c->h2_w, {D});
            if (r == "norm.bias") return f32(&c->h2_b, {D});
            if (r == "proj.weight") return b16(&c->h3_w, {V, D});
            if (r == "proj.bias") return f32(&c->h3_b, {V});
            return false;
        }
        if (key.rfind("plddt_head.", 0) == 0) {
            if (A == 0) return skip();
            const std::string r = key.substr(11);
            if (r == "0.weight") return b16(&c->p0_w, {D, D});
            if (r == "0.bias") return f32(&c->p0_b, {D});
            if (r == "2.weight") return f32(&c->p2_w, {D});
            if (r == "2.bias") return f32(&c->p2_b, {D});
            if (r == "3.weight") return b16(&c->p3_w, {A, D});
            if (r == "3.bias") return f32(&c->p3_b, {A});
            return false;
        }
        if (key.rfind("pairwise_classification_head.", 0) == 0) return skip();     // pTM / PAE: not on this path
        if (key.rfind("decoder_stack.", 0) != 0) return false;
        k = "transformer." + key.substr(14);
    } else
    if (key.rfind("sigma_embedder.mlp.", 0) == 0) {
        const std::string r = key.substr(19);
        if (r == "0.weight") return f32(&c->te_w0, {D, c->cfg.time_freq_dim});
        if (r == "0.bias") return f32(&c->te_b0, {D});
        if (r == "2.weight") return f32(&c->te_w2, {D, D});
        if (r == "2.bias") return f32(&c->te_b2, {D});
        return false;
    }
    if (c->cfg.model_kind == 0) {
        if (key.rfind("net.", 0) != 0) return false;
        k = key.substr(4);
    }
    if (k == "encoder.sequence_embed.weight") return f32(&c->seq_embed, {c->cfg.seq_vocab, D});
    if (k == "encoder.structure_tokens_embed.weight") return f32(&c->struct_embed, {c->cfg.struct_vocab, D});
    if (k == "encoder.plddt_projection.weight") return f32(&c->plddt_w, {D, 16});
    if (k == "encoder.plddt_projection.bias") return f32(&c->plddt_b, {D});
    if (k == "encoder.structure_per_res_plddt_projection.weight") return f32(&c->res_w, {D, 16});
    if (k == "encoder.structure_per_res_plddt_projection.bias") return f32(&c->res_b, {D});
    if (k == "encoder.ss8_embed.weight") return f32(&c->ss8, {11, D});
    if (k == "encoder.sasa_embed.weight") return f32(&c->sasa, {19, D});
    if (k.rfind("encoder.function_embed.", 0) == 0 || k == "encoder.residue_embed.weight") return skip();
    if (k == "transformer.norm.weight") return f32(&c->norm_w, {D});
    if (k.rfind("output_heads.structure_head.", 0) == 0) {
        const std::string r = k.substr(28);
        if (r == "0.weight") return b16(&c->h0_w, {D, D});
        if (r == "0.bias") return f32(&c->h0_b, {D});
        if (r == "2.weight") return f32(&c->h2_w, {D});
        if (r == "2.bias") return f32(&c->h2_b, {D});
        if (r == "3.weight") return b16(&c->h3_w, {V, D});
        if (r == "3.bias") return f32(&c->h3_b, {V});
        return false;
    }
    if (k.rfind("output_heads.", 0) == 0) return skip();      // other ESM3 heads: not on this path
    if (k.rfind("transformer.blocks.", 0) == 0) {
        const size_t dot = k.find('.', 19);
        if (dot == std::string::npos) return false;
        const int l = atoi(k.substr(19, dot - 19).c_str());
        if (l < 0 || l >= c->cfg.n_layers) return false;
        LayerW& w = c->layers[l];
        const std::string r = k.substr(dot + 1);
        if (r.rfind("geom_attn.", 0) == 0) {
            // exact zero without coordinates (A6); kept for esmdiff_set_structure_coords when v_heads is known
            const int64_t H = c->cfg.v_heads;
            if (H <= 0 || l != 0 || c->cfg.model_kind != 0) return skip();
            if (r == "geom_attn.s_norm.weight") return f32(&w.g_snorm, {D});
            if (r == "geom_attn.proj.weight") return b16(&w.g_proj, {15 * H, D});
            if (r == "geom_attn.out_proj.weight") return b16(&w.g_out, {D, 3 * H});
            if (r == "geom_attn.distance_scale_per_head") return f32(&w.g_wdist, {H});
            if (r == "geom_attn.rotation_scale_per_head") return f32(&w.g_wrot, {H});
            return false;
        }
        if (r == "attn.layernorm_qkv.0.weight") { f32(&w.ln1_w, {D}); s->layer = l; return true; }
        if (r == "attn.layernorm_qkv.0.bias") { f32(&w.ln1_b, {D}); s->layer = l; return true; }
        if (r == "attn.layernorm_qkv.1.weight") {
            b16(&w.wqkv, {3 * D, D});
            s->layer = l;
            if (c->ln_fold) s->master = &w.wqkv_f32;
            return true;
        }
        if (r == "attn.q_ln.weight") return f32(&w.qln_w, {D});
        if (r == "attn.k_ln.weight") return f32(&w.kln_w, {D});
        if (r == "attn.out_proj.weight") return b16(&w.wo, {D, D});
        if (r == "ffn.0.weight") { f32(&w.ln2_w, {D}); s->layer = l; return true; }
        if (r == "ffn.0.bias") { f32(&w.ln2_b, {D}); s->layer = l; return true; }
        if (r == "ffn.1.weight") {
            *s = {K_BF16_SWIGLU, (void**)&w.w1, {2 * F, D}};
            s->layer = l;
            if (c->ln_fold) s->master = &w.w1_f32;
            return true;
        }
        if (r == "ffn.3.weight") return b16(&w.w2, {D, F});
        return false;
    }
    return false;
}

static std::vector<std::string> required_keys(const esmdiff_ctx* c) {
    const char* per_block[] = {"attn.layernorm_qkv.0.weight", "attn.layernorm_qkv.0.bias",
                               "attn.layernorm_qkv.1.weight", "attn.q_ln.weight", "attn.k_ln.weight",
                               "attn.out_proj.weight", "ffn.0.weight", "ffn.0.bias", "ffn.1.weight",
                               "ffn.3.weight"};
    if (c->cfg.model_kind == 1) {
        std::vector<std::string> k = {"embed.weight", "decoder_stack.norm.weight",
                                      "affine_output_projection.ffn1.weight", "affine_output_projection.ffn1.bias",
                                      "affine_output_projection.norm.weight", "affine_output_projection.norm.bias",
                                      "affine_output_projection.proj.weight", "affine_output_projection.proj.bias"};
        if (c->cfg.n_aux_out > 0)
            for (const char* r : {"0.weight", "0.bias", "2.weight", "2.bias", "3.weight", "3.bias"})
                k.push_back(std::string("plddt_head.") + r);
        for (int l = 0; l < c->cfg.n_layers; ++l)
            for (const char* r : per_block) k.push_back("decoder_stack.blocks." + std::to_string(l) + "." + r);
        return k;
    }
    std::vector<std::string> k = {
        "net.encoder.sequence_embed.weight", "net.encoder.structure_tokens_embed.weight",
        "net.encoder.plddt_projection.weight", "net.encoder.plddt_projection.bias",
        "net.encoder.structure_per_res_plddt_projection.weight",
        "net.encoder.structure_per_res_plddt_projection.bias", "net.encoder.ss8_embed.weight",
        "net.encoder.sasa_embed.weight", "net.transformer.norm.weight",
        "net.output_heads.structure_head.0.weight", "net.output_heads.structure_head.0.bias",
        "net.output_heads.structure_head.2.weight", "net.output_heads.structure_head.2.bias",
        "net.output_heads.structure_head.3.weight", "net.output_heads.structure_head.3.bias",
        "sigma_embedder.mlp.0.weight", "sigma_embedder.mlp.0.bias", "sigma_embedder.mlp.2.weight",
        "sigma_embedder.mlp.2.bias"};
    const char* per[] = {"attn.layernorm_qkv.0.weight", "attn.layernorm_qkv.0.bias",
                         "attn.layernorm_qkv.1.weight", "attn.q_ln.weight", "attn.k_ln.weight",
                         "attn.out_proj.weight", "ffn.0.weight", "ffn.0.bias", "ffn.1.weight",
                         "ffn.3.weight"};
    for (int l = 0; l < c->cfg.n_layers; ++l)
        for (const char* r : per) k.push_back("net.transformer.blocks." + std::to_string(l) + "." + r);
    return k;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int esmdiff_abi_version(void) { return ESMDIFF_ABI_VERSION; }

const char* esmdiff_last_error(const esmdiff_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int esmdiff_create(const esmdiff_cfg* cfg, int device, esmdiff_ctx** out) {
    if (!cfg || !out) { g_create_error = "create: null argument"; return 1; }
    *out = nullptr;
    if (cfg->d_model <= 0 || cfg->d_model % 256 != 0 || cfg->d_model > 1536 ||
        cfg->n_heads * 64 != cfg->d_model || cfg->n_layers <= 0 || cfg->ffn_hidden % 128 != 0 ||
        cfg->ffn_hidden <= 0 || cfg->model_kind < 0 || cfg->model_kind > 1 ||
        (cfg->model_kind == 0 && (cfg->n_structure_heads <= ESMDIFF_STRUCTURE_MASK_TOKEN ||
                                  cfg->n_structure_heads > 4352 || cfg->time_freq_dim <= 0 || cfg->time_freq_dim % 2 != 0)) ||
        (cfg->model_kind == 1 && (cfg->n_structure_heads < 9 || cfg->n_structure_heads > 4352 || cfg->n_aux_out < 0 ||
                                  cfg->n_aux_out > 4352 || cfg->struct_vocab <= 0))) {
        g_create_error = "create: unsupported dimensions (need d_model % 256 == 0 <= 1536, d_head 64, "
                         "ffn_hidden % 128 == 0, 4096 < n_structure_heads <= 4352)";
        return 1;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("create: no CUDA device (") + cudaGetErrorString(e) +
                         "); esmdiff_b200 has no CPU fallback";
        return 2;
    }
    if (device < 0 || device >= ndev) { g_create_error = "create: bad device index"; return 1; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        g_create_error = "create: device is not sm_100 (Blackwell B200); kernels are sm_100a only";
        return 2;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return 1; }
    esmdiff_ctx* c = new esmdiff_ctx();
    c->cfg = *cfg;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->layers.resize(cfg->n_layers);
    if (const char* e = getenv("ESMDIFF_ATTN"))
        c->attn_variant = strcmp(e, "stream") == 0 ? 1 : strcmp(e, "tiles") == 0 ? 2 : 0;
    if (const char* e = getenv("ESMDIFF_LN")) c->ln_fold = strcmp(e, "separate") != 0;
    if (const char* e = getenv("ESMDIFF_QK")) c->qk_fused = strcmp(e, "separate") != 0;
    if (const char* e = getenv("ESMDIFF_PDL")) c->pdl = atoi(e) != 0;
    if (const char* e = getenv("ESMDIFF_SKIP_DENOISE_FORWARD")) c->skip_denoise = atoi(e) != 0;
    if (const char* e = getenv("ESMDIFF_QKV_RUN")) c->qkv_run = atoi(e);
    if (const char* e = getenv("ESMDIFF_SPLIT_ROWS")) c->split_rows = atoll(e);
    if (const char* e = getenv("ESMDIFF_RESID_BN")) c->resid_bn = atoi(e);
    if (const char* e = getenv("ESMDIFF_ATTN_QSPLIT")) c->attn_qsplit = atoi(e);
    if (const char* e = getenv("ESMDIFF_ATTN_FOLD")) c->attn_fold = atoi(e);
    c->qk_fused = c->qk_fused && c->ln_fold;
    if (const char* e = getenv("ESMDIFF_GRAPH")) c->graph_mode = atoi(e) != 0 ? 1 : 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) {
        g_create_error = "create: cuTensorMapEncodeTiled not available from the driver";
        delete c;
        return 1;
    }
    c->encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (c->alloc(&c->dev_err, 1) || c->alloc(&c->cond, cfg->d_model) || c->alloc(&c->te_hidden, cfg->d_model) ||
        c->alloc(&c->const_vec, cfg->d_model)) {
        g_create_error = c->err;
        delete c;
        return 1;
    }
    cudaMemset(c->dev_err, 0, sizeof(int));
    int zero = 0;
    cudaMemcpyToSymbol(g_abort_flag, &zero, sizeof(int));
    *out = c;
    return 0;
}

int esmdiff_destroy(esmdiff_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->drop_graphs();
    if (c->gstream) cudaStreamDestroy(c->gstream);
    for (int i = 0; i < 2; ++i) {
        if (c->sstream[i]) cudaStreamDestroy(c->sstream[i]);
        if (c->ev_sjoin[i]) cudaEventDestroy(c->ev_sjoin[i]);
    }
    if (c->ev_sfork) cudaEventDestroy(c->ev_sfork);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (void* p : c->owned)
        if (p) cudaFree(p);
    for (LayerW& w : c->layers) {
        if (w.wqkv_f32) cudaFree(w.wqkv_f32);
        if (w.w1_f32) cudaFree(w.w1_f32);
    }
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    delete c;
    return 0;
}

int esmdiff_set_weight(esmdiff_ctx* c, const char* key, const void* data, int on_device, int dtype,
                       const int64_t* shape, int ndim) {
    if (!c || !key || !data) return 1;
    CK(cudaSetDevice(c->device));
    Slot s;
    if (!resolve_key(c, key, &s)) return c->fail(std::string("set_weight: unexpected key ") + key);
    if (s.kind == K_SKIP) return 0;
    if ((int)s.shape.size() != ndim) return c->fail(std::string("set_weight: rank mismatch for ") + key);
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] != s.shape[i]) {
            char buf[256];
            snprintf(buf, sizeof buf, "set_weight: size mismatch for %s: dim %d is %lld, expected %lld", key, i,
                     (long long)shape[i], (long long)s.shape[i]);
            return c->fail(buf);
        }
        n *= shape[i];
    }
    if (dtype != ESMDIFF_F32 && dtype != ESMDIFF_BF16) return c->fail("set_weight: dtype must be f32 or bf16");
    // stage as fp32 on the device
    float* stage = nullptr;
    CK(cudaMalloc(&stage, n * sizeof(float)));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (dtype == ESMDIFF_F32) {
        CK(cudaMemcpy(stage, data, n * sizeof(float), kind));
    } else {
        bf16* tmp = nullptr;
        CK(cudaMalloc(&tmp, n * sizeof(bf16)));
        CK(cudaMemcpy(tmp, data, n * sizeof(bf16), kind));
        bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256>>>(tmp, stage, n);
        CK(cudaDeviceSynchronize());
        cudaFree(tmp);
    }
    if (*s.dst == nullptr) {
        void* q = nullptr;
        const size_t bytes = n * (s.kind == K_F32 ? sizeof(float) : sizeof(bf16));
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) { cudaFree(stage); return c->fail("set_weight: out of device memory"); }
        c->owned.push_back(q);
        *s.dst = q;
    }
    if (s.layer >= 0) c->layers[s.layer].dirty = true;
    if (s.master != nullptr) {
        // folded at finalize (needs this block's LayerNorm weight and bias as well)
        if (*s.master) cudaFree(*s.master);
        *s.master = stage;
        c->loaded.insert(key);
        c->finalized = false;
        return 0;
    }
    if (s.kind == K_F32) {
        CK(cudaMemcpy(*s.dst, stage, n * sizeof(float), cudaMemcpyDeviceToDevice));
    } else {
        const int64_t rows = s.shape[0], cols = s.shape[1];
        ew::convert_rows_bf16_kernel<<<(unsigned)((n + 255) / 256), 256>>>(
            stage, reinterpret_cast<bf16*>(*s.dst), rows, cols, s.kind == K_BF16_SWIGLU ? c->cfg.ffn_hidden : 0);
        CK(cudaGetLastError());
    }
    CK(cudaDeviceSynchronize());
    cudaFree(stage);
    c->loaded.insert(key);
    c->finalized = false;
    return 0;
}

int esmdiff_finalize_weights(esmdiff_ctx* c) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    std::string missing;
    int nmiss = 0;
    for (const std::string& k : required_keys(c))
        if (!c->loaded.count(k)) {
            if (nmiss < 8) missing += (nmiss ? ", " : "") + k;
            ++nmiss;
        }
    if (nmiss) {
        return c->fail("finalize_weights: Missing key(s) in state_dict: " + missing +
                       (nmiss > 8 ? " ... (" + std::to_string(nmiss) + " total)" : ""));
    }
    const int D = c->cfg.d_model;
    if (c->ln_fold) {
        const int64_t F = c->cfg.ffn_hidden;
        for (int l = 0; l < c->cfg.n_layers; ++l) {
            LayerW& w = c->layers[l];
            if (!w.dirty) continue;
            if (!w.wqkv_f32 || !w.w1_f32)
                return c->fail("finalize_weights: block " + std::to_string(l) + " had a LayerNorm or Linear weight "
                               "replaced after finalize; set attn.layernorm_qkv.{0,1} and ffn.{0,1} of that block together");
            if (!w.cqkv) {
                if (c->alloc(&w.cqkv, 3 * D) || c->alloc(&w.bqkv, 3 * D) || c->alloc(&w.c1, 2 * F) || c->alloc(&w.b1, 2 * F) ||
                    c->alloc(&w.qk_gamma, 2 * D))
                    return 1;
            }
            const float* colmean = nullptr;
            if (c->qk_fused) {
                // q_ln / k_ln centring folded into the weight: remove the column means of the q rows
                // and of the k rows (gemm.cuh EPI_QKV_ROPE_LN)
                if (!c->colmean && c->alloc(&c->colmean, 2 * D)) return 1;
                ew::column_mean_kernel<<<dim3((unsigned)((D + 255) / 256), 2), 256>>>(w.wqkv_f32, c->colmean, D, D);
                colmean = c->colmean;
            }
            ew::fold_layernorm_weight_kernel<<<(unsigned)((3 * D + 7) / 8), 256>>>(w.wqkv_f32, w.ln1_w, w.ln1_b, w.wqkv,
                                                                                   w.cqkv, w.bqkv, 3 * D, D, 0, colmean,
                                                                                   c->qk_fused ? 2 * D : 0, D);
            ew::fold_layernorm_weight_kernel<<<(unsigned)((2 * F + 7) / 8), 256>>>(w.w1_f32, w.ln2_w, w.ln2_b, w.w1, w.c1,
                                                                                   w.b1, 2 * F, D, (int)F);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            cudaFree(w.wqkv_f32);
            cudaFree(w.w1_f32);
            w.wqkv_f32 = nullptr;
            w.w1_f32 = nullptr;
            w.dirty = false;
        }
    }
    if (c->qk_fused)
        for (LayerW& w : c->layers) {            // q_ln / k_ln weights can be replaced without refolding
            ew::interleave_rotary_pairs_kernel<<<(D + 255) / 256, 256>>>(w.qln_w, w.qk_gamma, D);
            ew::interleave_rotary_pairs_kernel<<<(D + 255) / 256, 256>>>(w.kln_w, w.qk_gamma + D, D);
        }
    if (c->cfg.model_kind == 0)
        ew::default_tracks_kernel<<<(D + 127) / 128, 128>>>(c->plddt_w, c->plddt_b, c->res_w, c->res_b, c->ss8,
                                                            c->sasa, c->const_vec, D);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    c->finalized = true;
    return 0;
}

int esmdiff_time_embed(esmdiff_ctx* c, float sigma, float* cond_out, void* stream) {
    if (!c || !c->finalized) return c ? c->fail("time_embed: weights not finalized") : 1;
    CK(cudaSetDevice(c->device));
    return launch_time_embed(c, sigma, cond_out, (cudaStream_t)stream);
}

int esmdiff_set_structure_coords(esmdiff_ctx* c, const float* coords, int B, int T, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    c->drop_graphs();                                     // captured forwards do not contain the branch (or its frames)
    if (!coords) {
        c->geom_on = false;
        return 0;
    }
    if (c->cfg.model_kind != 0 || c->cfg.v_heads <= 0 || c->cfg.v_heads > 256 || (3 * c->cfg.v_heads) % 64 != 0)
        return c->fail("set_structure_coords: needs the sampling network created with 0 < v_heads <= 256, 3 v_heads % 64 == 0");
    if (B <= 0 || T <= 0) return c->fail("set_structure_coords: B and T must be positive");
    const int64_t M = (int64_t)B * T, H = c->cfg.v_heads;
    if (M > c->geom_rows) {
        CK(cudaStreamSynchronize(st));
        void* olds[] = {c->g_rot, c->g_trans, c->g_work, c->g_mask, c->g_projout};
        for (void* o : olds)
            if (o) {
                cudaFree(o);
                for (auto& q : c->owned)
                    if (q == o) q = nullptr;
            }
        c->g_rot = c->g_trans = c->g_work = nullptr; c->g_mask = nullptr; c->g_projout = nullptr;
        c->geom_rows = 0;
        c->tmaps.clear();
        if (c->alloc(&c->g_rot, M * 9) || c->alloc(&c->g_trans, M * 3) || c->alloc(&c->g_mask, M) ||
            c->alloc(&c->g_work, M * 15 * H) || c->alloc(&c->g_projout, M * 15 * H))
            return 1;
        c->geom_rows = M;
    }
    if (esmdiff_geom_frames(coords, B, T, c->g_rot, c->g_trans, c->g_mask, st)) return c->fail("set_structure_coords: launch failed");
    c->launches++;
    c->geom_on = true;
    c->geom_B = B;
    c->geom_T = T;
    return 0;
}

int esmdiff_forward(esmdiff_ctx* c, const int64_t* seq, const int64_t* xt, int B, int T, const float* aux,
                    int64_t aux_row_stride, float* logits, float* emb, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return forward_impl(c, seq, xt, B, T, aux, aux_row_stride, logits, emb, (cudaStream_t)stream);
}

int esmdiff_forward_sigma(esmdiff_ctx* c, const int64_t* seq, const int64_t* xt, int B, int T, float sigma,
                          float* logits, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    if (!c->finalized) return c->fail("forward: esmdiff_finalize_weights has not succeeded");
    if (launch_time_embed(c, sigma, c->cond, (cudaStream_t)stream)) return 1;
    return forward_impl(c, seq, xt, B, T, c->cond, 0, logits, nullptr, (cudaStream_t)stream);
}

int esmdiff_logits_parameterization(esmdiff_ctx* c, const float* logits, const int64_t* xt, int B, int T,
                                    float* logp, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_sampler<2>(c, logits, nullptr, const_cast<int64_t*>(xt), logp, B * T, 0.f, 0.f, 0, 0,
                             (cudaStream_t)stream);
}

int esmdiff_sample_step(esmdiff_ctx* c, int64_t* x, const float* logits, const float* u, float mc_t, float mc_s,
                        int B, int T, uint64_t seed, uint32_t step, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_sampler<0>(c, logits, u, x, nullptr, B * T, mc_t, mc_s, seed, step, (cudaStream_t)stream);
}

int esmdiff_denoise_argmax(esmdiff_ctx* c, int64_t* x, const float* logits, int B, int T, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_sampler<1>(c, logits, nullptr, x, nullptr, B * T, 0.f, 0.f, 0, 0, (cudaStream_t)stream);
}

int esmdiff_schedule(int steps, float eps, float noise_eps, float* sigma, float* mc_t, float* mc_s) {
    if (steps <= 0 || !sigma || !mc_t || !mc_s) return 1;
    // torch.linspace(1, eps, steps+1) in fp32: symmetric fill from both ends
    const int n = steps + 1;
    const float start = 1.0f, end = eps;
    const float stepv = (end - start) / static_cast<float>(n - 1);
    const float dt = static_cast<float>((1.0 - static_cast<double>(eps)) / steps);
    const float one_m = 1.0f - noise_eps;
    for (int i = 0; i < n; ++i) {
        const float t = i < n / 2 ? start + stepv * i : end - stepv * (n - 1 - i);
        sigma[i] = -log1pf(-one_m * t);
        if (i < steps) {
            mc_t[i] = 1.0f - expf(-sigma[i]);
            const float ts = t - dt;
            mc_s[i] = 1.0f - expf(log1pf(-one_m * ts));
        }
    }
    return 0;
}

int esmdiff_ddpm_sample(esmdiff_ctx* c, const int64_t* seq, const int64_t* prior, int B, int T, int steps,
                        const float* sigma, const float* mc_t, const float* mc_s, uint64_t seed,
                        int noise_removal, int64_t* out, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    if (!c->finalized) return c->fail("ddpm_sample: esmdiff_finalize_weights has not succeeded");
    if (steps <= 0 || !sigma || !mc_t || !mc_s) return c->fail("ddpm_sample: bad schedule");
    cudaStream_t st = (cudaStream_t)stream;
    const int M = B * T;
    if (c->activate(0)) return 1;
    if (prior) {
        CK(cudaMemcpyAsync(out, prior, (size_t)M * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    } else {
        fill_i64_kernel<<<(M + 255) / 256, 256, 0, st>>>(reinterpret_cast<long long*>(out),
                                                         ESMDIFF_STRUCTURE_MASK_TOKEN, M);
        c->launches++;
        CK(cudaGetLastError());
    }
    // Samples are independent (the reference repeats one row, sample_esmdiff.py:186), so a small batch
    // is walked as TWO halves on two streams, each with its own workspace: at a few thousand token rows
    // every GEMM is one or two waves of tiles on 74 CTA pairs (out_proj at 13 samples: 84 tiles = 2 waves
    // for 1.14 waves of work) and the attention grid barely exceeds the 296 resident CTAs, so the tail
    // of each kernel leaves most of the chip idle; with two independent kernel chains in flight the
    // block scheduler fills those tails with the other half's kernels.  Results are bit-identical to
    // the single-stream loop (per-row arithmetic does not depend on the batch; the Philox counter uses
    // the row index of the whole batch).
    struct Part { int b0, nb; cudaStream_t s; };
    Part parts[2] = {{0, B, st}, {0, 0, st}};
    int nparts = 1;
    if (B >= 2 && M <= c->split_rows && !c->prof && !c->geom_on) {
        if (!c->sstream[0]) {
            for (int i = 0; i < 2; ++i) {
                CK(cudaStreamCreateWithFlags(&c->sstream[i], cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&c->ev_sjoin[i], cudaEventDisableTiming));
            }
            CK(cudaEventCreateWithFlags(&c->ev_sfork, cudaEventDisableTiming));
        }
        nparts = 2;
        parts[0] = {0, (B + 1) / 2, c->sstream[0]};
        parts[1] = {(B + 1) / 2, B - (B + 1) / 2, c->sstream[1]};
        CK(cudaEventRecord(c->ev_sfork, st));
        for (int h = 0; h < 2; ++h) CK(cudaStreamWaitEvent(c->sstream[h], c->ev_sfork, 0));
    }
    const int total = steps + (noise_removal ? 1 : 0);
    for (int i = 0; i < total; ++i) {               // step-major enqueue: both chains stay fed even without graphs
        if (i == steps && c->skip_denoise && nparts == 1) {
            // noise removal only rewrites rows that still hold a MASK: none left -> the forward is dead work
            if (!c->dev_count && c->alloc(&c->dev_count, 1)) return 1;
            CK(cudaMemsetAsync(c->dev_count, 0, sizeof(int), st));
            count_equal_kernel<<<(M + 255) / 256, 256, 0, st>>>(reinterpret_cast<const long long*>(out),
                                                                ESMDIFF_STRUCTURE_MASK_TOKEN, M, c->dev_count);
            c->launches++;
            int left = 1;
            CK(cudaMemcpyAsync(&left, c->dev_count, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (left == 0) break;
        }
        for (int h = 0; h < nparts; ++h) {
            const Part& pt = parts[h];
            if (c->activate(h) || ensure_workspace(c, (int64_t)pt.nb * T)) { c->activate(0); return 1; }
            const int64_t off = (int64_t)pt.b0 * T;
            int rc = launch_time_embed(c, sigma[i], c->cond, pt.s);
            rc = rc || forward_step(c, seq + off, out + off, pt.nb, T, c->logits_ws, pt.s);
            if (!rc) {
                if (i < steps)
                    rc = launch_sampler<0>(c, c->logits_ws, nullptr, out + off, nullptr, pt.nb * T, mc_t[i], mc_s[i], seed,
                                           (uint32_t)i, pt.s, (uint32_t)off);
                else
                    rc = launch_sampler<1>(c, c->logits_ws, nullptr, out + off, nullptr, pt.nb * T, 0.f, 0.f, 0, 0, pt.s);
            }
            if (rc) { c->activate(0); return 1; }
        }
    }
    if (c->activate(0)) return 1;
    if (nparts == 2)
        for (int h = 0; h < 2; ++h) {
            CK(cudaEventRecord(c->ev_sjoin[h], c->sstream[h]));
            CK(cudaStreamWaitEvent(st, c->ev_sjoin[h], 0));
        }
    return 0;
}


// ------------------------------------------------------------------------------------------------
// --mode gibbs: esm's iterative structure-track sampler (gibbs.cuh)
// ------------------------------------------------------------------------------------------------
static int gibbs_step_impl(esmdiff_ctx* c, int64_t* x, const float* logits, const float* noise, int B, int T,
                           float temperature, float top_p, int k, uint64_t seed, uint32_t step, uint32_t row_offset,
                           cudaStream_t st) {
    const int V = c->cfg.n_structure_heads;
    const int64_t M = (int64_t)B * T;
    if (V > sampler::THREADS * sampler::MAX_PER_THREAD) return c->fail("gibbs: vocabulary too large");
    if (!(temperature > 0.f)) return c->fail("gibbs: temperature must be positive");
    if ((size_t)T * sizeof(float) > 48 * 1024) return c->fail("gibbs: sequence too long for the commit kernel");
    if (M > c->gibbs_rows) {
        if (c->alloc(&c->gibbs_cand, M) || c->alloc(&c->gibbs_ent, M)) return 1;
        c->gibbs_rows = M;
    }
    {
        ProfScope prof(c, ESMDIFF_PROF_SAMPLER, (noise ? 8.0 : 4.0) * M * V, st);
        gibbs::gibbs_rows_kernel<<<(unsigned)M, sampler::THREADS, 0, st>>>(
            logits, (long long)V, noise, reinterpret_cast<const long long*>(x), c->gibbs_cand, c->gibbs_ent, V,
            ESMDIFF_STRUCTURE_MASK_TOKEN /* ids below the first special id are valid */, ESMDIFF_STRUCTURE_MASK_TOKEN,
            1.0f / temperature, top_p, (unsigned long long)seed, step, row_offset);
        c->launches++;
    }
    gibbs::gibbs_commit_kernel<<<B, 256, (size_t)T * sizeof(float), st>>>(reinterpret_cast<long long*>(x), c->gibbs_cand,
                                                                        c->gibbs_ent, T, k);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int esmdiff_gibbs_step(esmdiff_ctx* c, int64_t* x, const float* logits, const float* noise, int B, int T,
                       float temperature, float top_p, int k, uint64_t seed, uint32_t step, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return gibbs_step_impl(c, x, logits, noise, B, T, temperature, top_p, k, seed, step, 0, (cudaStream_t)stream);
}

int esmdiff_gibbs_sample(esmdiff_ctx* c, const int64_t* seq, const int64_t* prior, int B, int T, int steps,
                         const int* k_per_step, float temperature, float top_p, uint64_t seed, int64_t* out,
                         void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    if (!c->finalized) return c->fail("gibbs_sample: esmdiff_finalize_weights has not succeeded");
    if (steps <= 0 || !k_per_step || !prior) return c->fail("gibbs_sample: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = (int64_t)B * T;
    if (c->activate(0) || ensure_workspace(c, M)) return 1;
    CK(cudaMemcpyAsync(out, prior, (size_t)M * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    // auxiliary_embeddings = None (esm's sampler knows nothing about the diffusion time): the conditioning row is zero
    CK(cudaMemsetAsync(c->cond, 0, (size_t)c->cfg.d_model * sizeof(float), st));
    for (int i = 0; i < steps; ++i) {
        if (forward_step(c, seq, out, B, T, c->logits_ws, st)) return 1;
        if (gibbs_step_impl(c, out, c->logits_ws, nullptr, B, T, temperature, top_p, k_per_step[i], seed, (uint32_t)i, 0, st))
            return 1;
    }
    return 0;
}

int esmdiff_decode_structure(esmdiff_ctx* c, const int64_t* tokens, int B, int T, float* bb_out, float* o_out,
                             float* plddt_out, float* affine_out, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    if (c->cfg.model_kind != 1) return c->fail("decode_structure: context was not created as a structure token decoder");
    if (!c->finalized) return c->fail("decode_structure: esmdiff_finalize_weights has not succeeded");
    if (B <= 0 || T < 3 || !tokens || !bb_out) return c->fail("decode_structure: bad arguments (T counts BOS and EOS)");
    if (plddt_out && c->cfg.n_aux_out == 0) return c->fail("decode_structure: this decoder has no pLDDT head");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M64 = (int64_t)B * T;
    if (M64 > (1ll << 30)) return c->fail("decode_structure: B*T too large");
    const int M = (int)M64;
    const int D = c->cfg.d_model, V = c->cfg.n_structure_heads;
    if (ensure_workspace(c, M)) return 1;
    if (c->qk_fused && ensure_rope(c, T, st)) return 1;
    {
        ProfScope prof(c, ESMDIFF_PROF_EMBED, 6.0 * M * D, st);
        dec::embed_tokens_kernel<<<(M + 7) / 8, 256, 0, st>>>(reinterpret_cast<const long long*>(tokens), c->dec_embed, c->x,
                                                             c->ln_fold ? c->xn : nullptr, c->stats, M, D,
                                                             c->cfg.struct_vocab, c->dev_err);
        c->stats_span = 128;
        c->stats_ptr = c->stats;
    }
    c->launches++;
    CK(cudaGetLastError());
    if (run_blocks(c, B, T, st)) return 1;
    if (launch_layernorm(c, c->x, c->norm_w, nullptr, c->xn, M, D, st)) return 1;
    // Dim6RotStructureHead: ffn1 + GELU, LayerNorm, proj (same graph as a RegressionHead)
    if (launch_gemm(c, gemm::EPI_BIAS_GELU_F32, c->xn, c->h0_w, M, D, D, c->headh, D, c->h0_b, 1.f, st)) return 1;
    if (launch_layernorm(c, c->headh, c->h2_w, c->h2_b, c->att, M, D, st)) return 1;
    if (launch_gemm(c, gemm::EPI_BIAS_F32, c->att, c->h3_w, M, V, D, c->logits_ws, V, c->h3_b, 1.f, st)) return 1;
    dec::backbone_frames_kernel<<<(M + 127) / 128, 128, 0, st>>>(c->logits_ws, V, bb_out, M);
    c->launches++;
    if (affine_out) CK(cudaMemcpyAsync(affine_out, c->logits_ws, (size_t)M * V * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (o_out) {
        dec::infer_oxygen_kernel<<<(M + 127) / 128, 128, 0, st>>>(bb_out, o_out, B, T);
        c->launches++;
    }
    if (plddt_out) {
        const int A = c->cfg.n_aux_out;
        if (launch_gemm(c, gemm::EPI_BIAS_GELU_F32, c->xn, c->p0_w, M, D, D, c->headh, D, c->p0_b, 1.f, st)) return 1;
        if (launch_layernorm(c, c->headh, c->p2_w, c->p2_b, c->att, M, D, st)) return 1;
        if (launch_gemm(c, gemm::EPI_BIAS_F32, c->att, c->p3_w, M, A, D, c->aux_ws, A, c->p3_b, 1.f, st)) return 1;
        dec::plddt_mean_kernel<<<(M + 7) / 8, 256, 0, st>>>(c->aux_ws, A, A, plddt_out, M);
        c->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

int esmdiff_synchronize(esmdiff_ctx* c, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    CK(cudaGetLastError());
    int abort_flag = 0, tok_err = 0;
    CK(cudaMemcpyFromSymbol(&abort_flag, g_abort_flag, sizeof(int)));
    CK(cudaMemcpy(&tok_err, c->dev_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (abort_flag) {
        int zero = 0;
        cudaMemcpyToSymbol(g_abort_flag, &zero, sizeof(int));
        return c->fail("pipeline watchdog: an mbarrier wait timed out inside a tcgen05 kernel");
    }
    if (tok_err) {
        cudaMemset(c->dev_err, 0, sizeof(int));
        return c->fail("IndexError: token id out of range in embedding lookup");
    }
    return 0;
}

int esmdiff_ddpm_sample_host(esmdiff_ctx* c, const int64_t* seq_host, const int64_t* prior_host, int B, int T,
                             int steps, float eps, uint64_t seed, int noise_removal, int64_t* out_host) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    const int64_t M = (int64_t)B * T;
    if (M > c->tok_rows) {
        if (c->alloc(&c->x_tok, 2 * M)) return 1;
        if (c->alloc(&c->seq_tok, M)) return 1;
        c->tok_rows = M;
    }
    std::vector<float> sg(steps + 1), a(steps), b(steps);
    if (esmdiff_schedule(steps, eps, 1e-3f, sg.data(), a.data(), b.data())) return c->fail("bad schedule");
    cudaStream_t st = 0;
    CK(cudaMemcpyAsync(c->seq_tok, seq_host, M * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    int64_t* prior_dev = nullptr;
    if (prior_host) {
        prior_dev = c->x_tok + M;
        CK(cudaMemcpyAsync(prior_dev, prior_host, M * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    }
    if (esmdiff_ddpm_sample(c, c->seq_tok, prior_dev, B, T, steps, sg.data(), a.data(), b.data(), seed,
                            noise_removal, c->x_tok, st))
        return 1;
    CK(cudaMemcpyAsync(out_host, c->x_tok, M * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    return esmdiff_synchronize(c, st);
}

int64_t esmdiff_launch_count(const esmdiff_ctx* c) { return c ? c->launches : 0; }

int esmdiff_profile_enable(esmdiff_ctx* c, int on) {
    if (!c) return 1;
    if (on) {
        c->prof_recs.clear();
        c->ev_used = 0;
    }
    c->prof = on != 0;
    return 0;
}

int esmdiff_profile_read(esmdiff_ctx* c, int kind, double* ms, double* work, int64_t* launches) {
    if (!c || !ms || !work || !launches) return 1;
    CK(cudaSetDevice(c->device));
    *ms = 0.0; *work = 0.0; *launches = 0;
    for (const auto& r : c->prof_recs) {
        if (r.kind != kind) continue;
        CK(cudaEventSynchronize(r.b));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, r.a, r.b));
        *ms += t; *work += r.work; *launches += 1;
    }
    return 0;
}

int esmdiff_op_gemm(esmdiff_ctx* c, int epi, const void* a, const void* w, int M, int N, int K, void* out,
                    int64_t ldo, const float* bias, float scale, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_gemm(c, epi, (const bf16*)a, (const bf16*)w, M, N, K, out, ldo, bias, scale, (cudaStream_t)stream);
}
int esmdiff_op_gemm_ln(esmdiff_ctx* c, int epi, const void* a, const void* w, int M, int N, int K, void* out,
                       int64_t ldo, const float* bias, float scale, const void* stats_in, const float* colsum,
                       void* stats_out, void* xb_out, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    GemmLN ln;
    ln.stats_in = (const float2*)stats_in;
    ln.colsum = colsum;
    ln.stats_out = (float2*)stats_out;
    ln.xb_out = (bf16*)xb_out;
    ln.stats_span = (stats_in != nullptr && stats_in == c->stats_ptr) ? c->stats_span : 128;   // as the producing call left it
    return launch_gemm(c, epi, (const bf16*)a, (const bf16*)w, M, N, K, out, ldo, bias, scale, (cudaStream_t)stream, ln);
}
int esmdiff_op_fold_layernorm(esmdiff_ctx* c, const float* w, const float* gamma, const float* beta, void* dst,
                              float* colsum, float* bias, int64_t rows, int64_t cols, int swiglu_hidden,
                              void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    ew::fold_layernorm_weight_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        w, gamma, beta, (bf16*)dst, colsum, bias, rows, cols, swiglu_hidden);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}
int esmdiff_op_layernorm(esmdiff_ctx* c, const float* x, const float* w, const float* b, void* y, int M, int D,
                         void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_layernorm(c, x, w, b, (bf16*)y, M, D, (cudaStream_t)stream);
}
int esmdiff_op_qk_norm_rope(esmdiff_ctx* c, void* qkv, const float* qw, const float* kw, int B, int T, int D,
                            void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_qk_norm_rope(c, (bf16*)qkv, qw, kw, B * T, T, D, (cudaStream_t)stream);
}
int esmdiff_op_attention(esmdiff_ctx* c, const void* qkv, void* out, int B, int T, int H, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_attention(c, (const bf16*)qkv, (bf16*)out, B, T, H, nullptr, (cudaStream_t)stream);
}
int esmdiff_op_attention_ln(esmdiff_ctx* c, const void* qkv, const float* qk_sumsq, void* out, int B, int T, int H,
                            void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    return launch_attention(c, (const bf16*)qkv, (bf16*)out, B, T, H, qk_sumsq, (cudaStream_t)stream);
}
int esmdiff_op_gemm_qkv_rope(esmdiff_ctx* c, const void* a, const void* w, int M, int N, int K, void* out, int64_t ldo,
                             const float* bias, const void* stats_in, const float* colsum, const float* qk_gamma,
                             float* qk_sumsq, int T, int n_rope, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    if (ensure_rope(c, T, (cudaStream_t)stream)) return 1;
    GemmLN ln;
    ln.stats_in = (const float2*)stats_in;
    ln.colsum = colsum;
    ln.rope = c->rope_rows; ln.qk_gamma = qk_gamma; ln.qk_sumsq = qk_sumsq; ln.T = T; ln.n_rope = n_rope;
    ln.stats_span = (stats_in != nullptr && stats_in == c->stats_ptr) ? c->stats_span : 128;
    return launch_gemm(c, gemm::EPI_QKV_ROPE_LN, (const bf16*)a, (const bf16*)w, M, N, K, out, ldo, bias, 1.f,
                       (cudaStream_t)stream, ln);
}
int esmdiff_op_fold_layernorm_centered(esmdiff_ctx* c, const float* w, const float* gamma, const float* beta, void* dst,
                                       float* colsum, float* bias, int64_t rows, int64_t cols, int64_t center_rows,
                                       int64_t center_block, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    if (center_rows % center_block != 0 || center_rows > rows) return c->fail("fold_layernorm_centered: bad centring blocks");
    float* cm = nullptr;
    const int64_t nb = center_rows / center_block;
    if (nb > 0) {
        CK(cudaMallocAsync(&cm, nb * cols * sizeof(float), (cudaStream_t)stream));
        ew::column_mean_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(
            w, cm, center_block, cols);
    }
    ew::fold_layernorm_weight_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        w, gamma, beta, (bf16*)dst, colsum, bias, rows, cols, 0, cm, center_rows, center_block);
    c->launches += nb > 0 ? 2 : 1;
    CK(cudaGetLastError());
    if (cm) CK(cudaFreeAsync(cm, (cudaStream_t)stream));
    return 0;
}
int esmdiff_op_stats_span(const esmdiff_ctx* c) { return c ? c->stats_span : 0; }
int esmdiff_set_time_conditioning(esmdiff_ctx* c, int on) {
    if (!c) return 1;
    c->cfg.time_conditioning = on ? 1 : 0;
    return 0;
}
int esmdiff_op_convert_bf16(esmdiff_ctx* c, const float* src, void* dst, int64_t rows, int64_t cols,
                            int swiglu_hidden, void* stream) {
    if (!c) return 1;
    CK(cudaSetDevice(c->device));
    const int64_t n = rows * cols;
    ew::convert_rows_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, (bf16*)dst, rows, cols, swiglu_hidden);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
