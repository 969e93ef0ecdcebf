// libdwb C ABI: plan life cycle, weight ingestion, forward and the graph-captured sampler.
// Host-side orchestration only; all arithmetic lives in the kernel translation units.
//
// Reference call sites this file stands in for:
//   generate.py:94-103  construct_model(cfg).cuda(); load_state_dict      -> dwb_plan_create / set_tensor / finalize
//   generate.py:51      net((x, t), mel_spec)                              -> dwb_forward
//   generate.py:23-55   sampling()                                         -> dwb_sample
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "fft_plan.cuh"
#include "kernels.h"

namespace dwb {
bool pdl_enabled() {
    // off by default: measured on one box (tools/ab_env.py) the 200-step loop runs at 20.0 clips/s with programmatic edges
    // between the hot kernels and at 20.5 without - the early-scheduled CTAs of the next kernel hold SMs while they wait
    static const bool on = [] { const char *e = getenv("DWB_PDL"); return e && atoi(e) == 1; }();
    return on;
}
}  // namespace dwb

namespace dwb {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    return DWB_ERR_CUDA;
}

struct Tensor {
    std::vector<int64_t> shape;
    int dtype = DWB_F32;
    void *dev = nullptr;
    size_t bytes = 0;
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

enum OpKind { OP_BLOCK = 0, OP_DOWN = 1, OP_UP = 2 };

struct Op {
    int kind;
    std::string prefix;
    int H;        // block: width; pools: input width
    int l;        // input length
    int Ho;       // pools: output width
    int s;        // pool factor
    int in_buf = -1, out_buf = -1, skip_buf = -1;
    // block weights
    float ln1_m = 0, ln1_s = 1, ln2_m = 0, ln2_s = 1;
    float *kf = nullptr, *k32 = nullptr;
    float *Wo_t = nullptr, *bo = nullptr, *W1_t = nullptr, *b1 = nullptr, *W2_t = nullptr, *b2 = nullptr;
    uint4 *Wo_f[2] = {nullptr, nullptr}, *W1_f[2] = {nullptr, nullptr}, *W2_f[2] = {nullptr, nullptr};
    bool mma = false;
    bool umma = false;             // tcgen05 path (mix_umma.cu)
    bool gemm = false;             // tcgen05 GEMM-per-contraction path (mix_gemm_umma.cu): widths mix_umma.cu has no tile for
    bool gemm2 = false;            // ... with 512 output columns per CTA (pool_umma.cu, H = 512); DWB_GEMM=1 keeps the first form
    uint8_t *Wimg = nullptr;
    float *bimg = nullptr;
    int F = 0;
    int part_off = 0;      // offset of this block's fc_t rows in the stacked embedding output
    int64_t cond_off = 0;  // float offset (per cond batch element) of this block's conditioning features
    // pool weights
    float *Wp_t = nullptr, *bp = nullptr;
    uint4 *Wp_f[2] = {nullptr, nullptr};
    bool pumma = false;            // tcgen05 pool (pool_umma.cu)
    uint8_t *Wp_img = nullptr;
};

struct WaveLayer {
    int dilation;
    float *Wd_t, *bd, *Wr_t, *br, *Ws_t, *bs;
    uint4 *Wd_f[2], *Wr_f[2], *Ws_f[2];
    bool mma;
    bool umma;                     // tcgen05 path (wave_umma.cu)
    uint8_t *Wimg;
    int part_off;
    int64_t cond_off;
};

struct StepKey {              // what the captured one-step graph depends on
    const void *cond = nullptr;
    int cond_batch = 0, B = 0, L = 0;
    bool operator==(const StepKey &o) const { return cond == o.cond && cond_batch == o.cond_batch && B == o.B && L == o.L; }
};

}  // namespace dwb

using namespace dwb;

struct dwb_plan {
    dwb_config cfg;
    int device = 0;
    bool finalized = false;
    std::map<std::string, Tensor> tensors;
    std::vector<void *> owned;
    int64_t launches = 0;
    bool use_mma = true;           // DWB_MIX=simt in the environment selects the exact-fp32 SIMT channel mixing
    bool use_umma = true;          // DWB_MIX=mma keeps the legacy mma.sync path instead of tcgen05

    // embedding
    float *eW1 = nullptr, *eb1 = nullptr, *eW2 = nullptr, *eb2 = nullptr, *Wt_all = nullptr, *bt_all = nullptr;
    int Mtot = 0;
    // input conv / head
    float *init_w = nullptr, *init_b = nullptr;
    float *Wf_t = nullptr, *bf = nullptr, *wz = nullptr;
    uint4 *Wf_f[2] = {nullptr, nullptr};      // split-bf16 A fragments of the head's C -> C weight
    uint8_t *Wf_img = nullptr;                // tcgen05 head (pool_umma.cu)
    float bz = 0.f, norm_m = 0.f, norm_s = 1.f;
    int headC = 0;

    std::vector<Op> ops;           // sashimi
    int n_bufs = 0;
    // S4 pointwise tables for sequence lengths other than cfg.L (models/s4.py:1387: L_kernel = min(L, l_max)), one entry
    // per block in op order.  r < l: the kernel's first r taps (kf).  r > l: overlap-save, one table per direction.
    struct LenTab { float *kf = nullptr, *kf_c = nullptr, *kf_a = nullptr; };
    std::map<int, std::vector<LenTab>> len_tabs;
    std::vector<int> len_lru;      // most recent last; at most 2 lengths are kept
    std::vector<WaveLayer> wl;     // wavenet
    int64_t cond_total = 0;        // floats of conditioning features per cond batch element
    struct CondBlock {             // in-model mel front end of one block (folded weights)
        float *w[2], *b[2];        // ConvTranspose2d taps [3][2 s_i], bias (1)
        int s[2];
        float *Wm_t, *bm;          // mel_conv [mel_bands][Hc], (Hc)
        int Hc, l;                 // l = 0: the caller's L (wavenet; `off` is then the layer index)
        int64_t off;
    };
    std::vector<CondBlock> cond_blocks;

    // workspace (grow-only, keyed on B*L)
    int ws_B = 0, ws_L = 0;
    std::vector<float *> bufs, stat_bufs;
    float *g_buf = nullptr, *emb_buf = nullptr, *part_buf = nullptr;
    float *hid_buf = nullptr;      // (B,F,l) hidden activations of the GEMM-per-contraction mixing path
    float *skip_acc = nullptr;     // wavenet
    std::vector<void *> ws_owned;

    // sampler: per-step records (T, Mtot + 4) = [fc_t outputs of every layer | c1, sqrt(alpha), sigma, noise slot];
    // the record of the current step is copied to rec_cur, which is what the captured ONE-STEP graph reads, so
    // the graph depends on neither T, the step, nor any caller buffer (x lives in x_cur, noise behind noise_base)
    float *table_emb = nullptr, *table_part = nullptr, *table_t = nullptr, *step_tab = nullptr;
    int table_T = 0;
    std::vector<float> tab_coef;
    float *x_cur = nullptr, *rec_cur = nullptr;     // workspace: (B,1,L) state, (Mtot + 4) current record
    const float **noise_base = nullptr;             // workspace: device word holding the call's noise base pointer
    cudaGraphExec_t graph_exec = nullptr;
    cudaStream_t cap_stream = nullptr;   // capture happens on a plan-owned stream (the legacy default stream cannot capture)
    StepKey graph_key;
    int64_t graph_nodes = 0;
};

namespace dwb {

static int dev_alloc(dwb_plan *p, size_t bytes, void **out, bool workspace = false) {
    void *d = nullptr;
    DWB_CUDA(cudaMalloc(&d, bytes ? bytes : 4));
    (workspace ? p->ws_owned : p->owned).push_back(d);
    *out = d;
    return DWB_OK;
}

static const Tensor *find(dwb_plan *p, const std::string &name) {
    auto it = p->tensors.find(name);
    return it == p->tensors.end() ? nullptr : &it->second;
}

static int need(dwb_plan *p, const std::string &name, int64_t numel, const Tensor **out) {
    const Tensor *t = find(p, name);
    DWB_REQUIRE(t, DWB_ERR_MISSING, "state_dict entry `%s` was not supplied", name.c_str());
    DWB_REQUIRE(t->dtype == DWB_F32, DWB_ERR_INVALID, "`%s` must be float32", name.c_str());
    DWB_REQUIRE(t->numel() == numel, DWB_ERR_INVALID, "`%s` has %lld elements, expected %lld", name.c_str(),
                (long long)t->numel(), (long long)numel);
    *out = t;
    return DWB_OK;
}

#define TRY(expr)                      \
    do {                               \
        int _rc = (expr);              \
        if (_rc != DWB_OK) return _rc; \
    } while (0)

static int scalar_of(dwb_plan *p, const std::string &name, float *out, cudaStream_t st) {
    const Tensor *t;
    TRY(need(p, name, 1, &t));
    DWB_CUDA(cudaMemcpyAsync(out, t->dev, sizeof(float), cudaMemcpyDeviceToHost, st));
    DWB_CUDA(cudaStreamSynchronize(st));
    return DWB_OK;
}

// weight-normed conv `prefix`.{weight_v,weight_g,bias} (or plain .weight/.bias) -> transposed fold
static int folded(dwb_plan *p, const std::string &prefix, int M, int Kin, int taps, bool weight_norm, float **Wt,
                  float **bias, cudaStream_t st) {
    const Tensor *v, *g = nullptr, *b;
    TRY(need(p, prefix + (weight_norm ? ".weight_v" : ".weight"), (int64_t)M * Kin * taps, &v));
    if (weight_norm) TRY(need(p, prefix + ".weight_g", M, &g));
    TRY(need(p, prefix + ".bias", M, &b));
    void *o;
    TRY(dev_alloc(p, (size_t)M * Kin * taps * sizeof(float), &o));
    TRY(fold_weight((const float *)v->dev, g ? (const float *)g->dev : nullptr, M, Kin, taps, (float *)o, st));
    p->launches += 1;
    *Wt = (float *)o;
    *bias = (float *)b->dev;
    return DWB_OK;
}

static int build_sashimi_ops(dwb_plan *p) {
    const dwb_config &c = p->cfg;
    int H = c.d_model, L = c.L, i = 0;
    p->ops.clear();
    for (int q = 0; q < c.n_pool; ++q) {
        const int s = c.pool[q];
        if (c.unet)
            for (int k = 0; k < c.n_layers; ++k) {
                Op o{};
                o.kind = OP_BLOCK; o.prefix = "d_layers." + std::to_string(i++) + "."; o.H = H; o.l = L;
                p->ops.push_back(o);
            }
        DWB_REQUIRE(L % s == 0, DWB_ERR_INVALID, "L=%d not divisible by pool %d", L, s);
        Op o{};
        o.kind = OP_DOWN; o.prefix = "d_layers." + std::to_string(i++) + "."; o.H = H; o.l = L; o.Ho = H * c.expand; o.s = s;
        p->ops.push_back(o);
        L /= s;
        H *= c.expand;
    }
    for (int k = 0; k < c.n_layers; ++k) {
        Op o{};
        o.kind = OP_BLOCK; o.prefix = "c_layers." + std::to_string(k) + "."; o.H = H; o.l = L;
        p->ops.push_back(o);
    }
    i = 0;
    for (int q = c.n_pool - 1; q >= 0; --q) {
        const int s = c.pool[q];
        Op o{};
        o.kind = OP_UP; o.prefix = "u_layers." + std::to_string(i++) + "."; o.H = H; o.l = L; o.Ho = H / c.expand; o.s = s;
        p->ops.push_back(o);
        H /= c.expand;
        L *= s;
        for (int k = 0; k < c.n_layers; ++k) {
            Op b{};
            b.kind = OP_BLOCK; b.prefix = "u_layers." + std::to_string(i++) + "."; b.H = H; b.l = L;
            p->ops.push_back(b);
        }
    }
    // ---- buffer assignment: UNet skip stack (sashimi.py:292-307) with a free list
    std::vector<int> stack, freel;
    std::vector<bool> saved;
    int nb = 0;
    auto alloc = [&]() {
        if (!freel.empty()) { int b = freel.back(); freel.pop_back(); return b; }
        saved.push_back(false);
        return nb++;
    };
    int cur = alloc();   // output of init_conv
    size_t idx = 0;
    const size_t n_d = (size_t)c.n_pool * ((c.unet ? c.n_layers : 0) + 1);
    for (; idx < n_d; ++idx) {          // down path: every input is saved
        Op &o = p->ops[idx];
        stack.push_back(cur); saved[cur] = true;
        o.in_buf = cur; o.out_buf = alloc(); cur = o.out_buf;
    }
    stack.push_back(cur); saved[cur] = true;
    for (int k = 0; k < c.n_layers; ++k, ++idx) {   // centre
        Op &o = p->ops[idx];
        o.in_buf = cur; o.out_buf = alloc();
        if (!saved[cur]) freel.push_back(cur);
        cur = o.out_buf;
    }
    {   // x = x + outputs.pop() after the centre stack: fused into the last centre block
        Op &o = p->ops[idx - 1];
        DWB_REQUIRE(c.n_layers >= 1, DWB_ERR_UNSUPPORTED, "n_layers must be >= 1");
        o.skip_buf = stack.back(); stack.pop_back();
    }
    std::vector<int> pending_free;   // skip buffers are recycled one op after the op that read them
    pending_free.push_back(p->ops[idx - 1].skip_buf);
    for (; idx < p->ops.size(); ++idx) {
        Op &o = p->ops[idx];
        for (int b : pending_free) { saved[b] = false; freel.push_back(b); }
        pending_free.clear();
        o.in_buf = cur; o.out_buf = alloc();
        if (o.kind == OP_UP || c.unet) {
            DWB_REQUIRE(!stack.empty(), DWB_ERR_INVALID, "skip stack underflow");
            o.skip_buf = stack.back(); stack.pop_back();
            pending_free.push_back(o.skip_buf);
        }
        if (!saved[cur]) freel.push_back(cur);
        cur = o.out_buf;
    }
    p->n_bufs = nb;
    return DWB_OK;
}

static int finalize_common(dwb_plan *p, const std::string &emb_prefix, const std::string &init_prefix, int C,
                           cudaStream_t st) {
    const dwb_config &c = p->cfg;
    const Tensor *t;
    TRY(need(p, emb_prefix + "fc_t1.weight", (int64_t)c.embed_mid * c.embed_in, &t)); p->eW1 = (float *)t->dev;
    TRY(need(p, emb_prefix + "fc_t1.bias", c.embed_mid, &t)); p->eb1 = (float *)t->dev;
    TRY(need(p, emb_prefix + "fc_t2.weight", (int64_t)c.embed_out * c.embed_mid, &t)); p->eW2 = (float *)t->dev;
    TRY(need(p, emb_prefix + "fc_t2.bias", c.embed_out, &t)); p->eb2 = (float *)t->dev;
    float *wt;
    TRY(folded(p, init_prefix, C, 1, 1, true, &wt, &p->init_b, st));
    p->init_w = wt;   // [1][C]
    return DWB_OK;
}

static int finalize_head(dwb_plan *p, int C, cudaStream_t st) {
    TRY(folded(p, "final_conv.0.conv", C, C, 1, true, &p->Wf_t, &p->bf, st));
    const Tensor *t;
    TRY(need(p, "final_conv.2.conv.weight", C, &t)); p->wz = (float *)t->dev;
    TRY(scalar_of(p, "final_conv.2.conv.bias", &p->bz, st));
    p->headC = C;
    if (p->use_mma && C % 16 == 0) {
        for (int q = 0; q < 2; ++q) {
            void *d;
            TRY(dev_alloc(p, (size_t)C * C * 2, &d));
            p->Wf_f[q] = (uint4 *)d;
        }
        TRY(frag_pack(p->Wf_t, C, C, (uint32_t *)p->Wf_f[0], (uint32_t *)p->Wf_f[1], st));
        p->launches += 1;
    }
    // DWB_HEAD=mma keeps the mma.sync head
    static const bool head_mma_only = [] { const char *e = getenv("DWB_HEAD"); return e && std::string(e) == "mma"; }();
    if (p->use_mma && p->use_umma && !head_mma_only && head_umma_supported(C)) {
        void *d;
        TRY(dev_alloc(p, head_umma_image_bytes(C), &d)); p->Wf_img = (uint8_t *)d;
        TRY(head_umma_pack(C, p->Wf_t, p->Wf_img, st));
        p->launches += 1;
    }
    return DWB_OK;
}

// stack every block's fc_t weight/bias into one (Mtot, E_out) matrix
static int stack_fc_t(dwb_plan *p, const std::vector<std::pair<std::string, int>> &blocks, cudaStream_t st) {
    const int E = p->cfg.embed_out;
    int tot = 0;
    for (auto &b : blocks) tot += b.second;
    p->Mtot = tot;
    void *W, *bb;
    TRY(dev_alloc(p, (size_t)tot * E * sizeof(float), &W));
    TRY(dev_alloc(p, (size_t)tot * sizeof(float), &bb));
    int off = 0;
    for (auto &b : blocks) {
        const Tensor *w, *bi;
        TRY(need(p, b.first + "fc_t.weight", (int64_t)b.second * E, &w));
        TRY(need(p, b.first + "fc_t.bias", b.second, &bi));
        DWB_CUDA(cudaMemcpyAsync((float *)W + (size_t)off * E, w->dev, (size_t)b.second * E * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
        DWB_CUDA(cudaMemcpyAsync((float *)bb + off, bi->dev, (size_t)b.second * sizeof(float), cudaMemcpyDeviceToDevice, st));
        off += b.second;
    }
    p->Wt_all = (float *)W;
    p->bt_all = (float *)bb;
    return DWB_OK;
}

static int finalize_sashimi(dwb_plan *p, cudaStream_t st) {
    const dwb_config &c = p->cfg;
    DWB_REQUIRE(c.n_pool >= 0 && c.n_pool <= DWB_MAX_POOL, DWB_ERR_INVALID, "n_pool=%d out of range", c.n_pool);
    DWB_REQUIRE(c.expand >= 1 && c.ff >= 1 && c.d_model >= 1 && c.L >= 2, DWB_ERR_INVALID, "bad sashimi config");
    TRY(build_sashimi_ops(p));
    TRY(finalize_common(p, "", "init_conv.0.conv", c.d_model, st));
    TRY(scalar_of(p, "norm.m", &p->norm_m, st));
    TRY(scalar_of(p, "norm.s", &p->norm_s, st));
    TRY(finalize_head(p, c.d_model, st));

    const int N = c.d_state_half;
    std::vector<std::pair<std::string, int>> fc;
    size_t max_khat = 0, max_k64 = 0;
    for (auto &o : p->ops)
        if (o.kind == OP_BLOCK) {
            max_khat = std::max(max_khat, (size_t)2 * o.H * (o.l / 2 + 1) * 2);
            max_k64 = std::max(max_k64, (size_t)2 * o.H * o.l);
        }
    double *khat = nullptr, *k64 = nullptr;
    DWB_CUDA(cudaMalloc(&khat, max_khat * sizeof(double)));
    cudaError_t e = cudaMalloc(&k64, max_k64 * sizeof(double));
    if (e != cudaSuccess) { cudaFree(khat); return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
    int rc = DWB_OK;
    int part_off = 0;
    int64_t cond_off = 0;
    for (auto &o : p->ops) {
        if (rc != DWB_OK) break;
        auto body = [&]() -> int {
            if (o.kind == OP_BLOCK) {
                const std::string k = o.prefix + "layer.kernel.kernel.";
                const int H = o.H, l = o.l;
                const Tensor *C, *B, *P, *iwr, *wim, *ldt, *D;
                TRY(need(p, k + "C", (int64_t)2 * H * N * 2, &C));
                TRY(need(p, k + "B", (int64_t)H * N * 2, &B));
                TRY(need(p, k + "P", (int64_t)H * N * 2, &P));
                TRY(need(p, k + "inv_w_real", (int64_t)H * N, &iwr));
                TRY(need(p, k + "w_imag", (int64_t)H * N, &wim));
                TRY(need(p, k + "log_dt", H, &ldt));
                TRY(need(p, o.prefix + "layer.D", H, &D));
                if (const Tensor *Lt = find(p, k + "L")) {
                    long long Lv = 0;
                    DWB_REQUIRE(Lt->dtype == DWB_I64 && Lt->numel() == 1, DWB_ERR_INVALID, "`%sL` must be a scalar int64", k.c_str());
                    DWB_CUDA(cudaMemcpyAsync(&Lv, Lt->dev, 8, cudaMemcpyDeviceToHost, st));
                    DWB_CUDA(cudaStreamSynchronize(st));
                    DWB_REQUIRE(Lv == l, DWB_ERR_STATE,
                                "`%sL` is %lld but the stage length is %d: apply the one-off C rewrite "
                                "(models/s4.py:525-551) on the host before loading", k.c_str(), Lv, l);
                }
                const Tensor *om = find(p, "nodes." + std::to_string(l));
                if (om) DWB_REQUIRE(om->numel() == (int64_t)(l / 2 + 1) * 2, DWB_ERR_INVALID, "nodes.%d has wrong size", l);
                void *k32, *kf;
                TRY(dev_alloc(p, (size_t)2 * H * l * sizeof(float), &k32));
                const int lg = fft_log2m_for(l);
                DWB_REQUIRE(lg > 0, DWB_ERR_UNSUPPORTED, "stage length %d exceeds the in-shared-memory FFT (max %d)", l, 1 << FFT_MAX_LOG2M);
                TRY(dev_alloc(p, (size_t)H * fft_table_floats(lg) * sizeof(float), &kf));
                TRY(s4_generate((const float *)C->dev, (const float *)B->dev, (const float *)P->dev, (const float *)iwr->dev,
                                (const float *)wim->dev, (const float *)ldt->dev, om ? (const float *)om->dev : nullptr, H, N, l,
                                khat, k64, (float *)k32, st, &p->launches));
                TRY(fftconv_prepare_f64(k64, (const float *)D->dev, H, l, (float *)kf, st));
                p->launches += 1;
                const float2 *tw;
                TRY(fft_twiddles(lg, st, &tw));
                if (fft_table_mode(lg, l) == 2) {            // created here: not allowed during graph capture
                    const float2 *tw2;
                    TRY(fft_pair_twiddles(lg, st, &tw2));
                }
                o.k32 = (float *)k32; o.kf = (float *)kf;
                TRY(scalar_of(p, o.prefix + "norm1.m", &o.ln1_m, st));
                TRY(scalar_of(p, o.prefix + "norm1.s", &o.ln1_s, st));
                TRY(scalar_of(p, o.prefix + "norm2.m", &o.ln2_m, st));
                TRY(scalar_of(p, o.prefix + "norm2.s", &o.ln2_s, st));
                o.F = c.ff * H;
                TRY(folded(p, o.prefix + "layer.output_linear.0", 2 * H, H, 1, false, &o.Wo_t, &o.bo, st));
                TRY(folded(p, o.prefix + "ff.ff.0.conv", o.F, H, 1, true, &o.W1_t, &o.b1, st));
                TRY(folded(p, o.prefix + "ff.ff.2.conv", H, o.F, 1, true, &o.W2_t, &o.b2, st));
                o.umma = p->use_mma && p->use_umma && mix_umma_supported(H, o.F, l);
                if (o.umma) {
                    void *d;
                    TRY(dev_alloc(p, mix_umma_image_bytes(H), &d)); o.Wimg = (uint8_t *)d;
                    TRY(dev_alloc(p, (size_t)5 * H * sizeof(float), &d)); o.bimg = (float *)d;
                    TRY(mix_umma_pack(H, o.Wo_t, o.W1_t, o.W2_t, o.bo, o.b1, o.b2, o.Wimg, o.bimg, st));
                    p->launches += 1;
                }
                o.gemm = !o.umma && p->use_mma && p->use_umma && mix_gemm_supported(H, o.F, l);
                static const bool gemm1_only = [] { const char *e = getenv("DWB_GEMM"); return e && atoi(e) == 1; }();
                o.gemm2 = o.gemm && !gemm1_only && mix_gemm2_supported(H, o.F, l);
                if (o.gemm2) {
                    void *d;
                    TRY(dev_alloc(p, mix_gemm2_image_bytes(H, o.F), &d)); o.Wimg = (uint8_t *)d;
                    TRY(mix_gemm2_pack(H, o.F, o.Wo_t, o.W1_t, o.W2_t, o.Wimg, st));
                    p->launches += 3;
                } else if (o.gemm) {
                    void *d;
                    TRY(dev_alloc(p, mix_gemm_image_bytes(H, o.F), &d)); o.Wimg = (uint8_t *)d;
                    TRY(mix_gemm_pack(H, o.F, o.Wo_t, o.W1_t, o.W2_t, o.Wimg, st));
                    p->launches += 1;
                }
                o.mma = !o.umma && !o.gemm && p->use_mma && mix_mma_supported(H, o.F, l);
                if (o.mma) {
                    auto pack = [&](const float *Wt, int M, int K, uint4 **f) -> int {
                        for (int q = 0; q < 2; ++q) {
                            void *d;
                            TRY(dev_alloc(p, (size_t)M * K * 2, &d));
                            f[q] = (uint4 *)d;
                        }
                        TRY(frag_pack(Wt, M, K, (uint32_t *)f[0], (uint32_t *)f[1], st));
                        p->launches += 1;
                        return DWB_OK;
                    };
                    TRY(pack(o.Wo_t, 2 * H, H, o.Wo_f));
                    TRY(pack(o.W1_t, o.F, H, o.W1_f));
                    TRY(pack(o.W2_t, H, o.F, o.W2_f));
                }
                o.part_off = part_off; part_off += H;
                o.cond_off = cond_off; cond_off += (int64_t)H * l;
                fc.push_back({o.prefix, H});
            } else {
                const bool up = o.kind == OP_UP;
                const int M = up ? o.Ho * o.s : o.Ho, K = up ? o.H : o.H * o.s;
                TRY(folded(p, o.prefix + "linear.conv", M, K, 1, true, &o.Wp_t, &o.bp, st));
                // DWB_POOL=mma keeps the pools on the mma.sync kernels
                static const bool pool_mma_only = [] { const char *e = getenv("DWB_POOL"); return e && std::string(e) == "mma"; }();
                o.pumma = p->use_mma && p->use_umma && !pool_mma_only && pool_umma_supported(o.H, o.Ho, o.s, up, o.l);
                if (o.pumma) {
                    void *d;
                    TRY(dev_alloc(p, pool_umma_image_bytes(o.H, o.Ho, o.s, up), &d)); o.Wp_img = (uint8_t *)d;
                    TRY(pool_umma_pack(o.H, o.Ho, o.s, up, o.Wp_t, o.Wp_img, st));
                    p->launches += 1;
                }
                o.mma = p->use_mma && pool_mma_supported(o.H, o.Ho, o.s, up);
                if (o.mma) {
                    for (int q = 0; q < 2; ++q) {
                        void *d;
                        TRY(dev_alloc(p, (size_t)M * K * 2, &d));
                        o.Wp_f[q] = (uint4 *)d;
                    }
                    TRY(frag_pack(o.Wp_t, M, K, (uint32_t *)o.Wp_f[0], (uint32_t *)o.Wp_f[1], st));
                    p->launches += 1;
                }
            }
            return DWB_OK;
        };
        rc = body();
    }
    cudaStreamSynchronize(st);
    cudaFree(khat);
    cudaFree(k64);
    if (rc != DWB_OK) return rc;
    p->cond_total = cond_off;
    TRY(stack_fc_t(p, fc, st));
    return DWB_OK;
}

static int finalize_wavenet(dwb_plan *p, cudaStream_t st) {
    const dwb_config &c = p->cfg;
    const int C = c.res_channels, S = c.skip_channels, N = c.num_res_layers;
    DWB_REQUIRE(C >= 1 && S >= 1 && N >= 1 && c.dilation_cycle >= 1, DWB_ERR_INVALID, "bad wavenet config");
    TRY(finalize_common(p, "residual_layer.", "init_conv.0.conv", C, st));
    TRY(finalize_head(p, S, st));
    std::vector<std::pair<std::string, int>> fc;
    p->wl.clear();
    int64_t cond_off = 0;
    for (int n = 0; n < N; ++n) {
        const std::string pre = "residual_layer.residual_blocks." + std::to_string(n) + ".";
        WaveLayer w{};
        w.dilation = 1 << (n % c.dilation_cycle);
        TRY(folded(p, pre + "dilated_conv_layer.conv", 2 * C, C, 3, true, &w.Wd_t, &w.bd, st));
        TRY(folded(p, pre + "res_conv", C, C, 1, true, &w.Wr_t, &w.br, st));
        TRY(folded(p, pre + "skip_conv", S, C, 1, true, &w.Ws_t, &w.bs, st));
        w.umma = p->use_mma && p->use_umma && wave_umma_supported(C, S);
        if (w.umma) {
            void *d;
            TRY(dev_alloc(p, wave_umma_image_bytes(C, S), &d)); w.Wimg = (uint8_t *)d;
            TRY(wave_umma_pack(C, S, w.Wd_t, w.Wr_t, w.Ws_t, w.Wimg, st));
            p->launches += 1;
        }
        w.mma = !w.umma && p->use_mma && wave_mma_supported(C, S);
        if (w.mma) {
            auto pack = [&](const float *Wt, int M, int K, uint4 **f) -> int {
                for (int q = 0; q < 2; ++q) {
                    void *d;
                    TRY(dev_alloc(p, (size_t)M * K * 2, &d));
                    f[q] = (uint4 *)d;
                }
                TRY(frag_pack(Wt, M, K, (uint32_t *)f[0], (uint32_t *)f[1], st));
                p->launches += 1;
                return DWB_OK;
            };
            TRY(pack(w.Wd_t, 2 * C, 3 * C, w.Wd_f));
            TRY(pack(w.Wr_t, C, C, w.Wr_f));
            TRY(pack(w.Ws_t, S, C, w.Ws_f));
        }
        w.part_off = n * C;
        w.cond_off = cond_off;
        p->wl.push_back(w);
        fc.push_back({pre, C});
    }
    TRY(stack_fc_t(p, fc, st));
    return DWB_OK;
}

// length of an op's input when the model runs on L samples instead of cfg.L (same pooling ratios)
static inline int run_len(const dwb_plan *p, int l_cfg, int L) { return (int)((int64_t)l_cfg * L / p->cfg.L); }

static int ensure_workspace(dwb_plan *p, int B, int L) {
    const dwb_config &c = p->cfg;
    const bool sash = c.model == DWB_MODEL_SASHIMI;
    if (p->ws_B == B && p->ws_L == L) return DWB_OK;
    DWB_CUDA(cudaDeviceSynchronize());
    for (void *d : p->ws_owned) cudaFree(d);
    p->ws_owned.clear();
    p->bufs.clear();
    p->stat_bufs.clear();
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    void *d;
    if (sash) {
        const size_t act = (size_t)B * c.d_model * L * sizeof(float);   // H*l never exceeds d_model*L when expand <= pool
        size_t maxact = act;
        for (auto &o : p->ops) {
            const size_t r = (size_t)run_len(p, o.l, L);
            maxact = std::max(maxact, (size_t)B * o.H * r * sizeof(float));
            if (o.kind != OP_BLOCK) maxact = std::max(maxact, (size_t)B * o.Ho * (o.kind == OP_DOWN ? r / o.s : r * o.s) * sizeof(float));
        }
        for (int i = 0; i < p->n_bufs; ++i) {
            TRY(dev_alloc(p, maxact, &d, true)); p->bufs.push_back((float *)d);
            TRY(dev_alloc(p, (size_t)B * L * 2 * sizeof(float), &d, true)); p->stat_bufs.push_back((float *)d);
        }
        TRY(dev_alloc(p, maxact, &d, true)); p->g_buf = (float *)d;
        size_t hid = 0;
        for (auto &o : p->ops)
            if (o.kind == OP_BLOCK && o.gemm) hid = std::max(hid, (size_t)B * o.F * run_len(p, o.l, L) * sizeof(float));
        p->hid_buf = nullptr;
        if (hid) { TRY(dev_alloc(p, hid, &d, true)); p->hid_buf = (float *)d; }
    } else {
        for (int i = 0; i < 2; ++i) {
            TRY(dev_alloc(p, (size_t)B * c.res_channels * L * sizeof(float), &d, true));
            p->bufs.push_back((float *)d);
        }
        TRY(dev_alloc(p, (size_t)B * c.skip_channels * L * sizeof(float), &d, true)); p->skip_acc = (float *)d;
    }
    TRY(dev_alloc(p, (size_t)B * c.embed_out * sizeof(float), &d, true)); p->emb_buf = (float *)d;
    TRY(dev_alloc(p, (size_t)B * p->Mtot * sizeof(float), &d, true)); p->part_buf = (float *)d;
    TRY(dev_alloc(p, (size_t)B * L * sizeof(float), &d, true)); p->x_cur = (float *)d;
    TRY(dev_alloc(p, (size_t)(p->Mtot + 4) * sizeof(float), &d, true)); p->rec_cur = (float *)d;
    TRY(dev_alloc(p, sizeof(float *), &d, true)); p->noise_base = (const float **)d;
    p->ws_B = B;
    p->ws_L = L;
    return DWB_OK;
}

// per-category device timing of one eager (non-graph) network evaluation: an event after every
// launch; the interval since the previous event is charged to the launch's category
struct Prof {
    std::vector<cudaEvent_t> ev;
    std::vector<int> cat;
    cudaStream_t st;
    int begin() {
        cudaEvent_t e;
        DWB_CUDA(cudaEventCreate(&e));
        DWB_CUDA(cudaEventRecord(e, st));
        ev.push_back(e);
        cat.push_back(-1);
        return DWB_OK;
    }
    int mark(int c) {
        cudaEvent_t e;
        DWB_CUDA(cudaEventCreate(&e));
        DWB_CUDA(cudaEventRecord(e, st));
        ev.push_back(e);
        cat.push_back(c);
        return DWB_OK;
    }
};
#define PROF(c) do { if (prof) TRY(prof->mark(c)); } while (0)

static void free_len_tabs(std::vector<dwb_plan::LenTab> &v) {
    for (auto &t : v) { cudaFree(t.kf); cudaFree(t.kf_c); cudaFree(t.kf_a); }
    v.clear();
}

// tables for run length L != cfg.L, built from the cached fp32 kernels (the reference convolves with fp32 kernels too).
// Not during stream capture; synchronises when it builds.
static int ensure_len_tabs(dwb_plan *p, int L, cudaStream_t st) {
    if (p->cfg.model != DWB_MODEL_SASHIMI || L == p->cfg.L) return DWB_OK;
    auto touch = [&]() {
        auto &lru = p->len_lru;
        lru.erase(std::remove(lru.begin(), lru.end(), L), lru.end());
        lru.push_back(L);
    };
    if (p->len_tabs.count(L)) { touch(); return DWB_OK; }
    DWB_CUDA(cudaDeviceSynchronize());
    while (p->len_lru.size() >= 2) {
        free_len_tabs(p->len_tabs[p->len_lru.front()]);
        p->len_tabs.erase(p->len_lru.front());
        p->len_lru.erase(p->len_lru.begin());
    }
    std::vector<dwb_plan::LenTab> tabs;
    int rc = DWB_OK;
    for (auto &o : p->ops) {
        dwb_plan::LenTab t;
        if (o.kind == OP_BLOCK && rc == DWB_OK) {
            const int r = run_len(p, o.l, L);
            const Tensor *D = find(p, o.prefix + "layer.D");
            auto make = [&](int taps, int dir, float **out) -> int {
                const int lg = fft_log2m_for(taps);
                DWB_REQUIRE(lg > 0, DWB_ERR_UNSUPPORTED, "stage length %d exceeds the in-shared-memory FFT (max %d)", taps, 1 << FFT_MAX_LOG2M);
                DWB_CUDA(cudaMalloc(out, (size_t)o.H * fft_table_floats(lg) * sizeof(float)));
                TRY(fftconv_prepare_f32(o.k32, o.l, (const float *)D->dev, o.H, taps, dir, *out, st));
                const float2 *tw;
                TRY(fft_twiddles(lg, st, &tw));
                if (dir == 0 && fft_table_mode(lg, taps) == 2) TRY(fft_pair_twiddles(lg, st, &tw));
                p->launches += 2;
                return DWB_OK;
            };
            if (r < 1) { set_error("sequence length %d leaves an empty stage", L); rc = DWB_ERR_INVALID; }
            else if (r <= o.l) rc = make(r, 0, &t.kf);
            else {
                rc = make(o.l, 1, &t.kf_c);
                if (rc == DWB_OK) rc = make(o.l, 2, &t.kf_a);
            }
        }
        tabs.push_back(t);
    }
    if (rc != DWB_OK) { free_len_tabs(tabs); return rc; }
    p->len_tabs[L] = tabs;
    touch();
    return DWB_OK;
}

static int stage_of(const dwb_plan *p, int l) {   // 0 = top stage, 1 = after first pool, ...
    int s = 0, cur = p->cfg.L;
    while (cur > l && s < p->cfg.n_pool) { cur /= p->cfg.pool[s]; ++s; }
    return s;
}

struct StepUpdate {           // fused DDPM update in the head (coefficients read from device memory), or plain eps
    const float *x = nullptr;
    const float *ctl = nullptr;                 // device: c1, sqrt(alpha), sigma, noise slot (int bits, -1 = none)
    const float *const *noise_base = nullptr;   // device: noise of slot i at *noise_base + i * B * L
};

// one network evaluation; part = (rows, Mtot) fc_t outputs with batch stride psb (0 = shared row)
static int run_network(dwb_plan *p, const float *x, const float *part, long long psb, const float *cond, int cond_batch,
                       float *out, const StepUpdate *upd, int B, int L, cudaStream_t st, Prof *prof = nullptr) {
    const dwb_config &c = p->cfg;
    HeadArgs h{};
    if (c.model == DWB_MODEL_SASHIMI) {
        TRY(init_conv_launch(x, p->init_w, p->init_b, B, c.d_model, L, p->bufs[0], p->stat_bufs[0], st));
        p->launches += 1;
        PROF(DWB_PROF_INIT);
        int last = 0;
        const std::vector<dwb_plan::LenTab> *lt = nullptr;
        if (L != c.L) {
            auto it = p->len_tabs.find(L);
            DWB_REQUIRE(it != p->len_tabs.end(), DWB_ERR_STATE, "tables for length %d were not prepared", L);
            lt = &it->second;
        }
        int64_t cond_off = 0;
        size_t oi = 0;
        for (auto &o : p->ops) {
            const int r = run_len(p, o.l, L);                  // = o.l when L == cfg.L
            if (o.kind == OP_BLOCK) {
                float *scratch = p->bufs[o.out_buf];           // the block's output buffer is free scratch here
                if (lt && r > o.l)
                    TRY(fftconv_ols_launch(p->bufs[o.in_buf], p->stat_bufs[o.in_buf], part + o.part_off, psb, o.ln1_m, o.ln1_s,
                                           (*lt)[oi].kf_c, (*lt)[oi].kf_a, p->g_buf, scratch, B, o.H, r, o.l, st));
                else
                    TRY(fftconv_launch(p->bufs[o.in_buf], p->stat_bufs[o.in_buf], part + o.part_off, psb, o.ln1_m, o.ln1_s,
                                       lt ? (*lt)[oi].kf : o.kf, p->g_buf, B, o.H, r, st, scratch));
                PROF(DWB_PROF_FFTCONV0 + std::min(stage_of(p, o.l), 3));
                MixArgs a{};
                a.g = p->g_buf; a.x = p->bufs[o.in_buf];
                a.skip = o.skip_buf >= 0 ? p->bufs[o.skip_buf] : nullptr;
                a.cond = cond ? cond + (size_t)cond_batch * cond_off : nullptr;
                cond_off += (int64_t)o.H * r;
                a.cond_stride_b = cond_batch > 1 ? 1 : 0;
                a.Wo_t = o.Wo_t; a.bo = o.bo; a.W1_t = o.W1_t; a.b1 = o.b1; a.W2_t = o.W2_t; a.b2 = o.b2;
                a.ln2_m = o.ln2_m; a.ln2_s = o.ln2_s;
                a.out = p->bufs[o.out_buf]; a.stats_out = p->stat_bufs[o.out_buf];
                a.H = o.H; a.F = o.F; a.l = r;
                a.Wo_fh = o.Wo_f[0]; a.Wo_fl = o.Wo_f[1]; a.W1_fh = o.W1_f[0]; a.W1_fl = o.W1_f[1];
                a.W2_fh = o.W2_f[0]; a.W2_fl = o.W2_f[1];
                a.Wimg = o.Wimg; a.bimg = o.bimg;
                TRY(o.umma ? mix_umma_launch(a, B, st)
                           : (o.gemm2 ? mix_gemm2_launch(a, p->hid_buf, B, st)
                                      : (o.gemm ? mix_gemm_launch(a, p->hid_buf, B, st) : (o.mma ? mix_mma_launch(a, B, st) : mix_launch(a, B, st)))));
                p->launches += ((lt && r > o.l) ? 3 : 2) + (o.gemm2 ? 3 : (o.gemm ? 4 : 0));
                PROF(DWB_PROF_MIX0 + std::min(stage_of(p, o.l), 3));
            } else {
                PoolArgs a{};
                a.x = p->bufs[o.in_buf];
                a.skip = o.skip_buf >= 0 ? p->bufs[o.skip_buf] : nullptr;
                a.W_t = o.Wp_t; a.bias = o.bp;
                a.out = p->bufs[o.out_buf]; a.stats_out = p->stat_bufs[o.out_buf];
                a.Hi = o.H; a.Ho = o.Ho; a.s = o.s; a.li = r;
                a.W_fh = o.Wp_f[0]; a.W_fl = o.Wp_f[1];
                if (o.pumma && pool_umma_supported(o.H, o.Ho, o.s, o.kind == OP_UP, r)) TRY(pool_umma_launch(a, o.Wp_img, o.kind == OP_UP, B, st));
                else if (o.mma) TRY(o.kind == OP_DOWN ? down_pool_mma_launch(a, B, st) : up_pool_mma_launch(a, B, st));
                else TRY(o.kind == OP_DOWN ? down_pool_launch(a, B, st) : up_pool_launch(a, B, st));
                p->launches += 1;
                PROF(DWB_PROF_POOL);
            }
            last = o.out_buf;
            ++oi;
        }
        h.x = p->bufs[last]; h.stats = p->stat_bufs[last];
        h.ln_m = p->norm_m; h.ln_s = p->norm_s; h.prescale = 1.f;
        h.C = c.d_model;
    } else {
        const int C = c.res_channels, S = c.skip_channels, N = c.num_res_layers;
        TRY(init_conv_launch(x, p->init_w, p->init_b, B, C, L, p->bufs[0], nullptr, st));
        p->launches += 1;
        PROF(DWB_PROF_INIT);
        int cur = 0;
        for (int n = 0; n < N; ++n) {
            const WaveLayer &w = p->wl[n];
            WaveBlockArgs a{};
            a.h = p->bufs[cur]; a.h_out = p->bufs[cur ^ 1];
            a.part_t = part + w.part_off; a.part_stride_b = psb;
            a.cond = cond ? cond + (size_t)cond_batch * n * 2 * C * L : nullptr;
            a.cond_stride_b = cond_batch > 1 ? 1 : 0;
            a.Wd_t = w.Wd_t; a.bd = w.bd; a.Wr_t = w.Wr_t; a.br = w.br; a.Ws_t = w.Ws_t; a.bs = w.bs;
            a.skip = p->skip_acc; a.first = n == 0;
            a.C = C; a.S = S; a.L = L; a.dilation = w.dilation;
            if (w.umma) {
                a.Wimg = w.Wimg;
                TRY(wave_block_umma_launch(a, B, st));
            } else if (w.mma) {
                a.Wd_fh = w.Wd_f[0]; a.Wd_fl = w.Wd_f[1]; a.Wr_fh = w.Wr_f[0]; a.Wr_fl = w.Wr_f[1];
                a.Ws_fh = w.Ws_f[0]; a.Ws_fl = w.Ws_f[1];
                TRY(wave_block_mma_launch(a, B, st));
            } else
                TRY(wave_block_launch(a, B, st));
            p->launches += 1;
            PROF(DWB_PROF_WAVEBLOCK);
            cur ^= 1;
        }
        h.x = p->skip_acc; h.stats = nullptr;
        h.prescale = sqrtf(1.0f / (float)N);
        h.C = S;
    }
    h.Wf_t = p->Wf_t; h.bf = p->bf; h.wz = p->wz; h.bz = p->bz;
    h.out = out; h.l = L;
    if (upd) { h.upd_x = upd->x; h.ctl = upd->ctl; h.noise_base = upd->noise_base; }
    h.Wf_fh = p->Wf_f[0]; h.Wf_fl = p->Wf_f[1];
    if (p->Wf_img) TRY(head_umma_launch(h, p->Wf_img, B, st));
    else if (h.Wf_fh) TRY(head_mma_launch(h, B, st));
    else TRY(head_launch(h, B, st));
    p->launches += 1;
    PROF(DWB_PROF_HEAD);
    return DWB_OK;
}

static int check_run(dwb_plan *p, int B, int L, const float *cond, int cond_batch) {
    DWB_REQUIRE(p && p->finalized, DWB_ERR_STATE, "plan is not finalized");
    DWB_REQUIRE(B >= 1 && L >= 1, DWB_ERR_INVALID, "bad sizes B=%d L=%d", B, L);
    DWB_REQUIRE(B <= 65535, DWB_ERR_UNSUPPORTED, "batch %d > 65535", B);
    if (p->cfg.model == DWB_MODEL_SASHIMI) {
        int prod = 1;
        for (int q = 0; q < p->cfg.n_pool; ++q) prod *= p->cfg.pool[q];
        DWB_REQUIRE(L % prod == 0, DWB_ERR_INVALID, "sequence length %d is not divisible by the pooling factor %d", L, prod);
    }
    if (cond) {
        DWB_REQUIRE(!p->cfg.unconditional, DWB_ERR_INVALID, "conditioning passed to an unconditional model");
        DWB_REQUIRE(cond_batch == 1 || cond_batch == B, DWB_ERR_INVALID, "cond_batch must be 1 or B");
    } else {
        DWB_REQUIRE(p->cfg.unconditional, DWB_ERR_INVALID, "conditional model called without conditioning features");
    }
    return DWB_OK;
}

}  // namespace dwb

// =========================================================================================
extern "C" {

int dwb_version(void) { return DWB_VERSION; }
const char *dwb_last_error(void) { return g_err; }

int dwb_device_count(int *count) {
    DWB_REQUIRE(count, DWB_ERR_INVALID, "null");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    return DWB_OK;
}

int dwb_plan_create(const dwb_config *cfg, int device, dwb_plan **out) {
    DWB_REQUIRE(cfg && out, DWB_ERR_INVALID, "dwb_plan_create: null pointer");
    DWB_REQUIRE(cfg->model == DWB_MODEL_WAVENET || cfg->model == DWB_MODEL_SASHIMI, DWB_ERR_INVALID, "unknown model %d", cfg->model);
    DWB_REQUIRE(cfg->embed_in >= 4 && cfg->embed_in % 2 == 0 && cfg->embed_mid >= 1 && cfg->embed_out >= 1, DWB_ERR_INVALID,
                "bad embedding dims %d/%d/%d", cfg->embed_in, cfg->embed_mid, cfg->embed_out);
    int n = 0;
    TRY(dwb_device_count(&n));
    DWB_REQUIRE(n > 0, DWB_ERR_CUDA, "no CUDA device: libdwb has no CPU fallback");
    DWB_REQUIRE(device >= 0 && device < n, DWB_ERR_INVALID, "device %d out of range (have %d)", device, n);
    DWB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DWB_CUDA(cudaGetDeviceProperties(&prop, device));
    DWB_REQUIRE(prop.major == 10, DWB_ERR_UNSUPPORTED, "device %d is sm_%d%d; libdwb is built for sm_100a only", device, prop.major, prop.minor);
    dwb_plan *p = new dwb_plan();
    if (const char *e = getenv("DWB_MIX")) {
        p->use_mma = std::string(e) != "simt";
        p->use_umma = std::string(e) != "mma";
    }
    p->cfg = *cfg;
    p->device = device;
    *out = p;
    return DWB_OK;
}

int dwb_plan_destroy(dwb_plan *p) {
    if (!p) return DWB_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
    for (auto &kv : p->tensors) cudaFree(kv.second.dev);
    for (void *d : p->owned) cudaFree(d);
    for (void *d : p->ws_owned) cudaFree(d);
    for (auto &kv : p->len_tabs) free_len_tabs(kv.second);
    cudaFree(p->table_emb); cudaFree(p->table_part); cudaFree(p->table_t); cudaFree(p->step_tab);
    delete p;
    return DWB_OK;
}

int dwb_plan_set_tensor(dwb_plan *p, const char *name, const void *data, int dtype, const int64_t *shape, int ndim,
                        int on_device, void *stream) {
    DWB_REQUIRE(p && name && data, DWB_ERR_INVALID, "dwb_plan_set_tensor: null pointer");
    DWB_REQUIRE(!p->finalized, DWB_ERR_STATE, "plan already finalized");
    DWB_REQUIRE(dtype == DWB_F32 || dtype == DWB_I64, DWB_ERR_INVALID, "`%s`: unsupported dtype %d", name, dtype);
    DWB_REQUIRE(ndim >= 0 && ndim <= 8, DWB_ERR_INVALID, "`%s`: bad ndim %d", name, ndim);
    DWB_CUDA(cudaSetDevice(p->device));
    Tensor t;
    t.dtype = dtype;
    for (int i = 0; i < ndim; ++i) {
        DWB_REQUIRE(shape[i] >= 0, DWB_ERR_INVALID, "`%s`: negative dim", name);
        t.shape.push_back(shape[i]);
    }
    t.bytes = (size_t)t.numel() * (dtype == DWB_F32 ? 4 : 8);
    DWB_CUDA(cudaMalloc(&t.dev, t.bytes ? t.bytes : 4));
    cudaError_t e = cudaMemcpyAsync(t.dev, data, t.bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                    (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) { cudaFree(t.dev); return cuda_fail(e, "copy tensor", __FILE__, __LINE__); }
    auto it = p->tensors.find(name);
    if (it != p->tensors.end()) { cudaFree(it->second.dev); p->tensors.erase(it); }
    p->tensors[name] = t;
    return DWB_OK;
}

// conditional models: fold the per-block mel front end (skipped when the host did not supply those weights)
static int finalize_cond(dwb_plan *p, cudaStream_t st) {
    p->cond_blocks.clear();
    if (p->cfg.unconditional) return DWB_OK;
    std::vector<std::pair<std::string, dwb_plan::CondBlock>> todo;
    if (p->cfg.model == DWB_MODEL_SASHIMI) {
        for (auto &o : p->ops)
            if (o.kind == OP_BLOCK) {
                dwb_plan::CondBlock cb{};
                cb.Hc = o.H; cb.l = o.l; cb.off = o.cond_off;
                todo.push_back({o.prefix, cb});
            }
    } else {
        for (size_t i = 0; i < p->wl.size(); ++i) {
            dwb_plan::CondBlock cb{};
            cb.Hc = 2 * p->cfg.res_channels; cb.l = 0; cb.off = (int64_t)i;      // layer index: offset = i * 2C * L at call time
            todo.push_back({"residual_layer.residual_blocks." + std::to_string(i) + ".", cb});
        }
    }
    if (todo.empty() || !find(p, todo[0].first + "upsample_conv2d.0.weight_v")) return DWB_OK;
    for (auto &t : todo) {
        dwb_plan::CondBlock cb = t.second;
        for (int i = 0; i < 2; ++i) {
            const std::string pre = t.first + "upsample_conv2d." + std::to_string(i);
            const Tensor *v = find(p, pre + ".weight_v");
            DWB_REQUIRE(v && v->shape.size() == 4 && v->shape[0] == 1 && v->shape[1] == 1 && v->shape[2] == 3 && v->shape[3] % 2 == 0,
                        DWB_ERR_INVALID, "`%s.weight_v` must be (1,1,3,2s)", pre.c_str());
            cb.s[i] = (int)(v->shape[3] / 2);
            DWB_REQUIRE(cb.s[i] % 2 == 0, DWB_ERR_UNSUPPORTED,
                        "`%s`: odd upsampling stride %d (ConvTranspose2d with padding s/2 then yields W*s + 1 columns; not built)",
                        pre.c_str(), cb.s[i]);
            TRY(folded(p, pre, 1, 1, 3 * 2 * cb.s[i], true, &cb.w[i], &cb.b[i], st));
        }
        TRY(folded(p, t.first + "mel_conv.conv", cb.Hc, p->cfg.mel_bands, 1, true, &cb.Wm_t, &cb.bm, st));
        p->cond_blocks.push_back(cb);
    }
    return DWB_OK;
}

int dwb_plan_finalize(dwb_plan *p, void *stream) {
    DWB_REQUIRE(p, DWB_ERR_INVALID, "null plan");
    DWB_REQUIRE(!p->finalized, DWB_ERR_STATE, "plan already finalized");
    DWB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = p->cfg.model == DWB_MODEL_SASHIMI ? finalize_sashimi(p, st) : finalize_wavenet(p, st);
    if (rc == DWB_OK) rc = finalize_cond(p, st);
    if (rc != DWB_OK) return rc;
    DWB_CUDA(cudaStreamSynchronize(st));
    p->finalized = true;
    return DWB_OK;
}

int dwb_forward(dwb_plan *p, const float *x, const float *t, const float *cond, int cond_batch, float *eps, int B, int L,
                void *stream) {
    DWB_REQUIRE(p && x && t && eps, DWB_ERR_INVALID, "dwb_forward: null pointer");
    TRY(check_run(p, B, L, cond, cond_batch));
    DWB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    TRY(ensure_workspace(p, B, L));
    TRY(ensure_len_tabs(p, L, st));
    const dwb_config &c = p->cfg;
    TRY(embed_launch(t, B, c.embed_in, c.embed_mid, c.embed_out, p->eW1, p->eb1, p->eW2, p->eb2, p->Wt_all, p->bt_all,
                     p->Mtot, p->emb_buf, p->part_buf, st));
    p->launches += 2;
    return run_network(p, x, p->part_buf, p->Mtot, cond, cond_batch, eps, nullptr, B, L, st);
}

// (T, Mtot + 4) per-step records; rebuilt when T or the schedule changes.  Synchronises.
static int ensure_step_table(dwb_plan *p, const float *coef_host, int T, cudaStream_t st) {
    const dwb_config &c = p->cfg;
    if (p->table_T == T && p->tab_coef.size() == (size_t)3 * T && std::equal(coef_host, coef_host + 3 * T, p->tab_coef.begin()))
        return DWB_OK;
    DWB_CUDA(cudaStreamSynchronize(st));
    const size_t rec = (size_t)p->Mtot + 4;
    if (p->table_T != T) {
        cudaFree(p->table_emb); cudaFree(p->table_part); cudaFree(p->table_t); cudaFree(p->step_tab);
        p->table_emb = p->table_part = p->table_t = p->step_tab = nullptr;
        p->table_T = 0;
        DWB_CUDA(cudaMalloc(&p->table_emb, (size_t)T * c.embed_out * sizeof(float)));
        DWB_CUDA(cudaMalloc(&p->table_part, (size_t)T * p->Mtot * sizeof(float)));
        DWB_CUDA(cudaMalloc(&p->table_t, (size_t)T * sizeof(float)));
        DWB_CUDA(cudaMalloc(&p->step_tab, (size_t)T * rec * sizeof(float)));
        std::vector<float> ts(T);
        for (int i = 0; i < T; ++i) ts[i] = (float)i;            // same step for the whole batch (generate.py:50)
        DWB_CUDA(cudaMemcpy(p->table_t, ts.data(), T * sizeof(float), cudaMemcpyHostToDevice));
        TRY(embed_launch(p->table_t, T, c.embed_in, c.embed_mid, c.embed_out, p->eW1, p->eb1, p->eW2, p->eb2, p->Wt_all,
                         p->bt_all, p->Mtot, p->table_emb, p->table_part, st));
        p->launches += 2;
        DWB_CUDA(cudaMemcpy2DAsync(p->step_tab, rec * sizeof(float), p->table_part, (size_t)p->Mtot * sizeof(float),
                                   (size_t)p->Mtot * sizeof(float), T, cudaMemcpyDeviceToDevice, st));
        DWB_CUDA(cudaStreamSynchronize(st));
        p->table_T = T;
    }
    std::vector<float> ctl((size_t)4 * T);
    for (int t = 0; t < T; ++t) {
        ctl[4 * t + 0] = coef_host[t];
        ctl[4 * t + 1] = coef_host[T + t];
        ctl[4 * t + 2] = coef_host[2 * T + t];
        const int slot = t > 0 ? T - 1 - t : -1;                 // draw i is used at step T-1-i; no draw at t = 0
        memcpy(&ctl[4 * t + 3], &slot, sizeof(int));
    }
    DWB_CUDA(cudaMemcpy2D(p->step_tab + p->Mtot, rec * sizeof(float), ctl.data(), 4 * sizeof(float), 4 * sizeof(float), T,
                          cudaMemcpyHostToDevice));
    p->tab_coef.assign(coef_host, coef_host + 3 * T);
    return DWB_OK;
}

// steps t_start, t_start-1, ..., t_start-n_steps+1 of the reverse loop on p->x_cur
static int run_steps(dwb_plan *p, const float *noise, const float *cond, int cond_batch, const float *coef_host, int T,
                     int t_start, int n_steps, int B, int L, int use_graph, cudaStream_t st) {
    TRY(ensure_step_table(p, coef_host, T, st));
    const size_t BL = (size_t)B * L, rec = (size_t)p->Mtot + 4;
    // slot i of the whole run lives at base + i*BL; the caller's pointer is slot T-1-t_start (never dereferenced below it)
    const float *base = noise ? noise - (ptrdiff_t)(T - 1 - t_start) * (ptrdiff_t)BL : nullptr;
    DWB_CUDA(cudaMemcpyAsync(p->noise_base, &base, sizeof(base), cudaMemcpyHostToDevice, st));   // pageable: staged before return
    StepUpdate u;
    u.x = p->x_cur; u.ctl = p->rec_cur + p->Mtot; u.noise_base = p->noise_base;
    auto one_step = [&](cudaStream_t s) { return run_network(p, p->x_cur, p->rec_cur, 0, cond, cond_batch, p->x_cur, &u, B, L, s); };
    if (use_graph) {
        StepKey key;
        key.cond = cond; key.cond_batch = cond_batch; key.B = B; key.L = L;
        if (!p->graph_exec || !(key == p->graph_key)) {
            if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
            if (!p->cap_stream) DWB_CUDA(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
            const int64_t before = p->launches;
            DWB_CUDA(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal));
            int rc = one_step(p->cap_stream);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(p->cap_stream, &graph);
            if (rc != DWB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
            p->graph_nodes = p->launches - before;
            p->launches = before;
            e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { p->graph_exec = nullptr; return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
            p->graph_key = key;
        }
    }
    for (int i = 0; i < n_steps; ++i) {
        const int t = t_start - i;
        DWB_CUDA(cudaMemcpyAsync(p->rec_cur, p->step_tab + (size_t)t * rec, rec * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (use_graph) {
            DWB_CUDA(cudaGraphLaunch(p->graph_exec, st));
            p->launches += p->graph_nodes;
        } else
            TRY(one_step(st));
    }
    return DWB_OK;
}

int dwb_sample(dwb_plan *p, const float *x_T, const float *noise, const float *cond, int cond_batch,
               const float *coef_host, int T, float *out, int B, int L, int use_graph, void *stream) {
    DWB_REQUIRE(p && x_T && out && coef_host, DWB_ERR_INVALID, "dwb_sample: null pointer");
    DWB_REQUIRE(T >= 1, DWB_ERR_INVALID, "T=%d", T);
    DWB_REQUIRE(noise || T == 1, DWB_ERR_INVALID, "noise is required for T > 1");
    TRY(check_run(p, B, L, cond, cond_batch));
    DWB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    TRY(ensure_workspace(p, B, L));
    TRY(ensure_len_tabs(p, L, st));
    const size_t bytes = (size_t)B * L * sizeof(float);
    DWB_CUDA(cudaMemcpyAsync(p->x_cur, x_T, bytes, cudaMemcpyDeviceToDevice, st));
    TRY(run_steps(p, noise, cond, cond_batch, coef_host, T, T - 1, T, B, L, use_graph, st));
    DWB_CUDA(cudaMemcpyAsync(out, p->x_cur, bytes, cudaMemcpyDeviceToDevice, st));
    return DWB_OK;
}

int dwb_sample_steps(dwb_plan *p, float *x, const float *noise, const float *cond, int cond_batch, const float *coef_host,
                     int T, int t_start, int n_steps, int B, int L, int use_graph, void *stream) {
    DWB_REQUIRE(p && x && coef_host, DWB_ERR_INVALID, "dwb_sample_steps: null pointer");
    DWB_REQUIRE(T >= 1 && n_steps >= 1 && t_start < T && t_start - n_steps + 1 >= 0, DWB_ERR_INVALID,
                "dwb_sample_steps: steps [%d, %d] outside [0, %d)", t_start - n_steps + 1, t_start, T);
    DWB_REQUIRE(noise || (t_start == 0 && n_steps == 1), DWB_ERR_INVALID, "noise is required for steps t > 0");
    TRY(check_run(p, B, L, cond, cond_batch));
    DWB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    TRY(ensure_workspace(p, B, L));
    TRY(ensure_len_tabs(p, L, st));
    const size_t bytes = (size_t)B * L * sizeof(float);
    DWB_CUDA(cudaMemcpyAsync(p->x_cur, x, bytes, cudaMemcpyDeviceToDevice, st));
    TRY(run_steps(p, noise, cond, cond_batch, coef_host, T, t_start, n_steps, B, L, use_graph, st));
    DWB_CUDA(cudaMemcpyAsync(x, p->x_cur, bytes, cudaMemcpyDeviceToDevice, st));
    return DWB_OK;
}

int dwb_plan_profile(dwb_plan *p, const float *x, const float *t, const float *cond, int cond_batch, float *eps, int B,
                     int L, int iters, double *ms, int64_t *counts, void *stream) {
    DWB_REQUIRE(p && x && t && eps && ms && counts && iters >= 1, DWB_ERR_INVALID, "dwb_plan_profile: bad arguments");
    TRY(check_run(p, B, L, cond, cond_batch));
    DWB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    TRY(ensure_workspace(p, B, L));
    TRY(ensure_len_tabs(p, L, st));
    const dwb_config &c = p->cfg;
    for (int i = 0; i < DWB_PROF_NCAT; ++i) { ms[i] = 0; counts[i] = 0; }
    for (int it = 0; it < iters; ++it) {
        Prof prof;
        prof.st = st;
        int rc = prof.begin();
        if (rc == DWB_OK) rc = embed_launch(t, B, c.embed_in, c.embed_mid, c.embed_out, p->eW1, p->eb1, p->eW2, p->eb2,
                                            p->Wt_all, p->bt_all, p->Mtot, p->emb_buf, p->part_buf, st);
        if (rc == DWB_OK) { p->launches += 2; rc = prof.mark(DWB_PROF_EMBED); }
        if (rc == DWB_OK) rc = run_network(p, x, p->part_buf, p->Mtot, cond, cond_batch, eps, nullptr, B, L, st, &prof);
        cudaError_t e = cudaStreamSynchronize(st);
        if (rc == DWB_OK && e == cudaSuccess)
            for (size_t i = 1; i < prof.ev.size(); ++i) {
                float dt = 0.f;
                cudaEventElapsedTime(&dt, prof.ev[i - 1], prof.ev[i]);
                ms[prof.cat[i]] += dt;
                counts[prof.cat[i]] += 1;
            }
        for (auto ev : prof.ev) cudaEventDestroy(ev);
        if (rc != DWB_OK) return rc;
        if (e != cudaSuccess) return cuda_fail(e, "profile sync", __FILE__, __LINE__);
    }
    return DWB_OK;
}

int dwb_plan_cond_layout(dwb_plan *p, int L, int *n_blocks, int *channels, int *lengths, int64_t *offsets) {
    DWB_REQUIRE(p && p->finalized && n_blocks, DWB_ERR_STATE, "plan is not finalized");
    int n = 0;
    if (p->cfg.model == DWB_MODEL_SASHIMI) {
        DWB_REQUIRE(L >= 1, DWB_ERR_INVALID, "dwb_plan_cond_layout: L=%d", L);
        int64_t off = 0;
        for (auto &o : p->ops)
            if (o.kind == OP_BLOCK) {
                const int r = run_len(p, o.l, L);
                if (channels) { channels[n] = o.H; lengths[n] = r; offsets[n] = off; }
                off += (int64_t)o.H * r;
                ++n;
            }
    } else {
        for (size_t i = 0; i < p->wl.size(); ++i) {
            if (channels) { channels[n] = 2 * p->cfg.res_channels; lengths[n] = L; offsets[n] = (int64_t)i * 2 * p->cfg.res_channels * L; }
            ++n;
        }
    }
    *n_blocks = n;
    return DWB_OK;
}

int dwb_plan_cond_features(dwb_plan *p, const float *mel, int cond_batch, int frames, int L, float *out, void *stream) {
    DWB_REQUIRE(p && p->finalized, DWB_ERR_STATE, "plan is not finalized");
    DWB_REQUIRE(mel && out && cond_batch >= 1 && frames >= 1 && L >= 1, DWB_ERR_INVALID, "dwb_plan_cond_features: bad arguments");
    DWB_REQUIRE(!p->cfg.unconditional, DWB_ERR_STATE, "unconditional model has no conditioning path");
    DWB_REQUIRE(!p->cond_blocks.empty(), DWB_ERR_MISSING, "upsample_conv2d / mel_conv weights were not supplied to the plan");
    DWB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int F = p->cfg.mel_bands;
    int w1max = 0, w2max = 0;
    for (auto &cb : p->cond_blocks) {
        w1max = std::max(w1max, frames * cb.s[0]);
        w2max = std::max(w2max, frames * cb.s[0] * cb.s[1]);
    }
    float *u1 = nullptr, *u2 = nullptr;
    DWB_CUDA(cudaMalloc(&u1, (size_t)cond_batch * F * w1max * sizeof(float)));
    cudaError_t e = cudaMalloc(&u2, (size_t)cond_batch * F * w2max * sizeof(float));
    if (e != cudaSuccess) { cudaFree(u1); return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
    int rc = DWB_OK;
    int64_t run_off = 0;
    for (auto &cb : p->cond_blocks) {
        const int l = cb.l ? run_len(p, cb.l, L) : L, W1 = frames * cb.s[0], W2 = W1 * cb.s[1];
        if (W2 < l) {
            set_error("upsampled mel has %d samples < %d", W2, l);
            rc = DWB_ERR_INVALID;
            break;
        }
        if ((rc = mel_upsample_launch(mel, cond_batch, F, frames, cb.s[0], cb.w[0], cb.b[0], u1, st)) != DWB_OK) break;
        if ((rc = mel_upsample_launch(u1, cond_batch, F, W1, cb.s[1], cb.w[1], cb.b[1], u2, st)) != DWB_OK) break;
        const int64_t off = cb.l ? run_off : cb.off * (int64_t)cb.Hc * L;
        run_off += (int64_t)cb.Hc * l;
        if ((rc = mel_conv_launch(u2, cond_batch, F, W2, cb.Wm_t, cb.bm, cb.Hc, l, out + (size_t)cond_batch * off, st)) != DWB_OK) break;
        p->launches += 3;
    }
    cudaError_t es = cudaStreamSynchronize(st);
    cudaFree(u1);
    cudaFree(u2);
    if (rc != DWB_OK) return rc;
    if (es != cudaSuccess) return cuda_fail(es, "dwb_plan_cond_features", __FILE__, __LINE__);
    return DWB_OK;
}

int dwb_plan_launch_count(dwb_plan *p, int64_t *count) {
    DWB_REQUIRE(p && count, DWB_ERR_INVALID, "null");
    *count = p->launches;
    return DWB_OK;
}

int dwb_plan_s4_blocks(dwb_plan *p, int *n_blocks) {
    DWB_REQUIRE(p && p->finalized && n_blocks, DWB_ERR_STATE, "plan is not finalized");
    int n = 0;
    for (auto &o : p->ops) n += o.kind == OP_BLOCK;
    *n_blocks = n;
    return DWB_OK;
}

int dwb_plan_s4_kernel(dwb_plan *p, int block, float *k_out, int64_t capacity, int *H, int *l) {
    DWB_REQUIRE(p && p->finalized, DWB_ERR_STATE, "plan is not finalized");
    int n = 0;
    for (auto &o : p->ops)
        if (o.kind == OP_BLOCK) {
            if (n == block) {
                if (H) *H = o.H;
                if (l) *l = o.l;
                if (k_out) {
                    DWB_REQUIRE(capacity >= (int64_t)2 * o.H * o.l, DWB_ERR_INVALID, "k_out too small");
                    DWB_CUDA(cudaMemcpy(k_out, o.k32, (size_t)2 * o.H * o.l * sizeof(float), cudaMemcpyDeviceToDevice));
                }
                return DWB_OK;
            }
            ++n;
        }
    set_error("block %d out of range", block);
    return DWB_ERR_INVALID;
}

int dwb_plan_mix_block(dwb_plan *p, int block, int exact, const float *g, const float *x, const float *skip, float *out,
                       float *stats_out, int B, void *stream) {
    DWB_REQUIRE(p && p->finalized, DWB_ERR_STATE, "plan is not finalized");
    DWB_REQUIRE(g && x && out && stats_out && B >= 1 && B <= 65535, DWB_ERR_INVALID, "dwb_plan_mix_block: bad arguments");
    DWB_CUDA(cudaSetDevice(p->device));
    int n = 0;
    for (auto &o : p->ops)
        if (o.kind == OP_BLOCK) {
            if (n++ != block) continue;
            MixArgs a{};
            a.g = g; a.x = x; a.skip = skip; a.cond = nullptr; a.cond_stride_b = 0;
            a.Wo_t = o.Wo_t; a.bo = o.bo; a.W1_t = o.W1_t; a.b1 = o.b1; a.W2_t = o.W2_t; a.b2 = o.b2;
            a.ln2_m = o.ln2_m; a.ln2_s = o.ln2_s; a.out = out; a.stats_out = stats_out;
            a.H = o.H; a.F = o.F; a.l = o.l;
            a.Wo_fh = o.Wo_f[0]; a.Wo_fl = o.Wo_f[1]; a.W1_fh = o.W1_f[0]; a.W1_fl = o.W1_f[1];
            a.W2_fh = o.W2_f[0]; a.W2_fl = o.W2_f[1];
            a.Wimg = o.Wimg; a.bimg = o.bimg;
            p->launches += 1;
            if (exact) return mix_launch(a, B, (cudaStream_t)stream);
            if (o.gemm) {
                float *hid = nullptr;
                DWB_CUDA(cudaMalloc(&hid, (size_t)B * o.F * o.l * sizeof(float)));
                int rc = o.gemm2 ? mix_gemm2_launch(a, hid, B, (cudaStream_t)stream) : mix_gemm_launch(a, hid, B, (cudaStream_t)stream);
                cudaStreamSynchronize((cudaStream_t)stream);
                cudaFree(hid);
                return rc;
            }
            return o.umma ? mix_umma_launch(a, B, (cudaStream_t)stream)
                          : (o.mma ? mix_mma_launch(a, B, (cudaStream_t)stream) : mix_launch(a, B, (cudaStream_t)stream));
        }
    set_error("block %d out of range", block);
    return DWB_ERR_INVALID;
}

/* debug (declared in dwb.h under "debug"): tcgen05 mixing of `block` with per-CTA phase timestamps */
int dwb_debug_mix_trace(dwb_plan *p, int block, const float *g, const float *x, float *out, float *stats_out, int B,
                        long long *trace, void *stream) {
    DWB_REQUIRE(p && p->finalized, DWB_ERR_STATE, "plan is not finalized");
    int n = 0;
    for (auto &o : p->ops)
        if (o.kind == OP_BLOCK) {
            if (n++ != block) continue;
            DWB_REQUIRE(o.umma, DWB_ERR_UNSUPPORTED, "block %d is not on the fused tcgen05 path", block);
            MixArgs a{};
            a.g = g; a.x = x; a.bo = o.bo; a.b1 = o.b1; a.b2 = o.b2;
            a.ln2_m = o.ln2_m; a.ln2_s = o.ln2_s; a.out = out; a.stats_out = stats_out;
            a.H = o.H; a.F = o.F; a.l = o.l; a.Wimg = o.Wimg; a.bimg = o.bimg; a.trace = trace;
            return mix_umma_launch(a, B, (cudaStream_t)stream);
        }
    set_error("block %d out of range", block);
    return DWB_ERR_INVALID;
}

/* debug (declared in dwb.h under "debug"): layer `layer` of a WaveNet plan on caller tensors with per-CTA phase timestamps */
int dwb_debug_wave_trace(dwb_plan *p, int layer, const float *h, const float *part, float *h_out, float *skip, int B, int L,
                         long long *trace, void *stream) {
    DWB_REQUIRE(p && p->finalized && p->cfg.model == DWB_MODEL_WAVENET, DWB_ERR_STATE, "needs a finalized WaveNet plan");
    DWB_REQUIRE(layer >= 0 && layer < (int)p->wl.size() && p->wl[layer].umma, DWB_ERR_UNSUPPORTED, "layer %d is not on the tcgen05 path", layer);
    const WaveLayer &w = p->wl[layer];
    WaveBlockArgs a{};
    a.h = h; a.h_out = h_out; a.part_t = part; a.part_stride_b = 0; a.bd = w.bd; a.br = w.br; a.bs = w.bs;
    a.skip = skip; a.first = 0; a.C = p->cfg.res_channels; a.S = p->cfg.skip_channels; a.L = L; a.dilation = w.dilation;
    a.Wimg = w.Wimg; a.trace = trace;
    return wave_block_umma_launch(a, B, (cudaStream_t)stream);
}

int dwb_plan_work(dwb_plan *p, int L, double *bytes, double *flops) {
    DWB_REQUIRE(p && p->finalized && bytes && flops, DWB_ERR_STATE, "plan is not finalized");
    const dwb_config &c = p->cfg;
    double by = 0, fl = 0;
    if (c.model == DWB_MODEL_SASHIMI) {
        // two passes per block (read x, write g | read g and x, write x') = 5*4*H*l,
        // UNet skip reads, pool in/out(+skip), init/head I/O.  GEMM flops 12 H^2 l (ff = 2) + FFT flops.
        by += 4.0 * L + 4.0 * c.d_model * L;
        for (auto &o : p->ops) {
            if (o.kind == OP_BLOCK) {
                const double Hl = (double)o.H * o.l, n = 2.0 * (1 << fft_log2m_for(o.l));
                by += 20.0 * Hl + (o.skip_buf >= 0 ? 4.0 * Hl : 0.0);
                fl += (4.0 + 4.0 * c.ff) * o.H * Hl + 2.0 * 2.5 * n * log2(n) * o.H + 6.0 * (n / 2 + 1) * o.H;
            } else {
                const double in = (double)o.H * o.l, outn = (double)o.Ho * (o.kind == OP_DOWN ? o.l / o.s : o.l * o.s);
                by += 4.0 * (in + outn + (o.skip_buf >= 0 ? outn : 0.0));
                fl += 2.0 * (o.kind == OP_DOWN ? (double)o.H * o.s * o.Ho * (o.l / o.s) : (double)o.H * o.Ho * o.s * o.l);
            }
        }
        by += 4.0 * c.d_model * L + 4.0 * L;
        fl += 2.0 * c.d_model * c.d_model * L + 2.0 * c.d_model * L + 2.0 * c.d_model * L;
    } else {
        const double C = c.res_channels, S = c.skip_channels, N = c.num_res_layers;
        fl = N * (12.0 * C * C * L + 2.0 * C * C * L + 2.0 * C * S * L) + 2.0 * S * S * L + 2.0 * S * L + 2.0 * C * L;
        by = N * 4.0 * (2.0 * C + 2.0 * S) * L + 4.0 * L * (2.0 + C + S);
    }
    *bytes = by;
    *flops = fl;
    return DWB_OK;
}

}  // extern "C"
