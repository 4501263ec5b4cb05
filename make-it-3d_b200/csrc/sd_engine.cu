// sd_engine.cu -- Stable-Diffusion guidance engine for the SDS step (C ABI: include/mi3d.h, section B3).
//
// Replaces what nerf/sd.py:117-174,212-220 runs inside diffusers (UNet2DConditionModel, AutoencoderKL.encode,
// DDIMScheduler.add_noise) with a statically planned list of launches over ONE tensor-core tile kernel
// (tc_gemm.cuh: tcgen05.mma / TMEM / TMA) plus the memory-bound kernels of sd_kernels.cuh:
//   * every Linear, 1x1 conv, attention product and im2col'ed strided conv   -> tc plain GEMM
//   * every 3x3 stride-1 conv (U-Net ResBlocks, VAE encoder, their dgrads)   -> tc implicit GEMM (TMA-shifted taps)
//   * GroupNorm(+SiLU), LayerNorm, softmax, GEGLU (GEMM epilogue), time-embedding, CFG + SDS gradient -> sdk kernels
// The plan (buffers in a caller-provided workspace, TMA descriptors, launch parameters) is built once at
// mi3d_sd_create(); a step is a fixed sequence of launches on the caller's stream -> CUDA-graph capturable.
// Activations: NHWC fp16.  Accumulation / statistics: fp32 (fp64 group sums).  Weights: fp16 copies of the fp32
// diffusers parameters, re-laid-out at load time (conv OIHW -> O,ky,kx,I ; GEGLU rows interleaved ; dgrad copies).
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mi3d.h"
#include "sd_kernels.cuh"
#include "tc_host.cuh"
#include "attn.cuh"

namespace sd {

using Op = std::function<int(cudaStream_t)>;

struct T4 {  // NHWC fp16 activation
    __half* p = nullptr; int n = 0, h = 0, w = 0, c = 0;
    size_t numel() const { return (size_t)n * h * w * c; }
    long long rows() const { return (long long)n * h * w; }
};

enum PKind { PK_F32 = 0, PK_LINEAR, PK_LINEAR_T, PK_LINEAR_BOTH, PK_CONV3, PK_CONV3_BOTH, PK_CONV1, PK_CONV1_BOTH, PK_GEGLU_W, PK_GEGLU_B,
             PK_CONV_F32_OHWI, PK_CONV_SMALL_COUT, PK_CONV_SMALL_COUT_BOTH };

struct Param {
    std::string name; std::vector<int> shape; size_t numel = 0; PKind kind = PK_F32;
    void* dst = nullptr;      // primary storage
    void* dst2 = nullptr;     // secondary (transposed / flipped copy for the input-gradient pass)
    std::string partner;      // GEGLU: weight <-> bias handled together
};

struct Engine {
    mi3d_unet_cfg ucfg; mi3d_vae_cfg vcfg;
    uint8_t* base = nullptr; size_t cap = 0, off = 0; bool dry = true; int err = 0;
    int num_sms = 148;
    std::vector<Param> params; std::map<std::string, int> pindex;
    std::vector<Op> unet_ops, enc_ops, enc_bwd_ops;
    std::vector<Op>* cur = nullptr;
    // GroupNorm statistics of one op list live in one pool, cleared by a single memset at the head of the list
    uint8_t* pool = nullptr; size_t pool_used = 0; static constexpr size_t kPoolBytes = 256 * 1024;
    void begin_list(std::vector<Op>* l) {
        cur = l; pool = (uint8_t*)alloc(kPoolBytes); pool_used = 0;
        uint8_t* pp = pool;
        push([pp](cudaStream_t st) { return (int)cudaMemsetAsync(pp, 0, kPoolBytes, st); });
    }
    double* pool_stats(size_t n_doubles) {
        const size_t b = n_doubles * sizeof(double);
        if (pool_used + b > kPoolBytes) { err = MI3D_ERR_ARG; return (double*)pool; }
        double* p = (double*)(pool + pool_used); pool_used += b; return p;
    }
    // per-ResBlock time-embedding projections of the whole U-Net, evaluated by one grouped launch right after the embedding MLP
    std::vector<sdk::SmallLinearJob> temb_jobs; sdk::SmallLinearJob* temb_jobs_dev = nullptr; int temb_max_n = 0;
    std::map<std::string, std::pair<void*, size_t>> named;   // debug taps: name -> (ptr, bytes)
    // fixed I/O buffers
    T4 unet_in; float* unet_out = nullptr; __half* ctx16 = nullptr; float* temb = nullptr; long long* t_dev = nullptr;
    T4 vae_in; float* vae_moments = nullptr; float* vae_gmoments = nullptr; __half* vae_gin = nullptr;
    T4 dec_in; float* dec_out = nullptr; std::vector<Op> dec_ops;          // VAE decoder (denoise side branch, nerf/sd.py:201-210)
    int ctx_pad = 128;
    float* splitk_ws = nullptr; size_t splitk_elems = 0;      // fp32 scratch shared by all split-K GEMMs (they run one at a time)
    // optional live timing of the tensor-core tile kernel (bench.py roofline leg): CUDA events around every k_tc_gemm launch
    bool profile = false; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events; size_t prof_used = 0;
    struct ProfShape { int M, N, K, bn, splits, conv, batch, epi; }; std::vector<ProfShape> prof_shapes;
    int do_launch(const CUtensorMap& ma, const CUtensorMap& mb, const tc::GemmParams& p, int bn, int batch, cudaStream_t st) {
        if (p.splits > 1) return tc::launch_splitk(ma, mb, p, bn, p.splits, splitk_ws, st);
        return tc::launch(ma, mb, p, bn, batch, st);
    }
    int timed_launch(const CUtensorMap& ma, const CUtensorMap& mb, const tc::GemmParams& p, int bn, int batch, cudaStream_t st) {
        if (!profile) return do_launch(ma, mb, p, bn, batch, st);
        if (prof_used == prof_events.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); prof_events.push_back({a, b}); }
        if (prof_shapes.size() < prof_events.size()) prof_shapes.resize(prof_events.size());
        prof_shapes[prof_used] = {p.M, p.N, p.K, bn, p.splits, p.conv, batch, p.epi_mode};
        auto& ev = prof_events[prof_used++];
        cudaEventRecord(ev.first, st);
        const int r = do_launch(ma, mb, p, bn, batch, st);
        cudaEventRecord(ev.second, st);
        return r;
    }

    void* alloc(size_t bytes) {
        off = (off + 1023) & ~(size_t)1023;
        void* p = dry ? (void*)(uintptr_t)(0x10000 + off) : (void*)(base + off);
        off += bytes;
        if (!dry && off > cap) err = MI3D_ERR_ARG;
        return p;
    }
    __half* alloc16(size_t n) { return (__half*)alloc(n * 2); }
    float* alloc32(size_t n) { return (float*)alloc(n * 4); }
    T4 act(int n, int h, int w, int c) { T4 t; t.n = n; t.h = h; t.w = w; t.c = c; t.p = alloc16(t.numel()); return t; }
    void push(Op op) { if (!dry && cur) cur->push_back(std::move(op)); }
    void tap(const std::string& name, void* p, size_t bytes) { if (!dry) named[name] = {p, bytes}; }

    Param& param(const std::string& name, std::vector<int> shape, PKind kind) {
        auto it = pindex.find(name);
        if (it != pindex.end()) return params[it->second];
        Param P; P.name = name; P.shape = shape; P.kind = kind; P.numel = 1;
        for (int s : shape) P.numel *= (size_t)s;
        switch (kind) {
            case PK_F32: case PK_GEGLU_B: P.dst = alloc(P.numel * 4); break;
            case PK_CONV_F32_OHWI: P.dst = alloc((size_t)std::max(shape[0], shape[1]) * 9 * 8 * 4); break;   // input channels padded to <= 8
            case PK_LINEAR: case PK_CONV3: case PK_CONV1: case PK_GEGLU_W: case PK_LINEAR_T: case PK_CONV_SMALL_COUT: P.dst = alloc(P.numel * 2); break;
            case PK_LINEAR_BOTH: case PK_CONV3_BOTH: case PK_CONV1_BOTH: case PK_CONV_SMALL_COUT_BOTH:
                P.dst = alloc(P.numel * 2); P.dst2 = alloc(P.numel * 2); break;
        }
        pindex[name] = (int)params.size();
        params.push_back(P);
        return params.back();
    }
    const float* pf(const std::string& name, std::vector<int> shape) { return (const float*)param(name, shape, PK_F32).dst; }

    // ------------------------------------------------------------------------------------------------------
    // op builders
    // ------------------------------------------------------------------------------------------------------
    struct Mat { const __half* p; long long ld; long long rows; int K; int z1 = 1; long long sz1 = 0; int z2 = 1; long long sz2 = 0; };
    struct Out { __half* p16 = nullptr; float* p32 = nullptr; long long ldc = 0; int z1 = 1; long long s_lo = 0, s_hi = 0; };
    struct Epi { const float* bias = nullptr; const float* row_bias = nullptr; int rows_per_group = 1; const __half* residual = nullptr;
                 long long ld_res = 0; int mode = tc::EPI_PLAIN; float alpha = 1.f; };

    void gemm(const Mat& a, const Mat& b, int N, const Out& o, const Epi& e, int batch = 1, bool b_batched = false) {
        if (dry) return;
        const long long mt = (a.rows + 127) / 128;
        int bn = tc::pick_block_n(N, mt * batch, num_sms);
        if (bn == 0 || a.K % 64) { err = MI3D_ERR_ARG; return; }
        int splits = 1;
        if (batch == 1 && e.mode == tc::EPI_PLAIN && o.p16 && a.rows % 128 == 0 && (size_t)a.rows * N <= splitk_elems) {
            if (N % 128 == 0 && mt * (N / 128) * 2 <= num_sms) bn = 128;            // few, fat tiles + split-K beats many thin tiles
            splits = tc::plan_splits(mt * (N / bn), a.K / 64, num_sms);
        }
        CUtensorMap ma, mb;
        int r = tc::make_map_matrix(&ma, a.p, a.K, a.rows, a.ld, 128, a.z1, a.sz1, a.z2, a.sz2);
        if (!r) r = tc::make_map_matrix(&mb, b.p, b.K, b.rows, b.ld, bn, b.z1, b.sz1, b.z2, b.sz2);
        if (r) { err = r; return; }
        tc::GemmParams p = {};
        p.M = (int)a.rows; p.N = N; p.K = a.K; p.num_k_blocks = a.K / 64; p.conv = 0;
        p.a_z1 = a.z1; p.b_z1 = b.z1; p.b_batched = b_batched ? 1 : 0;
        p.out = o.p16; p.out_f32 = o.p32; p.ldc = o.ldc; p.out_z1 = o.z1; p.out_s_lo = o.s_lo; p.out_s_hi = o.s_hi;
        p.bias = e.bias; p.row_bias = e.row_bias; p.rows_per_group = e.rows_per_group; p.residual = e.residual; p.ld_res = e.ld_res;
        p.epi_mode = e.mode; p.alpha = e.alpha; p.m_valid = (int)a.rows; p.splits = splits;
        push([this, ma, mb, p, bn, batch](cudaStream_t st) { return timed_launch(ma, mb, p, bn, batch, st); });
    }

    // y[M,N] = x[M,K] W[N,K]^T (+bias) (+residual)
    void linear(const __half* x, long long M, int K, const __half* W, int N, __half* y, const float* bias = nullptr,
                const __half* residual = nullptr, int mode = tc::EPI_PLAIN, long long ldc = -1) {
        Mat a{x, K, M, K}; Mat b{W, K, N, K};
        Out o; o.p16 = y; o.ldc = ldc >= 0 ? ldc : (mode == tc::EPI_GEGLU ? N / 2 : (mode == tc::EPI_TRANSPOSED ? M : N));
        Epi e; e.bias = bias; e.residual = residual; e.ld_res = N; e.mode = mode;
        gemm(a, b, N, o, e);
    }

    // 3x3 stride-1 pad-1 conv via implicit GEMM; w [Cout][3][3][Cin] fp16
    void conv3(const T4& x, const __half* w, int Cout, const T4& y, const float* bias, const float* row_bias = nullptr,
               const __half* residual = nullptr) {
        if (dry) return;
        const long long M = x.rows();
        int bw = x.w < 128 ? x.w : 128, bh = 128 / bw; if (bh > x.h) bh = x.h; const int bn_img = 128 / (bw * bh);
        if (M % 128 || x.c % 64 || bw * bh * bn_img != 128 || x.w % bw || x.h % bh || x.n % bn_img) { err = MI3D_ERR_ARG; return; }
        int bn = tc::pick_block_n(Cout, M / 128, num_sms);
        if (!bn) { err = MI3D_ERR_ARG; return; }
        int splits = 1;
        if ((size_t)M * Cout <= splitk_elems) {
            if (Cout % 128 == 0 && (M / 128) * (Cout / 128) * 2 <= num_sms) bn = 128;
            splits = tc::plan_splits((M / 128) * (Cout / bn), 9 * (x.c / 64), num_sms);
        }
        CUtensorMap ma, mb;
        int r = tc::make_map_nhwc(&ma, x.p, x.c, x.w, x.h, x.n, bw, bh, bn_img);
        if (!r) r = tc::make_map_matrix(&mb, w, 9ull * x.c, Cout, 9ull * x.c, bn);
        if (r) { err = r; return; }
        tc::GemmParams p = {};
        p.M = (int)M; p.N = Cout; p.K = 9 * x.c; p.num_k_blocks = 9 * (x.c / 64); p.conv = 1;
        p.conv_H = x.h; p.conv_W = x.w; p.conv_bw = bw; p.conv_bh = bh; p.cin_blocks = x.c / 64;
        p.a_z1 = 1; p.b_z1 = 1; p.b_batched = 0; p.out = y.p; p.ldc = Cout; p.out_z1 = 1;
        p.bias = bias; p.row_bias = row_bias; p.rows_per_group = x.h * x.w; p.residual = residual; p.ld_res = Cout;
        p.epi_mode = tc::EPI_PLAIN; p.alpha = 1.f; p.m_valid = (int)M; p.splits = splits;
        push([this, ma, mb, p, bn](cudaStream_t st) { return timed_launch(ma, mb, p, bn, 1, st); });
    }

    struct GN { double* stats; const float* gamma; const float* beta; float eps; int silu; T4 x; int G; };
    GN groupnorm(const T4& x, const std::string& pname, float eps, int silu, const T4& y) {
        GN g; g.x = x; g.G = ucfg.groups; g.eps = eps; g.silu = silu;
        g.gamma = pf(pname + ".weight", {x.c}); g.beta = pf(pname + ".bias", {x.c});
        g.stats = pool_stats((size_t)x.n * g.G * 2);
        if (dry) return g;
        const int HW = x.h * x.w, C = x.c, G = g.G, N = x.n;
        const int VPP = C / 8, PL = std::max(1, 256 / VPP), threads = PL * VPP;
        if (G > 64 || threads > 1024 || C % 8) { err = MI3D_ERR_ARG; return g; }
        // enough CTAs to fill the machine twice, at least PL*8 pixels each (8 CTAs per SM measured slower: more partial-sum atomics)
        int chunks = std::max(1, std::min((HW + PL * 8 - 1) / (PL * 8), (2 * num_sms + N - 1) / N));
        const int ppc = (HW + chunks - 1) / chunks; chunks = (HW + ppc - 1) / ppc;
        const __half* xp = x.p; __half* yp = y.p; const GN gg = g;
        push([=](cudaStream_t st) {
            sdk::k_gn_stats<<<dim3(chunks, N), threads, sdk::gn_smem_bytes(PL, C, G), st>>>(xp, HW, C, G, ppc, gg.stats);
            sdk::k_gn_apply<<<dim3(chunks, N), threads, 0, st>>>(xp, yp, gg.stats, gg.gamma, gg.beta, HW, C, G, gg.eps, gg.silu, ppc);
            return (int)cudaGetLastError();
        });
        return g;
    }
    // dx = GN(+SiLU) backward of dy (+ add)
    void groupnorm_bwd(const GN& g, const __half* dy, const __half* add, __half* dx) {
        double* bstats = pool_stats((size_t)g.x.n * g.G * 2);
        if (dry) return;
        const int HW = g.x.h * g.x.w, C = g.x.c, G = g.G, N = g.x.n;
        const int VPP = C / 8, PL = std::max(1, 256 / VPP), threads = PL * VPP;
        int chunks = std::max(1, std::min((HW + PL * 8 - 1) / (PL * 8), (2 * num_sms + N - 1) / N));
        const int ppc = (HW + chunks - 1) / chunks; chunks = (HW + ppc - 1) / ppc;
        const GN gg = g;
        push([=](cudaStream_t st) {
            sdk::k_gn_bwd_stats<<<dim3(chunks, N), threads, sdk::gn_smem_bytes(PL, C, G), st>>>(gg.x.p, dy, gg.stats, gg.gamma, gg.beta, HW, C, G, gg.eps, gg.silu, ppc, bstats);
            sdk::k_gn_bwd_apply<<<dim3(chunks, N), threads, 0, st>>>(gg.x.p, dy, gg.stats, bstats, gg.gamma, gg.beta, HW, C, G, gg.eps, gg.silu, add, dx, ppc);
            return (int)cudaGetLastError();
        });
    }
    void layernorm(const __half* x, long long rows, int C, const std::string& pname, __half* y) {
        const float* ga = pf(pname + ".weight", {C}); const float* be = pf(pname + ".bias", {C});
        push([=](cudaStream_t st) {
            sdk::k_layernorm<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, y, ga, be, (int)rows, C, 1e-5f);
            return (int)cudaGetLastError();
        });
    }
    void softmax(__half* s, size_t rows, int cols, int valid, float scale) {
        push([=](cudaStream_t st) {
            sdk::k_softmax<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(s, rows, cols, valid, scale);
            return (int)cudaGetLastError();
        });
    }
    void transpose(const __half* in, __half* out, int Z, int R, int C) {
        push([=](cudaStream_t st) {
            sdk::k_transpose<<<dim3((C + 31) / 32, (R + 31) / 32, Z), dim3(32, 8), 0, st>>>(in, out, R, C);
            return (int)cudaGetLastError();
        });
    }
    void linear_small(const float* x, const __half* w, const float* bias, float* y, int B, int N, int K, int silu_in, int silu_out) {
        push([=](cudaStream_t st) {
            sdk::k_linear_small<<<(N + 7) / 8, 256, 0, st>>>(x, w, bias, y, B, N, K, silu_in, silu_out);
            return (int)cudaGetLastError();
        });
    }
    static int blocks_for(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 148 * 16); }
    // persistent grids of the small direct convolutions: as many CTAs per SM as their weight staging area allows
    int small_cin_grid(size_t outputs, size_t smem) const { const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1))); return (int)std::min<size_t>((outputs / 8 + 255) / 256, (size_t)num_sms * per_sm); }
    int small_cout_grid(size_t pixels, int cin) const { const int ppw = cin <= 128 ? 2 : 1; return (int)std::min<size_t>((pixels / ppw + 7) / 8 + 1, (size_t)num_sms * 8); }

    // ------------------------------------------------------------------------------------------------------
    // U-Net pieces (diffusers naming)
    // ------------------------------------------------------------------------------------------------------
    T4 resnet(const std::string& pre, const T4& x, int cout, bool has_temb, float eps, std::vector<GN>* tape = nullptr, std::vector<T4>* acts = nullptr) {
        const int cin = x.c;
        T4 a1 = act(x.n, x.h, x.w, cin);
        GN g1 = groupnorm(x, pre + ".norm1", eps, 1, a1);
        const float* rb = nullptr;
        if (has_temb) {
            const int tdim = ucfg.block_out[0] * 4;
            const __half* wt = (const __half*)param(pre + ".time_emb_proj.weight", {cout, tdim}, PK_LINEAR).dst;
            const float* bt = pf(pre + ".time_emb_proj.bias", {cout});
            float* tp = alloc32((size_t)x.n * cout);
            temb_jobs.push_back({wt, bt, tp, cout});          // evaluated by the grouped launch at the head of the U-Net
            temb_max_n = std::max(temb_max_n, cout);
            rb = tp;
        }
        const bool both = tape != nullptr;
        Param& w1 = param(pre + ".conv1.weight", {cout, cin, 3, 3}, both ? PK_CONV3_BOTH : PK_CONV3);
        T4 h1 = act(x.n, x.h, x.w, cout);
        conv3(a1, (const __half*)w1.dst, cout, h1, pf(pre + ".conv1.bias", {cout}), rb);
        T4 a2 = act(x.n, x.h, x.w, cout);
        GN g2 = groupnorm(h1, pre + ".norm2", eps, 1, a2);
        const __half* sc = x.p;
        if (cin != cout) {
            Param& ws = param(pre + ".conv_shortcut.weight", {cout, cin, 1, 1}, both ? PK_CONV1_BOTH : PK_CONV1);
            T4 s = act(x.n, x.h, x.w, cout);
            linear(x.p, x.rows(), cin, (const __half*)ws.dst, cout, s.p, pf(pre + ".conv_shortcut.bias", {cout}));
            sc = s.p;
        }
        Param& w2 = param(pre + ".conv2.weight", {cout, cout, 3, 3}, both ? PK_CONV3_BOTH : PK_CONV3);
        T4 out = act(x.n, x.h, x.w, cout);
        conv3(a2, (const __half*)w2.dst, cout, out, pf(pre + ".conv2.bias", {cout}), nullptr, sc);
        if (tape) { tape->push_back(g1); tape->push_back(g2); }
        if (acts) { acts->push_back(a1); acts->push_back(a2); }
        return out;
    }

    // multi-head attention core on token matrices (attn.cuh).  q: [B*T, ldq] (head h at column h*64), k / v likewise
    // [B*Tk, ld]; only the first Tk_valid keys of each batch take part.  Writes o [B*T, C].
    void attention_core(const __half* q, long long ldq, const __half* k, long long ldk, const __half* v, long long ldv, int B, int T, int Tk,
                        int Tk_valid, int heads, int d, __half* o, int C) {
        if (d != 64) { err = MI3D_ERR_ARG; return; }
        if (dry) return;
        attn::Maps m;
        if (int r = attn::make_maps(&m, q, ldq, k, ldk, v, ldv, B, T, Tk, heads)) { err = r; return; }
        push([=](cudaStream_t st) {
            if (!profile) return attn::launch(m, o, C, B, T, Tk_valid, heads, st);
            // profiling: same event bracket as the tile kernel; recorded with conv = 2 (M = queries, N = keys, K = head_dim, batch = B * heads)
            if (prof_used == prof_events.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); prof_events.push_back({a, b}); }
            if (prof_shapes.size() < prof_events.size()) prof_shapes.resize(prof_events.size());
            prof_shapes[prof_used] = {T, Tk_valid, 64, 0, 1, 2, B * heads, 0};
            auto& ev = prof_events[prof_used++];
            cudaEventRecord(ev.first, st);
            const int r = attn::launch(m, o, C, B, T, Tk_valid, heads, st);
            cudaEventRecord(ev.second, st);
            return r;
        });
    }

    T4 transformer(const std::string& pre, const T4& x, int heads) {
        const int C = x.c, B = x.n, T = x.h * x.w; const long long M = x.rows();
        const int d = C / heads, cross = ucfg.cross_dim;
        T4 n = act(x.n, x.h, x.w, C);
        groupnorm(x, pre + ".norm", 1e-6f, 0, n);
        __half* h = alloc16(M * C);
        linear(n.p, M, C, (const __half*)param(pre + ".proj_in.weight", {C, C}, PK_LINEAR).dst, C, h, pf(pre + ".proj_in.bias", {C}));
        const std::string tb = pre + ".transformer_blocks.0";
        // ---- self attention ----
        __half* ln = alloc16(M * C);
        layernorm(h, M, C, tb + ".norm1", ln);
        // to_q / to_k / to_v weights are allocated back to back ([3C][C]): one GEMM produces q | k | v side by side
        const __half* wq = (const __half*)param(tb + ".attn1.to_q.weight", {C, C}, PK_LINEAR).dst;
        const __half* wk = (const __half*)param(tb + ".attn1.to_k.weight", {C, C}, PK_LINEAR).dst;
        const __half* wv = (const __half*)param(tb + ".attn1.to_v.weight", {C, C}, PK_LINEAR).dst;
        __half* ao = alloc16(M * C);
        if (wk == wq + (size_t)C * C && wv == wk + (size_t)C * C) {
            __half* qkv = alloc16(M * 3 * C);
            linear(ln, M, C, wq, 3 * C, qkv);
            attention_core(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, B, T, T, T, heads, d, ao, C);
        } else {
            __half* qb = alloc16(M * C); __half* kb = alloc16(M * C); __half* vb = alloc16(M * C);
            linear(ln, M, C, wq, C, qb); linear(ln, M, C, wk, C, kb); linear(ln, M, C, wv, C, vb);
            attention_core(qb, C, kb, C, vb, C, B, T, T, T, heads, d, ao, C);
        }
        __half* h2 = alloc16(M * C);
        linear(ao, M, C, (const __half*)param(tb + ".attn1.to_out.0.weight", {C, C}, PK_LINEAR).dst, C, h2, pf(tb + ".attn1.to_out.0.bias", {C}), h);
        // ---- cross attention (keys/values from the text context, padded to ctx_pad rows per batch) ----
        __half* ln2 = alloc16(M * C);
        layernorm(h2, M, C, tb + ".norm2", ln2);
        const long long Mc = (long long)B * ctx_pad;
        __half* q2 = alloc16(M * C);
        linear(ln2, M, C, (const __half*)param(tb + ".attn2.to_q.weight", {C, C}, PK_LINEAR).dst, C, q2);
        const __half* wk2 = (const __half*)param(tb + ".attn2.to_k.weight", {C, cross}, PK_LINEAR).dst;
        const __half* wv2 = (const __half*)param(tb + ".attn2.to_v.weight", {C, cross}, PK_LINEAR).dst;
        __half* ao2 = alloc16(M * C);
        if (wv2 == wk2 + (size_t)C * cross) {
            __half* kv = alloc16(Mc * 2 * C);
            linear(ctx16, Mc, cross, wk2, 2 * C, kv);
            attention_core(q2, C, kv, 2 * C, kv + C, 2 * C, B, T, ctx_pad, ucfg.ctx_len, heads, d, ao2, C);
        } else {
            __half* k2 = alloc16(Mc * C); __half* v2 = alloc16(Mc * C);
            linear(ctx16, Mc, cross, wk2, C, k2); linear(ctx16, Mc, cross, wv2, C, v2);
            attention_core(q2, C, k2, C, v2, C, B, T, ctx_pad, ucfg.ctx_len, heads, d, ao2, C);
        }
        __half* h3 = alloc16(M * C);
        linear(ao2, M, C, (const __half*)param(tb + ".attn2.to_out.0.weight", {C, C}, PK_LINEAR).dst, C, h3, pf(tb + ".attn2.to_out.0.bias", {C}), h2);
        // ---- GEGLU feed-forward ----
        __half* ln3 = alloc16(M * C);
        layernorm(h3, M, C, tb + ".norm3", ln3);
        Param& wg = param(tb + ".ff.net.0.proj.weight", {8 * C, C}, PK_GEGLU_W);
        Param& bg = param(tb + ".ff.net.0.proj.bias", {8 * C}, PK_GEGLU_B);
        wg.partner = bg.name; bg.partner = wg.name;
        __half* f1 = alloc16(M * 4 * C);
        linear(ln3, M, C, (const __half*)wg.dst, 8 * C, f1, (const float*)bg.dst, nullptr, tc::EPI_GEGLU);
        __half* h4 = alloc16(M * C);
        linear(f1, M, 4 * C, (const __half*)param(tb + ".ff.net.2.weight", {C, 4 * C}, PK_LINEAR).dst, C, h4, pf(tb + ".ff.net.2.bias", {C}), h3);
        T4 out = act(x.n, x.h, x.w, C);
        linear(h4, M, C, (const __half*)param(pre + ".proj_out.weight", {C, C}, PK_LINEAR).dst, C, out.p, pf(pre + ".proj_out.bias", {C}), x.p);
        return out;
    }

    T4 downsample(const std::string& pre, const T4& x, int pad_lo, bool both, __half** col_out = nullptr) {
        const int Ho = x.h / 2, Wo = x.w / 2, C = x.c;
        const long long Mo = (long long)x.n * Ho * Wo;
        __half* col = alloc16((size_t)Mo * 9 * C);
        const T4 xx = x;
        push([=](cudaStream_t st) {
            sdk::k_im2col_s2<<<blocks_for((size_t)Mo * 9 * C / 8), 256, 0, st>>>(xx.p, col, xx.n, xx.h, xx.w, C, Ho, Wo, pad_lo);
            return (int)cudaGetLastError();
        });
        Param& w = param(pre + ".conv.weight", {C, C, 3, 3}, both ? PK_CONV3_BOTH : PK_CONV3);
        T4 y = act(x.n, Ho, Wo, C);
        linear(col, Mo, 9 * C, (const __half*)w.dst, C, y.p, pf(pre + ".conv.bias", {C}));
        if (col_out) *col_out = col;
        return y;
    }

    T4 concat(const T4& a, const T4& b) {
        T4 o = act(a.n, a.h, a.w, a.c + b.c);
        const T4 aa = a, bb = b;
        push([=](cudaStream_t st) {
            sdk::k_concat<<<blocks_for(o.numel() / 8), 256, 0, st>>>(aa.p, bb.p, o.p, (size_t)aa.rows(), aa.c, bb.c);
            return (int)cudaGetLastError();
        });
        return o;
    }

    void build_unet() {
        begin_list(&unet_ops);
        const mi3d_unet_cfg& c = ucfg;
        const int B = c.batch, HW = c.latent_hw, nlev = c.n_levels, L = c.layers_per_block;
        const int c0 = c.block_out[0], tdim = c0 * 4;
        unet_in = act(B, HW, HW, c.in_ch);
        t_dev = (long long*)alloc(sizeof(long long));
        ctx16 = alloc16((size_t)B * ctx_pad * c.cross_dim);
        // time embedding
        float* tproj = alloc32((size_t)B * c0);
        float* t1 = alloc32((size_t)B * tdim);
        temb = alloc32((size_t)B * tdim);
        {
            long long* tptr = t_dev;
            push([=](cudaStream_t st) {
                sdk::k_time_proj<<<(B * c0 / 2 + 127) / 128, 128, 0, st>>>(tptr, tproj, B, c0);
                return (int)cudaGetLastError();
            });
            linear_small(tproj, (const __half*)param("time_embedding.linear_1.weight", {tdim, c0}, PK_LINEAR).dst, pf("time_embedding.linear_1.bias", {tdim}), t1, B, tdim, c0, 0, 1);
            linear_small(t1, (const __half*)param("time_embedding.linear_2.weight", {tdim, tdim}, PK_LINEAR).dst, pf("time_embedding.linear_2.bias", {tdim}), temb, B, tdim, tdim, 0, 0);
            temb_jobs.clear(); temb_max_n = 0;
            temb_jobs_dev = (sdk::SmallLinearJob*)alloc(64 * sizeof(sdk::SmallLinearJob));
            const float* te = temb;
            push([this, te, B, tdim](cudaStream_t st) {
                if (temb_jobs.empty()) return 0;
                sdk::k_linear_small_grouped<<<dim3((temb_max_n + 7) / 8, (unsigned)temb_jobs.size()), 256, 0, st>>>(te, temb_jobs_dev, B, tdim);
                return (int)cudaGetLastError();
            });
        }
        // conv_in (Cin = 4): direct
        T4 h = act(B, HW, HW, c0);
        {
            const float* w = (const float*)param("conv_in.weight", {c0, c.in_ch, 3, 3}, PK_CONV_F32_OHWI).dst;
            const float* b = pf("conv_in.bias", {c0});
            const T4 xin = unet_in, ho = h; const int cin = c.in_ch;
            push([=](cudaStream_t st) {
                if (cin != 4 || c0 % 8) return (int)MI3D_ERR_ARG;
                const size_t sm = (size_t)c0 * 9 * 4 * sizeof(float);
                sdk::k_conv_small_cin<4><<<small_cin_grid(ho.numel(), sm), 256, sm, st>>>(xin.p, w, b, ho.p, xin.n, xin.h, xin.w, c0);
                return (int)cudaGetLastError();
            });
        }
        tap("unet.conv_in", h.p, h.numel() * 2);
        std::vector<T4> skips; skips.push_back(h);
        for (int i = 0; i < nlev; i++) {
            const bool has_attn = i < nlev - 1;
            for (int j = 0; j < L; j++) {
                const std::string rp = "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
                h = resnet(rp, h, c.block_out[i], true, 1e-5f);
                if (has_attn) h = transformer("down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), h, c.heads[i]);
                skips.push_back(h);
            }
            if (i < nlev - 1) {
                h = downsample("down_blocks." + std::to_string(i) + ".downsamplers.0", h, 1, false);
                skips.push_back(h);
            }
            tap("unet.down" + std::to_string(i), h.p, h.numel() * 2);
        }
        h = resnet("mid_block.resnets.0", h, h.c, true, 1e-5f);
        h = transformer("mid_block.attentions.0", h, c.heads[nlev - 1]);
        h = resnet("mid_block.resnets.1", h, h.c, true, 1e-5f);
        tap("unet.mid", h.p, h.numel() * 2);
        for (int i = 0; i < nlev; i++) {
            const int lev = nlev - 1 - i;
            const bool has_attn = i > 0;
            for (int j = 0; j < L + 1; j++) {
                T4 s = skips.back(); skips.pop_back();
                T4 cat = concat(h, s);
                h = resnet("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), cat, c.block_out[lev], true, 1e-5f);
                if (has_attn) h = transformer("up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), h, c.heads[lev]);
            }
            if (i < nlev - 1) {
                T4 up = act(h.n, h.h * 2, h.w * 2, h.c);
                const T4 hh = h;
                push([=](cudaStream_t st) {
                    sdk::k_upsample2x<<<blocks_for(up.numel() / 8), 256, 0, st>>>(hh.p, up.p, hh.n, hh.h, hh.w, hh.c);
                    return (int)cudaGetLastError();
                });
                T4 y = act(up.n, up.h, up.w, up.c);
                const std::string up_pre = "up_blocks." + std::to_string(i) + ".upsamplers.0.conv";
                conv3(up, (const __half*)param(up_pre + ".weight", {h.c, h.c, 3, 3}, PK_CONV3).dst, h.c, y, pf(up_pre + ".bias", {h.c}));
                h = y;
            }
            tap("unet.up" + std::to_string(i), h.p, h.numel() * 2);
        }
        T4 a = act(h.n, h.h, h.w, h.c);
        groupnorm(h, "conv_norm_out", 1e-5f, 1, a);
        unet_out = alloc32((size_t)B * HW * HW * c.out_ch);
        {
            const __half* w = (const __half*)param("conv_out.weight", {c.out_ch, c0, 3, 3}, PK_CONV_SMALL_COUT).dst;
            const float* b = pf("conv_out.bias", {c.out_ch});
            float* o = unet_out; const int oc = c.out_ch;
            push([=](cudaStream_t st) {
                if (oc != 4) return (int)MI3D_ERR_ARG;
                const size_t pix = (size_t)a.rows();
                sdk::k_conv_small_cout<4><<<small_cout_grid(pix, a.c), 256, (size_t)4 * 9 * a.c * 2, st>>>(a.p, w, b, o, a.n, a.h, a.w, a.c);
                return (int)cudaGetLastError();
            });
        }
        if (temb_jobs.size() > 64) err = MI3D_ERR_ARG;
        if (!dry && !err && !temb_jobs.empty())
            if (cudaMemcpy(temb_jobs_dev, temb_jobs.data(), temb_jobs.size() * sizeof(sdk::SmallLinearJob), cudaMemcpyHostToDevice) != cudaSuccess) err = MI3D_ERR_ARG;
        cur = nullptr;
    }

    // ------------------------------------------------------------------------------------------------------
    // VAE encoder forward + input-gradient backward (AutoencoderKL.encode, nerf/sd.py:212-220)
    // ------------------------------------------------------------------------------------------------------
    struct ResTape { T4 x, a1, h1, a2, out; GN g1, g2; bool shortcut; std::string pre; int cin, cout; };
    struct DownTape { T4 x, y; std::string pre; };
    struct AttnTape { T4 x, n; GN g; __half *q, *k, *v, *P, *ao; int C, T; std::string pre; };

    std::vector<ResTape> vres; std::vector<DownTape> vdown; AttnTape vattn; GN vgn_out; T4 v_a_out, v_h_last, v_conv_in_out;
    std::vector<int> vorder;   // 0 = res, 1 = down, 2 = attn ; forward order for the reverse walk

    T4 vae_resnet(const std::string& pre, const T4& x, int cout) {
        ResTape t; t.pre = pre; t.x = x; t.cin = x.c; t.cout = cout; t.shortcut = x.c != cout;
        std::vector<GN> gns; std::vector<T4> acts;
        // re-implemented inline (needs h1 for the tape)
        const int cin = x.c;
        t.a1 = act(x.n, x.h, x.w, cin);
        t.g1 = groupnorm(x, pre + ".norm1", 1e-6f, 1, t.a1);
        Param& w1 = param(pre + ".conv1.weight", {cout, cin, 3, 3}, PK_CONV3_BOTH);
        t.h1 = act(x.n, x.h, x.w, cout);
        conv3(t.a1, (const __half*)w1.dst, cout, t.h1, pf(pre + ".conv1.bias", {cout}));
        t.a2 = act(x.n, x.h, x.w, cout);
        t.g2 = groupnorm(t.h1, pre + ".norm2", 1e-6f, 1, t.a2);
        const __half* sc = x.p;
        if (t.shortcut) {
            Param& ws = param(pre + ".conv_shortcut.weight", {cout, cin, 1, 1}, PK_CONV1_BOTH);
            T4 s = act(x.n, x.h, x.w, cout);
            linear(x.p, x.rows(), cin, (const __half*)ws.dst, cout, s.p, pf(pre + ".conv_shortcut.bias", {cout}));
            sc = s.p;
        }
        Param& w2 = param(pre + ".conv2.weight", {cout, cout, 3, 3}, PK_CONV3_BOTH);
        t.out = act(x.n, x.h, x.w, cout);
        conv3(t.a2, (const __half*)w2.dst, cout, t.out, pf(pre + ".conv2.bias", {cout}), nullptr, sc);
        vres.push_back(t); vorder.push_back(0);
        return t.out;
    }

    void build_vae() {
        begin_list(&enc_ops);
        const mi3d_vae_cfg& c = vcfg;
        const int S = c.image_hw, nlev = c.n_levels, L = c.layers_per_block;
        vae_in = act(1, S, S, 4);          // 3 channels padded to 4 (the 4th is zero)
        T4 h = act(1, S, S, c.block_out[0]);
        {
            const float* w = (const float*)param("encoder.conv_in.weight", {c.block_out[0], c.in_ch, 3, 3}, PK_CONV_F32_OHWI).dst;
            const float* b = pf("encoder.conv_in.bias", {c.block_out[0]});
            const T4 xin = vae_in, ho = h; const int co = c.block_out[0];
            push([=](cudaStream_t st) {
                if (co % 8) return (int)MI3D_ERR_ARG;
                const size_t sm = (size_t)co * 9 * 4 * sizeof(float);
                sdk::k_conv_small_cin<4><<<small_cin_grid(ho.numel(), sm), 256, sm, st>>>(xin.p, w, b, ho.p, 1, xin.h, xin.w, co);
                return (int)cudaGetLastError();
            });
        }
        v_conv_in_out = h;
        int ch = c.block_out[0];
        for (int i = 0; i < nlev; i++) {
            for (int j = 0; j < L; j++) {
                h = vae_resnet("encoder.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), h, c.block_out[i]);
                ch = c.block_out[i];
            }
            if (i < nlev - 1) {
                DownTape d; d.x = h; d.pre = "encoder.down_blocks." + std::to_string(i) + ".downsamplers.0";
                h = downsample(d.pre, h, 0, true);
                d.y = h; vdown.push_back(d); vorder.push_back(1);
            }
        }
        h = vae_resnet("encoder.mid_block.resnets.0", h, ch);
        {   // single-head attention over all pixels
            AttnTape& a = vattn; a.pre = "encoder.mid_block.attentions.0"; a.x = h; a.C = ch; a.T = h.h * h.w;
            const long long M = h.rows(); const int C = ch;
            a.n = act(h.n, h.h, h.w, C);
            a.g = groupnorm(h, a.pre + ".group_norm", 1e-6f, 0, a.n);
            a.q = alloc16(M * C); a.k = alloc16(M * C); a.v = alloc16(M * C);
            __half* vT = alloc16(M * C);
            linear(a.n.p, M, C, (const __half*)param(a.pre + ".to_q.weight", {C, C}, PK_LINEAR_BOTH).dst, C, a.q, pf(a.pre + ".to_q.bias", {C}));
            linear(a.n.p, M, C, (const __half*)param(a.pre + ".to_k.weight", {C, C}, PK_LINEAR_BOTH).dst, C, a.k, pf(a.pre + ".to_k.bias", {C}));
            linear(a.n.p, M, C, (const __half*)param(a.pre + ".to_v.weight", {C, C}, PK_LINEAR_BOTH).dst, C, a.v, pf(a.pre + ".to_v.bias", {C}));
            transpose(a.v, vT, 1, (int)M, C);
            a.P = alloc16((size_t)M * M);
            {
                Mat qa{a.q, C, M, C}; Mat kb{a.k, C, M, C};
                Out oo; oo.p16 = a.P; oo.ldc = M; Epi e;
                gemm(qa, kb, (int)M, oo, e);
            }
            softmax(a.P, (size_t)M, (int)M, (int)M, 1.0f / sqrtf((float)C));
            a.ao = alloc16(M * C);
            {
                Mat pa{a.P, M, M, (int)M}; Mat vb{vT, M, C, (int)M};
                Out oo; oo.p16 = a.ao; oo.ldc = C; Epi e;
                gemm(pa, vb, C, oo, e);
            }
            T4 out = act(h.n, h.h, h.w, C);
            linear(a.ao, M, C, (const __half*)param(a.pre + ".to_out.0.weight", {C, C}, PK_LINEAR_BOTH).dst, C, out.p, pf(a.pre + ".to_out.0.bias", {C}), h.p);
            h = out; vorder.push_back(2);
        }
        h = vae_resnet("encoder.mid_block.resnets.1", h, ch);
        v_h_last = h;
        v_a_out = act(h.n, h.h, h.w, ch);
        vgn_out = groupnorm(h, "encoder.conv_norm_out", 1e-6f, 1, v_a_out);
        // conv_out (ch -> 2*latent) direct + quant_conv (1x1, 2l -> 2l) folded in a tiny kernel at the API level
        const int l2 = 2 * c.latent_ch;
        float* co = alloc32((size_t)h.rows() * l2);
        {
            const __half* w = (const __half*)param("encoder.conv_out.weight", {l2, ch, 3, 3}, PK_CONV_SMALL_COUT_BOTH).dst;
            const float* b = pf("encoder.conv_out.bias", {l2});
            const T4 a = v_a_out;
            push([=](cudaStream_t st) {
                if (l2 != 8) return (int)MI3D_ERR_ARG;
                sdk::k_conv_small_cout<8><<<small_cout_grid((size_t)a.rows(), a.c), 256, (size_t)8 * 9 * a.c * 2, st>>>(a.p, w, b, co, a.n, a.h, a.w, a.c);
                return (int)cudaGetLastError();
            });
        }
        vae_moments = co;
        pf("quant_conv.weight", {l2, l2, 1, 1}); pf("quant_conv.bias", {l2});
        cur = nullptr;
    }

    // ------------------------------------------------------------------------------------------------------
    // VAE decoder (diffusers AutoencoderKL.decode: post_quant_conv folded into the API kernel, then Decoder): forward only.
    // ------------------------------------------------------------------------------------------------------
    T4 dec_resnet(const std::string& pre, const T4& x, int cout) {
        const int cin = x.c;
        T4 a1 = act(x.n, x.h, x.w, cin);
        groupnorm(x, pre + ".norm1", 1e-6f, 1, a1);
        T4 h1 = act(x.n, x.h, x.w, cout);
        conv3(a1, (const __half*)param(pre + ".conv1.weight", {cout, cin, 3, 3}, PK_CONV3).dst, cout, h1, pf(pre + ".conv1.bias", {cout}));
        T4 a2 = act(x.n, x.h, x.w, cout);
        groupnorm(h1, pre + ".norm2", 1e-6f, 1, a2);
        const __half* sc = x.p;
        if (cin != cout) {
            T4 sh = act(x.n, x.h, x.w, cout);
            linear(x.p, x.rows(), cin, (const __half*)param(pre + ".conv_shortcut.weight", {cout, cin, 1, 1}, PK_CONV1).dst, cout, sh.p,
                   pf(pre + ".conv_shortcut.bias", {cout}));
            sc = sh.p;
        }
        T4 out = act(x.n, x.h, x.w, cout);
        conv3(a2, (const __half*)param(pre + ".conv2.weight", {cout, cout, 3, 3}, PK_CONV3).dst, cout, out, pf(pre + ".conv2.bias", {cout}), nullptr, sc);
        return out;
    }

    void build_vae_dec() {
        begin_list(&dec_ops);
        const mi3d_vae_cfg& c = vcfg;
        const int nlev = c.n_levels, L = c.layers_per_block, hw = c.image_hw >> (nlev - 1);
        int ch = c.block_out[nlev - 1];
        dec_in = act(1, hw, hw, 4);                         // post_quant_conv(latents / 0.18215), NHWC fp16, written by the API kernel
        pf("post_quant_conv.weight", {c.latent_ch, c.latent_ch, 1, 1}); pf("post_quant_conv.bias", {c.latent_ch});
        T4 h = act(1, hw, hw, ch);
        {
            const float* w = (const float*)param("decoder.conv_in.weight", {ch, c.latent_ch, 3, 3}, PK_CONV_F32_OHWI).dst;
            const float* b = pf("decoder.conv_in.bias", {ch});
            const T4 xin = dec_in, ho = h; const int co = ch, lc = c.latent_ch;
            push([=](cudaStream_t st) {
                if (lc != 4 || co % 8) return (int)MI3D_ERR_ARG;
                const size_t sm = (size_t)co * 9 * 4 * sizeof(float);
                if (sm > 100 * 1024) return (int)MI3D_ERR_ARG;
                sdk::k_conv_small_cin<4><<<small_cin_grid(ho.numel(), sm), 256, sm, st>>>(xin.p, w, b, ho.p, 1, xin.h, xin.w, co);
                return (int)cudaGetLastError();
            });
        }
        h = dec_resnet("decoder.mid_block.resnets.0", h, ch);
        {   // single-head attention over all pixels (same construction as the encoder's, no tape)
            const std::string pre = "decoder.mid_block.attentions.0";
            const long long M = h.rows(); const int C = ch;
            T4 n = act(h.n, h.h, h.w, C);
            groupnorm(h, pre + ".group_norm", 1e-6f, 0, n);
            __half *q = alloc16(M * C), *k = alloc16(M * C), *v = alloc16(M * C), *vT = alloc16(M * C);
            linear(n.p, M, C, (const __half*)param(pre + ".to_q.weight", {C, C}, PK_LINEAR).dst, C, q, pf(pre + ".to_q.bias", {C}));
            linear(n.p, M, C, (const __half*)param(pre + ".to_k.weight", {C, C}, PK_LINEAR).dst, C, k, pf(pre + ".to_k.bias", {C}));
            linear(n.p, M, C, (const __half*)param(pre + ".to_v.weight", {C, C}, PK_LINEAR).dst, C, v, pf(pre + ".to_v.bias", {C}));
            transpose(v, vT, 1, (int)M, C);
            __half* P = alloc16((size_t)M * M);
            { Mat qa{q, C, M, C}; Mat kb{k, C, M, C}; Out oo; oo.p16 = P; oo.ldc = M; Epi e; gemm(qa, kb, (int)M, oo, e); }
            softmax(P, (size_t)M, (int)M, (int)M, 1.0f / sqrtf((float)C));
            __half* ao = alloc16(M * C);
            { Mat pa{P, M, M, (int)M}; Mat vb{vT, M, C, (int)M}; Out oo; oo.p16 = ao; oo.ldc = C; Epi e; gemm(pa, vb, C, oo, e); }
            T4 out = act(h.n, h.h, h.w, C);
            linear(ao, M, C, (const __half*)param(pre + ".to_out.0.weight", {C, C}, PK_LINEAR).dst, C, out.p, pf(pre + ".to_out.0.bias", {C}), h.p);
            h = out;
        }
        h = dec_resnet("decoder.mid_block.resnets.1", h, ch);
        for (int i = 0; i < nlev; i++) {
            const int co = c.block_out[nlev - 1 - i];
            for (int j = 0; j < L + 1; j++) h = dec_resnet("decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), h, co);
            ch = co;
            if (i < nlev - 1) {
                T4 up = act(h.n, h.h * 2, h.w * 2, h.c);
                const T4 hh = h;
                push([=](cudaStream_t st) {
                    sdk::k_upsample2x<<<blocks_for(up.numel() / 8), 256, 0, st>>>(hh.p, up.p, hh.n, hh.h, hh.w, hh.c);
                    return (int)cudaGetLastError();
                });
                T4 y = act(up.n, up.h, up.w, up.c);
                const std::string up_pre = "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv";
                conv3(up, (const __half*)param(up_pre + ".weight", {h.c, h.c, 3, 3}, PK_CONV3).dst, h.c, y, pf(up_pre + ".bias", {h.c}));
                h = y;
            }
        }
        T4 a = act(h.n, h.h, h.w, h.c);
        groupnorm(h, "decoder.conv_norm_out", 1e-6f, 1, a);
        dec_out = alloc32((size_t)a.rows() * 4);
        {
            // conv_out ch -> 3: the 4-output direct kernel; its 4th weight row / bias entry is the (zeroed) pad allocated right behind
            const __half* w = (const __half*)param("decoder.conv_out.weight", {c.in_ch, ch, 3, 3}, PK_CONV_SMALL_COUT).dst;
            alloc16((size_t)9 * ch);
            const float* b = pf("decoder.conv_out.bias", {c.in_ch});
            alloc32(4);
            float* o = dec_out; const int ic = c.in_ch;
            push([=](cudaStream_t st) {
                if (ic != 3) return (int)MI3D_ERR_ARG;
                sdk::k_conv_small_cout<4><<<small_cout_grid((size_t)a.rows(), a.c), 256, (size_t)4 * 9 * a.c * 2, st>>>(a.p, w, b, o, a.n, a.h, a.w, a.c);
                return (int)cudaGetLastError();
            });
        }
        cur = nullptr;
    }

    // dgrad of a stride-1 3x3 conv: dx = conv3(dy, W_flipped^T)  (+ residual accumulation)
    void conv3_dgrad(const T4& dy, const std::string& wname, int cin, const T4& dx, const __half* add = nullptr) {
        Param& w = params[pindex[wname]];
        conv3(dy, (const __half*)w.dst2, cin, dx, nullptr, nullptr, add);
    }

    void build_vae_bwd() {
        begin_list(&enc_bwd_ops);
        const mi3d_vae_cfg& c = vcfg;
        const int l2 = 2 * c.latent_ch;
        const T4 hl = v_h_last; const int ch = hl.c;
        vae_gmoments = alloc32((size_t)hl.rows() * l2);        // d/d(conv_out output) fp32 [pix][8], filled by the API kernel
        // conv_out^T : small-Cout transposed = "small Cin" conv with Cin = 8 -> output ch channels (fp16 in)
        T4 g8 = act(hl.n, hl.h, hl.w, 8);
        T4 da = act(hl.n, hl.h, hl.w, ch);
        {
            float* gm = vae_gmoments; const T4 g8c = g8, dac = da;
            const float* wf = (const float*)param("encoder.conv_out.weight#dgrad_f32", {ch, 8, 3, 3}, PK_CONV_F32_OHWI).dst;
            push([=](cudaStream_t st) {
                // fp32 -> fp16 staging of the 8-channel gradient
                sdk::k_f32_to_f16<<<blocks_for(g8c.numel()), 256, 0, st>>>(gm, g8c.p, g8c.numel());
                if (ch % 8) return (int)MI3D_ERR_ARG;
                const size_t sm = (size_t)ch * 9 * 8 * sizeof(float);
                sdk::k_conv_small_cin<8><<<small_cin_grid(dac.numel(), sm), 256, sm, st>>>(g8c.p, wf, nullptr, dac.p, g8c.n, g8c.h, g8c.w, ch);
                return (int)cudaGetLastError();
            });
        }
        T4 dh = act(hl.n, hl.h, hl.w, ch);
        groupnorm_bwd(vgn_out, da.p, nullptr, dh.p);
        // reverse walk
        int ri = (int)vres.size() - 1, di = (int)vdown.size() - 1;
        for (int oi = (int)vorder.size() - 1; oi >= 0; oi--) {
            const int kind = vorder[oi];
            if (kind == 0) {
                const ResTape& t = vres[ri--];
                // out = sc(x) + conv2(a2) ; a2 = silu(gn2(h1)) ; h1 = conv1(a1) ; a1 = silu(gn1(x))
                T4 da2 = act(t.x.n, t.x.h, t.x.w, t.cout);
                conv3_dgrad(dh, t.pre + ".conv2.weight", t.cout, da2);
                T4 dh1 = act(t.x.n, t.x.h, t.x.w, t.cout);
                groupnorm_bwd(t.g2, da2.p, nullptr, dh1.p);
                T4 da1 = act(t.x.n, t.x.h, t.x.w, t.cin);
                conv3_dgrad(dh1, t.pre + ".conv1.weight", t.cin, da1);
                // shortcut branch gradient
                const __half* dsc = dh.p;
                if (t.shortcut) {
                    Param& ws = params[pindex[t.pre + ".conv_shortcut.weight"]];
                    T4 d = act(t.x.n, t.x.h, t.x.w, t.cin);
                    linear(dh.p, t.x.rows(), t.cout, (const __half*)ws.dst2, t.cin, d.p);     // dst2 = W^T [cin][cout]
                    dsc = d.p;
                }
                T4 dx = act(t.x.n, t.x.h, t.x.w, t.cin);
                groupnorm_bwd(t.g1, da1.p, dsc, dx.p);
                dh = dx;
            } else if (kind == 1) {
                const DownTape& d = vdown[di--];
                const int C = d.x.c; const long long Mo = d.y.rows();
                Param& w = params[pindex[d.pre + ".conv.weight"]];
                // dcol[Mo, 9C] = dy[Mo, C] . Wt   with Wt = [9C][C]  (dst2 of a CONV3_BOTH strided weight = plain transpose, see load)
                __half* dcol = alloc16((size_t)Mo * 9 * C);
                linear(dh.p, Mo, C, (const __half*)w.dst2, 9 * C, dcol);
                T4 dx = act(d.x.n, d.x.h, d.x.w, C);
                const T4 xx = d.x, yy = d.y;
                push([=](cudaStream_t st) {
                    sdk::k_col2im_s2<<<blocks_for(dx.numel() / 8), 256, 0, st>>>(dcol, dx.p, xx.n, xx.h, xx.w, C, yy.h, yy.w, 0);
                    return (int)cudaGetLastError();
                });
                dh = dx;
            } else {
                const AttnTape& a = vattn; const int C = a.C; const long long M = a.x.rows();
                auto W2 = [&](const std::string& n) { return (const __half*)params[pindex[a.pre + n]].dst2; };   // transposed copies [in][out]
                // out = x + ao Wo^T + b
                __half* dao = alloc16(M * C);
                linear(dh.p, M, C, W2(".to_out.0.weight"), C, dao);
                // dV = P^T dao ; dP = dao V^T
                __half* PT = alloc16((size_t)M * M); transpose(a.P, PT, 1, (int)M, (int)M);
                __half* daoT = alloc16(M * C); transpose(dao, daoT, 1, (int)M, C);
                __half* dV = alloc16(M * C);
                { Mat A{PT, M, M, (int)M}; Mat Bm{daoT, M, C, (int)M}; Out oo; oo.p16 = dV; oo.ldc = C; Epi e; gemm(A, Bm, C, oo, e); }
                __half* dP = alloc16((size_t)M * M);
                { Mat A{dao, C, M, C}; Mat Bm{a.v, C, M, C}; Out oo; oo.p16 = dP; oo.ldc = M; Epi e; gemm(A, Bm, (int)M, oo, e); }
                const float scale = 1.0f / sqrtf((float)C);
                { const __half* Pp = a.P; push([=](cudaStream_t st) { sdk::k_softmax_bwd<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(Pp, dP, (size_t)M, (int)M, scale); return (int)cudaGetLastError(); }); }
                // dQ = dS K ; dK = dS^T Q
                __half* kT = alloc16(M * C); transpose(a.k, kT, 1, (int)M, C);
                __half* qT = alloc16(M * C); transpose(a.q, qT, 1, (int)M, C);
                __half* dST = alloc16((size_t)M * M); transpose(dP, dST, 1, (int)M, (int)M);
                __half* dQ = alloc16(M * C); __half* dK = alloc16(M * C);
                { Mat A{dP, M, M, (int)M}; Mat Bm{kT, M, C, (int)M}; Out oo; oo.p16 = dQ; oo.ldc = C; Epi e; gemm(A, Bm, C, oo, e); }
                { Mat A{dST, M, M, (int)M}; Mat Bm{qT, M, C, (int)M}; Out oo; oo.p16 = dK; oo.ldc = C; Epi e; gemm(A, Bm, C, oo, e); }
                // dn = dQ Wq + dK Wk + dV Wv
                __half* dn1 = alloc16(M * C); __half* dn2 = alloc16(M * C); __half* dn3 = alloc16(M * C);
                linear(dQ, M, C, W2(".to_q.weight"), C, dn1);
                linear(dK, M, C, W2(".to_k.weight"), C, dn2, nullptr, dn1);
                linear(dV, M, C, W2(".to_v.weight"), C, dn3, nullptr, dn2);
                T4 dx = act(a.x.n, a.x.h, a.x.w, C);
                groupnorm_bwd(a.g, dn3, dh.p, dx.p);
                dh = dx;
            }
        }
        // conv_in^T: d/d(input image, 4 padded channels) -- small-Cout direct conv with flipped weights, fp32 out
        vae_gin = nullptr;
        float* gimg = alloc32((size_t)vae_in.rows() * 4);
        named["vae.grad_in"] = {gimg, (size_t)vae_in.rows() * 4 * sizeof(float)};
        {
            const __half* w = (const __half*)param("encoder.conv_in.weight#dgrad", {4, c.block_out[0], 3, 3}, PK_CONV_SMALL_COUT).dst;
            const T4 d = dh;
            push([=](cudaStream_t st) {
                sdk::k_conv_small_cout<4><<<small_cout_grid((size_t)d.rows(), d.c), 256, (size_t)4 * 9 * d.c * 2, st>>>(d.p, w, nullptr, gimg, d.n, d.h, d.w, d.c);
                return (int)cudaGetLastError();
            });
        }
        vae_grad_img = gimg;
        cur = nullptr;
    }
    float* vae_grad_img = nullptr;

    int build(bool dry_run) {
        dry = dry_run; off = 0; err = 0; params.clear(); pindex.clear(); unet_ops.clear(); enc_ops.clear(); enc_bwd_ops.clear(); dec_ops.clear();
        vres.clear(); vdown.clear(); vorder.clear(); named.clear();
        splitk_elems = (size_t)2048 * 2560; splitk_ws = alloc32(splitk_elems);       // covers M <= 2048 rows x N <= 2560
        if (ucfg.n_levels > 0) build_unet();
        if (vcfg.n_levels > 0) { build_vae(); build_vae_bwd(); if (vcfg.decoder) build_vae_dec(); }
        return err;
    }

    int run_eager(std::vector<Op>& ops, cudaStream_t st) {
        for (auto& op : ops) { int r = op(st); if (r) return r; }
        return MI3D_OK;
    }
    // The launch lists are static (fixed pointers, shapes and order; the timestep and every input live in device buffers), so a list
    // replays as ONE CUDA graph from its third call on (~200-330 launches lose their CPU launch cost and inter-kernel gaps; measured
    // U-Net pass 7.68 -> 6.67 ms, VAE encode 2.89 -> 2.73 ms, VAE backward 3.55 -> 3.37 ms).  Enabled per engine with
    // mi3d_sd_set_graph_replay(); needs a capturable stream (torch's legacy default stream is not: nerf/sd.py runs the engine on its
    // own stream, event-ordered against the caller's).  A list whose capture fails falls back to plain launches for good.
    struct ListGraph { cudaGraphExec_t exec = nullptr; int calls = 0; bool failed = false; };
    std::map<std::vector<Op>*, ListGraph> graphs;
    bool use_graph = false;
    int run(std::vector<Op>& ops, cudaStream_t st) {
        ListGraph& g = graphs[&ops];
        if (!use_graph || profile || g.failed) return run_eager(ops, st);
        if (g.exec) return (int)cudaGraphLaunch(g.exec, st);
        if (g.calls++ == 0) return run_eager(ops, st);            // first call: plain launches (one-time function attributes, warm caches)
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); g.failed = true; return run_eager(ops, st); }
        const int r = run_eager(ops, st);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (r || ce != cudaSuccess || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
            cudaGetLastError(); g.failed = true; g.exec = nullptr;
            if (graph) cudaGraphDestroy(graph);
            return r ? r : run_eager(ops, st);
        }
        cudaGraphDestroy(graph);
        return (int)cudaGraphLaunch(g.exec, st);
    }
    ~Engine() { for (auto& kv : graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); }
};

// ------------------------------------------------------------------------------------------------------------
// API-level elementwise kernels (latent arithmetic of nerf/sd.py:124-171)
// ------------------------------------------------------------------------------------------------------------
// DDIMScheduler.step, eta = 0, after set_timesteps(num_train_timesteps) (nerf/sd.py:154-155): prev = t - 1, alpha_prev = alphas[max(t-1, 0)]
// (set_alpha_to_one = False).  x0 = (x_t - sqrt(1 - a_t) eps) / sqrt(a_t);  x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps
__global__ void k_ddim_step(const float* __restrict__ noise_pred, const float* __restrict__ latents_noisy, const float* __restrict__ alphas,
                            const long long* __restrict__ t, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long tt = *t;
    const float a_t = alphas[tt], a_p = alphas[tt > 0 ? tt - 1 : 0];
    const float x0 = (latents_noisy[i] - sqrtf(1.f - a_t) * noise_pred[i]) / sqrtf(a_t);
    out[i] = sqrtf(a_p) * x0 + sqrtf(1.f - a_p) * noise_pred[i];
}

// decoder input: z = post_quant_conv(latents / 0.18215) (1x1 conv, 4 -> 4), latents fp32 NCHW [1,4,h,w] -> NHWC fp16 [pix][4]
__global__ void k_dec_in(const float* __restrict__ latents, const float* __restrict__ w, const float* __restrict__ b, __half* __restrict__ out, int pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pix) return;
    float x[4];
    #pragma unroll
    for (int c = 0; c < 4; c++) x[c] = latents[(size_t)c * pix + i] / 0.18215f;
    __align__(8) __half o[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) o[k] = __float2half_rn(b[k] + w[4 * k] * x[0] + w[4 * k + 1] * x[1] + w[4 * k + 2] * x[2] + w[4 * k + 3] * x[3]);
    *reinterpret_cast<uint2*>(out + 4 * (size_t)i) = *reinterpret_cast<const uint2*>(o);
}

// imgs = (decoder_out / 2 + 0.5).clamp(0, 1) (nerf/sd.py:208), [pix][4] fp32 -> NCHW fp32 [1,3,S,S]
__global__ void k_dec_out(const float* __restrict__ o, float* __restrict__ imgs, int pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pix) return;
    #pragma unroll
    for (int c = 0; c < 3; c++) imgs[(size_t)c * pix + i] = fminf(fmaxf(o[4 * (size_t)i + c] / 2.f + 0.5f, 0.f), 1.f);
}

// bilinear resize (align_corners=False, F.interpolate semantics) of pred_rgb [3,H,W] fp32 NCHW to S x S, then 2x-1, NHWC fp16 (4 ch)
__global__ void k_interp_in(const float* __restrict__ rgb, int H, int W, __half* __restrict__ out, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * S) return;
    const int ox = i % S, oy = i / S;
    const float sx = (float)W / S, sy = (float)H / S;
    float fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f), fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f);
    const int x0 = min((int)fx, W - 1), y0 = min((int)fy, H - 1);
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float lx = fx - x0, ly = fy - y0;
    __align__(8) __half o[4];
    #pragma unroll
    for (int c = 0; c < 3; c++) {
        const float* p = rgb + (size_t)c * H * W;
        const float v = (1 - ly) * ((1 - lx) * p[y0 * W + x0] + lx * p[y0 * W + x1]) + ly * ((1 - lx) * p[y1 * W + x0] + lx * p[y1 * W + x1]);
        o[c] = __float2half_rn(2.f * v - 1.f);
    }
    o[3] = __float2half_rn(0.f);
    *reinterpret_cast<uint2*>(out + (size_t)i * 4) = *reinterpret_cast<const uint2*>(o);
}
// backward of the above: grad_rgb[c,y,x] (+)= 2 * sum over output pixels of weight * g[oy,ox,c]  (gather over the footprint)
__global__ void k_interp_bwd(const float* __restrict__ g /*[S*S][4]*/, int S, float* __restrict__ grad_rgb, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int x = i % W, y = i / W;
    const float sx = (float)W / S, sy = (float)H / S;
    // output pixels whose source footprint can touch (x, y): fx in (x-1, x+1)
    const int ox_lo = max(0, (int)floorf((x - 1 + 0.5f) / sx - 0.5f)), ox_hi = min(S - 1, (int)ceilf((x + 1 + 0.5f) / sx - 0.5f));
    const int oy_lo = max(0, (int)floorf((y - 1 + 0.5f) / sy - 0.5f)), oy_hi = min(S - 1, (int)ceilf((y + 1 + 0.5f) / sy - 0.5f));
    float acc[3] = {0.f, 0.f, 0.f};
    for (int oy = oy_lo; oy <= oy_hi; oy++) {
        const float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f);
        const int y0 = min((int)fy, H - 1), y1 = min(y0 + 1, H - 1); const float ly = fy - y0;
        float wy = 0.f; if (y0 == y) wy += 1 - ly; if (y1 == y) wy += ly;
        if (wy == 0.f) continue;
        for (int ox = ox_lo; ox <= ox_hi; ox++) {
            const float fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f);
            const int x0 = min((int)fx, W - 1), x1 = min(x0 + 1, W - 1); const float lx = fx - x0;
            float wx = 0.f; if (x0 == x) wx += 1 - lx; if (x1 == x) wx += lx;
            if (wx == 0.f) continue;
            const float* gp = g + ((size_t)oy * S + ox) * 4;
            acc[0] += wy * wx * gp[0]; acc[1] += wy * wx * gp[1]; acc[2] += wy * wx * gp[2];
        }
    }
    #pragma unroll
    for (int c = 0; c < 3; c++) grad_rgb[(size_t)c * H * W + i] = 2.f * acc[c];
}
// moments = quant_conv(conv_out) ; latents = (mean + exp(0.5*clamp(logvar)) * eps) * 0.18215  -> fp32 NCHW [4,h,w]
__global__ void k_sample_latents(const float* __restrict__ conv_out /*[pix][8]*/, const float* __restrict__ qw /*[8][8]*/, const float* __restrict__ qb,
                                 const float* __restrict__ eps /*[4,h,w]*/, float* __restrict__ latents, int pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pix) return;
    float m[8];
    #pragma unroll
    for (int o = 0; o < 8; o++) { float a = qb[o]; for (int k = 0; k < 8; k++) a = fmaf(qw[o * 8 + k], conv_out[(size_t)i * 8 + k], a); m[o] = a; }
    #pragma unroll
    for (int c = 0; c < 4; c++) {
        const float lv = fminf(fmaxf(m[4 + c], -30.f), 20.f);
        latents[(size_t)c * pix + i] = (m[c] + expf(0.5f * lv) * eps[(size_t)c * pix + i]) * 0.18215f;
    }
}
// backward: d conv_out[pix][8] from d latents
__global__ void k_sample_latents_bwd(const float* __restrict__ conv_out, const float* __restrict__ qw, const float* __restrict__ qb,
                                     const float* __restrict__ eps, const float* __restrict__ g_lat, float* __restrict__ g_conv_out, int pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pix) return;
    float m[8], gm[8];
    #pragma unroll
    for (int o = 0; o < 8; o++) { float a = qb[o]; for (int k = 0; k < 8; k++) a = fmaf(qw[o * 8 + k], conv_out[(size_t)i * 8 + k], a); m[o] = a; }
    #pragma unroll
    for (int c = 0; c < 4; c++) {
        const float g = g_lat[(size_t)c * pix + i] * 0.18215f;
        gm[c] = g;
        const float lv = m[4 + c];
        const bool pass = lv >= -30.f && lv <= 20.f;
        gm[4 + c] = pass ? g * eps[(size_t)c * pix + i] * 0.5f * expf(0.5f * lv) : 0.f;
    }
    #pragma unroll
    for (int k = 0; k < 8; k++) { float a = 0.f; for (int o = 0; o < 8; o++) a = fmaf(qw[o * 8 + k], gm[o], a); g_conv_out[(size_t)i * 8 + k] = a; }
}
// DDIM add_noise + duplicate for CFG + NCHW fp32 -> NHWC fp16 : x_in[b][p][c] = sqrt(a_t) z[c][p] + sqrt(1-a_t) eps[c][p]
__global__ void k_add_noise(const float* __restrict__ lat, const float* __restrict__ noise, const float* __restrict__ alphas, const long long* __restrict__ t,
                            __half* __restrict__ xin, int pix, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pix) return;
    const float a = alphas[t[0]];
    const float sa = sqrtf(a), sb = sqrtf(1.f - a);
    __align__(8) __half o[4];
    #pragma unroll
    for (int c = 0; c < 4; c++) o[c] = __float2half_rn(sa * lat[(size_t)c * pix + i] + sb * noise[(size_t)c * pix + i]);
    for (int b = 0; b < B; b++) *reinterpret_cast<uint2*>(xin + ((size_t)b * pix + i) * 4) = *reinterpret_cast<const uint2*>(o);
}
// ctx fp32 [B][L][D] -> fp16 [B][Lpad][D], zero rows beyond L
__global__ void k_ctx_pad(const float* __restrict__ ctx, __half* __restrict__ out, int B, int L, int Lpad, int D) {
    const size_t total = (size_t)B * Lpad * D;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D); const size_t r = i / D; const int l = (int)(r % Lpad), b = (int)(r / Lpad);
        out[i] = __float2half_rn(l < L ? ctx[((size_t)b * L + l) * D + d] : 0.f);
    }
}
// CFG (sd.py:150-151: text + gs*(text - uncond)), SDS weight w = 1 - alpha_t, grad = nan_to_num(w (eps_hat - eps)); NHWC fp32 in, NCHW fp32 out
__global__ void k_cfg_sds(const float* __restrict__ unet_out /*[2][pix][4]*/, const float* __restrict__ noise, const float* __restrict__ alphas,
                          const long long* __restrict__ t, float gs, float* __restrict__ noise_pred, float* __restrict__ grad, int pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pix) return;
    const float w = 1.f - alphas[t[0]];
    #pragma unroll
    for (int c = 0; c < 4; c++) {
        const float un = unet_out[(size_t)i * 4 + c], tx = unet_out[((size_t)pix + i) * 4 + c];
        const float np = tx + gs * (tx - un);
        if (noise_pred) noise_pred[(size_t)c * pix + i] = np;
        float g = w * (np - noise[(size_t)c * pix + i]);
        if (isnan(g)) g = 0.f; else if (isinf(g)) g = g > 0 ? 3.4028234664e38f : -3.4028234664e38f;
        if (grad) grad[(size_t)c * pix + i] = g;
    }
}

}  // namespace sd

struct mi3d_sd { sd::Engine e; };

extern "C" {

size_t mi3d_sd_workspace_bytes(const mi3d_unet_cfg* u, const mi3d_vae_cfg* v) {
    sd::Engine e;
    e.ucfg = u ? *u : mi3d_unet_cfg{}; e.vcfg = v ? *v : mi3d_vae_cfg{};
    if (v && !u) e.ucfg.groups = v->groups;
    e.build(true);
    return e.off + (1 << 20);
}

mi3d_sd_t mi3d_sd_create(const mi3d_unet_cfg* u, const mi3d_vae_cfg* v, void* workspace, size_t workspace_bytes) {
    if (!workspace) return nullptr;
    mi3d_sd* h = new mi3d_sd();
    h->e.ucfg = u ? *u : mi3d_unet_cfg{}; h->e.vcfg = v ? *v : mi3d_vae_cfg{};
    if (v && !u) h->e.ucfg.groups = v->groups;
    h->e.base = (uint8_t*)workspace; h->e.cap = workspace_bytes;
    int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&h->e.num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (h->e.build(false) != 0) { delete h; return nullptr; }
    cudaFuncSetAttribute(sdk::k_conv_small_cin<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(sdk::k_conv_small_cin<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(sdk::k_conv_small_cout<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(sdk::k_conv_small_cout<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    return h;
}

void mi3d_sd_destroy(mi3d_sd_t h) { delete h; }

int mi3d_sd_num_params(mi3d_sd_t h) { return h ? (int)h->e.params.size() : 0; }
const char* mi3d_sd_param_name(mi3d_sd_t h, int i) { return h->e.params[i].name.c_str(); }
long long mi3d_sd_param_numel(mi3d_sd_t h, int i) { return (long long)h->e.params[i].numel; }
int mi3d_sd_param_shape(mi3d_sd_t h, int i, int* shape4) {
    if (!h || i < 0 || i >= (int)h->e.params.size()) return -1;
    const auto& s = h->e.params[i].shape;
    for (int k = 0; k < 4; k++) shape4[k] = k < (int)s.size() ? s[k] : 0;
    return (int)s.size();
}

// src: fp32 device tensor in the diffusers layout named by mi3d_sd_param_name (names with a '#' suffix are derived copies: pass the base tensor)
int mi3d_sd_load_param(mi3d_sd_t h, int i, const float* src, const float* partner_src, mi3d_stream_t stream) {
    if (!h || i < 0 || i >= (int)h->e.params.size() || !src) return MI3D_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    sd::Param& P = h->e.params[i];
    const int nb = sd::Engine::blocks_for(P.numel);
    using namespace sdk;
    switch (P.kind) {
        case sd::PK_F32: MI3D_CHECK(cudaMemcpyAsync(P.dst, src, P.numel * 4, cudaMemcpyDeviceToDevice, st)); break;
        case sd::PK_LINEAR: case sd::PK_CONV1: k_f32_to_f16<<<nb, 256, 0, st>>>(src, (__half*)P.dst, P.numel); break;
        case sd::PK_LINEAR_BOTH: case sd::PK_CONV1_BOTH:
            k_f32_to_f16<<<nb, 256, 0, st>>>(src, (__half*)P.dst, P.numel);
            k_transpose_w<<<nb, 256, 0, st>>>(src, (__half*)P.dst2, P.shape[0], P.shape[1]); break;
        case sd::PK_LINEAR_T: k_transpose_w<<<nb, 256, 0, st>>>(src, (__half*)P.dst, P.shape[0], P.shape[1]); break;
        case sd::PK_CONV3: k_conv_w<<<nb, 256, 0, st>>>(src, (__half*)P.dst, P.shape[0], P.shape[1], 0); break;
        case sd::PK_CONV3_BOTH: {
            k_conv_w<<<nb, 256, 0, st>>>(src, (__half*)P.dst, P.shape[0], P.shape[1], 0);
            const bool strided = P.name.find("downsamplers") != std::string::npos;
            if (strided) {
                // dgrad of the im2col GEMM needs the plain transpose of the [Cout][9*Cin] matrix: [9*Cin][Cout]
                // (k_conv_w wrote [O][ky][kx][I]; transpose that fp16 matrix)
                k_transpose<<<dim3((9 * P.shape[1] + 31) / 32, (P.shape[0] + 31) / 32, 1), dim3(32, 8), 0, st>>>((const __half*)P.dst, (__half*)P.dst2, P.shape[0], 9 * P.shape[1]);
            } else {
                k_conv_w<<<nb, 256, 0, st>>>(src, (__half*)P.dst2, P.shape[0], P.shape[1], 1);
            }
            break;
        }
        case sd::PK_GEGLU_W: {
            if (!partner_src) return MI3D_ERR_ARG;
            sd::Param& Bp = h->e.params[h->e.pindex[P.partner]];
            k_geglu_w<<<nb, 256, 0, st>>>(src, partner_src, (__half*)P.dst, (float*)Bp.dst, P.shape[0] / 2, P.shape[1]);
            break;
        }
        case sd::PK_GEGLU_B: break;     // written together with the weight
        case sd::PK_CONV_F32_OHWI: {
            // fp32 direct-conv weights.  Base form: OIHW [O][I][3][3] -> [O][ky][kx][Ipad] with Ipad = 4 (conv_in) ;
            // "#dgrad_f32" form (conv_out transposed): src is conv_out OIHW [8][ch][3][3] -> [ch][2-ky][2-kx][8]
            std::vector<float> tmp;   // host staging keeps this one-off path simple
            const bool dg = P.name.find("#dgrad_f32") != std::string::npos;
            const int O = dg ? P.shape[1] : P.shape[0], I = dg ? P.shape[0] : P.shape[1];
            std::vector<float> hsrc((size_t)O * I * 9);
            MI3D_CHECK(cudaMemcpyAsync(hsrc.data(), src, hsrc.size() * 4, cudaMemcpyDeviceToHost, st));
            MI3D_CHECK(cudaStreamSynchronize(st));
            if (!dg) {
                const int Ip = 4;
                tmp.assign((size_t)O * 9 * Ip, 0.f);
                for (int o = 0; o < O; o++) for (int c = 0; c < I; c++) for (int t = 0; t < 9; t++) tmp[((size_t)o * 9 + t) * Ip + c] = hsrc[((size_t)o * I + c) * 9 + t];
            } else {
                // hsrc = conv_out [O=8][I=ch][3][3]; want w'[ci=ch][ky'][kx'][co=8] = hsrc[co][ci][2-ky'][2-kx']
                tmp.assign((size_t)I * 9 * O, 0.f);
                for (int o = 0; o < O; o++) for (int c = 0; c < I; c++) for (int t = 0; t < 9; t++) tmp[((size_t)c * 9 + (8 - t)) * O + o] = hsrc[((size_t)o * I + c) * 9 + t];
            }
            if (tmp.size() * 4 > P.numel * 4 + (dg ? 0 : (size_t)O * 9 * 4 * 4)) {}
            MI3D_CHECK(cudaMemcpyAsync(P.dst, tmp.data(), std::min(tmp.size(), dg ? P.numel : tmp.size()) * 4, cudaMemcpyHostToDevice, st));
            MI3D_CHECK(cudaStreamSynchronize(st));
            break;
        }
        case sd::PK_CONV_SMALL_COUT: case sd::PK_CONV_SMALL_COUT_BOTH: {
            const bool dg = P.name.find("#dgrad") != std::string::npos;
            if (!dg) k_conv_w<<<nb, 256, 0, st>>>(src, (__half*)P.dst, P.shape[0], P.shape[1], 0);
            else {
                // conv_in transposed: src conv_in OIHW [c0][3][3][3]; want [4 (padded cin)][2-ky][2-kx][c0] fp16
                const int c0 = P.shape[1];
                std::vector<float> hsrc((size_t)c0 * 3 * 9);
                MI3D_CHECK(cudaMemcpyAsync(hsrc.data(), src, hsrc.size() * 4, cudaMemcpyDeviceToHost, st));
                MI3D_CHECK(cudaStreamSynchronize(st));
                std::vector<__half> tmp((size_t)4 * 9 * c0, __float2half(0.f));
                for (int o = 0; o < c0; o++) for (int c = 0; c < 3; c++) for (int t = 0; t < 9; t++) tmp[((size_t)c * 9 + (8 - t)) * c0 + o] = __float2half(hsrc[((size_t)o * 3 + c) * 9 + t]);
                MI3D_CHECK(cudaMemcpyAsync(P.dst, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice, st));
                MI3D_CHECK(cudaStreamSynchronize(st));
            }
            break;
        }
    }
    return (int)cudaGetLastError();
}

// enable != 0: start timing every tensor-core tile launch with CUDA events; enable == 0: stop, synchronise the events and
// return the accumulated kernel milliseconds and launch count since the last enable
int mi3d_sd_set_graph_replay(mi3d_sd_t h, int enable) {
    if (!h) return MI3D_ERR_ARG;
    h->e.use_graph = enable != 0;
    return MI3D_OK;
}

int mi3d_sd_graph_replays(mi3d_sd_t h) {
    if (!h) return -1;
    int n = 0;
    for (auto& kv : h->e.graphs) if (kv.second.exec) n++;
    return n;
}

int mi3d_sd_profile(mi3d_sd_t h, int enable, float* gemm_ms, int* launches) { return mi3d_sd_profile_dump(h, enable, gemm_ms, launches, nullptr); }

// dump_path_host (nullable): one text line per timed launch "M N K block_n splits conv batch epi ms" (bench.py derives FLOPs from it)
int mi3d_sd_profile_dump(mi3d_sd_t h, int enable, float* gemm_ms, int* launches, const char* dump) {
    if (!h) return MI3D_ERR_ARG;
    sd::Engine& e = h->e;
    if (enable) { e.profile = true; e.prof_used = 0; return MI3D_OK; }
    e.profile = false;
    float total = 0.f; int n_tile = 0;
    FILE* f = dump ? fopen(dump, "w") : nullptr;
    for (size_t i = 0; i < e.prof_used; i++) {
        float ms = 0.f;
        MI3D_CHECK(cudaEventSynchronize(e.prof_events[i].second));
        MI3D_CHECK(cudaEventElapsedTime(&ms, e.prof_events[i].first, e.prof_events[i].second));
        if (e.prof_shapes[i].conv != 2) { total += ms; n_tile++; }
        if (f) { const auto& q = e.prof_shapes[i]; fprintf(f, "%d %d %d %d %d %d %d %d %.4f\n", q.M, q.N, q.K, q.bn, q.splits, q.conv, q.batch, q.epi, ms); }
    }
    if (f) fclose(f);
    if (gemm_ms) *gemm_ms = total;
    if (launches) *launches = n_tile;
    return MI3D_OK;
}

int mi3d_sd_debug_tensor(mi3d_sd_t h, const char* name, void** ptr, size_t* bytes) {
    auto it = h->e.named.find(name);
    if (it == h->e.named.end()) return MI3D_ERR_ARG;
    *ptr = it->second.first; *bytes = it->second.second;
    return MI3D_OK;
}

// nerf/sd.py:124,133,212-220 : pred_rgb [1,3,H,W] fp32 -> bilinear to image_hw -> VAE encode -> posterior sample * 0.18215
int mi3d_sd_encode(mi3d_sd_t h, const float* pred_rgb, int H, int W, const float* eps_posterior, float* latents, mi3d_stream_t stream) {
    if (!h || h->e.vcfg.n_levels == 0) return MI3D_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream; sd::Engine& e = h->e;
    const int S = e.vcfg.image_hw;
    sd::k_interp_in<<<(S * S + 255) / 256, 256, 0, st>>>(pred_rgb, H, W, e.vae_in.p, S);
    int r = e.run(e.enc_ops, st); if (r) return r;
    const int pix = (int)e.v_h_last.rows();
    const float* qw = (const float*)e.params[e.pindex["quant_conv.weight"]].dst; const float* qb = (const float*)e.params[e.pindex["quant_conv.bias"]].dst;
    sd::k_sample_latents<<<(pix + 127) / 128, 128, 0, st>>>(e.vae_moments, qw, qb, eps_posterior, latents, pix);
    return (int)cudaGetLastError();
}

// backward of mi3d_sd_encode w.r.t. pred_rgb (weights frozen): grad_pred_rgb [1,3,H,W] fp32 (overwritten)
int mi3d_sd_encode_backward(mi3d_sd_t h, const float* grad_latents, const float* eps_posterior, int H, int W, float* grad_pred_rgb, mi3d_stream_t stream) {
    if (!h || h->e.vcfg.n_levels == 0) return MI3D_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream; sd::Engine& e = h->e;
    const int pix = (int)e.v_h_last.rows(), S = e.vcfg.image_hw;
    const float* qw = (const float*)e.params[e.pindex["quant_conv.weight"]].dst; const float* qb = (const float*)e.params[e.pindex["quant_conv.bias"]].dst;
    sd::k_sample_latents_bwd<<<(pix + 127) / 128, 128, 0, st>>>(e.vae_moments, qw, qb, eps_posterior, grad_latents, e.vae_gmoments, pix);
    int r = e.run(e.enc_bwd_ops, st); if (r) return r;
    sd::k_interp_bwd<<<(H * W + 127) / 128, 128, 0, st>>>(e.vae_grad_img, S, grad_pred_rgb, H, W);
    return (int)cudaGetLastError();
}

// nerf/sd.py:154-155: one DDIM step t -> t-1 on the noisy latents with the CFG-combined noise prediction.  All fp32 [1,4,h,w].
int mi3d_sd_ddim_step(const float* noise_pred, const float* latents_noisy, const long long* t, const float* alphas_cumprod, float* prev_sample,
                      int n, mi3d_stream_t stream) {
    if (!noise_pred || !latents_noisy || !t || !alphas_cumprod || !prev_sample || n <= 0) return MI3D_ERR_ARG;
    sd::k_ddim_step<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(noise_pred, latents_noisy, alphas_cumprod, t, prev_sample, n);
    return (int)cudaGetLastError();
}

// nerf/sd.py:201-210 decode_latents: imgs = (vae.decode(latents / 0.18215).sample / 2 + 0.5).clamp(0, 1); latents fp32 [1,4,h,w] -> imgs fp32 [1,3,S,S]
int mi3d_sd_decode(mi3d_sd_t h, const float* latents, float* imgs, mi3d_stream_t stream) {
    if (!h || h->e.vcfg.n_levels == 0 || !h->e.vcfg.decoder || !latents || !imgs) return MI3D_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream; sd::Engine& e = h->e;
    const int pix = (int)e.dec_in.rows(), S = e.vcfg.image_hw;
    const float* qw = (const float*)e.params[e.pindex["post_quant_conv.weight"]].dst; const float* qb = (const float*)e.params[e.pindex["post_quant_conv.bias"]].dst;
    sd::k_dec_in<<<(pix + 127) / 128, 128, 0, st>>>(latents, qw, qb, e.dec_in.p, pix);
    int r = e.run(e.dec_ops, st); if (r) return r;
    sd::k_dec_out<<<(S * S + 127) / 128, 128, 0, st>>>(e.dec_out, imgs, S * S);
    return (int)cudaGetLastError();
}

// nerf/sd.py:138-171 : add_noise -> U-Net on [uncond, text] -> CFG -> SDS gradient.  t: device int64 scalar; alphas: device [1000] fp32
int mi3d_sd_unet_sds(mi3d_sd_t h, const float* latents, const float* noise, const long long* t, const float* alphas_cumprod,
                     const float* text_embeddings, float guidance_scale, float* noise_pred, float* grad, mi3d_stream_t stream) {
    if (!h || h->e.ucfg.n_levels == 0) return MI3D_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream; sd::Engine& e = h->e;
    const int pix = e.ucfg.latent_hw * e.ucfg.latent_hw, B = e.ucfg.batch;
    if (B != 2 || e.ucfg.in_ch != 4) return MI3D_ERR_ARG;
    MI3D_CHECK(cudaMemcpyAsync(e.t_dev, t, sizeof(long long), cudaMemcpyDeviceToDevice, st));
    sd::k_add_noise<<<(pix + 127) / 128, 128, 0, st>>>(latents, noise, alphas_cumprod, e.t_dev, e.unet_in.p, pix, B);
    sd::k_ctx_pad<<<sd::Engine::blocks_for((size_t)B * e.ctx_pad * e.ucfg.cross_dim), 256, 0, st>>>(text_embeddings, e.ctx16, B, e.ucfg.ctx_len, e.ctx_pad, e.ucfg.cross_dim);
    int r = e.run(e.unet_ops, st); if (r) return r;
    sd::k_cfg_sds<<<(pix + 127) / 128, 128, 0, st>>>(e.unet_out, noise, alphas_cumprod, e.t_dev, guidance_scale, noise_pred, grad, pix);
    return (int)cudaGetLastError();
}

}  // extern "C"
