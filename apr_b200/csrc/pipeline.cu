// pipeline.cu — native driver of the whole hot path: pyramid (collate) + KFE encoder blocks on one stream.
//
// Host-side mirror, in C++, of collate_fn_descriptor (/root/reference/Predator_APR/datasets/dataloader.py:72-198) and of
// the encoder loop of KPFCNN.forward (models/architectures.py:149-153) with SimpleBlock / ResnetBottleneckBlock.forward
// (models/blocks.py:581-593, :653-681). Every operation is one of this library's own C-ABI entry points; this file
// only sequences them, carves one caller-provided arena, and reads the three per-level point counts back through
// events so the host never waits for the encoder kernels it has already queued.
#include "common.cuh"
#include <new>
#include <vector>

namespace aprb {
constexpr int KFE_MAX_LEVELS = 8;
extern int g_fuse_stats;
extern int g_kpconv_f16;
extern int g_act_f16;
extern int g_dbg_skip_d2h;
extern int g_host_zero_copy;
extern int g_gemm_apply;
// aprb_set_option("blocking_sync"): the events a host thread waits on (per-level counts, host-output copies) are created
// with cudaEventBlockingSync, so a waiting thread sleeps instead of spinning. One process per GPU with several calls in
// flight each puts more waiting threads on the box than it has cores (8 ranks x 5 streams on 32 cores); spinning
// waiters then take the cores from the threads that launch kernels.
int g_blocking_sync = 1;
}

struct aprb_kfe {
    aprb_kfe_config cfg;
    std::vector<aprb_kfe_block> blocks;
    cudaEvent_t ev[aprb::KFE_MAX_LEVELS];
    int* h_counts;  // pinned: per level [M, status, lens[0..B)]
    int h_counts_cap;
    // state of the last forward (aprb_kfe_get)
    int B;
    int n[aprb::KFE_MAX_LEVELS];
    const float* pts[aprb::KFE_MAX_LEVELS];
    const int* lens[aprb::KFE_MAX_LEVELS];
    const int* conv[aprb::KFE_MAX_LEVELS];
    const int* pool[aprb::KFE_MAX_LEVELS];
    const int* up[aprb::KFE_MAX_LEVELS];
    const int* seg[aprb::KFE_MAX_LEVELS];   // per level: row offsets of the S normalisation segments (S+1 ints), or NULL
    const int* pool_w[aprb::KFE_MAX_LEVELS]; // per level: [S] width of the reference's pool matrix per segment, min(max_count, limit)
    int S;
    // per-block taps (aprb_kfe_set_tap): device copies of every block's output, for per-block parity tests
    char* tap_buf; size_t tap_cap, tap_off;
    struct Tap { size_t off; int rows, cols, f16, tag; };   // tag = 4 * block + kind (0 block output, 1 KPConv output, 2 KPConv input)
    std::vector<Tap> taps;
    struct BlockOut { const void* p; int rows, cols, f16; };
    std::vector<BlockOut> block_out;        // every encoder block's output of the last forward (lives in the arena)
    // fp16 activation mode: fp16 copies of the unary weights, owned by the handle (cudaMalloc at create), per block
    // [unary1, unary2, shortcut]; act16_ok = every block fits the fp16 kernels' shape constraints
    std::vector<void*> w16;
    bool act16_ok;
    // asynchronous host output (aprb_kfe_forward_host_async): the last block writes into out_dev[parity], a side stream
    // copies it to the caller's host buffer while the main stream already runs the next call
    float* out_dev[2];
    size_t out_dev_floats;
    cudaStream_t copy_st;
    cudaEvent_t ev_done, ev_copied[2];
    int parity;
    float* y_last;            // where the final block must write (NULL: carve from the arena)
    int y_last_rows_cap;
    int host_out_f16;         // aprb_kfe_set_host_output_f16: the asynchronous host output is fp16
    int y_last_f16;           // set while such a call is being queued: the final block stores fp16 into y_last
};

using namespace aprb;

namespace {

struct Arena {
    char* base;
    size_t off, cap;
    bool overflow;
    Arena(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes), overflow(false) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = align256(n * sizeof(T));
        if (off + bytes > cap) { overflow = true; return nullptr; }
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
    // everything after the persistent allocations is scratch for the op being launched (stream order makes reuse safe)
    void* scratch() const { return base + off; }
    size_t scratch_bytes() const { return cap - off; }
};

#define KFE_OK(expr)            \
    do {                        \
        int _rc = (expr);       \
        if (_rc) return _rc;    \
    } while (0)
#define KFE_ALLOC(var, T, n)                                                                     \
    T* var = A.take<T>(n);                                                                       \
    if (!var) { set_error("aprb_kfe_forward: arena too small (need > %zu bytes)", A.cap); return APRB_ERR_WORKSPACE; }

// Group statistics of a GEMM-produced tensor, handed from the producer's epilogue to the normalisation that follows
// (gs == nullptr: the consumer scans the tensor itself).
struct GStat {
    float* buf = nullptr;
    int written = 0;
    const float* get() const { return written ? buf : nullptr; }
};

// Debug tap (aprb_kfe_set_tap): stream-ordered device copy of an intermediate tensor of the running forward.
int tap_push(const aprb_kfe& hc, const void* p, int rows, int cols, int f16, int tag, cudaStream_t st) {
    aprb_kfe& h = const_cast<aprb_kfe&>(hc);
    if (!h.tap_buf) return APRB_OK;
    const size_t bytes = (size_t)rows * cols * (f16 ? 2 : 4);
    if (h.tap_off + bytes > h.tap_cap) { set_error("aprb_kfe_forward: tap buffer too small (tag %d needs %zu more bytes)", tag, bytes); return APRB_ERR_WORKSPACE; }
    APRB_CUDA_OK(cudaMemcpyAsync(h.tap_buf + h.tap_off, p, bytes, cudaMemcpyDeviceToDevice, st));
    h.taps.push_back({h.tap_off, rows, cols, f16, tag});
    h.tap_off += align256(bytes);
    return APRB_OK;
}

int kpconv_call(const aprb_kfe& h, const aprb_kfe_block& b, const float* q, const float* s, const int* idx, int ld,
                const float* x, int nq, int ns, int H, int cin, int cout, float* out, GStat* gs, Arena& A, cudaStream_t st) {
    size_t need = aprb_kpconv_ws_bytes(nq, ns, H, h.cfg.K, cin, cout);
    if (A.scratch_bytes() < need) { set_error("aprb_kfe_forward: arena too small for the KPConv workspace"); return APRB_ERR_WORKSPACE; }
    const bool f16 = b.kp_Wprep16 && aprb::g_kpconv_f16;
    return aprb_kpconv_forward_stats(q, s, idx, 0, ld, x, b.kp, b.kp_W, f16 ? (const float*)b.kp_Wprep16 : b.kp_Wprep, b.extent,
                                     nq, ns, H, h.cfg.K, cin, cout, out, f16 ? 3 : (b.kp_Wprep ? 0 : 1),
                                     gs ? gs->buf : nullptr, gs ? &gs->written : nullptr, A.scratch(), A.scratch_bytes(), st);
}

// InstanceNorm over the rows of each collated pair (segment) of level `lvl`
int norm_call(const aprb_kfe& h, int lvl, const float* x, int n, int c, float slope, const float* res, int norm_res, float* y,
              const GStat* gx, const GStat* gres, Arena& A, cudaStream_t st) {
    if (A.scratch_bytes() < aprb_instnorm_seg_ws_bytes(n, c, h.S)) { set_error("aprb_kfe_forward: arena too small for the norm workspace"); return APRB_ERR_WORKSPACE; }
    return aprb_instnorm_lrelu_seg_pre(x, n, c, h.seg[lvl], h.S, 1e-5f, slope, res, norm_res, 1, y, gx ? gx->get() : nullptr,
                                       gres ? gres->get() : nullptr, A.scratch(), A.scratch_bytes(), st);
}

int linear_call(const float* x, const float* W, int n, int cin, int cout, float* y, GStat* gs, Arena& A, cudaStream_t st) {
    size_t need = aprb_linear_tf32_ws_bytes(n, cin, cout);
    void* ws = A.scratch_bytes() >= need ? A.scratch() : nullptr;   // split-K is optional
    return aprb_linear_tf32_stats(x, W, n, cin, cout, y, gs ? gs->buf : nullptr, gs ? &gs->written : nullptr, ws,
                                  ws ? A.scratch_bytes() : 0, st);
}

// One encoder block. feat [ns_rows, in_dim] -> *out [nq, out_cols]. Temporaries are released when the block returns.
int run_block(const aprb_kfe& h, const aprb_kfe_block& b, const float* feat, const float** out, int* out_cols, Arena& A,
              cudaStream_t st) {
    const int l = b.layer;
    const int nq = b.strided ? h.n[l + 1] : h.n[l], ns = h.n[l];
    const int lq = b.strided ? l + 1 : l;                          // level of the block's output rows
    const float* q = b.strided ? h.pts[l + 1] : h.pts[l];
    const float* s = h.pts[l];
    const int* idx = b.strided ? h.pool[l] : h.conv[l];
    const int H = h.cfg.limits[l];
#define KFE_GSTAT(var, rows, cols)                                                                       \
    GStat var;                                                                                               \
    if (aprb::g_fuse_stats) {                                                                                \
        var.buf = (float*)A.take<char>(aprb_group_stats_bytes(rows, cols));                                  \
        if (!var.buf) { set_error("aprb_kfe_forward: arena too small (need > %zu bytes)", A.cap); return APRB_ERR_WORKSPACE; } \
    }
    if (b.type == 0) {                                             // SimpleBlock: KPConv -> IN -> LeakyReLU
        const int cout = b.out_dim / 2;
        KFE_ALLOC(y, float, (size_t)nq * cout);
        const size_t mark = A.off;
        KFE_ALLOC(t, float, (size_t)nq * cout);
        KFE_GSTAT(gt, nq, cout);
        KFE_OK(kpconv_call(h, b, q, s, idx, H, feat, nq, ns, H, b.in_dim, cout, t, &gt, A, st));
        KFE_OK(norm_call(h, lq, t, nq, cout, 0.1f, nullptr, 0, y, &gt, nullptr, A, st));
        A.off = mark;
        *out = y; *out_cols = cout;
        return APRB_OK;
    }
    const int mid = b.out_dim / 4, cout = b.out_dim;               // ResnetBottleneckBlock
    float* y = nullptr;
    if (&b == &h.blocks.back() && h.y_last) {                      // async host output: dedicated ping-pong buffer
        if (nq > h.y_last_rows_cap) { set_error("aprb_kfe_forward: output has %d rows, buffer holds %d", nq, h.y_last_rows_cap); return APRB_ERR_WORKSPACE; }
        y = h.y_last;
    } else {
        y = A.take<float>((size_t)nq * cout);
        if (!y) { set_error("aprb_kfe_forward: arena too small (need > %zu bytes)", A.cap); return APRB_ERR_WORKSPACE; }
    }
    const size_t mark = A.off;
    const float* x1 = feat;
    if (b.unary1_W) {                                              // unary1: Linear -> IN -> LeakyReLU
        KFE_ALLOC(t1, float, (size_t)ns * mid);
        KFE_GSTAT(g1, ns, mid);
        KFE_OK(linear_call(feat, b.unary1_W, ns, b.in_dim, mid, t1, &g1, A, st));
        KFE_OK(norm_call(h, l, t1, ns, mid, 0.1f, nullptr, 0, t1, &g1, nullptr, A, st));
        x1 = t1;
    }
    KFE_ALLOC(t2, float, (size_t)nq * mid);
    KFE_GSTAT(g2, nq, mid);
    KFE_OK(kpconv_call(h, b, q, s, idx, H, x1, nq, ns, H, mid, mid, t2, &g2, A, st));
    KFE_OK(norm_call(h, lq, t2, nq, mid, 0.1f, nullptr, 0, t2, &g2, nullptr, A, st));
    KFE_ALLOC(t3, float, (size_t)nq * cout);
    KFE_GSTAT(g3, nq, cout);
    KFE_OK(linear_call(t2, b.unary2_W, nq, mid, cout, t3, &g3, A, st)); // unary2 (IN folded into the final kernel)
    const float* sc = feat;
    if (b.strided) {                                               // shortcut = max_pool(features, pools)
        KFE_ALLOC(mp, float, (size_t)nq * b.in_dim);
        KFE_OK(aprb_max_pool_seg(feat, 0, idx, 0, H, nq, ns, H, b.in_dim, h.seg[lq], h.S, h.pool_w[l], mp, st));
        sc = mp;
    }
    if (b.shortcut_W) {                                            // LeakyReLU(IN(x3) + IN(Linear(sc)))
        KFE_ALLOC(t4, float, (size_t)nq * cout);
        KFE_GSTAT(g4, nq, cout);
        KFE_OK(linear_call(sc, b.shortcut_W, nq, b.in_dim, cout, t4, &g4, A, st));
        KFE_OK(norm_call(h, lq, t3, nq, cout, 0.1f, t4, 1, y, &g3, &g4, A, st));
    } else {                                                       // LeakyReLU(IN(x3) + sc)
        KFE_OK(norm_call(h, lq, t3, nq, cout, 0.1f, sc, 0, y, &g3, nullptr, A, st));
    }
#undef KFE_GSTAT
    A.off = mark;
    *out = y; *out_cols = cout;
    return APRB_OK;
}

// ---- fp16 activation mode -------------------------------------------------------------------------------------------
// Every normalised activation (the output of an InstanceNorm + LeakyReLU, already rounded to TF32's 10-bit mantissa) is
// stored in fp16 — exact for these values — so it costs half the bytes to write, to gather in KPConv / max-pool and to
// stream into the tensor cores (kind::f16). GEMM outputs, which feed the statistics, stay fp32.
int norm16_call(const aprb_kfe& h, int lvl, const float* x, int n, int c, const void* res, int res16, int norm_res, void* y,
                int out16, const GStat* gx, const GStat* gres, Arena& A, cudaStream_t st) {
    if (A.scratch_bytes() < aprb_instnorm_seg_ws_bytes(n, c, h.S)) { set_error("aprb_kfe_forward: arena too small for the norm workspace"); return APRB_ERR_WORKSPACE; }
    return aprb_instnorm_lrelu_seg_f16(x, n, c, h.seg[lvl], h.S, 1e-5f, 0.1f, res, res16, norm_res, 1, y, out16,
                                       gx ? gx->get() : nullptr, gres ? gres->get() : nullptr, A.scratch(), A.scratch_bytes(), st);
}

int run_block16(const aprb_kfe& h, size_t bi, const void* feat, bool feat16, bool last, const void** out, int* out_cols,
                Arena& A, cudaStream_t st) {
    const aprb_kfe_block& b = h.blocks[bi];
    const int l = b.layer;
    const int nq = b.strided ? h.n[l + 1] : h.n[l], ns = h.n[l];
    const int lq = b.strided ? l + 1 : l;
    const float* q = b.strided ? h.pts[l + 1] : h.pts[l];
    const float* s = h.pts[l];
    const int* idx = b.strided ? h.pool[l] : h.conv[l];
    const int H = h.cfg.limits[l];
    const int out16 = (last && !h.y_last_f16) ? 0 : 1;
#define KFE_GSTAT(var, rows, cols)                                                                       \
    GStat var;                                                                                               \
    if (aprb::g_fuse_stats) {                                                                                \
        var.buf = (float*)A.take<char>(aprb_group_stats_bytes(rows, cols));                                  \
        if (!var.buf) { set_error("aprb_kfe_forward: arena too small (need > %zu bytes)", A.cap); return APRB_ERR_WORKSPACE; } \
    }
    if (b.type == 0) {                                             // SimpleBlock on the raw fp32 input features
        if (feat16) { set_error("aprb_kfe_forward: SimpleBlock after the first block is not supported in fp16 mode"); return APRB_ERR_UNSUPPORTED; }
        const int cout = b.out_dim / 2;
        KFE_ALLOC(y, float, (size_t)nq * cout);
        const size_t mark = A.off;
        KFE_ALLOC(t, float, (size_t)nq * cout);
        KFE_GSTAT(gt, nq, cout);
        KFE_OK(kpconv_call(h, b, q, s, idx, H, (const float*)feat, nq, ns, H, b.in_dim, cout, t, &gt, A, st));
        KFE_OK(norm16_call(h, lq, t, nq, cout, nullptr, 0, 0, y, out16, &gt, nullptr, A, st));
        A.off = mark;
        *out = y; *out_cols = cout;
        return APRB_OK;
    }
    if (!feat16) { set_error("aprb_kfe_forward: fp16 mode expects fp16 features at a resnet block"); return APRB_ERR_UNSUPPORTED; }
    const int mid = b.out_dim / 4, cout = b.out_dim;
    void* const* w16 = &h.w16[bi * 3];
    float* y = nullptr;                                            // fp32-sized: holds fp16 or (last block) fp32
    if (last && h.y_last) {                                        // async host output: dedicated ping-pong buffer
        if (nq > h.y_last_rows_cap) { set_error("aprb_kfe_forward: output has %d rows, buffer holds %d", nq, h.y_last_rows_cap); return APRB_ERR_WORKSPACE; }
        y = h.y_last;
    } else {
        y = A.take<float>((size_t)nq * cout);
        if (!y) { set_error("aprb_kfe_forward: arena too small (need > %zu bytes)", A.cap); return APRB_ERR_WORKSPACE; }
    }
    const size_t mark = A.off;
    const void* x1 = feat;
    if (b.unary1_W) {
        KFE_ALLOC(t1raw, float, (size_t)ns * mid);
        KFE_ALLOC(t1, float, (size_t)ns * mid / 2 + 8);
        KFE_GSTAT(g1, ns, mid);
        KFE_OK(aprb_linear_f16_stats(feat, w16[0], ns, b.in_dim, mid, t1raw, g1.buf, g1.buf ? &g1.written : nullptr, st));
        KFE_OK(norm16_call(h, l, t1raw, ns, mid, nullptr, 0, 0, t1, 1, &g1, nullptr, A, st));
        x1 = t1;
    }
    KFE_ALLOC(t2raw, float, (size_t)nq * mid);
    KFE_ALLOC(t2, float, (size_t)nq * mid / 2 + 8);
    KFE_GSTAT(g2, nq, mid);
    {
        size_t need = aprb_kpconv_ws_bytes(nq, ns, H, h.cfg.K, mid, mid);
        if (A.scratch_bytes() < need) { set_error("aprb_kfe_forward: arena too small for the KPConv workspace"); return APRB_ERR_WORKSPACE; }
        // weighting stage on tcgen05 (mode 5) when the block carries the ck-ordered operand and the shape allows
        const bool tc = b.kp_Wprep16ck && aprb_kpconv_tc_supported(H, h.cfg.K, mid, mid, ns);
        KFE_OK(aprb_kpconv_forward_stats(q, s, idx, 0, H, (const float*)x1, b.kp, b.kp_W,
                                         (const float*)(tc ? b.kp_Wprep16ck : b.kp_Wprep16), b.extent, nq, ns,
                                         H, h.cfg.K, mid, mid, t2raw, tc ? 5 : 4, g2.buf, g2.buf ? &g2.written : nullptr, A.scratch(),
                                         A.scratch_bytes(), st));
    }
    KFE_OK(tap_push(h, x1, ns, mid, 1, 4 * (int)bi + 2, st));
    KFE_OK(tap_push(h, t2raw, nq, mid, 0, 4 * (int)bi + 1, st));
    KFE_OK(norm16_call(h, lq, t2raw, nq, mid, nullptr, 0, 0, t2, 1, &g2, nullptr, A, st));
    KFE_ALLOC(t3, float, (size_t)nq * cout);
    KFE_GSTAT(g3, nq, cout);
    const void* sc = feat;
    if (b.strided) {
        KFE_ALLOC(mp, float, (size_t)nq * b.in_dim / 2 + 8);
        KFE_OK(aprb_max_pool_seg(feat, 1, idx, 0, H, nq, ns, H, b.in_dim, h.seg[lq], h.S, h.pool_w[l], mp, st));
        sc = mp;
    }
    // Measured per block shape (tools/nrm_bench.py, profiles/r01_nrm_recompute.txt): recomputing wins while the products
    // are HBM-bound (1.15-1.41x for K_main + K_shortcut <= 1024) and loses once the repeated contraction is the cost
    // (0.86x at 512 + 1024 -> 2048, the last shortcut block): those keep the stored path.
    const int k_total = mid + (b.shortcut_W ? b.in_dim : 0);
    if (aprb::g_gemm_apply && g3.buf && (k_total <= 1024 || aprb::g_gemm_apply > 1)) {
        // Recompute path (include/aprb200.h, aprb_linear_f16_norm_apply): the [nq, cout] fp32 outputs of unary2 and of the
        // shortcut Linear are never materialised — statistics pass, per-segment mean / rstd, then the contraction again
        // with the normalisation, the shortcut and the activation in its epilogue. t3 / t4 only receive the ragged rows.
        const int nt = b.shortcut_W ? 2 : 1;
        KFE_OK(aprb_linear_f16_stats_ragged(t2, w16[1], nq, mid, cout, t3, g3.buf, h.seg[lq], h.S, st));
        float* t4 = nullptr; GStat g4;
        if (b.shortcut_W) {
            t4 = A.take<float>((size_t)nq * cout);
            g4.buf = (float*)A.take<char>(aprb_group_stats_bytes(nq, cout));
            if (!t4 || !g4.buf) { set_error("aprb_kfe_forward: arena too small (need > %zu bytes)", A.cap); return APRB_ERR_WORKSPACE; }
            KFE_OK(aprb_linear_f16_stats_ragged(sc, w16[2], nq, b.in_dim, cout, t4, g4.buf, h.seg[lq], h.S, st));
        }
        KFE_ALLOC(stt, float, (size_t)h.S * nt * 2 * cout);
        KFE_OK(aprb_instnorm_seg_stats(t3, t4, nq, cout, h.seg[lq], h.S, 1e-5f, g3.buf, g4.buf, stt, st));
        KFE_OK(aprb_linear_f16_norm_apply(t2, w16[1], nq, mid, cout, b.shortcut_W ? sc : nullptr, b.shortcut_W ? w16[2] : nullptr,
                                          b.in_dim, b.shortcut_W ? nullptr : sc, h.seg[lq], h.S, stt, 0.1f, y, out16, st));
        A.off = mark;
        *out = y; *out_cols = cout;
        return APRB_OK;
    }
    KFE_OK(aprb_linear_f16_stats(t2, w16[1], nq, mid, cout, t3, g3.buf, g3.buf ? &g3.written : nullptr, st));
    if (b.shortcut_W) {
        KFE_ALLOC(t4, float, (size_t)nq * cout);
        KFE_GSTAT(g4, nq, cout);
        KFE_OK(aprb_linear_f16_stats(sc, w16[2], nq, b.in_dim, cout, t4, g4.buf, g4.buf ? &g4.written : nullptr, st));
        KFE_OK(norm16_call(h, lq, t3, nq, cout, t4, 0, 1, y, out16, &g3, &g4, A, st));
    } else {
        KFE_OK(norm16_call(h, lq, t3, nq, cout, sc, 1, 0, y, out16, &g3, nullptr, A, st));
    }
#undef KFE_GSTAT
    A.off = mark;
    *out = y; *out_cols = cout;
    return APRB_OK;
}

__global__ void fill_kernel(float* p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

extern "C" int aprb_kfe_create(const aprb_kfe_config* cfg, const aprb_kfe_block* blocks, int nblocks, aprb_kfe** out) {
    APRB_REQUIRE(cfg && blocks && out && nblocks >= 1, "null argument");
    APRB_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= KFE_MAX_LEVELS, "num_layers out of range");
    APRB_REQUIRE(cfg->K >= 1 && cfg->K <= 16, "K out of range");
    for (int i = 0; i < nblocks; ++i) {
        const aprb_kfe_block& b = blocks[i];
        APRB_REQUIRE(b.layer >= 0 && b.layer < cfg->num_layers && (!b.strided || b.layer + 1 < cfg->num_layers), "block layer out of range");
        APRB_REQUIRE(b.kp && b.kp_W, "block without KPConv parameters");
        if (b.type == 1) {
            APRB_REQUIRE(b.unary2_W, "resnet block without unary2 weights");
            const int mid = b.out_dim / 4;
            bool ok = mid % 32 == 0 && b.out_dim % 16 == 0 && b.in_dim % 32 == 0;
            if (!ok) { set_error("aprb_kfe_create: block %d dims (%d -> %d) not supported by the tcgen05 Linear", i, b.in_dim, b.out_dim); return APRB_ERR_UNSUPPORTED; }
        }
    }
    aprb_kfe* h = new (std::nothrow) aprb_kfe();
    APRB_REQUIRE(h, "out of host memory");
    h->cfg = *cfg;
    h->blocks.assign(blocks, blocks + nblocks);
    h->h_counts = nullptr; h->h_counts_cap = 0; h->B = 0; h->S = 1;
    h->out_dev[0] = h->out_dev[1] = nullptr; h->out_dev_floats = 0; h->copy_st = nullptr; h->ev_done = nullptr;
    h->ev_copied[0] = h->ev_copied[1] = nullptr; h->parity = 0; h->y_last = nullptr; h->y_last_rows_cap = 0;
    h->host_out_f16 = 0; h->y_last_f16 = 0;
    h->tap_buf = nullptr; h->tap_cap = 0; h->tap_off = 0;
    for (int l = 0; l < KFE_MAX_LEVELS; ++l) h->pool_w[l] = nullptr;
    for (int l = 0; l < KFE_MAX_LEVELS; ++l) {
        h->ev[l] = nullptr; h->n[l] = 0; h->pts[l] = nullptr; h->lens[l] = nullptr; h->conv[l] = h->pool[l] = h->up[l] = nullptr;
        if (cudaEventCreateWithFlags(&h->ev[l], cudaEventDisableTiming | (aprb::g_blocking_sync ? cudaEventBlockingSync : 0)) != cudaSuccess) { set_error("cudaEventCreate failed"); delete h; return APRB_ERR_CUDA; }
    }
    // fp16 activation mode: shape constraints of the fp16 kernels, and fp16 copies of the unary weights
    h->act16_ok = nblocks >= 2 && blocks[0].type == 0;
    for (int i = 1; i < nblocks && h->act16_ok; ++i) {
        const aprb_kfe_block& b = blocks[i];
        const int mid = b.out_dim / 4;
        const bool slab_ok = mid == 64 || mid == 128 || mid == 256 || mid % 512 == 0;     // producer slabs hold whole rows
        h->act16_ok = b.type == 1 && b.kp_Wprep16 && b.in_dim % 64 == 0 && mid % 64 == 0 && b.out_dim % 16 == 0 && slab_ok &&
                      (!b.strided || (b.in_dim % 128 == 0 && b.in_dim <= 1024)) && cfg->limits[b.layer] <= 128;
    }
    h->w16.assign((size_t)nblocks * 3, nullptr);
    if (h->act16_ok) {
        for (int i = 1; i < nblocks; ++i) {
            const aprb_kfe_block& b = blocks[i];
            const float* src[3] = {b.unary1_W, b.unary2_W, b.shortcut_W};
            const size_t cnt[3] = {(size_t)(b.out_dim / 4) * b.in_dim, (size_t)b.out_dim * (b.out_dim / 4), (size_t)b.out_dim * b.in_dim};
            for (int j = 0; j < 3; ++j) {
                if (!src[j]) continue;
                void* p = nullptr;
                if (cudaMalloc(&p, cnt[j] * 2) != cudaSuccess || aprb_f32_to_f16(src[j], p, cnt[j], nullptr) != APRB_OK) {
                    set_error("aprb_kfe_create: fp16 weight copy failed");
                    aprb_kfe_destroy(h);
                    return APRB_ERR_CUDA;
                }
                h->w16[(size_t)i * 3 + j] = p;
            }
        }
        if (cudaStreamSynchronize(nullptr) != cudaSuccess) { set_error("aprb_kfe_create: sync failed"); aprb_kfe_destroy(h); return APRB_ERR_CUDA; }
    }
    *out = h;
    return APRB_OK;
}

extern "C" void aprb_kfe_destroy(aprb_kfe* h) {
    if (!h) return;
    for (void* p : h->w16) if (p) cudaFree(p);
    for (int i = 0; i < 2; ++i) { if (h->out_dev[i]) cudaFree(h->out_dev[i]); if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); }
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->copy_st) cudaStreamDestroy(h->copy_st);
    for (int l = 0; l < KFE_MAX_LEVELS; ++l) if (h->ev[l]) cudaEventDestroy(h->ev[l]);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    delete h;
}

// Arena size when level l holds at most N * ratio^l points (ratio in (0, 1]; 1 = the unconditional bound: a grid
// subsampling never returns more points than it was given).
extern "C" size_t aprb_kfe_arena_bytes_est(const aprb_kfe* h, int N, int B, float ratio) {
    if (!h || N < 0 || B < 1) return 0;
    if (!(ratio > 0.f) || ratio > 1.f) ratio = 1.f;
    const int L = h->cfg.num_layers;
    size_t nl[KFE_MAX_LEVELS + 1];
    double f = 1.0;
    for (int l = 0; l <= KFE_MAX_LEVELS; ++l) {
        size_t v = (size_t)((double)(N > 0 ? N : 1) * f) + 256;     // + slack for tiny clouds
        nl[l] = v < (size_t)(N > 0 ? N : 1) ? v : (size_t)(N > 0 ? N : 1);
        f *= ratio;
    }
    size_t pyramid = 0;
    for (int l = 0; l < L; ++l) {
        const size_t lim = (size_t)h->cfg.limits[l], n = nl[l];
        pyramid += 3 * align256(n * lim * 4) + align256(n * 12) + aprb_cell_grid_bytes((int)n, B) + 4096;
    }
    size_t feats = 0, worst_tmp = 0;
    for (const aprb_kfe_block& b : h->blocks) {
        const size_t ns = nl[b.layer], nq = b.strided ? nl[b.layer + 1] : ns;          // support / query rows of the block
        const size_t lim = (size_t)h->cfg.limits[b.layer];
        size_t cout = b.type == 0 ? b.out_dim / 2 : b.out_dim, cin_k = b.type == 0 ? b.in_dim : b.out_dim / 4;
        feats += align256(nq * cout * 4);
        size_t tmp = ns * 4 * (2 * (size_t)(b.out_dim / 4) + b.in_dim) + nq * 4 * (2 * (size_t)(b.out_dim / 4) + 2 * cout + b.in_dim) + 16 * 256;
        tmp += tmp / 16 + 8 * 256;                                  // group statistics: 1/16 of each GEMM output
        size_t kpw = aprb_kpconv_ws_bytes((int)nq, (int)ns, (int)lim, h->cfg.K, (int)cin_k, (int)(b.type == 0 ? cout : cin_k));
        size_t lin = aprb_linear_tf32_ws_bytes((int)ns, b.in_dim, (int)cout);
        size_t scratch = kpw > lin ? kpw : lin;
        if (tmp + scratch > worst_tmp) worst_tmp = tmp + scratch;
    }
    size_t sub = aprb_grid_subsample_ws_bytes((int)nl[0], B, 0);
    const int cps = h->cfg.clouds_per_segment > 0 ? h->cfg.clouds_per_segment : B;
    size_t norm = aprb_instnorm_seg_ws_bytes((int)nl[0], 2048, cdiv(B, cps));
    return pyramid + feats + worst_tmp + sub + norm + align256(nl[0] * 4) + (1u << 20);
}

// Unconditional bound (every level as large as level 0): 27 GB for a super-batch of 8 KITTI pairs. Callers that can
// retry use aprb_kfe_arena_bytes_est(ratio ~ 0.6; measured level ratios are 0.40-0.42) and fall back to this on
// APRB_ERR_WORKSPACE — aprb_kfe_forward checks every carve and never writes past the arena.
extern "C" size_t aprb_kfe_arena_bytes(const aprb_kfe* h, int N, int B) { return aprb_kfe_arena_bytes_est(h, N, B, 1.0f); }

extern "C" int aprb_kfe_forward(aprb_kfe* hp, const float* d_pts, const int32_t* d_lens, const float* d_feats, int N, int B,
                                void* d_arena, size_t arena_bytes, const float** out_feats, int* out_rows, int* out_cols,
                                void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(hp && d_pts && d_lens && d_arena && out_feats && out_rows && out_cols, "null argument");
    APRB_REQUIRE(N >= 1 && B >= 1, "need N >= 1 and B >= 1");
    aprb_kfe& h = *hp;
    const aprb_kfe_config& cfg = h.cfg;
    const int L = cfg.num_layers;
    if (h.h_counts_cap < (2 + B) * KFE_MAX_LEVELS) {
        if (h.h_counts) cudaFreeHost(h.h_counts);
        h.h_counts_cap = (2 + B) * KFE_MAX_LEVELS;
        APRB_CUDA_OK(cudaMallocHost(&h.h_counts, sizeof(int) * h.h_counts_cap));
    }
    Arena A(d_arena, arena_bytes);
    h.B = B;
    h.n[0] = N; h.pts[0] = d_pts; h.lens[0] = d_lens;
    const int cps = cfg.clouds_per_segment > 0 ? cfg.clouds_per_segment : B;
    h.S = cdiv(B, cps);
    for (int l = 0; l < KFE_MAX_LEVELS; ++l) { h.seg[l] = nullptr; h.pool_w[l] = nullptr; }
    h.taps.clear(); h.tap_off = 0;
    h.block_out.clear();
    if (h.S > 1) {
        KFE_ALLOC(seg0, int, (size_t)h.S + 1);
        KFE_OK(aprb_segment_offsets(d_lens, B, cps, seg0, st));
        h.seg[0] = seg0;
    }
    for (int l = 0; l < KFE_MAX_LEVELS; ++l) { h.conv[l] = h.pool[l] = h.up[l] = nullptr; if (l) { h.n[l] = 0; h.pts[l] = nullptr; h.lens[l] = nullptr; } }

    const float* x = d_feats;
    int xc = cfg.in_feats_dim;
    const bool act16 = aprb::g_act_f16 && aprb::g_kpconv_f16 && h.act16_ok;
    bool x_is16 = false;
    if (!x) {                                                      // features = ones (datasets/kitti.py:598-599)
        KFE_ALLOC(ones, float, (size_t)N * xc);
        APRB_TIMED("fill_kernel", st, 1, (fill_kernel<<<cdiv((long long)N * xc, 256), 256, 0, st>>>(ones, (size_t)N * xc, 1.0f)));
        x = ones;
    }

    float r = cfg.first_subsampling_dl * cfg.conv_radius;          // dataloader.py:93
    void* grid[KFE_MAX_LEVELS] = {nullptr};
    size_t grid_bytes[KFE_MAX_LEVELS] = {0};
    size_t bi = 0;                                                  // next encoder block
    auto run_blocks = [&](int level, int strided) -> int {
        while (bi < h.blocks.size() && h.blocks[bi].layer == level && (h.blocks[bi].strided != 0) == (strided != 0)) {
            const float* y = nullptr; int yc = 0;
            if (act16) {
                const void* y16 = nullptr;
                const bool last = bi + 1 == h.blocks.size();
                KFE_OK(run_block16(h, bi, x, x_is16, last, &y16, &yc, A, st));
                y = (const float*)y16; x_is16 = !last;
            } else {
                KFE_OK(run_block(h, h.blocks[bi], x, &y, &yc, A, st));
            }
            x = y; xc = yc;
            {
                const aprb_kfe_block& ob = h.blocks[bi];
                h.block_out.push_back({y, ob.strided ? h.n[ob.layer + 1] : h.n[ob.layer], yc,
                                       (act16 && (x_is16 || (bi + 1 == h.blocks.size() && h.y_last_f16))) ? 1 : 0});
            }
            if (h.tap_buf) {                                        // debug tap: keep a copy of this block's output
                const aprb_kfe_block& tb = h.blocks[bi];
                const int rows = tb.strided ? h.n[tb.layer + 1] : h.n[tb.layer];
                const int f16 = (act16 && (x_is16 || (bi + 1 == h.blocks.size() && h.y_last_f16))) ? 1 : 0;
                KFE_OK(tap_push(h, y, rows, yc, f16, 4 * (int)bi + 0, st));
            }
            ++bi;
        }
        return APRB_OK;
    };

    for (int l = 0; l < L; ++l) {
        const int lim = cfg.limits[l];
        APRB_REQUIRE(lim >= 1, "neighbourhood limit must be >= 1");
        if (l == 0) {
            grid_bytes[0] = aprb_cell_grid_bytes(h.n[0], B);
            grid[0] = A.take<char>(grid_bytes[0]);
            if (!grid[0]) { set_error("aprb_kfe_forward: arena too small"); return APRB_ERR_WORKSPACE; }
            KFE_OK(aprb_cell_grid_build(h.pts[0], h.lens[0], B, h.n[0], r, grid[0], grid_bytes[0], st));
        }
        KFE_ALLOC(conv, int, (size_t)h.n[l] * lim);
        KFE_OK(aprb_cell_grid_query(grid[l], grid_bytes[l], h.pts[l], h.lens[l], B, h.n[l], h.n[l], r, lim, conv, lim, nullptr, nullptr, st));
        h.conv[l] = conv;
        float* npts = nullptr; int* nlens = nullptr; int* dcnt = nullptr;
        int* hc = h.h_counts + (2 + B) * l;
        if (l + 1 < L) {                                           // subsample now, read the count back later
            npts = A.take<float>((size_t)h.n[l] * 3); nlens = A.take<int>(B); dcnt = A.take<int>(2);
            if (!npts || !nlens || !dcnt) { set_error("aprb_kfe_forward: arena too small"); return APRB_ERR_WORKSPACE; }
            const float dl = 2 * r / cfg.conv_radius;              // dataloader.py:138
            if (A.scratch_bytes() < aprb_grid_subsample_ws_bytes(h.n[l], B, 0)) { set_error("aprb_kfe_forward: arena too small"); return APRB_ERR_WORKSPACE; }
            KFE_OK(aprb_grid_subsample_batch(h.pts[l], h.lens[l], B, h.n[l], dl, 0, nullptr, 0, npts, nlens, dcnt, nullptr, dcnt + 1, 32,
                                             A.scratch(), A.scratch_bytes(), st));
            APRB_CUDA_OK(cudaMemcpyAsync(hc, dcnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
            APRB_CUDA_OK(cudaMemcpyAsync(hc + 2, nlens, B * sizeof(int), cudaMemcpyDeviceToHost, st));
            APRB_CUDA_OK(cudaEventRecord(h.ev[l], st));
        }
        KFE_OK(run_blocks(l, 0));                                  // queue this level's non-strided blocks first
        if (l + 1 < L) {
            APRB_CUDA_OK(cudaEventSynchronize(h.ev[l]));           // only waits for the subsample, not for the blocks
            if (hc[1] == 2) {                                      // grid needs the 64-bit key: redo (rare, synchronous)
                const float dl = 2 * r / cfg.conv_radius;
                KFE_OK(aprb_grid_subsample_batch(h.pts[l], h.lens[l], B, h.n[l], dl, 0, nullptr, 0, npts, nlens, dcnt, nullptr, dcnt + 1, 64,
                                                 A.scratch(), A.scratch_bytes(), st));
                APRB_CUDA_OK(cudaMemcpyAsync(hc, dcnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
                APRB_CUDA_OK(cudaStreamSynchronize(st));
            }
            if (hc[1] != 0) { set_error("aprb_kfe_forward: voxel grid of level %d does not fit the sort key", l); return APRB_ERR_UNSUPPORTED; }
            if (hc[0] < 1) { set_error("aprb_kfe_forward: level %d is empty", l + 1); return APRB_ERR_EMPTY; }
            h.n[l + 1] = hc[0]; h.pts[l + 1] = npts; h.lens[l + 1] = nlens;
            if (h.S > 1) {
                KFE_ALLOC(segn, int, (size_t)h.S + 1);
                KFE_OK(aprb_segment_offsets(nlens, B, cps, segn, st));
                h.seg[l + 1] = segn;
            }
            KFE_ALLOC(pool, int, (size_t)h.n[l + 1] * lim);
            KFE_ALLOC(poolw, int, (size_t)h.S);
            KFE_OK(aprb_cell_grid_query_seg(grid[l], grid_bytes[l], npts, nlens, B, h.n[l + 1], h.n[l], r, lim, pool, lim, cps, poolw, st));
            h.pool[l] = pool; h.pool_w[l] = poolw;
            grid_bytes[l + 1] = aprb_cell_grid_bytes(h.n[l + 1], B);
            grid[l + 1] = A.take<char>(grid_bytes[l + 1]);
            if (!grid[l + 1]) { set_error("aprb_kfe_forward: arena too small"); return APRB_ERR_WORKSPACE; }
            KFE_OK(aprb_cell_grid_build(npts, nlens, B, h.n[l + 1], 2 * r, grid[l + 1], grid_bytes[l + 1], st));
            if (cfg.build_upsamples == 2) {                        // nearest support only: the column closest_pool reads
                KFE_ALLOC(up, int, (size_t)h.n[l]);
                KFE_OK(aprb_cell_grid_query_nearest(grid[l + 1], grid_bytes[l + 1], h.pts[l], h.lens[l], B, h.n[l], h.n[l + 1], 2 * r, up, 1, st));
                h.up[l] = up;
            } else if (cfg.build_upsamples) {
                KFE_ALLOC(up, int, (size_t)h.n[l] * lim);
                KFE_OK(aprb_cell_grid_query(grid[l + 1], grid_bytes[l + 1], h.pts[l], h.lens[l], B, h.n[l], h.n[l + 1], 2 * r, lim, up, lim, nullptr, nullptr, st));
                h.up[l] = up;
            }
            KFE_OK(run_blocks(l, 1));                              // the strided block that closes the level
        }
        r *= 2;
    }
    if (bi != h.blocks.size()) { set_error("aprb_kfe_forward: %zu block(s) not scheduled (check layer/strided order)", h.blocks.size() - bi); return APRB_ERR_INVALID; }
    *out_feats = x; *out_rows = h.n[h.blocks.back().strided ? h.blocks.back().layer + 1 : h.blocks.back().layer]; *out_cols = xc;
    return APRB_OK;
}

extern "C" int aprb_kfe_forward_host(aprb_kfe* h, const float* h_pts, const int32_t* h_lens, int N, int B, void* d_arena,
                                     size_t arena_bytes, float* h_out, int h_out_rows_cap, int* out_rows, int* out_cols,
                                     void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(h && h_pts && h_lens && d_arena && h_out && out_rows && out_cols, "null argument");
    APRB_REQUIRE(N >= 1 && B >= 1, "need N >= 1 and B >= 1");
    size_t head = align256((size_t)N * 12) + align256((size_t)B * 4);
    if (arena_bytes <= head) { set_error("aprb_kfe_forward_host: arena too small"); return APRB_ERR_WORKSPACE; }
    float* d_pts = (float*)d_arena;
    int* d_lens = (int*)((char*)d_arena + align256((size_t)N * 12));
    APRB_CUDA_OK(cudaMemcpyAsync(d_pts, h_pts, (size_t)N * 12, cudaMemcpyHostToDevice, st));
    APRB_CUDA_OK(cudaMemcpyAsync(d_lens, h_lens, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    const float* y = nullptr;
    KFE_OK(aprb_kfe_forward(h, d_pts, d_lens, nullptr, N, B, (char*)d_arena + head, arena_bytes - head, &y, out_rows, out_cols, st));
    if (*out_rows > h_out_rows_cap) { set_error("aprb_kfe_forward_host: output has %d rows, buffer holds %d", *out_rows, h_out_rows_cap); return APRB_ERR_WORKSPACE; }
    APRB_CUDA_OK(cudaMemcpyAsync(h_out, y, (size_t)(*out_rows) * (*out_cols) * sizeof(float), cudaMemcpyDeviceToHost, st));
    APRB_CUDA_OK(cudaStreamSynchronize(st));
    return APRB_OK;
}

// Asynchronous form: H2D and the path on `stream`, the final block writes into one of two device buffers owned by the
// handle, and a side stream copies that buffer to h_out once the path is done — so the caller can queue the next call
// (which uses the other buffer) while this call's output is still crossing PCIe. aprb_kfe_wait_host(h, ticket) blocks
// until the copy of the call that returned `ticket` has landed. At most two calls may be in flight per handle; the
// caller alternates two host buffers.
extern "C" int aprb_kfe_forward_host_async(aprb_kfe* h, const float* h_pts, const int32_t* h_lens, int N, int B, void* d_arena,
                                           size_t arena_bytes, float* h_out, int h_out_rows_cap, int* out_rows, int* out_cols,
                                           int* ticket, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(h && h_pts && h_lens && d_arena && h_out && out_rows && out_cols && ticket, "null argument");
    APRB_REQUIRE(N >= 1 && B >= 1 && h_out_rows_cap >= 1, "need N >= 1, B >= 1 and a non-empty output buffer");
    if (!h->copy_st) {
        APRB_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking));
        APRB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i)
            APRB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming | (aprb::g_blocking_sync ? cudaEventBlockingSync : 0)));
    }
    const int cols = h->blocks.back().type == 0 ? h->blocks.back().out_dim / 2 : h->blocks.back().out_dim;
    const size_t need = (size_t)h_out_rows_cap * cols;
    if (h->out_dev_floats < need) {                                  // grow-only; waits for everything in flight
        APRB_CUDA_OK(cudaDeviceSynchronize());
        for (int i = 0; i < 2; ++i) {
            if (h->out_dev[i]) cudaFree(h->out_dev[i]);
            h->out_dev[i] = nullptr;
            APRB_CUDA_OK(cudaMalloc(&h->out_dev[i], need * sizeof(float)));
        }
        h->out_dev_floats = need;
    }
    const int p = h->parity;
    // Zero-copy output (aprb_set_option("host_zero_copy"), off by default): when the caller's pinned buffer is addressable
    // from the device, the final normalisation kernel stores the encoder output straight into it (posted PCIe writes
    // from the kernel instead of a DMA read-back). Measured on B200: slower than the copy-engine path (1991 vs 2393
    // clouds/s end to end), which itself costs ~17 % of the device-resident rate while it overlaps the next call.
    float* d_alias = nullptr;
    if (aprb::g_host_zero_copy) {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, h_out, 0) == cudaSuccess) d_alias = (float*)dp;
        else (void)cudaGetLastError();
    }
    size_t head = align256((size_t)N * 12) + align256((size_t)B * 4);
    if (arena_bytes <= head) { set_error("aprb_kfe_forward_host_async: arena too small"); return APRB_ERR_WORKSPACE; }
    float* d_pts = (float*)d_arena;
    int* d_lens = (int*)((char*)d_arena + align256((size_t)N * 12));
    APRB_CUDA_OK(cudaMemcpyAsync(d_pts, h_pts, (size_t)N * 12, cudaMemcpyHostToDevice, st));
    APRB_CUDA_OK(cudaMemcpyAsync(d_lens, h_lens, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    APRB_CUDA_OK(cudaStreamWaitEvent(st, h->ev_copied[p], 0));     // the copy that last read out_dev[p] (two calls ago)
    const float* y = nullptr;
    const bool act16 = aprb::g_act_f16 && aprb::g_kpconv_f16 && h->act16_ok;
    if (h->host_out_f16 && !act16) { set_error("aprb_kfe_forward_host_async: fp16 host output needs the fp16 activation mode"); return APRB_ERR_UNSUPPORTED; }
    h->y_last = d_alias ? d_alias : h->out_dev[p]; h->y_last_rows_cap = h_out_rows_cap; h->y_last_f16 = h->host_out_f16;
    int rc = aprb_kfe_forward(h, d_pts, d_lens, nullptr, N, B, (char*)d_arena + head, arena_bytes - head, &y, out_rows, out_cols, st);
    h->y_last = nullptr; h->y_last_rows_cap = 0; h->y_last_f16 = 0;
    if (rc) return rc;
    const size_t elt = h->host_out_f16 ? 2 : sizeof(float);
    if (d_alias) {                                                   // the output is complete when the stream gets here
        APRB_CUDA_OK(cudaEventRecord(h->ev_copied[p], st));
        *ticket = p;
        h->parity = p ^ 1;
        return APRB_OK;
    }
    APRB_CUDA_OK(cudaEventRecord(h->ev_done, st));
    APRB_CUDA_OK(cudaStreamWaitEvent(h->copy_st, h->ev_done, 0));
    if (!aprb::g_dbg_skip_d2h)   // diagnostic only (aprb_set_option("dbg_skip_d2h")): how much of e2e is the output copy
        APRB_CUDA_OK(cudaMemcpyAsync(h_out, y, (size_t)(*out_rows) * (*out_cols) * elt, cudaMemcpyDeviceToHost, h->copy_st));
    APRB_CUDA_OK(cudaEventRecord(h->ev_copied[p], h->copy_st));
    *ticket = p;
    h->parity = p ^ 1;
    return APRB_OK;
}

// The encoder output of aprb_kfe_forward_host_async as fp16: the final activation is already rounded to a 10-bit
// mantissa, so for |v| >= 2^-14 the fp16 value converts back to the identical fp32 value (below that: fp16 subnormals,
// absolute error <= 2^-25) and the device-to-host copy moves half the bytes. Needs the fp16 activation mode.
extern "C" int aprb_kfe_set_host_output_f16(aprb_kfe* h, int on) {
    APRB_REQUIRE(h, "null handle");
    h->host_out_f16 = on ? 1 : 0;
    return APRB_OK;
}

extern "C" int aprb_kfe_wait_host(aprb_kfe* h, int ticket) {
    APRB_REQUIRE(h && (ticket == 0 || ticket == 1), "bad argument");
    if (!h->ev_copied[ticket]) return APRB_OK;
    APRB_CUDA_OK(cudaEventSynchronize(h->ev_copied[ticket]));
    return APRB_OK;
}

extern "C" int aprb_kfe_set_tap(aprb_kfe* h, void* d_buf, size_t bytes) {
    APRB_REQUIRE(h, "null handle");
    h->tap_buf = (char*)d_buf; h->tap_cap = d_buf ? bytes : 0; h->tap_off = 0; h->taps.clear();
    return APRB_OK;
}

extern "C" int aprb_kfe_get_block_output(const aprb_kfe* h, int block, const void** d_ptr, int* rows, int* cols, int* is_f16) {
    APRB_REQUIRE(h && d_ptr && rows && cols && is_f16, "null argument");
    APRB_REQUIRE(block >= 0 && (size_t)block < h->block_out.size(), "block index out of range (run a forward first)");
    const aprb_kfe::BlockOut& o = h->block_out[(size_t)block];
    *d_ptr = o.p; *rows = o.rows; *cols = o.cols; *is_f16 = o.f16;
    return APRB_OK;
}

extern "C" int aprb_kfe_tap_count(const aprb_kfe* h) { return h ? (int)h->taps.size() : 0; }

extern "C" int aprb_kfe_get_tap(const aprb_kfe* h, int i, const void** d_ptr, int* rows, int* cols, int* is_f16, int* tag) {
    APRB_REQUIRE(h && d_ptr && rows && cols && is_f16 && tag, "null argument");
    APRB_REQUIRE(i >= 0 && (size_t)i < h->taps.size(), "tap index out of range");
    const aprb_kfe::Tap& t = h->taps[(size_t)i];
    *d_ptr = h->tap_buf + t.off; *rows = t.rows; *cols = t.cols; *is_f16 = t.f16; *tag = t.tag;
    return APRB_OK;
}

extern "C" int aprb_kfe_get(const aprb_kfe* h, int what, int level, const void** d_ptr, int* rows, int* cols) {
    APRB_REQUIRE(h && d_ptr && rows && cols, "null argument");
    APRB_REQUIRE(level >= 0 && level < h->cfg.num_layers, "level out of range");
    const int lim = h->cfg.limits[level];
    switch (what) {
        case 0: *d_ptr = h->pts[level]; *rows = h->n[level]; *cols = 3; break;
        case 1: *d_ptr = h->conv[level]; *rows = h->conv[level] ? h->n[level] : 0; *cols = lim; break;
        case 2: *d_ptr = h->pool[level]; *rows = h->pool[level] ? h->n[level + 1] : 0; *cols = lim; break;
        case 3: *d_ptr = h->up[level]; *rows = h->up[level] ? h->n[level] : 0; *cols = h->cfg.build_upsamples == 2 ? 1 : lim; break;
        case 4: *d_ptr = h->lens[level]; *rows = h->B; *cols = 1; break;
        default: set_error("aprb_kfe_get: unknown selector %d", what); return APRB_ERR_INVALID;
    }
    return APRB_OK;
}
