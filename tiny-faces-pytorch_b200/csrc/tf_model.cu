// DetectionModel forward / backward executor: the whole ResNet-101 res2-res4 trunk + heads of
// /root/reference/tinyfaces/models/model.py:89-128 (and the backward autograd derives for it,
// /root/reference/tinyfaces/trainer.py:86) as one stream-ordered sequence of this library's kernels.
//
// Data layout: activations NHWC fp32 in a caller-provided workspace (bump-allocated; every tensor needed by the
// backward stays resident -- ~25 GB at batch-8 960x1280, far below the 180 GB of HBM3e); weights are re-packed
// from the caller's OIHW parameters into [Cout][tap][Cin] each call.  The caller's tensors (image, parameters,
// output, gradients) keep the reference's NCHW / OIHW layouts.
//
// Precision modes: 1 = "fast": GEMM operands rounded to TF32 by their producers, one tensor-core product;
//                  2 = "parity": operands kept as exact (hi, lo) TF32 splits, three products (3xTF32).
// Stride-2 convolutions: fprop / wgrad read the full-resolution input through a TMA traversal stride; the dgrad runs as
// 4 parity-class GEMMs over dY (tfg::conv_dgrad_s2), no zero insertion.
#include "tf_common.cuh"
#include "tf_conv_gemm.h"
#include "tf_elementwise.h"
#include <algorithm>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace {

#define RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

struct ConvP { int w = -1; int cin = 0, cout = 0, k = 1, stride = 1; };
struct BnP { int gamma = -1, beta = -1, rm = -1, rv = -1, C = 0; };
struct BlockP { ConvP c1, c2, c3, cd; BnP b1, b2, b3, bd; bool has_ds = false; int stride = 1; };

struct Arena {
    char* base = nullptr; size_t cap = 0, off = 0, peak = 0; bool dry = true;
    float* f(size_t n) {
        off = tf_align_up(off, 1024);
        float* p = reinterpret_cast<float*>(base + off);
        off += n * sizeof(float);
        if (off > peak) peak = off;
        return p;
    }
};

// one conv + BN unit as the backward needs it
struct Unit {
    ConvP c; BnP bn;
    const float* x = nullptr; const float* x_lo = nullptr;      // GEMM input (after subsampling for 1x1/s2)
    int B = 0, H = 0, W = 0;                                    // spatial size the GEMM ran at
    int Ho = 0, Wo = 0;                                         // output size (after subsampling for 3x3/s2)
    float* y = nullptr;                                         // raw conv output at (Ho, Wo)
    float *mean = nullptr, *rstd = nullptr, *scale = nullptr, *shift = nullptr;
    float* a = nullptr; float* a_lo = nullptr;                  // post BN(+ReLU) activation (null for bn3 / ds)
    unsigned int* amask = nullptr;                              // 1-bit ReLU mask of a (training)
    float* wp = nullptr; float* wp_lo = nullptr;                // packed fprop weights
};
struct BlockS { Unit u1, u2, u3, ud; const float* x = nullptr; const float* x_lo = nullptr; int B = 0, H = 0, W = 0, Ho = 0, Wo = 0;
                float* out = nullptr; float* out_lo = nullptr; unsigned int* omask = nullptr; bool has_ds = false; };

struct Model {
    int T = 25, Cn = 125, Cp = 128;
    std::vector<std::string> names;
    ConvP stem; BnP stem_bn; std::vector<BlockP> blocks;
    int s3_w = -1, s3_b = -1, s4_w = -1, s4_b = -1, up_w = -1;
    // ---- state of the last forward (training) ----
    int B = 0, H = 0, W = 0, mode = 1, training = 0;
    Arena ar;
    std::vector<const void*> params;     // copied from the caller's table at every forward (backward reuses it)
    Unit stem_u; float* col = nullptr; float* col_lo = nullptr; int H2 = 0, W2 = 0, Hp = 0, Wp = 0;
    float* pool = nullptr; float* pool_lo = nullptr; unsigned char* pool_argmax = nullptr;
    std::vector<BlockS> bs;
    int H3 = 0, W3 = 0, H4 = 0, W4 = 0;
    const float* res3 = nullptr; const float* res3_lo = nullptr; const float* res4 = nullptr; const float* res4_lo = nullptr;
    float *w3p = nullptr, *w3p_lo = nullptr, *w4p = nullptr, *w4p_lo = nullptr, *b3p = nullptr, *b4p = nullptr, *up = nullptr, *offdiag = nullptr;
    float* partial = nullptr; float* bwd_slots = nullptr; float* coef = nullptr; float* dwtmp = nullptr; size_t fwd_mark = 0;
    float eps = 1e-5f, momentum = 0.1f;
    // weight-gradient GEMMs run on a side stream: they depend only on (x, dy) and nothing in the backward chain
    // depends on them, so their tensor-core time overlaps the HBM-bound BN-backward kernels of the main stream
    cudaStream_t side = nullptr;
    // The backward chain itself runs on an internal HIGH-priority stream (forked from / joined to the caller's stream):
    // when a dgrad GEMM of the chain and a weight-gradient GEMM of the side stream are both ready, the block scheduler
    // hands the SMs to the chain first (both kernels need a whole SM's shared memory, so they cannot co-reside).
    cudaStream_t chain = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pack = nullptr, ev_pack_dgrad = nullptr, ev_region[2] = {nullptr, nullptr}, ev_chain[2] = {nullptr, nullptr};
    // wgrad_mode 0: weight gradient forked before its dgrad is enqueued; 1: forked before, enqueued after the dgrad;
    //            2: deferred -- a block's three weight gradients are enqueued when the NEXT block's bn3 backward
    //               (the longest HBM-bound stretch of the chain, ~6 passes over a 1024-channel tensor) starts
    int wgrad_mode = 2, chain_priority = 1;
    struct PendingWgrad { tfg::WgradArgs w; size_t dw_elems; int O, I, taps, I_pad; float* gw; };
    std::vector<PendingWgrad> pending;
    // tf_model_backward_ex: event k is recorded (on the weight-gradient stream) once every gradient of the blocks with index
    // > bucket_blocks[k] is enqueued -- the caller's collective for that bucket waits on it and overlaps the rest of the backward
    std::vector<void*> bucket_events; std::vector<int> bucket_blocks;
    size_t bwd_region_bytes = 0;
    int plan_B = 0, plan_H = 0, plan_W = 0, plan_training = -1, plan_mode = 0, plan_epoch = -1; size_t plan_need = 0;   // cached dry run
    bool side_enabled = true;
    int side_dev = -1;                    // device the internal streams / events were created on
    void destroy_side() {
        if (!side) return;
        cudaStreamDestroy(side); cudaStreamDestroy(chain); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); cudaEventDestroy(ev_pack); cudaEventDestroy(ev_pack_dgrad);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(ev_region[i]); cudaEventDestroy(ev_chain[i]); }
        side = nullptr; chain = nullptr;
    }
    int ensure_side() {
        if (!side_enabled) return TF_OK;
        int dev = 0;
        TF_CHECK_CUDA(cudaGetDevice(&dev));
        if (side && side_dev != dev) destroy_side();          // the model moved to another GPU
        if (side) return TF_OK;
        side_dev = dev;
        TF_CHECK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        TF_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        TF_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        TF_CHECK_CUDA(cudaEventCreateWithFlags(&ev_pack, cudaEventDisableTiming));
        TF_CHECK_CUDA(cudaEventCreateWithFlags(&ev_pack_dgrad, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) TF_CHECK_CUDA(cudaEventCreateWithFlags(&ev_region[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) TF_CHECK_CUDA(cudaEventCreateWithFlags(&ev_chain[i], cudaEventDisableTiming));
        int least = 0, greatest = 0;
        TF_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        TF_CHECK_CUDA(cudaStreamCreateWithPriority(&chain, cudaStreamNonBlocking, greatest));
        return TF_OK;
    }
    ~Model() {
        destroy_side();
        for (int i = 0; i < TABLE_SLOTS; ++i) if (table_dev[i]) cudaFree(table_dev[i]);
    }
    // dW (OIHW) = unpack(wgrad(x, dy)) on stream s2 (the side stream, or the chain itself when there is none).
    // A 1x1 convolution's packed gradient [Cout][1][Cin] IS the OIHW tensor: it is accumulated straight into the
    // caller's gradient, no staging buffer / unpack pass.
    int wgrad_run(const PendingWgrad& j, cudaStream_t s2) {
        tfg::WgradArgs w = j.w;
        if (j.taps == 1 && j.I_pad == j.I && j.O == w.Cout) {
            TF_CHECK_CUDA(cudaMemsetAsync(j.gw, 0, j.dw_elems * 4, s2));
            w.dw = j.gw;
            return tfg::conv_wgrad(w, s2);
        }
        TF_CHECK_CUDA(cudaMemsetAsync(dwtmp, 0, j.dw_elems * 4, s2));
        w.dw = dwtmp;
        RC(tfg::conv_wgrad(w, s2));
        RC(tfe::unpack_wgrad(dwtmp, j.O, j.I, j.taps, j.I_pad, j.gw, s2));
        return TF_OK;
    }
    // Enqueue behind everything already on `st`.  fork_recorded: ev_fork was already recorded on st at the point the
    // inputs became final (wgrad_mode 1 records it before the dgrad so the side stream does not wait for that GEMM).
    int wgrad_async(tfg::WgradArgs w, size_t dw_elems, int O, int I, int taps, int I_pad, float* gw, cudaStream_t st,
                    bool fork_recorded = false) {
        PendingWgrad j = {w, dw_elems, O, I, taps, I_pad, gw};
        if (!side) return wgrad_run(j, st);
        if (wgrad_mode == 2) { pending.push_back(j); return TF_OK; }
        if (!fork_recorded) TF_CHECK_CUDA(cudaEventRecord(ev_fork, st));
        TF_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
        return wgrad_run(j, side);
    }
    // wgrad_mode 2: everything queued so far depends only on work already enqueued on st
    int wgrad_flush(cudaStream_t st) {
        if (pending.empty() || !side) return TF_OK;
        TF_CHECK_CUDA(cudaEventRecord(ev_fork, st));
        TF_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
        for (const PendingWgrad& j : pending) RC(wgrad_run(j, side));
        pending.clear();
        return TF_OK;
    }

    int add(const std::string& n) { names.push_back(n); return (int)names.size() - 1; }
    ConvP conv(const std::string& p, int cin, int cout, int k, int stride) { ConvP c; c.w = add(p + ".weight"); c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; return c; }
    BnP bnp(const std::string& p, int C) { BnP b; b.gamma = add(p + ".weight"); b.beta = add(p + ".bias"); b.rm = add(p + ".running_mean"); b.rv = add(p + ".running_var"); b.C = C; return b; }

    explicit Model(int templates) : T(templates), Cn(5 * templates) {
        Cp = 64; while (Cp < Cn) Cp *= 2;                        // padded head width (power of two: column reductions)
        stem = conv("model.conv1", 3, 64, 7, 2);
        stem_bn = bnp("model.bn1", 64);
        const int nblocks[3] = {3, 4, 23}, planes[3] = {64, 128, 256}, strides[3] = {1, 2, 2};
        int inpl = 64;
        for (int l = 0; l < 3; ++l)
            for (int i = 0; i < nblocks[l]; ++i) {
                const std::string p = "model.layer" + std::to_string(l + 1) + "." + std::to_string(i);
                BlockP b;
                b.stride = i == 0 ? strides[l] : 1;
                b.c1 = conv(p + ".conv1", inpl, planes[l], 1, 1); b.b1 = bnp(p + ".bn1", planes[l]);
                b.c2 = conv(p + ".conv2", planes[l], planes[l], 3, b.stride); b.b2 = bnp(p + ".bn2", planes[l]);
                b.c3 = conv(p + ".conv3", planes[l], planes[l] * 4, 1, 1); b.b3 = bnp(p + ".bn3", planes[l] * 4);
                if (i == 0) { b.has_ds = true; b.cd = conv(p + ".downsample.0", inpl, planes[l] * 4, 1, b.stride); b.bd = bnp(p + ".downsample.1", planes[l] * 4); }
                inpl = planes[l] * 4;
                blocks.push_back(b);
            }
        s3_w = add("score_res3.weight"); s3_b = add("score_res3.bias");
        s4_w = add("score_res4.weight"); s4_b = add("score_res4.bias");
        up_w = add("score4_upsample.weight");
    }
    const float* P(int i) const { return reinterpret_cast<const float*>(params[i]); }
    float* PW(int i) const { return reinterpret_cast<float*>(const_cast<void*>(params[i])); }

    // ------------------------------------------------------------------------------------------ forward pieces
    // Weight packing is batched: prepack_begin/prepack_add/prepack_flush re-layout every weight of a pass in ONE
    // launch; pack() then only looks the result up (key = parameter index, transposed?).
    std::map<std::pair<int, int>, std::pair<float*, float*>> packed;
    std::vector<tfe::PackJob> jobs;
    long long jobs_total = 0;
    std::map<int, std::pair<float*, float*>> eval_ss;     // BN gamma index -> (scale, shift) of the eval-mode affine form
    bool fold_scale = false;             // eval fast mode: the BN scale of a fused conv+BN(+ReLU) is folded into its packed weights
    const float* eval_scale_of(const BnP& b) const {
        if (!fold_scale) return nullptr;
        auto it = eval_ss.find(b.gamma);
        return it == eval_ss.end() ? nullptr : it->second.first;
    }
    void prepack_add(const ConvP& c, int O_pad, int I_pad, int transpose, const float* oscale = nullptr) {
        const int taps = c.k * c.k;
        const size_t n = (size_t)O_pad * taps * I_pad;
        float* wp = ar.f(n);
        float* wp_lo = mode == 2 ? ar.f(n) : nullptr;
        packed[{c.w, transpose}] = {wp, wp_lo};
        tfe::PackJob j;
        memset(&j, 0, sizeof(j));                               // (the table is compared bytewise: no stale padding)
        j.src = ar.dry ? nullptr : P(c.w); j.dst = wp; j.dst_lo = wp_lo; j.oscale = oscale;
        j.O_src = c.cout; j.I_src = c.cin; j.taps = taps; j.transpose = transpose; j.O_pad = O_pad; j.I_pad = I_pad;
        j.begin = jobs_total;
        jobs_total += (long long)n;
        jobs.push_back(j);
    }
    // Job tables (weight packing, eval-mode BN) live in PERSISTENT device slots with a host mirror: a forward whose table
    // equals the cached one -- every step after the first for a given shape / parameter set -- uploads nothing.  That
    // removes three small H2D copies per step and makes the call graph-capturable (a copy from pageable host memory
    // cannot be captured); a table that differs DURING capture is an error: run one eager step first.
    static constexpr int TABLE_SLOTS = 4, TABLE_BYTES = 24576;
    char* table_dev[TABLE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<char> table_host[TABLE_SLOTS];
    int table_dev_id = -1;
    int table_slot = 0;
    int upload_table(const void* src, size_t bytes, cudaStream_t st, const void** dev_out) {
        TF_REQUIRE(table_slot < TABLE_SLOTS && bytes <= (size_t)TABLE_BYTES, "job table too large (%zu bytes, slot %d)", bytes, table_slot);
        int dev = 0;
        TF_CHECK_CUDA(cudaGetDevice(&dev));
        if (table_dev_id != dev) {
            for (int i = 0; i < TABLE_SLOTS; ++i) { if (table_dev[i]) cudaFree(table_dev[i]); table_dev[i] = nullptr; table_host[i].clear(); }
            table_dev_id = dev;
        }
        const int k = table_slot++;
        if (!table_dev[k]) TF_CHECK_CUDA(cudaMalloc(&table_dev[k], TABLE_BYTES));
        std::vector<char>& h = table_host[k];
        if (h.size() != bytes || memcmp(h.data(), src, bytes) != 0) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            TF_CHECK_CUDA(cudaStreamIsCapturing(st, &cs));
            TF_REQUIRE(cs == cudaStreamCaptureStatusNone, "tf_model: the job tables changed during stream capture -- run one eager call with the same shape / parameters first");
            // Stream-ordered overwrite: every earlier reader of this slot ran on `st` or on a stream `st` has since joined
            // (forward: ev_pack), and the side stream is forked from `st` before it uploads.  The source is pageable, so
            // the runtime stages it before returning -- `src` may be reused immediately.
            TF_CHECK_CUDA(cudaMemcpyAsync(table_dev[k], src, bytes, cudaMemcpyHostToDevice, st));
            h.assign(reinterpret_cast<const char*>(src), reinterpret_cast<const char*>(src) + bytes);
        }
        *dev_out = table_dev[k];
        return TF_OK;
    }
    int prepack_flush(cudaStream_t st) {
        const size_t bytes = jobs.size() * sizeof(tfe::PackJob);
        if (!ar.dry && !jobs.empty()) {
            const void* dev = nullptr;
            RC(upload_table(jobs.data(), bytes, st, &dev));
            RC(tfe::pack_weights_batched(reinterpret_cast<const tfe::PackJob*>(dev), (int)jobs.size(), jobs_total, mode, st));
        }
        jobs.clear(); jobs_total = 0;
        return TF_OK;
    }
    int pack(const ConvP& c, int O_pad, int I_pad, int transpose, float** wp, float** wp_lo, cudaStream_t st) {
        auto it = packed.find({c.w, transpose});
        if (it != packed.end()) { *wp = it->second.first; *wp_lo = it->second.second; return TF_OK; }
        const int taps = c.k * c.k;
        const size_t n = (size_t)O_pad * taps * I_pad;
        *wp = ar.f(n);
        *wp_lo = mode == 2 ? ar.f(n) : nullptr;
        if (!ar.dry) RC(tfe::pack_weight(P(c.w), c.cout, c.cin, taps, transpose, O_pad, I_pad, *wp, *wp_lo, mode, st));
        return TF_OK;
    }
    // conv (+ stride handling) producing the raw output u.y at (Ho, Wo); fused != 0: eval epilogue scale/shift(/relu)
    int run_conv(Unit& u, const float* x, const float* x_lo, int B_, int H_, int W_, int fused, int relu, float* fused_out,
                 cudaStream_t st, int* stats_blocks = nullptr, const float* res = nullptr) {
        const ConvP& c = u.c;
        u.B = B_;
        const int Ho = c.stride == 2 ? (H_ + 1) / 2 : H_, Wo = c.stride == 2 ? (W_ + 1) / 2 : W_;
        u.Ho = Ho; u.Wo = Wo;
        RC(pack(c, c.cout, c.cin, 0, &u.wp, &u.wp_lo, st));
        tfg::ConvArgs a = {};
        a.B = B_; a.Cin = c.cin; a.Cout = c.cout; a.ksize = c.k; a.w = u.wp; a.w_lo = u.wp_lo;
        if (fused) { a.scale = fold_scale ? nullptr : u.scale; a.shift = u.shift; a.relu = relu; a.round_out = 1; a.res = res; }
        // stride 2 is a TMA traversal stride: the GEMM reads the full-resolution input directly
        u.x = x; u.x_lo = x_lo; u.H = H_; u.W = W_;
        a.x = u.x; a.x_lo = u.x_lo; a.H = u.H; a.W = u.W; a.stride = c.stride;
        float* dst = fused_out ? fused_out : ar.f((size_t)B_ * Ho * Wo * c.cout);
        a.y = dst;
        if (stats_blocks) { a.stats_partial = partial; a.stats_blocks = stats_blocks; }   // BN statistics in the epilogue
        if (!ar.dry) RC(tfg::conv_fprop(a, st));
        u.y = dst;
        return TF_OK;
    }
    // conv whose training-mode BN statistics come out of the GEMM epilogue (separate reduction only for 3x3/s2)
    int conv_with_stats(Unit& u, const float* x, const float* x_lo, int B_, int H_, int W_, cudaStream_t st) {
        if (!training) return run_conv(u, x, x_lo, B_, H_, W_, 0, 0, nullptr, st);
        int nblk = 0;
        RC(run_conv(u, x, x_lo, B_, H_, W_, 0, 0, nullptr, st, &nblk));
        RC(alloc_bn(u));
        const long long M = (long long)u.B * u.Ho * u.Wo;
        if (!ar.dry) RC(tfe::bn_finalize_train(partial, nblk, M, u.bn.C, P(u.bn.gamma), P(u.bn.beta), eps, momentum, PW(u.bn.rm),
                                               PW(u.bn.rv), u.scale, u.shift, u.mean, u.rstd, st));
        return TF_OK;
    }
    int alloc_bn(Unit& u) {
        const int C = u.bn.C;
        u.scale = ar.f(C); u.shift = ar.f(C); u.mean = ar.f(C); u.rstd = ar.f(C);
        return TF_OK;
    }
    // eval-mode scale/shift of every BN are computed by ONE launch at the start of the forward (prepare_eval_all)
    int prepare_eval_all(cudaStream_t st) {
        eval_ss.clear();
        std::vector<tfe::BnEvalJob> ej;
        int total = 0;
        auto add = [&](const BnP& b) {
            float* sc = ar.f(b.C); float* sh = ar.f(b.C);
            eval_ss[b.gamma] = {sc, sh};
            tfe::BnEvalJob j;
            memset(&j, 0, sizeof(j));
            j.gamma = ar.dry ? nullptr : P(b.gamma); j.beta = ar.dry ? nullptr : P(b.beta);
            j.run_mean = ar.dry ? nullptr : P(b.rm); j.run_var = ar.dry ? nullptr : P(b.rv);
            j.scale = sc; j.shift = sh; j.begin = total;
            total += b.C;
            ej.push_back(j);
        };
        add(stem_bn);
        for (const BlockP& bp : blocks) { add(bp.b1); add(bp.b2); add(bp.b3); if (bp.has_ds) add(bp.bd); }
        const size_t bytes = ej.size() * sizeof(tfe::BnEvalJob);
        if (!ar.dry) {
            const void* dev = nullptr;
            RC(upload_table(ej.data(), bytes, st, &dev));
            RC(tfe::bn_scale_shift_eval_batched(reinterpret_cast<const tfe::BnEvalJob*>(dev), (int)ej.size(), total, eps, st));
        }
        return TF_OK;
    }
    int bn_prepare_eval(Unit& u, cudaStream_t st) {
        auto it = eval_ss.find(u.bn.gamma);
        if (it != eval_ss.end()) { u.scale = it->second.first; u.shift = it->second.second; u.mean = nullptr; u.rstd = nullptr; return TF_OK; }
        RC(alloc_bn(u));
        if (!ar.dry) RC(tfe::bn_scale_shift_eval(u.bn.C, P(u.bn.gamma), P(u.bn.beta), P(u.bn.rm), P(u.bn.rv), eps, u.scale, u.shift, st));
        return TF_OK;
    }
    // conv + BN + ReLU -> activation (u.a)
    int conv_bn_relu(Unit& u, const float* x, const float* x_lo, int B_, int H_, int W_, cudaStream_t st) {
        const int Ho = u.c.stride == 2 ? (H_ + 1) / 2 : H_, Wo = u.c.stride == 2 ? (W_ + 1) / 2 : W_;
        const long long M = (long long)B_ * Ho * Wo;
        if (!training && mode == 1) {           // eval fast path: BN + ReLU + TF32 rounding in the GEMM epilogue
            RC(bn_prepare_eval(u, st));
            u.a = ar.f((size_t)M * u.c.cout); u.a_lo = nullptr;
            RC(run_conv(u, x, x_lo, B_, H_, W_, 1, 1, u.a, st));
            return TF_OK;
        }
        if (!training) RC(bn_prepare_eval(u, st));
        RC(conv_with_stats(u, x, x_lo, B_, H_, W_, st));
        u.a = ar.f((size_t)M * u.c.cout);
        u.a_lo = mode == 2 ? ar.f((size_t)M * u.c.cout) : nullptr;
        u.amask = training ? reinterpret_cast<unsigned int*>(ar.f((size_t)(M * u.c.cout + 31) / 32)) : nullptr;
        if (!ar.dry) RC(tfe::bn_apply(u.y, u.scale, u.shift, nullptr, nullptr, nullptr, 1, M, u.c.cout, u.a, u.a_lo, mode, u.amask, st));
        return TF_OK;
    }
    int block_forward(const BlockP& bp, BlockS& s, const float* x, const float* x_lo, int B_, int H_, int W_, float* out_pre,
                      float* out_lo_pre, cudaStream_t st) {
        s.has_ds = bp.has_ds;
        s.x = x; s.x_lo = x_lo; s.B = B_; s.H = H_; s.W = W_;
        s.u1.c = bp.c1; s.u1.bn = bp.b1; s.u2.c = bp.c2; s.u2.bn = bp.b2; s.u3.c = bp.c3; s.u3.bn = bp.b3; s.ud.c = bp.cd; s.ud.bn = bp.bd;
        RC(conv_bn_relu(s.u1, x, x_lo, B_, H_, W_, st));
        RC(conv_bn_relu(s.u2, s.u1.a, s.u1.a_lo, B_, H_, W_, st));
        const int Ho = s.u2.Ho, Wo = s.u2.Wo;
        s.Ho = Ho; s.Wo = Wo;
        const long long Mo = (long long)B_ * Ho * Wo;
        const int C4 = bp.c3.cout;
        if (!training && mode == 1 && !tfg::debug_flag(12)) {
            // inference fast path: conv3 + folded BN + shortcut + ReLU in ONE kernel (residual epilogue), no BN-apply pass
            const float* res = x;
            if (bp.has_ds) {
                RC(bn_prepare_eval(s.ud, st));
                RC(run_conv(s.ud, x, x_lo, B_, H_, W_, 1, 0, nullptr, st));            // BN folded into the epilogue
                res = s.ud.y;
            }
            RC(bn_prepare_eval(s.u3, st));
            s.out = out_pre ? out_pre : ar.f((size_t)Mo * C4);
            s.out_lo = nullptr; s.omask = nullptr;
            RC(run_conv(s.u3, s.u2.a, s.u2.a_lo, B_, Ho, Wo, 1, 1, s.out, st, nullptr, res));
            return TF_OK;
        }
        if (!training) RC(bn_prepare_eval(s.u3, st));
        RC(conv_with_stats(s.u3, s.u2.a, s.u2.a_lo, B_, Ho, Wo, st));
        const float* res = x; const float* rscale = nullptr; const float* rshift = nullptr;
        if (bp.has_ds) {
            if (!training && mode == 1) {
                RC(bn_prepare_eval(s.ud, st));
                RC(run_conv(s.ud, x, x_lo, B_, H_, W_, 1, 0, nullptr, st));        // BN folded into the epilogue
            } else {
                if (!training) RC(bn_prepare_eval(s.ud, st));
                RC(conv_with_stats(s.ud, x, x_lo, B_, H_, W_, st));
                rscale = s.ud.scale; rshift = s.ud.shift;
            }
            res = s.ud.y;
        }
        s.out = out_pre ? out_pre : ar.f((size_t)Mo * C4);
        s.out_lo = mode == 2 ? (out_lo_pre ? out_lo_pre : ar.f((size_t)Mo * C4)) : nullptr;
        if (!bp.has_ds && mode == 2) {
            // parity mode: the identity is x_hi + x_lo -- fold the lo part in first
            float* xsum = ar.f((size_t)Mo * C4);
            if (!ar.dry) RC(tfe::masked_add(x, nullptr, x_lo, Mo * C4, xsum, st));
            res = xsum;
        }
        s.omask = training ? reinterpret_cast<unsigned int*>(ar.f((size_t)(Mo * C4 + 31) / 32)) : nullptr;
        if (!ar.dry) RC(tfe::bn_apply(s.u3.y, s.u3.scale, s.u3.shift, res, rscale, rshift, 1, Mo, C4, s.out, s.out_lo, mode, s.omask, st));
        return TF_OK;
    }

    bool pack_pending = false;
    // API mode 3 ("mixed"): 3xTF32 forward (mode == 2: outputs inside the 1e-3 tolerance, exact ReLU masks / statistics) with the
    // single-product TF32 backward of mode 1 -- the saved (hi, lo) activations and weights are used through their hi parts only
    bool fast_bwd = false;
    int bmode() const { return (mode == 2 && fast_bwd) ? 1 : mode; }
    // every dgrad (transposed, tap-flipped) weight of the backward pass, one launch
    int prepack_dgrad_weights(cudaStream_t st) {
        jobs.clear(); jobs_total = 0;
        for (const BlockP& bp : blocks) {
            prepack_add(bp.c1, bp.c1.cin, bp.c1.cout, 1); prepack_add(bp.c2, bp.c2.cin, bp.c2.cout, 1);
            prepack_add(bp.c3, bp.c3.cin, bp.c3.cout, 1);
            if (bp.has_ds) prepack_add(bp.cd, bp.cd.cin, bp.cd.cout, 1);
        }
        ConvP t3; t3.w = s3_w; t3.cin = 512; t3.cout = Cn; t3.k = 1;
        ConvP t4; t4.w = s4_w; t4.cin = 1024; t4.cout = Cn; t4.k = 1;
        prepack_add(t3, 512, Cp, 1); prepack_add(t4, 1024, Cp, 1);
        return prepack_flush(st);
    }
    int forward(const float* x_nchw, float* out_nchw, cudaStream_t st) {
        table_slot = 0;
        H2 = (H - 1) / 2 + 1; W2 = (W - 1) / 2 + 1;
        Hp = (H2 - 1) / 2 + 1; Wp = (W2 - 1) / 2 + 1;
        partial = ar.f((size_t)tfe::COLREDUCE_MAX_BLOCKS * 2 * 2 * 1024);   // main + tail-launch statistics rows
        bwd_slots = ar.f((size_t)tfe::BN_BWD_SLOTS * 2 * 1024);
        coef = ar.f(3 * 1024);
        dwtmp = ar.f((size_t)1024 * 1024 + 4096);
        offdiag = ar.f(64);
        {   // Weight re-layout, batched.  Training: only the stem's weight is packed on the caller's stream; every other
            // fprop weight AND the transposed / tap-flipped dgrad weights of the backward are packed on the side stream
            // while the stem (im2col, GEMM, BN, max-pool: ~1.4 ms of HBM-bound work) runs -- off the critical path.
            packed.clear(); jobs.clear(); jobs_total = 0;
            fold_scale = !training && mode == 1;
            if (!training) RC(prepare_eval_all(st)); else eval_ss.clear();     // (scale, shift) of every BN: the packing below folds the scales
            if (!ar.dry && training) RC(ensure_side());
            const bool aside = training && (ar.dry || side);
            cudaStream_t ps = (aside && !ar.dry) ? side : st;
            ConvP c = stem; c.cin = 147; c.k = 1;
            prepack_add(c, 64, 160, 0, eval_scale_of(stem_bn));
            if (aside) {
                RC(prepack_flush(st));
                if (!ar.dry) { TF_CHECK_CUDA(cudaEventRecord(ev_fork, st)); TF_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0)); }
            }
            for (const BlockP& bp : blocks) {
                prepack_add(bp.c1, bp.c1.cout, bp.c1.cin, 0, eval_scale_of(bp.b1));
                prepack_add(bp.c2, bp.c2.cout, bp.c2.cin, 0, eval_scale_of(bp.b2));
                prepack_add(bp.c3, bp.c3.cout, bp.c3.cin, 0, tfg::debug_flag(12) ? nullptr : eval_scale_of(bp.b3));   // (flag 12: conv3's BN in a separate pass)
                if (bp.has_ds) prepack_add(bp.cd, bp.cd.cout, bp.cd.cin, 0, eval_scale_of(bp.bd));
            }
            ConvP h3; h3.w = s3_w; h3.cin = 512; h3.cout = Cn; h3.k = 1;
            ConvP h4; h4.w = s4_w; h4.cin = 1024; h4.cout = Cn; h4.k = 1;
            prepack_add(h3, Cp, 512, 0); prepack_add(h4, Cp, 1024, 0);
            RC(prepack_flush(ps));
            if (aside && !ar.dry) TF_CHECK_CUDA(cudaEventRecord(ev_pack, side));          // layer1 waits for the fprop weights only
            if (training) {
                RC(prepack_dgrad_weights(ps));
                if (aside && !ar.dry) TF_CHECK_CUDA(cudaEventRecord(ev_pack_dgrad, side));   // joined at the end of the forward
            }
            pack_pending = aside && !ar.dry;
        }
        // ---- stem: im2col + GEMM (K = 147 padded to 160), BN, ReLU, max-pool
        const long long M2 = (long long)B * H2 * W2;
        col = ar.f((size_t)M2 * 160);
        col_lo = mode == 2 ? ar.f((size_t)M2 * 160) : nullptr;
        if (!ar.dry) RC(tfe::stem_im2col(x_nchw, B, H, W, H2, W2, 160, col, col_lo, mode, st));
        Unit& u = stem_u;
        u = Unit();
        u.c = stem; u.bn = stem_bn; u.B = B; u.H = H2; u.W = W2; u.Ho = H2; u.Wo = W2; u.x = col; u.x_lo = col_lo;
        {   // packed stem weight [64][1][160]: rows are the OIHW-flattened filters, zero padded
            ConvP c = stem; c.cin = 147; c.k = 1;
            RC(pack(c, 64, 160, 0, &u.wp, &u.wp_lo, st));
        }
        tfg::ConvArgs a = {};
        a.x = col; a.x_lo = col_lo; a.B = 1; a.H = 1; a.W = (int)M2; a.Cin = 160; a.w = u.wp; a.w_lo = u.wp_lo; a.Cout = 64; a.ksize = 1;
        TF_REQUIRE(M2 < (1ll << 31), "forward: stem pixel count overflows");
        float* a0; float* a0_lo = nullptr;
        if (!training && mode == 1) {
            RC(bn_prepare_eval(u, st));
            a.scale = fold_scale ? nullptr : u.scale; a.shift = u.shift; a.relu = 1; a.round_out = 0;
            a0 = ar.f((size_t)M2 * 64);
            a.y = a0; u.y = a0;
            if (!ar.dry) RC(tfg::conv_fprop(a, st));
        } else {
            if (!training) RC(bn_prepare_eval(u, st));
            u.y = ar.f((size_t)M2 * 64);
            a.y = u.y;
            int nblk = 0;
            if (training) { a.stats_partial = partial; a.stats_blocks = &nblk; }
            if (!ar.dry) RC(tfg::conv_fprop(a, st));
            if (training) {
                RC(alloc_bn(u));
                if (!ar.dry) RC(tfe::bn_finalize_train(partial, nblk, M2, 64, P(u.bn.gamma), P(u.bn.beta), eps, momentum, PW(u.bn.rm),
                                                       PW(u.bn.rv), u.scale, u.shift, u.mean, u.rstd, st));
            }
            a0 = ar.f((size_t)M2 * 64);
            u.amask = training ? reinterpret_cast<unsigned int*>(ar.f((size_t)(M2 * 64 + 31) / 32)) : nullptr;
            if (!ar.dry) RC(tfe::bn_apply(u.y, u.scale, u.shift, nullptr, nullptr, nullptr, 1, M2, 64, a0, nullptr, 0, u.amask, st));
        }
        u.a = a0; u.a_lo = a0_lo;
        const long long Mp = (long long)B * Hp * Wp;
        pool = ar.f((size_t)Mp * 64);
        pool_lo = mode == 2 ? ar.f((size_t)Mp * 64) : nullptr;
        pool_argmax = training ? reinterpret_cast<unsigned char*>(ar.f((size_t)Mp * 64 / 4)) : nullptr;
        if (!ar.dry) RC(tfe::maxpool_fwd(a0, B, H2, W2, 64, Hp, Wp, pool, pool_lo, mode, pool_argmax, st));
        // ---- heads' buffers (model.py:104-126) are allocated first so that they survive the per-block resets
        auto half = [](int v) { return (v - 1) / 2 + 1; };
        H3 = half(Hp); W3 = half(Wp); H4 = half(H3); W4 = half(W3);
        const long long M3 = (long long)B * H3 * W3, M4 = (long long)B * H4 * W4;
        ConvP h3; h3.w = s3_w; h3.cin = 512; h3.cout = Cn; h3.k = 1;
        ConvP h4; h4.w = s4_w; h4.cin = 1024; h4.cout = Cn; h4.k = 1;
        RC(pack(h3, Cp, 512, 0, &w3p, &w3p_lo, st));
        RC(pack(h4, Cp, 1024, 0, &w4p, &w4p_lo, st));
        b3p = ar.f(Cp); b4p = ar.f(Cp); up = ar.f((size_t)Cn * 16);
        float* s3 = ar.f((size_t)M3 * Cp); float* s4 = ar.f((size_t)M4 * Cp);
        if (!ar.dry) {
            TF_CHECK_CUDA(cudaMemsetAsync(b3p, 0, Cp * 4, st)); TF_CHECK_CUDA(cudaMemsetAsync(b4p, 0, Cp * 4, st));
            TF_CHECK_CUDA(cudaMemcpyAsync(b3p, P(s3_b), Cn * 4, cudaMemcpyDeviceToDevice, st));
            TF_CHECK_CUDA(cudaMemcpyAsync(b4p, P(s4_b), Cn * 4, cudaMemcpyDeviceToDevice, st));
            RC(tfe::extract_upsample_diag(P(up_w), Cn, up, offdiag, st));
        }
        // ---- residual stages.  Training keeps every tensor for the backward; inference ping-pongs the block
        //      outputs and reuses one scratch region per block (bounded memory for the 5000x5000 pyramid level).
        bs.assign(blocks.size(), BlockS());
        float* pp[2] = {nullptr, nullptr}; float* pp_lo[2] = {nullptr, nullptr};
        if (!training) {
            size_t mx = 0; int th = Hp, tw_ = Wp;
            for (const BlockP& bp : blocks) {
                if (bp.stride == 2) { th = half(th); tw_ = half(tw_); }
                mx = std::max(mx, (size_t)B * th * tw_ * bp.c3.cout);
            }
            for (int k = 0; k < 2; ++k) { pp[k] = ar.f(mx); pp_lo[k] = mode == 2 ? ar.f(mx) : nullptr; }
        }
        const size_t scratch_mark = ar.off;
        if (pack_pending) TF_CHECK_CUDA(cudaStreamWaitEvent(st, ev_pack, 0));   // packed fprop weights are ready
        const float* cur = pool; const float* cur_lo = pool_lo;
        int ch = Hp, cw = Wp;
        for (size_t i = 0; i < blocks.size(); ++i) {
            if (!training) ar.off = scratch_mark;
            RC(block_forward(blocks[i], bs[i], cur, cur_lo, B, ch, cw, pp[i & 1], pp_lo[i & 1], st));
            cur = bs[i].out; cur_lo = bs[i].out_lo; ch = bs[i].Ho; cw = bs[i].Wo;
            if (i == 6) {
                res3 = cur; res3_lo = cur_lo;
                TF_REQUIRE(ch == H3 && cw == W3, "forward: res3 shape mismatch");
                if (!ar.dry) {
                    tfg::ConvArgs h = {};
                    h.x = res3; h.x_lo = res3_lo; h.B = B; h.H = H3; h.W = W3; h.Cin = 512; h.w = w3p; h.w_lo = w3p_lo; h.Cout = Cp; h.ksize = 1; h.shift = b3p; h.y = s3;
                    RC(tfg::conv_fprop(h, st));
                }
            }
        }
        res4 = cur; res4_lo = cur_lo;
        TF_REQUIRE(ch == H4 && cw == W4, "forward: res4 shape mismatch");
        if (!ar.dry) {
            tfg::ConvArgs h = {};
            h.x = res4; h.x_lo = res4_lo; h.B = B; h.H = H4; h.W = W4; h.Cin = 1024; h.w = w4p; h.w_lo = w4p_lo; h.Cout = Cp; h.ksize = 1; h.shift = b4p; h.y = s4;
            RC(tfg::conv_fprop(h, st));
            RC(tfe::head_combine_fwd(s3, s4, up, B, H3, W3, H4, W4, Cn, Cp, out_nchw, st));
        }
        // the dgrad weights were packed on the side stream behind the whole forward: join it (free by now)
        if (pack_pending) { TF_CHECK_CUDA(cudaStreamWaitEvent(st, ev_pack_dgrad, 0)); pack_pending = false; }
        fwd_mark = ar.off;
        return TF_OK;
    }

    // ------------------------------------------------------------------------------------------ backward pieces
    float* G(void* const* grads, int idx) const { return (grads && idx >= 0) ? reinterpret_cast<float*>(grads[idx]) : nullptr; }

    // BN backward of unit u: dout (masked by act > 0 when act != null) -> dy (+lo); dgamma/dbeta into the caller's grads
    int unit_bn_bwd(const Unit& u, const float* dout, const unsigned int* mask, float* gmask_out, float** dy, float** dy_lo,
                    void* const* grads, cudaStream_t st) {
        const long long M = (long long)u.B * u.Ho * u.Wo;
        const int C = u.bn.C;
        *dy = ar.f((size_t)M * C);
        *dy_lo = bmode() == 2 ? ar.f((size_t)M * C) : nullptr;
        if (!ar.dry) RC(tfe::bn_backward(dout, nullptr, mask, u.y, u.mean, u.rstd, P(u.bn.gamma), M, C, G(grads, u.bn.gamma), G(grads, u.bn.beta),
                                         *dy, *dy_lo, gmask_out, bmode(), bwd_slots, coef, st));
        return TF_OK;
    }
    // conv backward of unit u given dy at (Ho, Wo): weight gradient (always) and input gradient into dx
    // (dx_accumulate: add into dx instead of overwriting); dx == null skips the dgrad.
    int unit_conv_bwd(const Unit& u, const float* dy, const float* dy_lo, float* dx, int dx_accumulate, int dxH, int dxW,
                      void* const* grads, cudaStream_t st, const float* res = nullptr, const unsigned int* res_mask = nullptr) {
        const ConvP& c = u.c;
        const int taps = c.k * c.k;
        // ---- wgrad (a stride-2 conv reads x through the TMA traversal stride; dy stays at the output resolution)
        float* gw = G(grads, c.w);
        tfg::WgradArgs w = {};
        const bool want_w = gw && !ar.dry;
        if (want_w) {
            w.x = u.x; w.x_lo = u.x_lo; w.dy = dy; w.dy_lo = dy_lo; w.B = u.B; w.H = u.H; w.W = u.W; w.Cin = c.cin; w.Cout = c.cout;
            w.ksize = c.k; w.stride = c.stride;
            if (w.x_lo == nullptr || w.dy_lo == nullptr) { w.x_lo = nullptr; w.dy_lo = nullptr; }
            if (wgrad_mode != 1 || !side || !dx) RC(wgrad_async(w, (size_t)c.cout * taps * c.cin, c.cout, c.cin, taps, c.cin, gw, st));
            else TF_CHECK_CUDA(cudaEventRecord(ev_fork, st));          // inputs are final here; enqueue after the dgrad
        }
        // ---- dgrad: dx = conv(dy, w^T flipped)
        if (dx) {
            float *wt, *wt_lo;
            ConvP ct = c;
            RC(pack(ct, c.cin, c.cout, 1, &wt, &wt_lo, st));        // [Cin][taps][Cout]
            tfg::ConvArgs a = {};
            a.B = u.B; a.Cin = c.cout; a.w = wt; a.w_lo = wt_lo; a.Cout = c.cin; a.ksize = c.k;
            const bool phase_s2 = !tfg::debug_flag(11);     // tf_debug_set(11, 1): old zero-insert path (A/B)
            if (c.stride == 2 && phase_s2 && (c.k == 3 || dx_accumulate)) {
                // stride-2 dgrad as 4 parity-class GEMMs straight from dy (no zero insertion, 4x fewer FLOPs for the 3x3);
                // the 1x1 form only touches the even pixels, so it must ADD into a dx that already holds the main path
                a.x = dy; a.x_lo = dy_lo; a.H = u.H; a.W = u.W; a.y = dx; a.accumulate = dx_accumulate;
                if (a.x_lo == nullptr) a.w_lo = nullptr;
                if (!ar.dry) RC(tfg::conv_dgrad_s2(a, st));
            } else if (c.stride == 2 && c.k == 3) {
                // zero-insert dy to the input resolution, then a stride-1 conv with the flipped kernel
                float* z = ar.f((size_t)u.B * u.H * u.W * c.cout);
                float* z_lo = dy_lo ? ar.f((size_t)u.B * u.H * u.W * c.cout) : nullptr;
                if (!ar.dry) RC(tfe::zero_insert2(dy, dy_lo, u.B, u.H, u.W, c.cout, z, z_lo, st));
                a.x = z; a.x_lo = z_lo; a.H = u.H; a.W = u.W; a.y = dx; a.accumulate = dx_accumulate;
                if (a.x_lo == nullptr) a.w_lo = nullptr;
                if (!ar.dry) RC(tfg::conv_fprop(a, st));
            } else if (c.stride == 2) {
                // 1x1/s2: gradient w.r.t. the sampled pixels at the output resolution, scattered to the even positions
                float* dxs = ar.f((size_t)u.B * u.Ho * u.Wo * c.cin);
                a.x = dy; a.x_lo = dy_lo; a.H = u.Ho; a.W = u.Wo; a.y = dxs;
                if (a.x_lo == nullptr) a.w_lo = nullptr;
                if (!ar.dry) {
                    RC(tfg::conv_fprop(a, st));
                    TF_REQUIRE(!dx_accumulate, "unit_conv_bwd: accumulate into a strided dgrad is not supported");
                    RC(tfe::zero_insert2(dxs, nullptr, u.B, dxH, dxW, c.cin, dx, nullptr, st));
                }
            } else {
                a.x = dy; a.x_lo = dy_lo; a.H = u.H; a.W = u.W; a.y = dx; a.accumulate = dx_accumulate;
                if (res) { a.res = res; a.res_mask = res_mask; a.accumulate = 0; }     // dx = dgrad + masked upstream gradient
                if (a.x_lo == nullptr) a.w_lo = nullptr;
                if (!ar.dry) RC(tfg::conv_fprop(a, st));
            }
            if (want_w && wgrad_mode == 1 && side) RC(wgrad_async(w, (size_t)c.cout * taps * c.cin, c.cout, c.cin, taps, c.cin, gw, st, true));
        }
        return TF_OK;
    }
    int block_backward(const BlockS& s, const float* dout, float* dx, void* const* grads, cudaStream_t st) {
        const long long Mo = (long long)s.B * s.Ho * s.Wo;
        const int C4 = s.u3.c.cout;
        float *dy3, *dy3_lo;
        const bool has_ds = s.has_ds;
        // G = dout * [out > 0]: it feeds the downsample BN when there is one; with an identity shortcut it is never
        // materialised -- the conv1 dgrad's residual epilogue adds the masked dout itself (TMA-loaded boxes; -0.5 ms/step
        // against "dx starts as G, written by the BN-backward pass, and the dgrad reduce-adds onto it" = tf_debug_set(8, 2))
        const bool fuse_g = !has_ds && s.u1.c.k == 1 && s.u1.c.stride == 1 && bmode() == 1 && tfg::debug_flag(8) != 2;
        float* g = has_ds ? ar.f((size_t)Mo * C4) : (fuse_g ? nullptr : dx);
        RC(unit_bn_bwd(s.u3, dout, s.omask, g, &dy3, &dy3_lo, grads, st));
        const long long M2o = (long long)s.B * s.u2.Ho * s.u2.Wo;
        float* da2 = ar.f((size_t)M2o * s.u2.c.cout);
        RC(unit_conv_bwd(s.u3, dy3, dy3_lo, da2, 0, s.Ho, s.Wo, grads, st));
        float *dy2, *dy2_lo;
        RC(unit_bn_bwd(s.u2, da2, s.u2.amask, nullptr, &dy2, &dy2_lo, grads, st));
        const long long M1 = (long long)s.B * s.H * s.W;
        float* da1 = ar.f((size_t)M1 * s.u1.c.cout);
        RC(unit_conv_bwd(s.u2, dy2, dy2_lo, da1, 0, s.H, s.W, grads, st));
        float *dy1, *dy1_lo;
        RC(unit_bn_bwd(s.u1, da1, s.u1.amask, nullptr, &dy1, &dy1_lo, grads, st));
        if (has_ds && s.ud.c.stride == 2 && !tfg::debug_flag(11)) {
            // strided downsample: dx = main path first, then the 1x1/s2 dgrad adds into its even pixels
            float *dyd, *dyd_lo;
            RC(unit_bn_bwd(s.ud, g, nullptr, nullptr, &dyd, &dyd_lo, grads, st));
            RC(unit_conv_bwd(s.u1, dy1, dy1_lo, dx, 0, s.H, s.W, grads, st));          // dx = main path
            RC(unit_conv_bwd(s.ud, dyd, dyd_lo, dx, 1, s.H, s.W, grads, st));          // dx += downsample path
            return TF_OK;
        }
        if (has_ds) {
            float *dyd, *dyd_lo;
            RC(unit_bn_bwd(s.ud, g, nullptr, nullptr, &dyd, &dyd_lo, grads, st));
            RC(unit_conv_bwd(s.ud, dyd, dyd_lo, dx, 0, s.H, s.W, grads, st));          // dx = downsample path
        }
        if (fuse_g) RC(unit_conv_bwd(s.u1, dy1, dy1_lo, dx, 0, s.H, s.W, grads, st, dout, s.omask));   // dx = main path + G
        else RC(unit_conv_bwd(s.u1, dy1, dy1_lo, dx, 1, s.H, s.W, grads, st));         // dx += main path
        return TF_OK;
    }

    int backward(const float* dout_nchw, void* const* grads, cudaStream_t caller) {
        TF_REQUIRE(training, "tf_model_backward: the last forward was not a training forward");
        ar.off = fwd_mark;
        if (!ar.dry) RC(ensure_side());
        pending.clear();
        bool region_recorded[2] = {false, false};     // (waiting on a previous call's event is pointless and breaks stream capture)
        size_t next_event = 0;
        // experiment switches: tf_debug_set(5, 1 + mode) picks the weight-gradient schedule, tf_debug_set(6, 1) keeps the
        // chain on the caller's stream (no priority)
        wgrad_mode = tfg::debug_flag(5) ? tfg::debug_flag(5) - 1 : 2;
        chain_priority = tfg::debug_flag(6) ? 0 : 1;
        // fork: the whole chain runs on the high-priority stream, the caller's stream waits for it at the end
        cudaStream_t st = caller;
        if (!ar.dry && side && chain_priority) {
            TF_CHECK_CUDA(cudaEventRecord(ev_chain[0], caller));
            TF_CHECK_CUDA(cudaStreamWaitEvent(chain, ev_chain[0], 0));
            st = chain;
        }
        const long long M3 = (long long)B * H3 * W3, M4 = (long long)B * H4 * W4;
        float* ds3 = ar.f((size_t)M3 * Cp); float* ds4 = ar.f((size_t)M4 * Cp);
        // parity mode: the head gradients are exact (hi, lo) splits too, so that the head dgrad / wgrad GEMMs run the same
        // 3xTF32 products as the trunk (round 1 fed them single TF32 operands: 5e-4 on every gradient below the heads)
        float* ds3_lo = bmode() == 2 ? ar.f((size_t)M3 * Cp) : nullptr; float* ds4_lo = bmode() == 2 ? ar.f((size_t)M4 * Cp) : nullptr;
        if (!ar.dry) {
            TF_CHECK_CUDA(cudaMemsetAsync(bwd_slots, 0, (size_t)tfe::BN_BWD_SLOTS * 2 * 1024 * sizeof(float), st));   // every BN backward leaves it zeroed
            RC(tfe::head_combine_bwd(dout_nchw, up, B, H3, W3, H4, W4, Cn, Cp, ds3, ds4, st, ds3_lo, ds4_lo));
            // bias gradients: column sums of hi (+ lo): the lo parts are ~2^-12 of hi, their sum is added by a second pass
            if (G(grads, s3_b)) RC(bias_grad(ds3, ds3_lo, M3, G(grads, s3_b), st));
            if (G(grads, s4_b)) RC(bias_grad(ds4, ds4_lo, M4, G(grads, s4_b), st));
        }
        // head weight gradients and input gradients (as 1x1 "units" with padded Cout)
        Unit h3; h3.c.w = s3_w; h3.c.cin = 512; h3.c.cout = Cp; h3.c.k = 1; h3.x = res3; h3.x_lo = res3_lo; h3.B = B; h3.H = H3; h3.W = W3; h3.Ho = H3; h3.Wo = W3;
        Unit h4 = h3; h4.c.w = s4_w; h4.c.cin = 1024; h4.x = res4; h4.x_lo = res4_lo; h4.H = H4; h4.W = W4; h4.Ho = H4; h4.Wo = W4;
        float* dres4 = ar.f((size_t)M4 * 1024);
        // input gradients ping-pong between two buffers; one scratch region is reused by every other block
        size_t mx = 0;
        for (const BlockS& s : bs) mx = std::max(mx, (size_t)s.B * s.H * s.W * s.u1.c.cin);
        float* dxbuf[2] = {ar.f(mx), ar.f(mx)};
        RC(head_bwd(h4, ds4, ds4_lo, dres4, 0, grads, st));
        const size_t scratch_mark = ar.off;
        // Two scratch regions alternate between blocks: the side stream may still be reading block i's dy tensors
        // (weight gradients) while the chain already works on block i-1 (and, with deferred weight gradients, i-2 has
        // not started): region r is reused only after the side-stream work that read it has finished (ev_region[r]).
        const size_t region = ar.dry ? 0 : bwd_region_bytes;
        size_t max_used = 0;
        // ---- layer3 .. layer1
        const float* dcur = dres4;
        for (int i = (int)blocks.size() - 1; i >= 0; --i) {
            const BlockS& s = bs[i];
            const int r = i & 1;
            ar.off = scratch_mark + (size_t)r * region;
            if (!ar.dry && side) {
                // deferred weight gradients of block i+1 (other region): start them now, next to this block's bn3 backward
                RC(wgrad_flush(st));
                // bucket events (tf_model_backward_ex): every weight / BN gradient of the blocks > i is now enqueued
                while (next_event < bucket_events.size() && bucket_blocks[next_event] >= i) {
                    if (wgrad_mode != 2) break;          // (the other schedules have no per-block flush points)
                    TF_CHECK_CUDA(cudaEventRecord((cudaEvent_t)bucket_events[next_event], side));
                    ++next_event;
                }
                if (wgrad_mode == 2) { TF_CHECK_CUDA(cudaEventRecord(ev_region[r ^ 1], side)); region_recorded[r ^ 1] = true; }
                if (region_recorded[r]) TF_CHECK_CUDA(cudaStreamWaitEvent(st, ev_region[r], 0));
            }
            float* dx = dxbuf[i & 1];
            RC(block_backward(s, dcur, dx, grads, st));
            if (i == 7) RC(head_bwd(h3, ds3, ds3_lo, dx, 1, grads, st));   // res3 also feeds score_res3: dx(block 7 input) += head dgrad
            if (!ar.dry && side && wgrad_mode != 2) { TF_CHECK_CUDA(cudaEventRecord(ev_region[r], side)); region_recorded[r] = true; }
            max_used = std::max(max_used, ar.off - (scratch_mark + (size_t)r * region));
            dcur = dx;
        }
        if (ar.dry) bwd_region_bytes = tf_align_up(max_used, 1024);
        else if (side) RC(wgrad_flush(st));
        // the stem's scratch lives BEHIND the two regions (1.3 GB more workspace at batch-8 960x1280): the chain does not
        // have to wait for the layer-1 weight gradients still running on the side stream
        ar.off = scratch_mark + 2 * bwd_region_bytes;
        ar.peak = std::max(ar.peak, ar.off + 4096);
        // ---- stem
        const long long M2 = (long long)B * H2 * W2;
        float* da0 = ar.f((size_t)M2 * 64);
        if (!ar.dry) RC(tfe::maxpool_bwd(pool_argmax, dcur, B, H2, W2, 64, Hp, Wp, da0, st));
        float *dy0, *dy0_lo;
        RC(unit_bn_bwd(stem_u, da0, stem_u.amask, nullptr, &dy0, &dy0_lo, grads, st));
        float* gw = G(grads, stem.w);
        if (gw && !ar.dry) {
            tfg::WgradArgs w = {};
            w.x = col; w.x_lo = col_lo; w.dy = dy0; w.dy_lo = dy0_lo; w.B = 1; w.H = 1; w.W = (int)M2; w.Cin = 160; w.Cout = 64; w.ksize = 1;
            if (!w.x_lo || !w.dy_lo) { w.x_lo = nullptr; w.dy_lo = nullptr; }
            RC(wgrad_async(w, (size_t)64 * 160, 64, 147, 1, 160, gw, st));
        }
        if (!ar.dry && side) {                       // join: the caller's stream sees every gradient
            RC(wgrad_flush(st));
            for (; next_event < bucket_events.size(); ++next_event)       // remaining buckets (layer 1, stem): complete here
                TF_CHECK_CUDA(cudaEventRecord((cudaEvent_t)bucket_events[next_event], side));
            TF_CHECK_CUDA(cudaEventRecord(ev_join, side));
            TF_CHECK_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
            if (st != caller) {
                TF_CHECK_CUDA(cudaEventRecord(ev_chain[1], st));
                TF_CHECK_CUDA(cudaStreamWaitEvent(caller, ev_chain[1], 0));
            }
        } else if (!ar.dry) {
            for (; next_event < bucket_events.size(); ++next_event) TF_CHECK_CUDA(cudaEventRecord((cudaEvent_t)bucket_events[next_event], st));
        }
        return TF_OK;
    }
    // bias gradient of a head: column sums of ds (hi) and, in parity mode, of its lo part added on top
    int bias_grad(const float* ds, const float* ds_lo, long long M, float* gb, cudaStream_t st) {
        RC(tfe::column_sum(ds, M, Cp, Cn, gb, partial, st));
        if (ds_lo) RC(tfe::column_sum(ds_lo, M, Cp, Cn, gb, partial, st, 1));
        return TF_OK;
    }
    // head (score_res3 / score_res4) backward: weight gradient [Cn, Cin] and dres (+)= ds * W
    int head_bwd(const Unit& h, const float* ds, const float* ds_lo, float* dres, int accumulate, void* const* grads, cudaStream_t st) {
        const ConvP& c = h.c;
        float* gw = G(grads, c.w);
        if (gw && !ar.dry) {
            tfg::WgradArgs w = {};
            w.x = h.x; w.dy = ds; w.B = h.B; w.H = h.H; w.W = h.W; w.Cin = c.cin; w.Cout = Cp; w.ksize = 1;
            if (h.x_lo && ds_lo) { w.x_lo = h.x_lo; w.dy_lo = ds_lo; }
            RC(wgrad_async(w, (size_t)Cp * c.cin, Cn, c.cin, 1, c.cin, gw, st));
        }
        float *wt, *wt_lo;
        ConvP ct = c; ct.cout = Cn;
        RC(pack(ct, c.cin, Cp, 1, &wt, &wt_lo, st));                 // [Cin][1][Cp], columns >= Cn are zero
        tfg::ConvArgs a = {};
        a.x = ds; a.B = h.B; a.H = h.H; a.W = h.W; a.Cin = Cp; a.w = wt; a.Cout = c.cin; a.ksize = 1; a.y = dres; a.accumulate = accumulate;
        if (ds_lo && wt_lo) { a.x_lo = ds_lo; a.w_lo = wt_lo; }
        if (!ar.dry) RC(tfg::conv_fprop(a, st));
        return TF_OK;
    }
};

}  // namespace

TF_API int tf_model_backward_ex(void* handle, const float* dout, void* const* grads, int num_events, void* const* events,
                                const int* event_first_block, void* stream);

TF_API int tf_model_create(int num_templates, void** handle) {
    TF_REQUIRE(handle && num_templates > 0 && num_templates <= 32, "tf_model_create: bad args");
    *handle = new Model(num_templates);
    return TF_OK;
}
TF_API int tf_model_destroy(void* handle) { delete reinterpret_cast<Model*>(handle); return TF_OK; }
TF_API int tf_model_num_params(void* handle) { return handle ? (int)reinterpret_cast<Model*>(handle)->names.size() : -1; }
TF_API const char* tf_model_param_name(void* handle, int i) {
    Model* m = reinterpret_cast<Model*>(handle);
    return (m && i >= 0 && i < (int)m->names.size()) ? m->names[i].c_str() : nullptr;
}
TF_API int tf_model_output_shape(void* handle, int H, int W, int* H3, int* W3) {
    TF_REQUIRE(handle && H3 && W3 && H > 0 && W > 0, "tf_model_output_shape: bad args");
    auto half = [](int v) { return (v - 1) / 2 + 1; };
    *H3 = half(half(half(H))); *W3 = half(half(half(W)));
    return TF_OK;
}
// Workspace needed by forward (+ backward when training) for this shape / mode (dry run of the same code path).
TF_API int tf_model_workspace_bytes(void* handle, int B, int H, int W, int training, int mode, size_t* bytes) {
    TF_REQUIRE(handle && bytes && B > 0 && H >= 16 && W >= 16 && mode >= 1 && mode <= 3, "tf_model_workspace_bytes: bad args");
    Model* m = reinterpret_cast<Model*>(handle);
    m->plan_training = -1;                       // any explicit query invalidates the cached plan of tf_model_forward
    m->B = B; m->H = H; m->W = W; m->training = training; m->mode = mode == 3 ? 2 : mode; m->fast_bwd = mode == 3;
    m->ar = Arena(); m->ar.dry = true; m->ar.base = nullptr;
    RC(m->forward(nullptr, nullptr, nullptr));
    if (training) RC(m->backward(nullptr, nullptr, nullptr));
    *bytes = m->ar.peak + 4096;
    return TF_OK;
}
// x: [B,3,H,W] fp32 NCHW (device).  params: num_params device pointers in tf_model_param_name order.
// out: [B,5T,H3,W3] fp32 NCHW (device).  training != 0: batch-statistics BN, running stats updated in place,
// everything the backward needs stays in the workspace until the next forward.
TF_API int tf_model_forward(void* handle, const float* x, int B, int H, int W, const void* const* params, int training,
                            int mode, float bn_momentum, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(handle && x && params && out && workspace, "tf_model_forward: null pointer");
    TF_REQUIRE(B > 0 && H >= 16 && W >= 16 && mode >= 1 && mode <= 3, "tf_model_forward: bad shape/mode");
    Model* m = reinterpret_cast<Model*>(handle);
    size_t need;
    if (m->plan_B == B && m->plan_H == H && m->plan_W == W && m->plan_training == training && m->plan_mode == mode &&
        m->plan_epoch == tfg::debug_epoch()) {
        need = m->plan_need;                     // same shape as the last call: the dry-run plan is still valid
    } else {
        RC(tf_model_workspace_bytes(handle, B, H, W, training, mode, &need));
        m->plan_B = B; m->plan_H = H; m->plan_W = W; m->plan_training = training; m->plan_mode = mode; m->plan_need = need;
        m->plan_epoch = tfg::debug_epoch();
    }
    m->B = B; m->H = H; m->W = W; m->training = training; m->mode = mode == 3 ? 2 : mode; m->fast_bwd = mode == 3;
    if (workspace_bytes < need) { tf_set_error("tf_model_forward: workspace %zu < required %zu", workspace_bytes, need); return TF_ERR_WORKSPACE; }
    m->params.assign(params, params + m->names.size()); m->momentum = bn_momentum;
    m->ar = Arena(); m->ar.dry = false; m->ar.base = reinterpret_cast<char*>(workspace); m->ar.cap = workspace_bytes;
    return m->forward(x, out, (cudaStream_t)stream);
}
// dout: [B,5T,H3,W3] NCHW.  grads: num_params device pointers (null = not wanted); conv / head weights in OIHW,
// BN gamma/beta, head biases.  Must follow a training forward on the same workspace.
TF_API int tf_model_backward(void* handle, const float* dout, void* const* grads, void* stream) {
    return tf_model_backward_ex(handle, dout, grads, 0, nullptr, nullptr, stream);
}
// tf_model_backward that also records `num_events` caller-owned cudaEvent_t's while the backward is being enqueued: event k is
// recorded as soon as every gradient (conv weights, BN gamma / beta, head weights / biases) of the residual blocks with forward
// index >= event_first_block[k] has been enqueued (blocks 0..29 = layer1.0 .. layer3.22; score_res4 counts as block 29, score_res3 as block 7, the
// stem as block -1; event_first_block must be descending).  A gradient all-reduce bucket that waits on event k therefore
// overlaps the rest of the backward.  Events with first_block <= 0 complete at the end.
TF_API int tf_model_backward_ex(void* handle, const float* dout, void* const* grads, int num_events, void* const* events,
                                const int* event_first_block, void* stream) {
    TF_REQUIRE(handle && dout && grads, "tf_model_backward: null pointer");
    TF_REQUIRE(num_events >= 0 && num_events <= 64 && (num_events == 0 || (events && event_first_block)), "tf_model_backward_ex: bad events");
    Model* m = reinterpret_cast<Model*>(handle);
    TF_REQUIRE(!m->ar.dry && m->ar.base, "tf_model_backward: no forward state");
    m->bucket_events.clear(); m->bucket_blocks.clear();
    for (int k = 0; k < num_events; ++k) {
        TF_REQUIRE(events[k] && (k == 0 || event_first_block[k] <= event_first_block[k - 1]), "tf_model_backward_ex: events must be non-null, first blocks descending");
        m->bucket_events.push_back(events[k]);
        m->bucket_blocks.push_back(event_first_block[k] - 1);        // recorded when the loop reaches block index first_block - 1
    }
    const int rc = m->backward(dout, grads, (cudaStream_t)stream);
    m->bucket_events.clear(); m->bucket_blocks.clear();
    return rc;
}
// Debug / test hook: locate an internal NHWC activation of the last forward ("stem", "pool", "block<i>.out",
// "block<i>.u<1|2|3|d>.<y|a>") and copy it (device to device) into dst (capacity in floats).
TF_API int tf_model_get_tensor(void* handle, const char* name, float* dst, int64_t capacity, int* shape4, void* stream) {
    TF_REQUIRE(handle && name && shape4, "tf_model_get_tensor: bad args");
    Model* m = reinterpret_cast<Model*>(handle);
    TF_REQUIRE(!m->ar.dry && m->ar.base, "tf_model_get_tensor: no forward state");
    const std::string n(name);
    const float* src = nullptr; const float* src_lo = nullptr; int B = m->B, H = 0, W = 0, C = 0;
    if (n == "stem") { src = m->stem_u.a; H = m->H2; W = m->W2; C = 64; }
    else if (n == "stem.y") { src = m->stem_u.y; H = m->H2; W = m->W2; C = 64; }
    else if (n == "col") { src = m->col; H = m->H2; W = m->W2; C = 160; }
    else if (n == "pool") { src = m->pool; src_lo = m->pool_lo; H = m->Hp; W = m->Wp; C = 64; }
    else if (n.rfind("block", 0) == 0) {
        const size_t dot = n.find('.');
        TF_REQUIRE(dot != std::string::npos, "tf_model_get_tensor: bad name %s", name);
        const int i = atoi(n.substr(5, dot - 5).c_str());
        TF_REQUIRE(i >= 0 && i < (int)m->bs.size(), "tf_model_get_tensor: bad block index in %s", name);
        const BlockS& s = m->bs[i];
        const std::string rest = n.substr(dot + 1);
        if (rest == "out") { src = s.out; src_lo = s.out_lo; H = s.Ho; W = s.Wo; C = s.u3.c.cout; }
        else {
            const Unit* u = rest[1] == '1' ? &s.u1 : rest[1] == '2' ? &s.u2 : rest[1] == '3' ? &s.u3 : &s.ud;
            H = u->Ho; W = u->Wo; C = u->c.cout;
            src = rest.back() == 'y' ? u->y : u->a;
            if (rest.back() != 'y') src_lo = u->a_lo;
        }
    }
    TF_REQUIRE(src, "tf_model_get_tensor: unknown or empty tensor %s", name);
    shape4[0] = B; shape4[1] = H; shape4[2] = W; shape4[3] = C;
    const int64_t cnt = (int64_t)B * H * W * C;
    if (dst) {
        TF_REQUIRE(capacity >= cnt, "tf_model_get_tensor: capacity too small");
        if (src_lo) RC(tfe::masked_add(src, nullptr, src_lo, cnt, dst, (cudaStream_t)stream));     // parity mode: hi + lo
        else TF_CHECK_CUDA(cudaMemcpyAsync(dst, src, cnt * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return TF_OK;
}
// max |off-diagonal| of score4_upsample.weight seen by the last forward (host value; synchronises the stream)
TF_API int tf_model_upsample_offdiag(void* handle, float* value, void* stream) {
    TF_REQUIRE(handle && value, "tf_model_upsample_offdiag: bad args");
    Model* m = reinterpret_cast<Model*>(handle);
    TF_REQUIRE(m->offdiag && !m->ar.dry, "tf_model_upsample_offdiag: no forward state");
    TF_CHECK_CUDA(cudaMemcpyAsync(value, m->offdiag, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TF_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return TF_OK;
}
