// tcgen05 / TMA implicit-GEMM convolution kernels for NHWC fp32 activations (TF32 tensor-core math,
// fp32 accumulation in TMEM).  They stand in for the 96 nn.Conv2d calls on the reference's hot path
// (/root/reference/tinyfaces/models/model.py:90-106 -> torchvision resnet.py Bottleneck convs) and for the
// dgrad / wgrad GEMMs autograd runs for them (/root/reference/tinyfaces/trainer.py:86).
//
//  conv_gemm_kernel  (fprop + dgrad):  D[pixel, co] = sum_{tap, ci} A[pixel (+) tap, ci] * Wp[co, tap, ci]
//      A tiles: 128 pixels x 32 channels (128 B rows, SWIZZLE_128B) fetched by TMA straight from the NHWC
//      tensor -- a 2-D map for 1x1 convs, a 4-D map (C, W, H, B) for 3x3 where the tap shift is just a
//      coordinate offset and the zero padding is TMA's out-of-bounds fill.  Both operands K-major.
//      Persistent CTAs, warp-specialised: warp 0 TMA producer, warp 1 tcgen05.mma issuer (one thread),
//      warps 2-5 epilogue (tcgen05.ld -> swizzled smem -> TMA store).  Two TMEM accumulator stages so the
//      epilogue of tile i overlaps the MMAs of tile i+1.
//  conv_wgrad_kernel:  dW[co, tap, ci] += sum_{pixel} dY[pixel, co] * X[pixel (+) tap, ci]
//      Both operands MN-major (channels contiguous, reduction over pixels), split-K over pixel blocks,
//      epilogue = TMA reduce-add (fp32) into the packed gradient.
//
// A K loop may run over up to 3 (A, B) tensor-map pairs ("segments"): the 3xTF32 parity mode feeds
// (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo) through the same accumulator.
#include "tf_common.cuh"
#include "tf_umma.cuh"
#include <mutex>

using namespace tfu;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                 // fp32 elements = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * 128;
constexpr int EPI_BUF_BYTES = 128 * 128;    // 128 rows x 32 fp32
constexpr int GEMM_THREADS = 192;
constexpr int MAX_SEG = 3;
constexpr int MAX_TAPS = 9;

struct GemmMaps {
    CUtensorMap a[MAX_SEG];
    CUtensorMap b[MAX_SEG];
    CUtensorMap d;
    CUtensorMap res;              // RES kernels: the residual tensor, same geometry as d (32 x 32 boxes, SWIZZLE_128B)
};

struct GemmParams {
    int num_m_tiles, num_n_tiles;
    int spatial;                  // 0: A/D are 2-D [pixels, C]; 1: 4-D (C, W, H, B) with spatial tiles
    int tiles_x, tiles_y, tw, th; // spatial tiling of one image (tw * th == 128)
    int taps, kblocks, nseg;      // K loop = nseg x taps x kblocks blocks of 32 channels
    // tap t reads A at (x0 + tap_ox[t], y0 + tap_oy[t]) (x stride) and B at K block (tap_w[t] * kblocks + kb): the full 3x3
    // stencil by default; the parity classes of a stride-2 dgrad use 1, 2 or 4 taps with 0 / +1 offsets
    signed char tap_ox[9], tap_oy[9]; unsigned char tap_w[9];
    int stride;                   // spatial mode: input coordinate = stride * output coordinate + tap offset (TMA elementStrides)
    // Tail-wave split-K: work units [0, main_tiles) are whole tiles; every later tile is cut into `ksplit` K slices
    // (one unit each) whose partial sums meet in D through TMA reduce-add (the host zeroes those rows first).  Without
    // it 300 tiles on 148 SMs cost 3 rounds for 2.03 rounds of work.
    int main_tiles, ksplit;
    // Residual epilogue (RES kernels, flat pixel rows only): D[row, n] = f(acc + res[row, n]) -- res_mask (optional, 1 bit
    // per element) gates the residual.  Each epilogue warp TMA-loads the 32 x 32 residual box of the chunk two chunks
    // ahead into a private swizzled buffer (these kernels give one or two operand stages to those buffers).  Used by the
    // inference path (conv3 + folded BN + shortcut + ReLU in one kernel, no separate BN-apply pass).
    const float* res; const unsigned int* res_mask; long long m_rows;
    const float* scale;           // optional per-output-channel epilogue: v = v * scale[n] + shift[n]
    const float* shift;           //   (shift alone = bias), then optional ReLU and TF32 rounding
    int relu, round_out;
    int accumulate;               // 1: D += result (TMA reduce-add) instead of D = result
    float* stats_partial;         // optional BN statistics: [grid/num_n_tiles][2][Cout] per-channel (sum, sum^2) partials, one row per CTA
    int cout;                     //   of the raw accumulators (needs gridDim.x % num_n_tiles == 0: fixed n-tile per CTA)
    int img_w, img_h;             // spatial mode: rows of a patch that fall outside the image are not statistics
    int* err_flag;
};

template <int BN> struct GemmCfg {
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int B_STAGE_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * EPI_BUF_BYTES + 1024 /*barriers*/ + 1024 /*align*/;
    static constexpr int RES_DROP = (32768 + STAGE_BYTES - 1) / STAGE_BYTES;   // operand stages given to the residual buffers (4 warps x 2 x 4 KB)
    static constexpr int FUSED_SMEM_BYTES = SMEM_BYTES + 1024;   // + the CTA's per-channel shift vector (= the 227 KB maximum for BN = 256)
    static constexpr int TMEM_COLS = 2 * BN;
};

// 3x3 taps are enumerated row-major: tap = (dy+1)*3 + (dx+1)
__device__ __forceinline__ int tap_dx(int taps, int tap) { return taps == 9 ? tap % 3 - 1 : 0; }
__device__ __forceinline__ int tap_dy(int taps, int tap) { return taps == 9 ? tap / 3 - 1 : 0; }

__device__ __forceinline__ void named_bar_sync_epi() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// The staging buffers are reached through 32-bit shared-window addresses with explicit st.shared / ld.shared: a generic
// pointer derived from the aligned dynamic-smem base makes the compiler emit generic LD/ST (seen in the round-1 SASS),
// which are slower than LDS/STS and were the top long-scoreboard stall of the statistics epilogue.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void store_row_swizzled(uint32_t buf_s, int row, const uint32_t (&r)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        sts128(buf_s + row * 128 + ((c ^ (row & 7)) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
}

// One 32-row x 32-column chunk of an accumulator tile, by one epilogue warp (lane = row): optional fused per-channel
// scale / shift / ReLU / TF32 rounding (FUSED: the inference path), swizzled staging, TMA store or reduce-add of the
// warp's box, optional per-column (sum, sum^2) of the staged values for the BN statistics.
template <bool FUSED, bool SPATIAL>
__device__ __forceinline__ void epilogue_chunk(uint32_t (&r)[32], const GemmParams& p, const CUtensorMap* dmap, uint8_t* buf,
                                               int lane, int nb, bool reduce_out, bool do_stats, uint32_t rowmask,
                                               int c1, int c2, int c3, float& ssum, float& ssq, uint32_t shift_s = 0) {
    if (FUSED) {
        // shift_s: this chunk's 32 per-channel shifts, staged in shared memory at kernel start (every lane needs all 32:
        // 8 broadcast LDS.128 instead of 32 global loads); the per-channel scale is normally folded into the packed weights
        if (p.scale) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * __ldg(p.scale + nb + j));
        }
        if (shift_s) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 sv;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w) : "r"(shift_s + c * 16));
                r[4 * c] = __float_as_uint(__uint_as_float(r[4 * c]) + sv.x);
                r[4 * c + 1] = __float_as_uint(__uint_as_float(r[4 * c + 1]) + sv.y);
                r[4 * c + 2] = __float_as_uint(__uint_as_float(r[4 * c + 2]) + sv.z);
                r[4 * c + 3] = __float_as_uint(__uint_as_float(r[4 * c + 3]) + sv.w);
            }
        }
        if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.f));
        }
        if (p.round_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(tf_round_tf32(__uint_as_float(r[j])));
        }
    }
    const uint32_t buf_s = smem_u32(buf);
    if (lane == 0) tma_store_wait_read<1>();         // the store that last used this buffer has drained
    __syncwarp();
    store_row_swizzled(buf_s, lane, r);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (reduce_out) {
            if (SPATIAL) tma_reduce_add_4d(dmap, buf, nb, c1, c2, c3);
            else         tma_reduce_add_2d(dmap, buf, nb, c1);
        } else {
            if (SPATIAL) tma_store_4d(dmap, buf, nb, c1, c2, c3);
            else         tma_store_2d(dmap, buf, nb, c1);
        }
        tma_store_commit();
    }
    if (do_stats) {
        // column sums of this warp's 32 rows: lane = channel, conflict-free reads of the staged rows, all issued before
        // the first use; rowmask drops rows outside the image (a 3x3 tap can pull in-image data into such a row)
        float v[32];
#pragma unroll
        for (int l = 0; l < 32; ++l)
            v[l] = lds32(buf_s + l * 128 + (((lane >> 2) ^ (l & 7)) << 4) + ((lane & 3) << 2));
        if (SPATIAL && rowmask != 0xffffffffu) {
#pragma unroll
            for (int l = 0; l < 32; ++l) if (!((rowmask >> l) & 1u)) v[l] = 0.f;
        }
        float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int l = 0; l < 32; l += 2) {
            s0 += v[l]; t0 = fmaf(v[l], v[l], t0);
            s1 += v[l + 1]; t1 = fmaf(v[l + 1], v[l + 1], t1);
        }
        ssum += s0 + s1; ssq += t0 + t1;
    }
}

// End of a CTA's epilogue: the four warps' per-column (sum, sum^2) partials are combined through the (by now idle) staging
// buffers into ONE row of the partial table per CTA -- the finalize kernel behind the GEMM then reads 4x fewer rows.
template <int BN>
__device__ __forceinline__ void write_stats_row(uint8_t* epi, int q, int lane, const float (&st_sum)[BN / 32],
                                                const float (&st_sq)[BN / 32], float* __restrict__ dst /* row: [2][cout] at n0 */,
                                                int cout) {
    if (lane == 0) tma_store_wait_read<0>();             // this warp's stores have left its staging buffers
    __syncwarp();
    const uint32_t red = smem_u32(epi);                  // warp w keeps [2][BN] floats at the start of ITS 8 KB staging area
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(red + q * 8192 + (c * 32 + lane) * 4), "f"(st_sum[c]) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(red + q * 8192 + (BN + c * 32 + lane) * 4), "f"(st_sq[c]) : "memory");
    }
    named_bar_sync_epi();                                // the 128 epilogue threads
    for (int col = q * 32 + lane; col < BN; col += 128) {
        float s = 0.f, t = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            s += lds32(red + w * 8192 + col * 4);
            t += lds32(red + w * 8192 + (BN + col) * 4);
        }
        dst[col] = s;
        dst[cout + col] = t;
    }
}

struct WorkUnit { int tile, k0, k1, split; };
__device__ __forceinline__ WorkUnit decode_unit(const GemmParams& p, int u, int kiters) {
    WorkUnit w;
    if (u < p.main_tiles) { w.tile = u; w.k0 = 0; w.k1 = kiters; w.split = 0; }
    else {
        const int t = u - p.main_tiles, s = t % p.ksplit;
        w.tile = p.main_tiles + t / p.ksplit;
        w.k0 = (int)((long long)kiters * s / p.ksplit);
        w.k1 = (int)((long long)kiters * (s + 1) / p.ksplit);
        w.split = 1;
    }
    return w;
}

template <int BN, bool FUSED, bool SPATIAL, bool RES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = RES ? Cfg::STAGES - Cfg::RES_DROP : Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint8_t* resbuf = smem + STAGES * Cfg::STAGE_BYTES;             // RES: 4 warps x 2 x 4 KB in the dropped stage(s)
    uint64_t* full = reinterpret_cast<uint64_t*>(epi + 2 * EPI_BUF_BYTES);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint64_t* rbar = tempty + 4;                                    // RES: [4 warps][2 buffers]
    float* shift_smem = reinterpret_cast<float*>(epi + 2 * EPI_BUF_BYTES + 1024);    // FUSED only (FUSED_SMEM_BYTES)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (FUSED && p.shift)                               // every CTA keeps one n-tile: its BN shifts, once
        for (int i = threadIdx.x; i < BN; i += GEMM_THREADS) shift_smem[i] = __ldg(p.shift + (blockIdx.x % p.num_n_tiles) * BN + i);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) { prefetch_tmap(&maps.a[s]); prefetch_tmap(&maps.b[s]); }
        prefetch_tmap(&maps.d);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
        if (RES) { prefetch_tmap(&maps.res); for (int i = 0; i < 8; ++i) mbar_init(&rbar[i], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.num_m_tiles * p.num_n_tiles;
    const int total_units = p.main_tiles + (total_tiles - p.main_tiles) * p.ksplit;
    const int kiters = p.nseg * p.taps * p.kblocks;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            int stage = 0; uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
                const WorkUnit wu = decode_unit(p, unit, kiters);
                const int n0 = (wu.tile % p.num_n_tiles) * BN;
                const int mt = wu.tile / p.num_n_tiles;
                int m0 = mt * BLOCK_M, img = 0, x0 = 0, y0 = 0;
                if (SPATIAL) {
                    img = mt / tiles_per_img;
                    const int r = mt % tiles_per_img;
                    y0 = (r / p.tiles_x) * p.th;
                    x0 = (r % p.tiles_x) * p.tw;
                }
                // K index -> (segment, tap, 32-channel block), kb fastest
                int kb = wu.k0 % p.kblocks, tap = (wu.k0 / p.kblocks) % p.taps, seg = wu.k0 / (p.kblocks * p.taps);
                for (int k = wu.k0; k < wu.k1; ++k) {
                    mbar_wait(&empty[stage], phase ^ 1, p.err_flag, 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    if (SPATIAL)
                        tma_load_4d(sa, &maps.a[seg], &full[stage], kb * BLOCK_K, p.stride * x0 + p.tap_ox[tap], p.stride * y0 + p.tap_oy[tap], img);
                    else
                        tma_load_2d(sa, &maps.a[seg], &full[stage], kb * BLOCK_K, m0);
                    tma_load_2d(sb, &maps.b[seg], &full[stage], (p.tap_w[tap] * p.kblocks + kb) * BLOCK_K, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++kb == p.kblocks) { kb = 0; if (++tap == p.taps) { tap = 0; ++seg; } }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BN, MAJOR_K, MAJOR_K);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
                const WorkUnit wu = decode_unit(p, unit, kiters);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1, p.err_flag, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int k = wu.k0; k < wu.k1; ++k) {
                    mbar_wait(&full[stage], phase, p.err_flag, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int j = 0; j < BLOCK_K / 8; ++j) {
                        const uint64_t ad = make_smem_desc(sa + j * 32, 16, 1024);
                        const uint64_t bd = make_smem_desc(sb + j * 32, 16, 1024);
                        mma_tf32(d_tmem, ad, bd, idesc, (k != wu.k0) || j != 0);
                    }
                    tc_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[acc]);
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..5)
        // Each warp owns TMEM lanes [32q, 32q+32) = 32 rows of the tile and works independently: tcgen05.ld of the
        // next 32-column chunk is in flight while the current one is scaled / staged (two private 4 KB swizzled
        // buffers) and stored by the warp's own TMA (box = 32 rows), so there is no block-level barrier in the loop.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        uint8_t* wbuf = epi + q * 2 * 4096;
        int wx = 0, wy = 0;                     // position of the warp's 32 rows inside a spatial patch
        if (SPATIAL) { wy = (q * 32) / p.tw; wx = (q * 32) % p.tw; }
        int it = 0, ebuf = 0;
        float st_sum[BN / 32], st_sq[BN / 32];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) { st_sum[c] = 0.f; st_sq[c] = 0.f; }
        uint32_t rphase[2] = {0, 0};
        if (RES && lane == 0 && (int)blockIdx.x < total_units) {    // residual boxes of the first unit's chunks 0 and 1
            const WorkUnit fw = decode_unit(p, blockIdx.x, kiters);
            const int fn0 = (fw.tile % p.num_n_tiles) * BN, fm0 = (fw.tile / p.num_n_tiles) * BLOCK_M;
            for (int c = 0; c < 2 && c < BN / 32; ++c) {
                mbar_expect_tx(&rbar[q * 2 + c], 4096);
                tma_load_2d(resbuf + (q * 2 + c) * 4096, &maps.res, &rbar[q * 2 + c], fn0 + c * 32, fm0 + q * 32);
            }
        }
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++it) {
            const WorkUnit wu = decode_unit(p, unit, kiters);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n0 = (wu.tile % p.num_n_tiles) * BN;
            const int mt = wu.tile / p.num_n_tiles;
            int m0 = mt * BLOCK_M, img = 0, x0 = 0, y0 = 0;
            if (SPATIAL) {
                img = mt / tiles_per_img;
                const int r = mt % tiles_per_img;
                y0 = (r / p.tiles_x) * p.th;
                x0 = (r % p.tiles_x) * p.tw;
            }
            const bool reduce_out = p.accumulate || wu.split;       // K slices meet in D through reduce-add
            const bool do_stats = p.stats_partial && !wu.split;     // (the host reduces the split rows separately)
            uint32_t rowmask = 0xffffffffu;                         // bit l: row l of this warp is a pixel of the image
            if (SPATIAL && do_stats) {
                const int rw = q * 32 + lane;
                rowmask = __ballot_sync(0xffffffffu, x0 + rw % p.tw < p.img_w && y0 + rw / p.tw < p.img_h);
            }
            // RES: the residual boxes of chunks 0 and 1 of this unit were requested at the end of the previous unit (or
            // before the loop); chunk c's buffer is refilled with chunk c+2 (next unit's 0 / 1 at the end) right after use
            const long long res_row = (long long)m0 + q * 32 + lane;
            const bool res_valid = RES && res_row < p.m_rows;
            int nn0 = 0, nm0 = 0; bool has_next = false;            // first chunk coordinates of this CTA's next unit
            if (RES && unit + (int)gridDim.x < total_units) {
                const WorkUnit nw = decode_unit(p, unit + gridDim.x, kiters);
                nn0 = (nw.tile % p.num_n_tiles) * BN; nm0 = (nw.tile / p.num_n_tiles) * BLOCK_M; has_next = true;
            }
            mbar_wait(&tfull[acc], acc_phase, p.err_flag, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            uint32_t rr[2][32];
            tmem_ld_32x32(taddr, rr[0]);
#pragma unroll
            for (int chunk = 0; chunk < BN / 32; ++chunk) {
                uint32_t (&r)[32] = rr[chunk & 1];
                tmem_ld_wait();                                              // this chunk's registers are valid
                if (chunk + 1 < BN / 32) tmem_ld_32x32(taddr + (chunk + 1) * 32, rr[(chunk + 1) & 1]);
                else { tc_fence_before(); mbar_arrive(&tempty[acc]); }        // accumulator stage fully read
                if (RES) {
                    constexpr int NCH = BN / 32;
                    const int slot = chunk & 1;
                    uint32_t m = 0xffffffffu;
                    if (p.res_mask) m = res_valid ? __ldg(p.res_mask + ((res_row * p.cout + n0 + chunk * 32) >> 5)) : 0u;
                    mbar_wait(&rbar[q * 2 + slot], rphase[slot], p.err_flag, 5);
                    rphase[slot] ^= 1;
                    const uint32_t rb = smem_u32(resbuf + (q * 2 + slot) * 4096) + lane * 128;
                    uint4 rs[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rs[c].x), "=r"(rs[c].y), "=r"(rs[c].z), "=r"(rs[c].w)
                                     : "r"(rb + ((c ^ (lane & 7)) << 4)) : "memory");
                    fence_proxy_async();                                     // generic reads before the async-proxy refill
                    __syncwarp();                                            // every lane has read the buffer: refill it
                    if (lane == 0) {
                        if (chunk + 2 < NCH) {
                            mbar_expect_tx(&rbar[q * 2 + slot], 4096);
                            tma_load_2d(resbuf + (q * 2 + slot) * 4096, &maps.res, &rbar[q * 2 + slot], n0 + (chunk + 2) * 32, m0 + q * 32);
                        } else if (has_next && chunk + 2 - NCH < NCH) {
                            mbar_expect_tx(&rbar[q * 2 + slot], 4096);
                            tma_load_2d(resbuf + (q * 2 + slot) * 4096, &maps.res, &rbar[q * 2 + slot], nn0 + (chunk + 2 - NCH) * 32, nm0 + q * 32);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 v = rs[c];
                        if (m & (1u << (4 * c)))     r[4 * c]     = __float_as_uint(__uint_as_float(r[4 * c])     + __uint_as_float(v.x));
                        if (m & (2u << (4 * c)))     r[4 * c + 1] = __float_as_uint(__uint_as_float(r[4 * c + 1]) + __uint_as_float(v.y));
                        if (m & (4u << (4 * c)))     r[4 * c + 2] = __float_as_uint(__uint_as_float(r[4 * c + 2]) + __uint_as_float(v.z));
                        if (m & (8u << (4 * c)))     r[4 * c + 3] = __float_as_uint(__uint_as_float(r[4 * c + 3]) + __uint_as_float(v.w));
                    }
                }
                uint8_t* buf = wbuf + ebuf * 4096;
                epilogue_chunk<FUSED, SPATIAL>(r, p, &maps.d, buf, lane, n0 + chunk * 32, reduce_out, do_stats, rowmask,
                                               SPATIAL ? x0 + wx : m0 + q * 32, y0 + wy, img, st_sum[chunk], st_sq[chunk],
                                               (FUSED && p.shift) ? smem_u32(shift_smem) + chunk * 128 : 0u);
                ebuf ^= 1;
            }
        }
        if (p.stats_partial) {
            const int n0 = (blockIdx.x % p.num_n_tiles) * BN;
            write_stats_row<BN>(epi, q, lane, st_sum, st_sq, p.stats_partial + (size_t)(blockIdx.x / p.num_n_tiles) * 2 * p.cout + n0, p.cout);
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    __syncthreads();
    if (warp == 1) { __syncwarp(); tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
}

// ------------------------------------------------------------------------------------------------ 2-CTA fprop/dgrad
// conv_gemm2_kernel: the same implicit GEMM with tcgen05.mma.cta_group::2 -- a cluster of two CTAs (one TPC) computes a
// 256-pixel x 256-channel tile: each CTA stages ITS 128 pixel rows of A and HALF (128 channels) of the B tile, the
// leader's single MMA thread issues M=256 x N=256 x K=8 instructions that read both CTAs' shared memory, and each CTA
// owns the accumulator rows of its pixels in its own TMEM.  Per CTA the operand traffic drops from 48 KB to 32 KB per
// K=32 step (the one-CTA kernel is shared-memory-bandwidth bound on the K=256 layers), 6 pipeline stages fit.
// Barrier protocol: full[s] lives in the leader (count 1 = the leader producer's expect_tx of BOTH CTAs' bytes; both
// producers' TMA complete_tx on it); empty[s] / tfull[a] exist in both CTAs and are signalled by multicast commits;
// tempty[a] lives in the leader and collects the 256 epilogue threads of both CTAs.
constexpr int G2_BN = 256;
constexpr int G2_STAGES = 6;
constexpr int G2_B_STAGE_BYTES = 128 * 128;                        // this CTA's half of the B tile
constexpr int G2_STAGE_BYTES = A_STAGE_BYTES + G2_B_STAGE_BYTES;   // 32 KB
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 2 * EPI_BUF_BYTES + 1024 + 1024;
constexpr int G2_FUSED_SMEM_BYTES = G2_SMEM_BYTES + 1024;

template <bool FUSED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
conv_gemm2_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
    constexpr int BN = G2_BN;
    constexpr int STAGES = G2_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi = smem + STAGES * G2_STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(epi + 2 * EPI_BUF_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* shift_smem = reinterpret_cast<float*>(epi + 2 * EPI_BUF_BYTES + 1024);    // FUSED only

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs)
    if (FUSED && p.shift)
        for (int i = threadIdx.x; i < BN; i += GEMM_THREADS) shift_smem[i] = __ldg(p.shift + ((blockIdx.x >> 1) % p.num_n_tiles) * BN + i);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) { prefetch_tmap(&maps.a[s]); prefetch_tmap(&maps.b[s]); }
        prefetch_tmap(&maps.d);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 256); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2<2 * BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // the peer's barriers are initialised too
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int pair_tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;     // a "tile" = two consecutive m-tiles x BN
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer (both CTAs)
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < pair_tiles; tile += num_clusters) {
                const int n0 = (tile % p.num_n_tiles) * BN;
                const int mt = 2 * (tile / p.num_n_tiles) + (int)rank;
                int m0 = mt * BLOCK_M, img = 0, x0 = 0, y0 = 0;
                if (p.spatial) {
                    img = mt / tiles_per_img;                 // an odd tail m-tile lands in image index B: fully OOB -> zeros
                    const int r = mt % tiles_per_img;
                    y0 = (r / p.tiles_x) * p.th;
                    x0 = (r % p.tiles_x) * p.tw;
                }
                for (int seg = 0; seg < p.nseg; ++seg)
                    for (int tap = 0; tap < p.taps; ++tap)
                        for (int kb = 0; kb < p.kblocks; ++kb) {
                            mbar_wait(&empty[stage], phase ^ 1, p.err_flag, 21);
                            uint8_t* sa = smem + stage * G2_STAGE_BYTES;
                            uint8_t* sb = sa + A_STAGE_BYTES;
                            const uint32_t lead_full = map_to_cta(smem_u32(&full[stage]), 0);
                            if (rank == 0) mbar_expect_tx(&full[stage], 2 * G2_STAGE_BYTES);
                            if (p.spatial)
                                tma2_load_4d(sa, &maps.a[seg], lead_full, kb * BLOCK_K, p.stride * x0 + p.tap_ox[tap], p.stride * y0 + p.tap_oy[tap], img);
                            else
                                tma2_load_2d(sa, &maps.a[seg], lead_full, kb * BLOCK_K, m0);
                            tma2_load_2d(sb, &maps.b[seg], lead_full, (p.tap_w[tap] * p.kblocks + kb) * BLOCK_K, n0 + (int)rank * 128);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ------------------------------------------------------------ MMA issuer (leader CTA only)
            constexpr uint32_t idesc = make_idesc_tf32(256, BN, MAJOR_K, MAJOR_K);
            const int kiters = p.nseg * p.taps * p.kblocks;
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < pair_tiles; tile += num_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1, p.err_flag, 22);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int k = 0; k < kiters; ++k) {
                    mbar_wait(&full[stage], phase, p.err_flag, 23);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int j = 0; j < BLOCK_K / 8; ++j) {
                        const uint64_t ad = make_smem_desc(sa + j * 32, 16, 1024);
                        const uint64_t bd = make_smem_desc(sb + j * 32, 16, 1024);
                        mma2_tf32(d_tmem, ad, bd, idesc, (k | j) != 0);
                    }
                    tc_commit2(&empty[stage]);               // frees the stage in BOTH CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit2(&tfull[acc]);                     // accumulators ready in BOTH CTAs
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..5, both CTAs)
        const int q = warp & 3;
        uint8_t* wbuf = epi + q * 2 * 4096;
        int wx = 0, wy = 0;
        if (p.spatial) { wy = (q * 32) / p.tw; wx = (q * 32) % p.tw; }
        int it = 0, ebuf = 0;
        float st_sum[BN / 32], st_sq[BN / 32];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) { st_sum[c] = 0.f; st_sq[c] = 0.f; }
        for (int tile = cluster_id; tile < pair_tiles; tile += num_clusters, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n0 = (tile % p.num_n_tiles) * BN;
            const int mt = 2 * (tile / p.num_n_tiles) + (int)rank;
            int m0 = mt * BLOCK_M, img = 0, x0 = 0, y0 = 0;
            if (p.spatial) {
                img = mt / tiles_per_img;
                const int r = mt % tiles_per_img;
                y0 = (r / p.tiles_x) * p.th;
                x0 = (r % p.tiles_x) * p.tw;
            }
            const bool real_tile = mt < p.num_m_tiles;     // the odd tail pair has one phantom half
            uint32_t rowmask = 0xffffffffu;
            if (p.spatial && p.stats_partial) {
                const int rw = q * 32 + lane;
                rowmask = __ballot_sync(0xffffffffu, x0 + rw % p.tw < p.img_w && y0 + rw / p.tw < p.img_h);
            }
            mbar_wait(&tfull[acc], acc_phase, p.err_flag, 24);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            uint32_t rr[2][32];
            tmem_ld_32x32(taddr, rr[0]);
#pragma unroll
            for (int chunk = 0; chunk < BN / 32; ++chunk) {
                uint32_t (&r)[32] = rr[chunk & 1];
                tmem_ld_wait();
                if (chunk + 1 < BN / 32) tmem_ld_32x32(taddr + (chunk + 1) * 32, rr[(chunk + 1) & 1]);
                else { tc_fence_before(); mbar_arrive_cluster(map_to_cta(smem_u32(&tempty[acc]), 0)); }
                if (!real_tile) continue;
                uint8_t* buf = wbuf + ebuf * 4096;
                const uint32_t shs = (FUSED && p.shift) ? smem_u32(shift_smem) + chunk * 128 : 0u;
                if (p.spatial) epilogue_chunk<FUSED, true>(r, p, &maps.d, buf, lane, n0 + chunk * 32, p.accumulate != 0, p.stats_partial != nullptr, rowmask,
                                                          x0 + wx, y0 + wy, img, st_sum[chunk], st_sq[chunk], shs);
                else           epilogue_chunk<FUSED, false>(r, p, &maps.d, buf, lane, n0 + chunk * 32, p.accumulate != 0, p.stats_partial != nullptr, rowmask,
                                                           m0 + q * 32, 0, 0, st_sum[chunk], st_sq[chunk], shs);
                ebuf ^= 1;
            }
        }
        if (p.stats_partial) {
            // every cluster keeps one n-tile (num_clusters % num_n_tiles == 0): row = (cluster / n_tiles) * 2 + rank
            const int n0 = (cluster_id % p.num_n_tiles) * BN;
            write_stats_row<BN>(epi, q, lane, st_sum, st_sq,
                                p.stats_partial + (size_t)((cluster_id / p.num_n_tiles) * 2 + (int)rank) * 2 * p.cout + n0, p.cout);
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // neither CTA may exit (or free TMEM) while the peer can still signal it
    if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc2<2 * BN>(tmem_base); }
}

// ------------------------------------------------------------------------------------------------ wgrad
struct WgradMaps { CUtensorMap a[MAX_SEG], b[MAX_SEG], d; };   // a: dY (Cout, pixels), b: X (Cin, pixels), d: dW [Cout, taps*Cin]
struct WgradParams {
    int m_tiles, n_tiles, taps, splits, nseg;
    int spatial, tiles_x, tiles_y, tw, th;     // pixel blocks of 32: flat or (tw x th) patches of one image
    int num_pblocks;                            // total pixel blocks
    int cin, cout;                              // dW is [cout][taps*cin]; column = tap * cin + ci
    int b_groups;                               // 32-channel groups fetched per B-operand TMA (= min(BN, cin) / 32)
    int stride;                                 // X coordinate = stride * dY coordinate + tap offset
    int plain_store;                            // debug: overwrite instead of reduce-add (needs splits == 1)
    int* err_flag;
};

// MT = number of 128-row accumulator tiles a CTA keeps in TMEM (MT * BN columns).  MT = 2: the SAME B (activation) tile of
// a pipeline stage is multiplied with two dY tiles (Cout rows m0 .. m0+255), so a stage of 64 KB feeds twice the MMAs of
// the 48 KB stage of MT = 1: 65 instead of 44 flop per byte brought into shared memory.  The kernel is L2->SM fill bound
// (ncu, round 1: 48-56 % tensor pipe), so that ratio is what sets its speed.
template <int BN, int MT> struct WgradCfg {
    static constexpr int A_BYTES = MT * A_STAGE_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + BN * 128;
    static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*barriers*/ + 1024 /*align*/;      // epilogue overlays stage 0
    static constexpr int TMEM_COLS = MT * BN < 32 ? 32 : MT * BN;
    static_assert(2 * EPI_BUF_BYTES <= STAGES * STAGE_BYTES, "wgrad epilogue buffers must fit in the operand ring");
    static_assert(STAGES >= 3, "wgrad pipeline too shallow");
};

template <int BN, int MT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ WgradMaps maps, const WgradParams p) {
    using Cfg = WgradCfg<BN, MT>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // The epilogue staging buffers overlay the first operand stage: the epilogue only starts once every MMA (hence every
    // operand read and every TMA load) has completed.  That keeps the CTA at ~194 KB of shared memory, so ~30 KB stay free
    // for the chain's elementwise blocks (1 KB reserved each) this kernel is meant to run next to.
    uint8_t* epi = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) { prefetch_tmap(&maps.a[s]); prefetch_tmap(&maps.b[s]); }
        prefetch_tmap(&maps.d);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&tfull[0], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // GEMM view: D[Cout, taps*Cin] -- a column block of BN may span several taps (Cin < BN) or part of one
    const int units = p.m_tiles * p.n_tiles;
    const int unit = blockIdx.x % units, split = blockIdx.x / units;
    const int n0 = (unit % p.n_tiles) * BN;
    const int m0 = (unit / p.n_tiles) * (BLOCK_M * MT);
    const int ncols = p.taps * p.cin;
    const int b_boxes = min(BN / 32, (ncols - n0 + 31) / 32);           // 32-column chunks of this tile that exist in dW
    const int pb0 = (int)((long long)p.num_pblocks * split / p.splits);
    const int pb1 = (int)((long long)p.num_pblocks * (split + 1) / p.splits);
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (pb1 > pb0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int pb = pb0; pb < pb1; ++pb)
                for (int seg = 0; seg < p.nseg; ++seg) {
                    mbar_wait(&empty[stage], phase ^ 1, p.err_flag, 11);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    // One TMA per operand (per tap for B): the channel axis is split into (32, C/32) so that a box
                    // (32 ch, 32 px, g groups) lands as g consecutive [32 px][128 B] slabs -- the MN-major layout the
                    // MMA wants -- instead of g separate 4 KB copies.  Groups past the tensor edge are zero-filled.
                    // (MT = 2: the A box holds 8 groups = both 128-row dY tiles.)
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    const int gb = p.b_groups;                       // 32-channel groups per B instruction
                    if (p.spatial) {
                        const int img = pb / tiles_per_img, r = pb % tiles_per_img;
                        const int y0 = (r / p.tiles_x) * p.th, x0 = (r % p.tiles_x) * p.tw;
                        tma_load_5d(sa, &maps.a[seg], &full[stage], 0, x0, y0, img, m0 / 32);
                        for (int j = 0; j < BN / 32; j += gb) {
                            const int col = n0 + j * 32;
                            const int tap = col < ncols ? col / p.cin : 0;
                            const int grp = col < ncols ? (col - tap * p.cin) / 32 : p.cin / 32;    // past the end: all zero
                            tma_load_5d(sb + j * 4096, &maps.b[seg], &full[stage], 0, p.stride * x0 + tap_dx(p.taps, tap), p.stride * y0 + tap_dy(p.taps, tap), img, grp);
                        }
                    } else {
                        tma_load_3d(sa, &maps.a[seg], &full[stage], 0, pb * 32, m0 / 32);
                        for (int j = 0; j < BN / 32; j += gb) {
                            const int col = n0 + j * 32;
                            tma_load_3d(sb + j * 4096, &maps.b[seg], &full[stage], 0, pb * 32, col < ncols ? col / 32 : p.cin / 32);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BN, MAJOR_MN, MAJOR_MN);
                int stage = 0; uint32_t phase = 0;
                for (int pb = pb0; pb < pb1; ++pb)
                for (int seg = 0; seg < p.nseg; ++seg) {
                    mbar_wait(&full[stage], phase, p.err_flag, 13);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {       // 32 pixels = 4 MMAs of K = 8
                        // MN-major tf32: 32 channels (128 B) x 32 pixel rows per TMA box, SWIZZLE_128B_BASE32B atoms of
                        // 4 rows: LBO = next 32-channel box (4096 B), SBO = next 4-row group (512 B), K step = 8 rows
                        const uint64_t bd = make_smem_desc(sb + j * 1024, 4096, 512, LAYOUT_SW128_BASE32B);
#pragma unroll
                        for (int h = 0; h < MT; ++h) {
                            const uint64_t ad = make_smem_desc(sa + h * A_STAGE_BYTES + j * 1024, 4096, 512, LAYOUT_SW128_BASE32B);
                            mma_tf32(tmem_base + h * BN, ad, bd, idesc, (pb != pb0) || seg != 0 || j != 0);
                        }
                    }
                    tc_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[0]);
            }
        } else {
            const int q = warp & 3;
            const int row = q * 32 + lane;
            const bool store_thread = threadIdx.x == 64;
            mbar_wait(&tfull[0], 0, p.err_flag, 14);
            tc_fence_after();
            int ebuf = 0;
#pragma unroll 1
            for (int h = 0; h < MT; ++h) {
                if (m0 + h * BLOCK_M >= p.cout) break;                 // (a dY tile past the tensor edge holds zeros)
#pragma unroll 1
                for (int chunk = 0; chunk < b_boxes; ++chunk) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + h * BN + chunk * 32, r);
                    tmem_ld_wait();
                    uint8_t* buf = epi + ebuf * EPI_BUF_BYTES;
                    if (store_thread) tma_store_wait_read<1>();
                    named_bar_sync_epi();
                    store_row_swizzled(smem_u32(buf), row, r);
                    fence_proxy_async();
                    named_bar_sync_epi();
                    if (store_thread) {
                        if (p.plain_store) tma_store_2d(&maps.d, buf, n0 + chunk * 32, m0 + h * BLOCK_M);
                        else               tma_reduce_add_2d(&maps.d, buf, n0 + chunk * 32, m0 + h * BLOCK_M);
                        tma_store_commit();
                    }
                    ebuf ^= 1;
                }
            }
            if (store_thread) tma_store_wait_all<0>();
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    });
    return fn;
}

// 2-D fp32 tensor [rows, cols] (cols contiguous, row pitch = pitch_elems), box (box_cols<=32, box_rows)
int encode_2d(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t pitch_elems, uint32_t box_cols,
              uint32_t box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { tf_set_error("cuTensorMapEncodeTiled entry point unavailable"); return TF_ERR_CUDA; }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tf_set_error("cuTensorMapEncodeTiled(2d cols=%llu rows=%llu pitch=%llu box=%u,%u) failed: %d",
                                          (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_elems, box_cols, box_rows, (int)r); return TF_ERR_CUDA; }
    return TF_OK;
}
// 4-D NHWC fp32 tensor viewed as (C, W, H, B), box (bc<=32, bw, bh, 1)
int encode_4d(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t bc, uint32_t bw,
              uint32_t bh, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint32_t estride = 1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { tf_set_error("cuTensorMapEncodeTiled entry point unavailable"); return TF_ERR_CUDA; }
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 4, W * C * 4, H * W * C * 4};
    // with a traversal stride s the box spans s*n coordinates and loads every s-th element (n of them)
    cuuint32_t box[4] = {bc, bw * estride, bh * estride, 1};
    cuuint32_t es[4] = {1, estride, estride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tf_set_error("cuTensorMapEncodeTiled(4d C=%llu W=%llu H=%llu B=%llu box=%u,%u,%u) failed: %d",
                                          (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B, bc, bw, bh, (int)r); return TF_ERR_CUDA; }
    return TF_OK;
}

// 4-D view (C, Wq, Hq, B) of every second pixel in x and y of an NHWC tensor [B,H,W,C], starting at pixel (py, px)
int encode_4d_lattice2(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t B, int py, int px,
                       uint32_t bc, uint32_t bw, uint32_t bh) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { tf_set_error("cuTensorMapEncodeTiled entry point unavailable"); return TF_ERR_CUDA; }
    const uint64_t Wq = (W - px + 1) / 2, Hq = (H - py + 1) / 2;
    cuuint64_t dims[4] = {C, Wq, Hq, B};
    cuuint64_t strides[3] = {2 * C * 4, 2 * W * C * 4, H * W * C * 4};
    cuuint32_t box[4] = {bc, bw, bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const char* base = reinterpret_cast<const char*>(ptr) + ((size_t)py * W + px) * C * 4;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<char*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tf_set_error("cuTensorMapEncodeTiled(4d lattice C=%llu W=%llu H=%llu B=%llu) failed: %d",
                                          (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B, (int)r); return TF_ERR_CUDA; }
    return TF_OK;
}

// MN-major operand views for wgrad: channels split as (32, C/32) with the group axis outermost in the box
int encode_3d_grouped(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t rows, uint32_t box_rows, uint32_t groups) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { tf_set_error("cuTensorMapEncodeTiled entry point unavailable"); return TF_ERR_CUDA; }
    cuuint64_t dims[3] = {32, rows, C / 32};
    cuuint64_t strides[2] = {C * 4, 128};
    cuuint32_t box[3] = {32, box_rows, groups};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tf_set_error("cuTensorMapEncodeTiled(3d grouped C=%llu rows=%llu box=%u,%u) failed: %d",
                                          (unsigned long long)C, (unsigned long long)rows, box_rows, groups, (int)r); return TF_ERR_CUDA; }
    return TF_OK;
}
int encode_5d_grouped(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t bw, uint32_t bh,
                      uint32_t groups, uint32_t estride = 1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { tf_set_error("cuTensorMapEncodeTiled entry point unavailable"); return TF_ERR_CUDA; }
    cuuint64_t dims[5] = {32, W, H, B, C / 32};
    cuuint64_t strides[4] = {C * 4, W * C * 4, H * W * C * 4, 128};
    cuuint32_t box[5] = {32, bw * estride, bh * estride, 1, groups};
    cuuint32_t es[5] = {1, estride, estride, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tf_set_error("cuTensorMapEncodeTiled(5d grouped C=%llu W=%llu H=%llu B=%llu box=%u,%u,%u) failed: %d",
                                          (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B, bw, bh, groups, (int)r); return TF_ERR_CUDA; }
    return TF_OK;
}

void pick_tile(int W, int H, int area, int* tw, int* th) {
    long long best = -1;
    for (int w = area; w >= 1; w >>= 1) {
        const int h = area / w;
        if (w > 256 || h > 256) continue;
        const long long cover = (long long)((W + w - 1) / w) * w * ((H + h - 1) / h) * h;
        if (best < 0 || cover < best || (cover == best && w > *tw && w <= 32)) { best = cover; *tw = w; *th = h; }
    }
}

// ---- launch plan of one convolution GEMM: pure host arithmetic (no CUDA calls), exported as tf_conv_plan so that the
//      CPU test-suite can check its invariants (tests/test_cabi_cpu.py)
struct ConvPlan {
    int spatial, tw, th, tiles_x, tiles_y, m_tiles;   // m-tiles: flat rows of 128 pixels or (tw x th) patches of one image
    int bn, n_tiles, grid;                            // tile width, n-tiles, persistent CTAs (a multiple of n_tiles)
    int two_cta;                                      // cta_group::2 kernel (256 x 256 tiles over CTA pairs)
    int main_tiles, ksplit;                           // tail-wave split-K: tiles >= main_tiles are cut into ksplit K slices
    long long tail_pix0;                              // first output pixel of the split region (-1: none); contiguous to the end
    int kiters;
};
ConvPlan plan_conv(int B, int H, int W, int Cin, int Cout, int ksize, int stride, int nseg, bool plain, bool has_res,
                   int num_sms, int force_bn, int two_cta_switch /*0 auto, 1 on, 2 off*/, bool allow_split) {
    ConvPlan c = {};
    const int taps = ksize * ksize;
    const int Ho = stride == 2 ? (H + 1) / 2 : H, Wo = stride == 2 ? (W + 1) / 2 : W;
    c.spatial = ksize == 3 || stride == 2;
    const long long Mtot = (long long)B * Ho * Wo;
    if (!c.spatial) { c.m_tiles = (int)((Mtot + BLOCK_M - 1) / BLOCK_M); c.tiles_x = c.tiles_y = 1; c.tw = 128; c.th = 1; }
    else {
        pick_tile(Wo, Ho, BLOCK_M, &c.tw, &c.th);
        c.tiles_x = (Wo + c.tw - 1) / c.tw; c.tiles_y = (Ho + c.th - 1) / c.th;
        c.m_tiles = B * c.tiles_x * c.tiles_y;
    }
    // tile width: fewest (rounds over the SMs) x (time per tile ~ BN); narrower tiles pay more smem bandwidth per MMA
    c.bn = 64; double best = 1e30;
    for (int cand = 256; cand >= 64; cand >>= 1) {
        if (Cout % cand) continue;
        const long long tiles = (long long)c.m_tiles * (Cout / cand);
        const double rounds = (double)((tiles + num_sms - 1) / num_sms);
        const double cost = rounds * cand * (cand == 256 ? 1.0 : (cand == 128 ? 1.45 : 2.2));   // measured: N=128 tiles are smem-bandwidth bound
        if (cost < best) { best = cost; c.bn = cand; }
    }
    if (force_bn) c.bn = force_bn;
    c.n_tiles = Cout / c.bn;
    c.kiters = nseg * taps * (Cin / 32);
    // 2-CTA (cta_group::2, 256 x 256 tiles): measured on B200 at batch-8 960x1280 -- K = 1024 1x1 (N = 256): 48.1 vs 52.8 us
    // (49.4 vs 53.1 with the statistics epilogue); K = 256 -> N = 1024: 52.4 vs 53.8 (noise); 3x3: 77 vs 69 (the one-CTA
    // kernel has the tail split-K).  So: flat GEMMs with a long K loop and a single 256-wide n-tile.
    const bool two_cta_auto = !c.spatial && Cout == 256 && Cin >= 512 && nseg == 1;
    c.two_cta = c.bn == 256 && two_cta_switch != 2 && (two_cta_switch == 1 || two_cta_auto) && c.m_tiles >= 2 && !has_res;
    const int tiles = c.m_tiles * c.n_tiles;
    c.grid = (tiles < num_sms ? tiles : num_sms) / c.n_tiles * c.n_tiles;
    c.main_tiles = tiles; c.ksplit = 1; c.tail_pix0 = -1;
    if (c.two_cta) return c;
    // ---- tail-wave split-K: the m-tiles that do not fill a whole round over the CTAs are cut along K.  Only for a plain
    //      epilogue (K slices cannot be scaled / clamped separately) and the 3x3 GEMMs (measured: for a K = 1024 1x1 the
    //      two extra launches -- memset of the split rows, their BN statistics -- eat the 4 us).
    if (plain && allow_split && taps == 9 && c.kiters >= 16 && c.grid > 0 && tiles > c.grid && tiles % c.grid != 0) {
        const int full = tiles / c.grid;
        int main_m = (int)((long long)full * c.grid / c.n_tiles);
        if (c.spatial) main_m -= main_m % c.tiles_x;      // the split region starts at a tile-row boundary
        const int tail_tiles = (c.m_tiles - main_m) * c.n_tiles;
        int ks = tail_tiles > 0 ? c.grid / tail_tiles : 0;
        if (ks > c.kiters / 8) ks = c.kiters / 8;
        // rounds: ceil(main / grid) whole tiles + one round of 1/ks tiles (+ ~0.2 of a tile for the extra epilogue / launches)
        const int main_rounds = (main_m * c.n_tiles + c.grid - 1) / c.grid;
        if (ks >= 2 && main_m > 0 && main_rounds + 1.0 / ks + 0.2 < 0.92 * (double)((tiles + c.grid - 1) / c.grid)) {     // worth >= 8 %
            c.main_tiles = main_m * c.n_tiles; c.ksplit = ks;
            if (c.spatial) {
                const int tpi = c.tiles_x * c.tiles_y, img = main_m / tpi, ty = (main_m % tpi) / c.tiles_x;
                c.tail_pix0 = ((long long)img * Ho + (long long)ty * c.th) * Wo;
            } else c.tail_pix0 = (long long)main_m * BLOCK_M;
        }
    }
    return c;
}
// taps of parity class (py, px) of a stride-2 dgrad, in the stride-1 numbering t = (dy'+1)*3 + (dx'+1) of the flipped
// packed weights: dy' = 0 for even rows (dY offset 0); dy' = -1 (offset 0) and +1 (offset +1) for odd rows; likewise in x
int s2_class_taps(int ksize, int py, int px, int* tap_w, int* tap_ox, int* tap_oy) {
    if (ksize == 1) { if (py || px) return 0; tap_w[0] = 0; tap_ox[0] = 0; tap_oy[0] = 0; return 1; }
    int nt = 0;
    const int ny = py ? 2 : 1, nx = px ? 2 : 1;
    for (int iy = 0; iy < ny; ++iy)
        for (int ix = 0; ix < nx; ++ix) {
            const int dyp = py ? (iy ? 1 : -1) : 0, dxp = px ? (ix ? 1 : -1) : 0;
            tap_oy[nt] = dyp == 1 ? 1 : 0;
            tap_ox[nt] = dxp == 1 ? 1 : 0;
            tap_w[nt] = (dyp + 1) * 3 + (dxp + 1);
            ++nt;
        }
    return nt;
}

// per-device state (a process may drive several GPUs; the SM count and the pipeline-error flag belong to the
// device that is current when a GEMM is enqueued)
constexpr int MAX_DEVICES = 64;
int g_num_sms_dev[MAX_DEVICES] = {0};
int* g_err_flag_dev[MAX_DEVICES] = {nullptr};
thread_local int g_num_sms = 0;
thread_local int* g_err_flag = nullptr;
std::mutex g_dev_mutex;
int ensure_device_state() {
    int dev = 0;
    TF_CHECK_CUDA(cudaGetDevice(&dev));
    TF_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "device index %d out of range", dev);
    if (g_num_sms_dev[dev] == 0) {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (g_num_sms_dev[dev] == 0) {
            int sms = 0;
            TF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            TF_CHECK_CUDA(cudaMalloc(&g_err_flag_dev[dev], sizeof(int)));
            TF_CHECK_CUDA(cudaMemset(g_err_flag_dev[dev], 0, sizeof(int)));
            g_num_sms_dev[dev] = sms;
        }
    }
    g_num_sms = g_num_sms_dev[dev];
    g_err_flag = g_err_flag_dev[dev];
    return TF_OK;
}

void default_taps(GemmParams& p, int taps) {
    p.taps = taps;
    for (int t = 0; t < 9; ++t) {
        p.tap_ox[t] = (signed char)(taps == 9 ? t % 3 - 1 : 0);
        p.tap_oy[t] = (signed char)(taps == 9 ? t / 3 - 1 : 0);
        p.tap_w[t] = (unsigned char)t;
    }
}

template <int BN, bool FUSED, bool SPATIAL, bool RES>
int launch_gemm_variant(const GemmMaps& maps, const GemmParams& p, int grid, cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        TF_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, FUSED, SPATIAL, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           FUSED ? Cfg::FUSED_SMEM_BYTES : Cfg::SMEM_BYTES));
        attr_set = true;
    }
    conv_gemm_kernel<BN, FUSED, SPATIAL, RES><<<grid, GEMM_THREADS, FUSED ? Cfg::FUSED_SMEM_BYTES : Cfg::SMEM_BYTES, st>>>(maps, p);   // grid % num_n_tiles == 0: one n-tile per CTA (stats_partial)
    TF_LAUNCH_CHECK();
    return TF_OK;
}
// FUSED (per-channel scale / shift / ReLU / rounding in the epilogue), SPATIAL (4-D patch tiles vs flat pixel rows) and
// RES (masked residual added in the epilogue) are separate instantiations: the training path's plain epilogue stays
// small enough for the instruction cache
template <int BN>
int launch_gemm(const GemmMaps& maps, const GemmParams& p, int grid, cudaStream_t st) {
    const bool fused = p.scale || p.shift || p.relu || p.round_out;
    if (p.res) return fused ? launch_gemm_variant<BN, true, false, true>(maps, p, grid, st) : launch_gemm_variant<BN, false, false, true>(maps, p, grid, st);
    if (p.spatial) return fused ? launch_gemm_variant<BN, true, true, false>(maps, p, grid, st) : launch_gemm_variant<BN, false, true, false>(maps, p, grid, st);
    return fused ? launch_gemm_variant<BN, true, false, false>(maps, p, grid, st) : launch_gemm_variant<BN, false, false, false>(maps, p, grid, st);
}
template <bool FUSED>
int launch_gemm2_variant(const GemmMaps& maps, const GemmParams& p, int clusters, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        TF_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED ? G2_FUSED_SMEM_BYTES : G2_SMEM_BYTES));
        attr_set = true;
    }
    conv_gemm2_kernel<FUSED><<<2 * clusters, GEMM_THREADS, FUSED ? G2_FUSED_SMEM_BYTES : G2_SMEM_BYTES, st>>>(maps, p);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int launch_gemm2(const GemmMaps& maps, const GemmParams& p, int* stats_rows, cudaStream_t st) {
    const int pair_tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    int clusters = pair_tiles < g_num_sms / 2 ? pair_tiles : g_num_sms / 2;
    clusters = clusters / p.num_n_tiles * p.num_n_tiles;                 // every cluster keeps one n-tile
    if (stats_rows) *stats_rows = clusters / p.num_n_tiles * 2;
    if (p.scale || p.shift || p.relu || p.round_out) return launch_gemm2_variant<true>(maps, p, clusters, st);
    return launch_gemm2_variant<false>(maps, p, clusters, st);
}
template <int BN, int MT>
int launch_wgrad(const WgradMaps& maps, const WgradParams& p, cudaStream_t st) {
    using Cfg = WgradCfg<BN, MT>;
    static bool attr_set = false;
    if (!attr_set) {
        TF_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int grid = p.m_tiles * p.n_tiles * p.splits;
    conv_wgrad_kernel<BN, MT><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(maps, p);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

int g_debug[16] = {0};
int g_debug_epoch = 0;              // bumped by every tf_debug_set: cached workspace plans made under other switches are stale

}  // namespace

TF_API int tf_debug_set(int key, int value) {
    TF_REQUIRE(key >= 0 && key < 16, "tf_debug_set: bad key");
    g_debug[key] = value;
    ++g_debug_epoch;
    return TF_OK;
}

#include "tf_conv_gemm.h"
#include "tf_elementwise.h"
namespace tfg {

int debug_flag(int key) { return (key >= 0 && key < 16) ? g_debug[key] : 0; }
int debug_epoch() { return g_debug_epoch; }

int conv_fprop(const ConvArgs& a, cudaStream_t st) {
    TF_REQUIRE(a.x && a.w && a.y, "conv_fprop: null pointer");
    TF_REQUIRE((a.x_lo == nullptr) == (a.w_lo == nullptr), "conv_fprop: x_lo and w_lo must be given together");
    TF_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0 && a.Cin > 0 && a.Cin % 32 == 0 && a.Cout >= 64 && a.Cout % 64 == 0,
               "conv_fprop: unsupported shape B=%d H=%d W=%d Cin=%d Cout=%d", a.B, a.H, a.W, a.Cin, a.Cout);
    TF_REQUIRE(a.ksize == 1 || a.ksize == 3, "conv_fprop: ksize must be 1 or 3");
    TF_REQUIRE(a.stride == 0 || a.stride == 1 || a.stride == 2, "conv_fprop: stride must be 1 or 2");
    int rc = ensure_device_state();
    if (rc) return rc;
    const int B = a.B, H = a.H, W = a.W, Cin = a.Cin, Cout = a.Cout;
    const int taps = a.ksize * a.ksize;
    GemmMaps maps;
    GemmParams p = {};
    const int stride = a.stride == 2 ? 2 : 1;
    const int Ho = stride == 2 ? (H + 1) / 2 : H, Wo = stride == 2 ? (W + 1) / 2 : W;      // output size
    const long long Mtot = (long long)B * Ho * Wo;
    const bool plain = !a.scale && !a.shift && !a.relu && !a.round_out;
    const ConvPlan cp = plan_conv(B, H, W, Cin, Cout, a.ksize, stride, a.x_lo ? 3 : 1, plain, a.res != nullptr, g_num_sms,
                                  g_debug[2], g_debug[4], g_debug[3] != 2);
    const bool spatial = cp.spatial != 0;
    const int BN = cp.bn;
    default_taps(p, taps); p.kblocks = Cin / 32; p.nseg = a.x_lo ? 3 : 1;
    p.scale = a.scale; p.shift = a.shift; p.relu = a.relu; p.round_out = a.round_out;
    p.accumulate = a.accumulate || g_debug[1];
    p.res = a.res; p.res_mask = a.res_mask; p.m_rows = Mtot;
    TF_REQUIRE(!a.res || (!spatial && !p.accumulate && !a.stats_partial && !a.x_lo),
               "conv_fprop: the residual epilogue needs a flat (1x1 stride-1), non-accumulating, single-pass GEMM without statistics");
    p.stats_partial = a.stats_partial; p.cout = Cout; p.img_w = Wo; p.img_h = Ho;
    if (g_debug[7] && !p.stats_partial) {          // probe: time the BN-statistics epilogue without the model around it
        static float* scratch = nullptr;
        if (!scratch) TF_CHECK_CUDA(cudaMalloc(&scratch, (size_t)2 * 592 * 2 * 1024 * sizeof(float)));
        p.stats_partial = scratch;
    }
    p.err_flag = g_err_flag;
    p.num_n_tiles = Cout / BN;
    const float* as[3] = {a.x, a.x_lo, a.x};
    const float* bs[3] = {a.w, a.w, a.w_lo};
    const long long M = Mtot;
    p.stride = stride;
    p.spatial = cp.spatial; p.num_m_tiles = cp.m_tiles; p.tiles_x = cp.tiles_x; p.tiles_y = cp.tiles_y; p.tw = cp.tw; p.th = cp.th;
    if (!spatial) {
        for (int s = 0; s < p.nseg; ++s)
            if ((rc = encode_2d(&maps.a[s], as[s], Cin, M, Cin, 32, BLOCK_M))) return rc;
        if ((rc = encode_2d(&maps.d, a.y, Cout, M, Cout, 32, 32))) return rc;          // one store box per epilogue warp
        if (a.res && (rc = encode_2d(&maps.res, a.res, Cout, M, Cout, 32, 32))) return rc;
    } else {
        for (int s = 0; s < p.nseg; ++s)
            if ((rc = encode_4d(&maps.a[s], as[s], Cin, W, H, B, 32, p.tw, p.th, CU_TENSOR_MAP_SWIZZLE_128B, stride))) return rc;
        {   // per-warp store box: 32 consecutive tile rows = (bw x bh) pixels of the patch
            const int bw = p.tw < 32 ? p.tw : 32, bh = 32 / bw;
            if ((rc = encode_4d(&maps.d, a.y, Cout, Wo, Ho, B, 32, bw, bh))) return rc;
        }
    }
    const bool two_cta = cp.two_cta != 0;
    for (int s = 0; s < p.nseg; ++s)
        if ((rc = encode_2d(&maps.b[s], bs[s], (uint64_t)taps * Cin, Cout, (uint64_t)taps * Cin, 32, two_cta ? 128 : BN))) return rc;
    int stats_rows = 0;
    p.main_tiles = cp.main_tiles; p.ksplit = cp.ksplit;
    if (two_cta) {
        // 2-CTA (cta_group::2) kernel for 256-wide tiles: each CTA of the pair stages half (128 rows) of the B tile
        if ((rc = launch_gemm2(maps, p, &stats_rows, st))) return rc;
        if (a.stats_blocks) *a.stats_blocks = stats_rows;
        return TF_OK;
    }
    const int grid = cp.grid;
    stats_rows = grid / p.num_n_tiles;
    const long long tail_pix0 = cp.tail_pix0;            // tail-wave split-K region (plan_conv), contiguous to the end
    if (tail_pix0 >= 0 && !p.accumulate)
        TF_CHECK_CUDA(cudaMemsetAsync(a.y + tail_pix0 * Cout, 0, (size_t)(Mtot - tail_pix0) * Cout * sizeof(float), st));
    if (BN == 256) rc = launch_gemm<256>(maps, p, grid, st);
    else if (BN == 128) rc = launch_gemm<128>(maps, p, grid, st);
    else rc = launch_gemm<64>(maps, p, grid, st);
    if (rc) return rc;
    if (p.stats_partial && tail_pix0 >= 0) {
        // BN statistics of the split rows (their epilogues only saw partial sums): appended partial rows
        int nb = 0;
        if ((rc = tfe::column_stats(a.y + tail_pix0 * Cout, Mtot - tail_pix0, Cout, p.stats_partial + (size_t)stats_rows * 2 * Cout, &nb, st))) return rc;
        stats_rows += nb;
    }
    if (a.stats_blocks) *a.stats_blocks = stats_rows;
    return TF_OK;
}

// Input gradient of a stride-2 convolution (3x3 pad 1, or 1x1) WITHOUT zero insertion: the input pixels fall into 4
// parity classes (py, px); each class is a small stride-1 correlation over dy -- 1, 2, 2 and 4 of the 9 flipped taps (1 tap,
// class (0,0) only, for a 1x1) -- written through a tensor map of that class's quarter lattice of dx.  9/4 tap
// evaluations per dy pixel instead of the 9 per dx pixel (= 36) of "zero-insert, then stride-1 dgrad".
//   a.x = dy [B, ceil(H/2), ceil(W/2), Cin]  (Cin = the convolution's Cout), a.w = packed dgrad weights [Cout][taps][Cin]
//   a.y = dx [B, H, W, Cout]; a.H, a.W = the dx size; a.accumulate adds into dx (1x1: untouched pixels keep their value)
int conv_dgrad_s2(const ConvArgs& a, cudaStream_t st) {
    TF_REQUIRE(a.x && a.w && a.y && !a.scale && !a.shift && !a.relu && !a.round_out && !a.stats_partial && !a.res,
               "conv_dgrad_s2: plain epilogue only");
    TF_REQUIRE((a.x_lo == nullptr) == (a.w_lo == nullptr), "conv_dgrad_s2: x_lo and w_lo must be given together");
    TF_REQUIRE(a.B > 0 && a.H > 1 && a.W > 1 && a.Cin % 32 == 0 && a.Cout % 64 == 0 && (a.ksize == 1 || a.ksize == 3),
               "conv_dgrad_s2: unsupported shape B=%d H=%d W=%d Cin=%d Cout=%d k=%d", a.B, a.H, a.W, a.Cin, a.Cout, a.ksize);
    int rc = ensure_device_state();
    if (rc) return rc;
    const int B = a.B, H = a.H, W = a.W, Cin = a.Cin, Cout = a.Cout;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, taps = a.ksize * a.ksize;
    const float* as[3] = {a.x, a.x_lo, a.x};
    const float* bs[3] = {a.w, a.w, a.w_lo};
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            if (a.ksize == 1 && (py || px)) continue;               // a 1x1/s2 only ever read the even pixels
            const int Hq = (H - py + 1) / 2, Wq = (W - px + 1) / 2;
            if (Hq <= 0 || Wq <= 0) continue;
            GemmMaps maps;
            GemmParams p = {};
            p.kblocks = Cin / 32; p.nseg = a.x_lo ? 3 : 1; p.stride = 1; p.spatial = 1;
            p.accumulate = a.accumulate; p.cout = Cout; p.img_w = Wq; p.img_h = Hq; p.err_flag = g_err_flag;
            int tw_[9], tox[9], toy[9];
            const int nt = s2_class_taps(a.ksize, py, px, tw_, tox, toy);
            for (int t = 0; t < nt; ++t) { p.tap_w[t] = (unsigned char)tw_[t]; p.tap_ox[t] = (signed char)tox[t]; p.tap_oy[t] = (signed char)toy[t]; }
            p.taps = nt;
            pick_tile(Wq, Hq, BLOCK_M, &p.tw, &p.th);
            p.tiles_x = (Wq + p.tw - 1) / p.tw; p.tiles_y = (Hq + p.th - 1) / p.th;
            p.num_m_tiles = B * p.tiles_x * p.tiles_y;
            int BN = 64; double best = 1e30;
            for (int cand = 256; cand >= 64; cand >>= 1) {
                if (Cout % cand) continue;
                const long long tiles = (long long)p.num_m_tiles * (Cout / cand);
                const double cost = (double)((tiles + g_num_sms - 1) / g_num_sms) * cand * (cand == 256 ? 1.0 : (cand == 128 ? 1.45 : 2.2));
                if (cost < best) { best = cost; BN = cand; }
            }
            p.num_n_tiles = Cout / BN;
            for (int s = 0; s < p.nseg; ++s) {
                if ((rc = encode_4d(&maps.a[s], as[s], Cin, Wo, Ho, B, 32, p.tw, p.th))) return rc;
                if ((rc = encode_2d(&maps.b[s], bs[s], (uint64_t)taps * Cin, Cout, (uint64_t)taps * Cin, 32, BN))) return rc;
            }
            const int bw = p.tw < 32 ? p.tw : 32, bh = 32 / bw;
            if ((rc = encode_4d_lattice2(&maps.d, a.y, Cout, W, H, B, py, px, 32, bw, bh))) return rc;
            const int tiles = p.num_m_tiles * p.num_n_tiles;
            const int grid = (tiles < g_num_sms ? tiles : g_num_sms) / p.num_n_tiles * p.num_n_tiles;
            p.main_tiles = tiles; p.ksplit = 1;
            if (BN == 256) rc = launch_gemm<256>(maps, p, grid, st);
            else if (BN == 128) rc = launch_gemm<128>(maps, p, grid, st);
            else rc = launch_gemm<64>(maps, p, grid, st);
            if (rc) return rc;
        }
    return TF_OK;
}

int conv_wgrad(const WgradArgs& a, cudaStream_t st) {
    TF_REQUIRE(a.x && a.dy && a.dw, "conv_wgrad: null pointer");
    TF_REQUIRE((a.x_lo == nullptr) == (a.dy_lo == nullptr), "conv_wgrad: x_lo and dy_lo must be given together");
    TF_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0 && a.Cin > 0 && a.Cin % 32 == 0 && a.Cout > 0 && a.Cout % 32 == 0,
               "conv_wgrad: unsupported shape B=%d H=%d W=%d Cin=%d Cout=%d", a.B, a.H, a.W, a.Cin, a.Cout);
    TF_REQUIRE(a.ksize == 1 || a.ksize == 3, "conv_wgrad: ksize must be 1 or 3");
    int rc = ensure_device_state();
    if (rc) return rc;
    const int B = a.B, H = a.H, W = a.W, Cin = a.Cin, Cout = a.Cout;
    const int ncols = a.ksize * a.ksize * Cin;
    const int BN = ncols <= 64 ? 64 : (ncols <= 128 ? 128 : 256);
    WgradMaps maps;
    WgradParams p = {};
    p.taps = a.ksize * a.ksize; p.cin = Cin; p.cout = Cout; p.err_flag = g_err_flag;
    p.nseg = a.x_lo ? 3 : 1;
    // Two accumulator tiles per CTA (dY rows m0 .. m0+255 share every activation tile) for the 3x3 weight gradients, which
    // are bound by the L2->shared-memory fill rate: measured on B200 at M = 38400, 256->256: 95.3 -> 60.4 us (0.58 -> 0.91 of
    // the TF32 burst peak).  The 1x1 weight gradients stream 196 MB through HBM for 20 GFLOP (41 us = 73 % of the HBM peak,
    // 0.59 of the tensor peak either way) and keep the one-tile kernel.  tf_debug_set(10, 1): one tile everywhere (A/B).
    const int MT = (a.ksize == 3 && Cout % (2 * BLOCK_M) == 0 && !g_debug[10]) ? 2 : 1;
    p.m_tiles = (Cout + BLOCK_M * MT - 1) / (BLOCK_M * MT);
    p.n_tiles = (ncols + BN - 1) / BN;
    {   // groups per B instruction: must divide both the tile (BN/32) and a tap's channel groups (Cin/32)
        int x = (Cin < BN ? Cin : BN) / 32, y = BN / 32;
        while (y) { const int t = x % y; x = y; y = t; }
        p.b_groups = x;
    }
    TF_REQUIRE(Cin >= BN || BN % Cin == 0 || a.ksize == 1, "conv_wgrad: Cin=%d does not tile BN=%d", Cin, BN);
    const float* as[3] = {a.dy, a.dy_lo, a.dy};
    const float* bs[3] = {a.x, a.x, a.x_lo};
    const int stride = a.stride == 2 ? 2 : 1;
    const int Ho = stride == 2 ? (H + 1) / 2 : H, Wo = stride == 2 ? (W + 1) / 2 : W;      // dy is [B,Ho,Wo,Cout], x is [B,H,W,Cin]
    const long long M = (long long)B * Ho * Wo;
    p.stride = stride;
    if (a.ksize == 1 && stride == 1) {
        p.spatial = 0; p.tiles_x = p.tiles_y = 1; p.tw = 32; p.th = 1;
        p.num_pblocks = (int)((M + 31) / 32);
        for (int s = 0; s < p.nseg; ++s) {
            if ((rc = encode_3d_grouped(&maps.a[s], as[s], Cout, M, 32, MT * BLOCK_M / 32))) return rc;
            if ((rc = encode_3d_grouped(&maps.b[s], bs[s], Cin, M, 32, p.b_groups))) return rc;
        }
    } else {
        p.spatial = 1;
        pick_tile(Wo, Ho, 32, &p.tw, &p.th);
        p.tiles_x = (Wo + p.tw - 1) / p.tw; p.tiles_y = (Ho + p.th - 1) / p.th;
        p.num_pblocks = B * p.tiles_x * p.tiles_y;
        for (int s = 0; s < p.nseg; ++s) {
            if ((rc = encode_5d_grouped(&maps.a[s], as[s], Cout, Wo, Ho, B, p.tw, p.th, MT * BLOCK_M / 32))) return rc;
            if ((rc = encode_5d_grouped(&maps.b[s], bs[s], Cin, W, H, B, p.tw, p.th, p.b_groups, stride))) return rc;
        }
    }
    if ((rc = encode_2d(&maps.d, a.dw, (uint64_t)p.taps * Cin, Cout, (uint64_t)p.taps * Cin, 32, BLOCK_M))) return rc;
    const int units = p.m_tiles * p.n_tiles;
    int splits = g_num_sms / units;                       // one wave of CTAs
    if (splits > p.num_pblocks) splits = p.num_pblocks;
    if (splits < 1) splits = 1;
    if (g_debug[0]) { splits = 1; p.plain_store = 1; }
    p.splits = splits;
    if (MT == 2) {
        if (BN == 256) return launch_wgrad<256, 2>(maps, p, st);
        if (BN == 128) return launch_wgrad<128, 2>(maps, p, st);
        return launch_wgrad<64, 2>(maps, p, st);
    }
    if (BN == 256) return launch_wgrad<256, 1>(maps, p, st);
    if (BN == 128) return launch_wgrad<128, 1>(maps, p, st);
    return launch_wgrad<64, 1>(maps, p, st);
}

}  // namespace tfg

// y[B,H,W,Cout] = conv(x[B,H,W,Cin], w) with stride 1 and "same" zero padding, NHWC fp32.
//   w_packed : [Cout][ksize*ksize][Cin]  (tap-major K), Cout and Cin multiples of 32 (Cout multiple of 64)
//   x_lo / w_lo (both or neither): low-order TF32 split parts for the 3-product parity mode
//   bias: optional [Cout].  Also serves as dgrad (flipped taps, transposed weights) and as a plain
//   row-major GEMM  y[M,Cout] = x[M,Cin] * w[Cout,Cin]^T  (ksize 1, B=1, H=1, W=M).
TF_API int tf_conv2d_nhwc(const float* x, const float* x_lo, int B, int H, int W, int Cin, const float* w_packed,
                          const float* w_lo, int Cout, int ksize, const float* bias, float* y, void* stream) {
    tfg::ConvArgs a = {};
    a.x = x; a.x_lo = x_lo; a.B = B; a.H = H; a.W = W; a.Cin = Cin;
    a.w = w_packed; a.w_lo = w_lo; a.Cout = Cout; a.ksize = ksize; a.shift = bias; a.y = y;
    return tfg::conv_fprop(a, (cudaStream_t)stream);
}

// 1x1 stride-1 GEMM with the residual epilogue: y[M,Cout] = x[M,Cin] * w^T + (res_mask bit ? res : 0)  -- the kernel the
// training backward uses for "dx = conv1 dgrad + masked upstream gradient" (the masked term is never materialised) and
// the inference forward for conv3 + shortcut.  res_mask: optional, 1 bit per element of res ((M*Cout+31)/32 words).
TF_API int tf_conv2d_nhwc_res(const float* x, int B, int H, int W, int Cin, const float* w_packed, int Cout, const float* res,
                              const uint32_t* res_mask, float* y, void* stream) {
    tfg::ConvArgs a = {};
    a.x = x; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.w = w_packed; a.Cout = Cout; a.ksize = 1; a.y = y;
    a.res = res; a.res_mask = res_mask;
    return tfg::conv_fprop(a, (cudaStream_t)stream);
}

// dw_packed[Cout][ksize*ksize][Cin] += sum_pixels dy[pixel, co] * x[pixel (+) tap, ci]   (caller zeroes dw_packed)
TF_API int tf_conv2d_wgrad_nhwc(const float* x, const float* dy, int B, int H, int W, int Cin, int Cout, int ksize,
                                float* dw_packed, void* stream) {
    tfg::WgradArgs a = {};
    a.x = x; a.dy = dy; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.dw = dw_packed;
    return tfg::conv_wgrad(a, (cudaStream_t)stream);
}

// Strided variants (stride 1 or 2, padding ksize/2): y is [B, ceil(H/s), ceil(W/s), Cout]; dy likewise.
TF_API int tf_conv2d_nhwc_strided(const float* x, int B, int H, int W, int Cin, const float* w_packed, int Cout, int ksize,
                                  int stride, const float* bias, float* y, void* stream) {
    tfg::ConvArgs a = {};
    a.x = x; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.w = w_packed; a.Cout = Cout; a.ksize = ksize; a.stride = stride;
    a.shift = bias; a.y = y;
    return tfg::conv_fprop(a, (cudaStream_t)stream);
}
TF_API int tf_conv2d_wgrad_nhwc_strided(const float* x, const float* dy, int B, int H, int W, int Cin, int Cout, int ksize,
                                        int stride, float* dw_packed, void* stream) {
    tfg::WgradArgs a = {};
    a.x = x; a.dy = dy; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.stride = stride; a.dw = dw_packed;
    return tfg::conv_wgrad(a, (cudaStream_t)stream);
}

TF_API int tf_conv2d_dgrad_s2_nhwc(const float* dy, int B, int H, int W, int Cdy, const float* w_packed, int Cdx, int ksize,
                                   int accumulate, float* dx, void* stream) {
    tfg::ConvArgs a = {};
    a.x = dy; a.B = B; a.H = H; a.W = W; a.Cin = Cdy; a.w = w_packed; a.Cout = Cdx; a.ksize = ksize; a.accumulate = accumulate; a.y = dx;
    return tfg::conv_dgrad_s2(a, (cudaStream_t)stream);
}

// Host-only views of the launch planning, for the CPU test-suite (no CUDA calls):
//   tf_conv_plan: out[16] = {spatial, tw, th, tiles_x, tiles_y, m_tiles, bn, n_tiles, grid, two_cta, main_tiles, ksplit,
//                           kiters, tail_pix0 (low 31 bits), tail_pix0 >> 31, 0}
TF_API int tf_conv_plan(int B, int H, int W, int Cin, int Cout, int ksize, int stride, int nseg, int plain, int has_res,
                        int num_sms, int* out16) {
    TF_REQUIRE(out16 && B > 0 && H > 0 && W > 0 && Cin % 32 == 0 && Cout % 64 == 0 && (ksize == 1 || ksize == 3) && num_sms > 0,
               "tf_conv_plan: bad args");
    const ConvPlan c = plan_conv(B, H, W, Cin, Cout, ksize, stride == 2 ? 2 : 1, nseg, plain != 0, has_res != 0, num_sms, 0, 0, true);
    const int v[16] = {c.spatial, c.tw, c.th, c.tiles_x, c.tiles_y, c.m_tiles, c.bn, c.n_tiles, c.grid, c.two_cta, c.main_tiles,
                       c.ksplit, c.kiters, (int)(c.tail_pix0 < 0 ? -1 : (c.tail_pix0 & 0x7fffffff)),
                       (int)(c.tail_pix0 < 0 ? -1 : (c.tail_pix0 >> 31)), 0};
    for (int i = 0; i < 16; ++i) out16[i] = v[i];
    return TF_OK;
}
//   tf_dgrad_s2_taps: the taps of parity class (py, px) of a stride-2 dgrad; returns the count (0..4), -1 on bad args
TF_API int tf_dgrad_s2_taps(int ksize, int py, int px, int* tap_w, int* tap_ox, int* tap_oy) {
    if (!(ksize == 1 || ksize == 3) || (py | px) & ~1 || !tap_w || !tap_ox || !tap_oy) return -1;
    return s2_class_taps(ksize, py, px, tap_w, tap_ox, tap_oy);
}

// Reads (and clears) the current device's pipeline error flag (set by a timed-out mbarrier wait).
TF_API int tf_gemm_error_flag(int* value) {
    TF_REQUIRE(value, "tf_gemm_error_flag: null");
    *value = 0;
    int dev = 0;
    TF_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES || !g_err_flag_dev[dev]) return TF_OK;
    TF_CHECK_CUDA(cudaMemcpy(value, g_err_flag_dev[dev], sizeof(int), cudaMemcpyDeviceToHost));
    TF_CHECK_CUDA(cudaMemset(g_err_flag_dev[dev], 0, sizeof(int)));
    return TF_OK;
}
