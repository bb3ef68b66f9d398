// Tap-GEMM kernel for sm_100a: TMA-fed, tcgen05.mma with TMEM accumulators, warp-specialised,
// persistent.  See tapgemm.cuh for the contraction and DESIGN.md §3 for the tiling.
#include "tapgemm.cuh"

#include <cuda_bf16.h>
#include <stdio.h>

#include "sm100.cuh"

namespace la {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 bf16 = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kThreads = 256;                     // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps4-7 epilogue
constexpr int kEpiThreads = 128;
constexpr int kRedKinds = 5;                      // red_s, red_d, red_rgb[3]

template <int BN>
struct Cfg {
    static constexpr int kBBytes = BN * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;   // two accumulator stages
    static constexpr int kNCh = BN / 32;
    static constexpr int kRaccBytes = kRedKinds * kNCh * kEpiThreads * 4;   // per-thread reduction accumulators
    static constexpr int kSmemBytes = kStages * kStageBytes + kRaccBytes + 256 /*barriers*/ + 1024 /*align slack*/;
};

struct TileCoord {
    int prob, nblk, n0, h0, w0;
};

__device__ __forceinline__ TileCoord decode_tile(const TapGemmParams& P, int t) {
    TileCoord c;
    c.nblk = t / P.m_tiles;                  // M fastest: a CTA's contiguous chunk shares the weight tile
    int m = t - c.nblk * P.m_tiles;
    const int per_prob = P.tiles_n * P.tiles_h * P.tiles_w;
    c.prob = m / per_prob;
    int local = m - c.prob * per_prob;
    int tx = local % P.tiles_w;
    int ty = (local / P.tiles_w) % P.tiles_h;
    int tn = local / (P.tiles_w * P.tiles_h);
    c.n0 = tn * P.nb;
    c.h0 = ty * P.th;
    c.w0 = tx * P.tw;
    return c;
}

__device__ __forceinline__ void tile_range(int total, int& begin, int& end) {
    begin = static_cast<int>(static_cast<long long>(blockIdx.x) * total / gridDim.x);
    end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * total / gridDim.x);
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16lo_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// store 32 consecutive values as bf16 (hi plane) and, if lo != null, the bf16 residual (lo plane)
__device__ __forceinline__ void store_bf16x32(void* hi, void* lo, long long off, const float (&v)[32]) {
    uint4* dh = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(hi) + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
        o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
        o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        dh[j] = o;
    }
    if (lo) {
        uint4* dl = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(lo) + off);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = v[8 * j + i] - bf16_round(v[8 * j + i]);
            uint4 o;
            o.x = pack_bf16(r[0], r[1]);
            o.y = pack_bf16(r[2], r[3]);
            o.z = pack_bf16(r[4], r[5]);
            o.w = pack_bf16(r[6], r[7]);
            dl[j] = o;
        }
    }
}

__device__ __forceinline__ void load_bf16x32(const void* hi, const void* lo, long long off, float (&v)[32]) {
    const uint4* sh = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(hi) + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 a = __ldg(sh + j);
        v[8 * j + 0] = bf16lo_f(a.x); v[8 * j + 1] = bf16hi_f(a.x);
        v[8 * j + 2] = bf16lo_f(a.y); v[8 * j + 3] = bf16hi_f(a.y);
        v[8 * j + 4] = bf16lo_f(a.z); v[8 * j + 5] = bf16hi_f(a.z);
        v[8 * j + 6] = bf16lo_f(a.w); v[8 * j + 7] = bf16hi_f(a.w);
    }
    if (lo) {
        const uint4* sl = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(lo) + off);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 a = __ldg(sl + j);
            v[8 * j + 0] += bf16lo_f(a.x); v[8 * j + 1] += bf16hi_f(a.x);
            v[8 * j + 2] += bf16lo_f(a.y); v[8 * j + 3] += bf16hi_f(a.y);
            v[8 * j + 4] += bf16lo_f(a.z); v[8 * j + 5] += bf16hi_f(a.z);
            v[8 * j + 6] += bf16lo_f(a.w); v[8 * j + 7] += bf16hi_f(a.w);
        }
    }
}

__device__ __forceinline__ void load_f32x32(const float* p, float (&v)[32]) {
    const float4* s = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 a = __ldg(s + j);
        v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
    }
}

// Transposing warp reduction: on return lane L holds sum over lanes of v[L].  31 shuffles.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// ------------------------------------------------------------------------------------
// Epilogue shared by the tensor-core kernel and its SIMT twin.  128 threads, thread = one
// accumulator row (pixel), 32 columns per chunk.

struct RowCtx {
    bool valid;
    int n;
    long long pix;       // output pixel index (n*OH + oh)*OW + ow
    long long px_in_img; // oh*OW + ow
};

__device__ __forceinline__ RowCtx make_row(const TapGemmParams& P, const TileCoord& tc, int row) {
    RowCtx r;
    const int box_px = P.th * P.tw;
    const int ni = row / box_px;
    const int rem = row - ni * box_px;
    const int hi = rem / P.tw;
    const int wi = rem - hi * P.tw;
    const int n = tc.n0 + ni, h = tc.h0 + hi, w = tc.w0 + wi;
    r.valid = (ni < P.nb) && (n < P.batch) && (h < P.vh) && (w < P.vw);
    r.n = n;
    const TapProblem& pr = P.prob[tc.prob];
    const int oh = h * P.osy + pr.oy0, ow = w * P.osx + pr.ox0;
    r.px_in_img = static_cast<long long>(oh) * P.OW + ow;
    r.pix = static_cast<long long>(n) * P.OH * P.OW + r.px_in_img;
    return r;
}

// Reduction bookkeeping: per-thread accumulators in shared memory, keyed on (sample, column
// block); flushed to global memory with one atomicAdd per (kind, column) when the key changes.
template <int BN>
__device__ __forceinline__ void racc_flush(const TapGemmParams& P, float* sracc, int key, int tid) {
    constexpr int NCH = BN / 32;
    epi_bar();
    if (key >= 0) {
        const int n = key / P.n_blocks, nblk = key - n * P.n_blocks;
        const int kinds = P.bwd_last ? 1 : (P.g_rgb ? kRedKinds : 2);
        for (int e = tid; e < kinds * NCH * 32; e += kEpiThreads) {
            const int L = e & 31, ch = (e >> 5) % NCH, kind = e / (32 * NCH);
            float* base = sracc + (kind * NCH + ch) * kEpiThreads + L;
            const float s = (base[0] + base[32]) + (base[64] + base[96]);
            base[0] = base[32] = base[64] = base[96] = 0.f;
            const long long col = static_cast<long long>(n) * P.n_total + nblk * BN + ch * 32 + L;
            if (kind == 0) atomicAdd(P.red_s + col, s);
            else if (kind == 1) atomicAdd(P.red_d + col, s);
            else atomicAdd(P.red_rgb + static_cast<long long>(kind - 2) * P.batch * P.n_total + col, s);
        }
    }
    epi_bar();
}

template <int BN>
__device__ __forceinline__ void racc_add(const TapGemmParams& P, float* sracc, int kind, int ch, int tid, int lane,
                                         float (&prod)[32], bool smem_mode, bool warp_uniform, const RowCtx& rc,
                                         int col0) {
    constexpr int NCH = BN / 32;
    float* gbase = kind == 0 ? P.red_s : (kind == 1 ? P.red_d : P.red_rgb + static_cast<long long>(kind - 2) * P.batch * P.n_total);
    if (warp_uniform) {
        const float s = warp_transpose_sum(prod, lane);
        if (smem_mode) {
            sracc[(kind * NCH + ch) * kEpiThreads + tid] += s;
        } else {
            const int n = __shfl_sync(0xffffffffu, rc.n, 0);
            const bool any = __any_sync(0xffffffffu, rc.valid);
            if (any && n < P.batch) atomicAdd(gbase + static_cast<long long>(n) * P.n_total + col0 + lane, s);
        }
    } else if (rc.valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(gbase + static_cast<long long>(rc.n) * P.n_total + col0 + j, prod[j]);
    }
}

template <int BN, int EPI, class LoadChunk>
__device__ __forceinline__ void epilogue_tile(const TapGemmParams& P, const TileCoord& tc, int tid, float* sracc,
                                              int& cur_key, LoadChunk&& load_chunk) {
    constexpr int NCH = BN / 32;
    const int lane = tid & 31;
    const RowCtx rc = make_row(P, tc, tid);
    const int box_px = P.th * P.tw;
    const bool warp_uniform = box_px >= 32;          // the 32 rows of a warp belong to one sample
    const bool smem_mode = P.nb == 1;                // the whole tile belongs to one sample

    if (EPI == kEpiBwd && smem_mode) {
        const int key = tc.n0 * P.n_blocks + tc.nblk;
        if (key != cur_key) {
            if (cur_key >= 0) racc_flush<BN>(P, sracc, cur_key, tid);
            cur_key = key;
        }
    }

    float nz = 0.f;
    float4 grgb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rc.valid) {
        if (EPI == kEpiFwd && P.noise) nz = __ldg(P.noise + rc.n * P.noise_stride_n + rc.px_in_img) * P.noise_scale;
        if (EPI == kEpiBwd && P.noise_prev) nz = __ldg(P.noise_prev + rc.n * P.noise_prev_stride_n + rc.px_in_img) * P.noise_prev_scale;
        if (EPI == kEpiBwd && P.g_rgb) grgb = __ldg(P.g_rgb + rc.pix);
    }
    float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
    float best_s[8];
    int best_i[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { best_s[t] = __int_as_float(0x7f800000); best_i[t] = -1; }

#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
        float acc[32];
        load_chunk(ch, acc);
        const int col0 = tc.nblk * BN + ch * 32;
        const long long eoff = rc.pix * P.n_total + col0;             // element offset in [pixel][N] tensors
        const long long coff = static_cast<long long>(rc.n) * P.n_total + col0;   // offset in [batch][N] tensors

        if constexpr (EPI == kEpiRawF32) {
            if (rc.valid) {
                float4* dst = reinterpret_cast<float4*>(P.raw_out + eoff);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
        } else if constexpr (EPI == kEpiTopK) {
            if (rc.valid && rc.pix < P.n_queries) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int code = col0 + j;
                    if (code < P.n_codes) {
                        float sc = fmaf(-2.f, acc[j], __ldg(P.code_sqnorm + code));
                        int id = code;
                        if (sc < best_s[7]) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) {       // sorted insertion, ascending
                                if (sc < best_s[t]) {
                                    const float ts = best_s[t]; const int ti = best_i[t];
                                    best_s[t] = sc; best_i[t] = id;
                                    sc = ts; id = ti;
                                }
                            }
                        }
                    }
                }
            }
        } else if constexpr (EPI == kEpiFwd) {
            if (rc.valid) {
                float dm[32], bs[32];
                load_f32x32(P.demod + coff, dm);
                load_f32x32(P.bias + col0, bs);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float z = fmaf(acc[j], dm[j], nz) + bs[j];
                    z = (z > 0.f ? z : z * P.act_slope) * P.act_gain;
                    if (P.act_clamp >= 0.f) z = fminf(fmaxf(z, -P.act_clamp), P.act_clamp);
                    acc[j] = z;
                }
                store_bf16x32(P.x_hi, P.split ? P.x_lo : nullptr, eoff, acc);
                if (P.rgbw) {
                    const float4* rw = P.rgbw + coff;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float4 w4 = __ldg(rw + j);
                        rgb0 = fmaf(acc[j], w4.x, rgb0);
                        rgb1 = fmaf(acc[j], w4.y, rgb1);
                        rgb2 = fmaf(acc[j], w4.z, rgb2);
                    }
                }
                if (P.s_next) {
                    load_f32x32(P.s_next + coff, dm);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] *= dm[j];
                    store_bf16x32(P.xs_hi, P.split ? P.xs_lo : nullptr, eoff, acc);
                }
            }
        } else {   // kEpiBwd
            float xp[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) xp[j] = 0.f;
            if (rc.valid)
                load_bf16x32(P.xp_hi, P.split ? P.xp_lo : nullptr,
                             static_cast<long long>(rc.n) * P.xp_stride_n + rc.px_in_img * P.n_total + col0, xp);
            float prod[32];
            // (1) style-gradient reduction: sum_px g_xs * x_{l-1}
#pragma unroll
            for (int j = 0; j < 32; ++j) prod[j] = rc.valid ? acc[j] * xp[j] : 0.f;
            racc_add<BN>(P, sracc, 0, ch, tid, lane, prod, smem_mode, warp_uniform, rc, col0);
            if (!P.bwd_last) {
                float sc[32];
                if (rc.valid) {
                    load_f32x32(P.s_cur + coff, sc);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] *= sc[j];      // g_x (conv part)
                    if (P.g_rgb) {
                        const float4* rw = P.rgbw_prev + coff;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float4 w4 = __ldg(rw + j);
                            acc[j] += grgb.x * w4.x + grgb.y * w4.y + grgb.z * w4.z;
                        }
                    }
                    // activation backward of layer l-1 (decided by its saved output) and y recovery
                    load_f32x32(P.bias_prev + col0, sc);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float out = xp[j];
                        const bool pos = out > 0.f;
                        float gz = acc[j] * P.act_gain * (pos ? 1.f : P.act_slope);
                        if (P.act_clamp >= 0.f && !(fabsf(out) < P.act_clamp)) gz = 0.f;
                        const float z = pos ? out / P.act_gain : out / (P.act_gain * P.act_slope);
                        prod[j] = gz * (z - nz - sc[j]);
                        acc[j] = gz;
                    }
                    load_f32x32(P.demod_prev + coff, sc);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] *= sc[j];      // g_y of layer l-1
                    store_bf16x32(P.gy_hi, P.split ? P.gy_lo : nullptr, eoff, acc);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) prod[j] = 0.f;
                }
                racc_add<BN>(P, sracc, 1, ch, tid, lane, prod, smem_mode, warp_uniform, rc, col0);
                if (P.g_rgb) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) prod[j] = xp[j] * grgb.x;
                    racc_add<BN>(P, sracc, 2, ch, tid, lane, prod, smem_mode, warp_uniform, rc, col0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) prod[j] = xp[j] * grgb.y;
                    racc_add<BN>(P, sracc, 3, ch, tid, lane, prod, smem_mode, warp_uniform, rc, col0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) prod[j] = xp[j] * grgb.z;
                    racc_add<BN>(P, sracc, 4, ch, tid, lane, prod, smem_mode, warp_uniform, rc, col0);
                }
            }
        }
    }
    if (EPI == kEpiTopK && rc.valid && rc.pix < P.n_queries) {
        const long long base = (rc.pix * P.n_blocks + tc.nblk) * P.topk;
#pragma unroll
        for (int t = 0; t < 8; ++t)
            if (t < P.topk) { P.cand_score[base + t] = best_s[t]; P.cand_idx[base + t] = best_i[t]; }
    }
    if (EPI == kEpiFwd && P.rgbw && rc.valid)
        P.rgb_part[static_cast<long long>(tc.nblk) * P.batch * P.OH * P.OW + rc.pix] = make_float4(rgb0, rgb1, rgb2, 0.f);
}

template <int BN>
__device__ __forceinline__ void racc_init(float* sracc, int tid) {
    for (int e = tid; e < kRedKinds * (BN / 32) * kEpiThreads; e += kEpiThreads) sracc[e] = 0.f;
    epi_bar();
}

// ------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1) tapgemm_kernel(const __grid_constant__ TapGemmParams P) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* sracc = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kRaccBytes);
    uint64_t* empty_bar = full_bar + C::kStages;
    uint64_t* tfull_bar = empty_bar + C::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = P.m_tiles * P.n_blocks;
    int t_begin, t_end;
    tile_range(total_tiles, t_begin, t_end);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&P.a_map[0]);
        prefetch_tmap(&P.b_map);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<C::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t a_tx_bytes = static_cast<uint32_t>(P.nb * P.th * P.tw) * 128u;

    if (warp == 0 && lane == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
            const TileCoord tc = decode_tile(P, t);
            const TapProblem& pr = P.prob[tc.prob];
            for (int ti = 0; ti < pr.ntaps; ++ti) {
                const Tap tap = P.taps[pr.tap_begin + ti];
                const CUtensorMap* amap = &P.a_map[tap.src];
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, P.err_flag, 1);
                    uint8_t* sa = smem + stage * C::kStageBytes;
                    mbar_expect_tx(&full_bar[stage], a_tx_bytes + C::kBBytes);
                    tma_load_4d(sa, amap, &full_bar[stage], kc * kBlockK, tc.w0 + tap.dx, tc.h0 + tap.dy, tc.n0);
                    tma_load_3d(sa + kABytes, &P.b_map, &full_bar[stage], kc * kBlockK, tc.nblk * BN, tap.widx);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
            const TileCoord tc = decode_tile(P, t);
            const int ksteps = P.prob[tc.prob].ntaps * P.kchunks;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1, P.err_flag, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(&full_bar[stage], phase, P.err_flag, 3);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
                const uint64_t adesc = make_sw128_kmajor_desc(sa);
                const uint64_t bdesc = make_sw128_kmajor_desc(sa + kABytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                    umma_bf16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                              (ks | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (128 threads)
        const int q = warp - 4;               // TMEM lane quarter == warp % 4
        const int tid = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        int cur_key = -1;
        if (EPI == kEpiBwd) racc_init<BN>(sracc, tid);
        for (int t = t_begin; t < t_end; ++t) {
            const TileCoord tc = decode_tile(P, t);
            mbar_wait(&tfull_bar[acc], acc_phase, P.err_flag, 4);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
            epilogue_tile<BN, EPI>(P, tc, tid, sracc, cur_key, [&](int ch, float (&v)[32]) {
                uint32_t u[32];
                tmem_ld32(t_addr + ch * 32, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]);
            });
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI == kEpiBwd && cur_key >= 0) racc_flush<BN>(P, sracc, cur_key, tid);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------
// SIMT twin: 128 threads, same tile walk and the same epilogue; the accumulator row is
// computed by brute force (zero when the problem has no taps: the backward-chain seed).
template <int BN, int EPI>
__global__ void __launch_bounds__(kEpiThreads) tapgemm_simt_kernel(const __grid_constant__ TapGemmParams P,
                                                                   const TapSimtOperands ops) {
    __shared__ float sracc[kRedKinds * (BN / 32) * kEpiThreads];
    const int tid = threadIdx.x;
    const int total_tiles = P.m_tiles * P.n_blocks;
    int t_begin, t_end;
    tile_range(total_tiles, t_begin, t_end);
    int cur_key = -1;
    if (EPI == kEpiBwd) racc_init<BN>(sracc, tid);
    const int K = P.kchunks * kBlockK;
    const int box_px = P.th * P.tw;
    for (int t = t_begin; t < t_end; ++t) {
        const TileCoord tc = decode_tile(P, t);
        const TapProblem& pr = P.prob[tc.prob];
        const int ni = tid / box_px, rem = tid - ni * box_px;
        const int n = tc.n0 + ni, h = tc.h0 + rem / P.tw, w = tc.w0 + rem % P.tw;
        const bool in_box = ni < P.nb && n < P.batch;
        epilogue_tile<BN, EPI>(P, tc, tid, sracc, cur_key, [&](int ch, float (&v)[32]) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            if (!in_box) return;
            for (int ti = 0; ti < pr.ntaps; ++ti) {
                const Tap tap = P.taps[pr.tap_begin + ti];
                const int hh = h + tap.dy, ww = w + tap.dx;
                if (hh < 0 || hh >= ops.a_h || ww < 0 || ww >= ops.a_w) continue;
                const __nv_bfloat16* arow = reinterpret_cast<const __nv_bfloat16*>(ops.a_ptrs[tap.src]) +
                                            n * ops.a_sn + hh * ops.a_sh + ww * ops.a_sw;
                const __nv_bfloat16* wmat = reinterpret_cast<const __nv_bfloat16*>(ops.w) +
                                            (static_cast<long long>(tap.widx) * (EPI == kEpiTopK ? P.n_codes : P.n_total) + tc.nblk * BN + ch * 32) * K;
                const int rows_total = EPI == kEpiTopK ? P.n_codes : P.n_total;
                const int jmax = rows_total - (tc.nblk * BN + ch * 32);
                for (int k = 0; k < K; ++k) {
                    const float a = __bfloat162float(arow[k]);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < jmax) v[j] = fmaf(a, __bfloat162float(wmat[static_cast<long long>(j) * K + k]), v[j]);
                }
            }
        });
    }
    if (EPI == kEpiBwd && cur_key >= 0) racc_flush<BN>(P, sracc, cur_key, tid);
}

template <int BN, int EPI>
int launch_bn_epi(const TapGemmParams& p, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg<BN>::kSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        attr_set = true;
    }
    const int total = p.m_tiles * p.n_blocks;
    int grid = total < num_sms ? total : num_sms;
    if (grid <= 0) return 0;
    for (int i = 0; i < p.nprob; ++i)
        if (p.prob[i].ntaps <= 0) return static_cast<int>(cudaErrorInvalidValue);   // accumulator would be undefined
    tapgemm_kernel<BN, EPI><<<grid, kThreads, Cfg<BN>::kSmemBytes, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
}

template <int BN>
int launch_bn(const TapGemmParams& p, int num_sms, cudaStream_t stream) {
    switch (p.epilogue) {
        case kEpiRawF32: return launch_bn_epi<BN, kEpiRawF32>(p, num_sms, stream);
        case kEpiFwd: return launch_bn_epi<BN, kEpiFwd>(p, num_sms, stream);
        case kEpiBwd: return launch_bn_epi<BN, kEpiBwd>(p, num_sms, stream);
        case kEpiTopK: return launch_bn_epi<BN, kEpiTopK>(p, num_sms, stream);
        default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

template <int BN, int EPI>
int launch_simt_bn_epi(const TapGemmParams& p, const TapSimtOperands& ops, cudaStream_t stream) {
    const int total = p.m_tiles * p.n_blocks;
    int grid = total < 148 * 8 ? total : 148 * 8;
    if (grid <= 0) return 0;
    tapgemm_simt_kernel<BN, EPI><<<grid, kEpiThreads, 0, stream>>>(p, ops);
    return static_cast<int>(cudaGetLastError());
}

template <int BN>
int launch_simt_bn(const TapGemmParams& p, const TapSimtOperands& ops, cudaStream_t stream) {
    switch (p.epilogue) {
        case kEpiRawF32: return launch_simt_bn_epi<BN, kEpiRawF32>(p, ops, stream);
        case kEpiFwd: return launch_simt_bn_epi<BN, kEpiFwd>(p, ops, stream);
        case kEpiBwd: return launch_simt_bn_epi<BN, kEpiBwd>(p, ops, stream);
        case kEpiTopK: return launch_simt_bn_epi<BN, kEpiTopK>(p, ops, stream);
        default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

}  // namespace

int launch_tapgemm(const TapGemmParams& p, int num_sms, cudaStream_t stream) {
    const int bn = p.n_total / p.n_blocks;
    switch (bn) {
        case 256: return launch_bn<256>(p, num_sms, stream);
        case 128: return launch_bn<128>(p, num_sms, stream);
        case 64: return launch_bn<64>(p, num_sms, stream);
        default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

int launch_tapgemm_simt(const TapGemmParams& p, const TapSimtOperands& ops, cudaStream_t stream) {
    const int bn = p.n_total / p.n_blocks;
    switch (bn) {
        case 256: return launch_simt_bn<256>(p, ops, stream);
        case 128: return launch_simt_bn<128>(p, ops, stream);
        case 64: return launch_simt_bn<64>(p, ops, stream);
        default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return -1;
        fn = reinterpret_cast<EncodeFn>(ptr);
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr,
                    bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return static_cast<int>(r);
}

}  // namespace la
