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
constexpr int kThreads = 384;                     // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-11 epilogue
constexpr int kEpiThreads = 256;                  // two epilogue warpgroups; warp w owns TMEM lanes 32*(w%4)..+31
constexpr int kStileStride = 68;                  // floats per row of the 128x64 transpose tile (272 B: conflict-free both ways)
constexpr int kRegsProducer = 56, kRegsEpilogue = 224;   // setmaxnreg: 128*56 + 256*224 = 64512 <= 65536

template <int BN, int EPI>
struct Cfg {
    static constexpr int kBBytes = BN * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = EPI == kEpiBwd ? (BN == 256 ? 3 : (BN == 128 ? 5 : 7)) : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
    static constexpr int kTmemCols = 2 * BN;      // two accumulator stages
    static constexpr int kStileBytes = EPI == kEpiBwd ? kBlockM * kStileStride * 4 : 0;
    static constexpr int kFlushBytes = EPI == kEpiBwd ? 5 * BN * 4 : 0;
    static constexpr int kRowInfoBytes = EPI == kEpiBwd ? kBlockM * 4 : 0;
    static constexpr int kEpiSmemBytes = kStileBytes + kFlushBytes + kRowInfoBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiSmemBytes + 256 /*barriers*/ + 1024 /*align slack*/;
};

struct TileCoord {
    int prob, nblk, n0, h0, w0;
};

__device__ __forceinline__ TileCoord decode_tile(const TapGemmParams& P, int t) {
    TileCoord c;
    c.nblk = t / P.m_tiles;                  // M fastest: a CTA's contiguous chunk shares the weight tile
    int m = t - c.nblk * P.m_tiles;
    int p = 0;
#pragma unroll
    for (int i = 1; i < kMaxProblems; ++i)
        if (i < P.nprob && m >= P.prob[i].tile_begin) p = i;
    c.prob = p;
    const TapProblem& pr = P.prob[p];
    int local = m - pr.tile_begin;
    int tx = local % pr.tiles_w;
    int ty = (local / pr.tiles_w) % pr.tiles_h;
    int tn = local / (pr.tiles_w * pr.tiles_h);
    c.n0 = tn * P.nb;
    c.h0 = ty * P.th;
    c.w0 = tx * P.tw;
    return c;
}

// Contiguous, COST-balanced tile range of this CTA: a tile costs max(ntaps, 1) of its problem (the
// phases of the transposed convolution have 4 / 2 / 2 / 1 taps).  Tiles are ordered [nblk][problem][tile].
__device__ __forceinline__ long long tiles_before(const TapGemmParams& P, long long cost_per_nblk, long long x) {
    long long nb_full = x / cost_per_nblk;
    if (nb_full >= P.n_blocks) return static_cast<long long>(P.n_blocks) * P.m_tiles;
    long long rem = x - nb_full * cost_per_nblk, count = nb_full * P.m_tiles;
    for (int p = 0; p < P.nprob; ++p) {
        const TapProblem& pr = P.prob[p];
        const long long c = pr.ntaps > 0 ? pr.ntaps : 1;
        const long long tiles = static_cast<long long>(pr.tiles_h) * pr.tiles_w * P.tiles_n;
        if (rem >= tiles * c) { count += tiles; rem -= tiles * c; }
        else { count += (rem + c - 1) / c; break; }
    }
    return count;
}
__device__ __forceinline__ void tile_range(const TapGemmParams& P, int& begin, int& end) {
    long long cost = 0;
    for (int p = 0; p < P.nprob; ++p)
        cost += static_cast<long long>(P.prob[p].tiles_h) * P.prob[p].tiles_w * P.tiles_n * (P.prob[p].ntaps > 0 ? P.prob[p].ntaps : 1);
    const long long total = cost * P.n_blocks;
    begin = static_cast<int>(tiles_before(P, cost, total * blockIdx.x / gridDim.x));
    end = static_cast<int>(tiles_before(P, cost, total * (blockIdx.x + 1) / gridDim.x));
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

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


// ------------------------------------------------------------------------------------
// Epilogues shared by the tensor-core kernel and its SIMT twin.  256 threads: thread t owns
// accumulator row (t & 127) of the 32-column chunks with (chunk & 1) == (t >> 7).

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
    const TapProblem& pr = P.prob[tc.prob];
    r.valid = (ni < P.nb) && (n < P.batch) && (h < pr.vh) && (w < pr.vw);
    r.n = n;
    const int oh = h * P.osy + pr.oy0, ow = w * P.osx + pr.ox0;
    r.px_in_img = static_cast<long long>(oh) * P.OW + ow;
    r.pix = static_cast<long long>(n) * P.OH * P.OW + r.px_in_img;
    return r;
}

// ---- row-owner epilogues: raw store, forward activation, top-k
template <int BN, int EPI, class LoadChunk>
__device__ __forceinline__ void rowowner_tile(const TapGemmParams& P, const TileCoord& tc, int t, LoadChunk&& load_chunk) {
    constexpr int NCH = BN / 32;
    const int eg = t >> 7;
    const RowCtx rc = make_row(P, tc, t & 127);
    float nz = 0.f;
    if (EPI == kEpiFwd && rc.valid && P.noise) nz = __ldg(P.noise + rc.n * P.noise_stride_n + rc.px_in_img) * P.noise_scale;
    float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
    float best_s[8];
    int best_i[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best_s[k] = __int_as_float(0x7f800000); best_i[k] = -1; }
    const bool tk_valid = rc.valid && rc.pix < P.n_queries;

#pragma unroll 1
    for (int ch = eg; ch < NCH; ch += 2) {
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
        } else if constexpr (EPI == kEpiStoreBf16) {
            if (rc.valid) store_bf16x32(P.x_hi, P.split ? P.x_lo : nullptr, eoff, acc);
        } else if constexpr (EPI == kEpiTopK) {
            if (tk_valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int code = col0 + j;
                    if (code < P.n_codes) {
                        float sc = fmaf(-2.f, acc[j], __ldg(P.code_sqnorm + code));
                        int id = code;
                        if (sc < best_s[7]) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {       // sorted insertion, ascending
                                if (sc < best_s[k]) {
                                    const float ts = best_s[k]; const int ti = best_i[k];
                                    best_s[k] = sc; best_i[k] = id;
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
        }
    }
    if (EPI == kEpiTopK && tk_valid) {
        const long long base = ((rc.pix * P.n_blocks + tc.nblk) * 2 + eg) * P.topk;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < P.topk) { P.cand_score[base + k] = best_s[k]; P.cand_idx[base + k] = best_i[k]; }
    }
    if (EPI == kEpiFwd && P.rgbw && rc.valid)     // one partial per (column block, epilogue group): summed in fixed order later
        P.rgb_part[static_cast<long long>(tc.nblk * 2 + eg) * P.batch * P.OH * P.OW + rc.pix] = make_float4(rgb0, rgb1, rgb2, 0.f);
}

// ---- column-owner backward epilogue.
// Phase A (row owners): TMEM -> registers -> 128x64 fp32 tile in shared memory.  Phase B: thread
// (cp = t & 31, rg = t >> 5) owns columns 2cp, 2cp+1 of the chunk and rows rg*16..+15, so the
// per-column coefficients sit in registers, global accesses are 128-byte coalesced rows and the
// style-gradient column sums are plain per-thread accumulations (10 registers per 64-column chunk).
template <int BN>
struct BwdState {
    float racc[BN / 64][10];     // (red_s, red_d, red_rgb0..2) x 2 columns, per 64-column chunk
    int key;                      // (sample, column block) the accumulators belong to, -1 = empty
};

template <int BN>
__device__ __forceinline__ void bwd_state_init(BwdState<BN>& st) {
#pragma unroll
    for (int c = 0; c < BN / 64; ++c)
#pragma unroll
        for (int k = 0; k < 10; ++k) st.racc[c][k] = 0.f;
    st.key = -1;
}

__device__ __forceinline__ float* red_base(const TapGemmParams& P, int kind) {
    return kind == 0 ? P.red_s : (kind == 1 ? P.red_d : P.red_rgb + static_cast<long long>(kind - 2) * P.batch * P.n_total);
}

template <int BN>
__device__ __forceinline__ void bwd_flush(const TapGemmParams& P, BwdState<BN>& st, int t, float* sflush) {
    constexpr int NCH = BN / 64;
    const int cp = t & 31, rg = t >> 5;
    const int kinds = P.bwd_last ? 1 : (P.g_rgb ? 5 : 2);
    const int n = st.key / P.n_blocks, nblk = st.key - n * P.n_blocks;
    if (P.nb == 1) {
        // every thread of the CTA holds the same key: reduce the 8 row groups through shared memory
#pragma unroll 1
        for (int round = 0; round < 8; ++round) {
            if (rg == round) {
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        if ((k >> 1) < kinds) {
                            float* p = sflush + (k >> 1) * BN + c * 64 + 2 * cp + (k & 1);
                            *p = (round == 0 ? 0.f : *p) + st.racc[c][k];
                        }
                    }
            }
            epi_bar();
        }
        for (int e = t; e < kinds * BN; e += kEpiThreads) {
            const int kind = e / BN, col = e - kind * BN;
            atomicAdd(red_base(P, kind) + static_cast<long long>(n) * P.n_total + nblk * BN + col, sflush[e]);
        }
        epi_bar();
    } else if (n < P.batch) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int k = 0; k < 10; ++k)
                if ((k >> 1) < kinds)
                    atomicAdd(red_base(P, k >> 1) + static_cast<long long>(n) * P.n_total + nblk * BN + c * 64 + 2 * cp + (k & 1), st.racc[c][k]);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int k = 0; k < 10; ++k) st.racc[c][k] = 0.f;
}

template <int BN, bool kHaveAcc, class LoadChunk, class Release>
__device__ __forceinline__ void bwd_tile(const TapGemmParams& P, const TileCoord& tc, int t, float* stile, float* sflush, int* rowinfo,
                                         BwdState<BN>& st, LoadChunk&& load_chunk, Release&& release_acc) {
    constexpr int NCH = BN / 64;
    const int row = t & 127, eg = t >> 7;        // phase A
    const int cp = t & 31, rg = t >> 5;          // phase B
    const int box_px = P.th * P.tw;
    const int ni_rg = (rg * 16) / box_px;
    const int n_rg = tc.n0 + ni_rg;
    const bool n_ok = ni_rg < P.nb && n_rg < P.batch;
    const int new_key = (P.nb == 1 ? tc.n0 : n_rg) * P.n_blocks + tc.nblk;
    if (new_key != st.key) {
        if (st.key >= 0) bwd_flush<BN>(P, st, t, sflush);
        st.key = new_key;
    }
    if (eg == 0) {
        const RowCtx rc = make_row(P, tc, row);
        rowinfo[row] = rc.valid ? static_cast<int>(rc.px_in_img) : -1;
    }
    const float inv_gain = 1.f / P.act_gain, inv_gain_slope = 1.f / (P.act_gain * P.act_slope);
    const long long img_px = static_cast<long long>(P.OH) * P.OW;
    const unsigned* xph = reinterpret_cast<const unsigned*>(P.xp_hi);
    const unsigned* xpl = P.split ? reinterpret_cast<const unsigned*>(P.xp_lo) : nullptr;
    unsigned* gyh = reinterpret_cast<unsigned*>(P.gy_hi);
    unsigned* gyl = P.split ? reinterpret_cast<unsigned*>(P.gy_lo) : nullptr;

#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (kHaveAcc) {
            float acc[32];
            load_chunk(c * 2 + eg, acc);
            float4* dst = reinterpret_cast<float4*>(stile + row * kStileStride + eg * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            if (c == NCH - 1) release_acc();
        }
        epi_bar();
        const int col = tc.nblk * BN + c * 64 + 2 * cp;
        const long long coff = static_cast<long long>(n_rg) * P.n_total + col;
        float2 sc = make_float2(0.f, 0.f), dm = sc, bs = sc;
        float4 rw0 = make_float4(0.f, 0.f, 0.f, 0.f), rw1 = rw0;
        if (n_ok && !P.bwd_last) {
            sc = __ldg(reinterpret_cast<const float2*>(P.s_cur + coff));
            dm = __ldg(reinterpret_cast<const float2*>(P.demod_prev + coff));
            bs = __ldg(reinterpret_cast<const float2*>(P.bias_prev + col));
            if (P.g_rgb) { rw0 = __ldg(P.rgbw_prev + coff); rw1 = __ldg(P.rgbw_prev + coff + 1); }
        }
        float r[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) r[k] = 0.f;
        if (n_ok) {
            // rows in batches of 8: all global loads of a batch are issued before any use (memory-level parallelism)
            const unsigned* xbase_h = xph + ((static_cast<long long>(n_rg) * P.xp_stride_n + col) >> 1);
            const unsigned* xbase_l = xpl ? xpl + ((static_cast<long long>(n_rg) * P.xp_stride_n + col) >> 1) : nullptr;
            const float4* gbase = P.g_rgb ? P.g_rgb + static_cast<long long>(n_rg) * img_px : nullptr;
            const float* nbase = (P.noise_prev && !P.bwd_last) ? P.noise_prev + n_rg * P.noise_prev_stride_n : nullptr;
            const long long gybase = (static_cast<long long>(n_rg) * img_px * P.n_total + col) >> 1;
            const int half_n = P.n_total >> 1;
#pragma unroll
            for (int b8 = 0; b8 < 2; ++b8) {
                int pxi[8];
                unsigned xu[8], xl[8];
                float nzv[8];
                float4 gv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pxi[i] = rowinfo[rg * 16 + b8 * 8 + i];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int p = pxi[i] < 0 ? 0 : pxi[i];
                    xu[i] = __ldg(xbase_h + static_cast<long long>(p) * half_n);
                    xl[i] = xbase_l ? __ldg(xbase_l + static_cast<long long>(p) * half_n) : 0u;
                    nzv[i] = nbase ? __ldg(nbase + p) : 0.f;
                    gv[i] = gbase ? __ldg(gbase + p) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (pxi[i] < 0) continue;
                    const int rr = rg * 16 + b8 * 8 + i;
                    float2 a = make_float2(0.f, 0.f);
                    if (kHaveAcc) a = *reinterpret_cast<const float2*>(stile + rr * kStileStride + 2 * cp);
                    const float x0 = bf16lo_f(xu[i]) + bf16lo_f(xl[i]), x1 = bf16hi_f(xu[i]) + bf16hi_f(xl[i]);
                    r[0] = fmaf(a.x, x0, r[0]);
                    r[1] = fmaf(a.y, x1, r[1]);
                    if (!P.bwd_last) {
                        float g0 = a.x * sc.x, g1 = a.y * sc.y;
                        if (P.g_rgb) {
                            const float4 g = gv[i];
                            g0 += g.x * rw0.x + g.y * rw0.y + g.z * rw0.z;
                            g1 += g.x * rw1.x + g.y * rw1.y + g.z * rw1.z;
                            r[4] = fmaf(x0, g.x, r[4]); r[5] = fmaf(x1, g.x, r[5]);
                            r[6] = fmaf(x0, g.y, r[6]); r[7] = fmaf(x1, g.y, r[7]);
                            r[8] = fmaf(x0, g.z, r[8]); r[9] = fmaf(x1, g.z, r[9]);
                        }
                        const float nz = nzv[i] * P.noise_prev_scale;
                        // activation backward of layer l-1 decided by its saved output; y recovered from it
                        const bool p0 = x0 > 0.f, p1 = x1 > 0.f;
                        float gz0 = g0 * P.act_gain * (p0 ? 1.f : P.act_slope);
                        float gz1 = g1 * P.act_gain * (p1 ? 1.f : P.act_slope);
                        if (P.act_clamp >= 0.f) {
                            if (!(fabsf(x0) < P.act_clamp)) gz0 = 0.f;
                            if (!(fabsf(x1) < P.act_clamp)) gz1 = 0.f;
                        }
                        const float z0 = x0 * (p0 ? inv_gain : inv_gain_slope), z1 = x1 * (p1 ? inv_gain : inv_gain_slope);
                        r[2] = fmaf(gz0, z0 - nz - bs.x, r[2]);
                        r[3] = fmaf(gz1, z1 - nz - bs.y, r[3]);
                        const float y0 = gz0 * dm.x, y1 = gz1 * dm.y;
                        const long long goff = gybase + static_cast<long long>(pxi[i]) * half_n;
                        gyh[goff] = pack_bf16(y0, y1);
                        if (gyl) gyl[goff] = pack_bf16(y0 - bf16_round(y0), y1 - bf16_round(y1));
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) st.racc[c][k] += r[k];
        epi_bar();           // the tile / rowinfo may be overwritten from here on
    }
}

// ------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1) tapgemm_kernel(const __grid_constant__ TapGemmParams P) {
    using C = Cfg<BN, EPI>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + C::kStages * C::kStageBytes;
    float* stile = reinterpret_cast<float*>(epi_smem);
    float* sflush = reinterpret_cast<float*>(epi_smem + C::kStileBytes);
    int* rowinfo = reinterpret_cast<int*>(epi_smem + C::kStileBytes + C::kFlushBytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + C::kEpiSmemBytes);
    uint64_t* empty_bar = full_bar + C::kStages;
    uint64_t* tfull_bar = empty_bar + C::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int t_begin, t_end;
    tile_range(P, t_begin, t_end);

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
            mbar_init(&tempty_bar[a], kEpiThreads / 32);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<C::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t a_tx_bytes = static_cast<uint32_t>(P.nb * P.th * P.tw) * 128u;

    if (warp < 4) {
        setmaxnreg_dec<kRegsProducer>();
        if (warp == 0 && lane == 0) {
            // ---------------------------------------------------------------- TMA producer
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
            // ---------------------------------------------------------------- MMA issuer
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
        }
    } else {
        // -------------------------------------------------------------------- epilogue (256 threads)
        setmaxnreg_inc<kRegsEpilogue>();
        const int t = threadIdx.x - 128;
        const int q = warp & 3;               // TMEM lane quarter
        int acc = 0;
        uint32_t acc_phase = 0;
        BwdState<BN> st;
        if (EPI == kEpiBwd) bwd_state_init<BN>(st);
        for (int tile = t_begin; tile < t_end; ++tile) {
            const TileCoord tc = decode_tile(P, tile);
            mbar_wait(&tfull_bar[acc], acc_phase, P.err_flag, 4);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
            auto load_chunk = [&](int ch, float (&v)[32]) {
                uint32_t u[32];
                tmem_ld32(t_addr + ch * 32, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]);
            };
            auto release = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            };
            if constexpr (EPI == kEpiBwd) {
                bwd_tile<BN, true>(P, tc, t, stile, sflush, rowinfo, st, load_chunk, release);
            } else {
                rowowner_tile<BN, EPI>(P, tc, t, load_chunk);
                release();
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI == kEpiBwd && st.key >= 0) bwd_flush<BN>(P, st, t, sflush);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------
// SIMT twin: 256 threads, same tile walk and the same epilogue code; the accumulator chunk is
// computed by brute force.  Debug cross-check of the tensor-core path (tests), never a fallback.
template <int BN, int EPI>
__global__ void __launch_bounds__(kEpiThreads) tapgemm_simt_kernel(const __grid_constant__ TapGemmParams P,
                                                                   const TapSimtOperands ops) {
    __shared__ __align__(16) float stile[EPI == kEpiBwd ? kBlockM * kStileStride : 4];
    __shared__ float sflush[EPI == kEpiBwd ? 5 * BN : 1];
    __shared__ int rowinfo[kBlockM];
    const int t = threadIdx.x;
    int t_begin, t_end;
    tile_range(P, t_begin, t_end);
    BwdState<BN> st;
    if (EPI == kEpiBwd) bwd_state_init<BN>(st);
    const int K = P.kchunks * kBlockK;
    const int box_px = P.th * P.tw;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const TileCoord tc = decode_tile(P, tile);
        const TapProblem& pr = P.prob[tc.prob];
        const int row = t & 127;
        const int ni = row / box_px, rem = row - ni * box_px;
        const int n = tc.n0 + ni, h = tc.h0 + rem / P.tw, w = tc.w0 + rem % P.tw;
        const bool in_box = ni < P.nb && n < P.batch;
        auto load_chunk = [&](int ch, float (&v)[32]) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            if (!in_box) return;
            const int rows_total = EPI == kEpiTopK ? P.n_codes : P.n_total;
            const int jmax = rows_total - (tc.nblk * BN + ch * 32);
            for (int ti = 0; ti < pr.ntaps; ++ti) {
                const Tap tap = P.taps[pr.tap_begin + ti];
                const int hh = h + tap.dy, ww = w + tap.dx;
                if (hh < 0 || hh >= ops.a_hs[tap.src] || ww < 0 || ww >= ops.a_ws[tap.src]) continue;
                const __nv_bfloat16* arow = reinterpret_cast<const __nv_bfloat16*>(ops.a_ptrs[tap.src]) +
                                            n * ops.a_sn + hh * ops.a_sh + ww * ops.a_sw;
                const __nv_bfloat16* wmat = reinterpret_cast<const __nv_bfloat16*>(ops.w) +
                                            (static_cast<long long>(tap.widx) * rows_total + tc.nblk * BN + ch * 32) * K;
                for (int k = 0; k < K; ++k) {
                    const float a = __bfloat162float(arow[k]);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < jmax) v[j] = fmaf(a, __bfloat162float(wmat[static_cast<long long>(j) * K + k]), v[j]);
                }
            }
        };
        if constexpr (EPI == kEpiBwd) bwd_tile<BN, true>(P, tc, t, stile, sflush, rowinfo, st, load_chunk, [] {});
        else rowowner_tile<BN, EPI>(P, tc, t, load_chunk);
    }
    if (EPI == kEpiBwd && st.key >= 0) bwd_flush<BN>(P, st, t, sflush);
}

// Seed of the backward chain: the backward epilogue with a zero accumulator (activation backward
// of the top layer from the toRGB gradient alone).  Memory-bound elementwise pass.
template <int BN>
__global__ void __launch_bounds__(kEpiThreads, 2) tapgemm_seed_kernel(const __grid_constant__ TapGemmParams P) {
    __shared__ float sflush[5 * BN];
    __shared__ int rowinfo[kBlockM];
    const int t = threadIdx.x;
    int t_begin, t_end;
    tile_range(P, t_begin, t_end);
    BwdState<BN> st;
    bwd_state_init<BN>(st);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const TileCoord tc = decode_tile(P, tile);
        bwd_tile<BN, false>(P, tc, t, nullptr, sflush, rowinfo, st, [](int, float (&)[32]) {}, [] {});
    }
    if (st.key >= 0) bwd_flush<BN>(P, st, t, sflush);
}

template <int BN, int EPI>
int launch_bn_epi(const TapGemmParams& p, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg<BN, EPI>::kSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        attr_set = true;
    }
    const int total = p.m_tiles * p.n_blocks;
    int grid = total < num_sms ? total : num_sms;
    if (grid <= 0) return 0;
    for (int i = 0; i < p.nprob; ++i)
        if (p.prob[i].ntaps <= 0) return static_cast<int>(cudaErrorInvalidValue);   // accumulator would be undefined
    tapgemm_kernel<BN, EPI><<<grid, kThreads, Cfg<BN, EPI>::kSmemBytes, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
}

template <int BN>
int launch_bn(const TapGemmParams& p, int num_sms, cudaStream_t stream) {
    switch (p.epilogue) {
        case kEpiRawF32: return launch_bn_epi<BN, kEpiRawF32>(p, num_sms, stream);
        case kEpiFwd: return launch_bn_epi<BN, kEpiFwd>(p, num_sms, stream);
        case kEpiBwd: return launch_bn_epi<BN, kEpiBwd>(p, num_sms, stream);
        case kEpiTopK: return launch_bn_epi<BN, kEpiTopK>(p, num_sms, stream);
        case kEpiStoreBf16: return launch_bn_epi<BN, kEpiStoreBf16>(p, num_sms, stream);
        default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

template <int BN, int EPI>
int launch_simt_bn_epi(const TapGemmParams& p, const TapSimtOperands& ops, cudaStream_t stream) {
    const int total = p.m_tiles * p.n_blocks;
    int grid = total < 148 * 4 ? total : 148 * 4;
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
        case kEpiStoreBf16: return launch_simt_bn_epi<BN, kEpiStoreBf16>(p, ops, stream);
        default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

template <int BN>
int launch_seed_bn(const TapGemmParams& p, int num_sms, cudaStream_t stream) {
    const int total = p.m_tiles * p.n_blocks;
    int grid = total < num_sms * 6 ? total : num_sms * 6;
    if (grid <= 0) return 0;
    tapgemm_seed_kernel<BN><<<grid, kEpiThreads, 0, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace

int launch_tapgemm_seed(const TapGemmParams& p, int num_sms, cudaStream_t stream) {
    if (p.epilogue != kEpiBwd) return static_cast<int>(cudaErrorInvalidValue);
    switch (p.n_total / p.n_blocks) {
        case 256: return launch_seed_bn<256>(p, num_sms, stream);
        case 128: return launch_seed_bn<128>(p, num_sms, stream);
        case 64: return launch_seed_bn<64>(p, num_sms, stream);
        default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

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
