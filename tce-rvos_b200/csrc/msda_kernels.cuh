// msda_kernels.cuh — sm_100a kernels for multi-scale deformable attention.
//
// Written from scratch for B200; this is NOT the Deformable-DETR extension.  What the
// reference computes (semantics only) is defined by
//   forward  : models/ops/src/cuda/ms_deform_im2col_cuda.cuh:421-452 (point gating `h>-1 && w>-1
//              && h<H && w<W`), :34-85 (per-corner zero padding, weights hh*hw, hh*lw, lh*hw, lh*lw)
//   backward : ms_deform_im2col_cuda.cuh:88-160 (grad formulas), :457-563 (per (b,q,m) reduction)
//
// Design:
//  * The op is a gather: per (query, head) it reads L*P*4 corner rows of D channels.  For
//    D = 32 a corner row is exactly one 128-byte line (fp32) or one 64-byte half line (bf16).
//  * FAST PATH (D == 32, L*P == 16): a warp owns HPW heads of one query.  A head is served by
//    LPH lanes, each holding CPL = 32/LPH channels in one 16-byte vector, so ONE warp-wide
//    LDG.128 fetches the same corner for HPW heads (fp32: 4 heads x 128 B, bf16: 8 heads x 64 B).
//    The per-point work (coordinate scaling, floor, bounds masks, bilinear weights) is done ONCE
//    by the lane that loaded the point (coalesced 16-byte loads of sampling_loc/attn_weight) and
//    broadcast inside the LPH-lane group with 4 (forward) / 5 (backward) shuffles; the reference redoes it in all
//    32 channel threads.
//  * BACKWARD: grad_value goes out as 16-byte vector reductions (red.global.add.v4.f32 — one
//    instruction per lane per corner instead of 4 scalar atomics), after the corners that
//    consecutive points of one (query, head, level) share have been merged by the lanes that own
//    the points (the pass is bound by the number of line updates reaching the L2); the channel dot products
//    needed by grad_attn_weight / grad_sampling_loc are reduced across the LPH lanes with a
//    transposing butterfly (log2(LPH) exchange rounds that halve the live set each round) that
//    ends with every lane holding exactly the points it loaded, so the two gradients are written
//    once, coalesced, with no shared memory, no __syncthreads and no memset pass.
//  * GENERIC PATH: any D / L / P / dtype (incl. fp64 for gradcheck parity with the reference's
//    models/ops/test.py): warp per (query, head), lanes strided over channels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// --------------------------------------------------------------------------------------------
// scalar conversions
// --------------------------------------------------------------------------------------------
template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

template <typename A, typename T> __device__ __forceinline__ A cvt_to(T v) { return static_cast<A>(v); }
template <> __device__ __forceinline__ float cvt_to<float, __nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float cvt_to<float, __half>(__half v) { return __half2float(v); }

template <typename T, typename A> __device__ __forceinline__ T cvt_from(A v) { return static_cast<T>(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_from<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cvt_from<__half, float>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float floor_acc(float v) { return floorf(v); }
__device__ __forceinline__ double floor_acc(double v) { return floor(v); }

// --------------------------------------------------------------------------------------------
// One sampling point, fully resolved.  Gating and padding follow the reference exactly:
// the point contributes only if (h_im > -1 && w_im > -1 && h_im < H && w_im < W); each of the
// four corners is dropped individually when it falls outside the level (zero padding).
// --------------------------------------------------------------------------------------------
template <typename A>
struct PointGeom {
    A lh, lw;      // fractional parts
    int h_low, w_low;
    int mask;      // bit c set <=> corner c is read (c: 0=(lo,lo) 1=(lo,hi) 2=(hi,lo) 3=(hi,hi))
};

template <typename A>
__device__ __forceinline__ PointGeom<A> resolve_point(A x, A y, int H, int W)
{
    PointGeom<A> g;
    const A h_im = y * static_cast<A>(H) - static_cast<A>(0.5);
    const A w_im = x * static_cast<A>(W) - static_cast<A>(0.5);
    const bool inside = (h_im > static_cast<A>(-1)) && (w_im > static_cast<A>(-1)) &&
                        (h_im < static_cast<A>(H)) && (w_im < static_cast<A>(W));
    const A hf = floor_acc(h_im);
    const A wf = floor_acc(w_im);
    // A gated-off point must contribute exactly nothing, also when the location is non-finite or
    // overflows (inf - inf = NaN would otherwise reach the fast kernels' fmaf(w, 0, acc)); the
    // reference skips such points entirely (ms_deform_im2col_cuda.cuh:441).
    g.lh = inside ? h_im - hf : static_cast<A>(0);
    g.lw = inside ? w_im - wf : static_cast<A>(0);
    // Only converted when inside (finite, in range); otherwise the value is never used.
    g.h_low = inside ? static_cast<int>(hf) : 0;
    g.w_low = inside ? static_cast<int>(wf) : 0;
    const bool top = g.h_low >= 0, bot = g.h_low + 1 <= H - 1;
    const bool lft = g.w_low >= 0, rgt = g.w_low + 1 <= W - 1;
    int mask = 0;
    if (inside) {
        mask = (int)(top && lft) | ((int)(top && rgt) << 1) | ((int)(bot && lft) << 2) | ((int)(bot && rgt) << 3);
    }
    g.mask = mask;
    return g;
}

// --------------------------------------------------------------------------------------------
// vector atomics
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4_f32(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void red_add_v4_bf16x2(void *addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("red.relaxed.gpu.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

template <typename T> __device__ __forceinline__ void atomic_add_any(T *p, T v) { atomicAdd(p, v); }

// --------------------------------------------------------------------------------------------
// FAST PATH configuration
// --------------------------------------------------------------------------------------------
template <typename VT> struct FastCfg;
template <> struct FastCfg<float> {
    static constexpr int CPL = 4;            // channels per lane (one 16-byte vector)
    static constexpr int LPH = 8;            // lanes per head
    static constexpr int LOG_LPH = 3;
};
template <> struct FastCfg<__nv_bfloat16> {
    static constexpr int CPL = 8;
    static constexpr int LPH = 4;
    static constexpr int LOG_LPH = 2;
};

// Occupancy targets (CTAs of 256 threads per SM) per kernel family; overridable at build time
// for tuning sweeps (tools/variants.py).
#ifndef MSDA_FWD_MIN_BLOCKS_F32
#define MSDA_FWD_MIN_BLOCKS_F32 4
#endif
#ifndef MSDA_FWD_MIN_BLOCKS_BF16
#define MSDA_FWD_MIN_BLOCKS_BF16 3
#endif
#ifndef MSDA_BWD_MIN_BLOCKS_F32
#define MSDA_BWD_MIN_BLOCKS_F32 3
#endif
#ifndef MSDA_BWD_MIN_BLOCKS_BF16
#define MSDA_BWD_MIN_BLOCKS_BF16 2
#endif
#ifndef MSDA_FWD_BATCH
#define MSDA_FWD_BATCH 2
#endif
#ifndef MSDA_QMAP_FWD
#define MSDA_QMAP_FWD 0   // 1: the warps of a CTA take CONSECUTIVE QUERIES of one head group (8 queries x 1 group) instead of
#endif                    //    consecutive items (4 queries x 2 groups): neighbouring queries share corner rows in the SM's L1
#ifndef MSDA_QMAP_BWD
#define MSDA_QMAP_BWD 0
#endif
#ifndef MSDA_MERGE
#define MSDA_MERGE 1   // 1: backward: corners of the points of one (query, head, level) that fall on the same pixel are
                       // merged by the lanes that own the points before anything is broadcast (TCE geometry only)
#endif
#ifndef MSDA_EXP
#define MSDA_EXP 0     // timing experiments only (tools/variants.py); 0 = product
#endif

constexpr int kFastD = 32;
constexpr int kFastLP = 16;

// one vector load of CPL channels -> float[CPL] (16 bytes, or 8 bytes for 4 x bf16)
template <typename VT, int CPL> struct VecIOn;
template <> struct VecIOn<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4])
    {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void zero(float (&v)[4]) { v[0] = v[1] = v[2] = v[3] = 0.f; }
    static __device__ __forceinline__ void load_stream(const float *p, float (&v)[4])
    {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store_stream(float *p, const float (&v)[4])
    {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <> struct VecIOn<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8])
    {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i]     = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void zero(float (&v)[8])
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    static __device__ __forceinline__ void load_stream(const __nv_bfloat16 *p, float (&v)[8])
    {
        const uint4 t = __ldcs(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i]     = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store_stream(__nv_bfloat16 *p, const float (&v)[8])
    {
        uint4 t;
        t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
        t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
        __stcs(reinterpret_cast<uint4 *>(p), t);
    }
};

// 4 x bf16 (8 bytes): the backward's 8-lanes-per-head layout for 16-bit values
template <> struct VecIOn<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void unpack(const uint2 t, float (&v)[4])
    {
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[4]) { unpack(__ldg(reinterpret_cast<const uint2 *>(p)), v); }
    static __device__ __forceinline__ void zero(float (&v)[4]) { v[0] = v[1] = v[2] = v[3] = 0.f; }
    static __device__ __forceinline__ void load_stream(const __nv_bfloat16 *p, float (&v)[4]) { unpack(__ldcs(reinterpret_cast<const uint2 *>(p)), v); }
    static __device__ __forceinline__ void store_stream(__nv_bfloat16 *p, const float (&v)[4])
    {
        __stcs(reinterpret_cast<uint2 *>(p), make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3])));
    }
};
template <typename VT> using VecIO = VecIOn<VT, FastCfg<VT>::CPL>;

// Loads of the PPL points a lane owns: PPL*2 location scalars and PPL weights, contiguous.
template <typename LT, int PPL> struct PointIO;
template <> struct PointIO<float, 2> {
    static __device__ __forceinline__ void load_loc(const float *p, float (&xy)[4])
    {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        xy[0] = t.x; xy[1] = t.y; xy[2] = t.z; xy[3] = t.w;
    }
    static __device__ __forceinline__ void load_aw(const float *p, float (&a)[2])
    {
        const float2 t = __ldcs(reinterpret_cast<const float2 *>(p));
        a[0] = t.x; a[1] = t.y;
    }
    static __device__ __forceinline__ void store_loc(float *p, const float (&xy)[4])
    {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(xy[0], xy[1], xy[2], xy[3]));
    }
    static __device__ __forceinline__ void store_aw(float *p, const float (&a)[2])
    {
        __stcs(reinterpret_cast<float2 *>(p), make_float2(a[0], a[1]));
    }
};
template <> struct PointIO<float, 4> {
    static __device__ __forceinline__ void load_loc(const float *p, float (&xy)[8])
    {
        const float4 t0 = __ldcs(reinterpret_cast<const float4 *>(p));
        const float4 t1 = __ldcs(reinterpret_cast<const float4 *>(p) + 1);
        xy[0] = t0.x; xy[1] = t0.y; xy[2] = t0.z; xy[3] = t0.w;
        xy[4] = t1.x; xy[5] = t1.y; xy[6] = t1.z; xy[7] = t1.w;
    }
    static __device__ __forceinline__ void load_aw(const float *p, float (&a)[4])
    {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
    }
    static __device__ __forceinline__ void store_loc(float *p, const float (&xy)[8])
    {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(xy[0], xy[1], xy[2], xy[3]));
        __stcs(reinterpret_cast<float4 *>(p) + 1, make_float4(xy[4], xy[5], xy[6], xy[7]));
    }
    static __device__ __forceinline__ void store_aw(float *p, const float (&a)[4])
    {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(a[0], a[1], a[2], a[3]));
    }
};
template <> struct PointIO<__nv_bfloat16, 2> {
    static __device__ __forceinline__ void load_loc(const __nv_bfloat16 *p, float (&xy)[4])
    {
        const uint2 t = __ldcs(reinterpret_cast<const uint2 *>(p));
        xy[0] = __uint_as_float(t.x << 16); xy[1] = __uint_as_float(t.x & 0xffff0000u);
        xy[2] = __uint_as_float(t.y << 16); xy[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void load_aw(const __nv_bfloat16 *p, float (&a)[2])
    {
        const uint32_t t = __ldcs(reinterpret_cast<const uint32_t *>(p));
        a[0] = __uint_as_float(t << 16); a[1] = __uint_as_float(t & 0xffff0000u);
    }
    static __device__ __forceinline__ void store_loc(__nv_bfloat16 *p, const float (&xy)[4])
    {
        __stcs(reinterpret_cast<uint2 *>(p), make_uint2(pack_bf16x2(xy[0], xy[1]), pack_bf16x2(xy[2], xy[3])));
    }
    static __device__ __forceinline__ void store_aw(__nv_bfloat16 *p, const float (&a)[2])
    {
        __stcs(reinterpret_cast<uint32_t *>(p), pack_bf16x2(a[0], a[1]));
    }
};
template <> struct PointIO<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void load_loc(const __nv_bfloat16 *p, float (&xy)[8])
    {
        const uint4 t = __ldcs(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xy[2 * i]     = __uint_as_float(w[i] << 16);
            xy[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void load_aw(const __nv_bfloat16 *p, float (&a)[4])
    {
        const uint2 t = __ldcs(reinterpret_cast<const uint2 *>(p));
        a[0] = __uint_as_float(t.x << 16); a[1] = __uint_as_float(t.x & 0xffff0000u);
        a[2] = __uint_as_float(t.y << 16); a[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store_loc(__nv_bfloat16 *p, const float (&xy)[8])
    {
        uint4 t;
        t.x = pack_bf16x2(xy[0], xy[1]); t.y = pack_bf16x2(xy[2], xy[3]);
        t.z = pack_bf16x2(xy[4], xy[5]); t.w = pack_bf16x2(xy[6], xy[7]);
        __stcs(reinterpret_cast<uint4 *>(p), t);
    }
    static __device__ __forceinline__ void store_aw(__nv_bfloat16 *p, const float (&a)[4])
    {
        uint2 t;
        t.x = pack_bf16x2(a[0], a[1]); t.y = pack_bf16x2(a[2], a[3]);
        __stcs(reinterpret_cast<uint2 *>(p), t);
    }
};

// What one lane knows about a point it owns.  The first four members are what the other lanes
// of the head need (4 shuffles per point); the rest stays with the owner.
struct OwnedPoint {
    float ta, tb;  // (1-lh)*attn, lh*attn  -> corner weights are ta*hw, ta*lw, tb*hw, tb*lw
    float lw;
    int packed;    // (base << 4) | mask; base = pixel index in the frame of corner (h_low, w_low),
                   // possibly negative (arithmetic shift restores it); mask = corners that are read
    float lh, a;   // owner only (backward)
    int H, W;
};

// Level lookup straight from the device-resident int64 tables (L1-resident after first touch).
__device__ __forceinline__ void level_dims(const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                                           int l, int &H, int &W, int &start)
{
    H = static_cast<int>(__ldg(shapes + 2 * l));
    W = static_cast<int>(__ldg(shapes + 2 * l + 1));
    start = static_cast<int>(__ldg(lsi + l));
}

__device__ __forceinline__ OwnedPoint own_point(float x, float y, float a, int H, int W, int start)
{
    OwnedPoint o;
    const PointGeom<float> g = resolve_point<float>(x, y, H, W);
    o.ta = (1.f - g.lh) * a;
    o.tb = g.lh * a;
    o.lw = g.lw;
    o.packed = ((start + g.h_low * W + g.w_low) << 4) | g.mask;
    o.lh = g.lh; o.a = a; o.H = H; o.W = W;
    return o;
}

// A point as seen by every lane of the head after the broadcast.
struct PointView {
    float w0, w1, w2, w3;   // corner weights x attention weight
    int base, mask;
};

template <int LPH>
__device__ __forceinline__ PointView view_point(const OwnedPoint &o, int src)
{
    const float ta = __shfl_sync(kFull, o.ta, src, LPH);
    const float tb = __shfl_sync(kFull, o.tb, src, LPH);
    const float lw = __shfl_sync(kFull, o.lw, src, LPH);
    const int packed = __shfl_sync(kFull, o.packed, src, LPH);
    const float hw = 1.f - lw;
    PointView v;
    v.w0 = ta * hw; v.w1 = ta * lw; v.w2 = tb * hw; v.w3 = tb * lw;
    v.base = packed >> 4;
    v.mask = packed & 15;
    return v;
}

// --------------------------------------------------------------------------------------------
// Merging the duplicate corners of one (query, head, level).
//
// The P points a head samples on one level are close to each other (the module initialises them 1, 2, 3, 4 px
// along one direction, ms_deform_attn.py:63-71), so the 2x2 corner blocks of consecutive points overlap: on
// the benchmark's model-like inputs 22 % of all (point, corner) pairs hit a pixel that the previous point of
// the same level already touches (1.4 % for uniform random locations; profiles/r02_merge.md).  Every such pair
// is a 128-byte line reduced twice by the backward, and the backward is bound by exactly that number
// (DESIGN.md section 4).  All contributions of one (query, head) carry the same grad_out, so merging is scalar
// work: the duplicate corner's weight is ADDED to the first occurrence's and the duplicate is switched off.
// It is done once per point by the lane that owns the point, in a chain from the last point of the level to
// the first (D -> C -> B -> A), so that a pixel shared by three points ends up in one slot.  BACKWARD ONLY:
// in the forward a duplicate corner is a load that hits the line its twin has just brought into L1, the
// cheapest access there is, and the extra hand-over word costs more than it saves (measured: 170 -> 177 us).
//
// A merged point travels to the other lanes of the head as its four corner weights (no longer separable into
// (ta, tb) x (hw, lw)) and one word: base << 8 | load mask << 4 | active mask.  `load mask` = corners whose
// value row is read (the backward needs every corner's dot product, merged or not), `active mask` = corners
// that are reduced into grad_value.
// --------------------------------------------------------------------------------------------
struct MergedPoint {
    float w[4];
    int packed;        // base << 8 | load mask << 4 | active mask
    __device__ __forceinline__ int base() const { return packed >> 8; }
    __device__ __forceinline__ int lmask() const { return (packed >> 4) & 15; }
    __device__ __forceinline__ int amask() const { return packed & 15; }
};

__device__ __forceinline__ MergedPoint merged_from(const OwnedPoint &o)
{
    MergedPoint m;
    const float hw = 1.f - o.lw;
    m.w[0] = o.ta * hw; m.w[1] = o.ta * o.lw; m.w[2] = o.tb * hw; m.w[3] = o.tb * o.lw;
    const int mask = o.packed & 15;
    m.packed = ((o.packed >> 4) << 8) | (mask << 4) | mask;
    return m;
}

// Fold the corners of point b that coincide with an active corner of the point at (base_a, mask_a) into wa[] and
// switch them off in b.  Two active corners have the same pixel index only if they are the same pixel (both are
// inside the level).  `commit` = false makes the call a no-op (lanes that only take part in the exchange).
__device__ __forceinline__ void fold_into(float (&wa)[4], int mask_a, int base_a, MergedPoint &b, int W, bool commit)
{
    const int d = b.base() - base_a;
    const int off[4] = {0, 1, W, W + 1};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool hit = commit && (d + off[c] == off[k]) && ((b.packed >> c) & 1) && ((mask_a >> k) & 1);
            wa[k] += hit ? b.w[c] : 0.f;
            b.packed = hit ? (b.packed & ~(1 << c)) : b.packed;
        }
    }
}

// All points a lane owns lie on one level (TCE geometry: P = 4; PPL = 2 -> the level's points are split over the
// lane pair (2l, 2l+1), PPL = 4 -> one lane owns the whole level).
template <int PPL, int LPH>
__device__ __forceinline__ void merge_level(MergedPoint (&m)[PPL], int W, int j)
{
    if constexpr (PPL == 4) {
        fold_into(m[2].w, m[2].amask(), m[2].base(), m[3], W, true);
        fold_into(m[1].w, m[1].amask(), m[1].base(), m[2], W, true);
        fold_into(m[0].w, m[0].amask(), m[0].base(), m[1], W, true);
    } else {
        static_assert(PPL == 2, "lanes own 2 or 4 points");
        const bool odd = (j & 1) != 0;
        fold_into(m[0].w, m[0].amask(), m[0].base(), m[1], W, odd);             // D -> C on the odd lane
        // C -> B: B is the even lane's second point.  The odd lane gets B's position, works out what C adds to
        // each corner of B, and sends the four sums back.
        const int pb = __shfl_xor_sync(kFull, m[1].packed, 1, LPH);
        float add[4] = {0.f, 0.f, 0.f, 0.f};
        fold_into(add, pb & 15, pb >> 8, m[0], W, odd);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float r = __shfl_xor_sync(kFull, add[k], 1, LPH);
            m[1].w[k] += odd ? 0.f : r;
        }
        fold_into(m[0].w, m[0].amask(), m[0].base(), m[1], W, !odd);            // B -> A on the even lane
    }
}

template <int LPH>
__device__ __forceinline__ MergedPoint view_merged(const MergedPoint &m, int src)
{
    MergedPoint v;
#pragma unroll
    for (int k = 0; k < 4; ++k) v.w[k] = __shfl_sync(kFull, m.w[k], src, LPH);
    v.packed = __shfl_sync(kFull, m.packed, src, LPH);
    return v;
}

// --------------------------------------------------------------------------------------------
// Fused module prologue (models/ops/modules/ms_deform_attn.py:99-109).  In PROLOGUE mode the
// `loc` / `aw` kernel arguments are the RAW outputs of the two Linear layers (sampling offsets
// and pre-softmax logits); the lane that owns a point turns them into a sampling location
// (reference point + offset / normaliser, 2-d or 4-d reference form) and a softmax weight
// (max/sum over the 16 points of the head = LPH-lane shuffles), optionally materialising both
// because the decoder consumes them (tce_deformable_transformer.py:752-758).
// --------------------------------------------------------------------------------------------
struct ProArgs {
    const void *ref;        // [N, Lq, L, ref_dim], loc dtype
    int ref_dim;            // 2 or 4
    int n_levels;
    float n_points;         // P as float (the 4-d form divides by it)
    void *loc_out;          // optional [N, Lq, M, L, P, 2]
    void *aw_out;           // optional [N, Lq, M, L, P]
    float *grad_ref;        // backward, optional, fp32 [N, Lq, L, ref_dim], zero-filled by the host side
};

template <int LPH>
__device__ __forceinline__ float group_max(float v)
{
#pragma unroll
    for (int o = LPH >> 1; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o, LPH));
    return v;
}
template <int LPH>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int o = LPH >> 1; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o, LPH);
    return v;
}

// Per owned point: what the prologue did, kept for the backward chain rule.
struct ProPoint {
    float sx, sy;           // d(loc)/d(offset): 1/W,1/H (2-d form, applied as a division) or rw*0.5/P, rh*0.5/P
    float ox, oy;           // raw offsets (4-d form needs them for d/d(rw,rh))
};

template <typename LT, int LPH, int PPL>
__device__ __forceinline__ void prologue_points(float (&xy)[2 * PPL], float (&a)[PPL], const int (&Hs)[PPL],
                                                const int (&Ws)[PPL], const int (&lv)[PPL],
                                                const ProArgs &pro, int64_t nq, ProPoint (&pp_out)[PPL])
{
    // softmax over the 16 logits of this head
    float mx = a[0];
#pragma unroll
    for (int pp = 1; pp < PPL; ++pp) mx = fmaxf(mx, a[pp]);
    mx = group_max<LPH>(mx);
    float sum = 0.f;
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) { a[pp] = expf(a[pp] - mx); sum += a[pp]; }
    sum = group_sum<LPH>(sum);
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) a[pp] = a[pp] / sum;

    const LT *ref = static_cast<const LT *>(pro.ref) + nq * (pro.n_levels * pro.ref_dim);
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) {
        const LT *r = ref + lv[pp] * pro.ref_dim;
        const float ox = xy[2 * pp], oy = xy[2 * pp + 1];
        pp_out[pp].ox = ox; pp_out[pp].oy = oy;
        const float rx = cvt_to<float>(r[0]), ry = cvt_to<float>(r[1]);
        if (pro.ref_dim == 2) {
            xy[2 * pp]     = rx + ox / static_cast<float>(Ws[pp]);
            xy[2 * pp + 1] = ry + oy / static_cast<float>(Hs[pp]);
            pp_out[pp].sx = static_cast<float>(Ws[pp]);
            pp_out[pp].sy = static_cast<float>(Hs[pp]);
        } else {
            const float rw = cvt_to<float>(r[2]), rh = cvt_to<float>(r[3]);
            xy[2 * pp]     = rx + ox / pro.n_points * rw * 0.5f;
            xy[2 * pp + 1] = ry + oy / pro.n_points * rh * 0.5f;
            pp_out[pp].sx = rw;
            pp_out[pp].sy = rh;
        }
    }
}

// item -> (query, head group of the warp).  QMAP = false: item = nq * groups + hg.  QMAP = true: the items are taken in
// super-blocks of (warps per CTA) queries x groups, and inside a super-block the head group is the slow index, so that the
// warps of one CTA work on consecutive queries of the same heads.  Returns false for the padding of the last super-block.
template <bool QMAP>
__device__ __forceinline__ bool map_item(int64_t item, int64_t n_items, int groups, int64_t &nq, int &hg)
{
    if (!QMAP) {
        nq = item / groups;
        hg = static_cast<int>(item - nq * groups);
        return item < n_items;
    }
    const int wpb = static_cast<int>(blockDim.x >> 5);
    const int per = wpb * groups;
    const int64_t super = item / per;
    const int r = static_cast<int>(item - super * per);
    hg = r / wpb;
    nq = super * wpb + (r - hg * wpb);
    return nq < n_items / groups;
}

// --------------------------------------------------------------------------------------------
// FAST forward.  One warp = HPW heads of one (frame, query).
//   item = nq * (M / HPW) + head_group
// TCE = true fixes M = 8 heads, L = 4 levels, P = 4 points at compile time (TCE-RVOS / the
// Deformable-DETR default): pixel and row strides become immediates / 4 per-level registers.
// --------------------------------------------------------------------------------------------
template <typename VT, typename LT, bool PROLOGUE, bool TCE>
__global__ void __launch_bounds__(256, (sizeof(VT) == 4 ? MSDA_FWD_MIN_BLOCKS_F32 : MSDA_FWD_MIN_BLOCKS_BF16))
fwd_fast_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                const int64_t *__restrict__ lsi, const LT *__restrict__ loc,
                const LT *__restrict__ aw, VT *__restrict__ out,
                int64_t n_items, int64_t Lq, int64_t S, int M, int log2P, ProArgs pro)
{
    using Cfg = FastCfg<VT>;
    constexpr int CPL = Cfg::CPL, LPH = Cfg::LPH, HPW = kWarp / LPH, PPL = kFastLP / LPH;

    const int lane = threadIdx.x & 31;
    const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (TCE) { M = 8; log2P = 2; }
    const int groups = M / HPW;
    int64_t nq;
    int hg;
    if (!map_item<MSDA_QMAP_FWD != 0>(item, n_items, groups, nq, hg)) return;   // warp-uniform
    const int64_t n = nq / Lq;
    const int g = lane / LPH, j = lane % LPH;
    const int m = hg * HPW + g;

    // ---- own PPL points: coalesced 16-byte streaming loads, resolved once ----
    const int64_t qh0 = nq * M + hg * HPW;                 // first (query, head) of this warp
    float xy[2 * PPL], a[PPL];
    PointIO<LT, PPL>::load_loc(loc + qh0 * (kFastLP * 2) + lane * (2 * PPL), xy);
    PointIO<LT, PPL>::load_aw(aw + qh0 * kFastLP + lane * PPL, a);

    int Hs[PPL], Ws[PPL], starts[PPL], lv[PPL];
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) {
        lv[pp] = (j * PPL + pp) >> log2P;
        level_dims(shapes, lsi, lv[pp], Hs[pp], Ws[pp], starts[pp]);
    }
    if (PROLOGUE) {
        ProPoint unused[PPL];
        prologue_points<LT, LPH, PPL>(xy, a, Hs, Ws, lv, pro, nq, unused);
        if (pro.loc_out)
            PointIO<LT, PPL>::store_loc(static_cast<LT *>(pro.loc_out) + qh0 * (kFastLP * 2) + lane * (2 * PPL), xy);
        if (pro.aw_out)
            PointIO<LT, PPL>::store_aw(static_cast<LT *>(pro.aw_out) + qh0 * kFastLP + lane * PPL, a);
    }
    OwnedPoint own[PPL];
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp)
        own[pp] = own_point(xy[2 * pp], xy[2 * pp + 1], a[pp], Hs[pp], Ws[pp], starts[pp]);
    // All offsets inside a frame are 32-bit byte offsets.
    const int pix = M * kFastD;                            // elements between neighbouring pixels
    const int pixb = pix * static_cast<int>(sizeof(VT));   // bytes
    const char *vlane = reinterpret_cast<const char *>(value + n * S * pix + m * kFastD + j * CPL);
    int rowb[4];                                           // bytes between rows, per level (TCE only)
    if (TCE) {
#pragma unroll
        for (int l = 0; l < 4; ++l) rowb[l] = static_cast<int>(__ldg(shapes + 2 * l + 1)) * pixb;
    }

    float acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.f;

    // Points are processed in batches of NB: all 4*NB corner loads of a batch are issued before
    // the first FMA consumes one, so a warp waits on memory 16/NB times instead of 16 times.
    constexpr int NB = MSDA_FWD_BATCH;
#pragma unroll
    for (int p0 = 0; p0 < kFastLP; p0 += NB) {
        PointView pv[NB];
        float v[NB][4][CPL];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int p = p0 + b, src = p / PPL, pp = p % PPL;
            pv[b] = view_point<LPH>(own[pp], src);
            const int rb = TCE ? rowb[p >> 2] : __shfl_sync(kFull, own[pp].W, src, LPH) * pixb;
            const char *r0 = vlane + static_cast<int64_t>(pv[b].base) * pixb;
            const char *r1 = r0 + rb;
#pragma unroll
            for (int k = 0; k < 4; ++k) VecIO<VT>::zero(v[b][k]);
            if (pv[b].mask & 1) VecIO<VT>::load(reinterpret_cast<const VT *>(r0), v[b][0]);
            if (pv[b].mask & 2) VecIO<VT>::load(reinterpret_cast<const VT *>(r0 + pixb), v[b][1]);
            if (pv[b].mask & 4) VecIO<VT>::load(reinterpret_cast<const VT *>(r1), v[b][2]);
            if (pv[b].mask & 8) VecIO<VT>::load(reinterpret_cast<const VT *>(r1 + pixb), v[b][3]);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                acc[c] = fmaf(pv[b].w0, v[b][0][c], acc[c]);
                acc[c] = fmaf(pv[b].w1, v[b][1][c], acc[c]);
                acc[c] = fmaf(pv[b].w2, v[b][2][c], acc[c]);
                acc[c] = fmaf(pv[b].w3, v[b][3][c], acc[c]);
            }
        }
    }

    VecIO<VT>::store_stream(out + (nq * M + m) * kFastD + j * CPL, acc);
}

// --------------------------------------------------------------------------------------------
// Transposing reduction over the LPH lanes of a head, done as a tree so that only a handful of
// partial sums is ever live.  Lane jj owns points jj*PPL + pp.  tree<BIT, PBASE> returns, in every
// lane, the sum over the 2*BIT lanes of its subtree of the 4 corner dot products of the ONE point
// of that subtree the lane owns.  Each internal node exchanges 4 values with the partner lane
// (xor BIT) and keeps 4; leaves gather/scatter one point.  Total: (LP - PPL) * 4 exchanges per
// warp, the same as a full butterfly, but the live set is 4 * log2(LPH) floats instead of 64.
// --------------------------------------------------------------------------------------------
template <int V> struct IntC { static constexpr int value = V; };

template <int BIT, int PBASE, int PPL, int LPH, typename Leaf>
__device__ __forceinline__ void tree_reduce(int j, Leaf &leaf, float (&res)[4])
{
    if constexpr (BIT == 0) {
        leaf(IntC<PBASE>{}, res);
    } else {
        float lo[4], hi[4];
        tree_reduce<BIT / 2, PBASE, PPL, LPH>(j, leaf, lo);
        tree_reduce<BIT / 2, PBASE + BIT * PPL, PPL, LPH>(j, leaf, hi);
        const bool upper = (j & BIT) != 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float send = upper ? lo[c] : hi[c];
            const float keep = upper ? hi[c] : lo[c];
            res[c] = keep + __shfl_xor_sync(kFull, send, BIT, LPH);
        }
    }
}

// --------------------------------------------------------------------------------------------
// grad_value scatter: predicated 16-byte vector reductions (no branch around the instruction).
// GT is the type grad_value is accumulated in: VT itself (fp32 in place, or direct 16-bit vector
// atomics) or float (fp32 scratch for 16-bit VT).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4_f32_if(int pred, void *addr, float a, float b, float c, float d)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
                 "@p red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(pred) : "memory");
}
__device__ __forceinline__ void red_add_v4_bf16x2_if(int pred, void *addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
                 "@p red.relaxed.gpu.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};\n\t}"
                 :: "l"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(pred) : "memory");
}

// CPL = channels a lane holds.  16-bit values accumulate either in an fp32 scratch (GT = float; the
// backward then uses the 8-lanes-per-head layout, CPL = 4, so that every reduction instruction
// carries whole 128-byte lines like the fp32 kernel: with 4 lanes x 8 channels a line was two
// 64-byte requests and the L2 sustained 20 % fewer line updates) or directly in 16 bits
// (GT = VT, 4 lanes x 8 channels, one 64-byte bf16x2 request per line).
template <typename VT, typename GT, int CPL> struct GradScatter;
template <> struct GradScatter<float, float, 4> {
    static __device__ __forceinline__ void add(int pred, char *p, float w, const float (&go)[4])
    {
        red_add_v4_f32_if(pred, p, w * go[0], w * go[1], w * go[2], w * go[3]);
    }
};
template <> struct GradScatter<__nv_bfloat16, float, 4> {
    static __device__ __forceinline__ void add(int pred, char *p, float w, const float (&go)[4])
    {
        red_add_v4_f32_if(pred, p, w * go[0], w * go[1], w * go[2], w * go[3]);
    }
};
template <> struct GradScatter<__nv_bfloat16, __nv_bfloat16, 8> {
    static __device__ __forceinline__ void add(int pred, char *p, float w, const float (&go)[8])
    {
        red_add_v4_bf16x2_if(pred, p, pack_bf16x2(w * go[0], w * go[1]), pack_bf16x2(w * go[2], w * go[3]),
                             pack_bf16x2(w * go[4], w * go[5]), pack_bf16x2(w * go[6], w * go[7]));
    }
};

// lanes per head of the backward kernel for (value type, accumulation type)
template <typename VT, typename GT> struct BwdLanes { static constexpr int LPH = FastCfg<VT>::LPH; };
template <> struct BwdLanes<__nv_bfloat16, float> { static constexpr int LPH = 8; };

// --------------------------------------------------------------------------------------------
// FAST backward.
// --------------------------------------------------------------------------------------------
template <typename VT, typename LT, typename GT, bool PROLOGUE, bool TCE>
__global__ void __launch_bounds__(256, (BwdLanes<VT, GT>::LPH == 8 ? MSDA_BWD_MIN_BLOCKS_F32 : MSDA_BWD_MIN_BLOCKS_BF16))
bwd_fast_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                const int64_t *__restrict__ lsi, const LT *__restrict__ loc,
                const LT *__restrict__ aw, const VT *__restrict__ grad_out,
                GT *__restrict__ grad_value, LT *__restrict__ grad_loc, LT *__restrict__ grad_aw,
                int64_t n_items, int64_t Lq, int64_t S, int M, int log2P, ProArgs pro)
{
    constexpr int LPH = BwdLanes<VT, GT>::LPH, CPL = kFastD / LPH, HPW = kWarp / LPH, PPL = kFastLP / LPH;
    using IO = VecIOn<VT, CPL>;

    const int lane = threadIdx.x & 31;
    const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (TCE) { M = 8; log2P = 2; }
    const int groups = M / HPW;
    int64_t nq;
    int hg;
    if (!map_item<MSDA_QMAP_BWD != 0>(item, n_items, groups, nq, hg)) return;
    const int64_t n = nq / Lq;
    const int g = lane / LPH, j = lane % LPH;
    const int m = hg * HPW + g;

    const int64_t qh0 = nq * M + hg * HPW;
    float xy[2 * PPL], a[PPL];
    PointIO<LT, PPL>::load_loc(loc + qh0 * (kFastLP * 2) + lane * (2 * PPL), xy);
    PointIO<LT, PPL>::load_aw(aw + qh0 * kFastLP + lane * PPL, a);

    float go[CPL];
    IO::load_stream(grad_out + (nq * M + m) * kFastD + j * CPL, go);

    int Hs[PPL], Ws[PPL], starts[PPL], lv[PPL];
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) {
        lv[pp] = (j * PPL + pp) >> log2P;
        level_dims(shapes, lsi, lv[pp], Hs[pp], Ws[pp], starts[pp]);
    }
    ProPoint prop[PPL];
    if (PROLOGUE) prologue_points<LT, LPH, PPL>(xy, a, Hs, Ws, lv, pro, nq, prop);
    OwnedPoint own[PPL];
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp)
        own[pp] = own_point(xy[2 * pp], xy[2 * pp + 1], a[pp], Hs[pp], Ws[pp], starts[pp]);
    MergedPoint mp[PPL];
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) mp[pp] = merged_from(own[pp]);
#if MSDA_MERGE
    if (TCE) merge_level<PPL, LPH>(mp, Ws[0], j);
#endif

    const int pix = M * kFastD;
    const int pixb = pix * static_cast<int>(sizeof(VT));
    const int gpixb = pix * static_cast<int>(sizeof(GT));
    const int64_t head_off = n * S * pix + m * kFastD;
    const char *vlane = reinterpret_cast<const char *>(value + head_off + j * CPL);
    char *glane = reinterpret_cast<char *>(grad_value + head_off + j * CPL);
    int rowW[4];
    if (TCE) {
#pragma unroll
        for (int l = 0; l < 4; ++l) rowW[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));
    }

    // one sampling point: broadcast, gather 4 corners, scatter 4 vector reductions, 4 partial dots
    auto leaf = [&](auto pc, float (&d)[4]) {
        constexpr int p = decltype(pc)::value;
        constexpr int src = p / PPL, pp = p % PPL;
        const MergedPoint pv = view_merged<LPH>(mp[pp], src);
        const int W = TCE ? rowW[p >> 2] : __shfl_sync(kFull, own[pp].W, src, LPH);
        const char *r0 = vlane + static_cast<int64_t>(pv.base()) * pixb;
        const char *r1 = r0 + static_cast<int64_t>(W) * pixb;
        char *g0 = glane + static_cast<int64_t>(pv.base()) * gpixb;
        char *g1 = g0 + static_cast<int64_t>(W) * gpixb;
        float v0[CPL], v1[CPL], v2[CPL], v3[CPL];
        IO::zero(v0); IO::zero(v1); IO::zero(v2); IO::zero(v3);
        // MSDA_EXP 10..13: timing experiments that isolate the two halves of the backward
        // (profiles/r02_bwd_split.md); never product.
#if MSDA_EXP == 12
        constexpr bool kLoads = false;
#elif MSDA_EXP == 13
        constexpr bool kLoads = p < 8;
#else
        constexpr bool kLoads = true;
#endif
#if MSDA_EXP == 10
        constexpr bool kReds = false;
#elif MSDA_EXP == 11 || MSDA_EXP == 13
        constexpr bool kReds = p < 8;
#else
        constexpr bool kReds = true;
#endif
        if (kLoads) {
            // every corner's row is read (its dot product feeds the point's own gradients) ...
            if (pv.lmask() & 1) IO::load(reinterpret_cast<const VT *>(r0), v0);
            if (pv.lmask() & 2) IO::load(reinterpret_cast<const VT *>(r0 + pixb), v1);
            if (pv.lmask() & 4) IO::load(reinterpret_cast<const VT *>(r1), v2);
            if (pv.lmask() & 8) IO::load(reinterpret_cast<const VT *>(r1 + pixb), v3);
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c) { v0[c] = pv.w[0]; v1[c] = pv.w[1]; v2[c] = pv.w[2]; v3[c] = pv.w[3]; }
        }
        if (kReds) {
            // ... but a pixel shared with an earlier point of the level is reduced once, with the summed weight
            GradScatter<VT, GT, CPL>::add(pv.amask() & 1, g0, pv.w[0], go);
            GradScatter<VT, GT, CPL>::add(pv.amask() & 2, g0 + gpixb, pv.w[1], go);
            GradScatter<VT, GT, CPL>::add(pv.amask() & 4, g1, pv.w[2], go);
            GradScatter<VT, GT, CPL>::add(pv.amask() & 8, g1 + gpixb, pv.w[3], go);
        }
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            d0 = fmaf(go[c], v0[c], d0);
            d1 = fmaf(go[c], v1[c], d1);
            d2 = fmaf(go[c], v2[c], d2);
            d3 = fmaf(go[c], v3[c], d3);
        }
        d[0] = d0; d[1] = d1; d[2] = d2; d[3] = d3;
    };

    // lane ends up with the full channel dot products of its own PPL points
    float gxy[2 * PPL], ga[PPL];
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) {
        float d[4];
        if (pp == 0) tree_reduce<LPH / 2, 0, PPL, LPH>(j, leaf, d);
        if (pp == 1) tree_reduce<LPH / 2, 1, PPL, LPH>(j, leaf, d);
        if (PPL > 2) {
            if (pp == 2) tree_reduce<LPH / 2, 2 % PPL, PPL, LPH>(j, leaf, d);
            if (pp == 3) tree_reduce<LPH / 2, 3 % PPL, PPL, LPH>(j, leaf, d);
        }
        const OwnedPoint &o = own[pp];
        const float lh = o.lh, lw = o.lw, hh = 1.f - lh, hw = 1.f - lw;
        ga[pp] = hh * hw * d[0] + hh * lw * d[1] + lh * hw * d[2] + lh * lw * d[3];
        const float gw = hh * (d[1] - d[0]) + lh * (d[3] - d[2]);
        const float gh = hw * (d[2] - d[0]) + lw * (d[3] - d[1]);
        gxy[2 * pp]     = static_cast<float>(o.W) * gw * o.a;
        gxy[2 * pp + 1] = static_cast<float>(o.H) * gh * o.a;
    }
    if (PROLOGUE) {
        // chain rule through softmax: g_logit_i = a_i * (g_a_i - sum_k a_k g_a_k)
        float dot = 0.f;
#pragma unroll
        for (int pp = 0; pp < PPL; ++pp) dot = fmaf(a[pp], ga[pp], dot);
        dot = group_sum<LPH>(dot);
#pragma unroll
        for (int pp = 0; pp < PPL; ++pp) ga[pp] = a[pp] * (ga[pp] - dot);
        // chain rule through loc = ref + f(offset, ref)
#pragma unroll
        for (int pp = 0; pp < PPL; ++pp) {
            const float gx = gxy[2 * pp], gy = gxy[2 * pp + 1];
            float grw = 0.f, grh = 0.f;
            if (pro.ref_dim == 2) {
                gxy[2 * pp] = gx / prop[pp].sx;
                gxy[2 * pp + 1] = gy / prop[pp].sy;
            } else {
                gxy[2 * pp] = gx * 0.5f * prop[pp].sx / pro.n_points;
                gxy[2 * pp + 1] = gy * 0.5f * prop[pp].sy / pro.n_points;
                grw = gx * 0.5f * (prop[pp].ox / pro.n_points);
                grh = gy * 0.5f * (prop[pp].oy / pro.n_points);
            }
            if (pro.grad_ref) {
                float *gr = pro.grad_ref + (nq * pro.n_levels + lv[pp]) * pro.ref_dim;
                atomicAdd(gr, gx);
                atomicAdd(gr + 1, gy);
                if (pro.ref_dim == 4) { atomicAdd(gr + 2, grw); atomicAdd(gr + 3, grh); }
            }
        }
    }
    PointIO<LT, PPL>::store_loc(grad_loc + qh0 * (kFastLP * 2) + lane * (2 * PPL), gxy);
    PointIO<LT, PPL>::store_aw(grad_aw + qh0 * kFastLP + lane * PPL, ga);
}

// --------------------------------------------------------------------------------------------
// 16-bit grad_value epilogue: round the fp32 scratch once.
// --------------------------------------------------------------------------------------------
template <typename VT>
__global__ void __launch_bounds__(256)
cast_scratch_kernel(const float *__restrict__ src, VT *__restrict__ dst, int64_t n_vec8)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec8; i += stride) {
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        const float4 a = __ldcs(s4 + 2 * i);
        const float4 b = __ldcs(s4 + 2 * i + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        VT o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = cvt_from<VT, float>(v[k]);
        *(reinterpret_cast<uint4 *>(dst) + i) = *reinterpret_cast<const uint4 *>(o);
    }
}

template <typename VT>
__global__ void cast_scratch_tail_kernel(const float *__restrict__ src, VT *__restrict__ dst,
                                         int64_t begin, int64_t end)
{
    const int64_t i = begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < end) dst[i] = cvt_from<VT, float>(src[i]);
}

// --------------------------------------------------------------------------------------------
// GENERIC forward: warp per (frame, query, head); lanes strided over channels.
// THREE_D adds the fork's cross-frame blend (ms_deform_im2col_cuda.cuh:366-420).
// --------------------------------------------------------------------------------------------
template <typename VT, typename LT, bool THREE_D>
__global__ void __launch_bounds__(256)
fwd_generic_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                   const int64_t *__restrict__ lsi, const LT *__restrict__ loc,
                   const LT *__restrict__ aw, VT *__restrict__ out,
                   int64_t n_qh, int64_t N, int64_t Lq, int64_t S, int M, int D, int L, int P)
{
    using A = typename AccOf<VT>::type;
    constexpr int LD = THREE_D ? 3 : 2;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t pix = static_cast<int64_t>(M) * D;

    for (int64_t qh = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; qh < n_qh; qh += warps) {
        const int m = static_cast<int>(qh % M);
        const int64_t n = (qh / M) / Lq;
        const LT *lp = loc + qh * (static_cast<int64_t>(L) * P * LD);
        const LT *ap = aw + qh * (static_cast<int64_t>(L) * P);
        for (int c = lane; c < D; c += kWarp) {
            A acc = 0;
            for (int l = 0; l < L; ++l) {
                const int H = static_cast<int>(__ldg(shapes + 2 * l));
                const int W = static_cast<int>(__ldg(shapes + 2 * l + 1));
                const int64_t start = __ldg(lsi + l);
                for (int p = 0; p < P; ++p) {
                    const int k = l * P + p;
                    const A x = cvt_to<A>(lp[k * LD]), y = cvt_to<A>(lp[k * LD + 1]);
                    const A wgt = cvt_to<A>(ap[k]);
                    const PointGeom<A> g = resolve_point<A>(x, y, H, W);
                    if (!g.mask) continue;
                    const A hh = 1 - g.lh, hw = 1 - g.lw;
                    const A cw[4] = {hh * hw, hh * g.lw, g.lh * hw, g.lh * g.lw};
                    const int64_t o00 = (start + static_cast<int64_t>(g.h_low) * W + g.w_low) * pix + m * D + c;
                    const int64_t co[4] = {o00, o00 + pix, o00 + W * pix, o00 + W * pix + pix};
                    if (!THREE_D) {
                        const VT *vf = value + n * S * pix;
                        A s = 0;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            if (g.mask & (1 << k4)) s += cw[k4] * cvt_to<A>(vf[co[k4]]);
                        acc += s * wgt;
                    } else {
                        const A f_im = cvt_to<A>(lp[k * LD + 2]) * static_cast<A>(N) - static_cast<A>(0.5);
                        const A ff = floor_acc(f_im);
                        // frame index stays in range of int for any finite location the model produces
                        const int64_t f_low = static_cast<int64_t>(ff), f_high = f_low + 1;
                        A lo = 0, hi = 0;
                        if (f_low >= 0 && f_low < N) {
                            const VT *vf = value + f_low * S * pix;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                if (g.mask & (1 << k4)) lo += cw[k4] * cvt_to<A>(vf[co[k4]]);
                        }
                        if (f_high >= 0 && f_high < N) {
                            const VT *vf = value + f_high * S * pix;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                if (g.mask & (1 << k4)) hi += cw[k4] * cvt_to<A>(vf[co[k4]]);
                        }
                        acc += (lo * (static_cast<A>(f_high) - f_im) + hi * (f_im - static_cast<A>(f_low))) * wgt;
                    }
                }
            }
            out[qh * D + c] = cvt_from<VT, A>(acc);
        }
    }
}

template <typename A>
__device__ __forceinline__ A warp_sum(A v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// --------------------------------------------------------------------------------------------
// GENERIC backward (2-D and the fork's 3-D variant, ms_deform_im2col_cuda.cuh:162-242, :565-712).
// GT = accumulation type of grad_value (VT, or float scratch for 16-bit VT).
// --------------------------------------------------------------------------------------------
template <typename VT, typename LT, typename GT, bool THREE_D>
__global__ void __launch_bounds__(256)
bwd_generic_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                   const int64_t *__restrict__ lsi, const LT *__restrict__ loc,
                   const LT *__restrict__ aw, const VT *__restrict__ grad_out,
                   GT *__restrict__ grad_value, LT *__restrict__ grad_loc, LT *__restrict__ grad_aw,
                   int64_t n_qh, int64_t N, int64_t Lq, int64_t S, int M, int D, int L, int P)
{
    using A = typename AccOf<VT>::type;
    constexpr int LD = THREE_D ? 3 : 2;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t pix = static_cast<int64_t>(M) * D;

    for (int64_t qh = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; qh < n_qh; qh += warps) {
        const int m = static_cast<int>(qh % M);
        const int64_t n = (qh / M) / Lq;
        const LT *lp = loc + qh * (static_cast<int64_t>(L) * P * LD);
        const LT *ap = aw + qh * (static_cast<int64_t>(L) * P);
        LT *glp = grad_loc + qh * (static_cast<int64_t>(L) * P * LD);
        LT *gap = grad_aw + qh * (static_cast<int64_t>(L) * P);
        const VT *gop = grad_out + qh * D;

        for (int l = 0; l < L; ++l) {
            const int H = static_cast<int>(__ldg(shapes + 2 * l));
            const int W = static_cast<int>(__ldg(shapes + 2 * l + 1));
            const int64_t start = __ldg(lsi + l);
            for (int p = 0; p < P; ++p) {
                const int k = l * P + p;
                const A x = cvt_to<A>(lp[k * LD]), y = cvt_to<A>(lp[k * LD + 1]);
                const A wgt = cvt_to<A>(ap[k]);
                const PointGeom<A> g = resolve_point<A>(x, y, H, W);
                A s_a = 0, s_w = 0, s_h = 0, s_f = 0;
                if (g.mask) {
                    const A lh = g.lh, lw = g.lw, hh = 1 - lh, hw = 1 - lw;
                    const A cw[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
                    // d(weight)/d(w_im), d(weight)/d(h_im) per corner
                    const A dw[4] = {-hh, hh, -lh, lh};
                    const A dh[4] = {-hw, -lw, hw, lw};
                    const int64_t o00 = (start + static_cast<int64_t>(g.h_low) * W + g.w_low) * pix + m * D;
                    const int64_t co[4] = {o00, o00 + pix, o00 + W * pix, o00 + W * pix + pix};
                    if (!THREE_D) {
                        const int64_t fo = n * S * pix;
                        for (int c = lane; c < D; c += kWarp) {
                            const A go = cvt_to<A>(gop[c]);
                            const A gv = go * wgt;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                if (g.mask & (1 << k4)) {
                                    const A v = cvt_to<A>(value[fo + co[k4] + c]);
                                    s_a += go * cw[k4] * v;
                                    s_w += gv * dw[k4] * v;
                                    s_h += gv * dh[k4] * v;
                                    atomic_add_any(grad_value + fo + co[k4] + c, cvt_from<GT, A>(cw[k4] * gv));
                                }
                            }
                        }
                    } else {
                        const A f_im = cvt_to<A>(lp[k * LD + 2]) * static_cast<A>(N) - static_cast<A>(0.5);
                        const int64_t f_low = static_cast<int64_t>(floor_acc(f_im)), f_high = f_low + 1;
                        const A lf = f_im - static_cast<A>(f_low), hf = 1 - lf;
                        for (int side = 0; side < 2; ++side) {
                            const int64_t f = side ? f_high : f_low;
                            if (f < 0 || f >= N) continue;
                            const A fw = side ? lf : hf;          // blend factor of this frame
                            const A fs = side ? A(1) : A(-1);     // d(fw)/d(f_im)
                            const int64_t fo = f * S * pix;
                            for (int c = lane; c < D; c += kWarp) {
                                const A go = cvt_to<A>(gop[c]);
                                const A gv = go * wgt;
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4) {
                                    if (g.mask & (1 << k4)) {
                                        const A v = cvt_to<A>(value[fo + co[k4] + c]);
                                        s_a += go * cw[k4] * fw * v;
                                        s_w += gv * dw[k4] * fw * v;
                                        s_h += gv * dh[k4] * fw * v;
                                        s_f += gv * fs * cw[k4] * v;
                                        atomic_add_any(grad_value + fo + co[k4] + c, cvt_from<GT, A>(cw[k4] * fw * gv));
                                    }
                                }
                            }
                        }
                    }
                }
                s_a = warp_sum(s_a); s_w = warp_sum(s_w); s_h = warp_sum(s_h);
                if (THREE_D) s_f = warp_sum(s_f);
                if (lane == 0) {
                    gap[k] = cvt_from<LT, A>(s_a);
                    glp[k * LD] = cvt_from<LT, A>(static_cast<A>(W) * s_w);
                    glp[k * LD + 1] = cvt_from<LT, A>(static_cast<A>(H) * s_h);
                    if (THREE_D) glp[k * LD + 2] = cvt_from<LT, A>(static_cast<A>(N - 1) * s_f);
                }
            }
        }
    }
}

}  // namespace msda
