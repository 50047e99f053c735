// msda_capi.cu — the C-ABI shared library (libmsda_b200.so) declared in include/msda_b200.h.
// Host-side dispatch only; all device code is in msda_kernels.cuh / msda_prologue.cuh.
// No torch types, no device synchronisation in the device-pointer entry points.

#include "../../include/msda_b200.h"
#include "msda_kernels.cuh"

#include <atomic>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

std::atomic<int64_t> g_launches{0};

inline size_t dtype_size(int dt)
{
    switch (dt) {
        case MSDA_F32: return 4;
        case MSDA_BF16: return 2;
        case MSDA_F16: return 2;
        case MSDA_F64: return 8;
        default: return 0;
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int ilog2_exact(int v)
{
    int l = 0;
    while ((1 << l) < v) ++l;
    return ((1 << l) == v) ? l : -1;
}

struct Dims {
    int64_t N, S, Lq;
    int M, D, L, P;
    int vdt, ldt;
};

int check_shape(const msda_shape *s, Dims &d)
{
    if (!s) return MSDA_E_NULL;
    d = Dims{s->n_frames, s->spatial_size, s->n_queries, s->n_heads, s->channels,
             s->n_levels, s->n_points, s->value_dtype, s->loc_dtype};
    if (d.N < 0 || d.Lq < 0 || d.S <= 0 || d.M <= 0 || d.D <= 0 || d.L <= 0 || d.P <= 0) return MSDA_E_SHAPE;
    // index model: pixel indices and per-(query,head) point counts are 32-bit, everything that
    // scales with N is 64-bit (the reference overflows int32 at N>=960, level-set C: SURVEY 8a a3)
    if (d.S > (int64_t(1) << 27)) return MSDA_E_SHAPE;
    if (int64_t(d.M) * d.D > (int64_t(1) << 20)) return MSDA_E_SHAPE;
    if (int64_t(d.L) * d.P > (int64_t(1) << 16)) return MSDA_E_SHAPE;
    const bool ok =
        (d.vdt == MSDA_F32 && d.ldt == MSDA_F32) || (d.vdt == MSDA_F64 && d.ldt == MSDA_F64) ||
        (d.vdt == MSDA_BF16 && (d.ldt == MSDA_F32 || d.ldt == MSDA_BF16)) ||
        (d.vdt == MSDA_F16 && (d.ldt == MSDA_F32 || d.ldt == MSDA_F16));
    if (!ok) return MSDA_E_DTYPE;
    return MSDA_OK;
}

inline bool fast_eligible(const Dims &d)
{
    if (d.D != msda::kFastD || d.L * d.P != msda::kFastLP) return false;
    // offsets inside one frame are 32-bit byte offsets (also for the fp32 scratch of bf16 grads)
    if (d.S * d.M * d.D * 4 >= (int64_t(1) << 31)) return false;
    // a point travels between lanes as (pixel index << 8 | corner masks) in one 32-bit word
    if (d.S >= (int64_t(1) << 22)) return false;
    if (d.vdt == MSDA_F32 && d.ldt == MSDA_F32) return d.M % 4 == 0;
    if (d.vdt == MSDA_BF16 && (d.ldt == MSDA_F32 || d.ldt == MSDA_BF16)) return d.M % 8 == 0;
    return false;
}

inline int launch_status()
{
    const cudaError_t e = cudaGetLastError();
    return static_cast<int>(e);
}

inline unsigned grid_for_warps(int64_t warps, int threads)
{
    const int64_t wpb = threads / 32;
    return static_cast<unsigned>((warps + wpb - 1) / wpb);
}

// grid of the fast kernels: with the query-major item map (MSDA_QMAP_*) the queries are padded to whole CTAs per head group
inline unsigned grid_for_items(int64_t queries, int groups, int threads, bool qmap)
{
    const int64_t wpb = threads / 32;
    if (!qmap) return grid_for_warps(queries * groups, threads);
    return static_cast<unsigned>(((queries + wpb - 1) / wpb) * groups);
}

int sm_count_aux()
{
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// grid for the grid-stride generic kernels: enough CTAs to fill the machine a few times over
unsigned generic_grid(int64_t warps, int threads)
{
    const int sms = sm_count_aux();
    const int64_t need = (warps + threads / 32 - 1) / (threads / 32);
    return static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sms) * 32)));
}

// ---- forward ------------------------------------------------------------------------------

template <typename VT, typename LT, bool PRO = false>
int launch_fwd_fast(const Dims &d, const void *value, const int64_t *shapes, const int64_t *lsi,
                    const void *loc, const void *aw, void *out, cudaStream_t st,
                    const msda::ProArgs &pro = msda::ProArgs{})
{
    constexpr int HPW = 32 / msda::FastCfg<VT>::LPH;
    const int64_t items = d.N * d.Lq * (d.M / HPW);
    if (items == 0) return MSDA_OK;
    constexpr int threads = 256;
    auto kern = (d.M == 8 && d.P == 4) ? msda::fwd_fast_kernel<VT, LT, PRO, true> : msda::fwd_fast_kernel<VT, LT, PRO, false>;
    kern<<<grid_for_items(d.N * d.Lq, d.M / HPW, threads, MSDA_QMAP_FWD != 0), threads, 0, st>>>(
        static_cast<const VT *>(value), shapes, lsi, static_cast<const LT *>(loc),
        static_cast<const LT *>(aw), static_cast<VT *>(out), items, d.Lq, d.S, d.M, ilog2_exact(d.P), pro);
    ++g_launches;
    return launch_status();
}

template <typename VT, typename LT, bool THREE_D>
int launch_fwd_generic(const Dims &d, int64_t group, const void *value, const int64_t *shapes,
                       const int64_t *lsi, const void *loc, const void *aw, void *out, cudaStream_t st)
{
    constexpr int threads = 256;
    if (!THREE_D) {
        const int64_t qh = d.N * d.Lq * d.M;
        if (qh == 0) return MSDA_OK;
        msda::fwd_generic_kernel<VT, LT, false><<<generic_grid(qh, threads), threads, 0, st>>>(
            static_cast<const VT *>(value), shapes, lsi, static_cast<const LT *>(loc),
            static_cast<const LT *>(aw), static_cast<VT *>(out), qh, d.N, d.Lq, d.S, d.M, d.D, d.L, d.P);
        ++g_launches;
        return launch_status();
    }
    // 3-D: the temporal blend is confined to groups of `group` frames (the reference's
    // im2col_step chunk, ms_deform_attn_cuda.cu:54,77): one launch per group.
    const size_t ve = sizeof(VT), le = sizeof(LT);
    const int64_t pix = int64_t(d.M) * d.D;
    for (int64_t f0 = 0; f0 < d.N; f0 += group) {
        const int64_t nf = std::min(group, d.N - f0);
        const int64_t qh = nf * d.Lq * d.M;
        if (qh == 0) continue;
        msda::fwd_generic_kernel<VT, LT, true><<<generic_grid(qh, threads), threads, 0, st>>>(
            reinterpret_cast<const VT *>(static_cast<const char *>(value) + f0 * d.S * pix * ve), shapes, lsi,
            reinterpret_cast<const LT *>(static_cast<const char *>(loc) + f0 * d.Lq * d.M * d.L * d.P * 3 * le),
            reinterpret_cast<const LT *>(static_cast<const char *>(aw) + f0 * d.Lq * d.M * d.L * d.P * le),
            reinterpret_cast<VT *>(static_cast<char *>(out) + f0 * d.Lq * pix * ve),
            qh, nf, d.Lq, d.S, d.M, d.D, d.L, d.P);
        ++g_launches;
        const int e = launch_status();
        if (e) return e;
    }
    return MSDA_OK;
}

template <bool THREE_D>
int dispatch_fwd_generic(const Dims &d, int64_t group, const void *value, const int64_t *shapes,
                         const int64_t *lsi, const void *loc, const void *aw, void *out, cudaStream_t st)
{
#define MSDA_CASE(VDT, LDT, VT, LT) \
    if (d.vdt == VDT && d.ldt == LDT) return launch_fwd_generic<VT, LT, THREE_D>(d, group, value, shapes, lsi, loc, aw, out, st);
    MSDA_CASE(MSDA_F32, MSDA_F32, float, float)
    MSDA_CASE(MSDA_F64, MSDA_F64, double, double)
    MSDA_CASE(MSDA_BF16, MSDA_F32, __nv_bfloat16, float)
    MSDA_CASE(MSDA_BF16, MSDA_BF16, __nv_bfloat16, __nv_bfloat16)
    MSDA_CASE(MSDA_F16, MSDA_F32, __half, float)
    MSDA_CASE(MSDA_F16, MSDA_F16, __half, __half)
#undef MSDA_CASE
    return MSDA_E_DTYPE;
}

// ---- backward -----------------------------------------------------------------------------

template <typename VT>
int launch_cast(const float *scratch, void *grad_value, int64_t n, cudaStream_t st)
{
    const int64_t n8 = n / 8;
    if (n8 > 0) {
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n8 + 255) / 256, 148 * 16));
        msda::cast_scratch_kernel<VT><<<grid, 256, 0, st>>>(scratch, static_cast<VT *>(grad_value), n8);
        ++g_launches;
        const int e = launch_status();
        if (e) return e;
    }
    if (n8 * 8 < n) {
        msda::cast_scratch_tail_kernel<VT><<<1, 32, 0, st>>>(scratch, static_cast<VT *>(grad_value), n8 * 8, n);
        ++g_launches;
        return launch_status();
    }
    return MSDA_OK;
}

template <typename VT, typename LT, typename GT, bool PRO = false>
int launch_bwd_fast(const Dims &d, const void *value, const int64_t *shapes, const int64_t *lsi,
                    const void *loc, const void *aw, const void *go, GT *gv, void *gl, void *ga,
                    cudaStream_t st, const msda::ProArgs &pro = msda::ProArgs{})
{
    constexpr int HPW = 32 / msda::BwdLanes<VT, GT>::LPH;
    const int64_t items = d.N * d.Lq * (d.M / HPW);
    if (items == 0) return MSDA_OK;
    constexpr int threads = 256;
    auto kern = (d.M == 8 && d.P == 4) ? msda::bwd_fast_kernel<VT, LT, GT, PRO, true> : msda::bwd_fast_kernel<VT, LT, GT, PRO, false>;
    kern<<<grid_for_items(d.N * d.Lq, d.M / HPW, threads, MSDA_QMAP_BWD != 0), threads, 0, st>>>(
        static_cast<const VT *>(value), shapes, lsi, static_cast<const LT *>(loc),
        static_cast<const LT *>(aw), static_cast<const VT *>(go), gv, static_cast<LT *>(gl),
        static_cast<LT *>(ga), items, d.Lq, d.S, d.M, ilog2_exact(d.P), pro);
    ++g_launches;
    return launch_status();
}

template <typename VT, typename LT, typename GT, bool THREE_D>
int launch_bwd_generic(const Dims &d, int64_t group, const void *value, const int64_t *shapes,
                       const int64_t *lsi, const void *loc, const void *aw, const void *go, GT *gv,
                       void *gl, void *ga, cudaStream_t st)
{
    constexpr int threads = 256;
    constexpr int LD = THREE_D ? 3 : 2;
    const int64_t pix = int64_t(d.M) * d.D;
    const int64_t step = THREE_D ? group : std::max<int64_t>(d.N, 1);
    for (int64_t f0 = 0; f0 < d.N; f0 += step) {
        const int64_t nf = std::min(step, d.N - f0);
        const int64_t qh = nf * d.Lq * d.M;
        if (qh == 0) continue;
        const int64_t pts = f0 * d.Lq * d.M * d.L * d.P;
        msda::bwd_generic_kernel<VT, LT, GT, THREE_D><<<generic_grid(qh, threads), threads, 0, st>>>(
            static_cast<const VT *>(value) + f0 * d.S * pix, shapes, lsi,
            static_cast<const LT *>(loc) + pts * LD, static_cast<const LT *>(aw) + pts,
            static_cast<const VT *>(go) + f0 * d.Lq * pix, gv + f0 * d.S * pix,
            static_cast<LT *>(gl) + pts * LD, static_cast<LT *>(ga) + pts,
            qh, nf, d.Lq, d.S, d.M, d.D, d.L, d.P);
        ++g_launches;
        const int e = launch_status();
        if (e) return e;
    }
    return MSDA_OK;
}

template <bool THREE_D>
int backward_impl(const msda_shape *shape, int64_t group, const void *value, const int64_t *shapes,
                  const int64_t *lsi, const void *loc, const void *aw, const void *go,
                  void *gv, void *gl, void *ga, void *ws, size_t ws_bytes, void *stream)
{
    Dims d;
    int e = check_shape(shape, d);
    if (e) return e;
    if (!value || !shapes || !lsi || !loc || !aw || !go || !gv || !gl || !ga) {
        if (d.N * d.Lq == 0 && value && gv) { /* degenerate: fall through to the memset */ }
        else return MSDA_E_NULL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nval = d.N * d.S * d.M * d.D;
    const bool sixteen = (d.vdt == MSDA_BF16 || d.vdt == MSDA_F16);
    const bool use_ws = sixteen && ws != nullptr;
    if (use_ws && ws_bytes < static_cast<size_t>(nval) * sizeof(float)) return MSDA_E_WORKSPACE;

    // the only zero-fill pass: grad_value (or its fp32 scratch).  grad_loc / grad_aw are
    // written exactly once by the kernels (the reference memsets all three, .cu:145-147).
    cudaError_t ce;
    if (use_ws) ce = cudaMemsetAsync(ws, 0, static_cast<size_t>(nval) * sizeof(float), st);
    else        ce = cudaMemsetAsync(gv, 0, static_cast<size_t>(nval) * dtype_size(d.vdt), st);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    if (d.N * d.Lq == 0) return MSDA_OK;

    const bool fast = !THREE_D && fast_eligible(d) && aligned16(value) && aligned16(loc) && aligned16(aw) &&
                      aligned16(go) && aligned16(gv) && aligned16(gl) && aligned16(ga) && (!use_ws || aligned16(ws));
    if (fast) {
        if (d.vdt == MSDA_F32)
            return launch_bwd_fast<float, float, float>(d, value, shapes, lsi, loc, aw, go, static_cast<float *>(gv), gl, ga, st);
        if (d.vdt == MSDA_BF16) {
            if (use_ws) {
                e = (d.ldt == MSDA_F32)
                        ? launch_bwd_fast<__nv_bfloat16, float, float>(d, value, shapes, lsi, loc, aw, go, static_cast<float *>(ws), gl, ga, st)
                        : launch_bwd_fast<__nv_bfloat16, __nv_bfloat16, float>(d, value, shapes, lsi, loc, aw, go, static_cast<float *>(ws), gl, ga, st);
                if (e) return e;
                return launch_cast<__nv_bfloat16>(static_cast<const float *>(ws), gv, nval, st);
            }
            return (d.ldt == MSDA_F32)
                       ? launch_bwd_fast<__nv_bfloat16, float, __nv_bfloat16>(d, value, shapes, lsi, loc, aw, go, static_cast<__nv_bfloat16 *>(gv), gl, ga, st)
                       : launch_bwd_fast<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(d, value, shapes, lsi, loc, aw, go, static_cast<__nv_bfloat16 *>(gv), gl, ga, st);
        }
    }

#define MSDA_CASE(VDT, LDT, VT, LT)                                                                                   \
    if (d.vdt == VDT && d.ldt == LDT) {                                                                               \
        if (use_ws) {                                                                                                 \
            e = launch_bwd_generic<VT, LT, float, THREE_D>(d, group, value, shapes, lsi, loc, aw, go,                 \
                                                           static_cast<float *>(ws), gl, ga, st);                     \
            if (e) return e;                                                                                          \
            return launch_cast<VT>(static_cast<const float *>(ws), gv, nval, st);                                     \
        }                                                                                                             \
        return launch_bwd_generic<VT, LT, VT, THREE_D>(d, group, value, shapes, lsi, loc, aw, go,                     \
                                                       static_cast<VT *>(gv), gl, ga, st);                            \
    }
    MSDA_CASE(MSDA_BF16, MSDA_F32, __nv_bfloat16, float)
    MSDA_CASE(MSDA_BF16, MSDA_BF16, __nv_bfloat16, __nv_bfloat16)
    MSDA_CASE(MSDA_F16, MSDA_F32, __half, float)
    MSDA_CASE(MSDA_F16, MSDA_F16, __half, __half)
#undef MSDA_CASE
    if (d.vdt == MSDA_F32)
        return launch_bwd_generic<float, float, float, THREE_D>(d, group, value, shapes, lsi, loc, aw, go, static_cast<float *>(gv), gl, ga, st);
    if (d.vdt == MSDA_F64)
        return launch_bwd_generic<double, double, double, THREE_D>(d, group, value, shapes, lsi, loc, aw, go, static_cast<double *>(gv), gl, ga, st);
    return MSDA_E_DTYPE;
}

}  // namespace

// =============================================================================================
// exported C ABI
// =============================================================================================
extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

const char *msda_error_string(int code)
{
    switch (code) {
        case MSDA_OK: return "ok";
        case MSDA_E_NULL: return "msda: a required pointer is NULL";
        case MSDA_E_SHAPE: return "msda: invalid or too large dimension";
        case MSDA_E_DTYPE: return "msda: unsupported dtype combination";
        case MSDA_E_ALIGN: return "msda: pointer is not aligned for its element type";
        case MSDA_E_WORKSPACE: return "msda: workspace too small";
        case MSDA_E_UNSUPPORTED: return "msda: feature not supported for these arguments";
        case MSDA_E_IM2COL_STEP: return "msda: batch must be divisible by min(batch, im2col_step)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "msda: unknown error";
}

const char *msda_kernel_name(const msda_shape *shape, int backward)
{
    Dims d;
    if (check_shape(shape, d)) return "invalid";
    if (fast_eligible(d)) {
        if (d.vdt == MSDA_F32) return backward ? "bwd_fast_d32_f32" : "fwd_fast_d32_f32";
        return backward ? "bwd_fast_d32_bf16" : "fwd_fast_d32_bf16";
    }
    return backward ? "bwd_generic" : "fwd_generic";
}

int64_t msda_algorithmic_bytes(const msda_shape *shape, int pass)
{
    Dims d;
    if (check_shape(shape, d)) return -1;
    const int64_t ev = dtype_size(d.vdt), el = dtype_size(d.ldt);
    const int64_t val = d.N * d.S * d.M * d.D * ev;
    const int64_t outb = d.N * d.Lq * d.M * d.D * ev;
    const int64_t locb = d.N * d.Lq * d.M * d.L * d.P * 2 * el;
    const int64_t awb = d.N * d.Lq * d.M * d.L * d.P * el;
    const int64_t fwd = val + locb + awb + outb;                 // e*N*(256 S + 640 Lq) at TCE shape
    const int64_t bwd = 2 * val + 2 * locb + 2 * awb + outb;     // e*N*(512 S + 1024 Lq)
    if (pass == 0) return fwd;
    if (pass == 1) return bwd;
    return fwd + bwd;
}

int msda_check_im2col_step(int64_t n_frames, int im2col_step)
{
    if (n_frames <= 0) return MSDA_OK;
    if (im2col_step <= 0) return MSDA_E_IM2COL_STEP;
    const int64_t step = std::min<int64_t>(n_frames, im2col_step);
    return (n_frames % step == 0) ? MSDA_OK : MSDA_E_IM2COL_STEP;
}

int64_t msda_launch_count(void) { return g_launches.load(); }

int msda_forward(const msda_shape *shape, const void *value, const int64_t *shapes, const int64_t *lsi,
                 const void *loc, const void *aw, void *out, void *stream)
{
    Dims d;
    const int e = check_shape(shape, d);
    if (e) return e;
    if (d.N * d.Lq == 0) return MSDA_OK;
    if (!value || !shapes || !lsi || !loc || !aw || !out) return MSDA_E_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (fast_eligible(d) && aligned16(value) && aligned16(loc) && aligned16(aw) && aligned16(out)) {
        if (d.vdt == MSDA_F32) return launch_fwd_fast<float, float>(d, value, shapes, lsi, loc, aw, out, st);
        if (d.ldt == MSDA_F32) return launch_fwd_fast<__nv_bfloat16, float>(d, value, shapes, lsi, loc, aw, out, st);
        return launch_fwd_fast<__nv_bfloat16, __nv_bfloat16>(d, value, shapes, lsi, loc, aw, out, st);
    }
    return dispatch_fwd_generic<false>(d, d.N, value, shapes, lsi, loc, aw, out, st);
}

size_t msda_backward_workspace_bytes(const msda_shape *shape)
{
    Dims d;
    if (check_shape(shape, d)) return 0;
    if (d.vdt == MSDA_BF16 || d.vdt == MSDA_F16) return static_cast<size_t>(d.N * d.S * d.M * d.D) * sizeof(float);
    return 0;
}

int msda_backward(const msda_shape *shape, const void *value, const int64_t *shapes, const int64_t *lsi,
                  const void *loc, const void *aw, const void *go, void *gv, void *gl, void *ga,
                  void *ws, size_t ws_bytes, void *stream)
{
    return backward_impl<false>(shape, 0, value, shapes, lsi, loc, aw, go, gv, gl, ga, ws, ws_bytes, stream);
}

int msda_forward_3d(const msda_shape *shape, int64_t frames_per_group, const void *value, const int64_t *shapes,
                    const int64_t *lsi, const void *loc, const void *aw, void *out, void *stream)
{
    Dims d;
    const int e = check_shape(shape, d);
    if (e) return e;
    if (d.N * d.Lq == 0) return MSDA_OK;
    if (!value || !shapes || !lsi || !loc || !aw || !out) return MSDA_E_NULL;
    if (frames_per_group <= 0) return MSDA_E_SHAPE;
    return dispatch_fwd_generic<true>(d, std::min(frames_per_group, d.N), value, shapes, lsi, loc, aw, out,
                                      static_cast<cudaStream_t>(stream));
}

int msda_backward_3d(const msda_shape *shape, int64_t frames_per_group, const void *value, const int64_t *shapes,
                     const int64_t *lsi, const void *loc, const void *aw, const void *go, void *gv, void *gl,
                     void *ga, void *ws, size_t ws_bytes, void *stream)
{
    if (frames_per_group <= 0) return MSDA_E_SHAPE;
    const int64_t group = shape ? std::min<int64_t>(frames_per_group, std::max<int64_t>(shape->n_frames, 1)) : 1;
    return backward_impl<true>(shape, group, value, shapes, lsi, loc, aw, go, gv, gl, ga, ws, ws_bytes, stream);
}

}  // extern "C"

#include "msda_prologue_host.inl"
#include "msda_host_pipeline.inl"
#include "msda_aux_host.inl"
