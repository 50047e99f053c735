// msda_host_pipeline.inl — host-buffer entry points (msda_*_host).
// Included at the end of msda_capi.cu.
//
// Frames are independent (each frame's queries read only that frame's value rows), so a call on
// HOST tensors is cut into frame chunks and run as a three-stage pipeline on three streams:
//   h2d stream : chunk c inputs  host -> device
//   run stream : kernels of chunk c (waits on chunk c's copy-in event)
//   d2h stream : chunk c results device -> host (waits on chunk c's kernel event)
// PCIe is full duplex, so copy-in of chunk c+1, kernels of chunk c and copy-out of chunk c-1
// overlap.  Device staging is owned by the context and only ever grows.

struct msda_host_ctx {
    int device = 0;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_run;
    // staging
    void *buf[12] = {nullptr};
    size_t cap[12] = {0};
    std::mutex mu;
};

namespace {

enum { B_VALUE = 0, B_LOC, B_AW, B_OUT, B_GO, B_GV, B_GL, B_GA, B_WS, B_SHAPES, B_LSI, B_COUNT };

int ctx_reserve(msda_host_ctx *c, int slot, size_t bytes)
{
    if (bytes <= c->cap[slot]) return 0;
    if (c->buf[slot]) cudaFree(c->buf[slot]);
    c->buf[slot] = nullptr; c->cap[slot] = 0;
    const cudaError_t e = cudaMalloc(&c->buf[slot], bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    c->cap[slot] = bytes;
    return 0;
}

int ctx_events(msda_host_ctx *c, size_t n)
{
    while (c->ev_in.size() < n) {
        cudaEvent_t a, b;
        cudaError_t e = cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
        if (e != cudaSuccess) return static_cast<int>(e);
        c->ev_in.push_back(a); c->ev_run.push_back(b);
    }
    return 0;
}

#define MSDA_CU(call) do { const cudaError_t e__ = (call); if (e__ != cudaSuccess) return static_cast<int>(e__); } while (0)

int host_pipeline(msda_host_ctx *c, const msda_shape *shape, const void *value, const int64_t *shapes_h,
                  const int64_t *lsi_h, const void *loc, const void *aw, const void *go, void *out,
                  void *gv, void *gl, void *ga, bool with_backward)
{
    Dims d;
    int e = check_shape(shape, d);
    if (e) return e;
    if (!c) return MSDA_E_NULL;
    if (d.N * d.Lq == 0 && !with_backward) return MSDA_OK;
    if (!value || !shapes_h || !lsi_h || !loc || !aw || !out) return MSDA_E_NULL;
    if (with_backward && (!go || !gv || !gl || !ga)) return MSDA_E_NULL;

    std::lock_guard<std::mutex> lock(c->mu);
    int prev = 0;
    MSDA_CU(cudaGetDevice(&prev));
    MSDA_CU(cudaSetDevice(c->device));
    // Every exit below restores the caller's device; an exit on an error first drains the three streams, because
    // copies and kernels of earlier chunks may still be using the caller's host buffers and the staging.
    struct Guard {
        msda_host_ctx *c; int prev; bool done = false;
        ~Guard()
        {
            if (!done) { cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->s_run); cudaStreamSynchronize(c->s_out); }
            cudaSetDevice(prev);
        }
    } guard{c, prev};

    const size_t ev = dtype_size(d.vdt), el = dtype_size(d.ldt);
    const size_t f_val = static_cast<size_t>(d.S) * d.M * d.D * ev;            // bytes per frame
    const size_t f_out = static_cast<size_t>(d.Lq) * d.M * d.D * ev;
    const size_t f_loc = static_cast<size_t>(d.Lq) * d.M * d.L * d.P * 2 * el;
    const size_t f_aw = static_cast<size_t>(d.Lq) * d.M * d.L * d.P * el;
    const size_t N = static_cast<size_t>(d.N);
    const bool need_ws = with_backward && (d.vdt == MSDA_BF16 || d.vdt == MSDA_F16);

    int rc = 0;
    if ((rc = ctx_reserve(c, B_VALUE, N * f_val)) || (rc = ctx_reserve(c, B_LOC, N * f_loc)) ||
        (rc = ctx_reserve(c, B_AW, N * f_aw)) || (rc = ctx_reserve(c, B_OUT, N * f_out)) ||
        (rc = ctx_reserve(c, B_SHAPES, sizeof(int64_t) * 2 * d.L)) || (rc = ctx_reserve(c, B_LSI, sizeof(int64_t) * d.L)))
        return rc;
    if (with_backward) {
        if ((rc = ctx_reserve(c, B_GO, N * f_out)) || (rc = ctx_reserve(c, B_GV, N * f_val)) ||
            (rc = ctx_reserve(c, B_GL, N * f_loc)) || (rc = ctx_reserve(c, B_GA, N * f_aw)) ||
            (need_ws && (rc = ctx_reserve(c, B_WS, static_cast<size_t>(d.S) * d.M * d.D * sizeof(float) * N))))
            return rc;
    }

    // chunking: up to 16 chunks (MSDA_HOST_CHUNKS overrides, for tuning), at least one frame each
    static const int64_t max_chunks = [] {
        const char *e = std::getenv("MSDA_HOST_CHUNKS");
        const long v = e ? std::atol(e) : 0;
        return static_cast<int64_t>(v > 0 ? v : 16);
    }();
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(d.N, max_chunks));
    const int64_t per = (d.N + n_chunks - 1) / n_chunks;
    if ((rc = ctx_events(c, static_cast<size_t>(n_chunks)))) return rc;

    MSDA_CU(cudaMemcpyAsync(c->buf[B_SHAPES], shapes_h, sizeof(int64_t) * 2 * d.L, cudaMemcpyHostToDevice, c->s_in));
    MSDA_CU(cudaMemcpyAsync(c->buf[B_LSI], lsi_h, sizeof(int64_t) * d.L, cudaMemcpyHostToDevice, c->s_in));

    auto at = [](void *base, size_t off) { return static_cast<void *>(static_cast<char *>(base) + off); };
    auto cat = [](const void *base, size_t off) { return static_cast<const void *>(static_cast<const char *>(base) + off); };

    int64_t ci = 0;
    for (int64_t f0 = 0; f0 < d.N; f0 += per, ++ci) {
        const size_t nf = static_cast<size_t>(std::min(per, d.N - f0));
        const size_t o = static_cast<size_t>(f0);
        MSDA_CU(cudaMemcpyAsync(at(c->buf[B_VALUE], o * f_val), cat(value, o * f_val), nf * f_val, cudaMemcpyHostToDevice, c->s_in));
        MSDA_CU(cudaMemcpyAsync(at(c->buf[B_LOC], o * f_loc), cat(loc, o * f_loc), nf * f_loc, cudaMemcpyHostToDevice, c->s_in));
        MSDA_CU(cudaMemcpyAsync(at(c->buf[B_AW], o * f_aw), cat(aw, o * f_aw), nf * f_aw, cudaMemcpyHostToDevice, c->s_in));
        if (with_backward)
            MSDA_CU(cudaMemcpyAsync(at(c->buf[B_GO], o * f_out), cat(go, o * f_out), nf * f_out, cudaMemcpyHostToDevice, c->s_in));
        MSDA_CU(cudaEventRecord(c->ev_in[ci], c->s_in));

        MSDA_CU(cudaStreamWaitEvent(c->s_run, c->ev_in[ci], 0));
        msda_shape sub = *shape;
        sub.n_frames = static_cast<int64_t>(nf);
        e = msda_forward(&sub, at(c->buf[B_VALUE], o * f_val), static_cast<const int64_t *>(c->buf[B_SHAPES]),
                         static_cast<const int64_t *>(c->buf[B_LSI]), at(c->buf[B_LOC], o * f_loc),
                         at(c->buf[B_AW], o * f_aw), at(c->buf[B_OUT], o * f_out), c->s_run);
        if (e) return e;
        if (with_backward) {
            const size_t f_ws = static_cast<size_t>(d.S) * d.M * d.D * sizeof(float);
            e = msda_backward(&sub, at(c->buf[B_VALUE], o * f_val), static_cast<const int64_t *>(c->buf[B_SHAPES]),
                              static_cast<const int64_t *>(c->buf[B_LSI]), at(c->buf[B_LOC], o * f_loc),
                              at(c->buf[B_AW], o * f_aw), at(c->buf[B_GO], o * f_out), at(c->buf[B_GV], o * f_val),
                              at(c->buf[B_GL], o * f_loc), at(c->buf[B_GA], o * f_aw),
                              need_ws ? at(c->buf[B_WS], o * f_ws) : nullptr, need_ws ? nf * f_ws : 0, c->s_run);
            if (e) return e;
        }
        MSDA_CU(cudaEventRecord(c->ev_run[ci], c->s_run));

        MSDA_CU(cudaStreamWaitEvent(c->s_out, c->ev_run[ci], 0));
        MSDA_CU(cudaMemcpyAsync(at(out, o * f_out), at(c->buf[B_OUT], o * f_out), nf * f_out, cudaMemcpyDeviceToHost, c->s_out));
        if (with_backward) {
            MSDA_CU(cudaMemcpyAsync(at(gv, o * f_val), at(c->buf[B_GV], o * f_val), nf * f_val, cudaMemcpyDeviceToHost, c->s_out));
            MSDA_CU(cudaMemcpyAsync(at(gl, o * f_loc), at(c->buf[B_GL], o * f_loc), nf * f_loc, cudaMemcpyDeviceToHost, c->s_out));
            MSDA_CU(cudaMemcpyAsync(at(ga, o * f_aw), at(c->buf[B_GA], o * f_aw), nf * f_aw, cudaMemcpyDeviceToHost, c->s_out));
        }
    }
    MSDA_CU(cudaStreamSynchronize(c->s_out));
    MSDA_CU(cudaStreamSynchronize(c->s_run));
    MSDA_CU(cudaStreamSynchronize(c->s_in));
    guard.done = true;
    return MSDA_OK;
}

}  // namespace

extern "C" {

int msda_host_ctx_create(msda_host_ctx **ctx, int device)
{
    if (!ctx) return MSDA_E_NULL;
    *ctx = nullptr;
    int prev = 0;
    MSDA_CU(cudaGetDevice(&prev));
    MSDA_CU(cudaSetDevice(device));
    msda_host_ctx *c = new msda_host_ctx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { msda_host_ctx_destroy(c); return static_cast<int>(e); }
    *ctx = c;
    return MSDA_OK;
}

void msda_host_ctx_destroy(msda_host_ctx *c)
{
    if (!c) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    for (int i = 0; i < B_COUNT; ++i) if (c->buf[i]) cudaFree(c->buf[i]);
    for (cudaEvent_t ev : c->ev_in) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->ev_run) cudaEventDestroy(ev);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_run) cudaStreamDestroy(c->s_run);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    cudaSetDevice(prev);
    delete c;
}

int msda_forward_host(msda_host_ctx *ctx, const msda_shape *shape, const void *value,
                      const int64_t *shapes, const int64_t *lsi, const void *loc, const void *aw, void *out)
{
    return host_pipeline(ctx, shape, value, shapes, lsi, loc, aw, nullptr, out, nullptr, nullptr, nullptr, false);
}

int msda_forward_backward_host(msda_host_ctx *ctx, const msda_shape *shape, const void *value,
                               const int64_t *shapes, const int64_t *lsi, const void *loc, const void *aw,
                               const void *go, void *out, void *gv, void *gl, void *ga)
{
    return host_pipeline(ctx, shape, value, shapes, lsi, loc, aw, go, out, gv, gl, ga, true);
}

}  // extern "C"
