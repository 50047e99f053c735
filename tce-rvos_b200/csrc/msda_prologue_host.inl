// msda_prologue_host.inl — host dispatch of the fused prologue operators (row f1).
// Included at the end of msda_capi.cu (shares its anonymous-namespace helpers).

extern "C" {

int msda_prologue_supported(const msda_shape *shape, int ref_dim)
{
    Dims d;
    if (check_shape(shape, d)) return 0;
    return (fast_eligible(d) && (ref_dim == 2 || ref_dim == 4)) ? 1 : 0;
}

int msda_prologue_forward(const msda_shape *shape, const void *value, const int64_t *shapes,
                          const int64_t *lsi, const void *offsets, const void *logits,
                          const void *ref_points, int ref_dim, void *out, void *loc_out, void *aw_out,
                          void *stream)
{
    Dims d;
    const int e = check_shape(shape, d);
    if (e) return e;
    if (ref_dim != 2 && ref_dim != 4) return MSDA_E_SHAPE;
    if (!fast_eligible(d)) return MSDA_E_UNSUPPORTED;
    if (d.N * d.Lq == 0) return MSDA_OK;
    if (!value || !shapes || !lsi || !offsets || !logits || !ref_points || !out) return MSDA_E_NULL;
    if (!aligned16(value) || !aligned16(offsets) || !aligned16(logits) || !aligned16(out) ||
        (loc_out && !aligned16(loc_out)) || (aw_out && !aligned16(aw_out)))
        return MSDA_E_ALIGN;
    msda::ProArgs pro{};
    pro.ref = ref_points; pro.ref_dim = ref_dim; pro.n_levels = d.L; pro.n_points = static_cast<float>(d.P);
    pro.loc_out = loc_out; pro.aw_out = aw_out; pro.grad_ref = nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d.vdt == MSDA_F32) return launch_fwd_fast<float, float, true>(d, value, shapes, lsi, offsets, logits, out, st, pro);
    if (d.ldt == MSDA_F32) return launch_fwd_fast<__nv_bfloat16, float, true>(d, value, shapes, lsi, offsets, logits, out, st, pro);
    return launch_fwd_fast<__nv_bfloat16, __nv_bfloat16, true>(d, value, shapes, lsi, offsets, logits, out, st, pro);
}

int msda_prologue_backward(const msda_shape *shape, const void *value, const int64_t *shapes,
                           const int64_t *lsi, const void *offsets, const void *logits,
                           const void *ref_points, int ref_dim, const void *go, void *gv,
                           void *g_off, void *g_logit, float *g_ref, void *ws, size_t ws_bytes, void *stream)
{
    Dims d;
    int e = check_shape(shape, d);
    if (e) return e;
    if (ref_dim != 2 && ref_dim != 4) return MSDA_E_SHAPE;
    if (!fast_eligible(d)) return MSDA_E_UNSUPPORTED;
    if (!value || !shapes || !lsi || !offsets || !logits || !ref_points || !go || !gv || !g_off || !g_logit)
        return MSDA_E_NULL;
    if (!aligned16(value) || !aligned16(offsets) || !aligned16(logits) || !aligned16(go) || !aligned16(gv) ||
        !aligned16(g_off) || !aligned16(g_logit) || (ws && !aligned16(ws)))
        return MSDA_E_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nval = d.N * d.S * d.M * d.D;
    const bool use_ws = (d.vdt == MSDA_BF16) && ws != nullptr;
    if (use_ws && ws_bytes < static_cast<size_t>(nval) * sizeof(float)) return MSDA_E_WORKSPACE;
    cudaError_t ce = use_ws ? cudaMemsetAsync(ws, 0, static_cast<size_t>(nval) * sizeof(float), st)
                            : cudaMemsetAsync(gv, 0, static_cast<size_t>(nval) * dtype_size(d.vdt), st);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    if (g_ref) {
        ce = cudaMemsetAsync(g_ref, 0, static_cast<size_t>(d.N * d.Lq * d.L * ref_dim) * sizeof(float), st);
        if (ce != cudaSuccess) return static_cast<int>(ce);
    }
    if (d.N * d.Lq == 0) return MSDA_OK;
    msda::ProArgs pro{};
    pro.ref = ref_points; pro.ref_dim = ref_dim; pro.n_levels = d.L; pro.n_points = static_cast<float>(d.P);
    pro.loc_out = nullptr; pro.aw_out = nullptr; pro.grad_ref = g_ref;
    if (d.vdt == MSDA_F32)
        return launch_bwd_fast<float, float, float, true>(d, value, shapes, lsi, offsets, logits, go,
                                                          static_cast<float *>(gv), g_off, g_logit, st, pro);
    if (use_ws) {
        e = (d.ldt == MSDA_F32)
                ? launch_bwd_fast<__nv_bfloat16, float, float, true>(d, value, shapes, lsi, offsets, logits, go, static_cast<float *>(ws), g_off, g_logit, st, pro)
                : launch_bwd_fast<__nv_bfloat16, __nv_bfloat16, float, true>(d, value, shapes, lsi, offsets, logits, go, static_cast<float *>(ws), g_off, g_logit, st, pro);
        if (e) return e;
        return launch_cast<__nv_bfloat16>(static_cast<const float *>(ws), gv, nval, st);
    }
    return (d.ldt == MSDA_F32)
               ? launch_bwd_fast<__nv_bfloat16, float, __nv_bfloat16, true>(d, value, shapes, lsi, offsets, logits, go, static_cast<__nv_bfloat16 *>(gv), g_off, g_logit, st, pro)
               : launch_bwd_fast<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16, true>(d, value, shapes, lsi, offsets, logits, go, static_cast<__nv_bfloat16 *>(gv), g_off, g_logit, st, pro);
}

}  // extern "C"
