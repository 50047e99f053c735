// msda_aux_host.inl — C-ABI entry points of the value epilogue (f2) and the decoder's top-K sample
// selection (f3).  Included at the end of msda_capi.cu.

#include "msda_aux_kernels.cuh"

namespace {

template <typename ST, typename DT>
int launch_value_epilogue(int64_t rows, int width, const void *src, const uint8_t *mask, void *dst, cudaStream_t st)
{
    const int64_t n_vec = rows * (width / 8);
    if (n_vec == 0) return MSDA_OK;
    const int vpr = width / 8;
    const int shift = ilog2_exact(vpr);
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n_vec + 255) / 256, int64_t(sm_count_aux()) * 16));
    msda::value_epilogue_kernel<ST, DT><<<grid, 256, 0, st>>>(static_cast<const ST *>(src), mask, static_cast<DT *>(dst),
                                                              n_vec, vpr, shift);
    ++g_launches;
    return launch_status();
}

template <typename LT>
int launch_topk(const Dims &d, int k, const void *loc, const void *aw, const void *vr, void *samples, void *weights,
                int64_t *index, cudaStream_t st)
{
    const int64_t nq = d.N * d.Lq;
    if (nq == 0) return MSDA_OK;
    const int C = d.M * d.L * d.P;
    const unsigned grid = static_cast<unsigned>((nq + 3) / 4);           // 4 warps per CTA
#define MSDA_TOPK(E)                                                                                              \
    msda::topk_samples_kernel<LT, E><<<grid, 128, 0, st>>>(static_cast<const LT *>(loc), static_cast<const LT *>(aw), \
        static_cast<const LT *>(vr), static_cast<LT *>(samples), static_cast<LT *>(weights), index, nq, d.Lq, C, d.L, d.P, k)
    if (C <= 32) MSDA_TOPK(1);
    else if (C <= 64) MSDA_TOPK(2);
    else if (C <= 128) MSDA_TOPK(4);
    else if (C <= 256) MSDA_TOPK(8);
    else MSDA_TOPK(16);
#undef MSDA_TOPK
    ++g_launches;
    return launch_status();
}

}  // namespace

extern "C" {

int msda_value_epilogue(int64_t rows, int width, const void *src, int src_dtype, const uint8_t *padding_mask,
                        void *dst, int dst_dtype, void *stream)
{
    if (rows < 0 || width <= 0) return MSDA_E_SHAPE;
    if (rows == 0) return MSDA_OK;
    if (!src || !dst) return MSDA_E_NULL;
    if (width % 8 != 0) return MSDA_E_UNSUPPORTED;
    if (!aligned16(src) || !aligned16(dst)) return MSDA_E_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (src_dtype == MSDA_F32 && dst_dtype == MSDA_F32) return launch_value_epilogue<float, float>(rows, width, src, padding_mask, dst, st);
    if (src_dtype == MSDA_F32 && dst_dtype == MSDA_BF16) return launch_value_epilogue<float, __nv_bfloat16>(rows, width, src, padding_mask, dst, st);
    if (src_dtype == MSDA_BF16 && dst_dtype == MSDA_F32) return launch_value_epilogue<__nv_bfloat16, float>(rows, width, src, padding_mask, dst, st);
    if (src_dtype == MSDA_BF16 && dst_dtype == MSDA_BF16) return launch_value_epilogue<__nv_bfloat16, __nv_bfloat16>(rows, width, src, padding_mask, dst, st);
    return MSDA_E_DTYPE;
}

int msda_topk_samples(const msda_shape *shape, int k, const void *sampling_loc, const void *attn_weight,
                      const void *valid_ratios, void *samples_keep, void *weights_keep, int64_t *index_keep,
                      void *stream)
{
    if (!shape) return MSDA_E_NULL;
    // only the query-side dimensions and the loc dtype of the shape matter here
    Dims d{shape->n_frames, 1, shape->n_queries, shape->n_heads, 1, shape->n_levels, shape->n_points, MSDA_F32, shape->loc_dtype};
    if (d.N < 0 || d.Lq < 0 || d.M <= 0 || d.L <= 0 || d.P <= 0) return MSDA_E_SHAPE;
    const int64_t C = int64_t(d.M) * d.L * d.P;
    if (k <= 0 || k > C || C > 512) return MSDA_E_UNSUPPORTED;
    if (d.N * d.Lq == 0) return MSDA_OK;
    if (!sampling_loc || !attn_weight || !samples_keep || !weights_keep) return MSDA_E_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d.ldt == MSDA_F32) return launch_topk<float>(d, k, sampling_loc, attn_weight, valid_ratios, samples_keep, weights_keep, index_keep, st);
    if (d.ldt == MSDA_BF16) return launch_topk<__nv_bfloat16>(d, k, sampling_loc, attn_weight, valid_ratios, samples_keep, weights_keep, index_keep, st);
    return MSDA_E_DTYPE;
}

}  // extern "C"
