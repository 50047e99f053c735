/*
 * msda_b200.h — C ABI of the B200-native multi-scale deformable attention library
 * (libmsda_b200.so).
 *
 * This is the drop-in boundary for TCE-RVOS's hot path.  Every entry point
 * replaces one piece of the reference's native extension; citations are
 * file:line under the reference tree (haliphinx/TCE-RVOS):
 *
 *   msda_forward            <- ms_deform_attn_cuda_forward  (models/ops/src/cuda/ms_deform_attn_cuda.cu:21-102)
 *                              + ms_deformable_im2col_cuda  (models/ops/src/cuda/ms_deform_im2col_cuda.cuh:1232-1269)
 *   msda_backward           <- ms_deform_attn_cuda_backward (ms_deform_attn_cuda.cu:105-186)
 *                              + ms_deformable_col2im_cuda  (ms_deform_im2col_cuda.cuh:1271-1674)
 *   msda_forward_3d /       <- the is_3d=true branches of the same two functions
 *   msda_backward_3d           (ms_deform_im2col_cuda.cuh:366-420, :565-712; MSDeformAttnFunction3D,
 *                              models/ops/functions/ms_deform_attn_func.py:43-64)
 *   msda_prologue_forward   <- the elementwise prologue of MSDeformAttn.forward
 *                              (models/ops/modules/ms_deform_attn.py:99-109): softmax over L*P
 *                              and reference_points + offsets/normalizer, fused into the gather
 *   msda_*_host             <- same operators for callers that hold HOST buffers; the library
 *                              owns the device staging and the copy/compute overlap
 *
 * No torch type crosses this line: plain pointers, sizes, a cudaStream_t passed as
 * void*.  All device pointers must be valid on the current CUDA device.  Tensors are
 * dense and row-major with the reference's layouts:
 *
 *   value            [N, S, M, D]            value dtype
 *   spatial_shapes   [L, 2]  int64 (H_l,W_l) DEVICE memory (the reference passes CUDA int64 tensors)
 *   level_start_index[L]     int64           DEVICE memory
 *   sampling_loc     [N, Lq, M, L, P, 2]     loc dtype, (x, y) normalised to [0,1], may lie outside
 *                    [N, Lq, M, L, P, 3]     for the *_3d entry points: (x, y, frame)
 *   attn_weight      [N, Lq, M, L, P]        loc dtype
 *   out / grad_out   [N, Lq, M*D]            value dtype
 *
 * Return value: 0 on success; a positive value is a cudaError_t; a negative value is
 * one of the MSDA_E_* codes below.  msda_error_string() turns either into text.  Unlike
 * the reference, which only printf()s launch failures (ms_deform_im2col_cuda.cuh:1263-1267),
 * every launch error is returned to the caller.  No entry point synchronises the device
 * (the reference's forward calls cudaDeviceSynchronize(), .cuh:1260), so all device-pointer
 * entry points are CUDA-graph capturable.
 */
#ifndef MSDA_B200_H
#define MSDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 1

/* element types */
enum {
    MSDA_F32  = 0,
    MSDA_BF16 = 1,
    MSDA_F16  = 2,
    MSDA_F64  = 3
};

/* library error codes (negative; positive codes are cudaError_t) */
enum {
    MSDA_OK            = 0,
    MSDA_E_NULL        = -1,  /* a required pointer is NULL                       */
    MSDA_E_SHAPE       = -2,  /* a dimension is <= 0 or overflows the index model */
    MSDA_E_DTYPE       = -3,  /* unsupported dtype combination                    */
    MSDA_E_ALIGN       = -4,  /* pointer not aligned for the element type         */
    MSDA_E_WORKSPACE   = -5,  /* workspace too small                              */
    MSDA_E_UNSUPPORTED = -6,  /* feature not built                                */
    MSDA_E_IM2COL_STEP = -7   /* N % min(N, im2col_step) != 0 (ms_deform_attn_cuda.cu:54-56) */
};

/* Problem geometry.  Field order is ABI; append only. */
typedef struct msda_shape {
    int64_t n_frames;      /* N  : value.size(0) = batch x T frames            */
    int64_t spatial_size;  /* S  : sum_l H_l*W_l                               */
    int64_t n_queries;     /* Lq : queries per frame                           */
    int32_t n_heads;       /* M                                                */
    int32_t channels;      /* D  : channels per head                           */
    int32_t n_levels;      /* L                                                */
    int32_t n_points;      /* P                                                */
    int32_t value_dtype;   /* MSDA_* of value / out / grad_out / grad_value    */
    int32_t loc_dtype;     /* MSDA_* of sampling_loc / attn_weight + grads     */
} msda_shape;

/* ---- introspection ------------------------------------------------------------ */

int         msda_abi_version(void);
const char *msda_error_string(int code);
/* Name of the kernel family the dispatcher picks for this shape ("fast_d32_f32",
 * "fast_d32_bf16", "generic", ...); for logs, tests and profiles. */
const char *msda_kernel_name(const msda_shape *shape, int backward);
/* Algorithmic bytes (each tensor moved once) of one call: pass 0 = forward,
 * 1 = backward, 2 = forward+backward.  SURVEY.md section 8(d). */
int64_t     msda_algorithmic_bytes(const msda_shape *shape, int pass);
/* The reference's im2col_step precondition (ms_deform_attn_cuda.cu:54-56): returns 0
 * or MSDA_E_IM2COL_STEP.  The kernels themselves never chunk. */
int         msda_check_im2col_step(int64_t n_frames, int im2col_step);
/* Number of kernel launches issued by this library since load (for bench.py's
 * gpu_launches claim). */
int64_t     msda_launch_count(void);

/* ---- device-pointer operators --------------------------------------------------- */

/* out[n,q,m,:] = sum_{l,p} attn_weight * bilinear(value_l[n,:,m,:], loc)   (a3-a5) */
int msda_forward(const msda_shape *shape,
                 const void *value,
                 const int64_t *spatial_shapes,
                 const int64_t *level_start_index,
                 const void *sampling_loc,
                 const void *attn_weight,
                 void *out,
                 void *stream);

/* Bytes of scratch msda_backward wants for this shape (0 when none is needed: fp32/fp64
 * accumulate in place; bf16/f16 grad_value is accumulated in fp32 scratch and rounded once). */
size_t msda_backward_workspace_bytes(const msda_shape *shape);

/* grad_value (scatter-add), grad_sampling_loc, grad_attn_weight              (a6-a8)
 * grad_value is zero-filled by the callee; grad_sampling_loc / grad_attn_weight are
 * written exactly once per element (no memset pass).  workspace may be NULL when
 * msda_backward_workspace_bytes() == 0; when it is NULL for a 16-bit value dtype the
 * library falls back to 16-bit vector atomics directly into grad_value (faster, one
 * rounding per contribution instead of one per element). */
int msda_backward(const msda_shape *shape,
                  const void *value,
                  const int64_t *spatial_shapes,
                  const int64_t *level_start_index,
                  const void *sampling_loc,
                  const void *attn_weight,
                  const void *grad_out,
                  void *grad_value,
                  void *grad_sampling_loc,
                  void *grad_attn_weight,
                  void *workspace, size_t workspace_bytes,
                  void *stream);

/* The fork's cross-frame variant: a third location coordinate selects a fractional frame
 * index f*G-0.5 inside a group of G = min(N, frames_per_group) consecutive frames and the two
 * neighbouring frames are blended linearly.  frames_per_group is the reference's im2col_step:
 * its chunk loop (ms_deform_attn_cuda.cu:54,77) makes the blend chunk-relative. */
int msda_forward_3d(const msda_shape *shape, int64_t frames_per_group,
                    const void *value,
                    const int64_t *spatial_shapes,
                    const int64_t *level_start_index,
                    const void *sampling_loc,   /* [..., 3] */
                    const void *attn_weight,
                    void *out,
                    void *stream);

int msda_backward_3d(const msda_shape *shape, int64_t frames_per_group,
                     const void *value,
                     const int64_t *spatial_shapes,
                     const int64_t *level_start_index,
                     const void *sampling_loc,  /* [..., 3] */
                     const void *attn_weight,
                     const void *grad_out,
                     void *grad_value,
                     void *grad_sampling_loc,   /* [..., 3] */
                     void *grad_attn_weight,
                     void *workspace, size_t workspace_bytes,
                     void *stream);

/* Fused prologue + gather (row f1).  Consumes the raw outputs of the module's two
 * Linear layers instead of materialised locations / weights:
 *   offsets    [N, Lq, M, L, P, 2]  loc dtype   (sampling_offsets(query))
 *   logits     [N, Lq, M, L*P]      loc dtype   (attention_weights(query), pre-softmax)
 *   ref_points [N, Lq, L, ref_dim]  loc dtype, ref_dim = 2 or 4
 *                                   (ms_deform_attn.py:103-109)
 * loc_out / aw_out may be NULL (encoder / FTF discard them) or receive the materialised
 * sampling_locations / attention_weights (the decoder consumes them,
 * tce_deformable_transformer.py:752-758). */
/* 1 if the fused prologue kernels cover this shape (D == 32, L*P == 16, fp32 or bf16), else 0;
 * callers then compose the prologue themselves and use msda_forward / msda_backward. */
int msda_prologue_supported(const msda_shape *shape, int ref_dim);

int msda_prologue_forward(const msda_shape *shape,
                          const void *value,
                          const int64_t *spatial_shapes,
                          const int64_t *level_start_index,
                          const void *offsets,
                          const void *logits,
                          const void *ref_points, int ref_dim,
                          void *out,
                          void *loc_out,
                          void *aw_out,
                          void *stream);

/* Backward of msda_prologue_forward: grad wrt value, raw offsets, raw logits and
 * (optionally, may be NULL) reference points [N, Lq, L, ref_dim]. */
int msda_prologue_backward(const msda_shape *shape,
                           const void *value,
                           const int64_t *spatial_shapes,
                           const int64_t *level_start_index,
                           const void *offsets,
                           const void *logits,
                           const void *ref_points, int ref_dim,
                           const void *grad_out,
                           void *grad_value,
                           void *grad_offsets,
                           void *grad_logits,
                           float *grad_ref_points,   /* fp32 whatever the loc dtype; may be NULL */
                           void *workspace, size_t workspace_bytes,
                           void *stream);

/* ---- the two small passes either side of the sampling op ---- */

/* Value epilogue (f2; models/ops/modules/ms_deform_attn.py:95-98): dst[r, :] = 0 where
 * padding_mask[r] != 0, else convert(src[r, :]).  src / dst are [rows, width] (rows = N*S,
 * width = M*D, a multiple of 8), dtypes MSDA_F32 or MSDA_BF16 in any combination (fp32 -> bf16
 * feeds the 16-bit gather; bf16 -> fp32 with the same mask is its backward).  padding_mask is
 * [rows] bool bytes or NULL.  dst may alias src when the dtypes agree.  Masked rows of src are
 * not read. */
int msda_value_epilogue(int64_t rows, int width,
                        const void *src, int src_dtype,
                        const uint8_t *padding_mask,
                        void *dst, int dst_dtype,
                        void *stream);

/* Decoder sample selection (f3; models/tce_deformable_transformer.py:752-758): per (frame,
 * query) the k largest of the M*L*P attention weights, in descending order, ties broken by
 * the lower flat index (m, l, p) — torch.topk leaves that order unspecified — and the
 * sampling locations of those points divided by valid_ratios[n, l, :] (NULL = no division).
 *   sampling_loc [N, Lq, M, L, P, 2], attn_weight [N, Lq, M, L, P], valid_ratios [N, L, 2],
 *   all in the shape's loc dtype (fp32 or bf16);  M*L*P <= 512, 1 <= k <= M*L*P
 *   samples_keep [N, Lq, k, 2], weights_keep [N, Lq, k], index_keep [N, Lq, k] int64 or NULL.
 * The value dtype / spatial size / channels of the shape are ignored. */
int msda_topk_samples(const msda_shape *shape, int k,
                      const void *sampling_loc,
                      const void *attn_weight,
                      const void *valid_ratios,
                      void *samples_keep,
                      void *weights_keep,
                      int64_t *index_keep,
                      void *stream);

/* ---- host-buffer operators ---------------------------------------------------------
 * Same operators for a caller whose tensors live in HOST memory (pinned memory gives
 * full PCIe rate; pageable works).  spatial_shapes / level_start_index are HOST
 * pointers here.  Frames are independent, so the library splits
 * N into chunks and pipelines H2D copy / kernels / D2H copy on three streams.  The calls
 * return after all results are in the host buffers.  Device staging is cached in an
 * opaque context so that repeated calls do not allocate. */
typedef struct msda_host_ctx msda_host_ctx;

int  msda_host_ctx_create(msda_host_ctx **ctx, int device);
void msda_host_ctx_destroy(msda_host_ctx *ctx);

int msda_forward_host(msda_host_ctx *ctx, const msda_shape *shape,
                      const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index,
                      const void *sampling_loc, const void *attn_weight,
                      void *out);

int msda_forward_backward_host(msda_host_ctx *ctx, const msda_shape *shape,
                               const void *value, const int64_t *spatial_shapes,
                               const int64_t *level_start_index,
                               const void *sampling_loc, const void *attn_weight,
                               const void *grad_out,
                               void *out,
                               void *grad_value, void *grad_sampling_loc,
                               void *grad_attn_weight);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H */
