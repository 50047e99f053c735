// msda_aux_kernels.cuh — the two small kernels either side of the sampling op (SURVEY.md 8f rows
// f2 and f3).  Written from scratch for sm_100a.
//
//  * value_epilogue_kernel  (f2)  models/ops/modules/ms_deform_attn.py:95-98: rows of the value_proj
//    output whose padding-mask byte is set become zero; optionally the row is converted
//    (fp32 -> bf16 for the 16-bit gather, or bf16 -> fp32 on the way back = its own backward).
//    One pass: 32 bytes in, 32 or 16 bytes out per thread; HBM-bound.
//  * topk_samples_kernel    (f3)  models/tce_deformable_transformer.py:752-758: per (frame, query)
//    the K largest of the M*L*P attention weights, and the sampling locations of those points
//    divided by the level's valid ratio.  The reference runs div + topk + 2 gathers + repeat
//    (5 launches over the full [N, Lq, 128, 2] tensor); here one warp sorts the 128 candidates in
//    registers (bitonic network over packed 64-bit keys) and writes only the K survivors.
//    Ties are broken by the lower candidate index (torch.topk leaves the order of equal values
//    unspecified).
#pragma once

#include "msda_kernels.cuh"

namespace msda {

// ---------------------------------------------------------------------------------------------
// f2: masked fill + conversion, 8 elements per thread
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec8;
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[8])
    {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(p));
        const float4 b = __ldcs(reinterpret_cast<const float4 *>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[8])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *(reinterpret_cast<float4 *>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8])
    {
        const uint4 t = __ldcs(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8])
    {
        uint4 t;
        t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
        t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4 *>(p) = t;
    }
};

// vec_shift >= 0: vectors per row is 1 << vec_shift (TCE: 256 channels -> 32 vectors, shift 5)
template <typename ST, typename DT>
__global__ void __launch_bounds__(256)
value_epilogue_kernel(const ST *__restrict__ src, const uint8_t *__restrict__ mask, DT *__restrict__ dst,
                      int64_t n_vec, int vec_per_row, int vec_shift)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        bool masked = false;
        if (mask) {
            const int64_t row = vec_shift >= 0 ? (i >> vec_shift) : (i / vec_per_row);
            masked = __ldg(mask + row) != 0;
        }
        float v[8];
        if (masked) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;       // masked rows are not even read
        } else {
            Vec8<ST>::load(src + i * 8, v);
        }
        Vec8<DT>::store(dst + i * 8, v);
    }
}

// ---------------------------------------------------------------------------------------------
// f3: top-K of the attention weights of one query + the sampling locations of the survivors
// ---------------------------------------------------------------------------------------------
// order-preserving map float -> uint32 (larger float <=> larger unsigned; -0 < +0; NaNs sort above +inf,
// i.e. they win like they do in torch.topk)
__device__ __forceinline__ uint32_t float_order_key(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int ELEMS>
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long (&key)[ELEMS], int lane)
{
    constexpr int N = ELEMS * 32;
    // element index e = lane * ELEMS + r.  Standard bitonic network, final order descending.
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= ELEMS) {
                const int lj = j / ELEMS;               // partner lane distance
#pragma unroll
                for (int r = 0; r < ELEMS; ++r) {
                    const int e = lane * ELEMS + r;
                    const unsigned long long other = __shfl_xor_sync(kFull, key[r], lj);
                    const bool lower = (e & j) == 0;     // this element is the lower index of the pair
                    const bool desc = (e & k) == 0;      // direction of this k-block (descending overall)
                    const bool take_max = (lower == desc);
                    const unsigned long long mx = key[r] > other ? key[r] : other;
                    const unsigned long long mn = key[r] > other ? other : key[r];
                    key[r] = take_max ? mx : mn;
                }
            } else {
#pragma unroll
                for (int r = 0; r < ELEMS; ++r) {
                    if ((r & j) == 0) {
                        const int e = lane * ELEMS + r;
                        const bool desc = (e & k) == 0;
                        const unsigned long long a = key[r], b = key[r | j];
                        const bool swap = desc ? (a < b) : (a > b);
                        key[r] = swap ? b : a;
                        key[r | j] = swap ? a : b;
                    }
                }
            }
        }
    }
}

// One warp per (frame, query).  C = M*L*P candidates, padded to ELEMS*32 with -inf-like keys.
template <typename LT, int ELEMS>
__global__ void __launch_bounds__(128)
topk_samples_kernel(const LT *__restrict__ loc, const LT *__restrict__ aw, const LT *__restrict__ valid_ratios,
                    LT *__restrict__ samples_keep, LT *__restrict__ weights_keep, int64_t *__restrict__ index_keep,
                    int64_t n_queries, int64_t Lq, int C, int L, int P, int K)
{
    const int lane = threadIdx.x & 31;
    const int64_t nq = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (nq >= n_queries) return;
    const LT *w = aw + nq * C;
    unsigned long long key[ELEMS];
#pragma unroll
    for (int r = 0; r < ELEMS; ++r) {
        const int c = lane * ELEMS + r;
        // padding entries carry key 0: below every real candidate (real keys have a non-zero low word)
        key[r] = 0ull;
        if (c < C) {
            const float f = cvt_to<float>(w[c]);
            key[r] = (static_cast<unsigned long long>(float_order_key(f)) << 32) | (0xffffffffu - static_cast<uint32_t>(c));
        }
    }
    bitonic_sort_desc<ELEMS>(key, lane);
    const int64_t n = nq / Lq;
#pragma unroll
    for (int r = 0; r < ELEMS; ++r) {
        const int rank = lane * ELEMS + r;
        if (rank < K) {
            const int c = static_cast<int>(0xffffffffu - static_cast<uint32_t>(key[r]));
            const int l = (c / P) % L;
            const float x = cvt_to<float>(loc[(nq * C + c) * 2]), y = cvt_to<float>(loc[(nq * C + c) * 2 + 1]);
            float rx = 1.f, ry = 1.f;
            if (valid_ratios) {
                rx = cvt_to<float>(valid_ratios[(n * L + l) * 2]);
                ry = cvt_to<float>(valid_ratios[(n * L + l) * 2 + 1]);
            }
            samples_keep[(nq * K + rank) * 2] = cvt_from<LT, float>(x / rx);
            samples_keep[(nq * K + rank) * 2 + 1] = cvt_from<LT, float>(y / ry);
            weights_keep[nq * K + rank] = w[c];
            if (index_keep) index_keep[nq * K + rank] = c;
        }
    }
}

}  // namespace msda
