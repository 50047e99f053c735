// msda_tile_kernels.cuh — persistent, SM-resident kernels for the encoder-sized calls of the hot path
// (sm_100a, TCE geometry: 8 heads x 32 channels, 4 levels x 4 points, fp32 or bf16 values).
//
// Replaces (semantics only) ms_deformable_im2col_gpu_kernel (models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:320-455) and ms_deformable_col2im_gpu_kernel_*  (:457-563) for calls with
// enough queries to fill the machine; the one-shot kernels of msda_kernels.cuh keep the small calls.
//
// Why a second kernel family (round-1 measurements): the one-shot backward sends
// every corner contribution to the L2 as a 128-byte vector reduction and an SM can push only
// ~32 B/clk into the crossbar, so a row update costs 4-6.3 SM cycles; half of those updates go to
// the two coarse levels, which are only 300 pixels per frame.  Here a CTA stays on its SM, works on
// ONE (frame, group of 4 heads) at a time and keeps the grad_value rows of the coarse levels of that
// unit in shared memory (512 B per pixel = 4 heads x 32 fp32 channels; every suffix of levels that
// fits is taken, decided on the device from spatial_shapes).  A row update becomes LDS.128 + FADD +
// STS.128 (~2 cycles of the shared-memory pipe) and leaves the SM once per unit as a bulk
// reduction (cp.reduce.async.bulk .add.f32, 512 B per pixel).
//
// Race freedom without shared-memory float atomics (sm_100a has none; a CAS.128 loop costs 5.6
// cycles per row): the rows of one level form an ownership domain guarded by a warp-level lock.
// A warp does all the global work of its query (16 points: corner loads, dot products, the fine
// levels' vector reductions, grad_sampling_loc / grad_attn_weight) without any lock and then enters
// the critical section of each coarse level for only the 16 row updates of that level.  Inside the
// section a lane always touches the same 16 bytes of a row that it alone reads and writes, the four
// lane groups of a warp are four different heads, and the four corners of a point are four
// different pixels, so plain loads and stores are exact.
//
// Scheduling: the queries of a unit are cut into contiguous segments, one per CTA where possible
// (spatial locality of the corner loads in the SM's L1), each with a grab counter in a small
// workspace; warps take batches of consecutive queries from their CTA's segment; a CTA that runs
// dry first claims an untouched segment near its own position, then helps the segment with the most
// work left.  The accumulator is flushed only when the CTA moves to another unit.
//
// The forward kernel uses the same schedule and stages the coarse levels' VALUE rows of the unit
// in shared memory with bulk copies (cp.async.bulk + mbarrier), so that half of the corner reads
// never enter the L1 tag pipeline.  Both kernels hand a point's weights from the owning lane to the
// other lanes of the head through a 1 KB per-warp table in shared memory (one broadcast LDS.128 per
// point) instead of four shuffles.
#pragma once

#include "msda_kernels.cuh"

namespace msda {

struct TileParams {
    int64_t Lq, S;
    int n_units;         // frames x head groups (unit = n * 2 + hg)
    int segs_per_unit;
    int n_segs;
    int batch;           // consecutive queries per grab
    int cap_rows;        // coarse pixels (512-byte rows) the shared map can hold
    int *ctr;            // workspace, zero-filled per launch: [n_segs] grab counters, [n_segs] claimed flags
};

constexpr int kTileRowBytes = 512;       // one pixel of one unit: 4 heads x 32 ch fp32 (or 8 heads x 32 ch bf16)

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_reduce_add_f32(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n\t"
                 "DONE_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- per-CTA control block -----------------------------------------------------------------
#ifndef MSDA_TILE_SLEEP_NS
#define MSDA_TILE_SLEEP_NS 100u
#endif
struct TileLock { unsigned next, serving; };
struct TileCtl {
    unsigned long long best;
    uint64_t bar;          // mbarrier of the forward's value staging
    int seg;
    TileLock lock[4];      // one per level (only the coarse ones are used)
};

__device__ __forceinline__ void seg_range(const TileParams &tp, int s, int &unit, int &q0, int &q1)
{
    unit = s / tp.segs_per_unit;
    const int part = s - unit * tp.segs_per_unit;
    q0 = static_cast<int>(static_cast<int64_t>(part) * tp.Lq / tp.segs_per_unit);
    q1 = static_cast<int>(static_cast<int64_t>(part + 1) * tp.Lq / tp.segs_per_unit);
}

__device__ __forceinline__ int ld_volatile_int(const int *p) { return *reinterpret_cast<const volatile int *>(p); }

// All threads of the CTA call this (it contains barriers).  Returns the next segment to work on or -1.
//  1. the untouched, unclaimed segment (index >= gridDim.x; the first gridDim.x segments are the CTAs'
//     home segments) nearest after this CTA's position, claimed with a CAS;
//  2. otherwise the segment with the most queries left (ties spread over the CTAs by a hash).
__device__ int tile_next_segment(const TileParams &tp, TileCtl &ctl)
{
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    int *claimed = tp.ctr + tp.n_segs;
    for (;;) {
        if (tid == 0) ctl.best = 0ull;
        __syncthreads();
        unsigned long long mine = 0ull;
        const int extra = tp.n_segs - G;
        if (extra > 0) {
            const int s0 = static_cast<int>(static_cast<int64_t>(b) * extra / G);
            for (int e = tid; e < extra; e += nt) {
                const int s = G + e;
                if (ld_volatile_int(claimed + s) == 0 && ld_volatile_int(tp.ctr + s) == 0) {
                    const int d = (e - s0 + extra) % extra;
                    const unsigned long long key = (1ull << 62) | (static_cast<unsigned long long>(0x3fffffff - d) << 24) |
                                                   static_cast<unsigned long long>(s);
                    mine = key > mine ? key : mine;
                }
            }
        }
        if (mine == 0ull) {
            for (int s = tid; s < tp.n_segs; s += nt) {
                int unit, q0, q1;
                seg_range(tp, s, unit, q0, q1);
                const int rem = (q1 - q0) - ld_volatile_int(tp.ctr + s);
                if (rem > 0) {
                    const unsigned prio = (static_cast<unsigned>(s) * 2654435761u + static_cast<unsigned>(b) * 40503u) >> 16;
                    const unsigned long long key = (static_cast<unsigned long long>(min(rem, 0xfffff)) << 40) |
                                                   (static_cast<unsigned long long>(prio & 0xffffu) << 24) |
                                                   static_cast<unsigned long long>(s);
                    mine = key > mine ? key : mine;
                }
            }
        }
        if (mine) atomicMax(&ctl.best, mine);
        __syncthreads();
        const unsigned long long best = ctl.best;
        if (best == 0ull) return -1;
        const int s = static_cast<int>(best & 0xffffffull);
        if (!(best >> 62)) return s;
        if (tid == 0) ctl.seg = (atomicCAS(claimed + s, 0, 1) == 0) ? s : -1;
        __syncthreads();
        const int got = ctl.seg;
        if (got >= 0) return got;
        // lost the race for that segment: look again
    }
}

// Warp-level ticket lock in shared memory (FIFO).  Lane 0 takes a ticket and polls a plain load.
// No memory-fence instruction on purpose: a MEMBAR here waits for ALL of the warp's outstanding
// memory operations, including the vector reductions of the fine levels that are queued behind a
// saturated L2 (measured: the critical section then lasts ~2400 cycles and the kernel is lock-bound).
// What makes the hand-over correct without it: every shared-memory access of the SM is executed in
// issue order by the one LSU pipeline, the unlocking store is issued after the section's stores
// (same warp, __syncwarp between them), and the next holder issues its loads only after its poll
// has returned the new ticket.  The asm statements are volatile with a memory clobber, so the
// compiler keeps the order as well.
__device__ __forceinline__ unsigned lds_volatile_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u32(unsigned *p, unsigned v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned warp_lock(TileLock *lk, int lane)
{
    unsigned my = 0;
    if (lane == 0) {
        asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], 1;" : "=r"(my) : "r"(smem_u32(&lk->next)) : "memory");
        // Only the next in line polls in a tight loop; everybody further back sleeps in proportion to its
        // distance, so that waiting warps do not take issue slots and LSU slots from the holder
        // (measured with 23 tight pollers: the 100-instruction section took ~2900 cycles).
        for (;;) {
            const unsigned d = my - lds_volatile_u32(&lk->serving);
            if (d == 0u) break;
            if (d > 1u) __nanosleep(d * MSDA_TILE_SLEEP_NS);
        }
    }
    __syncwarp();
    return my;
}
__device__ __forceinline__ void warp_unlock(TileLock *lk, int lane, unsigned my)
{
    __syncwarp();
    if (lane == 0) sts_volatile_u32(&lk->serving, my + 1u);
}

// Which levels live in shared memory: the longest suffix of levels whose pixels fit cap_rows.
__device__ __forceinline__ void coarse_split(const int64_t *__restrict__ lsi, int64_t S, int cap_rows,
                                             int &cf, int &cstart, int &n_coarse)
{
    cf = 4; cstart = static_cast<int>(S); n_coarse = 0;
#pragma unroll
    for (int l = 3; l >= 0; --l) {
        const int64_t st = __ldg(lsi + l);
        if (cf == l + 1 && st >= 0 && st <= S && S - st <= cap_rows) { cf = l; cstart = static_cast<int>(st); n_coarse = static_cast<int>(S - st); }
    }
}

// A point as written by its owner lane into the warp's table: (ta, tb, lw, packed).
__device__ __forceinline__ PointView view_from_table(const float4 d)
{
    const float hw = 1.f - d.z;
    PointView v;
    v.w0 = d.x * hw; v.w1 = d.x * d.z; v.w2 = d.y * hw; v.w3 = d.y * d.z;
    const int packed = __float_as_int(d.w);
    v.base = packed >> 4;
    v.mask = packed & 15;
    return v;
}

template <typename VT> struct TileIO;           // corner load of 4 channels per lane (8 lanes per head)
template <> struct TileIO<float> {
    static __device__ __forceinline__ float4 ld(const char *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ float4 ld_stream(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
};
template <> struct TileIO<__nv_bfloat16> {
    static __device__ __forceinline__ float4 cvt(uint2 t)
    {
        return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u),
                           __uint_as_float(t.y << 16), __uint_as_float(t.y & 0xffff0000u));
    }
    static __device__ __forceinline__ float4 ld(const char *p) { return cvt(__ldg(reinterpret_cast<const uint2 *>(p))); }
    static __device__ __forceinline__ float4 ld_stream(const __nv_bfloat16 *p) { return cvt(__ldcs(reinterpret_cast<const uint2 *>(p))); }
};

#ifndef MSDA_TILE_BWD_WARPS
#define MSDA_TILE_BWD_WARPS 16
#endif
#ifndef MSDA_TILE_FWD_WARPS
#define MSDA_TILE_FWD_WARPS 24
#endif

// --------------------------------------------------------------------------------------------
// Persistent backward.  VT = value / grad_out type (float or bf16), grad_value accumulates in fp32
// (in place for fp32 values, in the fp32 scratch for bf16 values).  Lane layout as the one-shot
// kernel: lane = g * 8 + j, g = head inside the unit's group of 4, j = channel quad.
// Dynamic shared memory: [(cap_rows + 1) x 512 B accumulator, the last row is scratch][NWARPS x 1 KB point tables].
// --------------------------------------------------------------------------------------------
template <typename VT, typename LT, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1)
bwd_tile_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                const int64_t *__restrict__ lsi, const LT *__restrict__ loc,
                const LT *__restrict__ aw, const VT *__restrict__ grad_out,
                float *__restrict__ grad_value, LT *__restrict__ grad_loc, LT *__restrict__ grad_aw,
                TileParams tp)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ TileCtl ctl;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 3, j = lane & 7;
    float4 *acc = reinterpret_cast<float4 *>(tile_smem);
    float4 *table = reinterpret_cast<float4 *>(tile_smem + static_cast<size_t>(tp.cap_rows + 1) * kTileRowBytes) + warp * 64;
    const int dummy = tp.cap_rows * 32 + lane;          // this lane's 16 bytes of the scratch row behind the map

    int cf, cstart, n_coarse;
    coarse_split(lsi, tp.S, tp.cap_rows, cf, cstart, n_coarse);

    // the lane's own two points are always points 2j, 2j+1 of its head: both of level j >> 1
    int myH, myW, myStart;
    level_dims(shapes, lsi, j >> 1, myH, myW, myStart);
    int rowW[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) rowW[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));

    if (tid < 4) { ctl.lock[tid].next = 0u; ctl.lock[tid].serving = 0u; }
    for (int i = tid; i < n_coarse * 32; i += NWARPS * 32) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    constexpr int pixb = 8 * kFastD * static_cast<int>(sizeof(VT));     // bytes between pixels of value
    constexpr int gpixb = 8 * kFastD * 4;                                // ... of grad_value (fp32)

    // ---- everything one warp does for one (frame, 4 heads, query) ----
    auto item = [&](int64_t n, int hg, int q) {
        const int64_t nq = n * tp.Lq + q;
        const int64_t qh0 = nq * 8 + hg * 4;
        float xy[4], a[2];
        PointIO<LT, 2>::load_loc(loc + qh0 * 32 + lane * 4, xy);
        PointIO<LT, 2>::load_aw(aw + qh0 * 16 + lane * 2, a);
        const float4 go = TileIO<VT>::ld_stream(grad_out + (qh0 + g) * kFastD + j * 4);

        OwnedPoint own[2];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            own[pp] = own_point(xy[2 * pp], xy[2 * pp + 1], a[pp], myH, myW, myStart);
            table[(2 * j + pp) * 4 + g] = make_float4(own[pp].ta, own[pp].tb, own[pp].lw, __int_as_float(own[pp].packed));
        }
        __syncwarp();

        const int64_t head_off = (n * tp.S * 8 + hg * 4 + g) * kFastD + j * 4;
        const char *vlane = reinterpret_cast<const char *>(value + head_off);
        char *glane = reinterpret_cast<char *>(grad_value + head_off);

        auto leaf = [&](auto pc, float (&d)[4]) {
            constexpr int p = decltype(pc)::value;
            const PointView pv = view_from_table(table[p * 4 + g]);
            const int W = rowW[p >> 2];
            const char *r0 = vlane + static_cast<int64_t>(pv.base) * pixb;
            const char *r1 = r0 + static_cast<int64_t>(W) * pixb;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;
            if (pv.mask & 1) v0 = TileIO<VT>::ld(r0);
            if (pv.mask & 2) v1 = TileIO<VT>::ld(r0 + pixb);
            if (pv.mask & 4) v2 = TileIO<VT>::ld(r1);
            if (pv.mask & 8) v3 = TileIO<VT>::ld(r1 + pixb);
            if ((p >> 2) < cf) {          // fine level: vector reductions into the L2
                char *g0 = glane + static_cast<int64_t>(pv.base) * gpixb;
                char *g1 = g0 + static_cast<int64_t>(W) * gpixb;
                red_add_v4_f32_if(pv.mask & 1, g0, pv.w0 * go.x, pv.w0 * go.y, pv.w0 * go.z, pv.w0 * go.w);
                red_add_v4_f32_if(pv.mask & 2, g0 + gpixb, pv.w1 * go.x, pv.w1 * go.y, pv.w1 * go.z, pv.w1 * go.w);
                red_add_v4_f32_if(pv.mask & 4, g1, pv.w2 * go.x, pv.w2 * go.y, pv.w2 * go.z, pv.w2 * go.w);
                red_add_v4_f32_if(pv.mask & 8, g1 + gpixb, pv.w3 * go.x, pv.w3 * go.y, pv.w3 * go.z, pv.w3 * go.w);
            }
            d[0] = go.x * v0.x + go.y * v0.y + go.z * v0.z + go.w * v0.w;
            d[1] = go.x * v1.x + go.y * v1.y + go.z * v1.z + go.w * v1.w;
            d[2] = go.x * v2.x + go.y * v2.y + go.z * v2.z + go.w * v2.w;
            d[3] = go.x * v3.x + go.y * v3.y + go.z * v3.z + go.w * v3.w;
        };

        float gxy[4], ga[2];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            float d[4];
            if (pp == 0) tree_reduce<4, 0, 2, 8>(j, leaf, d);
            else         tree_reduce<4, 1, 2, 8>(j, leaf, d);
            const OwnedPoint &o = own[pp];
            const float lh = o.lh, lw = o.lw, hh = 1.f - lh, hw = 1.f - lw;
            ga[pp] = hh * hw * d[0] + hh * lw * d[1] + lh * hw * d[2] + lh * lw * d[3];
            const float gw = hh * (d[1] - d[0]) + lh * (d[3] - d[2]);
            const float gh = hw * (d[2] - d[0]) + lw * (d[3] - d[1]);
            gxy[2 * pp]     = static_cast<float>(myW) * gw * o.a;
            gxy[2 * pp + 1] = static_cast<float>(myH) * gh * o.a;
        }
        PointIO<LT, 2>::store_loc(grad_loc + qh0 * 32 + lane * 4, gxy);
        PointIO<LT, 2>::store_aw(grad_aw + qh0 * 16 + lane * 2, ga);

        // ---- coarse levels: the 16 row updates of a level inside that level's critical section ----
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int l = (warp & 1) ? 3 - i : i;
            if (l < cf) continue;
#if MSDA_EXP == 20
            continue;      // timing experiment: coarse updates dropped (wrong results)
#endif
            const int W32 = __shfl_sync(kFull, myW, (l << 1) & 7) * 32;     // lane 2l owns a point of level l
            // masked-off corners go to a scratch row behind the map (their sum is never read): the section
            // then has no predicates, no divergence between the four head groups and no register shuffling
            float w[4][4];
            int i0[4], i1[4], i2[4], i3[4];
#pragma unroll
            for (int pi = 0; pi < 4; ++pi) {
                const PointView pv = view_from_table(table[(l * 4 + pi) * 4 + g]);
                const int idx = ((pv.base - cstart) * 4 + g) * 8 + j;
                w[pi][0] = pv.w0; w[pi][1] = pv.w1; w[pi][2] = pv.w2; w[pi][3] = pv.w3;
                i0[pi] = (pv.mask & 1) ? idx : dummy;
                i1[pi] = (pv.mask & 2) ? idx + 32 : dummy;
                i2[pi] = (pv.mask & 4) ? idx + W32 : dummy;
                i3[pi] = (pv.mask & 8) ? idx + W32 + 32 : dummy;
            }
            const unsigned ticket = warp_lock(&ctl.lock[l], lane);
#pragma unroll
            for (int pi = 0; pi < 4; ++pi) {
                float4 t0 = acc[i0[pi]], t1 = acc[i1[pi]], t2 = acc[i2[pi]], t3 = acc[i3[pi]];
                t0.x = fmaf(w[pi][0], go.x, t0.x); t0.y = fmaf(w[pi][0], go.y, t0.y); t0.z = fmaf(w[pi][0], go.z, t0.z); t0.w = fmaf(w[pi][0], go.w, t0.w);
                t1.x = fmaf(w[pi][1], go.x, t1.x); t1.y = fmaf(w[pi][1], go.y, t1.y); t1.z = fmaf(w[pi][1], go.z, t1.z); t1.w = fmaf(w[pi][1], go.w, t1.w);
                t2.x = fmaf(w[pi][2], go.x, t2.x); t2.y = fmaf(w[pi][2], go.y, t2.y); t2.z = fmaf(w[pi][2], go.z, t2.z); t2.w = fmaf(w[pi][2], go.w, t2.w);
                t3.x = fmaf(w[pi][3], go.x, t3.x); t3.y = fmaf(w[pi][3], go.y, t3.y); t3.z = fmaf(w[pi][3], go.z, t3.z); t3.w = fmaf(w[pi][3], go.w, t3.w);
                acc[i0[pi]] = t0; acc[i1[pi]] = t1; acc[i2[pi]] = t2; acc[i3[pi]] = t3;
            }
            warp_unlock(&ctl.lock[l], lane, ticket);
        }
        __syncwarp();     // the table is rewritten by the next item
    };

    // ---- flush the unit's coarse rows: one 512-byte bulk reduction per pixel, then re-zero ----
    auto flush = [&](int unit, bool last) {
        if (n_coarse == 0) return;
        fence_proxy_async_smem();
        __syncthreads();
        const int64_t n = unit >> 1;
        const int hg = unit & 1;
        float *gbase = grad_value + ((n * tp.S + cstart) * 8 + hg * 4) * kFastD;
        for (int px = tid; px < n_coarse; px += NWARPS * 32)
            bulk_reduce_add_f32(gbase + static_cast<int64_t>(px) * (8 * kFastD), acc + px * 32, kTileRowBytes);
        bulk_commit_group();
        if (last) { bulk_wait_group0(); return; }
        bulk_wait_group_read0();
        __syncthreads();
        for (int i = tid; i < n_coarse * 32; i += NWARPS * 32) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    };

    int seg = (static_cast<int>(blockIdx.x) < tp.n_segs) ? static_cast<int>(blockIdx.x) : -1;
    int cur_unit = -1;
    for (;;) {
        if (seg < 0) {
            seg = tile_next_segment(tp, ctl);
            if (seg < 0) break;
        }
        int unit, q0, q1;
        seg_range(tp, seg, unit, q0, q1);
        if (unit != cur_unit) {
            if (cur_unit >= 0) flush(cur_unit, false);
            cur_unit = unit;
        }
        const int64_t n = unit >> 1;
        const int hg = unit & 1;
        const int len = q1 - q0;
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(tp.ctr + seg, tp.batch);
        nxt = __shfl_sync(kFull, nxt, 0);
        while (nxt < len) {
            const int cur = nxt;
            if (lane == 0) nxt = atomicAdd(tp.ctr + seg, tp.batch);     // next grab in flight while this batch runs
            const int end = min(cur + tp.batch, len);
            for (int q = q0 + cur; q < q0 + end; ++q) item(n, hg, q);
            nxt = __shfl_sync(kFull, nxt, 0);
        }
        __syncthreads();
        seg = -1;
    }
    if (cur_unit >= 0) flush(cur_unit, true);
}

// --------------------------------------------------------------------------------------------
// Persistent forward (fp32 values).  Same schedule; the coarse levels' value rows of the unit are
// staged in shared memory by bulk copies, the fine levels are gathered through L1 as before.
// Dynamic shared memory: [cap_rows x 512 B values][NWARPS x 1 KB point tables].
// --------------------------------------------------------------------------------------------
template <typename LT, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1)
fwd_tile_kernel(const float *__restrict__ value, const int64_t *__restrict__ shapes,
                const int64_t *__restrict__ lsi, const LT *__restrict__ loc,
                const LT *__restrict__ aw, float *__restrict__ out, TileParams tp)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ TileCtl ctl;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 3, j = lane & 7;
    const float4 *sval = reinterpret_cast<const float4 *>(tile_smem);
    float4 *table = reinterpret_cast<float4 *>(tile_smem + static_cast<size_t>(tp.cap_rows) * kTileRowBytes) + warp * 64;

    int cf, cstart, n_coarse;
    coarse_split(lsi, tp.S, tp.cap_rows, cf, cstart, n_coarse);

    int myH, myW, myStart;
    level_dims(shapes, lsi, j >> 1, myH, myW, myStart);
    int rowW[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) rowW[l] = static_cast<int>(__ldg(shapes + 2 * l + 1));

    if (tid == 0) {
        mbar_init(&ctl.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;

    constexpr int pixb = 8 * kFastD * 4;

    auto item = [&](int64_t n, int hg, int q) {
        const int64_t nq = n * tp.Lq + q;
        const int64_t qh0 = nq * 8 + hg * 4;
        float xy[4], a[2];
        PointIO<LT, 2>::load_loc(loc + qh0 * 32 + lane * 4, xy);
        PointIO<LT, 2>::load_aw(aw + qh0 * 16 + lane * 2, a);
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            const OwnedPoint o = own_point(xy[2 * pp], xy[2 * pp + 1], a[pp], myH, myW, myStart);
            table[(2 * j + pp) * 4 + g] = make_float4(o.ta, o.tb, o.lw, __int_as_float(o.packed));
        }
        __syncwarp();

        const char *vlane = reinterpret_cast<const char *>(value + (n * tp.S * 8 + hg * 4 + g) * kFastD + j * 4);
        const float4 *slane = sval + g * 8 + j;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

        constexpr int NB = 2;
#pragma unroll
        for (int p0 = 0; p0 < kFastLP; p0 += NB) {
            PointView pv[NB];
            float4 v[NB][4];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int p = p0 + b;
                pv[b] = view_from_table(table[p * 4 + g]);
                const int W = rowW[p >> 2];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[b][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((p >> 2) < cf) {
                    const char *r0 = vlane + static_cast<int64_t>(pv[b].base) * pixb;
                    const char *r1 = r0 + static_cast<int64_t>(W) * pixb;
                    if (pv[b].mask & 1) v[b][0] = __ldg(reinterpret_cast<const float4 *>(r0));
                    if (pv[b].mask & 2) v[b][1] = __ldg(reinterpret_cast<const float4 *>(r0 + pixb));
                    if (pv[b].mask & 4) v[b][2] = __ldg(reinterpret_cast<const float4 *>(r1));
                    if (pv[b].mask & 8) v[b][3] = __ldg(reinterpret_cast<const float4 *>(r1 + pixb));
                } else {
                    const float4 *r0 = slane + (pv[b].base - cstart) * 32;
                    const float4 *r1 = r0 + W * 32;
                    if (pv[b].mask & 1) v[b][0] = r0[0];
                    if (pv[b].mask & 2) v[b][1] = r0[32];
                    if (pv[b].mask & 4) v[b][2] = r1[0];
                    if (pv[b].mask & 8) v[b][3] = r1[32];
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                acc.x = fmaf(pv[b].w0, v[b][0].x, acc.x); acc.y = fmaf(pv[b].w0, v[b][0].y, acc.y);
                acc.z = fmaf(pv[b].w0, v[b][0].z, acc.z); acc.w = fmaf(pv[b].w0, v[b][0].w, acc.w);
                acc.x = fmaf(pv[b].w1, v[b][1].x, acc.x); acc.y = fmaf(pv[b].w1, v[b][1].y, acc.y);
                acc.z = fmaf(pv[b].w1, v[b][1].z, acc.z); acc.w = fmaf(pv[b].w1, v[b][1].w, acc.w);
                acc.x = fmaf(pv[b].w2, v[b][2].x, acc.x); acc.y = fmaf(pv[b].w2, v[b][2].y, acc.y);
                acc.z = fmaf(pv[b].w2, v[b][2].z, acc.z); acc.w = fmaf(pv[b].w2, v[b][2].w, acc.w);
                acc.x = fmaf(pv[b].w3, v[b][3].x, acc.x); acc.y = fmaf(pv[b].w3, v[b][3].y, acc.y);
                acc.z = fmaf(pv[b].w3, v[b][3].z, acc.z); acc.w = fmaf(pv[b].w3, v[b][3].w, acc.w);
            }
        }
        __stcs(reinterpret_cast<float4 *>(out + (qh0 + g) * kFastD + j * 4), acc);
        __syncwarp();
    };

    // stage the coarse value rows of a unit: 512 contiguous bytes per pixel (4 heads), pixel stride 1 KB
    auto stage = [&](int unit) {
        if (n_coarse == 0) return;
        __syncthreads();          // nobody reads the previous unit's rows any more
        const int64_t n = unit >> 1;
        const int hg = unit & 1;
        if (tid == 0) mbar_arrive_expect_tx(&ctl.bar, static_cast<uint32_t>(n_coarse) * kTileRowBytes);
        const float *vbase = value + ((n * tp.S + cstart) * 8 + hg * 4) * kFastD;
        for (int px = tid; px < n_coarse; px += NWARPS * 32)
            bulk_load(tile_smem + static_cast<size_t>(px) * kTileRowBytes, vbase + static_cast<int64_t>(px) * (8 * kFastD),
                      kTileRowBytes, &ctl.bar);
        mbar_wait(&ctl.bar, phase);
        phase ^= 1u;
    };

    int seg = (static_cast<int>(blockIdx.x) < tp.n_segs) ? static_cast<int>(blockIdx.x) : -1;
    int cur_unit = -1;
    for (;;) {
        if (seg < 0) {
            seg = tile_next_segment(tp, ctl);
            if (seg < 0) break;
        }
        int unit, q0, q1;
        seg_range(tp, seg, unit, q0, q1);
        if (unit != cur_unit) {
            stage(unit);
            cur_unit = unit;
        }
        const int64_t n = unit >> 1;
        const int hg = unit & 1;
        const int len = q1 - q0;
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(tp.ctr + seg, tp.batch);
        nxt = __shfl_sync(kFull, nxt, 0);
        while (nxt < len) {
            const int cur = nxt;
            if (lane == 0) nxt = atomicAdd(tp.ctr + seg, tp.batch);
            const int end = min(cur + tp.batch, len);
            for (int q = q0 + cur; q < q0 + end; ++q) item(n, hg, q);
            nxt = __shfl_sync(kFull, nxt, 0);
        }
        __syncthreads();
        seg = -1;
    }
}

}  // namespace msda
