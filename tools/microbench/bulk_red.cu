// Microbenchmark: is a 128-byte line update cheaper for the chip when it travels as ONE bulk
// reduction (cp.reduce.async.bulk, shared -> global, the TMA path) than as the 8 x 16-byte lanes of
// a red.global.add.v4.f32?  The backward scatter of MSDA is bound by the rate of line updates
// (DESIGN.md section 4), so this decides whether staging w*grad_out rows in shared memory pays.
// Developer tool; not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o bulk_red bulk_red.cu && ./bulk_red
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int STAGES = 4;

__device__ __forceinline__ void bulk_red_add_f32(float *gdst, const void *ssrc, unsigned bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 :: "l"(gdst), "r"(static_cast<unsigned>(__cvta_generic_to_shared(ssrc))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void red_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE 0: red.v4.f32, 4 scattered lines per warp instruction (what bwd_fast_kernel issues)
// MODE 1: STS.128 of the 4 lines + fence.proxy.async + 4 bulk reductions of 128 B (one per 8-lane group)
// MODE 2: like 1 but ONE bulk reduction of 512 contiguous bytes per warp step
// MODE 3: like 1, staging only (no reduction issued): cost of the shared-memory side alone
template <int MODE>
__global__ void __launch_bounds__(512) red_kernel(float *buf, unsigned span_lines, int iters)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 *stage = reinterpret_cast<float4 *>(smem) + wib * (STAGES * 32);       // STAGES x 512 B per warp
    unsigned s = warp * 2654435761u + 777u;
    const int q = lane >> 3;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        const unsigned r = s >> 7;
        const float v = 1.0f + (it & 3);
        if (MODE == 0) {
            const unsigned line = (r + q * 9973u) % span_lines;
            red_v4(buf + static_cast<size_t>(line) * 32 + (lane & 7) * 4, v, v, v, v);
        } else {
            const int st = it % STAGES;
            if (it >= STAGES) {
                if ((lane & 7) == 0) bulk_wait_read<STAGES - 1>();      // the group that last read this stage is done
                __syncwarp();
            }
            stage[st * 32 + lane] = make_float4(v, v, v, v);
            fence_async_smem();
            __syncwarp();
            if (MODE == 1 && (lane & 7) == 0) {
                const unsigned line = (r + q * 9973u) % span_lines;
                bulk_red_add_f32(buf + static_cast<size_t>(line) * 32, stage + st * 32 + q * 8, 128);
            }
            if (MODE == 2 && lane == 0) {
                const unsigned line = (r % (span_lines / 4)) * 4;
                bulk_red_add_f32(buf + static_cast<size_t>(line) * 32, stage + st * 32, 512);
            }
            if ((lane & 7) == 0) bulk_commit();
        }
    }
    if (MODE != 0 && (lane & 7) == 0) bulk_wait_read<0>();
}

template <typename F>
float time_ms(F f)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned span = 49u * 1024 * 1024 / 128;             // 49 MB of lines: grad_value of config 2
    float *buf;
    CK(cudaMalloc(&buf, static_cast<size_t>(span) * 128));
    CK(cudaMemset(buf, 0, static_cast<size_t>(span) * 128));
    const int iters = 1024, wpc = 16;
    const size_t smem = wpc * STAGES * 512;
    const char *names[4] = {"red.global.add.v4.f32, 4 scattered lines / warp step", "bulk reduce 4 x 128 B (scattered lines)",
                            "bulk reduce 1 x 512 B (4 contiguous lines)", "staging only (STS.128 + fence.proxy.async)"};
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
        const int ctas = sms * ctas_per_sm;
        for (int mode = 0; mode < 4; ++mode) {
            auto run = [&]() {
                switch (mode) {
                    case 0: red_kernel<0><<<ctas, wpc * 32, smem>>>(buf, span, iters); break;
                    case 1: red_kernel<1><<<ctas, wpc * 32, smem>>>(buf, span, iters); break;
                    case 2: red_kernel<2><<<ctas, wpc * 32, smem>>>(buf, span, iters); break;
                    case 3: red_kernel<3><<<ctas, wpc * 32, smem>>>(buf, span, iters); break;
                }
            };
            const float ms = time_ms(run);
            CK(cudaGetLastError());
            const double lines = static_cast<double>(ctas) * wpc * iters * 4;
            printf("%2d warps/SM  %-54s %8.3f ms  %7.2f Glines/s  %6.2f TB/s of line updates\n", ctas_per_sm * wpc, names[mode], ms,
                   lines / ms / 1e6, lines * 128 / ms / 1e9);
        }
    }
    // spot check: the sum of the buffer must equal what was added (bulk reductions are really applied)
    CK(cudaMemset(buf, 0, static_cast<size_t>(span) * 128));
    red_kernel<1><<<sms, wpc * 32, smem>>>(buf, span, 64);
    CK(cudaDeviceSynchronize());
    float *h = static_cast<float *>(malloc(static_cast<size_t>(span) * 128));
    CK(cudaMemcpy(h, buf, static_cast<size_t>(span) * 128, cudaMemcpyDeviceToHost));
    double sum = 0;
    for (size_t i = 0; i < static_cast<size_t>(span) * 32; ++i) sum += h[i];
    const double expect = static_cast<double>(sms) * wpc * 32 * 4 * (16 * (1 + 2 + 3 + 4));
    printf("bulk reduce checksum %.0f expected %.0f %s\n", sum, expect, sum == expect ? "OK" : "MISMATCH");
    return 0;
}
