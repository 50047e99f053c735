// Microbenchmark: L1TEX throughput on B200 for the load/shuffle/reduction shapes the MSDA kernels
// can be built from.  Developer tool (results recorded in DESIGN.md); not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o l1_shapes l1_shapes.cu && ./l1_shapes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;

// Each warp walks over a private L1-resident window (8 KB = 64 lines) of `buf`.
// MODE 0: LDG.32, 32 lanes -> 1 line per instruction
// MODE 1: LDG.128, each quarter-warp reads a full line, the 4 lines are DIFFERENT (scattered)
// MODE 2: LDG.128, 512 contiguous bytes (4 consecutive lines)
// MODE 3: LDG.64, each half-warp reads a full line, 2 different lines
// MODE 4: LDG.128, 4 lanes x 16 B = 64 B per group, 8 different lines (bf16 layout)
// MODE 5: LDG.128 all quarter-warps read the SAME line (broadcast across quarters)
template <int MODE>
__global__ void ld_kernel(const float *__restrict__ buf, float *out, int window_lines)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const char *base = reinterpret_cast<const char *>(buf) + (size_t)warp * window_lines * 128;
    float acc = 0.f;
    unsigned s = warp * 2654435761u + 12345u;
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
        s = s * 1664525u + 1013904223u;
        const unsigned r = s >> 8;
        if (MODE == 0) {
            const int line = r & (window_lines - 1);
            acc += __ldg(reinterpret_cast<const float *>(base + line * 128) + lane);
        } else if (MODE == 1) {
            const int q = lane >> 3;
            const int line = (r + q * 17) & (window_lines - 1);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + line * 128) + (lane & 7));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 2) {
            const int line = (r & (window_lines / 4 - 1)) * 4;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + line * 128) + lane);
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 3) {
            const int h = lane >> 4;
            const int line = (r + h * 17) & (window_lines - 1);
            const float2 v = __ldg(reinterpret_cast<const float2 *>(base + line * 128) + (lane & 15));
            acc += v.x + v.y;
        } else if (MODE == 4) {
            const int g = lane >> 2;
            const int line = (r + g * 7) & (window_lines - 1);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + line * 128 + (g & 1) * 64) + (lane & 3));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 5) {
            const int line = r & (window_lines - 1);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + line * 128) + (lane & 7));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 6) {
            // LDG.256: 4 lanes x 32 B cover a line, 8 DIFFERENT lines per warp instruction
            const int g = lane >> 2;
            const int line = (r + g * 7) & (window_lines - 1);
            float a, b, c, d, e, f, gg, h;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(gg), "=f"(h)
                         : "l"(base + line * 128 + (lane & 3) * 32));
            acc += a + b + c + d + e + f + gg + h;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// shuffles only
__global__ void shfl_kernel(float *out)
{
    float v = threadIdx.x;
#pragma unroll 8
    for (int it = 0; it < ITERS * 4; ++it) v = __shfl_sync(0xffffffffu, v, (it * 7) & 7, 8) + 1.f;
    if (v == 123.456f) out[0] = v;
}

// vector reductions to L2.  MODE 0: red.v4.f32, quarter-warp per line, 4 different lines, addresses
// spread over `span_lines`; MODE 1: scalar red.f32, 32 lanes one line; MODE 2: red.v4 4 contiguous lines
// MODE 3: red.v2.f32 half-warp per line
template <int MODE>
__global__ void red_kernel(float *buf, size_t span_lines, int iters)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned s = warp * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        const size_t r = (s >> 4);
        if (MODE == 0) {
            const size_t line = (r + (lane >> 3) * 977) % span_lines;
            float *p = buf + line * 32 + (lane & 7) * 4;
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
        } else if (MODE == 1) {
            const size_t line = r % span_lines;
            atomicAdd(buf + line * 32 + lane, 1.0f);
        } else if (MODE == 2) {
            const size_t line = (r % (span_lines / 4)) * 4;
            float *p = buf + line * 32 + lane * 4;
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
        } else if (MODE == 3) {
            const size_t line = (r + (lane >> 4) * 977) % span_lines;
            float *p = buf + line * 32 + (lane & 15) * 2;
            asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
        }
    }
}

// shared-memory accumulation of 128-byte lines by whole warps (32 lanes = 32 floats of one line)
// MODE 0: atomicAdd(float) (CAS loop on sm_100a)  MODE 1: plain LDS + FADD + STS (race-free only if the
// warp owns the line)  MODE 2: atomicAdd(int) native
template <int MODE>
__global__ void smem_acc_kernel(float *out)
{
    __shared__ float tile[8192];                      // 32 KB = 256 lines
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tile[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) / 32 * 2654435761u + 99u;
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        s = s * 1664525u + 1013904223u;
        const int line = (s >> 8) & 255;
        float *p = tile + line * 32 + lane;
        if (MODE == 0) atomicAdd(p, 1.0f);
        else if (MODE == 1) *p += 1.0f;
        else atomicAdd(reinterpret_cast<int *>(p), 1);
    }
    __syncthreads();
    if (tile[threadIdx.x] == 123.456f) out[0] = 1.f;
}

template <typename F>
float time_ms(F f)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();                                   // warm
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main(int argc, char **argv)
{
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("SMs %d, clock %.0f MHz (max)\n", sms, khz / 1e3);
    const double ghz = khz / 1e6;

    const int warps_per_cta = 16, ctas = sms * 2;             // 32 warps / SM
    // lines per warp: 32 -> 4 KB/warp = 128 KB/SM, L1 resident; argv[1] = 256 -> 1 MB/SM: served by the L2
    const int window = argc > 1 ? atoi(argv[1]) : 32;
    const bool loads_only = argc > 1;
    float *buf, *out;
    const size_t nwarps = (size_t)ctas * warps_per_cta;
    CK(cudaMalloc(&buf, nwarps * window * 128));
    CK(cudaMemset(buf, 0, nwarps * window * 128));
    CK(cudaMalloc(&out, 4));
    const char *names[7] = {"LDG.32  1 line/instr", "LDG.128 4 scattered lines (fp32 layout)", "LDG.128 4 contiguous lines",
                            "LDG.64  2 scattered lines", "LDG.128 8 scattered half-lines (bf16 layout)", "LDG.128 same line x4 quarters",
                            "LDG.256 8 scattered lines (4 lanes x 32 B)"};
    const int lines_per[7] = {1, 4, 4, 2, 8, 1, 8};
    for (int mode = 0; mode < 7; ++mode) {
        auto run = [&]() {
            switch (mode) {
                case 0: ld_kernel<0><<<ctas, warps_per_cta * 32>>>(buf, out, window); break;
                case 1: ld_kernel<1><<<ctas, warps_per_cta * 32>>>(buf, out, window); break;
                case 2: ld_kernel<2><<<ctas, warps_per_cta * 32>>>(buf, out, window); break;
                case 3: ld_kernel<3><<<ctas, warps_per_cta * 32>>>(buf, out, window); break;
                case 4: ld_kernel<4><<<ctas, warps_per_cta * 32>>>(buf, out, window); break;
                case 5: ld_kernel<5><<<ctas, warps_per_cta * 32>>>(buf, out, window); break;
                case 6: ld_kernel<6><<<ctas, warps_per_cta * 32>>>(buf, out, window); break;
            }
        };
        const float ms = time_ms(run);
        const double instr_per_sm = (double)nwarps / sms * ITERS;
        const double cyc = ms * 1e-3 * ghz * 1e9;
        printf("%-46s %8.3f ms  %6.2f cyc/instr/SM  %6.2f cyc/line  %7.1f B/clk/SM\n", names[mode], ms,
               cyc / instr_per_sm, cyc / instr_per_sm / lines_per[mode],
               (mode == 4 ? 64.0 : 128.0) * lines_per[mode] * instr_per_sm / cyc);
    }
    if (loads_only) return 0;
    {
        const float ms = time_ms([&]() { shfl_kernel<<<ctas, warps_per_cta * 32>>>(out); });
        const double instr_per_sm = (double)nwarps / sms * ITERS * 4;
        printf("%-46s %8.3f ms  %6.2f cyc/instr/SM (32 warps/SM, dependent chain per warp)\n", "SHFL.IDX width 8", ms,
               ms * 1e-3 * ghz * 1e9 / instr_per_sm);
    }
    {
        const char *sn[3] = {"smem atomicAdd(float) 32 lanes/line (CAS)", "smem LDS+FADD+STS 32 lanes/line (no atomics)", "smem atomicAdd(int) 32 lanes/line (native)"};
        for (int mode = 0; mode < 3; ++mode) {
            auto run = [&]() {
                if (mode == 0) smem_acc_kernel<0><<<ctas, warps_per_cta * 32>>>(out);
                else if (mode == 1) smem_acc_kernel<1><<<ctas, warps_per_cta * 32>>>(out);
                else smem_acc_kernel<2><<<ctas, warps_per_cta * 32>>>(out);
            };
            const float ms = time_ms(run);
            const double instr_per_sm = (double)nwarps / sms * ITERS;
            printf("%-46s %8.3f ms  %6.2f cyc/line-update/SM\n", sn[mode], ms, ms * 1e-3 * ghz * 1e9 / instr_per_sm);
        }
    }
    // reductions
    const char *rnames[4] = {"RED.v4.f32 4 scattered lines", "RED.f32 scalar x32, 1 line", "RED.v4.f32 4 contiguous lines", "RED.v2.f32 2 scattered lines"};
    const int rlines[4] = {4, 1, 4, 2};
    const size_t spans[3] = {(size_t)1 << 12, (size_t)1 << 16, (size_t)1 << 20};     // 512 KB, 8 MB, 128 MB of lines
    float *rbuf;
    CK(cudaMalloc(&rbuf, spans[2] * 128));
    CK(cudaMemset(rbuf, 0, spans[2] * 128));
    const int riters = 512;
    // SM-side or L2-side limit?  Same kernel on a quarter / half / all of the SMs.
    for (int frac = 4; frac >= 1; frac /= 2) {
        const int c2 = ctas / frac;
        const float ms = time_ms([&]() { red_kernel<0><<<c2, warps_per_cta * 32>>>(rbuf, spans[1], riters); });
        const double lines = (double)c2 * warps_per_cta * riters * 4;
        printf("RED.v4.f32 scattered, %3d CTAs (1/%d of the SMs x2): %7.2f Glines/s\n", c2, frac, lines / ms / 1e6);
    }
    for (int wpc = 4; wpc <= 32; wpc *= 2) {
        const float ms = time_ms([&]() { red_kernel<0><<<sms, wpc * 32>>>(rbuf, spans[1], riters); });
        const double lines = (double)sms * wpc * riters * 4;
        printf("RED.v4.f32 scattered, 1 CTA/SM x %2d warps: %7.2f Glines/s\n", wpc, lines / ms / 1e6);
    }
    for (int mode = 0; mode < 4; ++mode)
        for (int si = 0; si < 3; ++si) {
            auto run = [&]() {
                switch (mode) {
                    case 0: red_kernel<0><<<ctas, warps_per_cta * 32>>>(rbuf, spans[si], riters); break;
                    case 1: red_kernel<1><<<ctas, warps_per_cta * 32>>>(rbuf, spans[si], riters); break;
                    case 2: red_kernel<2><<<ctas, warps_per_cta * 32>>>(rbuf, spans[si], riters); break;
                    case 3: red_kernel<3><<<ctas, warps_per_cta * 32>>>(rbuf, spans[si], riters); break;
                }
            };
            const float ms = time_ms(run);
            const double lines = (double)nwarps * riters * rlines[mode];
            printf("%-34s span %7.1f MB  %8.3f ms  %7.2f Glines/s  %7.2f TB/s of 128B-line updates\n", rnames[mode],
                   spans[si] * 128 / 1e6, ms, lines / ms / 1e6, lines * 128 / ms / 1e9);
        }
    return 0;
}
