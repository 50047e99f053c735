// Microbenchmark: accumulating 128-byte fp32 rows into a shared-memory map shared by all warps of a
// CTA (the "coarse levels stay on the SM" variant of the MSDA backward).
// Access shape of the kernel: a warp instruction carries 4 rows (one per 8-lane group), each lane
// 16 bytes.  sm_100a has no native shared fp32 add (atomicAdd(float) = ATOMS.CAST.SPIN loop), so
// the candidates are compare-and-swap loops of 32, 64 or 128 bits per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o smem_acc smem_acc.cu && ./smem_acc
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void add_cas32(float *p, float4 v)
{
    atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
}

__device__ __forceinline__ void add_cas64(float *p, float4 v)
{
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float a = h ? v.z : v.x, b = h ? v.w : v.y;
        unsigned long long old = q[h], assumed;
        do {
            assumed = old;
            const float lo = __uint_as_float(static_cast<unsigned>(assumed)) + a;
            const float hi = __uint_as_float(static_cast<unsigned>(assumed >> 32)) + b;
            const unsigned long long nv = (static_cast<unsigned long long>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
            old = atomicCAS(q + h, assumed, nv);
        } while (old != assumed);
    }
}

__device__ __forceinline__ void add_cas128(float *p, float4 v)
{
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    uint4 old = *reinterpret_cast<uint4 *>(p);
    while (true) {
        uint4 nv;
        nv.x = __float_as_uint(__uint_as_float(old.x) + v.x);
        nv.y = __float_as_uint(__uint_as_float(old.y) + v.y);
        nv.z = __float_as_uint(__uint_as_float(old.z) + v.z);
        nv.w = __float_as_uint(__uint_as_float(old.w) + v.w);
        uint4 got;
        asm volatile("{\n\t.reg .b128 c, n, d;\n\t"
                     "mov.b128 c, {%4, %5, %6, %7};\n\t"
                     "mov.b128 n, {%8, %9, %10, %11};\n\t"
                     "atom.shared.cas.b128 d, [%12], c, n;\n\t"
                     "mov.b128 {%0, %1, %2, %3}, d;\n\t}"
                     : "=r"(got.x), "=r"(got.y), "=r"(got.z), "=r"(got.w)
                     : "r"(old.x), "r"(old.y), "r"(old.z), "r"(old.w), "r"(nv.x), "r"(nv.y), "r"(nv.z), "r"(nv.w), "r"(a)
                     : "memory");
        if (got.x == old.x && got.y == old.y && got.z == old.z && got.w == old.w) break;
        old = got;
    }
}

// MODE 0: 4 x CAS.32 loops per lane (atomicAdd(float))   MODE 1: 2 x CAS.64   MODE 2: 1 x CAS.128
// MODE 3: plain LDS.128 + FADD + STS.128 (NOT race free; the floor if rows were privately owned)
template <int MODE>
__global__ void __launch_bounds__(768, 1) acc_kernel(float *out, int rows, int iters)
{
    extern __shared__ __align__(16) float map[];
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) map[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x >> 5) * 2654435761u + 99u;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        const int row = ((s >> 9) + g * 7919u) % rows;
        float *p = map + row * 32 + j * 4;
        const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        if (MODE == 0) add_cas32(p, v);
        else if (MODE == 1) add_cas64(p, v);
        else if (MODE == 2) add_cas128(p, v);
        else {
            float4 o = *reinterpret_cast<float4 *>(p);
            o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
            *reinterpret_cast<float4 *>(p) = o;
        }
    }
    __syncthreads();
    float acc = 0.f;
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) acc += map[i];
    atomicAdd(out + blockIdx.x, acc);
}

int main()
{
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const double ghz = khz / 1e6;
    float *out;
    CK(cudaMalloc(&out, sms * sizeof(float)));
    const int iters = 4096, threads = 768;
    const char *names[4] = {"4 x CAS.32 (atomicAdd float)", "2 x CAS.64", "1 x CAS.128", "LDS.128+FADD+STS.128 (racy floor)"};
    for (int rows : {1200, 240}) {                 // 300 px x 4 heads (levels 2+3 of level set A); 60 px x 4 heads (level 3 only)
        const size_t smem = static_cast<size_t>(rows) * 128;
        for (int mode = 0; mode < 4; ++mode) {
            auto launch = [&]() {
                switch (mode) {
                    case 0: CK(cudaFuncSetAttribute(acc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); acc_kernel<0><<<sms, threads, smem>>>(out, rows, iters); break;
                    case 1: CK(cudaFuncSetAttribute(acc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); acc_kernel<1><<<sms, threads, smem>>>(out, rows, iters); break;
                    case 2: CK(cudaFuncSetAttribute(acc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); acc_kernel<2><<<sms, threads, smem>>>(out, rows, iters); break;
                    case 3: CK(cudaFuncSetAttribute(acc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); acc_kernel<3><<<sms, threads, smem>>>(out, rows, iters); break;
                }
            };
            CK(cudaMemset(out, 0, sms * sizeof(float)));
            launch();
            CK(cudaDeviceSynchronize());
            float h0 = 0; CK(cudaMemcpy(&h0, out, 4, cudaMemcpyDeviceToHost));
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
            CK(cudaEventRecord(a));
            launch();
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            const double rows_per_sm = (double)(threads / 32) * iters * 4;
            const double expect = (double)(threads / 32) * iters * 4 * 8 * 10.0;     // every row update adds 8 lanes x (1+2+3+4)
            printf("%5d rows  %-36s %8.3f ms  %6.2f cyc/row-update/SM   checksum %s (%.0f vs %.0f)\n", rows, names[mode], ms,
                   ms * 1e-3 * ghz * 1e9 / rows_per_sm, (mode == 3 || h0 == (float)expect) ? "ok" : "MISMATCH", h0, expect);
        }
    }
    return 0;
}
