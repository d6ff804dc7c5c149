// l2pin.cu -- microbenchmark: can a share of a 190 MB stream that is re-read every pass (the CG's K) be kept in
// B200's 126 MB L2?  Reads S bytes per pass; the first f*S bytes of every CTA's chunk use an "evict_last" load,
// the rest an "evict_first" load.  Prints effective GB/s (S / time per pass) for several mechanisms.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ double ld_pin(const double* p) {
    double v;
    if (MODE == 0) { v = __ldcs(p); }                                                     // no pinning (all streaming)
    else if (MODE == 1) { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
                          asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); }
    else if (MODE == 2) { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
                          asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); }
    else { v = __ldg(p); }                                                                // MODE 3: plain load (access policy window decides)
    return v;
}
template <int MODE> __device__ __forceinline__ double ld_str(const double* p) {
    double v;
    if (MODE == 1) { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                     asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); }
    else if (MODE == 4) { asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); }
    else { v = __ldcs(p); }
    return v;
}

// chunked: CTA b owns [b*chunk, (b+1)*chunk) doubles; first pin of them are "pinned"
template <int MODE>
__global__ void __launch_bounds__(256) k_read(const double* __restrict__ a, size_t chunk, size_t pin, double* out) {
    const double* base = a + (size_t)blockIdx.x * chunk;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    size_t i = threadIdx.x;
    for (; i + 768 < pin; i += 1024) {
        s0 += ld_pin<MODE>(base + i); s1 += ld_pin<MODE>(base + i + 256); s2 += ld_pin<MODE>(base + i + 512); s3 += ld_pin<MODE>(base + i + 768);
    }
    for (; i < pin; i += 256) s0 += ld_pin<MODE>(base + i);
    for (; i + 768 < chunk; i += 1024) {
        s0 += ld_str<MODE>(base + i); s1 += ld_str<MODE>(base + i + 256); s2 += ld_str<MODE>(base + i + 512); s3 += ld_str<MODE>(base + i + 768);
    }
    for (; i < chunk; i += 256) s0 += ld_str<MODE>(base + i);
    double s = s0 + s1 + s2 + s3;
    if (s == 1.2345e-300) out[0] = s;
}

template <int MODE>
float run(const double* a, size_t n, int grid, double f, double* out, cudaStream_t st, int passes) {
    size_t chunk = n / grid;
    size_t pin = (size_t)(f * chunk) / 256 * 256;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) k_read<MODE><<<grid, 256, 0, st>>>(a, chunk, pin, out);
    CK(cudaEventRecord(e0, st));
    for (int w = 0; w < passes; ++w) k_read<MODE><<<grid, 256, 0, st>>>(a, chunk, pin, out);
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / passes;
}

int main(int argc, char** argv) {
    const size_t MB = 1 << 20;
    size_t S = (argc > 1 ? atol(argv[1]) : 190) * MB;
    size_t n = S / 8;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s: L2 %.1f MB, persistingL2CacheMaxSize %.1f MB, accessPolicyMaxWindowSize %.1f MB, SMs %d\n", prop.name, prop.l2CacheSize / 1048576.0,
           prop.persistingL2CacheMaxSize / 1048576.0, prop.accessPolicyMaxWindowSize / 1048576.0, prop.multiProcessorCount);
    double *a, *out; CK(cudaMalloc(&a, S)); CK(cudaMalloc(&out, 8)); CK(cudaMemset(a, 0, S));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const int grid = prop.multiProcessorCount * 8;
    const int passes = 20;
    const double fs[] = {0.0, 0.15, 0.25, 0.33, 0.4, 0.5, 0.6};
    printf("S = %zu MB, grid %d x 256, chunked; GB/s effective per pass\n", S / MB, grid);
    printf("%-46s", "f (share of S marked evict_last):");
    for (double f : fs) printf(" %6.2f", f);
    printf("\n");
    auto row = [&](const char* name, auto fn) {
        printf("%-46s", name);
        for (double f : fs) { float ms = fn(f); printf(" %6.0f", S / 1e6 / ms); }
        printf("\n"); fflush(stdout);
    };
    row("mode0 all ld.cs (no pinning)", [&](double f) { return run<0>(a, n, grid, f, out, st, passes); });
    row("mode1 hint evict_last / hint evict_first", [&](double f) { return run<1>(a, n, grid, f, out, st, passes); });
    row("mode2 createpolicy evict_last hint / ld.cs", [&](double f) { return run<2>(a, n, grid, f, out, st, passes); });
    row("mode4 ldg / ld.L1::no_allocate (no L2 hints)", [&](double f) { return run<4>(a, n, grid, f, out, st, passes); });
    // persisting carve-out + the same hint loads
    size_t carve = prop.persistingL2CacheMaxSize;
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    printf("-- cudaLimitPersistingL2CacheSize = %.1f MB\n", carve / 1048576.0);
    row("mode1 with carve-out", [&](double f) { return run<1>(a, n, grid, f, out, st, passes); });
    row("mode2 with carve-out", [&](double f) { return run<2>(a, n, grid, f, out, st, passes); });
    // access policy window on the stream: persisting hit ratio over the whole buffer (window size limited)
    {
        printf("%-46s", "mode3 accessPolicyWindow hitRatio=f (plain ldg)");
        for (double f : fs) {
            cudaStreamAttrValue v{};
            size_t win = S < (size_t)prop.accessPolicyMaxWindowSize ? S : (size_t)prop.accessPolicyMaxWindowSize;
            v.accessPolicyWindow.base_ptr = a; v.accessPolicyWindow.num_bytes = win; v.accessPolicyWindow.hitRatio = (float)f;
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
            float ms = run<3>(a, n, grid, 1.0, out, st, passes);
            printf(" %6.0f", S / 1e6 / ms);
            CK(cudaCtxResetPersistingL2Cache());
        }
        printf("\n");
        cudaStreamAttrValue v{}; v.accessPolicyWindow.num_bytes = 0; CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
    }
    // smaller working sets for reference: what does an L2-resident stream reach?
    for (size_t s2 : {32 * MB, 64 * MB, 96 * MB, 112 * MB, 128 * MB}) {
        float ms = run<3>(a, s2 / 8, grid, 1.0, out, st, passes);
        printf("plain ldg, S = %3zu MB re-read every pass: %6.0f GB/s\n", s2 / MB, s2 / 1e6 / ms);
    }
    return 0;
}
