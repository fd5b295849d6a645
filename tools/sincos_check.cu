// tools/sincos_check.cu -- is CUDA's sincos(x) bit-identical to (sin(x), cos(x)), and are sin odd / cos even bit for bit?
// K4 relies on both when it shares one sincos between the values the reference computes with separate calls.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/sincos_check tools/sincos_check.cu && /tmp/sincos_check
#include <cstdio>
#include <cstdint>
__device__ unsigned long long d_bad[4];
__device__ double rnd(uint64_t& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(s >> 11) * (1.0 / 9007199254740992.0); }
__global__ void k(int per, double scale) {
    uint64_t s = 0x9E3779B97F4A7C15ULL * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1);
    for (int i = 0; i < per; ++i) {
        const double x = (rnd(s) * 2.0 - 1.0) * scale;
        double sn, cs;
        sincos(x, &sn, &cs);
        if (__double_as_longlong(sn) != __double_as_longlong(sin(x))) atomicAdd(&d_bad[0], 1ULL);
        if (__double_as_longlong(cs) != __double_as_longlong(cos(x))) atomicAdd(&d_bad[1], 1ULL);
        if (__double_as_longlong(sin(-x)) != __double_as_longlong(-sin(x))) atomicAdd(&d_bad[2], 1ULL);
        if (__double_as_longlong(cos(-x)) != __double_as_longlong(cos(x))) atomicAdd(&d_bad[3], 1ULL);
    }
}
int main() {
    const double scales[] = {1e-3, 1.0, 3.2, 7.0, 100.0, 1e6};
    for (double sc : scales) {
        unsigned long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(d_bad, z, sizeof(z));
        k<<<1024, 256>>>(256, sc);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(z, d_bad, sizeof(z));
        printf("|x| <= %g: %d samples, sincos!=sin %llu, sincos!=cos %llu, sin not odd %llu, cos not even %llu\n", sc, 1024 * 256 * 256, z[0], z[1], z[2], z[3]);
    }
    return 0;
}
