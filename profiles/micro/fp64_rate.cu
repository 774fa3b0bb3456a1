// Microbenchmark: FP64 vs FP32 FMA issue rate per SM on this GPU (nvcc -arch=sm_100a fp64_rate.cu -o fp64_rate)
#include <cstdio>
#include <cuda_runtime.h>
template <class T>
__global__ void fma_kernel(T* out, int iters) {
    T a0 = threadIdx.x * (T)1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T b = (T)1.0000001, c = (T)1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <class T>
double run(const char* name) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    T* out; cudaMalloc(&out, sizeof(T) * nsm * 8 * 256);
    const int iters = 20000;
    fma_kernel<T><<<nsm * 8, 256>>>(out, 1000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fma_kernel<T><<<nsm * 8, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)nsm * 8 * 256 * iters * 8;
    const double tflops = 2 * fmas / (ms * 1e-3) / 1e12;
    printf("%s: %.3f ms, %.2f TFLOP/s, %.1f FMA/clk/SM at %d MHz nominal\n", name, ms, tflops,
           fmas / (ms * 1e-3) / nsm / (khz * 1e3), khz / 1000);
    cudaFree(out);
    return tflops;
}
int main() { run<float>("fp32"); run<double>("fp64"); return 0; }
