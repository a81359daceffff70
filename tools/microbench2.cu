// Atomic-rate microbenchmarks (B200): global RED by operand type, shared ATOMS by operand type.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
template <typename T> __global__ void red_kernel(const int* __restrict__ cam, long n, T* y) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    T* p = y + 9 * (long)cam[i];
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(p + k, (T)(1 + k));
  }
}
template <typename T> __global__ void smem_kernel(const int* __restrict__ cam, long n, int nx, T* out) {
  extern __shared__ unsigned char raw[];
  T* sy = reinterpret_cast<T*>(raw);
  for (int i = threadIdx.x; i < nx; i += blockDim.x) sy[i] = 0;
  __syncthreads();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    T* p = sy + 9 * cam[i];
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(p + k, (T)(1 + k));
  }
  __syncthreads();
  if (sy[threadIdx.x] == (T)123457) out[0] = sy[0];
}
template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}
template <typename T> void run(const char* name, const int* cam, long N, int C, int sms) {
  T* y; CK(cudaMalloc(&y, 9 * C * sizeof(T))); CK(cudaMemset(y, 0, 9 * C * sizeof(T)));
  float ms = timeit([&] { red_kernel<T><<<sms * 8, 128>>>(cam, N, y); });
  printf("global RED %-4s: %.3f ms  %.1f G lane-ops/s (%.2f cyc/lane/SM)\n", name, ms, 9.0 * N / ms / 1e6, 1.94e9 * sms * ms * 1e-3 / (9.0 * N));
  auto k = smem_kernel<T>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * C * sizeof(T)));
  ms = timeit([&] { k<<<sms, 512, 9 * C * sizeof(T)>>>(cam, N, 9 * C, y); });
  printf("shared ATOMS %-4s (1 CTA/SM x512): %.3f ms  %.1f G lane-ops/s (%.2f cyc/lane/SM)\n", name, ms, 9.0 * N / ms / 1e6, 1.94e9 * sms * ms * 1e-3 / (9.0 * N));
  cudaFree(y);
}
int main() {
  const int C = 1723; const long N = 8L << 20;
  std::vector<int> h(N); srand(1); for (long i = 0; i < N; ++i) h[i] = rand() % C;
  int* cam; CK(cudaMalloc(&cam, N * 4)); CK(cudaMemcpy(cam, h.data(), N * 4, cudaMemcpyHostToDevice));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<double>("f64", cam, N, C, sms);
  run<unsigned long long>("u64", cam, N, C, sms);
  run<float>("f32", cam, N, C, sms);
  run<unsigned int>("u32", cam, N, C, sms);
  return 0;
}
