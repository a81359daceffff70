// Microbenchmarks that size the design of the fused implicit-Schur kernel on B200 (run under gpurun):
//   red      : FP64 RED.ADD throughput into a camera-sized vector (random / clustered cameras)
//   ldg      : scattered 72 B gathers of x_cam from global (L1/L2)
//   lds      : the same gathers with x staged in shared memory
//   stream   : TMA bulk-copy streaming read of a large array through a 4-stage smem ring (HBM ceiling for this access pattern)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__global__ void red_kernel(const int* __restrict__ cam, long n, double* y) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double* p = y + 9 * (long)cam[i];
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(p + k, 1.0 + k);
  }
}
// warp-aggregated: lanes with equal camera pre-reduce through shuffles, leader issues the REDs
__global__ void red_agg_kernel(const int* __restrict__ cam, long n, double* y) {
  for (long i0 = (blockIdx.x * (long)blockDim.x + threadIdx.x); i0 < n + 32; i0 += (long)gridDim.x * blockDim.x) {
    const bool act = i0 < n;
    const int c = act ? cam[i0] : -1 - (threadIdx.x & 31);
    const unsigned m = __match_any_sync(0xffffffffu, c);
    const int leader = __ffs(m) - 1;
    const int lane = threadIdx.x & 31;
    double g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = 1.0 + k;
    // reduce within the match group: iterate over set bits (groups are small)
    unsigned rest = m & ~(1u << leader);
    while (__any_sync(0xffffffffu, rest != 0)) {
      const int src = rest ? __ffs(rest) - 1 : lane;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double v = __shfl_sync(0xffffffffu, g[k], src);
        if (lane == leader && rest) g[k] += v;
      }
      rest &= rest - 1;
    }
    if (act && lane == leader) {
      double* p = y + 9 * (long)c;
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(p + k, g[k]);
    }
  }
}
__global__ void ldg_kernel(const int* __restrict__ cam, long n, const double* __restrict__ x, double* out) {
  double acc = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double* p = x + 9 * (long)cam[i];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc += p[k];
  }
  if (acc == 123.456) out[0] = acc;
}
__global__ void lds_kernel(const int* __restrict__ cam, long n, const double* __restrict__ x, int nx, double* out) {
  extern __shared__ double sx[];
  for (int i = threadIdx.x; i < nx; i += blockDim.x) sx[i] = x[i];
  __syncthreads();
  double acc = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double* p = sx + 9 * cam[i];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc += p[k];
  }
  if (acc == 123.456) out[0] = acc;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES, int BYTES>
__global__ void stream_kernel(const char* __restrict__ src, long nchunks, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * BYTES);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long first = blockIdx.x, stride = gridDim.x;
  long issued = 0;
  auto issue = [&](long c, int s) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + s)), "r"(BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)s * BYTES)), "l"(src + c * BYTES), "r"(BYTES), "r"(s32(full + s)) : "memory");
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) { long c = first + issued * stride; if (c < nchunks) issue(c, s); ++issued; }
  double acc = 0;
  long k = 0;
  for (long c = first; c < nchunks; c += stride, ++k) {
    const int s = k % STAGES;
    const uint32_t parity = (k / STAGES) & 1;
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(full + s)), "r"(parity) : "memory");
    const double2* d = reinterpret_cast<const double2*>(smem + (size_t)s * BYTES);
    for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) { double2 v = d[i]; acc += v.x + v.y; }
    __syncthreads();
    if (threadIdx.x == 0) { long cn = first + issued * stride; if (cn < nchunks) issue(cn, s); ++issued; }
  }
  if (acc == 123.456) out[0] = acc;
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const int C = 1723; const long N = 8L << 20;
  std::vector<int> h(N);
  double *y, *x, *out; int* cam; char* big;
  CK(cudaMalloc(&y, 9 * C * 8)); CK(cudaMalloc(&x, 9 * C * 8)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cam, N * 4));
  CK(cudaMemset(y, 0, 9 * C * 8)); CK(cudaMemset(x, 0, 9 * C * 8));
  const long BIG = 2L << 30; CK(cudaMalloc(&big, BIG)); CK(cudaMemset(big, 1, BIG));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int mode = 0; mode < 3; ++mode) {
    // mode 0: uniformly random cameras; 1: runs of 4 equal cameras; 2: runs of 16
    const int run = mode == 0 ? 1 : (mode == 1 ? 4 : 16);
    srand(1);
    for (long i = 0; i < N; i += run) { int c = rand() % C; for (int j = 0; j < run && i + j < N; ++j) h[i + j] = c; }
    CK(cudaMemcpy(cam, h.data(), N * 4, cudaMemcpyHostToDevice));
    for (int bps : {4, 8, 16}) {
      float ms = timeit([&] { red_kernel<<<sms * bps, 128>>>(cam, N, y); });
      printf("red   run=%2d blocks/SM=%2d : %.3f ms  %.1f G lane-RED/s  (%.2f cyc/lane/SM @1.94GHz)\n", run, bps, ms, 9.0 * N / ms / 1e6, 1.94e9 * sms * ms * 1e-3 / (9.0 * N));
    }
    float ms = timeit([&] { red_agg_kernel<<<sms * 8, 128>>>(cam, N, y); });
    printf("redagg run=%2d              : %.3f ms  %.1f G row-equivalent lane-RED/s\n", run, ms, 9.0 * N / ms / 1e6);
    ms = timeit([&] { ldg_kernel<<<sms * 8, 128>>>(cam, N, x, out); });
    printf("ldg   run=%2d               : %.3f ms  %.1f G lane-LDG/s (%.2f cyc/lane/SM)\n", run, ms, 9.0 * N / ms / 1e6, 1.94e9 * sms * ms * 1e-3 / (9.0 * N));
    CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * C * 8));
    ms = timeit([&] { lds_kernel<<<sms, 512, 9 * C * 8>>>(cam, N, x, 9 * C, out); });
    printf("lds   run=%2d (1 CTA/SM x512): %.3f ms  %.1f G lane-LDS/s (%.2f cyc/lane/SM)\n", run, ms, 9.0 * N / ms / 1e6, 1.94e9 * sms * ms * 1e-3 / (9.0 * N));
  }
  {
    constexpr int BYTES = 24576;
    const long nch = BIG / BYTES;
    auto k4 = stream_kernel<4, BYTES>;
    CK(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * BYTES + 64));
    for (int bps : {1, 2}) {
      float ms = timeit([&] { k4<<<sms * bps, 256, 4 * BYTES + 64>>>(big, nch, out); });
      printf("stream 4x24KB stages, %d CTA/SM x256 thr: %.3f ms  %.1f GB/s\n", bps, ms, (double)nch * BYTES / ms / 1e6);
    }
    auto k8 = stream_kernel<8, BYTES>;
    CK(cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * BYTES + 64));
    float ms = timeit([&] { k8<<<sms, 512, 8 * BYTES + 64>>>(big, nch, out); });
    printf("stream 8x24KB stages, 1 CTA/SM x512 thr: %.3f ms  %.1f GB/s\n", ms, (double)nch * BYTES / ms / 1e6);
  }
  return 0;
}
