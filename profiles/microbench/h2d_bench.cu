// h2d_bench.cu -- what does this box's PCIe link give for pinned host -> device transfers, by mechanism?
// (context for bench.py's e2e leg, which is bound by exactly this)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o h2d_bench h2d_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(err__), __LINE__); exit(1); } } while (0)

// every thread keeps UNROLL 16 B loads from system memory in flight
template <int UNROLL>
__global__ void __launch_bounds__(256) k_pull(const int4* __restrict__ src, int4* __restrict__ dst, uint64_t nvec) {
  const uint64_t stride = (uint64_t)gridDim.x * 256 * UNROLL;
  for (uint64_t base = (uint64_t)blockIdx.x * 256 * UNROLL; base < nvec; base += stride) {
    int4 v[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) { uint64_t i = base + threadIdx.x + k * 256; if (i < nvec) v[k] = __ldcs(src + i); }
#pragma unroll
    for (int k = 0; k < UNROLL; k++) { uint64_t i = base + threadIdx.x + k * 256; if (i < nvec) dst[i] = v[k]; }
  }
}

static float timed(cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) { float ms; CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); return ms; }

int main() {
  const size_t N = 2ull << 30;
  char* h; char* d;
  CK(cudaMallocHost(&h, N)); CK(cudaMalloc(&d, N));
  for (size_t i = 0; i < N; i += 4096) h[i] = (char)i;
  cudaStream_t st[8]; for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaEvent_t e0, e1, ev[8]; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); for (auto& e : ev) CK(cudaEventCreate(&e));
  auto report = [&](const char* what, float ms, size_t bytes) { printf("%-70s %8.2f ms  %6.1f GB/s\n", what, ms, bytes / ms * 1e-6); };

  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(e0, st[0])); CK(cudaMemcpyAsync(d, h, N, cudaMemcpyHostToDevice, st[0])); CK(cudaEventRecord(e1, st[0]));
    float ms = timed(st[0], e0, e1); if (rep) report("cudaMemcpyAsync, one 2 GiB copy", ms, N);
  }
  for (size_t piece : {512ull << 10, 2ull << 20, 16ull << 20}) {
    for (int ns : {1, 2, 4}) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, st[0]));
      for (int k = 1; k < ns; k++) CK(cudaStreamWaitEvent(st[k], e0, 0));
      size_t n = N / piece;
      for (size_t i = 0; i < n; i++) CK(cudaMemcpyAsync(d + i * piece, h + i * piece, piece, cudaMemcpyHostToDevice, st[i % ns]));
      for (int k = 1; k < ns; k++) { CK(cudaEventRecord(ev[k], st[k])); CK(cudaStreamWaitEvent(st[0], ev[k], 0)); }
      CK(cudaEventRecord(e1, st[0]));
      char buf[128]; snprintf(buf, sizeof buf, "cudaMemcpyAsync, %zu KiB pieces round-robin on %d stream(s)", piece >> 10, ns);
      report(buf, timed(st[0], e0, e1), N);
    }
  }
  const int4* hs; CK(cudaHostGetDevicePointer((void**)&hs, h, 0));
  for (int ctas : {16, 32, 64, 128, 296, 592}) {
    CK(cudaEventRecord(e0, st[0])); k_pull<8><<<ctas, 256, 0, st[0]>>>(hs, (int4*)d, N / 16); CK(cudaEventRecord(e1, st[0]));
    char buf[128]; snprintf(buf, sizeof buf, "SM pull kernel, %d CTAs x 256 thr x 8 x 16 B in flight", ctas);
    report(buf, timed(st[0], e0, e1), N);
  }
  for (int ctas : {32, 64, 128}) {
    CK(cudaEventRecord(e0, st[0])); k_pull<16><<<ctas, 256, 0, st[0]>>>(hs, (int4*)d, N / 16); CK(cudaEventRecord(e1, st[0]));
    char buf[128]; snprintf(buf, sizeof buf, "SM pull kernel, %d CTAs x 256 thr x 16 x 16 B in flight", ctas);
    report(buf, timed(st[0], e0, e1), N);
  }
  // copy engine and SMs together on disjoint halves
  for (int ctas : {32, 64}) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, st[0])); CK(cudaStreamWaitEvent(st[1], e0, 0));
    CK(cudaMemcpyAsync(d, h, N / 2, cudaMemcpyHostToDevice, st[0]));
    k_pull<8><<<ctas, 256, 0, st[1]>>>(hs + N / 32, (int4*)(d + N / 2), N / 32);
    CK(cudaEventRecord(ev[1], st[1])); CK(cudaStreamWaitEvent(st[0], ev[1], 0)); CK(cudaEventRecord(e1, st[0]));
    char buf[128]; snprintf(buf, sizeof buf, "copy engine (1 GiB) + SM pull kernel %d CTAs (1 GiB) concurrently", ctas);
    report(buf, timed(st[0], e0, e1), N);
  }
  // D2H for reference
  CK(cudaEventRecord(e0, st[0])); CK(cudaMemcpyAsync(h, d, N, cudaMemcpyDeviceToHost, st[0])); CK(cudaEventRecord(e1, st[0]));
  report("cudaMemcpyAsync D2H, one 2 GiB copy", timed(st[0], e0, e1), N);
  // bidirectional
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0, st[0])); CK(cudaStreamWaitEvent(st[1], e0, 0));
  CK(cudaMemcpyAsync(d, h, N / 2, cudaMemcpyHostToDevice, st[0]));
  CK(cudaMemcpyAsync(h + N / 2, d + N / 2, N / 2, cudaMemcpyDeviceToHost, st[1]));
  CK(cudaEventRecord(ev[1], st[1])); CK(cudaStreamWaitEvent(st[0], ev[1], 0)); CK(cudaEventRecord(e1, st[0]));
  report("H2D 1 GiB + D2H 1 GiB concurrently (bytes = 2 GiB)", timed(st[0], e0, e1), N);
  return 0;
}
