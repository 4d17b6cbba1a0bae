// Probe of the FP64 tensor pipe on sm_100a: DMMA.8x8x4 issue rate per SM
// sub-partition as a function of resident warps and independent accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe.bin tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_probe(double* out, int iters, long long* cyc) {
  double c[ILP][2];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_smsp, double* out, long long* cyc) {
  const int threads = warps_per_smsp * 4 * 32, iters = 4096;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_probe<ILP><<<sms, threads>>>(out, iters, cyc);
  cudaEventRecord(e0);
  k_probe<ILP><<<sms, threads>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double n_dmma_per_smsp = (double)iters * ILP * warps_per_smsp;
  const double tf = 2.0 * 256 * n_dmma_per_smsp * 4 * sms / (ms * 1e-3) / 1e12;
  printf("{\"warps_per_smsp\": %d, \"ilp\": %d, \"cycles_per_dmma_per_smsp\": %.3f, \"tflops\": %.2f, \"ms\": %.3f}\n",
         warps_per_smsp, ILP, (double)h / n_dmma_per_smsp, tf, ms);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  for (int w = 1; w <= 4; ++w) {
    run<4>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc); run<32>(w, out, cyc);
  }
  return 0;
}
