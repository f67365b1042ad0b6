// Developer microbenchmarks for design decisions of the membership kernel (not product code):
// FP64 latency / throughput, shared-memory atomic throughput, random LDS.128 gather throughput.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void dp_latency(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = __dadd_rn(__dmul_rn(a, b), b); }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x + 2] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void dp_tput(double* out, long long* cyc, int n) {
  double a0 = out[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    a0 = __dadd_rn(__dmul_rn(a0, b), b); a1 = __dadd_rn(__dmul_rn(a1, b), b);
    a2 = __dadd_rn(__dmul_rn(a2, b), b); a3 = __dadd_rn(__dmul_rn(a3, b), b);
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x + 2] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void fp_latency(float* out, long long* cyc, int n) {
  float a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = __fadd_rn(__fmul_rn(a, b), b); }
  long long t1 = clock64();
  out[threadIdx.x + 2] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// every lane ORs into its own random word (spread addresses)
__global__ void atoms_spread(uint32_t* out, long long* cyc, int n) {
  __shared__ uint32_t tab[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = 0;
  __syncthreads();
  uint32_t h = threadIdx.x * 2654435761u;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { h = h * 1664525u + 1013904223u; atomicOr(&tab[(h >> 12) & 4095], 1u << (h & 31)); }
  __syncthreads();
  long long t1 = clock64();
  out[threadIdx.x] = tab[threadIdx.x];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// random 16-byte gathers, row stride `stride` words, then a conflict-free variant
__global__ void lds128_gather(uint32_t* out, long long* cyc, int n, int stride_words, int rows) {
  extern __shared__ uint4 tab4[];
  uint32_t* tab = reinterpret_cast<uint32_t*>(tab4);
  for (int i = threadIdx.x; i < rows * stride_words; i += blockDim.x) tab[i] = i;
  __syncthreads();
  uint32_t h = threadIdx.x * 2654435761u + 12345u;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    h = h * 1664525u + 1013904223u;
    const int r = (h >> 10) % rows;
    const uint4 v = *reinterpret_cast<const uint4*>(tab + r * stride_words);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds32_gather(uint32_t* out, long long* cyc, int n, int rows) {
  extern __shared__ uint4 tab4[];
  uint32_t* tab = reinterpret_cast<uint32_t*>(tab4);
  for (int i = threadIdx.x; i < rows; i += blockDim.x) tab[i] = i;
  __syncthreads();
  uint32_t h = threadIdx.x * 2654435761u + 12345u;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    h = h * 1664525u + 1013904223u;
    acc ^= tab[(h >> 10) % rows];
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* dout; long long* dcyc; float* fout; uint32_t* uout;
  cudaMalloc(&dout, 8 * (2 + 148 * 1024)); cudaMalloc(&dcyc, 8 * 148); cudaMalloc(&fout, 4 * 2048); cudaMalloc(&uout, 4 * 2048);
  cudaMemset(dout, 0, 8 * (2 + 148 * 1024)); cudaMemset(fout, 0, 4 * 2048);
  long long c[148];
  const int n = 4096;
  for (int threads : {32, 256, 1024}) {
    dp_latency<<<1, threads>>>(dout, dcyc, n); cudaMemcpy(c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("dp dependent chain (dmul+dadd), %4d threads/SM: %.2f cycles per op\n", threads, (double)c[0] / (2.0 * n));
    dp_tput<<<1, threads>>>(dout, dcyc, n); cudaMemcpy(c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("dp 4 independent chains,          %4d threads/SM: %.2f cycles per warp-op per SM -> %.1f lanes/clk/SM\n", threads,
           (double)c[0] / (8.0 * n * (threads / 32)), 32.0 * 8.0 * n * (threads / 32) / (double)c[0]);
  }
  fp_latency<<<1, 32>>>(fout, dcyc, n); cudaMemcpy(c, dcyc, 8, cudaMemcpyDeviceToHost);
  printf("fp32 dependent chain: %.2f cycles per op\n", (double)c[0] / (2.0 * n));
  for (int threads : {32, 256, 1024}) {
    atoms_spread<<<1, threads>>>(uout, dcyc, 1024); cudaMemcpy(c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("smem atomicOr spread, %4d threads: %.1f cycles per warp-instr per SM\n", threads, (double)c[0] / (1024.0 * (threads / 32)));
  }
  cudaFuncSetAttribute(lds128_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int stride : {4, 8, 12, 16, 20}) {
    lds128_gather<<<1, 1024, 128 * stride * 4>>>(uout, dcyc, 2048, stride, 128); cudaMemcpy(c, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.128 random row gather, row stride %2d words, 1024 threads: %.1f cycles per warp-instr per SM\n", stride,
           (double)c[0] / (2048.0 * 32));
  }
  lds32_gather<<<1, 1024, 4096 * 4>>>(uout, dcyc, 2048, 4096); cudaMemcpy(c, dcyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS.32 random gather (4096 words), 1024 threads: %.1f cycles per warp-instr per SM\n", (double)c[0] / (2048.0 * 32));
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
