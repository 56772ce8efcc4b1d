// DRAM access-pattern probe for the slice coders (not part of the product): W warps each consume N pieces of 512 bytes.
//   pattern 0: group-major  - warp w owns one contiguous stream, piece p at (w * N + p) * 512   (the layout of vc2_common.cuh)
//   pattern 1: piece-major  - piece p of all warps is contiguous, (p * W + w) * 512
// `work` dependent integer operations per piece stand in for the coder's arithmetic; `stores` 4-byte stores per piece and
// lane go to a region of the lane's own (4 KB apart, like the packer's staging words).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/stream_probe.cu -o gpurun_out/stream_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128) probe(const int4* __restrict__ src, int* out, long long W, int N, int pattern, int work, int stores, unsigned* stage) {
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= W) return;
  int acc = 0;
  unsigned* wp = stage + ((w * 32 + lane) % (1 << 19)) * 1024;   // 4 KB per lane, 2 GB pool
  int4 nxt = __ldg(src + ((pattern ? w : w * N) * 32 + lane));
  for (int p = 0; p < N; ++p) {
    const int4 v = nxt;
    if (p + 1 < N) nxt = __ldg(src + ((pattern ? ((long long)(p + 1) * W + w) : (w * N + p + 1)) * 32 + lane));
    int x = v.x ^ v.y ^ v.z ^ v.w;
    for (int i = 0; i < work; ++i) x = x * 1664525 + 1013904223;
    acc += x;
    for (int i = 0; i < stores; ++i) *wp++ = (unsigned)x;
    if (((wp - stage) & 1023) > 1000) wp -= 1000;
  }
  if (acc == 0x12345678) out[0] = acc;
}
int main(int argc, char** argv) {
  const long long W = argc > 1 ? atoll(argv[1]) : 64800;   // 128 C3 pictures x 506.25 groups
  const int N = argc > 2 ? atoi(argv[2]) : 1024;
  const size_t bytes = (size_t)W * N * 512;
  int4* src; int* out; unsigned* stage;
  cudaMalloc(&src, bytes); cudaMalloc(&out, 4); cudaMemset(src, 1, bytes);
  cudaMalloc(&stage, (size_t)(1 << 19) * 4096 + 8192);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int work = 80; work <= 160; work += 80)
    for (int stores = 0; stores <= 2; ++stores) {
      const int pattern = 0;
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        probe<<<(unsigned)((W * 32 + 127) / 128), 128>>>(src, out, W, N, pattern, work, stores, stage);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      printf("work %3d stores/piece %d: %.3f ms  %.0f GB/s read  (%s)\n", work, stores, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
