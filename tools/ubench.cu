// Instruction-throughput probe for the integer pipes of sm_100a (not part of the product).  One CTA of 1024 threads per
// SM, every thread runs 8 independent chains of the same instruction; prints warp instructions per clock and SM.
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench.cu -o tools/_probe/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 512

struct OpIadd3   { static constexpr const char* name = "IADD3 (a+b+c)";        __device__ static int f(int x, int y, int z) { int r; asm volatile("{.reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, %3;}" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; } };
struct OpAdd     { static constexpr const char* name = "IADD (a+b)";           __device__ static int f(int x, int y, int z) { int r; asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; } };
struct OpImad    { static constexpr const char* name = "IMAD (a*b+c)";         __device__ static int f(int x, int y, int z) { int r; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; } };
struct OpImadImm { static constexpr const char* name = "IMAD imm (a*-9+c)";    __device__ static int f(int x, int y, int z) { int r; asm volatile("mad.lo.s32 %0, %1, -9, %2;" : "=r"(r) : "r"(x), "r"(z)); return r; } };
struct OpImadHi  { static constexpr const char* name = "IMAD.HI (mulhi+c)";    __device__ static int f(int x, int y, int z) { int r; asm volatile("mad.hi.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; } };
struct OpUmulHi  { static constexpr const char* name = "UMUL.HI";              __device__ static int f(int x, int y, int z) { int r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; } };
struct OpShrAdd  { static constexpr const char* name = "shr+add (LEA.HI.SX32?)"; __device__ static int f(int x, int y, int z) { int r; asm volatile("{.reg .s32 t; shr.s32 t, %1, 4; add.s32 %0, t, %2;}" : "=r"(r) : "r"(x), "r"(z)); return r; } };
struct OpShr     { static constexpr const char* name = "SHF.R.S32 (a>>4)";     __device__ static int f(int x, int y, int z) { int r; asm volatile("shr.s32 %0, %1, 4;" : "=r"(r) : "r"(x)); return r ^ z; } };
struct OpShrVar  { static constexpr const char* name = "SHF.R.U32 var";        __device__ static int f(int x, int y, int z) { int r; asm volatile("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; } };
struct OpPrmt    { static constexpr const char* name = "PRMT";                 __device__ static int f(int x, int y, int z) { int r; asm volatile("prmt.b32 %0, %1, %2, 0x4401;" : "=r"(r) : "r"(x), "r"(z)); return r; } };
struct OpLop3    { static constexpr const char* name = "LOP3";                 __device__ static int f(int x, int y, int z) { int r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; } };
struct OpAbs     { static constexpr const char* name = "IABS";                 __device__ static int f(int x, int y, int z) { int r; asm volatile("abs.s32 %0, %1;" : "=r"(r) : "r"(x)); return r - z; } };
struct OpMin     { static constexpr const char* name = "VIMNMX (min)";         __device__ static int f(int x, int y, int z) { int r; asm volatile("min.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; } };
struct OpClz     { static constexpr const char* name = "FLO (clz)";            __device__ static int f(int x, int y, int z) { int r; asm volatile("clz.b32 %0, %1;" : "=r"(r) : "r"(x)); return r + z; } };
struct OpPopc    { static constexpr const char* name = "POPC";                 __device__ static int f(int x, int y, int z) { int r; asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(x)); return r + z; } };
struct OpBfe     { static constexpr const char* name = "BFE.U32";              __device__ static int f(int x, int y, int z) { int r; asm volatile("bfe.u32 %0, %1, 6, 10;" : "=r"(r) : "r"(x)); return r + z; } };
struct OpSel     { static constexpr const char* name = "setp+selp";            __device__ static int f(int x, int y, int z) { int r; asm volatile("{.reg .pred p; setp.lt.s32 p, %1, %2; selp.s32 %0, %3, %1, p;}" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; } };
struct OpShfl    { static constexpr const char* name = "SHFL.IDX";             __device__ static int f(int x, int y, int z) { return __shfl_sync(0xFFFFFFFFu, x, y & 31); } };
struct OpDp4a    { static constexpr const char* name = "IDP4A";                __device__ static int f(int x, int y, int z) { int r; asm volatile("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; } };
struct OpDp2a    { static constexpr const char* name = "IDP2A.lo";             __device__ static int f(int x, int y, int z) { int r; asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; } };
struct OpVadd2   { static constexpr const char* name = "vadd2 (emulated?)";    __device__ static int f(int x, int y, int z) { return (int)__vadd2((unsigned)x, (unsigned)y); } };
struct OpCvtSat  { static constexpr const char* name = "cvt.sat.s16.s32";      __device__ static int f(int x, int y, int z) { int r; asm volatile("{.reg .s16 t; cvt.sat.s16.s32 t, %1; cvt.s32.s16 %0, t;}" : "=r"(r) : "r"(x)); return r + z; } };

template <class Op>
__global__ void __launch_bounds__(1024, 1) bench(int* out, long long* cycles, int y, int z) {
  int x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 7 + c;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = Op::f(x[c], y, z);
  }
  __syncthreads();
  const long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// mixed: a lifting-like step  o += (7 + a + d - 9*(b + c)) >> 4  written as the four instructions we hope for
__global__ void __launch_bounds__(1024, 1) bench_lift(int* out, long long* cycles, int y, int z) {
  int x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 7 + c;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      const int a = x[(c + 1) % CHAINS], b = x[(c + 2) % CHAINS], cc = x[(c + 3) % CHAINS], d = x[(c + 4) % CHAINS];
      const int t1 = b + cc;
      const int u = a + d + 7;
      const int w = t1 * -9 + u;
      x[c] += w >> 4;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared memory: 16-byte loads / stores, conflict free
__global__ void __launch_bounds__(1024, 1) bench_lds(int* out, long long* cycles, int store) {
  extern __shared__ int4 sm[];
  for (int i = threadIdx.x; i < 4096; i += 1024) sm[i] = make_int4(i, i, i, i);
  __syncthreads();
  int4 acc = make_int4(0, 0, 0, 0);
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      const int idx = (threadIdx.x + 1024 * (c & 3)) & 4095;
      if (store) sm[idx] = acc;
      else { const int4 q = sm[idx]; acc.x ^= q.x; acc.y ^= q.y; acc.z ^= q.z; acc.w ^= q.w; }
    }
    if (store) acc.x += i;
  }
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w ^ sm[threadIdx.x].x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int* d_out;
static long long* d_cyc;
static int nsm;

static void report(const char* name, double instr_per_thread_iter) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("%-28s error %s\n", name, cudaGetErrorString(e)); return; }
  static long long h[1024];
  cudaMemcpy(h, d_cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int i = 0; i < nsm; ++i) sum += (double)h[i];
  const double cyc = sum / nsm;
  const double winstr = 32.0 * ITERS * CHAINS * instr_per_thread_iter;   // warp instructions per CTA (= per SM)
  printf("%-28s %8.0f cycles  %6.3f warp-instr/clk/SM  (%5.1f lanes/clk/SM)\n", name, cyc, winstr / cyc, 32.0 * winstr / cyc);
}

template <class Op>
static void run(double n = 1.0) {
  bench<Op><<<nsm, 1024>>>(d_out, d_cyc, 3, 5);
  bench<Op><<<nsm, 1024>>>(d_out, d_cyc, 3, 5);
  report(Op::name, n);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  nsm = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, nsm);
  cudaMalloc(&d_out, sizeof(int) * 1024 * nsm);
  cudaMalloc(&d_cyc, sizeof(long long) * nsm);
  run<OpAdd>(); run<OpIadd3>(); run<OpImad>(); run<OpImadImm>(); run<OpImadHi>(); run<OpUmulHi>();
  run<OpShrAdd>(); run<OpShr>(2.0); run<OpShrVar>(); run<OpPrmt>(); run<OpLop3>(); run<OpAbs>(2.0); run<OpMin>();
  run<OpClz>(2.0); run<OpPopc>(2.0); run<OpBfe>(2.0); run<OpSel>(2.0); run<OpShfl>(); run<OpDp4a>(); run<OpDp2a>(); run<OpVadd2>();
  run<OpCvtSat>(2.0);
  bench_lift<<<nsm, 1024>>>(d_out, d_cyc, 3, 5);
  bench_lift<<<nsm, 1024>>>(d_out, d_cyc, 3, 5);
  report("lift step (4 instr hoped)", 4.0);
  cudaFuncSetAttribute(bench_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  bench_lds<<<nsm, 1024, 65536>>>(d_out, d_cyc, 0);
  bench_lds<<<nsm, 1024, 65536>>>(d_out, d_cyc, 0);
  report("LDS.128", 1.0);
  bench_lds<<<nsm, 1024, 65536>>>(d_out, d_cyc, 1);
  bench_lds<<<nsm, 1024, 65536>>>(d_out, d_cyc, 1);
  report("STS.128", 1.0);
  return 0;
}
