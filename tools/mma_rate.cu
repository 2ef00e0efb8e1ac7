// Microbenchmark: issue-rate / shared-memory-operand-bandwidth ceiling of tcgen05.mma kind::f16 for the
// operand shapes the UNet kernels can use (M=128, N in {64,128,256}, SWIZZLE_NONE K-major operands in SMEM).
// One CTA per SM, one thread issues NMMA back-to-back MMAs over resident shared memory (no loads at all),
// so the number printed is the hardware ceiling of each formulation.  Build: see tools/Makefile.
#include "../deepsensornz_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <cstdlib>

void cnp_set_error(const char*, ...) {}

template <int N, int NACC, int SAME_A>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles, int b_mis, int b_lbo, int acc_stride) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = tc::smem_u32(smem), b0 = a0 + 96 * 1024;
    // A: 128 rows x 16 k: two K core-matrix columns LBO apart, 8-row groups SBO=128 B apart
    uint64_t ad[8], bd[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) ad[i] = tc::make_smem_desc(a0 + (SAME_A ? 0 : i * 8192), 4096, 128);
#pragma unroll
    for (int i = 0; i < 4; ++i) bd[i] = tc::make_smem_desc(b0 + b_mis + (b_lbo ? i * 64 : i * 2 * N * 16), b_lbo ? b_lbo : N * 16, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        tc::mma_bf16_ss(tm + (j % NACC) * (acc_stride ? acc_stride : N), ad[j % 8], bd[(j / NACC) % 4], idesc, 1u);
    }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc::fence_after_sync(); tc::tmem_dealloc(tm, 512); }
}

template <int N, int NACC, int SAME_A>
void run(const char* name, int iters, int b_mis = 0, int b_lbo = 0, int acc_stride = 0) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  auto k = rate_kernel<N, NACC, SAME_A>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148, 128, 200 * 1024>>>(iters, d, b_mis, b_lbo, acc_stride);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<148, 128, 200 * 1024>>>(iters, d, b_mis, b_lbo, acc_stride);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double nmma = 16.0 * iters;
  const double flops = 148.0 * nmma * 2.0 * 128 * N * 16;
  printf("%-34s N=%3d acc=%d  %6.1f cyc/MMA  (floor %3d)  %7.1f TFLOP/s  %.3f ms  %s\n", name, N, NACC, mx / nmma,
         128 * N / 256, flops / (ms * 1e-3) / 1e12, ms, err == cudaSuccess ? "" : cudaGetErrorString(err));
  cudaFree(d);
}

int main() {
  const int iters = 4096;
  run<64, 4, 0>("M128 N64  distinct A (current)", iters);
  run<64, 4, 1>("M128 N64  same A", iters);
  run<64, 1, 0>("M128 N64  one accumulator", iters);
  run<128, 2, 0>("M128 N128 distinct A", iters);
  run<128, 2, 1>("M128 N128 same A", iters);
  run<256, 2, 0>("M128 N256 distinct A", iters);
  run<256, 2, 1>("M128 N256 same A", iters);
  run<256, 1, 0>("M128 N256 one accumulator", iters);
  run<160, 3, 0>("M128 N160 aligned dense", iters);
  run<160, 3, 0>("M128 N160 B +16B", iters, 16);
  run<160, 3, 0>("M128 N160 B +48B", iters, 48);
  run<160, 3, 0>("M128 N160 B +64B", iters, 64);
  run<160, 3, 0>("M128 N160 LBO 25152", iters, 0, 25152);
  run<160, 3, 0>("M128 N160 LBO 25152 +16B", iters, 16, 25152);
  run<160, 3, 0>("M128 N160 LBO 25600 (x1024)", iters, 0, 25600);
  run<160, 3, 0>("M128 N160 acc stride 128?", iters, 0, 0, 170);
  run<128, 3, 0>("M128 N128 B +16B", iters, 16);
  run<256, 2, 0>("M128 N256 B +16B", iters, 16);
  run<64, 4, 0>("M128 N64 B +16B", iters, 16);
  return 0;
}
