// Probe of the tcgen05.ld.16x256b register <-> (TMEM lane, column) mapping and of stmatrix.trans, used to design the
// conv epilogue's transposition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(float* out /*[2][32 threads][16 regs]*/, uint16_t* sm_out /*[32 px][40 halfwords]*/) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) uint16_t tile[32 * 40];
  const int lane = threadIdx.x;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  __syncwarp();
  const uint32_t tm = slot;
  // value(lane, col) = lane * 100 + col, written with the 32x32b shape (thread = lane)
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = __float_as_uint((float)(lane * 100 + c));
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(tm),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]));
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t packed[2][8];
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tm + ((uint32_t)(half * 16) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int k = 0; k < 16; ++k) out[(half * 32 + lane) * 16 + k] = __uint_as_float(r[k]);
    // pack register pairs (2k, 2k+1) as (lo = r[2k], hi = r[2k+1]) integers < 65536 for the stmatrix probe
    for (int k = 0; k < 8; ++k) packed[half][k] = ((uint32_t)__uint_as_float(r[2 * k]) & 0xffffu) | ((uint32_t)__uint_as_float(r[2 * k + 1]) << 16);
  }
  // stmatrix.x4.trans: matrices j = 0..3 from registers packed[half][2*cb + h8] ... probe with half 0, column blocks 0,1
  // thread t supplies the row address of matrix t/8, row t%8.  Rows of the destination = pixels, pitch 40 halfwords.
  for (int i = lane; i < 32 * 40; i += 32) tile[i] = 0xffff;
  __syncwarp();
  {
    // matrices: m0 = (lanes 0-7, cols 0-7) = packed[0][0]; m1 = (lanes 8-15, cols 0-7) = packed[0][1];
    //           m2 = (lanes 0-7, cols 8-15) = packed[0][2]; m3 = (lanes 8-15, cols 8-15) = packed[0][3]  (if the guess holds)
    const int mj = lane >> 3, rr = lane & 7;
    // transposed: matrix row index = pixel (col of the fragment).  dest row (pixel) = 8*(mj>>1) + rr, channel base = 8*(mj&1)
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&tile[(8 * (mj >> 1) + rr) * 40 + 8 * (mj & 1)]);
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(packed[0][0]),
                 "r"(packed[0][1]), "r"(packed[0][2]), "r"(packed[0][3]) : "memory");
  }
  __syncwarp();
  for (int i = lane; i < 32 * 40; i += 32) sm_out[i] = tile[i];
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

int main() {
  float* d; uint16_t* s;
  cudaMalloc(&d, 2 * 32 * 16 * 4); cudaMalloc(&s, 32 * 40 * 2);
  probe<<<1, 32>>>(d, s);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  float h[2 * 32 * 16]; uint16_t hs[32 * 40];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hs, s, sizeof(hs), cudaMemcpyDeviceToHost);
  for (int half = 0; half < 2; ++half)
    for (int t = 0; t < 32; t += (t < 8 ? 1 : 8)) {
      printf("half %d thread %2d:", half, t);
      for (int k = 0; k < 16; ++k) printf(" %5.0f", h[(half * 32 + t) * 16 + k]);
      printf("\n");
    }
  printf("stmatrix.trans result [pixel row][16 halfwords] (value = lane*100+col):\n");
  for (int p = 0; p < 16; ++p) { printf("px %2d:", p); for (int c = 0; c < 16; ++c) printf(" %5u", hs[p * 40 + c]); printf("\n"); }
  return 0;
}
