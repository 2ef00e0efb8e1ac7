// bf16 UNet path for sm_100a: tcgen05 implicit-GEMM convolution (forward, dgrad) plus the
// memory-bound helpers around it (weight packing, layout conversion, 1x1 input conv, bilinear x2
// resize forward/backward, space-to-depth for the stride-2 layers).
//
// Replaces the torch Conv2d stack of upstream neuralprocesses' UNet (reached
// from ConvNP.loss_fn nzdownscale/downscaler/train.py:370 and train_epoch train.py:388-394) for the
// bf16 mode of BASELINE.json's north_star ("UNet conv stack as tcgen05/TMA implicit-GEMM in bf16").
//
// ---- activation layout in HBM ("blocked") ---------------------------------------------------
//   [B][C/8][H+4][W+4][8] bf16 : channels in chunks of 8 (16 B per pixel-chunk), every plane
//   physically zero-padded by 2 pixels on each side.  The pad is zeroed once at allocation and
//   never written, so 5x5 / pad-2 convolutions need no boundary logic.
//
// ---- implicit GEMM ----------------------------------------------------------------------------
//   A tile = a (TH+4) x pitch window of the padded input (pitch = TW+4), for 4 channel chunks,
//   copied row by row with cp.async.bulk into shared memory as [chunk][window pixel][8 ch].
//   That is exactly the SWIZZLE_NONE K-major UMMA layout (core matrix = 8 consecutive pixels x
//   8 channels, 16 B rows), and a conv tap (ky,kx) is just a +(ky*pitch+kx)*16 B shift of the
//   descriptor start address: no im2col, every input byte is fetched from L2 once per tile.
//   M = 128 consecutive window pixels per accumulator, R <= 4 accumulators per tile (TMEM
//   columns [64r, 64r+64)), double-buffered across tiles (2 x 256 columns).  Window pixels that
//   are not valid outputs (the 4 halo columns per row) are computed and dropped: rows of D only
//   depend on the matching rows of A, so they cannot contaminate valid outputs.
//   N = 64 output channels, K = 32 input channels x taps per stage; weights are pre-packed per
//   stage as [tap][k8][n][8] (K-major, SWIZZLE_NONE) and streamed through a 3-deep ring.
//   Warp roles: 0 = bulk-copy producer, 1 = MMA issuer (one elected thread), 2 = TMEM allocator,
//   4..7 = epilogue (tcgen05.ld -> bias/ReLU/mask -> bf16 blocked or fp32 NCHW stores).
#include "tc_common.cuh"

#define CNP_MAX_AB 8
#define CNP_MAX_ST 24
#define CNP_MAX_TAP 5

struct cnp_conv_plan {
  int n_ab;
  int ab_chunk0[CNP_MAX_AB];       // first of the 4 source chunks of this A-block
  int ab_wci0[CNP_MAX_AB];         // weight input-channel base of this A-block (for packing)
  int ab_stage0[CNP_MAX_AB + 1];   // stages [ab_stage0[i], ab_stage0[i+1]) belong to A-block i
  int st_ntap[CNP_MAX_ST];
  short st_aoff[CNP_MAX_ST][CNP_MAX_TAP];  // pixel offset of the tap inside the A window
  signed char st_wky[CNP_MAX_ST][CNP_MAX_TAP];
  signed char st_wkx[CNP_MAX_ST][CNP_MAX_TAP];
};

struct cnp_conv_args {
  const __nv_bfloat16* x; long long x_bs; int x_Hp, x_Wp;
  const __nv_bfloat16* w;
  int B, H, W;
  int TW, TH, pitch, R, plane_sm, tiles_x, tiles_y;
  int out_mode;                  // 0: blocked bf16, 1: NCHW fp32
  void* out; long long out_bs; int out_c_off; int out_Hp, out_Wp;
  int sy, ay, sx, ax;            // output pixel = (y*sy+ay, x*sx+ax)
  const float* bias; int relu;
  const __nv_bfloat16* mask; long long mask_bs; int mask_cb_off;
  int accumulate;
  cnp_conv_plan plan;
};

namespace {

constexpr int N_OUT = 64;
constexpr int W_TAP_BYTES = 4 * N_OUT * 16;             // 4 k-chunks x 64 n x 16 B = 4096
constexpr int W_STAGE_BYTES = CNP_MAX_TAP * W_TAP_BYTES;  // 20480
constexpr int W_STAGES = 3;
constexpr int A_BUFS = 2;
constexpr int ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = 512;

#ifdef CNP_LEGACY_CONV_TC   // first formulation (pixels = M, N = 64), kept for A/B runs: make LEGACY=1
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ cnp_conv_args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int plane_bytes = a.plane_sm * 16;
  const int abuf_bytes = 4 * plane_bytes;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + A_BUFS * abuf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + W_STAGES * W_STAGE_BYTES);
  uint64_t* a_full = bars;             // [2]
  uint64_t* a_empty = bars + 2;        // [2]
  uint64_t* w_full = bars + 4;         // [3]
  uint64_t* w_empty = bars + 7;        // [3]
  uint64_t* acc_full = bars + 10;      // [2]
  uint64_t* acc_empty = bars + 12;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = a.B * a.tiles_x * a.tiles_y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(a_full + i, 1); tc::mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 3; ++i) { tc::mbar_init(w_full + i, 1); tc::mbar_init(w_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 4); }
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int rows = a.TH + 4;
  const uint32_t row_bytes = (uint32_t)a.pitch * 16u;
  const long long plane_g = (long long)a.x_Hp * a.x_Wp * 8;  // elements per chunk plane

  if (warp == 0) {
    // ===================== producer: bulk copies of A windows and weight stages ================
    if (tc::elect_one()) {
      uint32_t a_it = 0, w_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / (a.tiles_x * a.tiles_y), tr = tile % (a.tiles_x * a.tiles_y);
        const int y0 = (tr / a.tiles_x) * a.TH, x0 = (tr % a.tiles_x) * a.TW;
        const __nv_bfloat16* xb = a.x + (long long)b * a.x_bs + ((long long)y0 * a.x_Wp + x0) * 8;
        for (int ab = 0; ab < a.plan.n_ab; ++ab, ++a_it) {
          const int buf = a_it & 1;
          tc::mbar_wait(a_empty + buf, ((a_it >> 1) & 1) ^ 1);
          tc::mbar_expect_tx(a_full + buf, 4u * rows * row_bytes);
          uint8_t* dst = a_smem + buf * abuf_bytes;
          for (int c = 0; c < 4; ++c) {
            const __nv_bfloat16* src = xb + (long long)(a.plan.ab_chunk0[ab] + c) * plane_g;
            for (int r = 0; r < rows; ++r)
              tc::bulk_g2s(dst + c * plane_bytes + r * row_bytes, src + (long long)r * a.x_Wp * 8, row_bytes,
                           a_full + buf);
          }
          for (int s = a.plan.ab_stage0[ab]; s < a.plan.ab_stage0[ab + 1]; ++s, ++w_it) {
            const int ws = w_it % W_STAGES;
            tc::mbar_wait(w_empty + ws, ((w_it / W_STAGES) & 1) ^ 1);
            const uint32_t bytes = (uint32_t)a.plan.st_ntap[s] * W_TAP_BYTES;
            tc::mbar_expect_tx(w_full + ws, bytes);
            tc::bulk_g2s(w_smem + ws * W_STAGE_BYTES,
                         reinterpret_cast<const uint8_t*>(a.w) + (size_t)s * W_STAGE_BYTES, bytes, w_full + ws);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ===========================================================
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, N_OUT, 0, 0);
      uint32_t a_it = 0, w_it = 0, t_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_it) {
        const int as = t_it & 1;
        tc::mbar_wait(acc_empty + as, ((t_it >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t d0 = tmem_base + as * (TMEM_COLS / ACC_STAGES);
        bool first = true;
        for (int ab = 0; ab < a.plan.n_ab; ++ab, ++a_it) {
          const int buf = a_it & 1;
          tc::mbar_wait(a_full + buf, (a_it >> 1) & 1);
          const uint32_t a_base = tc::smem_u32(a_smem + buf * abuf_bytes);
          for (int s = a.plan.ab_stage0[ab]; s < a.plan.ab_stage0[ab + 1]; ++s, ++w_it) {
            const int ws = w_it % W_STAGES;
            tc::mbar_wait(w_full + ws, (w_it / W_STAGES) & 1);
            tc::fence_after_sync();
            const uint32_t w_base = tc::smem_u32(w_smem + ws * W_STAGE_BYTES);
            const int ntap = a.plan.st_ntap[s];
            for (int t = 0; t < ntap; ++t) {
              const uint32_t aoff = (uint32_t)a.plan.st_aoff[s][t] * 16u;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t bdesc = tc::make_smem_desc(w_base + t * W_TAP_BYTES + ks * (2 * N_OUT * 16),
                                                          N_OUT * 16, 128);
                for (int r = 0; r < a.R; ++r) {
                  const uint64_t adesc = tc::make_smem_desc(a_base + 2 * ks * plane_bytes + aoff + r * 2048,
                                                            plane_bytes, 128);
                  tc::mma_bf16_ss(d0 + r * N_OUT, adesc, bdesc, idesc, first ? 0u : 1u);
                }
                first = false;
              }
            }
            tc::mma_commit(w_empty + ws);
          }
          tc::mma_commit(a_empty + buf);
        }
        tc::mma_commit(acc_full + as);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue ==============================================================
    const int q = warp & 3;
    uint32_t t_it = 0;
    const long long oplane = (long long)a.out_Hp * a.out_Wp;  // blocked: pixels per chunk plane
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_it) {
      const int as = t_it & 1;
      const int b = tile / (a.tiles_x * a.tiles_y), tr = tile % (a.tiles_x * a.tiles_y);
      const int y0 = (tr / a.tiles_x) * a.TH, x0 = (tr % a.tiles_x) * a.TW;
      tc::mbar_wait(acc_full + as, (t_it >> 1) & 1);
      tc::fence_after_sync();
      for (int r = 0; r < a.R; ++r) {
        const int m = r * 128 + q * 32 + lane;
        const int ty = m / a.pitch, tx = m - ty * a.pitch;
        const int y = y0 + ty, x = x0 + tx;
        const bool valid = (ty < a.TH) && (tx < a.TW) && (y < a.H) && (x < a.W);
        const int oy = y * a.sy + a.ay, ox = x * a.sx + a.ax;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * (TMEM_COLS / ACC_STAGES) + r * N_OUT;
        if (a.out_mode == 0) {
          const long long pix = (long long)(oy + 2) * a.out_Wp + (ox + 2);
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(a.out) + (long long)b * a.out_bs +
                                 ((long long)a.out_c_off * oplane + pix) * 8;
          // issue every mask / read-modify-write load of this pixel up front (independent 16 B loads in
          // flight) so the epilogue is bandwidth- rather than latency-bound
          uint4 mk[8], old[8];
          if (a.mask && valid) {
            const __nv_bfloat16* mbase = a.mask + (long long)b * a.mask_bs + ((long long)a.mask_cb_off * oplane + pix) * 8;
#pragma unroll
            for (int c = 0; c < 8; ++c) mk[c] = __ldg(reinterpret_cast<const uint4*>(mbase + (long long)c * oplane * 8));
          }
          if (a.accumulate && valid) {
#pragma unroll
            for (int c = 0; c < 8; ++c) old[c] = *reinterpret_cast<const uint4*>(obase + (long long)c * oplane * 8);
          }
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            float v[32];
            tc::tmem_ld32(taddr + hc * 32, v);
            tc::tmem_ld_wait();
            if (a.bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += __ldg(a.bias + hc * 32 + i);
            }
            if (a.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = v[i] < 0.f ? 0.f : v[i];
            }
            if (valid) {
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const int chunk = hc * 4 + c8;
                float* vv = v + c8 * 8;
                if (a.mask) {
                  const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&mk[chunk]);
#pragma unroll
                  for (int i = 0; i < 8; ++i) if (!(__bfloat162float(mb[i]) > 0.f)) vv[i] = 0.f;
                }
                if (a.accumulate) {
                  const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(&old[chunk]);
#pragma unroll
                  for (int i = 0; i < 8; ++i) vv[i] += __bfloat162float(ob[i]);
                }
                uint4 pk;
                __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(vv[2 * i], vv[2 * i + 1]);
                *reinterpret_cast<uint4*>(obase + (long long)chunk * oplane * 8) = pk;
              }
            }
          }
        } else {
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            float v[32];
            tc::tmem_ld32(taddr + hc * 32, v);
            tc::tmem_ld_wait();
            if (a.bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += __ldg(a.bias + hc * 32 + i);
            }
            if (a.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = v[i] < 0.f ? 0.f : v[i];
            }
            if (valid) {
              float* dst = reinterpret_cast<float*>(a.out) + (long long)b * a.out_bs +
                           ((long long)(a.out_c_off + hc * 32) * a.out_Hp + oy) * a.out_Wp + ox;
              const long long cs = (long long)a.out_Hp * a.out_Wp;
#pragma unroll
              for (int i = 0; i < 32; ++i) dst[i * cs] = a.accumulate ? dst[i * cs] + v[i] : v[i];
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty + as);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc::fence_after_sync(); tc::tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// weight packing: torch fp32 [Cout][Cin][k][k] -> bf16 stages [stage][tap][k8][n][8]
//   transposed = 0: n = output channel (co_off + n), k = input channel ab_wci0 + k8*8 + c
//   transposed = 1 (dgrad): n = original input channel (co_off + n), k = original output channel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int transposed, int co_off,
                    __nv_bfloat16* __restrict__ wpk, const __grid_constant__ cnp_conv_plan plan) {
  const int nst = plan.ab_stage0[plan.n_ab];
  const int total = nst * CNP_MAX_TAP * 4 * N_OUT * 8;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
    const int c = e & 7, n = (e >> 3) & 63, k8 = (e >> 9) & 3, t = (e >> 11) % CNP_MAX_TAP, s = (e >> 11) / CNP_MAX_TAP;
    int ab = 0;
    while (ab + 1 < plan.n_ab && s >= plan.ab_stage0[ab + 1]) ++ab;
    float v = 0.f;
    if (t < plan.st_ntap[s]) {
      const int kc = plan.ab_wci0[ab] + k8 * 8 + c, nn = co_off + n;
      const int ky = plan.st_wky[s][t], kx = plan.st_wkx[s][t];
      if (!transposed) {
        if (nn < Cout && kc < Cin) v = w[(((size_t)nn * Cin + kc) * k + ky) * k + kx];
      } else {
        if (kc < Cout && nn < Cin) v = w[(((size_t)kc * Cin + nn) * k + ky) * k + kx];
      }
    }
    wpk[e] = __float2bfloat16_rn(v);
  }
}
#endif  // CNP_LEGACY_CONV_TC

// ---------------------------------------------------------------------------------------------
// layout conversion  NCHW fp32 <-> blocked bf16  (tests, encoder/decoder interfacing)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_blk_kernel(const float* __restrict__ src, long long src_bs, int C, int H, int W,
                   __nv_bfloat16* __restrict__ dst, long long dst_bs, int cb_off, int ones_ch,
                   unsigned long long shared_mask) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int Hp = H + 4, Wp = W + 4;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H * W; e += gridDim.x * 256) {
    const int y = e / W, x = e % W;
    uint4 pk;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = chunk * 8 + i;
      // channels flagged in shared_mask hold one field for the whole batch: read batch 0 (broadcast on the fly)
      const size_t bb = ((shared_mask >> (c & 63)) & 1ull) ? 0 : (size_t)b;
      v[i] = c < C ? src[bb * src_bs + ((size_t)c * H + y) * W + x] : (c == ones_ch ? 1.f : 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(dst + (size_t)b * dst_bs + (((size_t)(cb_off + chunk) * Hp + y + 2) * Wp + x + 2) * 8) = pk;
  }
}

__global__ void __launch_bounds__(256)
blk_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, long long src_bs, int cb_off, int C, int H, int W,
                   float* __restrict__ dst, long long dst_bs) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int Hp = H + 4, Wp = W + 4;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H * W; e += gridDim.x * 256) {
    const int y = e / W, x = e % W;
    const uint4 pk = *reinterpret_cast<const uint4*>(
        src + (size_t)b * src_bs + (((size_t)(cb_off + chunk) * Hp + y + 2) * Wp + x + 2) * 8);
    const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&pk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = chunk * 8 + i;
      if (c < C) dst[(size_t)b * dst_bs + ((size_t)c * H + y) * W + x] = __bfloat162float(pb[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// initial 1x1 conv: fp32 NCHW [B,Cin,H,W] (encoder output) -> blocked bf16 [B,Cout/8,..] (+bias)
// ---------------------------------------------------------------------------------------------
constexpr int IC_MAX_CIN = 40;

__global__ void __launch_bounds__(256)
conv1x1_in_kernel(const float* __restrict__ x, long long x_bs, int Cin, int H, int W,
                  const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                  __nv_bfloat16* __restrict__ out, long long out_bs, int cb_off) {
  __shared__ __align__(16) float w_s[IC_MAX_CIN][64];
  __shared__ __align__(16) float b_s[64];
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < Cin * 64; e += 256) {
    int ci = e / 64, co = e % 64;
    w_s[ci][co] = co < Cout ? w[(size_t)co * Cin + ci] : 0.f;
  }
  if (threadIdx.x < 64) b_s[threadIdx.x] = threadIdx.x < Cout ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int Hp = H + 4, Wp = W + 4;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H * W; e += gridDim.x * 256) {
    const int y = e / W, xx = e % W;
    float acc[64];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 bq = reinterpret_cast<const float4*>(b_s)[q];
      acc[4 * q] = bq.x; acc[4 * q + 1] = bq.y; acc[4 * q + 2] = bq.z; acc[4 * q + 3] = bq.w;
    }
    for (int c0 = 0; c0 < Cin; c0 += 8) {
      float xv[8];   // 8 independent loads in flight before the FMAs
#pragma unroll
      for (int k = 0; k < 8; ++k)
        xv[k] = (c0 + k < Cin) ? __ldg(x + (size_t)b * x_bs + (size_t)(c0 + k) * H * W + e) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (c0 + k < Cin) {
          const float v = xv[k];
          const float4* wr = reinterpret_cast<const float4*>(w_s[c0 + k]);   // broadcast 16 B reads
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 wq = wr[q];
            acc[4 * q] = fmaf(v, wq.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, wq.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, wq.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, wq.w, acc[4 * q + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int chunk = 0; chunk < 8; ++chunk) {
      uint4 pk;
      __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(acc[chunk * 8 + 2 * i], acc[chunk * 8 + 2 * i + 1]);
      *reinterpret_cast<uint4*>(out + (size_t)b * out_bs + (((size_t)(cb_off + chunk) * Hp + y + 2) * Wp + xx + 2) * 8) = pk;
    }
  }
}

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* v) {
  const uint4 pk = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&pk);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(pb[i]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
  uint4 pk;
  __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = pk;
}

// ---------------------------------------------------------------------------------------------
// weight/bias gradient of the initial 1x1 conv: dw[co][ci] += sum_px dy[co,px] * x[ci,px]
//   x fp32 NCHW (encoder output, Cin <= 16), dy blocked bf16 (64 channels).
//   Tile = 256 pixels staged in shared memory (dy as raw bf16 [px][64], x as fp32 [px][16]); warp w reduces
//   pixels w, w+8, ...; lane = (8 output channels, 4 input channels): 3 LDS.128 per 32 FMA, so the kernel is
//   FMA / HBM bound instead of LDS bound.  Partial sums meet in shared memory, one atomic per output per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int IW_PX = 256;
constexpr int IW_MAXCI = 32;   // input channels incl. the constant-1 slot that carries the bias gradient

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;   // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NH>   // NH = number of 16-channel halves of x in use
__global__ void __launch_bounds__(256)
conv1x1_in_wgrad_kernel(const float* __restrict__ x, long long x_bs, int Cin, int H, int W,
                        const __nv_bfloat16* __restrict__ dy, long long dy_bs, int dy_cb, int B,
                        float* __restrict__ dw, float* __restrict__ dbias) {
  extern __shared__ __align__(16) uint8_t iw_smem[];
  constexpr int NCI = 16 * NH;
  constexpr int STAGE_B = IW_PX * 64 * 2 + IW_PX * NCI * 4;     // dy tile (bf16) + x tile (fp32), two stages (LDGSTS)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co8 = lane >> 2, ci4 = lane & 3;
  const int HW = H * W, Hp = H + 4, Wp = W + 4;
  const int tiles_per_img = (HW + IW_PX - 1) / IW_PX;
  const int ntiles = B * tiles_per_img;
  const int one_slot = NCI - 1;                 // x channel holding the constant 1 (Cin < NCI is guaranteed)
  float acc[NH][8][4];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh)
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[hh][a][c] = 0.f;

  auto issue = [&](int t, int stage) {
    __nv_bfloat16 (*dy_s)[64] = reinterpret_cast<__nv_bfloat16 (*)[64]>(iw_smem + stage * STAGE_B);
    float (*x_s)[NCI] = reinterpret_cast<float (*)[NCI]>(iw_smem + stage * STAGE_B + IW_PX * 64 * 2);
    const int b = t / tiles_per_img, e0 = (t % tiles_per_img) * IW_PX;
    for (int e = threadIdx.x; e < IW_PX * 8; e += 256) {
      const int px = e & (IW_PX - 1), chunk = e >> 8, pe = e0 + px;
      const bool ok = pe < HW;
      const int yy = ok ? pe / W : 0, xx = ok ? pe - yy * W : 0;
      cp_async16(&dy_s[px][chunk * 8], dy + (size_t)b * dy_bs + (((size_t)(dy_cb + chunk) * Hp + yy + 2) * Wp + xx + 2) * 8, ok);
    }
    for (int e = threadIdx.x; e < NCI * IW_PX; e += 256) {
      const int ci = e >> 8, px = e & (IW_PX - 1), pe = e0 + px;
      if (ci < Cin) cp_async4(&x_s[px][ci], x + (size_t)b * x_bs + (size_t)ci * HW + (pe < HW ? pe : 0), pe < HW);
      else x_s[px][ci] = (ci == one_slot && pe < HW) ? 1.f : 0.f;
    }
    cp_async_commit();
  };

  int stage = 0;
  if ((int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, stage ^= 1) {
    const int tn = t + gridDim.x;
    if (tn < ntiles) { issue(tn, stage ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const __nv_bfloat16 (*dy_s)[64] = reinterpret_cast<const __nv_bfloat16 (*)[64]>(iw_smem + stage * STAGE_B);
    const float (*x_s)[NCI] = reinterpret_cast<const float (*)[NCI]>(iw_smem + stage * STAGE_B + IW_PX * 64 * 2);
#pragma unroll 2
    for (int px = warp; px < IW_PX; px += 8) {
      const uint4 g = *reinterpret_cast<const uint4*>(&dy_s[px][co8 * 8]);
      const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
      float gf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f2 = __bfloat1622float2(g2[i]); gf[2 * i] = f2.x; gf[2 * i + 1] = f2.y; }
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        const float4 xv = *reinterpret_cast<const float4*>(&x_s[px][hh * 16 + ci4 * 4]);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          acc[hh][a][0] = fmaf(gf[a], xv.x, acc[hh][a][0]); acc[hh][a][1] = fmaf(gf[a], xv.y, acc[hh][a][1]);
          acc[hh][a][2] = fmaf(gf[a], xv.z, acc[hh][a][2]); acc[hh][a][3] = fmaf(gf[a], xv.w, acc[hh][a][3]);
        }
      }
    }
    __syncthreads();   // the stage just consumed is refilled by the next iteration's issue()
  }
  // reduce the 8 warps through shared memory ([8][64][16] floats = 32 KB per half), one atomic per output
  float* red = reinterpret_cast<float*>(iw_smem);
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) red[(warp * 64 + co8 * 8 + a) * 16 + ci4 * 4 + c] = acc[hh][a][c];
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * 1024 + e];
      const int co = e >> 4, ci = hh * 16 + (e & 15);
      if (ci < Cin) atomicAdd(dw + (size_t)co * Cin + ci, s);
      else if (ci == one_slot && dbias) atomicAdd(dbias + co, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear x2 (align_corners=False) on blocked tensors; backward optionally masks by (act > 0)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void up2_src(int Y, int H, int* y0, int* y1, float* lam) {
  float src = ((float)Y + 0.5f) * 0.5f - 0.5f;
  src = fmaxf(src, 0.f);
  int i0 = (int)src;
  *y0 = i0; *y1 = min(i0 + 1, H - 1); *lam = src - (float)i0;
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): two adjacent 8-channel pixels of a blocked row per instruction,
// so that a warp's access covers whole 32 B sectors (two 16 B accesses at a 32 B stride each touch half a sector)
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void unpack8(const uint32_t* r, float* v) {   // 4 words = 8 bf16 -> fp32 (exact)
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(r[i] << 16); v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u); }
}
__device__ __forceinline__ void pack8(const float* v, uint32_t* r) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    r[i] = *reinterpret_cast<const uint32_t*>(&t);
  }
}

// forward: one thread per LOW-res pixel (u,v) and chunk: 3x3 source neighbourhood -> the 2x2 output block.
//   Y = 2u   : 0.25 x[u-1] + 0.75 x[u]   (u-1 clamped: at u = 0 the source coordinate clamps to 0 -> x[0])
//   Y = 2u+1 : 0.75 x[u]   + 0.25 x[u+1] (u+1 clamped)
// evaluated exactly like up2_src (lambda = 0.75 / 0.25, or 0 at the clamped border): horizontally once per source
// row (even / odd output column), then vertically; each output row leaves as one 32 B store.
__global__ void __launch_bounds__(256)
blk_upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long x_bs, int x_cb, int H, int W,
                          __nv_bfloat16* __restrict__ y, long long y_bs, int y_cb) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int Wp = W + 4, Hp = H + 4, W2p = 2 * W + 4, H2p = 2 * H + 4;
  const __nv_bfloat16* xc = x + (size_t)b * x_bs + (size_t)(x_cb + chunk) * Hp * Wp * 8;
  __nv_bfloat16* yc = y + (size_t)b * y_bs + (size_t)(y_cb + chunk) * H2p * W2p * 8;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H * W; e += gridDim.x * 256) {
    const int u = e / W, v = e - u * W;
    const int um = max(u - 1, 0), up = min(u + 1, H - 1), vm = max(v - 1, 0), vp = min(v + 1, W - 1);
    const int rr[3] = {um, u, up}, cc[3] = {vm, v, vp};
    uint4 raw[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        raw[i][j] = __ldg(reinterpret_cast<const uint4*>(xc + ((size_t)(rr[i] + 2) * Wp + cc[j] + 2) * 8));
    // lambda of the even output (towards index u from u-1) and of the odd one (towards u+1 from u)
    const float ly0 = u > 0 ? 0.75f : 0.f, ly1 = 0.25f, lx0 = v > 0 ? 0.75f : 0.f, lx1 = 0.25f;
    float ev[3][8], od[3][8];   // per source row: even / odd output column
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float a0[8], a1[8], a2[8];
      unpack8(reinterpret_cast<const uint32_t*>(&raw[i][0]), a0);
      unpack8(reinterpret_cast<const uint32_t*>(&raw[i][1]), a1);
      unpack8(reinterpret_cast<const uint32_t*>(&raw[i][2]), a2);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ev[i][k] = (1.f - lx0) * a0[k] + lx0 * a1[k];
        od[i][k] = (1.f - lx1) * a1[k] + lx1 * a2[k];
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int r0 = dy, r1 = dy + 1;                        // source rows (y0, y1) of this output row
      const float ly = dy == 0 ? ly0 : ly1;
      float o0[8], o1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o0[k] = (1.f - ly) * ev[r0][k] + ly * ev[r1][k];
        o1[k] = (1.f - ly) * od[r0][k] + ly * od[r1][k];
      }
      uint32_t pk[8];
      pack8(o0, pk);
      pack8(o1, pk + 4);
      stg256(yc + ((size_t)(2 * u + dy + 2) * W2p + 2 * v + 2) * 8, pk);
    }
  }
}

// dx[u,v] = sum over the <=4x4 hi-res pixels that reference (u,v); optional ReLU mask by act>0.  The 4 tap weights
// per dimension are (0.25, 0.75, 0.75, 0.25) away from the borders and follow up2_src at them.
__device__ __forceinline__ float up2_wgt(int Y, int H, int u) {
  if (Y < 0 || Y >= 2 * H) return 0.f;
  int y0, y1; float l;
  up2_src(Y, H, &y0, &y1, &l);
  return (y0 == u ? (1.f - l) : 0.f) + (y1 == u ? l : 0.f);
}

// Row kernel (W <= 1024): one block per low-res row u and chunk, thread v = low-res column.  Each thread loads only
// ITS two hi-res columns (2v, 2v+1) of the four rows -- one 32 B load per row -- and combines them vertically
// (V0, V1); the horizontal taps of the neighbours (V1 of v-1, V0 of v+1) come through shared memory.  Per thread:
// 4 loads and 64 conversions instead of 16 and 128 (the per-pixel kernel below was issue-bound at 1.75 TB/s).
// rows_per_block low-res rows per block: hi-res rows 2u+1, 2u+2 of one row are rows 2u'-1, 2u' of the next
__global__ void __launch_bounds__(512)
blk_upsample2x_bwd_row_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_bs, int dy_cb, int H, int W,
                              __nv_bfloat16* __restrict__ dx, long long dx_bs, int dx_cb,
                              const __nv_bfloat16* __restrict__ act, long long act_bs, int act_cb, int accumulate,
                              int rows_per_block) {
  extern __shared__ __align__(16) float4 up_sm[];          // [4][blockDim.x]: V0 lo, V0 hi, V1 lo, V1 hi
  const int chunk = blockIdx.y, b = blockIdx.z, v = threadIdx.x, T = blockDim.x;
  const int u0 = blockIdx.x * rows_per_block, u1 = min(u0 + rows_per_block, H);
  const int Wp = W + 4, Hp = H + 4, W2p = 2 * W + 4, H2p = 2 * H + 4;
  const __nv_bfloat16* dyc = dy + (size_t)b * dy_bs + (size_t)(dy_cb + chunk) * H2p * W2p * 8;
  __nv_bfloat16* dxc = dx + (size_t)b * dx_bs + (size_t)(dx_cb + chunk) * Hp * Wp * 8;
  const __nv_bfloat16* ac = act ? act + (size_t)b * act_bs + (size_t)(act_cb + chunk) * Hp * Wp * 8 : nullptr;
  const bool in = v < W;
  float wx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) wx[k] = in ? up2_wgt(2 * v - 1 + k, W, v) : 0.f;
  // this thread's two hi-res columns (2v, 2v+1) of the rows 2u-1 and 2u, carried from one low-res row to the next so
  // that every hi-res row is loaded once per block (rows outside the image are the zero pad of dy: weight 0, safe)
  uint32_t ra[8], rb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ra[i] = rb[i] = 0;
  if (in) {
    ldg256(dyc + ((size_t)(2 * u0 - 1 + 2) * W2p + 2 * v + 2) * 8, ra);
    ldg256(dyc + ((size_t)(2 * u0 + 2) * W2p + 2 * v + 2) * 8, rb);
  }
  for (int u = u0; u < u1; ++u) {
    uint32_t rc[8], rd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rc[i] = rd[i] = 0;
    if (in) {
      ldg256(dyc + ((size_t)(2 * u + 1 + 2) * W2p + 2 * v + 2) * 8, rc);
      ldg256(dyc + ((size_t)(2 * u + 2 + 2) * W2p + 2 * v + 2) * 8, rd);
    }
    float wy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wy[k] = up2_wgt(2 * u - 1 + k, H, u);
    float V0[8], V1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) V0[i] = V1[i] = 0.f;
    const uint32_t* rows[4] = {ra, rb, rc, rd};
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      float a0[8], a1[8];
      unpack8(rows[ky], a0);
      unpack8(rows[ky] + 4, a1);
#pragma unroll
      for (int i = 0; i < 8; ++i) { V0[i] = fmaf(wy[ky], a0[i], V0[i]); V1[i] = fmaf(wy[ky], a1[i], V1[i]); }
    }
    up_sm[v] = make_float4(V0[0], V0[1], V0[2], V0[3]);
    up_sm[T + v] = make_float4(V0[4], V0[5], V0[6], V0[7]);
    up_sm[2 * T + v] = make_float4(V1[0], V1[1], V1[2], V1[3]);
    up_sm[3 * T + v] = make_float4(V1[4], V1[5], V1[6], V1[7]);
    __syncthreads();
    if (in) {
      float s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = wx[1] * V0[i];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = fmaf(wx[2], V1[i], s[i]);
      if (v > 0) {           // hi-res column 2v-1 = V1 of the left neighbour
        const float4 l0 = up_sm[2 * T + v - 1], l1 = up_sm[3 * T + v - 1];
        const float L[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = fmaf(wx[0], L[i], s[i]);
      }
      if (v < W - 1) {       // hi-res column 2v+2 = V0 of the right neighbour
        const float4 r0 = up_sm[v + 1], r1 = up_sm[T + v + 1];
        const float R[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = fmaf(wx[3], R[i], s[i]);
      }
      const size_t pix = ((size_t)(u + 2) * Wp + v + 2) * 8;
      if (ac) {
        float av[8];
        ld8(ac + pix, av);
#pragma unroll
        for (int i = 0; i < 8; ++i) if (!(av[i] > 0.f)) s[i] = 0.f;
      }
      if (accumulate) {
        float old[8];
        ld8(dxc + pix, old);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += old[i];
      }
      st8(dxc + pix, s);
    }
    __syncthreads();        // the exchange buffer is rewritten by the next row
#pragma unroll
    for (int i = 0; i < 8; ++i) { ra[i] = rc[i]; rb[i] = rd[i]; }
  }
}

// Per-pixel kernel (any width): one thread per low-res pixel and chunk, all 16 loads issued before the arithmetic.
__global__ void __launch_bounds__(256)
blk_upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_bs, int dy_cb, int H, int W,
                          __nv_bfloat16* __restrict__ dx, long long dx_bs, int dx_cb,
                          const __nv_bfloat16* __restrict__ act, long long act_bs, int act_cb, int accumulate) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int Wp = W + 4, Hp = H + 4, H2 = 2 * H, W2 = 2 * W, W2p = W2 + 4, H2p = H2 + 4;
  const __nv_bfloat16* dyc = dy + (size_t)b * dy_bs + (size_t)(dy_cb + chunk) * H2p * W2p * 8;
  __nv_bfloat16* dxc = dx + (size_t)b * dx_bs + (size_t)(dx_cb + chunk) * Hp * Wp * 8;
  const __nv_bfloat16* ac = act ? act + (size_t)b * act_bs + (size_t)(act_cb + chunk) * Hp * Wp * 8 : nullptr;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H * W; e += gridDim.x * 256) {
    const int u = e / W, v = e % W;
    float wx[4], wy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { wx[k] = up2_wgt(2 * v - 1 + k, W, v); wy[k] = up2_wgt(2 * u - 1 + k, H, u); }
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    uint4 raw[4][4];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
      for (int kx = 0; kx < 4; ++kx)
        raw[ky][kx] = __ldg(reinterpret_cast<const uint4*>(dyc + ((size_t)(2 * u - 1 + ky + 2) * W2p + (2 * v - 1 + kx) + 2) * 8));
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
#pragma unroll
      for (int h2 = 0; h2 < 4; ++h2) {   // 4 bf16 pairs = 8 channels
        float hx = 0.f, hy = 0.f;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const __nv_bfloat162 pr = reinterpret_cast<const __nv_bfloat162*>(&raw[ky][kx])[h2];
          const float2 f2 = __bfloat1622float2(pr);
          hx = fmaf(wx[kx], f2.x, hx); hy = fmaf(wx[kx], f2.y, hy);
        }
        s[2 * h2] = fmaf(wy[ky], hx, s[2 * h2]);
        s[2 * h2 + 1] = fmaf(wy[ky], hy, s[2 * h2 + 1]);
      }
    }
    const size_t pix = ((size_t)(u + 2) * Wp + v + 2) * 8;
    if (ac) {
      float av[8];
      ld8(ac + pix, av);
#pragma unroll
      for (int i = 0; i < 8; ++i) if (!(av[i] > 0.f)) s[i] = 0.f;
    }
    if (accumulate) {
      float old[8];
      ld8(dxc + pix, old);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += old[i];
    }
    st8(dxc + pix, s);
  }
}

// space-to-depth for the stride-2 layers: phase (py,px) plane p = py*2+px holds x[2y+py, 2x+px];
// output chunk index = p*CB + chunk at half resolution.
__global__ void __launch_bounds__(256)
blk_space_to_depth_kernel(const __nv_bfloat16* __restrict__ x, long long x_bs, int x_cb, int CB, int H, int W,
                          __nv_bfloat16* __restrict__ y, long long y_bs) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int Hp = H + 4, Wp = W + 4, H2 = H / 2, W2 = W / 2, H2p = H2 + 4, W2p = W2 + 4;
  const __nv_bfloat16* xc = x + (size_t)b * x_bs + (size_t)(x_cb + chunk) * Hp * Wp * 8;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H * W; e += gridDim.x * 256) {
    const int yy = e / W, xx = e % W;
    const int p = (yy & 1) * 2 + (xx & 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xc + ((size_t)(yy + 2) * Wp + xx + 2) * 8));
    __nv_bfloat16* yc = y + (size_t)b * y_bs + (size_t)(p * CB + chunk) * H2p * W2p * 8;
    *reinterpret_cast<uint4*>(yc + ((size_t)(yy / 2 + 2) * W2p + xx / 2 + 2) * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// host-side plan construction
// ---------------------------------------------------------------------------------------------
enum { KIND_K5S1 = 0, KIND_K1 = 1, KIND_K5S2 = 2, KIND_K5S1_DGRAD = 3, KIND_K1_DGRAD = 4, KIND_K5S2_DGRAD = 5 };

int build_plan(int kind, int n_chunks, int pitch, int py, int px, cnp_conv_plan* p) {
  memset(p, 0, sizeof(*p));
  int st = 0;
  auto add_ab = [&](int chunk0, int wci0) {
    p->ab_chunk0[p->n_ab] = chunk0; p->ab_wci0[p->n_ab] = wci0; p->ab_stage0[p->n_ab] = st; ++p->n_ab;
  };
  auto add_tap = [&](int s, int aoff, int wky, int wkx) {
    int t = p->st_ntap[s]++;
    p->st_aoff[s][t] = (short)aoff; p->st_wky[s][t] = (signed char)wky; p->st_wkx[s][t] = (signed char)wkx;
  };
  if (kind == KIND_K5S1 || kind == KIND_K5S1_DGRAD) {
    CNP_REQUIRE(n_chunks == 8 || n_chunks == 16, "conv plan: 5x5 needs 8 or 16 source chunks");
    for (int g = 0; g < n_chunks / 4; ++g) {
      add_ab(4 * g, 32 * g);
      for (int ky = 0; ky < 5; ++ky, ++st)
        for (int kx = 0; kx < 5; ++kx)
          add_tap(st, ky * pitch + kx, kind == KIND_K5S1 ? ky : 4 - ky, kind == KIND_K5S1 ? kx : 4 - kx);
    }
  } else if (kind == KIND_K1 || kind == KIND_K1_DGRAD) {
    CNP_REQUIRE(n_chunks == 8 || n_chunks == 16, "conv plan: 1x1 needs 8 or 16 source chunks");
    for (int g = 0; g < n_chunks / 4; ++g) { add_ab(4 * g, 32 * g); add_tap(st, 2 * pitch + 2, 0, 0); ++st; }
  } else if (kind == KIND_K5S2) {
    CNP_REQUIRE(n_chunks == 32, "conv plan: stride-2 forward reads the 4x8-chunk phase tensor");
    for (int ph = 0; ph < 4; ++ph) {
      const int qy = ph >> 1, qx = ph & 1;
      for (int half = 0; half < 2; ++half) {
        add_ab(ph * 8 + 4 * half, 32 * half);
        for (int ky = qy; ky < 5; ky += 2, ++st)
          for (int kx = qx; kx < 5; kx += 2)
            add_tap(st, (2 + (ky - 2 - qy) / 2) * pitch + 2 + (kx - 2 - qx) / 2, ky, kx);
      }
    }
  } else if (kind == KIND_K5S2_DGRAD) {
    CNP_REQUIRE(n_chunks == 8, "conv plan: stride-2 dgrad reads the 8-chunk dy tensor");
    for (int half = 0; half < 2; ++half) {
      add_ab(4 * half, 32 * half);
      for (int ky = py; ky < 5; ky += 2, ++st)
        for (int kx = px; kx < 5; kx += 2)
          add_tap(st, (2 + (py + 2 - ky) / 2) * pitch + 2 + (px + 2 - kx) / 2, ky, kx);
    }
  } else {
    CNP_REQUIRE(false, "conv plan: unknown kind %d", kind);
  }
  p->ab_stage0[p->n_ab] = st;
  CNP_REQUIRE(st <= CNP_MAX_ST && p->n_ab <= CNP_MAX_AB, "conv plan: too many stages");
  return 0;
}

// pick the tile (TW, TH) maximising useful MMA rows; (TW+4)*TH <= 512 window pixels (R <= 4)
void choose_tile(int H, int W, int* TW, int* TH, int* R) {
  double best = -1.0;
  for (int tw = 4; tw <= 252 && tw <= W + 3; ++tw) {
    const int pitch = tw + 4;
    for (int th = 1; th <= 32; ++th) {
      const int win = pitch * th;
      if (win > 512) break;
      const int r = (win + 127) / 128;
      const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
      const double eff = (double)H * W / ((double)tiles * 128.0 * r);
      // mild preference for fewer, larger tiles (less halo re-read, fewer barriers)
      const double score = eff + 1e-4 * r;
      if (score > best) { best = score; *TW = tw; *TH = th; *R = r; }
    }
  }
}

int fill_geometry(cnp_conv_args* a) {
  int TW, TH, R;
  choose_tile(a->H, a->W, &TW, &TH, &R);
  a->TW = TW; a->TH = TH; a->R = R; a->pitch = TW + 4;
  a->tiles_x = cnp_cdiv(a->W, TW); a->tiles_y = cnp_cdiv(a->H, TH);
  a->plane_sm = 128 * R + 4 * a->pitch + 8;
  return 0;
}

size_t conv_smem_bytes(const cnp_conv_args& a) {
  return (size_t)A_BUFS * 4 * a.plane_sm * 16 + (size_t)W_STAGES * W_STAGE_BYTES + 16 * 8 + 16;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------

struct cnp_conv_out {
  int mode;               // 0: blocked bf16 (blk), 1: NCHW fp32 (f32 / f32_bstride / f32_ch_off / H x W)
  cnp_blk blk;
  float* f32; long long f32_bstride; int f32_ch_off;
  int sy, ay, sx, ax;     // output pixel = (y*sy+ay, x*sx+ax); (1,0,1,0) for a plain conv
  const float* bias;      // [64] or NULL
  int relu;
  const cnp_blk* mask;    // zero the result where mask <= 0 (same geometry as the output) or NULL
  int accumulate;
  const cnp_blk* s2d;     // (conv_tc2 only) optional space-to-depth copy of the output
};

#ifdef CNP_LEGACY_CONV_TC
// Bytes of packed weights for (kind, n_chunks): stages x 20480.
CNP_API long long cnp_conv_tc_packed_bytes(int kind, int n_chunks) {
  cnp_conv_plan p;
  if (build_plan(kind, n_chunks, 8, 0, 0, &p)) return -1;
  return (long long)p.ab_stage0[p.n_ab] * W_STAGE_BYTES;
}

// Pack torch-layout fp32 weights [Cout][Cin][k][k] for conv_tc (see pack_weights_kernel).
CNP_API int cnp_conv_tc_pack(const float* w, int Cout, int Cin, int k, int kind, int n_chunks, int py, int px,
                             int co_off, void* wpk, cudaStream_t st) {
  cnp_conv_plan p;
  if (int e = build_plan(kind, n_chunks, 8, py, px, &p)) return e;
  const int transposed = (kind >= KIND_K5S1_DGRAD) ? 1 : 0;
  pack_weights_kernel<<<64, 256, 0, st>>>(w, Cout, Cin, k, transposed, co_off, reinterpret_cast<__nv_bfloat16*>(wpk), p);
  CNP_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}

// Tensor-core convolution producing 64 output channels.  x: blocked source whose chunks
// [x->cb_off, x->cb_off + n_chunks) are the K dimension; (x->H, x->W) is the accumulator grid.
CNP_API int cnp_conv_tc(const cnp_blk* x, int n_chunks, const void* wpk, int kind, int py, int px,
                        const cnp_conv_out* o, int B, cudaStream_t st) {
  CNP_REQUIRE(x && o && wpk && B > 0, "conv_tc: bad arguments");
  cnp_conv_args a;
  memset(&a, 0, sizeof(a));
  a.H = x->H; a.W = x->W; a.B = B;
  a.x_Hp = x->H + 4; a.x_Wp = x->W + 4;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x->base) + (long long)x->cb_off * a.x_Hp * a.x_Wp * 8;
  a.x_bs = x->bstride;
  a.w = reinterpret_cast<const __nv_bfloat16*>(wpk);
  fill_geometry(&a);
  if (int e = build_plan(kind, n_chunks, a.pitch, py, px, &a.plan)) return e;
  a.out_mode = o->mode;
  a.sy = o->sy; a.ay = o->ay; a.sx = o->sx; a.ax = o->ax;
  CNP_REQUIRE(a.sy >= 1 && a.sx >= 1, "conv_tc: output scale must be >= 1");
  if (o->mode == 0) {
    CNP_REQUIRE(o->blk.H == x->H * o->sy && o->blk.W == x->W * o->sx, "conv_tc: output geometry mismatch");
    a.out = o->blk.base; a.out_bs = o->blk.bstride; a.out_c_off = o->blk.cb_off;
    a.out_Hp = o->blk.H + 4; a.out_Wp = o->blk.W + 4;
    if (o->mask) {
      CNP_REQUIRE(o->mask->H == o->blk.H && o->mask->W == o->blk.W, "conv_tc: mask geometry mismatch");
      a.mask = reinterpret_cast<const __nv_bfloat16*>(o->mask->base);
      a.mask_bs = o->mask->bstride; a.mask_cb_off = o->mask->cb_off;
    }
  } else {
    CNP_REQUIRE(o->f32 && o->sy == 1 && o->sx == 1 && !o->mask, "conv_tc: fp32 NCHW output is plain only");
    a.out = o->f32; a.out_bs = o->f32_bstride; a.out_c_off = o->f32_ch_off;
    a.out_Hp = x->H; a.out_Wp = x->W;
  }
  a.bias = o->bias; a.relu = o->relu; a.accumulate = o->accumulate;
  const size_t smem = conv_smem_bytes(a);
  CNP_REQUIRE(smem <= 227 * 1024, "conv_tc: tile needs %zu B of shared memory", smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cnp_set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr = smem;
  }
  const int ntiles = B * a.tiles_x * a.tiles_y;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  conv_tc_kernel<<<grid, 256, smem, st>>>(a);
  CNP_LAUNCH_CHECK("conv_tc_kernel");
  return 0;
}

#endif  // CNP_LEGACY_CONV_TC

CNP_API int cnp_blk_from_nchw_f32(const float* src, long long src_bstride, int B, int C, int H, int W,
                                  const cnp_blk* dst, cudaStream_t st) {
  CNP_REQUIRE(dst && dst->H == H && dst->W == W, "blk_from_nchw: geometry mismatch");
  dim3 grid(min(cnp_cdiv(H * W, 256), 128), cnp_cdiv(C, 8), B);
  nchw_to_blk_kernel<<<grid, 256, 0, st>>>(src, src_bstride, C, H, W, reinterpret_cast<__nv_bfloat16*>(dst->base),
                                           dst->bstride, dst->cb_off, -1, 0ull);
  CNP_LAUNCH_CHECK("nchw_to_blk_kernel");
  return 0;
}

// Same conversion into n_chunks chunks, with channel C set to 1 inside the image (the pad stays 0) and channels
// C+1.. zero: the input of the first UNet convolution once the initial 1x1 is folded into it (fold_in.cu) -- the
// constant channel carries the 1x1's bias through the zero-padded 5x5 window exactly.
// shared_mask: bit c set = channel c of `src` is only valid in batch 0 (a context set shared by all tasks, encoded once)
// and is broadcast to every task here instead of by a separate copy.
CNP_API int cnp_blk_from_nchw_f32_ones(const float* src, long long src_bstride, int B, int C, int H, int W,
                                       const cnp_blk* dst, int n_chunks, unsigned long long shared_mask, cudaStream_t st) {
  CNP_REQUIRE(dst && dst->H == H && dst->W == W, "blk_from_nchw_ones: geometry mismatch");
  CNP_REQUIRE(C + 1 <= n_chunks * 8 && C < 64, "blk_from_nchw_ones: need C + 1 <= 8 * n_chunks and C < 64");
  dim3 grid(min(cnp_cdiv(H * W, 256), 128), n_chunks, B);
  nchw_to_blk_kernel<<<grid, 256, 0, st>>>(src, src_bstride, C, H, W, reinterpret_cast<__nv_bfloat16*>(dst->base),
                                           dst->bstride, dst->cb_off, C, shared_mask);
  CNP_LAUNCH_CHECK("nchw_to_blk_kernel(ones)");
  return 0;
}

CNP_API int cnp_blk_to_nchw_f32(const cnp_blk* src, int B, int C, float* dst, long long dst_bstride, cudaStream_t st) {
  CNP_REQUIRE(src, "blk_to_nchw: null source");
  dim3 grid(min(cnp_cdiv(src->H * src->W, 256), 128), cnp_cdiv(C, 8), B);
  blk_to_nchw_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src->base), src->bstride, src->cb_off,
                                           C, src->H, src->W, dst, dst_bstride);
  CNP_LAUNCH_CHECK("blk_to_nchw_kernel");
  return 0;
}

CNP_API int cnp_conv1x1_in_bf16(const float* x, long long x_bstride, const float* w, const float* bias, int B, int Cin,
                                int Cout, const cnp_blk* out, cudaStream_t st) {
  CNP_REQUIRE(out && Cin >= 1 && Cin <= IC_MAX_CIN && Cout >= 1 && Cout <= 64, "conv1x1_in: need Cin<=%d, Cout<=64", IC_MAX_CIN);
  dim3 grid(cnp_cdiv(out->H * out->W, 256), B);
  conv1x1_in_kernel<<<grid, 256, 0, st>>>(x, x_bstride, Cin, out->H, out->W, w, bias, Cout,
                                          reinterpret_cast<__nv_bfloat16*>(out->base), out->bstride, out->cb_off);
  CNP_LAUNCH_CHECK("conv1x1_in_kernel");
  return 0;
}

CNP_API int cnp_blk_upsample2x_fwd(const cnp_blk* x, int n_chunks, const cnp_blk* y, int B, cudaStream_t st) {
  CNP_REQUIRE(x && y && y->H == 2 * x->H && y->W == 2 * x->W, "blk_upsample2x_fwd: geometry mismatch");
  CNP_REQUIRE((reinterpret_cast<uintptr_t>(y->base) & 31) == 0 && (y->bstride & 15) == 0,
              "blk_upsample2x_fwd: output must be 32-byte aligned (256-bit stores)");
  dim3 grid(cnp_cdiv(x->H * x->W, 256), n_chunks, B);
  blk_upsample2x_fwd_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x->base), x->bstride, x->cb_off,
                                                  x->H, x->W, reinterpret_cast<__nv_bfloat16*>(y->base), y->bstride,
                                                  y->cb_off);
  CNP_LAUNCH_CHECK("blk_upsample2x_fwd_kernel");
  return 0;
}

CNP_API int cnp_blk_upsample2x_bwd(const cnp_blk* dy, int n_chunks, const cnp_blk* dx, const cnp_blk* act, int accumulate,
                                   int B, cudaStream_t st) {
  CNP_REQUIRE(dy && dx && dy->H == 2 * dx->H && dy->W == 2 * dx->W, "blk_upsample2x_bwd: geometry mismatch");
  CNP_REQUIRE(!act || (act->H == dx->H && act->W == dx->W), "blk_upsample2x_bwd: mask geometry mismatch");
  const __nv_bfloat16* ap = act ? reinterpret_cast<const __nv_bfloat16*>(act->base) : nullptr;
  if (dx->W <= 512 && (reinterpret_cast<uintptr_t>(dy->base) & 31) == 0 && (dy->bstride & 15) == 0) {
    const int T = cnp_cdiv(dx->W, 32) * 32;
    const int rpb = dx->H >= 100 ? 8 : (dx->H >= 50 ? 4 : 1);     // small images need the blocks more than the reuse
    blk_upsample2x_bwd_row_kernel<<<dim3(cnp_cdiv(dx->H, rpb), n_chunks, B), T, (size_t)4 * T * sizeof(float4), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy->base), dy->bstride, dy->cb_off, dx->H, dx->W,
        reinterpret_cast<__nv_bfloat16*>(dx->base), dx->bstride, dx->cb_off, ap, act ? act->bstride : 0,
        act ? act->cb_off : 0, accumulate, rpb);
    CNP_LAUNCH_CHECK("blk_upsample2x_bwd_row_kernel");
    return 0;
  }
  dim3 grid(cnp_cdiv(dx->H * dx->W, 256), n_chunks, B);
  blk_upsample2x_bwd_kernel<<<grid, 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy->base), dy->bstride, dy->cb_off, dx->H, dx->W,
      reinterpret_cast<__nv_bfloat16*>(dx->base), dx->bstride, dx->cb_off, ap, act ? act->bstride : 0,
      act ? act->cb_off : 0, accumulate);
  CNP_LAUNCH_CHECK("blk_upsample2x_bwd_kernel");
  return 0;
}

CNP_API int cnp_blk_space_to_depth(const cnp_blk* x, int n_chunks, const cnp_blk* y, int B, cudaStream_t st) {
  CNP_REQUIRE(x && y && x->H % 2 == 0 && x->W % 2 == 0 && y->H == x->H / 2 && y->W == x->W / 2 && y->cb_off == 0,
              "blk_space_to_depth: geometry mismatch");
  dim3 grid(min(cnp_cdiv(x->H * x->W, 256), 128), n_chunks, B);
  blk_space_to_depth_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x->base), x->bstride, x->cb_off,
                                                  n_chunks, x->H, x->W, reinterpret_cast<__nv_bfloat16*>(y->base),
                                                  y->bstride);
  CNP_LAUNCH_CHECK("blk_space_to_depth_kernel");
  return 0;
}

CNP_API int cnp_conv1x1_in_wgrad(const float* x, long long x_bstride, int Cin, const cnp_blk* dy, int B, float* dw,
                                 float* dbias, cudaStream_t st) {
  CNP_REQUIRE(x && dy && dw && B > 0 && Cin >= 1 && Cin < IW_MAXCI, "conv1x1_in_wgrad: need Cin < %d", IW_MAXCI);
  const int nh = Cin < 16 ? 1 : 2;
  const size_t smem = 2 * ((size_t)IW_PX * 64 * 2 + (size_t)IW_PX * 16 * nh * 4);   // two stages
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv1x1_in_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (IW_PX * 64 * 2 + IW_PX * 16 * 4));
    cudaFuncSetAttribute(conv1x1_in_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (IW_PX * 64 * 2 + IW_PX * 32 * 4));
    attr = true;
  }
  const int tiles = B * cnp_cdiv(dy->H * dy->W, IW_PX);
  const int grid = tiles < 3 * num_sms() ? tiles : 3 * num_sms();
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy->base);
  if (nh == 1)
    conv1x1_in_wgrad_kernel<1><<<grid, 256, smem, st>>>(x, x_bstride, Cin, dy->H, dy->W, dyp, dy->bstride, dy->cb_off, B, dw, dbias);
  else
    conv1x1_in_wgrad_kernel<2><<<grid, 256, smem, st>>>(x, x_bstride, Cin, dy->H, dy->W, dyp, dy->bstride, dy->cb_off, B, dw, dbias);
  CNP_LAUNCH_CHECK("conv1x1_in_wgrad_kernel");
  return 0;
}
