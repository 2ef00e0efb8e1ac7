// bf16 weight gradient of the UNet convolutions on tcgen05 (sm_100a).
//
//   dW[co][ci][ky][kx] += sum_{b,y,x} dY[b,co,y,x] * X[b,ci, s*y+ky-2, s*x+kx-2]
//
// (torch autograd of Conv2d inside upstream neuralprocesses' UNet, reached from
// mean_batch_loss.backward() in deepsensor.train.train_epoch -- nzdownscale/downscaler/train.py:388-394.)
//
// GEMM view: D[ci, co] = sum_p Xs[p, ci] * dY[p, co] with the reduction over pixels p.  Both blocked
// tensors ([chunk][pixel][8 ch]) are *MN-major* UMMA operands as they lie in memory: 8 channels are
// contiguous (16 B) and consecutive pixels are 16 B apart, so a run of pixels copied with one
// cp.async.bulk per chunk plane is directly a SWIZZLE_NONE MN-major tile (SBO = plane stride,
// LBO = 128 B).  The reduction runs over the padded image in *linear* pixel order: the zero pad of dY
// annihilates every out-of-image product, so no 2-D tiling or boundary logic is needed, and a tap
// (ky,kx) is a constant linear shift of X.  One CTA owns one kernel row ky (a "pass") and keeps its
// taps as separate M=128 x N=64 fp32 accumulators in TMEM for the whole K loop:
//   * 128-channel inputs: M = the 16 chunks, one accumulator per kx (5 x 64 TMEM columns);
//   * 64-channel inputs : M = [8 chunks ; the same 8 chunks shifted by 3 pixels], so one MMA yields
//     taps kx and kx+3 (3 accumulators);
//   * stride-2 layers   : M = [phase (py,0) ; phase (py,1)] of the space-to-depth tensor.
// K is split across CTAs; partial sums are reduced with fp32 atomics straight into the torch-layout
// gradient (or through a workspace + reduce kernel for long K loops).
//
// N = 128 ("dup") variant, used for the 5x5 kinds: the M128 x N64 MMA is shared-memory bound (48 cycles against a
// 32-cycle floor, profiles/r01_mma_rate.txt) while N = 128 runs at its floor (64 cycles).  The dY tile is therefore
// loaded twice, the second copy shifted by one pixel, and used as ONE N = 128 operand [dY(p) ; dY(p-1)]: an MMA with
// the X operand at offset a then yields tap a (columns 0..63) and tap a+1 (columns 64..127).  A 128-channel layer
// needs 3 MMAs of 64 cycles per 16-pixel K step (taps {0,1},{2,3},{4,5}; 5 is discarded) instead of 5 of 48.
#include "tc_common.cuh"
#include <stdlib.h>

#define CNP_WG_MAX_PASS 10
#define CNP_WG_MAX_ACC 5


struct cnp_wg_pass {
  int chunk0;        // first source chunk of the M operand
  int shift_px;      // >=0: planes 8..15 = chunks chunk0..chunk0+7 shifted by shift_px pixels; -1: real chunks 8..15
  int base_off;      // linear pixel offset of the X tile relative to the dY tile
  int n_acc;
  int a0, astep;               // accumulator j uses the X operand at pixel offset a0 + j*astep
  int slot[CNP_WG_MAX_ACC][2][2];   // tap index (ky*k+kx) of [accumulator][row half][column half], -1 = discard
  int ci0, ci1;                // input-channel base of the two row halves
  int ky0, n_grp;              // narrow inputs: the pass covers kernel rows ky0 .. ky0 + n_grp - 1 (one row group each)
  int dy_chunk0;               // first dY chunk of this pass (phase kind: 16 * row phase)
  int img0, n_img;             // images this pass reduces over (n_img = 0: all B)
};

struct cnp_wg_args {
  const __nv_bfloat16* x; long long x_bs; long long x_plane;   // elements
  const __nv_bfloat16* dy; long long dy_bs; long long dy_plane;
  float* dw; float* dbias; int Cin, KK;
  float* dw2; int dw2_from_pass; // passes >= dw2_from_pass accumulate into dw2 instead of dw (row / column strips in one launch)
  float* ws;                     // optional partial-sum workspace [pass][ksplit][acc][128][64]; NULL = atomics into dw
  float* ws_bias;                // with ws: per-CTA bias partials [pass][ksplit][64]
  int B, P, p_start, tiles_per_img, ksplit, n_pass;
  int dup;                       // 1: N = 128 operand [dY(p) ; dY(p-1)] (column half 1 = tap a+1)
  int ws_acc;                    // accumulators per CTA in the workspace layout
  int cluster;                   // 5: the five ky passes of a K-split slice form a cluster and share the dY stream
  int narrow;                    // >0: chunks per row group -- the 16 X planes are [kernel row][chunk] (few input channels)
  int grp_shift;                 // narrow: pixel shift between consecutive row groups (= padded row pitch)
  int dy_pair;                   // 1: the N = 128 operand is [dY chunks c0..c0+7 (p) ; chunks c0+8..c0+15 (p-1)] -- two x-phases
  int bias_grp;                  // passes that see the same dY tiles (bias sums are dealt round-robin inside a group)
  cnp_wg_pass pass[CNP_WG_MAX_PASS];
};

namespace {

constexpr int WG_STAGES = 3;
constexpr int XPAD = 8;  // extra pixels per X plane in shared memory (max tap offset 4, keeps 128 B alignment)

// K steps [K0, K1) x NACC accumulators of one stage, fully unrolled (accumulator j: X operand at +a_j pixels)
template <int NACC, int K0, int K1>
__device__ __forceinline__ void wg_issue(uint32_t tmem, uint32_t ncols, uint32_t xs16, uint32_t astep, uint32_t ds16,
                                         uint32_t lbo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int ks = K0; ks < K1; ++ks) {
#pragma unroll
    for (int j = 0; j < NACC; ++j)
      tc::mma_bf16_ss_lohi(tmem + j * ncols, ((xs16 + ks * 16 + j * astep) & 0x3FFFu) | lbo, a_hi,
                           ((ds16 + ks * 16) & 0x3FFFu) | lbo, b_hi, idesc, ks > K0 ? 1u : acc0);
  }
}
__device__ __forceinline__ void wg_issue_n(int nacc, int k0, int k1, uint32_t tmem, uint32_t ncols, uint32_t xs16,
                                           uint32_t astep, uint32_t ds16, uint32_t lbo, uint32_t a_hi, uint32_t b_hi,
                                           uint32_t idesc, uint32_t acc0) {
  for (int ks = k0; ks < k1; ++ks)
    for (int j = 0; j < nacc; ++j)
      tc::mma_bf16_ss_lohi(tmem + j * ncols, ((xs16 + ks * 16 + j * astep) & 0x3FFFu) | lbo, a_hi,
                           ((ds16 + ks * 16) & 0x3FFFu) | lbo, b_hi, idesc, ks > k0 ? 1u : acc0);
}

// accumulator j, TMEM lane m, column half -> (tap index or -1, input channel)
__device__ __forceinline__ int wg_slot(const cnp_wg_args& a, const cnp_wg_pass& ps, int m, int j, int colhalf, int* ci) {
  if (a.narrow) {
    const int rows = a.narrow * 8, g = m / rows, kx = ps.a0 + j * ps.astep + colhalf;
    *ci = m - g * rows;
    return (g < ps.n_grp && kx < 5) ? (ps.ky0 + g) * 5 + kx : -1;
  }
  const int half = m >> 6;
  *ci = (half ? ps.ci1 : ps.ci0) + (m & 63);
  return ps.slot[j][half][colhalf];
}

// CL = true: instantiation with the cluster / multicast code (a kernel that contains cluster instructions is scheduled
// differently even when launched without a cluster, which costs ~15 % here, hence two instantiations)
template <bool CL>
__global__ void __launch_bounds__(224, 1)
wgrad_tc_kernel(const __grid_constant__ cnp_wg_args a) {
  const int a_cluster = CL ? a.cluster : 1;
  extern __shared__ __align__(128) uint8_t smem[];
  const int P = a.P;
  const int x_plane_b = (P + XPAD) * 16, dy_plane_b = P * 16;
  const int ndy = a.dup ? 16 : 8;              // dY planes per stage (dup: second copy shifted by one pixel)
  const int x_tile_b = 16 * x_plane_b, dy_tile_b = ndy * dy_plane_b;
  const int stage_b = x_tile_b + dy_tile_b;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * stage_b);
  uint64_t* full = bars;                 // [WG_STAGES]
  uint64_t* empty = bars + WG_STAGES;    // [WG_STAGES]
  uint64_t* done = bars + 2 * WG_STAGES; // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const cnp_wg_pass& ps = a.pass[blockIdx.y];
  const int total_tiles = (ps.n_img ? ps.n_img : a.B) * a.tiles_per_img;

  if (threadIdx.x == 0) {
    const bool bias_cta0 = (a.dbias != nullptr);
    // a stage is free when every reader is done with it: the MMAs (of all CTAs of the cluster, whose dY copies this
    // CTA's producer overwrites by multicast) and, with a bias gradient, the 4 summing warps (of every CTA)
    const uint32_t ncta = (CL && a_cluster > 1) ? (uint32_t)a_cluster : 1u;
    for (int i = 0; i < WG_STAGES; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, ncta * (bias_cta0 ? 5u : 1u)); }
    tc::mbar_init(done, 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  if ((CL && a_cluster > 1)) tc::cluster_sync();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = (CL && a_cluster > 1) ? tc::cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << a_cluster) - 1u);

  if (warp == 0) {
    if (tc::elect_one()) {
      // per-plane source offsets are fixed for the whole CTA: computed once, so that the per-stage work of this single
      // producer thread is 32 straight-line copy instructions (with the offsets recomputed per copy, or a loop with a
      // run-time bound, the producer becomes the bottleneck: 830 us instead of 446 us on the 128-channel layer)
      const int nxp = a.narrow ? ps.n_grp * a.narrow : 16;   // X planes really loaded (the rest is never read back)
      long long xoff[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (a.narrow) xoff[c] = (long long)(c % a.narrow) * a.x_plane + (long long)(c / a.narrow) * a.grp_shift * 8;
        else if (c < 8 || ps.shift_px < 0) xoff[c] = (long long)(ps.chunk0 + c) * a.x_plane;
        else xoff[c] = (long long)(ps.chunk0 + c - 8) * a.x_plane + (long long)ps.shift_px * 8;
      }
      const uint32_t tx = (uint32_t)nxp * (uint32_t)(P + XPAD) * 16u + (uint32_t)ndy * (uint32_t)P * 16u;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += a.ksplit, ++it) {
        const int s = it % WG_STAGES;
        tc::mbar_wait(empty + s, ((it / WG_STAGES) & 1) ^ 1);
        const int b = ps.img0 + t / a.tiles_per_img, ti = t % a.tiles_per_img;
        const long long p0 = a.p_start + (long long)ti * P;
        uint8_t* xs = smem + s * stage_b;
        tc::mbar_expect_tx(full + s, tx);       // the whole stage (X planes from here, dY planes from warp 6)
        const __nv_bfloat16* xb = a.x + (long long)b * a.x_bs + (p0 + ps.base_off) * 8;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < nxp) tc::bulk_g2s(xs + c * x_plane_b, xb + xoff[c], (uint32_t)(P + XPAD) * 16u, full + s);
      }
    }
  } else if (warp == 6) {
    // second producer thread: the dY planes of every stage (one thread issuing all ~32 copies of a stage limits the
    // kernel; the barrier's transaction count is armed by warp 0 -- completions may arrive first, the phase cannot
    // complete before warp 0's arrive.expect_tx)
    if (tc::elect_one()) {
      long long doff[16];
#pragma unroll
      for (int c = 0; c < 16; ++c)   // 8..15 (dup): one pixel earlier; phase pairs: the second x-phase's chunks
        doff[c] = (long long)(ps.dy_chunk0 + (a.dy_pair ? c : (c & 7))) * a.dy_plane - (c >> 3) * 8;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += a.ksplit, ++it) {
        const int s = it % WG_STAGES;
        tc::mbar_wait(empty + s, ((it / WG_STAGES) & 1) ^ 1);
        const int b = ps.img0 + t / a.tiles_per_img, ti = t % a.tiles_per_img;
        const long long p0 = a.p_start + (long long)ti * P;
        uint8_t* ds = smem + s * stage_b + x_tile_b;
        const __nv_bfloat16* db = a.dy + (long long)b * a.dy_bs + p0 * 8;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < ndy) {
            if ((CL && a_cluster > 1)) {            // every CTA of the cluster needs the same dY tile: load a fifth, multicast it
              if ((uint32_t)c % (uint32_t)a_cluster == crank)
                tc::bulk_g2s_mc(ds + c * dy_plane_b, db + doff[c], (uint32_t)P * 16u, full + s, cmask);
            } else {
              tc::bulk_g2s(ds + c * dy_plane_b, db + doff[c], (uint32_t)P * 16u, full + s);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one() && (int)blockIdx.x < total_tiles) {
      // MN-major operands: LBO = 128 B between the two 8-pixel K groups, SBO = plane stride between 8-channel groups.
      // Only the low descriptor word changes between MMAs (+1 per tap pixel, +16 per 16-pixel K step), and the wait
      // for the next stage is issued before the last K step so its latency overlaps queued MMAs.
      const uint32_t idesc = a.dup ? tc::make_idesc_bf16(128, 128, 1, 1) : tc::make_idesc_bf16(128, 64, 1, 1);
      const uint32_t ncols = a.dup ? 128u : 64u;
      const uint32_t a_hi = ((uint32_t)x_plane_b >> 4) | (1u << 14);
      const uint32_t b_hi = ((uint32_t)dy_plane_b >> 4) | (1u << 14);
      const uint32_t lbo = (128u >> 4) << 16;
      const int nks = P / 16;
      const uint32_t astep = (uint32_t)ps.astep;
      uint32_t it = 0;
      tc::mbar_wait(full, 0);
      tc::fence_after_sync();
      for (int t = blockIdx.x; t < total_tiles; t += a.ksplit, ++it) {
        const int s = it % WG_STAGES;
        const uint32_t xs16 = (tc::smem_u32(smem + s * stage_b) >> 4) + (uint32_t)ps.a0;
        const uint32_t ds16 = tc::smem_u32(smem + s * stage_b + x_tile_b) >> 4;
        const bool has_next = t + a.ksplit < total_tiles;
        const uint32_t acc0 = it > 0 ? 1u : 0u;
        const bool fast = nks == 8;
        if (fast && ps.n_acc == 3) wg_issue<3, 0, 7>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc0);
        else if (fast && ps.n_acc == 4) wg_issue<4, 0, 7>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc0);
        else if (fast && ps.n_acc == 2) wg_issue<2, 0, 7>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc0);
        else if (fast && ps.n_acc == 5) wg_issue<5, 0, 7>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc0);
        else wg_issue_n(ps.n_acc, 0, nks - 1, tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc0);
        if (has_next) {
          const uint32_t n = it + 1;
          tc::mbar_wait(full + n % WG_STAGES, (n / WG_STAGES) & 1);
          tc::fence_after_sync();
        }
        const uint32_t acc1 = (it > 0 || nks > 1) ? 1u : 0u;
        if (fast && ps.n_acc == 3) wg_issue<3, 7, 8>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc1);
        else if (fast && ps.n_acc == 4) wg_issue<4, 7, 8>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc1);
        else if (fast && ps.n_acc == 2) wg_issue<2, 7, 8>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc1);
        else if (fast && ps.n_acc == 5) wg_issue<5, 7, 8>(tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc1);
        else wg_issue_n(ps.n_acc, nks - 1, nks, tmem_base, ncols, xs16, astep, ds16, lbo, a_hi, b_hi, idesc, acc1);
        if ((CL && a_cluster > 1)) tc::mma_commit_mc(empty + s, cmask); else tc::mma_commit(empty + s);
      }
      tc::mma_commit(done);
    }
  } else if (warp < 6) {
    // epilogue warps 2..5 -> TMEM lane quadrant (warp & 3)
    const int q = warp & 3;
    const bool has_work = blockIdx.x < total_tiles;
    if (a.dbias != nullptr) {
      // bias gradient: these warps idle in the K loop, so they sum the dY tiles streamed anyway; every tile
      // is seen by all passes (blockIdx.y), tile #it is summed by pass it % n_pass to spread the LDS traffic
      const int et = (warp - 2) * 32 + lane;  // 0..127
      float bs[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) bs[i] = 0.f;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += a.ksplit, ++it) {
        const int s = it % WG_STAGES;
        tc::mbar_wait(full + s, (it / WG_STAGES) & 1);
        const uint8_t* ds = smem + s * stage_b + x_tile_b;
        if ((int)(it % (uint32_t)a.bias_grp) == (int)blockIdx.y % a.bias_grp)
        for (int px = et; px < P; px += 128) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 pk = *reinterpret_cast<const uint4*>(ds + c * dy_plane_b + px * 16);
            const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) bs[c * 8 + i] += __bfloat162float(pb[i]);
          }
          if (a.dy_pair) {     // the second x-phase (its copy lies one pixel earlier: the sum over all tiles is the same)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 pk = *reinterpret_cast<const uint4*>(ds + (8 + c) * dy_plane_b + px * 16);
              const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&pk);
#pragma unroll
              for (int i = 0; i < 8; ++i) bs[c * 8 + i] += __bfloat162float(pb[i]);
            }
          }
        }
        __syncwarp();
        if ((CL && a_cluster > 1)) { if (lane < a_cluster) tc::mbar_arrive_cluster(empty + s, (uint32_t)lane); }
        else if (lane == 0) tc::mbar_arrive(empty + s);
      }
      // fold the 4 epilogue warps in shared memory: ONE atomic per channel per CTA (64 hot addresses shared by
      // every CTA of the launch -- per-warp atomics cost ~35 us of serialised L2 traffic per launch)
      float* bred = reinterpret_cast<float*>(smem + WG_STAGES * stage_b + (2 * WG_STAGES + 1) * 8 + 16);   // [4][64]
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float v = bs[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) bred[(warp - 2) * 64 + i] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && has_work) {
        // with a workspace every CTA stores its partial and wgrad_reduce_kernel adds them in a fixed order
        // (run-to-run identical gradients); without one, one atomic per channel per CTA
        float* bws = a.ws ? a.ws_bias + ((size_t)blockIdx.y * a.ksplit + blockIdx.x) * 64 : nullptr;
        for (int i = lane; i < 64; i += 32) {
          const float v = bred[i] + bred[64 + i] + bred[128 + i] + bred[192 + i];
          if (bws) bws[i] = v; else atomicAdd(a.dbias + i, v);
        }
      }
    }
    if (has_work) {
      tc::mbar_wait(done, 0);
      tc::fence_after_sync();
      const int m = q * 32 + lane;
      const int ncols = a.dup ? 128 : 64;
      for (int j = 0; j < ps.n_acc; ++j) {
        float* wsp = a.ws ? a.ws + ((((size_t)blockIdx.y * a.ksplit + blockIdx.x) * a.ws_acc + j) * 128 + m) * 128 : nullptr;
        for (int hc = 0; hc < ncols / 32; ++hc) {
          int ci;
          const int slot = wg_slot(a, ps, m, j, hc >> 1, &ci);
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + j * ncols + hc * 32, v);
          tc::tmem_ld_wait();
          if (wsp) {
            // partial sums go to the workspace with plain 16 B stores; wgrad_reduce_kernel folds the K split
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(wsp + hc * 32)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else if (slot >= 0 && ci < a.Cin) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int co = (hc & 1) * 32 + i;
              atomicAdd(((a.dw2 && (int)blockIdx.y >= a.dw2_from_pass) ? a.dw2 : a.dw) + ((size_t)co * a.Cin + ci) * a.KK + slot, v[i]);
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if ((CL && a_cluster > 1)) tc::cluster_sync();   // no CTA exits while a partner may still signal its barriers
  if (warp == 1) { tc::fence_after_sync(); tc::tmem_dealloc(tmem_base, 512); }
}

// dw[co][ci][slot] += sum over the K-split CTAs of the partial accumulators written by wgrad_tc_kernel, in a FIXED order:
// gradients are run-to-run identical, so a CUDA-graph replay equals the eager step bit for bit.
// Block = 32 column quads (one 16 B load per partial per thread, a warp reads 512 contiguous bytes) x 8 K-slices (one
// warp each): slice s adds partials s, s+8, s+16, ... in order, the slices meet in shared memory and warp 0 adds them
// 0..7 in order.  Every (co, ci, tap) is owned by exactly one thread (wg_slot is one-to-one on the valid slots), so the
// final update is a plain read-modify-write.  (One thread walking all ~29 partials serially was latency-bound: 24 us
// per launch against 9 us for the atomics version; this form keeps 8x the loads in flight.)
// Block 0 also folds the per-CTA bias partials.
constexpr int WG_RED_KS = 8;
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const __grid_constant__ cnp_wg_args a) {
  __shared__ float4 part[WG_RED_KS][32];
  const int ncols = a.dup ? 128 : 64;
  const int total4 = a.n_pass * a.ws_acc * 128 * (ncols / 4);
  const int q = threadIdx.x & 31, sl = threadIdx.x >> 5;
  if (a.dbias != nullptr && blockIdx.x == 0 && threadIdx.x < 64) {
    float s = 0.f;
    const int n = a.n_pass * a.ksplit;
    for (int k = 0; k < n; ++k) s += a.ws_bias[(size_t)k * 64 + threadIdx.x];
    a.dbias[threadIdx.x] += s;
  }
  const int e = blockIdx.x * 32 + q;
  bool valid = e < total4;
  int col = 0, m = 0, j = 0, pass = 0, slot = -1, ci = 0;
  if (valid) {
    col = (e % (ncols / 4)) * 4; m = (e / (ncols / 4)) & 127;
    const int jj = e / ((ncols / 4) * 128);
    j = jj % a.ws_acc; pass = jj / a.ws_acc;
    const cnp_wg_pass& ps = a.pass[pass];
    valid = j < ps.n_acc;
    if (valid) { slot = wg_slot(a, ps, m, j, col >> 6, &ci); valid = slot >= 0 && ci < a.Cin; }
  }
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    const float* src = a.ws + (((size_t)pass * a.ksplit * a.ws_acc + j) * 128 + m) * 128 + col;
    const size_t stride = (size_t)a.ws_acc * 128 * 128;
#pragma unroll 4
    for (int k = sl; k < a.ksplit; k += WG_RED_KS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)k * stride));
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
  }
  part[sl][q] = sum;
  __syncthreads();
  if (sl == 0 && valid) {
#pragma unroll
    for (int s = 1; s < WG_RED_KS; ++s) { const float4 v = part[s][q]; sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w; }
    const int co = col & 63;
    float* dst = ((a.dw2 && pass >= a.dw2_from_pass) ? a.dw2 : a.dw) + ((size_t)co * a.Cin + ci) * a.KK + slot;
    const size_t cs = (size_t)a.Cin * a.KK;
    dst[0] += sum.x; dst[cs] += sum.y; dst[2 * cs] += sum.z; dst[3 * cs] += sum.w;
  }
}

// per-channel sum of a blocked tensor: out[c] += sum_{b,y,x} v[b,c,y,x]   (bias gradients)
__global__ void __launch_bounds__(256)
blk_channel_sum_kernel(const __nv_bfloat16* __restrict__ v, long long bs, int cb_off, int H, int W, int B,
                       float* __restrict__ out) {
  __shared__ float red[8][8];
  const int chunk = blockIdx.x;
  const int Hp = H + 4, Wp = W + 4;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  const long long total = (long long)B * H * W;
  for (long long e = (long long)blockIdx.y * 256 + threadIdx.x; e < total; e += (long long)gridDim.y * 256) {
    const int b = (int)(e / (H * W)), r = (int)(e % (H * W));
    const int y = r / W, x = r % W;
    const uint4 pk = __ldg(reinterpret_cast<const uint4*>(
        v + (size_t)b * bs + (((size_t)(cb_off + chunk) * Hp + y + 2) * Wp + x + 2) * 8));
    const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&pk);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += __bfloat162float(pb[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x >> 5][i] = s[i];
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + chunk * 8 + threadIdx.x, t);
  }
}

}  // namespace

enum { WG_K5S1 = 0, WG_K1 = 1, WG_K5S2 = 2, WG_K5S1_NARROW = 3, WG_UP_PHASE = 4, WG_K5S1_T = 5 };

// dw (+=) torch layout [64][Cin][k][k] fp32; dbias (+=) [64] fp32 or NULL (sum of dy over batch and pixels).  x: source view (n_chunks = 8 or 16; 32 = phase tensor for
// the stride-2 layers), dy: 8-chunk gradient view at the accumulator resolution.
// Bytes of the optional partial-sum workspace (max over kinds: 5 passes x <=148 CTAs x 5 accumulators x 32 KB).
CNP_API long long cnp_conv_tc_wgrad_workspace_bytes(void) {
  // <= 148 CTAs x <= 5 accumulators x [128][128] floats, + the per-CTA bias partials [148][64]
  return (long long)148 * 5 * 128 * 128 * sizeof(float) + (long long)148 * 64 * sizeof(float);
}

static int wgrad_launch(const cnp_blk* x, int n_chunks, const cnp_blk* dy, int kind, float* dw, float* dw2, int b_split,
                        float* dbias, int Cin, int B, void* workspace, long long workspace_bytes, cudaStream_t st);

CNP_API int cnp_conv_tc_wgrad(const cnp_blk* x, int n_chunks, const cnp_blk* dy, int kind, float* dw, float* dbias,
                              int Cin, int B, void* workspace, long long workspace_bytes, cudaStream_t st) {
  return wgrad_launch(x, n_chunks, dy, kind, dw, nullptr, 0, dbias, Cin, B, workspace, workspace_bytes, st);
}

// 5x5 stride-1 weight gradient of TWO image groups in one launch: images [0, b_split) accumulate into dw, images
// [b_split, B) into dw2 (both torch layout, +=).  The polyphase level's row strips and (transposed) column strips are
// such a pair (up_poly.cu); cnp_up_wgrad_fold adds dw2 tap-transposed.
CNP_API int cnp_conv_tc_wgrad_pair(const cnp_blk* x, int n_chunks, const cnp_blk* dy, float* dw, float* dw2, int b_split,
                                   float* dbias, int Cin, int B, void* workspace, long long workspace_bytes,
                                   cudaStream_t st) {
  CNP_REQUIRE(dw2 && b_split > 0 && b_split < B, "conv_tc_wgrad_pair: needs a second gradient and 0 < b_split < B");
  return wgrad_launch(x, n_chunks, dy, 0 /* WG_K5S1 */, dw, dw2, b_split, dbias, Cin, B, workspace, workspace_bytes, st);
}

static int wgrad_launch(const cnp_blk* x, int n_chunks, const cnp_blk* dy, int kind, float* dw, float* dw2, int b_split,
                        float* dbias, int Cin, int B, void* workspace, long long workspace_bytes, cudaStream_t st) {
  CNP_REQUIRE(x && dy && dw && B > 0, "conv_tc_wgrad: bad arguments");
  CNP_REQUIRE(x->H == dy->H && x->W == dy->W, "conv_tc_wgrad: x and dy must share the accumulator geometry");
  cnp_wg_args a;
  memset(&a, 0, sizeof(a));
  const bool tswap = kind == WG_K5S1_T;      // same reduction, gradient stored with the taps transposed (column strips)
  if (tswap) kind = WG_K5S1;
  const int H = dy->H, W = dy->W, Hp = H + 4, Wp = W + 4;
  a.x_plane = (long long)Hp * Wp * 8; a.dy_plane = a.x_plane;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x->base) + (long long)x->cb_off * a.x_plane; a.x_bs = x->bstride;
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy->base) + (long long)dy->cb_off * a.dy_plane; a.dy_bs = dy->bstride;
  a.dw = dw; a.dbias = dbias; a.Cin = Cin; a.B = B;
  // reduction tile: the zero pad after the last interior pixel (2*Wp+2 pixels) must cover the overshoot
  a.P = (2 * Wp + 2 >= 128) ? 128 : (2 * Wp + 2 >= 64 ? 64 : (2 * Wp + 2 >= 32 ? 32 : 16));
  CNP_REQUIRE(2 * Wp + 2 >= a.P, "conv_tc_wgrad: image too narrow");
  a.p_start = 2 * Wp + 2;
  const int p_end = (H + 1) * Wp + W + 2;
  a.tiles_per_img = cnp_cdiv(p_end - a.p_start, a.P);
  int np = 0;
  static const bool no_dup = getenv("CNP_WGRAD_NO_DUP") != nullptr;
  // only where it pays: 128-channel inputs (3 MMAs of 64 cycles instead of 5 of 48 per K step).  With 64-channel
  // inputs the second dY copy makes the stage L2-bound (measured 18 % slower), so those keep N = 64.
  // narrow kind: N = 128 measured 120 us against 151 us for five N = 64 MMAs per K step (16 channels, 304^2, B = 16)
  static const bool narrow_nodup = getenv("CNP_WGRAD_NARROW_NODUP") != nullptr;
  a.dup = ((!no_dup && kind == WG_K5S1 && n_chunks == 16) || (kind == WG_K5S1_NARROW && !narrow_nodup) ||
           kind == WG_UP_PHASE) ? 1 : 0;
  auto clear_slots = [](cnp_wg_pass& p) {
    for (int j = 0; j < CNP_WG_MAX_ACC; ++j)
      for (int h = 0; h < 2; ++h) p.slot[j][h][0] = p.slot[j][h][1] = -1;
  };
  auto tap = [tswap](int ky, int kx) { return (kx >= 0 && kx < 5) ? (tswap ? kx * 5 + ky : ky * 5 + kx) : -1; };
  if (kind == WG_K5S1) {
    CNP_REQUIRE(n_chunks == 8 || n_chunks == 16, "conv_tc_wgrad: 5x5 needs 8 or 16 source chunks");
    CNP_REQUIRE(Cin == n_chunks * 8, "conv_tc_wgrad: Cin mismatch");
    a.KK = 25;
    for (int ky = 0; ky < 5; ++ky, ++np) {
      cnp_wg_pass& p = a.pass[np];
      clear_slots(p);
      p.chunk0 = 0; p.base_off = (ky - 2) * Wp - 2; p.ci0 = 0;
      if (n_chunks == 16) {                 // rows = 128 input channels
        p.shift_px = -1; p.ci1 = 64;
        if (a.dup) {                        // X offsets 0,2,4: taps (a, a+1)
          p.n_acc = 3; p.a0 = 0; p.astep = 2;
          for (int j = 0; j < 3; ++j)
            for (int h = 0; h < 2; ++h) { p.slot[j][h][0] = tap(ky, 2 * j); p.slot[j][h][1] = tap(ky, 2 * j + 1); }
        } else {
          p.n_acc = 5; p.a0 = 0; p.astep = 1;
          for (int kx = 0; kx < 5; ++kx) p.slot[kx][0][0] = p.slot[kx][1][0] = tap(ky, kx);
        }
      } else {                              // rows = [64 channels ; the same 64 channels 3 pixels further]
        p.shift_px = 3; p.ci1 = 0;
        if (a.dup) {                        // offsets 0,1: row half 0 taps (a, a+1), row half 1 taps (a+3, a+4)
          p.n_acc = 2; p.a0 = 0; p.astep = 1;
          p.slot[0][0][0] = tap(ky, 0); p.slot[0][0][1] = tap(ky, 1); p.slot[0][1][0] = tap(ky, 3); p.slot[0][1][1] = tap(ky, 4);
          p.slot[1][0][1] = tap(ky, 2);     // (1, 4, 5 are already covered or out of range)
        } else {
          p.n_acc = 3; p.a0 = 0; p.astep = 1;
          for (int kx = 0; kx < 3; ++kx) { p.slot[kx][0][0] = tap(ky, kx); p.slot[kx][1][0] = tap(ky, kx + 3); }
        }
      }
    }
  } else if (kind == WG_K5S1_NARROW) {
    // few input channels (the folded first layer, fold_in.cu): the 128 M rows hold G = 16 / n_chunks kernel ROWS of
    // the same n_chunks source chunks (row group g = the source shifted by g image rows), and the N = 128 operand
    // [dY(p) ; dY(p-1)] adds the column pairs: 3 MMAs per 16-pixel K step cover G x 5 taps.
    CNP_REQUIRE(n_chunks >= 1 && n_chunks <= 8 && Cin == n_chunks * 8, "conv_tc_wgrad: narrow 5x5 needs 1..8 source chunks");
    a.KK = 25;
    a.narrow = n_chunks; a.grp_shift = Wp;
    const int G = 16 / n_chunks;
    for (int ky0 = 0; ky0 < 5; ky0 += G, ++np) {
      cnp_wg_pass& p = a.pass[np];
      clear_slots(p);
      p.chunk0 = 0; p.shift_px = -1; p.base_off = (ky0 - 2) * Wp - 2; p.ci0 = p.ci1 = 0;
      p.ky0 = ky0; p.n_grp = (5 - ky0 < G) ? 5 - ky0 : G;
      if (a.dup) { p.n_acc = 3; p.a0 = 0; p.astep = 2; } else { p.n_acc = 5; p.a0 = 0; p.astep = 1; }
    }
  } else if (kind == WG_K1) {
    CNP_REQUIRE(n_chunks == 8 && Cin == 64, "conv_tc_wgrad: 1x1 needs an 8-chunk source");
    a.KK = 1;
    cnp_wg_pass& p = a.pass[np++];
    clear_slots(p);
    p.chunk0 = 0; p.shift_px = 0; p.base_off = 0; p.n_acc = 1; p.a0 = 0; p.astep = 1; p.slot[0][0][0] = 0;
  } else if (kind == WG_K5S2) {
    CNP_REQUIRE(n_chunks == 32 && Cin == 64, "conv_tc_wgrad: stride-2 reads the 32-chunk phase tensor");
    a.KK = 25;
    for (int py = 0; py < 2; ++py)
      for (int ky = py; ky < 5; ky += 2, ++np) {
        cnp_wg_pass& p = a.pass[np];
        clear_slots(p);
        const int dyy = (ky - 2 - py) / 2;
        // rows = [x-phase 0 ; x-phase 1] of the space-to-depth tensor: offset j <-> taps kx = 2j (half 0), 2j+1 (half 1)
        p.chunk0 = 16 * py; p.shift_px = -1; p.base_off = dyy * Wp - 1; p.ci0 = 0; p.ci1 = 0;
        if (a.dup) {                        // offset j: taps (2j, 2j+2) for x-phase 0, (2j+1, 2j+3) for x-phase 1
          p.n_acc = 2; p.a0 = 0; p.astep = 1;
          p.slot[0][0][0] = tap(ky, 0); p.slot[0][0][1] = tap(ky, 2); p.slot[0][1][0] = tap(ky, 1); p.slot[0][1][1] = tap(ky, 3);
          p.slot[1][0][1] = tap(ky, 4);
        } else {
          p.n_acc = 3; p.a0 = 0; p.astep = 1;
          for (int j = 0; j < 3; ++j) { p.slot[j][0][0] = tap(ky, 2 * j); p.slot[j][1][0] = tap(ky, 2 * j + 1); }
        }
      }
  } else if (kind == WG_UP_PHASE) {
    // Polyphase resize-convolution (up_poly.cu): x = the LOW-res input (16 chunks), dy = the space-to-depth copy of dY
    // (32 chunks: (a*2+b)*8 + c).  dw = the phase gradients [a][b][64][Cin][4][4]:
    //   dWp[a,b,co,ci,p,q] = sum dY_ab[i, j] x[i + a + p - 2, j + b + q - 2]
    // One pass per (a, p); the N = 128 operand is [dY_a0(pix) ; dY_a1(pix - 1)], so the X operand at column offset j - 2
    // yields tap q = j of BOTH x-phases: 4 MMAs of N = 128 per 16-pixel K step, all 8 halves useful (64 tap-GEMMs per
    // low-res pixel against 4 x 30 slots of the 5x5 kernel on the upsampled tensor).
    CNP_REQUIRE(n_chunks == 16 && Cin == 128, "conv_tc_wgrad: the up-phase kind needs a 128-channel low-res input");
    a.KK = 16;
    a.dy_pair = 1;
    for (int ra = 0; ra < 2; ++ra)
      for (int pp = 0; pp < 4; ++pp, ++np) {
        cnp_wg_pass& p = a.pass[np];
        clear_slots(p);
        p.chunk0 = 0; p.shift_px = -1; p.ci0 = 0; p.ci1 = 64; p.dy_chunk0 = 16 * ra;
        p.base_off = (ra + pp - 2) * Wp - 2;
        p.n_acc = 4; p.a0 = 0; p.astep = 1;
        for (int j = 0; j < 4; ++j)
          for (int h = 0; h < 2; ++h)
            for (int b = 0; b < 2; ++b) p.slot[j][h][b] = (ra * 2 + b) * 64 * Cin * 16 + pp * 4 + j;
      }
  } else {
    CNP_REQUIRE(false, "conv_tc_wgrad: unknown kind %d", kind);
  }
  a.bias_grp = kind == WG_UP_PHASE ? 4 : np;
  if (dw2) {
    // second image group: the same passes again, over images [b_split, B), into dw2
    CNP_REQUIRE(kind == WG_K5S1 && 2 * np <= CNP_WG_MAX_PASS, "conv_tc_wgrad_pair: 5x5 stride-1 only");
    for (int i = 0; i < np; ++i) {
      a.pass[np + i] = a.pass[i];
      a.pass[i].img0 = 0; a.pass[i].n_img = b_split;
      a.pass[np + i].img0 = b_split; a.pass[np + i].n_img = B - b_split;
    }
    a.dw2 = dw2; a.dw2_from_pass = np;
    np *= 2;
  }
  if (a.dup) {
    // the shifted dY copy sums dY[p-1]: run one pixel further so that it still covers the last interior pixel
    a.tiles_per_img = cnp_cdiv(p_end - a.p_start + 1, a.P);
  }
  a.ws_acc = 1;
  for (int i = 0; i < np; ++i) a.ws_acc = a.pass[i].n_acc > a.ws_acc ? a.pass[i].n_acc : a.ws_acc;
  a.n_pass = np;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int total_tiles = (dw2 ? (b_split < B - b_split ? b_split : B - b_split) : B) * a.tiles_per_img;
  a.ksplit = sms / np;
  if (a.ksplit > total_tiles) a.ksplit = total_tiles;
  if (a.ksplit < 1) a.ksplit = 1;
  const size_t stage_b = (size_t)16 * (a.P + XPAD) * 16 + (size_t)(a.dup ? 16 : 8) * a.P * 16;
  const size_t smem = WG_STAGES * stage_b + (2 * WG_STAGES + 1) * 8 + 16 + 4 * 64 * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cnp_set_error("conv_tc_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr = smem;
  }
  // cluster of the five ky passes sharing the dY stream (dup path only: its stage is otherwise L2-bound); the K split
  // is limited to the number of clusters that can be resident at once
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 1; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(224); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attrs; cfg.numAttrs = 1;
  a.cluster = 1;
  // Measured on B200: the 5-CTA cluster halves the L2 traffic of the dup path but runs 2.2x SLOWER (only 26 clusters
  // fit, the five passes advance in lock-step, and cluster launches place CTAs differently), so it stays opt-in.
  static const bool use_cluster = getenv("CNP_WGRAD_CLUSTER") != nullptr;
  if (a.dup && np == 5 && use_cluster) {
    attrs[0].val.clusterDim.y = 5;
    cfg.gridDim = dim3(a.ksplit, np);
    int max_clusters = 0;
    const cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, wgrad_tc_kernel<true>, &cfg);
    if (getenv("CNP_WGRAD_VERBOSE")) fprintf(stderr, "wgrad: max active clusters of 5 = %d (%s), ksplit %d\n", max_clusters, cudaGetErrorString(oe), a.ksplit);
    if (oe == cudaSuccess && max_clusters >= 16) {
      a.cluster = 5;
      if (a.ksplit > max_clusters) a.ksplit = max_clusters;
    } else {
      cudaGetLastError();
      attrs[0].val.clusterDim.y = 1;
    }
  }
  a.ws = nullptr;
  // partial sums through the workspace + reduce kernel; only very short K loops keep the atomics
  // (with the 16 B / 8-partial reduce kernel the workspace wins down to small layers: 152^2 142.6 -> 130.6 us,
  // 38^2 41.8 -> 34.3 us against the atomics epilogue)
  // (a caller that hands over a workspace gets the ordered reduction for every size: deterministic gradients)
  static const int ws_min_tiles = getenv("CNP_WGRAD_WS_MIN_TILES") ? atoi(getenv("CNP_WGRAD_WS_MIN_TILES")) : 1;
  const long long ws_part = (long long)np * a.ksplit * a.ws_acc * 128 * 128 * (long long)sizeof(float);
  if (workspace && total_tiles >= ws_min_tiles && workspace_bytes >= ws_part + (long long)np * a.ksplit * 64 * (long long)sizeof(float)) {
    a.ws = reinterpret_cast<float*>(workspace);
    a.ws_bias = a.ws + ws_part / sizeof(float);
  }
  cfg.gridDim = dim3(a.ksplit, np);
  if (a.cluster > 1) {
    cudaError_t le = cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<true>, a);
    if (le != cudaSuccess) { cnp_set_error("wgrad_tc_kernel: %s", cudaGetErrorString(le)); return (int)le; }
  } else {   // plain launch: a launch carrying a cluster attribute (even 1x1x1) was measured ~15 % slower here
    wgrad_tc_kernel<false><<<cfg.gridDim, 224, smem, st>>>(a);
  }
  CNP_LAUNCH_CHECK("wgrad_tc_kernel");
  if (a.ws) {
    const int total4 = np * a.ws_acc * 128 * ((a.dup ? 128 : 64) / 4);
    wgrad_reduce_kernel<<<dim3(cnp_cdiv(total4, 32)), 256, 0, st>>>(a);
    CNP_LAUNCH_CHECK("wgrad_reduce_kernel");
  }
  return 0;
}

CNP_API int cnp_blk_channel_sum(const cnp_blk* v, int n_chunks, int B, float* out, cudaStream_t st) {
  CNP_REQUIRE(v && out && n_chunks > 0 && B > 0, "blk_channel_sum: bad arguments");
  dim3 grid(n_chunks, 32);
  blk_channel_sum_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(v->base), v->bstride, v->cb_off,
                                               v->H, v->W, B, out);
  CNP_LAUNCH_CHECK("blk_channel_sum_kernel");
  return 0;
}
