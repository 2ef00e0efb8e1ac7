// bf16 UNet convolutions on tcgen05, second formulation: WEIGHTS are the M operand, PIXELS the N operand.
//
// Replaces the torch Conv2d stack (forward and input-gradient) of upstream neuralprocesses' UNet
// (reached from ConvNP.loss_fn nzdownscale/downscaler/train.py:370 and
// train_epoch train.py:388-394).
//
// Why this shape.  tools/mma_rate.cu (profiles/r01_mma_rate.txt) measures 48 cycles per M128xN64xK16
// tcgen05.mma with both operands in shared memory -- the 6 KB of operand reads per MMA exceed the
// 128 B/clk shared-memory port, capping a pixels-as-M / 64-channels-as-N kernel at 2/3 of the tensor
// peak -- while N >= 128 runs at the floor (N/2 cycles).  A UNet layer only has 64 output channels, so
// here the roles are swapped:
//   D[m, n]  m = 128 TMEM lanes = 2 groups x 64 output channels, n = up to 256 consecutive pixels of a row
//   A (M operand) = packed weights  [128 x 16 input channels]  (4 KB, K-major, SWIZZLE_NONE)
//   B (N operand) = the activation window in shared memory, [chunk][pixel][8 ch]: K-major with 16 B pixel
//                   rows, so a conv tap (dy,dx) is again just a +(dy*pitch+dx)*16 B shift of the descriptor.
// The two lane groups are
//   PAIR mode (64 output channels): group 0 = output row r, group 1 = output row r+1.  The window row
//        r+i feeds tap ky=i of group 0 and tap ky=i-1 of group 1, so A = [W(i,kx) ; W(i-1,kx)] and a 5x5
//        conv takes 6x5 MMAs per 16 input channels for 2 output rows (5/6 of the MMA slots useful);
//   WIDE mode (128 output channels, the dgrad of the 128->64 layers): group g = channels 64g..64g+63.
// One accumulator = one output row (WIDE) or row pair (PAIR) x N pixels = N TMEM columns; a tile keeps
// floor(512/N) accumulators and streams K in blocks of 16 channels: window blocks (2 chunks x (TH+4) rows)
// double-buffered, weights through a 3-deep ring of 20 KB stages (5 positions x 4 KB), all by cp.async.bulk.
// Warps: 0 = window producer, 1 = weight producer, 2 = MMA issuer, 3 = TMEM allocator, 4..15 = epilogue
// (tcgen05.ld 32 pixels of one channel per thread -> bias/ReLU -> transpose through shared memory ->
// 16 B blocked stores, optional ReLU-mask / accumulate; or direct fp32 NCHW stores).
#include "tc_common.cuh"
#include <stdlib.h>

#define C2_MAX_KB 16
#define C2_MAX_TYPES 4
#define C2_MAX_POS 30

struct cnp_c2_plan {
  int n_kb;
  int regular;                  // every stage = 5 consecutive pixels of window row <stage index> (5x5 stride-1 kinds)
  int kb_chunk0[C2_MAX_KB];     // first of the 2 source chunks of this 16-channel K block
  int kb_wci0[C2_MAX_KB];       // weight input-channel base (packing)
  int kb_type[C2_MAX_KB];       // position list used by this K block
  int kb_wsel[C2_MAX_KB];       // which of several stacked weight tensors this K block reads (packing; phase dgrad)
  int t_npos[C2_MAX_TYPES];
  short t_boff[C2_MAX_TYPES][C2_MAX_POS];          // B start, in window pixels (row*pitch + col)
  signed char t_tap[C2_MAX_TYPES][C2_MAX_POS][4];  // (ky0,kx0) of group 0, (ky1,kx1) of group 1; -1 = zero block
};

struct cnp_c2_args {
  const __nv_bfloat16* x; long long x_bs; int x_Hp, x_Wp;
  const uint8_t* w;
  const uint8_t* w2; int w2_from_b;   // optional second packed weight tensor, used by the images b >= w2_from_b (strips)
  int B, H, W;
  int TW, TH, pitch, N, nacc, rpa, plane_sm, tiles_x, tiles_y;
  int nbuf;                     // 2: two accumulator sets of 256 TMEM columns -- the epilogue of tile t overlaps the MMAs of t+1
  int wide;
  int pxpair;                   // stride-2 dgrad: lane group g = output x-phase g (both phases of a row in one launch)
  int n_work, split_from, split;  // work items: tiles [0, split_from) whole, the leftover tiles of the last round
                                  // split into `split` single-accumulator items (finer tail, see cnp_conv_tc2)
  int cluster;                   // 1, or 2: CTA pairs share the weight stream (each loads half a stage and multicasts it)
  int out_mode;                  // 0: blocked bf16, 1: NCHW fp32
  void* out; long long out_bs; int out_c_off; int out_Hp, out_Wp;
  int sy, ay, sx, ax;
  const float* bias; int relu;
  const __nv_bfloat16* mask; long long mask_bs; int mask_cb_off;
  int accumulate;
  __nv_bfloat16* s2d; long long s2d_bs;   // optional second output: space-to-depth copy (32 chunks at half resolution)
  int s2d_c0, s2d_band;          // ... of the 8 output chunks from s2d_c0 on, skipping a band of s2d_band half-res pixels
  int dbg_flags;                 // profiling only: 1 = epilogue stops after the TMEM load, 2 = skips the global stores
  long long* dbg;                // optional [grid][8] cycle counters (cnp_conv_tc2_debug), else NULL
  cnp_c2_plan plan;
};

static long long* g_c2_dbg = nullptr;
static int g_c2_dbg_flags = 0;
static int g_c2_cluster = 1;   // 2: CTA pairs share the weight stream by multicast (cnp_conv_tc2_set_cluster)

namespace {

constexpr int C2_POS_BYTES = 4096;                       // 2 k-chunks x 128 rows x 16 B
constexpr int C2_STAGE_POS = 5;
constexpr int C2_STAGE_BYTES = C2_STAGE_POS * C2_POS_BYTES;
constexpr int C2_WSTAGES = 3;
constexpr int C2_ABUFS = 2;
constexpr int C2_EPI_WARPS = 12;                        // generic instantiation (123 registers per thread)
constexpr int C2_EPI_WARPS_SPEC = 16;                   // specialised epilogues (<= 102 registers): 4 warps per TMEM quadrant
constexpr int C2_THREADS = 32 * (4 + C2_EPI_WARPS);
constexpr int C2_STG_WORDS_BF16 = 32 * 20;               // per-warp transpose buffer [32 px][64 B of channels + 16 B pad]
constexpr int C2_STG_WORDS_F32 = 32 * 36;                // fp32 output: [32 ch][36 floats]

// ReLU-backward keep mask of a packed bf16 pair: 0xffff per half whose value is > 0 (non-zero magnitude, sign clear).
// Integer form, both halves at once: 5 instructions per word instead of two conversions, two compares and two selects --
// the mask arithmetic sits on the critical path of an epilogue warp (3 warps per scheduler, little to overlap with).
__device__ __forceinline__ uint32_t relu_keep(uint32_t w) {
  const uint32_t nz = ((w & 0x7fff7fffu) + 0x7fff7fffu) & 0x80008000u;   // bit 15 / 31: magnitude != 0
  return ((nz & ~w) >> 15) * 0xffffu;
}

struct c2_work { int b, y0, x0, nacc; };
__device__ __forceinline__ c2_work decode_work(const cnp_c2_args& a, int w) {
  int tile = w, j = 0;
  const bool sub = w >= a.split_from && a.split > 1;
  if (w >= a.split_from) { tile = a.split_from + (w - a.split_from) / a.split; j = (w - a.split_from) % a.split; }
  const int per_img = a.tiles_x * a.tiles_y;
  const int tr = tile % per_img;
  c2_work r;
  r.b = tile / per_img;
  r.y0 = (tr / a.tiles_x) * a.TH + j * a.rpa;
  r.x0 = (tr % a.tiles_x) * a.TW;
  const int need = (a.H - r.y0 + a.rpa - 1) / a.rpa;
  r.nacc = sub ? 1 : (need < a.nacc ? need : a.nacc);
  if (r.nacc < 1) r.nacc = 1;
  return r;
}

// positions [P0, P1) of a regular stage (consecutive window pixels of one row) x NACC accumulators, fully unrolled
template <int NACC, int P0, int P1>
__device__ __forceinline__ void issue_row(uint32_t tmem, uint32_t w_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                          uint32_t N, uint32_t acc_step16, uint32_t first) {
#pragma unroll
  for (int p = P0; p < P1; ++p) {
#pragma unroll
    for (int j = 0; j < NACC; ++j)
      tc::mma_bf16_ss_lohi(tmem + j * N, w_lo + p * (C2_POS_BYTES >> 4), desc_hi, b_lo + p + j * acc_step16, desc_hi, idesc,
                           p > 0 ? 1u : first);
  }
}
template <int P0, int P1>
__device__ __forceinline__ void issue_row_n(int nacc, uint32_t tmem, uint32_t w_lo, uint32_t b_lo, uint32_t desc_hi,
                                            uint32_t idesc, uint32_t N, uint32_t acc_step16, uint32_t first) {
#pragma unroll
  for (int p = P0; p < P1; ++p)
    for (int j = 0; j < nacc; ++j)
      tc::mma_bf16_ss_lohi(tmem + j * N, w_lo + p * (C2_POS_BYTES >> 4), desc_hi, b_lo + p + j * acc_step16, desc_hi, idesc,
                           p > 0 ? 1u : first);
}

// CL = true: instantiation with the cluster / multicast code.  A kernel that contains cluster instructions is scheduled
// differently even in a plain launch (measured ~15 % slower on the wgrad kernel), hence two instantiations.
// EPI: epilogue specialisation.  The epilogue is sensitive to every instruction and live register (adding code it never
// executed slowed the unmasked WIDE dgrad by 5 %), so the configurations the engine launches get their flags as
// compile-time constants: EPI < 0 = generic (every flag read at run time: fp32 output, debug counters, anything else),
// EPI >= 0 = blocked bf16 output with exactly the flag bits below set.
enum { EF_MASK = 1, EF_ACC = 2, EF_PX = 4, EF_WIDE = 8, EF_BIASRELU = 16, EF_S2D = 32 };
// EW: epilogue warps.  An item of the epilogue is a chain of latencies (TMEM load, transposition through shared memory,
// global accesses), so the epilogue's length is set by how many items are in flight: the specialised instantiations fit
// 16 epilogue warps (640 threads x <= 102 registers) where the generic one has 12.
template <bool CL, int EPI, int EW>
__global__ void __launch_bounds__(32 * (4 + EW), 1)
conv_tc2_kernel(const __grid_constant__ cnp_c2_args a) {
  const int a_cluster = CL ? a.cluster : 1;
  extern __shared__ __align__(128) uint8_t smem[];
  const int plane_bytes = a.plane_sm * 16;
  const int abuf_bytes = 2 * plane_bytes;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + C2_ABUFS * abuf_bytes;
  uint32_t* stg = reinterpret_cast<uint32_t*>(w_smem + C2_WSTAGES * C2_STAGE_BYTES);
  const int stg_words = a.out_mode == 1 ? C2_STG_WORDS_F32 : C2_STG_WORDS_BF16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + EW * stg_words);
  uint64_t* a_full = bars;             // [2]
  uint64_t* a_empty = bars + 2;        // [2]
  uint64_t* w_full = bars + 4;         // [3]
  uint64_t* w_empty = bars + 7;        // [3]
  uint64_t* acc_full = bars + 10;      // [2]
  uint64_t* acc_empty = bars + 12;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  __shared__ uint32_t pos_tbl[C2_MAX_TYPES][C2_MAX_POS + 2];   // B offsets (16 B units) of the generic plans

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = a.n_work;   // work items (tiles, with the tail of the last round split finer)

  if (threadIdx.x == 0) {
    for (int i = 0; i < C2_ABUFS; ++i) { tc::mbar_init(a_full + i, 1); tc::mbar_init(a_empty + i, 1); }
    for (int i = 0; i < C2_WSTAGES; ++i) { tc::mbar_init(w_full + i, 1); tc::mbar_init(w_empty + i, (uint32_t)a_cluster); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, EW); }
    tc::mbar_fence_init();
  }
  if (warp == 3) tc::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x >= 128 && threadIdx.x < 128 + C2_MAX_TYPES * C2_MAX_POS) {
    const int e = threadIdx.x - 128;
    pos_tbl[e / C2_MAX_POS][e % C2_MAX_POS] = (uint32_t)a.plan.t_boff[e / C2_MAX_POS][e % C2_MAX_POS];
  }
  tc::fence_before_sync();
  __syncthreads();
  if ((CL && a_cluster > 1)) tc::cluster_sync();   // the partner's barriers are initialised before anything is multicast to them
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_iter = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;   // a CTA without a tile in the last round still
  const bool ghost = (CL && a_cluster > 1) && (int)blockIdx.x + (n_iter - 1) * (int)gridDim.x >= ntiles;   // streams weights
  const uint32_t crank = (CL && a_cluster > 1) ? tc::cluster_ctarank() : 0u;

  const uint32_t row_bytes = (uint32_t)a.pitch * 16u;
  const long long plane_g = (long long)a.x_Hp * a.x_Wp * 8;

  if (warp == 0) {
    // ===================== window producer ======================================================
    if (tc::elect_one()) {
      uint32_t a_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const c2_work wk = decode_work(a, tile);
        const int b = wk.b, y0 = wk.y0, x0 = wk.x0;
        const int rows = wk.nacc * a.rpa + 4;          // window rows this work item needs
        const __nv_bfloat16* xb = a.x + (long long)b * a.x_bs + ((long long)y0 * a.x_Wp + x0) * 8;
        for (int kb = 0; kb < a.plan.n_kb; ++kb, ++a_it) {
          const int buf = a_it & 1;
          tc::mbar_wait(a_empty + buf, ((a_it >> 1) & 1) ^ 1);
          tc::mbar_expect_tx(a_full + buf, 2u * rows * row_bytes);
          uint8_t* dst = a_smem + buf * abuf_bytes;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const __nv_bfloat16* src = xb + (long long)(a.plan.kb_chunk0[kb] + c) * plane_g;
            for (int r = 0; r < rows; ++r)
              tc::bulk_g2s(dst + c * plane_bytes + r * row_bytes, src + (long long)r * a.x_Wp * 8, row_bytes,
                           a_full + buf);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== weight producer ======================================================
    if (tc::elect_one()) {
      uint32_t w_it = 0;
      for (int it = 0; it < n_iter; ++it) {
        if (a_cluster == 1 && (int)blockIdx.x + it * (int)gridDim.x >= ntiles) break;
        const uint8_t* wsrc = a.w;
        if (a.w2 && decode_work(a, (int)blockIdx.x + it * (int)gridDim.x).b >= a.w2_from_b) wsrc = a.w2;
        for (int kb = 0; kb < a.plan.n_kb; ++kb) {
          const int npos = a.plan.t_npos[a.plan.kb_type[kb]];
          for (int s0 = 0; s0 < npos; s0 += C2_STAGE_POS, ++w_it) {
            const int ws = w_it % C2_WSTAGES;
            const int np = min(C2_STAGE_POS, npos - s0);
            const uint32_t bytes = (uint32_t)np * C2_POS_BYTES;
            tc::mbar_wait(w_empty + ws, ((w_it / C2_WSTAGES) & 1) ^ 1);   // cluster: BOTH CTAs released the slot
            tc::mbar_expect_tx(w_full + ws, bytes);
            if ((CL && a_cluster > 1)) {
              // my half of the stage, delivered to both CTAs of the pair (one L2 read instead of two)
              const uint32_t half = bytes >> 1;
              tc::bulk_g2s_mc(w_smem + ws * C2_STAGE_BYTES + crank * half, wsrc + crank * half, half, w_full + ws, 0x3);
            } else {
              tc::bulk_g2s(w_smem + ws * C2_STAGE_BYTES, wsrc, bytes, w_full + ws);
            }
            wsrc += (size_t)bytes;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer ===========================================================
    // One flat loop over (tile, K block, stage).  The barrier waits for stage s+1 are issued BEFORE the last
    // position of stage s: the tensor pipe queues only a few MMAs, so a wait placed between two stages would
    // leave it idle for the wait's latency (measured: ~60 cycles per stage, fully exposed).
    if (tc::elect_one() && (int)blockIdx.x < ntiles) {
      const uint32_t idesc = tc::make_idesc_bf16(128, a.N, 0, 0);
      const uint32_t acc_step16 = (uint32_t)(a.rpa * a.pitch);        // accumulator-to-accumulator B shift, 16 B units
      const uint32_t b_lbo = ((uint32_t)plane_bytes >> 4) << 16;      // LBO field of the B descriptor (chunk plane stride)
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);              // SBO = 128 B, descriptor version 1
      const uint32_t w_lbo = 128u << 16;                              // LBO 2048 B of the packed weights
      const uint32_t Ncols = (uint32_t)a.N;
      long long c_acc = 0, c_a = 0, c_w = 0, t0 = 0;
      const long long t_begin = clock64();

      auto wait_stage = [&](int s0, uint32_t a_it, uint32_t w_it) {
        if (a.dbg) t0 = clock64();
        if (s0 == 0) tc::mbar_wait(a_full + (a_it & 1), (a_it >> 1) & 1);
        if (a.dbg) { const long long t1 = clock64(); c_a += t1 - t0; t0 = t1; }
        tc::mbar_wait(w_full + w_it % C2_WSTAGES, (w_it / C2_WSTAGES) & 1);
        if (a.dbg) c_w += clock64() - t0;
        tc::fence_after_sync();
      };
      auto tile_nacc = [&](int tile) { return decode_work(a, tile).nacc; };

      int tile = blockIdx.x, kb = 0, s0 = 0;
      uint32_t a_it = 0, w_it = 0, t_it = 0;
      int nacc = tile_nacc(tile);
      wait_stage(0, 0, 0);
      for (;;) {
        const int type = a.plan.kb_type[kb];
        const int npos = a.plan.t_npos[type];
        const int np = min(C2_STAGE_POS, npos - s0);
        const int ws = w_it % C2_WSTAGES, buf = a_it & 1;
        const uint32_t tmem_acc = tmem_base + (a.nbuf > 1 ? (t_it & 1u) * 256u : 0u);   // this tile's accumulator set
        int n_s0 = s0 + C2_STAGE_POS, n_kb = kb, n_tile = tile;
        uint32_t n_a_it = a_it;
        if (n_s0 >= npos) {
          n_s0 = 0; ++n_kb; ++n_a_it;
          if (n_kb == a.plan.n_kb) { n_kb = 0; n_tile += gridDim.x; }
        }
        const bool kb_end = n_s0 == 0, tile_end = n_tile != tile;
        const uint32_t w_base = tc::smem_u32(w_smem + ws * C2_STAGE_BYTES);
        const uint32_t a16 = tc::smem_u32(a_smem + buf * abuf_bytes) >> 4;
        const uint32_t first = (kb > 0 || s0 > 0) ? 1u : 0u;
        if (a.plan.regular) {
          // stage = window row s0/5, positions = 5 consecutive pixels: only 32-bit adds between MMAs
          const uint32_t w_lo = ((w_base >> 4) & 0x3FFFu) | w_lbo;
          const uint32_t b_lo = ((a16 + (uint32_t)(s0 / C2_STAGE_POS) * (uint32_t)a.pitch) & 0x3FFFu) | b_lbo;
          if (nacc == 3) issue_row<3, 0, 4>(tmem_acc, w_lo, b_lo, desc_hi, idesc, Ncols, acc_step16, first);
          else if (nacc == 2) issue_row<2, 0, 4>(tmem_acc, w_lo, b_lo, desc_hi, idesc, Ncols, acc_step16, first);
          else issue_row_n<0, 4>(nacc, tmem_acc, w_lo, b_lo, desc_hi, idesc, Ncols, acc_step16, first);
          if (!tile_end) wait_stage(n_s0, n_a_it, w_it + 1);
          if (nacc == 3) issue_row<3, 4, 5>(tmem_acc, w_lo, b_lo, desc_hi, idesc, Ncols, acc_step16, 1u);
          else if (nacc == 2) issue_row<2, 4, 5>(tmem_acc, w_lo, b_lo, desc_hi, idesc, Ncols, acc_step16, 1u);
          else issue_row_n<4, 5>(nacc, tmem_acc, w_lo, b_lo, desc_hi, idesc, Ncols, acc_step16, 1u);
        } else {
          // generic plans (stride-2, 1x1): offsets of the stage's <= 5 positions come from shared memory, loaded
          // together up front so the MMAs are separated only by 32-bit adds
          uint32_t boff[C2_STAGE_POS];
#pragma unroll
          for (int p = 0; p < C2_STAGE_POS; ++p) boff[p] = pos_tbl[type][min(s0 + p, C2_MAX_POS - 1)];
          const uint32_t w_lo0 = ((w_base >> 4) & 0x3FFFu) | w_lbo;
#pragma unroll
          for (int p = 0; p < C2_STAGE_POS; ++p) {
            if (p < np) {
              if (p == np - 1 && !tile_end) wait_stage(n_s0, n_a_it, w_it + 1);
              const uint32_t b_lo = ((a16 + boff[p]) & 0x3FFFu) | b_lbo;
              const uint32_t accf = (p > 0) ? 1u : first;
              if (nacc == 3) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  tc::mma_bf16_ss_lohi(tmem_acc + j * Ncols, w_lo0 + p * (C2_POS_BYTES >> 4), desc_hi, b_lo + j * acc_step16,
                                       desc_hi, idesc, accf);
              } else {
                for (int j = 0; j < nacc; ++j)
                  tc::mma_bf16_ss_lohi(tmem_acc + j * Ncols, w_lo0 + p * (C2_POS_BYTES >> 4), desc_hi, b_lo + j * acc_step16,
                                       desc_hi, idesc, accf);
              }
            }
          }
        }
        if ((CL && a_cluster > 1)) tc::mma_commit_mc(w_empty + ws, 0x3); else tc::mma_commit(w_empty + ws);
        if (kb_end) tc::mma_commit(a_empty + buf);
        ++w_it;
        if (tile_end) {
          tc::mma_commit(acc_full + (a.nbuf > 1 ? (t_it & 1u) : 0u));
          ++t_it;
          if (n_tile >= ntiles) break;
          nacc = tile_nacc(n_tile);
          if (a.dbg) t0 = clock64();
          // the epilogue has drained the accumulator set the next tile uses (two sets: its use before last)
          if (a.nbuf > 1) tc::mbar_wait(acc_empty + (t_it & 1u), ((t_it >> 1) & 1u) ^ 1u);
          else tc::mbar_wait(acc_empty, (t_it & 1) ^ 1);
          if (a.dbg) c_acc += clock64() - t0;
          wait_stage(0, n_a_it, w_it);
        }
        tile = n_tile; kb = n_kb; s0 = n_s0; a_it = n_a_it;
      }
      if (ghost) {
        // no tile in the last round, but the partner still needs this CTA's half of every weight stage: keep the
        // ring turning (wait for the data, release the slot in both CTAs)
        for (int kb2 = 0; kb2 < a.plan.n_kb; ++kb2) {
          const int npos = a.plan.t_npos[a.plan.kb_type[kb2]];
          for (int q0 = 0; q0 < npos; q0 += C2_STAGE_POS, ++w_it) {
            tc::mbar_wait(w_full + w_it % C2_WSTAGES, (w_it / C2_WSTAGES) & 1);
            tc::fence_after_sync();
            tc::mma_commit_mc(w_empty + w_it % C2_WSTAGES, 0x3);
          }
        }
      }
      if (a.dbg) {
        long long* d = a.dbg + blockIdx.x * 8;
        d[0] = clock64() - t_begin; d[1] = c_acc; d[2] = c_a; d[3] = c_w; d[4] = t_it;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue ==============================================================
    constexpr bool EG = EPI < 0;
    const bool f_mode0 = EG ? a.out_mode == 0 : true;                       // blocked bf16 output
    const bool f_px = EG ? a.pxpair != 0 : (EPI & EF_PX) != 0;
    const bool f_mask = EG ? a.mask != nullptr : (EPI & EF_MASK) != 0;
    const bool f_acc = EG ? a.accumulate != 0 : (EPI & EF_ACC) != 0;
    const bool f_wide = EG ? a.wide != 0 : (EPI & EF_WIDE) != 0;
    const bool f_bias = EG ? a.bias != nullptr : (EPI & EF_BIASRELU) != 0;
    const bool f_relu = EG ? a.relu != 0 : (EPI & EF_BIASRELU) != 0;
    const bool f_s2d = EG ? a.s2d != nullptr : (EPI & EF_S2D) != 0;
    const int f_dbgf = EG ? a.dbg_flags : 0;
    long long* const f_dbg = EG ? a.dbg : nullptr;
    const int ew = warp - 4;
    const int q = warp & 3;           // TMEM lane quadrant this warp may read
    const int h = ew >> 2;            // this quadrant's (accumulator, column block) items are dealt round-robin to its warps
    const int m = q * 32 + lane;
    const int g = m >> 6;
    const int ch = f_wide ? m : (m & 63);                    // output channel of this thread (TMEM lane)
    // x-phase pairs: the 32 lanes of a quadrant are [16 channels of x-phase 0 | the same 16 channels of x-phase 1] (see
    // pack2_kernel), so that a thread ends up with BOTH output pixels 2x, 2x+1 of two chunks: 32 B stores / loads
    const int chunk0 = a.out_c_off + (f_wide ? q * 4 : (f_px ? q * 2 : (q & 1) * 4));   // first output chunk of this warp
    const float bias_v = f_bias ? __ldg(a.bias + ch) : 0.f;
    // bias of the four channels this thread holds in the fragment distribution: warp channel base + lane/4 + {0,8,16,24}
    float bias4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      bias4[i] = !f_bias ? 0.f : (f_px ? __ldg(a.bias + 16 * q + (lane >> 2) + 8 * (i & 1))
                                           : __ldg(a.bias + (ch - lane) + (lane >> 2) + 8 * i));
    uint32_t* my_stg = stg + ew * stg_words;
    const long long oplane = (long long)a.out_Hp * a.out_Wp;
    const int ncb = a.N >> 5;
    // The epilogue is instruction-bound (ncu: 0.7 warp instructions per scheduler per cycle), so its index arithmetic is
    // kept off the per-item path: this warp's items (accumulator j, column block cb) are STEPPED -- h, h+3, h+6 ... of
    // nacc x ncb, the same sequence for every tile -- instead of divided out of an item index, and every global address is
    // a per-tile 64-bit base of this thread plus a 32-bit element offset of the item.
    const int it_j0 = h / ncb, it_cb0 = h - it_j0 * ncb;
    const int it_dj = (EW / 4) / ncb, it_dcb = (EW / 4) - it_dj * ncb;
    const int gsel = (f_wide || f_px) ? 0 : g;                      // PAIR: lane group = second row of the accumulator
    const int row_eoff = a.rpa * a.sy * a.out_Wp * 8;                   // element offset between accumulators (rows)
    const int col_eoff = 32 * a.sx * 8;                                 // ... between column blocks
    const int cplane = (int)(oplane * 8);                               // ... between chunk planes
    uint32_t t_it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_it) {
      const c2_work wk = decode_work(a, tile);
      const int b = wk.b, y0 = wk.y0, x0 = wk.x0;
      // per-tile bases of this thread (pixel x0 + lane of row y0 [+ lane group], chunk chunk0)
      const long long pix_t = (long long)((y0 + gsel) * a.sy + a.ay + 2) * a.out_Wp + ((x0 + lane) * a.sx + a.ax + 2);
      __nv_bfloat16* const obase_t = reinterpret_cast<__nv_bfloat16*>(a.out) + (long long)b * a.out_bs +
                                     ((long long)chunk0 * oplane + pix_t) * 8;
      const __nv_bfloat16* const mbase_t = f_mask ? a.mask + (long long)b * a.mask_bs +
          ((long long)(a.mask_cb_off + chunk0 - a.out_c_off) * oplane + pix_t) * 8 : nullptr;
      if (f_mode0 && (f_mask || f_acc)) {
        // The MMAs of this tile are still running and these warps would only wait: pull the ReLU-mask (and the
        // accumulate target) lines of the tile into L2 now, so the epilogue's loads are not DRAM round trips issued
        // while the tensor pipe is idle (masked WIDE dgrad at 304^2: 555 us against 444 us unmasked before this).
        for (int j = it_j0, cb = it_cb0; j < wk.nacc;) {
          const int xs = cb * 32;
          if (y0 + j * a.rpa + gsel < a.H && xs < a.TW && x0 + xs < a.W) {
            // lanes past the image edge prefetch the last valid pixel's line
            const int over = x0 + xs + lane - (a.W - 1);
            const int eoff = j * row_eoff + cb * col_eoff - (over > 0 ? over * a.sx * 8 : 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (f_px && c >= 2) break;
              if (f_mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(mbase_t + eoff + c * cplane));
              if (f_acc) asm volatile("prefetch.global.L2 [%0];" ::"l"(obase_t + eoff + c * cplane));
            }
          }
          cb += it_dcb; j += it_dj;
          if (cb >= ncb) { cb -= ncb; ++j; }
        }
      }
      const uint32_t abuf = a.nbuf > 1 ? (t_it & 1u) : 0u;
      const uint32_t tmem_acc = tmem_base + abuf * 256u;
      tc::mbar_wait(acc_full + abuf, a.nbuf > 1 ? ((t_it >> 1) & 1u) : (t_it & 1u));
      const long long te0 = f_dbg ? clock64() : 0;
      tc::fence_after_sync();
      for (int j = it_j0, cb = it_cb0, jn, cbn; j < wk.nacc; j = jn, cb = cbn) {
        cbn = cb + it_dcb; jn = j + it_dj;
        if (cbn >= ncb) { cbn -= ncb; ++jn; }
        const int y = y0 + j * a.rpa + gsel;
        if (y >= a.H) continue;                              // warp-uniform
        const int oy = y * a.sy + a.ay;
        const int eoff = j * row_eoff + cb * col_eoff;       // this item's element offset from the per-tile bases
        {
          const int xs = cb * 32;
          if (xs >= a.TW || x0 + xs >= a.W) continue;        // warp-uniform
          const int nvalid = min(min(a.TW - xs, a.W - x0 - xs), 32);   // valid pixels of this block
          uint32_t wv[16];
          if (f_mode0) {
            // ---- blocked bf16: [channel = TMEM lane][pixel] -> [pixel = thread][32 channels] ----
            // The accumulator block is read in the mma-fragment distribution (two 16-lane halves), bias / ReLU are
            // applied, pairs are packed to bf16 and four stmatrix.x4.trans write the block transposed into shared
            // memory as [pixel][8 channels] rows; every thread then reads its pixel's 64 B back with four 16 B loads.
            // (The first version -- thread = channel, 32 two-byte stores + 16 word loads per block -- kept the
            // shared-memory instruction pipe busy for half of the epilogue: tools/bench_conv.py stall counters.)
            uint32_t ra[16], rb[16];
            const uint32_t t0 = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * a.N + xs);
            tc::tmem_ld_16x256b_x4(t0, ra);
            tc::tmem_ld_16x256b_x4(t0 + (16u << 16), rb);
            tc::tmem_ld_wait();
            if (f_dbgf & 1) { uint32_t xacc = 0;
#pragma unroll
              for (int i = 0; i < 16; ++i) xacc ^= ra[i] ^ rb[i];
              if (xacc == 0x12345678u) f_dbg[0] = 1;
              continue; }
            const uint32_t srow = tc::smem_u32(my_stg) + (uint32_t)(lane & 7) * 80u + (uint32_t)(lane >> 3) * 16u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float f[8];
              f[0] = __uint_as_float(ra[4 * k]);     f[1] = __uint_as_float(ra[4 * k + 1]);
              f[2] = __uint_as_float(ra[4 * k + 2]); f[3] = __uint_as_float(ra[4 * k + 3]);
              f[4] = __uint_as_float(rb[4 * k]);     f[5] = __uint_as_float(rb[4 * k + 1]);
              f[6] = __uint_as_float(rb[4 * k + 2]); f[7] = __uint_as_float(rb[4 * k + 3]);
              if (f_bias) {                          // (the input gradients have none: 32 FADDs per item less)
                f[0] += bias4[0]; f[1] += bias4[0]; f[2] += bias4[1]; f[3] += bias4[1];
                f[4] += bias4[2]; f[5] += bias4[2]; f[6] += bias4[3]; f[7] += bias4[3];
              }
              if (f_relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
              }
              uint32_t m[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 pr = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                m[i] = *reinterpret_cast<const uint32_t*>(&pr);
              }
              tc::stmatrix_x4_trans(srow + (uint32_t)k * 8u * 80u, m[0], m[1], m[2], m[3]);
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 t4 = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(my_stg) + lane * 80 + c * 16);
              wv[4 * c] = t4.x; wv[4 * c + 1] = t4.y; wv[4 * c + 2] = t4.z; wv[4 * c + 3] = t4.w;
            }
            __syncwarp();
          } else {
          float v[32];
          tc::tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * a.N + xs), v);
          tc::tmem_ld_wait();
          if (f_dbgf & 1) { float sacc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sacc += v[i];
            if (sacc == 1.2345e-30f) f_dbg[0] = 1;
            continue; }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] += bias_v;
            if (f_relu) v[i] = v[i] < 0.f ? 0.f : v[i];
          }
          {
            // fp32 NCHW: stage [channel = lane][32 px] (row stride 36 floats), re-read as 4 channels x 8 groups of
            // 4 pixels so every warp store writes four full 128 B lines
            float* stf = reinterpret_cast<float*>(my_stg);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(stf + lane * 36 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            __syncwarp();
            const int px4 = (lane & 7) * 4;
            float* dst0 = reinterpret_cast<float*>(a.out) + (long long)b * a.out_bs +
                          ((long long)(a.out_c_off + ch - lane) * a.out_Hp + oy) * a.out_Wp + (x0 + xs + px4);
            const long long cstride = (long long)a.out_Hp * a.out_Wp;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = k * 4 + (lane >> 3);
              const float4 val = *reinterpret_cast<const float4*>(stf + c * 36 + px4);
              float* dst = dst0 + c * cstride;
              if (px4 + 3 < nvalid && !f_acc && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                *reinterpret_cast<float4*>(dst) = val;
              } else {
                const float vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (px4 + i < nvalid) dst[i] = f_acc ? dst[i] + vv[i] : vv[i];
              }
            }
            __syncwarp();
            continue;
          }
          }
          if (f_dbgf & 2) { uint32_t xacc = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) xacc ^= wv[i];
            if (xacc == 0x12345678u) f_dbg[0] = 1;
            continue; }
          if (f_px) {
            // wv = [x-phase 0: chunk0, chunk0+1 | x-phase 1: chunk0, chunk0+1]; output pixels 2x and 2x+1 of a chunk are
            // 32 contiguous bytes: one 256-bit access per chunk for the mask, the accumulate target and the store
            if (lane < nvalid) {
              __nv_bfloat16* obase = obase_t + eoff;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint32_t v8[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { v8[i] = wv[4 * c + i]; v8[4 + i] = wv[8 + 4 * c + i]; }
                if (f_mask) {
                  uint32_t mk[8];
                  tc::ldg256_nc(mbase_t + eoff + c * cplane, mk);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v8[i] &= relu_keep(mk[i]);
                }
                if (f_acc) {
                  uint32_t old[8];
                  tc::ld256(obase + c * cplane, old);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old[i]));
                    const float2 nf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v8[i]));
                    const __nv_bfloat162 r = __floats2bfloat162_rn(of.x + nf.x, of.y + nf.y);
                    v8[i] = *reinterpret_cast<const uint32_t*>(&r);
                  }
                }
                tc::st256(obase + c * cplane, v8);
              }
            }
          } else if (lane < nvalid) {
            __nv_bfloat16* obase = obase_t + eoff;
            if (f_mask) {
              const __nv_bfloat16* mbase = mbase_t + eoff;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 mk = __ldg(reinterpret_cast<const uint4*>(mbase + c * cplane));
                wv[c * 4] &= relu_keep(mk.x); wv[c * 4 + 1] &= relu_keep(mk.y);
                wv[c * 4 + 2] &= relu_keep(mk.z); wv[c * 4 + 3] &= relu_keep(mk.w);
              }
            }
            if (f_acc) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 old = *reinterpret_cast<const uint4*>(obase + c * cplane);
                const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 of = __bfloat1622float2(o2[i]);
                  const float2 nf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[c * 4 + i]));
                  const __nv_bfloat162 r = __floats2bfloat162_rn(of.x + nf.x, of.y + nf.y);
                  wv[c * 4 + i] = *reinterpret_cast<const uint32_t*>(&r);
                }
              }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(obase + c * cplane) =
                  make_uint4(wv[c * 4], wv[c * 4 + 1], wv[c * 4 + 2], wv[c * 4 + 3]);
            const int sc0 = chunk0 - a.out_c_off - a.s2d_c0;     // WIDE: only the quadrants holding chunks s2d_c0 .. +7
            const int ox = (x0 + xs + lane) * a.sx + a.ax;
            if (f_s2d && sc0 >= 0 && sc0 < 8 && (oy >> 1) >= a.s2d_band && (oy >> 1) < (a.H >> 1) - a.s2d_band &&
                (ox >> 1) >= a.s2d_band && (ox >> 1) < (a.W >> 1) - a.s2d_band) {
              // phase plane p = (y&1)*2 + (x&1) holds pixel (y/2, x/2): the input layout of the next stride-2 layer
              // (and, with a band, of the polyphase resize-convolution's backward: up_poly.cu)
              const int H2p = (a.H >> 1) + 4, W2p = (a.W >> 1) + 4;
              const int ph = (oy & 1) * 2 + (ox & 1);
              __nv_bfloat16* sb = a.s2d + (long long)b * a.s2d_bs +
                                  (((long long)(ph * 8 + sc0) * H2p + (oy >> 1) + 2) * W2p + (ox >> 1) + 2) * 8;
              const long long splane = (long long)H2p * W2p * 8;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(sb + c * splane) = make_uint4(wv[c * 4], wv[c * 4 + 1], wv[c * 4 + 2], wv[c * 4 + 3]);
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty + abuf);
      if (f_dbg && ew == 0 && lane == 0) f_dbg[blockIdx.x * 8 + 5] += clock64() - te0;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if ((CL && a_cluster > 1)) tc::cluster_sync();   // nobody exits while the partner may still multicast into its shared memory
  if (warp == 3) { tc::fence_after_sync(); tc::tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// weight packing: torch fp32 [Cout][Cin][k][k] -> bf16 [K block][position][k8 (2)][m (128)][8]
//   m = 64*g + n.  PAIR: group g uses tap g of the position, output channel co_off + n.
//                  WIDE: both groups use tap 0, output channel co_off + 64*g + n.
//   transposed (dgrad): "output channel" indexes the conv's INPUT channels and k its OUTPUT channels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack2_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int transposed, int wide, int pxpair, int co_off,
             __nv_bfloat16* __restrict__ wpk, const __grid_constant__ cnp_c2_plan plan, int total_pos,
             long long group_stride /* elements between the weight tensors of lane groups 0 and 1 (0: the same tensor) */,
             long long wsel_stride /* elements between the stacked weight tensors selected per K block (plan.kb_wsel) */) {
  // (K block, local position) of every global position, built once per block in shared memory: indexing the by-value
  // plan with run-time subscripts makes every thread keep a private copy of the 2 KB struct (the 256-position plan of the
  // phase dgrad took 124 us to pack that way)
  __shared__ unsigned char pos_kb[C2_MAX_KB * C2_MAX_POS];
  __shared__ unsigned char pos_lp[C2_MAX_KB * C2_MAX_POS];
  if (threadIdx.x == 0) {
    int g = 0;
    for (int kb = 0; kb < plan.n_kb; ++kb)
      for (int lp = 0; lp < plan.t_npos[plan.kb_type[kb]]; ++lp, ++g) { pos_kb[g] = (unsigned char)kb; pos_lp[g] = (unsigned char)lp; }
  }
  __syncthreads();
  const long long total = (long long)total_pos * 2048;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int c = (int)(e & 7), m = (int)((e >> 3) & 127), k8 = (int)((e >> 10) & 1);
    const int kb = pos_kb[e >> 11], gp = pos_lp[e >> 11];
    const int type = plan.kb_type[kb];
    // PAIR: m = 64 g + n.  x-phase pairs: every 32-lane quadrant holds [16 channels of group 0 | the same of group 1], so
    // that an epilogue thread owns both x-phases of its pixel (conv_tc2_kernel)
    const int g = pxpair ? (m >> 4) & 1 : m >> 6, n = pxpair ? (m >> 5) * 16 + (m & 15) : m & 63;
    const int ky = plan.t_tap[type][gp][wide ? 0 : 2 * g], kx = plan.t_tap[type][gp][wide ? 1 : 2 * g + 1];
    float v = 0.f;
    if (ky >= 0) {
      const int kc = plan.kb_wci0[kb] + k8 * 8 + c;
      const int nn = co_off + (wide ? m : n);
      if (!transposed) {
        if (nn < Cout && kc < Cin) v = w[(size_t)g * group_stride + (((size_t)nn * Cin + kc) * k + ky) * k + kx];
      } else {
        if (kc < Cout && nn < Cin) v = w[(size_t)plan.kb_wsel[kb] * wsel_stride + (((size_t)kc * Cin + nn) * k + ky) * k + kx];
      }
    }
    wpk[e] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------
// host-side plan construction
// ---------------------------------------------------------------------------------------------
enum { KIND_K5S1 = 0, KIND_K1 = 1, KIND_K5S2 = 2, KIND_K5S1_DGRAD = 3, KIND_K1_DGRAD = 4, KIND_K5S2_DGRAD = 5,
       KIND_UP_PHASE = 6, KIND_UP_PHASE_DGRAD = 7 };

struct Tap { int roff, coff, wky, wkx; };

// positions of one K-block type from the tap list of group 0 (window offsets relative to the output pixel's
// padded origin); PAIR adds the row-shifted copies for group 1
// (taps1 != NULL: group 1 has its own tap list at the SAME window offsets -- the second x-phase of a stride-2 dgrad)
int add_type(cnp_c2_plan* p, int type, const Tap* taps, int ntaps, int pitch, int wide, const Tap* taps1 = nullptr,
             int ntaps1 = 0) {
  int np = 0;
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) {
      int t0 = -1, t1 = -1;
      for (int t = 0; t < ntaps; ++t) {
        if (taps[t].roff == r && taps[t].coff == c) t0 = t;
        if (!wide && !taps1 && taps[t].roff + 1 == r && taps[t].coff == c) t1 = t;
      }
      for (int t = 0; t < ntaps1; ++t)
        if (taps1[t].roff == r && taps1[t].coff == c) t1 = t;
      if (t0 < 0 && t1 < 0) continue;
      CNP_REQUIRE(np < C2_MAX_POS, "conv plan: too many positions");
      const Tap* g1 = taps1 ? taps1 : taps;
      p->t_boff[type][np] = (short)(r * pitch + c);
      p->t_tap[type][np][0] = t0 >= 0 ? (signed char)taps[t0].wky : -1;
      p->t_tap[type][np][1] = t0 >= 0 ? (signed char)taps[t0].wkx : -1;
      p->t_tap[type][np][2] = t1 >= 0 ? (signed char)g1[t1].wky : -1;
      p->t_tap[type][np][3] = t1 >= 0 ? (signed char)g1[t1].wkx : -1;
      ++np;
    }
  p->t_npos[type] = np;
  return 0;
}

int build_plan2(int kind, int n_chunks, int pitch, int py, int px, int wide, cnp_c2_plan* p) {
  memset(p, 0, sizeof(*p));
  Tap taps[25];
  auto add_kb = [&](int chunk0, int wci0, int type) {
    p->kb_chunk0[p->n_kb] = chunk0; p->kb_wci0[p->n_kb] = wci0; p->kb_type[p->n_kb] = type; ++p->n_kb;
  };
  if (kind == KIND_K5S1 || kind == KIND_K5S1_DGRAD) {
    CNP_REQUIRE(n_chunks >= 2 && n_chunks <= 16 && n_chunks % 2 == 0, "conv plan: 5x5 needs 2, 4, .. 16 source chunks");
    int nt = 0;
    for (int ky = 0; ky < 5; ++ky)
      for (int kx = 0; kx < 5; ++kx)
        taps[nt++] = Tap{ky, kx, kind == KIND_K5S1 ? ky : 4 - ky, kind == KIND_K5S1 ? kx : 4 - kx};
    if (int e = add_type(p, 0, taps, nt, pitch, wide)) return e;
    for (int g = 0; g < n_chunks / 2; ++g) add_kb(2 * g, 16 * g, 0);
    p->regular = 1;
  } else if (kind == KIND_K1 || kind == KIND_K1_DGRAD) {
    CNP_REQUIRE(n_chunks == 8 || n_chunks == 16, "conv plan: 1x1 needs 8 or 16 source chunks");
    taps[0] = Tap{2, 2, 0, 0};
    if (int e = add_type(p, 0, taps, 1, pitch, wide)) return e;
    for (int g = 0; g < n_chunks / 2; ++g) add_kb(2 * g, 16 * g, 0);
  } else if (kind == KIND_K5S2) {
    CNP_REQUIRE(n_chunks == 32, "conv plan: stride-2 forward reads the 4x8-chunk phase tensor");
    for (int ph = 0; ph < 4; ++ph) {
      const int qy = ph >> 1, qx = ph & 1;
      int nt = 0;
      for (int ky = qy; ky < 5; ky += 2)
        for (int kx = qx; kx < 5; kx += 2)
          taps[nt++] = Tap{2 + (ky - 2 - qy) / 2, 2 + (kx - 2 - qx) / 2, ky, kx};
      if (int e = add_type(p, ph, taps, nt, pitch, wide)) return e;
      for (int g = 0; g < 4; ++g) add_kb(ph * 8 + 2 * g, 16 * g, ph);
    }
  } else if (kind == KIND_K5S2_DGRAD) {
    CNP_REQUIRE(n_chunks == 8, "conv plan: stride-2 dgrad reads the 8-chunk dy tensor");
    int nt = 0;
    if (px == 2) {
      // both x-phases of output row phase py in one launch: lane group 0 = phase px 0, group 1 = phase px 1, each with
      // its own taps at the shared window offsets (3 x 3 resp. 2 x 3 positions; x-phase 1 has no tap at column offset 1)
      CNP_REQUIRE(!wide, "conv plan: the x-phase pair produces 64 channels per phase");
      Tap taps1[25];
      int nt1 = 0;
      for (int ky = py; ky < 5; ky += 2) {
        for (int kx = 0; kx < 5; kx += 2) taps[nt++] = Tap{2 + (py + 2 - ky) / 2, 2 + (0 + 2 - kx) / 2, ky, kx};
        for (int kx = 1; kx < 5; kx += 2) taps1[nt1++] = Tap{2 + (py + 2 - ky) / 2, 2 + (1 + 2 - kx) / 2, ky, kx};
      }
      if (int e = add_type(p, 0, taps, nt, pitch, 0, taps1, nt1)) return e;
    } else {
      for (int ky = py; ky < 5; ky += 2)
        for (int kx = px; kx < 5; kx += 2)
          taps[nt++] = Tap{2 + (py + 2 - ky) / 2, 2 + (px + 2 - kx) / 2, ky, kx};
      if (int e = add_type(p, 0, taps, nt, pitch, wide)) return e;
    }
    for (int g = 0; g < 4; ++g) add_kb(2 * g, 16 * g, 0);
  } else if (kind == KIND_UP_PHASE) {
    // Polyphase form of conv5x5(bilinear_up2x(x)) (tools/polyphase_check.py): output row phase py of the high-resolution
    // result as a 4x4-tap convolution of the REPLICATE-padded low-resolution input; lane group g = output x-phase g.
    // Window offset of phase tap (p, q) of x-phase b: (py + p, b + q) -- 4 rows x 5 column offsets = 20 positions per
    // 16-channel K block (x-phase 0 has no tap at column offset 4, x-phase 1 none at 0).  The weights are the 4x4 phase
    // weights [b][Cout][Cin][4][4] of row phase py (cnp_up_phase_weights).
    CNP_REQUIRE(n_chunks >= 2 && n_chunks <= 16 && n_chunks % 2 == 0 && !wide && (py == 0 || py == 1),
                "conv plan: the up-phase kind needs 2..16 source chunks, 64 outputs per x-phase, py in {0,1}");
    Tap taps1[25];
    int nt = 0, nt1 = 0;
    for (int pp = 0; pp < 4; ++pp)
      for (int qq = 0; qq < 4; ++qq) {
        taps[nt++] = Tap{py + pp, qq, pp, qq};
        taps1[nt1++] = Tap{py + pp, 1 + qq, pp, qq};
      }
    if (int e = add_type(p, 0, taps, nt, pitch, 0, taps1, nt1)) return e;
    for (int g = 0; g < n_chunks / 2; ++g) add_kb(2 * g, 16 * g, 0);
  } else if (kind == KIND_UP_PHASE_DGRAD) {
    // Input gradient of the polyphase resize-convolution at LOW resolution: the source is the space-to-depth copy of dY
    // (chunk = (a*2+b)*8 + c holds dY[2i+a, 2j+b]), the reduction runs over the 4 phases x 64 channels, and phase (a, b)
    // contributes its 4x4 taps (p, q) at window offsets (4-a-p, 4-b-q) of a 5x5 window:
    //   dx[r, c] = sum_{a,b,p,q} Wp[a,b,:,ci,p,q] . dY_ab[r + 2 - a - p, c + 2 - b - q]
    // 16 K blocks x 16 positions = 256 tap-GEMMs per low-res pixel against 4 x 100 for dgrad + upsample^T.  The weights are
    // the four phase tensors [a][b][Cout][Cin][4][4] (cnp_up_phase_weights), selected per K block.
    CNP_REQUIRE(n_chunks == 32 && wide, "conv plan: the up-phase dgrad reads the 4x8-chunk phase tensor of dY and produces 128 channels");
    for (int ph = 0; ph < 4; ++ph) {
      const int a = ph >> 1, b = ph & 1;
      int nt = 0;
      for (int pp = 0; pp < 4; ++pp)
        for (int qq = 0; qq < 4; ++qq) taps[nt++] = Tap{4 - a - pp, 4 - b - qq, pp, qq};
      if (int e = add_type(p, ph, taps, nt, pitch, 1)) return e;     // one tap per position for both lane groups
      for (int g = 0; g < 4; ++g) { add_kb(ph * 8 + 2 * g, 16 * g, ph); p->kb_wsel[p->n_kb - 1] = ph; }
    }
  } else {
    CNP_REQUIRE(false, "conv plan: unknown kind %d", kind);
  }
  return 0;
}

int plan_total_pos(const cnp_c2_plan& p) {
  int t = 0;
  for (int kb = 0; kb < p.n_kb; ++kb) t += p.t_npos[p.kb_type[kb]];
  return t;
}

int g2_num_sms = 0;
int num_sms2() {
  if (g2_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g2_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g2_num_sms <= 0) g2_num_sms = 148;
  }
  return g2_num_sms;
}

size_t c2_smem_bytes(int plane_sm, int out_mode, int epi_warps) {
  const size_t stg = (size_t)epi_warps * (out_mode == 1 ? C2_STG_WORDS_F32 : C2_STG_WORDS_BF16) * 4;
  return (size_t)C2_ABUFS * 2 * plane_sm * 16 + (size_t)C2_WSTAGES * C2_STAGE_BYTES + stg + 16 * 8;
}

// tile geometry: N (pixels per MMA, multiple of 32, <= 256), nacc = floor(512/N) accumulators of rpa rows.
// Cost model (cycles per tile), calibrated with cnp_conv_tc2_debug on B200:
//   MMA   : an M128 x N x K16 MMA takes max(N/2, (4096 + 32 N)/128) cycles -- tensor floor vs the 128 B/clk
//           shared-memory operand port (tools/mma_rate.cu) -- and a tile issues mma_per_acc per accumulator;
//   L2    : weights (mma_per_acc x 4 KB) + windows are re-read per tile; ~30 B/clk/SM when all SMs stream;
//   epilogue (not overlapped with the MMAs): ~18 cycles per accumulator column.
// Minimise waves x (max(MMA, L2) + epilogue).
// Two accumulator sets (nbuf = 2, nacc * N <= 256 columns each): the epilogue of a tile overlaps the MMAs of the next one,
// a tile then costs max(main loop, epilogue) instead of their sum -- pays on the layers with a short K loop or a heavy
// epilogue (the folded first layer: 30 positions; the stride-2 input gradients: 40 / 60 positions and a masked
// read-modify-write epilogue), not on the 128-channel layers, whose weight stream needs 3 accumulators per tile to stay
// under the L2 rate.  epi_col: epilogue cycles per accumulator column (18 plain; the mask and the accumulate target are
// global loads issued from the epilogue).
void choose_geometry(cnp_c2_args* a, int mma_per_acc, double epi_col, bool allow_nbuf2, int epi_warps) {
  const int rpa = (a->wide || a->pxpair) ? 1 : 2;
  double best = 1e300;
  for (int N = 32; N <= 256; N += 32) {
    const int tw_max = N < a->W ? N : a->W;
    const int tiles_x = cnp_cdiv(a->W, tw_max);
    const int TW = cnp_cdiv(a->W, tiles_x);              // balanced tile width
    if (cnp_cdiv(TW, 32) * 32 != N) continue;            // a smaller N serves this width
    const double cyc = (N / 2.0 > 32.0 + N / 4.0) ? N / 2.0 : 32.0 + N / 4.0;
    int nacc = 512 / N;
    const int max_rows = cnp_cdiv(a->H, rpa);
    if (nacc > max_rows) nacc = max_rows;
    for (; nacc >= 1; --nacc) {
      const int TH = nacc * rpa;
      const int pitch = TW + 4;
      const int plane_sm = (TH + 4) * pitch + (N > pitch ? N - pitch : 0) + 8;
      if (c2_smem_bytes(plane_sm, a->out_mode, epi_warps) > 225 * 1024) continue;
      const int tiles_y = cnp_cdiv(a->H, TH);
      const long long tiles = (long long)a->B * tiles_x * tiles_y;
      const long long waves = (tiles + num_sms2() - 1) / num_sms2();
      const int n_kb = mma_per_acc / (a->plan.t_npos[0] > 0 ? a->plan.t_npos[0] : 1);
      const double mma_c = (double)mma_per_acc * nacc * cyc;
      // weights are shared by the CTA pair when launched as clusters of 2 (multicast): half the bytes per CTA
      const double l2_c = ((double)mma_per_acc * 4096.0 + (double)n_kb * 2.0 * (TH + 4) * pitch * 16.0) / 30.0;
      const double main_c = (mma_c > l2_c ? mma_c : l2_c) + 1500.0, epi_c = epi_col * nacc * N;
      for (int nbuf = 1; nbuf <= ((allow_nbuf2 && nacc * N <= 256) ? 2 : 1); ++nbuf) {
        const double cost = nbuf == 1 ? (double)waves * (main_c + epi_c)
                                      : (double)waves * (main_c > epi_c ? main_c : epi_c) + (main_c < epi_c ? main_c : epi_c);
        if (cost < best) {
          best = cost;
          a->N = N; a->nacc = nacc; a->rpa = rpa; a->TW = TW; a->TH = TH; a->pitch = pitch; a->plane_sm = plane_sm;
          a->tiles_x = tiles_x; a->tiles_y = tiles_y; a->nbuf = nbuf;
        }
      }
    }
  }
}

}  // namespace

struct cnp_conv_out {
  int mode;
  cnp_blk blk;
  float* f32; long long f32_bstride; int f32_ch_off;
  int sy, ay, sx, ax;
  const float* bias;
  int relu;
  const cnp_blk* mask;
  int accumulate;
  const cnp_blk* s2d;    // optional: also write the space-to-depth copy of the output (32 chunks, half resolution)
  int s2d_c0;            // first of the 8 output chunks that are copied (0; 8 = second half of a 128-channel output)
  int s2d_band;          // half-res pixels along every edge that are NOT written (stay zero)
};

// Debug aid: when buf != NULL every later cnp_conv_tc2 launch writes, per CTA, 8 int64 counters
// {MMA-thread cycles, waiting on the epilogue, on windows, on weights, tiles, epilogue cycles, -, -}.
CNP_API int cnp_conv_tc2_debug(long long* buf, int flags) { g_c2_dbg = buf; g_c2_dbg_flags = flags; return 0; }

// 1 (default): plain launch.  2: clusters of two CTAs, each loading half of every weight stage and multicasting it
// to its partner (halves the L2 -> SM weight traffic; measured neutral at the current tile sizes, kept for smaller
// tiles).  Process-wide switch.
CNP_API int cnp_conv_tc2_set_cluster(int cluster) {
  CNP_REQUIRE(cluster == 1 || cluster == 2, "conv_tc2_set_cluster: 1 or 2");
  g_c2_cluster = cluster;
  return 0;
}

CNP_API long long cnp_conv_tc2_packed_bytes(int kind, int n_chunks, int n_out) {
  cnp_c2_plan p;
  if (build_plan2(kind, n_chunks, 8, 0, 0, n_out == 128, &p)) return -1;
  return (long long)plan_total_pos(p) * C2_POS_BYTES;
}

// Pack torch-layout fp32 weights [Cout][Cin][k][k] for cnp_conv_tc2 (n_out = 64: PAIR, 128: WIDE).
CNP_API int cnp_conv_tc2_pack(const float* w, int Cout, int Cin, int k, int kind, int n_chunks, int py, int px,
                              int co_off, int n_out, void* wpk, cudaStream_t st) {
  CNP_REQUIRE(n_out == 64 || n_out == 128, "conv_tc2_pack: n_out must be 64 or 128");
  cnp_c2_plan p;
  if (int e = build_plan2(kind, n_chunks, 8, py, px, n_out == 128, &p)) return e;
  const int total_pos = plan_total_pos(p);
  CNP_REQUIRE((kind != KIND_UP_PHASE && kind != KIND_UP_PHASE_DGRAD) || k == 4,
              "conv_tc2_pack: the up-phase kinds pack 4x4 phase weights ([2][Cout][Cin][4][4] of one row phase / all four)");
  const int transposed = ((kind >= KIND_K5S1_DGRAD && kind <= KIND_K5S2_DGRAD) || kind == KIND_UP_PHASE_DGRAD) ? 1 : 0;
  const int pack_blocks = total_pos * 2048 / (256 * 4) < 128 ? 128 : (total_pos * 2048 / (256 * 4) > 592 ? 592 : total_pos * 2048 / (256 * 4));
  const int pxpair = ((kind == KIND_K5S2_DGRAD && px == 2) || kind == KIND_UP_PHASE) ? 1 : 0;
  pack2_kernel<<<pack_blocks, 256, 0, st>>>(w, Cout, Cin, k, transposed, n_out == 128, pxpair, co_off,
                                    reinterpret_cast<__nv_bfloat16*>(wpk), p, total_pos,
                                    kind == KIND_UP_PHASE ? (long long)Cout * Cin * 16 : 0ll,
                                    kind == KIND_UP_PHASE_DGRAD ? (long long)Cout * Cin * 16 : 0ll);
  CNP_LAUNCH_CHECK("pack2_kernel");
  return 0;
}

// Tensor-core convolution producing n_out (64 or 128) output channels at chunks [out.cb_off, +n_out/8).
// x: blocked source whose chunks [x->cb_off, x->cb_off + n_chunks) are the reduction dimension;
// (x->H, x->W) is the accumulator grid.
static int conv_tc2_launch(const cnp_blk* x, int n_chunks, const void* wpk, const void* wpk2, int w2_from_b, int kind, int py,
                           int px, int n_out, const cnp_conv_out* o, int B, cudaStream_t st);

CNP_API int cnp_conv_tc2(const cnp_blk* x, int n_chunks, const void* wpk, int kind, int py, int px, int n_out,
                         const cnp_conv_out* o, int B, cudaStream_t st) {
  return conv_tc2_launch(x, n_chunks, wpk, nullptr, 0, kind, py, px, n_out, o, B, st);
}

// Same launch with TWO packed weight tensors: images b < w2_from_b use wpk, the others wpk2 (same kind and geometry).
// One launch then serves the row strips and the tap-transposed column strips of a square level (up_poly.cu).
CNP_API int cnp_conv_tc2_w2(const cnp_blk* x, int n_chunks, const void* wpk, const void* wpk2, int w2_from_b, int kind,
                            int py, int px, int n_out, const cnp_conv_out* o, int B, cudaStream_t st) {
  CNP_REQUIRE(wpk2 && w2_from_b > 0 && w2_from_b < B, "conv_tc2_w2: needs a second weight tensor and 0 < w2_from_b < B");
  return conv_tc2_launch(x, n_chunks, wpk, wpk2, w2_from_b, kind, py, px, n_out, o, B, st);
}

static int conv_tc2_launch(const cnp_blk* x, int n_chunks, const void* wpk, const void* wpk2, int w2_from_b, int kind, int py,
                           int px, int n_out, const cnp_conv_out* o, int B, cudaStream_t st) {
  CNP_REQUIRE(x && o && wpk && B > 0, "conv_tc2: bad arguments");
  CNP_REQUIRE(n_out == 64 || n_out == 128, "conv_tc2: n_out must be 64 or 128");
  cnp_c2_args a;
  memset(&a, 0, sizeof(a));
  a.H = x->H; a.W = x->W; a.B = B;
  a.x_Hp = x->H + 4; a.x_Wp = x->W + 4;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x->base) + (long long)x->cb_off * a.x_Hp * a.x_Wp * 8;
  a.x_bs = x->bstride;
  a.w = reinterpret_cast<const uint8_t*>(wpk);
  a.w2 = reinterpret_cast<const uint8_t*>(wpk2); a.w2_from_b = w2_from_b;
  a.wide = n_out == 128;
  a.pxpair = ((kind == KIND_K5S2_DGRAD && px == 2) || kind == KIND_UP_PHASE) ? 1 : 0;
  CNP_REQUIRE(!a.pxpair || (o->mode == 0 && o->sx == 2 && o->ax == 0 && !o->s2d),
              "conv_tc2: the x-phase pair writes a blocked output with sx = 2, ax = 0");
  a.out_mode = o->mode;
  a.cluster = wpk2 ? 1 : g_c2_cluster;    // the multicast weight stream assumes one tensor for the CTA pair
  if (int e = build_plan2(kind, n_chunks, 8, py, px, a.wide, &a.plan)) return e;   // position count only
  // Opt-in (CNP_DOUBLE_ACC=1): measured on B200 it LOSES wherever the model picks it -- masked WIDE dgrad at 304^2 503 ->
  // 743 us, stride-2 dgrad phases 127 -> 373 us / 45 -> 155 us -- because a set of <= 256 columns means one accumulator of
  // N = 160, and every tile re-streams the whole packed weight tensor from L2 for a third of the MMAs: the weight stream,
  // not the epilogue, bounds short tiles (same finding as the two-round tail split above).  CNP_DOUBLE_ACC=narrow restricts it to the folded first layer
  // (one K block: N = 64 x 4 accumulators x 2 sets is chosen, 137 against 127 us).
  // epilogue instantiation (flag set known from the arguments) and, with it, the number of epilogue warps
  static const bool generic_only = getenv("CNP_C2_GENERIC_EPILOGUE") != nullptr;
  int epi = -1;
  if (!generic_only && a.cluster == 1 && !g_c2_dbg && g_c2_dbg_flags == 0 && o->mode == 0 && (o->bias != nullptr) == (o->relu != 0)) {
    epi = (o->mask ? EF_MASK : 0) | (o->accumulate ? EF_ACC : 0) | (a.pxpair ? EF_PX : 0) | (a.wide ? EF_WIDE : 0) |
          (o->bias ? EF_BIASRELU : 0) | (o->s2d ? EF_S2D : 0);
    static const int supported[] = {EF_BIASRELU, EF_BIASRELU | EF_S2D, EF_PX | EF_BIASRELU, EF_MASK | EF_WIDE,
                                    EF_MASK | EF_WIDE | EF_S2D, EF_WIDE, 0, EF_MASK, EF_PX | EF_MASK | EF_ACC, EF_MASK | EF_ACC};
    bool ok = false;
    for (int f : supported) ok = ok || f == epi;
    if (!ok) epi = -1;
  }
  const int epi_warps = epi >= 0 ? C2_EPI_WARPS_SPEC : C2_EPI_WARPS;
  static const char* nbuf2_env = getenv("CNP_DOUBLE_ACC");
  const bool nbuf2 = nbuf2_env != nullptr && (strcmp(nbuf2_env, "narrow") != 0 || a.plan.n_kb == 1);
  choose_geometry(&a, plan_total_pos(a.plan), 18.0 + (o->mode == 0 && o->mask ? 30.0 : 0.0) + (o->accumulate ? 40.0 : 0.0),
                  nbuf2, epi_warps);
  if (getenv("CNP_C2_VERBOSE"))
    fprintf(stderr, "conv_tc2 kind %d chunks %d %dx%d: N %d nacc %d nbuf %d tiles %d\n", kind, n_chunks, a.H, a.W, a.N, a.nacc,
            a.nbuf, B * a.tiles_x * a.tiles_y);
  CNP_REQUIRE(a.N > 0, "conv_tc2: no tile geometry for %d x %d", a.H, a.W);
  if (int e = build_plan2(kind, n_chunks, a.pitch, py, px, a.wide, &a.plan)) return e;
  a.out_mode = o->mode;
  a.sy = o->sy; a.ay = o->ay; a.sx = o->sx; a.ax = o->ax;
  CNP_REQUIRE(a.sy >= 1 && a.sx >= 1, "conv_tc2: output scale must be >= 1");
  if (o->mode == 0) {
    CNP_REQUIRE(o->blk.H == x->H * o->sy && o->blk.W == x->W * o->sx, "conv_tc2: output geometry mismatch");
    a.out = o->blk.base; a.out_bs = o->blk.bstride; a.out_c_off = o->blk.cb_off;
    a.out_Hp = o->blk.H + 4; a.out_Wp = o->blk.W + 4;
    if (o->s2d) {
      CNP_REQUIRE(o->sy == 1 && o->sx == 1 && o->blk.H % 2 == 0 && o->blk.W % 2 == 0 &&
                  o->s2d->H == o->blk.H / 2 && o->s2d->W == o->blk.W / 2 && o->s2d->cb_off == 0 && o->s2d_band >= 0 &&
                  (o->s2d_c0 == 0 || (n_out == 128 && o->s2d_c0 == 8)),
                  "conv_tc2: space-to-depth output geometry mismatch");
      a.s2d = reinterpret_cast<__nv_bfloat16*>(o->s2d->base); a.s2d_bs = o->s2d->bstride;
      a.s2d_c0 = o->s2d_c0; a.s2d_band = o->s2d_band;
    }
    if (o->mask) {
      CNP_REQUIRE(o->mask->H == o->blk.H && o->mask->W == o->blk.W, "conv_tc2: mask geometry mismatch");
      a.mask = reinterpret_cast<const __nv_bfloat16*>(o->mask->base);
      a.mask_bs = o->mask->bstride; a.mask_cb_off = o->mask->cb_off;
    }
  } else {
    CNP_REQUIRE(o->f32 && o->sy == 1 && o->sx == 1 && !o->mask, "conv_tc2: fp32 NCHW output is plain only");
    a.out = o->f32; a.out_bs = o->f32_bstride; a.out_c_off = o->f32_ch_off;
    a.out_Hp = x->H; a.out_Wp = x->W;
  }
  a.bias = o->bias; a.relu = o->relu; a.accumulate = o->accumulate;
  a.dbg = g_c2_dbg; a.dbg_flags = g_c2_dbg_flags;
  const size_t smem = c2_smem_bytes(a.plane_sm, a.out_mode, epi_warps);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaSuccess;
#define C2_ATTR(CLF, F, EWN) if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc2_kernel<CLF, (F), EWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#define C2_ATTR_S(F) C2_ATTR(false, F, C2_EPI_WARPS_SPEC)
    C2_ATTR(false, -1, C2_EPI_WARPS) C2_ATTR(true, -1, C2_EPI_WARPS) C2_ATTR_S(EF_BIASRELU) C2_ATTR_S(EF_BIASRELU | EF_S2D)
    C2_ATTR_S(EF_PX | EF_BIASRELU) C2_ATTR_S(EF_MASK | EF_WIDE) C2_ATTR_S(EF_MASK | EF_WIDE | EF_S2D)
    C2_ATTR_S(EF_WIDE) C2_ATTR_S(0) C2_ATTR_S(EF_MASK) C2_ATTR_S(EF_PX | EF_MASK | EF_ACC) C2_ATTR_S(EF_MASK | EF_ACC)
#undef C2_ATTR_S
#undef C2_ATTR
    if (e != cudaSuccess) { cnp_set_error("conv_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr = smem;
  }
  const int ntiles = B * a.tiles_x * a.tiles_y;
  int grid = ntiles < num_sms2() ? ntiles : num_sms2();
  if (grid < 2) a.cluster = 1;
  if (a.cluster == 2) grid &= ~1;
  // tail splitting: the tiles left over after the last full round are cut into single-accumulator items so that the
  // last round costs 1/nacc of a tile (1632 tiles on 148 SMs: 11.33 rounds instead of 12)
  a.n_work = ntiles; a.split_from = ntiles; a.split = 1;
  {
    // (measured: letting the split items take two rounds -- 76 left-over tiles of 3 accumulators -> 228 items -- is SLOWER
    // than one round of whole tiles, 2.86 vs 2.69 ms per step: a single-accumulator item still streams the whole weight
    // tensor, so it is L2-bound and costs far more than a third of a tile)
    const int full = (ntiles / grid) * grid, left = ntiles - full;
    if (left > 0 && a.nacc > 1 && left * a.nacc <= grid && full > 0) {
      a.split_from = full; a.split = a.nacc; a.n_work = full + left * a.nacc;
    }
  }
  // CTA pairs (cluster of 2) share the weight stream by multicast; needs an even grid (a CTA left without a tile in
  // the last round keeps streaming its half)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C2_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = a.cluster; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  if (a.cluster > 1) {
    cudaError_t le = cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, -1, C2_EPI_WARPS>, a);
    if (le != cudaSuccess) { cnp_set_error("conv_tc2_kernel: %s", cudaGetErrorString(le)); return (int)le; }
  } else {   // plain launch when no cluster is requested (cluster-attribute launches place CTAs differently)
#define C2_LAUNCH_EPI(F) case (F): conv_tc2_kernel<false, (F), C2_EPI_WARPS_SPEC><<<grid, 32 * (4 + C2_EPI_WARPS_SPEC), smem, st>>>(a); break;
    switch (epi) {
      C2_LAUNCH_EPI(EF_BIASRELU)                          // forward layers, strips
      C2_LAUNCH_EPI(EF_BIASRELU | EF_S2D)                 // forward + space-to-depth copy for the stride-2 layer behind it
      C2_LAUNCH_EPI(EF_PX | EF_BIASRELU)                  // polyphase forward
      C2_LAUNCH_EPI(EF_MASK | EF_WIDE)                    // input gradient of a 128-channel level
      C2_LAUNCH_EPI(EF_MASK | EF_WIDE | EF_S2D)           // ... that also hands dY to a polyphase level
      C2_LAUNCH_EPI(EF_WIDE)                              // ... unmasked (before Upsample^T; strips)
      C2_LAUNCH_EPI(0)                                    // 64-channel input gradient, unmasked
      C2_LAUNCH_EPI(EF_MASK)
      C2_LAUNCH_EPI(EF_PX | EF_MASK | EF_ACC)             // stride-2 input gradient onto the skip gradient
      C2_LAUNCH_EPI(EF_MASK | EF_ACC)                     // stride-1 down-path input gradient onto the skip gradient
      default: conv_tc2_kernel<false, -1, C2_EPI_WARPS><<<grid, C2_THREADS, smem, st>>>(a); break;
    }
#undef C2_LAUNCH_EPI
  }
  CNP_LAUNCH_CHECK("conv_tc2_kernel");
  return 0;
}
