// RoIAlign backward on the tensor cores (bf16 gradients, [R,C,7,7] layout) — sm_100a.
//
//   grad_in[pix, c] = sum_{r touching pix} sum_bin Wt_r[pix, bin] * g[r, c, bin]
//
// Gather formulation, no atomics: CTA = (image, 16x16 pixel tile, 256 channels) owns its slice of
// grad_input as TWO fp32 accumulators D[128 ch x 256 px] in TMEM (512 columns) and walks the RoIs that
// touch the tile in index order (deterministic).  Per RoI one K = 64 (49 bins, zero padded) contraction:
//   A = g[r, c0:c0+128, :]   [128 ch x 64 bins]  K-major  (bulk-copied raw, re-laid out to the 128B swizzle)
//   B = Wt_r^T               [256 px x 64 bins]  K-major  (built from the prep kernel's Wy/Wx tables)
// Roles: warp 0 = bulk-copy producer (gradient chunks + table slices, 3-deep ring), warp 1 = MMA issuer,
// warps 2..17 = operand builders (double-buffered operand tiles) and, at the end, the epilogue that streams
// TMEM to grad_input NHWC (lanes = consecutive channels: 128 B coalesced stores, every element written once).
// The CUDA-core kernel (roi_align.cu) measured 8 % of HBM peak, issue bound (profiles/r01_roi_align_ncu.md).
//
// kRHWC = gradients in [R,7,7,C] order (DA_ROI_OUT_RHWC): the A operand is then MN-major AS IT LIES IN MEMORY.  ONE rank-4 TMA box
// (64 ch, 16 * ksteps bins, 4 channel groups, 1 RoI) per pair lands 128B-swizzled in a 4-slot operand ring (a box per k-step
// measured slower: the producer warp's wait -> expect_tx -> issue sequences are ~400 cycles each and serialise): no raw
// gradient ring and no relayout by the builders (the relayout's reads of 98-byte-pitch rows and its swizzled stores were ~690 of
// the ~1600 shared-memory-port cycles per (RoI, tile) pair; profiles/r02_roi_bwd_pipeline_study.md).  Bin-major order also lets a
// pair fetch and multiply only the BIN ROWS whose support meets the tile's rows (the prep kernel leaves that range in the pad
// column of the Wy table): a RoI spanning several tiles vertically costs each of them 1-3 k-steps instead of 4 (bins past 49
// are out-of-bounds zero fill, bins of the range's last k-step that do not meet the tile multiply zero weights).
#include <stdlib.h>
#include <string.h>
#include "da_common.cuh"
#include "da_ptx.cuh"
#include "roi_common.cuh"

namespace da {

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

__constant__ unsigned char kRowPerm[64] = {0, 1, 34, 19, 4, 5, 38, 23, 8, 9, 42, 27, 12, 13, 46, 31, 16, 49, 50, 35, 20, 53, 54, 39, 24, 57, 58, 43, 28, 61, 62, 47,
                                           32, 17, 2, 3, 36, 21, 6, 7, 40, 25, 10, 11, 44, 29, 14, 15, 48, 33, 18, 51, 52, 37, 22, 55, 56, 41, 26, 59, 60, 45, 30, 63};

constexpr int BT_TY = 16, BT_TX = 16, BT_PX = BT_TY * BT_TX;   // 256 pixels = UMMA N
constexpr int BT_CH = 256;                                      // channels per CTA = 2 accumulators of 128
constexpr int BT_BUILDERS = 512;                                // warps 2..17: two half-rows per operand row
constexpr int BT_THREADS = 64 + BT_BUILDERS;
constexpr int BT_XWARPS = 3;                                    // kRHWC only: two more producer warps (Wx slices, gradient boxes) + a second MMA issuer
constexpr int BT_THREADS_RHWC = BT_THREADS + 32 * BT_XWARPS;
constexpr int BT_LIST = 1024;
constexpr int BT_WARPS = BT_THREADS / 32;
constexpr int BT_ROUNDS = (BT_LIST + BT_THREADS - 1) / BT_THREADS;
constexpr int BT_A_BYTES = 128 * 128;                           // [128 ch][64 bins] bf16
constexpr int BT_B_BYTES = BT_PX * 128;                         // [256 px][64 bins] bf16
constexpr int BT_OPS_BYTES = 2 * BT_A_BYTES + BT_B_BYTES;       // one operand set: 64 KB
constexpr int BT_RAW_G = BT_CH * PP * 2;                        // 25088 B
constexpr int BT_RAW_BYTES = 64 + BT_RAW_G + (BT_TY + BT_TX) * WROW * 4 + 64;   // header | g | wy | wx  (26240, 128-multiple)
constexpr int BT_PF = 8;                                       // L2 prefetch distance (pairs)
constexpr int BT_NRAW = 3;                                     // raw ring depth (copy latency ~3 pair times)
constexpr size_t BT_SMEM = 1024 + 2 * (size_t)BT_OPS_BYTES + BT_NRAW * (size_t)BT_RAW_BYTES + BT_LIST * 4 + 256;
// kRHWC layout: [BT_NA][A slot 32 KB] | [2][B 32 KB] | [BT_NRAW_RHWC][header | wy | wx] | list | barriers
// A slot = the k-steps of one pair: [4 channel groups][16 * ksteps bins][64 ch] bf16, 128B-swizzled rows
constexpr int BT_NA = 4;                                       // gradient ring depth in pairs.  Measured and not kept: a variable-size ring (18 units of
                                                               // 8 KB = 5.9 pairs in flight) and an L2 tensor prefetch 8 pairs ahead: both +-0
constexpr int BT_ASLOT = 2 * BT_A_BYTES;                       // 32 KB: 4 k-steps
constexpr int BT_NRAW_RHWC = 8;
constexpr int BT_RAWT_BYTES = 64 + (BT_TY + BT_TX) * WROW * 4 + 64;    // 1152 (128-multiple)
constexpr size_t BT_SMEM_RHWC = 1024 + (size_t)BT_NA * BT_ASLOT + 2 * (size_t)BT_B_BYTES + BT_NRAW_RHWC * (size_t)BT_RAWT_BYTES + BT_LIST * 4 + 512;
static_assert(BT_SMEM_RHWC <= 227 * 1024 && BT_NA * BT_ASLOT >= BT_PX * BT_CH * 2, "epilogue staging overlays the A ring");
constexpr int BT_STAGE_PITCH = BT_CH * 2 + 16;                 // epilogue staging row: 512 B of channels + 16 B (bank rotation)
static_assert((size_t)BT_PX * BT_STAGE_PITCH <= (size_t)BT_NA * BT_ASLOT + 2 * BT_B_BYTES, "epilogue staging fits the operand buffers");
struct BwdMaps { CUtensorMap m[4]; };                          // boxes of 16, 32, 48, 64 bins

// Work items = (image, pixel tile) x channel chunk; their cost is the number of RoIs touching the tile (0 ... 54 pairs at the bench
// size), and the hardware hands CTAs out in index order: with tiles in raster order the last CTAs to start were heavy ones and
// the SMs were busy 151 us of a 184 us kernel (tools/trace_roi_bwd.py).  roi_tile_order_kernel sorts the (image, tile) items
// by decreasing RoI count; CTA L takes item order[L / chunks], channel chunk L % chunks (longest processing time first).
__global__ void __launch_bounds__(1024)
roi_tile_order_kernel(const unsigned char* __restrict__ ws, int R, int N, int H, int W, int tiles_x, int tiles_y, int* __restrict__ order) {
  __shared__ int cnt[1024];
  const RoiMeta* metas = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off());
  const int tiles = tiles_x * tiles_y, n_items = N * tiles, t = threadIdx.x;
  if (t < n_items) cnt[t] = 0;
  __syncthreads();
  for (int r = t; r < R; r += blockDim.x) {
    const RoiMeta m = metas[r];
    if (m.b < 0 || m.b >= N || m.ny <= 0 || m.nx <= 0) continue;
    const int ty_a = m.y_lo / BT_TY, ty_b = min((m.y_lo + m.ny - 1) / BT_TY, tiles_y - 1);
    const int tx_a = m.x_lo / BT_TX, tx_b = min((m.x_lo + m.nx - 1) / BT_TX, tiles_x - 1);
    for (int ty = ty_a; ty <= ty_b; ++ty)
      for (int tx = tx_a; tx <= tx_b; ++tx) atomicAdd(&cnt[m.b * tiles + ty * tiles_x + tx], 1);
  }
  __syncthreads();
  if (t < n_items) {     // rank by counting (n_items <= 1024): stable, deterministic
    const int mine = cnt[t];
    int rank = 0;
    for (int j = 0; j < n_items; ++j) rank += (cnt[j] > mine) || (cnt[j] == mine && j < t);
    order[rank] = t;
  }
}

template <typename TO, bool kTrace, bool kRHWC>
__global__ void __launch_bounds__(kRHWC ? BT_THREADS_RHWC : BT_THREADS, 1)
roi_align_bwd_tc_kernel(const __grid_constant__ BwdMaps gmaps, const __nv_bfloat16* __restrict__ grad_out, int C, int H, int W, int R,
                        const unsigned char* __restrict__ ws, TO* __restrict__ grad_in, int tiles_x, int dbg, unsigned long long* trace,
                        const int* __restrict__ item_order, int tiles, int chunks) {
  if (dbg & 32) return;
  unsigned long long tr0 = 0, tr1 = 0, tr2 = 0, tr3 = 0;
  const int item_l = (int)blockIdx.x / chunks, chunk_l = (int)blockIdx.x % chunks;
  const int item = item_order ? item_order[item_l] : item_l;
  const int blk_b = item / tiles, blk_tile = item % tiles;
  unsigned long long* ptrace = (kTrace && trace && blockIdx.x == 3 * chunks + 1) ? trace + 8 * (size_t)gridDim.x : nullptr;
  // per-pair stamps (tools/trace_roi_bwd.py) exist only in the kTrace instantiation: even predicated off they were ~4 % of the
  // builder warps' issue slots (ncu source page, profiles/r02_roi_align_ncu.md)
#define PSTAMP(pair, k) do { if (kTrace && ptrace && (pair) < 64) ptrace[(pair) * 8 + (k)] = (unsigned long long)clock64(); } while (0)   /* SM cycles: globaltimer ticks every 256 ns here */
  if (kTrace && trace && threadIdx.x == 64) tr0 = globaltimer_ns();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ops0 = base;                                   // [2][A0 | A1 | B]   (kRHWC: A ring, then the two B tiles)
  constexpr int NRAW = kRHWC ? BT_NRAW_RHWC : BT_NRAW;          // table slots only in the kRHWC mode: cheap, so the producer runs further ahead
  constexpr int RAW_OFF = kRHWC ? BT_NA * BT_ASLOT + 2 * BT_B_BYTES : 2 * BT_OPS_BYTES;
  constexpr int RAW_PITCH = kRHWC ? BT_RAWT_BYTES : BT_RAW_BYTES;
  constexpr int RAW_TAB = kRHWC ? 0 : BT_RAW_G;                 // offset of the table slices behind a slot's 64-byte header
  const uint32_t raw0 = ops0 + RAW_OFF;                         // [NRAW] raw slots
  int* list = reinterpret_cast<int*>(gen + RAW_OFF + NRAW * RAW_PITCH);
  const uint32_t bars = smem_u32(list + BT_LIST);
  const uint32_t raw_full0 = bars, raw_empty0 = bars + 64, ops_ready0 = bars + 128, ops_free0 = bars + 144,
                 tfull = bars + 160, tslot = bars + 168, u_full0 = bars + 192, u_empty0 = bars + 320;
  static_assert(NRAW <= 8 && BT_NA <= 16, "barrier block layout");
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tslot - base));
  __shared__ int s_wcount[BT_ROUNDS * BT_WARPS];
  __shared__ int s_rows[2];     // per operand buffer: first tile row | (end tile row << 8) of the RoI's footprint in this tile

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blk_b;
  const int c0 = chunk_l * BT_CH;
  const int ty0 = (blk_tile / tiles_x) * BT_TY, tx0 = (blk_tile % tiles_x) * BT_TX;
  const int ty1 = min(ty0 + BT_TY, H), tx1 = min(tx0 + BT_TX, W);
  const RoiMeta* metas = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off());
  const float* tables = reinterpret_cast<const float*>(ws + ws_table_off(R));
  const int nch = min(BT_CH, C - c0);

  if (t == 0) {
    for (int i = 0; i < NRAW; ++i) {
      mbar_init(raw_full0 + 8 * i, 1);
      mbar_init(raw_empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(ops_ready0 + 8 * i, 1);
      mbar_init(ops_free0 + 8 * i, kRHWC ? 2 : 1);     // kRHWC: two MMA issuers (one per accumulator) commit
    }
    mbar_init(tfull, kRHWC ? 2 : 1);
    if (kRHWC)
      for (int i = 0; i < BT_NA; ++i) {
        mbar_init(u_full0 + 8 * i, 1);
        mbar_init(u_empty0 + 8 * i, 2);
      }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tslot_ptr;

  int seq = 0;  // pairs processed so far (same in every thread)
  for (int rbase = 0; rbase < R; rbase += BT_LIST) {
    // ---- ordered compaction of the RoIs of image b that touch this tile: all loads in flight first, then
    // one ballot per round, a shared prefix over (round, warp) counts, two block barriers per chunk
    const int rend = min(rbase + BT_LIST, R);
    bool hit[BT_ROUNDS];
    unsigned bal[BT_ROUNDS];
#pragma unroll
    for (int k = 0; k < BT_ROUNDS; ++k) {
      const int r = rbase + k * BT_THREADS + t;
      hit[k] = false;
      if (r < rend && warp < BT_WARPS) {      // (the extra producer warps of the kRHWC mode take no part in the compaction)
        const RoiMeta m = metas[r];
        hit[k] = (m.b == b) && m.ny > 0 && m.y_lo < ty1 && m.y_lo + m.ny > ty0 && m.x_lo < tx1 && m.x_lo + m.nx > tx0;
      }
    }
    __syncthreads();   // previous chunk's list / counts are no longer read
#pragma unroll
    for (int k = 0; k < BT_ROUNDS; ++k) {
      bal[k] = __ballot_sync(0xffffffffu, hit[k]);
      if (lane == 0 && warp < BT_WARPS) s_wcount[k * BT_WARPS + warp] = __popc(bal[k]);
    }
    __syncthreads();
    int total = 0;
#pragma unroll
    for (int k = 0; k < BT_ROUNDS; ++k) {
      int off = total;
      for (int w = 0; w < BT_WARPS; ++w) {
        const int cnt = s_wcount[k * BT_WARPS + w];
        if (w < warp) off += cnt;
        total += cnt;
      }
      if (hit[k]) list[off + __popc(bal[k] & ((1u << lane) - 1u))] = rbase + k * BT_THREADS + t;
    }
    __syncthreads();
    const int n_list = (dbg & 64) ? 0 : total;
    if (kTrace && trace && threadIdx.x == 64 && rbase == 0) tr1 = globaltimer_ns();

    if (warp == 0) {
      // ------------------------------------------------ producer: raw gradient chunks + table slices
      int my_r = 0;
      RoiMeta my_m = {};
      for (int li = 0; li < n_list; ++li) {
        const int sq = seq + li, slot = sq % NRAW;
        const uint32_t par = (uint32_t)(sq / NRAW) & 1u;
        if ((li & 31) == 0 && li + lane < n_list) {   // 32 metas per L2 round trip, off the ring's critical path
          my_r = list[li + lane];
          my_m = metas[my_r];
          if (kRHWC) {   // bin rows of this RoI that meet the tile's rows: pad column of its Wy rows (roi_prep_kernel)
            const int ya_ = max(my_m.y_lo, ty0), yb_ = min(my_m.y_lo + my_m.ny, ty1);
            const float* tab_ = tables + (size_t)my_r * (H + W) * WROW;
            const int fa = __float_as_int(tab_[(size_t)(ya_ - my_m.y_lo) * WROW + P]) & 255;
            const int lb = (__float_as_int(tab_[(size_t)(yb_ - 1 - my_m.y_lo) * WROW + P]) >> 8) & 255;   // 1 + last bin row
            const int pa_ = min(fa, P - 1), pb_ = max(min(lb, P), pa_ + 1);
            my_m.gh = pa_;                                   // (gh / gw are not used by this kernel)
            my_m.gw = ((pb_ - pa_) * P + 15) >> 4;           // k-steps: 1..4
          }
        }
        // pull the gradient chunk of the pair BT_PF ring slots ahead into L2 (the ring itself only covers ~3 us)
        if (!kRHWC && lane == 3 && li + BT_PF < n_list && !(dbg & 128))
          bulk_prefetch_l2(grad_out + ((size_t)list[li + BT_PF] * C + c0) * PP, (uint32_t)(nch * PP * 2));
        const int src = li & 31;
        const int r = __shfl_sync(0xffffffffu, my_r, src);
        RoiMeta m;
        m.y_lo = __shfl_sync(0xffffffffu, my_m.y_lo, src);
        m.ny = __shfl_sync(0xffffffffu, my_m.ny, src);
        m.x_lo = __shfl_sync(0xffffffffu, my_m.x_lo, src);
        m.nx = __shfl_sync(0xffffffffu, my_m.nx, src);
        m.count = __shfl_sync(0xffffffffu, my_m.count, src);
        const int pa = kRHWC ? __shfl_sync(0xffffffffu, my_m.gh, src) : 0, ksteps = kRHWC ? __shfl_sync(0xffffffffu, my_m.gw, src) : 4;
        const int ya = max(m.y_lo, ty0), yb = min(m.y_lo + m.ny, ty1);
        const int xa = max(m.x_lo, tx0), xb = min(m.x_lo + m.nx, tx1);
        mbar_wait(raw_empty0 + 8 * slot, par ^ 1u);
        if (lane == 0) PSTAMP(sq, 0);
        uint8_t* sl = gen + RAW_OFF + slot * RAW_PITCH;
        const uint32_t g_bytes = kRHWC ? 0u : (dbg & 16) ? 16u : (uint32_t)(nch * PP * 2);
        const uint32_t fb = raw_full0 + 8 * slot;
        if (lane == 0) {
          int* hdr = reinterpret_cast<int*>(sl);
          hdr[0] = ya; hdr[1] = yb; hdr[2] = xa; hdr[3] = xb;
          reinterpret_cast<float*>(sl)[4] = 1.f / (float)m.count;
          hdr[5] = pa; hdr[6] = ksteps;
          mbar_expect_tx(fb, g_bytes + (uint32_t)((yb - ya) + (xb - xa)) * WROW * 4);
        }
        __syncwarp();
        const float* tab = tables + (size_t)r * (H + W) * WROW;
        const uint32_t sbase = raw0 + slot * RAW_PITCH + 64;
        if (!kRHWC && lane == 0) bulk_g2s(sbase, grad_out + ((size_t)r * C + c0) * PP, g_bytes, fb);
        if (lane == 1) bulk_g2s(sbase + RAW_TAB, tab + (size_t)(ya - m.y_lo) * WROW, (uint32_t)(yb - ya) * WROW * 4, fb);
        if (!kRHWC && lane == 2) bulk_g2s(sbase + RAW_TAB + BT_TY * WROW * 4, tab + (size_t)H * WROW + (size_t)(xa - m.x_lo) * WROW,
                                          (uint32_t)(xb - xa) * WROW * 4, fb);
      }
    } else if (kRHWC && warp >= BT_WARPS && warp < BT_WARPS + 2) {
      // ------------------------------------------------ kRHWC: two more producer warps.  One wait -> expect_tx -> issue sequence
      // of a thread costs ~400 cycles whatever its size and the sequences of one warp serialise: with all three copies of a
      // pair on warp 0 the producer alone was ~1300 cycles per pair (the whole kernel without MMAs and weight build: 1390).
      // warp BT_WARPS: Wx slice (completes on the slot's raw_full, whose transaction count warp 0 posts);
      // warp BT_WARPS + 1: the pair's gradient box.
      const bool is_a = warp == BT_WARPS + 1;
      int my_r = 0;
      RoiMeta my_m = {};
      for (int li = 0; li < n_list; ++li) {
        const int sq = seq + li;
        if ((li & 31) == 0 && li + lane < n_list) {
          my_r = list[li + lane];
          my_m = metas[my_r];
          if (is_a) {   // bin rows of this RoI that meet the tile's rows: pad column of its Wy rows (roi_prep_kernel)
            const int ya_ = max(my_m.y_lo, ty0), yb_ = min(my_m.y_lo + my_m.ny, ty1);
            const float* tab_ = tables + (size_t)my_r * (H + W) * WROW;
            const int fa = __float_as_int(tab_[(size_t)(ya_ - my_m.y_lo) * WROW + P]) & 255;
            const int lb = (__float_as_int(tab_[(size_t)(yb_ - 1 - my_m.y_lo) * WROW + P]) >> 8) & 255;
            const int pa_ = min(fa, P - 1), pb_ = max(min(lb, P), pa_ + 1);
            my_m.gh = pa_;
            my_m.gw = ((pb_ - pa_) * P + 15) >> 4;
          }
        }
        const int src = li & 31;
        const int r = __shfl_sync(0xffffffffu, my_r, src);
        if (is_a) {
          const int pa = __shfl_sync(0xffffffffu, my_m.gh, src), ksteps = __shfl_sync(0xffffffffu, my_m.gw, src);
          const int aslot = sq % BT_NA;
          mbar_wait(u_empty0 + 8 * aslot, ((uint32_t)(sq / BT_NA) & 1u) ^ 1u);
          if (lane == 0) {
            // box (64 ch, 16 * ksteps bins from bin row pa on, 4 channel groups, 1 RoI) -> smem [grp][bin][64 ch], swizzled; bins
            // >= 49 and channel groups past C/64 are out of bounds = zero fill (counted in the transaction bytes)
            mbar_expect_tx(u_full0 + 8 * aslot, (uint32_t)ksteps * (BT_ASLOT / 4));
            tma_load_4d(ops0 + aslot * BT_ASLOT, &gmaps.m[ksteps - 1], u_full0 + 8 * aslot, 0, pa * P, c0 >> 6, r);
          }
        } else {
          const int x_lo = __shfl_sync(0xffffffffu, my_m.x_lo, src), nx = __shfl_sync(0xffffffffu, my_m.nx, src);
          const int xa = max(x_lo, tx0), xb = min(x_lo + nx, tx1);
          const int slot = sq % NRAW;
          mbar_wait(raw_empty0 + 8 * slot, ((uint32_t)(sq / NRAW) & 1u) ^ 1u);
          if (lane == 0)
            bulk_g2s(raw0 + slot * RAW_PITCH + 64 + RAW_TAB + BT_TY * WROW * 4,
                     tables + (size_t)r * (H + W) * WROW + (size_t)H * WROW + (size_t)(xa - x_lo) * WROW, (uint32_t)(xb - xa) * WROW * 4,
                     raw_full0 + 8 * slot);
        }
      }
    } else if (warp == 1 || (kRHWC && warp == BT_WARPS + 2)) {
      // ------------------------------------------------ MMA issuer(s)
      // kRHWC: TWO issuers, one per accumulator (128-channel block).  The cycle stamps showed the single issuer as the serial
      // bottleneck of the pair loop: ~1450 cycles per pair = two barrier waits that had long completed (~150-250 each), ~130-150
      // per tcgen05.mma whatever its N, the commits.  Each accumulator still sees its MMAs in RoI order (deterministic sums).
      const int my_cb = (warp == 1) ? 0 : 1;
      if (lane == 0) {
        for (int li = 0; li < n_list; ++li) {
          const int sq = seq + li, ob = sq & 1;
          const uint32_t par = (uint32_t)(sq >> 1) & 1u;
          mbar_wait(ops_ready0 + 8 * ob, par);
          if (my_cb == 0) PSTAMP(sq, 1);
          tc_fence_after();
          const uint32_t ops = ops0 + ob * BT_OPS_BYTES;
          const uint32_t b_tile = kRHWC ? ops0 + BT_NA * BT_ASLOT + ob * BT_B_BYTES : ops + 2 * BT_A_BYTES;
          // Only the tile rows the RoI touches take part: N = 16 px x (rows touched) instead of the whole 16x16 tile (the
          // weight rows of the other pixels are zero, and are not even built any more).  (Saves builder work only: the cycle
          // stamps show ~130-150 cycles per tcgen05.mma whatever N <= 256 is -- an M = 128 instruction is paced by its A
          // operand.  Fewer, wider MMAs would need the roles swapped: pixels as M, 256 channels as N)  The very first pair of the CTA runs
          // the full N = 256 with accumulate = 0: it is what initialises both TMEM accumulators.
          int r0 = 0, nrows = BT_TY, ksteps = 4;
          const int rr = *reinterpret_cast<volatile int*>(&s_rows[ob]);
          if (sq > 0) {
            r0 = rr & 255;
            nrows = ((rr >> 8) & 255) - r0;
          }
          if (kRHWC) ksteps = rr >> 16;
          const uint32_t idesc_n = make_idesc(128, sq > 0 ? BT_TX * nrows : BT_PX, kRHWC ? 1 : 0, 0);
          const uint64_t bd = desc_kmajor_sw128(b_tile + (uint32_t)r0 * (BT_TX * 128));
          if constexpr (kRHWC) {
            // ROLES SWAPPED against the [R,C,7,7] mode: D_h[128 px x 256 ch] += Wt_h[128 px x bins] . G[bins x 256 ch] for the
            // pixel half h (tile rows 8h .. 8h+7) of this issuer, skipped when the RoI misses that half.  An M = 128 tcgen05.mma
            // costs ~130-150 cycles whatever its N (cycle stamps, profiles/r02_roi_align_rhwc.md): restricting N to the touched
            // rows never saved tensor-pipe time; one full-N MMA per touched half does (1.45 instead of 2 per k-step at the bench
            // size).  A = weight tile (K-major, 128B swizzle, rows = pixels); B = gradient box, MN-major as it lies in memory:
            // 64-channel groups ksteps * 2 KB apart (LBO), 8-bin atoms 1 KB apart (SBO), 16 bins = 2 KB per k-step.
            // (the team leader waited for the pair's gradient box before it arrived on ops_ready: ONE barrier wait and ONE commit
            // per pair on this thread -- each already-completed mbarrier wait cost it 150-300 cycles, two of them plus two commits
            // were ~530 of its ~1090 cycles per pair)
            const int aslot = sq % BT_NA;
            if (my_cb == 0) PSTAMP(sq, 7);
            const uint32_t a_tile = ops0 + aslot * BT_ASLOT, gpitch = (uint32_t)ksteps * 2048u;
            const int ya_r = rr & 255, yb_r = (rr >> 8) & 255;
            const bool touched = sq == 0 || (ya_r < 8 * (my_cb + 1) && yb_r > 8 * my_cb);   // first pair: initialises the accumulator
            constexpr uint32_t idesc_sw = make_idesc(128, BT_CH, 0, 1);
            const uint64_t wd = desc_kmajor_sw128(b_tile + (uint32_t)my_cb * (128 * 128));
            if (touched && !(dbg & 1))
              for (int kk = 0; kk < ksteps; ++kk)
                umma_bf16(tmem_base + my_cb * BT_CH, wd + (uint64_t)(kk * 2), desc_mnmajor_sw128(a_tile + kk * 2048, gpitch), idesc_sw,
                          (sq > 0 || kk > 0) ? 1u : 0u);
            umma_commit(u_empty0 + 8 * aslot);
          } else {
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
              const uint64_t ad = desc_kmajor_sw128(ops + cb * BT_A_BYTES);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                if (!(dbg & 1)) umma_bf16(tmem_base + cb * BT_PX + r0 * BT_TX, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc_n, (sq > 0 || kk > 0) ? 1u : 0u);
            }
          }
          if (!kRHWC) umma_commit(ops_free0 + 8 * ob);     // kRHWC: the weight buffer is released by the pair's u_empty commit
          if (my_cb == 0) PSTAMP(sq, 2);
        }
      }
    } else {
      // ------------------------------------------------ operand builders (512 threads)
      // thread = (operand row, 64-byte half of its 128-byte K row); halves are warp-uniform
      // Lane -> operand row: the raw gradient rows have a 98-byte pitch, so 32 CONSECUTIVE rows put 8 of the 32 lanes of every
      // relayout LDS.32 on an already used bank (2 passes per load: ~430 of the ~1600 shared-memory-port cycles per pair,
      // profiles/r02_roi_bwd_pipeline_study.md).  Within each group of 64 rows the two warps take the row sets below instead:
      // floor(24.5 r) mod 32 is distinct over a warp (conflict-free loads) and every aligned octet of lanes holds all residues
      // r mod 8 (conflict-free 128B-swizzled STS.128, as before).  Pure re-assignment of work: results are bit-identical.
      // kRHWC: the builders only build the weight tile, and what a pair costs them is latency (two barrier waits, table reads,
      // proxy fence, team barrier, arrive: ~700-1000 cycles with ~150 of arithmetic), so they work as TWO TEAMS of 256 threads on
      // alternate pairs -- team t owns weight buffer t, a thread builds a whole 128-byte row -- and the latencies of consecutive
      // pairs overlap (tools/trace_roi_bwd.py: the kernel without MMAs and weight arithmetic was 1145 cycles per pair).
      const int bt = t - 64, row = (dbg & 256) ? (bt & 255) : ((bt & 192) | kRowPerm[bt & 63]);
      const int team = kRHWC ? (bt >> 8) : 0;
      const bool leader = kRHWC ? (bt & 255) == 0 : bt == 0;
      for (int li = 0; li < n_list; ++li) {
        const int sq = seq + li, slot = sq % NRAW, ob = sq & 1;
        if (kRHWC && ob != team) continue;
        mbar_wait(raw_full0 + 8 * slot, (uint32_t)(sq / NRAW) & 1u);
        if (bt == 0) PSTAMP(sq, 3);
        if (kRHWC) {     // weight buffer ob was last read by the MMAs of pair sq - 2: its u_empty completion releases it
          if (sq >= 2) mbar_wait(u_empty0 + 8 * ((sq - 2) % BT_NA), (uint32_t)((sq - 2) / BT_NA) & 1u);
        } else {
          mbar_wait(ops_free0 + 8 * ob, ((uint32_t)(sq >> 1) & 1u) ^ 1u);
        }
        if (bt == 0) PSTAMP(sq, 4);
        const uint8_t* sl = gen + RAW_OFF + slot * RAW_PITCH;
        const int* hdr = reinterpret_cast<const int*>(sl);
        const int ya = hdr[0], yb = hdr[1], xa = hdr[2], xb = hdr[3];
        const float inv_count = reinterpret_cast<const float*>(sl)[4];
        const float* wy_s = reinterpret_cast<const float*>(sl + 64 + RAW_TAB);
        const float* wx_s = wy_s + BT_TY * WROW;
        uint8_t* ops = gen + ob * BT_OPS_BYTES;
        uint8_t* b_ops = kRHWC ? gen + BT_NA * BT_ASLOT + ob * BT_B_BYTES : ops + 2 * BT_A_BYTES;
        const int pa = kRHWC ? hdr[5] : 0, ksteps = kRHWC ? hdr[6] : 4;
        // (1) A: 49 bf16 of channel `row` -> 64 (zero padded), 128B-swizzled K-major row.  Rows start on
        // 2-byte boundaries (98 B pitch): read aligned words and funnel-shift by 0 or 16 bits.
        if constexpr (!kRHWC) {
          const int half = bt >> 8;
          const int cb = row >> 7, mrow = row & 127;
          uint32_t pk[16];
          if (row < nch && !(dbg & 2)) {
            const uint32_t byte0 = (uint32_t)row * (PP * 2) + (uint32_t)half * 64u;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(sl + 64 + (byte0 & ~3u));
            const uint32_t sh = (byte0 & 2u) * 8u;
            if (half == 0) {
              uint32_t w[17];
#pragma unroll
              for (int j = 0; j < 17; ++j) w[j] = src[j];
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = __funnelshift_r(w[j], w[j + 1], sh);
            } else {
              uint32_t w[10];
#pragma unroll
              for (int j = 0; j < 10; ++j) w[j] = src[j];
#pragma unroll
              for (int j = 0; j < 9; ++j) pk[j] = __funnelshift_r(w[j], w[j + 1], sh);
              pk[8] &= 0xFFFFu;  // element 48 | pad
#pragma unroll
              for (int j = 9; j < 16; ++j) pk[j] = 0u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          uint8_t* arow = ops + cb * BT_A_BYTES + (mrow >> 3) * 1024 + (mrow & 7) * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(arow + (((half * 4 + k) ^ (mrow & 7)) << 4)) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        }
        // (2) B: Wt[px][bin] = Wy[y][ph] * Wx[x][pw] / count inside the footprint, else 0.  Tile rows the RoI does not
        // touch are skipped (the MMA only reads rows [ya, yb)); the first pair of the CTA still writes all 256 rows.
#pragma unroll
        for (int hh = 0; hh < (kRHWC ? 2 : 1); ++hh) {
        const int half = kRHWC ? hh : (bt >> 8);
        // ([R,C,7,7] mode: only the tile rows the RoI touches are read by the restricted-N MMAs; kRHWC: every row of a pixel half
        // the RoI touches is an M row of that half's MMA, rows outside the footprint are written as zeros)
        const int hrow = row >> 7;
        const bool row_used = kRHWC ? ((ya - ty0) < 8 * (hrow + 1) && (yb - ty0) > 8 * hrow)
                                    : ((ty0 + (row >> 4)) >= ya && (ty0 + (row >> 4)) < yb);
        if ((sq == 0 || row_used) && half * 2 < ksteps) {
          const int y = ty0 + (row >> 4), x = tx0 + (row & 15);
          const bool in = (y >= ya) && (y < yb) && (x >= xa) && (x < xb);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
          if (in && !(dbg & 4)) {
            const float4 wa = reinterpret_cast<const float4*>(wy_s)[(y - ya) * 2], wb = reinterpret_cast<const float4*>(wy_s)[(y - ya) * 2 + 1];
            const float4 xa4 = reinterpret_cast<const float4*>(wx_s)[(x - xa) * 2], xb4 = reinterpret_cast<const float4*>(wx_s)[(x - xa) * 2 + 1];
            float wyv[P] = {wa.x * inv_count, wa.y * inv_count, wa.z * inv_count, wa.w * inv_count,
                            wb.x * inv_count, wb.y * inv_count, wb.z * inv_count};
            if (kRHWC && pa > 0) {   // K column j of the pair is bin pa*7 + j: bin row pa + j/7 (zero weight past the last one)
              const float* wrow = wy_s + (y - ya) * WROW;
#pragma unroll
              for (int i = 0; i < P; ++i) wyv[i] = (pa + i < P) ? wrow[pa + i] * inv_count : 0.f;
            }
            const float wxv[P] = {xa4.x, xa4.y, xa4.z, xa4.w, xb4.x, xb4.y, xb4.z};
            if (half == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n0 = 2 * j, n1 = 2 * j + 1;
                __nv_bfloat162 v = __floats2bfloat162_rn(wyv[n0 / P] * wxv[n0 % P], wyv[n1 / P] * wxv[n1 % P]);
                pk[j] = *reinterpret_cast<uint32_t*>(&v);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 9; ++j) {
                const int n0 = 32 + 2 * j, n1 = 33 + 2 * j;
                const float v1 = (n1 < PP) ? wyv[(n1 < PP ? n1 : 0) / P] * wxv[n1 % P] : 0.f;
                __nv_bfloat162 v = __floats2bfloat162_rn(wyv[n0 / P] * wxv[n0 % P], v1);
                pk[j] = *reinterpret_cast<uint32_t*>(&v);
              }
            }
          }
          uint8_t* brow = b_ops + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(brow + (((half * 4 + k) ^ (row & 7)) << 4)) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        }
        }
        if (bt == 0) PSTAMP(sq, 5);
        fence_proxy_async();
        named_bar_sync(kRHWC ? 3 + team : 2, kRHWC ? BT_BUILDERS / 2 : BT_BUILDERS);
        if (bt == 0) PSTAMP(sq, 6);
        if (leader) {
          s_rows[ob] = (ya - ty0) | ((yb - ty0) << 8) | (ksteps << 16);     // read by the MMA thread after it has seen ops_ready (release/acquire)
          if (kRHWC) mbar_wait(u_full0 + 8 * (sq % BT_NA), (uint32_t)(sq / BT_NA) & 1u);   // the pair's gradient box has landed
          mbar_arrive(ops_ready0 + 8 * ob);
          mbar_arrive(raw_empty0 + 8 * slot);
        }
      }
    }
    seq += n_list;
  }

  if (kTrace && trace && threadIdx.x == 64) tr2 = globaltimer_ns();
  // ---- epilogue: TMEM -> grad_input (or zeros when no RoI touches the tile)
  if ((warp == 1 || (kRHWC && warp == BT_WARPS + 2)) && lane == 0 && seq > 0) umma_commit(tfull);
  if (kRHWC && warp >= 2 && warp < BT_WARPS) {
    // pixel-major accumulators: lane = pixel of half h, columns = the CTA's 256 channels.  A thread holds 32 consecutive channels of
    // its pixel per TMEM load: 64 (bf16) / 128 (fp32) contiguous bytes of grad_input NHWC, stored directly -- no transpose through
    // shared memory (the [R,C,7,7] mode's epilogue is 256 2-byte shared stores per thread + bulk stores: 3.7 us per CTA).
    const int q = warp & 3, h = ((warp - 2) >> 2) & 1, chalf = (warp - 2) >> 3;   // TMEM quadrant, pixel half, channel half
    const int px = h * 128 + q * 32 + lane;
    const int y = ty0 + (px >> 4), x = tx0 + (px & 15);
    if (seq > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    if (kTrace && trace && threadIdx.x == 64) tr3 = globaltimer_ns();
#pragma unroll 1
    for (int cc = chalf * 4; cc < chalf * 4 + 4; ++cc) {
      uint32_t v[32];
      if (seq > 0) {
        DA_TMEM_LD32(tmem_base + ((uint32_t)(q * 32) << 16) + h * BT_CH + cc * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int c = c0 + cc * 32;
      if (y < H && x < W && c < C && !(dbg & 8)) {      // C % 64 == 0: a 32-channel run is inside or outside as a whole
        TO* dst = grad_in + (((size_t)b * H + y) * W + x) * C + c;
        if constexpr (sizeof(TO) == 4) {
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int k = 0; k < 8; ++k) d4[k] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
      }
      if constexpr (sizeof(TO) == 2) {
        // bf16: stage[px][256 ch] over the (now idle) operand buffers with a 528-byte row pitch -- the 16-byte stores of 8
        // consecutive pixels then fall on 8 different bank groups -- and one 512 B bulk store per pixel (16-byte global stores
        // at a 4 KB stride measured 6.8 us per CTA)
        uint4* s4 = reinterpret_cast<uint4*>(gen + (size_t)px * BT_STAGE_PITCH + cc * 64);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[8 * k + 2 * i]), __uint_as_float(v[8 * k + 2 * i + 1]));
            pk[i] = *reinterpret_cast<const uint32_t*>(&t2);
          }
          s4[k] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    if constexpr (sizeof(TO) == 2) {
      fence_proxy_async();
      named_bar_sync(2, BT_BUILDERS);
      const int p2 = t - 64;
      if (p2 < BT_PX) {
        const int y2 = ty0 + (p2 >> 4), x2 = tx0 + (p2 & 15);
        if (y2 < H && x2 < W && !(dbg & 8)) {
          bulk_s2g(grad_in + (((size_t)b * H + y2) * W + x2) * C + c0, base + (uint32_t)p2 * BT_STAGE_PITCH, (uint32_t)nch * 2u);
          bulk_commit();
          bulk_wait_read0();
        }
      }
    }
    tc_fence_before();
  } else if (warp >= 2 && warp < BT_WARPS) {
    const int q = warp & 3, cb = ((warp - 2) >> 2) & 1, chalf = (warp - 2) >> 3;   // TMEM quadrant, accumulator, pixel half
    const int c = c0 + cb * 128 + q * 32 + lane;
    if (seq > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    if (kTrace && trace && threadIdx.x == 64) tr3 = globaltimer_ns();
#pragma unroll 1
    for (int cc = chalf * (BT_PX / 64); cc < (chalf + 1) * (BT_PX / 64); ++cc) {
      uint32_t v[32];
      if (seq > 0) {
        DA_TMEM_LD32(tmem_base + ((uint32_t)(q * 32) << 16) + cb * BT_PX + cc * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if constexpr (sizeof(TO) == 2) {
        // bf16: transpose through the (now idle) operand buffers to stage[px][256 ch]; 64 B conflict-free
        // warp stores, then one 512 B bulk store per pixel (2-byte scattered stores measured 10 us per CTA)
        __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(gen);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          stage[(cc * 32 + j) * BT_CH + cb * 128 + q * 32 + lane] = __float2bfloat16_rn(__uint_as_float(v[j]));
      } else {
        if (c < C) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int px = cc * 32 + j;
            const int y = ty0 + (px >> 4), x = tx0 + (px & 15);
            if (y < H && x < W && !(dbg & 8)) grad_in[(((size_t)b * H + y) * W + x) * C + c] = from_f32<TO>(__uint_as_float(v[j]));
          }
        }
      }
    }
    if constexpr (sizeof(TO) == 2) {
      fence_proxy_async();
      named_bar_sync(2, BT_BUILDERS);
      const int px = t - 64;
      if (px < BT_PX) {
        const int y = ty0 + (px >> 4), x = tx0 + (px & 15);
        if (y < H && x < W && !(dbg & 8)) {
          bulk_s2g(grad_in + (((size_t)b * H + y) * W + x) * C + c0, base + (uint32_t)px * (BT_CH * 2), (uint32_t)nch * 2u);
          bulk_commit();
          bulk_wait_read0();
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (kTrace && trace && threadIdx.x == 64) {
    unsigned long long* o = trace + 8 * (size_t)blockIdx.x;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    o[0] = tr0; o[1] = tr1; o[2] = tr2; o[3] = globaltimer_ns(); o[4] = (unsigned long long)seq; o[5] = smid; o[6] = tr3;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename TO, bool kRHWC>
static int launch_bwd_tc(const void* grad_out, int N, int C, int H, int W, int R, const void* ws, void* grad_in, cudaStream_t st) {
  const int tiles_x = (W + BT_TX - 1) / BT_TX, tiles_y = (H + BT_TY - 1) / BT_TY;
  const int tiles = tiles_x * tiles_y, chunks = (C + BT_CH - 1) / BT_CH;
  DA_REQUIRE((long long)N * tiles * chunks <= 0x7fffffffll, DA_ERR_UNSUPPORTED, "roi_align_backward tc: grid too large");
  dim3 grid((unsigned)(N * tiles * chunks));
  BwdMaps gmaps;
  memset(&gmaps, 0, sizeof(gmaps));
  if (kRHWC) {
    // grad_out [R][49][C] viewed as (64 ch, 49 bins, C/64 groups, R): one box = the k-steps ([16 k bins][256 ch]) of a (RoI, CTA) pair
    DA_REQUIRE(C % 64 == 0, DA_ERR_UNSUPPORTED, "roi_align_backward tc ([R,7,7,C] gradients): C must be a multiple of 64");
    const uint64_t dims[4] = {64, (uint64_t)PP, (uint64_t)(C / 64), (uint64_t)R};
    const uint64_t strides[3] = {(uint64_t)C * 2, 128, (uint64_t)PP * C * 2};
    for (int k = 1; k <= 4; ++k) {
      const uint32_t box[4] = {64, (uint32_t)(16 * k), BT_CH / 64, 1};
      int rc = encode_map(&gmaps.m[k - 1], grad_out, 4, dims, strides, box);
      if (rc) return rc;
    }
  }
  // longest-processing-time-first order of the (image, tile) items; the list lives in the workspace's RoI-order region (only the
  // forward reads that one, and every forward rewrites it)
  int* item_order = nullptr;
  if ((long long)N * tiles <= 1024 && (long long)N * tiles <= R && !(g_opt.roi_bwd_dbg & 512)) {
    item_order = reinterpret_cast<int*>(const_cast<unsigned char*>((const unsigned char*)ws) + ws_order_off(R, H, W));
    roi_tile_order_kernel<<<1, 1024, 0, st>>>((const unsigned char*)ws, R, N, H, W, tiles_x, tiles_y, item_order);
    DA_LAUNCH_CHECK();
  }
  const int dbg = g_opt.roi_bwd_dbg;   // timing experiments only (results are wrong when set)
  unsigned long long* trace = reinterpret_cast<unsigned long long*>(g_opt.roi_bwd_trace);   // tools/trace_roi_bwd.py
  constexpr size_t smem = kRHWC ? BT_SMEM_RHWC : BT_SMEM;
  static bool attr_set_dev[kMaxDevices] = {};     // function attributes are per device (one flag array per instantiation)
  bool& attr_set = attr_set_dev[cur_dev()];
  if (!attr_set) {
    DA_CUDA_OK(cudaFuncSetAttribute(roi_align_bwd_tc_kernel<TO, false, kRHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DA_CUDA_OK(cudaFuncSetAttribute(roi_align_bwd_tc_kernel<TO, true, kRHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (trace)
    roi_align_bwd_tc_kernel<TO, true, kRHWC><<<grid, kRHWC ? BT_THREADS_RHWC : BT_THREADS, smem, st>>>(gmaps, (const __nv_bfloat16*)grad_out, C, H, W, R, (const unsigned char*)ws,
                                                                                 static_cast<TO*>(grad_in), tiles_x, dbg, trace, item_order, tiles, chunks);
  else
    roi_align_bwd_tc_kernel<TO, false, kRHWC><<<grid, kRHWC ? BT_THREADS_RHWC : BT_THREADS, smem, st>>>(gmaps, (const __nv_bfloat16*)grad_out, C, H, W, R, (const unsigned char*)ws,
                                                                                  static_cast<TO*>(grad_in), tiles_x, dbg, trace, item_order, tiles, chunks);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

int roi_align_bwd_tc(const void* grad_out, int layout, int N, int C, int H, int W, int R, const void* ws, void* grad_in, int grad_in_dtype,
                     cudaStream_t st) {
  if (layout == DA_ROI_OUT_RHWC) {
    if (grad_in_dtype == DA_BF16) return launch_bwd_tc<__nv_bfloat16, true>(grad_out, N, C, H, W, R, ws, grad_in, st);
    return launch_bwd_tc<float, true>(grad_out, N, C, H, W, R, ws, grad_in, st);
  }
  if (grad_in_dtype == DA_BF16) return launch_bwd_tc<__nv_bfloat16, false>(grad_out, N, C, H, W, R, ws, grad_in, st);
  return launch_bwd_tc<float, false>(grad_out, N, C, H, W, R, ws, grad_in, st);
}

}  // namespace da
