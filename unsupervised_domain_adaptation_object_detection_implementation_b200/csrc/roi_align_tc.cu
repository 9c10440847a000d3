// RoIAlign forward on the tensor cores (bf16 features) — sm_100a.
//
// For one RoI the separable interpolation is a small dense contraction over the K footprint pixels:
//     out[c, (ph,pw)] = sum_k F[k, c] * Wt[k, (ph,pw)],    Wt[k,(ph,pw)] = Wy[y_k,ph] * Wx[x_k,pw] / count
// The CUDA-core kernels in roi_align.cu spend 23 instructions per (pixel, 64 channels) on it and are
// issue bound at ~8 % of HBM peak (profiles/r01_roi_align_ncu.md).  Here the FMA work moves to
// tcgen05.mma and the kernel is left with data movement:
//   * A operand = features, MN-major.  The footprint is tiled by "quads" of 4 rows x 4 pixels = the K = 16
//     of one tcgen05.mma.  The channel axis is viewed as (C/64) x 64, so ONE rank-4 TMA box
//     (64 ch, 4 px, 4 rows, 8 groups) = 16 KB fetches a quad for 512 channels (4 channel blocks of 128):
//     measured TMA cost is ~110 cycles per op + ~21 ns/KB, so small boxes (the first version used 1-2 KB
//     ones) were bound by the per-op overhead.  Per 64-channel group the box is two 128B-swizzle atoms of
//     the canonical MN-major UMMA layout; a block's second group sits 2048 B further (descriptor LBO).
//   * B operand = Wt as bf16, K-major (no-swizzle core-matrix layout, 2 KB per quad), built in shared
//     memory by a dedicated 4-warp builder team from the prep kernel's (Wy, Wx) tables, once per RoI (chunks of 16
//     quads = 256 pixels for larger footprints, two chunk buffers so that the builders run one chunk ahead of the
//     MMAs) and reused by all channel groups.
//   * D = [128 channels x 64 (49 bins)] fp32 in TMEM; the 4 accumulators of a channel group stay live
//     while its quads stream through, two groups ping-pong (8 accumulators, 512 columns) so the
//     epilogue of group g overlaps the MMAs of group g+1; the epilogue writes the [128][49] tile to a
//     staging buffer and one cp.async.bulk store moves it to out[r, c0:c0+128, 7, 7] (contiguous).
// Persistent CTAs (one per SM) walk the RoIs round-robin; empty footprints get zeros from the epilogue warps.
// Precision: weights are rounded to bf16 (rel. 2^-9), accumulation is fp32; stated with the bf16
// path's tolerance (tests: <= 1e-2 of the max magnitude vs the oracle on bf16-rounded features).
#include "da_common.cuh"
#include "da_ptx.cuh"
#include "roi_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace da {

constexpr int TC_NISSUE = 2;                    // MMA issuer warps: block j of a stage is issued by issuer j % 2
constexpr int TC_BUILD_WARP0 = 5 + TC_NISSUE;   // first of the 4 weight-builder warps
constexpr int TC_THREADS = 32 * (TC_BUILD_WARP0 + 4);    // warp 0: TMA, warp 1 + warps 6..: MMA, warps 2..5: epilogue, last 4: weights
constexpr int TC_MCH = 128;                     // channels per accumulator (UMMA M)
constexpr int TC_NB = 64;                       // 49 bins padded to the UMMA N
constexpr int TC_GB = 4;                        // channel blocks per group (512 channels share one TMA box)
constexpr int TC_ASTAGE = TC_GB * TC_MCH * 16 * 2;   // one quad (16 pixels) x 512 channels x bf16 = 16 KB
constexpr int TC_NSTAGE = 6;
constexpr int TC_BQUAD = TC_NB * 16 * 2;        // weights of one quad: 64 bins x 16 pixels, 2 KB
constexpr int TC_CHUNK_Q = 16;                  // quads per weight chunk (256 pixels); two chunk buffers: the builders run one ahead
constexpr int TC_BBUF = TC_CHUNK_Q * TC_BQUAD;  // 32 KB
constexpr int TC_NACC = 2 * TC_GB;
constexpr int TC_NSCHED = 4;                    // depth of the in-CTA work queue (RoI indices broadcast to every role)              // TMEM accumulators: two groups ping-pong

template <typename TOut> __host__ __device__ constexpr int tc_stage_bytes() { return ((TC_MCH * PP * (int)sizeof(TOut) + 127) / 128) * 128; }
template <typename TOut> __host__ __device__ inline size_t tc_smem_bytes(int H, int W) {
  return 1024 + (size_t)TC_NSTAGE * TC_ASTAGE + 2 * (size_t)TC_BBUF + 2 * tc_stage_bytes<TOut>() +
         (size_t)(H + W + 8) * WROW * 4 + 512;
}

struct TcRoi {
  int NQ, ncg, nchunks, row0, x_lo;
};
__device__ __forceinline__ bool tc_roi(const RoiMeta& m, int H, TcRoi& t) {
  t.ncg = (m.nx + 3) >> 2;
  t.NQ = ((m.ny + 3) >> 2) * t.ncg;
  if (t.NQ < 1) return false;
  t.nchunks = (t.NQ + TC_CHUNK_Q - 1) / TC_CHUNK_Q;
  t.row0 = m.b * H + m.y_lo;
  t.x_lo = m.x_lo;
  return true;
}

// Weight tile of one quad, K-major, 128B swizzle is NOT used here: the B tile of one MMA is
// [64 bins][16 pixels] = 32 B per row, stored as the canonical no-swizzle K-major layout
// (core matrices of 8 rows x 16 B): element (n, k) at  (n>>3)*256 + (k>>3)*128 + (n&7)*16 + (k&7)*2.
__device__ __forceinline__ uint64_t desc_kmajor_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46);
}

// Work queue.  A persistent CTA processes whole RoIs, whose cost (footprint quads) spans 1..64+; a static stride
// over the RoI index left the slowest CTA with 2x the mean work.  The last block of roi_prep_kernel counting-sorts the RoIs by
// decreasing footprint; the TMA producer of every CTA pops the next index with one atomicAdd (list scheduling in
// LPT order) and broadcasts it to the other roles through a small shared-memory ring.  The order inside a bucket
// depends on atomics, the results do not: every RoI's output is independent of where and when it is computed.
// consumer side of the in-CTA queue: returns the next RoI index or -1 (drained); `arrive` = this thread releases the slot
// kWarp: called by all 32 lanes (lane 0 releases after the whole warp has read); else by one elected thread
template <bool kWarp>
__device__ __forceinline__ int sched_pop(uint32_t sfull0, uint32_t sempty0, const volatile int* sched, int qi, bool arrive) {
  const int slot = qi % TC_NSCHED;
  mbar_wait(sfull0 + 8 * slot, (uint32_t)(qi / TC_NSCHED) & 1u);
  const int r = sched[slot];
  if (kWarp) __syncwarp();
  if (arrive) mbar_arrive(sempty0 + 8 * slot);
  return r;
}

// kRHWC: out is [R,7,7,C] (DA_ROI_OUT_RHWC): the [49][128] result tile is staged bin-major and leaves through ONE tensor store
// (box 128 ch x 49 bins of the [R*49, C] view); MMAs and their order are those of the [R,C,7,7] mode: bit-identical values.
template <typename TOut, bool kRHWC>
__global__ void __launch_bounds__(TC_THREADS, 1)
roi_align_fwd_tc_kernel(const __grid_constant__ CUtensorMap fmap, const __grid_constant__ CUtensorMap omap, int C, int H, int W, int R,
                        const unsigned char* __restrict__ ws, TOut* __restrict__ out, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_ring = base;
  const uint32_t b_base = a_ring + TC_NSTAGE * TC_ASTAGE;
  const uint32_t stage0 = b_base + 2 * TC_BBUF;
  constexpr int STG = tc_stage_bytes<TOut>();
  float* wy_s = reinterpret_cast<float*>(gen + (stage0 - base) + 2 * STG);
  float* wx_s = wy_s + (size_t)(H + 4) * WROW;
  const uint32_t bars = smem_u32(wx_s + (size_t)(W + 4) * WROW);
  const uint32_t full0 = bars, empty0 = bars + 8 * TC_NSTAGE, tfull0 = bars + 16 * TC_NSTAGE,
                 tempty0 = tfull0 + 8 * TC_NACC, b_ready = tempty0 + 8 * TC_NACC, b_free = b_ready + 16, sfull0 = b_free + 16,
                 sempty0 = sfull0 + 8 * TC_NSCHED, tslot = sempty0 + 8 * TC_NSCHED, sched_a = tslot + 8;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tslot - base));
  volatile int* sched = reinterpret_cast<volatile int*>(gen + (sched_a - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RoiMeta* metas = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off());
  const float* tables = reinterpret_cast<const float*>(ws + ws_table_off(R));
  const int nblk = (C + TC_MCH - 1) / TC_MCH;
  const int ngroups = (nblk + TC_GB - 1) / TC_GB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, TC_NISSUE); }
    for (int b = 0; b < TC_NACC; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    for (int b = 0; b < 2; ++b) { mbar_init(b_ready + 8 * b, 1); mbar_init(b_free + 8 * b, TC_NISSUE); }
    for (int i = 0; i < TC_NSCHED; ++i) { mbar_init(sfull0 + 8 * i, 1); mbar_init(sempty0 + 8 * i, TC_NISSUE + 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tslot, TC_NACC * TC_NB);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tslot_ptr;

  // Loop nest shared by all roles:  RoI -> group of 4 channel blocks -> chunk of 32 quads -> quad (= ring stage)
  //                                 -> channel block (one MMA each, 4 accumulators of the group)
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: ONE 16 KB box per stage
    if (lane == 0) {
      int kq = 0, qi = 0;
      int* counter = reinterpret_cast<int*>(const_cast<unsigned char*>(ws)) + 1;
      const int* order = reinterpret_cast<const int*>(ws + ws_order_off(R, H, W));
      int nxt_i = atomicAdd(counter, 1);
      for (;; ++qi) {
        const int r = nxt_i < R ? order[nxt_i] : -1;
        if (r >= 0) nxt_i = atomicAdd(counter, 1);   // next index is in flight while this RoI is processed
        const int sslot = qi % TC_NSCHED;
        mbar_wait(sempty0 + 8 * sslot, ((uint32_t)(qi / TC_NSCHED) & 1u) ^ 1u);
        sched[sslot] = r;
        mbar_arrive(sfull0 + 8 * sslot);
        if (r < 0) break;
        TcRoi t;
        if (!tc_roi(metas[r], H, t)) continue;
        for (int g = 0; g < ngroups; ++g) {
          int rg = 0, cg = 0;
          for (int quad = 0; quad < t.NQ; ++quad, ++kq) {
            const int slot = kq % TC_NSTAGE;
            const uint32_t ph = (uint32_t)(kq / TC_NSTAGE) & 1u;
            mbar_wait(empty0 + 8 * slot, ph ^ 1u);
            const uint32_t fb = full0 + 8 * slot;
            mbar_expect_tx(fb, TC_ASTAGE);
            // box (64 ch, 4 px, 4 rows, 8 channel groups of 64): smem = [grp][row][px][64 ch]
            tma_load_4d(a_ring + slot * TC_ASTAGE, &fmap, fb, 0, t.x_lo + cg * 4, t.row0 + rg * 4, g * (TC_GB * 2));
            if (++cg == t.ncg) { cg = 0; ++rg; }
          }
        }
      }
    }
  } else if (warp == 1 || (warp >= 6 && warp < TC_BUILD_WARP0)) {
    // ------------------------------------------------------------------ MMA issuers (one elected lane each)
    if (lane == 0) {
      const int me = (warp == 1) ? 0 : warp - 5;
      constexpr uint32_t idesc = make_idesc(TC_MCH, TC_NB, 1, 0);   // A MN-major (channels), B K-major
      // A: per block j the two 64-channel groups are 2048 B apart (LBO), the two 8-pixel atoms 1024 B (SBO)
      const uint64_t ad0 = desc_mnmajor_sw128(a_ring, 2048);
      const uint64_t bd0 = desc_kmajor_nosw(b_base, 128, 256);
      int kq = 0, nbuild = 0, gcount = 0, bb = 0;
      for (int qi = 0;; ++qi) {
        const int r = sched_pop<false>(sfull0, sempty0, sched, qi, true);
        if (r < 0) break;
        TcRoi t;
        if (!tc_roi(metas[r], H, t)) continue;
        for (int g = 0; g < ngroups; ++g, ++gcount) {
          const int nb_g = min(TC_GB, nblk - g * TC_GB);
          const int accb = (gcount & 1) * TC_GB;
          const uint32_t use = (uint32_t)(gcount >> 1);
          for (int j = me; j < nb_g; j += TC_NISSUE) {
            mbar_wait(tempty0 + 8 * (accb + j), (use & 1u) ^ 1u);   // epilogue has drained this accumulator
          }
          tc_fence_after();
          for (int ch = 0; ch < t.nchunks; ++ch) {
            const bool built = (t.nchunks > 1) || (g == 0);           // a fresh weight chunk was built for (g, ch)
            const bool last_use = (t.nchunks > 1) || (g == ngroups - 1);
            if (built) {
              bb = nbuild & 1;
              mbar_wait(b_ready + 8 * bb, (uint32_t)(nbuild >> 1) & 1u);
              tc_fence_after();
              ++nbuild;
            }
            const int nq_ch = min(TC_CHUNK_Q, t.NQ - ch * TC_CHUNK_Q);
            for (int ql = 0; ql < nq_ch; ++ql, ++kq) {
              const int slot = kq % TC_NSTAGE;
              const uint32_t ph = (uint32_t)(kq / TC_NSTAGE) & 1u;
              mbar_wait(full0 + 8 * slot, ph);
              tc_fence_after();
              const uint64_t ad = ad0 + (uint64_t)((slot * TC_ASTAGE) >> 4);
              const uint64_t bd = bd0 + (uint64_t)((bb * TC_BBUF + ql * TC_BQUAD) >> 4);
              const uint32_t accf = (ch > 0 || ql > 0) ? 1u : 0u;
              if (!(dbg & 1))
                for (int j = me; j < nb_g; j += TC_NISSUE)
                  umma_bf16(tmem_base + (accb + j) * TC_NB, ad + (uint64_t)(j * 256), bd, idesc, accf);
              umma_commit(empty0 + 8 * slot);
            }
            if (last_use) umma_commit(b_free + 8 * bb);   // all of THIS issuer's MMAs reading the weight chunk have completed
          }
          for (int j = me; j < nb_g; j += TC_NISSUE) umma_commit(tfull0 + 8 * (accb + j));
        }
      }
    }
  } else if (warp >= TC_BUILD_WARP0) {
    // ------------------------------------------------------------------ weight builders (128 threads), one chunk ahead of the MMAs
    // r01 had the epilogue warps build the weights: at every RoI boundary the tensor pipe (and behind it the TMA ring) waited for
    // the last group's epilogue + the table loads + the build.  A dedicated team with two chunk buffers takes all of that off
    // the critical path, and for footprints of several chunks build(ch+1) overlaps the MMAs of chunk ch.
    const int tid = threadIdx.x - 32 * TC_BUILD_WARP0;
    const int n = tid & 63, part = tid >> 6;
    const int ph_n = n / P, pw_n = n - ph_n * P;
    const bool n_ok = n < PP;
    int nbuild = 0;
    for (int qi = 0;; ++qi) {
      const int r = sched_pop<true>(sfull0, sempty0, sched, qi, lane == 0);
      if (r < 0) break;
      const RoiMeta m = metas[r];
      TcRoi t;
      if (!tc_roi(m, H, t)) continue;
      const int nrg = (m.ny + 3) >> 2;
      {   // tables -> smem (1/count folded into Wy; both zero-padded to whole quads); the previous RoI's builds are behind a barrier
        const float* tab = tables + (size_t)r * (H + W) * WROW;
        const float inv = 1.f / (float)m.count;
        for (int i = tid; i < nrg * 4 * WROW; i += 128) wy_s[i] = (i < m.ny * WROW) ? tab[i] * inv : 0.f;
        for (int i = tid; i < t.ncg * 4 * WROW; i += 128) wx_s[i] = (i < m.nx * WROW) ? tab[(size_t)H * WROW + i] : 0.f;
        named_bar_sync(3, 128);
      }
      for (int g = 0; g < ngroups; ++g) {
        for (int ch = 0; ch < t.nchunks; ++ch) {
          const bool built = (t.nchunks > 1) || (g == 0);
          if (!built) continue;
          const int bb = nbuild & 1;
          mbar_wait(b_free + 8 * bb, ((uint32_t)(nbuild >> 1) & 1u) ^ 1u);
          ++nbuild;
          uint8_t* bdst = gen + (b_base - base) + bb * TC_BBUF;
          const int q0 = ch * TC_CHUNK_Q, nq_ch = min(TC_CHUNK_Q, t.NQ - q0);
          // one 16-byte chunk (8 pixels of one bin row) per (atom, n): atom = 2*quad + (rows 0-1 | rows 2-3)
          for (int atom = part; atom < nq_ch * 2; atom += 2) {
            const int ql = atom >> 1;
            const int quad = q0 + ql;
            const int rg = quad / t.ncg, cg = quad - rg * t.ncg;
            const int ry0 = rg * 4 + (atom & 1) * 2;
            const float wy0 = n_ok ? wy_s[ry0 * WROW + ph_n] : 0.f;
            const float wy1 = n_ok ? wy_s[(ry0 + 1) * WROW + ph_n] : 0.f;
            const float* wxp = wx_s + (size_t)cg * 4 * WROW + pw_n;
            const float x0 = wxp[0], x1 = wxp[WROW], x2 = wxp[2 * WROW], x3 = wxp[3 * WROW];
            __nv_bfloat162 v0 = __floats2bfloat162_rn(wy0 * x0, wy0 * x1), v1 = __floats2bfloat162_rn(wy0 * x2, wy0 * x3);
            __nv_bfloat162 v2 = __floats2bfloat162_rn(wy1 * x0, wy1 * x1), v3 = __floats2bfloat162_rn(wy1 * x2, wy1 * x3);
            const uint32_t off = (uint32_t)ql * TC_BQUAD + (uint32_t)(n >> 3) * 256u + (uint32_t)(atom & 1) * 128u + (uint32_t)(n & 7) * 16u;
            *reinterpret_cast<uint4*>(bdst + off) =
                make_uint4(*reinterpret_cast<uint32_t*>(&v0), *reinterpret_cast<uint32_t*>(&v1),
                           *reinterpret_cast<uint32_t*>(&v2), *reinterpret_cast<uint32_t*>(&v3));
          }
          fence_proxy_async();
          named_bar_sync(3, 128);
          if (tid == 0) mbar_arrive(b_ready + 8 * bb);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int tid = threadIdx.x - 64;
    const int q = warp & 3;
    int nstore = 0, gcount = 0;
    for (int qi = 0;; ++qi) {
      const int r = sched_pop<true>(sfull0, sempty0, sched, qi, lane == 0);
      if (r < 0) break;
      TcRoi t;
      if (!tc_roi(metas[r], H, t)) {   // empty footprint (degenerate / outside / bad batch index): zeros, as the reference
        TOut* o = out + (size_t)r * C * PP;
        for (int i = tid; i < C * PP; i += 128) o[i] = from_f32<TOut>(0.f);
        continue;
      }
      for (int g = 0; g < ngroups; ++g, ++gcount) {
        const int nb_g = min(TC_GB, nblk - g * TC_GB);
        // epilogue of the group's channel blocks
        const int accb = (gcount & 1) * TC_GB;
        const uint32_t use = (uint32_t)(gcount >> 1);
        for (int j = 0; j < nb_g; ++j, ++nstore) {
          const int mb = g * TC_GB + j;
          const int sb = nstore & 1;
          if (tid == 0) bulk_wait_read<1>();       // the bulk store that last used staging[sb] has been read out
          named_bar_sync(2, 128);
          mbar_wait(tfull0 + 8 * (accb + j), use & 1u);
          tc_fence_after();
          uint32_t v0[32], v1[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (accb + j) * TC_NB;
          DA_TMEM_LD32(taddr, v0);
          DA_TMEM_LD32(taddr + 32, v1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * (accb + j));
          TOut* stg = reinterpret_cast<TOut*>(gen + (stage0 - base) + sb * STG) + (size_t)(q * 32 + lane) * (kRHWC ? 1 : PP);
          if constexpr (kRHWC) {
            // bin-major staging [49][128 ch]: a warp's store covers 32 consecutive channels of one bin (conflict-free)
#pragma unroll
            for (int k = 0; k < 32; ++k) stg[k * TC_MCH] = from_f32<TOut>(__uint_as_float(v0[k]));
#pragma unroll
            for (int k = 32; k < PP; ++k) stg[k * TC_MCH] = from_f32<TOut>(__uint_as_float(v1[k - 32]));
            fence_proxy_async();
            named_bar_sync(2, 128);
            if (tid == 0 && !(dbg & 2)) {
              tma_store_2d(&omap, stage0 + sb * STG, mb * TC_MCH, r * PP);     // channels past C are clipped by the tensor map
              bulk_commit();
            }
            continue;
          }
          bool packed = false;
          if constexpr (sizeof(TOut) == 2) packed = !(dbg & 4);       // bit 2 of roi_fwd_dbg: the old 2-byte stores (A/B)
          if (packed) {
            // 98-byte rows: 4-byte aligned for even channels, 2 mod 4 for odd ones.  One 2-byte store (element 48 resp. 0) and 24
            // packed 4-byte stores per row instead of 49 2-byte ones: the epilogue's conflicted stores were the largest single user
            // of the shared-memory port in this kernel (2 passes x 49 per warp and block), the port the TMA ingest competes for.
#define TC_V(k) __uint_as_float((k) < 32 ? v0[(k) & 31] : v1[((k) - 32) & 31])
            const bool odd = (lane & 1) != 0;
            stg[odd ? 0 : PP - 1] = from_f32<TOut>(odd ? TC_V(0) : TC_V(PP - 1));     // (only instantiated paths with 2-byte TOut get here)
            uint32_t* stg32 = reinterpret_cast<uint32_t*>(stg + (odd ? 1 : 0));
#pragma unroll
            for (int i = 0; i < (PP - 1) / 2; ++i) {
              const float lo = odd ? TC_V(2 * i + 1) : TC_V(2 * i), hi = odd ? TC_V(2 * i + 2) : TC_V(2 * i + 1);
              const __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
              stg32[i] = *reinterpret_cast<const uint32_t*>(&pk);
            }
#undef TC_V
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) stg[k] = from_f32<TOut>(__uint_as_float(v0[k]));
#pragma unroll
            for (int k = 32; k < PP; ++k) stg[k] = from_f32<TOut>(__uint_as_float(v1[k - 32]));
          }
          fence_proxy_async();
          named_bar_sync(2, 128);
          if (tid == 0 && !(dbg & 2)) {
            const int c0 = mb * TC_MCH;
            const int nch = min(TC_MCH, C - c0);
            bulk_s2g(out + ((size_t)r * C + c0) * PP, stage0 + sb * STG, (uint32_t)(nch * PP * sizeof(TOut)));
            bulk_commit();
          }
        }
      }
    }
    if (tid == 0) bulk_wait_read0();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_NACC * TC_NB);
  }
}

template <typename TOut, bool kRHWC>
static int launch_tc(const CUtensorMap& fmap, int C, int H, int W, int R, const void* ws, void* out, cudaStream_t st) {
  const size_t smem = tc_smem_bytes<TOut>(H, W);
  DA_REQUIRE(smem <= 227 * 1024, DA_ERR_UNSUPPORTED, "roi_align tc: H+W too large for shared memory");
  CUtensorMap omap;
  memset(&omap, 0, sizeof(omap));
  if (kRHWC) {
    DA_REQUIRE((long long)R * PP <= 0x7fffffffll, DA_ERR_UNSUPPORTED, "roi_align tc: too many RoIs");
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)R * PP};
    const uint64_t strides[1] = {(uint64_t)C * sizeof(TOut)};
    const uint32_t box[2] = {TC_MCH, PP};
    int rc = encode_map_plain(&omap, out, (int)sizeof(TOut), 2, dims, strides, box);
    if (rc) return rc;
  }
  auto k = roi_align_fwd_tc_kernel<TOut, kRHWC>;
  DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // the footprint-sorted RoI order was left in the workspace by the last block of roi_prep_kernel
  const int grid = R < num_sms() ? R : num_sms();
  k<<<grid, TC_THREADS, smem, st>>>(fmap, omap, C, H, W, R, (const unsigned char*)ws, (TOut*)out, g_opt.roi_fwd_dbg);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

// Requires C % 64 == 0: the channel axis is viewed as (C/64 groups) x 64 so that one rank-4 box
// (64 ch, 4 px, 4 rows, 8 groups) fetches a quad for 512 channels.
int roi_align_fwd_tc(const void* feat, int N, int C, int H, int W, int R, const void* ws, void* out, int out_dtype, int layout,
                     cudaStream_t st) {
  CUtensorMap fmap;
  const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)N * H, (uint64_t)(C / 64)};
  const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, 128};
  const uint32_t box[4] = {64, 4, 4, 2 * TC_GB};
  int rc = encode_map(&fmap, feat, 4, dims, strides, box);
  if (rc) return rc;
  if (layout == DA_ROI_OUT_RHWC) {
    if (out_dtype == DA_BF16) return launch_tc<__nv_bfloat16, true>(fmap, C, H, W, R, ws, out, st);
    return launch_tc<float, true>(fmap, C, H, W, R, ws, out, st);
  }
  if (out_dtype == DA_BF16) return launch_tc<__nv_bfloat16, false>(fmap, C, H, W, R, ws, out, st);
  return launch_tc<float, false>(fmap, C, H, W, R, ws, out, st);
}

}  // namespace da
