// tcgen05 / TMEM / TMA implicit-GEMM engine for the domain-classifier convs and FC stacks
// (DA_ENGINE_UMMA_BF16, DA_ENGINE_UMMA_BF16X3, DA_ENGINE_UMMA_BF16X6) — sm_100a only.
//
// One warp-specialised kernel serves forward and data-gradient ("NT" form: both operands
// K-major), a second one the weight gradient ("TN" form: both operands MN-major, the
// reduction runs over pixels):
//
//   warp 0 (1 thread)  TMA producer : cp.async.bulk.tensor boxes -> 128B-swizzled smem ring
//   warp 1 (1 thread)  MMA issuer   : tcgen05.mma.cta_group::1.kind::f16, fp32 accumulator in TMEM
//   warps 2..5         epilogue     : tcgen05.ld TMEM -> registers -> fused epilogue -> global
//
// Implicit GEMM without im2col: activations are NHWC, so for one filter tap the A tile of a
// BHxBW patch of output pixels is a rank-4 TMA box (64 channels, BW, BH, 1 image) whose
// start coordinate is shifted by the tap; out-of-bounds (padding) is zero-filled by TMA.
// Stride-s convolutions use s*s "parity" tensor maps (base pointer shifted by the parity,
// pixel strides multiplied by s), so no element-stride traversal is needed.  The data
// gradient of a strided conv is run per output-parity class with the subset of taps that
// reaches that class.  GRL: the data-gradient epilogue multiplies by out_scale (= -lambda),
// so the reversed gradient is emitted by the same pass (instance_da.py:20-23).
//
// BF16X3: fp32 operands are split x = hi + lo (bf16 each) and the kernel accumulates
// hi*hi + hi*lo + lo*hi into the same TMEM accumulator (3 k-passes), ~2^-16 relative.
// BF16X6: exact 3-way split x = hi + mid + lo, six product terms (everything down to 2^-16 of the
// product; the dropped mid*lo, lo*mid, lo*lo are <= 2^-23): fp32-class results (<= 1e-5 parity bar)
// on the tensor cores at 1/6 of the bf16 rate -- the same cost as 3xTF32 (half rate x 3 passes)
// without a second operand format, swizzle layout and MN-major restriction set.
#include "da_common.cuh"
#include "da_ptx.cuh"
#include <cuda.h>
#include <string.h>

namespace da {

// ---------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------
constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NT_THREADS = 192;      // weight-gradient kernel: TMA, MMA, 4 epilogue warps
constexpr int NT_EPI_WARPS = 8;      // forward / data-gradient kernel: 8 epilogue warps (2 per TMEM lane quadrant)
constexpr int NT_FWD_THREADS = 64 + 32 * NT_EPI_WARPS;
constexpr int MAX_TAPS = 16;
// Tile configuration: BN = 256 keeps the per-MMA shared-memory traffic (A 4 KB + B 8 KB per 128 cycles)
// under the 128 B/clk SMEM port; BN = 128 serves narrow outputs.  Two TMEM accumulators (2*BN
// columns) let the epilogue of tile i overlap the MMAs of tile i+1 (persistent kernel).
// BN = 512 (r02, CTA-pair mode only): ONE accumulator of 512 columns fills the TMEM, so the epilogue is not overlapped with the
// next tile -- worth it from ~24 k-steps per tile on (measured; FC1 forward: every operand byte is delivered to the SMs 3x instead of 4x; the
// 256-wide kernel moves 9.3 TB/s L2 -> SM at 64 % tensor-pipe activity, profiles/r02_hot_kernels_ncu.md).  The pair MMA is
// still N = 256: two instructions per k-substep share the A operand.
template <int kBN> struct TileCfg {
  static constexpr int STAGES = (kBN == 512) ? 4 : (kBN == 256) ? 4 : (kBN == 128 ? 6 : 8);   // 512: stages of the PAIR layout (A + B/2)
  static constexpr int B_BYTES = kBN * BK * 2;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * (A_BYTES + (kBN == 512 ? B_BYTES / 2 : B_BYTES)) + 256 + NT_EPI_WARPS * 4096;   // + epilogue staging
  static constexpr int TMEM_COLS = (kBN == 512) ? 512 : 2 * kBN;
  static constexpr int NBUF = (kBN == 512) ? 1 : 2;
};

struct TapInfo {
  int map;  // which parity tensor map of A
  int dh, dw;
  int bk;   // K offset of this tap inside the B matrix
};

struct NtParams {
  int dbg;                  // DA_UMMA_DBG timing experiments (results are wrong when set)
  CUtensorMap a_map[3][4];  // [split part hi/mid/lo][parity]
  CUtensorMap b_map[3];     // [split part]
  TapInfo taps[MAX_TAPS];
  int num_taps, kchunks;    // k iterations per term = num_taps * kchunks
  int num_terms;
  int term_a[6], term_b[6];
  int b_mn_major;           // B tile is [k rows][n contiguous] (weights read untransposed for dgrad)
  int b_col0;               // MN-major B: column offset of n = 0 inside the B matrix (unused)
  int flat;                 // A is a flat [M,K] matrix (1x1 / FC)
  int pix_fast;             // tile order: pixel tiles fastest (the weight operand is the big one and is streamed once)
  int BH, BW, bw_shift;     // patch shape (BH*BW == 128, powers of two)
  int TH, TW;               // extent of the tile grid in (class) pixels
  int tiles_h, tiles_w;
  long long M_flat;
  // output addressing: pixel (n, i, j) of the tile grid -> y[((n*OHf + i*os+oa)*OWf + j*os+ob)*Cout + c]
  int OHf, OWf, os, oa, ob, Cout;
  const float* scale;
  const float* shift;
  int relu;
  float drop_p;
  unsigned long long seed;
  const unsigned long long* seed_ctr;
  float out_scale;
  void* y;
  int y_dtype;
  float* partial;  // split-K: fp32 [splits][numel(y)]
  long long y_numel;
};

template <typename T> struct Pack;
template <> struct Pack<__nv_bfloat16> {
  __device__ static void store8(__nv_bfloat16* p, const float* v) {
    uint4 u;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<unsigned*>(&a); u.y = *reinterpret_cast<unsigned*>(&b);
    u.z = *reinterpret_cast<unsigned*>(&c); u.w = *reinterpret_cast<unsigned*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// Fused epilogue of one 32-column chunk of a warp's 32 accumulator rows.
// Phase 1: every lane applies the epilogue to the 32 values of ITS row.  Phase 2: the warp transposes the
// [32 rows][32 cols] chunk through a private shared-memory tile (16-byte pieces, XOR-swizzled) so that
// each store instruction writes whole contiguous row segments (64 B bf16 / 128 B fp32 per row) instead
// of 32 lanes hitting 32 different cache lines with 16 B each.
// Per-channel scale / shift of a 32-column chunk, one column per lane (column c_base + lane): ONE coalesced load per chunk,
// issued early by the caller (32 broadcast loads per lane inside the epilogue stalled every chunk on the first L1 miss).
struct LaneAffine { float scale, shift; };
__device__ __forceinline__ LaneAffine nt_lane_affine(const NtParams& P, int c_base, int lane) {
  LaneAffine a{1.f, 0.f};
  const int c = c_base + lane;
  if (c < P.Cout) {
    if (P.scale) a.scale = __ldg(P.scale + c);
    if (P.shift) a.shift = __ldg(P.shift + c);
  }
  return a;
}

__device__ __forceinline__ void nt_epilogue_chunk(const NtParams& P, const uint32_t* v, size_t row_off, bool valid,
                                                  int c_base, bool raw_partial, float* partial, uint8_t* wstage,
                                                  int lane, LaneAffine aff) {
  const int ncols = min(32, P.Cout - c_base);   // warp-uniform
  if (ncols <= 0) return;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  const bool out_f32 = raw_partial || P.y_dtype == DA_F32;
  // Every optional step is ONE warp-uniform branch around its own unrolled loop.  (r01 had them as per-element `if`s inside one
  // loop; the compiler if-converted the ~45-instruction dropout hash, so every conv with a bias / BN / ReLU epilogue issued
  // 1500 predicated-off instructions per chunk: +12 us per tile, tools/probe_mk.py.)
  if (!raw_partial) {
    const float os = P.out_scale;
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= os;
    if (P.scale) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= __shfl_sync(0xffffffffu, aff.scale, j);
    }
    if (P.shift) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += __shfl_sync(0xffffffffu, aff.shift, j);
    }
    if (P.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (P.drop_p > 0.f) {
      const uint32_t thr = drop_threshold(P.drop_p);
      const float keep_scale = 1.f / (1.f - P.drop_p);
      const unsigned long long seed = effective_seed(P.seed, P.seed_ctr);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        f[j] = (drop_hash(seed, (uint64_t)(row_off + c_base + j)) >= thr) ? f[j] * keep_scale : 0.f;
    }
  }
  uint8_t* gbase = raw_partial ? reinterpret_cast<uint8_t*>(partial) : reinterpret_cast<uint8_t*>(P.y);
  const int es = out_f32 ? 4 : 2;
  const bool fast = (ncols == 32) && ((P.Cout * es) % 16 == 0) && ((c_base * es) % 16 == 0) &&
                    ((reinterpret_cast<uintptr_t>(gbase) & 15) == 0);
  if (!fast) {   // ragged tail: plain per-row stores
    if (valid) {
      // compile-time indices only: a runtime-indexed loop would move f[] (and every access to it) to local memory
      if (out_f32) {
        float* o = reinterpret_cast<float*>(gbase) + row_off + c_base;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = f[j];
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(gbase) + row_off + c_base;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = __float2bfloat16_rn(f[j]);
      }
    }
    return;
  }
  // phase 2a: own row -> staging, 16-byte piece j at slot (j ^ (row & (npieces-1)))
  const int npieces = out_f32 ? 8 : 4;          // 16-byte pieces per row
  const int row_bytes = npieces * 16;
  uint4* srow = reinterpret_cast<uint4*>(wstage + lane * row_bytes);
  if (out_f32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      srow[j ^ (lane & 7)] = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 a = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]), b = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
      __nv_bfloat162 c = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]), d = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
      srow[j ^ (lane & 3)] = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b),
                                        *reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
    }
  }
  __syncwarp();
  // phase 2b: lanes cover contiguous row segments
  const int rows_per_it = 32 / npieces;
  const int sub = lane / npieces, piece = lane % npieces;
  const unsigned long long my_off = valid ? (unsigned long long)row_off : ~0ull;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (it * rows_per_it >= 32) break;
    const int row = it * rows_per_it + sub;
    const unsigned long long roff = __shfl_sync(0xffffffffu, my_off, row);
    const uint4 val = *reinterpret_cast<const uint4*>(wstage + row * row_bytes + ((piece ^ (row & (npieces - 1))) << 4));
    if (roff != ~0ull && !(P.dbg & 1))
      *reinterpret_cast<uint4*>(gbase + (roff + c_base) * es + piece * 16) = val;
  }
  __syncwarp();
}

// Persistent kernel: grid = min(#tiles, #SMs); tile = (pixel tile, Cout tile, k-split).
// kCluster == 2: the two CTAs of a cluster work on adjacent pixel tiles of the SAME Cout tile and
// k-split; each loads half of the shared B (weight) tile and TMA-multicasts it to both, which halves
// the L2->SM weight traffic (the 128x256 tile is L2-bandwidth bound otherwise).
// k2SM (CTA pair, kCluster == 2): ONE tcgen05.mma.cta_group::2 per k-step covers the 256 pixel rows of both CTAs.
// Each CTA stages its own 128 rows of A and HALF of the B tile (16 + 16 KB per k-step instead of 16 + 32 KB), so
// the bytes every SM has to receive per MMA cycle drop by a third -- the 1-SM 128x256 tile is bound by the ~46 B/clk
// an SM can take in, not by the tensor pipe.  The leader (cluster rank 0) owns the full barriers and issues the MMAs.
// kChunked (fp32-class engine, BN = 128): the fp32 accumulator in TMEM is updated with TRUNCATION, so the error of an
// accumulation chain is a bias of ~2^-24 of the running sum per tcgen05.mma (tools/probe_precision.py,
// profiles/r02_precision_probe.txt: K = 100352 -> 1.9e-4 on same-sign data).  In this mode a chain is cut after kChunkIters
// k-steps (32 MMAs): the two TMEM buffers ping-pong per CHUNK instead of per tile, and the epilogue warps add every finished
// chunk into fp32 REGISTER accumulators (64 per thread) with round-to-nearest adds; the fused epilogue then runs from registers.
constexpr int kChunkIters = 8;

template <int kBN, int kCluster, bool k2SM, bool kChunked = false>
__global__ void __launch_bounds__(NT_FWD_THREADS, 1)
umma_nt_kernel(const __grid_constant__ NtParams P, int pixel_tiles, int n_tiles, int splits) {
  using Cfg = TileCfg<kBN>;
  static_assert(!k2SM || kCluster == 2, "the CTA-pair mode is a cluster of exactly two CTAs");
  static_assert(kBN != 512 || k2SM, "512-wide tiles exist in the CTA-pair mode only");
  static_assert(!kChunked || (kBN == 128 && !k2SM), "chunked accumulation: 128-wide tiles, one CTA per MMA");
  constexpr int NBUF = Cfg::NBUF;
  constexpr int B_BYTES = k2SM ? Cfg::B_BYTES / 2 : Cfg::B_BYTES;                       // bytes of B staged by THIS CTA per k-step
  constexpr int STAGES = (kBN == 512) ? Cfg::STAGES : (Cfg::STAGES * (A_BYTES + Cfg::B_BYTES)) / (A_BYTES + B_BYTES);  // same ring bytes, deeper ring
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + STAGES * A_BYTES;
  const uint32_t bars = b_s + STAGES * B_BYTES;
  // full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem slot
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16,
                 tslot = tempty0 + 16;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* epi_stage = gen_base + (bars - base) + 256;   // NT_EPI_WARPS x 4 KB, 16-byte aligned
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tslot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tiles are enumerated per CLUSTER: super pixel tile q covers pixel tiles q*kCluster + rank
  const int crank = (kCluster > 1) ? (int)cluster_ctarank() : 0;
  const int super_tiles = (pixel_tiles + kCluster - 1) / kCluster;
  const int total_tiles = super_tiles * n_tiles * splits;
  const int tile0 = blockIdx.x / kCluster, tile_step = gridDim.x / kCluster;
  const int total_iters = P.num_terms * P.num_taps * P.kchunks;
  const int per_split = (total_iters + splits - 1) / splits;
  const int per_img = P.tiles_h * P.tiles_w;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, k2SM ? 1 : kCluster); }
    // pair mode: the epilogue warps of BOTH CTAs release an accumulator to the leader's MMA thread
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, k2SM ? 2 * NT_EPI_WARPS : NT_EPI_WARPS); }
    fence_barrier_init();
  }
  pdl_launch_dependents();   // the next kernel of the stream may run its prologue under this one
  if (warp == 1) { if (k2SM) tmem_alloc_2sm(tslot, Cfg::TMEM_COLS); else tmem_alloc(tslot, Cfg::TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // peer barriers are initialised before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tslot_ptr;
  pdl_wait();                // everything above overlapped the predecessor; global memory is touched only below

  if (warp == 0) {
    if (lane == 0) {
      int kq = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int sp = tile % splits, rest = tile / splits;
        const int nt = P.pix_fast ? rest / super_tiles : rest % n_tiles;
        const int pt = (P.pix_fast ? rest % super_tiles : rest / n_tiles) * kCluster + crank;
        const int c0 = nt * kBN;
        int n_img = 0, i0 = 0, j0 = 0, m0 = 0;
        if (P.flat) {
          m0 = pt * BM;
        } else {
          n_img = pt / per_img;
          const int t = pt % per_img;
          i0 = (t / P.tiles_w) * P.BH;
          j0 = (t % P.tiles_w) * P.BW;
        }
        const int it_begin = sp * per_split, it_end = min(it_begin + per_split, total_iters);
        for (int it = it_begin; it < it_end; ++it, ++kq) {
          const int s = kq % STAGES;
          const uint32_t ph = (uint32_t)(kq / STAGES) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const int per_term = P.num_taps * P.kchunks;
          const int term = it / per_term, rem = it % per_term;
          // taps innermost: the shifted A boxes of one channel chunk overlap almost completely, so eight of the nine loads hit L2
          // (channel chunks innermost re-streamed the whole activation tile once per tap: SRM data gradient, 158 MB x 9)
          const int kc = rem / P.num_taps, tap = rem % P.num_taps;
          const TapInfo ti = P.taps[tap];
          const uint32_t fb = full0 + 8 * s;
          const CUtensorMap* bm = &P.b_map[P.term_b[term]];
          if constexpr (k2SM) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            if (crank == 0) mbar_expect_tx(fb, 2 * (A_BYTES + B_BYTES));
            if (P.flat)
              tma_load_2d_2sm(a_s + s * A_BYTES, &P.a_map[P.term_a[term]][0], fb, kc * BK, m0);
            else
              tma_load_4d_2sm(a_s + s * A_BYTES, &P.a_map[P.term_a[term]][ti.map], fb, kc * BK, j0 + ti.dw, i0 + ti.dh, n_img);
            // per 256-column MMA this CTA stages output channels [h2*256 + crank*128, +128) of the tile
#pragma unroll
            for (int h2 = 0; h2 < kBN / 256; ++h2) {
              if (P.b_mn_major) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  tma_load_2d_2sm(b_s + s * B_BYTES + (h2 * 2 + j) * (64 * BK * 2), bm, fb, ti.bk + c0 + h2 * 256 + crank * 128 + j * 64, kc * BK);
              } else {
                tma_load_2d_2sm(b_s + s * B_BYTES + h2 * (128 * BK * 2), bm, fb, ti.bk + kc * BK, c0 + h2 * 256 + crank * 128);
              }
            }
            continue;
          }
          mbar_expect_tx(fb, A_BYTES + B_BYTES);
          if (P.flat)
            tma_load_2d(a_s + s * A_BYTES, &P.a_map[P.term_a[term]][0], fb, kc * BK, m0);
          else
            tma_load_4d(a_s + s * A_BYTES, &P.a_map[P.term_a[term]][ti.map], fb, kc * BK, j0 + ti.dw, i0 + ti.dh, n_img);
          if (kCluster == 1) {
            if (P.b_mn_major) {
              // boxes of (64 n, 64 k-rows): row = output channel chunk kc, column = tap block + n
#pragma unroll
              for (int j = 0; j < kBN / 64; ++j)
                tma_load_2d(b_s + s * B_BYTES + j * (64 * BK * 2), bm, fb, ti.bk + c0 + j * 64, kc * BK);
            } else {
              tma_load_2d(b_s + s * B_BYTES, bm, fb, ti.bk + kc * BK, c0);
            }
          } else {
            // this CTA fetches its half of the B tile and multicasts it into both CTAs
            constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);
            constexpr int HALF = kBN / kCluster;
            if (P.b_mn_major) {
#pragma unroll
              for (int j = 0; j < HALF / 64; ++j) {
                const int jj = crank * (HALF / 64) + j;
                tma_load_2d_mc(b_s + s * B_BYTES + jj * (64 * BK * 2), bm, fb, ti.bk + c0 + jj * 64, kc * BK, kMask);
              }
            } else {
              tma_load_2d_mc(b_s + s * B_BYTES + crank * (HALF * BK * 2), bm, fb, ti.bk + kc * BK, c0 + crank * HALF, kMask);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && (!k2SM || crank == 0)) {
      const uint32_t idesc = make_idesc(k2SM ? 2 * BM : BM, kBN == 512 ? 256 : kBN, 0, P.b_mn_major ? 1 : 0);
      int kq = 0, tcount = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int sp = tile % splits;
        const int it_begin = sp * per_split, it_end = min(it_begin + per_split, total_iters);
        // one accumulator hand-over per tile -- or, chunked, per kChunkIters k-steps (tcount counts hand-overs)
        int buf = tcount % NBUF;
        mbar_wait(tempty0 + 8 * buf, (((uint32_t)(tcount / NBUF)) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        uint32_t d_tmem = tmem_base + buf * kBN;
        for (int it = it_begin; it < it_end; ++it, ++kq) {
          const int rel = it - it_begin;
          if (kChunked && rel > 0 && rel % kChunkIters == 0) {     // hand the finished chunk over, start the next chain at zero
            umma_commit(tfull0 + 8 * buf);
            ++tcount;
            buf = tcount & 1;
            mbar_wait(tempty0 + 8 * buf, (((uint32_t)(tcount >> 1)) & 1u) ^ 1u);
            tc_fence_after();
            d_tmem = tmem_base + buf * kBN;
          }
          const bool fresh = kChunked ? (rel % kChunkIters == 0) : (it == it_begin);
          const int s = kq % STAGES;
          const uint32_t ph = (uint32_t)(kq / STAGES) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t ad = desc_kmajor_sw128(a_s + s * A_BYTES + kk * 32);
            const uint64_t bd = P.b_mn_major ? desc_mnmajor_sw128(b_s + s * B_BYTES + kk * 2048, 64 * BK * 2)
                                             : desc_kmajor_sw128(b_s + s * B_BYTES + kk * 32);
            if (kBN == 512) {     // two N = 256 pair MMAs share the A operand
              umma_bf16_2sm(d_tmem, ad, bd, idesc, (!fresh || kk > 0) ? 1u : 0u);
              const uint64_t bd2 = P.b_mn_major ? desc_mnmajor_sw128(b_s + s * B_BYTES + 2 * (64 * BK * 2) + kk * 2048, 64 * BK * 2)
                                                : desc_kmajor_sw128(b_s + s * B_BYTES + 128 * BK * 2 + kk * 32);
              umma_bf16_2sm(d_tmem + 256, ad, bd2, idesc, (!fresh || kk > 0) ? 1u : 0u);
            } else if (k2SM) umma_bf16_2sm(d_tmem, ad, bd, idesc, (!fresh || kk > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, ad, bd, idesc, (!fresh || kk > 0) ? 1u : 0u);
          }
          // the stage is reusable only when BOTH CTAs are done with it (multicast writes into both)
          if (k2SM) umma_commit_2sm(empty0 + 8 * s, 3);
          else if (kCluster == 1) umma_commit(empty0 + 8 * s);
          else umma_commit_mc(empty0 + 8 * s, (uint16_t)((1u << kCluster) - 1u));
        }
        if (k2SM) umma_commit_2sm(tfull0 + 8 * buf, 3);   // each CTA's epilogue drains its own 128 accumulator rows
        else umma_commit(tfull0 + 8 * buf);
        ++tcount;
      }
    }
  } else {
    // epilogue warps 2..9 -> TMEM lane quadrant (warp % 4); the two warps of a quadrant split the column chunks
    const int q = warp & 3;
    const int ehalf = (warp - 2) >> 2;
    const int r = q * 32 + lane;  // accumulator row == pixel within the tile
    int tcount = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int sp = tile % splits, rest = tile / splits;
      const int nt = P.pix_fast ? rest / super_tiles : rest % n_tiles;
      const int pt = (P.pix_fast ? rest % super_tiles : rest / n_tiles) * kCluster + crank;
      const int c0 = nt * kBN;
      const int it_begin = sp * per_split, it_end = min(it_begin + per_split, total_iters);
      const bool has_k = it_end > it_begin;
      bool valid;
      size_t row_off;
      if (P.flat) {
        const long long m = (long long)pt * BM + r;
        valid = m < P.M_flat;
        row_off = (size_t)m * P.Cout;
      } else {
        const int n_img = pt / per_img, t = pt % per_img;
        const int i = (t / P.tiles_w) * P.BH + (r >> P.bw_shift), j = (t % P.tiles_w) * P.BW + (r & (P.BW - 1));
        valid = (i < P.TH) && (j < P.TW) && (pt < pixel_tiles);   // phantom tile of an odd cluster tail
        row_off = (((size_t)n_img * P.OHf + (size_t)i * P.os + P.oa) * P.OWf + (size_t)j * P.os + P.ob) * P.Cout;
      }
      const bool raw = splits > 1;
      float* partial = raw ? P.partial + (size_t)sp * P.y_numel : nullptr;
      if constexpr (kChunked) {
        // this warp owns columns [ehalf*32, +32) and [(ehalf+2)*32, +32) of its 32 rows: 64 fp32 register accumulators
        float a0[32], a1[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
        const int nchunks = has_k ? (it_end - it_begin + kChunkIters - 1) / kChunkIters : 1;
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c, ++tcount) {
          const int buf = tcount & 1;
          mbar_wait(tfull0 + 8 * buf, ((uint32_t)(tcount >> 1)) & 1u);
          tc_fence_after();
          if (has_k) {
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN;
            DA_TMEM_LD32(taddr + ehalf * 32, v0);
            DA_TMEM_LD32(taddr + (ehalf + 2) * 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { a0[j] = __fadd_rn(a0[j], __uint_as_float(v0[j])); a1[j] = __fadd_rn(a1[j], __uint_as_float(v1[j])); }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
        }
        uint32_t u[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(a0[j]);
        nt_epilogue_chunk(P, u, row_off, valid, c0 + ehalf * 32, raw, partial, epi_stage + (warp - 2) * 4096, lane,
                          nt_lane_affine(P, c0 + ehalf * 32, lane));
#pragma unroll
        for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(a1[j]);
        nt_epilogue_chunk(P, u, row_off, valid, c0 + (ehalf + 2) * 32, raw, partial, epi_stage + (warp - 2) * 4096, lane,
                          nt_lane_affine(P, c0 + (ehalf + 2) * 32, lane));
        continue;
      }
      const int buf = tcount % NBUF;
      const uint32_t use = (uint32_t)(tcount / NBUF);
      ++tcount;
      LaneAffine aff = nt_lane_affine(P, c0 + ehalf * 32, lane);      // first chunk's scale / shift: in flight during the wait
      mbar_wait(tfull0 + 8 * buf, use & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cc = ehalf; cc < kBN / 32; cc += NT_EPI_WARPS / 4) {
        const LaneAffine aff_next = nt_lane_affine(P, c0 + (cc + NT_EPI_WARPS / 4) * 32, lane);   // next chunk's, under this chunk's work
        uint32_t v[32];
        if (has_k) {
          DA_TMEM_LD32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN + cc * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (!(P.dbg & 2)) nt_epilogue_chunk(P, v, row_off, valid, c0 + cc * 32, raw, partial, epi_stage + (warp - 2) * 4096, lane, aff);
        aff = aff_next;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (k2SM) mbar_arrive_cluster(tempty0 + 8 * buf, 0); else mbar_arrive(tempty0 + 8 * buf); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // no CTA exits while its peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    if (k2SM) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// split-K finish: y = epilogue(sum_s partial[s])
__global__ void nt_splitk_finish_kernel(const float* __restrict__ partial, int splits, long long numel, int Cout,
                                        const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                        float drop_p, unsigned long long seed0, const unsigned long long* seed_ctr,
                                        float out_scale, void* y, int y_dtype) {
  pdl_launch_dependents();
  const unsigned long long seed = effective_seed(seed0, seed_ctr);
  const uint32_t thr = drop_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(size_t)s * numel + i];
    const int c = (int)(i % Cout);
    float x = acc * out_scale;
    if (scale) x *= scale[c];
    if (shift) x += shift[c];
    if (relu) x = fmaxf(x, 0.f);
    if (drop_p > 0.f) x = (drop_hash(seed, (uint64_t)i) >= thr) ? x * keep_scale : 0.f;
    if (y_dtype == DA_BF16) reinterpret_cast<__nv_bfloat16*>(y)[i] = __float2bfloat16_rn(x);
    else reinterpret_cast<float*>(y)[i] = x;
  }
}

// four outputs per thread (numel % 4 == 0, Cout % 4 == 0, 16-byte aligned buffers): 16-byte loads of every split's partial
__global__ void nt_splitk_finish_vec4_kernel(const float4* __restrict__ partial, int splits, long long n4, int Cout,
                                             const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                             float drop_p, unsigned long long seed0, const unsigned long long* seed_ctr,
                                             float out_scale, void* y, int y_dtype) {
  pdl_launch_dependents();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 p = partial[(size_t)s * n4 + i];
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    float x[4] = {acc.x * out_scale, acc.y * out_scale, acc.z * out_scale, acc.w * out_scale};
    const int c = (int)((i * 4) % Cout);
    if (scale) { const float4 v = *reinterpret_cast<const float4*>(scale + c); x[0] *= v.x; x[1] *= v.y; x[2] *= v.z; x[3] *= v.w; }
    if (shift) { const float4 v = *reinterpret_cast<const float4*>(shift + c); x[0] += v.x; x[1] += v.y; x[2] += v.z; x[3] += v.w; }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.f);
    }
    if (drop_p > 0.f) {      // one uniform branch (see nt_epilogue_chunk)
      const unsigned long long seed = effective_seed(seed0, seed_ctr);
      const uint32_t thr = drop_threshold(drop_p);
      const float keep_scale = 1.f / (1.f - drop_p);
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = (drop_hash(seed, (uint64_t)(i * 4 + j)) >= thr) ? x[j] * keep_scale : 0.f;
    }
    if (y_dtype == DA_BF16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
      reinterpret_cast<uint2*>(y)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    } else {
      reinterpret_cast<float4*>(y)[i] = make_float4(x[0], x[1], x[2], x[3]);
    }
  }
}

// Warp-cooperative store of a [32 rows][32 fp32] accumulator chunk: lane L holds row L.  The chunk is transposed
// through a private 4 KB shared-memory tile (16-byte pieces, XOR-swizzled) so that every store instruction
// writes four whole 128-byte row segments instead of 32 lanes hitting 32 different lines with 16 B each.
// `row_ptr` = this lane's destination row (nullptr: row not stored); must be 16-byte aligned.
__device__ __forceinline__ void warp_store_rows_f32(const uint32_t* v, float* row_ptr, uint8_t* wstage, int lane) {
  uint4* srow = reinterpret_cast<uint4*>(wstage + lane * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int sub = lane >> 3, piece = lane & 7;
  const unsigned long long mine = reinterpret_cast<unsigned long long>(row_ptr);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + sub;
    const unsigned long long dst = __shfl_sync(0xffffffffu, mine, row);
    const uint4 val = *reinterpret_cast<const uint4*>(wstage + row * 128 + ((piece ^ (row & 7)) << 4));
    if (dst) *reinterpret_cast<uint4*>(dst + piece * 16) = val;
  }
  __syncwarp();
}

// Same data path as warp_store_rows_f32, but the 32x32 gradient chunk never reaches memory: the lanes that would store a
// 16-byte piece of a row load the matching pieces of the fp32 master and of the momentum buffer (all 16 loads of a lane
// are issued before the first use: 8 KB in flight per warp), apply torch.optim.SGD's rule with the operation order of
// sgd_step_kernel (bit-identical results) and store master, momentum and the bf16 operand copy.
// `w_row` = this lane's row in the master (nullptr: row outside the tensor).
__device__ __forceinline__ void warp_sgd_rows(const uint32_t* v, float* w_row, uint8_t* wstage, int lane, float* w_base,
                                              float* buf_base, __nv_bfloat16* shadow_base, float lr, float mu, float wd, int first) {
  uint4* srow = reinterpret_cast<uint4*>(wstage + lane * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int sub = lane >> 3, piece = lane & 7;
  const unsigned long long mine = reinterpret_cast<unsigned long long>(w_row);
  float4 wv[8], bv[8];
  unsigned long long dst[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + sub;
    dst[it] = __shfl_sync(0xffffffffu, mine, row);
    if (dst[it]) {
      dst[it] += piece * 16;
      wv[it] = *reinterpret_cast<const float4*>(dst[it]);
      bv[it] = first ? make_float4(0.f, 0.f, 0.f, 0.f)
                     : *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(buf_base) + (dst[it] - reinterpret_cast<unsigned long long>(w_base)));
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + sub;
    const float4 gv = *reinterpret_cast<const float4*>(wstage + row * 128 + ((piece ^ (row & 7)) << 4));
    if (dst[it]) {
      float* wp = &wv[it].x; float* bp = &bv[it].x; const float* gp = &gv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = fmaf(wd, wp[j], gp[j]);
        bp[j] = first ? d : fmaf(mu, bp[j], d);
        wp[j] = fmaf(-lr, bp[j], wp[j]);
      }
      const unsigned long long off = dst[it] - reinterpret_cast<unsigned long long>(w_base);
      *reinterpret_cast<float4*>(dst[it]) = wv[it];
      *reinterpret_cast<float4*>(reinterpret_cast<char*>(buf_base) + off) = bv[it];
      if (shadow_base) {
        __nv_bfloat162 a = __floats2bfloat162_rn(wv[it].x, wv[it].y), b = __floats2bfloat162_rn(wv[it].z, wv[it].w);
        *reinterpret_cast<uint2*>(reinterpret_cast<char*>(shadow_base) + (off >> 1)) =
            make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------
// weight gradient ("TN"): dW[co, tap, ci] = sum_pix dZ[pix, co] * X[pix_in(tap), ci]
// ---------------------------------------------------------------------------------------
constexpr int WK = 64;  // pixels per k-step
constexpr int W_A_BYTES = BM * WK * 2;

struct TnParams {
  CUtensorMap a_map[3];     // dZ [split part]: (Cout, OW, OH, N) or flat (Cout, M)
  CUtensorMap b_map[3][4];  // X  [split part][parity]
  TapInfo taps[MAX_TAPS];
  int num_taps, num_terms;
  int term_a[6], term_b[6];
  int flat;
  int BH, BW;               // pixel patch per k-step (BH*BW == 64)
  int tiles_h, tiles_w, NB; // patch grid
  long long M_flat;
  int Cout, Cin;
  float* dw;                // [Cout, taps, Cin] fp32 (or partial [splits][...])
  long long dw_numel;
  // fused optimizer (da_conv_backward_weight_sgd): when sgd_w != nullptr the epilogue applies the SGD rule to the fp32
  // master / momentum / bf16 operand copy at the tile's addresses instead of storing the gradient (same memory order as dw)
  float* sgd_w;
  float* sgd_buf;
  __nv_bfloat16* sgd_shadow;
  float sgd_lr, sgd_mu, sgd_wd;
  int sgd_first;
};

// Persistent kernel, one CTA per SM.  Tile = (Cout tile, Cin tile, tap, pixel split); tiles are enumerated per
// CLUSTER: the kCluster CTAs of a cluster take adjacent Cout tiles of the SAME (Cin tile, tap, split), so they
// share the X (B operand) tile: each CTA fetches 1/kCluster of it and TMA-multicasts it to all of them, which
// cuts the L2->SM traffic of the 128x256 tile from 48 KB to 16 + 32/kCluster KB per k-step (the kernel was
// L2-bandwidth bound at 505 TFLOP/s).  Two TMEM accumulators: the epilogue of tile t (128 KB of fp32 stores)
// runs under the main loop of tile t+1.
// k2SM (r02): the two CTAs of a cluster form ONE cta_group::2 MMA over 256 output channels: each stages its own 128 Cout rows of
// dZ and HALF of the X tile (16 + 16 KB per k-step instead of 16 + 32 KB received through multicast), so the bytes every SM
// has to take in per MMA drop by a third -- the same reasoning as the forward kernel's pair mode.  Same tile assignment as the
// multicast cluster (two Cout tiles sharing one X tile); the leader owns the full barriers and issues.
template <int kBN, int kCluster, bool k2SM = false>
__global__ void __launch_bounds__(NT_FWD_THREADS, 1)
umma_tn_kernel(const __grid_constant__ TnParams P, int co_tiles, int ci_tiles, int splits) {
  static_assert(!k2SM || kCluster == 2, "the CTA-pair mode is a cluster of exactly two CTAs");
  static_assert(kBN != 512 || k2SM, "512-wide tiles exist in the CTA-pair mode only");
  constexpr int NBUF = TileCfg<kBN>::NBUF;                           // 512-wide: ONE accumulator fills the TMEM (no epilogue overlap)
  constexpr int W_B_FULL = kBN * WK * 2;
  constexpr int W_B_BYTES = k2SM ? W_B_FULL / 2 : W_B_FULL;          // bytes of X staged in THIS CTA per k-step
  constexpr int STAGES = (kBN == 512) ? TileCfg<kBN>::STAGES
                                      : (TileCfg<kBN>::STAGES * (W_A_BYTES + W_B_FULL)) / (W_A_BYTES + W_B_BYTES);   // same ring bytes, deeper ring
  constexpr int BN = kBN;
  constexpr int B_BOXES = BN / 64, B_PER_CTA = B_BOXES / kCluster;
  static_assert(B_BOXES % kCluster == 0, "cluster size must divide the 64-channel boxes of the B tile");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + STAGES * W_A_BYTES;
  const uint32_t bars = b_s + STAGES * W_B_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16,
                 tslot = tempty0 + 16;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* epi_stage = gen_base + (bars - base) + 256;   // NT_EPI_WARPS x 4 KB, 16-byte aligned
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tslot - base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int crank = (kCluster > 1) ? (int)cluster_ctarank() : 0;
  const int co_super = (co_tiles + kCluster - 1) / kCluster;
  const long long total_tiles = (long long)co_super * ci_tiles * P.num_taps * splits;
  const int tile0 = blockIdx.x / kCluster, tile_step = gridDim.x / kCluster;
  const long long patches = P.flat ? (P.M_flat + WK - 1) / WK : (long long)P.NB * P.tiles_h * P.tiles_w;
  const long long total_iters = patches * P.num_terms;
  const long long per_split = (total_iters + splits - 1) / splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, k2SM ? 1 : kCluster); }
    // pair mode: the epilogue warps of BOTH CTAs release an accumulator to the leader's MMA thread
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, k2SM ? 2 * NT_EPI_WARPS : NT_EPI_WARPS); }
    fence_barrier_init();
  }
  pdl_launch_dependents();   // the next kernel of the stream may run its prologue under this one
  if (warp == 1) { if (k2SM) tmem_alloc_2sm(tslot, TileCfg<kBN>::TMEM_COLS); else tmem_alloc(tslot, 2 * BN); }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // peer barriers are initialised before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tslot_ptr;
  pdl_wait();                // everything above overlapped the predecessor; global memory is touched only below

  // tile index -> (co tile of this CTA, ci tile, tap, split); co is the fastest index so that concurrently
  // running clusters read the same X tiles (L2 reuse)
#define DA_TN_DECODE(tile)                                                        \
  const int cs_ = (int)((tile) % co_super);                                        \
  const long long r1_ = (tile) / co_super;                                         \
  const int cit = (int)(r1_ % ci_tiles);                                           \
  const int z_ = (int)(r1_ / ci_tiles);                                            \
  const int tap = z_ % P.num_taps, split = z_ / P.num_taps;                        \
  const int co0 = (cs_ * kCluster + crank) * BM, ci0 = cit * BN;                   \
  const long long it_begin = split * per_split;                                    \
  const long long it_end = (it_begin + per_split < total_iters) ? it_begin + per_split : total_iters; \
  const int n_iters = (int)((it_end > it_begin) ? it_end - it_begin : 0);

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);
      int kq = 0;
      for (long long tile = tile0; tile < total_tiles; tile += tile_step) {
        DA_TN_DECODE(tile)
        const TapInfo ti = P.taps[tap];
        for (int k = 0; k < n_iters; ++k, ++kq) {
          const long long it = it_begin + k;
          const int s = kq % STAGES;
          const uint32_t ph = (uint32_t)(kq / STAGES) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const int term = (int)(it / patches);
          const long long patch = it % patches;
          const uint32_t fb = full0 + 8 * s;
          const uint32_t ad = a_s + s * W_A_BYTES, bd = b_s + s * W_B_BYTES;
          if constexpr (k2SM) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            if (crank == 0) mbar_expect_tx(fb, 2 * (W_A_BYTES + W_B_BYTES));
            if (P.flat) {
              const int m = (int)(patch * WK);
              tma_load_2d_2sm(ad, &P.a_map[P.term_a[term]], fb, co0, m);
              tma_load_2d_2sm(ad + W_A_BYTES / 2, &P.a_map[P.term_a[term]], fb, co0 + 64, m);
#pragma unroll
              for (int j = 0; j < B_PER_CTA; ++j)   // box j: 256-column MMA j/2, this CTA's 128 columns of it, 64-column half j%2
                tma_load_2d_2sm(bd + j * (64 * WK * 2), &P.b_map[P.term_b[term]][0], fb, ci0 + (j >> 1) * 256 + crank * 128 + (j & 1) * 64, m);
            } else {
              const int per_img = P.tiles_h * P.tiles_w;
              const int n = (int)(patch / per_img), t = (int)(patch % per_img);
              const int i0 = (t / P.tiles_w) * P.BH, j0 = (t % P.tiles_w) * P.BW;
              tma_load_4d_2sm(ad, &P.a_map[P.term_a[term]], fb, co0, j0, i0, n);
              tma_load_4d_2sm(ad + W_A_BYTES / 2, &P.a_map[P.term_a[term]], fb, co0 + 64, j0, i0, n);
#pragma unroll
              for (int j = 0; j < B_PER_CTA; ++j)
                tma_load_4d_2sm(bd + j * (64 * WK * 2), &P.b_map[P.term_b[term]][ti.map], fb, ci0 + (j >> 1) * 256 + crank * 128 + (j & 1) * 64,
                                j0 + ti.dw, i0 + ti.dh, n);
            }
            continue;
          }
          mbar_expect_tx(fb, W_A_BYTES + W_B_BYTES);
          if (P.flat) {
            const int m = (int)(patch * WK);
            tma_load_2d(ad, &P.a_map[P.term_a[term]], fb, co0, m);
            tma_load_2d(ad + W_A_BYTES / 2, &P.a_map[P.term_a[term]], fb, co0 + 64, m);
#pragma unroll
            for (int j = 0; j < B_PER_CTA; ++j) {
              const int jj = crank * B_PER_CTA + j;
              if (kCluster == 1) tma_load_2d(bd + jj * (64 * WK * 2), &P.b_map[P.term_b[term]][0], fb, ci0 + jj * 64, m);
              else tma_load_2d_mc(bd + jj * (64 * WK * 2), &P.b_map[P.term_b[term]][0], fb, ci0 + jj * 64, m, kMask);
            }
          } else {
            const int per_img = P.tiles_h * P.tiles_w;
            const int n = (int)(patch / per_img), t = (int)(patch % per_img);
            const int i0 = (t / P.tiles_w) * P.BH, j0 = (t % P.tiles_w) * P.BW;
            tma_load_4d(ad, &P.a_map[P.term_a[term]], fb, co0, j0, i0, n);
            tma_load_4d(ad + W_A_BYTES / 2, &P.a_map[P.term_a[term]], fb, co0 + 64, j0, i0, n);
#pragma unroll
            for (int j = 0; j < B_PER_CTA; ++j) {
              const int jj = crank * B_PER_CTA + j;
              if (kCluster == 1)
                tma_load_4d(bd + jj * (64 * WK * 2), &P.b_map[P.term_b[term]][ti.map], fb, ci0 + jj * 64, j0 + ti.dw, i0 + ti.dh, n);
              else
                tma_load_4d_mc(bd + jj * (64 * WK * 2), &P.b_map[P.term_b[term]][ti.map], fb, ci0 + jj * 64, j0 + ti.dw, i0 + ti.dh, n, kMask);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && (!k2SM || crank == 0)) {
      constexpr uint32_t idesc = make_idesc(k2SM ? 2 * BM : BM, kBN == 512 ? 256 : BN, 1, 1);
      int kq = 0, tcount = 0;
      for (long long tile = tile0; tile < total_tiles; tile += tile_step, ++tcount) {
        DA_TN_DECODE(tile)
        (void)co0; (void)ci0; (void)tap;
        const int buf = tcount % NBUF;
        mbar_wait(tempty0 + 8 * buf, (((uint32_t)(tcount / NBUF)) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int k = 0; k < n_iters; ++k, ++kq) {
          const int s = kq % STAGES;
          const uint32_t ph = (uint32_t)(kq / STAGES) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < WK / 16; ++kk) {
            // 16 k-rows = 2 swizzle atoms of 1024 B
            const uint64_t adsc = desc_mnmajor_sw128(a_s + s * W_A_BYTES + kk * 2048, W_A_BYTES / 2);
            const uint64_t bdsc = desc_mnmajor_sw128(b_s + s * W_B_BYTES + kk * 2048, 64 * WK * 2);
            if (kBN == 512) {   // two N = 256 pair MMAs share the dZ operand; the second reads boxes 2, 3 of this CTA's X stage
              umma_bf16_2sm(d_tmem, adsc, bdsc, idesc, (k > 0 || kk > 0) ? 1u : 0u);
              const uint64_t bdsc2 = desc_mnmajor_sw128(b_s + s * W_B_BYTES + 2 * (64 * WK * 2) + kk * 2048, 64 * WK * 2);
              umma_bf16_2sm(d_tmem + 256, adsc, bdsc2, idesc, (k > 0 || kk > 0) ? 1u : 0u);
            } else if (k2SM) umma_bf16_2sm(d_tmem, adsc, bdsc, idesc, (k > 0 || kk > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, adsc, bdsc, idesc, (k > 0 || kk > 0) ? 1u : 0u);
          }
          // the stage is reusable only when EVERY CTA of the cluster is done with it (multicast writes into all / the pair
          // MMA reads both CTAs' shared memory)
          if (k2SM) umma_commit_2sm(empty0 + 8 * s, 3);
          else if (kCluster == 1) umma_commit(empty0 + 8 * s);
          else umma_commit_mc(empty0 + 8 * s, (uint16_t)((1u << kCluster) - 1u));
        }
        if (n_iters > 0) {
          if (k2SM) umma_commit_2sm(tfull0 + 8 * buf, 3);   // each CTA's epilogue drains its own 128 accumulator rows
          else umma_commit(tfull0 + 8 * buf);
        } else {
          mbar_arrive(tfull0 + 8 * buf);
          if (k2SM) mbar_arrive_cluster(tfull0 + 8 * buf, 1);
        }
      }
    }
  } else {
    // epilogue warps 2..9 -> TMEM lane quadrant (warp % 4); the two warps of a quadrant split the column chunks
    const int q = warp & 3, ehalf = (warp - 2) >> 2;
    int tcount = 0;
    for (long long tile = tile0; tile < total_tiles; tile += tile_step, ++tcount) {
      DA_TN_DECODE(tile)
      const int buf = tcount % NBUF;
      const int co = co0 + q * 32 + lane;
      float* out = P.dw + (size_t)split * P.dw_numel;
      mbar_wait(tfull0 + 8 * buf, ((uint32_t)(tcount / NBUF)) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cc = ehalf; cc < BN / 32; cc += NT_EPI_WARPS / 4) {
        uint32_t v[32];
        if (n_iters > 0) {
          DA_TMEM_LD32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + cc * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int cb = ci0 + cc * 32;
        if (cb < P.Cin) {   // warp-uniform
          if (P.sgd_w) {    // fused optimizer (launcher guarantees Cin % 32 == 0, splits == 1, 16-byte aligned tensors)
            float* wr = (co < P.Cout) ? P.sgd_w + ((size_t)co * P.num_taps + tap) * P.Cin + cb : nullptr;
            warp_sgd_rows(v, wr, epi_stage + (warp - 2) * 4096, lane, P.sgd_w, P.sgd_buf, P.sgd_shadow, P.sgd_lr, P.sgd_mu,
                          P.sgd_wd, P.sgd_first);
            continue;
          }
          float* o = (co < P.Cout) ? out + ((size_t)co * P.num_taps + tap) * P.Cin + cb : nullptr;
          const int ncols = min(32, P.Cin - cb);
          if (ncols == 32 && (P.Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            warp_store_rows_f32(v, o, epi_stage + (warp - 2) * 4096, lane);
          } else if (o) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = __uint_as_float(v[j]);   // compile-time indices keep v[] in registers
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (k2SM) mbar_arrive_cluster(tempty0 + 8 * buf, 0); else mbar_arrive(tempty0 + 8 * buf); }
    }
  }
#undef DA_TN_DECODE
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    if (k2SM) tmem_dealloc_2sm(tmem_base, TileCfg<kBN>::TMEM_COLS); else tmem_dealloc(tmem_base, 2 * BN);
  }
}

__global__ void sum_splits_vec4_kernel(const float4* __restrict__ part, int splits, long long n4, float4* __restrict__ out) {
  pdl_launch_dependents();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 p = part[(size_t)s * n4 + i];
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    out[i] = acc;
  }
}
__global__ void sum_splits_kernel(const float* __restrict__ part, int splits, long long n, float* __restrict__ out) {
  pdl_launch_dependents();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[(size_t)s * n + i];
    out[i] = acc;
  }
}

// [O][T][I] -> [I][T][O] (weights for the data gradient), with optional hi/lo split
template <typename TS>
__global__ void weight_oti_to_ito_kernel(const TS* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo, int O, int T, int I) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int o = o0 + j, i = i0 + threadIdx.x;
    if (o < O && i < I) tile[j][threadIdx.x] = to_f32<TS>(src[((size_t)o * T + t) * I + i]);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int i = i0 + j, o = o0 + threadIdx.x;
    if (o < O && i < I) {
      const float v = tile[threadIdx.x][j];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const size_t d = ((size_t)i * T + t) * O + o;
      hi[d] = h;
      if (lo) lo[d] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

__global__ void cast_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void split_hi_lo_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// 3-way split x = hi + mid + lo (bf16 each, 8 significant bits apiece: EXACT for every normal fp32 value)
__global__ void split_hi_mid_lo_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);          // exact
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);         // exact
    hi[i] = h;
    mid[i] = m;
    lo[i] = __float2bfloat16_rn(r2);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 tensor map, innermost dim first.  rank 2 or 4.
int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  EncodeTiledFn fn = get_encode();
  DA_REQUIRE(fn != nullptr, DA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  DA_REQUIRE(((uintptr_t)base & 15) == 0, DA_ERR_INVALID_ARG, "tensor map base %p is not 16-byte aligned", base);
  for (int i = 0; i < rank - 1; ++i)
    DA_REQUIRE((gs[i] & 15) == 0, DA_ERR_UNSUPPORTED, "tensor map stride %llu is not a multiple of 16 bytes (channel counts must be multiples of 8)", (unsigned long long)gs[i]);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DA_REQUIRE(r == CUDA_SUCCESS, DA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DA_OK;
}

int encode_map_plain(CUtensorMap* m, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  EncodeTiledFn fn = get_encode();
  DA_REQUIRE(fn != nullptr, DA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  DA_REQUIRE(elem_bytes == 2 || elem_bytes == 4, DA_ERR_INVALID_ARG, "tensor map: element size %d", elem_bytes);
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  DA_REQUIRE(((uintptr_t)base & 15) == 0, DA_ERR_INVALID_ARG, "tensor map base %p is not 16-byte aligned", base);
  for (int i = 0; i < rank - 1; ++i)
    DA_REQUIRE((gs[i] & 15) == 0, DA_ERR_UNSUPPORTED, "tensor map stride %llu is not a multiple of 16 bytes", (unsigned long long)gs[i]);
  CUresult r = fn(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                  const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DA_REQUIRE(r == CUDA_SUCCESS, DA_ERR_CUDA, "cuTensorMapEncodeTiled (plain) failed with CUresult %d", (int)r);
  return DA_OK;
}

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static inline int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }
static inline int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

// Pick the BHxBW patch (BH*BW == pixels, powers of two) that wastes the fewest tile pixels.
static void pick_patch(int TH, int TW, int pixels, int* BH, int* BW) {
  long long best = -1;
  for (int bw = 1; bw <= pixels; bw <<= 1) {
    const int bh = pixels / bw;
    if (bw > 256 || bh > 256) continue;
    const long long cover = (long long)((TH + bh - 1) / bh) * bh * ((TW + bw - 1) / bw) * bw;
    if (best < 0 || cover < best || (cover == best && bw > *BW)) { best = cover; *BH = bh; *BW = bw; }
  }
}

// Parity tensor maps of an NHWC activation tensor [N,H,W,C] for a stride-s gather:
// map (pa,pb): pixel (i,j) -> x[n, s*i+pa, s*j+pb, :].  box = (64 ch, BW, BH, 1).
static int make_parity_maps(CUtensorMap* maps, bool* present, const __nv_bfloat16* x, int N, int H, int W, int C, int s,
                            int BH, int BW) {
  for (int pa = 0; pa < s; ++pa)
    for (int pb = 0; pb < s; ++pb) {
      const int idx = pa * s + pb;
      const int Hp = (H - pa + s - 1) / s, Wp = (W - pb + s - 1) / s;
      present[idx] = (Hp > 0 && Wp > 0);
      if (!present[idx]) continue;
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)N};
      const uint64_t strides[3] = {(uint64_t)s * C * 2, (uint64_t)s * W * C * 2, (uint64_t)H * W * C * 2};
      const uint32_t box[4] = {64, (uint32_t)BW, (uint32_t)BH, 1};
      int rc = encode_map(&maps[idx], x + ((size_t)pa * W + pb) * C, 4, dims, strides, box);
      if (rc) return rc;
    }
  return DA_OK;
}

struct Geom {
  int N, H, W, Cin, Cout, KH, KW, s, p, OH, OW;
};
static Geom geom_of(const da_conv_desc* d) {
  Geom g{d->N, d->H, d->W, d->Cin, d->Cout, d->KH, d->KW, d->stride, d->pad, 0, 0};
  g.OH = (d->H + 2 * d->pad - d->KH) / d->stride + 1;
  g.OW = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  return g;
}
static inline bool is_flat(const Geom& g) { return g.KH == 1 && g.KW == 1 && g.s == 1 && g.p == 0; }

static int pick_splits(long long ctas, int k_iters) {
  const int sms = num_sms();
  if (ctas >= sms || k_iters < 8) return 1;
  int s = (int)((sms + ctas - 1) / ctas);
  if (s > k_iters / 16) s = k_iters / 16;
  if (s > 16) s = 16;
  return s < 1 ? 1 : s;
}

// ---- workspace layout -------------------------------------------------------------------
// [0] fp32 split-K partials / wgrad partials   (part_bytes)
// [1] operand staging: bf16 hi/lo copies of x (or dz) and of the weights (stage_bytes)
static size_t part_bytes(const Geom& g) {
  const size_t y = (size_t)g.N * g.OH * g.OW * g.Cout, x = (size_t)g.N * g.H * g.W * g.Cin;
  const size_t w = (size_t)g.Cout * g.KH * g.KW * g.Cin;
  size_t m = y > x ? y : x;
  if (w > m) m = w;
  // split-K only runs when fewer than one wave of 128x128 tiles exists, so the tensor that
  // is split never exceeds ~160 tiles
  const size_t cap = (size_t)160 * BM * 256;
  if (m > cap) m = cap;
  return align_up(m * 16 * sizeof(float), 256);
}
static size_t stage_bytes(const Geom& g) {
  const size_t y = (size_t)g.N * g.OH * g.OW * g.Cout, x = (size_t)g.N * g.H * g.W * g.Cin;
  const size_t w = (size_t)g.Cout * g.KH * g.KW * g.Cin;
  // worst case (wgrad BF16X6): hi+mid+lo of x and of dz; forward/dgrad: 3 parts of one activation + 3 parts of the weights
  return align_up(3 * 2 * (x + y) + 3 * 2 * w + 4096, 256);
}
size_t umma_workspace_bytes(const da_conv_desc* d) {
  if (d->engine == DA_ENGINE_SIMT_F32) return 0;
  const Geom g = geom_of(d);
  return part_bytes(g) + stage_bytes(g);
}

static int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

// Prepare the bf16 operand part(s) of the caller's tensor: parts[0] = hi (BF16), + lo (BF16X3), + mid, lo (BF16X6).
struct Parts {
  const __nv_bfloat16* p[3];
  int n;
};
static int prep_operand(const void* src, int dtype, long long n, int engine, uint8_t*& stage, Parts* out, cudaStream_t st) {
  out->p[0] = out->p[1] = out->p[2] = nullptr;
  out->n = 1;
  if (dtype == DA_BF16) {
    DA_REQUIRE(engine == DA_ENGINE_UMMA_BF16, DA_ERR_UNSUPPORTED, "the split-precision engines need fp32 operands");
    out->p[0] = (const __nv_bfloat16*)src;
    return DA_OK;
  }
  const int nparts = engine == DA_ENGINE_UMMA_BF16X6 ? 3 : (engine == DA_ENGINE_UMMA_BF16X3 ? 2 : 1);
  __nv_bfloat16* part[3] = {nullptr, nullptr, nullptr};
  for (int i = 0; i < nparts; ++i) {
    part[i] = (__nv_bfloat16*)stage;
    stage += align_up((size_t)n * 2, 256);
  }
  if (nparts == 3) split_hi_mid_lo_kernel<<<ew_blocks(n), 256, 0, st>>>((const float*)src, part[0], part[1], part[2], n);
  else if (nparts == 2) split_hi_lo_kernel<<<ew_blocks(n), 256, 0, st>>>((const float*)src, part[0], part[1], n);
  else cast_to_bf16_kernel<<<ew_blocks(n), 256, 0, st>>>((const float*)src, part[0], n);
  DA_LAUNCH_CHECK();
  for (int i = 0; i < nparts; ++i) out->p[i] = part[i];
  out->n = nparts;
  return DA_OK;
}

// Product terms of the split-precision engines (indices into the operand parts 0 = hi, 1 = mid / lo, 2 = lo).
//   BF16X3: x = hi + lo (16 bits):  hi*hi + hi*lo + lo*hi                      (dropped: lo*lo ~ 2^-16)
//   BF16X6: x = hi + mid + lo (24 bits, exact): hh + hm + mh + hl + lh + mm    (dropped: ml, lm, ll ~ 2^-24: fp32 class)
// Small terms first: they are accumulated before the fp32 sum grows.
static void set_terms(int engine, int* num_terms, int* ta, int* tb) {
  for (int i = 0; i < 6; ++i) ta[i] = tb[i] = 0;
  if (engine == DA_ENGINE_UMMA_BF16X6) {
    *num_terms = 6;
    const int a[6] = {1, 0, 2, 0, 1, 0}, b[6] = {1, 2, 0, 1, 0, 0};
    for (int i = 0; i < 6; ++i) { ta[i] = a[i]; tb[i] = b[i]; }
  } else if (engine == DA_ENGINE_UMMA_BF16X3) {
    *num_terms = 3;
    ta[0] = 0; tb[0] = 1;  // hi*lo
    ta[1] = 1; tb[1] = 0;  // lo*hi
    ta[2] = 0; tb[2] = 0;  // hi*hi
  } else {
    *num_terms = 1;
  }
}

// 128x256 tiles unless that leaves most SMs idle on a short-K problem (the instance-head FCs: M = N = 1024,
// K <= 1024): there 128x128 tiles double the CTA count and make split-K (and its finish kernel) unnecessary.
static inline int choose_bn(int ncols, long long m_rows, long long k_iters) {
  if (ncols <= 64) return 64;
  const long long mt = (m_rows + BM - 1) / BM;
  const bool short_k = k_iters < 64;
  if (ncols > 128 && (mt * ((ncols + 255) / 256) >= num_sms() / 2 || !short_k)) return 256;
  // short-K problems that do not fill the machine with 128-wide tiles: 128x64 tiles (twice the CTAs)
  if (short_k && mt * ((ncols + 127) / 128) < num_sms() / 2 && !g_opt.umma_no_bn64) return 64;
  return 128;
}

static int pick_splits_persistent(long long tiles, int k_iters) {
  // persistent kernel: keep #tile-units <= #SMs (one wave) when splitting
  const int sms = num_sms();
  if (tiles >= sms / 2 || k_iters < 16) return 1;
  int s = (int)(sms / tiles);
  if (s > k_iters / 32) s = k_iters / 32;   // a split must amortise the partial write + finish kernel
  if (s > 16) s = 16;
  return s < 1 ? 1 : s;
}

// Weight-gradient kernel of the fp32-class engine: its accumulation chains (pixels) are cut by split-K instead of in-kernel
// chunking (see kChunked above for why chains must be short): at most kX6ChainIters k-steps per chain, partials summed in
// fp32 round-to-nearest by sum_splits_kernel.
constexpr int kX6ChainIters = 32;
constexpr int kMaxSplits = 64;

template <int kBN, int kCluster, bool k2SM = false, bool kChunked = false>
static int launch_nt_t(NtParams& P, long long pixel_tiles, int k_iters, void* ws_part, size_t part_cap, cudaStream_t st) {
  const int n_tiles = (P.Cout + kBN - 1) / kBN;
  const long long super_tiles = (pixel_tiles + kCluster - 1) / kCluster;
  // split-K partials are indexed like y; a strided (parity-class) launch only owns part of y
  int splits = (P.os == 1) ? pick_splits_persistent(super_tiles * kCluster * n_tiles, k_iters) : 1;

  while (splits > 1 && (size_t)splits * P.y_numel * sizeof(float) > part_cap) --splits;
  P.partial = (float*)ws_part;
  const long long total = super_tiles * n_tiles * splits;   // cluster-level tile units
  DA_REQUIRE(total * kCluster <= 0x7fffffffll, DA_ERR_UNSUPPORTED, "umma: too many tiles");
  auto kern = umma_nt_kernel<kBN, kCluster, k2SM, kChunked>;
  static bool attr_set_dev[kMaxDevices] = {};     // function attributes are per device
  bool& attr_set = attr_set_dev[cur_dev()];
  if (!attr_set) {
    DA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<kBN>::SMEM));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(NT_FWD_THREADS);
  cfg.dynamicSmemBytes = TileCfg<kBN>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt.no_pdl ? 1 : 2;
  // persistent grid = the clusters that are co-resident (GPC sizes need not be multiples of the cluster size)
  static int hw_clusters_dev[kMaxDevices] = {};
  int& hw_clusters = hw_clusters_dev[cur_dev()];
  if (hw_clusters == 0) {
    cfg.gridDim = dim3((num_sms_physical() / kCluster) * kCluster);
    int n = 0;
    if (kCluster > 1 && cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) hw_clusters = n;
    else hw_clusters = num_sms_physical() / kCluster;
    (void)cudaGetLastError();
    if (hw_clusters > num_sms_physical() / kCluster) hw_clusters = num_sms_physical() / kCluster;
  }
  int max_clusters = num_sms() / kCluster;                 // honours da_set_sm_limit
  if (max_clusters > hw_clusters) max_clusters = hw_clusters;
  if (max_clusters < 1) max_clusters = 1;
  const int clusters = (int)(total < max_clusters ? total : max_clusters);
  cfg.gridDim = dim3(clusters * kCluster);
  DA_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, P, (int)pixel_tiles, n_tiles, splits));
  DA_LAUNCH_CHECK();
  if (splits > 1) {
    const bool vec4 = P.y_numel % 4 == 0 && P.Cout % 4 == 0 && ((reinterpret_cast<uintptr_t>(P.y) | reinterpret_cast<uintptr_t>(P.partial) |
                       reinterpret_cast<uintptr_t>(P.scale) | reinterpret_cast<uintptr_t>(P.shift)) & 15) == 0;
    if (vec4)
      nt_splitk_finish_vec4_kernel<<<ew_blocks(P.y_numel / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(P.partial), splits, P.y_numel / 4,
                                                                             P.Cout, P.scale, P.shift, P.relu, P.drop_p, P.seed, P.seed_ctr,
                                                                             P.out_scale, P.y, P.y_dtype);
    else
      nt_splitk_finish_kernel<<<ew_blocks(P.y_numel), 256, 0, st>>>(P.partial, splits, P.y_numel, P.Cout, P.scale, P.shift,
                                                                    P.relu, P.drop_p, P.seed, P.seed_ctr, P.out_scale, P.y, P.y_dtype);
    DA_LAUNCH_CHECK();
  }
  return DA_OK;
}
// Cluster shape of the forward / data-gradient GEMM:
//   pair (2 CTAs, ONE cta_group::2 MMA over 256 pixel rows) when the pixel tiles pair up and the tile is 256 wide;
//   else 2 CTAs sharing the weight tile through TMA multicast; else single CTAs.
static inline bool nt_pair_mma(int bn, long long pixel_tiles) {
  // an odd tile count leaves one phantom tile in the last pair (its A rows are TMA zero fill, its rows are not stored):
  // accepted when that is < 7 % of the work
  return bn == 256 && pixel_tiles >= 2 && (pixel_tiles % 2 == 0 || pixel_tiles >= 15) && !g_opt.umma_no_2sm;
}
static inline int nt_cluster(int bn, long long pixel_tiles) {
  (void)bn;
  return pixel_tiles >= 2 ? 2 : 1;
}

static int launch_nt(NtParams& P, int bn, long long pixel_tiles, int k_iters, void* ws_part, size_t part_cap, cudaStream_t st) {
  P.dbg = g_opt.umma_dbg;
  // concurrently running tiles share the operand whose index is NOT the fastest one; stream the bigger operand once
  P.pix_fast = ((long long)P.Cout > pixel_tiles * BM) ? 1 : 0;
  if (P.num_terms == 6)   // fp32-class engine: chunked accumulation (short TMEM chains, fp32 register sums), 128-wide tiles
    return nt_cluster(128, pixel_tiles) == 2 ? launch_nt_t<128, 2, false, true>(P, pixel_tiles, k_iters, ws_part, part_cap, st)
                                             : launch_nt_t<128, 1, false, true>(P, pixel_tiles, k_iters, ws_part, part_cap, st);
  if (bn == 512) return launch_nt_t<512, 2, true>(P, pixel_tiles, k_iters, ws_part, part_cap, st);
  if (nt_pair_mma(bn, pixel_tiles)) return launch_nt_t<256, 2, true>(P, pixel_tiles, k_iters, ws_part, part_cap, st);
  const int cl = nt_cluster(bn, pixel_tiles);   // CTAs sharing the weight tile through TMA multicast
  if (bn == 256)
    return cl == 2 ? launch_nt_t<256, 2>(P, pixel_tiles, k_iters, ws_part, part_cap, st)
                   : launch_nt_t<256, 1>(P, pixel_tiles, k_iters, ws_part, part_cap, st);
  if (bn == 64)   // an MN-major B tile of 64 channels is ONE 64x64 box: nothing to split across a multicast pair
    return (cl == 2 && !P.b_mn_major) ? launch_nt_t<64, 2>(P, pixel_tiles, k_iters, ws_part, part_cap, st)
                                      : launch_nt_t<64, 1>(P, pixel_tiles, k_iters, ws_part, part_cap, st);
  return cl == 2 ? launch_nt_t<128, 2>(P, pixel_tiles, k_iters, ws_part, part_cap, st)
                 : launch_nt_t<128, 1>(P, pixel_tiles, k_iters, ws_part, part_cap, st);
}

int umma_conv_forward(const da_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                      int relu, float drop_p, uint64_t seed, void* y, void* ws, size_t ws_bytes, cudaStream_t st) {
  const Geom g = geom_of(d);
  DA_REQUIRE(g.Cin % 8 == 0, DA_ERR_UNSUPPORTED, "umma conv: Cin=%d must be a multiple of 8", g.Cin);
  DA_REQUIRE(g.KH * g.KW <= MAX_TAPS, DA_ERR_UNSUPPORTED, "umma conv: at most %d filter taps", MAX_TAPS);
  DA_REQUIRE(g.s <= 2, DA_ERR_UNSUPPORTED, "umma conv: stride %d not built (1 or 2)", g.s);
  DA_REQUIRE(ws && ws_bytes >= umma_workspace_bytes(d), DA_ERR_WORKSPACE, "umma conv forward: workspace too small (%zu < %zu)", ws_bytes, umma_workspace_bytes(d));
  uint8_t* stage = (uint8_t*)ws + part_bytes(g);
  Parts xs, wsrc;
  int rc = prep_operand(x, d->x_dtype, (long long)g.N * g.H * g.W * g.Cin, d->engine, stage, &xs, st);
  if (rc) return rc;
  rc = prep_operand(w, d->x_dtype, (long long)g.Cout * g.KH * g.KW * g.Cin, d->engine, stage, &wsrc, st);
  if (rc) return rc;

  NtParams P;
  memset(&P, 0, sizeof(P));
  set_terms(d->engine, &P.num_terms, P.term_a, P.term_b);
  P.kchunks = (g.Cin + BK - 1) / BK;
  P.Cout = g.Cout; P.OHf = g.OH; P.OWf = g.OW; P.os = 1; P.oa = 0; P.ob = 0;
  P.scale = scale; P.shift = shift; P.relu = relu; P.drop_p = drop_p; P.seed = seed; P.seed_ctr = g_seed_counter; P.out_scale = 1.f;
  P.y = y; P.y_dtype = d->y_dtype; P.y_numel = (long long)g.N * g.OH * g.OW * g.Cout;
  const int Ktot = g.KH * g.KW * g.Cin;
  int bn = d->engine == DA_ENGINE_UMMA_BF16X6 ? 128 : choose_bn(g.Cout, (long long)g.N * g.OH * g.OW, (Ktot + BK - 1) / BK);
  long long pixel_tiles;
  if (is_flat(g)) {
    P.flat = 1;
    P.M_flat = (long long)g.N * g.H * g.W;
    P.num_taps = 1;
    P.taps[0] = TapInfo{0, 0, 0, 0};
    for (int t = 0; t < xs.n; ++t) {
      const uint64_t dims[2] = {(uint64_t)g.Cin, (uint64_t)P.M_flat};
      const uint64_t strides[1] = {(uint64_t)g.Cin * 2};
      const uint32_t box[2] = {BK, BM};
      rc = encode_map(&P.a_map[t][0], xs.p[t], 2, dims, strides, box);
      if (rc) return rc;
    }
    pixel_tiles = (P.M_flat + BM - 1) / BM;
  } else {
    P.flat = 0;
    P.TH = g.OH; P.TW = g.OW;
    pick_patch(g.OH, g.OW, BM, &P.BH, &P.BW);
    P.bw_shift = ilog2(P.BW);
    P.tiles_h = (g.OH + P.BH - 1) / P.BH; P.tiles_w = (g.OW + P.BW - 1) / P.BW;
    bool present[4] = {false, false, false, false};
    for (int t = 0; t < xs.n; ++t) {
      rc = make_parity_maps(P.a_map[t], present, xs.p[t], g.N, g.H, g.W, g.Cin, g.s, P.BH, P.BW);
      if (rc) return rc;
    }
    int nt = 0;
    for (int kh = 0; kh < g.KH; ++kh)
      for (int kw = 0; kw < g.KW; ++kw) {
        const int pa = posmod(kh - g.p, g.s), pb = posmod(kw - g.p, g.s);
        if (!present[pa * g.s + pb]) continue;
        P.taps[nt++] = TapInfo{pa * g.s + pb, floordiv(kh - g.p, g.s), floordiv(kw - g.p, g.s), (kh * g.KW + kw) * g.Cin};
      }
    P.num_taps = nt;
    pixel_tiles = (long long)g.N * P.tiles_h * P.tiles_w;
  }
  // long-K problems whose tiles pair up: 512-wide tiles (one TMEM accumulator, two N = 256 pair MMAs per k-substep)
  if (bn == 256 && d->engine == DA_ENGINE_UMMA_BF16 && nt_pair_mma(bn, pixel_tiles) && g.Cout % 512 == 0 &&
      (long long)P.num_taps * P.kchunks >= 24 && !g_opt.umma_no_bn512)
    bn = 512;
  for (int t = 0; t < wsrc.n; ++t) {
    // a 2-CTA cluster (pixel_tiles >= 2, see launch_nt) loads the weight tile as two multicast halves
    const uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)g.Cout};
    const uint64_t strides[1] = {(uint64_t)Ktot * 2};
    const uint32_t box[2] = {BK, (uint32_t)(bn == 512 ? 128 : bn / nt_cluster(bn, pixel_tiles))};
    rc = encode_map(&P.b_map[t], wsrc.p[t], 2, dims, strides, box);
    if (rc) return rc;
  }
  return launch_nt(P, bn, pixel_tiles, P.num_terms * P.num_taps * P.kchunks, ws, part_bytes(g), st);
}

int umma_conv_backward_data(const da_conv_desc* d, const void* dz, const void* w, float out_scale, void* dx, void* ws,
                            size_t ws_bytes, cudaStream_t st) {
  const Geom g = geom_of(d);
  DA_REQUIRE(g.Cin % 8 == 0 && g.Cout % 8 == 0, DA_ERR_UNSUPPORTED, "umma dgrad: channels must be multiples of 8");
  DA_REQUIRE(g.KH * g.KW <= MAX_TAPS && g.s <= 2, DA_ERR_UNSUPPORTED, "umma dgrad: unsupported filter/stride");
  DA_REQUIRE(ws && ws_bytes >= umma_workspace_bytes(d), DA_ERR_WORKSPACE, "umma dgrad: workspace too small");
  uint8_t* stage = (uint8_t*)ws + part_bytes(g);
  const int taps = g.KH * g.KW;
  Parts zs, wsrc;
  int rc = prep_operand(dz, d->x_dtype, (long long)g.N * g.OH * g.OW * g.Cout, d->engine, stage, &zs, st);
  if (rc) return rc;
  // B = the weights themselves, read UNtransposed as an MN-major operand: for tap t the B tile is
  // W[co, t*Cin + ci] viewed as [k = co rows][n = ci contiguous] -> boxes of (64 ci, 64 co)
  rc = prep_operand(w, d->x_dtype, (long long)g.Cout * taps * g.Cin, d->engine, stage, &wsrc, st);
  if (rc) return rc;
  NtParams base;
  memset(&base, 0, sizeof(base));
  set_terms(d->engine, &base.num_terms, base.term_a, base.term_b);
  base.b_mn_major = 1;
  base.kchunks = (g.Cout + BK - 1) / BK;
  base.Cout = g.Cin;  // GEMM N dimension = input channels
  base.OHf = g.H; base.OWf = g.W;
  base.relu = 0; base.drop_p = 0.f; base.out_scale = out_scale;
  base.y = dx; base.y_dtype = d->y_dtype; base.y_numel = (long long)g.N * g.H * g.W * g.Cin;
  int bn = d->engine == DA_ENGINE_UMMA_BF16X6 ? 128 : choose_bn(g.Cin, (long long)g.N * g.H * g.W, ((long long)taps * g.Cout + BK - 1) / BK);
  // data gradients with >= 24 k-steps per launch whose pixel tiles pair up: 512-wide tiles (see TileCfg)
  if (bn == 256 && d->engine == DA_ENGINE_UMMA_BF16 && g.Cin % 512 == 0 && (long long)taps * base.kchunks / (g.s * g.s) >= 24 &&
      !g_opt.umma_no_bn512) {
    long long ptiles;
    // every launch of this call (one per input-parity class for stride 2) must pair up
    bool ok = true;
    if (is_flat(g)) ok = nt_pair_mma(bn, ((long long)g.N * g.H * g.W + BM - 1) / BM);
    else
      for (int a = 0; a < g.s && ok; ++a)
        for (int b = 0; b < g.s && ok; ++b) {
          const int th = (g.H - a + g.s - 1) / g.s, tw = (g.W - b + g.s - 1) / g.s;
          if (th <= 0 || tw <= 0) continue;
          int bh, bw;
          pick_patch(th, tw, BM, &bh, &bw);
          ok = nt_pair_mma(bn, (long long)g.N * ((th + bh - 1) / bh) * ((tw + bw - 1) / bw));
        }
    if (ok) bn = 512;
  }
  for (int t = 0; t < wsrc.n; ++t) {
    const uint64_t dims[2] = {(uint64_t)taps * g.Cin, (uint64_t)g.Cout};
    const uint64_t strides[1] = {(uint64_t)taps * g.Cin * 2};
    const uint32_t box[2] = {64, 64};
    rc = encode_map(&base.b_map[t], wsrc.p[t], 2, dims, strides, box);
    if (rc) return rc;
  }
  if (is_flat(g)) {
    NtParams P = base;
    P.flat = 1; P.M_flat = (long long)g.N * g.H * g.W; P.num_taps = 1; P.taps[0] = TapInfo{0, 0, 0, 0};
    P.os = 1;
    for (int t = 0; t < zs.n; ++t) {
      const uint64_t dims[2] = {(uint64_t)g.Cout, (uint64_t)P.M_flat};
      const uint64_t strides[1] = {(uint64_t)g.Cout * 2};
      const uint32_t box[2] = {BK, BM};
      rc = encode_map(&P.a_map[t][0], zs.p[t], 2, dims, strides, box);
      if (rc) return rc;
    }
    return launch_nt(P, bn, (P.M_flat + BM - 1) / BM, P.num_terms * P.kchunks, ws, part_bytes(g), st);
  }
  // one launch per input-pixel parity class (a,b): h = s*i + a, w = s*j + b
  for (int a = 0; a < g.s; ++a)
    for (int b = 0; b < g.s; ++b) {
      NtParams P = base;
      P.flat = 0;
      P.TH = (g.H - a + g.s - 1) / g.s; P.TW = (g.W - b + g.s - 1) / g.s;
      if (P.TH <= 0 || P.TW <= 0) continue;
      P.os = g.s; P.oa = a; P.ob = b;
      pick_patch(P.TH, P.TW, BM, &P.BH, &P.BW);
      P.bw_shift = ilog2(P.BW);
      P.tiles_h = (P.TH + P.BH - 1) / P.BH; P.tiles_w = (P.TW + P.BW - 1) / P.BW;
      bool present[4];
      for (int t = 0; t < zs.n; ++t) {
        rc = make_parity_maps(P.a_map[t], present, zs.p[t], g.N, g.OH, g.OW, g.Cout, 1, P.BH, P.BW);
        if (rc) return rc;
      }
      int nt = 0;
      for (int kh = 0; kh < g.KH; ++kh)
        for (int kw = 0; kw < g.KW; ++kw) {
          const int th = a + g.p - kh, tw = b + g.p - kw;  // oh = i + th/s when s | th
          if (posmod(th, g.s) != 0 || posmod(tw, g.s) != 0) continue;
          P.taps[nt++] = TapInfo{0, floordiv(th, g.s), floordiv(tw, g.s), (kh * g.KW + kw) * g.Cin};
        }
      P.num_taps = nt;
      rc = launch_nt(P, bn, (long long)g.N * P.tiles_h * P.tiles_w, P.num_terms * nt * P.kchunks, ws, part_bytes(g), st);
      if (rc) return rc;
    }
  return DA_OK;
}

template <int kBN, int kCluster, bool k2SM = false>
static int launch_tn_t(const TnParams& P, int co_tiles, int ci_tiles, int splits, cudaStream_t st) {
  const long long co_super = (co_tiles + kCluster - 1) / kCluster;
  const long long total = co_super * ci_tiles * P.num_taps * splits;   // cluster-level tile units
  DA_REQUIRE(total * kCluster <= 0x7fffffffll, DA_ERR_UNSUPPORTED, "umma wgrad: too many tiles");
  auto kern = umma_tn_kernel<kBN, kCluster, k2SM>;
  static bool attr_set_dev[kMaxDevices] = {};     // function attributes are per device
  bool& attr_set = attr_set_dev[cur_dev()];
  if (!attr_set) {
    DA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<kBN>::SMEM));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(NT_FWD_THREADS);
  cfg.dynamicSmemBytes = TileCfg<kBN>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt.no_pdl ? 1 : 2;
  // persistent grid = the clusters that are co-resident (GPC sizes need not be multiples of the cluster size)
  static int hw_clusters_dev[kMaxDevices] = {};
  int& hw_clusters = hw_clusters_dev[cur_dev()];
  if (hw_clusters == 0) {
    cfg.gridDim = dim3((num_sms_physical() / kCluster) * kCluster);
    int n = 0;
    if (kCluster > 1 && cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) hw_clusters = n;
    else hw_clusters = num_sms_physical() / kCluster;
    (void)cudaGetLastError();
    if (hw_clusters > num_sms_physical() / kCluster) hw_clusters = num_sms_physical() / kCluster;
  }
  int max_clusters = num_sms() / kCluster;                 // honours da_set_sm_limit
  if (max_clusters > hw_clusters) max_clusters = hw_clusters;
  if (max_clusters < 1) max_clusters = 1;
  const int clusters = (int)(total < max_clusters ? total : max_clusters);
  cfg.gridDim = dim3(clusters * kCluster);
  DA_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, P, co_tiles, ci_tiles, splits));
  count_launch();
  return DA_OK;
}

int umma_conv_backward_weight(const da_conv_desc* d, const void* x, const void* dz, float* dw, void* ws,
                              size_t ws_bytes, cudaStream_t st, const da_sgd_fuse* sgd) {
  const Geom g = geom_of(d);
  if (sgd) {
    DA_REQUIRE(sgd->w && sgd->momentum_buf, DA_ERR_INVALID_ARG, "umma wgrad+sgd: null master / momentum");
    DA_REQUIRE(g.Cin % 32 == 0, DA_ERR_UNSUPPORTED, "umma wgrad+sgd: Cin must be a multiple of 32");
    DA_REQUIRE((((uintptr_t)sgd->w | (uintptr_t)sgd->momentum_buf) & 15) == 0 && ((uintptr_t)sgd->w_bf16 & 7) == 0,
               DA_ERR_INVALID_ARG, "umma wgrad+sgd: tensors must be 16-byte aligned");
  } else {
    DA_REQUIRE(dw, DA_ERR_INVALID_ARG, "umma wgrad: null output");
  }
  DA_REQUIRE(g.Cin % 8 == 0 && g.Cout % 8 == 0, DA_ERR_UNSUPPORTED, "umma wgrad: channels must be multiples of 8");
  DA_REQUIRE(g.KH * g.KW <= MAX_TAPS && g.s <= 2, DA_ERR_UNSUPPORTED, "umma wgrad: unsupported filter/stride");
  DA_REQUIRE(ws && ws_bytes >= umma_workspace_bytes(d), DA_ERR_WORKSPACE, "umma wgrad: workspace too small");
  uint8_t* stage = (uint8_t*)ws + part_bytes(g);
  Parts xs, zs;
  int rc = prep_operand(x, d->x_dtype, (long long)g.N * g.H * g.W * g.Cin, d->engine, stage, &xs, st);
  if (rc) return rc;
  rc = prep_operand(dz, d->x_dtype, (long long)g.N * g.OH * g.OW * g.Cout, d->engine, stage, &zs, st);
  if (rc) return rc;

  TnParams P;
  memset(&P, 0, sizeof(P));
  set_terms(d->engine, &P.num_terms, P.term_a, P.term_b);
  P.Cout = g.Cout; P.Cin = g.Cin;
  P.dw_numel = (long long)g.Cout * g.KH * g.KW * g.Cin;
  long long patches;
  if (is_flat(g)) {
    P.flat = 1; P.M_flat = (long long)g.N * g.H * g.W; P.num_taps = 1; P.taps[0] = TapInfo{0, 0, 0, 0};
    for (int t = 0; t < zs.n; ++t) {
      const uint64_t da_[2] = {(uint64_t)g.Cout, (uint64_t)P.M_flat};
      const uint64_t sa[1] = {(uint64_t)g.Cout * 2};
      const uint32_t box[2] = {64, WK};
      rc = encode_map(&P.a_map[t], zs.p[t], 2, da_, sa, box);
      if (rc) return rc;
      const uint64_t db[2] = {(uint64_t)g.Cin, (uint64_t)P.M_flat};
      const uint64_t sb[1] = {(uint64_t)g.Cin * 2};
      rc = encode_map(&P.b_map[t][0], xs.p[t], 2, db, sb, box);
      if (rc) return rc;
    }
    patches = (P.M_flat + WK - 1) / WK;
  } else {
    P.flat = 0;
    pick_patch(g.OH, g.OW, WK, &P.BH, &P.BW);
    P.tiles_h = (g.OH + P.BH - 1) / P.BH; P.tiles_w = (g.OW + P.BW - 1) / P.BW; P.NB = g.N;
    bool present[4] = {false, false, false, false}, pz[4];
    for (int t = 0; t < zs.n; ++t) {
      CUtensorMap tmp[4];
      rc = make_parity_maps(tmp, pz, zs.p[t], g.N, g.OH, g.OW, g.Cout, 1, P.BH, P.BW);
      if (rc) return rc;
      P.a_map[t] = tmp[0];
      rc = make_parity_maps(P.b_map[t], present, xs.p[t], g.N, g.H, g.W, g.Cin, g.s, P.BH, P.BW);
      if (rc) return rc;
    }
    int nt = 0;
    for (int kh = 0; kh < g.KH; ++kh)
      for (int kw = 0; kw < g.KW; ++kw) {
        const int pa = posmod(kh - g.p, g.s), pb = posmod(kw - g.p, g.s);
        // taps whose parity plane is empty contribute nothing; they keep a slot (map 0 with an
        // always-out-of-bounds shift) so that dW stays densely indexed by (kh,kw)
        if (!present[pa * g.s + pb]) { P.taps[nt++] = TapInfo{0, 1 << 20, 1 << 20, 0}; continue; }
        P.taps[nt++] = TapInfo{pa * g.s + pb, floordiv(kh - g.p, g.s), floordiv(kw - g.p, g.s), 0};
      }
    P.num_taps = nt;
    patches = (long long)g.N * P.tiles_h * P.tiles_w;
  }
  // weight-gradient tiles are [128 Cout x bn Cin] per tap; K runs over the pixels
  // weight-gradient tiles are [128 Cout x bn Cin] per tap; K runs over the pixels.  Short pixel loops: more, narrower
  // tiles instead of split-K.
  const long long co_t = (g.Cout + BM - 1) / BM;
  const bool short_k = patches * P.num_terms < 64;
  int bn = 128;
  if (g.Cin <= 64) bn = 64;
  else if (g.Cin > 128 && (co_t * ((g.Cin + 255) / 256) * P.num_taps >= num_sms() / 2 || !short_k)) bn = 256;
  else if (short_k && co_t * ((g.Cin + 127) / 128) * P.num_taps < num_sms() / 2 && !g_opt.umma_no_bn64) bn = 64;
  // long pixel loops, Cout tiles that pair up, no fused optimizer (its epilogue needs the overlap): 512-wide pair tiles
  // (one TMEM accumulator, two N = 256 pair MMAs per k-substep share the dZ operand: operand bytes per FLOP -25 %)
  if (bn == 256 && !sgd && d->engine == DA_ENGINE_UMMA_BF16 && g.Cin % 512 == 0 && co_t >= 2 && co_t % 2 == 0 &&
      patches * P.num_terms >= 32 && !g_opt.umma_no_2sm && !g_opt.umma_no_bn512)
    bn = 512;
  const long long tiles = (long long)((g.Cout + BM - 1) / BM) * ((g.Cin + bn - 1) / bn) * P.num_taps;
  long long k_iters = patches * P.num_terms;
  int splits = pick_splits_persistent(tiles, (int)(k_iters > 1000000 ? 1000000 : k_iters));
  if (P.num_terms == 6) {   // fp32-class engine: bounded accumulation chains (see kX6ChainIters)
    const long long need = (k_iters + kX6ChainIters - 1) / kX6ChainIters;
    if (need > splits) splits = (int)(need > kMaxSplits ? kMaxSplits : need);
  }
  while (splits > 1 && (size_t)splits * P.dw_numel * sizeof(float) > part_bytes(g)) --splits;
  if (sgd) {     // every gradient element must be complete inside one tile: no split-K
    splits = 1;
    P.sgd_w = sgd->w; P.sgd_buf = sgd->momentum_buf; P.sgd_shadow = (__nv_bfloat16*)sgd->w_bf16;
    P.sgd_lr = sgd->lr; P.sgd_mu = sgd->momentum; P.sgd_wd = sgd->weight_decay; P.sgd_first = sgd->first_step;
  }
  P.dw = splits > 1 ? (float*)ws : dw;
  const int co_tiles = (g.Cout + BM - 1) / BM, ci_tiles = (g.Cin + bn - 1) / bn;
  // Cout tiles that share an X tile form a cluster (TMA multicast of the shared operand)
  int rc2;
  if (bn == 512) {
    rc2 = launch_tn_t<512, 2, true>(P, co_tiles, ci_tiles, splits, st);
  } else if (bn == 256) {
    // clusters of 4 only fit 33 times on the 148 SMs (GPC sizes), pairs fit 74 times
    if (co_tiles >= 2 && co_tiles % 2 == 0 && !g_opt.umma_no_2sm) rc2 = launch_tn_t<256, 2, true>(P, co_tiles, ci_tiles, splits, st);   // CTA pair
    else if (co_tiles >= 2) rc2 = launch_tn_t<256, 2>(P, co_tiles, ci_tiles, splits, st);
    else rc2 = launch_tn_t<256, 1>(P, co_tiles, ci_tiles, splits, st);
  } else if (bn == 64) {
    rc2 = launch_tn_t<64, 1>(P, co_tiles, ci_tiles, splits, st);
  } else {
    if (co_tiles >= 2) rc2 = launch_tn_t<128, 2>(P, co_tiles, ci_tiles, splits, st);
    else rc2 = launch_tn_t<128, 1>(P, co_tiles, ci_tiles, splits, st);
  }
  if (rc2) return rc2;
  if (splits > 1) {
    if (P.dw_numel % 4 == 0 && ((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0)
      sum_splits_vec4_kernel<<<ew_blocks(P.dw_numel / 4), 256, 0, st>>>((const float4*)ws, splits, P.dw_numel / 4, (float4*)dw);
    else
      sum_splits_kernel<<<ew_blocks(P.dw_numel), 256, 0, st>>>((const float*)ws, splits, P.dw_numel, dw);
    DA_LAUNCH_CHECK();
  }
  return DA_OK;
}

}  // namespace da
