// One persistent kernel for a whole CHAIN of small dense contractions -- the instance-level domain classifier
// (mmdet/models/roi_heads/instance_da.py:42-148: NonLocalBlock over the k RoIs, FC 1024-512-512-2, sigmoid) fused with its
// loss (CE on the sigmoid outputs, detectors/DAFaster_rcnn_Orig.py:177-188), forward and backward -- sm_100a only.
//
// Why a chain and not one launch per layer: at k = 1024 RoIs every layer of the head is a 1-3 GFLOP GEMM that a B200
// finishes in ~2 us of tensor time, but as a kernel of its own it costs 9-13 us (launch, barrier init, TMEM allocation,
// tensor-map fetch, pipeline fill, tail): the head was 27 GEMM launches + 20 small kernels = 0.5 ms of a 1.9 ms step.
// Why not one CTA per row block that keeps the activations on chip: each CTA would have to pull EVERY weight matrix through
// its own L2->SM port (1.5 MB for FC1+FC2 = 17 us at ~90 GB/s per SM), slower than N-parallel tiles on all SMs.
//
// So: a PROGRAM of ops, grouped into barrier-delimited groups.  Inside a group the tiles of all GEMM ops are dealt
// round-robin to the persistent CTAs (warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..9 = epilogue, the
// mbarrier ring and both TMEM accumulators stay live across ops and groups), then every thread runs the group's
// elementwise ops (query-axis softmax, CE, column sums), then a grid-wide barrier (one atomic + acquire spin; epilogue
// writes are fenced generic->async proxy so the next group's TMA loads see them).  Operands are bf16 in global memory
// (L2 resident: the whole head's working set is < 30 MB), any of A / B may be K-major ([rows, K]) or MN-major ([K, rows]),
// with a row stride, so column slices of one buffer (theta | phi | g) and transposed uses (weights in the data gradient,
// activations in the weight gradient) need no copies.  Thin layers (FC 512 -> 2) ride the same path: TMA zero-fills the
// out-of-bounds part of their boxes.
#include "da_common.cuh"
#include "da_ptx.cuh"
#include <cuda.h>
#include <string.h>

namespace da {

// Tile width per op: 64 (more tiles: the small GEMMs are latency sized) or 128 (groups whose tiles fill the machine anyway:
// with every SM pulling operands the L2 fabric, not the SM port, is the limit, and a 128-wide tile needs 1/3 fewer operand
// bytes per FLOP).  The 192 KB operand ring is cut per GROUP: 8 stages of 16 + 8 KB, or 6 stages of 16 + 16 KB when the group
// has a 128-wide op (everything is drained at a group boundary; mbarrier phases are tracked per stage, not per ring turn).
constexpr int CH_BM = 128, CH_BN_MAX = 128, CH_BK = 64;
constexpr int CH_A_BYTES = CH_BM * CH_BK * 2;          // 16 KB
constexpr int CH_B_BYTES = CH_BN_MAX * CH_BK * 2;      // 16 KB (wide stage)
constexpr int CH_STAGES = 8;                           // barriers; a wide group uses 6 of them
constexpr int CH_RING_BYTES = 8 * (CH_A_BYTES + CH_B_BYTES / 2);   // 192 KB = 6 * (16 + 16) KB
constexpr int CH_THREADS = 320;                    // TMA warp, MMA warp, 8 epilogue warps
constexpr int CH_MAX_OPS = 28;
constexpr int CH_MAX_GROUPS = 16;
constexpr size_t CH_SMEM_FIXED = 1024 + (size_t)CH_RING_BYTES + 512;   // + the program copy (see CH_SMEM)

enum ChainKind {
  CH_GEMM = 0,
  CH_SOFTMAX_COL_FWD = 1,   // P[q,k] = exp(S[q,k]) / sum_q' exp(S[q',k])          (nn.Softmax(dim=1) on [b,q,k], Q11)
  CH_SOFTMAX_COL_BWD = 2,   // dS = P * (dP - sum_q P*dP)
  CH_CE_FWD = 3,            // pred = sigmoid(z); loss = mean_r CE(pred_r, label_r)    (single CTA, deterministic)
  CH_CE_BWD = 4,            // dz (padded to 8 bf16 columns) from dloss, dpred
  CH_COLSUM = 5             // dst[c] = sum_r src[r,c]                                (bias gradients)
};

struct __align__(128) ChainOp {
  CUtensorMap a_map, b_map;
  int kind, group;
  // ---- GEMM: out[M,N] = epilogue(A[M,K] * B[N,K]^T)
  int a_mn, b_mn;            // 0 = K-major ([rows, K] row-major), 1 = MN-major ([K, rows] row-major)
  int M, N, K;
  int bn;                    // tile width: 64 or 128
  int tiles_m, tiles_n;
  float alpha;               // v = alpha * (acc + res) + bias
  const float* bias;         // [N] or null
  const __nv_bfloat16* res;  // [M, ld_res] or null
  int ld_res;
  const __nv_bfloat16* gate; // [M, ld_gate] or null: v *= (gate > 0) ? gate_scale : 0   (ReLU + dropout derivative from the
  int ld_gate;               //                       stored post-activation: a dropped or clamped unit stored exactly 0)
  float gate_scale;
  int relu;
  float drop_p;              // dropout on the output: keep(seed, m*N + n) ? v / (1-p) : 0
  unsigned long long seed;
  void* out;
  int out_f32, ld_out;
  // ---- elementwise ops (meaning per kind, see the kernels)
  const void* p0;
  const void* p1;
  const void* p2;
  const void* p3;
  void* q0;
  void* q1;
  int i0, i1, i2, i3;
  float f0;
};

struct ChainParams {
  int nops, ngroups;
  int group_begin[CH_MAX_GROUPS + 1];
  unsigned int* barrier;                    // zeroed by the host before the launch
  const unsigned long long* seed_ctr;
  unsigned long long* trace;                // debug (da_set_option("chain_trace", ptr)): globaltimer stamps of CTA 0
  ChainOp ops[CH_MAX_OPS];
};

constexpr size_t CH_SMEM = CH_SMEM_FIXED + sizeof(ChainOp) * CH_MAX_OPS;

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// All CTAs of the (co-resident: grid <= #SMs, one CTA per SM) grid have finished everything before this point.
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int target) {
  fence_proxy_async_all();        // this thread's global writes -> visible to later async-proxy (TMA) reads
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(bar) < target) {
      if (clock64() - t0 > 4000000000ll) {
        printf("da_b200: chain grid barrier timed out (block %d, target %u, have %u)\n", blockIdx.x, target, ld_acquire_u32(bar));
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
}

__device__ __forceinline__ void ch_tile_of(const ChainOp* ops, int ob, int oe, int t, int& op, int& m0, int& n0) {
  // tiles of the group's GEMM ops, op after op; inside an op n is the fastest index (neighbouring CTAs share the A tile)
  op = -1;
  for (int o = ob; o < oe; ++o) {
    if (ops[o].kind != CH_GEMM) continue;
    const int n = ops[o].tiles_m * ops[o].tiles_n;
    if (t < n) { op = o; m0 = (t / ops[o].tiles_n) * CH_BM; n0 = (t % ops[o].tiles_n) * ops[o].bn; return; }
    t -= n;
  }
}

// ---- fused epilogue of one [32 rows x 32 columns] chunk: lane = row, 32 consecutive columns per lane ----------------
__device__ __forceinline__ void ch_epilogue_chunk(const ChainOp& o, const uint32_t* acc, int m, int nb, unsigned long long seed) {
  if (m >= o.M || nb >= o.N) return;
  const int ncols = min(32, o.N - nb);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  const bool full = (ncols == 32);
  if (o.res) {
    const __nv_bfloat16* r = o.res + (size_t)m * o.ld_res + nb;
    if (full && ((o.ld_res & 7) == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = *reinterpret_cast<const uint4*>(r + 8 * j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[8 * j + 2 * e] += f.x; v[8 * j + 2 * e + 1] += f.y; }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __bfloat162float(r[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= o.alpha;
  if (o.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(o.bias + nb + j);
  }
  if (o.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (o.drop_p > 0.f) {
    const uint32_t thr = drop_threshold(o.drop_p);
    const float ks = 1.f / (1.f - o.drop_p);
    const uint64_t base = (uint64_t)m * (uint64_t)o.N + (uint64_t)nb;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (drop_hash(seed, base + j) >= thr) ? v[j] * ks : 0.f;
  }
  if (o.gate) {
    const __nv_bfloat16* g = o.gate + (size_t)m * o.ld_gate + nb;
    if (full && ((o.ld_gate & 7) == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = *reinterpret_cast<const uint4*>(g + 8 * j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          v[8 * j + 2 * e] = f.x > 0.f ? v[8 * j + 2 * e] * o.gate_scale : 0.f;
          v[8 * j + 2 * e + 1] = f.y > 0.f ? v[8 * j + 2 * e + 1] * o.gate_scale : 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) v[j] = __bfloat162float(g[j]) > 0.f ? v[j] * o.gate_scale : 0.f;
    }
  }
  if (o.out_f32) {
    float* dst = reinterpret_cast<float*>(o.out) + (size_t)m * o.ld_out + nb;
    if (full && ((o.ld_out & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) dst[j] = v[j];
    }
  } else {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(o.out) + (size_t)m * o.ld_out + nb;
    if (full && ((o.ld_out & 7) == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), b = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
        __nv_bfloat162 c = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), d = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b),
                                                             *reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// ---- elementwise ops: executed by ALL threads of every CTA after the group's GEMM tiles ------------------------------
// Column-wise (query-axis) softmax of S [R rows (q), R columns (k)]: a CTA owns 8 columns at a time, its threads are
// (row group, column) pairs.  One read pass: a thread keeps its (up to CH_SM_ROWS) values of the column in registers, the 40
// row groups of the CTA are combined through shared memory in a fixed order (deterministic), then the registers are
// normalised and written.  Columns taller than 40 * CH_SM_ROWS rows take the three-pass form.
constexpr int CH_SM_ROWS = 26;      // 40 row groups x 26 = 1040 rows: covers the 1024 RoIs of a pair
__device__ void ch_softmax_col_fwd(const ChainOp& o, float* red) {
  const float* S = reinterpret_cast<const float*>(o.p0);
  __nv_bfloat16* Pm = reinterpret_cast<__nv_bfloat16*>(o.q0);
  const int R = o.i0, ldS = o.i1, ldP = o.i2;
  const int col = threadIdx.x & 7, rg = threadIdx.x >> 3, nrg = CH_THREADS / 8;   // 40 row groups
  const bool in_regs = R <= nrg * CH_SM_ROWS;
  for (int c0 = blockIdx.x * 8; c0 < R; c0 += gridDim.x * 8) {
    const int c = c0 + col;
    const bool ok = c < R;
    float v[CH_SM_ROWS];
    float mx = -INFINITY;
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < CH_SM_ROWS; ++i) {
        const int r = rg + i * nrg;
        v[i] = (ok && r < R) ? S[(size_t)r * ldS + c] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
    } else if (ok) {
      for (int r = rg; r < R; r += nrg) mx = fmaxf(mx, S[(size_t)r * ldS + c]);
    }
    red[rg * 8 + col] = mx;
    __syncthreads();
    for (int g = 0; g < nrg; ++g) mx = fmaxf(mx, red[g * 8 + col]);
    __syncthreads();
    float sum = 0.f;
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < CH_SM_ROWS; ++i) { v[i] = ok ? expf(v[i] - mx) : 0.f; sum += v[i]; }     // exp(-inf) = 0 for the padding rows
    } else if (ok) {
      for (int r = rg; r < R; r += nrg) sum += expf(S[(size_t)r * ldS + c] - mx);
    }
    red[rg * 8 + col] = sum;
    __syncthreads();
    sum = 0.f;
    for (int g = 0; g < nrg; ++g) sum += red[g * 8 + col];     // fixed order: deterministic
    __syncthreads();
    const float inv = 1.f / sum;
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < CH_SM_ROWS; ++i) {
        const int r = rg + i * nrg;
        if (ok && r < R) Pm[(size_t)r * ldP + c] = __float2bfloat16_rn(v[i] * inv);
      }
    } else if (ok) {
      for (int r = rg; r < R; r += nrg) Pm[(size_t)r * ldP + c] = __float2bfloat16_rn(expf(S[(size_t)r * ldS + c] - mx) * inv);
    }
  }
}

__device__ void ch_softmax_col_bwd(const ChainOp& o, float* red) {
  const __nv_bfloat16* Pm = reinterpret_cast<const __nv_bfloat16*>(o.p0);
  const float* dP = reinterpret_cast<const float*>(o.p1);
  __nv_bfloat16* dS = reinterpret_cast<__nv_bfloat16*>(o.q0);
  const int R = o.i0, ldP = o.i1, ldD = o.i2, ldO = o.i3;
  const int col = threadIdx.x & 7, rg = threadIdx.x >> 3, nrg = CH_THREADS / 8;
  const bool in_regs = R <= nrg * CH_SM_ROWS;
  for (int c0 = blockIdx.x * 8; c0 < R; c0 += gridDim.x * 8) {
    const int c = c0 + col;
    const bool ok = c < R;
    float p[CH_SM_ROWS], d[CH_SM_ROWS];
    float dot = 0.f;
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < CH_SM_ROWS; ++i) {
        const int r = rg + i * nrg;
        const bool in = ok && r < R;
        p[i] = in ? __bfloat162float(Pm[(size_t)r * ldP + c]) : 0.f;
        d[i] = in ? dP[(size_t)r * ldD + c] : 0.f;
        dot += p[i] * d[i];
      }
    } else if (ok) {
      for (int r = rg; r < R; r += nrg) dot += __bfloat162float(Pm[(size_t)r * ldP + c]) * dP[(size_t)r * ldD + c];
    }
    red[rg * 8 + col] = dot;
    __syncthreads();
    dot = 0.f;
    for (int g = 0; g < nrg; ++g) dot += red[g * 8 + col];
    __syncthreads();
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < CH_SM_ROWS; ++i) {
        const int r = rg + i * nrg;
        if (ok && r < R) dS[(size_t)r * ldO + c] = __float2bfloat16_rn(p[i] * (d[i] - dot));
      }
    } else if (ok) {
      for (int r = rg; r < R; r += nrg) {
        const float pv = __bfloat162float(Pm[(size_t)r * ldP + c]);
        dS[(size_t)r * ldO + c] = __float2bfloat16_rn(pv * (dP[(size_t)r * ldD + c] - dot));
      }
    }
  }
}

// pred = sigmoid(z); loss = mean_r (logsumexp(pred_r) - pred_r[label_r])  (nn.CrossEntropyLoss on the sigmoid outputs, Q4).
__device__ void ch_ce_fwd(const ChainOp& o, float* red) {
  if (blockIdx.x != 0) return;
  const float* z = reinterpret_cast<const float*>(o.p0);
  const int32_t* labels = reinterpret_cast<const int32_t*>(o.p1);
  float* pred = reinterpret_cast<float*>(o.q0);
  float* loss = reinterpret_cast<float*>(o.q1);
  const int R = o.i0;
  float acc = 0.f;
  for (int r = threadIdx.x; r < R; r += CH_THREADS) {
    const float u0 = sigmoidf_(z[2 * r]), u1 = sigmoidf_(z[2 * r + 1]);
    pred[2 * r] = u0;
    pred[2 * r + 1] = u1;
    const int l = labels[r];
    if (l == 0 || l == 1) {
      const float mx = fmaxf(u0, u1);
      acc += mx + logf(expf(u0 - mx) + expf(u1 - mx)) - (l ? u1 : u0);
    }
  }
  acc = block_sum<false>(acc, red);
  if (threadIdx.x == 0) loss[0] = acc / (float)R;
}

// dz[r, 0..1] (bf16, row padded to 8 columns so that it is a legal TMA operand) from dloss (device scalar * f0) and dpred.
__device__ void ch_ce_bwd(const ChainOp& o) {
  const float* z = reinterpret_cast<const float*>(o.p0);
  const int32_t* labels = reinterpret_cast<const int32_t*>(o.p1);
  const float* gl = reinterpret_cast<const float*>(o.p2);
  const float* gp = reinterpret_cast<const float*>(o.p3);
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(o.q0);
  float* dz32 = reinterpret_cast<float*>(o.q1);
  const int R = o.i0;
  const float g = (gl ? gl[0] : 1.f) * o.f0 / (float)R;
  for (int r = blockIdx.x * CH_THREADS + threadIdx.x; r < R; r += gridDim.x * CH_THREADS) {
    const float u0 = sigmoidf_(z[2 * r]), u1 = sigmoidf_(z[2 * r + 1]);
    const int l = labels[r];
    float d0 = 0.f, d1 = 0.f;
    if (l == 0 || l == 1) {
      const float mx = fmaxf(u0, u1);
      const float e0 = expf(u0 - mx), e1 = expf(u1 - mx);
      const float inv = 1.f / (e0 + e1);
      d0 = g * (e0 * inv - (l == 0 ? 1.f : 0.f));
      d1 = g * (e1 * inv - (l == 1 ? 1.f : 0.f));
    }
    if (gp) { d0 += gp[2 * r]; d1 += gp[2 * r + 1]; }
    d0 *= u0 * (1.f - u0);
    d1 *= u1 * (1.f - u1);
    __nv_bfloat162 a = __floats2bfloat162_rn(d0, d1);
    *reinterpret_cast<uint4*>(dz + 8 * (size_t)r) = make_uint4(*reinterpret_cast<uint32_t*>(&a), 0u, 0u, 0u);
    if (dz32) { dz32[2 * r] = d0; dz32[2 * r + 1] = d1; }
  }
}

// dst[c] = sum_r src[r, c]: a CTA owns 8 columns at a time; fixed summation order.
__device__ void ch_colsum(const ChainOp& o, float* red) {
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(o.p0);
  float* dst = reinterpret_cast<float*>(o.q0);
  const int R = o.i0, ld = o.i1, ncols = o.i2;
  const int col = threadIdx.x & 7, rg = threadIdx.x >> 3, nrg = CH_THREADS / 8;
  for (int c0 = blockIdx.x * 8; c0 < ncols; c0 += gridDim.x * 8) {
    const int c = c0 + col;
    float s = 0.f;
    if (c < ncols) for (int r = rg; r < R; r += nrg) s += __bfloat162float(src[(size_t)r * ld + c]);
    red[rg * 8 + col] = s;
    __syncthreads();
    if (rg == 0 && c < ncols) {
      float t = 0.f;
      for (int g = 0; g < nrg; ++g) t += red[g * 8 + col];
      dst[c] = t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(CH_THREADS, 1)
chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + CH_RING_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * CH_STAGES, tfull0 = bars + 16 * CH_STAGES, tempty0 = tfull0 + 16, tslot = tempty0 + 16;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tslot - base));
  float* red = reinterpret_cast<float*>(gen_base);      // elementwise scratch: the operand ring is idle while they run
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The program lives in the kernel-parameter space (14 KB: beyond what the constant cache keeps warm, and every group touches
  // op records no SM has read yet: ~1 us of cold misses per role and group in the first build).  Its scalar fields are read from
  // a shared-memory copy; only the TMA engine reads the tensor maps in place (prefetched into its descriptor cache here).
  ChainOp* ops = reinterpret_cast<ChainOp*>(gen_base + (bars - base) + 512);
  {
    const uint4* src = reinterpret_cast<const uint4*>(&P.ops[0]);
    uint4* dst = reinterpret_cast<uint4*>(ops);
    const int n16 = (int)(sizeof(ChainOp) * P.nops / 16);
    for (int i = threadIdx.x; i < n16; i += CH_THREADS) dst[i] = src[i];
    for (int o = threadIdx.x; o < P.nops; o += CH_THREADS)
      if (P.ops[o].kind == CH_GEMM) { prefetch_tensormap(&P.ops[o].a_map); prefetch_tensormap(&P.ops[o].b_map); }
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < CH_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tslot, 2 * CH_BN_MAX);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tslot_ptr;
  pdl_wait();      // NO early griddepcontrol.launch_dependents: a dependent grid's CTAs must not take SMs this grid barrier needs
  const unsigned long long seed_add = P.seed_ctr ? *P.seed_ctr : 0ull;

  int tcount = 0;               // accumulator hand-overs (MMA thread / epilogue warps)
  uint32_t ring_bits = 0;       // producer: uses of empty[s] mod 2; MMA thread: uses of full[s] mod 2
  int tr = 0;
#define CH_STAMP()                                                                                        \
  do {                                                                                                    \
    if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) {                                                 \
      unsigned long long t_;                                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)::"memory");                                    \
      P.trace[tr++] = t_;                                                                                 \
    }                                                                                                     \
  } while (0)
  CH_STAMP();
  for (int g = 0; g < P.ngroups; ++g) {
    const int ob = P.group_begin[g], oe = P.group_begin[g + 1];
    int total = 0;
    bool wide_g = false;
    for (int o = ob; o < oe; ++o)
      if (ops[o].kind == CH_GEMM) { total += ops[o].tiles_m * ops[o].tiles_n; wide_g |= (ops[o].bn == 128); }
    const int nstages = wide_g ? 6 : 8;
    const uint32_t stage_bytes = CH_A_BYTES + (wide_g ? CH_B_BYTES : CH_B_BYTES / 2);
    int rs = 0;                   // ring stage (producer / MMA thread); the ring is empty at a group boundary
    if (warp == 0) {
      if (lane == 0) {
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
          int op, m0, n0;
          ch_tile_of(ops, ob, oe, t, op, m0, n0);
          const ChainOp& o = ops[op];
          const CUtensorMap* amap = &P.ops[op].a_map;     // the TMA engine reads descriptors from the parameter space
          const CUtensorMap* bmap = &P.ops[op].b_map;
          const int kchunks = (o.K + CH_BK - 1) / CH_BK;
          for (int kc = 0; kc < kchunks; ++kc) {
            const int s = rs;
            rs = (rs + 1 == nstages) ? 0 : rs + 1;
            mbar_wait(empty0 + 8 * s, ((ring_bits >> s) & 1u) ^ 1u);
            ring_bits ^= 1u << s;
            const uint32_t fb = full0 + 8 * s;
            mbar_expect_tx(fb, CH_A_BYTES + o.bn * CH_BK * 2);
            const uint32_t ad = base + s * stage_bytes, bd = ad + CH_A_BYTES;
            if (o.a_mn) {
              tma_load_2d(ad, amap, fb, m0, kc * CH_BK);
              tma_load_2d(ad + CH_A_BYTES / 2, amap, fb, m0 + 64, kc * CH_BK);
            } else {
              tma_load_2d(ad, amap, fb, kc * CH_BK, m0);
            }
            if (o.b_mn) {
              tma_load_2d(bd, bmap, fb, n0, kc * CH_BK);
              if (o.bn == 128) tma_load_2d(bd + CH_B_BYTES / 2, bmap, fb, n0 + 64, kc * CH_BK);
            } else {
              tma_load_2d(bd, bmap, fb, kc * CH_BK, n0);      // box rows = o.bn (encoded per op)
            }
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++tcount) {
          int op, m0, n0;
          ch_tile_of(ops, ob, oe, t, op, m0, n0);
          const ChainOp& o = ops[op];
          const int kchunks = (o.K + CH_BK - 1) / CH_BK;
          const uint32_t idesc = make_idesc(CH_BM, o.bn, o.a_mn, o.b_mn);
          const int buf = tcount & 1;
          mbar_wait(tempty0 + 8 * buf, (((uint32_t)(tcount >> 1)) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * CH_BN_MAX;
          for (int kc = 0; kc < kchunks; ++kc) {
            const int s = rs;
            rs = (rs + 1 == nstages) ? 0 : rs + 1;
            mbar_wait(full0 + 8 * s, (ring_bits >> s) & 1u);
            ring_bits ^= 1u << s;
            const uint32_t a_st = base + s * stage_bytes, b_st = a_st + CH_A_BYTES;
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < CH_BK / 16; ++kk) {
              const uint64_t adsc = o.a_mn ? desc_mnmajor_sw128(a_st + kk * 2048, CH_A_BYTES / 2) : desc_kmajor_sw128(a_st + kk * 32);
              const uint64_t bdsc = o.b_mn ? desc_mnmajor_sw128(b_st + kk * 2048, CH_B_BYTES / 2) : desc_kmajor_sw128(b_st + kk * 32);
              umma_bf16(d_tmem, adsc, bdsc, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(empty0 + 8 * s);
          }
          if (kchunks > 0) umma_commit(tfull0 + 8 * buf);
          else mbar_arrive(tfull0 + 8 * buf);
        }
      }
    } else {
      const int q = warp & 3, chunk = (warp - 2) >> 2;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++tcount) {
        int op, m0, n0;
        ch_tile_of(ops, ob, oe, t, op, m0, n0);
        const ChainOp& o = ops[op];
        const int buf = tcount & 1;
        mbar_wait(tfull0 + 8 * buf, ((uint32_t)(tcount >> 1)) & 1u);
        tc_fence_after();
        uint32_t v[32];
        const int nch = o.bn == 128 ? 2 : 1;               // 32-column chunks of this warp: chunk, chunk + 2
#pragma unroll 1
        for (int ci = 0; ci < nch; ++ci) {
          const int cc = chunk + 2 * ci;
          if (o.K > 0) {
            DA_TMEM_LD32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * CH_BN_MAX + cc * 32, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          if (ci == nch - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);      // the accumulator is in registers: the next tile may start
          }
          ch_epilogue_chunk(o, v, m0 + q * 32 + lane, n0 + cc * 32, o.seed + seed_add);
        }
      }
    }
    if (P.trace && blockIdx.x == 0 && lane == 0 && (warp < 3 || warp == 9)) {     // per-role "done" stamps of the group
      unsigned long long t_;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)::"memory");
      P.trace[128 + 4 * g + (warp == 9 ? 3 : warp)] = t_;
    }
    // the operand ring doubles as scratch of the elementwise ops: every MMA that read it has completed (the epilogue
    // warps waited for the last accumulator) once all warps are here
    __syncthreads();
    CH_STAMP();        // GEMM tiles of the group done (this CTA)
    for (int o = ob; o < oe; ++o) {
      const ChainOp& e = ops[o];
      switch (e.kind) {
        case CH_SOFTMAX_COL_FWD: ch_softmax_col_fwd(e, red); break;
        case CH_SOFTMAX_COL_BWD: ch_softmax_col_bwd(e, red); break;
        case CH_CE_FWD: ch_ce_fwd(e, red); break;
        case CH_CE_BWD: ch_ce_bwd(e); break;
        case CH_COLSUM: ch_colsum(e, red); break;
        default: break;
      }
    }
    CH_STAMP();        // elementwise ops done
    if (g + 1 < P.ngroups) grid_sync(P.barrier, (unsigned int)(g + 1) * gridDim.x);
    // after the LAST grid barrier nothing in this grid waits for another CTA any more: the dependent grid may start its
    // prologue now (its CTAs only get an SM when one of ours exits)
    if (g + 2 == P.ngroups) pdl_launch_dependents();
    CH_STAMP();        // barrier passed
  }
#undef CH_STAMP
  tc_fence_before();
  __syncthreads();
  if (P.ngroups < 2) pdl_launch_dependents();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * CH_BN_MAX);
  }
}

// ---------------------------------------------------------------------------------------
// host: program builder
// ---------------------------------------------------------------------------------------
struct Builder {
  ChainParams P;
  int rc;
  struct Operands { const void* A; const void* B; int lda, ldb; } raw[CH_MAX_OPS];   // tensor maps are encoded in finish()
  Builder() : rc(DA_OK) {
    memset(&P, 0, sizeof(P));
    memset(raw, 0, sizeof(raw));
  }
  ChainOp* add(int kind, int group) {
    if (P.nops >= CH_MAX_OPS || group >= CH_MAX_GROUPS) { set_error("chain: program too long"); rc = DA_ERR_UNSUPPORTED; return nullptr; }
    ChainOp* o = &P.ops[P.nops++];
    o->kind = kind;
    o->group = group;
    o->alpha = 1.f;
    o->gate_scale = 1.f;
    if (group + 1 > P.ngroups) P.ngroups = group + 1;
    return o;
  }
  // operand: base pointer (bf16), logical [rows, K] when mn == 0 (row stride ld), [K, rows] when mn == 1
  static int encode(CUtensorMap* m, const void* base, int mn, int rows, int K, int ld, int box_rows) {
    const uint64_t dims[2] = {(uint64_t)(mn ? rows : K), (uint64_t)(mn ? K : rows)};
    const uint64_t strides[1] = {(uint64_t)ld * 2};
    const uint32_t box[2] = {64, (uint32_t)(mn ? 64 : box_rows)};
    return encode_map(m, base, 2, dims, strides, box);
  }
  ChainOp* gemm(int group, const void* A, int a_mn, int lda, const void* B, int b_mn, int ldb, int M, int N, int K, void* out,
                int out_f32, int ld_out) {
    ChainOp* o = add(CH_GEMM, group);
    if (!o) return nullptr;
    o->a_mn = a_mn; o->b_mn = b_mn; o->M = M; o->N = N; o->K = K;
    o->bn = 64;
    o->out = out; o->out_f32 = out_f32; o->ld_out = ld_out;
    raw[P.nops - 1] = Operands{A, B, lda, ldb};
    return o;
  }
  int finish() {
    if (rc) return rc;
    // tile width per group: 128 when the group's wide-eligible GEMMs still give (nearly) every SM a tile, else 64
    for (int g = 0; g < P.ngroups; ++g) {
      int wide_tiles = 0, narrow_other = 0;
      for (int i = 0; i < P.nops; ++i) {
        const ChainOp& o = P.ops[i];
        if (o.group != g || o.kind != CH_GEMM) continue;
        const int tm = (o.M + CH_BM - 1) / CH_BM;
        if (o.N % 128 == 0 && o.K >= 256) wide_tiles += tm * (o.N / 128);
        else narrow_other += tm * ((o.N + 63) / 64);
      }
      const bool wide = !g_opt.chain_no_bn128 && wide_tiles + narrow_other >= (num_sms() * 5) / 8;
      for (int i = 0; i < P.nops; ++i) {
        ChainOp& o = P.ops[i];
        if (o.group != g || o.kind != CH_GEMM) continue;
        o.bn = (wide && o.N % 128 == 0 && o.K >= 256) ? 128 : 64;
        o.tiles_m = (o.M + CH_BM - 1) / CH_BM;
        o.tiles_n = (o.N + o.bn - 1) / o.bn;
        int r = encode(&o.a_map, raw[i].A, o.a_mn, o.M, o.K, raw[i].lda, CH_BM);
        if (!r) r = encode(&o.b_map, raw[i].B, o.b_mn, o.N, o.K, raw[i].ldb, o.bn);
        if (r) return r;
      }
    }
    // ops were appended group by group, in order
    int g = 0;
    P.group_begin[0] = 0;
    for (int i = 0; i < P.nops; ++i) {
      if (P.ops[i].group < g) { set_error("chain: ops out of group order"); return DA_ERR_INVALID_ARG; }
      while (g < P.ops[i].group) P.group_begin[++g] = i;
    }
    while (g < P.ngroups) P.group_begin[++g] = P.nops;
    return DA_OK;
  }
};

static int launch_chain(Builder& b, unsigned int* barrier, cudaStream_t st) {
  int rc = b.finish();
  if (rc) return rc;
  b.P.barrier = barrier;
  b.P.seed_ctr = g_seed_counter;
  b.P.trace = reinterpret_cast<unsigned long long*>(g_opt.chain_trace);
  int max_tiles = 1;
  for (int g = 0; g < b.P.ngroups; ++g) {
    int t = 0;
    for (int o = b.P.group_begin[g]; o < b.P.group_begin[g + 1]; ++o)
      if (b.P.ops[o].kind == CH_GEMM) t += b.P.ops[o].tiles_m * b.P.ops[o].tiles_n;
    if (t > max_tiles) max_tiles = t;
  }
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[cur_dev()];
  if (!attr_set) {
    DA_CUDA_OK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
    attr_set = true;
  }
  DA_CUDA_OK(cudaMemsetAsync(barrier, 0, 16, st));
  // every CTA must be resident at once (grid barrier): one CTA per SM, never more CTAs than SMs; elementwise ops want the
  // whole machine, so the grid is not shrunk to the tile count
  const int grid = num_sms();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt.no_pdl ? 0 : 1;
  DA_CUDA_OK(cudaLaunchKernelEx(&cfg, chain_kernel, b.P));
  DA_LAUNCH_CHECK();
  (void)max_tiles;
  return DA_OK;
}

}  // namespace da

using namespace da;

static inline int pad8(int v) { return (v + 7) & ~7; }

extern "C" size_t da_instance_fc_workspace_bytes(int R) {
  // barrier words | S / dP fp32 [R, pad8(R)] | dS bf16 [R, pad8(R)] | dz padded bf16 [R, 8]
  const size_t ld = (size_t)pad8(R);
  return 256 + align_up((size_t)R * ld * 4, 256) + align_up((size_t)R * ld * 2, 256) + align_up((size_t)R * 16, 256);
}

static int check_desc(const da_instance_fc_desc* d, const char* who) {
  DA_REQUIRE(d != nullptr, DA_ERR_INVALID_ARG, "%s: null descriptor", who);
  DA_REQUIRE(d->R > 0 && d->C > 0 && d->H1 > 0 && d->H2 > 0, DA_ERR_INVALID_ARG, "%s: bad sizes", who);
  DA_REQUIRE(d->C % 64 == 0 && d->H1 % 64 == 0 && d->H2 % 64 == 0 && (!d->nlb || (d->I > 0 && d->I % 64 == 0)), DA_ERR_UNSUPPORTED,
             "%s: channel counts must be multiples of 64 (C=%d I=%d H1=%d H2=%d)", who, d->C, d->I, d->H1, d->H2);
  DA_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f, DA_ERR_INVALID_ARG, "%s: drop_p out of [0,1)", who);
  DA_REQUIRE(d->C0 >= 0 && d->C0 % 64 == 0, DA_ERR_UNSUPPORTED, "%s: C0=%d must be a multiple of 64 (0 = no feeding layer)", who, d->C0);
  return DA_OK;
}

extern "C" int da_instance_fc_forward(const da_instance_fc_desc* d, const da_instance_fc_tensors* t, void* workspace,
                                      size_t workspace_bytes, da_stream_t stream) {
  int rc = check_desc(d, "instance_fc_forward");
  if (rc) return rc;
  DA_REQUIRE(t && t->x && t->w1 && t->w2 && t->w3 && t->labels && t->h1 && t->h2 && t->z && t->pred && t->loss, DA_ERR_INVALID_ARG,
             "instance_fc_forward: null tensor");
  DA_REQUIRE(!d->nlb || (t->w_proj && t->w_mask && t->proj && t->attn && t->y && t->t), DA_ERR_INVALID_ARG,
             "instance_fc_forward: null NonLocalBlock tensor");
  DA_REQUIRE(workspace && workspace_bytes >= da_instance_fc_workspace_bytes(d->R), DA_ERR_WORKSPACE, "instance_fc_forward: workspace too small");
  const int R = d->R, C = d->C, I = d->I, H1 = d->H1, H2 = d->H2, ldp = pad8(R);
  uint8_t* ws = (uint8_t*)workspace;
  unsigned int* barrier = (unsigned int*)ws;
  float* S = (float*)(ws + 256);
  DA_REQUIRE(d->C0 == 0 || (t->xin && t->w0), DA_ERR_INVALID_ARG, "instance_fc_forward: null feeding-layer tensor");
  Builder b;
  int g = 0;
  // feeding layer: x = relu(xin W0^T + b0)                                                (convfc_bbox_head.py:229-237)
  if (d->C0 > 0)
    if (ChainOp* o = b.gemm(g++, t->xin, 0, d->C0, t->w0, 0, d->C0, R, C, d->C0, const_cast<void*>(t->x), 0, C)) { o->bias = t->b0; o->relu = 1; }
  const void* feat = t->x;      // input of FC1
  if (d->nlb) {
    const __nv_bfloat16* proj = (const __nv_bfloat16*)t->proj;
    // [theta | phi | g] = x * Wcat^T                                                     (instance_da.py:163-169)
    b.gemm(g++, t->x, 0, C, t->w_proj, 0, C, R, 3 * I, C, t->proj, 0, 3 * I);
    // S[q,k] = theta[q,:] . phi[k,:]                                                     (:170)
    b.gemm(g++, proj, 0, 3 * I, proj + I, 0, 3 * I, R, R, I, S, 1, ldp);
    // softmax over the QUERY axis (nn.Softmax(dim=1) on [b,q,k], :171)
    if (ChainOp* o = b.add(CH_SOFTMAX_COL_FWD, g++)) { o->p0 = S; o->q0 = t->attn; o->i0 = R; o->i1 = ldp; o->i2 = ldp; }
    // Y = P * g                                                                          (:172-173)
    b.gemm(g++, t->attn, 0, ldp, proj + 2 * I, 1, 3 * I, R, I, R, t->y, 0, I);
    // t = Y * Wmask^T + x                                                                (:174-175)
    if (ChainOp* o = b.gemm(g++, t->y, 0, I, t->w_mask, 0, I, R, C, I, t->t, 0, C)) { o->res = (const __nv_bfloat16*)t->x; o->ld_res = C; }
    feat = t->t;
  }
  // fc1 -> ReLU -> dropout, fc2 -> ReLU -> dropout, fc3                                  (:73-81, :125-131)
  if (ChainOp* o = b.gemm(g++, feat, 0, C, t->w1, 0, C, R, H1, C, t->h1, 0, H1)) { o->bias = t->b1; o->relu = 1; o->drop_p = d->drop_p; o->seed = d->seed1; }
  if (ChainOp* o = b.gemm(g++, t->h1, 0, H1, t->w2, 0, H1, R, H2, H1, t->h2, 0, H2)) { o->bias = t->b2; o->relu = 1; o->drop_p = d->drop_p; o->seed = d->seed2; }
  if (ChainOp* o = b.gemm(g++, t->h2, 0, H2, t->w3, 0, H2, R, 2, H2, t->z, 1, 2)) o->bias = t->b3;
  // sigmoid + CrossEntropyLoss on the sigmoid outputs (instance_da.py:82-85 + DAFaster_rcnn_Orig.py:177-188)
  if (ChainOp* o = b.add(CH_CE_FWD, g++)) { o->p0 = t->z; o->p1 = t->labels; o->q0 = t->pred; o->q1 = t->loss; o->i0 = R; }
  return launch_chain(b, barrier, (cudaStream_t)stream);
}

extern "C" int da_instance_fc_backward(const da_instance_fc_desc* d, const da_instance_fc_tensors* t, const da_instance_fc_grads* gr,
                                       void* workspace, size_t workspace_bytes, da_stream_t stream) {
  int rc = check_desc(d, "instance_fc_backward");
  if (rc) return rc;
  DA_REQUIRE(t && gr && t->x && t->w1 && t->w2 && t->w3 && t->labels && t->h1 && t->h2 && t->z, DA_ERR_INVALID_ARG, "instance_fc_backward: null tensor");
  DA_REQUIRE(gr->dx && gr->dw1 && gr->dw2 && gr->dw3 && gr->db1 && gr->db2 && gr->db3 && gr->dz2 && gr->dz1, DA_ERR_INVALID_ARG,
             "instance_fc_backward: null gradient buffer");
  DA_REQUIRE(!d->nlb || (t->w_proj && t->w_mask && t->proj && t->attn && t->y && t->t && gr->dw_proj && gr->dw_mask && gr->dt && gr->dy && gr->dproj),
             DA_ERR_INVALID_ARG, "instance_fc_backward: null NonLocalBlock tensor");
  DA_REQUIRE(d->C0 == 0 || (t->xin && t->w0 && gr->dxin && gr->dw0 && gr->db0 && (!d->gate_in || gr->db_in)), DA_ERR_INVALID_ARG,
             "instance_fc_backward: null feeding-layer tensor");
  DA_REQUIRE(workspace && workspace_bytes >= da_instance_fc_workspace_bytes(d->R), DA_ERR_WORKSPACE, "instance_fc_backward: workspace too small");
  const int R = d->R, C = d->C, I = d->I, H1 = d->H1, H2 = d->H2, ldp = pad8(R);
  uint8_t* ws = (uint8_t*)workspace;
  unsigned int* barrier = (unsigned int*)ws;
  float* dP = (float*)(ws + 256);
  __nv_bfloat16* dS = (__nv_bfloat16*)(ws + 256 + align_up((size_t)R * ldp * 4, 256));
  __nv_bfloat16* dzp = (__nv_bfloat16*)((uint8_t*)dS + align_up((size_t)R * ldp * 2, 256));
  const float keep = d->drop_p > 0.f ? 1.f / (1.f - d->drop_p) : 1.f;
  const void* feat = d->nlb ? t->t : t->x;
  Builder b;
  int g = 0;
  // dz = dCE/dz (+ dpred through the sigmoid)
  if (ChainOp* o = b.add(CH_CE_BWD, g++)) {
    o->p0 = t->z; o->p1 = t->labels; o->p2 = gr->grad_loss; o->p3 = gr->grad_pred; o->q0 = dzp; o->q1 = nullptr; o->i0 = R; o->f0 = gr->loss_scale;
  }
  // fc3: dW3 = dz^T h2, db3 = colsum(dz), dz2 = (dz W3) * relu'/dropout mask of h2
  b.gemm(g, dzp, 1, 8, t->h2, 1, H2, 2, H2, R, gr->dw3, 1, H2);
  if (ChainOp* o = b.gemm(g, dzp, 0, 8, t->w3, 1, H2, R, H2, 2, gr->dz2, 0, H2)) { o->gate = (const __nv_bfloat16*)t->h2; o->ld_gate = H2; o->gate_scale = keep; }
  if (ChainOp* o = b.add(CH_COLSUM, g++)) { o->p0 = dzp; o->q0 = gr->db3; o->i0 = R; o->i1 = 8; o->i2 = 2; }
  // fc2
  b.gemm(g, gr->dz2, 1, H2, t->h1, 1, H1, H2, H1, R, gr->dw2, 1, H1);
  if (ChainOp* o = b.gemm(g, gr->dz2, 0, H2, t->w2, 1, H1, R, H1, H2, gr->dz1, 0, H1)) { o->gate = (const __nv_bfloat16*)t->h1; o->ld_gate = H1; o->gate_scale = keep; }
  if (ChainOp* o = b.add(CH_COLSUM, g++)) { o->p0 = gr->dz2; o->q0 = gr->db2; o->i0 = R; o->i1 = H2; o->i2 = H2; }
  // fc1: dW1 = dz1^T feat, db1, d(feat) = dz1 W1
  b.gemm(g, gr->dz1, 1, H1, feat, 1, C, H1, C, R, gr->dw1, 1, C);
  if (d->nlb) {
    b.gemm(g, gr->dz1, 0, H1, t->w1, 1, C, R, C, H1, gr->dt, 0, C);
  } else {
    if (ChainOp* o = b.gemm(g, gr->dz1, 0, H1, t->w1, 1, C, R, C, H1, gr->dx, 0, C)) {
      o->alpha = d->grl;   // reversed gradient leaves here
      if (d->C0 > 0) { o->gate = (const __nv_bfloat16*)t->x; o->ld_gate = C; }
    }
  }
  if (ChainOp* o = b.add(CH_COLSUM, g++)) { o->p0 = gr->dz1; o->q0 = gr->db1; o->i0 = R; o->i1 = H1; o->i2 = H1; }
  if (d->nlb) {
    const __nv_bfloat16* proj = (const __nv_bfloat16*)t->proj;
    __nv_bfloat16* dproj = (__nv_bfloat16*)gr->dproj;
    // conv_mask: dWmask = dt^T Y, dY = dt Wmask
    b.gemm(g, gr->dt, 1, C, t->y, 1, I, C, I, R, gr->dw_mask, 1, I);
    b.gemm(g++, gr->dt, 0, C, t->w_mask, 1, I, R, I, C, gr->dy, 0, I);
    // Y = P g: dP = dY g^T (fp32), dg = P^T dY
    b.gemm(g, gr->dy, 0, I, proj + 2 * I, 0, 3 * I, R, R, I, dP, 1, ldp);
    b.gemm(g++, t->attn, 1, ldp, gr->dy, 1, I, R, I, R, dproj + 2 * I, 0, 3 * I);
    // softmax over q: dS = P * (dP - sum_q P dP)
    if (ChainOp* o = b.add(CH_SOFTMAX_COL_BWD, g++)) { o->p0 = t->attn; o->p1 = dP; o->q0 = dS; o->i0 = R; o->i1 = ldp; o->i2 = ldp; o->i3 = ldp; }
    // S = theta phi^T: dtheta = dS phi, dphi = dS^T theta
    b.gemm(g, dS, 0, ldp, proj + I, 1, 3 * I, R, I, R, dproj, 0, 3 * I);
    b.gemm(g++, dS, 1, ldp, proj, 1, 3 * I, R, I, R, dproj + I, 0, 3 * I);
    // projections: dWcat = dproj^T x, dx = grl * (dproj Wcat + dt)   (residual path; GRL weight folded, instance_da.py:20-23)
    b.gemm(g, dproj, 1, 3 * I, t->x, 1, C, 3 * I, C, R, gr->dw_proj, 1, C);
    if (ChainOp* o = b.gemm(g++, dproj, 0, 3 * I, t->w_proj, 1, C, R, C, 3 * I, gr->dx, 0, C)) {
      o->res = (const __nv_bfloat16*)gr->dt; o->ld_res = C; o->alpha = d->grl;
      if (d->C0 > 0) { o->gate = (const __nv_bfloat16*)t->x; o->ld_gate = C; }
    }
  }
  if (d->C0 > 0) {
    // feeding layer: dx now holds dz0 = grl * d(x) * (x > 0); dW0 = dz0^T xin, db0 = colsum(dz0), dxin = dz0 W0 [* (xin > 0)]
    const int C0 = d->C0;
    b.gemm(g, gr->dx, 1, C, t->xin, 1, C0, C, C0, R, gr->dw0, 1, C0);
    if (ChainOp* o = b.gemm(g, gr->dx, 0, C, t->w0, 1, C0, R, C0, C, gr->dxin, 0, C0))
      if (d->gate_in) { o->gate = (const __nv_bfloat16*)t->xin; o->ld_gate = C0; }
    if (ChainOp* o = b.add(CH_COLSUM, g++)) { o->p0 = gr->dx; o->q0 = gr->db0; o->i0 = R; o->i1 = C; o->i2 = C; }
    if (d->gate_in)
      if (ChainOp* o = b.add(CH_COLSUM, g++)) { o->p0 = gr->dxin; o->q0 = gr->db_in; o->i0 = R; o->i1 = C0; o->i2 = C0; }
  }
  return launch_chain(b, barrier, (cudaStream_t)stream);
}
