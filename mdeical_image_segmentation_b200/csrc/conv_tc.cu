// tcgen05 / TMEM / TMA implicit-GEMM kernels for the U-Net convolutions (sm_100a only).
//
// Data layout in HBM: activations are channels-last bf16 (N, D, H, W, C); one 128-row GEMM tile is
// a TD x TH x TW box of positions, fetched by ONE TMA tile-mode load per (tap, 64-channel chunk):
// the box start is shifted by the tap offset and the hardware zero-fills everything that falls
// outside the tensor, which is exactly the conv's zero padding. The box lands in shared memory as
// 128 rows x 128 B with the 128-byte swizzle, i.e. the canonical K-major UMMA operand.
// Weights are pre-packed [tap][n][k] bf16 so a (tap, k-chunk, n-tile) slab is one 3-D TMA box.
//
// Kernel anatomy (one CTA per SM, persistent over tiles):
//   warp 0 lane 0 : TMA producer      (full/empty mbarrier ring of `stages` slots)
//   warp 1 lane 0 : tcgen05.mma issuer (accumulators double-buffered in TMEM)
//   warp 2        : TMEM allocator
//   warps 4..7    : epilogue: tcgen05.ld -> bias / ReLU / mask -> bf16 -> global (concat slice)
#include "conv_tc.cuh"

#include "tmap.cuh"
#include <type_traits>

namespace b200 {

static const int kSmemBudget = 227 * 1024;
static const int kTileRows = 128;
static const int kABytes = kTileRows * 128;  // 128 rows x 64 bf16

struct TileGeom {
  int tdl, thl, twl;
  int tiles_d, tiles_h, tiles_w;
  int m_tiles;
};

// pick a power-of-two TD x TH x TW = 128 box minimising padded work, then halo surface
static TileGeom pick_tile(int pn, int pd, int ph, int pw, bool force_td1) {
  TileGeom best{};
  double best_cost = 1e300;
  for (int tdl = 0; tdl <= 7; ++tdl)
    for (int thl = 0; tdl + thl <= 7; ++thl) {
      int twl = 7 - tdl - thl;
      int td = 1 << tdl, th = 1 << thl, tw = 1 << twl;
      if (force_td1 && td != 1) continue;
      if (td > 2 * pd && td > 1) continue;
      long nd = (pd + td - 1) / td, nh = (ph + th - 1) / th, nw = (pw + tw - 1) / tw;
      double padded = (double)nd * td * nh * th * nw * tw;
      double halo = (double)(td + 2) * (th + 2) * (tw + 2) / 128.0;  // L2 re-read factor proxy
      double cost = padded * (1.0 + 0.02 * halo) + (tw < 4 ? padded * 0.05 : 0.0);
      if (cost < best_cost) {
        best_cost = cost;
        best.tdl = tdl;
        best.thl = thl;
        best.twl = twl;
        best.tiles_d = (int)nd;
        best.tiles_h = (int)nh;
        best.tiles_w = (int)nw;
        best.m_tiles = (int)(pn * nd * nh * nw);
      }
    }
  return best;
}

static int make_act_tmap(CUtensorMap* tm, const ActView& v, int tap_mode, const TileGeom& g) {
  uint64_t dims[5], str[5];
  uint32_t box[5];
  const bf16* base = v.ptr + v.coff;
  if (tap_mode == TC_TAP_SHIFT) {
    dims[0] = v.c; dims[1] = v.w; dims[2] = v.h; dims[3] = v.d; dims[4] = v.n;
    str[0] = 1; str[1] = v.ctot; str[2] = (uint64_t)v.w * v.ctot;
    str[3] = (uint64_t)v.h * v.w * v.ctot; str[4] = (uint64_t)v.d * v.h * v.w * v.ctot;
    box[0] = 64; box[1] = 1u << g.twl; box[2] = 1u << g.thl; box[3] = 1u << g.tdl; box[4] = 1;
  } else {
    // (N, 1, 2H, 2W, C) seen as (C, j, W, i, N*H)
    int H = v.h / 2, W = v.w / 2;
    dims[0] = v.c; dims[1] = 2; dims[2] = W; dims[3] = 2; dims[4] = (uint64_t)v.n * H;
    str[0] = 1; str[1] = v.ctot; str[2] = 2ull * v.ctot; str[3] = (uint64_t)v.w * v.ctot;
    str[4] = 2ull * v.w * v.ctot;
    box[0] = 64; box[1] = 1; box[2] = 1u << g.twl; box[3] = 1; box[4] = 1u << g.thl;
  }
  return make_tmap_bf16(tm, base, 5, dims, str, box);
}

struct TileCoord {
  int n, d0, h0, w0;
};
__device__ __forceinline__ TileCoord decode_tile(int mt, int tiles_d, int tiles_h, int tiles_w,
                                                 int tdl, int thl, int twl) {
  TileCoord t;
  int tw = mt % tiles_w; mt /= tiles_w;
  int th = mt % tiles_h; mt /= tiles_h;
  int td = mt % tiles_d; mt /= tiles_d;
  t.n = mt; t.d0 = td << tdl; t.h0 = th << thl; t.w0 = tw << twl;
  return t;
}


// ===================================================================== epilogue (both GEMM kernels)
// Accumulator rows live one per TMEM lane, so tcgen05.ld hands every thread one output position and 32
// consecutive channels: storing that directly makes each warp-wide 16-byte store touch 32 different
// 128-byte lines.  Instead each epilogue warp packs its 32 rows x 32 channels to bf16 in a private
// 2 KB shared-memory buffer (XOR-swizzled, conflict-free) and reads it back with 4 lanes per row, so
// a store instruction covers 8 rows x 64 contiguous bytes, the ReLU-mask / dot operands are fetched
// with the same coalesced pattern, and a lane always owns the same 8 channels — which makes the
// per-channel reductions that follow a convolution (bias gradient, BatchNorm / GroupNorm statistics
// and their backward sums) a few register adds in the epilogue instead of extra passes over HBM.
static const int kEpiStageBytes = 2048;   // per epilogue warp: 32 rows x 64 B
// control region at the start of dynamic shared memory: barriers | per-warp staging | column sums
__host__ __device__ inline int epi_ctrl_bytes(int epi_warps) { return 1024 + epi_warps * kEpiStageBytes + 2048; }
static inline int smem_overhead(int epi_warps) { return epi_ctrl_bytes(epi_warps) + 1024; }  // + alignment slack

// statistics accumulated per (sample, output channel) over the stored result r
enum { EPI_STATS_NONE = 0, EPI_STATS_SUM = 1, EPI_STATS_SUMSQ = 2, EPI_STATS_DOT = 3 };

struct EpiArgs {
  bf16* out;
  int out_ctot, out_coff;
  const float* bias;
  int relu;
  const bf16* mask;          // out = 0 where mask <= 0 (geometry of out)
  int mask_ctot, mask_coff;
  const bf16* dot;           // EPI_STATS_DOT: second statistic = sum out * dot
  int dot_ctot, dot_coff;
  double* stats;             // fp64 [pn][2][stats_cols] accumulated into, or NULL
  int stats_cols;
  int epi_warps;             // 4 or 8 epilogue warps (two warps per TMEM lane quarter split the chunks)
  int dbg;                   // bring-up knob (b200_set_tuning key 4): 1 = no global stores, 2 = no epilogue work
};

__device__ __forceinline__ void epi_bar(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t bf16x2_gt0_mask(uint32_t m) {
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&m);
  return __hgt2_mask(v, __float2bfloat162_rn(0.f));
}
__device__ __forceinline__ uint32_t bf16x2_relu(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  h = __hmax2(h, __float2bfloat162_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Mask / dot operands of one chunk, fetched one chunk ahead of their use so that the global-load
// latency hides behind the TMEM -> shared-memory staging of the previous chunk.
template <bool MASK, int STATS>
struct EpiAux {
  uint4 m[MASK ? 4 : 1];
  uint4 d[STATS == EPI_STATS_DOT ? 4 : 1];
};
// Row placement of one chunk for the read-back phase: lane (4 lanes per row) stores rows it = 0..3.
// Element offsets (position * ctot + coff) of each row's first channel in out / mask / dot; -1 = no row.
struct EpiRows {
  long ooff[4], moff[4], doff[4];
};
template <bool MASK, int STATS>
__device__ __forceinline__ void epi_prefetch(const EpiArgs& e, int col, int ncol, int lane,
                                             const EpiRows& rw, EpiAux<MASK, STATS>& a) {
  if ((lane & 3) * 8 >= ncol) return;   // 16-column chunks: lanes 2,3 of each row own nothing
  const int c = col + (lane & 3) * 8;
  if (MASK) {
#pragma unroll
    for (int it = 0; it < 4; ++it)   // rows that do not exist read row 0 of the tensor (never used)
      a.m[it] = ldg_nc_v4(e.mask + (rw.moff[it] < 0 ? (long)e.mask_coff : rw.moff[it]) + c);
  }
  if (STATS == EPI_STATS_DOT) {
#pragma unroll
    for (int it = 0; it < 4; ++it)
      a.d[it] = ldg_nc_v4(e.dot + (rw.doff[it] < 0 ? (long)e.dot_coff : rw.doff[it]) + c);
  }
}

// One warp, one chunk: v = 32 (or 16) accumulator columns of row `lane`.  col = first output channel
// of the chunk.  An epilogue warp shares its scheduler with at most one other epilogue warp, so every
// instruction costs most of its latency: the variants are compile-time (MASK, STATS) so that each
// carries only the code it needs, and all addresses are prepared by the caller.
template <bool MASK, int STATS>
__device__ __forceinline__ void epi_chunk(const EpiArgs& e, const uint32_t (&v)[32], int ncol, int col,
                                          uint8_t* stg, int lane, const EpiRows& rw, float (&acc)[16],
                                          const EpiAux<MASK, STATS>& aux) {
  uint32_t pk[16];
  const float* bias = e.bias;
  const bool relu = e.relu != 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h * 16 < ncol) {
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[h * 16 + j]);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + col + h * 16 + j));
        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[h * 8 + j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
      if (relu) {   // max(round(x), 0) == round(max(x, 0)): clamp the packed pairs (half the instructions)
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[h * 8 + j] = bf16x2_relu(pk[h * 8 + j]);
      }
    }
  }
  {
    const int sw = (lane >> 1) & 3;
    uint4* srow = reinterpret_cast<uint4*>(stg + lane * 64);
    srow[0 ^ sw] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    srow[1 ^ sw] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    if (ncol > 16) {
      srow[2 ^ sw] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
      srow[3 ^ sw] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
    }
  }
  __syncwarp();
  const int ch = lane & 3, rsub = lane >> 2;
  if (ch * 8 < ncol) {
    bf16* const obase = e.out + (col + ch * 8);
    const uint8_t* const sbase = stg + rsub * 64 + ((ch ^ ((rsub >> 1) & 3)) << 4);
    const int dbg = e.dbg;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      uint4 u = *reinterpret_cast<const uint4*>(sbase + it * 512);   // row it*8 + rsub (same swizzle phase)
      const long off = rw.ooff[it];
      if (off >= 0) {
        if (MASK) {
          const uint4 m = aux.m[it];
          u.x &= bf16x2_gt0_mask(m.x); u.y &= bf16x2_gt0_mask(m.y);
          u.z &= bf16x2_gt0_mask(m.z); u.w &= bf16x2_gt0_mask(m.w);
        }
        if (STATS != EPI_STATS_NONE) {
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
          uint32_t dw[4] = {0, 0, 0, 0};
          if (STATS == EPI_STATS_DOT) { dw[0] = aux.d[it].x; dw[1] = aux.d[it].y; dw[2] = aux.d[it].z; dw[3] = aux.d[it].w; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = unpack_bf16x2(uw[j]);
            acc[2 * j] += a.x; acc[2 * j + 1] += a.y;
            if (STATS == EPI_STATS_SUMSQ) {
              acc[8 + 2 * j] = fmaf(a.x, a.x, acc[8 + 2 * j]);
              acc[8 + 2 * j + 1] = fmaf(a.y, a.y, acc[8 + 2 * j + 1]);
            } else if (STATS == EPI_STATS_DOT) {
              const float2 b = unpack_bf16x2(dw[j]);
              acc[8 + 2 * j] = fmaf(a.x, b.x, acc[8 + 2 * j]);
              acc[8 + 2 * j + 1] = fmaf(a.y, b.y, acc[8 + 2 * j + 1]);
            }
          }
        }
        if (dbg != 1) *reinterpret_cast<uint4*>(obase + off) = u;
      }
    }
  }
  __syncwarp();
}

// Fold the per-lane partial column sums of one chunk into the CTA's shared column sums [2][bn].
// The 8 lanes that own the same 8 columns (lane bits 2..4) hold 8 or 16 partials each; a halving
// exchange (send the half the partner keeps, add the half received) leaves every lane with 1 or 2
// finished sums after 3 steps — 14 shuffles instead of the 48 of a plain butterfly.
template <int STATS>
__device__ __forceinline__ void epi_fold_acc(float (&acc)[16], float* colsum, int c0, int ncol, int bn,
                                             int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  const int ch = lane & 3;
  if (STATS == EPI_STATS_SUM) {
    float a[4], b[2];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      a[k] = (b4 ? acc[4 + k] : acc[k]) + __shfl_xor_sync(0xffffffffu, b4 ? acc[k] : acc[4 + k], 16);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      b[k] = (b3 ? a[2 + k] : a[k]) + __shfl_xor_sync(0xffffffffu, b3 ? a[k] : a[2 + k], 8);
    const float c = (b2 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, b2 ? b[0] : b[1], 4);
    if (ch * 8 < ncol) atomicAdd(colsum + c0 + ch * 8 + (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0), c);
  } else {
    float a[8], b[4], c[2];
#pragma unroll
    for (int k = 0; k < 8; ++k)   // bit 4 picks the statistic
      a[k] = (b4 ? acc[8 + k] : acc[k]) + __shfl_xor_sync(0xffffffffu, b4 ? acc[k] : acc[8 + k], 16);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      b[k] = (b3 ? a[4 + k] : a[k]) + __shfl_xor_sync(0xffffffffu, b3 ? a[k] : a[4 + k], 8);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      c[k] = (b2 ? b[2 + k] : b[k]) + __shfl_xor_sync(0xffffffffu, b2 ? b[k] : b[2 + k], 4);
    if (ch * 8 < ncol) {
      float* dst = colsum + (b4 ? bn : 0) + c0 + ch * 8 + (b3 ? 4 : 0) + (b2 ? 2 : 0);
      atomicAdd(dst, c[0]);
      atomicAdd(dst + 1, c[1]);
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
}

// all epilogue threads: shared column sums [2][bn] -> global fp64 stats[n][2][stats_cols]
__device__ __forceinline__ void epi_flush_colsum(const EpiArgs& e, float* colsum, int bn, int n, int col0,
                                                 int tid, int nthr) {
  epi_bar(nthr);
  for (int i = tid; i < 2 * bn; i += nthr) {
    const float v = colsum[i];
    const int s = i >= bn, cidx = i - s * bn;
    if (v != 0.f) atomicAdd(e.stats + ((long)n * 2 + s) * e.stats_cols + col0 + cidx, (double)v);
    colsum[i] = 0.f;
  }
  epi_bar(nthr);
}

// ===================================================================== forward / dgrad GEMM
struct GemmArgs {
  int pn, pd, ph, pw;
  int tdl, thl, twl, tiles_d, tiles_h, tiles_w;
  int m_tiles, n_tiles, total;
  int ntaps, kh, kw, pad_d, pad_h, pad_w;
  int tap_mode, out_mode;
  int kchunks, kdim, BN, stages;
  int cout_t;
  int chunk;             // epilogue chunk width: 32 columns (16 when a scatter block is not a multiple of 32)
  int b_resident;        // the CTA keeps one column tile and all its weight slabs in shared memory
  EpiArgs e;
  int use_table;
  TapTable tt;
};

template <bool MASK, int STATS>
__global__ void __launch_bounds__(384, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  uint64_t* tfull = empty + 16;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* tiles = smem + epi_ctrl_bytes(p.e.epi_warps);
  const int b_tile_bytes = p.BN * 128;
  const int stage_bytes = p.b_resident ? kABytes : kABytes + b_tile_bytes;
  uint8_t* bres = tiles + (size_t)p.stages * stage_bytes;   // resident weight slabs [tap][kchunk]
  uint64_t* bres_full = &full[15];
  // work items: (m tile, n tile) pairs strided over the grid, or — weight-resident — the m tiles of ONE
  // n tile per CTA (grid is a multiple of n_tiles)
  const int wstride = p.b_resident ? (int)gridDim.x / p.n_tiles : (int)gridDim.x;
  const int wfirst = p.b_resident ? (int)blockIdx.x / p.n_tiles : (int)blockIdx.x;
  const int wcount = p.b_resident ? p.m_tiles : p.total;
  const int nt_fixed = (int)blockIdx.x % p.n_tiles;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(bres_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 32 * p.e.epi_warps);
    }
    fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * p.BN)) tmem_cols <<= 1;
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kb = p.ntaps * p.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    if (p.b_resident && wfirst < wcount && elect_one()) {
      mbar_arrive_expect_tx(bres_full, (uint32_t)(num_kb * b_tile_bytes));
      for (int tap = 0; tap < p.ntaps; ++tap)
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_3d(bres + (size_t)(tap * p.kchunks + kc) * b_tile_bytes, &tmB, bres_full, kc * 64,
                      nt_fixed * p.BN, p.use_table ? p.tt.widx[tap] : tap);
    }
    __syncwarp();
    for (int work = wfirst; work < wcount; work += wstride) {
      const int mt = p.b_resident ? work : work / p.n_tiles;
      const int nt = p.b_resident ? nt_fixed : work - mt * p.n_tiles;
      const TileCoord tc = decode_tile(mt, p.tiles_d, p.tiles_h, p.tiles_w, p.tdl, p.thl, p.twl);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int kwi = tap % p.kw, khi = (tap / p.kw) % p.kh, kdi = tap / (p.kw * p.kh);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sA = tiles + (size_t)stage * stage_bytes;
            uint8_t* sB = sA + kABytes;
            mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
            int wtap = tap;
            if (p.use_table) {
              tma_load_5d(sA, &tmA, &full[stage], kc * 64, tc.w0 + p.tt.dw[tap], tc.h0 + p.tt.dh[tap],
                          tc.d0 + p.tt.dd[tap], tc.n + p.tt.dn[tap]);
              wtap = p.tt.widx[tap];
            } else if (p.tap_mode == TC_TAP_SHIFT)
              tma_load_5d(sA, &tmA, &full[stage], kc * 64, tc.w0 + kwi - p.pad_w,
                          tc.h0 + khi - p.pad_h, tc.d0 + kdi - p.pad_d, tc.n);
            else
              tma_load_5d(sA, &tmA, &full[stage], kc * 64, kwi, tc.w0, khi, tc.n * p.ph + tc.h0);
            if (!p.b_resident) tma_load_3d(sB, &tmB, &full[stage], kc * 64, nt * p.BN, wtap);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0, acc = 0, acc_phase = 0;
    if (p.b_resident && wfirst < wcount) mbar_wait(bres_full, 0);
    for (int work = wfirst; work < wcount; work += wstride) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int kc = kb % p.kchunks;
        int nk = (p.kdim - kc * 64 + 15) >> 4;
        nk = nk > 4 ? 4 : nk;
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t aaddr = smem_u32(tiles + (size_t)stage * stage_bytes);
        if (elect_one()) {
          const uint64_t da = make_smem_desc_sw128(aaddr, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(
              p.b_resident ? smem_u32(bres + (size_t)kb * b_tile_bytes) : aaddr + kABytes, 16, 1024);
#pragma unroll 4
          for (int k = 0; k < nk; ++k)
            umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4 && warp < 4 + p.e.epi_warps) {
    // ------------------------------------------------------------------ epilogue (see epi_chunk)
    // warp ew reads TMEM lanes 32*(ew%4).. (hardware rule); with 8 warps the two warps of a lane
    // quarter take alternate chunks
    const int ew = warp - 4, q = ew & 3, grp = ew >> 2, ngrp = p.e.epi_warps >> 2;
    const int tid = threadIdx.x - 128, nthr = 32 * p.e.epi_warps, rsub = lane >> 2;
    uint8_t* stg = smem + 1024 + ew * kEpiStageBytes;
    float* colsum = reinterpret_cast<float*>(smem + 1024 + p.e.epi_warps * kEpiStageBytes);
    const EpiArgs& e = p.e;
    constexpr bool kStats = STATS != EPI_STATS_NONE;
    constexpr bool kAux = MASK || STATS == EPI_STATS_DOT;
    const int BN = p.BN, chunk = p.chunk;
    const long out_ctot = e.out_ctot, out_coff = e.out_coff;
    if (kStats) {
      for (int i = tid; i < 2 * BN; i += nthr) colsum[i] = 0.f;
      epi_bar(nthr);
    }
    int cur_n = -1, cur_nt = -1;
    float sacc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) sacc[k] = 0.f;
    const int nchunks = (BN + chunk - 1) / chunk;
    uint32_t acc = 0, acc_phase = 0;
    for (int work = wfirst; work < wcount; work += wstride) {
      const int mt = p.b_resident ? work : work / p.n_tiles;
      const int nt = p.b_resident ? nt_fixed : work - mt * p.n_tiles;
      const TileCoord tc = decode_tile(mt, p.tiles_d, p.tiles_h, p.tiles_w, p.tdl, p.thl, p.twl);
      if (kStats && (tc.n != cur_n || nt != cur_nt)) {
        if (cur_n >= 0) epi_flush_colsum(e, colsum, BN, cur_n, cur_nt * BN, tid, nthr);
        cur_n = tc.n; cur_nt = nt;
      }
      if (kAux && work + wstride < wcount && p.out_mode == TC_OUT_CONV) {
        // pull the next tile's mask / dot lines into L2 (see the halo kernel)
        const int w2 = work + wstride;
        const int mt2 = p.b_resident ? w2 : w2 / p.n_tiles;
        const int nt2 = p.b_resident ? nt_fixed : w2 - mt2 * p.n_tiles;
        const TileCoord t2 = decode_tile(mt2, p.tiles_d, p.tiles_h, p.tiles_w, p.tdl, p.thl, p.twl);
        const int nl = (BN * 2 + 127) >> 7;
        for (int i = tid; i < 128 * nl; i += nthr) {
          const int l = i % nl, row = i / nl;
          const int w_ = t2.w0 + (row & ((1 << p.twl) - 1));
          const int h_ = t2.h0 + ((row >> p.twl) & ((1 << p.thl) - 1));
          const int d_ = t2.d0 + (row >> (p.twl + p.thl));
          if (w_ < p.pw && h_ < p.ph && d_ < p.pd) {
            const long pos = (((long)t2.n * p.pd + d_) * p.ph + h_) * p.pw + w_;
            if (MASK) prefetch_l2(e.mask + pos * e.mask_ctot + e.mask_coff + nt2 * BN + l * 64);
            if (STATS == EPI_STATS_DOT) prefetch_l2(e.dot + pos * e.dot_ctot + e.dot_coff + nt2 * BN + l * 64);
          }
        }
      }
      // the 4 tile rows this lane stores in the read-back phase: q*32 + it*8 + lane/4
      long cpos[4];       // conv mode: output position; scatter mode: position of pixel (2h, 2w)
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = q * 32 + it * 8 + rsub;
        const int wx = tc.w0 + (row & ((1 << p.twl) - 1));
        const int hy = tc.h0 + ((row >> p.twl) & ((1 << p.thl) - 1));
        const int dz = tc.d0 + (row >> (p.twl + p.thl));
        const bool valid = (wx < p.pw) && (hy < p.ph) && (dz < p.pd);
        if (p.out_mode == TC_OUT_CONV)
          cpos[it] = valid ? (((long)tc.n * p.pd + dz) * p.ph + hy) * p.pw + wx : -1;
        else
          cpos[it] = valid ? ((long)tc.n * (2 * p.ph) + 2 * hy) * (2 * p.pw) + 2 * wx : -1;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)BN;
      auto issue = [&](uint32_t (&v)[32], int ci) {
        const int c0 = ci * chunk;
        tmem_ld16(taddr + c0, v);
        if (chunk == 32 && BN - c0 > 16) tmem_ld16(taddr + c0 + 16, v + 16);
      };
      // output column and rows of chunk ci (scatter mode: column block ij -> pixel (2h+i, 2w+j))
      auto place = [&](int ci, int& ocol, EpiRows& rw) {
        ocol = nt * BN + ci * chunk;
        long shift = 0;
        if (p.out_mode != TC_OUT_CONV) {
          const int ij = ocol / p.cout_t;
          ocol -= ij * p.cout_t;
          shift = (long)(ij >> 1) * (2 * p.pw) + (ij & 1);
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          rw.ooff[it] = cpos[it] < 0 ? -1 : (cpos[it] + shift) * out_ctot + out_coff;
          if (MASK) rw.moff[it] = cpos[it] < 0 ? -1 : cpos[it] * e.mask_ctot + e.mask_coff;
          if (STATS == EPI_STATS_DOT) rw.doff[it] = cpos[it] < 0 ? -1 : cpos[it] * e.dot_ctot + e.dot_coff;
        }
      };
      auto prefetch = [&](EpiAux<MASK, STATS>& aux, int ci) {
        if (!kAux) return;
        int ocol;
        EpiRows rw;
        place(ci, ocol, rw);
        epi_prefetch<MASK, STATS>(e, ocol, (chunk == 32 && BN - ci * chunk > 16) ? 32 : 16, lane, rw, aux);
      };
      auto process = [&](const uint32_t (&v)[32], const EpiAux<MASK, STATS>& aux, int ci) {
        const int c0 = ci * chunk;
        const int ncol = (chunk == 32 && BN - c0 > 16) ? 32 : 16;
        int ocol;
        EpiRows rw;
        place(ci, ocol, rw);
        epi_chunk<MASK, STATS>(e, v, ncol, ocol, stg, lane, rw, sacc, aux);
        if (kStats) epi_fold_acc<STATS>(sacc, colsum, c0, ncol, BN, lane);
      };
      uint32_t va[32], vb[32];
      EpiAux<MASK, STATS> xa, xb;
      if (e.dbg != 2 && grp < nchunks) {
        issue(va, grp);
        prefetch(xa, grp);
        for (int ci = grp; ci < nchunks; ci += 2 * ngrp) {
          const int c1 = ci + ngrp, c2 = c1 + ngrp;
          tmem_ld_wait();
          if (c1 < nchunks) { issue(vb, c1); prefetch(xb, c1); }
          process(va, xa, ci);
          if (c1 < nchunks) {
            tmem_ld_wait();
            if (c2 < nchunks) { issue(va, c2); prefetch(xa, c2); }
            process(vb, xb, c1);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (kStats && cur_n >= 0) epi_flush_colsum(e, colsum, BN, cur_n, cur_nt * BN, tid, nthr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ===================================================================== halo-reuse forward / dgrad
// For 3x3(x3) convolutions on planes of at least 16 x 8 positions the per-tap TMA loads above re-read
// every activation 9 (27) times from L2.  This variant fetches ONE halo box per 64-channel chunk —
// (TW+2) x (TH+2) x kd input positions for a 1 x 16 x 8 output tile — and feeds all taps from it:
// the UMMA A descriptor of tap (kd,kh,kw) simply starts at halo row ((kd*18 + kh)*10 + kw) with an
// 8-row-group stride (SBO) of 10 rows = 1280 B.  Measured on B200 (tools/desc_probe.cu,
// profiles/r01_desc_probe_kmajor.log): SWIZZLE_128B operands may start at any 128-byte row and use
// any SBO that is a multiple of 128 B, because the XOR pattern follows absolute shared-memory address
// bits — exactly how TMA wrote the box.  L2->SM traffic for A drops 6.3x (3D) / 6.4x (2D).
static const int kHaloW = 10, kHaloH = 18;
static const int kHaloPlaneBytes = kHaloW * kHaloH * 128;  // 23040

struct HaloArgs {
  int pn, pd, ph, pw;
  int tiles_h, tiles_w, m_tiles, n_tiles, total;
  int kd;                // 1 or 3 depth taps (kh = kw = 3)
  int kchunks, kdim, BN;
  int R;                 // 16x8 row-blocks per super-tile (1, 2 or 4): R accumulators share each weight slab
  int halo_h;            // 16 R + 2
  int a_stage_bytes, a_stages, b_slots;
  int G;                 // taps per weight slot (3 = one kw row per TMA box, or 1)
  int b_resident;        // all weight slabs of the (single) column tile fit in shared memory: loaded once
  EpiArgs e;
};

// MMA issuer of the halo kernel.  One thread issues every tcgen05.mma, and with N = 64 an MMA retires
// every ~28 ns, so the issue loop itself was the limiter (measured: ~200 ns of uniform-datapath
// bookkeeping per tap + ~27 ns per MMA with run-time R / tap loops).  R (row-blocks) and G (taps per
// weight slot) are template parameters and the 9 in-plane taps are unrolled, so every descriptor is
// base + compile-time constant and a tap's R x 4 MMAs are issued back to back.
template <int R, int G, int NK>   // NK = K-steps per 64-channel chunk when uniform (4, 2, 1), 0 = run time
__device__ __forceinline__ void halo_mma_warp(const HaloArgs& p, uint32_t tmem_base, uint32_t sA_addr,
                                              uint32_t sB_addr, uint64_t* afull, uint64_t* aempty,
                                              uint64_t* bfull, uint64_t* bempty, uint64_t* tfull,
                                              uint64_t* tempty) {
  const int BN = p.BN, kchunks = p.kchunks, kd = p.kd, a_stages = p.a_stages, b_slots = p.b_slots;
  const int kdim = p.kdim, total = p.total, resident = p.b_resident;
  const uint32_t a_stage_bytes = (uint32_t)p.a_stage_bytes;
  const uint32_t tap_bytes = (uint32_t)BN * 128u, b_bytes = tap_bytes * G;
  const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
  const uint64_t adesc0 = make_smem_desc_sw128(0, 16, kHaloW * 128);
  const uint64_t bdesc0 = make_smem_desc_sw128(0, 16, 1024);
  int stage = 0, slot = 0;
  uint32_t aphase = 0, bphase = 0, acc = 0, acc_phase = 0;
  bool first_tile = true;
  for (int work = blockIdx.x; work < total; work += gridDim.x) {
    mbar_wait(&tempty[acc], acc_phase ^ 1);
    tc_fence_after();
    const uint32_t tmem_d = tmem_base + acc * (uint32_t)(R * BN);
    for (int kc = 0; kc < kchunks; ++kc) {
      int nk = (kdim - kc * 64 + 15) >> 4;
      nk = nk > 4 ? 4 : nk;
      for (int kdi = 0; kdi < kd; ++kdi) {
        mbar_wait(&afull[stage], aphase);
        tc_fence_after();
        const uint64_t da0 = adesc0 + (uint64_t)((sA_addr + (uint32_t)stage * a_stage_bytes) >> 4);
        const bool last_plane = kc == kchunks - 1 && kdi == kd - 1;
        // all MMAs of taps [t0, t1) of this plane; weight slab of tap t sits at slot_base + (t - t0) slabs
        auto issue_taps = [&](int t0, int t1, uint32_t slab_addr) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (t < t0 || t >= t1) continue;
            const int khi = t / 3, kwi = t - khi * 3;
            const uint64_t db = bdesc0 + (uint64_t)((slab_addr + (uint32_t)(t - t0) * tap_bytes) >> 4);
            const uint32_t fresh = (kc | kdi | t) == 0 ? 0u : 1u;   // first MMA of a tile overwrites
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const uint64_t da = da0 + (uint64_t)((((r * 16 + khi) * kHaloW + kwi) * 128) >> 4);
              const uint32_t td = tmem_d + (uint32_t)(r * BN);
              if (NK > 0) {   // uniform chunks (C_in = 16, 32 or a multiple of 64): fully unrolled
                umma_bf16(td, da, db, idesc, fresh);
#pragma unroll
                for (int k = 1; k < NK; ++k) umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
              } else {
                for (int k = 0; k < nk; ++k)
                  umma_bf16(td, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k ? 1u : fresh);
              }
            }
          }
        };
        if (resident && !first_tile) {
          // weights already in place: the whole plane (9 taps x R x NK MMAs) goes out in one elected block
          if (elect_one()) {
            issue_taps(0, 9, sB_addr + (uint32_t)slot * b_bytes);
            umma_commit(&aempty[stage]);
            if (last_plane) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          slot += 9 / G;
          if (slot >= b_slots) { slot -= b_slots; bphase ^= 1; }
        } else {
#pragma unroll
          for (int g = 0; g < 9 / G; ++g) {   // one weight slot (G taps) per elected block
            mbar_wait(&bfull[slot], bphase);
            tc_fence_after();
            if (elect_one()) {
              issue_taps(g * G, g * G + G, sB_addr + (uint32_t)slot * b_bytes);
              if (!resident) umma_commit(&bempty[slot]);
              if (g == 9 / G - 1) {
                umma_commit(&aempty[stage]);
                if (last_plane) umma_commit(&tfull[acc]);
              }
            }
            __syncwarp();
            if (++slot == b_slots) { slot = 0; bphase ^= 1; }
          }
        }
        if (++stage == a_stages) { stage = 0; aphase ^= 1; }
      }
    }
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
    first_tile = false;
  }
}

// Super-tile = 1 x (16 R) x 8 output positions of one depth plane.  A stage = the halo of ONE input plane
// (10 x (16R+2) positions x 64 channels); a weight slot = G taps; each weight slab is applied to all R
// row-blocks (R accumulators in TMEM), which divides the weight traffic — the limiter of narrow-C_out
// layers once the activations are halo-shared — by R.
template <bool MASK, int STATS>
__global__ void __launch_bounds__(384, 1)
    tc_gemm_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const HaloArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem);
  uint64_t* aempty = afull + 8;
  uint64_t* bfull = aempty + 8;
  uint64_t* bempty = bfull + 16;
  uint64_t* tfull = bempty + 16;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* sA = smem + epi_ctrl_bytes(p.e.epi_warps);
  uint8_t* sB = sA + (size_t)p.a_stages * p.a_stage_bytes;
  const int tap_bytes = p.BN * 128;
  const int b_bytes = tap_bytes * p.G;
  const int ntaps = p.kd * 9;
  const int thl = p.R == 4 ? 6 : (p.R == 2 ? 5 : 4);   // log2(16 R)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < p.b_slots; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 32 * p.e.epi_warps); }
    fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * p.R * p.BN)) tmem_cols <<= 1;
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- A (halo plane) producer
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t plane_bytes = (uint32_t)(kHaloW * p.halo_h * 128);
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const int mt = work / p.n_tiles;
      const TileCoord tc = decode_tile(mt, p.pd, p.tiles_h, p.tiles_w, 0, thl, 3);
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int kdi = 0; kdi < p.kd; ++kdi) {
          mbar_wait(&aempty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&afull[stage], plane_bytes);
            tma_load_5d(sA + (size_t)stage * p.a_stage_bytes, &tmA, &afull[stage], kc * 64, tc.w0 - 1,
                        tc.h0 - 1, tc.d0 + kdi - (p.kd >> 1), tc.n);
          }
          __syncwarp();
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 3) {
    // ---------------------------------------------------------------- B (weights) producer
    int slot = 0;
    uint32_t phase = 0;
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const int nt = work % p.n_tiles;
      if (p.b_resident && work != (int)blockIdx.x) break;   // weights stay put after the first tile
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int tap = 0; tap < ntaps; tap += p.G) {
          mbar_wait(&bempty[slot], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bfull[slot], (uint32_t)b_bytes);
            tma_load_3d(sB + (size_t)slot * b_bytes, &tmB, &bfull[slot], kc * 64, nt * p.BN, tap);
          }
          __syncwarp();
          if (++slot == p.b_slots) { slot = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (halo_mma_warp)
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
#define HALO_MMA(RR, GG, KK) \
  halo_mma_warp<RR, GG, KK>(p, tmem_base, sA_addr, sB_addr, afull, aempty, bfull, bempty, tfull, tempty)
    const int nku = (p.kdim & 63) == 0 ? 4 : (p.kdim == 32 ? 2 : (p.kdim == 16 ? 1 : 0));
#define HALO_MMA_NK(RR, GG)                                   \
  do {                                                        \
    if (nku == 4) HALO_MMA(RR, GG, 4);                        \
    else if (nku == 2) HALO_MMA(RR, GG, 2);                   \
    else if (nku == 1) HALO_MMA(RR, GG, 1);                   \
    else HALO_MMA(RR, GG, 0);                                 \
  } while (0)
    if (p.G == 3) {
      if (p.R == 4) HALO_MMA_NK(4, 3);
      else if (p.R == 2) HALO_MMA_NK(2, 3);
      else HALO_MMA_NK(1, 3);
    } else {
      if (p.R == 4) HALO_MMA_NK(4, 1);
      else if (p.R == 2) HALO_MMA_NK(2, 1);
      else HALO_MMA_NK(1, 1);
    }
#undef HALO_MMA_NK
#undef HALO_MMA
  } else if (warp >= 4 && warp < 4 + p.e.epi_warps) {
    // ---------------------------------------------------------------- epilogue (see epi_chunk)
    const int ew = warp - 4, q = ew & 3, grp = ew >> 2, ngrp = p.e.epi_warps >> 2;
    const int tid = threadIdx.x - 128, nthr = 32 * p.e.epi_warps, rsub = lane >> 2;
    uint8_t* stg = smem + 1024 + ew * kEpiStageBytes;
    float* colsum = reinterpret_cast<float*>(smem + 1024 + p.e.epi_warps * kEpiStageBytes);
    const EpiArgs& e = p.e;
    constexpr bool kStats = STATS != EPI_STATS_NONE;
    constexpr bool kAux = MASK || STATS == EPI_STATS_DOT;
    const int BN = p.BN, R = p.R, pw = p.pw, ph = p.ph;
    const int rsh = R == 4 ? 2 : (R == 2 ? 1 : 0);
    const long out_ctot = e.out_ctot, out_coff = e.out_coff;
    const long row_elems = (long)pw * out_ctot;   // elements between vertically adjacent positions
    const long mrow_elems = (long)pw * e.mask_ctot, drow_elems = (long)pw * e.dot_ctot;
    if (kStats) {
      for (int i = tid; i < 2 * BN; i += nthr) colsum[i] = 0.f;
      epi_bar(nthr);
    }
    int cur_n = -1, cur_nt = -1;
    float sacc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) sacc[k] = 0.f;
    const int nchunks = (BN + 31) / 32;
    const int items = nchunks * R;   // item = chunk * R + row-block
    uint32_t acc = 0, acc_phase = 0;
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const int mt = work / p.n_tiles, nt = work - mt * p.n_tiles;
      const TileCoord tc = decode_tile(mt, p.pd, p.tiles_h, p.tiles_w, 0, thl, 3);
      if (kStats && (tc.n != cur_n || nt != cur_nt)) {
        if (cur_n >= 0) epi_flush_colsum(e, colsum, BN, cur_n, cur_nt * BN, tid, nthr);
        cur_n = tc.n; cur_nt = nt;
      }
      if (kAux && work + (int)gridDim.x < p.total) {
        // The mask / dot tensors are read once, straight from HBM, by 16-byte loads of a handful of
        // warps: far too little memory-level parallelism to cover DRAM latency.  Pull the NEXT tile's
        // lines into L2 now (a couple of prefetches per thread); the loads one chunk ahead then hit L2.
        const int w2 = work + (int)gridDim.x;
        const int mt2 = w2 / p.n_tiles, nt2 = w2 - mt2 * p.n_tiles;
        const TileCoord t2 = decode_tile(mt2, p.pd, p.tiles_h, p.tiles_w, 0, thl, 3);
        const int nl = (BN * 2 + 127) >> 7;                 // 128-byte lines per position
        const int nlines = 16 * R * 8 * nl;
        const long plane2 = ((long)t2.n * p.pd + t2.d0) * ph;
        for (int i = tid; i < nlines; i += nthr) {
          const int l = i % nl, pp = i / nl;
          const int w_ = t2.w0 + (pp & 7), h_ = t2.h0 + (pp >> 3);
          if (w_ < pw && h_ < ph) {
            const long pos = (plane2 + h_) * pw + w_;
            if (MASK) prefetch_l2(e.mask + pos * e.mask_ctot + e.mask_coff + nt2 * BN + l * 64);
            if (STATS == EPI_STATS_DOT) prefetch_l2(e.dot + pos * e.dot_ctot + e.dot_coff + nt2 * BN + l * 64);
          }
        }
      }
      // read-back rows of this lane: tile row q*32 + it*8 + lane/4 -> w = lane/4, h = q*4 + it
      const int wx = tc.w0 + rsub;
      const int hy0 = tc.h0 + q * 4;
      const long pos0 = (wx < pw) ? (((long)tc.n * p.pd + tc.d0) * ph + hy0) * pw + wx : -1;
      const long off0 = pos0 * out_ctot + out_coff;
      const long moff0 = pos0 * e.mask_ctot + e.mask_coff, doff0 = pos0 * e.dot_ctot + e.dot_coff;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)(R * BN);
      auto issue = [&](uint32_t (&v)[32], int item) {
        const int ci = item >> rsh, r = item & (R - 1);
        const uint32_t ta = tbase + (uint32_t)(r * BN + ci * 32);
        tmem_ld16(ta, v);
        if (BN - ci * 32 > 16) tmem_ld16(ta + 16, v + 16);
      };
      auto place = [&](int r, EpiRows& rw) {
        const int hleft = ph - (hy0 + r * 16);              // rows it < hleft exist
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const bool ok = pos0 >= 0 && it < hleft;
          rw.ooff[it] = ok ? off0 + (long)(r * 16 + it) * row_elems : -1;
          if (MASK) rw.moff[it] = ok ? moff0 + (long)(r * 16 + it) * mrow_elems : -1;
          if (STATS == EPI_STATS_DOT) rw.doff[it] = ok ? doff0 + (long)(r * 16 + it) * drow_elems : -1;
        }
      };
      auto prefetch = [&](EpiAux<MASK, STATS>& aux, int item) {
        if (!kAux) return;
        const int ci = item >> rsh, r = item & (R - 1);
        EpiRows rw;
        place(r, rw);
        epi_prefetch<MASK, STATS>(e, nt * BN + ci * 32, BN - ci * 32 > 16 ? 32 : 16, lane, rw, aux);
      };
      auto process = [&](const uint32_t (&v)[32], const EpiAux<MASK, STATS>& aux, int item, int next) {
        const int ci = item >> rsh, r = item & (R - 1);
        const int c0 = ci * 32;
        const int ncol = BN - c0 > 16 ? 32 : 16;
        EpiRows rw;
        place(r, rw);
        epi_chunk<MASK, STATS>(e, v, ncol, nt * BN + c0, stg, lane, rw, sacc, aux);
        // this warp's partial sums of a chunk are complete when its next item belongs to another chunk
        if (kStats && (next >= items || (next >> rsh) != ci)) epi_fold_acc<STATS>(sacc, colsum, c0, ncol, BN, lane);
      };
      uint32_t va[32], vb[32];
      EpiAux<MASK, STATS> xa, xb;
      if (e.dbg != 2 && grp < items) {
        issue(va, grp);
        prefetch(xa, grp);
        for (int item = grp; item < items; item += 2 * ngrp) {
          const int i1 = item + ngrp, i2 = i1 + ngrp;
          tmem_ld_wait();
          if (i1 < items) { issue(vb, i1); prefetch(xb, i1); }
          process(va, xa, item, i1);
          if (i1 < items) {
            tmem_ld_wait();
            if (i2 < items) { issue(va, i2); prefetch(xa, i2); }
            process(vb, xb, i1, i2);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (kStats && cur_n >= 0) epi_flush_colsum(e, colsum, BN, cur_n, cur_nt * BN, tid, nthr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ===================================================================== depth-rolling forward / dgrad (N = 3 C_out)
// ncu on the halo kernel above for a C_out = 64 layer (192 -> 64 @ 128^3, profiles/r02_fwd_192_64_ncu.txt): the
// tensor pipe is 73 % busy while l1tex__data_pipe_tc_wavefronts_mem_shared runs at 90 % of its peak.  An M = 128,
// N = 64, K = 16 MMA has 32 cycles of math but fetches 4 KB of A + 2 KB of B from shared memory — 48 wavefronts of
// 128 B at one per cycle — so every N = 64 MMA is operand-fetch bound at 2/3 of the pipe, whatever the schedule.
// The fetch per FLOP only falls with a wider N, and a 3-D convolution has one to offer: the input plane z feeds
// THREE output planes, z+1, z and z-1, through the depth taps kd = 0, 1, 2.  This kernel walks a column of
// (16 R x 8) positions along the depth axis: every input halo plane is loaded once and multiplied, tap by tap,
// with the three depth slabs [kd=2 | kd=1 | kd=0] of the filter at once (N = 3 C_out = 192: 80 wavefronts for 96
// cycles of math, math-bound), accumulating into the TMEM accumulators of the three output planes it touches.
// The accumulators form a ring of S = 4 plane slots x R row-blocks (R S C_out = 512 columns = all of TMEM); a
// plane is complete — and handed to the epilogue warps — as soon as the input plane after it has been consumed,
// and its slot is recycled for the plane S further on.  A fresh accumulator is opened by splitting the first
// K step of an input plane into per-plane MMAs (accumulate = 0 on the new plane only); everything else is one
// MMA per (row-block, K step).  Activation traffic falls 3x (one halo plane per output plane instead of three).
struct RollArgs {
  int pn, pd, ph, pw;
  int tiles_h, tiles_w, nseg, segL, total;
  int kchunks, kdim, BN;
  int R, S;              // row-blocks per tile, accumulator slots (R * S * BN <= 512)
  int halo_h;            // 16 R + 2
  int a_stage_bytes, a_stages, b_slots;
  int b_resident;        // all 27 x kchunks filter slabs stay in shared memory (loaded once per CTA)
  EpiArgs e;
};

struct RollItem {
  int n, z0, Lw, h0, w0;
};
__device__ __forceinline__ RollItem roll_decode(const RollArgs& p, int work) {
  RollItem it;
  const int tw = work % p.tiles_w; work /= p.tiles_w;
  const int th = work % p.tiles_h; work /= p.tiles_h;
  const int seg = work % p.nseg; work /= p.nseg;
  it.n = work;
  it.z0 = seg * p.segL;
  it.Lw = p.pd - it.z0 < p.segL ? p.pd - it.z0 : p.segL;
  it.h0 = th * 16 * p.R;
  it.w0 = tw * 8;
  return it;
}

template <bool MASK, int STATS>
__global__ void __launch_bounds__(384, 1)
    tc_gemm_roll_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const RollArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem);
  uint64_t* aempty = afull + 8;
  uint64_t* bfull = aempty + 8;
  uint64_t* bempty = bfull + 16;
  uint64_t* tfull = bempty + 16;
  uint64_t* tempty = tfull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 8);
  uint8_t* sA = smem + epi_ctrl_bytes(p.e.epi_warps);
  uint8_t* sB = sA + (size_t)p.a_stages * p.a_stage_bytes;
  const int BN = p.BN, R = p.R, S = p.S;
  const int tap_bytes = BN * 128;
  const int b_bytes = 3 * tap_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < p.b_slots; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < S; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 32 * p.e.epi_warps); }
    fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(R * S * BN)) tmem_cols <<= 1;
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- A (halo plane) producer
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t plane_bytes = (uint32_t)(kHaloW * p.halo_h * 128);
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const RollItem it = roll_decode(p, work);
      const int zb = it.z0 > 0 ? it.z0 - 1 : 0;
      const int ze = it.z0 + it.Lw < p.pd ? it.z0 + it.Lw : p.pd - 1;
      for (int z = zb; z <= ze; ++z)
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&aempty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&afull[stage], plane_bytes);
            tma_load_5d(sA + (size_t)stage * p.a_stage_bytes, &tmA, &afull[stage], kc * 64, it.w0 - 1, it.h0 - 1, z,
                        it.n);
          }
          __syncwarp();
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 3) {
    // ---------------------------------------------------------------- B producer: slabs [kd=2 | kd=1 | kd=0] of (kh, kw)
    int slot = 0;
    uint32_t phase = 0;
    if (p.b_resident) {
      // slot (kc, g) holds its three depth slabs for the whole kernel
      if ((int)blockIdx.x < p.total && elect_one()) {
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int g = 0; g < 9; ++g) {
            const int sl = kc * 9 + g;
            uint8_t* dst = sB + (size_t)sl * b_bytes;
            mbar_arrive_expect_tx(&bfull[sl], (uint32_t)b_bytes);
#pragma unroll
            for (int j = 0; j < 3; ++j)
              tma_load_3d(dst + j * tap_bytes, &tmB, &bfull[sl], kc * 64, 0, (2 - j) * 9 + g);
          }
      }
      __syncwarp();
    }
    for (int work = blockIdx.x; work < p.total && !p.b_resident; work += gridDim.x) {
      const RollItem it = roll_decode(p, work);
      const int zb = it.z0 > 0 ? it.z0 - 1 : 0;
      const int ze = it.z0 + it.Lw < p.pd ? it.z0 + it.Lw : p.pd - 1;
      for (int z = zb; z <= ze; ++z)
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int g = 0; g < 9; ++g) {
            mbar_wait(&bempty[slot], phase ^ 1);
            if (elect_one()) {
              uint8_t* dst = sB + (size_t)slot * b_bytes;
              mbar_arrive_expect_tx(&bfull[slot], (uint32_t)b_bytes);
#pragma unroll
              for (int j = 0; j < 3; ++j)   // position j holds depth tap kd = 2 - j  <->  output plane z - 1 + j
                tma_load_3d(dst + j * tap_bytes, &tmB, &bfull[slot], kc * 64, 0, (2 - j) * 9 + g);
            }
            __syncwarp();
            if (++slot == p.b_slots) { slot = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    const uint32_t idesc1 = make_idesc_bf16(128, BN, 0, 0), idesc2 = make_idesc_bf16(128, 2 * BN, 0, 0),
                   idesc3 = make_idesc_bf16(128, 3 * BN, 0, 0);
    const uint64_t adesc0 = make_smem_desc_sw128(0, 16, kHaloW * 128);
    const uint64_t bdesc0 = make_smem_desc_sw128(0, 16, 1024);
    int stage = 0, slot = 0;
    uint32_t aphase = 0, bphase = 0;
    const bool resident = p.b_resident != 0;
    bool b_loaded = false;  // resident filters: the full barriers are waited for once
    long gi_base = 0;      // output planes this CTA has opened so far: plane index -> accumulator slot and parity
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const RollItem it = roll_decode(p, work);
      const int zb = it.z0 > 0 ? it.z0 - 1 : 0;
      const int ze = it.z0 + it.Lw < p.pd ? it.z0 + it.Lw : p.pd - 1;
      const int p_last = it.z0 + it.Lw - 1;
      for (int z = zb; z <= ze; ++z) {
        const int p_lo = z - 1 > it.z0 ? z - 1 : it.z0;
        const int p_hi = z + 1 < p_last ? z + 1 : p_last;
        const int cnt = p_hi - p_lo + 1;                  // 1..3 output planes fed by this input plane
        const int bpos = p_lo - (z - 1);                  // first filter slab position used
        // planes that receive their first contribution now: z + 1, and plane 0 at z = 0
        for (int pp = p_lo; pp <= p_hi; ++pp) {
          const int zfirst = pp > 0 ? pp - 1 : 0;
          if (zfirst == z) {
            const long gi = gi_base + (pp - it.z0);
            mbar_wait(&tempty[gi % S], (uint32_t)(((gi / S) & 1) ^ 1));
          }
        }
        tc_fence_after();
        const long gi_lo = gi_base + (p_lo - it.z0);
        const int s_lo = (int)(gi_lo % S);
        const int cnt1 = s_lo + cnt <= S ? cnt : S - s_lo;      // planes before the ring wraps
        const uint32_t tc_a = (uint32_t)(s_lo * BN);
        const uint32_t id_a = cnt1 == 3 ? idesc3 : (cnt1 == 2 ? idesc2 : idesc1);
        const bool two = cnt1 < cnt;
        const uint32_t id_b = cnt - cnt1 == 2 ? idesc2 : idesc1;
        const uint64_t boff_b = (uint64_t)((cnt1 * tap_bytes) >> 4);
        bool first_iter = true;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          int nk = (p.kdim - kc * 64 + 15) >> 4;
          nk = nk > 4 ? 4 : nk;
          mbar_wait(&afull[stage], aphase);
          tc_fence_after();
          const uint64_t da0 = adesc0 + (uint64_t)((sA_addr + (uint32_t)stage * (uint32_t)p.a_stage_bytes) >> 4);
          // the in-plane taps, fully unrolled: every descriptor is base + compile-time constant and the R x nk MMAs
          // of a tap go out back to back (the run-time version of this loop spent longer issuing than the tensor
          // pipe needs to execute: 755 -> TF/s, profiles/r02_roll_tune.txt)
          auto issue_plane = [&](auto nk_tag) {
            constexpr int NKC = decltype(nk_tag)::value;
            const int nkk = NKC > 0 ? NKC : nk;
#pragma unroll
            for (int g = 0; g < 9; ++g) {
              const int khi = g / 3, kwi = g - khi * 3;
              const int bsl = resident ? kc * 9 + g : slot;
              if (!resident || !b_loaded) {
                mbar_wait(&bfull[bsl], resident ? 0u : bphase);
                tc_fence_after();
              }
              if (elect_one()) {
                const uint64_t db0 = bdesc0 + (uint64_t)((sB_addr + (uint32_t)bsl * (uint32_t)b_bytes +
                                                          (uint32_t)bpos * (uint32_t)tap_bytes) >> 4);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                  if (r >= R) break;
                  const uint64_t da = da0 + (uint64_t)((((r * 16 + khi) * kHaloW + kwi) * 128) >> 4);
                  const uint32_t tcol = tmem_base + (uint32_t)(r * S * BN);
                  if (first_iter) {
                    // open / continue every plane separately: only a fresh plane starts from zero
                    for (int j = 0; j < cnt; ++j) {
                      const int pp = p_lo + j;
                      const int zfirst = pp > 0 ? pp - 1 : 0;
                      umma_bf16(tcol + (uint32_t)(((s_lo + j) % S) * BN), da, db0 + (uint64_t)((j * tap_bytes) >> 4),
                                idesc1, zfirst == z ? 0u : 1u);
                    }
                  } else {
                    umma_bf16(tcol + tc_a, da, db0, id_a, 1u);
                    if (two) umma_bf16(tcol, da, db0 + boff_b, id_b, 1u);
                  }
#pragma unroll
                  for (int k = 1; k < 4; ++k) {
                    if (k >= nkk) break;
                    umma_bf16(tcol + tc_a, da + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), id_a, 1u);
                    if (two) umma_bf16(tcol, da + (uint64_t)(2 * k), db0 + boff_b + (uint64_t)(2 * k), id_b, 1u);
                  }
                }
                if (!resident) umma_commit(&bempty[slot]);
                if (g == 8) umma_commit(&aempty[stage]);
              }
              __syncwarp();
              first_iter = false;
              if (!resident && ++slot == p.b_slots) { slot = 0; bphase ^= 1; }
            }
          };
          if (nk == 4) issue_plane(std::integral_constant<int, 4>{});
          else issue_plane(std::integral_constant<int, 0>{});
          if (resident && kc == p.kchunks - 1) b_loaded = true;
          if (++stage == p.a_stages) { stage = 0; aphase ^= 1; }
        }
        // planes whose last contribution this was: z - 1, and the volume's last plane at z = pd - 1
        if (elect_one()) {
          if (z - 1 >= it.z0 && z - 1 <= p_last) umma_commit(&tfull[(gi_base + (z - 1 - it.z0)) % S]);
          if (z == p.pd - 1 && z <= p_last) umma_commit(&tfull[(gi_base + (z - it.z0)) % S]);
        }
        __syncwarp();
      }
      gi_base += it.Lw;
    }
  } else if (warp >= 4 && warp < 4 + p.e.epi_warps) {
    // ---------------------------------------------------------------- epilogue (see epi_chunk)
    const int ew = warp - 4, q = ew & 3, grp = ew >> 2, ngrp = p.e.epi_warps >> 2;
    const int tid = threadIdx.x - 128, nthr = 32 * p.e.epi_warps, rsub = lane >> 2;
    uint8_t* stg = smem + 1024 + ew * kEpiStageBytes;
    float* colsum = reinterpret_cast<float*>(smem + 1024 + p.e.epi_warps * kEpiStageBytes);
    const EpiArgs& e = p.e;
    constexpr bool kStats = STATS != EPI_STATS_NONE;
    constexpr bool kAux = MASK || STATS == EPI_STATS_DOT;
    const int pw = p.pw, ph = p.ph;
    const long out_ctot = e.out_ctot, out_coff = e.out_coff;
    const long row_elems = (long)pw * out_ctot;
    const long mrow_elems = (long)pw * e.mask_ctot, drow_elems = (long)pw * e.dot_ctot;
    if (kStats) {
      for (int i = tid; i < 2 * BN; i += nthr) colsum[i] = 0.f;
      epi_bar(nthr);
    }
    int cur_n = -1;
    float sacc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) sacc[k] = 0.f;
    const int nchunks = (BN + 31) / 32;
    const int items = nchunks * R;       // item = chunk * R + row-block
    long gi_base = 0;
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const RollItem it = roll_decode(p, work);
      if (kStats && it.n != cur_n) {
        if (cur_n >= 0) epi_flush_colsum(e, colsum, BN, cur_n, 0, tid, nthr);
        cur_n = it.n;
      }
      const int wx = it.w0 + rsub;
      const int hy0 = it.h0 + q * 4;
      for (int pl = 0; pl < it.Lw; ++pl) {
        const int zp = it.z0 + pl;
        const long gi = gi_base + pl;
        const int sl = (int)(gi % S);
        if (kAux && pl + 1 < it.Lw) {
          // pull the NEXT plane's mask / dot lines of this tile into L2 (see the halo kernel)
          const int nl = (BN * 2 + 127) >> 7;
          const int nlines = 16 * R * 8 * nl;
          const long plane2 = ((long)it.n * p.pd + zp + 1) * ph;
          for (int i = tid; i < nlines; i += nthr) {
            const int l = i % nl, pp = i / nl;
            const int w_ = it.w0 + (pp & 7), h_ = it.h0 + (pp >> 3);
            if (w_ < pw && h_ < ph) {
              const long pos = (plane2 + h_) * pw + w_;
              if (MASK) prefetch_l2(e.mask + pos * e.mask_ctot + e.mask_coff + l * 64);
              if (STATS == EPI_STATS_DOT) prefetch_l2(e.dot + pos * e.dot_ctot + e.dot_coff + l * 64);
            }
          }
        }
        const long pos0 = (wx < pw) ? (((long)it.n * p.pd + zp) * ph + hy0) * pw + wx : -1;
        const long off0 = pos0 * out_ctot + out_coff;
        const long moff0 = pos0 * e.mask_ctot + e.mask_coff, doff0 = pos0 * e.dot_ctot + e.dot_coff;
        mbar_wait(&tfull[sl], (uint32_t)((gi / S) & 1));
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sl * BN);
        auto issue = [&](uint32_t (&v)[32], int item) {
          const int ci = item / R, r = item - ci * R;
          const uint32_t ta = tbase + (uint32_t)(r * S * BN + ci * 32);
          tmem_ld16(ta, v);
          if (BN - ci * 32 > 16) tmem_ld16(ta + 16, v + 16);
        };
        auto place = [&](int r, EpiRows& rw) {
          const int hleft = ph - (hy0 + r * 16);              // rows it < hleft exist
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            const bool ok = pos0 >= 0 && i2 < hleft;
            rw.ooff[i2] = ok ? off0 + (long)(r * 16 + i2) * row_elems : -1;
            if (MASK) rw.moff[i2] = ok ? moff0 + (long)(r * 16 + i2) * mrow_elems : -1;
            if (STATS == EPI_STATS_DOT) rw.doff[i2] = ok ? doff0 + (long)(r * 16 + i2) * drow_elems : -1;
          }
        };
        auto prefetch = [&](EpiAux<MASK, STATS>& aux, int item) {
          if (!kAux) return;
          const int ci = item / R, r = item - ci * R;
          EpiRows rw;
          place(r, rw);
          epi_prefetch<MASK, STATS>(e, ci * 32, BN - ci * 32 > 16 ? 32 : 16, lane, rw, aux);
        };
        auto process = [&](const uint32_t (&v)[32], const EpiAux<MASK, STATS>& aux, int item, int next) {
          const int ci = item / R, r = item - ci * R;
          const int c0 = ci * 32;
          const int ncol = BN - c0 > 16 ? 32 : 16;
          EpiRows rw;
          place(r, rw);
          epi_chunk<MASK, STATS>(e, v, ncol, c0, stg, lane, rw, sacc, aux);
          if (kStats && (next >= items || next / R != ci)) epi_fold_acc<STATS>(sacc, colsum, c0, ncol, BN, lane);
        };
        uint32_t va[32], vb[32];
        EpiAux<MASK, STATS> xa, xb;
        if (e.dbg != 2 && grp < items) {
          issue(va, grp);
          prefetch(xa, grp);
          for (int item = grp; item < items; item += 2 * ngrp) {
            const int i1 = item + ngrp, i2 = i1 + ngrp;
            tmem_ld_wait();
            if (i1 < items) { issue(vb, i1); prefetch(xb, i1); }
            process(va, xa, item, i1);
            if (i1 < items) {
              tmem_ld_wait();
              if (i2 < items) { issue(va, i2); prefetch(xa, i2); }
              process(vb, xb, i1, i2);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty[sl]);
      }
      gi_base += it.Lw;
    }
    if (kStats && cur_n >= 0) epi_flush_colsum(e, colsum, BN, cur_n, 0, tid, nthr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

thread_local int g_epi_dbg = 0;
static EpiArgs make_epi(const TcGemmProblem& p) {
  EpiArgs e{};
  e.out = p.out; e.out_ctot = p.out_ctot; e.out_coff = p.out_coff;
  e.bias = p.bias; e.relu = p.relu;
  e.mask = p.mask; e.mask_ctot = p.mask_ctot; e.mask_coff = p.mask_coff;
  e.dot = p.dot; e.dot_ctot = p.dot_ctot; e.dot_coff = p.dot_coff;
  e.stats = p.stats; e.stats_cols = p.ncols;
  e.epi_warps = 4;
  e.dbg = g_epi_dbg;
  return e;
}

static int stats_kind_of(const TcGemmProblem& p) {
  if (!p.stats) return EPI_STATS_NONE;
  if (p.dot) return EPI_STATS_DOT;
  return p.stats_kind == EPI_STATS_SUM ? EPI_STATS_SUM : EPI_STATS_SUMSQ;
}
thread_local int g_epi_warps = 0;  // tuning knob: 0 = heuristic, 4 / 8 = forced
// Two epilogue warps per TMEM lane quarter when a tile's MMA work is too short to hide a 4-warp
// epilogue (~1000 cycles per 32x32 chunk per warp, latency-bound): small K = taps * C_in.
static int pick_epi_warps(long k_total, bool aux) {
  if (g_epi_warps == 4 || g_epi_warps == 8) return g_epi_warps;
  return k_total <= (aux ? 2400 : 600) ? 8 : 4;   // measured: profiles/r01_epilogue_probe.log
}

template <class Args, class F>
static int launch_variant(F pick, bool mask, int kind, int grid, int block, int smem, cudaStream_t stream,
                          const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& a) {
  void (*fn)(CUtensorMap, CUtensorMap, Args) = pick(mask, kind);
  if (!fn) return B200_ERR_UNSUPPORTED;
  cudaError_t ce = ensure_dynamic_smem((const void*)fn, kSmemBudget);
  if (ce != cudaSuccess) return (int)ce;
  fn<<<grid, block, smem, stream>>>(tmA, tmB, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
#define B200_PICK_VARIANT(KERNEL, ARGS)                                                     \
  [](bool mask, int kind) -> void (*)(CUtensorMap, CUtensorMap, ARGS) {                     \
    if (!mask) {                                                                            \
      switch (kind) {                                                                       \
        case EPI_STATS_NONE: return KERNEL<false, EPI_STATS_NONE>;                          \
        case EPI_STATS_SUM: return KERNEL<false, EPI_STATS_SUM>;                            \
        case EPI_STATS_SUMSQ: return KERNEL<false, EPI_STATS_SUMSQ>;                        \
        case EPI_STATS_DOT: return KERNEL<false, EPI_STATS_DOT>;                            \
      }                                                                                     \
    } else {                                                                                \
      switch (kind) {                                                                       \
        case EPI_STATS_NONE: return KERNEL<true, EPI_STATS_NONE>;                           \
        case EPI_STATS_SUM: return KERNEL<true, EPI_STATS_SUM>;                             \
        case EPI_STATS_SUMSQ: return KERNEL<true, EPI_STATS_SUMSQ>;                         \
        case EPI_STATS_DOT: return KERNEL<true, EPI_STATS_DOT>;                             \
      }                                                                                     \
    }                                                                                       \
    return nullptr;                                                                         \
  }

static int pick_bn(int ncols, int limit) {
  for (int bn = limit; bn >= 16; bn -= 16)
    if (ncols % bn == 0) return bn;
  return 0;
}

bool tc_gemm_supported(const TcGemmProblem& p) {
  if (p.kdim < 16 || p.kdim % 8) return false;
  if (p.a.ctot % 8 || p.a.coff % 8) return false;
  if (p.out_ctot % 8 || p.out_coff % 8) return false;
  if (p.mask && (p.mask_ctot % 8 || p.mask_coff % 8)) return false;
  if (p.ncols % 16) return false;
  if (p.out_mode == TC_OUT_K2S2_SCATTER && (p.cout_t % 16 || p.stats)) return false;
  if (p.stats && (p.table || p.out_mode != TC_OUT_CONV)) return false;
  if (p.dot && (p.dot_ctot % 8 || p.dot_coff % 8)) return false;
  if (p.tap_mode == TC_TAP_K2S2 && (p.pd != 1 || p.a.h != 2 * p.ph || p.a.w != 2 * p.pw))
    return false;
  return true;
}

static int pick_bn(int ncols, int limit);
thread_local int g_halo_mode = 1;  // 0 disables the halo-reuse kernel (A/B testing from the bench tools)

thread_local int g_halo_rmax = 4;  // tuning knob: cap on row-blocks per super-tile
static bool halo_eligible(const TcGemmProblem& p) {
  return g_halo_mode && !p.table && p.tap_mode == TC_TAP_SHIFT && p.out_mode == TC_OUT_CONV && p.kh == 3 &&
         p.kw == 3 && (p.kd == 3 || p.kd == 1) && p.ph >= 12 && p.pw >= 8;
}

thread_local int g_halo_resident = 1;  // tuning knob: weight-resident mode of the halo kernel
thread_local int g_halo_bn = 0;        // tuning knob: 0 = heuristic column tile, else an upper bound on BN
thread_local int g_gemm_resident = 1;  // tuning knob: weight-resident mode of the per-tap kernel

static int tc_gemm_halo(const TcGemmProblem& p, cudaStream_t stream) {
  HaloArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  a.BN = pick_bn(p.ncols, g_halo_bn >= 16 ? g_halo_bn : 256);
  a.n_tiles = p.ncols / a.BN;
  a.kd = p.kd;
  a.kdim = p.kdim; a.kchunks = (p.kdim + 63) / 64;
  a.e = make_epi(p);
  const int kind = stats_kind_of(p);
  a.e.epi_warps = pick_epi_warps((long)p.kd * 9 * p.kdim, p.mask || kind != EPI_STATS_NONE);
  const int kSmemOverhead = smem_overhead(a.e.epi_warps);
  // R row-blocks share every weight slab: as many as TMEM (2 R BN <= 512 columns), the plane height and
  // shared memory allow
  int R = 1;
  for (int r : {4, 2}) {
    if (r > g_halo_rmax || 2 * r * a.BN > 512) continue;
    if (p.ph < 16 * r - (r == 2 ? 8 : 16)) continue;              // mostly-empty super-tiles
    const int stage = (kHaloW * (16 * r + 2) * 128 + 1023) / 1024 * 1024;
    if (kSmemBudget - kSmemOverhead - 2 * stage < 2 * 3 * a.BN * 128) continue;
    R = r;
    break;
  }
  // Weight-resident mode: when every weight slab of the layer fits in shared memory next to two
  // activation stages (2D layers with C_in * C_out <= 128 * 64) it is loaded once per CTA instead of
  // once per super-tile — those layers were bound by re-streaming 74-147 KB of weights from L2 per
  // 4 us of MMA work.  Weight sharing across row-blocks then buys nothing, so R only has to fit.
  const int ntaps_all = p.kd * 9;
  const long w_bytes = (long)ntaps_all * a.kchunks * a.BN * 128;
  a.b_resident = 0;
  if (g_halo_resident && a.n_tiles == 1 && ntaps_all * a.kchunks / 3 <= 16) {
    for (int r : {4, 2, 1}) {
      if (r > g_halo_rmax || 2 * r * a.BN > 512) continue;
      if (r == 4 && p.ph < 48) continue;
      if (r == 2 && p.ph < 24) continue;
      const long stage = (kHaloW * (16 * r + 2) * 128 + 1023) / 1024 * 1024;
      if (kSmemBudget - kSmemOverhead - w_bytes >= 2 * stage * (p.kd == 3 ? 2 : 1)) {
        a.b_resident = 1;
        R = r;
        break;
      }
    }
  }
  a.R = R;
  a.halo_h = 16 * R + 2;
  a.tiles_h = (p.ph + 16 * R - 1) / (16 * R); a.tiles_w = (p.pw + 7) / 8;
  a.m_tiles = p.pn * p.pd * a.tiles_h * a.tiles_w;
  a.total = a.m_tiles * a.n_tiles;
  a.a_stage_bytes = (kHaloW * a.halo_h * 128 + 1023) / 1024 * 1024;
  a.a_stages = 2;
  int rem = kSmemBudget - kSmemOverhead - a.a_stages * a.a_stage_bytes;
  int b_bytes;
  if (a.b_resident) {
    a.G = 3;
    b_bytes = a.G * a.BN * 128;
    a.b_slots = ntaps_all * a.kchunks / 3;
    rem -= (int)w_bytes;
    while (a.a_stages < 6 && rem >= a.a_stage_bytes) {
      ++a.a_stages;
      rem -= a.a_stage_bytes;
    }
  } else {
    a.G = (2 * 3 * a.BN * 128 <= rem) ? 3 : 1;
    b_bytes = a.G * a.BN * 128;
    // spend what is left on more activation stages first (cheap, hides TMA latency), then weight slots
    while (a.a_stages < 6 && rem - a.a_stage_bytes >= 3 * b_bytes) {
      ++a.a_stages;
      rem -= a.a_stage_bytes;
    }
    a.b_slots = rem / b_bytes;
    if (a.b_slots > 12) a.b_slots = 12;
    if (a.b_slots < 2) return B200_ERR_UNSUPPORTED;
  }

  CUtensorMap tmA, tmB;
  {
    const ActView& v = p.a;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, (uint32_t)kHaloW, (uint32_t)a.halo_h, 1, 1};
    int rc = make_tmap_bf16(&tmA, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const int ntaps = p.kd * 9;
    uint64_t dims[3] = {(uint64_t)p.kdim, (uint64_t)p.ncols, (uint64_t)ntaps};
    uint64_t str[3] = {1, (uint64_t)p.kdim, (uint64_t)p.kdim * p.ncols};
    uint32_t box[3] = {64, (uint32_t)a.BN, (uint32_t)a.G};
    int rc = make_tmap_bf16(&tmB, p.wp, 3, dims, str, box);
    if (rc) return rc;
  }
  const int smem = a.a_stages * a.a_stage_bytes + a.b_slots * b_bytes + kSmemOverhead;
  const int grid = a.total < sm_count() ? a.total : sm_count();
  return launch_variant<HaloArgs>(B200_PICK_VARIANT(tc_gemm_halo_kernel, HaloArgs), p.mask != nullptr, kind, grid,
                                  128 + 32 * a.e.epi_warps, smem, stream, tmA, tmB, a);
}

// tuning knob: depth-rolling N = 3 C_out kernel for narrow 3-D layers: 0 = off (default: measured slower than the halo
// kernel on B200, see roll_eligible), 1 = where the filters fit in shared memory, 2 = every eligible layer
thread_local int g_roll_mode = 0;
thread_local int g_roll_seg = 0;    // tuning knob: 0 = heuristic segment length, else planes per segment

// Where the 27 filter slabs fit in shared memory next to two halo stages (K * C_out <= 64 * 32: the first two
// layers of a 3-D U-Net and their data gradients) the walk streams activations only and runs math-bound.  With
// streamed filters it is bound by the L2 -> SM path instead — R = 2 row-blocks is all TMEM allows next to four
// plane slots, so each 24 KB filter slot lasts 768 cycles (37 B/clk per SM; ncu: tensor pipe 56 % busy,
// profiles/r02_roll_192_64_ncu.txt) and the halo kernel's R = 4 schedule wins (knob 10 = 2 forces the walk).
static long roll_filter_bytes(const TcGemmProblem& p) {
  return 27L * ((p.kdim + 63) / 64) * p.ncols * 128;
}
static bool roll_eligible(const TcGemmProblem& p) {
  if (!g_roll_mode || !halo_eligible(p) || p.kd != 3 || p.ncols > 64 || p.ncols % 16 || p.pd < 4 || p.ph < 24)
    return false;
  if ((p.kdim + 63) / 64 * 9 > 16) return false;          // barrier slots of the resident mode
  const long a_stage = (kHaloW * (16 * 2 + 2) * 128 + 1023) / 1024 * 1024;
  const bool fits = roll_filter_bytes(p) + 2 * a_stage + smem_overhead(4) <= kSmemBudget;
  return fits || g_roll_mode == 2;
}

static int tc_gemm_roll(const TcGemmProblem& p, cudaStream_t stream) {
  RollArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  a.BN = p.ncols;
  a.kdim = p.kdim; a.kchunks = (p.kdim + 63) / 64;
  a.e = make_epi(p);
  const int kind = stats_kind_of(p);
  a.e.epi_warps = g_epi_warps == 8 ? 8 : 4;     // a plane's MMAs (>= 6900 cycles) hide a 4-warp epilogue
  const int kSmemOverhead = smem_overhead(a.e.epi_warps);
  a.R = 2; a.S = 4;
  a.halo_h = 16 * a.R + 2;
  a.tiles_h = (p.ph + 16 * a.R - 1) / (16 * a.R); a.tiles_w = (p.pw + 7) / 8;
  // segment length: whole waves of the persistent grid, few segments (each re-reads two boundary planes)
  const long cols = (long)p.pn * a.tiles_h * a.tiles_w;
  int bestL = p.pd;
  double best = -1.0;
  for (int L : {4, 6, 8, 12, 16, 24, 32, 48, 64, 128}) {
    if (L > p.pd && L != 4) continue;
    const int LL = L > p.pd ? p.pd : L;
    const long items = cols * ((p.pd + LL - 1) / LL);
    const long waves = (items + sm_count() - 1) / sm_count();
    const double eff = (double)items / (double)(waves * sm_count()) * ((double)LL / (LL + 0.7));
    if (eff > best) { best = eff; bestL = LL; }
  }
  a.segL = g_roll_seg > 0 ? (g_roll_seg < p.pd ? g_roll_seg : p.pd) : bestL;
  a.nseg = (p.pd + a.segL - 1) / a.segL;
  a.total = (int)(cols * a.nseg);
  a.a_stage_bytes = (kHaloW * a.halo_h * 128 + 1023) / 1024 * 1024;
  const int b_bytes = 3 * a.BN * 128;
  // a filter slot (three depth slabs of one in-plane tap, 24 KB) lasts R x 4 MMAs = 768 cycles: at ~32 B/clk
  // the weight stream needs >= 4 slots in flight to cover L2 latency, so the halo planes get two stages only
  a.a_stages = 2;
  int rem = kSmemBudget - kSmemOverhead - a.a_stages * a.a_stage_bytes;
  a.b_resident = roll_filter_bytes(p) <= rem && a.kchunks * 9 <= 16;
  if (a.b_resident) {
    a.b_slots = a.kchunks * 9;
    rem -= a.b_slots * b_bytes;
    while (a.a_stages < 4 && rem >= a.a_stage_bytes) { ++a.a_stages; rem -= a.a_stage_bytes; }
  } else {
    a.b_slots = rem / b_bytes;
    if (a.b_slots > 12) a.b_slots = 12;
    if (a.b_slots < 2) return B200_ERR_UNSUPPORTED;
  }

  CUtensorMap tmA, tmB;
  {
    const ActView& v = p.a;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, (uint32_t)kHaloW, (uint32_t)a.halo_h, 1, 1};
    int rc = make_tmap_bf16(&tmA, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.kdim, (uint64_t)p.ncols, 27};
    uint64_t str[3] = {1, (uint64_t)p.kdim, (uint64_t)p.kdim * p.ncols};
    uint32_t box[3] = {64, (uint32_t)a.BN, 1};
    int rc = make_tmap_bf16(&tmB, p.wp, 3, dims, str, box);
    if (rc) return rc;
  }
  const int smem = a.a_stages * a.a_stage_bytes + a.b_slots * b_bytes + kSmemOverhead;
  const int grid = a.total < sm_count() ? a.total : sm_count();
  return launch_variant<RollArgs>(B200_PICK_VARIANT(tc_gemm_roll_kernel, RollArgs), p.mask != nullptr, kind, grid,
                                  128 + 32 * a.e.epi_warps, smem, stream, tmA, tmB, a);
}

int tc_gemm_conv(const TcGemmProblem& p, cudaStream_t stream) {
  if (!tc_gemm_supported(p)) return B200_ERR_UNSUPPORTED;
  if (roll_eligible(p)) return tc_gemm_roll(p, stream);
  if (halo_eligible(p)) return tc_gemm_halo(p, stream);
  GemmArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  const TileGeom g = pick_tile(p.pn, p.pd, p.ph, p.pw, p.tap_mode == TC_TAP_K2S2 || p.pd == 1);
  a.tdl = g.tdl; a.thl = g.thl; a.twl = g.twl;
  a.tiles_d = g.tiles_d; a.tiles_h = g.tiles_h; a.tiles_w = g.tiles_w;
  a.m_tiles = g.m_tiles;
  a.BN = pick_bn(p.ncols, 256);  // scatter mode: a tile may span several (i,j) blocks of cout_t columns
  a.chunk = (p.out_mode == TC_OUT_K2S2_SCATTER && p.cout_t % 32) ? 16 : 32;
  a.n_tiles = p.ncols / a.BN;
  a.total = a.m_tiles * a.n_tiles;
  a.ntaps = p.kd * p.kh * p.kw; a.kh = p.kh; a.kw = p.kw;
  a.pad_d = p.kd / 2; a.pad_h = p.kh / 2; a.pad_w = p.kw / 2;
  int wslabs = a.ntaps;
  if (p.table) {
    a.use_table = 1;
    a.tt = *p.table;
    a.ntaps = p.table->ntaps;
    wslabs = 0;
    for (int t = 0; t < a.ntaps; ++t) wslabs = wslabs > a.tt.widx[t] + 1 ? wslabs : a.tt.widx[t] + 1;
  }
  a.tap_mode = p.tap_mode; a.out_mode = p.out_mode;
  a.kdim = p.kdim; a.kchunks = (p.kdim + 63) / 64;
  a.cout_t = p.cout_t;
  a.e = make_epi(p);
  const int kind = stats_kind_of(p);
  a.e.epi_warps = pick_epi_warps((long)a.ntaps * p.kdim, p.mask || kind != EPI_STATS_NONE);
  const int kSmemOverhead = smem_overhead(a.e.epi_warps);
  // Weight-resident mode: with few taps and a short K (ConvTranspose 2x2 forward / backward-data) the
  // weight slabs of one column tile fit in shared memory; each CTA then keeps ONE column tile and only
  // streams activations (per-tile L2 traffic drops from A + B to A).
  const long w_bytes = (long)a.ntaps * a.kchunks * a.BN * 128;
  int grid = a.total < sm_count() ? a.total : sm_count();
  a.b_resident = 0;
  if (g_gemm_resident && kSmemBudget - kSmemOverhead - w_bytes >= 4L * kABytes &&
      a.m_tiles >= 2 * (sm_count() / a.n_tiles) && a.n_tiles <= sm_count()) {
    a.b_resident = 1;
    grid = (sm_count() / a.n_tiles) * a.n_tiles;
  }
  const int stage_bytes = a.b_resident ? kABytes : kABytes + a.BN * 128;
  a.stages = (int)((kSmemBudget - kSmemOverhead - (a.b_resident ? w_bytes : 0)) / stage_bytes);
  if (a.stages > 12) a.stages = 12;
  if (a.stages < 2) return B200_ERR_UNSUPPORTED;

  CUtensorMap tmA, tmB;
  int rc = make_act_tmap(&tmA, p.a, p.tap_mode, g);
  if (rc) return rc;
  {
    uint64_t dims[3] = {(uint64_t)p.kdim, (uint64_t)p.ncols, (uint64_t)wslabs};
    uint64_t str[3] = {1, (uint64_t)p.kdim, (uint64_t)p.kdim * p.ncols};
    uint32_t box[3] = {64, (uint32_t)a.BN, 1};
    rc = make_tmap_bf16(&tmB, p.wp, 3, dims, str, box);
    if (rc) return rc;
  }
  const int smem = a.stages * stage_bytes + kSmemOverhead + (a.b_resident ? (int)w_bytes : 0);
  return launch_variant<GemmArgs>(B200_PICK_VARIANT(tc_gemm_kernel, GemmArgs), p.mask != nullptr, kind, grid,
                                  128 + 32 * a.e.epi_warps, smem, stream, tmA, tmB, a);
}

// ===================================================================== weight gradient
// split-K factor for the weight-gradient kernels: the grid is base_ctas * splits one-shot CTAs, so pick
// the split count whose total fills whole waves of the SMs (a 2.007-wave grid costs 3 waves)
static int pick_splits(long base_ctas, int m_tiles, int sms, double epi_tiles = 2.0) {
  long best = 1;
  double best_score = -1.0;
  const long max_splits = m_tiles < 512 ? m_tiles : 512;
  for (long sp = 1; sp <= max_splits; ++sp) {
    const long total = base_ctas * sp;
    const long waves = (total + sms - 1) / sms;
    if (waves > 4 && sp > 1) break;
    const double eff = (double)total / (double)(waves * sms);          // SM occupancy of the schedule
    const long tiles = (m_tiles + sp - 1) / sp;
    const double work_eff = (double)m_tiles / (double)(tiles * sp);     // padding of the last split
    const double amort = (double)tiles / ((double)tiles + epi_tiles);   // prologue/epilogue per CTA, in tile times
    const double score = eff * work_eff * amort * (waves >= 2 ? 1.0 : 0.97);
    if (score > best_score) { best_score = score; best = sp; }
  }
  return (int)best;
}

// Both operands are MN-major: a [128 positions x 64 channels] TMA box is read by the tensor core
// with the positions as the reduction (K) dimension, 16 positions per tcgen05.mma.
struct WgradArgs {
  int pn, pd, ph, pw;
  int tdl, thl, twl, tiles_d, tiles_h, tiles_w, m_tiles;
  int ntaps, kh, kw, pad_d, pad_h, pad_w, tap_mode;
  int M, mblk, nblk, T, tap_groups, splits, tiles_per_split;
  int cm, cn;  // channel counts of the M / N operands
  int NB;      // 64-channel boxes of the N operand per MMA (N = 64 NB <= 256; LBO = one box)
  int stages;
  float* dw;
  long s_m, s_n;
  int use_table;
  TapTable tt;
};

__global__ void __launch_bounds__(256, 1)
    tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN,
                    const WgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  uint64_t* tfull = empty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
  uint8_t* tiles = smem + 1024;
  const int mboxes = p.M / 64;
  const int NB = p.NB, ncols = 64 * NB;
  const int stage_bytes = (mboxes + p.T * NB) * kABytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmN);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.T * ncols)) tmem_cols <<= 1;
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work decode: blockIdx.x -> (split, tap group, n block, m block)
  int w = blockIdx.x;
  const int mb = w % p.mblk; w /= p.mblk;
  const int nb = w % p.nblk; w /= p.nblk;
  const int tg = w % p.tap_groups; w /= p.tap_groups;
  const int split = w;
  const int mt_begin = split * p.tiles_per_split;
  int mt_end = mt_begin + p.tiles_per_split;
  if (mt_end > p.m_tiles) mt_end = p.m_tiles;
  const int tap0 = tg * p.T;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      const TileCoord tc = decode_tile(mt, p.tiles_d, p.tiles_h, p.tiles_w, p.tdl, p.thl, p.twl);
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
      uint8_t* sM = tiles + (size_t)stage * stage_bytes;
      uint8_t* sN = sM + mboxes * kABytes;
      mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
      for (int b = 0; b < mboxes; ++b) {
        const int c0 = mb * p.M + b * 64;
        if (p.tap_mode == TC_TAP_SHIFT)
          tma_load_5d(sM + b * kABytes, &tmM, &full[stage], c0, tc.w0, tc.h0, tc.d0, tc.n);
        else  // convT 2x2: the M operand is x, a plain (N,1,H,W,C) tensor
          tma_load_5d(sM + b * kABytes, &tmM, &full[stage], c0, tc.w0, tc.h0, 0, tc.n);
      }
      for (int t = 0; t < p.T; ++t) {
        const int tap = tap0 + t;
        const int kwi = tap % p.kw, khi = (tap / p.kw) % p.kh, kdi = tap / (p.kw * p.kh);
        for (int b = 0; b < NB; ++b) {
          uint8_t* dst = sN + (t * NB + b) * kABytes;
          const int c0 = (nb * NB + b) * 64;
          if (p.use_table)
            tma_load_5d(dst, &tmN, &full[stage], c0, tc.w0 + p.tt.dw[tap], tc.h0 + p.tt.dh[tap],
                        tc.d0 + p.tt.dd[tap], tc.n + p.tt.dn[tap]);
          else if (p.tap_mode == TC_TAP_SHIFT)
            tma_load_5d(dst, &tmN, &full[stage], c0, tc.w0 + kwi - p.pad_w, tc.h0 + khi - p.pad_h,
                        tc.d0 + kdi - p.pad_d, tc.n);
          else
            tma_load_5d(dst, &tmN, &full[stage], c0, kwi, tc.w0, khi, tc.n * p.ph + tc.h0);
        }
      }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // The T taps' boxes of the tap-addressed operand sit 16 KB apart, i.e. they ARE the atoms of one wider
    // MN-major operand: when T * ncols <= 256 a single MMA per K step covers all taps (N = 256 for the four taps
    // of a ConvTranspose 2x2, 192 for a three-tap group) instead of T fetch-bound N = 64 MMAs; the accumulator
    // layout (tap t at column t * ncols) is unchanged.
    const bool merged = p.T * ncols <= 256 && p.T > 1;
    const uint32_t idesc = make_idesc_bf16(p.M, merged ? p.T * ncols : ncols, 1, 1);
    const int tsteps = merged ? 1 : p.T;
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t maddr = smem_u32(tiles + (size_t)stage * stage_bytes);
      const uint32_t naddr = maddr + mboxes * kABytes;
      if (elect_one()) {
        const uint64_t da0 = make_smem_desc_sw128(maddr, kABytes, 1024);
        for (int t = 0; t < tsteps; ++t) {
          const uint64_t db0 = make_smem_desc_sw128(naddr + t * NB * kABytes, kABytes, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 positions (= 2 KB of rows) per MMA
            umma_bf16(tmem_base + t * ncols, da0 + (uint64_t)(k * 128), db0 + (uint64_t)(k * 128), idesc,
                      !(first && k == 0));
        }
        umma_commit(&empty[stage]);
        if (mt == mt_end - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      first = false;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (mt_end <= mt_begin && elect_one()) umma_commit(&tfull[0]);
  } else if (warp >= 4) {
    const int q = warp - 4;
    int m_local;
    bool lane_ok = true;
    if (p.M == 128) {
      m_local = q * 32 + lane;
    } else {  // M = 64: row m lives in TMEM lane (m % 16) + 32 * (m / 16)
      m_local = q * 16 + (lane & 15);
      lane_ok = lane < 16;
    }
    const int m = mb * p.M + m_local;
    const bool m_ok = lane_ok && m < p.cm && mt_end > mt_begin;
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < p.T; ++t) {
      const int tap = p.use_table ? p.tt.widx[tap0 + t] : tap0 + t;
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + t * ncols + c0, v);
        tmem_ld_wait();
        if (m_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = nb * ncols + c0 + j;
            if (n < p.cn) atomicAdd(p.dw + (long)m * p.s_m + (long)n * p.s_n + tap, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ===================================================================== halo-reuse weight gradient
// dw[co][ci][tap] += sum_pos dy[pos][co] * x[pos + tap][ci] for the 9 taps of one depth plane, from ONE
// x halo box (18 x 10 positions x 64 channels) and ONE dy tile (16 x 8 positions x <=64 channels) per
// 128-position reduction block.  The shifted operand sits on the MMA's M side: M = 128 rows are TWO
// taps x 64 input channels — the descriptor's leading byte offset (LBO) is simply the byte distance
// between the two taps' first halo rows — so narrow layers still issue full-rate M = 128 MMAs.
// Tap pairs (0,1)(2,3)(4,5)(6,7)(7,8); tap 7's duplicate is dropped in the epilogue.
// MN-major operands with arbitrary row starts / SBO = 1280 B verified on B200
// (tools/desc_probe_mn.cu, profiles/r01_desc_probe_mnmajor.log).
struct WgradHaloArgs {
  int pn, pd, ph, pw;
  int tiles_h, tiles_w, m_tiles;
  int kd;                       // depth taps (1 or 3); blockIdx decodes (split, kd, co blk, ci blk)
  int ciblk, coblk, splits, tiles_per_split;
  int cin, cout, N;             // N = co tile (multiple of 16, <= 64; 128 in the wide variant)
  int stages;
  float* dw;
  long s_co, s_ci;
  float* dbias;                 // optional: dbias[co] += sum_pos dy[pos][co] (conv bias gradient), or NULL
};

// WIDE variant (C_out >= 128): an N = 64 MMA reads 4 KB of A + 2 KB of B from shared memory per 32
// tensor-core cycles — more than the 128 B/clk shared memory delivers — so the narrow kernel tops out at
// ~70 % of the pipe (and spends 1 slot in 10 on the duplicated tap 7).  With N = 128 the same 4 KB of A
// feeds twice the math.  TMEM holds 512 columns = FOUR tap pairs x 128 output channels, so the wide
// kernel takes the pairs (0,1)(2,3)(5,6)(7,8) and leaves the un-shifted centre tap 4 — a plain
// dy^T * x GEMM — to tc_wgrad_kernel (N = 256 column blocks).
template <bool WIDE>
__global__ void __launch_bounds__(256, 1)
    tc_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                         const WgradHaloArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  uint64_t* tfull = empty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
  uint8_t* tiles = smem + 1024;
  constexpr int x_bytes = 23552;                 // 180 rows x 128 B, padded to 1024
  constexpr int kYBoxes = WIDE ? 2 : 1;          // dy tile: 128 rows x 128 B per 64 output channels
  constexpr int stage_bytes = x_bytes + kYBoxes * kABytes;
  constexpr int kGroups = WIDE ? 4 : 5;
  constexpr int kGroupCols = WIDE ? 128 : 64;
  constexpr int kCoBlock = WIDE ? 128 : 64;

  int w = blockIdx.x;
  const int cib = w % p.ciblk; w /= p.ciblk;
  const int cob = w % p.coblk; w /= p.coblk;
  const int kdi = w % p.kd; w /= p.kd;
  const int split = w;
  const int mt_begin = split * p.tiles_per_split;
  int mt_end = mt_begin + p.tiles_per_split;
  if (mt_end > p.m_tiles) mt_end = p.m_tiles;
  // Bias gradient = column sums of dy: the dy tiles pass through this CTA's shared memory anyway, so
  // the (otherwise idle until the end) epilogue warps of ONE CTA per (co block, split) add them up —
  // a stage is then released by the MMA commit plus one arrival per summing warp.
  const bool do_bias = p.dbias != nullptr && cib == 0 && kdi == 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], do_bias ? 5 : 1); }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      const TileCoord tc = decode_tile(mt, p.pd, p.tiles_h, p.tiles_w, 0, 4, 3);
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sX = tiles + (size_t)stage * stage_bytes;
        uint8_t* sY = sX + x_bytes;
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(kHaloPlaneBytes + kYBoxes * kABytes));
        tma_load_5d(sX, &tmX, &full[stage], cib * 64, tc.w0 - 1, tc.h0 - 1, tc.d0 + kdi - (p.kd >> 1), tc.n);
#pragma unroll
        for (int b = 0; b < kYBoxes; ++b)
          tma_load_5d(sY + b * kABytes, &tmDY, &full[stage], cob * kCoBlock + b * 64, tc.w0, tc.h0, tc.d0, tc.n);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, WIDE ? 128 : p.N, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t xaddr = smem_u32(tiles + (size_t)stage * stage_bytes);
      const uint32_t yaddr = xaddr + x_bytes;
      if (elect_one()) {
      const uint64_t db0 = make_smem_desc_sw128(yaddr, kABytes, 1024);
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        const int ta = WIDE ? (g < 2 ? 2 * g : 2 * g + 1) : (g < 4 ? 2 * g : 7), tb = ta + 1;
        const int ra = (ta / 3) * kHaloW + (ta % 3), rb = (tb / 3) * kHaloW + (tb % 3);
        const uint64_t da0 = make_smem_desc_sw128(xaddr + (uint32_t)(ra * 128), (uint32_t)((rb - ra) * 128),
                                                  kHaloW * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)  // 16 positions = 2 W-rows of the tile per MMA
          umma_bf16(tmem_base + g * kGroupCols, da0 + (uint64_t)(j * ((2 * kHaloW * 128) >> 4)),
                    db0 + (uint64_t)(j * (2048 >> 4)), idesc, !(first && j == 0));
      }
      umma_commit(&empty[stage]);
      if (mt == mt_end - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      first = false;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (mt_end <= mt_begin && elect_one()) umma_commit(&tfull[0]);
  } else if (warp >= 4) {
    const int q = warp - 4;
    if (do_bias) {
      // warp q sums positions [32 q, 32 q + 32) of every dy tile; lane owns channels (2 lane, 2 lane + 1)
      // of each 64-channel box: one conflict-free 128-byte row read per position (128-byte swizzle)
      float bs[kYBoxes][2];
#pragma unroll
      for (int b = 0; b < kYBoxes; ++b) bs[b][0] = bs[b][1] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = mt_begin; mt < mt_end; ++mt) {
        mbar_wait(&full[stage], phase);
        const uint8_t* sY = tiles + (size_t)stage * stage_bytes + x_bytes;
#pragma unroll
        for (int b = 0; b < kYBoxes; ++b) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const int row = q * 32 + r;
            const uint32_t v = *reinterpret_cast<const uint32_t*>(
                sY + b * kABytes + row * 128 + ((((lane >> 2) ^ (row & 7))) << 4) + ((lane & 3) << 2));
            const float2 f = unpack_bf16x2(v);
            bs[b][0] += f.x;
            bs[b][1] += f.y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int b = 0; b < kYBoxes; ++b) {
        const int co = cob * kCoBlock + b * 64 + 2 * lane;
        if (co < p.cout) atomicAdd(p.dbias + co, bs[b][0]);
        if (co + 1 < p.cout) atomicAdd(p.dbias + co + 1, bs[b][1]);
      }
    }
    const int m = q * 32 + lane;           // 0..127: (tap half, ci)
    const int half = m >> 6;
    const int ci = cib * 64 + (m & 63);
    const int ncols = WIDE ? 128 : p.N;
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < kGroups; ++g) {
      const int ta = WIDE ? (g < 2 ? 2 * g : 2 * g + 1) : (g < 4 ? 2 * g : 7);
      const int tap = ta + half;
      const bool dup = (!WIDE && g == 4 && half == 0);  // tap 7 was already produced by group 3
      const bool ok = !dup && ci < p.cin && mt_end > mt_begin;
      const long base = (long)ci * p.s_ci + (kdi * 9 + tap);
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + g * kGroupCols + c0, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = cob * kCoBlock + c0 + j;
            if (co < p.cout) atomicAdd(p.dw + (long)co * p.s_co + base, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ===================================================================================== 3-D weight gradient, N = 192
// ncu on the tap-pair kernel above (192 -> 64 @ 128^3, profiles/r02_wgrad_192_64_ncu.txt): the tensor pipe is 57 %
// active while l1tex__data_pipe_tc_wavefronts_mem_shared sits at 84 % of its peak — an M = 128, N = 64, K = 16
// MMA fetches 4 KB of A and 2 KB of B from shared memory (48 x 128-byte wavefronts, one per cycle) for 32 cycles of
// math.  The operand fetch, not the math, bounds every N = 64 MMA, and with C_out = 64 the N side has nothing
// more to give ... in ONE depth plane.  In 3-D the same x halo plane z meets THREE dy planes:
//     dw[kd] += sum_q dy[plane z + 1 - kd][q] * x[plane z][q + (kh, kw)]          kd = 0, 1, 2
// so the N operand is widened to 192 = (dy planes z-1 | z | z+1) x 64 output channels — one TMA box of depth 3
// lands the three planes 16 KB apart, which is exactly the leading-dimension byte offset of a 3-atom MN-major
// operand — and one MMA (80 wavefronts for 96 cycles of math: math-bound) produces all three depth taps of a
// tap pair.  TMEM holds 512 columns = two N = 192 tap-pair groups + 128 columns, so the nine in-plane taps are
// split over two CTA types that share nothing but the loads' L2 lines:
//     type 0: pairs (0,1) (2,3) at N = 192, tap 8 against dy planes (z-1 | z) at N = 128       256 cycles / K step
//     type 1: pairs (4,5) (6,7) at N = 192, tap 8 against dy plane  z+1      at N = 64        240 cycles / K step
// against 3 x 5 x 48 = 720 for the tap-pair kernel run once per depth tap.  Tiles are indexed by the x plane z
// (dy planes outside the volume are zero-filled by TMA), so there is no boundary case.  The bias gradient sums
// the middle dy plane only.  The separate centre-tap GEMM of the wide variant is not needed for 3-D layers.
struct WgradD3Args {
  int pn, pd, ph, pw;
  int tiles_h, tiles_w, m_tiles;
  int ciblk, coblk, splits, tiles_per_split;
  int cin, cout;
  int stages;
  float* dw;
  long s_co, s_ci;
  float* dbias;
};

static const int kD3XBytes = 23552;                    // 18 x 10 halo rows x 128 B, padded to 1024
static const int kD3YBytes = 3 * kABytes;              // three dy planes of 128 positions x 64 channels
static const int kD3StageBytes = kD3XBytes + kD3YBytes;

__global__ void __launch_bounds__(256, 1)
    tc_wgrad_d3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const WgradD3Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  uint64_t* tfull = empty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
  uint8_t* tiles = smem + 1024;

  int w = blockIdx.x;
  const int cib = w % p.ciblk; w /= p.ciblk;
  const int cob = w % p.coblk; w /= p.coblk;
  const int type = w % 2; w /= 2;
  const int split = w;
  const int mt_begin = split * p.tiles_per_split;
  int mt_end = mt_begin + p.tiles_per_split;
  if (mt_end > p.m_tiles) mt_end = p.m_tiles;
  const bool do_bias = p.dbias != nullptr && cib == 0 && type == 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], do_bias ? 5 : 1); }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      const TileCoord tc = decode_tile(mt, p.pd, p.tiles_h, p.tiles_w, 0, 4, 3);   // tc.d0 = x plane z
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sX = tiles + (size_t)stage * kD3StageBytes;
        uint8_t* sY = sX + kD3XBytes;
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(kHaloPlaneBytes + kD3YBytes));
        tma_load_5d(sX, &tmX, &full[stage], cib * 64, tc.w0 - 1, tc.h0 - 1, tc.d0, tc.n);
        tma_load_5d(sY, &tmDY, &full[stage], cob * 64, tc.w0, tc.h0, tc.d0 - 1, tc.n);   // planes z-1, z, z+1
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t id192 = make_idesc_bf16(128, 192, 1, 1);
    const uint32_t id_g2 = make_idesc_bf16(128, type == 0 ? 128 : 64, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t xaddr = smem_u32(tiles + (size_t)stage * kD3StageBytes);
      const uint32_t yaddr = xaddr + kD3XBytes;
      if (elect_one()) {
        const uint64_t db0 = make_smem_desc_sw128(yaddr, kABytes, 1024);                       // planes z-1 | z | z+1
        const uint64_t db2 = make_smem_desc_sw128(yaddr + (type == 0 ? 0 : 2 * kABytes), kABytes, 1024);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          // halo rows of the group's two taps: (0,1) (2,3) | (4,5) (6,7) | tap 8 twice (second half dropped)
          const int ta = g < 2 ? (type * 4 + 2 * g) : 8, tb = g < 2 ? ta + 1 : 8;
          const int ra = (ta / 3) * kHaloW + (ta % 3), rb = (tb / 3) * kHaloW + (tb % 3);
          const uint64_t da0 = make_smem_desc_sw128(xaddr + (uint32_t)(ra * 128), (uint32_t)((rb - ra) * 128),
                                                    kHaloW * 128);
          const uint64_t dbg = g < 2 ? db0 : db2;
          const uint32_t idesc = g < 2 ? id192 : id_g2;
#pragma unroll
          for (int j = 0; j < 8; ++j)   // 16 positions = 2 W-rows of the tile per MMA
            umma_bf16(tmem_base + g * 192, da0 + (uint64_t)(j * ((2 * kHaloW * 128) >> 4)),
                      dbg + (uint64_t)(j * (2048 >> 4)), idesc, !(first && j == 0));
        }
        umma_commit(&empty[stage]);
        if (mt == mt_end - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      first = false;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (mt_end <= mt_begin && elect_one()) umma_commit(&tfull[0]);
  } else if (warp >= 4) {
    const int q = warp - 4;
    if (do_bias) {
      // warp q sums positions [32 q, 32 q + 32) of the MIDDLE dy plane of every tile (each position once)
      float bs0 = 0.f, bs1 = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = mt_begin; mt < mt_end; ++mt) {
        mbar_wait(&full[stage], phase);
        const uint8_t* sY = tiles + (size_t)stage * kD3StageBytes + kD3XBytes + kABytes;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const int row = q * 32 + r;
          const uint32_t v = *reinterpret_cast<const uint32_t*>(
              sY + row * 128 + ((((lane >> 2) ^ (row & 7))) << 4) + ((lane & 3) << 2));
          const float2 f = unpack_bf16x2(v);
          bs0 += f.x;
          bs1 += f.y;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      const int co = cob * 64 + 2 * lane;
      if (co < p.cout) atomicAdd(p.dbias + co, bs0);
      if (co + 1 < p.cout) atomicAdd(p.dbias + co + 1, bs1);
    }
    const int m = q * 32 + lane;           // 0..127: (tap half, ci)
    const int half = m >> 6;
    const int ci = cib * 64 + (m & 63);
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < 3; ++g) {
      const int tap = g < 2 ? (type * 4 + 2 * g + half) : 8;
      const bool ok = !(g == 2 && half == 1) && ci < p.cin && mt_end > mt_begin;
      const int nblk = g < 2 ? 3 : (type == 0 ? 2 : 1);         // 64-column blocks = dy planes of the group
      for (int jb = 0; jb < nblk; ++jb) {
        const int kd = g < 2 ? 2 - jb : (type == 0 ? 2 - jb : 0);   // dy plane z-1+jb  <->  depth tap 2-jb
        const long base = (long)ci * p.s_ci + (kd * 9 + tap);
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + g * 192 + jb * 64 + c0, v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = cob * 64 + c0 + j;
              if (co < p.cout) atomicAdd(p.dw + (long)co * p.s_co + base, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ===================================================================================== 2-D weight gradient, N = 192
// The same widening for 3 x 3 layers without a depth axis: the three dy views are ROW shifts.  Tiles are 16 x 8
// positions p of the x frame; A = x[p + (0, kw-1)] (tap pair (kw, kw+1) on the M side, as above), B = the dy box
// of rows h0-1 .. h0+16 read at row offsets j = 0, 1, 2 — atom j = dy[p + (j-1, 0)], 1 KB after atom j-1, so the
// three atoms overlap in shared memory and the leading-dimension byte offset is one tile row:
//     D[(kw, ci)][(j, co)] = sum_p x[p + (0, kw-1)][ci] dy[p + (j-1, 0)][co] = dw[co][ci][kh = 2-j][kw].
// x is zero outside the image, so summing p over the image covers every term; dy rows outside are zero-filled.
// Two groups: kw pair (0,1) and kw = 2 (second half unused): 2 x 96 cycles per K step for nine taps, math-bound,
// against 5 x 48 (tap-pair kernel, fetch-bound) or 4 x 64 + the separate centre-tap GEMM (wide variant).
static const int kR2XBytes = 20480;                    // 16 x 10 positions x 128 B
static const int kR2YBytes = 18432;                    // 18 x 8 positions x 128 B
static const int kR2StageBytes = kR2XBytes + kR2YBytes;

__global__ void __launch_bounds__(256, 1)
    tc_wgrad_r2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const WgradD3Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  uint64_t* tfull = empty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
  uint8_t* tiles = smem + 1024;

  int w = blockIdx.x;
  const int cib = w % p.ciblk; w /= p.ciblk;
  const int cob = w % p.coblk; w /= p.coblk;
  const int split = w;
  const int mt_begin = split * p.tiles_per_split;
  int mt_end = mt_begin + p.tiles_per_split;
  if (mt_end > p.m_tiles) mt_end = p.m_tiles;
  const bool do_bias = p.dbias != nullptr && cib == 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], do_bias ? 5 : 1); }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      const TileCoord tc = decode_tile(mt, p.pd, p.tiles_h, p.tiles_w, 0, 4, 3);
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sX = tiles + (size_t)stage * kR2StageBytes;
        uint8_t* sY = sX + kR2XBytes;
        mbar_arrive_expect_tx(&full[stage], (uint32_t)kR2StageBytes);
        tma_load_5d(sX, &tmX, &full[stage], cib * 64, tc.w0 - 1, tc.h0, tc.d0, tc.n);       // 10 x 16 positions
        tma_load_5d(sY, &tmDY, &full[stage], cob * 64, tc.w0, tc.h0 - 1, tc.d0, tc.n);      // 8 x 18 positions
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, 192, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t xaddr = smem_u32(tiles + (size_t)stage * kR2StageBytes);
      const uint32_t yaddr = xaddr + kR2XBytes;
      if (elect_one()) {
        const uint64_t db0 = make_smem_desc_sw128(yaddr, 1024, 1024);            // atoms = dy rows shifted by 0 / 1 / 2
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          // M halves: kw = 0 | 1, then kw = 2 | 2 (second half dropped)
          const uint64_t da0 = make_smem_desc_sw128(xaddr + (uint32_t)(g * 2 * 128), g == 0 ? 128u : 0u,
                                                    kHaloW * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j)   // 16 positions = 2 rows of the tile per MMA
            umma_bf16(tmem_base + g * 192, da0 + (uint64_t)(j * ((2 * kHaloW * 128) >> 4)),
                      db0 + (uint64_t)(j * (2048 >> 4)), idesc, !(first && j == 0));
        }
        umma_commit(&empty[stage]);
        if (mt == mt_end - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      first = false;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (mt_end <= mt_begin && elect_one()) umma_commit(&tfull[0]);
  } else if (warp >= 4) {
    const int q = warp - 4;
    if (do_bias) {
      // warp q sums positions [32 q, 32 q + 32) of the tile's own 16 rows (box rows 1..16): each position once
      float bs0 = 0.f, bs1 = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = mt_begin; mt < mt_end; ++mt) {
        mbar_wait(&full[stage], phase);
        const uint8_t* sY = tiles + (size_t)stage * kR2StageBytes + kR2XBytes + 1024;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const int row = q * 32 + r;
          const uint32_t v = *reinterpret_cast<const uint32_t*>(
              sY + row * 128 + ((((lane >> 2) ^ (row & 7))) << 4) + ((lane & 3) << 2));
          const float2 f = unpack_bf16x2(v);
          bs0 += f.x;
          bs1 += f.y;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      const int co = cob * 64 + 2 * lane;
      if (co < p.cout) atomicAdd(p.dbias + co, bs0);
      if (co + 1 < p.cout) atomicAdd(p.dbias + co + 1, bs1);
    }
    const int m = q * 32 + lane;           // 0..127: (tap half, ci)
    const int half = m >> 6;
    const int ci = cib * 64 + (m & 63);
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < 2; ++g) {
      const int kwi = g == 0 ? half : 2;
      const bool ok = !(g == 1 && half == 1) && ci < p.cin && mt_end > mt_begin;
      for (int jb = 0; jb < 3; ++jb) {
        const long base = (long)ci * p.s_ci + ((2 - jb) * 3 + kwi);     // dy row shift jb - 1  <->  kh = 2 - jb
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + g * 192 + jb * 64 + c0, v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = cob * 64 + c0 + j;
              if (co < p.cout) atomicAdd(p.dw + (long)co * p.s_co + base, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

thread_local int g_wgrad_d3 = 1;   // tuning knob: N = 192 three-plane weight gradient for 3-D layers
thread_local int g_wgrad_d3_epi = 80;   // tuning knob: epilogue cost (65 k fp32 reductions per CTA) in tile times

thread_local int g_wgrad_halo_mode = 1;

static bool wgrad_halo_eligible(const TcWgradProblem& p) {
  return g_wgrad_halo_mode && !p.table && p.tap_mode == TC_TAP_SHIFT && p.kh == 3 && p.kw == 3 &&
         (p.kd == 3 || p.kd == 1) && p.ph >= 12 && p.pw >= 8;
}

thread_local int g_wgrad_wide = 1;   // tuning knob: N = 128 four-pair variant + centre-tap GEMM for C_out >= 128
static int tc_wgrad_generic(const TcWgradProblem& p, cudaStream_t stream, bool wide_n);

static int tc_wgrad_halo(const TcWgradProblem& p, cudaStream_t stream) {
  // conv: mop = dy (c_out), nop = x (c_in); dw index = co * s_m + ci * s_n + tap
  WgradHaloArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  a.tiles_h = (p.ph + 15) / 16; a.tiles_w = (p.pw + 7) / 8;
  a.m_tiles = p.pn * p.pd * a.tiles_h * a.tiles_w;
  a.kd = p.kd;
  a.cin = p.nop.c; a.cout = p.mop.c;
  // wide pays when the centre-tap GEMM can use N >= 128 (C_in >= 128) and there are enough tiles to split
  // (measured per layer: profiles/r01_kernel_perf_wgrad_wide.log, r01_layers_cfg3p_wide_vs_narrow.txt)
  const bool wide = g_wgrad_wide && a.cout >= 128 && a.cout % 64 == 0 && a.cin >= 128 &&
                    (a.cin >= 256 ? a.m_tiles >= 128 : a.m_tiles >= 2048);
  const int co_block = wide ? 128 : 64;
  a.ciblk = (a.cin + 63) / 64; a.coblk = (a.cout + co_block - 1) / co_block;
  a.N = wide ? 128 : (a.cout >= 64 ? 64 : ((a.cout + 15) / 16) * 16);
  const int stage_bytes = 23552 + (wide ? 2 : 1) * kABytes;
  a.stages = (kSmemBudget - 2048) / stage_bytes;
  if (a.stages > 6) a.stages = 6;
  const long base_ctas = (long)a.ciblk * a.coblk * a.kd;
  const long want = pick_splits(base_ctas, a.m_tiles, sm_count());
  a.tiles_per_split = (int)((a.m_tiles + want - 1) / want);
  a.splits = (a.m_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.dw = p.dw; a.s_co = p.s_m; a.s_ci = p.s_n;
  a.dbias = p.dbias;

  CUtensorMap tmX, tmDY;
  {
    const ActView& v = p.nop;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, (uint32_t)kHaloW, (uint32_t)kHaloH, 1, 1};
    int rc = make_tmap_bf16(&tmX, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const ActView& v = p.mop;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, 8, 16, 1, 1};
    int rc = make_tmap_bf16(&tmDY, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  const int smem = a.stages * stage_bytes + 2048;
  {
    cudaError_t e = ensure_dynamic_smem((const void*)tc_wgrad_halo_kernel<false>, kSmemBudget);
    if (e == cudaSuccess) e = ensure_dynamic_smem((const void*)tc_wgrad_halo_kernel<true>, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
  }
  const long grid = base_ctas * a.splits;
  if (!wide) {
    tc_wgrad_halo_kernel<false><<<(unsigned)grid, 256, smem, stream>>>(tmX, tmDY, a);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  if (g_wgrad_wide != 3) {   // (2 / 3: timing probes — pairs only / centre taps only)
    tc_wgrad_halo_kernel<true><<<(unsigned)grid, 256, smem, stream>>>(tmX, tmDY, a);
    B200_LAUNCH_CHECK();
  }
  if (g_wgrad_wide == 2) return B200_OK;
  // the centre tap of every depth plane: un-shifted in-plane, so a plain per-tap reduction GEMM
  TapTable tt{};
  tt.ntaps = p.kd;
  for (int k = 0; k < p.kd; ++k) {
    tt.dd[k] = (signed char)(k - (p.kd >> 1));
    tt.dh[k] = 0; tt.dw[k] = 0; tt.dn[k] = 0;
    tt.widx[k] = (signed char)(k * 9 + 4);
  }
  TcWgradProblem c = p;
  c.table = &tt;
  return tc_wgrad_generic(c, stream, true);
}

static bool wgrad_d3_eligible(const TcWgradProblem& p) {
  return g_wgrad_d3 && g_wgrad_halo_mode && !p.table && p.tap_mode == TC_TAP_SHIFT && p.kd == 3 && p.kh == 3 &&
         p.kw == 3 && p.ph >= 12 && p.pw >= 8 && p.pd >= 2 && p.mop.c >= 48;
}

static bool wgrad_r2_eligible(const TcWgradProblem& p) {
  return g_wgrad_d3 && g_wgrad_halo_mode && !p.table && p.tap_mode == TC_TAP_SHIFT && p.kd == 1 && p.kh == 3 &&
         p.kw == 3 && p.ph >= 12 && p.pw >= 8 && p.mop.c >= 48;
}

static int tc_wgrad_r2(const TcWgradProblem& p, cudaStream_t stream) {
  WgradD3Args a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  a.tiles_h = (p.ph + 15) / 16; a.tiles_w = (p.pw + 7) / 8;
  a.m_tiles = p.pn * p.pd * a.tiles_h * a.tiles_w;
  a.cin = p.nop.c; a.cout = p.mop.c;
  a.ciblk = (a.cin + 63) / 64; a.coblk = (a.cout + 63) / 64;
  a.stages = (kSmemBudget - 2048) / kR2StageBytes;
  if (a.stages > 6) a.stages = 6;
  const long base_ctas = (long)a.ciblk * a.coblk;
  const long want = pick_splits(base_ctas, a.m_tiles, sm_count(), (double)g_wgrad_d3_epi * 0.75);
  a.tiles_per_split = (int)((a.m_tiles + want - 1) / want);
  a.splits = (a.m_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.dw = p.dw; a.s_co = p.s_m; a.s_ci = p.s_n;
  a.dbias = p.dbias;
  CUtensorMap tmX, tmDY;
  {
    const ActView& v = p.nop;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, (uint32_t)kHaloW, 16, 1, 1};
    int rc = make_tmap_bf16(&tmX, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const ActView& v = p.mop;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, 8, 18, 1, 1};
    int rc = make_tmap_bf16(&tmDY, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  const int smem = a.stages * kR2StageBytes + 2048;
  cudaError_t e = ensure_dynamic_smem((const void*)tc_wgrad_r2_kernel, kSmemBudget);
  if (e != cudaSuccess) return (int)e;
  const long grid = base_ctas * a.splits;
  tc_wgrad_r2_kernel<<<(unsigned)grid, 256, smem, stream>>>(tmX, tmDY, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static int tc_wgrad_d3(const TcWgradProblem& p, cudaStream_t stream) {
  // conv: mop = dy (c_out), nop = x (c_in); dw index = co * s_m + ci * s_n + tap
  WgradD3Args a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  a.tiles_h = (p.ph + 15) / 16; a.tiles_w = (p.pw + 7) / 8;
  a.m_tiles = p.pn * p.pd * a.tiles_h * a.tiles_w;
  a.cin = p.nop.c; a.cout = p.mop.c;
  a.ciblk = (a.cin + 63) / 64; a.coblk = (a.cout + 63) / 64;
  a.stages = (kSmemBudget - 2048) / kD3StageBytes;
  if (a.stages > 4) a.stages = 4;
  if (a.stages < 2) return B200_ERR_UNSUPPORTED;
  const long base_ctas = (long)a.ciblk * a.coblk * 2;
  const long want = pick_splits(base_ctas, a.m_tiles, sm_count(), (double)g_wgrad_d3_epi);
  a.tiles_per_split = (int)((a.m_tiles + want - 1) / want);
  a.splits = (a.m_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.dw = p.dw; a.s_co = p.s_m; a.s_ci = p.s_n;
  a.dbias = p.dbias;

  CUtensorMap tmX, tmDY;
  {
    const ActView& v = p.nop;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, (uint32_t)kHaloW, (uint32_t)kHaloH, 1, 1};
    int rc = make_tmap_bf16(&tmX, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const ActView& v = p.mop;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.d, (uint64_t)v.n};
    uint64_t str[5] = {1, (uint64_t)v.ctot, (uint64_t)v.w * v.ctot, (uint64_t)v.h * v.w * v.ctot,
                       (uint64_t)v.d * v.h * v.w * v.ctot};
    uint32_t box[5] = {64, 8, 16, 3, 1};      // three depth planes: 16 KB apart in shared memory
    int rc = make_tmap_bf16(&tmDY, v.ptr + v.coff, 5, dims, str, box);
    if (rc) return rc;
  }
  const int smem = a.stages * kD3StageBytes + 2048;
  cudaError_t e = ensure_dynamic_smem((const void*)tc_wgrad_d3_kernel, kSmemBudget);
  if (e != cudaSuccess) return (int)e;
  const long grid = base_ctas * a.splits;
  tc_wgrad_d3_kernel<<<(unsigned)grid, 256, smem, stream>>>(tmX, tmDY, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

bool tc_wgrad_supported(const TcWgradProblem& p) {
  if (p.mop.c < 16 || p.nop.c < 16 || p.mop.c % 8 || p.nop.c % 8) return false;
  if (p.mop.ctot % 8 || p.mop.coff % 8 || p.nop.ctot % 8 || p.nop.coff % 8) return false;
  if (p.tap_mode == TC_TAP_K2S2 && p.pd != 1) return false;
  return true;
}

bool tc_wgrad_fuses_bias(const TcWgradProblem& p) {
  return tc_wgrad_supported(p) && wgrad_halo_eligible(p);
}

int tc_wgrad(const TcWgradProblem& p, cudaStream_t stream) {
  if (!tc_wgrad_supported(p)) return B200_ERR_UNSUPPORTED;
  if (wgrad_d3_eligible(p)) return tc_wgrad_d3(p, stream);
  if (wgrad_r2_eligible(p)) return tc_wgrad_r2(p, stream);
  if (wgrad_halo_eligible(p)) return tc_wgrad_halo(p, stream);
  return tc_wgrad_generic(p, stream, false);
}

// wide_n: one tap per CTA and up to four 64-channel boxes of the tap-addressed operand per MMA
// (N = 256): the centre-tap GEMM of the wide halo variant has no halo reuse, so its arithmetic
// intensity has to come from the tile shape.
static int tc_wgrad_generic(const TcWgradProblem& p, cudaStream_t stream, bool wide_n) {
  WgradArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  const TileGeom g = pick_tile(p.pn, p.pd, p.ph, p.pw, p.tap_mode == TC_TAP_K2S2 || p.pd == 1);
  a.tdl = g.tdl; a.thl = g.thl; a.twl = g.twl;
  a.tiles_d = g.tiles_d; a.tiles_h = g.tiles_h; a.tiles_w = g.tiles_w; a.m_tiles = g.m_tiles;
  a.ntaps = p.kd * p.kh * p.kw; a.kh = p.kh; a.kw = p.kw;
  a.pad_d = p.kd / 2; a.pad_h = p.kh / 2; a.pad_w = p.kw / 2;
  if (p.table) {
    a.use_table = 1;
    a.tt = *p.table;
    a.ntaps = p.table->ntaps;
  }
  a.tap_mode = p.tap_mode;
  a.cm = p.mop.c; a.cn = p.nop.c;
  a.M = a.cm > 64 ? 128 : 64;
  a.mblk = (a.cm + a.M - 1) / a.M;
  a.nblk = (a.cn + 63) / 64;
  a.T = (a.ntaps % 3 == 0) ? 3 : (a.ntaps % 4 == 0 ? 4 : 1);
  a.NB = 1;
  if (wide_n) {
    a.T = 1;
    const int boxes = (a.cn + 63) / 64;
    a.NB = boxes % 4 == 0 ? 4 : (boxes % 3 == 0 ? 3 : (boxes % 2 == 0 ? 2 : 1));
  }
  a.nblk = (a.cn + 64 * a.NB - 1) / (64 * a.NB);
  a.tap_groups = a.ntaps / a.T;
  const int stage_bytes = (a.M / 64 + a.T * a.NB) * kABytes;
  a.stages = (kSmemBudget - 2048) / stage_bytes;
  if (a.stages > 8) a.stages = 8;
  if (a.stages < 2) return B200_ERR_UNSUPPORTED;
  const long base_ctas = (long)a.mblk * a.nblk * a.tap_groups;
  const long want = pick_splits(base_ctas, a.m_tiles, sm_count());
  a.tiles_per_split = (int)((a.m_tiles + want - 1) / want);
  a.splits = (a.m_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.dw = p.dw; a.s_m = p.s_m; a.s_n = p.s_n;

  CUtensorMap tmM, tmN;
  int rc;
  if (p.tap_mode == TC_TAP_SHIFT) {
    rc = make_act_tmap(&tmM, p.mop, TC_TAP_SHIFT, g);
    if (rc) return rc;
    rc = make_act_tmap(&tmN, p.nop, TC_TAP_SHIFT, g);
  } else {
    rc = make_act_tmap(&tmM, p.mop, TC_TAP_SHIFT, g);
    if (rc) return rc;
    rc = make_act_tmap(&tmN, p.nop, TC_TAP_K2S2, g);
  }
  if (rc) return rc;
  const int smem = a.stages * stage_bytes + 2048;
  {
    cudaError_t e = ensure_dynamic_smem((const void*)tc_wgrad_kernel, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
  }
  const long grid = base_ctas * a.splits;
  tc_wgrad_kernel<<<(unsigned)grid, 256, smem, stream>>>(tmM, tmN, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace b200
