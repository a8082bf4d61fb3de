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

// ===================================================================== forward / dgrad GEMM
struct GemmArgs {
  int pn, pd, ph, pw;
  int tdl, thl, twl, tiles_d, tiles_h, tiles_w;
  int m_tiles, n_tiles, total;
  int ntaps, kh, kw, pad_d, pad_h, pad_w;
  int tap_mode, out_mode;
  int kchunks, kdim, BN, stages;
  bf16* out;
  int out_ctot, out_coff;
  int cout_t;
  const float* bias;
  int relu;
  const bf16* mask;
  int mask_ctot, mask_coff;
  int use_table;
  TapTable tt;
};

__global__ void __launch_bounds__(256, 1)
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
  uint8_t* tiles = smem + 1024;
  const int stage_bytes = kABytes + p.BN * 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 128);
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
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const int mt = work / p.n_tiles, nt = work - mt * p.n_tiles;
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
            tma_load_3d(sB, &tmB, &full[stage], kc * 64, nt * p.BN, wtap);
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
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
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
          const uint64_t db = make_smem_desc_sw128(aaddr + kABytes, 16, 1024);
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
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp - 4;
    const int row = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const int mt = work / p.n_tiles, nt = work - mt * p.n_tiles;
      const TileCoord tc = decode_tile(mt, p.tiles_d, p.tiles_h, p.tiles_w, p.tdl, p.thl, p.twl);
      const int wx = tc.w0 + (row & ((1 << p.twl) - 1));
      const int hy = tc.h0 + ((row >> p.twl) & ((1 << p.thl) - 1));
      const int dz = tc.d0 + (row >> (p.twl + p.thl));
      const bool valid = (wx < p.pw) && (hy < p.ph) && (dz < p.pd);
      long opos;
      int ocol0 = nt * p.BN;
      if (p.out_mode == TC_OUT_CONV) {
        opos = (((long)tc.n * p.pd + dz) * p.ph + hy) * p.pw + wx;
      } else {
        const int ij = ocol0 / p.cout_t;
        ocol0 -= ij * p.cout_t;
        opos = ((long)tc.n * (2 * p.ph) + (2 * hy + (ij >> 1))) * (2 * p.pw) + (2 * wx + (ij & 1));
      }
      bf16* orow = p.out + opos * p.out_ctot + p.out_coff + ocol0;
      const bf16* mrow = p.mask ? p.mask + opos * p.mask_ctot + p.mask_coff + ocol0 : nullptr;
      const float* brow = p.bias ? p.bias + ocol0 : nullptr;

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.BN;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (brow) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + c0 + j));
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (mrow) {
            const uint4 m0 = *reinterpret_cast<const uint4*>(mrow + c0);
            const uint4 m1 = *reinterpret_cast<const uint4*>(mrow + c0 + 8);
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 mv = unpack_bf16x2(mw[j]);
              if (!(mv.x > 0.f)) f[2 * j] = 0.f;
              if (!(mv.y > 0.f)) f[2 * j + 1] = 0.f;
            }
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
          o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
          o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
          *reinterpret_cast<uint4*>(orow + c0) = o0;
          *reinterpret_cast<uint4*>(orow + c0 + 8) = o1;
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
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
  bf16* out;
  int out_ctot, out_coff;
  const float* bias;
  int relu;
  const bf16* mask;
  int mask_ctot, mask_coff;
};

// Super-tile = 1 x (16 R) x 8 output positions of one depth plane.  A stage = the halo of ONE input plane
// (10 x (16R+2) positions x 64 channels); a weight slot = G taps; each weight slab is applied to all R
// row-blocks (R accumulators in TMEM), which divides the weight traffic — the limiter of narrow-C_out
// layers once the activations are halo-shared — by R.
__global__ void __launch_bounds__(256, 1)
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
  uint8_t* sA = smem + 1024;
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
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
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
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
    int stage = 0, slot = 0;
    uint32_t aphase = 0, bphase = 0, acc = 0, acc_phase = 0;
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)(p.R * p.BN);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        int nk = (p.kdim - kc * 64 + 15) >> 4;
        nk = nk > 4 ? 4 : nk;
        for (int kdi = 0; kdi < p.kd; ++kdi) {
          mbar_wait(&afull[stage], aphase);
          tc_fence_after();
          // descriptors differ only in their 14-bit start-address field: build once, then add
          const uint64_t da0 = make_smem_desc_sw128(smem_u32(sA + (size_t)stage * p.a_stage_bytes), 16,
                                                    kHaloW * 128);
          for (int khi = 0; khi < 3; ++khi) {
            for (int kwi = 0; kwi < 3; ++kwi) {
              const int tap = (kdi * 3 + khi) * 3 + kwi;
              const int sub = p.G == 3 ? kwi : 0;
              if (sub == 0) {
                mbar_wait(&bfull[slot], bphase);
                tc_fence_after();
              }
              const bool last_of_slot = sub == p.G - 1;
              const bool last_of_plane = khi == 2 && kwi == 2;
              if (elect_one()) {
                const uint64_t db = make_smem_desc_sw128(
                    smem_u32(sB + (size_t)slot * b_bytes + (size_t)sub * tap_bytes), 16, 1024);
                for (int r = 0; r < p.R; ++r) {
                  const uint64_t da = da0 + (uint64_t)((((r * 16 + khi) * kHaloW + kwi) * 128) >> 4);
#pragma unroll 4
                  for (int k = 0; k < nk; ++k)
                    umma_bf16(tmem_d + r * p.BN, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                              (kc | tap | k) != 0);
                }
                if (last_of_slot) umma_commit(&bempty[slot]);
                if (last_of_plane) {
                  umma_commit(&aempty[stage]);
                  if (kc == p.kchunks - 1 && kdi == p.kd - 1) umma_commit(&tfull[acc]);
                }
              }
              __syncwarp();
              if (last_of_slot && ++slot == p.b_slots) { slot = 0; bphase ^= 1; }
            }
          }
          if (++stage == p.a_stages) { stage = 0; aphase ^= 1; }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp - 4;
    const int row = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int work = blockIdx.x; work < p.total; work += gridDim.x) {
      const int mt = work / p.n_tiles, nt = work - mt * p.n_tiles;
      const TileCoord tc = decode_tile(mt, p.pd, p.tiles_h, p.tiles_w, 0, thl, 3);
      const int wx = tc.w0 + (row & 7);
      const int ocol0 = nt * p.BN;
      const float* brow = p.bias ? p.bias + ocol0 : nullptr;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      for (int r = 0; r < p.R; ++r) {
        const int hy = tc.h0 + r * 16 + (row >> 3);
        const bool valid = (wx < p.pw) && (hy < p.ph);
        const long opos = (((long)tc.n * p.pd + tc.d0) * p.ph + hy) * p.pw + wx;
        bf16* orow = p.out + opos * p.out_ctot + p.out_coff + ocol0;
        const bf16* mrow = p.mask ? p.mask + opos * p.mask_ctot + p.mask_coff + ocol0 : nullptr;
        const uint32_t taddr =
            tmem_base + ((uint32_t)(q * 32) << 16) + (acc * (uint32_t)p.R + r) * (uint32_t)p.BN;
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (valid) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if (brow) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + c0 + j));
                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (mrow) {
              const uint4 m0 = *reinterpret_cast<const uint4*>(mrow + c0);
              const uint4 m1 = *reinterpret_cast<const uint4*>(mrow + c0 + 8);
              const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 mv = unpack_bf16x2(mw[j]);
                if (!(mv.x > 0.f)) f[2 * j] = 0.f;
                if (!(mv.y > 0.f)) f[2 * j + 1] = 0.f;
              }
            }
            uint4 o0, o1;
            o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
            o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
            o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
            *reinterpret_cast<uint4*>(orow + c0) = o0;
            *reinterpret_cast<uint4*>(orow + c0 + 8) = o1;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
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
  if (p.out_mode == TC_OUT_K2S2_SCATTER) {
    if (p.cout_t % 16 || pick_bn(p.cout_t, 256) == 0) return false;
  }
  if (p.tap_mode == TC_TAP_K2S2 && (p.pd != 1 || p.a.h != 2 * p.ph || p.a.w != 2 * p.pw))
    return false;
  return true;
}

static int pick_bn(int ncols, int limit);
static int sm_count();
int g_halo_mode = 1;  // 0 disables the halo-reuse kernel (A/B testing from the bench tools)

static bool halo_eligible(const TcGemmProblem& p) {
  return g_halo_mode && !p.table && p.tap_mode == TC_TAP_SHIFT && p.out_mode == TC_OUT_CONV && p.kh == 3 &&
         p.kw == 3 && (p.kd == 3 || p.kd == 1) && p.ph >= 12 && p.pw >= 8;
}

int g_halo_rmax = 4;  // tuning knob: cap on row-blocks per super-tile

static int tc_gemm_halo(const TcGemmProblem& p, cudaStream_t stream) {
  HaloArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  a.BN = pick_bn(p.ncols, 256);
  a.n_tiles = p.ncols / a.BN;
  a.kd = p.kd;
  a.kdim = p.kdim; a.kchunks = (p.kdim + 63) / 64;
  // R row-blocks share every weight slab: as many as TMEM (2 R BN <= 512 columns), the plane height and
  // shared memory allow
  int R = 1;
  for (int r : {4, 2}) {
    if (r > g_halo_rmax || 2 * r * a.BN > 512) continue;
    if (p.ph < 16 * r - (r == 2 ? 8 : 16)) continue;              // mostly-empty super-tiles
    const int stage = (kHaloW * (16 * r + 2) * 128 + 1023) / 1024 * 1024;
    if (kSmemBudget - 2048 - 2 * stage < 2 * 3 * a.BN * 128) continue;
    R = r;
    break;
  }
  a.R = R;
  a.halo_h = 16 * R + 2;
  a.tiles_h = (p.ph + 16 * R - 1) / (16 * R); a.tiles_w = (p.pw + 7) / 8;
  a.m_tiles = p.pn * p.pd * a.tiles_h * a.tiles_w;
  a.total = a.m_tiles * a.n_tiles;
  a.a_stage_bytes = (kHaloW * a.halo_h * 128 + 1023) / 1024 * 1024;
  a.a_stages = 2;
  int rem = kSmemBudget - 2048 - a.a_stages * a.a_stage_bytes;
  a.G = (2 * 3 * a.BN * 128 <= rem) ? 3 : 1;
  int b_bytes = a.G * a.BN * 128;
  // spend what is left on more activation stages first (cheap, hides TMA latency), then weight slots
  while (a.a_stages < 6 && rem - a.a_stage_bytes >= 3 * b_bytes) {
    ++a.a_stages;
    rem -= a.a_stage_bytes;
  }
  a.b_slots = rem / b_bytes;
  if (a.b_slots > 12) a.b_slots = 12;
  if (a.b_slots < 2) return B200_ERR_UNSUPPORTED;
  a.out = p.out; a.out_ctot = p.out_ctot; a.out_coff = p.out_coff;
  a.bias = p.bias; a.relu = p.relu;
  a.mask = p.mask; a.mask_ctot = p.mask_ctot; a.mask_coff = p.mask_coff;

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
  const int smem = a.a_stages * a.a_stage_bytes + a.b_slots * b_bytes + 2048;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_halo_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = a.total < sm_count() ? a.total : sm_count();
  tc_gemm_halo_kernel<<<grid, 256, smem, stream>>>(tmA, tmB, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int tc_gemm_conv(const TcGemmProblem& p, cudaStream_t stream) {
  if (!tc_gemm_supported(p)) return B200_ERR_UNSUPPORTED;
  if (halo_eligible(p)) return tc_gemm_halo(p, stream);
  GemmArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  const TileGeom g = pick_tile(p.pn, p.pd, p.ph, p.pw, p.tap_mode == TC_TAP_K2S2 || p.pd == 1);
  a.tdl = g.tdl; a.thl = g.thl; a.twl = g.twl;
  a.tiles_d = g.tiles_d; a.tiles_h = g.tiles_h; a.tiles_w = g.tiles_w;
  a.m_tiles = g.m_tiles;
  a.BN = p.out_mode == TC_OUT_K2S2_SCATTER ? pick_bn(p.cout_t, 256) : pick_bn(p.ncols, 256);
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
  const int stage_bytes = kABytes + a.BN * 128;
  a.stages = (kSmemBudget - 2048) / stage_bytes;
  if (a.stages > 12) a.stages = 12;
  if (a.stages < 2) return B200_ERR_UNSUPPORTED;
  a.out = p.out; a.out_ctot = p.out_ctot; a.out_coff = p.out_coff; a.cout_t = p.cout_t;
  a.bias = p.bias; a.relu = p.relu;
  a.mask = p.mask; a.mask_ctot = p.mask_ctot; a.mask_coff = p.mask_coff;

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
  const int smem = a.stages * stage_bytes + 2048;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int grid = a.total < sm_count() ? a.total : sm_count();
  tc_gemm_kernel<<<grid, 256, smem, stream>>>(tmA, tmB, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ===================================================================== weight gradient
// split-K factor for the weight-gradient kernels: the grid is base_ctas * splits one-shot CTAs, so pick
// the split count whose total fills whole waves of the SMs (a 2.007-wave grid costs 3 waves)
static int pick_splits(long base_ctas, int m_tiles, int sms) {
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
    const double amort = (double)tiles / (double)(tiles + 2);           // prologue/epilogue per CTA
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
  const int stage_bytes = (mboxes + p.T) * kABytes;

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
  while (tmem_cols < (uint32_t)(p.T * 64)) tmem_cols <<= 1;
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
        if (p.use_table)
          tma_load_5d(sN + t * kABytes, &tmN, &full[stage], nb * 64, tc.w0 + p.tt.dw[tap],
                      tc.h0 + p.tt.dh[tap], tc.d0 + p.tt.dd[tap], tc.n + p.tt.dn[tap]);
        else if (p.tap_mode == TC_TAP_SHIFT)
          tma_load_5d(sN + t * kABytes, &tmN, &full[stage], nb * 64, tc.w0 + kwi - p.pad_w,
                      tc.h0 + khi - p.pad_h, tc.d0 + kdi - p.pad_d, tc.n);
        else
          tma_load_5d(sN + t * kABytes, &tmN, &full[stage], nb * 64, kwi, tc.w0, khi,
                      tc.n * p.ph + tc.h0);
      }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(p.M, 64, 1, 1);
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
        for (int t = 0; t < p.T; ++t) {
          const uint64_t db0 = make_smem_desc_sw128(naddr + t * kABytes, kABytes, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 positions (= 2 KB of rows) per MMA
            umma_bf16(tmem_base + t * 64, da0 + (uint64_t)(k * 128), db0 + (uint64_t)(k * 128), idesc,
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
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + t * 64 + c0, v);
        tmem_ld_wait();
        if (m_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = nb * 64 + c0 + j;
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
  int cin, cout, N;             // N = co tile (multiple of 16, <= 64)
  int stages;
  float* dw;
  long s_co, s_ci;
};

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
  const int x_bytes = 23552;                 // 180 rows x 128 B, padded to 1024
  const int stage_bytes = x_bytes + kABytes;  // + dy tile (128 rows x 128 B)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int w = blockIdx.x;
  const int cib = w % p.ciblk; w /= p.ciblk;
  const int cob = w % p.coblk; w /= p.coblk;
  const int kdi = w % p.kd; w /= p.kd;
  const int split = w;
  const int mt_begin = split * p.tiles_per_split;
  int mt_end = mt_begin + p.tiles_per_split;
  if (mt_end > p.m_tiles) mt_end = p.m_tiles;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      const TileCoord tc = decode_tile(mt, p.pd, p.tiles_h, p.tiles_w, 0, 4, 3);
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sX = tiles + (size_t)stage * stage_bytes;
        uint8_t* sY = sX + x_bytes;
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(kHaloPlaneBytes + kABytes));
        tma_load_5d(sX, &tmX, &full[stage], cib * 64, tc.w0 - 1, tc.h0 - 1, tc.d0 + kdi - (p.kd >> 1), tc.n);
        tma_load_5d(sY, &tmDY, &full[stage], cob * 64, tc.w0, tc.h0, tc.d0, tc.n);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, p.N, 1, 1);
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
      for (int g = 0; g < 5; ++g) {
        const int ta = g < 4 ? 2 * g : 7, tb = ta + 1;
        const int ra = (ta / 3) * kHaloW + (ta % 3), rb = (tb / 3) * kHaloW + (tb % 3);
        const uint64_t da0 = make_smem_desc_sw128(xaddr + (uint32_t)(ra * 128), (uint32_t)((rb - ra) * 128),
                                                  kHaloW * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)  // 16 positions = 2 W-rows of the tile per MMA
          umma_bf16(tmem_base + g * 64, da0 + (uint64_t)(j * ((2 * kHaloW * 128) >> 4)),
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
    const int m = q * 32 + lane;           // 0..127: (tap half, ci)
    const int half = m >> 6;
    const int ci = cib * 64 + (m & 63);
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < 5; ++g) {
      const int ta = g < 4 ? 2 * g : 7;
      const int tap = ta + half;
      const bool dup = (g == 4 && half == 0);  // tap 7 was already produced by group 3
      const bool ok = !dup && ci < p.cin && mt_end > mt_begin;
      const long base = (long)ci * p.s_ci + (kdi * 9 + tap);
      for (int c0 = 0; c0 < p.N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + g * 64 + c0, v);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int g_wgrad_halo_mode = 1;

static bool wgrad_halo_eligible(const TcWgradProblem& p) {
  return g_wgrad_halo_mode && !p.table && p.tap_mode == TC_TAP_SHIFT && p.kh == 3 && p.kw == 3 &&
         (p.kd == 3 || p.kd == 1) && p.ph >= 12 && p.pw >= 8;
}

static int tc_wgrad_halo(const TcWgradProblem& p, cudaStream_t stream) {
  // conv: mop = dy (c_out), nop = x (c_in); dw index = co * s_m + ci * s_n + tap
  WgradHaloArgs a{};
  a.pn = p.pn; a.pd = p.pd; a.ph = p.ph; a.pw = p.pw;
  a.tiles_h = (p.ph + 15) / 16; a.tiles_w = (p.pw + 7) / 8;
  a.m_tiles = p.pn * p.pd * a.tiles_h * a.tiles_w;
  a.kd = p.kd;
  a.cin = p.nop.c; a.cout = p.mop.c;
  a.ciblk = (a.cin + 63) / 64; a.coblk = (a.cout + 63) / 64;
  a.N = a.cout >= 64 ? 64 : ((a.cout + 15) / 16) * 16;
  const int stage_bytes = 23552 + kABytes;
  a.stages = (kSmemBudget - 2048) / stage_bytes;
  if (a.stages > 6) a.stages = 6;
  const long base_ctas = (long)a.ciblk * a.coblk * a.kd;
  const long want = pick_splits(base_ctas, a.m_tiles, sm_count());
  a.tiles_per_split = (int)((a.m_tiles + want - 1) / want);
  a.splits = (a.m_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.dw = p.dw; a.s_co = p.s_m; a.s_ci = p.s_n;

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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_halo_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long grid = base_ctas * a.splits;
  tc_wgrad_halo_kernel<<<(unsigned)grid, 256, smem, stream>>>(tmX, tmDY, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

bool tc_wgrad_supported(const TcWgradProblem& p) {
  if (p.mop.c < 16 || p.nop.c < 16 || p.mop.c % 8 || p.nop.c % 8) return false;
  if (p.mop.ctot % 8 || p.mop.coff % 8 || p.nop.ctot % 8 || p.nop.coff % 8) return false;
  if (p.tap_mode == TC_TAP_K2S2 && p.pd != 1) return false;
  return true;
}

int tc_wgrad(const TcWgradProblem& p, cudaStream_t stream) {
  if (!tc_wgrad_supported(p)) return B200_ERR_UNSUPPORTED;
  if (wgrad_halo_eligible(p)) return tc_wgrad_halo(p, stream);
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
  a.tap_groups = a.ntaps / a.T;
  const int stage_bytes = (a.M / 64 + a.T) * kABytes;
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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long grid = base_ctas * a.splits;
  tc_wgrad_kernel<<<(unsigned)grid, 256, smem, stream>>>(tmM, tmN, a);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace b200
