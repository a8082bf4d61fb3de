// HBM-bound kernels of the U-Net path: max-pool, nearest up-sampling, concat copies, ReLU backward,
// bias gradients, GroupNorm, BatchNorm and the fp32 NCHW <-> bf16 channels-last boundary.
// All of them stream channels-last bf16 with 128-bit accesses (8 channels per thread) and reduce
// with warp shuffles + one atomic per block.
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {

template <int V>
struct Vec;
template <>
struct Vec<8> {
  static __device__ __forceinline__ void load(const bf16* p, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
                 d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  static __device__ __forceinline__ void store(bf16* p, const float* f) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct Vec<1> {
  static __device__ __forceinline__ void load(const bf16* p, float* f) { f[0] = __bfloat162float(*p); }
  static __device__ __forceinline__ void store(bf16* p, const float* f) { *p = __float2bfloat16(f[0]); }
};

struct View {
  bf16* p;
  int ctot, coff;
  __device__ __forceinline__ bf16* at(long pos, int c) const { return p + pos * ctot + coff + c; }
};
static inline View mk(const b200_view& v) { return View{(bf16*)v.ptr, v.ctot, v.coff}; }
static inline bool vec8(int c, const b200_view& a) {
  return c % 8 == 0 && a.ctot % 8 == 0 && a.coff % 8 == 0 && ((uintptr_t)a.ptr % 16) == 0;
}
// out = in * 2^k: ATen's float index map floor(dst * (in / out)) is then exactly dst >> k
static inline bool pow2_factor(int in, int out) {
  if (in <= 0 || out % in) return false;
  const int f = out / in;
  return (f & (f - 1)) == 0;
}
static inline unsigned blocks_for(long total, int threads) {
  return (unsigned)((total + threads - 1) / threads);
}

// ------------------------------------------------------------------------------------ max-pool
template <int V>
__global__ void maxpool_fwd_kernel(int n, int d, int h, int w, int c, int pd, View x, View y) {
  const int od = d / pd, oh = h / 2, ow = w / 2, cv = c / V;
  const long total = (long)n * od * oh * ow * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int xw = (int)(r % ow); r /= ow;
    const int xh = (int)(r % oh); r /= oh;
    const int xd = (int)(r % od); r /= od;
    const int b = (int)r;
    float m[V];
    bool first = true;
    for (int a = 0; a < pd; ++a)
      for (int e = 0; e < 2; ++e)
        for (int g = 0; g < 2; ++g) {
          float f[V];
          const long pos = (((long)b * d + xd * pd + a) * h + xh * 2 + e) * w + xw * 2 + g;
          Vec<V>::load(x.at(pos, ch), f);
#pragma unroll
          for (int j = 0; j < V; ++j) m[j] = first ? f[j] : fmaxf(m[j], f[j]);
          first = false;
        }
    const long opos = (((long)b * od + xd) * oh + xh) * ow + xw;
    Vec<V>::store(y.at(opos, ch), m);
  }
}

// one thread per pooling WINDOW (including the partial windows an odd edge leaves over): reads the
// window once, routes dy to the first maximum in (d,h,w) scan order and writes every dx of the window
template <int V>
__global__ void maxpool_bwd_kernel(int n, int d, int h, int w, int c, int pd, View x, View dy, View add, View dx,
                                   int mask_relu) {
  const int od = d / pd, oh = h / 2, ow = w / 2, cv = c / V;
  const int wd = (d + pd - 1) / pd, wh = (h + 1) / 2, ww = (w + 1) / 2;
  const long total = (long)n * wd * wh * ww * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int pw_ = (int)(r % ww); r /= ww;
    const int ph_ = (int)(r % wh); r /= wh;
    const int pdd = (int)(r % wd); r /= wd;
    const int b = (int)r;
    const bool pooled = pw_ < ow && ph_ < oh && pdd < od;
    float m[V];
    int arg[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { m[j] = 0.f; arg[j] = -1; }
    if (pooled) {
      int k = 0;
      for (int a = 0; a < pd; ++a)
        for (int e = 0; e < 2; ++e)
          for (int g = 0; g < 2; ++g, ++k) {
            float f[V];
            const long pos = (((long)b * d + pdd * pd + a) * h + ph_ * 2 + e) * w + pw_ * 2 + g;
            Vec<V>::load(x.at(pos, ch), f);
#pragma unroll
            for (int j = 0; j < V; ++j)
              if (k == 0 || f[j] > m[j]) { m[j] = f[j]; arg[j] = k; }
          }
    }
    float g_[V];
#pragma unroll
    for (int j = 0; j < V; ++j) g_[j] = 0.f;
    if (pooled) Vec<V>::load(dy.at((((long)b * od + pdd) * oh + ph_) * ow + pw_, ch), g_);
    int k = 0;
    for (int a = 0; a < pd; ++a)
      for (int e = 0; e < 2; ++e)
        for (int g = 0; g < 2; ++g, ++k) {
          const int xd = pdd * pd + a, xh = ph_ * 2 + e, xw = pw_ * 2 + g;
          if (xd >= d || xh >= h || xw >= w) continue;
          const long pos = (((long)b * d + xd) * h + xh) * w + xw;
          float out[V];
#pragma unroll
          for (int j = 0; j < V; ++j) out[j] = 0.f;
          if (add.p) Vec<V>::load(add.at(pos, ch), out);
#pragma unroll
          for (int j = 0; j < V; ++j) out[j] += (arg[j] == k) ? g_[j] : 0.f;
          if (mask_relu) {  // x is a ReLU output: fold the ReLU backward of its producer in here
            float xk[V];
            Vec<V>::load(x.at(pos, ch), xk);
#pragma unroll
            for (int j = 0; j < V; ++j)
              if (!(xk[j] > 0.f)) out[j] = 0.f;
          }
          Vec<V>::store(dx.at(pos, ch), out);
        }
  }
}

// Even sizes, 8 channels per thread: the whole window (x, the skip gradient and dy) is fetched by
// independent 16-byte loads BEFORE any use — up to 17 loads in flight per thread, which is what it takes
// to cover HBM latency at 2048 threads per SM (the generic kernel above keeps one load in flight).
template <int PD, bool ADD, bool MASK>
__global__ void __launch_bounds__(256)
    maxpool_bwd_fast_kernel(int n, int d, int h, int w, int c, View x, View dy, View add, View dx) {
  constexpr int K = PD * 4;
  const int od = d / PD, oh = h / 2, ow = w / 2, cv = c / 8;
  const long total = (long)n * od * oh * ow * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * 8; r /= cv;
    const int pw_ = (int)(r % ow); r /= ow;
    const int ph_ = (int)(r % oh); r /= oh;
    const int pdd = (int)(r % od); r /= od;
    const int b = (int)r;
    long pos[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int a = k >> 2, e = (k >> 1) & 1, g = k & 1;
      pos[k] = (((long)b * d + pdd * PD + a) * h + ph_ * 2 + e) * w + pw_ * 2 + g;
    }
    uint4 xv[K], av[ADD ? K : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) xv[k] = ldg_nc_v4(x.at(pos[k], ch));
    if (ADD) {
#pragma unroll
      for (int k = 0; k < K; ++k) av[k] = ldg_nc_v4(add.at(pos[k], ch));
    }
    const uint4 gv = ldg_nc_v4(dy.at((((long)b * od + pdd) * oh + ph_) * ow + pw_, ch));
    uint4 ov[K];
#pragma unroll
    for (int q = 0; q < 4; ++q) {       // 32-bit word q = channels (2q, 2q + 1)
      const float2 gg = unpack_bf16x2((&gv.x)[q]);
      float2 m = unpack_bf16x2((&xv[0].x)[q]);
      int a0 = 0, a1 = 0;
#pragma unroll
      for (int k = 1; k < K; ++k) {
        const float2 f = unpack_bf16x2((&xv[k].x)[q]);
        if (f.x > m.x) { m.x = f.x; a0 = k; }
        if (f.y > m.y) { m.y = f.y; a1 = k; }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float2 o = make_float2(0.f, 0.f);
        if (ADD) o = unpack_bf16x2((&av[k].x)[q]);
        o.x += (a0 == k) ? gg.x : 0.f;
        o.y += (a1 == k) ? gg.y : 0.f;
        if (MASK) {
          const float2 f = unpack_bf16x2((&xv[k].x)[q]);
          if (!(f.x > 0.f)) o.x = 0.f;
          if (!(f.y > 0.f)) o.y = 0.f;
        }
        (&ov[k].x)[q] = pack_bf16x2(o.x, o.y);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) *reinterpret_cast<uint4*>(dx.at(pos[k], ch)) = ov[k];
  }
}

// ------------------------------------------------------------------------------------ UNet 3+ ops
// k x k / stride k max-pool on exact windows (nn.MaxPool2d(k, k, ceil_mode=True) of model/unet2d/unet.py:175-187;
// the network only accepts sizes divisible by 16, where ceil mode equals floor mode), 2D, 8 channels per thread.
__global__ void __launch_bounds__(256)
    maxpoolk_fwd_kernel(int n, int h, int w, int c, int k, View x, View y) {
  const int oh = h / k, ow = w / k, cv = c / 8;
  const long total = (long)n * oh * ow * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * 8; r /= cv;
    const int xw = (int)(r % ow); r /= ow;
    const int xh = (int)(r % oh); r /= oh;
    const long b = r;
    float m[8];
    for (int e = 0; e < k; ++e) {
#pragma unroll 4
      for (int g = 0; g < k; ++g) {
        float f[8];
        Vec<8>::load(x.at((b * h + xh * k + e) * w + xw * k + g, ch), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = (e | g) == 0 ? f[j] : fmaxf(m[j], f[j]);
      }
    }
    Vec<8>::store(y.at((b * oh + xh) * ow + xw, ch), m);
  }
}
// one thread per window: first maximum in (h, w) scan order receives dy; every dx of the window is written
// (+ add: a gradient of x's shape accumulated in, NULL = none)
__global__ void __launch_bounds__(256)
    maxpoolk_bwd_kernel(int n, int h, int w, int c, int k, View x, View dy, View add, View dx) {
  const int oh = h / k, ow = w / k, cv = c / 8;
  const long total = (long)n * oh * ow * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * 8; r /= cv;
    const int xw = (int)(r % ow); r /= ow;
    const int xh = (int)(r % oh); r /= oh;
    const long b = r;
    float m[8];
    int arg[8];
    for (int e = 0; e < k; ++e)
      for (int g = 0; g < k; ++g) {
        float f[8];
        Vec<8>::load(x.at((b * h + xh * k + e) * w + xw * k + g, ch), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if ((e | g) == 0 || f[j] > m[j]) { m[j] = f[j]; arg[j] = e * k + g; }
      }
    float gy[8];
    Vec<8>::load(dy.at((b * oh + xh) * ow + xw, ch), gy);
    for (int e = 0; e < k; ++e)
      for (int g = 0; g < k; ++g) {
        const long pos = (b * h + xh * k + e) * w + xw * k + g;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        if (add.p) Vec<8>::load(add.at(pos, ch), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += (arg[j] == e * k + g) ? gy[j] : 0.f;
        Vec<8>::store(dx.at(pos, ch), o);
      }
  }
}

// nn.Upsample(scale_factor = s, mode = 'bilinear') (align_corners = False), model/unet2d/unet.py:198-315:
// src = max((dst + 0.5) / s - 0.5, 0), i0 = floor(src), i1 = min(i0 + 1, in - 1), weights (1 - l, l).
__device__ __forceinline__ void bilinear_src(int dst, float inv_s, int in, int& i0, int& i1, float& l0, float& l1) {
  float src = inv_s * (dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i0 = i0 < in - 1 ? i0 : in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}
// Forward: one block per output row.  The row's two source rows and weights are found once; the threads walk
// the row's (pixel, 8-channel vector) elements with 32-bit index arithmetic (the first version spent its time
// in 64-bit div/mod: 1.5 TB/s of stores).  Four L2-resident loads, one 16-byte store per element.
__global__ void __launch_bounds__(256)
    bilinear2d_fwd_kernel(int n, int ih, int iw, int c, int s, View x, View y) {
  const int oh = ih * s, ow = iw * s;
  const unsigned cv = (unsigned)c / 8u;
  const float inv_s = 1.f / (float)s;
  const int row = blockIdx.x;
  const int b = row / oh, xh = row - b * oh;
  int h0, h1;
  float hl0, hl1;
  bilinear_src(xh, inv_s, ih, h0, h1, hl0, hl1);
  const long r0 = ((long)b * ih + h0) * iw, r1 = ((long)b * ih + h1) * iw, ro = (long)row * ow;
  const unsigned per_row = (unsigned)ow * cv;
#pragma unroll 2
  for (unsigned e = threadIdx.x; e < per_row; e += 256u) {
    const unsigned xw = e / cv;
    const int ch = (int)(e - xw * cv) * 8;
    int w0, w1;
    float wl0, wl1;
    bilinear_src((int)xw, inv_s, iw, w0, w1, wl0, wl1);
    float a00[8], a01[8], a10[8], a11[8], o[8];
    Vec<8>::load(x.at(r0 + w0, ch), a00);
    Vec<8>::load(x.at(r0 + w1, ch), a01);
    Vec<8>::load(x.at(r1 + w0, ch), a10);
    Vec<8>::load(x.at(r1 + w1, ch), a11);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = hl0 * (wl0 * a00[j] + wl1 * a01[j]) + hl1 * (wl0 * a10[j] + wl1 * a11[j]);
    Vec<8>::store(y.at(ro + xw, ch), o);
  }
}

// Forward for power-of-two scales (all the network uses): one thread per (source CELL, 8-channel vector).  Cell
// (cy, cx), cy in [-1, ih-1], holds the S x S outputs whose lower source index is (cy, cx) (border cells are
// half-size and collapse onto one source row / column): 4 corner loads feed S*S stores, the vertical lerp is done
// once per output row, column weights are compile-time constants.  ~30 instructions per 16-byte store instead of
// ~110, and 4/S^2 loads per output instead of 4.
template <int S>
__global__ void __launch_bounds__(256)
    bilinear2d_fwd_cell_kernel(int n, int ih, int iw, int c, View x, View y) {
  const int oh = ih * S, ow = iw * S, cv = c / 8;
  const long total = (long)n * (ih + 1) * (iw + 1) * cv;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long r = i;
  const int ch = (int)(r % cv) * 8; r /= cv;
  const int cx = (int)(r % (iw + 1)) - 1; r /= (iw + 1);
  const int cy = (int)(r % (ih + 1)) - 1;
  const long b = r / (ih + 1);
  const int h0 = cy < 0 ? 0 : cy, h1 = cy + 1 < ih ? cy + 1 : ih - 1;
  const int w0 = cx < 0 ? 0 : cx, w1 = cx + 1 < iw ? cx + 1 : iw - 1;
  const bool ey = h0 == h1, ex = w0 == w1;     // border cell (or a 1-pixel axis): a single source row / column
  float a00[8], a01[8], a10[8], a11[8];
  Vec<8>::load(x.at((b * ih + h0) * iw + w0, ch), a00);
  Vec<8>::load(x.at((b * ih + h0) * iw + w1, ch), a01);
  Vec<8>::load(x.at((b * ih + h1) * iw + w0, ch), a10);
  Vec<8>::load(x.at((b * ih + h1) * iw + w1, ch), a11);
  for (int ty = 0; ty < S; ++ty) {
    const int oy = S * cy + S / 2 + ty;
    if (oy < 0 || oy >= oh) continue;
    const float hl1 = ey ? 0.f : ((float)ty + 0.5f) * (1.f / S), hl0 = 1.f - hl1;
    float l[8], rr[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      l[q] = hl0 * a00[q] + hl1 * a10[q];
      rr[q] = hl0 * a01[q] + hl1 * a11[q];
    }
    const long orow = (b * oh + oy) * ow;
#pragma unroll
    for (int tx = 0; tx < S; ++tx) {
      const int ox = S * cx + S / 2 + tx;
      if (ox < 0 || ox >= ow) continue;
      const float wl1 = ex ? 0.f : ((float)tx + 0.5f) * (1.f / S), wl0 = 1.f - wl1;
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = wl0 * l[q] + wl1 * rr[q];
      Vec<8>::store(y.at(orow + ox, ch), o);
    }
  }
}

// Backward = the exact adjoint, separable: first along W (dy bf16 -> fp32 scratch [n*oh][iw][c]), then along H
// (scratch -> dx bf16, + add).  Each pass is ONE sequential walk along the axis per thread: source index i0 is
// non-decreasing in the destination index, so a two-entry register window (sums for out[j] and out[j+1]) slides
// along, every input element is read exactly once (plus a 2s-element lead-in per run), U loads are issued before
// they are consumed, no atomics, deterministic.  A thread owns outputs [j0, j1) of one (outer, 8-channel) line.
template <int U, class Load, class Emit>
__device__ __forceinline__ void bilinear_adjoint_walk(int s, int li, int j0, int j1, Load load, Emit emit) {
  const float inv_s = 1.f / (float)s;
  const int lo = li * s;
  int dlo = (j0 - 1) * s - 1, dhi = (j1 + 1) * s + 1;
  dlo = dlo < 0 ? 0 : dlo;
  dhi = dhi > lo ? lo : dhi;
  float a0[8], a1[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) a0[q] = a1[q] = 0.f;
  int j = -1000000;
  for (int d0 = dlo; d0 < dhi; d0 += U) {
    float g[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (d0 + u < dhi) load(d0 + u, g[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (d0 + u >= dhi) break;
      int i0, i1;
      float l0, l1;
      bilinear_src(d0 + u, inv_s, li, i0, i1, l0, l1);
      if (i0 != j) {                     // the window slides by one (or is placed for the first element)
        if (j >= j0 && j < j1) emit(j, a0);
#pragma unroll
        for (int q = 0; q < 8; ++q) { a0[q] = (i0 == j + 1) ? a1[q] : 0.f; a1[q] = 0.f; }
        j = i0;
      }
      const float w0 = (i1 != i0) ? l0 : l0 + l1, w1 = (i1 != i0) ? l1 : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) { a0[q] = fmaf(w0, g[u][q], a0[q]); a1[q] = fmaf(w1, g[u][q], a1[q]); }
    }
  }
  if (j >= j0 && j < j1) emit(j, a0);
  if (j + 1 >= j0 && j + 1 < j1) emit(j + 1, a1);
}
// Power-of-two scales (all the network uses): the same walk cell by cell.  Cell j (j = -1 .. li-1) holds the S
// destinations S*j + S/2 + t whose lower source index is j; slot t gives (1 - (t + .5)/S) to out[j] and (t + .5)/S to
// out[j+1] — compile-time weights, no per-element index search (the generic walk costs ~120 instructions per
// 16-byte load, this one ~30) — and the border cells (-1 and li-1, half-size) give everything to out[0] / out[li-1].
// out[j] = right part of cell j-1 + left part of cell j; 8 loads are issued before the first is consumed.
struct RawBf8 {
  uint4 v;
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float* f) const {
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
};
struct RawF8 {
  float4 a, b;
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float* f) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
template <int S, class Raw, class Load, class Emit>
__device__ __forceinline__ void bilinear_adjoint_cells(int li, int j0, int j1, Load load, Emit emit) {
  constexpr int NB = 8;                          // loads in flight per thread
  constexpr int G = S <= NB ? NB / S : 1;        // whole cells per batch (S <= 8)
  constexpr int PARTS = S <= NB ? 1 : S / NB;    // batches per cell (S = 16)
  const int lo = li * S;
  float prev[8], left[8], right[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) prev[q] = left[q] = right[q] = 0.f;
  for (int jb = j0 - 1; jb < j1; jb += G) {
#pragma unroll
    for (int part = 0; part < PARTS; ++part) {
      Raw g[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int j = jb + (S <= NB ? u / S : 0);
        const int t = S <= NB ? u % S : part * NB + u;
        const int d = S * j + S / 2 + t;
        if (j < j1 && d >= 0 && d < lo) load(d, g[u]); else g[u].zero();
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int j = jb + (S <= NB ? u / S : 0);
        const int t = S <= NB ? u % S : part * NB + u;
        const bool first = j < 0, last = j == li - 1;
        const float w1c = ((float)t + 0.5f) * (1.f / S);
        const float w1 = first ? 1.f : (last ? 0.f : w1c), w0 = 1.f - w1;
        float f[8];
        g[u].unpack(f);
#pragma unroll
        for (int q = 0; q < 8; ++q) { left[q] = fmaf(w0, f[q], left[q]); right[q] = fmaf(w1, f[q], right[q]); }
        if (t == S - 1) {                        // cell complete
          if (j >= j0 && j < j1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) left[q] += prev[q];
            emit(j, left);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) { prev[q] = right[q]; left[q] = right[q] = 0.f; }
        }
      }
    }
  }
}
template <int S>
__global__ void __launch_bounds__(256)
    bilinear2d_bwd_w_cell_kernel(long rows, int iw, int c, int runs, int per_run, View dy, float* __restrict__ tmp) {
  const int cv = c / 8;
  const long total = rows * runs * cv;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = (int)(i % cv) * 8;
  const long r = i / cv;
  const int run = (int)(r % runs);
  const long row = r / runs;
  const int j0 = run * per_run, j1 = min(iw, j0 + per_run);
  const bf16* ibase = dy.at(row * (long)iw * S, ch);
  const long istride = dy.ctot;
  float* obase = tmp + row * (long)iw * c + ch;
  bilinear_adjoint_cells<S, RawBf8>(
      iw, j0, j1, [&](int d, RawBf8& g) { g.v = ldg_nc_v4(ibase + d * istride); },
      [&](int j, const float* a) {
        float4* o = reinterpret_cast<float4*>(obase + (long)j * c);
        o[0] = make_float4(a[0], a[1], a[2], a[3]);
        o[1] = make_float4(a[4], a[5], a[6], a[7]);
      });
}
template <int S>
__global__ void __launch_bounds__(256)
    bilinear2d_bwd_h_cell_kernel(int n, int ih, int iw, int c, int runs, int per_run, const float* __restrict__ tmp,
                                 View add, View dx) {
  const int cv = c / 8;
  const long line = (long)iw * cv;
  const long total = (long)n * runs * line;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long q = i % line;
  const long r = i / line;
  const int run = (int)(r % runs);
  const int b = (int)(r / runs);
  const int xw = (int)(q / cv), ch = (int)(q % cv) * 8;
  const int j0 = run * per_run, j1 = min(ih, j0 + per_run);
  const float* ibase = tmp + ((long)b * ih * S * iw + xw) * c + ch;
  const long rowstride = (long)iw * c;
  bilinear_adjoint_cells<S, RawF8>(
      ih, j0, j1,
      [&](int d, RawF8& g) {
        const float4* p = reinterpret_cast<const float4*>(ibase + d * rowstride);
        g.a = p[0];
        g.b = p[1];
      },
      [&](int j, const float* a) {
        const long pos = ((long)b * ih + j) * iw + xw;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
        if (add.p) Vec<8>::load(add.at(pos, ch), o);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += a[k];
        Vec<8>::store(dx.at(pos, ch), o);
      });
}

// pass 1, along W: thread = (row of n*oh, run, channel vector)
__global__ void __launch_bounds__(256)
    bilinear2d_bwd_w_kernel(long rows, int iw, int c, int s, int runs, int per_run, View dy, float* __restrict__ tmp) {
  const int cv = c / 8;
  const long total = rows * runs * cv;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = (int)(i % cv) * 8;
  const long r = i / cv;
  const int run = (int)(r % runs);
  const long row = r / runs;
  const int j0 = run * per_run, j1 = min(iw, j0 + per_run);
  const long ibase = row * (long)iw * s;
  float* obase = tmp + row * (long)iw * c + ch;
  bilinear_adjoint_walk<8>(
      s, iw, j0, j1, [&](int d, float* g) { Vec<8>::load(dy.at(ibase + d, ch), g); },
      [&](int j, const float* a) {
        float4* o = reinterpret_cast<float4*>(obase + (long)j * c);
        o[0] = make_float4(a[0], a[1], a[2], a[3]);
        o[1] = make_float4(a[4], a[5], a[6], a[7]);
      });
}
// pass 2, along H: thread = (sample, run, (pixel of the row, channel vector))
__global__ void __launch_bounds__(256)
    bilinear2d_bwd_h_kernel(int n, int ih, int iw, int c, int s, int runs, int per_run,
                            const float* __restrict__ tmp, View add, View dx) {
  const int cv = c / 8;
  const long line = (long)iw * cv;
  const long total = (long)n * runs * line;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long q = i % line;
  const long r = i / line;
  const int run = (int)(r % runs);
  const int b = (int)(r / runs);
  const int xw = (int)(q / cv), ch = (int)(q % cv) * 8;
  const int j0 = run * per_run, j1 = min(ih, j0 + per_run);
  const float* ibase = tmp + ((long)b * ih * s * iw + xw) * c + ch;
  const long rowstride = (long)iw * c;
  bilinear_adjoint_walk<4>(
      s, ih, j0, j1,
      [&](int d, float* g) {
        const float4* p = reinterpret_cast<const float4*>(ibase + d * rowstride);
        const float4 u = p[0], v = p[1];
        g[0] = u.x; g[1] = u.y; g[2] = u.z; g[3] = u.w; g[4] = v.x; g[5] = v.y; g[6] = v.z; g[7] = v.w;
      },
      [&](int j, const float* a) {
        const long pos = ((long)b * ih + j) * iw + xw;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
        if (add.p) Vec<8>::load(add.at(pos, ch), o);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += a[k];
        Vec<8>::store(dx.at(pos, ch), o);
      });
}
// y = relu(a + b): residual join of ResNetBlock (model/unet3d/buildingblocks.py:319-323)
__global__ void __launch_bounds__(256) add_relu_kernel(long npos, int c, View a, View b, View y) {
  const int cv = c / 8;
  const long total = npos * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * 8;
    float u[8], v[8];
    Vec<8>::load(a.at(pos, ch), u);
    Vec<8>::load(b.at(pos, ch), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = fmaxf(u[j] + v[j], 0.f);
    Vec<8>::store(y.at(pos, ch), u);
  }
}
// dst += src (gradient accumulation where a tensor feeds several branches)
__global__ void __launch_bounds__(256) add_kernel(long npos, int c, View src, View dst) {
  const int cv = c / 8;
  const long total = npos * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * 8;
    float a[8], b[8];
    Vec<8>::load(dst.at(pos, ch), a);
    Vec<8>::load(src.at(pos, ch), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    Vec<8>::store(dst.at(pos, ch), a);
  }
}

// ------------------------------------------------------------------------------------ nearest
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) {
  const int s = (int)floorf(dst * scale);
  return s < in - 1 ? s : in - 1;
}
template <int V>
__global__ void upsample_fwd_kernel(int n, int id, int ih, int iw, int od, int oh, int ow, int c,
                                    View x, View y) {
  const int cv = c / V;
  const float sd = (float)id / od, sh = (float)ih / oh, sw = (float)iw / ow;
  const long total = (long)n * od * oh * ow * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int xw = (int)(r % ow); r /= ow;
    const int xh = (int)(r % oh); r /= oh;
    const int xd = (int)(r % od); r /= od;
    const int b = (int)r;
    const long ipos = (((long)b * id + nearest_src(xd, sd, id)) * ih + nearest_src(xh, sh, ih)) * iw +
                      nearest_src(xw, sw, iw);
    float f[V];
    Vec<V>::load(x.at(ipos, ch), f);
    Vec<V>::store(y.at((((long)b * od + xd) * oh + xh) * ow + xw, ch), f);
  }
}
// whole-factor up-sampling (the 2x of every U-Net decoder level): one thread per SOURCE vector writes
// its fd*fh*fw children, so the index arithmetic is paid once per 16 bytes read, not per 16 bytes written
template <int V>
__global__ void upsample_exact_fwd_kernel(int n, int id, int ih, int iw, int fd, int fh, int fw, int c,
                                          View x, View y) {
  const int cv = c / V;
  const int od = id * fd, oh = ih * fh, ow = iw * fw;
  const long total = (long)n * id * ih * iw * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int xw = (int)(r % iw); r /= iw;
    const int xh = (int)(r % ih); r /= ih;
    const int xd = (int)(r % id); r /= id;
    const int b = (int)r;
    float f[V];
    Vec<V>::load(x.at((((long)b * id + xd) * ih + xh) * iw + xw, ch), f);
    for (int a = 0; a < fd; ++a)
      for (int e = 0; e < fh; ++e) {
        const long row = (((long)b * od + xd * fd + a) * oh + xh * fh + e) * ow + (long)xw * fw;
        for (int g = 0; g < fw; ++g) Vec<V>::store(y.at(row + g, ch), f);
      }
  }
}
// children of source index s along one axis: [lo, hi)
__device__ __forceinline__ void children(int s, float scale, int in, int out, int& lo, int& hi) {
  int o = (int)floorf(s / scale) - 2;
  if (o < 0) o = 0;
  while (o < out && nearest_src(o, scale, in) < s) ++o;
  lo = o;
  while (o < out && nearest_src(o, scale, in) == s) ++o;
  hi = o;
}
template <int V>
__global__ void upsample_bwd_kernel(int n, int id, int ih, int iw, int od, int oh, int ow, int c,
                                    View dy, View dx, int exact) {
  const int cv = c / V;
  const float sd = (float)id / od, sh = (float)ih / oh, sw = (float)iw / ow;
  const long total = (long)n * id * ih * iw * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int xw = (int)(r % iw); r /= iw;
    const int xh = (int)(r % ih); r /= ih;
    const int xd = (int)(r % id); r /= id;
    const int b = (int)r;
    int d0, d1, h0, h1, w0, w1;
    if (exact) {   // whole factors: children of s are [s*f, s*f + f)
      d0 = xd * (od / id); d1 = d0 + od / id;
      h0 = xh * (oh / ih); h1 = h0 + oh / ih;
      w0 = xw * (ow / iw); w1 = w0 + ow / iw;
    } else {
      children(xd, sd, id, od, d0, d1);
      children(xh, sh, ih, oh, h0, h1);
      children(xw, sw, iw, ow, w0, w1);
    }
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int a = d0; a < d1; ++a)
      for (int e = h0; e < h1; ++e)
        for (int g = w0; g < w1; ++g) {
          float f[V];
          Vec<V>::load(dy.at((((long)b * od + a) * oh + e) * ow + g, ch), f);
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += f[j];
        }
    Vec<V>::store(dx.at((((long)b * id + xd) * ih + xh) * iw + xw, ch), acc);
  }
}

// ------------------------------------------------------------------------------------ parity planes
// The stride-2 transposed convolution is produced as 8 (3D) / 4 (2D) parity planes of the input's size:
// ct[z] with z = 2m + p lives in plane (pd,ph,pw) at position m.  These kernels fuse the gather from /
// scatter to that layout with the nearest resize to the skip's size (buildingblocks.py:703).
template <int V>
__global__ void planes_resize_fwd_kernel(int n, int d, int h, int w, int c, int kd, int od, int oh, int ow,
                                         View yp, View y) {
  const int cv = c / V;
  const int zd = kd == 1 ? 1 : 2 * d - 1, zh = 2 * h - 1, zw = 2 * w - 1;
  const float sd = (float)zd / od, sh = (float)zh / oh, sw = (float)zw / ow;
  const long total = (long)n * od * oh * ow * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int xw = (int)(r % ow); r /= ow;
    const int xh = (int)(r % oh); r /= oh;
    const int xd = (int)(r % od); r /= od;
    const int b = (int)r;
    const int a = nearest_src(xd, sd, zd), e = nearest_src(xh, sh, zh), g = nearest_src(xw, sw, zw);
    const int cls = kd == 1 ? ((e & 1) * 2 + (g & 1)) : (((a & 1) * 2 + (e & 1)) * 2 + (g & 1));
    const int md = kd == 1 ? 0 : a >> 1;
    const long ipos = ((((long)cls * n + b) * d + md) * h + (e >> 1)) * w + (g >> 1);
    float f[V];
    Vec<V>::load(yp.at(ipos, ch), f);
    Vec<V>::store(y.at((((long)b * od + xd) * oh + xh) * ow + xw, ch), f);
  }
}
template <int V>
__global__ void planes_resize_bwd_kernel(int n, int d, int h, int w, int c, int kd, int od, int oh, int ow,
                                         View dy, View dyp) {
  const int cv = c / V;
  const int classes = kd == 1 ? 4 : 8;
  const int zd = kd == 1 ? 1 : 2 * d - 1, zh = 2 * h - 1, zw = 2 * w - 1;
  const float sd = (float)zd / od, sh = (float)zh / oh, sw = (float)zw / ow;
  const long total = (long)classes * n * d * h * w * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int mw = (int)(r % w); r /= w;
    const int mh = (int)(r % h); r /= h;
    const int md = (int)(r % d); r /= d;
    const int b = (int)(r % n); r /= n;
    const int cls = (int)r;
    const int pd = kd == 1 ? 0 : (cls >> 2) & 1, ph = (cls >> 1) & 1, pw = cls & 1;
    const int a = kd == 1 ? 0 : 2 * md + pd, e = 2 * mh + ph, g = 2 * mw + pw;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    if (a < zd && e < zh && g < zw) {
      int d0, d1, h0, h1, w0, w1;
      children(a, sd, zd, od, d0, d1);
      children(e, sh, zh, oh, h0, h1);
      children(g, sw, zw, ow, w0, w1);
      for (int q = d0; q < d1; ++q)
        for (int s = h0; s < h1; ++s)
          for (int t = w0; t < w1; ++t) {
            float f[V];
            Vec<V>::load(dy.at((((long)b * od + q) * oh + s) * ow + t, ch), f);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] += f[j];
          }
    }
    Vec<V>::store(dyp.at(i / cv, ch), acc);
  }
}

// Exact 2x case (od, oh, ow) = 2 (d, h, w) of the 3D backward above: with src = max(dst - 1, 0) per axis a
// plane element has ONE child (dst = src + 1) unless src = 0 (children 0 and 1), so an interior thread
// moves the 8 parity classes of one (md, mh, mw) with 8 independent 16-byte loads in flight and no index
// search; only the three low faces take the summing path.
__global__ void __launch_bounds__(256)
    planes_resize_bwd2x_kernel(int n, int d, int h, int w, int c, View dy, View dyp) {
  const int cv = c / 8;
  const int od = 2 * d, oh = 2 * h, ow = 2 * w;
  const long total = (long)n * d * h * w * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * 8; r /= cv;
    const int mw = (int)(r % w); r /= w;
    const int mh = (int)(r % h); r /= h;
    const int md = (int)(r % d); r /= d;
    const long b = r;
    const long plane_pos = ((b * d + md) * h + mh) * w + mw;
    const long cls_stride = (long)n * d * h * w;
    if (md > 0 && mh > 0 && mw > 0) {
      uint4 v[8];
#pragma unroll
      for (int cls = 0; cls < 8; ++cls) {
        const int a = 2 * md + (cls >> 2), e = 2 * mh + ((cls >> 1) & 1), g = 2 * mw + (cls & 1);
        v[cls] = make_uint4(0, 0, 0, 0);
        if (a < od - 1 && e < oh - 1 && g < ow - 1)
          v[cls] = ldg_nc_v4(dy.at(((b * od + a + 1) * oh + e + 1) * ow + g + 1, ch));
      }
#pragma unroll
      for (int cls = 0; cls < 8; ++cls)
        *reinterpret_cast<uint4*>(dyp.at(cls * cls_stride + plane_pos, ch)) = v[cls];
    } else {
      for (int cls = 0; cls < 8; ++cls) {
        const int a = 2 * md + (cls >> 2), e = 2 * mh + ((cls >> 1) & 1), g = 2 * mw + (cls & 1);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (a < od - 1 && e < oh - 1 && g < ow - 1) {
          const int d0 = a ? a + 1 : 0, h0 = e ? e + 1 : 0, w0 = g ? g + 1 : 0;
          for (int q = d0; q <= a + 1; ++q)
            for (int t = h0; t <= e + 1; ++t)
              for (int u = w0; u <= g + 1; ++u) {
                float f[8];
                Vec<8>::load(dy.at(((b * od + q) * oh + t) * ow + u, ch), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
              }
        }
        Vec<8>::store(dyp.at(cls * cls_stride + plane_pos, ch), acc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ copies
template <int V>
__global__ void copy_kernel(long npos, int c, View src, View dst) {
  const int cv = c / V;
  const long total = npos * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    if (V == 8)
      *reinterpret_cast<uint4*>(dst.at(pos, ch)) = *reinterpret_cast<const uint4*>(src.at(pos, ch));
    else
      *dst.at(pos, ch) = *src.at(pos, ch);
  }
}

// ------------------------------------------------------------------------------------ column sums
// Generic "elementwise + per-channel reduction" skeleton. blockDim = 256; thread t owns channel
// vector (t % cvb) of the block's channel window and walks positions with stride rows = 256 / cvb.
template <int V, int NRED, bool DO_RED, class OutT, class F>
__device__ __forceinline__ void column_pass(long npos, int c, long chunk, F f, OutT* red_out[NRED]) {
  const int cv = c / V;
  const int cvb = cv < 256 ? cv : 256;  // channel vectors per block window
  const int rows = 256 / cvb;
  const int t_cv = threadIdx.x % cvb, t_row = threadIdx.x / cvb;
  const int windows = (cv + cvb - 1) / cvb;
  const int win = blockIdx.y % windows;
  const int my_cv = win * cvb + t_cv;
  float acc[NRED][V];
#pragma unroll
  for (int k = 0; k < NRED; ++k)
#pragma unroll
    for (int j = 0; j < V; ++j) acc[k][j] = 0.f;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < npos ? pbeg + chunk : npos;
  if (t_row < rows && my_cv < cv) {
#pragma unroll 2
    for (long pos = pbeg + t_row; pos < pend; pos += rows) f(pos, my_cv * V, acc);
  }
  if (!DO_RED) return;
  __shared__ float sred[256 * 8];
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < V; ++j) sred[threadIdx.x * V + j] = acc[k][j];
    __syncthreads();
    if (t_row == 0 && my_cv < cv) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += sred[(r * cvb + t_cv) * V + j];
        atomicAdd(red_out[k] + my_cv * V + j, (OutT)s);
      }
    }
  }
}
static inline void column_grid(long npos, int c, int V, dim3& grid, long& chunk) {
  const int cv = c / V;
  const int cvb = cv < 256 ? cv : 256;
  const int windows = (cv + cvb - 1) / cvb;
  long want = 148L * 8 / windows;
  if (want < 1) want = 1;
  chunk = (npos + want - 1) / want;
  const int rows = 256 / cvb;
  if (chunk < 4L * rows) chunk = 4L * rows;
  grid = dim3((unsigned)((npos + chunk - 1) / chunk), (unsigned)windows);
}

template <int V, bool MASK, bool BIAS>
__global__ void __launch_bounds__(256, 4)
    relu_bwd_kernel(long npos, int c, long chunk, View dy, View a, View dz, float* db) {
  float* outs[1] = {db};
  auto f = [&](long pos, int ch, float(*acc)[V]) {
    float g[V];
    Vec<V>::load(dy.at(pos, ch), g);
    if (MASK) {
      float av[V];
      Vec<V>::load(a.at(pos, ch), av);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(av[j] > 0.f)) g[j] = 0.f;
      Vec<V>::store(dz.at(pos, ch), g);
    }
    if (BIAS) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[0][j] += g[j];
    }
  };
  column_pass<V, 1, BIAS, float>(npos, c, chunk, f, outs);
}

// ------------------------------------------------------------------------------------ layout
__global__ void nchw_to_cl_kernel(int n, int c, long spatial, const float* src, View dst) {
  const long total = (long)n * spatial;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / spatial, s = i % spatial;
    for (int ch = 0; ch < c; ++ch)
      *dst.at(i, ch) = __float2bfloat16(src[(b * c + ch) * spatial + s]);
  }
}
__global__ void cl_to_nchw_kernel(int n, int c, long spatial, View src, float* dst) {
  const long total = (long)n * spatial;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / spatial, s = i % spatial;
    for (int ch = 0; ch < c; ++ch)
      dst[(b * c + ch) * spatial + s] = __bfloat162float(*src.at(i, ch));
  }
}

// ------------------------------------------------------------------------------------ narrow-input im2col
// The network's first convolution has 1-4 input channels.  As an implicit GEMM its weight gradient issues
// one K = 16 MMA per tap with 12-15 of the 16 K lanes zero (and M = 2 taps x 64 rows with 1-4 real ones);
// with the taps concatenated along K instead (a[pos][tap * c_in + ci], K' = taps * c_in rounded up to
// 16 <= 64) it is one short reduction GEMM that streams a and dy at HBM speed.
// One thread per position: the kd x 3 x 3 neighbourhood is gathered once (taps and channels are
// compile-time, so there is no per-element index arithmetic) and the K' row leaves as 16-byte stores.
// ONES: one extra column per tap that is 1 where the tap falls inside the volume (the "constant-one input
// channel" whose weight gradient gives the first norm's parameter gradients, engine._plant_ones) — synthesised,
// not read.
template <int KD, int CREAL, int ONES = 0>
__global__ void __launch_bounds__(256)
    im2col3_kernel(int n, int d, int h, int w, int kp, View x, bf16* __restrict__ a) {
  constexpr int CIN = CREAL + ONES;
  constexpr int NT = KD * 9, KR = NT * CIN, KV = (KR + 7) / 8;   // 16-byte vectors that hold real taps
  const int kv = kp / 8;
  const long total = (long)n * d * h * w;
  for (long pos = (long)blockIdx.x * blockDim.x + threadIdx.x; pos < total; pos += (long)gridDim.x * blockDim.x) {
    long r = pos;
    const int xw = (int)(r % w); r /= w;
    const int xh = (int)(r % h); r /= h;
    const int xd = (int)(r % d); r /= d;
    const long b = r;
    unsigned short e[KV * 8];
#pragma unroll
    for (int k = KR; k < KV * 8; ++k) e[k] = 0;
#pragma unroll
    for (int tap = 0; tap < NT; ++tap) {
      const int kwi = tap % 3, khi = (tap / 3) % 3, kdi = tap / 9;
      const int zd = xd + kdi - (KD >> 1), zh = xh + khi - 1, zw = xw + kwi - 1;
      const bool ok = zd >= 0 && zd < d && zh >= 0 && zh < h && zw >= 0 && zw < w;
      const unsigned short* src =
          reinterpret_cast<const unsigned short*>(x.at(ok ? ((b * d + zd) * h + zh) * w + zw : pos, 0));
#pragma unroll
      for (int ci = 0; ci < CREAL; ++ci) e[tap * CIN + ci] = ok ? src[ci] : (unsigned short)0;
      if (ONES) e[tap * CIN + CREAL] = ok ? (unsigned short)0x3F80 : (unsigned short)0;   // bf16 1.0
    }
    uint4* dst = reinterpret_cast<uint4*>(a + pos * kp);
#pragma unroll
    for (int v = 0; v < KV; ++v) {
      uint4 o;
      o.x = e[8 * v] | ((uint32_t)e[8 * v + 1] << 16); o.y = e[8 * v + 2] | ((uint32_t)e[8 * v + 3] << 16);
      o.z = e[8 * v + 4] | ((uint32_t)e[8 * v + 5] << 16); o.w = e[8 * v + 6] | ((uint32_t)e[8 * v + 7] << 16);
      dst[v] = o;
    }
    for (int v = KV; v < kv; ++v) dst[v] = make_uint4(0, 0, 0, 0);
  }
}
// K' = 64 (the 3-D one-channel first layer with its ones column: 128 bytes per position).  Same values, but a warp's
// 32 rows go through shared memory so that every store instruction writes one contiguous 512-byte run — a thread
// storing its own row writes 32 scattered 16-byte pieces per instruction (2.1 TB/s).  16-byte slots are XOR-swizzled
// by the row index: conflict-free both ways.
template <int KD, int CREAL, int ONES>
__global__ void __launch_bounds__(256)
    im2col3_k64_kernel(int n, int d, int h, int w, View x, bf16* __restrict__ a) {
  constexpr int CIN = CREAL + ONES;
  constexpr int NT = KD * 9, KR = NT * CIN;
  static_assert(KR <= 64, "one 128-byte row per position");
  __shared__ uint4 stage[8][32 * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long total = (long)n * d * h * w;
  const long nwarp_rows = (total + 31) / 32;
  for (long wr = (long)blockIdx.x * 8 + warp; wr < nwarp_rows; wr += (long)gridDim.x * 8) {
    const long pos = wr * 32 + lane;
    unsigned short e[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) e[k] = 0;
    if (pos < total) {
      long r = pos;
      const int xw = (int)(r % w); r /= w;
      const int xh = (int)(r % h); r /= h;
      const int xd = (int)(r % d); r /= d;
      const long b = r;
#pragma unroll
      for (int tap = 0; tap < NT; ++tap) {
        const int kwi = tap % 3, khi = (tap / 3) % 3, kdi = tap / 9;
        const int zd = xd + kdi - (KD >> 1), zh = xh + khi - 1, zw = xw + kwi - 1;
        const bool ok = zd >= 0 && zd < d && zh >= 0 && zh < h && zw >= 0 && zw < w;
        const unsigned short* src =
            reinterpret_cast<const unsigned short*>(x.at(ok ? ((b * d + zd) * h + zh) * w + zw : pos, 0));
#pragma unroll
        for (int ci = 0; ci < CREAL; ++ci) e[tap * CIN + ci] = ok ? src[ci] : (unsigned short)0;
        if (ONES) e[tap * CIN + CREAL] = ok ? (unsigned short)0x3F80 : (unsigned short)0;   // bf16 1.0
      }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      uint4 o;
      o.x = e[8 * v] | ((uint32_t)e[8 * v + 1] << 16); o.y = e[8 * v + 2] | ((uint32_t)e[8 * v + 3] << 16);
      o.z = e[8 * v + 4] | ((uint32_t)e[8 * v + 5] << 16); o.w = e[8 * v + 6] | ((uint32_t)e[8 * v + 7] << 16);
      stage[warp][lane * 8 + (v ^ (lane & 7))] = o;
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(a) + wr * 32 * 8;
    const long vec_left = (total - wr * 32) * 8;           // vectors of this warp's rows that exist
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = j * 32 + lane, p = idx >> 3, v = idx & 7;
      if (idx < vec_left) dst[idx] = stage[warp][p * 8 + (v ^ (p & 7))];
    }
    __syncwarp();
  }
}
// ------------------------------------------------------------------------------------ affine streams
// y = relu?(x * sc[ch] + sh[ch]) and dx = a[ch] * g + b[ch] * x + c0[ch] (g = dy, masked by y > 0):
// every normalisation forward / backward "apply" pass of the path has this form once the per-channel
// constants are folded.  A thread owns one 8-channel vector (constants live in registers) and walks
// positions 4 at a time, so there is no per-element index arithmetic and 4 x 16 B loads are in flight.
struct ColMap {
  int cv, cvb, rows, t_cv, t_row, my_cv;
  __device__ __forceinline__ ColMap(int c, int V) {
    cv = c / V;
    cvb = cv < 256 ? cv : 256;
    rows = 256 / cvb;
    t_cv = threadIdx.x % cvb;
    t_row = threadIdx.x / cvb;
    my_cv = blockIdx.y * cvb + t_cv;
  }
  __device__ __forceinline__ bool active() const { return t_row < rows && my_cv < cv; }
};

// mode 0: GroupNorm (stats = [n][groups][2] mean/rstd), mode 1: BatchNorm (p0 = mean[c], p1 = rstd[c])
template <int V>
__global__ void __launch_bounds__(256, 4)
    norm_apply_kernel(long spatial, int c, int groups, long chunk, int mode, View x, const float* p0,
                      const float* p1, const float* gamma, const float* beta, int relu, View y, int cpg = 0) {
  const ColMap m(c, V);
  if (!m.active()) return;
  const int b = blockIdx.z, ch = m.my_cv * V;
  float sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float mean, rstd;
    if (mode == 0) {
      const int g = (ch + j) / (cpg ? cpg : c / groups);   // cpg given: x is a channel prefix of the normalised tensor
      mean = p0[(b * groups + g) * 2];
      rstd = p0[(b * groups + g) * 2 + 1];
    } else {
      mean = p0[ch + j];
      rstd = p1[ch + j];
    }
    sc[j] = rstd * gamma[ch + j];
    sh[j] = beta[ch + j] - mean * sc[j];
  }
  const long base = (long)b * spatial;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < spatial ? pbeg + chunk : spatial;
  if (V == 8) {
    // 4 positions per trip, raw 16-byte loads all issued before the first use (4 x 16 B in flight per
    // thread at 1024 threads per SM is what HBM latency needs), unpacked only when consumed
    for (long p = pbeg + m.t_row; p < pend; p += 4L * m.rows) {
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long q = p + (long)u * m.rows;
        if (q < pend) r[u] = ldg_nc_v4(x.at(base + q, ch));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long q = p + (long)u * m.rows;
        if (q < pend) {
          uint4 o;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            float2 f = unpack_bf16x2((&r[u].x)[w]);
            f.x = fmaf(f.x, sc[2 * w], sh[2 * w]);
            f.y = fmaf(f.y, sc[2 * w + 1], sh[2 * w + 1]);
            if (relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
            (&o.x)[w] = pack_bf16x2(f.x, f.y);
          }
          *reinterpret_cast<uint4*>(y.at(base + q, ch)) = o;
        }
      }
    }
    return;
  }
  for (long p = pbeg + m.t_row; p < pend; p += 2L * m.rows) {
    float f[2][V];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (p + (long)u * m.rows < pend) Vec<V>::load(x.at(base + p + (long)u * m.rows, ch), f[u]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p + (long)u * m.rows < pend) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float v = fmaf(f[u][j], sc[j], sh[j]);
          f[u][j] = relu ? fmaxf(v, 0.f) : v;
        }
        Vec<V>::store(y.at(base + p + (long)u * m.rows, ch), f[u]);
      }
    }
  }
}

// mode 0: GroupNorm backward (p0 = stats, red = gsum [n][groups][2]); mode 1: BatchNorm backward
// (p0 = mean, p1 = rstd, red = [2][c] or NULL for frozen statistics); ro = ReLU output for masking
template <int V>
__global__ void __launch_bounds__(256, 4)
    norm_bwd_apply_kernel(long spatial, int c, int groups, long chunk, int mode, float inv_m, View x,
                          View dy, View ro, const float* p0, const float* p1, const float* gamma,
                          const float* red, View dx, int mask_x, int cpg = 0) {
  const ColMap m(c, V);
  if (!m.active()) return;
  const int b = blockIdx.z, ch = m.my_cv * V;
  if (cpg == 0) cpg = c / groups;   // cpg given: the kernel runs on a channel prefix of a wider GroupNorm
  float ka[V], kb[V], kc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float mean, rstd, r0, r1;
    if (mode == 0) {
      const int g = b * groups + (ch + j) / cpg;
      mean = p0[g * 2]; rstd = p0[g * 2 + 1];
      r0 = red[g * 2]; r1 = red[g * 2 + 1];
      ka[j] = rstd * gamma[ch + j];
      kb[j] = -rstd * rstd * r1 * inv_m;
      kc[j] = -rstd * inv_m * r0 - mean * kb[j];
    } else {
      mean = p0[ch + j]; rstd = p1[ch + j];
      r0 = red ? red[ch + j] : 0.f; r1 = red ? red[c + ch + j] : 0.f;
      const float gr = gamma[ch + j] * rstd;
      ka[j] = gr;
      kb[j] = -gr * rstd * r1 * inv_m;
      kc[j] = -gr * r0 * inv_m - mean * kb[j];
    }
  }
  const long base = (long)b * spatial;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < spatial ? pbeg + chunk : spatial;
  if (V == 8 && !ro.p) {
    // 2 positions per trip: the four raw 16-byte loads (x, dy of both positions) are issued before the first
    // use — with one position per trip the kernel ran at 3.6 TB/s (55 % of the HBM copy peak, round-1 profile)
    constexpr int U = 2;
    for (long p = pbeg + m.t_row; p < pend; p += (long)U * m.rows) {
      uint4 rx[U], rg[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long q = p + (long)u * m.rows;
        if (q < pend) {
          rx[u] = ldg_nc_v4(x.at(base + q, ch));
          rg[u] = ldg_nc_v4(dy.at(base + q, ch));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long q = p + (long)u * m.rows;
        if (q < pend) {
          uint4 o;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float2 xf = unpack_bf16x2((&rx[u].x)[w]);
            const float2 gf = unpack_bf16x2((&rg[u].x)[w]);
            float a0 = fmaf(ka[2 * w], gf.x, fmaf(kb[2 * w], xf.x, kc[2 * w]));
            float a1 = fmaf(ka[2 * w + 1], gf.y, fmaf(kb[2 * w + 1], xf.y, kc[2 * w + 1]));
            if (mask_x) {
              if (!(xf.x > 0.f)) a0 = 0.f;
              if (!(xf.y > 0.f)) a1 = 0.f;
            }
            (&o.x)[w] = pack_bf16x2(a0, a1);
          }
          *reinterpret_cast<uint4*>(dx.at(base + q, ch)) = o;
        }
      }
    }
    return;
  }
  for (long p = pbeg + m.t_row; p < pend; p += 1L * m.rows) {
    float xv[1][V], g[1][V], o[1][V];
#pragma unroll
    for (int u = 0; u < 1; ++u)
      if (p + (long)u * m.rows < pend) {
        const long q = base + p + (long)u * m.rows;
        Vec<V>::load(x.at(q, ch), xv[u]);
        Vec<V>::load(dy.at(q, ch), g[u]);
        if (ro.p) Vec<V>::load(ro.at(q, ch), o[u]);
      }
#pragma unroll
    for (int u = 0; u < 1; ++u)
      if (p + (long)u * m.rows < pend) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float gj = g[u][j];
          if (ro.p && !(o[u][j] > 0.f)) gj = 0.f;
          g[u][j] = fmaf(ka[j], gj, fmaf(kb[j], xv[u][j], kc[j]));
          if (mask_x && !(xv[u][j] > 0.f)) g[u][j] = 0.f;  // x is a ReLU output of the layer before
        }
        Vec<V>::store(dx.at(base + p + (long)u * m.rows, ch), g[u]);
      }
  }
}

// GroupNorm backward over the two halves of a decoder concat buffer (skip | nearest-2x-upsampled), with the
// backward of the up-sampling folded in: for the up-sampled channels every child of a source voxel q has
// the SAME x (= x_low[q]), so   dx_low[q] = sum_children (ka g + kb x + kc) = ka * sum g + nchild (kb x + kc)
// (masked by x > 0 when x is a ReLU output).  The full-resolution gradient of that half is never written
// and the stand-alone up-sampling backward pass disappears.  One thread = one source voxel x 8 channels,
// the up to 8 child loads are issued before the first use.
__global__ void __launch_bounds__(256, 4)
    gn_bwd_up_kernel(int n, int ld, int lh, int lw, int fd, int fh, int fw, int cx, int coff_up, int cpg,
                     int groups, float inv_m, View xcat, View dy, const float* stats, const float* gamma,
                     const float* gsum, View dlow, int mask_x, View xlow) {
  const int cv = cx / 8;
  const long total = (long)n * ld * lh * lw * cv;
  const int od = ld * fd, oh = lh * fh, ow = lw * fw;
  const float nchild = (float)(fd * fh * fw);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * 8; r /= cv;
    const int xw = (int)(r % lw); r /= lw;
    const int xh = (int)(r % lh); r /= lh;
    const int xd = (int)(r % ld); r /= ld;
    const int b = (int)r;
    uint4 rg[8];
    const long p0 = (((long)b * od + xd * fd) * oh + xh * fh) * ow + xw * fw;
#pragma unroll
    for (int q = 0; q < 8; ++q) {          // child q = (a, e, g): static register indices, predicated loads
      const int a = q >> 2, e = (q >> 1) & 1, g = q & 1;
      if (a < fd && e < fh && g < fw) rg[q] = ldg_nc_v4(dy.at(p0 + ((long)a * oh + e) * ow + g, coff_up + ch));
      else rg[q] = make_uint4(0, 0, 0, 0);
    }
    // x of the up-sampled half: its first child in the concat buffer, or (when the forward never materialised
    // that half, b200_gn_apply_up) the low-resolution source itself
    const uint4 rx = xlow.p ? ldg_nc_v4(xlow.at((((long)b * ld + xd) * lh + xh) * lw + xw, ch))
                            : ldg_nc_v4(xcat.at(p0, coff_up + ch));
    float sg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sg[j] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 f = unpack_bf16x2((&rg[q].x)[w]);
        sg[2 * w] += f.x; sg[2 * w + 1] += f.y;
      }
    }
    uint4 o;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 xf = unpack_bf16x2((&rx.x)[w]);
      float res[2];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int c_abs = coff_up + ch + 2 * w + h2;
        const int gi = b * groups + c_abs / cpg;
        const float mean = stats[gi * 2], rstd = stats[gi * 2 + 1];
        const float r0 = gsum[gi * 2], r1 = gsum[gi * 2 + 1];
        const float ka = rstd * gamma[c_abs];
        const float kb = -rstd * rstd * r1 * inv_m;
        const float kc = -rstd * inv_m * r0 - mean * kb;
        const float xv = h2 ? xf.y : xf.x;
        float v = fmaf(ka, sg[2 * w + h2], nchild * fmaf(kb, xv, kc));
        if (mask_x && !(xv > 0.f)) v = 0.f;
        res[h2] = v;
      }
      (&o.x)[w] = pack_bf16x2(res[0], res[1]);
    }
    *reinterpret_cast<uint4*>(dlow.at((((long)b * ld + xd) * lh + xh) * lw + xw, ch)) = o;
  }
}

// GroupNorm apply over the up-sampled half of a decoder concat WITHOUT materialising it: one thread = one source
// voxel x 8 channels of the low-resolution tensor, normalised once (+ ReLU) and written to its fd*fh*fw children in
// channels [coff_up, coff_up + cx) of y.  Replaces upsample_exact_fwd (1 read + 8 writes) followed by the apply's
// re-read of those 8 copies.
__global__ void __launch_bounds__(256, 4)
    gn_apply_up_kernel(int n, int ld, int lh, int lw, int fd, int fh, int fw, int cx, int coff_up, int cpg,
                       int groups, View xlow, const float* stats, const float* gamma, const float* beta,
                       int relu, View y) {
  const int cv = cx / 8;
  const long total = (long)n * ld * lh * lw * cv;
  const int oh = lh * fh, ow = lw * fw, od = ld * fd;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * 8; r /= cv;
    const int xw = (int)(r % lw); r /= lw;
    const int xh = (int)(r % lh); r /= lh;
    const int xd = (int)(r % ld); r /= ld;
    const int b = (int)r;
    const uint4 rx = ldg_nc_v4(xlow.at((((long)b * ld + xd) * lh + xh) * lw + xw, ch));
    uint4 o;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      float2 f = unpack_bf16x2((&rx.x)[w]);
      float res[2] = {f.x, f.y};
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int c_abs = coff_up + ch + 2 * w + h2;
        const int gi = b * groups + c_abs / cpg;
        const float sc = stats[gi * 2 + 1] * gamma[c_abs];
        const float v = fmaf(res[h2], sc, beta[c_abs] - stats[gi * 2] * sc);
        res[h2] = relu ? fmaxf(v, 0.f) : v;
      }
      (&o.x)[w] = pack_bf16x2(res[0], res[1]);
    }
    const long p0 = (((long)b * od + xd * fd) * oh + xh * fh) * ow + (long)xw * fw;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int a = q >> 2, e = (q >> 1) & 1, g = q & 1;
      if (a < fd && e < fh && g < fw)
        *reinterpret_cast<uint4*>(y.at(p0 + ((long)a * oh + e) * ow + g, coff_up + ch)) = o;
    }
  }
}

// ------------------------------------------------------------------------------------ GroupNorm
// sums: double [n][2][c] per-sample per-channel (sum x, sum x^2); finalize folds channels -> groups
template <int V>
__global__ void __launch_bounds__(256, 4)
    gn_sums_kernel(long spatial, int c, long chunk, View x, double* sums) {
  const int b = blockIdx.z;
  double* outs[2] = {sums + (long)b * c * 2, sums + (long)b * c * 2 + c};
  auto f = [&](long pos, int ch, float(*acc)[V]) {
    float xv[V];
    Vec<V>::load(x.at((long)b * spatial + pos, ch), xv);
#pragma unroll
    for (int j = 0; j < V; ++j) { acc[0][j] += xv[j]; acc[1][j] += xv[j] * xv[j]; }
  };
  column_pass<V, 2, true, double>(spatial, c, chunk, f, outs);
}
// 2x max-pool forward that also leaves the per-sample per-channel (sum y, sum y^2) of its output in sums
// fp64 [n][2][c]: the GroupNorm that opens the next encoder level (`gcr`) then needs no statistics pass over y.
template <int PD>
__global__ void __launch_bounds__(256, 4)
    maxpool_fwd_sums_kernel(int d, int h, int w, int c, long chunk, View x, View y, double* sums) {
  const int b = blockIdx.z;
  const int od = d / PD, oh = h / 2, ow = w / 2;
  const long ospatial = (long)od * oh * ow;
  double* outs[2] = {sums + (long)b * c * 2, sums + (long)b * c * 2 + c};
  auto f = [&](long pos, int ch, float(*acc)[8]) {
    const int xw = (int)(pos % ow), xh = (int)((pos / ow) % oh), xd = (int)(pos / ((long)ow * oh));
    uint4 r[PD * 4];
#pragma unroll
    for (int q = 0; q < PD * 4; ++q) {
      const int a = q >> 2, e = (q >> 1) & 1, g = q & 1;
      r[q] = ldg_nc_v4(x.at((((long)b * d + xd * PD + a) * h + xh * 2 + e) * w + xw * 2 + g, ch));
    }
    uint4 o;
#pragma unroll
    for (int w2 = 0; w2 < 4; ++w2) {
      float2 m = unpack_bf16x2((&r[0].x)[w2]);
#pragma unroll
      for (int q = 1; q < PD * 4; ++q) {
        const float2 v = unpack_bf16x2((&r[q].x)[w2]);
        m.x = fmaxf(m.x, v.x);
        m.y = fmaxf(m.y, v.y);
      }
      (&o.x)[w2] = pack_bf16x2(m.x, m.y);
      acc[0][2 * w2] += m.x; acc[0][2 * w2 + 1] += m.y;
      acc[1][2 * w2] = fmaf(m.x, m.x, acc[1][2 * w2]); acc[1][2 * w2 + 1] = fmaf(m.y, m.y, acc[1][2 * w2 + 1]);
    }
    *reinterpret_cast<uint4*>(y.at((long)b * ospatial + pos, ch)) = o;
  };
  column_pass<8, 2, true, double>(ospatial, c, chunk, f, outs);
}
__global__ void gn_finalize_kernel(int n, int c, int groups, double count, float eps,
                                   const double* sums, float* stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * groups) return;
  const int b = i / groups, g = i % groups, cpg = c / groups;
  double s1 = 0, s2 = 0;
  for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
    s1 += sums[(long)b * c * 2 + ch];
    s2 += sums[(long)b * c * 2 + c + ch];
  }
  const double mean = s1 / count;
  double var = s2 / count - mean * mean;
  if (var < 0) var = 0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}
template <int V>
__global__ void gn_apply_kernel(int n, long spatial, int c, int groups, View x, const float* stats,
                                const float* gamma, const float* beta, int relu, View y) {
  const int cv = c / V, cpg = c / groups;
  const long total = (long)n * spatial * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    const int b = (int)(pos / spatial);
    float f[V];
    Vec<V>::load(x.at(pos, ch), f);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int g = (ch + j) / cpg;
      const float mean = stats[(b * groups + g) * 2], rstd = stats[(b * groups + g) * 2 + 1];
      float v = (f[j] - mean) * rstd * gamma[ch + j] + beta[ch + j];
      f[j] = relu ? fmaxf(v, 0.f) : v;
    }
    Vec<V>::store(y.at(pos, ch), f);
  }
}
// per-sample per-channel sums: s[n][c][2] = (sum dy, sum dy * xhat)
template <int V>
__global__ void __launch_bounds__(256, 4)
    gn_bwd_reduce_kernel(long spatial, int c, int groups, long chunk, View x, View dy,
                         const float* stats, float* s) {
  const int b = blockIdx.z;
  const int cpg = c / groups;
  float* outs[2] = {s + (long)b * c * 2, s + (long)b * c * 2 + c};
  auto f = [&](long pos, int ch, float(*acc)[V]) {
    float xv[V], g[V];
    const long gp = (long)b * spatial + pos;
    Vec<V>::load(x.at(gp, ch), xv);
    Vec<V>::load(dy.at(gp, ch), g);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int gi = (ch + j) / cpg;
      const float mean = stats[(b * groups + gi) * 2], rstd = stats[(b * groups + gi) * 2 + 1];
      acc[0][j] += g[j];
      acc[1][j] += g[j] * (xv[j] - mean) * rstd;
    }
  };
  column_pass<V, 2, true, float>(spatial, c, chunk, f, outs);
}
// gsum[n][g][2] = (sum_c gamma*s1, sum_c gamma*s2); dgamma += sum_n s2; dbeta += sum_n s1
__global__ void gn_bwd_finalize_kernel(int n, int c, int groups, const float* gamma, const float* s,
                                       float* gsum, float* dgamma, float* dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cpg = c / groups;
  if (i < n * groups) {
    const int b = i / groups, g = i % groups;
    float a = 0.f, bb = 0.f;
    for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
      a += gamma[ch] * s[(long)b * c * 2 + ch];
      bb += gamma[ch] * s[(long)b * c * 2 + c + ch];
    }
    gsum[2 * i] = a;
    gsum[2 * i + 1] = bb;
  }
  if (i < c) {
    float a = 0.f, bb = 0.f;
    for (int b = 0; b < n; ++b) {
      a += s[(long)b * c * 2 + i];
      bb += s[(long)b * c * 2 + c + i];
    }
    if (dbeta) atomicAdd(dbeta + i, a);
    if (dgamma) atomicAdd(dgamma + i, bb);
  }
}
// fused variant: raw[n][2][c] fp64 = (sum dy, sum dy * x) from the dgrad epilogue;
// sum dy * xhat = rstd * (sum dy*x - mean * sum dy)
__global__ void gn_bwd_finalize_raw_kernel(int n, int c, int groups, const float* gamma,
                                           const float* stats, const double* raw, float* gsum,
                                           float* dgamma, float* dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cpg = c / groups;
  if (i < n * groups) {
    const int b = i / groups, g = i % groups;
    const double mean = stats[2 * i], rstd = stats[2 * i + 1];
    double a = 0, bb = 0;
    for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
      const double s1 = raw[((long)b * 2) * c + ch], sx = raw[((long)b * 2 + 1) * c + ch];
      a += gamma[ch] * s1;
      bb += gamma[ch] * rstd * (sx - mean * s1);
    }
    gsum[2 * i] = (float)a;
    gsum[2 * i + 1] = (float)bb;
  }
  if (i < c) {
    const int g = i / cpg;
    double a = 0, bb = 0;
    for (int b = 0; b < n; ++b) {
      const double mean = stats[2 * (b * groups + g)], rstd = stats[2 * (b * groups + g) + 1];
      const double s1 = raw[((long)b * 2) * c + i], sx = raw[((long)b * 2 + 1) * c + i];
      a += s1;
      bb += rstd * (sx - mean * s1);
    }
    if (dbeta) atomicAdd(dbeta + i, (float)a);
    if (dgamma) atomicAdd(dgamma + i, (float)bb);
  }
}
// Variant for a GroupNorm WITHOUT activation whose input was never materialised (b200_gn_apply_up): the data-gradient
// epilogue correlated dy with the norm's OUTPUT y = gamma * xhat + beta instead, raw[n][2][c] = (sum dy, sum dy * y):
//   gamma * sum dy * xhat = sum dy * y - beta * sum dy        (group term: no division)
//   dgamma = that / gamma                                     (channels with |gamma| * 8 < |beta|, where the bf16
// rounding of y ~ beta would dominate — gamma == 0 included —, are left to gn_dgamma_direct_kernel)
__device__ __forceinline__ bool gn_gamma_degenerate(float gamma, float beta) {
  return gamma == 0.f || fabsf(gamma) * 8.f < fabsf(beta);
}
__global__ void gn_bwd_finalize_y_kernel(int n, int c, int groups, const float* gamma, const float* beta,
                                         const double* raw, float* gsum, float* dgamma, float* dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cpg = c / groups;
  if (i < n * groups) {
    const int b = i / groups, g = i % groups;
    double a = 0, bb = 0;
    for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
      const double s1 = raw[((long)b * 2) * c + ch], sy = raw[((long)b * 2 + 1) * c + ch];
      a += gamma[ch] * s1;
      if (gamma[ch] != 0.f) bb += sy - (double)beta[ch] * s1;   // gamma == 0: y == beta, the true term is 0
    }
    gsum[2 * i] = (float)a;
    gsum[2 * i + 1] = (float)bb;
  }
  if (i < c) {
    double a = 0, bb = 0;
    for (int b = 0; b < n; ++b) {
      const double s1 = raw[((long)b * 2) * c + i], sy = raw[((long)b * 2 + 1) * c + i];
      a += s1;
      bb += sy - (double)beta[i] * s1;
    }
    if (dbeta) atomicAdd(dbeta + i, (float)a);
    if (dgamma && !gn_gamma_degenerate(gamma[i], beta[i])) atomicAdd(dgamma + i, (float)(bb / (double)gamma[i]));
  }
}
// dgamma_c = sum dy * xhat straight from the tensors, for the (rare) channels gn_bwd_finalize_y_kernel cannot serve:
// one block per channel, which returns at once unless its gamma is degenerate.  x of channel ch >= cs comes from the
// low-resolution tensor (nearest 1x/2x up-sampling).
__global__ void gn_dgamma_direct_kernel(int n, int ld, int lh, int lw, int fd, int fh, int fw, int cs, int c,
                                        int groups, View xskip, View xlow, View dy, const float* stats,
                                        const float* gamma, const float* beta, float* dgamma) {
  const int ch = blockIdx.x;
  if (!gn_gamma_degenerate(gamma[ch], beta[ch])) return;
  const int od = ld * fd, oh = lh * fh, ow = lw * fw, cpg = c / groups;
  const long spatial = (long)od * oh * ow;
  double acc = 0;
  for (int b = 0; b < n; ++b) {
    const float mean = stats[2 * (b * groups + ch / cpg)], rstd = stats[2 * (b * groups + ch / cpg) + 1];
    float part = 0.f;
    for (long p = threadIdx.x; p < spatial; p += blockDim.x) {
      const int xw = (int)(p % ow), xh = (int)((p / ow) % oh), xd = (int)(p / ((long)ow * oh));
      const long pos = (long)b * spatial + p;
      const float xv = ch < cs ? __bfloat162float(*xskip.at(pos, ch))
                               : __bfloat162float(*xlow.at((((long)b * ld + xd / fd) * lh + xh / fh) * lw + xw / fw,
                                                           ch - cs));
      part = fmaf(__bfloat162float(*dy.at(pos, ch)), (xv - mean) * rstd, part);
    }
    acc += (double)part;
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s2 = blockDim.x / 2; s2 > 0; s2 >>= 1) {
    if ((int)threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(dgamma + ch, (float)red[0]);
}
__global__ void bn_red_from_sums_kernel(int n, int c, const double* raw, const float* mean,
                                        const float* rstd, float* red) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  double a = 0, sx = 0;
  for (int b = 0; b < n; ++b) {
    a += raw[((long)b * 2) * c + i];
    sx += raw[((long)b * 2 + 1) * c + i];
  }
  red[i] = (float)a;
  red[c + i] = (float)((double)rstd[i] * (sx - (double)mean[i] * a));
}
template <int V>
__global__ void gn_bwd_apply_kernel(int n, long spatial, int c, int groups, View x, View dy,
                                    const float* stats, const float* gamma, const float* gsum,
                                    View dx) {
  const int cv = c / V, cpg = c / groups;
  const float inv_m = 1.f / ((float)cpg * (float)spatial);
  const long total = (long)n * spatial * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    const int b = (int)(pos / spatial);
    float xv[V], g[V];
    Vec<V>::load(x.at(pos, ch), xv);
    Vec<V>::load(dy.at(pos, ch), g);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int gi = b * groups + (ch + j) / cpg;
      const float mean = stats[gi * 2], rstd = stats[gi * 2 + 1];
      const float xhat = (xv[j] - mean) * rstd;
      g[j] = rstd * (g[j] * gamma[ch + j] - (gsum[gi * 2] + xhat * gsum[gi * 2 + 1]) * inv_m);
    }
    Vec<V>::store(dx.at(pos, ch), g);
  }
}

// ------------------------------------------------------------------------------------ BatchNorm
template <int V>
__global__ void __launch_bounds__(256, 4)
    bn_stats_kernel(long npos, int c, long chunk, View x, double* sums) {
  double* outs[2] = {sums, sums + c};
  auto f = [&](long pos, int ch, float(*acc)[V]) {
    float xv[V];
    Vec<V>::load(x.at(pos, ch), xv);
#pragma unroll
    for (int j = 0; j < V; ++j) { acc[0][j] += xv[j]; acc[1][j] += xv[j] * xv[j]; }
  };
  column_pass<V, 2, true, double>(npos, c, chunk, f, outs);
}
template <int V>
__global__ void bn_apply_kernel(long npos, int c, View x, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, int relu, View y) {
  const int cv = c / V;
  const long total = npos * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    float f[V];
    Vec<V>::load(x.at(pos, ch), f);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float v = (f[j] - mean[ch + j]) * rstd[ch + j] * gamma[ch + j] + beta[ch + j];
      f[j] = relu ? fmaxf(v, 0.f) : v;
    }
    Vec<V>::store(y.at(pos, ch), f);
  }
}
template <int V>
__global__ void __launch_bounds__(256, 4)
    bn_bwd_reduce_kernel(long npos, int c, long chunk, View x, View dy, View ro, const float* mean,
                         const float* rstd, float* red) {
  float* outs[2] = {red, red + c};
  auto f = [&](long pos, int ch, float(*acc)[V]) {
    float xv[V], g[V];
    Vec<V>::load(x.at(pos, ch), xv);
    Vec<V>::load(dy.at(pos, ch), g);
    if (ro.p) {
      float o[V];
      Vec<V>::load(ro.at(pos, ch), o);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(o[j] > 0.f)) g[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      acc[0][j] += g[j];
      acc[1][j] += g[j] * (xv[j] - mean[ch + j]) * rstd[ch + j];
    }
  };
  column_pass<V, 2, true, float>(npos, c, chunk, f, outs);
}
template <int V>
__global__ void bn_bwd_apply_kernel(long npos, int c, View x, View dy, View ro, const float* mean,
                                    const float* rstd, const float* gamma, const float* red,
                                    View dx) {
  const int cv = c / V;
  const float inv_m = 1.f / (float)npos;
  const long total = npos * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    float xv[V], g[V];
    Vec<V>::load(x.at(pos, ch), xv);
    Vec<V>::load(dy.at(pos, ch), g);
    if (ro.p) {
      float o[V];
      Vec<V>::load(ro.at(pos, ch), o);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!(o[j] > 0.f)) g[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xhat = (xv[j] - mean[ch + j]) * rstd[ch + j];
      g[j] = gamma[ch + j] * rstd[ch + j] * (g[j] - red[ch + j] * inv_m - xhat * red[c + ch + j] * inv_m);
    }
    Vec<V>::store(dx.at(pos, ch), g);
  }
}

}  // namespace b200

using namespace b200;

#define GRID1D(total) (blocks_for((total), 256) < 148u * 16u ? blocks_for((total), 256) : 148u * 16u)
#define DISPATCH_V(cond, KERNEL, GRID, ...)                                  \
  do {                                                                       \
    if (cond)                                                                \
      KERNEL<8><<<GRID, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);        \
    else                                                                     \
      KERNEL<1><<<GRID, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);        \
    B200_LAUNCH_CHECK();                                                     \
  } while (0)

extern "C" {

int b200_maxpool2_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d,
                      b200_view x, b200_view y, b200_stream_t stream) {
  if (pool_d != 1 && pool_d != 2) return B200_ERR_BAD_ARG;
  const long total = (long)n * (d / pool_d) * (h / 2) * (w / 2) * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, y);
  DISPATCH_V(v, maxpool_fwd_kernel, GRID1D(total / (v ? 8 : 1)), n, d, h, w, c, pool_d, mk(x), mk(y));
  return B200_OK;
}
int b200_maxpool2_fwd_sums(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d, b200_view x,
                           b200_view y, double* sums, b200_stream_t stream) {
  if (pool_d != 1 && pool_d != 2) return B200_ERR_BAD_ARG;
  if (!sums) return B200_ERR_BAD_ARG;
  if (!vec8(c, x) || !vec8(c, y) || d % pool_d || h % 2 || w % 2) return B200_ERR_UNSUPPORTED;
  const long ospatial = (long)(d / pool_d) * (h / 2) * (w / 2);
  if ((long)n * ospatial * c <= 0) return B200_OK;
  dim3 grid;
  long chunk;
  column_grid(ospatial, c, 8, grid, chunk);
  grid.z = n;
  if (pool_d == 2) maxpool_fwd_sums_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(d, h, w, c, chunk, mk(x), mk(y), sums);
  else maxpool_fwd_sums_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(d, h, w, c, chunk, mk(x), mk(y), sums);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_maxpool2_bwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d,
                      b200_view x, b200_view dy, b200_view add, b200_view dx, int32_t mask_relu,
                      b200_stream_t stream) {
  if (pool_d != 1 && pool_d != 2) return B200_ERR_BAD_ARG;
  if ((long)n * d * h * w * c <= 0) return B200_OK;
  const long total = (long)n * ((d + pool_d - 1) / pool_d) * ((h + 1) / 2) * ((w + 1) / 2) * c;
  const bool v = vec8(c, x) && vec8(c, dy) && vec8(c, dx) && (!add.ptr || vec8(c, add));
  if (v && d % pool_d == 0 && h % 2 == 0 && w % 2 == 0) {
    const unsigned grid = GRID1D(total / 8);
#define POOL_BWD_FAST(PD, ADD, MASK) \
  maxpool_bwd_fast_kernel<PD, ADD, MASK><<<grid, 256, 0, (cudaStream_t)stream>>>(n, d, h, w, c, mk(x), mk(dy), mk(add), mk(dx))
    const bool a = add.ptr != nullptr, m = mask_relu != 0;
    if (pool_d == 2) {
      if (a && m) POOL_BWD_FAST(2, true, true); else if (a) POOL_BWD_FAST(2, true, false);
      else if (m) POOL_BWD_FAST(2, false, true); else POOL_BWD_FAST(2, false, false);
    } else {
      if (a && m) POOL_BWD_FAST(1, true, true); else if (a) POOL_BWD_FAST(1, true, false);
      else if (m) POOL_BWD_FAST(1, false, true); else POOL_BWD_FAST(1, false, false);
    }
#undef POOL_BWD_FAST
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  DISPATCH_V(v, maxpool_bwd_kernel, GRID1D(total / (v ? 8 : 1)), n, d, h, w, c, pool_d, mk(x), mk(dy),
             mk(add), mk(dx), mask_relu);
  return B200_OK;
}
int b200_maxpoolk_fwd(int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, b200_view x, b200_view y,
                      b200_stream_t stream) {
  if (k < 2 || k > 16 || h % k || w % k || !vec8(c, x) || !vec8(c, y)) return B200_ERR_UNSUPPORTED;
  const long total = (long)n * (h / k) * (w / k) * (c / 8);
  if (total <= 0) return B200_OK;
  maxpoolk_fwd_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, h, w, c, k, mk(x), mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_maxpoolk_bwd(int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, b200_view x, b200_view dy,
                      b200_view add, b200_view dx, b200_stream_t stream) {
  if (k < 2 || k > 16 || h % k || w % k || !vec8(c, x) || !vec8(c, dy) || !vec8(c, dx) ||
      (add.ptr && !vec8(c, add)))
    return B200_ERR_UNSUPPORTED;
  const long total = (long)n * (h / k) * (w / k) * (c / 8);
  if (total <= 0) return B200_OK;
  maxpoolk_bwd_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, h, w, c, k, mk(x), mk(dy), mk(add), mk(dx));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_bilinear2d_fwd(int32_t n, int32_t ih, int32_t iw, int32_t c, int32_t scale, b200_view x,
                        b200_view y, b200_stream_t stream) {
  if (scale < 1 || scale > 32 || !vec8(c, x) || !vec8(c, y)) return B200_ERR_UNSUPPORTED;
  const long rows = (long)n * ih * scale;
  if (rows <= 0 || iw <= 0) return B200_OK;
  if (rows > 0x7fffffffL) return B200_ERR_UNSUPPORTED;
  const long cells = (long)n * (ih + 1) * (iw + 1) * (c / 8);
#define BILINEAR_CELL(S_)                                                                                         \
  case S_:                                                                                                        \
    bilinear2d_fwd_cell_kernel<S_><<<blocks_for(cells, 256), 256, 0, (cudaStream_t)stream>>>(n, ih, iw, c, mk(x), \
                                                                                             mk(y));              \
    break;
  switch (scale) {
    BILINEAR_CELL(2)
    BILINEAR_CELL(4)
    BILINEAR_CELL(8)
    BILINEAR_CELL(16)
    default:
      bilinear2d_fwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(n, ih, iw, c, scale, mk(x), mk(y));
  }
#undef BILINEAR_CELL
  B200_LAUNCH_CHECK();
  return B200_OK;
}
// outputs per thread so that the grid has >= ~600 k threads but a run keeps >= 8 outputs (lead-in <= 25 %)
static inline void adjoint_runs(long lines, int len, int* runs, int* per_run) {
  long want = (600000 + lines - 1) / lines;
  int r = (int)(want < 1 ? 1 : want);
  const int rmax = len / 8 > 0 ? len / 8 : 1;
  r = r > rmax ? rmax : r;
  *per_run = (len + r - 1) / r;
  *runs = (len + *per_run - 1) / *per_run;
}
int b200_bilinear2d_bwd(int32_t n, int32_t ih, int32_t iw, int32_t c, int32_t scale, b200_view dy,
                        b200_view add, b200_view dx, void* scratch, b200_stream_t stream) {
  if (scale < 1 || scale > 32 || !vec8(c, dy) || !vec8(c, dx) || (add.ptr && !vec8(c, add)) || !scratch ||
      ((uintptr_t)scratch % 16) != 0)
    return B200_ERR_UNSUPPORTED;
  if ((long)n * ih * iw <= 0) return B200_OK;
  const int cv = c / 8;
  const long rows = (long)n * ih * scale;
  int runs, per_run;
  cudaStream_t st = (cudaStream_t)stream;
  adjoint_runs(rows * cv, iw, &runs, &per_run);
  long total = rows * runs * cv;
#define BILINEAR_BWD_W(S_)                                                                                     \
  case S_:                                                                                                     \
    bilinear2d_bwd_w_cell_kernel<S_><<<blocks_for(total, 256), 256, 0, st>>>(rows, iw, c, runs, per_run, mk(dy), \
                                                                              (float*)scratch);                \
    break;
  switch (scale) {
    BILINEAR_BWD_W(2)
    BILINEAR_BWD_W(4)
    BILINEAR_BWD_W(8)
    BILINEAR_BWD_W(16)
    default:
      bilinear2d_bwd_w_kernel<<<blocks_for(total, 256), 256, 0, st>>>(rows, iw, c, scale, runs, per_run, mk(dy),
                                                                      (float*)scratch);
  }
#undef BILINEAR_BWD_W
  B200_LAUNCH_CHECK();
  adjoint_runs((long)n * iw * cv, ih, &runs, &per_run);
  total = (long)n * runs * iw * cv;
#define BILINEAR_BWD_H(S_)                                                                                   \
  case S_:                                                                                                   \
    bilinear2d_bwd_h_cell_kernel<S_><<<blocks_for(total, 256), 256, 0, st>>>(                                \
        n, ih, iw, c, runs, per_run, (const float*)scratch, mk(add), mk(dx));                                \
    break;
  switch (scale) {
    BILINEAR_BWD_H(2)
    BILINEAR_BWD_H(4)
    BILINEAR_BWD_H(8)
    BILINEAR_BWD_H(16)
    default:
      bilinear2d_bwd_h_kernel<<<blocks_for(total, 256), 256, 0, st>>>(n, ih, iw, c, scale, runs, per_run,
                                                                      (const float*)scratch, mk(add), mk(dx));
  }
#undef BILINEAR_BWD_H
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_add_relu(int64_t npos, int32_t c, b200_view a, b200_view b, b200_view y, b200_stream_t stream) {
  if (!vec8(c, a) || !vec8(c, b) || !vec8(c, y)) return B200_ERR_UNSUPPORTED;
  const long total = (long)npos * (c / 8);
  if (total <= 0) return B200_OK;
  add_relu_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>((long)npos, c, mk(a), mk(b), mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_add(int64_t npos, int32_t c, b200_view src, b200_view dst, b200_stream_t stream) {
  if (!vec8(c, src) || !vec8(c, dst)) return B200_ERR_UNSUPPORTED;
  const long total = (long)npos * (c / 8);
  if (total <= 0) return B200_OK;
  add_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>((long)npos, c, mk(src), mk(dst));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_upsample_nearest_fwd(int32_t n, int32_t id, int32_t ih, int32_t iw, int32_t od, int32_t oh,
                              int32_t ow, int32_t c, b200_view x, b200_view y, b200_stream_t stream) {
  const long total = (long)n * od * oh * ow * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, y);
  if (pow2_factor(id, od) && pow2_factor(ih, oh) && pow2_factor(iw, ow)) {   // floor(dst * in / out) == dst / f
    const long src = (long)n * id * ih * iw * c;
    DISPATCH_V(v, upsample_exact_fwd_kernel, GRID1D(src / (v ? 8 : 1)), n, id, ih, iw, od / id, oh / ih,
               ow / iw, c, mk(x), mk(y));
    return B200_OK;
  }
  DISPATCH_V(v, upsample_fwd_kernel, GRID1D(total / (v ? 8 : 1)), n, id, ih, iw, od, oh, ow, c, mk(x),
             mk(y));
  return B200_OK;
}
int b200_upsample_nearest_bwd(int32_t n, int32_t id, int32_t ih, int32_t iw, int32_t od, int32_t oh,
                              int32_t ow, int32_t c, b200_view dy, b200_view dx,
                              b200_stream_t stream) {
  const long total = (long)n * id * ih * iw * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, dy) && vec8(c, dx);
  const int exact = pow2_factor(id, od) && pow2_factor(ih, oh) && pow2_factor(iw, ow);
  DISPATCH_V(v, upsample_bwd_kernel, GRID1D(total / (v ? 8 : 1)), n, id, ih, iw, od, oh, ow, c, mk(dy),
             mk(dx), exact);
  return B200_OK;
}
int b200_deconv_planes_resize_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t kd,
                                  int32_t od, int32_t oh, int32_t ow, b200_view yp, b200_view y,
                                  b200_stream_t stream) {
  const long total = (long)n * od * oh * ow * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, yp) && vec8(c, y);
  DISPATCH_V(v, planes_resize_fwd_kernel, GRID1D(total / (v ? 8 : 1)), n, d, h, w, c, kd, od, oh, ow,
             mk(yp), mk(y));
  return B200_OK;
}
int b200_deconv_planes_resize_bwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t kd,
                                  int32_t od, int32_t oh, int32_t ow, b200_view dy, b200_view dyp,
                                  b200_stream_t stream) {
  const long total = (long)(kd == 1 ? 4 : 8) * n * d * h * w * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, dy) && vec8(c, dyp);
  if (v && kd == 3 && od == 2 * d && oh == 2 * h && ow == 2 * w) {
    planes_resize_bwd2x_kernel<<<GRID1D(total / 64), 256, 0, (cudaStream_t)stream>>>(n, d, h, w, c, mk(dy), mk(dyp));
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  DISPATCH_V(v, planes_resize_bwd_kernel, GRID1D(total / (v ? 8 : 1)), n, d, h, w, c, kd, od, oh, ow,
             mk(dy), mk(dyp));
  return B200_OK;
}
int b200_copy_channels(int64_t npos, int32_t c, b200_view src, b200_view dst, b200_stream_t stream) {
  const long total = (long)npos * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, src) && vec8(c, dst);
  DISPATCH_V(v, copy_kernel, GRID1D(total / (v ? 8 : 1)), (long)npos, c, mk(src), mk(dst));
  return B200_OK;
}

int b200_relu_bwd(int64_t npos, int32_t c, b200_view dy, b200_view a, b200_view dz, float* db,
                  b200_stream_t stream) {
  if (npos <= 0) return B200_OK;
  const bool v = vec8(c, dy) && vec8(c, a) && vec8(c, dz);
  dim3 grid;
  long chunk;
  column_grid(npos, c, v ? 8 : 1, grid, chunk);
  if (v) {
    if (db) relu_bwd_kernel<8, true, true><<<grid, 256, 0, (cudaStream_t)stream>>>(npos, c, chunk, mk(dy), mk(a), mk(dz), db);
    else relu_bwd_kernel<8, true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(npos, c, chunk, mk(dy), mk(a), mk(dz), db);
  } else {
    if (db) relu_bwd_kernel<1, true, true><<<grid, 256, 0, (cudaStream_t)stream>>>(npos, c, chunk, mk(dy), mk(a), mk(dz), db);
    else relu_bwd_kernel<1, true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(npos, c, chunk, mk(dy), mk(a), mk(dz), db);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_bias_grad(int64_t npos, int32_t c, b200_view dy, float* db, b200_stream_t stream) {
  if (npos <= 0 || !db) return B200_OK;
  const bool v = vec8(c, dy);
  dim3 grid;
  long chunk;
  column_grid(npos, c, v ? 8 : 1, grid, chunk);
  if (v) relu_bwd_kernel<8, false, true><<<grid, 256, 0, (cudaStream_t)stream>>>(npos, c, chunk, mk(dy), mk(dy), mk(dy), db);
  else relu_bwd_kernel<1, false, true><<<grid, 256, 0, (cudaStream_t)stream>>>(npos, c, chunk, mk(dy), mk(dy), mk(dy), db);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_nchw_f32_to_cl_bf16(int32_t n, int32_t c, int64_t spatial, const float* src, b200_view dst,
                             b200_stream_t stream) {
  const long total = (long)n * spatial;
  if (total <= 0) return B200_OK;
  nchw_to_cl_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, c, (long)spatial, src, mk(dst));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_im2col3(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c_in, int32_t kd, int32_t kp,
                 b200_view x, void* a, b200_stream_t stream) {
  const int ones = c_in < 0 ? 1 : 0;          // c_in = -(real channels): append the synthesised ones column
  if (ones) c_in = -c_in + 1;
  if ((kd != 1 && kd != 3) || c_in <= 0 || kp % 8 || kp < kd * 9 * c_in || ((uintptr_t)a % 16)) return B200_ERR_BAD_ARG;
  const long total = (long)n * d * h * w;
  if (total <= 0) return B200_OK;
  if (c_in > 4 || kd * 9 * c_in > 64) return B200_ERR_UNSUPPORTED;
#define IM2COL(KD, CIN) im2col3_kernel<KD, CIN><<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, d, h, w, kp, mk(x), (bf16*)a)
  if (ones) {
    if (kd == 3 && c_in == 2 && kp == 64) {
      const long blocks = (total + 255) / 256;
      im2col3_k64_kernel<3, 1, 1><<<(unsigned)(blocks < 148L * 16 ? blocks : 148L * 16), 256, 0, (cudaStream_t)stream>>>(
          n, d, h, w, mk(x), (bf16*)a);
    } else if (kd == 3 && c_in == 2) im2col3_kernel<3, 1, 1><<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, d, h, w, kp, mk(x), (bf16*)a);
    else if (kd == 1 && c_in == 2) im2col3_kernel<1, 1, 1><<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, d, h, w, kp, mk(x), (bf16*)a);
    else return B200_ERR_UNSUPPORTED;
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  if (kd == 1) {
    if (c_in == 1) IM2COL(1, 1); else if (c_in == 2) IM2COL(1, 2); else if (c_in == 3) IM2COL(1, 3); else IM2COL(1, 4);
  } else {
    if (c_in == 1) IM2COL(3, 1); else if (c_in == 2) IM2COL(3, 2); else return B200_ERR_UNSUPPORTED;
  }
#undef IM2COL
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_cl_bf16_to_nchw_f32(int32_t n, int32_t c, int64_t spatial, b200_view src, float* dst,
                             b200_stream_t stream) {
  const long total = (long)n * spatial;
  if (total <= 0) return B200_OK;
  cl_to_nchw_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, c, (long)spatial, mk(src), dst);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_gn_stats(int32_t n, int64_t spatial, int32_t c, int32_t groups, float eps, b200_view x,
                  double* sums, float* stats, b200_stream_t stream) {
  if (groups <= 0 || c % groups) return B200_ERR_BAD_ARG;
  if ((long)n * spatial <= 0) return B200_OK;
  const bool v = vec8(c, x);
  dim3 grid;
  long chunk;
  column_grid(spatial, c, v ? 8 : 1, grid, chunk);
  grid.z = n;
  if (v) gn_sums_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, chunk, mk(x), sums);
  else gn_sums_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, chunk, mk(x), sums);
  B200_LAUNCH_CHECK();
  gn_finalize_kernel<<<(n * groups + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      n, c, groups, (double)(c / groups) * (double)spatial, eps, sums, stats);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_gn_apply(int32_t n, int64_t spatial, int32_t c, int32_t groups, b200_view x,
                  const float* stats, const float* gamma, const float* beta, int32_t relu,
                  b200_view y, b200_stream_t stream) {
  if ((long)n * spatial * c <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, y);
  dim3 grid;
  long chunk;
  column_grid(spatial, c, v ? 8 : 1, grid, chunk);
  grid.z = n;
  if (v) norm_apply_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, 0, mk(x), stats, nullptr, gamma, beta, relu, mk(y));
  else norm_apply_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, 0, mk(x), stats, nullptr, gamma, beta, relu, mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_gn_apply_up(int32_t n, int32_t ld, int32_t lh, int32_t lw, int32_t fd, int32_t fh, int32_t fw,
                     int32_t cs, int32_t cx, int32_t groups, b200_view xskip, b200_view xlow, const float* stats,
                     const float* gamma, const float* beta, int32_t relu, b200_view y, b200_stream_t stream) {
  const int c = cs + cx;
  if (groups <= 0 || c % groups) return B200_ERR_BAD_ARG;
  if (fd < 1 || fh < 1 || fw < 1 || fd > 2 || fh > 2 || fw > 2) return B200_ERR_UNSUPPORTED;
  const b200_view up_y{y.ptr, y.ctot, y.coff + cs};
  if (cs % 8 || cx % 8 || !vec8(cs, xskip) || !vec8(cs, y) || !vec8(cx, xlow) || !vec8(cx, up_y))
    return B200_ERR_UNSUPPORTED;
  const long spatial = (long)ld * fd * lh * fh * lw * fw;
  if ((long)n * spatial <= 0) return B200_OK;
  const int cpg = c / groups;
  dim3 grid;
  long chunk;
  column_grid(spatial, cs, 8, grid, chunk);
  grid.z = n;
  norm_apply_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(spatial, cs, groups, chunk, 0, mk(xskip), stats, nullptr,
                                                              gamma, beta, relu, mk(y), cpg);
  B200_LAUNCH_CHECK();
  const long total = (long)n * ld * lh * lw * (cx / 8);
  gn_apply_up_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, ld, lh, lw, fd, fh, fw, cx, cs, cpg, groups,
                                                                      mk(xlow), stats, gamma, beta, relu, mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_gn_bwd(int32_t n, int64_t spatial, int32_t c, int32_t groups, b200_view x,
                const float* stats, const float* gamma, b200_view dy, b200_view dx, float* dgamma,
                float* dbeta, float* scratch, int32_t mask_relu, b200_stream_t stream) {
  if (groups <= 0 || c % groups) return B200_ERR_BAD_ARG;
  const long total = (long)n * spatial * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, dy) && vec8(c, dx);
  float* s = scratch;                    // [n][2][c]
  float* gsum = scratch + (long)n * c * 2;  // [n][groups][2]
  dim3 grid;
  long chunk;
  column_grid(spatial, c, v ? 8 : 1, grid, chunk);
  grid.z = n;
  if (v) gn_bwd_reduce_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, mk(x), mk(dy), stats, s);
  else gn_bwd_reduce_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, mk(x), mk(dy), stats, s);
  B200_LAUNCH_CHECK();
  const int work = n * groups > c ? n * groups : c;
  gn_bwd_finalize_kernel<<<(work + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, c, groups, gamma, s, gsum, dgamma, dbeta);
  B200_LAUNCH_CHECK();
  column_grid(spatial, c, v ? 8 : 1, grid, chunk);
  grid.z = n;
  const float inv_m = 1.f / ((float)(c / groups) * (float)spatial);
  b200_view none{nullptr, 0, 0};
  if (v) norm_bwd_apply_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, 0, inv_m, mk(x), mk(dy), mk(none), stats, nullptr, gamma, gsum, mk(dx), mask_relu);
  else norm_bwd_apply_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, 0, inv_m, mk(x), mk(dy), mk(none), stats, nullptr, gamma, gsum, mk(dx), mask_relu);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_gn_stats_finalize(int32_t n, int64_t spatial, int32_t c, int32_t groups, float eps,
                           const double* sums, float* stats, b200_stream_t stream) {
  if (groups <= 0 || c % groups) return B200_ERR_BAD_ARG;
  if ((long)n * spatial <= 0) return B200_OK;
  gn_finalize_kernel<<<(n * groups + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      n, c, groups, (double)(c / groups) * (double)spatial, eps, sums, stats);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_gn_bwd_fused(int32_t n, int64_t spatial, int32_t c, int32_t groups, b200_view x,
                      const float* stats, const float* gamma, b200_view dy, b200_view dx,
                      float* dgamma, float* dbeta, const double* colstats, float* scratch,
                      int32_t mask_relu, b200_stream_t stream) {
  if (groups <= 0 || c % groups || !colstats || !scratch) return B200_ERR_BAD_ARG;
  const long total = (long)n * spatial * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, dy) && vec8(c, dx);
  float* gsum = scratch;  // [n][groups][2]
  const int work = n * groups > c ? n * groups : c;
  gn_bwd_finalize_raw_kernel<<<(work + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      n, c, groups, gamma, stats, colstats, gsum, dgamma, dbeta);
  B200_LAUNCH_CHECK();
  dim3 grid;
  long chunk;
  column_grid(spatial, c, v ? 8 : 1, grid, chunk);
  grid.z = n;
  const float inv_m = 1.f / ((float)(c / groups) * (float)spatial);
  b200_view none{nullptr, 0, 0};
  if (v) norm_bwd_apply_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, 0, inv_m, mk(x), mk(dy), mk(none), stats, nullptr, gamma, gsum, mk(dx), mask_relu);
  else norm_bwd_apply_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, c, groups, chunk, 0, inv_m, mk(x), mk(dy), mk(none), stats, nullptr, gamma, gsum, mk(dx), mask_relu);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_gn_bwd_fused_up(int32_t n, int32_t ld, int32_t lh, int32_t lw, int32_t fd, int32_t fh, int32_t fw,
                         int32_t cs, int32_t cx, int32_t groups, b200_view x, const float* stats,
                         const float* gamma, b200_view dy, b200_view dskip, b200_view dlow, float* dgamma,
                         float* dbeta, const double* colstats, float* scratch, int32_t mask_relu,
                         b200_view xlow, const float* beta_if_y, b200_stream_t stream) {
  const int c = cs + cx;
  if (xlow.ptr && !vec8(cx, xlow)) return B200_ERR_UNSUPPORTED;
  if (beta_if_y && !xlow.ptr) return B200_ERR_BAD_ARG;
  if (groups <= 0 || c % groups || !colstats || !scratch) return B200_ERR_BAD_ARG;
  if (fd < 1 || fh < 1 || fw < 1 || fd > 2 || fh > 2 || fw > 2) return B200_ERR_UNSUPPORTED;
  const b200_view up_x{x.ptr, x.ctot, x.coff + cs}, up_dy{dy.ptr, dy.ctot, dy.coff + cs};
  if (cs % 8 || cx % 8 || !vec8(cs, x) || !vec8(cs, dy) || !vec8(cs, dskip) || !vec8(cx, up_x) ||
      !vec8(cx, up_dy) || !vec8(cx, dlow))
    return B200_ERR_UNSUPPORTED;
  const long spatial = (long)ld * fd * lh * fh * lw * fw;
  if ((long)n * spatial <= 0) return B200_OK;
  float* gsum = scratch;  // [n][groups][2]
  const int work = n * groups > c ? n * groups : c;
  if (beta_if_y) {
    gn_bwd_finalize_y_kernel<<<(work + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, c, groups, gamma, beta_if_y,
                                                                                colstats, gsum, dgamma, dbeta);
    B200_LAUNCH_CHECK();
    if (dgamma) {
      gn_dgamma_direct_kernel<<<c, 256, 0, (cudaStream_t)stream>>>(n, ld, lh, lw, fd, fh, fw, cs, c, groups, mk(x),
                                                                  mk(xlow), mk(dy), stats, gamma, beta_if_y, dgamma);
      B200_LAUNCH_CHECK();
    }
  } else {
    gn_bwd_finalize_raw_kernel<<<(work + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        n, c, groups, gamma, stats, colstats, gsum, dgamma, dbeta);
    B200_LAUNCH_CHECK();
  }
  const float inv_m = 1.f / ((float)(c / groups) * (float)spatial);
  const int cpg = c / groups;
  // skip half: the plain apply pass on the channel prefix [0, cs)
  dim3 grid;
  long chunk;
  column_grid(spatial, cs, 8, grid, chunk);
  grid.z = n;
  b200_view none{nullptr, 0, 0};
  norm_bwd_apply_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(spatial, cs, groups, chunk, 0, inv_m, mk(x), mk(dy),
                                                                  mk(none), stats, nullptr, gamma, gsum, mk(dskip),
                                                                  mask_relu, cpg);
  B200_LAUNCH_CHECK();
  // up-sampled half, reduced to the source resolution
  const long total = (long)n * ld * lh * lw * (cx / 8);
  gn_bwd_up_kernel<<<GRID1D(total), 256, 0, (cudaStream_t)stream>>>(n, ld, lh, lw, fd, fh, fw, cx, cs, cpg, groups,
                                                                    inv_m, mk(x), mk(dy), stats, gamma, gsum,
                                                                    mk(dlow), mask_relu, mk(xlow));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_bn_bwd_red_from_sums(int32_t n, int32_t c, const double* colstats, const float* mean,
                              const float* rstd, float* red, b200_stream_t stream) {
  if (n <= 0 || c <= 0 || !colstats || !red) return B200_ERR_BAD_ARG;
  bn_red_from_sums_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, c, colstats, mean, rstd, red);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_bn_stats(int64_t npos, int32_t c, b200_view x, double* sums, b200_stream_t stream) {
  if (npos <= 0) return B200_OK;
  const bool v = vec8(c, x);
  dim3 grid;
  long chunk;
  column_grid(npos, c, v ? 8 : 1, grid, chunk);
  if (v) bn_stats_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, chunk, mk(x), sums);
  else bn_stats_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, chunk, mk(x), sums);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_bn_apply(int64_t npos, int32_t c, b200_view x, const float* mean, const float* rstd,
                  const float* gamma, const float* beta, int32_t relu, b200_view y,
                  b200_stream_t stream) {
  if ((long)npos * c <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, y);
  dim3 grid;
  long chunk;
  column_grid(npos, c, v ? 8 : 1, grid, chunk);
  if (v) norm_apply_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, 1, chunk, 1, mk(x), mean, rstd, gamma, beta, relu, mk(y));
  else norm_apply_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, 1, chunk, 1, mk(x), mean, rstd, gamma, beta, relu, mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_bn_bwd_reduce(int64_t npos, int32_t c, b200_view x, b200_view dy, b200_view relu_out,
                       const float* mean, const float* rstd, float* red, b200_stream_t stream) {
  if (npos <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, dy) && (!relu_out.ptr || vec8(c, relu_out));
  dim3 grid;
  long chunk;
  column_grid(npos, c, v ? 8 : 1, grid, chunk);
  if (v) bn_bwd_reduce_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, chunk, mk(x), mk(dy), mk(relu_out), mean, rstd, red);
  else bn_bwd_reduce_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, chunk, mk(x), mk(dy), mk(relu_out), mean, rstd, red);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_bn_bwd_apply(int64_t npos, int32_t c, b200_view x, b200_view dy, b200_view relu_out,
                      const float* mean, const float* rstd, const float* gamma, const float* red,
                      b200_view dx, int32_t mask_relu, b200_stream_t stream) {
  if ((long)npos * c <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, dy) && vec8(c, dx) && (!relu_out.ptr || vec8(c, relu_out));
  dim3 grid;
  long chunk;
  column_grid(npos, c, v ? 8 : 1, grid, chunk);
  const float inv_m = 1.f / (float)npos;
  if (v) norm_bwd_apply_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, 1, chunk, 1, inv_m, mk(x), mk(dy), mk(relu_out), mean, rstd, gamma, red, mk(dx), mask_relu);
  else norm_bwd_apply_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((long)npos, c, 1, chunk, 1, inv_m, mk(x), mk(dy), mk(relu_out), mean, rstd, gamma, red, mk(dx), mask_relu);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // extern "C"
