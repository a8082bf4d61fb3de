// 1x1 segmentation head, the loss reductions (Dice / BCE / IoU / F1 / cross-entropy) and the
// evaluation-metric reductions. All HBM-bound: one pass over the logits + target, fp32 math,
// fp64 block-level accumulation (one atomic per block and quantity).
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {

struct View {
  bf16* p;
  int ctot, coff;
  __device__ __forceinline__ bf16* at(long pos, int c) const { return p + pos * ctot + coff + c; }
};
static inline View mk(const b200_view& v) { return View{(bf16*)v.ptr, v.ctot, v.coff}; }

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum_d(r);
  }
  return r;  // valid in thread 0
}

// ------------------------------------------------------------------------------------ head
// 8 threads per position, each owning 8 consecutive input channels (one 16-byte load); a warp
// therefore streams 512 contiguous bytes of x per pass.  c_in is processed in passes of 64.
#define HEAD_CO_LIMIT 8
__device__ __forceinline__ void ld8(const bf16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// Forward: one thread per position (grid.y = sample, so no 64-bit div/mod per position).  The thread
// streams its 128-byte-aligned channel row with 16-byte loads (each 32-byte sector is fetched once and
// consumed by two consecutive loads) and the per-class stores of a warp are 128 contiguous bytes.
template <int HEAD_MAX_CO>
__global__ void __launch_bounds__(256)
    head_fwd_kernel(long spatial, int c_in, int c_out, View x, const float* w, const float* b,
                    float* logits) {
  extern __shared__ float sw[];  // [c_out][c_in] + [c_out]
  for (int i = threadIdx.x; i < c_out * c_in; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < c_out; i += blockDim.x) sw[c_out * c_in + i] = b ? b[i] : 0.f;
  __syncthreads();
  const long bi = blockIdx.y;
  float* lbase = logits + bi * c_out * spatial;
  for (long sp = (long)blockIdx.x * blockDim.x + threadIdx.x; sp < spatial; sp += (long)gridDim.x * blockDim.x) {
    const bf16* xr = x.at(bi * spatial + sp, 0);
    float acc[HEAD_MAX_CO];
#pragma unroll
    for (int j = 0; j < HEAD_MAX_CO; ++j) acc[j] = 0.f;
#pragma unroll 2
    for (int c0 = 0; c0 < c_in; c0 += 16) {
      float f[16];
      ld8(xr + c0, f);
      if (c0 + 8 < c_in) ld8(xr + c0 + 8, f + 8);
#pragma unroll
      for (int j = 0; j < HEAD_MAX_CO; ++j) {
        if (j < c_out) {
          const float* wr = sw + j * c_in + c0;   // warp-uniform address: shared-memory broadcast
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[j] = fmaf(f[k], wr[k], acc[j]);
          if (c0 + 8 < c_in) {
#pragma unroll
            for (int k = 8; k < 16; ++k) acc[j] = fmaf(f[k], wr[k], acc[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < HEAD_MAX_CO; ++j)
      if (j < c_out) lbase[(long)j * spatial + sp] = acc[j] + sw[c_out * c_in + j];
  }
}
// One pass over x and dlogits: dx[pos][ci] = sum_co g[co] w[co][ci] (bf16 store) and
// dw[co][ci] += sum_pos g[co] x[pos][ci], db[co] += sum_pos g[co].
// 8 threads per position, each owning 8 consecutive channels (its dw partials live in registers);
// grid.y = sample; lane `oct` fetches g[oct] and the 8 lanes of a position share it by shuffle; two
// positions per thread are in flight.
template <int HEAD_MAX_CO>
__global__ void __launch_bounds__(256)
    head_bwd_kernel(long spatial, int c_in, int c_out, long chunk, View x, const float* w,
                    const float* dlogits, View dx, float* dw, float* db, int mask_relu) {
  extern __shared__ float sw[];       // [c_out][c_in] weights, then [32][c_out][8] reduction scratch
  float* sred = sw + c_out * c_in;
  for (int i = threadIdx.x; i < c_out * c_in; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long bi = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int oct = threadIdx.x & 7, sub = lane >> 3, prow = threadIdx.x >> 3;  // 32 positions per block pass
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < spatial ? pbeg + chunk : spatial;
  const float* gbase = dlogits + bi * c_out * spatial;
  const long xbase = bi * spatial;
  for (int cb = 0; cb < c_in; cb += 64) {
    const int c0 = cb + oct * 8;
    const bool cvalid = c0 < c_in;
    float aw[HEAD_MAX_CO][8];
    float ab[HEAD_MAX_CO];
#pragma unroll
    for (int j = 0; j < HEAD_MAX_CO; ++j) {
      ab[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) aw[j][k] = 0.f;
    }
    // warp-uniform trip count (the shuffles need every lane): a warp covers positions s0 .. s0+3 and +32
    for (long s0 = pbeg + warp * 4; s0 < pend; s0 += 64) {
      long sp[2] = {s0 + sub, s0 + 32 + sub};
      bool ok[2] = {sp[0] < pend, sp[1] < pend};
      float gl[2], f[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        gl[u] = (ok[u] && oct < c_out) ? __ldg(gbase + (long)oct * spatial + sp[u]) : 0.f;
        if (ok[u] && cvalid) ld8(x.at(xbase + sp[u], c0), f[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < HEAD_MAX_CO; ++j) {
          if (j < c_out) {
            const float g = __shfl_sync(0xffffffffu, gl[u], (lane & ~7) + j);
            if (ok[u] && cvalid) {
              const float* wr = sw + j * c_in + c0;
              ab[j] += g;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                aw[j][k] = fmaf(g, f[u][k], aw[j][k]);
                o[k] = fmaf(g, wr[k], o[k]);
              }
            }
          }
        }
        if (dx.p && ok[u] && cvalid) {
          if (mask_relu) {  // x is a ReLU output: dx doubles as the pre-activation gradient
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (!(f[u][k] > 0.f)) o[k] = 0.f;
          }
          uint4 q;
          q.x = pack_bf16x2(o[0], o[1]); q.y = pack_bf16x2(o[2], o[3]);
          q.z = pack_bf16x2(o[4], o[5]); q.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(dx.at(xbase + sp[u], c0)) = q;
        }
      }
    }
    // reduce the 32 position-rows of the block, one output channel at a time
#pragma unroll
    for (int j = 0; j < HEAD_MAX_CO; ++j) {
      if (j >= c_out) break;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) sred[(prow * 8 + oct) * 8 + k] = aw[j][k];
      __syncthreads();
      if (threadIdx.x < 64) {
        const int o8 = threadIdx.x >> 3, k = threadIdx.x & 7;
        float s_ = 0.f;
        for (int r = 0; r < 32; ++r) s_ += sred[(r * 8 + o8) * 8 + k];
        const int c = cb + o8 * 8 + k;
        if (c < c_in && dw) atomicAdd(dw + (long)j * c_in + c, s_);
      }
      if (cb == 0 && db) {
        __syncthreads();
        sred[threadIdx.x] = oct == 0 ? ab[j] : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
          float s_ = 0.f;
          for (int r = 0; r < 256; r += 8) s_ += sred[r];
          atomicAdd(db + j, s_);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ head, C_in = 64
// The U-Nets of the path always feed the head 64 channels (f_maps[0]); these variants are built for HBM
// bandwidth.  A warp owns 32 consecutive positions; 8 lanes share one position, each holding 8 channels
// (one 16-byte load, so a warp-wide load is 512 contiguous bytes) and ALL 8 loads of the 32 positions are
// issued before the first use.  Weights live in registers.
//
// Forward: every lane forms 8 passes x CO partial dot products; a halving exchange over the 8 lanes of a
// position group (send the half the partner keeps, add the half received: 7 CO shuffles instead of 24 CO)
// leaves lane `oct` with the CO finished logits of pass `oct`, i.e. of position 4 oct + sub — the warp's
// store of one class is then 128 contiguous bytes.
template <int CO>
__global__ void __launch_bounds__(256)
    head_fwd64_kernel(long spatial, View x, const float* __restrict__ w, const float* __restrict__ b,
                      float* __restrict__ logits) {
  const int lane = threadIdx.x & 31, oct = lane & 7, sub = lane >> 3;
  const long bi = blockIdx.y;
  float wr[CO][8], bias[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) {
    bias[j] = b ? __ldg(b + j) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[j][k] = __ldg(w + j * 64 + oct * 8 + k);
  }
  float* lbase = logits + bi * CO * spatial;
  const long ngroups = (spatial + 31) >> 5;
  const long wstride = (long)gridDim.x * (blockDim.x >> 5);
  for (long grp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); grp < ngroups; grp += wstride) {
    const long base = grp << 5;
    uint4 xv[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      long sp = base + 4 * p + sub;
      sp = sp < spatial ? sp : spatial - 1;
      xv[p] = ldg_nc_v4(x.at(bi * spatial + sp, oct * 8));
    }
    float v[8 * CO];   // v[p * CO + j]
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float2 a0 = unpack_bf16x2(xv[p].x), a1 = unpack_bf16x2(xv[p].y), a2 = unpack_bf16x2(xv[p].z),
                   a3 = unpack_bf16x2(xv[p].w);
      const float f[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
      for (int j = 0; j < CO; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(f[k], wr[j][k], acc);
        v[p * CO + j] = acc;
      }
    }
    // halving exchange on the pass index: bit 2 with lane^4, bit 1 with lane^2, bit 0 with lane^1
    const bool b2 = oct & 4, b1 = oct & 2, b0 = oct & 1;
    float v4[4 * CO], v2[2 * CO], v1[CO];
#pragma unroll
    for (int i = 0; i < 4 * CO; ++i)
      v4[i] = (b2 ? v[4 * CO + i] : v[i]) + __shfl_xor_sync(0xffffffffu, b2 ? v[i] : v[4 * CO + i], 4);
#pragma unroll
    for (int i = 0; i < 2 * CO; ++i)
      v2[i] = (b1 ? v4[2 * CO + i] : v4[i]) + __shfl_xor_sync(0xffffffffu, b1 ? v4[i] : v4[2 * CO + i], 2);
#pragma unroll
    for (int i = 0; i < CO; ++i)
      v1[i] = (b0 ? v2[CO + i] : v2[i]) + __shfl_xor_sync(0xffffffffu, b0 ? v2[i] : v2[CO + i], 1);
    const long sp = base + 4 * oct + sub;   // lane oct holds pass oct
    if (sp < spatial) {
#pragma unroll
      for (int j = 0; j < CO; ++j) lbase[(long)j * spatial + sp] = v1[j] + bias[j];
    }
  }
}

// Backward: dx[pos][ci] = (x > 0) * sum_j g[j] w[j][ci]; dw[j][ci] += g[j] x[pos][ci]; db[j] += g[j].
// dlogits are fetched coalesced (lane l = position base + l, one 128-byte load per class) and handed to
// the 8 lanes of each position by shuffle; dw partials stay in registers for the whole kernel.
template <int CO>
__global__ void __launch_bounds__(256, 2)
    head_bwd64_kernel(int n, long spatial, View x, const float* __restrict__ w,
                      const float* __restrict__ dlogits, View dx, float* dw, float* db, int mask_relu) {
  __shared__ float sred[8][CO][64];
  __shared__ float sdb[8][CO];
  __shared__ __align__(16) float swt[CO][64];   // weights: 32 more accumulator registers would spill
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, oct = lane & 7, sub = lane >> 3;
  for (int i = threadIdx.x; i < CO * 64; i += blockDim.x) swt[i >> 6][i & 63] = __ldg(w + i);
  __syncthreads();
  float aw[CO][8], ab[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) {
    ab[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) aw[j][k] = 0.f;
  }
  // work item = (sample, 32-position group), flattened so that the grid is exactly the resident blocks
  const long gps = (spatial + 31) >> 5, ngroups = gps * n;
  const long wstride = (long)gridDim.x * (blockDim.x >> 5);
  for (long item = (long)blockIdx.x * (blockDim.x >> 5) + warp; item < ngroups; item += wstride) {
    const long bi = item / gps;
    const long base = (item - bi * gps) << 5;
    const float* gbase = dlogits + bi * CO * spatial;
    float gl[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) gl[j] = base + lane < spatial ? __ldg(gbase + (long)j * spatial + base + lane) : 0.f;
#pragma unroll
    for (int j = 0; j < CO; ++j) ab[j] += gl[j];
    // two half-groups of 4 passes: 4 x 16 B per lane in flight keeps CO = 4 within 128 registers
#pragma unroll 1
    for (int hp = 0; hp < 8; hp += 4) {
      uint4 xv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        long sp = base + 4 * (hp + q) + sub;
        sp = sp < spatial ? sp : spatial - 1;
        xv[q] = ldg_nc_v4(x.at(bi * spatial + sp, oct * 8));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = hp + q;
        const float2 a0 = unpack_bf16x2(xv[q].x), a1 = unpack_bf16x2(xv[q].y), a2 = unpack_bf16x2(xv[q].z),
                     a3 = unpack_bf16x2(xv[q].w);
        const float f[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < CO; ++j) {
          const float g = __shfl_sync(0xffffffffu, gl[j], 4 * p + sub);   // 0 beyond the tail
          const float4 w0 = *reinterpret_cast<const float4*>(&swt[j][oct * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&swt[j][oct * 8 + 4]);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            aw[j][k] = fmaf(g, f[k], aw[j][k]);
            o[k] = fmaf(g, wv[k], o[k]);
          }
        }
        const long sp = base + 4 * p + sub;
        if (dx.p && sp < spatial) {
          if (mask_relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (!(f[k] > 0.f)) o[k] = 0.f;
          }
          uint4 qv;
          qv.x = pack_bf16x2(o[0], o[1]); qv.y = pack_bf16x2(o[2], o[3]);
          qv.z = pack_bf16x2(o[4], o[5]); qv.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(dx.at(bi * spatial + sp, oct * 8)) = qv;
        }
      }
    }
  }
  // dw: sum the 4 position sub-groups of the warp, then the 8 warps of the block
#pragma unroll
  for (int j = 0; j < CO; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = aw[j][k];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      if (sub == 0) sred[warp][j][oct * 8 + k] = a;
    }
    const float t = warp_sum(ab[j]);
    if (lane == 0) sdb[warp][j] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CO * 64; i += blockDim.x) {
    const int j = i >> 6, c = i & 63;
    float a = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) a += sred[wi][j][c];
    if (dw) atomicAdd(dw + j * 64 + c, a);
  }
  if (db && threadIdx.x < CO) {
    float a = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) a += sdb[wi][threadIdx.x];
    atomicAdd(db + threadIdx.x, a);
  }
}

// ------------------------------------------------------------------------------------ seg sums
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// numerically stable BCE-with-logits: max(x,0) - x*t + log(1 + exp(-|x|))
__device__ __forceinline__ float bce_logits(float x, float t) {
  return fmaxf(x, 0.f) - x * t + log1pf(__expf(-fabsf(x)));
}

__global__ void __launch_bounds__(256)
    seg_sums_kernel(int c, long spatial, long chunk, const float* logits, const float* target,
                    int normalization, double* sums) {
  __shared__ double sh[8];
  const int row = blockIdx.y;  // n * c + channel
  const int ch = row % c;
  const float* xr = logits + (long)row * spatial;
  const float* tr = target + (long)row * spatial;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < spatial ? pbeg + chunk : spatial;
  float a[6] = {0, 0, 0, 0, 0, 0};
  for (long i = pbeg + threadIdx.x; i < pend; i += blockDim.x) {
    const float x = xr[i], t = tr[i];
    const float p = normalization == 0 ? 1.f / (1.f + expf(-x)) : x;
    a[0] += p * t; a[1] += p * p; a[2] += t * t; a[3] += p; a[4] += t;
    if (normalization == 0) a[5] += bce_logits(x, t);
  }
  for (int k = 0; k < 6; ++k) {
    const double r = block_sum_d((double)a[k], sh);
    if (threadIdx.x == 0) atomicAdd(&sums[ch * 6 + k], r);
  }
}
__global__ void seg_loss_bwd_kernel(int c, long spatial, long total, const float* logits,
                                    const float* target, const float* coef, int normalization,
                                    float* dlogits) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)((i / spatial) % c);
    const float x = logits[i], t = target[i];
    const float* k = coef + ch * 4;
    if (normalization == 0) {
      const float p = 1.f / (1.f + expf(-x));
      dlogits[i] = k[0] * (p - t) + p * (1.f - p) * (k[1] * t + k[2] * p + k[3]);
    } else {
      dlogits[i] = k[1] * t + k[2] * x + k[3];
    }
  }
}

// ------------------------------------------------------------------------------------ cross entropy
// nn.CrossEntropyLoss() semantics: labels equal to ignore_index (-100) contribute nothing and do not count in
// the mean; labels outside [0, c) (other than -100) are counted in sums[2] — the host turns a non-zero count
// into a NaN loss instead of reading out of bounds.  sums = {sum nll, valid positions, bad labels}.
#define B200_CE_IGNORE_INDEX (-100)
__global__ void __launch_bounds__(256)
    ce_fwd_kernel(int n, int c, long spatial, const float* logits, const int64_t* labels,
                  double* sums) {
  __shared__ double sh[8];
  const long total = (long)n * spatial;
  float acc = 0.f;
  unsigned valid = 0, bad = 0;
  for (long pos = (long)blockIdx.x * blockDim.x + threadIdx.x; pos < total; pos += (long)gridDim.x * blockDim.x) {
    const int64_t lab = labels[pos];
    if (lab == B200_CE_IGNORE_INDEX) continue;
    if (lab < 0 || lab >= c) { ++bad; continue; }
    const long b = pos / spatial, s = pos % spatial;
    const float* xr = logits + b * c * spatial + s;
    float m = -INFINITY;
    for (int k = 0; k < c; ++k) m = fmaxf(m, xr[(long)k * spatial]);
    float se = 0.f;
    for (int k = 0; k < c; ++k) se += expf(xr[(long)k * spatial] - m);
    acc += (m + logf(se)) - xr[lab * spatial];
    ++valid;
  }
  const double r = block_sum_d((double)acc, sh);
  __syncthreads();
  const double v = block_sum_d((double)valid, sh);
  __syncthreads();
  const double bd = block_sum_d((double)bad, sh);
  if (threadIdx.x == 0) {
    atomicAdd(sums, r);
    atomicAdd(sums + 1, v);
    if (bd != 0.0) atomicAdd(sums + 2, bd);
  }
}
__global__ void ce_bwd_kernel(int n, int c, long spatial, const float* logits, const int64_t* labels,
                              const float* scale, float* dlogits) {
  const long total = (long)n * spatial;
  const float sc = scale[0];
  for (long pos = (long)blockIdx.x * blockDim.x + threadIdx.x; pos < total; pos += (long)gridDim.x * blockDim.x) {
    const long b = pos / spatial, s = pos % spatial;
    const float* xr = logits + b * c * spatial + s;
    float* gr = dlogits + b * c * spatial + s;
    const int64_t lab = labels[pos];
    if (lab < 0 || lab >= c) {   // ignored (or invalid) position: no gradient
      for (int k = 0; k < c; ++k) gr[(long)k * spatial] = 0.f;
      continue;
    }
    float m = -INFINITY;
    for (int k = 0; k < c; ++k) m = fmaxf(m, xr[(long)k * spatial]);
    float se = 0.f;
    for (int k = 0; k < c; ++k) se += expf(xr[(long)k * spatial] - m);
    const float inv = 1.f / se;
    for (int k = 0; k < c; ++k)
      gr[(long)k * spatial] = sc * (expf(xr[(long)k * spatial] - m) * inv - (k == (int)lab ? 1.f : 0.f));
  }
}

// ------------------------------------------------------------------------------------ metrics
__device__ __forceinline__ float metric_prob(float x, int mode) {
  if (mode == 0) return 1.f / (1.f + expf(-x) + 1e-6f);
  if (mode == 1) return 1.f / (1.f + expf(-x));
  return x;
}
__global__ void __launch_bounds__(256) metric_prob_sum_kernel(long numel, const float* logits, double* psum) {
  __shared__ double sh[8];
  float acc = 0.f;
  double dacc = 0.0;
  int cnt = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long)gridDim.x * blockDim.x) {
    acc += metric_prob(logits[i], 0);
    if (++cnt == 256) { dacc += acc; acc = 0.f; cnt = 0; }
  }
  dacc += acc;
  const double r = block_sum_d(dacc, sh);
  if (threadIdx.x == 0) atomicAdd(psum, r);
}
__global__ void __launch_bounds__(256)
    metric_counts_kernel(long per_sample, long chunk, const float* logits, const float* labels,
                         const float* thr, int mode, double* counts) {
  __shared__ double sh[8];
  const int b = blockIdx.y;
  const float th = thr[0];
  const float* xr = logits + (long)b * per_sample;
  const float* lr = labels + (long)b * per_sample;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < per_sample ? pbeg + chunk : per_sample;
  unsigned ci = 0, cp = 0, cl = 0;
  for (long i = pbeg + threadIdx.x; i < pend; i += blockDim.x) {
    const bool P = metric_prob(xr[i], mode) > th;
    const bool L = lr[i] > th;
    ci += (P && L); cp += P; cl += L;
  }
  const double r0 = block_sum_d((double)ci, sh);
  if (threadIdx.x == 0) atomicAdd(&counts[b * 3 + 0], r0);
  const double r1 = block_sum_d((double)cp, sh);
  if (threadIdx.x == 0) atomicAdd(&counts[b * 3 + 1], r1);
  const double r2 = block_sum_d((double)cl, sh);
  if (threadIdx.x == 0) atomicAdd(&counts[b * 3 + 2], r2);
}
#define ARGMAX_MAX_C 16
__global__ void __launch_bounds__(256)
    metric_argmax_counts_kernel(int c, long spatial, long chunk, const float* x, const float* target,
                                double* counts) {
  __shared__ double sh[8];
  const int b = blockIdx.y;
  const float* xr = x + (long)b * c * spatial;
  const float* tr = target + (long)b * c * spatial;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < spatial ? pbeg + chunk : spatial;
  unsigned inter[ARGMAX_MAX_C], uni[ARGMAX_MAX_C];
#pragma unroll
  for (int k = 0; k < ARGMAX_MAX_C; ++k) inter[k] = uni[k] = 0;
  for (long i = pbeg + threadIdx.x; i < pend; i += blockDim.x) {
    int am = 0;
    float m = xr[i];
    for (int k = 1; k < c; ++k) {
      const float v = xr[(long)k * spatial + i];
      if (v > m) { m = v; am = k; }
    }
#pragma unroll
    for (int k = 0; k < ARGMAX_MAX_C; ++k) {
      if (k < c) {
        const bool P = (k == am);
        const bool T = ((unsigned char)tr[(long)k * spatial + i]) != 0;  // `.byte()` of the reference
        inter[k] += (P && T); uni[k] += (P || T);
      }
    }
  }
  for (int k = 0; k < c; ++k) {
    const double r0 = block_sum_d((double)inter[k], sh);
    if (threadIdx.x == 0) atomicAdd(&counts[((long)b * c + k) * 2], r0);
    const double r1 = block_sum_d((double)uni[k], sh);
    if (threadIdx.x == 0) atomicAdd(&counts[((long)b * c + k) * 2 + 1], r1);
  }
}

}  // namespace b200

using namespace b200;

static inline unsigned grid1d(long total) {
  long b = (total + 255) / 256;
  return (unsigned)(b < 148L * 16 ? (b < 1 ? 1 : b) : 148L * 16);
}
static inline long pick_chunk(long per_row, long rows) {
  long want = 148L * 8 / (rows > 0 ? rows : 1);
  if (want < 1) want = 1;
  long chunk = (per_row + want - 1) / want;
  if (chunk < 1024) chunk = 1024;
  return chunk;
}

extern "C" {

int b200_head_fwd(int32_t n, int64_t spatial, int32_t c_in, int32_t c_out, b200_view x,
                  const float* w, const float* b, float* logits, b200_stream_t stream) {
  const long total = (long)n * spatial;
  if (total <= 0) return B200_OK;
  const size_t smem = ((size_t)c_out * c_in + c_out) * sizeof(float);
  if (smem > 48 * 1024 || c_out > HEAD_CO_LIMIT || c_in % 8 || x.ctot % 8 || x.coff % 8 || n > 65535)
    return B200_ERR_UNSUPPORTED;
  long gx = (spatial + 255) / 256;
  const long cap = (148L * 16 + n - 1) / n;
  if (gx > cap) gx = cap;
  const dim3 grid((unsigned)(gx < 1 ? 1 : gx), (unsigned)n);
  if (c_in == 64 && (c_out == 1 || c_out == 2 || c_out == 4 || c_out == 8) && ((uintptr_t)x.ptr % 16) == 0) {
#define HEAD_FWD64(CO) head_fwd64_kernel<CO><<<grid, 256, 0, (cudaStream_t)stream>>>((long)spatial, mk(x), w, b, logits)
    if (c_out == 1) HEAD_FWD64(1); else if (c_out == 2) HEAD_FWD64(2); else if (c_out == 4) HEAD_FWD64(4); else HEAD_FWD64(8);
#undef HEAD_FWD64
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
#define HEAD_FWD(CO) head_fwd_kernel<CO><<<grid, 256, smem, (cudaStream_t)stream>>>((long)spatial, c_in, c_out, mk(x), w, b, logits)
  if (c_out <= 1) HEAD_FWD(1); else if (c_out <= 2) HEAD_FWD(2); else if (c_out <= 4) HEAD_FWD(4); else HEAD_FWD(8);
#undef HEAD_FWD
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_head_bwd(int32_t n, int64_t spatial, int32_t c_in, int32_t c_out, b200_view x,
                  const float* w, const float* dlogits, b200_view dx, float* dw, float* db,
                  int32_t mask_relu, b200_stream_t stream) {
  const long total = (long)n * spatial;
  if (total <= 0) return B200_OK;
  const size_t smem = ((size_t)c_out * c_in + 2048) * sizeof(float);
  if (smem > 48 * 1024 || c_out > HEAD_CO_LIMIT || c_in % 8 || x.ctot % 8 || x.coff % 8 ||
      (dx.ptr && (dx.ctot % 8 || dx.coff % 8)) || n > 65535)
    return B200_ERR_UNSUPPORTED;
  if (c_in == 64 && (c_out == 1 || c_out == 2 || c_out == 4) && ((uintptr_t)x.ptr % 16) == 0 &&
      (!dx.ptr || ((uintptr_t)dx.ptr % 16) == 0)) {
    long gx = (total + 255) / 256;
    if (gx > 148L * 2) gx = 148L * 2;              // 2 resident blocks per SM, dw partials flushed once per block
    const dim3 grid64((unsigned)gx);
#define HEAD_BWD64(CO) head_bwd64_kernel<CO><<<grid64, 256, 0, (cudaStream_t)stream>>>(n, (long)spatial, mk(x), w, dlogits, mk(dx), dw, db, mask_relu)
    if (c_out == 1) HEAD_BWD64(1); else if (c_out == 2) HEAD_BWD64(2); else HEAD_BWD64(4);
#undef HEAD_BWD64
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  long per = (148L * 16 + n - 1) / n;              // blocks per sample
  long chunk = (spatial + per - 1) / per;
  chunk = (chunk + 63) / 64 * 64;                  // whole 64-position block passes
  if (chunk < 256) chunk = 256;
  const dim3 grid((unsigned)((spatial + chunk - 1) / chunk), (unsigned)n);
#define HEAD_BWD(CO) head_bwd_kernel<CO><<<grid, 256, smem, (cudaStream_t)stream>>>((long)spatial, c_in, c_out, chunk, mk(x), w, dlogits, mk(dx), dw, db, mask_relu)
  if (c_out <= 1) HEAD_BWD(1); else if (c_out <= 2) HEAD_BWD(2); else if (c_out <= 4) HEAD_BWD(4); else HEAD_BWD(8);
#undef HEAD_BWD
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_seg_sums(int32_t n, int32_t c, int64_t spatial, const float* logits, const float* target,
                  int32_t normalization, double* sums, b200_stream_t stream) {
  if ((long)n * c * spatial <= 0) return B200_OK;
  if ((long)n * c > 65535) return B200_ERR_UNSUPPORTED;
  const long chunk = pick_chunk(spatial, (long)n * c);
  dim3 grid((unsigned)((spatial + chunk - 1) / chunk), (unsigned)(n * c));
  seg_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, (long)spatial, chunk, logits, target, normalization, sums);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_seg_loss_bwd(int32_t n, int32_t c, int64_t spatial, const float* logits,
                      const float* target, const float* coef, int32_t normalization, float* dlogits,
                      b200_stream_t stream) {
  const long total = (long)n * c * spatial;
  if (total <= 0) return B200_OK;
  seg_loss_bwd_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(c, (long)spatial, total, logits, target, coef, normalization, dlogits);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_ce_fwd(int32_t n, int32_t c, int64_t spatial, const float* logits, const int64_t* labels,
                double* nll_sum, b200_stream_t stream) {
  const long total = (long)n * spatial;
  if (total <= 0) return B200_OK;
  ce_fwd_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(n, c, (long)spatial, logits, labels, nll_sum);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_ce_bwd(int32_t n, int32_t c, int64_t spatial, const float* logits, const int64_t* labels,
                const float* scale, float* dlogits, b200_stream_t stream) {
  const long total = (long)n * spatial;
  if (total <= 0) return B200_OK;
  ce_bwd_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(n, c, (long)spatial, logits, labels, scale, dlogits);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_metric_prob_sum(int64_t numel, const float* logits, double* psum, b200_stream_t stream) {
  if (numel <= 0) return B200_OK;
  metric_prob_sum_kernel<<<grid1d(numel), 256, 0, (cudaStream_t)stream>>>((long)numel, logits, psum);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_metric_counts(int32_t n, int64_t per_sample, const float* logits, const float* labels,
                       const float* thr, int32_t mode, double* counts, b200_stream_t stream) {
  if ((long)n * per_sample <= 0) return B200_OK;
  if (n > 65535) return B200_ERR_UNSUPPORTED;
  const long chunk = pick_chunk(per_sample, n);
  dim3 grid((unsigned)((per_sample + chunk - 1) / chunk), (unsigned)n);
  metric_counts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((long)per_sample, chunk, logits, labels, thr, mode, counts);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_metric_argmax_counts(int32_t n, int32_t c, int64_t spatial, const float* x,
                              const float* target, double* counts, b200_stream_t stream) {
  if ((long)n * spatial <= 0) return B200_OK;
  if (c > ARGMAX_MAX_C || n > 65535) return B200_ERR_UNSUPPORTED;
  const long chunk = pick_chunk(spatial, n);
  dim3 grid((unsigned)((spatial + chunk - 1) / chunk), (unsigned)n);
  metric_argmax_counts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, (long)spatial, chunk, x, target, counts);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // extern "C"
