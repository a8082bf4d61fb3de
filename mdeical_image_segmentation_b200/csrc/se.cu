// Concurrent spatial and channel squeeze-and-excitation (model/unet3d/se.py:18-113, ChannelSpatialSELayer3D) on a
// channels-last bf16 tensor y (n, spatial, c), as ResNetBlockSE applies it to its output
// (model/unet3d/buildingblocks.py:328-362):   z = max(y * a[n, c], y * b[n, pos]),
//   a = sigmoid(fc2(relu(fc1(mean_pos y))))   (cSE; the two tiny FCs stay host-side library GEMMs),
//   b = sigmoid(conv1x1x1(y))                 (sSE; the 1x1x1 conv is the head kernel).
// y is a ReLU output (>= 0) and both gates are positive, so z = y * max(a, b) and z > 0 exactly where y > 0.
// All kernels: one warp per position, lanes own 8-channel vectors (c <= 1024), HBM-bound single passes.
//   channel_sums : sums[n][c] += sum_pos y                                       (the squeeze)
//   gate_fwd     : z = y * max(a, sigmoid(s))
//   gate_bwd_red : da[n][c] += sum_pos g*y*[a >= b];  ds[n][pos] = sum_c g*y*[b > a] * b(1-b)
//   gate_bwd_apl : dy = g*max(a, b) + (y > 0 ? dmean[n][c] : 0) + dsse            (dsse: sSE conv data gradient)
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {

#define SE_MAXV 4   // 8-channel vectors per lane: c <= 32 * 8 * 4 = 1024

__device__ __forceinline__ void se_ld8(const bf16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void se_st8(bf16* p, const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float se_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

struct SeView { const bf16* p; int ctot, coff; };

// grid (blocks, n); a warp walks positions warp_global, warp_global + nwarps_total, ...
__global__ void __launch_bounds__(256)
    se_channel_sums_kernel(long spatial, int c, SeView y, double* __restrict__ sums) {
  const int lane = threadIdx.x & 31, cv = c / 8;
  const long nw = (long)gridDim.x * 8, w0 = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long b = blockIdx.y;
  float acc[SE_MAXV][8];
#pragma unroll
  for (int v = 0; v < SE_MAXV; ++v)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[v][q] = 0.f;
  for (long pos = w0; pos < spatial; pos += nw) {
    const bf16* row = y.p + (b * spatial + pos) * y.ctot + y.coff;
#pragma unroll
    for (int v = 0; v < SE_MAXV; ++v) {
      const int vec = lane + 32 * v;
      if (vec < cv) {
        float f[8];
        se_ld8(row + vec * 8, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[v][q] += f[q];
      }
    }
  }
#pragma unroll
  for (int v = 0; v < SE_MAXV; ++v) {
    const int vec = lane + 32 * v;
    if (vec < cv)
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(sums + b * c + vec * 8 + q, (double)acc[v][q]);
  }
}

__global__ void __launch_bounds__(256)
    se_gate_fwd_kernel(long spatial, int c, SeView y, const float* __restrict__ a, const float* __restrict__ s,
                       bf16* __restrict__ z, int zctot, int zcoff) {
  const int lane = threadIdx.x & 31, cv = c / 8;
  const long nw = (long)gridDim.x * 8, w0 = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long b = blockIdx.y;
  for (long pos = w0; pos < spatial; pos += nw) {
    const float bg = se_sigmoid(s[b * spatial + pos]);
    const bf16* row = y.p + (b * spatial + pos) * y.ctot + y.coff;
    bf16* orow = z + (b * spatial + pos) * zctot + zcoff;
#pragma unroll
    for (int v = 0; v < SE_MAXV; ++v) {
      const int vec = lane + 32 * v;
      if (vec < cv) {
        float f[8];
        se_ld8(row + vec * 8, f);
        const float4 a0 = *reinterpret_cast<const float4*>(a + b * c + vec * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(a + b * c + vec * 8 + 4);
        const float ag[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] *= fmaxf(ag[q], bg);
        se_st8(orow + vec * 8, f);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
    se_gate_bwd_reduce_kernel(long spatial, int c, SeView y, const float* __restrict__ a, const float* __restrict__ s,
                              SeView g, double* __restrict__ da, float* __restrict__ ds) {
  const int lane = threadIdx.x & 31, cv = c / 8;
  const long nw = (long)gridDim.x * 8, w0 = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long b = blockIdx.y;
  float acc[SE_MAXV][8], ag[SE_MAXV][8];
#pragma unroll
  for (int v = 0; v < SE_MAXV; ++v) {
    const int vec = lane + 32 * v;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      acc[v][q] = 0.f;
      ag[v][q] = vec < cv ? a[b * c + vec * 8 + q] : 0.f;
    }
  }
  for (long pos = w0; pos < spatial; pos += nw) {
    const float bg = se_sigmoid(s[b * spatial + pos]);
    const bf16* yrow = y.p + (b * spatial + pos) * y.ctot + y.coff;
    const bf16* grow = g.p + (b * spatial + pos) * g.ctot + g.coff;
    float srow = 0.f;
#pragma unroll
    for (int v = 0; v < SE_MAXV; ++v) {
      const int vec = lane + 32 * v;
      if (vec < cv) {
        float fy[8], fg[8];
        se_ld8(yrow + vec * 8, fy);
        se_ld8(grow + vec * 8, fg);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float gy = fg[q] * fy[q];
          if (ag[v][q] >= bg) acc[v][q] += gy; else srow += gy;
        }
      }
    }
    srow = warp_sum(srow);
    if (lane == 0) ds[b * spatial + pos] = srow * bg * (1.f - bg);
  }
#pragma unroll
  for (int v = 0; v < SE_MAXV; ++v) {
    const int vec = lane + 32 * v;
    if (vec < cv)
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(da + b * c + vec * 8 + q, (double)acc[v][q]);
  }
}

__global__ void __launch_bounds__(256)
    se_gate_bwd_apply_kernel(long spatial, int c, SeView y, const float* __restrict__ a, const float* __restrict__ s,
                             SeView g, const float* __restrict__ dmean, SeView dsse, bf16* __restrict__ dy, int dctot,
                             int dcoff) {
  const int lane = threadIdx.x & 31, cv = c / 8;
  const long nw = (long)gridDim.x * 8, w0 = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long b = blockIdx.y;
  for (long pos = w0; pos < spatial; pos += nw) {
    const float bg = se_sigmoid(s[b * spatial + pos]);
    const long rp = b * spatial + pos;
#pragma unroll
    for (int v = 0; v < SE_MAXV; ++v) {
      const int vec = lane + 32 * v;
      if (vec < cv) {
        float fy[8], fg[8], fd[8], o[8];
        se_ld8(y.p + rp * y.ctot + y.coff + vec * 8, fy);
        se_ld8(g.p + rp * g.ctot + g.coff + vec * 8, fg);
        se_ld8(dsse.p + rp * dsse.ctot + dsse.coff + vec * 8, fd);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float ag = a[b * c + vec * 8 + q];
          o[q] = fg[q] * fmaxf(ag, bg) + (fy[q] > 0.f ? dmean[b * c + vec * 8 + q] : 0.f) + fd[q];
        }
        se_st8(dy + rp * dctot + dcoff + vec * 8, o);
      }
    }
  }
}

static inline bool se_ok(int c, const b200_view& v) {
  return c % 8 == 0 && c <= 32 * 8 * SE_MAXV && v.ctot % 8 == 0 && v.coff % 8 == 0 && ((uintptr_t)v.ptr % 16) == 0;
}
static inline SeView se_mk(const b200_view& v) { return SeView{(const bf16*)v.ptr, v.ctot, v.coff}; }
static inline dim3 se_grid(long spatial, int n) {
  long blocks = (spatial + 7) / 8;
  const long cap = (148L * 8 + n - 1) / n;       // ~8 CTAs per SM over the whole batch
  blocks = blocks > cap ? cap : blocks;
  return dim3((unsigned)(blocks < 1 ? 1 : blocks), (unsigned)n);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_se_channel_sums(int32_t n, int64_t spatial, int32_t c, b200_view y, double* sums, b200_stream_t stream) {
  if (!se_ok(c, y) || !sums || n > 65535) return B200_ERR_UNSUPPORTED;
  if ((long)n * spatial <= 0) return B200_OK;
  se_channel_sums_kernel<<<se_grid(spatial, n), 256, 0, (cudaStream_t)stream>>>((long)spatial, c, se_mk(y), sums);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_se_gate_fwd(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s, b200_view z,
                     b200_stream_t stream) {
  if (!se_ok(c, y) || !se_ok(c, z) || !a || !s || n > 65535 || ((uintptr_t)a % 16)) return B200_ERR_UNSUPPORTED;
  if ((long)n * spatial <= 0) return B200_OK;
  se_gate_fwd_kernel<<<se_grid(spatial, n), 256, 0, (cudaStream_t)stream>>>((long)spatial, c, se_mk(y), a, s,
                                                                            (bf16*)z.ptr, z.ctot, z.coff);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_se_gate_bwd_reduce(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s,
                            b200_view g, double* da, float* ds, b200_stream_t stream) {
  if (!se_ok(c, y) || !se_ok(c, g) || !a || !s || !da || !ds || n > 65535) return B200_ERR_UNSUPPORTED;
  if ((long)n * spatial <= 0) return B200_OK;
  se_gate_bwd_reduce_kernel<<<se_grid(spatial, n), 256, 0, (cudaStream_t)stream>>>((long)spatial, c, se_mk(y), a, s,
                                                                                   se_mk(g), da, ds);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_se_gate_bwd_apply(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s,
                           b200_view g, const float* dmean, b200_view dsse, b200_view dy, b200_stream_t stream) {
  if (!se_ok(c, y) || !se_ok(c, g) || !se_ok(c, dsse) || !se_ok(c, dy) || !a || !s || !dmean || n > 65535)
    return B200_ERR_UNSUPPORTED;
  if ((long)n * spatial <= 0) return B200_OK;
  se_gate_bwd_apply_kernel<<<se_grid(spatial, n), 256, 0, (cudaStream_t)stream>>>(
      (long)spatial, c, se_mk(y), a, s, se_mk(g), dmean, se_mk(dsse), (bf16*)dy.ptr, dy.ctot, dy.coff);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // extern "C"
