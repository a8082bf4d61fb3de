// Concurrent spatial and channel squeeze-and-excitation (model/unet3d/se.py:18-113, ChannelSpatialSELayer3D) on a
// channels-last bf16 tensor y (n, spatial, c), as ResNetBlockSE applies it to its output
// (model/unet3d/buildingblocks.py:328-362):   z = max(y * a[n, c], y * b[n, pos]),
//   a = sigmoid(fc2(relu(fc1(mean_pos y))))   (cSE; the two tiny FCs stay host-side library GEMMs),
//   b = sigmoid(conv1x1x1(y))                 (sSE; the 1x1x1 conv is the head kernel).
// y is a ReLU output (>= 0) and both gates are positive, so z = y * max(a, b) and z > 0 exactly where y > 0.
// All kernels: lanes own 8-channel vectors, several positions per warp for narrow tensors (c <= 1024), HBM-bound
// single passes.
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

// Lane mapping.  CVL lanes (a power of two) share one position and own VPL 8-channel vectors each
// (vec = sub + CVL * v); a warp therefore holds PPW = 32 / CVL positions at once and walks U such groups per trip, so
// that narrow tensors (c = 64: 8 lanes per position) still keep >= 2 KB of loads in flight per warp — the first
// version (one position per warp) ran at ~5 % of the HBM roofline on the 64-channel levels.
template <int CVL>
struct SeMap {
  static constexpr int PPW = 32 / CVL;
  int sub, pg;
  long base0, stride;
  __device__ __forceinline__ SeMap() {
    const int lane = threadIdx.x & 31;
    sub = lane % CVL;
    pg = lane / CVL;
    base0 = ((long)blockIdx.x * 8 + (threadIdx.x >> 5)) * PPW;
    stride = (long)gridDim.x * 8 * PPW;
  }
};
// sum over the lanes that share `sub` (different positions of the warp)
template <int CVL>
__device__ __forceinline__ float se_sum_over_positions(float v) {
#pragma unroll
  for (int o = CVL; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the CVL lanes of one position
template <int CVL>
__device__ __forceinline__ float se_sum_over_channels(float v) {
#pragma unroll
  for (int o = CVL >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-channel partial sums of the 8 warps of a block are folded in shared memory before they reach the fp64
// accumulators: one atomic per (block, channel) instead of one per (warp, channel) — at the 64-channel levels the
// reducing kernels used to end in ~300 k atomics onto 64 addresses and ran at 10-14 % of the HBM roofline.
template <int CVL, int VPL>
__device__ __forceinline__ void se_block_reduce_atomic(const float (&acc)[VPL][8], const SeMap<CVL>& m, int c, int cv,
                                                       double* __restrict__ dst) {
  __shared__ float red[8][SE_MAXV * 32 * 8];
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = m.sub + CVL * v;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float t = se_sum_over_positions<CVL>(acc[v][q]);
      if (m.pg == 0 && vec < cv) red[warp][vec * 8 + q] = t;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) t += red[w2][i];
    atomicAdd(dst + i, (double)t);
  }
}

template <int CVL, int VPL, int U>
__global__ void __launch_bounds__(256)
    se_channel_sums_kernel(long spatial, int c, SeView y, double* __restrict__ sums) {
  const SeMap<CVL> m;
  const int cv = c / 8;
  const long b = blockIdx.y;
  float acc[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[v][q] = 0.f;
  for (long base = m.base0; base < spatial; base += m.stride * U) {
    float f[U][VPL][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pos = base + u * m.stride + m.pg;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = m.sub + CVL * v;
        if (pos < spatial && vec < cv) se_ld8(y.p + (b * spatial + pos) * y.ctot + y.coff + vec * 8, f[u][v]);
        else
#pragma unroll
          for (int q = 0; q < 8; ++q) f[u][v][q] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[v][q] += f[u][v][q];
  }
  se_block_reduce_atomic<CVL, VPL>(acc, m, c, cv, sums + b * c);
}

template <int CVL, int VPL, int U>
__global__ void __launch_bounds__(256)
    se_gate_fwd_kernel(long spatial, int c, SeView y, const float* __restrict__ a, const float* __restrict__ s,
                       bf16* __restrict__ z, int zctot, int zcoff) {
  const SeMap<CVL> m;
  const int cv = c / 8;
  const long b = blockIdx.y;
  float ag[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = m.sub + CVL * v;
#pragma unroll
    for (int q = 0; q < 8; ++q) ag[v][q] = vec < cv ? a[b * c + vec * 8 + q] : 0.f;
  }
  for (long base = m.base0; base < spatial; base += m.stride * U) {
    float f[U][VPL][8], sg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pos = base + u * m.stride + m.pg;
      sg[u] = pos < spatial ? s[b * spatial + pos] : 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = m.sub + CVL * v;
        if (pos < spatial && vec < cv) se_ld8(y.p + (b * spatial + pos) * y.ctot + y.coff + vec * 8, f[u][v]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pos = base + u * m.stride + m.pg;
      const float bg = se_sigmoid(sg[u]);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = m.sub + CVL * v;
        if (pos < spatial && vec < cv) {
#pragma unroll
          for (int q = 0; q < 8; ++q) f[u][v][q] *= fmaxf(ag[v][q], bg);
          se_st8(z + (b * spatial + pos) * zctot + zcoff + vec * 8, f[u][v]);
        }
      }
    }
  }
}

template <int CVL, int VPL, int U>
__global__ void __launch_bounds__(256)
    se_gate_bwd_reduce_kernel(long spatial, int c, SeView y, const float* __restrict__ a, const float* __restrict__ s,
                              SeView g, double* __restrict__ da, float* __restrict__ ds) {
  const SeMap<CVL> m;
  const int cv = c / 8;
  const long b = blockIdx.y;
  float acc[VPL][8], ag[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = m.sub + CVL * v;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      acc[v][q] = 0.f;
      ag[v][q] = vec < cv ? a[b * c + vec * 8 + q] : 0.f;
    }
  }
  for (long base = m.base0; base < spatial; base += m.stride * U) {
    float fy[U][VPL][8], fg[U][VPL][8], sg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pos = base + u * m.stride + m.pg;
      sg[u] = pos < spatial ? s[b * spatial + pos] : 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = m.sub + CVL * v;
        if (pos < spatial && vec < cv) {
          se_ld8(y.p + (b * spatial + pos) * y.ctot + y.coff + vec * 8, fy[u][v]);
          se_ld8(g.p + (b * spatial + pos) * g.ctot + g.coff + vec * 8, fg[u][v]);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) fy[u][v][q] = fg[u][v][q] = 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pos = base + u * m.stride + m.pg;
      const float bg = se_sigmoid(sg[u]);
      float srow = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float gy = fg[u][v][q] * fy[u][v][q];
          if (ag[v][q] >= bg) acc[v][q] += gy; else srow += gy;
        }
      srow = se_sum_over_channels<CVL>(srow);
      if (m.sub == 0 && pos < spatial) ds[b * spatial + pos] = srow * bg * (1.f - bg);
    }
  }
  se_block_reduce_atomic<CVL, VPL>(acc, m, c, cv, da + b * c);
}

template <int CVL, int VPL, int U>
__global__ void __launch_bounds__(256)
    se_gate_bwd_apply_kernel(long spatial, int c, SeView y, const float* __restrict__ a, const float* __restrict__ s,
                             SeView g, const float* __restrict__ dmean, SeView dsse, bf16* __restrict__ dy, int dctot,
                             int dcoff) {
  const SeMap<CVL> m;
  const int cv = c / 8;
  const long b = blockIdx.y;
  float ag[VPL][8], dm[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = m.sub + CVL * v;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      ag[v][q] = vec < cv ? a[b * c + vec * 8 + q] : 0.f;
      dm[v][q] = vec < cv ? dmean[b * c + vec * 8 + q] : 0.f;
    }
  }
  for (long base = m.base0; base < spatial; base += m.stride * U) {
    float fy[U][VPL][8], fg[U][VPL][8], fd[U][VPL][8], sg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pos = base + u * m.stride + m.pg;
      const long rp = b * spatial + pos;
      sg[u] = pos < spatial ? s[rp] : 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = m.sub + CVL * v;
        if (pos < spatial && vec < cv) {
          se_ld8(y.p + rp * y.ctot + y.coff + vec * 8, fy[u][v]);
          se_ld8(g.p + rp * g.ctot + g.coff + vec * 8, fg[u][v]);
          se_ld8(dsse.p + rp * dsse.ctot + dsse.coff + vec * 8, fd[u][v]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long pos = base + u * m.stride + m.pg;
      const long rp = b * spatial + pos;
      const float bg = se_sigmoid(sg[u]);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vec = m.sub + CVL * v;
        if (pos < spatial && vec < cv) {
          float o[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            o[q] = fg[u][v][q] * fmaxf(ag[v][q], bg) + (fy[u][v][q] > 0.f ? dm[v][q] : 0.f) + fd[u][v][q];
          se_st8(dy + rp * dctot + dcoff + vec * 8, o);
        }
      }
    }
  }
}

// cv = c / 8 vectors per position -> (lanes per position, vectors per lane, positions groups per trip)
#define SE_DISPATCH(KERNEL, cv, grid, stream, ...)                                                        \
  do {                                                                                                    \
    if ((cv) <= 1) KERNEL<1, 1, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__);                                \
    else if ((cv) <= 2) KERNEL<2, 1, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__);                           \
    else if ((cv) <= 4) KERNEL<4, 1, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__);                           \
    else if ((cv) <= 8) KERNEL<8, 1, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__);                           \
    else if ((cv) <= 16) KERNEL<16, 1, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__);                         \
    else if ((cv) <= 32) KERNEL<32, 1, 4><<<grid, 256, 0, stream>>>(__VA_ARGS__);                         \
    else if ((cv) <= 64) KERNEL<32, 2, 2><<<grid, 256, 0, stream>>>(__VA_ARGS__);                         \
    else KERNEL<32, 4, 1><<<grid, 256, 0, stream>>>(__VA_ARGS__);                                         \
  } while (0)

static inline bool se_ok(int c, const b200_view& v) {
  return c % 8 == 0 && c <= 32 * 8 * SE_MAXV && v.ctot % 8 == 0 && v.coff % 8 == 0 && ((uintptr_t)v.ptr % 16) == 0;
}
static inline SeView se_mk(const b200_view& v) { return SeView{(const bf16*)v.ptr, v.ctot, v.coff}; }
static inline dim3 se_grid(long spatial, int n, int c) {
  const int cv = c / 8;
  int cvl = 1;
  while (cvl < cv && cvl < 32) cvl <<= 1;
  const long per_block = 8L * (32 / cvl) * (cv <= 32 ? 4 : (cv <= 64 ? 2 : 1));   // positions per block trip
  long blocks = (spatial + per_block - 1) / per_block;
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
  SE_DISPATCH(se_channel_sums_kernel, c / 8, se_grid(spatial, n, c), (cudaStream_t)stream, (long)spatial, c, se_mk(y),
              sums);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_se_gate_fwd(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s, b200_view z,
                     b200_stream_t stream) {
  if (!se_ok(c, y) || !se_ok(c, z) || !a || !s || n > 65535 || ((uintptr_t)a % 16)) return B200_ERR_UNSUPPORTED;
  if ((long)n * spatial <= 0) return B200_OK;
  SE_DISPATCH(se_gate_fwd_kernel, c / 8, se_grid(spatial, n, c), (cudaStream_t)stream, (long)spatial, c, se_mk(y), a,
              s, (bf16*)z.ptr, z.ctot, z.coff);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_se_gate_bwd_reduce(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s,
                            b200_view g, double* da, float* ds, b200_stream_t stream) {
  if (!se_ok(c, y) || !se_ok(c, g) || !a || !s || !da || !ds || n > 65535) return B200_ERR_UNSUPPORTED;
  if ((long)n * spatial <= 0) return B200_OK;
  SE_DISPATCH(se_gate_bwd_reduce_kernel, c / 8, se_grid(spatial, n, c), (cudaStream_t)stream, (long)spatial, c,
              se_mk(y), a, s, se_mk(g), da, ds);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_se_gate_bwd_apply(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s,
                           b200_view g, const float* dmean, b200_view dsse, b200_view dy, b200_stream_t stream) {
  if (!se_ok(c, y) || !se_ok(c, g) || !se_ok(c, dsse) || !se_ok(c, dy) || !a || !s || !dmean || n > 65535)
    return B200_ERR_UNSUPPORTED;
  if ((long)n * spatial <= 0) return B200_OK;
  SE_DISPATCH(se_gate_bwd_apply_kernel, c / 8, se_grid(spatial, n, c), (cudaStream_t)stream, (long)spatial, c,
              se_mk(y), a, s, se_mk(g), dmean, se_mk(dsse), (bf16*)dy.ptr, dy.ctot, dy.coff);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // extern "C"
