// First layer of a 3-D U-Net on a ONE-channel volume (UNet3D(1, .): SingleConv1 of the first encoder,
// model/unet3d/buildingblocks.py:116-162 with in_channels = 1): nn.Conv3d(1, 32, 3, padding=1) forward on the CUDA
// cores.
//
// On the tensor cores this layer runs with the input zero-padded to 16 channels — 15/16 of every MMA is padding and
// an N = 32 MMA is operand-fetch bound on top (0.37 ms at 128^3 x 2, 317 TF/s of useful math).  The real work is
// 27 x 32 multiply-adds per voxel = 7.2 GFLOP: this kernel takes 0.30 ms with packed fp32x2 FMAs (ncu: 16 warps per
// SM at 128 registers, issue-bound).  It reads the volume from the channel-0 plane of the padded bf16 buffer the rest
// of the executor uses (2 of every 32 bytes: 134 MB of sectors, 0.02 ms) and keeps the numerics of the tensor-core
// path: bf16 operands (the filter is rounded to bf16 as the packed filter is), fp32 accumulation, bf16 stored
// output, fp64 column statistics of the STORED output for the norm that follows.
// The matching CUDA-core WEIGHT GRADIENT (lane = output channel, the 3 x 3 x 3 window sliding through registers)
// was written and measured at 0.59-0.62 ms against 0.48 ms for im2col + short-K tensor-core GEMM, and removed.
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {
namespace {

struct FView {
  bf16* p;
  int ctot, coff;
  __device__ __forceinline__ bf16* at(long pos, int c) const { return p + pos * ctot + coff + c; }
};
inline FView mk(const b200_view& v) { return FView{(bf16*)v.ptr, v.ctot, v.coff}; }

constexpr int kTZ = 4, kTY = 8, kTX = 32;            // voxels per block tile (256 threads)
constexpr int kHZ = kTZ + 2, kHY = kTY + 2, kHX = kTX + 2;

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

// the (kTZ+2) x (kTY+2) x (kTX+2) halo of one tile, zero outside the volume, as fp32 in shared memory
__device__ __forceinline__ void load_halo(float (*s_in)[kHY][kHX], const FView& x, int b, int d, int h, int w,
                                          int z0, int y0, int x0) {
  for (int i = threadIdx.x; i < kHZ * kHY * kHX; i += blockDim.x) {
    const int xx = i % kHX, yy = (i / kHX) % kHY, zz = i / (kHX * kHY);
    const int gz = z0 + zz - 1, gy = y0 + yy - 1, gx = x0 + xx - 1;
    float v = 0.f;
    if (gz >= 0 && gz < d && gy >= 0 && gy < h && gx >= 0 && gx < w)
      v = __bfloat162float(*x.at((((long)b * d + gz) * h + gy) * w + gx, 0));
    s_in[zz][yy][xx] = v;
  }
}

// Sum of v[c] over the 32 lanes of a warp for 32 channels at once: recursive halving — in step m a lane keeps the
// half of its values whose channel bit m equals its own lane bit m and receives the partner's partials of that
// half — leaves lane l with the total of channel l in v[0] after 16 + 8 + 4 + 2 + 1 shuffles.
__device__ __forceinline__ float warp_sum_by_channel(float (&v)[32], int lane) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const float send = up ? v[i] : v[i + m];
      const float keep = up ? v[i + m] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------- forward
// One thread = one (y, x) column of the tile: its 4 voxels along z, two at a time, 32 output channels each.
// The filter sits in shared memory as [tap][32] fp32 (every lane reads the same float4: a broadcast).
__global__ void __launch_bounds__(256, 2)
    conv3d_c1_fwd_kernel(int n, int d, int h, int w, FView x, const float* __restrict__ wgt,
                         const float* __restrict__ bias, int relu, FView y, double* __restrict__ stats, int tiles_z,
                         int tiles_y, int tiles_x) {
  constexpr int C = 32;
  __shared__ float s_in[kHZ][kHY][kHX];
  __shared__ __align__(16) float s_w[27][C];
  __shared__ float s_stat[2][C];
  __shared__ uint4 s_out[8][32 * 4];
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; t /= tiles_y;
  const int tz = t % tiles_z;
  const int b = t / tiles_z;
  const int z0 = tz * kTZ, y0 = ty * kTY, x0 = tx * kTX;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) s_w[i / C][i % C] = bf16_round(wgt[(i % C) * 27 + i / C]);
  if (threadIdx.x < 2 * C) s_stat[threadIdx.x / C][threadIdx.x % C] = 0.f;
  load_halo(s_in, x, b, d, h, w, z0, y0, x0);
  __syncthreads();
  const int lane = threadIdx.x & 31, lx = lane, ly = threadIdx.x >> 5;
  const int gy = y0 + ly, gx = x0 + lx;
  const bool col_ok = gy < h && gx < w;
#pragma unroll 1
  for (int pass = 0; pass < kTZ / 2; ++pass) {
    const int zo = pass * 2;
    // fp32x2 FMAs (FFMA2, sm_100): half the instructions of the scalar form for the same 27 x 32 x 2 products
    float2 acc[2][C / 2];
#pragma unroll
    for (int c = 0; c < C / 2; ++c) {
      const float2 bv = bias ? make_float2(bias[2 * c], bias[2 * c + 1]) : make_float2(0.f, 0.f);
      acc[0][c] = bv;
      acc[1][c] = bv;
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float2 in[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float v = s_in[zo + p][ly + kh][lx + kw];
          in[p] = make_float2(v, v);
        }
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const float4* wrow = reinterpret_cast<const float4*>(s_w[(kd * 3 + kh) * 3 + kw]);
#pragma unroll
          for (int q = 0; q < C / 4; ++q) {
            const float4 wv = wrow[q];
            const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
            acc[0][2 * q] = __ffma2_rn(in[kd], w01, acc[0][2 * q]);
            acc[0][2 * q + 1] = __ffma2_rn(in[kd], w23, acc[0][2 * q + 1]);
            acc[1][2 * q] = __ffma2_rn(in[kd + 1], w01, acc[1][2 * q]);
            acc[1][2 * q + 1] = __ffma2_rn(in[kd + 1], w23, acc[1][2 * q + 1]);
          }
        }
      }
    float s1[C], s2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) s1[c] = s2[c] = 0.f;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int gz = z0 + zo + v;
      const bool ok = col_ok && gz < d;
      uint32_t pk[C / 2];
#pragma unroll
      for (int c = 0; c < C; c += 2) {
        float a0 = acc[v][c / 2].x, a1 = acc[v][c / 2].y;
        if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
        pk[c / 2] = pack_bf16x2(a0, a1);
        if (stats) {
          const float2 r = unpack_bf16x2(pk[c / 2]);      // statistics of the STORED values
          if (ok) {
            s1[c] += r.x; s1[c + 1] += r.y;
            s2[c] = fmaf(r.x, r.x, s2[c]); s2[c + 1] = fmaf(r.y, r.y, s2[c + 1]);
          }
        }
      }
      // a thread holds the 64 bytes of its voxel; through shared memory (16-byte slots XOR-swizzled by the row) the
      // warp's 32 x-consecutive voxels leave as four contiguous 512-byte stores instead of 4 x 32 scattered pieces
      uint4* st = s_out[threadIdx.x >> 5];
#pragma unroll
      for (int q = 0; q < C / 8; ++q)
        st[lane * 4 + (q ^ (lane & 3))] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      __syncwarp();
      if (gy < h && gz < d) {
        const long row0 = (((long)b * d + gz) * h + gy) * w + x0;
#pragma unroll
        for (int j = 0; j < C / 8; ++j) {
          const int idx = j * 32 + lane, pp = idx >> 2, q = idx & 3;
          if (x0 + pp < w) *reinterpret_cast<uint4*>(y.at(row0 + pp, q * 8)) = st[pp * 4 + (q ^ (pp & 3))];
        }
      }
      __syncwarp();
    }
    if (stats) {
      const float t1 = warp_sum_by_channel(s1, lane);
      const float t2 = warp_sum_by_channel(s2, lane);
      atomicAdd(&s_stat[0][lane], t1);
      atomicAdd(&s_stat[1][lane], t2);
    }
  }
  if (stats) {
    __syncthreads();
    if (threadIdx.x < 2 * C)
      atomicAdd(stats + ((long)b * 2 + threadIdx.x / C) * C + threadIdx.x % C,
                (double)s_stat[threadIdx.x / C][threadIdx.x % C]);
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_conv3d_c1_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c_out, b200_view x, const float* wgt,
                       const float* bias, int32_t relu, b200_view y, double* colstats, b200_stream_t stream) {
  if (c_out != 32 || !x.ptr || !y.ptr || !wgt) return B200_ERR_UNSUPPORTED;
  if (d < 2 || h < 2 || w < 2) return B200_ERR_UNSUPPORTED;
  if (y.ctot % 8 || y.coff % 8 || ((uintptr_t)y.ptr % 16)) return B200_ERR_UNSUPPORTED;
  if ((long)n * d * h * w <= 0) return B200_OK;
  const int tz = (d + kTZ - 1) / kTZ, ty = (h + kTY - 1) / kTY, tx = (w + kTX - 1) / kTX;
  const long blocks = (long)n * tz * ty * tx;
  if (blocks > 0x7fffffffL) return B200_ERR_UNSUPPORTED;
  conv3d_c1_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, d, h, w, mk(x), wgt, bias, relu, mk(y),
                                                                          colstats, tz, ty, tx);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // extern "C"
