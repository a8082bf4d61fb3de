// HBM-bound kernels for the layer_order letters and encoder options beyond {c, g, b, r} of
// model/unet3d/buildingblocks.py: LeakyReLU ('l') and ELU ('e') forward / backward (optionally fused with
// the residual add of ResNetBlock), dropout ('d' element-wise, 'D' channel-wise) as a mask / per-channel
// scale multiply, and 2x average pooling (Encoder pool_type='avg').  Channels-last bf16, 128-bit accesses
// (8 channels per thread) when the layout allows, a scalar path otherwise.
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {
namespace {

struct AView {
  bf16* p;
  int ctot, coff;
  __device__ __forceinline__ bf16* at(long pos, int c) const { return p + pos * ctot + coff + c; }
};
inline AView mk(const b200_view& v) { return AView{(bf16*)v.ptr, v.ctot, v.coff}; }
inline bool vec8(int c, const b200_view& a) {
  return c % 8 == 0 && a.ctot % 8 == 0 && a.coff % 8 == 0 && ((uintptr_t)a.ptr % 16) == 0;
}
inline unsigned grid_for(long total) {
  const long b = (total + 255) / 256;
  const long cap = (long)sm_count() * 16;
  return (unsigned)(b < cap ? b : cap);
}

template <int V>
__device__ __forceinline__ void ld(const bf16* p, float* f) {
  if (V == 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  } else {
    f[0] = __bfloat162float(*p);
  }
}
template <int V>
__device__ __forceinline__ void st(bf16* p, const float* f) {
  if (V == 8) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  } else {
    *p = __float2bfloat16(f[0]);
  }
}

// kind: B200_ACT_RELU / LEAKY_RELU (param = negative slope) / ELU (param = alpha) / IDENTITY (plain a + b)
__device__ __forceinline__ float act_f(float x, int kind, float param) {
  if (kind == B200_ACT_LEAKY_RELU) return x > 0.f ? x : x * param;
  if (kind == B200_ACT_ELU) return x > 0.f ? x : param * expm1f(x);
  if (kind == B200_ACT_IDENTITY) return x;
  return fmaxf(x, 0.f);
}
// derivative from the (bf16-stored) OUTPUT y: ReLU y > 0; LeakyReLU y > 0 ? 1 : slope (y and x share their
// sign for slope > 0); ELU y > 0 ? 1 : y + alpha (= alpha * exp(x)); at x == 0 ATen uses the x <= 0 branch
__device__ __forceinline__ float act_df(float y, int kind, float param) {
  if (kind == B200_ACT_LEAKY_RELU) return y > 0.f ? 1.f : param;
  if (kind == B200_ACT_ELU) return y > 0.f ? 1.f : y + param;
  if (kind == B200_ACT_IDENTITY) return 1.f;
  return y > 0.f ? 1.f : 0.f;
}

template <int V>
__global__ void __launch_bounds__(256)
    act_fwd_kernel(long npos, int c, int kind, float param, AView a, AView b, AView y) {
  const int cv = c / V;
  const long total = npos * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    float u[V], v[V];
    ld<V>(a.at(pos, ch), u);
    if (b.p) {
      ld<V>(b.at(pos, ch), v);
#pragma unroll
      for (int j = 0; j < V; ++j) u[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < V; ++j) u[j] = act_f(u[j], kind, param);
    st<V>(y.at(pos, ch), u);
  }
}

template <int V>
__global__ void __launch_bounds__(256)
    act_bwd_kernel(long npos, int c, int kind, float param, AView dy, AView y, AView dx) {
  const int cv = c / V;
  const long total = npos * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    float g[V], o[V];
    ld<V>(dy.at(pos, ch), g);
    ld<V>(y.at(pos, ch), o);
#pragma unroll
    for (int j = 0; j < V; ++j) g[j] *= act_df(o[j], kind, param);
    st<V>(dx.at(pos, ch), g);
  }
}

// y = x * (keep ? 1 : 0) * scale * chan[n][c]; keep: uint8 [n*spatial][c] dense or NULL; chan fp32 [n][c] or NULL
template <int V>
__global__ void __launch_bounds__(256)
    scale_mask_kernel(int n, long spatial, int c, AView x, const uint8_t* keep, const float* chan, float scale,
                      AView y) {
  const int cv = c / V;
  const long total = (long)n * spatial * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pos = i / cv;
    const int ch = (int)(i % cv) * V;
    const long b = pos / spatial;
    float u[V];
    ld<V>(x.at(pos, ch), u);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float f = scale;
      if (keep) f = keep[pos * c + ch + j] ? f : 0.f;
      if (chan) f *= chan[b * c + ch + j];
      u[j] *= f;
    }
    st<V>(y.at(pos, ch), u);
  }
}

template <int V>
__global__ void __launch_bounds__(256)
    avgpool_fwd_kernel(int n, int d, int h, int w, int c, int pd, AView x, AView y) {
  const int od = d / pd, oh = h / 2, ow = w / 2, cv = c / V;
  const long total = (long)n * od * oh * ow * cv;
  const float inv = 1.f / (float)(4 * pd);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int xw = (int)(r % ow); r /= ow;
    const int xh = (int)(r % oh); r /= oh;
    const int xd = (int)(r % od); r /= od;
    const int b = (int)r;
    float s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = 0.f;
    for (int a = 0; a < pd; ++a)
      for (int e = 0; e < 2; ++e)
        for (int g = 0; g < 2; ++g) {
          float f[V];
          const long pos = (((long)b * d + xd * pd + a) * h + xh * 2 + e) * w + xw * 2 + g;
          ld<V>(x.at(pos, ch), f);
#pragma unroll
          for (int j = 0; j < V; ++j) s[j] += f[j];
        }
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] *= inv;
    st<V>(y.at((((long)b * od + xd) * oh + xh) * ow + xw, ch), s);
  }
}

// dx[pos] = dy[window of pos] / (4 pd) (0 for the rows / columns floor-mode pooling dropped) + add[pos];
// mask_relu: the result is zeroed where x <= 0 (x = the pooled tensor, a ReLU output)
template <int V>
__global__ void __launch_bounds__(256)
    avgpool_bwd_kernel(int n, int d, int h, int w, int c, int pd, AView x, AView dy, AView add, AView dx,
                       int mask_relu) {
  const int od = d / pd, oh = h / 2, ow = w / 2, cv = c / V;
  const long total = (long)n * d * h * w * cv;
  const float inv = 1.f / (float)(4 * pd);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int ch = (int)(r % cv) * V; r /= cv;
    const int xw = (int)(r % w); r /= w;
    const int xh = (int)(r % h); r /= h;
    const int xd = (int)(r % d); r /= d;
    const int b = (int)r;
    const long pos = (((long)b * d + xd) * h + xh) * w + xw;
    float g[V];
#pragma unroll
    for (int j = 0; j < V; ++j) g[j] = 0.f;
    const int zd = xd / pd, zh = xh / 2, zw = xw / 2;
    if (zd < od && zh < oh && zw < ow) {
      ld<V>(dy.at((((long)b * od + zd) * oh + zh) * ow + zw, ch), g);
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] *= inv;
    }
    if (add.p) {
      float a[V];
      ld<V>(add.at(pos, ch), a);
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] += a[j];
    }
    if (mask_relu) {
      float xv[V];
      ld<V>(x.at(pos, ch), xv);
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] = xv[j] > 0.f ? g[j] : 0.f;
    }
    st<V>(dx.at(pos, ch), g);
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

#define ACT_DISPATCH(vec, KERNEL, total, ...)                                              \
  do {                                                                                     \
    if (vec)                                                                               \
      KERNEL<8><<<grid_for((total) / 8), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);     \
    else                                                                                   \
      KERNEL<1><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);           \
  } while (0)

static bool act_kind_ok(int kind) {
  return kind == B200_ACT_RELU || kind == B200_ACT_LEAKY_RELU || kind == B200_ACT_ELU || kind == B200_ACT_IDENTITY;
}

int b200_act_fwd(int64_t npos, int32_t c, int32_t kind, float param, b200_view a, b200_view b, b200_view y,
                 b200_stream_t stream) {
  if (!act_kind_ok(kind) || !a.ptr || !y.ptr) return B200_ERR_BAD_ARG;
  const long total = (long)npos * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, a) && vec8(c, y) && (!b.ptr || vec8(c, b));
  ACT_DISPATCH(v, act_fwd_kernel, total, (long)npos, c, kind, param, mk(a), mk(b), mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_act_bwd(int64_t npos, int32_t c, int32_t kind, float param, b200_view dy, b200_view y, b200_view dx,
                 b200_stream_t stream) {
  if (!act_kind_ok(kind) || !dy.ptr || !y.ptr || !dx.ptr) return B200_ERR_BAD_ARG;
  const long total = (long)npos * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, dy) && vec8(c, y) && vec8(c, dx);
  ACT_DISPATCH(v, act_bwd_kernel, total, (long)npos, c, kind, param, mk(dy), mk(y), mk(dx));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_scale_mask(int32_t n, int64_t spatial, int32_t c, b200_view x, const uint8_t* keep, const float* chan,
                    float scale, b200_view y, b200_stream_t stream) {
  if (!x.ptr || !y.ptr) return B200_ERR_BAD_ARG;
  const long total = (long)n * spatial * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, y);
  ACT_DISPATCH(v, scale_mask_kernel, total, n, (long)spatial, c, mk(x), keep, chan, scale, mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_avgpool2_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d, b200_view x,
                      b200_view y, b200_stream_t stream) {
  if (pool_d != 1 && pool_d != 2) return B200_ERR_BAD_ARG;
  const long total = (long)n * (d / pool_d) * (h / 2) * (w / 2) * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, x) && vec8(c, y);
  ACT_DISPATCH(v, avgpool_fwd_kernel, total, n, d, h, w, c, pool_d, mk(x), mk(y));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_avgpool2_bwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d, b200_view x,
                      b200_view dy, b200_view add, b200_view dx, int32_t mask_relu, b200_stream_t stream) {
  if (pool_d != 1 && pool_d != 2) return B200_ERR_BAD_ARG;
  if (mask_relu && !x.ptr) return B200_ERR_BAD_ARG;
  const long total = (long)n * d * h * w * c;
  if (total <= 0) return B200_OK;
  const bool v = vec8(c, dy) && vec8(c, dx) && (!add.ptr || vec8(c, add)) && (!mask_relu || vec8(c, x));
  ACT_DISPATCH(v, avgpool_bwd_kernel, total, n, d, h, w, c, pool_d, mk(x), mk(dy), mk(add), mk(dx), mask_relu);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
