// GPU-side volume augmentation: the numpy / scipy transforms of the reference's
// augment/unet3d_augment/transforms.py on dense fp32 (C, D, H, W) device tensors — the step in front of the network
// once real datasets feed several GPUs.  All kernels are element-wise or gather passes (HBM-bound; the elastic
// deformation that costs the reference "~1 s per 64 x 200 x 200 patch" (transforms.py:136) is three separable
// Gaussian passes per displacement field plus one tri-linear gather).
//   flip_rot90     : RandomFlip (:25-50) + RandomRotate90 (:53-80) as one index map
//   affine_clip    : y = clip(a[c] x + b[c], lo, hi): RandomContrast (:115-135), Standardize (:495-523), Normalize (:547-605)
//   channel_stats  : per-channel sum, sum of squares, min, max (Standardize / Normalize statistics)
//   gaussian_axis  : scipy.ndimage.gaussian_filter along one axis, mode='reflect', truncate = 4
//   map_coordinates: scipy.ndimage.map_coordinates(order 0 | 1, mode='reflect') at index + displacement (:178-190)
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {
namespace {

inline unsigned aug_grid(long total) {
  const long b = (total + 255) / 256;
  const long cap = (long)sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b < cap ? b : cap));
}

// half-sample symmetric extension (d c b a | a b c d | d c b a): scipy's mode='reflect'
__device__ __forceinline__ int reflect_idx(long i, int n) {
  const long p = 2L * n;
  long j = i % p;
  if (j < 0) j += p;
  return (int)(j >= n ? p - 1 - j : j);
}

__global__ void __launch_bounds__(256)
    flip_rot90_kernel(int c, int d, int h, int w, int flip, int k, const float* __restrict__ src,
                      float* __restrict__ dst) {
  // source dims (d, h, w); flips first (np.flip per axis), then np.rot90(m, k, axes=(1, 2)) of the flipped volume
  const int oh = (k & 1) ? w : h, ow = (k & 1) ? h : w;
  const long total = (long)c * d * oh * ow;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int x = (int)(r % ow); r /= ow;
    const int y = (int)(r % oh); r /= oh;
    const int z = (int)(r % d); r /= d;
    const int ch = (int)r;
    int fy, fx;                       // position in the flipped (un-rotated) volume
    if (k == 0) { fy = y; fx = x; }
    else if (k == 1) { fy = x; fx = w - 1 - y; }      // rot90: out[i, j] = m[j, W - 1 - i]
    else if (k == 2) { fy = h - 1 - y; fx = w - 1 - x; }
    else { fy = h - 1 - x; fx = y; }                  // k = 3: out[i, j] = m[H - 1 - j, i]
    const int sz = (flip & 1) ? d - 1 - z : z;
    const int sy = (flip & 2) ? h - 1 - fy : fy;
    const int sx = (flip & 4) ? w - 1 - fx : fx;
    dst[i] = src[(((long)ch * d + sz) * h + sy) * w + sx];
  }
}

__global__ void __launch_bounds__(256)
    affine_clip_kernel(int c, long per_c, const float* __restrict__ src, const float* __restrict__ a,
                       const float* __restrict__ b, float lo, float hi, float* __restrict__ dst) {
  const long total = (long)c * per_c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i / per_c);
    const float v = fmaf(a[ch], src[i], b[ch]);
    dst[i] = fminf(fmaxf(v, lo), hi);
  }
}

// stats[c][4] = (sum, sum of squares, min, max); caller initialises to (0, 0, +inf, -inf)
__global__ void __launch_bounds__(256)
    channel_stats_kernel(long per_c, const float* __restrict__ src, double* __restrict__ stats) {
  __shared__ double sh[4][8];
  const int ch = blockIdx.y;
  const float* p = src + (long)ch * per_c;
  double s = 0.0, s2 = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per_c; i += (long)gridDim.x * blockDim.x) {
    const float v = p[i];
    s += v; s2 += (double)v * v;
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  double dmn = mn, dmx = mx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    dmn = fmin(dmn, __shfl_xor_sync(0xffffffffu, dmn, o));
    dmx = fmax(dmx, __shfl_xor_sync(0xffffffffu, dmx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = s2; sh[2][warp] = dmn; sh[3][warp] = dmx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = sh[0][0]; s2 = sh[1][0]; dmn = sh[2][0]; dmx = sh[3][0];
    for (int k = 1; k < 8; ++k) { s += sh[0][k]; s2 += sh[1][k]; dmn = fmin(dmn, sh[2][k]); dmx = fmax(dmx, sh[3][k]); }
    atomicAdd(stats + ch * 4, s);
    atomicAdd(stats + ch * 4 + 1, s2);
    // min / max through the ordered-integer trick on the fp64 bit pattern is overkill here: one CAS loop each
    unsigned long long* pm = reinterpret_cast<unsigned long long*>(stats + ch * 4 + 2);
    unsigned long long old = *pm, assumed;
    do {
      assumed = old;
      if (__longlong_as_double((long long)assumed) <= dmn) break;
      old = atomicCAS(pm, assumed, (unsigned long long)__double_as_longlong(dmn));
    } while (assumed != old);
    pm = reinterpret_cast<unsigned long long*>(stats + ch * 4 + 3);
    old = *pm;
    do {
      assumed = old;
      if (__longlong_as_double((long long)assumed) >= dmx) break;
      old = atomicCAS(pm, assumed, (unsigned long long)__double_as_longlong(dmx));
    } while (assumed != old);
  }
}

// one axis of scipy.ndimage.gaussian_filter(sigma, mode='reflect'): weights exp(-k^2 / 2 sigma^2), |k| <= int(4 sigma + .5)
__global__ void __launch_bounds__(256)
    gaussian_axis_kernel(int n0, int n1, int n2, int axis, float sigma, int radius, float inv_norm,
                         const float* __restrict__ src, float* __restrict__ dst) {
  const long total = (long)n0 * n1 * n2;
  const int len = axis == 0 ? n0 : (axis == 1 ? n1 : n2);
  const long stride = axis == 0 ? (long)n1 * n2 : (axis == 1 ? n2 : 1);
  const float coef = -0.5f / (sigma * sigma);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int x = (int)(r % n2); r /= n2;
    const int y = (int)(r % n1); r /= n1;
    const int z = (int)r;
    const int pos = axis == 0 ? z : (axis == 1 ? y : x);
    const long base = i - (long)pos * stride;
    float acc = 0.f;
    for (int k = -radius; k <= radius; ++k)
      acc = fmaf(__expf(coef * (float)(k * k)), src[base + (long)reflect_idx((long)pos + k, len) * stride], acc);
    dst[i] = acc * inv_norm;
  }
}

// out[c, p] = interp(src[c], p + (dz, dy, dx)[p]) with the volume extended by half-sample symmetric reflection
template <int ORDER>
__global__ void __launch_bounds__(256)
    map_coordinates_kernel(int c, int d, int h, int w, const float* __restrict__ src, const float* __restrict__ dz,
                           const float* __restrict__ dy, const float* __restrict__ dx, float* __restrict__ dst) {
  const long vol = (long)d * h * w;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < vol; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int x = (int)(r % w); r /= w;
    const int y = (int)(r % h); r /= h;
    const int z = (int)r;
    const float cz = (float)z + (dz ? dz[i] : 0.f), cy = (float)y + dy[i], cx = (float)x + dx[i];
    if (ORDER == 0) {
      const long iz = reflect_idx((long)floorf(cz + 0.5f), d), iy = reflect_idx((long)floorf(cy + 0.5f), h),
                 ix = reflect_idx((long)floorf(cx + 0.5f), w);
      for (int ch = 0; ch < c; ++ch) dst[ch * vol + i] = src[ch * vol + (iz * h + iy) * w + ix];
    } else {
      const float fz = floorf(cz), fy = floorf(cy), fx = floorf(cx);
      const float tz = cz - fz, ty = cy - fy, tx = cx - fx;
      const long z0 = reflect_idx((long)fz, d), z1 = reflect_idx((long)fz + 1, d);
      const long y0 = reflect_idx((long)fy, h), y1 = reflect_idx((long)fy + 1, h);
      const long x0 = reflect_idx((long)fx, w), x1 = reflect_idx((long)fx + 1, w);
      for (int ch = 0; ch < c; ++ch) {
        const float* s = src + ch * vol;
        const float v000 = s[(z0 * h + y0) * w + x0], v001 = s[(z0 * h + y0) * w + x1];
        const float v010 = s[(z0 * h + y1) * w + x0], v011 = s[(z0 * h + y1) * w + x1];
        const float v100 = s[(z1 * h + y0) * w + x0], v101 = s[(z1 * h + y0) * w + x1];
        const float v110 = s[(z1 * h + y1) * w + x0], v111 = s[(z1 * h + y1) * w + x1];
        const float a0 = (1.f - tz) * (1.f - ty), a1 = (1.f - tz) * ty, a2 = tz * (1.f - ty), a3 = tz * ty;
        dst[ch * vol + i] = a0 * ((1.f - tx) * v000 + tx * v001) + a1 * ((1.f - tx) * v010 + tx * v011) +
                            a2 * ((1.f - tx) * v100 + tx * v101) + a3 * ((1.f - tx) * v110 + tx * v111);
      }
    }
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

int b200_aug_flip_rot90(int32_t c, int32_t d, int32_t h, int32_t w, int32_t flip_bits, int32_t rot_k,
                        const float* src, float* dst, b200_stream_t stream) {
  if (!src || !dst || src == dst || rot_k < 0 || rot_k > 3 || (flip_bits & ~7)) return B200_ERR_BAD_ARG;
  const long total = (long)c * d * h * w;
  if (total <= 0) return B200_OK;
  flip_rot90_kernel<<<aug_grid(total), 256, 0, (cudaStream_t)stream>>>(c, d, h, w, flip_bits, rot_k, src, dst);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_aug_affine_clip(int32_t c, int64_t per_channel, const float* src, const float* a, const float* b,
                         float lo, float hi, float* dst, b200_stream_t stream) {
  if (!src || !dst || !a || !b) return B200_ERR_BAD_ARG;
  const long total = (long)c * per_channel;
  if (total <= 0) return B200_OK;
  affine_clip_kernel<<<aug_grid(total), 256, 0, (cudaStream_t)stream>>>(c, (long)per_channel, src, a, b, lo, hi, dst);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_aug_channel_stats(int32_t c, int64_t per_channel, const float* src, double* stats, b200_stream_t stream) {
  if (!src || !stats || c > 65535) return B200_ERR_BAD_ARG;
  if ((long)c * per_channel <= 0) return B200_OK;
  long blocks = (per_channel + 256 * 8 - 1) / (256 * 8);
  const long cap = ((long)sm_count() * 8 + c - 1) / c;
  blocks = blocks < 1 ? 1 : (blocks < cap ? blocks : cap);
  channel_stats_kernel<<<dim3((unsigned)blocks, (unsigned)c), 256, 0, (cudaStream_t)stream>>>((long)per_channel, src,
                                                                                              stats);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_aug_gaussian_axis(int32_t n0, int32_t n1, int32_t n2, int32_t axis, float sigma, const float* src,
                           float* dst, b200_stream_t stream) {
  if (!src || !dst || src == dst || axis < 0 || axis > 2 || !(sigma > 0.f)) return B200_ERR_BAD_ARG;
  const long total = (long)n0 * n1 * n2;
  if (total <= 0) return B200_OK;
  const int radius = (int)(4.0f * sigma + 0.5f);
  double norm = 0.0;
  for (int k = -radius; k <= radius; ++k) norm += exp(-0.5 * (double)k * k / ((double)sigma * sigma));
  gaussian_axis_kernel<<<aug_grid(total), 256, 0, (cudaStream_t)stream>>>(n0, n1, n2, axis, sigma, radius,
                                                                          (float)(1.0 / norm), src, dst);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_aug_map_coordinates(int32_t c, int32_t d, int32_t h, int32_t w, const float* src, const float* dz,
                             const float* dy, const float* dx, int32_t order, float* dst, b200_stream_t stream) {
  if (!src || !dst || src == dst || !dy || !dx || (order != 0 && order != 1)) return B200_ERR_BAD_ARG;
  const long vol = (long)d * h * w;
  if (vol * c <= 0) return B200_OK;
  if (order == 0)
    map_coordinates_kernel<0><<<aug_grid(vol), 256, 0, (cudaStream_t)stream>>>(c, d, h, w, src, dz, dy, dx, dst);
  else
    map_coordinates_kernel<1><<<aug_grid(vol), 256, 0, (cudaStream_t)stream>>>(c, d, h, w, src, dz, dy, dx, dst);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
