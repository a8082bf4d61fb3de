// MS-SSIM term of the reference's SegmentationLoss (model/unet2d/loss.py:21-28: `1 - MS_SSIM(data_range=1,
// channel=1)(sigmoid(logits), targets)`, third-party pytorch-msssim==1.0.0, requirements.txt:127), forward and
// backward, single-channel fp32.
//
// Algorithm (restated in oracle/msssim.py): 5 scales; per scale an 11-tap Gaussian (sigma 1.5), VALID separable
// filtering of X, Y, X^2, Y^2, XY -> cs = (2 s12 + C2) / (s1 + s2 + C2), ssim = lum * cs with
// lum = (2 mu1 mu2 + C1) / (mu1^2 + mu2^2 + C1); per-sample means of cs (scales 0-3) and ssim (scale 4); between
// scales a 2x2 average pool with padding = size % 2 (zeros counted).  The product over scales and the batch mean
// are five numbers per sample and stay in the host-side formula (as the other losses' six sums do).
//
// The data is tiny next to the network (N x H x W floats, L2-resident), so the kernels are written for simplicity:
// a 16 x 16 output tile per block, both filter passes in shared memory.  Backward is the exact adjoint: the per-
// position coefficients a = dL/dmu1, b = dL/dE[x^2], c = dL/dE[xy] are recomputed on the tile's 26 x 26 support,
// spread back with the same (symmetric) window — dX = G^T a + 2 X G^T b + Y G^T c — and the coarser scale's
// gradient is folded in through the average pool (x 1/4) in the same kernel; scale 0 also applies the sigmoid's
// derivative and writes d(loss)/d(logits).
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {

#define MS_LEVELS 5
#define MS_T 16           // output tile edge
#define MS_IN 26          // MS_T + 10: input support of a forward tile / coefficient support of a backward tile
#define MS_IN2 36         // MS_T + 20: input support of a backward tile

// exp(-(k-5)^2 / (2 * 1.5^2)) / sum, k = 0..10  (pytorch_msssim._fspecial_gauss_1d(11, 1.5))
__constant__ float ms_g[11] = {0.00102838f, 0.00759876f, 0.03600077f, 0.10936069f, 0.21300553f, 0.26601171f,
                               0.21300553f, 0.10936069f, 0.03600077f, 0.00759876f, 0.00102838f};

__device__ __forceinline__ double ms_block_sum(double v, double* sh) {
  v = warp_sum_d(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum_d(r);
  }
  return r;  // valid in thread 0
}

// scale 0: X = sigmoid(logits), Y = targets
__global__ void __launch_bounds__(256) ms_sigmoid_kernel(long total, const float* __restrict__ logits,
                                                         const float* __restrict__ targets, float* __restrict__ x,
                                                         float* __restrict__ y) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    x[i] = 1.f / (1.f + expf(-logits[i]));
    y[i] = targets[i];
  }
}
// F.avg_pool2d(kernel 2, stride 2, padding = size % 2), count_include_pad: always / 4
__global__ void __launch_bounds__(256) ms_pool_kernel(int n, int h, int w, int oh, int ow, int ph, int pw,
                                                      const float* __restrict__ xi, const float* __restrict__ yi,
                                                      float* __restrict__ xo, float* __restrict__ yo) {
  const long total = (long)n * oh * ow;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
    const long b = i / ((long)ow * oh);
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int iy = 2 * oy - ph + e, ix = 2 * ox - pw + g;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
          sx += xi[(b * h + iy) * w + ix];
          sy += yi[(b * h + iy) * w + ix];
        }
      }
    xo[i] = 0.25f * sx;
    yo[i] = 0.25f * sy;
  }
}

struct MsStats { float mu1, mu2, e11, e22, e12; };

__device__ __forceinline__ void ms_maps(const MsStats& m, float c1, float c2, float& cs, float& lum, float& B,
                                        float& Q) {
  const float s1 = m.e11 - m.mu1 * m.mu1, s2 = m.e22 - m.mu2 * m.mu2, s12 = m.e12 - m.mu1 * m.mu2;
  B = s1 + s2 + c2;
  Q = m.mu1 * m.mu1 + m.mu2 * m.mu2 + c1;
  cs = (2.f * s12 + c2) / B;
  lum = (2.f * m.mu1 * m.mu2 + c1) / Q;
}

// grid (tiles_x, tiles_y, n): sums[n] += sum over the tile's valid positions of cs (or ssim)
__global__ void __launch_bounds__(256)
    ms_fwd_kernel(int h, int w, const float* __restrict__ X, const float* __restrict__ Y, double* __restrict__ sums,
                  int want_ssim, float c1, float c2) {
  __shared__ float sx[MS_IN][MS_IN + 1], sy[MS_IN][MS_IN + 1];
  __shared__ float hp[5][MS_IN][MS_T];
  __shared__ double red[8];
  const int vh = h - 10, vw = w - 10;
  const int ty0 = blockIdx.y * MS_T, tx0 = blockIdx.x * MS_T;
  const long base = (long)blockIdx.z * h * w;
  for (int i = threadIdx.x; i < MS_IN * MS_IN; i += 256) {
    const int r = i / MS_IN, c = i % MS_IN;
    const int iy = ty0 + r, ix = tx0 + c;
    const bool in = iy < h && ix < w;
    sx[r][c] = in ? X[base + (long)iy * w + ix] : 0.f;
    sy[r][c] = in ? Y[base + (long)iy * w + ix] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MS_IN * MS_T; i += 256) {
    const int r = i / MS_T, c = i % MS_T;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = ms_g[k], x = sx[r][c + k], y = sy[r][c + k];
      m1 = fmaf(g, x, m1); m2 = fmaf(g, y, m2);
      e11 = fmaf(g, x * x, e11); e22 = fmaf(g, y * y, e22); e12 = fmaf(g, x * y, e12);
    }
    hp[0][r][c] = m1; hp[1][r][c] = m2; hp[2][r][c] = e11; hp[3][r][c] = e22; hp[4][r][c] = e12;
  }
  __syncthreads();
  const int r = threadIdx.x / MS_T, c = threadIdx.x % MS_T;
  double val = 0.0;
  if (ty0 + r < vh && tx0 + c < vw) {
    MsStats m = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = ms_g[k];
      m.mu1 = fmaf(g, hp[0][r + k][c], m.mu1); m.mu2 = fmaf(g, hp[1][r + k][c], m.mu2);
      m.e11 = fmaf(g, hp[2][r + k][c], m.e11); m.e22 = fmaf(g, hp[3][r + k][c], m.e22);
      m.e12 = fmaf(g, hp[4][r + k][c], m.e12);
    }
    float cs, lum, B, Q;
    ms_maps(m, c1, c2, cs, lum, B, Q);
    val = want_ssim ? (double)(lum * cs) : (double)cs;
  }
  const double tot = ms_block_sum(val, red);
  if (threadIdx.x == 0) atomicAdd(&sums[blockIdx.z], tot);
}

// grid (tiles_x, tiles_y, n) over the FULL image.  gmean[n] = d(loss)/d(mean of this scale's map for sample n);
// parent (may be NULL) = gradient w.r.t. the next coarser scale's X (ph x pw, pooled with padding pad_h / pad_w).
// final_sigmoid: X is sigmoid(logits); write d/d(logits) = dX * X * (1 - X).
__global__ void __launch_bounds__(256)
    ms_bwd_kernel(int h, int w, const float* __restrict__ X, const float* __restrict__ Y,
                  const float* __restrict__ gmean, int want_ssim, float c1, float c2,
                  const float* __restrict__ parent, int ph, int pw, int pad_h, int pad_w, int final_sigmoid,
                  float* __restrict__ dX) {
  __shared__ float sx[MS_IN2][MS_IN2 + 1], sy[MS_IN2][MS_IN2 + 1];
  __shared__ float hp[5][MS_IN2][MS_IN];
  __shared__ float abc[3][MS_IN][MS_IN + 1];
  __shared__ float hq[3][MS_IN][MS_T];
  const int vh = h - 10, vw = w - 10;
  const int ty0 = blockIdx.y * MS_T, tx0 = blockIdx.x * MS_T;
  const long base = (long)blockIdx.z * h * w;
  const float gm = gmean[blockIdx.z] / (float)((long)vh * vw);
  // inputs rows ty0 - 10 .. ty0 + 25
  for (int i = threadIdx.x; i < MS_IN2 * MS_IN2; i += 256) {
    const int r = i / MS_IN2, c = i % MS_IN2;
    const int iy = ty0 - 10 + r, ix = tx0 - 10 + c;
    const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
    sx[r][c] = in ? X[base + (long)iy * w + ix] : 0.f;
    sy[r][c] = in ? Y[base + (long)iy * w + ix] : 0.f;
  }
  __syncthreads();
  // horizontal pass at valid columns vx = tx0 - 10 + c, c < 26
  for (int i = threadIdx.x; i < MS_IN2 * MS_IN; i += 256) {
    const int r = i / MS_IN, c = i % MS_IN;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = ms_g[k], x = sx[r][c + k], y = sy[r][c + k];
      m1 = fmaf(g, x, m1); m2 = fmaf(g, y, m2);
      e11 = fmaf(g, x * x, e11); e22 = fmaf(g, y * y, e22); e12 = fmaf(g, x * y, e12);
    }
    hp[0][r][c] = m1; hp[1][r][c] = m2; hp[2][r][c] = e11; hp[3][r][c] = e22; hp[4][r][c] = e12;
  }
  __syncthreads();
  // vertical pass -> coefficients at valid positions v = (ty0 - 10 + r, tx0 - 10 + c)
  for (int i = threadIdx.x; i < MS_IN * MS_IN; i += 256) {
    const int r = i / MS_IN, c = i % MS_IN;
    const int vy = ty0 - 10 + r, vx = tx0 - 10 + c;
    float a = 0.f, b = 0.f, cc = 0.f;
    if (vy >= 0 && vy < vh && vx >= 0 && vx < vw) {
      MsStats m = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = ms_g[k];
        m.mu1 = fmaf(g, hp[0][r + k][c], m.mu1); m.mu2 = fmaf(g, hp[1][r + k][c], m.mu2);
        m.e11 = fmaf(g, hp[2][r + k][c], m.e11); m.e22 = fmaf(g, hp[3][r + k][c], m.e22);
        m.e12 = fmaf(g, hp[4][r + k][c], m.e12);
      }
      float cs, lum, B, Q;
      ms_maps(m, c1, c2, cs, lum, B, Q);
      const float dcs_mu1 = 2.f * (cs * m.mu1 - m.mu2) / B, dcs_e11 = -cs / B, dcs_e12 = 2.f / B;
      if (want_ssim) {
        const float dlum_mu1 = 2.f * (m.mu2 - lum * m.mu1) / Q;
        a = gm * (lum * dcs_mu1 + cs * dlum_mu1);
        b = gm * lum * dcs_e11;
        cc = gm * lum * dcs_e12;
      } else {
        a = gm * dcs_mu1; b = gm * dcs_e11; cc = gm * dcs_e12;
      }
    }
    abc[0][r][c] = a; abc[1][r][c] = b; abc[2][r][c] = cc;
  }
  __syncthreads();
  // adjoint, horizontal: hq[q][r][c] = sum_k g[k] * abc[q][r][c + 10 - k]   (output column tx0 + c)
  for (int i = threadIdx.x; i < MS_IN * MS_T; i += 256) {
    const int r = i / MS_T, c = i % MS_T;
    float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = ms_g[k];
      a = fmaf(g, abc[0][r][c + 10 - k], a); b = fmaf(g, abc[1][r][c + 10 - k], b);
      cc = fmaf(g, abc[2][r][c + 10 - k], cc);
    }
    hq[0][r][c] = a; hq[1][r][c] = b; hq[2][r][c] = cc;
  }
  __syncthreads();
  const int r = threadIdx.x / MS_T, c = threadIdx.x % MS_T;
  const int py = ty0 + r, px = tx0 + c;
  if (py < h && px < w) {
    float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = ms_g[k];
      a = fmaf(g, hq[0][r + 10 - k][c], a); b = fmaf(g, hq[1][r + 10 - k][c], b);
      cc = fmaf(g, hq[2][r + 10 - k][c], cc);
    }
    const float x = sx[r + 10][c + 10], y = sy[r + 10][c + 10];
    float d = a + 2.f * x * b + y * cc;
    if (parent) d += 0.25f * parent[((long)blockIdx.z * ph + ((py + pad_h) >> 1)) * pw + ((px + pad_w) >> 1)];
    if (final_sigmoid) d *= x * (1.f - x);
    dX[base + (long)py * w + px] = d;
  }
}

struct MsPyramid {
  int h[MS_LEVELS], w[MS_LEVELS];
  long off[MS_LEVELS + 1];   // float offsets of each scale in a packed pyramid buffer (16-byte aligned)
};
static bool ms_pyramid(int n, int h, int w, MsPyramid* p) {
  if (n < 0 || (h < w ? h : w) <= 160) return false;    // pytorch_msssim asserts smaller_side > (11 - 1) * 2^4
  long off = 0;
  for (int l = 0; l < MS_LEVELS; ++l) {
    p->h[l] = h; p->w[l] = w; p->off[l] = off;
    off += ((long)n * h * w + 3) / 4 * 4;
    h = (h + (h & 1)) / 2;
    w = (w + (w & 1)) / 2;
  }
  p->off[MS_LEVELS] = off;
  return true;
}
static inline unsigned ms_grid1d(long total) {
  const long b = (total + 255) / 256;
  return (unsigned)(b < 148L * 8 ? (b > 0 ? b : 1) : 148L * 8);
}

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_msssim_pyramid_floats(int32_t n, int32_t h, int32_t w) {
  MsPyramid p;
  return ms_pyramid(n, h, w, &p) ? (int64_t)p.off[MS_LEVELS] : (int64_t)B200_ERR_UNSUPPORTED;
}

int b200_msssim_fwd(int32_t n, int32_t h, int32_t w, const float* logits, const float* targets, float* pyr_x,
                    float* pyr_y, double* sums, float data_range, b200_stream_t stream) {
  MsPyramid p;
  if (!ms_pyramid(n, h, w, &p) || !logits || !targets || !pyr_x || !pyr_y || !sums) return B200_ERR_UNSUPPORTED;
  if (n == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * MS_LEVELS * n, st);
  if (e != cudaSuccess) return (int)e;
  const long total0 = (long)n * h * w;
  ms_sigmoid_kernel<<<ms_grid1d(total0), 256, 0, st>>>(total0, logits, targets, pyr_x, pyr_y);
  B200_LAUNCH_CHECK();
  for (int l = 0; l < MS_LEVELS; ++l) {
    if (l > 0) {
      const long tot = (long)n * p.h[l] * p.w[l];
      ms_pool_kernel<<<ms_grid1d(tot), 256, 0, st>>>(n, p.h[l - 1], p.w[l - 1], p.h[l], p.w[l], p.h[l - 1] & 1,
                                                     p.w[l - 1] & 1, pyr_x + p.off[l - 1], pyr_y + p.off[l - 1],
                                                     pyr_x + p.off[l], pyr_y + p.off[l]);
      B200_LAUNCH_CHECK();
    }
    const dim3 grid((p.w[l] - 10 + MS_T - 1) / MS_T, (p.h[l] - 10 + MS_T - 1) / MS_T, n);
    ms_fwd_kernel<<<grid, 256, 0, st>>>(p.h[l], p.w[l], pyr_x + p.off[l], pyr_y + p.off[l], sums + (long)l * n,
                                        l == MS_LEVELS - 1, c1, c2);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

int b200_msssim_bwd(int32_t n, int32_t h, int32_t w, const float* pyr_x, const float* pyr_y, const float* gmeans,
                    float* pyr_dx, float* dlogits, float data_range, b200_stream_t stream) {
  MsPyramid p;
  if (!ms_pyramid(n, h, w, &p) || !pyr_x || !pyr_y || !gmeans || !pyr_dx || !dlogits) return B200_ERR_UNSUPPORTED;
  if (n == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  for (int l = MS_LEVELS - 1; l >= 0; --l) {
    const bool last = l == MS_LEVELS - 1;
    const dim3 grid((p.w[l] + MS_T - 1) / MS_T, (p.h[l] + MS_T - 1) / MS_T, n);
    ms_bwd_kernel<<<grid, 256, 0, st>>>(p.h[l], p.w[l], pyr_x + p.off[l], pyr_y + p.off[l], gmeans + (long)l * n,
                                        last, c1, c2, last ? nullptr : pyr_dx + p.off[l + 1], last ? 0 : p.h[l + 1],
                                        last ? 0 : p.w[l + 1], p.h[l] & 1, p.w[l] & 1, l == 0,
                                        l == 0 ? dlogits : pyr_dx + p.off[l]);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

}  // extern "C"
