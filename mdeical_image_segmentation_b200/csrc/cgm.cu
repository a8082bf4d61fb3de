// Class-guided module of UNet_3Plus_DeepSup_CGM (model/unet2d/unet.py:795-1153).
//   cls  = Sequential(Dropout(0.5), Conv2d(1024, 2, 1), AdaptiveMaxPool2d(1), Sigmoid())(hd5)      (:999-1004)
//   gate = cls.argmax(dim=1).float()                                                            (:1036-1038)
//   d_i  = sigmoid(einsum("ijk,ij->ijk", d_i, gate))                                        (:1012-1017, :1147-1153)
// The gate is a 0/1 number per sample (arg-max: no gradient reaches the cls convolution, as in the reference),
// so the whole module is one small reduction over hd5 plus an element-wise pass over the five logit maps.
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {

// one block per sample: warps walk positions, lanes own 8-channel vectors; max over positions of the two 1x1-conv
// outputs (bias added after the max, sigmoid is monotone), then gate = (sigmoid(m1) > sigmoid(m0)) in fp32 so that
// saturation ties resolve to class 0 exactly as torch.argmax does.
__global__ void __launch_bounds__(256)
    cgm_gate_kernel(long npos, int c, const bf16* __restrict__ x, int ctot, int coff, const float* __restrict__ w,
                    const float* __restrict__ bias, const uint8_t* __restrict__ keep, float keep_scale,
                    float* __restrict__ gate, float* __restrict__ cls) {
  __shared__ float sm0[8], sm1[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long b = blockIdx.x;
  float m0 = -INFINITY, m1 = -INFINITY;
  for (long pos = warp; pos < npos; pos += nwarps) {
    const bf16* row = x + (b * npos + pos) * ctot + coff;
    const uint8_t* krow = keep ? keep + (b * npos + pos) * c : nullptr;
    float a0 = 0.f, a1 = 0.f;
    for (int ch = lane * 8; ch < c; ch += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + ch);
      const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
      float f[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (krow) f[q] = krow[ch + q] ? f[q] * keep_scale : 0.f;
        a0 = fmaf(f[q], w[ch + q], a0);
        a1 = fmaf(f[q], w[c + ch + q], a1);
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    m0 = fmaxf(m0, a0);
    m1 = fmaxf(m1, a1);
  }
  if (lane == 0) { sm0[warp] = m0; sm1[warp] = m1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nwarps; ++i) { m0 = fmaxf(m0, sm0[i]); m1 = fmaxf(m1, sm1[i]); }
    const float s0 = 1.f / (1.f + expf(-(m0 + bias[0]))), s1 = 1.f / (1.f + expf(-(m1 + bias[1])));
    gate[b] = s1 > s0 ? 1.f : 0.f;
    if (cls) { cls[2 * b] = s0; cls[2 * b + 1] = s1; }
  }
}

__global__ void __launch_bounds__(256)
    gate_sigmoid_fwd_kernel(long per_sample, long total, const float* __restrict__ x, const float* __restrict__ gate,
                            float* __restrict__ y) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
    y[i] = 1.f / (1.f + expf(-x[i] * gate[i / per_sample]));
}
__global__ void __launch_bounds__(256)
    gate_sigmoid_bwd_kernel(long per_sample, long total, const float* __restrict__ y, const float* __restrict__ dy,
                            const float* __restrict__ gate, float* __restrict__ dx) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float v = y[i];
    dx[i] = dy[i] * v * (1.f - v) * gate[i / per_sample];
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_cgm_gate(int32_t n, int64_t npos, int32_t c, b200_view x, const float* w, const float* bias,
                  const uint8_t* keep, float keep_scale, float* gate, float* cls, b200_stream_t stream) {
  if (c % 8 || x.ctot % 8 || x.coff % 8 || ((uintptr_t)x.ptr % 16) || !w || !bias || !gate || npos <= 0)
    return B200_ERR_UNSUPPORTED;
  if (n <= 0) return B200_OK;
  cgm_gate_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>((long)npos, c, (const bf16*)x.ptr, x.ctot, x.coff, w,
                                                                 bias, keep, keep_scale, gate, cls);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
static inline unsigned cgm_grid(long total) {
  const long b = (total + 255) / 256;
  return (unsigned)(b < 148L * 16 ? b : 148L * 16);
}
int b200_gate_sigmoid_fwd(int32_t n, int64_t per_sample, const float* x, const float* gate, float* y,
                          b200_stream_t stream) {
  const long total = (long)n * per_sample;
  if (total <= 0) return B200_OK;
  gate_sigmoid_fwd_kernel<<<cgm_grid(total), 256, 0, (cudaStream_t)stream>>>((long)per_sample, total, x, gate, y);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_gate_sigmoid_bwd(int32_t n, int64_t per_sample, const float* y, const float* dy, const float* gate,
                          float* dx, b200_stream_t stream) {
  const long total = (long)n * per_sample;
  if (total <= 0) return B200_OK;
  gate_sigmoid_bwd_kernel<<<cgm_grid(total), 256, 0, (cudaStream_t)stream>>>((long)per_sample, total, y, dy, gate, dx);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // extern "C"
