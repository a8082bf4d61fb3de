// Sliding-window prediction on the device (model/unet3d/predictor.py:65-165 StandardPredictor and
// dataset/unet3d_dataset/utils.py:40-131, 314-361): the reference mirror-pads the volume on the host, cuts
// patches, copies every predicted patch back (`prediction.cpu().numpy()`), and accumulates / normalises in numpy.
// Here the raw volume and the prediction / normalisation maps stay in HBM:
//   gather     : one (patch + halo) window read straight from the UNPADDED volume with reflect indexing
//                (np.pad(mode='reflect'), utils.py:314-342) — no padded copy is ever materialised;
//   accumulate : prediction_map[index] += pred without its halo, normalization_mask[index] += 1
//                (predictor.py:147-150, remove_padding utils.py:345-361); one launch per patch, so overlapping
//                patches are summed in the reference's order and the result is bit-identical and deterministic;
//   finalize   : prediction_map / normalization_mask, optional arg-max over channels (predictor.py:161-165).
// All three are plain HBM-bound element-wise passes over fp32 NCDHW data (the module API's own format).
#include "common.cuh"
#include "../../include/b200unet.h"

namespace b200 {

__device__ __forceinline__ int reflect_index(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i = i < 0 ? -i : i;
  i %= period;
  return i < n ? i : period - i;
}

__global__ void __launch_bounds__(256)
    patch_gather_kernel(int c, int d, int h, int w, const float* __restrict__ vol, int oz, int oy, int ox, int hz,
                        int hy, int hx, int pd, int ph, int pw, float* __restrict__ out) {
  const long total = (long)c * pd * ph * pw;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % pw), y = (int)((i / pw) % ph), z = (int)((i / ((long)pw * ph)) % pd);
    const int ch = (int)(i / ((long)pw * ph * pd));
    // window origin (oz, oy, ox) is in un-padded coordinates of the patch WITHOUT halo; the window starts hz before
    const int sz = reflect_index(oz - hz + z, d), sy = reflect_index(oy - hy + y, h), sx = reflect_index(ox - hx + x, w);
    out[i] = vol[(((long)ch * d + sz) * h + sy) * w + sx];
  }
}

__global__ void __launch_bounds__(256)
    patch_accumulate_kernel(int c, int d, int h, int w, int pd, int ph, int pw, int hz, int hy, int hx,
                            const float* __restrict__ pred, int channel, int oz, int oy, int ox,
                            float* __restrict__ map, uint8_t* __restrict__ norm) {
  const int id = pd - 2 * hz, ih = ph - 2 * hy, iw = pw - 2 * hx;      // inner (un-padded) patch
  const long total = (long)c * id * ih * iw;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % iw), y = (int)((i / iw) % ih), z = (int)((i / ((long)iw * ih)) % id);
    const int ch = (int)(i / ((long)iw * ih * id));
    const int src_ch = channel >= 0 ? channel : ch;
    const float v = pred[(((long)src_ch * pd + z + hz) * ph + y + hy) * pw + x + hx];
    const long dst = (((long)ch * d + oz + z) * h + oy + y) * w + ox + x;
    map[dst] += v;
    norm[dst] += 1;
  }
}

__global__ void __launch_bounds__(256)
    prediction_finalize_kernel(int c, long nvox, float* __restrict__ map, const uint8_t* __restrict__ norm,
                               int32_t* __restrict__ seg) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (long)gridDim.x * blockDim.x) {
    float best = 0.f;
    int arg = 0;
    for (int ch = 0; ch < c; ++ch) {
      const float v = map[ch * nvox + i] / (float)norm[ch * nvox + i];
      map[ch * nvox + i] = v;
      if (ch == 0 || v > best) { best = v; arg = ch; }       // first maximum, as np.argmax
    }
    if (seg) seg[i] = arg;
  }
}

static inline unsigned pred_grid(long total) {
  const long b = (total + 255) / 256;
  return (unsigned)(b < 148L * 16 ? (b > 0 ? b : 1) : 148L * 16);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_patch_gather(int32_t c, int32_t d, int32_t h, int32_t w, const float* vol, int32_t oz, int32_t oy,
                      int32_t ox, int32_t hz, int32_t hy, int32_t hx, int32_t pd, int32_t ph, int32_t pw, float* out,
                      b200_stream_t stream) {
  if (!vol || !out || c <= 0 || d <= 0 || h <= 0 || w <= 0 || pd <= 0 || ph <= 0 || pw <= 0 || hz < 0 || hy < 0 ||
      hx < 0 || hz >= d || hy >= h || hx >= w)       // np.pad(reflect) needs pad < size
    return B200_ERR_BAD_ARG;
  const long total = (long)c * pd * ph * pw;
  patch_gather_kernel<<<pred_grid(total), 256, 0, (cudaStream_t)stream>>>(c, d, h, w, vol, oz, oy, ox, hz, hy, hx, pd,
                                                                         ph, pw, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_patch_accumulate(int32_t c, int32_t d, int32_t h, int32_t w, int32_t pd, int32_t ph, int32_t pw, int32_t hz,
                          int32_t hy, int32_t hx, const float* pred, int32_t c_pred, int32_t channel, int32_t oz,
                          int32_t oy, int32_t ox, float* map, uint8_t* norm, b200_stream_t stream) {
  const int id = pd - 2 * hz, ih = ph - 2 * hy, iw = pw - 2 * hx;
  if (!pred || !map || !norm || c <= 0 || id <= 0 || ih <= 0 || iw <= 0 || hz < 0 || hy < 0 || hx < 0 ||
      channel >= c_pred || (channel < 0 && c != c_pred) || (channel >= 0 && c != 1) || oz < 0 || oy < 0 || ox < 0 ||
      oz + id > d || oy + ih > h || ox + iw > w)
    return B200_ERR_BAD_ARG;
  const long total = (long)c * id * ih * iw;
  patch_accumulate_kernel<<<pred_grid(total), 256, 0, (cudaStream_t)stream>>>(c, d, h, w, pd, ph, pw, hz, hy, hx, pred,
                                                                             channel, oz, oy, ox, map, norm);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int b200_prediction_finalize(int32_t c, int64_t nvox, float* map, const uint8_t* norm, int32_t* seg,
                             b200_stream_t stream) {
  if (!map || !norm || c <= 0) return B200_ERR_BAD_ARG;
  if (nvox <= 0) return B200_OK;
  prediction_finalize_kernel<<<pred_grid((long)nvox), 256, 0, (cudaStream_t)stream>>>(c, (long)nvox, map, norm, seg);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // extern "C"
