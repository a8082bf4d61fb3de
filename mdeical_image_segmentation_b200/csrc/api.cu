// extern "C" entry points for the convolution family: weight repacking and the dispatch between the
// tcgen05 implicit-GEMM kernels (conv_tc.cu) and the CUDA-core kernels (conv_direct.cu).
#include "../../include/b200unet.h"
#include "conv_tc.cuh"

namespace b200 {
extern thread_local int g_halo_mode;
extern thread_local int g_wgrad_halo_mode;
extern thread_local int g_halo_rmax;
extern thread_local int g_halo_resident;
extern thread_local int g_gemm_resident;
extern thread_local int g_epi_dbg;
extern thread_local int g_epi_warps;
extern thread_local int g_wgrad_wide;
extern thread_local int g_halo_bn;
extern thread_local int g_wgrad_d3;
extern thread_local int g_wgrad_d3_epi;
extern thread_local int g_roll_mode;
extern thread_local int g_roll_seg;
int direct_gemm_conv(const TcGemmProblem& p, cudaStream_t stream);
int direct_wgrad(const TcWgradProblem& p, cudaStream_t stream);
static thread_local int g_last_impl = 0;

// w (c_out, c_in, taps) fp32 -> fwd [tap][c_out][c_in], bwd [ntaps-1-tap][c_in][c_out]
__global__ void conv_repack_kernel(int c_out, int c_in, int ntaps, const float* w, bf16* fwd, bf16* bwd) {
  const long total = (long)c_out * c_in * ntaps;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % ntaps);
    const int ci = (int)((i / ntaps) % c_in);
    const int co = (int)(i / ((long)ntaps * c_in));
    const bf16 v = __float2bfloat16(w[i]);
    if (fwd) fwd[((long)tap * c_out + co) * c_in + ci] = v;
    if (bwd) bwd[((long)(ntaps - 1 - tap) * c_in + ci) * c_out + co] = v;
  }
}
// w (c_in, c_out, 2, 2) fp32 -> fwd [(ij)*c_out + co][c_in], bwd [ij][c_in][c_out]
__global__ void convt_repack_kernel(int c_in, int c_out, const float* w, bf16* fwd, bf16* bwd) {
  const long total = (long)c_in * c_out * 4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ij = (int)(i % 4);
    const int co = (int)((i / 4) % c_out);
    const int ci = (int)(i / (4L * c_out));
    const bf16 v = __float2bfloat16(w[i]);
    if (fwd) fwd[((long)ij * c_out + co) * c_in + ci] = v;
    if (bwd) bwd[((long)ij * c_in + ci) * c_out + co] = v;
  }
}
// w (c_in, c_out, taps) fp32 -> fwd [tap][c_out][c_in], bwd [tap][c_in][c_out]
__global__ void deconv_repack_kernel(int c_in, int c_out, int ntaps, const float* w, bf16* fwd, bf16* bwd) {
  const long total = (long)c_in * c_out * ntaps;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % ntaps);
    const int co = (int)((i / ntaps) % c_out);
    const int ci = (int)(i / ((long)ntaps * c_out));
    const bf16 v = __float2bfloat16(w[i]);
    if (fwd) fwd[((long)tap * c_out + co) * c_in + ci] = v;
    if (bwd) bwd[((long)tap * c_in + ci) * c_out + co] = v;
  }
}

// Tap lists of the stride-2 / kernel-3 / padding-1 transposed convolution, per axis:
//   output z = 2m + p:  p = 0 <- (k = 1, x[m]);   p = 1 <- (k = 2, x[m]) and (k = 0, x[m + 1])
//   input  i:           k = 1 -> plane p = 0 at m = i;  k = 2 -> p = 1 at m = i;  k = 0 -> p = 1 at m = i - 1
static void deconv_fwd_table(int kd, int cls, TapTable* t) {
  const int p[3] = {(cls >> 2) & 1, (cls >> 1) & 1, cls & 1};
  int kopt[3][2], sopt[3][2], nopt[3];
  for (int a = 0; a < 3; ++a) {
    if (a == 0 && kd == 1) { nopt[a] = 1; kopt[a][0] = 0; sopt[a][0] = 0; continue; }
    if (p[a] == 0) { nopt[a] = 1; kopt[a][0] = 1; sopt[a][0] = 0; }
    else { nopt[a] = 2; kopt[a][0] = 2; sopt[a][0] = 0; kopt[a][1] = 0; sopt[a][1] = 1; }
  }
  t->ntaps = 0;
  for (int i = 0; i < nopt[0]; ++i)
    for (int j = 0; j < nopt[1]; ++j)
      for (int k = 0; k < nopt[2]; ++k) {
        const int q = t->ntaps++;
        t->dd[q] = (signed char)sopt[0][i]; t->dh[q] = (signed char)sopt[1][j]; t->dw[q] = (signed char)sopt[2][k];
        t->dn[q] = 0;
        t->widx[q] = (signed char)(kd == 1 ? kopt[1][j] * 3 + kopt[2][k]
                                           : (kopt[0][i] * 3 + kopt[1][j]) * 3 + kopt[2][k]);
      }
}
static void deconv_bwd_table(int kd, int n, TapTable* t) {
  static const int pk[3] = {1, 0, 1};    // parity plane read by kernel index k
  static const int sk[3] = {-1, 0, 0};   // position shift for kernel index k
  t->ntaps = 0;
  for (int a = 0; a < (kd == 1 ? 1 : 3); ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) {
        const int q = t->ntaps++;
        const int pd = kd == 1 ? 0 : pk[a];
        t->dd[q] = (signed char)(kd == 1 ? 0 : sk[a]); t->dh[q] = (signed char)sk[b]; t->dw[q] = (signed char)sk[c];
        t->dn[q] = ((pd * 2 + pk[b]) * 2 + pk[c]) * n;
        t->widx[q] = (signed char)q;
      }
}

static int check_conv(const b200_conv_desc* d) {
  if (!d) return B200_ERR_BAD_ARG;
  if (d->n <= 0 || d->d <= 0 || d->h <= 0 || d->w <= 0 || d->c_in <= 0 || d->c_out <= 0)
    return B200_ERR_BAD_ARG;
  const int ks[3] = {d->kd, d->kh, d->kw};
  for (int k : ks)
    if (k != 1 && k != 3) return B200_ERR_UNSUPPORTED;
  if (d->in_coff < 0 || d->in_coff + d->c_in > d->in_ctot) return B200_ERR_BAD_ARG;
  if (d->out_coff < 0 || d->out_coff + d->c_out > d->out_ctot) return B200_ERR_BAD_ARG;
  return B200_OK;
}
static unsigned grid_for(long total) {
  long b = (total + 255) / 256;
  return (unsigned)(b > 148L * 8 ? 148L * 8 : (b < 1 ? 1 : b));
}

// Stand-alone column statistics of a stored tensor (what the tcgen05 epilogues fuse): one block row per
// sample, a thread per channel.  Only the CUDA-core fallback path uses it.
__global__ void colstats_kernel(long spatial, int c, const bf16* out, int octot, int ocoff, const bf16* dot,
                                int dctot, int dcoff, double* stats, int sum_only) {
  const int b = blockIdx.y;
  const long chunk = (spatial + gridDim.x - 1) / gridDim.x;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = pbeg + chunk < spatial ? pbeg + chunk : spatial;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (long pos = pbeg; pos < pend; ++pos) {
      const long gp = (long)b * spatial + pos;
      const float v = __bfloat162float(out[gp * octot + ocoff + ch]);
      const float d = dot ? __bfloat162float(dot[gp * dctot + dcoff + ch]) : v;
      s0 += v;
      s1 = fmaf(v, d, s1);
    }
    if (pend > pbeg) {
      atomicAdd(stats + ((long)b * 2 + 0) * c + ch, (double)s0);
      if (!sum_only) atomicAdd(stats + ((long)b * 2 + 1) * c + ch, (double)s1);
    }
  }
}

static int run_gemm(const TcGemmProblem& p, int impl, cudaStream_t s) {
  const bool tc_ok = tc_gemm_supported(p);
  if (impl == B200_IMPL_TCGEN05 && !tc_ok) return B200_ERR_UNSUPPORTED;
  if (impl == B200_IMPL_DIRECT || !tc_ok) {
    g_last_impl = B200_IMPL_DIRECT;
    int rc = direct_gemm_conv(p, s);
    if (rc || !p.stats) return rc;
    if (p.out_mode != TC_OUT_CONV) return B200_ERR_UNSUPPORTED;
    const long spatial = (long)p.pd * p.ph * p.pw;
    long gx = spatial / 256;
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    colstats_kernel<<<dim3((unsigned)gx, (unsigned)p.pn), 128, 0, s>>>(
        spatial, p.ncols, p.out, p.out_ctot, p.out_coff, p.dot, p.dot_ctot, p.dot_coff, p.stats,
        p.stats_kind == B200_STATS_SUM && !p.dot);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  g_last_impl = B200_IMPL_TCGEN05;
  return tc_gemm_conv(p, s);
}
static int run_wgrad(const TcWgradProblem& p, int impl, cudaStream_t s) {
  const bool tc_ok = tc_wgrad_supported(p);
  if (impl == B200_IMPL_TCGEN05 && !tc_ok) return B200_ERR_UNSUPPORTED;
  if (impl == B200_IMPL_DIRECT || !tc_ok) {
    g_last_impl = B200_IMPL_DIRECT;
    return direct_wgrad(p, s);
  }
  g_last_impl = B200_IMPL_TCGEN05;
  return tc_wgrad(p, s);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_version(void) { return 100; }

const char* b200_strerror(int code) {
  switch (code) {
    case B200_OK: return "success";
    case B200_ERR_BAD_ARG: return "b200: bad argument";
    case B200_ERR_UNSUPPORTED: return "b200: unsupported shape, alignment or kernel size";
    case B200_ERR_WORKSPACE: return "b200: workspace too small";
    case B200_ERR_ARCH: return "b200: device is not sm_100";
    case B200_ERR_DRIVER: return "b200: cuTensorMapEncodeTiled unavailable or failed";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "b200: unknown error";
}

int b200_last_impl(void) { return g_last_impl; }

int b200_set_tuning(int key, int value) {
  if (key == 0) { b200::g_halo_mode = value; return B200_OK; }
  if (key == 1) { b200::g_wgrad_halo_mode = value; return B200_OK; }
  if (key == 2) { b200::g_halo_rmax = value; return B200_OK; }
  if (key == 4) { b200::g_epi_dbg = value; return B200_OK; }
  if (key == 5) { b200::g_epi_warps = value; return B200_OK; }
  if (key == 3) { b200::g_halo_resident = value; b200::g_gemm_resident = value; return B200_OK; }
  if (key == 6) { b200::g_wgrad_wide = value; return B200_OK; }
  if (key == 7) { b200::g_halo_bn = value; return B200_OK; }
  if (key == 8) { b200::g_wgrad_d3 = value; return B200_OK; }
  if (key == 9) { b200::g_wgrad_d3_epi = value; return B200_OK; }
  if (key == 10) { b200::g_roll_mode = value; return B200_OK; }
  if (key == 11) { b200::g_roll_seg = value; return B200_OK; }
  return B200_ERR_BAD_ARG;
}

int b200_conv_repack_weight(const b200_conv_desc* d, const float* w, void* wp_fwd, void* wp_bwd,
                            b200_stream_t stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  const int ntaps = d->kd * d->kh * d->kw;
  const long total = (long)d->c_out * d->c_in * ntaps;
  conv_repack_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(d->c_out, d->c_in, ntaps, w,
                                                                        (bf16*)wp_fwd, (bf16*)wp_bwd);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ---- all conv filters of a network in ONE launch -------------------------------------------------
// A U-Net step re-packs 14-23 filters (the fp32 masters change every optimizer step); one launch per
// filter cost 0.39-0.68 ms per step, mostly launch latency and 2-byte scattered stores.  Here a block
// owns a 32 (c_out) x 32 (c_in) tile of one filter: the forward pack [tap][c_out][c_in] is written with
// c_in along the warp, the backward pack [flipped tap][c_in][c_out] goes through a shared-memory
// transpose so that c_out runs along the warp — both are contiguous 64-byte rows.
namespace b200 {
static const int kRepackMax = 24;
struct RepackTable {
  b200_repack_item it[kRepackMax];
  int first_block[kRepackMax + 1];
  int count;
};
__global__ void __launch_bounds__(1024) conv_repack_multi_kernel(const __grid_constant__ RepackTable t) {
  __shared__ bf16 tile[32][33];
  int k = 0;
  while (k + 1 < t.count && (int)blockIdx.x >= t.first_block[k + 1]) ++k;
  const b200_repack_item& it = t.it[k];
  const int local = blockIdx.x - t.first_block[k];
  const int ci_tiles = (it.c_in_pad + 31) / 32;
  const int co0 = (local / ci_tiles) * 32, ci0 = (local % ci_tiles) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int co = co0 + ty, ci = ci0 + tx;
  const bool in_src = co < it.c_out && ci < it.c_in;          // padding channels read as zero
  const float* src = it.w + ((long)co * it.c_in + ci) * it.ntaps;
  bf16* fwd = (bf16*)it.wp_fwd;
  bf16* bwd = (bf16*)it.wp_bwd;
  const int co_t = co0 + tx, ci_t = ci0 + ty;                  // transposed roles for the backward pack
  for (int tap = 0; tap < it.ntaps; ++tap) {
    const bf16 v = __float2bfloat16(in_src ? src[tap] : 0.f);
    if (fwd && co < it.c_out && ci < it.c_in_pad) fwd[((long)tap * it.c_out + co) * it.c_in_pad + ci] = v;
    if (bwd) {
      tile[ty][tx] = v;
      __syncthreads();
      if (co_t < it.c_out && ci_t < it.c_in_pad)
        bwd[((long)(it.ntaps - 1 - tap) * it.c_in_pad + ci_t) * it.c_out + co_t] = tile[tx][ty];
      __syncthreads();
    }
  }
}
}  // namespace b200

int b200_conv_repack_weights(int32_t count, const b200_repack_item* items, b200_stream_t stream) {
  if (count < 0 || (count > 0 && !items)) return B200_ERR_BAD_ARG;
  for (int base = 0; base < count; base += b200::kRepackMax) {
    b200::RepackTable t{};
    t.count = count - base < b200::kRepackMax ? count - base : b200::kRepackMax;
    int blocks = 0;
    for (int i = 0; i < t.count; ++i) {
      const b200_repack_item& it = items[base + i];
      if (!it.w || it.c_out <= 0 || it.c_in <= 0 || it.c_in_pad < it.c_in || it.ntaps <= 0) return B200_ERR_BAD_ARG;
      t.it[i] = it;
      t.first_block[i] = blocks;
      blocks += ((it.c_out + 31) / 32) * ((it.c_in_pad + 31) / 32);
    }
    t.first_block[t.count] = blocks;
    if (blocks == 0) continue;
    b200::conv_repack_multi_kernel<<<blocks, 1024, 0, (cudaStream_t)stream>>>(t);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

int b200_conv_fwd(const b200_conv_desc* d, const void* x, const void* wp_fwd, const float* bias,
                  void* y, b200_stream_t stream) {
  return b200_conv_fwd_stats(d, x, wp_fwd, bias, y, nullptr, stream);
}
int b200_conv_fwd_stats(const b200_conv_desc* d, const void* x, const void* wp_fwd, const float* bias,
                        void* y, double* colstats, b200_stream_t stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  TcGemmProblem p{};
  p.a = ActView{(const bf16*)x, d->n, d->d, d->h, d->w, d->c_in, d->in_ctot, d->in_coff};
  p.wp = (const bf16*)wp_fwd;
  p.ncols = d->c_out; p.kdim = d->c_in;
  p.kd = d->kd; p.kh = d->kh; p.kw = d->kw;
  p.tap_mode = TC_TAP_SHIFT; p.out_mode = TC_OUT_CONV;
  p.pn = d->n; p.pd = d->d; p.ph = d->h; p.pw = d->w;
  p.out = (bf16*)y; p.out_ctot = d->out_ctot; p.out_coff = d->out_coff;
  p.bias = bias; p.relu = d->relu;
  p.stats = colstats; p.stats_kind = B200_STATS_SUM_SQ;
  return run_gemm(p, d->impl, (cudaStream_t)stream);
}

int b200_conv_bwd_data(const b200_conv_desc* d, const void* dy, const void* wp_bwd,
                       const void* relu_mask, void* dx, b200_stream_t stream) {
  return b200_conv_bwd_data_stats(d, dy, wp_bwd, relu_mask, dx, B200_STATS_NONE, nullptr, 0, 0, nullptr,
                                  stream);
}
int b200_conv_bwd_data_stats(const b200_conv_desc* d, const void* dy, const void* wp_bwd,
                             const void* relu_mask, void* dx, int32_t stats_kind, const void* dot,
                             int32_t dot_ctot, int32_t dot_coff, double* colstats,
                             b200_stream_t stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  TcGemmProblem p{};
  p.a = ActView{(const bf16*)dy, d->n, d->d, d->h, d->w, d->c_out, d->out_ctot, d->out_coff};
  p.wp = (const bf16*)wp_bwd;
  p.ncols = d->c_in; p.kdim = d->c_out;
  p.kd = d->kd; p.kh = d->kh; p.kw = d->kw;
  p.tap_mode = TC_TAP_SHIFT; p.out_mode = TC_OUT_CONV;
  p.pn = d->n; p.pd = d->d; p.ph = d->h; p.pw = d->w;
  p.out = (bf16*)dx; p.out_ctot = d->in_ctot; p.out_coff = d->in_coff;
  p.mask = (const bf16*)relu_mask; p.mask_ctot = d->in_ctot; p.mask_coff = d->in_coff;
  if (stats_kind < B200_STATS_NONE || stats_kind > B200_STATS_SUM_DOT) return B200_ERR_BAD_ARG;
  if ((stats_kind != B200_STATS_NONE) != (colstats != nullptr)) return B200_ERR_BAD_ARG;
  if ((stats_kind == B200_STATS_SUM_DOT) != (dot != nullptr)) return B200_ERR_BAD_ARG;
  p.stats = colstats; p.stats_kind = stats_kind;
  p.dot = (const bf16*)dot; p.dot_ctot = dot_ctot; p.dot_coff = dot_coff;
  return run_gemm(p, d->impl, (cudaStream_t)stream);
}

int b200_conv_bwd_weight(const b200_conv_desc* d, const void* x, const void* dy, float* dw,
                         b200_stream_t stream) {
  return b200_conv_bwd_weight_bias(d, x, dy, dw, nullptr, stream);
}

int b200_conv_bwd_weight_bias(const b200_conv_desc* d, const void* x, const void* dy, float* dw,
                              float* dbias, b200_stream_t stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  TcWgradProblem p{};
  p.mop = ActView{(const bf16*)dy, d->n, d->d, d->h, d->w, d->c_out, d->out_ctot, d->out_coff};
  p.nop = ActView{(const bf16*)x, d->n, d->d, d->h, d->w, d->c_in, d->in_ctot, d->in_coff};
  p.kd = d->kd; p.kh = d->kh; p.kw = d->kw;
  p.tap_mode = TC_TAP_SHIFT;
  p.pn = d->n; p.pd = d->d; p.ph = d->h; p.pw = d->w;
  p.dw = dw;
  const int ntaps = d->kd * d->kh * d->kw;
  p.s_m = (long)d->c_in * ntaps; p.s_n = ntaps;
  if (dbias) {
    // fused into the halo weight-gradient kernel when that is the kernel that runs; otherwise one
    // HBM-speed column-sum pass
    if (d->impl != B200_IMPL_DIRECT && tc_wgrad_fuses_bias(p)) {
      p.dbias = dbias;
    } else {
      b200_view v{const_cast<void*>(dy), d->out_ctot, d->out_coff};
      rc = b200_bias_grad((int64_t)d->n * d->d * d->h * d->w, d->c_out, v, dbias, stream);
      if (rc) return rc;
    }
  }
  return run_wgrad(p, d->impl, (cudaStream_t)stream);
}

// ---- ConvTranspose2d k2 s2
static int check_convt(const b200_convt_desc* d) {
  if (!d || d->n <= 0 || d->h <= 0 || d->w <= 0 || d->c_in <= 0 || d->c_out <= 0) return B200_ERR_BAD_ARG;
  if (d->in_coff < 0 || d->in_coff + d->c_in > d->in_ctot) return B200_ERR_BAD_ARG;
  if (d->out_coff < 0 || d->out_coff + d->c_out > d->out_ctot) return B200_ERR_BAD_ARG;
  return B200_OK;
}
int b200_convt2x2_repack_weight(const b200_convt_desc* d, const float* w, void* wp_fwd, void* wp_bwd,
                                b200_stream_t stream) {
  int rc = check_convt(d);
  if (rc) return rc;
  convt_repack_kernel<<<grid_for((long)d->c_in * d->c_out * 4), 256, 0, (cudaStream_t)stream>>>(
      d->c_in, d->c_out, w, (bf16*)wp_fwd, (bf16*)wp_bwd);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_convt2x2_fwd(const b200_convt_desc* d, const void* x, const void* wp_fwd, const float* bias,
                      void* y, b200_stream_t stream) {
  int rc = check_convt(d);
  if (rc) return rc;
  TcGemmProblem p{};
  p.a = ActView{(const bf16*)x, d->n, 1, d->h, d->w, d->c_in, d->in_ctot, d->in_coff};
  p.wp = (const bf16*)wp_fwd;
  p.ncols = 4 * d->c_out; p.kdim = d->c_in;
  p.kd = p.kh = p.kw = 1;
  p.tap_mode = TC_TAP_SHIFT; p.out_mode = TC_OUT_K2S2_SCATTER;
  p.pn = d->n; p.pd = 1; p.ph = d->h; p.pw = d->w;
  p.out = (bf16*)y; p.out_ctot = d->out_ctot; p.out_coff = d->out_coff; p.cout_t = d->c_out;
  p.bias = bias;
  return run_gemm(p, d->impl, (cudaStream_t)stream);
}
int b200_convt2x2_bwd_data(const b200_convt_desc* d, const void* dy, const void* wp_bwd,
                           const void* relu_mask, void* dx, b200_stream_t stream) {
  return b200_convt2x2_bwd_data_stats(d, dy, wp_bwd, relu_mask, dx, B200_STATS_NONE, nullptr, stream);
}
int b200_convt2x2_bwd_data_stats(const b200_convt_desc* d, const void* dy, const void* wp_bwd,
                                 const void* relu_mask, void* dx, int32_t stats_kind, double* colstats,
                                 b200_stream_t stream) {
  int rc = check_convt(d);
  if (rc) return rc;
  TcGemmProblem p{};
  p.a = ActView{(const bf16*)dy, d->n, 1, 2 * d->h, 2 * d->w, d->c_out, d->out_ctot, d->out_coff};
  p.wp = (const bf16*)wp_bwd;
  p.ncols = d->c_in; p.kdim = d->c_out;
  p.kd = 1; p.kh = 2; p.kw = 2;
  p.tap_mode = TC_TAP_K2S2; p.out_mode = TC_OUT_CONV;
  p.pn = d->n; p.pd = 1; p.ph = d->h; p.pw = d->w;
  p.out = (bf16*)dx; p.out_ctot = d->in_ctot; p.out_coff = d->in_coff;
  p.mask = (const bf16*)relu_mask; p.mask_ctot = d->in_ctot; p.mask_coff = d->in_coff;
  if (stats_kind != B200_STATS_NONE && stats_kind != B200_STATS_SUM && stats_kind != B200_STATS_SUM_SQ)
    return B200_ERR_BAD_ARG;
  if ((stats_kind != B200_STATS_NONE) != (colstats != nullptr)) return B200_ERR_BAD_ARG;
  p.stats = colstats; p.stats_kind = stats_kind;
  return run_gemm(p, d->impl, (cudaStream_t)stream);
}
int b200_convt2x2_bwd_weight(const b200_convt_desc* d, const void* x, const void* dy, float* dw,
                             b200_stream_t stream) {
  int rc = check_convt(d);
  if (rc) return rc;
  TcWgradProblem p{};
  p.mop = ActView{(const bf16*)x, d->n, 1, d->h, d->w, d->c_in, d->in_ctot, d->in_coff};
  p.nop = ActView{(const bf16*)dy, d->n, 1, 2 * d->h, 2 * d->w, d->c_out, d->out_ctot, d->out_coff};
  p.kd = 1; p.kh = 2; p.kw = 2;
  p.tap_mode = TC_TAP_K2S2;
  p.pn = d->n; p.pd = 1; p.ph = d->h; p.pw = d->w;
  p.dw = dw;
  p.s_m = (long)d->c_out * 4; p.s_n = 4;  // dw (c_in, c_out, 2, 2)
  return run_wgrad(p, d->impl, (cudaStream_t)stream);
}

// ---- ConvTranspose k3 s2 p1 through parity planes
static int check_deconv(const b200_deconv_desc* d) {
  if (!d || d->n <= 0 || d->d <= 0 || d->h <= 0 || d->w <= 0 || d->c_in <= 0 || d->c_out <= 0)
    return B200_ERR_BAD_ARG;
  if (d->kd != 1 && d->kd != 3) return B200_ERR_BAD_ARG;
  if (d->kd == 1 && d->d != 1) return B200_ERR_BAD_ARG;
  if (d->in_coff < 0 || d->in_coff + d->c_in > d->in_ctot) return B200_ERR_BAD_ARG;
  return B200_OK;
}
int b200_deconv3s2_repack_weight(const b200_deconv_desc* d, const float* w, void* wp_fwd, void* wp_bwd,
                                 b200_stream_t stream) {
  int rc = check_deconv(d);
  if (rc) return rc;
  const int ntaps = d->kd * 9;
  deconv_repack_kernel<<<grid_for((long)d->c_in * d->c_out * ntaps), 256, 0, (cudaStream_t)stream>>>(
      d->c_in, d->c_out, ntaps, w, (bf16*)wp_fwd, (bf16*)wp_bwd);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
int b200_deconv3s2_fwd(const b200_deconv_desc* d, const void* x, const void* wp_fwd, void* yp,
                       b200_stream_t stream) {
  int rc = check_deconv(d);
  if (rc) return rc;
  const int classes = d->kd == 1 ? 4 : 8;
  const long plane = (long)d->n * d->d * d->h * d->w * d->c_out;
  for (int cls = 0; cls < classes; ++cls) {
    TapTable t;
    deconv_fwd_table(d->kd, cls, &t);
    TcGemmProblem p{};
    p.a = ActView{(const bf16*)x, d->n, d->d, d->h, d->w, d->c_in, d->in_ctot, d->in_coff};
    p.wp = (const bf16*)wp_fwd;
    p.ncols = d->c_out; p.kdim = d->c_in;
    p.kd = d->kd; p.kh = 3; p.kw = 3;
    p.tap_mode = TC_TAP_SHIFT; p.table = &t; p.out_mode = TC_OUT_CONV;
    p.pn = d->n; p.pd = d->d; p.ph = d->h; p.pw = d->w;
    p.out = (bf16*)yp + cls * plane; p.out_ctot = d->c_out; p.out_coff = 0;
    rc = run_gemm(p, d->impl, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return B200_OK;
}
int b200_deconv3s2_bwd_data(const b200_deconv_desc* d, const void* dyp, const void* wp_bwd, void* dx,
                            b200_stream_t stream) {
  int rc = check_deconv(d);
  if (rc) return rc;
  const int classes = d->kd == 1 ? 4 : 8;
  TapTable t;
  deconv_bwd_table(d->kd, d->n, &t);
  TcGemmProblem p{};
  p.a = ActView{(const bf16*)dyp, classes * d->n, d->d, d->h, d->w, d->c_out, d->c_out, 0};
  p.wp = (const bf16*)wp_bwd;
  p.ncols = d->c_in; p.kdim = d->c_out;
  p.kd = d->kd; p.kh = 3; p.kw = 3;
  p.tap_mode = TC_TAP_SHIFT; p.table = &t; p.out_mode = TC_OUT_CONV;
  p.pn = d->n; p.pd = d->d; p.ph = d->h; p.pw = d->w;
  p.out = (bf16*)dx; p.out_ctot = d->in_ctot; p.out_coff = d->in_coff;
  return run_gemm(p, d->impl, (cudaStream_t)stream);
}
int b200_deconv3s2_bwd_weight(const b200_deconv_desc* d, const void* x, const void* dyp, float* dw,
                              b200_stream_t stream) {
  int rc = check_deconv(d);
  if (rc) return rc;
  const int classes = d->kd == 1 ? 4 : 8;
  TapTable t;
  deconv_bwd_table(d->kd, d->n, &t);
  TcWgradProblem p{};
  p.mop = ActView{(const bf16*)x, d->n, d->d, d->h, d->w, d->c_in, d->in_ctot, d->in_coff};
  p.nop = ActView{(const bf16*)dyp, classes * d->n, d->d, d->h, d->w, d->c_out, d->c_out, 0};
  p.kd = d->kd; p.kh = 3; p.kw = 3;
  p.tap_mode = TC_TAP_SHIFT; p.table = &t;
  p.pn = d->n; p.pd = d->d; p.ph = d->h; p.pw = d->w;
  p.dw = dw;
  const int ntaps = d->kd * 9;
  p.s_m = (long)d->c_out * ntaps; p.s_n = ntaps;  // dw (c_in, c_out, taps)
  return run_wgrad(p, d->impl, (cudaStream_t)stream);
}

}  // extern "C"
