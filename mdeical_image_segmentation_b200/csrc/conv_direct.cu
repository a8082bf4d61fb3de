// CUDA-core implementations of the same implicit-GEMM problems as conv_tc.cu. They serve the layers
// the tensor-core path cannot take (first layers with 1..4 input channels, channel counts that are
// not multiples of 8) and act as an on-device cross-check of the tcgen05 kernels in the tests.
#include "conv_tc.cuh"

namespace b200 {

__device__ __forceinline__ float ldbf(const bf16* p) { return __bfloat162float(*p); }

// source row of the tap-addressed operand for output position (n,d,h,w); -1 if it is zero padding
__device__ __forceinline__ long tap_row(const ActView& v, int tap_mode, int n, int d, int h, int w,
                                        int kdi, int khi, int kwi, int pad_d, int pad_h,
                                        int pad_w) {
  if (tap_mode == TC_TAP_SHIFT) {
    const int sd = d + kdi - pad_d, sh = h + khi - pad_h, sw = w + kwi - pad_w;
    if (sd < 0 || sd >= v.d || sh < 0 || sh >= v.h || sw < 0 || sw >= v.w) return -1;
    return (((long)n * v.d + sd) * v.h + sh) * v.w + sw;
  }
  return ((long)n * v.h + (2 * h + khi)) * v.w + (2 * w + kwi);
}

#define DIRECT_CO 16
#define DIRECT_CK 16

__global__ void __launch_bounds__(128)
    direct_gemm_kernel(const TcGemmProblem p, const long npos, const int use_table, const TapTable tt) {
  __shared__ float ws[27 * DIRECT_CK * DIRECT_CO];
  const int ntaps = use_table ? tt.ntaps : p.kd * p.kh * p.kw;
  const int n0 = blockIdx.y * DIRECT_CO;
  const long pos = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pos < npos;
  int n = 0, d = 0, h = 0, w = 0;
  if (active) {
    long r = pos;
    w = (int)(r % p.pw); r /= p.pw;
    h = (int)(r % p.ph); r /= p.ph;
    d = (int)(r % p.pd); r /= p.pd;
    n = (int)r;
  }
  float acc[DIRECT_CO];
#pragma unroll
  for (int j = 0; j < DIRECT_CO; ++j) acc[j] = 0.f;
  const int pad_d = p.kd / 2, pad_h = p.kh / 2, pad_w = p.kw / 2;
  const bool shift = p.tap_mode == TC_TAP_SHIFT;

  for (int c0 = 0; c0 < p.kdim; c0 += DIRECT_CK) {
    __syncthreads();
    for (int i = threadIdx.x; i < ntaps * DIRECT_CK * DIRECT_CO; i += blockDim.x) {
      const int j = i % DIRECT_CO, ck = (i / DIRECT_CO) % DIRECT_CK, tap = i / (DIRECT_CO * DIRECT_CK);
      float v = 0.f;
      const int wt_ = use_table ? tt.widx[tap] : tap;
      if (c0 + ck < p.kdim && n0 + j < p.ncols)
        v = ldbf(p.wp + ((long)wt_ * p.ncols + (n0 + j)) * p.kdim + c0 + ck);
      ws[i] = v;
    }
    __syncthreads();
    if (!active) continue;
    const int ckn = min(DIRECT_CK, p.kdim - c0);
    for (int tap = 0; tap < ntaps; ++tap) {
      const int kwi = tap % p.kw, khi = (tap / p.kw) % p.kh, kdi = tap / (p.kw * p.kh);
      const long row = use_table
                           ? tap_row(p.a, TC_TAP_SHIFT, n + tt.dn[tap], d, h, w, tt.dd[tap], tt.dh[tap],
                                     tt.dw[tap], 0, 0, 0)
                           : tap_row(p.a, p.tap_mode, n, d, h, w, kdi, khi, kwi, shift ? pad_d : 0,
                                     shift ? pad_h : 0, shift ? pad_w : 0);
      if (row < 0) continue;
      const bf16* src = p.a.ptr + row * p.a.ctot + p.a.coff + c0;
      const float* wt = ws + tap * DIRECT_CK * DIRECT_CO;
      for (int ck = 0; ck < ckn; ++ck) {
        const float xv = ldbf(src + ck);
#pragma unroll
        for (int j = 0; j < DIRECT_CO; ++j) acc[j] = fmaf(xv, wt[ck * DIRECT_CO + j], acc[j]);
      }
    }
  }
  if (!active) return;
  long opos;
  int ocol0 = n0;
  if (p.out_mode == TC_OUT_CONV) {
    opos = pos;
  } else {
    const int ij = n0 / p.cout_t;
    ocol0 = n0 - ij * p.cout_t;
    opos = ((long)n * (2 * p.ph) + (2 * h + (ij >> 1))) * (2 * p.pw) + (2 * w + (ij & 1));
  }
  bf16* orow = p.out + opos * p.out_ctot + p.out_coff + ocol0;
  const bf16* mrow = p.mask ? p.mask + opos * p.mask_ctot + p.mask_coff + ocol0 : nullptr;
#pragma unroll
  for (int j = 0; j < DIRECT_CO; ++j) {
    if (n0 + j >= p.ncols) break;
    float v = acc[j];
    if (p.bias) v += p.bias[ocol0 + j];
    if (p.relu) v = fmaxf(v, 0.f);
    if (mrow && !(ldbf(mrow + j) > 0.f)) v = 0.f;
    orow[j] = __float2bfloat16(v);
  }
}

int direct_gemm_conv(const TcGemmProblem& p, cudaStream_t stream) {
  if (p.kd * p.kh * p.kw > 27) return B200_ERR_UNSUPPORTED;
  TapTable tt{};
  if (p.table) tt = *p.table;
  if (p.out_mode == TC_OUT_K2S2_SCATTER && p.cout_t % DIRECT_CO) return B200_ERR_UNSUPPORTED;
  const long npos = (long)p.pn * p.pd * p.ph * p.pw;
  if (npos == 0) return B200_OK;
  dim3 grid((unsigned)((npos + 127) / 128), (unsigned)((p.ncols + DIRECT_CO - 1) / DIRECT_CO));
  direct_gemm_kernel<<<grid, 128, 0, stream>>>(p, npos, p.table != nullptr, tt);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ---------------------------------------------------------------------------------- wgrad
// thread -> (m within a 64-wide block, slot group); each thread owns WG_ACC (n, tap) pairs and
// reduces over a chunk of positions, then one atomicAdd per accumulator.
#define WG_ACC 7
__global__ void __launch_bounds__(256)
    direct_wgrad_kernel(const TcWgradProblem p, const long npos, const long chunk, const int use_table,
                        const TapTable tt) {
  const int ntaps = use_table ? tt.ntaps : p.kd * p.kh * p.kw;
  const int m = blockIdx.y * 64 + (threadIdx.x & 63);
  const int grp = threadIdx.x >> 6;
  const int combos = p.nop.c * ntaps;
  const int cbase = (blockIdx.z * 4 + grp) * WG_ACC;
  int cn[WG_ACC], ckd[WG_ACC], ckh[WG_ACC], ckw[WG_ACC], ctap[WG_ACC];
  float acc[WG_ACC];
#pragma unroll
  for (int a = 0; a < WG_ACC; ++a) {
    const int c = cbase + a;
    acc[a] = 0.f;
    cn[a] = -1;
    ckd[a] = ckh[a] = ckw[a] = ctap[a] = 0;
    if (c < combos) {
      const int tap = c % ntaps;
      cn[a] = c / ntaps;
      ctap[a] = tap;
      ckw[a] = tap % p.kw; ckh[a] = (tap / p.kw) % p.kh; ckd[a] = tap / (p.kw * p.kh);
    }
  }
  const bool shift = p.tap_mode == TC_TAP_SHIFT;
  const int pad_d = shift ? p.kd / 2 : 0, pad_h = shift ? p.kh / 2 : 0, pad_w = shift ? p.kw / 2 : 0;
  const long pbeg = (long)blockIdx.x * chunk;
  const long pend = min(npos, pbeg + chunk);
  if (m < p.mop.c) {
    for (long pos = pbeg; pos < pend; ++pos) {
      long r = pos;
      const int w = (int)(r % p.pw); r /= p.pw;
      const int h = (int)(r % p.ph); r /= p.ph;
      const int d = (int)(r % p.pd); r /= p.pd;
      const int n = (int)r;
      const float mv = ldbf(p.mop.ptr + pos * p.mop.ctot + p.mop.coff + m);
#pragma unroll
      for (int a = 0; a < WG_ACC; ++a) {
        if (cn[a] < 0) continue;
        const long row = use_table ? tap_row(p.nop, TC_TAP_SHIFT, n + tt.dn[ctap[a]], d, h, w,
                                             tt.dd[ctap[a]], tt.dh[ctap[a]], tt.dw[ctap[a]], 0, 0, 0)
                                   : tap_row(p.nop, p.tap_mode, n, d, h, w, ckd[a], ckh[a], ckw[a], pad_d,
                                             pad_h, pad_w);
        if (row < 0) continue;
        acc[a] = fmaf(mv, ldbf(p.nop.ptr + row * p.nop.ctot + p.nop.coff + cn[a]), acc[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < WG_ACC; ++a) {
      if (cn[a] < 0) continue;
      const int tap = use_table ? tt.widx[ctap[a]] : (ckd[a] * p.kh + ckh[a]) * p.kw + ckw[a];
      atomicAdd(p.dw + (long)m * p.s_m + (long)cn[a] * p.s_n + tap, acc[a]);
    }
  }
}

int direct_wgrad(const TcWgradProblem& p, cudaStream_t stream) {
  const long npos = (long)p.pn * p.pd * p.ph * p.pw;
  if (npos == 0) return B200_OK;
  TapTable tt{};
  if (p.table) tt = *p.table;
  const int ntaps = p.table ? tt.ntaps : p.kd * p.kh * p.kw;
  const int combos = p.nop.c * ntaps;
  const unsigned gy = (p.mop.c + 63) / 64, gz = (combos + 4 * WG_ACC - 1) / (4 * WG_ACC);
  long want_x = (148L * 8 + (long)gy * gz - 1) / ((long)gy * gz);
  if (want_x < 1) want_x = 1;
  long chunk = (npos + want_x - 1) / want_x;
  if (chunk < 32) chunk = 32;
  const unsigned gx = (unsigned)((npos + chunk - 1) / chunk);
  if (gz > 65535 || gy > 65535) return B200_ERR_UNSUPPORTED;
  direct_wgrad_kernel<<<dim3(gx, gy, gz), 256, 0, stream>>>(p, npos, chunk, p.table != nullptr, tt);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace b200
