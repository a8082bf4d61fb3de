// Internal interface of the tcgen05 implicit-GEMM kernels (conv_tc.cu).
#pragma once
#include "common.cuh"

namespace b200 {

// Activation operand of an implicit GEMM: channels-last bf16 (N, D, H, W, C) slice.
struct ActView {
  const bf16* ptr;
  int n, d, h, w, c;
  int ctot, coff;
};

enum { TC_OUT_CONV = 0, TC_OUT_K2S2_SCATTER = 1 };
// how taps address the tap-dependent operand:
//   TC_TAP_SHIFT : taps shift the tile by (kd-1, kh-1, kw-1) with zero padding (3x3x3 conv)
//   TC_TAP_K2S2  : operand is (N, 1, 2H, 2W, C); tap (i, j) selects positions (2h+i, 2w+j)
enum { TC_TAP_SHIFT = 0, TC_TAP_K2S2 = 1 };

// Optional explicit tap list (overrides kd/kh/kw in TC_TAP_SHIFT mode): per tap a position shift, a
// batch-index offset (selects one of several stacked planes, e.g. the parity classes of a stride-2
// transposed convolution) and the index of the weight slab to use.
struct TapTable {
  int ntaps;
  signed char dd[27], dh[27], dw[27], widx[27];
  int dn[27];
};

// Forward / data-gradient implicit GEMM over position tiles of 128 rows:
//   out[pos][n] = sum_{tap, k} A[pos (+) tap][k] * Wp[tap][n][k]   (+ bias, ReLU, mask)
struct TcGemmProblem {
  ActView a;              // A operand (x for fwd, dy for dgrad); for TC_TAP_K2S2 h,w are 2H,2W
  const bf16* wp;         // [ntaps][ncols][kdim] bf16, K contiguous
  int ncols, kdim;        // GEMM N and per-tap K
  int kd, kh, kw;         // taps per axis (TC_TAP_K2S2: 1,2,2)
  int tap_mode;           // TC_TAP_*
  const TapTable* table;  // host pointer or NULL
  int out_mode;           // TC_OUT_*
  int pn, pd, ph, pw;     // position grid the 128-row tiles enumerate
  bf16* out;
  int out_ctot, out_coff;
  int cout_t;             // TC_OUT_K2S2_SCATTER: channels per (i,j) block (ncols = 4 * cout_t)
  const float* bias;      // per output channel or NULL
  int relu;
  const bf16* mask;       // optional: out = 0 where mask <= 0; same geometry as out
  int mask_ctot, mask_coff;
  // optional per-sample column statistics of the stored (bf16) output, accumulated into
  // stats fp64 [pn][2][ncols]: [.][0][c] += sum out, [.][1][c] += sum out^2, or sum out * dot when a
  // dot operand (same geometry as out) is given.  Fused into the epilogue of the tcgen05 kernels.
  double* stats;
  int stats_kind;         // 1 = only the sum is needed ([.][1][c] left untouched); otherwise sum + second moment
  const bf16* dot;
  int dot_ctot, dot_coff;
};
int tc_gemm_conv(const TcGemmProblem& p, cudaStream_t stream);
bool tc_gemm_supported(const TcGemmProblem& p);

// Weight-gradient implicit GEMM (reduction over positions):
//   dw[m*s_m + n*s_n + tap] += sum_pos Mop[pos][m] * Nop[pos (+) tap][n]
struct TcWgradProblem {
  ActView mop;            // un-shifted operand (dy for conv, x for convT 2x2)
  ActView nop;            // tap-addressed operand (x for conv, dy for convT 2x2)
  int kd, kh, kw;
  int tap_mode;
  const TapTable* table;  // host pointer or NULL (then taps = kd*kh*kw shifts)
  int pn, pd, ph, pw;     // position grid
  float* dw;
  long s_m, s_n;
  float* dbias;           // optional (only when tc_wgrad_fuses_bias): dbias[m] += sum_pos Mop[pos][m]
};
int tc_wgrad(const TcWgradProblem& p, cudaStream_t stream);
bool tc_wgrad_supported(const TcWgradProblem& p);
bool tc_wgrad_fuses_bias(const TcWgradProblem& p);

}  // namespace b200
