// Hardware probe #2: MN-major SWIZZLE_128B operands (the weight-gradient GEMM reduces over positions,
// so channels are the contiguous "MN" axis and positions are K).  Questions:
//   * may the B operand start at an arbitrary 128-byte row (position) with an 8-row K-group stride
//     (SBO) of 1280 B, i.e. read a shifted tap out of a halo tile?
//   * may B use N = 32 or 48 channels of the 64-channel rows, optionally starting 64 B into the row?
// D[m][n] = sum_k A[k][m] * B[rowmap(k)][ncol0 + n],  A = identity-like so D is readable.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tmap.cuh"
using namespace b200;

#define ROWS 384

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA,
                                             const __grid_constant__ CUtensorMap tmB, int start_row,
                                             int sbo_bytes, int N, int col_bytes, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  uint8_t* sB = smem + 1024;             // ROWS x 128 B "halo" (positions x 64 channels)
  uint8_t* sA = sB + ROWS * 128;         // 2 boxes of [128 positions x 64 channels] (M = 128)
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, ROWS * 128 + 2 * 128 * 128);
    for (int b = 0; b < ROWS / 128; ++b) tma_load_2d(sB + b * 128 * 128, &tmB, bar, 0, b * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sA + 128 * 128, &tmA, bar, 64, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
    for (int j = 0; j < 8; ++j) {  // 16 positions per MMA
      const uint64_t da = make_smem_desc_sw128(smem_u32(sA) + j * 2048, 128 * 128, 1024, 0);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sB) + start_row * 128 + col_bytes + j * 2 * sbo_bytes,
                                               128 * 128, sbo_bytes, 0);
      umma_bf16(tmem, da, db, idesc, j != 0);
    }
    umma_commit(bar2);
  }
  __syncthreads();
  mbar_wait(bar2, 0);
  tc_fence_after();
  const int row = threadIdx.x;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[row * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  // A[k][m] (128 positions x 128 channels): A[k][m] = 1 if m == k else 0  -> D[m][n] = B[rowmap(m)][n]
  std::vector<bf16> hA(128 * 128), hB(ROWS * 64);
  for (int k = 0; k < 128; ++k)
    for (int m = 0; m < 128; ++m) hA[k * 128 + m] = __float2bfloat16(m == k ? 1.f : 0.f);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < 64; ++c) hB[r * 64 + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251));
  bf16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  uint64_t dimsA[2] = {128, 128}, strA[2] = {1, 128};
  uint32_t boxA[2] = {64, 128};
  uint64_t dimsB[2] = {64, ROWS}, strB[2] = {1, 64};
  uint32_t boxB[2] = {64, 128};
  if (make_tmap_bf16(&tmA, dA, 2, dimsA, strA, boxA) || make_tmap_bf16(&tmB, dB, 2, dimsB, strB, boxB)) {
    printf("tensor map failed\n");
    return 1;
  }
  const int smem = 2048 + ROWS * 128 + 2 * 128 * 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hOut(128 * 64);
  const int starts[] = {0, 1, 3, 11};
  const int sbos[] = {1024, 1280};
  const int ns[] = {64, 48, 32, 16};
  const int cols[] = {0, 32, 64};
  for (int s : starts)
    for (int sbo : sbos)
      for (int N : ns)
        for (int cb : cols) {
          if (cb / 2 + N > 64) continue;
          cudaMemset(dOut, 0, 128 * 64 * 4);
          probe<<<1, 128, smem>>>(tmA, tmB, s, sbo, N, cb, dOut);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("start=%d sbo=%d N=%d col=%dB : CUDA error %s\n", s, sbo, N, cb, cudaGetErrorString(e));
            return 2;
          }
          cudaMemcpy(hOut.data(), dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int m = 0; m < 128; ++m) {
            const int r = s + (m / 8) * (sbo / 128) + (m % 8);
            for (int n = 0; n < N; ++n) {
              const float want = (float)((r * 7 + (cb / 2 + n) * 3) % 251);
              if (hOut[m * 64 + n] != want) ++bad;
            }
          }
          printf("MN-major B: start_row=%d sbo=%d N=%d col_off=%dB : %s (%d wrong)\n", s, sbo, N, cb,
                 bad ? "MISMATCH" : "EXACT", bad);
        }
  return 0;
}
