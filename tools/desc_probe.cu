// Hardware probe: does tcgen05.mma accept a K-major SWIZZLE_128B operand whose start address is a
// 128-byte row that is NOT 1024-byte aligned, and 8-row group strides (SBO) that are not multiples
// of 1024 B?  (Needed to read the kw-shifted taps of a conv from ONE halo tile in shared memory.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../mdeical_image_segmentation_b200/csrc desc_probe.cu -o desc_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tmap.cuh"
using namespace b200;

#define ROWS 384  // smem rows of 128 B filled by TMA (3 boxes of 128 rows)

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA,
                                             const __grid_constant__ CUtensorMap tmB, int start_row,
                                             int sbo_bytes, int base_off, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  uint8_t* sA = smem + 1024;
  uint8_t* sB = sA + ROWS * 128;
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
    mbar_arrive_expect_tx(bar, ROWS * 128 + 64 * 128);
    for (int b = 0; b < ROWS / 128; ++b) tma_load_2d(sA + b * 128 * 128, &tmA, bar, 0, b * 128);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t a0 = smem_u32(sA) + start_row * 128;
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = make_smem_desc_sw128(a0 + k * 32, 16, sbo_bytes, base_off);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024, 0);
      umma_bf16(tmem, da, db, idesc, k != 0);
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
  std::vector<bf16> hA(ROWS * 64), hB(64 * 64);
  for (int r = 0; r < ROWS; ++r)
    for (int k = 0; k < 64; ++k) hA[r * 64 + k] = __float2bfloat16((float)((r * 7 + k * 3) % 251));
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  bf16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  uint64_t dimsA[2] = {64, ROWS}, strA[2] = {1, 64};
  uint32_t boxA[2] = {64, 128};
  uint64_t dimsB[2] = {64, 64}, strB[2] = {1, 64};
  uint32_t boxB[2] = {64, 64};
  if (make_tmap_bf16(&tmA, dA, 2, dimsA, strA, boxA) || make_tmap_bf16(&tmB, dB, 2, dimsB, strB, boxB)) {
    printf("tensor map failed\n");
    return 1;
  }
  const int smem = 1024 + 1024 + ROWS * 128 + 64 * 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hOut(128 * 64);
  const int starts[] = {0, 1, 2, 3, 5, 8, 9};
  const int sbos[] = {1024, 1280, 2048, 2304};
  for (int s : starts)
    for (int sbo : sbos)
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        const int bo = bo_mode ? (s & 7) : 0;
        if (bo_mode && bo == 0) continue;
        cudaMemset(dOut, 0, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(tmA, tmB, s, sbo, bo, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("start=%d sbo=%d base_off=%d : CUDA error %s\n", s, sbo, bo, cudaGetErrorString(e));
          return 2;
        }
        cudaMemcpy(hOut.data(), dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first_bad = -1;
        for (int m = 0; m < 128; ++m) {
          const int r = s + (m / 8) * (sbo / 128) + (m % 8);
          for (int n = 0; n < 64; ++n) {
            const float want = (float)((r * 7 + n * 3) % 251);
            if (hOut[m * 64 + n] != want) {
              if (first_bad < 0) first_bad = m * 64 + n;
              ++bad;
            }
          }
        }
        printf("start_row=%d sbo=%d base_off=%d : %s (%d wrong, first at m=%d n=%d)\n", s, sbo, bo,
               bad ? "MISMATCH" : "EXACT", bad, first_bad / 64, first_bad % 64);
      }
  return 0;
}
