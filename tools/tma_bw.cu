// TMA box-shape bandwidth probe: persistent CTAs stream boxes of a (2,128,128,128,C) bf16 tensor into
// shared memory (ring of `stages` slots, nothing consumes the data) and report chip-wide GB/s.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tmap.cuh"
using namespace b200;

struct Args { int bw, bh, bd, nbox_per_stage, stage_bytes, stages, tiles_h, tiles_w, D, total; int step_w, step_h; };

__global__ void __launch_bounds__(128, 1) stream(const __grid_constant__ CUtensorMap tm, const Args a, int c0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (sraw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int box_bytes = a.bw * a.bh * a.bd * 128;
    int issued = 0, waited = 0;
    uint32_t phase_bits = 0;
    for (int work = blockIdx.x; work < a.total; work += gridDim.x) {
      if (issued - waited == a.stages) {
        const int s = waited % a.stages;
        mbar_wait(&full[s], (phase_bits >> s) & 1);
        phase_bits ^= 1u << s;
        ++waited;
      }
      int mt = work;
      const int tw = mt % a.tiles_w; mt /= a.tiles_w;
      const int th = mt % a.tiles_h; mt /= a.tiles_h;
      const int d = mt % a.D; mt /= a.D;
      const int n = mt;
      const int s = issued % a.stages;
      mbar_arrive_expect_tx(&full[s], (uint32_t)(box_bytes * a.nbox_per_stage));
      for (int b = 0; b < a.nbox_per_stage; ++b)
        tma_load_5d(buf + (size_t)s * a.stage_bytes + (size_t)b * box_bytes, &tm, &full[s], c0,
                    tw * a.step_w - 1, th * a.step_h - 1, d - 1 + (a.nbox_per_stage > 1 ? b : 0), n);
      ++issued;
    }
    while (waited < issued) {
      const int s = waited % a.stages;
      mbar_wait(&full[s], (phase_bits >> s) & 1);
      phase_bits ^= 1u << s;
      ++waited;
    }
  }
}

int main() {
  const int N = 2, D = 128, H = 128, W = 128;
  for (int C : {64, 192}) {
    bf16* x;
    size_t elems = (size_t)N * D * H * W * C;
    cudaMalloc(&x, elems * 2);
    cudaMemset(x, 0, elems * 2);
    struct V { const char* name; int bw, bh, bd, nbox; int sw, sh; };
    V vs[] = {{"halo 10x18x3 one box", 10, 18, 3, 1, 8, 16}, {"halo 10x18x1 x3 boxes", 10, 18, 1, 3, 8, 16},
              {"tile 8x16x1 one box (per-tap style)", 8, 16, 1, 1, 8, 16}, {"halo 16x18x3 one box", 16, 18, 3, 1, 8, 16},
              {"halo 10x18x1 one box (2D style)", 10, 18, 1, 1, 8, 16}, {"halo 18x10x3 (tile 16w x 8h)", 18, 10, 3, 1, 16, 8},
              {"halo 34x6x3 (tile 32w x 4h)", 34, 6, 3, 1, 32, 4}};
    for (auto& v : vs) {
      CUtensorMap tm;
      uint64_t dims[5] = {(uint64_t)C, W, H, D, N};
      uint64_t str[5] = {1, (uint64_t)C, (uint64_t)W * C, (uint64_t)H * W * C, (uint64_t)D * H * W * C};
      uint32_t box[5] = {64, (uint32_t)v.bw, (uint32_t)v.bh, (uint32_t)v.bd, 1};
      if (make_tmap_bf16(&tm, x, 5, dims, str, box)) { printf("tmap fail\n"); return 1; }
      Args a{};
      a.bw = v.bw; a.bh = v.bh; a.bd = v.bd; a.nbox_per_stage = v.nbox;
      a.stage_bytes = ((v.bw * v.bh * v.bd * v.nbox * 128) + 1023) / 1024 * 1024;
      a.step_w = v.sw; a.step_h = v.sh;
      a.tiles_w = W / v.sw; a.tiles_h = H / v.sh; a.D = D;
      a.total = N * D * a.tiles_h * a.tiles_w;
      for (int stages : {2, 3}) {
        a.stages = stages;
        const int smem = 2048 + a.stages * a.stage_bytes;
        if (smem > 227 * 1024) continue;
        cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        stream<<<148, 128, smem>>>(tm, a, 0);
        cudaEventRecord(e0);
        for (int it = 0; it < 3; ++it) stream<<<148, 128, smem>>>(tm, a, 0);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        const double bytes = (double)a.total * v.bw * v.bh * v.bd * v.nbox * 128;
        printf("C=%3d %-40s stages=%d : %.3f ms  %.0f GB/s into smem (%.2f us per stage-load per SM)\n", C, v.name, stages, ms,
               bytes / ms / 1e6, ms * 1e3 / (a.total / 148.0));
      }
    }
    cudaFree(x);
  }
  return 0;
}
