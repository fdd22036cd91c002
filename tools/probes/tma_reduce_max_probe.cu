// Does cp.reduce.async.bulk.tensor ... .max work on a BFLOAT16 tensor map (sm_100a), with the box clipped at the tensor edge and the
// 128-byte swizzle?  Used by the stem + max-pool fusion (conv_tc3.cuh, A_S2D pool mode).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_reduce_max_probe tools/probes/tma_reduce_max_probe.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int C = 64, W = 56, H = 8, N = 2, P = 64;   // box {64 ch, 64 px (clipped to 56), 1 row, 1 image}

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, float base) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* tile = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  // value(px, c) = base + px - c / 8 (differs per block), written with the 128-byte swizzle: 16-byte chunk index ^ (px & 7)
  for (int i = threadIdx.x; i < P * C; i += blockDim.x) {
    const int px = i / C, c = i % C;
    const int chunk = (c / 8) ^ (px & 7);
    reinterpret_cast<__nv_bfloat16*>(tile)[px * C + chunk * 8 + (c % 8)] = __float2bfloat16(base + (float)px - (float)(c / 8) + (float)blockIdx.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int h = blockIdx.x % H, n = blockIdx.x / H % N;
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.max.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tm), "r"(smem_u32(tile)),
                 "r"(0), "r"(0), "r"(h), "r"(n)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  std::vector<__nv_bfloat16> y((size_t)N * H * W * C, __float2bfloat16(0.f));
  __nv_bfloat16* dy;
  CK(cudaMalloc(&dy, y.size() * 2));
  CK(cudaMemset(dy, 0, y.size() * 2));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  CUtensorMap tm;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, P, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 + 1024));
  // pass 1: every (n, h) row once with base 1; pass 2: twice as many blocks (each row hit twice, values base -3 + block index)
  probe_kernel<<<N * H, 128, 2 * 8192 + 1024>>>(tm, 1.f);
  CK(cudaDeviceSynchronize());
  probe_kernel<<<2 * N * H, 128, 2 * 8192 + 1024>>>(tm, -3.f);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(y.data(), dy, y.size() * 2, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int n = 0; n < N; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c) {
    const int b1 = n * H + h, b2 = b1 + N * H;
    const float v1 = 1.f + w - c / 8 + b1, v2 = -3.f + w - c / 8 + b1, v3 = -3.f + w - c / 8 + b2;
    float want = 0.f;
    want = v1 > want ? v1 : want; want = v2 > want ? v2 : want; want = v3 > want ? v3 : want;
    const float got = __bfloat162float(y[(((size_t)n * H + h) * W + w) * C + c]);
    if (got != want && bad++ < 8) printf("  mismatch n %d h %d w %d c %d: got %g want %g\n", n, h, w, c, got, want);
  }
  printf("tma reduce max bf16 (clipped box, swizzle 128B): %s (%d mismatches)\n", bad ? "BAD" : "OK", bad);
  return bad != 0;
}
