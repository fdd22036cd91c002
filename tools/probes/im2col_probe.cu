// Empirical check of the im2col-mode TMA semantics on sm_100a (coordinates, im2col offsets, traversal stride,
// bounding-box corners, image wrap-around, OOB fill).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o im2col_probe im2col_probe.cu
// The input tensor is NHWC bf16 with channel 0/1/2 = n/h/w of the pixel and channel 3 = 1 (so zero-filled rows are recognisable).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int PIX, int CH>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int offw, int offh, __nv_bfloat16* out) {
  __shared__ __align__(1024) __nv_bfloat16 tile[PIX * CH];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < PIX * CH; i += blockDim.x) tile[i] = __float2bfloat16(-7.f);
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bar)), "r"(PIX * CH * 2) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(
            smem_u32(tile)),
        "l"(&tm), "r"(smem_u32(&bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"((unsigned short)offw), "h"((unsigned short)offh)
        : "memory");
  }
  // wait (bounded)
  uint32_t ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PIX * CH; i += blockDim.x) out[i] = tile[i];
  if (threadIdx.x == 0 && !ok) printf("  (mbarrier never completed)\n");
}

int main() {
  const int N = 2, H = 6, W = 7, C = 64, PIX = 16;
  std::vector<__nv_bfloat16> x((size_t)N * H * W * C);
  for (int n = 0; n < N; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c) {
    float v = c == 0 ? n : c == 1 ? h : c == 2 ? w : c == 3 ? 1.f : 0.25f * (c % 8);
    x[(((size_t)n * H + h) * W + w) * C + c] = __float2bfloat16(v);
  }
  __nv_bfloat16 *dx, *dout;
  CK(cudaMalloc(&dx, x.size() * 2));
  CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dout, PIX * C * 2));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fp, cudaEnableDefault, &q));
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fp);
  if (!enc) { printf("no cuTensorMapEncodeIm2col\n"); return 1; }

  struct Case { const char* name; int lw, lh, uw, uh, sw, sh; int c, w, h, n, ow, oh; };
  const Case cases[] = {
      {"A  3x3 p1 s1: coords=(lower corner)      off(0,0)", -1, -1, -1, -1, 1, 1, 0, -1, -1, 0, 0, 0},
      {"B  3x3 p1 s1: coords=(w2,h1)             off(0,0)", -1, -1, -1, -1, 1, 1, 0, 2, 1, 0, 0, 0},
      {"B2 3x3 p1 s1: coords=(w2,h1)             off(2,1)", -1, -1, -1, -1, 1, 1, 0, 2, 1, 0, 2, 1},
      {"C  3x3 p1 s1: coords=(0,0)               off(1,1)", -1, -1, -1, -1, 1, 1, 0, 0, 0, 0, 1, 1},
      {"D  3x3 p1 s2: coords=(lower corner)      off(0,0)", -1, -1, -1, -1, 2, 2, 0, -1, -1, 0, 0, 0},
      {"D2 3x3 p1 s2: coords=(w1,h1) (2nd pos)   off(2,2)", -1, -1, -1, -1, 2, 2, 0, 1, 1, 0, 2, 2},
      {"D3 3x3 p1 s2: coords=(w1? as q=1,p=0)    off(0,0)", -1, -1, -1, -1, 2, 2, 0, 1, -1, 0, 0, 0},
      {"E  1x1 p0 s2: lower 0 upper -1/-1        off(0,0)", 0, 0, 0, -1, 2, 2, 0, 0, 0, 0, 0, 0},
      {"F  3x3 p1 s1: start in image 1, row 5    off(1,1)", -1, -1, -1, -1, 1, 1, 0, 3, 4, 1, 1, 1},
  };
  for (const Case& cs : cases) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {cs.lw, cs.lh}, upper[2] = {cs.uw, cs.uh};
    cuuint32_t estr[4] = {1, (cuuint32_t)cs.sw, (cuuint32_t)cs.sh, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lower, upper, C, PIX, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%s  encode rc=%d\n", cs.name, (int)r);
    if (r != CUDA_SUCCESS) continue;
    probe_kernel<PIX, C><<<1, 128>>>(tm, cs.c, cs.w, cs.h, cs.n, cs.ow, cs.oh, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<__nv_bfloat16> o(PIX * C);
    CK(cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost));
    printf("  rows (n,h,w | valid): ");
    for (int p = 0; p < PIX; ++p)
      printf("(%g,%g,%g|%g) ", __bfloat162float(o[p * C + 0]), __bfloat162float(o[p * C + 1]), __bfloat162float(o[p * C + 2]), __bfloat162float(o[p * C + 3]));
    printf("\n");
  }
  return 0;
}
