// Host-callable launchers of every CUDA kernel in the library.  All of them enqueue on the given
// stream and never allocate or synchronise.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace b200 {

struct ActSpec;

// ---- implicit-GEMM convolution / GEMM on tcgen05 tensor cores (conv_tc.cu) -----------------
// GEMM view: D[M, Cout] = A[M, K] * W[Cout, K]^T with M = N*Ho*Wo output pixels and
// K = kh*kw*Cin_p gathered on the fly from the NHWC8 activation tensor.
struct ConvParams {
  const void* x = nullptr;     // activations NHWC [N,H,W,Cin_p], Cin_p % 8 == 0
  const void* w = nullptr;     // packed weights [Cout_alloc][Ktot], K order = (tap, cin), zero padded
  const float* bias = nullptr; // [Cout_alloc] fp32
  const void* res = nullptr;   // optional residual, same geometry as y
  void* y = nullptr;           // output [M, ldy]
  int N = 1, H = 1, W = 1, Cin_p = 8, Ho = 1, Wo = 1, Cout_p = 8;
  int kh = 1, kw = 1, sh = 1, sw = 1, dh = 1, dw = 1, pt = 0, pl = 0;
  int M = 0, Ktot = 0;
  int Cout_alloc = 0;          // N extent covered by tiles / entries of bias (multiple of the N tile)
  int w_ld = 0;                // weight row pitch in elements
  int w_rows = 0;              // valid weight rows (rows >= w_rows are read as zero)
  int w_k = 0;                 // valid K per weight row (k >= w_k is read as zero)
  int ldy = 0;                 // output row pitch in elements (>= Cout_p)
  int act1 = 0, act2 = 0;
  float a1 = 0, b1 = 0, a2 = 0, b2 = 0;
  int out_f32 = 0;             // write fp32 instead of the compute type
  // activation addressing override (elements; 0 = dense NHWC: Cin_p, W*Cin_p, H*W*Cin_p).  Used by the space-to-depth first
  // convolution, whose "pixels" are overlapping 64-element windows of a 16-channel image.  TMA paths only.
  int64_t x_sw = 0, x_sh = 0, x_sn = 0;
  // Walk the output tiles last-to-first.  The compiler alternates the direction from one kernel to the next so that a kernel starts on
  // the rows its producer wrote last, which are still resident in the 126 MB L2 (the tensors themselves are 100-400 MB at batch 256).
  int reverse = 0;
  // Second activation source of a K-concatenated GEMM (conv_tc3 only): a 1x1 convolution over x2 [N, H2, W2, K2] with stride (sh2, sw2)
  // producing the same [N, Ho, Wo] pixels.  K blocks [Ktot - K2, Ktot) of the packed weights belong to it.  Used to fold the projection
  // shortcut of a ResNet downsampling block into the block's last 1x1 convolution: conv_c(b) + conv_proj(x) = [b | x_strided] * [Wc ; Wp].
  const void* x2 = nullptr;
  int K2 = 0, H2 = 0, W2 = 0, sh2 = 1, sw2 = 1;
  // Second GEMM fused behind this one (conv_tc3f only): Z[M, N2] = act3(Y * w2[N2, Cout_p]^T + bias2), Y = this convolution's bf16 output.
  // Used to run the 1x1 reduce convolution of the next bottleneck block inside the kernel that produces its input.
  const void* w2 = nullptr;
  const float* bias2 = nullptr;
  void* z = nullptr;
  int N2 = 0, ldz = 0;
  int act3 = 0;
  float a3 = 0, b3 = 0;
  int batch = 1;               // blockIdx.z batches (batched MatMul); strides in elements
  // grouped weights (batched activation x activation MatMul flattened into one launch of the persistent kernel): output rows
  // [g*w_grp_m, (g+1)*w_grp_m) use weight rows [g*w_grp_rows, (g+1)*w_grp_rows); 0 = one weight matrix
  int w_grp_m = 0, w_grp_rows = 0;
  // grouped weights given MN-major: group g is a [Ktot][64] matrix (N contiguous; attention's V as it comes out of the head split) read
  // through an MN-major UMMA descriptor instead of being transposed first.  Single-CTA kernel, Cout == 64.
  int w_mn_major = 0;
  // the "weights" are an activation tensor written by an earlier kernel of the step (activation x activation MatMul): the kernels may not
  // read them before the programmatic-launch dependency has resolved (constant weights and bias are fetched while the producer drains)
  int w_dynamic = 0;
  // space-to-depth stem only (pair kernel): the 3x3 / stride-2 / pad-1 max pool that follows the ReLU is taken in the epilogue; y is the
  // POOLED tensor [N][Ho / 2][Wo / 2][Cout_p], zero-filled before the launch (the epilogue max-reduces into it; ReLU outputs are >= 0)
  int pool = 0;
  // head-split output (BERT's Reshape [N,S,H*64] -> [N,S,H,64] + Transpose (0,2,1,3) folded into the producing FullyConnected): row
  // n*S + s, columns [64 h, 64 h + 64) are stored at y[n][h][s][0..63]; 0 = plain [M][ldy] output.  Pair kernel only.
  int y_heads_s = 0, y_heads_h = 0;
  // the inverse (context MatMul + Transpose (0,2,1,3)): flat rows (n, h, s) of 64 columns are stored at y[n][s][h][0..63].  Single-CTA
  // kernel with grouped weights only (w_grp_m == y_heads_s).
  int y_heads_inv = 0;
  int64_t x_bs = 0, w_bs = 0, y_bs = 0;
};
// dtype: B200_BF16 or B200_F16.  bn: N tile (64 or 128).
void launch_conv_tc(const ConvParams& p, int dtype, int bn, cudaStream_t s);
int conv_tc_pick_bn(int cout_p);
void conv_tc_init_device();
// true when both fused activations are none / ReLU / Clamp (the two-instruction epilogue path of the tcgen05 kernels)
bool conv_acts_are_clamps(const ConvParams& p);
// Persistent TMA-fed variant (conv_tc2.cu): used whenever one weight matrix serves the whole launch.
bool conv_tc2_supported(const ConvParams& p);
bool conv_tc2_a_is_matrix(const ConvParams& p);
bool conv_tc2_a_is_im2col(const ConvParams& p);
bool conv_tc2_s2d_encode_ok();                    // the driver accepts the overlapping-window tensor map of the space-to-depth stem
bool conv_tc2_im2col_enabled();                   // B200_IM2COL != 0 and the driver exports cuTensorMapEncodeIm2col   // A tile by im2col-mode TMA (Cin_p % 64 == 0)
void launch_conv_tc2(const ConvParams& p, int dtype, int bn, cudaStream_t s);
void conv_tc2_init_device();
// CTA-pair variant (conv_tc3.cu): tcgen05.mma.cta_group::2 over 256 x {64,128,256} tiles, slab-ring epilogue, weight-stationary and
// 3x3 "row" modes.  Used for large launches (every SM pair has tiles to walk); B200_TC3=0 disables, =1 forces it wherever supported.
bool conv_tc3_supported(const ConvParams& p);
bool conv_tc3_preferred(const ConvParams& p);
// the stem + max-pool fusion (ConvParams::pool) applies to this launch (geometry without the pool flag set)
bool conv_tc3_s2d_pool_ok(const ConvParams& p);
int conv_tc3_pick_bn(const ConvParams& p);
const char* conv_tc3_mode_name(const ConvParams& p);
void launch_conv_tc3(const ConvParams& p, int dtype, int bn, cudaStream_t s);
void conv_tc3_init_device();
// Two chained 1x1 convolutions in one CTA-pair kernel (conv_tc3f.cu): ConvParams::w2 / z / N2 describe the second one.
bool conv_tc3f_supported(const ConvParams& p);
void launch_conv_tc3f(const ConvParams& p, int dtype, cudaStream_t s);
void conv_tc3f_init_device();
// CUDA-core fp32 path (fp32 mode: 1e-4 parity; also the reference for tcgen05 unit tests).
// Same parameter block; x/w/res/y are fp32, weights [Cout_alloc][Ktot] fp32.
void launch_conv_f32(const ConvParams& p, cudaStream_t s);

// ---- depthwise convolution (memory bound, CUDA cores) --------------------------------------
struct DwParams {
  const void* x; const void* w; const float* bias; void* y;  // w: [kh*kw][Cp] in compute type
  int N, H, W, Cp, Ho, Wo, kh, kw, sh, sw, dh, dw, pt, pl;
  int act; float a, b;
};
void launch_dwconv(const DwParams& p, int dtype, cudaStream_t s);
bool dwconv_strip_supported(const DwParams& p, int dtype);      // kernels_dw.cu: 3x3 stride 1/2, 16-bit data
void launch_dwconv_strip(const DwParams& p, int dtype, cudaStream_t s);
struct ConvParams;
bool conv_small_cin_supported(const ConvParams& p, int cin, int dtype);   // kernels_dw.cu: dense convolution over <= 4 input channels
void launch_conv_small_cin(const ConvParams& p, int cin, int dtype, cudaStream_t s);

// ---- fused inverted-residual block (mbconv.cu): expand 1x1 -> depthwise 3x3 -> project 1x1 (+ residual) in one tcgen05 kernel ----
struct MbParams {
  const void* x; void* y;              // NHWC8: [N][H][W][Cin_p] -> [N][Ho][Wo][Cout_p]
  const void* w1; const float* b1;     // expand weights [Ch_p][64 * ceil(Cin_p / 64)] (K-major, zero padded), bias [Ch_p]; unused without expand
  const void* wd; const float* bd;     // depthwise weights [9][Ch_p] in the compute type, bias [Ch_p]
  const void* w2; const float* b2;     // project weights [Cout_alloc][Ch_p], bias [Cout_alloc]
  int N, H, W, Cin_p, Ch_p, Cout_p, Cout_alloc, Ho, Wo, stride, pt, pl;   // Ch_p: hidden channels padded to a multiple of 64
  int has_expand, has_res;             // has_res: y += x (stride 1, Cin_p == Cout_p)
  float lo1, hi1, lo2, hi2;            // clamp bounds of the expand / depthwise activations (-inf / +inf: none)
};
bool mbconv_supported(const MbParams& p, int dtype);
void launch_mbconv(const MbParams& p, int dtype, cudaStream_t s);
void mbconv_init_device();

// ---- pooling ---------------------------------------------------------------------------------
struct PoolParams {
  const void* x; void* y;
  int N, H, W, Cp, Ho, Wo, kh, kw, sh, sw, pt, pl, pb, pr;
  int is_avg, exclude_pad;
};
void launch_pool(const PoolParams& p, int dtype, cudaStream_t s);
void launch_global_avgpool(const void* x, void* y, int N, int HW, int Cp, int dtype, cudaStream_t s);

// ---- elementwise -------------------------------------------------------------------------------
// unary activation over n physical elements
void launch_unary(const void* x, void* y, int64_t n, int dtype, int act, float a, float b, cudaStream_t s);
// FakeQuantize (ngraph reference fake_quantize.hpp:176-194) over n physical elements; ranges[4][cdim] = in_low, in_high, out_low, out_high per
// channel, channel of element i = (i / inner) % cdim (NHWC8: inner 1, cdim Cp; NCHW row-major: inner H*W, cdim C; per-tensor: cdim 1)
void launch_fake_quantize(const void* x, void* y, int64_t n, int dtype, const float* ranges, int cdim, int64_t inner, int levels, cudaStream_t s);
// binary with numpy broadcasting over logical row-major tensors (rank <= 8)
struct BinaryParams {
  const void* a; const void* b; void* y;
  int rank; int64_t oshape[8]; int64_t astride[8]; int64_t bstride[8];  // element strides, 0 = broadcast
  int64_t n; int op; int act; float aa, ab;
};
void launch_binary(const BinaryParams& p, int dtype, cudaStream_t s);
// same-shape fast path (any layout, physical element count), optional fused activation
void launch_binary_same(const void* a, const void* b, void* y, int64_t n, int dtype, int op, int act, float aa, float ab, cudaStream_t s);
// y[.., c] = act(x[.., c] * scale[c] + shift[c]) over NHWC8 (Cp) or rows of length Cp
void launch_scale_shift(const void* x, const float* scale, const float* shift, void* y, int64_t rows, int Cp, int dtype, int act,
                        float aa, float ab, cudaStream_t s);

// ---- softmax / MVN / gather ------------------------------------------------------------------
void launch_softmax(const void* x, void* y, int64_t outer, int64_t axis, int64_t inner, int dtype, cudaStream_t s);
// softmax(x * scale + mask[row / rows_per_mask]) over the last axis, one pass (mask may be null); cols % 8 == 0, cols <= 2048
bool softmax_fused_supported(int64_t cols, int dtype);
void launch_softmax_scale_mask(const void* x, const void* mask, void* y, int64_t rows, int64_t cols, float scale, int64_t rows_per_mask, int dtype,
                               cudaStream_t s);
void launch_mvn(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int64_t cols, int normalize_variance,
                int eps_outside, float eps, int dtype, cudaStream_t s);
void launch_gather_rows(const void* table, const int32_t* idx, void* y, int64_t n_idx, int64_t row_elems, int64_t table_rows,
                        int dtype, cudaStream_t s);
void launch_reduce(const void* x, void* y, int64_t outer, int64_t axis, int64_t inner, int dtype, int kind /*0 mean 1 sum 2 max*/, cudaStream_t s);

// ---- layout / dtype conversion ---------------------------------------------------------------
enum PhysLayout : int { PL_ROWMAJOR = 0, PL_NHWC8 = 1, PL_NHWC_DENSE = 2 };
// rank-4 logical [N,C,H,W]; converts dtype and layout in one pass
void launch_convert4(const void* src, int sdtype, int slayout, void* dst, int ddtype, int dlayout, int N, int C, int H, int W, cudaStream_t s);
// network input (any blob precision, NCHW or NHWC) -> space-to-depth [N, Hb, Wb, 16] in the compute type (Layout L_S2D16)
void launch_ingest_s2d(const void* src, int sdtype, int slayout, void* dst, int ddtype, int N, int C, int H, int W, int pt, int pl, int Hb, int Wb,
                       cudaStream_t s);
// flat dtype conversion
void launch_convert_flat(const void* src, int sdtype, void* dst, int ddtype, int64_t n, cudaStream_t s);
// generic permute of a dense row-major tensor: out[idx] = in[perm(idx)]
struct PermuteParams {
  const void* x; void* y; int rank; int64_t oshape[8]; int64_t istride[8]; int64_t n;
};
void launch_permute(const PermuteParams& p, int elem_size, cudaStream_t s);
// strided 2D copy: rows x row_bytes from src (pitch) to dst (pitch); used by concat / split
void launch_copy2d(const void* src, int64_t spitch, void* dst, int64_t dpitch, int64_t row_bytes, int64_t rows, cudaStream_t s);
void launch_fill_zero(void* dst, int64_t bytes, cudaStream_t s);

// ---- TopK / NMS / DetectionOutput (fp32 + i32, bit-exact ordering rules) ---------------------
void launch_topk(const float* x, float* vals, int32_t* idx, int64_t outer, int64_t axis, int64_t inner, int k, int mode_min, int sort_mode,
                 cudaStream_t s);
struct NmsParams {
  const float* boxes; const float* scores; int32_t* sel_idx; float* sel_scores; int32_t* valid;
  int B, C, Nb, max_out, center_format, sort_desc; float iou_thr, score_thr, soft_sigma;
  void* workspace; int64_t workspace_bytes;
};
int64_t nms_workspace_bytes(int B, int C, int Nb, int max_out, bool soft = false);
void launch_nms(const NmsParams& p, cudaStream_t s);
void nms_check(const NmsParams& p);                     // throws B200_ERR_UNSUPPORTED for geometry without a kernel (compile time)
// ---- fused attention (attention.cu) ---------------------------------------------------------------
struct AttnParams {
  const void* q; const void* k; const void* v;   // [N, H, S, dh] in the compute type (head-split layout)
  const void* mask;                              // [mask_batch, S] additive mask in the compute type, or null
  void* y;                                       // [N, S, H * dh] (heads merged)
  int N, H, S, dh, mask_batch;
  float scale;
};
bool attention_supported(const AttnParams& p, int dtype);
void launch_attention(const AttnParams& p, int dtype, cudaStream_t s);
void attention_init_device();
struct DetOutParams {
  const float* loc; const float* conf; const float* priors; float* out;
  int N, num_priors, num_classes, background, top_k, keep_top_k, code_center, share_location, var_in_target, normalized,
      clip_before, clip_after, input_h, input_w, decrease_label_id, num_loc_classes, priors_batch;
  float nms_thr, conf_thr;
  int64_t out_rows;
  void* workspace; int64_t workspace_bytes;
};
int64_t detout_workspace_bytes(const DetOutParams& p);
void detection_output_check(const DetOutParams& p);   // throws B200_ERR_UNSUPPORTED for geometry without a kernel (compile time)
void launch_detection_output(const DetOutParams& p, cudaStream_t s);
void detect_init_device();

}  // namespace b200
