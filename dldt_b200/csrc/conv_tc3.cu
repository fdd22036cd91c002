// CTA-pair implicit-GEMM convolution / GEMM on tcgen05 (sm_100a), version 3: cta_group::2, 256 x BN tiles.
//
//   D[M, Cout] = im2col(X)[M, K] * Wp[Cout, K]^T  (+ bias, activation, + residual, activation)
//
// Same arithmetic contract as conv_tc2.cu (ngraph reference/convolution.hpp:89-186; post-op chain of
// nodes/mkldnn_conv_node.cpp:355-403).  Two CTAs of a cluster (the two SMs of a TPC) compute one 256 x BN tile with ONE
// tcgen05.mma.cta_group::2 per 16 K elements issued by the leader: each CTA stages its own 128 activation rows and only HALF of the
// weight tile.  A 128 x 128 single-CTA tile reads 128 B/clk of operands from shared memory AND writes 128 B/clk of new operands into
// it, which caps it at 50 % of the tensor pipe; the 256 x 256 pair tile needs 64 + 64 B/clk.  Further traffic cuts:
//   * weight-stationary mode: a layer with a single N tile keeps its half of the packed weights resident in shared memory
//   * the epilogue works one 64-channel slab at a time through small staging rings (2 output slabs, 4 residual slabs) instead of
//     whole-tile double buffers, which is what lets BN = 256 (512 TMEM columns, two accumulators) fit beside the operand rings
//   * "halo" mode for 3x3 / stride-1 / pad-1 convolutions: the A operand of one 64-channel chunk is ONE tiled TMA box
//     {64 ch, P pixels, 128/P + 2 rows} (P = power of two >= W + 2) and the nine taps are nine MMAs whose descriptors start
//     (r * P + s) * 128 bytes into the box (one pixel = one 128-byte swizzled row; the hardware swizzle is a function of the absolute
//     shared-memory address, so a descriptor may start on any 128-byte row: probed on B200, profiles/r1_call_a), so the activations
//     cross the L2 -> SM fabric (R + 2) / R times instead of 9 times.  Output rows with w >= W are computed and dropped by a 3-D store map.
//
//   warp 0-7  epilogue (TMEM lane quarter = warp % 4, column half = warp / 4)   warp 8   MMA issuer (leader CTA) + TMEM alloc
//   warp 9    operand TMA issuer (A ring, B ring, resident weights)              warp 10  residual TMA issuer
//
// mbarriers: fullA/emptyA[s], fullB/emptyB[s] (operand rings; full* live in the leader and collect both CTAs' bytes),
// tfull[a] (multicast commit), tempty[a] (leader; one arrival per CTA), rfull/rempty[slot] (local), wfull (leader).
#include "conv_tc3.cuh"

namespace b200 {

// defined in conv_tc2.cu
CUtensorMap make_tmap_2d(const void* base, bool bf16, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer);
CUtensorMap make_tmap_im2col(const ConvParams& p, bool bf16);
PFN_cuTensorMapEncodeTiled_v12000 encode_tiled_fn();
tc::EpiAct make_epi(int kind, float a, float b);

namespace tc3 {
B200_TC3_DEFINE_TYPED(__nv_bfloat16)
}

namespace {

using namespace tc3;

int g_sm_count3 = 0;
int g_max_clusters3 = 0;   // CTA pairs that can be co-resident (cudaOccupancyMaxActiveClusters)
inline int pair_slots() { return g_max_clusters3 > 0 ? g_max_clusters3 : (g_sm_count3 > 0 ? g_sm_count3 : 148) / 2; }

// rank-3/4 tiled maps over NHWC activations (halo mode)
CUtensorMap make_tmap_tiled(const void* base, bool bf16, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  auto enc = encode_tiled_fn();
  B200_CHECK(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap m;
  cuuint64_t d[5]; cuuint64_t st[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = enc(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled (rank " << rank << ") failed (" << (int)r << ")");
  return m;
}

int rows_pitch(const ConvParams& p) {
  int P = 8;
  while (P < p.W + p.kw - 1) P <<= 1;
  return P;
}

// Halo mode applies to dense 3x3 stride-1, dilation-1, pad-1 convolutions whose rows fit a 64-pixel box.
bool halo_ok(const ConvParams& p) {
  const char* e = getenv("B200_TC3_HALO");
  if (e && e[0] == '0') return false;
  if (!(p.kw == 3 && p.kh == 3 && p.sh == 1 && p.sw == 1 && p.dh == 1 && p.dw == 1 && p.pl == 1 && p.pt == 1 && p.Wo == p.W && p.Ho == p.H &&
        p.Cin_p % 64 == 0 && p.x_sw == 0))
    return false;
  const int P = rows_pitch(p);
  if (P > 64) return false;
  const int R = BM / P;
  // a tile covers R image rows x P pixel slots; what hangs over the image is computed and dropped.  Worth it while at least half of
  // the MMA rows are real output pixels (14x14: 77 %, 7x7: 38 % -> im2col mode instead)
  const int tiles = (p.Ho + R - 1) / R;
  // measured on B200 (profiles/README.md): 56x56 and 28x28 (87.5 % useful rows) gain 25-35 % over im2col mode, 14x14 (77 %, one more
  // round of tiles) loses 15 %.  B200_TC3_HALO=2 forces halo mode wherever the geometry allows (tests of the partial-tile path).
  const double useful = (double)(p.Ho * p.Wo) / (double)(tiles * BM);
  return p.ldy == p.Cout_p && useful >= ((e && e[0] == '2') ? 0.3 : 0.8);
}

// The space-to-depth stem (compile.cu lower_stem_s2d): a 4x4 stride-1 window over 16-channel block pixels, described to the single-CTA
// kernel as "4 taps of 64 overlapping channels".  Here the box of block pixels is loaded ONCE per output row segment and every tap is a
// shifted 32-byte-swizzled view of it, so the input crosses the L2 -> SM fabric 4x (block rows) instead of 16x.
bool s2d_ok(const ConvParams& p) {
  const char* e = getenv("B200_TC3_S2D");
  if (e && e[0] == '0') return false;
  // block row pitch a multiple of 16 bytes, groups of 4 output pixels, no residual, one 64-channel N tile (four phase accumulators fill TMEM),
  // one segment per row (group 31 of the box reads group 32 = the next block row: it must be a dropped one, so Wo <= 124)
  return p.x_sw == 16 && p.kh == 4 && p.kw == 1 && p.Cin_p == 64 && p.Ktot == 256 && p.sh == 1 && p.sw == 1 && p.dh == 1 && p.pt == 0 && p.pl == 0 &&
         p.x_sh % 16 == 0 && p.x_sn % p.x_sh == 0 && p.Cout_alloc == 64 && p.ldy == p.Cout_p && p.Cout_p <= 64 && p.Cout_p % 8 == 0 && p.Wo % 4 == 0 && p.Wo / 4 < S2D_G && p.res == nullptr;
}

}  // namespace

// Stem + max-pool fusion (ConvParams::pool): whole 4-row tiles, pooled rows of at most 64 pixels (one staging slab per output row), a ReLU
// as the only activation (the pooled tensor is zero-filled and max-reduced into), four staging slabs (two tiles x 4 rows x 8 KB).
bool conv_tc3_s2d_pool_ok(const ConvParams& p) {
  const char* e = getenv("B200_FUSE_POOL");
  if (e && e[0] == '0') return false;
  ConvParams q = p;
  q.pool = 0;
  if (!s2d_ok(q)) return false;
  return p.Ho % 4 == 0 && p.Wo % 4 == 0 && p.Wo / 2 <= 64 && p.act1 == ACT_RELU && p.act2 == ACT_NONE && p.res == nullptr &&
         plan3(p.Cout_alloc, false, (p.Ktot / BK) * (p.Cout_alloc / 2) * 128, A_S2D, 0, p.Cout_alloc, p.Ktot / BK).nout == 4;
}

namespace {

int a_mode_of(const ConvParams& p) {
  if (s2d_ok(p)) return A_S2D;
  if (p.K2 > 0) return (p.kh == 1 && p.kw == 1 && p.sh == 1 && p.sw == 1 && p.pt == 0 && p.pl == 0 && p.Cin_p % 64 == 0 && p.K2 % 64 == 0 && p.x_sw == 0) ? A_MATRIX : 0;
  if (conv_tc2_a_is_matrix(p) && p.Cin_p % 64 == 0) return A_MATRIX;
  if (halo_ok(p)) return A_HALO;
  if (conv_tc2_a_is_im2col(p)) return A_IM2COL;
  return 0;
}

bool ws_applies(const ConvParams& p, int bn, int a_mode) {
  const int ws_bytes = (p.Ktot / BK) * (bn / 2) * 128;
  const char* no_ws = getenv("B200_NO_WS");
  // halo mode: one A stage feeds nine k-blocks (2300+ cycles of MMA), two stages are a full double buffer
  const char* big = getenv("B200_TC3_WS_BIG");
  const bool big_ok = !(big && big[0] == '0');
  if (!big_ok && ws_bytes > 80 * 1024) return false;
  // (im2col mode with 144 KB of resident weights and three A stages measured slower than the streamed ring: res3_0_b 83 -> 88+ us)
  const int min_sa = a_mode == A_HALO ? (big_ok ? 2 : 3) : 4;
  // One N tile: the whole weight matrix is resident.  Several N tiles: with a cluster count that is a multiple of the N tile count every
  // cluster sees ONE fixed N tile (tile = cluster + k * clusters, n = tile % n_tiles), so its half of THAT tile's weights can be resident
  // too (launch_tc3 rounds the grid accordingly).  B200_TC3_WSN=0 keeps the single-N-tile rule.
  const int n_tiles = p.Cout_alloc / bn;
  const char* wsn = getenv("B200_TC3_WSN");
  const bool multi_ok = !(wsn && wsn[0] == '0') && n_tiles >= 2 && n_tiles <= 8 && pair_slots() / n_tiles >= 8;
  return (n_tiles == 1 || multi_ok) && ws_bytes <= WS_MAX_BYTES &&
         plan3(bn, p.res != nullptr, ws_bytes, a_mode, rows_pitch(p), p.Cout_alloc, p.Ktot / BK).sa >= (n_tiles > 1 ? 3 : min_sa) &&
         !(no_ws && no_ws[0] == '1');
}

template <typename T>
void launch_tc3(const ConvParams& p, bool bf16, int bn, int a_mode, cudaStream_t s) {
  Tc3Args a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  { const char* e = getenv("B200_DEBUG_NO_RES"); if (e && e[0] == '1') a.p.res = nullptr; }   // timing experiment only: wrong results
  { const char* e = getenv("B200_DEBUG_TC3"); a.debug = e ? atoi(e) : 0; }                        // timing experiments only: wrong results
  a.e1 = make_epi(p.act1, p.a1, p.b1);
  a.e2 = make_epi(p.act2, p.a2, p.b2);
  a.gen = conv_acts_are_clamps(p) ? 0 : 1;
  {
    // a trailing ReLU rides on the float -> bf16/f16 conversion (B200_RELU_PACK=0: explicit max)
    const char* e = getenv("B200_RELU_PACK");
    const bool on = !(e && e[0] == '0');
    if (on && p.res != nullptr && p.act2 == ACT_RELU) a.e2.relu_in_pack = 1;
    if (on && p.res == nullptr && p.act1 == ACT_RELU) a.e1.relu_in_pack = 1;
  }
  a.n_tiles = p.Cout_alloc / bn;
  a.reverse = p.reverse;
  { const char* e = getenv("B200_TC3_WARP_STORE"); a.warp_store = (e && e[0] == '0') ? 0 : 1; }
  {
    // TMEM accumulator ring: as many accumulators as 512 columns hold (at most 4), so that the MMAs of the next tiles never wait for the
    // epilogue's latency chain of an earlier one.  B200_TC3_NACC=2 restores the double buffer, B200_TC3_EARLY=0 the late hand-back.
    const int acc_cols = a_mode == A_S2D ? 4 * bn : bn;
    int nacc = std::min(MAX_ACC, 512 / acc_cols);
    const char* e = getenv("B200_TC3_NACC");
    if (e && atoi(e) >= 2 && atoi(e) < nacc) nacc = atoi(e) >= 4 ? 4 : 2;
    a.acc_log2 = nacc >= 4 ? 2 : 1;
    const char* er = getenv("B200_TC3_EARLY");
    a.early_release = (er && er[0] == '0') ? 0 : 1;
  }
  { const char* e = getenv("B200_TC3_PREFETCH"); a.prefetch_tiles = e ? atoi(e) : 0; }   // measured: no gain on B200 (profiles/r1 call k)
  {
    // activations are read once per layer (the next reader is a different kernel, long after the tensor has left the 126 MB L2 at batch 256):
    // let them leave first so that what this kernel WRITES stays resident for its consumer.  B200_TC3_EVICT=0 turns the hints off.
    const char* e = getenv("B200_TC3_EVICT");
    const bool on = !(e && e[0] == '0');
    a.a_policy = on ? ptx::L2_EVICT_FIRST : ptx::L2_EVICT_NORMAL;
    a.r_policy = on ? ptx::L2_EVICT_FIRST : ptx::L2_EVICT_NORMAL;
  }
  int cta_tiles_m;
  if (a_mode == A_S2D) {
    a.nseg = (p.Wo / 4 + S2D_G - 1) / S2D_G;
    a.tiles_per_img = ((p.Ho + S2D_R - 1) / S2D_R) * a.nseg;
    cta_tiles_m = a.tiles_per_img * p.N;
  } else if (a_mode == A_HALO) {
    a.rows_p = rows_pitch(p);
    a.rows_r = BM / a.rows_p;
    a.tiles_per_img = (p.Ho + a.rows_r - 1) / a.rows_r;
    cta_tiles_m = a.tiles_per_img * p.N;
  } else {
    cta_tiles_m = (p.M + BM - 1) / BM;
  }
  a.num_tiles = ((cta_tiles_m + 1) / 2) * a.n_tiles;
  const bool ws = a_mode == A_S2D ? true : ws_applies(p, bn, a_mode);
  a.tm_w = make_tmap_2d(p.w, bf16, (uint64_t)p.Ktot, (uint64_t)p.Cout_alloc, (uint64_t)p.w_ld * 2, BK, bn / 2);
  if (a_mode == A_S2D) {
    // input: block pixels in groups of 4 (64 elements = one 128-byte row): {64, ceil(Wb / 4), Hb, N}; the last group of a row may run one
    // pixel into the next row (read only by output pixels the store drops)
    const uint64_t Wb = (uint64_t)(p.x_sh / 16), Hb = (uint64_t)(p.x_sn / p.x_sh), pitch = (uint64_t)p.ldy * 2;
    const uint64_t ad[4] = {64, (Wb + 3) / 4, Hb, (uint64_t)p.N};
    const uint64_t as[3] = {128, Wb * 32, Hb * Wb * 32};
    const uint32_t ab[4] = {64, (uint32_t)S2D_G, (uint32_t)(S2D_R + 3), 1};
    a.tm_a = make_tmap_tiled(p.x, bf16, 4, ad, as, ab);
    // output: {channel, phase (pixel 4 g + f), pixel group g, row, image}
    const uint64_t yd[5] = {(uint64_t)p.Cout_p, 4, (uint64_t)(p.Wo / 4), (uint64_t)p.Ho, (uint64_t)p.N};
    const uint64_t ys[4] = {pitch, 4 * pitch, (uint64_t)p.Wo * pitch, (uint64_t)p.Ho * p.Wo * pitch};
    const uint32_t yb[5] = {64, 1, (uint32_t)S2D_G, (uint32_t)S2D_R, 1};
    a.tm_y = make_tmap_tiled(p.y, bf16, 5, yd, ys, yb);
    a.tm_r = a.tm_y;
    const uint32_t ywb[5] = {64, 1, (uint32_t)S2D_G, 1, 1};         // one warp = one output row of the tile
    a.tm_yw = make_tmap_tiled(p.y, bf16, 5, yd, ys, ywb);
    if (p.pool) {
      // y is the pooled tensor {channel, Wo / 2, Ho / 2, image}; one box = one pooled row of 64 pixel slots (clipped at Wo / 2)
      B200_CHECK(conv_tc3_s2d_pool_ok(p), B200_ERR_INVALID, "conv_tc3: stem + max-pool fusion does not apply to this launch");
      const uint64_t pd[4] = {(uint64_t)p.Cout_p, (uint64_t)(p.Wo / 2), (uint64_t)(p.Ho / 2), (uint64_t)p.N};
      const uint64_t ps[3] = {pitch, (uint64_t)(p.Wo / 2) * pitch, (uint64_t)(p.Ho / 2) * (p.Wo / 2) * pitch};
      const uint32_t pb[4] = {64, 64, 1, 1};
      a.tm_y = make_tmap_tiled(p.y, bf16, 4, pd, ps, pb);
      a.tm_yw = a.tm_r = a.tm_y;
      a.pool = 1;
    }
  } else if (a_mode == A_HALO) {
    const uint64_t C = (uint64_t)p.Cin_p, pitch = (uint64_t)p.ldy * 2;
    const uint64_t ad[4] = {C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    const uint64_t as[3] = {C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2};
    const uint32_t ab[4] = {64, (uint32_t)a.rows_p, (uint32_t)(a.rows_r + 2), 1};
    a.tm_a = make_tmap_tiled(p.x, bf16, 4, ad, as, ab);
    // output / residual: {channels, w, h, n}; pixels beyond Wo and rows beyond Ho (last tile of an image when Ho % R != 0) are clipped
    // on store and zero-filled on load
    const uint64_t yd[4] = {(uint64_t)p.Cout_p, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    const uint64_t ys[3] = {pitch, (uint64_t)p.Wo * pitch, (uint64_t)p.Ho * p.Wo * pitch};
    const uint32_t yb[4] = {64, (uint32_t)a.rows_p, (uint32_t)a.rows_r, 1};
    a.tm_y = make_tmap_tiled(p.y, bf16, 4, yd, ys, yb);
    a.tm_r = p.res ? make_tmap_tiled(p.res, bf16, 4, yd, ys, yb) : a.tm_y;
    // one warp = 32 consecutive rows of the tile = 32 pixels of one image row (P >= 32) or 32 / P whole image rows
    const uint32_t ywb[4] = {64, (uint32_t)std::min(a.rows_p, 32), (uint32_t)std::max(1, 32 / a.rows_p), 1};
    a.tm_yw = make_tmap_tiled(p.y, bf16, 4, yd, ys, ywb);
  } else {
    a.tm_y = make_tmap_2d(p.y, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM);
    a.tm_yw = make_tmap_2d(p.y, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, 32);
    if (p.y_heads_s) {
      B200_CHECK(a_mode == A_MATRIX && p.y_heads_s % 32 == 0 && p.M % p.y_heads_s == 0 && p.Cout_p == 64 * p.y_heads_h && !p.res, B200_ERR_INVALID,
                 "conv_tc3: bad head-split output geometry");
      const uint64_t S = (uint64_t)p.y_heads_s, H = (uint64_t)p.y_heads_h;
      const uint64_t yd[4] = {64, S, H, (uint64_t)p.M / S};
      const uint64_t ys[3] = {128, S * 128, H * S * 128};
      const uint32_t ywb[4] = {64, 32, 1, 1};
      a.tm_yw = make_tmap_tiled(p.y, bf16, 4, yd, ys, ywb);
      a.warp_store = 1;
    }
    a.tm_r = p.res ? make_tmap_2d(p.res, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM) : a.tm_y;
    a.tm_a = a_mode == A_MATRIX ? make_tmap_2d(p.x, bf16, (uint64_t)p.Cin_p, (uint64_t)p.M, (uint64_t)p.Cin_p * 2, BK, BM) : make_tmap_im2col(p, bf16);
    if (p.K2 > 0) {
      B200_CHECK(a_mode == A_MATRIX && p.K2 % BK == 0 && p.Cin_p + p.K2 == p.Ktot, B200_ERR_INVALID, "conv_tc3: bad second-source geometry");
      ConvParams p2;                     // the shortcut as a 1x1 strided convolution over x2, same output pixels
      p2.x = p.x2; p2.N = p.N; p2.H = p.H2; p2.W = p.W2; p2.Cin_p = p.K2; p2.Ho = p.Ho; p2.Wo = p.Wo;
      p2.kh = p2.kw = 1; p2.sh = p.sh2; p2.sw = p.sw2; p2.dh = p2.dw = 1; p2.pt = p2.pl = 0;
      a.tm_a2 = make_tmap_im2col(p2, bf16);
    }
  }
  a.nkb1 = (p.K2 > 0 ? p.Cin_p : p.Ktot) / BK;
  int clusters = std::min(a.num_tiles, pair_slots());
  if (ws && a.n_tiles > 1) clusters -= clusters % a.n_tiles;      // resident weights of one fixed N tile per cluster (ws_applies)
  const int grid = 2 * clusters;
  const unsigned smem = plan3(bn, a.p.res != nullptr, ws ? (p.Ktot / BK) * (bn / 2) * 128 : 0, a_mode, a.rows_p, p.Cout_alloc, p.Ktot / BK).total;
  launch_tc3_typed<T>(a, bn, a_mode, ws, grid, smem, s);
}

}  // namespace

bool conv_tc3_supported(const ConvParams& p) {
  return conv_tc2_supported(p) && p.w_grp_m == 0 && !p.y_heads_inv && a_mode_of(p) != 0 && p.Cout_alloc % 64 == 0 && p.Ktot % BK == 0 && p.Cout_alloc <= BIAS_SMEM_FLOATS;
}

// Estimated cycles of one launch with N tile `bn` on `sms` SMs: rounds x k-blocks x max(MMA, operand bytes / fabric rate) + epilogue traffic.
static double tc3_cost(const ConvParams& p, int bn, int a_mode, int sms) {
  const double fabric = 42.0;                          // B/clk per SM (6.3 KB/clk chip-wide)
  int cta_m;
  double a_scale = 1.0;
  if (a_mode == A_HALO) {
    const int P = rows_pitch(p);
    cta_m = ((p.Ho + BM / P - 1) / (BM / P)) * p.N;
    a_scale = (BM / P + 2.0) / (BM / P) / (p.kh * p.kw);
  } else {
    cta_m = (p.M + BM - 1) / BM;
  }
  const int tiles = ((cta_m + 1) / 2) * (p.Cout_alloc / bn);
  const int rounds = (tiles + sms / 2 - 1) / (sms / 2);
  const int nkb = p.Ktot / BK;
  const bool ws = p.Cout_alloc == bn && nkb * (bn / 2) * 128 <= WS_MAX_BYTES;
  const double bytes_kb = A_BYTES * a_scale + (ws ? 0.0 : (bn / 2) * 128.0);
  const double mma_kb = bn * 2.0;                      // 128 x bn x 64 per CTA at 4096 MAC/clk
  const double epi = (p.res ? 2.0 : 1.0) * BM * bn * 2.0 / fabric;
  return rounds * (nkb * std::max(mma_kb, bytes_kb / fabric) + epi);
}

int conv_tc3_pick_bn(const ConvParams& p) {
  const int a_mode = a_mode_of(p);
  if (a_mode == A_S2D) return p.Cout_alloc;   // weights resident, one N tile
  { const char* e = getenv("B200_TC3_FORCE_BN"); const int f = e ? atoi(e) : 0; if ((f == 64 || f == 128 || f == 256) && p.Cout_alloc % f == 0) return f; }   // tests
  const int sms = 2 * pair_slots();
  int best = 0;
  double best_c = 0;
  for (int bn : {256, 128, 64}) {
    if (p.Cout_alloc % bn) continue;
    const double c = tc3_cost(p, bn, a_mode, sms);
    if (!best || c < best_c) { best = bn; best_c = c; }
  }
  return best;
}

// auto policy: the pair kernel pays off once every SM pair has a tile to walk
bool conv_tc3_preferred(const ConvParams& p) {
  const char* e = getenv("B200_TC3");
  if (e && e[0] == '0') return false;
  if (!conv_tc3_supported(p)) return false;
  if (e && e[0] == '1') return true;
  const int bn = conv_tc3_pick_bn(p);
  const int sms = 2 * pair_slots();
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * (p.Cout_alloc / bn);
  return tiles >= sms / 2;
}

const char* conv_tc3_mode_name(const ConvParams& p) {
  switch (a_mode_of(p)) {
    case A_MATRIX: return p.K2 > 0 ? "tma_kcat" : (p.y_heads_s > 0 ? "tma_heads" : "tma");
    case A_HALO: return "halo";
    case A_IM2COL: return "im2col";
    case A_S2D: return "s2d";
  }
  return "?";
}

void conv_tc3_f16_opt_in();   // conv_tc3_f16.cu

void conv_tc3_init_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count3, cudaDevAttrMultiProcessorCount, dev);
  tc3::opt_in_tc3_typed<__nv_bfloat16>();
  conv_tc3_f16_opt_in();
  {
    // a persistent grid larger than the number of co-resident pairs would run its last clusters as a second wave
    auto k = tc3::conv_tc3_kernel<__nv_bfloat16, 256, A_MATRIX, false, false>;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)g_sm_count3, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_LIMIT;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, k, &cfg) == cudaSuccess && nc > 0) g_max_clusters3 = std::min(nc, g_sm_count3 / 2);
    else cudaGetLastError();
  }
}

void launch_conv_tc3_f16(const ConvParams& p, int bn, int a_mode, cudaStream_t s);   // conv_tc3_f16.cu

void launch_conv_tc3(const ConvParams& p, int dtype, int bn, cudaStream_t s) {
  B200_CHECK(conv_tc3_supported(p), B200_ERR_UNSUPPORTED, "conv_tc3: unsupported launch");
  B200_CHECK(p.Cout_alloc % bn == 0 && p.Cout_p % 8 == 0 && p.ldy % 8 == 0, B200_ERR_INVALID, "conv_tc3: bad geometry");
  const bool bf16 = dtype == B200_BF16;
  B200_CHECK(bf16 || dtype == B200_F16, B200_ERR_UNSUPPORTED, "conv_tc3: dtype must be bf16/f16");
  const int a_mode = a_mode_of(p);
  if (bf16) launch_tc3<__nv_bfloat16>(p, true, bn, a_mode, s);
  else launch_tc3<__half>(p, false, bn, a_mode, s);
}

}  // namespace b200
