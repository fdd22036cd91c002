// NonMaxSuppression-5 and DetectionOutput-1 on the GPU, fp32 + i32, reproducing the reference
// ordering rules bit for bit (compiled with -fmad=false so the box arithmetic is not contracted):
//   NMS:  inference-engine/src/mkldnn_plugin/nodes/mkldnn_non_max_suppression_node.cpp:134-238 (execute),
//         :240-273 (intersectionOverUnion), :336-376 (nmsWithoutSoftSigma)
//   DetectionOutput: nodes/mkldnn_detection_output_node.cpp:129-352 (execute), :366-401 (JaccardOverlap),
//         :403-503 (decodeBBoxes), :505-536 (nms_cf)
// Candidates are ordered by 64-bit keys (orderable score bits << 32 | ~index): score descending, index ascending; keys are distinct.
// DetectionOutput fast path (top_k <= 512, the SSD case): confidences are transposed to [image][class][prior] once (coalesced, as
// the reference reorders them), then one block per (image, class) radix-SELECTS the top_k keys (8-bit digits, shared-memory
// histograms) instead of sorting all priors, rank-sorts those <= 512 keys, builds the pairwise-overlap BITMASK of the selected boxes
// in parallel and replays the greedy suppression as a serial scan over that mask (a box is dropped iff an earlier KEPT box overlaps
// it: the same recurrence as nms_cf, the overlaps computed by the same expression).  keep_top_k per image is a radix select too.
// NMS-5 and the DetectionOutput fall-back (top_k = -1 or > 512) sort all candidates with a block-wide bitonic sort in shared memory
// and run the greedy pass on one warp, the kept list checked 32 boxes at a time.
#include <algorithm>
#include <cfloat>

#include "common.h"
#include "kernels.h"

namespace b200 {
namespace {

constexpr int SORT_MAX = 16384;                 // keys per block-level sort in shared memory
constexpr int SORT_SMEM = SORT_MAX * 8;         // 128 KB

__device__ __forceinline__ uint32_t ord_f32(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) { return ((uint64_t)ord_f32(score) << 32) | (uint64_t)(0xFFFFFFFFu - idx); }
__device__ __forceinline__ uint32_t key_idx(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu); }

// sorts keys[0..n) descending; n must be a power of two; all threads of the block participate
__device__ void bitonic_sort_desc(uint64_t* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t a = keys[i], b = keys[ixj];
          bool desc_seg = (i & k) == 0;
          if ((a < b) == desc_seg) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}
inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---------------------------------------------------------------------------------------------
// NMS-5
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float nms_iou(const float* bi, const float* bj, int center) {
  float yminI, xminI, ymaxI, xmaxI, yminJ, xminJ, ymaxJ, xmaxJ;
  if (center) {
    yminI = bi[1] - bi[3] / 2.f; xminI = bi[0] - bi[2] / 2.f; ymaxI = bi[1] + bi[3] / 2.f; xmaxI = bi[0] + bi[2] / 2.f;
    yminJ = bj[1] - bj[3] / 2.f; xminJ = bj[0] - bj[2] / 2.f; ymaxJ = bj[1] + bj[3] / 2.f; xmaxJ = bj[0] + bj[2] / 2.f;
  } else {
    yminI = fminf(bi[0], bi[2]); xminI = fminf(bi[1], bi[3]); ymaxI = fmaxf(bi[0], bi[2]); xmaxI = fmaxf(bi[1], bi[3]);
    yminJ = fminf(bj[0], bj[2]); xminJ = fminf(bj[1], bj[3]); ymaxJ = fmaxf(bj[0], bj[2]); xmaxJ = fmaxf(bj[1], bj[3]);
  }
  float areaI = (ymaxI - yminI) * (xmaxI - xminI);
  float areaJ = (ymaxJ - yminJ) * (xmaxJ - xminJ);
  if (areaI <= 0.f || areaJ <= 0.f) return 0.f;
  float inter = fmaxf(fminf(ymaxI, ymaxJ) - fmaxf(yminI, yminJ), 0.f) * fmaxf(fminf(xmaxI, xmaxJ) - fmaxf(xminI, xminJ), 0.f);
  return inter / (areaI + areaJ - inter);
}

// one block per (batch, class)
__global__ void nms_select_kernel(NmsParams p, int n_pad, int32_t* sel, int32_t* cnt) {
  extern __shared__ uint64_t nms_keys[];
  __shared__ int s_count;
  const int b = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const float* scores = p.scores + ((int64_t)b * p.C + c) * p.Nb;
  const float* boxes = p.boxes + (int64_t)b * p.Nb * 4;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    uint64_t k = 0;
    if (i < p.Nb && scores[i] > p.score_thr) { k = make_key(scores[i], (uint32_t)i); ++local; }
    nms_keys[i] = k;
  }
  atomicAdd(&s_count, local);
  __syncthreads();
  bitonic_sort_desc(nms_keys, n_pad);
  const int count = s_count;
  int32_t* my_sel = sel + (int64_t)blockIdx.x * p.max_out;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int nsel = 0;
    for (int i = 0; i < count && nsel < p.max_out; ++i) {
      const int idx = (int)key_idx(nms_keys[i]);
      bool sup = false;
      for (int k = lane; k < nsel; k += 32)
        if (nms_iou(boxes + (int64_t)idx * 4, boxes + (int64_t)my_sel[k] * 4, p.center_format) >= p.iou_thr) sup = true;
      sup = __any_sync(0xffffffffu, sup);
      if (!sup) {
        if (lane == 0) my_sel[nsel] = idx;
        ++nsel;
      }
      __syncwarp();
    }
    if (lane == 0) cnt[blockIdx.x] = nsel;
  }
}

// Soft-NMS (soft_nms_sigma > 0): nodes/mkldnn_non_max_suppression_node.cpp:275-334.  The reference pops a priority queue (score
// descending, index ascending), decays the popped score by exp(scale * iou^2) against the boxes selected since the box was last looked
// at (suppress_begin_index), and either selects the box (score unchanged), re-inserts it (decayed, still above the threshold) or drops it.
// One block per (batch, class): the pop is a block-wide arg-max over the current scores (kept in the workspace), the inherently serial
// decay loop runs on thread 0.
__global__ void nms_soft_select_kernel(NmsParams p, float scale, float* cur, int32_t* sbi, int32_t* sel, float* sel_sc, int32_t* cnt) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_top;
  __shared__ int s_nsel;
  const int b = blockIdx.x / p.C;
  const float* scores = p.scores + (int64_t)blockIdx.x * p.Nb;
  const float* boxes = p.boxes + (int64_t)b * p.Nb * 4;
  float* my_cur = cur + (int64_t)blockIdx.x * p.Nb;
  int32_t* my_sbi = sbi + (int64_t)blockIdx.x * p.Nb;
  int32_t* my_sel = sel + (int64_t)blockIdx.x * p.max_out;
  float* my_sc = sel_sc + (int64_t)blockIdx.x * p.max_out;
  for (int i = threadIdx.x; i < p.Nb; i += blockDim.x) {
    const float v = scores[i];
    my_cur[i] = v > p.score_thr ? v : -INFINITY;
    my_sbi[i] = 0;
  }
  if (threadIdx.x == 0) s_nsel = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  while (true) {
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < p.Nb; i += blockDim.x) {
      const float v = my_cur[i];
      if (v > -INFINITY) best = max(best, (unsigned long long)make_key(v, (uint32_t)i));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) s_warp[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long top = 0ull;
      for (int w = 0; w < nwarps; ++w) top = max(top, s_warp[w]);
      int nsel = s_nsel;
      if (top != 0ull && nsel < p.max_out) {
        const int idx = (int)key_idx(top);
        float score = my_cur[idx];
        const float orig = score;
        bool selected = true;
        for (int k = nsel - 1; k >= my_sbi[idx]; --k) {
          const float iou = nms_iou(boxes + (int64_t)idx * 4, boxes + (int64_t)my_sel[k] * 4, p.center_format);
          const float weight = iou <= p.iou_thr ? expf(scale * iou * iou) : 0.0f;
          score *= weight;
          if (iou >= p.iou_thr) { selected = false; break; }
          if (score <= p.score_thr) break;
        }
        my_sbi[idx] = nsel;
        if (selected && score == orig) {
          my_sel[nsel] = idx; my_sc[nsel] = score;
          s_nsel = nsel + 1;
          my_cur[idx] = -INFINITY;
        } else if (selected && score > p.score_thr) {
          my_cur[idx] = score;            // back into the queue with the decayed score
        } else {
          my_cur[idx] = -INFINITY;
        }
      } else {
        top = 0ull;
      }
      s_top = top;
    }
    __syncthreads();
    if (s_top == 0ull) break;
  }
  if (threadIdx.x == 0) cnt[blockIdx.x] = s_nsel;
}

// single block: gather, (optionally) global sort, write padded outputs.  sel_sc: scores of the selections (soft-NMS decays them), or null
__global__ void nms_finalize_kernel(NmsParams p, const int32_t* sel, const int32_t* cnt, uint64_t* keys, int n_pad, int64_t out_rows, const float* sel_sc,
                                    float* out_sc) {
  __shared__ int s_total;
  const int BC = p.B * p.C;
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < BC; ++i) t += cnt[i];
    s_total = t;
  }
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) keys[i] = 0;
  __syncthreads();
  const int total = s_total;
  // position of every selected box in (batch, class, selection order); key low word keeps that order when unsorted
  for (int bc = threadIdx.x; bc < BC; bc += blockDim.x) {
    int base = 0;
    for (int i = 0; i < bc; ++i) base += cnt[i];
    for (int k = 0; k < cnt[bc]; ++k) {
      int box = sel[(int64_t)bc * p.max_out + k];
      float sc = sel_sc ? sel_sc[(int64_t)bc * p.max_out + k] : p.scores[(int64_t)bc * p.Nb + box];
      if (out_sc) out_sc[(int64_t)bc * p.Nb + box] = sc;       // decayed score, addressable by (batch, class, box) after the sort
      uint32_t lin = (uint32_t)((int64_t)bc * p.Nb + box);
      keys[base + k] = p.sort_desc ? make_key(sc, lin) : (((uint64_t)(0xFFFFFFFFu - (uint32_t)(base + k))) << 32) | (uint64_t)lin;
    }
  }
  __syncthreads();
  if (p.sort_desc) bitonic_sort_desc(keys, n_pad);
  const int valid = (int)min((int64_t)total, out_rows);
  for (int64_t r = threadIdx.x; r < out_rows; r += blockDim.x) {
    if (r < valid) {
      uint32_t lin = p.sort_desc ? key_idx(keys[r]) : (uint32_t)(keys[r] & 0xFFFFFFFFu);
      int box = (int)(lin % (uint32_t)p.Nb);
      int bc = (int)(lin / (uint32_t)p.Nb);
      int bb = bc / p.C, cc = bc % p.C;
      p.sel_idx[r * 3 + 0] = bb; p.sel_idx[r * 3 + 1] = cc; p.sel_idx[r * 3 + 2] = box;
      p.sel_scores[r * 3 + 0] = (float)bb; p.sel_scores[r * 3 + 1] = (float)cc;
      p.sel_scores[r * 3 + 2] = out_sc ? out_sc[(int64_t)bc * p.Nb + box] : p.scores[(int64_t)bc * p.Nb + box];
    } else {
      p.sel_idx[r * 3 + 0] = -1; p.sel_idx[r * 3 + 1] = -1; p.sel_idx[r * 3 + 2] = -1;
      p.sel_scores[r * 3 + 0] = -1.f; p.sel_scores[r * 3 + 1] = -1.f; p.sel_scores[r * 3 + 2] = -1.f;
    }
  }
  if (threadIdx.x == 0) p.valid[0] = valid;
}

// ---------------------------------------------------------------------------------------------
// DetectionOutput
// ---------------------------------------------------------------------------------------------
__global__ void detout_decode_kernel(DetOutParams p, float* boxes, float* sizes) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)p.N * p.num_loc_classes * p.num_priors;
  if (i >= total) return;
  const int pr = (int)(i % p.num_priors);
  const int c = (int)((i / p.num_priors) % p.num_loc_classes);
  const int n = (int)(i / ((int64_t)p.num_priors * p.num_loc_classes));
  const int prior_size = p.normalized ? 4 : 5, offs = p.normalized ? 0 : 1;
  const float* ppriors = p.priors;
  const float* pvar = p.priors + (int64_t)p.num_priors * prior_size;
  if (p.priors_batch > 1) {
    ppriors += p.var_in_target ? (int64_t)n * p.num_priors * prior_size : (int64_t)2 * n * p.num_priors * prior_size;
    pvar += p.var_in_target ? 0 : (int64_t)2 * n * p.num_priors * prior_size;
  }
  const float* loc = p.loc + (int64_t)n * 4 * p.num_loc_classes * p.num_priors + (int64_t)4 * pr * p.num_loc_classes + (int64_t)c * 4;
  float pxmin = ppriors[pr * prior_size + 0 + offs], pymin = ppriors[pr * prior_size + 1 + offs];
  float pxmax = ppriors[pr * prior_size + 2 + offs], pymax = ppriors[pr * prior_size + 3 + offs];
  float lxmin = loc[0], lymin = loc[1], lxmax = loc[2], lymax = loc[3];
  if (!p.normalized) {
    pxmin /= (float)p.input_w; pymin /= (float)p.input_h; pxmax /= (float)p.input_w; pymax /= (float)p.input_h;
  }
  float nxmin = 0.f, nymin = 0.f, nxmax = 0.f, nymax = 0.f;
  if (!p.code_center) {
    if (p.var_in_target) { nxmin = pxmin + lxmin; nymin = pymin + lymin; nxmax = pxmax + lxmax; nymax = pymax + lymax; }
    else {
      nxmin = pxmin + pvar[pr * 4 + 0] * lxmin; nymin = pymin + pvar[pr * 4 + 1] * lymin;
      nxmax = pxmax + pvar[pr * 4 + 2] * lxmax; nymax = pymax + pvar[pr * 4 + 3] * lymax;
    }
  } else {
    float pw = pxmax - pxmin, ph = pymax - pymin;
    float pcx = (pxmin + pxmax) / 2.0f, pcy = (pymin + pymax) / 2.0f;
    float cx, cy, w, h;
    if (p.var_in_target) {
      cx = lxmin * pw + pcx; cy = lymin * ph + pcy; w = expf(lxmax) * pw; h = expf(lymax) * ph;
    } else {
      cx = pvar[pr * 4 + 0] * lxmin * pw + pcx; cy = pvar[pr * 4 + 1] * lymin * ph + pcy;
      w = expf(pvar[pr * 4 + 2] * lxmax) * pw; h = expf(pvar[pr * 4 + 3] * lymax) * ph;
    }
    nxmin = cx - w / 2.0f; nymin = cy - h / 2.0f; nxmax = cx + w / 2.0f; nymax = cy + h / 2.0f;
  }
  if (p.clip_before) {
    nxmin = fmaxf(0.f, fminf(1.f, nxmin)); nymin = fmaxf(0.f, fminf(1.f, nymin));
    nxmax = fmaxf(0.f, fminf(1.f, nxmax)); nymax = fmaxf(0.f, fminf(1.f, nymax));
  }
  // decoded layout [N][loc_class][prior][4]
  float* o = boxes + (((int64_t)n * p.num_loc_classes + c) * p.num_priors + pr) * 4;
  o[0] = nxmin; o[1] = nymin; o[2] = nxmax; o[3] = nymax;
  sizes[((int64_t)n * p.num_loc_classes + c) * p.num_priors + pr] = (nxmax - nxmin) * (nymax - nymin);
}

__device__ __forceinline__ float jaccard(const float* bb, const float* sz, int i1, int i2) {
  float xmin1 = bb[i1 * 4 + 0], ymin1 = bb[i1 * 4 + 1], xmax1 = bb[i1 * 4 + 2], ymax1 = bb[i1 * 4 + 3];
  float xmin2 = bb[i2 * 4 + 0], ymin2 = bb[i2 * 4 + 1], xmax2 = bb[i2 * 4 + 2], ymax2 = bb[i2 * 4 + 3];
  if (xmin2 > xmax1 || xmax2 < xmin1 || ymin2 > ymax1 || ymax2 < ymin1) return 0.f;
  float iw = fminf(xmax1, xmax2) - fmaxf(xmin1, xmin2);
  float ih = fminf(ymax1, ymax2) - fmaxf(ymin1, ymin2);
  if (iw <= 0.f || ih <= 0.f) return 0.f;
  float inter = iw * ih;
  return inter / (sz[i1] + sz[i2] - inter);
}

// one block per (image, class): candidate filter, sort, top_k, greedy NMS (nms_cf)
__global__ void detout_nms_kernel(DetOutParams p, const float* boxes, const float* sizes, int n_pad, int max_keep, int32_t* kept, int32_t* cnt) {
  extern __shared__ uint64_t det_keys[];
  __shared__ int s_count;
  const int n = blockIdx.x / p.num_classes, c = blockIdx.x % p.num_classes;
  if (c == p.background) { if (threadIdx.x == 0) cnt[blockIdx.x] = 0; return; }
  const float* conf = p.conf + (int64_t)n * p.num_priors * p.num_classes;   // [prior][class]
  const int lc = p.share_location ? 0 : c;
  const float* bb = boxes + ((int64_t)n * p.num_loc_classes + lc) * p.num_priors * 4;
  const float* sz = sizes + ((int64_t)n * p.num_loc_classes + lc) * p.num_priors;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    uint64_t k = 0;
    if (i < p.num_priors) {
      float s = conf[(int64_t)i * p.num_classes + c];
      if (s > p.conf_thr) { k = make_key(s, (uint32_t)i); ++local; }
    }
    det_keys[i] = k;
  }
  atomicAdd(&s_count, local);
  __syncthreads();
  bitonic_sort_desc(det_keys, n_pad);
  const int count = s_count;
  const int num_scores = p.top_k == -1 ? count : min(p.top_k, count);
  int32_t* my = kept + (int64_t)blockIdx.x * max_keep;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int det = 0;
    for (int i = 0; i < num_scores; ++i) {
      const int idx = (int)key_idx(det_keys[i]);
      bool sup = false;
      for (int k = lane; k < det; k += 32)
        if (jaccard(bb, sz, idx, my[k]) > p.nms_thr) sup = true;
      sup = __any_sync(0xffffffffu, sup);
      if (!sup) {
        if (lane == 0) my[det] = idx;
        ++det;
      }
      __syncwarp();
    }
    if (lane == 0) cnt[blockIdx.x] = det;
  }
}

// one block per image: keep_top_k across classes (global sort by score), lists re-stored per class
__global__ void detout_keep_kernel(DetOutParams p, int n_pad, int max_keep, int32_t* kept, int32_t* cnt, int32_t* totals) {
  extern __shared__ uint64_t keep_keys[];
  __shared__ int s_total;
  const int n = blockIdx.x;
  int32_t* my_cnt = cnt + (int64_t)n * p.num_classes;
  int32_t* my_kept = kept + (int64_t)n * p.num_classes * max_keep;
  const float* conf = p.conf + (int64_t)n * p.num_priors * p.num_classes;
  if (threadIdx.x == 0) {
    int t = 0;
    for (int c = 0; c < p.num_classes; ++c) t += my_cnt[c];
    s_total = t;
  }
  __syncthreads();
  const int total = s_total;
  if (!(p.keep_top_k > -1 && total > p.keep_top_k)) {
    if (threadIdx.x == 0) totals[n] = total;
    return;
  }
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) keep_keys[i] = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < p.num_classes; c += blockDim.x) {
    int base = 0;
    for (int k = 0; k < c; ++k) base += my_cnt[k];
    for (int k = 0; k < my_cnt[c]; ++k) {
      int idx = my_kept[(int64_t)c * max_keep + k];
      keep_keys[base + k] = make_key(conf[(int64_t)idx * p.num_classes + c], (uint32_t)(c * p.num_priors + idx));
    }
  }
  __syncthreads();
  bitonic_sort_desc(keep_keys, n_pad);
  if (threadIdx.x == 0) {
    for (int c = 0; c < p.num_classes; ++c) my_cnt[c] = 0;
    for (int j = 0; j < p.keep_top_k; ++j) {
      uint32_t lin = key_idx(keep_keys[j]);
      int label = (int)(lin / (uint32_t)p.num_priors), idx = (int)(lin % (uint32_t)p.num_priors);
      my_kept[(int64_t)label * max_keep + my_cnt[label]] = idx;
      my_cnt[label]++;
    }
    totals[n] = p.keep_top_k;
  }
}

// single block: rows [image_id, label, conf, xmin, ymin, xmax, ymax] grouped by image, class, list order
__global__ void detout_pack_kernel(DetOutParams p, const float* boxes, int max_keep, const int32_t* kept, const int32_t* cnt,
                                   const int32_t* totals) {
  __shared__ int s_count;
  for (int64_t i = threadIdx.x; i < p.out_rows * 7; i += blockDim.x) p.out[i] = 0.f;
  __syncthreads();
  for (int nc = threadIdx.x; nc < p.N * p.num_classes; nc += blockDim.x) {
    const int n = nc / p.num_classes, c = nc % p.num_classes;
    int base = 0;
    for (int i = 0; i < n; ++i) base += totals[i];
    for (int k = 0; k < c; ++k) base += cnt[n * p.num_classes + k];
    const float* conf = p.conf + (int64_t)n * p.num_priors * p.num_classes;
    const int lc = p.share_location ? 0 : c;
    const float* bb = boxes + ((int64_t)n * p.num_loc_classes + lc) * p.num_priors * 4;
    for (int k = 0; k < cnt[nc]; ++k) {
      int64_t row = base + k;
      if (row >= p.out_rows) break;
      int idx = kept[(int64_t)nc * max_keep + k];
      float xmin = bb[idx * 4 + 0], ymin = bb[idx * 4 + 1], xmax = bb[idx * 4 + 2], ymax = bb[idx * 4 + 3];
      if (p.clip_after) {
        xmin = fmaxf(0.f, fminf(1.f, xmin)); ymin = fmaxf(0.f, fminf(1.f, ymin));
        xmax = fmaxf(0.f, fminf(1.f, xmax)); ymax = fmaxf(0.f, fminf(1.f, ymax));
      }
      float* o = p.out + row * 7;
      o[0] = (float)n; o[1] = (float)(p.decrease_label_id ? c - 1 : c); o[2] = conf[(int64_t)idx * p.num_classes + c];
      o[3] = xmin; o[4] = ymin; o[5] = xmax; o[6] = ymax;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < p.N; ++i) t += totals[i];
    s_count = t;
    if (t < p.out_rows) p.out[(int64_t)t * 7] = -1.f;   // end-of-list marker
  }
}


// ---- fast path ------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;
constexpr int SEL_KMAX = 512;        // candidates that enter the greedy pass on the fast path (top_k)
constexpr int KEEP_KMAX = 1024;      // keep_top_k on the fast path

// conf [N][prior][class] -> [N][class][prior] (32 x 32 tiles through shared memory, both sides coalesced)
__global__ void detout_transpose_conf_kernel(const float* __restrict__ in, float* __restrict__ out, int P, int C) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  const float* src = in + (int64_t)n * P * C;
  float* dst = out + (int64_t)n * P * C;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int pr = p0 + r, c = c0 + threadIdx.x;
    if (pr < P && c < C) tile[r][threadIdx.x] = src[(int64_t)pr * C + c];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, pr = p0 + threadIdx.x;
    if (pr < P && c < C) dst[(int64_t)c * P + pr] = tile[threadIdx.x][r];
  }
}

struct SelectShared {
  unsigned hist[256];
  unsigned long long prefix;
  int need, done, count;
};

// Threshold T such that exactly min(k, nonzero) of the DISTINCT keys key_at(0..n) are >= T (0 = absent key).  MSB-first radix select
// with 8-bit digits; stops as soon as the bucket that holds the k-th key is needed whole.  Every thread of the block calls it.
template <typename KeyAt>
__device__ uint64_t radix_select_desc(KeyAt key_at, int n, int k, int nonzero, SelectShared& sh) {
  if (nonzero <= k) return 1ull;
  if (threadIdx.x == 0) { sh.prefix = 0ull; sh.need = k; sh.done = 0; }
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh.hist[i] = 0u;
    __syncthreads();
    const unsigned long long prefix = sh.prefix;
    // warp-aggregated histogram: scores of one class crowd into a few bins (softmax outputs share their exponent; saturated heads give
    // thousands of EQUAL scores), and 32 shared-memory atomics on one address serialise.  Lanes with the same digit elect one to add.
    const int lane = threadIdx.x & 31;
    for (int base = threadIdx.x - lane; base < n; base += blockDim.x) {
      const int i = base + lane;
      const uint64_t key = i < n ? key_at(i) : 0ull;
      const bool counts = key != 0ull && (shift == 56 || ((key ^ prefix) >> (shift + 8)) == 0ull);
      const unsigned digit = (unsigned)(key >> shift) & 255u;
      const unsigned voters = __ballot_sync(0xffffffffu, counts);
      if (counts) {
        const unsigned peers = __match_any_sync(voters, digit);
        if (lane == __ffs(peers) - 1) atomicAdd(&sh.hist[digit], (unsigned)__popc(peers));
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns bins 255-8l .. 248-8l (descending); find the bin in which the running count reaches `need`
      const int lane = threadIdx.x, top = 255 - 8 * lane;
      unsigned local[8], lsum = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) { local[b] = sh.hist[top - b]; lsum += local[b]; }
      unsigned incl = lsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned need = (unsigned)sh.need, before = incl - lsum;
      const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
      if (hit != 0u && lane == __ffs(hit) - 1) {
        unsigned cum = before;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (cum + local[b] >= need) {
            sh.prefix = prefix | ((unsigned long long)(top - b) << shift);
            sh.need = (int)(need - cum);
            sh.done = (local[b] == need - cum) ? 1 : 0;
            break;
          }
          cum += local[b];
        }
      }
    }
    __syncthreads();
    if (sh.done) break;
  }
  return sh.prefix;
}

// keys[0..m) (distinct) -> sorted[0..m) descending: every key counts the keys above it (m <= 1024: shared-memory broadcast reads)
__device__ void rank_sort_desc(const uint64_t* keys, uint64_t* sorted, int m) {
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const uint64_t mine = keys[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += keys[j] > mine ? 1 : 0;
    sorted[rank] = mine;
  }
}

// one block per (image, class): select top_k, rank sort, overlap bitmask, greedy scan
// dynamic shared memory: scores [P] | keys [kcap] | sorted [kcap] | box [kcap][4] | size [kcap] | mask [kcap][kcap/32]
__global__ void __launch_bounds__(SEL_THREADS) detout_nms_fast_kernel(DetOutParams p, const float* __restrict__ conf_t, const float* __restrict__ boxes,
                                                                      const float* __restrict__ sizes, int kcap, int max_keep, int32_t* kept, int32_t* cnt) {
  extern __shared__ __align__(16) unsigned char det_smem[];
  __shared__ SelectShared sh;
  const int n = blockIdx.x / p.num_classes, c = blockIdx.x % p.num_classes;
  if (c == p.background) { if (threadIdx.x == 0) cnt[blockIdx.x] = 0; return; }
  const int P = p.num_priors, words = kcap / 32;
  float* sc = reinterpret_cast<float*>(det_smem);
  uint64_t* keys = reinterpret_cast<uint64_t*>(det_smem + (((int64_t)P * 4 + 15) / 16) * 16);
  uint64_t* sorted = keys + kcap;
  float4* sbox = reinterpret_cast<float4*>(sorted + kcap);
  float* ssz = reinterpret_cast<float*>(sbox + kcap);
  unsigned* mask = reinterpret_cast<unsigned*>(ssz + kcap);
  const float* conf = conf_t + ((int64_t)n * p.num_classes + c) * P;
  const int lc = p.share_location ? 0 : c;
  const float* bb = boxes + ((int64_t)n * p.num_loc_classes + lc) * P * 4;
  const float* sz = sizes + ((int64_t)n * p.num_loc_classes + lc) * P;
  if (threadIdx.x == 0) sh.count = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float v = conf[i];
    sc[i] = v;
    local += v > p.conf_thr ? 1 : 0;
  }
  if (local) atomicAdd(&sh.count, local);
  __syncthreads();
  const int count = sh.count;
  const int m = min(p.top_k, count);        // top_k in [1, SEL_KMAX] on this path
  if (m == 0) { if (threadIdx.x == 0) cnt[blockIdx.x] = 0; return; }
  const float thr = p.conf_thr;
  auto key_at = [&](int i) -> uint64_t { const float v = sc[i]; return v > thr ? make_key(v, (uint32_t)i) : 0ull; };
  const uint64_t T = radix_select_desc(key_at, P, m, count, sh);
  if (threadIdx.x == 0) sh.count = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const uint64_t key = key_at(i);
    if (key >= T && key != 0ull) keys[atomicAdd(&sh.count, 1)] = key;
  }
  __syncthreads();
  rank_sort_desc(keys, sorted, m);
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int idx = (int)key_idx(sorted[i]);
    sbox[i] = *reinterpret_cast<const float4*>(bb + (int64_t)idx * 4);
    ssz[i] = sz[idx];
  }
  __syncthreads();
  // mask[i][w] bit b: candidate j = 32 w + b (j > i) overlaps candidate i by more than the threshold: JaccardOverlap(j, i) of the
  // reference (mkldnn_detection_output_node.cpp:366-401) with box i held in registers and box j read as one 16-byte vector
  for (int cell = threadIdx.x; cell < m * words; cell += blockDim.x) {
    const int i = cell / words, w = cell % words;
    unsigned bits = 0u;
    const int j0 = max(w * 32, i + 1), j1 = min(w * 32 + 32, m);
    if (j0 < j1) {
      const float4 bi = sbox[i];
      const float si = ssz[i];
      for (int j = j0; j < j1; ++j) {
        const float4 bj = sbox[j];
        // jaccard(bb, sz, i1 = j, i2 = i): xmin2 = bi.x ... ; same expression order as the device function above
        if (bi.x > bj.z || bi.z < bj.x || bi.y > bj.w || bi.w < bj.y) continue;
        const float iw = fminf(bj.z, bi.z) - fmaxf(bj.x, bi.x);
        const float ih = fminf(bj.w, bi.w) - fmaxf(bj.y, bi.y);
        if (iw <= 0.f || ih <= 0.f) continue;
        const float inter = iw * ih;
        if (inter / (ssz[j] + si - inter) > p.nms_thr) bits |= 1u << (j & 31);
      }
    }
    mask[cell] = bits;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned removed = 0u;          // lane w holds word w of the "suppressed" set (words <= 16)
    int32_t* my = kept + (int64_t)blockIdx.x * max_keep;
    int det = 0;
    for (int i = 0; i < m; ++i) {
      const unsigned word = __shfl_sync(0xffffffffu, removed, i >> 5);
      if (!((word >> (i & 31)) & 1u)) {
        if (lane == 0) my[det] = (int)key_idx(sorted[i]);
        ++det;
        if (lane < words) removed |= mask[i * words + lane];
      }
    }
    if (lane == 0) cnt[blockIdx.x] = det;
  }
}

// one block per image: keep_top_k across classes by radix select; lists re-stored per class in score order
// dynamic shared memory: keys [cap] | sel [KEEP_KMAX] | sorted [KEEP_KMAX] | counts [num_classes] | bases [num_classes]
__global__ void __launch_bounds__(1024) detout_keep_fast_kernel(DetOutParams p, const float* __restrict__ conf_t, int cap, int max_keep, int32_t* kept,
                                                                int32_t* cnt, int32_t* totals) {
  extern __shared__ __align__(16) unsigned char keep_smem[];
  __shared__ SelectShared sh;
  uint64_t* keys = reinterpret_cast<uint64_t*>(keep_smem);
  uint64_t* sel = keys + cap;
  uint64_t* sorted = sel + KEEP_KMAX;
  int* counts = reinterpret_cast<int*>(sorted + KEEP_KMAX);
  int* bases = counts + p.num_classes;
  const int n = blockIdx.x, C = p.num_classes, P = p.num_priors;
  int32_t* my_cnt = cnt + (int64_t)n * C;
  int32_t* my_kept = kept + (int64_t)n * C * max_keep;
  const float* conf = conf_t + (int64_t)n * C * P;
  for (int c = threadIdx.x; c < C; c += blockDim.x) counts[c] = my_cnt[c];
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int c = 0; c < C; ++c) { bases[c] = t; t += counts[c]; }
    sh.count = t;
  }
  __syncthreads();
  const int total = sh.count;
  if (!(p.keep_top_k > -1 && total > p.keep_top_k)) {
    if (threadIdx.x == 0) totals[n] = total;
    return;
  }
  // one flat pass over the `total` kept boxes (class found by binary search over the prefix sums): every thread has its loads in flight
  // at once.  (A loop over classes serialised 81 dependent global-load round trips per image: 124 us of the 450 us DetectionOutput.)
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int lo_c = 0, hi_c = C - 1;
    while (lo_c < hi_c) {                 // last class whose base <= e (classes with zero boxes share a base: take the one that owns e)
      const int mid = (lo_c + hi_c + 1) >> 1;
      if (bases[mid] <= e) lo_c = mid; else hi_c = mid - 1;
    }
    const int c = lo_c, k = e - bases[c];
    const int idx = my_kept[(int64_t)c * max_keep + k];
    keys[e] = make_key(conf[(int64_t)c * P + idx], (uint32_t)(c * P + idx));
  }
  __syncthreads();
  auto key_at = [&](int i) -> uint64_t { return keys[i]; };
  const uint64_t T = radix_select_desc(key_at, total, p.keep_top_k, total, sh);
  if (threadIdx.x == 0) sh.count = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    if (keys[i] >= T) sel[atomicAdd(&sh.count, 1)] = keys[i];
  __syncthreads();
  rank_sort_desc(sel, sorted, p.keep_top_k);
  for (int c = threadIdx.x; c < C; c += blockDim.x) counts[c] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 0; j < p.keep_top_k; ++j) {
      const uint32_t lin = key_idx(sorted[j]);
      const int label = (int)(lin / (uint32_t)P), idx = (int)(lin % (uint32_t)P);
      my_kept[(int64_t)label * max_keep + counts[label]++] = idx;
    }
    totals[n] = p.keep_top_k;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) my_cnt[c] = counts[c];
}

// rows [image_id, label, conf, xmin, ymin, xmax, ymax] grouped by image, class, list order.  One block per image; the row offset of
// every (image, class) list comes from prefix sums held in shared memory.  conf_t: scores as [image][class][prior].
__global__ void detout_pack_fast_kernel(DetOutParams p, const float* __restrict__ conf_t, const float* __restrict__ boxes, int max_keep,
                                        const int32_t* __restrict__ kept, const int32_t* __restrict__ cnt, const int32_t* __restrict__ totals) {
  extern __shared__ int pack_smem[];      // [num_classes + 1] row offsets of this image's class lists
  __shared__ int s_base, s_all;
  const int n = blockIdx.x, C = p.num_classes, P = p.num_priors;
  for (int c = threadIdx.x; c < C; c += blockDim.x) pack_smem[c + 1] = cnt[n * C + c];
  if (threadIdx.x == 0) {
    int b = 0, all = 0;
    for (int i = 0; i < p.N; ++i) { const int t = totals[i]; if (i < n) b += t; all += t; }
    s_base = b; s_all = all;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    pack_smem[0] = 0;
    for (int c = 0; c < C; ++c) pack_smem[c + 1] += pack_smem[c];
  }
  __syncthreads();
  const int base = s_base, mine = pack_smem[C];
  const float* conf = conf_t + (int64_t)n * C * P;
  for (int c = 0; c < C; ++c) {
    const int lc = p.share_location ? 0 : c;
    const float* bb = boxes + ((int64_t)n * p.num_loc_classes + lc) * P * 4;
    const int first = pack_smem[c], len = pack_smem[c + 1] - first;
    for (int k = threadIdx.x; k < len; k += blockDim.x) {
      const int64_t row = (int64_t)base + first + k;
      if (row >= p.out_rows) break;
      const int idx = kept[((int64_t)n * C + c) * max_keep + k];
      float xmin = bb[idx * 4 + 0], ymin = bb[idx * 4 + 1], xmax = bb[idx * 4 + 2], ymax = bb[idx * 4 + 3];
      if (p.clip_after) {
        xmin = fmaxf(0.f, fminf(1.f, xmin)); ymin = fmaxf(0.f, fminf(1.f, ymin));
        xmax = fmaxf(0.f, fminf(1.f, xmax)); ymax = fmaxf(0.f, fminf(1.f, ymax));
      }
      float* o = p.out + row * 7;
      o[0] = (float)n; o[1] = (float)c; o[2] = conf[(int64_t)c * P + idx];
      o[3] = xmin; o[4] = ymin; o[5] = xmax; o[6] = ymax;
    }
  }
  // the last image also clears the unused tail and writes the end-of-list marker
  if (n == p.N - 1) {
    const int64_t used = min((int64_t)s_all, p.out_rows);
    for (int64_t i = used * 7 + threadIdx.x; i < p.out_rows * 7; i += blockDim.x) p.out[i] = (i == used * 7) ? -1.f : 0.f;
  }
  (void)mine;
}

bool detout_fast_path(const DetOutParams& p) {
  const char* off = getenv("B200_DETOUT_FAST");
  if (off && off[0] == '0') return false;
  const int max_keep = p.top_k > 0 ? std::min(p.top_k, p.num_priors) : p.num_priors;
  return p.top_k > 0 && p.top_k <= SEL_KMAX && (p.keep_top_k < 0 || p.keep_top_k <= KEEP_KMAX) && (int64_t)p.num_classes * max_keep <= SORT_MAX &&
         p.num_priors <= 32768;
}
int detout_kcap(const DetOutParams& p) { return (int)round_up((int64_t)std::min(p.top_k, p.num_priors), 32); }
size_t detout_nms_fast_smem(const DetOutParams& p) {
  const int64_t k = detout_kcap(p);
  return (size_t)(round_up((int64_t)p.num_priors * 4, 16) + k * 8 * 2 + k * 16 + k * 4 + k * (k / 32) * 4);
}
size_t detout_keep_fast_smem(const DetOutParams& p, int max_keep) {
  return (size_t)((int64_t)p.num_classes * max_keep * 8 + KEEP_KMAX * 8 * 2 + p.num_classes * 8 + 16);
}

int detout_max_keep(const DetOutParams& p) { return p.top_k > 0 ? std::min(p.top_k, p.num_priors) : p.num_priors; }

}  // namespace

void detect_init_device() {
  cudaFuncSetAttribute(nms_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM);
  cudaFuncSetAttribute(detout_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM);
  cudaFuncSetAttribute(detout_keep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM);
  cudaFuncSetAttribute(detout_nms_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(detout_keep_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

int64_t nms_workspace_bytes(int B, int C, int Nb, int max_out, bool soft) {
  int64_t sel = round_up((int64_t)B * C * max_out * 4, 256);
  int64_t cnt = round_up((int64_t)B * C * 4, 256);
  int64_t keys = (int64_t)next_pow2((int)std::min<int64_t>((int64_t)B * C * max_out, 1 << 30)) * 8;
  // soft-NMS: current score + suppress_begin_index + decayed score per (batch, class, box), decayed score per selection
  int64_t soft_bytes = soft ? 3 * round_up((int64_t)B * C * Nb * 4, 256) + round_up((int64_t)B * C * max_out * 4, 256) : 0;
  return sel + cnt + keys + soft_bytes + 1024;
}

void launch_nms(const NmsParams& p, cudaStream_t s) {
  nms_check(p);
  if (p.B * p.C == 0 || p.max_out == 0) return;
  char* ws = (char*)p.workspace;
  int32_t* sel = (int32_t*)ws;
  int64_t off = round_up((int64_t)p.B * p.C * p.max_out * 4, 256);
  int32_t* cnt = (int32_t*)(ws + off);
  off += round_up((int64_t)p.B * p.C * 4, 256);
  uint64_t* keys = (uint64_t*)(ws + off);
  const int64_t out_rows = (int64_t)p.B * p.C * p.max_out;
  const int tot_pad = next_pow2((int)std::max<int64_t>(out_rows, 2));
  if (p.soft_sigma > 0.f) {
    off += (int64_t)tot_pad * 8;
    const int64_t per = round_up((int64_t)p.B * p.C * p.Nb * 4, 256);
    float* cur = (float*)(ws + off); off += per;
    int32_t* sbi = (int32_t*)(ws + off); off += per;
    float* out_sc = (float*)(ws + off); off += per;
    float* sel_sc = (float*)(ws + off);
    const float scale = (float)(-0.5 / p.soft_sigma);      // mkldnn_non_max_suppression_node.cpp:158-160
    nms_soft_select_kernel<<<p.B * p.C, 256, 0, s>>>(p, scale, cur, sbi, sel, sel_sc, cnt);
    nms_finalize_kernel<<<1, 1024, 0, s>>>(p, sel, cnt, keys, tot_pad, out_rows, sel_sc, out_sc);
    return;
  }
  const int n_pad = next_pow2(std::max(p.Nb, 2));
  nms_select_kernel<<<p.B * p.C, 512, (size_t)n_pad * 8, s>>>(p, n_pad, sel, cnt);
  nms_finalize_kernel<<<1, 1024, 0, s>>>(p, sel, cnt, keys, tot_pad, out_rows, nullptr, nullptr);
}

int64_t detout_workspace_bytes(const DetOutParams& p) {
  int64_t boxes = round_up((int64_t)p.N * p.num_loc_classes * p.num_priors * 16, 256);
  int64_t sizes = round_up((int64_t)p.N * p.num_loc_classes * p.num_priors * 4, 256);
  int64_t kept = round_up((int64_t)p.N * p.num_classes * detout_max_keep(p) * 4, 256);
  int64_t cnt = round_up((int64_t)p.N * p.num_classes * 4, 256);
  int64_t totals = round_up((int64_t)p.N * 4, 256);
  int64_t conf_t = detout_fast_path(p) ? round_up((int64_t)p.N * p.num_priors * p.num_classes * 4, 256) : 0;
  return boxes + sizes + kept + cnt + totals + conf_t + 1024;
}

// Geometry the kernels cannot run: reported at COMPILE time (b200_compile -> LoadNetwork / QueryNetwork), not at the first inference
void detection_output_check(const DetOutParams& p) {
  B200_CHECK(!p.decrease_label_id, B200_ERR_UNSUPPORTED, "DetectionOutput: decrease_label_id (MXNet style) is not supported");
  if (detout_fast_path(p)) {
    B200_CHECK(detout_nms_fast_smem(p) <= 200 * 1024, B200_ERR_UNSUPPORTED, "DetectionOutput: " << p.num_priors << " priors x top_k " << p.top_k << " exceed shared memory");
    return;
  }
  B200_CHECK(p.num_priors <= SORT_MAX, B200_ERR_UNSUPPORTED, "DetectionOutput: more than " << SORT_MAX << " priors with top_k = " << p.top_k);
  B200_CHECK(p.keep_top_k < 0 || (int64_t)p.num_classes * detout_max_keep(p) <= SORT_MAX, B200_ERR_UNSUPPORTED,
             "DetectionOutput: num_classes*top_k > " << SORT_MAX);
}
void nms_check(const NmsParams& p) {
  B200_CHECK(p.soft_sigma >= 0.f, B200_ERR_INVALID, "NonMaxSuppression: negative soft_nms_sigma");
  B200_CHECK(p.Nb <= SORT_MAX, B200_ERR_UNSUPPORTED, "NonMaxSuppression: more than " << SORT_MAX << " boxes per batch");
  B200_CHECK((int64_t)p.B * p.C * p.Nb < (1ll << 32), B200_ERR_UNSUPPORTED, "NonMaxSuppression: problem too large");
}

void launch_detection_output(const DetOutParams& p, cudaStream_t s) {
  detection_output_check(p);
  const int max_keep = detout_max_keep(p);
  const int64_t all = (int64_t)p.num_classes * max_keep;
  char* ws = (char*)p.workspace;
  int64_t off = 0;
  float* boxes = (float*)(ws + off); off += round_up((int64_t)p.N * p.num_loc_classes * p.num_priors * 16, 256);
  float* sizes = (float*)(ws + off); off += round_up((int64_t)p.N * p.num_loc_classes * p.num_priors * 4, 256);
  int32_t* kept = (int32_t*)(ws + off); off += round_up((int64_t)p.N * p.num_classes * max_keep * 4, 256);
  int32_t* cnt = (int32_t*)(ws + off); off += round_up((int64_t)p.N * p.num_classes * 4, 256);
  int32_t* totals = (int32_t*)(ws + off); off += round_up((int64_t)p.N * 4, 256);
  int64_t nd = (int64_t)p.N * p.num_loc_classes * p.num_priors;
  detout_decode_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, s>>>(p, boxes, sizes);
  if (detout_fast_path(p)) {
    float* conf_t = (float*)(ws + off);
    dim3 tg((unsigned)((p.num_classes + 31) / 32), (unsigned)((p.num_priors + 31) / 32), (unsigned)p.N);
    detout_transpose_conf_kernel<<<tg, dim3(32, 8), 0, s>>>(p.conf, conf_t, p.num_priors, p.num_classes);
    detout_nms_fast_kernel<<<p.N * p.num_classes, SEL_THREADS, detout_nms_fast_smem(p), s>>>(p, conf_t, boxes, sizes, detout_kcap(p), max_keep, kept, cnt);
    detout_keep_fast_kernel<<<p.N, 1024, detout_keep_fast_smem(p, max_keep), s>>>(p, conf_t, (int)all, max_keep, kept, cnt, totals);
    detout_pack_fast_kernel<<<p.N, 256, (size_t)(p.num_classes + 1) * 4, s>>>(p, conf_t, boxes, max_keep, kept, cnt, totals);
    return;
  }
  const int n_pad = next_pow2(std::max(p.num_priors, 2));
  detout_nms_kernel<<<p.N * p.num_classes, 512, (size_t)n_pad * 8, s>>>(p, boxes, sizes, n_pad, max_keep, kept, cnt);
  const int keep_pad = next_pow2((int)std::max<int64_t>(std::min<int64_t>(all, SORT_MAX), 2));
  detout_keep_kernel<<<p.N, 1024, (size_t)keep_pad * 8, s>>>(p, keep_pad, max_keep, kept, cnt, totals);
  detout_pack_kernel<<<1, 1024, 0, s>>>(p, boxes, max_keep, kept, cnt, totals);
}

}  // namespace b200
