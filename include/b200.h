/*
 * b200.h - C ABI of libb200kernels.so, the sm_100a kernel library + graph engine that sits
 * BELOW the B200 Inference-Engine device plugin (plugin/src) and below the Python host mirror
 * (dldt_b200/).  Plain C: opaque handles, POD descriptors, int status codes, explicit streams
 * owned by the request objects.  No C++ types and no exceptions cross this boundary; this is the
 * only place CUDA appears (SURVEY.md 8b "Below the plugin").
 *
 * Which reference interface each group of entry points replaces (paths under /root/reference):
 *
 *   b200_graph_*            the node factory of MKLDNNGraph::Replicate
 *                           (inference-engine/src/mkldnn_plugin/mkldnn_graph.cpp:181-320): one call per
 *                           ngraph::Node of the (already nGraph-optimised) ngraph::Function.
 *   b200_compile            MKLDNNGraph::InitGraph (mkldnn_graph.cpp:322-351): fusion passes
 *                           (mkldnn_graph_optimizer.cpp:53-148), layout/impl selection, reorder insertion,
 *                           MemorySolver (mkldnn_memory_solver.cpp:51-96), weight repack
 *                           (ExecuteConstantNodesOnly).
 *   b200_network_*          MKLDNNExecNetwork metrics + GetExecGraphInfo
 *                           (mkldnn_exec_network.cpp:185-215, mkldnn_graph_dumper.cpp).
 *   b200_request_create     MKLDNNExecNetwork::CreateInferRequestImpl (mkldnn_exec_network.cpp:35-147).
 *   b200_request_start/wait MKLDNNInferRequest::InferImpl = PushInputData + MKLDNNGraph::Infer +
 *                           PullOutputData (mkldnn_infer_request.cpp:168-197, mkldnn_graph.cpp:708-834),
 *                           split into the non-blocking enqueue and the blocking wait that the 3-stage
 *                           pipeline of docs/template_plugin/src/template_async_infer_request.cpp:23-39 needs.
 *   b200_request_perf_*     MKLDNNInferRequest::GetPerformanceCounts (mkldnn_infer_request.cpp:199-205).
 *   b200_host_*             the default blob allocator (inference-engine/include/ie_allocator.hpp:26-40),
 *                           pinned so H2D/D2H run at PCIe speed and asynchronously.
 *
 * All functions return B200_OK (0) or a negative b200_status; b200_last_error() returns a thread-local
 * human readable message for the last failure on the calling thread.
 */
#ifndef B200_H_
#define B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#define B200_ABI_VERSION 1
#define B200_MAX_RANK 8
#define B200_MAX_INPUTS 16
#define B200_MAX_IATTR 40
#define B200_MAX_FATTR 8

typedef enum b200_status {
  B200_OK = 0,
  B200_ERR_INVALID = -1,      /* bad argument / malformed graph  -> IE ParameterMismatch / GeneralError */
  B200_ERR_UNSUPPORTED = -2,  /* op or attribute combination has no kernel -> IE NotImplemented        */
  B200_ERR_CUDA = -3,         /* CUDA runtime/driver error (sticky)       -> IE GeneralError           */
  B200_ERR_NO_DEVICE = -4,    /* no sm_100 device / bad DEVICE_ID         -> IE NotFound               */
  B200_ERR_BUSY = -5,         /* request already running                  -> IE RequestBusy            */
  B200_ERR_OOM = -6,
  B200_NOT_READY = 1,         /* wait() timed out                          -> IE RESULT_NOT_READY       */
  B200_ALREADY_PINNED = 2     /* host_register(): the memory is page-locked already, nothing to undo        */
} b200_status;

typedef enum b200_dtype {
  B200_F32 = 0, B200_F16 = 1, B200_BF16 = 2, B200_I32 = 3, B200_U8 = 4, B200_I64 = 5, B200_I8 = 6,
  B200_BOOL = 7
} b200_dtype;

/* Inference precision of the compiled network (analogue of ENFORCE_BF16, ie_plugin_config.hpp:373). */
typedef enum b200_precision { B200_PREC_FP32 = 0, B200_PREC_BF16 = 1, B200_PREC_FP16 = 2 } b200_precision;

/* Host-side blob layout for rank-4 I/O (InferenceEngine::Layout, ie_common.h:70-101). */
typedef enum b200_layout { B200_LAYOUT_ANY = 0, B200_LAYOUT_NCHW = 1, B200_LAYOUT_NHWC = 2 } b200_layout;

/* Operation set = the nGraph opset1-8 operations on the hot path.
 * Integer attributes live in b200_node_desc.i[], float attributes in .f[]; the layout per op is
 * documented next to each enumerator. */
typedef enum b200_op {
  B200_OP_PARAMETER = 0,     /* shape/dtype from desc                                              */
  B200_OP_CONSTANT = 1,      /* shape/dtype/data from desc (data copied)                           */
  B200_OP_RESULT = 2,        /* in: x                                                              */
  B200_OP_CONVOLUTION = 3,   /* in: x[N,C,H,W], w[O,I,kh,kw]; i={sh,sw,dh,dw,pt,pl,pb,pr}          */
  B200_OP_GROUP_CONVOLUTION = 4, /* in: x, w[G,O/G,I/G,kh,kw]; same attrs                          */
  B200_OP_BATCH_NORM = 5,    /* in: x,gamma,beta,mean,var; f={eps}  (opset5 order)                 */
  B200_OP_ADD = 6, B200_OP_SUBTRACT = 7, B200_OP_MULTIPLY = 8, B200_OP_DIVIDE = 9,
  B200_OP_MAXIMUM = 10, B200_OP_MINIMUM = 11,       /* numpy broadcast                             */
  B200_OP_RELU = 12, B200_OP_SIGMOID = 13, B200_OP_TANH = 14,
  B200_OP_CLAMP = 15,        /* f={lo,hi}                                                          */
  B200_OP_SWISH = 16,        /* f={beta}                                                           */
  B200_OP_GELU = 17,         /* i={0 erf | 1 tanh}                                                 */
  B200_OP_HSWISH = 18, B200_OP_HSIGMOID = 19, B200_OP_MISH = 20, B200_OP_EXP = 21,
  B200_OP_SQRT = 22, B200_OP_ABS = 23, B200_OP_NEGATIVE = 24, B200_OP_ERF = 25,
  B200_OP_ELU = 26,          /* f={alpha}                                                          */
  B200_OP_MAX_POOL = 27,     /* i={kh,kw,sh,sw,pt,pl,pb,pr,ceil_mode}                              */
  B200_OP_AVG_POOL = 28,     /* i={kh,kw,sh,sw,pt,pl,pb,pr,ceil_mode,exclude_pad}                  */
  B200_OP_REDUCE_MEAN = 29,  /* i={keep_dims,n_axes,axes...}                                       */
  B200_OP_MATMUL = 30,       /* i={transpose_a,transpose_b}                                        */
  B200_OP_SOFTMAX = 31,      /* i={axis}                                                           */
  B200_OP_MVN = 32,          /* i={normalize_variance,eps_mode(0 inside,1 outside),n_axes,axes...}; f={eps} */
  B200_OP_RESHAPE = 33,      /* target shape in desc.shape (fully resolved)                        */
  B200_OP_TRANSPOSE = 34,    /* i={perm...}                                                        */
  B200_OP_CONCAT = 35,       /* i={axis}                                                           */
  B200_OP_SPLIT = 36,        /* i={axis,num_splits} -> num_splits outputs                          */
  B200_OP_GATHER = 37,       /* in: data, indices(i32); i={axis}                                   */
  B200_OP_TOPK = 38,         /* i={k,axis,mode(0 max,1 min),sort(0 none,1 value,2 index)} -> values, indices(i32) */
  B200_OP_NMS = 39,          /* in: boxes[B,N,4], scores[B,C,N]; i={max_out_per_class,center_format,sort_desc};
                                f={iou_thr,score_thr,soft_sigma} -> indices[M,3] i32, scores[M,3] f32, valid[1] i32 */
  B200_OP_DETECTION_OUTPUT = 40, /* in: loc, conf, priors; i={num_classes,background,top_k,keep_top_k,code_type(0 corner,
                                1 center_size),share_location,var_in_target,normalized,clip_before,clip_after,
                                input_h,input_w,decrease_label_id}; f={nms_thr,conf_thr}              */
  B200_OP_CONVERT = 41,      /* target dtype in desc.dtype                                         */
  B200_OP_POWER = 42,        /* numpy broadcast x^y                                                */
  B200_OP_PRELU = 43,        /* in: x, slope (per-channel or scalar)                               */
  B200_OP_REDUCE_SUM = 44, B200_OP_REDUCE_MAX = 45, /* same attrs as REDUCE_MEAN                   */
  B200_OP_SQUARED_DIFFERENCE = 46,
  B200_OP_ROUND = 47,        /* i={mode: 0 half to even | 1 half away from zero}  (opset5 Round; what FakeQuantize decomposes to) */
  B200_OP_FAKE_QUANTIZE = 48, /* in: x, in_low, in_high, out_low, out_high (constants: one value, or one per channel = axis 1); i={levels} */
  B200_OP_STRIDED_SLICE = 49, /* in: x.  The input is VIEWED as shape v[r] (same element count), element (o0..or-1) of the [n0..nr-1] result is
                               * v[begin0 + o0 * stride0, ...] (stride may be negative), and the result is viewed as desc.shape.
                               * i={r, v[r], begin[r], stride[r], n[r]}: what ngraph::make_slice_plan reduces opset1 StridedSlice to */
  B200_OP__COUNT
} b200_op;

typedef struct b200_node_desc {
  int32_t op;                       /* b200_op */
  const char* name;                 /* friendly name (kept for exec-graph info / perf counters) */
  int32_t n_inputs;
  int32_t inputs[B200_MAX_INPUTS];  /* value ids returned by earlier b200_graph_add calls */
  int32_t i[B200_MAX_IATTR];
  float f[B200_MAX_FATTR];
  int32_t dtype;                    /* PARAMETER / CONSTANT / CONVERT */
  int32_t rank;                     /* PARAMETER / CONSTANT / RESHAPE */
  int64_t shape[B200_MAX_RANK];
  const void* data;                 /* CONSTANT payload in `dtype`, row-major */
  uint64_t data_bytes;
} b200_node_desc;

typedef struct b200_tensor_info {
  char name[256];
  int32_t dtype;                    /* dtype of the graph edge */
  int32_t rank;
  int64_t shape[B200_MAX_RANK];
} b200_tensor_info;

typedef struct b200_compile_options {
  int32_t abi_version;              /* = B200_ABI_VERSION */
  int32_t device;                   /* CUDA ordinal (DEVICE_ID) */
  int32_t precision;                /* b200_precision */
  int32_t conv_impl;                /* 0 auto | 1 force CUDA-core fp32-accumulate direct path (debug) */
  int32_t reserved[12];
} b200_compile_options;

typedef struct b200_exec_node_info {
  char name[256];                   /* friendly name of the node that produced the kernel's output */
  char type[64];                    /* kernel class: Convolution, FullyConnected, Pooling, Eltwise ... */
  char impl[64];                    /* e.g. tcgen05_bf16_128x128x64, simt_fp32, dw_nhwc ... */
  char fused[1024];                 /* comma separated original layer names folded into this kernel */
  char out_precision[16];
  char out_layout[16];
  int32_t exec_order;
  double flops;                     /* algorithmic FLOPs (2*MAC) of this kernel */
  double bytes;                     /* algorithmic bytes (activations in+out + weights once) */
} b200_exec_node_info;

typedef struct b200_graph b200_graph;
typedef struct b200_network b200_network;
typedef struct b200_request b200_request;

/* ---- library / device ------------------------------------------------------------------- */
B200_API int b200_abi_version(void);
B200_API const char* b200_last_error(void);
B200_API int b200_device_count(int* count);
B200_API int b200_device_name(int device, char* buf, size_t buf_len);
B200_API int b200_device_props(int device, int* sm_count, int* cc_major, int* cc_minor, uint64_t* mem_bytes);

/* ---- pinned host memory ----------------------------------------------------------------- */
B200_API int b200_host_alloc(void** ptr, size_t bytes);
B200_API int b200_host_free(void* ptr);
/* Page-locks caller-owned memory (B200_OK: undo with b200_host_unregister; B200_ALREADY_PINNED: nothing was done). */
B200_API int b200_host_register(void* ptr, size_t bytes);
B200_API int b200_host_unregister(void* ptr);

/* ---- device memory for remote blobs (inference-engine/include/ie_remote_context.hpp:58-178): the pointers b200_request_start takes
 * may be DEVICE addresses (any GPU of the box), in which case the "copy" is device to device and nothing crosses PCIe.
 * direction: 0 = host -> device, 1 = device -> host; synchronous. */
B200_API int b200_device_alloc(int device, void** dptr, size_t bytes);
B200_API int b200_device_free(int device, void* dptr);
B200_API int b200_device_copy(int device, void* dst, const void* src, size_t bytes, int direction);

/* ---- graph construction (host only, no CUDA) -------------------------------------------- */
B200_API int b200_graph_create(const char* name, b200_graph** out);
B200_API void b200_graph_destroy(b200_graph* g);
/* Adds a node, runs shape/type inference, writes the ids of its output values (n_outputs of them;
 * out_ids must have room for B200_MAX_INPUTS entries). */
B200_API int b200_graph_add(b200_graph* g, const b200_node_desc* desc, int32_t* out_ids, int32_t* n_outputs);
B200_API int b200_graph_value_info(const b200_graph* g, int32_t value_id, b200_tensor_info* info);
/* 1 if the op (with these attributes / input ranks) has a kernel: used by QueryNetwork. */
B200_API int b200_op_supported(int32_t op);

/* ---- compile ---------------------------------------------------------------------------- */
B200_API int b200_compile(const b200_graph* g, const b200_compile_options* opts, b200_network** out);
B200_API void b200_network_destroy(b200_network* net);
B200_API int b200_network_num_inputs(const b200_network* net);
B200_API int b200_network_num_outputs(const b200_network* net);
B200_API int b200_network_io_info(const b200_network* net, int is_output, int index, b200_tensor_info* info);
B200_API int b200_network_num_exec_nodes(const b200_network* net);
B200_API int b200_network_exec_node(const b200_network* net, int index, b200_exec_node_info* info);
/* key in {"arena_bytes","weight_bytes","flops","bytes","num_kernels","device","precision"} */
B200_API int b200_network_stat(const b200_network* net, const char* key, double* value);

/* Kernel selection (the per-layer choice between alternative kernels is MEASURED once per network, engine.cu autotune):
 * _tune runs it now on a temporary arena instead of at the first request; _get_tuning returns the decisions as text
 * ("<exec node>\t<impl>\n" lines; *needed = bytes including the terminator, buf may be NULL to query); _set_tuning installs
 * decisions of an earlier load so they are not measured again.  This table is what ExecutableNetwork::Export carries besides the IR
 * (docs/template_plugin/src/template_executable_network.cpp:179-200; inference-engine/src/inference_engine/ie_core.cpp:265-340). */
B200_API int b200_network_tune(b200_network* net);
B200_API int b200_network_get_tuning(const b200_network* net, char* buf, size_t capacity, size_t* needed);
B200_API int b200_network_set_tuning(b200_network* net, const char* text);

/* The static activation-memory planner on its own: boxes (first, last, bytes) over execution indices,
 * last < 0 = lives to the end; writes the offset of every box and the arena size.  Replaces
 * MemorySolver::solve/getOffset (inference-engine/src/mkldnn_plugin/mkldnn_memory_solver.cpp:51-96,126-131). */
B200_API int b200_plan_arena(int32_t n_boxes, const int64_t* bytes, const int32_t* first, const int32_t* last, int64_t* offsets,
                             int64_t* total_bytes);

/* ---- infer requests --------------------------------------------------------------------- */
B200_API int b200_request_create(b200_network* net, b200_request** out);
B200_API void b200_request_destroy(b200_request* req);
/* Host-side blob format of input/output `index` (default: inputs = graph dtype with i64->i32,
 * f16/bf16->f32; outputs f32 / i32; layout NCHW).  Normally called before the first start(); it may be
 * called again later (a blob with another layout arrived): the request is drained, the captured graph
 * is dropped and the next start() records it again. */
B200_API int b200_request_set_io_format(b200_request* req, int is_output, int index, int32_t dtype, int32_t layout);
B200_API int b200_request_io_bytes(const b200_request* req, int is_output, int index, uint64_t* bytes);
/* Enqueue H2D of every input from host_inputs[i], the captured CUDA graph, and D2H of every output
 * into host_outputs[i] on the request's stream.  Returns immediately. Host buffers should be pinned
 * (b200_host_alloc / b200_host_register) for the copies to be asynchronous; device addresses (b200_device_alloc,
 * or any CUDA allocation of this process) are accepted too (unified addressing picks the copy direction). */
B200_API int b200_request_start(b200_request* req, const void* const* host_inputs, void* const* host_outputs);
/* Block until the request's stream is drained.  timeout_ms < 0: forever; 0: poll. */
B200_API int b200_request_wait(b200_request* req, int64_t timeout_ms);
/* Device-resident path: device address of the staging buffer that holds input/output `index` in the
 * host blob format, and a launch that runs ingest -> kernels -> egress only (no PCIe). */
B200_API int b200_request_device_io(b200_request* req, int is_output, int index, void** dptr, uint64_t* bytes);
B200_API int b200_request_launch(b200_request* req);
/* Time `iters` back-to-back launches with CUDA events on the request's stream.
 * mode 0: device-resident launch; mode 1: start() with the given host buffers (H2D+graph+D2H). */
B200_API int b200_request_time(b200_request* req, int iters, int mode, const void* const* host_inputs,
                               void* const* host_outputs, float* total_ms);
/* Per exec-node device time (microseconds) of one eager, event-bracketed run (PERF_COUNT=YES). */
B200_API int b200_request_perf_counts(b200_request* req, float* usec, int capacity);
B200_API int b200_request_num_launches(const b200_request* req, int* kernels_per_infer);
/* Raw stream handle (cudaStream_t) for callers that want to enqueue their own work. */
B200_API int b200_request_stream(b200_request* req, void** stream);
/* Debug: copy the device contents of exec node `index`'s output, converted to f32 in the logical
 * (row-major, NCHW) layout, into host memory (count elements). */
B200_API int b200_request_dump_node(b200_request* req, int index, float* out, uint64_t count);

#ifdef __cplusplus
}
#endif
#endif /* B200_H_ */
