/* recom_b200.h — C ABI of the B200-native fused embedding-column path.
 *
 * This header is the drop-in boundary. Each entry point replaces one piece of
 * the plugin ABI the reference's TF ops bind with dlopen/dlsym
 * (paths relative to the reference root):
 *
 *   recom_plan_create / recom_plan_create_from_blob / recom_plan_destroy
 *       replace dlopen(dlpath) + dlsym("CreateConstBuffers") and the cudaFree in
 *       the op destructor
 *       (tensorflow_addons/custom_ops/feature_column_process/feature_column_process_op_gpu.cu.cc:51-62,
 *        feature_column_process_ops.h:47,
 *        tensorflow_addons/graph_optimizers/cuda_emitter.cc:2260-2301).
 *       The plan is what the reference's generated .cu file is: the per-column
 *       description of transform -> lookup -> combine -> concat slot.
 *   recom_plan_workspace_bytes
 *       replaces the arena sizing `buffer_size_sum = sum alignmem(expr*sizeof)`
 *       (cuda_emitter.cc:2151-2179) and the malloc_temp/malloc_buff callbacks
 *       (feature_column_process_op_gpu.cu.cc:99-111): the caller owns memory.
 *   recom_plan_run
 *       replaces dlsym("ProcessFeatureColumns")
 *       (feature_column_process_ops.h:37-43, cuda_emitter.cc:2306-2494) AND
 *       dlsym("ConcatOutputs")
 *       (tensorflow_addons/custom_ops/concat_outputs/concat_outputs_ops.h:35,
 *        concat_outputs_op_gpu.cu.cc:133-140): it writes the final row-major
 *       [batch, out_width] tensor directly, stream ordered, no host sync.
 *   recom_scatter_columns
 *   recom_dense_relu            (8(f) N4: the first dense layer behind the concat, tcgen05)
 *       keeps ConcatOutputs' `host_inputs` capability
 *       (concat_outputs_op_gpu.cu.cc:186-232): columns computed elsewhere are
 *       scattered into their concat slices.
 *   recom_segment_ids_to_row_splits
 *       replaces experiment::ComputeSegmentOffsets (cuda_emitter.cc:768-818).
 *   recom_plan_run_sharded, recom_shard_wait, recom_ipc_alloc/open/close/free
 *       no counterpart in the reference (one GPU + CPU spill, tf_cpu_gpu_pass):
 *       column groups on several GPUs of one NVSwitch box, every row's slice
 *       stored straight into the concat tensor of the GPU that owns the row,
 *       exchange and barrier inside the same kernel.
 *
 * No C++ types, no torch types: plain pointers and sizes only. Every function
 * returns a recom_status; recom_last_error() gives a thread-local message. The
 * library never calls exit() (the reference does, via CubDebugExit).
 */
#ifndef RECOM_B200_H_
#define RECOM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 4: + RECOM_SEG_COO_RESHAPE_I64 / RECOM_SEG_FIXED, recom_dense_relu, recom_plan_set_request_overlap
 *    (additive: every struct layout of version 3 is unchanged) */
#define RECOM_B200_ABI_VERSION 4

#if defined(__GNUC__)
#define RECOM_API __attribute__((visibility("default")))
#else
#define RECOM_API
#endif

typedef enum recom_status {
  RECOM_OK = 0,
  RECOM_ERR_INVALID_ARGUMENT = 1,
  RECOM_ERR_CUDA = 2,
  RECOM_ERR_WORKSPACE_TOO_SMALL = 3,
  RECOM_ERR_UNSUPPORTED = 4,
  RECOM_ERR_OUT_OF_MEMORY = 5,
  RECOM_ERR_ID_OUT_OF_RANGE = 6
} recom_status;

typedef enum recom_dtype {
  RECOM_F32 = 0,
  RECOM_BF16 = 1,
  RECOM_I32 = 2,
  RECOM_I64 = 3,
  RECOM_STR = 4, /* wire format: int32 offsets[n+1] slot + uint8 bytes slot */
  RECOM_F64 = 5
} recom_dtype;

/* The three lookup shapes LookupOptimizer produces
 * (tensorflow_addons/graph_optimizers/lookup_optimizer.cc:137-155,270-440). */
typedef enum recom_lookup {
  RECOM_LOOKUP_DENSE = 0,          /* exactly one id per sample: GatherV2 */
  RECOM_LOOKUP_GATHER_SCATTER = 1, /* <=1 id per sample: ScatterNd(Gather) */
  RECOM_LOOKUP_SEGMENT = 2         /* multi-hot: SparseSegment*WithNumSegments */
} recom_lookup;

typedef enum recom_combiner {
  RECOM_COMBINE_SUM = 0,
  RECOM_COMBINE_MEAN = 1,
  RECOM_COMBINE_SQRTN = 2
} recom_combiner;

/* What is fused ahead of the gather. */
typedef enum recom_transform {
  RECOM_XFORM_NONE = 0,      /* values are row ids already (i32 / i64) */
  RECOM_XFORM_BUCKETIZE = 1, /* TF Bucketize: upper_bound over float boundaries */
  RECOM_XFORM_HASH_STR = 2,  /* TF StringToHashBucketFast: Fingerprint64 % n */
  RECOM_XFORM_HASH_INT = 3   /* AsString(int) then StringToHashBucketFast */
} recom_transform;

/* How the sample (segment) of each id is given for non-DENSE lookups. */
typedef enum recom_seg_format {
  RECOM_SEG_NONE = 0,
  RECOM_SEG_ROW_SPLITS_I32 = 1, /* int32[batch+1], CSR */
  RECOM_SEG_COO_I64 = 2, /* int64[nnz, seg_stride] SparseTensor.indices, row id in
                           column 0, row-major sorted: what the reference reads in
                           place (cuda_emitter.cc:1836-1873) */
  RECOM_SEG_SPLIT_STR = 3, /* the values are ONE string per row (STR: int32
                           offsets[batch+1] + payload) and the row's ids are the tokens of
                           tf.strings.split(row, delimiter) (StringSplit, string_split_op.cc:37-99,
                           162-170; the reference's model E joins its ids with ';',
                           examples/python/dlrm.py:49-53), hashed with StringToHashBucketFast.
                           transform must be HASH_STR, seg_slot -1, and seg_stride holds the
                           delimiter: RECOM_SPLIT_SPEC(d0, d1, d2, n, skip_empty) - any of the n <= 3
                           characters splits (n == 0: every character is a token), empty tokens
                           are skipped when skip_empty (TF's default). */
  RECOM_SEG_COO_RESHAPE_I64 = 4, /* SparseTensor.indices of rank R behind a SparseReshape, inlined
                           into the load of the row id the way the reference's EmitInputInline does
                           (cuda_emitter.cc:1874-1922: linearise with the input shape, divide by
                           the trailing output dimensions): seg_slot = int64[nnz, R] indices,
                           seg_slot + 1 = int64[R] dense_shape, and - with has_new_shape -
                           seg_slot + 2 = int64[Ro] new shape (one entry may be -1, inferred like
                           SparseReshape does); without it the new shape is
                           [prod(shape[:-1]), shape[-1]], the reshape of
                           safe_embedding_lookup_sparse (embedding_ops.py:882-941). The id's row is
                           column 0 of the reshaped indices: linear(e) / prod(new_shape[1:]).
                           seg_stride = RECOM_RESHAPE_SPEC(R, has_new_shape), 1 <= R, Ro <= 8. */
  RECOM_SEG_FIXED = 5      /* every row owns the same number T of consecutive ids: the values are
                           a dense [batch, T] tensor and the column is Sum(axis=1) over the gathered
                           [batch, T, D] rows - the reference's EmitBatchColReduction
                           (cuda_emitter.cc:1180-1244) fused with the gather in front of it.
                           seg_slot -1; seg_stride = T, or 0: T = value elements / batch. */
} recom_seg_format;
#define RECOM_RESHAPE_SPEC(rank, has_new_shape) ((int32_t)((rank) | ((has_new_shape) ? 256 : 0)))
#define RECOM_SPLIT_SPEC(d0, d1, d2, n, skip_empty)                                        \
  ((int32_t)((uint32_t)(uint8_t)(d0) | (uint32_t)(uint8_t)(d1) << 8 | (uint32_t)(uint8_t)(d2) << 16 | \
             (uint32_t)(n) << 24 | (uint32_t)((skip_empty) ? 1 : 0) << 28))

/* One embedding column ("FC" in the reference, graph_info.h:41-42). */
typedef struct recom_column_desc {
  int32_t lookup;      /* recom_lookup */
  int32_t combiner;    /* recom_combiner (ignored for DENSE / GATHER_SCATTER) */
  int32_t transform;   /* recom_transform */
  int32_t value_dtype; /* recom_dtype of the values slot */
  int32_t table;       /* index into recom_request.tables */
  int32_t table_dtype; /* RECOM_F32 or RECOM_BF16 */
  int64_t table_rows;
  int32_t dim;         /* embedding width D */
  int32_t out_offset;  /* element offset of this column's slice in the concat row */
  int32_t value_slot;  /* blob slot of the values (for STR: of the int32 offsets) */
  int32_t bytes_slot;  /* STR only: blob slot of the uint8 payload, else -1 */
  int32_t seg_slot;    /* blob slot of row_splits / COO indices, -1 for DENSE */
  int32_t seg_format;  /* recom_seg_format */
  int32_t seg_stride;  /* COO only: int64s per entry (sparse rank) */
  int32_t num_boundaries;    /* BUCKETIZE only */
  int64_t boundaries_offset; /* BUCKETIZE only: first boundary in `boundaries` */
  int64_t num_buckets;       /* HASH_* only */
} recom_column_desc;

/* Pre-lookup id filtering fused into the stage. The reference's
 * pre_lookup_optimizer.cc:66-123,413-478 folds the graph's comparison / tf.where chains on the ids
 * into interval sets and emits three TF-CPU ops; here the same decisions are taken per id inside
 * the kernel, in front of the transform:
 *   RECOM_FILTER_KEEP     Addons>GatherIndiceValue (gather_indice_value_ops.cc:33-70): an id whose
 *                         raw value lies in none of the closed intervals [lo_i, hi_i] is dropped -
 *                         not gathered, not counted by mean / sqrtn. (The reference's comparator
 *                         `x >= l || x <= r` is always true; the documented intent - and the graph
 *                         it replaces, e.g. sparse.retain(ids, ids >= 0) - is `l <= x <= r`.)
 *   RECOM_FILTER_SELECT   Addons>SelectValue (select_value_ops.cc:33-56): such an id's raw value
 *                         is replaced by `substitute`.
 *   RECOM_FILTER_STR_NE   Addons>GatherStrValueGenIndice (gather_str_value_gen_indice_ops.cc:31-58):
 *                         a string equal to `not_equal` (usually "") is dropped.
 * KEEP / SELECT apply to I32 / I64 values, STR_NE to STR values. A one-id-per-row column whose id
 * is dropped writes a zero row. */
typedef enum recom_filter_kind {
  RECOM_FILTER_NONE = 0,
  RECOM_FILTER_KEEP = 1,
  RECOM_FILTER_SELECT = 2,
  RECOM_FILTER_STR_NE = 3
} recom_filter_kind;

#define RECOM_MAX_INTERVALS 4

typedef struct recom_column_filter {
  int32_t kind;          /* recom_filter_kind */
  int32_t num_intervals; /* KEEP / SELECT: 1..RECOM_MAX_INTERVALS */
  int64_t lo[RECOM_MAX_INTERVALS];
  int64_t hi[RECOM_MAX_INTERVALS];
  int64_t substitute;    /* SELECT */
  int32_t not_equal_len; /* STR_NE: 0..16 */
  uint8_t not_equal[16];
  int32_t pad;
} recom_column_filter;

typedef struct recom_plan_config {
  int32_t out_dtype;    /* RECOM_F32 or RECOM_BF16: dtype of the concat tensor */
  int32_t out_width;    /* sum of dims (row length of the concat tensor), elements */
  int32_t num_slots;    /* number of blob slots a request carries */
  int32_t num_tables;   /* number of table pointers a request carries */
  int32_t check_ids;    /* !=0: bounds-check every row id; bad ids read row 0 and
                           raise the request's status word */
  int32_t device;       /* CUDA device ordinal */
  int32_t max_ctas_per_sm; /* 0 = library default */
  int32_t row_shard;    /* Mode B, 0 = whole tables: world | rank << 16. The rows of every table are
                           sharded over `world` ranks - row r lives on rank r % world at local row
                           r / world, and recom_request.tables point to this rank's shards. The
                           plan then writes PARTIAL SUMS over the rows it owns (every combiner
                           acts as SUM, rows of other ranks contribute zeros); the caller reduces
                           the [batch, out_width] tensors over the ranks (ncclReduceScatter) and
                           calls recom_plan_scale_rows on the rows it keeps. Descriptors keep
                           the GLOBAL table_rows. */
} recom_plan_config;

typedef struct recom_plan recom_plan;
typedef struct CUstream_st *recom_stream_t; /* == cudaStream_t */

/* One inference request. All device pointers are borrowed for the duration of
 * the stream-ordered work; host arrays are consumed before the call returns. */
typedef struct recom_request {
  int64_t batch;               /* B: rows of the concat tensor */
  const void *blob;            /* device: ConcatInputs blob (one H2D copy) */
  const int64_t *slot_offsets; /* host [num_slots]: byte offset of each slot */
  const int64_t *slot_elems;   /* host [num_slots]: element count of each slot */
  const void *const *tables;   /* host [num_tables]: device pointers, row-major */
  void *out;                   /* device [batch, out_width] of out_dtype */
  void *workspace;             /* device, >= recom_plan_workspace_bytes() */
  size_t workspace_bytes;
  int32_t *ids_out;   /* optional device [sum over columns of their id count], in
                         column order: the row id every gather used (parity hook) */
  int32_t *status_out; /* optional device int32: set !=0 on id-range violation */
  recom_stream_t stream;
  int64_t args_offset; /* byte offset inside `blob` of the argument trailer that
                          recom_concat_inputs_with_args appended (the slot offsets and element
                          counts, device readable): recom_plan_run then uploads nothing per
                          request. -1: the blob has no trailer; the two host arrays are staged
                          through pinned memory into the workspace (one small H2D copy). */
} recom_request;

RECOM_API const char *recom_last_error(void);
RECOM_API int recom_abi_version(void);

/* `boundaries` holds exact float bits (the reference embeds std::to_string
 * literals, cuda_emitter.cc:1806-1810). Arrays are copied; caller may free. */
RECOM_API recom_status recom_plan_create(const recom_column_desc *cols, int32_t num_cols,
                               const float *boundaries, int64_t num_boundaries,
                               const recom_plan_config *cfg, recom_plan **out);
/* The same with one recom_column_filter per column (`filters` may be null: no filtering). */
RECOM_API recom_status recom_plan_create_filtered(const recom_column_desc *cols, int32_t num_cols,
                                        const float *boundaries, int64_t num_boundaries,
                                        const recom_column_filter *filters,
                                        const recom_plan_config *cfg, recom_plan **out);
RECOM_API recom_status recom_plan_destroy(recom_plan *plan);

/* Serialized plan: the payload that replaces the reference's `dlpath` attr. */
RECOM_API recom_status recom_plan_serialize(const recom_plan *plan, void *buf,
                                  size_t buf_bytes, size_t *needed);
RECOM_API recom_status recom_plan_create_from_blob(const void *buf, size_t bytes,
                                         int32_t device, recom_plan **out);

RECOM_API recom_status recom_plan_num_columns(const recom_plan *plan, int32_t *out);
RECOM_API recom_status recom_plan_out_width(const recom_plan *plan, int32_t *out);

/* What an op kernel with the reference's signature (Addons>FeatureColumnProcess[WithSymbols],
 * feature_column_process_op_gpu.cu.cc:140-175) needs to turn its inputs into a recom_request and
 * its result into the reference's outputs:
 *   recom_plan_slot_layout   per blob slot: rank of the host tensor ConcatInputs packed (its dims in
 *                            `concated_shapes`: 1, or 2 for SparseTensor.indices) and element size;
 *   recom_plan_infer_batch   the batch size from the slots' element counts (the reference's
 *                            generated code reads it from the same shapes);
 *   recom_plan_output_shapes `output_shapes`: [batch, dim] per column, concat order;
 *   recom_plan_output_ptrs   `output_ptrs`: device int64[num_columns], the address of every
 *                            column's slice in row 0 of the concat tensor at `out` (row pitch
 *                            out_width elements: the columns are already in place, what
 *                            Addons>ConcatOutputs would have produced; one tiny kernel on `stream`). */
RECOM_API recom_status recom_plan_num_slots(const recom_plan *plan, int32_t *out);
RECOM_API recom_status recom_plan_slot_layout(const recom_plan *plan, int32_t *ranks,
                                    int32_t *elem_bytes);
RECOM_API recom_status recom_plan_infer_batch(const recom_plan *plan, const int64_t *slot_elems,
                                    int64_t *batch);
RECOM_API recom_status recom_plan_output_shapes(const recom_plan *plan, int64_t batch,
                                      int32_t *shapes);
/* Mode B (recom_plan_config.row_shard), after the cross-rank reduction: applies the mean /
 * sqrtn scaling (TF-CPU rule: x 1/m below 10 ids, / m otherwise) to rows [row_begin, row_begin +
 * rows) of the summed concat tensor; out_rows = address of row `row_begin`. `req`: the request the
 * partial sums were computed from (blob with the argument trailer, stream, workspace). */
RECOM_API recom_status recom_plan_scale_rows(const recom_plan *plan, const recom_request *req,
                                   int64_t row_begin, int64_t rows, void *out_rows);
RECOM_API recom_status recom_plan_output_ptrs(const recom_plan *plan, void *out, int64_t *ptrs_dev,
                                    recom_stream_t stream);

/* Upper bound of the per-request scratch for a batch of `batch` rows carrying
 * `total_ids` ids over all non-DENSE columns. */
RECOM_API recom_status recom_plan_workspace_bytes(const recom_plan *plan, int64_t batch,
                                        int64_t total_ids, size_t *out);

/* Enqueue the whole fused stage on req->stream. Re-entrant per plan. */
RECOM_API recom_status recom_plan_run(const recom_plan *plan, const recom_request *req);

/* Back-to-back requests on one stream. A plain request whose stage is ONE launch (no helper kernel,
 * no ids_out hook, a grid that fills the GPU) is launched with programmatic stream serialization:
 * when the operation in front of it on the stream is the previous request's kernel, the new grid
 * becomes resident while that one drains, runs its prologue (work records, column contexts, TMA
 * staging of the small tables, L2 prefetch of ITS blob - reads of its own request and of the plan
 * only) and executes griddepcontrol.wait before its first store, i.e. nothing of request i + 1 is
 * written before request i is complete and flushed. In front of any other operation (a copy, an
 * event, someone else's kernel without an early trigger) the launch is ordered as usual. The
 * request's blob and tables must therefore be in place when the operation in front of the launch
 * STARTS if that operation is a kernel that triggers dependents early - true for every TF / torch
 * kernel today; pass on = 0 for a producer that does (default: on; RECOM_B200_REQUEST_PDL=0). */
RECOM_API recom_status recom_plan_set_request_overlap(recom_plan *plan, int32_t on);

/* Mode A with the exchange fused into the kernel (recom_b200/sharding.py):
 * this plan computes the columns [col_offset, col_offset + out_width) of a global concat row
 * of `global_width` elements; rows [g*rows_per_rank, (g+1)*rows_per_rank) of the batch are
 * stored straight into peer_out[g], rank g's [rows_per_rank, global_width] tensor (peer-mapped
 * device memory, e.g. opened with recom_ipc_open), from the kernel's phase 2 over NVLink. No
 * collective and no concatenation copy follow; the caller only has to order the consumers after
 * all ranks' kernels. req->out is not used. batch <= world * rows_per_rank,
 * rows_per_rank >= 32, world <= 8. The reference has no counterpart (single GPU).
 *
 * Completion signal (optional, peer_flags != NULL): peer_flags[g] is rank g's array of
 * RECOM_SHARD_FLAG_WORDS uint32 words (peer mapped, zeroed before the first step). When the
 * last warp of this rank's kernel has finished, and after a system-scope fence, the kernel
 * stores `epoch` into word `rank` of every peer's array (release, system scope). A consumer on
 * rank g enqueues recom_shard_wait(peer_flags[g], world, epoch, stream) ahead of its readers: a
 * one-warp kernel that spins until words [0, world) have all reached `epoch`. The exchange then
 * needs no host barrier and no collective at all. Epochs of one flag array must increase by
 * one per step (wrap-around is handled); with the output tensors rotated over >= 2 sets, a
 * rank that has passed the wait of step j-1 may overwrite the set of step j - 2 or older.
 *
 * RECOM_SHARD_WAIT_IN_KERNEL: the kernel's last warp, after it has signalled, itself waits for
 * the signals of all ranks (10 s time-out, reported like recom_shard_wait's): the launch then ends
 * only when every rank's slices of THIS rank's rows have landed, i.e. compute, exchange and
 * barrier are one kernel and anything enqueued behind it on the stream may read the rows. Every
 * rank must launch the step without waiting for another rank's launch of the same step. */
#define RECOM_SHARD_FLAG_WORDS 16
#define RECOM_SHARD_WAIT_IN_KERNEL 1u
typedef struct recom_shard_out {
  int32_t world;
  int32_t col_offset;   /* elements */
  int32_t global_width; /* elements */
  int32_t rank;         /* this plan's rank: the flag word it signals */
  int64_t rows_per_rank;
  void *const *peer_out;   /* host [world] */
  void *const *peer_flags; /* host [world], or NULL: no completion signal */
  uint32_t epoch;
  uint32_t flags; /* RECOM_SHARD_WAIT_IN_KERNEL or 0 */
} recom_shard_out;
RECOM_API recom_status recom_plan_run_sharded(const recom_plan *plan, const recom_request *req,
                                              const recom_shard_out *shard);
/* Stream-ordered wait for the completion signals of all `world` ranks (see above). If a signal
 * has not arrived after timeout_ms (0: 10 s) the kernel stores 1 into word
 * RECOM_SHARD_FLAG_WORDS - 1 of local_flags and returns; the host may inspect that word. */
RECOM_API recom_status recom_shard_wait(void *local_flags, int32_t world, uint32_t epoch,
                                        uint32_t timeout_ms, void *stream);

/* Peer-visible device buffers for recom_plan_run_sharded: one cudaMalloc'ed buffer per rank,
 * exported as a 64-byte CUDA IPC handle that the other ranks of the node open. */
RECOM_API recom_status recom_ipc_alloc(size_t bytes, void **dev_ptr, uint8_t handle[64]);
RECOM_API recom_status recom_ipc_open(const uint8_t handle[64], void **dev_ptr);
RECOM_API recom_status recom_ipc_close(void *dev_ptr);
RECOM_API recom_status recom_ipc_free(void *dev_ptr);

/* Number of kernels the last recom_plan_run on this thread launched. */
RECOM_API int recom_last_launch_count(void);

/* Algorithmic bytes of the request just described:
 * inputs consumed + row-splits + compulsory table rows + output written. */
RECOM_API recom_status recom_plan_algorithmic_bytes(const recom_plan *plan,
                                          const recom_request *req,
                                          int64_t *bytes_out);

/* out[p, col_offsets[i] + e] = cols[i][p, e] for e < col_dims[i]. `cols` is a
 * host array of device pointers; dtype is the plan-independent element size. */
RECOM_API recom_status recom_scatter_columns(const void *const *cols,
                                   const int32_t *col_dims,
                                   const int32_t *col_offsets, int32_t num_cols,
                                   int64_t batch, int32_t out_width,
                                   int32_t elem_bytes, void *out,
                                   recom_stream_t stream);

/* The same for columns that live in HOST memory (Addons>ConcatOutputs' `host_inputs`,
 * concat_outputs_op_gpu.cu.cc:186-232): uploaded into `d_scratch` (device, 16-byte aligned, the
 * columns' bytes rounded up to 16 each), then scattered. */
/* SURVEY 8(f) N4 - the layer on the far side of the path: out[m, n] = relu(A[m, k] * W[n, k]^T +
 * bias[n]), the first dense layer of the top MLP behind the concat (examples/python/dlrm.py:175-182:
 * tf.layers.dense(units, relu); the reference leaves it to TF's MatMul / BiasAdd / Relu). A is the
 * bf16 concat tensor a RECOM_BF16 plan has just written (row stride lda elements), W the layer's
 * kernel stored [n, k] (K-major: TF's [k, n] kernel transposed once at load time), both device
 * pointers, 16-byte aligned, k and lda multiples of 8; bias (fp32[n], device) may be null; out is
 * fp32 [m, n]. One tcgen05 / TMA kernel on `stream` (csrc/dense_layer.cu). */
RECOM_API recom_status recom_dense_relu(const void *a_bf16, int64_t m, int32_t k, int64_t lda,
                                        const void *w_nk_bf16, int32_t n, const float *bias,
                                        int32_t relu, float *out, recom_stream_t stream);

RECOM_API recom_status recom_scatter_host_columns(const void *const *host_cols,
                                        const int32_t *col_dims,
                                        const int32_t *col_offsets, int32_t num_cols,
                                        int64_t batch, int32_t out_width,
                                        int32_t elem_bytes, void *out, void *d_scratch,
                                        recom_stream_t stream);

/* Sorted COO row ids (int64, stride in int64s) -> int32 row_splits[batch+1]. */
RECOM_API recom_status recom_segment_ids_to_row_splits(const int64_t *coo, int64_t nnz,
                                             int32_t stride, int64_t batch,
                                             int32_t *row_splits,
                                             recom_stream_t stream);

/* Device-side TF StringToHashBucketFast over a packed string slot (parity hook
 * and stand-alone op): out[i] = Fingerprint64(bytes[off[i]:off[i+1]]) % n. */
RECOM_API recom_status recom_string_to_hash_bucket_fast(const uint8_t *bytes,
                                              const int32_t *offsets, int64_t n,
                                              int64_t num_buckets, int64_t *out,
                                              recom_stream_t stream);
/* Raw device Fingerprint64 (uint64 written as int64 bits). */
RECOM_API recom_status recom_fingerprint64(const uint8_t *bytes, const int32_t *offsets,
                                 int64_t n, uint64_t *out,
                                 recom_stream_t stream);

/* ---- host-side pieces of the op surface (no GPU needed) ------------------------------ */

/* Addons>ConcatInputs (tensorflow_addons/custom_ops/concat_inputs/concat_inputs_ops.cc:42-77):
 * copies `num_inputs` host buffers into one blob, every slot 16-byte aligned and zero
 * padded, and reports each slot's byte offset. `blob == NULL` only computes `*needed`. */
RECOM_API recom_status recom_concat_inputs(const void *const *inputs, const int64_t *nbytes,
                                 int32_t num_inputs, void *blob, size_t blob_bytes,
                                 int64_t *offsets, size_t *needed);

/* The same, followed by the request's argument trailer: {magic, num_inputs, offsets[num_inputs],
 * elems[num_inputs]} (int64 each) at byte *args_offset of the blob (16-byte aligned, past the last
 * slot). `elems` are the element counts of the inputs (the products of the `shapes` the
 * reference op outputs). A request whose blob carries the trailer (recom_request.args_offset)
 * makes recom_plan_run upload nothing: the one H2D copy of the blob brings the arguments along
 * (the reference uploads KnlArgs and SymbolResults on every call, cuda_emitter.cc:2182-2232). */
RECOM_API recom_status recom_concat_inputs_with_args(const void *const *inputs,
                                           const int64_t *nbytes, const int64_t *elems,
                                           int32_t num_inputs, void *blob, size_t blob_bytes,
                                           int64_t *offsets, int64_t *args_offset,
                                           size_t *needed);

/* Addons>ShapeConstruct (tensorflow_addons/custom_ops/shape_construct/shape_construct_ops.cc:34-153):
 * run-time evaluation of size expressions over named symbols (num_segments, reshape
 * shapes, the `symbols` input of FeatureColumnProcessWithSymbols). The expression strings
 * are the ones the reference stores in the op's `exprs` attr (C syntax as printed by
 * SymEngine: + - * / %, parentheses, max/min/pow/abs/floor/ceiling); symbols are bound as C
 * `int`, `/` truncates, exactly like the code the reference compiles with g++. */
typedef struct recom_shape_construct recom_shape_construct;
RECOM_API recom_status recom_shape_construct_create(const char *const *symbols, int32_t num_symbols,
                                          const char *const *exprs, int32_t num_exprs,
                                          recom_shape_construct **out);
RECOM_API recom_status recom_shape_construct_eval(const recom_shape_construct *sc,
                                        const int64_t *inputs, int32_t num_inputs,
                                        int64_t *outputs, int32_t num_outputs);
RECOM_API recom_status recom_shape_construct_destroy(recom_shape_construct *sc);

#ifdef __cplusplus
}
#endif
#endif /* RECOM_B200_H_ */
