/*
 * annb.h -- C ABI of libannlite_b200.so: the B200 (sm_100a) implementation of AnnLite's
 * PQ-ADC / PQ-HNSW search path.
 *
 * This is the drop-in boundary.  Every entry point below replaces one
 * native call the reference makes through `annlite.pq_bind` (Cython) or
 * `annlite.hnsw_bind.Index` (pybind11); the reference interface it stands in for is cited as
 * file:line relative to the reference tree.  Plain pointers and sizes only -- no torch, no
 * C++ types.  All functions return 0 on success and a negative ANNB_E* code on failure;
 * annb_last_error() returns the message for the calling thread (the strings mirror the
 * reference's exception texts so the Python shim can raise the same exceptions).
 *
 * Memory spaces: every data pointer is accompanied by (or covered by) a `space` argument:
 * ANNB_HOST = ordinary host memory (numpy), ANNB_DEVICE = device memory on the handle's GPU
 * (e.g. torch `tensor.data_ptr()`).  Host pointers are borrowed for the duration of the call
 * only.  Work is enqueued on the handle's own CUDA stream; calls taking host output buffers
 * synchronise that stream before returning, calls with device outputs return asynchronously
 * (use annb_sync / annb_stream).
 *
 * There is no CPU fallback: every compute entry point fails with ANNB_ENODEVICE when no CUDA
 * device is usable.
 */
#ifndef ANNB_H_
#define ANNB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANNB_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define ANNB_API __attribute__((visibility("default")))
#else
#define ANNB_API
#endif

/* status codes */
#define ANNB_OK 0
#define ANNB_EINVAL (-1)     /* bad argument / wrong dimensionality                              */
#define ANNB_ENODEVICE (-2)  /* no usable CUDA device (there is no CPU path)                     */
#define ANNB_ECUDA (-3)      /* CUDA runtime error (message has the cudaError string)            */
#define ANNB_ENOMEM (-4)
#define ANNB_ESTATE (-5)     /* call order: codebook / codes / graph not set                      */
#define ANNB_EFEWRESULTS (-6)/* a query found < k results: hnsw_bindings.cpp:342-345 / :461-464   */
#define ANNB_EIO (-7)        /* "Cannot open file" / "Index seems to be corrupted or unsupported" */
#define ANNB_ECAPACITY (-8)  /* "The number of elements exceeds the specified limit" hnswalg.h:1133 */
#define ANNB_ENOTFOUND (-9)  /* "Label not found" hnswalg.h:878-880                               */
#define ANNB_ELIMIT (-10)    /* outside this build's limits (ef > ANNB_MAX_EF, code width ...)    */

#define ANNB_HOST 0
#define ANNB_DEVICE 1

/* metric: annlite/enums.py:25-28 -> space names 'l2' / 'ip' / 'cosine' (hnsw/index.py:196-202) */
#define ANNB_METRIC_L2 0
#define ANNB_METRIC_IP 1
#define ANNB_METRIC_COSINE 2

#define ANNB_MAX_EF 512 /* top-candidate list lives in registers: 16 entries x 32 lanes */

typedef struct annb_index annb_index_t;

/* ---- library / error -------------------------------------------------------------------- */
ANNB_API int annb_version(void);
ANNB_API const char *annb_last_error(void);
ANNB_API int annb_device_count(void);

/* ---- lifetime --------------------------------------------------------------------------- */

/* Index(space, dim) + _loadPQ geometry (hnsw_bindings.cpp:87-112, :851-928).
 * dim == n_subvectors * d_subvector is enforced (:883-891); code width follows n_clusters:
 * <=256 -> u8, <=65536 -> u16 (:909-927; u32 codes are rejected with ANNB_ELIMIT).
 * device >= 0 selects the GPU.  device == -1 creates a HOST-ONLY handle that can hold, load, save and
 * extend (annb_add_items_with_tables) a graph but cannot compute anything: every kernel-backed entry
 * point returns ANNB_ENODEVICE on it.  It exists for tooling and CPU-side tests of the container. */
ANNB_API int annb_create(int device, int metric, int dim, int n_subvectors, int n_clusters,
                annb_index_t **out);
ANNB_API int annb_destroy(annb_index_t *h);

/* The codebook a trained PQCodec hands to _loadPQ via get_codebook() (hnsw_bindings.cpp:892-905,
 * annlite/core/codec/pq.py:231-237): (n_subvectors, n_clusters, d_subvector) fp32, C-contiguous. */
ANNB_API int annb_set_codebook(annb_index_t *h, const float *codebook, int space);

/* cudaStream_t of the handle as an integer (for torch.cuda.ExternalStream) and a blocking sync. */
ANNB_API int annb_stream(annb_index_t *h, uint64_t *stream_out);
ANNB_API int annb_sync(annb_index_t *h);

/* ---- K1: ADC tables -------------------------------------------------------------------- */

/* pq_bind.batch_precompute_adc_table (bindings/pq_bindings.pyx:149-210; single query :85-145),
 * pq_bind.batch_precompute_adc_table_ip (:214-274) followed by the `1/n_clusters - .` epilogue of
 * PQCodec.get_dist_mat (annlite/core/codec/pq.py:293-325).  The metric of the handle selects the
 * formula.  queries: (B, dim) fp32; out: (B, n_subvectors, n_clusters) fp32.
 * normalize != 0 applies annlite.math.l2_normalize (annlite/math.py:6-18) to each query first,
 * `normalize` times (HnswIndex.search normalises twice for COSINE: hnsw/index.py:28-29 and
 * pq.py:309-310).  Arithmetic order and rounding match the reference (sequential over j,
 * no FMA contraction). */
ANNB_API int annb_adc_table(annb_index_t *h, const float *queries, int q_space, int64_t B, int normalize,
                   float *out, int out_space);

/* ---- K2: exhaustive ADC scan ------------------------------------------------------------ */

/* The flat code matrix PQIndex keeps in `_data` (annlite/core/index/pq_index.py:25-27,
 * flat_index.py:47-56): (n, n_subvectors) u8/u16.  Independent of the HNSW graph. */
ANNB_API int annb_set_codes(annb_index_t *h, const void *codes, int space, int64_t n);

/* pq_bind.dist_pqcodes_to_codebooks (bindings/pq_bindings.pyx:52-80): one (M,Ks) table against
 * all n codes -> n fp32 distances. */
ANNB_API int annb_scan(annb_index_t *h, const float *table, int t_space, float *out_dists, int out_space);

/* PQIndex.search for a batch (pq_index.py:29-56 + annlite/math.py:94-120): per query the k
 * smallest ADC distances over the code matrix, ascending; ties ordered by row index.
 * `tables` may be NULL, in which case they are built from `queries` with K1 (fused path).
 * ids: (B,k) int64 row indices, dists: (B,k) fp32 (squared L2 / raw IP form, no sqrt). */
ANNB_API int annb_scan_topk(annb_index_t *h, const float *queries, const float *tables, int in_space,
                   int64_t B, int k, int64_t *ids, float *dists, int out_space);

/* ---- HNSW graph: import / export / build ------------------------------------------------- */

/* Index.init_index(max_elements, M, ef_construction, random_seed) (hnsw_bindings.cpp:144-163,
 * :941-943; hnswalg.h:27-69).  Creates an empty graph with the reference's memory layout. */
ANNB_API int annb_init_graph(annb_index_t *h, int64_t max_elements, int M, int ef_construction,
                    uint64_t random_seed);

/* Index.load_index(path, max_elements) / Index.save_index(path): hnswlib's binary format
 * (hnswalg.h:708-736, :738-846) -- files are interchangeable with the reference. */
ANNB_API int annb_load_index(annb_index_t *h, const char *path, int64_t max_elements);
ANNB_API int annb_save_index(annb_index_t *h, const char *path);

/* Index.__setstate__ / createFromParams + setAnnData (hnsw_bindings.cpp:691-841): adopt a graph
 * given as the raw arrays of the reference's pickle dict. */
ANNB_API int annb_set_graph(annb_index_t *h, const uint8_t *data_level0, uint64_t data_level0_bytes,
                   uint64_t size_data_per_element, uint64_t offset_data, uint64_t label_offset,
                   const uint8_t *link_lists, uint64_t link_lists_bytes, const int32_t *element_levels,
                   int64_t n_element_levels, uint64_t size_links_per_element,
                   int64_t cur_element_count, int64_t max_elements, int32_t max_level,
                   uint32_t enterpoint_node, int max_M, int max_M0, int M, int ef_construction,
                   double mult);
/* The three array extents (bytes of data_level0 / link_lists, entries of element_levels) are what the
 * caller really holds; the state is rejected (ANNB_EINVAL) if cur_element_count or the levels ask for
 * more, and then checked structurally like a loaded file. */

/* Index.__getstate__ (hnsw_bindings.cpp:549-671): sizes first, then copy-out. */
ANNB_API int annb_graph_info(annb_index_t *h, int64_t *cur_element_count, int64_t *max_elements,
                    uint64_t *size_data_per_element, uint64_t *link_lists_bytes, int32_t *max_level,
                    uint32_t *enterpoint_node, int *max_M, int *max_M0, int *M, int *ef_construction,
                    double *mult);
ANNB_API int annb_get_graph(annb_index_t *h, uint8_t *data_level0, uint8_t *link_lists, int32_t *element_levels);

/* Index.add_items(data, ids, num_threads, dtables) (hnsw_bindings.cpp:216-300 -> hnswalg.h:1108-1235).
 * `vectors`: (n, dim) fp32 rows (already l2-normalised by the caller for COSINE, as pre_process
 * does); the library encodes them (PQCodec.encode, pq.py:158-177) unless `codes` is given, builds
 * each row's ADC table on the fly (the reference materialises (n,M,Ks) tables; this does not)
 * and inserts with the reference's algorithm, including its PQ-mode neighbour heuristic
 * (SURVEY.md section 0.2).  num_threads == 1 reproduces the reference's single-threaded graph
 * bit for bit; > 1 inserts concurrently like ParallelFor (hnsw_bindings.cpp:24-77). */
ANNB_API int annb_add_items(annb_index_t *h, const float *vectors, const void *codes, const uint64_t *labels,
                   int64_t n, int num_threads);

/* The literal shape of Index.add_items(data=codes, ids, num_threads, dtables) (hnsw_bindings.cpp:286-300):
 * caller-supplied PQ codes (n, n_subvectors) and the materialised (n, n_subvectors, n_clusters) fp32
 * tables, both host memory.  Kept for call-site compatibility; annb_add_items is the path that
 * does not need O(n*M*Ks) host memory. */
ANNB_API int annb_add_items_with_tables(annb_index_t *h, const void *codes, const float *tables,
                               const uint64_t *labels, int64_t n, int num_threads);

/* PQCodec.encode (pq.py:158-177): nearest codeword per subspace, first minimum. (n,dim)->(n,M). */
ANNB_API int annb_encode(annb_index_t *h, const float *vectors, int v_space, int64_t n, void *codes, int c_space);

ANNB_API int annb_resize_index(annb_index_t *h, int64_t new_max_elements); /* hnswalg.h:680-706           */
ANNB_API int annb_mark_deleted(annb_index_t *h, uint64_t label);           /* hnswalg.h:875-902           */
ANNB_API int annb_unmark_deleted(annb_index_t *h, uint64_t label);         /* hnswalg.h:908-934           */
ANNB_API int annb_element_count(annb_index_t *h, int64_t *out);            /* hnsw_bindings.cpp:987-992   */
/* Index.get_ids_list (hnsw_bindings.cpp:539-547): labels in internal-id order. */
ANNB_API int annb_get_labels(annb_index_t *h, uint64_t *labels_out, int64_t cap);
/* Index.get_items(ids) (hnsw_bindings.cpp:518-537): stored PQ codes of the given labels. */
ANNB_API int annb_get_codes(annb_index_t *h, const uint64_t *labels, int64_t n, void *codes_out);

/* ---- K3: HNSW search ---------------------------------------------------------------------- */

/* Index.knn_query(data, k, num_threads, dtables) (hnsw_bindings.cpp:303-391 -> hnswalg.h:1237-1295)
 * and Index.knn_query_with_filter(data, filters, k, num_threads, dtables) (:393-516 ->
 * hnswalg.h:1297-1361).
 *   queries / tables : exactly one non-NULL.  `tables` = the (B,M,Ks) `dtables` argument;
 *                      `queries` = (B,dim) fp32, tables are then built on the device (K1) with
 *                      `normalize` rounds of l2_normalize first.
 *   ef               : Index.set_ef value; the search uses max(ef, k) (hnswalg.h:1279).
 *   filter_labels    : NULL = knn_query; else the `filters` array of allowed labels (host or
 *                      device per `filter_space`), n_filter entries.  Membership is exact (the
 *                      reference's binary-fuse-16 filter approximates the same set).
 *   labels_out       : (B,k) uint64, dists_out: (B,k) fp32, rows ascending by (dist, label) --
 *                      the order produced by the unload loop at hnsw_bindings.cpp:346-351.
 *   stats_out        : optional (B,3) int64 {hops, neighbours listed, distances computed}:
 *                      metric_hops / metric_distance_computations of hnswalg.h:279-282,:1256-1257.
 * Returns ANNB_EFEWRESULTS if any query found fewer than k results (outputs are still filled,
 * missing slots = UINT64_MAX / +inf), mirroring the RuntimeError of the reference. */
ANNB_API int annb_search(annb_index_t *h, const float *queries, const float *tables, int in_space, int64_t B,
                int normalize, int k, int ef, const uint64_t *filter_labels, int filter_space,
                int64_t n_filter, uint64_t *labels_out, float *dists_out, int out_space,
                int64_t *stats_out);

/* Exhaustive ADC over a label subset of the INDEXED nodes (the brute-force route for
 * very selective filters that HnswIndex.search leaves as a TODO, annlite/core/index/hnsw/index.py:152): the
 * codes of the nodes whose labels are listed are gathered from the graph, scanned with K2 for every query and
 * the k best returned as (labels, dists), ascending (dist, label order of the subset list).  Unknown and
 * deleted labels are skipped; fewer than k candidates -> missing slots = UINT64_MAX / +inf and
 * ANNB_EFEWRESULTS.  Opt-in from the Python side (it changes WHICH results come back compared with the
 * reference's filtered graph walk: these are exact over the subset). */
ANNB_API int annb_scan_subset(annb_index_t *h, const float *queries, int in_space, int64_t B, int normalize, int k,
                     const uint64_t *subset_labels, int64_t n_subset, uint64_t *labels_out, float *dists_out);

/* Streaming form of annb_search for serving loops (no stats).  annb_search_submit enqueues upload (host inputs),
 * the walk (tables built inside the kernel) and download (host outputs) of one batch on one of two internal lanes and
 * returns a ticket at once; annb_search_wait blocks until that batch is complete and reports ANNB_EFEWRESULTS like
 * annb_search.  Two batches can be in flight, so the copies of batch i+1 and the under-occupied tail of batch i's
 * launch overlap.  Buffers of a submitted batch (queries, filter labels, outputs) belong to the library until its
 * wait returns; a third submit first waits for the oldest ticket of its lane.  An index with deleted nodes is served
 * by the deletion-aware walk (searchBaseLayerST<true>, hnswalg.h:243-329).
 * annb_search_submit_filtered = knn_query_with_filter (bindings/hnsw_bindings.cpp:393-516) in the same form: the
 * filter label list is uploaded and turned into the by-id bitmap on the batch's own lane; filter_labels == NULL means
 * no filter, a non-NULL list with n_filter == 0 admits nothing.  Queries whose walk outgrew the register lists are
 * re-run on the bitmap walk inside annb_search_wait (annb_fallback_queries counts them) -- or earlier, by the next
 * blocking call on the handle (annb_search, annb_scan_topk, annb_adc_table, annb_add_items, ...): those share device
 * scratch with lane 0 and first settle any such batch, whose ticket stays valid for annb_search_wait. */
ANNB_API int annb_search_submit(annb_index_t *h, const float *queries, int in_space, int64_t B, int normalize, int k,
                       int ef, uint64_t *labels_out, float *dists_out, int out_space, int *ticket_out);
ANNB_API int annb_search_submit_filtered(annb_index_t *h, const float *queries, int in_space, int64_t B, int normalize,
                                int k, int ef, const uint64_t *filter_labels, int filter_space, int64_t n_filter,
                                uint64_t *labels_out, float *dists_out, int out_space, int *ticket_out);
ANNB_API int annb_search_wait(annb_index_t *h, int ticket);

/* ---- shard merge (CellContainer.ivf_search merge rule, annlite/container.py:130-138) ---------- */

/* G per-shard result lists (G,B,k) (as all-gathered over NCCL) -> (B,k) global top-k by
 * (dist, label).  All pointers are device pointers on the handle's GPU. */
ANNB_API int annb_merge_topk(annb_index_t *h, const uint64_t *labels_gbk, const float *dists_gbk, int G,
                    int64_t B, int k, uint64_t *labels_out, float *dists_out);

/* The sharded step without a host synchronisation (annlite/container.py:88-144 spread over G GPUs): a rank's
 * walk (annb_search_submit with DEVICE outputs) writes its (B,k) fp32 distances and (B,k) u64 labels into ONE
 * packed buffer ([dists | pad to 8][labels], rank_stride_bytes long), ONE all-gather of that buffer over NCCL is
 * enqueued on the same lane's stream (annb_lane_stream(ticket & 1) -> torch.cuda.ExternalStream), and
 * annb_merge_topk_packed enqueues the merge of the G gathered buffers on that stream again.  annb_search_wait
 * (ticket) then covers walk + gather + merge.  `lane` = ticket & 1. */
ANNB_API int annb_merge_topk_packed(annb_index_t *h, const void *packed_gathered, int G, int64_t B, int k,
                           int64_t rank_stride_bytes, int64_t labels_offset_bytes, uint64_t *labels_out,
                           float *dists_out, int lane);
ANNB_API int annb_lane_stream(annb_index_t *h, int lane, uint64_t *stream_out);

/* ---- introspection for bench / roofline ------------------------------------------------------ */
/* Device time of the last search / scan / table kernel in milliseconds (CUDA events on the
 * handle's stream) and launch counters since creation. */
ANNB_API int annb_last_kernel_ms(annb_index_t *h, float *table_ms, float *search_ms, float *scan_ms);
ANNB_API int annb_launch_count(annb_index_t *h, int64_t *out);
/* number of filtered/deleted batches that outgrew the single-list walk and were re-run on the bitmap walk */
ANNB_API int annb_fallback_count(annb_index_t *h, int64_t *out);
/* ... and how many queries those re-runs covered: only the queries whose list overflowed are redone (the reference's
 * searchBaseLayerSTWithFilter, hnswalg.h:332-440, has no list to overflow) */
ANNB_API int annb_fallback_queries(annb_index_t *h, int64_t *out);
/* How often the device copy of the graph was re-derived from the host graph in full, and how often only the records
 * a small host insertion rewrote were uploaded (the reference searches the structure it inserts into: hnswalg.h:1108;
 * here insertions happen on the host graph and the walk layout on the device follows). */
ANNB_API int annb_sync_counts(annb_index_t *h, int64_t *full_syncs, int64_t *patches);
/* Tuning / A-B / test switches (none changes results): "chunks" (host-buffer pipeline depth of annb_search, 1 = off),
 * "walk_kernel" (plain search: 0 hnsw_walk4 fused, 2 hnsw_walk4 over K1 tables, 1 round-1 kernels), "flagged_kernel"
 * (filter / deletions: 0 hnsw_walk4f where it applies, 1 hnsw_walk_flagged), "flagged_en" (force hnsw_walk4f's
 * traversed-only list to 32 x value entries), "flagged_epl" (force hnsw_walk_flagged with 32 x value entries),
 * "force_general" (1 = filtered route without a filter, 2 = bitmap walk), "scan_kernel" (K2: 0 query-tiled where it
 * applies, 1 round-1 kernel, 2 tiled even for small inputs, 3 first tiled version), "prefetch", "warps_per_cta",
 * "ctas_per_sm", "gpu_build", "gpu_build_frac", "timing", "ip_raw", "dump_tables", "reset_counters". */
ANNB_API int annb_set_option(annb_index_t *h, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* ANNB_H_ */
