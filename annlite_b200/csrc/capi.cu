// capi.cu -- the extern "C" boundary of libannlite_b200.so (see include/annb.h).
#include <math_constants.h>
#include <stdarg.h>

#include <algorithm>
#include <limits>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "annb_internal.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void annb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}


int annb_scratch(annb_index *h, int slot, size_t bytes, void **out) {
  if (bytes == 0) bytes = 16;
  if (h->scratch_cap[slot] < bytes) {
    if (h->d_scratch[slot]) {
      ANNB_CUDA(cudaStreamSynchronize(h->stream));
      if (h->stream2) ANNB_CUDA(cudaStreamSynchronize(h->stream2));
      ANNB_CUDA(cudaFree(h->d_scratch[slot]));
      h->d_scratch[slot] = nullptr;
      h->scratch_cap[slot] = 0;
    }
    size_t cap = bytes + bytes / 4;
    cap = (cap + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&h->d_scratch[slot], cap);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cap = (bytes + 255) / 256 * 256;
      e = cudaMalloc(&h->d_scratch[slot], cap);
    }
    if (e != cudaSuccess) ANNB_FAIL(ANNB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", cap, cudaGetErrorString(e));
    h->scratch_cap[slot] = cap;
  }
  *out = h->d_scratch[slot];
  return ANNB_OK;
}

int annb_pinned(annb_index *h, int slot, size_t bytes, void **out) {
  if (h->pinned_cap[slot] < bytes) {
    if (h->h_pinned[slot]) {
      ANNB_CUDA(cudaStreamSynchronize(h->stream));
      if (h->stream2) ANNB_CUDA(cudaStreamSynchronize(h->stream2));
      ANNB_CUDA(cudaFreeHost(h->h_pinned[slot]));
      h->h_pinned[slot] = nullptr;
      h->pinned_cap[slot] = 0;
    }
    ANNB_CUDA(cudaMallocHost(&h->h_pinned[slot], bytes));
    h->pinned_cap[slot] = bytes;
  }
  *out = h->h_pinned[slot];
  return ANNB_OK;
}

// Host-only handles (device == -1) exist for the graph container alone -- file I/O, pickle state and
// the host-side insertion algorithm fed with caller-supplied tables.  Every compute entry point
// (tables, scan, encode, search, merge, add_items from vectors) requires a GPU: ANNB_NEED_GPU.
#define ANNB_ENTER(h)                                                        \
  if (!(h)) ANNB_FAIL(ANNB_EINVAL, "null index handle");                     \
  std::lock_guard<std::mutex> _lk((h)->mu);                                  \
  if ((h)->device >= 0) ANNB_CUDA(cudaSetDevice((h)->device))
#define ANNB_NEED_GPU(h) \
  if ((h)->device < 0) ANNB_FAIL(ANNB_ENODEVICE, "this handle is host-only (graph I/O); compute needs a CUDA device: annlite_b200 has no CPU fallback")

#define ANNB_TRY(expr)       \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

// stage an input: returns a device pointer (the caller's own if it already is one)
static int stage_in(annb_index *h, const void *p, int space, size_t bytes, int slot, const void **out) {
  if (space == ANNB_DEVICE) {
    *out = p;
    return ANNB_OK;
  }
  void *d;
  ANNB_TRY(annb_scratch(h, slot, bytes, &d));
  ANNB_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, h->stream));
  *out = d;
  return ANNB_OK;
}

extern "C" {

int annb_version(void) { return ANNB_VERSION; }
const char *annb_last_error(void) { return g_err; }

int annb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int annb_create(int device, int metric, int dim, int n_subvectors, int n_clusters, annb_index_t **out) {
  if (!out) ANNB_FAIL(ANNB_EINVAL, "out is null");
  *out = nullptr;
  if (metric < ANNB_METRIC_L2 || metric > ANNB_METRIC_COSINE) ANNB_FAIL(ANNB_EINVAL, "Space name must be one of l2, ip, or cosine.");
  if (dim <= 0 || n_subvectors <= 0 || n_clusters <= 0) ANNB_FAIL(ANNB_EINVAL, "dim, n_subvectors and n_clusters must be positive");
  if (dim % n_subvectors != 0)
    ANNB_FAIL(ANNB_EINVAL, "Initialization Error, expect HNSW.dim == PQ.n_subvector*PQ.d_subvector, but got:\nHNSW.dim =%d, n_subvectors=%d", dim, n_subvectors);
  if (n_clusters > 65536)
    ANNB_FAIL(ANNB_ELIMIT, "PQ clustering exceed the maximum, annlite set the maximum of clusters = 65536, but got PQ.n_clusters=%d", n_clusters);
  if (device != -1) {
    int ndev = annb_device_count();
    if (ndev <= 0) ANNB_FAIL(ANNB_ENODEVICE, "no CUDA device available: annlite_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) ANNB_FAIL(ANNB_EINVAL, "device %d out of range (have %d)", device, ndev);
    ANNB_CUDA(cudaSetDevice(device));
  }
  annb_index *h = new annb_index();
  h->device = device;
  h->metric = metric;
  h->dim = dim;
  h->M = n_subvectors;
  h->Ks = n_clusters;
  h->ds = dim / n_subvectors;
  h->code_bytes = n_clusters <= 256 ? 1 : 2;
  if ((size_t)h->M * h->code_bytes > 128) {
    delete h;
    ANNB_FAIL(ANNB_ELIMIT, "code rows wider than 128 bytes are not supported (n_subvectors=%d)", n_subvectors);
  }
  if (device < 0) {
    *out = h;
    return ANNB_OK;
  }
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
  for (int i = 0; i < 6 && e == cudaSuccess; i++) e = cudaEventCreate(&h->ev[i]);
  if (e != cudaSuccess) {
    annb_set_error("CUDA error %s while creating stream/events", cudaGetErrorString(e));
    delete h;
    return ANNB_ECUDA;
  }
  *out = h;
  return ANNB_OK;
}

int annb_destroy(annb_index_t *h) {
  if (!h) return ANNB_OK;
  if (h->device < 0) {
    delete h;
    return ANNB_OK;
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &p : h->d_scratch)
    if (p) cudaFree(p);
  for (auto &p : h->h_pinned)
    if (p) cudaFreeHost(p);
  if (h->d_codebook) cudaFree(h->d_codebook);
  if (h->d_codebook_t) cudaFree(h->d_codebook_t);
  if (h->d_codes) cudaFree(h->d_codes);
  if (h->d_rec0) cudaFree(h->d_rec0);
  if (h->d_up) cudaFree(h->d_up);
  if (h->d_labels) cudaFree(h->d_labels);
  if (h->d_deleted) cudaFree(h->d_deleted);
  for (int i = 0; i < 6; i++)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return ANNB_OK;
}

int annb_set_codebook(annb_index_t *h, const float *codebook, int space) {
  ANNB_ENTER(h);
  if (!codebook) ANNB_FAIL(ANNB_EINVAL, "Passed PQ class is none");
  const size_t bytes = (size_t)h->M * h->Ks * h->ds * sizeof(float);
  ANNB_NEED_GPU(h);
  if (!h->d_codebook) ANNB_CUDA(cudaMalloc(&h->d_codebook, bytes));
  ANNB_CUDA(cudaMemcpyAsync(h->d_codebook, codebook, bytes, space == ANNB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                            h->stream));
  h->h_codebook.resize((size_t)h->M * h->Ks * h->ds);
  if (space == ANNB_DEVICE)
    ANNB_CUDA(cudaMemcpyAsync(h->h_codebook.data(), codebook, bytes, cudaMemcpyDeviceToHost, h->stream));
  else
    memcpy(h->h_codebook.data(), codebook, bytes);
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  // transposed copy for the table build inside the walk (hnsw_walk4): [m][j/V][c][V], V floats per load, so
  // that the 32 lanes of a warp (32 consecutive codewords c) read one contiguous 32*V*4-byte run
  h->cb_vec = (h->ds % 4 == 0) ? 4 : (h->ds % 2 == 0 ? 2 : 0);
  if (h->cb_vec) {
    const int V = h->cb_vec, nv = h->ds / V;
    std::vector<float> t(h->h_codebook.size());
    for (int m = 0; m < h->M; m++)
      for (int c = 0; c < h->Ks; c++)
        for (int j = 0; j < h->ds; j++)
          t[(((size_t)m * nv + j / V) * h->Ks + c) * V + j % V] = h->h_codebook[((size_t)m * h->Ks + c) * h->ds + j];
    if (!h->d_codebook_t) ANNB_CUDA(cudaMalloc(&h->d_codebook_t, bytes));
    ANNB_CUDA(cudaMemcpyAsync(h->d_codebook_t, t.data(), bytes, cudaMemcpyHostToDevice, h->stream));
    ANNB_CUDA(cudaStreamSynchronize(h->stream));
  }
  return ANNB_OK;
}

int annb_stream(annb_index_t *h, uint64_t *stream_out) {
  if (!h || !stream_out) ANNB_FAIL(ANNB_EINVAL, "null argument");
  *stream_out = (uint64_t)(uintptr_t)h->stream;
  return ANNB_OK;
}

static int lane_wait(annb_index *h, int lane);
static int lane_rerun_flagged(annb_index *h, int lane, int64_t n_over);

// A streamed filtered batch whose walk flagged queries is finished (those queries re-run on the bitmap walk) out of
// the lane's own scratch: queries, filter bitmap, device outputs.  Lane 0 shares these buffers with the blocking entry
// points, so they settle such a batch before they touch them; its ticket stays open for annb_search_wait, which then
// only reports `< k` results.
static int settle_flagged_lanes(annb_index *h) {
  for (int lane = 0; lane < 2; lane++) {
    annb_index::AsyncLane &L = h->lanes[lane];
    if (!L.busy || !L.flagged) continue;
    ANNB_CUDA(cudaStreamSynchronize(lane ? h->stream2 : h->stream));
    int64_t n_over = 0;
    for (int64_t b = 0; b < L.B; b++) n_over += L.hfound[b] < 0;
    if (n_over) ANNB_TRY(lane_rerun_flagged(h, lane, n_over));
    L.flagged = false;
  }
  return ANNB_OK;
}

int annb_sync(annb_index_t *h) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  int rc0 = lane_wait(h, 0), rc1 = lane_wait(h, 1);
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  ANNB_CUDA(cudaStreamSynchronize(h->stream2));
  return rc0 ? rc0 : rc1;
}

// ---- K1 ---------------------------------------------------------------------------------------
// queries (host/device) -> device tables; `d_tables_out` receives the device pointer used.
// queries (host/device) -> device queries, normalised `normalize` times (never the caller's own buffer)
static int stage_queries(annb_index *h, const float *queries, int q_space, int64_t B, int normalize, int slot, const float **dq_out) {
  if (!h->d_codebook) ANNB_FAIL(ANNB_ESTATE, "Please train the PQ before using HNSW quantization backend");
  const size_t qbytes = (size_t)B * h->dim * sizeof(float);
  const float *dq;
  if (normalize > 0 && q_space == ANNB_DEVICE) {  // never modify the caller's buffer
    void *d;
    ANNB_TRY(annb_scratch(h, slot, qbytes, &d));
    ANNB_CUDA(cudaMemcpyAsync(d, queries, qbytes, cudaMemcpyDeviceToDevice, h->stream));
    dq = (const float *)d;
  } else {
    ANNB_TRY(stage_in(h, queries, q_space, qbytes, slot, (const void **)&dq));
  }
  for (int r = 0; r < normalize; r++) ANNB_TRY(launch_l2_normalize(h, const_cast<float *>(dq), B, h->dim));
  *dq_out = dq;
  return ANNB_OK;
}

// queries (host/device) -> device tables (K1 as a kernel of its own: the literal pq_bind calls, K2, insertion)
static int build_tables(annb_index *h, const float *queries, int q_space, int64_t B, int normalize, float *d_tables) {
  const float *dq;
  if (h->opt_timing) cudaEventRecord(h->ev[0], h->stream);
  ANNB_TRY(stage_queries(h, queries, q_space, B, normalize, S_QUERIES, &dq));
  ANNB_TRY(launch_adc_table(h, dq, B, d_tables));
  if (h->opt_timing) cudaEventRecord(h->ev[1], h->stream);
  return ANNB_OK;
}

// fill the fused-build fields of a SearchParams (hnsw_walk4 builds each query's table in shared memory)
static void fuse_params(annb_index *h, SearchParams &p, const float *d_queries) {
  p.tables = nullptr;
  p.queries = d_queries;
  p.cbt = h->d_codebook_t;
  p.cb_vec = h->cb_vec;
  p.ds = h->ds;
  p.is_ip = h->metric != ANNB_METRIC_L2;
  p.bias = h->opt_ip_raw ? 0.f : (float)(1.0 / (double)h->Ks);
}

int annb_adc_table(annb_index_t *h, const float *queries, int q_space, int64_t B, int normalize, float *out, int out_space) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_TRY(settle_flagged_lanes(h));
  if (B < 0 || (B > 0 && (!queries || !out))) ANNB_FAIL(ANNB_EINVAL, "null queries/out");
  if (B == 0) return ANNB_OK;
  const size_t tbytes = (size_t)B * h->M * h->Ks * sizeof(float);
  float *d_tables = out;
  if (out_space != ANNB_DEVICE) ANNB_TRY(annb_scratch(h, S_TABLES, tbytes, (void **)&d_tables));
  ANNB_TRY(build_tables(h, queries, q_space, B, normalize, d_tables));
  if (out_space != ANNB_DEVICE) {
    ANNB_CUDA(cudaMemcpyAsync(out, d_tables, tbytes, cudaMemcpyDeviceToHost, h->stream));
    ANNB_CUDA(cudaStreamSynchronize(h->stream));
  }
  return ANNB_OK;
}

// ---- K2 ---------------------------------------------------------------------------------------
// A code is an index into a Ks-entry table row.  Where the code type can hold more than Ks (Ks < 256 with u8
// codes, ...) a value from outside -- a caller's array, a file -- is checked once on entry: the kernels and the
// host builder index the table with it unchecked, as the reference does (space_pq.h:30-35).
static int check_code_rows(annb_index *h, const uint8_t *rows, int64_t n, size_t stride) {
  const uint64_t full = h->code_bytes == 1 ? 256ull : (h->code_bytes == 2 ? 65536ull : (1ull << 32));
  if ((uint64_t)h->Ks >= full) return ANNB_OK;
  for (int64_t i = 0; i < n; i++) {
    const uint8_t *r = rows + (size_t)i * stride;
    for (int m = 0; m < h->M; m++) {
      uint32_t v = 0;
      memcpy(&v, r + (size_t)m * h->code_bytes, h->code_bytes);
      if (v >= (uint32_t)h->Ks)
        ANNB_FAIL(ANNB_EINVAL, "PQ code %u (row %lld, subvector %d) is not below n_clusters=%d", v, (long long)i, m, h->Ks);
    }
  }
  return ANNB_OK;
}

int annb_set_codes(annb_index_t *h, const void *codes, int space, int64_t n) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  if (n < 0 || (n > 0 && !codes)) ANNB_FAIL(ANNB_EINVAL, "null codes");
  if (space != ANNB_DEVICE) ANNB_TRY(check_code_rows(h, (const uint8_t *)codes, n, (size_t)h->M * h->code_bytes));
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->d_codes) {
    ANNB_CUDA(cudaFree(h->d_codes));
    h->d_codes = nullptr;
  }
  h->n_codes = 0;
  if (n == 0) return ANNB_OK;
  const size_t bytes = (size_t)n * h->M * h->code_bytes;
  ANNB_CUDA(cudaMalloc(&h->d_codes, bytes + 16));
  ANNB_CUDA(cudaMemcpyAsync(h->d_codes, codes, bytes, space == ANNB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                            h->stream));
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  h->n_codes = n;
  return ANNB_OK;
}

int annb_scan(annb_index_t *h, const float *table, int t_space, float *out_dists, int out_space) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_TRY(settle_flagged_lanes(h));
  if (!table || !out_dists) ANNB_FAIL(ANNB_EINVAL, "null table/out");
  if (!h->d_codes) ANNB_FAIL(ANNB_ESTATE, "no code matrix: call annb_set_codes first");
  const float *dt;
  ANNB_TRY(stage_in(h, table, t_space, (size_t)h->M * h->Ks * sizeof(float), S_TABLES, (const void **)&dt));
  float *dout = out_dists;
  if (out_space != ANNB_DEVICE) ANNB_TRY(annb_scratch(h, S_OUT_D, (size_t)h->n_codes * sizeof(float), (void **)&dout));
  if (h->opt_timing) cudaEventRecord(h->ev[4], h->stream);
  ANNB_TRY(launch_scan(h, dt, dout));
  if (h->opt_timing) cudaEventRecord(h->ev[5], h->stream);
  if (out_space != ANNB_DEVICE) {
    ANNB_CUDA(cudaMemcpyAsync(out_dists, dout, (size_t)h->n_codes * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    ANNB_CUDA(cudaStreamSynchronize(h->stream));
  }
  return ANNB_OK;
}

int annb_scan_topk(annb_index_t *h, const float *queries, const float *tables, int in_space, int64_t B, int k, int64_t *ids,
                   float *dists, int out_space) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_TRY(settle_flagged_lanes(h));
  if ((queries == nullptr) == (tables == nullptr)) ANNB_FAIL(ANNB_EINVAL, "exactly one of queries / tables must be given");
  if (B < 0 || k <= 0 || !ids || !dists) ANNB_FAIL(ANNB_EINVAL, "bad B/k/outputs");
  if (!h->d_codes) ANNB_FAIL(ANNB_ESTATE, "no code matrix: call annb_set_codes first");
  if (B == 0) return ANNB_OK;
  const size_t tbytes = (size_t)B * h->M * h->Ks * sizeof(float);
  const float *dt;
  if (tables) {
    ANNB_TRY(stage_in(h, tables, in_space, tbytes, S_TABLES, (const void **)&dt));
  } else {
    float *t;
    ANNB_TRY(annb_scratch(h, S_TABLES, tbytes, (void **)&t));
    // PQIndex.search builds its table with precompute_adc (L2 form, no normalisation): pq.py:200-224
    ANNB_TRY(build_tables(h, queries, in_space, B, 0, t));
    dt = t;
  }
  int64_t *dids = ids;
  float *dd = dists;
  if (out_space != ANNB_DEVICE) {
    ANNB_TRY(annb_scratch(h, S_OUT_L, (size_t)B * k * sizeof(int64_t), (void **)&dids));
    ANNB_TRY(annb_scratch(h, S_OUT_D, (size_t)B * k * sizeof(float), (void **)&dd));
  }
  if (h->opt_timing) cudaEventRecord(h->ev[4], h->stream);
  ANNB_TRY(launch_scan_topk(h, dt, B, k, dids, dd));
  if (h->opt_timing) cudaEventRecord(h->ev[5], h->stream);
  if (out_space != ANNB_DEVICE) {
    ANNB_CUDA(cudaMemcpyAsync(ids, dids, (size_t)B * k * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    ANNB_CUDA(cudaMemcpyAsync(dists, dd, (size_t)B * k * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    ANNB_CUDA(cudaStreamSynchronize(h->stream));
  }
  return ANNB_OK;
}

// ---- graph ------------------------------------------------------------------------------------
int annb_init_graph(annb_index_t *h, int64_t max_elements, int M, int ef_construction, uint64_t random_seed) {
  ANNB_ENTER(h);
  // M = 1 makes the level multiplier 1/ln(M) infinite (hnswalg.h:47) and the reference's level draw undefined
  if (max_elements < 0 || M < 2 || M > 1024) ANNB_FAIL(ANNB_EINVAL, "bad max_elements / M (2 <= M <= 1024)");
  ANNB_TRY(h->g.init(max_elements, M, ef_construction, random_seed, (size_t)h->M * h->code_bytes));
  h->dev_dirty = true;
  return ANNB_OK;
}

int annb_load_index(annb_index_t *h, const char *path, int64_t max_elements) {
  ANNB_ENTER(h);
  if (!path) ANNB_FAIL(ANNB_EINVAL, "null path");
  ANNB_TRY(h->g.load_file(path, max_elements, (size_t)h->M * h->code_bytes));
  h->dev_dirty = true;
  if (int rc = check_code_rows(h, h->g.level0 + h->g.offset_data, h->g.count.load(), h->g.size_per_elem)) {
    h->g.clear();
    return rc;
  }
  return ANNB_OK;
}

int annb_save_index(annb_index_t *h, const char *path) {
  ANNB_ENTER(h);
  if (!path) ANNB_FAIL(ANNB_EINVAL, "null path");
  if (!h->g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  return h->g.save_file(path);
}

int annb_set_graph(annb_index_t *h, const uint8_t *data_level0, uint64_t data_level0_bytes, uint64_t size_data_per_element,
                   uint64_t offset_data, uint64_t label_offset, const uint8_t *link_lists, uint64_t link_lists_bytes,
                   const int32_t *element_levels, int64_t n_element_levels, uint64_t size_links_per_element,
                   int64_t cur_element_count, int64_t max_elements, int32_t max_level, uint32_t enterpoint_node, int max_M,
                   int max_M0, int M, int ef_construction, double mult) {
  ANNB_ENTER(h);
  if (cur_element_count < 0 || cur_element_count > 0xfffffffell || n_element_levels < cur_element_count ||
      (size_data_per_element && data_level0_bytes / size_data_per_element < (uint64_t)cur_element_count) || M < 2 || M > 1024 ||
      ef_construction < 1 || !std::isfinite(mult) || mult < 0.0 || mult > 64.0)
    ANNB_FAIL(ANNB_EINVAL, "graph state is inconsistent with the sizes of its arrays");
  HostGraph &g = h->g;
  const size_t crow = (size_t)h->M * h->code_bytes;
  if (max_elements < cur_element_count) max_elements = cur_element_count;
  ANNB_TRY(g.init(max_elements, M, ef_construction, 100, crow));
  if (g.size_per_elem != size_data_per_element) ANNB_FAIL(ANNB_EINVAL, "Invalid value of size_data_per_element_ ");
  if (g.label_offset != label_offset) ANNB_FAIL(ANNB_EINVAL, "Invalid value of label_offset_ ");
  if (g.offset_data != offset_data) ANNB_FAIL(ANNB_EINVAL, "Invalid value of offsetData_ ");
  if (g.maxM != max_M) ANNB_FAIL(ANNB_EINVAL, "Invalid value of maxM_ ");
  if (g.maxM0 != max_M0) ANNB_FAIL(ANNB_EINVAL, "Invalid value of maxM0_ ");
  if (g.size_links_per_elem != size_links_per_element) ANNB_FAIL(ANNB_EINVAL, "Invalid value of size_links_per_element_ ");
  if (cur_element_count < 0 || (cur_element_count > 0 && (!data_level0 || !element_levels))) ANNB_FAIL(ANNB_EINVAL, "bad arguments");
  g.mult = mult;
  g.maxlevel = max_level;
  g.enterpoint = enterpoint_node;
  const size_t n = (size_t)cur_element_count;
  if (n) memcpy(g.level0, data_level0, n * g.size_per_elem);
  size_t off = 0;
  for (size_t i = 0; i < n; i++) {
    if (element_levels[i] < 0 || element_levels[i] >= 63 || (element_levels[i] > 0 && !link_lists))
      ANNB_FAIL(ANNB_EINVAL, "Invalid value of element_levels_[%zu]", i);
    g.levels[i] = element_levels[i];
    if (element_levels[i] > 0) {
      const size_t sz = g.size_links_per_elem * (size_t)element_levels[i];
      if (off + sz > link_lists_bytes) {
        for (size_t j = 0; j < i; j++) {
          free(g.upper[j]);
          g.upper[j] = nullptr;
          g.levels[j] = 0;
        }
        g.levels[i] = 0;
        ANNB_FAIL(ANNB_EINVAL, "element_levels_ ask for more link-list bytes than link_lists holds");
      }
      g.upper[i] = (uint8_t *)malloc(sz);
      if (!g.upper[i]) ANNB_FAIL(ANNB_ENOMEM, "Not enough memory: loadIndex failed to allocate linklist");
      memcpy(g.upper[i], link_lists + off, sz);
      off += sz;
    }
  }
  g.count = (int64_t)n;
  g.num_deleted = 0;
  int vrc = g.validate();
  if (vrc == ANNB_OK) vrc = check_code_rows(h, g.level0 + g.offset_data, (int64_t)n, g.size_per_elem);
  if (int rc = vrc) {  // leave an empty, initialised graph behind
    g.count = 0;
    for (size_t i = 0; i < n; i++) {
      free(g.upper[i]);
      g.upper[i] = nullptr;
      g.levels[i] = 0;
    }
    g.maxlevel = -1;
    g.enterpoint = 0xFFFFFFFFu;
    h->dev_dirty = true;
    return rc;
  }
  g.label_lookup.reserve(n);
  for (size_t i = 0; i < n; i++) {
    g.label_lookup[g.label((uint32_t)i)] = (uint32_t)i;
    if (g.deleted((uint32_t)i)) g.num_deleted++;
  }
  h->dev_dirty = true;
  return ANNB_OK;
}

int annb_graph_info(annb_index_t *h, int64_t *cur_element_count, int64_t *max_elements, uint64_t *size_data_per_element,
                    uint64_t *link_lists_bytes, int32_t *max_level, uint32_t *enterpoint_node, int *max_M, int *max_M0, int *M,
                    int *ef_construction, double *mult) {
  ANNB_ENTER(h);
  HostGraph &g = h->g;
  if (!g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  const int64_t n = g.count.load();
  uint64_t lb = 0;
  for (int64_t i = 0; i < n; i++)
    if (g.levels[i] > 0) lb += g.size_links_per_elem * (uint64_t)g.levels[i];
  if (cur_element_count) *cur_element_count = n;
  if (max_elements) *max_elements = g.max_elements;
  if (size_data_per_element) *size_data_per_element = g.size_per_elem;
  if (link_lists_bytes) *link_lists_bytes = lb;
  if (max_level) *max_level = g.maxlevel;
  if (enterpoint_node) *enterpoint_node = g.enterpoint;
  if (max_M) *max_M = g.maxM;
  if (max_M0) *max_M0 = g.maxM0;
  if (M) *M = g.M;
  if (ef_construction) *ef_construction = g.ef_construction;
  if (mult) *mult = g.mult;
  return ANNB_OK;
}

int annb_get_graph(annb_index_t *h, uint8_t *data_level0, uint8_t *link_lists, int32_t *element_levels) {
  ANNB_ENTER(h);
  HostGraph &g = h->g;
  if (!g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  const size_t n = (size_t)g.count.load();
  if (data_level0 && n) memcpy(data_level0, g.level0, n * g.size_per_elem);
  size_t off = 0;
  for (size_t i = 0; i < n; i++) {
    if (element_levels) element_levels[i] = g.levels[i];
    if (g.levels[i] > 0) {
      const size_t sz = g.size_links_per_elem * (size_t)g.levels[i];
      if (link_lists) memcpy(link_lists + off, g.upper[i], sz);
      off += sz;
    }
  }
  return ANNB_OK;
}

int annb_encode(annb_index_t *h, const float *vectors, int v_space, int64_t n, void *codes, int c_space) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_TRY(settle_flagged_lanes(h));
  if (n < 0 || (n > 0 && (!vectors || !codes))) ANNB_FAIL(ANNB_EINVAL, "null vectors/codes");
  if (!h->d_codebook) ANNB_FAIL(ANNB_ESTATE, "Please train the PQ before using HNSW quantization backend");
  if (n == 0) return ANNB_OK;
  const float *dx;
  ANNB_TRY(stage_in(h, vectors, v_space, (size_t)n * h->dim * sizeof(float), S_QUERIES, (const void **)&dx));
  void *dc = codes;
  const size_t cbytes = (size_t)n * h->M * h->code_bytes;
  if (c_space != ANNB_DEVICE) ANNB_TRY(annb_scratch(h, S_CODES, cbytes, &dc));
  ANNB_TRY(launch_encode(h, dx, n, dc));
  if (c_space != ANNB_DEVICE) {
    ANNB_CUDA(cudaMemcpyAsync(codes, dc, cbytes, cudaMemcpyDeviceToHost, h->stream));
    ANNB_CUDA(cudaStreamSynchronize(h->stream));
  }
  return ANNB_OK;
}

// add_items: tables for each chunk of rows come from K1, double-buffered through pinned memory
struct TableFeed {
  annb_index *h;
  const float *vectors;  // host
  int64_t n, chunk_rows;
  float *pinned[2];
  cudaEvent_t done[2];
  int64_t have_first[2];
  int rc;
};
static int feed_launch(TableFeed *f, int buf, int64_t first) {
  annb_index *h = f->h;
  const int64_t cnt = std::min(f->chunk_rows, f->n - first);
  float *dq, *dt;
  int rc;
  const int qslot = buf ? S_MISC : S_QUERIES, tslot = buf ? S_OUT_D : S_TABLES;
  if ((rc = annb_scratch(h, qslot, (size_t)f->chunk_rows * h->dim * sizeof(float), (void **)&dq))) return rc;
  if ((rc = annb_scratch(h, tslot, (size_t)f->chunk_rows * h->M * h->Ks * sizeof(float), (void **)&dt))) return rc;
  ANNB_CUDA(cudaMemcpyAsync(dq, f->vectors + (size_t)first * h->dim, (size_t)cnt * h->dim * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  // get_dist_mat re-normalises for COSINE (pq.py:309-310) on top of pre_process's pass
  if (h->metric == ANNB_METRIC_COSINE && (rc = launch_l2_normalize(h, dq, cnt, h->dim))) return rc;
  if ((rc = launch_adc_table(h, dq, cnt, dt))) return rc;
  ANNB_CUDA(cudaMemcpyAsync(f->pinned[buf], dt, (size_t)cnt * h->M * h->Ks * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  ANNB_CUDA(cudaEventRecord(f->done[buf], h->stream));
  f->have_first[buf] = first;
  return ANNB_OK;
}
static const float *feed_next(void *ctx, int64_t first, int64_t cnt) {
  TableFeed *f = (TableFeed *)ctx;
  (void)cnt;
  const int buf = (int)((first / f->chunk_rows) & 1);
  if (f->have_first[buf] != first && (f->rc = feed_launch(f, buf, first))) return nullptr;
  if (cudaEventSynchronize(f->done[buf]) != cudaSuccess) {
    annb_set_error("CUDA error while building insertion tables");
    f->rc = ANNB_ECUDA;
    return nullptr;
  }
  const int64_t nxt = first + f->chunk_rows;
  if (nxt < f->n && (f->rc = feed_launch(f, buf ^ 1, nxt))) return nullptr;
  return f->pinned[buf];
}

int annb_add_items(annb_index_t *h, const float *vectors, const void *codes, const uint64_t *labels, int64_t n, int num_threads) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_TRY(settle_flagged_lanes(h));
  if (n < 0 || (n > 0 && (!vectors || !labels))) ANNB_FAIL(ANNB_EINVAL, "null vectors/labels");
  if (!h->d_codebook) ANNB_FAIL(ANNB_ESTATE, "Please train the PQ before using HNSW quantization backend");
  if (!h->g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised: call annb_init_graph first");
  if (n == 0) return ANNB_OK;
  ANNB_CUDA(cudaStreamSynchronize(h->stream));   // insertion reuses search scratch: let streamed searches finish
  ANNB_CUDA(cudaStreamSynchronize(h->stream2));
  // num_threads == 1 is the reference's deterministic single-threaded build; any other value is an unordered
  // concurrent build in the reference too -- large batches of fresh rows then take the GPU builder
  if (num_threads != 1 && !codes && h->opt_gpu_build) {
    const int rc = gpu_build_run(h, vectors, labels, n, num_threads);
    if (rc != 1) return rc;
  }
  return hnsw_host_add(h, vectors, codes, labels, n, num_threads, nullptr);
}

}  // extern "C"

int hnsw_host_add(annb_index *h, const float *vectors, const void *codes, const uint64_t *labels, int64_t n, int num_threads,
                  const int32_t *forced_levels) {
  if (n == 0) return ANNB_OK;
  const size_t crow = (size_t)h->M * h->code_bytes;
  std::vector<uint8_t> own_codes;
  const uint8_t *hc = (const uint8_t *)codes;
  if (hc) ANNB_TRY(check_code_rows(h, hc, n, crow));
  if (!hc) {  // PQCodec.encode on the GPU, chunked
    own_codes.resize((size_t)n * crow);
    const int64_t step = 1 << 18;
    for (int64_t s = 0; s < n; s += step) {
      const int64_t c = std::min(step, n - s);
      float *dx;
      void *dc;
      ANNB_TRY(annb_scratch(h, S_QUERIES, (size_t)c * h->dim * sizeof(float), (void **)&dx));
      ANNB_TRY(annb_scratch(h, S_CODES, (size_t)c * crow, &dc));
      ANNB_CUDA(cudaMemcpyAsync(dx, vectors + (size_t)s * h->dim, (size_t)c * h->dim * sizeof(float), cudaMemcpyHostToDevice, h->stream));
      ANNB_TRY(launch_encode(h, dx, c, dc));
      ANNB_CUDA(cudaMemcpyAsync(own_codes.data() + (size_t)s * crow, dc, (size_t)c * crow, cudaMemcpyDeviceToHost, h->stream));
      ANNB_CUDA(cudaStreamSynchronize(h->stream));
    }
    hc = own_codes.data();
  }
  TableFeed f;
  f.h = h;
  f.vectors = vectors;
  f.n = n;
  const size_t TS = (size_t)h->M * h->Ks;
  f.chunk_rows = std::max<int64_t>(64, std::min<int64_t>(8192, (int64_t)((64u << 20) / (TS * sizeof(float)))));
  f.rc = 0;
  for (int b = 0; b < 2; b++) {
    ANNB_TRY(annb_pinned(h, b, (size_t)f.chunk_rows * TS * sizeof(float), (void **)&f.pinned[b]));
    ANNB_CUDA(cudaEventCreateWithFlags(&f.done[b], cudaEventDisableTiming));
    f.have_first[b] = -1;
  }
  // device copy current (possibly with earlier patches pending)?  then remember what this insertion rewrites
  const bool track = !h->dev_dirty && h->d_rec0 && h->gd.n > 0 && !forced_levels;
  BuildTrack tr;
  int rc = hnsw_insert_rows(h, hc, labels, n, num_threads, feed_next, &f, f.chunk_rows, forced_levels, track ? &tr : nullptr);
  cudaStreamSynchronize(h->stream);
  for (int b = 0; b < 2; b++) cudaEventDestroy(f.done[b]);
  if (track && rc == ANNB_OK && !tr.untracked) {
    h->patch.dirty0.insert(h->patch.dirty0.end(), tr.dirty0.begin(), tr.dirty0.end());
    h->patch.upper_dirty |= tr.upper_dirty;
    h->patch_pending = true;
  } else {
    h->dev_dirty = true;
  }
  if (rc == ANNB_ECUDA && f.rc) return f.rc;
  return rc;
}


extern "C" {

struct HostTables {
  const float *tables;
  size_t TS;
};
static const float *host_tables_next(void *ctx, int64_t first, int64_t) {
  HostTables *t = (HostTables *)ctx;
  return t->tables + (size_t)first * t->TS;
}
int annb_add_items_with_tables(annb_index_t *h, const void *codes, const float *tables, const uint64_t *labels, int64_t n,
                               int num_threads) {
  ANNB_ENTER(h);
  if (n < 0 || (n > 0 && (!codes || !tables || !labels))) ANNB_FAIL(ANNB_EINVAL, "null codes/tables/labels");
  if (!h->g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised: call annb_init_graph first");
  if (n == 0) return ANNB_OK;
  ANNB_TRY(check_code_rows(h, (const uint8_t *)codes, n, (size_t)h->M * h->code_bytes));
  HostTables t{tables, (size_t)h->M * h->Ks};
  int rc = hnsw_insert_rows(h, (const uint8_t *)codes, labels, n, num_threads, host_tables_next, &t, n);
  h->dev_dirty = true;
  return rc;
}

int annb_resize_index(annb_index_t *h, int64_t new_max_elements) {
  ANNB_ENTER(h);
  if (!h->g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  return h->g.resize(new_max_elements);
}

static int set_deleted(annb_index *h, uint64_t label, bool del) {
  HostGraph &g = h->g;
  auto it = g.label_lookup.find(label);
  if (it == g.label_lookup.end()) ANNB_FAIL(ANNB_ENOTFOUND, "Label not found");
  uint8_t *flags = g.rec0(it->second) + 2;
  if (del) {
    if (*flags & 1) ANNB_FAIL(ANNB_EINVAL, "The requested to delete element is already deleted");
    *flags |= 1;
    g.num_deleted++;
  } else {
    if (!(*flags & 1)) ANNB_FAIL(ANNB_EINVAL, "The requested to undelete element is not deleted");
    *flags &= (uint8_t)~1;
    g.num_deleted--;
  }
  h->deleted_dirty = true;
  return ANNB_OK;
}
int annb_mark_deleted(annb_index_t *h, uint64_t label) {
  ANNB_ENTER(h);
  return set_deleted(h, label, true);
}
int annb_unmark_deleted(annb_index_t *h, uint64_t label) {
  ANNB_ENTER(h);
  return set_deleted(h, label, false);
}

int annb_element_count(annb_index_t *h, int64_t *out) {
  if (!h || !out) ANNB_FAIL(ANNB_EINVAL, "null argument");
  *out = h->g.inited ? h->g.count.load() : 0;
  return ANNB_OK;
}

int annb_get_labels(annb_index_t *h, uint64_t *labels_out, int64_t cap) {
  ANNB_ENTER(h);
  const int64_t n = std::min<int64_t>(cap, h->g.inited ? h->g.count.load() : 0);
  for (int64_t i = 0; i < n; i++) labels_out[i] = h->g.label((uint32_t)i);
  return ANNB_OK;
}

int annb_get_codes(annb_index_t *h, const uint64_t *labels, int64_t n, void *codes_out) {
  ANNB_ENTER(h);
  HostGraph &g = h->g;
  for (int64_t i = 0; i < n; i++) {
    auto it = g.label_lookup.find(labels[i]);
    if (it == g.label_lookup.end() || g.deleted(it->second)) ANNB_FAIL(ANNB_ENOTFOUND, "Label not found");
    memcpy((uint8_t *)codes_out + (size_t)i * g.code_row_bytes, g.code(it->second), g.code_row_bytes);
  }
  return ANNB_OK;
}

}  // extern "C"

// ---- host graph -> device walk layout ---------------------------------------------------------
static int ensure_dev(void **p, size_t *cap, size_t bytes) {
  if (*cap >= bytes && *p) return ANNB_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes + 256;
    e = cudaMalloc(p, want);
  }
  if (e != cudaSuccess) ANNB_FAIL(ANNB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
  *cap = want;
  return ANNB_OK;
}

static int upload_deleted(annb_index *h) {
  HostGraph &g = h->g;
  const int64_t n = g.count.load();
  const size_t words = (size_t)(n + 31) / 32 + 1;
  std::vector<uint32_t> bm(words, 0);
  for (int64_t i = 0; i < n; i++)
    if (g.deleted((uint32_t)i)) bm[i >> 5] |= 1u << (i & 31);
  ANNB_TRY(ensure_dev((void **)&h->d_deleted, &h->cap_deleted, words * 4));
  ANNB_CUDA(cudaMemcpyAsync(h->d_deleted, bm.data(), words * 4, cudaMemcpyHostToDevice, h->stream));
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  h->gd.deleted = h->d_deleted;
  h->deleted_dirty = false;
  return ANNB_OK;
}

// upper levels of the walk layout + entry point, from the host graph (all of it: ~N/16 small records)
static int upload_upper(annb_index *h) {
  HostGraph &g = h->g;
  GraphDev &d = h->gd;
  const int64_t n = g.count.load();
  d.maxlevel = g.maxlevel;
  d.ep_node = g.enterpoint;
  if (g.maxlevel >= ANNB_MAX_LEVELS) ANNB_FAIL(ANNB_ELIMIT, "graph has %d levels (limit %d)", g.maxlevel + 1, ANNB_MAX_LEVELS);
  std::vector<uint8_t> up;
  std::vector<uint32_t> map_prev, map_cur;
  size_t off = 0;
  for (int l = 1; l <= g.maxlevel; l++) {
    std::vector<uint32_t> nodes;
    for (int64_t i = 0; i < n; i++)
      if (g.levels[i] >= l) nodes.push_back((uint32_t)i);
    map_cur.assign((size_t)n, 0xffffffffu);
    for (size_t r = 0; r < nodes.size(); r++) map_cur[nodes[r]] = (uint32_t)r;
    d.up_off[l] = off;
    up.resize(off + nodes.size() * (size_t)d.recu_bytes, 0);
    for (size_t r = 0; r < nodes.size(); r++) {
      uint8_t *rec = up.data() + off + r * (size_t)d.recu_bytes;
      const uint32_t u = nodes[r];
      const uint8_t *ll = g.list_at(u, l);
      uint16_t cnt;
      memcpy(&cnt, ll, 2);
      uint32_t *links = reinterpret_cast<uint32_t *>(rec);
      for (int j = 0; j < g.maxM; j++) {
        uint32_t lk = 0xffffffffu;
        if (j < (int)cnt) {
          uint32_t v;
          memcpy(&v, ll + 4 + 4 * j, 4);
          if (map_cur[v] == 0xffffffffu) ANNB_FAIL(ANNB_EINVAL, "Trying to make a link on a non-existent level");
          lk = map_cur[v];
          memcpy(rec + d.code_offu + (size_t)j * d.code_row, g.code(v), (size_t)d.code_row);
        }
        links[j] = lk;
      }
      uint32_t tail[2] = {u, l == 1 ? u : map_prev[u]};
      memcpy(rec + d.tail_offu, tail, 8);
    }
    off += nodes.size() * (size_t)d.recu_bytes;
    if (l == g.maxlevel) d.ep_rec = map_cur[g.enterpoint];
    map_prev.swap(map_cur);
  }
  if (g.maxlevel <= 0) d.ep_rec = g.enterpoint;
  ANNB_TRY(ensure_dev((void **)&h->d_up, &h->cap_up, std::max<size_t>(up.size(), 16)));
  if (!up.empty()) ANNB_CUDA(cudaMemcpyAsync(h->d_up, up.data(), up.size(), cudaMemcpyHostToDevice, h->stream));
  d.up = h->d_up;
  memcpy(d.ep_code, g.code(g.enterpoint), (size_t)d.code_row);
  return ANNB_OK;
}

// A small host insertion into a graph whose device copy was current: upload only what changed -- the rewritten
// level-0 records (packed on the host: links + the neighbours' codes), the new labels, and the upper levels if any
// of them (or the entry point) moved -- instead of re-deriving all N records.
static int patch_device_graph(annb_index *h) {
  HostGraph &g = h->g;
  GraphDev &d = h->gd;
  const int64_t n = g.count.load(), n_old = d.n;
  std::vector<uint32_t> &ids = h->patch.dirty0;
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  const size_t cnt = ids.size();
  if (h->patch.untracked || n < n_old || cnt > (size_t)n / 4 + 64 || (size_t)n * d.rec0_bytes > h->cap_rec0 || (size_t)n * 8 > h->cap_labels) {
    h->dev_dirty = true;   // too much changed, or the device buffers are too small: re-derive everything
    return 1;
  }
  uint8_t *hst;
  uint32_t *hid;
  ANNB_TRY(annb_pinned(h, 0, cnt * (size_t)d.rec0_bytes + 16, (void **)&hst));
  ANNB_TRY(annb_pinned(h, 1, cnt * 4 + 16, (void **)&hid));
  for (size_t t = 0; t < cnt; t++) {
    const uint32_t u = ids[t];
    hid[t] = u;
    uint8_t *rec = hst + t * (size_t)d.rec0_bytes;
    memset(rec, 0, (size_t)d.rec0_bytes);
    const uint8_t *ll = g.rec0(u);
    uint16_t c;
    memcpy(&c, ll, 2);
    uint32_t *links = reinterpret_cast<uint32_t *>(rec);
    for (int j = 0; j < g.maxM0; j++) {
      uint32_t lk = 0xffffffffu;
      if (j < (int)c) {
        memcpy(&lk, ll + 4 + 4 * j, 4);
        memcpy(rec + d.code_off0 + (size_t)j * d.code_row, g.code(lk), (size_t)d.code_row);
      }
      links[j] = lk;
    }
  }
  uint8_t *dst;
  uint32_t *did;
  ANNB_TRY(annb_scratch(h, S_RAW0, cnt * (size_t)d.rec0_bytes + 16, (void **)&dst));
  ANNB_TRY(annb_scratch(h, S_QMAP, cnt * 4 + 16, (void **)&did));
  ANNB_CUDA(cudaMemcpyAsync(dst, hst, cnt * (size_t)d.rec0_bytes, cudaMemcpyHostToDevice, h->stream));
  ANNB_CUDA(cudaMemcpyAsync(did, hid, cnt * 4, cudaMemcpyHostToDevice, h->stream));
  ANNB_TRY(launch_scatter_records(h, dst, did, (int64_t)cnt, d.rec0_bytes, h->d_rec0));
  if (n > n_old) {
    std::vector<uint64_t> labels((size_t)(n - n_old));
    for (int64_t i = n_old; i < n; i++) {
      labels[i - n_old] = g.label((uint32_t)i);
      h->max_label = std::max(h->max_label, labels[i - n_old]);
      h->labels_identity &= labels[i - n_old] == (uint64_t)i;
    }
    ANNB_CUDA(cudaMemcpyAsync(h->d_labels + n_old, labels.data(), labels.size() * 8, cudaMemcpyHostToDevice, h->stream));
    ANNB_CUDA(cudaStreamSynchronize(h->stream));  // `labels` goes out of scope
  }
  d.n = n;
  if (h->patch.upper_dirty) ANNB_TRY(upload_upper(h));
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  ANNB_TRY(upload_deleted(h));
  h->patch_pending = false;
  h->patch = BuildTrack();
  h->patches++;
  return ANNB_OK;
}

int sync_device_graph(annb_index *h) {
  HostGraph &g = h->g;
  if (!g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  if (!h->dev_dirty && !h->deleted_dirty && !h->patch_pending) return ANNB_OK;
  // device buffers are about to be rewritten / reallocated: nothing may still be walking them
  // (streamed searches of annb_search_submit run on both lanes)
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->stream2) ANNB_CUDA(cudaStreamSynchronize(h->stream2));
  if (!h->dev_dirty && h->patch_pending) {
    const int rc = patch_device_graph(h);
    if (rc != 1) return rc;   // 1 = not patchable: fall through to the full re-derivation
  }
  if (!h->dev_dirty) return upload_deleted(h);
  const int64_t n = g.count.load();
  GraphDev &d = h->gd;
  memset(&d, 0, sizeof(d));
  d.n = n;
  d.M = h->M;
  d.Ks = h->Ks;
  d.code_bytes = h->code_bytes;
  d.code_row = h->M * h->code_bytes;
  d.maxM = g.maxM;
  d.maxM0 = g.maxM0;
  auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
  d.code_off0 = (int)up16((size_t)4 * g.maxM0);
  d.rec0_bytes = (int)up16((size_t)d.code_off0 + (size_t)g.maxM0 * d.code_row);
  d.code_offu = (int)up16((size_t)4 * g.maxM);
  d.tail_offu = (int)(((size_t)d.code_offu + (size_t)g.maxM * d.code_row + 7) / 8 * 8);
  d.recu_bytes = (int)up16((size_t)d.tail_offu + 8);
  d.maxlevel = g.maxlevel;
  d.ep_node = g.enterpoint;
  if (n == 0) {
    h->dev_dirty = false;
    return ANNB_OK;
  }
  if (g.maxlevel >= ANNB_MAX_LEVELS) ANNB_FAIL(ANNB_ELIMIT, "graph has %d levels (limit %d)", g.maxlevel + 1, ANNB_MAX_LEVELS);

  // level 0: upload the raw records, then gather [links | neighbour codes] on the device
  uint8_t *raw;
  const size_t raw_bytes = (size_t)n * g.size_per_elem;
  ANNB_TRY(annb_scratch(h, S_RAW0, raw_bytes, (void **)&raw));
  ANNB_CUDA(cudaMemcpyAsync(raw, g.level0, raw_bytes, cudaMemcpyHostToDevice, h->stream));
  const int64_t n_alloc = std::max<int64_t>(n, h->reserve_nodes);  // the GPU builder appends nodes in place
  ANNB_TRY(ensure_dev((void **)&h->d_rec0, &h->cap_rec0, (size_t)n_alloc * d.rec0_bytes));
  d.rec0 = h->d_rec0;
  ANNB_TRY(launch_pack_rec0(h, raw, n));

  // labels
  std::vector<uint64_t> labels((size_t)n);
  uint64_t maxl = 0;
  bool ident = true;
  for (int64_t i = 0; i < n; i++) {
    labels[i] = g.label((uint32_t)i);
    maxl = std::max(maxl, labels[i]);
    ident &= labels[i] == (uint64_t)i;
  }
  h->max_label = maxl;
  h->labels_identity = ident;
  ANNB_TRY(ensure_dev((void **)&h->d_labels, &h->cap_labels, (size_t)n_alloc * 8));
  ANNB_CUDA(cudaMemcpyAsync(h->d_labels, labels.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  d.labels = h->d_labels;

  ANNB_TRY(upload_upper(h));
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  ANNB_TRY(upload_deleted(h));
  h->dev_dirty = false;
  h->patch_pending = false;
  h->patch = BuildTrack();
  h->full_syncs++;
  return ANNB_OK;
}

extern "C" {

// ---- K3 ---------------------------------------------------------------------------------------
int annb_search(annb_index_t *h, const float *queries, const float *tables, int in_space, int64_t B, int normalize, int k,
                int ef, const uint64_t *filter_labels, int filter_space, int64_t n_filter, uint64_t *labels_out,
                float *dists_out, int out_space, int64_t *stats_out) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_TRY(settle_flagged_lanes(h));
  if ((queries == nullptr) == (tables == nullptr)) ANNB_FAIL(ANNB_EINVAL, "exactly one of queries / tables must be given");
  if (B < 0 || k <= 0 || ef <= 0 || !labels_out || !dists_out) ANNB_FAIL(ANNB_EINVAL, "bad B/k/ef/outputs");
  if (!h->g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  ANNB_TRY(sync_device_graph(h));
  if (B == 0) return ANNB_OK;
  const int ef_eff = std::max(ef, k);  // hnswalg.h:1279
  if (ef_eff > ANNB_MAX_EF) ANNB_FAIL(ANNB_ELIMIT, "max(ef, k)=%d exceeds ANNB_MAX_EF=%d", ef_eff, ANNB_MAX_EF);
  const bool host_out = out_space != ANNB_DEVICE;
  uint64_t *dl = labels_out;
  float *dd = dists_out;
  int64_t *dstats = nullptr;
  int32_t *dfound;
  if (host_out) {
    ANNB_TRY(annb_scratch(h, S_OUT_L, (size_t)B * k * 8, (void **)&dl));
    ANNB_TRY(annb_scratch(h, S_OUT_D, (size_t)B * k * 4, (void **)&dd));
  }
  ANNB_TRY(annb_scratch(h, S_FOUND, (size_t)B * 4, (void **)&dfound));
  if (stats_out) {
    if (host_out) ANNB_TRY(annb_scratch(h, S_STATS, (size_t)B * 24, (void **)&dstats));
    else dstats = stats_out;
  }
  if (h->gd.n == 0) {  // empty index: searchKnn returns nothing (hnswalg.h:1240)
    ANNB_FAIL(ANNB_EFEWRESULTS, "Cannot return the results in a contigious 2D array. Probably ef or M is too small");
  }
  // ---- host-buffer fast path: chunked two-stream pipeline -------------------------------------------
  // H2D of chunk c+1 overlaps the walk of chunk c, the walk of chunk c+1 fills the SMs that chunk c's
  // persistent launch leaves idle in its tail, and D2H of chunk c overlaps the walk of chunk c+1.
  {
    const bool plain = !filter_labels && h->g.num_deleted == 0 && !h->opt_force_general;
    int nch = (int)h->opt_chunks;
    if (nch == 0) nch = (B >= 32768) ? 4 : (B >= 4096 ? 2 : 1);
    if (nch > 8) nch = 8;
    if (queries && in_space != ANNB_DEVICE && host_out && plain && nch > 1 && B >= 2 * nch) {
      const size_t TS = (size_t)h->M * h->Ks;
      const bool fuse = walk4_can_fuse(h);  // the walk builds its own tables: no (B,M,Ks) buffer at all
      float *dq, *dtab = nullptr;
      unsigned int *counters;
      int32_t *hfound;
      ANNB_TRY(annb_scratch(h, S_QUERIES, (size_t)B * h->dim * sizeof(float), (void **)&dq));
      if (!fuse) ANNB_TRY(annb_scratch(h, S_TABLES, (size_t)B * TS * sizeof(float), (void **)&dtab));
      ANNB_TRY(annb_scratch(h, S_COUNTER, 256 * 9, (void **)&counters));
      ANNB_TRY(annb_pinned(h, 2, (size_t)B * 4, (void **)&hfound));
      if (!h->d_codebook) ANNB_FAIL(ANNB_ESTATE, "Please train the PQ before using HNSW quantization backend");
      ANNB_CUDA(cudaStreamSynchronize(h->stream));  // scratch (re)allocation and earlier work are settled
      cudaStream_t lanes[2] = {h->stream, h->stream2};
      cudaStream_t saved = h->stream;
      int rc = ANNB_OK;
      for (int c = 0; c < nch && rc == ANNB_OK; c++) {
        const int64_t b0 = B * c / nch, b1 = B * (c + 1) / nch, nb = b1 - b0;
        h->stream = lanes[c & 1];
        if (cudaMemcpyAsync(dq + b0 * h->dim, queries + b0 * h->dim, (size_t)nb * h->dim * sizeof(float),
                            cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
          rc = ANNB_ECUDA;
          break;
        }
        for (int r = 0; r < normalize && rc == ANNB_OK; r++) rc = launch_l2_normalize(h, dq + b0 * h->dim, nb, h->dim);
        if (rc == ANNB_OK && !fuse) rc = launch_adc_table(h, dq + b0 * h->dim, nb, dtab + b0 * TS);
        if (rc != ANNB_OK) break;
        SearchParams p;
        memset(&p, 0, sizeof(p));
        if (fuse) fuse_params(h, p, dq + b0 * h->dim);
        else p.tables = dtab + b0 * TS;
        if (h->opt_dump_tables) p.dump_tables = reinterpret_cast<float *>(h->opt_dump_tables) + b0 * TS;
        p.B = nb;
        p.k = k;
        p.ef = ef_eff;
        p.out_labels = dl + b0 * k;
        p.out_dists = dd + b0 * k;
        p.out_found = dfound + b0;
        p.out_stats = dstats ? dstats + b0 * 3 : nullptr;
        p.work_counter = counters + 64 * (c + 1);
        if (c == 0 && h->opt_timing) cudaEventRecord(h->ev[2], h->stream);
        rc = launch_search(h, p, false);
        if (rc != ANNB_OK) break;
        if (c == nch - 1 && h->opt_timing) cudaEventRecord(h->ev[3], h->stream);
        cudaMemcpyAsync(hfound + b0, dfound + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(labels_out + b0 * k, dl + b0 * k, (size_t)nb * k * 8, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(dists_out + b0 * k, dd + b0 * k, (size_t)nb * k * 4, cudaMemcpyDeviceToHost, h->stream);
        if (stats_out) cudaMemcpyAsync(stats_out + b0 * 3, dstats + b0 * 3, (size_t)nb * 24, cudaMemcpyDeviceToHost, h->stream);
      }
      h->stream = saved;
      cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->stream2);
      if (rc != ANNB_OK) {
        if (rc == ANNB_ECUDA) annb_set_error("CUDA error in the chunked search pipeline: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
      }
      if (e1 != cudaSuccess || e2 != cudaSuccess)
        ANNB_FAIL(ANNB_ECUDA, "CUDA error %s in the chunked search pipeline", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      for (int64_t b = 0; b < B; b++)
        if (hfound[b] < k)
          ANNB_FAIL(ANNB_EFEWRESULTS, "Cannot return the results in a contigious 2D array. Probably ef or M is too small");
      return ANNB_OK;
    }
  }
  // tables: handed in (the literal dtables argument), built inside the walk (plain search, hnsw_walk4), or by K1
  const size_t tbytes = (size_t)B * h->M * h->Ks * sizeof(float);
  const float *dt = nullptr, *dq_fused = nullptr;
  const bool plain_search = !filter_labels && h->g.num_deleted == 0 && !h->opt_force_general;
  // fraction of the nodes a filtered / deletion-aware walk can admit: sizes its lists
  const double selectivity = filter_labels ? std::min<double>(1.0, (double)n_filter / (double)std::max<int64_t>(1, h->gd.n))
                                           : 1.0 - (double)h->g.num_deleted / (double)std::max<int64_t>(1, h->gd.n);
  // the table is built inside the walk for the plain search (hnsw_walk4) and for the filtered one (hnsw_walk4f)
  const bool fuse_walk = walk4_can_fuse(h) && (plain_search || (h->opt_force_general != 2 && walk4f_applicable(h, ef_eff, selectivity)));
  if (tables) {
    ANNB_TRY(stage_in(h, tables, in_space, tbytes, S_TABLES, (const void **)&dt));
  } else if (fuse_walk) {
    ANNB_TRY(stage_queries(h, queries, in_space, B, normalize, S_QUERIES, &dq_fused));
  } else {
    float *t;
    ANNB_TRY(annb_scratch(h, S_TABLES, tbytes, (void **)&t));
    ANNB_TRY(build_tables(h, queries, in_space, B, normalize, t));
    dt = t;
  }
  // filter
  const uint32_t *dfilter = nullptr;
  if (filter_labels) {
    uint32_t *by_id;
    ANNB_TRY(annb_scratch(h, S_FLT_BY_ID, ((size_t)(h->gd.n + 31) / 32 + 1) * 4, (void **)&by_id));
    if (h->max_label > (uint64_t)h->gd.n * 64 + (1ull << 30)) {
      // labels too sparse for a by-label bitmap on the device: resolve them through the host's
      // label -> internal id map (the same map mark_deleted uses) and upload the by-id bitmap
      std::vector<uint64_t> tmp;
      const uint64_t *hl = filter_labels;
      if (filter_space == ANNB_DEVICE) {
        tmp.resize((size_t)n_filter);
        ANNB_CUDA(cudaMemcpyAsync(tmp.data(), filter_labels, (size_t)n_filter * 8, cudaMemcpyDeviceToHost, h->stream));
        ANNB_CUDA(cudaStreamSynchronize(h->stream));
        hl = tmp.data();
      }
      std::vector<uint32_t> bm((size_t)(h->gd.n + 31) / 32 + 1, 0u);
      for (int64_t i = 0; i < n_filter; i++) {
        auto it = h->g.label_lookup.find(hl[i]);
        if (it != h->g.label_lookup.end()) bm[it->second >> 5] |= 1u << (it->second & 31);
      }
      ANNB_CUDA(cudaMemcpyAsync(by_id, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice, h->stream));
      ANNB_CUDA(cudaStreamSynchronize(h->stream));
    } else {
      const uint64_t *dfl;
      ANNB_TRY(stage_in(h, filter_labels, filter_space, (size_t)std::max<int64_t>(n_filter, 1) * 8, S_FLT_LABELS, (const void **)&dfl));
      uint32_t *by_label;
      ANNB_TRY(annb_scratch(h, S_FLT_BY_LABEL, ((size_t)(h->max_label >> 5) + 1) * 4, (void **)&by_label));
      ANNB_TRY(launch_filter_bitmap(h, dfl, n_filter, by_label, by_id));
    }
    dfilter = by_id;
  }
  SearchParams p;
  memset(&p, 0, sizeof(p));
  if (dq_fused) fuse_params(h, p, dq_fused);
  else p.tables = dt;
  if (h->opt_dump_tables) p.dump_tables = reinterpret_cast<float *>(h->opt_dump_tables);
  p.B = B;
  p.k = k;
  p.ef = ef_eff;
  p.filter = dfilter;
  p.out_labels = dl;
  p.out_dists = dd;
  p.out_found = dfound;
  p.out_stats = dstats;
  const bool general = dfilter != nullptr || h->g.num_deleted > 0 || h->opt_force_general;
  p.selectivity = (float)selectivity;
  int32_t *hfound;
  ANNB_TRY(annb_pinned(h, 2, (size_t)B * 4, (void **)&hfound));
  int mode = general ? (h->opt_force_general == 2 ? 2 : 1) : 0;
  for (;;) {
    if (h->opt_timing) cudaEventRecord(h->ev[2], h->stream);
    ANNB_TRY(launch_search(h, p, mode));
    if (h->opt_timing) cudaEventRecord(h->ev[3], h->stream);
    // found < k anywhere?  (hnsw_bindings.cpp:342-345).  A tiny reduction on the host side of `found`.
    ANNB_CUDA(cudaMemcpyAsync(hfound, dfound, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    if (host_out) {
      ANNB_CUDA(cudaMemcpyAsync(labels_out, dl, (size_t)B * k * 8, cudaMemcpyDeviceToHost, h->stream));
      ANNB_CUDA(cudaMemcpyAsync(dists_out, dd, (size_t)B * k * 4, cudaMemcpyDeviceToHost, h->stream));
      if (stats_out) ANNB_CUDA(cudaMemcpyAsync(stats_out, dstats, (size_t)B * 24, cudaMemcpyDeviceToHost, h->stream));
    }
    ANNB_CUDA(cudaStreamSynchronize(h->stream));
    if (mode == 2) break;
    // queries that outgrew the flagged walk's list (found = -1): only those are redone on the bitmap walk
    // (exact, any size) through a query-index map; their rows of the outputs are simply overwritten
    int64_t n_over = 0;
    for (int64_t b = 0; b < B; b++) n_over += hfound[b] < 0;
    if (n_over == 0) break;
    uint32_t *hmap, *dmap;
    ANNB_TRY(annb_pinned(h, 3, (size_t)n_over * 4, (void **)&hmap));
    ANNB_TRY(annb_scratch(h, S_QMAP, (size_t)n_over * 4, (void **)&dmap));
    int64_t w = 0;
    for (int64_t b = 0; b < B; b++)
      if (hfound[b] < 0) hmap[w++] = (uint32_t)b;
    ANNB_CUDA(cudaMemcpyAsync(dmap, hmap, (size_t)n_over * 4, cudaMemcpyHostToDevice, h->stream));
    if (!p.tables) {  // the walk built its tables in shared memory: the bitmap walk reads them from HBM (K1, all rows)
      float *t;
      ANNB_TRY(annb_scratch(h, S_TABLES, tbytes, (void **)&t));
      ANNB_TRY(launch_adc_table(h, p.queries, B, t));
      p.tables = t;
      p.queries = nullptr;
    }
    p.qmap = dmap;
    p.B = n_over;
    mode = 2;
    h->flagged_fallbacks++;
    h->flagged_fallback_queries += n_over;
  }
  for (int64_t b = 0; b < B; b++)
    if (hfound[b] < k)
      ANNB_FAIL(ANNB_EFEWRESULTS, "Cannot return the results in a contigious 2D array. Probably ef or M is too small");
  return ANNB_OK;
}

int annb_scan_subset(annb_index_t *h, const float *queries, int in_space, int64_t B, int normalize, int k,
                     const uint64_t *subset_labels, int64_t n_subset, uint64_t *labels_out, float *dists_out) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_TRY(settle_flagged_lanes(h));
  if (!queries || !labels_out || !dists_out || B < 0 || k <= 0 || n_subset < 0 || (n_subset > 0 && !subset_labels))
    ANNB_FAIL(ANNB_EINVAL, "bad arguments");
  if (!h->g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  if (B == 0) return ANNB_OK;
  HostGraph &g = h->g;
  const size_t crow = g.code_row_bytes;
  std::vector<uint8_t> codes;
  std::vector<uint64_t> labs;
  codes.reserve((size_t)n_subset * crow);
  labs.reserve((size_t)n_subset);
  for (int64_t i = 0; i < n_subset; i++) {
    auto it = g.label_lookup.find(subset_labels[i]);
    if (it == g.label_lookup.end() || g.deleted(it->second)) continue;
    const uint8_t *c = g.code(it->second);
    codes.insert(codes.end(), c, c + crow);
    labs.push_back(subset_labels[i]);
  }
  const int64_t n = (int64_t)labs.size();
  // the scan works on the handle's flat code matrix: keep the caller's one (annb_set_codes) intact
  uint8_t *saved_codes = h->d_codes;
  const int64_t saved_n = h->n_codes;
  h->d_codes = nullptr;
  h->n_codes = 0;
  int rc = ANNB_OK;
  std::vector<int64_t> ids((size_t)B * k);
  if (n > 0) {
    cudaError_t ce = cudaMalloc(&h->d_codes, (size_t)n * crow + 16);
    if (ce != cudaSuccess) rc = ANNB_ENOMEM;
    if (rc == ANNB_OK && cudaMemcpyAsync(h->d_codes, codes.data(), (size_t)n * crow, cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
      rc = ANNB_ECUDA;
    h->n_codes = n;
    float *t = nullptr, *dd = nullptr;
    int64_t *dids = nullptr;
    if (rc == ANNB_OK) rc = annb_scratch(h, S_TABLES, (size_t)B * h->M * h->Ks * sizeof(float), (void **)&t);
    if (rc == ANNB_OK) rc = annb_scratch(h, S_OUT_L, (size_t)B * k * sizeof(int64_t), (void **)&dids);
    if (rc == ANNB_OK) rc = annb_scratch(h, S_OUT_D, (size_t)B * k * sizeof(float), (void **)&dd);
    if (rc == ANNB_OK) rc = build_tables(h, queries, in_space, B, normalize, t);
    if (rc == ANNB_OK) rc = launch_scan_topk(h, t, B, k, dids, dd);
    if (rc == ANNB_OK) {
      cudaMemcpyAsync(ids.data(), dids, (size_t)B * k * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream);
      cudaMemcpyAsync(dists_out, dd, (size_t)B * k * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
      if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = ANNB_ECUDA;
    }
    if (h->d_codes) cudaFree(h->d_codes);
  }
  h->d_codes = saved_codes;
  h->n_codes = saved_n;
  if (rc == ANNB_ECUDA) annb_set_error("CUDA error in annb_scan_subset: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc == ANNB_ENOMEM) annb_set_error("cudaMalloc failed in annb_scan_subset");
  if (rc != ANNB_OK) return rc;
  for (int64_t i = 0; i < B * k; i++) {
    if (n > 0 && ids[i] >= 0) {
      labels_out[i] = labs[(size_t)ids[i]];
    } else {
      labels_out[i] = UINT64_MAX;
      dists_out[i] = std::numeric_limits<float>::infinity();
    }
  }
  if (n < k) ANNB_FAIL(ANNB_EFEWRESULTS, "Cannot return the results in a contigious 2D array. Probably ef or M is too small");
  return ANNB_OK;
}

// ---- streaming submit / wait ---------------------------------------------------------------------
// A filtered / deletion-aware batch whose walk flagged queries (found = -1: N outgrew its capacity) is finished at
// wait time: exactly those queries are re-run on the bitmap walk, on the lane's own stream and scratch.
static int lane_rerun_flagged(annb_index *h, int lane, int64_t n_over) {
  annb_index::AsyncLane &L = h->lanes[lane];
  cudaStream_t st = lane ? h->stream2 : h->stream;
  uint32_t *hmap, *dmap;
  float *t;
  ANNB_TRY(annb_pinned(h, 3, (size_t)n_over * 4, (void **)&hmap));
  ANNB_TRY(annb_scratch(h, S_QMAP, (size_t)n_over * 4, (void **)&dmap));
  ANNB_TRY(annb_scratch(h, lane ? S_L1_TABLES : S_TABLES, (size_t)L.B * h->M * h->Ks * sizeof(float), (void **)&t));
  int64_t w = 0;
  for (int64_t b = 0; b < L.B; b++)
    if (L.hfound[b] < 0) hmap[w++] = (uint32_t)b;
  cudaStream_t saved = h->stream;
  h->stream = st;
  int rc = ANNB_OK;
  if (cudaMemcpyAsync(dmap, hmap, (size_t)n_over * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = ANNB_ECUDA;
  if (rc == ANNB_OK) rc = launch_adc_table(h, L.dq, L.B, t);
  if (rc == ANNB_OK) {
    SearchParams p;
    memset(&p, 0, sizeof(p));
    p.tables = t;
    p.B = n_over;
    p.k = L.k;
    p.ef = L.ef;
    p.filter = L.dfilter;
    p.selectivity = L.selectivity;
    p.out_labels = L.dl;
    p.out_dists = L.dd;
    p.out_found = L.dfound;
    p.qmap = dmap;
    rc = launch_search(h, p, 2);
  }
  if (rc == ANNB_OK) {
    cudaMemcpyAsync(L.hfound, L.dfound, (size_t)L.B * 4, cudaMemcpyDeviceToHost, st);
    if (L.host_labels) {
      cudaMemcpyAsync(L.host_labels, L.dl, (size_t)L.B * L.k * 8, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(L.host_dists, L.dd, (size_t)L.B * L.k * 4, cudaMemcpyDeviceToHost, st);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = ANNB_ECUDA;
  }
  h->stream = saved;
  if (rc == ANNB_ECUDA) annb_set_error("CUDA error while re-running flagged queries of a streamed search: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc != ANNB_OK) return rc;
  h->flagged_fallbacks++;
  h->flagged_fallback_queries += n_over;
  return ANNB_OK;
}

static int lane_wait(annb_index *h, int lane) {
  annb_index::AsyncLane &L = h->lanes[lane];
  if (!L.busy) return ANNB_OK;
  cudaStream_t st = lane ? h->stream2 : h->stream;
  ANNB_CUDA(cudaStreamSynchronize(st));
  L.busy = false;
  if (L.flagged) {
    int64_t n_over = 0;
    for (int64_t b = 0; b < L.B; b++) n_over += L.hfound[b] < 0;
    if (n_over) ANNB_TRY(lane_rerun_flagged(h, lane, n_over));
  }
  for (int64_t b = 0; b < L.B; b++)
    if (L.hfound[b] < L.k)
      ANNB_FAIL(ANNB_EFEWRESULTS, "Cannot return the results in a contigious 2D array. Probably ef or M is too small");
  return ANNB_OK;
}

// one batch, enqueued entirely on one lane: H2D -> normalise -> [filter bitmap] -> walk (tables built inside) -> D2H
static int submit_impl(annb_index *h, const float *queries, int in_space, int64_t B, int normalize, int k, int ef,
                       const uint64_t *filter_labels, int filter_space, int64_t n_filter, uint64_t *labels_out, float *dists_out,
                       int out_space, int *ticket_out) {
  if (!queries || !labels_out || !dists_out || !ticket_out || B <= 0 || k <= 0 || ef <= 0) ANNB_FAIL(ANNB_EINVAL, "bad arguments");
  if (n_filter < 0 || (n_filter > 0 && !filter_labels)) ANNB_FAIL(ANNB_EINVAL, "bad filter");
  if (!h->g.inited) ANNB_FAIL(ANNB_ESTATE, "index not initialised");
  if (!h->d_codebook) ANNB_FAIL(ANNB_ESTATE, "Please train the PQ before using HNSW quantization backend");
  if (h->opt_force_general == 2) ANNB_FAIL(ANNB_EINVAL, "streamed searches do not run on the bitmap walk: use annb_search");
  if (h->dev_dirty || h->deleted_dirty || h->patch_pending) {  // re-upload of the graph: nothing may be in flight
    ANNB_TRY(lane_wait(h, 0));
    ANNB_TRY(lane_wait(h, 1));
    ANNB_TRY(sync_device_graph(h));
  }
  if (h->gd.n == 0) ANNB_FAIL(ANNB_EFEWRESULTS, "Cannot return the results in a contigious 2D array. Probably ef or M is too small");
  const int ef_eff = std::max(ef, k);
  if (ef_eff > ANNB_MAX_EF) ANNB_FAIL(ANNB_ELIMIT, "max(ef, k)=%d exceeds ANNB_MAX_EF=%d", ef_eff, ANNB_MAX_EF);
  const bool general = filter_labels != nullptr || h->g.num_deleted > 0 || h->opt_force_general;
  const double selectivity = filter_labels ? std::min<double>(1.0, (double)n_filter / (double)std::max<int64_t>(1, h->gd.n))
                                           : 1.0 - (double)h->g.num_deleted / (double)std::max<int64_t>(1, h->gd.n);
  const bool sparse_labels = h->max_label > (uint64_t)h->gd.n * 64 + (1ull << 30);
  const int ticket = (int)(h->next_ticket++ & 0x3fffffff);
  const int lane = ticket & 1;
  ANNB_TRY(lane_wait(h, lane));  // a lane holds one batch at a time
  const bool host_in = in_space != ANNB_DEVICE, host_out = out_space != ANNB_DEVICE;
  const size_t TS = (size_t)h->M * h->Ks;
  const bool fuse = walk4_can_fuse(h) && (!general || walk4f_applicable(h, ef_eff, selectivity));
  float *dq = nullptr, *dtab = nullptr, *dd = dists_out;
  uint64_t *dl = labels_out;
  int32_t *dfound;
  unsigned int *counters;
  uint32_t *by_id = nullptr, *by_label = nullptr;
  uint64_t *dflt = nullptr;
  // all scratch first: a (re)allocation synchronises both lanes, which is only safe before enqueuing
  if (host_in || normalize > 0)
    ANNB_TRY(annb_scratch(h, lane ? S_L1_QUERIES : S_QUERIES, (size_t)B * h->dim * sizeof(float), (void **)&dq));
  if (!fuse) ANNB_TRY(annb_scratch(h, lane ? S_L1_TABLES : S_TABLES, (size_t)B * TS * sizeof(float), (void **)&dtab));
  if (host_out) {
    ANNB_TRY(annb_scratch(h, lane ? S_L1_OUT_L : S_OUT_L, (size_t)B * k * 8, (void **)&dl));
    ANNB_TRY(annb_scratch(h, lane ? S_L1_OUT_D : S_OUT_D, (size_t)B * k * 4, (void **)&dd));
  }
  ANNB_TRY(annb_scratch(h, lane ? S_L1_FOUND : S_L0_FOUND, (size_t)B * 4, (void **)&dfound));
  ANNB_TRY(annb_scratch(h, S_LANE_COUNTERS, 512, (void **)&counters));
  if (filter_labels) {
    ANNB_TRY(annb_scratch(h, lane ? S_L1_FLT_BY_ID : S_FLT_BY_ID, ((size_t)(h->gd.n + 31) / 32 + 1) * 4, (void **)&by_id));
    if (!sparse_labels) {
      ANNB_TRY(annb_scratch(h, lane ? S_L1_FLT_BY_LABEL : S_FLT_BY_LABEL, ((size_t)(h->max_label >> 5) + 1) * 4, (void **)&by_label));
      if (filter_space != ANNB_DEVICE)
        ANNB_TRY(annb_scratch(h, lane ? S_L1_FLT_LABELS : S_FLT_LABELS, (size_t)std::max<int64_t>(n_filter, 1) * 8, (void **)&dflt));
    }
  }
  int32_t *hfound;
  ANNB_TRY(annb_pinned(h, 4 + lane, (size_t)B * 4, (void **)&hfound));

  cudaStream_t saved = h->stream;
  h->stream = lane ? h->stream2 : saved;
  int rc = ANNB_OK;
  const float *src = queries;
  if (dq) {
    if (cudaMemcpyAsync(dq, queries, (size_t)B * h->dim * sizeof(float), host_in ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                        h->stream) != cudaSuccess)
      rc = ANNB_ECUDA;
    src = dq;
  }
  for (int r = 0; r < normalize && rc == ANNB_OK; r++) rc = launch_l2_normalize(h, dq, B, h->dim);
  if (rc == ANNB_OK && filter_labels) {
    if (sparse_labels) {
      // labels too sparse for a by-label bitmap on the device: resolve them through the host's label -> id map
      std::vector<uint64_t> tmp;
      const uint64_t *hl = filter_labels;
      if (filter_space == ANNB_DEVICE) {
        tmp.resize((size_t)n_filter);
        if (cudaMemcpy(tmp.data(), filter_labels, (size_t)n_filter * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = ANNB_ECUDA;
        hl = tmp.data();
      }
      std::vector<uint32_t> bm((size_t)(h->gd.n + 31) / 32 + 1, 0u);
      for (int64_t i = 0; i < n_filter && rc == ANNB_OK; i++) {
        auto it = h->g.label_lookup.find(hl[i]);
        if (it != h->g.label_lookup.end()) bm[it->second >> 5] |= 1u << (it->second & 31);
      }
      if (rc == ANNB_OK && (cudaMemcpyAsync(by_id, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                            cudaStreamSynchronize(h->stream) != cudaSuccess))
        rc = ANNB_ECUDA;
    } else {
      const uint64_t *dfl = filter_labels;
      if (dflt) {
        if (cudaMemcpyAsync(dflt, filter_labels, (size_t)n_filter * 8, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) rc = ANNB_ECUDA;
        dfl = dflt;
      }
      if (rc == ANNB_OK) rc = launch_filter_bitmap(h, dfl, n_filter, by_label, by_id);
    }
  }
  if (rc == ANNB_OK && !fuse) rc = launch_adc_table(h, src, B, dtab);
  if (rc == ANNB_OK) {
    SearchParams p;
    memset(&p, 0, sizeof(p));
    if (fuse) fuse_params(h, p, src);
    else p.tables = dtab;
    p.B = B;
    p.k = k;
    p.ef = ef_eff;
    p.filter = by_id;
    p.out_labels = dl;
    p.out_dists = dd;
    p.out_found = dfound;
    p.work_counter = counters + 64 * lane;
    p.selectivity = (float)selectivity;
    rc = launch_search(h, p, general ? 1 : 0);
  }
  if (rc == ANNB_OK) {
    cudaMemcpyAsync(hfound, dfound, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream);
    if (host_out) {
      cudaMemcpyAsync(labels_out, dl, (size_t)B * k * 8, cudaMemcpyDeviceToHost, h->stream);
      cudaMemcpyAsync(dists_out, dd, (size_t)B * k * 4, cudaMemcpyDeviceToHost, h->stream);
    }
    if (cudaGetLastError() != cudaSuccess) rc = ANNB_ECUDA;
  }
  h->stream = saved;
  if (rc != ANNB_OK) {
    if (rc == ANNB_ECUDA) annb_set_error("CUDA error while enqueuing a streamed search");
    return rc;
  }
  annb_index::AsyncLane &L = h->lanes[lane];
  L.busy = true;
  L.B = B;
  L.k = k;
  L.hfound = hfound;
  L.flagged = general;
  L.ef = ef_eff;
  L.dq = src;
  L.dfilter = by_id;
  L.selectivity = (float)selectivity;
  L.dl = dl;
  L.dd = dd;
  L.dfound = dfound;
  L.host_labels = host_out ? labels_out : nullptr;
  L.host_dists = host_out ? dists_out : nullptr;
  *ticket_out = ticket;
  return ANNB_OK;
}

int annb_search_submit(annb_index_t *h, const float *queries, int in_space, int64_t B, int normalize, int k, int ef,
                       uint64_t *labels_out, float *dists_out, int out_space, int *ticket_out) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  return submit_impl(h, queries, in_space, B, normalize, k, ef, nullptr, ANNB_HOST, 0, labels_out, dists_out, out_space, ticket_out);
}

int annb_search_submit_filtered(annb_index_t *h, const float *queries, int in_space, int64_t B, int normalize, int k, int ef,
                                const uint64_t *filter_labels, int filter_space, int64_t n_filter, uint64_t *labels_out,
                                float *dists_out, int out_space, int *ticket_out) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  if (!filter_labels && n_filter != 0) ANNB_FAIL(ANNB_EINVAL, "bad filter");
  return submit_impl(h, queries, in_space, B, normalize, k, ef, filter_labels, filter_space, n_filter, labels_out, dists_out, out_space,
                     ticket_out);
}

int annb_search_wait(annb_index_t *h, int ticket) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  if (ticket < 0) ANNB_FAIL(ANNB_EINVAL, "bad ticket");
  return lane_wait(h, ticket & 1);
}

int annb_merge_topk(annb_index_t *h, const uint64_t *labels_gbk, const float *dists_gbk, int G, int64_t B, int k,
                    uint64_t *labels_out, float *dists_out) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  if (!labels_gbk || !dists_gbk || !labels_out || !dists_out || G <= 0 || k <= 0 || B < 0) ANNB_FAIL(ANNB_EINVAL, "bad arguments");
  return launch_merge_topk(h, labels_gbk, dists_gbk, G, B, k, B * (int64_t)k, B * (int64_t)k, labels_out, dists_out, h->stream);
}

int annb_merge_topk_packed(annb_index_t *h, const void *packed, int G, int64_t B, int k, int64_t rank_stride_bytes,
                           int64_t labels_offset_bytes, uint64_t *labels_out, float *dists_out, int lane) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  if (!packed || !labels_out || !dists_out || G <= 0 || k <= 0 || B < 0 || lane < 0 || lane > 1)
    ANNB_FAIL(ANNB_EINVAL, "bad arguments");
  if ((rank_stride_bytes & 7) || (labels_offset_bytes & 7) || labels_offset_bytes < B * (int64_t)k * 4 ||
      rank_stride_bytes < labels_offset_bytes + B * (int64_t)k * 8 || (reinterpret_cast<uintptr_t>(packed) & 7))
    ANNB_FAIL(ANNB_EINVAL, "packed shard results: offsets must be 8-byte aligned and hold (B,k) fp32 + (B,k) u64");
  const float *d = reinterpret_cast<const float *>(packed);
  const uint64_t *l = reinterpret_cast<const uint64_t *>(reinterpret_cast<const uint8_t *>(packed) + labels_offset_bytes);
  return launch_merge_topk(h, l, d, G, B, k, rank_stride_bytes / 8, rank_stride_bytes / 4, labels_out, dists_out,
                           lane ? h->stream2 : h->stream);
}

int annb_lane_stream(annb_index_t *h, int lane, uint64_t *stream_out) {
  if (!h || !stream_out || lane < 0 || lane > 1) ANNB_FAIL(ANNB_EINVAL, "bad arguments");
  *stream_out = (uint64_t)(uintptr_t)(lane ? h->stream2 : h->stream);
  return ANNB_OK;
}

int annb_last_kernel_ms(annb_index_t *h, float *table_ms, float *search_ms, float *scan_ms) {
  ANNB_ENTER(h);
  ANNB_NEED_GPU(h);
  ANNB_CUDA(cudaStreamSynchronize(h->stream));
  float v;
  if (table_ms) *table_ms = (cudaEventElapsedTime(&v, h->ev[0], h->ev[1]) == cudaSuccess) ? v : -1.f;
  if (search_ms) *search_ms = (cudaEventElapsedTime(&v, h->ev[2], h->ev[3]) == cudaSuccess) ? v : -1.f;
  if (scan_ms) *scan_ms = (cudaEventElapsedTime(&v, h->ev[4], h->ev[5]) == cudaSuccess) ? v : -1.f;
  cudaGetLastError();
  return ANNB_OK;
}

int annb_launch_count(annb_index_t *h, int64_t *out) {
  if (!h || !out) ANNB_FAIL(ANNB_EINVAL, "null argument");
  *out = h->launches;
  return ANNB_OK;
}

int annb_fallback_count(annb_index_t *h, int64_t *out) {
  if (!h || !out) ANNB_FAIL(ANNB_EINVAL, "null argument");
  *out = h->flagged_fallbacks;
  return ANNB_OK;
}

int annb_sync_counts(annb_index_t *h, int64_t *full_syncs, int64_t *patches) {
  if (!h) ANNB_FAIL(ANNB_EINVAL, "null argument");
  if (full_syncs) *full_syncs = h->full_syncs;
  if (patches) *patches = h->patches;
  return ANNB_OK;
}

int annb_fallback_queries(annb_index_t *h, int64_t *out) {
  if (!h || !out) ANNB_FAIL(ANNB_EINVAL, "null argument");
  *out = h->flagged_fallback_queries;
  return ANNB_OK;
}

int annb_set_option(annb_index_t *h, const char *name, int64_t value) {
  ANNB_ENTER(h);
  if (!name) ANNB_FAIL(ANNB_EINVAL, "null option name");
  if (!strcmp(name, "warps_per_cta")) h->opt_warps_per_cta = value;
  else if (!strcmp(name, "ctas_per_sm")) h->opt_ctas_per_sm = value;
  else if (!strcmp(name, "force_general")) h->opt_force_general = value;
  else if (!strcmp(name, "timing")) h->opt_timing = value;
  else if (!strcmp(name, "ip_raw")) h->opt_ip_raw = value;
  else if (!strcmp(name, "chunks")) h->opt_chunks = value;
  else if (!strcmp(name, "flagged_epl")) h->opt_flagged_epl = value;
  else if (!strcmp(name, "flagged_kernel")) h->opt_flagged_kernel = value;
  else if (!strcmp(name, "flagged_en")) h->opt_flagged_en = value;
  else if (!strcmp(name, "scan_kernel")) h->opt_scan_kernel = value;
  else if (!strcmp(name, "walk_kernel")) h->opt_walk_kernel = value;
  else if (!strcmp(name, "dump_tables")) h->opt_dump_tables = value;
  else if (!strcmp(name, "prefetch")) h->opt_prefetch = value;
  else if (!strcmp(name, "gpu_build")) h->opt_gpu_build = value;
  else if (!strcmp(name, "gpu_build_frac")) h->opt_gpu_build_frac = value;
  else if (!strcmp(name, "reset_counters")) h->flagged_fallbacks = h->flagged_fallback_queries = 0;
  else ANNB_FAIL(ANNB_EINVAL, "unknown option %s", name);
  return ANNB_OK;
}

}  // extern "C"
