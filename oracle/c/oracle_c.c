/* ORACLE — TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with the product).
 *
 * Plain-C restatement of the integer / edge-order arithmetic of the hot path: the part of the reference's behaviour that
 * lives in PyG primitives (`torch-geometric>=2.3.0`, unpinned, /root/reference/requirements.txt:3, not
 * vendored -> PARITY UNPINNED against real PyG) and in ATen's CPU scatter kernels.  Scalar loops, one thread, no library:
 * an implementation that shares nothing with either the PyTorch shim in oracle/pyg_shim or the CUDA kernels, so that
 * agreement between the three is evidence about the semantics and not about a shared bug.
 * tests/test_oracle_c.py pins it bit for bit against the shim (which is pinned against the unmodified reference modules).
 *
 * Every function names the reference call site / App. A paragraph it restates.  All matrices are dense row-major.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* The sort the reference never does explicitly (SURVEY §8a note): perm = argsort(key_row, stable), col = other[perm],
 * rowptr = [0, cumsum(bincount(key_row, N))].  edge_index is [2, E] (row 0 = src, row 1 = dst); by_src = 0 groups by dst
 * (GINConv's flow source_to_target, src/models/gnn.py:41, App. A.1), by_src = 1 groups by src (the backward pass). */
int oc_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int by_src, int32_t* rowptr, int32_t* col, int32_t* eid) {
  const int64_t* key = edge_index + (by_src ? 0 : E);
  const int64_t* other = edge_index + (by_src ? E : 0);
  int64_t* next = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
  if (!next) return -1;
  for (int64_t e = 0; e < E; ++e) {
    if (key[e] < 0 || key[e] >= N) { free(next); return -2; }
    next[key[e] + 1]++;
  }
  for (int64_t i = 0; i < N; ++i) next[i + 1] += next[i];
  for (int64_t i = 0; i <= N; ++i) rowptr[i] = (int32_t)next[i];
  for (int64_t e = 0; e < E; ++e) {            /* ascending e = stable: neighbours keep the original edge order */
    const int64_t p = next[key[e]]++;
    col[p] = (int32_t)other[e];
    if (eid) eid[p] = (int32_t)e;
  }
  free(next);
  return 0;
}

/* Batch.batch (sorted graph id per node, App. A.6) -> ptr [S+1]. */
void oc_segment_ptr(const int64_t* ids, int64_t n, int64_t S, int32_t* ptr) {
  int64_t r = 0;
  for (int64_t s = 0; s <= S; ++s) {
    while (r < n && ids[r] < s) ++r;
    ptr[s] = (int32_t)r;
  }
}

static int cmp_i64(const void* a, const void* b) {
  const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return (x > y) - (x < y);
}

/* coalesce (App. A.4; to_undirected at src/pretrain/tasks.py:108 is coalesce(cat([ei, ei.flip(0)]))): sort by row*N+col,
 * drop duplicates.  symmetrise != 0 adds the flipped edges first.  out must hold 2 * (symmetrise ? 2E : E) values and is
 * written as [2, count] row-major with leading dimension `count`.  Returns count. */
int64_t oc_coalesce(const int64_t* edge_index, int64_t E, int64_t N, int symmetrise, int64_t* out) {
  const int64_t M = symmetrise ? 2 * E : E;
  int64_t* key = (int64_t*)malloc((size_t)(M > 0 ? M : 1) * sizeof(int64_t));
  if (!key) return -1;
  for (int64_t e = 0; e < E; ++e) {
    key[e] = edge_index[e] * N + edge_index[E + e];
    if (symmetrise) key[E + e] = edge_index[E + e] * N + edge_index[e];
  }
  qsort(key, (size_t)M, sizeof(int64_t), cmp_i64);
  int64_t count = 0;
  for (int64_t i = 0; i < M; ++i)
    if (i == 0 || key[i] != key[i - 1]) key[count++] = key[i];
  for (int64_t i = 0; i < count; ++i) {
    out[i] = key[i] / N;
    out[count + i] = key[i] % N;
  }
  free(key);
  return count;
}

/* GINConv before its MLP (App. A.1/A.2, src/models/gnn.py:41): out = zeros; for e in edge order: out[dst_e] += x[src_e];
 * out += (1 + eps) * x.  transposed != 0 swaps the roles of src and dst (the backward pass w.r.t. x).
 * fp32 adds in edge order, product and sum rounded separately (no fused multiply-add) — what ATen's scatter_add_ and the
 * elementwise (1+eps)*x + out do. */
void oc_gin_aggregate(const float* x, int64_t N, int64_t F, const int64_t* edge_index, int64_t E, float eps, int transposed,
                      int with_self, float* out) {
  const int64_t* from = edge_index + (transposed ? E : 0);
  const int64_t* to = edge_index + (transposed ? 0 : E);
  memset(out, 0, (size_t)N * (size_t)F * sizeof(float));
  for (int64_t e = 0; e < E; ++e) {
    const float* s = x + from[e] * F;
    float* d = out + to[e] * F;
    for (int64_t f = 0; f < F; ++f) d[f] = d[f] + s[f];
  }
  if (with_self) {                           /* built with -ffp-contract=off: the product is rounded before the add */
    const float scale = 1.0f + eps;
    for (int64_t i = 0; i < N * F; ++i) {
      const float prod = scale * x[i];
      out[i] = out[i] + prod;
    }
  }
}

/* global_mean_pool / global_max_pool / sum (App. A.2/A.3; src/models/finetune_model.py:75, src/pretrain/tasks.py:241-246,
 * 299,331).  mode 0 = sum, 1 = mean (count clamped to >= 1), 2 = max on a ZERO-initialised output with include_self=False
 * (an empty graph keeps its 0 row). */
void oc_segment_pool(const float* x, const int64_t* batch, int64_t N, int64_t F, int64_t B, int mode, float* out) {
  memset(out, 0, (size_t)B * (size_t)F * sizeof(float));
  int64_t* count = (int64_t*)calloc((size_t)(B > 0 ? B : 1), sizeof(int64_t));
  for (int64_t r = 0; r < N; ++r) {
    const int64_t g = batch[r];
    float* o = out + g * F;
    const float* v = x + r * F;
    if (mode == 2) {
      for (int64_t f = 0; f < F; ++f)
        if (count[g] == 0 || v[f] > o[f] || (v[f] != v[f])) o[f] = v[f];     /* first row replaces the zero init; NaN propagates */
    } else {
      for (int64_t f = 0; f < F; ++f) o[f] = o[f] + v[f];
    }
    count[g]++;
  }
  if (mode == 1)
    for (int64_t g = 0; g < B; ++g) {
      const float c = (float)(count[g] > 0 ? count[g] : 1);
      for (int64_t f = 0; f < F; ++f) out[g * F + f] = out[g * F + f] / c;
    }
  free(count);
}

/* Backward of the native amax path (App. A.3, probed on torch 2.11 CPU): dx[r,c] = g[s,c] * [x[r,c] == out[s,c]] / k,
 * k = number of tied rows in the segment, + 1 when out[s,c] == 0.0 (the zero-initialised self slot counts as a tie). */
void oc_segment_max_bwd(const float* grad_out, const float* x, const float* out, const int64_t* batch, int64_t N, int64_t F,
                        int64_t B, float* grad_x) {
  float* ties = (float*)calloc((size_t)(B * F > 0 ? B * F : 1), sizeof(float));
  for (int64_t r = 0; r < N; ++r)
    for (int64_t f = 0; f < F; ++f)
      if (x[r * F + f] == out[batch[r] * F + f]) ties[batch[r] * F + f] += 1.0f;
  for (int64_t i = 0; i < B * F; ++i)
    if (out[i] == 0.0f) ties[i] += 1.0f;
  for (int64_t r = 0; r < N; ++r)
    for (int64_t f = 0; f < F; ++f) {
      const int64_t s = batch[r] * F + f;
      grad_x[r * F + f] = (x[r * F + f] == out[s]) ? grad_out[s] / ties[s] : 0.0f;
    }
  free(ties);
}

/* MLPLinkPredictor's decoder input (src/models/heads.py:59-65): [h_u + h_v, h_u * h_v, |h_u - h_v|] per edge. */
void oc_lp_features(const float* h, int64_t H, const int64_t* edges, int64_t E, float* feat) {
  for (int64_t e = 0; e < E; ++e) {
    const float* u = h + edges[e] * H;
    const float* v = h + edges[E + e] * H;
    float* o = feat + e * 3 * H;
    for (int64_t k = 0; k < H; ++k) {
      o[k] = u[k] + v[k];
      o[H + k] = u[k] * v[k];
      o[2 * H + k] = fabsf(u[k] - v[k]);
    }
  }
}

/* The deterministic branch of negative_sampling (App. A.5) that every TU-sized graph takes at the reference's call site
 * (src/pretrain/tasks.py:107-111): all ordered non-edges (self loops excluded) of an n-node graph in ascending code order
 * row*(n-1)+col', truncated to `want`.  edges: [2, E] local ids.  out: [2, count] with leading dimension `want_cap`
 * (= capacity).  Returns count. */
int64_t oc_all_non_edges(const int64_t* edges, int64_t E, int64_t n, int64_t want, int64_t want_cap, int64_t* out) {
  const int64_t pop = n * n - n;
  unsigned char* taken = (unsigned char*)calloc((size_t)(pop > 0 ? pop : 1), 1);
  for (int64_t e = 0; e < E; ++e) {
    const int64_t r = edges[e], c = edges[E + e];
    if (r == c) continue;
    taken[r * (n - 1) + (r < c ? c - 1 : c)] = 1;
  }
  int64_t count = 0;
  for (int64_t code = 0; code < pop && count < want && count < want_cap; ++code) {
    if (taken[code]) continue;
    const int64_t r = code / (n - 1);
    int64_t c = code % (n - 1);
    if (r <= c) c += 1;
    out[count] = r;
    out[want_cap + count] = c;
    ++count;
  }
  free(taken);
  return count;
}

/* NFM's masked rows (src/models/pretrain_model.py:84-86, src/pretrain/tasks.py:82): gather and broadcast-scatter of rows. */
void oc_rows_gather(const float* x, int64_t F, const int64_t* idx, int64_t M, float* out) {
  for (int64_t i = 0; i < M; ++i) memcpy(out + i * F, x + idx[i] * F, (size_t)F * sizeof(float));
}
void oc_rows_fill(float* x, int64_t F, const int64_t* idx, int64_t M, const float* row) {
  for (int64_t i = 0; i < M; ++i) memcpy(x + idx[i] * F, row, (size_t)F * sizeof(float));
}
