/*
 * TEST INFRASTRUCTURE — the oracle.  NOT product code.
 *
 * Plain-C restatement of the CPU algorithm of epsilla-cloud/vectordb's vector-search hot path
 * (SURVEY.md §8a rows A1-A14), each function citing the reference file:line it follows
 * (paths relative to /root/reference/engine).  It exists so the parity tests have a checker that
 * travels to the GPU box (the reference tree does not).
 *
 * Pinning: tests/test_oracle.py checks every function here against oracle/_ref/libepsilla_ref.so
 * (the reference's own sources compiled unmodified, oracle/Makefile) on seeded inputs — distances
 * BIT-exact, search results identical — and against the reference test-suite's known answers
 * (engine/test/engine/db/db_server.cpp: DenseVector orders :289-292, DenseVectorFilter :1620-1627,
 * half-circle exact top-500 :1133-1200) and the committed fixtures under tests/golden/ that were
 * generated from libepsilla_ref.so by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (vectordb_b200/) never does.
 *
 * Search restates the reference at IntraQueryThreads = 1, the only configuration in which the
 * reference is a pure function of (graph, vectors, query, L).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PORT_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* A1/A2: distances.                                                                            */
/*                                                                                              */
/* db/index/distance_simd.cpp:179-188 (fvec_inner_product) and :204-215 (fvec_L2sqr) are scalar */
/* loops compiled under "unroll-loops,associative-math,no-signed-zeros"                         */
/* (db/index/platform_macros.hpp:135-140) with no -march (CMakeLists.txt:4-8), i.e. SSE2.  What */
/* GCC 13.3 -O3 makes of them (objdump of oracle/_ref/obj/db/index/distance_simd.o) is: one     */
/* 4-lane accumulator, lane j summing elements j, j+4, j+8, ... in index order with separate    */
/* mul and add (SSE2 has no FMA); horizontal sum (l0+l2)+(l1+l3); a 2-element tail is added     */
/* lane-wise before the horizontal sum, a last odd element after it (L2) / scalar tail after it (IP); d<4 scalar.  */
/* This file is compiled with -ffp-contract=off-equivalent flags (no -march => no FMA) so the   */
/* order below IS the reference's arithmetic; test_oracle.py asserts bit equality vs _ref.      */
/* ------------------------------------------------------------------------------------------ */
static float sum_lanes(const float* x, const float* y, size_t d, int l2) {
  if (d < 4) {
    float res = 0.f;
    for (size_t i = 0; i < d; ++i) {
      float t = l2 ? (x[i] - y[i]) : x[i];
      float u = l2 ? t : y[i];
      res += t * u;
    }
    return res;
  }
  volatile float keep; /* forbid the compiler from re-vectorising into another order */
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  size_t d4 = d & ~(size_t)3;
  for (size_t i = 0; i < d4; i += 4) {
    float t0, t1, t2, t3;
    if (l2) {
      t0 = x[i] - y[i]; t1 = x[i + 1] - y[i + 1]; t2 = x[i + 2] - y[i + 2]; t3 = x[i + 3] - y[i + 3];
      t0 = t0 * t0; t1 = t1 * t1; t2 = t2 * t2; t3 = t3 * t3;
    } else {
      t0 = x[i] * y[i]; t1 = x[i + 1] * y[i + 1]; t2 = x[i + 2] * y[i + 2]; t3 = x[i + 3] * y[i + 3];
    }
    a0 += t0; a1 += t1; a2 += t2; a3 += t3;
  }
  float s0 = a0 + a2, s1 = a1 + a3;
  size_t rem = d - d4;
  float res;
  if (l2) {
    /* fvec_L2sqr tail: a 2-element remainder is added lane-wise before the horizontal sum, a last
     * odd element after it. */
    if (rem >= 2) {
      float t0 = x[d4] - y[d4], t1 = x[d4 + 1] - y[d4 + 1];
      t0 = t0 * t0; t1 = t1 * t1;
      res = (s0 + t0) + (s1 + t1);
      d4 += 2;
      rem -= 2;
    } else {
      res = s0 + s1;
    }
    if (rem == 1) { float t = x[d4] - y[d4]; res += t * t; }
  } else {
    /* fvec_inner_product tail: horizontal sum first, then the remainder one scalar at a time. */
    res = s0 + s1;
    for (size_t i = d4; i < d; ++i) res += x[i] * y[i];
  }
  keep = res;
  return keep;
}

/* db/index/distance_simd.cpp:204-215 */
PORT_API float port_l2sqr(const float* x, const float* y, int64_t d) { return sum_lanes(x, y, (size_t)d, 1); }
/* db/index/distance_simd.cpp:179-188 */
PORT_API float port_inner_product(const float* x, const float* y, int64_t d) { return sum_lanes(x, y, (size_t)d, 0); }

/* db/index/index.cpp:10-35 (GetDistFunc); space_l2.hpp:8-26 (L2Sqr), space_ip.hpp:8-20
 * (InnerProduct = -ip), space_cosine.hpp:8-16 (CosineDistance = 1 - 1.0f*ip).
 * metric: 1 EUCLIDEAN, 2 COSINE, 3 DOT_PRODUCT (db/catalog/meta_types.hpp MetricType); other -> L2. */
PORT_API float port_distance(int metric, const float* a, const float* b, int64_t d) {
  switch (metric) {
    case 2: return 1 - 1.0f * port_inner_product(a, b, d);
    case 3: return -port_inner_product(a, b, d);
    default: return port_l2sqr(a, b, d);
  }
}

/* A14: db/vector.cpp:60-69 (Normalize): scalar fp32 sum, sqrt, divide.  (The insert path applies
 * it only when the squared norm exceeds 1e-10, db/table_segment_mvp.cpp:574-587.) */
PORT_API void port_normalize(float* v, int64_t d) {
  volatile float sum = 0;
  for (int64_t i = 0; i < d; i++) sum += v[i] * v[i];
  float s = sqrtf(sum);
  for (int64_t i = 0; i < d; i++) v[i] /= s;
}

/* A13: utils/concurrent_bitset.cpp:9-11 */
static int bit_test(const uint8_t* bits, int64_t id) { return bits ? (bits[id >> 3] & (1 << (id & 7))) != 0 : 0; }

/* ------------------------------------------------------------------------------------------ */
/* A12: filter evaluation.  query/expr/expr_evaluator.cpp.                                      */
/* Node POD = the fields of ExprNode (query/expr/expr_types.hpp:77-90) that numeric/bool        */
/* predicates use, with field_name already resolved to its byte offset in the attribute row     */
/* (field_name_mem_offset_map_) or -2 for "@distance".                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t node_type;  /* NodeType, expr_types.hpp:11-48 */
  int64_t value_type; /* ValueType, expr_types.hpp:67-74 */
  int64_t left, right; /* size_t; "none" is (size_t)-1 */
  int64_t int_value;
  double double_value;
  int64_t bool_value;
  int64_t field_offset;
} port_node;

enum { /* NodeType ordinals */
  NT_Invalid, NT_IntConst, NT_StringConst, NT_DoubleConst, NT_BoolConst, NT_Int1Attr, NT_Int2Attr,
  NT_Int4Attr, NT_Int8Attr, NT_StringAttr, NT_DoubleAttr, NT_FloatAttr, NT_BoolAttr, NT_GeoPointAttr,
  NT_Add, NT_Subtract, NT_Multiply, NT_Divide, NT_Module, NT_LT, NT_LTE, NT_EQ, NT_GT, NT_GTE, NT_NE,
  NT_AND, NT_OR, NT_NOT
};
enum { VT_STRING, VT_INT, VT_DOUBLE, VT_BOOL };

typedef struct {
  const port_node* nodes;
  int64_t n_nodes;
  const char* attrs;
  int64_t stride;
} filt;

/* expr_evaluator.cpp:127-164 */
static double num_eval(const filt* f, int64_t idx, int64_t id, double distance) {
  const port_node* r = &f->nodes[idx];
  int64_t t = r->node_type;
  if (t == NT_IntConst) return (double)r->int_value;
  if (t == NT_DoubleConst) return r->double_value;
  if (t == NT_Int1Attr || t == NT_Int2Attr || t == NT_Int4Attr || t == NT_Int8Attr) {
    const char* p = f->attrs + r->field_offset + id * f->stride; /* :61-92 */
    int64_t v;
    if (t == NT_Int1Attr) { int8_t x; memcpy(&x, p, 1); v = x; }
    else if (t == NT_Int2Attr) { int16_t x; memcpy(&x, p, 2); v = x; }
    else if (t == NT_Int4Attr) { int32_t x; memcpy(&x, p, 4); v = x; }
    else { int64_t x; memcpy(&x, p, 8); v = x; }
    return (double)v;
  }
  if (t == NT_DoubleAttr || t == NT_FloatAttr) {
    if (r->field_offset == -2) return distance; /* "@distance", :143-145 */
    const char* p = f->attrs + r->field_offset + id * f->stride; /* :94-105 */
    if (t == NT_DoubleAttr) { double x; memcpy(&x, p, 8); return x; }
    float x; memcpy(&x, p, 4); return (double)x;
  }
  if (r->left != -1 && r->right != -1) {
    double a = num_eval(f, r->left, id, distance), b = num_eval(f, r->right, id, distance);
    switch (t) {
      case NT_Add: return a + b;
      case NT_Subtract: return a - b;
      case NT_Multiply: return a * b;
      case NT_Divide: return a / b;
      case NT_Module: return fmod(a, b);
    }
  }
  return 0.0;
}

/* expr_evaluator.cpp:170-258.  NB: NOT / AND / OR / bool-EQ children are evaluated through the
 * two-argument overload (:166-168), i.e. with distance 0 — "@distance" only sees the real
 * distance in a top-level comparison.  String, IN, LIKE, NEARBY nodes are out of scope (§2 row 9)
 * and never reach this evaluator (the host rejects them). */
static int logical_eval(const filt* f, int64_t idx, int64_t id, double distance) {
  if (idx < 0) return 1; /* :171-173 */
  const port_node* r = &f->nodes[idx];
  int64_t t = r->node_type;
  if (t == NT_BoolConst) return r->bool_value != 0;
  if (t == NT_BoolAttr) { /* :56-59: the byte VALUE is cast to a pointer -> "non-zero byte" */
    return f->attrs[r->field_offset + id * f->stride] != 0;
  }
  if (t == NT_NOT) return !logical_eval(f, r->left, id, 0);
  if (r->left != -1 && r->right != -1) {
    if (t == NT_EQ || t == NT_NE) {
      int64_t cvt = f->nodes[r->left].value_type;
      if (cvt == VT_BOOL) {
        int a = logical_eval(f, r->left, id, 0), b = logical_eval(f, r->right, id, 0);
        return t == NT_EQ ? a == b : a != b;
      }
      double a = num_eval(f, r->left, id, distance), b = num_eval(f, r->right, id, distance);
      return t == NT_EQ ? a == b : a != b;
    }
    if (t == NT_AND || t == NT_OR) {
      int a = logical_eval(f, r->left, id, 0), b = logical_eval(f, r->right, id, 0);
      return t == NT_AND ? (a && b) : (a || b);
    }
    double a = num_eval(f, r->left, id, distance), b = num_eval(f, r->right, id, distance);
    switch (t) {
      case NT_GT: return a > b;
      case NT_GTE: return a >= b;
      case NT_LT: return a < b;
      case NT_LTE: return a <= b;
    }
  }
  return 0;
}

PORT_API int port_filter_eval(const port_node* nodes, int64_t n_nodes, const char* attrs, int64_t stride,
                              int64_t id, double distance) {
  filt f = {nodes, n_nodes, attrs, stride};
  return logical_eval(&f, n_nodes - 1, id, distance); /* root = last node, vec_search_executor.cpp:848 */
}

/* ------------------------------------------------------------------------------------------ */
/* A7: candidate queue.  db/execution/candidate.hpp:7-23                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t id;
  float dist;
  uint8_t checked;
} cand;

static int cand_lt(const cand* a, const cand* b) { /* candidate.hpp:16-22 */
  if (a->dist != b->dist) return a->dist < b->dist;
  return a->id < b->id;
}
static int cand_cmp(const void* a, const void* b) {
  const cand *x = (const cand*)a, *y = (const cand*)b;
  return cand_lt(x, y) ? -1 : (cand_lt(y, x) ? 1 : 0);
}
static int64_t lower_bound(const cand* q, int64_t n, const cand* c) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = lo + (hi - lo) / 2;
    if (cand_lt(&q[mid], c)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* vec_search_executor.cpp:75-117 (AddIntoQueue), queue_start folded into q. */
static int64_t add_into_queue(cand* q, int64_t* size, int64_t cap, const cand* c) {
  if (*size == 0) { q[(*size)++] = *c; return 0; }
  int64_t end = *size;
  int64_t loc = lower_bound(q, end, c);
  if (loc != end) {
    if (c->id == q[loc].id) return cap; /* duplicate */
    if (*size >= cap) { --*size; --end; }
  } else {
    if (*size < cap) { q[loc] = *c; ++*size; return *size - 1; }
    return cap;
  }
  memmove(q + loc + 1, q + loc, (size_t)(end - loc) * sizeof(cand));
  q[loc] = *c;
  ++*size;
  return loc;
}

/* vec_search_executor.cpp:137-148 (InsertOneElementAt) */
static void insert_one_at(const cand* c, cand* q, int64_t idx, int64_t size) {
  memmove(q + idx + 1, q + idx, (size_t)(size - idx - 1) * sizeof(cand));
  q[idx] = *c;
}

/* vec_search_executor.cpp:150-217 (MergeTwoQueuesInto1stQueueSeqFixed) */
static int64_t merge_fixed(cand* q1, int64_t n1, cand* q2, int64_t n2) {
  int64_t ins = lower_bound(q1, n1, &q2[0]);
  if (ins == n1) return ins;
  if (ins == n1 - 1) { q1[ins] = q2[0]; return ins; }
  if (q2[0].id != q1[ins].id) insert_one_at(&q2[0], q1, ins, n1);
  else if (!q2[0].checked && q1[ins].checked) q1[ins].checked = 0;
  if (n2 == 1) return ins;
  int64_t i1 = ins + 1, i2 = 1;
  for (int64_t at = ins + 1; at < n1; ++at) {
    if (i1 >= n1 || i2 >= n2) break;
    if (cand_lt(&q1[i1], &q2[i2])) { ++i1; }
    else if (cand_lt(&q2[i2], &q1[i1])) { insert_one_at(&q2[i2++], q1, at, n1); ++i1; }
    else {
      if (!q2[i2].checked && q1[i1].checked) q1[i1].checked = 0;
      ++i2; ++i1;
    }
  }
  return ins;
}

PORT_API int64_t port_add_into_queue(int64_t* ids, float* dists, uint8_t* checked, int64_t* size, int64_t cap,
                                     int64_t id, float dist) {
  /* array-of-fields wrapper for unit tests */
  cand* q = (cand*)malloc(sizeof(cand) * (size_t)(cap + 1));
  for (int64_t i = 0; i < *size; ++i) { q[i].id = ids[i]; q[i].dist = dists[i]; q[i].checked = checked[i]; }
  cand c = {id, dist, 0};
  int64_t r = add_into_queue(q, size, cap, &c);
  for (int64_t i = 0; i < *size; ++i) { ids[i] = q[i].id; dists[i] = q[i].dist; checked[i] = q[i].checked; }
  free(q);
  return r;
}

/* ------------------------------------------------------------------------------------------ */
/* The index view the search functions read (the executor's ctor arguments + the segment fields */
/* of SURVEY.md §8b "Reads from the segment").                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t metric;         /* 1 L2, 2 COSINE, 3 IP */
  int32_t prefilter;      /* prefilter_enabled_ */
  int64_t dim;
  const float* vectors;   /* vector_tables_[f], row-major */
  int64_t total_rows;     /* table_segment->record_number_ at Search time */
  int64_t n_indexed;      /* ann_index->record_number_ */
  const int64_t* offsets; /* offset_table_[n_indexed+1] */
  const int64_t* nbrs;    /* neighbor_list_ */
  int64_t nav;            /* start_search_point_ */
  const uint8_t* deleted; /* ConcurrentBitset bytes (may be NULL) */
  const char* attrs;      /* attribute_table_ */
  int64_t attr_stride;    /* primitive_offset_ */
  const port_node* filter;
  int64_t n_filter;
  int64_t L_master, L_local;
} port_index;

static float dist_to(const port_index* ix, int64_t row, const float* q) {
  return port_distance(ix->metric, ix->vectors + ix->dim * row, q, ix->dim);
}

/* A4: vec_search_executor.cpp:487-516 (PrepareInitIds).  Q1 guard: the reference never terminates
 * when L > n_indexed; callers clamp L first (port_search does). */
PORT_API void port_prepare_init_ids(const int64_t* offsets, const int64_t* nbrs, int64_t nav, int64_t n_indexed,
                                    int64_t L, int64_t* init_ids) {
  uint8_t* sel = (uint8_t*)calloc((size_t)n_indexed, 1);
  int64_t end = 0;
  for (int64_t e = offsets[nav]; e < offsets[nav + 1] && end < L; ++e) {
    int64_t v = nbrs[e];
    if (sel[v]) continue;
    sel[v] = 1;
    init_ids[end++] = v;
  }
  int64_t tmp = nav + 1;
  while (end < L) {
    if (tmp == n_indexed) tmp = 0;
    int64_t v = tmp++;
    if (sel[v]) continue;
    sel[v] = 1;
    init_ids[end++] = v;
  }
  free(sel);
}

/* A5/A6: vec_search_executor.cpp:518-715 (SearchImpl) at num_threads_ = 1, with
 * :446-485 (InitializeSetLPara) and :384-444 (ExpandOneCandidate) inlined.
 * With one thread PickTopMToWorkers (:328-356) moves nothing (dest_queue 0 IS the master) and only
 * counts the unchecked entries at/after k_master; the "parallel" block is the master's own loop of
 * <= subsearch_iterations expansions; MergeAllQueuesToMaster has no workers.  The net effect is
 * best-first search: expand the first unchecked entry, k = (r <= k ? r : k+1), until no unchecked
 * entry remains.  dist_bound is a live alias of slot L-1 (:546). */
static void search_impl(const port_index* ix, const float* q, int64_t L, const int64_t* init_ids, cand* set_L,
                        uint8_t* visited, uint64_t* n_dist, uint64_t* n_expand) {
  int64_t size = 0;
  for (int64_t i = 0; i < L; ++i) visited[init_ids[i]] = 1;                 /* :455-457 */
  for (int64_t i = 0; i < L; ++i) {                                          /* :462-478 */
    set_L[i].id = init_ids[i];
    set_L[i].dist = dist_to(ix, init_ids[i], q);
    set_L[i].checked = 0;
  }
  qsort(set_L, (size_t)L, sizeof(cand), cand_cmp);                           /* :481-483 */
  size = L;
  const float* last_dist = &set_L[L - 1].dist;                               /* :546 */
  int64_t k = 0;
  for (;;) {
    /* first unchecked at/after k (PickTopMToWorkers count + worker loop skip, :339-343,:672-674) */
    while (k < size && set_L[k].checked) ++k;
    if (k >= size) break;
    set_L[k].checked = 1;
    int64_t c = set_L[k].id;
    int64_t nk = L;
    ++*n_expand;
    for (int64_t e = ix->offsets[c]; e < ix->offsets[c + 1]; ++e) {          /* :400 */
      int64_t nb = ix->nbrs[e];
      if (visited[nb]) continue;                                             /* :403-406 */
      visited[nb] = 1;
      ++*n_dist;
      float d = dist_to(ix, nb, q);
      if (d > *last_dist) continue;                                          /* :424 */
      cand cd = {nb, d, 0};
      int64_t r = add_into_queue(set_L, &size, L, &cd);                      /* :430-438 */
      if (r < nk) nk = r;
    }
    if (nk <= k) k = nk; else ++k;                                           /* :648-652 */
  }
}

/* A9: vec_search_executor.cpp:717-768 (BruteForceSearch) — distance for every row in [start,end),
 * then compaction dropping deleted / filter-failing rows (filter sees the distance), then sort. */
static int64_t brute_force(const port_index* ix, const float* q, int64_t start, int64_t end, cand* out,
                           uint64_t* n_dist) {
  filt f = {ix->filter, ix->n_filter, ix->attrs, ix->attr_stride};
  int64_t n = 0;
  for (int64_t v = start; v < end; ++v) {
    float d = dist_to(ix, v, q);
    ++*n_dist;
    if (!bit_test(ix->deleted, v) && logical_eval(&f, ix->n_filter - 1, v, (double)d)) {
      out[n].id = v; out[n].dist = d; out[n].checked = 0; ++n;
    }
  }
  qsort(out, (size_t)n, sizeof(cand), cand_cmp);
  return n;
}

/* A10: vec_search_executor.cpp:770-831 (PreFilterBruteForceSearch) — filter (distance 0) and
 * deleted first, distance only for passing rows. */
static int64_t prefilter_brute_force(const port_index* ix, const float* q, int64_t start, int64_t end, cand* out,
                                     uint64_t* n_dist) {
  filt f = {ix->filter, ix->n_filter, ix->attrs, ix->attr_stride};
  int64_t n = 0;
  for (int64_t v = start; v < end; ++v) {
    if (!bit_test(ix->deleted, v) && logical_eval(&f, ix->n_filter - 1, v, 0)) {
      out[n].id = v; out[n].dist = dist_to(ix, v, q); out[n].checked = 0; ++n;
      ++*n_dist;
    }
  }
  qsort(out, (size_t)n, sizeof(cand), cand_cmp);
  return n;
}

static int64_t min3(int64_t a, int64_t b, int64_t c) { int64_t m = a < b ? a : b; return m < c ? m : c; }

/* A11: vec_search_executor.cpp:833-935 (Search).  ids/dists must hold max(limit, L_master) entries
 * (Q2: the reference's PreFilter branch writes min(queue,limit) entries unguarded).
 * stats[0] += distance evaluations (seed included), stats[1] += expansions.
 * Returns result_size. */
PORT_API int64_t port_search(const port_index* ix, const float* q, int64_t limit, int64_t* ids, double* dists,
                             uint64_t* stats) {
  uint64_t n_dist = 0, n_expand = 0;
  filt f = {ix->filter, ix->n_filter, ix->attrs, ix->attr_stride};
  int64_t root = ix->n_filter - 1;
  int64_t total = ix->total_rows;
  int64_t result = 0;
  int brute = ix->n_indexed < 512; /* BruteforceThreshold, vec_search_executor.hpp:28, cpp:62 */
  if (ix->prefilter) {                                                       /* :855-861 */
    cand* bq = (cand*)malloc(sizeof(cand) * (size_t)(total > 0 ? total : 1));
    int64_t n = prefilter_brute_force(ix, q, 0, total, bq, &n_dist);
    result = n < limit ? n : limit;
    for (int64_t i = 0; i < result; ++i) { ids[i] = bq[i].id; dists[i] = bq[i].dist; }
    free(bq);
  } else if (brute) {                                                        /* :862-868 */
    cand* bq = (cand*)malloc(sizeof(cand) * (size_t)(total > 0 ? total : 1));
    int64_t n = brute_force(ix, q, 0, total, bq, &n_dist);
    result = min3(n, limit, ix->L_local);
    for (int64_t i = 0; i < result; ++i) { ids[i] = bq[i].id; dists[i] = bq[i].dist; }
    free(bq);
  } else {
    int64_t L = ix->L_master < ix->n_indexed ? ix->L_master : ix->n_indexed; /* Q1 clamp */
    int64_t search_limit = min3(ix->n_indexed, limit, ix->L_local);          /* :872 */
    cand* set_L = (cand*)malloc(sizeof(cand) * (size_t)(L + 1));
    int64_t* init_ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);
    uint8_t* visited = (uint8_t*)calloc((size_t)ix->n_indexed, 1);
    port_prepare_init_ids(ix->offsets, ix->nbrs, ix->nav, ix->n_indexed, L, init_ids);
    n_dist += (uint64_t)L;
    search_impl(ix, q, L, init_ids, set_L, visited, &n_dist, &n_expand);
    int64_t cand_num;
    if (total > ix->n_indexed) {                                             /* :885-904 */
      int64_t tail = total - ix->n_indexed;
      cand* bq = (cand*)malloc(sizeof(cand) * (size_t)tail);
      int64_t n = brute_force(ix, q, ix->n_indexed, total, bq, &n_dist);
      int64_t bsz = n < limit ? n : limit;
      if (bsz > 0) {
        merge_fixed(set_L, search_limit, bq, bsz);
        cand_num = L < total ? L : total;
      } else {
        cand_num = L < ix->n_indexed ? L : ix->n_indexed;
      }
      free(bq);
    } else {
      cand_num = L < ix->n_indexed ? L : ix->n_indexed;                      /* :917 */
    }
    if (cand_num > L) cand_num = L;
    for (int64_t k = 0; k < cand_num && result < search_limit; ++k) {        /* :906-914, :919-927 */
      int64_t id = set_L[k].id;
      if (bit_test(ix->deleted, id) || !logical_eval(&f, root, id, (double)set_L[k].dist)) continue;
      ids[result] = id;
      dists[result] = set_L[k].dist;
      ++result;
    }
    free(set_L); free(init_ids); free(visited);
  }
  if (stats) { stats[0] += n_dist; stats[1] += n_expand; }
  return result;
}

/* nq queries, outputs [nq x limit] (rows padded with id -1 / +inf). */
PORT_API void port_search_batch(const port_index* ix, const float* queries, int64_t nq, int64_t limit, int64_t* ids,
                                double* dists, int64_t* counts, uint64_t* stats) {
  int64_t cap = limit > ix->L_master ? limit : ix->L_master;
  int64_t* tid = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cap + 1));
  double* td = (double*)malloc(sizeof(double) * (size_t)(cap + 1));
  for (int64_t i = 0; i < nq; ++i) {
    int64_t n = port_search(ix, queries + i * ix->dim, limit, tid, td, stats);
    counts[i] = n;
    for (int64_t j = 0; j < limit; ++j) {
      ids[i * limit + j] = j < n ? tid[j] : -1;
      dists[i * limit + j] = j < n ? td[j] : INFINITY;
    }
  }
  free(tid); free(td);
}
