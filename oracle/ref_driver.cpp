// TEST INFRASTRUCTURE — not product code.
//
// C-ABI driver around the UNMODIFIED reference hot path (epsilla-cloud/vectordb), compiled from
// the sources where they lie under /root/reference/engine by oracle/Makefile into
// oracle/_ref/libepsilla_ref.so.  Nothing here restates reference logic: it only constructs the
// reference's own objects (TableSegmentMVP, ANNGraphSegment, VecSearchExecutor, Expr parser) and
// forwards calls, exactly as SURVEY.md Appendix A describes.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.
#include <omp.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "db/ann_graph_segment.hpp"
#include "db/execution/aggregation.hpp"
#include "db/execution/vec_search_executor.hpp"
#include "db/index/index.hpp"
#include "db/table_segment_mvp.hpp"
#include "db/vector.hpp"
#include "query/expr/expr.hpp"
#include "utils/common_util.hpp"

using namespace vectordb;
using namespace vectordb::engine;

namespace {

std::atomic<uint64_t> g_dist_count{0};
DenseVecDistFunc<float> g_inner = nullptr;
float CountingDist(const void* a, const void* b, const void* p) {
  g_dist_count.fetch_add(1, std::memory_order_relaxed);
  return g_inner(a, b, p);
}

struct RefCtx {
  meta::TableSchema schema;
  std::unique_ptr<TableSegmentMVP> seg;
  std::shared_ptr<ANNGraphSegment> ann;
  std::unordered_map<std::string, meta::FieldType> field_map;
  meta::MetricType metric;
  size_t dim = 0;  // dist_func_param points here (table_mvp.cpp:82)
  std::vector<std::shared_ptr<execution::VecSearchExecutor>> execs;
  bool counting = false;
};

}  // namespace

extern "C" {

// metric: 1 L2, 2 COSINE, 3 IP (meta::MetricType).  attr_types: meta::FieldType ints.
void* ref_create(int metric, int64_t dim, int64_t capacity, int n_attr, const int* attr_types,
                 const char* const* attr_names) {
  auto* c = new RefCtx();
  c->metric = static_cast<meta::MetricType>(metric);
  c->dim = static_cast<size_t>(dim);
  c->schema.id_ = 0;
  c->schema.name_ = "t";
  int64_t fid = 0;
  for (int i = 0; i < n_attr; ++i) {
    meta::FieldSchema f;
    f.id_ = fid++;
    f.name_ = attr_names[i];
    f.field_type_ = static_cast<meta::FieldType>(attr_types[i]);
    f.is_primary_key_ = false;
    c->schema.fields_.push_back(f);
    c->field_map[f.name_] = f.field_type_;
  }
  meta::FieldSchema v;
  v.id_ = fid++;
  v.name_ = "Vec";
  v.field_type_ = meta::FieldType::VECTOR_FLOAT;
  v.vector_dimension_ = dim;
  v.metric_type_ = c->metric;
  c->schema.fields_.push_back(v);
  c->field_map["@distance"] = meta::FieldType::DOUBLE;
  c->seg.reset(new TableSegmentMVP(c->schema, capacity, nullptr));
  c->ann = std::make_shared<ANNGraphSegment>(true);
  return c;
}

void ref_destroy(void* h) { delete static_cast<RefCtx*>(h); }

float* ref_vectors(void* h) { return static_cast<RefCtx*>(h)->seg->vector_tables_[0]; }
char* ref_attrs(void* h) { return static_cast<RefCtx*>(h)->seg->attribute_table_; }
int64_t ref_attr_stride(void* h) { return static_cast<RefCtx*>(h)->seg->primitive_offset_; }
int64_t ref_attr_offset(void* h, const char* name) {
  auto& m = static_cast<RefCtx*>(h)->seg->field_name_mem_offset_map_;
  auto it = m.find(name);
  return it == m.end() ? -1 : static_cast<int64_t>(it->second);
}
// String attribute of one row (TableSegmentMVP::var_len_attr_table_[column][row], table_segment_mvp.hpp:82).
int ref_set_string(void* h, const char* name, int64_t row, const char* value) {
  auto* c = static_cast<RefCtx*>(h);
  auto it = c->seg->field_name_mem_offset_map_.find(name);
  if (it == c->seg->field_name_mem_offset_map_.end() || it->second >= c->seg->var_len_attr_table_.size()) return -1;
  c->seg->var_len_attr_table_[it->second][row] = std::string(value);
  return 0;
}
void ref_set_rows(void* h, int64_t n) { static_cast<RefCtx*>(h)->seg->record_number_ = n; }
void ref_set_deleted(void* h, int64_t id, int flag) {
  auto* c = static_cast<RefCtx*>(h);
  if (flag) c->seg->deleted_->set(id); else c->seg->deleted_->clear(id);
}

// ANNGraphSegment::BuildFromVectorTable over rows [0, n) (ann_graph_segment.cpp:201).
int ref_build(void* h, int64_t n, int threads) {
  auto* c = static_cast<RefCtx*>(h);
  omp_set_num_threads(threads);
  c->ann = std::make_shared<ANNGraphSegment>(true);
  try {
    c->ann->BuildFromVectorTable(c->seg->vector_tables_[0], n, static_cast<int64_t>(c->dim), c->metric);
  } catch (...) {
    return -1;
  }
  return 0;
}

// Hand an externally built CSR (e.g. GPU-built) to the reference executor's pointer ctor.
int ref_set_graph(void* h, int64_t n_indexed, const int64_t* offsets, const int64_t* nbrs, int64_t nav) {
  auto* c = static_cast<RefCtx*>(h);
  c->ann = std::make_shared<ANNGraphSegment>(true);
  c->ann->record_number_ = n_indexed;
  delete[] c->ann->offset_table_;
  delete[] c->ann->neighbor_list_;
  c->ann->offset_table_ = new int64_t[n_indexed + 1];
  std::memcpy(c->ann->offset_table_, offsets, sizeof(int64_t) * (n_indexed + 1));
  int64_t e = n_indexed > 0 ? offsets[n_indexed] : 0;
  c->ann->neighbor_list_ = new int64_t[e > 0 ? e : 1];
  if (e > 0) std::memcpy(c->ann->neighbor_list_, nbrs, sizeof(int64_t) * e);
  c->ann->navigation_point_ = nav;
  return 0;
}

int64_t ref_graph(void* h, const int64_t** offsets, const int64_t** nbrs, int64_t* nav) {
  auto* c = static_cast<RefCtx*>(h);
  *offsets = c->ann->offset_table_;
  *nbrs = c->ann->neighbor_list_;
  *nav = c->ann->navigation_point_;
  return c->ann->record_number_;
}

// Build n_exec executors the way TableMVP does (table_mvp.cpp:72-89).
int ref_make_executors(void* h, int n_exec, int T, int64_t L_master, int64_t L_local, int64_t iters,
                       int prefilter, int counting) {
  auto* c = static_cast<RefCtx*>(h);
  c->execs.clear();
  c->counting = counting != 0;
  DistFunc f = GetDistFunc(meta::FieldType::VECTOR_FLOAT, c->metric);
  if (c->counting) {
    g_inner = std::get<DenseVecDistFunc<float>>(f);
    f = static_cast<DenseVecDistFunc<float>>(CountingDist);
  }
  for (int i = 0; i < n_exec; ++i) {
    c->execs.push_back(std::make_shared<execution::VecSearchExecutor>(
        static_cast<int64_t>(c->dim), c->ann->navigation_point_, c->ann, c->ann->offset_table_,
        c->ann->neighbor_list_, c->seg->vector_tables_[0], f, &c->dim, T, L_master, L_local, iters,
        prefilter != 0));
  }
  return 0;
}

// One VecSearchExecutor::Search call on executor `e` (vec_search_executor.cpp:833).
// Returns result_size, or -1 when the filter fails to parse.
int64_t ref_search(void* h, int e, const float* query, int64_t limit, const char* filter, int64_t* ids,
                   double* dists, uint64_t* n_dist) {
  auto* c = static_cast<RefCtx*>(h);
  std::vector<query::expr::ExprNodePtr> nodes;
  if (filter && filter[0]) {
    auto st = query::expr::Expr::ParseNodeFromStr(filter, nodes, c->field_map);
    if (!st.ok()) return -1;
  }
  auto& ex = c->execs[e];
  uint64_t before = g_dist_count.load();
  int64_t rs = 0;
  auto status = ex->Search(const_cast<float*>(query), c->seg.get(), static_cast<size_t>(limit), nodes, rs);
  if (!status.ok()) return -2;  // the reference always returns OK (:934); a replaced Search may refuse a filter
  if (n_dist) *n_dist = g_dist_count.load() - before;
  for (int64_t i = 0; i < rs; ++i) {
    ids[i] = ex->search_result_[i];
    dists[i] = ex->distance_[i];
  }
  return rs;
}

// nq queries dealt to the executors by std::threads (one thread per executor), mimicking
// concurrent single-query REST calls.  ids/dists are [nq x limit]; counts [nq].
int ref_search_batch(void* h, const float* queries, int64_t nq, int64_t limit, const char* filter,
                     int64_t* ids, double* dists, int64_t* counts) {
  auto* c = static_cast<RefCtx*>(h);
  std::vector<query::expr::ExprNodePtr> nodes0;
  if (filter && filter[0]) {
    auto st = query::expr::Expr::ParseNodeFromStr(filter, nodes0, c->field_map);
    if (!st.ok()) return -1;
  }
  std::atomic<int64_t> next{0};
  std::atomic<bool> failed{false};
  int ne = static_cast<int>(c->execs.size());
  auto worker = [&](int e) {
    auto nodes = nodes0;
    auto& ex = c->execs[e];
    for (;;) {
      int64_t q = next.fetch_add(1);
      if (q >= nq) break;
      int64_t rs = 0;
      auto status = ex->Search(const_cast<float*>(queries + q * c->dim), c->seg.get(), static_cast<size_t>(limit), nodes, rs);
      if (!status.ok()) { failed.store(true); rs = 0; }
      counts[q] = rs;
      for (int64_t i = 0; i < rs && i < limit; ++i) {
        ids[q * limit + i] = ex->search_result_[i];
        dists[q * limit + i] = ex->distance_[i];
      }
    }
  };
  std::vector<std::thread> th;
  for (int e = 1; e < ne; ++e) th.emplace_back(worker, e);
  worker(0);
  for (auto& t : th) t.join();
  return failed.load() ? -2 : 0;
}

// ann_graph_<field>.bin round trip through the reference's own writer / loader
// (ANNGraphSegment::SaveANNGraph db/ann_graph_segment.cpp:156-199, loading ctor :39-98).
int ref_save_graph(void* h, const char* dir, int64_t table_id, int64_t field_id) {
  auto* c = static_cast<RefCtx*>(h);
  server::CommonUtil::CreateDirectory(std::string(dir) + "/" + std::to_string(table_id));
  auto st = c->ann->SaveANNGraph(dir, table_id, field_id, true);
  return st.ok() ? 0 : -1;
}
int ref_load_graph(void* h, const char* dir, int64_t table_id, int64_t field_id) {
  auto* c = static_cast<RefCtx*>(h);
  try {
    c->ann = std::make_shared<ANNGraphSegment>(std::string(dir), table_id, field_id);
  } catch (...) {
    return -1;
  }
  return 0;
}

float ref_distance(int metric, const float* a, const float* b, int64_t dim) {
  size_t d = static_cast<size_t>(dim);
  DistFunc f = GetDistFunc(meta::FieldType::VECTOR_FLOAT, static_cast<meta::MetricType>(metric));
  return std::get<DenseVecDistFunc<float>>(f)(a, b, &d);
}

void ref_normalize(float* v, int64_t dim) { Normalize(v, static_cast<size_t>(dim)); }

// Evaluate a parsed filter on row `id` with distance `dist` (expr_evaluator.cpp:170).
// Returns 0/1, or -1 on parse failure.
int ref_filter_eval(void* h, const char* filter, int64_t id, double dist) {
  auto* c = static_cast<RefCtx*>(h);
  std::vector<query::expr::ExprNodePtr> nodes;
  if (filter && filter[0]) {
    auto st = query::expr::Expr::ParseNodeFromStr(filter, nodes, c->field_map);
    if (!st.ok()) return -1;
  }
  query::expr::ExprEvaluator ev(nodes, c->seg->field_name_mem_offset_map_, c->seg->primitive_offset_,
                                c->seg->var_len_attr_num_, c->seg->attribute_table_, c->seg->var_len_attr_table_);
  return ev.LogicalEvaluate(static_cast<int>(nodes.size()) - 1, id, dist) ? 1 : 0;
}

// Dump the parsed node array as PODs so tests can feed the same program to the CUDA path.
// Layout per node (8 x int64/double slots = 64 B): node_type, value_type, left, right, int_value,
// double_value(bits), bool_value, field_offset (-1 if none, -2 for @distance).
int64_t ref_filter_nodes(void* h, const char* filter, int64_t* out, int64_t max_nodes) {
  auto* c = static_cast<RefCtx*>(h);
  std::vector<query::expr::ExprNodePtr> nodes;
  if (filter && filter[0]) {
    auto st = query::expr::Expr::ParseNodeFromStr(filter, nodes, c->field_map);
    if (!st.ok()) return -1;
  }
  int64_t n = static_cast<int64_t>(nodes.size());
  if (n > max_nodes) return -2;
  for (int64_t i = 0; i < n; ++i) {
    auto& nd = nodes[i];
    int64_t* o = out + i * 8;
    o[0] = static_cast<int64_t>(nd->node_type);
    o[1] = static_cast<int64_t>(nd->value_type);
    o[2] = static_cast<int64_t>(nd->left);
    o[3] = static_cast<int64_t>(nd->right);
    o[4] = nd->int_value;
    std::memcpy(&o[5], &nd->double_value, 8);
    o[6] = nd->bool_value ? 1 : 0;
    int64_t off = -1;
    if (!nd->field_name.empty()) {
      if (nd->field_name == "@distance") off = -2;
      else {
        auto it = c->seg->field_name_mem_offset_map_.find(nd->field_name);
        off = it == c->seg->field_name_mem_offset_map_.end() ? -1 : static_cast<int64_t>(it->second);
      }
    }
    o[7] = off;
  }
  return n;
}

// Value expression (group-by key / aggregate input: Expr::ParseNodeFromStr(expr, nodes, map, false),
// db_server.cpp:406,438) dumped as the same PODs as ref_filter_nodes; *root_value_type = ValueType of the root.
int64_t ref_value_nodes(void* h, const char* expr, int64_t* out, int64_t max_nodes, int64_t* root_value_type) {
  auto* c = static_cast<RefCtx*>(h);
  std::vector<query::expr::ExprNodePtr> nodes;
  auto st = query::expr::Expr::ParseNodeFromStr(expr, nodes, c->field_map, false);
  if (!st.ok() || nodes.empty()) return -1;
  int64_t n = static_cast<int64_t>(nodes.size());
  if (n > max_nodes) return -2;
  for (int64_t i = 0; i < n; ++i) {
    auto& nd = nodes[i];
    int64_t* o = out + i * 8;
    o[0] = static_cast<int64_t>(nd->node_type);
    o[1] = static_cast<int64_t>(nd->value_type);
    o[2] = static_cast<int64_t>(nd->left);
    o[3] = static_cast<int64_t>(nd->right);
    o[4] = nd->int_value;
    std::memcpy(&o[5], &nd->double_value, 8);
    o[6] = nd->bool_value ? 1 : 0;
    int64_t off = -1;
    if (!nd->field_name.empty()) {
      if (nd->field_name == "@distance") off = -2;
      else {
        auto it = c->seg->field_name_mem_offset_map_.find(nd->field_name);
        off = it == c->seg->field_name_mem_offset_map_.end() ? -1 : static_cast<int64_t>(it->second);
      }
    }
    o[7] = off;
  }
  *root_value_type = static_cast<int64_t>(nodes[n - 1]->value_type);
  return n;
}

// FacetExecutor (db/execution/aggregation.hpp:124-413) over one id list, set up the way db_server.cpp:384-456 does:
// one group-by expression ("" = global group, expression "1"), aggregates "SUM(expr)" / "MIN(..)" / "MAX(..)" /
// "COUNT(..)".  Writes the Project() JSON array into out (NUL-terminated); returns its length, -1 on a parse error.
int64_t ref_facet(void* h, const char* group_expr, int n_aggs, const char* const* agg_exprs, const int64_t* ids,
                  const double* dists, int64_t n, int has_distance, char* out, int64_t cap) {
  auto* c = static_cast<RefCtx*>(h);
  const bool global = group_expr == nullptr || group_expr[0] == 0;
  std::vector<std::string> group_exprs{global ? std::string("1") : std::string(group_expr)};
  std::vector<std::vector<query::expr::ExprNodePtr>> group_nodes(1);
  if (!query::expr::Expr::ParseNodeFromStr(group_exprs[0], group_nodes[0], c->field_map, false).ok()) return -1;
  std::vector<std::string> aggs;
  std::vector<query::expr::NodeType> types;
  std::vector<std::vector<query::expr::ExprNodePtr>> agg_nodes;
  for (int i = 0; i < n_aggs; ++i) {
    std::string e = agg_exprs[i], up = e, inner;
    for (auto& ch : up) ch = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
    query::expr::NodeType t;
    if (up.rfind("SUM(", 0) == 0 && up.back() == ')') { t = query::expr::NodeType::SumAggregation; inner = e.substr(4, e.size() - 5); }
    else if (up.rfind("MAX(", 0) == 0 && up.back() == ')') { t = query::expr::NodeType::MaxAggregation; inner = e.substr(4, e.size() - 5); }
    else if (up.rfind("MIN(", 0) == 0 && up.back() == ')') { t = query::expr::NodeType::MinAggregation; inner = e.substr(4, e.size() - 5); }
    else if (up.rfind("COUNT(", 0) == 0 && up.back() == ')') { t = query::expr::NodeType::CountAggregation; inner = "1"; }
    else return -1;
    std::vector<query::expr::ExprNodePtr> nodes;
    if (!query::expr::Expr::ParseNodeFromStr(inner, nodes, c->field_map, false).ok()) return -1;
    aggs.push_back(e);
    types.push_back(t);
    agg_nodes.push_back(nodes);
  }
  execution::FacetExecutor fx(global, group_exprs, group_nodes, types, aggs, agg_nodes);
  std::vector<int64_t> idv(ids, ids + n);
  std::vector<double> dv;
  if (has_distance) dv.assign(dists, dists + n);
  fx.Aggregate(c->seg.get(), n, idv, has_distance != 0, dv);
  vectordb::Json result;
  fx.Project(result);
  std::string text = result.DumpToString();
  if (static_cast<int64_t>(text.size()) + 1 > cap) return -2;
  std::memcpy(out, text.c_str(), text.size() + 1);
  return static_cast<int64_t>(text.size());
}

}  // extern "C"
