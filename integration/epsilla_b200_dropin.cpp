// Reference-side binding: the drop-in for Epsilla's vector-search hot path.
//
// This translation unit is compiled against the reference's UNMODIFIED headers and provides strong
// definitions of exactly three reference symbols, which replace the engine's own (weakened at link time,
// see integration/Makefile) — no reference source file is edited:
//
//   vectordb::engine::execution::VecSearchExecutor::VecSearchExecutor(...)   engine/db/execution/vec_search_executor.cpp:29-73
//   vectordb::engine::execution::VecSearchExecutor::Search(...)              engine/db/execution/vec_search_executor.cpp:833-935
//   vectordb::engine::ANNGraphSegment::BuildFromVectorTable(...)             engine/db/ann_graph_segment.cpp:201-242
//
// Everything else of the class (SearchByAttribute, the public result arrays search_result_ / distance_ /
// dimension_ that TableMVP::Search reads after the call — engine/db/table_mvp.cpp:365-394) stays the
// reference's.  TableMVP, ExecutorPool, DBServer and the filter parser therefore run unchanged; the dense
// vector search and the graph build run on the B200 through include/epsilla_b200.h.
//
// State: one device mirror (eps_index) per (vector table, graph, executor parameters), shared by the
// NumExecutorPerField executors TableMVP creates for a field (engine/db/table_mvp.cpp:72-89) and destroyed
// with the last of them — ownership rides on the executor's own ann_index_ shared_ptr (aliasing
// constructor), so the reference's inline destructor needs no hook.  Calls on a mirror are serialised by a
// mutex; concurrent unfiltered single-query calls from the executors of a pool are COALESCED into one batched
// launch by a leader/follower batch former — the reference API stays one query per call.
//
// Scope limits, reported as a non-OK Status instead of silently computing on the CPU: sparse-vector fields
// and string / IN / LIKE / NEARBY filter nodes.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "db/ann_graph_segment.hpp"
#include "db/execution/vec_search_executor.hpp"
#include "db/index/index.hpp"
#include "epsilla_b200.h"

namespace vectordb {
namespace engine {

namespace b200 {

struct Mirror {
  std::shared_ptr<ANNGraphSegment> ann;  // keeps the graph alive, like the reference member does
  eps_index* ix = nullptr;
  std::mutex mu;
  int metric = EPS_METRIC_L2;
  int64_t dim = 0;
  float* host_vectors = nullptr;
  int64_t capacity = 0;
  int64_t attr_rows = 0;
  const char* attr_ptr = nullptr;
  int64_t L_master = 500, L_local = 500;
  bool prefilter = false;
  int64_t nav = 0;
  int64_t* offsets = nullptr;
  int64_t* nbrs = nullptr;
  // Coalescing of concurrent single-query calls: unfiltered searches that arrive while
  // another executor's call is in flight are gathered and served by ONE batched launch.
  struct Pending {
    const float* query;
    size_t limit;
    int64_t* ids;
    double* dists;
    int64_t count = 0;
    int rc = EPS_OK;
    bool done = false;
  };
  std::mutex qmu;
  std::condition_variable qcv;
  std::vector<Pending*> waiting;
  bool leader_active = false;
  ~Mirror() {
    if (ix) eps_index_destroy(ix);
  }
};

using Key = std::tuple<const float*, const ANNGraphSegment*, int64_t, int64_t, bool, int>;
static std::mutex g_mu;
static std::map<Key, std::weak_ptr<Mirror>> g_mirrors;

static int MetricOf(const DistFunc& f) {
  if (!std::holds_alternative<DenseVecDistFunc<float>>(f)) return -1;
  auto fp = std::get<DenseVecDistFunc<float>>(f);
  auto is = [&](meta::MetricType m) {
    return fp == std::get<DenseVecDistFunc<float>>(GetDistFunc(meta::FieldType::VECTOR_FLOAT, m));
  };
  if (is(meta::MetricType::EUCLIDEAN)) return EPS_METRIC_L2;
  if (is(meta::MetricType::COSINE)) return EPS_METRIC_COSINE;
  if (is(meta::MetricType::DOT_PRODUCT)) return EPS_METRIC_IP;
  return -1;  // a DistFunc the library does not know (e.g. a test wrapper)
}

static Status Fail(const char* what) {
  return Status(DB_UNEXPECTED_ERROR, std::string("epsilla_b200: ") + what + ": " + eps_last_error());
}

// Leader/follower batch former.  The first caller becomes the leader: it waits a short window for followers,
// then serves every queued request with the same limit in one eps_search_batch call (other limits in further
// calls), copies each result out and wakes the followers.  Requests keep their per-call semantics (one query,
// own result arrays); only the launch is shared.
static int SearchCoalesced(Mirror* m, const float* query, size_t limit, int64_t* ids, double* dists, int64_t* count) {
  Mirror::Pending me;
  me.query = query; me.limit = limit; me.ids = ids; me.dists = dists;
  std::unique_lock<std::mutex> q(m->qmu);
  m->waiting.push_back(&me);
  if (m->leader_active) {
    m->qcv.wait(q, [&] { return me.done || !m->leader_active; });
    if (me.done) { *count = me.count; return me.rc; }
    // the leader left before taking this request: fall through and lead
  }
  m->leader_active = true;
  while (!me.done) {
    m->qcv.wait_for(q, std::chrono::microseconds(100));  // batching window
    std::vector<Mirror::Pending*> batch;
    std::vector<Mirror::Pending*> rest;
    const size_t lim = m->waiting.front()->limit;
    for (auto* p : m->waiting) (p->limit == lim ? batch : rest).push_back(p);
    m->waiting.swap(rest);
    q.unlock();
    const int64_t nq = static_cast<int64_t>(batch.size());
    std::vector<float> qbuf(static_cast<size_t>(nq) * m->dim);
    for (int64_t i = 0; i < nq; ++i) std::memcpy(qbuf.data() + i * m->dim, batch[i]->query, sizeof(float) * m->dim);
    std::vector<int64_t> oi(static_cast<size_t>(nq) * lim), oc(static_cast<size_t>(nq));
    std::vector<double> od(static_cast<size_t>(nq) * lim);
    int rc;
    {
      std::lock_guard<std::mutex> lk(m->mu);  // the index itself is single-threaded
      rc = eps_search_batch(m->ix, qbuf.data(), nq, static_cast<int64_t>(lim), nullptr, 0, oi.data(), od.data(), oc.data(), nullptr);
    }
    q.lock();
    for (int64_t i = 0; i < nq; ++i) {
      auto* p = batch[i];
      p->rc = rc;
      if (rc == EPS_OK) {
        p->count = oc[i];
        std::memcpy(p->ids, oi.data() + i * lim, sizeof(int64_t) * lim);
        std::memcpy(p->dists, od.data() + i * lim, sizeof(double) * lim);
      }
      p->done = true;
    }
    m->qcv.notify_all();
  }
  m->leader_active = false;
  m->qcv.notify_all();  // a queued follower (if any) takes over as leader
  *count = me.count;
  return me.rc;
}

}  // namespace b200

namespace execution {

// Same signature and member initialisation as the reference constructor; the per-query CPU scratch
// (is_visited_, set_L_, brute_force_queue_) is left empty because the search runs on the device.
VecSearchExecutor::VecSearchExecutor(const int64_t dimension, const int64_t start_search_point,
                                     std::shared_ptr<ANNGraphSegment> ann_index, int64_t* offset_table,
                                     int64_t* neighbor_list,
                                     std::variant<DenseVectorColumnDataContainer, VariableLenAttrColumnContainer*> vector_column,
                                     DistFunc fstdistfunc, void* dist_func_param, int num_threads, int64_t L_master,
                                     int64_t L_local, int64_t subsearch_iterations, bool prefilter_enabled)
    : total_indexed_vector_(ann_index->record_number_),
      dimension_(dimension),
      start_search_point_(start_search_point),
      offset_table_(offset_table),
      neighbor_list_(neighbor_list),
      vector_column_(vector_column),
      fstdistfunc_(fstdistfunc),
      dist_func_param_(dist_func_param),
      num_threads_(num_threads),
      L_master_(L_master),
      L_local_(L_local),
      subsearch_iterations_(subsearch_iterations),
      search_result_(L_master),
      distance_(L_master),
      init_ids_(1, 0),
      local_queues_sizes_(num_threads, 0),
      local_queues_starts_(num_threads),
      brute_force_search_(ann_index->record_number_ < BruteforceThreshold),
      prefilter_enabled_(prefilter_enabled) {
  const int metric = b200::MetricOf(fstdistfunc);
  if (metric < 0 || !std::holds_alternative<DenseVectorColumnDataContainer>(vector_column)) {
    ann_index_ = ann_index;  // sparse / unknown metric: Search() reports NOT_IMPLEMENTED
    return;
  }
  float* table = std::get<DenseVectorColumnDataContainer>(vector_column);
  b200::Key key(table, ann_index.get(), L_master, L_local, prefilter_enabled, metric);
  std::shared_ptr<b200::Mirror> m;
  {
    std::lock_guard<std::mutex> lk(b200::g_mu);
    auto it = b200::g_mirrors.find(key);
    if (it != b200::g_mirrors.end()) m = it->second.lock();
    if (!m) {
      m = std::make_shared<b200::Mirror>();
      m->ann = ann_index;
      m->metric = metric;
      m->dim = dimension;
      m->host_vectors = table;
      m->L_master = L_master;
      m->L_local = L_local;
      m->prefilter = prefilter_enabled;
      m->nav = start_search_point;
      m->offsets = offset_table;
      m->nbrs = neighbor_list;
      b200::g_mirrors[key] = m;
    }
  }
  // the executor's own ann_index_ owns the mirror (aliasing constructor) and still points at the graph
  ann_index_ = std::shared_ptr<ANNGraphSegment>(m, ann_index.get());
  init_ids_[0] = reinterpret_cast<int64_t>(m.get());
}

Status VecSearchExecutor::Search(const VectorPtr query_data, vectordb::engine::TableSegmentMVP* table_segment,
                                 const size_t limit, std::vector<vectordb::query::expr::ExprNodePtr>& filter_nodes,
                                 int64_t& result_size) {
  result_size = 0;
  auto* m = reinterpret_cast<b200::Mirror*>(init_ids_[0]);
  if (m == nullptr || !std::holds_alternative<DenseVectorPtr>(query_data))
    return Status(NOT_IMPLEMENTED_ERROR, "epsilla_b200: sparse-vector search is out of scope of the GPU path");
  std::unique_lock<std::mutex> lk(m->mu);
  const int64_t total = table_segment->record_number_;  // snapshot (:839)
  if (m->ix == nullptr) {
    m->capacity = static_cast<int64_t>(table_segment->size_limit_);
    if (eps_index_create(&m->ix, m->metric, m->dim, m->host_vectors, m->capacity, 0) != EPS_OK) return b200::Fail("create");
    if (eps_index_sync_rows(m->ix, std::max<int64_t>(total, total_indexed_vector_)) != EPS_OK) return b200::Fail("sync_rows");
    if (total_indexed_vector_ > 0 &&
        eps_index_set_graph(m->ix, total_indexed_vector_, m->offsets, m->nbrs, m->nav) != EPS_OK)
      return b200::Fail("set_graph");
    if (eps_index_config(m->ix, m->L_master, m->L_local, m->prefilter ? 1 : 0, 0) != EPS_OK) return b200::Fail("config");
  }
  if (eps_index_sync_rows(m->ix, total) != EPS_OK) return b200::Fail("sync_rows");
  ConcurrentBitset& deleted = *(table_segment->deleted_);  // (:840)
  if (eps_index_set_deleted(m->ix, deleted.data(), (total + 7) / 8) != EPS_OK) return b200::Fail("set_deleted");

  // filter nodes -> PODs, field names resolved through the segment's offset map (:841-848)
  std::vector<eps_filter_node> nodes(filter_nodes.size());
  for (size_t i = 0; i < filter_nodes.size(); ++i) {
    const auto& s = *filter_nodes[i];
    eps_filter_node& d = nodes[i];
    d.node_type = static_cast<int64_t>(s.node_type);
    d.value_type = static_cast<int64_t>(s.value_type);
    d.left = static_cast<int64_t>(s.left);
    d.right = static_cast<int64_t>(s.right);
    d.int_value = s.int_value;
    d.double_value = s.double_value;
    d.bool_value = s.bool_value ? 1 : 0;
    d.field_offset = -1;
    if (!s.field_name.empty()) {
      if (s.field_name == "@distance") d.field_offset = -2;
      else {
        auto it = table_segment->field_name_mem_offset_map_.find(s.field_name);
        if (it != table_segment->field_name_mem_offset_map_.end()) d.field_offset = static_cast<int64_t>(it->second);
      }
    }
  }
  if (!nodes.empty() && (m->attr_rows != total || m->attr_ptr != table_segment->attribute_table_)) {
    if (eps_index_set_attrs(m->ix, table_segment->attribute_table_, table_segment->primitive_offset_, total) != EPS_OK)
      return b200::Fail("set_attrs");
    m->attr_rows = total;
    m->attr_ptr = table_segment->attribute_table_;
  }
  if (search_result_.size() < limit) {  // the reference overruns here when limit > L_master
    search_result_.resize(limit);
    distance_.resize(limit);
  }
  int64_t count = 0;
  int rc;
  if (nodes.empty()) {
    // hand the request to the mirror's batch former: release the mirror lock while queued so that the other
    // executors of the pool (engine/db/execution/executor_pool.hpp) can join the same batch
    lk.unlock();
    rc = b200::SearchCoalesced(m, std::get<DenseVectorPtr>(query_data), limit, search_result_.data(), distance_.data(), &count);
    lk.lock();
  } else {
    rc = eps_search_batch(m->ix, std::get<DenseVectorPtr>(query_data), 1, static_cast<int64_t>(limit), nodes.data(),
                          static_cast<int64_t>(nodes.size()), search_result_.data(), distance_.data(), &count, nullptr);
  }
  if (rc == EPS_ERR_UNSUPPORTED) return Status(NOT_IMPLEMENTED_ERROR, std::string("epsilla_b200: ") + eps_last_error());
  if (rc != EPS_OK) return b200::Fail("search");
  result_size = count;
  return Status::OK();  // (:934)
}

}  // namespace execution

// Graph build on the device, emitting the reference's own CSR members so SaveANNGraph / the loader
// (engine/db/ann_graph_segment.cpp:39-98,156-199) and every executor keep working unchanged.
void ANNGraphSegment::BuildFromVectorTable(VectorColumnData vector_column, int64_t n, int64_t dim,
                                           meta::MetricType metricType) {
  if (!std::holds_alternative<DenseVectorColumnDataContainer>(vector_column))
    throw std::runtime_error("epsilla_b200: sparse-vector graph build is out of scope of the GPU path");
  int metric = EPS_METRIC_L2;
  if (metricType == meta::MetricType::COSINE) metric = EPS_METRIC_COSINE;
  if (metricType == meta::MetricType::DOT_PRODUCT) metric = EPS_METRIC_IP;
  eps_index* ix = nullptr;
  auto die = [&](const char* what) {
    std::string msg = std::string("epsilla_b200 build: ") + what + ": " + eps_last_error();
    if (ix) eps_index_destroy(ix);
    throw std::runtime_error(msg);
  };
  if (eps_index_create(&ix, metric, dim, std::get<DenseVectorColumnDataContainer>(vector_column), n, 0) != EPS_OK) die("create");
  if (eps_index_sync_rows(ix, n) != EPS_OK) die("sync_rows");
  if (eps_index_build(ix, n, nullptr) != EPS_OK) die("build");
  int64_t ni = 0, ne = 0, nav = 0;
  if (eps_index_get_graph(ix, &ni, &ne, nullptr, nullptr, &nav) != EPS_OK) die("get_graph");
  int64_t* off = new int64_t[ni + 1];
  int64_t* nb = new int64_t[ne > 0 ? ne : 1];
  if (eps_index_get_graph(ix, nullptr, nullptr, off, nb, nullptr) != EPS_OK) { delete[] off; delete[] nb; die("get_graph"); }
  eps_index_destroy(ix);
  record_number_ = n;
  if (offset_table_ != nullptr) delete[] offset_table_;
  if (neighbor_list_ != nullptr) delete[] neighbor_list_;
  offset_table_ = off;
  neighbor_list_ = nb;
  navigation_point_ = nav;
}

}  // namespace engine
}  // namespace vectordb
