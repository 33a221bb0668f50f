// K2 — batched best-first graph search.  SURVEY.md §8a rows A4-A8.
//
// What the reference computes (engine/db/execution/vec_search_executor.cpp, IntraQueryThreads = 1, the
// only configuration in which it is a pure function of its inputs — SURVEY.md §5/§8c):
//   InitializeSetLPara (:446-485)  seed the queue with the L query-independent init ids, mark them
//                                  visited, sort by (distance,id);
//   SearchImpl (:518-715)          repeatedly expand the first unchecked queue entry;
//   ExpandOneCandidate (:384-444)  for each CSR neighbour: skip if visited, mark, distance, reject if
//                                  dist > worst-in-queue, else AddIntoQueue (:75-117, sorted insert with
//                                  eviction); return the lowest insert position r;
//   k = (r <= k) ? r : k+1 (:648-652); stop when no unchecked entry is left.
// The queue after one expansion is the top-L by (distance,id) of {queue ∪ unvisited neighbours}, whatever
// the insertion order, and r is the final position of the smallest inserted entry; so evaluating all
// neighbour distances of a vertex in parallel and merging them at once is equivalent.
//
// Mapping: a persistent grid, one CTA per in-flight query (queries are claimed from an atomic counter),
// the sorted queue and the query vector in shared memory, one warp per neighbour row (coalesced 128-bit
// streaming loads, warp-shuffle reduction), a per-CTA visited bitmap in global memory (L2-resident),
// block-parallel rank-and-shift merge instead of the reference's memmove insert.
#include <cstdlib>

#include "internal.h"

namespace eps {

constexpr int kCH = 64;  // neighbours handled per merge round

struct GSArgs {
  const float* vectors;
  const int64_t* offsets;
  const int32_t* nbrs;
  const int32_t* init_ids;
  const float* queries;
  uint32_t* visited;            // [slots x visited_words]
  unsigned long long* out_queue;  // [nq x L]
  int* work_counter;
  unsigned long long* stats;    // n_dist, n_expand, n_edges, n_seed
  int64_t visited_words;
  int dim, metric, vec4;
  int L, Lp;
  int nq;
};

__device__ __forceinline__ int lb_masked(const unsigned long long* a, int n, unsigned long long key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((a[mid] & kKeyMask) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128, 6) graph_search_kernel(GSArgs a) {
  extern __shared__ __align__(16) unsigned char gs_smem[];
  unsigned long long* queue = reinterpret_cast<unsigned long long*>(gs_smem);          // [Lp]
  unsigned long long* cand = queue + a.Lp;                                             // [kCH] accepted, unsorted
  unsigned long long* cs = cand + kCH;                                                 // [kCH] accepted, sorted
  float* qv = reinterpret_cast<float*>(cs + kCH);                                      // [dim padded to 4]
  int* pos = reinterpret_cast<int*>(qv + ((a.dim + 3) & ~3));                          // [kCH]
  int* fresh = pos + kCH;                                                              // [kCH]
  __shared__ int s_q, s_cur, s_nfresh, s_nacc, s_pmin;
  __shared__ unsigned long long s_ndist, s_nexp, s_nedge;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  uint32_t* visited = a.visited + static_cast<int64_t>(blockIdx.x) * a.visited_words;
  const int L = a.L;
  if (tid == 0) { s_ndist = 0; s_nexp = 0; s_nedge = 0; }

  for (;;) {
    __syncthreads();
    if (tid == 0) s_q = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int q = s_q;
    if (q >= a.nq) break;
    for (int i = tid; i < a.dim; i += blockDim.x) qv[i] = a.queries[static_cast<int64_t>(q) * a.dim + i];
    for (int i = L + tid; i < a.Lp; i += blockDim.x) queue[i] = kKeyInf;
    if (tid == 0) { s_nfresh = 0; s_nacc = 0; }
    __syncthreads();

    // ---- seed (InitializeSetLPara) ----
    for (int i = tid; i < L; i += blockDim.x) {
      uint32_t id = static_cast<uint32_t>(a.init_ids[i]);
      atomicOr(&visited[id >> 5], 1u << (id & 31));
    }
    for (int i = warp; i < L; i += nwarps) {
      const int id = a.init_ids[i];
      float d = warp_distance(a.metric, a.vec4 != 0, a.vectors + static_cast<int64_t>(id) * a.dim, qv, a.dim, lane);
      if (lane == 0) queue[i] = make_key(d, static_cast<uint32_t>(id));
    }
    __syncthreads();
    block_bitonic_sort(queue, a.Lp);

    // ---- best-first loop (SearchImpl) ----
    int k = 0;
    for (;;) {
      if (warp == 0) {
        int found = -1;
        for (int p = k; p < L; p += 32) {
          int idx = p + lane;
          bool un = idx < L && !(queue[idx] & kCheckedBit);
          unsigned b = __ballot_sync(kFull, un);
          if (b) { found = p + __ffs(b) - 1; break; }
        }
        if (lane == 0) {
          s_cur = found;
          s_pmin = L;
          if (found >= 0) queue[found] |= kCheckedBit;
        }
      }
      __syncthreads();
      const int cur = s_cur;
      if (cur < 0) break;
      const int c = static_cast<int>(key_id(queue[cur]));
      const int64_t e0 = a.offsets[c], e1 = a.offsets[c + 1];
      __syncthreads();  // everyone holds cur before warp 0 may publish the next one
      if (tid == 0) { ++s_nexp; s_nedge += static_cast<unsigned long long>(e1 - e0); }

      for (int64_t eb = e0; eb < e1; eb += kCH) {
        const int cnt = static_cast<int>(min(static_cast<int64_t>(kCH), e1 - eb));
        // visited test-and-set (ExpandOneCandidate :403-406)
        if (tid < cnt) {
          const uint32_t nb = static_cast<uint32_t>(a.nbrs[eb + tid]);
          const uint32_t bit = 1u << (nb & 31);
          const uint32_t old = atomicOr(&visited[nb >> 5], bit);
          if (!(old & bit)) fresh[atomicAdd(&s_nfresh, 1)] = static_cast<int>(nb);
        }
        __syncthreads();
        const int nfresh = s_nfresh;
        const unsigned long long bound = queue[L - 1] & kKeyMask;  // live worst entry (:546)
        for (int i = warp; i < nfresh; i += nwarps) {
          const int nb = fresh[i];
          float d = warp_distance(a.metric, a.vec4 != 0, a.vectors + static_cast<int64_t>(nb) * a.dim, qv, a.dim, lane);
          if (lane == 0) {
            unsigned long long key = make_key(d, static_cast<uint32_t>(nb));
            if (key < bound) cand[atomicAdd(&s_nacc, 1)] = key;  // dist > bound rejected (:424); ties by id
          }
        }
        __syncthreads();
        const int m = s_nacc;
        if (m > 0) {
          // 1. sort the accepted candidates (rank by counting; keys are distinct)
          for (int i = tid; i < m; i += blockDim.x) {
            const unsigned long long key = cand[i];
            int r = 0;
            for (int j = 0; j < m; ++j) r += (cand[j] < key);
            cs[r] = key;
          }
          __syncthreads();
          // 2. insertion points in the current queue
          for (int i = tid; i < m; i += blockDim.x) pos[i] = lb_masked(queue, L, cs[i]);
          __syncthreads();
          const int p0 = pos[0];
          // 3. shift old entries right by the number of candidates that precede them, top tile first
          for (int hi = L; hi > p0; hi -= blockDim.x) {
            const int j = hi - 1 - tid;
            unsigned long long key = 0;
            int dest = L;
            if (j >= p0) {
              key = queue[j];
              int lo = 0, up = m;  // s = #{i : pos[i] <= j}
              while (lo < up) { int mid = (lo + up) >> 1; if (pos[mid] <= j) lo = mid + 1; else up = mid; }
              dest = j + lo;
            }
            __syncthreads();
            if (dest < L) queue[dest] = key;
            __syncthreads();
          }
          // 4. drop the candidates into their holes
          for (int i = tid; i < m; i += blockDim.x) {
            const int f = pos[i] + i;
            if (f < L) queue[f] = cs[i];
          }
          if (tid == 0 && p0 < s_pmin) s_pmin = p0;
        }
        // every thread has read m (the m > 0 path passed barriers since; m == 0 rewrites 0 with 0)
        if (tid == 0) { s_ndist += static_cast<unsigned long long>(nfresh); s_nfresh = 0; s_nacc = 0; }
        __syncthreads();  // counters reset + queue settled before the next round
      }
      const int pmin = s_pmin;
      k = (pmin <= k) ? pmin : k + 1;  // :648-652 (only warp 0 consumes k)
    }

    // ---- results + visited reset (:711-714) ----
    unsigned long long* out = a.out_queue + static_cast<int64_t>(q) * L;
    for (int i = tid; i < L; i += blockDim.x) out[i] = queue[i];
    {
      uint4* v4 = reinterpret_cast<uint4*>(visited);
      const int64_t n4 = a.visited_words >> 2;
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int64_t i = tid; i < n4; i += blockDim.x) v4[i] = z;
    }
    if (tid == 0) s_ndist += static_cast<unsigned long long>(L);
  }
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&a.stats[0], s_ndist);
    atomicAdd(&a.stats[1], s_nexp);
    atomicAdd(&a.stats[2], s_nedge);
  }
}

// ---------------------------------------------------------------------------------------------------------
// v2 of the kernel: same algorithm and results, restructured against the two stalls the ncu capture of v1
// showed (profiles/r01_ncu_graph_search_v1_*: barrier 8.4 and long_scoreboard 7.5 per issue, 16 % DRAM):
//   * fixed-stride adjacency (kEll ids per vertex, -1 padded) => the neighbour ids are ONE load away from the
//     vertex id (CSR needs offsets first); rows longer than kEll continue in the CSR (rare: repair hubs);
//   * the adjacency row of the NEXT likely candidate (second unchecked entry) is loaded speculatively by the
//     otherwise idle warps 2-3 while warps 0-1 run the visited test of the current one;
//   * the seed distances arrive precomputed as a dense [B x L] tile product (the seed set is query-independent,
//     SURVEY §8a A4) instead of L warp-per-row evaluations per query;
//   * the merge shifts in place in super-tiles of 8 keys per thread (2 barriers per 1024 keys instead of per 128);
//   * a lane issues all 128-bit loads of a row before the first use (one memory round trip per row).
// Three barriers per expansion when nothing is accepted (the common case late in a search), six otherwise.
// ---------------------------------------------------------------------------------------------------------
constexpr int kEll = 64;

struct GS2Args {
  const float* vectors;
  const int64_t* offsets;
  const int32_t* nbrs;
  const int32_t* ell;             // [n x kEll]
  const int32_t* init_ids;
  const float* seed_dist;         // [nq x seed_ld]
  const float* queries;
  uint32_t* visited;
  unsigned long long* out_queue;
  int* work_counter;
  unsigned long long* stats;
  int64_t visited_words;
  int64_t seed_ld;
  int dim, metric, vec4;
  int L, Lp;
  int nq;
};

__global__ void __launch_bounds__(128, 7) graph_search_kernel_v2(GS2Args a) {
  extern __shared__ __align__(16) unsigned char gs_smem[];
  unsigned long long* qa = reinterpret_cast<unsigned long long*>(gs_smem);  // [Lp] current queue
  unsigned long long* cand = qa + a.Lp;                                     // [kCH]
  unsigned long long* cs = cand + kCH;                                      // [kCH]
  float* qv = reinterpret_cast<float*>(cs + kCH);                           // [dim4]
  int* pos = reinterpret_cast<int*>(qv + ((a.dim + 3) & ~3));               // [kCH]
  int* fresh = pos + kCH;                                                   // [kCH]
  int* spec = fresh + kCH;                                                  // [2][kEll]
  __shared__ int s_q, s_cur, s_nfresh, s_nacc, s_pmin, s_more;
  __shared__ int s_spec_id[2];
  __shared__ unsigned long long s_ndist, s_nexp, s_nedge;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  uint32_t* visited = a.visited + static_cast<int64_t>(blockIdx.x) * a.visited_words;
  const int L = a.L;
  if (tid == 0) { s_ndist = 0; s_nexp = 0; s_nedge = 0; }

  for (;;) {
    __syncthreads();
    if (tid == 0) s_q = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int q = s_q;
    if (q >= a.nq) break;
    for (int i = tid; i < a.dim; i += blockDim.x) qv[i] = a.queries[static_cast<int64_t>(q) * a.dim + i];
    for (int i = tid; i < a.Lp; i += blockDim.x) {
      unsigned long long key = kKeyInf;
      if (i < L) {
        const uint32_t id = static_cast<uint32_t>(a.init_ids[i]);
        atomicOr(&visited[id >> 5], 1u << (id & 31));
        key = make_key(a.seed_dist[static_cast<int64_t>(q) * a.seed_ld + i], id);
      }
      qa[i] = key;
    }
    if (tid == 0) { s_nfresh = 0; s_nacc = 0; s_spec_id[0] = -1; s_spec_id[1] = -1; }
    __syncthreads();
    block_bitonic_sort(qa, a.Lp);

    int k = 0;
    for (int it = 0;; ++it) {
      const int par = it & 1;
      if (warp == 0) {
        int found = -1, next = -1;
        for (int p = k; p < L; p += 32) {
          const int idx = p + lane;
          const bool un = idx < L && !(qa[idx] & kCheckedBit);
          unsigned b = __ballot_sync(kFull, un);
          if (found < 0 && b) { found = p + __ffs(b) - 1; b &= b - 1; }
          if (found >= 0 && b) { next = p + __ffs(b) - 1; break; }
          if (found >= 0 && p >= found + 96) break;  // bounded look-ahead for the speculation
        }
        if (lane == 0) {
          s_cur = found;
          s_pmin = L;
          s_more = 0;
          if (found >= 0) qa[found] |= kCheckedBit;
          s_spec_id[par] = next >= 0 ? static_cast<int>(key_id(qa[next])) : -1;
        }
      }
      __syncthreads();  // (1)
      const int cur = s_cur;
      if (cur < 0) break;
      const int c = static_cast<int>(key_id(qa[cur]));
      const bool hit = s_spec_id[par ^ 1] == c;
      if (tid < kEll) {
        const int nb = hit ? spec[(par ^ 1) * kEll + tid] : __ldg(a.ell + static_cast<int64_t>(c) * kEll + tid);
        if (nb >= 0) {
          const uint32_t bit = 1u << (nb & 31);
          const uint32_t old = atomicOr(&visited[nb >> 5], bit);
          if (!(old & bit)) fresh[atomicAdd(&s_nfresh, 1)] = nb;
          if (tid == kEll - 1) s_more = 1;  // full row: it may continue in the CSR
        }
        const unsigned vb = __ballot_sync(kFull, nb >= 0);
        if (lane == 0 && vb) atomicAdd(&s_nedge, static_cast<unsigned long long>(__popc(vb)));
      } else {
        const int c2 = s_spec_id[par];
        if (c2 >= 0) spec[par * kEll + (tid - kEll)] = __ldg(a.ell + static_cast<int64_t>(c2) * kEll + (tid - kEll));
      }
      __syncthreads();  // (2)
      int64_t e_next = 0, e_end = 0;
      if (s_more) { e_next = a.offsets[c] + kEll; e_end = a.offsets[c + 1]; }
      if (tid == 0) { ++s_nexp; }
      for (;;) {
        const int nfresh = s_nfresh;
        const unsigned long long bound = qa[L - 1] & kKeyMask;
        for (int i = warp; i < nfresh; i += nwarps) {
          const int nb = fresh[i];
          float d = warp_distance(a.metric, a.vec4 != 0, a.vectors + static_cast<int64_t>(nb) * a.dim, qv, a.dim, lane);
          if (lane == 0) {
            const unsigned long long key = make_key(d, static_cast<uint32_t>(nb));
            if (key < bound) cand[atomicAdd(&s_nacc, 1)] = key;
          }
        }
        __syncthreads();  // (3)
        const int m = s_nacc;
        if (m > 0) {
          for (int i = tid; i < m; i += blockDim.x) {
            const unsigned long long key = cand[i];
            int r = 0;
            for (int j = 0; j < m; ++j) r += (cand[j] < key);
            cs[r] = key;
          }
          __syncthreads();
          for (int i = tid; i < m; i += blockDim.x) pos[i] = lb_masked(qa, L, cs[i]);
          __syncthreads();
          const int p0 = pos[0];
          // in-place shift of [p0, L): super-tiles of 8 keys per thread, top first; each key moves right by
          // the number of accepted candidates that precede it (counted linearly: m is small)
          for (int hi = L; hi > p0; hi -= 8 * 128) {
            unsigned long long kreg[8];
            int dreg[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = hi - 1 - (u * 128 + tid);
              dreg[u] = L;
              if (j >= p0) {
                kreg[u] = qa[j];
                int sft = 0;
                if (m <= 8) { for (int i = 0; i < m; ++i) sft += (pos[i] <= j); }
                else { int lo = 0, up = m; while (lo < up) { const int mid = (lo + up) >> 1; if (pos[mid] <= j) lo = mid + 1; else up = mid; } sft = lo; }
                dreg[u] = j + sft;
              }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 8; ++u) if (dreg[u] < L) qa[dreg[u]] = kreg[u];
            __syncthreads();
          }
          for (int i = tid; i < m; i += blockDim.x) {
            const int f = pos[i] + i;
            if (f < L) qa[f] = cs[i];
          }
          if (tid == 0 && p0 < s_pmin) s_pmin = p0;
        }
        if (tid == 0) { s_ndist += static_cast<unsigned long long>(nfresh); s_nfresh = 0; s_nacc = 0; }
        __syncthreads();  // queue + counters settled
        if (e_next >= e_end) break;
        // rare: the row continues in the CSR beyond its first kEll entries
        const int cnt = static_cast<int>(min(static_cast<int64_t>(kCH), e_end - e_next));
        if (tid < cnt) {
          const uint32_t nb = static_cast<uint32_t>(a.nbrs[e_next + tid]);
          const uint32_t bit = 1u << (nb & 31);
          const uint32_t old = atomicOr(&visited[nb >> 5], bit);
          if (!(old & bit)) fresh[atomicAdd(&s_nfresh, 1)] = static_cast<int>(nb);
        }
        e_next += cnt;
        __syncthreads();
      }
      const int pmin = s_pmin;
      k = (pmin <= k) ? pmin : k + 1;
    }

    unsigned long long* out = a.out_queue + static_cast<int64_t>(q) * L;
    for (int i = tid; i < L; i += blockDim.x) out[i] = qa[i];
    {
      uint4* v4 = reinterpret_cast<uint4*>(visited);
      const int64_t n4 = a.visited_words >> 2;
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int64_t i = tid; i < n4; i += blockDim.x) v4[i] = z;
    }
    if (tid == 0) s_ndist += static_cast<unsigned long long>(L);
  }
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&a.stats[0], s_ndist);
    atomicAdd(&a.stats[1], s_nexp);
    atomicAdd(&a.stats[2], s_nedge);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Wide variant: W unchecked candidates are expanded per iteration (their neighbour rows are tested, gathered
// and merged together).  This is the device analogue of the reference's IntraQueryThreads > 1 mode
// (vec_search_executor.cpp:601-698: M candidates dealt to workers, expanded against a slightly stale bound,
// merged back) — like that mode it is NOT bit-identical to the sequential order (W = 1, kernels above, is);
// it trades that for W-fold fewer dependent round trips per query.  Same queue / visited / bound rules.
// ---------------------------------------------------------------------------------------------------------
constexpr int kWideMax = 8;
constexpr int kWCap = kWideMax * kEll;  // fresh / candidate slots per iteration

__global__ void __launch_bounds__(128, 7) graph_search_kernel_wide(GS2Args a, int W) {
  extern __shared__ __align__(16) unsigned char gs_smem[];
  unsigned long long* qa = reinterpret_cast<unsigned long long*>(gs_smem);  // [Lp]
  unsigned long long* cand = qa + a.Lp;                                     // [kWCap]
  unsigned long long* cs = cand + kWCap;                                    // [kWCap]
  float* qv = reinterpret_cast<float*>(cs + kWCap);                         // [dim4]
  int* pos = reinterpret_cast<int*>(qv + ((a.dim + 3) & ~3));               // [kWCap]
  int* fresh = pos + kWCap;                                                 // [kWCap]
  __shared__ int s_q, s_ncur, s_first, s_nfresh, s_nacc, s_pmin, s_more;
  __shared__ int s_cid[kWideMax];
  __shared__ unsigned long long s_ndist, s_nexp, s_nedge;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  uint32_t* visited = a.visited + static_cast<int64_t>(blockIdx.x) * a.visited_words;
  const int L = a.L;
  if (tid == 0) { s_ndist = 0; s_nexp = 0; s_nedge = 0; }

  for (;;) {
    __syncthreads();
    if (tid == 0) s_q = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int q = s_q;
    if (q >= a.nq) break;
    for (int i = tid; i < a.dim; i += blockDim.x) qv[i] = a.queries[static_cast<int64_t>(q) * a.dim + i];
    for (int i = tid; i < a.Lp; i += blockDim.x) {
      unsigned long long key = kKeyInf;
      if (i < L) {
        const uint32_t id = static_cast<uint32_t>(a.init_ids[i]);
        atomicOr(&visited[id >> 5], 1u << (id & 31));
        key = make_key(a.seed_dist[static_cast<int64_t>(q) * a.seed_ld + i], id);
      }
      qa[i] = key;
    }
    if (tid == 0) { s_nfresh = 0; s_nacc = 0; }
    __syncthreads();
    block_bitonic_sort(qa, a.Lp);

    int k = 0;
    for (;;) {
      if (warp == 0) {
        int cnt = 0, first = -1;
        for (int p = k; p < L && cnt < W; p += 32) {
          const int idx = p + lane;
          const bool un = idx < L && !(qa[idx] & kCheckedBit);
          unsigned b = __ballot_sync(kFull, un);
          while (b && cnt < W) {
            const int at = p + __ffs(b) - 1;
            if (first < 0) first = at;
            if (lane == 0) { s_cid[cnt] = static_cast<int>(key_id(qa[at])); qa[at] |= kCheckedBit; }
            b &= b - 1;
            ++cnt;
          }
        }
        if (lane == 0) { s_ncur = cnt; s_first = first; s_pmin = L; s_more = 0; }
      }
      __syncthreads();  // (1)
      const int ncur = s_ncur;
      if (ncur == 0) break;
      for (int slot = tid; slot < ncur * kEll; slot += blockDim.x) {
        const int ci = slot / kEll, e = slot % kEll;
        const int nb = __ldg(a.ell + static_cast<int64_t>(s_cid[ci]) * kEll + e);
        if (nb >= 0) {
          const uint32_t bit = 1u << (nb & 31);
          const uint32_t old = atomicOr(&visited[nb >> 5], bit);
          if (!(old & bit)) fresh[atomicAdd(&s_nfresh, 1)] = nb;
          if (e == kEll - 1) atomicOr(&s_more, 1 << ci);
        }
        const unsigned vb = __ballot_sync(kFull, nb >= 0);  // ncur * kEll is a multiple of 32: warp-uniform trip count
        if (lane == 0 && vb) atomicAdd(&s_nedge, static_cast<unsigned long long>(__popc(vb)));
      }
      __syncthreads();  // (2)
      if (tid == 0) s_nexp += static_cast<unsigned long long>(ncur);
      const int more = s_more;
      int over_ci = -1;  // candidate whose CSR continuation is being drained (rare: rows longer than kEll)
      int64_t e_next = 0, e_end = 0;
      for (;;) {
        const int nfresh = s_nfresh;
        const unsigned long long bound = qa[L - 1] & kKeyMask;
        for (int i = warp; i < nfresh; i += nwarps) {
          const int nb = fresh[i];
          float d = warp_distance(a.metric, a.vec4 != 0, a.vectors + static_cast<int64_t>(nb) * a.dim, qv, a.dim, lane);
          if (lane == 0) {
            const unsigned long long key = make_key(d, static_cast<uint32_t>(nb));
            if (key < bound) cand[atomicAdd(&s_nacc, 1)] = key;
          }
        }
        __syncthreads();  // (3)
        const int m = s_nacc;
        if (m > 0) {
          for (int i = tid; i < m; i += blockDim.x) {
            const unsigned long long key = cand[i];
            int r = 0;
            for (int j = 0; j < m; ++j) r += (cand[j] < key);
            cs[r] = key;
          }
          __syncthreads();
          for (int i = tid; i < m; i += blockDim.x) pos[i] = lb_masked(qa, L, cs[i]);
          __syncthreads();
          const int p0 = pos[0];
          for (int hi = L; hi > p0; hi -= 8 * 128) {
            unsigned long long kreg[8];
            int dreg[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = hi - 1 - (u * 128 + tid);
              dreg[u] = L;
              if (j >= p0) {
                kreg[u] = qa[j];
                int sft = 0;
                if (m <= 8) { for (int i = 0; i < m; ++i) sft += (pos[i] <= j); }
                else { int lo = 0, up = m; while (lo < up) { const int mid = (lo + up) >> 1; if (pos[mid] <= j) lo = mid + 1; else up = mid; } sft = lo; }
                dreg[u] = j + sft;
              }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 8; ++u) if (dreg[u] < L) qa[dreg[u]] = kreg[u];
            __syncthreads();
          }
          for (int i = tid; i < m; i += blockDim.x) {
            const int f = pos[i] + i;
            if (f < L) qa[f] = cs[i];
          }
          if (tid == 0 && p0 < s_pmin) s_pmin = p0;
        }
        if (tid == 0) { s_ndist += static_cast<unsigned long long>(nfresh); s_nfresh = 0; s_nacc = 0; }
        __syncthreads();
        // next CSR continuation chunk, if any candidate's row exceeded the fixed stride
        while (e_next >= e_end) {
          ++over_ci;
          while (over_ci < ncur && !((more >> over_ci) & 1)) ++over_ci;
          if (over_ci >= ncur) break;
          e_next = a.offsets[s_cid[over_ci]] + kEll;
          e_end = a.offsets[s_cid[over_ci] + 1];
        }
        if (over_ci >= ncur) break;
        const int cnt = static_cast<int>(min(static_cast<int64_t>(kEll), e_end - e_next));
        if (tid < cnt) {
          const uint32_t nb = static_cast<uint32_t>(a.nbrs[e_next + tid]);
          const uint32_t bit = 1u << (nb & 31);
          const uint32_t old = atomicOr(&visited[nb >> 5], bit);
          if (!(old & bit)) fresh[atomicAdd(&s_nfresh, 1)] = static_cast<int>(nb);
        }
        e_next += cnt;
        __syncthreads();
      }
      const int pmin = s_pmin, first = s_first;
      k = pmin < first + 1 ? pmin : first + 1;
    }

    unsigned long long* out = a.out_queue + static_cast<int64_t>(q) * L;
    for (int i = tid; i < L; i += blockDim.x) out[i] = qa[i];
    {
      uint4* v4 = reinterpret_cast<uint4*>(visited);
      const int64_t n4 = a.visited_words >> 2;
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int64_t i = tid; i < n4; i += blockDim.x) v4[i] = z;
    }
    if (tid == 0) s_ndist += static_cast<unsigned long long>(L);
  }
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&a.stats[0], s_ndist);
    atomicAdd(&a.stats[1], s_nexp);
    atomicAdd(&a.stats[2], s_nedge);
  }
}

__global__ void csr_to_ell_kernel(const int64_t* __restrict__ offsets, const int32_t* __restrict__ nbrs, int64_t n,
                                  int32_t* __restrict__ ell) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n * kEll) return;
  const int64_t v = i / kEll;
  const int s = static_cast<int>(i % kEll);
  const int64_t e = offsets[v] + s;
  ell[i] = e < offsets[v + 1] ? nbrs[e] : -1;
}

__global__ void gather_rows_kernel(const float* __restrict__ vectors, const int32_t* __restrict__ ids, int n, int dim,
                                   float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(n) * dim) return;
  const int r = static_cast<int>(i / dim), c = static_cast<int>(i % dim);
  out[i] = vectors[static_cast<int64_t>(ids[r]) * dim + c];
}

// PrepareInitIds (vec_search_executor.cpp:487-516): dedup'd out-neighbours of the navigation point, then
// ids nav+1, nav+2, ... (mod n) until L entries.  Pure index logic on <= L + deg entries; runs on the host
// over the navigation row copied back from the device.  L is clamped to n_indexed by the caller (the
// reference loops forever when L > n_indexed).
int prepare_init_ids(Index* ix, int64_t L) {
  if (ix->init_L == L && ix->d_init_ids) return EPS_OK;
  int64_t e[2];
  EPS_CUDA(cudaMemcpyAsync(e, ix->d_offsets + ix->nav, 16, cudaMemcpyDeviceToHost, ix->stream));
  EPS_CUDA(cudaStreamSynchronize(ix->stream));
  std::vector<int32_t> row(static_cast<size_t>(e[1] - e[0]));
  if (!row.empty()) {
    EPS_CUDA(cudaMemcpyAsync(row.data(), ix->d_nbrs + e[0], row.size() * 4, cudaMemcpyDeviceToHost, ix->stream));
    EPS_CUDA(cudaStreamSynchronize(ix->stream));
  }
  std::vector<int32_t> ids;
  ids.reserve(static_cast<size_t>(L));
  std::vector<bool> sel(static_cast<size_t>(ix->n_indexed), false);
  for (size_t i = 0; i < row.size() && static_cast<int64_t>(ids.size()) < L; ++i) {
    int32_t v = row[i];
    if (sel[v]) continue;
    sel[v] = true;
    ids.push_back(v);
  }
  int64_t tmp = ix->nav + 1;
  while (static_cast<int64_t>(ids.size()) < L) {
    if (tmp == ix->n_indexed) tmp = 0;
    int64_t v = tmp++;
    if (sel[v]) continue;
    sel[v] = true;
    ids.push_back(static_cast<int32_t>(v));
  }
  if (ix->d_init_ids) { cudaFree(ix->d_init_ids); ix->d_init_ids = nullptr; }
  EPS_CUDA(cudaMalloc(&ix->d_init_ids, static_cast<size_t>(L) * 4));
  EPS_CUDA(cudaMemcpyAsync(ix->d_init_ids, ids.data(), static_cast<size_t>(L) * 4, cudaMemcpyHostToDevice, ix->stream));
  EPS_CUDA(cudaStreamSynchronize(ix->stream));
  ix->init_L = L;
  return EPS_OK;
}

int graph_search(Index* ix, const float* d_queries, int64_t nq, int64_t L, unsigned long long* d_queue,
                 eps_stats* stats) {
  if (L < 1 || L > ix->n_indexed) return fail(EPS_ERR_INVALID_ARGUMENT, "graph_search: L out of range");
  const int Lp = next_pow2(static_cast<int>(L));
  if (Lp > 16384) return fail(EPS_ERR_UNSUPPORTED, "SearchQueueSize above 16384 is not supported by the graph kernel");
  EPS_TRY(prepare_init_ids(ix, L));
  const int dimp = (static_cast<int>(ix->dim) + 3) & ~3;
  const bool use_v2 = getenv("EPS_GRAPH_V1") == nullptr;
  const int width = use_v2 ? std::max(1, std::min(ix->search_width, kWideMax)) : 1;
  const size_t smem = width > 1 ? static_cast<size_t>(Lp) * 8 + 2 * kWCap * 8 + static_cast<size_t>(dimp) * 4 + 2 * kWCap * 4
                      : use_v2 ? static_cast<size_t>(Lp) * 8 + 2 * kCH * 8 + static_cast<size_t>(dimp) * 4 + 2 * kCH * 4 + 2 * kEll * 4
                             : static_cast<size_t>(Lp) * 8 + 2 * kCH * 8 + static_cast<size_t>(dimp) * 4 + 2 * kCH * 4;
  if (smem > 220 * 1024) return fail(EPS_ERR_UNSUPPORTED, "queue + query do not fit in shared memory");
  int per_sm = 0;
  if (width > 1) {
    EPS_CUDA(cudaFuncSetAttribute(graph_search_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    EPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, graph_search_kernel_wide, 128, smem));
  } else if (use_v2) {
    EPS_CUDA(cudaFuncSetAttribute(graph_search_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    EPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, graph_search_kernel_v2, 128, smem));
  } else {
    EPS_CUDA(cudaFuncSetAttribute(graph_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    EPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, graph_search_kernel, 128, smem));
  }
  if (per_sm < 1) per_sm = 1;
  int slots = static_cast<int>(std::min<int64_t>(nq, static_cast<int64_t>(per_sm) * ix->num_sms));
  const int64_t words = ((ix->n_indexed + 31) / 32 + 3) & ~3ll;
  if (ix->visited_slots < slots || ix->s_visited.cap < static_cast<size_t>(slots) * words * 4) {
    EPS_TRY(ix->s_visited.reserve(static_cast<size_t>(slots) * words * 4));
    ix->visited_slots = slots;
  }
  // bitmaps must start clean; the kernel leaves them clean.  (Re)zero when the geometry changed.
  if (ix->vis_clean_ptr != ix->s_visited.p || ix->vis_clean_words != words) {
    EPS_CUDA(cudaMemsetAsync(ix->s_visited.p, 0, ix->s_visited.cap, ix->stream));
    ix->vis_clean_ptr = ix->s_visited.p;
    ix->vis_clean_words = words;
  }
  EPS_TRY(ix->s_misc.reserve(64));
  EPS_CUDA(cudaMemsetAsync(ix->s_misc.p, 0, 64, ix->stream));
  uint64_t launches = 1;
  if (use_v2) {
    if (!ix->d_ell) {  // fixed-stride adjacency, built once per installed graph
      EPS_CUDA(cudaMalloc(&ix->d_ell, static_cast<size_t>(ix->n_indexed) * kEll * 4));
      const int64_t tot = ix->n_indexed * kEll;
      csr_to_ell_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, ix->stream>>>(ix->d_offsets, ix->d_nbrs,
                                                                                          ix->n_indexed, ix->d_ell);
      EPS_CUDA(cudaGetLastError());
    }
    if (ix->seed_rows_L != L) {  // contiguous copy of the query-independent seed rows
      EPS_TRY(ix->s_seed_rows.reserve(static_cast<size_t>(L) * ix->dim * 4));
      const int64_t tot = L * ix->dim;
      gather_rows_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, ix->stream>>>(
          ix->d_vectors, ix->d_init_ids, static_cast<int>(L), static_cast<int>(ix->dim), ix->s_seed_rows.as<float>());
      EPS_CUDA(cudaGetLastError());
      ix->seed_rows_L = L;
    }
    const int64_t seed_ld = (L + 3) & ~3ll;
    EPS_TRY(ix->s_seed_dist.reserve(static_cast<size_t>(nq) * seed_ld * 4));
    EPS_TRY(launch_distances(ix, ix->s_seed_rows.as<float>(), 0, L, d_queries, nq, ix->s_seed_dist.as<float>(), seed_ld,
                             &launches));
    GS2Args a;
    a.vectors = ix->d_vectors; a.offsets = ix->d_offsets; a.nbrs = ix->d_nbrs; a.ell = ix->d_ell;
    a.init_ids = ix->d_init_ids; a.seed_dist = ix->s_seed_dist.as<float>(); a.queries = d_queries;
    a.visited = ix->s_visited.as<uint32_t>(); a.out_queue = d_queue;
    a.work_counter = reinterpret_cast<int*>(ix->s_misc.as<unsigned char>() + 32);
    a.stats = ix->s_misc.as<unsigned long long>();
    a.visited_words = words; a.seed_ld = seed_ld; a.dim = static_cast<int>(ix->dim); a.metric = ix->metric;
    a.vec4 = ix->vec4 ? 1 : 0; a.L = static_cast<int>(L); a.Lp = Lp; a.nq = static_cast<int>(nq);
    if (width > 1) graph_search_kernel_wide<<<slots, 128, smem, ix->stream>>>(a, width);
    else graph_search_kernel_v2<<<slots, 128, smem, ix->stream>>>(a);
  } else {
    GSArgs a;
    a.vectors = ix->d_vectors; a.offsets = ix->d_offsets; a.nbrs = ix->d_nbrs; a.init_ids = ix->d_init_ids;
    a.queries = d_queries; a.visited = ix->s_visited.as<uint32_t>(); a.out_queue = d_queue;
    a.work_counter = reinterpret_cast<int*>(ix->s_misc.as<unsigned char>() + 32);
    a.stats = ix->s_misc.as<unsigned long long>();
    a.visited_words = words; a.dim = static_cast<int>(ix->dim); a.metric = ix->metric; a.vec4 = ix->vec4 ? 1 : 0;
    a.L = static_cast<int>(L); a.Lp = Lp; a.nq = static_cast<int>(nq);
    graph_search_kernel<<<slots, 128, smem, ix->stream>>>(a);
  }
  EPS_CUDA(cudaGetLastError());
  if (stats) {
    stats->n_seed += static_cast<uint64_t>(nq) * static_cast<uint64_t>(L);
    stats->kernel_launches += launches;
  }
  return EPS_OK;
}

// Device counters of the last graph_search launch (call after the stream has been synchronised).
int read_graph_counters(Index* ix, eps_stats* stats) {
  if (!stats || !ix->s_misc.p) return EPS_OK;
  unsigned long long h[4];
  EPS_CUDA(cudaMemcpy(h, ix->s_misc.p, 32, cudaMemcpyDeviceToHost));
  stats->n_dist += h[0];
  stats->n_expand += h[1];
  stats->n_edges += h[2];
  return EPS_OK;
}

}  // namespace eps
