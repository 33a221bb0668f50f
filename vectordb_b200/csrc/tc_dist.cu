// K1-TC — coarse distance tiles of the exact scan on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// The large-batch exact scan (BruteForceSearch / PreFilter branch, engine/db/execution/vec_search_executor.cpp
// :717-831, at batch B) is a dense contraction D[q,r] = <Q[q,:], X[r,:]> with 2*B*N*d flops against N*d*4 bytes
// (512 flop/B at B = 1024): the fp32 SIMT tile kernel in brute_force.cu sits at 75 % of the SIMT issue rate and
// still needs 583 ms per 10M x 768 x 1024 batch.  This kernel produces the same [B x chunk] distance tile with
// tcgen05.mma kind::tf32:
//   * operands stay the reference's fp32 rows — TMA (cp.async.bulk.tensor, 128-byte swizzle) stages 128-row x 32-
//     float blocks of the table and 256-query x 32-float blocks of the query batch into a 4-stage shared-memory
//     ring; the tensor core reads them as TF32 (no converted copy of the table is kept);
//   * one elected thread issues 4 MMAs (M128 x N256 x K8) per block into a TMEM accumulator (256 columns, two
//     accumulators = all 512 columns, so the epilogue of tile t overlaps the MMAs of tile t+1);
//   * the epilogue warps read TMEM with tcgen05.ld, turn dot products into the metric
//     (|x|^2 + |q|^2 - 2 dot | 1 - dot | -dot), and store the tile query-major, coalesced.
// TF32 products carry ~1e-3 relative error, so this is a COARSE pass: bf_select_kernel keeps k' > k candidates
// per query and rescore_kernel re-evaluates them with the exact fp32 direct form before the final (distance,id)
// sort — returned ids/distances are those of the fp32 path ("TF32 MMA with fp32 re-score of
// survivors").  Bound: tensor pipe / L2 operand traffic.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "async.cuh"
#include "internal.h"

namespace eps {

constexpr int kTcBM = 128;      // table rows per tile  (UMMA M)
constexpr int kTcBN = 256;      // queries per tile     (UMMA N)
constexpr int kTcBK = 32;       // floats per k-block = one 128-byte swizzle atom
constexpr int kTcStages = 4;
constexpr int kTcABytes = kTcBM * kTcBK * 4;   // 16 KB
constexpr int kTcBBytes = kTcBN * kTcBK * 4;   // 32 KB
constexpr int kTcStageBytes = kTcABytes + kTcBBytes;
constexpr int kTcThreads = 192;  // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-5 epilogue
constexpr int kTcStgCap = 512;   // staged survivors per epilogue warp (8-byte key + 2-byte query index each)
constexpr int kTcStgBytes = 4 * kTcStgCap * 10 + 16;
constexpr int kTcSmem = kTcStages * kTcStageBytes + 1024 /*align*/ + 8192 /*qnorm + thresholds*/ + 256 /*barriers*/ + kTcStgBytes;

struct TcArgs {
  int64_t row_start;   // absolute first row of this chunk
  int64_t n;           // rows in this chunk
  int64_t nq;
  int64_t ldd;
  const float* xnorm;  // |x|^2 per absolute row (L2 only)
  const float* qnorm;  // |q|^2 per query (L2 only)
  float* D;            // [nq x ldd]
  int dim;
  int metric;
  int n_row_tiles;
  int n_q_tiles;
  int kb_elems;        // elements per 128-byte k-block: 32 (fp32 read as TF32) or 64 (bf16 mirror)
  uint32_t idesc;      // UMMA instruction descriptor for the operand type
  // fused selection (D == nullptr): only entries below the query's running threshold leave the SM
  const float* thr;             // [nq] coarse k'-th best so far
  unsigned long long* cand;     // [nq x cand_cap] keys
  int* cand_cnt;                // [nq]
  const uint32_t* pass;         // deleted / static-filter bitmap relative to pass_base (may be null)
  int64_t pass_base;
  int cand_cap;
};

// Survivors of the fused selection are staged per epilogue warp in shared memory and flushed once per tile.  A push
// straight to the per-query list needs the value its global atomicAdd returns (a ~1 us round trip under contention)
// before the warp can go on, and while the running thresholds are still loose — the first fused launches of a scan,
// or a clustered table where whole blobs pass — those round trips, serialised per warp, were most of the launch
// (32 K rows took 0.7 ms, 20x their MMA time).  The stage is flushed when half full and at the end of the CTA's tiles,
// four independent global atomics per lane at a time.
struct TcStage {
  unsigned long long* key;  // [kTcStgCap]
  unsigned short* q;        // [kTcStgCap]
  int* cnt;
};

__device__ __forceinline__ void cand_push_global(const TcArgs& a, int q, unsigned long long key) {
  const int slot = atomicAdd(&a.cand_cnt[q], 1);
  if (slot < a.cand_cap) a.cand[static_cast<int64_t>(q) * a.cand_cap + slot] = key;
}

__device__ __forceinline__ void cand_flush(const TcArgs& a, const TcStage& st, int lane) {
  __syncwarp();
  const int n = min(*st.cnt, kTcStgCap);
  for (int base = 0; base < n; base += 128) {
    int q[4], slot[4];
    unsigned long long key[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + 32 * u + lane;
      q[u] = idx < n ? st.q[idx] : -1;
      key[u] = idx < n ? st.key[idx] : 0ull;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) slot[u] = q[u] >= 0 ? atomicAdd(&a.cand_cnt[q[u]], 1) : a.cand_cap;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (slot[u] < a.cand_cap) a.cand[static_cast<int64_t>(q[u]) * a.cand_cap + slot[u]] = key[u];
  }
  __syncwarp();
  if (lane == 0) *st.cnt = 0;
  __syncwarp();
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), descriptor
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);        // start address, 16-byte units
  d |= static_cast<uint64_t>(1) << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // stride byte offset = 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                        // version = 1
  d |= static_cast<uint64_t>(2) << 61;                        // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256 (InstrDescriptor bit layout, same header).
__host__ __device__ __forceinline__ uint32_t umma_idesc(uint32_t fmt /* 2 = TF32, 1 = BF16 */) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= fmt << 7;                     // a_format
  d |= fmt << 10;                    // b_format
  d |= (kTcBN >> 3) << 17;           // n_dim
  d |= (kTcBM >> 4) << 24;           // m_dim
  return d;
}
template <bool BF16>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (BF16) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
    return;
  }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// one warp reads its 32 TMEM lanes x 32 fp32 columns (asynchronous until tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32(uint32_t (&v)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue work on one 32-column chunk of the accumulator (v[j] = dot(row, query q0+j) for this thread's row).
__device__ __forceinline__ void epi_chunk(const TcArgs& a, const uint32_t (&v)[32], int q0, bool row_ok, float xn, int64_t i,
                                          int64_t row_abs, const float* thr_s, const float* qn_s, const TcStage& st) {
  if (a.D == nullptr) {
    // ---- fused selection: 2 instructions per element (FFMA + compare), survivors are rare ----
    if (row_ok) {
      const float m = a.metric == EPS_METRIC_L2 ? -2.0f : -1.0f;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 ct = *reinterpret_cast<const float4*>(thr_s + q0 + 4 * j4);
        const float t0 = fmaf(m, __uint_as_float(v[4 * j4 + 0]), xn), t1 = fmaf(m, __uint_as_float(v[4 * j4 + 1]), xn);
        const float t2 = fmaf(m, __uint_as_float(v[4 * j4 + 2]), xn), t3 = fmaf(m, __uint_as_float(v[4 * j4 + 3]), xn);
        if ((t0 < ct.x) | (t1 < ct.y) | (t2 < ct.z) | (t3 < ct.w)) {
          const float tt[4] = {t0, t1, t2, t3};
          const float cc[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (tt[u] < cc[u]) {
              const int q = q0 + 4 * j4 + u;
              float d = tt[u];
              if (a.metric == EPS_METRIC_L2) d = fmaxf(d + qn_s[q], 0.f);
              else if (a.metric == EPS_METRIC_COSINE) d = 1.0f + d;
              const unsigned long long key = make_key(d, static_cast<uint32_t>(row_abs));
              const int pos = atomicAdd(st.cnt, 1);  // shared memory
              if (pos < kTcStgCap) { st.key[pos] = key; st.q[pos] = static_cast<unsigned short>(q); }
              else cand_push_global(a, q, key);      // stage full (flushed at the next chunk boundary)
            }
          }
        }
      }
    }
  } else {
    // ---- distance tile to global memory (first chunk: seeds the running thresholds) ----
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int q = q0 + j;
      if (row_ok && q < a.nq) {
        const float dot = __uint_as_float(v[j]);
        float d;
        if (a.metric == EPS_METRIC_L2) d = fmaxf(xn + qn_s[q] - 2.0f * dot, 0.f);
        else if (a.metric == EPS_METRIC_COSINE) d = 1.0f - dot;
        else d = -dot;
        a.D[static_cast<int64_t>(q) * a.ldd + i] = d;
      }
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) tc_dist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, TcArgs a) {
  extern __shared__ unsigned char tc_smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle pattern
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  float* qn_s = reinterpret_cast<float*>(base + kTcStages * kTcStageBytes);  // [<=1024]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kTcStages * kTcStageBytes + 8192);
  // bars[0..3] full, [4..7] empty, [8..9] tmem_full, [10..11] tmem_empty, then the TMEM base slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  unsigned char* stg = base + kTcStages * kTcStageBytes + 8192 + 256;  // [4 x keys][4 x query indices][4 counters]
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 4), tfull0 = smem_u32(bars + 8), tempty0 = smem_u32(bars + 10);
  const uint32_t stage0 = smem_u32(base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (a.dim + a.kb_elems - 1) / a.kb_elems;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(tfull0 + 8 * t, 1); mbar_init(tempty0 + 8 * t, 4); }
    mbar_fence_init();
  }
  if (warp == 1) {  // TMEM: all 512 columns (two 256-column fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  float* thr_s = qn_s + 1024;  // [<=1024] thresholds (fused mode)
  // fused mode compares t = m*dot + xn (m = -2 for L2, -1 otherwise) with a per-query constant:
  //   L2: d = t + |q|^2 < thr  <=>  t < thr - |q|^2;   cosine: d = 1 + t < thr  <=>  t < thr - 1;   IP: d = t < thr.
  // Queries beyond nq get -inf, so they never pass and need no bounds test in the inner loop.
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const float qn = (a.metric == EPS_METRIC_L2 && i < a.nq) ? a.qnorm[i] : 0.f;
    qn_s[i] = qn;
    float c = -INFINITY;
    if (a.D == nullptr && i < a.nq) {
      const float th = a.thr[i];
      c = a.metric == EPS_METRIC_L2 ? th - qn : (a.metric == EPS_METRIC_COSINE ? th - 1.0f : th);
    }
    thr_s[i] = c;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_row = a.n_q_tiles;
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      uint32_t it = 0;
      for (int rt = blockIdx.x; rt < a.n_row_tiles; rt += gridDim.x) {
        const int row0 = static_cast<int>(a.row_start) + rt * kTcBM;
        for (int qt = 0; qt < tiles_per_row; ++qt) {
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            mbar_expect_tx(full0 + 8 * s, kTcStageBytes);
            const uint32_t sa = stage0 + s * kTcStageBytes;
            tma_load_2d(sa, &tmA, kb * a.kb_elems, row0, full0 + 8 * s);
            tma_load_2d(sa + kTcABytes, &tmB, kb * a.kb_elems, qt * kTcBN, full0 + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const bool bf16 = a.kb_elems == 64;
      const uint32_t idesc = bf16 ? umma_idesc(1u) : umma_idesc(2u);
      uint32_t it = 0, tc = 0;
      for (int rt = blockIdx.x; rt < a.n_row_tiles; rt += gridDim.x) {
        for (int qt = 0; qt < tiles_per_row; ++qt, ++tc) {
          const uint32_t acc = tc & 1, aph = (tc >> 1) & 1;
          mbar_wait(tempty0 + 8 * acc, aph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + acc * kTcBN;
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = stage0 + s * kTcStageBytes;
            const uint64_t ad = umma_desc(sa), bd = umma_desc(sa + kTcABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 MMAs per 128-byte block: K = 8 tf32 / 16 bf16 = 32 bytes = +2 (16-B units)
              if (bf16) umma_issue<true>(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
              else umma_issue<false>(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit(empty0 + 8 * s);  // frees the stage when these MMAs retire
          }
          umma_commit(tfull0 + 8 * acc);  // accumulator complete
        }
      }
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 =====
    const int lq = (warp & 3) * 32;
    TcStage st;
    st.key = reinterpret_cast<unsigned long long*>(stg) + (warp & 3) * kTcStgCap;
    st.q = reinterpret_cast<unsigned short*>(stg + 4 * kTcStgCap * 8) + (warp & 3) * kTcStgCap;
    st.cnt = reinterpret_cast<int*>(stg + 4 * kTcStgCap * 10) + (warp & 3);
    if (lane == 0) *st.cnt = 0;
    __syncwarp();
    uint32_t tc = 0;
    for (int rt = blockIdx.x; rt < a.n_row_tiles; rt += gridDim.x) {
      const int64_t i = static_cast<int64_t>(rt) * kTcBM + lq + lane;  // row index inside the chunk
      bool row_ok = i < a.n;
      float xn = 0.f;
      if (a.metric == EPS_METRIC_L2 && row_ok) xn = a.xnorm[a.row_start + i];
      const int64_t row_abs = a.row_start + i;
      if (a.D == nullptr && a.pass && row_ok) {
        const int64_t pi = row_abs - a.pass_base;
        row_ok = (a.pass[pi >> 5] >> (pi & 31)) & 1u;
      }
      for (int qt = 0; qt < tiles_per_row; ++qt, ++tc) {
        const uint32_t acc = tc & 1, aph = (tc >> 1) & 1;
        mbar_wait(tfull0 + 8 * acc, aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(lq) << 16) + acc * kTcBN;
        const int qbase = qt * kTcBN;
#pragma unroll 1
        for (int c = 0; c < kTcBN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(v, taddr0 + c * 32);
          tmem_ld_wait();
          epi_chunk(a, v, qbase + c * 32, row_ok, xn, i, row_abs, thr_s, qn_s, st);
          if (a.D == nullptr) {
            __syncwarp();
            if (*st.cnt >= kTcStgCap / 2) cand_flush(a, st, lane);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      }
    }
    if (a.D == nullptr) cand_flush(a, st, lane);  // what is left in the stage (it is also flushed whenever half full)
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// |x|^2 per row (fp32, warp per row) for the L2 expansion of the coarse pass.
__global__ void row_norm_kernel(const float* __restrict__ v, int64_t row0, int64_t n, int dim, float* __restrict__ out) {
  const int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const float* p = v + (row0 + w) * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s = fmaf(p[i], p[i], s);
  s = warp_sum(s);
  if (lane == 0) out[row0 + w] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* tm, const void* base, int64_t rows, int dim, int box_rows, bool bf16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(EPS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const int esz = bf16 ? 2 : 4;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};  // 128-byte rows
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
                   gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EPS_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(static_cast<int>(r)));
  return EPS_OK;
}

__global__ void to_bf16_kernel(const float* __restrict__ in, int64_t n, unsigned short* __restrict__ out) {
  const int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(out + i) = o;
  } else {
    for (int64_t j = i; j < n; ++j) {
      const __nv_bfloat16 h = __float2bfloat16_rn(in[j]);
      out[j] = *reinterpret_cast<const unsigned short*>(&h);
    }
  }
}

bool tc_dist_usable(const Index* ix, int64_t nq) {
  static const bool env_off = getenv("EPS_NO_TC") != nullptr;  // read once
  if (env_off || ix->coarse_mode == 0) return false;
  if (nq < 64 || nq > 1024) return false;            // per-query constants live in a 4 KB shared-memory table
  const int align = ix->coarse_mode == 2 ? 8 : 4;    // TMA: 16-byte row pitch
  if (ix->dim % align != 0 || ix->dim < 32) return false;
  if ((reinterpret_cast<uintptr_t>(ix->d_vectors) & 15) != 0) return false;
  return get_encode() != nullptr;
}

// Same contract as launch_distances() (D[q*ldd + i] for rows [row_start, row_start+n)), coarse values; with
// `fused` the tile is filtered against the running thresholds in the epilogue instead of being written.
int tc_launch_distances(Index* ix, int64_t row_start, int64_t n, const float* d_queries, int64_t nq, float* D,
                        int64_t ldd, uint64_t* launches, const TcFused* fused) {
  const int dim = static_cast<int>(ix->dim);
  const bool bf16 = ix->coarse_mode == 2;
  if (ix->metric == EPS_METRIC_L2) {
    if (ix->xnorm_rows < ix->n_rows) {  // row norms for rows appended since the last call
      EPS_TRY(ix->s_xnorm.reserve(static_cast<size_t>(ix->capacity > ix->n_rows ? ix->capacity : ix->n_rows) * 4));
      if (ix->s_xnorm.p != ix->xnorm_ptr) { ix->xnorm_rows = 0; ix->xnorm_ptr = ix->s_xnorm.p; }
      const int64_t cnt = ix->n_rows - ix->xnorm_rows;
      row_norm_kernel<<<static_cast<unsigned>((cnt * 32 + 255) / 256), 256, 0, ix->stream>>>(ix->d_vectors, ix->xnorm_rows, cnt,
                                                                                            dim, ix->s_xnorm.as<float>());
      ix->xnorm_rows = ix->n_rows;
      ++*launches;
    }
    EPS_TRY(ix->s_qnorm.reserve(static_cast<size_t>(nq) * 4));
    row_norm_kernel<<<static_cast<unsigned>((nq * 32 + 255) / 256), 256, 0, ix->stream>>>(d_queries, 0, nq, dim,
                                                                                         ix->s_qnorm.as<float>());
    ++*launches;
  }
  const void* a_base = ix->d_vectors;
  const void* b_base = d_queries;
  if (bf16) {
    // bf16 mirror of the table (coarse pass only; the re-score reads the fp32 rows), kept current incrementally
    if (ix->bf16_rows < ix->n_rows) {
      EPS_TRY(ix->s_bf16.reserve(static_cast<size_t>(ix->capacity > ix->n_rows ? ix->capacity : ix->n_rows) * dim * 2));
      if (ix->s_bf16.p != ix->bf16_ptr) { ix->bf16_rows = 0; ix->bf16_ptr = ix->s_bf16.p; }
      const int64_t cnt = (ix->n_rows - ix->bf16_rows) * dim;
      to_bf16_kernel<<<static_cast<unsigned>((cnt / 4 + 256) / 256), 256, 0, ix->stream>>>(
          ix->d_vectors + ix->bf16_rows * dim, cnt, ix->s_bf16.as<unsigned short>() + ix->bf16_rows * dim);
      ix->bf16_rows = ix->n_rows;
      ++*launches;
    }
    EPS_TRY(ix->s_qbf16.reserve(static_cast<size_t>(nq) * dim * 2));
    const int64_t cnt = nq * dim;
    to_bf16_kernel<<<static_cast<unsigned>((cnt / 4 + 256) / 256), 256, 0, ix->stream>>>(d_queries, cnt, ix->s_qbf16.as<unsigned short>());
    ++*launches;
    a_base = ix->s_bf16.p;
    b_base = ix->s_qbf16.p;
  }
  CUtensorMap tmA, tmB;
  EPS_TRY(make_map(&tmA, a_base, ix->n_rows, dim, kTcBM, bf16));
  EPS_TRY(make_map(&tmB, b_base, nq, dim, kTcBN, bf16));
  TcArgs a;
  a.row_start = row_start; a.n = n; a.nq = nq; a.ldd = ldd;
  a.xnorm = ix->s_xnorm.as<float>(); a.qnorm = ix->s_qnorm.as<float>(); a.D = D; a.dim = dim; a.metric = ix->metric;
  a.kb_elems = bf16 ? 64 : 32;
  a.idesc = umma_idesc(bf16 ? 1u : 2u);
  a.thr = nullptr; a.cand = nullptr; a.cand_cnt = nullptr; a.pass = nullptr; a.pass_base = 0; a.cand_cap = 0;
  if (fused) {
    a.D = nullptr;
    a.thr = fused->thr; a.cand = fused->cand; a.cand_cnt = fused->cand_cnt; a.pass = fused->pass;
    a.pass_base = fused->pass_base; a.cand_cap = fused->cand_cap;
  }
  a.n_row_tiles = static_cast<int>((n + kTcBM - 1) / kTcBM);
  a.n_q_tiles = static_cast<int>((nq + kTcBN - 1) / kTcBN);
  static bool attr_set = false;
  if (!attr_set) {
    EPS_CUDA(cudaFuncSetAttribute(tc_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
    attr_set = true;
  }
  const int grid = std::min(a.n_row_tiles, ix->num_sms);
  tc_dist_kernel<<<grid, kTcThreads, kTcSmem, ix->stream>>>(tmA, tmB, a);
  EPS_CUDA(cudaGetLastError());
  ++*launches;
  return EPS_OK;
}

}  // namespace eps
