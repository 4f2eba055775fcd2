// Single-pass Ising-model extraction on sm_100a (the headline kernel).
//
// Replaces, fused into ONE kernel launch:
//   lattice_symmetries Operator.batched_apply      (called at annealing_sign_problem/common.py:96)
//   _clipped_search_sorted + membership mask       (common.py:116-128, :173)
//   bsearch loop of build_matrix                   (cbits/build_matrix.c:33-51)
//   _make_ising_model_compute_elements             (common.py:71-82)
//   csr_matrix(...) + sort_indices                 (common.py:193-195)
//
// The path is bound by instruction issue, not by HBM: ~37 candidates are
// generated and looked up per row and only a tenth of them are couplings.  The kernel therefore
// spends its instructions where the candidates are:
//   A. applicability of all moves for the 32 rows of a warp is computed on bit planes (the
//      32 x N key matrix transposed with warp shuffles): one LOP3 decides a move for 32 rows;
//      transposed back, every lane holds the bit mask of the moves that apply to ITS row;
//   B. every lane walks its own mask (all 32 lanes busy) and pre-sieves each candidate in an
//      order-preserving blocked Bloom filter (one 8-byte word per candidate, two hashed bits:
//      ~2 % false positives) -- nine misses in ten end here, after ~25 instructions;
//   C. survivors are re-dealt evenly over the lanes, searched exactly (4-byte first-position
//      table over the top key bits + a short scan of the keys) and appended, with their in-row
//      rank, to the warp's hit list (global scratch, L2-resident);
//   D. a tile (the 32 rows of a warp; warps never synchronise with each other) publishes its
//      coupling count, obtains its CSR offset by decoupled look-back over earlier tiles (tiles
//      are handed out by an atomic ticket, so a tile only waits for tiles that already run) and
//      writes indptr / indices / data in place.
// Every lane enumerates its moves in ascending key-delta order, so hits get ascending columns;
// positions are exact indices into the sorted basis (bit-exact CSR) and values are
// (coef * |psi_j|) * |psi_i|, the association of the reference's live path (common.py:71-82), as in extract.cu.
#include <mutex>
#include <unordered_set>

#include "fused.cuh"

namespace asp {

constexpr int kFxWarps = 8;
constexpr int kFxThreads = kFxWarps * 32;
constexpr int kFxTileRows = 32;  // a tile is the 32 rows of one warp
static_assert(kFxTileRows == kFusedTileRows, "fused.cuh out of sync");
constexpr uint32_t kFxMaxMoves = 2046;  // tag layout: move 11 bits | row lane 5 bits | in-row rank 12 bits
constexpr int kFxSurvSlotsDefault = 24; // survivor slots per lane between two exact-search rounds
constexpr int kFxMaxCtasPerSM = 8;
#ifndef ASP_FX_IN_FLIGHT
#define ASP_FX_IN_FLIGHT 4  // filter words a lane has in flight in the candidate walk (4 or 2)
#endif
#ifndef ASP_FX_BACKOFF_NS
#define ASP_FX_BACKOFF_NS 32  // first sleep of a look-back that finds an unpublished tile
#endif
#ifndef ASP_FX_LAG
#define ASP_FX_LAG 1
#endif
#ifndef ASP_FX_SCRATCH_MB
#define ASP_FX_SCRATCH_MB 512
#endif
constexpr int kFxLag = ASP_FX_LAG;  // tiles a warp counts ahead of the tile whose CSR offset it fetches
constexpr size_t kFxScratchBudget = static_cast<size_t>(ASP_FX_SCRATCH_MB) << 20;  // hit-list scratch never exceeds this

constexpr unsigned long long kFlagAggregate = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

// Shared-memory layout.  CTA-wide tables, then one block per warp; inside a warp's block the
// bit planes (stage A) share their bytes with the survivor slots (stages B/C) and the apply
// masks (A/B) with the row offsets and amplitudes of the write phase.
struct FxLayout {
  uint32_t cand, flip, coef, desc, mask, need, groups, diag;  // byte offsets of the tables
  uint32_t tables;                                            // bytes of all tables
  uint32_t w_surv, w_amask, w_pre, w_cnt, w_pend;             // byte offsets inside a warp's block
  uint32_t per_warp;
};

struct FusedArgs {
  FxLayout layout;         // computed on the host: the kernel reads the offsets from constant memory
  const uint64_t *spins;   // [n_total] ascending, unique
  const double *psi;
  const uint32_t *starts;  // [2^tbits + 1] first position with key >= bucket << tshift
  int tshift;
  const uint2 *filter;     // [2^fbits] blocked Bloom filter, word = key >> fshift
  int fshift;
  uint64_t state_mask;     // keys with bits outside are never candidates
  uint32_t n_total;
  uint64_t row_begin, num_rows, num_tiles;
  const Move *moves;
  int n_moves, n_down, n_words;
  const DiagBond *diag;
  int n_diag;
  const DiagGroup *groups; // closed form of the diagonal (n_groups < 0: use the bond loop)
  int n_groups, diag_scale;
  long long diag_c0;
  int surv_slots;          // survivor slots per lane
  int planes_ok;           // every move mask has exactly two bits: stage A on bit planes
  uint2 *scratch;          // [gridDim.x * kFxWarps][kFxLag + 1][scratch_per_warp] hit lists {position, tag}
  uint32_t scratch_per_warp;
  unsigned long long *status;  // [num_tiles] look-back words (zeroed)
  unsigned int *ticket;        // zeroed
  uint64_t capacity;           // entries the caller's indices/data can hold
  int64_t *indptr;
  int32_t *indices;
  double *data;
  unsigned long long *nnz_out;          // running total after this launch (device)
  unsigned long long *nnz_mirror;       // optional copy of it in mapped host memory
  const unsigned long long *base_in;   // couplings emitted by earlier row chunks (NULL = 0)
};

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Two hashed bit positions of a key inside its filter word: bits 0..4 (first half-word) and bits
// 8..12 (second half-word) of a GF(2)-LINEAR mix, so hash(s ^ flip) = hash(s) ^ hash(flip): the
// kernel hashes every row once and every move once instead of every candidate.
__host__ __device__ __forceinline__ uint32_t filter_hash(uint64_t key) {
  const uint32_t lo = static_cast<uint32_t>(key), hi = static_cast<uint32_t>(key >> 32);
  uint32_t y = lo ^ ((hi << 7) | (hi >> 25));
  y ^= y >> 15;
  y ^= y << 11;
  y ^= y >> 7;
  y ^= y << 3;
  y ^= y >> 17;
  return y;
}

// The two bits as one 64-bit word (uint2 {x, y} little endian): the index kernels set them with ONE
// 64-bit atomic -- the index build is bound by L2 atomic throughput.
__host__ __device__ __forceinline__ unsigned long long filter_bits(uint32_t h) {
  return (1ull << (h & 31u)) | (1ull << (32u + ((h >> 8) & 31u)));
}

// Loads that bypass L1 allocation: filter words, table entries, keys and gathered amplitudes are
// touched once per SM, while the warp's hit list (written, then read back) should stay in L1.
#ifndef ASP_FX_STREAM_FILTER
#define ASP_FX_STREAM_FILTER 0
#endif
#ifndef ASP_FX_STREAM_SEARCH
#define ASP_FX_STREAM_SEARCH 0
#endif
#ifndef ASP_FX_STREAM_PSI
#define ASP_FX_STREAM_PSI 0
#endif
__device__ __forceinline__ uint2 ldg_filter_u2(const uint2 *p) {
#if ASP_FX_STREAM_FILTER
  uint2 v;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t *p) {
#if ASP_FX_STREAM_SEARCH
  uint32_t v;
  asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ uint64_t ldg_stream_u64(const uint64_t *p) {
#if ASP_FX_STREAM_SEARCH
  uint64_t v;
  asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ double ldg_stream_f64(const double *p) {
#if ASP_FX_STREAM_PSI
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
// Shared memory through 32-bit addresses (the walk of stage B keeps running addresses).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));  // ordered with the other volatile asm, transparent to plain loads
  return v;
}
__device__ __forceinline__ uint2 lds_table_u2(uint32_t addr) {  // read-only tables: free to schedule
  uint2 v;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)));
}

#include "exchange_kernels.cuh"  // index + X1 kernels (gather_index_*, gather_copy_tma_kernel, index_block_kernel)

// Position of c (< 2^number_spins) in the sorted basis, or -1.
__device__ __forceinline__ int32_t search_one(const FusedArgs &a, uint64_t c) {
  const uint32_t *bucket = a.starts + (c >> a.tshift);
  uint32_t lo = ldg_stream_u32(bucket), hi = ldg_stream_u32(bucket + 1);
  if (hi - lo > 16) {  // pathological bucket: bisect down to a short scan
    do {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (ldg_stream_u64(&a.spins[mid]) < c)
        lo = mid + 1;
      else
        hi = mid + 1;
    } while (hi - lo > 8);
  }
  for (; lo < hi; lo += 2) {  // two keys per step (independent loads)
    const uint64_t k0 = ldg_stream_u64(&a.spins[lo]);
    const bool second = lo + 1 < hi;
    const uint64_t k1 = second ? ldg_stream_u64(&a.spins[lo + 1]) : 0ull;
    if (k0 >= c) return k0 == c ? static_cast<int32_t>(lo) : -1;
    if (second && k1 >= c) return k1 == c ? static_cast<int32_t>(lo + 1) : -1;
  }
  return -1;
}

// 32 x 32 bit-matrix transpose across a warp: result bit k of lane l = bit l of lane k's x.
__device__ __forceinline__ uint32_t transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const uint32_t low = d == 16 ? 0x0000FFFFu : d == 8 ? 0x00FF00FFu : d == 4 ? 0x0F0F0F0Fu : d == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, d);
    x = (lane & d) ? ((x & ~low) | ((y >> d) & low)) : ((x & low) | ((y << d) & ~low));
  }
  return x;
}

// Diagonal matrix element of one basis word: bond contributions summed in (term, bond) order.
__device__ __forceinline__ double diagonal_element(uint64_t s, const DiagBond *s_diag, int n_diag) {
  double d = 0.0;
  for (int k = 0; k < n_diag; ++k) {
    const DiagBond &db = s_diag[k];
    const int idx = static_cast<int>(((s >> db.i) & 1) * 2 + ((s >> db.j) & 1));
    d += db.d[idx];
  }
  return d;
}

// Same value in closed form (operator.cuh: DiagGroup): exact integer arithmetic, one scaling.
__device__ __forceinline__ double diagonal_closed_form(uint64_t s, const DiagGroup *s_groups, int n_groups, long long c0, int scale) {
  long long acc = c0;
  for (int g = 0; g < n_groups; ++g) {
    const DiagGroup grp = s_groups[g];
    acc += static_cast<long long>(grp.weight) * __popcll(s & (s >> grp.shift) & grp.sites);
  }
  return scalbn(static_cast<double>(acc), -scale);
}

__host__ __device__ inline FxLayout fx_layout(int n_moves, int n_words, int n_groups, int n_diag, int planes_ok, int surv_slots) {
  FxLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    const uint32_t at = off;
    off += (bytes + 15u) & ~15u;
    return at;
  };
  L.cand = take(static_cast<uint32_t>(n_words) * 32u * 8u);  // {flip >> fshift, hash(flip)} per move, padded to whole words
  L.flip = take(static_cast<uint32_t>(n_moves) * 8u);
  L.coef = take(static_cast<uint32_t>(n_moves) * 8u);
  L.desc = take(planes_ok ? static_cast<uint32_t>(n_words) * 128u : 0u);
  L.mask = take(planes_ok ? 0u : static_cast<uint32_t>(n_moves) * 8u);
  L.need = take(planes_ok ? 0u : static_cast<uint32_t>(n_moves) * 8u);
  L.groups = take(n_groups >= 0 ? static_cast<uint32_t>(n_groups) * 16u : 0u);
  L.diag = take(n_groups >= 0 ? 0u : static_cast<uint32_t>(n_diag) * static_cast<uint32_t>(sizeof(DiagBond)));
  L.tables = off;
  off = 0;
  const uint32_t surv_bytes = static_cast<uint32_t>(surv_slots) * 64u;
  L.w_surv = take(surv_bytes > 256u ? surv_bytes : 256u);                              // | planes u32[64]
  L.w_amask = take(n_words * 128u > 384u ? static_cast<uint32_t>(n_words) * 128u : 384u);  // | abs_psi f64[32], row_off u32[32]
  L.w_pre = take(64u * 4u);  // exclusive prefix of survivor counts [32] | owner board [32]
  L.w_cnt = take(32u * 4u);
  L.w_pend = take(static_cast<uint32_t>(kFxLag) * 36u * 4u);  // per waiting tile: packed row counts [32], tile lo/hi, list length
  L.per_warp = off;
  return L;
}

#ifndef ASP_FX_MIN_CTAS
#define ASP_FX_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(kFxThreads, ASP_FX_MIN_CTAS) extract_csr_kernel(const FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FxLayout &L = a.layout;
  uint2 *s_cand = reinterpret_cast<uint2 *>(smem_raw + L.cand);
  uint64_t *s_flip = reinterpret_cast<uint64_t *>(smem_raw + L.flip);
  double *s_coef = reinterpret_cast<double *>(smem_raw + L.coef);
  uint32_t *s_desc = reinterpret_cast<uint32_t *>(smem_raw + L.desc);  // site i | site j << 8 | need_i << 16 | need_j << 17 | valid << 18
  uint64_t *s_mask = reinterpret_cast<uint64_t *>(smem_raw + L.mask);
  uint64_t *s_need = reinterpret_cast<uint64_t *>(smem_raw + L.need);
  DiagGroup *s_groups = reinterpret_cast<DiagGroup *>(smem_raw + L.groups);
  DiagBond *s_diag = reinterpret_cast<DiagBond *>(smem_raw + L.diag);
  uint32_t lane = threadIdx.x & 31;
  // opaque to the compiler: the lane index stays in its register instead of being re-derived from the thread index all over
  // the tile loop (under the 40-register cap that re-derivation cost spills: 1.735 -> 1.68 ms per call on the bench workload)
  asm volatile("" : "+r"(lane));
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned char *const wbase = smem_raw + L.tables + L.per_warp * warp;
  uint32_t *const w_planes = reinterpret_cast<uint32_t *>(wbase + L.w_surv);
  uint16_t *const w_surv = reinterpret_cast<uint16_t *>(wbase + L.w_surv);
  uint32_t *const w_amask = reinterpret_cast<uint32_t *>(wbase + L.w_amask);
  double *const w_abs_psi = reinterpret_cast<double *>(wbase + L.w_amask);
  uint32_t *const w_row_off = reinterpret_cast<uint32_t *>(wbase + L.w_amask + 256);
  uint32_t *const w_pre = reinterpret_cast<uint32_t *>(wbase + L.w_pre);
  uint32_t *const w_cnt = reinterpret_cast<uint32_t *>(wbase + L.w_cnt);
  uint32_t *const w_pend = reinterpret_cast<uint32_t *>(wbase + L.w_pend);

  for (int k = threadIdx.x; k < a.n_words * 32; k += kFxThreads) {
    uint2 cand = make_uint2(0u, 0u);
    uint32_t desc = 0;
    if (k < a.n_moves) {
      const Move mv = a.moves[k];
      s_flip[k] = mv.flip;
      s_coef[k] = mv.coef;
      cand = make_uint2(static_cast<uint32_t>(mv.flip >> a.fshift), filter_hash(mv.flip));
      if (a.planes_ok) {
        const int i = __ffsll(static_cast<long long>(mv.mask)) - 1;
        const int j = 63 - __clzll(static_cast<long long>(mv.mask));
        desc = static_cast<uint32_t>(i) | (static_cast<uint32_t>(j) << 8) | (static_cast<uint32_t>((mv.need >> i) & 1) << 16) |
               (static_cast<uint32_t>((mv.need >> j) & 1) << 17) | (1u << 18);
      } else {
        s_mask[k] = mv.mask;
        s_need[k] = mv.need;
      }
    }
    s_cand[k] = cand;
    if (a.planes_ok) s_desc[k] = desc;
  }
  if (a.n_groups >= 0) {
    for (int k = threadIdx.x; k < a.n_groups; k += kFxThreads) s_groups[k] = a.groups[k];
  } else {
    for (int k = threadIdx.x; k < a.n_diag; k += kFxThreads) s_diag[k] = a.diag[k];
  }
  __syncthreads();  // tables loaded; from here on the warps run independently
  uint2 *const my_lists = a.scratch + (static_cast<size_t>(blockIdx.x) * kFxWarps + warp) * (kFxLag + 1) * a.scratch_per_warp;  // kFxLag + 1 hit lists
  const uint32_t slots = static_cast<uint32_t>(a.surv_slots);
  const uint32_t cand_base = smem_addr(s_cand);
  const uint32_t amask_base = smem_addr(w_amask) + lane * 4u;
  const uint32_t surv_base = smem_addr(w_surv) + lane * 2u;
  const uint32_t surv_limit = surv_base + (slots >= 8u ? slots - 4u : 0u) * 64u;  // fill level checked every 4 candidates

  // Lag: a warp runs stages A-C of its next kFxLag tiles before it fetches the CSR offset of a
  // tile it has counted.  The offset needs every earlier tile's count, and a warp that finished
  // early would otherwise spin for the slowest of its predecessors (the SM's warp arbiter is not
  // fair, so some warps ARE much slower); by the time the next tiles are counted they are done.
  // (Counts are published right after stage C and never wait for anything, so the scheme cannot
  // deadlock.)  Waiting tiles keep their hit list in one of kFxLag + 1 scratch lists and their row
  // counts in shared memory.
  uint32_t taken = 0, waiting = 0;  // tiles this warp has counted / of those, not yet written (warp-uniform)
  for (;;) {
    // a tile = the 32 rows of one warp; tiles are handed out in order by an atomic ticket
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.ticket, 1u);
    const uint64_t tile = __shfl_sync(0xffffffffu, ticket, 0);
    const bool have_tile = tile < a.num_tiles;
    if (!have_tile && waiting == 0) break;
    uint2 *const my_list = my_lists + static_cast<size_t>(taken % (kFxLag + 1)) * a.scratch_per_warp;
    uint32_t packed_cnt = 0, list_count = 0;
    if (have_tile) {
    const uint64_t r = tile * kFxTileRows + lane;
    const bool live = r < a.num_rows;
    const uint64_t row = a.row_begin + r;
    const uint64_t s = live ? __ldg(&a.spins[row]) : 0ull;
    const uint32_t s_lo = static_cast<uint32_t>(s), s_hi = static_cast<uint32_t>(s >> 32);
    const bool generates = live && (s & ~a.state_mask) == 0;  // keys wider than the word have no images in the basis

    // =========================== A: which moves apply to which row ===========================
    if (a.planes_ok) {
      const uint32_t gen_mask = __ballot_sync(0xffffffffu, generates);
      w_planes[lane] = transpose32(s_lo, lane);
      w_planes[32 + lane] = transpose32(s_hi, lane);
      __syncwarp();
      for (int w = 0; w < a.n_words; ++w) {
        const uint32_t desc = s_desc[w * 32 + lane];
        const uint32_t pi = w_planes[desc & 63u], pj = w_planes[(desc >> 8) & 63u];
        const uint32_t xi = ((desc >> 16) & 1u) - 1u, xj = ((desc >> 17) & 1u) - 1u;  // need 1 -> 0, need 0 -> ~0
        uint32_t app = (pi ^ xi) & (pj ^ xj) & gen_mask;                              // rows this lane's move applies to
        if (!(desc & (1u << 18))) app = 0;
        w_amask[w * 32 + lane] = transpose32(app, lane);                              // moves that apply to this lane's row
      }
    } else {
      for (int w = 0; w < a.n_words; ++w) {
        uint32_t bits = 0;
        const int m_end = min(32, a.n_moves - w * 32);
        for (int k = 0; k < m_end; ++k)
          if ((s & s_mask[w * 32 + k]) == s_need[w * 32 + k]) bits |= 1u << k;
        w_amask[w * 32 + lane] = generates ? bits : 0u;
      }
    }
    w_cnt[lane] = 0;
    __syncwarp();  // planes are dead from here: their bytes become the survivor slots

    // =========================== B + C: sieve, search, record ================================
    uint32_t surv_addr = surv_base;   // next free survivor slot of this lane

    // C: deal the waiting survivors evenly over the lanes, search them exactly, record the hits
    auto flush = [&]() {
      const uint32_t h = (surv_addr - surv_base) >> 6;
      uint32_t incl = h;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      const uint32_t excl = incl - h;
      w_pre[lane] = excl;
      uint32_t carry = 0;  // lane that owns the survivor just before this batch
      for (uint32_t base = 0; base < total; base += 32) {
        // the lane whose survivors include number base + lane: lanes that START inside the batch
        // leave their index at the start position, everyone looks up the nearest start at or below
        const uint32_t start = excl - base;
        const bool starts_here = h != 0 && start < 32u;
        const uint32_t heads = __reduce_or_sync(0xffffffffu, starts_here ? 1u << start : 0u);
        if (starts_here) w_pre[32 + start] = lane;  // w_pre[32..] doubles as the owner board (entries 32..63)
        __syncwarp();
        const uint32_t e = base + lane;
        const bool valid = e < total;
        const uint32_t at_or_below = heads & (0xFFFFFFFFu >> (31u - lane));
        const uint32_t src = at_or_below ? w_pre[32 + (31 - __clz(at_or_below))] : carry;
        carry = __shfl_sync(0xffffffffu, src, 31);
        const uint32_t m = valid ? (static_cast<uint32_t>(w_surv[(e - w_pre[src]) * 32 + src]) - (cand_base >> 3)) & 0xFFFFu : 0u;
        const uint64_t s_src = (static_cast<uint64_t>(__shfl_sync(0xffffffffu, s_hi, src)) << 32) | __shfl_sync(0xffffffffu, s_lo, src);
        int32_t pos = -1;
        if (valid) pos = search_one(a, s_src ^ s_flip[m]);
        const bool hit = pos >= 0;
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (hits) {
          const bool below = static_cast<int>(m) < a.n_down;
          const uint32_t belows = __ballot_sync(0xffffffffu, hit && below);
          if (hit) {
            const uint32_t peers = __match_any_sync(hits, src);  // hits of the same row in this batch (ascending move order by lane)
            const uint32_t packed = w_cnt[src];
            __syncwarp(hits);
            const uint32_t before = __popc(peers & lt_mask);
            if (before == 0) w_cnt[src] = packed + __popc(peers) + (__popc(peers & belows) << 16);
            const uint32_t rank = (packed & 0xFFFFu) + before + (below ? 0u : 1u);  // the diagonal sits after the negative deltas
            my_list[list_count + __popc(hits & lt_mask)] = make_uint2(static_cast<uint32_t>(pos), m | (src << 11) | (rank << 16));
          }
          list_count += __popc(hits);
        }
        __syncwarp();
      }
      surv_addr = surv_base;
    };

    {
      // B: every lane walks the set bits of its own mask words.  A candidate needs its filter
      // word (index = (s >> fshift) ^ (flip >> fshift)) and its hash (= hash(s) ^ hash(flip)):
      // one 8-byte table read and two XORs.
      const uint32_t s_idx = static_cast<uint32_t>(s >> a.fshift), s_hash = filter_hash(s);
      uint32_t amask_addr = amask_base, tab_addr = cand_base;
      const uint32_t amask_last = amask_base + (a.n_words > 0 ? a.n_words - 1 : 0) * 128u;
      uint32_t cur = a.n_words ? lds_u32(amask_addr) : 0u;
      // first half of a step: next set bit -> table entry -> filter word on its way
      auto fetch = [&](uint32_t &entry_addr, uint32_t &hsh, uint2 &word) {  // harmless for a lane that has run out of moves
        if (cur == 0 && amask_addr < amask_last) {  // at most one word per step
          amask_addr += 128;
          tab_addr += 256;
          cur = lds_u32(amask_addr);
        }
        const bool act = cur != 0;
        entry_addr = tab_addr + ((static_cast<uint32_t>(__ffs(static_cast<int>(cur)) - 1) & 31u) << 3);
        cur &= cur - 1;
        const uint2 entry = lds_table_u2(entry_addr);
        hsh = s_hash ^ entry.y;
        word = make_uint2(0u, 0u);
        if (act) word = ldg_filter_u2(a.filter + (s_idx ^ entry.x));
      };
      // second half: both hashed bits set -> the candidate survives into this lane's slots
      auto sieve = [&](uint32_t entry_addr, uint32_t hsh, uint2 word) {
        if (__funnelshift_r(word.x, 0u, hsh) & __funnelshift_r(word.y, 0u, hsh >> 8) & 1u) {
          sts_u16(surv_addr, entry_addr >> 3);
          surv_addr += 64;
        }
      };
      auto step = [&]() {
        uint32_t e0, h0;
        uint2 w0;
        fetch(e0, h0, w0);
        sieve(e0, h0, w0);
      };
      if (slots >= 8u) {
        // four candidates per lane between two looks at the loop condition and the fill level,
        // their four filter words in flight together
        while (__any_sync(0xffffffffu, cur != 0 || amask_addr < amask_last)) {
#if ASP_FX_IN_FLIGHT == 4
          uint32_t e0, e1, e2, e3, h0, h1, h2, h3;
          uint2 w0, w1, w2, w3;
          fetch(e0, h0, w0);
          fetch(e1, h1, w1);
          fetch(e2, h2, w2);
          fetch(e3, h3, w3);
          sieve(e0, h0, w0);
          sieve(e1, h1, w1);
          sieve(e2, h2, w2);
          sieve(e3, h3, w3);
#else
          uint32_t e0, e1, h0, h1;
          uint2 w0, w1;
          fetch(e0, h0, w0);
          fetch(e1, h1, w1);
          sieve(e0, h0, w0);
          sieve(e1, h1, w1);
          fetch(e0, h0, w0);
          fetch(e1, h1, w1);
          sieve(e0, h0, w0);
          sieve(e1, h1, w1);
#endif
          if (__any_sync(0xffffffffu, surv_addr > surv_limit)) {
            __syncwarp();
            flush();
          }
        }
      } else {  // tiny survivor lists (tests): flush after every survivor
        while (__any_sync(0xffffffffu, cur != 0 || amask_addr < amask_last)) {
          step();
          if (__any_sync(0xffffffffu, surv_addr != surv_base)) {
            __syncwarp();
            flush();
          }
        }
      }
      __syncwarp();
      flush();
    }
    packed_cnt = w_cnt[lane] + (live ? 1u : 0u);  // couplings of the row (diagonal included) | those below the diagonal << 16
    // publish the tile's count (tile 0: its inclusive prefix) -- never waits
    uint32_t total = packed_cnt & 0xFFFFu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) {
      if (tile == 0)
        st_status(&a.status[0], kFlagPrefix | ((a.base_in ? *a.base_in : 0ull) + total));
      else
        st_status(&a.status[tile], kFlagAggregate | total);
    }
    }  // have_tile

    if (waiting == kFxLag || (!have_tile && waiting != 0)) {
      // ======================= D: CSR offset of the oldest waiting tile by decoupled look-back
      const uint32_t oldest = taken - waiting;
      const uint32_t *const pend = w_pend + (oldest % kFxLag) * 36;
      const uint32_t pend_packed = pend[lane], pend_count = pend[34];
      const uint64_t ptile = (static_cast<uint64_t>(pend[33]) << 32) | pend[32];
      const uint64_t r = ptile * kFxTileRows + lane;
      const bool live = r < a.num_rows;
      const uint64_t row = a.row_begin + r;
      const uint32_t my_cnt = pend_packed & 0xFFFFu, down_cnt = pend_packed >> 16;
      uint32_t incl = my_cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t tile_total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long exclusive = 0;
      if (ptile == 0) {
        if (a.base_in) exclusive = *a.base_in;
      } else {
        int64_t look = static_cast<int64_t>(ptile) - 1;  // window [look - 31, look]
        for (;;) {
          const int64_t idx = look - lane;
          unsigned long long st = kFlagPrefix;  // virtual tiles before tile 0: prefix 0
          if (idx >= 0) {
            unsigned backoff = ASP_FX_BACKOFF_NS;
            while (((st = ld_status(&a.status[idx])) >> 62) == 0) {
              __nanosleep(backoff);
              if (backoff < 32 * ASP_FX_BACKOFF_NS) backoff <<= 1;
            }
          }
          const uint32_t has_prefix = __ballot_sync(0xffffffffu, (st >> 62) == 2);
          const uint32_t first = has_prefix ? static_cast<uint32_t>(__ffs(has_prefix)) - 1u : 32u;  // nearest tile with a prefix
          // counts of the tiles in front of it (each < 2^16) add up in 32 bits; the prefix itself is 64-bit
          uint32_t counts = lane < first ? static_cast<uint32_t>(st) : 0u;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) counts += __shfl_xor_sync(0xffffffffu, counts, o);
          exclusive += counts;
          if (has_prefix) {
            const uint32_t p_lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(st), first);
            const uint32_t p_hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(st >> 32), first);
            exclusive += ((static_cast<unsigned long long>(p_hi) << 32) | p_lo) & kValueMask;
            break;
          }
          look -= 32;
        }
        if (lane == 0) st_status(&a.status[ptile], kFlagPrefix | (exclusive + tile_total));
      }
      if (lane == 0 && ptile == a.num_tiles - 1) {
        a.indptr[a.num_rows] = static_cast<int64_t>(exclusive + tile_total);
        *a.nnz_out = exclusive + tile_total;
        if (a.nnz_mirror) {
          *a.nnz_mirror = exclusive + tile_total;
          __threadfence_system();
        }
      }
      const uint64_t tile_base = exclusive;

      // ======================= write the previous tile's CSR rows ============================
      // (the apply masks are dead: their bytes now hold the row offsets and amplitudes)
      const uint2 *const list = my_lists + static_cast<size_t>(oldest % (kFxLag + 1)) * a.scratch_per_warp;
      const uint32_t my_off = incl - my_cnt;  // row start relative to the tile
      const uint64_t s = live ? __ldg(&a.spins[row]) : 0ull;
      const double a_i = live ? fabs(__ldg(&a.psi[row])) : 0.0;
      w_row_off[lane] = my_off;
      w_abs_psi[lane] = a_i;
      if (live) a.indptr[r] = static_cast<int64_t>(tile_base + my_off);
      __syncwarp();
      for (uint32_t k = lane; k < pend_count; k += 32) {
        const uint2 entry = list[k];
        const uint32_t pos = entry.x, m = entry.y & 0x7FFu, src = (entry.y >> 11) & 31u, rank = entry.y >> 16;
        const uint64_t dest = tile_base + w_row_off[src] + rank;
        if (dest < a.capacity) {
          a.indices[dest] = static_cast<int32_t>(pos);
          a.data[dest] = __dmul_rn(__dmul_rn(s_coef[m], fabs(ldg_stream_f64(&a.psi[pos]))), w_abs_psi[src]);  // (c |psi_j|) |psi_i|: common.py:71-82
        }
      }
      if (live) {
        const double d = a.n_groups >= 0 ? diagonal_closed_form(s, s_groups, a.n_groups, a.diag_c0, a.diag_scale)
                                         : diagonal_element(s, s_diag, a.n_diag);
        const uint64_t dest = tile_base + my_off + down_cnt;
        if (dest < a.capacity) {
          a.indices[dest] = static_cast<int32_t>(row);
          a.data[dest] = __dmul_rn(__dmul_rn(d, a_i), a_i);
        }
      }
      __syncwarp();  // the row offsets are read by all lanes; the next tile overwrites their bytes
      --waiting;
    }
    if (have_tile) {  // the tile just counted joins the queue
      uint32_t *const pend = w_pend + (taken % kFxLag) * 36;
      pend[lane] = packed_cnt;
      if (lane == 0) {
        pend[32] = static_cast<uint32_t>(tile);
        pend[33] = static_cast<uint32_t>(tile >> 32);
        pend[34] = list_count;
      }
      __syncwarp();
      ++taken;
      ++waiting;
    }
  }
}

constexpr int kFxMaxChunks = kFusedMaxChunks;  // row chunks of one pipelined host call

struct FusedWorkspace {
  uint32_t *starts;
  uint2 *filter;
  unsigned long long *status;  // [num_tiles + kFxMaxChunks]: every row chunk has its own slice
  unsigned int *tickets;       // [kFxMaxChunks] one per chunk, 64 B apart
  unsigned long long *totals;  // [kFxMaxChunks] running totals
  uint2 *scratch;
  uint64_t num_buckets, num_words;
  int tshift, fshift;
  uint32_t scratch_per_warp, scratch_ctas;
  size_t bytes, zero_offset, zero_bytes;
};

static int g_surv_entries_override = 0;
// optional CUDA-event bracket around the extraction kernel alone (bench.py's roofline figure)
static bool g_time_kernel = false;
constexpr int kEvRing = 64;  // the last kEvRing launches keep their event pair
static cudaEvent_t g_ev_begin[kEvRing] = {}, g_ev_end[kEvRing] = {};
static uint64_t g_ev_launches = 0;
static int g_filter_bits_delta = 0, g_table_bits_delta = 0, g_stage_a_mode = 0;
static int g_gather_mode = 2;  // asp_gather_index: 2 = one TMA kernel (default), 1 = one kernel with plain loads, 0 = copy engines + per-block index kernels

static FusedWorkspace carve_fused(void *base, const asp_operator *op, uint64_t n_total, uint64_t num_rows) {
  FusedWorkspace w;
  int lg = 0;
  while ((1ull << lg) < n_total) ++lg;
  const int key_bits = static_cast<int>(op->number_spins);
  int tbits = lg - 1 + g_table_bits_delta;  // about two slots per key: most buckets hold 0 or 1 keys
  tbits = std::max(4, std::min(tbits, 26));
  tbits = std::min(tbits, key_bits);
  w.tshift = key_bits - tbits;
  w.num_buckets = 1ull << tbits;
  int fbits = lg - 1 + g_filter_bits_delta;  // 8 bytes per 2 keys: ~1 % false positives
  fbits = std::max(4, std::min(fbits, 27));
  fbits = std::min(fbits, key_bits);
  w.fshift = key_bits - fbits;
  w.num_words = 1ull << fbits;
  const uint64_t tiles = (num_rows + kFxTileRows - 1) / kFxTileRows + kFxMaxChunks;
  w.scratch_per_warp = std::max<uint32_t>(32u * static_cast<uint32_t>(op->moves.size()), 32u);
  const size_t per_cta = static_cast<size_t>(w.scratch_per_warp) * (kFxLag + 1) * kFxWarps * sizeof(uint2);  // kFxLag + 1 hit lists per warp
  uint64_t ctas = std::min<uint64_t>(static_cast<uint64_t>(kNumSMs) * kFxMaxCtasPerSM, std::max<uint64_t>(tiles, 1));
  ctas = std::min<uint64_t>(ctas, std::max<uint64_t>(kFxScratchBudget / per_cta, kNumSMs));
  w.scratch_ctas = static_cast<uint32_t>(ctas);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void *p = base ? static_cast<char *>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.starts = static_cast<uint32_t *>(take((w.num_buckets + 1) * sizeof(uint32_t)));
  w.scratch = static_cast<uint2 *>(take(per_cta * ctas));
  w.zero_offset = off;
  w.filter = static_cast<uint2 *>(take(w.num_words * sizeof(uint2)));
  w.status = static_cast<unsigned long long *>(take(tiles * sizeof(unsigned long long)));
  w.tickets = static_cast<unsigned int *>(take(kFxMaxChunks * 64));
  w.totals = static_cast<unsigned long long *>(take(kFxMaxChunks * sizeof(unsigned long long)));
  w.zero_bytes = off - w.zero_offset;
  w.bytes = off;
  return w;
}

// The index asp_gather_index leaves in a workspace is single-use: the extraction consumes the tickets and the
// look-back words.  The library remembers which workspaces hold a fresh index so that a second extraction
// without a new asp_gather_index is refused instead of silently writing nothing.
static std::mutex g_indexed_mu;
static std::unordered_set<const void *> g_indexed;
void fused_mark_indexed(const void *workspace) {
  std::lock_guard<std::mutex> lock(g_indexed_mu);
  g_indexed.insert(workspace);
}
bool fused_consume_indexed(const void *workspace) {
  std::lock_guard<std::mutex> lock(g_indexed_mu);
  return g_indexed.erase(workspace) > 0;
}

size_t fused_workspace_bytes(const asp_operator *op, uint64_t n_total, uint64_t num_rows) {
  return carve_fused(nullptr, op, n_total, num_rows).bytes;
}

int fused_check_operator(const asp_operator *op) {
  ASP_REQUIRE(op != nullptr, "operator is NULL");
  ASP_REQUIRE(op->d_moves != nullptr, "operator has no device mirror (created without a CUDA device)");
  if (!op->sorted_emitter() || op->moves.size() > kFxMaxMoves) {
    set_error("fused extraction needs an unsymmetrised operator with distinct moves (at most %u); use the apply + build_matrix + canonicalise path", kFxMaxMoves);
    return ASP_ERR_UNSUPPORTED;
  }
  return ASP_OK;
}

// Zero the look-back state and index the sorted basis (once per call, before the chunks).
int fused_prepare(const asp_operator *op, uint64_t n_total, const uint64_t *d_spins, uint64_t num_rows, void *d_workspace,
                  size_t workspace_bytes, cudaStream_t s) {
  FusedWorkspace w = carve_fused(d_workspace, op, n_total, num_rows);
  if (d_workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return ASP_ERR_WORKSPACE;
  }
  ASP_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(d_workspace) + w.zero_offset, 0, w.zero_bytes, s));
  // two keys per thread: neighbours that share a filter word share one 64-bit atomic (the pass is bound by L2 atomics)
  index_block_kernel<<<static_cast<unsigned>((n_total + 511) / 512), 256, 0, s>>>(d_spins, static_cast<uint32_t>(n_total), 0u,
                                                                                 static_cast<uint32_t>(n_total), op->state_mask, w.tshift,
                                                                                 w.num_buckets, w.starts, w.fshift, w.filter);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

// X1 + index in one kernel: see gather_index_kernel.  Host arrays: shard_begin[world+1],
// d_shard_spins/d_shard_psi[world] (device pointers valid on this device: own memory or IPC-mapped).
int fused_prepare_gather(const asp_operator *op, uint32_t world, uint32_t rank, const uint64_t *shard_begin,
                         const uint64_t *const *d_shard_spins, const double *const *d_shard_psi, const uint64_t *d_ready,
                         uint64_t epoch, uint64_t *d_spins, double *d_psi, uint64_t num_rows, void *d_workspace,
                         size_t workspace_bytes, cudaStream_t s, bool tma) {
  ASP_REQUIRE(world >= 1 && world <= static_cast<uint32_t>(kGxMaxRanks) && rank < world, "world size must be in 1..16");
  const uint64_t n_total = shard_begin[world];
  ASP_REQUIRE(shard_begin[0] == 0, "shard_begin[0] must be 0");
  const bool copy_only = op == nullptr;  // asp_gather_blocks: gather_copy_tma_kernel, no index
  FusedWorkspace w{};
  if (!copy_only) {
    w = carve_fused(d_workspace, op, n_total, num_rows);
    if (d_workspace == nullptr || workspace_bytes < w.bytes) {
      set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
      return ASP_ERR_WORKSPACE;
    }
  }
  GatherArgs a{};
  a.world = static_cast<int>(world);
  uint64_t units = 0;
  for (uint32_t k = 0; k < world; ++k) {
    const uint32_t q = (rank + 1 + k) % world;  // own block last
    ASP_REQUIRE(shard_begin[q + 1] >= shard_begin[q], "shard_begin must be non-decreasing");
    a.order[k] = static_cast<int>(q);
    a.unit_begin[k] = units;
    units += (shard_begin[q + 1] - shard_begin[q] + 1) / 2;
  }
  a.unit_begin[world] = units;
  uint64_t chunks = 0;
  for (uint32_t k = 0; k < world; ++k) {
    a.chunk_begin[k] = chunks;
    chunks += (shard_begin[a.order[k] + 1] - shard_begin[a.order[k]] + 1023) / 1024;
  }
  a.chunk_begin[world] = chunks;
  for (uint32_t q = 0; q < world; ++q) {
    ASP_REQUIRE(shard_begin[q + 1] == shard_begin[q] || (d_shard_spins[q] && d_shard_psi[q]), "NULL shard pointer");
    ASP_REQUIRE((reinterpret_cast<uintptr_t>(d_shard_spins[q]) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_shard_psi[q]) & 15u) == 0,
                "shard buffers must be 16-byte aligned");
    a.shard_spins[q] = d_shard_spins[q];
    a.shard_psi[q] = d_shard_psi[q];
    a.begin[q] = shard_begin[q];
  }
  a.begin[world] = n_total;
  a.ready = reinterpret_cast<const unsigned long long *>(d_ready);
  a.epoch = epoch;
  a.spins = d_spins;
  a.psi = d_psi;
  a.n = static_cast<uint32_t>(n_total);
  if (copy_only) {
    ASP_CUDA_CHECK(cudaFuncSetAttribute(gather_copy_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCxSmem)));
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(chunks, static_cast<uint64_t>(kNumSMs)));
    gather_copy_tma_kernel<<<grid, 32, kCxSmem, s>>>(a);
    ASP_LAUNCH_CHECK();
    return ASP_OK;
  }
  a.state_mask = op->state_mask;
  a.num_buckets = w.num_buckets;
  a.tshift = w.tshift;
  a.fshift = w.fshift;
  a.starts = w.starts;
  a.filter = w.filter;
  ASP_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(d_workspace) + w.zero_offset, 0, w.zero_bytes, s));
  if (tma) {
    a.stages = kTxStages;
    const size_t smem = kTxSmem;
    ASP_CUDA_CHECK(cudaFuncSetAttribute(gather_index_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTxSmem)));
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(chunks, static_cast<uint64_t>(kNumSMs) * kTxCtasPerSM));
    gather_index_tma_kernel<<<grid, kTxThreads, smem, s>>>(a);
  } else {
    const uint64_t per_cta = static_cast<uint64_t>(kGxThreads) * kGxUnroll;
    gather_index_kernel<<<static_cast<unsigned>((units + per_cta - 1) / per_cta), kGxThreads, 0, s>>>(a);
  }
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

// X1 with the copy engines: per block (own block first, then ring order) [wait for its ready flag ->
// cudaMemcpyAsync keys + amplitudes into the private full copy] on a copy stream; on the caller's stream
// the block is indexed as soon as its copies are done, while the next blocks travel.
constexpr int kCopyStreamsMax = 8;
struct CopyLane {
  cudaStream_t copy[kCopyStreamsMax] = {};  // one copy engine each; the blocks of a gather are dealt over them
  cudaEvent_t start = nullptr, done[kCopyStreamsMax] = {}, block[kGxMaxRanks] = {};
  int device = -1;
};
static CopyLane g_lane;
static int g_copy_streams = 2;  // copy streams a gather uses (asp_set_copy_streams): blocks in flight at a time (8 GPUs: 2 -> 3.05 ms per step, 4 -> 3.60, 7 -> 3.42)

int fused_prepare_gather_ce(const asp_operator *op, uint32_t world, uint32_t rank, const uint64_t *shard_begin,
                            const uint64_t *const *d_shard_spins, const double *const *d_shard_psi, const uint64_t *d_ready,
                            uint64_t epoch, uint64_t *d_spins, double *d_psi, uint64_t num_rows, void *d_workspace,
                            size_t workspace_bytes, cudaStream_t s) {
  ASP_REQUIRE(world >= 1 && world <= static_cast<uint32_t>(kGxMaxRanks) && rank < world, "world size must be in 1..16");
  const uint64_t n_total = shard_begin[world];
  ASP_REQUIRE(shard_begin[0] == 0, "shard_begin[0] must be 0");
  const bool index = op != nullptr;  // op == NULL: copies only (asp_gather_blocks), the caller indexes later
  FusedWorkspace w{};
  if (index) {
    w = carve_fused(d_workspace, op, n_total, num_rows);
    if (d_workspace == nullptr || workspace_bytes < w.bytes) {
      set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
      return ASP_ERR_WORKSPACE;
    }
  }
  int dev = 0;
  ASP_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_lane.device != dev) {
    ASP_REQUIRE(g_lane.device < 0, "one device per process");
    for (auto &c : g_lane.copy) ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
    ASP_CUDA_CHECK(cudaEventCreateWithFlags(&g_lane.start, cudaEventDisableTiming));
    for (auto &e : g_lane.done) ASP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : g_lane.block) ASP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_lane.device = dev;
  }
  const int streams = std::max(1, std::min(g_copy_streams, kCopyStreamsMax));
  if (index) ASP_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(d_workspace) + w.zero_offset, 0, w.zero_bytes, s));
  ASP_CUDA_CHECK(cudaEventRecord(g_lane.start, s));  // the copies may overwrite the full copy only after the caller's earlier work
  for (int c = 0; c < streams; ++c) ASP_CUDA_CHECK(cudaStreamWaitEvent(g_lane.copy[c], g_lane.start, 0));
  SeamArgs seams{};
  for (uint32_t k = 0; k < world; ++k) {
    const uint32_t q = (rank + k) % world;  // own block first (no flag to wait for), then ring order
    ASP_REQUIRE(shard_begin[q + 1] >= shard_begin[q], "shard_begin must be non-decreasing");
    const uint64_t b0 = shard_begin[q], len = shard_begin[q + 1] - b0;
    seams.first[q] = len > 0 ? static_cast<uint32_t>(b0) : static_cast<uint32_t>(n_total);
    if (len == 0) continue;
    ASP_REQUIRE(d_shard_spins[q] && d_shard_psi[q], "NULL shard pointer");
    // a block travels on ONE stream (keys, then amplitudes); the blocks are dealt over the streams, so several peers
    // are pulled at a time, each by its own copy engine
    cudaStream_t lane = g_lane.copy[k % streams];
    if (d_ready != nullptr && q != rank) {
      wait_one_flag_kernel<<<1, 1, 0, lane>>>(reinterpret_cast<const unsigned long long *>(d_ready) + q, epoch);
      ASP_LAUNCH_CHECK();
    }
    ASP_CUDA_CHECK(cudaMemcpyAsync(d_spins + b0, d_shard_spins[q], len * sizeof(uint64_t), cudaMemcpyDeviceToDevice, lane));
    if (index) {  // the keys of this block have landed: index them while its amplitudes and the other blocks travel
      ASP_CUDA_CHECK(cudaEventRecord(g_lane.block[k], lane));
      ASP_CUDA_CHECK(cudaStreamWaitEvent(s, g_lane.block[k], 0));
      index_block_kernel<<<static_cast<unsigned>((len + 511) / 512), 256, 0, s>>>(d_spins, static_cast<uint32_t>(n_total), static_cast<uint32_t>(b0),
                                                                                 static_cast<uint32_t>(b0 + len), op->state_mask, w.tshift,
                                                                                 w.num_buckets, w.starts, w.fshift, w.filter);
      ASP_LAUNCH_CHECK();
    }
    ASP_CUDA_CHECK(cudaMemcpyAsync(d_psi + b0, d_shard_psi[q], len * sizeof(double), cudaMemcpyDeviceToDevice, lane));
  }
  // everything (amplitudes included) has landed before the caller's stream goes on
  for (int c = 0; c < streams; ++c) {
    ASP_CUDA_CHECK(cudaEventRecord(g_lane.done[c], g_lane.copy[c]));
    ASP_CUDA_CHECK(cudaStreamWaitEvent(s, g_lane.done[c], 0));
  }
  if (index) {
    index_seam_kernel<<<1, 32, 0, s>>>(d_spins, static_cast<uint32_t>(n_total), seams, static_cast<int>(world), op->state_mask, w.tshift,
                                       w.num_buckets, w.starts);
    ASP_LAUNCH_CHECK();
  }
  return ASP_OK;
}

// Rows [row_begin + chunk_begin, +chunk_rows) of the block that starts at row_begin: chunk
// number `chunk` (< kFxMaxChunks) of a call whose earlier chunks cover [0, chunk_begin).
// d_indptr is the BLOCK's indptr; offsets continue from the previous chunk's total.
// The running total lands in the workspace (fused_total()) and, when nnz_mirror != NULL, in
// that (mapped host) location too.
int fused_launch(const asp_operator *op, uint64_t n_total, const uint64_t *d_spins, const double *d_psi, uint64_t row_begin,
                 uint64_t num_rows, int chunk, uint64_t chunk_begin, uint64_t chunk_rows, void *d_workspace, uint64_t capacity,
                 int64_t *d_indptr, int32_t *d_indices, double *d_data, unsigned long long *nnz_mirror, cudaStream_t s) {
  ASP_REQUIRE(chunk >= 0 && chunk < kFxMaxChunks, "too many row chunks");
  ASP_REQUIRE(chunk_begin % kFxTileRows == 0, "row chunks start at tile boundaries");
  FusedWorkspace w = carve_fused(d_workspace, op, n_total, num_rows);
  FusedArgs a{};
  a.spins = d_spins;
  a.psi = d_psi;
  a.starts = w.starts;
  a.tshift = w.tshift;
  a.filter = w.filter;
  a.fshift = w.fshift;
  a.state_mask = op->state_mask;
  a.n_total = static_cast<uint32_t>(n_total);
  a.row_begin = row_begin + chunk_begin;
  a.num_rows = chunk_rows;
  a.num_tiles = (chunk_rows + kFxTileRows - 1) / kFxTileRows;
  a.moves = op->d_moves;
  a.n_moves = static_cast<int>(op->moves.size());
  a.n_down = static_cast<int>(op->n_down);
  a.n_words = (a.n_moves + 31) / 32;
  a.diag = op->d_diag;
  a.n_diag = static_cast<int>(op->diag.size());
  const bool closed_form = op->diag_scale >= 0 && g_stage_a_mode != 1;
  a.groups = op->d_diag_groups;
  a.n_groups = closed_form ? static_cast<int>(op->diag_groups.size()) : -1;
  a.diag_scale = op->diag_scale;
  a.diag_c0 = op->diag_c0;
  int slots = kFxSurvSlotsDefault;
  if (g_surv_entries_override > 0) slots = std::max(1, g_surv_entries_override / 32);
  a.surv_slots = slots;
  bool two_bit = true;
  for (const Move &mv : op->moves) two_bit = two_bit && __builtin_popcountll(mv.mask) == 2;
  a.planes_ok = (two_bit && g_stage_a_mode != 1) ? 1 : 0;
  a.scratch = w.scratch;
  a.scratch_per_warp = w.scratch_per_warp;
  a.status = w.status + chunk_begin / kFxTileRows + chunk;
  a.ticket = w.tickets + 16 * chunk;
  a.capacity = capacity;
  a.indptr = d_indptr + chunk_begin;
  a.indices = d_indices;
  a.data = d_data;
  a.nnz_out = w.totals + chunk;
  a.nnz_mirror = nnz_mirror;
  a.base_in = chunk == 0 ? nullptr : w.totals + (chunk - 1);
  const FxLayout layout = fx_layout(a.n_moves, a.n_words, a.n_groups, a.n_diag, a.planes_ok, slots);
  a.layout = layout;
  const size_t smem = layout.tables + static_cast<size_t>(layout.per_warp) * kFxWarps;
  ASP_REQUIRE(smem <= 200 * 1024, "operator too large for the fused kernel's shared-memory tables");
  ASP_CUDA_CHECK(cudaFuncSetAttribute(extract_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  ASP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extract_csr_kernel, kFxThreads, smem));
  ASP_REQUIRE(per_sm >= 1, "fused extraction kernel does not fit on an SM");
  const uint64_t resident = static_cast<uint64_t>(kNumSMs) * std::min(per_sm, kFxMaxCtasPerSM);
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(std::min<uint64_t>(a.num_tiles, resident), w.scratch_ctas));
  const int ev_slot = static_cast<int>(g_ev_launches % kEvRing);
  if (g_time_kernel) {
    if (!g_ev_begin[ev_slot]) {
      ASP_CUDA_CHECK(cudaEventCreate(&g_ev_begin[ev_slot]));
      ASP_CUDA_CHECK(cudaEventCreate(&g_ev_end[ev_slot]));
    }
    ASP_CUDA_CHECK(cudaEventRecord(g_ev_begin[ev_slot], s));
  }
  extract_csr_kernel<<<grid, kFxThreads, smem, s>>>(a);
  ASP_LAUNCH_CHECK();
  if (g_time_kernel) {
    ASP_CUDA_CHECK(cudaEventRecord(g_ev_end[ev_slot], s));
    ++g_ev_launches;
  }
  return ASP_OK;
}

const unsigned long long *fused_total(const asp_operator *op, uint64_t n_total, uint64_t num_rows, void *d_workspace, int chunk) {
  return carve_fused(d_workspace, op, n_total, num_rows).totals + chunk;
}

}  // namespace asp

using namespace asp;

extern "C" {

void asp_debug_set_hit_list_capacity(int entries_per_warp) { g_surv_entries_override = entries_per_warp; }

void asp_debug_set_extract_tuning(int filter_bits_delta, int table_bits_delta, int stage_a_mode) {
  g_filter_bits_delta = filter_bits_delta;
  g_table_bits_delta = table_bits_delta;
  g_stage_a_mode = stage_a_mode;
}

void asp_debug_time_extract_kernel(int enable) { g_time_kernel = enable != 0; }

float asp_debug_extract_kernel_ms(int back) {
  float ms = -1.0f;
  if (back < 0 || back >= kEvRing || static_cast<uint64_t>(back) >= g_ev_launches) return -1.0f;
  const int slot = static_cast<int>((g_ev_launches - 1 - static_cast<uint64_t>(back)) % kEvRing);
  if (!g_ev_begin[slot] || cudaEventSynchronize(g_ev_end[slot]) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, g_ev_begin[slot], g_ev_end[slot]) != cudaSuccess) return -1.0f;
  return ms;
}

float asp_debug_last_extract_kernel_ms(void) { return asp_debug_extract_kernel_ms(0); }

size_t asp_extract_csr_workspace_bytes(asp_operator const *op, uint64_t n_total, uint64_t num_rows) {
  if (!op) return 0;
  return fused_workspace_bytes(op, n_total, num_rows);
}

static int extract_csr_impl(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, double const *d_psi,
                            uint64_t row_begin, uint64_t num_rows, void *d_workspace, size_t workspace_bytes,
                            uint64_t capacity, int64_t *d_indptr, int32_t *d_indices, double *d_data, uint64_t *h_nnz,
                            void *stream, bool indexed) {
  auto s = static_cast<cudaStream_t>(stream);
  int rc = fused_check_operator(op);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(n_total < (1ull << 31), "int32 column indices need n_total < 2^31 (scipy picks int32 the same way)");
  ASP_REQUIRE(row_begin + num_rows <= n_total, "row block exceeds the basis");
  ASP_REQUIRE(d_indptr != nullptr, "d_indptr is NULL");
  ASP_REQUIRE(capacity == 0 || (d_indices && d_data), "NULL output buffer");
  if (num_rows == 0 || n_total == 0) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t), s));
    if (h_nnz) {
      ASP_CUDA_CHECK(cudaStreamSynchronize(s));
      *h_nnz = 0;
    }
    return ASP_OK;
  }
  ASP_REQUIRE(d_spins && d_psi, "NULL input buffer");
  if (indexed) {  // the workspace was zeroed and indexed for this very (n_total, num_rows) by asp_gather_index
    ASP_REQUIRE(d_workspace != nullptr && workspace_bytes >= fused_workspace_bytes(op, n_total, num_rows), "workspace too small");
    ASP_REQUIRE(fused_consume_indexed(d_workspace),
                "the workspace holds no fresh index: call asp_gather_index before every indexed extraction (the index is single-use)");
  } else {
    rc = fused_prepare(op, n_total, d_spins, num_rows, d_workspace, workspace_bytes, s);
    if (rc != ASP_OK) return rc;
  }
  rc = fused_launch(op, n_total, d_spins, d_psi, row_begin, num_rows, 0, 0, num_rows, d_workspace, capacity, d_indptr, d_indices,
                    d_data, nullptr, s);
  if (rc != ASP_OK) return rc;
  if (h_nnz) {
    unsigned long long total = 0;
    ASP_CUDA_CHECK(cudaMemcpyAsync(&total, fused_total(op, n_total, num_rows, d_workspace, 0), sizeof(total), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    *h_nnz = total;
    if (total > capacity) {
      set_error("output capacity too small: %llu couplings, room for %llu (indptr is complete; call again with the larger capacity)",
                total, static_cast<unsigned long long>(capacity));
      return ASP_ERR_WORKSPACE;
    }
  }
  return ASP_OK;
}

int asp_extract_csr(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, double const *d_psi,
                    uint64_t row_begin, uint64_t num_rows, void *d_workspace, size_t workspace_bytes,
                    uint64_t capacity, int64_t *d_indptr, int32_t *d_indices, double *d_data, uint64_t *h_nnz,
                    void *stream) {
  return extract_csr_impl(op, n_total, d_spins, d_psi, row_begin, num_rows, d_workspace, workspace_bytes, capacity, d_indptr,
                          d_indices, d_data, h_nnz, stream, /*indexed=*/false);
}

int asp_gather_index(asp_operator const *op, uint32_t world, uint32_t rank, uint64_t const *shard_begin,
                     uint64_t const *const *d_shard_spins, double const *const *d_shard_psi, uint64_t const *d_ready,
                     uint64_t epoch, uint64_t *d_spins, double *d_psi, uint64_t num_rows, void *d_workspace,
                     size_t workspace_bytes, void *stream) {
  int rc = fused_check_operator(op);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(shard_begin && d_shard_spins && d_shard_psi, "NULL shard table");
  ASP_REQUIRE(world >= 1 && world <= 16, "world size must be in 1..16");
  ASP_REQUIRE(shard_begin[world] > 0 && shard_begin[world] < (1ull << 31), "the gathered basis needs 0 < n_total < 2^31");
  ASP_REQUIRE(d_spins && d_psi, "NULL output buffer");
  if (g_gather_mode != 0)
    rc = fused_prepare_gather(op, world, rank, shard_begin, d_shard_spins, d_shard_psi, d_ready, epoch, d_spins, d_psi, num_rows,
                              d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream), g_gather_mode == 2);
  else
    rc = fused_prepare_gather_ce(op, world, rank, shard_begin, d_shard_spins, d_shard_psi, d_ready, epoch, d_spins, d_psi, num_rows,
                                 d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  if (rc == ASP_OK) fused_mark_indexed(d_workspace);
  return rc;
}

int asp_gather_blocks(uint32_t world, uint32_t rank, uint64_t const *shard_begin, uint64_t const *const *d_shard_spins,
                      double const *const *d_shard_psi, uint64_t const *d_ready, uint64_t epoch, uint64_t *d_spins, double *d_psi,
                      void *stream) {
  ASP_REQUIRE(shard_begin && d_shard_spins && d_shard_psi, "NULL shard table");
  ASP_REQUIRE(world >= 1 && world <= 16, "world size must be in 1..16");
  ASP_REQUIRE(d_spins && d_psi, "NULL output buffer");
  if (g_gather_mode != 0)  // one thread per CTA drives bulk copies in both directions (fits beside a resident extraction)
    return fused_prepare_gather(nullptr, world, rank, shard_begin, d_shard_spins, d_shard_psi, d_ready, epoch, d_spins, d_psi, 0, nullptr, 0,
                                static_cast<cudaStream_t>(stream), true);
  return fused_prepare_gather_ce(nullptr, world, rank, shard_begin, d_shard_spins, d_shard_psi, d_ready, epoch, d_spins, d_psi, 0, nullptr,
                                 0, static_cast<cudaStream_t>(stream));
}

void asp_set_gather_mode(int mode) { g_gather_mode = (mode == 0 || mode == 1) ? mode : 2; }

void asp_set_copy_streams(int streams) { g_copy_streams = streams < 1 ? 1 : (streams > kCopyStreamsMax ? kCopyStreamsMax : streams); }

int asp_extract_csr_indexed(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, double const *d_psi,
                            uint64_t row_begin, uint64_t num_rows, void *d_workspace, size_t workspace_bytes,
                            uint64_t capacity, int64_t *d_indptr, int32_t *d_indices, double *d_data, uint64_t *h_nnz,
                            void *stream) {
  return extract_csr_impl(op, n_total, d_spins, d_psi, row_begin, num_rows, d_workspace, workspace_bytes, capacity, d_indptr,
                          d_indices, d_data, h_nnz, stream, /*indexed=*/true);
}

}  // extern "C"
