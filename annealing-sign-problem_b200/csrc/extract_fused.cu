// Single-pass Ising-model extraction on sm_100a (the headline kernel).
//
// Replaces, fused into ONE kernel launch:
//   lattice_symmetries Operator.batched_apply      (called at annealing_sign_problem/common.py:96)
//   _clipped_search_sorted + membership mask       (common.py:116-128, :173)
//   bsearch loop of build_matrix                   (cbits/build_matrix.c:33-51)
//   _make_ising_model_compute_elements             (common.py:71-82)
//   csr_matrix(...) + sort_indices                 (common.py:193-195)
//
// Organisation.  ~37 candidates are generated and looked up per row and only a
// tenth of them are couplings, so the kernel is built around the cost of a MISS:
//   * one lane = one row, one warp = a tile of 32 CONSECUTIVE rows of the sorted basis; the warp
//     walks the operator's slots (operator.cuh: Slot, |delta|-sorted) together.  All lanes probe
//     the same slot at the same time, so their candidates s_r ^ flip = s_r +- |delta| are an
//     ascending run that lands in one short window of the index: the 32 probes of a step touch a
//     handful of 128-byte lines instead of 32;
//   * the index is ONE 8-byte word per bucket of the top key bits: {first position of the bucket,
//     16 order-preserving + 16 hashed presence bits}.  Both bit positions are GF(2)-linear in the
//     key, so a candidate costs two XORs on per-row / per-slot constants, one 8-byte load and two
//     funnel shifts; ~98 % of the misses end there;
//   * survivors are queued in shared memory and verified 32 at a time against the interleaved
//     {key, |psi|} records: the word's order bits give the position inside the bucket, so a hit is
//     normally ONE 16-byte load that also brings the amplitude's sector into L2;
//   * hits of a row arrive outward from the diagonal (down images descending, up images
//     ascending), so a hit's place in the row is its running count on its side; the tile's hit
//     list stays in shared memory (overflow spills to an L2-resident scratch list);
//   * a tile publishes its coupling count (never waits), fetches its CSR offset one tile late by a
//     two-level decoupled look-back (tile counts inside a group of 32 tiles, group totals across
//     groups) and writes indptr / indices / data in place.
// Values are (coef * |psi_j|) * |psi_i| and (d * |psi_i|) * |psi_i|: the association of the reference's
// live path (common.py:71-82), so pre-symmetrisation entries are bitwise the reference's.
#include "fused.cuh"

namespace asp {

constexpr int kFxWarps = 8;
constexpr int kFxThreads = kFxWarps * 32;
constexpr int kFxTileRows = 32;  // a tile is the 32 rows of one warp
static_assert(kFxTileRows == kFusedTileRows, "fused.cuh out of sync");
constexpr uint32_t kFxMaxSlots = 1023;  // hit tag: (slot * 2 + down) 11 bits | row lane 5 bits | count on its side 16 bits
constexpr int kFxUnroll = 4;            // slots probed together (index loads in flight per lane)
constexpr int kFxQueue = 64;            // survivor queue entries per warp (verified 32 at a time)
constexpr int kFxGroupTiles = 32;       // tiles per look-back group
constexpr int kFxMaxCtasPerSM = 8;
#ifndef ASP_FX_BACKOFF_NS
#define ASP_FX_BACKOFF_NS 32  // first sleep of a look-back that finds an unpublished count
#endif
#ifndef ASP_FX_LAG
#define ASP_FX_LAG 1
#endif
#ifndef ASP_FX_SCRATCH_MB
#define ASP_FX_SCRATCH_MB 512
#endif
constexpr int kFxLag = ASP_FX_LAG;  // tiles a warp counts ahead of the tile whose CSR offset it fetches
static_assert(kFxLag == 1, "the hit lists are double-buffered for a lag of one tile");
constexpr size_t kFxScratchBudget = static_cast<size_t>(ASP_FX_SCRATCH_MB) << 20;  // spill lists never exceed this

constexpr unsigned long long kPrefixValid = 1ull << 63;
constexpr int kGroupCountShift = 48;  // group word: tiles published << 48 | couplings

#ifdef ASP_FX_DEBUG  // bounds checks of every computed address (compute-sanitizer is not available on the GPU pool)
#define FX_CHECK(cond, code, value)                                                                                          \
  do {                                                                                                                       \
    if (!(cond)) {                                                                                                           \
      printf("FX_CHECK %d failed: value %llu (block %d thread %d)\n", code, static_cast<unsigned long long>(value), blockIdx.x, \
             threadIdx.x);                                                                                                   \
      __trap();                                                                                                              \
    }                                                                                                                        \
  } while (0)
#else
#define FX_CHECK(cond, code, value) \
  do {                              \
  } while (0)
#endif

struct __align__(16) Record {  // the sorted basis as the kernel reads it: key and |psi| share a sector
  uint64_t key;
  double amp;
};

// Shared-memory layout.  CTA-wide slot tables, then one block per warp.
struct FxLayout {
  uint32_t probe, side, flip, coef, groups, diag;  // byte offsets of the tables
  uint32_t tables;                                 // bytes of all tables
  uint32_t w_queue, w_key, w_cnt, w_pend, w_list;  // byte offsets inside a warp's block
  uint32_t list_entries;                           // entries of ONE of the two hit lists
  uint32_t per_warp;
};

struct FusedArgs {
  FxLayout layout;         // computed on the host: the kernel reads the offsets from constant memory
  const Record *rec;       // [n_total] ascending, unique keys + |amplitude|
  const uint2 *index;      // [2^bbits + 1] {first position of the bucket, presence bits}
  int bshift, oshift;      // bucket = key >> bshift; order bit = (key >> oshift) & 15
  uint64_t state_mask;     // keys with bits outside are never candidates
  uint32_t n_total;
  uint64_t row_begin, num_rows, num_tiles;
  const Slot *slots;
  int n_slots, n_slots_padded;
  const DiagBond *diag;
  int n_diag;
  const DiagGroup *groups; // closed form of the diagonal (n_groups < 0: use the bond loop)
  int n_groups, diag_scale;
  long long diag_c0;
  uint2 *scratch;          // [gridDim.x * kFxWarps][2][scratch_per_warp] spilled hits {position, tag}
  uint32_t scratch_per_warp;
  uint32_t *tile_count;            // [num_tiles] couplings of a tile | 1 << 31 once published (zeroed)
  unsigned long long *group_count; // [groups] tiles published << 48 | couplings (zeroed)
  unsigned long long *group_prefix;// [groups] couplings before the group | 1 << 63 once known (zeroed)
  unsigned int *ticket;            // zeroed
  uint64_t num_buckets, num_groups, scratch_lists;  // (bounds, debug checks only)
  uint64_t capacity;               // entries the caller's indices/data can hold
  int64_t *indptr;
  int32_t *indices;
  double *data;
  unsigned long long *nnz_out;          // running total after this launch (device)
  unsigned long long *nnz_mirror;       // optional copy of it in mapped host memory
  const unsigned long long *base_in;   // couplings emitted by earlier row chunks (NULL = 0)
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// GF(2)-LINEAR mix of a key: hash(s ^ flip) = hash(s) ^ hash(flip), so the kernel hashes every row once
// and every slot once instead of every candidate.
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
// The two presence-bit positions of a key inside its bucket's word, packed as ord | hsh << 8:
// ord = the four key bits below the bucket bits (order-preserving inside the bucket), hsh = four hashed bits.
// Both are linear in the key.  (The probing row adds 16 << 8: the hashed bit lives in the upper half-word.)
__host__ __device__ __forceinline__ uint32_t index_sub(uint64_t key, int oshift) {
  return (static_cast<uint32_t>(key >> oshift) & 15u) | ((filter_hash(key) & 15u) << 8);
}
__host__ __device__ __forceinline__ uint32_t index_bits(uint32_t sub) { return (1u << (sub & 15u)) | (1u << (16u + ((sub >> 8) & 15u))); }

// Shared memory through 32-bit addresses.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint4 lds_table_v4(uint32_t addr) {  // read-only tables: free to schedule
  uint4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

#include "exchange_kernels.cuh"  // index + X1 kernels (gather_index_*, gather_copy_tma_kernel, index_block_kernel)

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

__host__ __device__ inline FxLayout fx_layout(int n_slots_padded, int n_groups, int n_diag, uint32_t list_entries) {
  FxLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    const uint32_t at = off;
    off += (bytes + 15u) & ~15u;
    return at;
  };
  const uint32_t ns = static_cast<uint32_t>(n_slots_padded);
  L.probe = take(ns * 32u);  // {flip >> bshift, sub(flip), mask lo/hi, need_down lo/hi, need_down ^ need_up lo/hi}
  L.side = take(ns * 16u);   // {mask, need_down}
  L.flip = take(ns * 8u);
  L.coef = take(ns * 16u);   // [slot * 2 + down]
  L.groups = take(n_groups >= 0 ? static_cast<uint32_t>(n_groups) * 16u : 0u);
  L.diag = take(n_groups >= 0 ? 0u : static_cast<uint32_t>(n_diag) * static_cast<uint32_t>(sizeof(DiagBond)));
  L.tables = off;
  off = 0;
  L.w_queue = take(kFxQueue * 16u);  // survivors {start, bits, sub, tag} | write phase: row offsets [32], down counts [32], |psi| [32]
  L.w_key = take(32u * 8u);          // the tile's keys
  L.w_cnt = take(64u * 4u);          // hits so far per row: up [32], down [32]
  L.w_pend = take(36u * 4u);         // waiting tile: packed row counts [32], tile lo/hi, list length, spilled length
  L.w_list = take(2u * list_entries * 8u);  // two hit lists {position, tag}: the tile being counted and the waiting one
  L.list_entries = list_entries;
  L.per_warp = off;
  return L;
}

#ifndef ASP_FX_MIN_CTAS
#define ASP_FX_MIN_CTAS 4
#endif
template <bool kWide>  // keys wider than 32 bits
__global__ void __launch_bounds__(kFxThreads, ASP_FX_MIN_CTAS) extract_csr_kernel(const FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FxLayout &L = a.layout;
  uint4 *s_probe = reinterpret_cast<uint4 *>(smem_raw + L.probe);
  ulonglong2 *s_side = reinterpret_cast<ulonglong2 *>(smem_raw + L.side);
  uint64_t *s_flip = reinterpret_cast<uint64_t *>(smem_raw + L.flip);
  double *s_coef = reinterpret_cast<double *>(smem_raw + L.coef);
  DiagGroup *s_groups = reinterpret_cast<DiagGroup *>(smem_raw + L.groups);
  DiagBond *s_diag = reinterpret_cast<DiagBond *>(smem_raw + L.diag);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned char *const wbase = smem_raw + L.tables + L.per_warp * warp;
  uint4 *const w_queue = reinterpret_cast<uint4 *>(wbase + L.w_queue);
  uint32_t *const w_row_off = reinterpret_cast<uint32_t *>(wbase + L.w_queue);        // write phase (the queue is empty then)
  uint32_t *const w_down = reinterpret_cast<uint32_t *>(wbase + L.w_queue + 128);
  double *const w_abs_psi = reinterpret_cast<double *>(wbase + L.w_queue + 256);
  uint64_t *const w_key = reinterpret_cast<uint64_t *>(wbase + L.w_key);
  uint32_t *const w_cnt = reinterpret_cast<uint32_t *>(wbase + L.w_cnt);
  uint32_t *const w_pend = reinterpret_cast<uint32_t *>(wbase + L.w_pend);
  uint2 *const w_lists = reinterpret_cast<uint2 *>(wbase + L.w_list);
  const uint32_t list_cap = L.list_entries;

  for (int k = threadIdx.x; k < a.n_slots_padded; k += kFxThreads) {
    // padding slots never apply: (s & 0) ^ ~0 is neither 0 nor need_down ^ need_up = 0
    Slot sl{0ull, 0ull, ~0ull, ~0ull, 0.0, 0.0};
    if (k < a.n_slots) sl = a.slots[k];
    const uint64_t dd = sl.need_down ^ sl.need_up;
    s_probe[2 * k] = make_uint4(static_cast<uint32_t>(sl.flip >> a.bshift), index_sub(sl.flip, a.oshift), static_cast<uint32_t>(sl.mask),
                                static_cast<uint32_t>(sl.mask >> 32));
    s_probe[2 * k + 1] = make_uint4(static_cast<uint32_t>(sl.need_down), static_cast<uint32_t>(sl.need_down >> 32), static_cast<uint32_t>(dd),
                                    static_cast<uint32_t>(dd >> 32));
    s_side[k] = make_ulonglong2(sl.mask, sl.need_down);
    s_flip[k] = sl.flip;
    s_coef[2 * k] = sl.coef_up;
    s_coef[2 * k + 1] = sl.coef_down;
  }
  if (a.n_groups >= 0) {
    for (int k = threadIdx.x; k < a.n_groups; k += kFxThreads) s_groups[k] = a.groups[k];
  } else {
    for (int k = threadIdx.x; k < a.n_diag; k += kFxThreads) s_diag[k] = a.diag[k];
  }
  __syncthreads();  // tables loaded; from here on the warps run independently
  FX_CHECK((static_cast<uint64_t>(blockIdx.x) * kFxWarps + warp) * 2 + 1 < a.scratch_lists, 13, blockIdx.x);
  uint2 *const my_spill = a.scratch + (static_cast<size_t>(blockIdx.x) * kFxWarps + warp) * 2 * a.scratch_per_warp;  // one spill list per hit list
  const uint32_t probe_base = smem_addr(s_probe);
  const uint32_t queue_base = smem_addr(w_queue);
  const uint32_t lane_tag = lane << 11;

  // Lag: a warp counts its NEXT tile before it fetches the CSR offset of the tile it counted last.
  // The offset needs every earlier tile's count, and a warp that finished early would otherwise spin
  // for the slowest of its predecessors (the SM's warp arbiter is not fair, so some warps ARE much
  // slower); by the time the next tile is counted they are done.  Counts are published right after
  // counting and never wait for anything, so the scheme cannot deadlock.  The waiting tile keeps its
  // hit list in the other half of the warp's list memory and its row counts in w_pend.
  uint32_t taken = 0, waiting = 0;  // tiles this warp has counted / of those, not yet written (warp-uniform)
  for (;;) {
    // a tile = the 32 rows of one warp; tiles are handed out in order by an atomic ticket
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.ticket, 1u);
    const uint64_t tile = __shfl_sync(0xffffffffu, ticket, 0);
    const bool have_tile = tile < a.num_tiles;
    if (!have_tile && waiting == 0) break;
    uint2 *const my_list = w_lists + static_cast<size_t>(taken & 1u) * list_cap;
    uint2 *const my_spill_list = my_spill + static_cast<size_t>(taken & 1u) * a.scratch_per_warp;
    uint32_t packed_cnt = 0, list_count = 0, spilled = 0;
    if (have_tile) {
      const uint64_t r = tile * kFxTileRows + lane;
      const bool live = r < a.num_rows;
      const uint64_t row = a.row_begin + r;
      uint64_t s = 0ull;
      if (live) s = __ldg(&a.rec[row].key);
      const uint32_t s_lo = static_cast<uint32_t>(s), s_hi = static_cast<uint32_t>(s >> 32);
      const bool generates = live && (s & ~a.state_mask) == 0;  // keys wider than the word have no images in the basis
      // per-row halves of the probe: bucket index and the two presence-bit positions (hashed one in the upper half-word)
      const uint32_t s_idx = static_cast<uint32_t>(s >> a.bshift);
      const uint32_t s_sub = index_sub(s, a.oshift) | (16u << 8);
      w_key[lane] = s;
      w_cnt[lane] = 0;
      w_cnt[32 + lane] = 0;
      __syncwarp();

      uint32_t q_head = 0, q_count = 0;  // survivor queue (ring of kFxQueue entries)

      // Verify up to 32 queued survivors against the records, rank the hits, append them to the hit list.
      auto verify = [&](uint32_t batch) {
        const bool valid = lane < batch;
        const uint4 q = w_queue[(q_head + lane) & (kFxQueue - 1)];
        const uint32_t tag = valid ? q.w : 0u;  // entries beyond the batch hold stale bytes
        const uint32_t slot = tag & 0x7FFu, src = (tag >> 11) & 31u;
#ifdef ASP_FX_DEBUG
        if (valid && slot >= static_cast<uint32_t>(a.n_slots)) {
          printf("queue garbage: block %d warp %u lane %u taken %u q_head %u q_count %u batch %u entry %08x %08x %08x %08x\n", blockIdx.x, warp, lane, taken,
                 q_head, q_count, batch, q.x, q.y, q.z, q.w);
          __trap();
        }
#endif
        const ulonglong2 side = s_side[slot];
        const uint64_t key = w_key[src];
        const uint64_t c = key ^ s_flip[slot];
        const uint32_t down = (key & side.x) == side.y ? 1u : 0u;
        // the order bits below the candidate's own: that many keys of the bucket precede it (at least)
        uint32_t p = q.x + __popc(q.y & ((1u << (q.z & 15u)) - 1u));
        bool hit = false;
        FX_CHECK(!valid || (c >> a.bshift) < a.num_buckets, 2, c);
        if (valid && p < a.n_total) {
          uint64_t k = __ldg(&a.rec[p].key);
          if (k < c) {
            // keys of the bucket share an order bit (rare), or the bucket is crowded (clustered bases):
            // lower bound between the guess and the first position of the next bucket
            uint32_t lo = p + 1, hi = __ldg(&a.index[(c >> a.bshift) + 1].x);
            FX_CHECK(hi <= a.n_total, 3, hi);
            while (lo < hi) {
              const uint32_t mid = lo + ((hi - lo) >> 1);
              if (__ldg(&a.rec[mid].key) < c)
                lo = mid + 1;
              else
                hi = mid;
            }
            p = lo;
            k = p < a.n_total ? __ldg(&a.rec[p].key) : ~c;
          }
          hit = k == c;
        }
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (hits) {
          if (list_count + 32u > list_cap) {  // (warp-uniform) no room for a full batch: spill the list to the scratch list
            FX_CHECK(spilled + list_count <= a.scratch_per_warp, 4, spilled + list_count);
            for (uint32_t k = lane; k < list_count; k += 32) my_spill_list[spilled + k] = my_list[k];
            spilled += list_count;
            list_count = 0;
            __syncwarp();
          }
          // hits of one row on one side arrive outward from the diagonal: the count so far is the place.
          // Every lane takes part in the match and the barrier (a partial-mask barrier inside the divergent
          // branch can pair up with the full-mask barriers of the lanes outside it).
          const uint32_t who = hit ? (src | (down << 5)) : (64u + lane);
          const uint32_t peers = __match_any_sync(0xffffffffu, who);  // same row and side in this batch (ascending slot order by lane)
          const uint32_t before = __popc(peers & lt_mask);
          const uint32_t sofar = hit ? w_cnt[who] : 0u;
          __syncwarp();
          if (hit) {
            if (before == 0) w_cnt[who] = sofar + __popc(peers);
            FX_CHECK(list_count + __popc(hits & lt_mask) < list_cap, 5, list_count);
            FX_CHECK(sofar + before < 1024u, 6, sofar + before);
            my_list[list_count + __popc(hits & lt_mask)] = make_uint2(p, (slot * 2u + down) | (src << 11) | ((sofar + before) << 16));
          }
          list_count += __popc(hits);
        }
        __syncwarp();
        q_head += batch;
        q_count -= batch;
      };

      // The probe loop: kFxUnroll slots at a time, their index words in flight together.
      for (int g0 = 0; g0 < a.n_slots_padded; g0 += kFxUnroll) {
        uint2 word[kFxUnroll];
        uint32_t sub[kFxUnroll];
#pragma unroll
        for (int j = 0; j < kFxUnroll; ++j) {
          const uint32_t at = probe_base + static_cast<uint32_t>(g0 + j) * 32u;
          const uint4 p0 = lds_table_v4(at), p1 = lds_table_v4(at + 16u);
          // u = (s & mask) ^ need_down: 0 when the down move applies, need_down ^ need_up when the up move does
          const uint32_t u_lo = (s_lo & p0.z) ^ p1.x;
          const uint32_t u_hi = kWide ? (s_hi & p0.w) ^ p1.y : 0u;
          const bool applies = (((u_lo | u_hi) == 0u) || (((u_lo ^ p1.z) | (kWide ? (u_hi ^ p1.w) : 0u)) == 0u)) && generates;
          sub[j] = s_sub ^ p0.y;
          word[j] = make_uint2(0u, 0u);
          FX_CHECK(!applies || (s_idx ^ p0.x) < a.num_buckets, 7, s_idx ^ p0.x);
          if (applies) word[j] = __ldg(&a.index[s_idx ^ p0.x]);
        }
#pragma unroll
        for (int j = 0; j < kFxUnroll; ++j) {
          // both presence bits set -> the candidate survives into the queue (a row without a candidate holds word 0)
          const bool pass = (__funnelshift_r(word[j].y, 0u, sub[j]) & __funnelshift_r(word[j].y, 0u, sub[j] >> 8) & 1u) != 0u;
          const uint32_t passed = __ballot_sync(0xffffffffu, pass);
#ifdef ASP_FX_DEBUG
          {
            const uint32_t p_first = __shfl_sync(0xffffffffu, passed, 0), c_first = __shfl_sync(0xffffffffu, q_count, 0), h_first = __shfl_sync(0xffffffffu, q_head, 0);
            if (p_first != passed || c_first != q_count || h_first != q_head) {
              printf("not uniform: block %d warp %u lane %u taken %u slot %d passed %08x/%08x q_count %u/%u q_head %u/%u pass %d\n", blockIdx.x, warp, lane, taken,
                     g0 + j, passed, p_first, q_count, c_first, q_head, h_first, static_cast<int>(pass));
              __trap();
            }
          }
#endif
          if (passed) {
            if (pass)
              w_queue[(q_head + q_count + __popc(passed & lt_mask)) & (kFxQueue - 1)] =
                  make_uint4(word[j].x, word[j].y, sub[j], static_cast<uint32_t>(g0 + j) | lane_tag);
            q_count += __popc(passed);
            if (q_count >= 32u) {
              __syncwarp();
              verify(32u);
            }
          }
        }
      }
      __syncwarp();
      if (q_count) verify(q_count);
#ifdef ASP_FX_DEBUG
      for (uint32_t k = lane; k < list_count; k += 32) {  // the list as it stands right after counting
        const uint2 e = my_list[k];
        const uint32_t side = (e.y & 1u) << 5 | ((e.y >> 11) & 31u);
        FX_CHECK(e.x < a.n_total, 20, e.x);
        FX_CHECK((e.y >> 16) < w_cnt[side] || spilled != 0, 21, (static_cast<unsigned long long>(taken) << 32) | k);
      }
      __syncwarp();
#endif
      const uint32_t up_cnt = w_cnt[lane], down_cnt = w_cnt[32 + lane];
      packed_cnt = (up_cnt + down_cnt + (live ? 1u : 0u)) | (down_cnt << 16);  // couplings of the row (diagonal included) | those below the diagonal << 16
      // publish the tile's count -- never waits
      uint32_t total = packed_cnt & 0xFFFFu;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      FX_CHECK(tile / kFxGroupTiles < a.num_groups, 8, tile);
      if (lane == 0) {
        st_relaxed_u32(&a.tile_count[tile], total | 0x80000000u);
        atomicAdd(&a.group_count[tile / kFxGroupTiles], (1ull << kGroupCountShift) | total);
      }
    }  // have_tile

    if (waiting == kFxLag || (!have_tile && waiting != 0)) {
      // ======================= CSR offset of the waiting tile by two-level decoupled look-back
      const uint32_t oldest = taken - waiting;
      const uint32_t pend_packed = w_pend[lane], pend_count = w_pend[34], pend_spilled = w_pend[35];
      FX_CHECK(pend_count <= list_cap && pend_spilled <= a.scratch_per_warp, 12, pend_count);
      const uint64_t ptile = (static_cast<uint64_t>(w_pend[33]) << 32) | w_pend[32];
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
      const uint64_t group = ptile / kFxGroupTiles;
      const uint32_t in_group = static_cast<uint32_t>(ptile % kFxGroupTiles);
      // (1) the earlier tiles of this tile's group (they hold earlier tickets: counted or being counted)
      uint32_t earlier = 0;
      if (lane < in_group) {
        unsigned backoff = ASP_FX_BACKOFF_NS;
        unsigned long long t0 = 0;
        while (((earlier = ld_relaxed_u32(&a.tile_count[group * kFxGroupTiles + lane])) >> 31) == 0u) {
          __nanosleep(backoff);
          if (backoff < 32 * ASP_FX_BACKOFF_NS) {
            backoff <<= 1;
          } else {  // a count that stays away for seconds is a bug, not a slow tile: fail, never hang the device
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ull) __trap();
          }
        }
        earlier &= 0x7FFFFFFFu;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) earlier += __shfl_xor_sync(0xffffffffu, earlier, o);
      // (2) couplings before the group: its published prefix, or the nearest published prefix plus the complete
      // groups in between (32 groups per step; every group before this one is full)
      unsigned long long before_group = a.base_in ? *a.base_in : 0ull;
      if (group != 0) {
        // ONE lane reads the word and hands it to the others: the branch below must be taken by the whole warp,
        // and lanes that read for themselves can see the word before and after it is published
        unsigned long long known = 0ull;
        if (lane == 0) known = ld_relaxed_u64(&a.group_prefix[group]);
        known = __shfl_sync(0xffffffffu, known, 0);
        if (known & kPrefixValid) {
          before_group = known & ~kPrefixValid;
        } else {
          unsigned long long acc = 0;
          int64_t look = static_cast<int64_t>(group) - 1;  // window [look - 31, look]
          for (;;) {
            const int64_t g = look - lane;
            unsigned long long prefix = 0;
            bool has_prefix = g <= 0;  // group 0 starts at base_in
            if (g > 0) {
              prefix = ld_relaxed_u64(&a.group_prefix[g]);
              has_prefix = (prefix & kPrefixValid) != 0;
              prefix &= ~kPrefixValid;
            } else {
              prefix = before_group;
            }
            const uint32_t found = __ballot_sync(0xffffffffu, has_prefix);
            const uint32_t first = static_cast<uint32_t>(__ffs(static_cast<int>(found))) - 1u;  // nearest group with a prefix (lane 0 = nearest); found != 0 when look < 32
            const uint32_t upto = found ? first : 31u;
            unsigned long long count = 0;
            if (lane <= upto && g >= 0) {  // groups g .. look contribute their totals
              unsigned backoff = ASP_FX_BACKOFF_NS;
              unsigned long long t0 = 0;
              while (((count = ld_relaxed_u64(&a.group_count[g])) >> kGroupCountShift) != static_cast<unsigned long long>(kFxGroupTiles)) {
                __nanosleep(backoff);
                if (backoff < 32 * ASP_FX_BACKOFF_NS) {
                  backoff <<= 1;
                } else {
                  const unsigned long long now = global_timer_ns();
                  if (t0 == 0) t0 = now;
                  if (now - t0 > 4000000000ull) __trap();
                }
              }
              count &= (1ull << kGroupCountShift) - 1ull;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
            acc += count;
            if (found) {
              acc += __shfl_sync(0xffffffffu, prefix, first);
              break;
            }
            look -= 32;
          }
          before_group = acc;
          if (lane == 0) st_relaxed_u64(&a.group_prefix[group], acc | kPrefixValid);
        }
      }
      const unsigned long long exclusive = before_group + earlier;
      if (lane == 0 && ptile == a.num_tiles - 1) {
        a.indptr[a.num_rows] = static_cast<int64_t>(exclusive + tile_total);
        *a.nnz_out = exclusive + tile_total;
        if (a.nnz_mirror) {
          *a.nnz_mirror = exclusive + tile_total;
          __threadfence_system();
        }
      }
      const uint64_t tile_base = exclusive;

      // ======================= write the waiting tile's CSR rows ============================
      const uint2 *const list = w_lists + static_cast<size_t>(oldest & 1u) * list_cap;
      const uint2 *const spill_list = my_spill + static_cast<size_t>(oldest & 1u) * a.scratch_per_warp;
      const uint32_t my_off = incl - my_cnt;  // row start relative to the tile
      Record mine{0ull, 0.0};
      if (live) {
        const ulonglong2 raw = __ldg(reinterpret_cast<const ulonglong2 *>(&a.rec[row]));
        mine.key = raw.x;
        mine.amp = __longlong_as_double(static_cast<long long>(raw.y));
      }
      const double a_i = mine.amp;
      w_row_off[lane] = my_off;
      w_down[lane] = down_cnt;
      w_abs_psi[lane] = a_i;
      if (live) a.indptr[r] = static_cast<int64_t>(tile_base + my_off);
      __syncwarp();
      auto emit = [&](uint2 entry) {
        const uint32_t pos = entry.x, code = entry.y & 0x7FFu, src = (entry.y >> 11) & 31u, k = entry.y >> 16;
        const uint32_t dn = w_down[src];
        FX_CHECK(pos < a.n_total, 9, (static_cast<unsigned long long>(taken) << 40) | (static_cast<unsigned long long>(waiting) << 36) | (static_cast<unsigned long long>(have_tile) << 32) | pend_count);
        FX_CHECK(code < 2u * static_cast<uint32_t>(a.n_slots), 10, code);
        FX_CHECK((code & 1u) ? k < dn : true, 11, k);
        const uint64_t dest = tile_base + w_row_off[src] + ((code & 1u) ? dn - 1u - k : dn + 1u + k);
        if (dest < a.capacity) {
          a.indices[dest] = static_cast<int32_t>(pos);
          a.data[dest] = __dmul_rn(__dmul_rn(s_coef[code], __ldg(&a.rec[pos].amp)), w_abs_psi[src]);
        }
      };
      for (uint32_t k = lane; k < pend_spilled; k += 32) emit(spill_list[k]);
      for (uint32_t k = lane; k < pend_count; k += 32) emit(list[k]);
      if (live) {
        const double d = a.n_groups >= 0 ? diagonal_closed_form(mine.key, s_groups, a.n_groups, a.diag_c0, a.diag_scale)
                                         : diagonal_element(mine.key, s_diag, a.n_diag);
        const uint64_t dest = tile_base + my_off + down_cnt;
        if (dest < a.capacity) {
          a.indices[dest] = static_cast<int32_t>(row);
          a.data[dest] = __dmul_rn(__dmul_rn(d, a_i), a_i);
        }
      }
      __syncwarp();  // the row offsets are read by all lanes; the next tile's queue overwrites their bytes
      --waiting;
    }
    if (have_tile) {  // the tile just counted joins the queue
      w_pend[lane] = packed_cnt;
      if (lane == 0) {
        w_pend[32] = static_cast<uint32_t>(tile);
        w_pend[33] = static_cast<uint32_t>(tile >> 32);
        w_pend[34] = list_count;
        w_pend[35] = spilled;
      }
      __syncwarp();
      ++taken;
      ++waiting;
    }
  }
}

constexpr int kFxMaxChunks = kFusedMaxChunks;  // row chunks of one pipelined host call

struct FusedWorkspace {
  Record *rec;                       // [n_total] {key, |psi|}
  uint2 *index;                      // [num_buckets + 1] {first position, presence bits}
  uint32_t *tile_count;              // [tiles + kFxMaxChunks]: every row chunk has its own slice
  unsigned long long *group_count;   // [groups + kFxMaxChunks]
  unsigned long long *group_prefix;  // [groups + kFxMaxChunks]
  unsigned int *tickets;             // [kFxMaxChunks] one per chunk, 64 B apart
  unsigned long long *totals;        // [kFxMaxChunks] running totals
  uint2 *scratch;                    // spill lists
  uint64_t num_buckets;
  int bshift, oshift;
  uint32_t scratch_per_warp, scratch_ctas;
  size_t bytes, zero_offset, zero_bytes;
};

static int g_list_entries_override = 0;
// optional CUDA-event bracket around the extraction kernel alone (bench.py's roofline figure)
static bool g_time_kernel = false;
constexpr int kEvRing = 64;  // the last kEvRing launches keep their event pair
static cudaEvent_t g_ev_begin[kEvRing] = {}, g_ev_end[kEvRing] = {};
static uint64_t g_ev_launches = 0;
static int g_bucket_bits_delta = 0, g_diag_mode = 0;
static int g_gather_mode = 2;  // asp_gather_index: 2 = one TMA kernel (default), 1 = one kernel with plain loads, 0 = copy engines + per-block index kernels

static FusedWorkspace carve_fused(void *base, const asp_operator *op, uint64_t n_total, uint64_t num_rows) {
  FusedWorkspace w;
  int lg = 0;
  while ((1ull << lg) < n_total) ++lg;
  const int key_bits = static_cast<int>(op->number_spins);
  int bbits = lg - 2 + g_bucket_bits_delta;  // two to four keys per bucket: 8 bytes of index per ~3 keys
  bbits = std::max(2, std::min(bbits, 27));
  bbits = std::min(bbits, key_bits);
  w.bshift = key_bits - bbits;
  w.oshift = std::max(w.bshift - 4, 0);
  w.num_buckets = 1ull << bbits;
  const uint64_t tiles = (num_rows + kFxTileRows - 1) / kFxTileRows + kFxMaxChunks;
  const uint64_t groups = tiles / kFxGroupTiles + 1 + kFxMaxChunks;
  w.scratch_per_warp = std::max<uint32_t>(32u * static_cast<uint32_t>(op->moves.size()), 32u);
  const size_t per_cta = static_cast<size_t>(w.scratch_per_warp) * 2 * kFxWarps * sizeof(uint2);  // two spill lists per warp
  uint64_t ctas = std::min<uint64_t>(static_cast<uint64_t>(kNumSMs) * kFxMaxCtasPerSM, std::max<uint64_t>(tiles, 1));
  ctas = std::min<uint64_t>(ctas, std::max<uint64_t>(kFxScratchBudget / per_cta, kNumSMs));
  w.scratch_ctas = static_cast<uint32_t>(ctas);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void *p = base ? static_cast<char *>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.rec = static_cast<Record *>(take((n_total + 1) * sizeof(Record)));
  w.scratch = static_cast<uint2 *>(take(per_cta * ctas));
  w.zero_offset = off;
  w.index = static_cast<uint2 *>(take((w.num_buckets + 1) * sizeof(uint2)));
  w.tile_count = static_cast<uint32_t *>(take(tiles * sizeof(uint32_t)));
  w.group_count = static_cast<unsigned long long *>(take(groups * sizeof(unsigned long long)));
  w.group_prefix = static_cast<unsigned long long *>(take(groups * sizeof(unsigned long long)));
  w.tickets = static_cast<unsigned int *>(take(kFxMaxChunks * 64));
  w.totals = static_cast<unsigned long long *>(take(kFxMaxChunks * sizeof(unsigned long long)));
  w.zero_bytes = off - w.zero_offset;
  w.bytes = off;
  return w;
}

size_t fused_workspace_bytes(const asp_operator *op, uint64_t n_total, uint64_t num_rows) {
  return carve_fused(nullptr, op, n_total, num_rows).bytes;
}

int fused_check_operator(const asp_operator *op) {
  ASP_REQUIRE(op != nullptr, "operator is NULL");
  ASP_REQUIRE(op->d_moves != nullptr && op->d_slots != nullptr, "operator has no device mirror (created without a CUDA device)");
  if (!op->sorted_emitter() || op->slots.size() > kFxMaxSlots) {
    set_error("fused extraction needs an unsymmetrised operator with distinct moves (at most %u move pairs); use the apply + build_matrix + canonicalise path", kFxMaxSlots);
    return ASP_ERR_UNSUPPORTED;
  }
  return ASP_OK;
}

// Zero the look-back state and index the sorted basis (once per call, before the chunks).
int fused_prepare(const asp_operator *op, uint64_t n_total, const uint64_t *d_spins, const double *d_psi, uint64_t num_rows, void *d_workspace,
                  size_t workspace_bytes, cudaStream_t s) {
  FusedWorkspace w = carve_fused(d_workspace, op, n_total, num_rows);
  if (d_workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return ASP_ERR_WORKSPACE;
  }
  ASP_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(d_workspace) + w.zero_offset, 0, w.zero_bytes, s));
  // two keys per thread: neighbours that share a bucket share one atomic (the pass is bound by L2 atomics)
  index_block_kernel<<<static_cast<unsigned>((n_total + 511) / 512), 256, 0, s>>>(d_spins, d_psi, static_cast<uint32_t>(n_total), 0u,
                                                                                 static_cast<uint32_t>(n_total), op->state_mask, w.bshift,
                                                                                 w.oshift, w.num_buckets, w.index, w.rec);
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
  a.bshift = w.bshift;
  a.oshift = w.oshift;
  a.index = w.index;
  a.rec = w.rec;
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
struct CopyLane {
  cudaStream_t copy = nullptr, copy2 = nullptr;  // keys on one copy engine, amplitudes on another
  cudaEvent_t start = nullptr, flag = nullptr, done[2] = {}, block[kGxMaxRanks] = {}, block2[kGxMaxRanks] = {};
  int device = -1;
};
static CopyLane g_lane;

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
    ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&g_lane.copy, cudaStreamNonBlocking));
    ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&g_lane.copy2, cudaStreamNonBlocking));
    ASP_CUDA_CHECK(cudaEventCreateWithFlags(&g_lane.start, cudaEventDisableTiming));
    ASP_CUDA_CHECK(cudaEventCreateWithFlags(&g_lane.flag, cudaEventDisableTiming));
    for (auto &e : g_lane.block) ASP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : g_lane.block2) ASP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : g_lane.done) ASP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_lane.device = dev;
  }
  if (index) ASP_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(d_workspace) + w.zero_offset, 0, w.zero_bytes, s));
  ASP_CUDA_CHECK(cudaEventRecord(g_lane.start, s));  // the copies may overwrite the full copy only after the caller's earlier work
  ASP_CUDA_CHECK(cudaStreamWaitEvent(g_lane.copy, g_lane.start, 0));
  ASP_CUDA_CHECK(cudaStreamWaitEvent(g_lane.copy2, g_lane.start, 0));
  SeamArgs seams{};
  for (uint32_t k = 0; k < world; ++k) {
    const uint32_t q = (rank + k) % world;  // own block first (no flag to wait for), then ring order
    ASP_REQUIRE(shard_begin[q + 1] >= shard_begin[q], "shard_begin must be non-decreasing");
    const uint64_t b0 = shard_begin[q], len = shard_begin[q + 1] - b0;
    seams.first[q] = len > 0 ? static_cast<uint32_t>(b0) : static_cast<uint32_t>(n_total);
    if (len == 0) continue;
    ASP_REQUIRE(d_shard_spins[q] && d_shard_psi[q], "NULL shard pointer");
    if (d_ready != nullptr && q != rank) {
      wait_one_flag_kernel<<<1, 1, 0, g_lane.copy>>>(reinterpret_cast<const unsigned long long *>(d_ready) + q, epoch);
      ASP_LAUNCH_CHECK();
      ASP_CUDA_CHECK(cudaEventRecord(g_lane.flag, g_lane.copy));  // the second engine starts on the same flag
      ASP_CUDA_CHECK(cudaStreamWaitEvent(g_lane.copy2, g_lane.flag, 0));
    }
    ASP_CUDA_CHECK(cudaMemcpyAsync(d_spins + b0, d_shard_spins[q], len * sizeof(uint64_t), cudaMemcpyDeviceToDevice, g_lane.copy));
    ASP_CUDA_CHECK(cudaMemcpyAsync(d_psi + b0, d_shard_psi[q], len * sizeof(double), cudaMemcpyDeviceToDevice, g_lane.copy2));
    if (!index) continue;
    ASP_CUDA_CHECK(cudaEventRecord(g_lane.block[k], g_lane.copy));
    ASP_CUDA_CHECK(cudaEventRecord(g_lane.block2[k], g_lane.copy2));
    ASP_CUDA_CHECK(cudaStreamWaitEvent(s, g_lane.block[k], 0));  // keys and amplitudes of this block have landed: index them
    ASP_CUDA_CHECK(cudaStreamWaitEvent(s, g_lane.block2[k], 0));
    index_block_kernel<<<static_cast<unsigned>((len + 511) / 512), 256, 0, s>>>(d_spins, d_psi, static_cast<uint32_t>(n_total), static_cast<uint32_t>(b0),
                                                                               static_cast<uint32_t>(b0 + len), op->state_mask, w.bshift, w.oshift,
                                                                               w.num_buckets, w.index, w.rec);
    ASP_LAUNCH_CHECK();
  }
  // everything has landed before the caller's stream goes on
  ASP_CUDA_CHECK(cudaEventRecord(g_lane.done[0], g_lane.copy));
  ASP_CUDA_CHECK(cudaEventRecord(g_lane.done[1], g_lane.copy2));
  ASP_CUDA_CHECK(cudaStreamWaitEvent(s, g_lane.done[0], 0));
  ASP_CUDA_CHECK(cudaStreamWaitEvent(s, g_lane.done[1], 0));
  if (index) {
    index_seam_kernel<<<1, 32, 0, s>>>(d_spins, static_cast<uint32_t>(n_total), seams, static_cast<int>(world), op->state_mask, w.bshift,
                                       w.num_buckets, w.index);
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
  (void)d_spins;
  (void)d_psi;  // the kernel reads the {key, |psi|} records the index pass made of them
  a.rec = w.rec;
  a.index = w.index;
  a.bshift = w.bshift;
  a.oshift = w.oshift;
  a.state_mask = op->state_mask;
  a.n_total = static_cast<uint32_t>(n_total);
  a.row_begin = row_begin + chunk_begin;
  a.num_rows = chunk_rows;
  a.num_tiles = (chunk_rows + kFxTileRows - 1) / kFxTileRows;
  a.slots = op->d_slots;
  a.n_slots = static_cast<int>(op->slots.size());
  a.n_slots_padded = (a.n_slots + kFxUnroll - 1) / kFxUnroll * kFxUnroll;
  a.diag = op->d_diag;
  a.n_diag = static_cast<int>(op->diag.size());
  const bool closed_form = op->diag_scale >= 0 && g_diag_mode != 1;
  a.groups = op->d_diag_groups;
  a.n_groups = closed_form ? static_cast<int>(op->diag_groups.size()) : -1;
  a.diag_scale = op->diag_scale;
  a.diag_c0 = op->diag_c0;
  a.scratch = w.scratch;
  a.scratch_per_warp = w.scratch_per_warp;
  a.scratch_lists = static_cast<uint64_t>(w.scratch_ctas) * kFxWarps * 2;
  a.num_buckets = w.num_buckets;
  a.num_groups = (chunk_rows + kFxTileRows - 1) / kFxTileRows / kFxGroupTiles + 1;
  const uint64_t first_tile = chunk_begin / kFxTileRows;
  a.tile_count = w.tile_count + first_tile + chunk;
  a.group_count = w.group_count + first_tile / kFxGroupTiles + chunk;
  a.group_prefix = w.group_prefix + first_tile / kFxGroupTiles + chunk;
  a.ticket = w.tickets + 16 * chunk;
  a.capacity = capacity;
  a.indptr = d_indptr + chunk_begin;
  a.indices = d_indices;
  a.data = d_data;
  a.nnz_out = w.totals + chunk;
  a.nnz_mirror = nnz_mirror;
  a.base_in = chunk == 0 ? nullptr : w.totals + (chunk - 1);
  const bool wide = op->number_spins > 32;
  auto kernel = wide ? extract_csr_kernel<true> : extract_csr_kernel<false>;
  // Hit lists: as long as the shared memory of an SM allows with kFxTargetCtas resident CTAs (a tile of the
  // headline workload holds ~120 hits; a longer list only matters for dense bases, which spill).
  int device = 0, smem_sm = 0, smem_cta = 0;
  ASP_CUDA_CHECK(cudaGetDevice(&device));
  ASP_CUDA_CHECK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
  ASP_CUDA_CHECK(cudaDeviceGetAttribute(&smem_cta, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  const FxLayout fixed = fx_layout(a.n_slots_padded, a.n_groups, a.n_diag, 0);
  const size_t fixed_bytes = fixed.tables + static_cast<size_t>(fixed.per_warp) * kFxWarps;
  uint32_t list_entries = 0;
  for (int ctas = ASP_FX_MIN_CTAS; ctas >= 1 && list_entries < 96; --ctas) {  // prefer more resident CTAs; never a list below 96 entries if fewer CTAs allow one
    const size_t budget = std::min<size_t>(static_cast<size_t>(smem_cta), (static_cast<size_t>(smem_sm) - 32 * 1024) / ctas - 1024);  // 32 KB stay L1
    if (budget <= fixed_bytes) continue;
    list_entries = static_cast<uint32_t>(std::min<size_t>((budget - fixed_bytes) / (kFxWarps * 2 * sizeof(uint2)), 512)) / 32 * 32;
  }
  if (g_list_entries_override > 0) list_entries = std::max(32, g_list_entries_override / 32 * 32);
  ASP_REQUIRE(list_entries >= 32, "operator too large for the fused kernel's shared-memory tables");
  const FxLayout layout = fx_layout(a.n_slots_padded, a.n_groups, a.n_diag, list_entries);
  a.layout = layout;
  const size_t smem = layout.tables + static_cast<size_t>(layout.per_warp) * kFxWarps;
  ASP_REQUIRE(smem <= static_cast<size_t>(smem_cta), "operator too large for the fused kernel's shared-memory tables");
  ASP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  ASP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kFxThreads, smem));
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
  kernel<<<grid, kFxThreads, smem, s>>>(a);
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

void asp_debug_set_hit_list_capacity(int entries_per_warp) { g_list_entries_override = entries_per_warp; }

void asp_debug_set_extract_tuning(int bucket_bits_delta, int reserved, int diag_mode) {
  (void)reserved;
  g_bucket_bits_delta = bucket_bits_delta;
  g_diag_mode = diag_mode;
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
  } else {
    rc = fused_prepare(op, n_total, d_spins, d_psi, num_rows, d_workspace, workspace_bytes, s);
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
    return fused_prepare_gather(op, world, rank, shard_begin, d_shard_spins, d_shard_psi, d_ready, epoch, d_spins, d_psi, num_rows,
                                d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream), g_gather_mode == 2);
  return fused_prepare_gather_ce(op, world, rank, shard_begin, d_shard_spins, d_shard_psi, d_ready, epoch, d_spins, d_psi, num_rows,
                                 d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
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

int asp_extract_csr_indexed(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, double const *d_psi,
                            uint64_t row_begin, uint64_t num_rows, void *d_workspace, size_t workspace_bytes,
                            uint64_t capacity, int64_t *d_indptr, int32_t *d_indices, double *d_data, uint64_t *h_nnz,
                            void *stream) {
  return extract_csr_impl(op, n_total, d_spins, d_psi, row_begin, num_rows, d_workspace, workspace_bytes, capacity, d_indptr,
                          d_indices, d_data, h_nnz, stream, /*indexed=*/true);
}

}  // extern "C"
