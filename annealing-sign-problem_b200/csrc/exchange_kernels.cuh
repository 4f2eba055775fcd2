// Kernels of the exchange step X1 and of the index over the sorted basis.  Included by
// extract_fused.cu (after filter_hash / filter_bits and the shared-memory helpers, inside namespace asp);
// the host side (fused_prepare*, asp_gather_index, asp_gather_blocks) lives there too.
//
//   index_block_kernel        first-position table + Bloom filter of a block of the basis (2 keys per thread)
//   gather_index_kernel       X1 + index in one kernel, plain 16-byte loads over NVLink peer memory
//   gather_index_tma_kernel   the same on the TMA: cp.async.bulk into shared memory, mbarrier pipeline (default)
//   gather_copy_tma_kernel    X1 alone, bulk copies in both directions, one thread per CTA
//   index_seam_kernel, wait_one_flag_kernel   helpers of the copy-engine variant
#pragma once

// ---- X1 fused with the index build: pull every rank's row block over NVLink peer memory ----------
// One kernel replaces ncclAllGather(keys) + ncclAllGather(amplitudes) + the index pass: the blocks
// of the sorted basis live in buffers the other processes of the node have mapped (CUDA IPC);
// a thread pulls two consecutive keys and amplitudes of one block with 16-byte loads (several in
// flight), writes the rank's private full copy and indexes the keys (first-position table + Bloom
// filter) while the next loads travel.  Blocks are visited in ring order rank+1, rank+2, ... so that
// at any time the ranks pull from different peers.  A CTA waits (system-scope acquire) for the
// "ready" flag of the blocks it touches; 10 s without it is a dead peer -> trap (never hang the box).
constexpr int kGxMaxRanks = 16;
constexpr int kGxThreads = 256;
constexpr int kGxUnroll = 4;  // key pairs per thread

struct GatherArgs {
  const uint64_t *shard_spins[kGxMaxRanks];
  const double *shard_psi[kGxMaxRanks];
  uint64_t begin[kGxMaxRanks + 1];       // global index of the first key of every block
  uint64_t unit_begin[kGxMaxRanks + 1];  // first key pair of the k-th VISITED block
  uint64_t chunk_begin[kGxMaxRanks + 1]; // first 1024-key chunk of the k-th visited block (TMA variant)
  int stages;                            // bulk copies in flight per CTA (TMA variant)
  int order[kGxMaxRanks];                // k-th visited block
  int world;
  const unsigned long long *ready;  // local flags [world] (NULL: no waiting)
  unsigned long long epoch;
  uint64_t *spins;  // [n] private full copy (out)
  double *psi;
  uint32_t n;
  uint64_t state_mask, num_buckets;
  int tshift, fshift;
  uint32_t *starts;
  uint2 *filter;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_flag_or_trap(const unsigned long long *flag, unsigned long long value) {
  if (ld_acquire_sys(flag) >= value) return;
  const unsigned long long t0 = global_timer_ns();
  unsigned ns = 64;
  while (ld_acquire_sys(flag) < value) {
    __nanosleep(ns);
    if (ns < 2048) ns <<= 1;
    if (global_timer_ns() - t0 > 10000000000ull) __trap();
  }
}
// Coherent (not .nc) loads: the data was released by another GPU in this very epoch.
__device__ __forceinline__ ulonglong2 ld_peer_v2(const void *p) {
  ulonglong2 v;
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_peer_u64(const void *p) {
  unsigned long long v;
  asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// First-position table entries owed by key i (predecessor pk): buckets (bucket(pk), bucket(key)].
__device__ __forceinline__ void index_table(const GatherArgs &a, uint64_t i, uint64_t key, bool has_prev, uint64_t pk) {
  const uint64_t last = a.num_buckets;
  const uint64_t b = (key & ~a.state_mask) ? last : key >> a.tshift;
  const uint64_t prev = has_prev ? ((pk & ~a.state_mask) ? last : pk >> a.tshift) + 1 : 0;
  for (uint64_t k = prev; k <= b && k <= last; ++k) a.starts[k] = static_cast<uint32_t>(i);
  if (i == a.n - 1)
    for (uint64_t k = b + 1; k <= last; ++k) a.starts[k] = a.n;
}

// Private copy + index of one or two consecutive keys (global positions g, g + 1).
__device__ __forceinline__ void emit_pair(const GatherArgs &a, uint64_t g, uint32_t cnt, ulonglong2 keys, ulonglong2 amps, bool has_prev,
                                          uint64_t pk) {
  a.spins[g] = keys.x;
  reinterpret_cast<unsigned long long *>(a.psi)[g] = amps.x;
  index_table(a, g, keys.x, has_prev, pk);
  const bool in0 = (keys.x & ~a.state_mask) == 0;
  const uint64_t w0 = keys.x >> a.fshift;
  unsigned long long bits0 = filter_bits(filter_hash(keys.x));
  if (cnt == 2) {
    a.spins[g + 1] = keys.y;
    reinterpret_cast<unsigned long long *>(a.psi)[g + 1] = amps.y;
    index_table(a, g + 1, keys.y, true, keys.x);
    if ((keys.y & ~a.state_mask) == 0) {
      const uint64_t w1 = keys.y >> a.fshift;
      const unsigned long long bits1 = filter_bits(filter_hash(keys.y));
      if (in0 && w1 == w0)
        bits0 |= bits1;  // sorted keys: neighbours often share a filter word -> one atomic for both
      else
        atomicOr(reinterpret_cast<unsigned long long *>(a.filter + w1), bits1);
    }
  }
  if (in0) atomicOr(reinterpret_cast<unsigned long long *>(a.filter + w0), bits0);
}

__global__ void __launch_bounds__(kGxThreads) gather_index_kernel(const GatherArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t units = a.unit_begin[a.world];
  const uint64_t cta_first = static_cast<uint64_t>(blockIdx.x) * (kGxThreads * kGxUnroll);
  if (cta_first >= units) return;
  if (a.ready != nullptr) {
    if (threadIdx.x == 0) {
      const uint64_t cta_last = min(cta_first + kGxThreads * kGxUnroll, units) - 1;
      for (int k = 0; k < a.world; ++k)
        if (a.unit_begin[k] <= cta_last && a.unit_begin[k + 1] > cta_first) wait_flag_or_trap(a.ready + a.order[k], a.epoch);
    }
    __syncthreads();
  }
  ulonglong2 keys[kGxUnroll], amps[kGxUnroll];
  uint64_t g[kGxUnroll];   // global index of the pair's first key (n: none)
  uint32_t cnt[kGxUnroll]; // keys of the pair (0, 1 or 2)
  int blk[kGxUnroll];
  uint64_t loc[kGxUnroll];
#pragma unroll
  for (int j = 0; j < kGxUnroll; ++j) {
    const uint64_t u = cta_first + static_cast<uint64_t>(j) * kGxThreads + threadIdx.x;
    cnt[j] = 0;
    g[j] = a.n;
    blk[j] = 0;
    loc[j] = 0;
    keys[j] = make_ulonglong2(0, 0);
    amps[j] = make_ulonglong2(0, 0);
    if (u < units) {
      int k = 0;
      while (u >= a.unit_begin[k + 1]) ++k;
      const int q = a.order[k];
      const uint64_t l = 2 * (u - a.unit_begin[k]), len = a.begin[q + 1] - a.begin[q];
      blk[j] = q;
      loc[j] = l;
      g[j] = a.begin[q] + l;
      if (l + 1 < len) {
        cnt[j] = 2;
        keys[j] = ld_peer_v2(a.shard_spins[q] + l);
        amps[j] = ld_peer_v2(a.shard_psi[q] + l);
      } else {
        cnt[j] = 1;
        keys[j].x = ld_peer_u64(a.shard_spins[q] + l);
        amps[j].x = ld_peer_u64(a.shard_psi[q] + l);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kGxUnroll; ++j) {
    // predecessor of the pair's first key: the neighbouring lane holds it unless this pair opens a
    // block (then it is the last key of the block before) or the lane is 0
    const uint64_t up = __shfl_up_sync(0xffffffffu, keys[j].y, 1);
    if (cnt[j] == 0) continue;
    uint64_t pk = up;
    bool has_prev = true;
    if (loc[j] == 0) {
      has_prev = g[j] > 0;
      if (has_prev) {
        int p = blk[j] - 1;
        while (a.begin[p + 1] == a.begin[p]) --p;  // skip empty blocks; g > 0 => one is not
        if (a.ready != nullptr) wait_flag_or_trap(a.ready + p, a.epoch);  // a block this CTA may not have waited for
        pk = ld_peer_u64(a.shard_spins[p] + (a.begin[p + 1] - a.begin[p] - 1));
      }
    } else if (lane == 0) {
      pk = ld_peer_u64(a.shard_spins[blk[j]] + loc[j] - 1);
    }
    emit_pair(a, g[j], cnt[j], keys[j], amps[j], has_prev, pk);
  }
}

// ---- X1, TMA variant of the fused kernel: a persistent CTA keeps kTxStages bulk copies
// (cp.async.bulk global -> shared, completion on an mbarrier) of 1024 keys + 1024 amplitudes in flight --
// 128 KB per SM without a single register -- while its threads index the chunk that has landed and write
// the private copy.  One elected thread waits for the block's ready flag and issues the copies.
constexpr int kTxChunk = 1024;
constexpr int kTxStages = 4;
constexpr int kTxThreads = 256;
constexpr int kTxCtasPerSM = 2;
constexpr size_t kTxSmem = static_cast<size_t>(kTxStages) * kTxChunk * 16 + 64;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(kTxThreads, 6) gather_index_tma_kernel(const GatherArgs a) {
  extern __shared__ __align__(128) unsigned char tx_smem[];
  const uint32_t smem0 = smem_addr(tx_smem);
  const int stages = a.stages;
  const uint32_t bars = smem0 + static_cast<uint32_t>(stages) * kTxChunk * 16;  // mbarriers behind the stages
  const uint64_t total = a.chunk_begin[a.world];
  if (threadIdx.x == 0) {
    for (int st = 0; st < stages; ++st) mbar_init(bars + 8 * st, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t waited = 0;  // blocks whose ready flag this CTA's issuing thread has seen
  auto locate = [&](uint64_t c, int &q, uint64_t &l0, uint32_t &cnt) {
    int k = 0;
    while (c >= a.chunk_begin[k + 1]) ++k;
    q = a.order[k];
    l0 = (c - a.chunk_begin[k]) * kTxChunk;
    const uint64_t len = a.begin[q + 1] - a.begin[q];
    cnt = static_cast<uint32_t>(min(static_cast<uint64_t>(kTxChunk), len - l0));
  };
  auto issue = [&](uint64_t it) {  // thread 0
    const uint64_t c = blockIdx.x + it * gridDim.x;
    if (c >= total) return;
    int q;
    uint64_t l0;
    uint32_t cnt;
    locate(c, q, l0, cnt);
    if (a.ready != nullptr && !((waited >> q) & 1u)) {
      wait_flag_or_trap(a.ready + q, a.epoch);
      asm volatile("fence.proxy.async;" ::: "memory");  // the async proxy reads what the acquire made visible
      waited |= 1u << q;
    }
    const uint32_t st = static_cast<uint32_t>(it % stages);
    const uint32_t even = cnt & ~1u;  // whole 16-byte units; an odd last key is read with a plain load
    mbar_arrive_expect_tx(bars + 8 * st, even * 16u);
    if (even) {
      bulk_g2s(smem0 + st * (kTxChunk * 16), a.shard_spins[q] + l0, even * 8u, bars + 8 * st);
      bulk_g2s(smem0 + st * (kTxChunk * 16) + kTxChunk * 8, a.shard_psi[q] + l0, even * 8u, bars + 8 * st);
    }
  };
  if (threadIdx.x == 0)
    for (int it = 0; it < stages; ++it) issue(it);
  for (uint64_t it = 0;; ++it) {
    const uint64_t c = blockIdx.x + it * gridDim.x;
    if (c >= total) break;
    const uint32_t st = static_cast<uint32_t>(it % stages);
    int q;
    uint64_t l0;
    uint32_t cnt;
    locate(c, q, l0, cnt);
    mbar_wait(bars + 8 * st, static_cast<uint32_t>(it / stages) & 1u);
    const unsigned long long *s_keys = reinterpret_cast<const unsigned long long *>(tx_smem + st * (kTxChunk * 16));
    const unsigned long long *s_amps = s_keys + kTxChunk;
#pragma unroll
    for (int h = 0; h < kTxChunk / 2 / kTxThreads; ++h) {
      const uint32_t e = 2u * (h * kTxThreads + threadIdx.x);  // first entry of this thread's pair inside the chunk
      if (e >= cnt) continue;
      const bool two = e + 1 < cnt;
      ulonglong2 keys, amps;
      if (two) {
        keys = *reinterpret_cast<const ulonglong2 *>(s_keys + e);
        amps = *reinterpret_cast<const ulonglong2 *>(s_amps + e);
      } else {  // odd tail of the block: not part of the bulk copy
        keys = make_ulonglong2(ld_peer_u64(a.shard_spins[q] + l0 + e), 0ull);
        amps = make_ulonglong2(ld_peer_u64(a.shard_psi[q] + l0 + e), 0ull);
      }
      const uint64_t g = a.begin[q] + l0 + e;
      bool has_prev = true;
      uint64_t pk;
      if (e > 0) {
        pk = s_keys[e - 1];
      } else if (l0 > 0) {
        pk = ld_peer_u64(a.shard_spins[q] + l0 - 1);
      } else {
        has_prev = g > 0;
        pk = 0;
        if (has_prev) {
          int p = q - 1;
          while (a.begin[p + 1] == a.begin[p]) --p;
          if (a.ready != nullptr) wait_flag_or_trap(a.ready + p, a.epoch);
          pk = ld_peer_u64(a.shard_spins[p] + (a.begin[p + 1] - a.begin[p] - 1));
        }
      }
      emit_pair(a, g, two ? 2u : 1u, keys, amps, has_prev, pk);
    }
    __syncthreads();  // every thread is done with this stage: refill it
    if (threadIdx.x == 0) issue(it + stages);
  }
}

// ---- X1, copy only, on the TMA: ONE thread per CTA moves chunks peer global -> shared -> private copy with bulk
// copies in both directions (no register, no LSU instruction per byte).  32-thread CTAs with 64 KB of shared memory
// fit beside a resident extraction kernel, so this is what gathers the NEXT basis while the current one is extracted.
constexpr int kCxStages = 3;  // 48 KB: fits beside six resident extraction CTAs
constexpr size_t kCxSmem = static_cast<size_t>(kCxStages) * kTxChunk * 16 + 64;

__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(32) gather_copy_tma_kernel(const GatherArgs a) {
  extern __shared__ __align__(128) unsigned char cx_smem[];
  const bool leader = threadIdx.x == 0;
  const uint32_t smem0 = smem_addr(cx_smem);
  const uint32_t bars = smem0 + kCxStages * kTxChunk * 16;
  const uint64_t total = a.chunk_begin[a.world];
  if (leader) {
    for (int st = 0; st < kCxStages; ++st) mbar_init(bars + 8 * st, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t waited = 0;
  auto locate = [&](uint64_t c, int &q, uint64_t &l0, uint32_t &cnt) {
    int k = 0;
    while (c >= a.chunk_begin[k + 1]) ++k;
    q = a.order[k];
    l0 = (c - a.chunk_begin[k]) * kTxChunk;
    const uint64_t len = a.begin[q + 1] - a.begin[q];
    cnt = static_cast<uint32_t>(min(static_cast<uint64_t>(kTxChunk), len - l0));
  };
  auto load = [&](uint64_t it) {  // leader only
    const uint64_t c = blockIdx.x + it * gridDim.x;
    if (c >= total) return;
    int q;
    uint64_t l0;
    uint32_t cnt;
    locate(c, q, l0, cnt);
    if (a.ready != nullptr && !((waited >> q) & 1u)) {
      wait_flag_or_trap(a.ready + q, a.epoch);
      asm volatile("fence.proxy.async;" ::: "memory");
      waited |= 1u << q;
    }
    const uint32_t st = static_cast<uint32_t>(it % kCxStages), even = cnt & ~1u;
    mbar_arrive_expect_tx(bars + 8 * st, even * 16u);
    if (even) {
      bulk_g2s(smem0 + st * (kTxChunk * 16), a.shard_spins[q] + l0, even * 8u, bars + 8 * st);
      bulk_g2s(smem0 + st * (kTxChunk * 16) + kTxChunk * 8, a.shard_psi[q] + l0, even * 8u, bars + 8 * st);
    }
  };
  if (leader)
    for (int it = 0; it < kCxStages; ++it) load(it);
  for (uint64_t it = 0;; ++it) {
    const uint64_t c = blockIdx.x + it * gridDim.x;
    if (c >= total) break;
    const uint32_t st = static_cast<uint32_t>(it % kCxStages);
    int q;
    uint64_t l0;
    uint32_t cnt;
    locate(c, q, l0, cnt);
    mbar_wait(bars + 8 * st, static_cast<uint32_t>(it / kCxStages) & 1u);
    const uint64_t g = a.begin[q] + l0;
    const uint32_t even = cnt & ~1u;
    if ((g & 1ull) == 0) {  // bulk stores need 16-byte aligned destinations
      if (leader && even) {
        bulk_s2g(a.spins + g, smem0 + st * (kTxChunk * 16), even * 8u);
        bulk_s2g(a.psi + g, smem0 + st * (kTxChunk * 16) + kTxChunk * 8, even * 8u);
      }
    } else {  // a block that starts at an odd global position: the warp stores the chunk with plain 8-byte stores
      const unsigned long long *sk = reinterpret_cast<const unsigned long long *>(cx_smem + st * (kTxChunk * 16));
      for (uint32_t e = threadIdx.x; e < even; e += 32) {
        a.spins[g + e] = sk[e];
        reinterpret_cast<unsigned long long *>(a.psi)[g + e] = sk[kTxChunk + e];
      }
    }
    if (leader) {
      if (cnt & 1u) {  // odd tail of the block: not part of the bulk copy
        a.spins[g + even] = ld_peer_u64(a.shard_spins[q] + l0 + even);
        reinterpret_cast<unsigned long long *>(a.psi)[g + even] = ld_peer_u64(a.shard_psi[q] + l0 + even);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stores have read the stage
    }
    __syncwarp();  // ... and so have the lanes of the plain-store path: refill it
    if (leader) load(it + kCxStages);
  }
  if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every store has landed before the kernel ends
}

// ---- X1, copy-engine variant: the blocks are pulled by cudaMemcpyAsync (the copy engines move 700 GB/s over
// NVLink, more than SM loads reach) and indexed block by block on the SMs while later blocks still travel.
// A thread indexes two consecutive keys of the block [b0, b1); the block's first key owes its table entries
// to a key of ANOTHER block (which may not have arrived): the seam kernel adds them at the end.
__global__ void __launch_bounds__(256) index_block_kernel(const uint64_t *__restrict__ spins, uint32_t n, uint32_t b0, uint32_t b1,
                                                          uint64_t state_mask, int tshift, uint64_t num_buckets,
                                                          uint32_t *__restrict__ starts, int fshift, uint2 *__restrict__ filter) {
  const uint32_t i = b0 + 2u * (blockIdx.x * 256u + threadIdx.x);
  if (i >= b1) return;
  const uint64_t last = num_buckets;
  const bool two = i + 1 < b1;
  const uint64_t k0 = spins[i], k1 = two ? spins[i + 1] : 0ull;
  const uint64_t bk0 = (k0 & ~state_mask) ? last : k0 >> tshift;
  if (i > b0 || i == 0) {
    uint64_t prev = 0;
    if (i > 0) {
      const uint64_t pk = spins[i - 1];
      prev = ((pk & ~state_mask) ? last : pk >> tshift) + 1;
    }
    for (uint64_t k = prev; k <= bk0 && k <= last; ++k) starts[k] = i;
  }
  uint64_t b_last = bk0;
  if (two) {
    const uint64_t bk1 = (k1 & ~state_mask) ? last : k1 >> tshift;
    for (uint64_t k = bk0 + 1; k <= bk1 && k <= last; ++k) starts[k] = i + 1;
    b_last = bk1;
  }
  if ((two ? i + 1 : i) == n - 1)
    for (uint64_t k = b_last + 1; k <= last; ++k) starts[k] = n;
  const bool in0 = (k0 & ~state_mask) == 0, in1 = two && (k1 & ~state_mask) == 0;
  unsigned long long bits0 = filter_bits(filter_hash(k0));
  const uint64_t w0 = k0 >> fshift;
  if (in1) {
    const uint64_t w1 = k1 >> fshift;
    const unsigned long long bits1 = filter_bits(filter_hash(k1));
    if (in0 && w1 == w0)
      bits0 |= bits1;
    else
      atomicOr(reinterpret_cast<unsigned long long *>(filter + w1), bits1);
  }
  if (in0) atomicOr(reinterpret_cast<unsigned long long *>(filter + w0), bits0);
}

struct SeamArgs {
  uint32_t first[kGxMaxRanks];  // first key of every non-empty block but the one that starts at 0 (n: none)
};
__global__ void index_seam_kernel(const uint64_t *__restrict__ spins, uint32_t n, const SeamArgs seams, int world, uint64_t state_mask,
                                  int tshift, uint64_t num_buckets, uint32_t *__restrict__ starts) {
  const int q = threadIdx.x;
  if (q >= world) return;
  const uint32_t i = seams.first[q];
  if (i == 0 || i >= n) return;
  const uint64_t last = num_buckets;
  const uint64_t key = spins[i], pk = spins[i - 1];
  const uint64_t b = (key & ~state_mask) ? last : key >> tshift;
  for (uint64_t k = ((pk & ~state_mask) ? last : pk >> tshift) + 1; k <= b && k <= last; ++k) starts[k] = i;
}

__global__ void wait_one_flag_kernel(const unsigned long long *flag, unsigned long long value) { wait_flag_or_trap(flag, value); }

