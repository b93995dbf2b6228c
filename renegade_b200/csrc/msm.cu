// Pippenger multi-scalar multiplication on BN254 G1 for sm_100a.
//
// Replaces what the reference reaches through jf-primitives `UnivariateKzgPCS::commit` ->
// ark-ec 0.4.2 `VariableBaseMSM::msm_bigint(&powers_of_g, &coeffs)` (call sites
// /root/reference/crates/circuits/circuit-types/src/traits.rs:850,996 and
// circuits-core/src/zk_circuits/proof_linking/intent_only.rs:42-47; algorithm restated in
// SURVEY.md App. B).  The result is a group element, so after affine normalisation it is
// bit-identical to arkworks' whatever the bucket schedule.
//
// B200-first design (not arkworks' per-window rayon loop):
//   * bases are fixed (the SRS), HBM is 180 GB: at load time every base gets its window
//     multiples 2^(c*j)*P precomputed, so ALL windows share ONE bucket set — no per-window
//     reduction, no Horner doublings on the device; the tables are gathered at 64 B/point
//     (two full sectors), well inside the HBM budget of an integer-pipe-bound kernel;
//   * signed c-bit digits halve the bucket count;
//   * digits are counting-sorted by bucket (histogram -> scan -> scatter, L2-resident atomics),
//     then one thread per bucket folds its points with XYZZ mixed additions (8M+2S, no
//     inversion), prefetching the next 64-byte point while the current addition runs;
//   * bucket sums are folded with a chunked running sum + block tree; the last few hundred
//     bytes go to the host, which does the final Horner (only when windows are not fully
//     precomputed) and the single field inversion.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device_ctx.h"
#include "ec.cuh"

namespace b200 {

namespace {

constexpr uint32_t kIdxBits = 26;  // entry = sign(1) | table(5) | point index(26)
constexpr uint32_t kIdxMask = (1u << kIdxBits) - 1u;
#ifndef B200_SEG_LEN
#define B200_SEG_LEN 32
#endif
constexpr int kSegLen = B200_SEG_LEN;  // a bucket is folded in segments of at most this many points
constexpr int kCombineSeq = 64;    // buckets with more segments than this take the block-tree path
// Buckets per thread (2^chunk_log) in the running-sum reduction.  A lone MSM is latency-bound there and wants
// short serial chains (4 buckets); inside the prover, where several proofs share the GPU, the reduction's
// OPERATION COUNT is what matters (it competes for issue slots with other proofs' accumulation) and 16
// buckets per thread is faster overall: 247 -> 265 proofs/s, while the 2^20 MSM alone goes 3.15 -> 3.28 ms
// (profiles/r1p_reduce_chunk_seglen_variants.log).
constexpr int kReduceChunkLogLatency = kMsmReduceChunkLogLatency;
constexpr int kReduceThreads = 128;

// ---- scalar digits ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t get_bits(const uint32_t* s, int pos, int c) {
    const int limb = pos >> 5, off = pos & 31;
    uint32_t v = limb < 8 ? s[limb] >> off : 0u;
    if (off + c > 32 && limb + 1 < 8) v |= s[limb + 1] << (32 - off);
    return v & ((1u << c) - 1u);
}

// Calls f(bucket, code) for every non-zero signed digit of scalar i.
template <class F>
__device__ __forceinline__ void for_each_digit(const fe* scalars, size_t i, int montgomery,
                                               const MsmPlan& pl, uint32_t point_idx, F f) {
    fe s = fe_load_ro(scalars + i);
    if (montgomery) s = fe_from_mont<FrCfg>(s);
    uint32_t limbs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) limbs[k] = s.l[k];
    const uint32_t half = 1u << (pl.c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < pl.n_digits; ++w) {
        uint32_t raw = get_bits(limbs, w * pl.c, pl.c) + carry;
        uint32_t neg = 0, mag = raw;
        if (raw > half) {  // digit = raw - 2^c  (magnitude 2^c - raw, possibly 0 when raw = 2^c)
            mag = (1u << pl.c) - raw;
            neg = 1;
            carry = 1;
        } else {
            carry = 0;
        }
        if (mag == 0) continue;
        const uint32_t phys = (uint32_t)(w % pl.n_phys), table = (uint32_t)(w / pl.n_phys);
        const uint32_t bucket = phys * half + (mag - 1u);
        f(bucket, (neg << 31) | (table << kIdxBits) | point_idx);
    }
}

// blockIdx.y = index of the MSM inside a batch (same bases, `stride` scalars apart); every MSM owns
// its own range of `buckets_per_msm` buckets, so the later phases see one big bucket array.
__global__ void msm_count_kernel(const fe* scalars, size_t n, size_t stride, int montgomery, MsmPlan pl,
                                 uint32_t buckets_per_msm, uint32_t* counts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t* my = counts + (size_t)blockIdx.y * buckets_per_msm;
    for_each_digit(scalars + (size_t)blockIdx.y * stride, i, montgomery, pl, (uint32_t)i,
                   [&](uint32_t bucket, uint32_t) { atomicAdd(my + bucket, 1u); });
}

__global__ void msm_scatter_kernel(const fe* scalars, size_t n, size_t stride, int montgomery, MsmPlan pl,
                                   uint32_t base_off, uint32_t buckets_per_msm, uint32_t* cursor,
                                   uint32_t* entries) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t* my = cursor + (size_t)blockIdx.y * buckets_per_msm;
    for_each_digit(scalars + (size_t)blockIdx.y * stride, i, montgomery, pl, (uint32_t)i + base_off,
                   [&](uint32_t bucket, uint32_t code) {
                       const uint32_t pos = atomicAdd(my + bucket, 1u);
                       entries[pos] = code;
                   });
}

// ---- exclusive scan of the bucket histogram (three small kernels) -----------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = wi - ws;  // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1) *total = wi;
    }
    __syncthreads();
    const uint32_t r = incl - v + warp_sums[wid];
    __syncthreads();
    return r;
}

// scan inputs: the bucket histogram itself, or the per-bucket segment count derived from offsets
struct ScanCounts {
    const uint32_t* p;
    __device__ __forceinline__ uint32_t operator()(size_t i) const { return p[i]; }
};
struct ScanSegCounts {  // ceil(bucket size / kSegLen)
    const uint32_t* offsets;
    __device__ __forceinline__ uint32_t operator()(size_t i) const {
        return (offsets[i + 1] - offsets[i] + (uint32_t)kSegLen - 1u) / (uint32_t)kSegLen;
    }
};

template <class In>
__global__ void scan_chunk_sums_kernel(In in, size_t n, uint32_t* chunk_sums) {
    __shared__ uint32_t total;
    const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += in(base + k);
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = total;
}

__global__ void scan_chunk_offsets_kernel(uint32_t* chunk_sums, size_t n_chunks) {
    // single block; serial over tiles of kScanThreads chunks
    __shared__ uint32_t total;
    uint32_t running = 0;
    for (size_t base = 0; base < n_chunks; base += kScanThreads) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < n_chunks ? chunk_sums[i] : 0u;
        const uint32_t ex = block_exclusive_scan(v, &total);
        if (i < n_chunks) chunk_sums[i] = running + ex;
        running += total;
        __syncthreads();
    }
}

template <class In>
__global__ void scan_apply_kernel(In in, size_t n, const uint32_t* chunk_offsets,
                                  uint32_t* out /* n + 1 */) {
    __shared__ uint32_t total;
    const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? in(base + k) : 0u;
        s += v[k];
    }
    uint32_t ex = block_exclusive_scan(s, &total) + chunk_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}

// ---- bucket accumulation ---------------------------------------------------------------------------
// Buckets are folded in SEGMENTS of at most kSegLen sorted entries (a bucket of k entries is cut
// into ceil(k / kSegLen) near-equal segments), one thread per segment, so the work per thread is
// bounded whatever the digit distribution: the top window of a 254-bit scalar, small witness
// values or repeated scalars all pile points into a few buckets.
__device__ __forceinline__ g1_affine load_entry_point(const g1_affine* tables, size_t n,
                                                      uint32_t code) {
    const size_t idx = code & kIdxMask, table = (code >> kIdxBits) & 31u;
    return g1_affine_load_ro(tables + table * n + idx);
}

// seg_bucket[s] = bucket owning segment s
__global__ void msm_segfill_kernel(const uint32_t* __restrict__ seg_offsets, uint32_t n_buckets,
                                   uint32_t* __restrict__ seg_bucket) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    const uint32_t s0 = seg_offsets[b], s1 = seg_offsets[b + 1];
    for (uint32_t s = s0; s < s1; ++s) seg_bucket[s] = b;
}

// Segment lengths differ (Poisson bucket sizes); lanes of a warp that fold different numbers of
// points idle in lockstep.  A counting sort of the segment ids by length (kSegLen+1 bins, longest
// first) hands every warp segments of equal length.
__device__ __forceinline__ uint32_t segment_length(const uint32_t* offsets, const uint32_t* seg_offsets, uint32_t b,
                                                   uint32_t s) {
    const uint32_t j = s - seg_offsets[b], k = seg_offsets[b + 1] - seg_offsets[b];
    const uint32_t cnt = offsets[b + 1] - offsets[b];
    return cnt / k + (j < cnt % k ? 1u : 0u);
}
__global__ void msm_seglen_hist_kernel(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ seg_offsets,
                                       const uint32_t* __restrict__ seg_bucket, uint32_t n_buckets,
                                       uint32_t* __restrict__ hist /* kSegLen + 1 bins */) {
    __shared__ uint32_t sh[kSegLen + 1];
    if (threadIdx.x <= kSegLen) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < seg_offsets[n_buckets]) atomicAdd(&sh[segment_length(offsets, seg_offsets, seg_bucket[s], s)], 1u);
    __syncthreads();
    if (threadIdx.x <= kSegLen && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
// hist -> start offsets, longest length first (single small block)
__global__ void msm_seglen_starts_kernel(uint32_t* hist) {
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int l = kSegLen; l >= 0; --l) {
            const uint32_t c = hist[l];
            hist[l] = run;
            run += c;
        }
    }
}
__global__ void msm_seglen_scatter_kernel(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ seg_offsets,
                                          const uint32_t* __restrict__ seg_bucket, uint32_t n_buckets,
                                          uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= seg_offsets[n_buckets]) return;
    const uint32_t len = segment_length(offsets, seg_offsets, seg_bucket[s], s);
    // warp-aggregated: lanes with the same length take consecutive slots with one atomic
    const uint32_t active = __activemask();
    const uint32_t peers = __match_any_sync(active, len);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&cursor[len], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + __popc(peers & ((1u << (threadIdx.x & 31)) - 1u))] = s;
}

__global__ void __launch_bounds__(128) msm_accumulate_kernel(const uint32_t* __restrict__ entries,
                                                             const uint32_t* __restrict__ offsets,
                                                             const uint32_t* __restrict__ seg_offsets,
                                                             const uint32_t* __restrict__ seg_bucket,
                                                             const uint32_t* __restrict__ order,
                                                             const g1_affine* __restrict__ tables,
                                                             size_t n_points, uint32_t n_buckets,
                                                             g1_xyzz* __restrict__ seg_sums) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= seg_offsets[n_buckets]) return;  // the grid is sized for the worst case
    const uint32_t s = order[tid];
    const uint32_t b = seg_bucket[s];
    const uint32_t j = s - seg_offsets[b], k = seg_offsets[b + 1] - seg_offsets[b];
    const uint32_t first = offsets[b], cnt = offsets[b + 1] - first;
    const uint32_t base = cnt / k, rem = cnt % k;  // near-equal split of the bucket
    uint32_t e = first + j * base + min(j, rem);
    const uint32_t end = e + base + (j < rem ? 1u : 0u);
    g1_xyzz acc = g1_xyzz_inf();
    if (e < end) {
        uint32_t code = entries[e];
        g1_affine next = load_entry_point(tables, n_points, code);
        while (true) {
            g1_affine cur = next;
            const uint32_t cur_code = code;
            ++e;
            if (e < end) {  // prefetch the next point under the current addition
                code = entries[e];
                next = load_entry_point(tables, n_points, code);
            }
            if (cur_code >> 31) cur.y = fe_neg<FqCfg>(cur.y);
            acc = g1_add_mixed(acc, cur);
            if (e >= end) break;
        }
    }
    g1_xyzz_store(seg_sums + s, acc);
}

// The reduction-side kernels run one warp per SM sub-partition on long dependent chains; with
// g1_add inlined a dozen times their code no longer fits the instruction cache (ncu: 0.67
// "no instruction" stalls per issue).  Out-of-line copies keep those kernels small.
__device__ __noinline__ g1_xyzz xyzz_add(const g1_xyzz& a, const g1_xyzz& b) { return g1_add(a, b); }
__device__ __noinline__ g1_xyzz xyzz_dbl(const g1_xyzz& a) { return g1_dbl(a); }

// bucket = sum of its segment sums: sequential for ordinary buckets, deferred to a block tree
// (msm_heavy_combine_kernel) for buckets cut into more than kCombineSeq segments.
__global__ void __launch_bounds__(128) msm_bucket_combine_kernel(const g1_xyzz* __restrict__ seg_sums,
                                                                 const uint32_t* __restrict__ seg_offsets,
                                                                 uint32_t n_buckets,
                                                                 g1_xyzz* __restrict__ buckets,
                                                                 uint32_t* __restrict__ heavy_count,
                                                                 uint32_t* __restrict__ heavy_list) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    const uint32_t s0 = seg_offsets[b], s1 = seg_offsets[b + 1];
    if (s1 - s0 > (uint32_t)kCombineSeq) {
        heavy_list[atomicAdd(heavy_count, 1u)] = b;
        return;
    }
    g1_xyzz acc = g1_xyzz_inf();
    if (s0 < s1) {
        acc = g1_xyzz_load(seg_sums + s0);
        for (uint32_t s = s0 + 1; s < s1; ++s) acc = xyzz_add(acc, g1_xyzz_load(seg_sums + s));
    }
    g1_xyzz_store(buckets + b, acc);
}

__global__ void __launch_bounds__(kReduceThreads) msm_heavy_combine_kernel(const g1_xyzz* __restrict__ seg_sums,
                                                                           const uint32_t* __restrict__ seg_offsets,
                                                                           const uint32_t* __restrict__ heavy_count,
                                                                           const uint32_t* __restrict__ heavy_list,
                                                                           g1_xyzz* __restrict__ buckets) {
    __shared__ g1_xyzz sh[kReduceThreads];
    for (uint32_t h = blockIdx.x; h < *heavy_count; h += gridDim.x) {
        const uint32_t b = heavy_list[h];
        const uint32_t s0 = seg_offsets[b], s1 = seg_offsets[b + 1];
        g1_xyzz acc = g1_xyzz_inf();
        for (uint32_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) acc = xyzz_add(acc, g1_xyzz_load(seg_sums + s));
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int stride = kReduceThreads / 2; stride > 0; stride >>= 1) {
            if ((int)threadIdx.x < stride) sh[threadIdx.x] = xyzz_add(sh[threadIdx.x], sh[threadIdx.x + stride]);
            __syncthreads();
        }
        if (threadIdx.x == 0) g1_xyzz_store(buckets + b, sh[0]);
        __syncthreads();
    }
}

// ---- bucket reduction: sum_k (k+1) * B[k] per physical window -----------------------------------
// These kernels are latency-bound (one warp per SM sub-partition walking a chain of dependent XYZZ
// operations), so the formulation minimises the LENGTH of that chain, not the operation count:
//   thread t = bx*T + tid owns buckets [tK, (t+1)K): running sums give S_t = sum B and
//   A_t = sum (i+1) B[tK+i]; the window total is sum_t A_t + K * sum_t t * S_t.
//   sum_t t*S_t = T * sum_bx bx * (sum_tid S) + sum_bx sum_tid tid * S, and a weighted sum
//   sum_i i*X_i equals the sum of the inclusive suffix sums R_1 + R_2 + ... — a log-depth block scan
//   instead of a per-thread double-and-add by the chunk index (2(c-1) dependent operations).
// Chain per window: 2K (chunk) + log T (scan) + log K + 1 + log T (tree), then the same scan + tree
// once more over the per-block totals in msm_reduce_final_kernel.
constexpr int ilog2_c(unsigned v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }
constexpr int kReduceThreadsLog = ilog2_c(kReduceThreads);
static_assert((1 << kReduceThreadsLog) == kReduceThreads, "reduction block size is a power of two");

// v_tid <- sum_{j >= tid} v_j over the block (Hillis–Steele); entries at index >= valid are the identity
__device__ __forceinline__ g1_xyzz block_suffix_scan(g1_xyzz v, g1_xyzz* sh, uint32_t valid) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < (uint32_t)kReduceThreads && d < valid; d <<= 1) {
        const bool has = threadIdx.x + d < (uint32_t)kReduceThreads;
        g1_xyzz o = g1_xyzz_inf();
        if (has) o = sh[threadIdx.x + d];
        __syncthreads();
        v = xyzz_add(v, o);
        sh[threadIdx.x] = v;
        __syncthreads();
    }
    return v;
}

// partials[(window * gridDim.x + bx) * 2 + {0, 1}] = { V_bx, (K*T) * sum_tid S }  with
// V_bx = sum_tid (A_tid + K * R_tid [tid >= 1])
__global__ void __launch_bounds__(kReduceThreads) msm_reduce_kernel(const g1_xyzz* __restrict__ buckets,
                                                                    uint32_t buckets_per_window, int chunk_log,
                                                                    g1_xyzz* __restrict__ partials) {
    __shared__ g1_xyzz sh[kReduceThreads];
    const uint32_t window = blockIdx.y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t chunk = 1u << chunk_log;
    const uint32_t first = t << chunk_log;
    const uint32_t threads_needed = (buckets_per_window + chunk - 1) >> chunk_log;
    const uint32_t valid = min((uint32_t)kReduceThreads, threads_needed - blockIdx.x * blockDim.x);
    g1_xyzz run = g1_xyzz_inf(), acc = g1_xyzz_inf();
    if (first < buckets_per_window) {
        const g1_xyzz* B = buckets + (size_t)window * buckets_per_window + first;
        const uint32_t cnt = min(chunk, buckets_per_window - first);
        for (int i = (int)cnt - 1; i >= 0; --i) {
            run = xyzz_add(run, g1_xyzz_load(B + i));
            acc = xyzz_add(acc, run);
        }
    }
    const g1_xyzz R = block_suffix_scan(run, sh, valid);
    g1_xyzz V = acc;
    if (threadIdx.x >= 1 && threadIdx.x < valid) {
        g1_xyzz kr = R;
#pragma unroll 1
        for (int i = 0; i < chunk_log; ++i) kr = xyzz_dbl(kr);
        V = xyzz_add(V, kr);
    }
    // the block total, scaled by K*T, is produced by the last thread (its warp has no part in the tree
    // below) while the first warps fold V: its doublings are spread over the tree levels
    g1_xyzz Tp = g1_xyzz_inf();
    if (threadIdx.x == kReduceThreads - 1) Tp = sh[0];
    __syncthreads();
    sh[threadIdx.x] = V;
    __syncthreads();
    int dleft = chunk_log + kReduceThreadsLog, levels_left = kReduceThreadsLog;
    for (int stride = kReduceThreads / 2; stride > 0; stride >>= 1, --levels_left) {
        if ((int)threadIdx.x < stride) sh[threadIdx.x] = xyzz_add(sh[threadIdx.x], sh[threadIdx.x + stride]);
        if (threadIdx.x == kReduceThreads - 1) {
            const int per = (dleft + levels_left - 1) / levels_left;
            for (int i = 0; i < per; ++i) Tp = xyzz_dbl(Tp);
            dleft -= per;
        }
        __syncthreads();
    }
    g1_xyzz* out = partials + ((size_t)window * gridDim.x + blockIdx.x) * 2;
    if (threadIdx.x == 0) g1_xyzz_store(out, sh[0]);
    if (threadIdx.x == kReduceThreads - 1) g1_xyzz_store(out + 1, Tp);
}

// window sum = sum_b V_b + sum_b b * T'_b.  Thread tid owns blocks tid, tid + T, ...: with
// p = sum_m X_m and q = sum_m m * X_m (running sums), sum_b b * X_b = sum_{tid >= 1} R_tid + T * sum_tid q_tid.
__global__ void __launch_bounds__(kReduceThreads) msm_reduce_final_kernel(const g1_xyzz* __restrict__ partials,
                                                                          uint32_t n_blocks,
                                                                          g1_xyzz* __restrict__ window_sums) {
    __shared__ g1_xyzz sh[kReduceThreads];
    const uint32_t window = blockIdx.x;
    const g1_xyzz* P = partials + (size_t)window * n_blocks * 2;
    const int rounds = (int)((n_blocks + kReduceThreads - 1) / kReduceThreads);
    g1_xyzz vsum = g1_xyzz_inf(), p = g1_xyzz_inf(), q = g1_xyzz_inf();
    for (int m = rounds - 1; m >= 0; --m) {
        const uint32_t j = threadIdx.x + (uint32_t)m * kReduceThreads;
        if (j < n_blocks) {
            vsum = xyzz_add(vsum, g1_xyzz_load(P + 2 * (size_t)j));
            p = xyzz_add(p, g1_xyzz_load(P + 2 * (size_t)j + 1));
        }
        if (m > 0) q = xyzz_add(q, p);
    }
    const uint32_t valid = min((uint32_t)kReduceThreads, n_blocks);
    const g1_xyzz R = block_suffix_scan(p, sh, valid);
    g1_xyzz W = vsum;
    if (threadIdx.x >= 1 && threadIdx.x < valid) W = xyzz_add(W, R);
    if (rounds > 1) {
#pragma unroll 1
        for (int i = 0; i < kReduceThreadsLog; ++i) q = xyzz_dbl(q);
        W = xyzz_add(W, q);
    }
    sh[threadIdx.x] = W;
    __syncthreads();
    for (int stride = kReduceThreads / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) sh[threadIdx.x] = xyzz_add(sh[threadIdx.x], sh[threadIdx.x + stride]);
        __syncthreads();
    }
    if (threadIdx.x == 0) g1_xyzz_store(window_sums + window, sh[0]);
}

// ---- window tables: table[j][i] = 2^(shift*j) * P_i (affine) -------------------------------------
__global__ void msm_table_kernel(const g1_affine* __restrict__ prev, g1_affine* __restrict__ next,
                                 size_t n, int shift) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const g1_affine p = g1_affine_load_ro(prev + i);
    g1_xyzz a = g1_dbl_affine(p);
    for (int k = 1; k < shift; ++k) a = g1_dbl(a);
    g1_affine_store(next + i, g1_to_affine(a));
}

__global__ void g1_on_curve_kernel(const g1_affine* pts, size_t n, uint32_t* bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!g1_affine_on_curve(g1_affine_load_ro(pts + i))) atomicAdd(bad, 1u);
}

// ---- synthetic inputs: SplitMix64 field elements, known-dlog bases ---------------
__device__ __forceinline__ uint64_t splitmix_at(uint64_t seed, uint64_t j) {
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ fe splitmix_fr_canon(uint64_t seed, uint64_t i) {
    fe v;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t x = splitmix_at(seed, 4 * i + k);
        v.l[2 * k] = (uint32_t)x;
        v.l[2 * k + 1] = (uint32_t)(x >> 32);
    }
    // 2^256 / r < 6: a few conditional subtractions reduce mod r
    for (int it = 0; it < 6; ++it) {
        bool ge = true;
        for (int k = 7; k >= 0; --k) {
            if (v.l[k] != FrCfg::mod(k)) { ge = v.l[k] > FrCfg::mod(k); break; }
        }
        if (!ge) break;
        uint64_t borrow = 0;
        for (int k = 0; k < 8; ++k) {
            const uint64_t d = (uint64_t)v.l[k] - FrCfg::mod(k) - borrow;
            v.l[k] = (uint32_t)d;
            borrow = (d >> 32) & 1u;
        }
    }
    return v;
}
__global__ void splitmix_fr_kernel(uint64_t seed, size_t first, size_t n, int montgomery, fe* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe v = splitmix_fr_canon(seed, first + i);
    if (montgomery) v = fe_to_mont<FrCfg>(v);
    fe_store(out + i, v);
}
__global__ void known_dlog_bases_kernel(uint64_t seed, size_t first, size_t n, g1_affine* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fe a = splitmix_fr_canon(seed, first + i);
    g1_affine g;
    g.x = fe_one<FqCfg>();
    g.y = fe_from_u32<FqCfg>(2);
    g1_affine_store(out + i, g1_to_affine(g1_mul_bits(g, a, 254)));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
MsmPlan msm_choose_plan(size_t n, int c_override, size_t mem_budget_bytes) {
    MsmPlan best;
    double best_cost = 1e300;
    const int c_lo = c_override > 0 ? c_override : 6, c_hi = c_override > 0 ? c_override : 23;
    for (int c = c_lo; c <= c_hi; ++c) {
        const int W = (255 + c - 1) / c;
        if (W > 32) continue;
        // Cost model fitted to B200 measurements (profiles/r1b_msm_window_sweep.log), in ms:
        // bucket phase ~ n*W mixed additions at ~5.8e6/ms, digit sort ~ n*W at ~45e6/ms, bucket
        // reduction ~ a latency floor + 2^(c-1) buckets.  `shrt` = how many bits the top digit of
        // a 254-bit scalar falls short of a full window: a short top digit piles n / 2^t points
        // into 2^t buckets (hot atomics in the sort, more segments to combine).
        const int t = 254 - c * (W - 1);
        const int shrt = t >= c - 3 ? 0 : (c - 3 - t);  // up to 2 bits short showed no penalty
        const double nw = (double)n * W;
        const double cost = nw * (1.0 + 0.02 * shrt) / 5.8e6 + nw * (1.0 + 0.1 * shrt) / 45e6 + 0.42 +
                            (double)((size_t)1 << (c - 1)) / 2.0e6;
        if (cost < best_cost) {
            best_cost = cost;
            best.c = c;
            best.n_digits = W;
        }
    }
    // full precompute (one physical window) unless the tables exceed the memory budget;
    // B200_MSM_PHYS_WINDOWS forces a split (tuning / test knob for the host Horner path)
    int phys = 1;
    if (const char* env = std::getenv("B200_MSM_PHYS_WINDOWS")) {
        const int v = std::atoi(env);
        if (v >= 1 && v <= best.n_digits) phys = v;
    }
    while (true) {
        const int tables = (best.n_digits + phys - 1) / phys;
        if ((double)tables * (double)n * 64.0 <= (double)mem_budget_bytes || phys >= best.n_digits) {
            best.n_phys = phys;
            best.n_tables = tables;
            break;
        }
        ++phys;
    }
    return best;
}

static int build_tables(Bases* b, cudaStream_t st) {
    const int shift = b->plan.c * b->plan.n_phys;
    const unsigned bs = 128;
    const unsigned grid = (unsigned)((b->n + bs - 1) / bs);
    for (int j = 1; j < b->plan.n_tables; ++j) {
        msm_table_kernel<<<grid, bs, 0, st>>>(b->tables + (size_t)(j - 1) * b->n,
                                              b->tables + (size_t)j * b->n, b->n, shift);
    }
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(st));
    return B200_OK;
}

static size_t table_mem_budget() {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (size_t)8 << 30;
    return free_b / 2;
}

int bases_create_device(const g1_affine* d_points, size_t n, int c_override, cudaStream_t st,
                        Bases** out) {
    if (n == 0 || n > ((size_t)1 << kIdxBits)) {
        set_error("bases: n must be in [1, 2^26]");
        return B200_ERR_INVALID;
    }
    Bases* b = new Bases();
    b->n = n;
    b->plan = msm_choose_plan(n, c_override, table_mem_budget());
    cudaError_t e = cudaMalloc(&b->tables, (size_t)b->plan.n_tables * n * sizeof(g1_affine));
    if (e != cudaSuccess) {
        delete b;
        return cuda_fail(e, "cudaMalloc(window tables)");
    }
    e = cudaMemcpyAsync(b->tables, d_points, n * sizeof(g1_affine), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
        delete b;
        return cuda_fail(e, "cudaMemcpyAsync(bases)");
    }
    int rc = build_tables(b, st);
    if (rc != B200_OK) {
        delete b;
        return rc;
    }
    *out = b;
    return B200_OK;
}

int bases_create(const g1_affine* h_points, size_t n, int c_override, int check_on_curve,
                 cudaStream_t st, Bases** out) {
    if (n == 0 || n > ((size_t)1 << kIdxBits)) {
        set_error("bases: n must be in [1, 2^26]");
        return B200_ERR_INVALID;
    }
    g1_affine* d_pts = nullptr;
    B200_CUDA(cudaMalloc(&d_pts, n * sizeof(g1_affine)));
    cudaError_t e = cudaMemcpyAsync(d_pts, h_points, n * sizeof(g1_affine), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        cudaFree(d_pts);
        return cuda_fail(e, "cudaMemcpyAsync(bases H2D)");
    }
    if (check_on_curve) {  // srs.rs:178-179 asserts every SRS point is on the curve
        uint32_t* d_bad = nullptr;
        uint32_t h_bad = 0;
        cudaMalloc(&d_bad, 4);
        cudaMemsetAsync(d_bad, 0, 4, st);
        g1_on_curve_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_pts, n, d_bad);
        cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st);
        e = cudaStreamSynchronize(st);
        cudaFree(d_bad);
        if (e != cudaSuccess) {
            cudaFree(d_pts);
            return cuda_fail(e, "on-curve check");
        }
        if (h_bad) {
            cudaFree(d_pts);
            set_error("point not on curve");
            return B200_ERR_NOT_ON_CURVE;
        }
    }
    int rc = bases_create_device(d_pts, n, c_override, st, out);
    cudaFree(d_pts);
    return rc;
}

template <class In>
static int exclusive_scan_u32(In d_in, size_t n, uint32_t* d_out, DevBuf* chunk_buf,
                              cudaStream_t st) {
    const size_t n_chunks = (n + kScanChunk - 1) / kScanChunk;
    int rc = chunk_buf->reserve(n_chunks * sizeof(uint32_t));
    if (rc != B200_OK) return rc;
    uint32_t* chunk = (uint32_t*)chunk_buf->p;
    scan_chunk_sums_kernel<In><<<(unsigned)n_chunks, kScanThreads, 0, st>>>(d_in, n, chunk);
    scan_chunk_offsets_kernel<<<1, kScanThreads, 0, st>>>(chunk, n_chunks);
    scan_apply_kernel<In><<<(unsigned)n_chunks, kScanThreads, 0, st>>>(d_in, n, chunk, d_out);
    return B200_OK;
}

// `batch` MSMs over the same bases (scalar vectors `stride` elements apart) in one pass: the
// digit sort, the bucket folding and the reduction each run once over batch * buckets buckets,
// which is what fills 148 SMs at the prover's 2^16-point sizes.
int msm_device_batch(const Bases* b, size_t base_off, const fe* d_scalars, size_t n, size_t stride,
                     unsigned batch, int montgomery, MsmScratch* s, cudaStream_t st, g1_affine* out,
                     int* out_inf) {
    if (base_off + n > b->n) {
        set_error("msm: base_off + n exceeds the loaded bases");
        return B200_ERR_INVALID;
    }
    if (batch == 0) return B200_OK;
    if (n == 0) {
        for (unsigned i = 0; i < batch; ++i) {
            out[i].x = fe_zero();
            out[i].y = fe_zero();
            if (out_inf) out_inf[i] = 1;
        }
        return B200_OK;
    }
    const MsmPlan& pl = b->plan;
    const uint32_t half = 1u << (pl.c - 1);
    const size_t buckets_per_msm = (size_t)pl.n_phys * half;
    const size_t n_buckets = buckets_per_msm * batch;
    const size_t n_windows = (size_t)pl.n_phys * batch;
    const size_t max_entries = n * (size_t)pl.n_digits * batch;
    if (max_entries >= ((size_t)1 << 32) || n_buckets >= ((size_t)1 << 31)) {
        set_error("msm: batch * n * windows exceeds 2^32 entries");
        return B200_ERR_INVALID;
    }
    int rc;
    if ((rc = s->counts.reserve(n_buckets * 4)) != B200_OK) return rc;
    if ((rc = s->offsets.reserve((n_buckets + 1) * 4)) != B200_OK) return rc;
    if ((rc = s->cursor.reserve(n_buckets * 4)) != B200_OK) return rc;
    if ((rc = s->entries.reserve(max_entries * 4)) != B200_OK) return rc;
    if ((rc = s->buckets.reserve(n_buckets * sizeof(g1_xyzz))) != B200_OK) return rc;
    // segments: at most floor(entries / kSegLen) full ones plus one partial per bucket
    const size_t max_segs = max_entries / kSegLen + n_buckets;
    const size_t max_heavy = max_segs / (kCombineSeq + 1) + 1;
    if ((rc = s->seg_offsets.reserve((n_buckets + 1) * 4)) != B200_OK) return rc;
    if ((rc = s->seg_bucket.reserve(max_segs * 4)) != B200_OK) return rc;
    if ((rc = s->seg_sums.reserve(max_segs * sizeof(g1_xyzz))) != B200_OK) return rc;
    if ((rc = s->heavy.reserve((max_heavy + 1) * 4)) != B200_OK) return rc;
    if ((rc = s->seg_order.reserve((max_segs + kSegLen + 2) * 4)) != B200_OK) return rc;
    // lone MSM: short chains while the reduction is latency-bound (<= 2^16 buckets -> <= 128 blocks); with more
    // buckets the blocks outnumber the SMs and the operation count takes over, as inside the prover
    int chunk_log = s->reduce_chunk_log > 0 ? s->reduce_chunk_log
                                            : std::min(kMsmReduceChunkLogThroughput, std::max(kReduceChunkLogLatency, pl.c - 15));
    while (chunk_log > 0 && (half >> chunk_log) < 32) --chunk_log;  // tiny windows: keep a warp's worth of threads
    const uint32_t reduce_threads_needed = (half + (1u << chunk_log) - 1) >> chunk_log;
    const uint32_t reduce_blocks = (reduce_threads_needed + kReduceThreads - 1) / kReduceThreads;
    if ((rc = s->partials.reserve(n_windows * reduce_blocks * 2 * sizeof(g1_xyzz))) != B200_OK) return rc;
    if ((rc = s->window_sums.reserve(n_windows * sizeof(g1_xyzz))) != B200_OK) return rc;

    uint32_t* counts = (uint32_t*)s->counts.p;
    uint32_t* offsets = (uint32_t*)s->offsets.p;
    uint32_t* cursor = (uint32_t*)s->cursor.p;
    uint32_t* entries = (uint32_t*)s->entries.p;
    g1_xyzz* buckets = (g1_xyzz*)s->buckets.p;
    g1_xyzz* partials = (g1_xyzz*)s->partials.p;
    g1_xyzz* window_sums = (g1_xyzz*)s->window_sums.p;
    uint32_t* seg_offsets = (uint32_t*)s->seg_offsets.p;
    uint32_t* seg_bucket = (uint32_t*)s->seg_bucket.p;
    g1_xyzz* seg_sums = (g1_xyzz*)s->seg_sums.p;
    uint32_t* heavy_count = (uint32_t*)s->heavy.p;
    uint32_t* heavy_list = heavy_count + 1;
    uint32_t* seg_hist = (uint32_t*)s->seg_order.p;  // kSegLen + 1 bins, then the ordering itself
    uint32_t* seg_order = seg_hist + kSegLen + 2;

    if (s->timing && !s->ev_init) {
        for (auto& e : s->ev) B200_CUDA(cudaEventCreate(&e));
        s->ev_init = true;
    }
    if (s->timing) cudaEventRecord(s->ev[0], st);
    B200_CUDA(cudaMemsetAsync(counts, 0, n_buckets * 4, st));
    const unsigned bs = 256;
    const dim3 grid_n((unsigned)((n + bs - 1) / bs), batch);
    msm_count_kernel<<<grid_n, bs, 0, st>>>(d_scalars, n, stride, montgomery, pl, (uint32_t)buckets_per_msm, counts);
    if ((rc = exclusive_scan_u32(ScanCounts{counts}, n_buckets, offsets, &s->block_sums, st)) != B200_OK) return rc;
    B200_CUDA(cudaMemcpyAsync(cursor, offsets, n_buckets * 4, cudaMemcpyDeviceToDevice, st));
    msm_scatter_kernel<<<grid_n, bs, 0, st>>>(d_scalars, n, stride, montgomery, pl, (uint32_t)base_off,
                                              (uint32_t)buckets_per_msm, cursor, entries);
    // the long kernels of the MSM go to the low-priority companion stream (see b200_init)
    cudaStream_t hv = s->hv_stream ? s->hv_stream : st;
    auto fork = [&]() {
        if (s->hv_stream) {
            cudaEventRecord(s->hv_fork, st);
            cudaStreamWaitEvent(hv, s->hv_fork, 0);
        }
    };
    auto join = [&]() {
        if (s->hv_stream) {
            cudaEventRecord(s->hv_join, hv);
            cudaStreamWaitEvent(st, s->hv_join, 0);
        }
    };
    if ((rc = exclusive_scan_u32(ScanSegCounts{offsets}, n_buckets, seg_offsets, &s->block_sums, st)) != B200_OK) return rc;
    msm_segfill_kernel<<<(unsigned)((n_buckets + 255) / 256), 256, 0, st>>>(seg_offsets, (uint32_t)n_buckets, seg_bucket);
    B200_CUDA(cudaMemsetAsync(heavy_count, 0, 4, st));
    B200_CUDA(cudaMemsetAsync(seg_hist, 0, (kSegLen + 2) * 4, st));
    const unsigned seg_grid = (unsigned)((max_segs + 255) / 256);
    msm_seglen_hist_kernel<<<seg_grid, 256, 0, st>>>(offsets, seg_offsets, seg_bucket, (uint32_t)n_buckets, seg_hist);
    msm_seglen_starts_kernel<<<1, 32, 0, st>>>(seg_hist);
    msm_seglen_scatter_kernel<<<seg_grid, 256, 0, st>>>(offsets, seg_offsets, seg_bucket, (uint32_t)n_buckets, seg_hist, seg_order);
    if (s->timing) cudaEventRecord(s->ev[1], st);
    fork();
    msm_accumulate_kernel<<<(unsigned)((max_segs + 127) / 128), 128, 0, hv>>>(
        entries, offsets, seg_offsets, seg_bucket, seg_order, b->tables, b->n, (uint32_t)n_buckets, seg_sums);
    join();
    msm_bucket_combine_kernel<<<(unsigned)((n_buckets + 127) / 128), 128, 0, st>>>(
        seg_sums, seg_offsets, (uint32_t)n_buckets, buckets, heavy_count, heavy_list);
    msm_heavy_combine_kernel<<<(unsigned)std::min<size_t>(max_heavy, 4096), kReduceThreads, 0, st>>>(
        seg_sums, seg_offsets, heavy_count, heavy_list, buckets);
    if (s->timing) cudaEventRecord(s->ev[2], st);
    msm_reduce_kernel<<<dim3(reduce_blocks, (unsigned)n_windows), kReduceThreads, 0, st>>>(buckets, half, chunk_log, partials);
    msm_reduce_final_kernel<<<(unsigned)n_windows, kReduceThreads, 0, st>>>(partials, reduce_blocks, window_sums);
    B200_CUDA(cudaGetLastError());

    std::vector<g1_xyzz> h_sums(n_windows);
    if (s->timing) cudaEventRecord(s->ev[3], st);
    B200_CUDA(cudaMemcpyAsync(h_sums.data(), window_sums, n_windows * sizeof(g1_xyzz), cudaMemcpyDeviceToHost, st));
    if (s->timing) cudaEventRecord(s->ev[4], st);
    B200_CUDA(cudaStreamSynchronize(st));
    if (s->timing) {
        cudaEventElapsedTime(&s->ms[0], s->ev[0], s->ev[4]);
        cudaEventElapsedTime(&s->ms[1], s->ev[0], s->ev[1]);
        cudaEventElapsedTime(&s->ms[2], s->ev[1], s->ev[2]);
        cudaEventElapsedTime(&s->ms[3], s->ev[2], s->ev[3]);
        s->tot_acc_ms += s->ms[2];
        s->tot_pairs += (double)n * batch;
        s->tot_launches += 1;
    }

    // host epilogue: per MSM a Horner over the physical windows (none when fully precomputed), then
    // ONE field inversion for the whole batch (Montgomery's trick over the ZZZ coordinates) — a few
    // hundred bytes of work.
    std::vector<g1_xyzz> totals(batch);
    std::vector<fe> prefix(batch);
    fe run = fe_one<FqCfg>();
    for (unsigned i = 0; i < batch; ++i) {
        const g1_xyzz* hs = h_sums.data() + (size_t)i * pl.n_phys;
        g1_xyzz total = hs[pl.n_phys - 1];
        for (int p = pl.n_phys - 2; p >= 0; --p) {
            for (int k = 0; k < pl.c; ++k) total = g1_dbl(total);
            total = g1_add(total, hs[p]);
        }
        totals[i] = total;
        prefix[i] = run;
        if (!g1_xyzz_is_inf(total)) run = fe_mul<FqCfg>(run, total.zzz);
    }
    fe inv = fe_inv<FqCfg>(run);
    for (unsigned i = batch; i-- > 0;) {
        const g1_xyzz& t = totals[i];
        if (g1_xyzz_is_inf(t)) {
            out[i].x = fe_zero();
            out[i].y = fe_zero();
            if (out_inf) out_inf[i] = 1;
            continue;
        }
        const fe i3 = fe_mul<FqCfg>(inv, prefix[i]);  // 1 / ZZZ_i
        inv = fe_mul<FqCfg>(inv, t.zzz);
        const fe izz = fe_mul<FqCfg>(fe_sqr<FqCfg>(t.zz), fe_sqr<FqCfg>(i3));  // 1/ZZ = ZZ^2 / ZZZ^2
        out[i].x = fe_mul<FqCfg>(t.x, izz);
        out[i].y = fe_mul<FqCfg>(t.y, i3);
        if (out_inf) out_inf[i] = 0;
    }
    return B200_OK;
}

int msm_device(const Bases* b, size_t base_off, const fe* d_scalars, size_t n, int montgomery,
               MsmScratch* s, cudaStream_t st, g1_affine* out, int* out_inf) {
    return msm_device_batch(b, base_off, d_scalars, n, n, 1, montgomery, s, st, out, out_inf);
}

int g1_known_dlog_bases_device(uint64_t seed, size_t first, size_t n, g1_affine* d_out,
                               cudaStream_t st) {
    if (n == 0) return B200_OK;
    known_dlog_bases_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seed, first, n, d_out);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int splitmix_fr_device(uint64_t seed, size_t first, size_t n, int montgomery, fe* d_out,
                       cudaStream_t st) {
    if (n == 0) return B200_OK;
    splitmix_fr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seed, first, n, montgomery, d_out);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

}  // namespace b200
