// g2v_paths.cu -- glue between the two hot paths: canonical form of the walks and exact
// duplicate / cross-group detection.
//
// Reference: `path = tuple(sorted(path))` into a Python set (G2Vec.py:345,351) and
// `commonPath = pathSetList[0].intersection(pathSetList[1])` (G2Vec.py:313).
//
//   paths_canon_kernel   one warp per walk row: the <= L node ids are sorted with a bitonic network in
//                        shared memory (padding -> INT32_MAX at the end) and a 64-bit key is formed from
//                        the sorted ids.  Equal rows have equal keys; the converse is NOT assumed.
//   paths_mark_kernel    rows are visited in key order (perm = argsort(key), a radix sort of 8-byte keys
//                        done by the caller).  A row is a duplicate iff an EARLIER row of the same key run
//                        has identical content -- the run is walked backwards with full row comparisons, so
//                        the result is exact even if two different rows ever shared a key.  In cross-group
//                        mode a row is marked iff a row of the OTHER group in its key run is identical.
// Sorting keys instead of rows replaces the lexicographic row sort (80-column merge sort) of the torch
// implementation; compaction and the CSR assembly stay with the caller.
#include "g2v_common.cuh"

namespace g2v {

constexpr int kPathWarps = 8;
constexpr int32_t kPad = kPathPad;

__global__ void __launch_bounds__(kPathWarps * 32)
paths_canon_kernel(const int32_t *__restrict__ nodes, int64_t n, int32_t L, int32_t P,
                   int32_t *__restrict__ sorted, unsigned long long *__restrict__ key) {
    extern __shared__ int32_t sh[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *s = sh + (size_t)warp * P;
    const int64_t nwarps = (int64_t)gridDim.x * kPathWarps;
    for (int64_t r = (int64_t)blockIdx.x * kPathWarps + warp; r < n; r += nwarps) {
        const int32_t *row = nodes + (size_t)r * L;
        for (int i = lane; i < P; i += 32) {
            const int32_t v = i < L ? __ldg(row + i) : kPad;
            s[i] = v < 0 ? kPad : v;                                  // -1 padding of the sampler
        }
        __syncwarp();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < (P >> 1); t += 32) {
                    const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
                    const bool up = (i & k) == 0;
                    const int32_t a = s[i], b = s[l];
                    if ((a > b) == up) { s[i] = b; s[l] = a; }
                }
                __syncwarp();
            }
        uint64_t h = 0;
        for (int i = lane; i < L; i += 32) {
            const int32_t v = s[i];
            sorted[(size_t)r * L + i] = v;
            if (v != kPad) h += path_key_term(v, i);
        }
        h = warp_sum_u64(h);
        if (lane == 0) key[r] = path_key_finish(h);
        __syncwarp();
    }
}

__device__ __forceinline__ bool rows_equal(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int32_t L,
                                           int lane) {
    bool same = true;
    for (int i = lane; i < L; i += 32) same = same && (__ldg(a + i) == __ldg(b + i));
    return __all_sync(0xffffffffu, same);
}

// flag[i] refers to sorted position i (row perm[i]).  group == nullptr: flag = 1 iff the row's content did not
// occur earlier in its key run (first occurrence).  group != nullptr: flag = 1 iff NO row of the other group in
// the key run has the same content (the row survives the cross-group removal).
__global__ void __launch_bounds__(kPathWarps * 32)
paths_mark_kernel(const int32_t *__restrict__ rows, const long long *__restrict__ key_sorted,
                  const long long *__restrict__ perm, const uint8_t *__restrict__ group, int64_t n, int32_t L,
                  uint8_t *__restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kPathWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kPathWarps;
    for (int64_t i = warp; i < n; i += nwarps) {
        const long long k = key_sorted[i];
        const long long me = perm[i];
        const int32_t *mine = rows + (size_t)me * L;
        bool hit = false;
        for (int64_t j = i - 1; !hit && j >= 0 && key_sorted[j] == k; --j) {
            const long long other = perm[j];
            if (group && group[other] == group[me]) continue;
            hit = rows_equal(mine, rows + (size_t)other * L, L, lane);
        }
        if (group)
            for (int64_t j = i + 1; !hit && j < n && key_sorted[j] == k; ++j) {
                const long long other = perm[j];
                if (group[other] == group[me]) continue;
                hit = rows_equal(mine, rows + (size_t)other * L, L, lane);
            }
        if (lane == 0) flag[i] = hit ? 0 : 1;
    }
}

}  // namespace g2v

using namespace g2v;

extern "C" int g2v_paths_canonicalise(const int32_t *nodes, int64_t n, int32_t L, int32_t *sorted, int64_t *key,
                                      void *stream) {
    G2V_REQUIRE(n >= 0 && L >= 1 && L <= 4096, "g2v_paths_canonicalise: bad sizes (n=%lld L=%d)", (long long)n, L);
    if (n == 0) return 0;
    G2V_REQUIRE(nodes && sorted && key, "g2v_paths_canonicalise: null pointer");
    DeviceProps dp;
    if (device_props(&dp)) return 1;
    int P = 32;
    while (P < L) P <<= 1;
    const size_t smem = (size_t)kPathWarps * P * sizeof(int32_t);
    G2V_REQUIRE(smem <= (size_t)dp.max_smem_optin, "g2v_paths_canonicalise: lenPath %d needs %zu B of shared memory", L, smem);
    if (smem > 48 * 1024)
        G2V_CUDA_OK(cudaFuncSetAttribute(paths_canon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n + kPathWarps - 1) / kPathWarps;
    const int64_t cap = (int64_t)dp.sm_count * 8;
    if (blocks > cap) blocks = cap;
    paths_canon_kernel<<<(unsigned)blocks, kPathWarps * 32, smem, (cudaStream_t)stream>>>(
        nodes, n, L, P, sorted, reinterpret_cast<unsigned long long *>(key));
    G2V_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int g2v_paths_mark(const int32_t *rows, const int64_t *key_sorted, const int64_t *perm,
                              const uint8_t *group, int64_t n, int32_t L, uint8_t *flag, void *stream) {
    G2V_REQUIRE(n >= 0 && L >= 1, "g2v_paths_mark: bad sizes");
    if (n == 0) return 0;
    G2V_REQUIRE(rows && key_sorted && perm && flag, "g2v_paths_mark: null pointer");
    DeviceProps dp;
    if (device_props(&dp)) return 1;
    int64_t blocks = (n + kPathWarps - 1) / kPathWarps;
    const int64_t cap = (int64_t)dp.sm_count * 8;
    if (blocks > cap) blocks = cap;
    paths_mark_kernel<<<(unsigned)blocks, kPathWarps * 32, 0, (cudaStream_t)stream>>>(
        rows, reinterpret_cast<const long long *>(key_sorted), reinterpret_cast<const long long *>(perm), group, n, L,
        flag);
    G2V_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

// =====================================================================================================
// Sort-free set pipeline: `pathSet.add(path)` (G2Vec.py:351), `pathSet - commonPath` (:313),
// the multi-hot rows of integrate_pathSet (:316-320, here CSR windows) and count_geneFreq (:288-308) without any
// sort: a hash table on the 64-bit row keys decides which rows survive, a prefix sum places them, and one
// kernel emits the CSR windows and the per-gene label counts.
//
//   paths_insert_kernel   slot(key) <- min row index of the key, per group (atomicCAS on the key, atomicMin on
//                         the index): the FIRST occurrence in input order represents a path
//   paths_flag_kernel     keep[i] = row i is its group's representative AND the other group has no row of the
//                         same content.  Contents are compared in full against the representatives; a key shared
//                         by two different contents (probability ~ n^2 / 2^64) is counted in `collisions` and the
//                         host then takes the exact sort-based path (paths_mark_kernel) instead.
//   scan kernels          exclusive prefix sums of keep and keep*len (window count and nnz), three phases
//   paths_emit_kernel     rowptr / gene / label of the kept rows in input order (group 0 first, as the
//                         reference's `for label, pathSet in enumerate(pathSetList)`), freq[label][gene] += 1
//   paths_code_kernel     0 more good paths / 1 more poor / 2 tie / -1 gene in no path (G2Vec.py:299-307)
namespace g2v {

struct PathSlot {
    unsigned long long key;
    uint32_t idx[2];
};
constexpr unsigned long long kSlotEmpty = ~0ull;
constexpr uint32_t kNoRow = 0xffffffffu;

__global__ void __launch_bounds__(256)
paths_insert_kernel(const unsigned long long *__restrict__ key, const uint8_t *__restrict__ group, int64_t n,
                    PathSlot *__restrict__ table, uint64_t mask) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = key[i];
        const int g = group ? (int)group[i] : 0;
        uint64_t h = mix64(k) & mask;
        while (true) {
            const unsigned long long prev = atomicCAS(&table[h].key, kSlotEmpty, k);
            if (prev == kSlotEmpty || prev == k) {
                atomicMin(&table[h].idx[g], (uint32_t)i);
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

__global__ void __launch_bounds__(kPathWarps * 32)
paths_flag_kernel(const int32_t *__restrict__ rows, const unsigned long long *__restrict__ key,
                  const uint8_t *__restrict__ group, int64_t n, int32_t L, const PathSlot *__restrict__ table,
                  uint64_t mask, uint8_t *__restrict__ keep, unsigned long long *__restrict__ collisions) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kPathWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kPathWarps;
    for (int64_t i = warp; i < n; i += nwarps) {
        const unsigned long long k = key[i];
        const int g = group ? (int)group[i] : 0;
        uint64_t h = mix64(k) & mask;
        while (table[h].key != k) h = (h + 1) & mask;            // the key was inserted: the probe terminates
        const uint32_t rep = table[h].idx[g], other = table[h].idx[1 - g];
        const int32_t *mine = rows + (size_t)i * L;
        bool clash = false, common = false;
        if (rep != (uint32_t)i) clash = !rows_equal(mine, rows + (size_t)rep * L, L, lane);
        if (other != kNoRow) {
            common = rows_equal(mine, rows + (size_t)other * L, L, lane);
            clash = clash || !common;
        }
        if (lane == 0) {
            keep[i] = (rep == (uint32_t)i && other == kNoRow) ? 1 : 0;
            if (clash) atomicAdd(collisions, 1ull);
        }
    }
}

constexpr int kScanThreads = 256, kScanItems = 2, kScanTile = kScanThreads * kScanItems;   // small tiles: enough CTAs to fill the chip at 200k rows

// phase A: per-tile totals of (keep, keep*len)
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const uint8_t *__restrict__ keep, const int32_t *__restrict__ len, int64_t n,
                  unsigned long long *__restrict__ tile_cnt, unsigned long long *__restrict__ tile_len) {
    __shared__ unsigned long long sc[kScanThreads / 32], sl[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    unsigned long long c = 0, l = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < n && keep[i]) { c += 1; l += (unsigned long long)len[i]; }
    }
    c = warp_sum_u64(c); l = warp_sum_u64(l);
    if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = c; sl[threadIdx.x >> 5] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tc = 0, tl = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) { tc += sc[w]; tl += sl[w]; }
        tile_cnt[blockIdx.x] = tc; tile_len[blockIdx.x] = tl;
    }
}

// phase B: exclusive scan of the tile totals by one block (sequential over chunks of its size), totals out
__global__ void __launch_bounds__(1024)
scan_spine_kernel(unsigned long long *__restrict__ tile_cnt, unsigned long long *__restrict__ tile_len, int64_t n_tiles,
                  long long *__restrict__ totals) {
    __shared__ unsigned long long wc[32], wl[32];
    __shared__ unsigned long long carry_c, carry_l;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_c = 0; carry_l = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const unsigned long long c = i < n_tiles ? tile_cnt[i] : 0, l = i < n_tiles ? tile_len[i] : 0;
        unsigned long long ic = c, il = l;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long uc = __shfl_up_sync(0xffffffffu, ic, o), ul = __shfl_up_sync(0xffffffffu, il, o);
            if (lane >= o) { ic += uc; il += ul; }
        }
        if (lane == 31) { wc[warp] = ic; wl[warp] = il; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long xc = wc[lane], xl = wl[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long uc = __shfl_up_sync(0xffffffffu, xc, o), ul = __shfl_up_sync(0xffffffffu, xl, o);
                if (lane >= o) { xc += uc; xl += ul; }
            }
            wc[lane] = xc; wl[lane] = xl;                          // inclusive over warps
        }
        __syncthreads();
        const unsigned long long oc = carry_c + (warp ? wc[warp - 1] : 0) + ic - c;
        const unsigned long long ol = carry_l + (warp ? wl[warp - 1] : 0) + il - l;
        if (i < n_tiles) { tile_cnt[i] = oc; tile_len[i] = ol; }
        __syncthreads();
        if (threadIdx.x == 0) { carry_c += wc[31]; carry_l += wl[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = (long long)carry_c; totals[1] = (long long)carry_l; }
}

// phase C + emit: every kept row gets its window index and its offset in gene[] and is written out by one warp
__global__ void __launch_bounds__(kScanThreads)
paths_emit_kernel(const int32_t *__restrict__ rows, const uint8_t *__restrict__ group, const int32_t *__restrict__ len,
                  const uint8_t *__restrict__ keep, int64_t n, int32_t L, const unsigned long long *__restrict__ tile_cnt,
                  const unsigned long long *__restrict__ tile_len, int32_t *__restrict__ rowptr,
                  int32_t *__restrict__ gene, uint8_t *__restrict__ label, int32_t *__restrict__ freq, int32_t V) {
    __shared__ unsigned long long pc[kScanTile], pl[kScanTile];   // exclusive prefixes inside the tile
    __shared__ unsigned long long sc[kScanThreads], sl[kScanThreads];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    // thread t owns the kScanItems CONSECUTIVE items base + t*kScanItems ..., so that a serial pass per thread plus
    // one block scan of the thread totals gives the prefixes in input order
    unsigned long long c = 0, l = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        pc[threadIdx.x * kScanItems + k] = c; pl[threadIdx.x * kScanItems + k] = l;
        if (i < n && keep[i]) { c += 1; l += (unsigned long long)len[i]; }
    }
    sc[threadIdx.x] = c; sl[threadIdx.x] = l;
    __syncthreads();
    if (threadIdx.x == 0) {                                       // 256 values: a serial exclusive scan is enough
        unsigned long long ac = tile_cnt[blockIdx.x], al = tile_len[blockIdx.x];
        for (int t = 0; t < kScanThreads; ++t) {
            const unsigned long long xc = sc[t], xl = sl[t];
            sc[t] = ac; sl[t] = al; ac += xc; al += xl;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int item = warp; item < kScanTile; item += kScanThreads / 32) {
        const int64_t i = base + item;
        if (i >= n || !keep[i]) continue;                         // warp-uniform
        const int t = item / kScanItems;
        const unsigned long long w = sc[t] + pc[item], off = sl[t] + pl[item];
        const int32_t ln = len[i];
        const int lab = group ? (int)group[i] : 0;
        if (lane == 0) { rowptr[w] = (int32_t)off; label[w] = (uint8_t)lab; }
        for (int k = lane; k < ln; k += 32) {
            const int32_t g = __ldg(rows + (size_t)i * L + k);
            gene[off + k] = g;
            if (freq && g >= 0 && g < V) atomicAdd(freq + (size_t)lab * V + g, 1);
        }
    }
}

__global__ void paths_code_kernel(const int32_t *__restrict__ freq, int32_t V, int8_t *__restrict__ code) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= V) return;
    const int32_t fg = freq[g], fp = freq[V + g];
    code[g] = (fg + fp == 0) ? -1 : (fg > fp ? 0 : (fg < fp ? 1 : 2));
}

static uint64_t table_slots(int64_t n) {
    uint64_t s = 1024;
    while (s < 2ull * (uint64_t)n) s <<= 1;
    return s;
}

}  // namespace g2v

extern "C" size_t g2v_paths_set_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    const size_t tiles = (size_t)((n + kScanTile - 1) / kScanTile);
    return table_slots(n) * sizeof(PathSlot) + 2 * tiles * sizeof(unsigned long long) + 256;
}

// Phase 1: keep[] (+ totals[0] = kept rows, totals[1] = their total length, totals[2] = key collisions; device).
extern "C" int g2v_paths_set_select(const int32_t *rows, const int64_t *key, const uint8_t *group, const int32_t *len,
                                    int64_t n, int32_t L, void *workspace, uint8_t *keep, int64_t *totals, void *stream) {
    G2V_REQUIRE(n >= 0 && n < 0xffffffffll && L >= 1, "g2v_paths_set_select: bad sizes");
    G2V_REQUIRE(totals, "g2v_paths_set_select: null totals");
    cudaStream_t st = (cudaStream_t)stream;
    G2V_CUDA_OK(cudaMemsetAsync(totals, 0, 3 * sizeof(int64_t), st));
    if (n == 0) return 0;
    G2V_REQUIRE(rows && key && len && workspace && keep, "g2v_paths_set_select: null pointer");
    DeviceProps dp;
    if (device_props(&dp)) return 1;
    const uint64_t slots = table_slots(n);
    PathSlot *table = reinterpret_cast<PathSlot *>(workspace);
    unsigned long long *tile_cnt = reinterpret_cast<unsigned long long *>(table + slots);
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    unsigned long long *tile_len = tile_cnt + tiles;
    G2V_CUDA_OK(cudaMemsetAsync(table, 0xff, slots * sizeof(PathSlot), st));
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)dp.sm_count * 16;
    paths_insert_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, st>>>(
        reinterpret_cast<const unsigned long long *>(key), group, n, table, slots - 1);
    G2V_CUDA_OK(cudaGetLastError());
    blocks = (n + kPathWarps - 1) / kPathWarps;
    paths_flag_kernel<<<(unsigned)(blocks > cap ? cap : blocks), kPathWarps * 32, 0, st>>>(
        rows, reinterpret_cast<const unsigned long long *>(key), group, n, L, table, slots - 1, keep,
        reinterpret_cast<unsigned long long *>(totals + 2));
    G2V_CUDA_OK(cudaGetLastError());
    scan_tiles_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(keep, len, n, tile_cnt, tile_len);
    G2V_CUDA_OK(cudaGetLastError());
    scan_spine_kernel<<<1, 1024, 0, st>>>(tile_cnt, tile_len, tiles, reinterpret_cast<long long *>(totals));
    G2V_CUDA_OK(cudaGetLastError());
    count_launch(4);
    return 0;
}

// Phase 2 (after the host has sized the outputs from totals): CSR windows of the kept rows + gene-frequency codes.
// rowptr [kept+1], gene [total length], label [kept], freq [2*V] scratch, code [V] (nullable pair).
extern "C" int g2v_paths_set_emit(const int32_t *rows, const uint8_t *group, const int32_t *len, const uint8_t *keep,
                                  int64_t n, int32_t L, int32_t V, const void *workspace, int64_t kept, int64_t nnz,
                                  int32_t *rowptr, int32_t *gene, uint8_t *label, int32_t *freq, int8_t *code,
                                  void *stream) {
    G2V_REQUIRE(n >= 0 && L >= 1 && V > 0 && kept >= 0 && nnz >= 0 && nnz < (1ll << 31), "g2v_paths_set_emit: bad sizes");
    G2V_REQUIRE(rowptr, "g2v_paths_set_emit: null rowptr");
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t last = (int32_t)nnz;
    G2V_CUDA_OK(cudaMemcpyAsync(rowptr + kept, &last, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (freq) G2V_CUDA_OK(cudaMemsetAsync(freq, 0, sizeof(int32_t) * 2 * (size_t)V, st));
    if (n > 0 && kept > 0) {
        G2V_REQUIRE(rows && len && keep && workspace && (nnz == 0 || gene) && label, "g2v_paths_set_emit: null pointer");
        const uint64_t slots = table_slots(n);
        const unsigned long long *tile_cnt =
            reinterpret_cast<const unsigned long long *>(reinterpret_cast<const PathSlot *>(workspace) + slots);
        const int64_t tiles = (n + kScanTile - 1) / kScanTile;
        paths_emit_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(rows, group, len, keep, n, L, tile_cnt, tile_cnt + tiles,
                                                                     rowptr, gene, label, freq, V);
        G2V_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    if (freq && code) {
        paths_code_kernel<<<(V + 255) / 256, 256, 0, st>>>(freq, V, code);
        G2V_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    G2V_CUDA_OK(cudaStreamSynchronize(st));                       // `last` is on this call's stack
    return 0;
}
