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
