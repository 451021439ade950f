// g2v_walk.cu -- HOT PATH 1: self-avoiding weighted random walks on CSR, one warp per walker.
//
// Replaces generate_pathSet / generate_randomPath (/root/reference/G2Vec.py:324-352).
// Per step of a walker at node `cur` (G2Vec.py:331-344):
//     path.append(cur)                                    -> path[] in shared memory
//     prob = adjMat[cur]; prob[path] = 0                  -> CSR row, visited test per neighbour
//     if prob.sum() > 0: cur = choice(p = prob / sum)     -> integer inverse CDF, one Philox draw
//     else: break
//
// Layout: the group's graph as CSR in HBM (rowptr int32 [V+1], col int32 [E] ascending per
// row, qw uint32 [E]); both are read with coalesced 128 B warp loads (32 neighbours per
// request).  Per warp in shared memory: the path (L ints, written back once, coalesced) and
// the visited set -- a V-bit bitmap (one LDS per neighbour) while 8 warps' bitmaps fit in 56 KB
// (V <= ~46k at L = 80), otherwise an open-addressing hash set of >= 3L slots whose size is
// independent of V (200k-node graphs keep full occupancy).  Neighbour chunks are kept in registers between the two passes
// (total, then selection); rows longer than 32*KC neighbours re-read the tail (L1/L2 hits).
// Walkers are handed out by an atomic ticket so that warps whose walker dead-ends early
// (62 % of ex_* start nodes have no out-edge) immediately take the next one.
//
// Integer arithmetic only on the selection path => bit-exact against oracle/g2v_oracle.c for
// any scan order:  T = sum of unvisited qw (uint64), r = mulhi64(x, T), first inclusive
// prefix > r.
#include <stdlib.h>

#include "g2v_common.cuh"

namespace g2v {

constexpr int kWalkWarps = 8;   // warps per CTA
constexpr int kKC = 4;          // neighbour chunks (of 32) cached in registers

__device__ __forceinline__ uint32_t hash_slot(int32_t c, int shift) {
    return ((uint32_t)c * 2654435761u) >> shift;
}

// true if node c is in the warp's visited set.  BITMAP: hs is a V-bit bitmap (1 LDS);
// otherwise an open-addressing hash set of node ids (-1 = empty).
template <bool BITMAP>
__device__ __forceinline__ bool visited(const int32_t *__restrict__ hs, uint32_t mask, int shift,
                                        int32_t c) {
    if (BITMAP) return (hs[c >> 5] >> (c & 31)) & 1;
    uint32_t i = hash_slot(c, shift);
    while (true) {
        int32_t x = hs[i];
        if (x == c) return true;
        if (x < 0) return false;
        i = (i + 1) & mask;
    }
}

template <bool BITMAP>
__global__ void __launch_bounds__(kWalkWarps * 32)
walk_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
            const uint32_t *__restrict__ qw, int32_t V, int32_t L, int32_t Lpad, int32_t H,
            int32_t hshift, uint64_t seed, uint32_t group, int64_t walker_begin,
            int64_t n_walkers, int64_t walker_stride, int32_t *__restrict__ out_nodes,
            int32_t *__restrict__ out_len, unsigned long long *__restrict__ ticket) {
    extern __shared__ int32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *path = smem + (size_t)warp * (Lpad + H);
    int32_t *hs = path + Lpad;
    const uint32_t hmask = (uint32_t)H - 1u;

    for (int i = lane; i < H; i += 32) hs[i] = BITMAP ? 0 : -1;
    __syncwarp();

    while (true) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((int64_t)t >= n_walkers) break;
        const int64_t w = walker_begin + (int64_t)t * walker_stride;
        const uint64_t subseq = ((uint64_t)group << 40) + (uint64_t)w;
        int32_t cur = (int32_t)(w % V);
        int32_t n = 0;
        bool dirty = false;
        uint32_t dlo = 0, dhi = 0;                       // lane k holds the draw of step (s & ~31) + k
        int32_t dbase = -1;

        for (int32_t s = 0; s < L; ++s) {
            path[n] = cur;                               // every lane stores the same value: no divergence
            ++n;
            if (s == L - 1) break;                       // last draw is never appended
            const int32_t b = __ldg(rowptr + cur), e = __ldg(rowptr + cur + 1);
            if (b == e) break;                           // no out-edges: dead end
            if (BITMAP) {                                // visited.insert(cur), uniform across the warp
                const int32_t wv = hs[cur >> 5];
                __syncwarp();
                hs[cur >> 5] = wv | (1 << (cur & 31));
            } else {
                uint32_t i = hash_slot(cur, hshift);
                while (hs[i] >= 0) i = (i + 1) & hmask;
                __syncwarp();
                hs[i] = cur;
            }
            dirty = true;
            __syncwarp();

            // ---- pass 1: weight of the unvisited out-neighbours, per chunk of 32 (REDUX.SUM)
            uint32_t mq[kKC], tot[kKC];
            int32_t mc[kKC];
            unsigned long long T = 0;
#pragma unroll
            for (int k = 0; k < kKC; ++k) {
                mq[k] = 0; mc[k] = -1; tot[k] = 0;
                if (b + k * 32 < e) {                    // warp-uniform
                    const int32_t j = b + k * 32 + lane;
                    if (j < e) {
                        const int32_t c = __ldg(col + j);
                        const uint32_t q = __ldg(qw + j);
                        mc[k] = c;
                        mq[k] = visited<BITMAP>(hs, hmask, hshift, c) ? 0u : q;
                    }
                    tot[k] = __reduce_add_sync(0xffffffffu, mq[k]);   // <= 32 * 2^24
                    T += tot[k];
                }
            }
            for (int32_t jb = b + kKC * 32; jb < e; jb += 32) {       // rows longer than 128 neighbours
                const int32_t j = jb + lane;
                uint32_t q = 0;
                if (j < e) q = visited<BITMAP>(hs, hmask, hshift, __ldg(col + j)) ? 0u : __ldg(qw + j);
                T += __reduce_add_sync(0xffffffffu, q);
            }
            if (T == 0) break;                           // every neighbour already visited

            // ---- one 64-bit Philox draw per step, r uniform in [0, T); 32 steps are drawn at once,
            //      one per lane (counter-based: lane k evaluates step dbase + k)
            if ((s & ~31) != dbase) {
                dbase = s & ~31;
                const uint64_t d = draw64(seed, subseq, (uint32_t)(dbase + lane));
                dlo = (uint32_t)d; dhi = (uint32_t)(d >> 32);
            }
            const uint64_t x = ((uint64_t)__shfl_sync(0xffffffffu, dhi, s & 31) << 32) |
                               __shfl_sync(0xffffffffu, dlo, s & 31);
            unsigned long long rem = __umul64hi(x, T);   // r - (weight of the chunks already skipped)

            // ---- pass 2: chunk that contains r (uniform scalar search), then one warp scan inside it
            int32_t nxt = -1;
            bool found = false;
#pragma unroll
            for (int k = 0; k < kKC; ++k) {
                if (!found && b + k * 32 < e) {
                    if (rem < (unsigned long long)tot[k]) {
                        const uint32_t incl = warp_inclusive_scan_u32(mq[k], lane);
                        const unsigned hit = __ballot_sync(0xffffffffu, incl > (uint32_t)rem);
                        nxt = __shfl_sync(0xffffffffu, mc[k], __ffs(hit) - 1);
                        found = true;
                    } else {
                        rem -= tot[k];
                    }
                }
            }
            for (int32_t jb = b + kKC * 32; !found && jb < e; jb += 32) {
                const int32_t j = jb + lane;
                int32_t c = -1;
                uint32_t q = 0;
                if (j < e) {
                    c = __ldg(col + j);
                    q = visited<BITMAP>(hs, hmask, hshift, c) ? 0u : __ldg(qw + j);
                }
                const uint32_t ct = __reduce_add_sync(0xffffffffu, q);
                if (rem < (unsigned long long)ct) {
                    const uint32_t incl = warp_inclusive_scan_u32(q, lane);
                    const unsigned hit = __ballot_sync(0xffffffffu, incl > (uint32_t)rem);
                    nxt = __shfl_sync(0xffffffffu, c, __ffs(hit) - 1);
                    found = true;
                } else {
                    rem -= ct;
                }
            }
            cur = nxt;
        }

        __syncwarp();
        // ---- write the row once, coalesced; -1 padding
        int32_t *row = out_nodes + (size_t)t * (size_t)L;
        for (int i = lane; i < L; i += 32) row[i] = (i < n) ? path[i] : -1;
        if (lane == 0) out_len[t] = n;
        if (dirty) {
            __syncwarp();
            if (BITMAP) {
                for (int i = lane; i < n; i += 32) hs[path[i] >> 5] = 0;   // only the words this walk touched
            } else {
                for (int i = lane; i < H; i += 32) hs[i] = -1;
            }
        }
        __syncwarp();
    }
}

__global__ void test_draws_kernel(uint64_t seed, uint64_t subseq, int32_t n, uint64_t *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = draw64(seed, subseq, (uint32_t)i);
}

}  // namespace g2v

using namespace g2v;

extern "C" size_t g2v_walk_workspace_bytes(void) { return 256; }

extern "C" int g2v_walk_launch(const int32_t *rowptr, const int32_t *col, const uint32_t *qw,
                               int32_t V, int64_t E, int32_t L, uint64_t seed, uint32_t group,
                               int64_t walker_begin, int64_t walker_end, int64_t walker_stride,
                               int32_t *out_nodes, int32_t *out_len, void *workspace,
                               void *stream) {
    G2V_REQUIRE(V > 0 && E >= 0, "g2v_walk_launch: V must be > 0 and E >= 0 (V=%d E=%lld)", V, (long long)E);
    G2V_REQUIRE(L >= 1 && L <= 4096, "g2v_walk_launch: lenPath must be in [1, 4096] (got %d)", L);
    G2V_REQUIRE(walker_stride >= 1 && walker_begin >= 0, "g2v_walk_launch: bad walker range");
    G2V_REQUIRE(rowptr && out_nodes && out_len && workspace, "g2v_walk_launch: null pointer");
    G2V_REQUIRE(E == 0 || (col && qw), "g2v_walk_launch: null col/qw with E > 0");
    const int64_t n_walkers =
        walker_end > walker_begin ? (walker_end - walker_begin + walker_stride - 1) / walker_stride : 0;
    if (n_walkers == 0) return 0;
    DeviceProps dp;
    if (device_props(&dp)) return 1;
    G2V_REQUIRE(dp.cc_major == 10, "g2v_walk_launch: needs an sm_100 device (found sm_%d%d)", dp.cc_major, dp.cc_minor);

    // visited set per warp: a V-bit bitmap when it is small enough to keep >= 4 CTAs per SM, else a hash set
    const int Lpad = (L + 31) & ~31;
    const int bm_words = (V + 31) / 32;
    const char *force = getenv("G2V_WALK_VISITED");             // test hook: "hash" / "bitmap"
    bool bitmap = (size_t)kWalkWarps * (Lpad + bm_words) * sizeof(int32_t) <= 56 * 1024;
    if (force && force[0] == 'h') bitmap = false;
    int H = 64, hshift = 26;
    if (bitmap) {
        H = bm_words;
    } else {
        while (H < 3 * L) { H <<= 1; --hshift; }
    }
    const size_t smem = (size_t)kWalkWarps * (Lpad + H) * sizeof(int32_t);
    G2V_REQUIRE(smem <= (size_t)dp.max_smem_optin, "g2v_walk_launch: lenPath %d needs %zu B of shared memory", L, smem);
    cudaStream_t st = (cudaStream_t)stream;
    auto kern = bitmap ? walk_kernel<true> : walk_kernel<false>;
    static bool attr_done[2] = {false, false};
    if (!attr_done[bitmap]) {
        G2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dp.max_smem_optin));
        G2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_done[bitmap] = true;
    }
    int per_sm = 0;
    G2V_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWalkWarps * 32, smem));
    G2V_REQUIRE(per_sm > 0, "g2v_walk_launch: kernel does not fit on an SM");
    int64_t grid = (int64_t)dp.sm_count * per_sm;                 // persistent: whole chip resident
    const int64_t need = (n_walkers + kWalkWarps - 1) / kWalkWarps;
    if (grid > need) grid = need;
    G2V_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(unsigned long long), st));
    kern<<<(unsigned)grid, kWalkWarps * 32, smem, st>>>(
        rowptr, col, qw, V, L, Lpad, H, hshift, seed, group, walker_begin, n_walkers, walker_stride,
        out_nodes, out_len, (unsigned long long *)workspace);
    G2V_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int g2v_walk_host(const int32_t *rowptr, const int32_t *col, const uint32_t *qw,
                             int32_t V, int64_t E, int32_t L, uint64_t seed, uint32_t group,
                             int64_t walker_begin, int64_t walker_end, int64_t walker_stride,
                             int32_t *out_nodes, int32_t *out_len) {
    G2V_REQUIRE(V > 0 && E >= 0 && L >= 1 && walker_stride >= 1, "g2v_walk_host: bad arguments");
    const int64_t n = walker_end > walker_begin ? (walker_end - walker_begin + walker_stride - 1) / walker_stride : 0;
    if (n == 0) return 0;
    int32_t *d_rowptr = nullptr, *d_col = nullptr, *d_nodes = nullptr, *d_len = nullptr;
    uint32_t *d_qw = nullptr;
    void *d_ws = nullptr;
    int rc = 1;
    cudaStream_t st = nullptr;
    do {
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) break;
        if (cudaMalloc(&d_rowptr, sizeof(int32_t) * (size_t)(V + 1)) != cudaSuccess) break;
        if (cudaMalloc(&d_col, sizeof(int32_t) * (size_t)(E > 0 ? E : 1)) != cudaSuccess) break;
        if (cudaMalloc(&d_qw, sizeof(uint32_t) * (size_t)(E > 0 ? E : 1)) != cudaSuccess) break;
        if (cudaMalloc(&d_nodes, sizeof(int32_t) * (size_t)n * (size_t)L) != cudaSuccess) break;
        if (cudaMalloc(&d_len, sizeof(int32_t) * (size_t)n) != cudaSuccess) break;
        if (cudaMalloc(&d_ws, g2v_walk_workspace_bytes()) != cudaSuccess) break;
        if (cudaMemcpyAsync(d_rowptr, rowptr, sizeof(int32_t) * (size_t)(V + 1), cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        if (E > 0) {
            if (cudaMemcpyAsync(d_col, col, sizeof(int32_t) * (size_t)E, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
            if (cudaMemcpyAsync(d_qw, qw, sizeof(uint32_t) * (size_t)E, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        }
        rc = g2v_walk_launch(d_rowptr, d_col, d_qw, V, E, L, seed, group, walker_begin, walker_end,
                             walker_stride, d_nodes, d_len, d_ws, st);
        if (rc) break;
        rc = 1;
        if (cudaMemcpyAsync(out_nodes, d_nodes, sizeof(int32_t) * (size_t)n * (size_t)L, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaMemcpyAsync(out_len, d_len, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc == 1) {
        cudaError_t e = cudaGetLastError();
        set_error("g2v_walk_host: CUDA failure: %s", cudaGetErrorString(e));
    }
    cudaFree(d_rowptr); cudaFree(d_col); cudaFree(d_qw); cudaFree(d_nodes); cudaFree(d_len); cudaFree(d_ws);
    if (st) cudaStreamDestroy(st);
    return rc;
}

extern "C" int g2v_test_draws(uint64_t seed, uint64_t subsequence, int32_t n, uint64_t *out_dev,
                              void *stream) {
    G2V_REQUIRE(n >= 0 && out_dev, "g2v_test_draws: bad arguments");
    if (n == 0) return 0;
    test_draws_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, subsequence, n, out_dev);
    G2V_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
