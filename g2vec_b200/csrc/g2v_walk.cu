// g2v_walk.cu -- HOT PATH 1: self-avoiding weighted random walks on CSR, one warp per walker.
//
// Replaces generate_pathSet / generate_randomPath (/root/reference/G2Vec.py:324-352).
// Per step of a walker at node `cur` (G2Vec.py:331-344):
//     path.append(cur)                                    -> path[] in shared memory
//     prob = adjMat[cur]; prob[path] = 0                  -> CSR row + the walker's visited set
//     if prob.sum() > 0: cur = choice(p = prob / sum)     -> rejection sampling over the row's static prefix
//     else: break                                            sums, exact masked inverse CDF as the fallback
//
// Layout: the group's graph as CSR in HBM (rowptr int32 [V+1], col int32 [E] ascending per row, psum uint32
// [E] = inclusive prefix sums of the quantised weights inside each row).  A step reads the row's prefix sums
// with coalesced 128 B warp loads, ONE col entry per attempt and one word of the visited set -- a V-bit
// bitmap (one LDS) while 8 warps' bitmaps fit 56 KB (V <= ~46k at L = 80), otherwise an open-addressing hash
// set of >= 3L slots whose size is independent of V.  The path lives in shared memory and is written back
// once, coalesced.  Philox draws of the first attempts are evaluated 32 steps at a time, one step per lane.
// Walkers are handed out by an atomic ticket so that warps whose walker dead-ends early (62 % of ex_* start
// nodes have no out-edge) immediately take the next one.
//
// Integer arithmetic only => bit-exact against oracle/g2v_oracle.c for any evaluation order.
#include <stdlib.h>

#include "g2v_common.cuh"

namespace g2v {

constexpr int kWalkWarps = 8;   // warps per CTA

__device__ __forceinline__ uint32_t hash_slot(int32_t c, int shift) {
    return ((uint32_t)c * 2654435761u) >> shift;
}

// Is node c in the walker's visited set?  `hs` indexes the dynamic shared array (an integer offset, so that
// every access is a plain LDS/STS with a register offset).
extern __shared__ int32_t g2v_walk_smem[];
template <bool BITMAP>
__device__ __forceinline__ bool is_visited(int hs, uint32_t mask, int shift, int32_t c) {
    if (BITMAP) return ((uint32_t)g2v_walk_smem[hs + (c >> 5)] >> (c & 31)) & 1u;
    uint32_t i = hash_slot(c, shift);
    while (true) {
        const int32_t x = g2v_walk_smem[hs + i];
        if (x == c) return true;
        if (x < 0) return false;
        i = (i + 1) & mask;
    }
}

constexpr int kAttempts = 4;          // rejection attempts per step before the exact fallback
constexpr uint32_t kMaxLen = 4096;    // draw index of (step s, attempt a) = a * kMaxLen + s

// the rare draws (attempts after the first, the fallback) stay out of line: less register pressure in the loop
__device__ __noinline__ uint64_t draw64_rare(uint64_t seed, uint64_t subseq, uint32_t k) { return draw64(seed, subseq, k); }

// r = floor(x * T / 2^64) for T < 2^32: two 32x32 multiplies instead of a 64x64 high multiply
__device__ __forceinline__ uint32_t mulhi64_32(uint64_t x, uint32_t T) {
    const unsigned long long lo = (unsigned long long)(uint32_t)x * T;
    return (uint32_t)(((unsigned long long)(uint32_t)(x >> 32) * T + (lo >> 32)) >> 32);
}

// One warp per walker.  Per step (G2Vec.py:331-344), with `psum` the inclusive prefix sums of the row's
// quantised weights (static, built once on the host):
//   T_all = psum[e-1];  up to kAttempts times:  r = floor(x*T_all/2^64), candidate = first neighbour whose
//   prefix exceeds r (ballot over the register-cached prefix chunks), accepted if not yet visited (ONE visited
//   test and ONE col load per attempt instead of a scan of the whole row);  if every attempt hit a visited
//   node: exact inverse CDF over the unvisited neighbours (dead end if none).  KC = prefix chunks of 32 kept
//   in registers (2 -> 32 registers, 8 CTAs per SM; 4 for graphs with longer rows).
template <bool BITMAP, int KC>
__global__ void __launch_bounds__(kWalkWarps * 32, KC == 2 ? 8 : 6)
walk_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
            const uint32_t *__restrict__ psum, int32_t V, int32_t L, int32_t Lpad, int32_t H,
            int32_t hshift, uint64_t seed, uint32_t group, int64_t walker_begin,
            int64_t n_walkers, int64_t walker_stride, int32_t *__restrict__ out_nodes,
            int32_t *__restrict__ out_len, unsigned long long *__restrict__ ticket) {
    int32_t *const smem = g2v_walk_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int path = warp * (Lpad + H);                 // offsets into smem (ints), not pointers
    const int hs = path + Lpad;
    const uint32_t hmask = (uint32_t)H - 1u;

    for (int i = lane; i < H; i += 32) smem[hs + i] = BITMAP ? 0 : -1;
    __syncwarp();

    while (true) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((int64_t)t >= n_walkers) break;
        const int64_t w = walker_begin + (int64_t)t * walker_stride;
        const uint64_t subseq = ((uint64_t)group << 40) + (uint64_t)w;
        int32_t cur = (int32_t)(w % V);
        int32_t n = 0, dbase = -1;
        bool dirty = false;
        uint32_t dlo = 0, dhi = 0;                       // lane k holds first-attempt draw of step dbase + k

        for (int32_t s = 0; s < L; ++s) {
            smem[path + n] = cur;                        // every lane stores the same value
            ++n;
            if (s == L - 1) break;                       // the L-th node is appended, never expanded
            const int32_t b = __ldg(rowptr + cur), e = __ldg(rowptr + cur + 1);
            if (b == e) break;                           // no out-edges: dead end
            if (BITMAP) {                                // visited.insert(cur), uniform across the warp
                const int32_t wv = smem[hs + (cur >> 5)];
                __syncwarp();
                smem[hs + (cur >> 5)] = wv | (1 << (cur & 31));
            } else {
                uint32_t i = hash_slot(cur, hshift);
                while (smem[hs + i] >= 0) i = (i + 1) & hmask;
                __syncwarp();
                smem[hs + i] = cur;
            }
            dirty = true;
            __syncwarp();

            // the row's prefix sums: KC chunks in registers (lanes past the end hold UINT32_MAX, which can
            // never be the first prefix > r because r < T_all), T_all by a broadcast load
            uint32_t pk[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int32_t j = b + k * 32 + lane;
                pk[k] = (j < e) ? __ldg(psum + j) : 0xffffffffu;
            }
            const uint32_t Tall = __ldg(psum + e - 1);

            int32_t nxt = -1;
            for (int a = 0; a < kAttempts && nxt < 0; ++a) {
                uint64_t x;
                if (a == 0) {                            // first attempts of 32 consecutive steps: one Philox
                    if ((s & ~31) != dbase) {            // evaluation, one step per lane
                        dbase = s & ~31;
                        const uint64_t d = draw64(seed, subseq, (uint32_t)(dbase + lane));
                        dlo = (uint32_t)d; dhi = (uint32_t)(d >> 32);
                    }
                    x = ((uint64_t)__shfl_sync(0xffffffffu, dhi, s & 31) << 32) | __shfl_sync(0xffffffffu, dlo, s & 31);
                } else {
                    x = draw64_rare(seed, subseq, (uint32_t)a * kMaxLen + (uint32_t)s);
                }
                const uint32_t r = mulhi64_32(x, Tall);
                int32_t jstar = -1;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    if (jstar < 0 && (k == 0 || b + k * 32 < e)) {
                        const unsigned hit = __ballot_sync(0xffffffffu, pk[k] > r);
                        if (hit) jstar = b + k * 32 + __ffs(hit) - 1;
                    }
                }
                for (int32_t jb = b + KC * 32; jstar < 0 && jb < e; jb += 32) {   // rows longer than KC*32
                    const int32_t j = jb + lane;
                    const uint32_t p = (j < e) ? __ldg(psum + j) : 0xffffffffu;
                    const unsigned hit = __ballot_sync(0xffffffffu, p > r);
                    if (hit) jstar = jb + __ffs(hit) - 1;
                }
                const int32_t c = __ldg(col + jstar);    // same address in every lane: one broadcast load
                if (!is_visited<BITMAP>(hs, hmask, hshift, c)) nxt = c;
            }

            if (nxt < 0) {
                // ---- exact fallback: inverse CDF over the unvisited neighbours (weights = prefix differences)
                unsigned long long T = 0;
                for (int32_t jb = b; jb < e; jb += 32) {
                    const int32_t j = jb + lane;
                    uint32_t q = 0;
                    if (j < e && !is_visited<BITMAP>(hs, hmask, hshift, __ldg(col + j)))
                        q = __ldg(psum + j) - (j > b ? __ldg(psum + j - 1) : 0u);
                    T += __reduce_add_sync(0xffffffffu, q);
                }
                if (T == 0) break;                       // every neighbour already visited: dead end
                uint32_t rem = mulhi64_32(draw64_rare(seed, subseq, (uint32_t)kAttempts * kMaxLen + (uint32_t)s), (uint32_t)T);
                for (int32_t jb = b; nxt < 0 && jb < e; jb += 32) {
                    const int32_t j = jb + lane;
                    int32_t c = -1;
                    uint32_t q = 0;
                    if (j < e) {
                        c = __ldg(col + j);
                        if (!is_visited<BITMAP>(hs, hmask, hshift, c))
                            q = __ldg(psum + j) - (j > b ? __ldg(psum + j - 1) : 0u);
                    }
                    const uint32_t ct = __reduce_add_sync(0xffffffffu, q);
                    if (rem < ct) {
                        const uint32_t incl = warp_inclusive_scan_u32(q, lane);
                        const unsigned hit = __ballot_sync(0xffffffffu, incl > rem);
                        nxt = __shfl_sync(0xffffffffu, c, __ffs(hit) - 1);
                    } else {
                        rem -= ct;
                    }
                }
            }
            cur = nxt;
        }

        __syncwarp();
        // ---- write the row once, coalesced; -1 padding
        int32_t *row = out_nodes + (size_t)t * (size_t)L;
        for (int i = lane; i < L; i += 32) row[i] = (i < n) ? smem[path + i] : -1;
        if (lane == 0) out_len[t] = n;
        if (dirty) {
            __syncwarp();
            if (BITMAP) {
                for (int i = lane; i < n; i += 32) smem[hs + (smem[path + i] >> 5)] = 0;   // only the touched words
            } else {
                for (int i = lane; i < H; i += 32) smem[hs + i] = -1;
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

extern "C" int g2v_walk_launch(const int32_t *rowptr, const int32_t *col, const uint32_t *psum,
                               int32_t V, int64_t E, int32_t L, uint64_t seed, uint32_t group,
                               int64_t walker_begin, int64_t walker_end, int64_t walker_stride,
                               int32_t *out_nodes, int32_t *out_len, void *workspace,
                               void *stream) {
    G2V_REQUIRE(V > 0 && E >= 0, "g2v_walk_launch: V must be > 0 and E >= 0 (V=%d E=%lld)", V, (long long)E);
    G2V_REQUIRE(L >= 1 && L <= 4096, "g2v_walk_launch: lenPath must be in [1, 4096] (got %d)", L);
    G2V_REQUIRE(walker_stride >= 1 && walker_begin >= 0, "g2v_walk_launch: bad walker range");
    const int64_t n_walkers =
        walker_end > walker_begin ? (walker_end - walker_begin + walker_stride - 1) / walker_stride : 0;
    if (n_walkers == 0) return 0;                                 // empty range: nothing to write
    G2V_REQUIRE(rowptr && out_nodes && out_len && workspace, "g2v_walk_launch: null pointer");
    G2V_REQUIRE(E == 0 || (col && psum), "g2v_walk_launch: null col/psum with E > 0");
    DeviceProps dp;
    if (device_props(&dp)) return 1;
    G2V_REQUIRE(dp.cc_major == 10, "g2v_walk_launch: needs an sm_100 device (found sm_%d%d)", dp.cc_major, dp.cc_minor);

    // visited set per walker: a V-bit bitmap when a CTA's bitmaps fit 56 KB (>= 4 CTAs per SM), else a hash set
    const double mean_deg = (double)E / (double)V;
    const int Lpad = (L + 31) & ~31;
    const int bm_words = (V + 31) / 32;
    const char *force = getenv("G2V_WALK_VISITED");               // test hook: "hash" / "bitmap"
    int Hh = 64, hshift = 26;                                     // hash set: >= 3L slots, power of two
    while (Hh < 3 * L) { Hh <<= 1; --hshift; }
    const size_t per_warp = (size_t)kWalkWarps * sizeof(int32_t);
    const size_t bm_smem = per_warp * (Lpad + bm_words), hash_smem = per_warp * (Lpad + Hh);
    bool bitmap = bm_smem <= 56 * 1024 || bm_smem <= hash_smem;   // occupancy first, then whichever is smaller
    if (force && force[0] == 'h') bitmap = false;
    const int H = bitmap ? bm_words : Hh;
    const size_t smem = (size_t)kWalkWarps * (Lpad + H) * sizeof(int32_t);
    G2V_REQUIRE(smem <= (size_t)dp.max_smem_optin, "g2v_walk_launch: lenPath %d needs %zu B of shared memory", L, smem);
    cudaStream_t st = (cudaStream_t)stream;
    typedef void (*kern_t)(const int32_t *, const int32_t *, const uint32_t *, int32_t, int32_t, int32_t, int32_t,
                           int32_t, uint64_t, uint32_t, int64_t, int64_t, int64_t, int32_t *, int32_t *,
                           unsigned long long *);
    // prefix chunks cached in registers: 2 when rows mostly fit 64 neighbours, 4 otherwise (G2V_WALK_KC overrides)
    const char *fk = getenv("G2V_WALK_KC");
    int kc = mean_deg <= 64.0 ? 2 : 4;
    if (fk && (atoi(fk) == 2 || atoi(fk) == 4)) kc = atoi(fk);
    static const kern_t table[2][2] = {{walk_kernel<false, 2>, walk_kernel<false, 4>},
                                       {walk_kernel<true, 2>, walk_kernel<true, 4>}};
    kern_t kern = table[bitmap][kc == 4];
    // per-device function attributes (set on every call: the process may have switched device)
    G2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dp.max_smem_optin));
    G2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    G2V_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWalkWarps * 32, smem));
    G2V_REQUIRE(per_sm > 0, "g2v_walk_launch: kernel does not fit on an SM");
    int64_t grid = (int64_t)dp.sm_count * per_sm;                 // persistent: whole chip resident
    const int64_t need = (n_walkers + kWalkWarps - 1) / kWalkWarps;
    if (grid > need) grid = need;
    G2V_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(unsigned long long), st));
    kern<<<(unsigned)grid, kWalkWarps * 32, smem, st>>>(
        rowptr, col, psum, V, L, Lpad, H, hshift, seed, group, walker_begin, n_walkers, walker_stride,
        out_nodes, out_len, (unsigned long long *)workspace);
    G2V_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int g2v_walk_host(const int32_t *rowptr, const int32_t *col, const uint32_t *psum,
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
            if (cudaMemcpyAsync(d_qw, psum, sizeof(uint32_t) * (size_t)E, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
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
