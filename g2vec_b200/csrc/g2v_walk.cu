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
// independent of V (200k-node graphs keep full occupancy).  KC (2 or 4) neighbour chunks are kept in
// registers between the two passes (per-chunk totals with REDUX.SUM, then one scan inside the selected
// chunk); rows longer than 32*KC neighbours re-read the tail (L1/L2 hits).  Philox draws are evaluated
// 32 steps at a time, one step per lane.
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
#ifndef G2V_WALK_MINB
#define G2V_WALK_MINB 6         // resident CTAs per SM the KC = 2 kernels are compiled for (measured: 8 -> 2.73 ms, 6 -> 2.65 ms)
#endif

__device__ __forceinline__ uint32_t hash_slot(int32_t c, int shift) {
    return ((uint32_t)c * 2654435761u) >> shift;
}

// Returns q if node c is NOT in the visited set, else 0.  `hs` indexes the dynamic shared array (kept as
// an integer offset so that every access is a plain LDS/STS with a register offset).
extern __shared__ int32_t g2v_walk_smem[];
template <bool BITMAP>
__device__ __forceinline__ uint32_t unvisited_weight(int hs, uint32_t mask, int shift, int32_t c, uint32_t q) {
    if (BITMAP) {
        const uint32_t bit = ((uint32_t)g2v_walk_smem[hs + (c >> 5)] >> (c & 31)) & 1u;
        return q & (bit - 1u);                            // bit = 1 -> 0, bit = 0 -> q
    }
    uint32_t i = hash_slot(c, shift);
    while (true) {
        const int32_t x = g2v_walk_smem[hs + i];
        if (x == c) return 0u;
        if (x < 0) return q;
        i = (i + 1) & mask;
    }
}

// One TILE of lanes (8, 16 or 32) per walker, 32/TILE walkers per warp.  The loop is a flat state
// machine -- every iteration is "one step for every tile of the warp" -- so that tiles whose walkers end
// at different times stay converged: finishing a walk (row write-out, visited-set reset) and fetching
// the next ticket are short predicated sections of the same iteration.
// KC = neighbour chunks (of TILE) kept in registers between the two passes: 2 for graphs whose rows
// mostly fit 64 neighbours (fewer registers -> 8 resident CTAs per SM), 4 otherwise.
template <bool BITMAP, int TILE, int KC>
__global__ void __launch_bounds__(kWalkWarps * 32, KC == 2 ? G2V_WALK_MINB : 6)
walk_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
            const uint32_t *__restrict__ qw, int32_t V, int32_t L, int32_t Lpad, int32_t H,
            int32_t hshift, uint64_t seed, uint32_t group, int64_t walker_begin,
            int64_t n_walkers, int64_t walker_stride, int32_t *__restrict__ out_nodes,
            int32_t *__restrict__ out_len, unsigned long long *__restrict__ ticket) {
    int32_t *const smem = g2v_walk_smem;
    constexpr int NT = 32 / TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = lane / TILE, tl = lane % TILE, tbase = tile * TILE;
    const unsigned tmask = TILE == 32 ? 0xffffffffu : (((1u << TILE) - 1u) << tbase);
    const int path = (warp * NT + tile) * (Lpad + H);   // offsets into smem (ints), not pointers
    const int hs = path + Lpad;
    const uint32_t hmask = (uint32_t)H - 1u;

    for (int i = tl; i < H; i += TILE) smem[hs + i] = BITMAP ? 0 : -1;
    __syncwarp(tmask);

    bool have = false, done = false, dirty = false;
    unsigned long long t = 0;
    uint64_t subseq = 0;
    int32_t cur = 0, n = 0, s = 0, dbase = -1;
    uint32_t dlo = 0, dhi = 0;                           // lane tl holds the draw of step dbase + tl

    while (true) {
        if (!have && !done) {                            // take the next walker
            unsigned long long tk = 0;
            if (tl == 0) tk = atomicAdd(ticket, 1ull);
            tk = __shfl_sync(tmask, tk, tbase);
            if ((int64_t)tk >= n_walkers) {
                done = true;
            } else {
                t = tk;
                const int64_t w = walker_begin + (int64_t)tk * walker_stride;
                subseq = ((uint64_t)group << 40) + (uint64_t)w;
                cur = (int32_t)(w % V);
                n = 0; s = 0; dbase = -1; dirty = false; have = true;
            }
        }
        if (TILE == 32) {
            if (done) break;                             // one walker per warp: nothing to wait for
        } else {
            if (__all_sync(0xffffffffu, done)) break;
            if (!have) continue;
        }

        // ---------------------------------------------------------------- one step of this tile's walker
        smem[path + n] = cur;                            // every lane of the tile stores the same value
        ++n;
        bool end = (s == L - 1);                         // the L-th node is appended, never expanded
        int32_t b = 0, e = 0;
        if (!end) {
            b = __ldg(rowptr + cur); e = __ldg(rowptr + cur + 1);
            end = (b == e);                              // no out-edges: dead end
        }
        if (!end) {
            if (BITMAP) {                                // visited.insert(cur), uniform across the tile
                const int32_t wv = smem[hs + (cur >> 5)];
                __syncwarp(tmask);
                smem[hs + (cur >> 5)] = wv | (1 << (cur & 31));
            } else {
                uint32_t i = hash_slot(cur, hshift);
                while (smem[hs + i] >= 0) i = (i + 1) & hmask;
                __syncwarp(tmask);
                smem[hs + i] = cur;
            }
            dirty = true;
            __syncwarp(tmask);

            // ---- pass 1: weight of the unvisited out-neighbours, per chunk of TILE (REDUX.SUM)
            uint32_t mq[KC], tot[KC];
            int32_t mc[KC];
            unsigned long long T = 0;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                mq[k] = 0; mc[k] = -1; tot[k] = 0;
                if (k == 0 || b + k * TILE < e) {        // tile-uniform (chunk 0 always exists: b < e)
                    const int32_t j = b + k * TILE + tl;
                    const bool in = j < e;
                    const int32_t c = in ? __ldg(col + j) : 0;       // predicated loads, no branch
                    const uint32_t q = in ? __ldg(qw + j) : 0u;
                    mc[k] = c;
                    mq[k] = unvisited_weight<BITMAP>(hs, hmask, hshift, c, q);
                    tot[k] = __reduce_add_sync(tmask, mq[k]);         // <= 32 * 2^24
                    T += tot[k];
                }
            }
            for (int32_t jb = b + KC * TILE; jb < e; jb += TILE) {   // rows longer than KC*TILE neighbours
                const int32_t j = jb + tl;
                uint32_t q = 0;
                if (j < e) q = unvisited_weight<BITMAP>(hs, hmask, hshift, __ldg(col + j), __ldg(qw + j));
                T += __reduce_add_sync(tmask, q);
            }
            const bool has_tail = b + KC * TILE < e;
            if (T == 0) {
                end = true;                              // every neighbour already visited
            } else {
                // ---- one 64-bit Philox draw per step, r uniform in [0, T); TILE steps are drawn at once,
                //      one per lane (counter-based: lane tl evaluates step dbase + tl)
                if ((s / TILE) * TILE != dbase) {
                    dbase = (s / TILE) * TILE;
                    const uint64_t d = draw64(seed, subseq, (uint32_t)(dbase + tl));
                    dlo = (uint32_t)d; dhi = (uint32_t)(d >> 32);
                }
                const int src = tbase + (s % TILE);
                const uint64_t x = ((uint64_t)__shfl_sync(tmask, dhi, src) << 32) | __shfl_sync(tmask, dlo, src);
                // r = floor(x*T / 2^64).  Without a tail T < 2^32 (KC chunk totals of at most 2^29), so the
                // product needs two 32x32 multiplies instead of a 64x64 high multiply.
                unsigned long long rem;                  // r - (weight of the chunks already skipped)
                if (!has_tail) {
                    const uint32_t T32 = (uint32_t)T;
                    const unsigned long long lo = (unsigned long long)(uint32_t)x * T32;
                    rem = ((unsigned long long)(uint32_t)(x >> 32) * T32 + (lo >> 32)) >> 32;
                } else {
                    rem = __umul64hi(x, T);
                }

                // ---- pass 2: chunk that contains r (tile-uniform scalar search), then one scan inside it
                int32_t nxt = -1;
                bool found = false;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    if (!found && (k == 0 || b + k * TILE < e)) {
                        if (rem < (unsigned long long)tot[k]) {
                            uint32_t incl = mq[k];
#pragma unroll
                            for (int o = 1; o < TILE; o <<= 1) {
                                const uint32_t up = __shfl_up_sync(tmask, incl, o, TILE);
                                if (tl >= o) incl += up;
                            }
                            const unsigned hit = __ballot_sync(tmask, incl > (uint32_t)rem);
                            nxt = __shfl_sync(tmask, mc[k], __ffs(hit) - 1);
                            found = true;
                        } else {
                            rem -= tot[k];
                        }
                    }
                }
                for (int32_t jb = b + KC * TILE; !found && jb < e; jb += TILE) {
                    const int32_t j = jb + tl;
                    int32_t c = -1;
                    uint32_t q = 0;
                    if (j < e) {
                        c = __ldg(col + j);
                        q = unvisited_weight<BITMAP>(hs, hmask, hshift, c, __ldg(qw + j));
                    }
                    const uint32_t ct = __reduce_add_sync(tmask, q);
                    if (rem < (unsigned long long)ct) {
                        uint32_t incl = q;
#pragma unroll
                        for (int o = 1; o < TILE; o <<= 1) {
                            const uint32_t up = __shfl_up_sync(tmask, incl, o, TILE);
                            if (tl >= o) incl += up;
                        }
                        const unsigned hit = __ballot_sync(tmask, incl > (uint32_t)rem);
                        nxt = __shfl_sync(tmask, c, __ffs(hit) - 1);
                        found = true;
                    } else {
                        rem -= ct;
                    }
                }
                cur = nxt;
                ++s;
            }
        }
        if (end) {                                       // walk finished: write the row once, coalesced
            __syncwarp(tmask);
            int32_t *row = out_nodes + (size_t)t * (size_t)L;
            for (int i = tl; i < L; i += TILE) row[i] = (i < n) ? smem[path + i] : -1;
            if (tl == 0) out_len[t] = n;
            if (dirty) {
                __syncwarp(tmask);
                if (BITMAP) {
                    for (int i = tl; i < n; i += TILE) smem[hs + (smem[path + i] >> 5)] = 0;   // only the touched words
                } else {
                    for (int i = tl; i < H; i += TILE) smem[hs + i] = -1;
                }
            }
            __syncwarp(tmask);
            have = false;
        }
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
    const int64_t n_walkers =
        walker_end > walker_begin ? (walker_end - walker_begin + walker_stride - 1) / walker_stride : 0;
    if (n_walkers == 0) return 0;                                 // empty range: nothing to write
    G2V_REQUIRE(rowptr && out_nodes && out_len && workspace, "g2v_walk_launch: null pointer");
    G2V_REQUIRE(E == 0 || (col && qw), "g2v_walk_launch: null col/qw with E > 0");
    DeviceProps dp;
    if (device_props(&dp)) return 1;
    G2V_REQUIRE(dp.cc_major == 10, "g2v_walk_launch: needs an sm_100 device (found sm_%d%d)", dp.cc_major, dp.cc_minor);

    // lanes per walker.  A full warp per walker is the fastest width on every graph measured -- syn10k
    // (mean degree 50): 3.3 ms at 32 lanes, 4.5 ms at 16, 6.0 ms at 8; ex_* (mean degree 3.4, 62 % of the
    // walks are singletons): 0.35 / 0.51 / 0.76 ms (profiles/README.md) -- so 8 and 16 are reachable only
    // through the G2V_WALK_TILE hook that the tests use.
    const double mean_deg = (double)E / (double)V;
    const char *ft = getenv("G2V_WALK_TILE");
    int tile = 32;
    if (ft && (atoi(ft) == 8 || atoi(ft) == 16 || atoi(ft) == 32)) tile = atoi(ft);
    const int nt = 32 / tile;
    // visited set per walker: a V-bit bitmap when a CTA's bitmaps fit 56 KB (>= 4 CTAs per SM), else a hash set
    const int Lpad = (L + 31) & ~31;
    const int bm_words = (V + 31) / 32;
    const char *force = getenv("G2V_WALK_VISITED");               // test hook: "hash" / "bitmap"
    int Hh = 64, hshift = 26;                                     // hash set: >= 3L slots, power of two
    while (Hh < 3 * L) { Hh <<= 1; --hshift; }
    const size_t per_tile = (size_t)kWalkWarps * nt * sizeof(int32_t);
    const size_t bm_smem = per_tile * (Lpad + bm_words), hash_smem = per_tile * (Lpad + Hh);
    bool bitmap = bm_smem <= 56 * 1024 || bm_smem <= hash_smem;   // occupancy first, then whichever is smaller
    if (force && force[0] == 'h') bitmap = false;
    const int H = bitmap ? bm_words : Hh;
    const size_t smem = (size_t)kWalkWarps * nt * (Lpad + H) * sizeof(int32_t);
    G2V_REQUIRE(smem <= (size_t)dp.max_smem_optin, "g2v_walk_launch: lenPath %d needs %zu B of shared memory", L, smem);
    cudaStream_t st = (cudaStream_t)stream;
    typedef void (*kern_t)(const int32_t *, const int32_t *, const uint32_t *, int32_t, int32_t, int32_t, int32_t,
                           int32_t, uint64_t, uint32_t, int64_t, int64_t, int64_t, int32_t *, int32_t *,
                           unsigned long long *);
    // chunks cached in registers: 2 when rows mostly fit 64 neighbours, 4 otherwise (G2V_WALK_KC overrides)
    const char *fk = getenv("G2V_WALK_KC");
    int kc = mean_deg <= 64.0 ? 2 : 4;   // measured: syn10k (deg 50) 2.96 vs 3.33 ms, syn20k (deg 100) 8.80 vs 8.18 ms
    if (fk && (atoi(fk) == 2 || atoi(fk) == 4)) kc = atoi(fk);
    if (tile != 32) kc = 4;
    static const kern_t table[2][4] = {
        {walk_kernel<false, 8, 4>, walk_kernel<false, 16, 4>, walk_kernel<false, 32, 4>, walk_kernel<false, 32, 2>},
        {walk_kernel<true, 8, 4>, walk_kernel<true, 16, 4>, walk_kernel<true, 32, 4>, walk_kernel<true, 32, 2>}};
    const int ti = tile == 8 ? 0 : (tile == 16 ? 1 : (kc == 4 ? 2 : 3));
    kern_t kern = table[bitmap][ti];
    // per-device function attributes (set on every call: the process may have switched device)
    G2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dp.max_smem_optin));
    G2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    G2V_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWalkWarps * 32, smem));
    G2V_REQUIRE(per_sm > 0, "g2v_walk_launch: kernel does not fit on an SM");
    int64_t grid = (int64_t)dp.sm_count * per_sm;                 // persistent: whole chip resident
    const int64_t need = (n_walkers + kWalkWarps * nt - 1) / (kWalkWarps * nt);
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
