// g2v_capi.cu -- error plumbing, device facts and the curand cross-check hook of the C ABI.
#include <curand_kernel.h>
#include <stdarg.h>

#include <atomic>

#include "g2v_common.cuh"

namespace g2v {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local const int32_t *g_skip = nullptr;
const int32_t *loop_skip_flag() { return g_skip; }
void set_loop_skip_flag(const int32_t *p) { g_skip = p; }

int device_props(DeviceProps *out) {
    static thread_local int cached_dev = -1;
    static thread_local DeviceProps cached;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no usable CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
        return 1;
    }
    if (dev != cached_dev) {
        int l2 = 0;
        if ((e = cudaDeviceGetAttribute(&cached.sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&cached.cc_major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&cached.cc_minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&cached.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) {
            set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
            return 1;
        }
        cached.l2_bytes = l2;
        cached_dev = dev;
    }
    *out = cached;
    return 0;
}

__global__ void curand_draws_kernel(uint64_t seed, uint64_t subseq, int32_t n, uint64_t *out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, subseq, 0, &st);
        for (int i = 0; i < n; ++i) {
            uint32_t lo = curand(&st);
            uint32_t hi = curand(&st);
            out[i] = ((uint64_t)hi << 32) | lo;
        }
    }
}

}  // namespace g2v

using namespace g2v;

extern "C" int g2v_abi_version(void) { return G2V_ABI_VERSION; }
extern "C" const char *g2v_last_error(void) { return g_err; }
extern "C" int64_t g2v_launch_count(void) { return g_launches.load(); }

extern "C" int g2v_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, int64_t *l2_bytes) {
    DeviceProps dp;
    if (device_props(&dp)) return 1;
    if (sm_count) *sm_count = dp.sm_count;
    if (cc_major) *cc_major = dp.cc_major;
    if (cc_minor) *cc_minor = dp.cc_minor;
    if (l2_bytes) *l2_bytes = dp.l2_bytes;
    return 0;
}

extern "C" int g2v_test_curand_draws(uint64_t seed, uint64_t subsequence, int32_t n, uint64_t *out_dev,
                                     void *stream) {
    G2V_REQUIRE(n >= 0 && out_dev, "g2v_test_curand_draws: bad arguments");
    if (n == 0) return 0;
    curand_draws_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(seed, subsequence, n, out_dev);
    G2V_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
