// common.cu — error plumbing, device info, shared small kernels, CUB scan/sort wrappers.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

namespace aprb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_free_events;
static std::mutex g_prof_mu;

static cudaEvent_t take_event() {
    if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(const char* name, cudaStream_t s, int nlaunch) : slot(-1), st(s) {
    g_launches.fetch_add(nlaunch, std::memory_order_relaxed);
    if (g_prof_on.load(std::memory_order_relaxed)) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        ProfRec r{name, take_event(), take_event()};
        cudaEventRecord(r.a, st);
        g_recs.push_back(r);
        slot = (int)g_recs.size() - 1;
    }
}
ProfScope::~ProfScope() {
    if (slot >= 0) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (slot < (int)g_recs.size()) cudaEventRecord(g_recs[slot].b, st);
    }
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return dev;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

__device__ void block_offsets(const int* __restrict__ lens, int B, int* __restrict__ off, int* s_part, int* s_carry) {
    if (B <= 64) {   // the hot path has B = 2: a sequential prefix by one thread beats a block scan
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int b = 0; b < B; ++b) { off[b] = acc; acc += lens[b]; }
            off[B] = acc;
        }
        return;
    }
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += blockDim.x) {   // chunked Hillis-Steele scan carrying a running prefix
        int i = base + threadIdx.x;
        int v = i < B ? lens[i] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < blockDim.x; d <<= 1) {
            int t = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = s_part[threadIdx.x];
        int carry = *s_carry;
        if (i < B) off[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) *s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[B] = *s_carry;
}

__global__ void setup_kernel(const int* __restrict__ lensA, int* __restrict__ offA, const int* __restrict__ lensB,
                             int* __restrict__ offB, int B, int* __restrict__ bbox, int* __restrict__ zero, int nzero) {
    __shared__ int s_part[256];
    __shared__ int s_carry;
    if (bbox)
        for (int i = threadIdx.x; i < 6 * B; i += blockDim.x) bbox[i] = (i % 6) < 3 ? 0x7FFFFFFF : (int)0x80000000;
    for (int i = threadIdx.x; i < nzero; i += blockDim.x) zero[i] = 0;
    block_offsets(lensA, B, offA, s_part, &s_carry);
    if (lensB) {
        __syncthreads();
        block_offsets(lensB, B, offB, s_part, &s_carry);
    }
}

__global__ void bbox_kernel(const float* __restrict__ pts, int N, const int* __restrict__ off, int B,
                            int* __restrict__ bbox) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < N;
    int b = valid ? find_cloud(off, B, i) : -1;
    float x = 0.f, y = 0.f, z = 0.f;
    if (valid) { x = pts[3 * (size_t)i]; y = pts[3 * (size_t)i + 1]; z = pts[3 * (size_t)i + 2]; }
    int mnx = valid ? f2ord(x) : 0x7FFFFFFF, mny = valid ? f2ord(y) : 0x7FFFFFFF, mnz = valid ? f2ord(z) : 0x7FFFFFFF;
    int mxx = valid ? f2ord(x) : (int)0x80000000, mxy = valid ? f2ord(y) : (int)0x80000000,
        mxz = valid ? f2ord(z) : (int)0x80000000;
    // warp-level reduction when the whole warp lies in one cloud (the common case)
    int b0 = __shfl_sync(0xffffffffu, b, 0);
    bool uniform = __all_sync(0xffffffffu, b == b0 || !valid) && b0 >= 0;
    if (uniform) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, d)); mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, d));
            mnz = min(mnz, __shfl_xor_sync(0xffffffffu, mnz, d)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, d)); mxz = max(mxz, __shfl_xor_sync(0xffffffffu, mxz, d));
        }
        if ((threadIdx.x & 31) == 0) {
            int* bb = bbox + 6 * b0;
            atomicMin(bb + 0, mnx); atomicMin(bb + 1, mny); atomicMin(bb + 2, mnz);
            atomicMax(bb + 3, mxx); atomicMax(bb + 4, mxy); atomicMax(bb + 5, mxz);
        }
    } else if (valid) {
        int* bb = bbox + 6 * b;
        atomicMin(bb + 0, mnx); atomicMin(bb + 1, mny); atomicMin(bb + 2, mnz);
        atomicMax(bb + 3, mxx); atomicMax(bb + 4, mxy); atomicMax(bb + 5, mxz);
    }
}

// ---- CUB wrappers -----------------------------------------------------------------------------------------------
size_t scan_temp_bytes(int n) {
    size_t bytes = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, n) != cudaSuccess) {
        (void)cudaGetLastError();
        bytes = (size_t)n / 64 + (1u << 16);  // no device visible: generous analytic bound
    }
    return align256(bytes + 256);
}

size_t sort_temp_bytes(int n) {
    size_t bytes = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                        (const int*)nullptr, (int*)nullptr, n) != cudaSuccess) {
        (void)cudaGetLastError();
        bytes = (size_t)n * 4 + (1u << 20);
    }
    return align256(bytes + 256);
}

int exclusive_scan_i32(const int* d_in, int* d_out, int n, void* d_temp, size_t temp_bytes, cudaStream_t st) {
    ProfScope ps("cub_exclusive_scan", st, 2);
    APRB_CUDA_OK(cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_in, d_out, n, st));
    return APRB_OK;
}

int sort_pairs_i32(const uint64_t* k_in, uint64_t* k_out, const int* v_in, int* v_out, int n, void* d_temp,
                   size_t temp_bytes, cudaStream_t st) {
    ProfScope ps("cub_radix_sort_pairs64", st, 10);
    APRB_CUDA_OK(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k_in, k_out, v_in, v_out, n, 0, 64, st));
    return APRB_OK;
}

int sort_pairs_i32(const uint32_t* k_in, uint32_t* k_out, const int* v_in, int* v_out, int n, void* d_temp,
                   size_t temp_bytes, cudaStream_t st) {
    ProfScope ps("cub_radix_sort_pairs32", st, 6);
    APRB_CUDA_OK(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k_in, k_out, v_in, v_out, n, 0, 32, st));
    return APRB_OK;
}

}  // namespace aprb

extern "C" {

int aprb_version(void) { return 100; }

long long aprb_launch_count(void) { return aprb::g_launches.load(); }

int aprb_prof_enable(int on) {
    aprb::g_prof_on.store(on ? 1 : 0);
    return APRB_OK;
}

// Synchronises the device, folds all recorded (name, start, end) event pairs into per-name totals and writes them as
// "name count total_ms\n" lines into buf (NUL-terminated, truncated to cap). Clears the records.
int aprb_prof_report(char* buf, size_t cap) {
    using namespace aprb;
    APRB_REQUIRE(buf && cap > 0, "null buffer");
    APRB_CUDA_OK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::vector<std::string> names;
    std::vector<double> ms;
    std::vector<long long> cnt;
    for (auto& r : g_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) { (void)cudaGetLastError(); t = 0.f; }
        size_t i = 0;
        for (; i < names.size(); ++i) if (names[i] == r.name) break;
        if (i == names.size()) { names.push_back(r.name); ms.push_back(0.0); cnt.push_back(0); }
        ms[i] += t; cnt[i] += 1;
        g_free_events.push_back(r.a); g_free_events.push_back(r.b);
    }
    g_recs.clear();
    std::string out;
    char line[256];
    for (size_t i = 0; i < names.size(); ++i) {
        snprintf(line, sizeof(line), "%s %lld %.6f\n", names[i].c_str(), cnt[i], ms[i]);
        out += line;
    }
    size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
    return APRB_OK;
}

const char* aprb_last_error(void) { return aprb::g_err; }

int aprb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    APRB_CUDA_OK(cudaGetDevice(&dev));
    if (sm_count) APRB_CUDA_OK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) APRB_CUDA_OK(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) APRB_CUDA_OK(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return APRB_OK;
}

}  // extern "C"
