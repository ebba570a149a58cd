// Library lifecycle, staging memory and launch accounting for the C ABI.
#include <cstdlib>
#include "common.cuh"
#include <atomic>
#include <mutex>

namespace stwo_b200 {
static std::atomic<uint64_t> g_launches{0};
static std::mutex g_mu;
static int g_device = -1;
static uint8_t *g_dev = nullptr;
static size_t g_dev_cap = 0;
static cudaStream_t g_stream = nullptr;

void note_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool device_ready() { return g_device >= 0; }
uint8_t *stage_dev() { return g_dev; }
cudaStream_t stage_stream() { return g_stream; }

int32_t stage_reserve(size_t dev_bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (dev_bytes <= g_dev_cap) return STWO_B200_OK;
    if (g_dev) cudaFree(g_dev);
    g_dev = nullptr; g_dev_cap = 0;
    size_t cap = align_up(dev_bytes + dev_bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&g_dev, cap);
    if (e != cudaSuccess) return -(int32_t)e;
    g_dev_cap = cap;
    return STWO_B200_OK;
}
}  // namespace stwo_b200

using namespace stwo_b200;

extern "C" int32_t stwo_b200_init(int32_t device) {
    // The batch drivers spread a pass over a dozen streams (worker pools, pipeline stages, shape groups).  With the driver's default of 8
    // hardware work queues, streams alias onto queues and independent chains wait for each other (measured on B200: 512-proof pipeline
    // 2.79 -> 2.33 ms, mixed 256-proof batch 14.6 -> 12.1 ms with 32).  Read when the context is created: this helps only if the library
    // is initialised before anything else touches CUDA in the process; a host that creates the context itself exports the variable itself.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return STWO_B200_E_NO_DEVICE;
    if (device < 0 || device >= n) return STWO_B200_E_BAD_ARG;
    STWO_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    STWO_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) return STWO_B200_E_NO_DEVICE;   // sm_100a code only; nothing else is built
    std::lock_guard<std::mutex> lk(g_mu);
    // One device per process: the staging area, the stream pools and events of the batch drivers and every uploaded circuit live on the
    // device of the first init.  A second init on ANOTHER device is refused instead of leaving those bound to the old one.
    if (g_device >= 0 && g_device != device) return STWO_B200_E_BAD_ARG;
    if (g_device != device) {
        STWO_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
        g_device = device;
    }
    return STWO_B200_OK;
}

extern "C" int32_t stwo_b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_dev) cudaFree(g_dev);
    g_dev = nullptr; g_dev_cap = 0;
    if (g_stream) cudaStreamDestroy(g_stream);
    g_stream = nullptr;
    stwo_b200::verify_pools_destroy();          // worker-stream pools and their events
    stwo_b200::circuit_streams_destroy();       // the side stream / events of the trace pass
    g_device = -1;
    return STWO_B200_OK;
}

extern "C" uint32_t stwo_b200_version(void) { return 0x00000100u; }
extern "C" uint64_t stwo_b200_launch_count(void) { return g_launches.load(); }
