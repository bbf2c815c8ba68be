// Diagnostics + process-wide state of libdsrl_b200.so.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace dsrl {

static thread_local char t_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_last_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

static std::once_flag g_dev_once;
static int g_sm_count = 0, g_cc_major = 0, g_dev_status = DSRL_ERR_CUDA;
static char g_dev_msg[256] = "";

static void probe_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        snprintf(g_dev_msg, sizeof(g_dev_msg), "no CUDA device (%s); libdsrl_b200 has no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    if (g_cc_major != 10) {
        snprintf(g_dev_msg, sizeof(g_dev_msg), "device compute capability %d.x; this library is sm_100a only", g_cc_major);
        return;
    }
    g_dev_status = DSRL_OK;
}

int require_device() {
    std::call_once(g_dev_once, probe_device);
    if (g_dev_status != DSRL_OK) set_last_error("%s", g_dev_msg);
    return g_dev_status;
}

int device_sm_count() {
    std::call_once(g_dev_once, probe_device);
    return g_sm_count > 0 ? g_sm_count : 148;
}


// Ticket words: kTicketSlots recycled round robin by eager launches (a word is back to 0 when its kernel ends, and 4096
// launches later that kernel has long finished), followed by kCaptureSlots handed out ONCE each to launches recorded into a
// CUDA graph -- a graph replays with the word it captured, so that word must never be given to anyone else.
static std::mutex g_ticket_mu;
static std::atomic<unsigned *> g_ticket_pool[64];
static std::atomic<unsigned> g_ticket_next{0};
static std::atomic<unsigned> g_capture_next[64];
constexpr unsigned kTicketSlots = 4096, kCaptureSlots = 16384;

unsigned *next_ticket_slot(cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        set_last_error("next_ticket_slot: cudaGetDevice failed");
        return nullptr;
    }
    unsigned *pool = g_ticket_pool[dev].load(std::memory_order_acquire);
    if (!pool) {
        std::lock_guard<std::mutex> lk(g_ticket_mu);
        pool = g_ticket_pool[dev].load(std::memory_order_acquire);
        if (!pool) {
            unsigned *p = nullptr;
            cudaError_t e = cudaMalloc(&p, (kTicketSlots + kCaptureSlots) * sizeof(unsigned));
            if (e == cudaSuccess) e = cudaMemset(p, 0, (kTicketSlots + kCaptureSlots) * sizeof(unsigned));
            if (e == cudaSuccess) e = cudaDeviceSynchronize();      // the memset is ordered before a first launch on ANY stream
            if (e != cudaSuccess) {
                set_last_error("ticket pool allocation failed (%s); if this happened inside a CUDA graph capture, run the "
                               "call once eagerly first", cudaGetErrorString(e));
                (void)cudaGetLastError();
                return nullptr;
            }
            g_ticket_pool[dev].store(p, std::memory_order_release);
            pool = p;
        }
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { (void)cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    if (cap == cudaStreamCaptureStatusActive) {
        const unsigned i = g_capture_next[dev].fetch_add(1, std::memory_order_relaxed);
        if (i >= kCaptureSlots) {
            set_last_error("more than %u kernel launches with a ticket word have been captured into CUDA graphs on device %d", kCaptureSlots, dev);
            return nullptr;
        }
        return pool + kTicketSlots + i;
    }
    return pool + (g_ticket_next.fetch_add(1, std::memory_order_relaxed) % kTicketSlots);
}

}  // namespace dsrl

extern "C" {
int dsrl_version(void) { return DSRL_B200_VERSION; }
const char *dsrl_last_error(void) { return dsrl::t_err; }
uint64_t dsrl_launch_count(void) { return dsrl::g_launches.load(std::memory_order_relaxed); }
}
