// Library-wide plumbing of the C ABI (include/imk.h): error reporting, launch
// accounting, device probing, and the host-buffer pipelines that stand in for the
// per-directory loops of create_pseudo_labels_im_* (functions.py:2844-2887,
// 2932-2980, 3020-3066) minus PNG I/O.
#include <stdarg.h>
#include <algorithm>
#include <map>
#include <vector>
#include "imk_unet.cuh"

namespace imk {

static thread_local char g_err[1024] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int64_t &launch_counter() { return g_launches; }

int num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
    if (!cache[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
        cache[dev] = n;
    }
    return cache[dev];
}

// ---- per-kernel profiler ------------------------------------------------------------
struct ProfRec { const char *name; int tag; cudaEvent_t a, b; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> *g_prof = nullptr;
static thread_local std::vector<cudaEvent_t> *g_prof_pool = nullptr;

static cudaEvent_t prof_event() {
    if (!g_prof_pool) g_prof_pool = new std::vector<cudaEvent_t>();
    if (!g_prof_pool->empty()) { cudaEvent_t e = g_prof_pool->back(); g_prof_pool->pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

bool profiling_active() { return g_prof_on; }

ProfileScope::ProfileScope(const char *name, int tag, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof_on) return;
    ProfRec r{name, tag, prof_event(), prof_event()};
    cudaEventRecord(r.a, s);
    g_prof->push_back(r);
    slot = (int)g_prof->size() - 1;
}
ProfileScope::~ProfileScope() {
    if (slot >= 0) cudaEventRecord((*g_prof)[slot].b, stream);
}

}  // namespace imk

using namespace imk;

extern "C" int imk_version(void) { return IMK_VERSION; }
extern "C" const char *imk_last_error(void) { return g_err; }
extern "C" int64_t imk_launch_count(void) { return g_launches; }

extern "C" int imk_device_available(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n > 0 ? 1 : 0;
}

extern "C" int imk_profile_begin(void) {
    if (!g_prof) g_prof = new std::vector<ProfRec>();
    g_prof->clear();
    g_prof_on = true;
    return IMK_OK;
}

extern "C" int imk_profile_end(imk_profile_entry *out, int cap, int *n_out) {
    IMK_REQUIRE(n_out && (out || cap == 0), "imk_profile_end: NULL argument");
    g_prof_on = false;
    *n_out = 0;
    if (!g_prof) return IMK_OK;
    IMK_CUDA(cudaDeviceSynchronize());
    std::map<std::pair<std::string, int>, std::pair<int64_t, double>> acc;
    for (ProfRec &r : *g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto &e = acc[{r.name, r.tag}];
            e.first += 1;
            e.second += ms;
        } else {
            cudaGetLastError();
        }
        g_prof_pool->push_back(r.a);
        g_prof_pool->push_back(r.b);
    }
    g_prof->clear();
    int n = 0;
    for (auto &kv : acc) {
        if (n < cap) {
            snprintf(out[n].name, sizeof(out[n].name), "%s", kv.first.first.c_str());
            out[n].tag = kv.first.second;
            out[n].launches = kv.second.first;
            out[n].total_ms = kv.second.second;
        }
        ++n;
    }
    *n_out = n;
    return IMK_OK;
}

// ---------------------------------------------------------------------------------------
//  Host-buffer pipeline.  Two slots, each with its own stream and device buffers:
//  slot s uploads chunk i+1 while slot s^1 computes chunk i and downloads chunk i-1's
//  results -- copies run on the copy engines concurrently with the kernels.
// ---------------------------------------------------------------------------------------
namespace {

// Device-side staging of the host pipeline.  Allocated once per thread and grown on demand (no cudaMalloc /
// cudaFree -- both synchronise the device -- on the per-directory call path); three slots so that the upload of
// chunk i+1, the kernels of chunk i and the download of chunk i-1 overlap on three streams / two copy engines.
constexpr int kSlots = 3;

struct Slot {
    uint8_t *img = nullptr, *img_out = nullptr, *labels = nullptr, *im = nullptr;
    int64_t *im_size = nullptr, *pred_size = nullptr;
    uint8_t *lists_equal = nullptr;
    uint8_t *bits = nullptr;                // packed layout: label bit planes, then the IM bit plane
    cudaEvent_t up_done = nullptr, comp_done = nullptr, down_done = nullptr;
};

struct Pipeline {
    Slot slot[kSlots];
    cudaStream_t up = nullptr, compute = nullptr, down = nullptr;
    size_t cap_img = 0, cap_px = 0, cap_lab = 0;
    int64_t cap_chunk = 0;
    int device = -1;
    void release() {
        for (Slot &s : slot) {
            cudaFree(s.img); cudaFree(s.img_out); cudaFree(s.labels); cudaFree(s.im);
            cudaFree(s.im_size); cudaFree(s.pred_size); cudaFree(s.lists_equal); cudaFree(s.bits);
            if (s.up_done) cudaEventDestroy(s.up_done);
            if (s.comp_done) cudaEventDestroy(s.comp_done);
            if (s.down_done) cudaEventDestroy(s.down_done);
            s = Slot{};
        }
        if (up) cudaStreamDestroy(up);
        if (compute) cudaStreamDestroy(compute);
        if (down) cudaStreamDestroy(down);
        up = compute = down = nullptr;
        cap_img = cap_px = cap_lab = 0; cap_chunk = 0;
    }
    ~Pipeline() { release(); }
};

static thread_local Pipeline g_pipe;

int pipeline_reserve(Pipeline &P, int64_t chunk, size_t img_bytes, size_t px_bytes, size_t lab_bytes, int planes) {
    int dev = 0;
    IMK_CUDA(cudaGetDevice(&dev));
    if (P.device == dev && P.cap_chunk >= chunk && P.cap_img >= img_bytes && P.cap_px >= px_bytes && P.cap_lab >= lab_bytes) return IMK_OK;
    IMK_CUDA(cudaDeviceSynchronize());
    P.release();
    P.device = dev;
    IMK_CUDA(cudaStreamCreateWithFlags(&P.up, cudaStreamNonBlocking));
    IMK_CUDA(cudaStreamCreateWithFlags(&P.compute, cudaStreamNonBlocking));
    IMK_CUDA(cudaStreamCreateWithFlags(&P.down, cudaStreamNonBlocking));
    for (Slot &s : P.slot) {
        IMK_CUDA(cudaMalloc(&s.img, img_bytes));
        IMK_CUDA(cudaMalloc(&s.img_out, img_bytes));
        IMK_CUDA(cudaMalloc(&s.labels, lab_bytes));
        IMK_CUDA(cudaMalloc(&s.im, px_bytes));
        IMK_CUDA(cudaMalloc(&s.im_size, sizeof(int64_t) * chunk));
        IMK_CUDA(cudaMalloc(&s.pred_size, sizeof(int64_t) * chunk * (planes > 3 ? planes : 3)));
        IMK_CUDA(cudaMalloc(&s.lists_equal, (size_t)chunk));
        IMK_CUDA(cudaMalloc(&s.bits, (lab_bytes + px_bytes) / 8 + 64));
        IMK_CUDA(cudaEventCreateWithFlags(&s.up_done, cudaEventDisableTiming));
        IMK_CUDA(cudaEventCreateWithFlags(&s.comp_done, cudaEventDisableTiming));
        IMK_CUDA(cudaEventCreateWithFlags(&s.down_done, cudaEventDisableTiming));
    }
    P.cap_chunk = chunk; P.cap_img = img_bytes; P.cap_px = px_bytes; P.cap_lab = lab_bytes;
    return IMK_OK;
}

int run_host_pipeline(imk_unet_t *const *nets, int M, bool multiclass, const uint8_t *images, int64_t N, int swap_rb,
                      float thr, int strict, int block_in, int block_out,
                      uint8_t *img_out, uint8_t *labels, uint8_t *im, int64_t *im_size, int64_t *pred_size,
                      uint8_t *lists_equal, int64_t chunk, const char *who, bool packed = false) {
    IMK_REQUIRE(nets && M >= 1 && nets[0], "%s: no models", who);
    IMK_REQUIRE(images && N >= 0, "%s: NULL images or N < 0", who);
    if (N == 0) return IMK_OK;
    const imk_unet_desc &d = nets[0]->desc;
    const int64_t HW = (int64_t)d.height * d.width;
    const int K = d.num_outputmasks;
    const int planes = multiclass ? 1 : K;
    IMK_REQUIRE(!packed || HW % 8 == 0, "%s: the packed layout needs H*W to be a multiple of 8", who);
    // default: large chunks (kernel efficiency), but at least four of them in flight through the three slots -- and a
    // short ramp (a quarter and a half chunk) at both ends: the first upload and the last download are the only copies
    // nothing overlaps, so they are made small.  An explicit chunk > 0 is used as given.
    const bool auto_chunk = chunk <= 0;
    if (auto_chunk) chunk = std::min<int64_t>(kMaxChunk, std::max<int64_t>(64, (N + 3) / 4));
    chunk = std::min<int64_t>(chunk, N);
    std::vector<int64_t> sizes;
    {
        int64_t rest = N;
        const bool ramp = auto_chunk && chunk >= 128 && N >= 4 * chunk;
        if (ramp) { sizes.push_back(chunk / 4); sizes.push_back(chunk / 2); rest -= chunk / 4 + chunk / 2 + chunk / 2 + chunk / 4; }
        for (; rest > 0; rest -= chunk) sizes.push_back(std::min<int64_t>(chunk, rest));
        if (ramp) { sizes.push_back(chunk / 2); sizes.push_back(chunk / 4); }
    }
    Pipeline &P = g_pipe;
    int rc = pipeline_reserve(P, chunk, (size_t)chunk * HW * d.in_channels, (size_t)chunk * HW, (size_t)chunk * HW * planes, planes);
    if (rc) return rc;
    // Each model keeps ONE workspace, so the kernels of consecutive chunks are issued to a single compute stream;
    // uploads and downloads run on their own streams, ordered by events.
    const int64_t n_chunks = (int64_t)sizes.size();
    int64_t n_next = 0;
    for (int64_t i = 0; i < n_chunks; ++i) {
        Slot &S = P.slot[i % kSlots];
        const int64_t n0 = n_next, n = sizes[(size_t)i];
        n_next += n;
        // the slot is free again once chunk i - kSlots has been downloaded
        if (i >= kSlots) IMK_CUDA(cudaStreamWaitEvent(P.up, S.down_done, 0));
        IMK_CUDA(cudaMemcpyAsync(S.img, images + n0 * HW * d.in_channels, (size_t)n * HW * d.in_channels, cudaMemcpyHostToDevice, P.up));
        IMK_CUDA(cudaEventRecord(S.up_done, P.up));
        IMK_CUDA(cudaStreamWaitEvent(P.compute, S.up_done, 0));
        if (i >= kSlots) IMK_CUDA(cudaStreamWaitEvent(P.compute, S.down_done, 0));
        if (multiclass)
            rc = imk_ensemble_im_multiclass(nets, M, S.img, n, swap_rb, block_in, block_out, img_out ? S.img_out : nullptr, S.labels, S.im,
                                            S.im_size, lists_equal ? S.lists_equal : nullptr, P.compute);
        else
            rc = imk_ensemble_im_binary(nets, M, S.img, n, swap_rb, thr, strict, block_in, block_out, img_out ? S.img_out : nullptr, S.labels,
                                        S.im, S.im_size, pred_size ? S.pred_size : nullptr, P.compute);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        // packed layout: 0/255 planes leave the device as bits (binary label planes + IM; a multiclass label keeps its ids)
        const bool pack_lab = packed && !multiclass && labels, pack_im = packed && im;
        const int64_t lab_bits = (int64_t)planes * n * HW / 8;
        if (pack_lab && (rc = imk_pack_bits(S.labels, (int64_t)planes * n * HW, S.bits, P.compute))) { cudaDeviceSynchronize(); return rc; }
        if (pack_im && (rc = imk_pack_bits(S.im, n * HW, S.bits + lab_bits, P.compute))) { cudaDeviceSynchronize(); return rc; }
        IMK_CUDA(cudaEventRecord(S.comp_done, P.compute));
        IMK_CUDA(cudaStreamWaitEvent(P.down, S.comp_done, 0));
        if (img_out) IMK_CUDA(cudaMemcpyAsync(img_out + n0 * HW * d.in_channels, S.img_out, (size_t)n * HW * d.in_channels, cudaMemcpyDeviceToHost, P.down));
        if (labels && pack_lab)
            for (int k = 0; k < planes; ++k)
                IMK_CUDA(cudaMemcpyAsync(labels + ((int64_t)k * N * HW + n0 * HW) / 8, S.bits + (int64_t)k * n * HW / 8, (size_t)(n * HW / 8), cudaMemcpyDeviceToHost, P.down));
        else if (labels)
            for (int k = 0; k < planes; ++k)     // slot planes are n*HW apart, host planes N*HW apart
                IMK_CUDA(cudaMemcpyAsync(labels + (int64_t)k * N * HW + n0 * HW, S.labels + (int64_t)k * n * HW, (size_t)n * HW, cudaMemcpyDeviceToHost, P.down));
        if (im && pack_im) IMK_CUDA(cudaMemcpyAsync(im + n0 * HW / 8, S.bits + lab_bits, (size_t)(n * HW / 8), cudaMemcpyDeviceToHost, P.down));
        else if (im) IMK_CUDA(cudaMemcpyAsync(im + n0 * HW, S.im, (size_t)n * HW, cudaMemcpyDeviceToHost, P.down));
        if (im_size) IMK_CUDA(cudaMemcpyAsync(im_size + n0, S.im_size, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, P.down));
        if (pred_size && !multiclass)
            for (int k = 0; k < planes; ++k)
                IMK_CUDA(cudaMemcpyAsync(pred_size + (int64_t)k * N + n0, S.pred_size + (int64_t)k * n, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, P.down));
        if (lists_equal && multiclass) IMK_CUDA(cudaMemcpyAsync(lists_equal + n0, S.lists_equal, (size_t)n, cudaMemcpyDeviceToHost, P.down));
        IMK_CUDA(cudaEventRecord(S.down_done, P.down));
    }
    IMK_CUDA(cudaStreamSynchronize(P.down));
    IMK_CUDA(cudaStreamSynchronize(P.compute));
    return IMK_OK;
}

}  // namespace

extern "C" int imk_pseudo_label_binary_host(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                            float thr, int strict_gt, int block_in, int block_out,
                                            uint8_t *img_out_host, uint8_t *labels_host, uint8_t *im_host,
                                            int64_t *im_size_host, int64_t *pred_size_host, int64_t chunk) {
    return run_host_pipeline(nets, M, false, images_host, N, swap_rb, thr, strict_gt, block_in, block_out, img_out_host, labels_host,
                             im_host, im_size_host, pred_size_host, nullptr, chunk, "imk_pseudo_label_binary_host");
}

extern "C" int imk_pseudo_label_multiclass_host(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                                int block_in, int block_out,
                                                uint8_t *img_out_host, uint8_t *label_host, uint8_t *im_host,
                                                int64_t *im_size_host, uint8_t *lists_equal_host, int64_t chunk) {
    return run_host_pipeline(nets, M, true, images_host, N, swap_rb, 0.f, 1, block_in, block_out, img_out_host, label_host, im_host,
                             im_size_host, nullptr, lists_equal_host, chunk, "imk_pseudo_label_multiclass_host");
}

extern "C" int imk_pseudo_label_binary_host_packed(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                                   float thr, int strict_gt, int block_in, int block_out,
                                                   uint8_t *img_out_host, uint8_t *label_bits_host, uint8_t *im_bits_host,
                                                   int64_t *im_size_host, int64_t *pred_size_host, int64_t chunk) {
    return run_host_pipeline(nets, M, false, images_host, N, swap_rb, thr, strict_gt, block_in, block_out, img_out_host, label_bits_host,
                             im_bits_host, im_size_host, pred_size_host, nullptr, chunk, "imk_pseudo_label_binary_host_packed", true);
}

extern "C" int imk_pseudo_label_multiclass_host_packed(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                                       int block_in, int block_out,
                                                       uint8_t *img_out_host, uint8_t *label_host, uint8_t *im_bits_host,
                                                       int64_t *im_size_host, uint8_t *lists_equal_host, int64_t chunk) {
    return run_host_pipeline(nets, M, true, images_host, N, swap_rb, 0.f, 1, block_in, block_out, img_out_host, label_host, im_bits_host,
                             im_size_host, nullptr, lists_equal_host, chunk, "imk_pseudo_label_multiclass_host_packed", true);
}
