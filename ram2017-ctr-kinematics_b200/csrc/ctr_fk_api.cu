// ctr_fk_api.cu — the extern "C" boundary (include/ctr_fk.h) over the CUDA kernels.
// No CPU fallback anywhere in this file: if a device call fails the CUDA error code is returned.
#include <cuda_runtime.h>

#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/ctr_fk.h"
#include "../host/ctr_sampler.h"
#include "ctr_fk_internal.h"
#include "ctr_fk_krobot.h"
#include "ctr_fk_launch.h"

using namespace ctrk;

struct ctr_fk_engine {
    ctr_robot_desc desc;
    KRobot         rb;
    int            key;
    int            device;
    int            njv;
    int            ntubes;
    int            sm_count;
    static constexpr int kStreams = 8;      // created; ctr_fk_batch_host uses host_streams of them
    int            host_streams; // pipeline depth of the host-pointer path (4; CTRFK_HOST_STREAMS overrides, for tuning)
    int            force_fine;   // -1: fine_sampling() decides which form of the sweep the tip kernels run; 0 / 1 force one
                                 // (CTRFK_FORCE_FINE, for tests and tuning)
    int            host_chunk_div;   // ... divided by this (1; CTRFK_HOST_CHUNK_DIV, for tuning)
    int            host_chunk_waves; // chunk size of the host-pointer path in waves of the tip kernel (1; CTRFK_HOST_CHUNK_WAVES)
    int            host_zero_copy;   // take the zero-copy form of ctr_fk_batch_host without being asked by CTR_FK_FLAG_HOST_ZEROCOPY
                                     // (0; CTRFK_HOST_ZEROCOPY=1 in the environment sets it)
    cudaStream_t   streams[kStreams];
    std::mutex     host_mu;      // serialises ctr_fk_batch_host / ctr_fk_single calls on one handle
    char*          single_buf;   // device scratch of ctr_fk_single (lazily allocated)
    char*          stage_buf;    // device staging slabs of ctr_fk_batch_host (grow-only, cached)
    size_t         stage_bytes;
    cudaMemPool_t  pool;         // private pool of the stream-ordered temporaries (binning order, tickets, staging
                                 // slabs of the host-pointer calls): freed blocks stay here for the next call and go
                                 // back to the driver with the handle, never parked in the device's default pool
                                 // that the caller's own allocator (torch, ...) shares
};

namespace {

// configurations below this count are not worth three extra launches for binning
constexpr int64_t kMinBinBatch = 8192;
// the binning kernel keeps 2 x (MaxSamples + 2) counters in shared memory
constexpr int kMaxBinSamples = 4096;
// host-pointer path: configurations per pipelined chunk
constexpr int64_t kHostChunk = 65536;   // 6.3 MB out per chunk: PCIe-bound, short pipeline fill/drain

struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

int cuda_fail(const char* what, cudaError_t e)
{
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

// The kernels use 16-byte vector loads/stores on joint vectors (even joint counts), poses and
// frames; device buffers from cudaMalloc / torch are 256-B aligned, sub-views may not be.
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// joint vectors are read as 16-byte pairs only when the joint count is even
int check_align(ctr_fk_handle h, const void* jv, const void* tip, const void* backbone)
{
    if (((h->njv & 1) == 0 && !aligned16(jv)) || !aligned16(tip) || !aligned16(backbone)) {
        set_error("joint-vector (even joint count), pose and backbone buffers must be 16-byte aligned");
        return CTR_FK_E_ARG;
    }
    return 0;
}

int check_common(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples)
{
    if (!h) { set_error("null engine handle"); return CTR_FK_E_NULL; }
    if (B < 0 || B > 2147483647ll) { set_error("batch size %lld out of range [0, 2^31)", (long long)B); return CTR_FK_E_ARG; }
    if (B > 0 && !jv) { set_error("null joint-value pointer"); return CTR_FK_E_NULL; }
    if (mode == CTR_FK_MODE_STEP) {
        if (!(step > 0.0) || !std::isfinite(step)) { set_error("arc-length step must be finite and > 0"); return CTR_FK_E_ARG; }
    } else if (mode == CTR_FK_MODE_NSAMPLE) {
        if (nsamples < 1) { set_error("nsamples must be >= 1"); return CTR_FK_E_ARG; }
    } else {
        set_error("unknown mode %d", mode);
        return CTR_FK_E_ARG;
    }
    return 0;
}

// Build `order` (longest sweeps first inside every window of kBinWindow = 4096 configurations) for a step-mode
// batch.  One stream-ordered allocation, freed by the caller after the solver launch.
cudaError_t make_order(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t ns,
                       int32_t** order_out, cudaStream_t st)
{
    int32_t* order = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync((void**)&order, sizeof(int32_t) * (size_t)B, h->pool, st);
    if (e != cudaSuccess) return e;
    e = launch_bin_local(h->key, h->rb, jv, B, mode, step, ns, order, st);
    if (e != cudaSuccess) { cudaFreeAsync(order, st); return e; }
    *order_out = order;
    return cudaSuccess;
}

// Device-side alias of a HOST range when the whole range is page-locked and mapped into the device's address
// space (cudaHostAlloc / cudaHostRegister, torch pin_memory()): kernels can then read and write it over PCIe
// directly.  nullptr for pageable memory (or a null pointer).
void* device_alias(const void* p, size_t bytes)
{
    if (!p || bytes == 0) return nullptr;
    cudaPointerAttributes lo, hi;
    if (cudaPointerGetAttributes(&lo, p) != cudaSuccess || cudaPointerGetAttributes(&hi, (const char*)p + bytes - 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    if (lo.type != cudaMemoryTypeHost || hi.type != cudaMemoryTypeHost || !lo.devicePointer || !hi.devicePointer) return nullptr;
    if ((const char*)hi.devicePointer - (const char*)lo.devicePointer != (ptrdiff_t)(bytes - 1)) return nullptr;
    return lo.devicePointer;
}

int run_batch(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
              uint32_t flags, double* tip, double* backbone, double* radius, int32_t* n_out,
              double* step_out, double* alpha_base, int32_t* status, cudaStream_t st)
{
    if (B == 0) return 0;
    FkArgs a;
    std::memset(&a, 0, sizeof a);
    a.jv = jv; a.B = B; a.mode = mode; a.step = step; a.nsamples = nsamples;
    a.tip = tip; a.n_out = n_out; a.step_out = step_out; a.alpha_base = alpha_base; a.status = status;
    a.backbone = backbone; a.radius = radius; a.soa = (flags & CTR_FK_FLAG_SOA) ? 1 : 0;

    int32_t* order = nullptr;
    if (mode == CTR_FK_MODE_STEP && !(flags & CTR_FK_FLAG_NO_BINNING) && B >= kMinBinBatch && h->rb.maxs <= kMaxBinSamples) {
        cudaError_t e = make_order(h, jv, B, mode, step, nsamples, &order, st);
        if (e != cudaSuccess) return cuda_fail("step-mode binning", e);
        a.order = order;
    }
    const bool strict = (flags & CTR_FK_FLAG_STRICT) != 0;
    a.fine = (h->force_fine >= 0) ? h->force_fine : (fine_sampling(h->rb, mode, step, nsamples) ? 1 : 0);
    cudaError_t e;
    if (backbone) e = strict ? launch_backbone_strict(h->key, h->rb, a, st) : launch_backbone_fast(h->key, h->rb, a, st);
    else          e = strict ? launch_tip_strict(h->key, h->rb, a, st) : launch_tip_fast(h->key, h->rb, a, st);
    if (order) cudaFreeAsync(order, st);
    if (e != cudaSuccess) return cuda_fail("solver kernel launch", e);
    return 0;
}

}  // namespace

extern "C" {

int ctr_fk_create(const ctr_robot_desc* desc, int device, ctr_fk_handle* out)
{
    if (!desc || !out) { set_error("ctr_fk_create: null argument"); return CTR_FK_E_NULL; }
    *out = nullptr;
    KRobot rb;
    int rc = make_krobot(desc, &rb);
    if (rc) { set_error("ctr_fk_create: invalid robot descriptor (code %d)", rc); return rc; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_create: no usable CUDA device (this library has no CPU path)", e);
    if (device < 0 || device >= ndev) { set_error("ctr_fk_create: device %d out of range (%d visible)", device, ndev); return CTR_FK_E_ARG; }
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return cuda_fail("ctr_fk_create: cudaSetDevice", g.err);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail("cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        set_error("ctr_fk_create: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return (int)cudaErrorNoKernelImageForDevice;
    }
    ctr_fk_engine* h = new (std::nothrow) ctr_fk_engine;
    if (!h) { set_error("out of host memory"); return CTR_FK_E_NOMEM; }
    h->desc = *desc;
    h->rb = rb;
    h->pool = nullptr;
    {
        cudaMemPoolProps pp = {};
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        cudaError_t pe = cudaMemPoolCreate(&h->pool, &pp);
        if (pe != cudaSuccess) { delete h; return cuda_fail("ctr_fk_create: cudaMemPoolCreate", pe); }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    h->key = layout_key(desc);
    h->device = device;
    h->ntubes = desc_ntubes(desc);
    h->njv = 1 + desc->nsections + h->ntubes;
    h->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < ctr_fk_engine::kStreams; ++i) h->streams[i] = nullptr;
    h->host_streams = 4;
    h->host_chunk_waves = 1;
    h->host_chunk_div = 1;
    h->force_fine = -1;
    if (const char* ev = std::getenv("CTRFK_FORCE_FINE")) { const int v = std::atoi(ev); if (v == 0 || v == 1) h->force_fine = v; }
    if (const char* ev = std::getenv("CTRFK_HOST_CHUNK_DIV")) { const int v = std::atoi(ev); if (v >= 1 && v <= 16) h->host_chunk_div = v; }
    if (const char* ev = std::getenv("CTRFK_HOST_STREAMS")) { const int v = std::atoi(ev); if (v >= 1 && v <= ctr_fk_engine::kStreams) h->host_streams = v; }
    h->host_zero_copy = 0;
    if (const char* ev = std::getenv("CTRFK_HOST_ZEROCOPY")) h->host_zero_copy = std::atoi(ev) != 0;
    if (const char* ev = std::getenv("CTRFK_HOST_CHUNK_WAVES")) { const int v = std::atoi(ev); if (v >= 1 && v <= 64) h->host_chunk_waves = v; }
    h->single_buf = nullptr;
    h->stage_buf = nullptr;
    h->stage_bytes = 0;
    for (int i = 0; i < ctr_fk_engine::kStreams; ++i) {
        if ((e = cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking)) != cudaSuccess) {
            for (int j = 0; j < i; ++j) cudaStreamDestroy(h->streams[j]);
            delete h;
            return cuda_fail("cudaStreamCreate", e);
        }
    }
    *out = h;
    return 0;
}

int ctr_fk_destroy(ctr_fk_handle h)
{
    if (!h) return 0;
    {
        DeviceGuard g(h->device);
        for (int i = 0; i < ctr_fk_engine::kStreams; ++i)
            if (h->streams[i]) { cudaStreamSynchronize(h->streams[i]); cudaStreamDestroy(h->streams[i]); }
        if (h->single_buf) cudaFree(h->single_buf);
        if (h->stage_buf) cudaFree(h->stage_buf);
        if (h->pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(h->pool); }
    }
    delete h;
    return 0;
}

int ctr_fk_get_desc(ctr_fk_handle h, ctr_robot_desc* out)
{
    if (!h || !out) { set_error("ctr_fk_get_desc: null argument"); return CTR_FK_E_NULL; }
    *out = h->desc;
    return 0;
}

int ctr_fk_batch(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
                 uint32_t flags, double* tip, double* backbone, double* radius, int32_t* n_out,
                 double* step_out, double* alpha_base, int32_t* status, void* stream)
{
    int rc = check_common(h, jv, B, mode, step, nsamples);
    if (rc) return rc;
    if ((rc = check_align(h, jv, tip, backbone))) return rc;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    return run_batch(h, jv, B, mode, step, nsamples, flags, tip, backbone, radius, n_out, step_out,
                     alpha_base, status, (cudaStream_t)stream);
}

int ctr_fk_batch_gather(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
                        uint32_t flags, void* const* tip_dst, int32_t ndst, int64_t first_row, int32_t* n_out,
                        double* step_out, int32_t* status, void* stream)
{
    int rc = check_common(h, jv, B, mode, step, nsamples);
    if (rc) return rc;
    if (flags & CTR_FK_FLAG_STRICT) { set_error("ctr_fk_batch_gather: FAST arithmetic only"); return CTR_FK_E_ARG; }
    const bool mc = (flags & CTR_FK_FLAG_GATHER_MULTICAST) != 0;
    if (!tip_dst || ndst < 1 || ndst > 8 || (mc && ndst != 1) || first_row < 0) {
        set_error("ctr_fk_batch_gather: 1..8 destination buffers (exactly 1 multicast address), first_row >= 0");
        return CTR_FK_E_ARG;
    }
    for (int p = 0; p < ndst; ++p)
        if (!tip_dst[p] || !aligned16(tip_dst[p])) { set_error("ctr_fk_batch_gather: destination %d null or not 16-byte aligned", p); return CTR_FK_E_ARG; }
    if (!aligned16(jv)) { set_error("ctr_fk_batch_gather: the joint-vector buffer must be 16-byte aligned"); return CTR_FK_E_ARG; }
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    cudaStream_t st = (cudaStream_t)stream;
    FkArgs a;
    std::memset(&a, 0, sizeof a);
    a.jv = jv; a.B = B; a.mode = mode; a.step = step; a.nsamples = nsamples;
    a.n_out = n_out; a.step_out = step_out; a.status = status;
    for (int p = 0; p < ndst; ++p) a.gather[p] = (double*)tip_dst[p];
    a.ngather = ndst; a.gather_mc = mc ? 1 : 0; a.gather_row = first_row;
    a.gather_rowwise = (flags & CTR_FK_FLAG_GATHER_ROWWISE) ? 1 : 0;
    int32_t* order = nullptr;
    if (mode == CTR_FK_MODE_STEP && !(flags & CTR_FK_FLAG_NO_BINNING) && B >= kMinBinBatch && h->rb.maxs <= kMaxBinSamples) {
        cudaError_t eo = make_order(h, jv, B, mode, step, nsamples, &order, st);
        if (eo != cudaSuccess) return cuda_fail("step-mode binning", eo);
        a.order = order;
    }
    a.fine = (h->force_fine >= 0) ? h->force_fine : (fine_sampling(h->rb, mode, step, nsamples) ? 1 : 0);
    cudaError_t e = launch_tip_gather(h->key, h->rb, a, st);
    if (order) cudaFreeAsync(order, st);
    if (e != cudaSuccess) return cuda_fail("solver kernel launch", e);
    return 0;
}

// ---- peer-visible buffers for ctr_fk_batch_gather (CUDA IPC; one process per GPU) -------------------------
int ctr_fk_peer_alloc(ctr_fk_handle h, size_t bytes, void** dptr, unsigned char* handle64)
{
    if (!h || !dptr || !handle64) { set_error("ctr_fk_peer_alloc: null argument"); return CTR_FK_E_NULL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_peer_alloc: cudaMalloc", e);
    cudaIpcMemHandle_t hd;
    e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail("ctr_fk_peer_alloc: cudaIpcGetMemHandle", e); }
    std::memcpy(handle64, &hd, 64);
    *dptr = p;
    return 0;
}

int ctr_fk_peer_open(ctr_fk_handle h, const unsigned char* handle64, void** dptr)
{
    if (!h || !dptr || !handle64) { set_error("ctr_fk_peer_open: null argument"); return CTR_FK_E_NULL; }
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_peer_open: cudaIpcOpenMemHandle", e);
    return 0;
}

int ctr_fk_peer_close(ctr_fk_handle h, void* dptr)
{
    if (!h) { set_error("null engine handle"); return CTR_FK_E_NULL; }
    if (!dptr) return 0;
    DeviceGuard g(h->device);
    cudaError_t e = cudaIpcCloseMemHandle(dptr);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_peer_close", e);
    return 0;
}

int ctr_fk_peer_free(ctr_fk_handle h, void* dptr)
{
    if (!h) { set_error("null engine handle"); return CTR_FK_E_NULL; }
    if (!dptr) return 0;
    DeviceGuard g(h->device);
    cudaError_t e = cudaFree(dptr);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_peer_free", e);
    return 0;
}

int ctr_fk_batch_host(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
                      uint32_t flags, double* tip, int32_t* n_out, double* step_out, int32_t* status)
{
    int rc = check_common(h, jv, B, mode, step, nsamples);
    if (rc) return rc;
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    std::lock_guard<std::mutex> lock(h->host_mu);

    // ---- zero-copy form (CTR_FK_FLAG_HOST_ZEROCOPY): page-locked host buffers are read and written by the solver
    // kernel itself.  One launch over the whole batch: every CTA fetches its joint vectors from host memory as whole
    // lines, solves, stages its 128 poses in shared memory and writes them to host memory as whole lines
    // (fk_tip_gather_kernel with the host buffer as its one destination).  No staging slabs, no chunk pipeline to fill
    // and drain.  Taken on request when every buffer passed in is page-locked (pinned or registered) and 16-byte
    // aligned and the arithmetic is FAST.  It was the default while the tip kernel took 2.04 ms per 1M configurations
    // (4.58e8 against 4.40e8 configurations/s for the copy pipeline); with the round-2 kernel (1.82 ms) the copy
    // engines are ahead (4.76e8 against 4.60e8: SM stores cross PCIe in smaller packets than the DMA engines'), so
    // the pipeline below is the default again and this form is for callers who want no device staging memory.
    // Step-mode batches that get binned stay on the staged form: binning permutes the rows of a CTA, and permuted
    // 96-byte rows would cross PCIe as scattered 16-byte writes (and the binning pass would read the joint vectors
    // over PCIe a second time).
    const bool binned = mode == CTR_FK_MODE_STEP && !(flags & CTR_FK_FLAG_NO_BINNING) && B >= kMinBinBatch && h->rb.maxs <= kMaxBinSamples;
    if ((h->host_zero_copy || (flags & CTR_FK_FLAG_HOST_ZEROCOPY)) && !binned && !(flags & (CTR_FK_FLAG_STRICT | CTR_FK_FLAG_HOST_STAGED | CTR_FK_FLAG_F32)) && tip && aligned16(jv) && aligned16(tip)) {
        const size_t nb = (size_t)B;
        double*  z_jv   = (double*)device_alias(jv, sizeof(double) * nb * h->njv);
        double*  z_tip  = (double*)device_alias(tip, sizeof(double) * nb * 12);
        int32_t* z_n    = n_out ? (int32_t*)device_alias(n_out, sizeof(int32_t) * nb) : nullptr;
        double*  z_step = step_out ? (double*)device_alias(step_out, sizeof(double) * nb) : nullptr;
        int32_t* z_stat = status ? (int32_t*)device_alias(status, sizeof(int32_t) * nb) : nullptr;
        if (z_jv && z_tip && (!n_out || z_n) && (!step_out || z_step) && (!status || z_stat)) {
            void* dst[1] = {z_tip};
            cudaStream_t st = h->streams[0];
            rc = ctr_fk_batch_gather(h, z_jv, B, mode, step, nsamples, flags, dst, 1, 0, z_n, z_step, z_stat, st);
            cudaError_t es = cudaStreamSynchronize(st);
            if (rc) return rc;
            if (es != cudaSuccess) return cuda_fail("ctr_fk_batch_host (zero-copy): sync", es);
            return 0;
        }
    }

    const int NS = h->host_streams;
    // one chunk = one full wave of the tip kernel (3 or 4 CTAs of kBlock threads per SM): a chunk that
    // fills only part of the resident slots would pay a whole wave time for less work
    int64_t wave = (int64_t)h->sm_count * tip_blocks_per_sm(h->ntubes, h->desc.nsections) * kBlock * h->host_chunk_waves;
    if (h->host_chunk_div > 1) wave = std::max<int64_t>(kBlock, wave / h->host_chunk_div / kBlock * kBlock);
    const int64_t chunk = std::min<int64_t>(wave > 0 ? wave : kHostChunk, B);
    const int njv = h->njv;
    // one slab per stream: jv | tip | n_out | step_out | status
    // every sub-buffer starts on a 256-B boundary (the kernels use 16-B vector accesses; an odd
    // joint count times an odd chunk would otherwise misalign the pose buffer)
    auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t off_tip  = up256(sizeof(double) * (size_t)chunk * njv);
    const size_t off_n    = off_tip + up256(sizeof(double) * (size_t)chunk * 12);
    const size_t off_step = off_n + up256(sizeof(int32_t) * (size_t)chunk);
    const size_t off_stat = off_step + up256(sizeof(double) * (size_t)chunk);
    const size_t slab     = off_stat + up256(sizeof(int32_t) * (size_t)chunk);
    const int nslab = (int)std::min<int64_t>(NS, (B + chunk - 1) / chunk);
    // staging slabs are cached in the handle: cudaMalloc/cudaFree per call would add milliseconds
    // and a device-wide synchronisation to every batch
    cudaError_t e = cudaSuccess;
    if (h->stage_bytes < slab * (size_t)nslab) {
        if (h->stage_buf) cudaFree(h->stage_buf);
        h->stage_buf = nullptr;
        h->stage_bytes = 0;
        e = cudaMalloc((void**)&h->stage_buf, slab * (size_t)nslab);
        if (e != cudaSuccess) return cuda_fail("ctr_fk_batch_host: device staging allocation", e);
        h->stage_bytes = slab * (size_t)nslab;
    }
    char* dbuf = h->stage_buf;

    int ret = 0;
    int64_t done = 0;
    for (int c = 0; done < B && ret == 0; ++c, done += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, B - done);
        cudaStream_t st = h->streams[c % NS];
        char* base = dbuf + slab * (size_t)(c % nslab);
        double*  d_jv   = (double*)base;
        double*  d_tip  = tip ? (double*)(base + off_tip) : nullptr;
        int32_t* d_n    = n_out ? (int32_t*)(base + off_n) : nullptr;
        double*  d_step = step_out ? (double*)(base + off_step) : nullptr;
        int32_t* d_stat = status ? (int32_t*)(base + off_stat) : nullptr;
        // stream order protects slab reuse: chunk c and chunk c+nslab share a stream
        if ((e = cudaMemcpyAsync(d_jv, jv + done * njv, sizeof(double) * (size_t)nb * njv, cudaMemcpyHostToDevice, st)) != cudaSuccess) { ret = cuda_fail("H2D copy", e); break; }
        ret = (flags & CTR_FK_FLAG_F32)
                  ? ctr_fk_batch_f32(h, d_jv, nb, mode, step, nsamples, flags & ~(uint32_t)(CTR_FK_FLAG_F32 | CTR_FK_FLAG_HOST_STAGED), d_tip, d_n, d_step, d_stat, st)
                  : run_batch(h, d_jv, nb, mode, step, nsamples, flags, d_tip, nullptr, nullptr, d_n, d_step, nullptr, d_stat, st);
        if (ret) break;
        if (tip && (e = cudaMemcpyAsync(tip + done * 12, d_tip, sizeof(double) * (size_t)nb * 12, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (n_out && (e = cudaMemcpyAsync(n_out + done, d_n, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (step_out && (e = cudaMemcpyAsync(step_out + done, d_step, sizeof(double) * (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (status && (e = cudaMemcpyAsync(status + done, d_stat, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
    }
    for (int i = 0; i < NS; ++i) {
        e = cudaStreamSynchronize(h->streams[i]);
        if (e != cudaSuccess && ret == 0) ret = cuda_fail("ctr_fk_batch_host: stream sync", e);
    }
    return ret;
}

// Host-pointer batch WITH the backbone: every frame and radius of every sample back to host memory
// ([B][max_samples][12], [B][max_samples]).  104 * max_samples bytes per configuration cross PCIe, so the path is
// bound by the D2H copy (26.6 KB per configuration at 256 samples: ~2e6 configurations/s at 55 GB/s); chunks are
// sized by their staging slab (<= 256 MB), not by the kernel's wave.  Own staging allocation per call
// (stream-ordered, from the cached pool); ctr_fk_batch_host's slabs are left alone.
int ctr_fk_batch_host_backbone(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
                               uint32_t flags, double* tip, double* backbone, double* radius, int32_t* n_out,
                               double* step_out, int32_t* status)
{
    int rc = check_common(h, jv, B, mode, step, nsamples);
    if (rc) return rc;
    if (flags & CTR_FK_FLAG_SOA) { set_error("ctr_fk_batch_host_backbone: [B][max_samples][12] layout only"); return CTR_FK_E_ARG; }
    if (!backbone) { set_error("ctr_fk_batch_host_backbone: null backbone pointer (use ctr_fk_batch_host for poses only)"); return CTR_FK_E_NULL; }
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    std::lock_guard<std::mutex> lock(h->host_mu);

    constexpr int NS = 3;
    const size_t maxs = (size_t)h->rb.maxs;
    const int    njv = h->njv;
    const size_t per_cfg = sizeof(double) * (njv + 12 + 12 * maxs + maxs + 1) + 2 * sizeof(int32_t);
    int64_t chunk = (int64_t)((size_t)(256u << 20) / per_cfg);
    chunk = std::max<int64_t>(kBlock, chunk / kBlock * kBlock);
    chunk = std::min<int64_t>(chunk, B);
    auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t off_tip  = up256(sizeof(double) * (size_t)chunk * njv);
    const size_t off_bb   = off_tip + up256(sizeof(double) * (size_t)chunk * 12);
    const size_t off_rad  = off_bb + up256(sizeof(double) * (size_t)chunk * 12 * maxs);
    const size_t off_n    = off_rad + up256(sizeof(double) * (size_t)chunk * maxs);
    const size_t off_step = off_n + up256(sizeof(int32_t) * (size_t)chunk);
    const size_t off_stat = off_step + up256(sizeof(double) * (size_t)chunk);
    const size_t slab     = off_stat + up256(sizeof(int32_t) * (size_t)chunk);
    const int nslab = (int)std::min<int64_t>(NS, (B + chunk - 1) / chunk);

    char* dbuf = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync((void**)&dbuf, slab * (size_t)nslab, h->pool, h->streams[0]);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_batch_host_backbone: device staging allocation", e);
    if ((e = cudaStreamSynchronize(h->streams[0])) != cudaSuccess) { cudaFreeAsync(dbuf, h->streams[0]); return cuda_fail("ctr_fk_batch_host_backbone: sync", e); }

    int ret = 0;
    int64_t done = 0;
    for (int c = 0; done < B && ret == 0; ++c, done += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, B - done);
        cudaStream_t st = h->streams[c % nslab];
        char* base = dbuf + slab * (size_t)(c % nslab);
        double*  d_jv   = (double*)base;
        double*  d_tip  = (double*)(base + off_tip);
        double*  d_bb   = (double*)(base + off_bb);
        double*  d_rad  = (double*)(base + off_rad);
        int32_t* d_n    = (int32_t*)(base + off_n);
        double*  d_step = (double*)(base + off_step);
        int32_t* d_stat = (int32_t*)(base + off_stat);
        if ((e = cudaMemcpyAsync(d_jv, jv + done * njv, sizeof(double) * (size_t)nb * njv, cudaMemcpyHostToDevice, st)) != cudaSuccess) { ret = cuda_fail("H2D copy", e); break; }
        ret = run_batch(h, d_jv, nb, mode, step, nsamples, flags, d_tip, d_bb, d_rad, d_n, d_step, nullptr, d_stat, st);
        if (ret) break;
        if ((e = cudaMemcpyAsync(backbone + (size_t)done * 12 * maxs, d_bb, sizeof(double) * (size_t)nb * 12 * maxs, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (radius && (e = cudaMemcpyAsync(radius + (size_t)done * maxs, d_rad, sizeof(double) * (size_t)nb * maxs, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (tip && (e = cudaMemcpyAsync(tip + done * 12, d_tip, sizeof(double) * (size_t)nb * 12, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (n_out && (e = cudaMemcpyAsync(n_out + done, d_n, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (step_out && (e = cudaMemcpyAsync(step_out + done, d_step, sizeof(double) * (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
        if (status && (e = cudaMemcpyAsync(status + done, d_stat, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { ret = cuda_fail("D2H copy", e); break; }
    }
    for (int i = 0; i < nslab; ++i) {
        e = cudaStreamSynchronize(h->streams[i]);
        if (e != cudaSuccess && ret == 0) ret = cuda_fail("ctr_fk_batch_host_backbone: stream sync", e);
    }
    cudaFreeAsync(dbuf, h->streams[0]);
    return ret;
}

// ---- page-locked host buffers placed for the engine's GPU ------------------------------------------------------
// The zero-copy form of ctr_fk_batch_host moves every byte over the GPU's PCIe link straight from / into the caller's
// pages, so WHERE those pages live matters on a multi-socket host: a buffer on the other socket makes every transfer
// cross the inter-socket link too, and eight ranks doing so share it.  ctr_fk_host_alloc binds the calling thread to
// the CPUs local to the GPU's PCIe root (sysfs local_cpulist of the device's bus id) while cudaHostAlloc faults the
// pages in — first touch puts them on that NUMA node — and restores the thread's affinity afterwards.
namespace {
bool gpu_local_cpus(int device, cpu_set_t* out)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    char path[128];
    std::snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = std::fopen(path, "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool got = std::fgets(line, sizeof line, f) != nullptr;
    std::fclose(f);
    if (!got) return false;
    CPU_ZERO(out);
    int n = 0;
    for (char* p = line; *p && *p != '\n';) {             // "0-15,32-47"
        char* end = nullptr;
        long a = std::strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        p = end;
        if (*p == '-') { b = std::strtol(p + 1, &end, 10); p = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, out); ++n; }
        if (*p == ',') ++p;
    }
    return n > 0;
}
}  // namespace

int ctr_fk_host_alloc(ctr_fk_handle h, size_t bytes, uint32_t flags, void** ptr)
{
    if (!h || !ptr) { set_error("ctr_fk_host_alloc: null argument"); return CTR_FK_E_NULL; }
    *ptr = nullptr;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    cpu_set_t before, local, both;
    bool bound = false;
    if (sched_getaffinity(0, sizeof before, &before) == 0 && gpu_local_cpus(h->device, &local)) {
        CPU_AND(&both, &before, &local);
        if (CPU_COUNT(&both) > 0 && !CPU_EQUAL(&both, &before)) bound = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    unsigned int fl = cudaHostAllocMapped | cudaHostAllocPortable;
    if (flags & CTR_FK_HOST_WRITE_COMBINED) fl |= cudaHostAllocWriteCombined;
    const cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 16, fl);
    if (bound) sched_setaffinity(0, sizeof before, &before);
    if (e != cudaSuccess) { *ptr = nullptr; return cuda_fail("ctr_fk_host_alloc: cudaHostAlloc", e); }
    return 0;
}

int ctr_fk_host_free(ctr_fk_handle h, void* ptr)
{
    if (!h) { set_error("null engine handle"); return CTR_FK_E_NULL; }
    if (!ptr) return 0;
    DeviceGuard g(h->device);
    const cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_host_free", e);
    return 0;
}

// CPUs local to the engine's GPU as a cpulist string ("0-15,32-47"); empty when the platform does not say.
int ctr_fk_host_local_cpus(ctr_fk_handle h, char* buf, size_t len)
{
    if (!h || !buf || len < 2) { set_error("ctr_fk_host_local_cpus: null argument"); return CTR_FK_E_NULL; }
    buf[0] = 0;
    cpu_set_t local;
    if (!gpu_local_cpus(h->device, &local)) return 0;
    size_t off = 0;
    for (int c = 0; c < CPU_SETSIZE;) {
        if (!CPU_ISSET(c, &local)) { ++c; continue; }
        int e = c;
        while (e + 1 < CPU_SETSIZE && CPU_ISSET(e + 1, &local)) ++e;
        char piece[32];
        if (e > c) std::snprintf(piece, sizeof piece, "%s%d-%d", off ? "," : "", c, e);
        else       std::snprintf(piece, sizeof piece, "%s%d", off ? "," : "", c);
        const size_t n = std::strlen(piece);
        if (off + n + 1 > len) break;
        std::memcpy(buf + off, piece, n + 1);
        off += n;
        c = e + 1;
    }
    return 0;
}

int ctr_fk_single(ctr_fk_handle h, const double* jv, int mode, double step, int32_t nsamples, uint32_t flags,
                  double* tip, double* backbone, double* radius, int32_t* n_out, double* step_out,
                  double* alpha_base)
{
    int rc = check_common(h, jv, 1, mode, step, nsamples);
    if (rc) return rc;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    std::lock_guard<std::mutex> lock(h->host_mu);
    const size_t maxs = (size_t)h->rb.maxs;
    // layout (doubles): jv[16] | tip[12] | step[1] | alpha[8] | n,status (1 double) | radius[maxs] | backbone[12 maxs]
    // the frames start on a 32-byte boundary (the kernels store them with 16- and 32-byte vectors;
    // an odd MaxSamples would otherwise leave them 8-byte aligned)
    const size_t o_tip = 16, o_step = 28, o_al = 29, o_n = 37, o_rad = 38, o_bb = (38 + maxs + 3) / 4 * 4;
    const size_t total = (o_bb + 12 * maxs) * sizeof(double);
    cudaError_t e;
    if (!h->single_buf && (e = cudaMalloc((void**)&h->single_buf, total)) != cudaSuccess)
        return cuda_fail("ctr_fk_single: scratch allocation", e);
    double* d = (double*)h->single_buf;
    cudaStream_t st = h->streams[0];
    if ((e = cudaMemcpyAsync(d, jv, sizeof(double) * (size_t)h->njv, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail("H2D copy", e);
    int32_t* d_n = (int32_t*)(d + o_n);
    rc = run_batch(h, d, 1, mode, step, nsamples, flags, d + o_tip, (backbone || radius) ? d + o_bb : nullptr,
                   radius ? d + o_rad : nullptr, d_n, d + o_step, d + o_al, d_n + 1, st);
    if (rc) return rc;
    double head[38];
    if ((e = cudaMemcpyAsync(head, d, sizeof head, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail("D2H copy", e);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail("ctr_fk_single: sync", e);
    int32_t nn[2];
    std::memcpy(nn, head + o_n, sizeof nn);
    if (tip) std::memcpy(tip, head + o_tip, 12 * sizeof(double));
    if (step_out) *step_out = head[o_step];
    if (alpha_base) std::memcpy(alpha_base, head + o_al, sizeof(double) * (size_t)h->ntubes);
    if (n_out) *n_out = nn[0];
    if (nn[0] > 0) {
        if (radius && (e = cudaMemcpy(radius, d + o_rad, sizeof(double) * (size_t)nn[0], cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail("D2H copy", e);
        if (backbone && (e = cudaMemcpy(backbone, d + o_bb, sizeof(double) * 12 * (size_t)nn[0], cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail("D2H copy", e);
    }
    return nn[1] ? CTR_FK_E_RANGE : 0;
}

int ctr_fk_batch_f32(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
                     uint32_t flags, double* tip, int32_t* n_out, double* step_out, int32_t* status, void* stream)
{
    int rc = check_common(h, jv, B, mode, step, nsamples);
    if (rc) return rc;
    if ((rc = check_align(h, jv, tip, nullptr))) return rc;
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    FkArgs a;
    std::memset(&a, 0, sizeof a);
    a.jv = jv; a.B = B; a.mode = mode; a.step = step; a.nsamples = nsamples;
    a.tip = tip; a.n_out = n_out; a.step_out = step_out; a.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* order = nullptr;
    if (mode == CTR_FK_MODE_STEP && !(flags & CTR_FK_FLAG_NO_BINNING) && B >= kMinBinBatch && h->rb.maxs <= kMaxBinSamples) {
        cudaError_t e = make_order(h, jv, B, mode, step, nsamples, &order, st);
        if (e != cudaSuccess) return cuda_fail("step-mode binning", e);
        a.order = order;
    }
    cudaError_t e = launch_tip_f32(h->key, h->rb, a, st);
    if (order) cudaFreeAsync(order, st);
    if (e != cudaSuccess) return cuda_fail("f32 solver kernel launch", e);
    return 0;
}

int ctr_fk_jacobian_batch_ex(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
                             int32_t i_sample, double fd_trans, double fd_rot, uint32_t flags, double* jac_tip,
                             double* jac_sample, double* tip, double* sample, int32_t* n_out, void* stream)
{
    const bool given = (flags & CTR_FK_FLAG_JAC_GIVEN_POSES) != 0;
    if (given && B > 0 && (!tip || (i_sample >= 0 && !sample))) {
        set_error("CTR_FK_FLAG_JAC_GIVEN_POSES needs the tip (and sample) poses as inputs");
        return CTR_FK_E_NULL;
    }
    int rc = check_common(h, jv, B, mode, step, nsamples);
    if (rc) return rc;
    if ((rc = check_align(h, jv, tip, sample))) return rc;
    if (!jac_tip && B > 0) { set_error("null Jacobian pointer"); return CTR_FK_E_NULL; }
    if (i_sample >= 0 && !jac_sample && B > 0) { set_error("i_sample >= 0 needs a jac_sample buffer"); return CTR_FK_E_NULL; }
    if (i_sample >= h->rb.maxs) { set_error("i_sample %d out of range (MaxSamples %d)", i_sample, h->rb.maxs); return CTR_FK_E_ARG; }
    if (!(fd_trans != 0.0) || !(fd_rot != 0.0) || !std::isfinite(fd_trans) || !std::isfinite(fd_rot)) {
        set_error("finite-difference steps must be finite and non-zero");
        return CTR_FK_E_ARG;
    }
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    JacArgs a;
    a.jv = jv; a.B = B; a.nsamples = nsamples; a.fd_trans = fd_trans; a.fd_rot = fd_rot; a.jac = jac_tip; a.tip = tip;
    a.mode = mode; a.step = step; a.i_sample = i_sample < 0 ? -1 : i_sample;
    a.jac_sample = i_sample >= 0 ? jac_sample : nullptr; a.sample = i_sample >= 0 ? sample : nullptr; a.n_out = n_out;
    a.given = given ? 1 : 0;
    a.order = nullptr;
    int32_t* order = nullptr;
    if (mode == CTR_FK_MODE_STEP && !(flags & CTR_FK_FLAG_NO_BINNING) && B >= kMinBinBatch && h->rb.maxs <= kMaxBinSamples) {
        cudaError_t eo = make_order(h, jv, B, mode, step, nsamples, &order, (cudaStream_t)stream);
        if (eo != cudaSuccess) return cuda_fail("step-mode binning", eo);
        a.order = order;
    }
    cudaError_t e = (flags & CTR_FK_FLAG_STRICT) ? launch_jacobian_strict(h->key, h->rb, a, (cudaStream_t)stream)
                                                 : launch_jacobian(h->key, h->rb, a, (cudaStream_t)stream);
    if (order) cudaFreeAsync(order, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail("jacobian kernel launch", e);
    return 0;
}

int ctr_fk_jacobian_batch(ctr_fk_handle h, const double* jv, int64_t B, int32_t nsamples, double fd_trans,
                          double fd_rot, double* jac, double* tip, void* stream)
{
    return ctr_fk_jacobian_batch_ex(h, jv, B, CTR_FK_MODE_NSAMPLE, 0.0, nsamples, -1, fd_trans, fd_rot, 0, jac, nullptr,
                                    tip, nullptr, nullptr, stream);
}

int ctr_fk_jacobian_host_ex(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step, int32_t nsamples,
                            int32_t i_sample, double fd_trans, double fd_rot, uint32_t flags, double* jac_tip,
                            double* jac_sample, double* tip, double* sample, int32_t* n_out)
{
    const bool given = (flags & CTR_FK_FLAG_JAC_GIVEN_POSES) != 0;
    if (given && B > 0 && (!tip || (i_sample >= 0 && !sample))) {
        set_error("CTR_FK_FLAG_JAC_GIVEN_POSES needs the tip (and sample) poses as inputs");
        return CTR_FK_E_NULL;
    }
    int rc = check_common(h, jv, B, mode, step, nsamples);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!jac_tip) { set_error("null Jacobian pointer"); return CTR_FK_E_NULL; }
    if (i_sample >= 0 && !jac_sample) { set_error("i_sample >= 0 needs a jac_sample buffer"); return CTR_FK_E_NULL; }
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    const size_t njv = (size_t)h->njv;
    auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
    const bool   ws = i_sample >= 0;
    const size_t b_jv = up256(sizeof(double) * njv * (size_t)B), b_jac = up256(sizeof(double) * 6 * njv * (size_t)B),
                 b_tip = up256(sizeof(double) * 12 * (size_t)B), b_n = up256(sizeof(int32_t) * (size_t)B);
    // stream-ordered allocation from the device pool (kept cached, see ctr_fk_create): no device-wide
    // synchronisation per call, which matters for the one-configuration calls of ctrb200::Robot
    char* d = nullptr;
    cudaStream_t st = h->streams[0];
    cudaError_t e = cudaMallocFromPoolAsync((void**)&d, b_jv + 2 * b_jac + 2 * b_tip + b_n, h->pool, st);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_jacobian_host: allocation", e);
    double*  d_jv   = (double*)d;
    double*  d_jac  = (double*)(d + b_jv);
    double*  d_jacs = (double*)(d + b_jv + b_jac);
    double*  d_tip  = (double*)(d + b_jv + 2 * b_jac);
    double*  d_smp  = (double*)(d + b_jv + 2 * b_jac + b_tip);
    int32_t* d_n    = (int32_t*)(d + b_jv + 2 * b_jac + 2 * b_tip);
    e = cudaMemcpyAsync(d_jv, jv, sizeof(double) * njv * (size_t)B, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && given) e = cudaMemcpyAsync(d_tip, tip, sizeof(double) * 12 * (size_t)B, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && given && ws) e = cudaMemcpyAsync(d_smp, sample, sizeof(double) * 12 * (size_t)B, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        rc = ctr_fk_jacobian_batch_ex(h, d_jv, B, mode, step, nsamples, i_sample, fd_trans, fd_rot, flags, d_jac, ws ? d_jacs : nullptr,
                                      tip ? d_tip : nullptr, (ws && sample) ? d_smp : nullptr, n_out ? d_n : nullptr, st);
        if (rc == 0) {
            e = cudaMemcpyAsync(jac_tip, d_jac, sizeof(double) * 6 * njv * (size_t)B, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && ws) e = cudaMemcpyAsync(jac_sample, d_jacs, sizeof(double) * 6 * njv * (size_t)B, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && tip && !given) e = cudaMemcpyAsync(tip, d_tip, sizeof(double) * 12 * (size_t)B, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && ws && sample && !given) e = cudaMemcpyAsync(sample, d_smp, sizeof(double) * 12 * (size_t)B, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && n_out) e = cudaMemcpyAsync(n_out, d_n, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, st);
        }
    }
    cudaFreeAsync(d, st);
    cudaError_t es = cudaStreamSynchronize(st);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail("ctr_fk_jacobian_host: copy", e);
    if (es != cudaSuccess) return cuda_fail("ctr_fk_jacobian_host: sync", es);
    return 0;
}

int ctr_fk_jacobian_host(ctr_fk_handle h, const double* jv, int64_t B, int32_t nsamples, double fd_trans,
                         double fd_rot, double* jac, double* tip)
{
    return ctr_fk_jacobian_host_ex(h, jv, B, CTR_FK_MODE_NSAMPLE, 0.0, nsamples, -1, fd_trans, fd_rot, 0, jac, nullptr, tip,
                                   nullptr, nullptr);
}

namespace {
// The scan `scan += angle_step; scan <= end` (robot_i.h:793, :861) must advance everywhere on [0, 2 pi]: a step that
// is absorbed by rounding would scan forever (the reference hangs; a persistent kernel would hang the GPU).
int check_angle_step(double angle_step, double min_degree)
{
    const double two_pi = 2.0 * 3.14159265358979323846;
    if (!(angle_step > 0.0) || !std::isfinite(angle_step) || !std::isfinite(min_degree) || !(two_pi + angle_step > two_pi)) {
        set_error("angle step must be finite, > 0 and large enough to advance a scan at 2 pi (> 4.5e-16); min_degree finite");
        return CTR_FK_E_ARG;
    }
    return 0;
}
}  // namespace

int ctr_fk_stability_batch(ctr_fk_handle h, const double* jv, int64_t B, double angle_step, double arc_step,
                           double min_degree, int32_t* stable, double* degree, void* stream)
{
    return ctr_fk_stability_batch_ex(h, jv, B, angle_step, arc_step, min_degree, 0, stable, degree, stream);
}

int ctr_fk_stability_batch_ex(ctr_fk_handle h, const double* jv, int64_t B, double angle_step, double arc_step,
                              double min_degree, uint32_t flags, int32_t* stable, double* degree, void* stream)
{
    int rc = check_common(h, jv, B, CTR_FK_MODE_STEP, arc_step, 0);
    if (rc) return rc;
    if ((rc = check_align(h, jv, nullptr, nullptr))) return rc;
    if ((rc = check_angle_step(angle_step, min_degree))) return rc;
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    StabArgs a;
    a.jv = jv; a.B = B; a.angle_step = angle_step; a.arc_step = arc_step; a.min_degree = min_degree;
    a.stable = stable; a.degree = degree;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* ticket = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync((void**)&ticket, sizeof(unsigned long long), h->pool, st);
    if (e != cudaSuccess) return cuda_fail("ticket allocation", e);
    e = cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), st);
    a.ticket = ticket;
    if (e == cudaSuccess) e = (flags & CTR_FK_FLAG_STRICT) ? launch_stability_strict(h->key, h->rb, a, h->sm_count, st)
                                                           : launch_stability(h->key, h->rb, a, h->sm_count, st);
    cudaFreeAsync(ticket, st);
    if (e != cudaSuccess) return cuda_fail("stability kernel launch", e);
    return 0;
}

int ctr_fk_stability_host(ctr_fk_handle h, const double* jv, int64_t B, double angle_step, double arc_step,
                          double min_degree, int32_t* stable, double* degree)
{
    return ctr_fk_stability_host_ex(h, jv, B, angle_step, arc_step, min_degree, 0, stable, degree);
}

int ctr_fk_stability_host_ex(ctr_fk_handle h, const double* jv, int64_t B, double angle_step, double arc_step,
                             double min_degree, uint32_t flags, int32_t* stable, double* degree)
{
    int rc = check_common(h, jv, B, CTR_FK_MODE_STEP, arc_step, 0);
    if (rc) return rc;
    if ((rc = check_angle_step(angle_step, min_degree))) return rc;
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t n_jv = sizeof(double) * (size_t)h->njv * (size_t)B;
    const size_t b_jv = up256(n_jv), b_deg = up256(sizeof(double) * (size_t)B), b_st = sizeof(int32_t) * (size_t)B;
    char* d = nullptr;
    cudaStream_t st = h->streams[0];
    cudaError_t e = cudaMallocFromPoolAsync((void**)&d, b_jv + b_deg + b_st, h->pool, st);
    if (e != cudaSuccess) return cuda_fail("ctr_fk_stability_host: allocation", e);
    double* d_jv = (double*)d;
    double* d_deg = (double*)(d + b_jv);
    int32_t* d_st = (int32_t*)(d + b_jv + b_deg);
    e = cudaMemcpyAsync(d_jv, jv, n_jv, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        rc = ctr_fk_stability_batch_ex(h, d_jv, B, angle_step, arc_step, min_degree, flags, d_st, d_deg, st);
        if (rc == 0) {
            if (stable) e = cudaMemcpyAsync(stable, d_st, b_st, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && degree) e = cudaMemcpyAsync(degree, d_deg, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st);
        }
    }
    cudaFreeAsync(d, st);
    cudaError_t es = cudaStreamSynchronize(st);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail("ctr_fk_stability_host: copy", e);
    if (es != cudaSuccess) return cuda_fail("ctr_fk_stability_host: sync", es);
    return 0;
}

int ctr_fk_batch_base(ctr_fk_handle h, const double* jv_base, int64_t B, int mode, double step, int32_t nsamples,
                      double tol, int32_t max_iter, double* jv_tip, double* tip, int32_t* iters, int32_t* status,
                      void* stream)
{
    return ctr_fk_batch_base_ex(h, jv_base, nullptr, B, mode, step, nsamples, tol, max_iter, 0, jv_tip, tip, iters, status,
                                nullptr, stream);
}

int ctr_fk_batch_base_ex(ctr_fk_handle h, const double* jv_base, const double* jv_guess, int64_t B, int mode, double step,
                         int32_t nsamples, double tol, int32_t max_iter, uint32_t flags, double* jv_tip, double* tip,
                         int32_t* iters, int32_t* status, double* resid, void* stream)
{
    int rc = check_common(h, jv_base, B, mode, step, nsamples);
    if (rc) return rc;
    if ((rc = check_align(h, jv_base, tip, nullptr)) || (rc = check_align(h, jv_tip, nullptr, nullptr)) ||
        (rc = check_align(h, jv_guess, nullptr, nullptr))) return rc;
    if (flags & ~(uint32_t)CTR_FK_FLAG_RK4) { set_error("ctr_fk_batch_base_ex: unknown flag (CTR_FK_FLAG_RK4 only)"); return CTR_FK_E_ARG; }
    if (B > 0 && !jv_tip) { set_error("jv_tip output is required"); return CTR_FK_E_NULL; }
    if (!(tol > 0.0) || !std::isfinite(tol) || max_iter < 1) { set_error("tol must be > 0 and max_iter >= 1"); return CTR_FK_E_ARG; }
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* ticket = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync((void**)&ticket, sizeof(unsigned long long), h->pool, st);
    if (e != cudaSuccess) return cuda_fail("ticket allocation", e);
    e = cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), st);
    int32_t* order = nullptr;
    if (e == cudaSuccess && mode == CTR_FK_MODE_STEP && B >= kMinBinBatch && h->rb.maxs <= kMaxBinSamples)
        e = make_order(h, jv_base, B, mode, step, nsamples, &order, st);   // longest sweeps first
    if (e == cudaSuccess) {
        ShootArgs a;
        a.jv_base = jv_base; a.B = B; a.mode = mode; a.step = step; a.nsamples = nsamples; a.tol = tol;
        a.max_iter = max_iter; a.order = order; a.jv_tip = jv_tip; a.iters = iters; a.status = status; a.ticket = ticket;
        a.jv_guess = jv_guess; a.resid = resid; a.rk4 = (flags & CTR_FK_FLAG_RK4) ? 1 : 0;
        e = launch_shoot(h->key, h->rb, a, h->sm_count, st);
    }
    if (order) cudaFreeAsync(order, st);
    cudaFreeAsync(ticket, st);
    if (e != cudaSuccess) return cuda_fail("shooting kernel launch", e);
    if (tip) return run_batch(h, jv_tip, B, mode, step, nsamples, 0u, tip, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st);
    return 0;
}

int ctr_fk_sample_joints_scaled(ctr_fk_handle h, uint64_t seed, int64_t first_config, int64_t B, int policy,
                                double p0, double p1, int32_t lut_size, double* jv, void* stream)
{
    if (!h) { set_error("null engine handle"); return CTR_FK_E_NULL; }
    if (B < 0 || first_config < 0) { set_error("negative batch size or offset"); return CTR_FK_E_ARG; }
    if (policy < CTR_FK_SCALE_NONE || policy > CTR_FK_SCALE_PROPORTIONAL || !std::isfinite(p0) || !std::isfinite(p1) ||
        lut_size < 0 || (policy == CTR_FK_SCALE_GAMMA && !(p0 > 0.0)) ||
        (policy == CTR_FK_SCALE_PROPORTIONAL && !(p0 >= 0.0 && p1 >= p0 && p1 <= 1.0))) {
        set_error("bad scale policy parameters");
        return CTR_FK_E_ARG;
    }
    if (B == 0) return 0;
    if (!jv) { set_error("null joint-value pointer"); return CTR_FK_E_NULL; }
    DeviceGuard g(h->device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    SamplerArgs a;
    make_sampler_table(&h->desc, &a.tb);
    a.tb.policy = policy; a.tb.p0 = p0; a.tb.p1 = p1; a.tb.lut_size = lut_size;
    a.seed = seed; a.first = first_config; a.B = B; a.jv = jv;
    cudaError_t e = launch_sampler(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail("sampler kernel launch", e);
    return 0;
}

int ctr_fk_sample_joints(ctr_fk_handle h, uint64_t seed, int64_t first_config, int64_t B, double* jv, void* stream)
{
    return ctr_fk_sample_joints_scaled(h, seed, first_config, B, CTR_FK_SCALE_NONE, 1.0, 1.0, 0, jv, stream);
}

static int peak_probe(int device, int32_t iters, double* flops_issued, void* stream, bool f32);

int ctr_fk_fp64_probe(int device, int32_t iters, double* flops_issued, void* stream) { return peak_probe(device, iters, flops_issued, stream, false); }
int ctr_fk_fp32_probe(int device, int32_t iters, double* flops_issued, void* stream) { return peak_probe(device, iters, flops_issued, stream, true); }

static int peak_probe(int device, int32_t iters, double* flops_issued, void* stream, bool f32)
{
    if (iters < 1) { set_error("iters must be >= 1"); return CTR_FK_E_ARG; }
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return cuda_fail("cudaSetDevice", g.err);
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return cuda_fail("cudaDeviceGetAttribute", e);
    static double* sinks[64] = {nullptr};   // one 8-byte word per device, never written
    static std::mutex mu;
    double* sink = nullptr;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (device < 0 || device >= 64) { set_error("device index out of range"); return CTR_FK_E_ARG; }
        if (!sinks[device] && (e = cudaMalloc((void**)&sinks[device], sizeof(double))) != cudaSuccess) return cuda_fail("cudaMalloc", e);
        sink = sinks[device];
    }
    const int blocks = sms * 8;
    e = f32 ? launch_fp32x2_probe(iters, blocks, sink, (cudaStream_t)stream) : launch_fp64_probe(iters, blocks, sink, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail("probe kernel launch", e);
    if (flops_issued)
        *flops_issued = (f32 ? 4.0 : 2.0) * (double)blocks * kProbeThreads * kProbeChains * 8.0 * (double)iters;
    return 0;
}

}  // extern "C"
