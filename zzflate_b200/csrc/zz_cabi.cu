// C-ABI layer (include/zzgpu.h): per-device context, scratch, staging and the kernel pipeline.
// Replaces WriteDeflateStream (zzflate/zzflate.cpp:81-156) and the checksum calls of
// AppendChecksum (zzflate.cpp:170-192) for the host driver in zz_host.cpp.
#include "../../include/zzgpu.h"
#include "zz_kernels.cuh"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cstring>
#include <cstdio>
#include <algorithm>

namespace {

using namespace zz;

constexpr uint32_t kMaxSlots = 16384;       // chunks per batch (scratch is sized for one batch)
constexpr int kMaxDevices = 16;
constexpr int kMaxLanes = 4;
uint64_t g_pieceFirst = 888, g_pieceMid = 1332, g_pieceLast = 888;   // piece schedule of host-buffer calls, in chunks
int g_optLanes = 3;                               // host-buffer calls: kernel streams that consecutive pieces alternate on ("lanes" option)
size_t g_segBytes = (size_t)2 << 30;             // host-buffer calls: input bytes staged on the device at a time ("segment_mib" option)
constexpr size_t kPipelineMin = (size_t)32 << 20;   // host inputs from this size on are cut into overlapping pieces
constexpr size_t kDefaultSlice = 1000000;   // outputbitstream.h:183

thread_local std::string t_lastError;
thread_local int t_device = -1;
thread_local long long t_sinkPieces = 0, t_sinkFirstH2dDone = 0;

struct Ctx {
    int device = -1;
    bool ready = false;
    // one call at a time per device; a held stream (zzgpu_deflate_hold) keeps the context until the fetch
    std::mutex mu;
    std::condition_variable cv;
    bool busy = false;
    bool held = false;
    std::thread::id holder;
    size_t heldLen = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t extra[kMaxLanes - 1] = {};  // host-buffer path: piece p runs on lane p % lanes (lane 0 is `stream`)
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    uint32_t slots = 0;
    uint32_t slotChunk = 0;
    uint16_t* cand = nullptr; uint8_t* info = nullptr; uint32_t* tokA = nullptr; uint16_t* tokD = nullptr; uint32_t* hist = nullptr;
    ChunkCodes* codes = nullptr; ChunkState* state = nullptr;
    uint64_t* total = nullptr;              // device [4]
    uint64_t* hTotal = nullptr;             // pinned [4]
    uint32_t* ck = nullptr; size_t ckCap = 0;
    uint32_t* hCk = nullptr; size_t hCkCap = 0;
    uint8_t* dIn = nullptr; size_t dInCap = 0;
    uint8_t* dOut = nullptr; size_t dOutCap = 0;
    cudaStream_t copyIn = nullptr, copyOut = nullptr;     // host-buffer path: H2D / D2H overlap the kernels
    uint8_t* stageIn[2] = { nullptr, nullptr };           // pinned blocks between pageable caller memory and the device
    uint8_t* stageOut[2] = { nullptr, nullptr };
    cudaEvent_t stageInEv[2] = { nullptr, nullptr }, stageOutEv[2] = { nullptr, nullptr };
    std::vector<cudaEvent_t> pieceEv;
    uint64_t* hPiece = nullptr; size_t hPieceCap = 0;     // pinned: running totals after each piece
    uint64_t* dPiece = nullptr; size_t dPieceCap = 0;     // device: the totals as K-OFFS of each piece left them
    std::vector<cudaEvent_t> stageEv;       // pool of events bracketing each stage launch
    std::vector<int> stageOf;               // stage id of the interval that ENDS at event i (-1: start marker)
    size_t stageUsed = 0;
};

Ctx g_ctx[kMaxDevices];
std::mutex g_mu;

int fail(int status, const char* what, cudaError_t e = cudaSuccess)
{
    t_lastError = what;
    if (e != cudaSuccess) { t_lastError += ": "; t_lastError += cudaGetErrorString(e); }
    return status;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ZZGPU_E_CUDA, #call, e_); } while (0)

// Exclusive use of a device context for the duration of one call (or until the fetch of a held stream).
class Lease {
public:
    explicit Lease(Ctx& c) : c_(c)
    {
        std::unique_lock<std::mutex> lk(c.mu);
        const std::thread::id me = std::this_thread::get_id();
        if (c.busy && c.held && c.holder == me) { resumed_ = true; return; }       // the holder comes back (fetch / release)
        c.cv.wait(lk, [&] { return !c.busy; });
        c.busy = true; c.holder = me; c.held = false;
    }
    ~Lease()
    {
        std::lock_guard<std::mutex> lk(c_.mu);
        if (resumed_ && !consume_) return;          // the hold goes on (the call was refused)
        if (keep_) { c_.held = true; return; }
        c_.busy = false; c_.held = false;
        c_.cv.notify_all();
    }
    bool resumed() const { return resumed_; }       // this thread already held a stream on the context
    void keep() { keep_ = true; }                   // the call leaves a held stream behind
    void consume() { consume_ = true; }             // the call ends this thread's hold
private:
    Ctx& c_;
    bool resumed_ = false, keep_ = false, consume_ = false;
};

void freeScratch(Ctx& c)
{
    cudaFree(c.cand); cudaFree(c.info); cudaFree(c.tokA); cudaFree(c.tokD); cudaFree(c.hist); cudaFree(c.codes); cudaFree(c.state);
    c.cand = nullptr; c.info = nullptr; c.tokA = nullptr; c.tokD = nullptr; c.hist = nullptr; c.codes = nullptr; c.state = nullptr;
    c.slots = 0; c.slotChunk = 0;
}

// Releases everything a context owns and returns it to the default-constructed state (shutdown, failed set-up).
void destroyCtx(Ctx& c)
{
    if (c.device >= 0) cudaSetDevice(c.device);
    if (c.stream) cudaStreamSynchronize(c.stream);
    for (auto& x : c.extra) if (x) cudaStreamSynchronize(x);
    if (c.copyIn) cudaStreamSynchronize(c.copyIn);
    if (c.copyOut) cudaStreamSynchronize(c.copyOut);
    freeScratch(c);
    cudaFree(c.total); cudaFreeHost(c.hTotal); cudaFree(c.ck); cudaFreeHost(c.hCk); cudaFree(c.dIn); cudaFree(c.dOut);
    cudaFreeHost(c.hPiece); cudaFree(c.dPiece);
    for (auto& e : c.pieceEv) cudaEventDestroy(e);
    for (auto& e : c.stageEv) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (c.stageIn[i]) cudaFreeHost(c.stageIn[i]);
        if (c.stageOut[i]) cudaFreeHost(c.stageOut[i]);
        if (c.stageInEv[i]) cudaEventDestroy(c.stageInEv[i]);
        if (c.stageOutEv[i]) cudaEventDestroy(c.stageOutEv[i]);
    }
    if (c.copyIn) cudaStreamDestroy(c.copyIn);
    if (c.copyOut) cudaStreamDestroy(c.copyOut);
    for (auto& x : c.extra) if (x) cudaStreamDestroy(x);
    for (auto& e : c.ev) if (e) cudaEventDestroy(e);
    if (c.stream) cudaStreamDestroy(c.stream);
    (void)cudaGetLastError();
    c.device = -1; c.ready = false; c.heldLen = 0;
    c.stream = nullptr; for (auto& x : c.extra) x = nullptr; for (auto& e : c.ev) e = nullptr;
    c.total = nullptr; c.hTotal = nullptr; c.ck = nullptr; c.ckCap = 0; c.hCk = nullptr; c.hCkCap = 0;
    c.dIn = nullptr; c.dInCap = 0; c.dOut = nullptr; c.dOutCap = 0; c.copyIn = nullptr; c.copyOut = nullptr;
    for (int i = 0; i < 2; ++i) { c.stageIn[i] = c.stageOut[i] = nullptr; c.stageInEv[i] = c.stageOutEv[i] = nullptr; }
    c.pieceEv.clear(); c.hPiece = nullptr; c.hPieceCap = 0; c.dPiece = nullptr; c.dPieceCap = 0;
    c.stageEv.clear(); c.stageOf.clear(); c.stageUsed = 0;
}

int setUpCtx(Ctx& c, int dev)
{
    CK(cudaSetDevice(dev));
    c.device = dev;
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (auto& e : c.ev) CK(cudaEventCreate(&e));
    CK(cudaMalloc(&c.total, 4 * sizeof(uint64_t)));
    CK(cudaMallocHost(&c.hTotal, 4 * sizeof(uint64_t)));
    CK(configure_kernels());
    return ZZGPU_OK;
}

int ensureCtx(Ctx*& out)
{
    int dev = t_device;
    if (dev < 0) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) return fail(ZZGPU_E_NO_DEVICE, "no CUDA device", e);
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        t_device = dev;
    }
    if (dev >= kMaxDevices) return fail(ZZGPU_E_ARG, "device index too large");
    Ctx& c = g_ctx[dev];
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!c.ready) {
            const int rc = setUpCtx(c, dev);
            if (rc) { const std::string keep = t_lastError; destroyCtx(c); t_lastError = keep; return rc; }
            c.ready = true;
        }
    }
    CK(cudaSetDevice(dev));
    out = &c;
    return ZZGPU_OK;
}

int ensureScratch(Ctx& c, uint32_t slots, uint32_t chunk)
{
    if (c.slots >= slots && c.slotChunk >= chunk) return ZZGPU_OK;
    freeScratch(c);
    cudaError_t e = cudaSuccess;
    auto alloc = [&](auto& p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(&p, bytes); };
    alloc(c.cand, (size_t)slots * chunk * sizeof(uint16_t));
    alloc(c.info, (size_t)slots * chunk);
    alloc(c.tokA, (size_t)slots * kMaxTokens * sizeof(uint32_t));
    alloc(c.tokD, (size_t)slots * kMaxTokens * sizeof(uint16_t));
    alloc(c.hist, (size_t)slots * kHistStride * sizeof(uint32_t));
    alloc(c.codes, (size_t)slots * sizeof(ChunkCodes));
    alloc(c.state, (size_t)slots * sizeof(ChunkState));
    if (e != cudaSuccess) { freeScratch(c); (void)cudaGetLastError(); return fail(ZZGPU_E_NOMEM, "scratch allocation failed", e); }
    c.slots = slots; c.slotChunk = chunk;
    return ZZGPU_OK;
}

template <class T>
int ensureBuf(T*& p, size_t& cap, size_t need, bool pinned = false)
{
    if (cap >= need && p) return ZZGPU_OK;
    if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; cap = 0; }
    size_t bytes = std::max<size_t>(need, 256) * sizeof(T);
    cudaError_t e = pinned ? cudaMallocHost(&p, bytes) : cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; (void)cudaGetLastError(); return fail(ZZGPU_E_NOMEM, "allocation failed", e); }
    cap = std::max<size_t>(need, 256);
    return ZZGPU_OK;
}

bool validParams(int level, uint32_t chunk, uint32_t dict)
{
    return level >= 0 && level <= 3 && chunk >= 1024 && chunk <= ZZGPU_MAX_CHUNK && (chunk % 32) == 0 &&
           dict <= ZZGPU_MAX_DICT;
}

int markStage(Ctx& c, int stage, cudaStream_t st = nullptr)
{
    if (!st) st = c.stream;
    if (c.stageUsed == c.stageEv.size()) {
        cudaEvent_t e; CK(cudaEventCreate(&e));
        c.stageEv.push_back(e); c.stageOf.push_back(-1);
    }
    c.stageOf[c.stageUsed] = stage;
    CK(cudaEventRecord(c.stageEv[c.stageUsed], st));
    c.stageUsed++;
    return ZZGPU_OK;
}

void collectStages(Ctx& c, zzgpu_stats* stats)
{
    if (!stats) return;
    for (size_t i = 1; i < c.stageUsed; ++i) {
        const int st = c.stageOf[i];
        if (st < 0) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c.stageEv[i - 1], c.stageEv[i]) == cudaSuccess) { stats->stage_ms[st] += ms; stats->stage_launches[st]++; }
    }
    (void)cudaGetLastError();                    // timing queries must never poison the next call
}

// Parallel memcpy between pageable caller memory and the pinned staging blocks (a single host thread moves
// ~10 GB/s, less than the PCIe link).  Workers are shared by all contexts of the process.
class CopyPool {
public:
    static CopyPool& get() { static CopyPool p; return p; }
    void copy(void* dst, const void* src, size_t n)
    {
        const size_t part = (size_t)2 << 20;
        if (n <= 2 * part || workers.empty()) { memcpy(dst, src, n); return; }
        std::atomic<size_t> left{ (n + part - 1) / part };
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t off = 0; off < n; off += part)
                q.push_back(Task{ (char*)dst + off, (const char*)src + off, std::min(part, n - off), &left });
        }
        cv.notify_all();
        for (;;) {                                   // the caller helps, then waits for stragglers
            Task t;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (q.empty()) break;
                t = q.front(); q.pop_front();
            }
            memcpy(t.d, t.s, t.n); t.left->fetch_sub(1);
        }
        while (left.load() != 0) std::this_thread::yield();
    }
private:
    struct Task { char* d; const char* s; size_t n; std::atomic<size_t>* left; };
    CopyPool()
    {
        unsigned k = std::thread::hardware_concurrency();
        k = k > 16 ? 7 : (k > 3 ? k / 2 - 1 : 0);
        for (unsigned i = 0; i < k; ++i) workers.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& w : workers) w.join();
    }
    void run()
    {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [this] { return stop || !q.empty(); });
                if (stop && q.empty()) return;
                t = q.front(); q.pop_front();
            }
            memcpy(t.d, t.s, t.n); t.left->fetch_sub(1);
        }
    }
    std::vector<std::thread> workers; std::mutex mu; std::condition_variable cv; std::deque<Task> q; bool stop = false;
};

bool isPinnedHost(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

constexpr size_t kStageBlock = (size_t)32 << 20;    // pinned staging block for pageable caller buffers

int preparePipeline(Ctx& c, size_t n, uint32_t chunk, int wantCk)
{
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    const uint32_t slots = (uint32_t)std::min<uint64_t>(nchunks, kMaxSlots);
    int rc = ensureScratch(c, slots, chunk); if (rc) return rc;
    if (wantCk) { rc = ensureBuf(c.ck, c.ckCap, 2 * nchunks); if (rc) return rc; }
    CK(cudaMemsetAsync(c.total, 0, 4 * sizeof(uint64_t), c.stream));
    c.stageUsed = 0;
    return ZZGPU_OK;
}

thread_local int t_mode = 0;                // zzgpu_deflate_mode: mode of the call in progress on this thread

Job makeJob(Ctx& c, const uint8_t* d_src, size_t n, size_t history, int final, uint8_t* d_dst, size_t cap, int level,
            uint32_t chunk, uint32_t dict, int wantCk, uint64_t first, uint32_t count, uint32_t slotBase = 0)
{
    Job job{};
    job.mode = t_mode;
    job.src = d_src; job.n = n; job.history = history; job.chunk = chunk; job.dict = dict;
    job.first_chunk = first; job.nchunks = count;
    job.final_stream = final; job.level = level; job.want_checksums = wantCk;
    const size_t sb = slotBase;                  // the job's rows of the per-chunk scratch arrays start here
    job.cand = c.cand + sb * chunk; job.info = c.info + sb * chunk; job.tokA = c.tokA + sb * kMaxTokens; job.tokD = c.tokD + sb * kMaxTokens;
    job.hist = c.hist + sb * kHistStride; job.codes = c.codes + sb; job.state = c.state + sb;
    job.dst = d_dst; job.cap = cap; job.total = c.total; job.ck = c.ck;
    return job;
}

// A piece of the host-buffer path that runs beside its neighbours: its own stream and half of the scratch rows.  K-OFFS
// carries the running output offset from piece to piece, so it waits for the previous piece's K-OFFS and leaves its own
// totals in `snapshot` for the host (the live totals move on with the next piece).
struct Lane {
    cudaStream_t st; uint32_t slotBase, slotCap;
    cudaEvent_t waitOffs, doneOffs;             // previous piece's K-OFFS done (or null) / this piece's
    uint64_t* snapshot;                         // device [4]
};

// Kernel pipeline over chunks [firstChunk, lastChunk) of the call (geometry is always that of the whole call):
// batches of up to 16 384 chunks, kernels back to back on one stream.
int runChunks(Ctx& c, const uint8_t* d_src, size_t n, size_t history, int final, uint8_t* d_dst, size_t cap,
              int level, uint32_t chunk, uint32_t dict, int wantCk, uint64_t firstChunk, uint64_t lastChunk, uint64_t& launches,
              const Lane* lane = nullptr)
{
    int rc;
    const uint32_t slotCap = lane ? lane->slotCap : c.slots, slotBase = lane ? lane->slotBase : 0;
    cudaStream_t st = lane ? lane->st : c.stream;
    for (uint64_t first = firstChunk; first < lastChunk; first += slotCap) {
        const Job job = makeJob(c, d_src, n, history, final, d_dst, cap, level, chunk, dict, wantCk, first,
                                (uint32_t)std::min<uint64_t>(slotCap, lastChunk - first), slotBase);
        const bool firstBatch = first == firstChunk, lastBatch = first + slotCap >= lastChunk;
        rc = markStage(c, -1, st); if (rc) return rc;
        if (level == 0) {                                   // every size is known beforehand: one kernel does it all
            launches += launch_stored(job, st); rc = markStage(c, ZZGPU_STAGE_EMIT, st); if (rc) return rc;
            continue;
        }
        if (level >= 2) {
            launches += launch_candidates(job, st); rc = markStage(c, ZZGPU_STAGE_CAND, st); if (rc) return rc;
            launches += launch_lz(job, st); rc = markStage(c, ZZGPU_STAGE_LZ, st); if (rc) return rc;
        }
        if (level == 1) {
            launches += launch_fixed(job, st); rc = markStage(c, ZZGPU_STAGE_FIXED, st); if (rc) return rc;
        } else {
            launches += launch_huffman(job, st); rc = markStage(c, ZZGPU_STAGE_HUFF, st); if (rc) return rc;
        }
        if (lane && firstBatch && lane->waitOffs) { CK(cudaStreamWaitEvent(st, lane->waitOffs, 0)); rc = markStage(c, -1, st); if (rc) return rc; }
        launches += launch_offsets(job, st); rc = markStage(c, ZZGPU_STAGE_OFFS, st); if (rc) return rc;
        if (lane && lastBatch) {
            CK(cudaMemcpyAsync(lane->snapshot, c.total, 4 * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
            CK(cudaEventRecord(lane->doneOffs, st));
            rc = markStage(c, -1, st); if (rc) return rc;
        }
        if (level == 1) { launches += launch_gather(job, st); rc = markStage(c, ZZGPU_STAGE_GATHER, st); if (rc) return rc; }
        else { launches += launch_emit(job, st); rc = markStage(c, ZZGPU_STAGE_EMIT, st); if (rc) return rc; }
        if (wantCk && !(level >= 2 && wantCk == 1)) {       // levels 2/3 with Adler-32 only (zlib): the sum is taken inside K-LZ
            launches += launch_checksums(job, st); rc = markStage(c, ZZGPU_STAGE_CKSUM, st); if (rc) return rc;
        }
    }
    CK(cudaGetLastError());
    return ZZGPU_OK;
}

// Runs the device pipeline over all chunks of the call.  d_src points at stream position 0 of the call in
// device memory (history bytes before it), d_dst receives the stream.
int runPipeline(Ctx& c, const uint8_t* d_src, size_t n, size_t history, int final, uint8_t* d_dst, size_t cap,
                int level, uint32_t chunk, uint32_t dict, int wantCk, uint64_t& launches)
{
    int rc = preparePipeline(c, n, chunk, wantCk); if (rc) return rc;
    return runChunks(c, d_src, n, history, final, d_dst, cap, level, chunk, dict, wantCk, 0, (n + chunk - 1) / chunk, launches);
}

// Piece schedule of the host-buffer path: small pieces first (the kernels start after one short H2D), large in
// the middle (full occupancy, one dictionary-priming pass per long run of chunks), small at the end (short drain).
std::vector<uint64_t> pieceSchedule(uint64_t nchunks, uint32_t chunk)
{
    // Every piece costs one launch of each kernel (the Huffman kernel alone is ~0.4 ms however few chunks it gets), and the
    // kernels of a piece start when its H2D is complete.  The link moves a chunk in about the time the kernels need for it,
    // so what counts is the sum of the pieces' fixed costs plus the first H2D and the last piece's kernels and D2H: a
    // short first piece, equal middle pieces, a short last one.  Sizes are multiples of 444 chunks = one wave of K-LZ
    // (148 SMs x 3 CTAs); 888 is a whole number of waves of K-CAND and K-EMIT too.
    std::vector<uint64_t> ends;
    const uint64_t first = g_pieceFirst, mid = g_pieceMid, last = g_pieceLast, unit = 444;
    if (nchunks * chunk < kPipelineMin || nchunks < first + last + unit) { ends.push_back(nchunks); return ends; }
    ends.push_back(first);
    const uint64_t middle = nchunks - first - last;
    const uint64_t k = std::max<uint64_t>(1, (middle + mid / 2) / mid);            // number of middle pieces
    uint64_t pos = first;
    for (uint64_t i = 0; i < k; ++i) {
        uint64_t take = i + 1 == k ? nchunks - last - pos : ((middle / k + unit / 2) / unit) * unit;
        if (take == 0) take = unit;
        if (pos + take > nchunks - last) take = nchunks - last - pos;
        if (take == 0) break;
        pos += take; ends.push_back(pos);
    }
    if (pos < nchunks) ends.push_back(nchunks);
    return ends;
}

// Where the stream of a host-buffer call goes.
struct HostOut {
    enum Kind { Direct, Sink, Hold } kind = Direct;
    uint8_t* dst = nullptr; size_t cap = 0; bool pinned = false;        // Direct: caller's buffer
    zzgpu_sink_fn sink = nullptr; void* user = nullptr; size_t slice = kDefaultSlice;   // Sink: slices in order
    size_t written = 0;         // bytes of the stream delivered (Direct, Sink) or held on the device (Hold) so far
    size_t d2h = 0;
    bool firstSlice = true;
};

int ensureStaging(Ctx& c)
{
    if (!c.copyIn) { CK(cudaStreamCreateWithFlags(&c.copyIn, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c.copyOut, cudaStreamNonBlocking)); }
    for (auto& x : c.extra) if (!x) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        if (!c.stageIn[i]) CK(cudaMallocHost(&c.stageIn[i], kStageBlock));
        if (!c.stageOut[i]) CK(cudaMallocHost(&c.stageOut[i], kStageBlock));
        if (!c.stageInEv[i]) CK(cudaEventCreateWithFlags(&c.stageInEv[i], cudaEventDisableTiming));
        if (!c.stageOutEv[i]) CK(cudaEventCreateWithFlags(&c.stageOutEv[i], cudaEventDisableTiming));
    }
    return ZZGPU_OK;
}

// Moves d_from[0, len) (device) to the call's destination.  Direct + pinned: one asynchronous copy on the copy-out
// stream.  Direct + pageable / Sink: through the two pinned staging blocks, the D2H of block k+1 running while block k
// is copied out (or handed to the sink in slices).  `probe`: the segment's H2D-complete events (diagnostic counter).
struct PieceProbe { const std::vector<cudaEvent_t>* ev; size_t issued, total; };

int deliver(Ctx& c, HostOut& out, const uint8_t* d_from, size_t len, const PieceProbe* probe)
{
    if (len == 0) return ZZGPU_OK;
    if (out.kind == HostOut::Hold) { out.written += len; return ZZGPU_OK; }
    if (out.kind == HostOut::Direct && out.pinned) {
        CK(cudaMemcpyAsync(out.dst + out.written, d_from, len, cudaMemcpyDeviceToHost, c.copyOut));
        out.written += len; out.d2h += len;
        return ZZGPU_OK;
    }
    size_t blocks = 0;
    size_t pendLen = 0; int pendSb = 0;
    for (size_t off = 0; off < len || pendLen; ) {
        size_t cur = 0; int sb = 0;
        if (off < len) {
            cur = std::min(kStageBlock, len - off); sb = (int)(blocks++ & 1);
            CK(cudaMemcpyAsync(c.stageOut[sb], d_from + off, cur, cudaMemcpyDeviceToHost, c.copyOut));
            CK(cudaEventRecord(c.stageOutEv[sb], c.copyOut));
        }
        if (pendLen) {
            CK(cudaEventSynchronize(c.stageOutEv[pendSb]));
            if (out.kind == HostOut::Direct) {
                CopyPool::get().copy(out.dst + out.written, c.stageOut[pendSb], pendLen);
            } else {
                if (out.firstSlice) {
                    out.firstSlice = false;
                    long long doneH2d = 0;
                    if (probe) for (size_t q = 0; q < probe->issued; ++q) if (cudaEventQuery((*probe->ev)[3 * q]) == cudaSuccess) ++doneH2d;
                    (void)cudaGetLastError();
                    t_sinkFirstH2dDone = doneH2d; t_sinkPieces = probe ? (long long)probe->total : 0;
                }
                for (size_t s = 0; s < pendLen; s += out.slice) out.sink(c.stageOut[pendSb] + s, std::min(out.slice, pendLen - s), out.user);
            }
            out.written += pendLen; out.d2h += pendLen;
        }
        pendLen = cur; pendSb = sb;
        off += cur;
    }
    return ZZGPU_OK;
}

// One segment (<= g_segBytes of input) of a host-buffer call: H2D copies, kernels and D2H copies of successive pieces
// overlap on three streams.  `histSrc` holds the `hist` bytes of stream before src.  d_out/d_cap: device buffer that
// receives the segment's stream.
int runHostSegment(Ctx& c, const uint8_t* src, size_t n, const uint8_t* histSrc, size_t hist, int final,
                   HostOut& out, uint8_t* d_out, size_t d_cap, int level, uint32_t chunk, uint32_t dict, int wantCk,
                   uint64_t& launches)
{
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    int rc = ensureBuf(c.dIn, c.dInCap, hist + n + 64); if (rc) return rc;
    rc = ensureStaging(c); if (rc) return rc;
    const bool srcPinned = isPinnedHost(src);
    const std::vector<uint64_t> ends = pieceSchedule(nchunks, chunk);
    const size_t np = ends.size();
    while (c.pieceEv.size() < 3 * np) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c.pieceEv.push_back(e); }
    rc = ensureBuf(c.hPiece, c.hPieceCap, 4 * np, true); if (rc) return rc;
    rc = ensureBuf(c.dPiece, c.dPieceCap, 4 * np); if (rc) return rc;
    rc = preparePipeline(c, n, chunk, wantCk); if (rc) return rc;
    // Two lanes: consecutive pieces alternate between two streams and the two halves of the scratch rows, so that the
    // latency-bound kernels of one piece (K-HUFF, K-OFFS, the tails of the others) run beside the next piece's K-CAND / K-LZ.
    uint64_t maxPiece = 0;
    for (size_t p = 0; p < np; ++p) maxPiece = std::max(maxPiece, ends[p] - (p ? ends[p - 1] : 0));
    int lanes = g_optLanes;
    while (lanes > 1 && (uint64_t)lanes * maxPiece > c.slots) --lanes;
    if (np < 2 || level == 0) lanes = 1;
    const bool twoLanes = lanes > 1;
    // the copy streams must not touch the staging buffers before earlier work on the main stream is done with them
    CK(cudaEventRecord(c.ev[0], c.stream));
    CK(cudaStreamWaitEvent(c.copyIn, c.ev[0], 0));
    CK(cudaStreamWaitEvent(c.copyOut, c.ev[0], 0));
    for (auto& x : c.extra) CK(cudaStreamWaitEvent(x, c.ev[0], 0));
    if (hist) CK(cudaMemcpyAsync(c.dIn, histSrc, hist, cudaMemcpyHostToDevice, c.copyIn));
    const uint8_t* d_src = c.dIn + hist;
    size_t inBlocks = 0, done = 0, nextDrain = 0, issued = 0;
    uint64_t firstChunk = 0;

    auto issue = [&](size_t p) -> int {
        const size_t lo = (size_t)(firstChunk * chunk), hi = (size_t)std::min<uint64_t>(n, ends[p] * chunk);
        if (srcPinned) {
            CK(cudaMemcpyAsync(c.dIn + hist + lo, src + lo, hi - lo, cudaMemcpyHostToDevice, c.copyIn));
        } else {
            for (size_t off = lo; off < hi; off += kStageBlock, ++inBlocks) {
                const size_t len = std::min(kStageBlock, hi - off);
                const int sb = (int)(inBlocks & 1);
                CK(cudaEventSynchronize(c.stageInEv[sb]));                   // the block's previous H2D has drained
                CopyPool::get().copy(c.stageIn[sb], src + off, len);
                CK(cudaMemcpyAsync(c.dIn + hist + off, c.stageIn[sb], len, cudaMemcpyHostToDevice, c.copyIn));
                CK(cudaEventRecord(c.stageInEv[sb], c.copyIn));
            }
        }
        CK(cudaEventRecord(c.pieceEv[3 * p], c.copyIn));
        const uint32_t ln = (uint32_t)(p % (size_t)lanes);
        cudaStream_t st = ln ? c.extra[ln - 1] : c.stream;
        CK(cudaStreamWaitEvent(st, c.pieceEv[3 * p], 0));
        int r;
        if (twoLanes) {
            const Lane lane = { st, ln * (c.slots / lanes), c.slots / lanes, p ? c.pieceEv[3 * (p - 1) + 2] : nullptr,
                                c.pieceEv[3 * p + 2], c.dPiece + 4 * p };
            r = runChunks(c, d_src, n, hist, final, d_out, d_cap, level, chunk, dict, wantCk, firstChunk, ends[p], launches, &lane);
            if (r) return r;
            // offset, match and stored counts as this piece's K-OFFS left them; the flags as they are now
            CK(cudaMemcpyAsync(c.hPiece + 4 * p, c.dPiece + 4 * p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(c.hPiece + 4 * p + 1, c.total + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        } else {
            r = runChunks(c, d_src, n, hist, final, d_out, d_cap, level, chunk, dict, wantCk, firstChunk, ends[p], launches);
            if (r) return r;
            CK(cudaMemcpyAsync(c.hPiece + 4 * p, c.total, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        }
        CK(cudaEventRecord(c.pieceEv[3 * p + 1], st));
        firstChunk = ends[p];
        issued = p + 1;
        return ZZGPU_OK;
    };
    // returns 0 when piece p was delivered, 1 when it is not ready yet (non-blocking mode), < 0 = -status
    auto drain = [&](size_t p, bool blocking) -> int {
        if (!blocking) {
            const cudaError_t q = cudaEventQuery(c.pieceEv[3 * p + 1]);
            if (q == cudaErrorNotReady) { (void)cudaGetLastError(); return 1; }
            if (q != cudaSuccess) return -fail(ZZGPU_E_CUDA, "cudaEventQuery", q);
        } else {
            const cudaError_t q = cudaEventSynchronize(c.pieceEv[3 * p + 1]);
            if (q != cudaSuccess) return -fail(ZZGPU_E_CUDA, "cudaEventSynchronize", q);
        }
        const uint64_t upto = c.hPiece[4 * p], flags = c.hPiece[4 * p + 1];
        if (flags & 1) return -fail(ZZGPU_E_CAPACITY, "destination too small");
        if (flags & ~1ull) return -fail(ZZGPU_E_CUDA, "internal consistency check failed (emit size mismatch)");
        if (upto > d_cap) return -fail(ZZGPU_E_CAPACITY, "destination too small");
        if (upto > done) {
            const PieceProbe probe = { &c.pieceEv, issued, np };
            const int r = deliver(c, out, d_out + done, (size_t)upto - done, &probe);
            if (r) return -r;
        }
        done = (size_t)upto;
        return 0;
    };

    for (size_t p = 0; p < np; ++p) {
        rc = issue(p); if (rc) return rc;
        while (nextDrain < p) {                  // hand over what has finished meanwhile (pageable input: the issue above took a while)
            const int r = drain(nextDrain, false);
            if (r < 0) return -r;
            if (r) break;
            ++nextDrain;
        }
    }
    for (; nextDrain < np; ++nextDrain) { const int r = drain(nextDrain, true); if (r < 0) return -r; }
    CK(cudaStreamSynchronize(c.copyOut));        // the device output buffer is reused by the next segment / call
    for (int i = 0; i < 4; ++i) c.hTotal[i] = c.hPiece[4 * (np - 1) + i];
    return ZZGPU_OK;
}

int foldChecksums(Ctx& c, size_t n, uint32_t chunk, int wantCk, uint32_t* adler0, uint32_t* crc)
{
    if (!wantCk) return ZZGPU_OK;
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    int rc = ensureBuf(c.hCk, c.hCkCap, 2 * nchunks, true); if (rc) return rc;
    CK(cudaMemcpyAsync(c.hCk, c.ck, 2 * nchunks * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    uint32_t a = 0, r = 0;
    const uint32_t shiftFull = crc32_shift_operator(chunk);          // x^(8*chunk): same for every full chunk
    for (uint64_t k = 0; k < nchunks; ++k) {
        const size_t len = (size_t)std::min<uint64_t>(chunk, n - k * chunk);
        a = adler32_combine(a, c.hCk[2 * k], len);
        r = len == chunk ? crc32_apply_shift(r, shiftFull) ^ c.hCk[2 * k + 1] : crc32_combine(r, c.hCk[2 * k + 1], len);
    }
    if (adler0) *adler0 = a;
    if (crc) *crc = r;
    return ZZGPU_OK;
}

void drainStreams(Ctx& c)
{
    if (c.copyIn) cudaStreamSynchronize(c.copyIn);
    if (c.stream) cudaStreamSynchronize(c.stream);
    for (auto& x : c.extra) if (x) cudaStreamSynchronize(x);
    if (c.copyOut) cudaStreamSynchronize(c.copyOut);
    (void)cudaGetLastError();
}

// Host-buffer call: segments of <= g_segBytes, each pipelined in pieces.  Checksums of the segments are folded here.
int deflateHost(Ctx& c, const uint8_t* src, size_t n, const uint8_t* histSrc, size_t hist, int final, HostOut& out,
                int level, uint32_t chunk, uint32_t dict, int wantCk, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats)
{
    const auto t0 = std::chrono::steady_clock::now();
    const size_t segBytes = std::max<size_t>(1, g_segBytes / chunk) * chunk;
    if (out.kind == HostOut::Hold) { int rc = ensureBuf(c.dOut, c.dOutCap, zzgpu_bound(n, level, chunk) + 64); if (rc) return rc; }
    uint32_t a = 0, r = 0;
    uint64_t launches = 0, matches = 0, stored = 0;
    size_t h2d = 0;
    for (size_t off = 0; off < n; off += segBytes) {
        const size_t len = std::min(segBytes, n - off);
        const int segFinal = final && off + len == n;
        const size_t h = off == 0 ? hist : std::min<size_t>(off + hist, (size_t)dict + kPreExtra);
        const uint8_t* hs = off == 0 ? histSrc : src + off - h;
        uint8_t* d_out; size_t d_cap;
        const size_t bound = zzgpu_bound(len, level, chunk);
        if (out.kind == HostOut::Hold) { d_out = c.dOut + out.written; d_cap = bound; }
        else {
            d_cap = out.kind == HostOut::Direct ? std::min(out.cap - out.written, bound) : bound;
            int rc = ensureBuf(c.dOut, c.dOutCap, d_cap + 64); if (rc) return rc;
            d_out = c.dOut;
        }
        int rc = runHostSegment(c, src + off, len, hs, h, segFinal, out, d_out, d_cap, level, chunk, dict, wantCk, launches);
        if (rc) { drainStreams(c); return rc; }
        uint32_t sa = 0, sr = 0;
        rc = foldChecksums(c, len, chunk, wantCk, &sa, &sr); if (rc) return rc;
        a = adler32_combine(a, sa, len); r = crc32_combine(r, sr, len);
        matches += c.hTotal[2]; stored += c.hTotal[3];
        h2d += h + len;
        collectStages(c, stats);
    }
    if (adler0) *adler0 = a;
    if (crc) *crc = r;
    if (stats) {
        stats->chunks = (n + chunk - 1) / chunk;
        stats->matches = matches; stats->stored_chunks = stored; stats->kernel_launches = launches;
        stats->device_ms = 0;                        // kernels interleave with copies: the stages' sum
        for (int i = 0; i < ZZGPU_NSTAGES; ++i) stats->device_ms += stats->stage_ms[i];
        stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        stats->h2d_bytes = h2d; stats->d2h_bytes = out.d2h;
    }
    return ZZGPU_OK;
}

const uint8_t kEmptyFinalStored[5] = { 0x01, 0x00, 0x00, 0xFF, 0xFF };     // R7: empty input still gets one final block

struct CallArgs {
    const uint8_t* src; size_t n; const uint8_t* histSrc; size_t history; int final; int src_mem;
    uint8_t* dst; size_t cap; int dst_mem;
    int level; uint32_t chunk; uint32_t dict; int want;
    HostOut::Kind kind; zzgpu_sink_fn sink; void* user; size_t slice;
};

int deflateCall(const CallArgs& a, size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats)
{
    const uint32_t chunk = a.chunk ? a.chunk : ZZGPU_DEFAULT_CHUNK;
    if (!validParams(a.level, chunk, a.dict) || !out_len || (!a.src && a.n)) return fail(ZZGPU_E_ARG, "invalid argument");
    if (a.kind == HostOut::Direct && !a.dst && a.cap) return fail(ZZGPU_E_ARG, "invalid argument");
    if (a.kind == HostOut::Sink && !a.sink) return fail(ZZGPU_E_ARG, "invalid argument");
    Ctx* cp = nullptr;
    int rc = ensureCtx(cp); if (rc) return rc;
    Ctx& c = *cp;
    Lease lease(c);
    if (lease.resumed()) return fail(ZZGPU_E_ARG, "this thread holds a stream on the device: fetch or release it first");
    if (stats) memset(stats, 0, sizeof *stats);
    if (adler0) *adler0 = 0;
    if (crc) *crc = 0;
    *out_len = 0;
    const bool hostBoth = a.src_mem == ZZGPU_MEM_HOST && a.dst_mem == ZZGPU_MEM_HOST;

    if (a.n == 0) {
        if (a.final) {
            if (a.kind == HostOut::Sink) { a.sink(kEmptyFinalStored, 5, a.user); }
            else if (a.kind == HostOut::Hold) {
                rc = ensureBuf(c.dOut, c.dOutCap, 64); if (rc) return rc;
                CK(cudaMemcpyAsync(c.dOut, kEmptyFinalStored, 5, cudaMemcpyHostToDevice, c.stream)); CK(cudaStreamSynchronize(c.stream));
            } else {
                if (a.cap < 5) return fail(ZZGPU_E_CAPACITY, "destination too small");
                if (a.dst_mem == ZZGPU_MEM_HOST) memcpy(a.dst, kEmptyFinalStored, 5);
                else { CK(cudaMemcpyAsync(a.dst, kEmptyFinalStored, 5, cudaMemcpyHostToDevice, c.stream)); CK(cudaStreamSynchronize(c.stream)); }
            }
            *out_len = 5;
        }
        if (a.kind == HostOut::Hold) { c.heldLen = *out_len; lease.keep(); }
        return ZZGPU_OK;
    }

    const size_t hist = std::min<size_t>(a.history, (size_t)a.dict + kPreExtra);    // bytes the kernels may look at
    if (hostBoth) {
        HostOut out;
        out.kind = a.kind; out.dst = a.dst; out.cap = a.cap; out.pinned = a.kind == HostOut::Direct && isPinnedHost(a.dst);
        out.sink = a.sink; out.user = a.user; out.slice = a.slice ? a.slice : kDefaultSlice;
        const uint8_t* hs = a.histSrc ? a.histSrc + (a.history - hist) : a.src - hist;
        rc = deflateHost(c, a.src, a.n, hs, hist, a.final, out, a.level, chunk, a.dict, a.want, adler0, crc, stats);
        if (rc) return rc;
        *out_len = out.written;
        if (a.kind == HostOut::Hold) { c.heldLen = out.written; lease.keep(); }
        return ZZGPU_OK;
    }

    // device-resident (or mixed) buffers: the caller's device memory is used in place
    uint64_t launches = 0;
    size_t h2d = 0, d2h = 0;
    CK(cudaEventRecord(c.ev[0], c.stream));
    const uint8_t* d_src = a.src;
    if (a.src_mem == ZZGPU_MEM_HOST) {
        rc = ensureBuf(c.dIn, c.dInCap, hist + a.n + 64); if (rc) return rc;
        CK(cudaMemcpyAsync(c.dIn, a.src - hist, hist + a.n, cudaMemcpyHostToDevice, c.stream));
        d_src = c.dIn + hist;
        h2d = hist + a.n;
    }
    uint8_t* d_dst = a.dst;
    size_t d_cap = a.cap;
    if (a.dst_mem == ZZGPU_MEM_HOST) {
        d_cap = std::min(a.cap, zzgpu_bound(a.n, a.level, chunk));
        rc = ensureBuf(c.dOut, c.dOutCap, d_cap + 64); if (rc) return rc;
        d_dst = c.dOut;
    }
    CK(cudaEventRecord(c.ev[1], c.stream));
    rc = runPipeline(c, d_src, a.n, hist, a.final, d_dst, d_cap, a.level, chunk, a.dict, a.want, launches);
    if (rc) return rc;
    CK(cudaEventRecord(c.ev[2], c.stream));
    CK(cudaMemcpyAsync(c.hTotal, c.total, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    const uint64_t total = c.hTotal[0], flags = c.hTotal[1];
    if (flags & 1) return fail(ZZGPU_E_CAPACITY, "destination too small");
    if (flags & ~1ull) return fail(ZZGPU_E_CUDA, "internal consistency check failed (emit size mismatch)");
    if (total > a.cap) return fail(ZZGPU_E_CAPACITY, "destination too small");
    if (a.dst_mem == ZZGPU_MEM_HOST) {
        CK(cudaMemcpyAsync(a.dst, c.dOut, total, cudaMemcpyDeviceToHost, c.stream));
        d2h = total;
    }
    CK(cudaEventRecord(c.ev[3], c.stream));
    rc = foldChecksums(c, a.n, chunk, a.want, adler0, crc); if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    *out_len = (size_t)total;
    if (stats) {
        stats->chunks = (a.n + chunk - 1) / chunk;
        stats->matches = c.hTotal[2];
        stats->stored_chunks = c.hTotal[3];
        stats->kernel_launches = launches;
        collectStages(c, stats);
        cudaEventElapsedTime(&stats->total_ms, c.ev[0], c.ev[3]);
        cudaEventElapsedTime(&stats->device_ms, c.ev[1], c.ev[2]);
        (void)cudaGetLastError();
        stats->h2d_bytes = h2d; stats->d2h_bytes = d2h;
    }
    return ZZGPU_OK;
}

}  // namespace

extern "C" {

int zzgpu_init(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(ZZGPU_E_NO_DEVICE, "no CUDA device", e);
    if (device < 0 || device >= count) return fail(ZZGPU_E_ARG, "bad device index");
    t_device = device;
    Ctx* c = nullptr;
    return ensureCtx(c);
}

void zzgpu_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& c : g_ctx) {
        if (!c.ready) continue;
        {   // wait for a running call on this context (a stream somebody still holds is dropped with it)
            std::unique_lock<std::mutex> cl(c.mu);
            c.cv.wait(cl, [&] { return !c.busy || c.held; });
            c.busy = true; c.held = false;
        }
        destroyCtx(c);
        {
            std::lock_guard<std::mutex> cl(c.mu);
            c.busy = false; c.held = false;
        }
        c.cv.notify_all();
    }
}

int zzgpu_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return count;
}

const char* zzgpu_strerror(int status)
{
    switch (status) {
    case ZZGPU_OK: return "ok";
    case ZZGPU_E_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case ZZGPU_E_CUDA: return "CUDA error";
    case ZZGPU_E_ARG: return "invalid argument";
    case ZZGPU_E_CAPACITY: return "destination buffer too small";
    case ZZGPU_E_NOMEM: return "out of memory";
    default: return "unknown status";
    }
}

const char* zzgpu_last_error(void) { return t_lastError.c_str(); }

size_t zzgpu_bound(size_t n, int level, uint32_t chunk)
{
    if (chunk == 0) chunk = ZZGPU_DEFAULT_CHUNK;
    const size_t chunks = n ? (n + chunk - 1) / chunk : 1;
    const size_t per = level == 1 ? ((size_t)chunk * 9 + 7) / 8 + 16 : (size_t)chunk + 16;
    return chunks * per;
}

int zzgpu_deflate_ex(const uint8_t* src, size_t n, size_t history, int final, int src_mem,
                     uint8_t* dst, size_t cap, int dst_mem,
                     int level, uint32_t chunk, uint32_t dict, int want_checksums,
                     size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats)
{
    const CallArgs a = { src, n, nullptr, history, final, src_mem, dst, cap, dst_mem, level, chunk, dict, want_checksums,
                         HostOut::Direct, nullptr, nullptr, 0 };
    return deflateCall(a, out_len, adler0, crc, stats);
}

int zzgpu_deflate_mode(const uint8_t* src, size_t n, size_t history, int final, int src_mem,
                       uint8_t* dst, size_t cap, int dst_mem,
                       int level, uint32_t chunk, uint32_t dict, int want_checksums, int mode,
                       size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats)
{
    if (mode != 0 && mode != 1) return fail(ZZGPU_E_ARG, "invalid mode");
    t_mode = mode;
    const int rc = zzgpu_deflate_ex(src, n, history, final, src_mem, dst, cap, dst_mem, level, chunk, dict, want_checksums,
                                    out_len, adler0, crc, stats);
    t_mode = 0;
    return rc;
}

int zzgpu_deflate_hist(const uint8_t* src, size_t n, const uint8_t* hist, size_t hist_len, int final,
                       uint8_t* dst, size_t cap, int level, uint32_t chunk, uint32_t dict, size_t* out_len)
{
    if (hist_len && !hist) return fail(ZZGPU_E_ARG, "invalid argument");
    const CallArgs a = { src, n, hist, hist_len, final, ZZGPU_MEM_HOST, dst, cap, ZZGPU_MEM_HOST, level, chunk, dict, 0,
                         HostOut::Direct, nullptr, nullptr, 0 };
    return deflateCall(a, out_len, nullptr, nullptr, nullptr);
}

int zzgpu_deflate_sink(const uint8_t* src, size_t n, size_t history, int final,
                       int level, uint32_t chunk, uint32_t dict, int want_checksums,
                       zzgpu_sink_fn sink, void* user, size_t slice,
                       size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats)
{
    t_sinkPieces = 0; t_sinkFirstH2dDone = 0;
    const CallArgs a = { src, n, nullptr, history, final, ZZGPU_MEM_HOST, nullptr, 0, ZZGPU_MEM_HOST, level, chunk, dict,
                         want_checksums, HostOut::Sink, sink, user, slice };
    return deflateCall(a, out_len, adler0, crc, stats);
}

int zzgpu_deflate_hold(const uint8_t* src, size_t n, size_t history, int final,
                       int level, uint32_t chunk, uint32_t dict, int want_checksums,
                       size_t* out_len, uint32_t* adler0, uint32_t* crc, zzgpu_stats* stats)
{
    const CallArgs a = { src, n, nullptr, history, final, ZZGPU_MEM_HOST, nullptr, 0, ZZGPU_MEM_HOST, level, chunk, dict,
                         want_checksums, HostOut::Hold, nullptr, nullptr, 0 };
    return deflateCall(a, out_len, adler0, crc, stats);
}

int zzgpu_fetch(uint8_t* dst, size_t cap, zzgpu_sink_fn sink, void* user, size_t slice)
{
    Ctx* cp = nullptr;
    int rc = ensureCtx(cp); if (rc) return rc;
    Ctx& c = *cp;
    Lease lease(c);                                  // ends the hold when it goes out of scope
    if (!lease.resumed()) return fail(ZZGPU_E_ARG, "no held stream on this device for the calling thread");
    if ((dst != nullptr) == (sink != nullptr)) return fail(ZZGPU_E_ARG, "exactly one of dst / sink");
    lease.consume();
    const size_t len = c.heldLen;
    c.heldLen = 0;
    if (dst && len > cap) return fail(ZZGPU_E_CAPACITY, "destination too small");
    rc = ensureStaging(c); if (rc) return rc;
    HostOut out;
    out.kind = dst ? HostOut::Direct : HostOut::Sink;
    out.dst = dst; out.cap = cap; out.pinned = dst && isPinnedHost(dst);
    out.sink = sink; out.user = user; out.slice = slice ? slice : kDefaultSlice; out.firstSlice = false;
    rc = deliver(c, out, c.dOut, len, nullptr);
    if (rc) { drainStreams(c); return rc; }
    CK(cudaStreamSynchronize(c.copyOut));
    return ZZGPU_OK;
}

void zzgpu_release(void)
{
    Ctx* cp = nullptr;
    if (ensureCtx(cp)) return;
    Lease lease(*cp);
    lease.consume();
    cp->heldLen = 0;
}

int zzgpu_deflate(const uint8_t* src, size_t n, int src_mem, uint8_t* dst, size_t cap, int dst_mem,
                  int level, uint32_t chunk, uint32_t dict,
                  size_t* out_len, uint32_t* adler, uint32_t* crc, zzgpu_stats* stats)
{
    const int want = (adler ? 1 : 0) | (crc ? 2 : 0);
    uint32_t a0 = 0, r = 0;
    int rc = zzgpu_deflate_ex(src, n, 0, 1, src_mem, dst, cap, dst_mem, level, chunk, dict, want, out_len, &a0, &r, stats);
    if (rc) return rc;
    if (adler) *adler = zz::adler32_combine(1, a0, n);
    if (crc) *crc = r;
    return ZZGPU_OK;
}

int zzgpu_checksums(const uint8_t* src, size_t n, int src_mem, uint32_t adler_start, uint32_t crc_start,
                    uint32_t* adler, uint32_t* crc)
{
    if (!src && n) return fail(ZZGPU_E_ARG, "invalid argument");
    Ctx* cp = nullptr;
    int rc = ensureCtx(cp); if (rc) return rc;
    Ctx& c = *cp;
    Lease lease(c);
    if (lease.resumed()) return fail(ZZGPU_E_ARG, "this thread holds a stream on the device: fetch or release it first");
    uint32_t a = adler_start, r = crc_start;
    const uint32_t chunk = ZZGPU_MAX_CHUNK;
    const size_t slice = src_mem == ZZGPU_MEM_HOST ? (size_t)1 << 30 : n;      // host input: bounded device staging
    for (size_t off = 0; off < n; off += slice) {
        const size_t len = std::min(slice, n - off);
        const uint8_t* d_src = src + off;
        if (src_mem == ZZGPU_MEM_HOST) {
            rc = ensureBuf(c.dIn, c.dInCap, len + 64); if (rc) return rc;
            CK(cudaMemcpyAsync(c.dIn, src + off, len, cudaMemcpyHostToDevice, c.stream));
            d_src = c.dIn;
        }
        const uint64_t nchunks = (len + chunk - 1) / chunk;
        rc = ensureBuf(c.ck, c.ckCap, 2 * nchunks); if (rc) return rc;
        for (uint64_t first = 0; first < nchunks; first += 32768) {
            Job job{};
            job.src = d_src; job.n = len; job.chunk = chunk; job.dict = 0; job.first_chunk = first;
            job.nchunks = (uint32_t)std::min<uint64_t>(32768, nchunks - first);
            job.final_stream = 1; job.ck = c.ck; job.want_checksums = (adler ? 1 : 0) | (crc ? 2 : 0);
            launch_checksums(job, c.stream);
        }
        CK(cudaGetLastError());
        uint32_t a0 = 0, r0 = 0;
        rc = foldChecksums(c, len, chunk, 3, &a0, &r0); if (rc) return rc;
        a = zz::adler32_combine(a, a0, len);
        r = zz::crc32_combine(r, r0, len);
    }
    if (adler) *adler = a;
    if (crc) *crc = r;
    return ZZGPU_OK;
}

int zzgpu_set_option(const char* name, int value)
{
    if (name && !strcmp(name, "piece_first") && value >= 1 && value <= 65536) { g_pieceFirst = (uint64_t)value; return ZZGPU_OK; }
    if (name && !strcmp(name, "piece_mid") && value >= 1 && value <= 65536) { g_pieceMid = (uint64_t)value; return ZZGPU_OK; }
    if (name && !strcmp(name, "piece_last") && value >= 1 && value <= 65536) { g_pieceLast = (uint64_t)value; return ZZGPU_OK; }
    if (name && !strcmp(name, "lanes") && value >= 1 && value <= kMaxLanes) { g_optLanes = (int)value; return ZZGPU_OK; }
    if (name && !strcmp(name, "segment_mib") && value >= 1 && value <= 65536) { g_segBytes = (size_t)value << 20; return ZZGPU_OK; }
    if (name && set_kernel_option(name, value)) return ZZGPU_OK;
    return fail(ZZGPU_E_ARG, "unknown option");
}

long long zzgpu_get_counter(const char* name)
{
    if (name && !strcmp(name, "sink_pieces")) return t_sinkPieces;
    if (name && !strcmp(name, "sink_first_h2d_done")) return t_sinkFirstH2dDone;
    return -1;
}

uint32_t zzgpu_adler32_combine(uint32_t first, uint32_t second_start0, size_t len_second)
{
    return zz::adler32_combine(first, second_start0, len_second);
}

uint32_t zzgpu_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
    return zz::crc32_combine(crc1, crc2, len2);
}

int zzgpu_debug_chunk(const uint8_t* src, size_t n, int src_mem, int level, uint32_t chunk, uint32_t dict,
                      uint64_t chunk_index, uint16_t* cand, uint32_t* tokens, uint32_t max_tokens, uint32_t* n_tokens,
                      uint32_t* hist, uint8_t* lengths, uint32_t* info)
{
    if (chunk == 0) chunk = ZZGPU_DEFAULT_CHUNK;
    if (!validParams(level, chunk, dict) || level < 2 || !src || n == 0) return fail(ZZGPU_E_ARG, "invalid argument");
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    if (nchunks > kMaxSlots || chunk_index >= nchunks) return fail(ZZGPU_E_ARG, "debug tap needs a single batch");
    Ctx* cp = nullptr;
    int rc = ensureCtx(cp); if (rc) return rc;
    Ctx& c = *cp;
    Lease lease(c);
    if (lease.resumed()) return fail(ZZGPU_E_ARG, "this thread holds a stream on the device: fetch or release it first");
    const uint8_t* d_src = src;
    if (src_mem == ZZGPU_MEM_HOST) {
        rc = ensureBuf(c.dIn, c.dInCap, n + 64); if (rc) return rc;
        CK(cudaMemcpyAsync(c.dIn, src, n, cudaMemcpyHostToDevice, c.stream));
        d_src = c.dIn;
    }
    const size_t cap = zzgpu_bound(n, level, chunk);
    rc = ensureBuf(c.dOut, c.dOutCap, cap + 64); if (rc) return rc;
    uint64_t launches = 0;
    rc = runPipeline(c, d_src, n, 0, 1, c.dOut, cap, level, chunk, dict, 0, launches); if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    const size_t slot = (size_t)chunk_index;
    ChunkState st;
    CK(cudaMemcpy(&st, c.state + slot, sizeof st, cudaMemcpyDeviceToHost));
    if (cand) {
        // K-EMIT reuses the candidate rows as scratch: run K-CAND again for the tap
        const Job job = makeJob(c, d_src, n, 0, 1, c.dOut, cap, level, chunk, dict, 0, 0, (uint32_t)nchunks);
        launch_candidates(job, c.stream);
        CK(cudaStreamSynchronize(c.stream));
        CK(cudaMemcpy(cand, c.cand + slot * chunk, (size_t)chunk * 2, cudaMemcpyDeviceToHost));
    }
    if (n_tokens) *n_tokens = st.ntok;
    if (tokens) {
        const uint32_t cnt = std::min(st.ntok, max_tokens);
        std::vector<uint32_t> a(cnt); std::vector<uint16_t> d(cnt);
        if (cnt) {
            CK(cudaMemcpy(a.data(), c.tokA + slot * kMaxTokens, cnt * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(d.data(), c.tokD + slot * kMaxTokens, cnt * 2, cudaMemcpyDeviceToHost));
        }
        for (uint32_t i = 0; i < cnt; ++i) { tokens[3 * i] = a[i] & 0xFFFF; tokens[3 * i + 1] = a[i] >> 16; tokens[3 * i + 2] = d[i]; }
    }
    if (hist) CK(cudaMemcpy(hist, c.hist + slot * kHistStride, 316 * 4, cudaMemcpyDeviceToHost));
    if (lengths) {
        ChunkCodes* cc = c.codes + slot;
        CK(cudaMemcpy(lengths, cc->lens, 335, cudaMemcpyDeviceToHost));
    }
    if (info) { info[0] = st.block_type; info[1] = st.hdr_bits; info[2] = st.out_bytes; info[3] = (uint32_t)st.total_bits; }
    return ZZGPU_OK;
}

}  // extern "C"
