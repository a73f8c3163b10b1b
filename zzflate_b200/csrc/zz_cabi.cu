// C-ABI layer (include/zzgpu.h): per-device context, scratch, staging and the kernel pipeline.
// Replaces WriteDeflateStream (zzflate/zzflate.cpp:81-156) and the checksum calls of
// AppendChecksum (zzflate.cpp:170-192) for the host driver in zz_host.cpp.
#include "../../include/zzgpu.h"
#include "zz_kernels.cuh"

#include <atomic>
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

thread_local std::string t_lastError;
thread_local int t_device = -1;

struct Ctx {
    int device = -1;
    bool ready = false;
    std::mutex mu;
    cudaStream_t stream = nullptr;
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
    cudaStream_t lane[3] = { nullptr, nullptr, nullptr };  // overlap option: front-end and back-end streams
    cudaEvent_t laneEv[3] = { nullptr, nullptr, nullptr };
    std::vector<cudaEvent_t> offsEv;        // offsets scan of batch k (the next batch's scan continues its running total)
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

void freeScratch(Ctx& c)
{
    cudaFree(c.cand); cudaFree(c.info); cudaFree(c.tokA); cudaFree(c.tokD); cudaFree(c.hist); cudaFree(c.codes); cudaFree(c.state);
    c.cand = nullptr; c.info = nullptr; c.tokA = nullptr; c.tokD = nullptr; c.hist = nullptr; c.codes = nullptr; c.state = nullptr;
    c.slots = 0;
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
            CK(cudaSetDevice(dev));
            CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
            for (auto& e : c.ev) CK(cudaEventCreate(&e));
            CK(cudaMalloc(&c.total, 4 * sizeof(uint64_t)));
            CK(cudaMallocHost(&c.hTotal, 4 * sizeof(uint64_t)));
            CK(configure_kernels());
            c.device = dev;
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
    CK(cudaMalloc(&c.cand, (size_t)slots * chunk * sizeof(uint16_t)));
    CK(cudaMalloc(&c.info, (size_t)slots * chunk));
    CK(cudaMalloc(&c.tokA, (size_t)slots * kMaxTokens * sizeof(uint32_t)));
    CK(cudaMalloc(&c.tokD, (size_t)slots * kMaxTokens * sizeof(uint16_t)));
    CK(cudaMalloc(&c.hist, (size_t)slots * kHistStride * sizeof(uint32_t)));
    CK(cudaMalloc(&c.codes, (size_t)slots * sizeof(ChunkCodes)));
    CK(cudaMalloc(&c.state, (size_t)slots * sizeof(ChunkState)));
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
    if (e != cudaSuccess) return fail(ZZGPU_E_NOMEM, "allocation failed", e);
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

constexpr uint32_t kLaneChunks = 4096;      // chunks per batch when batches overlap
int g_overlap = 0;                          // zzgpu_set_option("overlap", 0/1)

Job makeJob(Ctx& c, const uint8_t* d_src, size_t n, size_t history, int final, uint8_t* d_dst, size_t cap, int level,
            uint32_t chunk, uint32_t dict, int wantCk, uint64_t first, uint32_t count, size_t so)
{
    Job job{};
    job.src = d_src; job.n = n; job.history = history; job.chunk = chunk; job.dict = dict;
    job.first_chunk = first; job.nchunks = count;
    job.final_stream = final; job.level = level; job.want_checksums = wantCk;
    job.cand = c.cand + so * chunk; job.info = c.info + so * chunk; job.tokA = c.tokA + so * kMaxTokens; job.tokD = c.tokD + so * kMaxTokens;
    job.hist = c.hist + so * kHistStride; job.codes = c.codes + so; job.state = c.state + so;
    job.dst = d_dst; job.cap = cap; job.total = c.total; job.ck = c.ck;
    return job;
}

// Kernel pipeline over chunks [firstChunk, lastChunk) of the call (geometry is always that of the whole call).
//
// Default: batches of up to 16 384 chunks, kernels back to back on one stream.
// "overlap" option: batches of 4096 chunks on two streams with two scratch slices.  Stream A runs K-CAND and K-MATCH,
// stream B runs K-HUFF, K-OFFS, K-EMIT, K-CKSUM.  K-MATCH needs whole SMs (its CTA takes all shared memory), so it
// is ordered after the previous batch's K-EMIT; what overlaps is K-CAND of batch k+1 (one warp + 34 KiB per CTA) with
// the back end of batch k.
int runChunks(Ctx& c, const uint8_t* d_src, size_t n, size_t history, int final, uint8_t* d_dst, size_t cap,
              int level, uint32_t chunk, uint32_t dict, int wantCk, uint64_t firstChunk, uint64_t lastChunk, uint64_t& launches)
{
    int rc;
    const uint64_t count = lastChunk - firstChunk;
    const bool overlap = g_overlap && level >= 2 && count >= 2 * kLaneChunks && c.slots >= 2 * kLaneChunks;
    if (!overlap) {
        for (uint64_t first = firstChunk; first < lastChunk; first += c.slots) {
            const Job job = makeJob(c, d_src, n, history, final, d_dst, cap, level, chunk, dict, wantCk, first,
                                    (uint32_t)std::min<uint64_t>(c.slots, lastChunk - first), 0);
            cudaStream_t st = c.stream;
            rc = markStage(c, -1, st); if (rc) return rc;
            if (level >= 2) {
                launches += launch_candidates(job, st); rc = markStage(c, ZZGPU_STAGE_CAND, st); if (rc) return rc;
                launches += launch_info(job, st); rc = markStage(c, ZZGPU_STAGE_INFO, st); if (rc) return rc;
                launches += launch_parse(job, st); rc = markStage(c, ZZGPU_STAGE_PARSE, st); if (rc) return rc;
            }
            if (level == 1) {
                launches += launch_fixed(job, st); rc = markStage(c, ZZGPU_STAGE_FIXED, st); if (rc) return rc;
            } else {
                launches += launch_huffman(job, st); rc = markStage(c, ZZGPU_STAGE_HUFF, st); if (rc) return rc;
            }
            launches += launch_offsets(job, st); rc = markStage(c, ZZGPU_STAGE_OFFS, st); if (rc) return rc;
            if (level == 1) { launches += launch_gather(job, st); rc = markStage(c, ZZGPU_STAGE_GATHER, st); if (rc) return rc; }
            else { launches += launch_emit(job, st); rc = markStage(c, ZZGPU_STAGE_EMIT, st); if (rc) return rc; }
            if (wantCk) { launches += launch_checksums(job, st); rc = markStage(c, ZZGPU_STAGE_CKSUM, st); if (rc) return rc; }
        }
        CK(cudaGetLastError());
        return ZZGPU_OK;
    }

    for (int i = 0; i < 2; ++i)
        if (!c.lane[i]) { CK(cudaStreamCreateWithFlags(&c.lane[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c.laneEv[i], cudaEventDisableTiming)); }
    const size_t nb = (size_t)((count + kLaneChunks - 1) / kLaneChunks);
    while (c.offsEv.size() < 2 * nb + 1) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c.offsEv.push_back(e); }
    cudaStream_t sA = c.lane[0], sB = c.lane[1];
    CK(cudaEventRecord(c.offsEv[2 * nb], c.stream));               // everything queued so far on the main stream
    CK(cudaStreamWaitEvent(sA, c.offsEv[2 * nb], 0));
    CK(cudaStreamWaitEvent(sB, c.offsEv[2 * nb], 0));
    auto evParse = [&](size_t k) { return c.offsEv[2 * k]; };
    auto evEmit = [&](size_t k) { return c.offsEv[2 * k + 1]; };
    size_t k = 0;
    for (uint64_t first = firstChunk; first < lastChunk; first += kLaneChunks, ++k) {
        const Job job = makeJob(c, d_src, n, history, final, d_dst, cap, level, chunk, dict, wantCk, first,
                                (uint32_t)std::min<uint64_t>(kLaneChunks, lastChunk - first), (k & 1) * (size_t)kLaneChunks);
        // stream A: candidates as soon as the scratch slice is free, the match kernel once the SMs are
        if (k >= 2) CK(cudaStreamWaitEvent(sA, evEmit(k - 2), 0));
        rc = markStage(c, -1, sA); if (rc) return rc;
        launches += launch_candidates(job, sA); rc = markStage(c, ZZGPU_STAGE_CAND, sA); if (rc) return rc;
        launches += launch_info(job, sA); rc = markStage(c, ZZGPU_STAGE_INFO, sA); if (rc) return rc;
        if (k >= 1) CK(cudaStreamWaitEvent(sA, evEmit(k - 1), 0));
        rc = markStage(c, -1, sA); if (rc) return rc;
        launches += launch_parse(job, sA); rc = markStage(c, ZZGPU_STAGE_PARSE, sA); if (rc) return rc;
        CK(cudaEventRecord(evParse(k), sA));
        // stream B: back end of the batch
        CK(cudaStreamWaitEvent(sB, evParse(k), 0));
        rc = markStage(c, -1, sB); if (rc) return rc;
        launches += launch_huffman(job, sB); rc = markStage(c, ZZGPU_STAGE_HUFF, sB); if (rc) return rc;
        launches += launch_offsets(job, sB); rc = markStage(c, ZZGPU_STAGE_OFFS, sB); if (rc) return rc;
        launches += launch_emit(job, sB); rc = markStage(c, ZZGPU_STAGE_EMIT, sB); if (rc) return rc;
        if (wantCk) { launches += launch_checksums(job, sB); rc = markStage(c, ZZGPU_STAGE_CKSUM, sB); if (rc) return rc; }
        CK(cudaEventRecord(evEmit(k), sB));
    }
    CK(cudaStreamWaitEvent(c.stream, evEmit(k - 1), 0));
    CK(cudaStreamWaitEvent(c.stream, evParse(k - 1), 0));
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
std::vector<uint64_t> pieceSchedule(uint64_t nchunks)
{
    // Every piece costs one launch of each kernel (the Huffman kernel alone is ~0.4 ms however few chunks it gets), so
    // pieces are few.  Their sizes are multiples of 888 chunks = 148 SMs x 6 resident K-CAND warps, which is also a
    // whole number of waves of K-MATCH / K-EMIT (296 CTAs) and K-INFO (148 chunks): 888, 1776, 3552, then 4440 chunks,
    // and the last <= 7104 chunks in two pieces (about 60/40) so that the final D2H is short.
    std::vector<uint64_t> ends;
    const uint64_t unit = 888, big = 5 * unit;
    uint64_t pos = 0, size = unit;
    while (nchunks - pos > big + 3 * unit) {
        pos += size; ends.push_back(pos);
        size = size < 4 * unit ? size * 2 : big;
    }
    const uint64_t rem = nchunks - pos;
    if (rem > 2 * unit) { uint64_t take = ((rem * 3 / 5 + unit / 2) / unit) * unit; if (take == 0 || take >= rem) take = rem / 2; pos += take; ends.push_back(pos); }
    if (pos < nchunks) ends.push_back(nchunks);
    return ends;
}

// Host buffers on both sides: H2D copies, kernels and D2H copies of successive pieces overlap on three streams.
int runHostPipelined(Ctx& c, const uint8_t* src, size_t n, size_t hist, int final, uint8_t* dst, size_t cap,
                     int level, uint32_t chunk, uint32_t dict, int wantCk, uint64_t& launches, size_t& total, size_t& d2h)
{
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    int rc = ensureBuf(c.dIn, c.dInCap, hist + n + 64); if (rc) return rc;
    const size_t d_cap = std::min(cap, zzgpu_bound(n, level, chunk));
    rc = ensureBuf(c.dOut, c.dOutCap, d_cap + 64); if (rc) return rc;
    if (!c.copyIn) { CK(cudaStreamCreateWithFlags(&c.copyIn, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c.copyOut, cudaStreamNonBlocking)); }
    const bool srcPinned = isPinnedHost(src), dstPinned = isPinnedHost(dst);
    if (!srcPinned || !dstPinned) {
        for (int i = 0; i < 2; ++i) {
            if (!c.stageIn[i]) { CK(cudaMallocHost(&c.stageIn[i], kStageBlock)); CK(cudaMallocHost(&c.stageOut[i], kStageBlock));
                                 CK(cudaEventCreateWithFlags(&c.stageInEv[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c.stageOutEv[i], cudaEventDisableTiming)); }
        }
    }
    size_t inBlocks = 0, outBlocks = 0;
    const std::vector<uint64_t> ends = pieceSchedule(nchunks);
    const size_t np = ends.size();
    while (c.pieceEv.size() < 2 * np) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c.pieceEv.push_back(e); }
    rc = ensureBuf(c.hPiece, c.hPieceCap, 4 * np, true); if (rc) return rc;
    rc = preparePipeline(c, n, chunk, wantCk); if (rc) return rc;
    cudaEvent_t start = c.pieceEv[0];                   // copies must not start before earlier work on the main stream is done
    CK(cudaEventRecord(c.ev[0], c.stream));
    CK(cudaStreamWaitEvent(c.copyIn, c.ev[0], 0));
    CK(cudaStreamWaitEvent(c.copyOut, c.ev[0], 0));
    (void)start;
    const uint8_t* d_src = c.dIn + hist;
    uint64_t firstChunk = 0;
    for (size_t p = 0; p < np; ++p) {
        const size_t lo = (size_t)(firstChunk * chunk), hi = (size_t)std::min<uint64_t>(n, ends[p] * chunk);
        const size_t from = p == 0 ? 0 : hist + lo, to = hist + hi;          // piece 0 also carries the history
        if (srcPinned) {
            CK(cudaMemcpyAsync(c.dIn + from, src - hist + from, to - from, cudaMemcpyHostToDevice, c.copyIn));
        } else {
            for (size_t off = from; off < to; off += kStageBlock, ++inBlocks) {
                const size_t len = std::min(kStageBlock, to - off);
                const int sb = (int)(inBlocks & 1);
                CK(cudaEventSynchronize(c.stageInEv[sb]));                   // the block's previous H2D has drained
                CopyPool::get().copy(c.stageIn[sb], src - hist + off, len);
                CK(cudaMemcpyAsync(c.dIn + off, c.stageIn[sb], len, cudaMemcpyHostToDevice, c.copyIn));
                CK(cudaEventRecord(c.stageInEv[sb], c.copyIn));
            }
        }
        CK(cudaEventRecord(c.pieceEv[2 * p], c.copyIn));
        CK(cudaStreamWaitEvent(c.stream, c.pieceEv[2 * p], 0));
        rc = runChunks(c, d_src, n, hist, final, c.dOut, d_cap, level, chunk, dict, wantCk, firstChunk, ends[p], launches);
        if (rc) return rc;
        CK(cudaMemcpyAsync(c.hPiece + 4 * p, c.total, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        CK(cudaEventRecord(c.pieceEv[2 * p + 1], c.stream));
        firstChunk = ends[p];
    }
    CK(cudaEventRecord(c.ev[2], c.stream));
    size_t done = 0;
    for (size_t p = 0; p < np; ++p) {
        CK(cudaEventSynchronize(c.pieceEv[2 * p + 1]));
        const uint64_t upto = c.hPiece[4 * p], flags = c.hPiece[4 * p + 1];
        if (flags & 1) return fail(ZZGPU_E_CAPACITY, "destination too small");
        if (flags & ~1ull) return fail(ZZGPU_E_CUDA, "internal consistency check failed (emit size mismatch)");
        if (upto > cap) return fail(ZZGPU_E_CAPACITY, "destination too small");
        if (upto > done) {
            if (dstPinned) {
                CK(cudaMemcpyAsync(dst + done, c.dOut + done, upto - done, cudaMemcpyDeviceToHost, c.copyOut));
            } else {
                // D2H of block k+1 runs while block k is copied out of its pinned stage
                size_t pendOff = 0, pendLen = 0; int pendSb = 0;
                for (size_t off = done; off < upto || pendLen; ) {
                    size_t len = 0; int sb = 0;
                    if (off < upto) {
                        len = std::min(kStageBlock, (size_t)upto - off); sb = (int)(outBlocks++ & 1);
                        CK(cudaMemcpyAsync(c.stageOut[sb], c.dOut + off, len, cudaMemcpyDeviceToHost, c.copyOut));
                        CK(cudaEventRecord(c.stageOutEv[sb], c.copyOut));
                    }
                    if (pendLen) { CK(cudaEventSynchronize(c.stageOutEv[pendSb])); CopyPool::get().copy(dst + pendOff, c.stageOut[pendSb], pendLen); }
                    pendOff = off; pendLen = len; pendSb = sb;
                    off += len;
                }
            }
        }
        done = (size_t)upto;
    }
    CK(cudaEventRecord(c.ev[3], c.copyOut));
    CK(cudaStreamWaitEvent(c.stream, c.ev[3], 0));
    CK(cudaStreamSynchronize(c.copyOut));
    for (int i = 0; i < 4; ++i) c.hTotal[i] = c.hPiece[4 * (np - 1) + i];
    total = done; d2h = done;
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

const uint8_t kEmptyFinalStored[5] = { 0x01, 0x00, 0x00, 0xFF, 0xFF };     // R7: empty input still gets one final block

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
        cudaSetDevice(c.device);
        cudaStreamSynchronize(c.stream);
        freeScratch(c);
        cudaFree(c.total); cudaFreeHost(c.hTotal); cudaFree(c.ck); cudaFreeHost(c.hCk); cudaFree(c.dIn); cudaFree(c.dOut);
        cudaFreeHost(c.hPiece); c.hPiece = nullptr; c.hPieceCap = 0;
        for (auto& e : c.pieceEv) cudaEventDestroy(e);
        c.pieceEv.clear();
        for (auto& e : c.stageEv) cudaEventDestroy(e);
        c.stageEv.clear(); c.stageOf.clear(); c.stageUsed = 0;
        if (c.copyIn) { cudaStreamDestroy(c.copyIn); cudaStreamDestroy(c.copyOut); c.copyIn = nullptr; c.copyOut = nullptr; }
        for (int i = 0; i < 2; ++i) if (c.stageIn[i]) { cudaFreeHost(c.stageIn[i]); cudaFreeHost(c.stageOut[i]); cudaEventDestroy(c.stageInEv[i]); cudaEventDestroy(c.stageOutEv[i]); c.stageIn[i] = c.stageOut[i] = nullptr; }
        for (auto& e : c.ev) cudaEventDestroy(e);
        cudaStreamDestroy(c.stream);
        c.total = nullptr; c.hTotal = nullptr; c.ck = nullptr; c.hCk = nullptr; c.dIn = nullptr; c.dOut = nullptr;
        c.ckCap = c.hCkCap = c.dInCap = c.dOutCap = 0;
        c.ready = false;
    }
}

int zzgpu_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
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
    if (chunk == 0) chunk = ZZGPU_DEFAULT_CHUNK;
    if (!validParams(level, chunk, dict) || !out_len || (!src && n) || (!dst && cap)) return fail(ZZGPU_E_ARG, "invalid argument");
    Ctx* cp = nullptr;
    int rc = ensureCtx(cp); if (rc) return rc;
    Ctx& c = *cp;
    std::lock_guard<std::mutex> lk(c.mu);
    if (stats) memset(stats, 0, sizeof *stats);
    if (adler0) *adler0 = 0;
    if (crc) *crc = 0;

    if (n == 0) {
        *out_len = 0;
        if (final) {
            if (cap < 5) return fail(ZZGPU_E_CAPACITY, "destination too small");
            if (dst_mem == ZZGPU_MEM_HOST) memcpy(dst, kEmptyFinalStored, 5);
            else { CK(cudaMemcpyAsync(dst, kEmptyFinalStored, 5, cudaMemcpyHostToDevice, c.stream)); CK(cudaStreamSynchronize(c.stream)); }
            *out_len = 5;
        }
        return ZZGPU_OK;
    }

    uint64_t launches = 0;
    const size_t hist = std::min<size_t>(history, (size_t)dict + kPreExtra);    // bytes the kernels may look at
    size_t h2d = 0, d2h = 0;
    uint64_t total = 0;
    if (src_mem == ZZGPU_MEM_HOST && dst_mem == ZZGPU_MEM_HOST && n >= ((size_t)32 << 20)) {
        size_t tot = 0;
        rc = runHostPipelined(c, src, n, hist, final, dst, cap, level, chunk, dict, want_checksums, launches, tot, d2h);
        if (rc) { cudaStreamSynchronize(c.copyIn); cudaStreamSynchronize(c.stream); cudaStreamSynchronize(c.copyOut); return rc; }
        total = tot; h2d = hist + n;
    } else {
        CK(cudaEventRecord(c.ev[0], c.stream));
        const uint8_t* d_src = src;
        if (src_mem == ZZGPU_MEM_HOST) {
            rc = ensureBuf(c.dIn, c.dInCap, hist + n + 64); if (rc) return rc;
            CK(cudaMemcpyAsync(c.dIn, src - hist, hist + n, cudaMemcpyHostToDevice, c.stream));
            d_src = c.dIn + hist;
            h2d = hist + n;
        }
        uint8_t* d_dst = dst;
        size_t d_cap = cap;
        if (dst_mem == ZZGPU_MEM_HOST) {
            d_cap = std::min(cap, zzgpu_bound(n, level, chunk));
            rc = ensureBuf(c.dOut, c.dOutCap, d_cap + 64); if (rc) return rc;
            d_dst = c.dOut;
        }
        CK(cudaEventRecord(c.ev[1], c.stream));
        rc = runPipeline(c, d_src, n, hist, final, d_dst, d_cap, level, chunk, dict, want_checksums, launches);
        if (rc) return rc;
        CK(cudaEventRecord(c.ev[2], c.stream));
        CK(cudaMemcpyAsync(c.hTotal, c.total, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        total = c.hTotal[0];
        const uint64_t flags = c.hTotal[1];
        if (flags & 1) return fail(ZZGPU_E_CAPACITY, "destination too small");
        if (flags & ~1ull) return fail(ZZGPU_E_CUDA, "internal consistency check failed (emit size mismatch)");
        if (total > cap) return fail(ZZGPU_E_CAPACITY, "destination too small");
        if (dst_mem == ZZGPU_MEM_HOST) {
            CK(cudaMemcpyAsync(dst, c.dOut, total, cudaMemcpyDeviceToHost, c.stream));
            d2h = total;
        }
        CK(cudaEventRecord(c.ev[3], c.stream));
    }
    rc = foldChecksums(c, n, chunk, want_checksums, adler0, crc); if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    *out_len = (size_t)total;
    if (stats) {
        stats->chunks = (n + chunk - 1) / chunk;
        stats->matches = c.hTotal[2];
        stats->stored_chunks = c.hTotal[3];
        stats->kernel_launches = launches;
        collectStages(c, stats);
        cudaEventElapsedTime(&stats->total_ms, c.ev[0], c.ev[3]);
        if (h2d + d2h > 0) {
            stats->device_ms = 0;                    // host buffers: kernels interleave with copies; sum the stages
            for (int i = 0; i < ZZGPU_NSTAGES; ++i) stats->device_ms += stats->stage_ms[i];
        } else {
            cudaEventElapsedTime(&stats->device_ms, c.ev[1], c.ev[2]);
        }
        (void)cudaGetLastError();                    // timing queries must never poison the next call
        stats->h2d_bytes = h2d; stats->d2h_bytes = d2h;
    }
    return ZZGPU_OK;
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
    std::lock_guard<std::mutex> lk(c.mu);
    uint32_t a0 = 0, r = 0;
    if (n) {
        const uint8_t* d_src = src;
        if (src_mem == ZZGPU_MEM_HOST) {
            rc = ensureBuf(c.dIn, c.dInCap, n + 64); if (rc) return rc;
            CK(cudaMemcpyAsync(c.dIn, src, n, cudaMemcpyHostToDevice, c.stream));
            d_src = c.dIn;
        }
        const uint32_t chunk = ZZGPU_MAX_CHUNK;
        const uint64_t nchunks = (n + chunk - 1) / chunk;
        rc = ensureBuf(c.ck, c.ckCap, 2 * nchunks); if (rc) return rc;
        for (uint64_t first = 0; first < nchunks; first += 32768) {
            Job job{};
            job.src = d_src; job.n = n; job.chunk = chunk; job.dict = 0; job.first_chunk = first;
            job.nchunks = (uint32_t)std::min<uint64_t>(32768, nchunks - first);
            job.final_stream = 1; job.ck = c.ck;
            launch_checksums(job, c.stream);
        }
        CK(cudaGetLastError());
        rc = foldChecksums(c, n, chunk, 3, &a0, &r); if (rc) return rc;
    }
    if (adler) *adler = zz::adler32_combine(adler_start, a0, n);
    if (crc) *crc = zz::crc32_combine(crc_start, r, n);
    return ZZGPU_OK;
}

int zzgpu_set_option(const char* name, int value)
{
    if (name && !strcmp(name, "overlap")) { g_overlap = value ? 1 : 0; return ZZGPU_OK; }
    if (name && !strcmp(name, "emit")) { set_emit_variant(value ? 1 : 0); return ZZGPU_OK; }       // 0: position-range K-EMIT, 1: token-parallel (default)
    return fail(ZZGPU_E_ARG, "unknown option");
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
    std::lock_guard<std::mutex> lk(c.mu);
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
        const Job job = makeJob(c, d_src, n, 0, 1, c.dOut, cap, level, chunk, dict, 0, 0, (uint32_t)nchunks, 0);
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
