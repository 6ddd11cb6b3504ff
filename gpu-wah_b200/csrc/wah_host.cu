// wah_host.cu -- the host-buffer entry points of the C ABI (wah_compress_host / wah_decompress_host):
// what the reference's compress() / decompress() do around their kernels (compress.cu:41-209,
// decompress.cu:18-141), rebuilt for a PCIe-attached B200.
//
// The reference allocates and frees every device buffer inside each call and copies pageable memory
// synchronously.  Here a process-wide context keeps the device buffers, a ring of pinned bounce
// buffers and a small pool of copy threads alive between calls:
//   * a host buffer that is already page locked is DMAed directly;
//   * a pageable buffer moves in 16 MiB chunks through the pinned ring, the copy threads moving
//     chunk k+1 between user memory and the ring while the DMA engine moves chunk k;
//   * the result comes from calloc() because the reference's callers free() it (compress.cu:181,208); the copy
//     threads fill it in parallel and skip the 4 KiB blocks that are all zero, whose pages are then never touched.
#include "../../include/wah_b200.h"
#include "wah_kernels.h"

#include <emmintrin.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

using namespace wahb200;

int wah_set_error(int code, const char *fmt, ...);   // wah_capi.cu

#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return wah_set_error(e__ == cudaErrorMemoryAllocation ? WAH_ERR_NOMEM : WAH_ERR_CUDA, "%s: %s", \
                                 #expr, cudaGetErrorString(e__));                                       \
    } while (0)

namespace {

constexpr size_t CHUNK = 16u << 20;  // bytes per pinned bounce buffer
constexpr int NSLOT = 4;             // bounce buffers in the ring
// Sparse transfers: a bitvector worth compressing is mostly zero, and a block of zeros need not cross PCIe -- the input
// buffer on the device is cleared and only the 4 KiB blocks that hold a set bit are sent (packed by the copy threads,
// put in place by a small kernel); the decoded vector's non-zero blocks are packed on the device and only those come
// back, into a result that calloc() handed out as zeros.  A chunk without a zero block moves as before.
constexpr size_t XBLK = 4096;                        // bytes per block
constexpr size_t BLKS = CHUNK / XBLK;                // blocks per chunk
constexpr size_t SLOT = CHUNK + BLKS * 4;            // a bounce buffer: the chunk (or its packed blocks) + their block numbers
constexpr size_t SPARSE_MIN = 8u << 20;              // smaller transfers are not worth the bookkeeping

// ------------------------------------------------------------------ copy threads

// Copy with non-temporal stores: the destination (a result buffer the caller reads later, or a bounce buffer
// the DMA engine reads) is not wanted in this core's cache, and streaming stores spare the read-for-ownership
// of every destination line.
inline void stream_copy(void *dst, const void *src, size_t bytes)
{
    char *d = static_cast<char *>(dst);
    const char *s = static_cast<const char *>(src);
    const size_t head = std::min(bytes, (size_t)((16 - ((uintptr_t)d & 15)) & 15));
    memcpy(d, s, head);
    d += head;
    s += head;
    bytes -= head;
    size_t n16 = bytes / 16;
    for (; n16 >= 4; n16 -= 4, d += 64, s += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(s)), b = _mm_loadu_si128((const __m128i *)(s + 16));
        const __m128i c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
        _mm_stream_si128((__m128i *)(d), a);
        _mm_stream_si128((__m128i *)(d + 16), b);
        _mm_stream_si128((__m128i *)(d + 32), c);
        _mm_stream_si128((__m128i *)(d + 48), e);
    }
    for (; n16; n16--, d += 16, s += 16) _mm_stream_si128((__m128i *)d, _mm_loadu_si128((const __m128i *)s));
    _mm_sfence();
    memcpy(d, s, bytes & 15);
}

inline bool all_zero(const char *p, size_t bytes)
{
    size_t i = 0;
    for (; i < bytes && ((uintptr_t)(p + i) & 15); i++)
        if (p[i]) return false;
    for (; i + 64 <= bytes; i += 64) {
        const __m128i a = _mm_load_si128((const __m128i *)(p + i)), b = _mm_load_si128((const __m128i *)(p + i + 16));
        const __m128i c = _mm_load_si128((const __m128i *)(p + i + 32)), d = _mm_load_si128((const __m128i *)(p + i + 48));
        const __m128i o = _mm_or_si128(_mm_or_si128(a, b), _mm_or_si128(c, d));
        if (_mm_movemask_epi8(_mm_cmpeq_epi8(o, _mm_setzero_si128())) != 0xFFFF) return false;
    }
    for (; i < bytes; i++)
        if (p[i]) return false;
    return true;
}

class CopyPool {
   public:
    explicit CopyPool(int n) : n_(n)
    {
        for (int i = 1; i < n_; i++) threads_.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    // run job(i) for i in [0, size()) on the pool; the calling thread takes i = 0
    void parallel(const std::function<void(int)> &job)
    {
        if (n_ == 1) {
            job(0);
            return;
        }
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = &job;
            pending_ = n_ - 1;
            gen_++;
        }
        cv_.notify_all();
        job(0);
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [this] { return pending_ == 0; });
    }
    // memcpy split over the pool: page-aligned pieces of the destination, dealt round robin (a freshly
    // malloc()ed destination is faulted in by the copy itself, every page by exactly one thread)
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (bytes < (256u << 10) || n_ == 1) {
            memcpy(dst, src, bytes);
            return;
        }
        static const uintptr_t piece = [] {
            const char *e = getenv("WAH_B200_PIECE_KB");
            const long kb = e ? atol(e) : 1024;
            uintptr_t v = 64u << 10;   // a power of two (the pieces are aligned by masking), at least 64 KiB
            while ((long)(v >> 10) * 2 <= kb && v < ((uintptr_t)1 << 30)) v <<= 1;
            return v;
        }();
        const uintptr_t d0 = (uintptr_t)dst, d1 = d0 + bytes, first = d0 & ~(piece - 1);
        parallel([&](int i) {
            for (uintptr_t r = first + (uintptr_t)i * piece; r < d1; r += (uintptr_t)n_ * piece) {
                const uintptr_t a = std::max(r, d0), b = std::min(r + piece, d1);
                stream_copy((void *)a, (const char *)src + (a - d0), b - a);
            }
        });
    }

    // The same into a destination that is known to hold zeros already (a fresh calloc() block): 4 KiB blocks of the
    // source that are all zero are not written at all.  A decoded bitvector of the fill-dominated kind is mostly such
    // blocks; their pages of the result are never touched, i.e. never faulted in and zeroed by the kernel either --
    // they stay the shared zero page until the caller writes to them.
    // Populate the page tables of a freshly allocated buffer, every thread its own contiguous part, with one
    // madvise(MADV_POPULATE_WRITE) per part (Linux >= 5.14) instead of one page fault per page; falls back to
    // touching the pages.
    void prefault(void *p, size_t bytes)
    {
        const uintptr_t lo = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)p + bytes) & ~(uintptr_t)4095;
        if (hi <= lo) return;
        const size_t per = (((hi - lo) / n_) + 4095) & ~(size_t)4095;
        parallel([&](int i) {
            const uintptr_t a = std::min(hi, lo + per * i), b = i == n_ - 1 ? hi : std::min(hi, lo + per * (i + 1));
            if (b <= a) return;
#ifdef MADV_POPULATE_WRITE
            if (madvise((void *)a, b - a, MADV_POPULATE_WRITE) == 0) return;
#endif
            for (uintptr_t o = a; o < b; o += 4096) *(volatile char *)o = 0;
        });
    }

    void copy_into_zeroed(void *dst, const void *src, size_t bytes)
    {
        constexpr uintptr_t BLK = 4096;
        const uintptr_t d0 = (uintptr_t)dst, d1 = d0 + bytes;
        const uintptr_t first = d0 & ~(BLK - 1);
        const size_t nblk = (d1 - first + BLK - 1) / BLK;
        auto work = [&](int i) {
            // contiguous runs of blocks per thread (16 blocks at a time), dealt round robin
            for (size_t b0 = (size_t)i * 16; b0 < nblk; b0 += (size_t)n_ * 16) {
                for (size_t b = b0; b < b0 + 16 && b < nblk; b++) {
                    const uintptr_t a = std::max(first + b * BLK, d0), e = std::min(first + (b + 1) * BLK, d1);
                    const char *sp = (const char *)src + (a - d0);
                    if (!all_zero(sp, e - a)) stream_copy((void *)a, sp, e - a);
                }
            }
        };
        if (bytes < (256u << 10) || n_ == 1) {
            for (int i = 0; i < n_; i++) work(i);
            return;
        }
        parallel(work);
    }

    // Packs the blocks of src (len bytes) that hold a non-zero byte into `packed`, in order, and their numbers into
    // `list`; returns how many.  The last block may be short: it is padded with zeros.
    size_t pack_nonzero(char *packed, uint32_t *list, const char *src, size_t len)
    {
        const size_t nblk = (len + XBLK - 1) / XBLK;
        const size_t per = (nblk + n_ - 1) / n_;
        nz_.assign(nblk, 0);
        cnt_.assign(n_ + 1, 0);
        parallel([&](int i) {
            const size_t b0 = std::min(nblk, per * i), b1 = std::min(nblk, per * (i + 1));
            size_t c = 0;
            for (size_t b = b0; b < b1; b++) {
                const bool z = all_zero(src + b * XBLK, std::min(XBLK, len - b * XBLK));
                nz_[b] = !z;
                c += !z;
            }
            cnt_[i + 1] = c;
        });
        for (int i = 0; i < n_; i++) cnt_[i + 1] += cnt_[i];
        parallel([&](int i) {
            const size_t b0 = std::min(nblk, per * i), b1 = std::min(nblk, per * (i + 1));
            size_t pos = cnt_[i];
            for (size_t b = b0; b < b1; b++) {
                if (!nz_[b]) continue;
                const size_t n = std::min(XBLK, len - b * XBLK);
                stream_copy(packed + pos * XBLK, src + b * XBLK, n);
                if (n < XBLK) memset(packed + pos * XBLK + n, 0, XBLK - n);
                list[pos++] = (uint32_t)b;
            }
        });
        return cnt_[n_];
    }

    // the reverse on the way back: block list[i] of dst (dst_len bytes) = packed[i]
    void scatter_blocks(char *dst, size_t dst_len, const char *packed, const uint32_t *list, size_t n)
    {
        // (a contiguous stretch of the list per thread: threads that fault in neighbouring pages of a fresh result
        //  fight over the same page-table lock -- dealt round robin this was 2 us of copying and 6 us of waiting per block)
        const size_t per = (n + n_ - 1) / n_;
        auto work = [&](int i) {
            const size_t k1 = std::min(n, per * (i + 1));
            for (size_t k = std::min(n, per * i); k < k1; k++) {
                const size_t off = (size_t)list[k] * XBLK;
                if (off < dst_len) stream_copy(dst + off, packed + k * XBLK, std::min(XBLK, dst_len - off));
            }
        };
        if (n < 64 || n_ == 1) {
            for (int i = 0; i < n_; i++) work(i);
            return;
        }
        parallel(work);
    }

   private:
    std::vector<uint8_t> nz_;
    std::vector<size_t> cnt_;
    void loop(int i)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *job;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = job_;
            }
            (*job)(i);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    uint64_t gen_ = 0;
    bool stop_ = false;
    int pending_ = 0;
    const std::function<void(int)> *job_ = nullptr;
};

// ------------------------------------------------------------------ context

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        // grow in 2 MiB steps so that slightly different sizes reuse the buffer
        const size_t want = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostCtx {
    std::mutex mu;
    int device = -1;
    DevBuf a, b, ws, small;          // input, output, workspace, 64 B of scalars
    DevBuf stg, pack, flags, lists, counts;   // sparse transfers: staging slots; packed blocks, block flags, lists, counts
    uint64_t h2d_bytes = 0, d2h_bytes = 0;    // bytes the last entry point moved over PCIe
    char *pin = nullptr;             // NSLOT * SLOT bytes, page locked
    cudaStream_t stream = nullptr;   // everything is ordered on this stream
    cudaEvent_t slot_ev[NSLOT] = {};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    CopyPool *pool = nullptr;

    int init()
    {
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (device == dev && stream) return WAH_OK;
        release();
        CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaHostAlloc((void **)&pin, NSLOT * SLOT, cudaHostAllocDefault));
        for (int i = 0; i < NSLOT; i++) CUDA_TRY(cudaEventCreateWithFlags(&slot_ev[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreate(&t0));
        CUDA_TRY(cudaEventCreate(&t1));
        CUDA_TRY(small.reserve(64));
        if (!pool) {
            int n = (int)std::thread::hardware_concurrency();
            if (const char *e = getenv("WAH_B200_COPY_THREADS")) n = atoi(e);
            pool = new CopyPool(std::max(1, std::min(n, 16)));
        }
        device = dev;
        return WAH_OK;
    }
    void release()
    {
        a.release();
        b.release();
        ws.release();
        small.release();
        stg.release();
        pack.release();
        flags.release();
        lists.release();
        counts.release();
        if (pin) cudaFreeHost(pin);
        pin = nullptr;
        for (auto &e : slot_ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        t0 = t1 = nullptr;
        if (stream) {
            cudaStreamSynchronize(stream);
            forget_stream(stream);   // the per-device launch order must not record events on it any more
            cudaStreamDestroy(stream);
        }
        stream = nullptr;
        device = -1;
    }
    float lap()   // milliseconds since the previous lap (stream drained)
    {
        float ms = 0.f;
        cudaEventRecord(t1, stream);
        cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1);
        cudaEventRecord(t0, stream);
        return ms;
    }
};

HostCtx g_ctx;

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

bool sparse_copy();
int result_strategy();

// host -> device, asynchronous on ctx.stream for pinned memory, pipelined through the ring otherwise
int upload(HostCtx &c, void *d_dst, const void *h_src, size_t bytes)
{
    if (bytes == 0) return WAH_OK;
    if (is_pinned(h_src)) {
        CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c.stream));
        c.h2d_bytes += bytes;
        return WAH_OK;
    }
    const size_t n = (bytes + CHUNK - 1) / CHUNK;
    const bool sparse = sparse_copy() && bytes >= SPARSE_MIN;
    if (sparse) {
        CUDA_TRY(c.stg.reserve(NSLOT * SLOT));
        CUDA_TRY(cudaMemsetAsync(d_dst, 0, (bytes + 15) & ~(size_t)15, c.stream));   // (the device buffers have 16 bytes of slack)
    }
    for (size_t k = 0; k < n; k++) {
        const int s = (int)(k % NSLOT);
        const size_t off = k * CHUNK, len = std::min(CHUNK, bytes - off);
        char *slot = c.pin + s * SLOT;
        if (k >= NSLOT) CUDA_TRY(cudaEventSynchronize(c.slot_ev[s]));   // the DMA out of this slot is done
        if (!sparse) {
            c.pool->copy(slot, (const char *)h_src + off, len);
            CUDA_TRY(cudaMemcpyAsync((char *)d_dst + off, slot, len, cudaMemcpyHostToDevice, c.stream));
            c.h2d_bytes += len;
        } else {
            uint32_t *list = reinterpret_cast<uint32_t *>(slot + CHUNK);
            const size_t nblk = (len + XBLK - 1) / XBLK;
            const size_t nnz = c.pool->pack_nonzero(slot, list, (const char *)h_src + off, len);
            if (nnz == nblk) {   // no zero block in the chunk: straight into place
                CUDA_TRY(cudaMemcpyAsync((char *)d_dst + off, slot, len, cudaMemcpyHostToDevice, c.stream));
                c.h2d_bytes += len;
            } else if (nnz != 0) {
                char *d_slot = (char *)c.stg.p + s * SLOT;
                CUDA_TRY(cudaMemcpyAsync(d_slot, slot, nnz * XBLK, cudaMemcpyHostToDevice, c.stream));
                CUDA_TRY(cudaMemcpyAsync(d_slot + CHUNK, list, nnz * 4, cudaMemcpyHostToDevice, c.stream));
                CUDA_TRY(launch_scatter_blocks((char *)d_dst + off, (len + 15) & ~(size_t)15, d_slot,
                                               reinterpret_cast<const uint32_t *>(d_slot + CHUNK), (uint32_t)nnz, c.stream));
                c.h2d_bytes += nnz * (XBLK + 4);
            }
        }
        CUDA_TRY(cudaEventRecord(c.slot_ev[s], c.stream));
    }
    return WAH_OK;
}

// How a result buffer is allocated and filled (WAH_B200_RESULT): 0 malloc + parallel copy; 1 malloc, advised towards
// huge pages, page tables populated by the copy threads while the first chunks are in flight; 2 calloc + zero blocks
// skipped; 3 = 2 with the huge-page advice.
int result_strategy()
{
    static const int st = [] {
        const char *e = getenv("WAH_B200_RESULT");
        return e ? atoi(e) : 2;
    }();
    return st;
}

bool sparse_copy()
{
    static const bool on = [] {
        const char *e = getenv("WAH_B200_SPARSE_COPY");
        return e ? e[0] != '0' : true;
    }();
    return on;
}

// device -> host into a fresh result the caller will free(): only the blocks with a set bit.  How the result is
// handed out depends on how much of it is non-zero: mostly zeros -- calloc() and nothing else, the pages of the zero
// blocks are never touched (never faulted in, never zeroed by the kernel); otherwise the pages are populated up front
// by the copy threads, as huge pages where the kernel allows it (a fault per 4 KiB page of a 2 GiB result costs more
// than the transfer: a 2 GiB round trip with nearly every block non-zero took 195 ms that way, 122 ms populated).
int download_sparse(HostCtx &c, uint32_t **h_out, const void *d_src, size_t bytes)
{
    const size_t n_blocks = (bytes + XBLK - 1) / XBLK, n = (bytes + CHUNK - 1) / CHUNK;
    CUDA_TRY(c.pack.reserve(n * CHUNK));
    CUDA_TRY(c.flags.reserve(n_blocks));
    CUDA_TRY(c.lists.reserve(n * BLKS * 4));
    CUDA_TRY(c.counts.reserve(n * 4));
    CUDA_TRY(launch_pack_nonzero_blocks(d_src, bytes, (uint32_t)BLKS, (uint8_t *)c.flags.p, (uint32_t *)c.lists.p, (uint32_t *)c.counts.p,
                                        c.pack.p, c.stream));
    static const bool trace = getenv("WAH_B200_HOST_TRACE") != nullptr;
    const auto T0 = std::chrono::steady_clock::now();
    double t_wait = 0, t_scatter = 0;
    std::vector<uint32_t> cnt(n);
    CUDA_TRY(cudaMemcpyAsync(cnt.data(), c.counts.p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    const auto T1 = std::chrono::steady_clock::now();
    std::vector<size_t> live;   // chunks that hold anything
    size_t nnz = 0;
    for (size_t k = 0; k < n; k++) {
        if (cnt[k] != 0) live.push_back(k);
        nnz += cnt[k];
    }
    char *h_dst = (char *)calloc(bytes, 1);
    if (!h_dst) return wah_set_error(WAH_ERR_NOMEM, "calloc of %zu bytes failed", bytes);
    *h_out = (uint32_t *)h_dst;
    if (nnz * 2 >= n_blocks) {   // half of the blocks or more (measured: 0.29 of them is still 8 ms better left alone)
        const uintptr_t lo = ((uintptr_t)h_dst + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
        const uintptr_t hi = ((uintptr_t)h_dst + bytes) & ~(uintptr_t)((2u << 20) - 1);
        if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
        c.pool->prefault(h_dst, bytes);
    }
    auto chunk_len = [&](size_t k) { return std::min(CHUNK, bytes - k * CHUNK); };
    auto dense = [&](size_t k) { return cnt[k] == (chunk_len(k) + XBLK - 1) / XBLK; };
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t k = live[i];
        char *slot = c.pin + (i % NSLOT) * SLOT;
        cudaError_t e;
        if (dense(k)) {   // no zero block: straight from the decoded vector
            e = cudaMemcpyAsync(slot, (const char *)d_src + k * CHUNK, chunk_len(k), cudaMemcpyDeviceToHost, c.stream);
            c.d2h_bytes += chunk_len(k);
        } else {
            e = cudaMemcpyAsync(slot, (const char *)c.pack.p + k * CHUNK, (size_t)cnt[k] * XBLK, cudaMemcpyDeviceToHost, c.stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(slot + CHUNK, (const char *)c.lists.p + k * BLKS * 4, (size_t)cnt[k] * 4, cudaMemcpyDeviceToHost, c.stream);
            c.d2h_bytes += (size_t)cnt[k] * (XBLK + 4);
        }
        if (e != cudaSuccess) return e;
        return cudaEventRecord(c.slot_ev[i % NSLOT], c.stream);
    };
    for (size_t i = 0; i < live.size() && i < NSLOT; i++) CUDA_TRY(issue(i));
    for (size_t i = 0; i < live.size(); i++) {
        const size_t k = live[i];
        const char *slot = c.pin + (i % NSLOT) * SLOT;
        const auto a0 = std::chrono::steady_clock::now();
        CUDA_TRY(cudaEventSynchronize(c.slot_ev[i % NSLOT]));
        const auto a1 = std::chrono::steady_clock::now();
        if (dense(k))
            c.pool->copy((char *)h_dst + k * CHUNK, slot, chunk_len(k));
        else
            c.pool->scatter_blocks((char *)h_dst + k * CHUNK, chunk_len(k), slot, reinterpret_cast<const uint32_t *>(slot + CHUNK), cnt[k]);
        const auto a2 = std::chrono::steady_clock::now();
        t_wait += std::chrono::duration<double, std::milli>(a1 - a0).count();
        t_scatter += std::chrono::duration<double, std::milli>(a2 - a1).count();
        if (i + NSLOT < live.size()) CUDA_TRY(issue(i + NSLOT));
    }
    if (trace)
        fprintf(stderr, "download_sparse: %zu chunks (%zu live), pack+sync %.2f ms, DMA waits %.2f ms, scatter %.2f ms\n", n, live.size(),
                std::chrono::duration<double, std::milli>(T1 - T0).count(), t_wait, t_scatter);
    return WAH_OK;
}

// device -> host; returns with the data in place
int download(HostCtx &c, void *h_dst, const void *d_src, size_t bytes, bool zeroed)
{
    if (bytes == 0) return WAH_OK;
    if (is_pinned(h_dst)) {
        CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c.stream));
        CUDA_TRY(cudaStreamSynchronize(c.stream));
        c.d2h_bytes += bytes;
        return WAH_OK;
    }
    const size_t n = (bytes + CHUNK - 1) / CHUNK;
    c.d2h_bytes += bytes;
    auto issue = [&](size_t k) -> cudaError_t {
        const int s = (int)(k % NSLOT);
        const size_t off = k * CHUNK, len = std::min(CHUNK, bytes - off);
        cudaError_t e = cudaMemcpyAsync(c.pin + s * SLOT, (const char *)d_src + off, len, cudaMemcpyDeviceToHost, c.stream);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(c.slot_ev[s], c.stream);
    };
    for (size_t k = 0; k < n && k < NSLOT; k++) CUDA_TRY(issue(k));
    if (zeroed && result_strategy() == 1 && bytes >= (4u << 20)) c.pool->prefault(h_dst, bytes);   // while the first chunks are in flight
    if (result_strategy() < 2) zeroed = false;
    // `zeroed`: the destination is a fresh calloc() block: blocks of zeros are not written (nor their pages faulted in)
    for (size_t k = 0; k < n; k++) {
        const int s = (int)(k % NSLOT);
        const size_t off = k * CHUNK, len = std::min(CHUNK, bytes - off);
        CUDA_TRY(cudaEventSynchronize(c.slot_ev[s]));
        if (zeroed)
            c.pool->copy_into_zeroed((char *)h_dst + off, c.pin + s * SLOT, len);
        else
            c.pool->copy((char *)h_dst + off, c.pin + s * SLOT, len);
        if (k + NSLOT < n) CUDA_TRY(issue(k + NSLOT));
    }
    return WAH_OK;
}

// A result the caller will free() (the reference's contract: compress.cu:181,208, decompress.cu:127,140).  calloc():
// a block this size comes straight from mmap, i.e. it IS zero without anybody writing to it, and the copy threads
// leave the parts of it that stay zero alone (CopyPool::copy_into_zeroed).
uint32_t *alloc_result(uint64_t words)
{
    const size_t bytes = (size_t)(words ? words : 1) * 4;
    const int st = result_strategy();
    char *p = (char *)(st >= 2 ? calloc(bytes, 1) : malloc(bytes));
    if (p && (st == 1 || st == 3) && bytes >= (8u << 20)) {
        // large blocks are advised towards huge pages: their first touch costs one fault per 2 MiB where the kernel allows it
        const uintptr_t lo = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
        const uintptr_t hi = ((uintptr_t)p + bytes) & ~(uintptr_t)((2u << 20) - 1);
        if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
    }
    return (uint32_t *)p;
}

// the malloc()ed result of compress() / decompress(), filled from the device
int fetch_result(HostCtx &c, uint32_t **h_out, const void *d_src, uint64_t words)
{
    const size_t bytes = (size_t)words * 4;
    if (sparse_copy() && bytes >= SPARSE_MIN) {
        const int rc = download_sparse(c, h_out, d_src, bytes);
        if (rc != WAH_OK && *h_out) {
            free(*h_out);
            *h_out = nullptr;
        }
        return rc;
    }
    uint32_t *host = alloc_result(words);
    if (!host) return wah_set_error(WAH_ERR_NOMEM, "malloc of %llu words failed", (unsigned long long)words);
    if (int rc = download(c, host, d_src, bytes, true)) {
        free(host);
        return rc;
    }
    *h_out = host;
    return WAH_OK;
}

}  // namespace

extern "C" void wah_free(void *p) { free(p); }

extern "C" void wah_host_last_transfer_bytes(uint64_t *h2d_bytes, uint64_t *d2h_bytes)
{
    std::lock_guard<std::mutex> g(g_ctx.mu);
    if (h2d_bytes) *h2d_bytes = g_ctx.h2d_bytes;
    if (d2h_bytes) *d2h_bytes = g_ctx.d2h_bytes;
}

extern "C" void wah_host_release(void)
{
    std::lock_guard<std::mutex> g(g_ctx.mu);
    g_ctx.release();
}

extern "C" int wah_compress_host(const uint32_t *h_in, uint64_t n_words, int mode, uint32_t **h_out,
                                 uint64_t *out_words, float *ms_h2d, float *ms_compute, float *ms_d2h)
{
    if (mode != WAH_BLOCK1024 && mode != WAH_CANONICAL) return wah_set_error(WAH_ERR_INVALID, "unknown mode %d", mode);
    if (!h_out) return wah_set_error(WAH_ERR_INVALID, "h_out is null");
    if (n_words && !h_in) return wah_set_error(WAH_ERR_INVALID, "h_in is null");
    *h_out = nullptr;
    std::lock_guard<std::mutex> g(g_ctx.mu);
    HostCtx &c = g_ctx;
    if (int rc = c.init()) return rc;
    c.h2d_bytes = c.d2h_bytes = 0;
    CUDA_TRY(cudaEventRecord(c.t0, c.stream));
    // -- segment 1: buffers (kept between calls) + H2D (compress.cu:57-120)
    const uint64_t cap = wah_max_compressed_words(n_words);
    const size_t ws_bytes = wah_compress_workspace_bytes(n_words);
    CUDA_TRY(c.a.reserve(n_words * 4 + 16));
    CUDA_TRY(c.b.reserve(cap * 4 + 16));
    CUDA_TRY(c.ws.reserve(ws_bytes));
    if (int rc = upload(c, c.a.p, h_in, n_words * 4)) return rc;
    const float t_h2d = c.lap();
    // -- segment 2: compute (compress.cu:125-172)
    uint64_t *d_cnt = (uint64_t *)c.small.p;
    if (int rc = wah_compress_device((const uint32_t *)c.a.p, n_words, mode, (uint32_t *)c.b.p, cap, d_cnt, c.ws.p,
                                     ws_bytes, c.stream))
        return rc;
    uint64_t cw = 0;
    CUDA_TRY(cudaMemcpyAsync(&cw, d_cnt, sizeof(cw), cudaMemcpyDeviceToHost, c.stream));
    const float t_compute = c.lap();   // drains the stream: cw is valid
    if (cw == ~0ull)
        return wah_set_error(WAH_ERR_CUDA, "the compress kernel gave up waiting for part of its grid (is another context holding the GPU?)");
    // -- segment 3: D2H into a malloc()ed buffer (compress.cu:177-202)
    uint32_t *host = nullptr;
    if (int rc = fetch_result(c, &host, c.b.p, cw)) return rc;
    const float t_d2h = c.lap();
    *h_out = host;
    if (out_words) *out_words = cw;
    if (ms_h2d) *ms_h2d = t_h2d;
    if (ms_compute) *ms_compute = t_compute;
    if (ms_d2h) *ms_d2h = t_d2h;
    return WAH_OK;
}

namespace {

// Decodes the stream in c.a into c.b and returns its size.  The reference sizes its output between two of its kernels
// (decompress.cu:72-97: two 8-byte copies and a scan, then cudaMalloc); here the output buffer of the previous call is
// still there, so one launch does both whenever that buffer is large enough -- the decoder reports the true size
// whatever the capacity -- and only a stream that needs more is decoded a second time into a larger buffer.
int decode_resident(HostCtx &c, uint64_t c_words, uint64_t want_cap, uint64_t *words_out)
{
    uint64_t *d_info = (uint64_t *)c.small.p;
    uint64_t info[3] = {0, 0, 0};
    uint64_t cap = c.b.cap / 4;
    if (want_cap != 0) {
        CUDA_TRY(c.b.reserve(want_cap * 4 + 16));
        cap = want_cap;
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        if (cap == 0) {
            // nothing to decode into yet: size query (the scan phase alone)
            const size_t ws0 = wah_decompress_workspace_bytes(c_words, 0);
            CUDA_TRY(c.ws.reserve(ws0));
            if (int rc = wah_decoded_size_device((const uint32_t *)c.a.p, c_words, d_info, c.ws.p, ws0, c.stream)) return rc;
        } else {
            const size_t ws_bytes = wah_decompress_workspace_bytes(c_words, cap);
            CUDA_TRY(c.ws.reserve(ws_bytes));
            if (int rc = wah_decompress_device((const uint32_t *)c.a.p, c_words, (uint32_t *)c.b.p, cap, d_info, c.ws.p,
                                               ws_bytes, c.stream))
                return rc;
        }
        CUDA_TRY(cudaMemcpyAsync(info, d_info, sizeof(info), cudaMemcpyDeviceToHost, c.stream));
        CUDA_TRY(cudaStreamSynchronize(c.stream));
        if (info[2] & WAH_STATUS_TIMEOUT)
            return wah_set_error(WAH_ERR_CUDA, "the decode kernel gave up waiting for part of its grid (is another context holding the GPU?)");
        if (info[2] & WAH_STATUS_BAD_WORDS_MASK)
            return wah_set_error(WAH_ERR_FORMAT, "%llu zero-length fill words in the stream",
                                 (unsigned long long)(info[2] & WAH_STATUS_BAD_WORDS_MASK));
        if (cap != 0 && info[0] <= cap) break;
        if (want_cap != 0) break;   // the caller's capacity is what it is
        CUDA_TRY(c.b.reserve(info[0] * 4 + 16));
        cap = c.b.cap / 4;
    }
    *words_out = info[0];
    return WAH_OK;
}

}  // namespace

extern "C" int wah_decompress_host(const uint32_t *h_in, uint64_t c_words, uint32_t **h_out,
                                   uint64_t *out_words, float *ms_h2d, float *ms_compute, float *ms_d2h)
{
    if (!h_out) return wah_set_error(WAH_ERR_INVALID, "h_out is null");
    if (c_words && !h_in) return wah_set_error(WAH_ERR_INVALID, "h_in is null");
    *h_out = nullptr;
    std::lock_guard<std::mutex> g(g_ctx.mu);
    HostCtx &c = g_ctx;
    if (int rc = c.init()) return rc;
    c.h2d_bytes = c.d2h_bytes = 0;
    CUDA_TRY(cudaEventRecord(c.t0, c.stream));
    // -- segment 1: buffers + H2D (decompress.cu:34-56)
    CUDA_TRY(c.a.reserve(c_words * 4 + 16));
    if (int rc = upload(c, c.a.p, h_in, c_words * 4)) return rc;
    const float t_h2d = c.lap();
    // -- segment 2: output size, output buffer, expansion (decompress.cu:66-124)
    uint64_t words = 0;
    if (c_words != 0) {
        if (int rc = decode_resident(c, c_words, 0, &words)) return rc;
    }
    const float t_compute = c.lap();
    // -- segment 3: D2H into a malloc()ed buffer (decompress.cu:127-133)
    uint32_t *host = nullptr;
    if (int rc = fetch_result(c, &host, c.b.p, words)) return rc;
    const float t_d2h = c.lap();
    *h_out = host;
    if (out_words) *out_words = words;
    if (ms_h2d) *ms_h2d = t_h2d;
    if (ms_compute) *ms_compute = t_compute;
    if (ms_d2h) *ms_d2h = t_d2h;
    return WAH_OK;
}

// caller-provided result buffers (page-locked ones are written by DMA directly)
extern "C" int wah_compress_host_into(const uint32_t *h_in, uint64_t n_words, int mode, uint32_t *h_out,
                                      uint64_t out_capacity_words, uint64_t *out_words)
{
    if (mode != WAH_BLOCK1024 && mode != WAH_CANONICAL) return wah_set_error(WAH_ERR_INVALID, "unknown mode %d", mode);
    if (!out_words || (out_capacity_words && !h_out)) return wah_set_error(WAH_ERR_INVALID, "null output");
    if (n_words && !h_in) return wah_set_error(WAH_ERR_INVALID, "h_in is null");
    std::lock_guard<std::mutex> g(g_ctx.mu);
    HostCtx &c = g_ctx;
    if (int rc = c.init()) return rc;
    c.h2d_bytes = c.d2h_bytes = 0;
    const uint64_t cap = wah_max_compressed_words(n_words);
    const size_t ws_bytes = wah_compress_workspace_bytes(n_words);
    CUDA_TRY(c.a.reserve(n_words * 4 + 16));
    CUDA_TRY(c.b.reserve(cap * 4 + 16));
    CUDA_TRY(c.ws.reserve(ws_bytes));
    if (int rc = upload(c, c.a.p, h_in, n_words * 4)) return rc;
    uint64_t *d_cnt = (uint64_t *)c.small.p;
    if (int rc = wah_compress_device((const uint32_t *)c.a.p, n_words, mode, (uint32_t *)c.b.p, cap, d_cnt, c.ws.p,
                                     ws_bytes, c.stream))
        return rc;
    uint64_t cw = 0;
    CUDA_TRY(cudaMemcpyAsync(&cw, d_cnt, sizeof(cw), cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    if (cw == ~0ull)
        return wah_set_error(WAH_ERR_CUDA, "the compress kernel gave up waiting for part of its grid (is another context holding the GPU?)");
    *out_words = cw;
    if (cw > out_capacity_words)
        return wah_set_error(WAH_ERR_CAPACITY, "result needs %llu words, buffer holds %llu", (unsigned long long)cw,
                             (unsigned long long)out_capacity_words);
    return download(c, h_out, c.b.p, cw * 4, false);
}

extern "C" int wah_decompress_host_into(const uint32_t *h_in, uint64_t c_words, uint32_t *h_out,
                                        uint64_t out_capacity_words, uint64_t *out_words)
{
    if (!out_words || (out_capacity_words && !h_out)) return wah_set_error(WAH_ERR_INVALID, "null output");
    if (c_words && !h_in) return wah_set_error(WAH_ERR_INVALID, "h_in is null");
    std::lock_guard<std::mutex> g(g_ctx.mu);
    HostCtx &c = g_ctx;
    if (int rc = c.init()) return rc;
    c.h2d_bytes = c.d2h_bytes = 0;
    CUDA_TRY(c.a.reserve(c_words * 4 + 16));
    if (int rc = upload(c, c.a.p, h_in, c_words * 4)) return rc;
    // the caller's capacity bounds the output, so one pass does both the size and the expansion
    uint64_t words = 0;
    if (c_words != 0 && out_capacity_words != 0) {
        if (int rc = decode_resident(c, c_words, out_capacity_words, &words)) return rc;
    } else if (c_words != 0) {
        uint64_t *d_info = (uint64_t *)c.small.p;
        uint64_t info[3] = {0, 0, 0};
        const size_t ws0 = wah_decompress_workspace_bytes(c_words, 0);
        CUDA_TRY(c.ws.reserve(ws0));
        if (int rc = wah_decoded_size_device((const uint32_t *)c.a.p, c_words, d_info, c.ws.p, ws0, c.stream)) return rc;
        CUDA_TRY(cudaMemcpyAsync(info, d_info, sizeof(info), cudaMemcpyDeviceToHost, c.stream));
        CUDA_TRY(cudaStreamSynchronize(c.stream));
        words = info[0];
    }
    *out_words = words;
    if (words > out_capacity_words)
        return wah_set_error(WAH_ERR_CAPACITY, "result needs %llu words, buffer holds %llu",
                             (unsigned long long)words, (unsigned long long)out_capacity_words);
    return download(c, h_out, c.b.p, words * 4, false);
}
