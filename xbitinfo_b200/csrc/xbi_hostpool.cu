// Host side of the host-buffer entry points: a small pool of threads that moves slabs between
// the caller's (pageable) numpy memory and the library's pinned staging ring.
//
// cudaMemcpyAsync from or to pageable memory is staged by the driver on one thread (measured on
// the B200 boxes: 11 GB/s host->device, 4 GB/s for a fresh output array, against 55 GB/s from pinned
// memory).  Copying the slab into pinned memory with several threads first -- and out of it
// afterwards, which also spreads the first-touch page faults of a fresh output array over the
// threads -- leaves the DMA engine with pinned buffers only.
#include <pthread.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "xbi_internal.h"

namespace xbi {

namespace {

struct CopyJob {
    char* dst;
    const char* src;
    size_t dst_pitch, src_pitch, row_bytes, rows;
};

class CopyPool {
  public:
    explicit CopyPool(int n) : n_(n) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { worker(i); });
    }
    int size() const { return n_; }

    // blocks until the whole job is copied; the calling thread takes part 0
    void run(const CopyJob& job) {
        std::lock_guard<std::mutex> serial(run_mutex_);  // one job at a time
        if (n_ == 1) {
            copy_part(job, 0, 1);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = job;
            remaining_ = n_ - 1;
            ++generation_;
        }
        wake_.notify_all();
        copy_part(job, 0, n_);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return remaining_ == 0; });
    }

  private:
    static void copy_part(const CopyJob& j, int part, int parts) {
        if (j.rows == 1 || (j.dst_pitch == j.row_bytes && j.src_pitch == j.row_bytes)) {
            // contiguous: split the byte range on 4 KiB boundaries (whole pages per thread)
            const size_t total = j.row_bytes * j.rows;
            const size_t pages = (total + 4095) / 4096;
            const size_t lo = (pages * part / parts) * 4096, hi_p = (pages * (part + 1) / parts) * 4096;
            const size_t hi = hi_p < total ? hi_p : total;
            if (hi > lo) std::memcpy(j.dst + lo, j.src + lo, hi - lo);
        } else {
            const size_t lo = j.rows * part / parts, hi = j.rows * (part + 1) / parts;
            for (size_t r = lo; r < hi; ++r) std::memcpy(j.dst + r * j.dst_pitch, j.src + r * j.src_pitch, j.row_bytes);
        }
    }

    void worker(int id) {
        unsigned long long seen = 0;
        for (;;) {
            CopyJob job;
            {
                std::unique_lock<std::mutex> lk(m_);
                wake_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                job = job_;
            }
            copy_part(job, id, n_);
            {
                std::lock_guard<std::mutex> lk(m_);
                --remaining_;
            }
            done_.notify_one();
        }
    }

    const int n_;
    std::vector<std::thread> workers_;
    std::mutex run_mutex_, m_;
    std::condition_variable wake_, done_;
    CopyJob job_{};
    unsigned long long generation_ = 0;
    int remaining_ = 0;
};

// Leaked on purpose: the workers block on a condition variable for the life of the process, and a
// forked child (which inherits none of the threads) simply starts a pool of its own.
CopyPool* g_pool = nullptr;
std::mutex g_pool_mutex;
void forget_pool_in_child() {
    g_pool = nullptr;
    new (&g_pool_mutex) std::mutex();
}

int configured_threads() {
    if (const char* env = std::getenv("XBI_HOST_THREADS")) {
        const long v = std::atol(env);
        if (v >= 1 && v <= 64) return (int)v;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const int n = hw == 0 ? 4 : (int)hw;
    return n > 8 ? 8 : n;
}

CopyPool& pool() {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    if (!g_pool) {
        static bool atfork_installed = false;
        if (!atfork_installed) {
            pthread_atfork(nullptr, nullptr, forget_pool_in_child);
            atfork_installed = true;
        }
        g_pool = new CopyPool(configured_threads());
    }
    return *g_pool;
}

}  // namespace

int host_copy_threads() { return pool().size(); }

void host_copy_parallel(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows) {
    if (row_bytes == 0 || rows == 0) return;
    CopyJob job{static_cast<char*>(dst), static_cast<const char*>(src), dst_pitch, src_pitch, row_bytes, rows};
    if (row_bytes * rows < (size_t)1 << 20) {  // not worth waking anybody up
        for (size_t r = 0; r < rows; ++r) std::memcpy(job.dst + r * dst_pitch, job.src + r * src_pitch, row_bytes);
        return;
    }
    pool().run(job);
}

}  // namespace xbi
