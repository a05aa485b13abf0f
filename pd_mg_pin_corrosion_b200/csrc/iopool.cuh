// iopool.cuh -- pinned staging pool + writer threads shared by the output paths (vti.cu,
// checkpoint.cu): data leaves the device in chunks through a few pinned buffers; writer threads put
// every chunk at its final file offset with pwrite, so the D2H copy of one chunk overlaps the
// page-cache copies of others.
#pragma once
#include <fcntl.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace pdio {

constexpr size_t kIoChunk = (size_t)32 << 20;
constexpr int kIoBufs = 6, kIoWriters = 3;
struct IoPool {
    char* buf[kIoBufs] = {nullptr};
    cudaEvent_t ev[kIoBufs] = {nullptr};
    std::mutex busy;   // one writer run / checkpoint load per device at a time (contexts on several host threads)
};
inline IoPool g_io_pool[64];

inline int io_pool(pdgpu_ctx* c, IoPool** out) {
    IoPool* p = &g_io_pool[c->device < 64 ? c->device : 0];
    std::lock_guard<std::mutex> guard(p->busy);
    for (int k = 0; k < kIoBufs; ++k) {
        if (!p->buf[k]) CUDA_OK(cudaMallocHost(&p->buf[k], kIoChunk));
        if (!p->ev[k]) CUDA_OK(cudaEventCreateWithFlags(&p->ev[k], cudaEventDisableTiming));
    }
    *out = p;
    return 0;
}

struct IoJob { int k; size_t len; off_t off; };
class IoRun {
   public:
    IoRun(IoPool* io, int fd, int device) : io_(io), fd_(fd), device_(device) {
        if (!io_) { err_ = true; return; }
        io_->busy.lock();
        locked_ = true;
        for (int k = 0; k < kIoBufs; ++k) free_.push_back(k);
        for (int w = 0; w < kIoWriters; ++w) th_.emplace_back([this] { work(); });
    }
    ~IoRun() { finish(); }
    int acquire() {                               // a free staging buffer (blocks), -1 after a failure
        std::unique_lock<std::mutex> l(m_);
        cv_free_.wait(l, [this] { return !free_.empty() || err_; });
        if (err_) return -1;
        int k = free_.back();
        free_.pop_back();
        return k;
    }
    void submit(int k, size_t len, off_t off) {
        { std::lock_guard<std::mutex> l(m_); jobs_.push_back({k, len, off}); }
        cv_job_.notify_one();
    }
    void fail() { { std::lock_guard<std::mutex> l(m_); err_ = true; } cv_free_.notify_all(); }
    bool failed() { std::lock_guard<std::mutex> l(m_); return err_; }
    void finish() {
        { std::lock_guard<std::mutex> l(m_); done_ = true; }
        cv_job_.notify_all();
        for (std::thread& t : th_) if (t.joinable()) t.join();
        th_.clear();
        if (locked_) { io_->busy.unlock(); locked_ = false; }
    }

   private:
    void work() {
        cudaSetDevice(device_);
        for (;;) {
            IoJob j;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_job_.wait(l, [this] { return !jobs_.empty() || done_; });
                if (jobs_.empty()) return;
                j = jobs_.front();
                jobs_.pop_front();
            }
            bool ok = cudaEventSynchronize(io_->ev[j.k]) == cudaSuccess;
            size_t w = 0;
            while (ok && w < j.len) {
                ssize_t r = ::pwrite(fd_, io_->buf[j.k] + w, j.len - w, j.off + (off_t)w);
                if (r <= 0) ok = false; else w += (size_t)r;
            }
            {
                std::lock_guard<std::mutex> l(m_);
                if (!ok) err_ = true;
                free_.push_back(j.k);
            }
            cv_free_.notify_all();
        }
    }
    IoPool* io_;
    int fd_, device_;
    std::mutex m_;
    std::condition_variable cv_job_, cv_free_;
    std::deque<IoJob> jobs_;
    std::vector<int> free_;
    std::vector<std::thread> th_;
    bool done_ = false, err_ = false, locked_ = false;
};


}  // namespace pdio
