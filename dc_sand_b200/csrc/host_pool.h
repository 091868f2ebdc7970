// A few persistent host threads for the byte-moving parts of the host path (pageable -> pinned staging, complex64 ->
// complex128 widening): starting std::threads per 8 MB piece cost 3 x 30-50 us each time, a parked worker wakes in a few.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace ddch {

class HostPool {
public:
    explicit HostPool(int workers) {
        for (int i = 0; i < workers; ++i) th_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    HostPool(const HostPool&) = delete;
    HostPool& operator=(const HostPool&) = delete;
    int workers() const { return (int)th_.size(); }

    // fn(part) for part = 0 .. parts-1, spread over the workers and the calling thread; returns when all are done.
    // One job at a time (the host path of a handle is single-threaded by contract).
    void run(int parts, const std::function<void(int)>& fn) {
        if (parts <= 0) return;
        if (parts == 1 || th_.empty()) {
            for (int i = 0; i < parts; ++i) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            parts_ = parts;
            next_ = 0;
            left_ = parts;
            ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return left_ == 0; });
        fn_ = nullptr;   // under the mutex: no worker can pick up a part of a finished job
    }

private:
    // every piece of job state is read and written under the mutex (parts are ~1 MB copies: the lock is noise), so a worker
    // that comes late to one job can only ever join the job that is current
    void work() {
        for (;;) {
            const std::function<void(int)>* fn;
            int i;
            {
                std::lock_guard<std::mutex> g(m_);
                if (fn_ == nullptr || next_ >= parts_) return;
                i = next_++;
                fn = fn_;
            }
            (*fn)(i);
            std::lock_guard<std::mutex> g(m_);
            if (--left_ == 0) done_.notify_all();
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || (gen_ != seen && fn_ != nullptr); });
                if (stop_) return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_ = 0, parts_ = 0, left_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

}  // namespace ddch
