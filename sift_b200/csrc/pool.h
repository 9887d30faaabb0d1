// Worker pool of the host side of the pipeline (per-image order replay, frame packing).  Plain C++11 threads, no CUDA.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace siftgpu {

// One job at a time: begin() hands fn(0..n-1) to the workers and returns, end() lets the caller take what is left and waits
// for the rest.  Items are drawn from a ticket that carries the job's generation next to the item index, so a worker that
// is still leaving the previous job (descheduled between two instructions for however long) can never draw, run or count
// an item of the next one.  tests/native/pool_check.cpp hammers this with back-to-back jobs.
class Pool {
   public:
    explicit Pool(int n) {
        for (int i = 0; i < n; ++i) workers.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(m); stop = true; }
        cv.notify_all();
        for (auto& t : workers) t.join();
    }
    void begin(int n, std::function<void(int)> fn) {
        held = std::move(fn);  // no worker touches `held` between jobs: it only runs items it drew from the current ticket
        held_n = n;
        if (n <= 0 || workers.empty()) return;
        {
            std::lock_guard<std::mutex> g(m);
            ++gen;
            total = n;
            pending.store(n);
            ticket.store((uint64_t)gen << 32);
            active = true;
        }
        cv.notify_all();
    }
    void end() {
        const int n = held_n;
        held_n = 0;
        if (n <= 0) return;
        if (workers.empty()) { for (int i = 0; i < n; ++i) held(i); return; }
        work(gen, n);  // only this thread ever changes gen
        std::unique_lock<std::mutex> lk(m);
        done_cv.wait(lk, [this] { return pending.load() == 0; });
        active = false;
    }
    // Between begin() and end(): no further items are started.  Returns how many were (items 0 .. result-1 run to completion, the
    // caller's end() waits for the ones still in flight and then runs nothing itself).
    int cancel_rest() {
        const int n = held_n;
        if (n <= 0) return 0;
        if (workers.empty()) { held_n = 0; return 0; }
        const uint64_t exhausted = ((uint64_t)gen << 32) | (uint32_t)n;
        uint64_t cur = ticket.load();
        for (;;) {
            const int drawn = (int)(uint32_t)cur;
            if (drawn >= n) return n;
            if (!ticket.compare_exchange_weak(cur, exhausted)) continue;  // cur reloaded
            const int rest = n - drawn;
            if (pending.fetch_sub(rest) == rest) { std::lock_guard<std::mutex> lk(m); done_cv.notify_all(); }
            return drawn;
        }
    }
    void parallel_for(int n, const std::function<void(int)>& fn) {
        begin(n, fn);
        end();
    }

   private:
    // runs items of generation g until the job is exhausted or another job has taken the ticket over
    void work(uint32_t g, int n) {
        uint64_t cur = ticket.load();
        for (;;) {
            if ((uint32_t)(cur >> 32) != g || (int)(uint32_t)cur >= n) return;
            if (!ticket.compare_exchange_weak(cur, cur + 1)) continue;  // cur reloaded
            held((int)(uint32_t)cur);
            if (pending.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(m); done_cv.notify_all(); }
            cur = ticket.load();
        }
    }
    void loop() {
        uint32_t seen = 0;  // the last generation this worker took part in
        for (;;) {
            uint32_t g;
            int n;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || (active && gen != seen); });
                if (stop) return;
                g = gen; n = total; seen = g;
            }
            work(g, n);
        }
    }
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv, done_cv;
    std::function<void(int)> held;
    int held_n = 0;
    uint32_t gen = 0;     // guarded by m (written by the owner thread only)
    int total = 0;        // guarded by m
    bool active = false;  // guarded by m
    std::atomic<uint64_t> ticket{0};  // generation << 32 | next item
    std::atomic<int> pending{0};
    bool stop = false;
};

}  // namespace siftgpu
