// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header).
// Persistent worker pool standing in for the reference's ThreadPool
// (src/controller/concurrency.hpp:30-216): `threads` long-lived workers, one task per worker
// per update, the caller blocks until all have finished (the futures barrier of
// mppi.cpp:305-306). Kept separate from /root/reference so the oracle builds on the GPU box.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace oracle {

class Pool {
public:
    explicit Pool(unsigned n) : m_tasks(n), m_pending(0) {
        for (unsigned i = 0; i < n; i++) m_workers.emplace_back([this, i] { run(i); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(m_mutex); m_stop = true; }
        m_wake.notify_all();
        for (auto &w : m_workers) w.join();
    }
    // Hand worker `i` a task for this round (at most one each), then wait().
    void submit(unsigned i, std::function<void()> f) {
        std::lock_guard<std::mutex> l(m_mutex);
        m_tasks[i] = std::move(f);
        ++m_pending;
    }
    void launch() { m_wake.notify_all(); }
    void wait() {
        std::unique_lock<std::mutex> l(m_mutex);
        m_done.wait(l, [this] { return m_pending == 0; });
    }
    unsigned size() const { return (unsigned)m_workers.size(); }

private:
    void run(unsigned i) {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_mutex);
                m_wake.wait(l, [&] { return m_stop || (bool)m_tasks[i]; });
                if (m_stop) return;
                f = std::move(m_tasks[i]);
                m_tasks[i] = nullptr;
            }
            f();
            {
                std::lock_guard<std::mutex> l(m_mutex);
                --m_pending;
            }
            m_done.notify_all();
        }
    }
    std::vector<std::thread> m_workers;
    std::vector<std::function<void()>> m_tasks;
    std::mutex m_mutex;
    std::condition_variable m_wake, m_done;
    unsigned m_pending;
    bool m_stop = false;
};

}  // namespace oracle
