// Stand-in for the Boost subset the reference uses. Written for this repository; see the README.md of oracle/refshim.
// boost::asio::io_service as the reference's logger uses it (basic.hpp:106-264): post() queues a handler, run() drains the
// queue on the logger thread until the last io_service::work object is destroyed and the queue is empty.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <sys/ioctl.h>   // the reference relies on asio pulling in ioctl / winsize / STDOUT_FILENO (basic.hpp:160-170)
#include <unistd.h>
namespace boost { namespace asio {
class io_service {
public:
    class work {
    public:
        explicit work(io_service& io) : io_(io) { std::lock_guard<std::mutex> l(io_.m_); io_.work_++; }
        ~work() { { std::lock_guard<std::mutex> l(io_.m_); io_.work_--; } io_.cv_.notify_all(); }
    private:
        io_service& io_;
    };
    template <class F> void post(F f) { { std::lock_guard<std::mutex> l(m_); q_.emplace_back(std::move(f)); } cv_.notify_one(); }
    std::size_t run() {
        std::size_t n = 0;
        std::unique_lock<std::mutex> l(m_);
        while (true) {
            cv_.wait(l, [this] { return !q_.empty() || work_ == 0; });
            if (q_.empty()) break;
            std::function<void()> f = std::move(q_.front());
            q_.pop_front();
            l.unlock(); f(); n++; l.lock();
        }
        return n;
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    int work_ = 0;
};
}}
