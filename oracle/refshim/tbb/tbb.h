// Stand-in for the TBB subset the reference uses (basic.hpp:128-135, src/nanogi.cpp:83,229,281). Written for this repository;
// see ../README.md. parallel_for runs the grains of a blocked_range on std::threads pulling from an atomic cursor.
#pragma once
#include <atomic>
#include <list>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace tbb {

namespace detail { inline int& num_threads() { static int n = 0; return n; } }

class task_scheduler_init {
public:
    static const int deferred = -2;
    explicit task_scheduler_init(int n = -1) { if (n > 0) detail::num_threads() = n; }
    void initialize(int n) { if (n > 0) detail::num_threads() = n; }
};

template <class T> class blocked_range {
public:
    blocked_range(T b, T e, T grain = 1) : b_(b), e_(e), g_(grain < 1 ? 1 : grain) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    T grainsize() const { return g_; }
private:
    T b_, e_, g_;
};

template <class T, class F> void parallel_for(const blocked_range<T>& r, const F& f) {
    int n = detail::num_threads();
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    std::atomic<T> next(r.begin());
    auto worker = [&]() {
        while (true) {
            const T b = next.fetch_add(r.grainsize());
            if (b >= r.end()) break;
            const T e = b + r.grainsize() < r.end() ? b + r.grainsize() : r.end();
            f(blocked_range<T>(b, e, r.grainsize()));
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

template <class T> class enumerable_thread_specific {
public:
    typedef typename std::list<T>::iterator iterator;
    T& local() {
        std::lock_guard<std::mutex> l(m_);
        const auto id = std::this_thread::get_id();
        auto it = idx_.find(id);
        if (it != idx_.end()) return *it->second;
        items_.emplace_back();
        idx_[id] = &items_.back();
        return items_.back();
    }
    iterator begin() { return items_.begin(); }
    iterator end() { return items_.end(); }
    template <class F> void combine_each(F f) { for (auto& x : items_) f(x); }
private:
    std::mutex m_;
    std::list<T> items_;
    std::map<std::thread::id, T*> idx_;
};

template <class K, class V> class concurrent_hash_map {
public:
    class accessor {
    public:
        std::pair<const K, V>* operator->() const { return p_; }
        std::pair<const K, V>* p_ = nullptr;
        std::unique_lock<std::mutex> lock_;
    };
    bool insert(accessor& a, const K& key) {
        a.lock_ = std::unique_lock<std::mutex>(m_);
        auto r = map_.emplace(key, V());
        a.p_ = &*r.first;
        return r.second;
    }
private:
    std::mutex m_;
    std::unordered_map<K, V> map_;
};

}  // namespace tbb
