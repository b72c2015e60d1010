// Stub of boost/mpi/communicator.hpp for the oracle build (test infrastructure only).
// One std::thread per rank inside one process; messages are heap copies passed
// through a mutex + condition-variable mailbox keyed by (src, dst, tag).
#pragma once
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
namespace boost { namespace mpi {
struct shim_world {
    std::mutex mu;
    std::condition_variable cv;
    std::map<std::tuple<int, int, int>, std::deque<std::shared_ptr<void>>> box;
    int size = 1;
    static shim_world& get() { static shim_world w; return w; }
};
inline int& shim_rank() { static thread_local int r = 0; return r; }
class communicator {
  public:
    int rank() const { return shim_rank(); }
    int size() const { return shim_world::get().size; }
    template <class T> void send(int dst, int tag, T const& v) const {
        auto& w = shim_world::get();
        std::unique_lock<std::mutex> lk(w.mu);
        w.box[{shim_rank(), dst, tag}].push_back(std::make_shared<T>(v));
        w.cv.notify_all();
    }
    template <class T> void recv(int src, int tag, T& v) const {
        auto& w = shim_world::get();
        std::unique_lock<std::mutex> lk(w.mu);
        auto key = std::make_tuple(src, shim_rank(), tag);
        w.cv.wait(lk, [&] { auto it = w.box.find(key); return it != w.box.end() && !it->second.empty(); });
        auto& q = w.box[key];
        v = *std::static_pointer_cast<T>(q.front());
        q.pop_front();
    }
};
} }
