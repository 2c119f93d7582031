// TEST INFRASTRUCTURE ONLY.
//
// A thread-backed stand-in for the handful of boost::mpi calls the reference's parallel-tempering driver uses
// (detqmcpt.h:215-217, 282-299, 705-735, 813-835, 874, 932, 967-1012, 1092-1105, 1140-1145;
// mpiobservablehandlerpt.h:153-155, mpiobservablehandlerpt.cpp:79-83, 177-197, 222-224, 260-262):
// communicator::rank / size / barrier, gather (value and pointer forms), scatter, broadcast.  The development
// container has no MPI, so the reference's DetQMCPT cannot be built as shipped; with this header every "process" of the
// ladder is a thread of one process (oracle/ref_pt_harness.cpp) and the UNMODIFIED detqmcpt.h / mpiobservablehandlerpt.*
// run the replica-exchange simulation that the goldens of tests/golden/pt_reference.npz come from.
// It is force-included with -DBOOST_MPI_HPP so that the vendored boost/mpi.hpp (which needs <mpi.h>) stays out.
#ifndef FAKE_BOOST_MPI_HPP_
#define FAKE_BOOST_MPI_HPP_

#include <condition_variable>
#include <mutex>
#include <vector>

namespace boost { namespace mpi {

struct fake_world {
    int size = 1;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    long generation = 0;
    std::vector<const void*> slot;          // one pointer per rank: the argument a collective publishes
    static fake_world& get() { static fake_world w; return w; }
    static int& rank() { static thread_local int r = 0; return r; }
    void init(int n) { size = n; slot.assign(n, nullptr); waiting = 0; generation = 0; }
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const long gen = generation;
        if (++waiting == size) {
            waiting = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen; });
        }
    }
};

class communicator {
public:
    int rank() const { return fake_world::rank(); }
    int size() const { return fake_world::get().size; }
    void barrier() const { fake_world::get().barrier(); }
};

// every rank publishes a pointer to its contribution, the root copies between two barriers
template <class T>
void gather(const communicator& c, const T& in, std::vector<T>& out, int root) {
    fake_world& w = fake_world::get();
    w.slot[c.rank()] = &in;
    w.barrier();
    if (c.rank() == root) {
        out.resize(w.size);
        for (int r = 0; r < w.size; ++r) out[r] = *static_cast<const T*>(w.slot[r]);
    }
    w.barrier();
}
template <class T>
void gather(const communicator& c, const T* in, int n, std::vector<T>& out, int root) {
    fake_world& w = fake_world::get();
    w.slot[c.rank()] = in;
    w.barrier();
    if (c.rank() == root) {
        out.resize(size_t(w.size) * n);
        for (int r = 0; r < w.size; ++r)
            for (int i = 0; i < n; ++i) out[size_t(r) * n + i] = static_cast<const T*>(w.slot[r])[i];
    }
    w.barrier();
}
template <class T>
void scatter(const communicator& c, const std::vector<T>& in, T& out, int root) {
    fake_world& w = fake_world::get();
    if (c.rank() == root) w.slot[root] = &in;
    w.barrier();
    out = (*static_cast<const std::vector<T>*>(w.slot[root]))[c.rank()];
    w.barrier();
}
template <class T>
void broadcast(const communicator& c, T& value, int root) {
    fake_world& w = fake_world::get();
    if (c.rank() == root) w.slot[root] = &value;
    w.barrier();
    if (c.rank() != root) value = *static_cast<const T*>(w.slot[root]);
    w.barrier();
}

}}  // namespace boost::mpi

#endif
