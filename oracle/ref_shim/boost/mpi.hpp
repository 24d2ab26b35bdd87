// TEST INFRASTRUCTURE ONLY (oracle/): a stand-in for <boost/mpi.hpp> so that the reference's own
// translation units can be compiled in place from /root/reference without Boost.MPI or an MPI
// library.  "Ranks" are threads of one process; messages are buffered copies passed through an
// in-process mailbox keyed by (source, destination, tag).  Only the subset of the Boost.MPI API the
// reference calls is provided (communicator::{rank,size,barrier,isend,irecv,send,recv}, request,
// wait_all, broadcast, gather, all_gather, environment).  Nothing here is product code.
#ifndef CHIML_ORACLE_SHIM_BOOST_MPI_HPP
#define CHIML_ORACLE_SHIM_BOOST_MPI_HPP

#include <algorithm>
#include <array>
#include <complex>
#include <condition_variable>
#include <iostream>
#include <numeric>
#include <string>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace boost { namespace mpi {

namespace shim {

struct World
{
    int nranks = 1;
    std::mutex mtx;
    std::condition_variable cv;
    // mailbox: (src, dst, tag) -> queue of type-erased payloads
    std::map<std::tuple<int,int,int>, std::deque<std::shared_ptr<void>>> box;
    // barrier state
    int barCount = 0;
    long barGen = 0;
};

inline World& world() { static World w; return w; }
inline int& myRank() { static thread_local int r = 0; return r; }

inline void post(int src, int dst, int tag, std::shared_ptr<void> payload)
{
    World& w = world();
    {
        std::lock_guard<std::mutex> lk(w.mtx);
        w.box[std::make_tuple(src, dst, tag)].push_back(std::move(payload));
    }
    w.cv.notify_all();
}

inline std::shared_ptr<void> take(int src, int dst, int tag)
{
    World& w = world();
    std::unique_lock<std::mutex> lk(w.mtx);
    auto key = std::make_tuple(src, dst, tag);
    w.cv.wait(lk, [&]{ auto it = w.box.find(key); return it != w.box.end() && !it->second.empty(); });
    auto& q = w.box[key];
    std::shared_ptr<void> p = q.front();
    q.pop_front();
    return p;
}

inline void barrier()
{
    World& w = world();
    std::unique_lock<std::mutex> lk(w.mtx);
    long gen = w.barGen;
    if(++w.barCount == w.nranks)
    {
        w.barCount = 0;
        ++w.barGen;
        w.cv.notify_all();
    }
    else
        w.cv.wait(lk, [&]{ return w.barGen != gen; });
}

// tags reserved for the collectives (the reference's own tags are non-negative)
enum { TAG_BCAST = -101, TAG_GATHER = -102, TAG_ALLGATHER = -103 };

} // namespace shim

class environment
{
public:
    environment() {}
    environment(int&, char**&) {}
};

class request
{
public:
    std::function<void()> complete_; // empty => already complete
    request() {}
    void wait() { if(complete_) { complete_(); complete_ = nullptr; } }
};

class communicator
{
public:
    communicator() {}
    int rank() const { return shim::myRank(); }
    int size() const { return shim::world().nranks; }
    void barrier() const { shim::barrier(); }

    // pointer + count flavour (contiguous PODs)
    template <typename T> request isend(int dest, int tag, const T* values, int n) const
    {
        auto buf = std::make_shared<std::vector<T>>(values, values + n);
        shim::post(rank(), dest, tag, buf);
        return request();
    }
    template <typename T> request irecv(int source, int tag, T* values, int n) const
    {
        request r;
        int me = rank();
        r.complete_ = [source, me, tag, values, n]() {
            auto p = std::static_pointer_cast<std::vector<T>>(shim::take(source, me, tag));
            if(int(p->size()) != n) throw std::runtime_error("mpi shim: message size mismatch");
            std::copy(p->begin(), p->end(), values);
        };
        return r;
    }
    template <typename T> void send(int dest, int tag, const T* values, int n) const { isend(dest, tag, values, n); }
    template <typename T> void recv(int source, int tag, T* values, int n) const { irecv(source, tag, values, n).wait(); }

    // whole-object flavour (what Boost would serialise)
    template <typename T> request isend(int dest, int tag, const T& value) const
    {
        shim::post(rank(), dest, tag, std::make_shared<T>(value));
        return request();
    }
    template <typename T> request irecv(int source, int tag, T& value) const
    {
        request r;
        int me = rank();
        T* out = &value;
        r.complete_ = [source, me, tag, out]() { *out = *std::static_pointer_cast<T>(shim::take(source, me, tag)); };
        return r;
    }
    template <typename T> void send(int dest, int tag, const T& value) const { isend(dest, tag, value); }
    template <typename T> void recv(int source, int tag, T& value) const { irecv(source, tag, value).wait(); }
};

template <typename It> void wait_all(It first, It last) { for(; first != last; ++first) first->wait(); }

template <typename T> void broadcast(const communicator& comm, T& value, int root)
{
    if(comm.size() == 1) return;
    if(comm.rank() == root)
    {
        for(int r = 0; r < comm.size(); ++r)
            if(r != root) shim::post(root, r, shim::TAG_BCAST, std::make_shared<T>(value));
    }
    else
        value = *std::static_pointer_cast<T>(shim::take(root, comm.rank(), shim::TAG_BCAST));
    comm.barrier();
}

template <typename T> void broadcast(const communicator& comm, T* values, int n, int root)
{
    if(comm.size() == 1) return;
    if(comm.rank() == root)
    {
        for(int r = 0; r < comm.size(); ++r)
            if(r != root) shim::post(root, r, shim::TAG_BCAST, std::make_shared<std::vector<T>>(values, values + n));
    }
    else
    {
        auto p = std::static_pointer_cast<std::vector<T>>(shim::take(root, comm.rank(), shim::TAG_BCAST));
        std::copy(p->begin(), p->end(), values);
    }
    comm.barrier();
}

template <typename T> void gather(const communicator& comm, const T& in, std::vector<T>& out, int root)
{
    if(comm.rank() == root)
    {
        out.resize(comm.size());
        for(int r = 0; r < comm.size(); ++r)
        {
            if(r == root) out[r] = in;
            else out[r] = *std::static_pointer_cast<T>(shim::take(r, root, shim::TAG_GATHER));
        }
    }
    else
        shim::post(comm.rank(), root, shim::TAG_GATHER, std::make_shared<T>(in));
    comm.barrier();
}

template <typename T> void gather(const communicator& comm, const T& in, int root)
{
    if(comm.rank() == root) throw std::runtime_error("mpi shim: root must pass an output vector to gather");
    shim::post(comm.rank(), root, shim::TAG_GATHER, std::make_shared<T>(in));
    comm.barrier();
}

template <typename T> void all_gather(const communicator& comm, const T& in, std::vector<T>& out)
{
    int n = comm.size();
    int me = comm.rank();
    for(int r = 0; r < n; ++r)
        if(r != me) shim::post(me, r, shim::TAG_ALLGATHER, std::make_shared<T>(in));
    out.resize(n);
    for(int r = 0; r < n; ++r)
    {
        if(r == me) out[r] = in;
        else out[r] = *std::static_pointer_cast<T>(shim::take(r, me, shim::TAG_ALLGATHER));
    }
    comm.barrier();
}

}} // namespace boost::mpi

#endif
