// TEST INFRASTRUCTURE (CPU suite only): a thread-rendezvous stand-in for the few NCCL entry points csrc/dfd_exchange.cu
// resolves with dlopen("libnccl.so.2") — built as libnccl.so.2 into pytest's temporary directory and put on
// LD_LIBRARY_PATH of the sub-process that runs the multi-worker harness test, where every worker is a THREAD of one process and
// "device" memory is host memory (fake_cudart.cpp).  Collectives block until every rank of the communicator has arrived.
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

extern "C" {
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;    // ncclSuccess == 0
typedef int ncclDataType_t;  // ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4 ...
typedef int ncclRedOp_t;

struct Group {
    int world = 0, arrived = 0, generation = 0, joined = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<const void*> send;
    std::vector<long long> scratch;
    std::map<std::pair<int, int>, std::deque<std::vector<char>>> mail;  // (source, destination) -> messages in flight
    std::condition_variable mail_cv;
    void barrier(std::unique_lock<std::mutex>& lk) {
        const int gen = generation;
        if (++arrived == world) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    }
};
struct Comm { Group* g; int rank; };
typedef Comm* ncclComm_t;

static std::mutex g_mu;
static std::map<std::string, Group*> g_groups;

static size_t type_size(ncclDataType_t t) { return t == 0 || t == 1 ? 1 : t == 2 || t == 3 || t == 7 ? 4 : t == 6 ? 2 : 8; }

ncclResult_t ncclGetVersion(int* v) { *v = 22809; return 0; }
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
    static std::mt19937_64 rng(std::random_device{}());
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < sizeof id->internal; i += 8) { const unsigned long long r = rng(); memcpy(id->internal + i, &r, 8); }
    return 0;
}
ncclResult_t ncclCommInitRank(ncclComm_t* comm, int world, ncclUniqueId id, int rank) {
    Group* g;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Group*& slot = g_groups[std::string(id.internal, sizeof id.internal)];
        if (!slot) { slot = new Group(); slot->world = world; slot->send.assign((size_t)world, nullptr); }
        g = slot;
    }
    if (g->world != world || rank < 0 || rank >= world) return 4;  // ncclInvalidArgument
    *comm = new Comm{g, rank};
    std::unique_lock<std::mutex> lk(g->mu);
    g->barrier(lk);  // like the real call: returns once every rank has joined
    return 0;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return 0; }
ncclResult_t ncclCommAbort(ncclComm_t c) { delete c; return 0; }
const char* ncclGetErrorString(ncclResult_t r) { return r == 0 ? "no error" : "fake NCCL error (CPU test harness)"; }

ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, ncclDataType_t t, ncclComm_t c, void*) {
    Group* g = c->g;
    const size_t nb = count * type_size(t);
    std::unique_lock<std::mutex> lk(g->mu);
    g->send[(size_t)c->rank] = send;
    g->barrier(lk);
    for (int r = 0; r < g->world; ++r) memmove((char*)recv + (size_t)r * nb, g->send[(size_t)r], nb);
    g->barrier(lk);  // nobody's send buffer is reused before everyone has copied it
    return 0;
}
ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t t, ncclRedOp_t, ncclComm_t c, void*) {
    if (t != 2 && t != 4) return 4;  // int32 / int64 sums are all the library uses
    Group* g = c->g;
    std::unique_lock<std::mutex> lk(g->mu);
    if (g->scratch.size() < count) g->scratch.assign(count, 0);
    g->barrier(lk);
    for (size_t i = 0; i < count; ++i) g->scratch[i] += t == 2 ? (long long)((const int32_t*)send)[i] : ((const long long*)send)[i];
    g->barrier(lk);
    for (size_t i = 0; i < count; ++i) {
        if (t == 2) ((int32_t*)recv)[i] = (int32_t)g->scratch[i];
        else ((long long*)recv)[i] = g->scratch[i];
    }
    g->barrier(lk);
    if (c->rank == 0) std::fill(g->scratch.begin(), g->scratch.end(), 0);
    g->barrier(lk);
    return 0;
}
// point-to-point: a send never blocks (the payload is copied into the (source, destination) mailbox), a receive waits for
// the next message of that pair — the order of a pair's messages is the order of the calls, as NCCL guarantees inside and
// across groups; ncclGroupStart / ncclGroupEnd therefore have nothing to do.
ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void*) {
    Group* g = c->g;
    if (peer < 0 || peer >= g->world) return 4;
    const size_t nb = count * type_size(t);
    std::unique_lock<std::mutex> lk(g->mu);
    g->mail[{c->rank, peer}].emplace_back((const char*)buf, (const char*)buf + nb);
    g->mail_cv.notify_all();
    return 0;
}
ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void*) {
    Group* g = c->g;
    if (peer < 0 || peer >= g->world) return 4;
    const size_t nb = count * type_size(t);
    std::unique_lock<std::mutex> lk(g->mu);
    auto& q = g->mail[{peer, c->rank}];
    if (!g->mail_cv.wait_for(lk, std::chrono::seconds(60), [&] { return !q.empty(); })) return 6;  // ncclRemoteError: the sender never came
    if (q.front().size() != nb) return 4;  // both sides must agree on the message size (they derive it from the same count matrix)
    memcpy(buf, q.front().data(), nb);
    q.pop_front();
    return 0;
}
ncclResult_t ncclGroupStart(void) { return 0; }
ncclResult_t ncclGroupEnd(void) { return 0; }
}
