// tests/cpp/test_host_mirror.cpp — exercises the C++ host mirror (include/dfd_b200.hpp) through the C ABI and
// checks it against the C oracle (test infrastructure).  Built and run by tests/test_cpp_host.py:
//   g++ -std=c++17 -I include -I oracle tests/cpp/test_host_mirror.cpp -o ... -L<lib dirs> -ldfd_b200 -ldf_oracle
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "df_oracle.h"
#include "dfd_b200.hpp"

#define REQUIRE(c)                                                         \
    do {                                                                   \
        if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const bool compile_only = argc > 1 && strcmp(argv[1], "--no-gpu") == 0;
    using namespace dfd;
    // surface checks that need no GPU (the shapes the reference's plan tests assert)
    Partitioning hash = Partitioning::Hash({0}, 4);
    NetworkShuffleExec node = NetworkShuffleExec::try_new(hash, {}, 1, /*task_count=*/1, /*input_task_count=*/1);
    REQUIRE(std::string(node.name()) == "NetworkShuffleExec");
    REQUIRE(node.output_partitioning().partition_count == 4);
    REQUIRE(node.input_stage().plan.partition_count == 4);
    REQUIRE(NetworkShuffleExec::try_new(hash, {}, 1, 2, 2).input_stage().plan.partition_count == 8);
    bool threw = false;
    try { NetworkShuffleExec::try_new(Partitioning::Hash({}, 4), {}, 1, 1, 1); } catch (const Error& e) { threw = e.status() == DFD_ERR_INVALID_ARGUMENT; }
    REQUIRE(threw);
    if (compile_only) { printf("CPP_HOST_MIRROR_SURFACE_OK\n"); return 0; }

    // data: cfg-1 shape, (k: Int64, v: Int64)
    const int64_t n = 300007;
    const uint32_t N = 4;
    std::vector<int64_t> k(n), v(n);
    uint64_t s = 88172645463325252ULL;
    for (int64_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; k[i] = (int64_t)s; v[i] = i; }
    WorkerContext ctx(0);
    const size_t bytes = (size_t)n * 8;
    void* d_k = ctx.device_alloc(bytes); void* d_v = ctx.device_alloc(bytes);
    void* o_k = ctx.device_alloc(bytes); void* o_v = ctx.device_alloc(bytes);
    ctx.h2d(d_k, k.data(), bytes); ctx.h2d(d_v, v.data(), bytes);

    // oracle
    const void* cols[2] = {k.data(), v.data()};
    int32_t widths[2] = {8, 8}, keys[1] = {0};
    std::vector<int64_t> ref_k(n), ref_v(n), counts(N), starts(N + 1);
    void* outs[2] = {ref_k.data(), ref_v.data()};
    REQUIRE(orc_repartition_table(cols, widths, 2, n, keys, 1, N, 8192, 1, outs, counts.data(), starts.data()) == 0);

    // 1. BatchPartitioner mirror
    HashPartitioner part(ctx, hash);
    std::vector<dfd_column> in = {fixed_column(d_k, 8), fixed_column(d_v, 8)}, out = {fixed_column(o_k, 8), fixed_column(o_v, 8)};
    std::vector<int64_t> ps = part.partition(in, n, out);
    REQUIRE(ps == starts);
    std::vector<int64_t> got(n);
    ctx.d2h(got.data(), o_v, bytes);
    REQUIRE(got == ref_v);
    ctx.d2h(got.data(), o_k, bytes);
    REQUIRE(got == ref_k);

    // 2. NetworkShuffleExec mirror, one worker, fused exchange: execute(p) == the oracle's partition p
    ShuffleExchange x(ctx, 0, 1, nullptr);
    x.setup_window(2 * bytes + (1 << 20));
    std::vector<dfd_column> recv;
    node.shuffle(ctx, x, in, n, recv);
    for (uint32_t p = 0; p < N; ++p) {
        auto range = node.execute(p, DistributedTaskContext{0, 1});
        REQUIRE(range.first == starts[p] && range.second == starts[p + 1]);
    }
    ctx.d2h(got.data(), recv[1].values, bytes);
    REQUIRE(got == ref_v);

    // 3. error behaviour: C status codes surface as dfd::Error
    threw = false;
    try { HashPartitioner bad(ctx, Partitioning::Hash({0}, 0)); } catch (const Error& e) { threw = e.status() == DFD_ERR_INVALID_ARGUMENT; }
    REQUIRE(threw);
    REQUIRE(ctx.metrics().kernel_launches >= 6);
    ctx.device_free(d_k); ctx.device_free(d_v); ctx.device_free(o_k); ctx.device_free(o_v);
    printf("CPP_HOST_MIRROR_OK\n");
    return 0;
}
