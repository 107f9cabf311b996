// TEST HARNESS (not product): zerokit_b200/csrc/coalesce.hpp under ThreadSanitizer — many threads submit single requests, a fake
// batch function squares numbers (and sleeps like a GPU step would).  Checks: every request answered exactly once with its own
// result, no batch above the limit, concurrent callers really share batches, a lone caller gets a batch of one.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "coalesce.hpp"

struct Req {
    long in = 0, out = -1;
    int served = 0;
    bool done = false;
};

int main(int argc, char** argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 16, per_thread = argc > 2 ? atoi(argv[2]) : 40;
    const size_t max_batch = argc > 3 ? (size_t)atoi(argv[3]) : 8;
    zk::Coalescer<Req> co;
    std::atomic<long> batches{0}, items{0}, biggest{0};
    std::atomic<int> in_flight{0}, overlap{0};
    auto run = [&](std::vector<Req*>& b) {
        if (in_flight.fetch_add(1) != 0) overlap++;     // run() must never overlap with itself
        if (b.size() > max_batch || b.empty()) { fprintf(stderr, "bad batch size %zu\n", b.size()); abort(); }
        std::this_thread::sleep_for(std::chrono::microseconds(300));
        for (Req* r : b) { r->out = r->in * r->in; r->served++; }
        batches++;
        items += (long)b.size();
        long s = (long)b.size(), cur = biggest.load();
        while (s > cur && !biggest.compare_exchange_weak(cur, s)) {}
        in_flight.fetch_sub(1);
    };
    {   // a lone caller: one batch of one
        Req r; r.in = 7;
        co.submit(r, max_batch, run);
        if (!r.done || r.out != 49 || r.served != 1 || batches != 1) { fprintf(stderr, "lone caller failed\n"); return 1; }
    }
    std::atomic<int> bad{0};
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; t++)
        ts.emplace_back([&, t] {
            for (int i = 0; i < per_thread; i++) {
                Req r; r.in = 1000L * t + i;
                co.submit(r, max_batch, run);
                if (!r.done || r.out != r.in * r.in || r.served != 1) bad++;
            }
        });
    for (auto& t : ts) t.join();
    const long total = 1 + (long)threads * per_thread;
    printf("requests %ld, batches %ld, largest batch %ld, overlap %d, bad %d\n", total, batches.load(), biggest.load(), overlap.load(), bad.load());
    if (bad || overlap || items != total) return 1;
    if (threads > 1 && batches.load() >= total) { fprintf(stderr, "no coalescing happened\n"); return 1; }
    printf("coalesce ok\n");
    return 0;
}
