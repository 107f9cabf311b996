// Request coalescing for the single-item entry points of the C ABI.
//
// The reference's proving / verifying calls take &self and are re-entrant: a server proves or verifies from many
// threads on one handle, one item per call.  On the GPU one item costs the latency of a whole batch step (16.7 ms per proof,
// 44 ms per verification) while 4 096 items cost barely more, so serialising such callers behind the handle's mutex throws the
// device away.  Coalescer turns concurrent single calls into batches without a background thread: every caller queues its
// request; whoever finds no batch in flight becomes the leader, takes what is queued (its own request and everything that
// arrived while the previous batch ran), runs it as ONE batch and wakes the owners.  A lone caller pays one mutex round trip.
//
// Host logic only; exercised without a GPU by tests/host_fuzz/coalesce_test.cpp (ThreadSanitizer).
#pragma once
#include <condition_variable>
#include <cstddef>
#include <mutex>
#include <vector>

namespace zk {

// Req needs a public `bool done` (false on entry).  `run(std::vector<Req*>&)` processes every request of the batch and must not
// throw (record failures inside the request); it is never called concurrently with itself for one Coalescer.
template <class Req>
class Coalescer {
public:
    template <class Run>
    void submit(Req& r, size_t max_batch, Run&& run) {
        std::unique_lock<std::mutex> lk(mu_);
        queue_.push_back(&r);
        for (;;) {
            if (r.done) return;
            if (!busy_) {   // lead: take the oldest requests (not necessarily including our own if the queue is longer than a batch)
                busy_ = true;
                const size_t take = queue_.size() < max_batch ? queue_.size() : max_batch;
                std::vector<Req*> batch(queue_.begin(), queue_.begin() + take);
                queue_.erase(queue_.begin(), queue_.begin() + take);
                lk.unlock();
                run(batch);
                lk.lock();
                for (Req* b : batch) b->done = true;
                busy_ = false;
                cv_.notify_all();
                continue;
            }
            cv_.wait(lk);
        }
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Req*> queue_;
    bool busy_ = false;
};

}  // namespace zk
