// lesv_gpus.hpp -- one feeder thread per GPU for the drop-in tools.
//
// The reference runs T pthreads over a shared work list and serialises output with a mutex
// (app/map/chain_and_extend_kmer_matches.c:381-388, app/necat2sv/find_sv_signature.c:383-411).  Here one
// host process drives every visible GPU: each GPU has its own lesvgpu_ctx and its own feeder thread, the feeders
// pull sub-batches from one shared queue (an atomic counter over a list fixed BEFORE the work starts) and park
// their results in the slot of the sub-batch, so what is written afterwards does not depend on how many GPUs
// took part or on which of them was faster.  Nothing is exchanged between GPUs: the subject .pac and the index
// are replicated per context.
//
// Device selection: LESV_GPUS="0,2,3" (list), else LESV_GPU=<n> (one device: one process per GPU, as `-grid`
// and torchrun-style launches use it), else every device the CUDA runtime shows.
#pragma once
#include "../../include/lesvgpu.h"
#include "lesv_db.hpp"

#include <atomic>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace lesvgpus {

inline std::vector<int> devices()
{
    std::vector<int> d;
    if (const char* l = getenv("LESV_GPUS")) {
        for (const char* p = l; *p;) {
            char* e = nullptr;
            const long v = strtol(p, &e, 10);
            if (e == p) break;
            d.push_back((int)v);
            p = *e == ',' ? e + 1 : e;
        }
    } else if (const char* g = getenv("LESV_GPU")) {
        d.push_back(atoi(g));
    } else {
        const int n = lesvgpu_device_count();
        for (int i = 0; i < n; ++i) d.push_back(i);
    }
    if (d.empty()) lesvdb::die("no usable B200 (no CUDA device visible); there is no CPU path");
    return d;
}

struct Pool {
    std::vector<int> dev;
    std::vector<lesvgpu_ctx*> ctx;

    void open()
    {
        dev = devices();
        ctx.assign(dev.size(), nullptr);
        each([&](int g) {
            if (lesvgpu_ctx_create(dev[g], 0, &ctx[g]) != LESVGPU_OK)
                lesvdb::die("no usable B200 as device " + std::to_string(dev[g]) + " (lesvgpu_ctx_create failed); there is no CPU path");
        });
    }
    void close()
    {
        for (auto* c : ctx) lesvgpu_ctx_destroy(c);
        ctx.clear();
    }
    int size() const { return (int)ctx.size(); }

    // fn(g) once per GPU, each on its own thread
    void each(const std::function<void(int)>& fn) const
    {
        if (dev.size() == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int g = 0; g < (int)dev.size(); ++g) th.emplace_back(fn, g);
        for (auto& t : th) t.join();
    }
    // fn(g, item) for item = 0 .. n_items-1, every item exactly once, handed out in order to whichever GPU is free
    void queue(int n_items, const std::function<void(int, int)>& fn) const
    {
        std::atomic<int> next{0};
        each([&](int g) {
            for (;;) {
                const int item = next.fetch_add(1);
                if (item >= n_items) break;
                fn(g, item);
            }
        });
    }
};

}  // namespace lesvgpus
