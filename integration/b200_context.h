/*
 * integration/b200_context.h -- one GPU context per host thread, shared by the
 * re-defined members: the reference runs one DepthOptimizer (and its
 * StereoViews) per pool thread (app/smvsrecon.cc:558,658-733), so pool thread
 * k works on device k mod (number of devices): the views of a scene shard over
 * all GPUs of the box without any change to smvsrecon.
 *
 *   SMVSB_DEVICES=0,2,3   restricts / orders the devices used (default: all)
 */
#ifndef SMVS_B200_CONTEXT_H
#define SMVS_B200_CONTEXT_H

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "smvs_b200.hpp"

namespace smvs_b200_integration {

/* Devices the pool threads are spread over. */
inline std::vector<int> const&
device_list (void)
{
    static std::vector<int> const list = [] {
        std::vector<int> out;
        int const count = smvsb_device_count();
        char const* env = std::getenv("SMVSB_DEVICES");
        if (env != nullptr)
        {
            std::string s(env);
            std::size_t pos = 0;
            while (pos < s.size())
            {
                std::size_t const end = s.find(',', pos);
                std::string const tok = s.substr(pos, end == std::string::npos
                    ? std::string::npos : end - pos);
                if (!tok.empty())
                {
                    int const d = std::atoi(tok.c_str());
                    if (d >= 0 && d < count)
                        out.push_back(d);
                }
                if (end == std::string::npos)
                    break;
                pos = end + 1;
            }
        }
        if (out.empty())
            for (int d = 0; d < count; ++d)
                out.push_back(d);
        if (out.empty())
            out.push_back(0);     /* smvsb_create reports the missing GPU */
        return out;
    }();
    return list;
}

/* Ordinal of the calling thread among the threads that asked so far (the
 * pool's workers, in the order they first reach the GPU path). */
inline int
thread_ordinal (void)
{
    static std::atomic<int> next(0);
    static thread_local int const mine = next.fetch_add(1);
    return mine;
}

inline int
thread_device (void)
{
    std::vector<int> const& list = device_list();
    return list[static_cast<std::size_t>(thread_ordinal()) % list.size()];
}

inline smvsb::Context&
thread_context (void)
{
    static thread_local std::unique_ptr<smvsb::Context> ctx;
    if (!ctx)
        ctx.reset(new smvsb::Context(thread_device()));
    return *ctx;
}

/* Bumped by every StereoView::set_scale on this thread: image data a context
 * holds from before is then stale, whatever addresses the new images have. */
inline std::uint64_t&
views_generation (void)
{
    static thread_local std::uint64_t gen = 1;
    return gen;
}

}

#endif
