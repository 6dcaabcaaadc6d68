/*
 * integration/b200_context.h -- one GPU context per host thread, shared by the
 * re-defined members: the reference runs one DepthOptimizer (and its
 * StereoViews) per pool thread (app/smvsrecon.cc:658-733).
 */
#ifndef SMVS_B200_CONTEXT_H
#define SMVS_B200_CONTEXT_H

#include <memory>

#include "smvs_b200.hpp"

namespace smvs_b200_integration {

inline smvsb::Context&
thread_context (void)
{
    static thread_local std::unique_ptr<smvsb::Context> ctx;
    if (!ctx)
        ctx.reset(new smvsb::Context(0));
    return *ctx;
}

}

#endif
