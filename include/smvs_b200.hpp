/*
 * smvs_b200.hpp -- thin C++11 RAII layer over the C ABI (smvs_b200.h) for the
 * reference's C++ host: errors become exceptions the way the reference
 * reports them (std::invalid_argument / std::runtime_error), one Context per
 * DepthOptimizer (per pool thread).
 */
#ifndef SMVS_B200_HPP
#define SMVS_B200_HPP

#include <stdexcept>
#include <string>

#include "smvs_b200.h"

namespace smvsb {

class Context
{
public:
    explicit Context (int device = 0) : ctx(nullptr)
    {
        if (smvsb_create(device, &ctx) != SMVSB_OK)
            throw std::runtime_error(std::string("smvs_b200: ")
                + smvsb_last_error(nullptr));
    }
    ~Context (void) { smvsb_destroy(ctx); }
    Context (Context const&) = delete;
    Context& operator= (Context const&) = delete;

    smvsb_ctx* get (void) const { return ctx; }

    /* Maps a status code to the exception type the reference would throw. */
    void check (int rc) const
    {
        if (rc == SMVSB_OK)
            return;
        std::string const msg = std::string("smvs_b200: ")
            + smvsb_last_error(ctx);
        if (rc == SMVSB_ERR_INVALID)
            throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }

private:
    smvsb_ctx* ctx;
};

} /* namespace smvsb */

#endif /* SMVS_B200_HPP */
