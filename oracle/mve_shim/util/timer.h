/* Shim of MVE util/timer.h. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_UTIL_TIMER_HEADER
#define SHIM_UTIL_TIMER_HEADER

#include <chrono>
#include <cstddef>

namespace util {

class WallTimer
{
public:
    WallTimer (void) { this->reset(); }
    void reset (void) { this->start = std::chrono::high_resolution_clock::now(); }
    std::size_t get_elapsed (void) const
    {
        return std::chrono::duration_cast<std::chrono::milliseconds>(
            std::chrono::high_resolution_clock::now() - start).count();
    }
    float get_elapsed_sec (void) const
    { return (1.0f / 1000.0f) * static_cast<float>(this->get_elapsed()); }
private:
    std::chrono::high_resolution_clock::time_point start;
};

}

#endif
