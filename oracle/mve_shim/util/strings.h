/* Shim of MVE util/strings.h (debug names only). TEST INFRA ONLY. */
#ifndef SHIM_UTIL_STRINGS_HEADER
#define SHIM_UTIL_STRINGS_HEADER

#include <iomanip>
#include <sstream>
#include <string>

namespace util {
namespace string {

template <typename T>
inline std::string
get (T const& value)
{
    std::stringstream ss;
    ss << value;
    return ss.str();
}

template <typename T>
inline std::string
get_filled (T const& value, int width, char fill = '0')
{
    std::stringstream ss;
    ss << std::setw(width) << std::setfill(fill) << value;
    return ss.str();
}

}
}

#endif
