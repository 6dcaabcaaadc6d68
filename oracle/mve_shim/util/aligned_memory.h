/* Shim of MVE util/aligned_memory.h: std::vector with aligned storage. */
#ifndef SHIM_UTIL_ALIGNED_MEMORY_HEADER
#define SHIM_UTIL_ALIGNED_MEMORY_HEADER

#include <cstddef>
#include <cstdlib>
#include <new>
#include <vector>

namespace util {

template <typename T, std::size_t MODULO = 16>
struct AlignedAllocator
{
    typedef T value_type;
    typedef T* pointer;
    typedef T const* const_pointer;
    typedef T& reference;
    typedef T const& const_reference;
    typedef std::size_t size_type;
    typedef std::ptrdiff_t difference_type;
    template <class U> struct rebind { typedef AlignedAllocator<U, MODULO> other; };

    AlignedAllocator (void) {}
    template <class U>
    AlignedAllocator (AlignedAllocator<U, MODULO> const&) {}

    pointer allocate (size_type n)
    {
        if (n == 0)
            return nullptr;
        void* p = nullptr;
        std::size_t const align = MODULO < sizeof(void*) ? sizeof(void*) : MODULO;
        if (posix_memalign(&p, align, n * sizeof(T)) != 0)
            throw std::bad_alloc();
        return static_cast<pointer>(p);
    }
    void deallocate (pointer p, size_type) { std::free(p); }
    bool operator== (AlignedAllocator const&) const { return true; }
    bool operator!= (AlignedAllocator const&) const { return false; }
};

template <typename T, std::size_t MODULO = 16>
using AlignedMemory = std::vector<T, AlignedAllocator<T, MODULO> >;

}

#endif
