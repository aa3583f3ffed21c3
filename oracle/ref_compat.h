// oracle/ref_compat.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Force-included (-include) when oracle/Makefile compiles the UNMODIFIED
// reference sources in place from /root/reference.  GCC 13's libstdc++ rejects
// the reference's ArrayAllocator (LightCTR/common/memory_pool.h:104-107: its
// rebind<U>::other is std::allocator<U>, tripping the "rebind_alloc<value_type>
// must be A" static_assert).  We cannot edit the read-only reference tree and we
// do not copy it, so this header pre-defines that file's include guard and
// supplies a standards-conforming allocator with the same names.  No arithmetic
// of the hot path lives in memory_pool.h; only the allocation plumbing changes.
#ifndef LCTR_ORACLE_REF_COMPAT_H
#define LCTR_ORACLE_REF_COMPAT_H
#ifdef __cplusplus
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <list>
#include <memory>
#include <mutex>
#include <new>

#define memory_pool_h  // suppress LightCTR/common/memory_pool.h

class MemoryPool {
public:
    static MemoryPool& Instance() { static MemoryPool pool; return pool; }
    inline void leak_checkpoint() {}
    inline void* allocate(size_t size) { void* p = std::calloc(1, size ? size : 1); assert(p); return p; }
    inline void deallocate(void* ptr) { std::free(ptr); }
};

template <typename T>
class ArrayAllocator {
public:
    typedef T value_type;
    typedef T* pointer;
    typedef const T* const_pointer;
    typedef T& reference;
    typedef const T& const_reference;
    typedef size_t size_type;
    typedef ptrdiff_t difference_type;
    template <typename U> struct rebind { typedef ArrayAllocator<U> other; };
    ArrayAllocator() {}
    template <typename U> ArrayAllocator(const ArrayAllocator<U>&) {}
    pointer allocate(size_type n, const void* = 0) {
        return (T*)MemoryPool::Instance().allocate(n * sizeof(T));
    }
    void deallocate(pointer p, size_type) { MemoryPool::Instance().deallocate(p); }
    size_type max_size() const { return size_type(UINTMAX_MAX / sizeof(T)); }
};
template <typename A, typename B>
inline bool operator==(const ArrayAllocator<A>&, const ArrayAllocator<B>&) { return true; }
template <typename A, typename B>
inline bool operator!=(const ArrayAllocator<A>&, const ArrayAllocator<B>&) { return false; }
#endif  // __cplusplus
#endif
