// Neon.h — core vocabulary of the C++ host veneer over libneon_lbm.so.
//
// The reference is C++17/CUDA and that toolchain exists in this image, so the host side of the drop-in is C++ too:
// a header-only layer that keeps the names the lid-driven-cavity benchmark uses (Neon::init, Neon::index_3d,
// Neon::DataView, Neon::Runtime, NeonException/NEON_THROW, Neon::Pattern a.k.a. Neon::Compute) and routes every
// device operation of the LBM path through the C ABI of include/neon_lbm.h.  Nothing here computes.
//
// Mirrors (names and meaning, not code):
//   libNeonCore/include/Neon/core/types/{vec/vec3d_generic.h, DataView.h:7-12, Execution.h, Exceptions.h:19-24}
//   libNeonSys/src/sys/Neon.cpp:4 (Neon::init)
#pragma once

#include <cstdint>
#include <exception>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "neon_lbm.h"

/* libNeonCore/include/Neon/core/types/Macros.h:80-106 — host code compiled by g++ sees empty qualifiers */
#ifdef __CUDACC__
#define NEON_CUDA_HOST_DEVICE __host__ __device__
#define NEON_CUDA_DEVICE_ONLY __device__
#define NEON_CUDA_HOST_ONLY __host__
#else
#define NEON_CUDA_HOST_DEVICE
#define NEON_CUDA_DEVICE_ONLY
#define NEON_CUDA_HOST_ONLY
#endif

/* NVTX ranges as in the reference (compile-time NEON_USE_NVTX, cmake/Nvtx.cmake:3): "Skeleton" around Skeleton::run
 * (Skeleton.h:59) and one range per container (Graph.cpp:1002,1017).  nvtx3 is header-only. */
#ifdef NEON_USE_NVTX
#include <nvtx3/nvToolsExt.h>
#define NEON_NVTX_PUSH(name) nvtxRangePushA(name)
#define NEON_NVTX_POP() nvtxRangePop()
#else
#define NEON_NVTX_PUSH(name) ((void)0)
#define NEON_NVTX_POP() ((void)0)
#endif

namespace Neon {

// ---------------------------------------------------------------------------------------------------- small vectors
template <typename T>
struct Vec_3d
{
    union
    {
        struct
        {
            T x, y, z;
        };
        T v[3];
    };
    NEON_CUDA_HOST_DEVICE constexpr Vec_3d() : x(0), y(0), z(0) {}
    NEON_CUDA_HOST_DEVICE constexpr Vec_3d(T a) : x(a), y(a), z(a) {}
    NEON_CUDA_HOST_DEVICE constexpr Vec_3d(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename U>
    NEON_CUDA_HOST_DEVICE constexpr explicit Vec_3d(const Vec_3d<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z))
    {
    }
    NEON_CUDA_HOST_DEVICE constexpr bool operator==(const Vec_3d& o) const { return x == o.x && y == o.y && z == o.z; }
    NEON_CUDA_HOST_DEVICE constexpr bool operator!=(const Vec_3d& o) const { return !(*this == o); }
    NEON_CUDA_HOST_DEVICE constexpr Vec_3d operator+(const Vec_3d& o) const { return {T(x + o.x), T(y + o.y), T(z + o.z)}; }
    NEON_CUDA_HOST_DEVICE constexpr Vec_3d operator-(const Vec_3d& o) const { return {T(x - o.x), T(y - o.y), T(z - o.z)}; }
    NEON_CUDA_HOST_DEVICE constexpr Vec_3d operator-() const { return {T(-x), T(-y), T(-z)}; }
    template <typename K = size_t>
    NEON_CUDA_HOST_DEVICE constexpr K rMul() const
    {
        return K(x) * K(y) * K(z);
    }
    std::string to_string() const
    {
        std::ostringstream s;
        s << "(" << x << ", " << y << ", " << z << ")";
        return s.str();
    }
};
using index_3d = Vec_3d<int32_t>;
using int32_3d = Vec_3d<int32_t>;
using int8_3d = Vec_3d<int8_t>;
using double_3d = Vec_3d<double>;
using float_3d = Vec_3d<float>;

// ---------------------------------------------------------------------------------------------------- enums
enum class DataView
{
    STANDARD = NLBM_VIEW_STANDARD,
    INTERNAL = NLBM_VIEW_INTERNAL,
    BOUNDARY = NLBM_VIEW_BOUNDARY
};
struct DataViewUtil
{
    static const char* toString(DataView v)
    {
        return v == DataView::STANDARD ? "STANDARD" : v == DataView::INTERNAL ? "INTERNAL" : "BOUNDARY";
    }
};
enum class Runtime
{
    none,
    system,
    stream, /* CUDA streams: the only runtime that computes here */
    openmp  /* host logic only: there is no CPU compute path behind this veneer */
};
enum class Execution
{
    device,
    host
};
enum class Pattern
{
    MAP,
    STENCIL,
    REDUCE
};
using Compute = Pattern; /* older spelling, benchmarks/lbm-flow-over-sphere/src/LbmContainers.h:455-459 */
enum class computeMode_t
{
    par,
    seq
};
enum class MemoryLayout
{
    structOfArrays,
    arrayOfStructs
};
enum class IoFileType
{
    ASCII,
    BINARY
};

// ---------------------------------------------------------------------------------------------------- errors
class NeonException : public std::exception
{
   public:
    NeonException() = default;
    explicit NeonException(const std::string& where) : mWhere(where) {}
    NeonException(const NeonException& o) : mWhere(o.mWhere), mWhat(o.mWhat) { mMsg << o.mMsg.str(); }
    template <typename T>
    NeonException& operator<<(const T& v)
    {
        mMsg << v;
        return *this;
    }
    const char* what() const noexcept override
    {
        mWhat = "[NeonException] " + mWhere + ": " + mMsg.str();
        return mWhat.c_str();
    }

   private:
    std::string         mWhere;
    std::ostringstream  mMsg;
    mutable std::string mWhat;
};
#define NEON_THROW(exc) throw(exc)
#define NEON_THROW_UNSUPPORTED_OPERATION(msg)                       \
    {                                                               \
        Neon::NeonException neonExc_(__func__);                     \
        neonExc_ << "unsupported operation: " << std::string(msg); \
        throw neonExc_;                                             \
    }
#define NEON_DEV_UNDER_CONSTRUCTION(msg) NEON_THROW_UNSUPPORTED_OPERATION(std::string("under construction ") + std::string(msg))

namespace detail {
/* status of a C-ABI call -> NeonException (the reference checks every CUDA return, GpuDevice.h:165-188) */
inline void check(int rc, const char* what)
{
    if (rc != NLBM_OK) {
        NeonException e(what);
        e << "nlbm status " << rc << ": " << nlbm_last_error();
        NEON_THROW(e);
    }
}
}  // namespace detail

/* Neon::init (libNeonSys/src/sys/Neon.cpp:4): verifies that the kernel library this veneer was compiled against is
 * the one loaded.  There is no CPU fallback to select. */
inline void init()
{
    if (nlbm_abi_version() != NLBM_ABI_VERSION) {
        NeonException e("Neon::init");
        e << "libneon_lbm.so has ABI " << nlbm_abi_version() << ", headers expect " << NLBM_ABI_VERSION;
        NEON_THROW(e);
    }
}

}  // namespace Neon
