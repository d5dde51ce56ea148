// base.h -- shared infrastructure of libamgx_b200: error type, CUDA checks, device buffers,
// precision dispatch.  Host C++17 + CUDA runtime only.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <memory>
#include <stdexcept>
#include "../../include/amgx_b200.h"

namespace amgxb {

// Exception carried up to the C-ABI boundary, where it becomes an AMGX_RC
// (the reference does the same with amgx_exception -> getCAPIerror_x, include/amgx_c_common.h:26-47).
struct Error : public std::exception {
    AMGX_RC rc;
    std::string msg;
    Error(AMGX_RC r, std::string m) : rc(r), msg(std::move(m)) {}
    const char *what() const noexcept override { return msg.c_str(); }
};

[[noreturn]] inline void fatal(AMGX_RC rc, const std::string &m) { throw Error(rc, m); }

#define AMGXB_CUDA_CHECK(expr)                                                                   \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _b[512];                                                                        \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),         \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                                \
            throw ::amgxb::Error(_e == cudaErrorMemoryAllocation ? AMGX_RC_NO_MEMORY             \
                                                                 : AMGX_RC_CUDA_FAILURE, _b);    \
        }                                                                                        \
    } while (0)

#define AMGXB_LAUNCH_CHECK() AMGXB_CUDA_CHECK(cudaGetLastError())

// Output routed through the user print callback (AMGX_register_print_callback).
void amgx_output(const char *msg, int len);
void amgx_printf(const char *fmt, ...);

// Opt a kernel in to `smem` bytes of dynamic shared memory.  The attribute belongs to the (function, device) pair, so what has been
// granted is remembered per device: a process that drives several GPUs opts in on each of them (misc.cu).
void smem_opt_in(const void *kernel, size_t smem);

// Global launch counter (our kernels only): incremented by every launcher in k_*.cu.
extern long long g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches += n; }

// ---------------------------------------------------------------------------------------------
// Scalar precisions.  dDDI = (mat f64, vec f64); dDFI = (mat f32, vec f64); dFFI = (f32, f32)
// (reference mode arithmetic: include/amgx_config.h:81-124).
// ---------------------------------------------------------------------------------------------
enum class Prec : int { F64 = 0, F32 = 1 };
inline size_t prec_size(Prec p) { return p == Prec::F64 ? 8 : 4; }

struct ModeInfo {
    bool host;      // h* mode requested (computation still runs on the GPU)
    Prec vec, mat;
};
inline ModeInfo decode_mode(int mode)
{
    int mem = mode % 16, vec = (mode / 16) % 16, mat = (mode / 256) % 16, ind = (mode / 4096) % 16;
    if ((mem != 0 && mem != 1) || (vec != 0 && vec != 1) || (mat != 0 && mat != 1) || ind != 2 ||
        (vec == 1 && mat == 0))
        fatal(AMGX_RC_BAD_MODE, "unsupported AMGX_Mode");
    return ModeInfo{mem == 0, vec == 0 ? Prec::F64 : Prec::F32, mat == 0 ? Prec::F64 : Prec::F32};
}

// Dispatch a generic lambda on (MatT, VecT).
#define AMGXB_DISPATCH(matp, vecp, ...)                                                          \
    do {                                                                                         \
        if ((matp) == ::amgxb::Prec::F64 && (vecp) == ::amgxb::Prec::F64) {                      \
            using MatT = double; using VecT = double; __VA_ARGS__                                \
        } else if ((matp) == ::amgxb::Prec::F32 && (vecp) == ::amgxb::Prec::F64) {               \
            using MatT = float; using VecT = double; __VA_ARGS__                                 \
        } else if ((matp) == ::amgxb::Prec::F32 && (vecp) == ::amgxb::Prec::F32) {               \
            using MatT = float; using VecT = float; __VA_ARGS__                                  \
        } else ::amgxb::fatal(AMGX_RC_BAD_MODE, "unsupported precision combination");           \
    } while (0)

#define AMGXB_DISPATCH_VEC(vecp, ...)                                                            \
    do {                                                                                         \
        if ((vecp) == ::amgxb::Prec::F64) { using VecT = double; __VA_ARGS__ }                   \
        else { using VecT = float; __VA_ARGS__ }                                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Device buffer: raw bytes, 256-byte aligned by cudaMalloc, over-allocated by 64 bytes so that
// 16-byte-granular bulk copies (TMA) may over-read the tail of an array safely.
// ---------------------------------------------------------------------------------------------
struct DevBytes {
    void *p = nullptr;
    size_t bytes = 0;      // logical size
    size_t cap = 0;        // allocated size
    DevBytes() = default;
    DevBytes(const DevBytes &) = delete;
    DevBytes &operator=(const DevBytes &) = delete;
    DevBytes(DevBytes &&o) noexcept { *this = std::move(o); }
    DevBytes &operator=(DevBytes &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; cap = o.cap; o.p = nullptr; o.bytes = o.cap = 0; }
        return *this;
    }
    ~DevBytes() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; bytes = cap = 0;
    }
    void resize(size_t nbytes)   // contents undefined after growth
    {
        if (nbytes + 64 > cap) {
            release();
            size_t want = nbytes + 64;
            AMGXB_CUDA_CHECK(cudaMalloc(&p, want));
            cap = want;
        }
        bytes = nbytes;
    }
    void swap(DevBytes &o) { std::swap(p, o.p); std::swap(bytes, o.bytes); std::swap(cap, o.cap); }
};

template <class T> struct DevBuf {
    DevBytes b;
    size_t n = 0;
    T *ptr() { return (T *)b.p; }
    const T *ptr() const { return (const T *)b.p; }
    size_t size() const { return n; }
    void resize(size_t count) { b.resize(count * sizeof(T)); n = count; }
    void zero(cudaStream_t s = 0) { if (n) AMGXB_CUDA_CHECK(cudaMemsetAsync(b.p, 0, n * sizeof(T), s)); }
    void release() { b.release(); n = 0; }
    void swap(DevBuf &o) { b.swap(o.b); std::swap(n, o.n); }
    void from_any(const T *src, size_t count, cudaStream_t s = 0)   // host or device pointer
    {
        resize(count);
        if (count) AMGXB_CUDA_CHECK(cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyDefault, s));
    }
    std::vector<T> to_host(cudaStream_t s = 0) const
    {
        std::vector<T> h(n);
        if (n) {
            AMGXB_CUDA_CHECK(cudaMemcpyAsync(h.data(), b.p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
            AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        }
        return h;
    }
};

// A typed-at-runtime vector of scalars (fp32 or fp64) on the device.
struct DevVec {
    DevBytes b;
    Prec prec = Prec::F64;
    size_t n = 0;
    void *ptr() { return b.p; }
    const void *ptr() const { return b.p; }
    template <class T> T *as() { return (T *)b.p; }
    template <class T> const T *as() const { return (const T *)b.p; }
    void resize(size_t count, Prec p) { prec = p; b.resize(count * prec_size(p)); n = count; }
    void zero(cudaStream_t s = 0) { if (n) AMGXB_CUDA_CHECK(cudaMemsetAsync(b.p, 0, n * prec_size(prec), s)); }
    void swap(DevVec &o) { b.swap(o.b); std::swap(prec, o.prec); std::swap(n, o.n); }
    size_t nbytes() const { return n * prec_size(prec); }
};

// Grid-stride kernels cap their grids at a small multiple of the SM count of the ONE target of this library (sm_100a, B200: 148 SMs on two
// dies); the persistent tile kernels size their grids from Resources::num_sms (the device's own count) and the occupancy of the instantiation.
constexpr int B200_SMS = 148;

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace amgxb
