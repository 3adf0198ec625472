// grb_ops.cuh -- device-side operator algebra: typed binary ops, monoid identities, semiring
// functors (compile-time specialised for the hot set, run-time op codes for the rest) and the
// atomic "combine" used by the push (scatter) SpMSpV kernel and the SpGEMM hash accumulators.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <type_traits>

#include "grb_internal.h"

// BOOL is stored as one byte holding 0/1; kernels see it as this wrapper so that the
// arithmetic ops get their logical meaning (plus=lor, times=land, min=land, max=lor, minus=lxor)
struct gbool {
    uint8_t v;
    gbool() = default;
    __host__ __device__ explicit gbool(bool b) : v(b ? 1 : 0) {}
};

template <typename T> struct is_gbool : std::false_type {};
template <> struct is_gbool<gbool> : std::true_type {};

template <typename T> __host__ __device__ __forceinline__ bool truthy(T x) { return x != (T)0; }
template <> __host__ __device__ __forceinline__ bool truthy<gbool>(gbool x) { return x.v != 0; }

template <typename T> __host__ __device__ __forceinline__ T from_bool(bool b) { return (T)(b ? 1 : 0); }
template <> __host__ __device__ __forceinline__ gbool from_bool<gbool>(bool b) { return gbool(b); }

template <typename T> __host__ __device__ __forceinline__ T one_of() { return (T)1; }
template <> __host__ __device__ __forceinline__ gbool one_of<gbool>() { return gbool(true); }

// ---- generic typed binary op with a (possibly compile-time constant) op code ----
// pow is a long routine (through double for the integer types): kept OUT of line on the device, because binop() is inlined
// with a run-time opcode into every kernel that applies an accumulator or a dynamic semiring, and none of the hot paths uses it
template <typename T> __host__ __device__ inline T pow_value(T x, T y) {
    if constexpr (std::is_floating_point<T>::value) {
        return (T)pow(x, y);
    } else {
        typedef typename std::make_unsigned<T>::type U;   // through double, saturating, NaN -> 0 (how SuiteSparse defines integer pow)
        const double r = pow((double)x, (double)y);
        if (r != r) return (T)0;
        const T tmin = std::is_signed<T>::value ? (T)((U)1 << (sizeof(T) * 8 - 1)) : (T)0;
        const T tmax = std::is_signed<T>::value ? (T)(~((U)1 << (sizeof(T) * 8 - 1))) : (T)~(U)0;
        if (r <= (double)tmin) return tmin;
        if (r >= (double)tmax) return tmax;
        return (T)r;
    }
}
#ifdef __CUDACC__
template <typename T> __device__ __noinline__ T pow_out_of_line(T x, T y) { return pow_value<T>(x, y); }
#endif
template <typename T> __host__ __device__ __forceinline__ T pow_op(T x, T y) {
#ifdef __CUDA_ARCH__
    return pow_out_of_line<T>(x, y);
#else
    return pow_value<T>(x, y);
#endif
}

// WITH_POW = false: the same switch without pow -- for the accumulator fused into the SpMV kernels (spmv_common.cuh epi_write; the
// host never fuses a pow accumulator): with pow in it the int64 segmented kernel spilled 56 bytes and the masked pull went from 32 to 48
// registers (cuobjdump --dump-resource-usage), an SSSP sweep from 0.745 to 0.84 ms
template <typename T, bool WITH_POW = true> __host__ __device__ __forceinline__ T binop(int op, T x, T y) {
    if constexpr (is_gbool<T>::value) {
        switch (op) {
            case OP_FIRST: case OP_DIV: return x;
            case OP_SECOND: case OP_RDIV: case OP_ANY: return y;   // any(x, y) may return either; y makes any(identity, p) = p
            case OP_PAIR: return gbool(true);
            case OP_PLUS: case OP_LOR: case OP_MAX: return gbool(x.v | y.v);
            case OP_TIMES: case OP_LAND: case OP_MIN: return gbool(x.v & y.v);
            case OP_MINUS: case OP_RMINUS: case OP_LXOR: case OP_NE: case OP_ISNE: return gbool((x.v ^ y.v) != 0);
            case OP_LXNOR: case OP_EQ: case OP_ISEQ: return gbool(x.v == y.v);
            case OP_GT: return gbool(x.v > y.v);
            case OP_LT: return gbool(x.v < y.v);
            case OP_GE: return gbool(x.v >= y.v);
            case OP_LE: return gbool(x.v <= y.v);
            case OP_POW: return gbool(x.v || !y.v);
            case OP_RPOW: return gbool(y.v || !x.v);
        }
        return x;
    } else if constexpr (std::is_floating_point<T>::value) {
        switch (op) {
            case OP_FIRST: return x;
            case OP_SECOND: case OP_ANY: return y;   // any(x, y) may return either; y makes any(identity, p) = p
            case OP_PAIR: return (T)1;
            case OP_PLUS: return x + y;
            case OP_MINUS: return x - y;
            case OP_RMINUS: return y - x;
            case OP_TIMES: return x * y;
            case OP_DIV: return x / y;
            case OP_RDIV: return y / x;
            case OP_MIN: return (x < y || y != y) ? x : y;   // fmin semantics: NaN loses
            case OP_MAX: return (x > y || y != y) ? x : y;
            case OP_LOR: return (T)((x != 0) || (y != 0));
            case OP_LAND: return (T)((x != 0) && (y != 0));
            case OP_LXOR: return (T)((x != 0) != (y != 0));
            case OP_LXNOR: return (T)((x != 0) == (y != 0));
            case OP_ISEQ: case OP_EQ: return (T)(x == y);   // comparison codes: 1 / 0 in T (callers that need BOOL use cmpop)
            case OP_ISNE: case OP_NE: return (T)(x != y);
            case OP_GT: return (T)(x > y);
            case OP_LT: return (T)(x < y);
            case OP_GE: return (T)(x >= y);
            case OP_LE: return (T)(x <= y);
            case OP_POW: if constexpr (WITH_POW) return pow_op<T>(x, y); else break;
            case OP_RPOW: if constexpr (WITH_POW) return pow_op<T>(y, x); else break;
        }
        return x;
    } else {
        typedef typename std::make_unsigned<T>::type U;   // wrap-around arithmetic
        switch (op) {
            case OP_FIRST: return x;
            case OP_SECOND: case OP_ANY: return y;   // any(x, y) may return either; y makes any(identity, p) = p
            case OP_PAIR: return (T)1;
            case OP_PLUS: return (T)((U)x + (U)y);
            case OP_MINUS: return (T)((U)x - (U)y);
            case OP_RMINUS: return (T)((U)y - (U)x);
            case OP_TIMES: return (T)((U)x * (U)y);
            case OP_DIV:   // C division truncating toward zero; x/0 follows SuiteSparse's convention
                if (y == 0) return std::is_signed<T>::value ? (x == 0 ? (T)0 : (x < 0 ? (T)((U)1 << (sizeof(T) * 8 - 1)) : (T)(~((U)1 << (sizeof(T) * 8 - 1))))) : (x == 0 ? (T)0 : (T)~(U)0);
                if (std::is_signed<T>::value && y == (T)-1) return (T)((U)0 - (U)x);
                return (T)(x / y);
            case OP_RDIV:
                if (x == 0) return std::is_signed<T>::value ? (y == 0 ? (T)0 : (y < 0 ? (T)((U)1 << (sizeof(T) * 8 - 1)) : (T)(~((U)1 << (sizeof(T) * 8 - 1))))) : (y == 0 ? (T)0 : (T)~(U)0);
                if (std::is_signed<T>::value && x == (T)-1) return (T)((U)0 - (U)y);
                return (T)(y / x);
            case OP_MIN: return x < y ? x : y;
            case OP_MAX: return x > y ? x : y;
            case OP_LOR: return (T)((x != 0) || (y != 0));
            case OP_LAND: return (T)((x != 0) && (y != 0));
            case OP_LXOR: return (T)((x != 0) != (y != 0));
            case OP_LXNOR: return (T)((x != 0) == (y != 0));
            case OP_ISEQ: case OP_EQ: return (T)(x == y);
            case OP_ISNE: case OP_NE: return (T)(x != y);
            case OP_GT: return (T)(x > y);
            case OP_LT: return (T)(x < y);
            case OP_GE: return (T)(x >= y);
            case OP_LE: return (T)(x <= y);
            case OP_POW: if constexpr (WITH_POW) return pow_op<T>(x, y); else break;
            case OP_RPOW: if constexpr (WITH_POW) return pow_op<T>(y, x); else break;
        }
        return x;
    }
}

// comparison ops returning bool (for eWise EQ etc.)
template <typename T> __host__ __device__ __forceinline__ bool cmpop(int op, T x, T y) {
    if constexpr (is_gbool<T>::value) {
        switch (op) {
            case OP_EQ: return x.v == y.v; case OP_NE: return x.v != y.v; case OP_GT: return x.v > y.v;
            case OP_LT: return x.v < y.v; case OP_GE: return x.v >= y.v; case OP_LE: return x.v <= y.v;
        }
        return false;
    } else {
        switch (op) {
            case OP_EQ: return x == y; case OP_NE: return x != y; case OP_GT: return x > y;
            case OP_LT: return x < y; case OP_GE: return x >= y; case OP_LE: return x <= y;
        }
        return false;
    }
}

// ---- monoid identity ----
template <typename T> __host__ __device__ __forceinline__ T monoid_identity(int op) {
    if constexpr (is_gbool<T>::value) {
        switch (op) {
            case OP_TIMES: case OP_LAND: case OP_MIN: case OP_LXNOR: case OP_EQ: return gbool(true);
            default: return gbool(false);
        }
    } else if constexpr (std::is_floating_point<T>::value) {
        switch (op) {
            case OP_TIMES: case OP_LAND: return (T)1;
            case OP_MIN: return (T)INFINITY;
            case OP_MAX: return (T)-INFINITY;
            default: return (T)0;
        }
    } else {
        typedef typename std::make_unsigned<T>::type U;
        switch (op) {
            case OP_TIMES: case OP_LAND: return (T)1;
            case OP_MIN: return std::is_signed<T>::value ? (T)(~((U)1 << (sizeof(T) * 8 - 1))) : (T)~(U)0;
            case OP_MAX: return std::is_signed<T>::value ? (T)((U)1 << (sizeof(T) * 8 - 1)) : (T)0;
            default: return (T)0;
        }
    }
}

// ---- semiring functors ----
// compile-time: every method folds to straight-line code; flags let kernels skip loads the multiply ignores
template <int ADD, int MUL, typename T> struct SRStatic {
    typedef T value_type;
    static constexpr bool kStatic = true;
    static constexpr bool kReadsA = (MUL != OP_SECOND && MUL != OP_PAIR);   // matrix-side operand of mul(a, b)
    static constexpr bool kReadsB = (MUL != OP_FIRST && MUL != OP_PAIR);
    static constexpr bool kAddIsAny = (ADD == OP_ANY);
    __host__ __device__ __forceinline__ int add_op() const { return ADD; }
    __host__ __device__ __forceinline__ int mul_op() const { return MUL; }
    __host__ __device__ __forceinline__ T mul(T a, T b) const { return binop<T>(MUL, a, b); }
    __host__ __device__ __forceinline__ T add(T x, T y) const { return binop<T>(ADD, x, y); }
    __host__ __device__ __forceinline__ T identity() const { return monoid_identity<T>(ADD); }
    __host__ __device__ __forceinline__ bool reads_a() const { return kReadsA; }
    __host__ __device__ __forceinline__ bool reads_b() const { return kReadsB; }
};
// run-time op codes (uniform branches; the path for every builtin semiring outside the hot set)
template <typename T> struct SRDyn {
    typedef T value_type;
    static constexpr bool kStatic = false;
    static constexpr bool kReadsA = true;
    static constexpr bool kReadsB = true;
    static constexpr bool kAddIsAny = false;
    int a_op, m_op;
    __host__ __device__ __forceinline__ int add_op() const { return a_op; }
    __host__ __device__ __forceinline__ int mul_op() const { return m_op; }
    __host__ __device__ __forceinline__ T mul(T a, T b) const { return binop<T>(m_op, a, b); }
    __host__ __device__ __forceinline__ T add(T x, T y) const { return binop<T>(a_op, x, y); }
    __host__ __device__ __forceinline__ T identity() const { return monoid_identity<T>(a_op); }
    __host__ __device__ __forceinline__ bool reads_a() const { return m_op != OP_SECOND && m_op != OP_PAIR; }
    __host__ __device__ __forceinline__ bool reads_b() const { return m_op != OP_FIRST && m_op != OP_PAIR; }
};

// ---- atomic combine: *addr = add(*addr, v) for any monoid/type (global or shared memory) ----
template <typename W, typename T, typename F> __device__ __forceinline__ void atomic_cas_loop(T *addr, T v, F f) {
    // T is 4 or 8 bytes wide here
    W *wa = reinterpret_cast<W *>(addr);
    W old = *wa, assumed;
    do {
        assumed = old;
        T cur;
        memcpy(&cur, &assumed, sizeof(T));
        T nxt = f(cur, v);
        W nw;
        memcpy(&nw, &nxt, sizeof(T));
        if (nw == assumed) break;
        old = atomicCAS(wa, assumed, nw);
    } while (old != assumed);
}
// 1- and 2-byte types: CAS on the enclosing aligned 32-bit word
template <typename T, typename F> __device__ __forceinline__ void atomic_cas_small(T *addr, T v, F f) {
    uintptr_t a = reinterpret_cast<uintptr_t>(addr);
    unsigned int *wa = reinterpret_cast<unsigned int *>(a & ~(uintptr_t)3);
    unsigned int shift = (unsigned int)(a & 3) * 8;
    unsigned int mask = (sizeof(T) == 1 ? 0xffu : 0xffffu) << shift;
    unsigned int old = *wa, assumed;
    do {
        assumed = old;
        typename std::conditional<sizeof(T) == 1, uint8_t, uint16_t>::type bits = (assumed & mask) >> shift;
        T cur;
        memcpy(&cur, &bits, sizeof(T));
        T nxt = f(cur, v);
        typename std::conditional<sizeof(T) == 1, uint8_t, uint16_t>::type nb;
        memcpy(&nb, &nxt, sizeof(T));
        unsigned int nw = (assumed & ~mask) | ((unsigned int)nb << shift);
        if (nw == assumed) break;
        old = atomicCAS(wa, assumed, nw);
    } while (old != assumed);
}

template <typename SR, typename T> __device__ __forceinline__ void atomic_combine(const SR &sr, T *addr, T v) {
    const int op = sr.add_op();
    if constexpr (SR::kStatic) {
        // native fast paths for the hot monoids
        if constexpr (std::is_same<T, float>::value || std::is_same<T, double>::value) {
            if (op == OP_PLUS) { atomicAdd(addr, v); return; }
        }
        if constexpr (std::is_same<T, int32_t>::value) {
            if (op == OP_PLUS) { atomicAdd((int *)addr, (int)v); return; }
            if (op == OP_MIN) { atomicMin((int *)addr, (int)v); return; }
            if (op == OP_MAX) { atomicMax((int *)addr, (int)v); return; }
        }
        if constexpr (std::is_same<T, uint32_t>::value) {
            if (op == OP_PLUS) { atomicAdd((unsigned int *)addr, (unsigned int)v); return; }
            if (op == OP_MIN) { atomicMin((unsigned int *)addr, (unsigned int)v); return; }
            if (op == OP_MAX) { atomicMax((unsigned int *)addr, (unsigned int)v); return; }
        }
        if constexpr (std::is_same<T, int64_t>::value) {
            if (op == OP_PLUS) { atomicAdd((unsigned long long *)addr, (unsigned long long)v); return; }
            if (op == OP_MIN) { atomicMin((long long *)addr, (long long)v); return; }
            if (op == OP_MAX) { atomicMax((long long *)addr, (long long)v); return; }
        }
        if constexpr (std::is_same<T, uint64_t>::value) {
            if (op == OP_PLUS) { atomicAdd((unsigned long long *)addr, (unsigned long long)v); return; }
            if (op == OP_MIN) { atomicMin((unsigned long long *)addr, (unsigned long long)v); return; }
            if (op == OP_MAX) { atomicMax((unsigned long long *)addr, (unsigned long long)v); return; }
        }
        if (op == OP_ANY) { *addr = v; return; }   // any value is a valid result; plain store
    }
    auto f = [&](T a, T b) { return sr.add(a, b); };
    if constexpr (sizeof(T) == 8) atomic_cas_loop<unsigned long long>(addr, v, f);
    else if constexpr (sizeof(T) == 4) atomic_cas_loop<unsigned int>(addr, v, f);
    else atomic_cas_small(addr, v, f);
}

template <typename T> constexpr int type_code_of() {
    if (std::is_same<T, gbool>::value) return TC_BOOL;
    if (std::is_same<T, int8_t>::value) return TC_INT8;
    if (std::is_same<T, int16_t>::value) return TC_INT16;
    if (std::is_same<T, int32_t>::value) return TC_INT32;
    if (std::is_same<T, int64_t>::value) return TC_INT64;
    if (std::is_same<T, uint8_t>::value) return TC_UINT8;
    if (std::is_same<T, uint16_t>::value) return TC_UINT16;
    if (std::is_same<T, uint32_t>::value) return TC_UINT32;
    if (std::is_same<T, uint64_t>::value) return TC_UINT64;
    if (std::is_same<T, float>::value) return TC_FP32;
    return TC_FP64;
}

// ---- host-side dispatch helpers ----
#define GRB_DISPATCH_TYPE(tc, T, ...)                                                         \
    switch (tc) {                                                                             \
        case TC_BOOL: { typedef gbool T; __VA_ARGS__; } break;                                \
        case TC_INT8: { typedef int8_t T; __VA_ARGS__; } break;                               \
        case TC_INT16: { typedef int16_t T; __VA_ARGS__; } break;                             \
        case TC_INT32: { typedef int32_t T; __VA_ARGS__; } break;                             \
        case TC_INT64: { typedef int64_t T; __VA_ARGS__; } break;                             \
        case TC_UINT8: { typedef uint8_t T; __VA_ARGS__; } break;                             \
        case TC_UINT16: { typedef uint16_t T; __VA_ARGS__; } break;                           \
        case TC_UINT32: { typedef uint32_t T; __VA_ARGS__; } break;                           \
        case TC_UINT64: { typedef uint64_t T; __VA_ARGS__; } break;                           \
        case TC_FP32: { typedef float T; __VA_ARGS__; } break;                                \
        case TC_FP64: { typedef double T; __VA_ARGS__; } break;                               \
        default: break;                                                                       \
    }

// Semiring dispatch: the hot set gets SRStatic instantiations, everything else SRDyn<T>.
// Usage: GRB_DISPATCH_SEMIRING(add, mul, T, SRT, sr, body-using-SRT-and-sr)
#define GRB_SR_CASE(A, M, T, SRT, sr, ...)                                                    \
    if (_add == (A) && _mul == (M)) { typedef SRStatic<A, M, T> SRT; SRT sr; __VA_ARGS__; } else
#define GRB_DISPATCH_SEMIRING(add, mul, T, SRT, sr, ...)                                      \
    do {                                                                                      \
        const int _add = (add), _mul = (mul);                                                 \
        if constexpr (is_gbool<T>::value) {                                                   \
            GRB_SR_CASE(OP_LOR, OP_LAND, T, SRT, sr, __VA_ARGS__)                             \
            GRB_SR_CASE(OP_ANY, OP_PAIR, T, SRT, sr, __VA_ARGS__)                             \
            { typedef SRDyn<T> SRT; SRT sr; sr.a_op = _add; sr.m_op = _mul; __VA_ARGS__; }    \
        } else {                                                                              \
            GRB_SR_CASE(OP_PLUS, OP_TIMES, T, SRT, sr, __VA_ARGS__)                           \
            GRB_SR_CASE(OP_MIN, OP_PLUS, T, SRT, sr, __VA_ARGS__)                             \
            GRB_SR_CASE(OP_PLUS, OP_SECOND, T, SRT, sr, __VA_ARGS__)                          \
            GRB_SR_CASE(OP_PLUS, OP_FIRST, T, SRT, sr, __VA_ARGS__)                           \
            GRB_SR_CASE(OP_ANY, OP_PAIR, T, SRT, sr, __VA_ARGS__)                             \
            { typedef SRDyn<T> SRT; SRT sr; sr.a_op = _add; sr.m_op = _mul; __VA_ARGS__; }    \
        }                                                                                     \
    } while (0)
