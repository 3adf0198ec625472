/* ORACLE -- test infrastructure only.  Never linked, imported or executed by the
 * product path (python-graphblas_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement (plain C + OpenMP) of the multiply  T = A (+).(x) B  that
 * GrB_mxm / GrB_mxv / GrB_vxm perform inside SuiteSparse:GraphBLAS, the
 * third-party C library python-graphblas calls at graphblas/core/base.py:27
 * (reached from core/base.py:503 with the names set at core/matrix.py:2254,2321
 * and core/vector.py:1370).  That library (PyPI suitesparse-graphblas >=7.4.0.0,
 * reference pyproject.toml:64) is NOT under /root/reference and is not
 * installed, so this file restates the published algorithm of the GraphBLAS C
 * API 2.0 (reference README.md:52, docs/user_guide/operations.rst:4-153):
 *
 *   T(i,j) present  <=>  exists k : A(i,k) present and B(k,j) present
 *   T(i,j) = (+)_k A(i,k) (x) B(k,j)   over present k only; integers wrap;
 *   no entry is dropped for being zero / the identity.
 *
 * Mask / accum / replace write-back is done on top of T by oracle/bigref.py
 * (vectorised numpy) and, for small cases, by oracle/semantics.py.  Both are
 * pinned against the reference's known-answer tests in tests/golden/.
 *
 * Methods: row-wise Gustavson with a dense accumulator per thread (SpGEMM),
 * row-wise dot (mxv, "pull") and sequential scatter (vxm, "push").
 * Indices are int64 throughout; values are passed as raw typed arrays.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* type codes -- identical to include/grb_cuda.h GrB_Type codes */
enum { T_BOOL = 0, T_INT8, T_INT16, T_INT32, T_INT64, T_UINT8, T_UINT16, T_UINT32, T_UINT64, T_FP32, T_FP64 };
/* binary op codes used for the monoid (add) and the multiply */
enum { OP_FIRST = 1, OP_SECOND, OP_PAIR, OP_PLUS, OP_MINUS, OP_TIMES, OP_DIV, OP_MIN, OP_MAX, OP_LOR, OP_LAND, OP_LXOR, OP_ANY };

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- typed binary ops.  Integer arithmetic is done in the unsigned type of the same
 * width so that overflow wraps (two's complement), matching C-cast semantics. ---- */
#define DEF_INT_OPS(T, UT, SFX)                                                                   \
    static inline T op_##SFX(int op, T x, T y) {                                                  \
        switch (op) {                                                                             \
            case OP_FIRST: return x;                                                              \
            case OP_SECOND: return y;                                                             \
            case OP_PAIR: return (T)1;                                                            \
            case OP_PLUS: return (T)((UT)x + (UT)y);                                              \
            case OP_MINUS: return (T)((UT)x - (UT)y);                                             \
            case OP_TIMES: return (T)((UT)x * (UT)y);                                             \
            case OP_MIN: return x < y ? x : y;                                                    \
            case OP_MAX: return x > y ? x : y;                                                    \
            case OP_LOR: return (T)((x != 0) || (y != 0));                                        \
            case OP_LAND: return (T)((x != 0) && (y != 0));                                       \
            case OP_LXOR: return (T)((x != 0) != (y != 0));                                       \
            case OP_ANY: return x;                                                                \
        }                                                                                         \
        return x;                                                                                 \
    }
#define DEF_FP_OPS(T, SFX)                                                                        \
    static inline T op_##SFX(int op, T x, T y) {                                                  \
        switch (op) {                                                                             \
            case OP_FIRST: return x;                                                              \
            case OP_SECOND: return y;                                                             \
            case OP_PAIR: return (T)1;                                                            \
            case OP_PLUS: return x + y;                                                           \
            case OP_MINUS: return x - y;                                                          \
            case OP_TIMES: return x * y;                                                          \
            case OP_DIV: return x / y;                                                            \
            case OP_MIN: return (x < y || y != y) ? x : y;                                        \
            case OP_MAX: return (x > y || y != y) ? x : y;                                        \
            case OP_LOR: return (T)((x != 0) || (y != 0));                                        \
            case OP_LAND: return (T)((x != 0) && (y != 0));                                       \
            case OP_LXOR: return (T)((x != 0) != (y != 0));                                       \
            case OP_ANY: return x;                                                                \
        }                                                                                         \
        return x;                                                                                 \
    }
static inline uint8_t op_b(int op, uint8_t x, uint8_t y) {
    switch (op) {
        case OP_FIRST: return x;
        case OP_SECOND: return y;
        case OP_PAIR: return 1;
        case OP_PLUS: case OP_LOR: case OP_MAX: return (uint8_t)(x | y);
        case OP_TIMES: case OP_LAND: case OP_MIN: return (uint8_t)(x & y);
        case OP_MINUS: case OP_LXOR: return (uint8_t)(x ^ y);
        case OP_ANY: return x;
    }
    return x;
}
DEF_INT_OPS(int8_t, uint8_t, i8)
DEF_INT_OPS(int16_t, uint16_t, i16)
DEF_INT_OPS(int32_t, uint32_t, i32)
DEF_INT_OPS(int64_t, uint64_t, i64)
DEF_INT_OPS(uint8_t, uint8_t, u8)
DEF_INT_OPS(uint16_t, uint16_t, u16)
DEF_INT_OPS(uint32_t, uint32_t, u32)
DEF_INT_OPS(uint64_t, uint64_t, u64)
DEF_FP_OPS(float, f32)
DEF_FP_OPS(double, f64)


/* hot semirings get a compile-time specialised copy of every loop (so the CPU baseline is
 * not slowed by the op switch); everything else uses the generic copy. */
#define PICK(FN, SFX, ...)                                                                        \
    do {                                                                                          \
        if (add == OP_PLUS && mul == OP_TIMES) FN##_##SFX##_pt(__VA_ARGS__);                      \
        else if (add == OP_MIN && mul == OP_PLUS) FN##_##SFX##_mp(__VA_ARGS__);                   \
        else if (add == OP_PLUS && mul == OP_SECOND) FN##_##SFX##_ps(__VA_ARGS__);                \
        else if (add == OP_PLUS && mul == OP_FIRST) FN##_##SFX##_pf(__VA_ARGS__);                 \
        else if (add == OP_ANY && mul == OP_PAIR) FN##_##SFX##_ap(__VA_ARGS__);                   \
        else if (add == OP_LOR && mul == OP_LAND) FN##_##SFX##_ll(__VA_ARGS__);                   \
        else FN##_##SFX##_gen(__VA_ARGS__);                                                       \
    } while (0)
#define PICK_RC(RC, FN, SFX, ...)                                                                 \
    do {                                                                                          \
        if (add == OP_PLUS && mul == OP_TIMES) RC = FN##_##SFX##_pt(__VA_ARGS__);                 \
        else if (add == OP_MIN && mul == OP_PLUS) RC = FN##_##SFX##_mp(__VA_ARGS__);              \
        else if (add == OP_PLUS && mul == OP_SECOND) RC = FN##_##SFX##_ps(__VA_ARGS__);           \
        else if (add == OP_PLUS && mul == OP_FIRST) RC = FN##_##SFX##_pf(__VA_ARGS__);            \
        else if (add == OP_ANY && mul == OP_PAIR) RC = FN##_##SFX##_ap(__VA_ARGS__);              \
        else if (add == OP_LOR && mul == OP_LAND) RC = FN##_##SFX##_ll(__VA_ARGS__);              \
        else RC = FN##_##SFX##_gen(__VA_ARGS__);                                                  \
    } while (0)

/* =====================================================================================
 * mxv (pull):  t(i) = (+)_k  mul(A(i,k), x(k))      flip=0
 *              t(i) = (+)_k  mul(x(k), A(i,k))      flip=1  (vxm with A transposed: u'A' = A u)
 * xp == NULL means x is full.  tp[i] = 1 iff row i produced an entry.
 * ===================================================================================== */
#define DEF_MXV_K(T, SFX, K, ADDE, MULE)                                                                        \
    static void mxv_##SFX##_##K(int add, int mul, int flip, int64_t nrows, const int64_t *Ap,           \
                          const int64_t *Aj, const T *Ax, const T *x, const uint8_t *xp, T *tv,   \
                          uint8_t *tp) {                                                          \
        _Pragma("omp parallel for schedule(dynamic, 1024)") for (int64_t i = 0; i < nrows; i++) { \
            T acc = 0;                                                                            \
            int has = 0;                                                                          \
            for (int64_t p = Ap[i]; p < Ap[i + 1]; p++) {                                         \
                int64_t k = Aj[p];                                                                \
                if (xp && !xp[k]) continue;                                                       \
                T a = Ax ? Ax[p] : (T)1;                                                          \
                T prod = flip ? op_##SFX(MULE, x[k], a) : op_##SFX(MULE, a, x[k]);                  \
                if (has) acc = op_##SFX(ADDE, acc, prod);                                          \
                else { acc = prod; has = 1; }                                                     \
            }                                                                                     \
            tv[i] = acc;                                                                          \
            tp[i] = (uint8_t)has;                                                                 \
        }                                                                                         \
    }
#define DEF_MXV(T, SFX) \
    DEF_MXV_K(T, SFX, gen, add, mul) \
    DEF_MXV_K(T, SFX, pt, OP_PLUS, OP_TIMES) \
    DEF_MXV_K(T, SFX, mp, OP_MIN, OP_PLUS) \
    DEF_MXV_K(T, SFX, ps, OP_PLUS, OP_SECOND) \
    DEF_MXV_K(T, SFX, pf, OP_PLUS, OP_FIRST) \
    DEF_MXV_K(T, SFX, ap, OP_ANY, OP_PAIR) \
    DEF_MXV_K(T, SFX, ll, OP_LOR, OP_LAND)
DEF_MXV(uint8_t, b)
DEF_MXV(int8_t, i8)
DEF_MXV(int16_t, i16)
DEF_MXV(int32_t, i32)
DEF_MXV(int64_t, i64)
DEF_MXV(uint8_t, u8)
DEF_MXV(uint16_t, u16)
DEF_MXV(uint32_t, u32)
DEF_MXV(uint64_t, u64)
DEF_MXV(float, f32)
DEF_MXV(double, f64)

#define DISPATCH_TYPE(type, CALL)                                                                 \
    switch (type) {                                                                               \
        case T_BOOL: CALL(uint8_t, b); break;                                                     \
        case T_INT8: CALL(int8_t, i8); break;                                                     \
        case T_INT16: CALL(int16_t, i16); break;                                                  \
        case T_INT32: CALL(int32_t, i32); break;                                                  \
        case T_INT64: CALL(int64_t, i64); break;                                                  \
        case T_UINT8: CALL(uint8_t, u8); break;                                                   \
        case T_UINT16: CALL(uint16_t, u16); break;                                                \
        case T_UINT32: CALL(uint32_t, u32); break;                                                \
        case T_UINT64: CALL(uint64_t, u64); break;                                                \
        case T_FP32: CALL(float, f32); break;                                                     \
        case T_FP64: CALL(double, f64); break;                                                    \
        default: return -3;                                                                       \
    }

int oracle_mxv(int add, int mul, int type, int flip, int64_t nrows, const int64_t *Ap,
               const int64_t *Aj, const void *Ax, const void *x, const uint8_t *xp, void *tv,
               uint8_t *tp) {
#define CALL(T, SFX) PICK(mxv, SFX, add, mul, flip, nrows, Ap, Aj, (const T *)Ax, (const T *)x, xp, (T *)tv, tp)
    DISPATCH_TYPE(type, CALL)
#undef CALL
    return 0;
}

/* =====================================================================================
 * vxm (push):  t(j) = (+)_i  mul(u(i), A(i,j))   flip=0
 *              t(j) = (+)_i  mul(A(i,j), u(i))   flip=1  (mxv with A transposed)
 * Sequential scatter in increasing i: deterministic accumulation order.
 * ===================================================================================== */
#define DEF_PUSH_K(T, SFX, K, ADDE, MULE)                                                                       \
    static void push_##SFX##_##K(int add, int mul, int flip, int64_t nrows, int64_t ncols,              \
                           const int64_t *Ap, const int64_t *Aj, const T *Ax, const T *u,         \
                           const uint8_t *up, T *tv, uint8_t *tp) {                               \
        memset(tp, 0, (size_t)ncols);                                                             \
        for (int64_t i = 0; i < nrows; i++) {                                                     \
            if (up && !up[i]) continue;                                                           \
            for (int64_t p = Ap[i]; p < Ap[i + 1]; p++) {                                         \
                int64_t j = Aj[p];                                                                \
                T a = Ax ? Ax[p] : (T)1;                                                          \
                T prod = flip ? op_##SFX(MULE, a, u[i]) : op_##SFX(MULE, u[i], a);                  \
                if (tp[j]) tv[j] = op_##SFX(ADDE, tv[j], prod);                                    \
                else { tv[j] = prod; tp[j] = 1; }                                                 \
            }                                                                                     \
        }                                                                                         \
    }
#define DEF_PUSH(T, SFX) \
    DEF_PUSH_K(T, SFX, gen, add, mul) \
    DEF_PUSH_K(T, SFX, pt, OP_PLUS, OP_TIMES) \
    DEF_PUSH_K(T, SFX, mp, OP_MIN, OP_PLUS) \
    DEF_PUSH_K(T, SFX, ps, OP_PLUS, OP_SECOND) \
    DEF_PUSH_K(T, SFX, pf, OP_PLUS, OP_FIRST) \
    DEF_PUSH_K(T, SFX, ap, OP_ANY, OP_PAIR) \
    DEF_PUSH_K(T, SFX, ll, OP_LOR, OP_LAND)
DEF_PUSH(uint8_t, b)
DEF_PUSH(int8_t, i8)
DEF_PUSH(int16_t, i16)
DEF_PUSH(int32_t, i32)
DEF_PUSH(int64_t, i64)
DEF_PUSH(uint8_t, u8)
DEF_PUSH(uint16_t, u16)
DEF_PUSH(uint32_t, u32)
DEF_PUSH(uint64_t, u64)
DEF_PUSH(float, f32)
DEF_PUSH(double, f64)

int oracle_vxm_push(int add, int mul, int type, int flip, int64_t nrows, int64_t ncols,
                    const int64_t *Ap, const int64_t *Aj, const void *Ax, const void *u,
                    const uint8_t *up, void *tv, uint8_t *tp) {
#define CALL(T, SFX) PICK(push, SFX, add, mul, flip, nrows, ncols, Ap, Aj, (const T *)Ax, (const T *)u, up, (T *)tv, tp)
    DISPATCH_TYPE(type, CALL)
#undef CALL
    return 0;
}

/* =====================================================================================
 * mxm: two-phase Gustavson.  Rows [r0, r1) of A only (bounded samples for the CPU baseline).
 *   count:  Cp_out[i - r0] = nnz of row i of A*B            (i in [r0, r1))
 *   fill :  given Cp (exclusive scan, Cp[0] = 0), write sorted Cj and Cx.
 * ===================================================================================== */
int oracle_mxm_count(int64_t r0, int64_t r1, int64_t ncolsB, const int64_t *Ap, const int64_t *Aj,
                     const int64_t *Bp, const int64_t *Bj, int64_t *row_nnz) {
    int err = 0;
#pragma omp parallel
    {
        int64_t *mark = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncolsB > 0 ? ncolsB : 1));
        if (!mark) {
#pragma omp atomic write
            err = -102;
        } else {
            for (int64_t j = 0; j < ncolsB; j++) mark[j] = -1;
#pragma omp for schedule(dynamic, 256)
            for (int64_t i = r0; i < r1; i++) {
                int64_t cnt = 0;
                for (int64_t p = Ap[i]; p < Ap[i + 1]; p++) {
                    int64_t k = Aj[p];
                    for (int64_t q = Bp[k]; q < Bp[k + 1]; q++) {
                        int64_t j = Bj[q];
                        if (mark[j] != i) { mark[j] = i; cnt++; }
                    }
                }
                row_nnz[i - r0] = cnt;
            }
            free(mark);
        }
    }
    return err;
}

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

#define DEF_MXM_K(T, SFX, K, ADDE, MULE)                                                                        \
    static int mxm_fill_##SFX##_##K(int add, int mul, int64_t r0, int64_t r1, int64_t ncolsB,           \
                              const int64_t *Ap, const int64_t *Aj, const T *Ax,                  \
                              const int64_t *Bp, const int64_t *Bj, const T *Bx,                  \
                              const int64_t *Cp, int64_t *Cj, T *Cx) {                            \
        int err = 0;                                                                              \
        _Pragma("omp parallel") {                                                                 \
            size_t nc = (size_t)(ncolsB > 0 ? ncolsB : 1);                                        \
            int64_t *mark = (int64_t *)malloc(sizeof(int64_t) * nc);                              \
            T *acc = (T *)malloc(sizeof(T) * nc);                                                 \
            if (!mark || !acc) {                                                                  \
                _Pragma("omp atomic write") err = -102;                                           \
            } else {                                                                              \
                for (int64_t j = 0; j < ncolsB; j++) mark[j] = -1;                                \
                _Pragma("omp for schedule(dynamic, 256)") for (int64_t i = r0; i < r1; i++) {     \
                    int64_t base = Cp[i - r0], cnt = 0;                                           \
                    for (int64_t p = Ap[i]; p < Ap[i + 1]; p++) {                                 \
                        int64_t k = Aj[p];                                                        \
                        T a = Ax ? Ax[p] : (T)1;                                                  \
                        for (int64_t q = Bp[k]; q < Bp[k + 1]; q++) {                             \
                            int64_t j = Bj[q];                                                    \
                            T prod = op_##SFX(MULE, a, Bx ? Bx[q] : (T)1);                         \
                            if (mark[j] != i) { mark[j] = i; acc[j] = prod; Cj[base + cnt++] = j; } \
                            else acc[j] = op_##SFX(ADDE, acc[j], prod);                            \
                        }                                                                         \
                    }                                                                             \
                    qsort(Cj + base, (size_t)cnt, sizeof(int64_t), cmp_i64);                      \
                    for (int64_t c = 0; c < cnt; c++) Cx[base + c] = acc[Cj[base + c]];           \
                }                                                                                 \
            }                                                                                     \
            free(mark);                                                                           \
            free(acc);                                                                            \
        }                                                                                         \
        return err;                                                                               \
    }
#define DEF_MXM(T, SFX) \
    DEF_MXM_K(T, SFX, gen, add, mul) \
    DEF_MXM_K(T, SFX, pt, OP_PLUS, OP_TIMES) \
    DEF_MXM_K(T, SFX, mp, OP_MIN, OP_PLUS) \
    DEF_MXM_K(T, SFX, ps, OP_PLUS, OP_SECOND) \
    DEF_MXM_K(T, SFX, pf, OP_PLUS, OP_FIRST) \
    DEF_MXM_K(T, SFX, ap, OP_ANY, OP_PAIR) \
    DEF_MXM_K(T, SFX, ll, OP_LOR, OP_LAND)
DEF_MXM(uint8_t, b)
DEF_MXM(int8_t, i8)
DEF_MXM(int16_t, i16)
DEF_MXM(int32_t, i32)
DEF_MXM(int64_t, i64)
DEF_MXM(uint8_t, u8)
DEF_MXM(uint16_t, u16)
DEF_MXM(uint32_t, u32)
DEF_MXM(uint64_t, u64)
DEF_MXM(float, f32)
DEF_MXM(double, f64)

int oracle_mxm_fill(int add, int mul, int type, int64_t r0, int64_t r1, int64_t ncolsB,
                    const int64_t *Ap, const int64_t *Aj, const void *Ax, const int64_t *Bp,
                    const int64_t *Bj, const void *Bx, const int64_t *Cp, int64_t *Cj, void *Cx) {
    int rc = 0;
#define CALL(T, SFX) PICK_RC(rc, mxm_fill, SFX, add, mul, r0, r1, ncolsB, Ap, Aj, (const T *)Ax, Bp, Bj, (const T *)Bx, Cp, Cj, (T *)Cx)
    DISPATCH_TYPE(type, CALL)
#undef CALL
    return rc;
}


/* =====================================================================================
 * CPU BASELINE for bench.py (plus_times fp32 only): the same product the way a tuned CPU library organises it, so that the
 * reported CPU number is not handicapped by the parity oracle's conveniences (two passes, 64-bit indices, sorted rows):
 *   one numeric pass, rows left unsorted (as SuiteSparse:GraphBLAS leaves them, lazily), 32-bit column indices, and per row
 *   the accumulator SuiteSparse's saxpy3 method would pick -- an open-addressing hash table that stays in L1/L2 for rows with
 *   few products, the dense Gustavson workspace for rows with many.  Output rows go to a staging CSR addressed by the flops
 *   prefix (upper bound per row); row_nnz receives the exact counts.  Returns 0, or -102 when a workspace cannot be allocated.
 * Semantics identical to mxm_fill (verified by tests/test_oracle.py::test_fast_cpu_baseline_matches_oracle).
 * ===================================================================================== */
static inline uint32_t hash32(uint32_t k) { return k * 0x9E3779B1u; }

int oracle_mxm_baseline_f32(int64_t r0, int64_t r1, int64_t ncolsB, const int64_t *Ap, const int32_t *Aj, const float *Ax,
                            const int64_t *Bp, const int32_t *Bj, const float *Bx, const int64_t *Sp /* flops prefix, r1-r0+1 */,
                            int32_t *Cj, float *Cx, int64_t *row_nnz, int64_t hash_max_flops) {
    int err = 0;
#pragma omp parallel
    {
        const size_t nc = (size_t)(ncolsB > 0 ? ncolsB : 1);
        int32_t *mark = (int32_t *)malloc(sizeof(int32_t) * nc);     /* dense workspace: row stamp per column */
        float *acc = (float *)malloc(sizeof(float) * nc);
        int64_t hcap = 16;
        while (hcap < 2 * hash_max_flops) hcap <<= 1;
        int32_t *hk = (int32_t *)malloc(sizeof(int32_t) * (size_t)hcap);   /* hash workspace: sized per row inside this block */
        float *hv = (float *)malloc(sizeof(float) * (size_t)hcap);
        if (!mark || !acc || !hk || !hv) {
#pragma omp atomic write
            err = -102;
        } else {
            for (size_t j = 0; j < nc; j++) mark[j] = -1;
#pragma omp for schedule(dynamic, 256)
            for (int64_t i = r0; i < r1; i++) {
                const int64_t base = Sp[i - r0], flops = Sp[i - r0 + 1] - base;
                int64_t cnt = 0;
                if (flops == 0) { row_nnz[i - r0] = 0; continue; }
                if (flops <= hash_max_flops) {
                    uint32_t size = 16;
                    while (size < 2 * (uint64_t)flops) size <<= 1;
                    const uint32_t mask = size - 1;
                    int shift = 32;
                    for (uint32_t t = size; t > 1; t >>= 1) shift--;
                    for (uint32_t t = 0; t < size; t++) hk[t] = -1;
                    for (int64_t p = Ap[i]; p < Ap[i + 1]; p++) {
                        const int32_t k = Aj[p];
                        const float a = Ax[p];
                        for (int64_t q = Bp[k]; q < Bp[k + 1]; q++) {
                            const int32_t j = Bj[q];
                            const float prod = a * Bx[q];
                            uint32_t h = hash32((uint32_t)j) >> shift;
                            while (hk[h] != j && hk[h] != -1) h = (h + 1) & mask;
                            if (hk[h] == j) hv[h] += prod;
                            else { hk[h] = j; hv[h] = prod; cnt++; }
                        }
                    }
                    int64_t out = base;
                    for (uint32_t t = 0; t < size; t++)
                        if (hk[t] != -1) { Cj[out] = hk[t]; Cx[out] = hv[t]; out++; }
                } else {
                    const int32_t stamp = (int32_t)(i - r0);
                    for (int64_t p = Ap[i]; p < Ap[i + 1]; p++) {
                        const int32_t k = Aj[p];
                        const float a = Ax[p];
                        for (int64_t q = Bp[k]; q < Bp[k + 1]; q++) {
                            const int32_t j = Bj[q];
                            const float prod = a * Bx[q];
                            if (mark[j] != stamp) { mark[j] = stamp; acc[j] = prod; Cj[base + cnt++] = j; }
                            else acc[j] += prod;
                        }
                    }
                    for (int64_t c = 0; c < cnt; c++) Cx[base + c] = acc[Cj[base + c]];
                }
                row_nnz[i - r0] = cnt;
            }
        }
        free(mark); free(acc); free(hk); free(hv);
    }
    return err;
}

/* flops(i) = sum over A(i,:) of nnz(B(k,:)) for rows [r0, r1): written to flops[i - r0] */
void oracle_row_flops32(int64_t r0, int64_t r1, const int64_t *Ap, const int32_t *Aj, const int64_t *Bp, int64_t *flops) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = r0; i < r1; i++) {
        int64_t f = 0;
        for (int64_t p = Ap[i]; p < Ap[i + 1]; p++) f += Bp[Aj[p] + 1] - Bp[Aj[p]];
        flops[i - r0] = f;
    }
}
