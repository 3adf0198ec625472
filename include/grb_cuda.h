/* grb_cuda.h -- C ABI of libgrb_cuda.so, the B200-native GraphBLAS semiring engine.
 *
 * This is the drop-in boundary: the entry points are the subset of the
 * GraphBLAS C API 2.0 that python-graphblas's "vanilla" backend consumes for the
 * GrB_mxm / GrB_mxv / GrB_vxm path, with the same names, argument order and GrB_Info
 * convention, so the reference's dispatcher
 *
 *     graphblas/core/base.py:23-54   call(cfunc_name, args) -> getattr(lib, cfunc_name)(*cargs)
 *
 * can bind them unchanged (binding shown in INTEGRATION.md).  Plain pointers and sizes only;
 * no torch / C++ types.  All index arrays crossing this boundary are GrB_Index (uint64_t)
 * exactly as the reference hands them over (graphblas/core/utils.py:58-69).
 *
 * Each declaration cites the reference call site it serves.
 */
#ifndef GRB_CUDA_H
#define GRB_CUDA_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t GrB_Index;

/* GrB_Info values: GraphBLAS C API 2.0; consumed by graphblas/exceptions.py:123-157 */
typedef enum {
    GrB_SUCCESS = 0,
    GrB_NO_VALUE = 1,
    GrB_UNINITIALIZED_OBJECT = -1,
    GrB_NULL_POINTER = -2,
    GrB_INVALID_VALUE = -3,
    GrB_INVALID_INDEX = -4,
    GrB_DOMAIN_MISMATCH = -5,
    GrB_DIMENSION_MISMATCH = -6,
    GrB_OUTPUT_NOT_EMPTY = -7,
    GrB_NOT_IMPLEMENTED = -8,
    GrB_PANIC = -101,
    GrB_OUT_OF_MEMORY = -102,
    GrB_INSUFFICIENT_SPACE = -103,
    GrB_INVALID_OBJECT = -104,
    GrB_INDEX_OUT_OF_BOUNDS = -105,
    GrB_EMPTY_OBJECT = -106
} GrB_Info;

typedef enum { GrB_NONBLOCKING = 0, GrB_BLOCKING = 1 } GrB_Mode;
typedef enum { GrB_COMPLETE = 0, GrB_MATERIALIZE = 1 } GrB_WaitMode;
typedef enum { GrB_CSR_FORMAT = 0, GrB_CSC_FORMAT = 1, GrB_COO_FORMAT = 2 } GrB_Format;
typedef enum { GrB_OUTP = 0, GrB_MASK = 1, GrB_INP0 = 2, GrB_INP1 = 3 } GrB_Desc_Field;
typedef enum { GrB_DEFAULT = 0, GrB_REPLACE = 1, GrB_COMP = 2, GrB_TRAN = 3, GrB_STRUCTURE = 4 } GrB_Desc_Value;

/* opaque handles (graphblas/core/matrix.py:196 holds them in 1-element cells) */
typedef struct GrB_Type_opaque *GrB_Type;
typedef struct GrB_UnaryOp_opaque *GrB_UnaryOp;
typedef struct GrB_BinaryOp_opaque *GrB_BinaryOp;
typedef struct GrB_Monoid_opaque *GrB_Monoid;
typedef struct GrB_Semiring_opaque *GrB_Semiring;
typedef struct GrB_Descriptor_opaque *GrB_Descriptor;
typedef struct GrB_Matrix_opaque *GrB_Matrix;
typedef struct GrB_Vector_opaque *GrB_Vector;
typedef struct GrB_Scalar_opaque *GrB_Scalar;
typedef struct GrB_IndexUnaryOp_opaque *GrB_IndexUnaryOp;

/* ------------------------------------------------------------------ context
 * graphblas/__init__.py:143,158-173 initialize(blocking=...), is_initialized() */
GrB_Info GrB_init(GrB_Mode mode);
GrB_Info GrB_finalize(void);
GrB_Info GrB_getVersion(unsigned int *version, unsigned int *subversion);

/* ------------------------------------------------------------------ builtin objects
 * Every builtin type / operator / monoid / semiring / descriptor is also exported as a data
 * symbol with its C-API name (GrB_INT64, GrB_PLUS_TIMES_SEMIRING_FP32, GxB_ANY_PAIR_INT64,
 * GrB_DESC_RSC, ...) because the reference's operator registry discovers them by scanning
 * dir(lib) (graphblas/core/operator/base.py:690,803-893; semiring.py:185-219).
 * GrB_cuda_lookup returns the same handle by name (NULL if unknown); GrB_cuda_symbol_names
 * fills a NUL-separated list (returns required size) so a binding can enumerate them. */
void *GrB_cuda_lookup(const char *name);
size_t GrB_cuda_symbol_names(char *buf, size_t buflen);

extern GrB_Type GrB_BOOL, GrB_INT8, GrB_INT16, GrB_INT32, GrB_INT64, GrB_UINT8, GrB_UINT16, GrB_UINT32,
    GrB_UINT64, GrB_FP32, GrB_FP64;

/* ------------------------------------------------------------------ Matrix lifecycle
 * graphblas/core/matrix.py:190-203 (new), :218-225 (free), :482-494 (nvals), :764-789 (wait) */
GrB_Info GrB_Matrix_new(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols);
GrB_Info GrB_Matrix_free(GrB_Matrix *A);
GrB_Info GrB_Matrix_dup(GrB_Matrix *C, const GrB_Matrix A);
GrB_Info GrB_Matrix_clear(GrB_Matrix A);
GrB_Info GrB_Matrix_nrows(GrB_Index *nrows, const GrB_Matrix A);
GrB_Info GrB_Matrix_ncols(GrB_Index *ncols, const GrB_Matrix A);
GrB_Info GrB_Matrix_nvals(GrB_Index *nvals, const GrB_Matrix A);
GrB_Info GrB_Matrix_wait(GrB_Matrix A, GrB_WaitMode mode);
GrB_Info GrB_Matrix_error(const char **error, const GrB_Matrix A); /* graphblas/exceptions.py:171-189 */

/* ------------------------------------------------------------------ Vector lifecycle
 * graphblas/core/vector.py:159-170 (new), :184-191 (free) */
GrB_Info GrB_Vector_new(GrB_Vector *v, GrB_Type type, GrB_Index n);
GrB_Info GrB_Vector_free(GrB_Vector *v);
GrB_Info GrB_Vector_dup(GrB_Vector *w, const GrB_Vector u);
GrB_Info GrB_Vector_clear(GrB_Vector v);
GrB_Info GrB_Vector_size(GrB_Index *n, const GrB_Vector v);
GrB_Info GrB_Vector_nvals(GrB_Index *nvals, const GrB_Vector v);
GrB_Info GrB_Vector_wait(GrB_Vector v, GrB_WaitMode mode);
GrB_Info GrB_Vector_error(const char **error, const GrB_Vector v);

/* ------------------------------------------------------------------ the hot path
 * GrB_mxm : graphblas/core/matrix.py:2319-2328 (and core/vector.py:1778-1786 outer)
 * GrB_mxv : graphblas/core/matrix.py:2252-2259
 * GrB_vxm : graphblas/core/vector.py:1368-1375 (and :1734-1741 inner)
 * argument list assembled at graphblas/core/base.py:496-503:  [C, mask, accum, op, A, B, desc] */
GrB_Info GrB_mxm(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_Semiring op,
                 const GrB_Matrix A, const GrB_Matrix B, const GrB_Descriptor desc);
GrB_Info GrB_mxv(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring op,
                 const GrB_Matrix A, const GrB_Vector u, const GrB_Descriptor desc);
GrB_Info GrB_vxm(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring op,
                 const GrB_Vector u, const GrB_Matrix A, const GrB_Descriptor desc);

/* ------------------------------------------------------------------ O(n) vector operations that keep
 * BFS / SSSP / PageRank iterations on the device
 * eWiseAdd/eWiseMult: graphblas/core/vector.py:1050-1055,1142 ; apply: core/matrix.py:2440-2533 ;
 * reduce: core/vector.py:1669-1681 ; assign scalar under mask: core/vector.py:2020-2035 */
GrB_Info GrB_Vector_eWiseAdd_BinaryOp(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum,
                                      const GrB_BinaryOp op, const GrB_Vector u, const GrB_Vector v,
                                      const GrB_Descriptor desc);
GrB_Info GrB_Vector_eWiseMult_BinaryOp(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum,
                                       const GrB_BinaryOp op, const GrB_Vector u, const GrB_Vector v,
                                       const GrB_Descriptor desc);
GrB_Info GrB_Vector_apply(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_UnaryOp op,
                          const GrB_Vector u, const GrB_Descriptor desc);
/* reduce to a C scalar with a monoid; *nvals_out (optional) receives u's nvals so the caller can tell
 * "empty" (GrB_Vector_reduce_T leaves *val untouched then).  val is accum'ed when accum != NULL. */
GrB_Info GrB_cuda_Vector_reduce(void *val, GrB_Type val_type, const GrB_BinaryOp accum, const GrB_Monoid op,
                                const GrB_Vector u, GrB_Index *nvals_out);
/* w<mask>(all) = accum(w, scalar): GrB_Vector_assign_T with GrB_ALL */
GrB_Info GrB_cuda_Vector_assign_scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum,
                                       const void *val, GrB_Type val_type, const GrB_Descriptor desc);

/* ------------------------------------------------------------------ data in / out (type-generic cores;
 * the typed C-API names GrB_Matrix_import_FP32 ... are exported too and forward to these)
 * import/export: graphblas/core/matrix.py:992-1068, :1601-1645 ; build: :627-681 ;
 * extractTuples: :525-594 ; Vector build/extractTuples: core/vector.py:465-568 */
GrB_Info GrB_cuda_Matrix_import(GrB_Matrix *A, GrB_Type type, GrB_Type xtype /* type of Ax; NULL = type */,
                                GrB_Index nrows, GrB_Index ncols,
                                const GrB_Index *Ap, const GrB_Index *Ai, const void *Ax, GrB_Index Ap_len,
                                GrB_Index Ai_len, GrB_Index Ax_len, GrB_Format format);
GrB_Info GrB_Matrix_exportSize(GrB_Index *Ap_len, GrB_Index *Ai_len, GrB_Index *Ax_len, GrB_Format format,
                               GrB_Matrix A);
GrB_Info GrB_cuda_Matrix_export(GrB_Index *Ap, GrB_Index *Ai, void *Ax, GrB_Type type, GrB_Index *Ap_len,
                                GrB_Index *Ai_len, GrB_Index *Ax_len, GrB_Format format, GrB_Matrix A);
GrB_Info GrB_cuda_Matrix_build(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, const void *X,
                               GrB_Type xtype, GrB_Index nvals, const GrB_BinaryOp dup);
GrB_Info GrB_cuda_Matrix_extractTuples(GrB_Index *I, GrB_Index *J, void *X, GrB_Type xtype, GrB_Index *nvals,
                                       const GrB_Matrix A);
GrB_Info GrB_cuda_Matrix_extractElement(void *x, GrB_Type xtype, const GrB_Matrix A, GrB_Index i, GrB_Index j);
GrB_Info GrB_cuda_Vector_build(GrB_Vector w, const GrB_Index *I, const void *X, GrB_Type xtype, GrB_Index nvals,
                               const GrB_BinaryOp dup);
GrB_Info GrB_cuda_Vector_extractTuples(GrB_Index *I, void *X, GrB_Type xtype, GrB_Index *nvals,
                                       const GrB_Vector v);
GrB_Info GrB_cuda_Vector_setElement(GrB_Vector w, const void *x, GrB_Type xtype, GrB_Index i);
GrB_Info GrB_cuda_Vector_extractElement(void *x, GrB_Type xtype, const GrB_Vector v, GrB_Index i);
GrB_Info GrB_Vector_removeElement(GrB_Vector w, GrB_Index i);

/* ------------------------------------------------------------------ typed C-API names (one set per builtin type)
 * These are the exact symbols the reference formats and looks up, e.g. f"GrB_Matrix_import_{dtype.name}"
 * (graphblas/core/matrix.py:1040-1060), f"GrB_Matrix_export_{dtype_name}" (:1618-1621),
 * f"GrB_Matrix_build_{dtype_name}" (:660-676), f"GrB_Vector_extractTuples_{dtype_name}" (core/vector.py:500-510). */
#define GRB_CUDA_DECLARE_TYPED(SFX, CT)                                                                              \
    GrB_Info GrB_Matrix_import_##SFX(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols,               \
                                     const GrB_Index *Ap, const GrB_Index *Ai, const CT *Ax, GrB_Index Ap_len,      \
                                     GrB_Index Ai_len, GrB_Index Ax_len, GrB_Format format);                        \
    GrB_Info GrB_Matrix_export_##SFX(GrB_Index *Ap, GrB_Index *Ai, CT *Ax, GrB_Index *Ap_len, GrB_Index *Ai_len,    \
                                     GrB_Index *Ax_len, GrB_Format format, GrB_Matrix A);                           \
    GrB_Info GrB_Matrix_build_##SFX(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, const CT *X,              \
                                    GrB_Index nvals, const GrB_BinaryOp dup);                                       \
    GrB_Info GrB_Matrix_extractTuples_##SFX(GrB_Index *I, GrB_Index *J, CT *X, GrB_Index *nvals, const GrB_Matrix A); \
    GrB_Info GrB_Matrix_extractElement_##SFX(CT *x, const GrB_Matrix A, GrB_Index i, GrB_Index j);                  \
    GrB_Info GrB_Vector_build_##SFX(GrB_Vector w, const GrB_Index *I, const CT *X, GrB_Index nvals,                 \
                                    const GrB_BinaryOp dup);                                                        \
    GrB_Info GrB_Vector_extractTuples_##SFX(GrB_Index *I, CT *X, GrB_Index *nvals, const GrB_Vector v);             \
    GrB_Info GrB_Vector_setElement_##SFX(GrB_Vector w, CT x, GrB_Index i);                                          \
    GrB_Info GrB_Vector_extractElement_##SFX(CT *x, const GrB_Vector v, GrB_Index i);                               \
    GrB_Info GrB_Vector_reduce_##SFX(CT *val, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Vector u,    \
                                     const GrB_Descriptor desc);                                                    \
    GrB_Info GrB_Vector_assign_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, CT val,         \
                                     const GrB_Index *indices, GrB_Index ni, const GrB_Descriptor desc);
GRB_CUDA_DECLARE_TYPED(BOOL, bool)
GRB_CUDA_DECLARE_TYPED(INT8, int8_t)
GRB_CUDA_DECLARE_TYPED(INT16, int16_t)
GRB_CUDA_DECLARE_TYPED(INT32, int32_t)
GRB_CUDA_DECLARE_TYPED(INT64, int64_t)
GRB_CUDA_DECLARE_TYPED(UINT8, uint8_t)
GRB_CUDA_DECLARE_TYPED(UINT16, uint16_t)
GRB_CUDA_DECLARE_TYPED(UINT32, uint32_t)
GRB_CUDA_DECLARE_TYPED(UINT64, uint64_t)
GRB_CUDA_DECLARE_TYPED(FP32, float)
GRB_CUDA_DECLARE_TYPED(FP64, double)
extern const GrB_Index *GrB_ALL;

/* ------------------------------------------------------------------ GrB_Scalar objects
 * graphblas/core/scalar.py:83 (GrB_Scalar_new), :235-246 (nvals), :262-280 (clear / setElement_<T>), :199-216
 * (extractElement_<T>); SURVEY.md section 8b lists them among the required lifecycle exports.  A scalar holds one value of
 * a builtin type, or nothing (extractElement then returns GrB_NO_VALUE). */
GrB_Info GrB_Scalar_new(GrB_Scalar *s, GrB_Type type);
GrB_Info GrB_Scalar_free(GrB_Scalar *s);
GrB_Info GrB_Scalar_dup(GrB_Scalar *t, const GrB_Scalar s);
GrB_Info GrB_Scalar_clear(GrB_Scalar s);
GrB_Info GrB_Scalar_nvals(GrB_Index *nvals, const GrB_Scalar s);
GrB_Info GrB_Scalar_wait(GrB_Scalar s, GrB_WaitMode mode);
GrB_Info GrB_Scalar_error(const char **error, const GrB_Scalar s);
/* reduce to a GrB_Scalar: what the reference's default v.reduce() / A.reduce_scalar() call
 * (graphblas/core/vector.py:1670, core/matrix.py:2750); an empty input leaves an empty scalar */
GrB_Info GrB_Vector_reduce_Monoid_Scalar(GrB_Scalar s, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Vector u,
                                         const GrB_Descriptor desc);
GrB_Info GrB_Matrix_reduce_Monoid_Scalar(GrB_Scalar s, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Matrix A,
                                         const GrB_Descriptor desc);
/* bind-1st / bind-2nd apply with a GrB_Scalar (graphblas/core/vector.py:1479, 1525; core/matrix.py:2474, 2520) */
GrB_Info GrB_Vector_apply_BinaryOp1st_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                             const GrB_Scalar x, const GrB_Vector u, const GrB_Descriptor desc);
GrB_Info GrB_Vector_apply_BinaryOp2nd_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                             const GrB_Vector u, const GrB_Scalar y, const GrB_Descriptor desc);
GrB_Info GrB_Matrix_apply_BinaryOp1st_Scalar(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                             const GrB_Scalar x, const GrB_Matrix A, const GrB_Descriptor desc);
GrB_Info GrB_Matrix_apply_BinaryOp2nd_Scalar(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                             const GrB_Matrix A, const GrB_Scalar y, const GrB_Descriptor desc);
/* select with a builtin GrB_IndexUnaryOp (GrB_TRIL, GrB_TRIU, GrB_DIAG, GrB_OFFDIAG, GrB_COLLE, GrB_COLGT, GrB_ROWLE, GrB_ROWGT,
 * GrB_VALUE{EQ,NE,GT,GE,LT,LE}_<T>): graphblas/core/vector.py:1622-1624, core/matrix.py:2621-2623 */
GrB_Info GrB_Vector_select_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                  const GrB_Vector u, const GrB_Scalar y, const GrB_Descriptor desc);
GrB_Info GrB_Matrix_select_Scalar(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                  const GrB_Matrix A, const GrB_Scalar y, const GrB_Descriptor desc);
/* type-generic cores of the typed select names below: thunk passed by pointer + type */
GrB_Info GrB_cuda_Vector_select(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                const GrB_Vector u, const void *thunk, GrB_Type thunk_type, const GrB_Descriptor desc);
GrB_Info GrB_cuda_Matrix_select(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                const GrB_Matrix A, const void *thunk, GrB_Type thunk_type, const GrB_Descriptor desc);
/* w<mask> accum= u and C<Mask> accum= A over all indices (GrB_ALL): graphblas/core/vector.py:1928, core/matrix.py:3300 */
GrB_Info GrB_Vector_assign(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Vector u,
                           const GrB_Index *indices, GrB_Index ni, const GrB_Descriptor desc);
GrB_Info GrB_Vector_assign_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Scalar s,
                                  const GrB_Index *indices, GrB_Index ni, const GrB_Descriptor desc);
GrB_Info GrB_Matrix_assign(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_Matrix A,
                           const GrB_Index *rows, GrB_Index nrows, const GrB_Index *cols, GrB_Index ncols, const GrB_Descriptor desc);
/* the typed names of the above, one set per builtin type: f"GrB_Scalar_setElement_{T}" (core/scalar.py:272),
 * f"GrB_Vector_apply_BinaryOp1st_{T}" (core/vector.py:1477), f"GrB_Matrix_reduce_{T}" (core/matrix.py:2754),
 * f"GrB_Vector_select_{T}" (core/vector.py:1622) ... */
#define GRB_CUDA_DECLARE_TYPED2(SFX, CT)                                                                                     \
    GrB_Info GrB_Scalar_setElement_##SFX(GrB_Scalar s, CT x);                                                                \
    GrB_Info GrB_Scalar_extractElement_##SFX(CT *x, const GrB_Scalar s);                                                     \
    GrB_Info GrB_Matrix_setElement_##SFX(GrB_Matrix C, CT x, GrB_Index i, GrB_Index j);                                      \
    GrB_Info GrB_Matrix_reduce_##SFX(CT *val, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Matrix A,              \
                                     const GrB_Descriptor desc);                                                             \
    GrB_Info GrB_Vector_apply_BinaryOp1st_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum,               \
                                                const GrB_BinaryOp op, CT x, const GrB_Vector u, const GrB_Descriptor desc);  \
    GrB_Info GrB_Vector_apply_BinaryOp2nd_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum,               \
                                                const GrB_BinaryOp op, const GrB_Vector u, CT y, const GrB_Descriptor desc);  \
    GrB_Info GrB_Matrix_apply_BinaryOp1st_##SFX(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum,               \
                                                const GrB_BinaryOp op, CT x, const GrB_Matrix A, const GrB_Descriptor desc);  \
    GrB_Info GrB_Matrix_apply_BinaryOp2nd_##SFX(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum,               \
                                                const GrB_BinaryOp op, const GrB_Matrix A, CT y, const GrB_Descriptor desc);  \
    GrB_Info GrB_Vector_select_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op, \
                                     const GrB_Vector u, CT y, const GrB_Descriptor desc);                                   \
    GrB_Info GrB_Matrix_select_##SFX(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op, \
                                     const GrB_Matrix A, CT y, const GrB_Descriptor desc);
GRB_CUDA_DECLARE_TYPED2(BOOL, bool)
GRB_CUDA_DECLARE_TYPED2(INT8, int8_t)
GRB_CUDA_DECLARE_TYPED2(INT16, int16_t)
GRB_CUDA_DECLARE_TYPED2(INT32, int32_t)
GRB_CUDA_DECLARE_TYPED2(INT64, int64_t)
GRB_CUDA_DECLARE_TYPED2(UINT8, uint8_t)
GRB_CUDA_DECLARE_TYPED2(UINT16, uint16_t)
GRB_CUDA_DECLARE_TYPED2(UINT32, uint32_t)
GRB_CUDA_DECLARE_TYPED2(UINT64, uint64_t)
GRB_CUDA_DECLARE_TYPED2(FP32, float)
GRB_CUDA_DECLARE_TYPED2(FP64, double)
/* w<mask> accum= op(scalar, u) (scalar_first != 0) or op(u, scalar): GrB_Vector_apply_BinaryOp1st/2nd_<T> */
GrB_Info GrB_cuda_Vector_apply_binop(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                     const GrB_Vector u, const void *scalar, GrB_Type scalar_type, int scalar_first,
                                     const GrB_Descriptor desc);
size_t GrB_cuda_kernel_names(char *buf, size_t buflen);

/* ------------------------------------------------------------------ device-side extensions (the GxB_ analogue;
 * precedent: graphblas/core/ss/descriptor.py:77-83 axb_method, ss/_core.py:129-138 gpu_id) */
/* fast import with the device's own index widths (int64 row pointers, int32 column indices); `on_device`
 * != 0 means the three pointers are device pointers (copied, not adopted) */
GrB_Info GrB_cuda_Matrix_import_csr32(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols,
                                      const int64_t *Ap, const int32_t *Aj, const void *Ax, GrB_Index nvals,
                                      int on_device, int sorted);
GrB_Info GrB_cuda_Matrix_export_csr32(int64_t *Ap, int32_t *Aj, void *Ax, GrB_Index nvals_capacity,
                                      GrB_Matrix A, int sort);
/* the same copies on a separate copy stream, ordered after the work enqueued so far: the D2H of one result overlaps the
 * computation of the next.  Host arrays (pinned) and the matrix must stay alive until GrB_cuda_copy_sync() returns. */
/* fence on the copy stream: the ticket names everything enqueued so far; copy_wait(ticket) returns when that has drained while later
   copies keep running (release the source of block k while block k + 1 is still leaving the device); at most 16 tickets outstanding */
GrB_Info GrB_cuda_copy_fence(int *ticket);
GrB_Info GrB_cuda_copy_wait(int ticket);
GrB_Info GrB_cuda_Matrix_export_csr32_async(int64_t *Ap, int32_t *Aj, void *Ax, GrB_Index nvals_capacity, GrB_Matrix A);
GrB_Info GrB_cuda_copy_sync(void);
/* raw device views (valid until the object is next modified) */
GrB_Info GrB_cuda_Matrix_device_csr(const GrB_Matrix A, int64_t **Ap, int32_t **Aj, void **Ax);
/* matrix element-wise operations, all on the device: the C-API-2.0 names the reference calls
   (graphblas/core/base.py:401-411 transpose; core/matrix.py:1972-2165 eWise; :2440-2533 apply) plus bind-1st/2nd and
   reduce-to-scalar with the scalar passed by pointer + type (same convention as the GrB_cuda_Vector_* variants) */
GrB_Info GrB_transpose(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_Matrix A, const GrB_Descriptor desc);
GrB_Info GrB_Matrix_apply(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_UnaryOp op, const GrB_Matrix A,
                          const GrB_Descriptor desc);
GrB_Info GrB_cuda_Matrix_apply_binop(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                     const GrB_Matrix A, const void *scalar, GrB_Type scalar_type, int scalar_first,
                                     const GrB_Descriptor desc);
GrB_Info GrB_Matrix_eWiseAdd_BinaryOp(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                      const GrB_Matrix A, const GrB_Matrix B, const GrB_Descriptor desc);
GrB_Info GrB_Matrix_eWiseMult_BinaryOp(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                       const GrB_Matrix A, const GrB_Matrix B, const GrB_Descriptor desc);
GrB_Info GrB_cuda_Matrix_reduce(void *val, GrB_Type val_type, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Matrix A,
                                GrB_Index *nvals);
/* v as an n x 1 matrix (column vector), built on the device: what Vector._as_matrix provides for Vector.inner / Vector.outer
   (reference graphblas/core/vector.py:193-209, 1715-1787) */
GrB_Info GrB_cuda_Matrix_from_Vector(GrB_Matrix *A, const GrB_Vector v);
/* C-API 2.0: C = square matrix of order size(v) + |k| with v on its k-th diagonal (reference graphblas/core/vector.py:627) */
GrB_Info GrB_Matrix_diag(GrB_Matrix *C, const GrB_Vector v, int64_t k);
GrB_Info GrB_cuda_Vector_device_arrays(const GrB_Vector v, void **vals, uint8_t **present);
GrB_Info GrB_cuda_Vector_import_dense(GrB_Vector *v, GrB_Type type, GrB_Index n, const void *vals,
                                      const uint8_t *present /* NULL = full */, int on_device);
GrB_Info GrB_cuda_Vector_export_dense(void *vals, uint8_t *present, const GrB_Vector v);
GrB_Info GrB_cuda_Vector_touch(GrB_Vector v); /* arrays were modified through device_arrays(): drop caches */
GrB_Info GrB_cuda_Vector_assume_full(GrB_Vector v); /* ... and the caller vouches that every position now holds an entry (no recount) */
GrB_Info GrB_cuda_Matrix_sort(GrB_Matrix A);          /* finish a lazily "jumbled" result now */
/* squeeze a row-end product (rows in order, unused slots between them: how GrB_mxm leaves a result that hardly compresses)
   into the compact CSR now; a no-op on a compact matrix.  GrB_Matrix_wait(A, GrB_MATERIALIZE) does this and the sort. */
GrB_Info GrB_cuda_Matrix_compact(GrB_Matrix A);
GrB_Info GrB_cuda_Matrix_build_transpose(GrB_Matrix A); /* prebuild + cache the CSR of A' */
/* mxm symbolic phase only: *flops, *nvals_out of A(+).(x)B without forming it */
GrB_Info GrB_cuda_mxm_symbolic(GrB_Index *flops, GrB_Index *nvals_out, const GrB_Matrix A, const GrB_Matrix B,
                               const GrB_Descriptor desc);

/* ------------------------------------------------------------------ fused multiply + exchange over peer memory
 * (the per-iteration all-gather of the row-partitioned BFS / SSSP / PageRank, done by the SpMV
 * epilogue itself with NVLink P2P stores instead of a separate NCCL collective).
 * peer_alloc: cudaMalloc'd, zeroed, exportable with CUDA IPC; ipc_get / ipc_open move the mapping between the ranks of
 * one node; Vector_wrap presents such a buffer (values + presence bytes) as a library vector;
 * set_peer_targets(n, vals[], present[] (or NULL), offset, scale): from now on GrB_mxv / GrB_vxm also store every finished
 * output position i at position offset + i of the n target vectors (values optionally multiplied by scale[i], a device
 * array of the result type); n = 0 switches it off.  The targets must outlive the multiply; ordering between ranks
 * (nobody reads a target before every writer's kernel has completed) is the caller's, e.g. the all-reduce of the
 * convergence flag. */
GrB_Info GrB_cuda_peer_alloc(void **ptr, size_t bytes);
GrB_Info GrB_cuda_peer_free(void *ptr);
GrB_Info GrB_cuda_ipc_get(void *ptr, unsigned char handle[64]);
GrB_Info GrB_cuda_ipc_open(const unsigned char handle[64], void **ptr);
GrB_Info GrB_cuda_ipc_close(void *ptr);
GrB_Info GrB_cuda_Vector_wrap(GrB_Vector *v, GrB_Type type, GrB_Index n, void *vals, uint8_t *present);
GrB_Info GrB_cuda_set_peer_targets(int n, void *const *vals, uint8_t *const *present, GrB_Index offset, const void *scale);

/* streams / timing / options / introspection */
GrB_Info GrB_cuda_set_stream(void *cuda_stream); /* NULL restores the library's own stream */
void *GrB_cuda_get_stream(void);
/* order the library stream against another CUDA stream without a host sync: direction 0 = library waits for `other`,
   1 = `other` waits for the library; NULL / (void*)1 = the legacy NULL stream */
GrB_Info GrB_cuda_stream_order(void *other_stream, int direction);
GrB_Info GrB_cuda_sync(void);
GrB_Info GrB_cuda_set_device(int device);
GrB_Info GrB_cuda_timer_start(void);          /* cudaEventRecord on the library stream */
GrB_Info GrB_cuda_timer_stop(float *ms);      /* record + synchronize + elapsed */
GrB_Info GrB_cuda_set_option(const char *key, const char *value);
const char *GrB_cuda_get_option(const char *key);
uint64_t GrB_cuda_launch_count(void);         /* kernels launched by this library so far */
GrB_Info GrB_cuda_kernel_time(const char *name, double *total_ms, uint64_t *launches); /* when option profile=1 */
size_t GrB_cuda_memory_in_use(void);
const char *GrB_cuda_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GRB_CUDA_H */
