// grb_internal.h -- object model, error plumbing and device-side op algebra of libgrb_cuda.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/grb_cuda.h"

// ------------------------------------------------------------------ codes
enum TypeCode : int {
    TC_BOOL = 0, TC_INT8, TC_INT16, TC_INT32, TC_INT64, TC_UINT8, TC_UINT16, TC_UINT32, TC_UINT64, TC_FP32, TC_FP64,
    TC_COUNT
};
// binary op codes (the first 13 are shared with oracle/grb_oracle.c)
enum OpCode : int {
    OP_NONE = 0, OP_FIRST = 1, OP_SECOND, OP_PAIR, OP_PLUS, OP_MINUS, OP_TIMES, OP_DIV, OP_MIN, OP_MAX, OP_LOR,
    OP_LAND, OP_LXOR, OP_ANY, OP_RMINUS, OP_RDIV, OP_LXNOR, OP_EQ, OP_NE, OP_GT, OP_LT, OP_GE, OP_LE, OP_ISEQ,
    OP_ISNE, OP_POW, OP_RPOW /* pow(y, x): internal, the operand swap of vxm */, OP_COUNT
};
enum UnaryCode : int { UOP_IDENTITY = 1, UOP_AINV, UOP_MINV, UOP_LNOT, UOP_ABS, UOP_ONE, UOP_BNOT,
                       UOP_SQRT, UOP_EXP, UOP_LOG, UOP_EXP2, UOP_LOG2, UOP_LOG10, UOP_FLOOR, UOP_CEIL, UOP_ROUND, UOP_TRUNC, UOP_SIGNUM };   // floating point only
// GrB_IndexUnaryOp codes (select): positional ones take an int64 thunk, VALUE* compare the entry with a thunk of the op's type
enum IndexOpCode : int {
    IOP_TRIL = 1, IOP_TRIU, IOP_DIAG, IOP_OFFDIAG, IOP_COLLE, IOP_COLGT, IOP_ROWLE, IOP_ROWGT, IOP_VALUEEQ, IOP_VALUENE,
    IOP_VALUEGT, IOP_VALUEGE, IOP_VALUELT, IOP_VALUELE, IOP_ROWINDEX, IOP_COLINDEX, IOP_DIAGINDEX
};

static inline size_t type_size(int tc) {
    static const size_t s[TC_COUNT] = {1, 1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
    return s[tc];
}

// ------------------------------------------------------------------ opaque objects
struct GrB_Type_opaque { int code; size_t size; const char *name; };
struct GrB_UnaryOp_opaque { int opcode; int type; const char *name; };
struct GrB_BinaryOp_opaque { int opcode; int type; int ztype; const char *name; };
struct GrB_Monoid_opaque { int opcode; int type; const char *name; };
struct GrB_Semiring_opaque { int add; int mul; int type; const char *name; };
struct GrB_IndexUnaryOp_opaque { int opcode; int type; const char *name; };   // type: TC_* of VALUE* ops, -1 for positional ones
struct GrB_Descriptor_opaque { bool replace, comp, structure, t0, t1; const char *name; };

#define GRB_MAGIC_MATRIX 0x4d61747269784742ull
#define GRB_MAGIC_VECTOR 0x566563746f724742ull
#define GRB_MAGIC_SCALAR 0x5363616c61724742ull
#define GRB_MAGIC_FREED 0x4672656564474221ull

struct CsrArrays {
    int64_t *ptr = nullptr;   // nrows+1
    int32_t *idx = nullptr;   // nvals
    void *val = nullptr;      // nvals * type_size
    // "row-end" form (an mxm result before anything needs it compact): row i occupies [ptr[i], end[i]) of idx / val, which hold
    // `cap` slots; ptr is monotone (rows stay in row order) but ptr[i + 1] - end[i] slots may be unused.  This is the 4-array CSR of
    // the sparse BLAS (pointerB / pointerE); the multiply kernels read it as it is (they take a row-end pointer: ptr + 1 for a
    // compact CSR), everything else first calls matrix_materialize(), which squeezes the gaps out once.  nullptr: compact.
    int64_t *end = nullptr;
    int64_t *canon = nullptr; // nrows+1 compact row pointers of a row-end CSR (kept from the count scan, so compaction is one kernel)
    int64_t cap = 0;
    int64_t *tile_starts = nullptr;  // merge-path tile row coordinates (cached; see spmv.cu)
    int64_t n_tiles = 0;
    int tile_items = 0;
    // hot-column cache for pull SpMV (spmv.cu / hotcols.cu): the most referenced columns of this CSR, by rank
    int32_t *hot_remap = nullptr;    // nvals: column index, or HOT_FLAG | rank for a hot column
    int32_t *hot_cols = nullptr;     // hot_n: rank -> column
    int64_t *hot_prefix = nullptr;   // HOST array (malloc): references covered by ranks 0..r
    int hot_n = 0;
    int hot_state = 0;               // 0 not analysed, 1 built, -1 not worthwhile
    int pull_calls = 0;              // pull SpMV calls seen (the analysis is paid for on the second one)
    // auto mode of the pull SpMV (spmv.cu run_pull): timed trial of merge-path / segmented / segmented + hot-column cache on the
    // first multiplies with this CSR, per element-size class (<= 4 bytes, 8 bytes); the winner is kept
    int pull_choice[2] = {0, 0};     // 0 undecided, 1 merge, 2 seg, 3 seg + hot columns, 4 column-banded
    int pull_stage[2] = {0, 0};      // next candidate to time
    float pull_ms[2][4] = {{-1.f, -1.f, -1.f, -1.f}, {-1.f, -1.f, -1.f, -1.f}};
    // row-boundary metadata of the segmented pull SpMV (spmv_seg.cu), built on first use
    uint8_t *seg_flags = nullptr;    // bit (k & 7) of byte (k >> 3): entry k is the first of its row
    int32_t *seg_rows = nullptr;     // rows that have at least one entry, ascending
    int32_t *seg_tile_ord = nullptr; // per 128 entries: ordinal (in seg_rows) of the row in progress before the tile
    int64_t seg_nonempty = 0;
    int seg_state = 0;               // 0 not built, 1 built
    // column-banded copy of this CSR for the banded pull SpMV (spmv_band.cu), per element-size class; built on first use
    void *band_fmt[2] = {nullptr, nullptr};
};

struct GrB_Matrix_opaque {
    uint64_t magic;
    int type;
    int64_t nrows, ncols, nvals;
    CsrArrays csr;        // canonical storage: CSR, int64 row pointers, int32 column indices
    bool jumbled;         // column indices within a row may be unsorted (lazy sort, like the reference's C library)
    CsrArrays twin;       // cached CSR of the transpose (== CSC of this matrix); valid iff has_twin
    bool has_twin;
    std::string err;
};

struct GrB_Vector_opaque {
    uint64_t magic;
    int type;
    int64_t n;
    void *vals;           // n * type_size   (nullptr until first data)
    uint8_t *present;     // n bytes, 1 = entry exists
    int64_t nvals;        // -1 = unknown (count lazily)
    bool external = false;   // vals / present belong to the caller (GrB_cuda_Vector_wrap): never freed by the library
    std::string err;
};

// a GrB_Scalar lives on the host: one value of a builtin type, or empty
struct GrB_Scalar_opaque {
    uint64_t magic;
    int type;
    bool has;
    unsigned char buf[8];
    std::string err;
};
static inline bool valid(const GrB_Scalar s) { return s && s->magic == GRB_MAGIC_SCALAR; }
// host-side cast between builtin types (C semantics; anything -> BOOL is "!= 0")
void host_cast(void *dst, int dst_type, const void *src, int src_type);

// ------------------------------------------------------------------ runtime services (runtime.cu)
extern cudaStream_t g_stream;
extern bool g_initialized;
extern int g_num_sms;
void note_launch(const char *name);
struct KernelTimer { KernelTimer(const char *name); ~KernelTimer(); const char *name; cudaEvent_t a, b; bool on; };
#define LAUNCH_NOTE(name) note_launch(name); KernelTimer _kt(name)
cudaStream_t aux_stream(int i);      // two helper streams for concurrent kernels inside one call (nullptr when unavailable)
bool profiling();                    // per-kernel timing mode: everything stays on g_stream
void phase_mark(const char *name);   // option trace_host=1: host wall time since the previous mark goes to kernel-time slot "host:<name>" (nullptr: restart)

GrB_Info set_error(std::string *slot, GrB_Info info, const char *fmt, ...);
void set_last_error(const char *msg);
GrB_Info cuda_fail(std::string *slot, cudaError_t e, const char *what);
void *dev_alloc(size_t bytes);             // stream-ordered; returns nullptr on failure (error recorded)
void dev_free(void *p);
void *ws_acquire(int slot, size_t bytes);   // cached scratch (see runtime.cu)
void ws_release(int slot, void *p);
void ws_detach(int slot, void *p);          // the caller keeps the block (it leaves the cache; free it with dev_free)
void ws_trim();
template <typename T> static inline T *dev_alloc_t(size_t n) { return (T *)dev_alloc(n * sizeof(T)); }
const char *opt_get(const char *key, const char *dflt);
long opt_get_int(const char *key, long dflt);

#define CUDA_TRY(slot, expr)                                                   \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return cuda_fail(slot, _e, #expr);              \
    } while (0)
#define GRB_TRY(expr)                                  \
    do {                                               \
        GrB_Info _i = (expr);                          \
        if (_i != GrB_SUCCESS) return _i;              \
    } while (0)
#define CHECK_INIT()                                                                     \
    do {                                                                                 \
        if (!g_initialized) { GrB_Info _i = GrB_init(GrB_NONBLOCKING); if (_i) return _i; } \
    } while (0)

static inline bool valid(const GrB_Matrix A) { return A && A->magic == GRB_MAGIC_MATRIX; }
static inline bool valid(const GrB_Vector v) { return v && v->magic == GRB_MAGIC_VECTOR; }

// object helpers (objects.cu)
void csr_free(CsrArrays &c);
void csr_drop_hot(CsrArrays &c);                   // forget the hot-column cache (column order changed)
void csr_drop_seg(CsrArrays &c);
GrB_Info csr_ensure_hot(CsrArrays &c, int64_t ncols, int64_t nvals, int max_hot, std::string *err);   // hotcols.cu
void matrix_drop_twin(GrB_Matrix A);
void matrix_release(GrB_Matrix A);                 // free device arrays, keep shell
GrB_Info matrix_alloc_csr(GrB_Matrix A, int64_t nvals);  // allocates csr.ptr/idx/val for nvals entries
void matrix_take(GrB_Matrix dst, GrB_Matrix src);  // move arrays of src into dst (src becomes empty shell)
GrB_Info matrix_ensure_sorted(GrB_Matrix A);       // sort.cu
GrB_Info matrix_ensure_twin(GrB_Matrix A);         // transpose.cu
GrB_Info matrix_new_shell(GrB_Matrix *A, int type, int64_t nrows, int64_t ncols);
GrB_Info matrix_materialize(GrB_Matrix A);         // make sure csr.ptr exists (all-zero for an empty matrix) and the CSR is compact
GrB_Info matrix_ensure_ptr(GrB_Matrix A);          // csr.ptr exists; a row-end CSR is left as it is (multiply operands)
GrB_Info csr_compact(GrB_Matrix A);                // row-end CSR -> compact CSR (spgemm.cu)
static inline const int64_t *csr_row_end(const CsrArrays &c) { return c.end ? c.end : c.ptr + 1; }
static inline int64_t csr_slots(const CsrArrays &c, int64_t nvals) { return c.end ? c.cap : nvals; }
GrB_Info vector_ensure_arrays(GrB_Vector v);       // allocate vals/present (present zeroed) if missing
GrB_Info vector_count(GrB_Vector v);               // make nvals known
void vector_release(GrB_Vector v);
void vector_take_arrays(GrB_Vector v, void *vals, uint8_t *present, int64_t nvals);
GrB_Info vector_new_shell(GrB_Vector *v, int type, int64_t n);

// casts (convert.cu)
GrB_Info cast_array(void *dst, int dst_type, const void *src, int src_type, int64_t n, std::string *err);
// returns src itself when types agree, else a temp device copy in *tmp (caller dev_free(*tmp))
GrB_Info cast_view(const void **out, void **tmp, const void *src, int src_type, int dst_type, int64_t n,
                   std::string *err);

// generic epilogues (epilogue.cu)
GrB_Info mask_effective_bytes(const uint8_t **out, void **tmp, const uint8_t *present, const void *vals, int type,
                              int64_t n, bool structure, std::string *err);
struct MaskSpec { const uint8_t *present; const void *vals; int type; bool structure, comp; bool has; };
GrB_Info vector_write_back(GrB_Vector w, void *t_vals, uint8_t *t_present, int t_type, const GrB_Vector mask,
                           const GrB_BinaryOp accum, const GrB_Descriptor desc, bool t_owned);
GrB_Info matrix_write_back(GrB_Matrix C, GrB_Matrix T, const GrB_Matrix M, const GrB_BinaryOp accum,
                           const GrB_Descriptor desc);

// multiply cores
// t = M (+).(x) u where M = A (use_transpose=false) or A' (true); flip: multiply is mul(u_k, a) instead of mul(a, u_k).
// mask_eff (nullable): byte per output position, rows/positions ruled out by the mask may be skipped.
// epi (nullable): write-back parameters; when the traversal can finish each output position exactly once (pull
// kernels) it is applied inside the kernel and *fused is set -- the returned arrays are then the FINAL output.
// Peer targets of a fused multiply + exchange (GrB_cuda_set_peer_targets): every finished output position `row` is also stored
// at position offset + row of n remote (or local) vectors, straight from the kernel's epilogue -- the all-gather of the
// row-partitioned iteration without a collective.  `scale` (optional, device array of the result type,
// one value per local row) multiplies the value on its way out (PageRank exchanges damping * t / d, not t).
constexpr int MAX_PEERS = 8;
struct PeerTargets { int n; void *vals[MAX_PEERS]; uint8_t *present[MAX_PEERS]; int64_t offset; const void *scale; };
extern PeerTargets g_peer;
struct VecEpiHost { const void *c_vals; const uint8_t *c_present; const uint8_t *mask; int has_mask, comp, replace, accum; const PeerTargets *peer; };
GrB_Info multiply_mat_vec_impl(void **t_vals, uint8_t **t_present, int64_t *t_len, const GrB_Semiring op, GrB_Matrix A,
                               bool use_transpose, GrB_Vector u, bool flip, const uint8_t *mask_eff, bool mask_comp,
                               std::string *err, const VecEpiHost *epi, bool *fused);
GrB_Info spgemm(GrB_Matrix *T, const GrB_Semiring op, GrB_Matrix A, bool at, GrB_Matrix B, bool bt,
                const GrB_Matrix M, bool mask_comp, bool mask_struct, std::string *err, bool symbolic_only,
                uint64_t *flops_out, uint64_t *nvals_out);

// builtin tables (builtins.cu)
const GrB_Type_opaque *type_of_code(int code);
void builtins_init();

// small device utilities (util.cu)
GrB_Info exclusive_scan_i64(int64_t *data, int64_t n, std::string *err);  // in place; data[n-1] must be a pad slot
GrB_Info fill_bytes(void *p, int value, size_t bytes);
int64_t read_i64(const int64_t *dptr);  // synchronous D2H of one value
