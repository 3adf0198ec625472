// objects.cu -- GrB_Matrix / GrB_Vector lifecycle and storage helpers
#include <cub/cub.cuh>

#include "grb_internal.h"

// ------------------------------------------------------------------ util
GrB_Info exclusive_scan_i64(int64_t *data, int64_t n, std::string *err) {
    if (n <= 0) return GrB_SUCCESS;
    size_t tmp_bytes = 0;
    CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, data, data, n, g_stream));
    void *tmp = dev_alloc(tmp_bytes);
    if (!tmp) return set_error(err, GrB_OUT_OF_MEMORY, "scan scratch");
    note_launch("cub_exclusive_sum");
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, data, data, n, g_stream);
    dev_free(tmp);
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

GrB_Info fill_bytes(void *p, int value, size_t bytes) {
    if (bytes == 0) return GrB_SUCCESS;
    CUDA_TRY(nullptr, cudaMemsetAsync(p, value, bytes, g_stream));
    return GrB_SUCCESS;
}

int64_t read_i64(const int64_t *dptr) {
    int64_t v = 0;
    cudaMemcpyAsync(&v, dptr, sizeof v, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    return v;
}

// ------------------------------------------------------------------ matrix storage
void csr_drop_hot(CsrArrays &c) {
    dev_free(c.hot_remap);
    dev_free(c.hot_cols);
    free(c.hot_prefix);
    c.hot_remap = nullptr;
    c.hot_cols = nullptr;
    c.hot_prefix = nullptr;
    c.hot_n = 0;
    c.hot_state = 0;
    c.pull_calls = 0;
    for (int q = 0; q < 2; q++) {
        c.pull_choice[q] = 0;
        c.pull_stage[q] = 0;
        for (int k = 0; k < 4; k++) c.pull_ms[q][k] = -1.f;
    }
}

void csr_drop_seg(CsrArrays &c) {
    dev_free(c.seg_flags);
    dev_free(c.seg_rows);
    dev_free(c.seg_tile_ord);
    c.seg_flags = nullptr;
    c.seg_rows = nullptr;
    c.seg_tile_ord = nullptr;
    c.seg_nonempty = 0;
    c.seg_state = 0;
}

void csr_drop_band(CsrArrays &c);   // spmv_band.cu
void csr_free(CsrArrays &c) {
    csr_drop_band(c);
    dev_free(c.ptr);
    dev_free(c.idx);
    dev_free(c.val);
    dev_free(c.end);
    dev_free(c.canon);
    dev_free(c.tile_starts);
    csr_drop_hot(c);
    csr_drop_seg(c);
    c = CsrArrays();
}

void matrix_drop_twin(GrB_Matrix A) {
    if (A->has_twin) csr_free(A->twin);
    A->has_twin = false;
}

void matrix_release(GrB_Matrix A) {
    csr_free(A->csr);
    matrix_drop_twin(A);
    A->nvals = 0;
    A->jumbled = false;
}

GrB_Info matrix_new_shell(GrB_Matrix *A, int type, int64_t nrows, int64_t ncols) {
    GrB_Matrix M = new (std::nothrow) GrB_Matrix_opaque();
    if (!M) return GrB_OUT_OF_MEMORY;
    M->magic = GRB_MAGIC_MATRIX;
    M->type = type;
    M->nrows = nrows;
    M->ncols = ncols;
    M->nvals = 0;
    M->jumbled = false;
    M->has_twin = false;
    *A = M;
    return GrB_SUCCESS;
}

GrB_Info matrix_alloc_csr(GrB_Matrix A, int64_t nvals) {
    csr_free(A->csr);
    matrix_drop_twin(A);
    A->csr.ptr = dev_alloc_t<int64_t>((size_t)A->nrows + 1);
    A->csr.idx = dev_alloc_t<int32_t>((size_t)(nvals > 0 ? nvals : 1));
    A->csr.val = dev_alloc((size_t)(nvals > 0 ? nvals : 1) * type_size(A->type));
    if (!A->csr.ptr || !A->csr.idx || !A->csr.val) {
        csr_free(A->csr);
        return set_error(&A->err, GrB_OUT_OF_MEMORY, "cannot allocate CSR for %lld entries", (long long)nvals);
    }
    A->nvals = nvals;
    return GrB_SUCCESS;
}

// an empty matrix still needs a valid (all-zero) row pointer array for kernels
GrB_Info matrix_ensure_ptr(GrB_Matrix A) {
    if (A->csr.ptr) return GrB_SUCCESS;
    GRB_TRY(matrix_alloc_csr(A, 0));
    return fill_bytes(A->csr.ptr, 0, sizeof(int64_t) * ((size_t)A->nrows + 1));
}
GrB_Info matrix_materialize(GrB_Matrix A) {
    GRB_TRY(matrix_ensure_ptr(A));
    if (A->csr.end) GRB_TRY(csr_compact(A));
    return GrB_SUCCESS;
}

void matrix_take(GrB_Matrix dst, GrB_Matrix src) {
    csr_free(dst->csr);
    matrix_drop_twin(dst);
    dst->csr = src->csr;
    dst->twin = src->twin;
    dst->has_twin = src->has_twin;
    dst->nvals = src->nvals;
    dst->jumbled = src->jumbled;
    src->csr = CsrArrays();
    src->twin = CsrArrays();
    src->has_twin = false;
    src->nvals = 0;
    src->jumbled = false;
}

extern "C" GrB_Info GrB_Matrix_new(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols) {
    CHECK_INIT();
    if (!A || !type) return set_error(nullptr, GrB_NULL_POINTER, "GrB_Matrix_new: null argument");
    if (nrows > ((GrB_Index)1 << 60) || ncols > ((GrB_Index)1 << 60))
        return set_error(nullptr, GrB_INVALID_VALUE, "GrB_Matrix_new: dimension exceeds GrB_INDEX_MAX");
    if (ncols > (GrB_Index)INT32_MAX || nrows > (GrB_Index)INT32_MAX)
        return set_error(nullptr, GrB_NOT_IMPLEMENTED,
                         "GrB_Matrix_new: this backend stores 32-bit indices; dimensions must be < 2^31");
    return matrix_new_shell(A, type->code, (int64_t)nrows, (int64_t)ncols);
}

extern "C" GrB_Info GrB_Matrix_free(GrB_Matrix *A) {
    if (!A || !*A) return GrB_SUCCESS;
    if (!valid(*A)) return GrB_SUCCESS;
    matrix_release(*A);
    (*A)->magic = GRB_MAGIC_FREED;
    delete *A;
    *A = nullptr;
    return GrB_SUCCESS;
}

static GrB_Info csr_copy(CsrArrays &dst, const CsrArrays &src, int64_t nrows, int64_t nvals, int type, std::string *err) {
    size_t nv = (size_t)(nvals > 0 ? nvals : 1);
    dst.ptr = dev_alloc_t<int64_t>((size_t)nrows + 1);
    dst.idx = dev_alloc_t<int32_t>(nv);
    dst.val = dev_alloc(nv * type_size(type));
    if (!dst.ptr || !dst.idx || !dst.val) {
        csr_free(dst);
        return set_error(err, GrB_OUT_OF_MEMORY, "dup: out of device memory");
    }
    CUDA_TRY(err, cudaMemcpyAsync(dst.ptr, src.ptr, sizeof(int64_t) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, g_stream));
    if (nvals > 0) {
        CUDA_TRY(err, cudaMemcpyAsync(dst.idx, src.idx, sizeof(int32_t) * (size_t)nvals, cudaMemcpyDeviceToDevice, g_stream));
        CUDA_TRY(err, cudaMemcpyAsync(dst.val, src.val, type_size(type) * (size_t)nvals, cudaMemcpyDeviceToDevice, g_stream));
    }
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_Matrix_dup(GrB_Matrix *C, const GrB_Matrix A) {
    CHECK_INIT();
    if (!C) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    GrB_Matrix M;
    GRB_TRY(matrix_new_shell(&M, A->type, A->nrows, A->ncols));
    if (A->csr.ptr) {
        GrB_Info info = matrix_materialize(A);
        if (!info) info = csr_copy(M->csr, A->csr, A->nrows, A->nvals, A->type, &A->err);
        if (info) { delete M; return info; }
        M->nvals = A->nvals;
        M->jumbled = A->jumbled;
    }
    *C = M;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_Matrix_clear(GrB_Matrix A) {
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    matrix_release(A);
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Matrix_nrows(GrB_Index *n, const GrB_Matrix A) {
    if (!n) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    *n = (GrB_Index)A->nrows;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Matrix_ncols(GrB_Index *n, const GrB_Matrix A) {
    if (!n) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    *n = (GrB_Index)A->ncols;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Matrix_nvals(GrB_Index *n, const GrB_Matrix A) {
    if (!n) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    *n = (GrB_Index)A->nvals;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Matrix_wait(GrB_Matrix A, GrB_WaitMode mode) {
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    if (mode == GrB_MATERIALIZE) {   // finish everything the object still owes: compact row-end storage, sorted rows
        GRB_TRY(matrix_materialize(A));
        GRB_TRY(matrix_ensure_sorted(A));
    }
    CUDA_TRY(&A->err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Matrix_error(const char **error, const GrB_Matrix A) {
    if (!error) return GrB_NULL_POINTER;
    if (!valid(A)) { *error = GrB_cuda_last_error(); return GrB_SUCCESS; }
    *error = A->err.c_str();
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ vector storage
GrB_Info vector_new_shell(GrB_Vector *v, int type, int64_t n) {
    GrB_Vector V = new (std::nothrow) GrB_Vector_opaque();
    if (!V) return GrB_OUT_OF_MEMORY;
    V->magic = GRB_MAGIC_VECTOR;
    V->type = type;
    V->n = n;
    V->vals = nullptr;
    V->present = nullptr;
    V->nvals = 0;
    *v = V;
    return GrB_SUCCESS;
}

void vector_release(GrB_Vector v) {
    if (!v->external) {
        dev_free(v->vals);
        dev_free(v->present);
    }
    v->external = false;
    v->vals = nullptr;
    v->present = nullptr;
    v->nvals = 0;
}

GrB_Info vector_ensure_arrays(GrB_Vector v) {
    if (v->vals && v->present) return GrB_SUCCESS;
    if (v->n > (int64_t)INT32_MAX)   // positions are 32-bit on the build / scatter / frontier paths
        return set_error(&v->err, GrB_OUT_OF_MEMORY,
                         "vector of size %lld cannot be held in the dense device layout (at most 2^31 - 1 positions)", (long long)v->n);
    size_t n = (size_t)(v->n > 0 ? v->n : 1);
    v->vals = dev_alloc(n * type_size(v->type));
    v->present = (uint8_t *)dev_alloc(n);
    if (!v->vals || !v->present) {
        vector_release(v);
        return set_error(&v->err, GrB_OUT_OF_MEMORY, "cannot allocate vector of size %lld", (long long)v->n);
    }
    CUDA_TRY(&v->err, cudaMemsetAsync(v->present, 0, n, g_stream));
    CUDA_TRY(&v->err, cudaMemsetAsync(v->vals, 0, n * type_size(v->type), g_stream));
    v->nvals = 0;
    return GrB_SUCCESS;
}

void vector_take_arrays(GrB_Vector v, void *vals, uint8_t *present, int64_t nvals) {
    if (!v->external) {
        if (v->vals != vals) dev_free(v->vals);
        if (v->present != present) dev_free(v->present);
    }
    v->external = false;   // the new arrays are the library's
    v->vals = vals;
    v->present = present;
    v->nvals = nvals;
}

__global__ void count_present_kernel(const uint8_t *__restrict__ p, int64_t n, unsigned long long *out) {
    unsigned long long local = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // 16 bytes per thread per step where aligned
    const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
    int64_t n16 = n / 16;
    for (int64_t k = i; k < n16; k += stride) {
        uint4 w = p4[k];
        local += __popc(w.x & 0x01010101u) + __popc(w.y & 0x01010101u) + __popc(w.z & 0x01010101u) + __popc(w.w & 0x01010101u);
    }
    for (int64_t k = n16 * 16 + i; k < n; k += stride) local += p[k] != 0;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

GrB_Info vector_count(GrB_Vector v) {
    if (v->nvals >= 0) return GrB_SUCCESS;
    if (!v->present) { v->nvals = 0; return GrB_SUCCESS; }
    unsigned long long *d = dev_alloc_t<unsigned long long>(1);
    if (!d) return GrB_OUT_OF_MEMORY;
    CUDA_TRY(&v->err, cudaMemsetAsync(d, 0, 8, g_stream));
    int blocks = (int)std::min<int64_t>((v->n / 16 + 255) / 256 + 1, (int64_t)g_num_sms * 8);
    {
        LAUNCH_NOTE("count_present");
        count_present_kernel<<<blocks, 256, 0, g_stream>>>(v->present, v->n, d);
    }
    v->nvals = read_i64((const int64_t *)d);
    dev_free(d);
    CUDA_TRY(&v->err, cudaGetLastError());
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_Vector_new(GrB_Vector *v, GrB_Type type, GrB_Index n) {
    CHECK_INIT();
    if (!v || !type) return set_error(nullptr, GrB_NULL_POINTER, "GrB_Vector_new: null argument");
    if (n > ((GrB_Index)1 << 60)) return set_error(nullptr, GrB_INVALID_VALUE, "GrB_Vector_new: size exceeds GrB_INDEX_MAX");
    return vector_new_shell(v, type->code, (int64_t)n);
}
extern "C" GrB_Info GrB_Vector_free(GrB_Vector *v) {
    if (!v || !*v) return GrB_SUCCESS;
    if (!valid(*v)) return GrB_SUCCESS;
    vector_release(*v);
    (*v)->magic = GRB_MAGIC_FREED;
    delete *v;
    *v = nullptr;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Vector_dup(GrB_Vector *w, const GrB_Vector u) {
    CHECK_INIT();
    if (!w) return GrB_NULL_POINTER;
    if (!valid(u)) return GrB_UNINITIALIZED_OBJECT;
    GrB_Vector V;
    GRB_TRY(vector_new_shell(&V, u->type, u->n));
    if (u->vals) {
        size_t n = (size_t)(u->n > 0 ? u->n : 1);
        V->vals = dev_alloc(n * type_size(u->type));
        V->present = (uint8_t *)dev_alloc(n);
        if (!V->vals || !V->present) { vector_release(V); delete V; return set_error(&u->err, GrB_OUT_OF_MEMORY, "dup"); }
        CUDA_TRY(&u->err, cudaMemcpyAsync(V->vals, u->vals, n * type_size(u->type), cudaMemcpyDeviceToDevice, g_stream));
        CUDA_TRY(&u->err, cudaMemcpyAsync(V->present, u->present, n, cudaMemcpyDeviceToDevice, g_stream));
        V->nvals = u->nvals;
    }
    *w = V;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Vector_clear(GrB_Vector v) {
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    vector_release(v);
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Vector_size(GrB_Index *n, const GrB_Vector v) {
    if (!n) return GrB_NULL_POINTER;
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    *n = (GrB_Index)v->n;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Vector_nvals(GrB_Index *n, const GrB_Vector v) {
    if (!n) return GrB_NULL_POINTER;
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(vector_count(v));
    *n = (GrB_Index)v->nvals;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Vector_wait(GrB_Vector v, GrB_WaitMode mode) {
    (void)mode;
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    CUDA_TRY(&v->err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Vector_error(const char **error, const GrB_Vector v) {
    if (!error) return GrB_NULL_POINTER;
    if (!valid(v)) { *error = GrB_cuda_last_error(); return GrB_SUCCESS; }
    *error = v->err.c_str();
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_Vector_touch(GrB_Vector v) {
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    v->nvals = -1;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_Vector_assume_full(GrB_Vector v) {
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    v->nvals = v->n;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_Vector_device_arrays(const GrB_Vector v, void **vals, uint8_t **present) {
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(vector_ensure_arrays(v));
    if (vals) *vals = v->vals;
    if (present) *present = v->present;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_Matrix_device_csr(const GrB_Matrix A, int64_t **Ap, int32_t **Aj, void **Ax) {
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(matrix_materialize(A));
    if (Ap) *Ap = A->csr.ptr;
    if (Aj) *Aj = A->csr.idx;
    if (Ax) *Ax = A->csr.val;
    return GrB_SUCCESS;
}

// A vector over caller-owned device arrays (values + presence bytes of n positions): the replicated input vector of the fused
// multiply + exchange lives in an IPC-exportable cudaMalloc buffer the peers store into.  The library reads the arrays in
// place and never frees them; nvals is recounted on demand (GrB_cuda_Vector_touch after remote writes).
extern "C" GrB_Info GrB_cuda_Vector_wrap(GrB_Vector *v, GrB_Type type, GrB_Index n, void *vals, uint8_t *present) {
    CHECK_INIT();
    if (!v || !type || !vals || !present) return GrB_NULL_POINTER;
    if (n > (GrB_Index)INT32_MAX) return GrB_NOT_IMPLEMENTED;
    GrB_Vector w;
    GRB_TRY(vector_new_shell(&w, type->code, (int64_t)n));
    w->vals = vals;
    w->present = present;
    w->external = true;
    w->nvals = -1;
    *v = w;
    return GrB_SUCCESS;
}
