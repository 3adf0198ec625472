// io.cu -- data in / out at the C-ABI boundary.  Host arrays are GrB_Index (uint64) + typed values,
// exactly what the reference passes (graphblas/core/utils.py:58-114, core/matrix.py:992-1068,
// :1601-1645, :627-681, :525-594; core/vector.py:465-568).  On the device indices become
// int32 columns / int64 row pointers; conversion happens once, here.
#include <cub/cub.cuh>
#include <thread>
#include <vector>

#include "grb_ops.cuh"

GrB_Info matrix_check_sorted(GrB_Matrix A, bool *sorted);
GrB_Info expand_row_ids(const GrB_Matrix A, int32_t *rowid);

// ------------------------------------------------------------------ index width conversion kernels
__global__ void u64_to_i64_kernel(int64_t *dst, const uint64_t *src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = (int64_t)src[i];
}
__global__ void u64_to_i32_kernel(int32_t *dst, const uint64_t *src, int64_t n, uint64_t bound, int *err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) {
        uint64_t v = src[i];
        if (v >= bound) *err = 1;
        dst[i] = (int32_t)v;
    }
}
__global__ void i64_to_u64_kernel(uint64_t *dst, const int64_t *src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = (uint64_t)src[i];
}
__global__ void i32_to_u64_kernel(uint64_t *dst, const int32_t *src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = (uint64_t)(uint32_t)src[i];
}
__global__ void check_ptr_kernel(const int64_t *ptr, int64_t nrows, int64_t nnz, int *err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nrows) return;
    if (i == 0 && ptr[0] != 0) *err = 2;
    if (i == nrows && ptr[nrows] != nnz) *err = 2;
    if (i < nrows && ptr[i] > ptr[i + 1]) *err = 2;
}
static inline int grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    return (int)std::min<int64_t>(b, (int64_t)g_num_sms * 16);
}

// upload a host array into a fresh device buffer
static void *upload(const void *host, size_t bytes, std::string *err) {
    void *d = dev_alloc(bytes ? bytes : 16);
    if (!d) { set_error(err, GrB_OUT_OF_MEMORY, "upload buffer (%zu bytes)", bytes); return nullptr; }
    if (bytes && cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, g_stream) != cudaSuccess) {
        cudaGetLastError();
        dev_free(d);
        set_error(err, GrB_PANIC, "host to device copy failed");
        return nullptr;
    }
    return d;
}

// ------------------------------------------------------------------ CSR / CSC import
static GrB_Info import_csr_device(GrB_Matrix A, const uint64_t *dAp, const uint64_t *dAi, const void *dAx, int xtype,
                                  int64_t nnz) {
    std::string *err = &A->err;
    GRB_TRY(matrix_alloc_csr(A, nnz));
    int *flag = dev_alloc_t<int>(1);
    if (!flag) return GrB_OUT_OF_MEMORY;
    cudaMemsetAsync(flag, 0, 4, g_stream);
    note_launch("u64_to_i64");
    u64_to_i64_kernel<<<grid_for(A->nrows + 1), 256, 0, g_stream>>>(A->csr.ptr, dAp, A->nrows + 1);
    note_launch("check_ptr");
    check_ptr_kernel<<<(unsigned)((A->nrows + 1 + 255) / 256), 256, 0, g_stream>>>(A->csr.ptr, A->nrows, nnz, flag);
    if (nnz > 0) {
        note_launch("u64_to_i32");
        u64_to_i32_kernel<<<grid_for(nnz), 256, 0, g_stream>>>(A->csr.idx, dAi, nnz, (uint64_t)A->ncols, flag);
        GrB_Info ci = cast_array(A->csr.val, A->type, dAx, xtype, nnz, err);
        if (ci) { dev_free(flag); return ci; }
    }
    int h = 0;
    cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    dev_free(flag);
    CUDA_TRY(err, cudaGetLastError());
    if (h == 1) return set_error(err, GrB_INDEX_OUT_OF_BOUNDS, "import: an index is out of bounds");
    if (h == 2) return set_error(err, GrB_INVALID_VALUE, "import: malformed pointer array");
    bool sorted = true;
    GRB_TRY(matrix_check_sorted(A, &sorted));
    A->jumbled = !sorted;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Matrix_import(GrB_Matrix *Aout, GrB_Type type, GrB_Type xtype, GrB_Index nrows, GrB_Index ncols,
                                           const GrB_Index *Ap, const GrB_Index *Ai, const void *Ax, GrB_Index Ap_len,
                                           GrB_Index Ai_len, GrB_Index Ax_len, GrB_Format format) {
    CHECK_INIT();
    if (!xtype) xtype = type;
    if (!Aout || !type || !Ap || (!Ai && Ai_len) || (!Ax && Ax_len)) return set_error(nullptr, GrB_NULL_POINTER, "import: null argument");
    if (format == GrB_COO_FORMAT) {
        GrB_Matrix A;
        GRB_TRY(GrB_Matrix_new(&A, type, nrows, ncols));
        if (Ap_len < Ax_len || Ai_len < Ax_len) { GrB_Matrix_free(&A); return set_error(nullptr, GrB_INVALID_VALUE, "import COO: index arrays shorter than values"); }
        GrB_Info info = GrB_cuda_Matrix_build(A, Ap, Ai, Ax, xtype, Ax_len, nullptr);
        if (info) { GrB_Matrix_free(&A); return info; }
        *Aout = A;
        return GrB_SUCCESS;
    }
    const bool csc = format == GrB_CSC_FORMAT;
    const GrB_Index major = csc ? ncols : nrows, minor = csc ? nrows : ncols;
    if (Ap_len < major + 1) return set_error(nullptr, GrB_INVALID_VALUE, "import: pointer array too short (%llu < %llu)", (unsigned long long)Ap_len, (unsigned long long)(major + 1));
    const int64_t nnz = (int64_t)Ap[major];
    if ((GrB_Index)nnz > Ai_len || (GrB_Index)nnz > Ax_len) return set_error(nullptr, GrB_INVALID_VALUE, "import: index/value arrays shorter than Ap[n]=%lld", (long long)nnz);
    GrB_Matrix A;
    GRB_TRY(GrB_Matrix_new(&A, type, major, minor));
    std::string *err = &A->err;
    uint64_t *dAp = (uint64_t *)upload(Ap, sizeof(uint64_t) * (size_t)(major + 1), err);
    uint64_t *dAi = (uint64_t *)upload(Ai, sizeof(uint64_t) * (size_t)nnz, err);
    void *dAx = upload(Ax, type_size(xtype->code) * (size_t)nnz, err);
    GrB_Info info = (!dAp || !dAi || !dAx) ? GrB_OUT_OF_MEMORY : import_csr_device(A, dAp, dAi, dAx, xtype->code, nnz);
    dev_free(dAp); dev_free(dAi); dev_free(dAx);
    if (!info && csc) {
        // the arrays described A' in CSR form: build the twin and swap roles
        info = matrix_ensure_sorted(A);
        if (!info) info = matrix_ensure_twin(A);
        if (!info) {
            std::swap(A->csr, A->twin);
            std::swap(A->nrows, A->ncols);
            A->jumbled = false;
        }
    }
    if (info) {
        set_last_error(A->err.c_str());
        GrB_Matrix_free(&A);
        return info;
    }
    *Aout = A;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Matrix_import_csr32(GrB_Matrix *Aout, GrB_Type type, GrB_Index nrows, GrB_Index ncols,
                                                 const int64_t *Ap, const int32_t *Aj, const void *Ax, GrB_Index nvals,
                                                 int on_device, int sorted) {
    CHECK_INIT();
    if (!Aout || !type || !Ap) return set_error(nullptr, GrB_NULL_POINTER, "import_csr32: null argument");
    GrB_Matrix A;
    GRB_TRY(GrB_Matrix_new(&A, type, nrows, ncols));
    GrB_Info info = matrix_alloc_csr(A, (int64_t)nvals);
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!info) {
        cudaError_t e = cudaMemcpyAsync(A->csr.ptr, Ap, sizeof(int64_t) * (size_t)(nrows + 1), kind, g_stream);
        if (e == cudaSuccess && nvals) e = cudaMemcpyAsync(A->csr.idx, Aj, sizeof(int32_t) * (size_t)nvals, kind, g_stream);
        if (e == cudaSuccess && nvals) e = cudaMemcpyAsync(A->csr.val, Ax, type_size(type->code) * (size_t)nvals, kind, g_stream);
        if (e != cudaSuccess) info = cuda_fail(&A->err, e, "import_csr32 copy");
    }
    if (!info) {
        if (sorted) A->jumbled = false;
        else {
            bool s = true;
            info = matrix_check_sorted(A, &s);
            A->jumbled = !s;
        }
    }
    if (info) { set_last_error(A->err.c_str()); GrB_Matrix_free(&A); return info; }
    *Aout = A;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ export
extern "C" GrB_Info GrB_Matrix_exportSize(GrB_Index *Ap_len, GrB_Index *Ai_len, GrB_Index *Ax_len, GrB_Format format,
                                          GrB_Matrix A) {
    if (!Ap_len || !Ai_len || !Ax_len) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    *Ai_len = *Ax_len = (GrB_Index)A->nvals;
    *Ap_len = format == GrB_CSR_FORMAT ? (GrB_Index)A->nrows + 1 : format == GrB_CSC_FORMAT ? (GrB_Index)A->ncols + 1 : (GrB_Index)A->nvals;
    return GrB_SUCCESS;
}

// ---- large device -> pageable-host downloads (what the reference's to_csr hands us: plain numpy arrays).
// A single cudaMemcpy into pageable memory runs at a few GB/s (the driver stages through one pinned buffer with one host thread)
// and the uint64 widening of int32 indices would double the PCIe bytes.  Instead: the array crosses PCIe in its DEVICE width in
// 128 MB chunks into two pinned buffers, and a handful of host threads widen / copy chunk k into the destination (taking its
// page faults in parallel) while chunk k + 1 is in flight.
static struct { void *buf[2]; size_t bytes; cudaEvent_t ev[2]; } g_pin = {{nullptr, nullptr}, 0, {nullptr, nullptr}};
constexpr size_t PIN_CHUNK_BYTES = (size_t)128 << 20;
static bool pinned_ready() {
    if (g_pin.bytes) return true;
    for (int q = 0; q < 2; q++) {
        if (cudaHostAlloc(&g_pin.buf[q], PIN_CHUNK_BYTES, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&g_pin.ev[q], cudaEventDisableTiming) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
    }
    g_pin.bytes = PIN_CHUNK_BYTES;
    return true;
}
// mode 0: plain copy of `es`-byte elements; 1: int32 -> uint64; 2: int64 -> uint64 (bit copy)
static void host_chunk(void *dst, const void *src, int64_t n, size_t es, int mode) {
    const int nt = (int)std::min<int64_t>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())), (n + (1 << 20) - 1) >> 20);
    auto work = [=](int t) {
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        if (mode == 1) {
            const int32_t *s32 = (const int32_t *)src;
            uint64_t *d = (uint64_t *)dst;
            for (int64_t i = lo; i < hi; i++) d[i] = (uint64_t)(uint32_t)s32[i];
        } else {
            memcpy((char *)dst + (size_t)lo * es, (const char *)src + (size_t)lo * es, (size_t)(hi - lo) * es);
        }
    };
    if (nt <= 1) { work(0); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
}
static GrB_Info download_pipelined(void *host, const void *dev, int64_t n, size_t dev_es, int mode, std::string *err) {
    const size_t host_es = mode == 1 ? 8 : dev_es;
    const int64_t per = (int64_t)(PIN_CHUNK_BYTES / dev_es);
    const int64_t nchunks = (n + per - 1) / per;
    std::thread worker[2];
    cudaError_t e = cudaSuccess;
    for (int64_t k = 0; k <= nchunks && e == cudaSuccess; k++) {
        const int q = (int)(k & 1);
        if (k < nchunks) {
            if (worker[q].joinable()) worker[q].join();   // chunk k - 2 has left this pinned buffer
            const int64_t off = k * per, cnt = std::min<int64_t>(per, n - off);
            e = cudaMemcpyAsync(g_pin.buf[q], (const char *)dev + (size_t)off * dev_es, (size_t)cnt * dev_es, cudaMemcpyDeviceToHost, g_stream);
            if (e == cudaSuccess) e = cudaEventRecord(g_pin.ev[q], g_stream);
        }
        if (k >= 1 && e == cudaSuccess) {   // chunk k - 1 has landed: hand it to the host threads while chunk k is in flight
            const int p = (int)((k - 1) & 1);
            e = cudaEventSynchronize(g_pin.ev[p]);
            const int64_t off = (k - 1) * per, cnt = std::min<int64_t>(per, n - off);
            void *dst = (char *)host + (size_t)off * host_es;
            const void *src = g_pin.buf[p];
            if (e == cudaSuccess) worker[p] = std::thread([=]() { host_chunk(dst, src, cnt, dev_es, mode); });
        }
    }
    for (int q = 0; q < 2; q++)
        if (worker[q].joinable()) worker[q].join();
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

static GrB_Info download_as_u64(GrB_Index *host, const void *dev, bool is32, int64_t n, std::string *err) {
    if (n <= 0) return GrB_SUCCESS;
    if ((size_t)n * (is32 ? 4 : 8) >= PIN_CHUNK_BYTES / 2 && opt_get_int("export_pipelined", 1) != 0 && pinned_ready())
        return download_pipelined(host, dev, n, is32 ? 4 : 8, is32 ? 1 : 2, err);
    uint64_t *tmp = dev_alloc_t<uint64_t>((size_t)n);
    if (!tmp) return set_error(err, GrB_OUT_OF_MEMORY, "export staging");
    note_launch("to_u64");
    if (is32) i32_to_u64_kernel<<<grid_for(n), 256, 0, g_stream>>>(tmp, (const int32_t *)dev, n);
    else i64_to_u64_kernel<<<grid_for(n), 256, 0, g_stream>>>(tmp, (const int64_t *)dev, n);
    cudaError_t e = cudaMemcpyAsync(host, tmp, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, g_stream);
    dev_free(tmp);
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

static GrB_Info download_vals(void *host, int host_type, const void *dev, int dev_type, int64_t n, std::string *err) {
    if (n <= 0) return GrB_SUCCESS;
    const void *src;
    void *tmp;
    GRB_TRY(cast_view(&src, &tmp, dev, dev_type, host_type, n, err));
    GrB_Info info = GrB_SUCCESS;
    if ((size_t)n * type_size(host_type) >= PIN_CHUNK_BYTES / 2 && opt_get_int("export_pipelined", 1) != 0 && pinned_ready()) {
        info = download_pipelined(host, src, n, type_size(host_type), 0, err);
        dev_free(tmp);
        return info;
    }
    cudaError_t e = cudaMemcpyAsync(host, src, type_size(host_type) * (size_t)n, cudaMemcpyDeviceToHost, g_stream);
    dev_free(tmp);
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Matrix_export(GrB_Index *Ap, GrB_Index *Ai, void *Ax, GrB_Type type, GrB_Index *Ap_len,
                                           GrB_Index *Ai_len, GrB_Index *Ax_len, GrB_Format format, GrB_Matrix A) {
    CHECK_INIT();
    if (!Ap_len || !Ai_len || !Ax_len || !type) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    GrB_Index np, ni, nx;
    GRB_TRY(GrB_Matrix_exportSize(&np, &ni, &nx, format, A));
    if (*Ap_len < np || *Ai_len < ni || *Ax_len < nx)
        return set_error(&A->err, GrB_INSUFFICIENT_SPACE, "export: output arrays too small");
    if ((np && !Ap) || (ni && !Ai) || (nx && !Ax)) return GrB_NULL_POINTER;
    GRB_TRY(matrix_materialize(A));
    GRB_TRY(matrix_ensure_sorted(A));
    std::string *err = &A->err;
    const int64_t nnz = A->nvals;
    if (format == GrB_CSR_FORMAT) {
        GRB_TRY(download_as_u64(Ap, A->csr.ptr, false, A->nrows + 1, err));
        GRB_TRY(download_as_u64(Ai, A->csr.idx, true, nnz, err));
        GRB_TRY(download_vals(Ax, type->code, A->csr.val, A->type, nnz, err));
    } else if (format == GrB_CSC_FORMAT) {
        GRB_TRY(matrix_ensure_twin(A));
        GRB_TRY(download_as_u64(Ap, A->twin.ptr, false, A->ncols + 1, err));
        GRB_TRY(download_as_u64(Ai, A->twin.idx, true, nnz, err));
        GRB_TRY(download_vals(Ax, type->code, A->twin.val, A->type, nnz, err));
    } else {
        int32_t *rowid = dev_alloc_t<int32_t>((size_t)(nnz > 0 ? nnz : 1));
        if (!rowid) return set_error(err, GrB_OUT_OF_MEMORY, "export COO");
        GrB_Info info = expand_row_ids(A, rowid);
        if (!info) info = download_as_u64(Ap, rowid, true, nnz, err);
        if (!info) info = download_as_u64(Ai, A->csr.idx, true, nnz, err);
        if (!info) info = download_vals(Ax, type->code, A->csr.val, A->type, nnz, err);
        dev_free(rowid);
        GRB_TRY(info);
    }
    *Ap_len = np; *Ai_len = ni; *Ax_len = nx;
    CUDA_TRY(err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Matrix_export_csr32(int64_t *Ap, int32_t *Aj, void *Ax, GrB_Index cap, GrB_Matrix A, int sort) {
    CHECK_INIT();
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    if ((GrB_Index)A->nvals > cap) return set_error(&A->err, GrB_INSUFFICIENT_SPACE, "export_csr32: capacity %llu < nvals %lld", (unsigned long long)cap, (long long)A->nvals);
    GRB_TRY(matrix_materialize(A));
    if (sort) GRB_TRY(matrix_ensure_sorted(A));
    std::string *err = &A->err;
    if (Ap) CUDA_TRY(err, cudaMemcpyAsync(Ap, A->csr.ptr, sizeof(int64_t) * (size_t)(A->nrows + 1), cudaMemcpyDeviceToHost, g_stream));
    if (Aj && A->nvals) CUDA_TRY(err, cudaMemcpyAsync(Aj, A->csr.idx, sizeof(int32_t) * (size_t)A->nvals, cudaMemcpyDeviceToHost, g_stream));
    if (Ax && A->nvals) CUDA_TRY(err, cudaMemcpyAsync(Ax, A->csr.val, type_size(A->type) * (size_t)A->nvals, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}

// ---- asynchronous export: the copies run on a separate copy stream, ordered after everything enqueued so far on the compute
// stream, so the D2H of one result overlaps the computation of the next (a product formed in row blocks leaves the device at
// PCIe speed while later blocks are still being multiplied).  The matrix and the (pinned) host arrays must stay alive until
// GrB_cuda_copy_sync() returns.
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_copy_ev = nullptr;
extern "C" GrB_Info GrB_cuda_Matrix_export_csr32_async(int64_t *Ap, int32_t *Aj, void *Ax, GrB_Index cap, GrB_Matrix A) {
    CHECK_INIT();
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    if ((GrB_Index)A->nvals > cap) return set_error(&A->err, GrB_INSUFFICIENT_SPACE, "export_csr32: capacity %llu < nvals %lld", (unsigned long long)cap, (long long)A->nvals);
    GRB_TRY(matrix_materialize(A));
    std::string *err = &A->err;
    if (!g_copy_stream) {
        CUDA_TRY(err, cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(err, cudaEventCreateWithFlags(&g_copy_ev, cudaEventDisableTiming));
    }
    CUDA_TRY(err, cudaEventRecord(g_copy_ev, g_stream));
    CUDA_TRY(err, cudaStreamWaitEvent(g_copy_stream, g_copy_ev, 0));
    if (Ap) CUDA_TRY(err, cudaMemcpyAsync(Ap, A->csr.ptr, sizeof(int64_t) * (size_t)(A->nrows + 1), cudaMemcpyDeviceToHost, g_copy_stream));
    if (Aj && A->nvals) CUDA_TRY(err, cudaMemcpyAsync(Aj, A->csr.idx, sizeof(int32_t) * (size_t)A->nvals, cudaMemcpyDeviceToHost, g_copy_stream));
    if (Ax && A->nvals) CUDA_TRY(err, cudaMemcpyAsync(Ax, A->csr.val, type_size(A->type) * (size_t)A->nvals, cudaMemcpyDeviceToHost, g_copy_stream));
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_copy_sync(void) {
    if (g_copy_stream) CUDA_TRY(nullptr, cudaStreamSynchronize(g_copy_stream));
    return GrB_SUCCESS;
}
// A fence on the copy stream: *ticket names everything enqueued on it so far; GrB_cuda_copy_wait(ticket) returns when that has
// drained, while LATER copies keep running -- so the caller can release the source of copy k while copy k + 1 is in flight.
constexpr int COPY_FENCES = 16;
static cudaEvent_t g_copy_fence[COPY_FENCES];
static bool g_copy_fence_init = false;
static int g_copy_fence_next = 0;
extern "C" GrB_Info GrB_cuda_copy_fence(int *ticket) {
    if (!ticket) return GrB_NULL_POINTER;
    if (!g_copy_stream) { *ticket = -1; return GrB_SUCCESS; }
    if (!g_copy_fence_init) {
        for (int q = 0; q < COPY_FENCES; q++) CUDA_TRY(nullptr, cudaEventCreateWithFlags(&g_copy_fence[q], cudaEventDisableTiming));
        g_copy_fence_init = true;
    }
    const int t = g_copy_fence_next;
    g_copy_fence_next = (g_copy_fence_next + 1) % COPY_FENCES;
    CUDA_TRY(nullptr, cudaEventRecord(g_copy_fence[t], g_copy_stream));
    *ticket = t;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_cuda_copy_wait(int ticket) {
    if (ticket < 0 || ticket >= COPY_FENCES || !g_copy_fence_init) return GrB_SUCCESS;
    CUDA_TRY(nullptr, cudaEventSynchronize(g_copy_fence[ticket]));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Matrix_extractTuples(GrB_Index *I, GrB_Index *J, void *X, GrB_Type xtype, GrB_Index *nvals,
                                                  const GrB_Matrix A) {
    if (!nvals) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    if (*nvals < (GrB_Index)A->nvals) return set_error(&A->err, GrB_INSUFFICIENT_SPACE, "extractTuples: arrays too small");
    GrB_Index n = *nvals, a = *nvals, b = *nvals;
    // COO export; NULL output arrays are allowed by the spec (skip them) -> use scratch-free paths
    CHECK_INIT();
    GRB_TRY(matrix_materialize(A));
    GRB_TRY(matrix_ensure_sorted(A));
    std::string *err = &A->err;
    const int64_t nnz = A->nvals;
    if (I && nnz) {
        int32_t *rowid = dev_alloc_t<int32_t>((size_t)nnz);
        if (!rowid) return set_error(err, GrB_OUT_OF_MEMORY, "extractTuples");
        GrB_Info info = expand_row_ids(A, rowid);
        if (!info) info = download_as_u64(I, rowid, true, nnz, err);
        dev_free(rowid);
        GRB_TRY(info);
    }
    if (J) GRB_TRY(download_as_u64(J, A->csr.idx, true, nnz, err));
    if (X) GRB_TRY(download_vals(X, xtype ? xtype->code : A->type, A->csr.val, A->type, nnz, err));
    (void)n; (void)a; (void)b;
    *nvals = (GrB_Index)nnz;
    CUDA_TRY(err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Matrix_extractElement(void *x, GrB_Type xtype, const GrB_Matrix A, GrB_Index i, GrB_Index j) {
    CHECK_INIT();
    if (!x || !xtype) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    if (i >= (GrB_Index)A->nrows || j >= (GrB_Index)A->ncols) return set_error(&A->err, GrB_INVALID_INDEX, "extractElement: index out of range");
    if (!A->csr.ptr || A->nvals == 0) return GrB_NO_VALUE;
    GRB_TRY(matrix_materialize(A));
    int64_t be[2];
    cudaMemcpyAsync(be, A->csr.ptr + i, 16, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    int64_t len = be[1] - be[0];
    if (len <= 0) return GrB_NO_VALUE;
    std::vector<int32_t> cols((size_t)len);
    cudaMemcpyAsync(cols.data(), A->csr.idx + be[0], sizeof(int32_t) * (size_t)len, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    for (int64_t k = 0; k < len; k++) {
        if ((GrB_Index)cols[(size_t)k] == j) {
            return download_vals(x, xtype->code, (const char *)A->csr.val + (be[0] + k) * type_size(A->type), A->type, 1, &A->err) ||
                           cudaStreamSynchronize(g_stream) != cudaSuccess
                       ? GrB_PANIC : GrB_SUCCESS;
        }
    }
    return GrB_NO_VALUE;
}

// ------------------------------------------------------------------ build (COO -> sorted unique, duplicates reduced by `dup`)
__global__ void make_keys_kernel(uint64_t *keys, const uint64_t *I, const uint64_t *J, int64_t n, uint64_t nrows, uint64_t ncols, int *err) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += s) {
        uint64_t i = I[k], j = J ? J[k] : 0;
        if (i >= nrows || j >= ncols) *err = 1;
        keys[k] = i * ncols + j;
    }
}
__global__ void head_flags_kernel(const uint64_t *keys, int64_t n, int64_t *heads, int *dupflag) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += s) {
        bool head = k == 0 || keys[k] != keys[k - 1];
        heads[k] = head ? 1 : 0;
        if (!head) *dupflag = 1;
    }
}
// heads[] holds the exclusive scan; one thread per run folds the duplicates in input order
template <typename T>
__global__ void reduce_runs_kernel(const uint64_t *keys, const int64_t *pos, const int64_t *scan, int64_t n, const T *xin,
                                   int dup_op, uint64_t ncols, int32_t *out_row, int32_t *out_col, T *out_val) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += s) {
        bool head = k == 0 || keys[k] != keys[k - 1];
        if (!head) continue;
        T acc = xin[pos[k]];
        for (int64_t q = k + 1; q < n && keys[q] == keys[k]; q++) acc = binop<T>(dup_op, acc, xin[pos[q]]);
        int64_t o = scan[k];
        if (out_row) { out_row[o] = (int32_t)(keys[k] / ncols); out_col[o] = (int32_t)(keys[k] % ncols); }
        else out_col[o] = (int32_t)keys[k];   // vectors: the key is the index itself
        out_val[o] = acc;
    }
}
__global__ void row_ptr_from_sorted_rows_kernel(const int32_t *rows, int64_t n, int64_t nrows, int64_t *ptr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nrows) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rows[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    ptr[r] = lo;
}
__global__ void iota64_kernel(int64_t *p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) p[i] = i;
}

struct BuildResult { int32_t *rows = nullptr, *cols = nullptr; void *vals = nullptr; int64_t n = 0; };

// I, J (J may be null for vectors), X are HOST arrays
static GrB_Info build_sorted_unique(BuildResult *out, const GrB_Index *I, const GrB_Index *J, const void *X, int xtype,
                                    int out_type, int64_t n, uint64_t nrows, uint64_t ncols, const GrB_BinaryOp dup,
                                    std::string *err) {
    if (n == 0) return GrB_SUCCESS;
    if (dup && (dup->type != out_type || dup->ztype != out_type))
        return set_error(err, GrB_DOMAIN_MISMATCH, "build: dup operator %s is not typed like the object", dup->name);
    uint64_t *dI = (uint64_t *)upload(I, 8 * (size_t)n, err), *dJ = J ? (uint64_t *)upload(J, 8 * (size_t)n, err) : nullptr;
    void *dXraw = upload(X, type_size(xtype) * (size_t)n, err);
    uint64_t *keys = dev_alloc_t<uint64_t>((size_t)n), *keys_s = dev_alloc_t<uint64_t>((size_t)n);
    int64_t *pos = dev_alloc_t<int64_t>((size_t)n), *pos_s = dev_alloc_t<int64_t>((size_t)n), *heads = dev_alloc_t<int64_t>((size_t)n + 1);
    int *flags = dev_alloc_t<int>(2);
    void *dX = nullptr, *tmp = nullptr;
    GrB_Info info = GrB_SUCCESS;
    if (!dI || (J && !dJ) || !dXraw || !keys || !keys_s || !pos || !pos_s || !heads || !flags) info = set_error(err, GrB_OUT_OF_MEMORY, "build scratch");
    if (!info) {
        dX = dev_alloc(type_size(out_type) * (size_t)n);
        if (!dX) info = set_error(err, GrB_OUT_OF_MEMORY, "build scratch");
    }
    if (!info) info = cast_array(dX, out_type, dXraw, xtype, n, err);
    if (!info) {
        cudaMemsetAsync(flags, 0, 8, g_stream);
        note_launch("make_keys");
        make_keys_kernel<<<grid_for(n), 256, 0, g_stream>>>(keys, dI, dJ, n, nrows, ncols, flags);
        note_launch("iota");
        iota64_kernel<<<grid_for(n), 256, 0, g_stream>>>(pos, n);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_s, pos, pos_s, n, 0, 64, g_stream);
        tmp = dev_alloc(tb);
        if (!tmp) info = set_error(err, GrB_OUT_OF_MEMORY, "build sort scratch");
        else {
            note_launch("cub_radix_sort");
            cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_s, pos, pos_s, n, 0, 64, g_stream);
            if (e != cudaSuccess) info = cuda_fail(err, e, "cub radix sort (build)");
        }
    }
    int64_t uniq = 0;
    if (!info) {
        note_launch("head_flags");
        head_flags_kernel<<<grid_for(n), 256, 0, g_stream>>>(keys_s, n, heads, flags + 1);
        cudaMemsetAsync(heads + n, 0, 8, g_stream);
        info = exclusive_scan_i64(heads, n + 1, err);
    }
    if (!info) {
        int h[2] = {0, 0};
        cudaMemcpyAsync(h, flags, 8, cudaMemcpyDeviceToHost, g_stream);
        uniq = read_i64(heads + n);
        if (h[0]) info = set_error(err, GrB_INDEX_OUT_OF_BOUNDS, "build: an index is out of bounds");
        else if (h[1] && !dup) info = set_error(err, GrB_INVALID_VALUE, "build: duplicate indices and no dup operator");
    }
    if (!info) {
        out->rows = J ? dev_alloc_t<int32_t>((size_t)uniq) : nullptr;
        out->cols = dev_alloc_t<int32_t>((size_t)uniq);
        out->vals = dev_alloc(type_size(out_type) * (size_t)uniq);
        out->n = uniq;
        if ((J && !out->rows) || !out->cols || !out->vals) info = set_error(err, GrB_OUT_OF_MEMORY, "build result");
    }
    if (!info) {
        LAUNCH_NOTE("reduce_runs");
        GRB_DISPATCH_TYPE(out_type, T,
                          (reduce_runs_kernel<T><<<grid_for(n), 256, 0, g_stream>>>(keys_s, pos_s, heads, n, (const T *)dX, dup ? dup->opcode : OP_SECOND,
                                                                               J ? ncols : 1, out->rows, out->cols, (T *)out->vals)));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(err, e, "build");
    }
    dev_free(dI); dev_free(dJ); dev_free(dXraw); dev_free(keys); dev_free(keys_s); dev_free(pos); dev_free(pos_s);
    dev_free(heads); dev_free(flags); dev_free(dX); dev_free(tmp);
    if (info) { dev_free(out->rows); dev_free(out->cols); dev_free(out->vals); *out = BuildResult(); }
    return info;
}

extern "C" GrB_Info GrB_cuda_Matrix_build(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, const void *X, GrB_Type xtype,
                                          GrB_Index nvals, const GrB_BinaryOp dup) {
    CHECK_INIT();
    if (!valid(C)) return GrB_UNINITIALIZED_OBJECT;
    if (nvals && (!I || !J || !X)) return set_error(&C->err, GrB_NULL_POINTER, "build: null array");
    if (!xtype) return GrB_NULL_POINTER;
    if (C->nvals != 0) return set_error(&C->err, GrB_OUTPUT_NOT_EMPTY, "build: output already has entries");
    BuildResult r;
    GRB_TRY(build_sorted_unique(&r, I, J, X, xtype->code, C->type, (int64_t)nvals, (uint64_t)C->nrows, (uint64_t)C->ncols, dup, &C->err));
    matrix_release(C);
    C->csr.ptr = dev_alloc_t<int64_t>((size_t)C->nrows + 1);
    if (!C->csr.ptr) { dev_free(r.rows); dev_free(r.cols); dev_free(r.vals); return set_error(&C->err, GrB_OUT_OF_MEMORY, "build"); }
    if (r.n == 0) {
        fill_bytes(C->csr.ptr, 0, sizeof(int64_t) * ((size_t)C->nrows + 1));
        C->csr.idx = dev_alloc_t<int32_t>(1);
        C->csr.val = dev_alloc(16);
    } else {
        note_launch("row_ptr_from_rows");
        row_ptr_from_sorted_rows_kernel<<<(unsigned)((C->nrows + 1 + 255) / 256), 256, 0, g_stream>>>(r.rows, r.n, C->nrows, C->csr.ptr);
        C->csr.idx = r.cols;
        C->csr.val = r.vals;
        dev_free(r.rows);
    }
    C->nvals = r.n;
    C->jumbled = false;
    CUDA_TRY(&C->err, cudaGetLastError());
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ vectors
template <typename T>
__global__ void scatter_vec_kernel(const int32_t *idx, const T *vals, int64_t n, T *out_vals, uint8_t *out_present) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += s) {
        out_vals[idx[k]] = vals[k];
        out_present[idx[k]] = 1;
    }
}

extern "C" GrB_Info GrB_cuda_Vector_build(GrB_Vector w, const GrB_Index *I, const void *X, GrB_Type xtype, GrB_Index nvals,
                                          const GrB_BinaryOp dup) {
    CHECK_INIT();
    if (!valid(w)) return GrB_UNINITIALIZED_OBJECT;
    if (nvals && (!I || !X)) return set_error(&w->err, GrB_NULL_POINTER, "build: null array");
    if (!xtype) return GrB_NULL_POINTER;
    GRB_TRY(vector_count(w));
    if (w->nvals != 0) return set_error(&w->err, GrB_OUTPUT_NOT_EMPTY, "build: output already has entries");
    if (nvals == 0) return GrB_SUCCESS;
    // a vector too long for the dense layout (2^59 + 1 in the reference's tests) must fail cleanly: bad indices first, then the
    // size; for every other vector the bounds check is done on the device while the sort keys are made
    if ((uint64_t)w->n > (uint64_t)INT32_MAX)
        for (GrB_Index k = 0; k < nvals; k++)
            if (I[k] >= (GrB_Index)w->n) return set_error(&w->err, GrB_INDEX_OUT_OF_BOUNDS, "build: index %llu out of bounds", (unsigned long long)I[k]);
    GRB_TRY(vector_ensure_arrays(w));
    BuildResult r;
    GRB_TRY(build_sorted_unique(&r, I, nullptr, X, xtype->code, w->type, (int64_t)nvals, (uint64_t)w->n, 1, dup, &w->err));
    {
        LAUNCH_NOTE("scatter_vec");
        GRB_DISPATCH_TYPE(w->type, T, (scatter_vec_kernel<T><<<grid_for(r.n), 256, 0, g_stream>>>(r.cols, (const T *)r.vals, r.n, (T *)w->vals, w->present)));
    }
    w->nvals = r.n;
    dev_free(r.cols); dev_free(r.vals);
    CUDA_TRY(&w->err, cudaGetLastError());
    return GrB_SUCCESS;
}

__global__ void compact_indices64_kernel(const uint8_t *present, int64_t n, const int64_t *scan, uint64_t *out_idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) if (present[i]) out_idx[scan[i]] = (uint64_t)i;
}
template <typename T> __global__ void compact_vals_kernel(const uint8_t *present, int64_t n, const int64_t *scan, const T *vals, T *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) if (present[i]) out[scan[i]] = vals[i];
}
__global__ void present_to_i64_kernel(const uint8_t *present, int64_t n, int64_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) out[i] = present[i] ? 1 : 0;
}

extern "C" GrB_Info GrB_cuda_Vector_extractTuples(GrB_Index *I, void *X, GrB_Type xtype, GrB_Index *nvals, const GrB_Vector v) {
    CHECK_INIT();
    if (!nvals) return GrB_NULL_POINTER;
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(vector_count(v));
    if (*nvals < (GrB_Index)v->nvals) return set_error(&v->err, GrB_INSUFFICIENT_SPACE, "extractTuples: arrays too small");
    const int64_t n = v->n, nv = v->nvals;
    *nvals = (GrB_Index)nv;
    if (nv == 0) return GrB_SUCCESS;
    std::string *err = &v->err;
    int64_t *scan = dev_alloc_t<int64_t>((size_t)n + 1);
    uint64_t *didx = dev_alloc_t<uint64_t>((size_t)nv);
    const int ht = xtype ? xtype->code : v->type;
    void *dval = dev_alloc(type_size(v->type) * (size_t)nv);
    if (!scan || !didx || !dval) { dev_free(scan); dev_free(didx); dev_free(dval); return set_error(err, GrB_OUT_OF_MEMORY, "extractTuples"); }
    note_launch("present_to_i64");
    present_to_i64_kernel<<<grid_for(n), 256, 0, g_stream>>>(v->present, n, scan);
    cudaMemsetAsync(scan + n, 0, 8, g_stream);
    GrB_Info info = exclusive_scan_i64(scan, n + 1, err);
    if (!info) {
        note_launch("compact_indices");
        compact_indices64_kernel<<<grid_for(n), 256, 0, g_stream>>>(v->present, n, scan, didx);
        LAUNCH_NOTE("compact_vals");
        GRB_DISPATCH_TYPE(v->type, T, (compact_vals_kernel<T><<<grid_for(n), 256, 0, g_stream>>>(v->present, n, scan, (const T *)v->vals, (T *)dval)));
        if (I) cudaMemcpyAsync(I, didx, 8 * (size_t)nv, cudaMemcpyDeviceToHost, g_stream);
        if (X) info = download_vals(X, ht, dval, v->type, nv, err);
    }
    dev_free(scan); dev_free(didx); dev_free(dval);
    GRB_TRY(info);
    CUDA_TRY(err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Vector_setElement(GrB_Vector w, const void *x, GrB_Type xtype, GrB_Index i) {
    CHECK_INIT();
    if (!x || !xtype) return GrB_NULL_POINTER;
    if (!valid(w)) return GrB_UNINITIALIZED_OBJECT;
    if (i >= (GrB_Index)w->n) return set_error(&w->err, GrB_INVALID_INDEX, "setElement: index out of range");
    GRB_TRY(vector_ensure_arrays(w));
    void *d = upload(x, type_size(xtype->code), &w->err);
    if (!d) return GrB_OUT_OF_MEMORY;
    GrB_Info info = cast_array((char *)w->vals + i * type_size(w->type), w->type, d, xtype->code, 1, &w->err);
    dev_free(d);
    GRB_TRY(info);
    CUDA_TRY(&w->err, cudaMemsetAsync(w->present + i, 1, 1, g_stream));
    w->nvals = -1;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Vector_extractElement(void *x, GrB_Type xtype, const GrB_Vector v, GrB_Index i) {
    CHECK_INIT();
    if (!x || !xtype) return GrB_NULL_POINTER;
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    if (i >= (GrB_Index)v->n) return set_error(&v->err, GrB_INVALID_INDEX, "extractElement: index out of range");
    if (!v->present) return GrB_NO_VALUE;
    uint8_t p = 0;
    cudaMemcpyAsync(&p, v->present + i, 1, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    if (!p) return GrB_NO_VALUE;
    GRB_TRY(download_vals(x, xtype->code, (const char *)v->vals + i * type_size(v->type), v->type, 1, &v->err));
    CUDA_TRY(&v->err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_Vector_removeElement(GrB_Vector w, GrB_Index i) {
    if (!valid(w)) return GrB_UNINITIALIZED_OBJECT;
    if (i >= (GrB_Index)w->n) return set_error(&w->err, GrB_INVALID_INDEX, "removeElement: index out of range");
    if (!w->present) return GrB_SUCCESS;
    CUDA_TRY(&w->err, cudaMemsetAsync(w->present + i, 0, 1, g_stream));
    w->nvals = -1;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Vector_import_dense(GrB_Vector *vout, GrB_Type type, GrB_Index n, const void *vals,
                                                 const uint8_t *present, int on_device) {
    CHECK_INIT();
    if (!vout || !type || (!vals && n)) return GrB_NULL_POINTER;
    GrB_Vector v;
    GRB_TRY(GrB_Vector_new(&v, type, n));
    GrB_Info info = vector_ensure_arrays(v);
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!info && n) {
        cudaError_t e = cudaMemcpyAsync(v->vals, vals, type_size(type->code) * (size_t)n, kind, g_stream);
        if (e == cudaSuccess) e = present ? cudaMemcpyAsync(v->present, present, (size_t)n, kind, g_stream) : cudaMemsetAsync(v->present, 1, (size_t)n, g_stream);
        if (e != cudaSuccess) info = cuda_fail(&v->err, e, "import_dense");
    }
    if (info) { set_last_error(v->err.c_str()); GrB_Vector_free(&v); return info; }
    v->nvals = present ? -1 : (int64_t)n;
    *vout = v;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Vector_export_dense(void *vals, uint8_t *present, const GrB_Vector v) {
    CHECK_INIT();
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(vector_ensure_arrays(v));
    if (vals && v->n) CUDA_TRY(&v->err, cudaMemcpyAsync(vals, v->vals, type_size(v->type) * (size_t)v->n, cudaMemcpyDeviceToHost, g_stream));
    if (present && v->n) CUDA_TRY(&v->err, cudaMemcpyAsync(present, v->present, (size_t)v->n, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(&v->err, cudaStreamSynchronize(g_stream));
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ vector -> n x 1 matrix (column vector)
// The reference's Vector.inner / Vector.outer run GrB_vxm / GrB_mxm on the vector "cast" to a matrix
// (graphblas/core/vector.py:193-209 `_as_matrix`, :1715-1787); on the vanilla backend that is a fresh n x 1 Matrix filled by
// a column assign.  Here: presence bytes -> row pointers (exclusive scan), values compacted, every column index 0.
__global__ void vec_present_to_counts_kernel(int64_t n, const uint8_t *__restrict__ present, int64_t *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i <= n; i += s) cnt[i] = (i < n && present[i]) ? 1 : 0;
}
__global__ void vec_to_column_kernel(int64_t n, const uint8_t *__restrict__ present, const unsigned char *__restrict__ vals,
                                     size_t es, const int64_t *__restrict__ ptr, int32_t *__restrict__ idx,
                                     unsigned char *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) {
        if (!present[i]) continue;
        const int64_t k = ptr[i];
        idx[k] = 0;
        for (size_t b = 0; b < es; b++) out[(size_t)k * es + b] = vals[(size_t)i * es + b];
    }
}

extern "C" GrB_Info GrB_cuda_Matrix_from_Vector(GrB_Matrix *Aout, const GrB_Vector v) {
    CHECK_INIT();
    if (!Aout) return GrB_NULL_POINTER;
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(vector_ensure_arrays(v));
    GRB_TRY(vector_count(v));
    GrB_Matrix A = nullptr;
    GRB_TRY(matrix_new_shell(&A, v->type, v->n, 1));
    GrB_Info info = matrix_alloc_csr(A, v->nvals);
    if (!info) {
        const int blocks = (int)std::min<int64_t>((v->n + 256) / 256, (int64_t)g_num_sms * 16);
        note_launch("vec_present_to_counts");
        vec_present_to_counts_kernel<<<blocks, 256, 0, g_stream>>>(v->n, v->present, A->csr.ptr);
        info = exclusive_scan_i64(A->csr.ptr, v->n + 1, &A->err);
    }
    if (!info && v->nvals > 0) {
        const int blocks = (int)std::min<int64_t>((v->n + 255) / 256, (int64_t)g_num_sms * 16);
        note_launch("vec_to_column");
        vec_to_column_kernel<<<blocks, 256, 0, g_stream>>>(v->n, v->present, (const unsigned char *)v->vals, type_size(v->type),
                                                           A->csr.ptr, A->csr.idx, (unsigned char *)A->csr.val);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(&A->err, e, "vector to column matrix");
    }
    if (info) { set_last_error(A->err.c_str()); GrB_Matrix_free(&A); return info; }
    A->nvals = v->nvals;
    A->jumbled = false;
    *Aout = A;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ GrB_Matrix_diag (reference core/vector.py:605-628: Vector.diag)
__global__ void diag_counts_kernel(int64_t n, int64_t vn, int64_t roff, const uint8_t *__restrict__ present, int64_t *__restrict__ cnt) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; r <= n; r += s) {
        const int64_t i = r - roff;
        cnt[r] = (r < n && i >= 0 && i < vn && present[i]) ? 1 : 0;
    }
}
__global__ void diag_fill_kernel(int64_t vn, int64_t roff, int64_t coff, const uint8_t *__restrict__ present,
                                 const unsigned char *__restrict__ vals, size_t es, const int64_t *__restrict__ ptr,
                                 int32_t *__restrict__ idx, unsigned char *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < vn; i += s) {
        if (!present[i]) continue;
        const int64_t k = ptr[i + roff];
        idx[k] = (int32_t)(i + coff);
        for (size_t b = 0; b < es; b++) out[(size_t)k * es + b] = vals[(size_t)i * es + b];
    }
}
// C = the (n + |k|) x (n + |k|) matrix with v on its k-th diagonal: v(i) at (i, i + k) for k >= 0, at (i - k, i) for k < 0
extern "C" GrB_Info GrB_Matrix_diag(GrB_Matrix *C, const GrB_Vector v, int64_t k) {
    CHECK_INIT();
    if (!C) return GrB_NULL_POINTER;
    if (!valid(v)) return GrB_UNINITIALIZED_OBJECT;
    const int64_t n = v->n + (k < 0 ? -k : k);
    if (n > (int64_t)INT32_MAX) return set_error(nullptr, GrB_NOT_IMPLEMENTED, "GrB_Matrix_diag: dimensions must be < 2^31");
    GRB_TRY(vector_ensure_arrays(v));
    GRB_TRY(vector_count(v));
    GrB_Matrix A = nullptr;
    GRB_TRY(matrix_new_shell(&A, v->type, n, n));
    GrB_Info info = matrix_alloc_csr(A, v->nvals);
    const int64_t roff = k < 0 ? -k : 0, coff = k > 0 ? k : 0;
    if (!info) {
        const int blocks = (int)std::min<int64_t>((n + 256) / 256, (int64_t)g_num_sms * 16);
        note_launch("diag_counts");
        diag_counts_kernel<<<blocks, 256, 0, g_stream>>>(n, v->n, roff, v->present, A->csr.ptr);
        info = exclusive_scan_i64(A->csr.ptr, n + 1, &A->err);
    }
    if (!info && v->nvals > 0) {
        const int blocks = (int)std::min<int64_t>((v->n + 255) / 256, (int64_t)g_num_sms * 16);
        note_launch("diag_fill");
        diag_fill_kernel<<<blocks, 256, 0, g_stream>>>(v->n, roff, coff, v->present, (const unsigned char *)v->vals, type_size(v->type),
                                                       A->csr.ptr, A->csr.idx, (unsigned char *)A->csr.val);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(&A->err, e, "GrB_Matrix_diag");
    }
    if (info) { set_last_error(A->err.c_str()); GrB_Matrix_free(&A); return info; }
    A->nvals = v->nvals;
    A->jumbled = false;
    *C = A;
    return GrB_SUCCESS;
}
