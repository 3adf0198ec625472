// structure.cu -- structural maintenance of the device CSR: lazy row sort ("jumbled" results, as the
// reference's C library also leaves them: graphblas/core/matrix.py:1631-1644 sorts after export)
// and the cached CSR-of-the-transpose twin used by vxm / A.T operands.
#include <cub/cub.cuh>
#include <limits.h>
#include <vector>

#include "grb_ops.cuh"
void csr_drop_band(CsrArrays &c);   // spmv_band.cu

// ------------------------------------------------------------------ sortedness check
__global__ void check_sorted_kernel(int64_t nrows, const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                    int *__restrict__ flag) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w; i < nrows; i += nw) {
        int64_t b = ptr[i], e = ptr[i + 1];
        bool bad = false;
        for (int64_t k = b + 1 + lane; k < e; k += 32) bad |= idx[k - 1] >= idx[k];
        if (bad) *flag = 1;
    }
}

// classify rows by length: <=32 handled in place by the warp kernel; medium rows and long rows are listed
__global__ void sort_classify_kernel(int64_t nrows, const int64_t *__restrict__ ptr, int mid_cap, int mid2_cap, int big_cap,
                                     int32_t *__restrict__ mid_rows, int32_t *__restrict__ mid2_rows, int32_t *__restrict__ big_rows,
                                     int32_t *__restrict__ huge_rows, unsigned int *__restrict__ counters) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    int64_t len = ptr[i + 1] - ptr[i];
    if (len <= 32) return;
    if (len <= mid_cap) mid_rows[atomicAdd(&counters[0], 1u)] = (int32_t)i;
    else if (len <= mid2_cap) mid2_rows[atomicAdd(&counters[3], 1u)] = (int32_t)i;
    else if (len <= big_cap) big_rows[atomicAdd(&counters[1], 1u)] = (int32_t)i;
    else huge_rows[atomicAdd(&counters[2], 1u)] = (int32_t)i;
}

template <typename V>
__global__ void sort_rows_warp_kernel(int64_t nrows, const int64_t *__restrict__ ptr, int32_t *__restrict__ idx,
                                      V *__restrict__ val) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w; i < nrows; i += nw) {
        const int64_t b = ptr[i];
        const int len = (int)(ptr[i + 1] - b);
        if (len < 2 || len > 32) continue;
        int key = lane < len ? idx[b + lane] : INT_MAX;
        int src = lane;
        V v = V();
        if (lane < len) v = val[b + lane];
        // bitonic sort of (key, src) across the warp
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                int okey = __shfl_xor_sync(0xffffffffu, key, j);
                int osrc = __shfl_xor_sync(0xffffffffu, src, j);
                bool up = ((lane & k) == 0);
                bool lower = ((lane & j) == 0);
                bool take = (lower == up) ? (okey < key) : (okey > key);
                if (take) { key = okey; src = osrc; }
            }
        }
        // fetch the value that belongs to the key now held by this lane
        V sv;
        if constexpr (sizeof(V) == 8) {
            unsigned long long bits;
            memcpy(&bits, &v, 8);
            bits = __shfl_sync(0xffffffffu, bits, src);
            memcpy(&sv, &bits, 8);
        } else {
            unsigned int bits = 0;
            memcpy(&bits, &v, sizeof(V));
            bits = __shfl_sync(0xffffffffu, bits, src);
            memcpy(&sv, &bits, sizeof(V));
        }
        if (lane < len) { idx[b + lane] = key; val[b + lane] = sv; }
    }
}

// block per listed row; bitonic sort of (key, position) in shared memory, then values are permuted through registers
template <typename V, int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS)
sort_rows_block_kernel(const int32_t *__restrict__ rows, const int64_t *__restrict__ ptr, int32_t *__restrict__ idx,
                       V *__restrict__ val) {
    extern __shared__ int s_mem[];
    int *s_key = s_mem;
    int *s_pos = s_mem + CAP;
    const int64_t i = rows[blockIdx.x];
    const int64_t b = ptr[i];
    const int len = (int)(ptr[i + 1] - b);
    int n = 64;
    while (n < len) n <<= 1;
    for (int t = threadIdx.x; t < n; t += THREADS) {
        s_key[t] = t < len ? idx[b + t] : INT_MAX;
        s_pos[t] = t;
    }
    __syncthreads();
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n; t += THREADS) {
                int p = t ^ j;
                if (p > t) {
                    bool up = ((t & k) == 0);
                    int a = s_key[t], c = s_key[p];
                    if ((a > c) == up) {
                        s_key[t] = c; s_key[p] = a;
                        int q = s_pos[t]; s_pos[t] = s_pos[p]; s_pos[p] = q;
                    }
                }
            }
            __syncthreads();
        }
    }
    constexpr int PER = (CAP + THREADS - 1) / THREADS;
    V regs[PER];
#pragma unroll
    for (int r = 0; r < PER; r++) {
        int t = threadIdx.x + r * THREADS;
        if (t < len) regs[r] = val[b + s_pos[t]];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PER; r++) {
        int t = threadIdx.x + r * THREADS;
        if (t < len) { val[b + t] = regs[r]; idx[b + t] = s_key[t]; }
    }
}

template <typename V> __global__ void gather_vals_kernel(V *__restrict__ dst, const V *__restrict__ src, const int64_t *__restrict__ perm, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[perm[i]];
}
__global__ void iota_kernel(int64_t *p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = i;
}

static GrB_Info gather_by_size(void *dst, const void *src, const int64_t *perm, int64_t n, size_t esize, std::string *err) {
    if (n <= 0) return GrB_SUCCESS;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)g_num_sms * 16);
    LAUNCH_NOTE("gather_vals");
    switch (esize) {
        case 1: gather_vals_kernel<uint8_t><<<blocks, 256, 0, g_stream>>>((uint8_t *)dst, (const uint8_t *)src, perm, n); break;
        case 2: gather_vals_kernel<uint16_t><<<blocks, 256, 0, g_stream>>>((uint16_t *)dst, (const uint16_t *)src, perm, n); break;
        case 4: gather_vals_kernel<uint32_t><<<blocks, 256, 0, g_stream>>>((uint32_t *)dst, (const uint32_t *)src, perm, n); break;
        default: gather_vals_kernel<uint64_t><<<blocks, 256, 0, g_stream>>>((uint64_t *)dst, (const uint64_t *)src, perm, n); break;
    }
    CUDA_TRY(err, cudaGetLastError());
    return GrB_SUCCESS;
}

// all rows longer than the shared-memory sorter's capacity, in ONE batch: their lengths are scanned into a compact scratch
// layout, the rows are copied there, cub::DeviceSegmentedSort sorts every row as a segment, and the result is copied back
// (the previous per-row host loop cost ~2.5 ms and two synchronisations per row)
__global__ void huge_lens_kernel(int64_t n_huge, const int32_t *__restrict__ rows, const int64_t *__restrict__ ptr, int64_t *__restrict__ off) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_huge) return;
    off[r] = r < n_huge ? ptr[rows[r] + 1] - ptr[rows[r]] : 0;
}
template <typename V>
__global__ void huge_copy_kernel(int64_t n_huge, const int32_t *__restrict__ rows, const int64_t *__restrict__ ptr,
                                 const int64_t *__restrict__ off, int32_t *__restrict__ idx, V *__restrict__ val,
                                 int32_t *__restrict__ tk, V *__restrict__ tv, int to_scratch) {
    for (int64_t r = blockIdx.y; r < n_huge; r += gridDim.y) {
        const int64_t src = ptr[rows[r]], dst = off[r], len = off[r + 1] - dst;
        for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < len; k += (int64_t)gridDim.x * blockDim.x) {
            if (to_scratch) { tk[dst + k] = idx[src + k]; tv[dst + k] = val[src + k]; }
            else { idx[src + k] = tk[dst + k]; val[src + k] = tv[dst + k]; }
        }
    }
}
template <typename V> static GrB_Info sort_huge_rows(GrB_Matrix A, const int32_t *huge, int64_t n_huge) {
    std::string *err = &A->err;
    int64_t *off = dev_alloc_t<int64_t>((size_t)n_huge + 1);
    if (!off) return set_error(err, GrB_OUT_OF_MEMORY, "row sort offsets");
    note_launch("huge_lens");
    huge_lens_kernel<<<(unsigned)((n_huge + 1 + 255) / 256), 256, 0, g_stream>>>(n_huge, huge, A->csr.ptr, off);
    GrB_Info info = exclusive_scan_i64(off, n_huge + 1, err);
    int64_t total = 0;
    if (!info) total = read_i64(off + n_huge);
    int32_t *k_in = nullptr, *k_out = nullptr;
    V *v_in = nullptr, *v_out = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    if (!info && total > 0) {
        k_in = dev_alloc_t<int32_t>((size_t)total); k_out = dev_alloc_t<int32_t>((size_t)total);
        v_in = dev_alloc_t<V>((size_t)total); v_out = dev_alloc_t<V>((size_t)total);
        if (!k_in || !k_out || !v_in || !v_out) info = set_error(err, GrB_OUT_OF_MEMORY, "row sort scratch (%lld entries)", (long long)total);
        const dim3 grid(64, (unsigned)std::min<int64_t>(n_huge, 32768));
        if (!info) {
            note_launch("huge_copy");
            huge_copy_kernel<V><<<grid, 256, 0, g_stream>>>(n_huge, huge, A->csr.ptr, off, A->csr.idx, (V *)A->csr.val, k_in, v_in, 1);
            cub::DeviceSegmentedSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, total, (int64_t)n_huge, off, off + 1, g_stream);
            tmp = dev_alloc(tb);
            if (!tmp) info = set_error(err, GrB_OUT_OF_MEMORY, "row sort scratch");
        }
        if (!info) {
            note_launch("cub_segmented_sort");
            cudaError_t e = cub::DeviceSegmentedSort::SortPairs(tmp, tb, k_in, k_out, v_in, v_out, total, (int64_t)n_huge, off, off + 1, g_stream);
            if (e != cudaSuccess) info = cuda_fail(err, e, "cub segmented sort");
        }
        if (!info) {
            note_launch("huge_copy");
            huge_copy_kernel<V><<<grid, 256, 0, g_stream>>>(n_huge, huge, A->csr.ptr, off, A->csr.idx, (V *)A->csr.val, k_out, v_out, 0);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) info = cuda_fail(err, e, "row sort copy back");
        }
    }
    dev_free(off); dev_free(k_in); dev_free(k_out); dev_free(v_in); dev_free(v_out); dev_free(tmp);
    return info;
}

template <typename V> static GrB_Info sort_rows_typed(GrB_Matrix A) {
    std::string *err = &A->err;
    const int64_t nrows = A->nrows;
    // rows of 513 - 2048 entries get their own class: the bitonic network is a chain of barriers, so what it wants is many CTAs
    // per SM, and a 16 KB CTA fits four times more often than the 64 KB one of the 8192 class (which took 166 of the 222 ms
    // of sorting the 2.5 G-entry scale-22 product when it served all rows above 512)
    constexpr int MID_CAP = 512, MID2_CAP = 2048, BIG_CAP = 8192;
    int32_t *mid = dev_alloc_t<int32_t>((size_t)nrows), *mid2 = dev_alloc_t<int32_t>((size_t)nrows), *big = dev_alloc_t<int32_t>((size_t)nrows),
            *huge = dev_alloc_t<int32_t>((size_t)nrows);
    unsigned int *counters = dev_alloc_t<unsigned int>(4);
    if (!mid || !mid2 || !big || !huge || !counters) {
        dev_free(mid); dev_free(mid2); dev_free(big); dev_free(huge); dev_free(counters);
        return set_error(err, GrB_OUT_OF_MEMORY, "row sort lists");
    }
    cudaMemsetAsync(counters, 0, 16, g_stream);
    {
        LAUNCH_NOTE("sort_classify");
        sort_classify_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, g_stream>>>(nrows, A->csr.ptr, MID_CAP, MID2_CAP, BIG_CAP, mid, mid2, big, huge, counters);
    }
    {
        int blocks = (int)std::min<int64_t>((nrows + 7) / 8, (int64_t)g_num_sms * 32);
        LAUNCH_NOTE("sort_rows_warp");
        sort_rows_warp_kernel<V><<<blocks, 256, 0, g_stream>>>(nrows, A->csr.ptr, A->csr.idx, (V *)A->csr.val);
    }
    unsigned int h[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(h, counters, 16, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    GrB_Info info = GrB_SUCCESS;
    if (h[0]) {
        LAUNCH_NOTE("sort_rows_block512");
        sort_rows_block_kernel<V, MID_CAP, 128><<<h[0], 128, MID_CAP * 8, g_stream>>>(mid, A->csr.ptr, A->csr.idx, (V *)A->csr.val);
    }
    if (h[3]) {
        LAUNCH_NOTE("sort_rows_block2048");
        sort_rows_block_kernel<V, MID2_CAP, 256><<<h[3], 256, MID2_CAP * 8, g_stream>>>(mid2, A->csr.ptr, A->csr.idx, (V *)A->csr.val);
    }
    if (h[1]) {
        cudaFuncSetAttribute(sort_rows_block_kernel<V, BIG_CAP, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG_CAP * 8);
        LAUNCH_NOTE("sort_rows_block8192");
        sort_rows_block_kernel<V, BIG_CAP, 512><<<h[1], 512, BIG_CAP * 8, g_stream>>>(big, A->csr.ptr, A->csr.idx, (V *)A->csr.val);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) info = cuda_fail(err, e, "row sort");
    if (!info && h[2]) info = sort_huge_rows<V>(A, huge, (int64_t)h[2]);
    dev_free(mid); dev_free(mid2); dev_free(big); dev_free(huge); dev_free(counters);
    return info;
}

GrB_Info matrix_ensure_sorted(GrB_Matrix A) {
    if (A->csr.end) GRB_TRY(matrix_materialize(A));   // the sort kernels walk a compact CSR
    if (!A->jumbled || !A->csr.ptr || A->nvals == 0) { A->jumbled = false; return GrB_SUCCESS; }
    // the values are only moved, never interpreted: sort by byte width
    GrB_Info info;
    switch (type_size(A->type)) {
        case 1: info = sort_rows_typed<uint8_t>(A); break;
        case 2: info = sort_rows_typed<uint16_t>(A); break;
        case 4: info = sort_rows_typed<uint32_t>(A); break;
        default: info = sort_rows_typed<uint64_t>(A); break;
    }
    if (!info) A->jumbled = false;
    csr_drop_hot(A->csr);   // positions of the column indices moved
    csr_drop_band(A->csr);
    return info;
}

extern "C" GrB_Info GrB_cuda_Matrix_compact(GrB_Matrix A) {
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    return matrix_materialize(A);
}

extern "C" GrB_Info GrB_cuda_Matrix_sort(GrB_Matrix A) {
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    return matrix_ensure_sorted(A);
}

// returns 1 if some row has unsorted / duplicate column indices
GrB_Info matrix_check_sorted(GrB_Matrix A, bool *sorted) {
    *sorted = true;
    if (!A->csr.ptr || A->nvals == 0) return GrB_SUCCESS;
    GRB_TRY(matrix_materialize(A));
    int *flag = dev_alloc_t<int>(1);
    if (!flag) return GrB_OUT_OF_MEMORY;
    cudaMemsetAsync(flag, 0, 4, g_stream);
    int blocks = (int)std::min<int64_t>((A->nrows + 7) / 8, (int64_t)g_num_sms * 32);
    {
        LAUNCH_NOTE("check_sorted");
        check_sorted_kernel<<<blocks, 256, 0, g_stream>>>(A->nrows, A->csr.ptr, A->csr.idx, flag);
    }
    int h = 0;
    cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    dev_free(flag);
    CUDA_TRY(&A->err, cudaGetLastError());
    *sorted = (h == 0);
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ transpose twin
__global__ void expand_rows_kernel(int64_t nrows, const int64_t *__restrict__ ptr, int32_t *__restrict__ rowid) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w; i < nrows; i += nw)
        for (int64_t k = ptr[i] + lane; k < ptr[i + 1]; k += 32) rowid[k] = (int32_t)i;
}
// ptr[c] = first position whose sorted key >= c
__global__ void lower_bound_ptr_kernel(const int32_t *__restrict__ keys, int64_t n, int64_t nseg, int64_t *__restrict__ ptr) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nseg) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < c) lo = mid + 1;
        else hi = mid;
    }
    ptr[c] = lo;
}
__global__ void gather_i32_kernel(int32_t *__restrict__ dst, const int32_t *__restrict__ src, const int64_t *__restrict__ perm, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[perm[i]];
}

GrB_Info expand_row_ids(const GrB_Matrix A, int32_t *rowid) {
    if (A->nrows == 0 || A->nvals == 0) return GrB_SUCCESS;
    int blocks = (int)std::min<int64_t>((A->nrows + 7) / 8, (int64_t)g_num_sms * 32);
    LAUNCH_NOTE("expand_rows");
    expand_rows_kernel<<<blocks, 256, 0, g_stream>>>(A->nrows, A->csr.ptr, rowid);
    CUDA_TRY(&A->err, cudaGetLastError());
    return GrB_SUCCESS;
}

GrB_Info matrix_ensure_twin(GrB_Matrix A) {
    if (A->has_twin) return GrB_SUCCESS;
    GRB_TRY(matrix_materialize(A));
    std::string *err = &A->err;
    const int64_t nnz = A->nvals, ncols = A->ncols;
    const size_t es = type_size(A->type);
    CsrArrays tw;
    tw.ptr = dev_alloc_t<int64_t>((size_t)ncols + 1);
    tw.idx = dev_alloc_t<int32_t>((size_t)(nnz > 0 ? nnz : 1));
    tw.val = dev_alloc((size_t)(nnz > 0 ? nnz : 1) * es);
    if (!tw.ptr || !tw.idx || !tw.val) { csr_free(tw); return set_error(err, GrB_OUT_OF_MEMORY, "transpose twin"); }
    GrB_Info info = GrB_SUCCESS;
    if (nnz == 0) {
        fill_bytes(tw.ptr, 0, sizeof(int64_t) * ((size_t)ncols + 1));
    } else {
        int32_t *rowid = dev_alloc_t<int32_t>((size_t)nnz), *keys_out = dev_alloc_t<int32_t>((size_t)nnz);
        int64_t *pos_in = dev_alloc_t<int64_t>((size_t)nnz), *pos_out = dev_alloc_t<int64_t>((size_t)nnz);
        void *tmp = nullptr;
        if (!rowid || !keys_out || !pos_in || !pos_out) info = set_error(err, GrB_OUT_OF_MEMORY, "transpose scratch");
        if (!info) info = expand_row_ids(A, rowid);
        if (!info) {
            int blocks = (int)std::min<int64_t>((nnz + 255) / 256, (int64_t)g_num_sms * 16);
            note_launch("iota");
            iota_kernel<<<blocks, 256, 0, g_stream>>>(pos_in, nnz);
            int bits = 1;
            while (bits < 32 && ((int64_t)1 << bits) < ncols) bits++;
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, A->csr.idx, keys_out, pos_in, pos_out, nnz, 0, bits, g_stream);
            tmp = dev_alloc(tb);
            if (!tmp) info = set_error(err, GrB_OUT_OF_MEMORY, "transpose sort scratch");
            else {
                note_launch("cub_radix_sort");
                cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, A->csr.idx, keys_out, pos_in, pos_out, nnz, 0, bits, g_stream);
                if (e != cudaSuccess) info = cuda_fail(err, e, "cub radix sort (transpose)");
            }
        }
        if (!info) {
            int blocks = (int)std::min<int64_t>((nnz + 255) / 256, (int64_t)g_num_sms * 16);
            note_launch("gather_i32");
            gather_i32_kernel<<<blocks, 256, 0, g_stream>>>(tw.idx, rowid, pos_out, nnz);
            info = gather_by_size(tw.val, A->csr.val, pos_out, nnz, es, err);
        }
        if (!info) {
            note_launch("lower_bound_ptr");
            lower_bound_ptr_kernel<<<(unsigned)((ncols + 1 + 255) / 256), 256, 0, g_stream>>>(keys_out, nnz, ncols, tw.ptr);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) info = cuda_fail(err, e, "transpose");
        }
        dev_free(rowid); dev_free(keys_out); dev_free(pos_in); dev_free(pos_out); dev_free(tmp);
    }
    if (info) { csr_free(tw); return info; }
    A->twin = tw;
    A->has_twin = true;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_cuda_Matrix_build_transpose(GrB_Matrix A) {
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    return matrix_ensure_twin(A);
}
