// hotcols.cu -- analysis behind the hot-column cache of the pull SpMV (spmv.cu).
//
// The gather x[colidx[k]] of a CSR SpMV costs one L1 wavefront per distinct 128-byte line a warp touches; on a
// power-law graph that, not HBM, bounds the kernel.  Most references, however, go to few columns: the analysis
// ranks the columns of one CSR by reference count, keeps the `max_hot` most referenced ones and writes a second
// index array in which a reference to a hot column is replaced by HOT_FLAG | rank.  The SpMV kernel keeps
// x[hot column] in shared memory (bank-parallel, no tag lookup) and reads the remapped array INSTEAD of colidx, so
// the HBM traffic per multiply is unchanged.  Cached per CSR like the merge-path tile table; dropped when the
// column order changes (lazy sort) or the matrix is released.
//
// Serves GrB_mxv / pull GrB_vxm (reference core/matrix.py:2252-2259, core/vector.py:1368-1375) when the same matrix
// is multiplied repeatedly (BFS / SSSP / PageRank iterations).
#include <cub/cub.cuh>
#include <stdlib.h>

#include <algorithm>

#include <vector>

#include "grb_internal.h"

constexpr int32_t HOT_FLAG = (int32_t)0x80000000;

__global__ void col_histogram_kernel(const int32_t *__restrict__ idx, int64_t nnz, unsigned int *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) atomicAdd(&cnt[idx[i]], 1u);
}
__global__ void iota32_kernel(int32_t *__restrict__ p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = (int32_t)i;
}
__global__ void rank_scatter_kernel(const int32_t *__restrict__ sorted_cols, int n_hot, int32_t *__restrict__ rank_of) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_hot) rank_of[sorted_cols[r]] = r;
}
__global__ void remap_kernel(const int32_t *__restrict__ idx, int64_t nnz, const int32_t *__restrict__ rank_of,
                             int32_t *__restrict__ remap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        const int32_t c = idx[i];
        const int32_t r = rank_of[c];
        remap[i] = r >= 0 ? (HOT_FLAG | r) : c;
    }
}

GrB_Info csr_ensure_hot(CsrArrays &c, int64_t ncols, int64_t nvals, int max_hot, std::string *err) {
    if (c.hot_state != 0) return GrB_SUCCESS;
    c.hot_state = -1;
    if (nvals <= 0 || ncols <= 0 || !c.idx) return GrB_SUCCESS;
    const int n_top = (int)std::min<int64_t>(max_hot, ncols);
    unsigned int *cnt = dev_alloc_t<unsigned int>((size_t)ncols), *cnt_sorted = dev_alloc_t<unsigned int>((size_t)ncols);
    int32_t *cols = dev_alloc_t<int32_t>((size_t)ncols), *cols_sorted = dev_alloc_t<int32_t>((size_t)ncols);
    int32_t *rank_of = dev_alloc_t<int32_t>((size_t)ncols);
    void *tmp = nullptr;
    GrB_Info info = GrB_SUCCESS;
    std::vector<unsigned int> top((size_t)n_top);
    if (!cnt || !cnt_sorted || !cols || !cols_sorted || !rank_of) info = set_error(err, GrB_OUT_OF_MEMORY, "hot-column analysis");
    if (!info) {
        cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * (size_t)ncols, g_stream);
        cudaMemsetAsync(rank_of, 0xff, sizeof(int32_t) * (size_t)ncols, g_stream);   // -1
        const int blocks = (int)std::min<int64_t>((nvals + 255) / 256, (int64_t)g_num_sms * 16);
        const int cblocks = (int)std::min<int64_t>((ncols + 255) / 256, (int64_t)g_num_sms * 16);
        note_launch("hot_histogram");
        col_histogram_kernel<<<blocks, 256, 0, g_stream>>>(c.idx, nvals, cnt);
        note_launch("hot_iota");
        iota32_kernel<<<cblocks, 256, 0, g_stream>>>(cols, ncols);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, cnt, cnt_sorted, cols, cols_sorted, ncols, 0, 32, g_stream);
        tmp = dev_alloc(tb);
        if (!tmp) info = set_error(err, GrB_OUT_OF_MEMORY, "hot-column sort scratch");
        if (!info) {
            note_launch("hot_sort");
            cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(tmp, tb, cnt, cnt_sorted, cols, cols_sorted, ncols, 0, 32, g_stream);
            if (e != cudaSuccess) info = cuda_fail(err, e, "hot-column sort");
        }
    }
    if (!info) {
        cudaMemcpyAsync(top.data(), cnt_sorted, sizeof(unsigned int) * (size_t)n_top, cudaMemcpyDeviceToHost, g_stream);
        cudaStreamSynchronize(g_stream);
        int n_hot = 0;
        while (n_hot < n_top && top[(size_t)n_hot] >= 2) n_hot++;   // a column referenced once gains nothing from the cache
        if (n_hot > 0) {
            c.hot_prefix = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_hot);
            c.hot_cols = dev_alloc_t<int32_t>((size_t)n_hot);
            c.hot_remap = dev_alloc_t<int32_t>((size_t)nvals);
            if (!c.hot_prefix || !c.hot_cols || !c.hot_remap) {
                csr_drop_hot(c);
                c.hot_state = -1;
                info = set_error(err, GrB_OUT_OF_MEMORY, "hot-column remap");
            }
        }
        if (!info && n_hot > 0) {
            int64_t run = 0;
            for (int r = 0; r < n_hot; r++) { run += top[(size_t)r]; c.hot_prefix[r] = run; }
            cudaMemcpyAsync(c.hot_cols, cols_sorted, sizeof(int32_t) * (size_t)n_hot, cudaMemcpyDeviceToDevice, g_stream);
            note_launch("hot_rank_scatter");
            rank_scatter_kernel<<<(n_hot + 255) / 256, 256, 0, g_stream>>>(cols_sorted, n_hot, rank_of);
            const int blocks = (int)std::min<int64_t>((nvals + 255) / 256, (int64_t)g_num_sms * 16);
            note_launch("hot_remap");
            remap_kernel<<<blocks, 256, 0, g_stream>>>(c.idx, nvals, rank_of, c.hot_remap);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                csr_drop_hot(c);
                c.hot_state = -1;
                info = cuda_fail(err, e, "hot-column remap");
            } else {
                c.hot_n = n_hot;
                c.hot_state = 1;
            }
        }
    }
    dev_free(cnt); dev_free(cnt_sorted); dev_free(cols); dev_free(cols_sorted); dev_free(rank_of); dev_free(tmp);
    return info;
}
