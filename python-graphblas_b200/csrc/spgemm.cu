// spgemm.cu -- T = A (+).(x) B for GrB_mxm: two-phase, flop-binned hash SpGEMM (row-wise Gustavson).
//
//   1. row_flops : flops(i) = sum_{k in A(i,:)} nnz(B(k,:))                       (upper bound on nnz(T(i,:)))
//   2. symbolic  : rows binned by flops; per bin a hash-set kernel counts the distinct columns of the row
//                  (warp-per-row / CTA-per-row with the table in shared memory, global-memory table for the
//                  few rows whose bound exceeds the largest shared table)
//   3. scan      : row_nnz -> row pointers of T, exact allocation
//   4. numeric   : rows re-binned by exact nnz; same kernels with a value array next to the keys, semiring
//                  multiply + atomic monoid combine in shared memory, then compaction into T (unsorted within
//                  a row: T is marked "jumbled" and sorted lazily, exactly as the reference's C library
//                  allows -- graphblas/core/matrix.py:1631-1644)
//
// Within a row, sub-groups of 8 lanes take one A(i,k) each and stride the B(k,:) row, so B is read with
// contiguous 32-byte segments.  Table sizes are run-time (dynamic shared memory), so one kernel
// instantiation per (semiring, type) serves every bin.
//
// Serves GrB_mxm: reference graphblas/core/matrix.py:2319-2328 (call assembled at core/base.py:496-503).
#include <cub/cub.cuh>
#include <vector>

#include "grb_ops.cuh"

constexpr int HASH_EMPTY = -1;
constexpr int LPE = 8;   // lanes per A entry

__device__ __forceinline__ unsigned hash_slot(int key, int shift) { return ((unsigned)key * 0x9E3779B1u) >> shift; }

// ------------------------------------------------------------------ flops per row
__global__ void row_flops_kernel(int64_t nrows, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj,
                                 const int64_t *__restrict__ Bp, int64_t *__restrict__ flops) {
    // 8 lanes per row, 4 rows per warp per step; the whole warp runs the same trip count (shuffles inside)
    const int lane = threadIdx.x & 7, sub = (threadIdx.x & 31) >> 3;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 4; base < nrows; base += nwarps * 4) {
        const int64_t i = base + sub;
        int64_t f = 0;
        if (i < nrows)
            for (int64_t k = Ap[i] + lane; k < Ap[i + 1]; k += 8) {
                int32_t r = Aj[k];
                f += Bp[r + 1] - Bp[r];
            }
        f += __shfl_down_sync(0xffffffffu, f, 4, 8);
        f += __shfl_down_sync(0xffffffffu, f, 2, 8);
        f += __shfl_down_sync(0xffffffffu, f, 1, 8);
        if (lane == 0 && i < nrows) flops[i] = f;
    }
}

// ------------------------------------------------------------------ binning
constexpr int NBINS = 10;
// bin b holds rows with count in (kBinMax[b-1], kBinMax[b]]; bin 0 = empty rows; bin 9 = global-table rows
__constant__ int64_t c_bin_max[NBINS] = {0, 32, 128, 256, 512, 1024, 2048, 4096, 8192, INT64_MAX};
static const int h_bin_table[NBINS] = {0, 64, 256, 512, 1024, 2048, 4096, 8192, 16384, 0};
static const int h_bin_threads[NBINS] = {0, 256, 256, 64, 128, 128, 256, 256, 512, 512};

__device__ __forceinline__ int bin_of(int64_t c) {
    int b = 0;
#pragma unroll
    for (int q = 0; q < NBINS - 1; q++) b += (c > c_bin_max[q]);
    return b;
}
__global__ void bin_count_kernel(int64_t nrows, const int64_t *__restrict__ cnt, unsigned long long *__restrict__ bin_counts) {
    __shared__ unsigned int s[NBINS];
    if (threadIdx.x < NBINS) s[threadIdx.x] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += stride) atomicAdd(&s[bin_of(cnt[i])], 1u);
    __syncthreads();
    if (threadIdx.x < NBINS && s[threadIdx.x]) atomicAdd(&bin_counts[threadIdx.x], (unsigned long long)s[threadIdx.x]);
}
__global__ void bin_fill_kernel(int64_t nrows, const int64_t *__restrict__ cnt, unsigned long long *__restrict__ cursors,
                                int32_t *__restrict__ bin_rows) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += stride) {
        int b = bin_of(cnt[i]);
        if (b == 0) continue;
        unsigned long long pos = atomicAdd(&cursors[b], 1ull);
        bin_rows[pos] = (int32_t)i;
    }
}

// ------------------------------------------------------------------ the hash kernels
// One "group" (a warp when WARP_ROWS, else the whole CTA) owns one row and one hash table.
template <typename SR, typename T, bool NUMERIC, bool WARP_ROWS>
__global__ void spgemm_hash_kernel(SR sr, const int32_t *__restrict__ rows, int64_t n_rows, int table_size, int shift,
                                   const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj, const T *__restrict__ Ax,
                                   const int64_t *__restrict__ Bp, const int32_t *__restrict__ Bj, const T *__restrict__ Bx,
                                   int64_t *__restrict__ row_nnz,       // symbolic output
                                   const int64_t *__restrict__ Cp, int32_t *__restrict__ Cj, T *__restrict__ Cx,  // numeric output
                                   int *g_keys, T *g_vals, const int64_t *__restrict__ g_offsets) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int group_threads = WARP_ROWS ? 32 : blockDim.x;
    const int gtid = WARP_ROWS ? (threadIdx.x & 31) : threadIdx.x;
    const int groups_per_block = WARP_ROWS ? (blockDim.x >> 5) : 1;
    const int group_in_block = WARP_ROWS ? (threadIdx.x >> 5) : 0;
    const int64_t g = (int64_t)blockIdx.x * groups_per_block + group_in_block;
    const bool active = g < n_rows;   // whole group shares this predicate
    __shared__ int s_count_blk;

    int *keys;
    T *vals = nullptr;
    int *count;
    int tsize = table_size, tshift = shift;
    if (g_keys) {   // global-memory table for this row
        int64_t off = active ? g_offsets[g] : 0;
        int64_t sz = active ? g_offsets[g + 1] - off : 2;
        keys = g_keys + off;
        if (NUMERIC) vals = g_vals + off;
        tsize = (int)sz;
        tshift = 32 - (31 - __clz((unsigned)tsize));
        count = &s_count_blk;
    } else if (WARP_ROWS) {
        const size_t per = (size_t)table_size * (NUMERIC ? (4 + sizeof(T)) : 4) + 8;
        unsigned char *base = s_raw + ((per + 7) & ~(size_t)7) * group_in_block;
        count = reinterpret_cast<int *>(base);
        keys = reinterpret_cast<int *>(base + 8);
        if (NUMERIC) vals = reinterpret_cast<T *>(base + 8 + (size_t)table_size * 4);
    } else {
        count = &s_count_blk;
        keys = reinterpret_cast<int *>(s_raw);
        if (NUMERIC) vals = reinterpret_cast<T *>(s_raw + (size_t)table_size * 4);
    }
    const int mask = tsize - 1;

    // init
    if (active) {
        for (int t = gtid; t < tsize; t += group_threads) {
            keys[t] = HASH_EMPTY;
            if (NUMERIC) vals[t] = sr.identity();
        }
    }
    if (gtid == 0) *count = 0;
    if (WARP_ROWS) __syncwarp(); else __syncthreads();

    int local_new = 0;
    if (active) {
        const int64_t row = rows[g];
        const int sub = gtid / LPE, lane = gtid % LPE, nsub = group_threads / LPE;
        const int64_t a_beg = Ap[row], a_end = Ap[row + 1];
        for (int64_t k = a_beg + sub; k < a_end; k += nsub) {
            const int32_t br = Aj[k];
            T a = one_of<T>();
            if (NUMERIC && sr.reads_a()) a = Ax[k];
            const int64_t b_beg = Bp[br], b_end = Bp[br + 1];
            for (int64_t q = b_beg + lane; q < b_end; q += LPE) {
                const int j = Bj[q];
                T p = T();
                if (NUMERIC) {
                    T b = sr.reads_b() ? Bx[q] : one_of<T>();
                    p = sr.mul(a, b);
                }
                unsigned h = hash_slot(j, tshift) & (unsigned)mask;
                while (true) {
                    int cur = keys[h];
                    if (cur == HASH_EMPTY) {
                        cur = atomicCAS(&keys[h], HASH_EMPTY, j);
                        if (cur == HASH_EMPTY) { local_new++; cur = j; }
                    }
                    if (cur == j) {
                        if (NUMERIC) atomic_combine(sr, &vals[h], p);
                        break;
                    }
                    h = (h + 1) & (unsigned)mask;
                }
            }
        }
    }

    if (!NUMERIC) {
        // row_nnz = number of successful first inserts
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_new += __shfl_down_sync(0xffffffffu, local_new, o);
        if (WARP_ROWS) {
            if (gtid == 0 && active) row_nnz[rows[g]] = local_new;
        } else {
            if ((threadIdx.x & 31) == 0 && local_new) atomicAdd(count, local_new);
            __syncthreads();
            if (threadIdx.x == 0 && active) row_nnz[rows[g]] = *count;
        }
        return;
    } else {
        if (WARP_ROWS) __syncwarp(); else __syncthreads();
        if (active) {
            const int64_t row = rows[g];
            const int64_t base = Cp[row];
            for (int t = gtid; t < tsize; t += group_threads) {
                int key = keys[t];
                if (key != HASH_EMPTY) {
                    int pos = atomicAdd(count, 1);
                    Cj[base + pos] = key;
                    Cx[base + pos] = vals[t];
                }
            }
        }
    }
}

__global__ void gtable_sizes_kernel(const int32_t *__restrict__ rows, int64_t n, const int64_t *__restrict__ cnt,
                                    int64_t *__restrict__ sizes) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c = cnt[rows[i]] * 2;
    int64_t s = 1024;
    while (s < c) s <<= 1;
    sizes[i] = s;
}
__global__ void i64_copy_kernel(int64_t *dst, const int64_t *src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = src[i];
}
__global__ void reduce_sum_max_kernel(const int64_t *__restrict__ v, int64_t n, unsigned long long *__restrict__ out) {
    unsigned long long sum = 0, mx = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) { unsigned long long x = (unsigned long long)v[i]; sum += x; mx = x > mx ? x : mx; }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        unsigned long long om = __shfl_down_sync(0xffffffffu, mx, o);
        mx = om > mx ? om : mx;
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], sum); atomicMax(&out[1], mx); }
}

// ------------------------------------------------------------------ host orchestration
struct Bins {
    int32_t *rows = nullptr;           // row ids grouped by bin
    unsigned long long count[NBINS];   // rows per bin
    unsigned long long start[NBINS];   // offset of each bin inside rows[]
};

static GrB_Info make_bins(Bins *bins, int64_t nrows, const int64_t *cnt, std::string *err) {
    unsigned long long *d = dev_alloc_t<unsigned long long>(2 * NBINS);
    bins->rows = dev_alloc_t<int32_t>((size_t)(nrows > 0 ? nrows : 1));
    if (!d || !bins->rows) { dev_free(d); dev_free(bins->rows); bins->rows = nullptr; return set_error(err, GrB_OUT_OF_MEMORY, "spgemm bins"); }
    cudaMemsetAsync(d, 0, sizeof(unsigned long long) * 2 * NBINS, g_stream);
    int blocks = (int)std::min<int64_t>((nrows + 255) / 256 + 1, (int64_t)g_num_sms * 8);
    {
        LAUNCH_NOTE("spgemm_bin_count");
        bin_count_kernel<<<blocks, 256, 0, g_stream>>>(nrows, cnt, d);
    }
    cudaMemcpyAsync(bins->count, d, sizeof(unsigned long long) * NBINS, cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);
    unsigned long long off = 0, cursors[NBINS];
    for (int b = 0; b < NBINS; b++) {
        bins->start[b] = off;
        cursors[b] = off;
        if (b > 0) off += bins->count[b];
    }
    cudaMemcpyAsync(d + NBINS, cursors, sizeof(cursors), cudaMemcpyHostToDevice, g_stream);
    {
        LAUNCH_NOTE("spgemm_bin_fill");
        bin_fill_kernel<<<blocks, 256, 0, g_stream>>>(nrows, cnt, d + NBINS, bins->rows);
    }
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(g_stream);   // `cursors` is a host stack array read by the async copy
    dev_free(d);
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

struct HashArgs {
    const int64_t *Ap; const int32_t *Aj; const void *Ax;
    const int64_t *Bp; const int32_t *Bj; const void *Bx;
    int64_t *row_nnz; const int64_t *Cp; int32_t *Cj; void *Cx;
    const int64_t *cnt;   // per-row bound (flops or nnz) used for global table sizing
};

template <typename SR, typename T, bool NUMERIC>
static GrB_Info run_bins(const SR &sr, const Bins &bins, const HashArgs &a, std::string *err) {
    const size_t entry = NUMERIC ? 4 + sizeof(T) : 4;
    for (int b = 1; b < NBINS; b++) {
        const int64_t n = (int64_t)bins.count[b];
        if (n == 0) continue;
        const int32_t *rows = bins.rows + bins.start[b];
        if (b < NBINS - 1) {
            const int table = h_bin_table[b], threads = h_bin_threads[b];
            int shift = 32;
            for (int s = table; s > 1; s >>= 1) shift--;
            const bool warp_rows = (b <= 2);
            if (warp_rows) {
                const int rpb = threads / 32;
                size_t per = (((size_t)table * entry + 8) + 7) & ~(size_t)7;
                size_t smem = per * rpb;
                auto kern = spgemm_hash_kernel<SR, T, NUMERIC, true>;
                if (smem > 40 * 1024) CUDA_TRY(err, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_warp" : "spgemm_symbolic_warp");
                kern<<<(unsigned)((n + rpb - 1) / rpb), threads, smem, g_stream>>>(sr, rows, n, table, shift, a.Ap, a.Aj, (const T *)a.Ax, a.Bp, a.Bj, (const T *)a.Bx, a.row_nnz, a.Cp, a.Cj, (T *)a.Cx, nullptr, nullptr, nullptr);
            } else {
                size_t smem = (size_t)table * entry;
                auto kern = spgemm_hash_kernel<SR, T, NUMERIC, false>;
                if (smem > 40 * 1024) CUDA_TRY(err, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_block" : "spgemm_symbolic_block");
                kern<<<(unsigned)n, threads, smem, g_stream>>>(sr, rows, n, table, shift, a.Ap, a.Aj, (const T *)a.Ax, a.Bp, a.Bj, (const T *)a.Bx, a.row_nnz, a.Cp, a.Cj, (T *)a.Cx, nullptr, nullptr, nullptr);
            }
            {
                cudaError_t le = cudaGetLastError();
                if (le != cudaSuccess)
                    return set_error(err, GrB_PANIC, "spgemm %s kernel launch failed in bin %d (table %d, %d threads, %lld rows, entry %zu B): %s",
                                     NUMERIC ? "numeric" : "symbolic", b, table, threads, (long long)n, entry, cudaGetErrorString(le));
            }
        } else {
            // rows whose bound exceeds the largest shared table: global-memory tables, in batches that fit a budget
            const int64_t budget_entries = (int64_t)opt_get_int("spgemm_gtable_entries", (long)1 << 30);
            int64_t *sizes = dev_alloc_t<int64_t>((size_t)n + 1);
            if (!sizes) return set_error(err, GrB_OUT_OF_MEMORY, "global table sizes");
            note_launch("gtable_sizes");
            gtable_sizes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(rows, n, a.cnt, sizes);
            std::vector<int64_t> hs((size_t)n + 1);
            cudaMemcpyAsync(hs.data(), sizes, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, g_stream);
            cudaStreamSynchronize(g_stream);
            int64_t i0 = 0;
            GrB_Info info = GrB_SUCCESS;
            while (i0 < n && !info) {
                int64_t i1 = i0, tot = 0;
                while (i1 < n && (i1 == i0 || tot + hs[(size_t)i1] <= budget_entries)) tot += hs[(size_t)i1++];
                std::vector<int64_t> offs((size_t)(i1 - i0) + 1);
                offs[0] = 0;
                for (int64_t q = i0; q < i1; q++) offs[(size_t)(q - i0) + 1] = offs[(size_t)(q - i0)] + hs[(size_t)q];
                int64_t *doffs = dev_alloc_t<int64_t>(offs.size());
                int *gk = dev_alloc_t<int>((size_t)tot);
                T *gv = NUMERIC ? dev_alloc_t<T>((size_t)tot) : nullptr;
                if (!doffs || !gk || (NUMERIC && !gv)) info = set_error(err, GrB_OUT_OF_MEMORY, "global hash tables (%lld entries)", (long long)tot);
                if (!info) {
                    cudaMemcpyAsync(doffs, offs.data(), sizeof(int64_t) * offs.size(), cudaMemcpyHostToDevice, g_stream);
                    LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_global" : "spgemm_symbolic_global");
                    spgemm_hash_kernel<SR, T, NUMERIC, false><<<(unsigned)(i1 - i0), 512, 0, g_stream>>>(sr, rows + i0, i1 - i0, 2, 31, a.Ap, a.Aj, (const T *)a.Ax, a.Bp, a.Bj, (const T *)a.Bx, a.row_nnz, a.Cp, a.Cj, (T *)a.Cx, gk, gv, doffs);
                    cudaError_t e = cudaGetLastError();
                    cudaStreamSynchronize(g_stream);   // offs is host memory
                    if (e != cudaSuccess) info = cuda_fail(err, e, "spgemm global-table kernel");
                }
                dev_free(doffs); dev_free(gk); dev_free(gv);
                i0 = i1;
            }
            dev_free(sizes);
            GRB_TRY(info);
        }
    }
    return GrB_SUCCESS;
}

struct SpgemmPlan {
    CsrArrays *A, *B;
    int64_t m, k, n, annz, bnnz;
    int a_type, b_type;
};

template <typename T>
static GrB_Info spgemm_numeric_typed(GrB_Matrix Tm, const GrB_Semiring op, const SpgemmPlan &p, const Bins &bins,
                                     const int64_t *row_nnz, std::string *err) {
    GrB_Info info = GrB_SUCCESS;
    const int T_code = type_code_of<T>();
    GRB_DISPATCH_SEMIRING(op->add, op->mul, T, SRT, sr, {
        const void *ax = nullptr, *bx = nullptr;
        void *atmp = nullptr, *btmp = nullptr;
        if (sr.reads_a()) info = cast_view(&ax, &atmp, p.A->val, p.a_type, T_code, p.annz, err);
        if (!info && sr.reads_b()) info = cast_view(&bx, &btmp, p.B->val, p.b_type, T_code, p.bnnz, err);
        if (!info) {
            HashArgs a{p.A->ptr, p.A->idx, ax, p.B->ptr, p.B->idx, bx, nullptr, Tm->csr.ptr, Tm->csr.idx, Tm->csr.val, row_nnz};
            info = run_bins<SRT, T, true>(sr, bins, a, err);
        }
        dev_free(atmp);
        dev_free(btmp);
    });
    return info;
}

GrB_Info spgemm(GrB_Matrix *Tout, const GrB_Semiring op, GrB_Matrix A, bool at, GrB_Matrix B, bool bt, const GrB_Matrix M,
                bool mask_comp, bool mask_struct, std::string *err, bool symbolic_only, uint64_t *flops_out,
                uint64_t *nvals_out) {
    (void)M; (void)mask_comp; (void)mask_struct;   // the write-back applies the mask; see DESIGN.md
    GRB_TRY(matrix_materialize(A));
    GRB_TRY(matrix_materialize(B));
    if (at) GRB_TRY(matrix_ensure_twin(A));
    if (bt) GRB_TRY(matrix_ensure_twin(B));
    SpgemmPlan p;
    p.A = at ? &A->twin : &A->csr;
    p.B = bt ? &B->twin : &B->csr;
    p.m = at ? A->ncols : A->nrows;
    p.k = at ? A->nrows : A->ncols;
    const int64_t bk = bt ? B->ncols : B->nrows;
    p.n = bt ? B->nrows : B->ncols;
    p.annz = A->nvals; p.bnnz = B->nvals;
    p.a_type = A->type; p.b_type = B->type;
    if (p.k != bk)
        return set_error(err, GrB_DIMENSION_MISMATCH, "mxm: inner dimensions differ (%lld vs %lld)", (long long)p.k, (long long)bk);
    const int D = op ? op->type : TC_INT64;

    int64_t *flops = dev_alloc_t<int64_t>((size_t)p.m + 1), *row_nnz = dev_alloc_t<int64_t>((size_t)p.m + 1);
    unsigned long long *red = dev_alloc_t<unsigned long long>(2);
    Bins sbins, nbins;
    GrB_Matrix Tm = nullptr;
    GrB_Info info = GrB_SUCCESS;
    if (!flops || !row_nnz || !red) info = set_error(err, GrB_OUT_OF_MEMORY, "spgemm row arrays");
    unsigned long long hred[2] = {0, 0};
    if (!info && p.m > 0) {
        cudaMemsetAsync(red, 0, 16, g_stream);
        cudaMemsetAsync(row_nnz, 0, sizeof(int64_t) * ((size_t)p.m + 1), g_stream);
        int blocks = (int)std::min<int64_t>((p.m + 31) / 32, (int64_t)g_num_sms * 32);
        {
            LAUNCH_NOTE("spgemm_row_flops");
            row_flops_kernel<<<blocks, 256, 0, g_stream>>>(p.m, p.A->ptr, p.A->idx, p.B->ptr, flops);
        }
        {
            LAUNCH_NOTE("reduce_sum_max");
            reduce_sum_max_kernel<<<std::min(blocks, g_num_sms * 4), 256, 0, g_stream>>>(flops, p.m, red);
        }
        cudaMemcpyAsync(hred, red, 16, cudaMemcpyDeviceToHost, g_stream);
        info = make_bins(&sbins, p.m, flops, err);
    }
    if (flops_out) *flops_out = hred[0];
    // ---- symbolic
    if (!info && p.m > 0) {
        HashArgs a{p.A->ptr, p.A->idx, nullptr, p.B->ptr, p.B->idx, nullptr, row_nnz, nullptr, nullptr, nullptr, flops};
        SRDyn<int32_t> dummy;
        dummy.a_op = OP_ANY; dummy.m_op = OP_PAIR;
        info = run_bins<SRDyn<int32_t>, int32_t, false>(dummy, sbins, a, err);
    }
    int64_t total = 0;
    if (!info) {
        GrB_Info i2 = matrix_new_shell(&Tm, D, p.m, p.n);
        if (i2) info = i2;
    }
    if (!info) {
        Tm->csr.ptr = dev_alloc_t<int64_t>((size_t)p.m + 1);
        if (!Tm->csr.ptr) info = set_error(err, GrB_OUT_OF_MEMORY, "spgemm row pointers");
    }
    if (!info) {
        note_launch("i64_copy");
        i64_copy_kernel<<<(unsigned)std::min<int64_t>((p.m + 256) / 256, (int64_t)g_num_sms * 8), 256, 0, g_stream>>>(Tm->csr.ptr, row_nnz, p.m + 1);
        info = exclusive_scan_i64(Tm->csr.ptr, p.m + 1, err);
        if (!info) total = read_i64(Tm->csr.ptr + p.m);
    }
    if (nvals_out) *nvals_out = (uint64_t)total;
    if (!info && !symbolic_only) {
        size_t nv = (size_t)(total > 0 ? total : 1);
        Tm->csr.idx = dev_alloc_t<int32_t>(nv);
        Tm->csr.val = dev_alloc(nv * type_size(D));
        Tm->nvals = total;
        Tm->jumbled = true;
        if (!Tm->csr.idx || !Tm->csr.val) info = set_error(err, GrB_OUT_OF_MEMORY, "mxm result needs %lld entries (%.1f GB)", (long long)total, (double)total * (4 + type_size(D)) / 1e9);
        if (!info && total > 0) info = make_bins(&nbins, p.m, row_nnz, err);
        if (!info && total > 0) {
            GrB_Info i3 = GrB_NOT_IMPLEMENTED;
            GRB_DISPATCH_TYPE(D, T, i3 = spgemm_numeric_typed<T>(Tm, op, p, nbins, row_nnz, err));
            info = i3;
        }
    }
    dev_free(flops); dev_free(row_nnz); dev_free(red); dev_free(sbins.rows); dev_free(nbins.rows);
    if (info || symbolic_only) {
        if (Tm) GrB_Matrix_free(&Tm);
        return info;
    }
    *Tout = Tm;
    return GrB_SUCCESS;
}
