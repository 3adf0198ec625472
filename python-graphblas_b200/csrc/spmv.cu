// spmv.cu -- t = M (+).(x) u for GrB_mxv / GrB_vxm, M = A or A'.
//
//  pull  (M rows available as CSR):  merge-path CSR SpMV.  One CTA consumes a fixed-size slice of the
//        merge of (row-end offsets, nonzero indices); products are gathered with coalesced loads into
//        shared memory, each thread then walks its own equal share of the merge path and emits finished
//        rows; partial rows are stitched with a warp-shuffle segmented (reduce-by-key) scan inside the
//        CTA and a tiny fix-up kernel across CTAs.  Load balance is independent of the degree skew
//        (R-MAT max degree 97k vs mean 16).  Algorithmic bytes: nnz*(4 + rho*s_val) + (nrows+1)*8 +
//        ncols*s_x + nrows*(s_y+1)   (SURVEY.md section 8d).
//  pull  (masked): warp-per-row kernel that skips rows the mask rules out and, for the ANY monoid,
//        stops at the first hit (bottom-up BFS step).
//  push  (only CSR of M' available, u sparse): frontier compaction + warp-per-frontier-vertex scatter with
//        atomic monoid combine; the mask is tested in-register before the atomic.
//
// Serves: GrB_mxv (reference core/matrix.py:2252-2259), GrB_vxm (core/vector.py:1368-1375).
#include <limits.h>

#include "spmv_common.cuh"

constexpr int SPMV_BLOCK = 256;

// items per thread: odd, so that the per-thread walk over shared memory (stride IPT words) is bank-conflict free
#ifndef SPMV_IPT4
#define SPMV_IPT4 7
#endif
#ifndef SPMV_IPT8
#define SPMV_IPT8 5
#endif
template <typename T> struct SpmvCfg { static constexpr int IPT = (sizeof(T) >= 8 ? SPMV_IPT8 : SPMV_IPT4); };

// (row, value, has) triple of a partially reduced row; combine = reduce-by-key, associative
template <typename T> struct Carry { int row; int has; T val; };
template <typename SR, typename T>
__device__ __forceinline__ Carry<T> carry_combine(const SR &sr, const Carry<T> &a, const Carry<T> &b) {
    Carry<T> r = b;
    if (a.row == b.row && a.has) {
        r.val = b.has ? sr.add(a.val, b.val) : a.val;
        r.has = 1;
    }
    return r;
}


// ------------------------------------------------------------------ merge-path tile search
__global__ void merge_search_kernel(const int64_t *__restrict__ rowptr, int64_t nrows, int64_t nnz, int tile_items,
                                    int64_t n_tiles, int64_t *__restrict__ tile_starts) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    int64_t diag = t * (int64_t)tile_items;
    if (diag > nrows + nnz) diag = nrows + nnz;
    int64_t lo = diag > nnz ? diag - nnz : 0;
    int64_t hi = diag < nrows ? diag : nrows;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid + 1] <= diag - mid - 1) lo = mid + 1;
        else hi = mid;
    }
    tile_starts[t] = lo;
}

// ------------------------------------------------------------------ merge-path SpMV
template <typename SR, typename T, bool XFULL>
__global__ void __launch_bounds__(SPMV_BLOCK, (sizeof(T) >= 8 ? 6 : 8))
spmv_merge_kernel(SR sr, int64_t nrows, int64_t nnz, const int64_t *__restrict__ rowptr,
                  const int32_t *__restrict__ colidx, const T *__restrict__ avals, const T *__restrict__ x,
                  const uint8_t *__restrict__ xp, const int64_t *__restrict__ tile_starts, bool flip,
                  T *__restrict__ t_vals, uint8_t *__restrict__ t_present, int64_t *__restrict__ carry_row,
                  T *__restrict__ carry_val, uint8_t *__restrict__ carry_has, VecEpi<T> epi) {
    constexpr int IPT = SpmvCfg<T>::IPT;
    constexpr int TILE = SPMV_BLOCK * IPT;
    __shared__ int s_rowend[TILE + 1];
    __shared__ T s_prod[TILE];
    __shared__ uint8_t s_has[XFULL ? 1 : TILE];
    __shared__ Carry<T> s_warp[SPMV_BLOCK / 32];

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t total = nrows + nnz;
    const int64_t diag0 = tile * (int64_t)TILE;
    const int64_t diag1 = (diag0 + TILE < total) ? diag0 + TILE : total;
    const int64_t r0 = tile_starts[tile], r1 = tile_starts[tile + 1];
    const int64_t k0 = diag0 - r0, k1 = diag1 - r1;
    const int tile_rows = (int)(r1 - r0), tile_nnz = (int)(k1 - k0);
    const int tile_items = tile_rows + tile_nnz;

    const bool carry_in = rowptr[r0 < nrows ? r0 : nrows] < k0;   // the tile's first row began in an earlier tile
    for (int i = tid; i < tile_rows; i += SPMV_BLOCK) s_rowend[i] = (int)(rowptr[r0 + i + 1] - k0);
    if (tid == 0) s_rowend[tile_rows] = INT_MAX;

    // coalesced gather phase: all loads of the unrolled batch are issued before any use
    {
        int32_t c[IPT];
        T a[IPT];
#pragma unroll
        for (int it = 0; it < IPT; it++) {
            int idx = tid + it * SPMV_BLOCK;
            c[it] = 0;
            if (idx < tile_nnz) {
                c[it] = colidx[k0 + idx];
                if (sr.reads_a()) a[it] = avals[k0 + idx];
            }
        }
#pragma unroll
        for (int it = 0; it < IPT; it++) {
            int idx = tid + it * SPMV_BLOCK;
            if (idx < tile_nnz) {
                T av = sr.reads_a() ? a[it] : one_of<T>();
                T xv = sr.reads_b() ? x[c[it]] : one_of<T>();
                s_prod[idx] = flip ? sr.mul(xv, av) : sr.mul(av, xv);
                if (!XFULL) s_has[idx] = xp[c[it]];
            }
        }
    }
    __syncthreads();

    // each thread's share of the merge path
    int d = tid * IPT;
    if (d > tile_items) d = tile_items;
    int lo = d > tile_nnz ? d - tile_nnz : 0;
    int hi = d < tile_rows ? d : tile_rows;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s_rowend[mid] <= d - mid - 1) lo = mid + 1;
        else hi = mid;
    }
    int xr = lo, yk = d - lo;
    const int x_first = xr;
    T emit_val[IPT];          // indexed by the (static) step at which the row ended
    unsigned emit_mask = 0, emit_has = 0;
    T acc = sr.identity();
    int has = 0;
    int row_end = s_rowend[xr];   // re-read only when the row advances
#pragma unroll
    for (int it = 0; it < IPT; it++) {
        emit_val[it] = acc;
        if (d + it < tile_items) {
            if (yk < row_end) {
                bool ph = XFULL ? true : (s_has[yk] != 0);
                if (ph) {
                    acc = has ? sr.add(acc, s_prod[yk]) : s_prod[yk];
                    has = 1;
                }
                yk++;
            } else {
                emit_val[it] = acc;
                emit_mask |= 1u << it;
                emit_has |= (unsigned)has << it;
                xr++;
                row_end = s_rowend[xr];
                acc = sr.identity();
                has = 0;
            }
        }
    }

    // segmented inclusive scan of the per-thread leftovers (key = row)
    Carry<T> cur;
    cur.row = xr;
    cur.has = has;
    cur.val = acc;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Carry<T> up;
        up.row = __shfl_up_sync(0xffffffffu, cur.row, o);
        up.has = __shfl_up_sync(0xffffffffu, cur.has, o);
        up.val = shfl_up_any(cur.val, o);
        if (lane >= o) cur = carry_combine(sr, up, cur);
    }
    if (lane == 31) s_warp[warp] = cur;
    __syncthreads();
    if (warp == 0) {
        Carry<T> w;
        if (lane < SPMV_BLOCK / 32) w = s_warp[lane];
        else { w.row = -1; w.has = 0; w.val = sr.identity(); }
#pragma unroll
        for (int o = 1; o < SPMV_BLOCK / 32; o <<= 1) {
            Carry<T> up;
            up.row = __shfl_up_sync(0xffffffffu, w.row, o);
            up.has = __shfl_up_sync(0xffffffffu, w.has, o);
            up.val = shfl_up_any(w.val, o);
            if (lane >= o) w = carry_combine(sr, up, w);
        }
        if (lane < SPMV_BLOCK / 32) s_warp[lane] = w;
    }
    __syncthreads();
    // exclusive prefix for this thread
    Carry<T> prev;
    prev.row = __shfl_up_sync(0xffffffffu, cur.row, 1);
    prev.has = __shfl_up_sync(0xffffffffu, cur.has, 1);
    prev.val = shfl_up_any(cur.val, 1);
    if (lane == 0) { prev.row = -1; prev.has = 0; prev.val = sr.identity(); }
    if (warp > 0) {
        Carry<T> wp = s_warp[warp - 1];
        prev = (lane == 0) ? wp : carry_combine(sr, wp, prev);
    }
    // write finished rows; the first one may continue a row begun by earlier threads of this tile
#pragma unroll
    for (int it = 0; it < IPT; it++) {
        if (emit_mask & (1u << it)) {
            const int e = __popc(emit_mask & ((1u << it) - 1u));
            T v = emit_val[it];
            int h = (emit_has >> it) & 1;
            if (e == 0 && prev.row == x_first && prev.has) {
                v = h ? sr.add(prev.val, v) : prev.val;
                h = 1;
            }
            const int64_t row = r0 + x_first + e;
            if (e == 0 && x_first == 0 && carry_in) {   // earlier tiles hold part of this row: the fix-up kernel finishes it
                t_vals[row] = h ? v : T();
                t_present[row] = (uint8_t)h;
            } else {
                epi_write<T, false>(epi, row, v, h, t_vals, t_present);
            }
        }
    }
    // Fused exchange: the rows that ended in this tile form ONE contiguous run of the output; the CTA pushes that run
    // into the next input vector of every rank with coalesced stores (consecutive lanes -> consecutive positions: whole 128-byte
    // NVLink packets) instead of one scattered store per row and peer from the emission loop above -- measured 6.1 -> see
    // DESIGN.md section 5.  A row that began in an earlier tile is finished (and pushed) by the fix-up kernel.
    if (epi.active && epi.npeer) {
        __syncthreads();   // this CTA's own global writes above are visible to all of its threads
        for (int i = (carry_in ? 1 : 0) + tid; i < tile_rows; i += SPMV_BLOCK) {
            const int64_t row = r0 + i;
            const uint8_t p = t_present[row];
            T out = p ? t_vals[row] : T();
            if (p && epi.pscale) out = binop<T>(OP_TIMES, out, epi.pscale[row]);
#pragma unroll 1
            for (int k = 0; k < epi.npeer; k++) {
                epi.pv[k][epi.poff + row] = out;
                if (epi.pp[k]) epi.pp[k][epi.poff + row] = p;
            }
        }
    }
    // tile carry-out: partial sum of the row that continues into the next tile
    if (tid == SPMV_BLOCK - 1) {
        Carry<T> tot = (warp > 0) ? carry_combine(sr, s_warp[warp - 1], cur) : cur;
        // `cur` is warp-inclusive; combining with the previous warps' total gives the block-inclusive value
        carry_row[tile] = r0 + tot.row;
        carry_val[tile] = tot.val;
        carry_has[tile] = (uint8_t)tot.has;
    }
}

// cross-tile fix-up: the last tile that carries into a row folds the whole chain of carries into the partial left by
// the tile where the row ends, then applies the write-back (that tile deferred it -- same predicate on both sides:
// the row began before the first nonzero of the tile in which it ends).
template <typename SR, typename T>
__global__ void spmv_merge_fixup_kernel(SR sr, int64_t n_tiles, int64_t nrows, int tile_items, const int64_t *__restrict__ rowptr,
                                        const int64_t *__restrict__ tile_starts, const int64_t *__restrict__ carry_row,
                                        const T *__restrict__ carry_val, const uint8_t *__restrict__ carry_has,
                                        T *__restrict__ t_vals, uint8_t *__restrict__ t_present, VecEpi<T> epi) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t row = carry_row[t];
    if (row >= nrows) return;
    if (t + 1 >= n_tiles) return;                      // cannot happen for row < nrows (the last tile consumes every row end)
    if (carry_row[t + 1] == row) return;               // a later tile owns this chain
    const int64_t k0_next = (t + 1) * (int64_t)tile_items - tile_starts[t + 1];
    if (rowptr[row] >= k0_next) return;                // the row starts in the next tile: it was emitted there, complete
    T acc = T();
    int has = 0;
    for (int64_t q = t; q >= 0 && carry_row[q] == row; q--) {
        if (carry_has[q]) {
            acc = has ? sr.add(carry_val[q], acc) : carry_val[q];
            has = 1;
        }
    }
    if (t_present[row]) {
        acc = has ? sr.add(acc, t_vals[row]) : t_vals[row];
        has = 1;
    }
    epi_write(epi, row, acc, has, t_vals, t_present);
}

// ------------------------------------------------------------------ warp-per-row pull with mask skip / ANY early exit
template <typename SR, typename T, bool XFULL>
__global__ void __launch_bounds__(256)
spmv_rowwarp_kernel(SR sr, int64_t nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                    const T *__restrict__ avals, const T *__restrict__ x, const uint8_t *__restrict__ xp, bool flip,
                    const uint8_t *__restrict__ mask, bool mask_comp, T *__restrict__ t_vals,
                    uint8_t *__restrict__ t_present, VecEpi<T> epi) {
    // A warp takes 32 consecutive rows: every lane first filters ITS row (mask, row length) and writes the empty result of a
    // masked-out / empty row itself -- coalesced, and without a dependent load chain per skipped row (in a bottom-up BFS step
    // most rows are skipped) -- then the warp walks the rows that are left, one at a time, all lanes on one row.
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = w * 32; base < nrows; base += nw * 32) {
        const int64_t mine = base + lane;
        long long b = 0, e = 0;
        bool need = false;
        if (mine < nrows) {
            const bool m = mask ? ((mask[mine] != 0) != mask_comp) : true;
            if (m) {
                b = rowptr[mine];
                e = rowptr[mine + 1];
                need = e > b;
            }
            if (!need) epi_write(epi, mine, T(), 0, t_vals, t_present);   // masked out (T(i) irrelevant, row never read) or no entries
        }
        unsigned todo = __ballot_sync(FULL, need);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t rb = __shfl_sync(FULL, b, src), re = __shfl_sync(FULL, e, src);
            T acc = sr.identity();
            int has = 0;
            for (int64_t k = rb + lane; k < re; k += 32) {
                int32_t c = colidx[k];
                bool ph = XFULL ? true : (xp[c] != 0);
                if (ph) {
                    T av = sr.reads_a() ? avals[k] : one_of<T>();
                    T xv = sr.reads_b() ? x[c] : one_of<T>();
                    T p = flip ? sr.mul(xv, av) : sr.mul(av, xv);
                    acc = has ? sr.add(acc, p) : p;
                    has = 1;
                }
                if (SR::kAddIsAny && __any_sync(__activemask(), has)) break;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                T ov = shfl_down_any(acc, o);
                int oh = __shfl_down_sync(FULL, has, o);
                if (oh) {
                    acc = has ? sr.add(acc, ov) : ov;
                    has = 1;
                }
            }
            if (lane == 0) epi_write(epi, base + src, acc, has, t_vals, t_present);
        }
    }
}

// ------------------------------------------------------------------ push (SpMSpV)
__global__ void compact_present_kernel(const uint8_t *__restrict__ present, int64_t n, int32_t *__restrict__ out,
                                       unsigned long long *__restrict__ counter) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t base = i - lane; base < n; base += stride) {
        int64_t k = base + lane;
        bool p = k < n && present[k] != 0;
        unsigned bal = __ballot_sync(0xffffffffu, p);
        if (bal) {
            unsigned long long pos = 0;
            if (lane == 0) pos = atomicAdd(counter, (unsigned long long)__popc(bal));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (p) out[pos + __popc(bal & ((1u << lane) - 1))] = (int32_t)k;
        }
    }
}

template <typename T> __global__ void fill_value_kernel(T *__restrict__ p, T v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// one product of the push traversal: mask tested in-register, then the atomic monoid combine
template <typename SR, typename T>
__device__ __forceinline__ void push_edge(const SR &sr, int64_t k, T uv, const int32_t *__restrict__ colidx, const T *__restrict__ avals,
                                          bool flip, const uint8_t *__restrict__ mask, bool mask_comp, T *__restrict__ t_vals,
                                          uint8_t *__restrict__ t_present) {
    const int32_t j = colidx[k];
    if (mask && ((mask[j] != 0) == mask_comp)) return;   // mask applied in-register before the write
    if (SR::kAddIsAny && !sr.reads_a() && !sr.reads_b()) {
        t_present[j] = 1;   // any_pair: value is the constant 1 written by the fill
    } else {
        const T av = sr.reads_a() ? avals[k] : one_of<T>();
        const T p = flip ? sr.mul(uv, av) : sr.mul(av, uv);
        atomic_combine(sr, &t_vals[j], p);
        t_present[j] = 1;
    }
}

// A warp per frontier vertex.  A vertex with >= PUSH_HEAVY_DEG edges would keep ONE warp busy for milliseconds (the BFS source of
// a Graph500 R-MAT has ~10^5 neighbours): such vertices are only queued here and the whole grid then walks each of their
// adjacency lists together (spmspv_push_heavy_kernel).
constexpr int64_t PUSH_HEAVY_DEG = 4096, PUSH_GRID_DEG = 32768;
template <typename SR, typename T>
__global__ void __launch_bounds__(256)
spmspv_push_kernel(SR sr, const int32_t *__restrict__ frontier, int64_t n_frontier, const int64_t *__restrict__ rowptr,
                   const int32_t *__restrict__ colidx, const T *__restrict__ avals, const T *__restrict__ u, bool flip,
                   const uint8_t *__restrict__ mask, bool mask_comp, T *__restrict__ t_vals, uint8_t *__restrict__ t_present,
                   int32_t *__restrict__ heavy, unsigned long long *__restrict__ heavy_count) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t f = w; f < n_frontier; f += nw) {
        const int32_t i = frontier[f];
        const int64_t b = rowptr[i], e = rowptr[i + 1];
        if (heavy && e - b >= PUSH_HEAVY_DEG) {
            if (lane == 0) heavy[atomicAdd(heavy_count, 1ull)] = i;
            continue;
        }
        const T uv = sr.reads_b() ? u[i] : one_of<T>();
        for (int64_t k = b + lane; k < e; k += 32) push_edge<SR, T>(sr, k, uv, colidx, avals, flip, mask, mask_comp, t_vals, t_present);
    }
}
template <typename SR, typename T>
__global__ void __launch_bounds__(256)
spmspv_push_heavy_kernel(SR sr, const int32_t *__restrict__ heavy, const unsigned long long *__restrict__ heavy_count,
                         const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const T *__restrict__ avals,
                         const T *__restrict__ u, bool flip, const uint8_t *__restrict__ mask, bool mask_comp,
                         T *__restrict__ t_vals, uint8_t *__restrict__ t_present) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    const unsigned long long nh = *heavy_count;
    // heavy vertices (< PUSH_GRID_DEG edges): one CTA each, round-robin -- a frontier can hold hundreds of them
    for (unsigned long long h = blockIdx.x; h < nh; h += gridDim.x) {
        const int32_t i = heavy[h];
        const int64_t b = rowptr[i], e = rowptr[i + 1];
        if (e - b >= PUSH_GRID_DEG) continue;
        const T uv = sr.reads_b() ? u[i] : one_of<T>();
        for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x) push_edge<SR, T>(sr, k, uv, colidx, avals, flip, mask, mask_comp, t_vals, t_present);
    }
    // the few giant ones: the whole grid walks each adjacency list together
    for (unsigned long long h = 0; h < nh; h++) {
        const int32_t i = heavy[h];
        const int64_t b = rowptr[i], e = rowptr[i + 1];
        if (e - b < PUSH_GRID_DEG) continue;
        const T uv = sr.reads_b() ? u[i] : one_of<T>();
        for (int64_t k = b + tid; k < e; k += nt) push_edge<SR, T>(sr, k, uv, colidx, avals, flip, mask, mask_comp, t_vals, t_present);
    }
}

// ------------------------------------------------------------------ host side
static GrB_Info ensure_tiles(CsrArrays &c, int64_t nrows, int64_t nnz, int tile_items, std::string *err) {
    if (c.tile_starts && c.tile_items == tile_items) return GrB_SUCCESS;
    dev_free(c.tile_starts);
    c.tile_starts = nullptr;
    int64_t n_tiles = (nrows + nnz + tile_items - 1) / tile_items;
    c.tile_starts = dev_alloc_t<int64_t>((size_t)n_tiles + 1);
    if (!c.tile_starts) return set_error(err, GrB_OUT_OF_MEMORY, "merge-path tile table");
    c.n_tiles = n_tiles;
    c.tile_items = tile_items;
    LAUNCH_NOTE("merge_search");
    merge_search_kernel<<<(unsigned)((n_tiles + 1 + 255) / 256), 256, 0, g_stream>>>(c.ptr, nrows, nnz, tile_items, n_tiles,
                                                                                 c.tile_starts);
    CUDA_TRY(err, cudaGetLastError());
    return GrB_SUCCESS;
}

template <typename SR, typename T>
static GrB_Info run_pull(const SR &sr, CsrArrays &M, int64_t mrows, int64_t nnz, const T *avals, const T *x, int64_t x_len,
                         const uint8_t *xp, bool flip, const uint8_t *mask, bool mask_comp, T *t_vals,
                         uint8_t *t_present, const VecEpi<T> &epi, std::string *err, int native_val_type = -1) {
    if (mrows == 0) return GrB_SUCCESS;
    const char *method = opt_get("spmv", "auto");
    bool use_rowwarp = !strcmp(method, "rowwarp") || (!strcmp(method, "auto") && mask != nullptr);
    if (use_rowwarp) {
        int64_t warps_needed = (mrows + 31) / 32;   // a warp filters 32 rows at a time
        int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((warps_needed + 7) / 8, (int64_t)g_num_sms * 32));
        LAUNCH_NOTE("spmv_rowwarp");
        if (xp) spmv_rowwarp_kernel<SR, T, false><<<blocks, 256, 0, g_stream>>>(sr, mrows, M.ptr, M.idx, avals, x, xp, flip, mask, mask_comp, t_vals, t_present, epi);
        else spmv_rowwarp_kernel<SR, T, true><<<blocks, 256, 0, g_stream>>>(sr, mrows, M.ptr, M.idx, avals, x, xp, flip, mask, mask_comp, t_vals, t_present, epi);
        CUDA_TRY(err, cudaGetLastError());
        return GrB_SUCCESS;
    }
    // ---- which pull kernel.  Explicit options pick one; "auto" keeps the fused write-back on the merge-path kernel (its
    // per-row emission is coalesced; the segmented kernel emits row by row) and otherwise settles the choice between
    // merge-path / segmented / segmented + hot-column cache by a timed trial on the first multiplies with this CSR:
    // which one wins depends on the value width and on how the labels are laid out (natural R-MAT labels keep the hot
    // columns adjacent, so L1 already serves them; permuted labels make the shared-memory cache win by ~15 %).
    int candidate = 1;   // 1 merge, 2 seg, 3 seg + hot columns, 4 column-banded (spmv_band.cu)
    int trial_cls = -1;
    static cudaEvent_t trial_ev[2] = {nullptr, nullptr};
    // the banded kernel reads the matrix values in place (no typecast copy) and produces plain T
    // (opt-in for the timed trial -- option spmv_band=1 -- or forced with spmv=band: building its second, band-major copy of the
    // matrix costs a sort of all entries, and on the bench matrices it only wins for sparse input vectors: 449 vs 498 us)
    const bool band_forced = !strcmp(method, "band");
    const bool band_ok = native_val_type >= 0 && !epi.active && !flip && (band_forced || opt_get_int("spmv_band", 0) != 0);
    const int n_cand = band_ok && nnz >= opt_get_int("spmv_band_min_nnz", 1 << 22) ? 4 : 3;
    if (!strcmp(method, "seg")) candidate = 2;
    else if (!strcmp(method, "band")) candidate = band_ok ? 4 : 1;
    else if (!strcmp(method, "auto") && !epi.active) {
        const int cls = sizeof(T) >= 8 ? 1 : 0;
        candidate = cls ? 2 : 1;
        const char *hot_opt = opt_get("spmv_hot", "auto");
        if (!strcmp(hot_opt, "1")) candidate = 2;   // forced cache: segmented kernel, spmv_seg_run honours the option
        else if (SR::kStatic && nnz >= opt_get_int("spmv_trial_min_nnz", 1 << 20) && opt_get_int("spmv_trial", 1) != 0) {
            if (M.pull_choice[cls] && (M.pull_choice[cls] != 4 || band_ok)) candidate = M.pull_choice[cls];
            else if (M.pull_choice[cls]) candidate = cls ? 2 : 1;   // the banded winner does not apply to this call
            else {
                candidate = 1 + M.pull_stage[cls] / 2;   // every candidate runs twice: once to set up its cached metadata, once timed
                if (candidate == 3 && !strcmp(hot_opt, "0")) {   // cache switched off: skip its two stages
                    M.pull_ms[cls][2] = 1e30f;
                    M.pull_stage[cls] = 6;
                    candidate = 4;
                }
                if (candidate > n_cand) {   // nothing left to try: settle
                    int best = 0;
                    for (int k = 1; k < 4; k++)
                        if (M.pull_ms[cls][k] >= 0.f && (M.pull_ms[cls][best] < 0.f || M.pull_ms[cls][k] < M.pull_ms[cls][best])) best = k;
                    M.pull_choice[cls] = best + 1;
                    candidate = M.pull_choice[cls];
                } else {
                    trial_cls = cls;
                    if ((M.pull_stage[cls] & 1) == 0) { M.pull_stage[cls]++; trial_cls = -1; }   // warm-up run, untimed
                    for (int q = 0; q < 2 && trial_cls >= 0; q++)
                        if (!trial_ev[q] && cudaEventCreate(&trial_ev[q]) != cudaSuccess) {
                            (void)cudaGetLastError();
                            trial_cls = -1;
                            M.pull_choice[cls] = candidate = cls ? 2 : 1;   // cannot time: keep the static default
                        }
                    if (trial_cls >= 0) cudaEventRecord(trial_ev[0], g_stream);
                }
            }
        }
    }
    auto finish_trial = [&](int ran) {   // `ran`: the candidate that actually executed
        if (trial_cls < 0) return;
        float ms = 0.f;
        cudaEventRecord(trial_ev[1], g_stream);
        if (cudaEventSynchronize(trial_ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, trial_ev[0], trial_ev[1]) != cudaSuccess) {
            (void)cudaGetLastError();
            M.pull_choice[trial_cls] = trial_cls ? 2 : 1;
            return;
        }
        const int stage = M.pull_stage[trial_cls] / 2;
        M.pull_ms[trial_cls][stage] = (stage + 1 != ran) ? 1e30f : ms;   // the candidate did not apply and another kernel ran: it cannot win
        if (++M.pull_stage[trial_cls] >= 2 * n_cand) {
            int best = 0;
            for (int k = 1; k < n_cand; k++)
                if (M.pull_ms[trial_cls][k] < M.pull_ms[trial_cls][best]) best = k;
            M.pull_choice[trial_cls] = best + 1;
        }
    };
    if (candidate == 4) {
        bool handled = false;
        GRB_TRY(spmv_band_run(type_code_of<T>(), sr.add_op(), sr.mul_op(), M, mrows, x_len, nnz, native_val_type, x, xp, t_vals, t_present,
                              err, &handled));
        if (handled) {
            finish_trial(4);
            return GrB_SUCCESS;
        }
    }
    if (candidate >= 2) {
        bool handled = false, used_hot = false;
        const int hot_mode = !strcmp(method, "seg") || !strcmp(opt_get("spmv_hot", "auto"), "1") ? -1 : (candidate == 3 ? 1 : 0);
        GRB_TRY(spmv_seg_run(type_code_of<T>(), sr.add_op(), sr.mul_op(), M, mrows, x_len, nnz, avals, x, xp, t_vals, t_present,
                             &epi, err, &handled, hot_mode, &used_hot));
        if (handled) {
            finish_trial(used_hot ? 3 : 2);
            return GrB_SUCCESS;
        }
    }
    constexpr int TILE = SPMV_BLOCK * SpmvCfg<T>::IPT;
    GRB_TRY(ensure_tiles(M, mrows, nnz, TILE, err));
    const int64_t n_tiles = M.n_tiles;
    int64_t *carry_row = dev_alloc_t<int64_t>((size_t)n_tiles);
    T *carry_val = dev_alloc_t<T>((size_t)n_tiles);
    uint8_t *carry_has = dev_alloc_t<uint8_t>((size_t)n_tiles);
    if (!carry_row || !carry_val || !carry_has) {
        dev_free(carry_row); dev_free(carry_val); dev_free(carry_has);
        return set_error(err, GrB_OUT_OF_MEMORY, "merge-path carry arrays");
    }
    {
        LAUNCH_NOTE("spmv_merge");
        if (xp) spmv_merge_kernel<SR, T, false><<<(unsigned)n_tiles, SPMV_BLOCK, 0, g_stream>>>(sr, mrows, nnz, M.ptr, M.idx, avals, x, xp, M.tile_starts, flip, t_vals, t_present, carry_row, carry_val, carry_has, epi);
        else spmv_merge_kernel<SR, T, true><<<(unsigned)n_tiles, SPMV_BLOCK, 0, g_stream>>>(sr, mrows, nnz, M.ptr, M.idx, avals, x, xp, M.tile_starts, flip, t_vals, t_present, carry_row, carry_val, carry_has, epi);
    }
    {
        LAUNCH_NOTE("spmv_merge_fixup");
        spmv_merge_fixup_kernel<SR, T><<<(unsigned)((n_tiles + 255) / 256), 256, 0, g_stream>>>(sr, n_tiles, mrows, TILE, M.ptr, M.tile_starts, carry_row, carry_val, carry_has, t_vals, t_present, epi);
    }
    cudaError_t e = cudaGetLastError();
    dev_free(carry_row); dev_free(carry_val); dev_free(carry_has);
    CUDA_TRY(err, e);
    finish_trial(1);
    return GrB_SUCCESS;
}

template <typename SR, typename T>
static GrB_Info run_push(const SR &sr, const CsrArrays &A, int64_t arows, int64_t out_len, const T *avals, const T *u,
                         const uint8_t *up, int64_t u_nvals, bool flip, const uint8_t *mask, bool mask_comp, T *t_vals,
                         uint8_t *t_present, int64_t nnz_total, std::string *err) {
    CUDA_TRY(err, cudaMemsetAsync(t_present, 0, (size_t)(out_len > 0 ? out_len : 1), g_stream));
    {
        T init = (SR::kAddIsAny && !sr.reads_a() && !sr.reads_b()) ? one_of<T>() : sr.identity();
        int blocks = (int)std::min<int64_t>((out_len + 255) / 256 + 1, (int64_t)g_num_sms * 16);
        LAUNCH_NOTE("fill_identity");
        fill_value_kernel<T><<<blocks, 256, 0, g_stream>>>(t_vals, init, out_len);
    }
    if (u_nvals == 0 || arows == 0) return GrB_SUCCESS;
    int32_t *frontier = dev_alloc_t<int32_t>((size_t)u_nvals);
    unsigned long long *counter = dev_alloc_t<unsigned long long>(2);   // [0] frontier cursor, [1] heavy-vertex cursor
    // at most nnz / PUSH_HEAVY_DEG vertices can be heavy
    const int64_t heavy_cap = std::min<int64_t>(u_nvals, nnz_total / PUSH_HEAVY_DEG + 1);
    int32_t *heavy = dev_alloc_t<int32_t>((size_t)heavy_cap);
    if (!frontier || !counter || !heavy) { dev_free(frontier); dev_free(counter); dev_free(heavy); return set_error(err, GrB_OUT_OF_MEMORY, "frontier"); }
    CUDA_TRY(err, cudaMemsetAsync(counter, 0, 16, g_stream));
    {
        int blocks = (int)std::min<int64_t>((arows + 255) / 256, (int64_t)g_num_sms * 16);
        LAUNCH_NOTE("compact_frontier");
        compact_present_kernel<<<blocks, 256, 0, g_stream>>>(up, arows, frontier, counter);
    }
    {
        int blocks = (int)std::min<int64_t>((u_nvals + 7) / 8, (int64_t)g_num_sms * 32);
        LAUNCH_NOTE("spmspv_push");
        spmspv_push_kernel<SR, T><<<blocks, 256, 0, g_stream>>>(sr, frontier, u_nvals, A.ptr, A.idx, avals, u, flip, mask, mask_comp, t_vals, t_present, heavy, counter + 1);
    }
    if (nnz_total >= PUSH_HEAVY_DEG) {   // a heavy vertex can exist: the whole grid walks each queued adjacency list (no host sync: the count is read on the device)
        LAUNCH_NOTE("spmspv_push_heavy");
        spmspv_push_heavy_kernel<SR, T><<<g_num_sms * 4, 256, 0, g_stream>>>(sr, heavy, counter + 1, A.ptr, A.idx, avals, u, flip, mask, mask_comp, t_vals, t_present);
    }
    cudaError_t e = cudaGetLastError();
    dev_free(frontier); dev_free(counter); dev_free(heavy);
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

struct MatVecArgs {
    GrB_Matrix A; bool use_transpose; GrB_Vector u; bool flip; const uint8_t *mask; bool mask_comp;
    int add, mul; int64_t out_len; void *t_vals; uint8_t *t_present; std::string *err;
    const VecEpiHost *epi; bool *fused;
};

template <typename T> static GrB_Info mat_vec_typed(const MatVecArgs &a) {
    GrB_Matrix A = a.A;
    GrB_Vector u = a.u;
    GrB_Info info = GrB_SUCCESS;
    // the entry count of u picks the direction of a transposed multiply; a plain pull only uses it to skip the presence bytes of
    // a full vector, so an unknown count (arrays just refilled by a collective / by peers) is NOT recounted there: that would be
    // an O(n) kernel plus a host round trip in every iteration of a partitioned loop
    if (a.use_transpose) GRB_TRY(vector_count(u));
    // decide the traversal: rows of M are directly available (pull) unless M = A' and no CSC twin is wanted
    bool push = false;
    if (a.use_transpose) {
        const char *method = opt_get("vxm_method", "auto");
        if (!strcmp(method, "push")) push = true;
        else if (!strcmp(method, "pull")) push = false;
        else {
            long ratio = opt_get_int("push_ratio", 16);
            push = u->nvals * ratio <= A->nrows;
        }
        if (a.epi && a.epi->peer) push = false;   // the fused exchange lives in the pull kernels' row emission
        if (!push) GRB_TRY(matrix_ensure_twin(A));
    }
    CsrArrays &M = (a.use_transpose && !push) ? A->twin : A->csr;
    const int64_t mrows = (a.use_transpose && !push) ? A->ncols : A->nrows;
    const int T_code = type_code_of<T>();
    // vxm multiplies mul(u_k, a): fold the operand swap into the operator so kernels always compute mul(a, x)
    int mul = a.mul;
    if (a.flip) {
        switch (mul) {
            case OP_FIRST: mul = OP_SECOND; break;
            case OP_SECOND: mul = OP_FIRST; break;
            case OP_MINUS: mul = OP_RMINUS; break;
            case OP_RMINUS: mul = OP_MINUS; break;
            case OP_DIV: mul = OP_RDIV; break;
            case OP_RDIV: mul = OP_DIV; break;
            case OP_POW: mul = OP_RPOW; break;
            case OP_RPOW: mul = OP_POW; break;
            case OP_GT: mul = OP_LT; break;
            case OP_LT: mul = OP_GT; break;
            case OP_GE: mul = OP_LE; break;
            case OP_LE: mul = OP_GE; break;
            default: break;   // commutative multiplies
        }
    }
    const bool kflip = false;
    GRB_DISPATCH_SEMIRING(a.add, mul, T, SRT, sr, {
        const void *av = nullptr, *uv = nullptr;
        void *atmp = nullptr, *utmp = nullptr;
        if (sr.reads_a()) info = cast_view(&av, &atmp, M.val, A->type, T_code, A->nvals, a.err);
        if (!info && sr.reads_b()) info = cast_view(&uv, &utmp, u->vals, u->type, T_code, u->n, a.err);
        if (!info) {
            const uint8_t *up = (u->nvals == u->n) ? nullptr : u->present;
            if (push)
                info = run_push<SRT, T>(sr, A->csr, A->nrows, a.out_len, (const T *)av, (const T *)uv, u->present,
                                        u->nvals, kflip, a.mask, a.mask_comp, (T *)a.t_vals, a.t_present, A->nvals, a.err);
            else {
                VecEpi<T> epi;
                memset(&epi, 0, sizeof epi);
                // pull kernels finish every row exactly once: the write-back is applied there, in registers -- except for 8-byte
                // values without a mask, where the segmented kernel (whose row emission would be uncoalesced with the write-back
                // fused) plus the separate O(n) write-back pass beats the fused merge-path kernel (SSSP: 455 vs 580 us)
                // ... and for every width when the banded kernel (plain output only) has won this CSR's timed trial or is asked for
                const int cls = sizeof(T) >= 8 ? 1 : 0;
                const bool band_wins = !strcmp(opt_get("spmv", "auto"), "band") || (opt_get_int("spmv_band", 0) != 0 && (M.pull_choice[cls] == 4 || M.pull_choice[cls] == 0));
                const bool unfuse = a.epi && !a.epi->peer && !a.mask && !a.epi->has_mask && (sizeof(T) >= 8 || band_wins) &&
                                    (!strcmp(opt_get("spmv", "auto"), "auto") || !strcmp(opt_get("spmv", "auto"), "band")) &&
                                    A->nvals >= opt_get_int("spmv_trial_min_nnz", 1 << 20) && opt_get_int("spmv_unfuse_wide", 1) != 0;
                const int native = (!sr.reads_a() || av == (const void *)M.val) ? A->type : -1;
                if (a.epi && !unfuse) {
                    epi.active = 1;
                    epi.c_vals = (const T *)a.epi->c_vals; epi.c_present = a.epi->c_present; epi.mask = a.epi->mask;
                    epi.has_mask = a.epi->has_mask; epi.comp = a.epi->comp; epi.replace = a.epi->replace; epi.accum = a.epi->accum;
                    if (a.epi->peer) {
                        const PeerTargets &pt = *a.epi->peer;
                        epi.npeer = pt.n;
                        epi.poff = pt.offset;
                        epi.pscale = (const T *)pt.scale;
                        for (int k = 0; k < pt.n; k++) { epi.pv[k] = (T *)pt.vals[k]; epi.pp[k] = pt.present[k]; }
                    }
                    if (a.fused) *a.fused = true;
                }
                info = run_pull<SRT, T>(sr, M, mrows, A->nvals, (const T *)av, (const T *)uv, u->n, up, kflip, a.mask,
                                        a.mask_comp, (T *)a.t_vals, a.t_present, epi, a.err, native);
            }
        }
        dev_free(atmp);
        dev_free(utmp);
    });
    return info;
}

// mask_eff: byte array (1 = mask entry counts) or nullptr; see api.cu for how value masks are reduced to bytes
GrB_Info multiply_mat_vec_impl(void **t_vals_out, uint8_t **t_present_out, int64_t *t_len, const GrB_Semiring op,
                               GrB_Matrix A, bool use_transpose, GrB_Vector u, bool flip, const uint8_t *mask_eff,
                               bool mask_comp, std::string *err, const VecEpiHost *epi, bool *fused) {
    if (fused) *fused = false;
    const int64_t out_len = use_transpose ? A->ncols : A->nrows;
    const int64_t in_len = use_transpose ? A->nrows : A->ncols;
    if (u->n != in_len)
        return set_error(err, GrB_DIMENSION_MISMATCH, "matrix-vector multiply: matrix is %lldx%lld%s, vector has size %lld",
                         (long long)A->nrows, (long long)A->ncols, use_transpose ? " (transposed)" : "", (long long)u->n);
    GRB_TRY(matrix_materialize(A));
    GRB_TRY(vector_ensure_arrays(u));
    const int D = op->type;
    size_t n = (size_t)(out_len > 0 ? out_len : 1);
    void *tv = dev_alloc(n * type_size(D));
    uint8_t *tp = (uint8_t *)dev_alloc(n);
    if (!tv || !tp) { dev_free(tv); dev_free(tp); return set_error(err, GrB_OUT_OF_MEMORY, "result vector"); }
    MatVecArgs a{A, use_transpose, u, flip, mask_eff, mask_comp, op->add, op->mul, out_len, tv, tp, err, epi, fused};
    GrB_Info info = GrB_NOT_IMPLEMENTED;
    GRB_DISPATCH_TYPE(D, T, info = mat_vec_typed<T>(a));
    if (info) { dev_free(tv); dev_free(tp); return info; }
    *t_vals_out = tv;
    *t_present_out = tp;
    *t_len = out_len;
    return GrB_SUCCESS;
}
