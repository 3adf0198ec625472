// spmv_seg.cu -- segmented pull SpMV: t = M (+).(x) u over the CSR rows of M, the default kernel behind GrB_mxv and
// pull GrB_vxm (reference core/matrix.py:2252-2259, core/vector.py:1368-1375).
//
// The entries (not the rows) are split evenly: every lane owns NPL consecutive entries per step and fetches them with
// 128-bit loads of the column-index and value arrays, gathers x, multiplies, and reduces its run in registers.  Row
// boundaries come from one bit per entry ("this entry starts a row", cached per matrix together with the list of
// non-empty rows and the row ordinal in progress at every 128th entry), so the kernel never touches the row pointers
// and never searches: a popc prefix over the warp gives each lane its row ordinal, a warp-shuffle segmented scan
// stitches rows that span lanes, the carry of a row that spans steps stays in registers because a warp walks one
// contiguous range of the matrix, and the few rows that span warps are finished by a tiny fix-up kernel.  About 15
// instructions per entry instead of the ~95 of the merge-path kernel; load balance is exact whatever the degree skew.
//
// Hot-column cache (hotcols.cu): with one persistent CTA per SM, x[most referenced columns] lives in ~200 KB of shared
// memory; the kernel then reads the remapped column array instead of colidx.  The gather is what limits SpMV on a
// power-law graph (one L1 wavefront per distinct line per warp); shared memory serves a lane per bank per cycle.
//
// mask / accum / replace are applied in registers when a row is emitted (epi_write, spmv_common.cuh); rows without
// entries are written by a pre-fill pass.
//
// Algorithmic bytes per multiply (adapted: the row pointers are not read): nnz*(4 + rho*s_val) + nnz/8
// + nonempty_rows*4 + ncols*s_x + nrows*(s_y + 1).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>

#include "spmv_common.cuh"

#ifndef SEG_THREADS_CFG
#define SEG_THREADS_CFG 1024
#endif
constexpr int SEG_THREADS = SEG_THREADS_CFG;   // one persistent CTA per SM
constexpr int SEG_WARPS = SEG_THREADS / 32;
constexpr int SEG_SMEM_BYTES = 227 * 1024 - 1024;

template <typename T> struct SegCfg {
    static constexpr int NPL = sizeof(T) >= 8 ? 4 : 8;   // entries per lane per step
    static constexpr int TILE = 32 * NPL;                // entries per warp per step
};

// ------------------------------------------------------------------ row-boundary metadata (once per CSR)
__global__ void seg_flag_kernel(int64_t nrows, const int64_t *__restrict__ rowptr, unsigned int *__restrict__ flag_words) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += stride) {
        const int64_t b = rowptr[i];
        if (rowptr[i + 1] > b) atomicOr(&flag_words[b >> 5], 1u << (b & 31));   // little endian: bit (b & 7) of byte (b >> 3)
    }
}
struct SegRowNonEmpty {
    const int64_t *p;
    __device__ __forceinline__ bool operator()(const int &i) const { return p[i + 1] > p[i]; }
};
// tile_ord[t] = (number of rows starting at an entry position < 128 t) - 1
__global__ void seg_tile_ord_kernel(int64_t n_entries, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ rows,
                                    const int *__restrict__ n_rows_nonempty, int32_t *__restrict__ tile_ord) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_entries) return;
    const int64_t pos = t * 128;
    int64_t lo = 0, hi = *n_rows_nonempty;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rowptr[rows[mid]] < pos) lo = mid + 1;
        else hi = mid;
    }
    tile_ord[t] = (int32_t)lo - 1;
}

static int64_t seg_tile_ord_len(int64_t nnz) { return 2 * ((nnz + 255) / 256) + 2; }

static GrB_Info ensure_seg(CsrArrays &c, int64_t nrows, int64_t nnz, std::string *err) {
    if (c.seg_state == 1) return GrB_SUCCESS;
    csr_drop_seg(c);
    const size_t flag_bytes = (((size_t)nnz + 7) / 8 + 64 + 3) & ~(size_t)3;
    const int64_t n_ord = seg_tile_ord_len(nnz);
    c.seg_flags = (uint8_t *)dev_alloc(flag_bytes);
    c.seg_rows = dev_alloc_t<int32_t>((size_t)(nrows > 0 ? nrows : 1));
    c.seg_tile_ord = dev_alloc_t<int32_t>((size_t)n_ord);
    int *d_count = dev_alloc_t<int>(1);
    void *tmp = nullptr;
    GrB_Info info = GrB_SUCCESS;
    if (!c.seg_flags || !c.seg_rows || !c.seg_tile_ord || !d_count) info = set_error(err, GrB_OUT_OF_MEMORY, "segmented SpMV metadata");
    if (!info) {
        cudaMemsetAsync(c.seg_flags, 0, flag_bytes, g_stream);
        const int blocks = (int)std::min<int64_t>((nrows + 255) / 256, (int64_t)g_num_sms * 16);
        note_launch("seg_flags");
        seg_flag_kernel<<<blocks, 256, 0, g_stream>>>(nrows, c.ptr, reinterpret_cast<unsigned int *>(c.seg_flags));
        size_t tb = 0;
        thrust::counting_iterator<int> it(0);
        cub::DeviceSelect::If(nullptr, tb, it, c.seg_rows, d_count, (int)nrows, SegRowNonEmpty{c.ptr}, g_stream);
        tmp = dev_alloc(tb);
        if (!tmp) info = set_error(err, GrB_OUT_OF_MEMORY, "segmented SpMV metadata scratch");
        if (!info) {
            note_launch("seg_rows");
            cudaError_t e = cub::DeviceSelect::If(tmp, tb, it, c.seg_rows, d_count, (int)nrows, SegRowNonEmpty{c.ptr}, g_stream);
            if (e != cudaSuccess) info = cuda_fail(err, e, "segmented SpMV row list");
        }
    }
    if (!info) {
        note_launch("seg_tile_ord");
        seg_tile_ord_kernel<<<(unsigned)((n_ord + 255) / 256), 256, 0, g_stream>>>(n_ord, c.ptr, c.seg_rows, d_count, c.seg_tile_ord);
        int h = 0;
        cudaMemcpyAsync(&h, d_count, sizeof(int), cudaMemcpyDeviceToHost, g_stream);
        cudaError_t e = cudaStreamSynchronize(g_stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(err, e, "segmented SpMV metadata");
        c.seg_nonempty = h;
    }
    dev_free(d_count);
    dev_free(tmp);
    if (info) { csr_drop_seg(c); return info; }
    c.seg_state = 1;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ device helpers
template <typename T> struct SegTile {
    int32_t c[SegCfg<T>::NPL];
    T a[SegCfg<T>::NPL];
    unsigned f;        // row-start bits of this lane's entries
    unsigned vmask;    // entries that exist (all ones except in the last tile)
};

template <typename SR, typename T>
__device__ __forceinline__ void seg_load_tile(const SR &sr, SegTile<T> &r, int64_t tile, int lane, int64_t nnz,
                                              const int32_t *__restrict__ cols, const T *__restrict__ avals,
                                              const uint8_t *__restrict__ flags) {
    constexpr int NPL = SegCfg<T>::NPL;
    const int64_t base = tile * SegCfg<T>::TILE + (int64_t)lane * NPL;
    const unsigned fb = flags[base >> 3];   // the flag array is zero padded past the last entry
    r.f = NPL == 8 ? fb : ((fb >> (base & 4)) & 0xFu);
    if (base + NPL <= nnz) {
        unsigned cw[NPL];
        load_words<NPL>(cw, cols + base);
#pragma unroll
        for (int j = 0; j < NPL; j++) r.c[j] = (int32_t)cw[j];
        if (sr.reads_a()) {
            constexpr int VW = NPL * (int)sizeof(T) / 4;
            unsigned vw[VW];
            load_words<VW>(vw, avals + base);
#pragma unroll
            for (int j = 0; j < NPL; j++) r.a[j] = word_elem<T, VW>(vw, j);
        }
        r.vmask = (1u << NPL) - 1u;
    } else {
        r.vmask = 0;
#pragma unroll
        for (int j = 0; j < NPL; j++) {
            r.c[j] = 0;
            r.a[j] = T();
            if (base + j < nnz) {
                r.c[j] = cols[base + j];
                if (sr.reads_a()) r.a[j] = avals[base + j];
                r.vmask |= 1u << j;
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ void seg_emit(const VecEpi<T> &epi, const int32_t *__restrict__ seg_rows, int ord, T v, int h,
                                         T *__restrict__ t_vals, uint8_t *__restrict__ t_present) {
    if (ord < 0) return;
    epi_write(epi, (int64_t)seg_rows[ord], v, h, t_vals, t_present);
}

// per-warp-range boundary records, finished by seg_fixup_kernel
template <typename T> struct SegBounds {
    uint8_t *flag;       // the range contains at least one row start
    T *head; uint8_t *head_has;     // partial of the row in progress at the start of the range, up to the first row start
    T *tail; uint8_t *tail_has;     // partial after the last row start (the whole range when it has none)
    int32_t *ord_first, *ord_last;  // row ordinal in progress at the start / at the end of the range
};

// shared-memory layout of the persistent CTA: per-warp staging of the rows completed in one tile (value + presence
// byte, indexed by ordinal - ordinal at tile start), then the hot-column copy of x
template <typename T> struct SegSmem {
    static constexpr int TILE = SegCfg<T>::TILE;
    static constexpr size_t stage_val_bytes = (size_t)SEG_WARPS * TILE * sizeof(T);
    static constexpr size_t stage_bytes = stage_val_bytes + (size_t)SEG_WARPS * TILE;   // multiple of 16
};

template <typename SR, typename T, bool XFULL, bool HOT>
__global__ void __launch_bounds__(SEG_THREADS, 1)
spmv_seg_kernel(SR sr, int64_t nnz, int64_t n_tiles, const int32_t *__restrict__ cols, const T *__restrict__ avals,
                const T *__restrict__ x, const uint8_t *__restrict__ xp, const uint8_t *__restrict__ flags,
                const int32_t *__restrict__ seg_rows, int n_rows_nonempty, const int32_t *__restrict__ tile_ord,
                const T *__restrict__ xhot, const uint8_t *__restrict__ xhotp, int hot_k, const int32_t *__restrict__ hot_cols,
                T *__restrict__ t_vals, uint8_t *__restrict__ t_present, VecEpi<T> epi, SegBounds<T> bd) {
    constexpr int NPL = SegCfg<T>::NPL;
    constexpr int TILE = SegCfg<T>::TILE;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    T *s_val = reinterpret_cast<T *>(s_dyn) + (size_t)wib * TILE;
    uint8_t *s_has = s_dyn + SegSmem<T>::stage_val_bytes + (size_t)wib * TILE;
    T *s_hot = reinterpret_cast<T *>(s_dyn + SegSmem<T>::stage_bytes);
    uint8_t *s_hotp = s_dyn + SegSmem<T>::stage_bytes + (sr.reads_b() ? (((size_t)hot_k * sizeof(T) + 15) & ~(size_t)15) : 0);
    if (HOT) {
        if (sr.reads_b())
            for (int r = threadIdx.x; r < hot_k; r += SEG_THREADS) s_hot[r] = xhot[r];
        if (!XFULL)
            for (int r = threadIdx.x; r < hot_k; r += SEG_THREADS) s_hotp[r] = xhotp[r];
        __syncthreads();
    }
    const int64_t n_warps = (int64_t)gridDim.x * SEG_WARPS;
    const int64_t w = (int64_t)blockIdx.x * SEG_WARPS + wib;
    const int64_t t_lo = (w * n_tiles) / n_warps, t_hi = ((w + 1) * n_tiles) / n_warps;

    int ord = tile_ord[t_lo * (TILE / 128)];   // row ordinal in progress before the first entry of the range
    const int ord_first = ord;
    T carry_v = sr.identity();                 // partial of the row in progress since the last row start seen by this warp
    int carry_h = 0;
    bool seen = false;                         // a row start has been seen in this range

    SegTile<T> cur, nxt;
    if (t_lo < t_hi) seg_load_tile(sr, cur, t_lo, lane, nnz, cols, avals, flags);
    for (int64_t t = t_lo; t < t_hi; t++) {
        if (t + 1 < t_hi) seg_load_tile(sr, nxt, t + 1, lane, nnz, cols, avals, flags);   // in flight while this tile is reduced
        // rows completed in this tile have the consecutive ordinals ord, ord+1, ...: fetch their row numbers now,
        // coalesced and off the critical path
        int my_row = -1;
        {
            const int o = ord + lane;
            if (o >= 0 && o < n_rows_nonempty) my_row = seg_rows[o];
        }
        // ---- gather x and multiply
        T p[NPL];
        unsigned pm = cur.vmask;   // entries whose product exists
        unsigned hotmask = 0;
#pragma unroll
        for (int j = 0; j < NPL; j++) {
            int32_t c = cur.c[j];
            bool from_global = (cur.vmask >> j) & 1u;
            if (HOT && c < 0) {
                const int rk = c & 0x7fffffff;
                if (rk < hot_k) {
                    hotmask |= 1u << j;
                    from_global = false;
                    cur.c[j] = rk;
                } else {
                    c = hot_cols[rk];
                }
            }
            p[j] = one_of<T>();
            if (from_global) {
                if (sr.reads_b()) p[j] = x[c];
                if (!XFULL) { if (xp[c] == 0) pm &= ~(1u << j); }
            }
        }
        if (HOT) {
#pragma unroll
            for (int j = 0; j < NPL; j++) {
                if ((hotmask >> j) & 1u) {
                    const int rk = cur.c[j];
                    if (sr.reads_b()) p[j] = s_hot[rk];
                    if (!XFULL) { if (s_hotp[rk] == 0) pm &= ~(1u << j); }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NPL; j++) p[j] = sr.mul(sr.reads_a() ? cur.a[j] : one_of<T>(), p[j]);

        // ---- row ordinals: popc prefix of the row-start bits over the warp
        const unsigned f = cur.f;
        const int nfl = __popc(f);
        int incl = nfl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const int excl = incl - nfl;   // completed-row slot of the row in progress when this lane's run begins
        const int total_flags = __shfl_sync(0xffffffffu, incl, 31);

        // ---- in-lane walk: rows that start and end inside the run go to the staging slots excl+1, excl+2, ...
        T acc = sr.identity(), head = sr.identity();
        int acc_h = 0, head_h = 0, slot = excl;
        bool first = true;
#pragma unroll
        for (int j = 0; j < NPL; j++) {
            if ((f >> j) & 1u) {
                if (first) { head = acc; head_h = acc_h; first = false; }
                else { s_val[slot] = acc; s_has[slot] = (uint8_t)acc_h; }
                slot++;
                acc_h = 0;
            }
            if ((pm >> j) & 1u) {
                acc = acc_h ? sr.add(acc, p[j]) : p[j];
                acc_h = 1;
            }
        }
        // ---- segmented scan over lanes: (fl, v, vh) = (run has a row start, partial after its last row start)
        int fl = nfl > 0;
        T v = acc;
        int vh = acc_h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int fh = __shfl_up_sync(0xffffffffu, fl | (vh << 1), o);
            const T vu = shfl_up_any(v, o);
            if (lane >= o) {
                if (!fl) pv_combine(sr, vu, fh >> 1, v, vh);
                fl |= fh & 1;
            }
        }
        // exclusive prefix = what the row in progress has collected in earlier lanes (and, if none of them starts a row,
        // in earlier tiles of this warp)
        int efh = __shfl_up_sync(0xffffffffu, fl | (vh << 1), 1);
        T ev = shfl_up_any(v, 1);
        if (lane == 0) { efh = 0; ev = sr.identity(); }
        int eh = efh >> 1;
        if (!(efh & 1)) pv_combine(sr, carry_v, carry_h, ev, eh);
        if (nfl > 0) {   // the row in progress when this run began ends at the run's first row start
            pv_combine(sr, ev, eh, head, head_h);
            if (!seen && excl == 0) {   // first row start of the whole range: earlier warps may hold part of that row
                bd.head[w] = head;
                bd.head_has[w] = (uint8_t)head_h;
            } else {
                s_val[excl] = head;
                s_has[excl] = (uint8_t)head_h;
            }
        }
        // warp carry for the next tile = inclusive value of lane 31
        {
            const int fh31 = __shfl_sync(0xffffffffu, fl | (vh << 1), 31);
            const T v31 = shfl_any(v, 31);
            if (fh31 & 1) { carry_v = v31; carry_h = fh31 >> 1; }
            else {
                T nv = v31;
                int nh = fh31 >> 1;
                pv_combine(sr, carry_v, carry_h, nv, nh);
                carry_v = nv;
                carry_h = nh;
            }
        }
        // ---- emit the completed rows: slot i is row ordinal ord + i; slot 0 is skipped when it is the range's first
        __syncwarp();
        for (int i = lane; i < total_flags; i += 32) {
            const int o = ord + i;
            if (o < 0 || (i == 0 && !seen)) continue;
            const int row = i < 32 ? my_row : seg_rows[o];
            epi_write(epi, (int64_t)row, s_val[i], (int)s_has[i], t_vals, t_present);
        }
        __syncwarp();
        seen = seen || total_flags > 0;
        ord += total_flags;
        if (t + 1 < t_hi) cur = nxt;
    }
    if (lane == 0) {
        bd.flag[w] = seen ? 1 : 0;
        bd.tail[w] = carry_v;
        bd.tail_has[w] = (uint8_t)carry_h;
        bd.ord_first[w] = ord_first;
        bd.ord_last[w] = ord;
    }
}

// rows that span warp ranges: range w (one that contains a row start) finishes the row that was in progress at its
// start by folding the tails of the ranges before it, back to the nearest one that contains a row start.  Thread
// n_warps finishes the row in progress at the very end of the matrix.
template <typename SR, typename T>
__global__ void seg_fixup_kernel(SR sr, int64_t n_warps, const int32_t *__restrict__ seg_rows, SegBounds<T> bd,
                                 T *__restrict__ t_vals, uint8_t *__restrict__ t_present, VecEpi<T> epi) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_warps) return;
    if (w < n_warps && !bd.flag[w]) return;
    const int ord = w < n_warps ? bd.ord_first[w] : bd.ord_last[n_warps - 1];
    T v = w < n_warps ? bd.head[w] : sr.identity();
    int h = w < n_warps ? bd.head_has[w] : 0;
    for (int64_t q = w - 1; q >= 0; q--) {
        pv_combine(sr, bd.tail[q], (int)bd.tail_has[q], v, h);
        if (bd.flag[q]) break;
    }
    seg_emit(epi, seg_rows, ord, v, h, t_vals, t_present);
}

// rows without entries (and, before the main kernel overwrites them, all others): T is absent there
template <typename T>
__global__ void seg_prefill_kernel(int64_t n, VecEpi<T> epi, T *__restrict__ t_vals, uint8_t *__restrict__ t_present) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) epi_write(epi, i, T(), 0, t_vals, t_present);
}

template <typename T>
__global__ void seg_hot_gather_kernel(const T *__restrict__ x, const uint8_t *__restrict__ xp, const int32_t *__restrict__ hot_cols,
                                      int hot_k, T *__restrict__ xhot, uint8_t *__restrict__ xhotp) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= hot_k) return;
    const int32_t c = hot_cols[r];
    if (x) xhot[r] = x[c];
    if (xp) xhotp[r] = xp[c];
}

// ------------------------------------------------------------------ host side
// number of hot ranks to keep in shared memory for this multiply (0 = plain kernel)
static int seg_hot_plan(CsrArrays &M, int64_t ncols, int64_t nnz, size_t elem_bytes, size_t hot_budget, bool xfull, bool reads_b,
                        int hot_mode, std::string *err) {
    const char *mode = opt_get("spmv_hot", "auto");
    if (hot_mode == 0 || (hot_mode < 0 && strcmp(mode, "1")) || (!reads_b && xfull)) return 0;   // off unless forced / asked for
    const bool force = hot_mode < 0;   // option "1": no thresholds (tests); hot_mode 1: the caller's trial, thresholds apply
    const int max_ranks = SEG_SMEM_BYTES / 4;   // ranks kept in the remap: what a 4-byte table could ever hold
    if (M.hot_state == 0) {
        // the analysis costs a few multiplies' worth of time: the caller's trial asks for it on the third multiply with this CSR
        if (!force && nnz < opt_get_int("spmv_hot_min_nnz", 1 << 20)) return 0;
        if (csr_ensure_hot(M, ncols, nnz, max_ranks, err) != GrB_SUCCESS) return 0;
    }
    if (M.hot_state != 1 || M.hot_n == 0) return 0;
    const size_t per = (reads_b ? elem_bytes : 0) + (xfull ? 0 : 1);
    if (hot_budget <= 64) return 0;
    int k = (int)std::min<int64_t>(M.hot_n, (int64_t)((hot_budget - 64) / per));
    const long test_cap = opt_get_int("spmv_hot_cap", 0);   // tests: shrink the cache so that ranks >= k take the global path
    if (test_cap > 0 && k > test_cap) k = (int)test_cap;
    if (k <= 0) return 0;
    const double cover = (double)M.hot_prefix[k - 1] / (double)nnz;
    if (!force && cover < 0.01 * (double)opt_get_int("spmv_hot_min_cover_pct", 20)) return 0;
    return k;
}

template <typename SR, typename T>
static GrB_Info seg_run_typed(const SR &sr, CsrArrays &M, int64_t mrows, int64_t ncols, int64_t nnz, const T *avals, const T *x,
                              const uint8_t *xp, T *t_vals, uint8_t *t_present, const VecEpi<T> &epi, std::string *err,
                              bool *handled, int hot_mode, bool *used_hot) {
    *handled = false;
    if (mrows <= 0) { *handled = true; return GrB_SUCCESS; }
    // pre-fill: the result where T has no entry (every row for now; rows with entries are rewritten below)
    {
        if (epi.active) {
            const int blocks = (int)std::min<int64_t>((mrows + 255) / 256, (int64_t)g_num_sms * 16);
            LAUNCH_NOTE("spmv_seg_prefill");
            seg_prefill_kernel<T><<<blocks, 256, 0, g_stream>>>(mrows, epi, t_vals, t_present);
        } else {
            CUDA_TRY(err, cudaMemsetAsync(t_vals, 0, (size_t)mrows * sizeof(T), g_stream));
            CUDA_TRY(err, cudaMemsetAsync(t_present, 0, (size_t)mrows, g_stream));
        }
    }
    *handled = true;
    if (nnz <= 0) return GrB_SUCCESS;
    GRB_TRY(ensure_seg(M, mrows, nnz, err));
    // shared memory: per-warp staging + hot table.  What is not given to shared memory stays L1, which the gathers that
    // miss the hot table still need (cache lines double as miss buffers): `spmv_hot_kb` bounds the total carve-out.
    const size_t stage_bytes = SegSmem<T>::stage_bytes;
    size_t smem_total = (size_t)opt_get_int("spmv_hot_kb", 132) * 1024 - 1024;
    if (smem_total > (size_t)SEG_SMEM_BYTES) smem_total = SEG_SMEM_BYTES;
    const size_t hot_budget = smem_total > stage_bytes ? smem_total - stage_bytes : 0;
    int hot_k = SR::kStatic ? seg_hot_plan(M, ncols, nnz, sizeof(T), hot_budget, xp == nullptr, sr.reads_b(), hot_mode, err) : 0;
    if (hot_mode == 0) hot_k = 0;
    const int32_t *cols = hot_k ? M.hot_remap : M.idx;
    if (used_hot) *used_hot = hot_k > 0;

    constexpr int TILE = SegCfg<T>::TILE;
    const int64_t n_tiles = (nnz + TILE - 1) / TILE;
    const int grid = (int)std::min<int64_t>((n_tiles + SEG_WARPS - 1) / SEG_WARPS, (int64_t)g_num_sms);
    const int64_t n_warps = (int64_t)grid * SEG_WARPS;
    // boundary records
    const size_t rec = sizeof(T) * 2 + 3 + 8;
    unsigned char *brec = (unsigned char *)dev_alloc((size_t)n_warps * rec + 64);
    T *xhot = (hot_k && sr.reads_b()) ? dev_alloc_t<T>((size_t)hot_k) : nullptr;
    uint8_t *xhotp = (hot_k && xp) ? dev_alloc_t<uint8_t>((size_t)hot_k) : nullptr;
    if (!brec || (hot_k && sr.reads_b() && !xhot) || (hot_k && xp && !xhotp)) {
        dev_free(brec); dev_free(xhot); dev_free(xhotp);
        return set_error(err, GrB_OUT_OF_MEMORY, "segmented SpMV scratch");
    }
    SegBounds<T> bd;
    {
        unsigned char *q = brec;   // 8-byte fields first so that every array is naturally aligned
        bd.head = reinterpret_cast<T *>(q); q += (((size_t)n_warps * sizeof(T)) + 7) & ~(size_t)7;
        bd.tail = reinterpret_cast<T *>(q); q += (((size_t)n_warps * sizeof(T)) + 7) & ~(size_t)7;
        bd.ord_first = reinterpret_cast<int32_t *>(q); q += (size_t)n_warps * 4;
        bd.ord_last = reinterpret_cast<int32_t *>(q); q += (size_t)n_warps * 4;
        bd.flag = q; q += n_warps;
        bd.head_has = q; q += n_warps;
        bd.tail_has = q;
    }
    cudaError_t e = cudaSuccess;
    if (hot_k) {
        LAUNCH_NOTE("spmv_hot_gather");
        seg_hot_gather_kernel<T><<<(hot_k + 255) / 256, 256, 0, g_stream>>>(sr.reads_b() ? x : nullptr, xp, M.hot_cols, hot_k, xhot, xhotp);
    }
    {
        const size_t smem = hot_k ? smem_total : stage_bytes;
        LAUNCH_NOTE(hot_k ? "spmv_seg_hot" : "spmv_seg");
#define SEG_LAUNCH(XF, HT)                                                                                                     \
    do {                                                                                                                       \
        auto kern = spmv_seg_kernel<SR, T, XF, HT>;                                                                            \
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e == cudaSuccess)                                                                                                  \
            kern<<<grid, SEG_THREADS, smem, g_stream>>>(sr, nnz, n_tiles, cols, avals, x, xp, M.seg_flags, M.seg_rows,         \
                                                        (int)M.seg_nonempty, M.seg_tile_ord, xhot, xhotp, hot_k, M.hot_cols,   \
                                                        t_vals, t_present, epi, bd);                                           \
    } while (0)
        bool launched = false;
        if constexpr (SR::kStatic) {   // the hot-column variant exists for the compile-time specialised semirings only
            if (hot_k) {
                if (xp) SEG_LAUNCH(false, true);
                else SEG_LAUNCH(true, true);
                launched = true;
            }
        }
        if (!launched) {
            if (xp) SEG_LAUNCH(false, false);
            else SEG_LAUNCH(true, false);
        }
#undef SEG_LAUNCH
    }
    if (e == cudaSuccess) {
        LAUNCH_NOTE("spmv_seg_fixup");
        seg_fixup_kernel<SR, T><<<(unsigned)((n_warps + 1 + 255) / 256), 256, 0, g_stream>>>(sr, n_warps, M.seg_rows, bd, t_vals, t_present, epi);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    dev_free(brec); dev_free(xhot); dev_free(xhotp);
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

GrB_Info spmv_seg_run(int type_code, int add_op, int mul_op, CsrArrays &M, int64_t mrows, int64_t ncols, int64_t nnz,
                      const void *avals, const void *x, const uint8_t *xp, void *t_vals, uint8_t *t_present,
                      const void *epi_typed, std::string *err, bool *handled, int hot_mode, bool *used_hot) {
    *handled = false;
    if (used_hot) *used_hot = false;
    // 128-bit loads: the index and value arrays must be 16-byte aligned (library allocations always are)
    if (((uintptr_t)M.idx & 15) || ((uintptr_t)avals & 15)) return GrB_SUCCESS;
    GrB_Info info = GrB_SUCCESS;
    GRB_DISPATCH_TYPE(type_code, T, {
        GRB_DISPATCH_SEMIRING(add_op, mul_op, T, SRT, sr, {
            info = seg_run_typed<SRT, T>(sr, M, mrows, ncols, nnz, (const T *)avals, (const T *)x, xp, (T *)t_vals, t_present,
                                         *reinterpret_cast<const VecEpi<T> *>(epi_typed), err, handled, hot_mode, used_hot);
        });
    });
    return info;
}
