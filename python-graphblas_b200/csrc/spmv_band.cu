// spmv_band.cu -- column-banded pull SpMV: t = M (+).(x) u with the gathers served from shared memory.
//
// What bounds a pull SpMV on a power-law graph is the gather x[col]: from L1 / L2 it costs one wavefront per distinct 32-byte
// sector a warp touches (372 G gathers/s on the scale-22 R-MAT = 180 us for one multiply, whatever the load flavour), from
// shared memory a lane per bank per cycle (1 000 G/s measured, profiles/gather_bench_r01.txt).  So the matrix is kept a second
// time in COLUMN-BAND-MAJOR order (built once per CSR, cached like the transpose twin): band b holds the entries whose column
// lies in [b * 2^BITS, (b + 1) * 2^BITS) -- a window of x that fits 128 KB of shared memory -- sorted by row inside the band.
//   * a persistent CTA per SM walks a contiguous range of that order; when it enters a band it stages the band's x window in
//     shared memory (bank-swizzled: R-MAT's popular columns would otherwise all sit in bank 0), then its 32 warps stream the
//     band's entries, the next tile's loads in flight while the current one is reduced:
//     128-bit loads of 16-bit LOCAL column indices (half the index traffic of the CSR) and of the values, gather from the
//     window, multiply, and reduce runs of equal row ("segments") with the in-lane walk + warp-shuffle segmented scan of
//     the segmented kernel;
//   * segments are cut at every tile (256 entries, 128 for 8-byte values) boundary when the format is built and bands are
//     padded to whole tiles, so a tile is self-contained: no carries between tiles, warps or CTAs, no fix-up kernel;
//   * a finished segment is one atomic monoid combine into t(row) (a row meets at most one segment per band and tile cut):
//     RED.ADD / MIN / MAX for the hot monoids, a CAS loop otherwise.  t is pre-filled with the monoid identity.
// Algorithmic bytes per multiply: nnz * (2 + rho * s_val) + nnz / 8 + segments * 4 + ncols * s_x + nrows * (s_y + 1);
// the scale-22 Graph500 matrix has 19.6 M segments for its 65.2 M entries (3.3 entries per segment at 2^15 columns per band).
// Floating-point sums are accumulated in atomic order (not run-to-run deterministic in the last bits); integer, min / max
// and boolean results are exact.  Plain T output only: mask / accum are applied by the separate write-back pass.
//
// Serves GrB_mxv / pull GrB_vxm (reference core/matrix.py:2252-2259, core/vector.py:1368-1375) when the library's timed trial
// (spmv.cu run_pull) finds it fastest for the CSR at hand, or with option spmv=band.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "spmv_common.cuh"

constexpr int BAND_THREADS = 1024;   // one persistent CTA per SM
constexpr int BAND_WARPS = BAND_THREADS / 32;

template <int ES> struct BandCls {
    static constexpr int NPL = ES >= 8 ? 4 : 8;     // entries per lane per tile
    static constexpr int TILE = 32 * NPL;
    static constexpr int BITS = ES >= 8 ? 14 : 15;  // 2^BITS columns per band: a 128 KB window of 4- or 8-byte values
};

struct BandFormat {
    int bits = 0, tile = 0, n_bands = 0, val_type = -1;
    int64_t n_tiles = 0, n_seg = 0, nnz_pad = 0;
    uint16_t *col16 = nullptr;     // nnz_pad local column indices
    void *vals = nullptr;          // nnz_pad values in the matrix's own type
    uint8_t *flags = nullptr;      // bit per entry: starts a segment
    uint32_t *seg_base = nullptr;  // n_tiles: ordinal of the tile's first segment
    int32_t *seg_row = nullptr;    // n_seg: row of the segment (-1: padding)
    int32_t *band_tile = nullptr;  // n_bands + 1: first tile of every band
    int32_t *tile_band = nullptr;  // n_tiles
};

static void band_free(BandFormat *f) {
    if (!f) return;
    dev_free(f->col16); dev_free(f->vals); dev_free(f->flags); dev_free(f->seg_base); dev_free(f->seg_row); dev_free(f->band_tile);
    dev_free(f->tile_band);
    delete f;
}
void csr_drop_band(CsrArrays &c) {
    for (int q = 0; q < 2; q++) {
        band_free(reinterpret_cast<BandFormat *>(c.band_fmt[q]));
        c.band_fmt[q] = nullptr;
    }
}

// ------------------------------------------------------------------ format construction (once per CSR and element-size class)
__global__ void band_keys_kernel(int64_t nnz, const int32_t *__restrict__ idx, int bits, uint16_t *__restrict__ key, int32_t *__restrict__ ent) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; k < nnz; k += s) { key[k] = (uint16_t)((uint32_t)idx[k] >> bits); ent[k] = (int32_t)k; }
}
__global__ void band_starts_kernel(int64_t nnz, const uint16_t *__restrict__ key_s, int64_t *__restrict__ start) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; k < nnz; k += s)
        if (k == 0 || key_s[k] != key_s[k - 1]) start[key_s[k]] = k;
}
// sorted position q -> padded position; fills the local column, the row (temporary) and the value of the entry
template <typename U>
__global__ void band_scatter_kernel(int64_t nnz, int64_t nrows, const uint16_t *__restrict__ key_s, const int32_t *__restrict__ perm,
                                    const int64_t *__restrict__ start, const int64_t *__restrict__ pstart, const int64_t *__restrict__ ptr,
                                    const int32_t *__restrict__ idx, const U *__restrict__ val, int bits, uint16_t *__restrict__ col16,
                                    int32_t *__restrict__ rowtmp, U *__restrict__ vals) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; q < nnz; q += s) {
        const int b = key_s[q];
        const int64_t pq = pstart[b] + (q - start[b]);
        const int64_t e = perm[q];
        int64_t lo = 0, hi = nrows;   // row of entry e: the last r with ptr[r] <= e
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (ptr[mid] <= e) lo = mid;
            else hi = mid - 1;
        }
        col16[pq] = (uint16_t)((uint32_t)idx[e] & ((1u << bits) - 1u));
        rowtmp[pq] = (int32_t)lo;
        if (val) vals[pq] = val[e];
    }
}
__global__ void band_heads_kernel(int64_t nnz_pad, int tile, const int32_t *__restrict__ rowtmp, int32_t *__restrict__ head) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; k <= nnz_pad; k += s) head[k] = (k < nnz_pad && (k % tile == 0 || rowtmp[k] != rowtmp[k - 1])) ? 1 : 0;
}
// segid = exclusive scan of the heads: flag bits, first segment of every tile, row of every segment
__global__ void band_finish_kernel(int64_t nnz_pad, int tile, const int32_t *__restrict__ rowtmp, const int32_t *__restrict__ segid,
                                   uint8_t *__restrict__ flags, uint32_t *__restrict__ seg_base, int32_t *__restrict__ seg_row) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 entries
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; g * 8 < nnz_pad; g += s) {
        unsigned byte = 0;
        for (int j = 0; j < 8; j++) {
            const int64_t k = g * 8 + j;
            if (k >= nnz_pad) break;
            const bool h = segid[k + 1] != segid[k];
            if (h) {
                byte |= 1u << j;
                seg_row[segid[k]] = rowtmp[k];
            }
            if (k % tile == 0) seg_base[k / tile] = (uint32_t)segid[k];
        }
        flags[g] = (uint8_t)byte;
    }
}
__global__ void band_tile_band_kernel(int64_t n_tiles, int n_bands, const int32_t *__restrict__ band_tile, int32_t *__restrict__ tile_band) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    int lo = 0, hi = n_bands - 1;   // the last band whose first tile is <= t (bands without entries share their successor's start)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (band_tile[mid] <= t) lo = mid;
        else hi = mid - 1;
    }
    tile_band[t] = lo;
}

static GrB_Info band_build(BandFormat **out, const CsrArrays &c, int64_t nrows, int64_t ncols, int64_t nnz, int val_type, int es_cls,
                           std::string *err) {
    const int bits = es_cls ? BandCls<8>::BITS : BandCls<4>::BITS, tile = es_cls ? BandCls<8>::TILE : BandCls<4>::TILE;
    const int64_t nb64 = (ncols + ((int64_t)1 << bits) - 1) >> bits;
    if (nb64 > 65535 || nnz >= ((int64_t)1 << 31) - 2 * 65536 * tile) return GrB_NO_VALUE;   // 16-bit band keys, 32-bit positions
    const int n_bands = (int)std::max<int64_t>(1, nb64);
    BandFormat *f = new (std::nothrow) BandFormat();
    if (!f) return GrB_OUT_OF_MEMORY;
    f->bits = bits; f->tile = tile; f->n_bands = n_bands; f->val_type = val_type;
    const size_t vs = c.val ? type_size(val_type) : 0;
    uint16_t *key = dev_alloc_t<uint16_t>((size_t)nnz), *key_s = dev_alloc_t<uint16_t>((size_t)nnz);
    int32_t *ent = dev_alloc_t<int32_t>((size_t)nnz), *perm = dev_alloc_t<int32_t>((size_t)nnz);
    int64_t *start = dev_alloc_t<int64_t>((size_t)n_bands + 1), *pstart = dev_alloc_t<int64_t>((size_t)n_bands + 1);
    void *tmp = nullptr;
    int32_t *rowtmp = nullptr, *head = nullptr;
    GrB_Info info = GrB_SUCCESS;
    auto fail = [&](GrB_Info i, const char *what) { info = set_error(err, i, "banded SpMV format: %s", what); };
    const int blocks = (int)std::min<int64_t>((nnz + 255) / 256 + 1, (int64_t)g_num_sms * 16);
    if (!key || !key_s || !ent || !perm || !start || !pstart) fail(GrB_OUT_OF_MEMORY, "sort arrays");
    if (!info) {
        note_launch("band_keys");
        band_keys_kernel<<<blocks, 256, 0, g_stream>>>(nnz, c.idx, bits, key, ent);
        int key_bits = 1;
        while ((1 << key_bits) < n_bands) key_bits++;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key_s, ent, perm, nnz, 0, key_bits, g_stream);
        tmp = dev_alloc(tb);
        if (!tmp) fail(GrB_OUT_OF_MEMORY, "sort scratch");
        else {
            note_launch("cub_radix_sort");
            cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, key, key_s, ent, perm, nnz, 0, key_bits, g_stream);   // stable: (row, col) order survives inside a band
            if (e != cudaSuccess) info = cuda_fail(err, e, "banded SpMV format: radix sort");
        }
    }
    std::vector<int64_t> hstart((size_t)n_bands + 1), hp((size_t)n_bands + 1);
    std::vector<int32_t> hbt((size_t)n_bands + 1);
    if (!info) {
        cudaMemsetAsync(start, 0xff, sizeof(int64_t) * ((size_t)n_bands + 1), g_stream);   // -1: band without entries
        note_launch("band_starts");
        band_starts_kernel<<<blocks, 256, 0, g_stream>>>(nnz, key_s, start);
        cudaMemcpyAsync(hstart.data(), start, sizeof(int64_t) * (size_t)n_bands, cudaMemcpyDeviceToHost, g_stream);
        cudaError_t e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) info = cuda_fail(err, e, "banded SpMV format: band starts");
    }
    if (!info) {
        hstart[(size_t)n_bands] = nnz;
        for (int b = n_bands - 1; b >= 0; b--)
            if (hstart[(size_t)b] < 0) hstart[(size_t)b] = hstart[(size_t)b + 1];   // empty band: zero length
        int64_t run = 0;
        for (int b = 0; b < n_bands; b++) {
            hp[(size_t)b] = run;
            hbt[(size_t)b] = (int32_t)(run / tile);
            const int64_t cnt = hstart[(size_t)b + 1] - hstart[(size_t)b];
            run += (cnt + tile - 1) / tile * tile;
        }
        hp[(size_t)n_bands] = run;
        hbt[(size_t)n_bands] = (int32_t)(run / tile);
        f->nnz_pad = run;
        f->n_tiles = run / tile;
        cudaMemcpyAsync(start, hstart.data(), sizeof(int64_t) * ((size_t)n_bands + 1), cudaMemcpyHostToDevice, g_stream);
        cudaMemcpyAsync(pstart, hp.data(), sizeof(int64_t) * ((size_t)n_bands + 1), cudaMemcpyHostToDevice, g_stream);
        const size_t np = (size_t)std::max<int64_t>(run, 1);
        f->col16 = dev_alloc_t<uint16_t>(np + 64);
        f->vals = vs ? dev_alloc(np * vs + 64) : nullptr;
        f->flags = (uint8_t *)dev_alloc(np / 8 + 64);
        f->seg_base = dev_alloc_t<uint32_t>((size_t)f->n_tiles + 1);
        f->band_tile = dev_alloc_t<int32_t>((size_t)n_bands + 1);
        f->tile_band = dev_alloc_t<int32_t>((size_t)f->n_tiles + 1);
        rowtmp = dev_alloc_t<int32_t>(np + 1);
        head = dev_alloc_t<int32_t>(np + 2);
        if (!f->col16 || (vs && !f->vals) || !f->flags || !f->seg_base || !f->band_tile || !f->tile_band || !rowtmp || !head)
            fail(GrB_OUT_OF_MEMORY, "band arrays");
    }
    if (!info && f->nnz_pad > 0) {
        const size_t np = (size_t)f->nnz_pad;
        cudaMemsetAsync(f->col16, 0, np * 2, g_stream);
        if (vs) cudaMemsetAsync(f->vals, 0, np * vs, g_stream);
        cudaMemsetAsync(rowtmp, 0xff, (np + 1) * 4, g_stream);   // -1: padding
        cudaMemcpyAsync(f->band_tile, hbt.data(), sizeof(int32_t) * ((size_t)n_bands + 1), cudaMemcpyHostToDevice, g_stream);
        note_launch("band_scatter");
        switch (vs) {
            case 0: band_scatter_kernel<uint8_t><<<blocks, 256, 0, g_stream>>>(nnz, nrows, key_s, perm, start, pstart, c.ptr, c.idx, nullptr, bits, f->col16, rowtmp, nullptr); break;
            case 1: band_scatter_kernel<uint8_t><<<blocks, 256, 0, g_stream>>>(nnz, nrows, key_s, perm, start, pstart, c.ptr, c.idx, (const uint8_t *)c.val, bits, f->col16, rowtmp, (uint8_t *)f->vals); break;
            case 2: band_scatter_kernel<uint16_t><<<blocks, 256, 0, g_stream>>>(nnz, nrows, key_s, perm, start, pstart, c.ptr, c.idx, (const uint16_t *)c.val, bits, f->col16, rowtmp, (uint16_t *)f->vals); break;
            case 4: band_scatter_kernel<uint32_t><<<blocks, 256, 0, g_stream>>>(nnz, nrows, key_s, perm, start, pstart, c.ptr, c.idx, (const uint32_t *)c.val, bits, f->col16, rowtmp, (uint32_t *)f->vals); break;
            default: band_scatter_kernel<uint64_t><<<blocks, 256, 0, g_stream>>>(nnz, nrows, key_s, perm, start, pstart, c.ptr, c.idx, (const uint64_t *)c.val, bits, f->col16, rowtmp, (uint64_t *)f->vals); break;
        }
        note_launch("band_heads");
        band_heads_kernel<<<blocks, 256, 0, g_stream>>>(f->nnz_pad, tile, rowtmp, head);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, head, head, f->nnz_pad + 1, g_stream);
        void *tmp2 = dev_alloc(tb);
        if (!tmp2) fail(GrB_OUT_OF_MEMORY, "scan scratch");
        else {
            note_launch("cub_exclusive_sum");
            cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp2, tb, head, head, f->nnz_pad + 1, g_stream);
            int32_t nseg = 0;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&nseg, head + f->nnz_pad, sizeof(int32_t), cudaMemcpyDeviceToHost, g_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
            dev_free(tmp2);
            if (e != cudaSuccess) info = cuda_fail(err, e, "banded SpMV format: segment scan");
            f->n_seg = nseg;
        }
    }
    if (!info && f->nnz_pad > 0) {
        f->seg_row = dev_alloc_t<int32_t>((size_t)std::max<int64_t>(f->n_seg, 1));
        if (!f->seg_row) fail(GrB_OUT_OF_MEMORY, "segment rows");
        else {
            note_launch("band_finish");
            band_finish_kernel<<<blocks, 256, 0, g_stream>>>(f->nnz_pad, tile, rowtmp, head, f->flags, f->seg_base, f->seg_row);
            band_tile_band_kernel<<<(unsigned)((f->n_tiles + 255) / 256), 256, 0, g_stream>>>(f->n_tiles, n_bands, f->band_tile, f->tile_band);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) info = cuda_fail(err, e, "banded SpMV format: finish");
        }
    }
    dev_free(key); dev_free(key_s); dev_free(ent); dev_free(perm); dev_free(start); dev_free(pstart); dev_free(tmp); dev_free(rowtmp); dev_free(head);
    if (info) { band_free(f); return info; }
    *out = f;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ the kernel
template <typename T> struct BandArgs {
    int64_t n_tiles, ncols;
    const uint16_t *col16; const T *vals; const uint8_t *flags; const uint32_t *seg_base; const int32_t *seg_row;
    const int32_t *band_tile; const int32_t *tile_band;
    const T *x; const uint8_t *xp;
    T *t_vals; uint8_t *t_present;
};

template <typename T> struct BandSmem {
    static constexpr int TILE = BandCls<sizeof(T)>::TILE;
    static constexpr int WIN = 1 << BandCls<sizeof(T)>::BITS;
    static constexpr size_t win_bytes = (size_t)WIN * sizeof(T);                       // 128 KB for 4- and 8-byte values
    static constexpr size_t stage_val = (size_t)BAND_WARPS * TILE * sizeof(T);
    static constexpr size_t stage_has = (size_t)BAND_WARPS * TILE;
    static size_t total(bool xfull) { return 16 + win_bytes + stage_val + (xfull ? 0 : stage_has + (size_t)WIN); }
};

// shared-memory position of local column c: the low five bits (the bank of a 4-byte value) are XORed with a hash of the higher
// bits.  R-MAT's popular columns are the ones with few set bits -- 0, 32, 64, 1024, ... all live in bank 0 of a linear window
// (measured: 7-way bank conflicts on the gathers) -- the hash spreads them; within an aligned group of 32 columns it is a
// permutation, so filling the window stays conflict free.
__device__ __forceinline__ int band_swz(int c) { return c ^ (int)((((unsigned)c >> 5) * 0x9E3779B1u) >> 27); }

template <typename T> struct BandTile {
    static constexpr int NPL = BandCls<sizeof(T)>::NPL;
    unsigned cw[NPL / 2];
    T av[NPL];
    unsigned f;
    uint32_t sb;
};
template <typename SR, typename T>
__device__ __forceinline__ void band_load_tile(const SR &sr, BandTile<T> &r, const BandArgs<T> &a, int64_t tt, int lane) {
    constexpr int NPL = BandCls<sizeof(T)>::NPL;
    constexpr int TILE = BandCls<sizeof(T)>::TILE;
    const int64_t base = tt * TILE + (int64_t)lane * NPL;
    const unsigned fb = a.flags[base >> 3];
    r.f = NPL == 8 ? fb : ((fb >> (base & 4)) & 0xFu);
    load_words<NPL / 2>(r.cw, a.col16 + base);
    if (sr.reads_a()) {
        constexpr int VW = (NPL * (int)sizeof(T) / 4) >= 2 ? NPL * (int)sizeof(T) / 4 : 2;
        unsigned vw[VW];
        load_words<VW>(vw, a.vals + base);
#pragma unroll
        for (int j = 0; j < NPL; j++) r.av[j] = word_elem<T, VW>(vw, j);
    }
    r.sb = a.seg_base[tt];
}

template <typename SR, typename T, bool XFULL>
__global__ void __launch_bounds__(BAND_THREADS, 1)
spmv_band_kernel(SR sr, BandArgs<T> a) {
    constexpr int NPL = BandCls<sizeof(T)>::NPL;
    constexpr int TILE = BandCls<sizeof(T)>::TILE;
    constexpr int BITS = BandCls<sizeof(T)>::BITS;
    constexpr int WIN = 1 << BITS;
    constexpr int NSR = TILE / 64;   // segment rows fetched ahead per lane: covers the first TILE / 2 segments of a tile
    extern __shared__ __align__(128) unsigned char s_band[];
    T *xs = reinterpret_cast<T *>(s_band + 16);
    T *s_val_all = reinterpret_cast<T *>(s_band + 16 + BandSmem<T>::win_bytes);
    uint8_t *s_has_all = s_band + 16 + BandSmem<T>::win_bytes + BandSmem<T>::stage_val;
    uint8_t *xps = s_has_all + BandSmem<T>::stage_has;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    T *s_val = s_val_all + (size_t)wib * TILE;
    uint8_t *s_has = s_has_all + (size_t)wib * TILE;
    const int64_t t_lo = ((int64_t)blockIdx.x * a.n_tiles) / gridDim.x, t_hi = ((int64_t)(blockIdx.x + 1) * a.n_tiles) / gridDim.x;
    int64_t t = t_lo;
    while (t < t_hi) {
        const int b = a.tile_band[t];
        int64_t t_end = a.band_tile[b + 1];
        if (t_end > t_hi) t_end = t_hi;
        // ---- stage the band's window of x in shared memory (swizzled, see band_swz): coalesced loads, conflict-free stores
        const int64_t c0 = (int64_t)b << BITS;
        const int wn = (int)((a.ncols - c0 < WIN) ? (a.ncols - c0) : WIN);
        BandTile<T> cur, nxt;
        if (t + wib < t_end) band_load_tile(sr, cur, a, t + wib, lane);   // in flight while the window is staged
        __syncthreads();   // every warp is done with the previous window
        if (sr.reads_b())
            for (int i = tid; i < wn; i += BAND_THREADS) xs[band_swz(i)] = a.x[c0 + i];
        if (!XFULL)
            for (int i = tid; i < wn; i += BAND_THREADS) xps[band_swz(i)] = a.xp[c0 + i];
        __syncthreads();
        // ---- the band's tiles go round the warps; a tile is self-contained; the next tile's loads overlap this tile's work
        for (int64_t tt = t + wib; tt < t_end; tt += BAND_WARPS) {
            if (tt + BAND_WARPS < t_end) band_load_tile(sr, nxt, a, tt + BAND_WARPS, lane);
            const unsigned f = cur.f;
            const uint32_t sb = cur.sb;
            // ---- segment ordinals: popc prefix of the start bits over the warp (the tile's first entry always starts one)
            const int nfl = __popc(f);
            int incl = nfl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const int excl = incl - nfl;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            // rows of the tile's segments: fetched now, used after the reduction
            int srow[NSR];
#pragma unroll
            for (int k = 0; k < NSR; k++) srow[k] = (lane + 32 * k < total) ? a.seg_row[sb + lane + 32 * k] : -1;
            T p[NPL];
            unsigned pm = (1u << NPL) - 1u;
#pragma unroll
            for (int j = 0; j < NPL; j++) {
                const int c = band_swz((int)((cur.cw[j >> 1] >> (16 * (j & 1))) & 0xffffu));
                T xv = one_of<T>();
                if (sr.reads_b()) xv = xs[c];
                if (!XFULL) { if (xps[c] == 0) pm &= ~(1u << j); }
                p[j] = sr.mul(sr.reads_a() ? cur.av[j] : one_of<T>(), xv);
            }
            // ---- in-lane walk: segments that start and end inside the run go straight to their staging slot
            T acc = sr.identity(), head = sr.identity();
            int acc_h = 0, head_h = 0, cs = excl - 1;
            bool first = true;
#pragma unroll
            for (int j = 0; j < NPL; j++) {
                if ((f >> j) & 1u) {
                    if (first) { head = acc; head_h = acc_h; first = false; }
                    else { s_val[cs] = acc; if (!XFULL) s_has[cs] = (uint8_t)acc_h; }
                    cs++;
                    acc_h = 0;
                }
                if ((pm >> j) & 1u) {
                    acc = acc_h ? sr.add(acc, p[j]) : p[j];
                    acc_h = 1;
                }
            }
            // ---- segmented scan over lanes: (fl, v, vh) = (run has a start, partial after its last start)
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
            int efh = __shfl_up_sync(0xffffffffu, fl | (vh << 1), 1);
            T ev = shfl_up_any(v, 1);
            if (lane == 0) { efh = 0; ev = sr.identity(); }
            if (nfl > 0 && excl >= 1) {   // the segment in progress when this run began ends at the run's first start
                int eh = efh >> 1;
                pv_combine(sr, ev, eh, head, head_h);
                s_val[excl - 1] = head;
                if (!XFULL) s_has[excl - 1] = (uint8_t)head_h;
            }
            if (lane == 31) {   // the tile's last segment ends with the tile
                s_val[total - 1] = v;
                if (!XFULL) s_has[total - 1] = (uint8_t)vh;
            }
            __syncwarp();
            // ---- one atomic monoid combine per finished segment (dense x: presence comes from the row pointers afterwards)
#pragma unroll
            for (int k = 0; k < NSR; k++) {
                const int i = lane + 32 * k;
                if (i < total && srow[k] >= 0) {
                    bool has = true;
                    if (!XFULL) has = s_has[i] != 0;
                    if (has) {
                        atomic_combine(sr, &a.t_vals[srow[k]], s_val[i]);
                        if (!XFULL) a.t_present[srow[k]] = 1;
                    }
                }
            }
            for (int i = lane + 32 * NSR; i < total; i += 32) {
                const int row = a.seg_row[sb + i];
                if (row < 0) continue;
                if (!XFULL) { if (!s_has[i]) continue; }
                atomic_combine(sr, &a.t_vals[row], s_val[i]);
                if (!XFULL) a.t_present[row] = 1;
            }
            __syncwarp();
            if (tt + BAND_WARPS < t_end) cur = nxt;
        }
        t = t_end;
    }
}

template <typename T> __global__ void band_prefill_kernel(int64_t n, T init, T *__restrict__ t_vals, uint8_t *__restrict__ t_present) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) { t_vals[i] = init; t_present[i] = 0; }
}
// positions that received nothing hold the identity: GraphBLAS stores nothing there, the dense layout keeps T() for tidiness
// rowptr != nullptr (dense x): a row has a result exactly when it has entries
template <typename T> __global__ void band_tidy_kernel(int64_t n, T *__restrict__ t_vals, uint8_t *__restrict__ t_present, const int64_t *__restrict__ rowptr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) {
        bool p;
        if (rowptr) { p = rowptr[i + 1] > rowptr[i]; t_present[i] = p ? 1 : 0; }
        else p = t_present[i] != 0;
        if (!p) t_vals[i] = T();
    }
}

template <typename SR, typename T>
static GrB_Info band_run_typed(const SR &sr, CsrArrays &M, int64_t mrows, int64_t ncols, int64_t nnz, int val_type, const T *x,
                               const uint8_t *xp, T *t_vals, uint8_t *t_present, std::string *err, bool *handled) {
    *handled = false;
    const int cls = sizeof(T) >= 8 ? 1 : 0;
    BandFormat *f = reinterpret_cast<BandFormat *>(M.band_fmt[cls]);
    if (f && f->val_type != val_type) { band_free(f); f = nullptr; M.band_fmt[cls] = nullptr; }
    if (!f) {
        GrB_Info bi = band_build(&f, M, mrows, ncols, nnz, val_type, cls, err);
        if (bi == GrB_NO_VALUE) return GrB_SUCCESS;   // does not apply: the caller falls back
        GRB_TRY(bi);
        M.band_fmt[cls] = f;
    }
    const int pblocks = (int)std::min<int64_t>((mrows + 255) / 256, (int64_t)g_num_sms * 16);
    {
        LAUNCH_NOTE("spmv_band_prefill");
        const T init = (SR::kAddIsAny && !sr.reads_a() && !sr.reads_b()) ? one_of<T>() : sr.identity();
        band_prefill_kernel<T><<<pblocks, 256, 0, g_stream>>>(mrows, init, t_vals, t_present);
    }
    *handled = true;
    if (f->n_tiles == 0) return GrB_SUCCESS;
    BandArgs<T> a;
    a.n_tiles = f->n_tiles; a.ncols = ncols; a.col16 = f->col16; a.vals = (const T *)f->vals; a.flags = f->flags; a.seg_base = f->seg_base;
    a.seg_row = f->seg_row; a.band_tile = f->band_tile; a.tile_band = f->tile_band; a.x = x; a.xp = xp; a.t_vals = t_vals; a.t_present = t_present;
    const int grid = (int)std::min<int64_t>(f->n_tiles, (int64_t)g_num_sms);
    cudaError_t e = cudaSuccess;
    const size_t smem = BandSmem<T>::total(xp == nullptr);
    {
        LAUNCH_NOTE("spmv_band");
        if (xp) {
            auto kern = spmv_band_kernel<SR, T, false>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) kern<<<grid, BAND_THREADS, smem, g_stream>>>(sr, a);
        } else {
            auto kern = spmv_band_kernel<SR, T, true>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) kern<<<grid, BAND_THREADS, smem, g_stream>>>(sr, a);
        }
    }
    if (e == cudaSuccess) {
        LAUNCH_NOTE("spmv_band_tidy");
        band_tidy_kernel<T><<<pblocks, 256, 0, g_stream>>>(mrows, t_vals, t_present, xp ? nullptr : M.ptr);
        e = cudaGetLastError();
    }
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

GrB_Info spmv_band_run(int type_code, int add_op, int mul_op, CsrArrays &M, int64_t mrows, int64_t ncols, int64_t nnz, int val_type,
                       const void *x, const uint8_t *xp, void *t_vals, uint8_t *t_present, std::string *err, bool *handled) {
    *handled = false;
    if (mrows <= 0 || nnz <= 0 || ncols <= 0) return GrB_SUCCESS;
    GrB_Info info = GrB_SUCCESS;
    GRB_DISPATCH_TYPE(type_code, T, {
        GRB_DISPATCH_SEMIRING(add_op, mul_op, T, SRT, sr, {
            info = band_run_typed<SRT, T>(sr, M, mrows, ncols, nnz, val_type, (const T *)x, xp, (T *)t_vals, t_present, err, handled);
        });
    });
    return info;
}
