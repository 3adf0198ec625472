// spgemm_tile.cuh -- row-ordered, TMA-staged hash SpGEMM tiles with decoupled look-back output (included by spgemm.cu).
//
// Why: the binned one-pass kernels hash every row somewhere in a flops-sized staging CSR and a second kernel compacts it
// (20 GB written + 40 GB moved again for the scale-22 product), and their insert loop sits behind the dependent global
// load chain Aj -> Bp -> Bj/Bx of one row.  Here
//   * rows are taken IN ROW ORDER by persistent CTAs (an atomic ticket hands out tiles = runs of consecutive rows whose
//     flop bounds sum to < W + R); a tile's exact output size is known once its rows are hashed, its position in C by a
//     decoupled look-back over the tiles before it -- so the drain of the shared-memory table IS the final write of
//     C's column / value arrays and of C's row pointers: no staging CSR, no compaction pass, no scan;
//   * one PRODUCER warp per CTA runs ahead of the consumer warps: it reads the tile's A entries, looks up the B rows they
//     select and has the TMA unit copy those rows (1-D bulk copies, cp.async.bulk + mbarrier complete_tx) into a ring of
//     shared-memory stages; the CONSUMER warps hash products straight out of shared memory (column / value reads are
//     conflict-free LDS, no global load anywhere in the insert loop), one packed 64-bit atomicCAS per product;
//   * every row of a tile owns a slice of the CTA's table (1.5 x its flop bound), so rows never mix.
// Rows whose bound exceeds R ("holes") are hashed beforehand by the binned kernels into a small staging area; the tile
// kernel only accounts for their (already exact) counts when it assigns row pointers, and place_rows_kernel copies them
// into the holes afterwards.
//
// Bulk copies need 16-byte aligned global sources and sizes: a B row is copied as the aligned superset of its entries
// (up to AL-1 pad slots either side, dead lanes in the insert loop); arrays handed out by dev_alloc carry 64 bytes of
// slack so the last row's superset stays inside its allocation.
#pragma once
#include "tma.cuh"

constexpr int TILE_CH = 64;        // A entries per ring stage (two per producer lane)
constexpr int TILE_RMAX = 256;     // rows per tile
constexpr int TILE_NSTAGE = 2;
constexpr int TILE_UNROLL = 2;
#define LB_FLAG_AGG (1ull << 62)
#define LB_FLAG_PREFIX (2ull << 62)
#define LB_MASK ((1ull << 62) - 1)
enum : int { CHUNK_FIRST = 1, CHUNK_LAST = 2, CHUNK_DONE = 4 };
// phase counters (option spgemm_tile_dbg): consumer warp 0 and the producer warp add their clock64 deltas
enum : int { TD_C_WAIT = 0, TD_C_INSERT, TD_C_COUNT, TD_C_LOOKBACK, TD_C_WRITE, TD_C_BAR, TD_P_WAIT_TILE, TD_P_TILE, TD_P_WAIT_EMPTY, TD_P_LOAD, TD_P_ISSUE, TD_CHUNKS, TD_TILES, TD_N };

struct TileArgs {
    const int64_t *Ap; const int32_t *Aj; const void *Ax;
    const int64_t *Bp; const int32_t *Bj; const void *Bx;
    const int64_t *flops;        // per-row bound
    const int64_t *hole_nnz;     // exact count of rows with flops > R (written by the binned pass), 0 elsewhere
    const int64_t *tile_start;   // n_tiles + 1 row indices
    const int *n_tiles;          // device scalar
    int *ticket;                 // device scalar, zeroed
    unsigned long long *status;  // look-back words, zeroed
    int64_t *Cp; int32_t *Cj; void *Cx;
    int64_t nrows;
    int64_t R;
    int tcap, scap, tf8, cas_first;
    int poll;                    // 1: mbarrier waits poll with test_wait instead of the suspending try_wait
    unsigned long long *dbg;     // optional phase counters (clock cycles), see TILE_DBG_*
};

__device__ __forceinline__ void consumer_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += up;
    }
    return v;
}
// slice of a row with flop bound f: tf8/8 x f slots, a multiple of 32 so that a 32-slot piece of the table lies in one row
__host__ __device__ __forceinline__ int tile_slice_size(int64_t f, int tf8) { return (int)((((f * tf8) >> 3) + 2 + 31) & ~(int64_t)31); }

// shared-memory layout (byte offsets), the same arithmetic on host and device
template <typename T> struct TileLayout {
    static constexpr int AL = sizeof(T) >= 4 ? 4 : 16 / (int)sizeof(T);   // elements per 16-byte bulk-copy granule of the narrower array
    size_t bars, hdr, tinfo, ent, av, tbase, apr, hole, rcnt, rctr, rpos, misc, sj, sx, table, total;
    __host__ __device__ TileLayout(int tcap, int scap, bool packed) {
        size_t o = 0;
        bars = o; o += 8 * (2 * TILE_NSTAGE + 2);
        hdr = o; o += 16 * TILE_NSTAGE;
        tinfo = o; o += 32 * 2;
        o = (o + 15) & ~(size_t)15;
        ent = o; o += (size_t)16 * (TILE_CH + 1) * TILE_NSTAGE;
        av = o; o += ((sizeof(T) * TILE_CH + 15) & ~(size_t)15) * TILE_NSTAGE;
        tbase = o; o += (size_t)4 * (TILE_RMAX + 4) * 2;
        apr = o; o += (size_t)8 * (TILE_RMAX + 2) * 2;
        hole = o; o += (size_t)4 * (TILE_RMAX + 4) * 2;
        rcnt = o; o += (size_t)4 * (TILE_RMAX + 4);
        rctr = o; o += (size_t)4 * (TILE_RMAX + 4);
        o = (o + 7) & ~(size_t)7;
        rpos = o; o += (size_t)8 * (TILE_RMAX + 4);
        misc = o; o += 16;
        o = (o + 127) & ~(size_t)127;
        sj = o; o += (size_t)4 * scap * TILE_NSTAGE;
        o = (o + 127) & ~(size_t)127;
        sx = o; o += sizeof(T) * (size_t)scap * TILE_NSTAGE;
        o = (o + 127) & ~(size_t)127;
        table = o; o += packed ? (size_t)8 * tcap : (((size_t)4 * tcap + 15) & ~(size_t)15) + sizeof(T) * (size_t)tcap;
        total = o;
    }
};

struct TileInfo { long long row0; long long a0; int nrows; int tile; int last; int pad; };

// decoupled look-back (one full warp): publishes this tile's aggregate, returns the sum of all tiles before it and
// publishes the inclusive prefix.  A status word carries flag and value together, so no fences are involved.
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lb_store(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void tile_publish(unsigned long long *status, int tile, long long aggregate) {
    lb_store(status + tile, (tile == 0 ? LB_FLAG_PREFIX : LB_FLAG_AGG) | (unsigned long long)aggregate);
}
// One full warp.  The walk back is a chain of dependent L2 round trips, so every step looks at LB_WIN x 32 tiles at once.
constexpr int LB_WIN = 4;
__device__ __forceinline__ long long tile_lookback(unsigned long long *status, int tile, long long aggregate, int lane) {
    if (tile == 0) return 0;
    long long excl = 0;
    int idx = tile - 1;
    while (true) {
        unsigned long long s[LB_WIN];
        bool pending;
        do {
            pending = false;
#pragma unroll
            for (int k = 0; k < LB_WIN; k++) {
                const int my = idx - lane - 32 * k;
                s[k] = my >= 0 ? lb_load(status + my) : LB_FLAG_PREFIX;
            }
#pragma unroll
            for (int k = 0; k < LB_WIN; k++) pending |= (s[k] >> 62) == 0;
        } while (__any_sync(0xffffffffu, pending));
        bool done = false;
#pragma unroll
        for (int k = 0; k < LB_WIN; k++) {
            if (done) break;
            const unsigned pm = __ballot_sync(0xffffffffu, (s[k] >> 62) == 2);
            const int first = pm ? __ffs(pm) - 1 : 31;
            long long v = lane <= first ? (long long)(s[k] & LB_MASK) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            done = pm != 0;
        }
        if (done) break;
        idx -= 32 * LB_WIN;
    }
    if (lane == 0) lb_store(status + tile, LB_FLAG_PREFIX | (unsigned long long)(excl + aggregate));
    return excl;
}

template <typename SR, typename T, bool PACK>
__global__ void __launch_bounds__(512)
spgemm_tile_kernel(SR sr, TileArgs g) {
    typedef HashTable<SR, T, true, PACK> Table;
    constexpr bool kPacked = Table::kPacked;
    constexpr int AL = TileLayout<T>::AL;
    extern __shared__ __align__(128) unsigned char s_tile_raw[];
    unsigned char *s_raw = s_tile_raw;
    const TileLayout<T> L(g.tcap, g.scap, kPacked);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(s_raw + L.bars);
    uint64_t *bar_empty = bar_full + TILE_NSTAGE;
    uint64_t *bar_tile = bar_empty + TILE_NSTAGE;                            // [2]: the tile that used this slot is drained
    int4 *s_hdr = reinterpret_cast<int4 *>(s_raw + L.hdr);                   // {n_e, P, flags, tslot}
    TileInfo *s_tinfo = reinterpret_cast<TileInfo *>(s_raw + L.tinfo);
    int4 *s_ent = reinterpret_cast<int4 *>(s_raw + L.ent);                   // [stage][CH + 1]: {beg, end, slice base, slice size << 8 | row}
    constexpr size_t AV_STRIDE = ((sizeof(T) * TILE_CH + 15) & ~(size_t)15) / sizeof(T);
    T *s_av = reinterpret_cast<T *>(s_raw + L.av);                           // [stage][CH]
    int *s_tbase = reinterpret_cast<int *>(s_raw + L.tbase);                 // [tslot][RMAX + 4]
    long long *s_apr = reinterpret_cast<long long *>(s_raw + L.apr);         // [tslot][RMAX + 2]
    int *s_hole = reinterpret_cast<int *>(s_raw + L.hole);                   // [tslot][RMAX + 4]
    int *s_rcnt = reinterpret_cast<int *>(s_raw + L.rcnt);                   // new keys per row, counted while inserting
    int *s_rctr = reinterpret_cast<int *>(s_raw + L.rctr);                   // output cursor per row while draining
    long long *s_rpos = reinterpret_cast<long long *>(s_raw + L.rpos);
    long long *s_misc = reinterpret_cast<long long *>(s_raw + L.misc);
    int32_t *s_j = reinterpret_cast<int32_t *>(s_raw + L.sj);                // [stage][scap]
    T *s_x = reinterpret_cast<T *>(s_raw + L.sx);
    unsigned char *s_table = s_raw + L.table;
    unsigned char *s_tvals = s_table + (((size_t)4 * g.tcap + 15) & ~(size_t)15);   // split layout only

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncw = (blockDim.x >> 5) - 1, nct = ncw * 32;
    const int scap = g.scap;

    // ---- one-time set-up: barriers, empty table, zero row counters
    if (tid == 0) {
        for (int s = 0; s < TILE_NSTAGE; s++) { mbar_init(&bar_full[s], 32); mbar_init(&bar_empty[s], ncw); }
        mbar_init(&bar_tile[0], 1);
        mbar_init(&bar_tile[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        Table t0;
        t0.bind(s_table, (unsigned)g.tcap, s_tvals);
        t0.init(sr, tid, blockDim.x);
    }
    for (int r = tid; r < TILE_RMAX + 4; r += blockDim.x) { s_rcnt[r] = 0; s_rctr[r] = 0; }
    __syncthreads();
    const int n_tiles = *g.n_tiles;
    const bool dbg = g.dbg != nullptr && lane == 0 && warp <= 1;   // producer lane 0 and lane 0 of consumer warp 0
    long long dbg_t = dbg ? clock64() : 0;
    unsigned long long dbg_acc[TD_N];
    if (dbg)
        for (int q = 0; q < TD_N; q++) dbg_acc[q] = 0;
#define TILE_DBG(slot_)                                   \
    if (dbg) {                                            \
        const long long now_ = clock64();                 \
        dbg_acc[slot_] += (unsigned long long)(now_ - dbg_t); \
        dbg_t = now_;                                     \
    }

    if (warp == 0) {
        // =========================================================== producer
        int chunk_ctr = 0, tile_ctr = 0;
        for (;;) {
            // ONE ticket per CTA, taken when the previous tile is completely drained: whoever holds a ticket is hashing it, so a
            // look-back only ever waits for tiles that are being worked on.  (Taking tickets ahead -- to prefetch the next tile --
            // lets blocked tiles pile up behind each other: measured 15x to 100x slower.)
            TILE_DBG(TD_P_ISSUE);
            if (tile_ctr > 0) mbar_wait(&bar_tile[(tile_ctr - 1) & 1], (unsigned)((tile_ctr - 1) >> 1) & 1u, g.poll);
            TILE_DBG(TD_P_WAIT_TILE);
            int tile = 0;
            if (lane == 0) tile = atomicAdd(g.ticket, 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            const int tslot = tile_ctr & 1;
            if (tile >= n_tiles) {
                const int slot = chunk_ctr % TILE_NSTAGE;
                mbar_wait(&bar_empty[slot], ((chunk_ctr / TILE_NSTAGE) & 1) ^ 1, g.poll);
                if (lane == 0) s_hdr[slot] = make_int4(0, 0, CHUNK_DONE, 0);
                mbar_arrive(&bar_full[slot]);
                break;
            }
            const long long r0 = g.tile_start[tile], r1 = g.tile_start[tile + 1];
            const int nrows = (int)(r1 - r0);
            const long long a0 = g.Ap[r0], a1 = g.Ap[r1];
            int *tbase = s_tbase + tslot * (TILE_RMAX + 4);
            long long *apr = s_apr + tslot * (TILE_RMAX + 2);
            int *hole = s_hole + tslot * (TILE_RMAX + 4);
            int carry = 0;
            for (int rb = 0; rb < nrows; rb += 32) {
                const int r = rb + lane;
                int ts = 0, hn = 0;
                long long ap = a1 - a0;
                if (r < nrows) {
                    const long long f = g.flops[r0 + r];
                    if (f > g.R) hn = (int)g.hole_nnz[r0 + r];
                    else if (f > 0) ts = tile_slice_size(f, g.tf8);
                    ap = g.Ap[r0 + r] - a0;
                }
                const int incl = warp_incl_scan(ts, lane);
                if (r < nrows) { tbase[r] = carry + incl - ts; apr[r] = ap; hole[r] = hn; }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) {
                if (carry > g.tcap) __trap();   // the host sizes the tiles so that their slices fit the table
                tbase[nrows] = carry;
                apr[nrows] = a1 - a0;
                TileInfo ti;
                ti.row0 = r0; ti.a0 = a0; ti.nrows = nrows; ti.tile = tile; ti.last = (r1 >= g.nrows) ? 1 : 0; ti.pad = 0;
                s_tinfo[tslot] = ti;
            }
            __syncwarp();
            TILE_DBG(TD_P_TILE);
            if (dbg) dbg_acc[TD_TILES]++;
            // ---- chunks of this tile: up to TILE_CH A entries / scap staged slots each; a long B row may span chunks
            long long k = a0;
            long long resume = -1;   // element index in B where the first entry of the window resumes (-1: from its start)
            bool first = true;
            do {
                const int slot = chunk_ctr % TILE_NSTAGE;
                TILE_DBG(TD_P_ISSUE);
                mbar_wait(&bar_empty[slot], ((chunk_ctr / TILE_NSTAGE) & 1) ^ 1, g.poll);
                TILE_DBG(TD_P_WAIT_EMPTY);
                if (dbg) dbg_acc[TD_CHUNKS]++;
                int4 *ent = s_ent + slot * (TILE_CH + 1);
                T *av = s_av + slot * AV_STRIDE;
                long long s_el[2];
                int len[2], pl[2], tb[2], tsr[2];
                bool val[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const long long kk = k + 2 * lane + u;
                    val[u] = kk < a1;
                    s_el[u] = 0; len[u] = 0; pl[u] = 0; tb[u] = 0; tsr[u] = 0;
                    if (val[u]) {
                        const int32_t br = g.Aj[kk];
                        if (sr.reads_a()) av[2 * lane + u] = reinterpret_cast<const T *>(g.Ax)[kk];
                        long long bs = g.Bp[br];
                        const long long be = g.Bp[br + 1];
                        if (2 * lane + u == 0 && resume >= 0) bs = resume;
                        // owning row: the last r with apr[r] <= kk - a0 (rows without entries share their successor's start)
                        const long long x = kk - a0;
                        int lo = 0, hi = nrows - 1;
                        while (lo < hi) {
                            const int mid = (lo + hi + 1) >> 1;
                            if (apr[mid] <= x) lo = mid;
                            else hi = mid - 1;
                        }
                        tb[u] = tbase[lo];
                        const int tsz = tbase[lo + 1] - tb[u];
                        tsr[u] = (tsz << 8) | lo;
                        s_el[u] = bs;
                        len[u] = tsz > 0 ? (int)(be - bs) : 0;   // hole rows (and rows without products) contribute nothing
                        if (len[u] > 0) {
                            const long long as = bs & ~(long long)(AL - 1);
                            pl[u] = (int)(((bs + len[u] + AL - 1) & ~(long long)(AL - 1)) - as);
                        }
                    }
                }
                const int lsum = pl[0] + pl[1];
                const int incl = warp_incl_scan(lsum, lane);
                TILE_DBG(TD_P_LOAD);
                int pos[2];
                pos[0] = incl - lsum;
                pos[1] = pos[0] + pl[0];
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                // whole: fits entirely; part: starts inside the stage and is cut at its end; else: next chunk
                bool whole[2], part[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    whole[u] = val[u] && pos[u] + pl[u] <= scap;
                    part[u] = val[u] && !whole[u] && pos[u] < scap;
                }
                const unsigned w0 = __ballot_sync(0xffffffffu, whole[0]), w1 = __ballot_sync(0xffffffffu, whole[1]);
                const unsigned p0 = __ballot_sync(0xffffffffu, part[0]), p1 = __ballot_sync(0xffffffffu, part[1]);
                const int n_whole = __popc(w0) + __popc(w1);
                const bool any_part = (p0 | p1) != 0;
                const int n_e = n_whole + (any_part ? 1 : 0);
                unsigned tx = 0;
                long long my_resume = -1;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (whole[u] || part[u]) {
                        const long long as = s_el[u] & ~(long long)(AL - 1);
                        const int plu = whole[u] ? pl[u] : scap - pos[u];
                        const int beg = pos[u] + (len[u] > 0 ? (int)(s_el[u] - as) : 0);   // keeps beg monotone over empty entries
                        const int take = whole[u] ? len[u] : (int)(as + plu - s_el[u]);
                        ent[2 * lane + u] = make_int4(beg, beg + take, tb[u], tsr[u]);
                        if (part[u]) my_resume = as + plu;
                        if (plu > 0) tx += (unsigned)plu * 4u + (sr.reads_b() ? (unsigned)plu * (unsigned)sizeof(T) : 0u);
                    }
                }
                // the cut entry (at most one) tells every lane where the next window resumes
                const unsigned pm = p0 | p1;
                long long nresume = -1;
                if (pm) nresume = __shfl_sync(0xffffffffu, my_resume, __ffs(pm) - 1);
                const long long k_next = k + n_whole;
                if (lane == 0) {
                    ent[n_e] = make_int4(0x7fffffff, 0x7fffffff, 0, 0);
                    const int P = total < scap ? total : scap;
                    const bool last = k_next >= a1 && !any_part;
                    s_hdr[slot] = make_int4(n_e, P, (first ? CHUNK_FIRST : 0) | (last ? CHUNK_LAST : 0), tslot);
                }
                if (tx) mbar_arrive_expect_tx(&bar_full[slot], tx);
                else mbar_arrive(&bar_full[slot]);
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (whole[u] || part[u]) {
                        const long long as = s_el[u] & ~(long long)(AL - 1);
                        const int plu = whole[u] ? pl[u] : scap - pos[u];
                        if (plu > 0) {
                            bulk_g2s(s_j + (size_t)slot * scap + pos[u], g.Bj + as, (unsigned)plu * 4u, &bar_full[slot]);
                            if (sr.reads_b())
                                bulk_g2s(s_x + (size_t)slot * scap + pos[u], reinterpret_cast<const T *>(g.Bx) + as, (unsigned)plu * (unsigned)sizeof(T), &bar_full[slot]);
                        }
                    }
                }
                k = k_next;
                resume = nresume;
                first = false;
                chunk_ctr++;
            } while (k < a1);
            tile_ctr++;
        }
    } else {
        // =========================================================== consumers
        const int cw = warp - 1, ctid = tid - 32;
        int chunk_ctr = 0;
        for (;;) {
            const int slot = chunk_ctr % TILE_NSTAGE;
            TILE_DBG(TD_C_WRITE);
            mbar_wait(&bar_full[slot], (chunk_ctr / TILE_NSTAGE) & 1, g.poll);
            TILE_DBG(TD_C_WAIT);
            const int4 h = s_hdr[slot];
            if (h.z & CHUNK_DONE) break;
            const int n_e = h.x, P = h.y;
            const int4 *ent = s_ent + slot * (TILE_CH + 1);
            const T *av = s_av + slot * AV_STRIDE;
            const int32_t *sj = s_j + (size_t)slot * scap;
            const T *sx = s_x + (size_t)slot * scap;
            for (int seg = cw * (32 * TILE_UNROLL); seg < P; seg += ncw * (32 * TILE_UNROLL)) {
                __syncwarp();   // lanes leave the probe loops at different times: bring the warp back together
                int jj[TILE_UNROLL], tb[TILE_UNROLL], tsr[TILE_UNROLL];
                T pr[TILE_UNROLL];
                int p = seg + lane;
                int lo = 0, hi = n_e - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (ent[mid].x <= p) lo = mid;
                    else hi = mid - 1;
                }
#pragma unroll
                for (int u = 0; u < TILE_UNROLL; u++) {
                    jj[u] = HASH_EMPTY;
                    tsr[u] = 0;
                    if (p < P) {
                        while (ent[lo + 1].x <= p) lo++;
                        const int4 e = ent[lo];
                        if (p >= e.x && p < e.y) {
                            jj[u] = sj[p];
                            tb[u] = e.z;
                            tsr[u] = e.w;
                            pr[u] = sr.mul(sr.reads_a() ? av[lo] : one_of<T>(), sr.reads_b() ? sx[p] : one_of<T>());
                        }
                    }
                    p += 32;
                }
#pragma unroll
                for (int u = 0; u < TILE_UNROLL; u++) {
                    int fresh = 0;
                    if (jj[u] != HASH_EMPTY) {
                        Table t;
                        t.bind(s_table + (size_t)tb[u] * (kPacked ? 8 : 4), (unsigned)(tsr[u] >> 8), s_tvals + (size_t)tb[u] * sizeof(T), g.cas_first);
                        fresh = t.insert(sr, jj[u], pr[u]);
                    }
                    __syncwarp();
                    // new keys per row, warp-aggregated: a batch of 32 consecutive products lies in one row, seldom two or three
                    unsigned fm = __ballot_sync(0xffffffffu, fresh != 0);
                    const int row = tsr[u] & 255;
                    while (fm) {
                        const int leader = __ffs(fm) - 1;
                        const int r0 = __shfl_sync(0xffffffffu, row, leader);
                        const unsigned same = __ballot_sync(0xffffffffu, fresh != 0 && row == r0);
                        if (lane == leader) atomicAdd(&s_rcnt[r0], __popc(same));
                        fm &= ~same;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[slot]);
            chunk_ctr++;
            TILE_DBG(TD_C_INSERT);
            if (!(h.z & CHUNK_LAST)) continue;

            // ---- the tile is hashed: its exact size is known, its position comes from the look-back, the drain is the final write
            const int tslot = h.w;
            consumer_bar(nct);
            TILE_DBG(TD_C_BAR);
            const TileInfo ti = s_tinfo[tslot];
            const int *tbase = s_tbase + tslot * (TILE_RMAX + 4);
            const int *hole = s_hole + tslot * (TILE_RMAX + 4);
            const int nrows = ti.nrows;
            if (cw == 0) {
                long long carry = 0;
                for (int rb = 0; rb < nrows; rb += 32) {
                    const int r = rb + lane;
                    long long c = 0;
                    if (r < nrows) {
                        c = tbase[r + 1] - tbase[r] > 0 ? s_rcnt[r] : hole[r];
                        s_rcnt[r] = 0;   // ready for the next tile
                        s_rctr[r] = 0;
                    }
                    long long incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const long long up = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += up;
                    }
                    if (r < nrows) s_rpos[r] = carry + incl - c;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                TILE_DBG(TD_C_COUNT);
                if (lane == 0) tile_publish(g.status, ti.tile, carry);
                const long long base = tile_lookback(g.status, ti.tile, carry, lane);
                if (lane == 0) {
                    s_misc[0] = base;
                    if (ti.last) g.Cp[g.nrows] = base + carry;
                }
                TILE_DBG(TD_C_LOOKBACK);
            }
            consumer_bar(nct);
            TILE_DBG(TD_C_BAR);
            const long long base = s_misc[0];
            for (int r = ctid; r < nrows; r += nct) g.Cp[ti.row0 + r] = base + s_rpos[r];
            // drain: 32-slot pieces of the table go round the consumer warps (slices are multiples of 32 slots, so a piece lies in
            // one row); entries of a row may come out in any order -- the result is "jumbled" anyway -- so a piece only needs a
            // cursor bump of its row, no prefix over the row's other pieces
            T *Cx = reinterpret_cast<T *>(g.Cx);
            const int npieces = tbase[nrows] >> 5;
            for (int c = cw; c < npieces; c += ncw) {
                const int t = (c << 5) + lane;
                int lo = 0, hi = nrows - 1;   // the row whose slice holds slot c * 32: the last r with tbase[r] <= c * 32
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (tbase[mid] <= (c << 5)) lo = mid;
                    else hi = mid - 1;
                }
                int key;
                T v;
                if (kPacked) {
                    unsigned long long *ep = reinterpret_cast<unsigned long long *>(s_table) + t;
                    const unsigned long long e = *ep;
                    key = (int)(e >> 32);
                    v = unpack_value<T>(e);
                    *ep = HASH_EMPTY64;
                } else {
                    int *kp = reinterpret_cast<int *>(s_table) + t;
                    T *vp = reinterpret_cast<T *>(s_tvals) + t;
                    key = *kp;
                    v = *vp;
                    *kp = HASH_EMPTY;
                    *vp = sr.identity();
                }
                const unsigned m = __ballot_sync(0xffffffffu, key >= 0);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    int off = 0;
                    if (lane == leader) off = atomicAdd(&s_rctr[lo], __popc(m));
                    off = __shfl_sync(0xffffffffu, off, leader);
                    if (key >= 0) {
                        const long long pos = base + s_rpos[lo] + off + __popc(m & ((1u << lane) - 1u));
                        g.Cj[pos] = key;
                        Cx[pos] = v;
                    }
                }
            }
            TILE_DBG(TD_C_WRITE);
            consumer_bar(nct);   // the table is clean again: the producer may take the next ticket
            TILE_DBG(TD_C_BAR);
            if (ctid == 0) mbar_arrive(&bar_tile[tslot]);
        }
    }
    if (dbg)
        for (int q = 0; q < TD_N; q++)
            if (dbg_acc[q]) atomicAdd(&g.dbg[q], dbg_acc[q]);
#undef TILE_DBG
}

// ---- tile construction: a row starts a tile when it is the first of its flop window or of a block of TILE_RMAX rows
// small[i] = table slots row i needs in the tile kernel (0: no products, or a hole), big[i] = flop bound of a hole row
__global__ void tile_small_flops_kernel(int64_t nrows, const int64_t *__restrict__ flops, int64_t R, int tf8, int64_t *__restrict__ small,
                                        int64_t *__restrict__ big) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i <= nrows; i += s) {
        const int64_t f = i < nrows ? flops[i] : 0;
        small[i] = (f > 0 && f <= R) ? tile_slice_size(f, tf8) : 0;
        big[i] = f <= R ? 0 : f;
    }
}
__global__ void tile_heads_kernel(int64_t nrows, const int64_t *__restrict__ Fs, int64_t W, uint8_t *__restrict__ head) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += s) head[i] = (i == 0 || (i & (TILE_RMAX - 1)) == 0 || Fs[i] / W != Fs[i - 1] / W) ? 1 : 0;
}
__global__ void tile_sentinel_kernel(int64_t *tile_start, const int *n_tiles, int64_t nrows) { tile_start[*n_tiles] = nrows; }

// staged hole rows -> their final place: one warp per listed row
template <typename T>
__global__ void __launch_bounds__(256)
place_rows_kernel(const int32_t *__restrict__ rows, int64_t n_rows, const int64_t *__restrict__ Sp, const int64_t *__restrict__ Cp,
                  const int64_t *__restrict__ row_nnz, const int32_t *__restrict__ Sj, const T *__restrict__ Sx, int32_t *__restrict__ Cj,
                  T *__restrict__ Cx) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < n_rows; w += nw) {
        const int64_t i = rows[w];
        const int64_t src = Sp[i], dst = Cp[i], n = row_nnz[i];
        int64_t k = lane;
        for (; k + 96 < n; k += 128) {
            const int32_t j0 = __ldcs(Sj + src + k), j1 = __ldcs(Sj + src + k + 32), j2 = __ldcs(Sj + src + k + 64), j3 = __ldcs(Sj + src + k + 96);
            const T x0 = __ldcs(Sx + src + k), x1 = __ldcs(Sx + src + k + 32), x2 = __ldcs(Sx + src + k + 64), x3 = __ldcs(Sx + src + k + 96);
            Cj[dst + k] = j0; Cj[dst + k + 32] = j1; Cj[dst + k + 64] = j2; Cj[dst + k + 96] = j3;
            Cx[dst + k] = x0; Cx[dst + k + 32] = x1; Cx[dst + k + 64] = x2; Cx[dst + k + 96] = x3;
        }
        for (; k < n; k += 32) {
            Cj[dst + k] = __ldcs(Sj + src + k);
            Cx[dst + k] = __ldcs(Sx + src + k);
        }
    }
}
