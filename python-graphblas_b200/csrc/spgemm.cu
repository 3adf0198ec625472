// spgemm.cu -- T = A (+).(x) B for GrB_mxm: flop-binned hash SpGEMM (row-wise Gustavson), sm_100a.
//
//   row_flops : flops(i) = sum_{k in A(i,:)} nnz(B(k,:))   -- upper bound on nnz(T(i,:)); rows are binned by it.
//   one-pass  (default when the flops-sized arrays fit): every row is hashed ONCE and written into its own slots of arrays
//               addressed by the flops prefix, its exact count recorded.  A product that hardly compresses (nnz(T) within 1/8 of
//               the bound: the R-MAT squares) KEEPS those arrays as a row-end CSR (grb_internal.h CsrArrays::end) -- no copy;
//               otherwise a scan and a streaming compaction produce the compact CSR at once.  No symbolic pass at all.
//   two-pass  (memory-tight products, e.g. Graph500 skew): symbolic hash-set pass counts nnz per row, exact allocation, rows
//               re-binned by nnz, numeric pass.
//
// Hash kernels.  Rows with <= 150 products: one warp per row, table in shared memory (spgemm_warp_kernel).  Larger rows, unmasked
// numeric: spgemm_rows_kernel -- persistent CTAs walk the bin's rows; the A row is staged in shared memory a chunk at a time
// together with the start / length of every B row it selects, the chunk's products form one flat index space of which every warp
// owns a contiguous range (128 products per step, one binary search, loads issued before the first insert); an entry is ONE 64-bit
// word (key:value) claimed by a single atomicCAS for value types of <= 4 bytes; products that open a slot append it to a list kept
// in the row's own output slots, and the drain walks that list (no table clear, no table scan).  Masked products, the symbolic
// pass and the two-pass numeric phase use spgemm_block_kernel (CTA per row, clear + scan).  Tables are sized per row (2.5 x count,
// any size: multiply-shift range reduction instead of a power-of-two mask) inside the bin's shared-memory allocation; rows beyond
// the largest shared table are SPLIT over several CTAs by a second hash of the column (each part fits a shared table), or -- under
// a mask -- use global-memory tables.  Rows come out unsorted ("jumbled"); sorting is lazy (structure.cu), as the reference's C
// library allows (graphblas/core/matrix.py:1631-1644).
//
// Serves GrB_mxm: reference graphblas/core/matrix.py:2319-2328 (call assembled at core/base.py:496-503).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <vector>

#include "grb_ops.cuh"

constexpr int HASH_EMPTY = -1;
constexpr unsigned long long HASH_EMPTY64 = ~0ull;
constexpr int MASK_FLAG = (int)0x80000000;   // column indices are < 2^31 - 1, so bit 31 is free
constexpr int LPE = 8;    // lanes per A entry in the warp-per-row kernel
constexpr int MAX_THREADS = 1024;

__device__ __forceinline__ unsigned hash_slot(int key, unsigned size) {
    return (unsigned)(((unsigned long long)((unsigned)key * 0x9E3779B1u) * size) >> 32);   // multiply-shift into [0, size)
}
__device__ __forceinline__ int table_size_for(int64_t cnt, int cap, int tf8) {
    int64_t s = ((cnt * tf8) >> 3) + 8;   // tf8/8 x count
    if (s < 32) s = 32;
    return (int)(s > cap ? cap : s);
}
// heavy rows are split over several CTAs by a second, independent hash of the column (spgemm_block_kernel, `part_offs`)
__device__ __forceinline__ int part_of_key(int key, int parts) {
    return (int)(((unsigned long long)((unsigned)key * 0x85EBCA6Bu) * (unsigned)parts) >> 32);
}
// parts of a row with flop bound c when a part's table holds `maxc` keys: 1/8 head-room for the imbalance of the split
__host__ __device__ __forceinline__ int64_t parts_of(int64_t c, int64_t maxc) { return (c + (c >> 3) + maxc - 1) / maxc; }
template <typename T> struct Packed { static constexpr bool value = sizeof(T) <= 4; };
template <typename T> __device__ __forceinline__ unsigned long long pack_entry(int key, T v) {
    unsigned int bits = 0;
    memcpy(&bits, &v, sizeof(T));
    return ((unsigned long long)(unsigned)key << 32) | bits;
}
template <typename T> __device__ __forceinline__ T unpack_value(unsigned long long e) {
    unsigned int bits = (unsigned int)e;
    T v;
    memcpy(&v, &bits, sizeof(T));
    return v;
}

// ---- the table: symbolic (keys only), numeric packed (key:value in one 64-bit word), numeric split (keys + values)
template <typename SR, typename T, bool NUMERIC, bool PACK> struct HashTable {
    static constexpr bool kPacked = NUMERIC && PACK && Packed<T>::value;
    int *keys;
    T *vals;
    unsigned long long *ent;
    unsigned size;
    int cas_first;   // packed insert: claim with atomicCAS straight away instead of probing with a plain load first
    __device__ __forceinline__ void bind(void *base, unsigned sz, void *vals_base, int casf = 0) {
        size = sz;
        cas_first = casf;
        keys = reinterpret_cast<int *>(base);
        ent = reinterpret_cast<unsigned long long *>(base);
        vals = reinterpret_cast<T *>(vals_base);
    }
    static __host__ __device__ constexpr size_t entry_bytes() { return NUMERIC ? (kPacked ? 8 : 4 + sizeof(T)) : 4; }
    __device__ __forceinline__ void init(const SR &sr, int tid, int nthreads) {
        for (unsigned t = tid; t < size; t += nthreads) {
            if (kPacked) ent[t] = HASH_EMPTY64;
            else {
                keys[t] = HASH_EMPTY;
                if (NUMERIC) vals[t] = sr.identity();
            }
        }
    }
    // returns 1 when the key was new.  GUARD: trap when the table is full instead of probing for ever (split rows size their
    // tables from an expectation, not from an upper bound)
    template <bool GUARD = false>
    __device__ __forceinline__ int insert(const SR &sr, int j, T p) {
        unsigned h = hash_slot(j, size);
        if (kPacked) {
            const unsigned long long mine = pack_entry<T>(j, p);
            if (cas_first) {   // most products of a low-compression row open a new slot: one atomic, no probe load
                // (looking first with a plain load on the later probes and spending the atomic only on an empty slot was measured
                //  slower: 41.1 vs 36.9 ms for the CTA-per-row bins of the scale-22 product -- the extra LDS lengthens the chain)
                unsigned budget = size;   // a full table would probe for ever: trap instead (tables are sized from an upper bound,
                while (true) {            // a split row's part from an expectation with head-room -- this is the safety net)
                    const unsigned long long cur = atomicCAS(&ent[h], HASH_EMPTY64, mine);
                    if (cur == HASH_EMPTY64) return 1;
                    if ((int)(cur >> 32) == j) {
                        atomic_combine(sr, reinterpret_cast<T *>(&ent[h]), p);
                        return 0;
                    }
                    h = (h + 1 == size) ? 0 : h + 1;
                    if (GUARD && --budget == 0) __trap();
                }
            }
            while (true) {
                unsigned long long cur = ent[h];
                if (cur == HASH_EMPTY64) {
                    cur = atomicCAS(&ent[h], HASH_EMPTY64, mine);
                    if (cur == HASH_EMPTY64) return 1;
                }
                if ((int)(cur >> 32) == j) {
                    atomic_combine(sr, reinterpret_cast<T *>(&ent[h]), p);   // value half (little endian: low word)
                    return 0;
                }
                h = (h + 1 == size) ? 0 : h + 1;
            }
        } else {
            unsigned budget = size;
            while (true) {
                int cur = keys[h];
                int fresh = 0;
                if (cur == HASH_EMPTY) {
                    cur = atomicCAS(&keys[h], HASH_EMPTY, j);
                    if (cur == HASH_EMPTY) { fresh = 1; cur = j; }
                }
                if (cur == j) {
                    if (NUMERIC) atomic_combine(sr, &vals[h], p);
                    return fresh;
                }
                h = (h + 1 == size) ? 0 : h + 1;
                if (GUARD && --budget == 0) __trap();
            }
        }
    }
    // insert that also reports WHERE the key lives (for the slot list of spgemm_rows_kernel); returns 1 when the key was new
    __device__ __forceinline__ int insert_at(const SR &sr, int j, T p, unsigned &slot) {
        unsigned h = hash_slot(j, size);
        if (kPacked) {
            const unsigned long long mine = pack_entry<T>(j, p);
            while (true) {
                const unsigned long long cur = atomicCAS(&ent[h], HASH_EMPTY64, mine);
                if (cur == HASH_EMPTY64) { slot = h; return 1; }
                if ((int)(cur >> 32) == j) {
                    atomic_combine(sr, reinterpret_cast<T *>(&ent[h]), p);
                    return 0;
                }
                h = (h + 1 == size) ? 0 : h + 1;
            }
        } else {
            while (true) {
                int cur = keys[h];
                int fresh = 0;
                if (cur == HASH_EMPTY) {
                    cur = atomicCAS(&keys[h], HASH_EMPTY, j);
                    if (cur == HASH_EMPTY) { fresh = 1; cur = j; }
                }
                if (cur == j) {
                    if (NUMERIC) atomic_combine(sr, &vals[h], p);
                    slot = h;
                    return fresh;
                }
                h = (h + 1 == size) ? 0 : h + 1;
            }
        }
    }
    // ---- masked product (C<M> = A*B, M not complemented): the row's table is pre-loaded with the mask row's columns,
    // each tagged MASK_FLAG ("allowed, nothing accumulated yet").  A product whose column is not in the table is dropped
    // before any arithmetic is stored; the first product that hits a column clears the tag (idempotent atomicAnd),
    // values start at the monoid identity so every hit is just an atomic combine.
    __device__ __forceinline__ void insert_mask(const SR &sr, int j) {
        unsigned h = hash_slot(j, size);
        if (kPacked) {
            const unsigned long long mine = pack_entry<T>(j | MASK_FLAG, sr.identity());
            while (true) {
                unsigned long long cur = atomicCAS(&ent[h], HASH_EMPTY64, mine);
                if (cur == HASH_EMPTY64 || ((int)(cur >> 32) & ~MASK_FLAG) == j) return;
                h = (h + 1 == size) ? 0 : h + 1;
            }
        } else {
            while (true) {
                int cur = atomicCAS(&keys[h], HASH_EMPTY, j | MASK_FLAG);
                if (cur == HASH_EMPTY || (cur & ~MASK_FLAG) == j) return;
                h = (h + 1 == size) ? 0 : h + 1;
            }
        }
    }
    // returns 1 when this product is the first to land on its (allowed) column, 0 otherwise (incl. "not in the mask")
    __device__ __forceinline__ int accumulate_masked(const SR &sr, int j, T p) {
        unsigned h = hash_slot(j, size);
        while (true) {
            int *kp = kPacked ? reinterpret_cast<int *>(&ent[h]) + 1 : &keys[h];
            const int cur = *kp;
            if (cur == HASH_EMPTY) return 0;
            if ((cur & ~MASK_FLAG) == j) {
                int fresh = 0;
                if (cur & MASK_FLAG) fresh = (atomicAnd(kp, ~MASK_FLAG) & MASK_FLAG) != 0;
                if (NUMERIC) atomic_combine(sr, kPacked ? reinterpret_cast<T *>(&ent[h]) : &vals[h], p);
                return fresh;
            }
            h = (h + 1 == size) ? 0 : h + 1;
        }
    }
    // ---- complemented mask (C<!M> = A*B): the table is pre-loaded with the mask row's columns exactly as above, but here a
    // tagged column is FORBIDDEN: a product that lands on it is dropped, every other product is inserted as usual (tagged
    // keys are negative, real keys are not, so the two never compare equal).  The unmasked product is never formed.
    // Returns 1 when the key was new.
    __device__ __forceinline__ int insert_comp(const SR &sr, int j, T p) {
        unsigned h = hash_slot(j, size);
        if (kPacked) {
            const unsigned long long mine = pack_entry<T>(j, p);
            while (true) {
                unsigned long long cur = ent[h];
                if (cur == HASH_EMPTY64) {
                    cur = atomicCAS(&ent[h], HASH_EMPTY64, mine);
                    if (cur == HASH_EMPTY64) return 1;
                }
                const int k = (int)(cur >> 32);
                if (k == j) {
                    atomic_combine(sr, reinterpret_cast<T *>(&ent[h]), p);
                    return 0;
                }
                if (k == (j | MASK_FLAG)) return 0;
                h = (h + 1 == size) ? 0 : h + 1;
            }
        } else {
            while (true) {
                int cur = keys[h];
                int fresh = 0;
                if (cur == HASH_EMPTY) {
                    cur = atomicCAS(&keys[h], HASH_EMPTY, j);
                    if (cur == HASH_EMPTY) { fresh = 1; cur = j; }
                }
                if (cur == j) {
                    if (NUMERIC) atomic_combine(sr, &vals[h], p);
                    return fresh;
                }
                if (cur == (j | MASK_FLAG)) return 0;
                h = (h + 1 == size) ? 0 : h + 1;
            }
        }
    }
    // copy every occupied slot to out[*count ...] (count is a shared-memory counter)
    __device__ __forceinline__ void drain(int tid, int nthreads, int *count, int32_t *__restrict__ oj, T *__restrict__ ox) {
        for (unsigned t = tid; t < size; t += nthreads) {
            if (kPacked) {
                const unsigned long long e = ent[t];
                if ((int)(e >> 32) >= 0) {   // neither empty nor an untouched mask column
                    const int pos = atomicAdd(count, 1);
                    oj[pos] = (int)(e >> 32);
                    ox[pos] = unpack_value<T>(e);
                }
            } else {
                const int key = keys[t];
                if (key >= 0) {
                    const int pos = atomicAdd(count, 1);
                    oj[pos] = key;
                    ox[pos] = vals[t];
                }
            }
        }
    }
};

// ------------------------------------------------------------------ flops per row
// (row-end pointers: Ae[i] is where row i stops -- Ap + 1 for a compact CSR, the end array of a row-end CSR; Be likewise)
__global__ void row_flops_kernel(int64_t nrows, const int64_t *__restrict__ Ap, const int64_t *__restrict__ Ae, const int32_t *__restrict__ Aj,
                                 const int64_t *__restrict__ Bp, const int64_t *__restrict__ Be, int64_t *__restrict__ flops) {
    // 8 lanes per row, 4 rows per warp per step; the whole warp runs the same trip count (shuffles inside)
    const int lane = threadIdx.x & 7, sub = (threadIdx.x & 31) >> 3;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 4; base < nrows; base += nwarps * 4) {
        const int64_t i = base + sub;
        int64_t f = 0;
        if (i < nrows)
            for (int64_t k = Ap[i] + lane; k < Ae[i]; k += 8) {
                int32_t r = Aj[k];
                f += Be[r] - Bp[r];
            }
        f += __shfl_down_sync(0xffffffffu, f, 4, 8);
        f += __shfl_down_sync(0xffffffffu, f, 2, 8);
        f += __shfl_down_sync(0xffffffffu, f, 1, 8);
        if (lane == 0 && i < nrows) flops[i] = f;
    }
}

// ------------------------------------------------------------------ binning
constexpr int NBINS = 13;   // 0: empty rows, 1-2: warp per row, 3-11: CTA per row (shared table), 12: global table
constexpr int BIN_LAST_SHARED = NBINS - 2;
struct BinSpec { int64_t maxcount[NBINS]; int cap[NBINS]; int threads[NBINS]; int threads_rows[NBINS]; int tf8[NBINS]; int flags;
                 int64_t split_keys; /* keys a part of a split row is sized for (its table is the largest shared one) */ };

// shared memory of one CTA-per-row block beyond its hash table: per-thread staging of the A-row chunk
static inline size_t block_stage_bytes(int threads, size_t val_bytes) {
    return (size_t)threads * 8 + (size_t)((threads + 4) & ~3) * 4 + (((size_t)threads * val_bytes + 15) & ~(size_t)15);
}

static BinSpec make_bin_spec(size_t entry_bytes) {
    BinSpec s;
    const int caps[NBINS] = {0, 64, 256, 512, 1024, 2048, 4096, 6144, 8192, 12288, 16384, 0, 0};
    // small rows: narrow CTAs (a row has only a few products per thread; many independent CTAs per SM hide the dependent
    // load chain A -> Bp -> Bj of each row); big tables: occupancy is shared-memory bound, so more warps per CTA
    const int thr[NBINS] = {0, 256, 256, 64, 64, 128, 128, 256, 256, 512, 512, 1024, 1024};
    // the persistent slot-list kernel (spgemm_rows_kernel) has no per-row clear / scan to amortise and prefers wider CTAs for the
    // bigger tables: 40.3 -> 37.1 ms per scale-22 step with these (64 / 128-thread CTAs for the 512 / 1024 bins stay best)
    const int thr_rows[NBINS] = {0, 256, 256, 64, 64, 256, 256, 512, 512, 1024, 1024, 1024, 1024};
    int maxcap = (int)((204 * 1024) / entry_bytes);   // 227 KB minus the chunk staging of 1024 threads and the static arrays
    maxcap -= maxcap % 256;
    const int tf_small = std::max(10, (int)opt_get_int("spgemm_table_factor8", 20));     // table = tf8/8 x count
    const int tf_big = std::max(10, (int)opt_get_int("spgemm_table_factor8_big", tf_small));
    const int big_from = (int)opt_get_int("spgemm_big_from_bin", 8);
    for (int b = 0; b < NBINS; b++) {
        s.cap[b] = caps[b];
        s.threads[b] = thr[b];
        s.threads_rows[b] = thr_rows[b];
        s.tf8[b] = b >= big_from ? tf_big : tf_small;
        char key[32];
        snprintf(key, sizeof key, "spgemm_thr_%d", b);
        const long t = opt_get_int(key, 0);
        if (b >= 3 && t >= 64 && t <= 1024 && t % 32 == 0) s.threads[b] = s.threads_rows[b] = (int)t;
    }
    s.cap[BIN_LAST_SHARED] = maxcap > 16384 + 2048 ? maxcap : 16384;
    s.maxcount[0] = 0;
    for (int b = 1; b <= BIN_LAST_SHARED; b++) {
        s.maxcount[b] = (((int64_t)s.cap[b] - 8) * 8) / s.tf8[b] - 1;   // tf8/8*count + 8 <= cap
        if (s.maxcount[b] < s.maxcount[b - 1]) s.maxcount[b] = s.maxcount[b - 1];   // a tighter factor must not reorder the bins
    }
    s.maxcount[NBINS - 1] = INT64_MAX;
    // a part of a split row fills its table to about split_load8 / 8 / 1.125 = 0.56: fewer parts mean fewer CTAs that each re-read
    // the whole row (measured: 3.59 ms at the bins' own 0.4, 3.32 at 0.44, 3.05 at 0.56); parts_of() keeps 1/8 head-room for the
    // imbalance of the hash split, and the insert traps rather than spins should a table ever fill up
    s.split_keys = ((int64_t)s.cap[BIN_LAST_SHARED] * std::min<long>(6, std::max<long>(2, opt_get_int("spgemm_split_load8", 5)))) / 8;
    s.flags = opt_get_int("spgemm_cas_first", 1) != 0 ? 1 : 0;
    return s;
}
__device__ __forceinline__ int bin_of(const BinSpec &s, int64_t c) {
    int b = 0;
#pragma unroll
    for (int q = 0; q < NBINS - 1; q++) b += (c > s.maxcount[q]);
    return b;
}
__host__ __device__ __forceinline__ int64_t gtable_size_of(int64_t c) { return (2 * c + 1024) & ~(int64_t)3; }   // multiple of 4 keeps the value array 16-byte aligned
// bin_counts[0..NBINS) = rows per bin; bin_counts[NBINS] = total global-table entries the rows of the last bin need
// bin_counts[NBINS + 1] = CTAs the rows of the last bin need when each is split into parts that fit the largest shared table
__global__ void bin_count_kernel(BinSpec spec, int64_t nrows, const int64_t *__restrict__ cnt, unsigned long long *__restrict__ bin_counts) {
    __shared__ unsigned int s[NBINS];
    __shared__ unsigned long long s_g, s_p;
    if (threadIdx.x < NBINS) s[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_g = 0; s_p = 0; }
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += stride) {
        const int64_t c = cnt[i];
        const int b = bin_of(spec, c);
        atomicAdd(&s[b], 1u);
        if (b == NBINS - 1) {
            atomicAdd(&s_g, (unsigned long long)gtable_size_of(c));
            atomicAdd(&s_p, (unsigned long long)parts_of(c, spec.split_keys));
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS && s[threadIdx.x]) atomicAdd(&bin_counts[threadIdx.x], (unsigned long long)s[threadIdx.x]);
    if (threadIdx.x == 0 && s_g) { atomicAdd(&bin_counts[NBINS], s_g); atomicAdd(&bin_counts[NBINS + 1], s_p); }
}
struct BinCursors { unsigned long long v[NBINS]; };
__global__ void bin_cursors_kernel(BinCursors c, unsigned long long *__restrict__ cursors) {
    if (threadIdx.x < NBINS) cursors[threadIdx.x] = c.v[threadIdx.x];
}
// two-level: CTA-local histogram + one global atomic per (CTA, bin) reserves a contiguous range
__global__ void bin_fill_kernel(BinSpec spec, int64_t nrows, const int64_t *__restrict__ cnt, unsigned long long *__restrict__ cursors,
                                int32_t *__restrict__ bin_rows) {
    __shared__ unsigned int s_cnt[NBINS];
    __shared__ unsigned long long s_base[NBINS];
    const int64_t per_block = ((nrows + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per_block;
    const int64_t hi = lo + per_block < nrows ? lo + per_block : nrows;
    if (threadIdx.x < NBINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        int b = bin_of(spec, cnt[i]);
        if (b) atomicAdd(&s_cnt[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NBINS) {
        s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
        s_cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        int b = bin_of(spec, cnt[i]);
        if (b) bin_rows[s_base[b] + atomicAdd(&s_cnt[b], 1u)] = (int32_t)i;
    }
}

struct MaskArgs { const int64_t *Mp; const int32_t *Mj; const uint8_t *Meff; int comp; };   // Mp == nullptr: unmasked; comp: the mask row FORBIDS its columns

// ------------------------------------------------------------------ warp-per-row kernel (tiny rows)
template <typename SR, typename T, bool NUMERIC, bool PACK>
__global__ void __launch_bounds__(256)
spgemm_warp_kernel(SR sr, const int32_t *__restrict__ rows, int64_t n_rows, int cap, int tf8, const int64_t *__restrict__ cnt,
                   const int64_t *__restrict__ Ap, const int64_t *__restrict__ Ae, const int32_t *__restrict__ Aj, const T *__restrict__ Ax,
                   const int64_t *__restrict__ Bp, const int64_t *__restrict__ Be, const int32_t *__restrict__ Bj, const T *__restrict__ Bx,
                   int64_t *__restrict__ row_nnz, const int64_t *__restrict__ Op, int32_t *__restrict__ Oj, T *__restrict__ Ox,
                   MaskArgs mk) {
    typedef HashTable<SR, T, NUMERIC, PACK> Table;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane32 = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const bool active = g < n_rows;
    const size_t per = (((size_t)cap * Table::entry_bytes() + 8) + 15) & ~(size_t)15;
    unsigned char *base = s_raw + per * wib;
    int *count = reinterpret_cast<int *>(base);
    Table tab;
    const int64_t row = active ? rows[g] : 0;
    const unsigned tsize = active ? (unsigned)table_size_for(cnt[row], cap, tf8) : 32u;
    tab.bind(base + 8, tsize, base + 8 + (size_t)cap * 4);
    if (active) tab.init(sr, lane32, 32);
    if (lane32 == 0) *count = 0;
    __syncwarp();
    if (active && mk.Mp) {
        for (int64_t k = mk.Mp[row] + lane32; k < mk.Mp[row + 1]; k += 32)
            if (!mk.Meff || mk.Meff[k]) tab.insert_mask(sr, mk.Mj[k]);
    }
    __syncwarp();
    int local_new = 0;
    if (active) {
        const int sub = lane32 / LPE, lane = lane32 % LPE;
        const int64_t a_beg = Ap[row], a_end = Ae[row];
        for (int64_t k = a_beg + sub; k < a_end; k += 32 / LPE) {
            const int32_t br = Aj[k];
            T a = one_of<T>();
            if (NUMERIC && sr.reads_a()) a = Ax[k];
            const int64_t b_beg = Bp[br], b_end = Be[br];
            for (int64_t q = b_beg + lane; q < b_end; q += LPE) {
                const int j = Bj[q];
                T p = T();
                if (NUMERIC) p = sr.mul(a, sr.reads_b() ? Bx[q] : one_of<T>());
                local_new += mk.Mp ? (mk.comp ? tab.insert_comp(sr, j, p) : tab.accumulate_masked(sr, j, p)) : tab.insert(sr, j, p);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_new += __shfl_down_sync(0xffffffffu, local_new, o);
    if (active) {
        if (lane32 == 0 && row_nnz) row_nnz[row] = local_new;
        if (NUMERIC) {
            const int64_t ob = Op[row];
            tab.drain(lane32, 32, count, Oj + ob, Ox + ob);
        }
    }
}

// ------------------------------------------------------------------ CTA-per-row kernel, products flattened per thread
// The A row is staged in shared memory a chunk (blockDim entries) at a time: B-row start, A value and the exclusive
// prefix of the B-row lengths.  The chunk's products form one index space [0, P); thread t takes products
// t, t+blockDim, ... (consecutive lanes -> consecutive entries of the same B row -> coalesced loads), finds the
// owning A entry with a binary search on the prefix (lanes of a warp follow almost the same search path, so the
// shared-memory reads broadcast), issues UNROLL independent loads, then inserts.  Every lane has work regardless of
// how short the B rows are.  GLOBAL selects a global-memory table (rows too big for shared memory); it is a
// template parameter so that the shared-memory instantiation compiles to ATOMS/LDS/STS rather than generic atomics.
constexpr int UNROLL = 4;
#ifndef SPGEMM_PIPE
#define SPGEMM_PIPE 0   // measured: 37.74 (pipelined) vs 37.75 ms (plain) for the CTA bins of the scale-22 product, and 12 more registers
#endif
template <typename SR, typename T, bool NUMERIC, bool PACK, bool GLOBAL, bool SPLIT = false>
__global__ void __launch_bounds__(MAX_THREADS) spgemm_block_kernel(SR sr, const int32_t *__restrict__ rows, int cap, int tf8, int flags, const int64_t *__restrict__ cnt,
                                    const int64_t *__restrict__ Ap, const int64_t *__restrict__ Ae, const int32_t *__restrict__ Aj, const T *__restrict__ Ax,
                                    const int64_t *__restrict__ Bp, const int64_t *__restrict__ Be, const int32_t *__restrict__ Bj, const T *__restrict__ Bx,
                                    int64_t *__restrict__ row_nnz, const int64_t *__restrict__ Op, int32_t *__restrict__ Oj,
                                    T *__restrict__ Ox, unsigned char *g_table, const int64_t *__restrict__ g_offsets, MaskArgs mk,
                                    const int64_t *__restrict__ part_offs, int n_split_rows, int64_t part_maxc) {
    typedef HashTable<SR, T, NUMERIC, PACK> Table;
    // dynamic shared memory: [hash table: cap entries (none when GLOBAL)] [s_bs: B-row starts] [s_off: product prefix] [s_av: A values]
    // -- the staging arrays are sized by blockDim, so a 128-thread CTA of a small bin costs 2 KB, not the 16 KB of 1024 threads
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_wsum[MAX_THREADS / 32];
    __shared__ int s_count, s_new;
    __shared__ long long s_base;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const size_t tbytes = GLOBAL ? 0 : (((size_t)cap * Table::entry_bytes() + 15) & ~(size_t)15);
    int64_t *s_bs = reinterpret_cast<int64_t *>(s_raw + tbytes);
    int *s_off = reinterpret_cast<int *>(s_bs + nthreads);
    T *s_av = reinterpret_cast<T *>(s_off + ((nthreads + 4) & ~3));
    // SPLIT rows (part_offs != nullptr; unmasked, shared tables only): row rows[r] is hashed by part_offs[r + 1] - part_offs[r]
    // CTAs; CTA `part` of a row reads ALL of the row's products (B rows are streamed from L2) but keeps only the columns whose
    // part hash is `part`, so the pieces are disjoint, each fits a shared-memory table, and a heavy row no longer crawls through
    // a global-memory table on one SM.  Every CTA reserves its piece of the row's output with one atomic on the row's counter
    // (the order of a row's entries is free: results are "jumbled" anyway).
    int parts = 1, part = 0;
    int64_t row;
    if (SPLIT) {
        int lo = 0, hi = n_split_rows - 1;
        const int64_t me = blockIdx.x;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (part_offs[mid] <= me) lo = mid;
            else hi = mid - 1;
        }
        row = rows[lo];
        parts = (int)(part_offs[lo + 1] - part_offs[lo]);
        part = (int)(me - part_offs[lo]);
    } else {
        row = rows[blockIdx.x];
    }

    Table tab;
    if (GLOBAL) {   // global-memory table: [offset, offset + size) entries of this row
        const int64_t off = g_offsets[blockIdx.x];
        const unsigned sz = (unsigned)(g_offsets[blockIdx.x + 1] - off);
        const int64_t total = g_offsets[gridDim.x];
        unsigned char *kbase = g_table + (size_t)off * (Table::kPacked ? 8 : 4);
        unsigned char *vbase = g_table + (size_t)total * 4 + (size_t)off * sizeof(T);
        tab.bind(kbase, sz, vbase, flags & 1);
    } else {
        // a part expects 1 / parts of the row's keys; parts_of() left 1/8 head-room below part_maxc
        const int64_t c_eff = SPLIT ? part_maxc : cnt[row];
        tab.bind(s_raw, (unsigned)table_size_for(c_eff, cap, tf8), s_raw + (size_t)cap * 4, flags & 1);
    }
    tab.init(sr, tid, nthreads);
    if (tid == 0) { s_count = 0; s_new = 0; }
    if (mk.Mp) {
        __syncthreads();
        for (int64_t k = mk.Mp[row] + tid; k < mk.Mp[row + 1]; k += nthreads)
            if (!mk.Meff || mk.Meff[k]) tab.insert_mask(sr, mk.Mj[k]);
    }

    const int wlane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int64_t a_beg = Ap[row], a_end = Ae[row];
    int local_new = 0;
    for (int64_t c0 = a_beg; c0 < a_end; c0 += nthreads) {
        const int chunk_n = (int)((a_end - c0 < nthreads) ? (a_end - c0) : nthreads);
        int len = 0;
        if (tid < chunk_n) {
            const int64_t k = c0 + tid;
            const int32_t br = Aj[k];
            const int64_t bs = Bp[br];
            len = (int)(Be[br] - bs);
            s_bs[tid] = bs;
            if (NUMERIC && sr.reads_a()) s_av[tid] = Ax[k];
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (wlane >= o) incl += up;
        }
        if (wlane == 31) s_wsum[warp] = incl;
        __syncthreads();   // also orders the table init / the previous chunk's inserts before this chunk's
        int base = 0;
        for (int w = 0; w < warp; w++) base += s_wsum[w];
        s_off[tid] = base + incl - len;
        if (tid == nthreads - 1) s_off[nthreads] = base + incl;
        __syncthreads();
        const int P = s_off[nthreads];
        // A warp takes 32*UNROLL consecutive products per step: lane l handles products seg + l, seg + l + 32, ... -- consecutive
        // lanes read consecutive entries of one B row (coalesced), and the owning A entry is found by ONE binary search per
        // step; the later products of the lane lie 32 further on, a short forward walk over the prefix (s_off[nthreads] = P is
        // the sentinel, entries beyond chunk_n hold P as well).
        // SPGEMM_PIPE=1: software pipeline -- the loads of the warp's NEXT step are issued before the inserts of the current one, so
        // the L2 round trip of Bj / Bx overlaps the shared-memory atomics (the compiler cannot move loads across them).  Measured
        // neutral (the SM is short of warps, not of loads in flight: 6 CTAs x 4 warps, 13 cycles per issued instruction), so off.
        // Every warp owns ONE contiguous range of the chunk's products (P / nwarps, rounded up to whole warp rows of 32) and walks it
        // in steps of 32 * UNROLL: all warps do the same amount of work to within 32 products.  (Handing out 128-product steps
        // round-robin left a 1 100-product row as 3 + 2 + 2 + 2 steps over four warps: ncu showed 20 % of the samples of the
        // largest bin waiting at the barrier below, profiles/ncu_full_spgemm_r02.txt.)
        const int per_warp = (((P + nwarps - 1) / nwarps) + 31) & ~31;
        const int w_beg = warp * per_warp;
        const int w_end = w_beg + per_warp < P ? w_beg + per_warp : P;
        auto load_step = [&](int seg, int (&jj)[UNROLL], T (&bb)[UNROLL], T (&aa)[UNROLL]) {
            int p = seg + wlane;
            int lo = 0;
            if (p < w_end) {
                int hi = chunk_n - 1;   // last entry e with s_off[e] <= p (zero-length rows are skipped)
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_off[mid] <= p) lo = mid;
                    else hi = mid - 1;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                jj[u] = HASH_EMPTY;
                if (p < w_end) {
                    while (s_off[lo + 1] <= p) lo++;
                    const int64_t q = s_bs[lo] + (p - s_off[lo]);
                    jj[u] = Bj[q];
                    if (NUMERIC && sr.reads_b()) bb[u] = Bx[q];
                    if (NUMERIC && sr.reads_a()) aa[u] = s_av[lo];
                }
                p += 32;
            }
        };
        auto insert_step = [&](const int (&jj)[UNROLL], const T (&bb)[UNROLL], const T (&aa)[UNROLL]) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (jj[u] == HASH_EMPTY) continue;
                if (SPLIT && part_of_key(jj[u], parts) != part) continue;
                T pr = T();
                if (NUMERIC) pr = sr.mul(sr.reads_a() ? aa[u] : one_of<T>(), sr.reads_b() ? bb[u] : one_of<T>());
                local_new += mk.Mp ? (mk.comp ? tab.insert_comp(sr, jj[u], pr) : tab.accumulate_masked(sr, jj[u], pr)) : tab.template insert<SPLIT>(sr, jj[u], pr);
            }
        };
        const int step = 32 * UNROLL;
#if SPGEMM_PIPE
        {
            int jj[UNROLL], nj[UNROLL];
            T bb[UNROLL], aa[UNROLL], nb[UNROLL], na[UNROLL];
            int seg = w_beg;
            if (seg < w_end) load_step(seg, jj, bb, aa);
            while (seg < w_end) {
                const int nseg = seg + step;
                const bool more = nseg < w_end;   // warp-uniform
                if (more) load_step(nseg, nj, nb, na);
                insert_step(jj, bb, aa);
                if (more) {
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) { jj[u] = nj[u]; bb[u] = nb[u]; aa[u] = na[u]; }
                }
                seg = nseg;
            }
        }
#else
        for (int seg = w_beg; seg < w_end; seg += step) {
            int jj[UNROLL];
            T bb[UNROLL], aa[UNROLL];
            load_step(seg, jj, bb, aa);
            insert_step(jj, bb, aa);
        }
#endif
        __syncthreads();   // s_* arrays are rewritten by the next chunk
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_new += __shfl_down_sync(0xffffffffu, local_new, o);
    if (SPLIT) {
        // this CTA's piece: its size is the number of keys it opened; one atomic on the row's counter places it
        if (wlane == 0 && local_new) atomicAdd(&s_new, local_new);
        __syncthreads();
        if (tid == 0) s_base = (long long)atomicAdd(reinterpret_cast<unsigned long long *>(&row_nnz[row]), (unsigned long long)s_new);
        __syncthreads();
        if (NUMERIC) {
            const int64_t ob = Op[row] + s_base;
            tab.drain(tid, nthreads, &s_count, Oj + ob, Ox + ob);
        }
        return;
    }
    if (NUMERIC) {
        const int64_t ob = Op[row];
        tab.drain(tid, nthreads, &s_count, Oj + ob, Ox + ob);
        if (row_nnz) {
            __syncthreads();
            if (tid == 0) row_nnz[row] = s_count;
        }
    } else {
        if (wlane == 0 && local_new) atomicAdd(&s_count, local_new);
        __syncthreads();
        if (tid == 0) row_nnz[row] = s_count;
    }
}

// ------------------------------------------------------------------ persistent CTA-per-row kernel with a slot list (unmasked numeric)
// The CTA-per-row kernel above pays for its table twice per row: it clears all 2.5 x count slots before and scans all of them
// after the inserts (ncu, 4096-slot bin: 43.7 M warp-iterations each for 18.6 M insert steps; the scan with its warp-aggregated
// compaction is 15 % of the kernel).  Here a CTA stays resident and walks rows rows[b], rows[b + G], ...; every product that
// OPENS a slot appends the slot's index to a list -- kept in the row's own, still unused output slots Oj[Op[row] ...] (one
// coalesced 128-byte store per warp step, L2 resident) -- so the drain touches exactly the used slots: entry -> final (column,
// value) at the list position, fully coalesced, and the slot is reset to EMPTY on the way, which leaves the table clean for the
// CTA's next row: no clear pass, no scan, no per-slot compaction atomics.
template <typename SR, typename T, bool PACK>
__global__ void __launch_bounds__(MAX_THREADS)
spgemm_rows_kernel(SR sr, const int32_t *__restrict__ rows, int n_rows, int cap, int tf8, const int64_t *__restrict__ cnt,
                   const int64_t *__restrict__ Ap, const int64_t *__restrict__ Ae, const int32_t *__restrict__ Aj, const T *__restrict__ Ax,
                   const int64_t *__restrict__ Bp, const int64_t *__restrict__ Be, const int32_t *__restrict__ Bj, const T *__restrict__ Bx,
                   int64_t *__restrict__ row_nnz, const int64_t *__restrict__ Op, int32_t *Oj, T *Ox) {
    typedef HashTable<SR, T, true, PACK> Table;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_wsum[MAX_THREADS / 32];
    __shared__ int s_count[2];   // list length of the row in flight; the two slots alternate per row (see the end of the row loop)
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const size_t tbytes = (((size_t)cap * Table::entry_bytes() + 15) & ~(size_t)15);
    int64_t *s_bs = reinterpret_cast<int64_t *>(s_raw + tbytes);
    int *s_off = reinterpret_cast<int *>(s_bs + nthreads);
    T *s_av = reinterpret_cast<T *>(s_off + ((nthreads + 4) & ~3));
    const int wlane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const unsigned lt_mask = (1u << wlane) - 1u;

    Table tab;
    tab.bind(s_raw, (unsigned)cap, s_raw + (size_t)cap * 4, 1);
    tab.init(sr, tid, nthreads);   // the only full clear: every row's drain leaves the table empty again
    if (tid == 0) { s_count[0] = 0; s_count[1] = 0; }
    __syncthreads();

    int par = 0;
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x, par ^= 1) {
        const int64_t row = rows[r];
        tab.bind(s_raw, (unsigned)table_size_for(cnt[row], cap, tf8), s_raw + (size_t)cap * 4, 1);
        const int64_t ob = Op[row];
        int32_t *lj = Oj + ob;
        const int64_t a_beg = Ap[row], a_end = Ae[row];
        for (int64_t c0 = a_beg; c0 < a_end; c0 += nthreads) {
            const int chunk_n = (int)((a_end - c0 < nthreads) ? (a_end - c0) : nthreads);
            int len = 0;
            if (tid < chunk_n) {
                const int64_t k = c0 + tid;
                const int32_t br = Aj[k];
                const int64_t bs = Bp[br];
                len = (int)(Be[br] - bs);
                s_bs[tid] = bs;
                if (sr.reads_a()) s_av[tid] = Ax[k];
            }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (wlane >= o) incl += up;
            }
            if (wlane == 31) s_wsum[warp] = incl;
            __syncthreads();
            if (tid == 0 && c0 == a_beg) s_count[par ^ 1] = 0;   // the previous row's list length: everybody has read it by now
            int base = 0;
            for (int w = 0; w < warp; w++) base += s_wsum[w];
            s_off[tid] = base + incl - len;
            if (tid == nthreads - 1) s_off[nthreads] = base + incl;
            __syncthreads();
            const int P = s_off[nthreads];
            const int per_warp = (((P + nwarps - 1) / nwarps) + 31) & ~31;   // one contiguous range of the chunk's products per warp
            const int w_beg = warp * per_warp;
            const int w_end = w_beg + per_warp < P ? w_beg + per_warp : P;
            for (int seg = w_beg; seg < w_end; seg += 32 * UNROLL) {
                int jj[UNROLL];
                T bb[UNROLL], aa[UNROLL];
                int p = seg + wlane;
                int lo = 0;
                if (p < w_end) {
                    int hi = chunk_n - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (s_off[mid] <= p) lo = mid;
                        else hi = mid - 1;
                    }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    jj[u] = HASH_EMPTY;
                    if (p < w_end) {
                        while (s_off[lo + 1] <= p) lo++;
                        const int64_t q = s_bs[lo] + (p - s_off[lo]);
                        jj[u] = Bj[q];
                        if (sr.reads_b()) bb[u] = Bx[q];
                        if (sr.reads_a()) aa[u] = s_av[lo];
                    }
                    p += 32;
                }
                unsigned slot[UNROLL], fm[UNROLL];
                int fresh[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    fresh[u] = 0;
                    slot[u] = 0;
                    if (jj[u] != HASH_EMPTY) {
                        const T pr = sr.mul(sr.reads_a() ? aa[u] : one_of<T>(), sr.reads_b() ? bb[u] : one_of<T>());
                        fresh[u] = tab.insert_at(sr, jj[u], pr, slot[u]);
                    }
                }
                __syncwarp();
                // one reservation of list positions per warp step; new slots of the step go out as (at most UNROLL) coalesced runs
                int total = 0;
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    fm[u] = __ballot_sync(0xffffffffu, fresh[u] != 0);
                    total += __popc(fm[u]);
                }
                if (total) {
                    int pos = 0;
                    if (wlane == 0) pos = atomicAdd(&s_count[par], total);
                    pos = __shfl_sync(0xffffffffu, pos, 0);
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) {
                        if (fresh[u]) lj[pos + __popc(fm[u] & lt_mask)] = (int32_t)slot[u];
                        pos += __popc(fm[u]);
                    }
                }
            }
            __syncthreads();   // s_* arrays are rewritten by the next chunk; after the last chunk: every insert and list store is done
        }
        const int count = s_count[par];
        // drain by the list: position i of the row's output receives the entry of slot list[i]; the slot is emptied
        // (two list reads in flight per thread: they come back from L2, and 10 % of the samples sat on that load issued singly)
        T *lx = Ox + ob;
        for (int i0 = tid; i0 < count; i0 += 2 * nthreads) {
            const int i1 = i0 + nthreads;
            const unsigned sl0 = (unsigned)lj[i0];
            const unsigned sl1 = i1 < count ? (unsigned)lj[i1] : 0u;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i = u ? i1 : i0;
                const unsigned sl = u ? sl1 : sl0;
                if (i >= count) break;
                if (Table::kPacked) {
                    const unsigned long long e = tab.ent[sl];
                    tab.ent[sl] = HASH_EMPTY64;
                    lj[i] = (int)(e >> 32);
                    lx[i] = unpack_value<T>(e);
                } else {
                    lj[i] = tab.keys[sl];
                    lx[i] = tab.vals[sl];
                    tab.keys[sl] = HASH_EMPTY;
                    tab.vals[sl] = sr.identity();
                }
            }
        }
        if (row_nnz && tid == 0) row_nnz[row] = count;
        // No barrier here: the next row touches the table and the list counter only after the two barriers of its staging, and a
        // thread reaches those only after its share of this drain.  The counter of THIS row is reset by thread 0 after the next
        // row's first staging barrier (below), when every thread has read it; the next row counts in the other slot.
    }
    // Tried and dropped: prefetching the next row's descriptor, A entries and B-row extents into registers while the current row
    // is hashed (the four dependent round trips of a row's set-up, 12 % of the ncu samples).  It needs 21 more registers: at 53
    // the 256-thread CTAs lose a third of their occupancy and the step went 37.7 -> 42.1 ms; capped at 40 registers 40.1 ms.
    // Also dropped: a single-barrier staging path for chunks of <= 32 A entries (warp 0's shuffle scan as the prefix): 35.8 -> 37.7 ms.
}

// parts per listed row (sizes[n]: pad slot of the in-place exclusive scan)
__global__ void split_parts_kernel(const int32_t *__restrict__ rows, int64_t n, const int64_t *__restrict__ cnt, int64_t maxc,
                                   int64_t *__restrict__ sizes) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    sizes[i] = i < n ? parts_of(cnt[rows[i]], maxc) : 0;
}

__global__ void gtable_sizes_kernel(const int32_t *__restrict__ rows, int64_t n, const int64_t *__restrict__ cnt,
                                    int64_t *__restrict__ sizes) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    sizes[i] = i < n ? gtable_size_of(cnt[rows[i]]) : 0;   // sizes[n]: pad slot of the in-place exclusive scan
}
__global__ void i64_copy_kernel(int64_t *dst, const int64_t *src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = src[i];
}
// out[0] = sum, out[1] = max, out[2] = sum of the values above R, out[3] = how many are above R
__global__ void reduce_sum_max_kernel(const int64_t *__restrict__ v, int64_t n, int64_t R, unsigned long long *__restrict__ out) {
    unsigned long long sum = 0, mx = 0, big = 0, nbig = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) {
        const unsigned long long x = (unsigned long long)v[i];
        sum += x;
        mx = x > mx ? x : mx;
        if ((int64_t)x > R) { big += x; nbig++; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        big += __shfl_down_sync(0xffffffffu, big, o);
        nbig += __shfl_down_sync(0xffffffffu, nbig, o);
        unsigned long long om = __shfl_down_sync(0xffffffffu, mx, o);
        mx = om > mx ? om : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], sum);
        atomicMax(&out[1], mx);
        if (nbig) { atomicAdd(&out[2], big); atomicAdd(&out[3], nbig); }
    }
}
// masked product: the table holds the mask row, the work is the row's flops; bin by whichever asks for the bigger CTA
__global__ void masked_count_kernel(int64_t nrows, const int64_t *__restrict__ flops, const int64_t *__restrict__ Mp,
                                    int64_t work_cap, int64_t *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += s) {
        const int64_t mn = Mp[i + 1] - Mp[i], f = flops[i];
        int64_t w = f >> 5;
        if (w > work_cap) w = work_cap;
        cnt[i] = (mn > 0 && f > 0) ? (mn > w ? mn : w) : 0;
    }
}
// complemented mask: a row's table holds its forbidden columns AND its products
__global__ void comp_count_kernel(int64_t nrows, const int64_t *__restrict__ flops, const int64_t *__restrict__ Mp, int64_t *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i <= nrows; i += s) {
        const int64_t f = i < nrows ? flops[i] : 0;
        cnt[i] = f > 0 ? f + (Mp[i + 1] - Mp[i]) : 0;
    }
}
// staging (addressed by the flops prefix) -> final CSR: one warp per row, coalesced both ways, 4 loads in flight
template <typename T>
__global__ void __launch_bounds__(256)
compact_rows_kernel(int64_t nrows, const int64_t *__restrict__ Sp, const int64_t *__restrict__ Cp,
                    const int32_t *__restrict__ Sj, const T *__restrict__ Sx, int32_t *__restrict__ Cj, T *__restrict__ Cx) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w; i < nrows; i += nw) {
        const int64_t src = Sp[i], dst = Cp[i], n = Cp[i + 1] - dst;
        int64_t k = lane;
        for (; k + 224 < n; k += 256) {   // 16 independent loads in flight per lane: the copy is latency x bytes-in-flight bound
            int32_t j[8];
            T x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) j[u] = __ldcs(Sj + src + k + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = __ldcs(Sx + src + k + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; u++) __stcs(Cj + dst + k + 32 * u, j[u]);
#pragma unroll
            for (int u = 0; u < 8; u++) __stcs(Cx + dst + k + 32 * u, x[u]);
        }
        for (; k + 96 < n; k += 128) {
            int32_t j0 = __ldcs(Sj + src + k), j1 = __ldcs(Sj + src + k + 32), j2 = __ldcs(Sj + src + k + 64), j3 = __ldcs(Sj + src + k + 96);
            T x0 = __ldcs(Sx + src + k), x1 = __ldcs(Sx + src + k + 32), x2 = __ldcs(Sx + src + k + 64), x3 = __ldcs(Sx + src + k + 96);
            __stcs(Cj + dst + k, j0); __stcs(Cj + dst + k + 32, j1); __stcs(Cj + dst + k + 64, j2); __stcs(Cj + dst + k + 96, j3);
            __stcs(Cx + dst + k, x0); __stcs(Cx + dst + k + 32, x1); __stcs(Cx + dst + k + 64, x2); __stcs(Cx + dst + k + 96, x3);
        }
        for (; k < n; k += 32) {
            Cj[dst + k] = __ldcs(Sj + src + k);
            Cx[dst + k] = __ldcs(Sx + src + k);
        }
    }
}

#include "spgemm_tile.cuh"

// ------------------------------------------------------------------ host orchestration
struct Bins {
    int32_t *rows = nullptr;           // row ids grouped by bin
    unsigned long long count[NBINS];   // rows per bin
    unsigned long long start[NBINS];   // offset of each bin inside rows[]
    unsigned long long gtable_entries = 0;   // global-table entries the rows of the last bin need in total
    unsigned long long split_parts = 0;      // CTAs of the last bin when its rows are split into shared-table sized parts
    BinSpec spec;
};

static GrB_Info make_bins(Bins *bins, size_t entry_bytes, int64_t nrows, const int64_t *cnt, std::string *err) {
    bins->spec = make_bin_spec(entry_bytes);
    unsigned long long *d = dev_alloc_t<unsigned long long>(2 * NBINS + 3);   // counts[NBINS], global-table entries, split parts, cursors[NBINS]
    bins->rows = dev_alloc_t<int32_t>((size_t)(nrows > 0 ? nrows : 1));
    if (!d || !bins->rows) { dev_free(d); dev_free(bins->rows); bins->rows = nullptr; return set_error(err, GrB_OUT_OF_MEMORY, "spgemm bins"); }
    cudaMemsetAsync(d, 0, sizeof(unsigned long long) * (2 * NBINS + 3), g_stream);
    int blocks = (int)std::min<int64_t>((nrows + 255) / 256 + 1, (int64_t)g_num_sms * 8);
    {
        LAUNCH_NOTE("spgemm_bin_count");
        bin_count_kernel<<<blocks, 256, 0, g_stream>>>(bins->spec, nrows, cnt, d);
    }
    unsigned long long hcount[NBINS + 2];
    cudaMemcpyAsync(hcount, d, sizeof(unsigned long long) * (NBINS + 2), cudaMemcpyDeviceToHost, g_stream);
    cudaStreamSynchronize(g_stream);   // the ONE host round trip of the binning: grid sizes of the per-bin launches
    BinCursors cur;
    unsigned long long off = 0;
    for (int b = 0; b < NBINS; b++) {
        bins->count[b] = hcount[b];
        bins->start[b] = off;
        cur.v[b] = off;
        if (b > 0) off += bins->count[b];
    }
    bins->gtable_entries = hcount[NBINS];
    bins->split_parts = hcount[NBINS + 1];
    {
        LAUNCH_NOTE("spgemm_bin_fill");
        bin_cursors_kernel<<<1, 32, 0, g_stream>>>(cur, d + NBINS + 2);   // cursors by value: no host buffer to keep alive, no sync
        bin_fill_kernel<<<blocks, 256, 0, g_stream>>>(bins->spec, nrows, cnt, d + NBINS + 2, bins->rows);
    }
    cudaError_t e = cudaGetLastError();
    dev_free(d);
    CUDA_TRY(err, e);
    return GrB_SUCCESS;
}

struct HashArgs {
    const int64_t *Ap; const int64_t *Ae; const int32_t *Aj; const void *Ax;   // Ae / Be: row-end pointers (csr_row_end)
    const int64_t *Bp; const int64_t *Be; const int32_t *Bj; const void *Bx;
    int64_t *row_nnz;                  // written when non-null (symbolic count, or exact count of a one-pass row)
    const int64_t *Op; int32_t *Oj; void *Ox;   // numeric output: row i's entries go to O*[Op[i] ...]
    const int64_t *cnt;                // per-row bound the bins / table sizes were derived from
    MaskArgs mk;                       // mask row pattern (Mp == nullptr: unmasked)
};

template <typename SR, typename T, bool NUMERIC, bool PACK>
static GrB_Info run_bins(const SR &sr, const Bins &bins, const HashArgs &a, std::string *err) {
    typedef HashTable<SR, T, NUMERIC, PACK> Table;
    const size_t entry = Table::entry_bytes();
    const char *phase = NUMERIC ? "numeric" : "symbolic";
    // The bins are independent (disjoint rows).  Big-table bins run one or two latency-bound CTAs per SM, small-row bins many
    // short ones: on two helper streams they share the SMs instead of queueing behind each other, and the global-table bin
    // (host-synchronising batches, launched last) runs on the library stream beside both.  Per-kernel profiling
    // keeps everything on the library stream.
    cudaStream_t aux0 = nullptr, aux1 = nullptr;
    static cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    bool multi = !profiling() && opt_get_int("spgemm_streams", 0) != 0;
    if (multi) {
        aux0 = aux_stream(0);
        aux1 = aux_stream(1);
        if (!ev_fork) cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
        for (int q = 0; q < 2; q++)
            if (!ev_join[q]) cudaEventCreateWithFlags(&ev_join[q], cudaEventDisableTiming);
        multi = aux0 && aux1 && ev_fork && ev_join[0] && ev_join[1];
        (void)cudaGetLastError();
    }
    if (multi) {
        cudaEventRecord(ev_fork, g_stream);
        cudaStreamWaitEvent(aux0, ev_fork, 0);
        cudaStreamWaitEvent(aux1, ev_fork, 0);
    }
    const int big_from = (int)opt_get_int("spgemm_stream_split_bin", 8);
    auto process = [&](int b, cudaStream_t st) -> GrB_Info {
        const int64_t n = (int64_t)bins.count[b];
        if (n == 0) return GrB_SUCCESS;
        const int32_t *rows = bins.rows + bins.start[b];
        const int cap = bins.spec.cap[b], threads = bins.spec.threads[b];
        if (b <= 2) {
            const int rpb = threads / 32;
            const size_t per = (((size_t)cap * entry + 8) + 15) & ~(size_t)15;
            const size_t smem = per * rpb;
            auto kern = spgemm_warp_kernel<SR, T, NUMERIC, PACK>;
            if (smem > 40 * 1024) CUDA_TRY(err, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_warp" : "spgemm_symbolic_warp");
            kern<<<(unsigned)((n + rpb - 1) / rpb), threads, smem, st>>>(sr, rows, n, cap, bins.spec.tf8[b], a.cnt, a.Ap, a.Ae, a.Aj, (const T *)a.Ax, a.Bp, a.Be, a.Bj, (const T *)a.Bx, a.row_nnz, a.Op, a.Oj, (T *)a.Ox, a.mk);
        } else if (b < NBINS - 1 && NUMERIC && !a.mk.Mp && n < ((int64_t)1 << 31) && opt_get_int("spgemm_rows", 1) != 0) {
            // unmasked numeric rows: persistent CTAs, slot-list drain (spgemm_rows_kernel)
            if constexpr (NUMERIC) {
                const int threads = bins.spec.threads_rows[b];
                const size_t smem = (((size_t)cap * entry + 15) & ~(size_t)15) + block_stage_bytes(threads, sizeof(T));
                auto kern = spgemm_rows_kernel<SR, T, PACK>;
                CUDA_TRY(err, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                int per_sm = 0;
                CUDA_TRY(err, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
                if (per_sm < 1) per_sm = 1;
                const long waves = std::max<long>(1, opt_get_int("spgemm_rows_waves", 4));   // CTAs queued beyond the resident ones even out the tail
                const unsigned grid = (unsigned)std::min<int64_t>(n, (int64_t)g_num_sms * per_sm * waves);
                LAUNCH_NOTE("spgemm_numeric_block");
                kern<<<grid, threads, smem, st>>>(sr, rows, (int)n, cap, bins.spec.tf8[b], a.cnt, a.Ap, a.Ae, a.Aj, (const T *)a.Ax, a.Bp, a.Be, a.Bj, (const T *)a.Bx, a.row_nnz, a.Op, a.Oj, (T *)a.Ox);
            }
        } else if (b < NBINS - 1) {
            const size_t smem = (((size_t)cap * entry + 15) & ~(size_t)15) + block_stage_bytes(threads, sizeof(T));
            auto kern = spgemm_block_kernel<SR, T, NUMERIC, PACK, false>;
            CUDA_TRY(err, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_block" : "spgemm_symbolic_block");
            kern<<<(unsigned)n, threads, smem, st>>>(sr, rows, cap, bins.spec.tf8[b], bins.spec.flags, a.cnt, a.Ap, a.Ae, a.Aj, (const T *)a.Ax, a.Bp, a.Be, a.Bj, (const T *)a.Bx, a.row_nnz, a.Op, a.Oj, (T *)a.Ox, nullptr, nullptr, a.mk, nullptr, 0, 0);
        } else {
            // rows whose bound exceeds the largest shared table: global-memory tables.  The total table size came back with the
            // bin counts, so the usual case is fully asynchronous: sizes -> device scan -> one launch, no host round trip
            // unmasked products with exact-count output (row_nnz): split every such row over several CTAs with shared tables
            const int64_t maxc = bins.spec.split_keys;
            if (!a.mk.Mp && a.row_nnz && bins.split_parts > 0 && bins.split_parts < ((unsigned long long)1 << 31) && n < ((int64_t)1 << 31) &&
                opt_get_int("spgemm_split", 1) != 0) {
                int64_t *poffs = dev_alloc_t<int64_t>((size_t)n + 1);
                if (!poffs) return set_error(err, GrB_OUT_OF_MEMORY, "split row offsets");
                note_launch("split_parts");
                split_parts_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, g_stream>>>(rows, n, a.cnt, maxc, poffs);
                GrB_Info sinfo = exclusive_scan_i64(poffs, n + 1, err);
                if (!sinfo) {
                    const int scap = bins.spec.cap[BIN_LAST_SHARED], sthreads = bins.spec.threads[BIN_LAST_SHARED];
                    const size_t smem = (((size_t)scap * entry + 15) & ~(size_t)15) + block_stage_bytes(sthreads, sizeof(T));
                    auto kern = spgemm_block_kernel<SR, T, NUMERIC, PACK, false, true>;
                    CUDA_TRY(err, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                    LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_split" : "spgemm_symbolic_split");
                    kern<<<(unsigned)bins.split_parts, sthreads, smem, g_stream>>>(sr, rows, scap, bins.spec.tf8[BIN_LAST_SHARED], bins.spec.flags, a.cnt, a.Ap, a.Ae, a.Aj, (const T *)a.Ax, a.Bp, a.Be, a.Bj, (const T *)a.Bx, a.row_nnz, a.Op, a.Oj, (T *)a.Ox, nullptr, nullptr, a.mk, poffs, (int)n, maxc);
                    cudaError_t ge = cudaGetLastError();
                    if (ge != cudaSuccess) sinfo = cuda_fail(err, ge, "split-row spgemm kernel");
                }
                dev_free(poffs);
                return sinfo;
            }
            const int64_t budget_entries = (int64_t)opt_get_int("spgemm_gtable_entries", (long)1 << 30);
            const int64_t tot_all = (int64_t)bins.gtable_entries;
            if (tot_all <= budget_entries) {
                int64_t *doffs = dev_alloc_t<int64_t>((size_t)n + 1);
                const size_t bytes = Table::kPacked ? (size_t)tot_all * 8 : (size_t)tot_all * 4 + (NUMERIC ? (size_t)tot_all * sizeof(T) + 16 : 0);
                unsigned char *gt = (unsigned char *)dev_alloc(bytes);
                GrB_Info ginfo = GrB_SUCCESS;
                if (!doffs || !gt) ginfo = set_error(err, GrB_OUT_OF_MEMORY, "global hash tables (%lld entries)", (long long)tot_all);
                if (!ginfo) {
                    note_launch("gtable_sizes");
                    gtable_sizes_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, g_stream>>>(rows, n, a.cnt, doffs);
                    ginfo = exclusive_scan_i64(doffs, n + 1, err);
                }
                if (!ginfo) {
                    LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_global" : "spgemm_symbolic_global");
                    spgemm_block_kernel<SR, T, NUMERIC, PACK, true><<<(unsigned)n, 1024, block_stage_bytes(1024, sizeof(T)), g_stream>>>(sr, rows, 0, bins.spec.tf8[b], bins.spec.flags, a.cnt, a.Ap, a.Ae, a.Aj, (const T *)a.Ax, a.Bp, a.Be, a.Bj, (const T *)a.Bx, a.row_nnz, a.Op, a.Oj, (T *)a.Ox, gt, doffs, a.mk, nullptr, 0, 0);
                    cudaError_t ge = cudaGetLastError();
                    if (ge != cudaSuccess) ginfo = cuda_fail(err, ge, "global-table spgemm kernel");
                }
                dev_free(doffs); dev_free(gt);   // stream-ordered
                return ginfo;
            }
            // larger than the budget: batches sized on the host
            int64_t *sizes = dev_alloc_t<int64_t>((size_t)n + 1);
            if (!sizes) return set_error(err, GrB_OUT_OF_MEMORY, "global table sizes");
            note_launch("gtable_sizes");
            gtable_sizes_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, g_stream>>>(rows, n, a.cnt, sizes);
            std::vector<int64_t> hs((size_t)n + 1);
            cudaMemcpyAsync(hs.data(), sizes, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, g_stream);
            cudaStreamSynchronize(g_stream);
            dev_free(sizes);
            int64_t i0 = 0;
            GrB_Info info = GrB_SUCCESS;
            while (i0 < n && !info) {
                int64_t i1 = i0, tot = 0;
                while (i1 < n && (i1 == i0 || tot + hs[(size_t)i1] <= budget_entries)) tot += hs[(size_t)i1++];
                std::vector<int64_t> offs((size_t)(i1 - i0) + 1);
                offs[0] = 0;
                for (int64_t q = i0; q < i1; q++) offs[(size_t)(q - i0) + 1] = offs[(size_t)(q - i0)] + hs[(size_t)q];
                int64_t *doffs = dev_alloc_t<int64_t>(offs.size());
                // packed / symbolic: one array of `tot` entries; split numeric: tot keys (4 B) then tot values
                const size_t bytes = Table::kPacked ? (size_t)tot * 8 : (size_t)tot * 4 + (NUMERIC ? (size_t)tot * sizeof(T) + 16 : 0);
                unsigned char *gt = (unsigned char *)dev_alloc(bytes);
                if (!doffs || !gt) info = set_error(err, GrB_OUT_OF_MEMORY, "global hash tables (%lld entries)", (long long)tot);
                if (!info) {
                    cudaMemcpyAsync(doffs, offs.data(), sizeof(int64_t) * offs.size(), cudaMemcpyHostToDevice, g_stream);
                    LAUNCH_NOTE(NUMERIC ? "spgemm_numeric_global" : "spgemm_symbolic_global");
                    spgemm_block_kernel<SR, T, NUMERIC, PACK, true><<<(unsigned)(i1 - i0), 1024, block_stage_bytes(1024, sizeof(T)), g_stream>>>(sr, rows + i0, 0, bins.spec.tf8[b], bins.spec.flags, a.cnt, a.Ap, a.Ae, a.Aj, (const T *)a.Ax, a.Bp, a.Be, a.Bj, (const T *)a.Bx, a.row_nnz, a.Op, a.Oj, (T *)a.Ox, gt, doffs, a.mk, nullptr, 0, 0);
                    cudaStreamSynchronize(g_stream);   // offs is host memory
                }
                dev_free(doffs); dev_free(gt);
                i0 = i1;
            }
            GRB_TRY(info);
        }
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess)
            return set_error(err, GrB_PANIC, "spgemm %s kernel launch failed in bin %d (cap %d, %d threads, %lld rows, entry %zu B): %s",
                             phase, b, cap, threads, (long long)n, entry, cudaGetErrorString(le));
        return GrB_SUCCESS;
    };
    GrB_Info info = GrB_SUCCESS;
    for (int b = NBINS - 2; b >= 1 && !info; b--) info = process(b, multi ? (b >= big_from ? aux0 : aux1) : g_stream);
    if (!info) info = process(NBINS - 1, g_stream);   // global-table rows: library stream (its batches synchronise the host)
    if (multi) {
        cudaEventRecord(ev_join[0], aux0);
        cudaEventRecord(ev_join[1], aux1);
        cudaStreamWaitEvent(g_stream, ev_join[0], 0);
        cudaStreamWaitEvent(g_stream, ev_join[1], 0);
    }
    return info;
}



struct SpgemmPlan {
    CsrArrays *A, *B;
    int64_t m, k, n, annz, bnnz;
    int a_type, b_type;
};

static bool use_packed() { return opt_get_int("spgemm_pack", 1) != 0; }
template <typename T> static size_t numeric_entry_bytes() { return (Packed<T>::value && use_packed()) ? 8 : 4 + sizeof(T); }

// numeric pass writing row i at O*[Op[i]...]; row_nnz (optional) receives exact counts
template <typename T>
static GrB_Info spgemm_numeric_typed(const GrB_Semiring op, const SpgemmPlan &p, const Bins &bins, const int64_t *cnt,
                                     int64_t *row_nnz, const int64_t *Op, int32_t *Oj, void *Ox, MaskArgs mk, std::string *err) {
    GrB_Info info = GrB_SUCCESS;
    const int T_code = type_code_of<T>();
    GRB_DISPATCH_SEMIRING(op->add, op->mul, T, SRT, sr, {
        const void *ax = nullptr, *bx = nullptr;
        void *atmp = nullptr, *btmp = nullptr;
        if (sr.reads_a()) info = cast_view(&ax, &atmp, p.A->val, p.a_type, T_code, p.annz, err);
        if (!info && sr.reads_b()) info = cast_view(&bx, &btmp, p.B->val, p.b_type, T_code, p.bnnz, err);
        if (!info) {
            HashArgs a{p.A->ptr, csr_row_end(*p.A), p.A->idx, ax, p.B->ptr, csr_row_end(*p.B), p.B->idx, bx, row_nnz, Op, Oj, Ox, cnt, mk};
            if (!info) {
                if (Packed<T>::value && use_packed()) info = run_bins<SRT, T, true, true>(sr, bins, a, err);
                else info = run_bins<SRT, T, true, false>(sr, bins, a, err);
            }
        }
        dev_free(atmp);
        dev_free(btmp);
    });
    return info;
}

__global__ void row_end_kernel(int64_t nrows, const int64_t *__restrict__ Sp, int64_t *__restrict__ cnt_to_end) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i <= nrows; i += s) cnt_to_end[i] = Sp[i] + (i < nrows ? cnt_to_end[i] : 0);
}
__global__ void row_len_kernel(int64_t nrows, const int64_t *__restrict__ beg, const int64_t *__restrict__ end, int64_t *__restrict__ len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i <= nrows; i += s) len[i] = i < nrows ? end[i] - beg[i] : 0;
}

template <typename T>
static void launch_compact(int64_t m, const int64_t *Sp, const int64_t *Cp, const int32_t *Sj, const void *Sx, int32_t *Cj, void *Cx) {
    int blocks = (int)std::min<int64_t>((m + 7) / 8, (int64_t)g_num_sms * 32);
    LAUNCH_NOTE("spgemm_compact");
    compact_rows_kernel<T><<<blocks, 256, 0, g_stream>>>(m, Sp, Cp, Sj, (const T *)Sx, Cj, (T *)Cx);
}


// row-end CSR -> compact CSR, once, for consumers that walk ptr[i] .. ptr[i + 1] (matrix_materialize)
GrB_Info csr_compact(GrB_Matrix A) {
    CsrArrays &c = A->csr;
    if (!c.end) return GrB_SUCCESS;
    const int64_t m = A->nrows;
    const size_t es = type_size(A->type);
    std::string *err = &A->err;
    if (!c.canon) {
        c.canon = dev_alloc_t<int64_t>((size_t)m + 1);
        if (!c.canon) return set_error(err, GrB_OUT_OF_MEMORY, "compaction row pointers");
        note_launch("row_len");
        row_len_kernel<<<(unsigned)std::min<int64_t>((m + 256) / 256, (int64_t)g_num_sms * 8), 256, 0, g_stream>>>(m, c.ptr, c.end, c.canon);
        GRB_TRY(exclusive_scan_i64(c.canon, m + 1, err));
    }
    const size_t nv = (size_t)(A->nvals > 0 ? A->nvals : 1);
    int32_t *nj = dev_alloc_t<int32_t>(nv);
    void *nx = dev_alloc(nv * es);
    if (!nj || !nx) {
        dev_free(nj); dev_free(nx);
        return set_error(err, GrB_OUT_OF_MEMORY, "compaction of a %lld-entry product needs %.1f GB more", (long long)A->nvals, (double)nv * (4 + es) / 1e9);
    }
    if (A->nvals > 0) {
        switch (es) {
            case 1: launch_compact<uint8_t>(m, c.ptr, c.canon, c.idx, c.val, nj, nx); break;
            case 2: launch_compact<uint16_t>(m, c.ptr, c.canon, c.idx, c.val, nj, nx); break;
            case 4: launch_compact<uint32_t>(m, c.ptr, c.canon, c.idx, c.val, nj, nx); break;
            default: launch_compact<uint64_t>(m, c.ptr, c.canon, c.idx, c.val, nj, nx); break;
        }
        CUDA_TRY(err, cudaGetLastError());
    }
    dev_free(c.ptr); dev_free(c.end); dev_free(c.idx); dev_free(c.val);
    c.ptr = c.canon; c.canon = nullptr; c.end = nullptr; c.cap = 0;
    c.idx = nj; c.val = nx;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ tiled one-pass (spgemm_tile.cuh): host side
struct TileCfg { bool on = false; int threads = 256, ctas = 2, tcap = 0, scap = 0, tf8 = 12; int64_t R = 0, W = 0; size_t smem = 0; };
struct TileScratch {
    int64_t *small = nullptr, *big = nullptr, *tile_start = nullptr, *Sp = nullptr;
    uint8_t *head = nullptr;
    int *scalars = nullptr;            // [0] n_tiles, [1] ticket
    unsigned long long *status = nullptr;
    void *tmp = nullptr;
    int32_t *Sj = nullptr; void *Sx = nullptr;
    void release() {
        dev_free(small); dev_free(big); dev_free(tile_start); dev_free(Sp); dev_free(head); dev_free(scalars); dev_free(status); dev_free(tmp);
        ws_release(0, Sj); ws_release(1, Sx);
        small = big = tile_start = Sp = nullptr; head = nullptr; scalars = nullptr; status = nullptr; tmp = nullptr; Sj = nullptr; Sx = nullptr;
    }
};

// table / stage sizes from the shared-memory budget of `ctas` CTAs per SM; R = largest row bound hashed by the tile kernel,
// W = flop window of a tile (a tile holds < W + R products, its slices <= tf8/8 (W + R) + 2 TILE_RMAX table entries)
template <typename T> static TileCfg make_tile_cfg() {
    TileCfg c;
    const bool packed = Packed<T>::value && use_packed();
    c.threads = (int)std::min<long>(512, std::max<long>(64, opt_get_int("spgemm_tile_threads", 256))) & ~31;
    c.ctas = (int)std::min<long>(4, std::max<long>(1, opt_get_int("spgemm_tile_ctas", 2)));
    c.scap = (int)std::max<long>(256, opt_get_int("spgemm_tile_scap", 1792)) & ~15;
    c.tf8 = (int)std::max<long>(9, opt_get_int("spgemm_tile_tf8", 16));
    const size_t budget = (size_t)(227 * 1024) / c.ctas - 1024 - 64;   // 1 KB per CTA is reserved by the driver
    const size_t fixed = TileLayout<T>(0, c.scap, packed).total;
    const size_t entry = packed ? 8 : 4 + sizeof(T);
    if (budget < fixed + 4096 * entry) return c;   // not worth it
    c.tcap = (int)((budget - fixed - 16) / entry) & ~63;
    const long tc = opt_get_int("spgemm_tile_tcap", 0);
    if (tc >= 1024 && tc < c.tcap) c.tcap = (int)tc & ~63;
    // the largest row takes at most `spgemm_tile_rfrac8`/8 of the table; W = slot window of a tile, so that a tile's slices
    // (rows starting inside one window) sum to < W + slice(R) <= tcap
    const int64_t rslots = ((int64_t)c.tcap * std::min<long>(7, std::max<long>(1, opt_get_int("spgemm_tile_rfrac8", 5)))) / 8;
    c.R = ((rslots - 34) * 8) / c.tf8;
    const long ropt = opt_get_int("spgemm_tile_r", 0);
    if (ropt > 0 && ropt < c.R) c.R = ropt;
    if (c.R < 1) return c;
    c.W = (int64_t)c.tcap - tile_slice_size(c.R, c.tf8);
    if (c.W < 64) return c;
    c.smem = TileLayout<T>(c.tcap, c.scap, packed).total;
    c.on = true;
    return c;
}

template <typename SR, typename T, bool PACK>
static GrB_Info launch_tile_kernel(const SR &sr, const TileCfg &c, const TileArgs &ta, std::string *err) {
    auto kern = spgemm_tile_kernel<SR, T, PACK>;
    CUDA_TRY(err, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    int per_sm = 0;
    CUDA_TRY(err, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, c.threads, c.smem));
    if (per_sm < 1) return set_error(err, GrB_PANIC, "spgemm tile kernel does not fit an SM (%zu bytes of shared memory)", c.smem);
    per_sm = std::min(per_sm, c.ctas);
    LAUNCH_NOTE("spgemm_numeric_tile");
    kern<<<(unsigned)(g_num_sms * per_sm), c.threads, c.smem, g_stream>>>(sr, ta);
    CUDA_TRY(err, cudaGetLastError());
    return GrB_SUCCESS;
}

template <typename T>
static GrB_Info spgemm_tiled_typed(const GrB_Semiring op, const SpgemmPlan &p, const TileCfg &c, int64_t *flops, int64_t *row_nnz,
                                   uint64_t total_flops, uint64_t big_flops, uint64_t n_big, TileScratch *x, Bins *bins, GrB_Matrix Tm,
                                   int64_t *total_out, std::string *err) {
    const int64_t m = p.m;
    const size_t es = sizeof(T);
    const unsigned blocks = (unsigned)std::min<int64_t>((m + 256) / 256, (int64_t)g_num_sms * 8);
    // ---- 1. bounds of the rows the tile kernel hashes / of the hole rows, tile list in row order
    x->small = dev_alloc_t<int64_t>((size_t)m + 1);
    x->big = dev_alloc_t<int64_t>((size_t)m + 1);
    x->head = dev_alloc_t<uint8_t>((size_t)m + 1);
    x->tile_start = dev_alloc_t<int64_t>((size_t)m + 2);
    x->scalars = dev_alloc_t<int>(4);
    x->status = dev_alloc_t<unsigned long long>((size_t)m + 1);
    if (!x->small || !x->big || !x->head || !x->tile_start || !x->scalars || !x->status) return set_error(err, GrB_OUT_OF_MEMORY, "spgemm tile arrays");
    {
        LAUNCH_NOTE("spgemm_tile_plan");
        tile_small_flops_kernel<<<blocks, 256, 0, g_stream>>>(m, flops, c.R, c.tf8, x->small, x->big);
    }
    GRB_TRY(exclusive_scan_i64(x->small, m + 1, err));   // x->small is now the prefix Fs
    {
        LAUNCH_NOTE("spgemm_tile_plan");
        tile_heads_kernel<<<blocks, 256, 0, g_stream>>>(m, x->small, c.W, x->head);
        size_t tb = 0;
        thrust::counting_iterator<int64_t> it(0);
        CUDA_TRY(err, cub::DeviceSelect::Flagged(nullptr, tb, it, x->head, x->tile_start, x->scalars, (int)m, g_stream));
        x->tmp = dev_alloc(tb);
        if (!x->tmp) return set_error(err, GrB_OUT_OF_MEMORY, "spgemm tile scratch");
        CUDA_TRY(err, cub::DeviceSelect::Flagged(x->tmp, tb, it, x->head, x->tile_start, x->scalars, (int)m, g_stream));
        tile_sentinel_kernel<<<1, 1, 0, g_stream>>>(x->tile_start, x->scalars, m);
        CUDA_TRY(err, cudaMemsetAsync(x->scalars + 1, 0, sizeof(int), g_stream));
        CUDA_TRY(err, cudaMemsetAsync(x->status, 0, sizeof(unsigned long long) * ((size_t)m + 1), g_stream));
    }
    // ---- 2. hole rows: the binned kernels hash them into a staging area addressed by the prefix of their bounds
    int64_t n_listed = 0;
    if (n_big > 0) {
        GRB_TRY(make_bins(bins, numeric_entry_bytes<T>(), m, x->big, err));
        for (int b = 1; b < NBINS; b++) n_listed += (int64_t)bins->count[b];
        x->Sp = dev_alloc_t<int64_t>((size_t)m + 1);
        x->Sj = (int32_t *)ws_acquire(0, (size_t)big_flops * 4);
        x->Sx = ws_acquire(1, (size_t)big_flops * es);
        if (!x->Sp || !x->Sj || !x->Sx) return set_error(err, GrB_OUT_OF_MEMORY, "spgemm hole staging (%llu products)", (unsigned long long)big_flops);
        note_launch("i64_copy");
        i64_copy_kernel<<<blocks, 256, 0, g_stream>>>(x->Sp, x->big, m + 1);
        GRB_TRY(exclusive_scan_i64(x->Sp, m + 1, err));
        GRB_TRY(spgemm_numeric_typed<T>(op, p, *bins, x->big, row_nnz, x->Sp, x->Sj, x->Sx, MaskArgs{nullptr, nullptr, nullptr, 0}, err));
    }
    // ---- 3. result arrays bounded by the flops; the tile kernel writes row pointers, columns and values in place
    const size_t bound = (size_t)(total_flops > 0 ? total_flops : 1);
    Tm->csr.idx = dev_alloc_t<int32_t>(bound);
    Tm->csr.val = dev_alloc(bound * es);
    if (!Tm->csr.idx || !Tm->csr.val) return set_error(err, GrB_OUT_OF_MEMORY, "mxm result bound needs %llu entries (%.1f GB)", (unsigned long long)bound, (double)bound * (4 + es) / 1e9);
    GrB_Info info = GrB_SUCCESS;
    const int T_code = type_code_of<T>();
    GRB_DISPATCH_SEMIRING(op->add, op->mul, T, SRT, sr, {
        const void *ax = nullptr, *bx = nullptr;
        void *atmp = nullptr, *btmp = nullptr;
        if (sr.reads_a()) info = cast_view(&ax, &atmp, p.A->val, p.a_type, T_code, p.annz, err);
        if (!info && sr.reads_b()) info = cast_view(&bx, &btmp, p.B->val, p.b_type, T_code, p.bnnz, err);
        if (!info) {
            TileArgs ta;
            ta.Ap = p.A->ptr; ta.Aj = p.A->idx; ta.Ax = ax;
            ta.Bp = p.B->ptr; ta.Bj = p.B->idx; ta.Bx = bx;
            ta.flops = flops; ta.hole_nnz = row_nnz; ta.tile_start = x->tile_start; ta.n_tiles = x->scalars; ta.ticket = x->scalars + 1;
            ta.status = x->status; ta.Cp = Tm->csr.ptr; ta.Cj = Tm->csr.idx; ta.Cx = Tm->csr.val; ta.nrows = m; ta.R = c.R;
            ta.tcap = c.tcap; ta.scap = c.scap; ta.tf8 = c.tf8; ta.cas_first = opt_get_int("spgemm_cas_first", 1) != 0 ? 1 : 0;
            ta.poll = (int)opt_get_int("spgemm_tile_poll", 0);
            ta.dbg = nullptr;
            unsigned long long *dbg = nullptr;
            if (opt_get_int("spgemm_tile_dbg", 0) != 0) {   // phase counters of the producer / first consumer warp, printed to stderr
                dbg = dev_alloc_t<unsigned long long>(TD_N);
                if (dbg) cudaMemsetAsync(dbg, 0, sizeof(unsigned long long) * TD_N, g_stream);
                ta.dbg = dbg;
            }
            info = launch_tile_kernel<SRT, T, Packed<T>::value>(sr, c, ta, err);
            if (dbg) {
                unsigned long long h[TD_N];
                cudaMemcpyAsync(h, dbg, sizeof h, cudaMemcpyDeviceToHost, g_stream);
                cudaStreamSynchronize(g_stream);
                static const char *nm[TD_N] = {"c_wait_full", "c_insert", "c_scan", "c_lookback", "c_write", "c_bar", "p_wait_tile", "p_tile", "p_wait_empty", "p_load", "p_issue", "chunks", "tiles"};
                fprintf(stderr, "[spgemm_tile] tcap %d scap %d R %lld W %lld threads %d ctas %d smem %zu |", c.tcap, c.scap, (long long)c.R, (long long)c.W, c.threads, c.ctas, c.smem);
                for (int q = 0; q < TD_N; q++) fprintf(stderr, " %s=%.3g", nm[q], (double)h[q]);
                fprintf(stderr, "\n");
                dev_free(dbg);
            }
        }
        dev_free(atmp);
        dev_free(btmp);
    });
    GRB_TRY(info);
    // ---- 4. hole rows into their place
    if (n_listed > 0) {
        LAUNCH_NOTE("spgemm_place_rows");
        const unsigned pb = (unsigned)std::min<int64_t>((n_listed + 7) / 8, (int64_t)g_num_sms * 16);
        place_rows_kernel<T><<<pb, 256, 0, g_stream>>>(bins->rows, n_listed, x->Sp, Tm->csr.ptr, row_nnz, x->Sj, (const T *)x->Sx, Tm->csr.idx, (T *)Tm->csr.val);
        CUDA_TRY(err, cudaGetLastError());
    }
    const int64_t total = read_i64(Tm->csr.ptr + m);
    Tm->nvals = total;
    Tm->jumbled = true;
    *total_out = total;
    // ---- 5. a product that compresses well would keep most of the bound allocated for nothing: move it to exact arrays
    if ((uint64_t)total + total_flops / 8 < total_flops && total_flops > ((uint64_t)1 << 18)) {
        const size_t nv = (size_t)(total > 0 ? total : 1);
        int32_t *nj = dev_alloc_t<int32_t>(nv);
        void *nx = dev_alloc(nv * es);
        if (nj && nx) {
            CUDA_TRY(err, cudaMemcpyAsync(nj, Tm->csr.idx, (size_t)total * 4, cudaMemcpyDeviceToDevice, g_stream));
            CUDA_TRY(err, cudaMemcpyAsync(nx, Tm->csr.val, (size_t)total * es, cudaMemcpyDeviceToDevice, g_stream));
            dev_free(Tm->csr.idx); dev_free(Tm->csr.val);
            Tm->csr.idx = nj; Tm->csr.val = nx;
        } else {
            dev_free(nj); dev_free(nx);   // keep the generous arrays
        }
    }
    return GrB_SUCCESS;
}

GrB_Info spgemm(GrB_Matrix *Tout, const GrB_Semiring op, GrB_Matrix A, bool at, GrB_Matrix B, bool bt, const GrB_Matrix M,
                bool mask_comp, bool mask_struct, std::string *err, bool symbolic_only, uint64_t *flops_out,
                uint64_t *nvals_out) {
    // operands may be row-end CSRs (earlier products): the hash kernels read them as they are; only the tile kernel (bulk
    // copies over the row pointer) and the transposed twins need the compact form
    const bool want_tile = !symbolic_only && opt_get_int("spgemm_tile", 0) != 0;
    GRB_TRY(want_tile ? matrix_materialize(A) : matrix_ensure_ptr(A));
    GRB_TRY(want_tile ? matrix_materialize(B) : matrix_ensure_ptr(B));
    if (at) GRB_TRY(matrix_ensure_twin(A));
    if (bt) GRB_TRY(matrix_ensure_twin(B));
    SpgemmPlan p;
    p.A = at ? &A->twin : &A->csr;
    p.B = bt ? &B->twin : &B->csr;
    p.m = at ? A->ncols : A->nrows;
    p.k = at ? A->nrows : A->ncols;
    const int64_t bk = bt ? B->ncols : B->nrows;
    p.n = bt ? B->nrows : B->ncols;
    p.annz = csr_slots(*p.A, A->nvals); p.bnnz = csr_slots(*p.B, B->nvals);   // value slots (a typecast copies all of them)
    p.a_type = A->type; p.b_type = B->type;
    if (p.k != bk)
        return set_error(err, GrB_DIMENSION_MISMATCH, "mxm: inner dimensions differ (%lld vs %lld)", (long long)p.k, (long long)bk);
    const int D = op ? op->type : TC_INT64;
    const size_t es = type_size(D);

    phase_mark(nullptr);
    int64_t *flops = dev_alloc_t<int64_t>((size_t)p.m + 1), *row_nnz = dev_alloc_t<int64_t>((size_t)p.m + 1);
    int64_t *Sp = nullptr;
    int32_t *Sj = nullptr;
    void *Sx = nullptr;
    TileScratch tsx;   // tiled one-pass (spgemm_tile.cuh)
    int64_t *ccnt = nullptr;   // complemented mask: table-size bounds
    unsigned long long *red = dev_alloc_t<unsigned long long>(4);
    Bins fbins, nbins;
    GrB_Matrix Tm = nullptr;
    GrB_Info info = GrB_SUCCESS;
    if (!flops || !row_nnz || !red) info = set_error(err, GrB_OUT_OF_MEMORY, "spgemm row arrays");
    unsigned long long hred[4] = {0, 0, 0, 0};
    // tiled one-pass (default for unmasked products of >= 4-byte domains): rows with a bound above tile.R are "holes"
    TileCfg tile;
    if (!symbolic_only && es >= 4 && opt_get_int("spgemm_tile", 0) != 0 && !(M && opt_get_int("spgemm_mask", 1) != 0))
        GRB_DISPATCH_TYPE(D, T, if constexpr (sizeof(T) >= 4) tile = make_tile_cfg<T>());
    if (!info) {
        cudaMemsetAsync(red, 0, 32, g_stream);
        cudaMemsetAsync(row_nnz, 0, sizeof(int64_t) * ((size_t)p.m + 1), g_stream);
        cudaMemsetAsync(flops, 0, sizeof(int64_t) * ((size_t)p.m + 1), g_stream);
    }
    if (!info && p.m > 0) {
        int blocks = (int)std::min<int64_t>((p.m + 31) / 32, (int64_t)g_num_sms * 32);
        {
            LAUNCH_NOTE("spgemm_row_flops");
            row_flops_kernel<<<blocks, 256, 0, g_stream>>>(p.m, p.A->ptr, csr_row_end(*p.A), p.A->idx, p.B->ptr, csr_row_end(*p.B), flops);
        }
        {
            LAUNCH_NOTE("reduce_sum_max");
            reduce_sum_max_kernel<<<std::min(blocks, g_num_sms * 4), 256, 0, g_stream>>>(flops, p.m, tile.on ? tile.R : INT64_MAX, red);
        }
        cudaMemcpyAsync(hred, red, 32, cudaMemcpyDeviceToHost, g_stream);
        cudaStreamSynchronize(g_stream);
    }
    const uint64_t total_flops = hred[0];
    if (flops_out) *flops_out = total_flops;
    phase_mark("mxm_row_flops");

    // one pass (no symbolic phase) when the flops-sized staging copy plus the result fit comfortably
    bool onepass = false;
    if (!info && !symbolic_only && total_flops > 0) {
        const char *mode = opt_get("spgemm_mode", "auto");
        static size_t total_b = 0;   // device memory size: asked once (cudaMemGetInfo is a driver round trip)
        if (total_b == 0) {
            size_t free_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
        }
        // staged one-pass: staging + result; tiled: flops-bounded result + staging of the hole rows only
        const double need = tile.on ? ((double)total_flops + (double)hred[2]) * (double)(4 + es) : 2.0 * (double)total_flops * (double)(4 + es);
        const double avail = (double)total_b * 0.9 - (double)GrB_cuda_memory_in_use();
        onepass = !strcmp(mode, "onepass") || (!strcmp(mode, "auto") && need < 0.8 * avail);
        if (!strcmp(mode, "twopass")) onepass = false;
    }

    if (!info) {
        GrB_Info i2 = matrix_new_shell(&Tm, D, p.m, p.n);
        if (i2) info = i2;
    }
    if (!info) {
        Tm->csr.ptr = dev_alloc_t<int64_t>((size_t)p.m + 1);
        if (!Tm->csr.ptr) info = set_error(err, GrB_OUT_OF_MEMORY, "spgemm row pointers");
    }
    int64_t total = 0;
    const unsigned copy_blocks = (unsigned)std::min<int64_t>((p.m + 256) / 256, (int64_t)g_num_sms * 8);

    const bool masked = M && !mask_comp && !symbolic_only && total_flops > 0 && opt_get_int("spgemm_mask", 1) != 0;
    if (!info && masked) {
        // ---- C<M> = A*B, M not complemented: hash tables pre-loaded with the mask rows; output bounded by nnz(M),
        //      so the staging CSR simply has M's row pointers and no symbolic pass is needed
        GrB_Info im = matrix_materialize(M);
        const uint8_t *meff = nullptr;
        void *mtmp = nullptr;
        if (!im && !mask_struct && M->nvals > 0) im = mask_effective_bytes(&meff, &mtmp, nullptr, M->csr.val, M->type, M->nvals, false, err);
        info = im;
        size_t entry = 12;
        GRB_DISPATCH_TYPE(D, T, entry = numeric_entry_bytes<T>());
        int64_t *mcnt = dev_alloc_t<int64_t>((size_t)p.m + 1);
        int32_t *Mj_stage = nullptr;
        void *Mx_stage = nullptr;
        if (!info && !mcnt) info = set_error(err, GrB_OUT_OF_MEMORY, "masked spgemm counts");
        if (!info) {
            BinSpec spec = make_bin_spec(entry);
            note_launch("masked_count");
            masked_count_kernel<<<copy_blocks, 256, 0, g_stream>>>(p.m, flops, M->csr.ptr, spec.maxcount[BIN_LAST_SHARED], mcnt);
            info = make_bins(&fbins, entry, p.m, mcnt, err);
        }
        if (!info) {
            const size_t cap = (size_t)(M->nvals > 0 ? M->nvals : 1);
            Mj_stage = dev_alloc_t<int32_t>(cap);
            Mx_stage = dev_alloc(cap * es);
            if (!Mj_stage || !Mx_stage) info = set_error(err, GrB_OUT_OF_MEMORY, "masked spgemm staging");
        }
        if (!info) {
            GrB_Info i3 = GrB_NOT_IMPLEMENTED;
            MaskArgs mk{M->csr.ptr, M->csr.idx, meff, 0};
            GRB_DISPATCH_TYPE(D, T, i3 = spgemm_numeric_typed<T>(op, p, fbins, mcnt, row_nnz, M->csr.ptr, Mj_stage, Mx_stage, mk, err));
            info = i3;
        }
        if (!info) {
            note_launch("i64_copy");
            i64_copy_kernel<<<copy_blocks, 256, 0, g_stream>>>(Tm->csr.ptr, row_nnz, p.m + 1);
            info = exclusive_scan_i64(Tm->csr.ptr, p.m + 1, err);
            if (!info) total = read_i64(Tm->csr.ptr + p.m);
        }
        if (!info) {
            size_t nv = (size_t)(total > 0 ? total : 1);
            Tm->csr.idx = dev_alloc_t<int32_t>(nv);
            Tm->csr.val = dev_alloc(nv * es);
            Tm->nvals = total;
            Tm->jumbled = true;
            if (!Tm->csr.idx || !Tm->csr.val) info = set_error(err, GrB_OUT_OF_MEMORY, "masked mxm result");
        }
        if (!info && total > 0) {
            switch (es) {
                case 1: launch_compact<uint8_t>(p.m, M->csr.ptr, Tm->csr.ptr, Mj_stage, Mx_stage, Tm->csr.idx, Tm->csr.val); break;
                case 2: launch_compact<uint16_t>(p.m, M->csr.ptr, Tm->csr.ptr, Mj_stage, Mx_stage, Tm->csr.idx, Tm->csr.val); break;
                case 4: launch_compact<uint32_t>(p.m, M->csr.ptr, Tm->csr.ptr, Mj_stage, Mx_stage, Tm->csr.idx, Tm->csr.val); break;
                default: launch_compact<uint64_t>(p.m, M->csr.ptr, Tm->csr.ptr, Mj_stage, Mx_stage, Tm->csr.idx, Tm->csr.val); break;
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) info = cuda_fail(err, e, "masked spgemm compaction");
        }
        dev_free(mtmp); dev_free(mcnt); dev_free(Mj_stage); dev_free(Mx_stage);
    } else if (!info && onepass && tile.on) {
        // ---- tiled one-pass: rows hashed in row order, output written in place (spgemm_tile.cuh)
        GrB_Info i3 = GrB_NOT_IMPLEMENTED;
        GRB_DISPATCH_TYPE(D, T, if constexpr (sizeof(T) >= 4) i3 = spgemm_tiled_typed<T>(op, p, tile, flops, row_nnz, total_flops, hred[2], hred[3], &tsx, &fbins, Tm, &total, err));
        info = i3;
    } else if (!info && onepass) {
        // ---- bins and tables from the flops bound; staging addressed by the flops prefix
        size_t entry = 12;
        GRB_DISPATCH_TYPE(D, T, entry = numeric_entry_bytes<T>());
        phase_mark("mxm_plan");
        // C<!M> = A*B: the mask rows go into the tables as forbidden columns (insert_comp), so the unmasked product is never
        // materialised; tables are sized for products + mask entries, the staging CSR still by the flop bound
        const int64_t *bin_cnt = flops;
        MaskArgs cmk{nullptr, nullptr, nullptr, 0};
        const uint8_t *cmeff = nullptr;
        void *cmtmp = nullptr;
        if (M && mask_comp && opt_get_int("spgemm_mask", 1) != 0) {
            info = matrix_materialize(M);
            if (!info && !mask_struct && M->nvals > 0) info = mask_effective_bytes(&cmeff, &cmtmp, nullptr, M->csr.val, M->type, M->nvals, false, err);
            if (!info) {
                ccnt = dev_alloc_t<int64_t>((size_t)p.m + 1);
                if (!ccnt) info = set_error(err, GrB_OUT_OF_MEMORY, "masked spgemm counts");
            }
            if (!info) {
                note_launch("comp_count");
                comp_count_kernel<<<copy_blocks, 256, 0, g_stream>>>(p.m, flops, M->csr.ptr, ccnt);
                bin_cnt = ccnt;
                cmk = MaskArgs{M->csr.ptr, M->csr.idx, cmeff, 1};
            }
        }
        if (!info) info = make_bins(&fbins, entry, p.m, bin_cnt, err);
        phase_mark("mxm_bins");
        if (!info) {
            Sp = dev_alloc_t<int64_t>((size_t)p.m + 1);
            Sj = (int32_t *)ws_acquire(0, (size_t)total_flops * 4);
            Sx = ws_acquire(1, (size_t)total_flops * es);
            if (!Sp || !Sj || !Sx) info = set_error(err, GrB_OUT_OF_MEMORY, "spgemm staging (%llu products)", (unsigned long long)total_flops);
        }
        if (!info) {
            note_launch("i64_copy");
            i64_copy_kernel<<<copy_blocks, 256, 0, g_stream>>>(Sp, flops, p.m + 1);
            info = exclusive_scan_i64(Sp, p.m + 1, err);
        }
        if (!info) {
            GrB_Info i3 = GrB_NOT_IMPLEMENTED;
            phase_mark("mxm_staging_alloc");
            GRB_DISPATCH_TYPE(D, T, i3 = spgemm_numeric_typed<T>(op, p, fbins, bin_cnt, row_nnz, Sp, Sj, Sx, cmk, err));
            info = i3;
            phase_mark("mxm_numeric_launch");
        }
        if (!info) {
            note_launch("i64_copy");
            i64_copy_kernel<<<copy_blocks, 256, 0, g_stream>>>(Tm->csr.ptr, row_nnz, p.m + 1);
            info = exclusive_scan_i64(Tm->csr.ptr, p.m + 1, err);
            if (!info) total = read_i64(Tm->csr.ptr + p.m);
            phase_mark("mxm_numeric_wait");
        }
        // A product that hardly compresses (nnz(C) within 1/8 of the flop bound: the R-MAT squares) stays where the hash
        // kernels put it: the staging arrays BECOME the result as a row-end CSR (rows in row order, row i at [Sp[i], Sp[i] +
        // count[i])), and nothing is copied.  Later multiplies read that form directly; anything that needs a compact CSR
        // (export, sort, SpMV, element-wise ops) squeezes the gaps out once (matrix_materialize -> csr_compact).  Products that
        // do compress are compacted here, because their staging would pin (flops - nnz) slots of memory for nothing.
        if (!info && total > 0 && opt_get_int("spgemm_row_end", 1) != 0 && (uint64_t)total + total_flops / 8 >= total_flops) {
            note_launch("row_end");
            row_end_kernel<<<copy_blocks, 256, 0, g_stream>>>(p.m, Sp, row_nnz);
            Tm->csr.canon = Tm->csr.ptr;
            Tm->csr.ptr = Sp;
            Tm->csr.end = row_nnz;
            Tm->csr.idx = Sj;
            Tm->csr.val = Sx;
            Tm->csr.cap = (int64_t)total_flops;
            ws_detach(0, Sj); ws_detach(1, Sx);   // the cached scratch blocks now belong to the matrix
            Sp = nullptr; row_nnz = nullptr; Sj = nullptr; Sx = nullptr;
            Tm->nvals = total;
            Tm->jumbled = true;
        } else {
        if (!info) {
            size_t nv = (size_t)(total > 0 ? total : 1);
            Tm->csr.idx = dev_alloc_t<int32_t>(nv);
            Tm->csr.val = dev_alloc(nv * es);
            phase_mark("mxm_result_alloc");
            Tm->nvals = total;
            Tm->jumbled = true;
            if (!Tm->csr.idx || !Tm->csr.val) info = set_error(err, GrB_OUT_OF_MEMORY, "mxm result needs %lld entries (%.1f GB)", (long long)total, (double)total * (4 + es) / 1e9);
        }
        if (!info && total > 0) {
            switch (es) {
                case 1: launch_compact<uint8_t>(p.m, Sp, Tm->csr.ptr, Sj, Sx, Tm->csr.idx, Tm->csr.val); break;
                case 2: launch_compact<uint16_t>(p.m, Sp, Tm->csr.ptr, Sj, Sx, Tm->csr.idx, Tm->csr.val); break;
                case 4: launch_compact<uint32_t>(p.m, Sp, Tm->csr.ptr, Sj, Sx, Tm->csr.idx, Tm->csr.val); break;
                default: launch_compact<uint64_t>(p.m, Sp, Tm->csr.ptr, Sj, Sx, Tm->csr.idx, Tm->csr.val); break;
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) info = cuda_fail(err, e, "spgemm compaction");
        }
        }
        dev_free(cmtmp);
    } else if (!info) {
        // ---- two-pass: symbolic count, exact allocation, numeric
        if (p.m > 0 && total_flops > 0) {
            info = make_bins(&fbins, 4, p.m, flops, err);
            if (!info) {
                HashArgs a{p.A->ptr, csr_row_end(*p.A), p.A->idx, nullptr, p.B->ptr, csr_row_end(*p.B), p.B->idx, nullptr, row_nnz, nullptr, nullptr, nullptr, flops, MaskArgs{nullptr, nullptr, nullptr, 0}};
                SRDyn<int32_t> dummy;
                dummy.a_op = OP_ANY; dummy.m_op = OP_PAIR;
                info = run_bins<SRDyn<int32_t>, int32_t, false, false>(dummy, fbins, a, err);
            }
        }
        if (!info) {
            note_launch("i64_copy");
            i64_copy_kernel<<<copy_blocks, 256, 0, g_stream>>>(Tm->csr.ptr, row_nnz, p.m + 1);
            info = exclusive_scan_i64(Tm->csr.ptr, p.m + 1, err);
            if (!info) total = read_i64(Tm->csr.ptr + p.m);
        }
        if (!info && !symbolic_only) {
            size_t nv = (size_t)(total > 0 ? total : 1);
            Tm->csr.idx = dev_alloc_t<int32_t>(nv);
            Tm->csr.val = dev_alloc(nv * es);
            Tm->nvals = total;
            Tm->jumbled = true;
            if (!Tm->csr.idx || !Tm->csr.val) info = set_error(err, GrB_OUT_OF_MEMORY, "mxm result needs %lld entries (%.1f GB)", (long long)total, (double)total * (4 + es) / 1e9);
            size_t entry = 12;
            GRB_DISPATCH_TYPE(D, T, entry = numeric_entry_bytes<T>());
            if (!info && total > 0) info = make_bins(&nbins, entry, p.m, row_nnz, err);
            if (!info && total > 0) {
                GrB_Info i3 = GrB_NOT_IMPLEMENTED;
                GRB_DISPATCH_TYPE(D, T, i3 = spgemm_numeric_typed<T>(op, p, nbins, row_nnz, nullptr, Tm->csr.ptr, Tm->csr.idx, Tm->csr.val, MaskArgs{nullptr, nullptr, nullptr, 0}, err));
                info = i3;
            }
        }
    }
    if (nvals_out) *nvals_out = (uint64_t)total;
    phase_mark("mxm_compact_launch");
    dev_free(flops); dev_free(row_nnz); dev_free(red); dev_free(fbins.rows); dev_free(nbins.rows);
    dev_free(Sp); ws_release(0, Sj); ws_release(1, Sx);
    tsx.release();
    dev_free(ccnt);
    if (info || symbolic_only) {
        if (Tm) GrB_Matrix_free(&Tm);
        return info;
    }
    *Tout = Tm;
    phase_mark("mxm_free");
    return GrB_SUCCESS;
}
