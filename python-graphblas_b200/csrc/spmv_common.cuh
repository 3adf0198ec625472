// spmv_common.cuh -- pieces shared by the pull SpMV kernels (spmv.cu: merge-path / warp-per-row; spmv_seg.cu: segmented).
#pragma once
#include "grb_ops.cuh"

// ---- shuffle helpers for arbitrary 1..8 byte value types ----
template <typename T> __device__ __forceinline__ T shfl_up_any(T v, int delta) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long b;
        memcpy(&b, &v, 8);
        b = __shfl_up_sync(0xffffffffu, b, delta);
        memcpy(&v, &b, 8);
        return v;
    } else {
        unsigned int b = 0;
        memcpy(&b, &v, sizeof(T));
        b = __shfl_up_sync(0xffffffffu, b, delta);
        memcpy(&v, &b, sizeof(T));
        return v;
    }
}
template <typename T> __device__ __forceinline__ T shfl_down_any(T v, int delta) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long b;
        memcpy(&b, &v, 8);
        b = __shfl_down_sync(0xffffffffu, b, delta);
        memcpy(&v, &b, 8);
        return v;
    } else {
        unsigned int b = 0;
        memcpy(&b, &v, sizeof(T));
        b = __shfl_down_sync(0xffffffffu, b, delta);
        memcpy(&v, &b, sizeof(T));
        return v;
    }
}

template <typename T> __device__ __forceinline__ T shfl_any(T v, int src_lane) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long b;
        memcpy(&b, &v, 8);
        b = __shfl_sync(0xffffffffu, b, src_lane);
        memcpy(&v, &b, 8);
        return v;
    } else {
        unsigned int b = 0;
        memcpy(&b, &v, sizeof(T));
        b = __shfl_sync(0xffffffffu, b, src_lane);
        memcpy(&v, &b, sizeof(T));
        return v;
    }
}

// ------------------------------------------------------------------ write-back fused into the row emission
// w<M, replace> accum= t, applied in registers at the moment a row's reduction is complete.
// `active == 0` means plain T output (no mask, no accumulator).  c_* is the OLD content of the output vector.
template <typename T> struct VecEpi {
    const T *c_vals; const uint8_t *c_present; const uint8_t *mask;
    int active, has_mask, comp, replace, accum;
    // fused exchange: the finished value also goes to position poff + row of npeer peer vectors (NVLink P2P stores)
    int npeer;
    long long poff;
    const T *pscale;
    T *pv[MAX_PEERS];
    uint8_t *pp[MAX_PEERS];
};
// PEER = false: the caller pushes whole runs of finished rows to the peers itself (merge-path kernel: one coalesced run per tile)
template <typename T, bool PEER = true>
__device__ __forceinline__ void epi_write(const VecEpi<T> &e, int64_t row, T t, int tp, T *__restrict__ w_vals,
                                          uint8_t *__restrict__ w_present) {
    if (!e.active) {
        w_vals[row] = tp ? t : T();
        w_present[row] = (uint8_t)tp;
        return;
    }
    const bool m = e.has_mask ? ((e.mask[row] != 0) != (e.comp != 0)) : !e.comp;
    const bool cp = e.c_present ? e.c_present[row] != 0 : false;
    const T c = cp ? e.c_vals[row] : T();
    T z = t;
    bool zp = tp != 0;
    if (e.accum != OP_NONE && cp) {
        z = zp ? binop<T, false>(e.accum, c, z) : c;   // no pow here (api.cu never fuses a pow accumulator): keeps the hot kernels lean
        zp = true;
    }
    if (!m) {
        if (e.replace) zp = false;
        else { z = c; zp = cp; }
    }
    w_vals[row] = zp ? z : T();
    w_present[row] = zp ? 1 : 0;
    if (PEER && e.npeer) {
        const T out = zp ? (e.pscale ? binop<T>(OP_TIMES, z, e.pscale[row]) : z) : T();
#pragma unroll 1
        for (int k = 0; k < e.npeer; k++) {
            e.pv[k][e.poff + row] = out;
            if (e.pp[k]) e.pp[k][e.poff + row] = zp ? 1 : 0;
        }
    }
}


// 128-bit loads straight into registers (no local staging array): WORDS 32-bit words from a 4*WORDS-aligned address
template <int WORDS> __device__ __forceinline__ void load_words(unsigned (&w)[WORDS], const void *src) {
    if constexpr (WORDS == 8) {
        const uint4 v0 = reinterpret_cast<const uint4 *>(src)[0], v1 = reinterpret_cast<const uint4 *>(src)[1];
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    } else if constexpr (WORDS == 4) {
        const uint4 v0 = *reinterpret_cast<const uint4 *>(src);
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
    } else {
        static_assert(WORDS == 2, "unsupported vector width");
        const uint2 v0 = *reinterpret_cast<const uint2 *>(src);
        w[0] = v0.x; w[1] = v0.y;
    }
}
// element j of a packed run of T held in 32-bit words
template <typename T, int WORDS> __device__ __forceinline__ T word_elem(const unsigned (&w)[WORDS], int j) {
    T v;
    if constexpr (sizeof(T) == 8) {
        const unsigned long long u = (unsigned long long)w[2 * j] | ((unsigned long long)w[2 * j + 1] << 32);
        memcpy(&v, &u, 8);
    } else if constexpr (sizeof(T) == 4) {
        const unsigned u = w[j];
        memcpy(&v, &u, 4);
    } else if constexpr (sizeof(T) == 2) {
        const uint16_t u = (uint16_t)(w[j >> 1] >> (16 * (j & 1)));
        memcpy(&v, &u, 2);
    } else {
        const uint8_t u = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
        memcpy(&v, &u, 1);
    }
    return v;
}

template <typename SR, typename T> __device__ __forceinline__ void pv_combine(const SR &sr, T av, int ah, T &bv, int &bh) {
    // (bv, bh) <- (av, ah) (+) (bv, bh): the earlier partial on the left
    if (ah) {
        bv = bh ? sr.add(av, bv) : av;
        bh = 1;
    }
}


// segmented pull SpMV (spmv_seg.cu).  Returns GrB_SUCCESS with *handled = false when the inputs do not fit the kernel
// (unaligned arrays); the caller then falls back to the merge-path kernel.  `epi` may be null (plain T output).
// hot_mode: -1 = follow option spmv_hot ("1" forces the hot-column cache), 0 = plain, 1 = hot-column cache when it is viable;
// *used_hot reports whether the cache variant actually ran.
GrB_Info spmv_seg_run(int type_code, int add_op, int mul_op, CsrArrays &M, int64_t mrows, int64_t ncols, int64_t nnz,
                      const void *avals, const void *x, const uint8_t *xp, void *t_vals, uint8_t *t_present,
                      const void *epi_typed, std::string *err, bool *handled, int hot_mode, bool *used_hot);

// column-banded pull SpMV (spmv_band.cu): x staged one 128 KB column window at a time in shared memory.  *handled = false when the
// format does not apply (values would need a typecast, tiny matrix); plain T output only (no fused write-back).
GrB_Info spmv_band_run(int type_code, int add_op, int mul_op, CsrArrays &M, int64_t mrows, int64_t ncols, int64_t nnz, int val_type,
                       const void *x, const uint8_t *xp, void *t_vals, uint8_t *t_present, std::string *err, bool *handled);
void csr_drop_band(CsrArrays &c);
