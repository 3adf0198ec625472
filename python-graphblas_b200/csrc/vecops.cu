// vecops.cu -- the O(n) vector operations that surround the multiply in BFS / SSSP / PageRank loops, so a
// whole iteration stays on the device:
//   eWiseAdd / eWiseMult (reference core/vector.py:1050-1055, 1142), apply (core/matrix.py:2440-2533),
//   reduce to scalar (core/vector.py:1669-1681), scalar assign under a mask (core/vector.py:2020-2035).
// Each computes T in fresh buffers and goes through the common write-back (epilogue.cu).
#include <vector>

#include "grb_ops.cuh"


static inline int grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    return (int)std::min<int64_t>(b, (int64_t)g_num_sms * 16);
}

// ------------------------------------------------------------------ eWise
template <typename T, bool UNION, bool CMP>
__global__ void ewise_kernel(int64_t n, int op, const T *__restrict__ u, const uint8_t *__restrict__ up,
                             const T *__restrict__ v, const uint8_t *__restrict__ vp, void *__restrict__ tvals,
                             uint8_t *__restrict__ tp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const bool a = up[i] != 0, b = vp[i] != 0;
        bool p;
        T z = T();
        bool zb = false;
        if (a && b) {
            p = true;
            if (CMP) zb = cmpop<T>(op, u[i], v[i]);
            else z = binop<T>(op, u[i], v[i]);
        } else if (UNION && (a || b)) {
            p = true;
            z = a ? u[i] : v[i];
            if (CMP) zb = truthy<T>(z);
        } else {
            p = false;
        }
        if (CMP) ((uint8_t *)tvals)[i] = (p && zb) ? 1 : 0;
        else ((T *)tvals)[i] = p ? z : T();
        tp[i] = p ? 1 : 0;
    }
}

static GrB_Info ewise(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op, const GrB_Vector u,
                      const GrB_Vector v, const GrB_Descriptor desc, bool is_union) {
    CHECK_INIT();
    if (!valid(w)) return GrB_UNINITIALIZED_OBJECT;
    if (!op || !valid(u) || !valid(v)) return set_error(&w->err, GrB_NULL_POINTER, "eWise: null or uninitialised argument");
    if (mask && !valid(mask)) return GrB_UNINITIALIZED_OBJECT;
    if (u->n != w->n || v->n != w->n || (mask && mask->n != w->n))
        return set_error(&w->err, GrB_DIMENSION_MISMATCH, "eWise: sizes differ (%lld, %lld, %lld)", (long long)w->n, (long long)u->n, (long long)v->n);
    const bool cmp = op->ztype != op->type;
    const int64_t n = w->n;
    GRB_TRY(vector_ensure_arrays(u));
    GRB_TRY(vector_ensure_arrays(v));
    const void *uv, *vv;
    void *utmp, *vtmp;
    GRB_TRY(cast_view(&uv, &utmp, u->vals, u->type, op->type, n, &w->err));
    GrB_Info info = cast_view(&vv, &vtmp, v->vals, v->type, op->type, n, &w->err);
    if (info) { dev_free(utmp); return info; }
    size_t nn = (size_t)(n > 0 ? n : 1);
    void *tv = dev_alloc(nn * type_size(op->ztype));
    uint8_t *tp = (uint8_t *)dev_alloc(nn);
    if (!tv || !tp) { dev_free(utmp); dev_free(vtmp); dev_free(tv); dev_free(tp); return set_error(&w->err, GrB_OUT_OF_MEMORY, "eWise result"); }
    if (n > 0) {
        LAUNCH_NOTE(is_union ? "ewise_add" : "ewise_mult");
        GRB_DISPATCH_TYPE(op->type, T, {
            if (is_union) {
                if (cmp) ewise_kernel<T, true, true><<<grid_for(n), 256, 0, g_stream>>>(n, op->opcode, (const T *)uv, u->present, (const T *)vv, v->present, tv, tp);
                else ewise_kernel<T, true, false><<<grid_for(n), 256, 0, g_stream>>>(n, op->opcode, (const T *)uv, u->present, (const T *)vv, v->present, tv, tp);
            } else {
                if (cmp) ewise_kernel<T, false, true><<<grid_for(n), 256, 0, g_stream>>>(n, op->opcode, (const T *)uv, u->present, (const T *)vv, v->present, tv, tp);
                else ewise_kernel<T, false, false><<<grid_for(n), 256, 0, g_stream>>>(n, op->opcode, (const T *)uv, u->present, (const T *)vv, v->present, tv, tp);
            }
        });
    }
    cudaError_t e = cudaGetLastError();
    dev_free(utmp); dev_free(vtmp);
    if (e != cudaSuccess) { dev_free(tv); dev_free(tp); return cuda_fail(&w->err, e, "eWise"); }
    return vector_write_back(w, tv, tp, op->ztype, mask, accum, desc, true);
}

extern "C" GrB_Info GrB_Vector_eWiseAdd_BinaryOp(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                 const GrB_Vector u, const GrB_Vector v, const GrB_Descriptor desc) {
    return ewise(w, mask, accum, op, u, v, desc, true);
}
extern "C" GrB_Info GrB_Vector_eWiseMult_BinaryOp(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                  const GrB_Vector u, const GrB_Vector v, const GrB_Descriptor desc) {
    return ewise(w, mask, accum, op, u, v, desc, false);
}

// ------------------------------------------------------------------ apply
template <typename T> __device__ __forceinline__ T unop(int op, T x) {
    if constexpr (is_gbool<T>::value) {
        switch (op) {
            case UOP_LNOT: case UOP_BNOT: return gbool(!x.v);
            case UOP_ONE: return gbool(true);
            default: return x;   // identity, ainv, minv, abs are the identity on BOOL
        }
    } else if constexpr (std::is_floating_point<T>::value) {
        switch (op) {
            case UOP_AINV: return -x;
            case UOP_MINV: return (T)1 / x;
            case UOP_LNOT: return (T)(x == 0);
            case UOP_ABS: return x < 0 ? -x : x;
            case UOP_ONE: return (T)1;
            case UOP_SQRT: return sqrt(x);     // correctly rounded in both precisions (no fast-math)
            case UOP_EXP: return exp(x);
            case UOP_LOG: return log(x);
            case UOP_EXP2: return exp2(x);
            case UOP_LOG2: return log2(x);
            case UOP_LOG10: return log10(x);
            case UOP_FLOOR: return floor(x);
            case UOP_CEIL: return ceil(x);
            case UOP_ROUND: return rint(x);    // GxB_ROUND: C round-half-even is what numpy / the reference's tests expect of rint
            case UOP_TRUNC: return trunc(x);
            case UOP_SIGNUM: return x != x ? x : (T)((x > 0) - (x < 0));
            default: return x;
        }
    } else {
        typedef typename std::make_unsigned<T>::type U;
        switch (op) {
            case UOP_AINV: return (T)((U)0 - (U)x);
            case UOP_MINV: return binop<T>(OP_DIV, (T)1, x);
            case UOP_LNOT: return (T)(x == 0);
            case UOP_ABS: return (std::is_signed<T>::value && x < 0) ? (T)((U)0 - (U)x) : x;
            case UOP_ONE: return (T)1;
            case UOP_BNOT: return (T)~(U)x;
            default: return x;
        }
    }
}
// mode 0: unary op; 1: binop(scalar, x); 2: binop(x, scalar)
template <typename T>
__global__ void apply_kernel(int64_t n, int mode, int op, T scalar, const T *__restrict__ u, const uint8_t *__restrict__ up,
                             T *__restrict__ tv, uint8_t *__restrict__ tp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const bool p = up ? up[i] != 0 : true;   // up == nullptr: every position holds an entry (matrix value arrays)
        T z = T();
        if (p) z = mode == 0 ? unop<T>(op, u[i]) : mode == 1 ? binop<T>(op, scalar, u[i]) : binop<T>(op, u[i], scalar);
        tv[i] = z;
        if (tp) tp[i] = p ? 1 : 0;
    }
}

static GrB_Info apply_common(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, int mode, int opcode, int optype,
                             const void *scalar_host, int scalar_type, const GrB_Vector u, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(w) || !valid(u)) return GrB_UNINITIALIZED_OBJECT;
    if (mask && !valid(mask)) return GrB_UNINITIALIZED_OBJECT;
    if (u->n != w->n || (mask && mask->n != w->n)) return set_error(&w->err, GrB_DIMENSION_MISMATCH, "apply: sizes differ");
    const int64_t n = w->n;
    GRB_TRY(vector_ensure_arrays(u));
    const void *uv;
    void *utmp;
    GRB_TRY(cast_view(&uv, &utmp, u->vals, u->type, optype, n, &w->err));
    size_t nn = (size_t)(n > 0 ? n : 1);
    void *tv = dev_alloc(nn * type_size(optype));
    uint8_t *tp = (uint8_t *)dev_alloc(nn);
    if (!tv || !tp) { dev_free(utmp); dev_free(tv); dev_free(tp); return set_error(&w->err, GrB_OUT_OF_MEMORY, "apply result"); }
    // scalar -> op type on the host (11x11 small switch via double/int64 is enough for builtin scalars)
    unsigned char sbuf[8] = {0};
    if (mode != 0 && scalar_host) {
        double d = 0; int64_t l = 0; uint64_t ul = 0; bool isf = false, isu = false;
        switch (scalar_type) {
            case TC_BOOL: l = *(const uint8_t *)scalar_host != 0; break;
            case TC_INT8: l = *(const int8_t *)scalar_host; break;
            case TC_INT16: l = *(const int16_t *)scalar_host; break;
            case TC_INT32: l = *(const int32_t *)scalar_host; break;
            case TC_INT64: l = *(const int64_t *)scalar_host; break;
            case TC_UINT8: ul = *(const uint8_t *)scalar_host; isu = true; break;
            case TC_UINT16: ul = *(const uint16_t *)scalar_host; isu = true; break;
            case TC_UINT32: ul = *(const uint32_t *)scalar_host; isu = true; break;
            case TC_UINT64: ul = *(const uint64_t *)scalar_host; isu = true; break;
            case TC_FP32: d = *(const float *)scalar_host; isf = true; break;
            default: d = *(const double *)scalar_host; isf = true; break;
        }
#define PUT(CT) { CT x = isf ? (CT)d : isu ? (CT)ul : (CT)l; memcpy(sbuf, &x, sizeof x); }
        switch (optype) {
            case TC_BOOL: { uint8_t x = isf ? d != 0 : isu ? ul != 0 : l != 0; sbuf[0] = x; } break;
            case TC_INT8: PUT(int8_t) break; case TC_INT16: PUT(int16_t) break; case TC_INT32: PUT(int32_t) break;
            case TC_INT64: PUT(int64_t) break; case TC_UINT8: PUT(uint8_t) break; case TC_UINT16: PUT(uint16_t) break;
            case TC_UINT32: PUT(uint32_t) break; case TC_UINT64: PUT(uint64_t) break; case TC_FP32: PUT(float) break;
            default: PUT(double) break;
        }
#undef PUT
    }
    if (n > 0) {
        LAUNCH_NOTE("apply");
        GRB_DISPATCH_TYPE(optype, T, {
            T s;
            memcpy(&s, sbuf, sizeof(T));
            apply_kernel<T><<<grid_for(n), 256, 0, g_stream>>>(n, mode, opcode, s, (const T *)uv, u->present, (T *)tv, tp);
        });
    }
    cudaError_t e = cudaGetLastError();
    dev_free(utmp);
    if (e != cudaSuccess) { dev_free(tv); dev_free(tp); return cuda_fail(&w->err, e, "apply"); }
    return vector_write_back(w, tv, tp, optype, mask, accum, desc, true);
}

extern "C" GrB_Info GrB_Vector_apply(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_UnaryOp op,
                                     const GrB_Vector u, const GrB_Descriptor desc) {
    if (!op) return GrB_NULL_POINTER;
    return apply_common(w, mask, accum, 0, op->opcode, op->type, nullptr, 0, u, desc);
}
// w<mask> accum= op(scalar, u) (scalar_first != 0) or op(u, scalar)
extern "C" GrB_Info GrB_cuda_Vector_apply_binop(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                const GrB_Vector u, const void *scalar, GrB_Type scalar_type, int scalar_first,
                                                const GrB_Descriptor desc) {
    if (!op || !scalar || !scalar_type) return GrB_NULL_POINTER;
    if (op->ztype != op->type) return GrB_NOT_IMPLEMENTED;
    return apply_common(w, mask, accum, scalar_first ? 1 : 2, op->opcode, op->type, scalar, scalar_type->code, u, desc);
}

// ------------------------------------------------------------------ assign scalar (GrB_ALL)
template <typename T> __global__ void fill_full_kernel(int64_t n, T s, T *tv, uint8_t *tp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) { tv[i] = s; tp[i] = 1; }
}

extern "C" GrB_Info GrB_cuda_Vector_assign_scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const void *val,
                                                  GrB_Type val_type, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(w)) return GrB_UNINITIALIZED_OBJECT;
    if (!val || !val_type) return GrB_NULL_POINTER;
    if (mask && (!valid(mask) || mask->n != w->n)) return set_error(&w->err, GrB_DIMENSION_MISMATCH, "assign: mask size differs");
    const int64_t n = w->n;
    size_t nn = (size_t)(n > 0 ? n : 1);
    const int t = val_type->code;
    void *tv = dev_alloc(nn * type_size(t));
    uint8_t *tp = (uint8_t *)dev_alloc(nn);
    if (!tv || !tp) { dev_free(tv); dev_free(tp); return set_error(&w->err, GrB_OUT_OF_MEMORY, "assign"); }
    if (n > 0) {
        LAUNCH_NOTE("fill_full");
        GRB_DISPATCH_TYPE(t, T, {
            T s;
            memcpy(&s, val, sizeof(T));
            fill_full_kernel<T><<<grid_for(n), 256, 0, g_stream>>>(n, s, (T *)tv, tp);
        });
    }
    CUDA_TRY(&w->err, cudaGetLastError());
    return vector_write_back(w, tv, tp, t, mask, accum, desc, true);
}

// ------------------------------------------------------------------ reduce to scalar (deterministic two-pass)
template <typename T>
__global__ void reduce_partial_kernel(int64_t n, int op, const T *__restrict__ u, const uint8_t *__restrict__ up,
                                      T *__restrict__ partial, unsigned long long *__restrict__ pcount) {
    __shared__ T s_val[32];
    __shared__ unsigned long long s_cnt[32];
    T acc = monoid_identity<T>(op);
    unsigned long long cnt = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        if (!up || up[i]) { acc = binop<T>(op, acc, u[i]); cnt++; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        T other;
        if constexpr (sizeof(T) == 8) {
            unsigned long long b; memcpy(&b, &acc, 8); b = __shfl_down_sync(0xffffffffu, b, o); memcpy(&other, &b, 8);
        } else {
            unsigned int b = 0; memcpy(&b, &acc, sizeof(T)); b = __shfl_down_sync(0xffffffffu, b, o); memcpy(&other, &b, sizeof(T));
        }
        // combine only sides that hold values: any(x, y) returns y, so an empty partner must not overwrite a real value
        const unsigned long long ocnt = __shfl_down_sync(0xffffffffu, cnt, o);
        if (ocnt) acc = cnt ? binop<T>(op, acc, other) : other;
        cnt += ocnt;
    }
    if (lane == 0) { s_val[warp] = acc; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        T a = s_val[0];
        unsigned long long c = s_cnt[0];
        for (int q = 1; q < (int)(blockDim.x >> 5); q++) {
            if (s_cnt[q]) a = c ? binop<T>(op, a, s_val[q]) : s_val[q];
            c += s_cnt[q];
        }
        partial[blockIdx.x] = a;
        pcount[blockIdx.x] = c;
    }
}

// a scalar of any builtin type in its widest faithful representation
struct WideScalar { double d; int64_t l; uint64_t ul; bool isf, isu; };
template <typename T> static WideScalar fold_partials(const unsigned char *hp, int blocks, int opcode, const unsigned long long *counts = nullptr) {
    T acc = monoid_identity<T>(opcode);
    bool have = false;
    for (int b = 0; b < blocks; b++) {
        if (counts && counts[b] == 0) continue;   // a block that saw no entry holds the identity, which is not neutral for ANY
        T x;
        memcpy(&x, hp + (size_t)b * sizeof(T), sizeof(T));
        acc = have ? binop<T>(opcode, acc, x) : x;
        have = true;
    }
    WideScalar w;
    if constexpr (is_gbool<T>::value) { w.d = acc.v; w.l = acc.v; w.ul = acc.v; w.isf = false; w.isu = true; }
    else { w.d = (double)acc; w.l = (int64_t)acc; w.ul = (uint64_t)acc; w.isf = std::is_floating_point<T>::value; w.isu = std::is_unsigned<T>::value; }
    return w;
}
template <typename CT> static void put_scalar(void *val, int accum, const WideScalar &w) {
    CT x = w.isf ? (CT)w.d : w.isu ? (CT)w.ul : (CT)w.l;
    if (accum != OP_NONE) {
        CT old;
        memcpy(&old, val, sizeof old);
        x = binop<CT>(accum, old, x);
    }
    memcpy(val, &x, sizeof x);
}
static void store_scalar(void *val, int val_type, int accum, const WideScalar &w) {
    switch (val_type) {
        case TC_BOOL: {
            gbool x(w.isf ? w.d != 0 : w.isu ? w.ul != 0 : w.l != 0);
            if (accum != OP_NONE) { gbool o; o.v = *(uint8_t *)val; x = binop<gbool>(accum, o, x); }
            *(uint8_t *)val = x.v;
        } break;
        case TC_INT8: put_scalar<int8_t>(val, accum, w); break;
        case TC_INT16: put_scalar<int16_t>(val, accum, w); break;
        case TC_INT32: put_scalar<int32_t>(val, accum, w); break;
        case TC_INT64: put_scalar<int64_t>(val, accum, w); break;
        case TC_UINT8: put_scalar<uint8_t>(val, accum, w); break;
        case TC_UINT16: put_scalar<uint16_t>(val, accum, w); break;
        case TC_UINT32: put_scalar<uint32_t>(val, accum, w); break;
        case TC_UINT64: put_scalar<uint64_t>(val, accum, w); break;
        case TC_FP32: put_scalar<float>(val, accum, w); break;
        default: put_scalar<double>(val, accum, w); break;
    }
}

extern "C" GrB_Info GrB_cuda_Vector_reduce(void *val, GrB_Type val_type, const GrB_BinaryOp accum, const GrB_Monoid op,
                                           const GrB_Vector u, GrB_Index *nvals_out) {
    CHECK_INIT();
    if (!val || !val_type || !op) return GrB_NULL_POINTER;
    if (!valid(u)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(vector_ensure_arrays(u));
    const int64_t n = u->n;
    const int mt = op->type;
    const void *uv;
    void *utmp;
    GRB_TRY(cast_view(&uv, &utmp, u->vals, u->type, mt, n, &u->err));
    const int blocks = std::max(1, std::min(grid_for(n), g_num_sms * 4));
    void *partial = dev_alloc((size_t)blocks * 8);
    unsigned long long *pcount = dev_alloc_t<unsigned long long>((size_t)blocks);
    if (!partial || !pcount) { dev_free(utmp); dev_free(partial); dev_free(pcount); return set_error(&u->err, GrB_OUT_OF_MEMORY, "reduce"); }
    {
        LAUNCH_NOTE("reduce_partial");
        GRB_DISPATCH_TYPE(mt, T, (reduce_partial_kernel<T><<<blocks, 256, 0, g_stream>>>(n, op->opcode, (const T *)uv, u->present, (T *)partial, pcount)));
    }
    std::vector<unsigned char> hp((size_t)blocks * 8);
    std::vector<unsigned long long> hc((size_t)blocks);
    cudaMemcpyAsync(hp.data(), partial, (size_t)blocks * type_size(mt), cudaMemcpyDeviceToHost, g_stream);
    cudaMemcpyAsync(hc.data(), pcount, (size_t)blocks * 8, cudaMemcpyDeviceToHost, g_stream);
    cudaError_t e = cudaStreamSynchronize(g_stream);
    dev_free(utmp); dev_free(partial); dev_free(pcount);
    CUDA_TRY(&u->err, e);
    unsigned long long total = 0;
    for (auto c : hc) total += c;
    if (nvals_out) *nvals_out = total;
    u->nvals = (int64_t)total;
    // fold the per-block partials in block order on the host, then accum + cast into *val
    WideScalar ws;
    GRB_DISPATCH_TYPE(mt, T, ws = fold_partials<T>(hp.data(), blocks, op->opcode, hc.data()));
    store_scalar(val, val_type->code, accum ? accum->opcode : OP_NONE, ws);
    return GrB_SUCCESS;
}

// ================================================================== matrix element-wise operations
// GrB_transpose (reference core/base.py:401-411), GrB_Matrix_apply incl. bind-1st / bind-2nd (core/matrix.py:2440-2533),
// GrB_Matrix_eWiseAdd / eWiseMult_BinaryOp (core/matrix.py:1972-2165), reduce to scalar (core/matrix.py:2703-2735) -- the
// last two are what the reference's own parity predicate Matrix.isequal is made of (core/matrix.py:408-415).  Each builds T
// in fresh CSR arrays and goes through the common matrix write-back (mask / accum / replace, epilogue.cu).

static GrB_Info scalar_to_type(unsigned char *sbuf, const void *scalar_host, int scalar_type, int optype) {
    double d = 0; int64_t l = 0; uint64_t ul = 0; bool isf = false, isu = false;
    switch (scalar_type) {
        case TC_BOOL: l = *(const uint8_t *)scalar_host != 0; break;
        case TC_INT8: l = *(const int8_t *)scalar_host; break;
        case TC_INT16: l = *(const int16_t *)scalar_host; break;
        case TC_INT32: l = *(const int32_t *)scalar_host; break;
        case TC_INT64: l = *(const int64_t *)scalar_host; break;
        case TC_UINT8: ul = *(const uint8_t *)scalar_host; isu = true; break;
        case TC_UINT16: ul = *(const uint16_t *)scalar_host; isu = true; break;
        case TC_UINT32: ul = *(const uint32_t *)scalar_host; isu = true; break;
        case TC_UINT64: ul = *(const uint64_t *)scalar_host; isu = true; break;
        case TC_FP32: d = *(const float *)scalar_host; isf = true; break;
        default: d = *(const double *)scalar_host; isf = true; break;
    }
#define PUT(CT) { CT x = isf ? (CT)d : isu ? (CT)ul : (CT)l; memcpy(sbuf, &x, sizeof x); }
    switch (optype) {
        case TC_BOOL: { uint8_t x = isf ? d != 0 : isu ? ul != 0 : l != 0; sbuf[0] = x; } break;
        case TC_INT8: PUT(int8_t) break; case TC_INT16: PUT(int16_t) break; case TC_INT32: PUT(int32_t) break;
        case TC_INT64: PUT(int64_t) break; case TC_UINT8: PUT(uint8_t) break; case TC_UINT16: PUT(uint16_t) break;
        case TC_UINT32: PUT(uint32_t) break; case TC_UINT64: PUT(uint64_t) break; case TC_FP32: PUT(float) break;
        default: PUT(double) break;
    }
#undef PUT
    return GrB_SUCCESS;
}

// the CSR a matrix operand presents: its own arrays, or the cached transpose twin under GrB_DESC_T0 / T1
struct OperandCsr { const CsrArrays *c; int64_t nrows, ncols; };
static GrB_Info operand_csr(OperandCsr *o, GrB_Matrix A, bool transposed, bool need_sorted) {
    GRB_TRY(matrix_materialize(A));
    if (need_sorted) GRB_TRY(matrix_ensure_sorted(A));
    if (transposed) {
        GRB_TRY(matrix_ensure_twin(A));   // built by a stable sort: rows of the twin are sorted
        o->c = &A->twin; o->nrows = A->ncols; o->ncols = A->nrows;
    } else {
        o->c = &A->csr; o->nrows = A->nrows; o->ncols = A->ncols;
    }
    return GrB_SUCCESS;
}

// T with the pattern of `src` (row pointers and column indices copied) and room for values of `type`
static GrB_Info matrix_with_pattern(GrB_Matrix *Tout, int type, const OperandCsr &src, int64_t nvals, bool jumbled, std::string *err) {
    GrB_Matrix T = nullptr;
    GRB_TRY(matrix_new_shell(&T, type, src.nrows, src.ncols));
    GrB_Info info = matrix_alloc_csr(T, nvals);
    if (!info) {
        cudaError_t e = cudaMemcpyAsync(T->csr.ptr, src.c->ptr, sizeof(int64_t) * ((size_t)src.nrows + 1), cudaMemcpyDeviceToDevice, g_stream);
        if (e == cudaSuccess && nvals > 0) e = cudaMemcpyAsync(T->csr.idx, src.c->idx, sizeof(int32_t) * (size_t)nvals, cudaMemcpyDeviceToDevice, g_stream);
        if (e != cudaSuccess) info = cuda_fail(err, e, "copy of the CSR pattern");
    }
    if (info) { GrB_Matrix_free(&T); return info; }
    T->nvals = nvals;
    T->jumbled = jumbled;
    *Tout = T;
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_transpose(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_Matrix A,
                                  const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(C)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "GrB_transpose: output matrix is not initialised");
    if (!valid(A)) return set_error(&C->err, GrB_NULL_POINTER, "GrB_transpose: null or uninitialised argument");
    if (Mask && !valid(Mask)) return set_error(&C->err, GrB_UNINITIALIZED_OBJECT, "GrB_transpose: bad mask");
    const bool t0 = desc && desc->t0;   // transpose of the transposed input = the input itself
    OperandCsr src;
    GRB_TRY(operand_csr(&src, A, !t0, false));
    if (C->nrows != src.nrows || C->ncols != src.ncols || (Mask && (Mask->nrows != C->nrows || Mask->ncols != C->ncols)))
        return set_error(&C->err, GrB_DIMENSION_MISMATCH, "GrB_transpose: C(%lldx%lld) = A(%lldx%lld)'", (long long)C->nrows,
                         (long long)C->ncols, (long long)src.ncols, (long long)src.nrows);
    GrB_Matrix T = nullptr;
    GRB_TRY(matrix_with_pattern(&T, A->type, src, A->nvals, t0 ? A->jumbled : false, &C->err));
    if (A->nvals > 0)
        CUDA_TRY(&C->err, cudaMemcpyAsync(T->csr.val, src.c->val, type_size(A->type) * (size_t)A->nvals, cudaMemcpyDeviceToDevice, g_stream));
    GrB_Info info = matrix_write_back(C, T, Mask, accum, desc);
    GrB_Matrix_free(&T);
    return info;
}

static GrB_Info matrix_apply_common(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, int mode, int opcode, int optype,
                                    const void *scalar_host, int scalar_type, GrB_Matrix A, const GrB_Descriptor desc) {
    // the matrix is the second input of a bind-1st apply (GrB_INP1 transposes it), the first input otherwise
    const bool transposed = desc && (mode == 1 ? desc->t1 : desc->t0);
    CHECK_INIT();
    if (!valid(C)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "apply: output matrix is not initialised");
    if (!valid(A)) return set_error(&C->err, GrB_NULL_POINTER, "apply: null or uninitialised argument");
    if (Mask && !valid(Mask)) return set_error(&C->err, GrB_UNINITIALIZED_OBJECT, "apply: bad mask");
    OperandCsr src;
    GRB_TRY(operand_csr(&src, A, transposed, false));
    if (C->nrows != src.nrows || C->ncols != src.ncols || (Mask && (Mask->nrows != C->nrows || Mask->ncols != C->ncols)))
        return set_error(&C->err, GrB_DIMENSION_MISMATCH, "apply: C(%lldx%lld) vs A(%lldx%lld)", (long long)C->nrows, (long long)C->ncols,
                         (long long)src.nrows, (long long)src.ncols);
    const int64_t nv = A->nvals;
    GrB_Matrix T = nullptr;
    GRB_TRY(matrix_with_pattern(&T, optype, src, nv, transposed ? false : A->jumbled, &C->err));
    const void *av = nullptr;
    void *atmp = nullptr;
    GrB_Info info = nv > 0 ? cast_view(&av, &atmp, src.c->val, A->type, optype, nv, &C->err) : GrB_SUCCESS;
    if (!info && nv > 0) {
        unsigned char sbuf[8] = {0};
        if (mode != 0 && scalar_host) scalar_to_type(sbuf, scalar_host, scalar_type, optype);
        LAUNCH_NOTE("matrix_apply");
        GRB_DISPATCH_TYPE(optype, T_, {
            T_ sc;
            memcpy(&sc, sbuf, sizeof(T_));
            apply_kernel<T_><<<grid_for(nv), 256, 0, g_stream>>>(nv, mode, opcode, sc, (const T_ *)av, nullptr, (T_ *)T->csr.val, nullptr);
        });
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(&C->err, e, "matrix apply");
    }
    dev_free(atmp);
    if (!info) info = matrix_write_back(C, T, Mask, accum, desc);
    GrB_Matrix_free(&T);
    return info;
}
extern "C" GrB_Info GrB_Matrix_apply(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_UnaryOp op,
                                     const GrB_Matrix A, const GrB_Descriptor desc) {
    if (!op) return GrB_NULL_POINTER;
    return matrix_apply_common(C, Mask, accum, 0, op->opcode, op->type, nullptr, 0, A, desc);
}
// C<Mask> accum= op(scalar, A) (scalar_first != 0) or op(A, scalar)
extern "C" GrB_Info GrB_cuda_Matrix_apply_binop(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                const GrB_Matrix A, const void *scalar, GrB_Type scalar_type, int scalar_first,
                                                const GrB_Descriptor desc) {
    if (!op || !scalar || !scalar_type) return GrB_NULL_POINTER;
    if (op->ztype != op->type) return GrB_NOT_IMPLEMENTED;
    return matrix_apply_common(C, Mask, accum, scalar_first ? 1 : 2, op->opcode, op->type, scalar, scalar_type->code, A, desc);
}

// ---- eWise on sorted CSR rows: a warp per row would idle on average-degree-16 rows, so a thread merges one row with two
// pointers (count pass, scan, fill pass); both passes read the same data, the second from L2
template <typename T, bool UNION, bool CMP, bool FILL>
__global__ void mat_ewise_kernel(int64_t nrows, int op, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj,
                                 const T *__restrict__ Ax, const int64_t *__restrict__ Bp, const int32_t *__restrict__ Bj,
                                 const T *__restrict__ Bx, int64_t *__restrict__ Tp, int32_t *__restrict__ Tj, void *__restrict__ Tx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += stride) {
        int64_t a = Ap[i], b = Bp[i];
        const int64_t ae = Ap[i + 1], be = Bp[i + 1];
        int64_t out = FILL ? Tp[i] : 0;
        while (a < ae || b < be) {
            const int32_t ja = a < ae ? Aj[a] : INT32_MAX, jb = b < be ? Bj[b] : INT32_MAX;
            if (ja == jb) {
                if (FILL) {
                    Tj[out] = ja;
                    if (CMP) ((uint8_t *)Tx)[out] = cmpop<T>(op, Ax[a], Bx[b]) ? 1 : 0;
                    else ((T *)Tx)[out] = binop<T>(op, Ax[a], Bx[b]);
                }
                out++; a++; b++;
            } else if (ja < jb) {
                if (UNION) {
                    if (FILL) {
                        Tj[out] = ja;
                        if (CMP) ((uint8_t *)Tx)[out] = truthy<T>(Ax[a]) ? 1 : 0;
                        else ((T *)Tx)[out] = Ax[a];
                    }
                    out++;
                }
                a++;
            } else {
                if (UNION) {
                    if (FILL) {
                        Tj[out] = jb;
                        if (CMP) ((uint8_t *)Tx)[out] = truthy<T>(Bx[b]) ? 1 : 0;
                        else ((T *)Tx)[out] = Bx[b];
                    }
                    out++;
                }
                b++;
            }
        }
        if (!FILL) Tp[i] = out;
    }
}

static GrB_Info matrix_ewise(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op, GrB_Matrix A,
                             GrB_Matrix B, const GrB_Descriptor desc, bool is_union) {
    CHECK_INIT();
    if (!valid(C)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "eWise: output matrix is not initialised");
    if (!op || !valid(A) || !valid(B)) return set_error(&C->err, GrB_NULL_POINTER, "eWise: null or uninitialised argument");
    if (Mask && !valid(Mask)) return set_error(&C->err, GrB_UNINITIALIZED_OBJECT, "eWise: bad mask");
    OperandCsr a, b;
    GRB_TRY(operand_csr(&a, A, desc && desc->t0, true));
    GRB_TRY(operand_csr(&b, B, desc && desc->t1, true));
    if (a.nrows != b.nrows || a.ncols != b.ncols || C->nrows != a.nrows || C->ncols != a.ncols ||
        (Mask && (Mask->nrows != C->nrows || Mask->ncols != C->ncols)))
        return set_error(&C->err, GrB_DIMENSION_MISMATCH, "eWise: C(%lldx%lld), A(%lldx%lld), B(%lldx%lld)", (long long)C->nrows,
                         (long long)C->ncols, (long long)a.nrows, (long long)a.ncols, (long long)b.nrows, (long long)b.ncols);
    const bool cmp = op->ztype != op->type;
    const int64_t m = a.nrows;
    const void *av = nullptr, *bv = nullptr;
    void *atmp = nullptr, *btmp = nullptr;
    GrB_Info info = A->nvals > 0 ? cast_view(&av, &atmp, a.c->val, A->type, op->type, A->nvals, &C->err) : GrB_SUCCESS;
    if (!info && B->nvals > 0) info = cast_view(&bv, &btmp, b.c->val, B->type, op->type, B->nvals, &C->err);
    GrB_Matrix T = nullptr;
    if (!info) info = matrix_new_shell(&T, op->ztype, a.nrows, a.ncols);
    if (!info) {
        T->csr.ptr = dev_alloc_t<int64_t>((size_t)m + 1);
        if (!T->csr.ptr) info = set_error(&C->err, GrB_OUT_OF_MEMORY, "eWise row pointers");
    }
    int64_t total = 0;
    const int blocks = grid_for(m > 0 ? m : 1);
#define MAT_EWISE(FILLV)                                                                                                             \
    GRB_DISPATCH_TYPE(op->type, T_, {                                                                                              \
        if (is_union) {                                                                                                            \
            if (cmp) mat_ewise_kernel<T_, true, true, FILLV><<<blocks, 256, 0, g_stream>>>(m, op->opcode, a.c->ptr, a.c->idx, (const T_ *)av, b.c->ptr, b.c->idx, (const T_ *)bv, T->csr.ptr, T->csr.idx, T->csr.val);   \
            else mat_ewise_kernel<T_, true, false, FILLV><<<blocks, 256, 0, g_stream>>>(m, op->opcode, a.c->ptr, a.c->idx, (const T_ *)av, b.c->ptr, b.c->idx, (const T_ *)bv, T->csr.ptr, T->csr.idx, T->csr.val);       \
        } else {                                                                                                                   \
            if (cmp) mat_ewise_kernel<T_, false, true, FILLV><<<blocks, 256, 0, g_stream>>>(m, op->opcode, a.c->ptr, a.c->idx, (const T_ *)av, b.c->ptr, b.c->idx, (const T_ *)bv, T->csr.ptr, T->csr.idx, T->csr.val);  \
            else mat_ewise_kernel<T_, false, false, FILLV><<<blocks, 256, 0, g_stream>>>(m, op->opcode, a.c->ptr, a.c->idx, (const T_ *)av, b.c->ptr, b.c->idx, (const T_ *)bv, T->csr.ptr, T->csr.idx, T->csr.val);     \
        }                                                                                                                          \
    })
    if (!info) {
        cudaMemsetAsync(T->csr.ptr, 0, sizeof(int64_t) * ((size_t)m + 1), g_stream);
        if (m > 0) {
            LAUNCH_NOTE("matrix_ewise_count");
            MAT_EWISE(false);
        }
        info = exclusive_scan_i64(T->csr.ptr, m + 1, &C->err);
        if (!info) total = read_i64(T->csr.ptr + m);
    }
    if (!info) {
        const size_t nv = (size_t)(total > 0 ? total : 1);
        T->csr.idx = dev_alloc_t<int32_t>(nv);
        T->csr.val = dev_alloc(nv * type_size(op->ztype));
        T->nvals = total;
        T->jumbled = false;
        if (!T->csr.idx || !T->csr.val) info = set_error(&C->err, GrB_OUT_OF_MEMORY, "eWise result (%lld entries)", (long long)total);
    }
    if (!info && total > 0) {
        LAUNCH_NOTE("matrix_ewise_fill");
        MAT_EWISE(true);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(&C->err, e, "matrix eWise");
    }
#undef MAT_EWISE
    dev_free(atmp); dev_free(btmp);
    if (!info) info = matrix_write_back(C, T, Mask, accum, desc);
    if (T) GrB_Matrix_free(&T);
    return info;
}
extern "C" GrB_Info GrB_Matrix_eWiseAdd_BinaryOp(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                 const GrB_Matrix A, const GrB_Matrix B, const GrB_Descriptor desc) {
    return matrix_ewise(C, Mask, accum, op, A, B, desc, true);
}
extern "C" GrB_Info GrB_Matrix_eWiseMult_BinaryOp(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                  const GrB_Matrix A, const GrB_Matrix B, const GrB_Descriptor desc) {
    return matrix_ewise(C, Mask, accum, op, A, B, desc, false);
}

// all entries of A folded with a monoid (deterministic two-pass, as for vectors); *nvals_out = nvals(A)
extern "C" GrB_Info GrB_cuda_Matrix_reduce(void *val, GrB_Type val_type, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Matrix A,
                                           GrB_Index *nvals_out) {
    CHECK_INIT();
    if (!val || !val_type || !op) return GrB_NULL_POINTER;
    if (!valid(A)) return GrB_UNINITIALIZED_OBJECT;
    GRB_TRY(matrix_materialize(A));
    const int64_t n = A->nvals;
    if (nvals_out) *nvals_out = (GrB_Index)n;
    const int mt = op->type;
    WideScalar ws;
    if (n == 0) {
        GRB_DISPATCH_TYPE(mt, T, ws = fold_partials<T>(nullptr, 0, op->opcode));
        store_scalar(val, val_type->code, accum ? accum->opcode : OP_NONE, ws);
        return GrB_SUCCESS;
    }
    const void *av;
    void *atmp;
    GRB_TRY(cast_view(&av, &atmp, A->csr.val, A->type, mt, n, &A->err));
    const int blocks = std::max(1, std::min(grid_for(n), g_num_sms * 4));
    void *partial = dev_alloc((size_t)blocks * 8);
    unsigned long long *pcount = dev_alloc_t<unsigned long long>((size_t)blocks);
    if (!partial || !pcount) { dev_free(atmp); dev_free(partial); dev_free(pcount); return set_error(&A->err, GrB_OUT_OF_MEMORY, "reduce"); }
    {
        LAUNCH_NOTE("matrix_reduce_partial");
        GRB_DISPATCH_TYPE(mt, T, (reduce_partial_kernel<T><<<blocks, 256, 0, g_stream>>>(n, op->opcode, (const T *)av, nullptr, (T *)partial, pcount)));
    }
    std::vector<unsigned char> hp((size_t)blocks * 8);
    std::vector<unsigned long long> hc((size_t)blocks);
    cudaMemcpyAsync(hp.data(), partial, (size_t)blocks * type_size(mt), cudaMemcpyDeviceToHost, g_stream);
    cudaMemcpyAsync(hc.data(), pcount, (size_t)blocks * 8, cudaMemcpyDeviceToHost, g_stream);
    cudaError_t e = cudaStreamSynchronize(g_stream);
    dev_free(atmp); dev_free(partial); dev_free(pcount);
    CUDA_TRY(&A->err, e);
    GRB_DISPATCH_TYPE(mt, T, ws = fold_partials<T>(hp.data(), blocks, op->opcode, hc.data()));
    store_scalar(val, val_type->code, accum ? accum->opcode : OP_NONE, ws);
    return GrB_SUCCESS;
}

// ================================================================== select with a builtin GrB_IndexUnaryOp, whole-object assign
// select keeps the entries for which op(x, i, j, thunk) holds and leaves their values untouched
// (reference core/vector.py:1560-1631, core/matrix.py:2560-2630; GraphBLAS C API 2.0 section 4.3.9).  For a vector j = 0.
template <typename T>
__device__ __forceinline__ bool index_op_keeps(int op, int64_t i, int64_t j, T x, T y, int64_t yi) {
    switch (op) {
        case IOP_TRIL: return j <= i + yi;
        case IOP_TRIU: return j >= i + yi;
        case IOP_DIAG: return j == i + yi;
        case IOP_OFFDIAG: return j != i + yi;
        case IOP_COLLE: return j <= yi;
        case IOP_COLGT: return j > yi;
        case IOP_ROWLE: return i <= yi;
        case IOP_ROWGT: return i > yi;
        case IOP_VALUEEQ: return cmpop<T>(OP_EQ, x, y);
        case IOP_VALUENE: return cmpop<T>(OP_NE, x, y);
        case IOP_VALUEGT: return cmpop<T>(OP_GT, x, y);
        case IOP_VALUEGE: return cmpop<T>(OP_GE, x, y);
        case IOP_VALUELT: return cmpop<T>(OP_LT, x, y);
        case IOP_VALUELE: return cmpop<T>(OP_LE, x, y);
        case IOP_ROWINDEX: return i + yi != 0;       // index-valued ops used as a predicate: result cast to bool
        case IOP_COLINDEX: return j + yi != 0;
        case IOP_DIAGINDEX: return j - i + yi != 0;
    }
    return false;
}
static inline bool index_op_positional(int op) { return op < IOP_VALUEEQ || op > IOP_VALUELE; }

template <typename T>
__global__ void vec_select_kernel(int64_t n, int op, T y, int64_t yi, const T *__restrict__ u, const uint8_t *__restrict__ up,
                                  uint8_t *__restrict__ tp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) tp[i] = (up[i] && index_op_keeps<T>(op, i, 0, u ? u[i] : T(), y, yi)) ? 1 : 0;
}

extern "C" GrB_Info GrB_cuda_Vector_select(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                           const GrB_Vector u, const void *thunk, GrB_Type thunk_type, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(w)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "select: output vector is not initialised");
    if (!op || !valid(u) || !thunk || !thunk_type) return set_error(&w->err, GrB_NULL_POINTER, "select: null or uninitialised argument");
    if (mask && !valid(mask)) return set_error(&w->err, GrB_UNINITIALIZED_OBJECT, "select: bad mask");
    if (u->n != w->n || (mask && mask->n != w->n)) return set_error(&w->err, GrB_DIMENSION_MISMATCH, "select: sizes differ");
    const int64_t n = w->n;
    GRB_TRY(vector_ensure_arrays(u));
    const bool positional = index_op_positional(op->opcode);
    const int ct = positional ? u->type : op->type;   // the type the predicate compares in
    const void *uv = u->vals;
    void *utmp = nullptr;
    if (!positional) GRB_TRY(cast_view(&uv, &utmp, u->vals, u->type, ct, n, &w->err));
    int64_t yi = 0;
    unsigned char ybuf[8] = {0};
    host_cast(&yi, TC_INT64, thunk, thunk_type->code);
    host_cast(ybuf, ct, thunk, thunk_type->code);
    const size_t nn = (size_t)(n > 0 ? n : 1);
    void *tv = dev_alloc(nn * type_size(u->type));
    uint8_t *tp = (uint8_t *)dev_alloc(nn);
    if (!tv || !tp) { dev_free(utmp); dev_free(tv); dev_free(tp); return set_error(&w->err, GrB_OUT_OF_MEMORY, "select result"); }
    cudaError_t e = cudaSuccess;
    if (n > 0) {
        e = cudaMemcpyAsync(tv, u->vals, (size_t)n * type_size(u->type), cudaMemcpyDeviceToDevice, g_stream);
        LAUNCH_NOTE("vec_select");
        GRB_DISPATCH_TYPE(ct, T, {
            T y;
            memcpy(&y, ybuf, sizeof(T));
            vec_select_kernel<T><<<grid_for(n), 256, 0, g_stream>>>(n, op->opcode, y, yi, positional ? nullptr : (const T *)uv, u->present, tp);
        });
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    dev_free(utmp);
    if (e != cudaSuccess) { dev_free(tv); dev_free(tp); return cuda_fail(&w->err, e, "select"); }
    return vector_write_back(w, tv, tp, u->type, mask, accum, desc, true);
}

// matrix: a warp per row flags and counts the kept entries, a scan gives the row pointers, a second pass compacts
// (order within a row is preserved, so a sorted operand gives a sorted result)
template <typename T>
__global__ void __launch_bounds__(256)
mat_select_flags_kernel(int64_t nrows, int op, T y, int64_t yi, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj,
                        const T *__restrict__ Ax, uint8_t *__restrict__ keep, int64_t *__restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < nrows; w += nw) {
        const int64_t b = Ap[w], e = Ap[w + 1];
        int c = 0;
        for (int64_t k0 = b; k0 < e; k0 += 32) {
            const int64_t k = k0 + lane;
            bool kp = false;
            if (k < e) {
                kp = index_op_keeps<T>(op, w, (int64_t)Aj[k], Ax ? Ax[k] : T(), y, yi);
                keep[k] = kp ? 1 : 0;
            }
            c += __popc(__ballot_sync(0xffffffffu, kp));
        }
        if (lane == 0) cnt[w] = c;
    }
}
template <typename U>
__global__ void __launch_bounds__(256)
mat_select_compact_kernel(int64_t nrows, const int64_t *__restrict__ Ap, const int32_t *__restrict__ Aj, const U *__restrict__ Ax,
                          const uint8_t *__restrict__ keep, const int64_t *__restrict__ Tp, int32_t *__restrict__ Tj, U *__restrict__ Tx) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < nrows; w += nw) {
        const int64_t b = Ap[w], e = Ap[w + 1];
        int64_t out = Tp[w];
        for (int64_t k0 = b; k0 < e; k0 += 32) {
            const int64_t k = k0 + lane;
            const bool kp = k < e && keep[k];
            const unsigned m = __ballot_sync(0xffffffffu, kp);
            if (kp) {
                const int64_t pos = out + __popc(m & ((1u << lane) - 1u));
                Tj[pos] = Aj[k];
                Tx[pos] = Ax[k];
            }
            out += __popc(m);
        }
    }
}

extern "C" GrB_Info GrB_cuda_Matrix_select(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                           const GrB_Matrix A, const void *thunk, GrB_Type thunk_type, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(C)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "select: output matrix is not initialised");
    if (!op || !valid(A) || !thunk || !thunk_type) return set_error(&C->err, GrB_NULL_POINTER, "select: null or uninitialised argument");
    if (Mask && !valid(Mask)) return set_error(&C->err, GrB_UNINITIALIZED_OBJECT, "select: bad mask");
    OperandCsr src;
    GRB_TRY(operand_csr(&src, A, desc && desc->t0, false));
    if (C->nrows != src.nrows || C->ncols != src.ncols || (Mask && (Mask->nrows != C->nrows || Mask->ncols != C->ncols)))
        return set_error(&C->err, GrB_DIMENSION_MISMATCH, "select: C(%lldx%lld) vs A(%lldx%lld)", (long long)C->nrows, (long long)C->ncols,
                         (long long)src.nrows, (long long)src.ncols);
    const int64_t m = src.nrows, nv = A->nvals;
    const bool positional = index_op_positional(op->opcode);
    const int ct = positional ? A->type : op->type;
    int64_t yi = 0;
    unsigned char ybuf[8] = {0};
    host_cast(&yi, TC_INT64, thunk, thunk_type->code);
    host_cast(ybuf, ct, thunk, thunk_type->code);
    GrB_Matrix T = nullptr;
    GRB_TRY(matrix_new_shell(&T, A->type, m, src.ncols));
    T->csr.ptr = dev_alloc_t<int64_t>((size_t)m + 1);
    uint8_t *keep = (uint8_t *)dev_alloc((size_t)(nv > 0 ? nv : 1));
    const void *av = src.c->val;
    void *atmp = nullptr;
    GrB_Info info = (!T->csr.ptr || !keep) ? set_error(&C->err, GrB_OUT_OF_MEMORY, "select scratch") : GrB_SUCCESS;
    if (!info && !positional && nv > 0) info = cast_view(&av, &atmp, src.c->val, A->type, ct, nv, &C->err);
    const unsigned blocks = (unsigned)std::min<int64_t>((m + 7) / 8 + 1, (int64_t)g_num_sms * 16);
    if (!info) {
        cudaMemsetAsync(T->csr.ptr, 0, sizeof(int64_t) * ((size_t)m + 1), g_stream);
        if (m > 0) {
            LAUNCH_NOTE("mat_select_flags");
            GRB_DISPATCH_TYPE(ct, T_, {
                T_ y;
                memcpy(&y, ybuf, sizeof(T_));
                mat_select_flags_kernel<T_><<<blocks, 256, 0, g_stream>>>(m, op->opcode, y, yi, src.c->ptr, src.c->idx, positional ? nullptr : (const T_ *)av, keep, T->csr.ptr);
            });
        }
        info = exclusive_scan_i64(T->csr.ptr, m + 1, &C->err);
    }
    int64_t total = 0;
    if (!info) total = read_i64(T->csr.ptr + m);
    if (!info) {
        const size_t tn = (size_t)(total > 0 ? total : 1);
        T->csr.idx = dev_alloc_t<int32_t>(tn);
        T->csr.val = dev_alloc(tn * type_size(A->type));
        T->nvals = total;
        T->jumbled = (desc && desc->t0) ? false : A->jumbled;
        if (!T->csr.idx || !T->csr.val) info = set_error(&C->err, GrB_OUT_OF_MEMORY, "select result");
    }
    if (!info && total > 0) {
        LAUNCH_NOTE("mat_select_compact");
        switch (type_size(A->type)) {
            case 1: mat_select_compact_kernel<uint8_t><<<blocks, 256, 0, g_stream>>>(m, src.c->ptr, src.c->idx, (const uint8_t *)src.c->val, keep, T->csr.ptr, T->csr.idx, (uint8_t *)T->csr.val); break;
            case 2: mat_select_compact_kernel<uint16_t><<<blocks, 256, 0, g_stream>>>(m, src.c->ptr, src.c->idx, (const uint16_t *)src.c->val, keep, T->csr.ptr, T->csr.idx, (uint16_t *)T->csr.val); break;
            case 4: mat_select_compact_kernel<uint32_t><<<blocks, 256, 0, g_stream>>>(m, src.c->ptr, src.c->idx, (const uint32_t *)src.c->val, keep, T->csr.ptr, T->csr.idx, (uint32_t *)T->csr.val); break;
            default: mat_select_compact_kernel<uint64_t><<<blocks, 256, 0, g_stream>>>(m, src.c->ptr, src.c->idx, (const uint64_t *)src.c->val, keep, T->csr.ptr, T->csr.idx, (uint64_t *)T->csr.val); break;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(&C->err, e, "select");
    }
    dev_free(atmp);
    dev_free(keep);
    if (!info) info = matrix_write_back(C, T, Mask, accum, desc);
    GrB_Matrix_free(&T);
    return info;
}

// w<mask> accum= u / C<Mask> accum= A over GrB_ALL: the standard write-back with T = a copy of the operand
extern "C" GrB_Info GrB_Vector_assign(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Vector u,
                                      const GrB_Index *indices, GrB_Index, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(w)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "assign: output vector is not initialised");
    if (!valid(u)) return set_error(&w->err, GrB_NULL_POINTER, "assign: null or uninitialised argument");
    if (mask && !valid(mask)) return set_error(&w->err, GrB_UNINITIALIZED_OBJECT, "assign: bad mask");
    if (indices != nullptr && indices != GrB_ALL) return set_error(&w->err, GrB_NOT_IMPLEMENTED, "assign: only GrB_ALL is on this backend's path");
    if (u->n != w->n || (mask && mask->n != w->n)) return set_error(&w->err, GrB_DIMENSION_MISMATCH, "assign: sizes differ");
    GRB_TRY(vector_ensure_arrays(u));
    const int64_t n = w->n;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    void *tv = dev_alloc(nn * type_size(u->type));
    uint8_t *tp = (uint8_t *)dev_alloc(nn);
    if (!tv || !tp) { dev_free(tv); dev_free(tp); return set_error(&w->err, GrB_OUT_OF_MEMORY, "assign"); }
    if (n > 0) {
        CUDA_TRY(&w->err, cudaMemcpyAsync(tv, u->vals, (size_t)n * type_size(u->type), cudaMemcpyDeviceToDevice, g_stream));
        CUDA_TRY(&w->err, cudaMemcpyAsync(tp, u->present, (size_t)n, cudaMemcpyDeviceToDevice, g_stream));
    }
    return vector_write_back(w, tv, tp, u->type, mask, accum, desc, true);
}
extern "C" GrB_Info GrB_Matrix_assign(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_Matrix A, const GrB_Index *rows,
                                      GrB_Index, const GrB_Index *cols, GrB_Index, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(C)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "assign: output matrix is not initialised");
    if (!valid(A)) return set_error(&C->err, GrB_NULL_POINTER, "assign: null or uninitialised argument");
    if ((rows != nullptr && rows != GrB_ALL) || (cols != nullptr && cols != GrB_ALL))
        return set_error(&C->err, GrB_NOT_IMPLEMENTED, "assign: only GrB_ALL is on this backend's path");
    OperandCsr src;
    GRB_TRY(operand_csr(&src, A, desc && desc->t0, false));
    if (C->nrows != src.nrows || C->ncols != src.ncols || (Mask && (Mask->nrows != C->nrows || Mask->ncols != C->ncols)))
        return set_error(&C->err, GrB_DIMENSION_MISMATCH, "assign: C(%lldx%lld) vs A(%lldx%lld)", (long long)C->nrows, (long long)C->ncols,
                         (long long)src.nrows, (long long)src.ncols);
    GrB_Matrix T = nullptr;
    GRB_TRY(matrix_with_pattern(&T, A->type, src, A->nvals, (desc && desc->t0) ? false : A->jumbled, &C->err));
    if (A->nvals > 0)
        CUDA_TRY(&C->err, cudaMemcpyAsync(T->csr.val, src.c->val, type_size(A->type) * (size_t)A->nvals, cudaMemcpyDeviceToDevice, g_stream));
    GrB_Info info = matrix_write_back(C, T, Mask, accum, desc);
    GrB_Matrix_free(&T);
    return info;
}
