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
        const bool p = up[i] != 0;
        T z = T();
        if (p) z = mode == 0 ? unop<T>(op, u[i]) : mode == 1 ? binop<T>(op, scalar, u[i]) : binop<T>(op, u[i], scalar);
        tv[i] = z;
        tp[i] = p ? 1 : 0;
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
        if (up[i]) { acc = binop<T>(op, acc, u[i]); cnt++; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        T other;
        if constexpr (sizeof(T) == 8) {
            unsigned long long b; memcpy(&b, &acc, 8); b = __shfl_down_sync(0xffffffffu, b, o); memcpy(&other, &b, 8);
        } else {
            unsigned int b = 0; memcpy(&b, &acc, sizeof(T)); b = __shfl_down_sync(0xffffffffu, b, o); memcpy(&other, &b, sizeof(T));
        }
        acc = binop<T>(op, acc, other);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { s_val[warp] = acc; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        T a = s_val[0];
        unsigned long long c = s_cnt[0];
        for (int q = 1; q < (int)(blockDim.x >> 5); q++) { a = binop<T>(op, a, s_val[q]); c += s_cnt[q]; }
        partial[blockIdx.x] = a;
        pcount[blockIdx.x] = c;
    }
}

// a scalar of any builtin type in its widest faithful representation
struct WideScalar { double d; int64_t l; uint64_t ul; bool isf, isu; };
template <typename T> static WideScalar fold_partials(const unsigned char *hp, int blocks, int opcode) {
    T acc = monoid_identity<T>(opcode);
    for (int b = 0; b < blocks; b++) {
        T x;
        memcpy(&x, hp + (size_t)b * sizeof(T), sizeof(T));
        acc = binop<T>(opcode, acc, x);
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
    GRB_DISPATCH_TYPE(mt, T, ws = fold_partials<T>(hp.data(), blocks, op->opcode));
    store_scalar(val, val_type->code, accum ? accum->opcode : OP_NONE, ws);
    return GrB_SUCCESS;
}
