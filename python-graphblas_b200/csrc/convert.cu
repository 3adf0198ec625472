// convert.cu -- typecasts between the 11 builtin types (C semantics; float->int saturates, NaN -> 0,
// anything -> BOOL is (x != 0)) and index-width conversions at the uint64 boundary.
#include "grb_ops.cuh"

template <typename S, typename D> __device__ __forceinline__ D cast_one(S x) {
    if constexpr (is_gbool<D>::value) {
        return gbool(truthy<S>(x));
    } else if constexpr (is_gbool<S>::value) {
        return (D)(x.v ? 1 : 0);
    } else if constexpr (std::is_floating_point<S>::value && std::is_integral<D>::value) {
        if (x != x) return (D)0;
        if constexpr (std::is_unsigned<D>::value) {
            if (x <= (S)0) return (D)0;
            if (sizeof(D) == 8) return (x >= (S)18446744073709551615.0) ? (D)~(D)0 : (D)(unsigned long long)x;
            return (x >= (S)(D) ~(D)0) ? (D)~(D)0 : (D)(unsigned long long)x;
        } else {
            const long long lo = (sizeof(D) == 8) ? LLONG_MIN : -(1ll << (sizeof(D) * 8 - 1));
            const long long hi = (sizeof(D) == 8) ? LLONG_MAX : (1ll << (sizeof(D) * 8 - 1)) - 1;
            if (x <= (S)lo) return (D)lo;
            if (x >= (S)hi) return (D)hi;
            return (D)(long long)x;
        }
    } else {
        return (D)x;
    }
}

template <typename S, typename D>
__global__ void cast_kernel(D *__restrict__ dst, const S *__restrict__ src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = cast_one<S, D>(src[i]);
}

template <typename S> static GrB_Info cast_from(void *dst, int dst_type, const S *src, int64_t n, std::string *err) {
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)g_num_sms * 16);
    if (blocks < 1) blocks = 1;
    LAUNCH_NOTE("cast");
    GRB_DISPATCH_TYPE(dst_type, D, (cast_kernel<S, D><<<blocks, 256, 0, g_stream>>>((D *)dst, src, n)));
    CUDA_TRY(err, cudaGetLastError());
    return GrB_SUCCESS;
}

GrB_Info cast_array(void *dst, int dst_type, const void *src, int src_type, int64_t n, std::string *err) {
    if (n <= 0) return GrB_SUCCESS;
    if (dst_type == src_type) {
        CUDA_TRY(err, cudaMemcpyAsync(dst, src, (size_t)n * type_size(src_type), cudaMemcpyDeviceToDevice, g_stream));
        return GrB_SUCCESS;
    }
    GrB_Info info = GrB_PANIC;
    GRB_DISPATCH_TYPE(src_type, S, info = cast_from<S>(dst, dst_type, (const S *)src, n, err));
    return info;
}

GrB_Info cast_view(const void **out, void **tmp, const void *src, int src_type, int dst_type, int64_t n,
                   std::string *err) {
    *tmp = nullptr;
    if (src_type == dst_type || n <= 0) {
        *out = src;
        return GrB_SUCCESS;
    }
    void *t = dev_alloc((size_t)n * type_size(dst_type));
    if (!t) return set_error(err, GrB_OUT_OF_MEMORY, "typecast scratch");
    GrB_Info info = cast_array(t, dst_type, src, src_type, n, err);
    if (info) { dev_free(t); return info; }
    *tmp = t;
    *out = t;
    return GrB_SUCCESS;
}
