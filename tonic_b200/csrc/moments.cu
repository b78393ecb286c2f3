// K16 running observation statistics.
// Reference: tonic/torch/normalizers/mean_stds.py:44-48 (record: per-sample
// sum / sum of squares) and :50-74 (update: count-weighted merge, std floor
// 1e-2).  The reference accumulates float32 sums sequentially over workers; here
// the sums are accumulated in float64 (more accurate, same limit).
#include "common.cuh"

namespace tb {

constexpr int kMomCols = 4;   // columns per thread when dim > 256

__global__ void __launch_bounds__(256)
moments_record_kernel(const float* __restrict__ x, int64_t n_rows, int dim, double* sums,
                      int64_t rows_per_block) {
    extern __shared__ double sacc[];                 // [R][2*dim] when dim <= 256
    const int64_t r0 = blockIdx.x * rows_per_block;
    const int64_t r1 = min(n_rows, r0 + rows_per_block);
    if (r0 >= r1) return;
    if (dim <= 256) {
        const int R = 256 / dim;                     // rows handled per pass
        const int active = R * dim;
        double s = 0.0, q = 0.0;
        const int col = threadIdx.x % dim, rl = threadIdx.x / dim;
        if ((int)threadIdx.x < active) {
            for (int64_t r = r0 + rl; r < r1; r += R) {
                const double v = (double)x[r * dim + col];
                s += v;
                q += v * v;
            }
            sacc[(size_t)rl * 2 * dim + col] = s;
            sacc[(size_t)rl * 2 * dim + dim + col] = q;
        }
        __syncthreads();
        if ((int)threadIdx.x < dim) {
            double ts = 0.0, tq = 0.0;
            for (int r = 0; r < R; ++r) {
                ts += sacc[(size_t)r * 2 * dim + threadIdx.x];
                tq += sacc[(size_t)r * 2 * dim + dim + threadIdx.x];
            }
            atomicAdd(&sums[threadIdx.x], ts);
            atomicAdd(&sums[dim + threadIdx.x], tq);
        }
    } else {
        double s[kMomCols], q[kMomCols];
#pragma unroll
        for (int c = 0; c < kMomCols; ++c) s[c] = q[c] = 0.0;
        for (int64_t r = r0; r < r1; ++r) {
#pragma unroll
            for (int c = 0; c < kMomCols; ++c) {
                const int col = threadIdx.x + c * 256;
                if (col < dim) {
                    const double v = (double)x[r * dim + col];
                    s[c] += v;
                    q[c] += v * v;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kMomCols; ++c) {
            const int col = threadIdx.x + c * 256;
            if (col < dim) {
                atomicAdd(&sums[col], s[c]);
                atomicAdd(&sums[dim + col], q[c]);
            }
        }
    }
    if (threadIdx.x == 0) atomicAdd(&sums[2 * dim], (double)(r1 - r0));
}

__global__ void moments_update_kernel(double* sums, float* running, double* count,
                                      float* mean_out, float* std_out, int dim, float eps) {
    const double new_count = sums[2 * dim];
    if (new_count == 0.0) return;
    const double old_count = count[0];
    const double total = old_count + new_count;
    // mean_stds.py:54-57: python-float weights applied to float32 arrays
    const float w_old = (float)(old_count / total), w_new = (float)(new_count / total);
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        const float new_mean = (float)(sums[j] / new_count);
        const float new_mean_sq = (float)(sums[dim + j] / new_count);
        const float mean = __fadd_rn(__fmul_rn(w_old, running[j]), __fmul_rn(w_new, new_mean));
        const float mean_sq =
            __fadd_rn(__fmul_rn(w_old, running[dim + j]), __fmul_rn(w_new, new_mean_sq));
        running[j] = mean;
        running[dim + j] = mean_sq;
        const float var = fmaxf(__fsub_rn(mean_sq, __fmul_rn(mean, mean)), 0.0f);   // :66-67
        mean_out[j] = mean;
        std_out[j] = fmaxf(sqrtf(var), eps);                                         // :68-69
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * dim + 1; j += blockDim.x) sums[j] = 0.0;
    if (threadIdx.x == 0) count[0] = total;
}

}  // namespace tb

extern "C" int tb_moments_record(const float* d_x, int64_t n_rows, int32_t dim,
                                 double* d_sums, void* stream) {
    tb::ProfScope prof_scope("tb_moments_record", stream);
    TB_REQUIRE(d_x && d_sums && n_rows > 0 && dim > 0 && dim <= 256 * tb::kMomCols, TB_EINVAL,
               "tb_moments_record: bad arguments (dim=%d)", dim);
    int64_t rows_per_block = 512;
    int64_t blocks = (n_rows + rows_per_block - 1) / rows_per_block;
    const size_t smem = dim <= 256 ? (size_t)(256 / dim) * 2 * dim * sizeof(double) : 0;
    tb::moments_record_kernel<<<(int)blocks, 256, smem, tb::as_stream(stream)>>>(
        d_x, n_rows, dim, d_sums, rows_per_block);
    return tb::check_launch("tb_moments_record");
}

extern "C" int tb_moments_update(double* d_sums, float* d_running, double* d_count,
                                 float* d_mean, float* d_std, int32_t dim, float eps,
                                 void* stream) {
    tb::ProfScope prof_scope("tb_moments_update", stream);
    TB_REQUIRE(d_sums && d_running && d_count && d_mean && d_std && dim > 0, TB_EINVAL,
               "tb_moments_update: bad arguments");
    tb::moments_update_kernel<<<1, 256, 0, tb::as_stream(stream)>>>(
        d_sums, d_running, d_count, d_mean, d_std, dim, eps);
    return tb::check_launch("tb_moments_update");
}
