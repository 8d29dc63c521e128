// conv.cu -- scatter half of the time-major 1-D convolution backward
// (taiyaki/layers.py:744-850 `Convolution`; forward and the two weight/column
// GEMMs are library GEMMs on a window-gathered matrix, see layers.py
// `_ConvTimeMajor`).  Given the column gradient dcols [T_out][N][C*k] this
// sums, for every input sample, the <= ceil(k/stride) windows that contain it:
//     dx[t][n][c] = sum_{j = (t+pad) mod stride, step stride, j < k}
//                       dcols[(t + pad - j) / stride][n][c*k + j]
// (a gather, so no atomics and a deterministic sum).  HBM-bound: it reads
// dcols once and writes dx once.
#include "common.cuh"

namespace ty {

__global__ void __launch_bounds__(256) col2im_tm_kernel(const float *__restrict__ dcols, int Tout,
                                                        int N, int C, int k, int stride, int pad,
                                                        int T, float *__restrict__ dx) {
    const size_t total = (size_t)T * N * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t tn = i / C;
        const int n = (int)(tn % N);
        const int t = (int)(tn / N);
        const int tau = t + pad;
        float acc = 0.f;
        for (int j = tau % stride; j < k; j += stride) {
            const int to = (tau - j) / stride;
            if (to >= 0 && to < Tout && tau - j >= 0)
                acc += dcols[((size_t)to * N + n) * ((size_t)C * k) + (size_t)c * k + j];
        }
        dx[i] = acc;
    }
}

}  // namespace ty

using namespace ty;

extern "C" int ty_col2im_time_major(const float *dcols, int Tout, int N, int C, int k, int stride,
                                    int pad_left, int T, float *dx, void *stream) {
    if (!dcols || !dx || Tout <= 0 || N <= 0 || C <= 0 || k <= 0 || stride <= 0 || T <= 0 ||
        pad_left < 0) {
        set_error("ty_col2im_time_major: bad argument");
        return TY_EINVAL;
    }
    const size_t total = (size_t)T * N * C;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    col2im_tm_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dcols, Tout, N, C, k, stride, pad_left, T, dx);
    return check_launch("col2im_tm_kernel");
}
