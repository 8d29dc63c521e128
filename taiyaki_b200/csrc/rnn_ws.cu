// rnn_ws.cu -- C entry points of the warp-specialised recurrence (kernels: rnn_ws.cuh;
// instantiations per cell: rnn_ws_lstm.cu, rnn_ws_gru.cu, compiled in parallel).
#include "rnn_ws.cuh"

using namespace ty;

extern "C" int ty_rnn_forward_um(int cell, const float *xproj, const float *bias,
                                 const float *w_hh, int T, int N, int H, int reverse, float *y,
                                 void *y_bf16, void *reserve, void *stream) {
    if (!xproj || !w_hh || !y || !reserve || T <= 0 || N <= 0 || H <= 0 || (cell != kLstm && cell != kGru)) {
        set_error("ty_rnn_forward_um: bad argument (cell=%d T=%d N=%d H=%d)", cell, T, N, H);
        return TY_EINVAL;
    }
    RnnWsArgs a{};
    a.xproj = xproj; a.bias = bias; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse;
    a.y = y; a.y16 = static_cast<__nv_bfloat16 *>(y_bf16);
    a.gates = static_cast<float *>(reserve);
    a.cstate = a.gates + (size_t)T * N * H * 4;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return cell == kLstm ? launch_rnn_ws_lstm(false, a, H, s) : launch_rnn_ws_gru(false, a, H, s);
}

extern "C" int ty_rnn_backward_um(int cell, const float *dy, const float *w_hh, int T, int N,
                                  int H, int reverse, const float *y, const void *reserve,
                                  void *dxproj_bf16, void *dhid_bf16, float *dbias, void *stream) {
    if (!dy || !w_hh || !reserve || !dxproj_bf16 || T <= 0 || N <= 0 || H <= 0 ||
        (cell != kLstm && cell != kGru) || (cell == kGru && (!y || !dhid_bf16))) {
        set_error("ty_rnn_backward_um: bad argument (cell=%d T=%d N=%d H=%d)", cell, T, N, H);
        return TY_EINVAL;
    }
    RnnWsArgs a{};
    a.dy = dy; a.w_hh = w_hh; a.T = T; a.N = N; a.reverse = reverse;
    a.y = const_cast<float *>(y);
    a.gates = const_cast<float *>(static_cast<const float *>(reserve));
    a.cstate = a.gates + (size_t)T * N * H * 4;
    a.dx16 = static_cast<__nv_bfloat16 *>(dxproj_bf16);
    a.dhid16 = static_cast<__nv_bfloat16 *>(dhid_bf16);
    a.dbias = dbias;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return cell == kLstm ? launch_rnn_ws_lstm(true, a, H, s) : launch_rnn_ws_gru(true, a, H, s);
}

extern "C" int ty_rnn_um_supported(int hidden) {
    switch (hidden) {
        case 32: case 64: case 96: case 128: case 160: case 192: case 224: case 256: case 320: case 384: case 448:
            return 1;
        default:
            return 0;
    }
}
