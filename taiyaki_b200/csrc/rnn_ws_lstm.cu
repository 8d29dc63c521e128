// rnn_ws_lstm.cu -- LSTM instantiations of the warp-specialised recurrence (rnn_ws.cuh)
#include "rnn_ws.cuh"

namespace ty {
int launch_rnn_ws_lstm(bool backward, const RnnWsArgs &a, int H, cudaStream_t s) {
    return launch_rnn_ws<kLstm>(backward, a, H, s);
}
}  // namespace ty
