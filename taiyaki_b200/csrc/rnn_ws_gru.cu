// rnn_ws_gru.cu -- GRU instantiations of the warp-specialised recurrence (rnn_ws.cuh)
#include "rnn_ws.cuh"

namespace ty {
int launch_rnn_ws_gru(bool backward, const RnnWsArgs &a, int H, cudaStream_t s) {
    return launch_rnn_ws<kGru>(backward, a, H, s);
}
}  // namespace ty
