// comm.cu -- multi-GPU exchange (replaces 2decomp&FFT's update_halo / transpose_* and the scalar
// mpi_allreduce calls).  Filled in by the slab-decomposition milestone; a single-rank context never
// reaches these functions.
#include "fen_internal.cuh"

namespace fen {

struct Comm {};

int halo_exchange(fen_ctx* c, double* const* f, int n) {
    (void)c; (void)f; (void)n;
    return set_error(FEN_ERR_COMM, "multi-GPU halo exchange: fen_gpu_comm_connect has not been called");
}
int comm_allreduce(fen_ctx* c, double* d_vals, int n, int op) {
    (void)c; (void)d_vals; (void)n; (void)op;
    return set_error(FEN_ERR_COMM, "multi-GPU reduction: fen_gpu_comm_connect has not been called");
}
int comm_transpose_fwd(fen_ctx* c) {
    (void)c;
    return set_error(FEN_ERR_COMM, "multi-GPU transpose: fen_gpu_comm_connect has not been called");
}
int comm_transpose_bwd(fen_ctx* c) { return comm_transpose_fwd(c); }
void comm_destroy(fen_ctx* c) { (void)c; }

}  // namespace fen

extern "C" {
int fen_gpu_comm_handle_bytes(void) { return 0; }
int fen_gpu_comm_export(fen_ctx* c, void* out) { (void)c; (void)out; return fen::set_error(FEN_ERR_COMM, "not built yet"); }
int fen_gpu_comm_connect(fen_ctx* c, const void* all) { (void)c; (void)all; return fen::set_error(FEN_ERR_COMM, "not built yet"); }
}
