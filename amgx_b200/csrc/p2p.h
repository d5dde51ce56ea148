// p2p.h -- NVLink peer-memory halo exchange and scalar all-reduce (p2p.cu).  See the header of p2p.cu.
#pragma once
#include "base.h"
#include "matrix.h"
#include "kernels.h"

namespace amgxb {

constexpr int P2P_MAX_NEIGHBORS = 32;   // neighbours per manager on the peer-memory path (more: NCCL path)
constexpr int P2P_MAX_WORLD = 32;       // lanes of the all-reduce kernel

struct P2PLink;                         // per-manager device state (p2p.cu)

void p2p_init(Resources *rsc);                      // collective, after the NCCL communicator exists: window allocation + CUDA IPC exchange
void p2p_shutdown(Resources *rsc);
bool p2p_available(const Resources *rsc);
void p2p_manager_setup(const Matrix &A);            // collective: receive window of A.dist, addresses exchanged with the neighbours
bool p2p_exchange_blocking(const Matrix &A, void *x, Prec prec, int bsize, cudaStream_t s);   // push + flags + wait + unpack: one kernel on stream s
// op 0 sum / 2 max; post 0: FinOp epilogue, post 1: norm epilogue (sqrt when do_sqrt; host mirror when mirror)
bool p2p_allreduce_scalar(const Matrix &A, const ReduceCtx &red, int slot, int op, int post, int fin_op, int do_sqrt, bool mirror, cudaStream_t s);

}  // namespace amgxb
