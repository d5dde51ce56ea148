// matrix.h -- device-resident block-CSR matrix, vectors and resources of the engine.
// Layout follows the reference's Matrix<TConfig> (include/matrix.h:65-200): row_offsets[n+1],
// col_indices[nnz], values[nnz*bs] row-major blocks, diag[n] = index of each row's diagonal
// entry; an external diagonal (AMGX_matrix_upload_all diag_data != NULL) is stored after the
// off-diagonal values at block index nnz + row (src/matrix.cu:713-721).
#pragma once
#include "base.h"
#include "config.h"

namespace amgxb {

struct DistManager;   // dist.h

struct Resources {
    std::shared_ptr<Config> cfg;
    int device = 0;
    cudaStream_t stream = nullptr;       // compute stream of this engine
    cudaStream_t side_stream = nullptr;  // halo exchange / copies
    int num_sms = 148;
    int rank = 0, world = 1;
    void *nccl_comm = nullptr;           // ncclComm_t when world > 1
    void *p2p = nullptr;                 // PeerWindow (p2p.cu): CUDA-IPC mapped windows of all ranks, null -> NCCL send/recv path
    ~Resources();
};

// How the scalar-CSR kernels tile a matrix: fixed TILE_ROWS consecutive rows per tile; per tile the
// col/val ranges are staged into shared memory with TMA bulk copies (k_spmv.cu).
struct TilePlan {
    int tile_rows = 0;          // rows per tile (== CTA size)
    int num_tiles = 0;
    int max_tile_nnz = 0;       // capacity each smem stage must hold (already padded for alignment)
    int stages = 0;             // pipeline depth chosen for this matrix
    size_t smem_bytes = 0;      // dynamic shared memory per CTA
    bool use_tiles = false;     // false -> generic warp-per-row kernel
    int split = 0;              // distributed: first row (multiple of 4) of the segment that reads halo columns
    int unroll = 4;             // gathers in flight per consumer step (4 or 8)
    int ctas_per_sm = 1;        // resident CTAs per SM the grid is sized for
    bool use_perm = false;      // rows of a tile are handed to the threads sorted by length (Matrix::tile_perm)
    int max_row_nnz = 0;
};

// Sliding-window plan of the banded-matrix kernel (k_spmv_win.cu): x[r0 - w, r0 + stages * T + w) of a CTA's tile range in a shared-memory ring
struct WinPlan {
    bool on = false;
    DevBuf<unsigned char> sv;   // values in sliced-ELL order (per tile: rows sorted by length, 32-row slices stored entry-major)
    DevBuf<unsigned short> so;  // column - row + 32768 in the same order (0: use the 32-bit column of the CSR arrays)
    DevBuf<unsigned char> perm; // sorted position -> local row, per tile
    DevBuf<long long> tbase;    // first entry of every tile (num_tiles + 1)
    DevBuf<int> sbase;          // first entry of every slice relative to its tile
    DevBuf<unsigned short> smeta; // local row | length << 8 of every sorted position (streaming form of the kernel)
    DevBuf<int> slens;          // entries per row of every slice
    bool stream = false;        // per-warp chunk queues instead of whole-tile stages
    DevBuf<int> long_rows;      // rows longer than lmax: left out of the sliced-ELL copy, summed by the warp-per-row side kernel
    int num_long = 0, lmax = 0;
    int cap = 0;                // entries of the longest tile
    int ring = 0, w = 0, stages = 0, tiles_per_cta = 0, grid = 0;
    size_t smem_bytes = 0, stream_smem_bytes = 0;
    double inside = 0.0;        // fraction of the entries whose column lies within +-w of the row
};

// Optional compressed column stream of the tile kernels (k_spmv_enc.cu, AMGXB_COLENC=1; experimental, default off).  Per tile either
// 8-bit codes into a dictionary of (column - row) offsets, 16-bit offsets from the tile's smallest column, or the raw 32-bit columns.
struct ColEnc {
    bool on = false;
    DevBuf<unsigned char> codes;    // tile t's code segment starts at byte align16(2 * sa_t) + 32 * t and covers the aligned entry range [sa_t, ea_t)
    DevBuf<int> dict;               // 256 ints per tile: the sorted offsets (8-bit), or the base column in entry 0 (16-bit)
    DevBuf<int> meta;               // 4 ints per tile: column encoding (0 raw, 1 dict8, 2 off16), its dictionary length (padded to a multiple of 4),
                                    // value encoding (0 raw, 1 dict8), its dictionary length
    DevBuf<unsigned char> vcodes;   // value codes: tile t's segment starts at byte align16(sa_t) + 32 * t
    DevBuf<unsigned char> vdict;    // 256 matrix values per tile: the distinct values of the tile, ascending bit patterns
    bool values_encoded = false;    // the value codes exist and must follow in-place changes of the values (csr_values_changed)
    size_t smem_bytes = 0;          // dynamic shared memory of the encoded kernel
    int tiles_dict8 = 0, tiles_off16 = 0, tiles_raw = 0, tiles_val8 = 0;
    int max_dlen = 0, max_vdlen = 0;               // longest (padded) dictionaries of the level
    DevBuf<unsigned char> pcodes;   // pair codes: one byte per entry into the tile's (column offset, value) table (tiles with pmeta > 0)
    DevBuf<unsigned char> pdict;    // 256 pairs per tile
    DevBuf<int> pmeta;              // per tile: (padded) length of its pair table, 0 = not pair-coded
    int tiles_pair = 0, max_pdlen = 0;
    DevBuf<unsigned char> rowcodes; // row-pattern ids: one byte per row (tiles with rmeta > 0)
    DevBuf<unsigned char> rpat;     // 64 patterns of 16 bytes per tile (7 pair codes + length)
    DevBuf<int> rmeta;              // per tile: number of row patterns, 0 = rows keep their per-entry pair codes
    int tiles_rowpat = 0, max_rplen = 0;
    DevBuf<int> xahead;             // per tile: largest column offset of a dictionary-coded tile (x rows the producer prefetches into L2), else 0
    int col_w = 4, val_w = 8, dict_cap = 0, vdict_cap = 0;   // stage layout of the level (finalize_layout)
    int stages = 2, ctas_per_sm = 1;
    int num_tiles = 0;                             // tiles over all row segments ([0, split) then [split, n) on a row-partitioned matrix)
};

struct Matrix {
    std::shared_ptr<Resources> rsc;
    int mode = AMGX_mode_dDDI;
    Prec mat_prec = Prec::F64, vec_prec = Prec::F64;
    int n = 0;            // owned block rows
    int n_cols = 0;       // owned + halo block columns (== n on a single GPU)
    int split_row = 0;    // distributed: number of interior rows (rows [0, split_row) reference no halo column)
    int nnz = 0;          // stored blocks (without external diagonal)
    int bx = 1, by = 1;
    bool has_ext_diag = false;
    bool merged_ext_diag = false;   // scalar matrix uploaded with diag_data: merged into CSR, diagonal first
    ColEnc colenc;
    WinPlan win;
    bool dist_pending = false;      // comm maps were supplied; the next upload_all is a local distributed upload
    struct CommMaps {               // AMGX_matrix_comm_from_maps[_one_ring]: neighbours, rows to send, halo columns to receive into
        std::vector<int> neighbors;
        std::vector<std::vector<int>> send, recv;
    };
    std::shared_ptr<CommMaps> comm_maps;
    bool initialized = false;
    int level = 0;

    DevBuf<int> row_ptr, col_idx, diag_idx;
    DevVec values;        // (nnz [+ n if ext diag]) * bx*by scalars of mat_prec
    TilePlan plan;
    DevBuf<unsigned char> tile_perm;   // plan.use_perm: for every tile (numbered over the row segments) and thread, the row of the tile it takes

    std::shared_ptr<DistManager> dist;   // null on a single GPU

    // colouring (multicolour smoothers)
    int num_colors = 0;
    DevBuf<int> row_colors, sorted_rows_by_color;
    std::vector<int> color_offsets;      // host, [num_colors+1]
    bool user_coloring = false;

    int bs() const { return bx * by; }
    cudaStream_t stream() const { return rsc->stream; }
    void compute_diag_and_plan();        // diag_idx + tile plan (k_spmv.cu / k_setup.cu)
};

struct Vector {
    std::shared_ptr<Resources> rsc;
    int mode = AMGX_mode_dDDI;
    Prec prec = Prec::F64;
    int n = 0;            // block entries
    int block_dim = 1;
    DevVec data;          // n*block_dim scalars (+ halo tail when bound to a distributed matrix)
    std::shared_ptr<DistManager> dist;
    bool user_order = true;   // distributed: data is in the caller's (partition) order
};

}  // namespace amgxb
