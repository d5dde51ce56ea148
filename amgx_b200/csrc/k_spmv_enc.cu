// k_spmv_enc.cu -- EXPERIMENTAL (AMGXB_COLENC=1, default off; written without a device): the CSR tile kernel with a
// compressed column stream.  Same pipeline, same per-row FMA order and same epilogues as csr_tile_kernel (k_spmv.cu), hence the same
// bits; what changes is how many bytes of column information a tile moves through TMA:
//   enc 1  8-bit codes + a per-tile dictionary of the distinct (column - row) offsets (<= 256): 12 -> 9 B per entry
//          (a 7-point stencil has 7 offsets; the first aggregation levels a few dozen),
//   enc 2  16-bit offsets from the tile's smallest column (coarser, irregular levels): 12 -> 10 B per entry,
//   enc 0  the raw 32-bit columns (fallback, identical traffic to the plain kernel).
// The encoding is chosen per tile by colenc_build_kernel at plan time; it only reads row_ptr / col_idx, so value updates
// (AMGX_matrix_replace_coefficients) leave it valid.
// AMGXB_COLENC bit 1 (values 2, 3) adds the same idea for the VALUES ("value indexing"): a tile whose entries take <= 256 distinct bit
// patterns (stencil matrices and their Galerkin products: 2 on the fine 7-point level, a few dozen below) moves 8-bit codes plus a
// dictionary of the exact values instead of 8 bytes per entry -- lossless, the FMA operands are bit-identical.  Together: 12 -> 2 B
// per entry.  The value codes follow the values: csr_values_changed() rebuilds them after replace_coefficients / in-place scaling.
// Not used for: distributed matrices (two row segments), the aggregation-fused prolongation gather, the long-row fallback.
#include "kernels.h"
#include <climits>
#include <map>

namespace amgxb {
namespace {

#include "tile_common.cuh"

constexpr int DICT_SLOTS = 256;          // dictionary entries per tile (ints)
constexpr int HASH_SLOTS = 1024;

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
// byte offset of tile `t`'s code segment; sa = first (4-aligned) entry the tile stages.  2 bytes per entry are reserved whatever the
// encoding, 32 bytes of slack per tile absorb the 16-byte rounding of the copies and the <= 3 entries consecutive tiles share.
__host__ __device__ inline size_t code_offset(int sa, int t) { return align16((size_t)2 * (size_t)sa) + (size_t)32 * (size_t)t; }
// the same for the value codes (1 byte per entry)
__host__ __device__ inline size_t vcode_offset(int sa, int t) { return align16((size_t)sa) + (size_t)32 * (size_t)t; }
constexpr int META = 4;                  // ints per tile: column encoding, column dictionary length, value encoding, value dictionary length

// ---------------------------------------------------------------------------------------------
// plan: one CTA per tile (grid-stride), TILE_ROWS threads
// ---------------------------------------------------------------------------------------------
template <int TILE_ROWS>
__global__ void __launch_bounds__(TILE_ROWS) colenc_build_kernel(const int *__restrict__ rp, const int *__restrict__ ci, int row0, int n, int num_tiles, int tile_base,
                                                                 unsigned char *codes, int *dict, int *meta, int *stats)
{
    __shared__ int keys[HASH_SLOTS];
    __shared__ int list[DICT_SLOTS], sorted[DICT_SLOTS];
    __shared__ int s_count, s_min, s_max, s_overflow;
    const int tid = threadIdx.x;
    for (int ltile = blockIdx.x; ltile < num_tiles; ltile += gridDim.x) {
        const int tile = tile_base + ltile;             // tiles are numbered over the row segments of the matrix: [0, split) then [split, n)
        const int r0 = row0 + ltile * TILE_ROWS, r1 = min(r0 + TILE_ROWS, n);
        const int nz0 = rp[r0], nz1 = rp[r1];
        const int sa = nz0 & ~3, ea = (nz1 + 3) & ~3;
        for (int i = tid; i < HASH_SLOTS; i += TILE_ROWS) keys[i] = INT_MAX;
        if (tid == 0) { s_count = 0; s_min = INT_MAX; s_max = INT_MIN; s_overflow = 0; }
        __syncthreads();
        const int row = r0 + tid;
        const bool active = row < r1;
        // ---- pass 1: column range of the tile, set of distinct offsets
        if (active) {
            for (int k = rp[row]; k < rp[row + 1]; k++) {
                const int c = ci[k], delta = c - row;
                atomicMin(&s_min, c);
                atomicMax(&s_max, c);
                if (delta == INT_MAX) { s_overflow = 1; continue; }     // the empty-slot sentinel cannot be a key
                unsigned h = ((unsigned)delta * 2654435761u) >> 22;      // 10 bits
                for (int probe = 0; probe < HASH_SLOTS; probe++) {
                    if (*(volatile int *)&s_overflow) break;
                    const int old = atomicCAS(&keys[h], INT_MAX, delta);
                    if (old == INT_MAX) {
                        if (atomicAdd(&s_count, 1) >= DICT_SLOTS) s_overflow = 1;
                        break;
                    }
                    if (old == delta) break;
                    h = (h + 1) & (HASH_SLOTS - 1);
                }
            }
        }
        __syncthreads();
        const bool use_dict = !s_overflow && s_count <= DICT_SLOTS && nz1 > nz0;
        const bool use_off16 = !use_dict && nz1 > nz0 && (long long)s_max - (long long)s_min < 65536ll;
        const int count = s_count;
        __syncthreads();
        unsigned char *seg = codes + code_offset(sa, tile);
        if (use_dict) {
            // ---- compact the hash set, rank-sort it (ascending: the dictionary, hence the codes, do not depend on insertion order)
            if (tid == 0) s_count = 0;
            __syncthreads();
            for (int i = tid; i < HASH_SLOTS; i += TILE_ROWS)
                if (keys[i] != INT_MAX) list[atomicAdd(&s_count, 1)] = keys[i];
            __syncthreads();
            for (int i = tid; i < count; i += TILE_ROWS) {
                const int v = list[i];
                int rank = 0;
                for (int j = 0; j < count; j++) rank += (list[j] < v);
                sorted[rank] = v;
            }
            __syncthreads();
            const int padded = (count + 3) & ~3;
            for (int i = tid; i < padded; i += TILE_ROWS) dict[(size_t)tile * DICT_SLOTS + i] = sorted[min(i, count - 1)];
            for (int k = sa + tid; k < ea; k += TILE_ROWS)
                if (k < nz0 || k >= nz1) seg[k - sa] = 0;                 // alignment padding: entries of neighbouring tiles, never decoded
            if (active) {
                for (int k = rp[row]; k < rp[row + 1]; k++) {
                    const int delta = ci[k] - row;
                    int lo = 0, hi = count - 1;                           // binary search in the sorted dictionary
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (sorted[mid] < delta) lo = mid + 1;
                        else hi = mid;
                    }
                    seg[k - sa] = (unsigned char)lo;
                }
            }
            if (tid == 0) { meta[META * tile] = 1; meta[META * tile + 1] = padded; atomicAdd(stats + 0, 1); atomicMax(stats + 4, padded); }
        } else if (use_off16) {
            const int base = s_min;
            unsigned short *seg16 = reinterpret_cast<unsigned short *>(seg);
            for (int i = tid; i < 4; i += TILE_ROWS) dict[(size_t)tile * DICT_SLOTS + i] = base;
            for (int k = sa + tid; k < ea; k += TILE_ROWS)
                seg16[k - sa] = (k < nz0 || k >= nz1) ? (unsigned short)0 : (unsigned short)(ci[k] - base);
            if (tid == 0) { meta[META * tile] = 2; meta[META * tile + 1] = 4; atomicAdd(stats + 1, 1); }
        } else {
            if (tid == 0) { meta[META * tile] = 0; meta[META * tile + 1] = 0; atomicAdd(stats + 2, 1); }
        }
        __syncthreads();
    }
}

// value dictionary of a tile: the distinct bit patterns of its entries, ascending as unsigned integers
template <class T> __device__ __forceinline__ unsigned long long val_bits(T v);
template <> __device__ __forceinline__ unsigned long long val_bits<double>(double v) { return (unsigned long long)__double_as_longlong(v); }
template <> __device__ __forceinline__ unsigned long long val_bits<float>(float v) { return (unsigned long long)__float_as_uint(v); }
template <class T> __device__ __forceinline__ T val_from_bits(unsigned long long b);
template <> __device__ __forceinline__ double val_from_bits<double>(unsigned long long b) { return __longlong_as_double((long long)b); }
template <> __device__ __forceinline__ float val_from_bits<float>(unsigned long long b) { return __uint_as_float((unsigned)b); }

template <class MatT, int TILE_ROWS>
__global__ void __launch_bounds__(TILE_ROWS) valenc_build_kernel(const int *__restrict__ rp, const MatT *__restrict__ va, int row0, int n, int num_tiles, int tile_base,
                                                                 unsigned char *vcodes, MatT *vdict, int *meta, int *stats)
{
    constexpr unsigned long long EMPTY = ~0ull;
    __shared__ unsigned long long keys[HASH_SLOTS];
    __shared__ unsigned long long list[DICT_SLOTS], sorted[DICT_SLOTS];
    __shared__ int s_count, s_overflow;
    const int tid = threadIdx.x;
    for (int ltile = blockIdx.x; ltile < num_tiles; ltile += gridDim.x) {
        const int tile = tile_base + ltile;
        const int r0 = row0 + ltile * TILE_ROWS, r1 = min(r0 + TILE_ROWS, n);
        const int nz0 = rp[r0], nz1 = rp[r1];
        const int sa = nz0 & ~3, ea = (nz1 + 3) & ~3;
        for (int i = tid; i < HASH_SLOTS; i += TILE_ROWS) keys[i] = EMPTY;
        if (tid == 0) { s_count = 0; s_overflow = 0; }
        __syncthreads();
        for (int k = nz0 + tid; k < nz1; k += TILE_ROWS) {
            if (*(volatile int *)&s_overflow) break;
            const unsigned long long b = val_bits<MatT>(va[k]);
            if (b == EMPTY) { s_overflow = 1; break; }                   // the empty-slot sentinel cannot be a key
            unsigned h = ((unsigned)(b ^ (b >> 32)) * 2654435761u) >> 22;
            for (int probe = 0; probe < HASH_SLOTS; probe++) {
                if (*(volatile int *)&s_overflow) break;
                const unsigned long long old = atomicCAS(&keys[h], EMPTY, b);
                if (old == EMPTY) {
                    if (atomicAdd(&s_count, 1) >= DICT_SLOTS) s_overflow = 1;
                    break;
                }
                if (old == b) break;
                h = (h + 1) & (HASH_SLOTS - 1);
            }
        }
        __syncthreads();
        const bool use_dict = !s_overflow && s_count <= DICT_SLOTS && nz1 > nz0;
        const int count = s_count;
        __syncthreads();
        if (use_dict) {
            if (tid == 0) s_count = 0;
            __syncthreads();
            for (int i = tid; i < HASH_SLOTS; i += TILE_ROWS)
                if (keys[i] != EMPTY) list[atomicAdd(&s_count, 1)] = keys[i];
            __syncthreads();
            for (int i = tid; i < count; i += TILE_ROWS) {
                const unsigned long long v = list[i];
                int rank = 0;
                for (int j = 0; j < count; j++) rank += (list[j] < v);
                sorted[rank] = v;
            }
            __syncthreads();
            const int padded = (count + 3) & ~3;
            for (int i = tid; i < padded; i += TILE_ROWS) vdict[(size_t)tile * DICT_SLOTS + i] = val_from_bits<MatT>(sorted[min(i, count - 1)]);
            unsigned char *seg = vcodes + vcode_offset(sa, tile);
            for (int k = sa + tid; k < ea; k += TILE_ROWS) {
                int lo = 0;
                if (k >= nz0 && k < nz1) {
                    const unsigned long long b = val_bits<MatT>(va[k]);
                    int hi = count - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (sorted[mid] < b) lo = mid + 1;
                        else hi = mid;
                    }
                }
                seg[k - sa] = (unsigned char)lo;                          // alignment padding (entries of neighbouring tiles) gets code 0, never decoded
            }
            if (tid == 0) { meta[META * tile + 2] = 1; meta[META * tile + 3] = padded; atomicAdd(stats + 3, 1); atomicMax(stats + 5, padded); }
        } else {
            if (tid == 0) { meta[META * tile + 2] = 0; meta[META * tile + 3] = 0; }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// Pair dictionary (r02): a tile whose columns AND values are dictionary-coded and whose entries take <= 256 distinct
// (column code, value code) combinations gets ONE code byte per entry into a per-tile table of (column offset, value) pairs
// (7 pairs on a 7-point level, tens to ~150 on the first aggregation level).  The consumer then needs one byte load and one
// 16-byte look-up per entry instead of two byte loads and two look-ups: the coded kernels are bound by the SM's load/store
// wavefronts, not by HBM (ncu: 46 % DRAM, 64 % SM), so this is where their time is.  Built after the column and value codes;
// the pair flag lives beside them (pmeta), so value changes simply rebuild value codes and pairs.
// ---------------------------------------------------------------------------------------------
template <class MatT> struct EncPair;
template <> struct __align__(16) EncPair<double> { int off; int pad; double val; };
template <> struct __align__(8) EncPair<float> { int off; float val; };

template <class MatT, int TILE_ROWS>
__global__ void __launch_bounds__(TILE_ROWS) pairenc_build_kernel(const int *__restrict__ rp, int row0, int n, int num_tiles, int tile_base, const unsigned char *__restrict__ codes,
                                                                  const unsigned char *__restrict__ vcodes, const int *__restrict__ dict, const MatT *__restrict__ vdict,
                                                                  const int *__restrict__ meta, unsigned char *pcodes, EncPair<MatT> *pdict, int *pmeta, int *stats)
{
    __shared__ unsigned bitmap[2048];         // one bit per (column code, value code) combination
    __shared__ int wprefix[2048];
    __shared__ int part[TILE_ROWS];
    const int tid = threadIdx.x;
    for (int ltile = blockIdx.x; ltile < num_tiles; ltile += gridDim.x) {
        const int tile = tile_base + ltile;
        const int r0 = row0 + ltile * TILE_ROWS, r1 = min(r0 + TILE_ROWS, n);
        const int nz0 = rp[r0], nz1 = rp[r1];
        const int sa = nz0 & ~3, ea = (nz1 + 3) & ~3;
        if (meta[META * tile] != 1 || meta[META * tile + 2] != 1 || nz1 == nz0) {
            if (tid == 0) pmeta[tile] = 0;
            continue;                                                     // uniform over the CTA
        }
        const unsigned char *cseg = codes + code_offset(sa, tile), *vseg = vcodes + vcode_offset(sa, tile);
        for (int i = tid; i < 2048; i += TILE_ROWS) bitmap[i] = 0u;
        __syncthreads();
        for (int k = nz0 + tid; k < nz1; k += TILE_ROWS) {
            const unsigned key = ((unsigned)cseg[k - sa] << 8) | (unsigned)vseg[k - sa];
            atomicOr(&bitmap[key >> 5], 1u << (key & 31));
        }
        __syncthreads();
        // exclusive prefix of the per-word popcounts: each thread owns 2048 / TILE_ROWS consecutive words
        constexpr int WPT = 2048 / TILE_ROWS;
        int local = 0;
        for (int w = 0; w < WPT; w++) local += __popc(bitmap[tid * WPT + w]);
        part[tid] = local;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int t = 0; t < TILE_ROWS; t++) { const int v = part[t]; part[t] = run; run += v; }
            wprefix[0] = run;                                             // total, parked until everybody has read its partial
        }
        __syncthreads();
        const int count = wprefix[0];
        __syncthreads();
        int run = part[tid];
        for (int w = 0; w < WPT; w++) { wprefix[tid * WPT + w] = run; run += __popc(bitmap[tid * WPT + w]); }
        __syncthreads();
        if (count <= DICT_SLOTS) {
            unsigned char *pseg = pcodes + vcode_offset(sa, tile);
            for (int k = sa + tid; k < ea; k += TILE_ROWS) {
                int code = 0;
                if (k >= nz0 && k < nz1) {
                    const unsigned key = ((unsigned)cseg[k - sa] << 8) | (unsigned)vseg[k - sa];
                    code = wprefix[key >> 5] + __popc(bitmap[key >> 5] & ((1u << (key & 31)) - 1u));
                }
                pseg[k - sa] = (unsigned char)code;
            }
            const int *dt = dict + (size_t)tile * DICT_SLOTS;
            const MatT *vt = vdict + (size_t)tile * DICT_SLOTS;
            EncPair<MatT> *pt = pdict + (size_t)tile * DICT_SLOTS;
            for (int w = tid; w < 2048; w += TILE_ROWS) {
                unsigned bits = bitmap[w];
                int rank = wprefix[w];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const unsigned key = ((unsigned)w << 5) | (unsigned)b;
                    EncPair<MatT> pr;
                    memset(&pr, 0, sizeof(pr));
                    pr.off = dt[key >> 8];
                    pr.val = vt[key & 255u];
                    pt[rank++] = pr;
                }
            }
            const int padded = (count + 3) & ~3;
            __syncthreads();
            for (int i = count + tid; i < padded; i += TILE_ROWS) pt[i] = pt[count - 1];
            if (tid == 0) { pmeta[tile] = padded; atomicAdd(stats + 6, 1); atomicMax(stats + 7, padded); }
        } else if (tid == 0) pmeta[tile] = 0;
        __syncthreads();
    }
}


// x rows worth prefetching for a tile (see l2_prefetch_span): on a dictionary-coded tile the k-th neighbours of consecutive rows are
// consecutive in x, and the neighbour at the largest offset is the first reader of its x rows -- a DRAM miss on the consumers' critical
// path unless the producer asks for those rows ahead of time.  Tiles with other column codings (no such structure) get 0.
__global__ void xahead_kernel(const int *__restrict__ meta, const int *__restrict__ dict, int num_tiles, int *__restrict__ xahead)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    int best = 0;
    if (meta[META * t] == 1) {
        const int dlen = meta[META * t + 1];
        for (int i = 0; i < dlen; i++) best = max(best, dict[(size_t)t * DICT_SLOTS + i]);
    }
    xahead[t] = best;
}

// ---------------------------------------------------------------------------------------------
// Row patterns (r02): on a stencil level most rows of a tile are the SAME sequence of (offset, value) pairs -- all interior rows of a
// 7-point tile are one pattern.  A pair-coded tile whose rows have <= 7 entries and take <= 64 distinct sequences of pair codes moves
// ONE byte per ROW: the row's pattern id into a per-tile table of patterns (7 pair-code bytes + length).  The consumer loads its pattern
// with one 16-byte look-up and peels the pair codes out of a register -- no per-entry code loads, no row_ptr slice.  HBM per row:
// 1 byte + the vectors.  Built after the pair tables (rowpat_build_kernel); value changes rebuild pairs and patterns.
// ---------------------------------------------------------------------------------------------
constexpr int RP_SLOTS = 64;             // patterns per tile
constexpr int RP_MAX_LEN = 7;            // entries per row: 7 code bytes + the length in the top byte of the 64-bit pattern key
struct __align__(16) RowPattern { unsigned long long codes; int len; int pad; };

template <int TILE_ROWS>
__global__ void __launch_bounds__(TILE_ROWS) rowpat_build_kernel(const int *__restrict__ rp, int row0, int n, int num_tiles, int tile_base, const unsigned char *__restrict__ pcodes,
                                                                 const int *__restrict__ pmeta, unsigned char *rowcodes, RowPattern *rpat, int *rmeta, int *stats)
{
    constexpr int HS = 256;
    constexpr unsigned long long EMPTY = ~0ull;
    __shared__ unsigned long long hkey[HS];           // open addressing; a key is its own length tag (top byte), so one 64-bit CAS claims or matches a slot
    __shared__ unsigned long long lkey[RP_SLOTS];
    __shared__ int srank[RP_SLOTS];
    __shared__ int s_count, s_bad;
    const int tid = threadIdx.x;
    for (int ltile = blockIdx.x; ltile < num_tiles; ltile += gridDim.x) {
        const int tile = tile_base + ltile;
        const int r0 = row0 + ltile * TILE_ROWS, r1 = min(r0 + TILE_ROWS, n);
        if (pmeta[tile] <= 0) {
            if (tid == 0) rmeta[tile] = 0;
            continue;
        }
        const int nz0 = rp[r0], sa = nz0 & ~3;
        const unsigned char *pseg = pcodes + vcode_offset(sa, tile);
        for (int i = tid; i < HS; i += TILE_ROWS) hkey[i] = EMPTY;
        if (tid == 0) { s_count = 0; s_bad = 0; }
        __syncthreads();
        const int row = r0 + tid;
        unsigned long long key = 0ull;                 // rows past the end of the segment: the empty pattern
        if (row < r1) {
            const int k0 = rp[row], len = rp[row + 1] - k0;
            if (len > RP_MAX_LEN) s_bad = 1;
            else {
                for (int j = 0; j < len; j++) key |= (unsigned long long)pseg[k0 + j - sa] << (8 * j);
                key |= (unsigned long long)len << 56;
            }
        }
        __syncthreads();
        if (!s_bad) {
            unsigned h = (unsigned)(((key ^ (key >> 29)) * 0xff51afd7ed558ccdull) >> 56) & (HS - 1);
            for (int probe = 0; probe < HS; probe++) {
                const unsigned long long old = atomicCAS(&hkey[h], EMPTY, key);
                if (old == EMPTY) { if (atomicAdd(&s_count, 1) >= RP_SLOTS) s_bad = 1; break; }
                if (old == key) break;
                h = (h + 1) & (HS - 1);
            }
        }
        __syncthreads();
        const int count = s_count;
        const bool ok = !s_bad && count <= RP_SLOTS;
        __syncthreads();
        if (ok) {
            // compact + rank (ascending key, i.e. by length then codes: the table, hence the row codes, do not depend on insertion order)
            if (tid == 0) s_count = 0;
            __syncthreads();
            for (int i = tid; i < HS; i += TILE_ROWS)
                if (hkey[i] != EMPTY) lkey[atomicAdd(&s_count, 1)] = hkey[i];
            __syncthreads();
            for (int i = tid; i < count; i += TILE_ROWS) {
                int rank = 0;
                for (int j = 0; j < count; j++) rank += lkey[j] < lkey[i];
                srank[i] = rank;
                RowPattern pt;
                pt.codes = lkey[i];
                pt.len = (int)(lkey[i] >> 56);
                pt.pad = 0;
                rpat[(size_t)tile * RP_SLOTS + rank] = pt;
            }
            __syncthreads();
            int code = 0;
            for (int i = 0; i < count; i++)
                if (lkey[i] == key) { code = srank[i]; break; }
            rowcodes[(size_t)tile * TILE_ROWS + tid] = (unsigned char)code;
            if (tid == 0) { rmeta[tile] = count; atomicAdd(stats + 0, 1); atomicMax(stats + 1, count); }
        } else if (tid == 0) rmeta[tile] = 0;
        __syncthreads();
    }
}

__device__ __forceinline__ void ld_pair(const EncPair<double> *p, int &off, double &val);
__device__ __forceinline__ void ld_pair(const EncPair<float> *p, int &off, float &val);

// one row of a row-pattern tile: the pair codes come out of a register
template <class MatT, class VecT, int N>
__device__ __forceinline__ VecT row_pat_fixed(const unsigned long long codes, const EncPair<MatT> *__restrict__ pdict, const VecT *__restrict__ x, const int row)
{
    int off[N];
    MatT val[N];
    VecT xv[N];
    VecT sum = 0;
#pragma unroll
    for (int j = 0; j < N; j++) ld_pair(pdict + (unsigned)((codes >> (8 * j)) & 0xffull), off[j], val[j]);
#pragma unroll
    for (int j = 0; j < N; j++) xv[j] = __ldg(x + (row + off[j]));
#pragma unroll
    for (int j = 0; j < N; j++) sum = fma((VecT)val[j], xv[j], sum);
    return sum;
}
template <class MatT, class VecT>
__device__ __forceinline__ VecT row_dot_pattern(const RowPattern *__restrict__ rpat, const unsigned char code, const EncPair<MatT> *__restrict__ pdict, const int row,
                                                const VecT *__restrict__ x)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(rpat + code);
    const unsigned long long codes = ((unsigned long long)v.y << 32) | v.x;
    switch ((int)v.z) {
    case 7: return row_pat_fixed<MatT, VecT, 7>(codes, pdict, x, row);
    case 6: return row_pat_fixed<MatT, VecT, 6>(codes, pdict, x, row);
    case 5: return row_pat_fixed<MatT, VecT, 5>(codes, pdict, x, row);
    case 4: return row_pat_fixed<MatT, VecT, 4>(codes, pdict, x, row);
    case 3: return row_pat_fixed<MatT, VecT, 3>(codes, pdict, x, row);
    case 2: return row_pat_fixed<MatT, VecT, 2>(codes, pdict, x, row);
    case 1: return row_pat_fixed<MatT, VecT, 1>(codes, pdict, x, row);
    default: return (VecT)0;
    }
}

// ---------------------------------------------------------------------------------------------
// The encoded tile kernel.  blockDim.x = TILE_ROWS + 32 (last warp = producer).
// Stage layout, sized PER LEVEL from what its tiles actually use (a level whose tiles are all coded stages 2 bytes per entry, so
// many more CTAs fit an SM -- the consumers are latency-bound, occupancy is what buys bandwidth here):
//   value stream  cap * val_w bytes (val_w = 1 when every tile has a value dictionary, else sizeof(MatT))
//   column stream cap * col_w bytes (col_w = 1: all dict8; 2: dict8 / off16; 4: some tile keeps raw columns)
//   dict[dict_cap] ints | vdict[vdict_cap] MatT | rp[TILE_ROWS+4]
// ---------------------------------------------------------------------------------------------
struct EncArgs {
    const unsigned char *codes;
    const int *dict;
    const int *meta;
    const unsigned char *vcodes;
    const void *vdict;
    const unsigned char *pcodes;   // pair codes (1 byte per entry) and pair tables of the tiles with pmeta[tile] > 0
    const void *pdict;
    const int *pmeta;
    const unsigned char *rowcodes; // row-pattern ids (1 byte per row) and pattern tables of the tiles with rmeta[tile] > 0
    const RowPattern *rpat;
    const int *rmeta;
    const int *xahead;             // per tile: the largest column offset of a dictionary-coded (stencil-like) tile, 0 = none: x rows to prefetch into L2
    int val_w, col_w, dict_cap, vdict_cap;   // vdict_cap counts BYTES of the value / pair dictionary region
    int tile_base;      // global index of the segment's first tile (meta / dictionaries / code segments are numbered over all segments)
    int x_len;          // entries of x a column may address (rows + halo)
};

// one row of an encoded tile (separate column / value codes): rows are dispatched on their exact length to straight-line code (N code loads,
// N dictionary look-ups, N gathers of x in flight, N FMAs in storage order, no predicates) -- the coded kernels are bound by instruction issue
template <class MatT, class VecT, int ENC, int VENC, int N>
__device__ __forceinline__ VecT row_enc_fixed(const unsigned char *__restrict__ vstream, const unsigned char *__restrict__ cstream, const int *__restrict__ dict,
                                              const MatT *__restrict__ vdict, const int k, const int base, const VecT *__restrict__ x, VecT sum)
{
    int c[N];
    VecT xv[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (ENC == 1) c[j] = base + dict[cstream[k + j]];
        else if (ENC == 2) c[j] = base + (int)reinterpret_cast<const unsigned short *>(cstream)[k + j];
        else c[j] = reinterpret_cast<const int *>(cstream)[k + j];
    }
#pragma unroll
    for (int j = 0; j < N; j++) xv[j] = __ldg(x + c[j]);
#pragma unroll
    for (int j = 0; j < N; j++) {
        const MatT v = VENC ? vdict[vstream[k + j]] : reinterpret_cast<const MatT *>(vstream)[k + j];
        sum = fma((VecT)v, xv[j], sum);
    }
    return sum;
}

template <class MatT, class VecT, int ENC, int VENC>
__device__ __forceinline__ VecT row_dot_enc(const unsigned char *__restrict__ vstream, const unsigned char *__restrict__ cstream, const int *__restrict__ dict,
                                            const MatT *__restrict__ vdict, int k, const int kend, const int row, const VecT *__restrict__ x)
{
    const int base = (ENC == 2) ? dict[0] : row;
    int len = kend - k;
    VecT sum = 0;
    for (; len >= 8; len -= 8, k += 8) sum = row_enc_fixed<MatT, VecT, ENC, VENC, 8>(vstream, cstream, dict, vdict, k, base, x, sum);
    switch (len) {
    case 7: sum = row_enc_fixed<MatT, VecT, ENC, VENC, 7>(vstream, cstream, dict, vdict, k, base, x, sum); break;
    case 6: sum = row_enc_fixed<MatT, VecT, ENC, VENC, 6>(vstream, cstream, dict, vdict, k, base, x, sum); break;
    case 5: sum = row_enc_fixed<MatT, VecT, ENC, VENC, 5>(vstream, cstream, dict, vdict, k, base, x, sum); break;
    case 4: sum = row_enc_fixed<MatT, VecT, ENC, VENC, 4>(vstream, cstream, dict, vdict, k, base, x, sum); break;
    case 3: sum = row_enc_fixed<MatT, VecT, ENC, VENC, 3>(vstream, cstream, dict, vdict, k, base, x, sum); break;
    case 2: sum = row_enc_fixed<MatT, VecT, ENC, VENC, 2>(vstream, cstream, dict, vdict, k, base, x, sum); break;
    case 1: sum = row_enc_fixed<MatT, VecT, ENC, VENC, 1>(vstream, cstream, dict, vdict, k, base, x, sum); break;
    default: break;
    }
    return sum;
}

// one table entry with ONE shared-memory load (LDS.128 / LDS.64) instead of one per member
__device__ __forceinline__ void ld_pair(const EncPair<double> *p, int &off, double &val)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    off = (int)v.x;
    val = __hiloint2double((int)v.w, (int)v.z);
}
__device__ __forceinline__ void ld_pair(const EncPair<float> *p, int &off, float &val)
{
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    off = (int)v.x;
    val = __uint_as_float(v.y);
}

template <class MatT, class VecT, int N>
__device__ __forceinline__ VecT row_pair_fixed(const unsigned char *__restrict__ cs, const EncPair<MatT> *__restrict__ pdict, const VecT *__restrict__ x, const int row, VecT sum)
{
    int off[N];
    MatT val[N];
    VecT xv[N];
#pragma unroll
    for (int j = 0; j < N; j++) ld_pair(pdict + cs[j], off[j], val[j]);
#pragma unroll
    for (int j = 0; j < N; j++) xv[j] = __ldg(x + (row + off[j]));      // 32-bit column, then one scaled 64-bit add
#pragma unroll
    for (int j = 0; j < N; j++) sum = fma((VecT)val[j], xv[j], sum);
    return sum;
}

template <class MatT, class VecT>
__device__ __forceinline__ VecT row_dot_pair(const unsigned char *__restrict__ cstream, const EncPair<MatT> *__restrict__ pdict, int k, const int kend, const int row,
                                             const VecT *__restrict__ x)
{
    const unsigned char *cs = cstream + k;
    int len = kend - k;
    VecT sum = 0;
    for (; len >= 8; len -= 8, cs += 8) sum = row_pair_fixed<MatT, VecT, 8>(cs, pdict, x, row, sum);
    switch (len) {
    case 7: sum = row_pair_fixed<MatT, VecT, 7>(cs, pdict, x, row, sum); break;
    case 6: sum = row_pair_fixed<MatT, VecT, 6>(cs, pdict, x, row, sum); break;
    case 5: sum = row_pair_fixed<MatT, VecT, 5>(cs, pdict, x, row, sum); break;
    case 4: sum = row_pair_fixed<MatT, VecT, 4>(cs, pdict, x, row, sum); break;
    case 3: sum = row_pair_fixed<MatT, VecT, 3>(cs, pdict, x, row, sum); break;
    case 2: sum = row_pair_fixed<MatT, VecT, 2>(cs, pdict, x, row, sum); break;
    case 1: sum = row_pair_fixed<MatT, VecT, 1>(cs, pdict, x, row, sum); break;
    default: break;
    }
    return sum;
}

template <class MatT, class VecT, int TILE_ROWS, int EPI>
__global__ void __launch_bounds__(TILE_ROWS + PRODUCER_THREADS, (TILE_ROWS == 256 ? 5 : 9)) csr_tile_enc_kernel(const TileArgs<MatT, VecT> a, const EncArgs e)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + MAX_STAGES;
    double *smem_red = reinterpret_cast<double *>(smem_raw + 2 * MAX_STAGES * sizeof(uint64_t));
    unsigned char *stage_base = smem_raw + 512;
    const size_t vals_bytes = align16((size_t)a.cap * e.val_w);
    const size_t cols_bytes = align16((size_t)a.cap * e.col_w);
    const size_t dict_bytes = (size_t)e.dict_cap * sizeof(int);
    const size_t vdict_bytes = align16((size_t)e.vdict_cap);
    const size_t rp_bytes = (size_t)(TILE_ROWS + 4) * sizeof(int);
    const size_t hdr_off = vals_bytes + cols_bytes + dict_bytes + vdict_bytes + rp_bytes;      // 16-byte tile header: what the producer read from meta / pmeta
    const size_t stage_bytes = hdr_off + 16;
    constexpr int CONSUMER_WARPS = TILE_ROWS / 32;
    constexpr bool HAS_RED = (EPI == EPI_SPMV_DOT || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < a.stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc = 0.0;
    const int my_tiles = (a.num_tiles > (int)blockIdx.x) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid >= TILE_ROWS) {
        // ------------------------------- producer warp -------------------------------
        if (tid == TILE_ROWS) {
            int s = 0;             // stage and mbarrier phase advance by counting: no runtime division per tile
            unsigned ph = 0u;
            for (int it = 0; it < my_tiles; it++, s = (s + 1 == a.stages) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
                const int tile = blockIdx.x + it * gridDim.x;
                if (it >= a.stages) mbar_wait(&empty[s], ph ^ 1u);
                const int r0 = a.row0 + tile * TILE_ROWS;
                const int r1 = min(r0 + TILE_ROWS, a.n);
                const int nz0 = __ldg(a.row_ptr + r0), nz1 = __ldg(a.row_ptr + r1);
                const int sa = nz0 & ~3, ea = (nz1 + 3) & ~3;
                const int gtile = e.tile_base + tile;
                const int4 m = __ldg(reinterpret_cast<const int4 *>(e.meta) + gtile);
                const int enc = m.x, dlen = m.y, venc = m.z, vdlen = m.w;
                const int pdlen = __ldg(e.pmeta + gtile);
                unsigned char *st = stage_base + (size_t)s * stage_bytes;
                const unsigned rp_copy = (unsigned)(((r1 - r0 + 1 + 3) & ~3) * sizeof(int));
                const unsigned cnt = (unsigned)(ea - sa);
                const int rplen = pdlen > 0 ? __ldg(e.rmeta + gtile) : 0;
                *reinterpret_cast<int4 *>(st + hdr_off) = make_int4(enc, venc, pdlen, rplen);   // released to the consumers by the arrive below
                if (a.l2pf) {     // `stages` tiles ahead of the consumers: their per-row vector loads, and the x rows only this tile's furthest neighbour has reached
                    if (EPI == EPI_RESID || EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2 || EPI == EPI_JACOBI_L1 || EPI == EPI_ADD)
                        l2_prefetch_span(a.b + r0, r1 - r0);
                    if (EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_JACOBI_L1) l2_prefetch_span(a.d + r0, r1 - r0);
                    const int ahead = __ldg(e.xahead + gtile);
                    if (ahead > 0) l2_prefetch_span(a.x + (size_t)r0 + ahead, min(r1 - r0, e.x_len - r0 - ahead));
                }
                if (rplen > 0) {      // row-pattern tile: one byte per ROW into the column stream, patterns into the dictionary region, pairs beside them
                    const unsigned rc_copy = (unsigned)TILE_ROWS, rp2_copy = (unsigned)rplen * (unsigned)sizeof(RowPattern),
                                   pd_copy = (unsigned)align16((size_t)pdlen * sizeof(EncPair<MatT>));
                    mbar_expect_tx(&full[s], rc_copy + rp2_copy + pd_copy);
                    tma_bulk_g2s(st + vals_bytes, e.rowcodes + (size_t)gtile * TILE_ROWS, rc_copy, &full[s]);
                    tma_bulk_g2s(st + vals_bytes + cols_bytes, e.rpat + (size_t)gtile * RP_SLOTS, rp2_copy, &full[s]);
                    tma_bulk_g2s(st + vals_bytes + cols_bytes + dict_bytes, reinterpret_cast<const EncPair<MatT> *>(e.pdict) + (size_t)gtile * DICT_SLOTS, pd_copy, &full[s]);
                    continue;
                }
                if (pdlen > 0) {      // pair-coded tile: code bytes into the column stream, the pair table into the value-dictionary region
                    const unsigned pc_copy = (unsigned)align16(cnt), pd_copy = (unsigned)align16((size_t)pdlen * sizeof(EncPair<MatT>));
                    mbar_expect_tx(&full[s], rp_copy + pc_copy + pd_copy);
                    tma_bulk_g2s(st + vals_bytes + cols_bytes + dict_bytes + vdict_bytes, a.row_ptr + r0, rp_copy, &full[s]);
                    tma_bulk_g2s(st + vals_bytes, e.pcodes + vcode_offset(sa, gtile), pc_copy, &full[s]);
                    tma_bulk_g2s(st + vals_bytes + cols_bytes + dict_bytes, reinterpret_cast<const EncPair<MatT> *>(e.pdict) + (size_t)gtile * DICT_SLOTS, pd_copy, &full[s]);
                    continue;
                }
                unsigned col_copy = 0, val_copy = 0;
                if (cnt) col_copy = enc == 1 ? (unsigned)align16(cnt) : enc == 2 ? (unsigned)align16((size_t)cnt * 2) : cnt * (unsigned)sizeof(int);
                if (cnt) val_copy = venc == 1 ? (unsigned)align16(cnt) : cnt * (unsigned)sizeof(MatT);
                const unsigned dict_copy = (unsigned)dlen * (unsigned)sizeof(int);
                const unsigned vdict_copy = (unsigned)align16((size_t)vdlen * sizeof(MatT));
                mbar_expect_tx(&full[s], rp_copy + val_copy + col_copy + dict_copy + vdict_copy);
                tma_bulk_g2s(st + vals_bytes + cols_bytes + dict_bytes + vdict_bytes, a.row_ptr + r0, rp_copy, &full[s]);
                if (cnt) {
                    if (venc == 0) tma_bulk_g2s(st, a.val + sa, val_copy, &full[s]);
                    else tma_bulk_g2s(st, e.vcodes + vcode_offset(sa, gtile), val_copy, &full[s]);
                    if (enc == 0) tma_bulk_g2s(st + vals_bytes, a.col + sa, col_copy, &full[s]);
                    else tma_bulk_g2s(st + vals_bytes, e.codes + code_offset(sa, gtile), col_copy, &full[s]);
                }
                if (dict_copy) tma_bulk_g2s(st + vals_bytes + cols_bytes, e.dict + (size_t)gtile * DICT_SLOTS, dict_copy, &full[s]);
                if (vdict_copy)
                    tma_bulk_g2s(st + vals_bytes + cols_bytes + dict_bytes, reinterpret_cast<const MatT *>(e.vdict) + (size_t)gtile * DICT_SLOTS, vdict_copy, &full[s]);
            }
        }
    } else {
        // ------------------------------- consumers: one row per thread -------------------------------
        int s = 0;
        unsigned ph = 0u;
        for (int it = 0; it < my_tiles; it++, s = (s + 1 == a.stages) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int lrow = a.perm ? (int)__ldg(a.perm + (size_t)(e.tile_base + tile) * TILE_ROWS + tid) : tid;      // length-sorted rows (k_spmv.cu)
            const int row = a.row0 + tile * TILE_ROWS + lrow;
            const bool active = row < a.n;
            VecT bi = 0, xi = 0;
            MatT di = 1;
            if (active) {
                if (EPI == EPI_RESID || EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2 || EPI == EPI_JACOBI_L1 || EPI == EPI_ADD)
                    bi = __ldg(a.b + row);
                if (EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_JACOBI_L1) {
                    di = __ldg(a.d + row);
                    xi = __ldg(a.x + row);
                }
                if (EPI == EPI_SPMV_DOT) xi = __ldg(a.x + row);
            }
            const unsigned char *st = stage_base + (size_t)s * stage_bytes;
            const unsigned char *vstream = st;
            const unsigned char *cstream = st + vals_bytes;
            const int *dict = reinterpret_cast<const int *>(st + vals_bytes + cols_bytes);
            const MatT *vdict = reinterpret_cast<const MatT *>(st + vals_bytes + cols_bytes + dict_bytes);
            const int *rp = reinterpret_cast<const int *>(st + vals_bytes + cols_bytes + dict_bytes + vdict_bytes);
            mbar_wait(&full[s], ph);
            const int4 hdr = *reinterpret_cast<const int4 *>(st + hdr_off);      // the tile's encoding, uniform over the CTA
            const int enc = hdr.x, venc = hdr.y, pd = hdr.z;
            if (hdr.w > 0) {
                // row-pattern tile: no row_ptr slice, no per-entry codes
                if (active) {
                    const VecT sum = row_dot_pattern<MatT, VecT>(reinterpret_cast<const RowPattern *>(dict), cstream[lrow], reinterpret_cast<const EncPair<MatT> *>(vdict), row, a.x);
                    acc += tile_epilogue<MatT, VecT, EPI>(a, row, sum, bi, di, xi);
                }
            } else
            if (active) {
                const int sa = rp[0] & ~3;
                const int k = rp[lrow] - sa, kend = rp[lrow + 1] - sa;
                // the tile's encoding is uniform over the CTA: the switch does not diverge
                VecT sum;
                if (pd > 0) sum = row_dot_pair<MatT, VecT>(cstream, reinterpret_cast<const EncPair<MatT> *>(vdict), k, kend, row, a.x);
                else switch (enc * 2 + venc) {
                case 3: sum = row_dot_enc<MatT, VecT, 1, 1>(vstream, cstream, dict, vdict, k, kend, row, a.x); break;
                case 2: sum = row_dot_enc<MatT, VecT, 1, 0>(vstream, cstream, dict, vdict, k, kend, row, a.x); break;
                case 5: sum = row_dot_enc<MatT, VecT, 2, 1>(vstream, cstream, dict, vdict, k, kend, row, a.x); break;
                case 4: sum = row_dot_enc<MatT, VecT, 2, 0>(vstream, cstream, dict, vdict, k, kend, row, a.x); break;
                case 1: sum = row_dot_enc<MatT, VecT, 0, 1>(vstream, cstream, dict, vdict, k, kend, row, a.x); break;
                default: sum = row_dot_enc<MatT, VecT, 0, 0>(vstream, cstream, dict, vdict, k, kend, row, a.x); break;
                }
                acc += tile_epilogue<MatT, VecT, EPI>(a, row, sum, bi, di, xi);
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        }
    }
    if (HAS_RED) block_reduce_finish(acc, smem_red, a.red, a.fin_op, a.fin_slot, a.mirror);
}

template <class MatT, class VecT, int TILE_ROWS, int EPI>
void launch_enc(const Matrix &A, const TileArgs<MatT, VecT> &ta, const EncArgs &ea, int grid, cudaStream_t s)
{
    const size_t smem = A.colenc.smem_bytes;
    auto k = csr_tile_enc_kernel<MatT, VecT, TILE_ROWS, EPI>;
    smem_opt_in(reinterpret_cast<const void *>(k), smem);      // exactly what this kernel needs, once per size and device
    // the persistent grid must not exceed what is RESIDENT for this instantiation (registers differ per epilogue): a CTA that waits for a
    // slot runs its tiles after everybody else's (r02: 52 registers -> 4 resident of 5 launched per SM cost 50 % on the Jacobi sweep)
    static std::map<size_t, int> occ_by_smem;
    auto it = occ_by_smem.find(smem);
    if (it == occ_by_smem.end()) {
        int occ = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, TILE_ROWS + PRODUCER_THREADS, smem) != cudaSuccess) { cudaGetLastError(); occ = 1; }
        it = occ_by_smem.emplace(smem, std::max(occ, 1)).first;
    }
    const int sms = A.rsc ? A.rsc->num_sms : 148;
    grid = std::max(1, std::min(grid, sms * it->second));
    k<<<grid, TILE_ROWS + PRODUCER_THREADS, smem, s>>>(ta, ea);
}

template <class MatT, class VecT, int EPI> void launch_enc_epi(const Matrix &A, const TileArgs<MatT, VecT> &ta, const EncArgs &ea, cudaStream_t s)
{
    const int sms = A.rsc ? A.rsc->num_sms : 148;
    const int grid = std::max(1, std::min(sms * A.colenc.ctas_per_sm, ta.num_tiles));
    if (A.plan.tile_rows == 256) launch_enc<MatT, VecT, 256, EPI>(A, ta, ea, grid, s);
    else launch_enc<MatT, VecT, 128, EPI>(A, ta, ea, grid, s);
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

}  // namespace

// AMGXB_COLENC: bit 0 = compressed columns, bit 1 = value dictionaries (0 / unset: the plain kernels)
static int colenc_flags()
{
    static const int f = getenv("AMGXB_COLENC") ? atoi(getenv("AMGXB_COLENC")) & 3 : 3;     // default: coded columns AND values (r02: 256 -> 341 it/s on config 2)
    return f;
}
bool colenc_requested() { return colenc_flags() != 0; }

// Row segments the tile kernels run over: the whole matrix, or (row-partitioned matrices) the interior rows [0, split) and the rows
// [split, n) that read halo columns.  Tiles are numbered over the segments in this order.
struct EncSeg { int row0, row1, tiles, base; };
static int enc_segments(const Matrix &A, EncSeg seg[2])
{
    const int T = A.plan.tile_rows;
    if (A.plan.split > 0 && A.plan.split < A.n) {
        seg[0] = EncSeg{0, A.plan.split, ceil_div(A.plan.split, T), 0};
        seg[1] = EncSeg{A.plan.split, A.n, ceil_div(A.n - A.plan.split, T), seg[0].tiles};
        return 2;
    }
    seg[0] = EncSeg{0, A.n, ceil_div(A.n, T), 0};
    return 1;
}

static void build_value_codes(Matrix &A, cudaStream_t s)
{
    ColEnc &E = A.colenc;
    const int T = A.plan.tile_rows, nt = A.plan.num_tiles;
    DevBuf<int> stats;
    stats.resize(8);
    stats.zero(s);
    EncSeg seg[2];
    const int nseg = enc_segments(A, seg);
    for (int g = 0; g < nseg; g++) {
        const int grid = std::max(1, std::min(seg[g].tiles, B200_SMS * 8));
        if (A.mat_prec == Prec::F64) {
            if (T == 256) valenc_build_kernel<double, 256><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), A.values.as<double>(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.vcodes.ptr(), (double *)E.vdict.ptr(), E.meta.ptr(), stats.ptr());
            else valenc_build_kernel<double, 128><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.values.as<double>(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.vcodes.ptr(), (double *)E.vdict.ptr(), E.meta.ptr(), stats.ptr());
        } else {
            if (T == 256) valenc_build_kernel<float, 256><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), A.values.as<float>(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.vcodes.ptr(), (float *)E.vdict.ptr(), E.meta.ptr(), stats.ptr());
            else valenc_build_kernel<float, 128><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.values.as<float>(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.vcodes.ptr(), (float *)E.vdict.ptr(), E.meta.ptr(), stats.ptr());
        }
        count_launch();
    }
    (void)nt;
    AMGXB_LAUNCH_CHECK();
    const std::vector<int> h = stats.to_host(s);
    E.tiles_val8 = h[3];
    E.max_vdlen = h[5];
}

// pair tables of the tiles whose columns and values are both dictionary-coded (pairenc_build_kernel); off with AMGXB_ENC_PAIRS=0
static void build_pair_codes(Matrix &A, cudaStream_t s)
{
    ColEnc &E = A.colenc;
    const int T = A.plan.tile_rows, nt = E.num_tiles;
    static const int pairs_on = getenv("AMGXB_ENC_PAIRS") ? atoi(getenv("AMGXB_ENC_PAIRS")) : 1;
    E.tiles_pair = 0;
    E.max_pdlen = 0;
    E.pmeta.resize((size_t)std::max(nt, 1));
    E.pmeta.zero(s);
    if (!pairs_on || !E.values_encoded || E.tiles_dict8 == 0 || E.tiles_val8 == 0) { if (E.pcodes.size() == 0) { E.pcodes.resize(64); E.pdict.resize(64); } return; }
    const size_t msz = prec_size(A.mat_prec), psz = msz == 8 ? 16 : 8;
    E.pcodes.resize(align16((size_t)A.nnz + 8) + (size_t)32 * nt + 64);
    E.pdict.resize((size_t)nt * DICT_SLOTS * psz);
    DevBuf<int> stats;
    stats.resize(8);
    stats.zero(s);
    EncSeg seg[2];
    const int nseg = enc_segments(A, seg);
    for (int g = 0; g < nseg; g++) {
        const int grid = std::max(1, std::min(seg[g].tiles, B200_SMS * 8));
        if (A.mat_prec == Prec::F64) {
            if (T == 256) pairenc_build_kernel<double, 256><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.codes.ptr(), E.vcodes.ptr(), E.dict.ptr(), (const double *)E.vdict.ptr(), E.meta.ptr(), E.pcodes.ptr(), (EncPair<double> *)E.pdict.ptr(), E.pmeta.ptr(), stats.ptr());
            else pairenc_build_kernel<double, 128><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.codes.ptr(), E.vcodes.ptr(), E.dict.ptr(), (const double *)E.vdict.ptr(), E.meta.ptr(), E.pcodes.ptr(), (EncPair<double> *)E.pdict.ptr(), E.pmeta.ptr(), stats.ptr());
        } else {
            if (T == 256) pairenc_build_kernel<float, 256><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.codes.ptr(), E.vcodes.ptr(), E.dict.ptr(), (const float *)E.vdict.ptr(), E.meta.ptr(), E.pcodes.ptr(), (EncPair<float> *)E.pdict.ptr(), E.pmeta.ptr(), stats.ptr());
            else pairenc_build_kernel<float, 128><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.codes.ptr(), E.vcodes.ptr(), E.dict.ptr(), (const float *)E.vdict.ptr(), E.meta.ptr(), E.pcodes.ptr(), (EncPair<float> *)E.pdict.ptr(), E.pmeta.ptr(), stats.ptr());
        }
        count_launch();
    }
    AMGXB_LAUNCH_CHECK();
    const std::vector<int> h = stats.to_host(s);
    E.tiles_pair = h[6];
    E.max_pdlen = h[7];
}

// row patterns of the pair-coded tiles (rowpat_build_kernel); off with AMGXB_ENC_ROWPAT=0
static void build_row_patterns(Matrix &A, cudaStream_t s)
{
    ColEnc &E = A.colenc;
    const int T = A.plan.tile_rows, nt = E.num_tiles;
    static const int on = getenv("AMGXB_ENC_ROWPAT") ? atoi(getenv("AMGXB_ENC_ROWPAT")) : 1;
    E.tiles_rowpat = 0;
    E.max_rplen = 0;
    E.rmeta.resize((size_t)std::max(nt, 1));
    E.rmeta.zero(s);
    // the T row codes travel in the tile's column-stream region: it must hold them (cap * col_w >= cap >= T)
    if (!on || E.tiles_pair == 0 || A.plan.max_tile_nnz < T) { if (E.rowcodes.size() == 0) { E.rowcodes.resize(64); E.rpat.resize(64); } return; }
    E.rowcodes.resize((size_t)nt * T + 64);
    E.rpat.resize((size_t)nt * RP_SLOTS * sizeof(RowPattern));
    DevBuf<int> stats;
    stats.resize(8);
    stats.zero(s);
    EncSeg seg[2];
    const int nseg = enc_segments(A, seg);
    for (int g = 0; g < nseg; g++) {
        const int grid = std::max(1, std::min(seg[g].tiles, B200_SMS * 8));
        if (T == 256) rowpat_build_kernel<256><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.pcodes.ptr(), E.pmeta.ptr(), E.rowcodes.ptr(), (RowPattern *)E.rpat.ptr(), E.rmeta.ptr(), stats.ptr());
        else rowpat_build_kernel<128><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.pcodes.ptr(), E.pmeta.ptr(), E.rowcodes.ptr(), (RowPattern *)E.rpat.ptr(), E.rmeta.ptr(), stats.ptr());
        count_launch();
    }
    AMGXB_LAUNCH_CHECK();
    const std::vector<int> h = stats.to_host(s);
    E.tiles_rowpat = h[0];
    E.max_rplen = h[1];
}

// Stage layout and occupancy of the encoded kernel for this level, from what its tiles use (see the kernel's header comment).
static void finalize_layout(Matrix &A, cudaStream_t s)
{
    ColEnc &E = A.colenc;
    const int T = A.plan.tile_rows, nt = E.num_tiles;
    const size_t msz = prec_size(A.mat_prec);
    // what the tiles that are NOT pair-coded need (pair-coded tiles stage one code byte per entry and their table, no value stream)
    int raw_c = E.tiles_raw, off16 = E.tiles_off16, val_raw = nt - E.tiles_val8, val8 = E.tiles_val8;
    if (E.tiles_pair > 0) {
        const std::vector<int> hm = E.meta.to_host(s), hp = E.pmeta.to_host(s);
        raw_c = off16 = val_raw = val8 = 0;
        for (int t = 0; t < nt; t++) {
            if (hp[t] > 0) continue;
            raw_c += hm[(size_t)META * t] == 0;
            off16 += hm[(size_t)META * t] == 2;
            val_raw += hm[(size_t)META * t + 2] == 0;
            val8 += hm[(size_t)META * t + 2] == 1;
        }
    }
    E.col_w = raw_c > 0 ? 4 : (off16 > 0 ? 2 : 1);
    E.val_w = val_raw > 0 ? (int)msz : (val8 > 0 ? 1 : 0);
    E.dict_cap = std::max(std::max(E.max_dlen, E.tiles_off16 > 0 ? 4 : 0), E.max_rplen * (int)(sizeof(RowPattern) / 4));      // ints; the region also holds a tile's row patterns
    E.vdict_cap = (int)std::max((size_t)E.max_vdlen * msz, (size_t)E.max_pdlen * (msz == 8 ? 16 : 8));      // bytes
    const size_t cap = (size_t)A.plan.max_tile_nnz;
    const size_t stage = align16(cap * E.val_w) + align16(cap * E.col_w) + (size_t)E.dict_cap * 4 + align16((size_t)E.vdict_cap) + (size_t)(T + 4) * 4 + 16;
    static const int env_stages = getenv("AMGXB_ENC_STAGES") ? atoi(getenv("AMGXB_ENC_STAGES")) : 0;
    static const int env_ctas = getenv("AMGXB_ENC_CTAS") ? atoi(getenv("AMGXB_ENC_CTAS")) : 0;
    const int by_threads = std::min(2048 / (T + PRODUCER_THREADS), 65536 / ((T + PRODUCER_THREADS) * 40));   // threads and registers (40 per thread, -Xptxas -v)
    E.on = false;
    int best_st = 0, best_ctas = 0;
    for (int st = (env_stages >= 2 && env_stages <= MAX_STAGES) ? env_stages : MAX_STAGES; st >= 2; st--) {
        const size_t smem = 512 + (size_t)st * stage;
        if (smem > (size_t)216 * 1024) continue;
        const int ctas = std::max(1, std::min(by_threads, (int)((size_t)227 * 1024 / (smem + 1024))));
        if (ctas > best_ctas) { best_ctas = ctas; best_st = st; }      // most resident CTAs; ties keep the deeper pipeline
        if (env_stages) break;
    }
    if (!best_st) return;
    E.stages = best_st;
    E.ctas_per_sm = env_ctas > 0 ? std::min(env_ctas, best_ctas) : best_ctas;
    E.smem_bytes = 512 + (size_t)best_st * stage;
    E.on = (E.tiles_dict8 + E.tiles_off16 + E.tiles_val8) > 0;        // nothing to gain when every tile stays raw
}

// called at the end of csr_build_plan
void csr_build_colenc(Matrix &A, cudaStream_t s)
{
    A.colenc.on = false;
    A.colenc.values_encoded = false;
    if (!colenc_requested() || !A.plan.use_tiles || A.n == 0 || A.bs() != 1) return;
    EncSeg seg[2];
    const int nseg = enc_segments(A, seg);
    const int T = A.plan.tile_rows, nt = seg[nseg - 1].base + seg[nseg - 1].tiles;
    const size_t msz = prec_size(A.mat_prec);
    ColEnc &E = A.colenc;
    E.num_tiles = nt;
    E.meta.resize((size_t)META * nt);
    E.meta.zero(s);                                  // encoding 0 everywhere: raw columns, raw values
    E.tiles_dict8 = E.tiles_off16 = E.tiles_val8 = 0;
    E.max_dlen = E.max_vdlen = 0;
    E.tiles_raw = nt;
    if (colenc_flags() & 1) {
        E.codes.resize(align16((size_t)2 * ((size_t)A.nnz + 8)) + (size_t)32 * nt + 64);
        E.codes.zero(s);
        E.dict.resize((size_t)nt * DICT_SLOTS);
        E.dict.zero(s);
        DevBuf<int> stats;
        stats.resize(8);
        stats.zero(s);
        for (int g = 0; g < nseg; g++) {
            const int grid = std::max(1, std::min(seg[g].tiles, B200_SMS * 8));
            if (T == 256) colenc_build_kernel<256><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.codes.ptr(), E.dict.ptr(), E.meta.ptr(), stats.ptr());
            else colenc_build_kernel<128><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), seg[g].row0, seg[g].row1, seg[g].tiles, seg[g].base, E.codes.ptr(), E.dict.ptr(), E.meta.ptr(), stats.ptr());
            count_launch();
        }
        AMGXB_LAUNCH_CHECK();
        const std::vector<int> h = stats.to_host(s);
        E.tiles_dict8 = h[0];
        E.tiles_off16 = h[1];
        E.tiles_raw = h[2];
        E.max_dlen = h[4];
    } else {
        E.codes.resize(64);                          // never dereferenced (every tile has column encoding 0); keeps the pointers valid
        E.dict.resize(64);
    }
    if (colenc_flags() & 2) {
        E.vcodes.resize(align16((size_t)A.nnz + 8) + (size_t)32 * nt + 64);
        E.vcodes.zero(s);
        E.vdict.resize((size_t)nt * DICT_SLOTS * msz);
        E.vdict.zero(s);
        build_value_codes(A, s);
        E.values_encoded = true;
    } else {
        E.vcodes.resize(64);
        E.vdict.resize(64);
    }
    E.xahead.resize((size_t)std::max(nt, 1));
    if (colenc_flags() & 1) {
        xahead_kernel<<<ceil_div(nt, 256), 256, 0, s>>>(E.meta.ptr(), E.dict.ptr(), nt, E.xahead.ptr());
        count_launch();
        AMGXB_LAUNCH_CHECK();
    } else E.xahead.zero(s);
    build_pair_codes(A, s);
    build_row_patterns(A, s);
    finalize_layout(A, s);
    if (getenv("AMGXB_COLENC_VERBOSE"))
        fprintf(stderr, "[amgx_b200] colenc level %d: %d tiles of %d rows: columns dict8 %d, off16 %d, raw %d; values dict8 %d; pair-coded %d, row patterns %d | stage widths col %d val %d B, dict %d / %d B, "
                        "%d stages, %zu B smem, %d CTAs/SM%s\n", A.level, nt, T, E.tiles_dict8, E.tiles_off16, E.tiles_raw, E.tiles_val8, E.tiles_pair, E.tiles_rowpat, E.col_w, E.val_w, E.dict_cap,
                E.vdict_cap, E.stages, E.smem_bytes, E.ctas_per_sm, E.on ? "" : " (off)");
}

// The values of A were changed in place (AMGX_matrix_replace_coefficients, DIAGONAL_SYMMETRIC scaling): the value codes follow them.
void csr_values_changed(Matrix &A, cudaStream_t s)
{
    ColEnc &E = A.colenc;
    csr_window_values_changed(A, s);
    if (!E.values_encoded) return;
    const int nt = A.plan.num_tiles;
    // the column half of the per-tile descriptors stays, the value half is rewritten by the build kernel
    build_value_codes(A, s);
    build_pair_codes(A, s);
    build_row_patterns(A, s);
    finalize_layout(A, s);
    (void)nt;
}

// csr_op entry of the encoded path; returns false when the caller must use the plain kernels
bool csr_op_enc(const Matrix &A, CsrEpi epi, const CsrOpArgs &g, cudaStream_t s, int segment)
{
    if (!A.colenc.on || g.agg) return false;
    EncSeg seg[2];
    const int nseg = enc_segments(A, seg);
    if ((nseg == 1) != (segment == 0)) return false;      // a split matrix applied as a whole (or the reverse): the plain kernels
    const EncSeg &sg = seg[segment == 2 ? 1 : 0];
    EncArgs ea;
    ea.tile_base = sg.base;
    ea.x_len = A.n;
    ea.codes = A.colenc.codes.ptr();
    ea.dict = A.colenc.dict.ptr();
    ea.meta = A.colenc.meta.ptr();
    ea.vcodes = A.colenc.vcodes.ptr();
    ea.vdict = A.colenc.vdict.ptr();
    ea.pcodes = A.colenc.pcodes.ptr();
    ea.pdict = A.colenc.pdict.ptr();
    ea.pmeta = A.colenc.pmeta.ptr();
    ea.rowcodes = A.colenc.rowcodes.ptr();
    ea.rpat = (const RowPattern *)A.colenc.rpat.ptr();
    ea.rmeta = A.colenc.rmeta.ptr();
    ea.xahead = A.colenc.xahead.ptr();
    ea.val_w = A.colenc.val_w;
    ea.col_w = A.colenc.col_w;
    ea.dict_cap = A.colenc.dict_cap;
    ea.vdict_cap = A.colenc.vdict_cap;
    AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
        TileArgs<MatT, VecT> ta;
        ta.row_ptr = A.row_ptr.ptr();
        ta.col = A.col_idx.ptr();
        ta.val = A.values.as<MatT>();
        ta.n = sg.row1;
        ta.row0 = sg.row0;
        ta.num_tiles = sg.tiles;
        ta.cap = A.plan.max_tile_nnz;
        ta.stages = A.colenc.stages;
        ta.unroll = 8;
        ta.perm = A.plan.use_perm ? A.tile_perm.ptr() : nullptr;
        ta.tile_base = sg.base;
        ta.l2pf = (l2_prefetch_flags() & 2) != 0;
        ta.x = (const VecT *)g.x;
        ta.agg = nullptr;
        ta.b = (const VecT *)g.b;
        ta.d = (const MatT *)g.d;
        ta.y = (VecT *)g.y;
        ta.omega = g.omega;
        ta.red = g.red;
        ta.fin_op = g.fin_op;
        ta.fin_slot = g.fin_slot;
        ta.mirror = g.mirror;
        switch (epi) {
        case EPI_SPMV: launch_enc_epi<MatT, VecT, EPI_SPMV>(A, ta, ea, s); break;
        case EPI_RESID: launch_enc_epi<MatT, VecT, EPI_RESID>(A, ta, ea, s); break;
        case EPI_ADD: launch_enc_epi<MatT, VecT, EPI_ADD>(A, ta, ea, s); break;
        case EPI_JACOBI:
        case EPI_JACOBI_L1: launch_enc_epi<MatT, VecT, EPI_JACOBI>(A, ta, ea, s); break;
        case EPI_SPMV_DOT: launch_enc_epi<MatT, VecT, EPI_SPMV_DOT>(A, ta, ea, s); break;
        case EPI_JACOBI_DOT: launch_enc_epi<MatT, VecT, EPI_JACOBI_DOT>(A, ta, ea, s); break;
        case EPI_RESID_NRM2: launch_enc_epi<MatT, VecT, EPI_RESID_NRM2>(A, ta, ea, s); break;
        }
    });
    return true;
}

}  // namespace amgxb
