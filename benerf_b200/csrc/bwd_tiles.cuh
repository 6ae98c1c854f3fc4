// Tile format of the training path and the two tensor-core kernels of the backward pass that consume it (bwd_tiles.cu).
//
// A "tile matrix" of width W (64, 128 or 256 columns) holds a [rows, W] activation / gradient matrix as 128-row tiles of
// pre-split 16-bit operands in the exact shared-memory image tcgen05.mma reads:
//
//   tile t : | hi part: W/64 K-blocks | lo part: W/64 K-blocks |        K-block = [128 rows x 64 cols] 16-bit, 16 KB,
//                                                                        rows of 128 B, SWIZZLE_128B (16-byte chunk ^= row & 7)
//
// value = hi + lo, both bf16 (keeps the fp32 exponent range, so gradients need no scaling; kind::f16 MMAs want one format
// for both operands).  Activations are written by the forward kernel's epilogue (mlp_tc3.cu; mlp_tc2.cu in its mode) next to its own fp16 A
// operand, gradients by the dgrad epilogues.  The
// same block serves as a K-major operand (contraction over its 64 columns: dgrad) and as an MN-major operand
// (contraction over its rows: wgrad), so no kernel of the backward pass converts or transposes anything: operands move
// HBM -> shared memory with cp.async.bulk and go straight into the tensor core.
#pragma once
#include "common.cuh"

namespace bnrf {
namespace bwt {

constexpr int kTileRows = 128;
constexpr size_t kKbBytes = 16384;
__host__ __device__ inline size_t tile_part_bytes(int W) { return (size_t)(W / 64) * kKbBytes; }
__host__ __device__ inline size_t tile_bytes(int W) { return 2 * tile_part_bytes(W); }
__host__ __device__ inline int64_t tile_count(int64_t rows) { return (rows + kTileRows - 1) / kTileRows; }
// the forward kernel walks tiles in pairs, so tile matrices are allocated with an even tile count
__host__ __device__ inline int64_t tile_alloc(int64_t rows) { return 2 * ((rows + 2 * kTileRows - 1) / (2 * kTileRows)); }

// 8 fp32 values -> packed bf16 hi / lo words (value = hi + lo to ~2^-17)
__device__ __forceinline__ void split8_bf16_pub(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        const float b0 = __uint_as_float(h[i] << 16), b1 = __uint_as_float(h[i] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(v[2 * i + 1] - b1), "f"(v[2 * i] - b0));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

enum { DG_TILE = 0, DG_TILE_MASKED = 1, DG_F32_STORE = 2, DG_F32_ACCUM = 3 };

// out[row, n] = sum_k A[row, k] * B[n, k]  (+ r_row[row] * r_col[n]) (* (mask[row, n] != 0))
struct DgradArgs {
    const unsigned char* a_tiles; int K;        // bf16 tile matrix of width K = 128 | 256
    const unsigned char* b_img; int N;          // packed weights: per 64-wide K-block [hi: N x 128 B][lo: N x 128 B], SW128; N = 64 | 256
    int epi;                                    // DG_*
    const unsigned char* mask_tiles;            // DG_TILE_MASKED: activation tile matrix of width 256 (hi part is read)
    const float* r_row; int64_t r_stride; const float* r_col;   // optional rank-1 term (r_row == NULL: none)
    unsigned char* out_tiles;                   // DG_TILE*: bf16 tile matrix of width N (= 256)
    float* out_f32; int64_t ld_out;             // DG_F32_*: [rows, ld_out] fp32 (N = 64)
    int64_t rows; int tiles;
};

// dW[m, col0 + n] += sum_rows dz[row, m] * h[row, n]  for n < n_valid;  dB[m] += sum_rows dz[row, m];
// dWv[n] += sum_rows wrow[row * wrow_stride] * h[row, n];  dBv[0] += sum_rows wrow[row * wrow_stride]
struct WgradJob {
    const unsigned char* dz_tiles; int M;       // bf16 tile matrix, width M = 128 | 256
    const unsigned char* h_tiles; int N;        // bf16 tile matrix, width N = 64 | 256
    float* dW; int ldw; int col0; int n_valid;
    float* dB;                                  // nullable
    const float* col_scale;                     // nullable: the partial of column n is multiplied by col_scale[n] (BARF c2f weights)
    const float* wrow; int wrow_stride; float* dWv; float* dBv;   // nullable (alpha_linear rides on the feature job)
    int cta0, ctas;                             // CTAs [cta0, cta0 + ctas) split the tiles of this job
};
constexpr int kMaxWgradJobs = 12;
struct WgradParams { WgradJob job[kMaxWgradJobs]; int n_jobs; int tiles; int64_t rows; };

// wgrad_pair.cu: the wide jobs on CTA pairs.  a_tiles: the M-side operand (width 256 = accumulator rows), b_tiles: the N-side
// operand (width NB = 256 | 128).
struct WgradPairJob {
    const unsigned char* a_tiles; const unsigned char* b_tiles; int NB;
    float* dW; int ldw; int col0; int n_valid;   // !transposed: dW[m * ldw + col0 + n] += D[m][n] for n < n_valid
    int transposed;                              //  transposed: dW[n * ldw + m] += D[m][n]  (the merged view step's G)
    float* dB;                                   // += colsum(A operand) (nullable)
    const float* wrow; int wrow_stride; float* dWv; float* dBv;   // nullable: dWv[m] += sum_rows wrow[row] * A[row][m], dBv += sum wrow
    const float* wrow_base; int wrow_col;        // wrow == wrow_base + wrow_col: column wrow_col of a [rows, wrow_stride] fp32 matrix (16-byte aligned rows)
    int64_t unit0; int weight;                   // assigned by the launcher: position on the work line, units per row tile
};
constexpr int kMaxPairJobs = 8;
struct WgradPairParams { WgradPairJob job[kMaxPairJobs]; int n_jobs; int tiles; int64_t rows; int64_t units; };
int launch_tile_wgrad_pair(bnrf_ctx* ctx, WgradPairParams& p, cudaStream_t st);

int launch_tile_dgrad(bnrf_ctx* ctx, const DgradArgs& a, cudaStream_t st);
int launch_tile_wgrad(bnrf_ctx* ctx, WgradParams& p, cudaStream_t st);   // assigns cta0 / ctas
// fp32 [rows, W] (row stride ld) -> tile matrix (fmt 0 = fp16 hi/lo, 1 = bf16 hi/lo); rows beyond `rows` are zero
int launch_to_tiles(bnrf_ctx* ctx, const float* src, int64_t rows, int W, int64_t ld, int fmt, unsigned char* tiles, cudaStream_t st);
int launch_from_tiles(bnrf_ctx* ctx, const unsigned char* tiles, int64_t rows, int W, int fmt, float* dst, cudaStream_t st);
// weights as dgrad B operands: image i <- fp32 [N, K] row-major (rows k0 .. k0 + N of the k-major copy wt[s], i.e.
// B[n][k] = wt[s][k0 + n][k]); all images of a network in one launch
struct DgImageSeg { const float* src; unsigned char* img; int N, K; };
struct DgImageTable { DgImageSeg seg[12]; int n; };
int pack_dgrad_images(bnrf_ctx* ctx, const DgImageTable& t, cudaStream_t st);
__host__ __device__ inline size_t dgrad_image_bytes(int N, int K) { return (size_t)(K / 64) * 2 * N * 128; }

}  // namespace bwt
}  // namespace bnrf
