// Packed-parameter layout shared by the MLP kernels (pack / SIMT reference / tcgen05 fwd+bwd).
//
// The 8x256 NeRF MLP (reference models/nerf.py:107-127, forward :129-175) has 12 nn.Linear:
//   id 0..7  xyz_encoding_1..8   (256 x {63,256,256,256,319,256,256,256})  ReLU
//   id 8     xyz_encoding_final  (256 x 256)                               no activation
//   id 9     dir_encoding.0      (128 x 256)                               ReLU
//   id 10    sigma               (1 x 256)                                 raw
//   id 11    rgb.0               (3 x 128)                                 sigmoid
//
// What the tensor cores run ("GEMM layers" g = 0..9):
//   g 0..7   the eight trunk layers, N = 256
//   g 8      the HEAD layer, N = 144, K = 256 (input h8):
//              rows   0..127  W' = W_dir . W_final   (xyz_encoding_final has no activation, so
//                             dir(final(h8)) = (W_dir W_final) h8 + (W_dir b_final + b_dir): one
//                             128x256 layer instead of 256x256 followed by 128x256; exact in real
//                             arithmetic, and one bf16 rounding of the intermediate less)
//              row    128     w_sigma  (the density head rides along as one more output column)
//              rows 129..143  zero     (UMMA N must be a multiple of 16)
//   g 9      the rgb head, N = 16 (rows 0..2 = W_rgb, rest zero), K = 128 (input c = relu(head[0:128]))
// The fused W', b' are formed in fp32 by the pack kernel; the backward returns dW', db' and
// mlp_unfuse_grad_kernel applies the chain rule to W_final / W_dir / b_final / b_dir.
//
// packed buffer = [ fwd images | dgrad images | fp32 bias block | fused fp32 block | flat fp32 copy ]
//  * image = one K-chunk (64 bf16 = 128 B per row) of a B operand, rows = output features,
//    stored exactly as the UMMA K-major SWIZZLE_128B canonical layout wants it in shared
//    memory (16-byte unit u of row r lives at r*128 + ((u ^ (r&7))*16), 8-row groups 1024 B
//    apart), so one cp.async.bulk (TMA bulk copy) moves a chunk with no tensor map.
//  * fwd chunk order = streaming order of the forward kernel: for g, for kc, then the layer's
//    bias slab.  Layer 0 has one chunk (63 inputs + zero pad); layer 4 has five: chunk 0 =
//    encoding columns 0..62 of W5, chunks 1..4 = hidden columns 63..318.
//  * bias slab (g 0..8): the bias enters through the tensor core as one more K = 16 step against
//    the encoding image's K-step 3, whose last column (63, the pad) holds the constant 1 (or the
//    row's scale in tangent mode).  B operand of that step: N rows x 16 k-values, bf16, all zero
//    except k = 15 = bias[row]; K-major without swizzle: 8-row groups of 256 B = two 128-byte core
//    matrices (k 0..7, k 8..15), element (row r, k) at (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2.
//    The density bias and the rgb bias are added in fp32 by the epilogue (slab rows 128.. are zero).
//  * a CTA pair (cta_group::2) splits every B operand by rows: CTA rank r stages rows
//    [r*N/2, (r+1)*N/2) = the r-th half of the bytes of a chunk image / slab.
//  * dgrad images hold W^T (rows = input features, K = output features) in streaming order of
//    the backward kernel (see mlp_bwd.cu).
#pragma once
#include <stdint.h>

namespace mlp {

constexpr int NLIN = 12;
constexpr int NG = 10;
constexpr int W = 256;
constexpr int ENC = 63;
constexpr int HEAD_N = 144;         // 128 colour features + sigma + pad
constexpr int RGB_N = 16;           // 3 + pad

__host__ __device__ constexpr int lin_out(int id) { return id <= 8 ? 256 : id == 9 ? 128 : id == 10 ? 1 : 3; }
__host__ __device__ constexpr int lin_in(int id) { return id == 0 ? 63 : id == 4 ? 319 : id == 11 ? 128 : 256; }

// flat fp32 layout (also the layout of the gradient vector): for id: weight (out*in) then bias (out)
__host__ __device__ constexpr int64_t flat_w_off(int id) {
    int64_t o = 0;
    for (int i = 0; i < id; ++i) o += (int64_t)lin_out(i) * lin_in(i) + lin_out(i);
    return o;
}
__host__ __device__ constexpr int64_t flat_b_off(int id) { return flat_w_off(id) + (int64_t)lin_out(id) * lin_in(id); }
constexpr int64_t FLAT_FLOATS = flat_w_off(NLIN);                 // 592 388
// gradient vector = flat layout + scratch for the fused head layer: dW' (128 x 256) then db' (128)
constexpr int64_t GRAD_FUSED_W = FLAT_FLOATS;
constexpr int64_t GRAD_FUSED_B = GRAD_FUSED_W + 128 * 256;
constexpr int64_t GRAD_FLOATS = GRAD_FUSED_B + 128;

// ---- forward images
__host__ __device__ constexpr int g_N(int g) { return g < 8 ? 256 : g == 8 ? HEAD_N : RGB_N; }
__host__ __device__ constexpr int g_chunks(int g) { return g == 0 ? 1 : g == 4 ? 5 : g == 9 ? 2 : 4; }
__host__ __device__ constexpr uint32_t g_chunk_bytes(int g) { return (uint32_t)g_N(g) * 128u; }
__host__ __device__ constexpr uint32_t g_bias_bytes(int g) { return g < 9 ? (uint32_t)g_N(g) * 32u : 0u; }
__host__ __device__ constexpr int64_t fwd_chunk_off(int g, int kc) {     // kc == g_chunks(g): the bias slab
    int64_t o = 0;
    for (int i = 0; i < g; ++i) o += (int64_t)g_chunks(i) * g_chunk_bytes(i) + g_bias_bytes(i);
    return o + (int64_t)kc * g_chunk_bytes(g);
}
__host__ __device__ constexpr int64_t fwd_bias_off(int g) { return fwd_chunk_off(g, g_chunks(g)); }
constexpr int64_t FWD_BYTES = fwd_chunk_off(NG, 0);
constexpr int FWD_CHUNKS = 1 + 4 * 3 + 5 + 4 * 3 + 4 + 2;         // 36
constexpr int FWD_SLABS = 9;
// byte offset of element (row r, k in [0,16)) inside a bias slab
__host__ __device__ constexpr uint32_t slab_off(int r, int k) {
    return (uint32_t)(r >> 3) * 256u + (uint32_t)(k >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 7) * 2u;
}

// ---- dgrad images (W^T): step s of the backward kernel, see mlp_bwd.cu
//  s 0: rgb^T   rows 128 (c features),  K = 16  (1 chunk, K-step 0 only: columns 0..2 = W_rgb[j][row])
//  s 1: head^T  rows 256 (h8 features), K = 144 (chunks 0,1 = W'^T; chunk 2: column 0 = w_sigma[row], K-step 0 only)
//  s 2..4: g7,g6,g5 ^T  rows 256, K 256 (4 each)
//  s 5: g4^T hidden part: rows = in 63..318 (256), K 256 (4)
//  s 6: g4^T encoding part: rows = in 0..62 (+1 zero row) = 64, K 256 (4 chunks of 8 KB)
//  s 7..9: g3,g2,g1 ^T rows 256, K 256 (4 each)
//  s 10: g0^T rows 64 (63 + zero), K 256 (4 chunks of 8 KB)
constexpr int NBS = 11;
__host__ __device__ constexpr int bs_layer(int s) { return s == 0 ? 9 : s == 1 ? 8 : s <= 4 ? 9 - s : s == 5 ? 4 : s == 6 ? 4 : s <= 9 ? 10 - s : 0; }
__host__ __device__ constexpr int bs_rows(int s) { return s == 0 ? 128 : (s == 6 || s == 10) ? 64 : 256; }
__host__ __device__ constexpr int bs_chunks(int s) { return s == 0 ? 1 : s == 1 ? 3 : 4; }
__host__ __device__ constexpr int bs_ksteps(int s, int kc) { return (s == 0 || (s == 1 && kc == 2)) ? 1 : 4; }
__host__ __device__ constexpr int bs_in0(int s) { return s == 5 ? 63 : 0; }             // first input feature of the rows
__host__ __device__ constexpr int bs_in_valid(int s) { return (s == 6 || s == 10) ? 63 : 256; }
__host__ __device__ constexpr uint32_t bs_chunk_bytes(int s) { return (uint32_t)bs_rows(s) * 128u; }
__host__ __device__ constexpr int64_t bwd_chunk_off(int s, int kc) {
    int64_t o = FWD_BYTES;
    for (int i = 0; i < s; ++i) o += (int64_t)bs_chunks(i) * bs_chunk_bytes(i);
    return o + (int64_t)kc * bs_chunk_bytes(s);
}
constexpr int64_t IMG_BYTES = bwd_chunk_off(NBS, 0);
constexpr int BWD_CHUNKS = 1 + 3 + 4 * 9;                         // 40

// ---- fp32 bias block: [10][256], row g = bias the accumulators of GEMM layer g start from
//      (g 8: b' (128), b_sigma, zeros;  g 9: b_rgb (3), zeros)
constexpr int64_t SMALL_OFF = (IMG_BYTES + 1023) / 1024 * 1024;   // bytes
constexpr int SM_BIAS = 0;
constexpr int SMALL_FLOATS = 2560;
// ---- fused fp32 block: W' (128 x 256 row-major) then b' (128)
constexpr int64_t FUSED_OFF = SMALL_OFF + ((SMALL_FLOATS * 4 + 1023) / 1024) * 1024;
constexpr int FUSED_FLOATS = 128 * 256 + 128;
constexpr int64_t FLAT_OFF = FUSED_OFF + ((FUSED_FLOATS * 4 + 1023) / 1024) * 1024;
constexpr int64_t PACKED_BYTES = FLAT_OFF + ((FLAT_FLOATS * 4 + 1023) / 1024) * 1024;

// byte offset of element (row r, column c in [0,64)) inside a chunk image
__host__ __device__ constexpr uint32_t img_off(int r, int c) {
    return (uint32_t)r * 128u + (uint32_t)((((c >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1));
}

// ---- tiling: the forward / dgrad kernels run as CTA pairs; one pair iteration owns PAIR_POINTS compacted points
//      (CTA rank r: points [r*256, r*256+256) of the iteration, two 128-row tiles each), so the point images are
//      always written in whole groups of 4 tiles: tile index = compact point index / 128
constexpr int PAIR_POINTS = 512;
__host__ __device__ constexpr int64_t n_tiles_for(int64_t n) { return (n + PAIR_POINTS - 1) / PAIR_POINTS * 4; }
// ---- forward stash per 128-row tile (training): bf16 images + 1-bit ReLU masks
constexpr int64_t ST_ENC = 0;                      // 16 KB image
constexpr int64_t ST_H = 16384;                    // h1..h8: 8 x 64 KB images
constexpr int64_t ST_C = ST_H + 8 * 65536;         // c = relu(head[0:128]): 32 KB (2 chunks)
constexpr int64_t ST_MASK = ST_C + 32768;          // h1..h8 masks: 8 x [8 blocks][128 rows] x 4 B
constexpr int64_t ST_CMASK = ST_MASK + 8 * 4096;   // [4 blocks][128 rows] x 4 B
constexpr int64_t ST_TILE = ST_CMASK + 2048;       // 608 256
// ---- dY scratch per 128-row tile (backward): pre-activation gradient images
constexpr int64_t DY_RGB = 0;                       // d rgb_pre: 1 chunk, columns 0..2 (16 KB)
constexpr int64_t DY_HEAD = 16384;                  // d head_pre: chunks 0,1 = d c_pre, chunk 2 column 0 = d sigma (48 KB)
constexpr int64_t DY_H = DY_HEAD + 49152;           // dpre of trunk layer g (0..7) at DY_H + g*64 KB
constexpr int64_t DY_TILE = DY_H + 8 * 65536;       // 589 824

}  // namespace mlp
