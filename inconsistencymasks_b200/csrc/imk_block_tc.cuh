// Descriptors of the block-fused tcgen05 engine (imk_block_tc.cu).
#pragma once
#include <cuda.h>
#include <vector>
#include "imk_common.cuh"

namespace imk {

constexpr int kBtMaxBlocks = 24;        // M blocks (128 flat positions) per stage per tile
constexpr int kBtHeadMaxK = 3;          // widest output layer the in-epilogue head takes (on at most 32 channels)

struct BtStage {
    int taps, ksteps, n, nb;            // 1|9, Cin_p/16, Cout_p (UMMA N), M blocks
    int w_off;                          // byte offset of [tap][kstep][2][n][8] fp16 in the weight region
    int par_off;                        // float offset of bias[n] | scale[n] | shift[n]
    int col;                            // first TMEM column of the stage's accumulator region
    int cs;                             // TMEM columns between consecutive M blocks (n, or 8 for an n8 stage)
    // 8-channel maps (int(16 * alpha) <= 8, the reference's alpha = 0.5 networks) keep ONE 16-byte plane per position:
    //   kin8: the stage's A operand is a single plane.  K = 16 of an MMA then spans TWO positions: the descriptor's
    //         leading-dimension offset is the distance between two TAPS (f, f + 1: LBO = 16 bytes; (0,2) -> (1,0):
    //         LBO = (pitch - 2) * 16), so a 3x3 stage is 5 MMAs instead of 9 (operand B = the two taps' weights stacked,
    //         zeros under the odd tap); a 1x1 stage reads its plane twice (LBO = 0) against zero weights in the second chunk
    //   n8:   only accumulator columns 0..7 are real (UMMA N stays 16): the epilogue loads 8 columns and writes one plane
    int kin8, n8;
};

struct BtArgs {
    const void *in;                     // uint8 image (FRONT) or fp16 NHWC map
    const __half *in_lo;                // DEC: the half-resolution map that is upsampled and added
    __half *out;
    __half *out_pool;                   // optional: MaxPooling2D 2x2 of `out` (fp16 NHWC at half resolution), written by the store warp
    const uint8_t *wpk;                 // packed weights of all stages (device)
    const float *par;                   // bias / BN parameters of all stages (device)
    int w_bytes, par_floats;
    int H, W, in_c, swap_rb, load_kind, ld_cp;      // ld_cp: channels (padded) of the loaded operand
    int in_f32;                         // FRONT: the image is float32 [N,H,W,c] instead of uint8
    int Th, Tw, pitch, tiles_x, tiles_y;
    long long n_tiles;
    int has_s1;
    BtStage s1, s2, s3;
    int Pn0, Pn1, Pn2;
    int par_off_b, a0_off, a1_off, a2_off, bar_off; // byte offsets in dynamic shared memory
    int a1_stride;                      // A1 is double-buffered (tile parity): bytes between the two copies
    unsigned pitch_magic;               // ceil(2^32 / pitch)
    int tmem_cols;                      // TMEM columns to allocate (power of two >= the three accumulator regions)
    int o_off;                          // output staging tile [Th][Tw][out_c] fp16
    int out_c;                          // channels per pixel of the output map in HBM (s3.n, or 8 when s3.n8)
    int lut_off;                        // FRONT: 256-entry table of x/255 as fp16 hi | lo << 16
    // FRONT, uint8 images: raw tile staging by TMA (u8_issue): two buffers of `u8_panels` panels [Th + 2][u8_pw] bytes
    int u8_tma, u8_off, u8_pw, u8_panels, u8_pstride, u8_bstride;
    unsigned u8_pw_magic;               // ceil(2^32 / u8_pw)
    int tm_flat8;                       // tm_in describes an 8-channel map as uint64 [N][H][2W] (contiguous tile rows)
    alignas(64) CUtensorMap tm_in;      // TMA map (ENC / DEC): {8 ch, pitch, Th + 2, 1} boxes of the fp16 NHWC input / skip map
    // Epilogue constants, read as constant-bank operands (no loads): the BN scale is folded into the weights, so a
    // stage's epilogue is  h = fp16(acc + cpar[s][c]);  h = min(max(h, clo[s][c]), chi[s][c])  on packed halves
    //   scale > 0: (b', lo, hi) = (scale*bias + shift, shift, +inf)   scale < 0: (.., -inf, shift)   ReLU only: (bias, 0, +inf)
    float cpar[3][64];
    uint32_t clo[3][32], chi[3][32];    // half2 pairs of the clamp bounds, rounded to fp16
    int has_hi[3];                      // some channel of the stage has a finite upper bound (a BN scale <= 0)
    // load_kind 3 (FRONT, grayscale uint8 images): the input block x/255 -> 1x1 conv + ReLU -> BN (unet.py:4-9) is a
    // 256-entry table of finished fp16 rows, built in fp32 at kernel start; the loader warps copy rows straight into the
    // 3x3 stage's operand buffer: v = x/255 * fw[0][ch] + fb[ch], clamped to [flo, fhi] (BN scale folded as for the stages)
    float fw[4][32], fb[32], flo[32], fhi[32];
    // Head (level-0 decoder of a network with K <= 3 outputs on <= 32 channels -- the reference's binary configs):
    // the output layer `out` (unet.py:63) runs INSIDE the last epilogue.  E3 already holds a pixel's finished c9 row
    // (fp16 values) in registers: K * C0 fp32 FMAs with the weights as constant-bank operands give the logits
    // (p = b; p = fma(x_c, w_c, p) in channel order, the arithmetic of the generic output kernel), the activation /
    // decision follows in the same thread and ONE byte (or K floats) per pixel leaves the SM.  c9 is never written:
    // no staging tile, no bulk store, no 32 B/px map in HBM.
    // head_mode 0: fp32 probabilities [N,H,W,K] (.predict)   1: byte of threshold votes (bit k = head k fires)
    //           2: argmax class id
    int has_head, head_K, head_mode, head_act, head_strict;
    float head_thr, head_dstar;
    float hw[3][32], hb[4];
    float *head_probs;
    uint8_t *head_dec;
    int cgm1, cgm3;                     // commit-group size - 1 of the 1x1 stages (see commit_idx)
    int dbg_skip;                       // first tile (of CTA 0) the timeline records (IMK_BT_TL_SKIP)
    long long *dbg;                     // optional timeline buffer (IMK_BT_TIMELINE=1): [3 roles][16 tiles][8 events] clocks of CTA 0
};


struct FusedBlock {                     // one fused U-Net block of one model (device-resident packs + plan)
    bool ok = false;
    BtArgs args{};
    int w_bytes = 0, par_floats = 0;
    int ctas_per_sm = 1;                // launch shape: 1 -> block_tc_kernel<16, 8>, 2 -> block_tc_kernel<8, 4>
    size_t smem = 0;
};

// What the head of a level-0 decoder block writes (fused_block_build kind 4)
struct HeadOut {
    int mode;                           // 0: fp32 probabilities, 1: threshold votes (bit k of a byte), 2: argmax class id
    float thr, dstar;                   // mode 1: threshold; dstar > 0: the exact `1 + exp(-z) <= dstar` form of it
    int strict;
    float *probs;                       // mode 0: [n,H,W,K]
    uint8_t *dec;                       // mode 1 / 2: [n,H,W]
};

struct ConvHost {                       // host view of one Conv2D (+BN) while imk_unet_create runs
    const float *hwio, *bias;
    const float *bn_scale, *bn_shift;   // folded BN (gamma / sqrt(var + eps), beta - mean * scale) or null
    int ks, cin, cout;
};

// TMA map of an fp16 NHWC map [n, h, w, c] with boxes {8 ch, box_w, box_h, 1}; out-of-bounds elements read as zero
int make_map(CUtensorMap *map, const void *base, int64_t n, int h, int w, int c, int box_w, int box_h);

// kind: 0 FRONT (in 1x1, conv3, conv1 of level 0), 1 ENC (conv3, conv1), 2 DEC (conv1a, conv3, conv1b),
//       3 FRONT for uint8 images with the input block computed by the loader (a chain of two),
//       4 DEC + head (conv1a, conv3, conv1b, out): L[3] is the output layer, `head_act` its activation.
// Leaves fb.ok == false (and returns IMK_OK) when the block does not fit the resident-weight design.
// allow8: maps of <= 8 channels are stored / staged as ONE 8-channel plane (BtStage::kin8 / n8); every producer and
// consumer of such a map must then be a fused block built with allow8 (imk_unet_create checks this).
int fused_block_build(FusedBlock &fb, int kind_, const ConvHost *L, int H, int W, int in_c, std::vector<void *> &owned, int head_act = 0, bool allow8 = false);
// out_pool (optional, needs fused_block_can_pool): the 2x2 max-pooled map is written next to `out` by the same kernel.
bool fused_block_can_pool(const FusedBlock &fb);
int fused_block_launch(const FusedBlock &fb, const void *in, const __half *in_lo, __half *out, __half *out_pool, int64_t n,
                       int swap_rb, int in_f32, cudaStream_t stream, const HeadOut *head = nullptr);

}  // namespace imk
