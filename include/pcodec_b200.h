/*
 * pcodec_b200 — C-ABI of the B200 (sm_100a) hot path of the progressive codec.
 *
 * Drop-in boundary for EIDOSLAB/ProgressiveCodec's inference path
 * (ChannelProgresssiveWACNN.forward / compress / decompress).  Every entry point takes plain
 * pointers + sizes + an explicit CUDA stream (passed as void*, i.e. a cudaStream_t) and returns an
 * int status: 0 = OK, >0 = a pcodec error (PCODEC_ERR_*), <0 = -(cudaError_t).
 * No torch types cross this boundary.  Unless stated otherwise every pointer is a DEVICE pointer.
 *
 * Reference interfaces replaced (paths under /root/reference/src/compress/):
 *   cpp_exts/rans/rans_interface.cpp:99-204   RansEncoder / BufferedRansEncoder   -> pcodec_rans_encode_batch
 *   cpp_exts/rans/rans_interface.cpp:206-350  RansDecoder                         -> pcodec_rans_decode_batch
 *   cpp_exts/ops/ops.cpp:10-67                pmf_to_quantized_cdf                -> pcodec_pmf_to_quantized_cdf
 *   layers/masking.py:205-223                 ChannelMask 'point-based-std'       -> pcodec_quantile_threshold
 *   entropy_models/entropy_models.py:126-165, 661-666 quantize/dequantize/build_indexes
 *                                                                                 -> pcodec_slice_quantize / _indexes / _dequantize
 *   entropy_models/entropy_models.py:626-659  GaussianConditional likelihood      -> pcodec_slice_quantize (lik output)
 *   entropy_models/entropy_models.py:400-433, 446-522 EntropyBottleneck           -> pcodec_bottleneck_*
 *   models/utils.py:186-204, layers/layers.py:15-29, layers/gdn.py:50-63          -> pcodec_conv_taps
 *   layers/win_attention.py:84-115,153-207    shifted-window attention core       -> pcodec_window_attention
 *   layers/masking.py:171-194 (cust_map)      custom importance-map masks          -> pcodec_slice_quantize_cust
 *   models/CHProgREM.py:73-85,375-400         REM refinement gate                  -> pcodec_masked_residual
 *   (no reference equivalent) truncatable progressive container                   -> pcodec_layer_partition,
 *                                                                                    pcodec_rans_encode_segments / _decode_segments
 *
 * Activation layout: NHWC fp32 ("pixel-major"): element (n,h,w,c) of a tensor with pixel stride PS
 * lives at base[((n*H + h)*W + w)*PS + c]; PS >= C lets several tensors share one concat buffer.
 * Symbol / index / likelihood / mask tensors that feed the entropy coder use the reference's
 * NCHW order, because that order defines the bit stream (entropy_models.py:227-235).
 */
#ifndef PCODEC_B200_H_
#define PCODEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCODEC_OK 0
#define PCODEC_ERR_BAD_ARG 1
#define PCODEC_ERR_UNSUPPORTED 2
#define PCODEC_ERR_OVERFLOW 3 /* an output / scratch capacity was too small; retry with a larger one */

/* library / device info ------------------------------------------------------------------- */
int pcodec_version(void);
/* HOST. Fills sm count and compute capability of the current device; returns status. */
int pcodec_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* HOST. Number of kernels this library has launched since load / last reset (bench 'gpu_launches'). */
int64_t pcodec_launch_count(void);
void pcodec_reset_launch_count(void);
/* HOST. Debug: when on (also PCODEC_SYNC_LAUNCHES=1 in the environment at load time) every entry point synchronises
 * the device after its launches and names itself on stderr if the device faulted; the status is -cudaError_t.
 * The reference has no counterpart (python exceptions only). */
void pcodec_set_sync_launches(int on);
/* HOST. Debug: writes the entry points of the last (up to 64) launches of this process, oldest first, one per line
 * ("#seq thread <id> <entry point>") into buf; returns the number of characters written. */
int pcodec_recent_launches(char *buf, int cap);
/* HOST. Text for a status returned by any entry point (positive: PCODEC_ERR_*, negative: -cudaError_t). */
const char *pcodec_error_string(int status);

/* ------------------------------------------------------------------------------------------
 * Entropy coder
 * ---------------------------------------------------------------------------------------- */

/* HOST pointers. ops.cpp:10-67. cdf_out has n+1 entries. Returns PCODEC_ERR_BAD_ARG when no frequency
 * can be stolen for a zero-width bin (the reference asserts). */
int pcodec_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf_out);

/* Batched rANS encode of n_streams independent streams of n_per_stream symbols each
 * (one stream = one (image, slice) of the reference: rans_interface.cpp:193-204).
 *   symbols, indexes : int32 [n_streams][n_per_stream]
 *   cdfs             : int32 [n_tables][cdf_stride]; cdf_sizes, offsets: int32 [n_tables]
 *   scratch          : uint32 [n_streams][scratch_words]  (per-stream backward write area)
 *   stream_words     : int32 [n_streams] workspace; on return the number of 32-bit words of each stream
 *   out_bytes        : the streams, concatenated in stream order; capacity out_cap bytes
 *   out_offsets      : int64 [n_streams+1] byte offsets into out_bytes
 *   status           : int32 [1], set to PCODEC_ERR_OVERFLOW by the device when scratch_words or out_cap
 *                      was too small (the host reads it together with out_offsets)
 * The bytes of every stream are identical to RansEncoder.encode_with_indexes on the same inputs. */
int pcodec_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int n_streams, int64_t n_per_stream,
                             const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                             int n_tables, uint32_t *scratch, int64_t scratch_words, int32_t *stream_words,
                             uint8_t *out_bytes, int64_t out_cap, int64_t *out_offsets, int32_t *status, void *stream);

/* Batched rANS decode (rans_interface.cpp:206-275).  in_bytes/in_offsets as produced by the encoder (every
 * stream must start 4-byte aligned, i.e. offsets are multiples of 4, which holds for rANS word streams).
 *   out_symbols : int32 [n_streams][n_per_stream] */
int pcodec_rans_decode_batch(const uint8_t *in_bytes, const int64_t *in_offsets, int n_streams, int64_t n_per_stream,
                             const int32_t *indexes, const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                             const int32_t *offsets, int n_tables, int32_t *out_symbols, void *stream);

/* Variable-length streams ("segments") for the truncatable progressive container (the reference has
 * no container format — its compress() returns python lists, CHProg_cnn.py:847).  Stream s codes the elements
 * [seg_start[s], seg_start[s] + seg_count[s]) of the flat symbols / indexes arrays with the same arithmetic as
 * pcodec_rans_encode_batch (each segment is a stand-alone reference-decodable rANS stream). */
int pcodec_rans_encode_segments(const int32_t *symbols, const int32_t *indexes, const int64_t *seg_start,
                                const int32_t *seg_count, int n_streams, const int32_t *cdfs, int cdf_stride,
                                const int32_t *cdf_sizes, const int32_t *offsets, int n_tables, uint32_t *scratch,
                                int64_t scratch_words, int32_t *n_words, uint8_t *out_bytes, int64_t out_cap,
                                int64_t *out_offsets, int32_t *status, void *stream);
int pcodec_rans_decode_segments(const uint8_t *in_bytes, const int64_t *starts, const int64_t *ends, int n_streams,
                                const int64_t *seg_start, const int32_t *seg_count, const int32_t *indexes,
                                const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                                int n_tables, int32_t *out_symbols, void *stream);

/* REM wrapper (LatentRateReduction.forward tail + apply_latent_enhancement's attention mask, CHProgREM.py:73-85,
 * 375-400): out = ret * (star - bar) + identity, where star / bar are the variance-aware masks of the ORIGINAL sigma
 * at the current quality and at the preceding check level: mask_x = (mode_x == ONES) ? 1 : (mode_x == ZEROS) ? 0 :
 * (sigma >= thr_x[b]).  ret / identity / out have `channels` channels (32, or 64 when mu_std: the 32-channel mask is
 * used for both halves); sigma has 32 channels.  NHWC fp32 with pixel strides. */
int pcodec_masked_residual(const float *ret, int ret_ps, const float *identity, int id_ps, const float *sigma,
                           int sigma_ps, int batch, int64_t hw, int channels, int mode_star, const float *thr_star,
                           int mode_bar, const float *thr_bar, float *out, int out_ps, void *stream);

/* Progressive-layer partition of one 32-channel slice (replaces nothing in the reference: it turns the nested
 * variance-aware masks of masking.py:205-223 into an embedded layer order).  thresholds: float [n_levels][batch] in
 * non-increasing order per image (NaN = "everything", i.e. pr >= 10).  layer(e) = #{k : sigma_e < thr_k}.
 *   scatter == 0: in_a / in_b (NCHW-order planes int32 [batch][channels*hw], either may be NULL) are stably
 *                 partitioned by layer into out_a / out_b; counts int32 [batch][16] receives the layer sizes.
 *   scatter != 0: in_a holds layer-major compacted values; out_a[e] = value if layer(e) < avail[b] else 0. */
int pcodec_layer_partition(const float *sigma, int sigma_ps, int batch, int64_t hw, int channels, const float *thresholds,
                           int n_levels, const int32_t *in_a, const int32_t *in_b, int32_t *out_a, int32_t *out_b,
                           int32_t *counts, const int32_t *avail, int scatter, void *stream);

/* Same decoder for streams that are NOT consecutive in in_bytes: stream s occupies bytes [starts[s], ends[s]).  Lets
 * one launch decode several slices of a sub-batch (e.g. base slices 5..9, whose parameters do not depend on each
 * other, CHProg_cnn.py:878) out of a slice-major stream container. */
int pcodec_rans_decode_ranges(const uint8_t *in_bytes, const int64_t *starts, const int64_t *ends, int n_streams,
                              int64_t n_per_stream, const int32_t *indexes, const int32_t *cdfs, int cdf_stride,
                              const int32_t *cdf_sizes, const int32_t *offsets, int n_tables, int32_t *out_symbols,
                              void *stream);

/* HOST-only scalar walk over the same state arithmetic as the kernels (csrc/rans_core.h); exists so CPU unit
 * tests can pin that arithmetic against the oracle.  Not used by the product path.  Words are written at the END
 * of `words`; returns the number of words used or -1 on overflow. */
int64_t pcodec_selftest_rans_core_encode(const int32_t *symbols, const int32_t *indexes, int64_t n, const int32_t *cdfs,
                                         int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                                         uint32_t *words, int64_t cap_words);

/* ------------------------------------------------------------------------------------------
 * Variance-aware masking + quantisation + CDF index
 * ---------------------------------------------------------------------------------------- */

/* torch.quantile(scale[b].ravel(), q) for every image b, bit-compatible with ATen's linear interpolation
 * (masking.py:216).  scale: NHWC fp32 [B][HW][ps] of which C channels are used; thr: fp32 [B].
 * workspace: uint32 [B][4] scratch. */
int pcodec_quantile_threshold(const float *scale, int batch, int64_t hw, int channels, int pixel_stride, float q,
                              float *thr, uint32_t *workspace, void *stream);

/* mask modes */
#define PCODEC_MASK_ONES 0      /* pr >= 10, base slices, 'two-levels' pr != 0 */
#define PCODEC_MASK_ZEROS 1     /* pr == 0 */
#define PCODEC_MASK_THRESHOLD 2 /* mask = scale >= thr[b] */

/* Fused encoder-side slice step (CHProg_cnn.py:744-755, 819-834 / forward :626-632):
 *   v      = y - (y_sub ? y_sub : 0) - mu          (delta_encode subtracts the base slice first)
 *   m      = mask(scale)                            (mode above)
 *   sym    = rint(v * m)  [compress]   or   rint(v) * m  [forward]; both are the same integers
 *   index  = #{ j < n_levels-1 : table[j] < max(scale*m, scale_bound) }
 *   y_hat  = sym + mu
 *   lik    = Gaussian likelihood of sym under max(scale*m, bound), lower-bounded 1e-9 (optional)
 * Inputs are NHWC with their own pixel strides; symbols/indexes/mask/lik are written in NCHW order
 * [B][C][H*W]; y_hat is NHWC with its own stride.  Any of symbols/indexes/mask_out/lik/y_hat may be NULL. */
int pcodec_slice_quantize(const float *y, int y_ps, const float *y_sub, int y_sub_ps, const float *mu, int mu_ps,
                          const float *scale, int scale_ps, int batch, int64_t hw, int channels, int mask_mode,
                          const float *thr, const float *scale_table, int n_levels, float scale_bound,
                          int32_t *symbols, int32_t *indexes, float *mask_out, float *lik, float *y_hat, int y_hat_ps,
                          void *stream);

/* pcodec_slice_quantize with a custom importance map (ChannelMask.forward(cust_map=...), masking.py:171-194, reached
 * from compress/decompress at CHProg_cnn.py:721,823,850,964): in PCODEC_MASK_THRESHOLD mode the mask is
 * mask_src >= thr[b] instead of scale >= thr[b]; everything else (index from scale*mask, symbol, y_hat) is unchanged. */
int pcodec_slice_quantize_cust(const float *y, int y_ps, const float *y_sub, int y_sub_ps, const float *mu, int mu_ps,
                               const float *scale, int scale_ps, int batch, int64_t hw, int channels, int mask_mode,
                               const float *thr, const float *scale_table, int n_levels, float scale_bound,
                               int32_t *symbols, int32_t *indexes, float *mask_out, float *lik, float *y_hat,
                               int y_hat_ps, const float *mask_src, int mask_src_ps, void *stream);

/* Decoder-side: index (and mask) only (CHProg_cnn.py:891, 960-968). */
int pcodec_slice_indexes(const float *scale, int scale_ps, int batch, int64_t hw, int channels, int mask_mode,
                         const float *thr, const float *scale_table, int n_levels, float scale_bound,
                         int32_t *indexes, void *stream);

/* Decoder-side: y_hat = sym + mu (CHProg_cnn.py:894-896); symbols NCHW int32, mu / y_hat NHWC. */
int pcodec_slice_dequantize(const int32_t *symbols, const float *mu, int mu_ps, int batch, int64_t hw, int channels,
                            float *y_hat, int y_hat_ps, void *stream);

/* ------------------------------------------------------------------------------------------
 * EntropyBottleneck (z path)
 * ---------------------------------------------------------------------------------------- */

/* symbols = rint(z - median[c]) (NCHW int32), indexes = c, z_hat = sym + median[c] (NHWC).  z: NHWC. */
int pcodec_bottleneck_quantize(const float *z, int z_ps, const float *medians, int batch, int64_t hw, int channels,
                               int32_t *symbols, int32_t *indexes, float *z_hat, int z_hat_ps, void *stream);
/* z_hat = sym + median[c] from decoded NCHW symbols; also (re)writes indexes when non-NULL. */
int pcodec_bottleneck_dequantize(const int32_t *symbols, const float *medians, int batch, int64_t hw, int channels,
                                 float *z_hat, int z_hat_ps, void *stream);
/* indexes[b][c][hw] = c */
int pcodec_bottleneck_indexes(int batch, int64_t hw, int channels, int32_t *indexes, void *stream);
/* Factorised-density likelihood of z_hat (entropy_models.py:421-433), NCHW output.
 * params: per channel 58 floats: softplus(matrix0..4) (3,9,9,9,3), bias0..4 (3,3,3,3,1), tanh(factor0..3) (3,3,3,3). */
int pcodec_bottleneck_likelihood(const float *z_hat, int z_ps, const float *params, int batch, int64_t hw, int channels,
                                 float *lik, void *stream);

/* ------------------------------------------------------------------------------------------
 * Convolution as a sum of shifted-tap GEMMs (implicit GEMM, NHWC)
 * ---------------------------------------------------------------------------------------- */

#define PCODEC_MAX_SEGMENTS 4
#define PCODEC_MAX_TAPS 25

/* epilogues: acc is the fp32 accumulator + bias[co] */
#define PCODEC_EPI_LINEAR 0        /* acc */
#define PCODEC_EPI_GELU 1          /* gelu_erf(acc) */
#define PCODEC_EPI_ADD 2           /* acc + r1 */
#define PCODEC_EPI_ADD_GELU 3      /* gelu_erf(acc + r1)                     (ResidualUnit, layers.py:52-58) */
#define PCODEC_EPI_GATE 4          /* r2 * sigmoid(acc) + r1                 (Win_noShift_Attention, layers.py:69-75) */
#define PCODEC_EPI_GDN 5           /* r1 * rsqrt(acc)  with A = x*x          (gdn.py:53-63) */
#define PCODEC_EPI_IGDN 6          /* r1 * sqrt(acc) */
#define PCODEC_EPI_LRP 7           /* r1 + 0.5*tanh(acc) (+ r2 if given)     (CHProg_cnn.py:759-762, 840-843) */
#define PCODEC_EPI_CLAMP01 8       /* min(max(acc,0),1)                      (CHProg_cnn.py:909, 988) */
#define PCODEC_EPI_LEAKY 9         /* leaky_relu(acc, 0.01)                  (ResidualBlock, models/utils.py:59-87) */
#define PCODEC_EPI_LEAKY_ADD 10    /* leaky_relu(acc, 0.01) + r1             (ResidualBlock output: + identity / skip) */

#define PCODEC_FLAG_SQUARE_INPUT 1   /* A = x*x (GDN) */
#define PCODEC_FLAG_PIXEL_SHUFFLE2 2 /* output channel co -> pixel (2h + (co>>1&1), 2w + (co&1)), channel co>>2; applied
                                        after the epilogue (subpel_conv3x3, layers.py:20-24) */
#define PCODEC_FLAG_NO_F32_OUT 8     /* fp16 path: write only the output planes (out_hi/out_lo), not `out` */
#define PCODEC_FLAG_SQUARE_OUT_PLANES 16 /* fp16 path: the output PLANES hold (v * 2^-4)^2 instead of v — the x*x operand
                                        of the GDN that follows (gdn.py:50-63), so no separate squaring pass runs */
#define PCODEC_FLAG_SUBPIXEL_NCHW 4  /* the 4 sub-pixel phases of a stride-2 transposed convolution computed as ONE
                                        3x3-neighbourhood GEMM: conv channel co = (2*py + px) * C + c is stored to the
                                        NCHW image out[n][c][2h+py][2w+px], C = out_channels (cout >= 4*C, tcgen05 path
                                        only).  Used for the image layer of g_s (deconv(N, 3), CHProg_cnn.py:148-161) */

typedef struct {
  const float *ptr; /* NHWC base of this channel segment */
  int channels;     /* channels taken from this segment (multiple of 4) */
  int pixel_stride; /* floats between consecutive pixels */
} pcodec_segment;

/* Split-fp16 planes of an NHWC activation (the operand format of the fp16 tensor-core path, pcodec_split_planes):
 *   x ~= hi + lo * 2^-11,  hi = fp16(x),  lo = fp16((x - hi) * 2^11)      (22 significant bits, |x| < 65504)
 * two fp16 tensors with the geometry of the fp32 one; `pixel_stride` counts fp16 elements (a multiple of 8). */
typedef struct {
  const uint16_t *hi, *lo; /* fp16 bit patterns; hi == NULL: this segment has no planes */
  int pixel_stride;
} pcodec_planes;

typedef struct {
  /* input: virtual channel-concatenation of n_segments tensors sharing [batch, in_h, in_w] */
  pcodec_segment seg[PCODEC_MAX_SEGMENTS];
  int n_segments;
  int batch, in_h, in_w;
  /* taps: out(h,w) += in(h*in_step + dy[t], w*in_step + dx[t]) . W[t]; zero outside the input */
  int n_taps;
  int8_t dy[PCODEC_MAX_TAPS], dx[PCODEC_MAX_TAPS];
  int in_step;
  /* weights: fp32 [n_taps][cin_total][cout] (cout contiguous); bias fp32 [cout] or NULL */
  const float *weight;
  const float *bias;
  int cin_total, cout;
  /* output grid: M = batch*grid_h*grid_w GEMM rows; row (n,h,w) is stored at pixel
   * (h*out_step + out_off_y, w*out_step + out_off_x) of an [batch, out_h, out_w] NHWC tensor */
  int grid_h, grid_w, out_step, out_off_y, out_off_x, out_h, out_w;
  float *out;
  int out_pixel_stride;
  /* epilogue */
  int epilogue, flags;
  const float *r1; int r1_pixel_stride; /* indexed like `out` (same pixel, same channel) */
  const float *r2; int r2_pixel_stride;
  /* tensor-core path: handle from pcodec_conv_tc_prepare for these weights (NULL = fp32 SIMT kernel only) and the
   * number of TF32 products per MAC: 3 = split accumulation (fp32-class accuracy), 1 = plain TF32 */
  const void *tc_weights;
  int tc_split;
  /* fp16 tensor-core path (impl 3): the segments' split-fp16 planes (same order / channel windows as seg[]), optional
   * planes of the OUTPUT written by the epilogue (out_hi != NULL; with PCODEC_FLAG_NO_F32_OUT `out` is not written and
   * may be NULL), and a caller-owned scratch of PCODEC_CONV_PLAN_BYTES that pcodec_conv_plan() fills once per
   * descriptor (tiling + TMA tensor maps of the activation planes) and pcodec_conv_taps() reads at every launch. */
  pcodec_planes seg16[PCODEC_MAX_SEGMENTS];
  uint16_t *out_hi, *out_lo;
  int out_plane_stride;
  void *plan;
  /* fp16 path: residual operands given as split planes (used when r1 / r2 is NULL; value = hi + lo * 2^-11) */
  pcodec_planes r1_16, r2_16;
} pcodec_conv_desc;

#define PCODEC_CONV_PLAN_BYTES 2048

/* HOST descriptor, DEVICE tensors.  impl: 0 = auto (the fp16-split tcgen05 kernel when desc->plan is set, else the
 * 3xTF32 tcgen05 kernel when desc->tc_weights is set and the shape is supported, else SIMT), 1 = fp32 SIMT,
 * 2 = 3xTF32 tcgen05, 3 = fp16-split tcgen05 (2, 3: PCODEC_ERR_UNSUPPORTED if they cannot run this descriptor). */
int pcodec_conv_taps(const pcodec_conv_desc *desc, int impl, void *stream);

/* HOST. Build the launch plan of the fp16-split tcgen05 kernel into desc->plan (PCODEC_CONV_PLAN_BYTES, caller-owned):
 * every segment needs planes (seg16[i].hi/lo), desc->tc_weights a handle of pcodec_conv_tc_prepare.  Returns
 * PCODEC_ERR_UNSUPPORTED (and leaves the plan invalid) for shapes the kernel does not take. */
int pcodec_conv_plan(pcodec_conv_desc *desc);

/* fp32 NHWC window -> split-fp16 planes (see pcodec_planes).  src: [n_pixels] pixels of src_ps floats, `channels`
 * (multiple of 8) converted; hi/lo: planes with pixel stride dst_ps (fp16 elements, multiple of 8).
 * square != 0: the planes hold (x * 2^-4)^2 (GDN's x*x operand, gdn.py:50-63, pre-scaled for the fp16 range;
 * the GDN convolution's weights carry the 2^8). */
int pcodec_split_planes(const float *src, int src_ps, int64_t n_pixels, int channels, uint16_t *hi, uint16_t *lo,
                        int dst_ps, int square, void *stream);

/* Prepare weights for the tcgen05 kernel: takes the SIMT layout [n_taps][cin_total][cout] (device), builds the
 * K-major TF32 hi/lo copies [cout][n_taps*cin_total] and their TMA tensor maps; *handle_out is an opaque HOST handle
 * to put in pcodec_conv_desc.tc_weights.  PCODEC_ERR_UNSUPPORTED when cout is not a multiple of 16 (e.g. the
 * 3-channel output layer) or the driver lacks cuTensorMapEncodeTiled. */
int pcodec_conv_tc_prepare(const float *w_tap_major, int n_taps, int cin_total, int cout, void **handle_out,
                           void *stream);
void pcodec_conv_tc_release(void *handle);

/* ------------------------------------------------------------------------------------------
 * Shifted-window attention core (win_attention.py:84-115, 153-207)
 * ---------------------------------------------------------------------------------------- */
/* qkv: NHWC [B][H][W][3*C] (q | k | v, each head-major), out: NHWC [B][H][W][C] = softmax(q k^T * scale + relpos
 * + shift mask) v, already un-shifted back to image coordinates.  rel_bias: fp32 [heads][T][T] with T = ws*ws. */
int pcodec_window_attention(const float *qkv, int qkv_ps, float *out, int out_ps, const float *rel_bias, int batch,
                            int height, int width, int channels, int heads, int window, int shift, void *stream);

/* ------------------------------------------------------------------------------------------
 * Layout helpers
 * ---------------------------------------------------------------------------------------- */
/* NCHW [B][C][HW] -> NHWC (pixel stride dst_ps, channels beyond C up to c_pad are zero-filled). */
int pcodec_nchw_to_nhwc(const float *src, float *dst, int batch, int channels, int64_t hw, int dst_ps, int c_pad,
                        void *stream);
/* Patch extraction for the first analysis conv (3 input channels are too few for a 16-channel K chunk):
 * src NCHW [B][C][H][W] -> dst NHWC [B][OH][OW][k_pad] with dst[.., (ky*k+kx)*C + c] = src[n][c][oh*stride+ky-pad][ow*stride+kx-pad]
 * (zero outside the image and for entries >= k*k*C). */
int pcodec_im2col_nchw(const float *src, float *dst, int batch, int channels, int height, int width, int k, int stride,
                       int pad, int out_h, int out_w, int k_pad, void *stream);
/* NHWC (first C channels, pixel stride src_ps) -> NCHW. */
int pcodec_nhwc_to_nchw(const float *src, int src_ps, float *dst, int batch, int channels, int64_t hw, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PCODEC_B200_H_ */
