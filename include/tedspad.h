/*
 * tedspad.h -- C ABI of the B200-native TeD-SPAD snippet feature-extraction hot path.
 *
 * The reference (UCF-CRCV/TeD-SPAD) is pure Python/PyTorch and has no FFI of its own: every
 * GPU kernel on this path is an ATen/cuDNN library call made from an nn.Module.forward body.
 * Each entry point below therefore cites the reference call site(s) whose library kernels it
 * replaces.  The host side that binds these symbols (ctypes) lives in ted-spad_b200/_lib.py and
 * mirrors aux_code/model_loaders.py; see INTEGRATION.md for the stub a reference maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success, non-zero on error; tedspad_last_error() then returns
 *     a thread-local, NUL-terminated description (the Python host raises RuntimeError with it);
 *   - functions are re-entrant per (device, stream); the library keeps no mutable global state
 *     except the lazily resolved driver entry point for tensor-map encoding;
 *   - activations are channels-last bf16: [N][D+2pd][H+2ph][W+2pw][ld] with a zero halo of
 *     (pd,ph,pw) pixels on every side (2-D tensors use D=1, pd=0).  A `tedspad_tensor` is a view
 *     of channels [coff, coff+C) of such a buffer, which is how skip/branch concatenation
 *     (unet_parts.py:67, i3d.py:149) is done without a copy.
 */
#ifndef TEDSPAD_H_
#define TEDSPAD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEDSPAD_ABI_VERSION 9

enum { TEDSPAD_ACT_NONE = 0, TEDSPAD_ACT_RELU = 1, TEDSPAD_ACT_SIGMOID = 2 };
/* A-operand feed of the implicit GEMM: AUTO picks FLAT when legal. */
enum { TEDSPAD_FEED_AUTO = 0, TEDSPAD_FEED_FLAT_TMA = 1, TEDSPAD_FEED_GATHER = 2 };
/* resampling filters of tedspad_preprocess */
enum { TEDSPAD_RESAMPLE_AA_FLOAT = 0, TEDSPAD_RESAMPLE_PIL_U8 = 1 };

typedef struct tedspad_tensor {
  void* ptr;                /* base of the allocation (halo included) */
  int32_t N, D, H, W, C;    /* logical extents of the view */
  int32_t pd, ph, pw;       /* zero halo on each side of D, H, W */
  int32_t ld;               /* channel stride of the allocation, elements */
  int32_t coff;             /* first channel of the view */
} tedspad_tensor;

/*
 * Convolution + folded BatchNorm + bias + optional residual add + activation, as an implicit
 * GEMM on tcgen05 tensor cores (fp32 accumulate in TMEM).
 * Replaces: nn.Conv2d+BatchNorm2d+ReLU of DoubleConv (aux_code/models/unet_parts.py:15-22),
 *           Unit3D.forward incl. TF-SAME padding (aux_code/models/i3d.py:89-120),
 *           Bottleneck.forward convs/BN/residual/ReLU (aux_code/models/large_i3d.py:61-84),
 *           I3Res50 conv1/bn1/relu (large_i3d.py:251-253), torchvision VideoResNet
 *           BasicStem/BasicBlock convs (video/resnet.py:87-121,173-181), and the Linear head of
 *           wrapper_r3d_18 (aux_code/model_loaders.py:204-213) as a 1x1x1 convolution.
 * Weights `w` are bf16 [Cout_pad][K_pad], K-major, K ordered (kd,kh,kw,cin) with cin padded to
 * Cin_pad = x.C; BN scale is folded into them in fp32 before rounding; `bias` is fp32 [Cout_pad].
 * Padding is given as FRONT pads; the back pad is implied by the output extents (this is what
 * expresses the asymmetric TF-SAME padding of i3d.py:82-109).
 * FLAT feed requires stride 1, x.C % 64 == 0, symmetric pads equal to the kernel half-width,
 * y with the same (N,D,H,W) and halo as x and halo >= pads; it keeps the zero halo of y intact.
 */
typedef struct tedspad_conv {
  tedspad_tensor x;         /* bf16 input view */
  tedspad_tensor y;         /* output view: bf16, or fp32 when y_fp32 != 0 */
  const void* w;
  const float* bias;
  const void* res;          /* optional bf16 residual with y's geometry/halo; NULL = none */
  int32_t res_ld, res_coff;
  int32_t Cout, Cout_pad, K_pad;
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t pd, ph, pw;       /* front pads */
  int32_t act;              /* TEDSPAD_ACT_* */
  int32_t y_fp32;
  int32_t feed;             /* TEDSPAD_FEED_* */
  int32_t n_tile;           /* UMMA N per tile; 0 = auto */
  int32_t max_ctas;         /* persistent grid cap; 0 = number of SMs */
  /* Optional extra destinations: convolutions that read the SAME input (the 1x1 branches b0 / b1a / b2a of an
   * InceptionModule, aux_code/models/i3d.py:144-149) run as ONE GEMM over their concatenated weight rows; output
   * columns [0, y.C) go to y, [y2_begin, y2_begin + y2.C) to y2 and [y3_begin, y3_begin + y3.C) to y3.  Begins are
   * multiples of 16 in increasing order; y2/y3 share y's (N,D,H,W) and halo; Cout = end of the last segment; no
   * residual, bf16 outputs.  ptr NULL = none (y3 needs y2). */
  tedspad_tensor y2, y3;
  int32_t y2_begin, y3_begin;
} tedspad_conv;

int tedspad_conv_forward(const tedspad_conv* p, void* stream);

/*
 * SLAB feed of the implicit GEMM: convolutions whose A operand is read in place from ONE spatial
 * slab of the input held in shared memory, every filter tap being a shifted UMMA descriptor over
 * that slab (no im2col copies, one TMA box per tile and K stage, weights resident in shared
 * memory for the whole persistent CTA).  An output tile is 16 rows x (8*tm) columns of one image.
 *   TEDSPAD_SLAB_3X3     Conv2d 3x3 stride 1 pad 1, Cin % 64 == 0, Cout in {16..256} with
 *                        9*Cin*Cout*2 bytes <= ~150 KB: the 64/128-channel DoubleConv layers of the
 *                        anonymizer at 224^2 / 112^2 (aux_code/models/unet_parts.py:15-22).  Taps outside
 *                        the tensor are zero-filled by TMA.  128-byte swizzled slab rows (one pixel
 *                        x 64 channels); tap (ky,kx) = descriptor start + (ky*slab_w + kx) rows.
 *   TEDSPAD_SLAB_STEM2D  Conv2d 3x3 stride 1 pad 1 over a Cin<=8 image stored with 8 channels per
 *                        pixel (16 bytes): the anonymizer's first convolution (unet_parts.py:15).
 *                        Un-swizzled K-major descriptors with OVERLAPPING K-adjacent core matrices
 *                        (LBO = 16 B = one pixel): row m of tap kx is pixel m + kx of the slab row.
 *   TEDSPAD_SLAB_STEM3D  Conv3d (kd,7,7) stride (sd,2,2) over a Cin<=4 clip stored with 4 channels
 *                        per pixel (8 bytes): Conv3d_1a_7x7 (aux_code/models/i3d.py:238-239),
 *                        I3Res50.conv1 (large_i3d.py:135), torchvision BasicStem (video/resnet.py:
 *                        173-181).  Same overlapped descriptors; one UMMA row step (16 B) = the
 *                        stride of 2 pixels, one K chunk = 2 pixels x 4 channels.
 *   TEDSPAD_SLAB_3X3_STREAM  Conv2d 3x3 / Conv3d 3x3x3 (or (1,3,3)) and, with ONE spatial tap per K stage, 1x1x1 /
 *                        (3,1,1) convolutions (Bottleneck conv1 / conv3 of large_i3d.py:47-58, the Inception pool
 *                        branch and Conv3d_2b of i3d.py), stride 1, same padding, Cin % 64
 *                        == 0, any Cout_pad <= 2048 (multiple of 32): same slab for the activations, but
 *                        the weights are too large to stay resident and stream through their own ring
 *                        of [n_tile x 64] blocks, one per filter tap (2-D TMA on the STANDARD packed
 *                        layout: pass it as `w_image`, with K_pad).  A K stage = (temporal tap, 64-channel
 *                        block).  The 128-channel DoubleConv layers (unet_parts.py:15-22), Conv3d_2c_3x3
 *                        and the Inception 3x3x3 branches (aux_code/models/i3d.py:244,132-136).  Taps
 *                        outside the tensor are zero-filled by TMA: no halo required.  1x1x1 convolutions may also be
 *                        STRIDED (strides <= 2, pads 0: the down-sample projections of the ResNet encoders): the TMA box
 *                        then walks the input with the convolution's stride.
 *   TEDSPAD_SLAB_3X3_PAIR  TEDSPAD_SLAB_3X3 executed by CTA PAIRS (clusters of two CTAs on the SMs of one TPC,
 *                        tcgen05.mma.cta_group::2, M = 256): each CTA owns one tile and HALF of the weight rows, so
 *                        the shared-memory operand reads that bound the N = 64 tensor-core rate at 67 % drop to
 *                        (128 + 32) rows per step = 80 %.  The 64-output-channel DoubleConv layers.  Needs
 *                        Cout_pad % 32 == 0, W > 8 and an even tile count; same fused epilogues as 3X3; the image
 *                        from tedspad_conv_slab_pack(kind = PAIR) holds the two halves back to back.
 *   TEDSPAD_SLAB_3X3_STREAM_PAIR  TEDSPAD_SLAB_3X3_STREAM executed by CTA pairs: every CTA streams HALF of each weight
 *                        block's rows.  One N tile (Cout_pad <= 256, multiple of 32), even tile count.  The N = 128
 *                        layers, whose single-CTA MMA reads shared memory at exactly its 128 B/clk limit.
 *   TEDSPAD_SLAB_STEM3D_PAIR  TEDSPAD_SLAB_STEM3D executed by CTA pairs (same reason as 3X3_PAIR: the stems have
 *                        N = 64): each CTA holds the weight image of HALF the output channels.  Even tile count;
 *                        tedspad_conv_slab_pack(kind = STEM3D_PAIR) writes the two halves back to back.
 *   TEDSPAD_SLAB_3X3_KX_PAIR  Conv2d 3x3 stride 1 pad 1 with Cout_pad 32 or 64 on CTA pairs, the three taps of a filter ROW
 *                        sharing one A-operand fetch: B stacks the weights of filter columns kx = 0, 1, 2 along N
 *                        (N = 3 * Cout_pad), a tile is 8 rows x 16 slab columns (14 output columns) and the epilogue adds
 *                        the three column blocks of neighbouring lanes.  3 instead of 9 shared-memory reads of every
 *                        slab pixel: for the layers whose tensor-core rate is bound by exactly those reads (N <= 64).
 *                        No fused pool / OutConv / residual / up-sampling.
 * Optional fused producer (single-CTA 3X3 kinds, 2-D): `up` = the low-resolution tensor of Up.forward; its x2 bilinear
 * (align_corners=True) up-sampling is computed by four producer warps straight into the shared-memory slab,
 * so the up-sampled half of torch.cat([x2, x1]) (unet_parts.py:67) is never written to or read from HBM.
 * Optional fused epilogues (TEDSPAD_SLAB_3X3 only): MaxPool2d(2) of the output written to `pool`
 * (unet_parts.py:33), and OutConv 1x1 (Cout->3) + sigmoid written as planar [N][3][H][W] images
 * (unet_parts.py:71-77, unet_model.py:36-37) in which case y.ptr may be NULL.
 * `w_image` holds the weights in the exact shared-memory image the kernel reads; build it with
 * tedspad_conv_slab_pack() from the standard packed layout of tedspad_conv.
 */
enum { TEDSPAD_SLAB_3X3 = 0, TEDSPAD_SLAB_STEM2D = 1, TEDSPAD_SLAB_STEM3D = 2, TEDSPAD_SLAB_3X3_STREAM = 3,
       TEDSPAD_SLAB_3X3_PAIR = 4, TEDSPAD_SLAB_3X3_STREAM_PAIR = 5, TEDSPAD_SLAB_STEM3D_PAIR = 6,
       TEDSPAD_SLAB_3X3_KX_PAIR = 7 };

typedef struct tedspad_conv_slab {
  tedspad_tensor x;         /* bf16 input view (see kinds above) */
  tedspad_tensor y;         /* bf16 output view, interior written only; ptr may be NULL with oc_w */
  const void* w_image;      /* device: weights, shared-memory image (tedspad_conv_slab_pack) */
  const float* bias;        /* device: fp32 [Cout_pad] */
  tedspad_tensor pool;      /* optional fused MaxPool2d(2) output view (ptr NULL = none) */
  tedspad_tensor up;        /* optional fused Up.forward input (unet_parts.py:50,57-67): the convolution sees
                               [x | upsample2x(up)] along the channels; ptr NULL = none */
  const float* oc_w;        /* optional fused OutConv: device fp32 [3][Cout]; NULL = none */
  const float* oc_b;        /* device fp32 [3] */
  void* oc_planes;          /* device bf16 [N][3][H][W] (may be NULL when oc_clip is given) */
  float* oc_frames;         /* optional device fp32 [N][3][H][W] */
  int32_t kind;             /* TEDSPAD_SLAB_* */
  int32_t Cout, Cout_pad;   /* Cout_pad = UMMA N (multiple of 16, <= 256) */
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t pd, ph, pw;       /* front pads */
  int32_t act;              /* TEDSPAD_ACT_* */
  int32_t tm;               /* 8-column groups per tile: 1 or 2; 0 = auto */
  int32_t max_ctas;         /* persistent grid cap; 0 = number of SMs */
  int32_t n_tile;           /* STREAM kind: UMMA N per tile (multiple of 32 dividing Cout_pad); 0 = auto */
  int32_t K_pad;            /* STREAM kind: row length of the standard packed weights */
  const void* res;          /* optional bf16 residual laid out like y (same pixel geometry and halo): y = act(conv + bias +
                               res), the Bottleneck tail of large_i3d.py:72-79.  Excludes pool / OutConv / up.  NULL = none */
  int32_t res_ld, res_coff; /* residual row length and first channel, elements */
  tedspad_tensor oc_clip;   /* optional fused OutConv destination: the ENCODER clip [B][T][H][W][>=3] (bf16), written
                               through the raw-reshape glue of dali_extraction.py:171-173 (plane 3t+c of clip b ->
                               channel (3t+c)/T, time (3t+c)%T); channels >= 3 are not touched.  With it oc_planes may be
                               NULL.  ptr NULL = none */
  int32_t oc_T;             /* frames per clip for oc_clip (x.N % oc_T == 0) */
  int32_t stack_rows;       /* 2-D 3X3 kinds over buffers with zero halo rows (x.ph >= 1): tile the rows of all N images
                               as one column of N*(H+2ph) rows (no per-image remainder).  0 = when it saves tensor
                               time, 1 = always (when legal), -1 = never */
} tedspad_conv_slab;

/* Everything the kernel derives from a tedspad_conv_slab: exposed so that the CPU test-suite can
 * replay the TMA box / UMMA descriptor arithmetic without a GPU (tests/_slabsim.py). */
#define TEDSPAD_SLAB_MAX_MMA 112
typedef struct tedspad_slab_plan {
  int32_t tm, n_tile, k_stages, n_mma, stages, tmem_cols;
  int32_t n_grp, nk, a_kstep, b_kstep;   /* a K stage = n_grp table groups x nk K=16 steps (byte steps per K step) */
  int32_t box[5];           /* TMA box, elements: {c, w, h, d, n} */
  int32_t tdim[5];          /* TMA tensor dims, elements */
  int64_t tstride[4];       /* TMA global strides, bytes (dims 1..4) */
  int64_t tbase_off;        /* byte offset of the TMA base from x.ptr */
  int32_t swizzle128;       /* slab written with the 128-byte swizzle */
  int32_t merged_cw;        /* stems: TMA dims are {W*C, H, D, N, 1}; the box x coordinate is scaled by 8 */
  int32_t slab_bytes, slab_stride, w_bytes, smem_bytes;
  int32_t a_layout, a_lbo, a_sbo, b_layout, b_lbo, b_sbo;   /* UMMA smem descriptor fields, bytes */
  int32_t half_a_off;       /* A byte offset of the second 8-column group */
  int32_t c_step, x_step, x_off, y_step, y_off, z_step, z_off, z_kstep;  /* slab origin per tile / K stage */
  int32_t tiles_x, tiles_y, tiles_z, total_tiles;
  int32_t b_stream, b_stages, b_stride, cb_n, cin, num_n_tiles, tab_per_stage;
  int32_t up_cb_first;      /* channel blocks >= this one are interpolated from `up` instead of loaded by TMA */
  int32_t stack_hp, stack_ph, stack_n;   /* stacked rows: padded image height, halo rows, batch (stack_hp 0 = off):
                               tile row g of row-tile ty is stacked row R = ph + 16*ty + g = image R / hp, row R % hp - ph */
  int32_t acc_stages;       /* depth of the TMEM accumulator ring (2..4) */
  int32_t pair;             /* 1: CTA pairs (cluster of 2, cta_group::2); w_bytes / tab B offsets are per CTA (N/2 rows) */
  uint32_t tab[2 * TEDSPAD_SLAB_MAX_MMA];   /* per (k_stage, group): {A byte offset in slab, B byte offset in image} */
} tedspad_slab_plan;

int tedspad_conv_slab_plan(const tedspad_conv_slab* p, tedspad_slab_plan* out);   /* host only, no GPU needed */
/* Standard packed weights (bf16 [Cout_pad][K_pad], K = (kd,kh,kw,cin_pad), see tedspad_conv) ->
 * shared-memory image for `kind` (device to device).  Returns the image size through *image_bytes
 * when image == NULL (no GPU needed for that query). */
int tedspad_conv_slab_pack(int32_t kind, const void* w_std, int32_t Cout_pad, int32_t K_pad, int32_t cin_pad,
                           int32_t kd, int32_t kh, int32_t kw, int32_t pw_front, void* image, int64_t* image_bytes,
                           void* stream);
int tedspad_conv_slab_forward(const tedspad_conv_slab* p, void* stream);

/* Planar bf16 images [B*T][3][H][W] (the anonymizer's output) -> encoder input view [B][T][H][W][>=3]
 * through the raw-reshape glue of feature_extraction/dali_extraction.py:171-173: plane p = 3*t + c
 * of clip b becomes encoder channel p / T at time p % T.  Channels >= 3 of y are written as zero. */
int tedspad_planes_to_clip(const void* planes, const tedspad_tensor* y, int32_t T, void* stream);

/*
 * Max pooling over (D,H,W) windows, channels-last.  Out-of-range taps contribute `0` when
 * zero_pad != 0 (MaxPool3dSamePadding pads with zeros: aux_code/models/i3d.py:21-45) and are
 * ignored otherwise.  Replaces nn.MaxPool2d(2) (unet_parts.py:33), MaxPool3dSamePadding
 * (i3d.py:13-45), I3Res50.maxpool1/maxpool2 (large_i3d.py:138-139).  Writes only the interior of
 * y (its halo must already be zero).
 */
int tedspad_maxpool(const tedspad_tensor* x, const tedspad_tensor* y, int32_t kd, int32_t kh, int32_t kw,
                    int32_t sd, int32_t sh, int32_t sw, int32_t pd, int32_t ph, int32_t pw, int32_t zero_pad,
                    void* stream);

/*
 * x2 bilinear up-sampling with align_corners=True written into a channel slice of the
 * concatenation buffer, centred with F.pad when sizes differ.
 * Replaces Up.forward's nn.Upsample + F.pad + torch.cat (aux_code/models/unet_parts.py:50,57-67).
 */
int tedspad_upsample2x(const tedspad_tensor* x, const tedspad_tensor* y, void* stream);

/*
 * x2 nearest-neighbour up-sampling written into a channel slice of a concatenation buffer (y.H == 2 x.H,
 * y.W == 2 x.W).  Replaces F.interpolate(scale_factor=2, mode="nearest") + torch.cat of smp 0.3.3's
 * unetplusplus DecoderBlock.forward (the arch='unet++' anonymizer, aux_code/model_loaders.py:18-30).
 */
int tedspad_upsample2x_nearest(const tedspad_tensor* x, const tedspad_tensor* y, void* stream);

/*
 * Channels-last anonymizer output [B*T][1][H][W][>=3] (bf16; the UNet++ segmentation head, activation=None) ->
 * encoder clip [B][T][H][W][4|8] through the raw-reshape glue of feature_extraction/dali_extraction.py:171-173
 * (plane 3t+c of clip b -> encoder channel (3t+c)/T, time (3t+c)%T); pad channels written as zero.  `frames_out`
 * (optional) receives the un-scattered fp32 frames [B*T][3][H][W], the fa_model return value.
 * s2d != 0: x is the space-to-depth form [B*T][1][H/2][W/2][>=12] the UNet++ tail is computed in (channel
 * (2*(h&1) + (w&1))*3 + c of low-resolution pixel (h/2, w/2) is colour c of pixel (h, w)).
 */
int tedspad_frames_to_clip(const tedspad_tensor* x, const tedspad_tensor* y, int32_t T, int32_t s2d, float* frames_out,
                           void* stream);

/*
 * OutConv (1x1, C->3) + sigmoid + the anonymizer->encoder raw-reshape glue: plane p = 3*t + c of
 * frame t lands at encoder channel p / T, time p % T.
 * Replaces OutConv.forward + nn.Sigmoid (unet_parts.py:71-77, unet_model.py:36-37) and the
 * view/reshape at feature_extraction/dali_extraction.py:171-173.
 * x: [B*T frames] view with C input channels; w: fp32 [3][C]; b: fp32 [3];
 * y: encoder input view [B][T][H][W] with >=3 channels (bf16).  When `frames_out` is non-NULL the
 * un-scattered fp32 anonymized frames [B*T][3][H][W] (NCHW, the fa_model return value) are also
 * written.
 */
int tedspad_outconv_sigmoid(const tedspad_tensor* x, const float* w, const float* b, const tedspad_tensor* y,
                            int32_t T, float* frames_out, void* stream);

/*
 * Mean over a (kd, H, W) window sliding over D with stride 1 -> fp32 features [N][D-kd+1][C].
 * kd <= 0 means kd = D (global pooling).
 * Replaces nn.AvgPool3d([2,7,7]) (i3d.py:293-294,340) and AdaptiveAvgPool3d(1) (large_i3d.py:262,
 * torchvision video/resnet.py:261).
 */
int tedspad_avgpool_features(const tedspad_tensor* x, int32_t kd, float* out, void* stream);

/* In-place row-wise L2 normalisation of an fp32 [rows][cols] matrix: x / max(||x||_2, eps).  Replaces
 * nn.functional.normalize(x, p=2, dim=1) of the embedding head mlp.forward (aux_code/model_loaders.py:249-253). */
int tedspad_l2_normalize_rows(float* x, int32_t rows, int32_t cols, float eps, void* stream);

/*
 * The consumer's view of a feature matrix, computed on the device: what anomaly_detection_mgfn's
 * Dataset.__getitem__ builds on the host from a loaded .npy (datasets/dataset.py:51-132, utils/utils.py:34-42).
 * feats: fp32 [T][ncrops][F] (a [T][F] matrix is ncrops = 1: dataset.py:70-71 expand_dims).
 *   train != 0: out fp32 [ncrops][seg][F+1]: every crop resampled to `seg` segments by process_feat (segment s = mean
 *               of snippet rows bounds[s] .. bounds[s+1]-1, or the single row bounds[s] when empty; `bounds` = device
 *               int32 [seg+1] = np.linspace(0, T, seg+1, dtype=int)), L2 magnitude of the segment row as column F.
 *   train == 0: out fp32 [T][ncrops][F+1]: the rows and their L2 magnitude (test mode, dataset.py:68-86).
 */
int tedspad_mgfn_rows(const float* feats, int32_t T, int32_t ncrops, int32_t F, const int32_t* bounds, int32_t seg,
                      int32_t train, float* out, void* stream);

/*
 * Crop + resize + normalise decoded uint8 frames into the anonymizer's bf16 input layout.
 * Replaces DALIDataloader.val_augmentations (feature_extraction/dali_extraction.py:38-50) with
 * TEDSPAD_RESAMPLE_AA_FLOAT (antialiased bilinear on /255 floats) and
 * shanghai_frames_dataset.augmentation (feature_extraction/shanghai_dl.py:27-40) with
 * TEDSPAD_RESAMPLE_PIL_U8 (Pillow's 8-bit two-pass bilinear, re-quantised, then /255).
 * frames: uint8 [F][Hs][Ws][3]; desc: int32 [n_out][4] = {src_frame (<0: all-zero image, DALI
 * pad_sequences), top, left, hflip} (hflip: crop taken from the horizontally flipped frame, as
 * torchvision ten_crop does); all images of one call share the crop size crop_h x crop_w;
 * y: bf16 view [n_out][1][Ho][Wo] with >= 3 channels (extra channels are written as zero).
 * `frames_f32` (optional) receives the fp32 NCHW [n_out][3][Ho][Wo] result for parity tests.
 */
int tedspad_preprocess(const uint8_t* frames, int32_t F, int32_t Hs, int32_t Ws, const int32_t* desc,
                       int32_t n_out, int32_t crop_h, int32_t crop_w, const tedspad_tensor* y, int32_t resample,
                       float* frames_f32, void* stream);

/* fp32 NCHW / NCDHW tensor -> bf16 channels-last view (module-boundary adapter used when a caller
 * hands the nn.Module an fp32 torch tensor, e.g. dali_extraction.py:173,176).  x has Cx channels;
 * channels [Cx, y.C) of the view are written as zero. */
int tedspad_nchw_to_cl(const float* x, int32_t Cx, const tedspad_tensor* y, void* stream);

int tedspad_abi_version(void);
/* Struct layout as THIS library was compiled: out[0..9] = sizeof(tedspad_tensor), sizeof(tedspad_conv),
 * sizeof(tedspad_conv_slab), sizeof(tedspad_slab_plan), offsetof(tedspad_conv, y2), offsetof(tedspad_conv_slab, kind),
 * offsetof(tedspad_conv_slab, res), offsetof(tedspad_conv_slab, oc_clip), offsetof(tedspad_conv_slab, stack_rows),
 * offsetof(tedspad_slab_plan, tab).  A binding checks its own struct declarations against these (returns the count). */
int tedspad_abi_layout(int32_t* out, int32_t n);
int tedspad_num_sms(void);
const char* tedspad_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* TEDSPAD_H_ */
