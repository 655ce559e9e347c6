/*
 * bihome_b200 -- C ABI of the B200-native biHomE hot path (libbihome_b200.so).
 *
 * The reference (NeurAI-Lab/biHomE) is pure Python: it has no FFI.  Its "plugin API" for this path
 * is the set of Python call sites listed next to every entry point below; each entry point replaces
 * the ATen/kornia op chain behind that call site with one hand-written sm_100a kernel.  The host side
 * (bihome_b200/functional.py, loaded through ctypes) keeps the reference's Python signatures.
 *
 * Conventions (all entry points):
 *   - return int: 0 = ok, <0 = bad argument (BH_E_*), >0 = cudaError_t of the launch;
 *   - stateless, re-entrant, stream ordered: nothing is allocated, nothing synchronises; the caller
 *     owns every buffer including workspaces; `stream` is a cudaStream_t (0 = legacy default);
 *   - every pointer is a DEVICE pointer to contiguous float32 unless stated otherwise, 16-byte aligned
 *     (anything from cudaMalloc / a fresh torch tensor is);
 *   - H is a row-major 3x3 homography stored as 9 floats per sample, H[8] is carried (1 for DLT-4);
 *   - "pixel coordinates" are pixel centres, i.e. F.grid_sample(align_corners=True) un-normalised.
 */
#ifndef BIHOME_B200_H_
#define BIHOME_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* bh_stream_t; /* == cudaStream_t */

#define BH_VERSION 100 /* 0.1.0 */

enum {
    BH_OK = 0,
    BH_E_NULL = -1,      /* required pointer is NULL */
    BH_E_SHAPE = -2,     /* non-positive or unsupported dimension */
    BH_E_ALIGN = -3,     /* pointer not 16-byte aligned */
    BH_E_WORKSPACE = -4, /* workspace too small */
    BH_E_UNSUPPORTED = -5
};

int bh_version(void);
/* human readable text for a return code of any entry point */
const char* bh_strerror(int code);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
unsigned long long bh_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1  4-point offsets -> homography (8x8 DLT, LU with partial pivoting, one 8-lane group per sample)
 *
 * replaces: src/data/utils.py:20-24  four_point_to_homography (torch branch)
 *             -> kornia.get_perspective_transform (16x cat/stack + torch.solve)
 *           src/data/utils.py:36-51  image_shape_to_corners  (corners == NULL => the canonical
 *             [[0,0],[W,0],[W,Hh],[0,Hh]] the reference builds there)
 *
 * corners [B,4,2] or NULL, delta [B,4,2] -> H [B,9] mapping corner_i -> corner_i + delta_i, H[8] = 1.
 * bwd: gH [B,9] (gH[8] ignored) -> gDelta [B,4,2] (d/d delta; also d/d dst) and, if non-NULL,
 *      gCorners [B,4,2] (total derivative w.r.t. corners, dst = corners + delta included).
 * ------------------------------------------------------------------------------------------- */
int bh_dlt4_fwd(const float* corners, const float* delta, float* H, int B, float W, float Hh, bh_stream_t stream);
int bh_dlt4_bwd(const float* corners, const float* delta, const float* H, const float* gH, float* gDelta,
                float* gCorners, int B, float W, float Hh, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2  homography warp: out[b,c,y,x] = bilinear(src[b,c], proj(H_b [x,y,1]^T)), zeros padding
 *
 * replaces: src/data/utils.py:54-59  warp_image(inverse=True) -> torch.inverse + kornia.warp_perspective
 *             (normalize_homography, 2 more inversions, create_meshgrid, transform_points, F.grid_sample)
 *           src/heads/PerceptualHead.py:339-340,380-382,401  warp(ones) masks and
 *           src/heads/PerceptualHead.py:447-459              their AvgPool2d(pool): mask_pooled is the
 *             analytic coverage mx(u)*my(v) averaged over pool x pool output pixels, no source read.
 *
 * src [B,C,Hs,Ws] (channels_last=0) or [B,Hs,Ws,C] (channels_last=1); out likewise with Ho,Wo.
 * src/out may both be NULL (masks only).  mask_pooled [B,Ho/pool,Wo/pool] or NULL (then pool is ignored).
 * bwd: gOut (layout of out, or NULL), gMaskPooled or NULL -> gH [B,9] (overwritten, gH[8] = d/dh33),
 *      and if gSrc != NULL the image gradient is ACCUMULATED into gSrc (caller zero-fills).
 *      workspace: bh_warp_bwd_workspace_bytes(...) bytes (may be 0 -> pass NULL).
 * ------------------------------------------------------------------------------------------- */
int bh_warp_fwd(const float* src, const float* H, float* out, float* mask_pooled, int B, int C, int Hs, int Ws,
                int Ho, int Wo, int pool, int channels_last, bh_stream_t stream);
size_t bh_warp_bwd_workspace_bytes(int B, int C, int Hs, int Ws, int Ho, int Wo, int channels_last);
int bh_warp_bwd(const float* src, const float* H, const float* gOut, const float* gMaskPooled, float* gH,
                float* gSrc, int B, int C, int Hs, int Ws, int Ho, int Wo, int pool, int channels_last,
                void* workspace, size_t workspace_bytes, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K3  bidirectional perceptual (biHomE) loss, fused forward + backward, one pass over the features
 *
 * replaces: src/heads/PerceptualHead.py:559-561 (l1 distances), :609-665 (double-line, margin 'inf',
 *           channel-agnostic: ln1, ln2, ln3, loss) and the ~110 autograd kernels behind them.
 *
 *   W1 = m1w*m2, W2 = m2w*m1 (pooled masks, NULL m1/m2 == ones)      S1 = sum W1, den1 = max(S1,1)
 *   D1 = sum_c|f1w-f2| - sum_c|f1-f2|,  D2 = sum_c|f2w-f1| - sum_c|f1-f2|
 *   loss_b = sum_hw W1*D1/den1 + sum_hw W2*D2/den2 + mu*||H12 H21 - I||_F^2
 *
 * features [B,C,h,w] (channels_last=0) or [B,h,w,C] (=1); masks [B,h,w]; H12,H21 [B,9].
 * outputs: loss [B]; parts [B,5] = {ln1, ln2, S1, S2, ln3} (ln3 without mu);
 *          g_f1w,g_f2w (feature layout), g_m1w,g_m2w [B,h,w], gH12,gH21 [B,9]: d loss_b / d(.) ;
 *          g_f1,g_f2: optional (NULL when the extractor inputs need no gradient, the shipped configs).
 * bh_bihome_rescale multiplies every gradient of sample b by gscale[b] (device vector) and is a no-op
 * launch when gscale[b] == 1 -- the autograd backward calls it with the upstream gradient.
 * One or two stream-ordered launches depending on layout and batch (cluster kernel; TMA-ring cluster kernel; persistent
 * TMA stream + per-sample finish -- DESIGN.md section 4); g_m1w / g_m2w double as scratch between them, every output is
 * final when the call's last launch completes.  The kernel is chosen from the arguments alone (bh_tune_set below can
 * force one for the microbenchmark).
 * ------------------------------------------------------------------------------------------- */
int bh_bihome_fwd_bwd(const float* f1, const float* f2, const float* f1w, const float* f2w, const float* m1,
                      const float* m2, const float* m1w, const float* m2w, const float* H12, const float* H21,
                      float mu, float* loss, float* parts, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2,
                      float* g_m1w, float* g_m2w, float* gH12, float* gH21, int B, int C, int h, int w,
                      int channels_last, bh_stream_t stream);
int bh_bihome_rescale(const float* gscale, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2, float* g_m1w,
                      float* g_m2w, float* gH12, float* gH21, int B, int C, int h, int w, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K3g  the masked triplet loss in every other variant of the reference, fused forward + backward
 *
 * replaces: src/heads/PerceptualHead.py:465-538 (one-line: l1 / cosine, numeric margin, MASK_CRD weights),
 *           :555-665 (double-line: l1 / l2 / cosine distances, margin 'inf' or numeric, channel-aware or
 *           channel-agnostic aggregation, mu * ||H12 H21 - I||^2), src/heads/TripletHead.py:78-153 (the same algebra on
 *           one-channel full-resolution maps with learned masks -- the shipped zhang-orig loss) and the ~40 element-wise
 *           ATen kernels (+ as many in autograd) behind each.
 *
 *   line 1: x = f1w, y = f2, W = a1*b2 (mask_crd: a1);  line 2 (lines == 2): x = f2w, y = f1, W = a2*b1
 *   distance 0: l1, per channel |x - y|;  1: l2, channel mean of (x - y)^2;  2: 1 - cosine_similarity (eps 1e-8)
 *   hinge 0: g = sum_c (d(x,y) - d(f1,f2))                           (TRIPLET_MARGIN 'inf')
 *         1: g = sum_c max(|x-y| - |f1-f2| + margin, 0)              (numeric margin, channel-aware; distance 0 only)
 *         2: g = max(d(x,y) - d(f1,f2) + margin, 0) on channel-aggregated distances (one-line; channel-agnostic)
 *   ln = scale * sum_hw W g / max(sum_hw W, 1);   loss_b = ln1 [+ ln2 + mu * ||H12 H21 - I||_F^2]
 *
 * features [B,C,h,w] (channels_last = 0) or [B,h,w,C] (= 1); masks [B,h,w] at the feature resolution, b2 / b1 NULL ==
 * ones; lines == 1 ignores f2w, a2, b1, H12, H21 and their gradients (may be NULL).
 * outputs: loss [B]; parts [B,5] = {ln1, ln2, S1, S2, ln3}; g_f1w, g_f2w (feature layout); g_f1, g_f2 optional (both or
 * none); g_a1, g_a2 [B,h,w] required (scratch between the launches, final on return); g_b2, g_b1 optional; gH12, gH21.
 * Three stream-ordered launches (mask sums, streaming pass, per-sample finish).  bh_triplet_rescale multiplies every
 * non-NULL gradient of sample b by gscale[b] (no-op launch for 1).
 * ------------------------------------------------------------------------------------------- */
int bh_triplet_fwd_bwd(const float* f1, const float* f2, const float* f1w, const float* f2w, const float* a1, const float* b2,
                       const float* a2, const float* b1, const float* H12, const float* H21, int lines, int distance,
                       int hinge, int mask_crd, float margin1, float margin2, float scale1, float scale2, float mu,
                       float* loss, float* parts, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2, float* g_a1,
                       float* g_b2, float* g_a2, float* g_b1, float* gH12, float* gH21, int B, int C, int h, int w,
                       int channels_last, bh_stream_t stream);
int bh_triplet_rescale(const float* gscale, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2, float* g_a1, float* g_b2,
                       float* g_a2, float* g_b1, float* gH12, float* gH21, int B, int C, int h, int w, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4  N-point normalised DLT (Zeng / DSAC branch), one warp per hypothesis
 *
 * replaces: src/heads/ransac_utils.py:58-72 (gather + kornia.find_homography_dlt: normalize_points,
 *           A^T A, torch.svd, denormalise, /H33) and src/heads/PerceptualHead.py:175-178 (corner
 *           projection kornia.transform_points(H, four_points) - four_points).
 *
 * Correspondences come either as point lists p1, p2 [B,N,2] (field == NULL) or, fused with the reference's
 * forward_map_field (PerceptualHead.py:125-146), as a perspective field [B,2,Hf,Wf] (p1, p2 == NULL, N = Hf*Wf,
 * Wf = field width): p1 = pixel grid (x, y), p2 = p1 + field[b,:,y,x].  choice [B,M] int64 = the sampled indices
 * into the N points (the torch.multinomial draw; NULL = all N points in order, M == N).  four [4,2] corner points.
 * fwd -> Hn [B,9] (H / (H33 + 1e-8)) and, if delta != NULL, delta [B,4,2] = proj(Hn, four) - four.
 * bwd: gHn [B,9] or NULL, gDelta [B,4,2] or NULL -> d/d p2 ACCUMULATED (atomics; caller zero-fills) into
 *      gP2 [B,N,2] (point mode) or gField [B,2,Hf,Wf] (field mode).  p1 carries no gradient (constants in the
 *      reference).  The solve runs in float64 internally; the adjoint recomputes the forward.
 * ------------------------------------------------------------------------------------------- */
int bh_dltn_fwd(const float* p1, const float* p2, const float* field, const int64_t* choice, const float* four,
                float* Hn, float* delta, int B, int N, int M, int Wf, bh_stream_t stream);
int bh_dltn_bwd(const float* p1, const float* p2, const float* field, const int64_t* choice, const float* four,
                const float* gHn, const float* gDelta, float* gP2, float* gField, int B, int N, int M, int Wf,
                bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K5  synthetic (PD)S-COCO pair generation on the GPU
 *
 * replaces: src/data/transforms.py:456-576,724-725 HomographyNetPrep.__call__ (photometric distortion
 *           :296-330, random patch position and corner offsets, cv2.getPerspectiveTransform,
 *           cv2.warpPerspective with 1/32-px taps, crop), :344-354 DictToGrayscale,
 *           :369-378 DictStandardize, :728-743 DictToTensor, train.py:308-309 (.float()).
 *
 * images: uint8 [n_img,Hi,Wi,3] RGB pool resident in HBM; index [B] int32 image per sample.
 * params: double [B, BH_PAIR_NPARAM] per-sample draws, layout (see oracle/pairgen.py pack_params):
 *   2 x {b_on, b_delta, contrast_first, c_on, c_alpha, s_on, s_alpha, h_on, h_delta, l_on, l_perm},
 *   pos_x, pos_y, delta[8].
 * bh_pairgen_draw fills params (and index) from a counter-based generator (seed, step, sample);
 * bh_pairgen_apply renders patch1, patch2 [B,1,P,P] (grayscale, standardised) and delta [B,4,2];
 * bh_pairgen_image renders image1 [B,1,Hi,Wi]: the whole first image through its photometric chain, grayscale,
 *   standardised -- the 'image_1' entry the reference's PhotometricHead reads (src/heads/PhotometricHead.py:24,
 *   config/s-coco/nguyen-orig-lr-5e-3.yaml LEARNING_KEYS).
 * ------------------------------------------------------------------------------------------- */
#define BH_PAIR_NPARAM 32
int bh_pairgen_draw(double* params, int32_t* index, int B, int n_img, int Hi, int Wi, int rho, int P,
                    float max_delta, uint64_t seed, uint64_t step, bh_stream_t stream);
int bh_pairgen_apply(const uint8_t* images, const int32_t* index, const double* params, float* patch1,
                     float* patch2, float* delta, int B, int n_img, int Hi, int Wi, int P, double mean, double std,
                     bh_stream_t stream);
int bh_pairgen_image(const uint8_t* images, const int32_t* index, const double* params, float* image1, int B, int n_img,
                     int Hi, int Wi, double mean, double std, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K6  perspective-field head of the Zeng backbone (per-pixel 16 -> 128 -> 2 network with folded BatchNorm)
 *
 * replaces: src/backbones/Rethinking.py:144-147,293  layer8 = Conv2d(16,128,1) -> BatchNorm2d(128) -> ReLU ->
 *           Conv2d(128,2,1): ~21 ATen/cuDNN passes over a [B,128,P,P] tensor per backbone pass.
 *
 * x [n_pix, cin] is the channels-last input (n_pix = B*HW pixels, cin = 16); the field `out` and its gradient `gOut`
 * are planar [B,2,HW] (what K4 reads).  W1 [hid,cin], b1 [hid] are the FOLDED first layer (BatchNorm's batch or
 * running statistics applied by the caller: bihome_b200/functional.py), W2 [2,hid], b2 [2].
 *   bh_fieldhead_supported   1 for the compiled geometry (cin = 16, hid = 128), else 0 -> the caller keeps ATen.
 *   bh_fieldhead_grid        CTAs the moments (what = 0) / backward (what = 1) launch uses == rows of `partials`.
 *   bh_fieldhead_moments     partials [grid, cin + cin*cin] double: per-CTA sums of x_i, then of x_i x_k (the full
 *                            symmetric matrix, row-major); the caller adds the rows (fixed order).
 *   bh_fieldhead_fwd         out = W2 relu(W1 x + b1) + b2.  fwd / bwd run on the tensor cores (mma.sync TF32,
 *                            csrc/fieldhead_mma.cu) when HW % 32 == 0, on scalar per-pixel kernels otherwise.  tf32 = 0:
 *                            operands split into a TF32 head and a float32 remainder, three products per step --
 *                            float32-faithful; tf32 = 1: operands rounded to TF32, one product -- the arithmetic of the
 *                            cuDNN convolutions this replaces under torch.backends.cudnn.allow_tf32 (the caller passes it)
 *   bh_fieldhead_bwd         gx [n_pix, cin] = d/dx (overwritten); partials [grid, hid*cin + hid + 2*hid + 2] float:
 *                            per-CTA { gW1 | gb1 | gW2 | gb2 }, the caller adds the rows.
 *   bh_fieldhead_affine      gx (+)= a + M x per pixel, a [cin], M [cin,cin]: the adjoint of the moments (how the
 *                            batch statistics feed back into the input); accumulate = 0 overwrites gx.
 * ------------------------------------------------------------------------------------------- */
int bh_fieldhead_supported(int cin, int hid);
int bh_fieldhead_grid(int what, long long n_pix);
int bh_fieldhead_moments(const float* x, double* partials, long long n_pix, int cin, bh_stream_t stream);
int bh_fieldhead_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* out,
                     int B, int HW, int cin, int hid, int tf32, bh_stream_t stream);
int bh_fieldhead_bwd(const float* x, const float* W1, const float* b1, const float* W2, const float* gOut, float* gx,
                     float* partials, int B, int HW, int cin, int hid, int tf32, bh_stream_t stream);
int bh_fieldhead_affine(const float* x, const float* a, const float* M, float* gx, long long n_pix, int cin,
                        int accumulate, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K7  ResNet stem: BatchNorm2d (batch statistics) -> ReLU -> MaxPool2d(3, stride 2, padding 1), channels-last, one stage
 *
 * replaces: src/heads/PerceptualHead.py:56-58  resnet.bn1 -> relu -> maxpool of the frozen extractor (AuxiliaryResnet, in
 *             train mode: batch statistics, four passes per step) and
 *           src/backbones/Rethinking.py:31-36,293  layer1's BatchNorm2d -> ReLU -> MaxPool2d (two passes per step):
 *           six ATen / cuDNN passes over a [B,64,P/2,P/2] tensor forward (pooling keeps int64 indices), as many backward.
 *
 * x [N,H,W,C] channels-last (the convolution's output), y [N,Ho,Wo,C] with Ho = (H-1)/2 + 1, Wo = (W-1)/2 + 1;
 * C a power of two, 4 <= C <= 1024 (bh_stem_supported).  gamma / beta [C] or NULL (affine = False).
 *   bh_stem_fwd   batch mean / biased variance per channel (float64 reduction, fixed order), running_mean / running_var
 *                 (NULL = not tracked) updated in place with `momentum` and the unbiased variance, then
 *                 y = maxpool(relu(x * scale + shift)).  stats [4,C] out: scale = gamma / std, shift = beta - mean * scale,
 *                 mean, 1/std (saved for the backward).  code [N,Ho,Wo,C] uint8 out or NULL (no backward wanted): window
 *                 position 3 ky + kx of each maximum, first maximum in row-major order as ATen's max_pool2d.
 *   bh_stem_bwd   gy [N,Ho,Wo,C] -> gx [N,H,W,C] (overwritten) through the pooling, the ReLU and BatchNorm's batch-statistics
 *                 backward; ggamma / gbeta [C] (overwritten) or NULL.
 * ws: bh_stem_workspace_bytes(C) bytes, 16-byte aligned, contents irrelevant between calls.
 * ------------------------------------------------------------------------------------------- */
int bh_stem_supported(int C);
size_t bh_stem_workspace_bytes(int C);
int bh_stem_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                float eps, float* y, uint8_t* code, float* stats, void* ws, size_t ws_bytes, int N, int H, int W, int C,
                bh_stream_t stream);
int bh_stem_bwd(const float* x, const float* stats, const uint8_t* code, const float* gy, float* gx, float* ggamma, float* gbeta,
                void* ws, size_t ws_bytes, int N, int H, int W, int C, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K7b  BatchNorm2d (batch statistics) [+ residual] -> ReLU, channels-last: the inner stages of the residual blocks
 *
 * replaces: src/backbones/utils.py (ResNet34ConvBlock / IdentityBlock / ResNet50* blocks: BatchNorm2d -> ReLU inside the
 *           upper branch, `relu(upper_branch(x) + lower_branch(x))` at the end) and torchvision's BasicBlock / Bottleneck of
 *           the frozen extractor (src/heads/PerceptualHead.py:60-68): BatchNorm, the residual add and the ReLU as separate
 *           ATen / cuDNN passes forward, threshold_backward + batch_norm_backward backward.
 *
 * x, residual (or NULL), y, gy, gx, gresidual: [n_pix, C] rows (= channels-last [N,C,H,W], n_pix = N*H*W); C as for K7.
 *   bh_bnact_fwd   statistics / running statistics / stats [4,C] as bh_stem_fwd; y = relu(x * scale + shift [+ residual]).
 *   bh_bnact_bwd   gresidual == NULL: the ReLU decision is recomputed from x (y may be NULL).  gresidual != NULL: y (the saved
 *                  output) decides, gresidual = gy where y > 0 (overwritten) is the gradient of the residual input.
 *                  gx (overwritten) through BatchNorm's batch-statistics backward; ggamma / gbeta [C] or NULL.
 * ws: bh_stem_workspace_bytes(C) bytes.
 * ------------------------------------------------------------------------------------------- */
int bh_bnact_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                 float* running_var, float momentum, float eps, float* y, float* stats, void* ws, size_t ws_bytes, long long n_pix,
                 int C, bh_stream_t stream);
int bh_bnact_bwd(const float* x, const float* y, const float* stats, const float* gy, float* gx, float* gresidual, float* ggamma,
                 float* gbeta, void* ws, size_t ws_bytes, long long n_pix, int C, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K7c  y = relu(BatchNorm_a(a) + BatchNorm_b(b)): the end of a residual block whose skip path has its own BatchNorm
 *
 * replaces: src/backbones/utils.py  `relu(upper_branch(x) + lower_branch(x))` of ResNet34ConvBlock (projection skip) and the
 *           up-sampling blocks: two cuDNN BatchNorms, an add and a ReLU forward; threshold_backward + two batch_norm_backward.
 *
 * a, b, y, gy, ga, gb: [n_pix, C] rows; stats_a / stats_b [4,C] as in K7.  Neither the skip's normalised tensor nor its
 * gradient is materialised.  ws: 2 * bh_stem_workspace_bytes(C) bytes.
 * ------------------------------------------------------------------------------------------- */
int bh_bnact2_fwd(const float* a, const float* b, const float* gamma_a, const float* beta_a, float* running_mean_a,
                  float* running_var_a, float momentum_a, float eps_a, const float* gamma_b, const float* beta_b,
                  float* running_mean_b, float* running_var_b, float momentum_b, float eps_b, float* y, float* stats_a,
                  float* stats_b, void* ws, size_t ws_bytes, long long n_pix, int C, bh_stream_t stream);
int bh_bnact2_bwd(const float* a, const float* b, const float* y, const float* stats_a, const float* stats_b, const float* gy,
                  float* ga, float* gb, float* ggamma_a, float* gbeta_a, float* ggamma_b, float* gbeta_b, void* ws, size_t ws_bytes,
                  long long n_pix, int C, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K8  per-channel bias of a transposed convolution, channels-last
 *
 * replaces: the bias of nn.ConvTranspose2d in the up-sampling blocks (src/backbones/utils.py:65-66, 139-140; default
 *           bias=True).  cuDNN's transposed convolution has no bias epilogue: ATen adds the bias in a strided broadcast pass
 *           (`output.add_(reshape_bias(...))`) and reduces `grad_output.sum((0, 2, 3))` with a generic reduction.
 *
 * y, gy: [n_pix, C] rows (= channels-last [N,C,H,W]); C as for K7 (bh_stem_supported).
 *   bh_bias_add    y[p, c] += bias[c], in place.
 *   bh_bias_grad   gbias[c] (overwritten) = sum over p of gy[p, c], float64 accumulation in a fixed order.
 * ws: bh_stem_workspace_bytes(C) bytes, 16-byte aligned.
 * ------------------------------------------------------------------------------------------- */
int bh_bias_add(float* y, const float* bias, long long n_pix, int C, bh_stream_t stream);
int bh_bias_grad(const float* gy, float* gbias, void* ws, size_t ws_bytes, long long n_pix, int C, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MACE: mean over B*4 corners of ||delta_gt - delta_hat||_2  (train.py:401-404, eval.py:133-134)
 * out: 1 float (overwritten).
 * ------------------------------------------------------------------------------------------- */
int bh_mace(const float* delta_gt, const float* delta_hat, float* out, int B, bh_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Microbenchmark-only switch (tools/microbench.py): forces one of several equivalent kernels behind an entry point so
 * that they can be timed against each other.  Process-global; every key defaults to 0 = "choose from the arguments",
 * which is the only state the product path (bihome_b200/, train.py, eval.py, bench.py) ever runs in.
 *   "warp_path"    0 tile kernels (TMA box per 32x32 tile), 1 the persistent ring kernels
 *   "warp_variant" tile kernels: bit 0 = four warps per tile (8 rows each) instead of two (16 rows each)
 *   "loss_variant" 0 auto, 1 ldg cluster kernel, 2 TMA cluster kernel, 3 persistent TMA stream
 *   "loss_cluster" 0 auto, 1 | 2 | 4 | 8 CTAs per cluster
 *   "fieldhead_variant" 0 auto (the tensor-core kernels when HW % 32 == 0), 1 the scalar per-pixel kernels
 * returns BH_E_UNSUPPORTED for an unknown key.
 * ------------------------------------------------------------------------------------------- */
int bh_tune_set(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* BIHOME_B200_H_ */
