/*
 * ojdf.h -- C ABI of the B200-native per-frame TSDF-fusion hot path.
 *
 * The reference (suryanshkumar/online-joint-depthfusion-and-semantic @ a4f9e19) has no
 * FFI layer: its hot path is ~1500 ATen calls per frame issued from
 *   modules/extractor.py:24-79   Extractor.forward
 *   modules/integrator.py:15-126 Integrator.forward
 *   modules/pipeline.py:137-171  Pipeline._prepare_volume_update
 * Each entry point below replaces one of those call sites with hand-written sm_100a
 * kernels.  The host-side mirror of the reference classes (the package's modules/*.py)
 * binds these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every `*_dev` pointer is a DEVICE pointer owned by the caller (torch tensors);
 *     `*_host` pointers are small HOST arrays read synchronously during the call;
 *   - fp16 volumes are raw IEEE binary16 arrays, dense row-major (X,Y,Z), z fastest,
 *     linear index (x*Y + y)*Z + z           (modules/integrator.py:57);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the
 *     call returns without synchronising;
 *   - nothing is allocated: scratch comes from a caller-provided workspace;
 *   - return value 0 = success, OJDF_ERR_* (< 0) = argument error, > 0 = cudaError_t.
 *   - no torch types, no C++ types, no exceptions cross this boundary.
 */
#ifndef OJDF_H
#define OJDF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OJDF_VERSION 100

#define OJDF_ERR_BADARG (-1)      /* null pointer / non-positive size / P even or > 33 */
#define OJDF_ERR_WORKSPACE (-2)   /* workspace missing or too small */
#define OJDF_ERR_TOOLARGE (-3)    /* grid or entry count does not fit 32-bit keys */

int ojdf_version(void);
/* Human-readable text for a return code of this library. */
const char *ojdf_error_string(int code);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t ojdf_launch_count(void);

/* ---- a5: Extractor.compute_coordinates (modules/extractor.py:82-120) -----------------
 * depth (h,w) f32 -> world (h*w,3) f32.  Kinv = intrinsics.float().inverse() (3x3 row
 * major, computed by the host exactly as the reference does), E = cam->world rows 0..2
 * (3x4 row major).  Each output is the f32 FMA chain t=fl(a0*b0); t=fma(a1,b1,t); ...
 * (what the reference's BLAS does at the benchmark shapes). */
int ojdf_unproject(const float *depth_dev, int h, int w, const float *Kinv_host, const float *E_host,
                   float *world_dev, void *stream);

/* ---- a4/a6/a7/a8: Extractor.forward (modules/extractor.py:24-79, 309-345, 533-681) ---
 * One pass per frame: per-ray voxel-space sample points (f64, P = n_points samples at
 * 1-voxel spacing, index 0 nearest the camera), 8 trilinear corners per sample, gather of
 * the fp16 TSDF and weight volumes (out-of-grid corners read -0.1 / 0,
 * modules/extractor.py:663-664), f64 weighted sum in the reference's order -> f32.
 *
 *   depth_dev     (h,w) f32, or NULL when world_in_dev is given
 *   world_in_dev  optional (h*w,3) f32: use these world points instead of unprojecting
 *                 (parity tests feed the oracle's `pcl` here)
 *   eye is E[:,3] (modules/extractor.py:57)
 *   out_vals_dev / out_wts_dev   (N,P) f32   = values['fusion_values' / 'fusion_weights']
 *   out_world_dev   optional (N,3) f32       = values['pcl']
 *   out_ray_dev     REQUIRED (N,6) f64       = per-ray (centre_voxel xyz, unit direction xyz);
 *                                              written by the ray-setup kernel, read by the gather
 *                                              kernel, and the compact form ojdf_integrate() consumes
 *   out_points_dev  optional (N,P,3) f64     = values['points']
 *   out_idx_dev     optional (N,P,8,3) i64   = values['indices']
 *   out_w_dev       optional (N,P,8) f64     = values['weights']
 */
int ojdf_extract(const float *depth_dev, const float *world_in_dev, int h, int w,
                 const float *Kinv_host, const float *E_host, const double *origin_host, double resolution,
                 const void *tsdf_dev, const void *wvol_dev, int X, int Y, int Z, int P,
                 float *out_vals_dev, float *out_wts_dev, float *out_world_dev, double *out_ray_dev,
                 double *out_points_dev, int64_t *out_idx_dev, double *out_w_dev, void *stream);

/* The same step in two halves (the per-ray records depend only on depth and pose, so the pipeline computes them first
 * and plans the integration with them while the networks run):
 *   ojdf_rays   -- a5 + the record part of a6: world points (optional) and the (N,6) f64 per-ray records;
 *   ojdf_gather -- a6-a8 from existing records, one kernel; optionally also writes FusionNet's pixel-major input
 *                  rows [values(P) | weights(P) | last] (modules/pipeline.py:74-102) into pack_a_dev / pack_b_dev
 *                  ((N, pack_stride) f32; last_*_dev (N) f32 = depth frame / normalised label frame; pack_b optional),
 *                  which replaces a separate packing pass. */
int ojdf_rays(const float *depth_dev, const float *world_in_dev, int h, int w, const float *Kinv_host,
              const float *E_host, const double *origin_host, double resolution, float *out_world_dev,
              double *out_ray_dev, void *stream);
int ojdf_gather(const double *ray_dev, int h, int w, const void *tsdf_dev, const void *wvol_dev, int X, int Y, int Z,
                int P, float *out_vals_dev, float *out_wts_dev, double *out_points_dev, int64_t *out_idx_dev,
                double *out_w_dev, float *pack_a_dev, float *pack_b_dev, const float *last_a_dev,
                const float *last_b_dev, int pack_stride, void *stream);

/* Bytes of scratch ojdf_integrate*() needs for up to `max_entries` (ray,sample,corner)
 * entries per call: N*tail*8 for the frame form, M1*8 for the updates form. */
size_t ojdf_integrate_workspace_bytes(int64_t max_entries);
/* Put a fresh (or dirty, after a failed call) workspace into its idle state.  Must be
 * enqueued once before the first ojdf_integrate*() on that workspace. */
int ojdf_integrate_workspace_init(void *workspace_dev, size_t workspace_bytes, void *stream);
/* Introspection for tests: the first N bytes of a workspace of this size hold the hash table and
 * control words, and are all zero whenever no ojdf_integrate*() call is in flight. */
size_t ojdf_integrate_workspace_idle_bytes(size_t workspace_bytes);

/* ---- a13/a14/a15: _prepare_volume_update + Integrator.forward, whole-frame form ------
 * (modules/pipeline.py:137-171, modules/integrator.py:29-124).  Consumes the extractor's
 * per-ray record instead of the materialised indices/weights.
 *   ray_dev         (N,6) f64 from ojdf_extract(out_ray_dev)
 *   filt_depth_dev  (N) f32: masked depth; rays with 0 are skipped (pipeline.py:143-146)
 *   est_dev         (N,P) f32 network output; samples 0..tail-1 are integrated after
 *                   clamping to +-clamp_value (pipeline.py:156-159)
 *   pix_ids_dev (N) u8 / pix_scores_dev (N) f32: per-pixel label and score, broadcast
 *                   over the ray's samples (pipeline.py:161-169); used when do_semantics
 *   volumes are updated IN PLACE.  Per voxel the contributions are summed in fp32 in
 *   ascending (ray, sample, corner) order -- the reference's CPU index_add_ order -- and
 *   the semantic "last writer" is the highest entry, so the
 *   result is deterministic and bit-identical to the single-threaded reference.
 */
int ojdf_integrate(const double *ray_dev, const float *filt_depth_dev, const float *est_dev, int64_t N,
                   int P, int tail, float clamp_value,
                   void *tsdf_dev, void *wvol_dev, int X, int Y, int Z,
                   const uint8_t *pix_ids_dev, const float *pix_scores_dev,
                   uint8_t *ids_vol_dev, void *scores_vol_dev, int do_semantics,
                   void *workspace_dev, size_t workspace_bytes, void *stream);

/* The same step in two halves, so that the part that does not need the network can leave the critical path:
 *   ojdf_integrate_plan  -- needs only geometry (per-ray records + masked depth): groups the frame's entries by voxel
 *                           into the workspace.  Independent of FusionNet / AdapNet++: the pipeline issues it on a side
 *                           stream right after the per-ray records exist, while the networks run;
 *   ojdf_integrate_apply -- the network output, labels and scores are gathered per entry, each voxel's entries are
 *                           visited in ascending order and the volumes are updated; the workspace returns to idle.
 * ojdf_integrate == plan followed by apply on one stream.  Between the two calls the workspace must not be used by
 * another frame; arguments N, P, tail, X, Y, Z must be the same in both. */
int ojdf_integrate_plan(const double *ray_dev, const float *filt_depth_dev, int64_t N, int P, int tail,
                        int X, int Y, int Z, void *workspace_dev, size_t workspace_bytes, void *stream);
int ojdf_integrate_apply(const float *est_dev, int64_t N, int P, int tail, float clamp_value,
                         void *tsdf_dev, void *wvol_dev, int X, int Y, int Z,
                         const uint8_t *pix_ids_dev, const float *pix_scores_dev,
                         uint8_t *ids_vol_dev, void *scores_vol_dev, int do_semantics,
                         void *workspace_dev, size_t workspace_bytes, void *stream);

/* ---- a14/a15 in the reference's own `updates` form (modules/integrator.py:15-126) ------
 *   values_dev (M1) f32 already clamped, idx_dev (M1,8,3) i64, w_dev (M1,8) f64,
 *   ids_dev (M1) u8, scores_dev (M1) f32; entry e = m*8 + c.  Same ordering guarantees. */
int ojdf_integrate_updates(const float *values_dev, const int64_t *idx_dev, const double *w_dev, int64_t M1,
                           void *tsdf_dev, void *wvol_dev, int X, int Y, int Z,
                           const uint8_t *ids_dev, const float *scores_dev,
                           uint8_t *ids_vol_dev, void *scores_vol_dev, int do_semantics,
                           void *workspace_dev, size_t workspace_bytes, void *stream);

/* ---- a10 / a17: FusionNet and AdapNet++ layers (modules/model.py:4-283, modules/adapnet.py:12-415), pixel-major (NHWC)
 * fp32 activations.  One "tap GEMM" covers every convolution of both networks:
 *   out[p, coff+co] = out_mul * act(scale[co] * sum_tap sum_ci in[p + tap*dilation, ci] * W[tap,ci,co] + shift[co] (+ residual))
 * taps = 1 (1x1) or 9 (3x3, zero padding = dilation); act: 0 none, 1 ReLU, 2 LeakyReLU(slope), 3 tanh, 4 sigmoid,
 * 5 = sigmoid(v) * residual (the SSMA gate, modules/adapnet.py:352; residual_dev required).
 * scale/shift carry the conv bias and the inference BatchNorm.  in: (H*W, in_stride) floats, in_stride % 4 == 0, channels
 * [0,cin) are read; out: (H*W, out_stride) floats.  Up to 8 independent convolutions of identical shape (cin, cout, taps,
 * H, W, activation) go in ONE launch -- the two FusionNet heads, the four VortexPooling branches (each with its own
 * dilation), the two AdapNet++ encoders.  `problems_host` is a host array read during the call. */
typedef struct ojdf_conv_problem {
    const float *in_dev;
    const float *weights_dev;
    const float *scale_dev;
    const float *shift_dev;
    float *out_dev;
    const float *residual_dev;      /* optional (H*W, residual_stride): added before the activation */
    int in_stride, out_stride, out_coffset, dilation, residual_stride;
    int in_step;                    /* tensor-core kernel only: 0/1 = dense, 2 = read every other input pixel of every
                                     * other row (stride-2 1x1 convolutions, ResNet down-sampling); H, W stay OUTPUT sizes */
    int in_width;                   /* pixels per input row when in_step > 1 (the input image is in_width wide) */
    int out_step;                   /* tensor-core kernel only: 0/1 = dense, k = output pixel (y, x) of this problem is pixel
                                     * (k*y, k*x) of an image that is out_width pixels wide, counted from out_dev (one phase of a
                                     * transposed convolution; needs the TMA-store epilogue) */
    int out_width;
    int tap_mask;                   /* 3x3 only: bit t set = tap t (ky*3+kx) contributes; 0 = all nine.  Taps whose shifted
                                     * window lies entirely outside the image are dropped automatically */
} ojdf_conv_problem;
/* ---- the tap GEMM on the tensor cores (csrc/ojdf_conv_tc.cu, csrc/ojdf_conv_ss.cu),
 * computed by tcgen05.mma (kind::tf32, fp32 TMEM accumulators) with a 3xTF32 split-precision product
 * (x = hi + lo; hi*hi + lo*hi + hi*lo), i.e. fp32-grade results (~1e-6 relative).  Replaces the
 * reference's per-layer cuDNN convolution + BatchNorm + activation launches of modules/model.py:4-283 and
 * modules/adapnet.py:12-415.  Persistent CTAs (one per SM); one TMA halo box per 32-channel K chunk serves
 * all 9 taps; the activation operand is split into hi/lo in registers and kept in tensor memory; weight
 * stages are shared by up to 4 128-pixel M-tiles; double-buffered accumulators; TMA-store epilogue.
 *   - weights must be packed by ojdf_conv_tc_pack_weights (host side) with the same npad_req as the launch;
 *     `weights_dev` of each problem points at the device copy of the packed image;
 *   - in_dev 16-byte aligned, in_stride % 4 == 0; input channels >= cin are never read (TMA bounds), so a
 *     dense-block buffer can be consumed while it grows;
 *   - npad_req: 0 = default channel grouping (<= 128 output channels per CTA), or 16..128 (multiple of 16) to
 *     force narrower groups (more CTAs on small feature maps);
 *   - flags: 1 = the caller owns the pad channels [coff+cout, coff+round_up(cout,4)) of the output rows (they
 *     receive zeros or are left untouched; lets a width that is not a multiple of 4 use the TMA-store epilogue, which also needs
 *     out_coffset % 4 == 0, out_stride % 4 == 0 and a 16-byte aligned out_dev -- otherwise a coalesced plain
 *     store path is taken); experiments: 2 = one M-tile per group, 4 = never use halo boxes, 8 = per-thread
 *     stores, 16..2048 and bits 12-15 = timing probes (see the source), 4096 = never split the K loop;
 *   - scratch_dev (optional, scratch_bytes): lets the K loop be split over several CTAs when the feature map
 *     is too small to fill the SMs; partial sums are reduced in a fixed order by a second kernel (deterministic).
 *     flags 16384 (experimental, slower today): the CTA that delivers the last K slice of a tile reduces it inside the
 *     kernel; the first 16 KB of the scratch are then per-tile slice counters that must be ZERO before the first launch
 *     (every launch leaves them zero). */
int ojdf_conv_tc_layout(int cout, int npad_req, int *npad, int *groups);
size_t ojdf_conv_tc_weight_floats(int cin, int cout, int taps, int npad_req);
/* w_host: (cout, cin, taps) fp32 as in nn.Conv2d.weight (tap = ky*3 + kx); packed_host:
 * ojdf_conv_tc_weight_floats() floats. */
int ojdf_conv_tc_pack_weights(const float *w_host, int cin, int cout, int taps, int npad_req, float *packed_host);
int ojdf_conv_tc_batched(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H, int W,
                         int taps, int act, float slope, float out_mul, int npad_req, int flags, float *scratch_dev,
                         size_t scratch_bytes, void *stream);
/* ---- a10: chains of 1x1 convolutions kept on chip (csrc/ojdf_conv_chain.cu) --------------------------------------
 * FusionNet's Pred stack (modules/model.py:24-52: eleven 1x1 conv + BatchNorm + LeakyReLU layers, 114 -> ... -> 9) and the
 * end of every VortexPooling block (modules/model.py:131-141,157-159: four 19 -> 114 conv + BN + ReLU whose concatenation
 * feeds the 456 -> 114 `final` conv) as ONE launch each: a 128-pixel tile walks all steps inside a CTA, the activated
 * output of a step is split into tf32 hi / lo and written back to tensor memory as the A operand of the next step, so
 * activations never leave the SM between layers (same 3xTF32 numerics as ojdf_conv_tc_batched).
 *   step = D[acc] (+)= A . W^T  (cout <= 128), A = input `src` (a pixel-major global buffer, any cin) or, with src = -1,
 *   the activated output of the previous "epi = 1" step (then cin must equal that step's cout);
 *   fresh = 1 overwrites accumulator `acc` (0 or 1), fresh = 0 adds to it (sum over concatenated inputs);
 *   epi = 0: nothing (a later step adds to the same accumulator), 1: act(scale * D + shift) -> next step's A,
 *   2: out = out_mul * act(scale * D + shift) -> out_dev (the last step, and only the last step).
 * Up to 2 problems of identical shape per launch (the two FusionNet heads): pointer arrays are indexed by problem.
 * weights: ojdf_conv_tc_pack_weights(taps = 1, npad_req = 0) images.  flags: 1 = the caller owns the pad channels of
 * the output rows (TMA-store epilogue for widths that are not a multiple of 4), 8 = per-thread stores, 64 = 1xTF32. */
typedef struct ojdf_chain_input {
    const float *in_dev[2];         /* (H*W, in_stride) f32 per problem, 16-byte aligned */
    int in_stride, cin;             /* in_stride % 4 == 0; channels [0, cin) are read */
} ojdf_chain_input;
typedef struct ojdf_chain_step {
    const float *weights_dev[2];
    const float *scale_dev[2];
    const float *shift_dev[2];      /* scale / shift: cout floats, read by epi = 1 / 2 steps */
    int src, cin, cout, acc, fresh, epi, act;
    float slope;
} ojdf_chain_step;
int ojdf_conv_chain(const ojdf_chain_input *inputs_host, int n_inputs, const ojdf_chain_step *steps_host, int n_steps,
                    int n_problems, int H, int W, float *const *out_dev_host, const int *out_coffset_host, int out_stride,
                    float out_mul, int flags, void *stream);
/* ---- a17 / a3: the two ends of AdapNet++ that are not tap GEMMs (csrc/ojdf_adapnet_aux.cu) ------
 * ojdf_adapnet_stem: ResNet-50 conv1 7x7 / 2 / pad 3 (3 -> 64) + BatchNorm(eval, folded into scale/shift) + ReLU +
 * max-pool 3x3 / 2 / pad 1 (modules/adapnet.py:101,134-137) in one kernel: in_dev (3,H,W) f32 NCHW, weights_dev
 * (147, 64) f32 = [ci*49 + ky*7 + kx][co] (16-byte aligned), out_dev pixel-major (H/4 * W/4, out_stride) at channel
 * out_coffset.  Up to 2 problems (the RGB and the depth encoder) per launch; H, W multiples of 4 (AdapNet++ needs 16). */
typedef struct ojdf_stem_problem {
    const float *in_dev;
    const float *weights_dev;
    const float *scale_dev;
    const float *shift_dev;
    float *out_dev;
    int out_stride, out_coffset;
} ojdf_stem_problem;
int ojdf_adapnet_stem(const ojdf_stem_problem *problems_host, int n_problems, int H, int W, void *stream);
/* softmax over the C class logits of every pixel (pixel-major, `stride` floats per pixel), its maximum -> scores_dev
 * (npix) f32, its arg-max -> ids_dev (npix) u8, and (optional) the normalised label frame (1 + id) / n_classes ->
 * sem_frame_dev (npix) f32: modules/pipeline.py:57-58,96,184 in one pass. */
int ojdf_softmax_max(const float *logits_dev, int stride, int C, int npix, int n_classes, float *scores_dev,
                     uint8_t *ids_dev, float *sem_frame_dev, void *stream);
/* nn.AvgPool2d(3, stride 1, padding 1) of VortexPooling (modules/model.py:114-116), C % 4 == 0. */
int ojdf_avgpool3_nhwc(const float *in_dev, int in_stride, int H, int W, int C, float *out_dev, int out_stride,
                       void *stream);
/* Up to 8 equally shaped 3x3 average pools (stride 1, zero padding counted in the divisor) in one launch; a problem
 * with scale_dev != NULL also applies out = [relu](scale[c] * pool + shift[c]) (scale/shift: C floats, 16-byte
 * aligned).  Used for VortexPooling's cascaded pools, which commute with the branch's first 1x1 convolution
 * (modules/model.py:114-135): the engine pools W.x (19 channels) instead of x (114 channels). */
typedef struct ojdf_pool_problem {
    const float *in_dev;
    float *out_dev;
    const float *scale_dev;
    const float *shift_dev;
    int in_stride, out_stride;
    int identity;                   /* 1: no pooling, only the scale / shift (/ ReLU) epilogue (VortexPooling's branch 0, whose 1x1
                                     * product comes out of the same merged launch as the pooled branches') */
} ojdf_pool_problem;
int ojdf_avgpool3_batched(const ojdf_pool_problem *problems_host, int n_problems, int H, int W, int C, int relu, void *stream);
/* VortexPooling global branch (modules/model.py:107-112) folded into the bias of the `final` 1x1 conv:
 * shift_out[co] = f_shift[co] + f_scale[co] * sum_c wf1[co,c] * (g_scale[c]*(wg[c,:].mean_pixels(in)) + g_shift[c]).
 * partial_dev: scratch of partial_blocks*C floats. */
int ojdf_vortex_bias(const float *in_dev, int in_stride, int npix, int C, const float *wg_dev,
                     const float *g_scale_dev, const float *g_shift_dev, int Cg, const float *wf1_dev,
                     const float *f_scale_dev, const float *f_shift_dev, int Cout, float *partial_dev,
                     int partial_blocks, float *shift_out_dev, void *stream);
/* Generalisation used by AdapNet++'s eASPP branch 5 (modules/adapnet.py:201-205): C <= 2048 and an
 * optional ReLU on the pooled branch (v_relu).  Wide branches (C * Cg >= 65536) spread the first matrix-vector product
 * over many blocks and keep one C-float block of the scratch for its result. */
int ojdf_gap_bias(const float *in_dev, int in_stride, int npix, int C, const float *wg_dev,
                  const float *g_scale_dev, const float *g_shift_dev, int Cg, int v_relu, const float *wf1_dev,
                  const float *f_scale_dev, const float *f_shift_dev, int Cout, float *partial_dev,
                  int partial_blocks, float *shift_out_dev, void *stream);
/* Decoder skip join of AdapNet++ (modules/adapnet.py:305-315): gate[c] = relu(b[c] + W[c,:] . mean_pixels(x[:, :C])) from the
 * C (<= 1024, % 4 == 0) decoder features x (npix, x_stride), out[p, c] = skip[p, c] * gate[c] for the Cg (<= 32) skip
 * channels -- torch.mean + conv + relu + multiply + the copy into the concatenated buffer in two launches.
 * w_dev: (Cg, C) f32, b_dev: (Cg); partial_dev: scratch of partial_blocks*C floats. */
int ojdf_adapnet_skip_join(const float *x_dev, int x_stride, int C, int npix, const float *w_dev, const float *b_dev,
                           int Cg, const float *skip_dev, int skip_stride, float *out_dev, int out_stride,
                           float *partial_dev, int partial_blocks, void *stream);
/* (C, npix) fp32 <-> pixel-major (npix, stride) with a channel offset: hand-over between NCHW tensors
 * and the pixel-major kernels. */
int ojdf_nchw_to_nhwc(const float *in_dev, int C, int npix, float *out_dev, int out_stride, int out_coffset,
                      void *stream);
int ojdf_nhwc_to_nchw(const float *in_dev, int in_stride, int in_coffset, int C, int npix, float *out_dev,
                      void *stream);
/* Network input assembly (modules/pipeline.py:74-102, modules/model.py:269,274): head A gets
 * [values(P) | weights(P) | last_a] (the depth frame), head B (optional) [values | weights | last_b]
 * (the normalised label frame), pixel-major with `stride` floats per pixel. */
int ojdf_pack_fusion_input(const float *vals_dev, const float *wts_dev, const float *last_a_dev,
                           const float *last_b_dev, int npix, int P, float *out_a_dev, float *out_b_dev,
                           int stride, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OJDF_H */
