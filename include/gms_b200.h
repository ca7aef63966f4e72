/*
 * gms_b200.h -- C ABI of libgms_b200.so, the B200-native (sm_100a) mesh-Gaussian rasterizer.
 *
 * This is the drop-in boundary for the ONE hot path of waczjoan/gaussian-mesh-splatting:
 *   mesh -> Gaussian expansion, preprocess (3D->2D + SH), tile binning/sort, per-tile alpha
 *   compositing forward and backward.
 * Plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*.  No torch types.
 * Every function returns 0 on success, a negative GMS_E_* code otherwise; gms_last_error()
 * returns a human-readable message for the calling thread's last failure.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference; the rasterizer
 * itself is the un-vendored submodule submodules/diff-gaussian-rasterization, so for it the cited lines
 * are the reference's CALL SITES and [upstream] names the function of graphdeco-inria/diff-gaussian-rasterization
 * that the reference's pybind module `diff_gaussian_rasterization._C` exports):
 *
 *   gms_rasterize_forward    <- _C.rasterize_gaussians            [upstream rasterize_points.cu: RasterizeGaussiansCUDA]
 *                               reached from renderer/gaussian_renderer/__init__.py:94-102 (and :104-112, :97-105,
 *                               :99-107 of the animated / points / flame renderers)
 *   gms_rasterize_backward   <- _C.rasterize_gaussians_backward   [upstream RasterizeGaussiansBackwardCUDA]
 *                               reached from loss.backward(), train.py:108
 *   gms_mark_visible         <- _C.mark_visible                   [upstream markVisible]
 *   gms_expand_forward       <- GaussianMeshModel.update_alpha + _calc_xyz + prepare_scaling_rot + rot_to_quat_batch
 *                               games/mesh_splatting/scene/gaussian_mesh_model.py:86-169, utils/general_utils.py:19-96,
 *                               and the activation getters scene/gaussian_model.py:95-115 (optional fused outputs)
 *   gms_expand_backward      <- the autograd graph of the above (train.py:108 -> vertices/_alpha/_scale .grad)
 *
 * Scratch memory follows the stock extension's ownership model: the CALLER owns every byte.  The library
 * asks for its three scratch regions (per-Gaussian "geom", per-duplicate "binning", per-pixel "image")
 * through a callback, exactly like the resize-lambdas the stock _C module hands to CudaRasterizer
 * [upstream rasterize_points.cu: resizeFunctional]; the Python shim serves them from torch uint8 tensors
 * and keeps them alive for backward (ctx.save_for_backward in the stock shim).
 */
/* Threading: the library keeps per-PROCESS state (tuning options, launch counter, kernel timers, the pinned word the
 * synchronising forward reads N through).  Calls must come from one host thread at a time; concurrent calls from several
 * threads -- or interleaving a forward on one stream with option changes -- are not supported (the stock extension has
 * the same restriction: it launches on the legacy default stream and blocks on a cudaMemcpy).  Use one process per GPU,
 * as bench.py / torchrun do.  Error strings (gms_last_error) are per thread. */
#ifndef GMS_B200_H
#define GMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMS_OK 0
#define GMS_E_ARG (-1)        /* bad argument combination (e.g. both shs and colors_precomp) */
#define GMS_E_CUDA (-2)       /* a CUDA call or launch failed; see gms_last_error() */
#define GMS_E_ALLOC (-3)      /* the allocation callback returned NULL */
#define GMS_E_UNSUPPORTED (-4)

#define GMS_BUF_GEOM 0
#define GMS_BUF_BINNING 1
#define GMS_BUF_IMAGE 2

/* Scratch allocator: must return a device pointer to at least `bytes` bytes, 256-byte aligned, valid on
 * `stream` order (a torch.empty(uint8) tensor's data_ptr qualifies), or NULL. */
typedef void* (*gms_alloc_fn)(void* user, int which, size_t bytes);

/* The 13 fields of GaussianRasterizationSettings (renderer/gaussian_renderer/__init__.py:43-57) plus sizes.
 * Matrices/vectors stay on the DEVICE (they are CUDA tensors at the call site). */
typedef struct gms_raster_settings {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    const float* bg;          /* device [3] */
    float scale_modifier;
    const float* viewmatrix;  /* device [16], world_view_transform (transposed W2C), scene/cameras.py:54 */
    const float* projmatrix;  /* device [16], full_proj_transform, scene/cameras.py:56 */
    int32_t sh_degree;        /* active degree 0..3 */
    const float* campos;      /* device [3] */
    int32_t prefiltered;
    int32_t debug;            /* 1: synchronise + check after every launch (stock --debug behaviour) */
    int32_t antialiasing;
} gms_raster_settings;

/* Inputs of GaussianRasterizer.forward (renderer/gaussian_renderer/__init__.py:94-102). Exactly one of
 * shs / colors_precomp and exactly one of (scales, rotations) / cov3D_precomp must be non-NULL. */
typedef struct gms_raster_inputs {
    int32_t P;                   /* number of Gaussians */
    int32_t M;                   /* SH coefficients stored per Gaussian (16 for degree 3); 0 if no shs */
    const float* means3D;        /* [P,3] */
    const float* opacities;      /* [P,1] activated */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL */
    const float* scales;         /* [P,3] activated, or NULL */
    const float* rotations;      /* [P,4] (w,x,y,z) normalised, or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
} gms_raster_inputs;

/* Forward outputs (caller-allocated). */
#define GMS_FORWARD_ONLY 1    /* flags bit 0: no backward will follow (inference / no_grad): skip the survivor lists */
typedef struct gms_raster_outputs {
    float* out_color;     /* [3,H,W] */
    int32_t* radii;       /* [P] */
    float* out_invdepth;  /* [1,H,W] */
    int32_t flags;        /* 0, or GMS_FORWARD_ONLY */
} gms_raster_outputs;

/* What forward hands back for backward: scratch base pointers (as returned by the callback) and N. */
typedef struct gms_raster_saved {
    void* geom;
    void* binning;
    void* image;
    int64_t num_rendered;   /* N = number of (tile, Gaussian) duplicates */
    int64_t num_visible;    /* Gaussians with radii > 0 (statistics; may be -1 if not computed) */
    int64_t binning_capacity; /* duplicates the binning region was sized for (== num_rendered on the synchronising call) */
    int32_t flags;          /* bit 0: counting binning layout (point list first; no key arrays); bit 1: per-quad survivor lists
                               of the forward compositing pass follow the point list (consumed by backward); bit 2: the sorted
                               tile keys are 16-bit (tile count <= 65535) -- gms_debug_get_views then reports no key array */
} gms_raster_saved;

/* Gradients produced by backward (caller-allocated; NULL where the corresponding input was NULL). */
typedef struct gms_raster_grads {
    float* dL_dmeans3D;       /* [P,3] */
    float* dL_dmeans2D;       /* [P,3] NDC-scaled screen-space gradient, z = 0 */
    float* dL_dopacities;     /* [P,1] */
    float* dL_dshs;           /* [P,M,3] */
    float* dL_dcolors_precomp;/* [P,3] */
    float* dL_dscales;        /* [P,3] */
    float* dL_drotations;     /* [P,4] */
    float* dL_dcov3D_precomp; /* [P,6] */
    float* dL_dcolors_sh;     /* [P,3] optional, with shs: the clamp-masked colour gradient.  dL/dSH[k][c] = basis_k(dir) * this[c]
                                 (rank 1 per camera), so with it dL_dshs may be NULL: 12 B instead of 192 B per Gaussian
                                 (gms_adam_sh_factored consumes it) */
} gms_raster_grads;

/* ---- rasterizer ------------------------------------------------------------------------------- */

/* Bytes of the per-Gaussian and per-pixel scratch regions (binning depends on N and is requested
 * through the callback once N is known). */
int gms_scratch_bytes(int32_t P, int32_t W, int32_t H, size_t* geom_bytes, size_t* image_bytes);
size_t gms_binning_bytes(int64_t num_rendered, int32_t P);

int gms_rasterize_forward(const gms_raster_settings* settings, const gms_raster_inputs* in,
                          const gms_raster_outputs* out, gms_alloc_fn alloc, void* alloc_user,
                          gms_raster_saved* saved, void* cuda_stream);

/* The same forward WITHOUT the stock pipeline's per-frame host synchronisation ([upstream rasterizer_impl.cu:
 * cudaMemcpy(&num_rendered, ...)]): the caller sizes the binning region for `binning_capacity` duplicates, N is computed
 * and consumed on the device, saved->num_rendered is -1.  `n_host_mapped` (optional) is device-accessible pinned HOST
 * memory [2]: the binning kernel stores N and an overflow flag there, readable later without a sync.  If N exceeds the
 * capacity the frame degrades to the background image with zero gradients (flag = 1) -- grow and re-run.  No allocation,
 * no blocking call: the whole frame can be captured in a CUDA graph. */
int gms_rasterize_forward_nosync(const gms_raster_settings* settings, const gms_raster_inputs* in,
                                 const gms_raster_outputs* out, gms_alloc_fn alloc, void* alloc_user,
                                 gms_raster_saved* saved, int64_t binning_capacity, uint32_t* n_host_mapped,
                                 void* cuda_stream);

/* dL_dout_invdepth may be NULL (train.py never puts a loss on render_pkg["depth"]). */
int gms_rasterize_backward(const gms_raster_settings* settings, const gms_raster_inputs* in,
                           const int32_t* radii, const gms_raster_saved* saved,
                           const float* dL_dout_color /*[3,H,W]*/, const float* dL_dout_invdepth /*[1,H,W]|NULL*/,
                           const gms_raster_grads* grads, void* cuda_stream);

int gms_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present /*[P] bool*/, void* cuda_stream);

/* Read-only views into the scratch regions, for parity tests and statistics (device pointers). */
typedef struct gms_debug_views {
    const float* means2D;        /* packed records [P,12] floats (see gms_debug_unpack for the stock layouts) */
    const float* depths;         /* unused (NULL) */
    const float* cov3D;          /* [P,6] */
    const float* conic_opacity;  /* unused (NULL) */
    const float* rgb;            /* unused (NULL) */
    const uint8_t* clamped;      /* unused (NULL) */
    const uint32_t* tiles_touched; /* [P] */
    const uint32_t* point_list;  /* [N] Gaussian index per duplicate, sorted by (tile, depth bits, index) */
    const uint32_t* tile_keys;   /* [N] tile id per duplicate, sorted */
    const int32_t* ranges;       /* [T,2] */
    const float* final_T;        /* [H,W] */
    const int32_t* n_contrib;    /* [H,W] */
    const float* dgeom;          /* [P,12] after a backward call: per-Gaussian sums the composite backward accumulated --
                                    dL/dmean2D.xy (NDC-scaled), dL/dconic.xyz (stock half convention for xy), dL/d(conic_opacity.w),
                                    dL/drgb.rgb, dL/dinvdepth, 2 unused -- i.e. the input of the preprocess backward */
} gms_debug_views;
int gms_debug_get_views(const gms_raster_saved* saved, int32_t P, int32_t W, int32_t H, gms_debug_views* views);
/* The per-Gaussian preprocess results live in packed 48-byte records; this unpacks them into the stock
 * layouts (caller-allocated device buffers, any may be NULL): means2D [P,2], depths [P],
 * conic_opacity [P,4], rgb [P,3], clamped [P,3] bytes.  Rows of culled Gaussians are zero. */
int gms_debug_unpack(const gms_raster_saved* saved, int32_t P, const int32_t* radii, float* means2D, float* depths,
                     float* conic_opacity, float* rgb, uint8_t* clamped, void* cuda_stream);

/* ---- mesh -> Gaussian expansion --------------------------------------------------------------- */

typedef struct gms_expand_args {
    int32_t V, F, K;             /* vertices, faces, splats per face (P = F*K) */
    const float* vertices;       /* [V,3] */
    const int64_t* faces;        /* [F,3] int64 (torch.long), gaussian_mesh_model.py:74 */
    const float* triangles_in;   /* [F,3,3] or NULL: animated path passes transformed triangles directly
                                    (renderer/gaussian_animated_renderer/__init__.py:61-73); then vertices/faces unused */
    const float* alpha_raw;      /* _alpha [F,K,3] */
    const float* scale_raw;      /* _scale [P,1] */
    float eps;                   /* eps_s0 = 1e-8, gaussian_mesh_model.py:43 */
    /* outputs; any may be NULL */
    float* alpha;                /* [F,K,3] normalised barycentrics (pc.alpha) */
    float* triangles;            /* [F,3,3] (pc.triangles) */
    float* xyz;                  /* [P,3] (pc._xyz) */
    float* scaling_log;          /* [P,3] (pc._scaling) */
    float* rotation_raw;         /* [P,4] (pc._rotation) */
    float* scaling_act;          /* [P,3] = exp(_scaling)        (get_scaling, fused E4) */
    float* rotation_act;         /* [P,4] = normalize(_rotation) (get_rotation, fused E4) */
} gms_expand_args;

int gms_expand_forward(const gms_expand_args* a, void* cuda_stream);

typedef struct gms_expand_grads {
    /* incoming (any may be NULL = zero) */
    const float* dL_dxyz;          /* [P,3] */
    const float* dL_dscaling_log;  /* [P,3] */
    const float* dL_drotation_raw; /* [P,4] */
    const float* dL_dscaling_act;  /* [P,3] gradient w.r.t. exp(_scaling) */
    const float* dL_drotation_act; /* [P,4] gradient w.r.t. normalize(_rotation) */
    /* outgoing */
    float* dL_dvertices;           /* [V,3] ACCUMULATED with atomics: caller zero-fills; NULL if triangles_in */
    float* dL_dtriangles;          /* [F,3,3] written (NULL allowed) */
    float* dL_dalpha_raw;          /* [F,K,3] */
    float* dL_dscale_raw;          /* [P,1] */
} gms_expand_grads;

int gms_expand_backward(const gms_expand_args* a, const gms_expand_grads* g, void* cuda_stream);

/* gs_points pseudo-mesh path: one triangle per Gaussian -> (xyz = v1, 2 log-scales, quaternion).  Replaces
 * PointsGaussianModel.prepare_scaling_rot (games/flat_splatting/scene/points_gaussian_model.py:61-104) as called per
 * frame by renderer/gaussian_points_animated_renderer/__init__.py:61-66.  Forward only (render scripts, no_grad). */
typedef struct gms_points_args {
    int32_t P;
    const float* triangles;      /* [P,3,3] */
    float eps;                   /* 1e-8 */
    float* xyz;                  /* [P,3] = triangles[:,0]            (any output may be NULL) */
    float* scaling_log;          /* [P,2] (pc._scaling) */
    float* rotation_raw;         /* [P,4] (pc._rotation) */
    float* scaling_act;          /* [P,3] = (eps, exp(_scaling))      (get_scaling) */
    float* rotation_act;         /* [P,4] = normalize(_rotation)      (get_rotation) */
} gms_points_args;
int gms_points_expand_forward(const gms_points_args* a, void* cuda_stream);

/* gs_points, the other direction: every flat Gaussian -> its pseudo-mesh triangle (v1 = xyz, v2/v3 = xyz + s*axis, the
 * longer arm first).  Replaces PointsGaussianModel.prepare_vertices (games/flat_splatting/scene/points_gaussian_model.py:
 * 28-59, with get_scaling :106-109 and build_rotation utils/general_utils.py:158-179), run once when a trained gs_flat
 * model is loaded for editing (scripts/render_points_time_animated.py).  Forward only. */
typedef struct gms_points_vertices_args {
    int32_t P;
    const float* xyz;            /* [P,3]  pc._xyz */
    const float* scaling_log;    /* [P,scaling_cols]  pc._scaling; the LAST two columns are the in-plane log-scales */
    int32_t scaling_cols;        /* 2 (gs_points) or 3 (a gs_flat checkpoint) */
    const float* rotation_raw;   /* [P,4]  pc._rotation (w,x,y,z), not normalised */
    float* triangles;            /* out [P,3,3] */
} gms_points_vertices_args;
int gms_points_prepare_vertices(const gms_points_vertices_args* a, void* cuda_stream);

/* ---- training-step glue on the same stream (the callers either side of the path) ---- */

/* L = (1-lambda)*L1 + lambda*(1-SSIM) and dL/dimg in two launches.  Replaces utils/loss_utils.py:17-64 as used by
 * train.py:105-108 (5 grouped conv2d forward + 5 backward + ~30 elementwise launches per frame). */
typedef struct gms_loss_args {
    int32_t C, H, W;
    const float* img;        /* [C,H,W] rendered image */
    const float* gt;         /* [C,H,W] ground truth */
    float lambda_dssim;      /* 0.2, arguments/__init__.py:86 */
    const float* dL_dloss;   /* device scalar upstream gradient, or NULL for 1 */
    float* loss;             /* device [3]: loss, L1 mean, SSIM mean */
    float* dL_dimg;          /* [C,H,W] or NULL (forward only) */
    void* scratch;           /* gms_loss_scratch_bytes() bytes */
    size_t scratch_bytes;
} gms_loss_args;
int gms_loss_scratch_bytes(int32_t C, int32_t H, int32_t W, size_t* bytes);
int gms_l1_ssim_loss(const gms_loss_args* a, void* cuda_stream);

/* torch.optim.Adam(groups, eps=1e-15) of gaussian_mesh_model.py:171-183 (train.py:146-148) over ONE flat fp32
 * parameter buffer, one launch; consumes and (optionally, fully or only a leading range) zeroes the gradient in the same pass.
 * Segment i covers flat indices [seg_end[i-1], seg_end[i]); lr = lr0[i], or -- when period[i] > 0 --
 * lr0[i] where ((index - segment start) / inner[i]) % period[i] == 0 and lr1[i] elsewhere (DC vs rest SH). */
typedef struct gms_adam_args {
    int64_t n;               /* elements this call updates */
    int64_t offset;          /* flat index of element 0 (sharded optimizer: each rank updates [offset, offset+n)); 0 otherwise */
    float* p; float* g; float* m; float* v;   /* pointers to element `offset` of the respective flat buffers */
    int32_t nseg;
    int64_t seg_end[8];
    double lr0[8]; double lr1[8];   /* doubles, like torch's Python-float hyper-parameters: lr / (1 - beta1^t), 1 - beta */
    int32_t inner[8]; int32_t period[8];
    double beta1, beta2, eps;       /* and sqrt(1 - beta2^t) are formed in double and rounded to fp32 once */
    int32_t step;            /* 1-based step count (bias correction) */
    int32_t zero_grad;       /* 0: leave g; 1: zero every consumed element; 2: zero only flat indices < zero_end */
    int64_t zero_end;        /* (mode 2) e.g. the end of the vertices segment when every other gradient is overwritten
                                by the next frame (gms_train_frame) */
} gms_adam_args;
int gms_adam_step(const gms_adam_args* a, void* cuda_stream);

/* One gs_mesh training frame in ONE call (train.py:100-108 + :154-157 of the reference: render + loss + backward +
 * re-expansion), all launches on `cuda_stream`, no Python / autograd in between:
 *   expansion fwd (activated scales/rotations) -> sigmoid(opacity) -> rasterizer fwd -> L1+SSIM -> rasterizer bwd ->
 *   sigmoid bwd -> expansion bwd.
 * Model tensors are the RAW parameters; gradients are written (d_vertices: accumulated with atomics, keep it zeroed)
 * into caller buffers, e.g. views of the flat gradient buffer gms_adam_step / the NCCL exchange operate on.
 * `workspace` holds the per-frame intermediates (gms_frame_workspace_bytes); the rasterizer's scratch still comes
 * through the allocation callback. */
typedef struct gms_frame_args {
    int32_t V, F, K, M;
    const float* vertices; const int64_t* faces; const float* alpha_raw; const float* scale_raw;
    const float* features;      /* [P,M,3] packed SH (get_features) */
    const float* opacity_raw;   /* [P,1] logits */
    float eps;                  /* eps_s0 */
    float* d_vertices; float* d_alpha_raw; float* d_scale_raw; float* d_features; float* d_opacity_raw;
    gms_raster_settings settings;
    const float* gt;            /* [3,H,W] */
    float lambda_dssim;
    float* loss;                /* device [3]: loss, L1, SSIM */
    void* workspace; size_t workspace_bytes;
    int64_t* num_rendered;      /* host, optional (-1 on the sync-free path) */
    int64_t binning_capacity;   /* > 0: sync-free frame (gms_rasterize_forward_nosync semantics); 0: stock-style, one 4-byte D2H */
    uint32_t* n_host_mapped;    /* optional mapped pinned host [2]: N, overflow flag */
    float* d_color_sh;          /* optional [3P + 3]: factored SH gradient (dL_dcolors_sh) followed by this frame's camera centre;
                                   then d_features may be NULL and no SH gradient rows are written */
    void* event_sh_ready;       /* optional cudaEvent_t recorded right after the preprocess backward: d_color_sh (or d_features) is
                                   final from there on, so a data-parallel caller can start exchanging it on another stream while the
                                   opacity / expansion backward still runs */
    void* event_loss_ready;     /* optional cudaEvent_t recorded right after the loss kernels (about 40 % into the frame): a caller that
                                   logs the loss every step copies it to the host from another stream and can queue the next
                                   frame while this one's backward pass still runs */
} gms_frame_args;
size_t gms_frame_workspace_bytes(int32_t P, int32_t W, int32_t H);
/* Device pointers into a frame workspace (valid after gms_train_frame): this step's expansion outputs and images. */
typedef struct gms_frame_view {
    const float* xyz; const float* scales; const float* rotations; const float* opacities; const int32_t* radii;
    const float* image; const float* invdepth;
} gms_frame_view;
int gms_frame_views(void* workspace, int32_t P, int32_t W, int32_t H, gms_frame_view* v);

/* Adam step of the packed SH parameter [P,16,3] with its gradient rebuilt from FACTORS instead of read from memory:
 *   dL/dSH_i[k][c] = grad_scale * sum_r basis_k(normalize(xyz_i - campos_r)) * dcolor_r[i][c],   r = 0..R-1
 * `exchange` holds R slots of `slot_floats` floats, slot r = [3P colour gradients of frame r | its camera centre (3) | pad]
 * (what gms_train_frame writes through d_color_sh).  Data parallel: the slots are all-gathered (12 B per Gaussian and
 * rank instead of a 192 B reduce-scatter + all-gather), every rank runs this kernel on all Gaussians (replicated moments),
 * so no parameter all-gather is needed either.  Learning rates / bias correction as in gms_adam_step (DC coefficient:
 * lr_dc, the other 15: lr_rest; arguments_games/__init__.py:17-30). */
typedef struct gms_adam_sh_args {
    int32_t P, M, sh_degree, R;
    const float* xyz;          /* [P,3] Gaussian centres of this step (gms_frame_views) */
    const float* exchange;     /* [R, slot_floats] */
    int64_t slot_floats;
    float grad_scale;          /* 1/R */
    float* p; float* m; float* v;
    double lr_dc, lr_rest, beta1, beta2, eps;
    int32_t step;
} gms_adam_sh_args;
int gms_adam_sh_factored(const gms_adam_sh_args* a, void* cuda_stream);
int gms_train_frame(const gms_frame_args* a, gms_alloc_fn alloc, void* alloc_user, void* cuda_stream);

/* ---- image sink / source ---------------------------------------- */

/* float [C,H,W] -> 8-bit interleaved rows, byte = clamp(x * 255 + 0.5, 0, 255) truncated: the device half of
 * torchvision.utils.save_image (scripts/render_time_animated.py:86-87, scripts/render.py) in one pass.  Output row stride is
 * row_prefix + W*C bytes; row_prefix = 1 reserves PNG's filter byte (written 0), 0 gives plain HWC (PPM / raw video). */
int gms_image_quantize(const float* chw, uint8_t* out, int32_t C, int32_t H, int32_t W, int32_t row_prefix, void* cuda_stream);
/* 8-bit image ([H,W,C] if src_is_hwc else [C,H,W]) -> float [C,H,W] = byte / 255 (ToTensor / PILtoTorch,
 * utils/general_utils.py:105-112): ground-truth images can stay 8-bit on the host and on the device. */
int gms_image_dequantize(const uint8_t* src, int32_t src_is_hwc, float* chw, int32_t C, int32_t H, int32_t W, void* cuda_stream);

/* ---- misc ------------------------------------------------------------------------------------- */
const char* gms_last_error(void);
const char* gms_version(void);
/* Number of kernel launches issued by this library since the last call with reset != 0 (bench.py's
 * "gpu_launches" claim is counted, not guessed). */
int64_t gms_launch_count(int reset);
/* Per-kernel device time, measured with CUDA events recorded on the launching stream around every launch made
 * while option "time_kernels" is 1.  Fills up to max_kernels entries (accumulated ms, launch count, name) and
 * returns the number of kernel slots. */
int gms_kernel_times(int reset, int max_kernels, double* ms_out, int64_t* count_out, const char** names_out);
/* Tuning knobs (round-over-round experiments): "quad_masks", "warp_emit", "time_kernels", "composite_version",
 * "composite_fwd", "composite_bwd", "bwd_minblocks", "tile_order", "sort_impl", "expand_staged", "sh_staged"
 * (DESIGN.md lists what each selects).  Returns the previous value; unknown keys return -1. */
int gms_set_option(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* GMS_B200_H */
