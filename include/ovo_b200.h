/* ovo_b200 — C ABI of the B200-native hot path of tberriel/OVO
 * (CLIP/ViT region encoder -> 3D association/fusion into the point map -> text-vs-map cosine query).
 *
 * The reference has no FFI on this path: its seam is the Python class surface of `OVO`,
 * `CLIPGenerator`, `MaskGenerator`, `Instance3D` (the modules under ovo/entities).  ovo_b200/ mirrors those classes in
 * Python and calls the entry points below through ctypes; INTEGRATION.md shows the binding.
 * Each entry point cites the reference lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer named *_dev is a CUDA device pointer owned by the caller (PyTorch); *_host is host memory;
 *   - all calls return 0 on success or a negative OVO_E_* code; ovo_last_error() gives the message
 *     (thread-local); no C++ exception crosses the boundary;
 *   - kernels are enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     calls are asynchronous except where a host result is documented;
 *   - a handle is not thread-safe; distinct handles are.  One handle per GPU.
 */
#ifndef OVO_B200_H
#define OVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVO_OK 0
#define OVO_E_INVALID (-1)  /* bad argument / unsupported shape */
#define OVO_E_CUDA (-2)     /* CUDA runtime / driver error */
#define OVO_E_NOMEM (-3)    /* workspace allocation failed */
#define OVO_E_STATE (-4)    /* call sequence error (e.g. fuse before associate) */

typedef struct ovo_encoder ovo_encoder_t;
typedef struct ovo_map ovo_map_t;

const char* ovo_last_error(void);
int ovo_version(void);
/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as `gpu_launches`). */
long long ovo_launch_count(int reset);

/* ------------------------------------------------------------------------------------------------
 * Encoder: PE ViT (vision tower + text tower)
 *   thirdParty/perception_models/core/vision_encoder/pe.py:293-533 (VisionTransformer),
 *   pe.py:552-695 (TextTransformer), config.py:101-118 (PE-Core-L14-336).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int image_size; /* 336 */
  int patch_size; /* 14 */
  int width;      /* 1024; head_dim = width/heads: 64 (tcgen05 attention) or 32 / 80 / 96 (generic mma.sync attention, e.g. the
                   * ViT-H/14-shaped encoder of BASELINE config 4: 1280 / 16 = 80) */
  int layers;     /* 24 */
  int heads;      /* 16 */
  int mlp_width;  /* 4096 */
  int output_dim; /* 1024 */
  float ln_eps;   /* 1e-5 (pe.py:302) */
  /* text tower (0 layers = no text tower) */
  int text_ctx;   /* 32 */
  int text_width; /* 1024 */
  int text_heads; /* 16 */
  int text_layers;
  int text_mlp_width;
  int vocab_size; /* 49408 */
  int text_output_dim;
} ovo_vit_cfg;

/* Device pointers to one transformer block's parameters (reference state_dict names):
 *   ln_1.{weight,bias}, attn.in_proj_{weight [3W,W], bias}, attn.out_proj.{weight [W,W], bias},
 *   ln_2.{weight,bias}, mlp.c_fc.{weight [F,W], bias}, mlp.c_proj.{weight [W,F], bias}.
 * Matrices are bf16 row-major exactly as nn.Linear stores them ([out,in]); vectors are f32. */
typedef struct {
  const float* ln1_w; const float* ln1_b;
  const void* qkv_w;  const float* qkv_b;
  const void* out_w;  const float* out_b;
  const float* ln2_w; const float* ln2_b;
  const void* fc_w;   const float* fc_b;
  const void* proj_w; const float* proj_b;
} ovo_block_weights;

typedef struct {
  /* vision tower */
  const void* patch_w;      /* bf16 [width, kpad]: conv1.weight flattened (c,ky,kx) and zero padded to kpad */
  int patch_kpad;           /* 3*patch*patch rounded up to a multiple of 64 (640 for 14x14) */
  const float* cls_pos0;    /* f32 [width]: class_embedding + positional_embedding[0] (pe.py:512-519) */
  const float* pos;         /* f32 [1+grid^2, width] positional_embedding */
  const float* ln_pre_w; const float* ln_pre_b;
  const float* ln_post_w; const float* ln_post_b;
  const ovo_block_weights* blocks;          /* host array [layers] of device-pointer structs */
  /* region pooling (textregion.py:163-195 closed form): r = normalize(mean . pool_w^T + pool_b),
   * pool_w = (W_v^T W_o^T proj)^T as bf16 [output_dim, width], pool_b f32 [output_dim] */
  const void* pool_w; const float* pool_b;
  const float* pool_b_empty; /* f32 [output_dim] = out_proj.bias @ proj: a mask that covers no token has every
                              * key padded; torch's MHA (safe softmax, torch >= 2.5) then attends to nothing and
                              * the region feature is normalize(out_proj.bias @ proj) */
  /* text tower */
  const float* tok_emb;     /* f32 [vocab, text_width] token_embedding.weight */
  const float* text_pos;    /* f32 [ctx, text_width] */
  const ovo_block_weights* text_blocks;     /* host array [text_layers] */
  const float* ln_final_w; const float* ln_final_b;
  const void* text_proj_w;  /* bf16 [text_output_dim, text_width] = text_projection^T */
} ovo_vit_weights;

/* Creates an encoder able to process up to max_images 336x336 crops per call (activations are
 * pre-allocated), frames of at most max_h x max_w and max_masks masks per call. */
int ovo_encoder_create(const ovo_vit_cfg* cfg, const ovo_vit_weights* w, int max_images, int max_h, int max_w,
                       int max_masks, ovo_encoder_t** out);
void ovo_encoder_destroy(ovo_encoder_t* enc);

/* E1: PETextRegion.get_img_features crops + T.Resize(antialias) + Normalize (textregion.py:104-134,
 * transforms.py:19-26).  rgb uint8 [n_frames,H,W,3] -> im2col'd patches inside the encoder.  Returns the
 * number of 336x336 images per frame in *n_img_per_frame (1 + floor(H/336)*floor(W/336)). */
int ovo_encoder_preprocess(ovo_encoder_t* enc, const uint8_t* rgb_dev, int n_frames, int H, int W,
                           int* n_img_per_frame, void* stream);
/* Test tap: same as above but from already normalised float pixels [n_img,3,S,S] (skips the resize). */
int ovo_encoder_load_pixels(ovo_encoder_t* enc, const float* pixels_dev, int n_img, void* stream);
/* E2: VisionTransformer.forward_features(norm=True) (pe.py:499-533) on the images loaded by
 * preprocess/load_pixels.  n_layers < 0 = all.  tokens_out_dev (optional, may be NULL) receives f32
 * [n_img, 1+grid^2, width]; apply_ln_post selects the final ln_post. */
int ovo_encoder_forward(ovo_encoder_t* enc, int n_img, int n_layers, int apply_ln_post, float* tokens_out_dev,
                        void* stream);
/* E3-E5: resize_features + get_features_mask + pe_value_with_sam2_attn (textregion.py:9-28,145-195) for
 * ONE frame whose n_img images start at image index img0 of the last forward.
 * masks uint8 [M,H,W] (non-zero = set) -> out f32 [M, output_dim], unit norm (a mask that covers no token
 * gets normalize(pool_b_empty), as the reference does under torch >= 2.5). */
int ovo_encoder_pool_regions(ovo_encoder_t* enc, int img0, int H, int W, const uint8_t* masks_dev, int M,
                             float* out_dev, void* stream);
/* CLIPGenerator.extract_clip, TextRegion branch (clip_generator.py:125-135): E1..E5 for a batch of
 * frames.  masks are concatenated over frames, masks_per_frame_host[f] of them belong to frame f. */
int ovo_encode_regions(ovo_encoder_t* enc, const uint8_t* rgb_dev, int n_frames, int H, int W,
                       const uint8_t* masks_dev, const int* masks_per_frame_host, float* out_dev, void* stream);
/* Q1: CLIP.encode_text (pe.py:671-695,725): tokens int32 [T, ctx] -> f32 [T, text_output_dim], NOT
 * normalised (the caller applies clip_generator.py:170-173,193-196). */
int ovo_encode_text(ovo_encoder_t* enc, const int32_t* tokens_dev, int T, float* out_dev, void* stream);

/* get_embed_txt_similarity's text side (clip_generator.py:161-173,186-196): tokens int32 [Q*T, ctx] (T templates
 * per query, query-major) -> f32 [Q, text_output_dim] = normalize(mean_t(normalize(encode_text))). */
int ovo_text_bank(ovo_encoder_t* enc, const int32_t* tokens_dev, int Q, int T, float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Nearest neighbours and label transfer
 * ---------------------------------------------------------------------------------------------- */
/* Exact k nearest neighbours (k <= 8) of every query among N points, ascending by distance, ties by point index:
 * what scipy.spatial.KDTree(points).query(queries, k) returns in match_labels_to_vtx (ovo/utils/eval_utils.py:22-27)
 * and, with k = 1, Open3D's compute_point_cloud_distance in same_instance (ovo/utils/instance_utils.py:16-22).
 * points f32 [N,3], queries f32 [Q,3] -> idx int32 [Q,k], dist f64 [Q,k] (optional, may be NULL).
 * cell_size <= 0 picks the grid cell from the point density.  Needs N >= k.  Two host synchronisations (the
 * bounding box sizes the grid; the count of far-away queries sizes the exhaustive fallback). */
int ovo_knn(const float* points_dev, int64_t N, const float* queries_dev, int64_t Q, int k, float cell_size,
            int32_t* idx_out_dev, double* dist_out_dev, void* stream);
/* Diagnostics of the calling thread's last ovo_knn: the grid cell it used, the number of cells, and how many queries
 * went through the exhaustive fallback. */
void ovo_knn_stats(float* cell_size, int* n_cells, int* n_fallback);
/* torch.mode over the k labels a query's neighbours carry (eval_utils.py:29-30): labels int32 [N], idx int32 [Q,k]
 * -> out int32 [Q] = the most frequent label, the smallest one on ties. */
int ovo_knn_mode(const int32_t* labels_dev, const int32_t* idx_dev, int64_t Q, int k, int32_t* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Crop-based descriptors (CLIPGenerator.extract_clip's crop branch,
 * ovo/entities/clip_generator.py:136-158): embed types vanilla / fixed_weights / hovsg / adaptive_weights /
 * concept_fusion.  Each mask contributes a masked crop and a margin crop (ovo/utils/segment_utils.py:29-182), every
 * crop goes through the full encode_image = ViT + attention pooling + projection (pe.py:44-87,535-543), and the
 * frame's global descriptor, the masked-crop descriptor and the margin-crop descriptor are blended by fuse_clips
 * (ovo/utils/clip_utils.py:21-48).
 * ---------------------------------------------------------------------------------------------- */
/* The attention-pooling head of pe.VisionTransformer (`visual.attn_pool.*`, `visual.proj`); device pointers, matrices
 * bf16 [out,in] as nn.Linear stores them, vectors f32. */
typedef struct {
  int heads;                /* attn_pooler_heads (8 for PE-Core-L14-336, config.py:46); width/heads <= 256 */
  int mlp_width;            /* 4 * width (pe.py:52,70) */
  const float* q;           /* f32 [width] = (probe @ Wq^T + bq) * head_dim^-0.5: the probe is a parameter, so the query
                             * projection of pe.py:83-84 does not depend on the image */
  const void* kv_w;         /* bf16 [2*width, width] = attn.in_proj_weight[width:3*width] (keys | values) */
  const float* kv_b;        /* f32 [2*width] */
  const void* out_w; const float* out_b;     /* attn.out_proj */
  const float* ln_w; const float* ln_b;      /* layernorm */
  const void* fc_w; const float* fc_b;       /* mlp.c_fc [mlp_width, width] */
  const void* proj_w; const float* proj_b;   /* mlp.c_proj [width, mlp_width] */
  const void* vis_proj_w;   /* bf16 [output_dim, width] = visual.proj^T (pe.py:540-541) */
} ovo_pool_head_weights;
/* Installs the head (allocates its small workspaces).  Required before ovo_encode_images / ovo_encode_crops. */
int ovo_encoder_set_pool_head(ovo_encoder_t* enc, const ovo_pool_head_weights* w);
/* pe.CLIP.encode_image (pe.py:717-719, normalize=False) on already normalised pixels f32 [n,3,S,S]
 * -> f32 [n, output_dim].  Test tap of the head; n <= max_images. */
int ovo_encode_images(ovo_encoder_t* enc, const float* pixels_dev, int n, float* out_dev, void* stream);

#define OVO_EMBED_VANILLA 0
#define OVO_EMBED_FIXED_WEIGHTS 1
#define OVO_EMBED_HOVSG 2
#define OVO_EMBED_ADAPTIVE_WEIGHTS 3
#define OVO_EMBED_CONCEPT_FUSION 4
typedef struct {
  int embed_type;   /* OVO_EMBED_* */
  int return_all;   /* clip_generator.py:151-152: out is [M,3,D] = (global, masked crop, margin crop), no fusion */
  int mask_res;     /* side of the crops before the encoder's own resize (config `mask_res`, clip_generator.py:16) */
  int bbox_margin;  /* 50 (segment_utils.py:29) */
  float w_masked;   /* 0.4418 (clip_generator.py:33) */
  float w_global;   /* 0.1    (clip_generator.py:34) */
} ovo_crop_params;
/* batched_mask_to_box + batched_box_xyxy_to_xywh (segment_utils.py:43-104): masks uint8 [M,H,W] -> int32 [M,4]
 * (x, y, w, h) with w = right - left, h = bottom - top (the reference's convention); empty mask -> 0,0,0,0. */
int ovo_mask_boxes(const uint8_t* masks_dev, int M, int H, int W, int32_t* xywh_dev, void* stream);
/* extract_clip, crop branch, for one frame: rgb uint8 [H,W,3], masks uint8 [M,H,W] -> out f32 [M, output_dim] unit norm
 * ([M,3,output_dim] with return_all).  crops_out_dev (optional test tap, may be NULL) receives the uint8 crops
 * [n_crops, mask_res, mask_res, 3] (masked crops first, then margin crops; `vanilla` has masked crops only).
 * One host synchronisation (the boxes decide the crop geometry).  A mask whose box has zero width or height
 * (`vanilla`: zero width AND height) makes the reference's F.resize raise; here the call returns OVO_E_INVALID. */
int ovo_encode_crops(ovo_encoder_t* enc, const uint8_t* rgb_dev, int H, int W, const uint8_t* masks_dev, int M,
                     const ovo_crop_params* prm, float* out_dev, uint8_t* crops_out_dev, void* stream);
/* fuse_clips (clip_utils.py:21-48): g f32 [D] (the frame's global descriptor), seg / bbox f32 [M,D], unit norm
 * -> out f32 [M,D].  embed_type 1..4. */
int ovo_fuse_clips(const float* g_dev, const float* seg_dev, const float* bbox_dev, int M, int D, int embed_type,
                   float w_masked, float w_global, float* out_dev, void* stream);
/* embed_type `learned`: WeightsPredictorMerger (ovo/entities/clips_merging.py:26-56; wired at clip_generator.py:18-29,153).
 * Device pointers; matrices bf16 [out,in], vectors f32, names as in the module's state_dict. */
typedef struct {
  const void* in_w; const float* in_b;      /* att_encoder.layers.{l}.self_attn.in_proj_{weight [3d,d], bias} */
  const void* out_w; const float* out_b;    /* self_attn.out_proj */
  const float* ln1_w; const float* ln1_b;   /* norm1 */
  const void* ff1_w; const float* ff1_b;    /* linear1 [ff,d] */
  const void* ff2_w; const float* ff2_b;    /* linear2 [d,ff] */
  const float* ln2_w; const float* ln2_b;   /* norm2 */
} ovo_merger_layer;
typedef struct {
  int d_model, nhead, dim_feedforward, n_layers;
  const ovo_merger_layer* layers;           /* host array [n_layers] */
  int n_linear;                             /* linears of `mlp` (hparams n_layers + 2) */
  const void* const* mlp_w;                 /* host array [n_linear] of device pointers, bf16 [out_j, in_j]; in_0 = 3*d_model */
  const float* const* mlp_b;                /* host array [n_linear] of device pointers */
  const int* mlp_out;                       /* host array [n_linear]: out_j; the last is 3*d_model (per-channel weights) or 3 */
  float ln_eps;                             /* 1e-5 (nn.TransformerEncoderLayer default) */
} ovo_merger_weights;
/* clips f32 [B,3,D] (global, masked crop, margin crop; unit norm) -> out f32 [B,D] unit norm. */
int ovo_merge_clips_learned(const ovo_merger_weights* w, const float* clips_dev, int B, float* out_dev, void* stream);
/* siglip_cosine_similarity (clip_utils.py:10-14) applied in place to a similarity matrix of n entries:
 * sim <- sigmoid(sim * exp(logit_scale) + logit_bias). */
int ovo_siglip_similarity(float* sim_dev, int64_t n, float logit_scale, float logit_bias, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Map: 3D association, instance vote, fusion, query
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_matched;    /* points matched into this mask                        (ovo.py:257) */
  int32_t n_assigned;   /* of those, points that already carry an instance id   (ovo.py:260) */
  int32_t n_unassigned;
  int32_t mode_id;      /* torch.mode of the assigned ids, ties -> smallest     (ovo.py:264); -1 if none */
  int32_t ins_id;       /* instance id given to the mask, -1 if none            (ovo.py:263-276) */
  int32_t is_new;       /* 1 if ins_id was newly allocated from next_ins_id     (ovo.py:271-276) */
  int32_t area;         /* (seg_map == mask).sum()                              (ovo.py:259) */
  int32_t reserved;
} ovo_vote_row;

typedef struct {
  const float* depth_dev;   /* [h,w] f32 metres, 0 = invalid */
  int h, w;
  const int32_t* seg_map_dev; /* [H,W] i32, -1 = no mask */
  int H, W;
  int n_masks;              /* seg_map.max()+1 */
  float c2w[16];            /* camera-to-world, row major */
  float w2c[16];            /* its inverse (torch.linalg.inv(c2w), ovo.py:216) */
  float K[9];               /* intrinsics, row major */
  float match_th;           /* semantic.match_distance_th (0.05) */
  int track_th;             /* semantic.track_th (100) */
  int depth_filter;         /* semantic.depth_filter */
  /* rgb/depth resolution fix-up (ovo.py:218-221): 0 = none, else u' = int((u+crop_edge)*ratio_w) */
  int has_ratio; float ratio_h, ratio_w; int crop_edge;
  /* Batched calls only (NULL otherwise): depth_dev is ALREADY the map the points are matched against (the depth filter was run
   * where the frame lives, ovo_depth_filter) and depth_range_dev -> f32 [2] = min / max of the RAW depth > 0 (ovo_depth_range),
   * from which the frustum is built (ovo.py:209).  A sharded map gathers these instead of filtering every keyframe on every rank. */
  const float* depth_range_dev;
} ovo_frame;

int ovo_map_create(ovo_map_t** out);
void ovo_map_destroy(ovo_map_t* map);
/* Optional: sizes the handle's workspaces once for a map of up to max_points points, max_instances instances, max_masks masks
 * per keyframe and max_matches matched points per keyframe (any argument <= 0 leaves that workspace alone).  Without it the
 * workspaces grow on demand, each growth being a cudaFree + cudaMalloc (a device-wide synchronisation: isolated frames of tens of
 * milliseconds in a stream whose map keeps growing).  Call it before the first association: match lists of earlier keyframes that
 * live in a re-allocated slot are dropped. */
int ovo_map_reserve(ovo_map_t* map, int64_t max_points, int max_instances, int max_masks, int64_t max_matches);

/* geometry_utils.depth_filter (geometry_utils.py:92-96): 7x7 gaussian (sigma 2.5, reflect) high-pass;
 * |d - blur| > 0.05 -> -1. */
int ovo_depth_filter(const float* depth_dev, int h, int w, float* out_dev, void* stream);
/* Both for n_frames depth maps [n_frames,h,w] in ONE launch: out_dev [n_frames,h,w] = the filtered maps (NULL: skip),
 * ranges_out_dev f32 [n_frames,2] = min / max of every map's values > 0 (NULL: skip) — what a rank of a sharded map computes for
 * its own keyframes before they are gathered (ovo_frame.depth_range_dev). */
int ovo_depth_filter_batch(const float* depth_dev, int n_frames, int h, int w, float* out_dev, float* ranges_out_dev, void* stream);
/* min / max of the depth values > 0 (what compute_camera_frustum_corners takes from the raw depth, geometry_utils.py:110-111)
 * -> range_out_dev f32 [2]. */
int ovo_depth_range(const float* depth_dev, int64_t n, float* range_out_dev, void* stream);

/* OVO._match_and_track_instances minus the Python bookkeeping (ovo.py:204-229, 240-282):
 * frustum cull (geometry_utils.py:99-129,252-277) + projection/depth match (:26-89) + seg lookup +
 * per-mask instance vote + id assignment to unassigned points, in two streaming passes over the map.
 *   xyz_dev [N,3] f32, ins_ids_dev [N] i32 updated IN PLACE (-1 = unassigned),
 *   next_ins_id: in = OVO.next_ins_id, out = advanced by the number of new instances,
 *   votes_host [n_masks] filled on return (this call synchronises `stream` once),
 *   n_matched_host = len(matched_points_idxs).
 * The matched (point, mask) list of this keyframe is kept inside the handle in slot `kf_slot`
 * (0 <= kf_slot < 64) for a later ovo_map_fuse_dense. */
int ovo_map_associate(ovo_map_t* map, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N, const ovo_frame* frame,
                      int* next_ins_id, ovo_vote_row* votes_host, int* n_matched_host, int kf_slot, void* stream);
/* ovo_map_associate in two halves: `launch` enqueues everything (kernels + the read-back of the rows) and returns at once,
 * `wait` is the one host synchronisation and fills votes_host / n_matched / next_ins_id.  Between the two the caller does its own
 * host work (the drop-in's per-keyframe bookkeeping of the PREVIOUS keyframe overlaps the GPU's association of this one).  One
 * association in flight per handle. */
int ovo_map_associate_launch(ovo_map_t* map, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N, const ovo_frame* frame,
                             int next_ins_id, int kf_slot, void* stream);
int ovo_map_associate_wait(ovo_map_t* map, int* next_ins_id, ovo_vote_row* votes_host, int* n_matched_host);
/* The same association split in two for a map SHARDED over several GPUs: every rank calls
 * ovo_map_vote on its own points (no host sync), the ranks sum the returned table (n_masks*(n_ins+1) vote counts
 * followed by one n_matched counter; the call returns that length) with an all-reduce, then every rank calls
 * ovo_map_apply with the summed table: all ranks take identical decisions and update their own points. */
int ovo_map_vote(ovo_map_t* map, const float* xyz_dev, const int32_t* ins_ids_dev, int64_t N, const ovo_frame* frame,
                 int n_ins, int32_t* table_out_dev, int kf_slot, void* stream);
int ovo_map_apply(ovo_map_t* map, const int32_t* table_in_dev, int32_t* ins_ids_dev, int* next_ins_id,
                  ovo_vote_row* votes_host, int* n_matched_host, void* stream);
/* ovo_map_associate for a BATCH of keyframes (same results as n_frames consecutive calls, ids bit for bit): the map is
 * streamed ONCE for all keyframes (the geometry of a keyframe does not depend on the instance ids), the votes are then taken
 * keyframe by keyframe on the device with next_ins_id left there, and the host reads every keyframe's rows back with ONE
 * synchronisation.  frames / kf_slots host arrays [n_frames] (n_frames <= 64; kf_slots may be NULL: no dense fusion later),
 * votes_host [n_frames][votes_stride], n_matched_host [n_frames], mask_ins_out_dev (optional) i32 [n_frames][votes_stride]:
 * the instance id given to every mask (-1 = none), for callers that keep going on the device. */
int ovo_map_associate_batch(ovo_map_t* map, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N, const ovo_frame* frames,
                            int n_frames, const int* kf_slots, int* next_ins_id, ovo_vote_row* votes_host, int votes_stride,
                            int* n_matched_host, int32_t* mask_ins_out_dev, void* stream);
/* The same batch in stages, for a map SHARDED over several GPUs: every rank runs begin on its own points, then
 * per keyframe f in order: ovo_map_batch_vote (votes of its points; *table_dev / *table_len = the keyframe's table
 * [n_matched, 0, 0, 0 | n_masks x (n_ins+1) vote counts] inside tables_dev) -> the ranks SUM that table (all-reduce, in place) ->
 * ovo_map_batch_decide (identical decisions on every rank, next_ins_id stays on the device); ovo_map_batch_end assigns the
 * last keyframe's ids, reads all rows back and synchronises once.  tables_dev i32 [tables_cap] is caller-owned (NULL: a
 * workspace of the handle); it needs sum_f (4 + max(n_masks_f,1) * (next_ins_id + sum_{g<f} n_masks_g + 1), rounded up to 4) ints. */
int ovo_map_batch_begin(ovo_map_t* map, const float* xyz_dev, const int32_t* ins_ids_dev, int64_t N, const ovo_frame* frames,
                        int n_frames, const int* kf_slots, int next_ins_id, int32_t* tables_dev, int64_t tables_cap, void* stream);
int ovo_map_batch_vote(ovo_map_t* map, int f, int32_t* ins_ids_dev, int32_t** table_dev, int* table_len, void* stream);
int ovo_map_batch_decide(ovo_map_t* map, int f, void* stream);
int ovo_map_batch_end(ovo_map_t* map, int32_t* ins_ids_dev, int* next_ins_id, ovo_vote_row* votes_host, int votes_stride,
                      int* n_matched_host, int32_t* mask_ins_out_dev, void* stream);
/* Keyframe f of the pending batch: its mask count and the DEVICE address of the batch's instance counter (n_ins). */
int ovo_map_batch_info(ovo_map_t* map, int f, int* n_masks, const int32_t** n_ins_dev);

/* Device-side exchange of the vote tables of a sharded map (one process per GPU of one box, peer memory over NVLink/NVSwitch):
 * instead of a host-launched all-reduce per keyframe, ONE kernel per keyframe writes this rank's compact table into every
 * peer's inbox (16-byte peer stores), raises a flag there (st.release.sys), waits for the flags of its own inbox and sums the
 * world tables in place — the result is the same on every rank.  Set-up: every rank creates its exchange (slots = keyframes per
 * batch, table_ints = largest table), publishes its 64-byte IPC handle (ovo_xchg_ipc_handle) to the others by any means
 * (torch.distributed all_gather) and opens theirs (ovo_xchg_open_peers, handles [world][64] in rank order).  All ranks must
 * call ovo_xchg_exchange in the same order with the same slot. */
typedef struct ovo_xchg ovo_xchg_t;
int ovo_xchg_create(int rank, int world, int slots, int64_t table_ints, ovo_xchg_t** out);
int ovo_xchg_ipc_handle(ovo_xchg_t* x, void* handle_out_64);
int ovo_xchg_open_peers(ovo_xchg_t* x, const void* handles_world_x_64);
/* table_dev i32 [n_ints] (16-byte aligned) <- sum over the ranks.  n_ins_dev / n_masks (optional): only the first
 * 4 + max(n_masks,1) * (*n_ins_dev + 1) ints travel (the layout of ovo_map_batch_vote's tables). */
int ovo_xchg_exchange(ovo_xchg_t* x, int32_t* table_dev, int n_ints, const int32_t* n_ins_dev, int n_masks, int slot, void* stream);
void ovo_xchg_destroy(ovo_xchg_t* x);
/* Map growth of a sharded map: packs n freshly mapped points (xyz f32 [n,3], ids i32 [n]) into fixed-size per-shard runs for ONE
 * all-to-all without a host synchronisation: records_out_dev f32 [world * cap_per_dst][4] = (x, y, z, id bits); slot
 * dst * cap_per_dst + k holds this rank's k-th point (creation order) of shard dst = hash of its voxel (cell metres, the rule of
 * ovo_b200.sharding.shard_of_points); unused slots hold (far, far, far, id -1); *overflow_dev += points that did not fit. */
int ovo_route_pack(const float* xyz_dev, const int32_t* ids_dev, int n, int world, float cell, int cap_per_dst, float far_value,
                   float* records_out_dev, int32_t* overflow_dev, void* stream);
/* ovo_map_associate_batch on a shard, the per-keyframe tables summed through `xchg`: the whole batch is one call. */
int ovo_map_associate_batch_sharded(ovo_map_t* map, ovo_xchg_t* xchg, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N,
                                    const ovo_frame* frames, int n_frames, const int* kf_slots, int* next_ins_id,
                                    ovo_vote_row* votes_host, int votes_stride, int* n_matched_host, int32_t* mask_ins_out_dev,
                                    void* stream);
/* Copies the matched list of a slot to the caller: pairs (point index, mask index), n from associate. */
int ovo_map_get_matches(ovo_map_t* map, int kf_slot, int32_t* pairs_dev, int max_pairs, void* stream);

/* Dense per-point running mean (north-star F6; per-point analogue of instance3d.py:19-21).  The bank keeps the mean of the
 * descriptors of the masks a point fell into as TWO bf16 planes: bank_dev = the mean rounded to bf16 (the operand of
 * ovo_query_dense), bank_lo_dev = bf16(mean - bank): together 16-17 significant bits, so the increment of a long stream
 * ((e - f)/count, below half a bf16 ulp of f once count is in the hundreds) is not lost.
 * For every matched point p of slot kf_slot whose mask m has mask_row_dev[m] = r >= 0 (f32 operations, no FMA contraction):
 *   f = bank[p] + bank_lo[p];  e = bf16(feats[r]);  count[p] += 1;  f' = f + (e - 1*f) * (1 / count[p]);
 *   bank[p] = bf16(f');  bank_lo[p] = bf16(f' - bank[p]).
 * bank_dev, bank_lo_dev bf16 [N, D], counts_dev i32 [N], feats_dev f32 [n_rows, D], mask_row_dev i32 [n_masks]. */
int ovo_map_fuse_dense(ovo_map_t* map, int kf_slot, void* bank_dev, void* bank_lo_dev, int32_t* counts_dev, int64_t N, int D,
                       const float* feats_dev, int n_rows, const int32_t* mask_row_dev, int n_masks, void* stream);

/* The same mean for SEVERAL keyframes in one pass over the bank: a point matched in k keyframes of the batch has its rows
 * read once and written once, its k descriptors e_1..e_k (keyframe order) summed in f32:
 *   s = e_1 + ... + e_k;  count += k;  f' = f + (s - k*f) * (1 / count)
 * (equal to k single updates up to the rounding of the two-plane format).  feats_dev f32 [n_rows, D] holds the descriptors of
 * all keyframes; mask_row_dev i32 [n_slots, n_masks] maps (keyframe, mask) to a row of feats_dev or -1.  n_slots <= 64.  The
 * slots are either all filled by ovo_map_associate (match lists) or are consecutive slots of ONE ovo_map_associate_batch. */
int ovo_map_fuse_dense_batch(ovo_map_t* map, const int* kf_slots_host, int n_slots, void* bank_dev, void* bank_lo_dev,
                             int32_t* counts_dev, int64_t N, int D, const float* feats_dev, int n_rows,
                             const int32_t* mask_row_dev, int n_masks, void* stream);

/* Instance-bank running mean, fusion 'avg_pooling' (instance3d.py:19-21,157-189):
 * bank[row[i]] = (bank[row[i]]*cnt + feats[i]) / (cnt+1), not re-normalised. bank f32 [I,D]. */
int ovo_bank_update_mean(float* bank_dev, int32_t* counts_dev, int D, const float* feats_dev,
                         const int32_t* rows_dev, int n, void* stream);

/* The same fusion when a few views are ADDED to instances whose other views are already fused (the common case of
 * Instance3D.update_clip with 'avg_pooling': O(new views) per instance instead of re-averaging every stored view):
 * bank[b] = (n * bank[b] + sum_{i in [i0,i1)} store[idx[i]]) / (n + i1 - i0) for quads_dev i32 [m][4] = (b, n, i0, i1). */
int ovo_bank_add_views(float* bank_dev, int D, const float* store_dev, const int32_t* quads_dev, const int32_t* idx_dev, int m,
                       void* stream);

/* clip_cosine_similarity (clip_utils.py:16-19) on the DENSE map: bank bf16 [N,D] x text f32 [Q,D]^T ->
 * out f32 [N,Q].  tcgen05 GEMM streaming the bank once from HBM. */
int ovo_query_dense(ovo_map_t* map, const void* bank_dev, int64_t N, int D, const float* text_dev, int Q,
                    float* out_dev, void* stream);
/* Same, instance bank (OVO.query / get_objs_clips, ovo.py:495-527): out[i] = bank[rows[i]] . text^T in f32.
 * bank f32 [*,D], rows_dev i32 [I] (NULL = identity), text f32 [Q,D] -> out f32 [I,Q]. */
int ovo_query_instances(const float* bank_dev, const int32_t* rows_dev, int I, int D, const float* text_dev, int Q,
                        float* out_dev, void* stream);
/* OVO._fuse_masks_with_same_ins_id (ovo.py:284-324): out[r] = OR of masks m with group_dev[m] == r (uint8 0/1,
 * [R,H,W]); areas_dev[r] = number of set pixels.  masks uint8 [M,H,W]. */
int ovo_merge_masks(const uint8_t* masks_dev, int M, int H, int W, const int32_t* group_dev, int R, uint8_t* out_dev,
                    int32_t* areas_dev, void* stream);
/* Instance3D.update_clip for a batch of instances (instance3d.py:157-189): instance j fuses the descriptor rows
 * store[idx[off[j]..off[j+1])] into bank[out_rows[j]]; mode 0 avg_pooling (:19-21, not re-normalised),
 * 1 l1_medoid (:9-12), 2 cossim_medoid (:14-17); chosen_dev[j] (optional) = index of the medoid view. */
int ovo_fuse_views(const float* store_dev, int D, const int32_t* idx_dev, const int32_t* off_dev, int n_instances,
                   int mode, float* bank_dev, const int32_t* out_rows_dev, int32_t* chosen_dev, void* stream);
/* Map producer, VanillaMapper.map (ovo/slam/vanilla_mapper.py:46-85): depth pixels not yet explained
 * by a map point (same cull/project/depth test as the association, match_th 0.03, 3x3 erosion of the free mask,
 * every `downscale`-th pixel) are un-projected with c2w and APPENDED at row N of the caller's pre-reserved buffers
 * (xyz f32 [capacity,3], ids i32, ins_ids i32 = -1, colors u8 [capacity,3] optional) — no vstack re-allocation.
 * Synchronises `stream` once to return the number of new points. */
int ovo_map_integrate(ovo_map_t* map, float* xyz_dev, int32_t* ids_dev, int32_t* ins_ids_dev, uint8_t* colors_dev, int64_t N,
                      int64_t capacity, const float* depth_dev, const uint8_t* rgb_dev, int h, int w, const float* c2w,
                      const float* w2c, const float* K, float match_th, int downscale, int k_pool, int next_point_id,
                      int* n_new_host, void* stream);

/* S2 — mask post-processing of the proposals (ovo/utils/segment_utils.py).
 * ovo_mask_nms = masks_update + mask_nms + filter (:173-259): masks uint8 [M,H,W], scores f32 [M] (= stability *
 * predicted_iou) -> keep_dev uint8 [M] in the ORIGINAL mask order.  OVO passes iou_thr 0.8, score_thr 0.7, inner_thr 0.5
 * (mask_generator.py:25-27).
 * ovo_mask2segmap = mask2segmap (:12-27): masks painted in descending stability, earlier masks win overlaps:
 * seg_map_dev i32 [H,W] (-1 = none), maps_out_dev uint8 [M,H,W] = masks in painted order, order_dev i32 [M] = painted
 * position -> input index. */
int ovo_mask_nms(const uint8_t* masks_dev, const float* scores_dev, int M, int H, int W, float iou_thr, float score_thr,
                 float inner_thr, uint8_t* keep_dev, void* stream);
int ovo_mask2segmap(const uint8_t* masks_dev, const float* stability_dev, int M, int H, int W, int32_t* seg_map_dev,
                    uint8_t* maps_out_dev, int32_t* order_dev, void* stream);


/* ------------------------------------------------------------------------------------------------
 * S1 — SAM-2 mask proposal (image path): thirdParty/segment-anything-2/sam2/
 *   utils/transforms.py:15-40 (resize 1024 + normalize), modeling/backbones/hieradet.py:39-291 (Hiera trunk),
 *   modeling/backbones/image_encoder.py:45-134 (FPN neck), modeling/sam2_base.py:467-479 (conv_s0/s1),
 *   sam2_image_predictor.py:86-127,337-432 (set_image / _predict), modeling/sam/prompt_encoder.py:81-182,
 *   modeling/sam/transformer.py:44-286, modeling/sam/mask_decoder.py:110-245,
 *   automatic_mask_generator.py:170-375 + utils/amg.py (grid prompts, filters, box NMS),
 *   and OVO's wrapper ovo/entities/mask_generator.py:102-120 (segment).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ovo_sam ovo_sam_t;

/* One Hiera block (MultiScaleBlock, hieradet.py:84-166).  Matrices bf16 [out,in] as nn.Linear stores them, vectors f32. */
typedef struct {
  int dim, dim_out, heads, window /* 0 = global */, q_pool /* MaxPool2d(2,2) on q + shortcut */, grid_in;
  const float* norm1_w; const float* norm1_b;
  const void* qkv_w;    const float* qkv_b;     /* [3*dim_out, dim] rows [q;k;v], each head-major */
  const void* proj_w;   const float* proj_b;    /* [dim_out, dim_out] */
  const float* norm2_w; const float* norm2_b;
  const void* fc1_w;    const float* fc1_b;     /* [4*dim_out, dim_out] */
  const void* fc2_w;    const float* fc2_b;     /* [dim_out, 4*dim_out] */
  const void* short_w;  const float* short_b;   /* [dim_out, dim] shortcut projection of transition blocks, else NULL */
} ovo_hiera_block;

/* sam/transformer.py Attention: q/k/v/out projections, bf16 [out,in] + f32 bias */
typedef struct {
  const void* q_w; const float* q_b; const void* k_w; const float* k_b;
  const void* v_w; const float* v_b; const void* o_w; const float* o_b;
} ovo_sam_attn;

/* TwoWayAttentionBlock (sam/transformer.py:137-212) */
typedef struct {
  ovo_sam_attn self_attn, t2i, i2t;
  const float* norm_w[4]; const float* norm_b[4];
  const void* mlp0_w; const float* mlp0_b; const void* mlp1_w; const float* mlp1_b;
} ovo_sam_dec_layer;

typedef struct {
  int image_size;          /* 1024 */
  int n_blocks;
  int embed_dim;           /* 144 */
  int stage_end[4];        /* block index closing each stage (hieradet.py:193) */
  int decoder_depth;       /* 2 */
  float trunk_ln_eps;      /* 1e-6 */
  int max_batch;           /* frames one trunk pass may take (ovo_sam_set_images); 0 or 1 = single frame */
} ovo_sam_cfg;

typedef struct {
  /* trunk */
  const void* patch_w; int patch_kpad; const float* patch_b;   /* bf16 [embed, kpad]: conv 7x7 s4 p3 flattened (c,ky,kx), zero padded */
  const float* pos;                  /* f32 [g*g, embed]: Hiera._get_pos_embed (bicubic bkg + tiled window), weights only */
  const ovo_hiera_block* blocks;     /* host array [n_blocks] */
  /* neck (image_encoder.py:102-134) with conv_s0/conv_s1 (sam2_base.py:467-479) folded into the two lateral convs that feed
   * them, and no_mem_embed (sam2_image_predictor.py:118-121) folded into the bias of the 64x64 level */
  const void* neck3_w; const float* neck3_b;   /* [256, C_stage4]  (top level, only feeds the top-down path) */
  const void* neck2_w; const float* neck2_b;   /* [256, C_stage3]  -> image_embed */
  const void* s1_w;    const float* s1_b;      /* [64,  C_stage2]  -> feat_s1 */
  const void* s0_w;    const float* s0_b;      /* [32,  C_stage1]  -> feat_s0 */
  /* prompt encoder (prompt_encoder.py) */
  const float* gauss;          /* f32 [2,128] positional_encoding_gaussian_matrix */
  const float* point_embed;    /* f32 [256] point_embeddings[1] (foreground point) */
  const float* not_a_point;    /* f32 [256] */
  const float* dense_pe;       /* f32 [64*64, 256] get_dense_pe(), weights only */
  const float* no_mask_embed;  /* f32 [256] */
  /* mask decoder (mask_decoder.py) */
  const float* out_tokens;     /* f32 [6,256]: obj_score_token, iou_token, mask_tokens[0..3] */
  const ovo_sam_dec_layer* layers;   /* host array [decoder_depth] */
  ovo_sam_attn final_attn; const float* norm_final_w; const float* norm_final_b;
  const void* up0_w; const float* up0_b;       /* ConvTranspose2d 256->64 k2 s2 as bf16 [(ky*2+kx)*64+o, 256]; bias tiled [256] */
  const float* up_ln_w; const float* up_ln_b;  /* LayerNorm2d(64), eps 1e-6 */
  const void* up1_w; const float* up1_b;       /* ConvTranspose2d 64->32 k2 s2 as bf16 [(ky*2+kx)*32+o, 64]; bias tiled [128] */
  const void* hyper_w[4][3]; const float* hyper_b[4][3];   /* output_hypernetworks_mlps[i].layers[j] */
  const void* iou_w[3]; const float* iou_b[3];             /* iou_prediction_head (sigmoid output) */
} ovo_sam_weights;

typedef struct {
  int points_per_side;        /* ovo.yaml:32 -> 16 */
  float pred_iou_thresh;      /* segment_utils.py:298 <- sam.nms_iou_th (0.8) */
  float stability_thresh;     /* :299 <- sam.stability_score_th (0.95) */
  float stability_offset;     /* 1.0 (automatic_mask_generator.py:44) */
  float box_nms_thresh;       /* 0.7 (:46) */
  float nms_iou_th, nms_score_th, nms_inner_th;   /* OVO's second NMS, mask_generator.py:25-27 (0.8, 0.7, 0.5) */
} ovo_amg_params;

int ovo_sam_create(const ovo_sam_cfg* cfg, const ovo_sam_weights* w, int max_h, int max_w, int max_prompts, ovo_sam_t** out);
void ovo_sam_destroy(ovo_sam_t* sam);
/* SAM2ImagePredictor.set_image: rgb uint8 [H,W,3] -> image embedding + high-res features kept inside the handle.
 * Optional taps (may be NULL): pixels f32 [3,S,S]; embed f32 [g*g,256] (token major); feat_s0 f32 [16*g*g,32];
 * feat_s1 f32 [4*g*g,64], g = image_size/16.  n_blocks < 0 = all; block_out (optional) receives the f32 token grid
 * [grid^2, dim] after the last executed block (layer-by-layer parity tests). */
int ovo_sam_set_image(ovo_sam_t* sam, const uint8_t* rgb_dev, int H, int W, float* pixels_out_dev, float* embed_out_dev,
                      float* feat_s0_out_dev, float* feat_s1_out_dev, int n_blocks, float* block_out_dev, void* stream);
/* The same for n <= cfg.max_batch frames [n,H,W,3] in ONE trunk pass (replay / MaskGenerator.precompute: the stage-3 GEMMs of
 * a single frame have only 4096 token rows); the decoder then works on the frame chosen with ovo_sam_select_image. */
int ovo_sam_set_images(ovo_sam_t* sam, const uint8_t* rgb_dev, int n, int H, int W, void* stream);
int ovo_sam_select_image(ovo_sam_t* sam, int index);
/* Test tap: run the trunk from already normalised pixels f32 [3,S,S] instead of the resize. */
int ovo_sam_set_pixels(ovo_sam_t* sam, const float* pixels_dev, float* embed_out_dev, float* feat_s0_out_dev,
                       float* feat_s1_out_dev, int n_blocks, float* block_out_dev, void* stream);
/* SAM2ImagePredictor._predict, one foreground point per prompt, multimask_output=True: points f32 [P,2] in model-frame
 * pixels (transforms.py:59-65) -> low_res f32 [P,3,4g,4g] mask logits (not clamped), iou f32 [P,3]. */
int ovo_sam_predict(ovo_sam_t* sam, const float* points_dev, int P, float* low_res_out_dev, float* iou_out_dev, void* stream);
/* SAM2AutomaticMaskGenerator._process_batch/_process_crop filters for the whole-image crop
 * (automatic_mask_generator.py:251-375) on given logits: bilinear up-sampling to HxW (transforms.py:117), predicted-IoU
 * filter, stability score (utils/amg.py:158-178), threshold, boxes (:305-348), box NMS (torchvision batched_nms).
 *   low_res f32 [P,3,h,w], iou f32 [P,3]  ->  masks_out uint8 [K,H,W] (0/1) in the reference's order (descending predicted
 *   IoU), iou_out/stab_out f32 [K], boxes_out i32 [K,4] XYXY, src_out i32 [K] = index into the flattened [P*3] list.
 * Synchronises `stream` once; returns K in *n_out (K <= max_out, else OVO_E_INVALID). */
int ovo_sam_postprocess(ovo_sam_t* sam, const float* low_res_dev, const float* iou_dev, int P, int h, int w, int H, int W,
                        const ovo_amg_params* prm, uint8_t* masks_out_dev, float* iou_out_dev, float* stab_out_dev,
                        int32_t* boxes_out_dev, int32_t* src_out_dev, int max_out, int* n_out, void* stream);
/* Measurement / test aid: after this call ovo_sam_generate(_batch) still runs the whole network but hands low_dev f32
 * [n_prompts,3,4g,4g] / iou_dev f32 [n_prompts,3] (caller-owned, kept alive) to the AMG post-processing instead of the decoder's
 * own logits — with random-init weights the stock thresholds reject every proposal, so a benchmark without checkpoints would
 * time an empty NMS / seg-map stage.  NULL, NULL, 0 removes the override. */
int ovo_sam_override_logits(ovo_sam_t* sam, const float* low_dev, const float* iou_dev, int n_prompts);
/* MaskGenerator.segment (ovo/entities/mask_generator.py:102-120) end to end: set_image, grid prompts, decoder, AMG filters,
 * OVO's masks_update (segment_utils.py:173-259) and mask2segmap (:12-27).
 *   -> seg_map i32 [H,W] (-1 = none), masks_out uint8 [M,H,W] in painted order, M in *n_masks (M <= max_masks). */
int ovo_sam_generate(ovo_sam_t* sam, const uint8_t* rgb_dev, int H, int W, const ovo_amg_params* prm, int32_t* seg_map_dev,
                     uint8_t* masks_out_dev, int max_masks, int* n_masks, void* stream);

/* ovo_sam_generate for n_frames <= cfg.max_batch frames with one batched trunk pass: seg_maps i32 [n,H,W], masks_out uint8
 * [n, max_masks, H, W], n_masks_host[n]. */
int ovo_sam_generate_batch(ovo_sam_t* sam, const uint8_t* rgb_dev, int n_frames, int H, int W, const ovo_amg_params* prm,
                           int32_t* seg_maps_dev, uint8_t* masks_out_dev, int max_masks, int* n_masks_host, void* stream);

/* OVO.classify_instances (ovo.py:486-491): argmax over queries + threshold. sim f32 [n,Q] ->
 * cls i32 [n] (-1 if max <= th), conf f32 [n] (0 if max <= th). */
int ovo_classify(const float* sim_dev, int64_t n, int Q, float th, int32_t* cls_dev, float* conf_dev, void* stream);

/* Per-launch timing for bench.py's roofline: after ovo_profile_begin() every kernel launch of this library is
 * bracketed by CUDA events on its stream (the encoder then runs eagerly instead of replaying its CUDA graph);
 * ovo_profile_report() stops, synchronises and sums per kernel class: 0 gemm, 1 attention, 2 layernorm,
 * 3 preprocess, 4 region pooling, 5 association, 6 dense fusion, 7 query, 8 other.  Returns the class count. */
void ovo_profile_begin(void);
int ovo_profile_report(int n_classes, float* ms_host, double* flops_host, double* bytes_host, int* counts_host);

/* Tuning aid: thread-block cluster size (TMA multicast of the weight tile) used by the wide GEMM tiles:
 * 0 = automatic (default), 1, 2 or 4. */
void ovo_set_gemm_cluster(int cluster_size);

/* Measurement tap: while `buf_dev` (device memory, 8 x 16 x 16 int64, zeroed by the caller) is set, CTA 0 of every
 * attention_fwd_kernel launch writes clock64() stamps of its softmax thread 0 and of its MMA thread:
 * buf[(item % 8) * 256 + block * 16 + phase]; null switches it off (tools/attn_trace.py reads it). */
void ovo_attn_trace(long long* buf_dev);

/* Test tap for the GEMM machinery: C[M,N] f32 = A[M,K] bf16 . B[N,K]^T bf16 (+bias f32 [N]). */
int ovo_gemm_bf16(const void* A_dev, int lda, const void* B_dev, int ldb, int M, int N, int K,
                  const float* bias_dev, float* C_dev, int ldc, int force_bn, void* stream);

/* Measurement tap: average time of `iters` launches of the GEMM with a fused epilogue (0 f32, 1 bf16, 2 bf16+GELU,
 * 3 f32+residual, 6 bf16+ReLU) on synthetic operands, CUDA events on `stream`. */
int ovo_gemm_bench(int epi, int M, int N, int K, int force_bn, int iters, float* ms_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OVO_B200_H */
