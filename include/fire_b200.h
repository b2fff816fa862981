/* fire_b200.h - C ABI of libfire_b200.so, the sm_100a identification hot path for FIRE.
 *
 * The reference (IvanYachUkr/FACE-Identification-in-Real-time-Environments-FIRE) is pure Python;
 * it has no FFI of its own.  Its operator boundary for this path is three Python surfaces, and
 * each entry point below replaces the third-party native call that surface makes today:
 *
 *   fire_preprocess        <- cv2.resize(INTER_AREA) + /255      modules/encoder.py:19-27
 *                             crop slice                         modules/face_recognition.py:412-420
 *   fire_align_warp        <- cv2.getAffineTransform + warpAffine  yunet_face_detector.py:135-160 (enrol path)
 *   fire_ingest_f32        <- the NHWC float batch handed to     modules/encoder.py:16-17
 *   fire_facenet_*         <- onnxruntime InferenceSession.run   facenet_gpu.py:72,127
 *   fire_knn_create/add    <- hnswlib Index.init_index/add_items modules/hnsw_manager.py:29,127,137
 *   fire_knn_search        <- hnswlib Index.knn_query            modules/hnsw_manager.py:147,237
 *   fire_knn_merge         <- (new) multi-GPU partial top-k merge after the NCCL all-gather
 *
 * Conventions
 *   - plain C, no exceptions; every function returns FIRE_OK (0) or a negative FIRE_ERR_* code and
 *     leaves a message retrievable with fire_last_error() (thread-local).
 *   - pointers are DEVICE pointers owned by the caller unless the parameter is named host_*.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     and the functions do not synchronise, except the *_host convenience calls, which copy
 *     host<->device and return when the result is in the host buffers.
 *   - one handle lives on the device that was current at *_create; a handle is not re-entrant.
 *   - there is no CPU fallback: without a usable sm_100 device every compute call fails.
 */
#ifndef FIRE_B200_H_
#define FIRE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIRE_OK 0
#define FIRE_ERR_ARG (-1)         /* bad argument */
#define FIRE_ERR_CUDA (-2)        /* a CUDA runtime / driver call failed */
#define FIRE_ERR_STATE (-3)       /* handle in the wrong state (e.g. index full, k > count) */
#define FIRE_ERR_UNSUPPORTED (-4) /* device is not sm_100, or a size outside the kernels' range */

typedef void* fire_stream_t;
typedef struct fire_net fire_net_t;
typedef struct fire_knn fire_knn_t;

/* ---- library ---------------------------------------------------------------------------- */
int fire_init(int device);                 /* cudaSetDevice + checks compute capability 10.x */
const char* fire_last_error(void);
const char* fire_version(void);
uint64_t fire_launch_count(void);          /* number of kernels this library has launched so far */

/* ---- K1: crop + resize + normalise  (modules/encoder.py:19-27, face_recognition.py:412-420) */
#define FIRE_PRE_REFERENCE 0 /* cv2 INTER_AREA semantics on uint8, re-quantised to uint8, then /255 */
#define FIRE_PRE_NORTHSTAR 1 /* float half-pixel bilinear + per-crop prewhiten (not in the reference) */
#define FIRE_PRE_FLAG_SWAP_RB 16 /* OR into mode: reverse channel order (BGR<->RGB) while sampling */

/* frames      : uint8 HWC3 images packed in one device allocation
 * frame_desc  : int64 [n_frames][4] = {byte offset into frames, H, W, row stride in bytes}
 * boxes_xywh  : int32 [n_boxes][4]  = x, y, w, h exactly as the detector/tracker emits them; the
 *               reference's clamp rule is applied here: each of x,y,w,h = max(0, .) independently,
 *               then the far edge is clipped to the frame (numpy slicing).  Empty crops produce an
 *               all-zero output and status 1 in box_status.
 * box_frame   : int32 [n_boxes] frame index of every box
 * out_f16     : fp16 [n_boxes][80][80][16]  network input, space-to-depth: position (Y, X) holds the 2 x 2 pixel
 *               block (2Y + dy, 2X + dx) as channels (dy * 2 + dx) * 3 + c, channels 12..15 zero; pixel scale
 *               (0..255 == 0..1 in the reference).                     (may be NULL)
 * out_f32     : float [n_boxes][160][160][3] exactly what preprocess_for_encoder returns (may be NULL)
 * box_status  : int32 [n_boxes] 0 = ok, 1 = empty crop (may be NULL)
 */
int fire_preprocess(const uint8_t* frames, const int64_t* frame_desc, int n_frames, const int32_t* boxes_xywh,
                    const int32_t* box_frame, int n_boxes, int mode, void* out_f16, float* out_f32,
                    int32_t* box_status, fire_stream_t stream);

/* ROI upload for frame-sized inputs (BASELINE configs[4]): HOST-side packing of the crop rectangles of every box (the
 * crop rule above applied on the host, like the reference's image[y:y+h, x:x+w]) into one pinned staging buffer, so that only
 * the pixels a crop reads cross PCIe instead of whole frames.  host_packed receives, from offset 0, the tables
 * int64 frame_desc[n][4] | int32 boxes[n][4] | int32 box_frame[n] (fire_roi_meta_bytes(n) bytes in all) and then the
 * rectangles (rows padded to 16 bytes).  After ONE copy of *bytes_used bytes to a device buffer D, call
 * fire_preprocess(D, (int64*)D, n, (int32*)(D + 32 n), (int32*)(D + 48 n), n, ...): results are bit-identical to passing the
 * whole frames.  n_threads worker threads do the row copies (staging only - there is no compute here). */
size_t fire_roi_meta_bytes(int n_boxes);
int fire_pack_rois_host(const uint8_t* host_frames, const int64_t* host_frame_desc, int n_frames,
                        const int32_t* host_boxes_xywh, const int32_t* host_box_frame, int n_boxes,
                        uint8_t* host_packed, size_t host_packed_bytes, size_t* bytes_used, int n_threads);

/* The same upload with the copy engine doing the gather (no host copy): tables -> host_tables (pinned, fire_roi_meta_bytes(n)
 * bytes, untouched until `stream` has passed this call) -> one cudaMemcpyAsync; every crop rectangle goes straight from the
 * pinned frames to its place in dev_packed with one cudaMemcpy2DAsync.  Same device layout as fire_pack_rois_host + copy. */
int fire_upload_rois_dma(const uint8_t* host_frames, const int64_t* host_frame_desc, int n_frames,
                         const int32_t* host_boxes_xywh, const int32_t* host_box_frame, int n_boxes,
                         uint8_t* host_tables, uint8_t* dev_packed, size_t dev_packed_bytes, size_t* bytes_used,
                         fire_stream_t stream);

/* Aligned crop of the enrol path: cv2.warpAffine(image, M, (160, 160)) (INTER_LINEAR, constant 0 border) bit for bit,
 * optionally followed by the reference's [:, :, ::-1]            yunet_face_detector.py:135-165 (and the MediaPipe /
 * RetinaFace twins).  matrices: double [n_faces][6], the FORWARD 2x3 matrix cv2.getAffineTransform returns (the host
 * shim computes it with the same LU elimination, fire_b200/preprocess.py); face_frame: int32 [n_faces].
 * out_u8: uint8 [n_faces][160][160][3] (what extract_faces returns), out_f16: the network input of those crops
 * (layout above); either may be NULL. */
int fire_align_warp(const uint8_t* frames, const int64_t* frame_desc, int n_frames, const double* matrices,
                    const int32_t* face_frame, int n_faces, int swap_rb, uint8_t* out_u8, void* out_f16,
                    fire_stream_t stream);

/* float NHWC [B][160][160][3] in the reference's [0,1] scale -> fp16 [B][80][80][16] network input (layout above) */
int fire_ingest_f32(const float* in_nhwc3, int B, void* out_f16, fire_stream_t stream);

/* ---- K2: FaceNet (Inception-ResNet-v1) forward  (facenet_gpu.py:116-129) -------------------- */
/* host_blob: plan + folded fp16 weights as produced by fire_b200.weights.pack(). */
int fire_facenet_create(const void* host_blob, size_t bytes, fire_net_t** out);
int fire_facenet_destroy(fire_net_t* net);
int fire_facenet_dim(const fire_net_t* net);                    /* 128 or 512 */
size_t fire_facenet_workspace(const fire_net_t* net, int B);    /* bytes of scratch forward() needs */
double fire_facenet_flops(const fire_net_t* net);               /* algorithmic FLOP per image */
int fire_facenet_num_ops(const fire_net_t* net);
/* Kernel launches one forward() enqueues for the plan's ops (without the optional L2-norm launch): ops inside a fused
 * chain (block17_fused_kernel, block35_fused_kernel) share one launch. */
int fire_facenet_num_launches(const fire_net_t* net);
/* in_f16: fp16 [B][80][80][16] (fire_preprocess / fire_ingest_f32 output); out_raw: float [B][D] un-normalised (what encode() returns);
 * out_l2: float [B][D] rows divided by their L2 norm (face_recognition.py:225-229), may be NULL. */
int fire_facenet_forward(fire_net_t* net, const void* in_f16, int B, float* out_raw, float* out_l2,
                         void* workspace, size_t ws_bytes, fire_stream_t stream);
/* Profiling aid: runs forward with a CUDA-event pair around every op; host_ms[n_ops] gets the
 * per-op device time, host_flops[n_ops] the per-op algorithmic FLOP (0 for pools).  Synchronises.
 * A fused chain (fire_facenet_num_launches) is reported on its first op: time and FLOP of the whole launch there, 0 on the rest. */
int fire_facenet_profile(fire_net_t* net, const void* in_f16, int B, void* workspace, size_t ws_bytes,
                         float* host_ms, double* host_flops, int n_ops, fire_stream_t stream);
/* Debug aid: copy an internal activation buffer (index into the plan's buffer table) to the host as fp16.  Buffers that
 * only exist inside a fused chain (the branch tensors of Block17 / Block35) are never written unless the chain runs layer
 * by layer (FIRE_B200_FUSE17=0 / FIRE_B200_FUSE35=0); the arena also recycles dead buffers. */
int fire_facenet_read_buffer(fire_net_t* net, int buf, int B, const void* in_f16, const void* workspace,
                             void* host_out, size_t bytes);

/* ---- K3: exact cosine top-k over the enrolled gallery  (modules/hnsw_manager.py:135-149) ---- */
int fire_knn_create(int D, size_t capacity, fire_knn_t** out);
int fire_knn_destroy(fire_knn_t* h);
int fire_knn_reset(fire_knn_t* h);                                  /* count = 0 */
size_t fire_knn_count(const fire_knn_t* h);
size_t fire_knn_capacity(const fire_knn_t* h);
int fire_knn_dim(const fire_knn_t* h);
/* Append n rows; each is normalised with 1/(||x||+1e-30) like hnswlib's cosine space. */
int fire_knn_add(fire_knn_t* h, const float* rows, size_t n, fire_stream_t stream);
int fire_knn_add_host(fire_knn_t* h, const float* host_rows, size_t n);
/* Copy normalised rows [first, first+n) back to the host (persistence, tests). */
int fire_knn_get_rows_host(fire_knn_t* h, size_t first, size_t n, float* host_out);
/* Exact top-k by (distance asc, id asc), distance = 1 - <q^, g^>; ids = row index + id_offset.
 * Requires 1 <= k <= min(count, 64).  out_dist [Q][k] float, out_ids [Q][k] int64. */
int fire_knn_search(fire_knn_t* h, const float* queries, int Q, int k, int64_t id_offset, float* out_dist,
                    int64_t* out_ids, fire_stream_t stream);
int fire_knn_search_host(fire_knn_t* h, const float* host_queries, int Q, int k, int64_t id_offset,
                         float* host_out_dist, int64_t* host_out_ids);
/* Same search with the STORED rows [first, first+Q) as the queries, used as they are (they were normalised when they were
 * added, which is exactly what knn_query would make of the vector the row came from).  This is the bulk form of the
 * reference's find_similar_embeddings loop in shrink_db_ids (modules/face_recognition.py:265-315, hnsw_manager.py:227-244:
 * one top-50 query per stored embedding); the caller tiles first/Q over the gallery. */
int fire_knn_search_rows(fire_knn_t* h, size_t first, int Q, int k, int64_t id_offset, float* out_dist, int64_t* out_ids,
                         fire_stream_t stream);
/* Shard form for the multi-GPU exchange: results as ONE array of 12-byte records, int32 [Q][k][3] =
 * {distance bits, id low word, id high word}, so a single all-gather carries distances and ids; id = row * id_stride +
 * id_offset (contiguous shards: stride 1, offset = first global row; interleaved shards: stride = world size, offset =
 * rank).  Unlike fire_knn_search, k may exceed the rows this shard holds (an empty or tiny shard of a sharded gallery):
 * missing entries are (FLT_MAX, -1), which fire_knn_merge* sorts last. */
int fire_knn_search_packed(fire_knn_t* h, const float* queries, int Q, int k, int64_t id_offset, int64_t id_stride,
                           void* out_packed, fire_stream_t stream);
/* Merge G partial results (dists/ids laid out [G][Q][k], each row ascending) into the global top-k. */
int fire_knn_merge(const float* dists, const int64_t* ids, int Q, int k, int G, float* out_dist,
                   int64_t* out_ids, fire_stream_t stream);
/* Same merge over packed records [G][Q][k][3] as written by fire_knn_search_packed and gathered with one all-gather. */
int fire_knn_merge_packed(const void* packed, int Q, int k, int G, float* out_dist, int64_t* out_ids,
                          fire_stream_t stream);
/* Counters: queries answered so far / queries that needed the exact fp32 fallback scan. */
int fire_knn_stats(fire_knn_t* h, uint64_t* host_queries_total, uint64_t* host_queries_fallback);
/* host_out4 = {queries answered, queries whose merged-list proof failed (exact re-rank of every split's candidates),
 * of those: queries that then scanned ONE gallery split, queries that scanned the whole shard}. */
int fire_knn_stats_ex(fire_knn_t* h, uint64_t* host_out4);
/* Test hook: widen the fp16-filter safety margin (default: 3e-5 on top of the measured fp16 rounding bound) to force the fallback path. */
int fire_knn_set_margin(fire_knn_t* h, float eps);

#ifdef __cplusplus
}
#endif
#endif /* FIRE_B200_H_ */
