// roi_upload.cu - host side of the ROI upload for frame-sized inputs (BASELINE configs[4]).
//
// The reference crops on the host (modules/face_recognition.py:412-420: x, y, w, h = max(0, .) each, then
// image[y:y+h, x:x+w]) and only the crop ever reaches the encoder.  Uploading whole 1080p frames (6.2 MB each) to
// use ~150 KB per face made the frames path PCIe-bound in round 1.  fire_pack_rois_host applies the same crop rule
// on the host, copies ONLY the crop rectangles into one pinned staging buffer (a few worker threads, plain row
// memcpy - staging, not compute) and writes, at the head of the same buffer, the descriptor tables fire_preprocess
// needs when every rectangle is its own small "frame".  One cudaMemcpyAsync of `bytes_used` bytes then moves
// everything; on the device the tables and rectangles are addressed at the same offsets:
//
//   [0, 32 n)          int64 frame_desc[n][4] = {byte offset of rectangle i, rows, cols, row pitch}
//   [32 n, 48 n)       int32 boxes[n][4]      = {0, 0, cols, rows}        (the whole rectangle)
//   [48 n, 52 n)       int32 box_frame[n]     = i
//   [meta_bytes, ...)  the rectangles, rows padded to a multiple of 16 bytes, each starting on a 256-byte boundary
//
// Empty crops (box outside the frame, zero area) get a 0 x 0 descriptor: fire_preprocess reports status 1 for them,
// exactly as it does for an empty crop of a full frame.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "fire_internal.h"

using namespace fire;

namespace {
struct Roi {
  const uint8_t* src;
  long long src_stride;
  int rows, cols;
  long long dst_off, dst_pitch;
};
}  // namespace

// Layout shared by both upload routes: tables at the head, rectangles behind them (see the file header).
static int plan_rois(const uint8_t* host_frames, const int64_t* host_frame_desc, int n_frames, const int32_t* host_boxes_xywh,
                     const int32_t* host_box_frame, int n_boxes, uint8_t* tables, size_t capacity, std::vector<Roi>& rois, size_t* bytes_used,
                     const char* who);

extern "C" size_t fire_roi_meta_bytes(int n_boxes) {
  return (static_cast<size_t>(std::max(n_boxes, 0)) * 52 + 255) & ~static_cast<size_t>(255);
}

extern "C" int fire_pack_rois_host(const uint8_t* host_frames, const int64_t* host_frame_desc, int n_frames,
                                   const int32_t* host_boxes_xywh, const int32_t* host_box_frame, int n_boxes,
                                   uint8_t* host_packed, size_t host_packed_bytes, size_t* bytes_used, int n_threads) {
  if (!host_frames || !host_frame_desc || !host_boxes_xywh || !host_box_frame || !host_packed || !bytes_used)
    return fail(FIRE_ERR_ARG, "fire_pack_rois_host: NULL argument");
  if (n_boxes <= 0 || n_frames <= 0) return fail(FIRE_ERR_ARG, "fire_pack_rois_host: n_boxes=%d n_frames=%d", n_boxes, n_frames);
  if ((reinterpret_cast<uintptr_t>(host_packed) & 15) != 0) return fail(FIRE_ERR_ARG, "fire_pack_rois_host: staging buffer must be 16-byte aligned");
  std::vector<Roi> rois;
  const int rc = plan_rois(host_frames, host_frame_desc, n_frames, host_boxes_xywh, host_box_frame, n_boxes, host_packed, host_packed_bytes, rois,
                           bytes_used, "fire_pack_rois_host");
  if (rc != FIRE_OK) return rc;
  std::atomic<int> next{0};
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < n_boxes; i = next.fetch_add(1)) {
      const Roi& r = rois[static_cast<size_t>(i)];
      uint8_t* d = host_packed + r.dst_off;
      const size_t row_bytes = static_cast<size_t>(r.cols) * 3;
      for (int yy = 0; yy < r.rows; ++yy) {
        memcpy(d + static_cast<long long>(yy) * r.dst_pitch, r.src + static_cast<long long>(yy) * r.src_stride, row_bytes);
        if (row_bytes < static_cast<size_t>(r.dst_pitch)) memset(d + static_cast<long long>(yy) * r.dst_pitch + row_bytes, 0, static_cast<size_t>(r.dst_pitch) - row_bytes);
      }
    }
  };
  const int nt = std::max(1, std::min(n_threads, std::min(n_boxes, 16)));
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return FIRE_OK;
}

static int plan_rois(const uint8_t* host_frames, const int64_t* host_frame_desc, int n_frames, const int32_t* host_boxes_xywh,
                     const int32_t* host_box_frame, int n_boxes, uint8_t* tables, size_t capacity, std::vector<Roi>& rois, size_t* bytes_used,
                     const char* who) {
  const size_t meta = fire_roi_meta_bytes(n_boxes);
  if (capacity < meta) return fail(FIRE_ERR_ARG, "%s: staging buffer smaller than the tables", who);
  int64_t* desc = reinterpret_cast<int64_t*>(tables);
  int32_t* boxes = reinterpret_cast<int32_t*>(tables + static_cast<size_t>(n_boxes) * 32);
  int32_t* bframe = reinterpret_cast<int32_t*>(tables + static_cast<size_t>(n_boxes) * 48);
  rois.assign(static_cast<size_t>(n_boxes), Roi());
  size_t off = meta;
  for (int i = 0; i < n_boxes; ++i) {
    const int f = host_box_frame[i];
    if (f < 0 || f >= n_frames) return fail(FIRE_ERR_ARG, "%s: box %d names frame %d of %d", who, i, f, n_frames);
    const int64_t* fd = host_frame_desc + 4 * static_cast<long long>(f);
    const int H = static_cast<int>(fd[1]), W = static_cast<int>(fd[2]);
    const int32_t* b = host_boxes_xywh + 4 * static_cast<long long>(i);
    // the crop rule of face_recognition.py:412-420 + numpy slice clipping (same arithmetic as crop_geometry in preprocess.cu)
    const int x = std::max(0, b[0]), y = std::max(0, b[1]), w = std::max(0, b[2]), h = std::max(0, b[3]);
    const int x1 = std::min(W, x + w), y1 = std::min(H, y + h), x0 = std::min(x, W), y0 = std::min(y, H);
    Roi& r = rois[static_cast<size_t>(i)];
    r.cols = std::max(0, x1 - x0); r.rows = std::max(0, y1 - y0);
    if (r.cols == 0 || r.rows == 0) r.cols = r.rows = 0;
    r.src = host_frames + fd[0] + static_cast<long long>(y0) * fd[3] + static_cast<long long>(x0) * 3;
    r.src_stride = fd[3];
    r.dst_pitch = (static_cast<long long>(r.cols) * 3 + 15) & ~15ll;
    r.dst_off = static_cast<long long>(off);
    off += (static_cast<size_t>(r.dst_pitch) * r.rows + 255) & ~static_cast<size_t>(255);
    if (off > capacity) return fail(FIRE_ERR_ARG, "%s: staging buffer of %zu bytes is too small (box %d needs %zu)", who, capacity, i, off);
    desc[4 * i] = r.dst_off; desc[4 * i + 1] = r.rows; desc[4 * i + 2] = r.cols; desc[4 * i + 3] = r.dst_pitch;
    boxes[4 * i] = 0; boxes[4 * i + 1] = 0; boxes[4 * i + 2] = r.cols; boxes[4 * i + 3] = r.rows;
    bframe[i] = i;
  }
  *bytes_used = off;
  return FIRE_OK;
}

// The same upload with the copy engine doing the gather: no host copy at all.  The tables are written into
// `host_tables` (pinned, fire_roi_meta_bytes(n) bytes, must stay untouched until the stream has passed this call) and
// moved with one cudaMemcpyAsync; every rectangle goes straight from the pinned frames to its place in dev_packed with
// one cudaMemcpy2DAsync (source pitch = frame row stride, destination pitch = 16-byte padded row).  The up-to-15 padding
// bytes behind each destination row are not written: the kernel's 16-byte staging loads fetch them, no output reads them.
// host_frames must be pinned for the copies to be asynchronous.
extern "C" int fire_upload_rois_dma(const uint8_t* host_frames, const int64_t* host_frame_desc, int n_frames,
                                    const int32_t* host_boxes_xywh, const int32_t* host_box_frame, int n_boxes,
                                    uint8_t* host_tables, uint8_t* dev_packed, size_t dev_packed_bytes, size_t* bytes_used,
                                    fire_stream_t stream) {
  if (!host_frames || !host_frame_desc || !host_boxes_xywh || !host_box_frame || !host_tables || !dev_packed || !bytes_used)
    return fail(FIRE_ERR_ARG, "fire_upload_rois_dma: NULL argument");
  if (n_boxes <= 0 || n_frames <= 0) return fail(FIRE_ERR_ARG, "fire_upload_rois_dma: n_boxes=%d n_frames=%d", n_boxes, n_frames);
  std::vector<Roi> rois;
  // plan against the DEVICE capacity; only the tables are written on the host
  const int rc = plan_rois(host_frames, host_frame_desc, n_frames, host_boxes_xywh, host_box_frame, n_boxes, host_tables, dev_packed_bytes, rois,
                           bytes_used, "fire_upload_rois_dma");
  if (rc != FIRE_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FIRE_CUDA(cudaMemcpyAsync(dev_packed, host_tables, fire_roi_meta_bytes(n_boxes), cudaMemcpyHostToDevice, st));
  for (int i = 0; i < n_boxes; ++i) {
    const Roi& r = rois[static_cast<size_t>(i)];
    if (r.rows == 0) continue;
    FIRE_CUDA(cudaMemcpy2DAsync(dev_packed + r.dst_off, static_cast<size_t>(r.dst_pitch), r.src, static_cast<size_t>(r.src_stride),
                                static_cast<size_t>(r.cols) * 3, static_cast<size_t>(r.rows), cudaMemcpyHostToDevice, st));
  }
  return FIRE_OK;
}
