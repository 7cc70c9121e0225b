// polyred_cuda.cu — C ABI (include/polyred_cuda.h) over the sm_100a kernels in prc_kernels.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared -Xcompiler -fPIC
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <unistd.h>

#include "prc_kernels.cuh"
#include "prc_peer.cuh"

#define PRC_SHADE_BANDS 8  // row bands of the shading pass when the frame is read back (copy of band k overlaps shading of band k+1)
#define PRC_SHADE_BANDS_MAX 32  // PRC_SHADE_BANDS=n in the environment overrides the default (tuning: a smaller last band = a shorter exposed copy)

using namespace prc;

namespace {

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct prc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // PRC_FMA = exact | mixed (default) | fast, see DESIGN.md "Arithmetic contract":
  //   exact: every math.FMA is a float64 fma rounded to float32; mixed: exact for everything that decides the
  //   G-buffer and the shadow maps (geometry, raster, resolve), fmaf in the deferred shading; fast: fmaf everywhere.
  bool exact = true, exact_shade = false;

  // scene
  bool has_scene = false;
  DevScene S{};
  std::vector<DBuf*> scene_bufs;
  DBuf d_pos, d_nor, d_uv, d_col, d_mat, d_meta, d_mats, d_objstart, d_texfirst, d_lw, d_lh, d_loff, d_tex;
  uint32_t n_objects = 0;
  unsigned long long n_valid = 0;
  bool any_ao = false;

  // frame buffers
  int W = 0, H = 0;
  DBuf d_keys, d_ga, d_gb, d_gc, d_gd, d_ao, d_image, d_special, d_counters;
  DBuf d_large, d_clipq, d_tilecount, d_tilestart, d_cursor, d_bins, d_active, d_chunksum, d_targets;
  uint32_t n_targets = 1;
  uint32_t target_of_light[64] = {0};
  bool light_affine[64] = {false};
  DBuf d_xf, d_lights, d_ambient, d_gamma, d_frame, d_aoc, d_chunkbox, d_vis, d_cverts, d_cvoff, d_lidx;
  DBuf d_chunklist;  // compacted list of the chunks a raster pass can touch (k_chunk_cull_views)
  unsigned int* h_list_hint = nullptr;  // page-locked, device-visible: list length per raster pass of the last frame (grid sizing hint)
  int pass_slot = 0;                    // raster passes enqueued so far in the current frame
  bool no_chunk_cull = false, force_chunk_cull = false;  // PRC_NO_CHUNK_CULL / PRC_FORCE_CHUNK_CULL
  bool need_bins = false;               // a frame queued a record too large for k_medium_raster: the binned tile path runs from then on
  uint32_t n_chunks = 0;
  uint64_t n_cverts = 0;  // distinct chunk-local vertices (k_chunk_dedupe)
  std::vector<DevLight> h_lights;    // host staging of the per-frame light table
  TileTargets h_targets{};
  std::vector<DBuf> d_shadow_trans;  // per light
  // shadow maps of the casting lights, persistent, in ONE contiguous allocation (casting order) so that the
  // multi-GPU exchange is a single in-place all-gather; shadow_ptr[i] is light i's map or nullptr
  DBuf d_shadow_all;
  std::vector<float*> shadow_ptr;
  // Peer groups (prc_render_peer): the raster passes of a rank cover its share of the triangles and write PRIVATE buffers — d_keys
  // and d_shadow_mine (same layout as d_shadow_all) — which k_peer_push merges into every rank's d_shadow_all / d_mkeys
  // (merged visibility keys: 2 frame parities x [H][W], + 2 NaN-mode planes); shading reads the merged ones.
  DBuf d_shadow_mine, d_mkeys, d_dirty;  // d_dirty: row flags of the private buffers (Counters.dirty)
  bool skip_key_clear = false;            // peer frames: k_peer_push leaves the private key plane empty (clear-on-read)
  bool peer_private = false;               // connected: raster targets are the private buffers
  uint32_t part_rank = 0, part_world = 1;  // while a peer frame's raster passes are enqueued: this rank's share of the chunks
  bool part_active = false;
  unsigned long long* keys_shade = nullptr;   // while a peer frame is enqueued: the merged key plane of this frame (nullptr: d_keys)
  unsigned long long* first_shade = nullptr;  // ... and its NaN-mode plane
  uint32_t n_cast_alloc = 0;
  unsigned int large_cap = 0, clip_cap = 0, bins_cap = 0;
  AoConsts ao{};
  uint8_t* h_img[2] = {nullptr, nullptr};  // page-locked host images, used alternately
  size_t h_img_cap = 0;
  int h_img_cur = 0;
  Counters* h_counters = nullptr;  // pinned

  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // readback overlapped with shading: the image leaves in row bands on a second stream while the next band is shaded
  cudaStream_t copy_stream = nullptr;
  cudaStream_t copy_stream2 = nullptr;  // peer frames: the shadow sweep runs here, beside the camera pass on `stream`
  // PRC_FRAME_IMAGE_AT_SYNC: an image consumer waits for its peers' strips on img_stream, beside the frame's stream
  cudaStream_t img_stream = nullptr;
  cudaEvent_t ev_own_shaded = nullptr, ev_image_done = nullptr;
  bool img_pending = false;        // img_stream holds waits of frames not yet finished by prc_sync
  bool late_free_pending = false;  // the previous frame's "image free" announcement is img_stream's job (not the next frame's first operation)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_band[PRC_SHADE_BANDS_MAX] = {}, ev_copied = nullptr;
  int shade_bands = PRC_SHADE_BANDS;
  uint8_t* rb_dst = nullptr;  // page-locked destination of this frame's image (nullptr: no readback)
  // prc_render_batch: page-locked staging of the per-view uniforms, two halves used alternately (see upload())
  uint8_t* stage = nullptr;
  size_t stage_cap = 0, stage_off = 0, stage_end = 0;
  bool stage_on = false;
  cudaEvent_t ev_view[2] = {nullptr, nullptr};
  // MSAA: the shaded frame is W x H = msaa x the output; k_resize writes the (W/msaa) x (H/msaa) frame that is handed back
  int pending_async = 0;  // PRC_FRAME_ASYNC frames submitted since the last finish (their spans / overflow flag are still open)
  int msaa = 1;
  DBuf d_image_out, d_rz_cx, d_rz_sx, d_rz_cy, d_rz_sy;
  int rz_key[4] = {0, 0, 0, 0}, rz_flx = 0, rz_fly = 0;
  // per-kernel-class event pairs of the current frame
  std::vector<cudaEvent_t> evpool;
  struct Span { int cls; size_t a, b; };
  std::vector<Span> spans;
  size_t ev_used = 0;
  prc_timings tm{};
  unsigned long long launches = 0;
  bool gbuffer_valid = false, uniforms_valid = false, frame_uploaded = false;
  // NaN mode (bug-list 8, see nan_first in prc_kernels.cuh): entered, with a re-render of the frame, when a camera pass reported a
  // NaN-depth fragment; left at the next scene upload. d_keys then holds a second [H][W] u64 plane, the per-pixel first fragment.
  bool nan_mode = false;
  DevFrame h_frame{};
  bool capturing = false;
  bool no_fused_shade = false;  // PRC_NO_FUSED_SHADE
  bool two_streams = false;     // PRC_TWO_STREAMS=1: one-GPU frames run the shadow sweep beside the camera pass (second stream), as peer frames do
  bool zero_copy_out = false;   // PRC_ZERO_COPY_OUT=1: the shading kernels store the frame into the page-locked host image themselves (no device->host copies)
  bool stage_single = false;    // PRC_STAGE_UNIFORMS=1: a synchronous prc_render stages its uniforms in page-locked memory (one memcpy + async DMAs)
  bool ktimers = true;  // PRC_NO_KTIMERS=1: no per-kernel-class event brackets (prc_timings.kernel_ms stays 0)
  bool ktimers_frame = true;  // ... for the frame being enqueued (PRC_FRAME_NO_KERNEL_TIMERS)
  uint32_t n_lights_alloc = 0;

  // peer exchange (prc_peer_* / prc_render_peer): peers.world == 0 <=> not connected
  DBuf d_peer_signals, d_peer_err;
  PeerTable peers{};
  uint8_t* peer_image[PRC_PEER_MAX] = {};
  std::vector<void*> peer_opened;  // bases returned by cudaIpcOpenMemHandle
  void *peer_shadow_self = nullptr, *peer_image_self = nullptr;  // the exported buffers (a reallocation invalidates the group)
  uint32_t peer_epoch = 0;
  unsigned long long peer_timeout_ns = PRC_PEER_TIMEOUT_NS;  // PRC_PEER_TIMEOUT_MS
  // rows of this rank's two merged key planes that the last frame of that parity received keys for ([rr0, rr1) and [0, ax1)):
  // exactly these are cleared before the plane is used again, so a plane is all zero whatever strips the next frame brings
  int plane_rows[2][3] = {{0, 0, 0}, {0, 0, 0}};
  uint8_t* ext_img = nullptr;    // caller-owned page-locked host image (prc_set_host_image)
  bool defer_copy_join = false, copy_pending = false;  // see do_main: readback of peer frames overlaps the next frame's geometry
  size_t ext_img_bytes = 0, ext_img_off = 0;  // ext_img_off: prc_set_host_image_offset (double-buffered host images inside one registration)
  bool ext_img_registered = false;
  // MSAA frames cut into strips (prc_render_peer only): the rank shades `msaa` extra supersampled rows on either side of the
  // rows it owns — all the downsample filter reaches — and resizes only its own output rows; no exchange is needed
  // PRC_PEER_TRACE=1: CUDA-event brackets around the device-side waits, summed per signal kind by prc_sync (prc_peer_wait_ms):
  // where a rank idles for its peers
  bool peer_trace = false;
  std::vector<prc_ctx::Span> peer_spans;
  float peer_wait_ms[4] = {0, 0, 0, 0};
  bool allow_msaa_strips = false;
  int own_row0 = 0, own_row1 = 0;  // SCREEN rows (supersampled) whose output rows this context produces
};

namespace {

// peer exchange (defined at the end of this file)
void peer_release(prc_ctx* ctx);            // unmaps the peers' buffers
int32_t peer_check(prc_ctx* ctx);           // after a synchronisation: did a device-side wait for a peer give up?

#define PRC_RETRY 1
#define CK(call)                                                                             \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
      return PRC_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

int32_t ensure(prc_ctx* ctx, DBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return PRC_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  if (bytes == 0) bytes = 16;
  CK(cudaMalloc(&b.p, bytes));
  b.cap = bytes;
  return PRC_OK;
}
#define ENSURE(buf, bytes)                         \
  do {                                             \
    int32_t r_ = ensure(ctx, buf, bytes);          \
    if (r_ != PRC_OK) return r_;                   \
  } while (0)

int32_t upload(prc_ctx* ctx, DBuf& b, const void* src, size_t bytes) {
  ENSURE(b, bytes);
  if (!bytes) return PRC_OK;
  if (ctx->stage_on && ctx->stage_off + bytes <= ctx->stage_end) {
    // view batches: a copy from pageable memory makes the driver wait for the stream before it stages the data, which would
    // serialise every view's uniforms behind the previous view's kernels — stage them in page-locked memory here instead
    memcpy(ctx->stage + ctx->stage_off, src, bytes);
    CK(cudaMemcpyAsync(b.p, ctx->stage + ctx->stage_off, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stage_off += (bytes + 255) & ~(size_t)255;
    return PRC_OK;
  }
  CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return PRC_OK;
}
#define UPLOAD(buf, src, bytes)                    \
  do {                                             \
    int32_t r_ = upload(ctx, buf, src, bytes);     \
    if (r_ != PRC_OK) return r_;                   \
  } while (0)

void free_buf(DBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

// where the shadow raster of light i writes: the map itself, or this rank's private copy in a peer group
inline float* raster_map(prc_ctx* ctx, uint32_t i) {
  float* m = ctx->shadow_ptr[i];
  if (!m || !ctx->peer_private) return m;
  return (float*)ctx->d_shadow_mine.p + (m - (float*)ctx->d_shadow_all.p);
}

// Rows a context rasterises / resolves for the strip [row0, row1) it shades (screen rows of the frame buffer): the strip widened by
// the MSAA filter reach and the AO halo, and — for an upper strip of a frame with an AO material — rows [0, ax1) as well (see do_main)
inline void row_needs(bool any_ao, int msaa, bool msaa_strip, int H, int row0, int row1, int& s0, int& s1, int& rr0, int& rr1, int& ax1) {
  s0 = row0; s1 = row1;
  if (msaa_strip) {
    // output row y of imageutil.Resize reads supersampled rows [msaa*y - msaa, msaa*y + 2*msaa): shade that much beyond the strip
    s0 = std::max(0, row0 - msaa);
    s1 = std::min(H, row1 + msaa);
  }
  // AO marches up to 99 pixels from the shaded pixel (material/ao.go:44-46): widen the rasterised rows
  const int halo = any_ao ? 100 : 0;
  rr0 = std::max(0, s0 - halo);
  rr1 = std::min(H, s1 + halo);
  if (any_ao && rr0 <= 100) rr0 = 0;  // rows [0, 100) are needed for pixel (0,0)'s AO anyway: keep one contiguous range
  ax1 = (any_ao && rr0 > 0) ? std::min(100, rr0) : 0;
}

// CUDA-event bracket around one kernel class (timed on the launching stream)
struct KTimer {
  prc_ctx* ctx;
  size_t a;
  int cls;
  KTimer(prc_ctx* c, int k) : ctx(c), cls(k) {
    if (!ctx->ktimers_frame || k < 0) { cls = -1; return; }
    while (ctx->evpool.size() < ctx->ev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); ctx->evpool.push_back(e); }
    a = ctx->ev_used;
    ctx->ev_used += 2;
    cudaEventRecordWithFlags(ctx->evpool[a], ctx->stream, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
  }
  ~KTimer() {
    if (cls < 0) return;
    cudaEventRecordWithFlags(ctx->evpool[a + 1], ctx->stream, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
    ctx->spans.push_back({cls, a, a + 1});
  }
};

// bit i set <=> entry i is 0 or +-2^k with |k| <= 32 (see apply4m in prc_math.cuh)
uint32_t plain_mask(const float* m) {
  uint32_t pm = 0;
  for (int i = 0; i < 16; i++) {
    float a = std::fabs(m[i]);
    int e = 0;
    if (a == 0.0f || (std::isfinite(a) && std::frexp(a, &e) == 0.5f && e >= -31 && e <= 33)) pm |= 1u << i;
  }
  return pm;
}

inline unsigned int cdiv(unsigned long long a, unsigned int b) { return (unsigned int)((a + b - 1) / b); }
inline size_t ipc_round(size_t bytes) { return (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1); }

// ---- one raster pass (camera, or up to 8 shadow views in one sweep): geometry + small raster through the CTA queue; large triangles and
// (camera only) triangles needing clipping are queued. No host round trip. -----------------------------------
template <bool E, bool SHADOW>
int32_t raster_pass(prc_ctx* ctx, const DevFrame& F, const GeomViews& Vin) {
  cudaStream_t st = ctx->stream;
  Counters* cnt = (Counters*)ctx->d_counters.p;
  unsigned long long* keys = (unsigned long long*)ctx->d_keys.p;
  LargeRec* large = (LargeRec*)ctx->d_large.p;
  unsigned int* clipq = (unsigned int*)ctx->d_clipq.p;
  DBuf& fb = ctx->d_frame;  // device-resident copy of the frame for the rare generic path (uploaded by build_frame)
  GeomViews V = Vin;
  bool list_mode = false;
  unsigned int grid = cdiv(ctx->S.n_tris, PRC_GEOM_THREADS);
  const bool part = ctx->part_active;  // a rank of a peer group: its share of the chunks, every row
  if (part) {
    const unsigned int nb = cdiv(ctx->n_chunks, PRC_PART_BLOCK);  // blocks of chunks, dealt round-robin
    const unsigned int mine = ctx->part_rank < nb ? cdiv(nb - ctx->part_rank, ctx->part_world) : 0u;
    grid = std::max(1u, mine * PRC_PART_BLOCK);
    V.part_first = ctx->part_rank * PRC_PART_BLOCK;
    V.part_stride = ctx->part_world * PRC_PART_BLOCK;
    V.n_chunks = ctx->n_chunks;
    V.dirty = (unsigned char*)ctx->d_dirty.p;
  }
  if (ctx->S.n_tris && !ctx->no_chunk_cull && !part) {
    // chunk culling pays when a view covers only part of the rows (multi-GPU strips / shadow shards): one launch tests every
    // view of the pass and compacts the chunks some view can touch; the geometry grid then walks that list
    CullViews C{};
    C.n = SHADOW ? V.n : 1;
    C.stride = SHADOW ? 16 : 32;
    C.need_pixel00 = (!SHADOW && F.rr0 > 0) ? 1 : 0;
    bool any_test = false;
    for (int v = 0; v < C.n; v++) {
      C.r0[v] = SHADOW ? V.r0[v] : F.rr0;
      C.r1[v] = SHADOW ? V.r1[v] : F.rr1;
      C.test[v] = (!(C.r0[v] <= 0 && C.r1[v] >= F.H) || ctx->force_chunk_cull) ? 1 : 0;
      any_test = any_test || C.test[v];
      C.trans[v] = SHADOW ? V.trans[v] : (const float*)ctx->d_xf.p;
      C.vis[v] = (unsigned char*)ctx->d_vis.p + (size_t)(SHADOW ? 1 + v : 0) * ctx->n_chunks;
    }
    const int slot = ctx->pass_slot++;
    if (any_test && slot < 16) {
      ENSURE(ctx->d_chunklist, ((size_t)ctx->n_chunks + 4) * 4);
      unsigned int* n_list = &cnt->n_list[slot];  // zeroed with the frame's counters
      k_chunk_cull_views<<<cdiv(ctx->n_chunks, 256), 256, 0, st>>>((const ChunkBox*)ctx->d_chunkbox.p, ctx->n_chunks, C,
                                                                   (const float*)((const char*)fb.p + offsetof(DevFrame, viewport)), F.W,
                                                                   (unsigned int*)ctx->d_chunklist.p, n_list);
      ctx->launches++;
      for (int v = 0; v < C.n; v++) V.vis[v] = C.vis[v];
      V.any_vis = 1;
      V.list = (const unsigned int*)ctx->d_chunklist.p;
      V.n_list = n_list;
      V.n_list_hint = ctx->h_list_hint ? ctx->h_list_hint + slot : nullptr;
      list_mode = true;
      // grid from the list length of the previous frame's pass in this slot (0 = not known yet: one CTA per chunk)
      const unsigned int hint = ctx->h_list_hint ? ctx->h_list_hint[slot] : 0u;
      if (hint) grid = std::min(grid, std::max(148u * 2u, hint + hint / 8u + 32u));
      V.list_from = grid;
    }
  }
  {
    // (one event bracket per pass: an event pair costs ~2.6 us of stream time, so the clip kernel — a fixed grid that finds an
    // empty queue on most frames — is timed inside the camera pass's bracket; the "clip" class stays 0)
    KTimer kt(ctx, SHADOW ? PRC_K_GEOM_SHADOW : PRC_K_GEOM_CAMERA);
  if (ctx->S.n_tris) {
    const DevFrame* Fg = (const DevFrame*)fb.p;
    const bool tail = list_mode && grid < cdiv(ctx->S.n_tris, PRC_GEOM_THREADS);  // the list may have outgrown the hinted grid
#define PRC_LAUNCH_GEOM(NM_, LIST_, GRID_) \
  k_geom_raster<E, SHADOW, NM_, LIST_><<<GRID_, PRC_GEOM_THREADS, 0, st>>>(ctx->S, F, V, keys, large, ctx->large_cap, clipq, ctx->clip_cap, cnt, Fg)
    bool done = false;
    if constexpr (!SHADOW) {
      if (ctx->nan_mode) {
        if (list_mode) { PRC_LAUNCH_GEOM(true, 1, grid); if (tail) PRC_LAUNCH_GEOM(true, 2, 148u); }
        else if (part) PRC_LAUNCH_GEOM(true, 3, grid);
        else PRC_LAUNCH_GEOM(true, 0, grid);
        done = true;
      }
    }
    if (!done) {
      if (list_mode) { PRC_LAUNCH_GEOM(false, 1, grid); if (tail) PRC_LAUNCH_GEOM(false, 2, 148u); }
      else if (part) PRC_LAUNCH_GEOM(false, 3, grid);
      else PRC_LAUNCH_GEOM(false, 0, grid);
    }
#undef PRC_LAUNCH_GEOM
    ctx->launches += tail ? 2 : 1;
  }
  if (!SHADOW) {
    // fixed grid, reads its work count on the device (grid-stride loop)
    if (ctx->nan_mode) k_clip_raster<E, true><<<148, 128, 0, st>>>(ctx->S, F, clipq, keys, large, ctx->large_cap, cnt);
    else k_clip_raster<E><<<148, 128, 0, st>>>(ctx->S, F, clipq, keys, large, ctx->large_cap, cnt);
    ctx->launches++;
  }
  }
  CK(cudaGetLastError());
  return PRC_OK;
}

// ---- everything queued so far (camera keys and shadow maps), once per frame (or per phase in the split multi-GPU API): records up
// to PRC_MEDIUM_MAX_PIXELS by k_medium_raster, larger ones through the binned tile path — which is only launched once a frame
// has reported such a record (ctx->need_bins, sticky; the reporting frame is re-rendered like a queue overflow) --------
template <bool E>
int32_t flush_large(prc_ctx* ctx, const DevFrame& F) {
  cudaStream_t st = ctx->stream;
  Counters* cnt = (Counters*)ctx->d_counters.p;
  LargeRec* large = (LargeRec*)ctx->d_large.p;
  {
    KTimer kt(ctx, PRC_K_MEDIUM);
    if (ctx->nan_mode)
      k_medium_raster<E, true><<<148 * 2, 256, 0, st>>>(large, cnt, ctx->large_cap, F.W, F.H, (unsigned long long*)ctx->d_keys.p, (const TileTargets*)ctx->d_targets.p, ctx->need_bins ? 1 : 0);
    else
      k_medium_raster<E><<<148 * 2, 256, 0, st>>>(large, cnt, ctx->large_cap, F.W, F.H, (unsigned long long*)ctx->d_keys.p, (const TileTargets*)ctx->d_targets.p, ctx->need_bins ? 1 : 0);
    ctx->launches++;
  }
  if (ctx->need_bins) {
    const int tiles_x = (F.W + PRC_TILE - 1) / PRC_TILE, tiles_y = (F.H + PRC_TILE - 1) / PRC_TILE;
    const int n_tiles = tiles_x * tiles_y, n_vt = n_tiles * (int)ctx->n_targets;
    const unsigned int n_chunks = (unsigned int)((n_vt + 1 + 4095) / 4096);
    unsigned int* tile_count = (unsigned int*)ctx->d_tilecount.p;
    unsigned int* tile_start = (unsigned int*)ctx->d_tilestart.p;
    unsigned int* cursor = (unsigned int*)ctx->d_cursor.p;
    unsigned int* active_hdr = (unsigned int*)ctx->d_active.p;
    CK(cudaMemsetAsync(tile_count, 0, (size_t)n_chunks * 4096 * 4, st));
    CK(cudaMemsetAsync(active_hdr, 0, 16, st));
    {
      KTimer kt(ctx, PRC_K_BIN);
      k_bin_count<<<148 * 4, 256, 0, st>>>(large, cnt, ctx->large_cap, tiles_x, n_tiles, tile_count);
      k_scan_sums<<<n_chunks, 1024, 0, st>>>(tile_count, (unsigned int*)ctx->d_chunksum.p, cnt);
      k_scan_apply<<<n_chunks, 1024, 0, st>>>(tile_count, tile_start, cursor, (const unsigned int*)ctx->d_chunksum.p, cnt, ctx->bins_cap, ctx->large_cap,
                                              active_hdr + 4, active_hdr);
      k_bin_fill<<<148 * 4, 256, 0, st>>>(large, cnt, ctx->large_cap, tiles_x, n_tiles, cursor, (unsigned int*)ctx->d_bins.p, ctx->bins_cap);
      ctx->launches += 4;
    }
    {
      KTimer kt(ctx, PRC_K_TILE);
      if (ctx->nan_mode)
        k_tile_raster<E, true><<<148 * 8, PRC_TILE * PRC_TILE, 0, st>>>(large, tile_start, (unsigned int*)ctx->d_bins.p, tiles_x, n_tiles, F.W, F.H,
                                                                         (unsigned long long*)ctx->d_keys.p, (const TileTargets*)ctx->d_targets.p, cnt, active_hdr + 4, active_hdr);
      else
        k_tile_raster<E><<<148 * 8, PRC_TILE * PRC_TILE, 0, st>>>(large, tile_start, (unsigned int*)ctx->d_bins.p, tiles_x, n_tiles, F.W, F.H,
                                                                   (unsigned long long*)ctx->d_keys.p, (const TileTargets*)ctx->d_targets.p, cnt, active_hdr + 4, active_hdr);
      ctx->launches++;
    }
  }
  // the queue is consumed: the next phase starts an empty one (stats were accumulated by the kernels above)
  CK(cudaMemsetAsync(cnt, 0, 16, st));
  CK(cudaGetLastError());
  return PRC_OK;
}

int32_t build_frame(prc_ctx* ctx, const prc_frame* fr, DevFrame& F) {
  if (!fr || fr->abi_version != PRC_ABI_VERSION) { ctx->err = "prc_frame: bad abi_version"; return PRC_ERR_INVALID; }
  if (!ctx->has_scene) { ctx->err = "no scene uploaded"; return PRC_ERR_NO_SCENE; }
  if (fr->n_objects != ctx->n_objects) { ctx->err = "prc_frame.n_objects differs from the uploaded scene"; return PRC_ERR_INVALID; }
  if (fr->width == 0 || fr->height == 0 || fr->width > 16384 || fr->height > 16384) { ctx->err = "bad frame size"; return PRC_ERR_INVALID; }
  if (fr->row1 > fr->height || fr->row0 >= fr->row1) { ctx->err = "bad row range"; return PRC_ERR_INVALID; }
  for (uint32_t i = 0; i < fr->n_lights; i++) {
    if (fr->lights[i].kind > PRC_LIGHT_DIRECTIONAL) { ctx->err = "unsupported light kind"; return PRC_ERR_UNSUPPORTED; }
    if (fr->lights[i].cast_shadow && !fr->lights[i].shadow_trans) { ctx->err = "casting light without shadow_trans"; return PRC_ERR_INVALID; }
  }
  const int W = fr->width, H = fr->height;
  const size_t npx = (size_t)W * H;
  const int msaa = fr->msaa > 1 ? (int)fr->msaa : 1;
  if (msaa > 8 || W % msaa || H % msaa) { ctx->err = "msaa must be 1..8 and divide the frame size"; return PRC_ERR_INVALID; }
  const bool msaa_strip = msaa > 1 && (fr->row0 != 0 || fr->row1 != fr->height);
  if (msaa_strip && !ctx->allow_msaa_strips) { ctx->err = "MSAA with a partial row range is only supported by prc_render_peer"; return PRC_ERR_UNSUPPORTED; }
  if (msaa_strip && (fr->row0 % msaa || fr->row1 % msaa)) { ctx->err = "MSAA strips must start and end on multiples of msaa"; return PRC_ERR_INVALID; }
  ctx->msaa = msaa;
  ctx->ktimers_frame = ctx->ktimers && !(fr->flags & PRC_FRAME_NO_KERNEL_TIMERS);
  cudaStream_t st = ctx->stream;
  if (W != ctx->W || H != ctx->H || fr->n_lights != ctx->n_lights_alloc) {
    // resetBufs + initShadowMaps: new size => fresh (zero) shadow maps
    ctx->shadow_ptr.assign(fr->n_lights, nullptr);
    ctx->n_cast_alloc = 0xFFFFFFFFu;  // force (re)allocation + zeroing below
    for (auto& b : ctx->d_shadow_trans) free_buf(b);
    ctx->d_shadow_trans.assign(fr->n_lights, DBuf());
    ctx->W = W; ctx->H = H; ctx->n_lights_alloc = fr->n_lights;
    ctx->gbuffer_valid = false;
  }
  ENSURE(ctx->d_keys, npx * 8 * (ctx->nan_mode ? 2 : 1));
  for (DBuf* g : {&ctx->d_ga, &ctx->d_gb, &ctx->d_gc, &ctx->d_gd}) {
    // zeroed once when (re)allocated: the resolve only writes covered pixels and prc_read_gbuffer copies whole planes
    // (compute-sanitizer initcheck flagged the never-written texels of that debug read)
    const size_t before = g->p ? g->cap : 0;
    ENSURE(*g, npx * 16);
    if (g->cap != before) CK(cudaMemsetAsync(g->p, 0, g->cap, st));
  }
  if (ctx->any_ao) ENSURE(ctx->d_ao, npx * 4);
  // slack: the multi-GPU image all-gather uses equal, padded strips. Whole 2 MiB pages: the buffer can be exported through CUDA
  // IPC (prc_peer_export), which shares entire allocation blocks — an exported buffer must not share its block with others
  ENSURE(ctx->d_image, ipc_round(npx * 4 + (size_t)64 * W * 4));
  ENSURE(ctx->d_special, 16);
  const int n_tiles = ((W + PRC_TILE - 1) / PRC_TILE) * ((H + PRC_TILE - 1) / PRC_TILE);
  // targets of the tile path: 0 = camera, 1 + k = k-th casting light
  {

    uint32_t nt = 1;
    for (uint32_t i = 0; i < fr->n_lights && i < 64; i++) {
      ctx->target_of_light[i] = 0;
      if (fr->lights[i].cast_shadow && (fr->flags & PRC_FRAME_SHADOWMAP)) {
        if (nt > 32) { ctx->err = "more than 32 shadow-casting lights"; return PRC_ERR_UNSUPPORTED; }
        ctx->target_of_light[i] = nt;
        nt++;
      }
    }
    ctx->n_targets = nt;
    const size_t n_vt_pad = (((size_t)n_tiles * nt + 1 + 4095) / 4096) * 4096;
    ENSURE(ctx->d_tilecount, n_vt_pad * 4); ENSURE(ctx->d_tilestart, n_vt_pad * 4); ENSURE(ctx->d_cursor, n_vt_pad * 4);
    ENSURE(ctx->d_active, (n_vt_pad + 8) * 4);
    ENSURE(ctx->d_chunksum, 1024 * 4);
    ENSURE(ctx->d_targets, sizeof(TileTargets));
  }
  if (!ctx->d_bins.p) {
    const char* e = getenv("PRC_BINS_INIT");  // initial (tile, triangle) capacity; grown on demand by re-rendering the frame
    ENSURE(ctx->d_bins, e ? (size_t)atoll(e) * 4 : ((size_t)16 << 20));
    ctx->bins_cap = (unsigned int)(ctx->d_bins.cap / 4);
  }
  bool realloc_shadow = false;
  {
    uint32_t ncast = 0;
    for (uint32_t i = 0; i < fr->n_lights; i++)
      if (fr->lights[i].cast_shadow && (fr->flags & PRC_FRAME_SHADOWMAP)) ncast++;
    bool same = ncast == ctx->n_cast_alloc;
    for (uint32_t i = 0, k = 0; same && i < fr->n_lights; i++) {
      const bool c = fr->lights[i].cast_shadow && (fr->flags & PRC_FRAME_SHADOWMAP);
      same = c ? (ctx->shadow_ptr[i] == (float*)ctx->d_shadow_all.p + (size_t)(k++) * npx) : (ctx->shadow_ptr[i] == nullptr);
    }
    if (!same) {  // new set of casting lights: fresh zero maps (what initShadowMaps does, shadow.go:87)
      realloc_shadow = true;
      const size_t bytes = ipc_round(((size_t)ncast * npx + (size_t)64 * W) * 4);  // slack: the all-gather chunks are padded; whole 2 MiB pages (IPC export)
      ENSURE(ctx->d_shadow_all, bytes);
      CK(cudaMemsetAsync(ctx->d_shadow_all.p, 0, bytes, st));
      for (uint32_t i = 0, k = 0; i < fr->n_lights; i++) {
        const bool c = fr->lights[i].cast_shadow && (fr->flags & PRC_FRAME_SHADOWMAP);
        ctx->shadow_ptr[i] = c ? (float*)ctx->d_shadow_all.p + (size_t)(k++) * npx : nullptr;
      }
      ctx->n_cast_alloc = ncast;
    }
  }
  // uniforms
  const bool resident = (fr->flags & PRC_FRAME_UNIFORMS_RESIDENT) && ctx->uniforms_valid && !realloc_shadow;
  if (!resident) UPLOAD(ctx->d_xf, fr->objects, (size_t)fr->n_objects * sizeof(prc_object_xf));
  std::vector<DevLight>& hl = ctx->h_lights;
  hl.assign(fr->n_lights, DevLight());
  for (uint32_t i = 0; i < fr->n_lights; i++) {
    const prc_light& l = fr->lights[i];
    DevLight& d = hl[i];
    d.kind = l.kind;
    d.cast_shadow = (l.cast_shadow && (fr->flags & PRC_FRAME_SHADOWMAP)) ? 1u : 0u;
    memcpy(d.pos, l.pos, 12);
    d.intensity = l.intensity;
    d.color = l.color_rgba;
    for (int k = 0; k < 3; k++) d.colf[k] = (float)((l.color_rgba >> (8 * k)) & 0xffu);
    memcpy(d.view, l.view, 64);
    memcpy(d.proj, l.proj, 64);
    d.pm_view = plain_mask(l.view);
    d.pm_proj = plain_mask(l.proj);
    {
      // rigid view [R t; 0 0 0 1] with a perspective [a 0 0 0; 0 b 0 0; 0 0 c d; 0 0 -1 0] (1) or an orthographic
      // [a 0 0 b; 0 c 0 d; 0 0 e f; 0 0 0 1] (2) projection (math.Mat4 LookAt / Perspective / Orthographic; initShadowMaps
      // fits an orthographic light camera, render/shadow.go:67-86)
      const float *v = l.view, *q = l.proj;
      bool fin = true;
      for (int k = 0; k < 16; k++) fin = fin && std::isfinite(v[k]) && std::isfinite(q[k]);
      const bool rigid = v[12] == 0 && v[13] == 0 && v[14] == 0 && v[15] == 1;
      const bool persp = q[1] == 0 && q[2] == 0 && q[3] == 0 && q[4] == 0 && q[6] == 0 && q[7] == 0 && q[8] == 0 && q[9] == 0 && q[12] == 0 && q[13] == 0 &&
                         q[14] == -1 && q[15] == 0;
      const bool ortho = q[1] == 0 && q[2] == 0 && q[4] == 0 && q[6] == 0 && q[8] == 0 && q[9] == 0 && q[12] == 0 && q[13] == 0 && q[14] == 0 && q[15] == 1;
      d.persp_cam = (!fin || !rigid || getenv("PRC_NO_PERSP_CAM")) ? 0u : persp ? 1u : ortho ? 2u : 0u;
    }
    d.shadow_map = ctx->shadow_ptr[i];
    if (d.cast_shadow && !resident) {
      UPLOAD(ctx->d_shadow_trans[i], l.shadow_trans, (size_t)fr->n_objects * 64);
      bool aff = true;
      for (uint32_t o = 0; o < fr->n_objects && aff; o++) {
        const float* m = l.shadow_trans + (size_t)o * 16;
        aff = m[12] == 0.0f && m[13] == 0.0f && m[14] == 0.0f && m[15] == 1.0f;
      }
      if (i < 64) ctx->light_affine[i] = aff && !getenv("PRC_NO_AFFINE");
    }
  }
  if (!resident) UPLOAD(ctx->d_lights, hl.data(), hl.size() * sizeof(DevLight));
  if (!resident) {
    ctx->h_targets = TileTargets{};
    for (uint32_t i = 0; i < fr->n_lights && i < 64; i++)
      if (ctx->target_of_light[i]) ctx->h_targets.smap[ctx->target_of_light[i]] = raster_map(ctx, i);
    UPLOAD(ctx->d_targets, &ctx->h_targets, sizeof(TileTargets));
  }
  if (!resident) {
    UPLOAD(ctx->d_ambient, fr->ambient_intensity, (size_t)fr->n_ambient * 4);
    UPLOAD(ctx->d_gamma, fr->gamma_lut, 256);
  }
  ctx->uniforms_valid = true;
  F.W = W; F.H = H;
  F.cullW = (float)(msaa * W); F.cullH = (float)(msaa * H);
  ctx->own_row0 = (int)fr->row0; ctx->own_row1 = (int)fr->row1;
  {
    int s0, s1, rr0, rr1, ax1;
    row_needs(ctx->any_ao, msaa, msaa_strip, H, (int)fr->row0, (int)fr->row1, s0, s1, rr0, rr1, ax1);
    F.row0 = s0; F.row1 = s1; F.rr0 = rr0; F.rr1 = rr1;
  }
  F.flags = fr->flags;
  F.n_lights = fr->n_lights; F.n_ambient = fr->n_ambient;
  F.background = fr->background_rgba;
  memcpy(F.viewport, fr->viewport, 64); memcpy(F.viewport_inv, fr->viewport_inv, 64);
  memcpy(F.proj_inv, fr->proj_inv, 64); memcpy(F.view_inv, fr->view_inv, 64); memcpy(F.vtw, fr->viewport_to_world, 64);
  F.pm_viewport = plain_mask(F.viewport); F.pm_viewport_inv = plain_mask(F.viewport_inv); F.pm_proj_inv = plain_mask(F.proj_inv);
  F.pm_view_inv = plain_mask(F.view_inv); F.pm_vtw = plain_mask(F.vtw);
  {
    const float* v = F.viewport;
    F.vp_std = (v[1] == 0 && v[2] == 0 && v[4] == 0 && v[6] == 0 && v[8] == 0 && v[9] == 0 && v[10] == 1 && v[11] == 0 && v[12] == 0 && v[13] == 0 &&
                v[14] == 0 && v[15] == 1 && std::isfinite(v[0]) && std::isfinite(v[3]) && std::isfinite(v[5]) && std::isfinite(v[7]) && v[3] != 0 && v[7] != 0)
                   ? 1u : 0u;
    if (getenv("PRC_NO_VPSTD")) F.vp_std = 0;
    const float *a = F.viewport_inv, *q = F.proj_inv, *r = F.view_inv;
    bool ok = a[1] == 0 && a[2] == 0 && a[4] == 0 && a[6] == 0 && a[8] == 0 && a[9] == 0 && a[10] == 1 && a[11] == 0 && a[12] == 0 && a[13] == 0 && a[14] == 0 &&
              a[15] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0 && q[4] == 0 && q[6] == 0 && q[7] == 0 && q[8] == 0 && q[9] == 0 && q[10] == 0 && q[12] == 0 &&
              q[13] == 0 && r[12] == 0 && r[13] == 0 && r[14] == 0 && r[15] == 1;
    for (int k = 0; k < 16; k++) ok = ok && std::isfinite(a[k]) && std::isfinite(q[k]) && std::isfinite(r[k]);
    F.unproj_std = (ok && !getenv("PRC_NO_UNPROJ_STD")) ? 1u : 0u;
  }
  if (getenv("PRC_NO_PLAIN")) F.pm_viewport = F.pm_viewport_inv = F.pm_proj_inv = F.pm_view_inv = F.pm_vtw = 0;
  memcpy(F.cam, fr->cam_pos, 12);
  F.xf = (const prc_object_xf*)ctx->d_xf.p;
  F.lights = (const DevLight*)ctx->d_lights.p;
  F.ambient = (const float*)ctx->d_ambient.p;
  F.gamma = (const uint8_t*)ctx->d_gamma.p;
  // device-resident copy of the frame for the rare non-inlined generic path (see geom_generic)
  if (!ctx->frame_uploaded || memcmp(&ctx->h_frame, &F, sizeof(DevFrame)) != 0) {
    ctx->h_frame = F;
    UPLOAD(ctx->d_frame, &ctx->h_frame, sizeof(DevFrame));
    ctx->frame_uploaded = true;
  }
  return PRC_OK;
}

struct ShadowUnit { uint32_t light; int r0, r1; };

// shadow passes: up to 8 (light, row range) units share one sweep over the triangles
template <bool E>
int32_t do_shadows(prc_ctx* ctx, const prc_frame* fr, const DevFrame& F, const std::vector<ShadowUnit>& units, bool flush) {
  if (!(fr->flags & PRC_FRAME_SHADOWMAP)) return PRC_OK;
  if ((fr->flags & PRC_FRAME_SHADOW_RESET) && ctx->d_shadow_all.p) CK(cudaMemsetAsync(ctx->d_shadow_all.p, 0, ctx->d_shadow_all.cap, ctx->stream));
  GeomViews V{};
  const int per_sweep = getenv("PRC_SHADOW_FUSE") ? std::max(1, std::min(8, atoi(getenv("PRC_SHADOW_FUSE")))) : 8;
  for (const ShadowUnit& u : units) {
    const uint32_t i = u.light;
    if (i >= fr->n_lights || !fr->lights[i].cast_shadow || u.r0 >= u.r1) continue;
    if (ctx->light_affine[i & 63]) V.affine |= 1u << V.n;
    V.trans[V.n] = (const float*)ctx->d_shadow_trans[i].p;
    V.smap[V.n] = raster_map(ctx, i);
    V.target[V.n] = ctx->target_of_light[i];
    V.r0[V.n] = u.r0;
    V.r1[V.n] = u.r1;
    if (++V.n == per_sweep) {
      int32_t r = raster_pass<E, true>(ctx, F, V);
      if (r != PRC_OK) return r;
      V = GeomViews{};
    }
  }
  if (V.n) {
    int32_t r = raster_pass<E, true>(ctx, F, V);
    if (r != PRC_OK) return r;
  }
  if (flush) return flush_large<E>(ctx, F);
  return PRC_OK;
}

std::vector<ShadowUnit> units_from_mask(const prc_frame* fr, uint32_t light_mask, int r0, int r1) {
  std::vector<ShadowUnit> u;
  for (uint32_t i = 0; i < fr->n_lights; i++)
    if (fr->lights[i].cast_shadow && ((light_mask >> (i & 31)) & 1u)) u.push_back({i, r0, r1});
  return u;
}

// createWeights8 (internal/imageutil/resize.go:146-164) with the `linear` kernel (:64-70): float32 arithmetic in the
// reference's order (host code is not FMA-contracted), int16 coefficients in [-256, 256].
static void make_weights8(int dy, int filter_length, float scale, std::vector<short>& coeffs, std::vector<int>& start, int& flen) {
  const float cs = (float)std::ceil((double)scale);
  flen = filter_length * (int)(cs > 1.0f ? cs : 1.0f);
  const float inv = 1.0f / scale;
  const float filter_factor = inv < 1.0f ? inv : 1.0f;
  coeffs.assign((size_t)dy * flen, 0);
  start.assign(dy, 0);
  for (int y = 0; y < dy; y++) {
    volatile float t = (float)y + 0.5f;
    volatile float m = scale * t;
    float interp = m - 0.5f;
    start[y] = (int)interp - flen / 2 + 1;
    interp -= (float)start[y];
    for (int i = 0; i < flen; i++) {
      volatile float d = interp - (float)i;
      float in = std::fabs(d * filter_factor);
      volatile float k = in <= 1.0f ? 1.0f - in : 0.0f;
      coeffs[(size_t)y * flen + i] = (short)(int)(k * 256.0f);
    }
  }
}

// device tables of the MSAA downsample (iw x ih -> ow x oh), rebuilt only when the sizes change
static int32_t ensure_resize_tables(prc_ctx* ctx, int iw, int ih, int ow, int oh) {
  if (ctx->rz_key[0] == iw && ctx->rz_key[1] == ih && ctx->rz_key[2] == ow && ctx->rz_key[3] == oh) return PRC_OK;
  std::vector<short> cx, cy;
  std::vector<int> sx, sy;
  make_weights8(ow, 2, (float)iw / (float)ow, cx, sx, ctx->rz_flx);  // calcFactors (:114-131) with both sizes given
  make_weights8(oh, 2, (float)ih / (float)oh, cy, sy, ctx->rz_fly);
  CK(cudaStreamSynchronize(ctx->stream));  // a previous frame may still read the old tables
  UPLOAD(ctx->d_rz_cx, cx.data(), cx.size() * sizeof(short));
  UPLOAD(ctx->d_rz_sx, sx.data(), sx.size() * sizeof(int));
  UPLOAD(ctx->d_rz_cy, cy.data(), cy.size() * sizeof(short));
  UPLOAD(ctx->d_rz_sy, sy.data(), sy.size() * sizeof(int));
  CK(cudaStreamSynchronize(ctx->stream));  // the vectors above are pageable host memory
  ctx->rz_key[0] = iw; ctx->rz_key[1] = ih; ctx->rz_key[2] = ow; ctx->rz_key[3] = oh;
  return PRC_OK;
}

template <bool E>
int32_t do_main(prc_ctx* ctx, const prc_frame* fr, const DevFrame& F,
                int phases = 15 /* bit0: camera raster passes, bit2: queued records, bit3: NaN fix + G-buffer resolve, bit1: deferred shading */,
                int fuse = -1) {
  cudaStream_t st = ctx->stream;
  // the raster passes write d_keys; resolve and shading read the same plane, or — peer frames — the merged plane of the group
  unsigned long long* keys_w = ctx->keys_shade ? ctx->keys_shade : (unsigned long long*)ctx->d_keys.p;
  const unsigned long long* keys = keys_w;
  const uint32_t* special = (const uint32_t*)ctx->d_special.p;
  uint32_t* image = (uint32_t*)ctx->d_image.p;
  const AoConsts* aoc = (const AoConsts*)ctx->d_aoc.p;
  GBuf G{(float4*)ctx->d_ga.p, (float4*)ctx->d_gb.p, (float4*)ctx->d_gc.p, (float4*)ctx->d_gd.p, ctx->any_ao ? (float*)ctx->d_ao.p : nullptr};
  // one kernel for resolve + shading when nothing reads the G-buffer afterwards (k_resolve_shade)
  // (a caller that enqueues the phases one by one — peer frames — says so with fuse = 1: nothing reads the G-buffer in between)
  const bool fused = (phases == 15 || fuse == 1) && !(fr->flags & PRC_FRAME_KEEP_GBUFFER) && !ctx->any_ao && !ctx->no_fused_shade;
  const bool es = ctx->exact_shade;  // exact FMA in the shading-only arithmetic too (PRC_FMA=exact)

  // Row ranges to rasterise and resolve: the strip (widened by the AO halo) and, for an upper strip of a frame with an AO
  // material, rows [0, 100) as well — the colour of every uncovered pixel comes from shading pixel (0,0) (bug-list 3), whose
  // AO rays read the depths of up to 99 rows / columns around it.
  std::vector<std::pair<int, int>> ranges;
  ranges.push_back({F.rr0, F.rr1});
  if (ctx->any_ao && F.rr0 > 0) ranges.push_back({0, std::min(100, F.rr0)});
  unsigned long long* first = (unsigned long long*)ctx->d_keys.p + (size_t)F.W * F.H;  // NaN mode: first fragment per pixel

  if (phases & 1) {
    // clear the visibility keys of the rasterised rows (+ pixel (0,0))
    if (!ctx->skip_key_clear)
    for (const auto& rg : ranges) {
      CK(cudaMemsetAsync((unsigned long long*)ctx->d_keys.p + (size_t)rg.first * F.W, 0, (size_t)(rg.second - rg.first) * F.W * 8, st));
      if (ctx->nan_mode) CK(cudaMemsetAsync(first + (size_t)rg.first * F.W, 0xFF, (size_t)(rg.second - rg.first) * F.W * 8, st));
    }
    if (F.rr0 > 0 && ranges.size() == 1 && !ctx->skip_key_clear) {
      CK(cudaMemsetAsync(ctx->d_keys.p, 0, 8, st));
      if (ctx->nan_mode) CK(cudaMemsetAsync(first, 0xFF, 8, st));
    }
    for (const auto& rg : ranges) {
      DevFrame Fr = F;
      Fr.rr0 = rg.first; Fr.rr1 = rg.second;
      GeomViews V0{};
      V0.n = 1;
      int32_t r = raster_pass<E, false>(ctx, Fr, V0);
      if (r != PRC_OK) return r;
    }
  }

  if (phases & 4) {
    int32_t r = flush_large<E>(ctx, F);  // also rasterises what the shadow passes of this frame queued
    if (r != PRC_OK) return r;
  }

  if (phases & 8) {
    if (ctx->nan_mode) {
      const unsigned long long* first_s = ctx->first_shade ? ctx->first_shade : first;
      for (const auto& rg : ranges) {
        const size_t i0 = (size_t)rg.first * F.W, i1 = (size_t)rg.second * F.W;
        k_nan_fix<<<cdiv(i1 - i0, 256), 256, 0, st>>>(keys_w, first_s, i0, i1, (F.rr0 > 0 && ranges.size() == 1) ? 1 : 0);
        ctx->launches++;
      }
    }
    CK(cudaEventRecordWithFlags(ctx->ev[2], st, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    const bool stores00 = F.rr0 > 0 && ranges.size() == 1 && (fr->flags & PRC_FRAME_KEEP_GBUFFER);
    if (!fused || stores00) {
    KTimer kt(ctx, PRC_K_RESOLVE);
    if (!fused) {
      for (const auto& rg : ranges) {
        DevFrame Fr = F;
        Fr.rr0 = rg.first; Fr.rr1 = rg.second;
        const dim3 rgd((F.W + 31) / 32, (Fr.rr1 - Fr.rr0 + 3) / 4);
        if (es) k_resolve<E, E><<<rgd, 128, 0, st>>>(ctx->S, Fr, keys, G);
        else k_resolve<E, false><<<rgd, 128, 0, st>>>(ctx->S, Fr, keys, G);
        ctx->launches++;
      }
    }
    // k_shade_special resolves pixel (0,0) itself; only a G-buffer readback needs it STORED when it lies outside the rasterised rows
    if (stores00) {
      if (es) k_resolve00<E, E><<<1, 1, 0, st>>>(ctx->S, F, keys, G);
      else k_resolve00<E, false><<<1, 1, 0, st>>>(ctx->S, F, keys, G);
      ctx->launches++;
    }
    }
  }

  if (phases & 2) {
    if (ctx->copy_pending) {  // (peer frames) the previous frame's strip is still leaving through the copy stream: it reads d_image
      CK(cudaStreamWaitEvent(st, ctx->ev_copied, 0));
      ctx->copy_pending = false;
    }
    {
      KTimer kts(ctx, getenv("PRC_TIME_SPECIAL") ? PRC_K_RESOLVE : -1);
      if (es) k_shade_special<E, E><<<1, 32, 0, st>>>(ctx->S, F, aoc, keys, G, (uint32_t*)ctx->d_special.p);
      else k_shade_special<E, false><<<1, 32, 0, st>>>(ctx->S, F, aoc, keys, G, (uint32_t*)ctx->d_special.p);
      ctx->launches++;
    }
    KTimer kt(ctx, PRC_K_SHADE);
    // With a readback pending the strip is shaded in PRC_SHADE_BANDS row bands, top image rows first; each band's
    // device->host DMA runs on the copy stream while the next band is shaded (only the last band's copy is exposed).
    bool banded_copy = ctx->rb_dst && ctx->msaa == 1;
    uint32_t* host_image = nullptr;
    if (banded_copy && ctx->zero_copy_out) {
      void* dp = nullptr;
      if (cudaHostGetDevicePointer(&dp, ctx->rb_dst, 0) == cudaSuccess && dp) { host_image = (uint32_t*)dp; banded_copy = false; }
      else (void)cudaGetLastError();
    }
    const int rows = F.row1 - F.row0;
    const int nb = (banded_copy && rows >= 64 * ctx->shade_bands) ? ctx->shade_bands : 1;
    const int band = ((rows + nb - 1) / nb + 3) & ~3;
    for (int b = 0; b < nb; b++) {
      DevFrame Fb = F;
      Fb.row1 = F.row1 - b * band;  // image row r = screen y = H-1-r: the highest screen rows are the first image rows
      Fb.row0 = std::max(F.row0, Fb.row1 - band);
      if (Fb.row0 >= Fb.row1) break;
      const dim3 sg((F.W + 31) / 32, (Fb.row1 - Fb.row0 + 3) / 4);
      if (fused) {
        if (es) k_resolve_shade<E, E><<<sg, 128, 0, st>>>(ctx->S, Fb, aoc, keys, special, image, host_image);
        else k_resolve_shade<E, false><<<sg, 128, 0, st>>>(ctx->S, Fb, aoc, keys, special, image, host_image);
      } else {
        if (es) k_shade<true><<<sg, 128, 0, st>>>(ctx->S, Fb, aoc, keys, G, special, image, host_image);
        else k_shade<false><<<sg, 128, 0, st>>>(ctx->S, Fb, aoc, keys, G, special, image, host_image);
      }
      ctx->launches++;
      if (banded_copy) {
        const size_t off = (size_t)(F.H - Fb.row1) * F.W * 4, bytes = (size_t)(Fb.row1 - Fb.row0) * F.W * 4;
        CK(cudaEventRecord(ctx->ev_band[b], st));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_band[b], 0));
        CK(cudaMemcpyAsync(ctx->rb_dst + off, (uint8_t*)ctx->d_image.p + off, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
      }
    }
    if (banded_copy) {  // the frame's stream ends after the last copy
      CK(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
      // peer frames are submitted back to back: the NEXT frame's geometry may run under this copy, only its shading (which
      // rewrites d_image) and prc_sync wait for it
      if (ctx->defer_copy_join) ctx->copy_pending = true;
      else CK(cudaStreamWaitEvent(st, ctx->ev_copied, 0));
    }
    if (ctx->msaa > 1) {
      // passAntialiasing (raster.go:377): r.outBuf = imageutil.Resize(cfg.Width, cfg.Height, CurrBuffer().Image())
      const int ow = F.W / ctx->msaa, oh = F.H / ctx->msaa;
      const int32_t rr = ensure_resize_tables(ctx, F.W, F.H, ow, oh);
      if (rr != PRC_OK) return rr;
      ENSURE(ctx->d_image_out, (size_t)ow * oh * 4);
      // output rows of the strip this context owns (image row r = screen y = H-1-r); the whole frame on one GPU
      const int o0 = (F.H - ctx->own_row1) / ctx->msaa, o1 = (F.H - ctx->own_row0) / ctx->msaa;
      if (o0 == 0 && o1 == oh) {
        k_resize<<<dim3((ow + 31) / 32, (oh + 7) / 8), 256, 0, st>>>(image, F.W, F.H, (uint32_t*)ctx->d_image_out.p, ow, oh,
                                                                     (const short*)ctx->d_rz_cx.p, (const int*)ctx->d_rz_sx.p, ctx->rz_flx,
                                                                     (const short*)ctx->d_rz_cy.p, (const int*)ctx->d_rz_sy.p, ctx->rz_fly);
        ctx->launches++;
        if (ctx->rb_dst) CK(cudaMemcpyAsync(ctx->rb_dst, ctx->d_image_out.p, (size_t)ow * oh * 4, cudaMemcpyDeviceToHost, st));
      } else if (o1 > o0) {
        // a strip: the same kernel on the sub-range of output rows (its row tables and its output are offset, the
        // supersampled input rows it reads are absolute and were shaded with the halo above)
        k_resize<<<dim3((ow + 31) / 32, (o1 - o0 + 7) / 8), 256, 0, st>>>(
            image, F.W, F.H, (uint32_t*)ctx->d_image_out.p + (size_t)o0 * ow, ow, o1 - o0, (const short*)ctx->d_rz_cx.p, (const int*)ctx->d_rz_sx.p,
            ctx->rz_flx, (const short*)ctx->d_rz_cy.p + (size_t)o0 * ctx->rz_fly, (const int*)ctx->d_rz_sy.p + o0, ctx->rz_fly);
        ctx->launches++;
        if (ctx->rb_dst)
          CK(cudaMemcpyAsync(ctx->rb_dst + (size_t)o0 * ow * 4, (const uint8_t*)ctx->d_image_out.p + (size_t)o0 * ow * 4, (size_t)(o1 - o0) * ow * 4,
                             cudaMemcpyDeviceToHost, st));
      }
    }
    ctx->gbuffer_valid = !fused;
  }
  CK(cudaGetLastError());
  return PRC_OK;
}

// Two library-owned page-locked images used alternately, like the reference's double buffer (raster.go:86,
// 201-206): the device->host copy is DMA at PCIe speed; a caller that passes rgba_out == NULL reads the frame in
// place through prc_host_image (zero copy, valid until two frames later), one that passes its own buffer pays an
// extra host memcpy. readback_begin() picks this frame's buffer BEFORE the frame is enqueued (do_main issues the
// copies band by band behind the shading kernels); readback_end() waits for them.
int32_t readback_begin(prc_ctx* ctx, const prc_frame* fr) {
  ctx->rb_dst = nullptr;
  if (fr->flags & PRC_FRAME_NO_READBACK) return PRC_OK;
  const uint32_t ms = fr->msaa > 1 ? fr->msaa : 1;
  const size_t total = (size_t)(fr->width / ms) * (fr->height / ms) * 4;  // the frame handed back (after the MSAA downsample)
  if (ctx->h_img_cap < total) {
    CK(cudaStreamSynchronize(ctx->stream));  // (a view batch may still be copying into the old images)
    for (auto& p : ctx->h_img) { if (p) { cudaHostUnregister(p); free(p); } p = nullptr; }
    for (auto& p : ctx->h_img) {
      // first-touched by this thread (NUMA-local), then page-locked
      void* q = nullptr;
      if (posix_memalign(&q, 4096, (total + 4095) & ~(size_t)4095) != 0) { ctx->err = "out of host memory"; return PRC_ERR_CUDA; }
      memset(q, 0, total);
      p = (uint8_t*)q;
      CK(cudaHostRegister(p, (total + 4095) & ~(size_t)4095, cudaHostRegisterMapped));  // mapped: PRC_ZERO_COPY_OUT stores into it from the shading kernel
    }
    ctx->h_img_cap = total;
  }
  ctx->h_img_cur ^= 1;
  ctx->rb_dst = ctx->h_img[ctx->h_img_cur];
  return PRC_OK;
}

int32_t readback_end(prc_ctx* ctx, const DevFrame& F, uint8_t* rgba_out) {
  // image rows of the strip: screen rows [row0,row1) -> image rows [H-row1, H-row0)
  size_t off = (size_t)(F.H - F.row1) * F.W * 4, bytes = (size_t)(F.row1 - F.row0) * F.W * 4;
  if (ctx->msaa > 1) { off = 0; bytes = (size_t)(F.W / ctx->msaa) * (F.H / ctx->msaa) * 4; }
  CK(cudaStreamSynchronize(ctx->stream));  // the copies were joined into the frame's stream by do_main
  if (rgba_out && ctx->rb_dst) memcpy(rgba_out + off, ctx->rb_dst + off, bytes);
  return PRC_OK;
}

int32_t finish_timings(prc_ctx* ctx) {
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->copy_pending) {  // a peer frame's strip readback runs on the copy stream
    CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->copy_pending = false;
  }
  if (ctx->img_pending) {  // PRC_FRAME_IMAGE_AT_SYNC: the peers' strips of the last frames (the side stream's waits also end by timeout)
    CK(cudaStreamSynchronize(ctx->img_stream));
    ctx->img_pending = false;
  }
  float a = 0, b = 0, c = 0, d = 0;
  cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&d, ctx->ev[0], ctx->ev[3]);
  ctx->tm.abi_version = PRC_ABI_VERSION;
  ctx->tm.shadow_ms = a; ctx->tm.forward_ms = b; ctx->tm.shade_ms = c; ctx->tm.total_ms = d;
  ctx->tm.n_valid_tris = ctx->n_valid;
  CK(cudaMemcpy(ctx->h_counters, ctx->d_counters.p, sizeof(Counters), cudaMemcpyDeviceToHost));
  ctx->tm.n_nan_frags = ctx->h_counters->n_nan + ctx->h_counters->n_nan_shadow;
  // Conditions that invalidate the frame and are cured by rendering it again (every caller loops on PRC_RETRY; frames submitted
  // back to back surface it as PRC_ERR_RETRY from prc_sync). All of them are handled in one go.
  bool retry = false;
  if (ctx->h_counters->n_nan && !ctx->nan_mode && !getenv("PRC_NO_NAN_MODE")) {
    // a NaN-depth fragment in the camera pass: whether it shows depends on the draw order (bug-list 8) — NaN mode adds the
    // first-fragment plane behind the keys (the stream is idle here; the frame is re-rendered from its key clear on)
    ctx->nan_mode = true;
    ENSURE(ctx->d_keys, (size_t)ctx->W * ctx->H * 16);
    retry = true;
  }
  if (ctx->h_counters->need_bins && !ctx->need_bins) {
    // a queued record was too large for k_medium_raster and the binned tile path was off: switch it on for good
    ctx->need_bins = true;
    retry = true;
  }
  if (ctx->h_counters->large_overflow) {
    if (ctx->h_counters->max_bins > ctx->bins_cap) {  // grow the bin array
      ENSURE(ctx->d_bins, (size_t)ctx->h_counters->max_bins * 5);
      ctx->bins_cap = (unsigned int)(ctx->d_bins.cap / 4);
      retry = true;
    }
    if (ctx->h_counters->stat_large > ctx->large_cap) {  // large-triangle queue (stat_large = records pushed over the frame)
      const size_t want = (size_t)ctx->h_counters->stat_large * 5 / 4 + 1024;
      ENSURE(ctx->d_large, want * sizeof(LargeRec));
      ctx->large_cap = (unsigned int)std::min<size_t>(ctx->d_large.cap / sizeof(LargeRec), 0xFFFFFFF0u);
      retry = true;
    }
    if (!retry) {
      ctx->spans.clear();
      ctx->ev_used = 0;
      ctx->err = "internal queue overflow (large/clip queue)";
      return PRC_ERR_UNSUPPORTED;
    }
  }
  if (retry) {
    ctx->spans.clear();
    ctx->ev_used = 0;
    return PRC_RETRY;
  }
  ctx->tm.gpu_launches = ctx->launches;
  ctx->tm.n_large_items = ctx->h_counters->stat_large; ctx->tm.n_clipped = ctx->h_counters->stat_clip; ctx->tm.n_bin_entries = ctx->h_counters->stat_bins;
  for (int k = 0; k < 8; k++) { ctx->tm.kernel_ms[k] = 0; ctx->tm.kernel_launches[k] = 0; }
  for (auto& sp : ctx->spans) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->evpool[sp.a], ctx->evpool[sp.b]);
    ctx->tm.kernel_ms[sp.cls] += ms;
    ctx->tm.kernel_launches[sp.cls]++;
  }
  ctx->spans.clear();
  ctx->ev_used = 0;
  return PRC_OK;
}

}  // namespace

extern "C" {

uint32_t prc_abi_version(void) { return PRC_ABI_VERSION; }

int32_t prc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int32_t prc_open(int32_t device, prc_ctx** out) {
  if (!out) return PRC_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return PRC_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return PRC_ERR_CUDA;
  prc_ctx* ctx = new prc_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PRC_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PRC_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PRC_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->img_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PRC_ERR_CUDA; }
  cudaEventCreateWithFlags(&ctx->ev_own_shaded, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_image_done, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  for (auto& e : ctx->ev) cudaEventCreate(&e);
  for (auto& e : ctx->ev_band) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming);
  cudaMallocHost((void**)&ctx->h_counters, sizeof(Counters));
  cudaMalloc(&ctx->d_counters.p, sizeof(Counters));
  ctx->d_counters.cap = sizeof(Counters);
  cudaMemset(ctx->d_counters.p, 0, sizeof(Counters));
  const char* mode = getenv("PRC_FMA");
  ctx->exact = !(mode && strcmp(mode, "fast") == 0);
  ctx->exact_shade = mode && strcmp(mode, "exact") == 0;
  if (cudaHostAlloc((void**)&ctx->h_list_hint, 16 * sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) memset(ctx->h_list_hint, 0, 16 * sizeof(unsigned int));
  else { ctx->h_list_hint = nullptr; (void)cudaGetLastError(); }
  ctx->no_fused_shade = getenv("PRC_NO_FUSED_SHADE") != nullptr;
  ctx->no_chunk_cull = getenv("PRC_NO_CHUNK_CULL") != nullptr;
  ctx->force_chunk_cull = getenv("PRC_FORCE_CHUNK_CULL") != nullptr;
  ctx->need_bins = getenv("PRC_FORCE_BINS") != nullptr;
  ctx->ktimers = !(getenv("PRC_NO_KTIMERS") && atoi(getenv("PRC_NO_KTIMERS")) != 0);
  if (const char* sb = getenv("PRC_SHADE_BANDS")) ctx->shade_bands = std::max(1, std::min(PRC_SHADE_BANDS_MAX, atoi(sb)));
  if (const char* pt = getenv("PRC_PEER_TIMEOUT_MS")) ctx->peer_timeout_ns = (unsigned long long)std::max(1, atoi(pt)) * 1000000ull;
  ctx->two_streams = getenv("PRC_TWO_STREAMS") != nullptr && atoi(getenv("PRC_TWO_STREAMS")) != 0;
  ctx->zero_copy_out = getenv("PRC_ZERO_COPY_OUT") != nullptr && atoi(getenv("PRC_ZERO_COPY_OUT")) != 0;
  ctx->stage_single = getenv("PRC_STAGE_UNIFORMS") != nullptr && atoi(getenv("PRC_STAGE_UNIFORMS")) != 0;
  // AO constants (material/ao.go:28-32): a accumulates float32(Pi/4) in float32; Cos/Sin via float64.
  const float pi = 3.14159265358979323846f, q = pi / 4;
  float a = 0.0f;
  for (int k = 0; k < 8; k++) {
    ctx->ao.cosv[k] = (float)std::cos((double)a);
    ctx->ao.sinv[k] = (float)std::sin((double)a);
    a += q;
  }
  ctx->ao.half_pi = pi / 2;
  ctx->ao.four_pi = pi * 4;
  cudaMalloc(&ctx->d_aoc.p, sizeof(AoConsts));
  ctx->d_aoc.cap = sizeof(AoConsts);
  cudaMemcpy(ctx->d_aoc.p, &ctx->ao, sizeof(AoConsts), cudaMemcpyHostToDevice);
  *out = ctx;
  return PRC_OK;
}

int32_t prc_close(prc_ctx* ctx) {
  if (!ctx) return PRC_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  peer_release(ctx);
  free_buf(ctx->d_chunklist);
  if (ctx->ext_img && ctx->ext_img_registered) cudaHostUnregister(ctx->ext_img);
  free_buf(ctx->d_peer_signals);
  free_buf(ctx->d_peer_err);
  DBuf* all[] = {&ctx->d_pos, &ctx->d_nor, &ctx->d_uv, &ctx->d_col, &ctx->d_mat, &ctx->d_meta, &ctx->d_mats, &ctx->d_objstart, &ctx->d_texfirst,
                 &ctx->d_lw, &ctx->d_lh, &ctx->d_loff, &ctx->d_tex, &ctx->d_keys, &ctx->d_ga, &ctx->d_gb, &ctx->d_gc, &ctx->d_gd, &ctx->d_ao,
                 &ctx->d_image, &ctx->d_special, &ctx->d_counters, &ctx->d_large, &ctx->d_clipq, &ctx->d_tilecount, &ctx->d_tilestart, &ctx->d_cursor,
                 &ctx->d_bins, &ctx->d_active, &ctx->d_chunksum, &ctx->d_targets, &ctx->d_xf, &ctx->d_lights, &ctx->d_ambient, &ctx->d_gamma, &ctx->d_frame, &ctx->d_aoc, &ctx->d_chunkbox, &ctx->d_vis, &ctx->d_cverts, &ctx->d_cvoff, &ctx->d_lidx, &ctx->d_image_out, &ctx->d_rz_cx, &ctx->d_rz_sx, &ctx->d_rz_cy, &ctx->d_rz_sy};
  for (DBuf* b : all) free_buf(*b);
  free_buf(ctx->d_shadow_all);
  free_buf(ctx->d_shadow_mine);
  free_buf(ctx->d_mkeys);
  free_buf(ctx->d_dirty);
  for (auto& b : ctx->d_shadow_trans) free_buf(b);
  for (auto& p : ctx->h_img) if (p) { cudaHostUnregister(p); free(p); }
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  if (ctx->stage) cudaFreeHost(ctx->stage);
  for (auto& e : ctx->ev_view) if (e) cudaEventDestroy(e);
  if (ctx->h_list_hint) cudaFreeHost(ctx->h_list_hint);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->evpool) cudaEventDestroy(e);
  for (auto& e : ctx->ev_band) if (e) cudaEventDestroy(e);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
  if (ctx->img_stream) cudaStreamDestroy(ctx->img_stream);
  if (ctx->ev_own_shaded) cudaEventDestroy(ctx->ev_own_shaded);
  if (ctx->ev_image_done) cudaEventDestroy(ctx->ev_image_done);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PRC_OK;
}

const char* prc_last_error(prc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int32_t prc_set_exact_fma(prc_ctx* ctx, int32_t exact) {
  if (!ctx) return PRC_ERR_INVALID;
  ctx->exact = exact != 0;
  ctx->exact_shade = exact != 0;
  return PRC_OK;
}

int32_t prc_scene_upload(prc_ctx* ctx, const prc_scene* s) {
  if (!ctx) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  if (!s || s->abi_version != PRC_ABI_VERSION) { ctx->err = "prc_scene: bad abi_version"; return PRC_ERR_INVALID; }
  if (s->n_tris >= (1ull << 29)) { ctx->err = "too many triangles (limit 2^29)"; return PRC_ERR_UNSUPPORTED; }
  if (s->n_objects >= (1u << 24)) { ctx->err = "too many objects (limit 2^24)"; return PRC_ERR_UNSUPPORTED; }
  for (uint32_t i = 0; i < s->n_materials; i++) {
    const prc_material& m = s->materials[i];
    if (!(m.flags & PRC_MAT_NIL) && (m.texture < 0 || (uint32_t)m.texture >= s->n_textures)) { ctx->err = "material without texture"; return PRC_ERR_INVALID; }
  }
  CK(cudaSetDevice(ctx->device));
  ctx->has_scene = false;
  const uint64_t n = s->n_tris;
  ENSURE(ctx->d_pos, n * 36 + 16);  // (+16: slack for vectorised reads of the last triangle)
  UPLOAD(ctx->d_pos, s->pos, n * 36);
  UPLOAD(ctx->d_nor, s->nor, n * 36);
  UPLOAD(ctx->d_uv, s->uv, n * 24);
  UPLOAD(ctx->d_col, s->col, n * 12);
  UPLOAD(ctx->d_mat, s->mat, n * 4);
  UPLOAD(ctx->d_mats, s->materials, (size_t)s->n_materials * sizeof(prc_material));
  UPLOAD(ctx->d_objstart, s->obj_tri_start, (size_t)(s->n_objects + 1) * 8);
  UPLOAD(ctx->d_texfirst, s->tex_first_level, (size_t)(s->n_textures + 1) * 4);
  UPLOAD(ctx->d_lw, s->level_w, (size_t)s->n_tex_levels * 4);
  UPLOAD(ctx->d_lh, s->level_h, (size_t)s->n_tex_levels * 4);
  UPLOAD(ctx->d_loff, s->level_offset, (size_t)s->n_tex_levels * 8);
  UPLOAD(ctx->d_tex, s->tex_data, s->tex_bytes);
  ENSURE(ctx->d_meta, n * 4);
  ctx->large_cap = (unsigned int)std::min<uint64_t>(std::max<uint64_t>(n / 4, 1u << 20), 0xFFFFFFF0ull);  // grown on demand (frame re-render)
  ctx->clip_cap = (unsigned int)std::min<uint64_t>(n + 16, 0xFFFFFFF0ull);
  ENSURE(ctx->d_large, (size_t)ctx->large_cap * sizeof(LargeRec));
  ENSURE(ctx->d_clipq, (size_t)ctx->clip_cap * 4);
  ctx->any_ao = false;
  ctx->nan_mode = false;
  for (uint32_t i = 0; i < s->n_materials; i++)
    if (!(s->materials[i].flags & PRC_MAT_NIL) && (s->materials[i].flags & PRC_MAT_AMBIENT_OCCLUSION)) ctx->any_ao = true;
  // Triangle.IsValid + object index, once per scene
  Counters* cnt = (Counters*)ctx->d_counters.p;
  CK(cudaMemsetAsync(cnt, 0, offsetof(Counters, dirty), ctx->stream));
  if (n) k_validate<<<cdiv(n, 256), 256, 0, ctx->stream>>>((const float*)ctx->d_pos.p, (const uint64_t*)ctx->d_objstart.p, s->n_objects, n,
                                                           (uint32_t*)ctx->d_meta.p, &cnt->n_valid);
  ctx->n_chunks = cdiv(n, PRC_GEOM_THREADS);
  ENSURE(ctx->d_chunkbox, (size_t)std::max<uint32_t>(1, ctx->n_chunks) * sizeof(ChunkBox));
  ENSURE(ctx->d_vis, (size_t)std::max<uint32_t>(1, ctx->n_chunks) * 9);
  if (n) k_chunk_aabb<<<ctx->n_chunks, 256, 0, ctx->stream>>>((const float*)ctx->d_pos.p, (const uint32_t*)ctx->d_meta.p, n, (ChunkBox*)ctx->d_chunkbox.p);
  CK(cudaMemcpyAsync(ctx->h_counters, cnt, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
  // chunk-local vertex sharing (k_chunk_dedupe): count, scan on the host (n/256 integers, once per scene), fill
  ENSURE(ctx->d_cvoff, (size_t)(ctx->n_chunks + 1) * 4);
  ENSURE(ctx->d_lidx, std::max<uint64_t>(1, n) * 4);
  uint64_t n_cverts = 0;
  {
    std::vector<uint32_t> off((size_t)ctx->n_chunks + 1, 0u);
    if (n) {
      k_chunk_dedupe<<<ctx->n_chunks, 256, 0, ctx->stream>>>((const float*)ctx->d_pos.p, (const uint32_t*)ctx->d_meta.p, n, nullptr, (uint32_t*)ctx->d_cvoff.p,
                                                             nullptr, nullptr);
      CK(cudaMemcpyAsync(off.data(), ctx->d_cvoff.p, (size_t)ctx->n_chunks * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (uint32_t c = 0; c < ctx->n_chunks; c++) { const uint32_t k = off[c]; off[c] = (uint32_t)n_cverts; n_cverts += k; }
      off[ctx->n_chunks] = (uint32_t)n_cverts;  // < 3 * 2^29 fits 32 bits
    }
    CK(cudaMemcpyAsync(ctx->d_cvoff.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    ENSURE(ctx->d_cverts, std::max<uint64_t>(1, n_cverts) * 16);
    if (n)
      k_chunk_dedupe<<<ctx->n_chunks, 256, 0, ctx->stream>>>((const float*)ctx->d_pos.p, (const uint32_t*)ctx->d_meta.p, n, (const uint32_t*)ctx->d_cvoff.p,
                                                             nullptr, (float4*)ctx->d_cverts.p, (uint32_t*)ctx->d_lidx.p);
    CK(cudaStreamSynchronize(ctx->stream));  // `off` is pageable host memory
  }
  ctx->n_cverts = n_cverts;
  CK(cudaGetLastError());
  ctx->n_valid = ctx->h_counters->n_valid;
  DevScene& S = ctx->S;
  S.pos = (const float*)ctx->d_pos.p; S.nor = (const float*)ctx->d_nor.p; S.uv = (const float*)ctx->d_uv.p;
  S.col = (const uint32_t*)ctx->d_col.p; S.mat = (const int32_t*)ctx->d_mat.p; S.meta = (const uint32_t*)ctx->d_meta.p;
  S.n_tris = n;
  S.cverts = (const float4*)ctx->d_cverts.p; S.cvoff = (const uint32_t*)ctx->d_cvoff.p; S.lidx = (const uint32_t*)ctx->d_lidx.p;
  S.mats = (const prc_material*)ctx->d_mats.p; S.n_mats = s->n_materials;
  S.tex_first = (const uint32_t*)ctx->d_texfirst.p; S.level_w = (const uint32_t*)ctx->d_lw.p; S.level_h = (const uint32_t*)ctx->d_lh.p;
  S.level_off = (const uint64_t*)ctx->d_loff.p; S.tex_data = (const uint8_t*)ctx->d_tex.p;
  ctx->n_objects = s->n_objects;
  ctx->has_scene = true;
  ctx->gbuffer_valid = false;
  return PRC_OK;
}

int32_t prc_shadow_reset(prc_ctx* ctx) {
  if (!ctx) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  CK(cudaSetDevice(ctx->device));
  if (ctx->d_shadow_all.p) CK(cudaMemsetAsync(ctx->d_shadow_all.p, 0, ctx->d_shadow_all.cap, ctx->stream));
  if (ctx->d_shadow_mine.p) CK(cudaMemsetAsync(ctx->d_shadow_mine.p, 0, ctx->d_shadow_mine.cap, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PRC_OK;
}

static int32_t render_shadow_units(prc_ctx* ctx, const prc_frame* fr, const std::vector<ShadowUnit>& units) {
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  CK(cudaSetDevice(ctx->device));
  DevFrame F;
  int32_t r = build_frame(ctx, fr, F);
  if (r != PRC_OK) return r;
  for (const ShadowUnit& u : units)
    if (u.r1 > (int)fr->height || u.r0 < 0 || u.r0 >= u.r1) { ctx->err = "bad shadow row range"; return PRC_ERR_INVALID; }
  for (int attempt = 0; attempt < 4; attempt++) {
    ctx->launches = 0;
    ctx->pass_slot = 0;
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, offsetof(Counters, n_valid), ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    r = ctx->exact ? do_shadows<true>(ctx, fr, F, units, true) : do_shadows<false>(ctx, fr, F, units, true);
    if (r != PRC_OK) return r;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->h_counters->large_overflow && !(ctx->h_counters->need_bins && !ctx->need_bins)) return PRC_OK;
    bool grown = false;
    if (ctx->h_counters->need_bins && !ctx->need_bins) { ctx->need_bins = true; grown = true; }
    if (ctx->h_counters->max_bins > ctx->bins_cap) {
      ENSURE(ctx->d_bins, (size_t)ctx->h_counters->max_bins * 5);
      ctx->bins_cap = (unsigned int)(ctx->d_bins.cap / 4);
      grown = true;
    }
    if (ctx->h_counters->stat_large > ctx->large_cap) {
      ENSURE(ctx->d_large, ((size_t)ctx->h_counters->stat_large * 5 / 4 + 1024) * sizeof(LargeRec));
      ctx->large_cap = (unsigned int)std::min<size_t>(ctx->d_large.cap / sizeof(LargeRec), 0xFFFFFFF0u);
      grown = true;
    }
    if (!grown) { ctx->err = "internal queue overflow (clip queue)"; return PRC_ERR_UNSUPPORTED; }
    ctx->spans.clear();
    ctx->ev_used = 0;
  }
  ctx->err = "bin array kept overflowing";
  return PRC_ERR_UNSUPPORTED;
}

int32_t prc_render_shadows(prc_ctx* ctx, const prc_frame* fr, uint32_t light_mask, uint32_t srow0, uint32_t srow1) {
  if (!ctx) return PRC_ERR_INVALID;
  if (!fr || fr->abi_version != PRC_ABI_VERSION) { ctx->err = "prc_frame: bad abi_version"; return PRC_ERR_INVALID; }
  return render_shadow_units(ctx, fr, units_from_mask(fr, light_mask, (int)srow0, (int)srow1));
}

int32_t prc_render_shadow_units(prc_ctx* ctx, const prc_frame* fr, uint32_t n, const uint32_t* light, const uint32_t* row0, const uint32_t* row1) {
  if (!ctx) return PRC_ERR_INVALID;
  std::vector<ShadowUnit> u;
  for (uint32_t k = 0; k < n; k++) u.push_back({light[k], (int)row0[k], (int)row1[k]});
  return render_shadow_units(ctx, fr, u);
}

int32_t prc_render_main(prc_ctx* ctx, const prc_frame* fr, uint8_t* rgba_out) {
  if (!ctx) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  CK(cudaSetDevice(ctx->device));
  DevFrame F;
  int32_t r = build_frame(ctx, fr, F);
  if (r != PRC_OK) return r;
  r = readback_begin(ctx, fr);
  if (r != PRC_OK) return r;
  for (int attempt = 0; attempt < 4; attempt++) {
    ctx->pass_slot = 8;  // (the split-phase shadow call of this frame used the first slots)
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, offsetof(Counters, n_nan), ctx->stream));
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    r = ctx->exact ? do_main<true>(ctx, fr, F) : do_main<false>(ctx, fr, F);
    if (r != PRC_OK) return r;
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (!(fr->flags & PRC_FRAME_NO_READBACK)) {
      r = readback_end(ctx, F, rgba_out);
      if (r != PRC_OK) return r;
    }
    r = finish_timings(ctx);
    if (r != PRC_RETRY) return r;
  }
  ctx->err = "bin array kept overflowing";
  return PRC_ERR_UNSUPPORTED;
}

// Split of prc_render_main for overlapping the shadow-map exchange with the camera pass:
//   prc_render_forward  = camera geometry + raster + tile path + resolve (needs no shadow map), asynchronous
//   prc_render_deferred = shading (+ optional readback), synchronises and reports timings / queue overflow
int32_t prc_render_forward(prc_ctx* ctx, const prc_frame* fr) {
  if (!ctx) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  CK(cudaSetDevice(ctx->device));
  DevFrame F;
  int32_t r = build_frame(ctx, fr, F);
  if (r != PRC_OK) return r;
  ctx->rb_dst = nullptr;
  ctx->pass_slot = 8;
  CK(cudaMemsetAsync(ctx->d_counters.p, 0, offsetof(Counters, n_nan), ctx->stream));
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  return ctx->exact ? do_main<true>(ctx, fr, F, 13) : do_main<false>(ctx, fr, F, 13);
}

int32_t prc_render_deferred(prc_ctx* ctx, const prc_frame* fr, uint8_t* rgba_out) {
  if (!ctx) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  CK(cudaSetDevice(ctx->device));
  DevFrame F;
  int32_t r = build_frame(ctx, fr, F);
  if (r != PRC_OK) return r;
  r = readback_begin(ctx, fr);
  if (r != PRC_OK) return r;
  r = ctx->exact ? do_main<true>(ctx, fr, F, 2) : do_main<false>(ctx, fr, F, 2);
  if (r != PRC_OK) return r;
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  if (!(fr->flags & PRC_FRAME_NO_READBACK)) {
    r = readback_end(ctx, F, rgba_out);
    if (r != PRC_OK) return r;
  }
  r = finish_timings(ctx);
  if (r == PRC_RETRY) {  // a queue overflowed during the forward phase: redo both phases with the grown queues
    r = prc_render_main(ctx, fr, rgba_out);
  }
  return r;
}

// The key clear of do_main (phase 1), issued on `s`: the rasterised rows (+ the AO rows of pixel (0,0), + pixel (0,0) itself).
static int32_t clear_keys(prc_ctx* ctx, const DevFrame& F, cudaStream_t s) {
  unsigned long long* keys = (unsigned long long*)ctx->d_keys.p;
  unsigned long long* first = keys + (size_t)F.W * F.H;
  int r0s[2] = {F.rr0, 0}, r1s[2] = {F.rr1, (ctx->any_ao && F.rr0 > 0) ? std::min(100, F.rr0) : 0};
  for (int k = 0; k < 2; k++) {
    if (r1s[k] <= r0s[k]) continue;
    CK(cudaMemsetAsync(keys + (size_t)r0s[k] * F.W, 0, (size_t)(r1s[k] - r0s[k]) * F.W * 8, s));
    if (ctx->nan_mode) CK(cudaMemsetAsync(first + (size_t)r0s[k] * F.W, 0xFF, (size_t)(r1s[k] - r0s[k]) * F.W * 8, s));
  }
  if (F.rr0 > 0 && r1s[1] == 0) {
    CK(cudaMemsetAsync(keys, 0, 8, s));
    if (ctx->nan_mode) CK(cudaMemsetAsync(first, 0xFF, 8, s));
  }
  return PRC_OK;
}

// the launches of one whole frame (shadow sweeps, camera pass, tile path, resolve, shading)
static int32_t enqueue_frame(prc_ctx* ctx, const prc_frame* fr, const DevFrame& F) {
  const unsigned int rec = ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault;
  ctx->pass_slot = 0;
  if (ctx->pending_async == 0) {
    ctx->launches = 0;
    ctx->spans.clear();
    ctx->ev_used = 0;
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, offsetof(Counters, n_valid), ctx->stream));
  } else {
    // behind unfinished asynchronous frames: keep the sticky overflow flag / statistics, reset the per-pass counters and list lengths only
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, offsetof(Counters, large_overflow), ctx->stream));
  }
  CK(cudaEventRecordWithFlags(ctx->ev[0], ctx->stream, rec));
  int32_t r;
  if (ctx->two_streams && !ctx->capturing && (fr->flags & PRC_FRAME_SHADOWMAP) && !(fr->flags & PRC_FRAME_KEEP_GBUFFER)) {
    // (tuning) the shadow sweep and the camera pass write different targets and share only the atomically filled record queue:
    // side by side on two streams each fills the other's tail waves; the queued records are rasterised after the join
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->copy_stream2, ctx->ev_fork, 0));
    ctx->stream = ctx->copy_stream2;
    r = ctx->exact ? do_shadows<true>(ctx, fr, F, units_from_mask(fr, 0xFFFFFFFFu, 0, F.H), false)
                   : do_shadows<false>(ctx, fr, F, units_from_mask(fr, 0xFFFFFFFFu, 0, F.H), false);
    cudaEventRecord(ctx->ev_join, ctx->copy_stream2);
    ctx->stream = st;
    if (r == PRC_OK) r = ctx->exact ? do_main<true>(ctx, fr, F, 1, 1) : do_main<false>(ctx, fr, F, 1, 1);
    CK(cudaEventRecordWithFlags(ctx->ev[1], st, rec));
    CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    if (r == PRC_OK) r = ctx->exact ? do_main<true>(ctx, fr, F, 4 | 8 | 2, 1) : do_main<false>(ctx, fr, F, 4 | 8 | 2, 1);
    if (r != PRC_OK) return r;
    CK(cudaEventRecordWithFlags(ctx->ev[3], st, rec));
    return PRC_OK;
  }
  // The 66 MB key clear of a 4K frame (0.02 ms) is only needed by the camera pass: it runs on a second stream under the shadow
  // sweep, which never touches the keys (forked here = behind the previous frame's shading, joined before the camera pass).
  static const bool early_clear_on = !(getenv("PRC_NO_EARLY_CLEAR") != nullptr && atoi(getenv("PRC_NO_EARLY_CLEAR")) != 0);
  const bool early_clear = early_clear_on && !ctx->capturing && (fr->flags & PRC_FRAME_SHADOWMAP) && ctx->n_cast_alloc > 0 && ctx->n_cast_alloc != 0xFFFFFFFFu;
  if (early_clear) {
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream2, ctx->ev_fork, 0));
    r = clear_keys(ctx, F, ctx->copy_stream2);
    if (r != PRC_OK) return r;
    CK(cudaEventRecord(ctx->ev_join, ctx->copy_stream2));
  }
  r = ctx->exact ? do_shadows<true>(ctx, fr, F, units_from_mask(fr, 0xFFFFFFFFu, 0, F.H), false)
                 : do_shadows<false>(ctx, fr, F, units_from_mask(fr, 0xFFFFFFFFu, 0, F.H), false);
  if (r != PRC_OK) return r;
  CK(cudaEventRecordWithFlags(ctx->ev[1], ctx->stream, rec));
  if (early_clear) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    ctx->skip_key_clear = true;
  }
  r = ctx->exact ? do_main<true>(ctx, fr, F) : do_main<false>(ctx, fr, F);
  ctx->skip_key_clear = false;
  if (r != PRC_OK) return r;
  CK(cudaEventRecordWithFlags(ctx->ev[3], ctx->stream, rec));
  return PRC_OK;
}

int32_t prc_render(prc_ctx* ctx, const prc_frame* fr, uint8_t* rgba_out) {
  if (!ctx) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ctx->peer_private) { ctx->err = "this context is connected to a peer group (its raster targets are the group's private buffers): prc_peer_disconnect first"; return PRC_ERR_INVALID; }
  DevFrame F;
  const bool stage = ctx->stage_single && fr && !(fr->flags & (PRC_FRAME_ASYNC | PRC_FRAME_UNIFORMS_RESIDENT)) && ctx->pending_async == 0;
  if (stage) {
    // (tuning) a synchronous frame: nothing of the previous call is in flight, so ONE page-locked staging area takes the uniforms
    // with a host memcpy and the DMAs run behind the first launches — each copy from pageable memory is staged by the driver
    // before the call returns, six times per frame
    size_t need = (size_t)fr->n_objects * sizeof(prc_object_xf) + (size_t)fr->n_lights * sizeof(DevLight) + sizeof(TileTargets) + (size_t)fr->n_ambient * 4 + 256 +
                  sizeof(DevFrame) + 256 * 8;
    for (uint32_t i = 0; i < fr->n_lights && fr->lights; i++)
      if (fr->lights[i].cast_shadow) need += (size_t)fr->n_objects * 64 + 256;
    need = (need + 4095) & ~(size_t)4095;
    if (ctx->stage_cap < need) {
      CK(cudaStreamSynchronize(ctx->stream));
      if (ctx->stage) cudaFreeHost(ctx->stage);
      ctx->stage = nullptr; ctx->stage_cap = 0;
      CK(cudaMallocHost((void**)&ctx->stage, need));
      ctx->stage_cap = need;
    }
    ctx->stage_on = true; ctx->stage_off = 0; ctx->stage_end = ctx->stage_cap;
  }
  int32_t r = build_frame(ctx, fr, F);
  ctx->stage_on = false;
  if (r != PRC_OK) return r;
  r = readback_begin(ctx, fr);
  if (r != PRC_OK) return r;
  // (A CUDA-graph replay of the ~35 stream operations of a frame was measured: 1.894 vs 1.886 ms — the gaps between
  // the kernels are device-side launch latency, not host enqueue time, so the frame is launched directly.)
  if (fr->flags & PRC_FRAME_ASYNC) {
    if (!(fr->flags & PRC_FRAME_NO_READBACK) || (fr->flags & PRC_FRAME_KEEP_GBUFFER)) { ctx->err = "PRC_FRAME_ASYNC needs PRC_FRAME_NO_READBACK and no KEEP_GBUFFER"; return PRC_ERR_INVALID; }
    r = enqueue_frame(ctx, fr, F);
    if (r != PRC_OK) return r;
    ctx->pending_async++;
    if (ctx->ev_used > 8192) return prc_sync(ctx);  // bound the timing-event pool
    return PRC_OK;
  }
  if (ctx->pending_async) {  // a synchronous frame behind asynchronous ones: finish those first
    r = prc_sync(ctx);
    if (r != PRC_OK) return r;
  }
  for (int attempt = 0; attempt < 4; attempt++) {
    r = enqueue_frame(ctx, fr, F);
    if (r != PRC_OK) return r;
    if (!(fr->flags & PRC_FRAME_NO_READBACK)) {
      r = readback_end(ctx, F, rgba_out);
      if (r != PRC_OK) return r;
    }
    r = finish_timings(ctx);
    if (r != PRC_RETRY) return r;  // shadow maps only grow (atomicMax), so re-running the frame is idempotent
  }
  ctx->err = "bin array kept overflowing";
  return PRC_ERR_UNSUPPORTED;
}

// A batch of views (BASELINE configs[4]): every view is a whole Render() with its own uniforms. The views are submitted back to
// back; view v's uniforms are staged in page-locked memory and uploaded behind view v-1's kernels, and while the GPU renders view v
// the host copies view v-1 out of the library's page-locked double buffer into the caller's memory.
int32_t prc_render_batch(prc_ctx* ctx, uint32_t n, const prc_frame* frames, uint8_t* const* rgba_out) {
  if (!ctx) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return PRC_OK;
  if (!frames) { ctx->err = "prc_render_batch: frames is NULL"; return PRC_ERR_INVALID; }
  if (ctx->peer_private) { ctx->err = "this context is connected to a peer group: prc_peer_disconnect first"; return PRC_ERR_INVALID; }
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }
  size_t half = 0;
  for (uint32_t v = 0; v < n; v++) {
    const prc_frame& fr = frames[v];
    if (fr.abi_version != PRC_ABI_VERSION) { ctx->err = "prc_frame: bad abi_version"; return PRC_ERR_INVALID; }
    if (fr.flags & (PRC_FRAME_ASYNC | PRC_FRAME_KEEP_GBUFFER | PRC_FRAME_UNIFORMS_RESIDENT)) {
      ctx->err = "prc_render_batch: PRC_FRAME_ASYNC, PRC_FRAME_KEEP_GBUFFER and PRC_FRAME_UNIFORMS_RESIDENT are not supported";
      return PRC_ERR_INVALID;
    }
    size_t need = (size_t)fr.n_objects * sizeof(prc_object_xf) + (size_t)fr.n_lights * sizeof(DevLight) + sizeof(TileTargets) + (size_t)fr.n_ambient * 4 + 256 +
                  sizeof(DevFrame) + 256 * 8;
    for (uint32_t i = 0; i < fr.n_lights && fr.lights; i++)
      if (fr.lights[i].cast_shadow) need += (size_t)fr.n_objects * 64 + 256;
    half = std::max(half, need);
  }
  half = (half + 4095) & ~(size_t)4095;
  if (ctx->stage_cap < 2 * half) {
    if (ctx->stage) cudaFreeHost(ctx->stage);
    ctx->stage = nullptr; ctx->stage_cap = 0;
    CK(cudaMallocHost((void**)&ctx->stage, 2 * half));
    ctx->stage_cap = 2 * half;
  }
  for (auto& e : ctx->ev_view)
    if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto abandon = [&](int32_t code) {  // a view failed to enqueue: finish what is in flight, drop the batch's timing state
    ctx->stage_on = false;
    cudaStreamSynchronize(ctx->stream);
    ctx->pending_async = 0;
    ctx->spans.clear();
    ctx->ev_used = 0;
    return code;
  };
  for (int attempt = 0; attempt < 4; attempt++) {
    uint8_t* src[2] = {nullptr, nullptr};
    size_t bytes[2] = {0, 0};
    for (uint32_t v = 0; v < n; v++) {
      const prc_frame* fr = &frames[v];
      ctx->stage_on = true;
      ctx->stage_off = (v & 1u) * half;
      ctx->stage_end = ctx->stage_off + half;
      DevFrame F;
      int32_t r = build_frame(ctx, fr, F);
      ctx->stage_on = false;
      if (r != PRC_OK) return abandon(r);
      if ((r = readback_begin(ctx, fr)) != PRC_OK) return abandon(r);
      ctx->pending_async = v ? 1 : 0;  // (views after the first keep the batch's sticky overflow flag and statistics, see enqueue_frame)
      if (ctx->ev_used > 8192) { ctx->spans.clear(); ctx->ev_used = 0; }  // bound the timing-event pool: drop the oldest views' spans
      if ((r = enqueue_frame(ctx, fr, F)) != PRC_OK) return abandon(r);
      if (cudaEventRecord(ctx->ev_view[v & 1u], ctx->stream) != cudaSuccess) return abandon(PRC_ERR_CUDA);
      const uint32_t ms = fr->msaa > 1 ? fr->msaa : 1;
      src[v & 1u] = ctx->rb_dst;
      bytes[v & 1u] = (size_t)(fr->width / ms) * (fr->height / ms) * 4;
      if (v > 0) {
        const uint32_t p = (v - 1) & 1u;
        if (cudaEventSynchronize(ctx->ev_view[p]) != cudaSuccess) return abandon(PRC_ERR_CUDA);
        if (rgba_out && rgba_out[v - 1] && src[p]) memcpy(rgba_out[v - 1], src[p], bytes[p]);
      }
    }
    const uint32_t p = (n - 1) & 1u;
    if (cudaEventSynchronize(ctx->ev_view[p]) != cudaSuccess) return abandon(PRC_ERR_CUDA);
    ctx->pending_async = 0;
    const int32_t r = finish_timings(ctx);  // timings and statistics are sums over the views
    if (r == PRC_RETRY) continue;           // a queue overflowed in some view: every view again (idempotent, the maps only grow or are reset per view)
    if (r == PRC_OK && rgba_out && rgba_out[n - 1] && src[p]) memcpy(rgba_out[n - 1], src[p], bytes[p]);
    return r;
  }
  ctx->err = "bin array kept overflowing";
  return PRC_ERR_UNSUPPORTED;
}

int32_t prc_read_gbuffer(prc_ctx* ctx, prc_gbuffer_host* g) {
  if (!ctx || !g) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  if (!ctx->gbuffer_valid) { ctx->err = "no G-buffer (render a frame first)"; return PRC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->W * ctx->H;
  std::vector<unsigned long long> keys(n);
  std::vector<float4> a(n), b(n), c(n), d(n);
  CK(cudaMemcpy(keys.data(), ctx->d_keys.p, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(a.data(), ctx->d_ga.p, n * 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), ctx->d_gb.p, n * 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(c.data(), ctx->d_gc.p, n * 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(d.data(), ctx->d_gd.p, n * 16, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; i++) {
    const bool ok = keys[i] != 0;
    const uint32_t seq = 0xFFFFFFFFu - (uint32_t)keys[i];
    if (g->ok) g->ok[i] = ok;
    if (g->tri) g->tri[i] = ok ? (int32_t)(seq >> 3) : -1;
    if (g->sub) g->sub[i] = ok ? (int32_t)(seq & 7u) : 0;
    if (g->depth) g->depth[i] = ok ? a[i].x : 0.0f;
    if (g->uv) { g->uv[2 * i] = ok ? a[i].y : 0.0f; g->uv[2 * i + 1] = ok ? a[i].z : 0.0f; }
    if (g->dudv) { g->dudv[2 * i] = ok ? a[i].w : 0.0f; g->dudv[2 * i + 1] = ok ? b[i].x : 0.0f; }
    if (g->nor) { g->nor[3 * i] = ok ? b[i].y : 0.0f; g->nor[3 * i + 1] = ok ? b[i].z : 0.0f; g->nor[3 * i + 2] = ok ? b[i].w : 0.0f; }
    if (g->facenor) { g->facenor[3 * i] = ok ? c[i].x : 0.0f; g->facenor[3 * i + 1] = ok ? c[i].y : 0.0f; g->facenor[3 * i + 2] = ok ? c[i].z : 0.0f; }
    if (g->wpos) { g->wpos[3 * i] = ok ? c[i].w : 0.0f; g->wpos[3 * i + 1] = ok ? d[i].x : 0.0f; g->wpos[3 * i + 2] = ok ? d[i].y : 0.0f; }
    if (g->col) { uint32_t u; memcpy(&u, &d[i].z, 4); g->col[i] = ok ? u : 0u; }
    if (g->mat) { int32_t m; memcpy(&m, &d[i].w, 4); g->mat[i] = ok ? m : 0; }
  }
  return PRC_OK;
}

int32_t prc_read_shadowmap(prc_ctx* ctx, uint32_t light, float* out) {
  if (!ctx || !out) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  if (light >= ctx->shadow_ptr.size() || !ctx->shadow_ptr[light]) { ctx->err = "no such shadow map"; return PRC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(out, ctx->shadow_ptr[light], (size_t)ctx->W * ctx->H * 4, cudaMemcpyDeviceToHost));
  return PRC_OK;
}

int32_t prc_read_image(prc_ctx* ctx, uint8_t* rgba_out) {
  if (!ctx || !rgba_out) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->d_image.p || !ctx->W) { ctx->err = "no frame rendered yet"; return PRC_ERR_INVALID; }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(rgba_out, ctx->d_image.p, (size_t)ctx->W * ctx->H * 4, cudaMemcpyDeviceToHost));
  return PRC_OK;
}

int32_t prc_get_timings(prc_ctx* ctx, prc_timings* out) {
  if (!ctx || !out) return PRC_ERR_INVALID;
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }  // finish asynchronous frames first
  *out = ctx->tm;
  return PRC_OK;
}

int32_t prc_device_image(prc_ctx* ctx, uint64_t* dev_ptr, uint64_t* bytes, uint64_t* capacity) {
  if (!ctx || !ctx->d_image.p) return PRC_ERR_INVALID;
  *dev_ptr = (uint64_t)(uintptr_t)ctx->d_image.p;
  *bytes = (uint64_t)ctx->W * ctx->H * 4;
  *capacity = (uint64_t)ctx->d_image.cap;
  return PRC_OK;
}

int32_t prc_device_shadowmap(prc_ctx* ctx, uint32_t light, uint64_t* dev_ptr, uint64_t* bytes) {
  if (!ctx || light >= ctx->shadow_ptr.size() || !ctx->shadow_ptr[light]) return PRC_ERR_INVALID;
  *dev_ptr = (uint64_t)(uintptr_t)ctx->shadow_ptr[light];
  *bytes = (uint64_t)ctx->W * ctx->H * 4;
  return PRC_OK;
}

int32_t prc_device_shadow_all(prc_ctx* ctx, uint64_t* dev_ptr, uint64_t* bytes, uint64_t* capacity) {
  if (!ctx || !ctx->d_shadow_all.p) return PRC_ERR_INVALID;
  *dev_ptr = (uint64_t)(uintptr_t)ctx->d_shadow_all.p;
  *bytes = (uint64_t)ctx->n_cast_alloc * ctx->W * ctx->H * 4;
  *capacity = (uint64_t)ctx->d_shadow_all.cap;
  return PRC_OK;
}

int32_t prc_host_image(prc_ctx* ctx, uint64_t* host_ptr, uint64_t* bytes) {
  if (!ctx || !ctx->h_img[ctx->h_img_cur]) return PRC_ERR_INVALID;
  *host_ptr = (uint64_t)(uintptr_t)ctx->h_img[ctx->h_img_cur];
  *bytes = (uint64_t)(ctx->W / ctx->msaa) * (ctx->H / ctx->msaa) * 4;
  return PRC_OK;
}

int32_t prc_stream(prc_ctx* ctx, uint64_t* stream) {
  if (!ctx) return PRC_ERR_INVALID;
  *stream = (uint64_t)(uintptr_t)ctx->stream;
  return PRC_OK;
}

int32_t prc_sync(prc_ctx* ctx) {
  if (!ctx) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pending_async) {
    ctx->pending_async = 0;
    const int32_t r = finish_timings(ctx);  // synchronises; timings are sums over the asynchronous frames
    if (ctx->peers.world) {
      const int32_t pr = peer_check(ctx);
      if (pr != PRC_OK) return pr;
    }
    if (r == PRC_RETRY) { ctx->err = "frames submitted back to back must be submitted again (a queue overflowed and was grown, or the binned tile path / NaN mode was switched on)"; return PRC_ERR_RETRY; }
    return r;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return PRC_OK;
}

}  // extern "C"

// =====================================================================================================================
// Multi-GPU frames over NVLink peer memory (include/polyred_cuda.h "prc_render_peer"; kernels in prc_peer.cuh).
// The reference has no counterpart: the contract is "the same frame as one GPU, bit for bit".
// =====================================================================================================================
namespace {

// offset of a device pointer inside its allocation: cudaIpcGetMemHandle exports the whole allocation block and
// cudaIpcOpenMemHandle returns the block's base (small cudaMalloc buffers share blocks)
int32_t alloc_offset(prc_ctx* ctx, const void* p, uint64_t* off) {
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { ctx->err = "cuMemGetAddressRange not available"; return PRC_ERR_CUDA; }
  unsigned long long base = 0;
  size_t size = 0;
  if (((range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)p) != 0) { ctx->err = "cuMemGetAddressRange failed"; return PRC_ERR_CUDA; }
  *off = (uint64_t)(uintptr_t)p - base;
  return PRC_OK;
}

// Loads every kernel of this library into the current context NOW. CUDA 12 loads a kernel lazily at its first launch, and that load
// may synchronise the whole context. Ranks that share one device (several contexts of one process on one GPU: the single-GPU form
// of the group tests) share its primary context: rank B's first launch of some kernel then waits for everything running on the
// device, including rank A's one-warp wait kernel, which spins for a signal rank B has not enqueued yet — B is stuck in the load —
// until the wait gives up after PRC_PEER_TIMEOUT_NS (seen as PRC_ERR_PEER in tests/test_gpu_group.py, first on an 8-GPU box:
// whether it happens is a race between the two submit threads). Called by prc_peer_connect, i.e. before any wait kernel exists.
// Ranks on different devices (the real thing) were never affected: a load synchronises its own context only.
void preload_kernels() {
  static std::mutex mu;
  static std::vector<int> done;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return; }
  if (getenv("PRC_NO_PRELOAD")) return;  // (to reproduce the stall: tests/test_gpu_group.py, first test)
  std::lock_guard<std::mutex> lk(mu);
  for (int d : done)
    if (d == dev) return;
  done.push_back(dev);
  typedef int (*get_module_fn)(void**, void*);
  typedef int (*count_fn)(unsigned int*, void*);
  typedef int (*enum_fn)(void**, unsigned int, void*);
  typedef int (*load_fn)(void*);
  void *f_mod = nullptr, *f_cnt = nullptr, *f_enum = nullptr, *f_load = nullptr;
  cudaDriverEntryPointQueryResult q;
  bool ok = cudaGetDriverEntryPoint("cuFuncGetModule", &f_mod, cudaEnableDefault, &q) == cudaSuccess && f_mod && q == cudaDriverEntryPointSuccess;
  ok = ok && cudaGetDriverEntryPoint("cuModuleGetFunctionCount", &f_cnt, cudaEnableDefault, &q) == cudaSuccess && f_cnt && q == cudaDriverEntryPointSuccess;
  ok = ok && cudaGetDriverEntryPoint("cuModuleEnumerateFunctions", &f_enum, cudaEnableDefault, &q) == cudaSuccess && f_enum && q == cudaDriverEntryPointSuccess;
  ok = ok && cudaGetDriverEntryPoint("cuFuncLoad", &f_load, cudaEnableDefault, &q) == cudaSuccess && f_load && q == cudaDriverEntryPointSuccess;
  cudaFunction_t fn = nullptr;
  ok = ok && cudaGetFuncBySymbol(&fn, (const void*)k_peer_wait) == cudaSuccess && fn;
  void* mod = nullptr;
  unsigned int n = 0;
  ok = ok && ((get_module_fn)f_mod)(&mod, (void*)fn) == 0 && mod && ((count_fn)f_cnt)(&n, mod) == 0 && n > 0;
  if (ok) {
    std::vector<void*> fns(n, nullptr);
    if (((enum_fn)f_enum)(fns.data(), n, mod) == 0)
      for (void* f : fns)
        if (f) (void)((load_fn)f_load)(f);
  }
  (void)cudaGetLastError();
  if (getenv("PRC_DEBUG_PRELOAD")) fprintf(stderr, "[polyred_cuda] preload_kernels(device %d): %s, %u functions\n", dev, ok ? "loaded" : "driver entry points unavailable", n);
}

void peer_release(prc_ctx* ctx) {
  if (ctx->img_stream) cudaStreamSynchronize(ctx->img_stream);
  ctx->img_pending = false; ctx->late_free_pending = false;
  for (void* b : ctx->peer_opened) cudaIpcCloseMemHandle(b);
  ctx->peer_opened.clear();
  ctx->peers = PeerTable{};
  if (ctx->peer_private && ctx->d_counters.p) {  // back to one GPU: nothing marks rows any more
    cudaMemsetAsync((char*)ctx->d_counters.p + offsetof(Counters, dirty), 0, sizeof(unsigned char*), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  ctx->peer_private = false;
  ctx->uniforms_valid = false;
  for (auto& p : ctx->peer_image) p = nullptr;
  ctx->peer_shadow_self = ctx->peer_image_self = nullptr;
  ctx->peer_epoch = 0;
  memset(ctx->plane_rows, 0, sizeof(ctx->plane_rows));
}

int32_t peer_check(prc_ctx* ctx) {
  for (int k = 0; k < 4; k++) ctx->peer_wait_ms[k] = 0;
  for (auto& sp : ctx->peer_spans) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->evpool[sp.a], ctx->evpool[sp.b]) == cudaSuccess && sp.cls >= 0 && sp.cls < 4) ctx->peer_wait_ms[sp.cls] += ms;
  }
  ctx->peer_spans.clear();
  (void)cudaGetLastError();
  unsigned int n = 0;
  CK(cudaMemcpy(&n, ctx->d_peer_err.p, 4, cudaMemcpyDeviceToHost));
  if (n) {
    CK(cudaMemset(ctx->d_peer_err.p, 0, 4));
    ctx->err = "a device-side wait for a peer rank timed out (" + std::to_string(n) + " waits): the frames submitted since the last prc_sync are invalid";
    return PRC_ERR_PEER;
  }
  return PRC_OK;
}

inline void peer_wait(prc_ctx* ctx, uint32_t kind, uint32_t epoch, uint32_t mask) {
  const PeerTable& P = ctx->peers;
  if (!(mask & ~(1u << P.self))) return;
  size_t a = 0;
  if (ctx->peer_trace) {
    while (ctx->evpool.size() < ctx->ev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); ctx->evpool.push_back(e); }
    a = ctx->ev_used;
    ctx->ev_used += 2;
    cudaEventRecord(ctx->evpool[a], ctx->stream);
  }
  k_peer_wait<<<1, PRC_PEER_MAX, 0, ctx->stream>>>(P.signals[P.self], P.world, P.self, kind, epoch, mask, (unsigned int*)ctx->d_peer_err.p, ctx->peer_timeout_ns);
  ctx->launches++;
  if (ctx->peer_trace) {
    cudaEventRecord(ctx->evpool[a + 1], ctx->stream);
    ctx->peer_spans.push_back({(int)kind, a, a + 1});
  }
}

inline void peer_signal(prc_ctx* ctx, uint32_t kind, uint32_t epoch, uint32_t mask) {
  const PeerTable& P = ctx->peers;
  if (!(mask & ~(1u << P.self))) return;
  k_peer_signal<<<1, PRC_PEER_MAX, 0, ctx->stream>>>(P, kind, epoch, mask);
  ctx->launches++;
}

// signal, then wait, in one launch (k_peer_signal_wait); falls back to the single kernels when one half is empty
inline void peer_signal_wait(prc_ctx* ctx, uint32_t sig_kind, uint32_t sig_epoch, uint32_t sig_mask, uint32_t wait_kind, uint32_t wait_epoch, uint32_t wait_mask) {
  const PeerTable& P = ctx->peers;
  const uint32_t others = ~(1u << P.self);
  static const bool fuse = !(getenv("PRC_PEER_NO_FUSED_SIGNALS") != nullptr && atoi(getenv("PRC_PEER_NO_FUSED_SIGNALS")) != 0);
  if (!fuse || !(sig_mask & others) || !(wait_mask & others)) {
    peer_signal(ctx, sig_kind, sig_epoch, sig_mask);
    peer_wait(ctx, wait_kind, wait_epoch, wait_mask);
    return;
  }
  size_t a = 0;
  if (ctx->peer_trace) {
    while (ctx->evpool.size() < ctx->ev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); ctx->evpool.push_back(e); }
    a = ctx->ev_used;
    ctx->ev_used += 2;
    cudaEventRecord(ctx->evpool[a], ctx->stream);
  }
  k_peer_signal_wait<<<1, PRC_PEER_MAX, 0, ctx->stream>>>(P, sig_kind, sig_epoch, sig_mask, wait_kind, wait_epoch, wait_mask, (unsigned int*)ctx->d_peer_err.p, ctx->peer_timeout_ns);
  ctx->launches++;
  if (ctx->peer_trace) {
    cudaEventRecord(ctx->evpool[a + 1], ctx->stream);
    ctx->peer_spans.push_back({(int)wait_kind, a, a + 1});
  }
}

// One frame of the group on this rank, enqueued without waiting for anything (prc_peer.cuh has the partition):
//   clear the private keys and the NEXT frame's merged plane -> camera pass and fused shadow sweep over this rank's share of the
//   triangles (private buffers; no peer is involved, so this runs while slower peers still shade the previous frame) -> queued
//   records -> wait until nobody reads the previous frame's merged maps -> k_peer_push -> signal, wait for the peers' pushes ->
//   fused resolve + shading of the strip from the merged buffers -> signal, copy the strip to the image consumers.
// The merged key planes alternate with the frame parity: plane (e+1)&1 is cleared at the start of frame e, after this rank's shading
// of frame e-1 (stream order) and before any peer can push frame e+1 (a peer pushes e+1 only after it has seen this rank's
// SHADOW(e), which follows the clear in this stream).
template <bool E>
int32_t enqueue_peer_frame(prc_ctx* ctx, const prc_frame* fr, const DevFrame& F, const PushJob& needs, uint32_t image_mask) {
  const PeerTable& P = ctx->peers;
  const uint32_t all = P.world >= 32 ? 0xFFFFFFFFu : ((1u << P.world) - 1u), me = 1u << P.self;
  cudaStream_t st = ctx->stream;
  const bool shadows = (fr->flags & PRC_FRAME_SHADOWMAP) && ctx->n_cast_alloc > 0 && ctx->n_cast_alloc != 0xFFFFFFFFu;
  const size_t npx = (size_t)F.W * F.H;
  if ((size_t)(1u + (shadows ? ctx->n_cast_alloc : 0u)) * F.H * ((F.W + PRC_DIRTY_SEG - 1) / PRC_DIRTY_SEG) * PRC_DIRTY_STRIDE > ctx->d_dirty.cap) {
    ctx->err = "prc_render_peer: the frame has more shadow-casting lights than the group was exported for; export and connect again";
    return PRC_ERR_INVALID;
  }
  const uint32_t e = ++ctx->peer_epoch;
  ctx->pass_slot = 0;
  if (ctx->pending_async == 0) {
    ctx->launches = 0;
    ctx->spans.clear();
    ctx->peer_spans.clear();
    ctx->ev_used = 0;
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, offsetof(Counters, n_valid), st));
  } else {
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, offsetof(Counters, large_overflow), st));  // behind unfinished frames: the overflow flag and the statistics are sticky
  }
  // A rank that receives this frame's image tells the pushers its image buffer is free: whatever it held (the previous
  // frame's image, if it was a consumer then) has been read by everything the caller enqueued on this stream, or finished
  // on the host, before this call. Announcing "free up to e-1" per frame lets the consumer set change between frames.
  // (after a PRC_FRAME_IMAGE_AT_SYNC frame that this rank received, the side stream announces "free up to e-1" itself, once the
  // strips of e-1 have landed; announcing it here, earlier, would let a peer overwrite rows whose e-1 strip is still in flight)
  if ((image_mask & me) && !ctx->late_free_pending) peer_signal(ctx, PRC_SIG_IMAGE_FREE, e - 1, all);
  ctx->late_free_pending = false;
  CK(cudaEventRecord(ctx->ev[0], st));
  // merged planes of this frame (read by resolve / shading) and of the next one (cleared now)
  unsigned long long* mk = (unsigned long long*)ctx->d_mkeys.p;
  unsigned long long* plane_cur = mk + (size_t)(e & 1u) * npx;
  unsigned long long* plane_next = mk + (size_t)((e + 1u) & 1u) * npx;
  unsigned long long* first_cur = mk + (size_t)(2u + (e & 1u)) * npx;
  unsigned long long* first_next = mk + (size_t)(2u + ((e + 1u) & 1u)) * npx;
  {
    // The plane of frame e+1 last received keys in frame e-1, for the rows this rank resolved THEN (the strips move between frames:
    // re-balancing). Clearing this frame's rows instead left the keys of an older camera in rows the rank lost and later regained.
    int* pr = ctx->plane_rows[(e + 1u) & 1u];
    const int r0s[2] = {pr[0], 0}, r1s[2] = {pr[1], pr[2]};
    pr[0] = pr[1] = pr[2] = 0;
    int* pc = ctx->plane_rows[e & 1u];  // what this frame's plane receives
    pc[0] = F.rr0; pc[1] = F.rr1; pc[2] = needs.ax1[P.self];
    for (int k = 0; k < 2; k++) {
      if (r1s[k] <= r0s[k]) continue;
      CK(cudaMemsetAsync(plane_next + (size_t)r0s[k] * F.W, 0, (size_t)(r1s[k] - r0s[k]) * F.W * 8, st));
      if (ctx->nan_mode) CK(cudaMemsetAsync(first_next + (size_t)r0s[k] * F.W, 0xFF, (size_t)(r1s[k] - r0s[k]) * F.W * 8, st));
    }
    CK(cudaMemsetAsync(plane_next, 0, 8, st));  // pixel (0,0) is merged into every rank
    if (ctx->nan_mode) CK(cudaMemsetAsync(first_next, 0xFF, 8, st));
  }
  // strip readback into the caller's (shared) host image, band by band behind the shading kernels
  ctx->rb_dst = (!(fr->flags & PRC_FRAME_NO_READBACK) && ctx->ext_img) ? ctx->ext_img + ctx->ext_img_off : nullptr;
  // ---- raster passes: this rank's share of the triangles, every row, private targets
  DevFrame Fr = F;
  Fr.rr0 = 0; Fr.rr1 = F.H;
  ctx->part_rank = P.self; ctx->part_world = P.world; ctx->part_active = true;  // (a group of one runs the same code path)
  ctx->keys_shade = nullptr; ctx->first_shade = nullptr;
  ctx->skip_key_clear = true;  // empty since the last push
  // The camera pass and the shadow sweep are independent (different targets, one shared queue filled by atomics) and, at 1/N of
  // the triangles, each is only a few waves of CTAs long: they run on two streams so that one fills the other's tail.
  static const bool two_streams = !(getenv("PRC_PEER_ONE_STREAM") != nullptr && atoi(getenv("PRC_PEER_ONE_STREAM")) != 0);
  static const bool early_push = !(getenv("PRC_PEER_NO_EARLY_PUSH") != nullptr && atoi(getenv("PRC_PEER_NO_EARLY_PUSH")) != 0);
  static const bool full_push = getenv("PRC_PEER_FULL_PUSH") != nullptr && atoi(getenv("PRC_PEER_FULL_PUSH")) != 0;
  // k_peer_push (prc_peer.cuh) merges the flagged 256-pixel segments of this rank's private buffers into the peers
  PushJob J0 = needs;
  J0.W = F.W; J0.H = F.H;
  J0.n_sh = shadows ? ctx->n_cast_alloc : 0u;
  J0.dirty = (unsigned char*)ctx->d_dirty.p;
  J0.nseg = (F.W + PRC_DIRTY_SEG - 1) / PRC_DIRTY_SEG;
  J0.sh_mine = (float*)ctx->d_shadow_mine.p;
  J0.k_mine = (unsigned long long*)ctx->d_keys.p;
  J0.k_off = (unsigned long long)(plane_cur - mk);
  J0.f_mine = ctx->nan_mode ? (unsigned long long*)ctx->d_keys.p + npx : nullptr;
  J0.f_off = (unsigned long long)(first_cur - mk);
  J0.full = full_push ? 1 : 0;
  auto launch_push = [&](long long seg0, long long seg1) {
    PushJob J = J0;
    J.seg0 = seg0; J.seg1 = seg1;
    // one CTA per batch of PRC_PUSH_BATCH segment flags; a grid of at most 148 x 8 CTAs strides over the rest
    const unsigned int grid = (unsigned int)std::min<long long>(148 * 8, std::max<long long>(1, (seg1 - seg0 + PRC_PUSH_BATCH - 1) / PRC_PUSH_BATCH));
    k_peer_push<<<grid, 256, 0, ctx->stream>>>(P, J);
    ctx->launches++;
  };
  int32_t r = PRC_OK;
  // A frame that fails while it is being enqueued (a launch or allocation error) has already advanced this rank's epoch, and its
  // peers will wait for this frame's signals: publish them, so that the peers finish at once instead of spinning until the wait
  // times out. The frame is invalid on every rank; this rank's caller gets the error and drops the connection
  // (prc_group_render disconnects, PeerFrames.finish() votes).
  auto fail = [&](int32_t code) {
    ctx->stream = st;
    ctx->part_rank = 0; ctx->part_world = 1; ctx->part_active = false;
    ctx->skip_key_clear = false; ctx->defer_copy_join = false;
    ctx->keys_shade = nullptr; ctx->first_shade = nullptr;
    (void)cudaGetLastError();
    peer_signal(ctx, PRC_SIG_SHADOW, e, all);
    if (shadows) peer_signal(ctx, PRC_SIG_SHADED, e, all);
    peer_signal(ctx, PRC_SIG_IMAGE, e, image_mask & all);
    (void)cudaGetLastError();
    return code;
  };
  if (shadows && two_streams) {
    CK(cudaEventRecord(ctx->ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->copy_stream2, ctx->ev_fork, 0));
    ctx->stream = ctx->copy_stream2;
    r = do_shadows<E>(ctx, fr, Fr, units_from_mask(fr, 0xFFFFFFFFu, 0, F.H), false);
    cudaEventRecord(ctx->ev_join, ctx->copy_stream2);
    ctx->stream = st;
  }
  if (r == PRC_OK) r = do_main<E>(ctx, fr, Fr, 1, 1);
  ctx->skip_key_clear = false;
  if (r == PRC_OK && early_push && P.world > 1) {
    // The visibility keys of the camera pass leave NOW, under the shadow sweep that is still running on the other stream: the merge
    // of ~0.6 M keys per rank is 0.04 ms of reductions over NVLink, whatever the number of ranks, and the camera pass is the shorter
    // of the two. No wait is needed first: the peers cleared this frame's key plane before they signalled SHADOW(e-1), which this
    // rank awaited before it shaded frame e-1. What the queued records add to the key plane later goes with the second push.
    KTimer kt(ctx, PRC_K_EXCHANGE);
    launch_push(0, (long long)J0.H * J0.nseg);
  }
  CK(cudaEventRecord(ctx->ev[1], st));
  if (shadows && two_streams) CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
  else if (r == PRC_OK && shadows) r = do_shadows<E>(ctx, fr, Fr, units_from_mask(fr, 0xFFFFFFFFu, 0, F.H), false);
  if (r == PRC_OK) r = do_main<E>(ctx, fr, Fr, 4, 1);  // queued records of the camera AND the shadow passes
  ctx->part_rank = 0; ctx->part_world = 1; ctx->part_active = false;
  if (r != PRC_OK) return fail(r);
  // ---- exchange: what the queued records added to the key plane, and the shadow planes
  if (shadows) peer_wait(ctx, PRC_SIG_SHADED, e - 1, all);  // nobody may still be shading the previous frame from the maps about to be merged into
  {
    KTimer kt(ctx, PRC_K_EXCHANGE);
    launch_push(0, (long long)(1u + J0.n_sh) * J0.H * J0.nseg);
  }
  peer_signal_wait(ctx, PRC_SIG_SHADOW, e, all, PRC_SIG_SHADOW, e, all);
  // ---- resolve + shading of the strip from the merged buffers
  // (a previous PRC_FRAME_IMAGE_AT_SYNC frame: its strips must have landed before this frame's shading writes the image — normally long
  // satisfied; it only matters when the strips moved between the two frames)
  if (ctx->img_pending) CK(cudaStreamWaitEvent(st, ctx->ev_image_done, 0));
  ctx->keys_shade = plane_cur; ctx->first_shade = first_cur;
  ctx->defer_copy_join = true;
  r = do_main<E>(ctx, fr, F, 8 | 2, 1);
  ctx->defer_copy_join = false;
  ctx->keys_shade = nullptr; ctx->first_shade = nullptr;
  if (r != PRC_OK) return fail(r);
  // image strip: screen rows [row0,row1) = image rows [H-row1, H-row0)
  const size_t off = (size_t)(F.H - F.row1) * F.W * 4, bytes = (size_t)(F.row1 - F.row0) * F.W * 4;
  const uint32_t consumers = image_mask & all;
  if (!(consumers & ~me)) {
    if (shadows) peer_signal(ctx, PRC_SIG_SHADED, e, all);
  } else {
    // "my shading is done" and "wait until the consumers are done with the previous frame's image" in one launch
    peer_signal_wait(ctx, PRC_SIG_SHADED, e, shadows ? all : 0u, PRC_SIG_IMAGE_FREE, e - 1, consumers);
    for (uint32_t c = 0; c < P.world; c++)
      if (((consumers >> c) & 1u) && c != P.self)
        CK(cudaMemcpyAsync(ctx->peer_image[c] + off, (const uint8_t*)ctx->d_image.p + off, bytes, cudaMemcpyDefault, st));
    peer_signal(ctx, PRC_SIG_IMAGE, e, consumers);
  }
  if (consumers & me) {
    if ((fr->flags & PRC_FRAME_IMAGE_AT_SYNC) && (all & ~me)) {
      // beside the stream: own strip shaded (and copied to the other consumers) -> wait for the peers' strips -> "image free up to e"
      // -> ev_image_done. This rank's next frame does not wait for any of it; prc_sync does.
      CK(cudaEventRecord(ctx->ev_own_shaded, st));
      CK(cudaStreamWaitEvent(ctx->img_stream, ctx->ev_own_shaded, 0));
      ctx->stream = ctx->img_stream;
      peer_wait(ctx, PRC_SIG_IMAGE, e, all);
      peer_signal(ctx, PRC_SIG_IMAGE_FREE, e, all);
      ctx->stream = st;
      CK(cudaEventRecord(ctx->ev_image_done, ctx->img_stream));
      ctx->img_pending = true;
      ctx->late_free_pending = true;
    } else {
      peer_wait(ctx, PRC_SIG_IMAGE, e, all);  // every strip has landed in this rank's image
    }
  }
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaGetLastError());
  return PRC_OK;
}

}  // namespace

extern "C" {

int32_t prc_peer_export(prc_ctx* ctx, const prc_frame* fr, prc_peer_handle* out) {
  if (!ctx || !out) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }
  peer_release(ctx);
  DevFrame F;
  ctx->allow_msaa_strips = true;
  int32_t r = build_frame(ctx, fr, F);  // allocates (and, for a new size / light set, zeroes) the shadow and image buffers
  ctx->allow_msaa_strips = false;
  if (r != PRC_OK) return r;
  if (ctx->msaa > 1) {
    // the downsample tables are built here: building them inside the first frame would wait on the host for a stream that
    // may itself be waiting for a peer
    r = ensure_resize_tables(ctx, (int)fr->width, (int)fr->height, (int)fr->width / ctx->msaa, (int)fr->height / ctx->msaa);
    if (r != PRC_OK) return r;
    ENSURE(ctx->d_image_out, (size_t)(fr->width / ctx->msaa) * (fr->height / ctx->msaa) * 4);
  }
  // private shadow maps (raster target of this rank) and the merged key planes: 2 frame parities + 2 NaN-mode planes
  ENSURE(ctx->d_shadow_mine, ctx->d_shadow_all.cap);
  CK(cudaMemsetAsync(ctx->d_shadow_mine.p, 0, ctx->d_shadow_mine.cap, ctx->stream));
  {
    const size_t npx_ = (size_t)fr->width * fr->height;
    ENSURE(ctx->d_mkeys, ipc_round(npx_ * 8 * 4));
    CK(cudaMemsetAsync(ctx->d_mkeys.p, 0, npx_ * 8 * 2, ctx->stream));
    CK(cudaMemsetAsync((unsigned long long*)ctx->d_mkeys.p + npx_ * 2, 0xFF, npx_ * 8 * 2, ctx->stream));
    ENSURE(ctx->d_keys, npx_ * 16);  // room for the private NaN-mode plane, so that entering NaN mode never reallocates
    // the private planes start empty and are emptied again by every push (clear-on-read): keys 0, first-fragment plane ~0
    CK(cudaMemsetAsync(ctx->d_keys.p, 0, npx_ * 8, ctx->stream));
    CK(cudaMemsetAsync((unsigned long long*)ctx->d_keys.p + npx_, 0xFF, npx_ * 8, ctx->stream));
    // segment flags of the private buffers: [1 + casting lights][H][ceil(W / 256)] flags, one per 32-byte sector
    const size_t dirty_bytes = (size_t)(1u + (ctx->n_cast_alloc == 0xFFFFFFFFu ? 0u : ctx->n_cast_alloc)) * fr->height * ((fr->width + PRC_DIRTY_SEG - 1) / PRC_DIRTY_SEG) * PRC_DIRTY_STRIDE;
    ENSURE(ctx->d_dirty, dirty_bytes);
    CK(cudaMemsetAsync(ctx->d_dirty.p, 0, ctx->d_dirty.cap, ctx->stream));
    struct { unsigned char* p; int h; } dd = {(unsigned char*)ctx->d_dirty.p, (int)fr->height};
    CK(cudaMemcpyAsync((char*)ctx->d_counters.p + offsetof(Counters, dirty), &dd.p, sizeof(dd.p), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync((char*)ctx->d_counters.p + offsetof(Counters, dirty_h), &dd.h, sizeof(dd.h), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // dd is a stack object
  }
  ENSURE(ctx->d_peer_signals, ipc_round((size_t)PRC_SIG_KINDS * PRC_PEER_MAX * 4));  // its own 2 MiB block (see build_frame)
  ENSURE(ctx->d_peer_err, 16);
  CK(cudaMemsetAsync(ctx->d_peer_signals.p, 0, (size_t)PRC_SIG_KINDS * PRC_PEER_MAX * 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_peer_err.p, 0, 16, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memset(out, 0, sizeof(*out));
  out->abi_version = PRC_ABI_VERSION;
  out->device = (uint32_t)ctx->device;
  out->pid = (uint64_t)getpid();
  out->shadow_ptr = (uint64_t)(uintptr_t)ctx->d_shadow_all.p;
  out->image_ptr = (uint64_t)(uintptr_t)ctx->d_image.p;
  out->signals_ptr = (uint64_t)(uintptr_t)ctx->d_peer_signals.p;
  out->mkeys_ptr = (uint64_t)(uintptr_t)ctx->d_mkeys.p;
  out->shadow_bytes = ctx->d_shadow_all.cap;
  out->image_bytes = ctx->d_image.cap;
  out->mkeys_bytes = ctx->d_mkeys.cap;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "prc_peer_handle carries 64-byte IPC handles");
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->shadow_ipc, ctx->d_shadow_all.p));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->image_ipc, ctx->d_image.p));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->signals_ipc, ctx->d_peer_signals.p));
  CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out->mkeys_ipc, ctx->d_mkeys.p));
  if ((r = alloc_offset(ctx, ctx->d_mkeys.p, &out->mkeys_off)) != PRC_OK) return r;
  if ((r = alloc_offset(ctx, ctx->d_shadow_all.p, &out->shadow_off)) != PRC_OK) return r;
  if ((r = alloc_offset(ctx, ctx->d_image.p, &out->image_off)) != PRC_OK) return r;
  if ((r = alloc_offset(ctx, ctx->d_peer_signals.p, &out->signals_off)) != PRC_OK) return r;
  ctx->peer_shadow_self = ctx->d_shadow_all.p;
  ctx->peer_image_self = ctx->d_image.p;
  return PRC_OK;
}

int32_t prc_peer_connect(prc_ctx* ctx, uint32_t rank, uint32_t world, const prc_peer_handle* all) {
  if (!ctx || !all) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (world == 0 || world > PRC_PEER_MAX || rank >= world) { ctx->err = "prc_peer_connect: bad rank / world (at most 16 ranks)"; return PRC_ERR_INVALID; }
  if (!ctx->peer_shadow_self || ctx->peer_shadow_self != ctx->d_shadow_all.p || ctx->peer_image_self != ctx->d_image.p) {
    ctx->err = "prc_peer_connect: call prc_peer_export first";
    return PRC_ERR_INVALID;
  }
  preload_kernels();
  const prc_peer_handle& mine = all[rank];
  if (mine.pid != (uint64_t)getpid() || mine.shadow_ptr != (uint64_t)(uintptr_t)ctx->d_shadow_all.p) { ctx->err = "prc_peer_connect: all[rank] is not this context's handle"; return PRC_ERR_INVALID; }
  PeerTable T{};
  T.world = world;
  T.self = rank;
  uint8_t* image[PRC_PEER_MAX] = {};
  std::vector<void*> opened;
  auto fail = [&](const std::string& msg) {
    for (void* b : opened) cudaIpcCloseMemHandle(b);
    ctx->err = msg;
    return PRC_ERR_CUDA;
  };
  for (uint32_t p = 0; p < world; p++) {
    const prc_peer_handle& h = all[p];
    if (h.abi_version != PRC_ABI_VERSION) return fail("prc_peer_connect: bad abi_version in a peer handle");
    if (h.shadow_bytes != mine.shadow_bytes || h.image_bytes != mine.image_bytes || h.mkeys_bytes != mine.mkeys_bytes) return fail("prc_peer_connect: the ranks' frame buffers differ in size");
    if (p == rank || h.pid == mine.pid) {
      // same process (this rank, or another context of a single-process test harness): the pointers are valid as they are
      if (p != rank && (int)h.device != ctx->device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess((int)h.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        (void)cudaGetLastError();
      }
      T.shadow[p] = (float*)(uintptr_t)h.shadow_ptr;
      image[p] = (uint8_t*)(uintptr_t)h.image_ptr;
      T.signals[p] = (uint32_t*)(uintptr_t)h.signals_ptr;
      T.mkeys[p] = (unsigned long long*)(uintptr_t)h.mkeys_ptr;
      continue;
    }
    // another process: map its allocations (identical handles = buffers sharing one allocation block are mapped once)
    const uint8_t* ipc[4] = {h.shadow_ipc, h.image_ipc, h.signals_ipc, h.mkeys_ipc};
    void* base[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; k++) {
      for (int j = 0; j < k; j++)
        if (memcmp(ipc[k], ipc[j], 64) == 0) base[k] = base[j];
      if (base[k]) continue;
      cudaIpcMemHandle_t mh;
      memcpy(&mh, ipc[k], 64);
      const cudaError_t e = cudaIpcOpenMemHandle(&base[k], mh, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return fail(std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(p) + "): " + cudaGetErrorString(e));
      opened.push_back(base[k]);
    }
    T.shadow[p] = (float*)((uint8_t*)base[0] + h.shadow_off);
    image[p] = (uint8_t*)base[1] + h.image_off;
    T.signals[p] = (uint32_t*)((uint8_t*)base[2] + h.signals_off);
    T.mkeys[p] = (unsigned long long*)((uint8_t*)base[3] + h.mkeys_off);
  }
  ctx->peers = T;
  ctx->peer_private = true;
  ctx->uniforms_valid = false;  // the raster targets changed: the next frame uploads its tables again
  ctx->peer_trace = getenv("PRC_PEER_TRACE") != nullptr && atoi(getenv("PRC_PEER_TRACE")) != 0;
  ctx->peer_spans.clear();
  for (uint32_t p = 0; p < PRC_PEER_MAX; p++) ctx->peer_image[p] = image[p];
  ctx->peer_opened = opened;
  ctx->peer_epoch = 0;
  memset(ctx->plane_rows, 0, sizeof(ctx->plane_rows));  // prc_peer_export zeroed both planes
  return PRC_OK;
}

int32_t prc_set_host_image(prc_ctx* ctx, void* ptr, uint64_t bytes) {
  if (!ctx) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->ext_img) {
    if (ctx->ext_img_registered) cudaHostUnregister(ctx->ext_img);
    ctx->ext_img = nullptr;
    ctx->ext_img_bytes = 0;
    ctx->ext_img_off = 0;
  }
  if (!ptr) return PRC_OK;
  {
    // several contexts of ONE process may share the image (single-process tests): the first one registers it
    const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered) { ctx->err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return PRC_ERR_CUDA; }
    (void)cudaGetLastError();
    ctx->ext_img_registered = e == cudaSuccess;
  }
  ctx->ext_img = (uint8_t*)ptr;
  ctx->ext_img_bytes = (size_t)bytes;
  ctx->ext_img_off = 0;
  return PRC_OK;
}

int32_t prc_set_host_image_offset(prc_ctx* ctx, uint64_t offset) {
  if (!ctx) return PRC_ERR_INVALID;
  if (!ctx->ext_img || offset >= ctx->ext_img_bytes) { ctx->err = "prc_set_host_image_offset: no host image registered, or the offset lies outside it"; return PRC_ERR_INVALID; }
  ctx->ext_img_off = (size_t)offset;  // host-side state read when the next frame is enqueued: no wait needed
  return PRC_OK;
}

int32_t prc_peer_disconnect(prc_ctx* ctx) {
  if (!ctx) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int32_t r = PRC_OK;
  if (ctx->pending_async) r = prc_sync(ctx);
  else CK(cudaStreamSynchronize(ctx->stream));
  peer_release(ctx);
  return r;
}

int32_t prc_render_peer(prc_ctx* ctx, const prc_frame* fr, uint32_t n_ranks, const uint32_t* row0, const uint32_t* row1, uint32_t image_mask) {
  if (!ctx) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->peers.world) { ctx->err = "prc_render_peer: not connected (prc_peer_export / prc_peer_connect)"; return PRC_ERR_INVALID; }
  if (!fr || fr->abi_version != PRC_ABI_VERSION) { ctx->err = "prc_frame: bad abi_version"; return PRC_ERR_INVALID; }
  if (fr->flags & (PRC_FRAME_KEEP_GBUFFER | PRC_FRAME_SHADOW_RESET)) {
    ctx->err = "prc_render_peer: PRC_FRAME_KEEP_GBUFFER and PRC_FRAME_SHADOW_RESET are not supported";
    return PRC_ERR_UNSUPPORTED;
  }
  const uint32_t ms = fr->msaa > 1 ? fr->msaa : 1;
  if (ms > 1 && image_mask != 0) {
    ctx->err = "prc_render_peer: MSAA frames have no device-side gather (image_mask must be 0): the downsampled strips leave through prc_set_host_image";
    return PRC_ERR_UNSUPPORTED;
  }
  if (n_ranks != ctx->peers.world || !row0 || !row1) { ctx->err = "prc_render_peer: row0/row1 must list the strip of every rank of the group"; return PRC_ERR_INVALID; }
  if (row0[ctx->peers.self] != fr->row0 || row1[ctx->peers.self] != fr->row1) { ctx->err = "prc_render_peer: frame.row0/row1 differ from this rank's entry of row0/row1"; return PRC_ERR_INVALID; }
  if (!(fr->flags & PRC_FRAME_NO_READBACK) && ctx->ext_img && ctx->ext_img_bytes < ctx->ext_img_off + (size_t)(fr->width / ms) * (fr->height / ms) * 4) {
    ctx->err = "prc_render_peer: the host image registered with prc_set_host_image is smaller than the frame";
    return PRC_ERR_INVALID;
  }
  DevFrame F;
  ctx->allow_msaa_strips = true;
  int32_t r = build_frame(ctx, fr, F);
  ctx->allow_msaa_strips = false;
  if (r != PRC_OK) return r;
  if (ctx->d_shadow_all.p != ctx->peer_shadow_self || ctx->d_image.p != ctx->peer_image_self) {
    // the frame size or the set of casting lights changed: the peers still map the old buffers
    peer_release(ctx);
    ctx->err = "prc_render_peer: the exported buffers were reallocated; export and connect again on every rank";
    return PRC_ERR_INVALID;
  }
  // which rows every rank resolves (all ranks derive the same table from the same strips and the same scene)
  PushJob needs{};
  for (uint32_t p = 0; p < n_ranks; p++) {
    if (row1[p] > fr->height || row0[p] >= row1[p]) { ctx->err = "prc_render_peer: bad strip (every rank needs at least one row)"; return PRC_ERR_INVALID; }
    const bool strip = ms > 1 && (row0[p] != 0 || row1[p] != fr->height);
    int s0, s1;
    row_needs(ctx->any_ao, (int)ms, strip, (int)fr->height, (int)row0[p], (int)row1[p], s0, s1, needs.rr0[p], needs.rr1[p], needs.ax1[p]);
  }
  r = ctx->exact ? enqueue_peer_frame<true>(ctx, fr, F, needs, image_mask) : enqueue_peer_frame<false>(ctx, fr, F, needs, image_mask);
  if (r != PRC_OK) return r;
  ctx->pending_async++;
  if (ctx->ev_used > 8192) {
    // bound the timing-event pool without a host wait: keep the events, drop the spans of the oldest frames
    ctx->spans.clear();
    ctx->peer_spans.clear();
    ctx->ev_used = 0;
  }
  return PRC_OK;
}

int32_t prc_frame_state(prc_ctx* ctx, uint32_t* state) {
  if (!ctx || !state) return PRC_ERR_INVALID;
  *state = (ctx->nan_mode ? PRC_STATE_NAN_MODE : 0u) | (ctx->need_bins ? PRC_STATE_TILE_PATH : 0u);
  return PRC_OK;
}

int32_t prc_set_frame_state(prc_ctx* ctx, uint32_t state) {
  if (!ctx) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK && pr_ != PRC_ERR_RETRY) return pr_; }
  if ((state & PRC_STATE_NAN_MODE) && !ctx->nan_mode) {
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nan_mode = true;
    if (ctx->W && ctx->H) ENSURE(ctx->d_keys, (size_t)ctx->W * ctx->H * 16);
  }
  if (state & PRC_STATE_TILE_PATH) ctx->need_bins = true;
  return PRC_OK;
}

int32_t prc_count_covered(prc_ctx* ctx, uint64_t* covered) {
  if (!ctx || !covered) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }
  if (!ctx->d_keys.p || !ctx->W) { ctx->err = "no frame rendered yet"; return PRC_ERR_INVALID; }
  ENSURE(ctx->d_special, 16);
  unsigned long long* out = (unsigned long long*)ctx->d_special.p + 1;  // (the second half of the 16-byte `special` block is free between frames)
  CK(cudaMemsetAsync(out, 0, 8, ctx->stream));
  k_count_covered<<<148 * 8, 256, 0, ctx->stream>>>((const unsigned long long*)ctx->d_keys.p, (size_t)ctx->W * ctx->H, out);
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, out, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *covered = v;
  return PRC_OK;
}

int32_t prc_measure_fp32_peak(prc_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return PRC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pending_async) { const int32_t pr_ = prc_sync(ctx); if (pr_ != PRC_OK) return pr_; }
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  ENSURE(ctx->d_special, 16);
  const int iters = 4096, grid = 148 * 8 * 4;
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {  // first repetitions warm the clocks up
    CK(cudaEventRecord(a, ctx->stream));
    k_fma_peak<<<grid, 256, 0, ctx->stream>>>((float*)ctx->d_special.p, iters, 0.999f, 1e-3f);
    CK(cudaEventRecord(b, ctx->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double tf = 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *tflops = best;
  return PRC_OK;
}

int32_t prc_peer_wait_ms(prc_ctx* ctx, float out[4]) {
  if (!ctx || !out) return PRC_ERR_INVALID;
  for (int k = 0; k < 4; k++) out[k] = ctx->peer_wait_ms[k];
  return PRC_OK;
}

}  // extern "C"
