/*
 * polyred_cuda.h — C ABI of libpolyred_cuda.so, the B200 (sm_100a) backend for
 * polyred's render pass `render.NewRenderer(...).Render()`.
 *
 * Every citation below is a path:line inside the reference tree (poly.red).
 *
 * The ABI is deliberately purego-friendly (reference FFI style:
 * gpu/backend_gl_lib_linux.go:19-26, gpu/backend_gl.go:236-283): every export
 * takes and returns integers / pointers only (purego.SyscallN passes uintptr),
 * floats travel inside fixed-layout little-endian POD structs, and every struct
 * starts with `abi_version`.
 *
 * What each call replaces in the reference:
 *   prc_open / prc_close      gpu.Open / Device.Close (gpu/device.go:77-89) as used by
 *                             render.GPU(dev) (render/options.go:103-110)
 *   prc_scene_upload          the per-frame walk of Geometry.Triangles()/Materials()
 *                             (render/raster.go:241-270) — done once, scene stays in HBM
 *   prc_shadow_reset          the zeroing of shadow depth maps in initShadowMaps()
 *                             (render/shadow.go:33-90) reached from NewRenderer/Options
 *   prc_render                (*Renderer).Render() (render/raster.go:155-199): passShadows,
 *                             passForward (cpuForwardPass/draw/drawClipped), passDeferred
 *                             (shade, FragmentShader, shadingVisibility, AO), passAntialiasing
 *                             (GammaCorrection) — one whole-frame call instead of the three
 *                             runPass seams (render/raster.go:67-76, 223, 300, 370)
 *   prc_read_gbuffer          buffer.FragmentBuffer.Get (buffer/buffer.go:209-219), parity only
 *   prc_read_shadowmap        shadowInfo.depths (render/shadow.go:26-31), parity only
 *   prc_get_timings           profiling.Timed under render.Debug (render/raster.go:156-161)
 *
 * Error convention: every export returns int32 (0 = PRC_OK, <0 = error). There is NO CPU
 * fallback (the reference falls back per pass, render/raster.go:68-75; north_star forbids it):
 * the caller must surface the error. prc_last_error(ctx) returns a ctx-owned string valid
 * until the next call on that ctx. Thread-compatibility: one call at a time per ctx, from
 * any OS thread (each export does cudaSetDevice itself).
 */
#ifndef POLYRED_CUDA_H
#define POLYRED_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRC_ABI_VERSION 1u

/* error codes */
#define PRC_OK 0
#define PRC_ERR_INVALID -1      /* bad argument / abi_version mismatch */
#define PRC_ERR_CUDA -2         /* CUDA runtime error, see prc_last_error */
#define PRC_ERR_UNSUPPORTED -3  /* scene/option the path does not implement (no fallback) */
#define PRC_ERR_NO_SCENE -4
#define PRC_ERR_NCCL -5          /* reserved: collectives are driven by the host (torch.distributed), the library itself never calls NCCL */
#define PRC_ERR_RETRY -6        /* prc_sync after PRC_FRAME_ASYNC frames: a queue overflowed, it has been grown; submit those frames again */
#define PRC_ERR_PEER -7         /* prc_sync after prc_render_peer frames: a wait for a peer rank timed out (the frames are invalid) */

/* prc_material.flags (material.Standard, material/material.go:23-30) */
#define PRC_MAT_FLAT_SHADING 1u
#define PRC_MAT_AMBIENT_OCCLUSION 2u
#define PRC_MAT_RECEIVE_SHADOW 4u
#define PRC_MAT_NIL 8u          /* table slot holds a non-BlinnPhong material (raster.go:254): use vertex colour */
#define PRC_MAT_NO_MIPMAP 16u   /* Texture.useMipmap == false (buffer/texture.go:104-106) */

/* prc_light.kind (shader/blinn_cpu.go:72-80) */
#define PRC_LIGHT_POINT 0u
#define PRC_LIGHT_DIRECTIONAL 1u

/* prc_frame.flags */
#define PRC_FRAME_PERSPECT 1u     /* option.Perspect (render/options.go:49-56) */
#define PRC_FRAME_SHADOWMAP 2u    /* render.ShadowMap(true) */
#define PRC_FRAME_GAMMA 4u        /* render.GammaCorrection(true) */
#define PRC_FRAME_KEEP_GBUFFER 8u /* keep the G-buffer for prc_read_gbuffer (parity/debug) */
#define PRC_FRAME_NO_READBACK 16u /* leave the RGBA on the device (rgba_out may be NULL) */
#define PRC_FRAME_UNIFORMS_RESIDENT 32u /* the arrays behind objects/lights/shadow_trans/ambient/gamma are unchanged since
                                           the previous call on this ctx: skip their host->device copies (split-phase calls) */
#define PRC_FRAME_BGRA 128u        /* render.PixelFormat(buffer.PixelFormatBGRA): colour bytes stored B,G,R,A (buffer/buffer.go:242-263) */
#define PRC_FRAME_ASYNC 256u       /* prc_render only, with PRC_FRAME_NO_READBACK: enqueue the frame and return without waiting. Completion,
                                     * timings (summed over the frames) and the queue-overflow check happen in prc_sync(); for callers that
                                     * keep the frames on the device (present, multi-view batches) this removes the host bubble between frames */
#define PRC_FRAME_NO_KERNEL_TIMERS 512u /* no per-kernel-class CUDA-event brackets for this frame (prc_timings.kernel_ms stays 0): an event
                                        * pair costs ~2.6 us of stream time, which matters for the 0.2 ms frames of an 8-GPU group */
#define PRC_FRAME_IMAGE_AT_SYNC 1024u /* prc_render_peer, on a rank that receives the gathered image (image_mask): do not wait INSIDE the stream for
                                       * the peers' strips. The wait (and the "image buffer free" announcement that follows it) runs beside the
                                       * stream, so this rank's next frame starts while the strips of this one still arrive over NVLink (29 MB
                                       * into rank 0 at 8 GPUs = 0.04 ms of a 0.28 ms frame). The gathered image is then complete after
                                       * prc_sync — not for work enqueued on prc_stream() — and stays valid until the next prc_render_peer */
#define PRC_FRAME_SHADOW_RESET 64u /* zero the shadow maps at the start of this frame, stream-ordered: what Options()
                                     * does between views (render/options.go:125-141), without prc_shadow_reset's host sync */

/* Colours are packed R | G<<8 | B<<16 | A<<24 (image/color.RGBA byte order in memory). */

/* material.BlinnPhong (material/material.go:44-52) as tabulated into Renderer.matTable
 * (render/raster.go:252-256). `texture` indexes prc_scene textures; <0 is invalid for a
 * non-NIL material (the reference would nil-deref). */
typedef struct prc_material {
  uint32_t diffuse_rgba;
  uint32_t specular_rgba;
  float shininess;
  int32_t texture;
  uint32_t flags;
  uint32_t _pad;
} prc_material;

/* buffer.Texture (buffer/texture.go:28-69): a finished RGBA8 mip chain. Level l of texture t
 * is level index tex_first_level[t] + l; its texels start at tex_data + level_offset[idx]
 * (bytes), row-major, stride 4*level_w[idx]. The chain is an INPUT (the host builds it with
 * the reference's own imageutil.Resize). */
typedef struct prc_scene {
  uint32_t abi_version;
  uint32_t flags;
  uint64_t n_tris;
  /* Triangle soup in draw order = scene.IterObjects order x Geometry.Triangles() order
   * (render/raster.go:241-270). primitive.Vertex (geometry/primitive/vertex.go:15-21):
   * Pos.W is taken as 1 and Nor.W as 0 (what every reference loader produces,
   * model/load.go:162-170). */
  const float* pos;       /* [n_tris][3 verts][xyz]  */
  const float* nor;       /* [n_tris][3 verts][xyz]  */
  const float* uv;        /* [n_tris][3 verts][uv]   */
  const uint32_t* col;    /* [n_tris][3 verts] RGBA8 */
  const int32_t* mat;     /* [n_tris] FLAT material id = base + local (raster.go:259-262); <0: vertex colour */
  uint32_t n_objects;
  uint32_t n_materials;
  const uint64_t* obj_tri_start; /* [n_objects+1] triangle range of each Geometry */
  const prc_material* materials; /* [n_materials] the flat matTable */
  uint32_t n_textures;
  uint32_t n_tex_levels;
  const uint32_t* tex_first_level; /* [n_textures+1] */
  const uint32_t* level_w;         /* [n_tex_levels] */
  const uint32_t* level_h;         /* [n_tex_levels] */
  const uint64_t* level_offset;    /* [n_tex_levels] byte offset into tex_data */
  const uint8_t* tex_data;
  uint64_t tex_bytes;
} prc_scene;

/* Per-Geometry uniforms, computed by the host exactly as cpuForwardPass does
 * (render/raster.go:242-243, 382): trans = Proj.MulM(View).MulM(Model),
 * normal = Model.Inv().T(). Row-major X00..X33 (math/mat4.go:36-43). */
typedef struct prc_object_xf {
  float trans[16];
  float normal[16];
} prc_object_xf;

/* light.Source (light/interface.go:30-37). For a shadow-casting light the host also passes
 * the light camera fitted by initShadowMaps (render/shadow.go:41-86): view, proj, and per
 * object shadow_trans = lightProj.MulM(lightView).MulM(Model) (render/shadow.go:155). */
typedef struct prc_light {
  uint32_t kind;
  uint32_t cast_shadow;
  float pos[3]; /* Position() for point, Dir() for directional */
  float intensity;
  uint32_t color_rgba;
  uint32_t _pad;
  float view[16];
  float proj[16];
  const float* shadow_trans; /* [n_objects][16], NULL unless cast_shadow */
} prc_light;

typedef struct prc_frame {
  uint32_t abi_version;
  uint32_t flags;
  uint32_t width, height; /* size of the frame buffer = render.Size x MSAA (render/raster.go:149); see `msaa` */
  uint32_t n_objects;
  uint32_t n_lights;      /* light sources in Scene.Lights() order (scene/scene.go:35-47) */
  uint32_t n_ambient;
  uint32_t background_rgba;
  const prc_object_xf* objects;
  const prc_light* lights;
  const float* ambient_intensity; /* [n_ambient] Environment.Intensity() */
  float viewport[16];          /* math.ViewportMatrix(w,h) (math/math.go:270-277) */
  float viewport_inv[16];      /* mvp.ViewportInv (render/raster.go:246) */
  float proj_inv[16];          /* mvp.ProjInv (raster.go:245) */
  float view_inv[16];          /* mvp.ViewInv (raster.go:244) */
  float viewport_to_world[16]; /* ViewInv.MulM(ProjInv).MulM(VPInv) (raster.go:287) */
  float cam_pos[3];            /* Camera.Position() */
  uint32_t msaa;               /* render.MSAA(n), 0 or 1 = off (render/options.go:66). n > 1: width/height above are n times render.Size,
                                * the cull / clip box is n times LARGER again (the double-MSAA quirk, raster.go:416-423,438-439,
                                * cull.go:17) and the frame handed back is imageutil.Resize()d to (width/n, height/n) (raster.go:377) */
  uint8_t gamma_lut[256];      /* u8 -> u8 table of shader.GammaCorrection (shader/gamma.go:13-18) */
  /* Multi-GPU screen partition: this ctx shades rows [row0, row1) of SCREEN y (0 = bottom,
   * buffer.go:213). row0 = 0,row1 = height on one GPU. */
  uint32_t row0, row1;
} prc_frame;

/* Parity/debug view of the G-buffer in SCREEN coordinates: element [y*width + x] is what
 * FragmentBuffer.Get(x, y) returns (buffer/buffer.go:209-219). Any pointer may be NULL. */
typedef struct prc_gbuffer_host {
  uint32_t abi_version;
  uint32_t _pad;
  uint8_t* ok;      /* Fragment.Ok */
  int32_t* tri;     /* index of the winning triangle in prc_scene order, -1 if !ok */
  int32_t* sub;     /* 0 = drawn unclipped; k>=1 = k-th fan triangle of clipTriangle (clipping.go:73) */
  float* depth;     /* Fragment.Depth */
  float* uv;        /* [2] U,V */
  float* dudv;      /* [2] Du,Dv */
  float* nor;       /* [3] */
  float* facenor;   /* [3] */
  float* wpos;      /* [3] WordPos */
  uint32_t* col;    /* RGBA8 */
  int32_t* mat;     /* MaterialID */
} prc_gbuffer_host;

typedef struct prc_timings {
  uint32_t abi_version;
  uint32_t _pad;
  float shadow_ms;  /* passShadows, all casting lights */
  float forward_ms; /* passForward: geometry + raster (+ G-buffer resolve when kept) */
  float shade_ms;   /* passDeferred + gamma */
  float total_ms;   /* device time of the whole frame, excluding host copies */
  uint64_t n_valid_tris;  /* triangles passing Triangle.IsValid */
  uint64_t n_nan_frags;   /* fragments with NaN depth seen by the raster (bug-list 8), not resolved by key */
  uint64_t gpu_launches;  /* kernels launched by the last prc_render */
  /* per kernel class, CUDA-event time on the launching stream, summed over the frame:
   * 0 geom+raster (shadow lights)  1 geom+raster (camera)  2 clip  3 binning (count+scan+fill; only when a record needs the tile path)
   * 4 medium raster (queued records, one warp each)  5 tile raster  6 resolve  7 shade */
  float kernel_ms[8];
  uint32_t kernel_launches[8];
  uint64_t n_large_items; /* triangles routed to the tile path, summed over the passes of the frame */
  uint64_t n_clipped;     /* triangles that went through clipTriangle */
  uint64_t n_bin_entries; /* (tile, triangle) pairs */
} prc_timings;

#define PRC_K_GEOM_SHADOW 0
#define PRC_K_GEOM_CAMERA 1
#define PRC_K_CLIP 2      /* (timed inside the camera pass's bracket) */
#define PRC_K_EXCHANGE 2  /* peer groups: k_peer_push */
#define PRC_K_BIN 3
#define PRC_K_MEDIUM 4
#define PRC_K_TILE 5
#define PRC_K_RESOLVE 6
#define PRC_K_SHADE 7

typedef struct prc_ctx prc_ctx;

uint32_t prc_abi_version(void);
int32_t prc_device_count(void);
/* Opens a context on CUDA device `device`. */
int32_t prc_open(int32_t device, prc_ctx** out);
int32_t prc_close(prc_ctx* ctx);
const char* prc_last_error(prc_ctx* ctx);
/* Copies the scene into HBM (host arrays are borrowed for the call only) and precomputes
 * Triangle.IsValid (geometry/primitive/triangle.go:63-80), which is frame-invariant. */
int32_t prc_scene_upload(prc_ctx* ctx, const prc_scene* scene);
/* Zeroes the persistent shadow depth maps (they are never cleared by Render(),
 * render/shadow.go:221-228; only NewRenderer/Options re-create them). */
int32_t prc_shadow_reset(prc_ctx* ctx);
/* Renders one frame. rgba_out: caller-owned width*height*4 bytes in image order
 * (row r = screen y = height-1-r, buffer.go:160-166, 225). With row0/row1 set, only image
 * rows of that strip are written. */
int32_t prc_render(prc_ctx* ctx, const prc_frame* frame, uint8_t* rgba_out);
/* A batch of views of the uploaded scene (BASELINE configs[4]; what a caller of the reference writes as a loop of
 * Renderer.Options(render.Camera(c)) + Render(), render/options.go:125-141): frames[v] is the complete prc_frame of view v (its own
 * camera, light cameras and per-object matrices; PRC_FRAME_SHADOW_RESET when the view re-fits the light cameras, as Options() does),
 * rgba_out[v] the caller-owned image of view v (rgba_out or an entry may be NULL: that view is not copied out). The views are
 * submitted back to back: view v+1's uniforms are uploaded and its geometry runs while view v is copied to the host.
 * PRC_FRAME_ASYNC / KEEP_GBUFFER / UNIFORMS_RESIDENT are rejected. prc_get_timings then reports sums over the views. */
int32_t prc_render_batch(prc_ctx* ctx, uint32_t n_views, const prc_frame* frames, uint8_t* const* rgba_out);
/* Zero-copy result: with rgba_out == NULL the frame is left in one of two library-owned page-locked host images
 * used alternately (the reference's double buffer, render/raster.go:86,201-206: the returned *image.RGBA aliases
 * the buffer until two frames later). Returns that image (width*height*4 bytes, image order). */
int32_t prc_host_image(prc_ctx* ctx, uint64_t* host_ptr, uint64_t* bytes);
int32_t prc_read_gbuffer(prc_ctx* ctx, prc_gbuffer_host* out);
int32_t prc_read_shadowmap(prc_ctx* ctx, uint32_t light, float* out /* [width*height], idx = x + y*width */);
/* The device-resident RGBA8 image (width*height*4 bytes, image order, before any MSAA downsample) copied to host memory:
 * debug/parity view of frames rendered with PRC_FRAME_NO_READBACK (buf.Image(), buffer/buffer.go:160-166). Waits for the stream. */
int32_t prc_read_image(prc_ctx* ctx, uint8_t* rgba_out);
int32_t prc_get_timings(prc_ctx* ctx, prc_timings* out);

/* ---- multi-GPU plumbing (one process per GPU; collectives are driven by the host through
 * torch.distributed/NCCL on these device pointers, on the stream returned here) ---- */
/* Device pointer + byte size of the RGBA8 image (image order; `capacity` >= bytes includes the padding that lets
 * N equal strips be all-gathered in place) and of shadow map `light`. */
int32_t prc_device_image(prc_ctx* ctx, uint64_t* dev_ptr, uint64_t* bytes, uint64_t* capacity);
int32_t prc_device_shadowmap(prc_ctx* ctx, uint32_t light, uint64_t* dev_ptr, uint64_t* bytes);
/* Split frame: phase 1 = shadow passes for the lights/rows this rank owns; the host then
 * all-gathers the maps; phase 2 = forward + deferred for rows [row0,row1). */
int32_t prc_render_shadows(prc_ctx* ctx, const prc_frame* frame, uint32_t light_mask,
                           uint32_t srow0, uint32_t srow1);
/* Same, for n (light, row range) units in one call: up to 8 units share one sweep over the triangles. */
int32_t prc_render_shadow_units(prc_ctx* ctx, const prc_frame* frame, uint32_t n, const uint32_t* light, const uint32_t* row0,
                                const uint32_t* row1);
int32_t prc_render_main(prc_ctx* ctx, const prc_frame* frame, uint8_t* rgba_out);
/* prc_render_main in two halves, so the host can overlap the shadow-map exchange with the camera pass:
 * forward = geometry + raster + resolve for rows [row0,row1) (asynchronous, needs no shadow map);
 * deferred = shading (+ readback), synchronises. */
int32_t prc_render_forward(prc_ctx* ctx, const prc_frame* frame);
int32_t prc_render_deferred(prc_ctx* ctx, const prc_frame* frame, uint8_t* rgba_out);
/* All shadow maps of the casting lights (casting order, [Ls][height][width] float32) as ONE device range, so the
 * exchange is a single in-place all-gather; `capacity` >= bytes includes padding for equal-sized chunks. */
int32_t prc_device_shadow_all(prc_ctx* ctx, uint64_t* dev_ptr, uint64_t* bytes, uint64_t* capacity);
/* cudaStream_t the ctx launches on (as uint64) so the host can order NCCL after it. */
int32_t prc_stream(prc_ctx* ctx, uint64_t* stream);
/* Waits for everything submitted on this context. After PRC_FRAME_ASYNC frames it also finishes them (prc_get_timings then
 * reports sums over those frames) and returns PRC_ERR_RETRY if one of them overflowed an internal queue. */
int32_t prc_sync(prc_ctx* ctx);

/* ---- multi-GPU frames over NVLink peer memory (no host wait and no collective inside a frame) --------------------------
 * One process per GPU. Each rank exports its shadow buffer, image buffer and signal words (CUDA IPC), the host exchanges the
 * handles (any transport: torch.distributed.all_gather_object in polyred_b200/distributed.py) and every rank connects.
 * prc_render_peer then submits one frame of the group WITHOUT waiting on the host:
 *   this rank's shadow units -> their non-empty texels are stored into every peer's maps -> camera pass for rows [row0,row1)
 *   -> wait (on the device) for the peers' shadow rows -> shading -> the image strip is copied into the image of every
 *   rank in `image_mask` (bit r = rank r receives the whole frame; north_star: rank 0, mask 1).
 * Ranks are ordered by epoch words in peer memory (release/acquire at system scope); every rank must submit the same
 * sequence of prc_render_peer calls, each with the same image_mask on every rank. Frames stay on the device unless prc_set_host_image
 * registered a host image (PRC_FRAME_KEEP_GBUFFER and PRC_FRAME_SHADOW_RESET are rejected; MSAA frames — strips must start and
 * end on multiples of msaa — are downsampled per strip on the rank that shaded it, which shades `msaa` extra rows on either
 * side for the filter, and need image_mask = 0: they leave through the host image); a consumer's image of a frame stays valid until its
 * next prc_render_peer (readers: the host after prc_sync, or — unless the frame carried PRC_FRAME_IMAGE_AT_SYNC — work enqueued on prc_stream() before that call). prc_sync() finishes the submitted frames:
 * PRC_ERR_RETRY = a queue overflowed on THIS rank (grown now; all ranks must agree to submit the frames again),
 * PRC_ERR_PEER = a device-side wait for a peer gave up after 4 s.
 * The result is the 1-GPU frame bit for bit: depth maxima do not depend on who rasterised which rows. */
typedef struct prc_peer_handle {
  uint32_t abi_version;
  uint32_t device;        /* CUDA device ordinal of the exporting context */
  uint64_t pid;           /* exporting process: within one process the pointers below are used directly */
  uint64_t shadow_ptr, image_ptr, signals_ptr;  /* device addresses in the exporting process */
  uint64_t shadow_off, image_off, signals_off;  /* offsets of those buffers inside their IPC allocations */
  uint64_t shadow_bytes, image_bytes;           /* capacities; all ranks of a group must agree */
  uint8_t shadow_ipc[64], image_ipc[64], signals_ipc[64]; /* cudaIpcMemHandle_t */
  uint64_t mkeys_ptr, mkeys_off, mkeys_bytes;   /* the merged visibility-key planes (2 frame parities + 2 NaN-mode planes) */
  uint8_t mkeys_ipc[64];
} prc_peer_handle;

/* Allocates this context's buffers for `frame` (size, casting lights), zeroes its signal words and fills `out`.
 * Call on every rank, exchange, then prc_peer_connect; repeat after a change of frame size or light set. */
int32_t prc_peer_export(prc_ctx* ctx, const prc_frame* frame, prc_peer_handle* out);
/* `all` = the world's handles in rank order (all[rank] is this context's own). world <= 16. */
int32_t prc_peer_connect(prc_ctx* ctx, uint32_t rank, uint32_t world, const prc_peer_handle* all);
int32_t prc_peer_disconnect(prc_ctx* ctx);
/* row0[p], row1[p]: the strip of screen rows rank p shades, for every rank of the group (n_ranks = world; this rank's entry
 * must equal frame->row0/row1; every strip needs at least one row). */
int32_t prc_render_peer(prc_ctx* ctx, const prc_frame* frame, uint32_t n_ranks, const uint32_t* row0, const uint32_t* row1,
                        uint32_t image_mask);
/* With PRC_PEER_TRACE=1 in the environment at prc_peer_connect: milliseconds this context's stream spent in the device-side
 * waits of the frames finished by the last prc_sync, per signal kind: [0] peers' shadow rows, [1] peers done shading the
 * previous frame, [2] strips arriving in this rank's image, [3] consumers done with the previous image. Zeros otherwise. */
int32_t prc_peer_wait_ms(prc_ctx* ctx, float out[4]);
/* A caller-owned host image as the readback destination of prc_render_peer frames submitted WITHOUT PRC_FRAME_NO_READBACK:
 * [ptr, ptr + bytes) is page-locked and every such frame's strip (image rows of [row0,row1)) is DMA'd into it band by band
 * behind the shading kernels, complete after prc_sync. Meant for ONE shared-memory image mapped by every rank of a group:
 * each GPU then delivers its own strip over its own PCIe link and no device-side gather is needed (image_mask = 0).
 * ptr == NULL unregisters. */
int32_t prc_set_host_image(prc_ctx* ctx, void* ptr, uint64_t bytes);
/* The frames submitted from now on are read back to ptr + offset of the registered range: several images inside ONE registration
 * (page-locking is expensive), e.g. two used alternately like the reference's double buffer (render/raster.go:86,201-206). */
int32_t prc_set_host_image_offset(prc_ctx* ctx, uint64_t offset);

/* Sticky per-context render state that a frame can switch on (the reporting frame is re-rendered, PRC_ERR_RETRY for frames submitted
 * back to back): NaN mode (a camera pass produced a NaN-depth fragment: the first-fragment plane of bug-list 8 is maintained from then
 * on) and the binned tile path (a queued triangle was too large for the one-warp raster). The ranks of a peer group must agree on
 * NaN mode — the first fragment of a pixel may come from any rank — so after a PRC_ERR_RETRY the host ORs the ranks' states and
 * sets the result on every rank before the frames are submitted again (polyred_b200/distributed.py PeerFrames.finish). */
#define PRC_STATE_NAN_MODE 1u
#define PRC_STATE_TILE_PATH 2u
int32_t prc_frame_state(prc_ctx* ctx, uint32_t* state);
int32_t prc_set_frame_state(prc_ctx* ctx, uint32_t state);
/* Number of covered pixels (Fragment.Ok, buffer/buffer.go:209-219) of the rows this context rasterised in its last frame —
 * for the coverage-weighted roofline of the measurement harness. */
int32_t prc_count_covered(prc_ctx* ctx, uint64_t* covered);
/* Measured FP32 FMA throughput of this device in TFLOP/s (2 flop per FMA; a pure-FMA micro-benchmark, best of a few launches):
 * the denominator of the shading kernels' roofline. No reference counterpart (measurement only). */
int32_t prc_measure_fp32_peak(prc_ctx* ctx, double* tflops);

/* ---- device groups: ONE process, one context per device, one submit thread per device inside the library ----------------
 * What `render.CUDA(devices ...int)` binds when it is given more than one device: the multi-device
 * counterpart of render.GPU(dev) (render/options.go:103-110; the reference's gpu.Open takes one device, gpu/device.go:77-89).
 * A group owns its contexts; the calls below mirror the single-context ones and hide the whole peer protocol above
 * (export / connect over cudaDeviceEnablePeerAccess, the strip partition and its load balancing, the retry vote):
 *   prc_group_render       one frame over all devices = prc_render_peer on every context from the group's own threads:
 *                          raster passes partitioned by triangles, shading by screen strips balanced by measured time, every
 *                          device DMAs its strip into ONE page-locked host image (rgba_out may be NULL: read it in place through
 *                          prc_group_host_image). frame->row0/row1 must be 0/height. The result is the 1-GPU frame bit for bit.
 *                          With PRC_FRAME_NO_READBACK the strips are gathered into context 0's device image instead
 *                          (prc_group_ctx(g, 0) + prc_device_image), and with PRC_FRAME_ASYNC on top the call does not wait:
 *                          prc_group_sync finishes the frames (PRC_ERR_RETRY = submit them again, as for prc_sync).
 *   prc_group_render_views a batch of views dealt round-robin to the devices (view v on device v mod n), each device running
 *                          prc_render_batch on its share: no exchange at all (BASELINE configs[4]).
 * A device may be listed more than once (several contexts of one GPU: how the single-GPU tests exercise the protocol).
 * Every call returns 0 or a PRC_ERR_* code; prc_group_last_error names the failing rank. One call at a time per group. */
typedef struct prc_group prc_group;
int32_t prc_group_open(const int32_t* devices, uint32_t n, prc_group** out);
int32_t prc_group_close(prc_group* g);
const char* prc_group_last_error(prc_group* g);
uint32_t prc_group_size(prc_group* g);
/* The context of rank `rank` (owned by the group) for the read-only / debug calls: prc_get_timings, prc_read_shadowmap,
 * prc_device_image, prc_peer_wait_ms. */
int32_t prc_group_ctx(prc_group* g, uint32_t rank, prc_ctx** out);
int32_t prc_group_scene_upload(prc_group* g, const prc_scene* scene); /* replicated on every device, uploads run in parallel */
int32_t prc_group_shadow_reset(prc_group* g);
int32_t prc_group_render(prc_group* g, const prc_frame* frame, uint8_t* rgba_out);
int32_t prc_group_sync(prc_group* g);
int32_t prc_group_host_image(prc_group* g, uint64_t* host_ptr, uint64_t* bytes);
/* The strip (screen rows [row0, row1)) every rank shaded in the last frame; arrays of prc_group_size entries. */
int32_t prc_group_strips(prc_group* g, uint32_t* row0, uint32_t* row1);
int32_t prc_group_render_views(prc_group* g, uint32_t n_views, const prc_frame* frames, uint8_t* const* rgba_out);

/* Arithmetic mode of this context, overriding the PRC_FMA environment variable read by prc_open:
 * exact != 0 -> math.FMA[float32] emulated bit-exactly everywhere (float64 fma rounded to float32, math/math.go FMA),
 * exact == 0 -> single-rounding fmaf everywhere ("fast"). The default, "mixed", is only selectable through PRC_FMA. */
int32_t prc_set_exact_fma(prc_ctx* ctx, int32_t exact);

#ifdef __cplusplus
}
#endif
#endif /* POLYRED_CUDA_H */
