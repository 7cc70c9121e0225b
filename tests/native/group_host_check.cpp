// CPU check of the device-group host logic (polyred_b200/csrc/prc_group.cpp) against STUB contexts: no CUDA, no GPU.
// prc_group.cpp is built on the library's public calls only, so this file defines those calls as recording stubs and includes
// the group source itself. What is checked (tests/test_group_host_native.py runs it and reads the OK / FAIL lines):
//   * one submit thread per rank really submits: every rank receives every frame exactly once, with ITS strip and the strip table
//     of all ranks, frame.row0/row1 rewritten, PRC_FRAME_ASYNC stripped, PRC_FRAME_IMAGE_AT_SYNC added to asynchronous frames;
//   * the strips always partition the frame (top image rows = rank 0), also while the group re-balances them by the measured
//     resolve + shading times, and converge towards equal cost on a frame whose cost per row is skewed;
//   * a PRC_ERR_RETRY from ONE rank makes EVERY rank render the frame again after the ranks' frame states were OR-ed and set
//     on all of them; asynchronous frames report it from prc_group_sync;
//   * an error on one rank fails the call with that rank named, finishes the others and drops the connection (the next frame
//     exports and connects again);
//   * a new frame size or light set reconnects; the host image is registered once per size on every rank, offsets alternate;
//   * view batches are dealt round-robin and disconnect the group first.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/polyred_cuda.h"

// ---------------------------------------------------------------------------------------------------------------------------------
struct Call {
  uint32_t flags, row0, row1, image_mask, width, height;
  std::vector<uint32_t> r0, r1;
};
struct prc_ctx {
  int device = 0, rank = -1;
  std::string err;
  bool exported = false, connected = false;
  uint32_t world = 0;
  int exports = 0, connects = 0, disconnects = 0, syncs = 0, uploads = 0, shadow_resets = 0, state_sets = 0;
  uint32_t state = 0, state_set_to = 0;
  std::vector<Call> frames;
  int pending = 0;
  int fail_retry_at = -1;   // index of the frame (per ctx) whose prc_sync reports PRC_ERR_RETRY once
  int fail_error_at = -1;   // index of the frame whose prc_render_peer fails
  void* host = nullptr; uint64_t host_bytes = 0, host_off = 0; int host_sets = 0;
  std::vector<uint64_t> offsets;
  std::vector<uint32_t> batch_sizes;
  prc_timings tm{};
};
static std::mutex g_mu;
static std::vector<prc_ctx*> g_all;
// cost model of the fake GPU: image row r (0 = top) costs w(r); a rank's shading time = the sum over its strip
static double row_cost(int image_row, int H) { return image_row < H / 2 ? 0.2 : 1.8; }  // the lower half of the frame is 9x as expensive

extern "C" {
int32_t prc_open(int32_t device, prc_ctx** out) {
  if (device < 0 || device >= 64) return PRC_ERR_CUDA;
  prc_ctx* c = new prc_ctx();
  c->device = device;
  std::lock_guard<std::mutex> lk(g_mu);
  g_all.push_back(c);
  *out = c;
  return PRC_OK;
}
int32_t prc_close(prc_ctx*) { return PRC_OK; }  // (kept alive: the checks read the records afterwards)
const char* prc_last_error(prc_ctx* c) { return c ? c->err.c_str() : "null"; }
int32_t prc_scene_upload(prc_ctx* c, const prc_scene*) { c->uploads++; return PRC_OK; }
int32_t prc_shadow_reset(prc_ctx* c) { c->shadow_resets++; return PRC_OK; }
int32_t prc_render(prc_ctx* c, const prc_frame* f, uint8_t*) { c->frames.push_back({f->flags, f->row0, f->row1, 0, f->width, f->height, {}, {}}); return PRC_OK; }
int32_t prc_render_batch(prc_ctx* c, uint32_t n, const prc_frame*, uint8_t* const*) { c->batch_sizes.push_back(n); return PRC_OK; }
int32_t prc_host_image(prc_ctx*, uint64_t*, uint64_t*) { return PRC_ERR_INVALID; }
int32_t prc_peer_export(prc_ctx* c, const prc_frame* f, prc_peer_handle* out) {
  c->exports++; c->exported = true; c->connected = false;
  memset(out, 0, sizeof(*out));
  out->abi_version = PRC_ABI_VERSION; out->device = (uint32_t)c->device; out->shadow_ptr = (uint64_t)(uintptr_t)c;
  (void)f;
  return PRC_OK;
}
int32_t prc_peer_connect(prc_ctx* c, uint32_t rank, uint32_t world, const prc_peer_handle* all) {
  if (!c->exported || all[rank].shadow_ptr != (uint64_t)(uintptr_t)c) { c->err = "connect without export / wrong handle order"; return PRC_ERR_INVALID; }
  c->connects++; c->connected = true; c->rank = (int)rank; c->world = world;
  return PRC_OK;
}
int32_t prc_peer_disconnect(prc_ctx* c) { c->disconnects++; c->connected = false; c->exported = false; c->pending = 0; return PRC_OK; }
int32_t prc_render_peer(prc_ctx* c, const prc_frame* f, uint32_t n, const uint32_t* r0, const uint32_t* r1, uint32_t image_mask) {
  if (!c->connected) { c->err = "not connected"; return PRC_ERR_INVALID; }
  if ((int)c->frames.size() == c->fail_error_at) { c->fail_error_at = -1; c->frames.push_back({}); c->err = "injected failure"; return PRC_ERR_CUDA; }
  Call k{f->flags, f->row0, f->row1, image_mask, f->width, f->height, std::vector<uint32_t>(r0, r0 + n), std::vector<uint32_t>(r1, r1 + n)};
  c->frames.push_back(k);
  c->pending++;
  // fake timings of this frame: shading time = cost of the strip's image rows
  const int H = (int)f->height;
  double cost = 0;
  for (uint32_t y = f->row0; y < f->row1; y++) cost += row_cost(H - 1 - (int)y, H);
  memset(&c->tm, 0, sizeof(c->tm));
  c->tm.abi_version = PRC_ABI_VERSION;
  c->tm.kernel_ms[PRC_K_SHADE] = (float)(cost * 1e-3);
  c->tm.kernel_ms[PRC_K_GEOM_CAMERA] = 0.1f;
  return PRC_OK;
}
int32_t prc_sync(prc_ctx* c) {
  c->syncs++;
  const int last = (int)c->frames.size() - 1;
  c->pending = 0;
  if (c->fail_retry_at >= 0 && last >= c->fail_retry_at) { c->fail_retry_at = -1; c->state |= PRC_STATE_TILE_PATH; c->err = "queue grown"; return PRC_ERR_RETRY; }
  return PRC_OK;
}
int32_t prc_get_timings(prc_ctx* c, prc_timings* out) { *out = c->tm; return PRC_OK; }
int32_t prc_frame_state(prc_ctx* c, uint32_t* s) { *s = c->state; return PRC_OK; }
int32_t prc_set_frame_state(prc_ctx* c, uint32_t s) { c->state_sets++; c->state_set_to = s; c->state |= s; return PRC_OK; }
int32_t prc_set_host_image(prc_ctx* c, void* p, uint64_t bytes) { c->host = p; c->host_bytes = bytes; if (p) c->host_sets++; return PRC_OK; }
int32_t prc_set_host_image_offset(prc_ctx* c, uint64_t off) {
  if (!c->host || off >= c->host_bytes) { c->err = "offset outside the registered image"; return PRC_ERR_INVALID; }
  c->host_off = off; c->offsets.push_back(off);
  return PRC_OK;
}
}

#include "../../polyred_b200/csrc/prc_group.cpp"

// ---------------------------------------------------------------------------------------------------------------------------------
static int g_fail = 0;
#define CHECK(cond, ...)                          \
  do {                                            \
    if (!(cond)) {                                \
      g_fail++;                                   \
      printf("FAIL %s:%d  ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                        \
      printf("\n");                               \
    }                                             \
  } while (0)

static prc_frame make_frame(uint32_t w, uint32_t h, uint32_t flags, prc_light* lights, uint32_t n_lights) {
  prc_frame f;
  memset(&f, 0, sizeof(f));
  f.abi_version = PRC_ABI_VERSION;
  f.width = w; f.height = h; f.flags = flags; f.row0 = 0; f.row1 = h; f.msaa = 1;
  f.lights = lights; f.n_lights = n_lights;
  return f;
}

// the strips of the last frame every rank received: a partition of [0, H), rank 0 on top, identical tables on every rank
static void check_partition(const std::vector<prc_ctx*>& cs, uint32_t H, const char* what) {
  const size_t n = cs.size();
  const Call& a = cs[0]->frames.back();
  CHECK(a.r0.size() == n && a.r1.size() == n, "%s: strip table has %zu entries for %zu ranks", what, a.r0.size(), n);
  for (size_t r = 0; r < n; r++) {
    const Call& k = cs[r]->frames.back();
    CHECK(k.r0 == a.r0 && k.r1 == a.r1, "%s: rank %zu got a different strip table", what, r);
    CHECK(k.row0 == a.r0[r] && k.row1 == a.r1[r], "%s: rank %zu frame.row0/row1 = %u..%u, table says %u..%u", what, r, k.row0, k.row1, a.r0[r], a.r1[r]);
    CHECK(k.row0 < k.row1, "%s: rank %zu has an empty strip", what, r);
  }
  CHECK(a.r1[0] == H && a.r0[n - 1] == 0, "%s: the strips do not span the frame (%u..%u)", what, a.r0[n - 1], a.r1[0]);
  for (size_t r = 0; r + 1 < n; r++) CHECK(a.r0[r] == a.r1[r + 1], "%s: gap / overlap between rank %zu and %zu", what, r, r + 1);
}

int main() {
  const uint32_t W = 640, H = 480;
  prc_light lights[2];
  memset(lights, 0, sizeof(lights));
  lights[0].cast_shadow = 1;
  float dummy_trans[16] = {0};
  lights[0].shadow_trans = dummy_trans;

  // ---- 4 ranks: synchronous frames into the host image, re-balancing -----------------------------------------------------------
  {
    const int32_t devs[4] = {0, 1, 2, 3};
    prc_group* g = nullptr;
    CHECK(prc_group_open(devs, 4, &g) == PRC_OK && g, "prc_group_open");
    std::vector<prc_ctx*> cs(g_all.end() - 4, g_all.end());
    prc_scene sc;
    memset(&sc, 0, sizeof(sc));
    CHECK(prc_group_scene_upload(g, &sc) == PRC_OK, "scene upload");
    for (prc_ctx* c : cs) CHECK(c->uploads == 1, "every rank uploads the scene once (got %d)", c->uploads);
    prc_frame f = make_frame(W, H, PRC_FRAME_SHADOWMAP | PRC_FRAME_GAMMA, lights, 2);
    std::vector<uint8_t> out((size_t)W * H * 4);
    double first_spread = 0, last_spread = 0;
    for (int k = 0; k < 14; k++) {
      const int32_t rc = prc_group_render(g, &f, out.data());
      CHECK(rc == PRC_OK, "frame %d: rc %d (%s)", k, rc, prc_group_last_error(g));
      for (size_t r = 0; r < 4; r++) CHECK((int)cs[r]->frames.size() == k + 1, "frame %d: rank %zu has rendered %zu frames", k, r, cs[r]->frames.size());
      check_partition(cs, H, "sync frame");
      for (prc_ctx* c : cs) {
        const Call& last = c->frames.back();
        CHECK(!(last.flags & PRC_FRAME_ASYNC) && !(last.flags & PRC_FRAME_IMAGE_AT_SYNC), "a synchronous frame must carry neither ASYNC nor IMAGE_AT_SYNC (flags %u)", last.flags);
        CHECK(last.image_mask == 0, "frames for the host leave through the host image: image_mask must be 0, got %u", last.image_mask);
      }
      double mn = 1e30, mx = 0;
      for (prc_ctx* c : cs) { mn = std::min(mn, (double)c->tm.kernel_ms[PRC_K_SHADE]); mx = std::max(mx, (double)c->tm.kernel_ms[PRC_K_SHADE]); }
      if (k == 0) first_spread = mx / mn;
      last_spread = mx / mn;
    }
    CHECK(first_spread > 5.0, "the cost model is skewed (equal strips: max/min shading time %.2f)", first_spread);
    CHECK(last_spread < 1.35, "re-balancing did not converge: max/min shading time %.2f after 14 frames (%.2f at the start)", last_spread, first_spread);
    for (prc_ctx* c : cs) {
      CHECK(c->exports == 1 && c->connects == 1, "one export / connect for 14 frames of one signature (got %d / %d)", c->exports, c->connects);
      CHECK(c->host_sets == 1, "the host image is registered once per size on every rank (got %d)", c->host_sets);
      CHECK(c->offsets.size() == 14, "every frame selects its half of the double-buffered host image");
      for (size_t i = 1; i < c->offsets.size(); i++) CHECK(c->offsets[i] != c->offsets[i - 1], "consecutive frames must alternate between the two host images");
    }
    uint32_t r0[4], r1[4];
    CHECK(prc_group_strips(g, r0, r1) == PRC_OK && r1[0] == H && r0[3] == 0, "prc_group_strips");
    CHECK((r1[0] - r0[0]) > 2 * (r1[3] - r0[3]), "the cheap top rows should end up in a much taller strip (%u vs %u rows)", r1[0] - r0[0], r1[3] - r0[3]);

    // ---- a queue overflow on ONE rank: every rank renders the frame again, states agreed first --------------------------------
    const size_t before = cs[0]->frames.size();
    cs[2]->fail_retry_at = (int)cs[2]->frames.size();
    CHECK(prc_group_render(g, &f, out.data()) == PRC_OK, "frame with one retry: %s", prc_group_last_error(g));
    for (size_t r = 0; r < 4; r++) {
      CHECK(cs[r]->frames.size() == before + 2, "rank %zu must render the frame twice after rank 2's PRC_ERR_RETRY (rendered %zu)", r, cs[r]->frames.size() - before);
      CHECK(cs[r]->state_sets >= 1 && (cs[r]->state_set_to & PRC_STATE_TILE_PATH), "rank %zu must be given the OR of the ranks' frame states before the retry", r);
    }
    check_partition(cs, H, "retried frame");

    // ---- a new frame size: export + connect again, equal strips again ---------------------------------------------------------
    prc_frame f2 = make_frame(W, 240, PRC_FRAME_SHADOWMAP, lights, 2);
    std::vector<uint8_t> out2((size_t)W * 240 * 4);
    CHECK(prc_group_render(g, &f2, out2.data()) == PRC_OK, "smaller frame: %s", prc_group_last_error(g));
    for (prc_ctx* c : cs) CHECK(c->exports == 2 && c->connects == 2, "a new frame size reconnects (exports %d, connects %d)", c->exports, c->connects);
    check_partition(cs, 240, "new size");
    CHECK(cs[0]->frames.back().row1 - cs[0]->frames.back().row0 == 60, "a fresh connection starts from equal strips");
    // a new set of casting lights too
    lights[1].cast_shadow = 1; lights[1].shadow_trans = dummy_trans;
    CHECK(prc_group_render(g, &f2, out2.data()) == PRC_OK, "second caster: %s", prc_group_last_error(g));
    for (prc_ctx* c : cs) CHECK(c->exports == 3, "a new set of casting lights reconnects (exports %d)", c->exports);
    lights[1].cast_shadow = 0;

    // ---- an error on one rank: reported with the rank, the connection is dropped, the next frame reconnects ------------------
    cs[1]->fail_error_at = (int)cs[1]->frames.size();
    const int32_t rc = prc_group_render(g, &f2, out2.data());
    CHECK(rc == PRC_ERR_CUDA, "an injected failure on rank 1 must fail the frame (rc %d)", rc);
    CHECK(std::string(prc_group_last_error(g)).find("rank 1") != std::string::npos, "the error names the failing rank: '%s'", prc_group_last_error(g));
    for (prc_ctx* c : cs) CHECK(!c->connected, "a failed frame drops the connection on every rank");
    const int exports_before = cs[0]->exports;
    CHECK(prc_group_render(g, &f2, out2.data()) == PRC_OK, "the frame after a failure: %s", prc_group_last_error(g));
    for (prc_ctx* c : cs) CHECK(c->exports == exports_before + 1 && c->connected, "the next frame exports and connects again");

    // ---- rejected inputs ----------------------------------------------------------------------------------------------------------
    prc_frame bad = f2;
    bad.row0 = 8;
    CHECK(prc_group_render(g, &bad, out2.data()) == PRC_ERR_INVALID, "a caller-chosen row range is rejected");
    bad = f2; bad.flags |= PRC_FRAME_KEEP_GBUFFER;
    CHECK(prc_group_render(g, &bad, out2.data()) == PRC_ERR_UNSUPPORTED, "KEEP_GBUFFER is rejected");
    bad = f2; bad.flags |= PRC_FRAME_ASYNC;
    CHECK(prc_group_render(g, &bad, nullptr) == PRC_ERR_INVALID, "ASYNC without NO_READBACK is rejected");
    prc_frame tiny = make_frame(64, 2, 0, lights, 2);
    CHECK(prc_group_render(g, &tiny, nullptr) == PRC_ERR_INVALID, "2 rows cannot be cut into 4 strips");

    // ---- asynchronous device-resident frames -----------------------------------------------------------------------------------
    prc_frame fa = make_frame(W, H, PRC_FRAME_SHADOWMAP | PRC_FRAME_NO_READBACK | PRC_FRAME_ASYNC, lights, 2);
    const size_t n0 = cs[0]->frames.size();
    const int syncs0 = cs[0]->syncs;
    for (int k = 0; k < 5; k++) CHECK(prc_group_render(g, &fa, nullptr) == PRC_OK, "async frame %d: %s", k, prc_group_last_error(g));
    for (size_t r = 0; r < 4; r++) {
      CHECK(cs[r]->frames.size() == n0 + 5, "rank %zu: 5 asynchronous frames submitted", r);
      const Call& last = cs[r]->frames.back();
      CHECK((last.flags & PRC_FRAME_IMAGE_AT_SYNC) && !(last.flags & PRC_FRAME_ASYNC) && (last.flags & PRC_FRAME_NO_READBACK), "asynchronous frames: IMAGE_AT_SYNC set, ASYNC stripped (flags %u)", last.flags);
      CHECK(last.image_mask == 1, "device-resident frames are gathered into rank 0's image (mask %u)", last.image_mask);
    }
    CHECK(cs[0]->syncs == syncs0 + 1 || cs[0]->syncs == syncs0, "asynchronous frames do not wait (syncs %d -> %d; one is the reconnect's)", syncs0, cs[0]->syncs);
    check_partition(cs, H, "async frame");
    cs[3]->fail_retry_at = 0;
    CHECK(prc_group_sync(g) == PRC_ERR_RETRY, "prc_group_sync reports a rank's PRC_ERR_RETRY");
    CHECK(prc_group_sync(g) == PRC_OK, "and is clean afterwards");

    // ---- view batches: dealt round-robin, no exchange ---------------------------------------------------------------------------
    std::vector<prc_frame> views(10, make_frame(W, H, PRC_FRAME_SHADOWMAP | PRC_FRAME_SHADOW_RESET, lights, 2));
    CHECK(prc_group_render_views(g, 10, views.data(), nullptr) == PRC_OK, "view batch: %s", prc_group_last_error(g));
    const uint32_t want[4] = {3, 3, 2, 2};
    for (size_t r = 0; r < 4; r++) {
      CHECK(cs[r]->batch_sizes.size() == 1 && cs[r]->batch_sizes[0] == want[r], "rank %zu renders %u of 10 views", r, want[r]);
      CHECK(!cs[r]->connected, "a view batch disconnects the group first (whole frames per context)");
    }
    CHECK(prc_group_shadow_reset(g) == PRC_OK, "shadow reset");
    for (prc_ctx* c : cs) CHECK(c->shadow_resets == 1, "every rank zeroes its maps");
    CHECK(prc_group_close(g) == PRC_OK, "close");
  }
  // ---- a group of one is the plain context ------------------------------------------------------------------------------------
  {
    const int32_t dev = 5;
    prc_group* g = nullptr;
    CHECK(prc_group_open(&dev, 1, &g) == PRC_OK, "group of one");
    prc_ctx* c = g_all.back();
    prc_frame f = make_frame(W, H, PRC_FRAME_SHADOWMAP, lights, 2);
    CHECK(prc_group_render(g, &f, nullptr) == PRC_OK, "group of one renders through prc_render");
    CHECK(c->frames.size() == 1 && c->exports == 0 && c->frames[0].row0 == 0 && c->frames[0].row1 == H, "no peer machinery for one device");
    prc_group_close(g);
    const int32_t none[1] = {4096};
    CHECK(prc_group_open(none, 1, &g) != PRC_OK, "a device that does not exist fails prc_group_open");
    CHECK(prc_group_open(nullptr, 0, &g) == PRC_ERR_INVALID, "no devices");
  }
  printf(g_fail ? "FAILED %d checks\n" : "OK all checks passed\n", g_fail);
  return g_fail ? 1 : 0;
}
