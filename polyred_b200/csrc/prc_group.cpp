// prc_group.cpp — device groups behind the C ABI (include/polyred_cuda.h "device groups"): ONE process drives N contexts, one per
// device, each from its own submit thread, so that a cgo-free Go host (purego cannot call torch.distributed) gets multi-GPU frames
// from a single call. Built on the library's own public calls: prc_peer_export / prc_peer_connect (same process: the peers'
// pointers are used directly after cudaDeviceEnablePeerAccess), prc_render_peer, prc_sync, prc_set_host_image, prc_render_batch.
// The reference has no counterpart (gpu.Open takes one device, gpu/device.go:77-89): the contract is
// "the same frame as one context, bit for bit" (tests/test_gpu_group.py).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/polyred_cuda.h"

namespace {

// One submit thread per context. A job is posted to all workers at once; a worker spins briefly for the next job (frames of an
// 8-GPU group take ~0.2 ms: a condition-variable wake-up per frame would be a tenth of that) and then blocks.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<uint64_t> posted{0}, done{0};
  int32_t rc = 0;
  std::atomic<bool> quit{false};
};

}  // namespace

struct prc_group {
  std::vector<prc_ctx*> ctx;
  std::vector<int32_t> devices;
  std::vector<Worker*> workers;
  std::function<int32_t(uint32_t)> job;
  std::string err;
  // connection state: the signature of the frame the contexts were exported / connected for
  bool connected = false;
  uint32_t sig_w = 0, sig_h = 0, sig_msaa = 0, sig_lights = 0, sig_flags = 0;
  std::vector<uint8_t> sig_cast;
  // shading partition: bounds[k] .. bounds[k+1] = OUTPUT image rows of rank k (image row r = screen y = H-1-r)
  std::vector<int> bounds;
  std::vector<uint32_t> row0, row1;  // screen rows of the frame buffer per rank (what prc_render_peer takes)
  uint64_t frames_since_connect = 0;
  bool balance = true;
  int stagger_ms = 0;  // PRC_GROUP_STAGGER_MS (tests): rank r submits every frame r x this many ms late, so that its peers are already waiting for it
  // page-locked host images shared by all contexts (each DMAs its own strip): TWO images inside one registration, used alternately
  // like the reference's double buffer (render/raster.go:86,201-206): a frame read in place stays valid during the next Render()
  uint8_t* host_img = nullptr;
  size_t host_bytes = 0, host_cap = 0, host_stride = 0;
  int host_cur = 0;
  int pending = 0;  // PRC_FRAME_ASYNC frames submitted since the last sync
};

namespace {

uint32_t world(const prc_group* g) { return (uint32_t)g->ctx.size(); }

void worker_main(prc_group* g, uint32_t rank) {
  Worker& w = *g->workers[rank];
  uint64_t seen = 0;
  for (;;) {
    // spin a little, then block
    int spins = 0;
    while (w.posted.load(std::memory_order_acquire) == seen) {
      if (++spins < 20000) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        continue;
      }
      std::unique_lock<std::mutex> lk(w.mu);
      w.cv.wait(lk, [&] { return w.posted.load(std::memory_order_acquire) != seen || w.quit.load(); });
      break;
    }
    if (w.quit.load()) return;
    seen = w.posted.load(std::memory_order_acquire);
    w.rc = g->job(rank);
    w.done.store(seen, std::memory_order_release);
  }
}

// Runs fn(rank) for every rank concurrently and waits; rcs[rank] = its return code.
void run_all(prc_group* g, const std::function<int32_t(uint32_t)>& fn, std::vector<int32_t>& rcs) {
  const uint32_t n = world(g);
  rcs.assign(n, 0);
  if (n == 1) { rcs[0] = fn(0); return; }
  g->job = fn;
  for (uint32_t r = 0; r < n; r++) {
    Worker& w = *g->workers[r];
    {
      std::lock_guard<std::mutex> lk(w.mu);
      w.posted.fetch_add(1, std::memory_order_release);
    }
    w.cv.notify_one();
  }
  for (uint32_t r = 0; r < n; r++) {
    Worker& w = *g->workers[r];
    const uint64_t want = w.posted.load(std::memory_order_acquire);
    while (w.done.load(std::memory_order_acquire) != want) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    rcs[r] = w.rc;
  }
}

// first failing rank -> group error string; returns its code (0 if none). `skip` = a code that does not count as a failure.
int32_t first_error(prc_group* g, const std::vector<int32_t>& rcs, const char* what, int32_t skip = 0) {
  for (uint32_t r = 0; r < rcs.size(); r++)
    if (rcs[r] != 0 && rcs[r] != skip) {
      g->err = std::string(what) + ": rank " + std::to_string(r) + " (device " + std::to_string(g->devices[r]) + "): " + prc_last_error(g->ctx[r]);
      return rcs[r];
    }
  return 0;
}

// ---- the shading partition (polyred_b200/partition.py is the Python original; tests/test_partition_gloo.py pins its behaviour) ----
std::vector<int> equal_bounds(int total, int parts) {
  std::vector<int> b(parts + 1);
  const int chunk = (total + parts - 1) / parts;
  for (int k = 0; k <= parts; k++) b[k] = std::min(total, k * chunk);
  return b;
}

// New boundaries such that every range gets the same share of the measured cost, the cost of a range spread evenly over its rows;
// `damping` < 1 moves part of the way; at least `min_size` rows per range; multiples of `align` (MSAA strips).
std::vector<int> balanced_bounds(const std::vector<int>& bounds, const std::vector<double>& cost, double damping, int min_size) {
  const int n = (int)cost.size();
  const int lo = bounds.front(), hi = bounds.back();
  double total = 0;
  for (double c : cost) total += std::max(0.0, c);
  if (n == 1 || hi - lo <= 0 || total <= 0.0) return bounds;
  const double target = total / n;
  std::vector<int> nb(1, lo);
  int k = 0;
  double done = 0.0;
  for (int j = 1; j < n; j++) {
    const double want = j * target;
    while (k < n - 1 && done + std::max(0.0, cost[k]) < want) { done += std::max(0.0, cost[k]); k++; }
    const int size = bounds[k + 1] - bounds[k];
    const double c = std::max(0.0, cost[k]);
    double x = bounds[k] + ((c > 0.0 && size > 0) ? size * (want - done) / c : 0.0);
    x = bounds[j] + damping * (x - bounds[j]);
    nb.push_back((int)std::lround(x));
  }
  nb.push_back(hi);
  const int m = (hi - lo) >= n * min_size ? min_size : 0;
  for (int j = 1; j < n; j++) nb[j] = std::max(nb[j], nb[j - 1] + m);
  for (int j = n - 1; j > 0; j--) nb[j] = std::min(nb[j], nb[j + 1] - m);
  return nb;
}

void apply_bounds(prc_group* g, int out_h, int msaa) {
  const uint32_t n = world(g);
  g->row0.resize(n);
  g->row1.resize(n);
  for (uint32_t k = 0; k < n; k++) {
    g->row0[k] = (uint32_t)((out_h - g->bounds[k + 1]) * msaa);
    g->row1[k] = (uint32_t)((out_h - g->bounds[k]) * msaa);
  }
}

int32_t disconnect(prc_group* g) {
  if (!g->connected) return PRC_OK;
  std::vector<int32_t> rcs;
  run_all(g, [&](uint32_t r) { return prc_peer_disconnect(g->ctx[r]); }, rcs);
  g->connected = false;
  g->pending = 0;
  return first_error(g, rcs, "prc_peer_disconnect", PRC_ERR_RETRY);
}

int32_t set_host_image(prc_group* g, size_t bytes) {
  std::vector<int32_t> rcs;
  if (g->host_img && 2 * ((bytes + 4095) & ~(size_t)4095) <= g->host_cap) {
    g->host_bytes = bytes;
    g->host_stride = (bytes + 4095) & ~(size_t)4095;
    return PRC_OK;
  }
  if (g->host_img) {
    run_all(g, [&](uint32_t r) { return prc_set_host_image(g->ctx[r], nullptr, 0); }, rcs);
    free(g->host_img);
    g->host_img = nullptr;
    g->host_cap = g->host_bytes = 0;
    const int32_t e = first_error(g, rcs, "prc_set_host_image(NULL)");
    if (e) return e;
  }
  if (!bytes) return PRC_OK;
  g->host_stride = (bytes + 4095) & ~(size_t)4095;
  const size_t cap = 2 * g->host_stride;
  void* p = nullptr;
  if (posix_memalign(&p, 4096, cap) != 0) { g->err = "out of host memory"; return PRC_ERR_CUDA; }
  memset(p, 0, cap);
  g->host_img = (uint8_t*)p;
  g->host_cap = cap;
  g->host_bytes = bytes;
  // rank 0 page-locks the image (portable: every context of the process sees it), the others find it registered
  int32_t e = prc_set_host_image(g->ctx[0], g->host_img, cap);
  if (e) { g->err = std::string("prc_set_host_image: rank 0: ") + prc_last_error(g->ctx[0]); return e; }
  for (uint32_t r = 1; r < world(g); r++) {
    e = prc_set_host_image(g->ctx[r], g->host_img, cap);
    if (e) { g->err = "prc_set_host_image: rank " + std::to_string(r) + ": " + prc_last_error(g->ctx[r]); return e; }
  }
  return PRC_OK;
}

bool same_signature(const prc_group* g, const prc_frame* fr) {
  if (!g->connected || g->sig_w != fr->width || g->sig_h != fr->height || g->sig_msaa != (fr->msaa > 1 ? fr->msaa : 1u) || g->sig_lights != fr->n_lights ||
      g->sig_flags != (fr->flags & PRC_FRAME_SHADOWMAP) || g->sig_cast.size() != fr->n_lights)
    return false;
  for (uint32_t i = 0; i < fr->n_lights; i++)
    if (g->sig_cast[i] != (fr->lights[i].cast_shadow ? 1 : 0)) return false;
  return true;
}

// export on every context, then connect every context to all of them
int32_t connect(prc_group* g, const prc_frame* fr) {
  int32_t e = disconnect(g);
  if (e) return e;
  const uint32_t n = world(g);
  const int msaa = fr->msaa > 1 ? (int)fr->msaa : 1;
  const int out_h = (int)fr->height / msaa;
  if (out_h < (int)n) { g->err = "a frame of " + std::to_string(out_h) + " rows cannot be cut into " + std::to_string(n) + " strips"; return PRC_ERR_INVALID; }
  std::vector<prc_peer_handle> handles(n);
  std::vector<int32_t> rcs;
  run_all(g, [&](uint32_t r) {
    prc_frame f = *fr;
    f.flags &= ~(uint32_t)(PRC_FRAME_ASYNC | PRC_FRAME_UNIFORMS_RESIDENT);
    f.row0 = 0; f.row1 = f.height;
    return prc_peer_export(g->ctx[r], &f, &handles[r]);
  }, rcs);
  if ((e = first_error(g, rcs, "prc_peer_export"))) return e;
  run_all(g, [&](uint32_t r) { return prc_peer_connect(g->ctx[r], r, n, handles.data()); }, rcs);
  if ((e = first_error(g, rcs, "prc_peer_connect"))) {
    std::vector<int32_t> rc2;
    run_all(g, [&](uint32_t r) { return prc_peer_disconnect(g->ctx[r]); }, rc2);
    return e;
  }
  g->connected = true;
  g->sig_w = fr->width; g->sig_h = fr->height; g->sig_msaa = (uint32_t)msaa; g->sig_lights = fr->n_lights;
  g->sig_flags = fr->flags & PRC_FRAME_SHADOWMAP;
  g->sig_cast.assign(fr->n_lights, 0);
  for (uint32_t i = 0; i < fr->n_lights; i++) g->sig_cast[i] = fr->lights[i].cast_shadow ? 1 : 0;
  g->bounds = equal_bounds(out_h, (int)n);
  apply_bounds(g, out_h, msaa);
  g->frames_since_connect = 0;
  g->pending = 0;
  return PRC_OK;
}

// After a synchronised frame with kernel timers: move the strip boundaries towards equal resolve + shading time per rank.
void rebalance(prc_group* g, int out_h, int msaa) {
  const uint32_t n = world(g);
  std::vector<double> cost(n, 0.0);
  for (uint32_t r = 0; r < n; r++) {
    prc_timings t;
    memset(&t, 0, sizeof(t));
    if (prc_get_timings(g->ctx[r], &t) != PRC_OK) return;
    cost[r] = (double)t.kernel_ms[PRC_K_RESOLVE] + (double)t.kernel_ms[PRC_K_SHADE];
  }
  g->bounds = balanced_bounds(g->bounds, cost, 0.7, std::max(1, 16 / msaa));
  apply_bounds(g, out_h, msaa);
}

// finishes the submitted frames on every rank; PRC_ERR_RETRY (states agreed) if they must be submitted again
int32_t sync_all(prc_group* g) {
  std::vector<int32_t> rcs;
  run_all(g, [&](uint32_t r) { return prc_sync(g->ctx[r]); }, rcs);
  g->pending = 0;
  const int32_t e = first_error(g, rcs, "prc_sync", PRC_ERR_RETRY);
  if (e) return e;
  bool retry = false;
  for (int32_t rc : rcs) retry = retry || rc == PRC_ERR_RETRY;
  if (!retry) return PRC_OK;
  // the ranks must agree on NaN mode / the tile path before the frames are submitted again (include/polyred_cuda.h prc_frame_state)
  uint32_t state = 0;
  for (prc_ctx* c : g->ctx) {
    uint32_t s = 0;
    prc_frame_state(c, &s);
    state |= s;
  }
  run_all(g, [&](uint32_t r) { return prc_set_frame_state(g->ctx[r], state); }, rcs);
  const int32_t e2 = first_error(g, rcs, "prc_set_frame_state");
  if (e2) return e2;
  g->err = "frames submitted back to back must be submitted again (a queue was grown, or the tile path / NaN mode was switched on)";
  return PRC_ERR_RETRY;
}

}  // namespace

extern "C" {

int32_t prc_group_open(const int32_t* devices, uint32_t n, prc_group** out) {
  if (!out) return PRC_ERR_INVALID;
  *out = nullptr;
  if (!devices || n == 0 || n > 16) return PRC_ERR_INVALID;
  prc_group* g = new prc_group();
  for (uint32_t r = 0; r < n; r++) {
    prc_ctx* c = nullptr;
    const int32_t e = prc_open(devices[r], &c);
    if (e != PRC_OK) {
      for (prc_ctx* o : g->ctx) prc_close(o);
      delete g;
      return e;
    }
    g->ctx.push_back(c);
    g->devices.push_back(devices[r]);
  }
  const char* b = getenv("PRC_GROUP_BALANCE");
  g->balance = !(b && atoi(b) == 0);
  if (const char* sg = getenv("PRC_GROUP_STAGGER_MS")) g->stagger_ms = std::max(0, std::min(1000, atoi(sg)));
  if (n > 1) {
    // all Worker objects first, then the threads: a worker reads g->workers[rank] as soon as it starts, and a push_back for the
    // next worker could reallocate the vector under it (found by ThreadSanitizer, tests/test_group_host_native.py)
    for (uint32_t r = 0; r < n; r++) g->workers.push_back(new Worker());
    for (uint32_t r = 0; r < n; r++) g->workers[r]->th = std::thread(worker_main, g, r);
  }
  *out = g;
  return PRC_OK;
}

int32_t prc_group_close(prc_group* g) {
  if (!g) return PRC_ERR_INVALID;
  if (g->connected) disconnect(g);
  if (g->host_img) set_host_image(g, 0);
  for (Worker* w : g->workers) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->quit.store(true);
      w->posted.fetch_add(1, std::memory_order_release);
    }
    w->cv.notify_one();
    w->th.join();
    delete w;
  }
  for (prc_ctx* c : g->ctx) prc_close(c);
  delete g;
  return PRC_OK;
}

const char* prc_group_last_error(prc_group* g) { return g ? g->err.c_str() : "null group"; }

uint32_t prc_group_size(prc_group* g) { return g ? world(g) : 0u; }

int32_t prc_group_ctx(prc_group* g, uint32_t rank, prc_ctx** out) {
  if (!g || !out || rank >= world(g)) return PRC_ERR_INVALID;
  *out = g->ctx[rank];
  return PRC_OK;
}

int32_t prc_group_scene_upload(prc_group* g, const prc_scene* scene) {
  if (!g) return PRC_ERR_INVALID;
  if (g->pending) { const int32_t e = sync_all(g); if (e != PRC_OK && e != PRC_ERR_RETRY) return e; }
  std::vector<int32_t> rcs;
  run_all(g, [&](uint32_t r) { return prc_scene_upload(g->ctx[r], scene); }, rcs);
  return first_error(g, rcs, "prc_scene_upload");
}

int32_t prc_group_shadow_reset(prc_group* g) {
  if (!g) return PRC_ERR_INVALID;
  // no frame may be in flight on any rank while the maps are zeroed (a peer's merged texels would be lost)
  if (g->pending) { const int32_t e = sync_all(g); if (e != PRC_OK && e != PRC_ERR_RETRY) return e; }
  std::vector<int32_t> rcs;
  run_all(g, [&](uint32_t r) { return prc_shadow_reset(g->ctx[r]); }, rcs);
  return first_error(g, rcs, "prc_shadow_reset");
}

int32_t prc_group_sync(prc_group* g) {
  if (!g) return PRC_ERR_INVALID;
  return sync_all(g);
}

int32_t prc_group_render(prc_group* g, const prc_frame* fr, uint8_t* rgba_out) {
  if (!g) return PRC_ERR_INVALID;
  if (!fr || fr->abi_version != PRC_ABI_VERSION) { g->err = "prc_frame: bad abi_version"; return PRC_ERR_INVALID; }
  if (world(g) == 1) {
    // a group of one device is the plain single-context frame (no private buffers, no merge; every prc_render option applies)
    const int32_t e1 = prc_render(g->ctx[0], fr, rgba_out);
    if (e1 != PRC_OK) g->err = std::string("prc_render: ") + prc_last_error(g->ctx[0]);
    else if (fr->flags & PRC_FRAME_ASYNC) g->pending++;
    g->row0.assign(1, fr->row0);
    g->row1.assign(1, fr->row1);
    return e1;
  }
  if (fr->row0 != 0 || fr->row1 != fr->height) { g->err = "prc_group_render: frame.row0/row1 must be 0/height (the group cuts the strips itself)"; return PRC_ERR_INVALID; }
  if (fr->flags & (PRC_FRAME_KEEP_GBUFFER | PRC_FRAME_SHADOW_RESET)) {
    g->err = "prc_group_render: PRC_FRAME_KEEP_GBUFFER / PRC_FRAME_SHADOW_RESET are not supported (prc_group_shadow_reset zeroes the maps of every rank)";
    return PRC_ERR_UNSUPPORTED;
  }
  if (fr->n_lights && !fr->lights) { g->err = "prc_frame.lights is NULL"; return PRC_ERR_INVALID; }
  const bool on_device = (fr->flags & PRC_FRAME_NO_READBACK) != 0, async = (fr->flags & PRC_FRAME_ASYNC) != 0;
  if (async && !on_device) { g->err = "PRC_FRAME_ASYNC needs PRC_FRAME_NO_READBACK"; return PRC_ERR_INVALID; }
  const uint32_t n = world(g);
  const int msaa = fr->msaa > 1 ? (int)fr->msaa : 1;
  if (msaa > 8 || fr->width % msaa || fr->height % msaa) { g->err = "msaa must be 1..8 and divide the frame size"; return PRC_ERR_INVALID; }
  if (msaa > 1 && on_device) { g->err = "prc_group_render: MSAA frames leave through the host image (no PRC_FRAME_NO_READBACK)"; return PRC_ERR_UNSUPPORTED; }
  const int out_w = (int)fr->width / msaa, out_h = (int)fr->height / msaa;
  int32_t e;
  if (!same_signature(g, fr)) {
    if (g->pending && (e = sync_all(g)) != PRC_OK && e != PRC_ERR_RETRY) return e;
    if ((e = connect(g, fr)) != PRC_OK) return e;
  }
  if (!on_device) {
    if ((e = set_host_image(g, (size_t)out_w * out_h * 4)) != PRC_OK) return e;
    g->host_cur ^= 1;
    for (uint32_t r = 0; r < n; r++)
      if ((e = prc_set_host_image_offset(g->ctx[r], (uint64_t)g->host_cur * g->host_stride)) != PRC_OK) {
        g->err = std::string("prc_set_host_image_offset: ") + prc_last_error(g->ctx[r]);
        return e;
      }
  }
  // every 64th synchronised frame (and the first ones after connecting) runs with the per-class event brackets and moves the strips
  const bool measure = g->balance && n > 1 && !async && (g->frames_since_connect < 10 || g->frames_since_connect % 64 == 0);
  std::vector<int32_t> rcs;
  for (int attempt = 0; attempt < 4; attempt++) {
    run_all(g, [&](uint32_t r) {
      prc_frame f = *fr;
      f.flags &= ~(uint32_t)PRC_FRAME_ASYNC;
      if (!measure && n > 1) f.flags |= PRC_FRAME_NO_KERNEL_TIMERS;
      if (async) f.flags |= PRC_FRAME_IMAGE_AT_SYNC;  // frames submitted back to back: rank 0 does not stop for the strips of each one (prc_group_sync does)
      f.row0 = g->row0[r];
      f.row1 = g->row1[r];
      if (g->stagger_ms && r) std::this_thread::sleep_for(std::chrono::milliseconds((long)g->stagger_ms * r));
      int32_t rc = prc_render_peer(g->ctx[r], &f, n, g->row0.data(), g->row1.data(), on_device ? 1u : 0u);
      if (rc != PRC_OK || async) return rc;
      return prc_sync(g->ctx[r]);
    }, rcs);
    if ((e = first_error(g, rcs, "prc_render_peer", PRC_ERR_RETRY))) {
      // a rank that failed to submit leaves its peers waiting for its signals: finish (their waits time out) and drop the connection
      std::vector<int32_t> rc2;
      run_all(g, [&](uint32_t r) { return prc_sync(g->ctx[r]); }, rc2);
      const std::string keep = g->err;
      disconnect(g);
      g->err = keep;
      return e;
    }
    if (async) { g->pending++; g->frames_since_connect++; return PRC_OK; }
    bool retry = false;
    for (int32_t rc : rcs) retry = retry || rc == PRC_ERR_RETRY;
    if (!retry) {
      g->frames_since_connect++;
      if (measure) rebalance(g, out_h, msaa);
      if (!on_device && rgba_out) memcpy(rgba_out, g->host_img + (size_t)g->host_cur * g->host_stride, (size_t)out_w * out_h * 4);
      return PRC_OK;
    }
    uint32_t state = 0;
    for (prc_ctx* c : g->ctx) {
      uint32_t s = 0;
      prc_frame_state(c, &s);
      state |= s;
    }
    run_all(g, [&](uint32_t r) { return prc_set_frame_state(g->ctx[r], state); }, rcs);
    if ((e = first_error(g, rcs, "prc_set_frame_state"))) return e;
  }
  g->err = "a queue kept overflowing";
  return PRC_ERR_UNSUPPORTED;
}

int32_t prc_group_host_image(prc_group* g, uint64_t* host_ptr, uint64_t* bytes) {
  if (g && host_ptr && bytes && world(g) == 1) return prc_host_image(g->ctx[0], host_ptr, bytes);
  if (!g || !host_ptr || !bytes || !g->host_img) return PRC_ERR_INVALID;
  *host_ptr = (uint64_t)(uintptr_t)(g->host_img + (size_t)g->host_cur * g->host_stride);
  *bytes = (uint64_t)g->host_bytes;
  return PRC_OK;
}

int32_t prc_group_strips(prc_group* g, uint32_t* row0, uint32_t* row1) {
  if (!g || !row0 || !row1 || g->row0.size() != g->ctx.size()) return PRC_ERR_INVALID;
  for (uint32_t r = 0; r < world(g); r++) { row0[r] = g->row0[r]; row1[r] = g->row1[r]; }
  return PRC_OK;
}

int32_t prc_group_render_views(prc_group* g, uint32_t n_views, const prc_frame* frames, uint8_t* const* rgba_out) {
  if (!g) return PRC_ERR_INVALID;
  if (n_views == 0) return PRC_OK;
  if (!frames) { g->err = "prc_group_render_views: frames is NULL"; return PRC_ERR_INVALID; }
  int32_t e;
  if (g->pending && (e = sync_all(g)) != PRC_OK && e != PRC_ERR_RETRY) return e;
  // views need no exchange: every context renders whole frames into its own buffers
  if ((e = disconnect(g)) != PRC_OK) return e;
  const uint32_t n = world(g);
  std::vector<std::vector<prc_frame>> mine(n);
  std::vector<std::vector<uint8_t*>> outs(n);
  for (uint32_t v = 0; v < n_views; v++) {
    mine[v % n].push_back(frames[v]);
    outs[v % n].push_back(rgba_out ? rgba_out[v] : nullptr);
  }
  std::vector<int32_t> rcs;
  run_all(g, [&](uint32_t r) {
    if (mine[r].empty()) return (int32_t)PRC_OK;
    return prc_render_batch(g->ctx[r], (uint32_t)mine[r].size(), mine[r].data(), outs[r].data());
  }, rcs);
  return first_error(g, rcs, "prc_render_batch");
}

}  // extern "C"
