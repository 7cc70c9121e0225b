// prc_peer.cuh — the multi-GPU exchange of one frame over NVLink peer memory (no reference counterpart: polyred has no
// multi-GPU path; the parity oracle of this file is "the same frame as one GPU").
//
// Partition ("sort-last" for the geometry, screen strips for the shading; scene replicated on every GPU):
//   * every rank runs the geometry + raster passes — the camera pass and the fused sweep over all shadow-casting lights, the same
//     kernels and launch shape as on one GPU — over ITS SHARE OF THE TRIANGLES (16-chunk blocks dealt round-robin), into private
//     full-frame visibility keys and private shadow maps. (Round 1 cut the raster passes by screen rows instead: on C3 half of
//     all triangles project into one eighth of the rows, and a rank's pass cost 0.11 ms of the 0.175 ms one GPU needs.)
//   * k_peer_push then merges what the rank produced into its peers over NVLink with atomicMax: non-empty shadow texels into
//     EVERY rank's maps (every pixel can look up any texel), non-empty visibility keys into the rank(s) that shade that row.
//     Depth maxima / key maxima do not depend on who contributes what, so the merged buffers equal the 1-GPU buffers bit for bit.
//   * every rank shades its strip of rows from the merged keys and maps, and copies the strip into the root's image.
//   * ordering between ranks is by monotone epoch words written with st.release.sys into the PEER's memory
//     (k_peer_signal) and awaited with ld.acquire.sys by a one-warp kernel on the consumer's stream (k_peer_wait), so the
//     host never waits inside a frame and frames can be submitted back to back.
// A wait gives up after PRC_PEER_TIMEOUT_NS and counts the event (prc_sync then reports PRC_ERR_PEER): a missing peer
// turns into an error, not into a hung GPU.
#pragma once
#include <cstdint>

namespace prc {

#define PRC_PEER_MAX 16              // ranks in one exchange group (one NVSwitch domain)
#define PRC_PEER_TIMEOUT_NS 4000000000ull

// signal words of one rank, written by its peers: word [kind][source rank] holds the last epoch the source finished
enum PeerSignal : uint32_t {
  PRC_SIG_SHADOW = 0,    // source's shadow texels and visibility keys of this epoch are merged into my buffers
  PRC_SIG_SHADED = 1,    // source no longer reads its shadow maps of this epoch (its shading is done)
  PRC_SIG_IMAGE = 2,     // source's image strip of this epoch is in my image
  PRC_SIG_IMAGE_FREE = 3,  // source (an image consumer) is done with the image of this epoch
  PRC_SIG_KINDS = 4
};

struct PeerTable {
  float* shadow[PRC_PEER_MAX];     // merged stacked shadow maps of every rank (own entry = local pointer)
  uint32_t* signals[PRC_PEER_MAX];  // [PRC_SIG_KINDS][PRC_PEER_MAX] words of every rank
  unsigned long long* mkeys[PRC_PEER_MAX];  // merged visibility keys of every rank: 2 frame parities x [H][W] (x2 with the NaN-mode plane)
  uint32_t world, self;
};

// what one k_peer_push launch merges into the peers
struct PushJob {
  int W, H;
  uint32_t n_sh;                 // casting lights: rows [H, (1 + n_sh) H) of the row space are shadow rows (light k: [(1+k) H, (2+k) H))
  unsigned char* dirty;          // [(1 + n_sh) H][PRC_DIRTY_STRIDE] row flags (one per 32-byte sector) of this rank's private buffers (Counters.dirty), cleared here
  float* sh_mine;                // private stacked shadow maps [n_sh][H][W], zeroed where read (clear-on-read)
  unsigned long long* k_mine;    // private key plane [H][W], zeroed where read
  unsigned long long k_off;      // offset of this frame's parity plane inside mkeys[], in keys
  int rr0[PRC_PEER_MAX], rr1[PRC_PEER_MAX], ax1[PRC_PEER_MAX];  // rank p resolves rows [rr0, rr1) and [0, ax1) (+ pixel (0,0), always)
  // NaN mode: first-fragment plane, merged with atomicMin (nullptr otherwise; reset to ~0 where read); f_off = offset of the merged plane
  unsigned long long* f_mine;
  unsigned long long f_off;
  int full;                      // PRC_PEER_FULL_PUSH: send every non-empty shadow texel, also those the peers provably hold
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Waits until signal word [kind][src] of THIS rank has reached `epoch`, for every src != self in `mask`.
// One thread per source rank; <<<1, PRC_PEER_MAX>>>.
__global__ void k_peer_wait(const uint32_t* __restrict__ my_signals, uint32_t world, uint32_t self, uint32_t kind, uint32_t epoch, uint32_t mask,
                            unsigned int* __restrict__ timeouts) {
  const uint32_t src = threadIdx.x;
  if (src >= world || src == self || !((mask >> src) & 1u)) return;
  const uint32_t* w = my_signals + kind * PRC_PEER_MAX + src;
  const unsigned long long t0 = global_ns();
  unsigned int spins = 0;
  while ((int32_t)(ld_acquire_sys(w) - epoch) < 0) {  // epochs are compared modulo 2^32
    if ((++spins & 1023u) == 0 && global_ns() - t0 > PRC_PEER_TIMEOUT_NS) {
      atomicAdd(timeouts, 1u);
      return;
    }
    __nanosleep(64);
  }
}

// Publishes `epoch` into word [kind][self] of every rank in `mask` (never into this rank's own words).
// Stream order puts this kernel after the writes it announces; the release at system scope is cumulative over them.
__global__ void k_peer_signal(PeerTable P, uint32_t kind, uint32_t epoch, uint32_t mask) {
  const uint32_t dst = threadIdx.x;
  if (dst >= P.world || dst == P.self || !((mask >> dst) & 1u)) return;
  __threadfence_system();
  st_release_sys(P.signals[dst] + kind * PRC_PEER_MAX + P.self, epoch);
}

// Merges what this rank rasterised into its peers (and into its own merged buffers) — see the file header. Shadow depths are
// positive floats, which order like their int bits; depth 0 / key 0 is "nothing stored" and is skipped. Only rows flagged by the
// raster kernels are read (on C3 the scene covers a quarter of the screen rows and of each light's rows), and what is read
// is reset, so the private buffers are empty again for the next frame without a clearing pass. The reductions are
// fire-and-forget (RED over NVLink). One CTA per flagged row at a time.
__global__ void __launch_bounds__(256) k_peer_push(const __grid_constant__ PeerTable P, const __grid_constant__ PushJob J) {
  const int n_rows = (int)(1u + J.n_sh) * J.H;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    if (!J.dirty[(size_t)row * PRC_DIRTY_STRIDE]) continue;  // (uniform over the CTA)
    __syncthreads();              // every thread has read the flag
    if (threadIdx.x == 0) J.dirty[(size_t)row * PRC_DIRTY_STRIDE] = 0;
    if (row >= J.H) {
      // ---- a shadow row -> every rank's merged maps. A depth that does not exceed what this rank's own merged map holds is
      // not sent: every value in a merged map arrived by a push that goes to ALL ranks (and completes before its sender's
      // signal, which every rank awaits before shading), so the peers hold — or are about to hold — at least that value.
      // The maps are persistent and only grow (render/shadow.go:221-228), so for a scene that does not move the exchange
      // shrinks to the texels that changed; PRC_PEER_FULL_PUSH=1 sends every non-empty texel every frame (`full`).
      const size_t base = (size_t)(row - J.H) * J.W;
      const float* __restrict__ merged = P.shadow[P.self] + base;
      for (int x = threadIdx.x; x < J.W; x += blockDim.x) {
        const float v = J.sh_mine[base + x];
        if (v == 0.0f) continue;
        J.sh_mine[base + x] = 0.0f;
        if (!J.full && !(v > merged[x])) continue;
        for (uint32_t p = 0; p < P.world; p++) atomicMax(reinterpret_cast<int*>(P.shadow[p]) + base + x, __float_as_int(v));
      }
    } else {
      // ---- a row of visibility keys -> the ranks that resolve it (pixel (0,0): every rank)
      const int y = row;
      uint32_t dst = 0;
      for (uint32_t p = 0; p < P.world; p++)
        if ((y >= J.rr0[p] && y < J.rr1[p]) || y < J.ax1[p]) dst |= 1u << p;
      const size_t base = (size_t)y * J.W;
      for (int x = threadIdx.x; x < J.W; x += blockDim.x) {
        const unsigned long long k = J.k_mine[base + x];
        const unsigned long long f = J.f_mine ? J.f_mine[base + x] : ~0ull;
        if (!k && f == ~0ull) continue;
        if (k) J.k_mine[base + x] = 0ull;
        if (f != ~0ull) J.f_mine[base + x] = ~0ull;
        const uint32_t d = (base + x == 0) ? ((1u << P.world) - 1u) : dst;
        for (uint32_t p = 0; p < P.world; p++) {
          if (!((d >> p) & 1u)) continue;
          if (k) atomicMax(P.mkeys[p] + J.k_off + base + x, k);
          if (f != ~0ull) atomicMin(P.mkeys[p] + J.f_off + base + x, f);
        }
      }
    }
  }
}

}  // namespace prc
