// prc_peer.cuh — the multi-GPU exchange of one frame over NVLink peer memory (no reference counterpart: polyred has no
// multi-GPU path; the parity oracle of this file is "the same frame as one GPU").
//
// Every rank (one process per GPU) maps its peers' shadow buffer, image buffer and signal words (CUDA IPC) and the frame
// stays on the device end to end:
//   * shadow rows a rank rasterised are PUSHED into every peer's copy of the stacked maps by k_shadow_push — only texels
//     that hold a depth (a shadow map is mostly zeros: 1.6-5 % written on C3), instead of an all-gather of every byte;
//   * a rank's shaded image strip is copied into the root's image (copy engine, peer-to-peer);
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
  PRC_SIG_SHADOW = 0,    // source's shadow rows of this epoch are in my maps
  PRC_SIG_SHADED = 1,    // source no longer reads its shadow maps of this epoch (its shading is done)
  PRC_SIG_IMAGE = 2,     // source's image strip of this epoch is in my image
  PRC_SIG_IMAGE_FREE = 3,  // source (an image consumer) is done with the image of this epoch
  PRC_SIG_KINDS = 4
};

struct PeerTable {
  float* shadow[PRC_PEER_MAX];     // stacked shadow maps of every rank (own entry = local pointer)
  uint32_t* signals[PRC_PEER_MAX];  // [PRC_SIG_KINDS][PRC_PEER_MAX] words of every rank
  uint32_t world, self;
};

struct PushUnits {
  // up to 32 (offset, count) float ranges of the stacked maps owned by this rank (one per shadow unit)
  unsigned long long off[32];
  unsigned long long cnt[32];
  uint32_t n;
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

// Pushes the non-empty texels of the rows this rank owns into every peer's maps. The owner is the only rank that
// rasterises those rows and its values only grow (shadowDepthTest keeps the maximum, render/shadow.go:221-228), so a peer's
// texel is always an older value of the owner's: a plain store of the current value is the all-gather's result.
// Depth 0 is "nothing stored" (maps start at 0 and only z > 0 is ever kept), so zero texels are skipped.
template <int VEC>
__global__ void __launch_bounds__(256) k_shadow_push(PeerTable P, PushUnits U) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const float* __restrict__ mine = P.shadow[P.self];
  for (uint32_t u = 0; u < U.n; u++) {
    const unsigned long long base = U.off[u], n = U.cnt[u];
    if (VEC == 4) {
      const float4* __restrict__ src = reinterpret_cast<const float4*>(mine + base);
      for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
        const float4 v = __ldg(src + i);
        if (v.x == 0.0f && v.y == 0.0f && v.z == 0.0f && v.w == 0.0f) continue;
        for (uint32_t p = 0; p < P.world; p++)
          if (p != P.self) reinterpret_cast<float4*>(P.shadow[p] + base)[i] = v;
      }
    } else {
      for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __ldg(mine + base + i);
        if (v == 0.0f) continue;
        for (uint32_t p = 0; p < P.world; p++)
          if (p != P.self) P.shadow[p][base + i] = v;
      }
    }
  }
}

}  // namespace prc
