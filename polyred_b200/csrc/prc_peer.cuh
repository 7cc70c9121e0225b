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
#define PRC_PEER_TIMEOUT_NS 4000000000ull  // default; PRC_PEER_TIMEOUT_MS in the environment of prc_open overrides it (a host that may stall longer between submits)

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
  unsigned char* dirty;          // [(1 + n_sh) H][nseg][PRC_DIRTY_STRIDE] segment flags of this rank's private buffers (Counters.dirty), cleared here
  int nseg;                      // 256-pixel segments per row = ceil(W / PRC_DIRTY_SEG) of this rank's private buffers (Counters.dirty), cleared here
  float* sh_mine;                // private stacked shadow maps [n_sh][H][W], zeroed where read (clear-on-read)
  unsigned long long* k_mine;    // private key plane [H][W], zeroed where read
  unsigned long long k_off;      // offset of this frame's parity plane inside mkeys[], in keys
  int rr0[PRC_PEER_MAX], rr1[PRC_PEER_MAX], ax1[PRC_PEER_MAX];  // rank p resolves rows [rr0, rr1) and [0, ax1) (+ pixel (0,0), always)
  // NaN mode: first-fragment plane, merged with atomicMin (nullptr otherwise; reset to ~0 where read); f_off = offset of the merged plane
  unsigned long long* f_mine;
  unsigned long long f_off;
  long long seg0, seg1;          // segments this launch scans: [0, H nseg) = the key plane, [H nseg, (1 + n_sh) H nseg) = the shadow planes
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
                            unsigned int* __restrict__ timeouts, unsigned long long timeout_ns) {
  const uint32_t src = threadIdx.x;
  if (src >= world || src == self || !((mask >> src) & 1u)) return;
  const uint32_t* w = my_signals + kind * PRC_PEER_MAX + src;
  const unsigned long long t0 = global_ns();
  unsigned int spins = 0;
  while ((int32_t)(ld_acquire_sys(w) - epoch) < 0) {  // epochs are compared modulo 2^32
    if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) {
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

// k_peer_signal followed by k_peer_wait as ONE launch (every launch on a frame's critical path costs 2-4 us of a ~0.2 ms 8-GPU
// frame): thread t publishes `sig_epoch` to rank t, then waits for rank t's word. The threads are independent, so this is exactly
// the two kernels back to back.
__global__ void k_peer_signal_wait(PeerTable P, uint32_t sig_kind, uint32_t sig_epoch, uint32_t sig_mask, uint32_t wait_kind, uint32_t wait_epoch,
                                   uint32_t wait_mask, unsigned int* __restrict__ timeouts, unsigned long long timeout_ns) {
  const uint32_t t = threadIdx.x;
  const bool peer = t < P.world && t != P.self;
  if (peer && ((sig_mask >> t) & 1u)) {
    __threadfence_system();
    st_release_sys(P.signals[t] + sig_kind * PRC_PEER_MAX + P.self, sig_epoch);
  }
  __syncwarp();  // (one warp) every signal is on its way before any lane starts to spin
  if (!peer || !((wait_mask >> t) & 1u)) return;
  const uint32_t* w = P.signals[P.self] + wait_kind * PRC_PEER_MAX + t;
  const unsigned long long t0 = global_ns();
  unsigned int spins = 0;
  while ((int32_t)(ld_acquire_sys(w) - wait_epoch) < 0) {
    if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) {
      atomicAdd(timeouts, 1u);
      return;
    }
    __nanosleep(64);
  }
}

// Merges what this rank rasterised into its peers (and into its own merged buffers) — see the file header. Shadow depths are
// positive floats, which order like their int bits; depth 0 / key 0 is "nothing stored" and is skipped. The raster kernels flag
// every 256-pixel SEGMENT of a row they write (on C3 the scene covers a quarter of the screen), only flagged segments are read,
// and what is read is reset, so the private buffers are empty again for the next frame without a clearing pass. The reductions
// are fire-and-forget (RED over NVLink). One warp per batch of 32 segments: the lanes read 32 flags at once, then the warp walks
// the flagged segments with all loads of a segment issued before the first reduction (round 2: the first version walked whole
// rows with one dependent load per iteration and took 0.048 ms per frame whatever the number of ranks).
__device__ __forceinline__ void push_shadow_texel(const PeerTable& P, size_t i, float v, float merged, int full) {
  // A depth that does not exceed what this rank's own merged map holds is not sent: every value in a merged map arrived by a
  // push that goes to ALL ranks (and completes before its sender's signal, which every rank awaits before shading), so the peers
  // hold — or are about to hold — at least that value. The maps are persistent and only grow (render/shadow.go:221-228), so for
  // a scene that does not move the exchange shrinks to the texels that changed; PRC_PEER_FULL_PUSH=1 sends every non-empty
  // texel every frame (`full`).
  if (v == 0.0f || (!full && !(v > merged))) return;
  for (uint32_t p = 0; p < P.world; p++) atomicMax(reinterpret_cast<int*>(P.shadow[p]) + i, __float_as_int(v));
}
__device__ __forceinline__ void push_key(const PeerTable& P, const PushJob& J, size_t i, unsigned long long k, unsigned long long f, uint32_t dst) {
  if (!k && f == ~0ull) return;
  const uint32_t d = (i == 0) ? ((P.world >= 32u) ? 0xFFFFFFFFu : ((1u << P.world) - 1u)) : dst;  // pixel (0,0): every rank
  for (uint32_t p = 0; p < P.world; p++) {
    if (!((d >> p) & 1u)) continue;
    if (k) atomicMax(P.mkeys[p] + J.k_off + i, k);
    if (f != ~0ull) atomicMin(P.mkeys[p] + J.f_off + i, f);
  }
}
// one flagged segment, handled by one warp: every load of the segment is issued before the first reduction
__device__ __forceinline__ void push_segment(const PeerTable& P, const PushJob& J, long long sg, int lane, bool vec) {
  const int n_seg_row = J.nseg;
  const int row = (int)(sg / n_seg_row), x0 = (int)(sg - (long long)row * n_seg_row) * PRC_DIRTY_SEG;
  const int nx = min(PRC_DIRTY_SEG, J.W - x0);
  if (row >= J.H) {
    // ---- 256 shadow texels -> every rank's merged maps
    const size_t base = (size_t)(row - J.H) * J.W + x0;
    const float* __restrict__ merged = P.shadow[P.self] + base;
    float* mine = J.sh_mine + base;
    if (vec && nx == PRC_DIRTY_SEG) {
      const float4 v0 = reinterpret_cast<const float4*>(mine)[lane], v1 = reinterpret_cast<const float4*>(mine)[lane + 32];
      const float4 m0 = reinterpret_cast<const float4*>(merged)[lane], m1 = reinterpret_cast<const float4*>(merged)[lane + 32];
      if (v0.x != 0.0f || v0.y != 0.0f || v0.z != 0.0f || v0.w != 0.0f) reinterpret_cast<float4*>(mine)[lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (v1.x != 0.0f || v1.y != 0.0f || v1.z != 0.0f || v1.w != 0.0f) reinterpret_cast<float4*>(mine)[lane + 32] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      const size_t i0 = base + (size_t)lane * 4, i1 = i0 + 128;
      push_shadow_texel(P, i0, v0.x, m0.x, J.full); push_shadow_texel(P, i0 + 1, v0.y, m0.y, J.full);
      push_shadow_texel(P, i0 + 2, v0.z, m0.z, J.full); push_shadow_texel(P, i0 + 3, v0.w, m0.w, J.full);
      push_shadow_texel(P, i1, v1.x, m1.x, J.full); push_shadow_texel(P, i1 + 1, v1.y, m1.y, J.full);
      push_shadow_texel(P, i1 + 2, v1.z, m1.z, J.full); push_shadow_texel(P, i1 + 3, v1.w, m1.w, J.full);
    } else {
      for (int x = lane; x < nx; x += 32) {
        const float v = mine[x];
        if (v == 0.0f) continue;
        mine[x] = 0.0f;
        push_shadow_texel(P, base + x, v, merged[x], J.full);
      }
    }
  } else {
    // ---- 256 visibility keys -> the ranks that resolve this row (pixel (0,0): every rank)
    const int y = row;
    uint32_t dst = 0;
    for (uint32_t p = 0; p < P.world; p++)
      if ((y >= J.rr0[p] && y < J.rr1[p]) || y < J.ax1[p]) dst |= 1u << p;
    const size_t base = (size_t)y * J.W + x0;
    unsigned long long* mine = J.k_mine + base;
    if (vec && nx == PRC_DIRTY_SEG && !J.f_mine) {
      ulonglong2 k[4];
#pragma unroll
      for (int j = 0; j < 4; j++) k[j] = reinterpret_cast<const ulonglong2*>(mine)[lane + 32 * j];
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (k[j].x | k[j].y) reinterpret_cast<ulonglong2*>(mine)[lane + 32 * j] = make_ulonglong2(0ull, 0ull);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const size_t i = base + (size_t)(lane + 32 * j) * 2;
        push_key(P, J, i, k[j].x, ~0ull, dst);
        push_key(P, J, i + 1, k[j].y, ~0ull, dst);
      }
    } else {
      for (int x = lane; x < nx; x += 32) {
        const unsigned long long k = mine[x];
        const unsigned long long f = J.f_mine ? J.f_mine[base + x] : ~0ull;
        if (!k && f == ~0ull) continue;
        if (k) mine[x] = 0ull;
        if (f != ~0ull) J.f_mine[base + x] = ~0ull;
        push_key(P, J, base + x, k, f, dst);
      }
    }
  }
}
// A CTA takes batches of PRC_PUSH_BATCH segment flags: 64 threads read (and clear) one flag each and compact the flagged segments
// into shared memory, then the CTA's 8 warps share that list. (First segment version: one warp per 32 flags walking its flagged
// segments alone — up to 32 dependent round trips to memory per warp, 0.077 ms at 2 GPUs; whole rows before that: 0.048 ms.)
#define PRC_PUSH_BATCH 64
__global__ void __launch_bounds__(256) k_peer_push(const __grid_constant__ PeerTable P, const __grid_constant__ PushJob J) {
  __shared__ unsigned int list[PRC_PUSH_BATCH];
  __shared__ unsigned int n_list;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_seg = J.seg1;  // this launch merges the segments [seg0, seg1): the key plane alone, or every plane
  const bool vec = (J.W & 3) == 0;  // 16-byte loads need rows of whole float4 / ulonglong2
  for (long long s0 = J.seg0 + (long long)blockIdx.x * PRC_PUSH_BATCH; s0 < n_seg; s0 += (long long)gridDim.x * PRC_PUSH_BATCH) {
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    if (threadIdx.x < PRC_PUSH_BATCH) {
      const long long sl = s0 + threadIdx.x;
      if (sl < n_seg) {
        unsigned char* flag = J.dirty + (size_t)sl * PRC_DIRTY_STRIDE;
        if (*flag != 0) {
          *flag = 0;
          list[atomicAdd(&n_list, 1u)] = threadIdx.x;
        }
      }
    }
    __syncthreads();
    const unsigned int n = n_list;
    for (unsigned int i = warp; i < n; i += 8) push_segment(P, J, s0 + list[i], lane, vec);
    __syncthreads();  // the list is rewritten by the next batch
  }
}

}  // namespace prc
