"""One frame over N GPUs, one process per GPU (torch.distributed / NCCL over NVLink for the plumbing).

The reference has no multi-GPU path. Partition (north_star): the scene is replicated, the
screen is cut into N equal horizontal strips, the Ls shadow maps are cut into (light, row
range) shards, one contiguous chunk of the stacked rows per rank. Per frame:
  1. every rank rasterises its chunk of the stacked shadow maps                (prc_render_shadows)
  2. ONE in-place NCCL all-gather over the library's contiguous shadow buffer  (prc_device_shadow_all), asynchronous
  3. meanwhile: camera geometry + raster + resolve for this rank's strip       (prc_render_forward)
  4. shading once the maps have arrived                                        (prc_render_deferred)
  5. ONE in-place all-gather over the image buffer assembles the frame         (prc_device_image; equal strips)
The result is bit-identical to the 1-GPU frame: atomicMax keys / depth maxima do not depend on who
rasterised what, and every rank also rasterises pixel (0,0) (bug-list 3) and the AO halo rows.

PeerFrames (below) drives prc_render_peer, the default multi-GPU path: no collective and no host wait inside the
frame. The ranks map each other's buffers (CUDA IPC over NVLink); every rank rasterises ITS SHARE OF THE TRIANGLES
(camera pass + all shadow lights, full frame) into private buffers, the library merges them into the peers
with atomicMax over NVLink (shadow texels into every rank, visibility keys into the rank that shades the row),
every rank shades its strip and copies it into rank 0's image (csrc/prc_peer.cuh). torch.distributed only carries
the handles and the retry vote.
"""
from __future__ import annotations

import numpy as np

from . import partition


import os as _os

# PRC_PEER_NO_LATE_IMAGE=1: gathered device frames keep the wait for the peers' strips inside the consumer's stream (A/B switch for
# PRC_FRAME_IMAGE_AT_SYNC, include/polyred_cuda.h)
_NO_LATE_IMAGE = _os.environ.get("PRC_PEER_NO_LATE_IMAGE", "0") not in ("", "0")


def _cai(ptr, nbytes):
    class _V:
        pass
    v = _V()
    v.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return v


class DistributedFrame:
    def __init__(self, renderer, rank: int, world: int, device: int):
        import torch
        self.torch = torch
        self.r, self.be = renderer, renderer._backend
        self.rank, self.world, self.device = rank, world, torch.device("cuda", device)
        c = renderer.cfg
        self.w, self.h = c.Width, c.Height
        sources, _ = c.Scene.Lights()
        self.cast = [i for i, l in enumerate(sources) if l.cast_shadow] if c.ShadowMap else []
        # validated here, on EVERY rank alike (same h, world and options everywhere): a rank that raised alone later — the
        # library rejects an empty strip — would leave the others waiting in the all-gather
        from ._lib import PolyredCudaError
        from . import _abi as A
        if max(1, int(getattr(c, "MSAA", 1))) != 1:
            raise PolyredCudaError(A.PRC_ERR_UNSUPPORTED, "DistributedFrame: MSAA frames are cut into strips by PeerFrames / prc_group_render only")
        if world < 1 or partition.strips(self.h, world)[1][-1][0] >= partition.strips(self.h, world)[1][-1][1]:
            raise PolyredCudaError(A.PRC_ERR_INVALID, f"DistributedFrame: {self.h} rows cannot be cut into {world} equal non-empty strips")
        if self.cast and len(self.cast) * self.h < world:
            raise PolyredCudaError(A.PRC_ERR_INVALID, f"DistributedFrame: {len(self.cast)} shadow maps of {self.h} rows cannot be cut into {world} shards")
        self.img_chunk, self.rows = partition.strips(self.h, world)
        self.chunk, _ = partition.shadow_chunks(self.h, world, len(self.cast))
        self.units = partition.shadow_units(self.h, world, self.cast)
        self.stream = torch.cuda.ExternalStream(self.be.stream(), device=self.device)
        self._views = {}
        self._host = self._host_np = None
        import os
        self.overlap = os.environ.get("PRC_MGPU_OVERLAP", "1") != "0"

    def prepare(self, fd):
        fd.struct.row0, fd.struct.row1 = self.rows[self.rank]
        return fd

    def _view(self, ptr, nbytes):
        key = (ptr, nbytes)
        if key not in self._views:
            self._views[key] = self.torch.as_tensor(_cai(ptr, nbytes), device=self.device)
        return self._views[key]

    def render(self, fd, host_out: bool = False):
        """Render one frame over all ranks. host_out=True: rank 0 returns the (H, W, 4) u8 image in page-locked host memory."""
        import torch.distributed as dist
        torch, be, w, h = self.torch, self.be, self.w, self.h
        with torch.cuda.stream(self.stream):  # NCCL orders itself after / before the library's stream
            work = None
            mine = [(li, a, b) for li, a, b, owner in self.units if owner == self.rank]
            if self.cast:
                be.render_shadow_units(fd, mine)  # (no library call — and no uniform upload — when this rank owns no unit)
                ptr, nbytes, cap = be.device_shadow_all()
                cb = self.chunk * w * 4
                assert cb * self.world <= cap, "shadow buffer padding too small for this world size"
                full = self._view(ptr, cb * self.world)
                # one in-place all-gather: every rank contributes the stacked rows it rasterised
                work = dist.all_gather_into_tensor(full, full[self.rank * cb:(self.rank + 1) * cb], async_op=True)
                if not self.overlap:
                    work.wait()
                    work = None
            from . import _abi as A
            flags = fd.struct.flags
            if mine:
                fd.struct.flags = flags | A.PRC_FRAME_UNIFORMS_RESIDENT  # uploaded by prc_render_shadow_units above
            be.render_forward(fd)      # camera geometry + raster + resolve overlap the exchange
            if work is not None:
                work.wait()
            fd.struct.flags = flags | A.PRC_FRAME_UNIFORMS_RESIDENT
            be.render_deferred(fd, None)
            fd.struct.flags = flags
            ptr, nbytes, cap = be.device_image()
            cb = self.img_chunk * w * 4
            assert cb * self.world <= cap, "image buffer padding too small for this world size"
            full = self._view(ptr, cb * self.world)
            # one in-place all-gather assembles the frame (every rank, rank 0 included, ends up with the image)
            dist.all_gather_into_tensor(full, full[self.rank * cb:(self.rank + 1) * cb])
            if host_out and self.rank == 0:
                # page-locked staging owned by this object, returned as a numpy view (no extra host memcpy)
                if self._host is None or self._host.numel() != nbytes:
                    self._host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
                    self._host_np = self._host.numpy().reshape(h, w, 4)
                self._host.copy_(full[:nbytes], non_blocking=True)
                self.stream.synchronize()
                return self._host_np
            self.stream.synchronize()
            return None


class PeerFrames:
    """Frames of an N-GPU group submitted back to back through prc_render_peer (include/polyred_cuda.h).

    Every rank constructs one with the same renderer configuration and calls submit()/finish() in lockstep.
    finish() returns after all submitted frames are complete on this rank; rank 0 (`root`) then holds the
    whole image of the LAST frame on the device (image()), the other ranks hold their own strip only."""

    def __init__(self, renderer, rank: int, world: int, device: int, root: int = 0, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if group is None:
            # host-only plumbing (272-byte handles, one status vote per finish()): keep it off the GPUs when gloo is usable
            try:
                group = dist.new_group(backend="gloo")
            except Exception:  # noqa: BLE001 - same outcome on every rank of a node; the default (NCCL) group works too
                group = None
        self.group = group
        self.r, self.be = renderer, renderer._backend
        self.rank, self.world, self.root = rank, world, root
        self.device = torch.device("cuda", device)
        c = renderer.cfg
        self.w, self.h = c.Width, c.Height          # the frame handed back (render.Size)
        self.msaa = max(1, int(getattr(c, "MSAA", 1)))             # the frame buffer and the shadow maps are msaa times larger (raster.go:149)
        self.hs = self.h * self.msaa
        sources, _ = c.Scene.Lights()
        self.cast = [i for i, l in enumerate(sources) if l.cast_shadow] if c.ShadowMap else []
        if self.h < world:
            # (every rank sees the same h and world: all of them raise, none is left waiting in a collective)
            from ._lib import PolyredCudaError
            from . import _abi as A
            raise PolyredCudaError(A.PRC_ERR_INVALID, f"PeerFrames: a frame of {self.h} rows cannot be cut into {world} strips")
        # contiguous ranges of OUTPUT image rows per rank (the shading partition): equal to begin with, rebalance() moves them.
        # The raster passes are partitioned by triangles inside the library and need no bounds.
        self.img_bounds = partition.equal_bounds(self.h, world)
        self._apply_bounds()
        # MSAA frames are downsampled per strip on the rank that shaded it and have no device-side gather (share_host_image())
        self.image_mask = (1 << root) if self.msaa == 1 else 0
        self._submitted = []
        self._host = self._host_np = None
        self._connect(renderer.frame_desc(no_readback=True))

    def _connect(self, fd):
        """Export, exchange, connect. Every rank performs the same two host collectives whatever happens locally, so a failure on
        one rank (no peer access, IPC refused) surfaces as an exception on ALL ranks instead of a hang in a collective."""
        from ._lib import PolyredCudaError
        from . import _abi as A
        err = handle = None
        try:
            handle = self.be.peer_export(self.prepare(fd))
        except PolyredCudaError as e:
            err = str(e)
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle, group=self.group)
        if err is None and all(h is not None for h in handles):
            try:
                self.be.peer_connect(self.rank, self.world, handles)
            except PolyredCudaError as e:
                err = str(e)
        elif err is None:
            err = "a peer could not export its buffers"
        # doubles as the barrier: nobody signals before everybody has zeroed its words and mapped its peers
        errs = [None] * self.world
        self.dist.all_gather_object(errs, err, group=self.group)
        if any(e is not None for e in errs):
            try:
                self.be.peer_disconnect()
            except PolyredCudaError:
                pass
            raise PolyredCudaError(A.PRC_ERR_PEER, "PeerFrames: connecting the ranks failed: " + "; ".join(f"rank {k}: {e}" for k, e in enumerate(errs) if e))

    def prepare(self, fd):
        fd.struct.row0, fd.struct.row1 = self.rows[self.rank]
        return fd

    def options(self, *opts):
        """Renderer.Options(...) for the whole group (render/options.go:125-141): e.g. a new camera re-fits the light cameras and
        zeroes the shadow maps. Every rank calls it with the same options; no frame may be in flight on any rank while the
        maps are zeroed (a peer's pushed texels would be lost), hence finish() before and a barrier after."""
        self.finish()
        scene_before, sd = self.r.cfg.Scene, self.r._scene_desc
        self.r.Options(*opts)
        if self.r.cfg.Scene is scene_before:
            self.r._scene_desc = sd  # Options() drops the flattened scene; the geometry did not change (as in render.RenderViews)
        self.r._ensure_uploaded()  # applies the pending shadow-map reset on this rank
        self.be.sync()
        done = [None] * self.world
        self.dist.all_gather_object(done, True, group=self.group)

    def _apply_bounds(self):
        m = self.msaa
        self.rows = [(a * m, b * m) for a, b in partition.strips_from_bounds(self.h, self.img_bounds)]  # screen rows of the frame buffer, every rank
        # the arrays prc_render_peer takes, built once per partition (a frame at 8 GPUs is ~0.2 ms: host microseconds count)
        be = getattr(self, "be", None)
        self._row_arrays = be.row_arrays(self.rows) if hasattr(be, "row_arrays") else None

    def rebalance(self, damping: float = 0.7, min_rows: int = 16):
        """Move the strip boundaries so that every rank gets the same share of the time the last finished frames spent in the
        row-partitioned kernels (resolve + shading, from prc_get_timings; the raster passes are partitioned by triangles and
        cost every rank the same). Any partition renders the same frame bit for bit, so this only moves work. Collective: call
        on every rank after finish(); returns the per-rank (raster ms, shading ms) it balanced on."""
        t = self.be.timings()
        k = list(t.kernel_ms)
        mine = (float(k[0] + k[1] + k[2] + k[3] + k[4] + k[5]), float(k[6] + k[7]))
        costs = [None] * self.world
        self.dist.all_gather_object(costs, mine, group=self.group)
        self.img_bounds = partition.balanced_bounds(self.img_bounds, [c[1] for c in costs], damping, min_rows)
        self._apply_bounds()
        return costs

    def share_host_image(self):
        """One host image in shared memory, mapped and page-locked by every rank (prc_set_host_image): frames submitted with
        readback (frame_desc(no_readback=False)) then leave each GPU as its own strip over its own PCIe link, in parallel, and
        need no device-side gather. Returns the (H, W, 4) u8 view (valid on every rank after finish())."""
        from multiprocessing import shared_memory
        from ._lib import PolyredCudaError
        from . import _abi as A
        nbytes = self.w * self.h * 4
        name = err = None
        if self.rank == self.root:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name = self._shm.name
        names = [None] * self.world
        self.dist.all_gather_object(names, name, group=self.group)
        try:
            if self.rank != self.root:
                self._shm = shared_memory.SharedMemory(name=names[self.root])
            self._shm_np = np.ndarray((self.h, self.w, 4), np.uint8, buffer=self._shm.buf)
            self.be.set_host_image(self._shm_np.ctypes.data, nbytes)
        except (PolyredCudaError, OSError) as e:
            err = str(e)
        errs = [None] * self.world
        self.dist.all_gather_object(errs, err, group=self.group)  # same collectives on every rank, error or not
        if any(e is not None for e in errs):
            raise PolyredCudaError(A.PRC_ERR_CUDA, "PeerFrames: sharing the host image failed: " + "; ".join(f"rank {k}: {e}" for k, e in enumerate(errs) if e))
        return self._shm_np

    def submit(self, fd, gather: bool = True):
        """Enqueue one frame; does not wait for the GPU or for the peers. gather=False: no device-side gather of the strips
        (frames that leave through the shared host image, share_host_image())."""
        if self.msaa > 1 and fd.struct.flags & 16 and gather:  # PRC_FRAME_NO_READBACK: an MSAA frame would stay scattered over the ranks
            from ._lib import PolyredCudaError
            from . import _abi as A
            raise PolyredCudaError(A.PRC_ERR_UNSUPPORTED, "PeerFrames: MSAA frames leave through share_host_image() (frame_desc(no_readback=False), gather=False)")
        from . import _abi as A
        flags = fd.struct.flags
        if gather and self.image_mask and (flags & A.PRC_FRAME_NO_READBACK) and not _NO_LATE_IMAGE:
            # the gathered image is read after finish() (image()): root need not stop inside every frame for its peers' strips
            fd.struct.flags = flags | A.PRC_FRAME_IMAGE_AT_SYNC
        try:
            if self._row_arrays is not None:
                self.be.render_peer_arrays(self.prepare(fd), self._row_arrays[0], self._row_arrays[1], self.image_mask if gather else 0)
            else:
                self.be.render_peer(self.prepare(fd), self.rows, self.image_mask if gather else 0)
        finally:
            fd.struct.flags = flags
        self._submitted.append((fd, gather))

    def _shm_votes(self, vote):
        """The same exchange as all_gather_object(vote) through a few words of POSIX shared memory (one cache line per rank:
        [phase, code]): microseconds instead of a pickled gloo collective — what finish(fast=True) pays per frame."""
        from multiprocessing import shared_memory
        if getattr(self, "_bar", None) is None:
            name = None
            if self.rank == self.root:
                self._bar_shm = shared_memory.SharedMemory(create=True, size=self.world * 64)
                self._bar_shm.buf[:self.world * 64] = bytes(self.world * 64)
                name = self._bar_shm.name
            names = [None] * self.world
            self.dist.all_gather_object(names, name, group=self.group)
            if self.rank != self.root:
                self._bar_shm = shared_memory.SharedMemory(name=names[self.root])
            self._bar = np.ndarray((self.world, 8), np.int64, buffer=self._bar_shm.buf)
            self._bar_phase = 0
        self._bar_phase += 1
        bar, ph = self._bar, self._bar_phase
        bar[self.rank, 1] = vote[0]
        bar[self.rank, 3] = vote[2]
        bar[self.rank, 0] = ph          # (x86: stores are not reordered with older stores)
        col = bar[:, 0]
        while int(col.min()) < ph:
            pass
        codes = [int(c) for c in bar[:, 1]]
        states = [int(c) for c in bar[:, 3]]
        # nobody may overwrite its code for the next phase before everybody has read this one: second half of the barrier
        bar[self.rank, 2] = ph
        col2 = bar[:, 2]
        while int(col2.min()) < ph:
            pass
        return [(c, "" if c == 0 else f"error {c} (see that rank's log)", st) for c, st in zip(codes, states)]

    def finish(self, max_retries: int = 3, fast: bool = False):
        """Wait for the submitted frames. A queue overflow on ANY rank (the library has grown the queue) makes
        every rank submit its frames again — the shadow maps only grow, so that is idempotent. fast=True exchanges the
        per-rank status through shared memory instead of a gloo object collective (per-frame callers)."""
        from ._lib import PolyredCudaError
        from . import _abi as A
        for _ in range(max_retries + 1):
            vote = (0, "")
            try:
                self.be.sync()
            except PolyredCudaError as e:
                vote = (e.code, str(e))
            vote = (*vote, self.be.frame_state() if hasattr(self.be, "frame_state") else 0)
            if fast:
                votes = self._shm_votes(vote)
            else:
                votes = [None] * self.world
                self.dist.all_gather_object(votes, vote, group=self.group)  # same collective on every rank, error or not
            fatal = [(k, c, m) for k, (c, m, _) in enumerate(votes) if c not in (0, A.PRC_ERR_RETRY)]
            if fatal:
                self._submitted = []
                raise PolyredCudaError(fatal[0][1], "PeerFrames: " + "; ".join(f"rank {k}: {m}" for k, _, m in fatal))
            if not any(c for c, _, _ in votes):
                self._submitted = []
                return
            # the ranks must agree on NaN mode (the first fragment of a pixel may come from any rank's triangles)
            state = 0
            for _, _, st in votes:
                state |= st
            if hasattr(self.be, "set_frame_state"):
                self.be.set_frame_state(state)
            again, self._submitted = self._submitted, []
            for fd, gather in again:
                self.submit(fd, gather)
        raise PolyredCudaError(A.PRC_ERR_UNSUPPORTED, "PeerFrames: a queue kept overflowing")

    def image(self, host: bool = True):
        """The last frame on `root` (None elsewhere): (H, W, 4) u8 in page-locked host memory, or the device tensor."""
        if self.rank != self.root:
            return None
        torch = self.torch
        ptr, nbytes, _ = self.be.device_image()
        dev = torch.as_tensor(_cai(ptr, nbytes), device=self.device)
        if not host:
            return dev.view(self.h, self.w, 4)
        if self._host is None or self._host.numel() != nbytes:
            self._host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self._host_np = self._host.numpy().reshape(self.h, self.w, 4)
        stream = torch.cuda.ExternalStream(self.be.stream(), device=self.device)
        with torch.cuda.stream(stream):
            self._host.copy_(dev, non_blocking=True)
        stream.synchronize()
        return self._host_np

    def close(self):
        self.be.peer_disconnect()
        bar = getattr(self, "_bar_shm", None)
        if bar is not None:
            self._bar = None
            bar.close()
            if self.rank == self.root:
                bar.unlink()
            self._bar_shm = None
        shm = getattr(self, "_shm", None)
        if shm is not None:
            self.be.set_host_image(None)
            self._shm_np = None
            shm.close()
            if self.rank == self.root:
                shm.unlink()
            self._shm = None
