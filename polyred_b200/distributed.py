"""One frame over N GPUs, one process per GPU (torch.distributed / NCCL over NVLink for the plumbing).

The reference has no multi-GPU path. Partition (north_star): the scene is replicated, the
screen is cut into N horizontal strips (16-row aligned), the Ls shadow maps are cut into (light, row
range) shards dealt round-robin. Per frame:
  1. every rank rasterises its shadow shards into its local copy of the maps   (prc_render_shadows)
  2. each shard is broadcast from its owner into the other ranks' maps         (NCCL, device pointers)
  3. every rank runs forward + deferred for its strip                          (prc_render_main)
  4. the RGBA8 strips are sent to rank 0's image buffer                        (NCCL send/recv)
The result is bit-identical to the 1-GPU frame: atomicMax keys / depth maxima do not depend on who
rasterised what, and every rank also rasterises pixel (0,0) (bug-list 3) and the AO halo rows.
"""
from __future__ import annotations

import numpy as np

from . import partition


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
        self.cuts = partition.strips(self.h, world)
        self.units = partition.shadow_units(self.h, world, self.cast)
        self._views = {}

    def prepare(self, fd):
        fd.struct.row0, fd.struct.row1 = self.cuts[self.rank], self.cuts[self.rank + 1]
        return fd

    def _view(self, ptr, nbytes):
        key = (ptr, nbytes)
        if key not in self._views:
            self._views[key] = self.torch.as_tensor(_cai(ptr, nbytes), device=self.device)
        return self._views[key]

    def render(self, fd, host_out: np.ndarray | None = None):
        import torch.distributed as dist
        torch, be, w, h = self.torch, self.be, self.w, self.h
        for li, a, b, owner in self.units:
            if owner == self.rank:
                be.render_shadows(fd, 1 << li, a, b)
        if self.units:
            be.sync()
            for li, a, b, owner in self.units:
                ptr, _ = be.device_shadowmap(li)
                dist.broadcast(self._view(ptr + a * w * 4, (b - a) * w * 4), src=owner)
            torch.cuda.synchronize()
        be.render_main(fd, None)
        ptr, nbytes = be.device_image()
        img = self._view(ptr, nbytes)
        for k in range(1, self.world):
            ia, ib = partition.image_rows(h, self.cuts[k], self.cuts[k + 1])
            seg = img[ia * w * 4:ib * w * 4]
            if self.rank == k:
                dist.send(seg, dst=0)
            elif self.rank == 0:
                dist.recv(seg, src=k)
        torch.cuda.synchronize()
        if host_out is not None and self.rank == 0:
            host_out.reshape(-1)[:] = img.cpu().numpy()
