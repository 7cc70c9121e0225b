"""Run under torchrun on N GPUs: the N-GPU frame must be bit-identical to rank 0's own 1-GPU frame.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from polyred_b200 import render, synth
    from polyred_b200.distributed import DistributedFrame, PeerFrames
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, (s, cam), w, h in (
        ("city", synth.city_scene(n_objects=64, obj_stacks=16, obj_slices=16, ground_cells=80, tex_size=64), 960, 540),
        ("c2-ao", synth.mesh_scene(subdiv=40, with_ground=True, shadows=True, ao=True), 320, 200),
    ):
        opts = [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
        r = render.NewRenderer(*opts, render.CUDA(local))
        r._ensure_uploaded()
        df = DistributedFrame(r, rank, world, local)
        for _ in range(2):
            out = df.render(df.prepare(r.frame_desc(no_readback=True)), True)
        peer_out = rebalanced_out = shared_out = None
        moving_nd = -1
        if os.environ.get("PRC_CHECK_PEER", "1") != "0":
            # the same frame through prc_render_peer (NVLink peer memory, no collective): three frames back to back
            r2 = render.NewRenderer(*opts, render.CUDA(local))
            r2._ensure_uploaded()
            pf = PeerFrames(r2, rank, world, local, root=0)
            for _ in range(3):
                pf.submit(r2.frame_desc(no_readback=True))
            pf.finish()
            peer_out = pf.image(host=True)
            if peer_out is not None:
                peer_out = peer_out.copy()
            # any partition must give the same frame: move the boundaries by the measured per-rank times, render again
            costs = pf.rebalance(damping=1.0, min_rows=8)
            for _ in range(2):
                pf.submit(r2.frame_desc(no_readback=True))
            pf.finish()
            rebalanced_out = pf.image(host=True)
            if rebalanced_out is not None:
                rebalanced_out = rebalanced_out.copy()  # image() hands out one reused page-locked buffer
            if rank == 0:
                print(f"[multigpu_check] {name}: rebalanced strips {[b - a for a, b in pf.rows]} from per-rank (raster, shading) ms {[(round(c[0], 3), round(c[1], 3)) for c in costs]}")
            # and the e2e form: every rank reads its own strip back into one shared host image, no device-side gather
            shared = pf.share_host_image()
            pf.submit(r2.frame_desc(no_readback=False), gather=False)
            pf.finish()
            shared_out = shared.copy() if rank == 0 else None
            dist.barrier()
            if name == "city":
                # strips that move back and forth while the camera moves (re-balancing under a moving camera): a row a rank loses and
                # regains two frames later must not still hold the keys of the old camera (tests/test_gpu_peer.py has the one-process form)
                from polyred_b200 import partition
                bounds = [partition.equal_bounds(h, world), list(pf.img_bounds)]
                one = render.NewRenderer(*opts, render.CUDA(local)) if rank == 0 else None
                for k in range(1, 7):
                    cam_k = synth.orbit_camera(0.3 * k, aspect=w / h)
                    pf.img_bounds = bounds[k % 2]
                    pf._apply_bounds()
                    r2.cfg.Camera = cam_k
                    pf.submit(r2.frame_desc(no_readback=True))
                    pf.finish()
                    img = pf.image(host=True)
                    if rank == 0:
                        one.cfg.Camera = cam_k
                        nd = int((img != one.Render()).any(axis=2).sum())
                        moving_nd = max(moving_nd, nd)
                dist.barrier()
            pf.close()
        if rank == 0:
            ref = render.NewRenderer(*opts, render.CUDA(local)).Render()
            nd = int((np.abs(out.astype(int) - ref.astype(int)).max(axis=2) > 0).sum())
            print(f"[multigpu_check] {name} {w}x{h} world={world}: pixels differing from the 1-GPU frame = {nd}")
            ok = ok and nd == 0
            if peer_out is not None:
                for label, img in (("peer memory", peer_out), ("peer memory, rebalanced strips", rebalanced_out), ("peer memory, strips read back into one shared host image", shared_out)):
                    nd = int((np.abs(img.astype(int) - ref.astype(int)).max(axis=2) > 0).sum())
                    print(f"[multigpu_check] {name} {w}x{h} world={world} ({label}): pixels differing from the 1-GPU frame = {nd}")
                    ok = ok and nd == 0
                if moving_nd >= 0:
                    print(f"[multigpu_check] {name} {w}x{h} world={world} (peer memory, 6 frames: camera moving, strips alternating between two partitions): worst frame, pixels differing from the 1-GPU frame = {moving_nd}")
                    ok = ok and moving_nd == 0
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
