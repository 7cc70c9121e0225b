"""CPU, world_size 2 over gloo: the multi-GPU host logic (screen strips, shadow shards, image gather)
reassembles exactly the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polyred_b200 import partition


def test_strips_and_units_cover_everything_once():
    for h in (2160, 1080, 270, 100, 17):
        for world in (1, 2, 4, 8):
            chunk, rows = partition.strips(h, world)
            assert chunk * world >= h and chunk * world - h < world
            covered = sorted(rows)
            assert covered[0][0] == 0 and covered[-1][1] == h and all(a[1] == b[0] for a, b in zip(covered, covered[1:]) if a != b or a[1] > a[0])
            assert rows[0][1] == h  # rank 0 owns the TOP image rows
            units = partition.shadow_units(h, world, [0, 2, 4, 6])
            for li in (0, 2, 4, 6):
                rows = sorted((a, b) for l, a, b, _ in units if l == li)
                assert rows[0][0] == 0 and rows[-1][1] == h and all(x[1] == y[0] for x, y in zip(rows, rows[1:]))
            if world == 8 and h % 2 == 0:
                assert sorted(o for *_, o in units) == list(range(8))  # one equal shard per GPU (C3)
            chunk, per_rank = partition.shadow_chunks(h, world, 4)
            assert chunk * world >= 4 * h and chunk * world - 4 * h < world  # padding of the in-place all-gather
    assert partition.image_rows(2160, 0, 272) == (1888, 2160)


def test_bounds_helpers_reproduce_the_equal_partition():
    for h in (2160, 270, 100, 17):
        for world in (1, 2, 3, 8):
            _, rows = partition.strips(h, world)
            assert partition.strips_from_bounds(h, partition.equal_bounds(h, world)) == rows
            for cast in ([0, 2, 4, 6], [1], []):
                assert partition.shadow_units_from_bounds(h, cast, partition.equal_bounds(len(cast) * h, world)) == partition.shadow_units(h, world, cast)


def test_balanced_bounds_equalise_measured_cost():
    """Load balancing of the peer-memory frames: boundaries move so that every rank gets the same share of the measured
    time. Simulated frame: cost per image row = sky (cheap) at the top, dense geometry in the lower half, plus a fixed
    per-rank cost the density model does not know about — repeated rebalancing must still converge."""
    assert partition.balanced_bounds([0, 50, 100], [50, 150]) == [0, 67, 100]
    assert partition.balanced_bounds([0, 25, 50, 75, 100], [1, 1, 10, 10]) == [0, 59, 72, 86, 100]
    assert partition.balanced_bounds([0, 100], [3.0]) == [0, 100]
    assert partition.balanced_bounds([0, 5, 10], [0.0, 0.0]) == [0, 5, 10]            # nothing measured: unchanged
    b = partition.balanced_bounds([0, 25, 50, 75, 100], [0, 0, 0, 10], min_size=4)
    assert b[0] == 0 and b[-1] == 100 and all(y - x >= 4 for x, y in zip(b, b[1:]))
    h, world = 2160, 8
    row_cost = np.where(np.arange(h) < 700, 0.02, 1.0) + 2.0 * np.exp(-((np.arange(h) - 1500) / 120.0) ** 2)  # image rows, top first
    fixed = 30.0

    def measure(bounds):
        return [fixed + float(row_cost[a:b].sum()) for a, b in zip(bounds, bounds[1:])]

    bounds = partition.equal_bounds(h, world)
    first = measure(bounds)
    for _ in range(6):
        bounds = partition.balanced_bounds(bounds, measure(bounds), damping=0.7, min_size=16)
        assert bounds[0] == 0 and bounds[-1] == h and all(y - x >= 16 for x, y in zip(bounds, bounds[1:]))
    last = measure(bounds)
    assert max(first) / (sum(first) / world) > 1.8                       # equal strips: the slowest rank takes ~2x the mean
    assert max(last) / (sum(last) / world) < 1.08, (bounds, last)        # balanced: within 8 % of the mean
    # stacked shadow rows: the units of a balanced partition still tile every light's rows exactly once
    sb = partition.balanced_bounds(partition.equal_bounds(4 * 100, 3), [1.0, 5.0, 2.0])
    units = partition.shadow_units_from_bounds(100, [0, 2, 4, 6], sb)
    for li in (0, 2, 4, 6):
        rows = sorted((a, b) for l, a, b, _ in units if l == li)
        assert rows[0][0] == 0 and rows[-1][1] == 100 and all(x[1] == y[0] for x, y in zip(rows, rows[1:]))


def _worker(rank, world, port, h, w, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    full_img = rng.integers(0, 255, size=(h, w, 4), dtype=np.uint8)      # what one GPU would render
    full_maps = {li: rng.random((h, w)).astype(np.float32) for li in (0, 2)}
    chunk, rows = partition.strips(h, world)
    units = partition.shadow_units(h, world, [0, 2])
    # phase 1: each rank "rasterises" only its shadow units, then every unit is broadcast from its owner
    maps = {li: np.zeros((h, w), np.float32) for li in (0, 2)}
    for li, a, b, owner in units:
        if owner == rank:
            maps[li][a:b] = full_maps[li][a:b]
    for li, a, b, owner in units:
        t = torch.from_numpy(maps[li][a:b])
        dist.broadcast(t, src=owner)
    ok_maps = all(np.array_equal(maps[li], full_maps[li]) for li in (0, 2))
    # phase 2: each rank "shades" its strip; ONE all-gather of equal (padded) strips assembles the frame everywhere
    img = np.zeros((chunk * world, w, 4), np.uint8)
    r0, r1 = partition.image_rows(h, *rows[rank])
    img[r0:r1] = full_img[r0:r1]
    parts = [torch.zeros((chunk, w, 4), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(img[rank * chunk:(rank + 1) * chunk].copy()))
    ok_img = np.array_equal(torch.cat(parts).numpy()[:h], full_img)
    q.put((rank, ok_maps, ok_img))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_reassembles_the_frame():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 100, 40, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(m and i for _, m, i in res), res


class _FakeBackend:
    """Stands in for CudaBackend on CPU: records the prc_render_peer calls; rank 1's first prc_sync reports a grown queue."""

    def __init__(self, rank, fail_connect=False, fail_sync=None):
        self.rank, self.calls, self.syncs, self.connected = rank, [], 0, None
        self.fail_connect, self.fail_sync = fail_connect, fail_sync

    def peer_export(self, fd):
        from polyred_b200 import _abi as A
        h = A.prc_peer_handle(abi_version=A.PRC_ABI_VERSION, device=self.rank, pid=os.getpid())
        assert (fd.struct.row0, fd.struct.row1) != (0, 0)
        return bytes(h)

    def peer_connect(self, rank, world, handles):
        if self.fail_connect:
            from polyred_b200 import _abi as A
            from polyred_b200._lib import PolyredCudaError
            raise PolyredCudaError(A.PRC_ERR_CUDA, "cudaIpcOpenMemHandle: refused")
        self.connected = (rank, world, [len(b) for b in handles])

    def render_peer(self, fd, rows, image_mask):
        self.calls.append((fd.tag, fd.struct.row0, fd.struct.row1, tuple(rows), image_mask, int(fd.struct.flags)))

    def frame_state(self):
        return 1 if (self.rank == 1 and self.syncs >= 1) else 0  # rank 1 entered NaN mode with its retry

    def set_frame_state(self, state):
        self.state_set = state

    def sync(self):
        from polyred_b200 import _abi as A
        from polyred_b200._lib import PolyredCudaError
        self.syncs += 1
        if self.fail_sync is not None:
            raise PolyredCudaError(self.fail_sync, "a device-side wait for a peer rank timed out")
        if self.rank == 1 and self.syncs == 1:
            raise PolyredCudaError(A.PRC_ERR_RETRY, "queue grown")

    def timings(self):
        from polyred_b200 import _abi as A
        t = A.prc_timings()
        t.kernel_ms[0] = 1.0 + 3.0 * self.rank      # raster passes (partitioned by triangles: not what the strips are balanced on)
        t.kernel_ms[7] = 4.0 - 3.0 * self.rank      # shading: rank 0 four times slower
        return t

    def set_host_image(self, address, nbytes=0):
        self.host_image = (address, nbytes)

    def peer_disconnect(self):
        pass


def _peer_worker(rank, world, port, q):
    import types
    from polyred_b200 import _abi as A
    from polyred_b200 import light, scene
    from polyred_b200.distributed import PeerFrames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def frame(tag):
        return types.SimpleNamespace(struct=A.prc_frame(abi_version=A.PRC_ABI_VERSION, width=40, height=100, flags=A.PRC_FRAME_NO_READBACK), tag=tag)

    sc = scene.Scene(light.Point(cast_shadow=True), light.Point(), light.Point(cast_shadow=True))
    be = _FakeBackend(rank)
    r = types.SimpleNamespace(cfg=types.SimpleNamespace(Width=40, Height=100, Scene=sc, ShadowMap=True), _backend=be,
                              frame_desc=lambda no_readback=True: frame("connect"))
    pf = PeerFrames(r, rank, world, 0, root=0)
    mine = [frame(k) for k in range(3)]
    for f in mine:
        pf.submit(f)
    pf.finish()
    # gathered device frames carry PRC_FRAME_IMAGE_AT_SYNC into the library (rank 0 does not stop for its peers' strips inside
    # every frame; image() is read after finish()), and the caller's frame keeps the flags it came with
    assert all(c[5] & A.PRC_FRAME_IMAGE_AT_SYNC for c in be.calls), be.calls
    assert all(f.struct.flags == A.PRC_FRAME_NO_READBACK for f in mine)
    # rebalance(): both ranks gather the same timings and derive the same new boundaries
    costs = pf.rebalance(damping=1.0, min_rows=4)
    rebalanced = (costs, pf.img_bounds, pf.rows[rank], getattr(be, "state_set", None))
    pf.img_bounds = partition.equal_bounds(100, world)
    pf._apply_bounds()
    # one shared host image: every rank maps the same pages and "reads back" its own strip into them
    img = pf.share_host_image()
    r0, r1 = partition.image_rows(100, *pf.rows[rank])
    img[r0:r1] = rank + 1
    dist.barrier()
    shared_ok = be.host_image == (img.ctypes.data, 100 * 40 * 4) and bool((img[:50] == 1).all()) and bool((img[50:] == 2).all())
    dist.barrier()
    pf.close()
    assert be.host_image == (None, 0)
    # failures on ONE rank must surface on BOTH (every rank runs the same host collectives, error or not): no hang
    from polyred_b200._lib import PolyredCudaError
    r.group = pf.group
    seen = []
    try:
        r._backend = _FakeBackend(rank, fail_connect=(rank == 1))
        PeerFrames(r, rank, world, 0, root=0, group=pf.group)
    except PolyredCudaError as e:
        seen.append(("connect", e.code, "rank 1" in str(e)))
    r._backend = _FakeBackend(rank, fail_sync=(A.PRC_ERR_PEER if rank == 0 else None))
    pf2 = PeerFrames(r, rank, world, 0, root=0, group=pf.group)
    pf2.submit(frame(7))
    try:
        pf2.finish()
    except PolyredCudaError as e:
        seen.append(("finish", e.code, "rank 0" in str(e)))
    q.put((rank, be.connected, be.calls, be.syncs, len(pf._submitted), seen, shared_ok, rebalanced))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_frames_host_logic_two_ranks():
    """PeerFrames (prc_render_peer driver): handles are gathered in rank order, every rank submits its own strip together with
    the strips of all ranks, and a retry on ONE rank makes BOTH ranks submit the batch again (lockstep epochs) after agreeing on
    the sticky frame state (NaN mode)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import ctypes as C
    from polyred_b200 import _abi as A
    for rank, connected, calls, syncs, left, seen, shared_ok, rebalanced in res:
        assert shared_ok
        costs, img_bounds, my_rows, state_set = rebalanced
        assert costs == [(1.0, 4.0), (4.0, 1.0)]
        assert img_bounds == [0, 31, 100]      # rank 0 shades four times slower: fewer image rows
        assert my_rows == ((69, 100) if rank == 0 else (0, 69))
        assert state_set == 1                  # rank 1's NaN mode was set on both ranks before the frames were submitted again
        assert seen == [("connect", A.PRC_ERR_PEER, True), ("finish", A.PRC_ERR_PEER, True)], seen
        assert connected == (rank, 2, [C.sizeof(A.prc_peer_handle)] * 2)
        assert [c[0] for c in calls] == [0, 1, 2, 0, 1, 2] and syncs == 2 and left == 0  # one retry, on both ranks
        assert all(c[4] == 1 for c in calls)  # the image goes to rank 0
        rows = {(c[1], c[2]) for c in calls}
        assert rows == {(50, 100)} if rank == 0 else rows == {(0, 50)}  # rank 0 owns the top image rows = the high screen rows
        assert calls[0][3] == ((50, 100), (0, 50))  # every call carries the strips of ALL ranks, in rank order


def test_peer_frames_msaa_rows_are_supersampled_and_aligned():
    """With render.MSAA(m) the strips are cut in OUTPUT rows and handed to the library in frame-buffer rows (multiples of m, which
    prc_render_peer requires); the shadow maps live at the supersampled size."""
    from polyred_b200.distributed import PeerFrames
    for m, h, world in ((2, 135, 4), (3, 50, 3), (1, 100, 8)):
        pf = PeerFrames.__new__(PeerFrames)
        pf.msaa, pf.h, pf.hs, pf.rank, pf.cast = m, h, h * m, 1, [0, 2]
        pf.img_bounds = partition.balanced_bounds(partition.equal_bounds(h, world), [1.0 + k for k in range(world)])
        pf._apply_bounds()
        assert all(a % m == 0 and b % m == 0 for a, b in pf.rows)
        assert sorted(pf.rows)[0][0] == 0 and sorted(pf.rows)[-1][1] == h * m and sum(b - a for a, b in pf.rows) == h * m


def test_balanced_bounds_properties_hypothesis():
    """Whatever the measured costs, the new boundaries cover the same range, never decrease, keep min_size rows per rank when
    there is room, and a perfectly balanced partition of a uniform density is a fixed point."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 9), st.integers(0, 5000), st.data())
    def prop(n, total, data):
        cuts = sorted(data.draw(st.lists(st.integers(0, total), min_size=n - 1, max_size=n - 1)))
        bounds = [0] + cuts + [total]
        cost = data.draw(st.lists(st.floats(0, 1e3, allow_nan=False, allow_infinity=False), min_size=n, max_size=n))
        damping = data.draw(st.sampled_from([0.3, 0.7, 1.0]))
        min_size = data.draw(st.sampled_from([1, 4, 16]))
        new = partition.balanced_bounds(bounds, cost, damping, min_size)
        assert len(new) == n + 1 and new[0] == 0 and new[-1] == total
        assert all(a <= b for a, b in zip(new, new[1:]))
        if sum(cost) > 0 and total >= n * min_size and n > 1 and total > 0:
            assert all(b - a >= min_size for a, b in zip(new, new[1:])), (bounds, cost, new)

    prop()
    for n in (2, 3, 8):
        eq = partition.equal_bounds(240 * n, n)
        assert partition.balanced_bounds(eq, [5.0] * n) == eq


def test_distributed_frame_rejects_impossible_partitions_on_every_rank():
    """equal strips of ceil(H/world) rows leave trailing ranks an EMPTY strip when H is small, and only those
    ranks failed (inside the library) while the others sat in the all-gather. The constructor now refuses such a partition —
    and MSAA, and more ranks than shadow rows — from the configuration alone, i.e. on every rank alike, before any CUDA call."""
    import types
    import pytest
    from polyred_b200._lib import PolyredCudaError
    from polyred_b200.distributed import DistributedFrame

    def fake(h, msaa=1, n_cast=1):
        lights = [types.SimpleNamespace(cast_shadow=True) for _ in range(n_cast)]
        cfg = types.SimpleNamespace(Width=64, Height=h, MSAA=msaa, ShadowMap=True, Scene=types.SimpleNamespace(Lights=lambda: (lights, [])))
        return types.SimpleNamespace(cfg=cfg, _backend=None)

    for rank in range(8):  # 9 rows over 8 ranks: strips of 2 rows, ranks 5..7 would be empty
        with pytest.raises(PolyredCudaError, match="non-empty strips"):
            DistributedFrame(fake(9), rank, 8, 0)
    with pytest.raises(PolyredCudaError, match="MSAA"):
        DistributedFrame(fake(64, msaa=2), 0, 2, 0)
    with pytest.raises(PolyredCudaError, match="non-empty strips"):
        DistributedFrame(fake(3), 0, 4, 0)
