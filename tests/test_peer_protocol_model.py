"""CPU: a model of the prc_render_peer protocol (polyred_b200/csrc/polyred_cuda.cu: enqueue_peer_frame) under random
schedules. Every rank's stream executes the operations of its frames in order; waits block on the epoch words the peers
write; the host of each rank submits at its own pace. The model checks what the device code relies on:

  * no deadlock, whatever the interleaving of the ranks' GPUs and hosts (also when the set of image consumers changes from
    frame to frame and when a rank's host runs many frames ahead of the others);
  * when a rank shades frame e, every row of its copy of the shadow maps holds exactly the owner's frame-e state (no push
    of frame e+1 has landed yet, none of frame e is missing);
  * when a consumer reads the image of frame e, every strip in it is frame e's (no strip of e+1 has overwritten it);
  * epoch words only ever increase.

The operation list below is a transcription of enqueue_peer_frame; if the code changes, change it here too."""
import random

import pytest

SHADOW, SHADED, IMAGE, IMAGE_FREE = range(4)


def frame_ops(rank, world, e, mask, shadows=True, late=False, prev_late=False):
    """The stream operations of frame e on `rank` (enqueue_peer_frame, in order): (main stream ops, side stream ops).
    late = PRC_FRAME_IMAGE_AT_SYNC (a consumer waits for its peers' strips beside the stream); prev_late = this rank received
    frame e-1 that way (its "image free" announcement and its image-complete event belong to the side stream)."""
    me, everyone = 1 << rank, (1 << world) - 1
    others = everyone & ~me
    ops, side = [], []
    if (mask & me) and not prev_late:
        ops.append(("signal", IMAGE_FREE, e - 1, others))
    ops.append(("clear_next_keys", e))  # the merged key plane of frame e+1 (the planes alternate with the frame parity)
    ops.append(("forward", e))  # the camera pass needs no shadow map: it runs while slower peers still shade frame e-1
    ops.append(("push_keys", e))  # early key push: no wait (the peers cleared plane e&1 before they signalled SHADOW(e-1))
    if shadows:
        ops.append(("raster_shadow", e))
        ops.append(("wait", SHADED, e - 1, others))
        ops.append(("push", e))
        ops.append(("push_keys", e))  # what the queued records added
        ops.append(("signal", SHADOW, e, others))
        ops.append(("wait", SHADOW, e, others))
        ops.append(("signal", SHADOW, e, others))
        ops.append(("wait", SHADOW, e, others))
    if prev_late:
        ops.append(("await", ("image_done", e - 1)))  # the previous frame's strips have landed before this shading writes the image
    ops.append(("shade", e))
    if shadows:
        ops.append(("signal", SHADED, e, others))
    consumers = mask & everyone
    if consumers & ~me:
        ops.append(("wait", IMAGE_FREE, e - 1, consumers & ~me))
        ops.append(("copy_strip", e, consumers & ~me))
        ops.append(("signal", IMAGE, e, consumers & ~me))
    if consumers & me:
        if late and others:
            ops.append(("record", ("own_shaded", e)))
            side.append(("await", ("own_shaded", e)))
            side.append(("wait", IMAGE, e, others))
            side.append(("read_image", e))  # complete here, and until a frame e+1 is submitted (what prc_sync hands out)
            side.append(("signal", IMAGE_FREE, e, others))
            side.append(("record", ("image_done", e)))
        else:
            ops.append(("wait", IMAGE, e, others))
            ops.append(("read_image", e))  # a stream-ordered reader enqueued by the caller before the next frame
    return ops, side


def simulate(world, masks, seed, host_lookahead, late=None, starve=None):
    """late: per-frame booleans (PRC_FRAME_IMAGE_AT_SYNC on the consumers of that frame); None = never.
    starve = (rank, queue): that stream only runs when nothing else can (an adversarial schedule)."""
    rng = random.Random(seed)
    frames = len(masks)
    late = late or [False] * frames
    words = [[[0] * world for _ in range(4)] for _ in range(world)]       # words[dst][kind][src]
    maps = [[0] * world for _ in range(world)]                            # maps[holder][owner] = frame whose rows are in holder's copy
    image = [[0] * world for _ in range(world)]                           # image[holder][strip owner]
    keys = [[dict(), dict()] for _ in range(world)]                       # keys[holder][parity] = {pusher: frame} of the merged key planes
    queue = [[[], []] for _ in range(world)]                              # per rank: [main stream, side stream], submitted and not yet executed
    events = set()                                                        # (rank, name) recorded
    submitted = [0] * world                                               # frames submitted by each host
    executed_frames = [0] * world
    was_late_consumer = [False] * world

    def runnable(r, q):
        if not queue[r][q]:
            return False
        op = queue[r][q][0]
        if op[0] == "wait":
            _, kind, epoch, mask = op
            return all(words[r][kind][s] >= epoch for s in range(world) if (mask >> s) & 1)
        if op[0] == "await":
            return (r, op[1]) in events
        return True

    steps = 0
    while any(q for rq in queue for q in rq) or any(s < frames for s in submitted):
        steps += 1
        assert steps < 400000
        choices = [("gpu", r, q) for r in range(world) for q in (0, 1) if runnable(r, q)]
        # a host may run ahead of its own GPU by `host_lookahead` frames (bounded launch queue), independently of the others
        choices += [("host", r, 0) for r in range(world) if submitted[r] < frames and submitted[r] - executed_frames[r] < host_lookahead]
        if starve is not None and any(c[1:] != starve or c[0] == "host" for c in choices):
            choices = [c for c in choices if c[1:] != starve or c[0] == "host"]
        assert choices, f"deadlock: world={world} masks={masks} late={late} seed={seed} heads={[[q[0] if q else None for q in rq] for rq in queue]}"
        what, r, q = rng.choice(choices)
        if what == "host":
            e = submitted[r] + 1
            ops, side = frame_ops(r, world, e, masks[e - 1], late=late[e - 1], prev_late=was_late_consumer[r])
            was_late_consumer[r] = bool(side)
            queue[r][0].extend(ops + [("frame_done", e)])
            queue[r][1].extend(side)
            submitted[r] = e
            continue
        op = queue[r][q].pop(0)
        kind = op[0]
        if kind == "signal":
            _, k, epoch, mask = op
            for d in range(world):
                if (mask >> d) & 1:
                    assert words[d][k][r] <= epoch, "epoch words must not decrease"
                    words[d][k][r] = epoch
        elif kind == "record":
            events.add((r, op[1]))
        elif kind == "clear_next_keys":
            e = op[1]
            plane = keys[r][(e + 1) & 1]
            assert all(f < e + 1 for f in plane.values()), f"rank {r} clears keys of frame {e + 1} that a peer has already pushed"
            plane.clear()
        elif kind == "push_keys":
            e = op[1]
            for p in range(world):
                plane = keys[p][e & 1]
                assert all(f == e for f in plane.values()), f"rank {r} pushes keys of frame {e} into a plane of rank {p} that still holds {plane}"
                plane[r] = e
        elif kind == "raster_shadow":
            maps[r][r] = op[1]
        elif kind == "push":
            for p in range(world):
                if p != r:
                    maps[p][r] = op[1]
        elif kind == "shade":
            e = op[1]
            assert maps[r] == [e] * world, f"rank {r} shades frame {e} from shadow rows of frames {maps[r]}"
            assert keys[r][e & 1] == {p: e for p in range(world)}, f"rank {r} shades frame {e} from keys {keys[r][e & 1]}"
            image[r][r] = e
        elif kind == "copy_strip":
            _, e, mask = op
            for c in range(world):
                if (mask >> c) & 1:
                    assert image[c][r] <= e, "a strip of an older frame overwrites a newer one"
                    image[c][r] = e
        elif kind == "read_image":
            e = op[1]
            assert image[r] == [e] * world, f"consumer {r} reads frame {e} but its image holds strips of frames {image[r]}"
        elif kind == "frame_done":
            executed_frames[r] = op[1]
    return steps


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_protocol_model_random_schedules(world):
    rng = random.Random(1000 + world)
    everyone = (1 << world) - 1
    for trial in range(60 if world <= 3 else 25):
        frames = rng.randint(1, 7)
        style = trial % 4
        if style == 0:
            masks = [1] * frames                                  # north_star: the image goes to rank 0
        elif style == 1:
            masks = [everyone] * frames                           # every rank receives the image
        elif style == 2:
            masks = [0] * frames                                  # no device-side gather (strips leave through the shared host image)
        else:
            masks = [rng.randint(0, everyone) for _ in range(frames)]  # the consumer set changes from frame to frame
        for lookahead in (1, 2, frames + 1):
            simulate(world, masks, seed=rng.randint(0, 1 << 30), host_lookahead=lookahead)
            # PRC_FRAME_IMAGE_AT_SYNC on every frame, and switched on and off from frame to frame
            simulate(world, masks, seed=rng.randint(0, 1 << 30), host_lookahead=lookahead, late=[True] * frames)
            simulate(world, masks, seed=rng.randint(0, 1 << 30), host_lookahead=lookahead, late=[rng.random() < 0.5 for _ in range(frames)])


def test_model_detects_a_broken_protocol():
    """The checker is not vacuous: without the SHADED wait a fast rank pushes frame e+1 into maps a slow rank still shades
    frame e from, and some schedule exposes it."""
    global frame_ops
    good = frame_ops

    def broken(rank, world, e, mask, shadows=True, late=False, prev_late=False):
        ops, side = good(rank, world, e, mask, shadows, late, prev_late)
        return [op for op in ops if not (op[0] == "wait" and op[1] == SHADED)], side

    frame_ops = broken
    try:
        with pytest.raises(AssertionError, match="shades frame"):
            for seed in range(300):
                simulate(3, [1, 1, 1, 1], seed=seed, host_lookahead=4)
    finally:
        frame_ops = good


# ---- rows of the merged key planes when the strips move between frames (re-balancing) ----
def _plane_rows_model(strips_per_frame, height, rule):
    """One rank's two merged key planes, row by row: plane[parity][row] = frame whose keys the row holds (0 = empty). Frame e
    clears plane (e+1)&1 by `rule`, receives keys for its strip of frame e in plane e&1, then resolves that strip."""
    plane = [[0] * height, [0] * height]
    last = [None, None]  # rows the last frame of each parity received (enqueue_peer_frame: ctx->plane_rows)
    for e, (r0, r1) in enumerate(strips_per_frame, start=1):
        nxt = (e + 1) & 1
        if rule == "current_strip":  # round 2's first version
            lo, hi = r0, r1
        else:                        # "last_written": what the plane last received
            lo, hi = last[nxt] if last[nxt] else (0, 0)
            last[nxt] = None
        for y in range(lo, hi):
            plane[nxt][y] = 0
        for y in range(r0, r1):      # atomicMax merge: an older frame's keys survive under the new ones
            if plane[e & 1][y] not in (0, e):
                return e, y
            plane[e & 1][y] = e
        last[e & 1] = (r0, r1)
    return None


def test_key_planes_stay_clean_when_strips_move_back_and_forth():
    rng = random.Random(7)
    H = 64
    for trial in range(300):
        strips = []
        for _ in range(rng.randint(2, 12)):
            a = rng.randint(0, H - 1)
            strips.append((a, rng.randint(a + 1, H)))
        assert _plane_rows_model(strips, H, "last_written") is None, strips
    # the checker is not vacuous: clearing the CURRENT strip's rows leaves stale keys in a row lost and regained two frames later
    assert _plane_rows_model([(0, 40), (0, 10), (0, 40)], H, "current_strip") == (3, 10)


def test_model_detects_an_early_image_free_announcement():
    """PRC_FRAME_IMAGE_AT_SYNC: "image free up to e" must come from the side stream, after the strips of frame e have landed. If
    the consumer's next frame announced it as its first operation (what frames without the flag do), a fast peer's strip of frame
    e+1 could land while a slow peer's strip of frame e is still on its way: the image handed out for e would be mixed."""
    global frame_ops
    good = frame_ops

    def broken(rank, world, e, mask, shadows=True, late=False, prev_late=False):
        ops, side = good(rank, world, e, mask, shadows, late, prev_late)
        if (mask & (1 << rank)) and prev_late:
            ops = [("signal", IMAGE_FREE, e - 1, ((1 << world) - 1) & ~(1 << rank))] + ops
        return ops, side

    frame_ops = broken
    try:
        with pytest.raises(AssertionError, match="reads frame"):
            for seed in range(100):  # rank 0's side stream is starved: it runs only when every other stream is blocked
                simulate(3, [1, 1, 1, 1], seed=seed, host_lookahead=4, late=[True] * 4, starve=(0, 1))
    finally:
        frame_ops = good
    for seed in range(300):  # the real operation list passes the same adversarial schedules (and with the main stream starved)
        simulate(3, [1, 1, 1, 1], seed=seed, host_lookahead=4, late=[True] * 4, starve=(0, 1))
        simulate(3, [1, 1, 1, 1], seed=seed, host_lookahead=4, late=[True] * 4, starve=(0, 0))
