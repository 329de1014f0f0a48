"""WebSocket host: the reference's ClientMessage / binary record protocol
(src/server.rs:80-126, 173-190).  CPU tests drive the host with a stub job; the GPU test renders."""
import asyncio
import json
import threading
import time

import numpy as np
import pytest

websockets = pytest.importorskip("websockets")


def record(x, y, px):
    px = np.asarray(px, dtype=np.uint8).reshape(-1, 3)
    return bytes([0, px.shape[0]]) + int(x).to_bytes(2, "little") + int(y).to_bytes(2, "little") + px.tobytes()


class StubJob:
    """Yields a W x H gradient in the reference's record order, slowly enough to be stopped."""

    def __init__(self, w, h, delay=0.0):
        self.w, self.h, self.delay = w, h, delay
        self.stopped = False
        self.closed = False

    def messages(self):
        for y in range(self.h):
            for x in range(0, self.w, 60):
                if self.stopped:
                    return
                n = min(60, self.w - x)
                px = np.stack([np.arange(x, x + n) % 256, np.full(n, y % 256), np.full(n, 7)], axis=1)
                if self.delay:
                    time.sleep(self.delay)
                yield record(x, y, px)

    def stop(self):
        self.stopped = True

    def close(self):
        self.closed = True
        return self.stopped


def run_server(server):
    """starts server.listen on a free port in a background thread; returns (port, stop)"""
    ready = {}
    loop = asyncio.new_event_loop()

    def target():
        asyncio.set_event_loop(loop)

        async def go():
            async with websockets.serve(server.handle_connection, "127.0.0.1", 0) as srv:
                ready["port"] = srv.sockets[0].getsockname()[1]
                ready["stop"] = asyncio.Event()
                await ready["stop"].wait()

        loop.run_until_complete(go())

    t = threading.Thread(target=target, daemon=True)
    t.start()
    for _ in range(200):
        if "stop" in ready:
            break
        time.sleep(0.01)

    def stop():
        loop.call_soon_threadsafe(ready["stop"].set)
        t.join(5)

    return ready["port"], stop


def collect(port, requests, w, h, stop_after=None, timeout=30):
    async def go():
        nonlocal stop_after
        frame = np.zeros((h, w, 3), dtype=np.uint8)
        seen = 0
        async with websockets.connect(f"ws://127.0.0.1:{port}", max_size=1 << 20) as ws:
            for r in requests:
                await ws.send(r if isinstance(r, str) else json.dumps(r))
            while seen < w * h:
                try:
                    m = await asyncio.wait_for(ws.recv(), timeout=2.0 if stop_after else timeout)
                except asyncio.TimeoutError:
                    break
                assert isinstance(m, bytes) and m[0] == 0 and len(m) == 6 + 3 * m[1]
                n, x, y = m[1], int.from_bytes(m[2:4], "little"), int.from_bytes(m[4:6], "little")
                frame[y, x: x + n] = np.frombuffer(m[6:], dtype=np.uint8).reshape(n, 3)
                seen += n
                if stop_after and seen >= stop_after:
                    await ws.send(json.dumps({"type": "stop_rendering"}))
                    stop_after = None
                    await asyncio.sleep(0.3)
        return frame, seen

    return asyncio.run(go())


def make_server(rtb_mod, jobs, w=130, h=20, delay=0.0):
    from raytracer_server_b200.server import Server

    def factory(scene, spp, passes=1):
        if scene != "cornell_box":
            return None
        j = StubJob(w, h, delay)
        jobs.append((scene, spp, passes, j))
        return j

    return Server({}, width=w, height=h, job_factory=factory, log=lambda *a: None)


def test_render_request_streams_records(rtb):
    jobs = []
    port, stop = run_server(make_server(rtb, jobs))
    try:
        frame, seen = collect(port, [{"type": "render", "scene": "cornell_box", "spp": 64}], 130, 20)
    finally:
        stop()
    assert seen == 130 * 20 and jobs[0][:3] == ("cornell_box", 64, 1)
    assert (frame[:, :, 0] == (np.arange(130) % 256)[None, :]).all() and (frame[:, :, 1] == np.arange(20)[:, None]).all()
    assert jobs[0][3].closed


def test_ignored_messages_and_one_job_per_connection(rtb):
    jobs = []
    port, stop = run_server(make_server(rtb, jobs, delay=0.002))
    try:
        reqs = [{"type": "stop_rendering"},                                    # idle: ignored (src/server.rs:108-112)
                {"type": "render", "scene": "nope", "spp": 4},                  # unknown scene: nothing happens
                {"type": "render", "scene": "cornell_box", "spp": 8},
                {"type": "render", "scene": "cornell_box", "spp": 16},          # already running: ignored (:94)
                {"type": "dance"}]
        frame, seen = collect(port, reqs, 130, 20)
    finally:
        stop()
    assert seen == 130 * 20 and len(jobs) == 1 and jobs[0][1] == 8


def test_stop_rendering_cancels(rtb):
    jobs = []
    port, stop = run_server(make_server(rtb, jobs, h=200, delay=0.002))
    try:
        frame, seen = collect(port, [{"type": "render", "scene": "cornell_box", "spp": 8}], 130, 200, stop_after=130 * 5)
    finally:
        stop()
    assert 130 * 5 <= seen < 130 * 200 and jobs[0][3].stopped and jobs[0][3].closed


def test_bad_json_closes_the_connection(rtb):
    jobs = []
    port, stop = run_server(make_server(rtb, jobs))

    async def go():
        async with websockets.connect(f"ws://127.0.0.1:{port}") as ws:
            await ws.send("{not json")
            with pytest.raises(websockets.ConnectionClosed):
                await asyncio.wait_for(ws.recv(), timeout=5)

    try:
        asyncio.run(go())
    finally:
        stop()
    assert jobs == []


def test_connection_ids(rtb):
    from raytracer_server_b200.server import Server

    s = Server({}, log=lambda *a: None)
    ids = {s.generate_connection_id() for _ in range(200)}
    assert len(ids) == 200 and all(len(i) == 5 and len(set(i)) == 5 and i.islower() for i in ids)


@pytest.mark.gpu
def test_server_renders_reference_frame_on_gpu(rtb, gpu_scene):
    from raytracer_server_b200.server import Server

    scene = gpu_scene("cornell_box")
    srv = Server({"cornell_box": scene}, log=lambda *a: None)      # 600 x 450 like the reference
    port, stop = run_server(srv)
    try:
        frame, seen = collect(port, [{"type": "render", "scene": "cornell_box", "spp": 16}], 600, 450, timeout=60)
    finally:
        stop()
    assert seen == 600 * 450
    ref = scene.render(600, 450, 16, seed=1).astype(int)
    # the server draws a fresh seed per request (the reference is unseeded): compare statistically
    assert np.allclose(frame.reshape(-1, 3).mean(0), ref.reshape(-1, 3).mean(0), rtol=0.02)
    assert np.abs(frame.astype(int) - ref).mean() < 30     # two independent 16-spp frames: Monte-Carlo noise only
