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


@pytest.mark.parametrize("spp", ["64", 64.0, True, None, 2 ** 31])
def test_spp_must_be_a_json_i32(rtb, spp):
    # ClientMessage::Render { scene: String, spp: i32 } (src/server.rs:121-126): serde rejects anything else
    jobs = []
    port, stop = run_server(make_server(rtb, jobs))

    async def go():
        async with websockets.connect(f"ws://127.0.0.1:{port}") as ws:
            await ws.send(json.dumps({"type": "render", "scene": "cornell_box", "spp": spp}))
            with pytest.raises(websockets.ConnectionClosed):
                await asyncio.wait_for(ws.recv(), timeout=5)

    try:
        asyncio.run(go())
    finally:
        stop()
    assert jobs == []


def test_refused_request_keeps_the_connection(rtb):
    # the library refusing a request (RtbError, e.g. spp beyond the sample-index field) must not drop the client;
    # a negative spp is a valid i32 and reaches the factory (the reference renders a black frame for it)
    from raytracer_server_b200.host import RtbError
    from raytracer_server_b200.server import Server

    seen_spp = []

    def factory(scene, spp, passes=1):
        seen_spp.append(spp)
        if spp > 1000:
            raise RtbError(-7, "spp too large")
        return StubJob(130, 20)

    port, stop = run_server(Server({}, width=130, height=20, job_factory=factory, log=lambda *a: None))
    try:
        frame, seen = collect(port, [{"type": "render", "scene": "cornell_box", "spp": 5_000_000},
                                     {"type": "render", "scene": "cornell_box", "spp": -8}], 130, 20)
    finally:
        stop()
    assert seen == 130 * 20 and seen_spp == [5_000_000, -8]


def test_default_factory_clamps_negative_spp(rtb, monkeypatch):
    import raytracer_server_b200.host as host
    from raytracer_server_b200.server import default_job_factory

    made = []
    monkeypatch.setattr(host, "RenderJob", lambda scene, w, h, spp, **kw: made.append((w, h, spp, kw["passes"])) or "job")
    make = default_job_factory({"cornell_box": object()}, 600, 450)
    assert make("nope", 4) is None
    assert make("cornell_box", -12, 3) == "job" and made == [(600, 450, 0, 3)]


def test_stop_never_touches_a_freed_job(rtb, monkeypatch):
    # RenderJob.stop() (event-loop thread) racing RenderJob.close() (executor thread): rtb_job_cancel must never be
    # called with a handle that rtb_job_end has been given (it deletes the job)
    import ctypes as C

    import raytracer_server_b200.host as host

    calls, ended = [], threading.Event()

    class FakeLib:
        def rtb_job_stats(self, h, s):
            return 0

        def rtb_job_cancel(self, h):
            calls.append(("cancel", h.value, ended.is_set()))
            return 0

        def rtb_job_end(self, h):
            ended.set()
            time.sleep(0.05)            # the window in which the old code still saw a non-null handle
            calls.append(("end", h.value, True))
            return 1

    monkeypatch.setattr(host._abi, "lib", lambda: FakeLib())
    for _ in range(20):
        calls.clear()
        ended.clear()
        job = object.__new__(host.RenderJob)
        job._h, job._lock, job.cancelled, job.final_stats = C.c_void_p(1234), threading.Lock(), False, None
        t = threading.Thread(target=job.close)
        t.start()
        while t.is_alive():
            job.stop()
        t.join()
        job.stop()
        assert ("end", 1234, True) in calls
        assert not any(c[0] == "cancel" and c[2] for c in calls), "rtb_job_cancel after rtb_job_end was entered"
        assert job.close() is True and job._h.value is None


def test_connection_ids(rtb):
    from raytracer_server_b200.server import Server

    s = Server({}, log=lambda *a: None)
    ids = {s.generate_connection_id() for _ in range(200)}
    assert len(ids) == 200 and all(len(i) == 5 and len(set(i)) == 5 and i.islower() for i in ids)


def seeded_factory(scenes, w, h, seeds):
    from raytracer_server_b200.host import RenderJob

    def make(scene_name, spp, passes=1):
        if scene_name not in scenes:
            return None
        return RenderJob(scenes[scene_name], w, h, max(0, spp), passes=passes, seed=seeds[scene_name, spp])

    return make


@pytest.mark.gpu
def test_server_renders_reference_frame_on_gpu(rtb, gpu_scene, parity_log):
    # the reference's own example: cornell_box at the server's 600 x 450, 64 spp — through the socket, record by record.
    # Checked (a) against the blocking render of the same seed and (b) against the image the REFERENCE produced
    # (examples/cornell_box.png statistics, tests/golden/reference_pins.npz)
    from raytracer_server_b200.server import Server
    from test_reference_pins import check_against_reference

    scene = gpu_scene("cornell_box")
    srv = Server({}, job_factory=seeded_factory({"cornell_box": scene}, 600, 450, {("cornell_box", 64): 4242}), log=lambda *a: None)
    port, stop = run_server(srv)
    try:
        frame, seen = collect(port, [{"type": "render", "scene": "cornell_box", "spp": 64}], 600, 450, timeout=60)
    finally:
        stop()
    assert seen == 600 * 450
    assert np.abs(frame.astype(int) - scene.render(600, 450, 64, seed=4242).astype(int)).max() <= 1
    check_against_reference(frame, "cornell_box", "gpu_server", parity_log)


@pytest.mark.gpu
def test_four_connections_render_concurrently(rtb, gpu_scene, parity_log):
    # scenes shared by all connections, one job per connection, jobs of different connections run
    # concurrently on one device (a stream + context per job).  Each client must receive exactly its own frame.
    from raytracer_server_b200.server import Server

    W, H = 600, 450
    reqs = [("cornell_box", 64), ("cubes", 32), ("flying_unicorn", 32), ("cornell_box", 16)]
    scenes = {n: gpu_scene(n) for n, _ in reqs}
    seeds = {r: 9000 + i for i, r in enumerate(reqs)}
    want = {r: scenes[r[0]].render(W, H, r[1], seed=seeds[r]).astype(int) for r in reqs}
    srv = Server({}, job_factory=seeded_factory(scenes, W, H, seeds), log=lambda *a: None)
    port, stop = run_server(srv)
    got = {}

    def client(r):
        got[r] = collect(port, [{"type": "render", "scene": r[0], "spp": r[1]}], W, H, timeout=120)

    try:
        t0 = time.time()
        ts = [threading.Thread(target=client, args=(r,)) for r in reqs]
        [t.start() for t in ts]
        [t.join() for t in ts]
        dt = time.time() - t0
    finally:
        stop()
    for r in reqs:
        frame, seen = got[r]
        assert seen == W * H and np.abs(frame.astype(int) - want[r]).max() <= 1, r
    total = sum(W * H * (spp // 4) * 4 for _, spp in reqs)
    parity_log("gpu/server_concurrent_jobs", connections=len(reqs), samples=total, seconds=dt, msamples_per_s_through_sockets=total / dt / 1e6)
