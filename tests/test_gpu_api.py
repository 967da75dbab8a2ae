"""C-ABI behaviour on the GPU: error convention, checkpoint/resume, odd resolutions, sum mode."""
import types

import numpy as np
import pytest

from fspt_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def _frame(ctx, cam):
    return ctx.frame(cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), cam["env_theta"])


def test_error_convention(small_bunny):
    sa, cam = small_bunny
    ctx = capi.Context(64, 48)
    try:
        with pytest.raises(capi.FsptError) as e:  # render before scene upload
            ctx.render(_frame(ctx, cam), 0, [1.0], [2.0])
        assert e.value.code == -3 and "scene_upload" in str(e.value)
        bad = types.SimpleNamespace(**{k: getattr(sa, k) for k in ("bvh", "tris", "mats", "norms", "uvs", "atlas", "env", "bins")})
        bad.leaf_size = 8
        with pytest.raises(capi.FsptError) as e:
            ctx.scene_upload(bad)
        assert e.value.code == -1 and "LEAF_SIZE" in str(e.value)
        # a chain deeper than the reference's int stack[64] (tracer.fs:368) is refused, not silently truncated
        n = 70
        bvh = np.zeros((2 * n + 1, 9), np.float32)
        hdr = bvh.view(np.int32)
        for i in range(n):            # interior i: left = leaf 2i+1 ... laid out as [interior, leaf, interior, leaf ...]
            hdr[2 * i, 0], hdr[2 * i, 1], hdr[2 * i, 2] = 2 * i + 1, 2 * i + 2, -1
            hdr[2 * i + 1, 2] = 0
        hdr[2 * n, 2] = 0
        bvh[:, 3:6], bvh[:, 6:9] = -1.0, 1.0
        deep = types.SimpleNamespace(bvh=bvh, tris=sa.tris[:4], mats=sa.mats[:4], norms=sa.norms[:4], uvs=sa.uvs[:4],
                                     atlas=sa.atlas, env=sa.env, bins=sa.bins, leaf_size=4)
        with pytest.raises(capi.FsptError) as e:
            ctx.scene_upload(deep)
        assert e.value.code == -4 and "stack" in str(e.value)
        cyc = bvh[:3].copy()
        cyc.view(np.int32)[0, :3] = (0, 0, -1)  # node 0 is its own child
        with pytest.raises(capi.FsptError):
            ctx.scene_upload(types.SimpleNamespace(bvh=cyc, tris=sa.tris[:4], mats=sa.mats[:4], norms=sa.norms[:4],
                                                   uvs=sa.uvs[:4], atlas=sa.atlas, env=sa.env, bins=sa.bins, leaf_size=4))
        ctx.scene_upload(sa)  # the context is still usable after errors
        ctx.render(_frame(ctx, cam), 0, [1.0], [2.0])
        assert np.isfinite(ctx.read_accum()).all()
    finally:
        ctx.close()
    with pytest.raises(capi.FsptError):
        capi.Context(64, 48, device=99)


def test_checkpoint_resume_and_scene_swap(small_bunny, oracle_mod):
    sa, cam = small_bunny
    W, H, N = 72, 40, 5  # 40 rows / 72 columns: tiled path ordering
    rc, rt = scenes.rand_bases(N, 8)
    a = capi.Context(W, H)
    b = capi.Context(W, H)
    try:
        a.scene_upload(sa)
        a.render(_frame(a, cam), 0, rc, rt)
        full = a.read_accum()
        # stop after 2 samples, move the accumulator to another context, continue there
        a.clear()
        a.render(_frame(a, cam), 0, rc[:2], rt[:2])
        b.scene_upload(sa)
        b.write_accum(a.read_accum(), 2)
        b.render(_frame(b, cam), 2, rc[2:], rt[2:])
        assert np.array_equal(b.read_accum().view(np.uint32), full.view(np.uint32))
        # uploading a different scene into the same context reuses buffers and gives that scene's image
        sb, camb = scenes.quad_scene()
        a.scene_upload(sb)
        a.clear()
        a.render(_frame(a, camb), 0, rc[:1], rt[:1])
        O = oracle_mod.Oracle(sb)
        pos, d = oracle_mod.camera(W, H, camb["eye"], camb["dir"], camb["fov_scale"], scenes.lens_features(camb), rc[0])
        ref, _ = O.trace(pos, d, W, H, 0, rt[0], camb["env_theta"])
        assert np.array_equal(a.read_accum()[..., :3].view(np.uint32), ref[..., :3].view(np.uint32))
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("res", [(50, 37), (8, 4), (1, 1), (33, 64)])
def test_odd_resolutions_bit_exact(small_bunny, oracle_mod, res):
    sa, cam = small_bunny
    W, H = res
    O = oracle_mod.Oracle(sa)
    rc, rt = scenes.rand_bases(2, 4)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        ctx.render(_frame(ctx, cam), 0, rc, rt)
        fb = None
        for k in range(2):
            pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
            fb, _ = O.trace(pos, d, W, H, k, rt[k], cam["env_theta"], fb_prev=fb)
        assert np.array_equal(ctx.read_accum()[..., :3].view(np.uint32), fb[..., :3].view(np.uint32))
        assert np.array_equal(ctx.resolve(denoise=True), oracle_mod.draw(fb, denoise=True))
    finally:
        ctx.close()


def test_sum_mode_single_gpu(small_bunny, oracle_mod):
    sa, cam = small_bunny
    W, H, N = 64, 48, 4
    O = oracle_mod.Oracle(sa)
    rc, rt = scenes.rand_bases(N, 6)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        ctx.set_accum_mode(1)
        ctx.render(_frame(ctx, cam), 0, rc, rt)
        ref = np.zeros((H, W, 3), np.float32)
        for k in range(N):
            pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
            _, col, _ = O.trace(pos, d, W, H, 0, rt[k], cam["env_theta"], want_color=True)
            ref = ref + col[..., :3]
        assert np.array_equal(ctx.read_accum()[..., :3].view(np.uint32), ref.view(np.uint32))
        ptr, n, s = ctx.accum_device_ptr()
        assert ptr and n == W * H * 4 and s == N
    finally:
        ctx.close()


@pytest.mark.parametrize("wave,n", [("3", 10), (None, 70)])
def test_renders_longer_than_one_wave_bit_exact(small_bunny, oracle_mod, monkeypatch, wave, n):
    """A render call with more samples than fit in flight is cut into waves (3+3+3+1 samples here, 64+6 with the
    default sizing); the running mean must continue across the cuts exactly as tracer.fs:517 does per pass."""
    sa, cam = small_bunny
    W, H = 40, 24
    if wave:
        monkeypatch.setenv("FSPT_WAVE_SAMPLES", wave)
    O = oracle_mod.Oracle(sa)
    rc, rt = scenes.rand_bases(n, 9)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        ctx.render(_frame(ctx, cam), 0, rc, rt)
        fb, rays = None, 0
        for k in range(n):
            pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
            fb, _, st = O.trace(pos, d, W, H, k, rt[k], cam["env_theta"], fb_prev=fb, want_color=True)
            rays += st["rays"]
        assert np.array_equal(ctx.read_accum()[..., :3].view(np.uint32), fb[..., :3].view(np.uint32))
        assert ctx.stats()["last_rays"] == rays
    finally:
        ctx.close()


def test_two_triangle_scene_root_is_a_leaf(oracle_mod):
    """<= 4 triangles: the BVH root is a leaf (bvh.js:22) and the traversal starts on a leaf reference."""
    sa, cam = scenes.quad_scene()
    W, H = 48, 32
    O = oracle_mod.Oracle(sa)
    rc, rt = scenes.rand_bases(3, 2)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        fr = _frame(ctx, cam)
        idx, t, cnt, pos, d = ctx.debug_primary(fr, rc[0])
        oidx, ot, ocnt, _ = O.bvh_test(pos, d)
        assert np.array_equal(idx, oidx) and np.array_equal(t.view(np.uint32), ot.view(np.uint32)) and np.array_equal(cnt, ocnt)
        assert int(cnt.max()) == 1
        ctx.render(fr, 0, rc, rt)
        fb = None
        for k in range(3):
            pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
            fb, _ = O.trace(pos, d, W, H, k, rt[k], cam["env_theta"], fb_prev=fb)
        assert np.array_equal(ctx.read_accum()[..., :3].view(np.uint32), fb[..., :3].view(np.uint32))
    finally:
        ctx.close()


def test_plain_atlas_fallback_is_bit_identical(monkeypatch):
    """fspt_scene_upload falls back to the plain RGBA8 layered atlas when the material-interleaved one would not fit;
    both sampling routines must produce the same bits (FSPT_PLAIN_ATLAS forces the fallback)."""
    sa, cam = scenes.pbr_scene(atlas_res=64, subdiv=2, env_size=(128, 64))
    W, H = 96, 64
    rc, rt = scenes.rand_bases(4, 13)
    out = []
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("FSPT_PLAIN_ATLAS", "1")
        ctx = capi.Context(W, H)
        try:
            ctx.scene_upload(sa)
            ctx.render(_frame(ctx, cam), 0, rc, rt)
            out.append(ctx.read_accum().copy())
        finally:
            ctx.close()
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
    assert float(out[0][..., :3].max()) > 0.0


def test_interleaved_atlas_allocation_failure_falls_back_to_the_plain_atlas(monkeypatch):
    """When the material-interleaved array (or its pinned staging) cannot be allocated the upload must not fail: it takes
    the plain-atlas branch on the SAME context, after an interleaved scene was resident (FSPT_FORCE_MAT_TEX=0 makes every
    interleaved allocation 'fail')."""
    sa, cam = scenes.pbr_scene(atlas_res=64, subdiv=2, env_size=(128, 64))
    sb, _ = scenes.pbr_scene(atlas_res=32, subdiv=2, env_size=(128, 64))   # another atlas size: forces a re-allocation
    W, H = 96, 64
    rc, rt = scenes.rand_bases(3, 15)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sb)
        ctx.render(_frame(ctx, cam), 0, rc, rt)
        ctx.scene_upload(sa)
        ctx.clear()
        ctx.render(_frame(ctx, cam), 0, rc, rt)
        ref = ctx.read_accum().copy()
        ctx.scene_upload(sb)
        monkeypatch.setenv("FSPT_FORCE_MAT_TEX", "0")
        ctx.scene_upload(sa)            # interleaved array "cannot be allocated" -> plain RGBA8 layers
        ctx.clear()
        ctx.render(_frame(ctx, cam), 0, rc, rt)
        assert np.array_equal(ref.view(np.uint32), ctx.read_accum().view(np.uint32))
    finally:
        ctx.close()


def test_gpu_and_host_atlas_interleave_are_bit_identical(monkeypatch):
    """The material-interleaved atlas is built on the host (default for scenes whose materials use most of their four
    maps) or by k_interleave_atlas from the raw layers (chosen when few maps vary); FSPT_ATLAS_INTERLEAVE forces one."""
    sa, cam = scenes.pbr_scene(atlas_res=64, subdiv=2, env_size=(128, 64))
    W, H = 96, 64
    rc, rt = scenes.rand_bases(4, 21)
    out = []
    for mode in ("cpu", "gpu", "cpu"):  # the third upload re-creates the array without the surface flag
        monkeypatch.setenv("FSPT_ATLAS_INTERLEAVE", mode)
        if not out:
            ctx = capi.Context(W, H)
        ctx.scene_upload(sa)
        ctx.clear()
        ctx.render(_frame(ctx, cam), 0, rc, rt)
        out.append(ctx.read_accum().copy())
    ctx.close()
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
    assert np.array_equal(out[0].view(np.uint32), out[2].view(np.uint32))
    assert float(out[0][..., :3].max()) > 0.0


def test_tiles_rendered_one_after_the_other_equal_the_whole_frame(small_bunny, oracle_mod):
    """fspt_set_tile: a pixel's samples do not depend on which rectangle contains it (same gl_FragCoord, same
    resolution, same rand bases), so rendering the frame as bands -- even unaligned ones that use the row-major path
    ordering -- reproduces the whole-frame accumulation bit for bit, in both accumulation modes."""
    sa, cam = small_bunny
    W, H, N = 64, 48, 3
    rc, rt = scenes.rand_bases(N, 12)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        fr = _frame(ctx, cam)
        for mode in (0, 1):
            ctx.set_accum_mode(mode)
            ctx.set_tile(0, 0, W, H)
            ctx.clear()
            ctx.render(fr, 0, rc, rt)
            whole = ctx.read_accum()
            ctx.clear()
            for rect in [(0, 0, W, 16), (0, 16, 24, 32), (24, 16, 40, 13), (24, 29, 40, 19)]:
                ctx.set_tile(*rect)
                ctx.render(fr, 0, rc, rt)
            tiled = ctx.read_accum()
            assert np.array_equal(whole.view(np.uint32), tiled.view(np.uint32)), mode
        with pytest.raises(capi.FsptError):
            ctx.set_tile(0, 0, W + 1, H)
        # the oracle agrees with the tiled result as well (running mean)
        ctx.set_accum_mode(0)
        ctx.set_tile(0, 0, W, H)
        O = oracle_mod.Oracle(sa)
        fb = None
        for k in range(N):
            pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
            fb, _ = O.trace(pos, d, W, H, k, rt[k], cam["env_theta"], fb_prev=fb)
        ctx.clear()
        for rect in [(0, 0, W, 20), (0, 20, W, 28)]:
            ctx.set_tile(*rect)
            ctx.render(fr, 0, rc, rt)
        assert np.array_equal(ctx.read_accum()[..., :3].view(np.uint32), fb[..., :3].view(np.uint32))
    finally:
        ctx.close()


def test_back_to_back_renders_do_not_wait_and_stay_exact(small_bunny, oracle_mod):
    """fspt_render is asynchronous across calls (rand bases are staged through a ring of pinned slots): twenty
    single-sample calls enqueued without any synchronisation equal one twenty-sample call bit for bit, and a
    fspt_debug_primary right behind a render does not disturb the render's rand bases."""
    sa, cam = small_bunny
    W, H, N = 48, 32, 20
    rc, rt = scenes.rand_bases(N, 14)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        fr = _frame(ctx, cam)
        ctx.render(fr, 0, rc, rt)
        one = ctx.read_accum()
        ctx.clear()
        for k in range(N):
            ctx.render(fr, k, rc[k:k + 1], rt[k:k + 1])
        ctx.debug_primary(fr, 123.0)
        many = ctx.read_accum()
        assert np.array_equal(one.view(np.uint32), many.view(np.uint32))
    finally:
        ctx.close()


def test_collectives_need_a_communicator(small_bunny):
    sa, _ = small_bunny
    ctx = capi.Context(32, 16)
    try:
        ctx.scene_upload(sa)
        for call in (ctx.reduce_accum, ctx.scene_broadcast):
            with pytest.raises(capi.FsptError) as e:
                call(0)
            assert e.value.code == -3  # FSPT_E_STATE
    finally:
        ctx.close()


@pytest.mark.parametrize("mode", ["cpu", "gpu", "plain"])
def test_asynchronous_upload_equals_the_synchronous_one(monkeypatch, mode):
    """fspt_scene_upload_async returns before the atlas has been staged; a render enqueued right behind it starts its
    primary traversal at once and waits for the staging before its first shading launch.  Same bits as the synchronous
    upload for every staging path (host interleave, GPU interleave, plain layers), across scene swaps without a render
    in between, and with the source atlas overwritten as soon as fspt_scene_upload_wait has returned."""
    if mode == "plain":
        monkeypatch.setenv("FSPT_PLAIN_ATLAS", "1")
    else:
        monkeypatch.setenv("FSPT_ATLAS_INTERLEAVE", mode)
    sa, cam = scenes.pbr_scene(atlas_res=256, subdiv=2, env_size=(128, 64))
    sb, _ = scenes.pbr_scene(atlas_res=64, subdiv=2, env_size=(128, 64))
    W, H = 96, 64
    rc, rt = scenes.rand_bases(3, 31)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        fr = _frame(ctx, cam)
        ctx.render(fr, 0, rc, rt)
        ref = ctx.read_accum().copy()
        assert float(ref[..., :3].max()) > 0.0
        for _ in range(3):
            ctx.scene_upload(sb, wait=False)       # replaced before anything used it: the next upload joins the staging
            ctx.scene_upload(sa, wait=False)
            ctx.clear()
            ctx.render(fr, 0, rc, rt)              # joins the staging thread behind the primary traversal launch
            assert np.array_equal(ref.view(np.uint32), ctx.read_accum().view(np.uint32))
        # the borrow ends at upload_wait: scribbling over a COPY of the atlas afterwards must not reach the device
        import copy
        sc = copy.copy(sa)
        sc.atlas = np.array(sa.atlas, copy=True)
        ctx.scene_upload(sc, wait=False)
        ctx.upload_wait()
        sc.atlas[...] = 0
        ctx.clear()
        ctx.render(fr, 0, rc, rt)
        assert np.array_equal(ref.view(np.uint32), ctx.read_accum().view(np.uint32))
    finally:
        ctx.close()


@pytest.mark.parametrize("slots,threads", [("1", "4"), ("3", "16"), (None, "5")])
def test_geometry_ring_reuse_and_worker_counts_are_bit_identical(small_bunny, monkeypatch, slots, threads):
    """The geometry records are staged chunk by chunk through a ring of pinned slots by the context's host workers;
    a slot is rewritten only after the copies that read it have completed.  Forcing the ring down to one or three slots
    (every chunk then reuses a slot) and changing the number of workers must not change a bit of the render."""
    sa, cam = small_bunny
    W, H = 64, 48
    rc, rt = scenes.rand_bases(2, 41)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        fr = _frame(ctx, cam)
        ctx.render(fr, 0, rc, rt)
        ref = ctx.read_accum().copy()
        if slots:
            monkeypatch.setenv("FSPT_RING_SLOTS", slots)
        monkeypatch.setenv("FSPT_UPLOAD_THREADS", threads)
        for wait in (True, False):
            ctx.scene_upload(sa, wait=wait)
            ctx.clear()
            ctx.render(fr, 0, rc, rt)
            assert np.array_equal(ref.view(np.uint32), ctx.read_accum().view(np.uint32)), (slots, threads, wait)
    finally:
        ctx.close()


def test_page_locked_atlas_and_env_are_uploaded_in_place(monkeypatch):
    """fspt_host_register: a page-locked atlas is DMA'd from where it lies (distinct non-constant layers only, per-material
    texels built by k_interleave_atlas) and a page-locked environment likewise -- same bits as the staged upload, in the
    synchronous and the asynchronous call, and again after the buffers have been unlocked."""
    import copy
    sa0, cam = scenes.pbr_scene(atlas_res=128, subdiv=2, env_size=(128, 64))
    sa = copy.copy(sa0)
    sa.atlas = np.array(sa0.atlas, copy=True)
    sa.env = np.array(sa0.env, copy=True)
    W, H = 96, 64
    rc, rt = scenes.rand_bases(3, 51)
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa0)
        fr = _frame(ctx, cam)
        ctx.render(fr, 0, rc, rt)
        ref = ctx.read_accum().copy()
        assert float(ref[..., :3].max()) > 0.0
        capi.host_register(sa.atlas)
        capi.host_register(sa.env)
        capi.host_register(sa.env)      # registering twice is not an error
        try:
            for wait in (True, False, False):
                ctx.scene_upload(sa, wait=wait)
                ctx.clear()
                ctx.render(fr, 0, rc, rt)
                assert np.array_equal(ref.view(np.uint32), ctx.read_accum().view(np.uint32)), wait
            ctx.upload_wait()
            # the borrow of a page-locked buffer ends when its copies have COMPLETED: at the return of the synchronous
            # call, at fspt_scene_upload_wait for the asynchronous one -- scribbling over it then must not reach the device
            for wait in (True, False):
                sa.atlas[...] = sa0.atlas
                sa.env[...] = sa0.env
                ctx.scene_upload(sa, wait=wait)
                if not wait:
                    ctx.upload_wait()
                sa.atlas[...] = 0
                sa.env[...] = 0
                ctx.clear()
                ctx.render(fr, 0, rc, rt)
                assert np.array_equal(ref.view(np.uint32), ctx.read_accum().view(np.uint32)), ("scribble", wait)
            sa.atlas[...] = sa0.atlas
            sa.env[...] = sa0.env
        finally:
            ctx.synchronize()
            capi.host_unregister(sa.atlas)
            capi.host_unregister(sa.env)
        ctx.scene_upload(sa, wait=False)
        ctx.clear()
        ctx.render(fr, 0, rc, rt)
        assert np.array_equal(ref.view(np.uint32), ctx.read_accum().view(np.uint32))
    finally:
        ctx.close()
