"""Procedural stand-ins for the assets the reference repo does not ship (.MISSING_LARGE_BLOBS) and for the
synthetic benchmark configs of BASELINE.json: icosphere / "bunny-class" lumpy sphere, triangle soup,
RGBE environment with a sun, PBR texture maps.  Only +,-,*,/,sqrt and integer arithmetic are used so that the
same bytes come out on every machine (fixtures generated here must match scenes regenerated on the GPU box).
"""
import numpy as np


def icosphere(subdiv):
    """Unit icosphere: 20 * 4**subdiv triangles.  Returns (vertices (V,3) f64, faces (F,3) i64)."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v = v / np.sqrt((v * v).sum(axis=1))[:, None]
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        key = e[:, 0] * (v.shape[0] + 1) + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // (v.shape[0] + 1), uniq % (v.shape[0] + 1)
        mid = (v[a] + v[b]) * 0.5
        mid = mid / np.sqrt((mid * mid).sum(axis=1))[:, None]
        base = v.shape[0]
        v = np.concatenate([v, mid], axis=0)
        F = f.shape[0]
        m01, m12, m20 = base + inv[:F], base + inv[F:2 * F], base + inv[2 * F:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return v, f


def lumpy(vertices, amount=0.35):
    """Radial polynomial displacement: turns the sphere into a non-convex 'bunny-class' blob (ears, dents)."""
    x, y, z = vertices[:, 0], vertices[:, 1], vertices[:, 2]
    bump = (x * y * 2.0 + y * z * z * 3.0 - x * x * z * 2.5 + (y * y * y - 0.3 * y) * 2.0 + x * z * (x * x - z * z) * 4.0)
    return vertices * (1.0 + amount * bump)[:, None]


QUAD_VERTS = np.array([[0.5, 0.0, 0.5], [0.5, 0.0, -0.5], [-0.5, 0.0, -0.5], [-0.5, 0.0, 0.5]], np.float64)
QUAD_UVS = np.array([[0.0, 0.0], [0.0, 1.0], [1.0, 1.0], [1.0, 0.0]], np.float64)
QUAD_FACES = np.array([[0, 2, 1], [2, 0, 3]], np.int64)          # asset_packs/misc/top_mono.obj:11-12
QUAD_FACE_UVS = QUAD_UVS[QUAD_FACES]                             # `f 1/1 3/3 2/2`, `f 3/3 1/1 4/4`


def triangle_soup(n, seed=1234, extent=1.0, edge=(0.01, 0.05)):
    """n random small triangles, centres U[-extent,extent]^3 (BASELINE config 3/5).  (n,3,3) f64."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n, 1, 3))
    s = rng.uniform(edge[0], edge[1], (n, 1, 1))
    off = rng.uniform(-1.0, 1.0, (n, 3, 3))
    return c + off * s


def _rgbe_encode(rgb):
    """float RGB (H,W,3) -> RGBE8 such that decode rgb8/255 * 2^(e-128) (tracer.fs:412) reproduces it."""
    m = rgb.max(axis=2)
    _, ex = np.frexp(m)                       # m = f * 2^ex, f in [0.5,1)
    ex = np.where(m > 0, ex, -128).astype(np.int64)
    ex = np.clip(ex, -127, 127)
    scale = np.ldexp(1.0, -ex)[..., None]
    q = np.floor(rgb * scale * 255.0 + 0.5)
    out = np.empty(rgb.shape[:2] + (4,), np.uint8)
    out[..., :3] = np.clip(q, 0, 255).astype(np.uint8)
    out[..., 3] = (ex + 128).astype(np.uint8)
    return out


def environment(width=2048, height=1024, sun=(0.30, 0.28), sun_radius=0.012, sun_radiance=400.0):
    """Sky gradient + warm horizon + dark ground + one bright sun disc, as an RGBE-in-RGBA8 lat-long image
    (row 0 = +Y).  The sun makes ProcessEnvRadiance refine its bins (env_sampler.js:27-49)."""
    v = ((np.arange(height, dtype=np.float64) + 0.5) / height)[:, None]
    u = ((np.arange(width, dtype=np.float64) + 0.5) / width)[None, :]
    sky = np.clip(1.0 - v * 2.0, 0.0, 1.0)           # 1 at zenith .. 0 at horizon
    gnd = np.clip(v * 2.0 - 1.0, 0.0, 1.0)           # 0 at horizon .. 1 at nadir
    hor = 1.0 - np.clip(np.abs(v - 0.5) * 6.0, 0.0, 1.0)
    up = (v < 0.5).astype(np.float64)
    r = up * (0.25 + 0.15 * (1 - sky)) + (1 - up) * (0.18 - 0.10 * gnd) + hor * 0.55
    g = up * (0.45 + 0.20 * (1 - sky)) + (1 - up) * (0.15 - 0.08 * gnd) + hor * 0.45
    b = up * (0.95 - 0.10 * (1 - sky)) + (1 - up) * (0.10 - 0.05 * gnd) + hor * 0.30
    wob = 1.0 + 0.15 * (u * (1.0 - u) * 4.0 - 0.5)   # mild azimuthal variation
    rgb = np.stack([r * wob, g * wob, b * wob], axis=2)
    du = np.minimum(np.abs(u - sun[0]), 1.0 - np.abs(u - sun[0]))
    d2 = du * du + ((v - sun[1]) * 0.5) ** 2
    disc = (d2 < sun_radius * sun_radius).astype(np.float64)[..., None]
    halo = np.clip(1.0 - d2 / (16 * sun_radius * sun_radius), 0.0, 1.0)[..., None]
    rgb = rgb + disc * np.array([1.0, 0.9, 0.7]) * sun_radiance + halo * halo * np.array([1.0, 0.8, 0.5]) * 2.0
    return _rgbe_encode(rgb)


def constant_environment(width, height, value):
    rgb = np.empty((height, width, 3), np.float64)
    rgb[...] = value
    return _rgbe_encode(rgb)


def _tri_wave(x):
    f = x - np.floor(x)
    return 1.0 - np.abs(f * 2.0 - 1.0)


def pbr_maps(res=2048, seed=7, tag="A"):
    """Four RGBA8 maps (row 0 = image top) with the channel semantics the shader expects: baseColor (sRGB),
    metallicRoughness (.r metallic, .g roughness, tracer.fs:455-457), emissive, tangent-space normal."""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.arange(res, dtype=np.float64), np.arange(res, dtype=np.float64), indexing="ij")
    u, v = (x + 0.5) / res, (y + 0.5) / res
    tiles = 16
    tx, ty = np.floor(u * tiles).astype(np.int64), np.floor(v * tiles).astype(np.int64)
    pal = rng.uniform(0.15, 0.95, (tiles, tiles, 3))
    grout = ((_tri_wave(u * tiles) < 0.06) | (_tri_wave(v * tiles) < 0.06)).astype(np.float64)[..., None]
    base = pal[ty, tx] * (1 - grout) + 0.08 * grout
    base = base * (0.85 + 0.15 * _tri_wave(u * 97.0 + v * 31.0))[..., None]
    met_t = (rng.uniform(0, 1, (tiles, tiles)) < 0.3).astype(np.float64)
    rough_t = rng.uniform(0.05, 0.9, (tiles, tiles))
    mr = np.stack([met_t[ty, tx] * (1 - grout[..., 0]), rough_t[ty, tx] * (1 - grout[..., 0]) + 0.9 * grout[..., 0],
                   np.zeros_like(u)], axis=2)
    em_t = (rng.uniform(0, 1, (tiles, tiles)) < 0.04).astype(np.float64)
    core = ((_tri_wave(u * tiles) > 0.5) & (_tri_wave(v * tiles) > 0.5)).astype(np.float64)
    em = (em_t[ty, tx] * core)[..., None] * pal[ty, tx] * 0.6
    hx = _tri_wave(u * tiles * 4.0) - 0.5
    hy = _tri_wave(v * tiles * 4.0) - 0.5
    nx, ny = -hx * 0.6, -hy * 0.6
    nz = np.sqrt(np.clip(1.0 - nx * nx - ny * ny, 0.0, 1.0))
    nrm = np.stack([nx * 0.5 + 0.5, ny * 0.5 + 0.5, nz], axis=2)  # decode: (rgb-(.5,.5,0))*(2,2,1), tracer.fs:456

    def to8(a):
        o = np.empty((res, res, 4), np.uint8)
        o[..., :3] = np.floor(np.clip(a, 0, 1) * 255.0 + 0.5).astype(np.uint8)
        o[..., 3] = 255
        return o

    return {
        "baseColor": {"src": "procedural://%s/baseColor" % tag, "pixels": to8(base)},
        "metallicRoughness": {"src": "procedural://%s/metallicRoughness" % tag, "pixels": to8(mr)},
        "emissive": {"src": "procedural://%s/emissive" % tag, "pixels": to8(em)},
        "normal": {"src": "procedural://%s/normal" % tag, "pixels": to8(nrm)},
    }
