"""Triangle preparation: the part of obj_loader.js that turns indexed meshes into the reference's
`Triangle` records (verts, normals, tangents, bitangents, uvs), vectorised with numpy in float64
(JavaScript numbers are doubles).  Reference: obj_loader.js:19-38 (transforms), :40-52 (normals),
:62-101 (calcTangents), :103-162 (parseTriangle), :194-203 (smooth normals); vector.js:90-118.
"""
from dataclasses import dataclass, field
import math

import numpy as np

JS_EPSILON = 2.0 ** -52  # Number.EPSILON


def cross(a, b):
    """Vec3.cross (vector.js:113-118)."""
    x = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    y = -(a[..., 0] * b[..., 2] - a[..., 2] * b[..., 0])
    z = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    return np.stack([x, y, z], axis=-1)


def dot(a, b):
    """Vec3.dot (vector.js:43-45): left-to-right sum."""
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def normalize(v):
    """Vec3.normalize (vector.js:10-13): scale by 1/magnitude."""
    m = np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2])
    with np.errstate(divide="ignore", invalid="ignore"):
        return v * (1.0 / m)[..., None]


def rotate_arbitrary(v, axis, angle):
    """Vec3.rotateArbitrary (vector.js:90-102) applied to an (N,3) array."""
    x, y, z = (float(axis[0]), float(axis[1]), float(axis[2]))
    s, c = math.sin(angle), math.cos(angle)
    oc = 1.0 - c
    mat = [oc * x * x + c, oc * x * y - z * s, oc * z * x + y * s,
           oc * x * y + z * s, oc * y * y + c, oc * y * z - x * s,
           oc * z * x - y * s, oc * y * z + x * s, oc * z * z + c]
    out = np.empty_like(v)
    for i in range(3):  # matVecMultiply: dot(row, vec), left-to-right
        out[..., i] = mat[3 * i] * v[..., 0] + mat[3 * i + 1] * v[..., 1] + mat[3 * i + 2] * v[..., 2]
    return out


def apply_vector_transforms(v, transforms, world_transforms=None, rotation_only=False):
    """applyVectorTransforms (obj_loader.js:24-38): rotate -> scale -> translate -> world transforms."""
    v = np.asarray(v, dtype=np.float64)
    for r in transforms.get("rotate", []) or []:
        v = rotate_arbitrary(v, r["axis"], r["angle"])
    if not rotation_only:
        v = v * float(transforms.get("scale", 1.0)) + np.asarray(transforms.get("translate", [0, 0, 0]), np.float64)
    else:
        v = v * 1.0 + np.zeros(3)
    for t in world_transforms or []:
        if t.get("rotate"):
            for r in t["rotate"]:
                v = rotate_arbitrary(v, r["axis"], r["angle"])
        elif t.get("translate") and not rotation_only:
            v = v + np.asarray(t["translate"], np.float64)
    return v


@dataclass
class TriangleSet:
    """A batch of reference `Triangle`s (bvh.js:200-216) sharing one material record."""
    verts: np.ndarray       # (T,3,3) f64 world space
    normals: np.ndarray     # (T,3,3)
    tangents: np.ndarray    # (T,3,3)
    bitangents: np.ndarray  # (T,3,3)
    uvs: np.ndarray         # (T,3,2)
    material: dict = field(default_factory=dict)  # diffuseIndex, specularIndex, normalIndex, roughnessIndex, ior, dielectric, emittance

    @property
    def count(self):
        return self.verts.shape[0]


def mesh_to_triangles(vertices, faces, transforms, world_transforms=None, face_uvs=None, mesh_normals=None,
                      face_normal_idx=None):
    """parseTriangle + smooth normals + calcTangents for one OBJ group.

    vertices (V,3) object space, faces (F,3) zero-based vertex ids, face_uvs (F,3,2) or None (spherical
    mapping, obj_loader.js:63-70), transforms = the prop dict (rotate/scale/translate/normals).
    """
    vertices = np.asarray(vertices, np.float64)
    faces = np.asarray(faces, np.int64)
    wv = apply_vector_transforms(vertices, transforms, world_transforms)
    tv = wv[faces]  # (F,3,3)
    mode = transforms.get("normals")
    if mode == "mesh":
        mn = normalize(apply_vector_transforms(np.asarray(mesh_normals, np.float64), transforms, world_transforms, True))
        normals = mn[np.asarray(face_normal_idx, np.int64)]
    else:
        e1 = tv[:, 1] - tv[:, 0]
        e2 = tv[:, 2] - tv[:, 0]
        fn = normalize(cross(e1, e2))  # getNormal, obj_loader.js:40-44
        normals = np.repeat(fn[:, None, :], 3, axis=1)
        if mode == "smooth":  # averageNormals of the per-vertex lists, in triangle order (obj_loader.js:46-52,153-159,196-202)
            total = np.zeros((vertices.shape[0], 3))
            cnt = np.zeros(vertices.shape[0])
            flat_idx = faces.reshape(-1)
            np.add.at(total, flat_idx, np.repeat(fn, 3, axis=0))
            np.add.at(cnt, flat_idx, 1.0)
            with np.errstate(divide="ignore", invalid="ignore"):
                avg = total * (1.0 / cnt)[:, None]
            normals = avg[faces]
    # calcTangents, obj_loader.js:62-101
    if face_uvs is None:
        d = normalize(tv)
        u = np.arctan2(d[..., 2], d[..., 0]) / (math.pi * 2)
        v = np.arcsin(-d[..., 1]) / math.pi + 0.5
        uvs = np.stack([u, v], axis=-1)
    else:
        uvs = np.array(face_uvs, np.float64, copy=True)
    for i in range(3):
        uvs[:, i, :] += JS_EPSILON * (i + 1)
    dp0 = tv[:, 1] - tv[:, 0]
    dp1 = tv[:, 2] - tv[:, 0]
    du0 = uvs[:, 1] - uvs[:, 0]
    du1 = uvs[:, 2] - uvs[:, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = 1.0 / ((du0[:, 0] * du1[:, 1]) - (du0[:, 1] * du1[:, 0]))
        pre_t = normalize((dp0 * du1[:, 1:2] - dp1 * du0[:, 1:2]) * r[:, None])
        tangents = np.empty_like(tv)
        bitangents = np.empty_like(tv)
        bad = np.zeros((tv.shape[0], 3), bool)
        for i in range(3):
            n = normals[:, i]
            pre_b = normalize(cross(n, pre_t))
            t = normalize(cross(pre_b, n))
            b = normalize(cross(n, t))
            tangents[:, i] = t
            bitangents[:, i] = b
            bad[:, i] = np.isnan(dot(t, b))
    # the reference's NaN fallback assigns by index and then pushes anyway (obj_loader.js:93-99); emulate the
    # resulting list contents literally for the (rare) affected triangles
    for f in np.nonzero(bad.any(axis=1))[0]:
        tl, bl = [], []
        for i in range(3):
            n = normals[f, i]
            if bad[f, i]:
                t = cross(n, np.array([0.0, 1.0, 0.0]))
                b = cross(t, n)
                for lst, val in ((tl, t), (bl, b)):
                    if len(lst) > i:
                        lst[i] = val
                    else:
                        while len(lst) < i:
                            lst.append(np.full(3, np.nan))
                        lst.append(val)
            tl.append(tangents[f, i].copy())
            bl.append(bitangents[f, i].copy())
        for i in range(3):
            tangents[f, i] = tl[i]
            bitangents[f, i] = bl[i]
    return TriangleSet(tv, normals, tangents, bitangents, uvs)
