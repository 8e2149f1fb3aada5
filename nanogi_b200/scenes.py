"""Synthetic scene generators for the BASELINE.json configs.

A scene "spec" is a list of primitive dicts in the vocabulary of the reference's schema.yml
(type list, mesh, params). `to_scene_data` flattens a spec straight into the POD arrays of the C ABI;
`write_scene_files` writes the same spec as `scene.yml` + OBJ meshes so it can also go through the C++
front end / CLI exactly like a reference scene file.

  cornell_box()        C1: the Cornell box (38 triangles, D walls + one [L, D] area light, pinhole).
                       Geometry is the published Cornell box data, as in the reference's fixture
                       utils/runc/data/cornelbox/ (which is not copied into this repository).
  cornell_spheres()    C2: Cornell box + a G (conductor, roughness 0.1) and an S-fresnel icosphere.
  instanced_spheres()  C3: ~1.0 M triangles, 48 flattened icospheres (subdiv 5) in a room, D/G/S mix, 4 quad lights.
  interior()           C4: ~10 M triangle nave/aisle/column interior with 256 small quad lights.
  furnace()            known-answer scene: closed box, every wall [L, D] with Le = 1 and albedo rho.
  light_over_plane()   known-answer scene: one quad light over a diffuse plane.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import capi

Spec = List[dict]


# ---- mesh helpers -----------------------------------------------------------------------------
def quads_to_tris(quads: np.ndarray) -> np.ndarray:
    """[n,4,3] -> [2n,3,3], fan from the first vertex (0,1,2),(0,2,3) like aiProcess_Triangulate on convex quads."""
    q = np.asarray(quads, dtype=np.float64).reshape(-1, 4, 3)
    t = np.empty((q.shape[0] * 2, 3, 3), np.float64)
    t[0::2] = q[:, [0, 1, 2]]
    t[1::2] = q[:, [0, 2, 3]]
    return t


def flat_normals(tris: np.ndarray) -> np.ndarray:
    t = np.asarray(tris, dtype=np.float64)
    n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-300)
    return np.repeat(n[:, None, :], 3, axis=1)


def icosphere(subdiv: int) -> Tuple[np.ndarray, np.ndarray]:
    """Unit icosphere: returns (tris [20*4^s,3,3], smooth normals). Outward CCW winding."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    tris = v[f]
    for _ in range(subdiv):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab = a + b; ab /= np.linalg.norm(ab, axis=1, keepdims=True)
        bc = b + c; bc /= np.linalg.norm(bc, axis=1, keepdims=True)
        ca = c + a; ca /= np.linalg.norm(ca, axis=1, keepdims=True)
        tris = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1),
                               np.stack([ab, bc, ca], 1)], axis=0)
    return tris, tris.copy()


def box_quads(lo, hi, inward: bool) -> np.ndarray:
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    q = np.array([
        [[x0, y0, z0], [x1, y0, z0], [x1, y0, z1], [x0, y0, z1]],   # floor  (normal +y when inward)
        [[x0, y1, z0], [x0, y1, z1], [x1, y1, z1], [x1, y1, z0]],   # ceiling (normal -y)
        [[x0, y0, z0], [x0, y0, z1], [x0, y1, z1], [x0, y1, z0]],   # x = x0 (normal +x)
        [[x1, y0, z0], [x1, y1, z0], [x1, y1, z1], [x1, y0, z1]],   # x = x1 (normal -x)
        [[x0, y0, z0], [x0, y1, z0], [x1, y1, z0], [x1, y0, z0]],   # z = z0 (normal +z)
        [[x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]],   # z = z1 (normal -z)
    ], dtype=np.float64)
    # as listed the normals point INTO the box for floor: cross((x1-x0,0,0),(x1-x0,0,z1-z0)) = (0,-,0)?  fix below
    tris = quads_to_tris(q)
    n = flat_normals(tris)[:, 0]
    c = tris.mean(axis=1)
    centre = (np.asarray(lo, dtype=np.float64) + np.asarray(hi, dtype=np.float64)) * 0.5
    points_in = np.einsum("ij,ij->i", n, centre - c) > 0
    flip = points_in != inward
    q2 = q.copy()
    flipq = flip[0::2]
    q2[flipq] = q2[flipq][:, ::-1]
    return q2


def mesh_prim(types: List[str], tris: np.ndarray, normals: Optional[np.ndarray] = None, name: str = "mesh", uv: Optional[np.ndarray] = None,
              **params) -> dict:
    """`uv`: [nTri, 3, 2] texture coordinates; D / G params may carry "TexR": float array [H, W, 3] (row 0 = top) instead of "R"."""
    tris = np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)
    if normals is None:
        normals = flat_normals(tris)
    mesh = {"name": name, "tris": tris, "normals": np.asarray(normals, dtype=np.float64)}
    if uv is not None:
        mesh["uv"] = np.asarray(uv, dtype=np.float64).reshape(-1, 3, 2)
    return {"type": types, "mesh": mesh, "params": params}


def pinhole(eye, center, up, fov_deg: float) -> dict:
    return {"type": ["E"], "mesh": None, "params": {"E": {"type": "pinhole", "We": [1, 1, 1], "eye": list(eye), "center": list(center),
                                                          "up": list(up), "fov": float(fov_deg)}}}


def area_sensor(quad, we=(1.0, 1.0, 1.0), name: str = "sensor") -> dict:
    """E.area ("raw") sensor, reference rt.hpp:412-420 / :1908-1928: a quad [4, 3] (corners in uv order (0,0) (1,0) (1,1) (0,1)) whose uv
    coordinates ARE the raster position (rt.hpp:1386-1391); it looks along its geometric normal (p1-p0) x (p3-p0)."""
    q = np.asarray(quad, dtype=np.float64).reshape(4, 3)
    tris = np.array([[q[0], q[1], q[2]], [q[0], q[2], q[3]]])
    uv = np.array([[[0, 0], [1, 0], [1, 1]], [[0, 0], [1, 1], [0, 1]]], dtype=np.float64)
    pr = mesh_prim(["E"], tris, name=name, uv=uv, E={"type": "area", "We": list(we)})
    return pr


# ---- spec -> C ABI ----------------------------------------------------------------------------
def to_scene_data(spec: Spec, aspect: float = 1.0, name: str = "scene") -> capi.SceneData:
    pos, nrm, uvs, prims, textures = [], [], [], [], []
    any_uv = any(pr.get("mesh") is not None and "uv" in pr["mesh"] for pr in spec)
    first = 0

    def tex_index(img):
        for i, t in enumerate(textures):
            if t is img:
                return i
        textures.append(img)
        return len(textures) - 1

    for pr in spec:
        tbits = 0
        for s in pr["type"]:
            tbits |= {"D": capi.TYPE_D, "G": capi.TYPE_G, "S": capi.TYPE_S, "L": capi.TYPE_L, "E": capi.TYPE_E}[s]
        kw: Dict = {"type": tbits}
        if pr.get("mesh") is not None:
            t = np.asarray(pr["mesh"]["tris"], dtype=np.float32)
            n = np.asarray(pr["mesh"]["normals"], dtype=np.float32)
            kw["first_tri"] = first
            kw["num_tris"] = int(t.shape[0])
            first += int(t.shape[0])
            pos.append(t)
            nrm.append(n)
            if any_uv:
                uvs.append(np.asarray(pr["mesh"].get("uv", np.zeros((t.shape[0], 3, 2))), dtype=np.float32))
        P = pr["params"]
        if "L" in P:
            L = P["L"]
            kw["l_type"] = {"area": capi.L_AREA, "point": capi.L_POINT, "directional": capi.L_DIRECTIONAL}[L["type"]]
            kw["l_le"] = L["Le"]
            if L["type"] == "point":
                kw["l_vec"] = L["position"]
            if L["type"] == "directional":
                kw["l_vec"] = L["direction"]
        if "E" in P:
            E = P["E"]
        if "E" in P and P["E"]["type"] == "area":
            kw.update(e_type=capi.E_AREA, e_we=P["E"].get("We", [1, 1, 1]), e_aspect=float(aspect))
        elif "E" in P:
            eye, center, up = (np.asarray(E[k], dtype=np.float64) for k in ("eye", "center", "up"))
            vz = eye - center
            vz /= np.linalg.norm(vz)
            vx = np.cross(up, vz)
            vx /= np.linalg.norm(vx)
            vy = np.cross(vz, vx)
            kw.update(e_type=capi.E_PINHOLE, e_position=eye, e_vx=vx, e_vy=vy, e_vz=vz, e_fov=math.radians(E["fov"]),
                      e_aspect=float(aspect), e_we=E.get("We", [1, 1, 1]))
        if "D" in P:
            if "TexR" in P["D"]:
                kw["d_tex"] = tex_index(P["D"]["TexR"])
            else:
                kw["d_r"] = P["D"]["R"]
        if "G" in P:
            kw.update(g_eta=P["G"]["Eta"], g_k=P["G"]["K"], g_roughness=float(P["G"]["Roughness"]))
            if "TexR" in P["G"]:
                kw["g_tex"] = tex_index(P["G"]["TexR"])
            else:
                kw["g_r"] = P["G"]["R"]
        if "S" in P:
            S = P["S"]
            kw["s_type"] = {"reflection": capi.S_REFLECTION, "refraction": capi.S_REFRACTION, "fresnel": capi.S_FRESNEL}[S["type"]]
            kw["s_r"] = S["R"]
            if S["type"] != "reflection":
                kw["s_eta1"] = float(S["eta1"])
                kw["s_eta2"] = float(S["eta2"])
        prims.append(capi.make_prim(**kw))
    positions = np.concatenate(pos, axis=0) if pos else np.zeros((0, 3, 3), np.float32)
    normals = np.concatenate(nrm, axis=0) if nrm else np.zeros((0, 3, 3), np.float32)
    texcoords = np.concatenate(uvs, axis=0) if (any_uv and uvs) else None
    return capi.SceneData(positions, normals, prims, texcoords, name=name, textures=[np.asarray(t, dtype=np.float32) for t in textures])


def write_pfm(path: str, img: np.ndarray) -> None:
    """Little-endian colour PFM (rows bottom-up in the file); `img` is [H, W, 3] with row 0 = top."""
    img = np.asarray(img, dtype="<f4")
    with open(path, "wb") as f:
        f.write(f"PF\n{img.shape[1]} {img.shape[0]}\n-1.0\n".encode())
        f.write(img[::-1].tobytes())


def write_texture(directory: str, index: int, lobe: str, img) -> str:
    name = f"{index:03d}_{lobe}_tex.pfm"
    write_pfm(os.path.join(directory, name), img)
    return name


def write_scene_files(spec: Spec, directory: str, version: int = 5) -> str:
    """Writes scene.yml + one OBJ per mesh primitive (v / vn [/ vt] / f) + one PFM per TexR texture. Returns the YAML path."""
    os.makedirs(directory, exist_ok=True)
    lines = [f"version: {version}", "scene:", "  primitives:"]

    def vec(v):
        return "[" + ", ".join(repr(float(x)) for x in v) + "]"

    for i, pr in enumerate(spec):
        lines.append(f"    - type: [{', '.join(pr['type'])}]")
        if pr.get("mesh") is not None:
            fname = f"{i:03d}_{pr['mesh'].get('name', 'mesh')}.obj"
            t = np.asarray(pr["mesh"]["tris"], dtype=np.float32).reshape(-1, 3)
            n = np.asarray(pr["mesh"]["normals"], dtype=np.float32).reshape(-1, 3)
            with open(os.path.join(directory, fname), "w") as f:
                f.write(f"# generated by nanogi_b200.scenes\no {pr['mesh'].get('name', 'mesh')}\n")
                np.savetxt(f, t, fmt="v %.9g %.9g %.9g")
                np.savetxt(f, n, fmt="vn %.9g %.9g %.9g")
                idx = np.arange(1, t.shape[0] + 1).reshape(-1, 3)
                if "uv" in pr["mesh"]:
                    np.savetxt(f, np.asarray(pr["mesh"]["uv"], dtype=np.float32).reshape(-1, 2), fmt="vt %.9g %.9g")
                    np.savetxt(f, np.repeat(idx, 3, axis=1), fmt="f %d/%d/%d %d/%d/%d %d/%d/%d")
                else:
                    np.savetxt(f, np.repeat(idx, 2, axis=1), fmt="f %d//%d %d//%d %d//%d")
            lines += ["      mesh:", f"        path: '{fname}'"]
        lines.append("      params:")
        P = pr["params"]
        if "L" in P:
            L = P["L"]
            lines += ["        L:", f"          type: {L['type']}", f"          {L['type']}:", f"            Le: {vec(L['Le'])}"]
            if L["type"] == "point":
                lines.append(f"            position: {vec(L['position'])}")
            if L["type"] == "directional":
                lines.append(f"            direction: {vec(L['direction'])}")
        if "E" in P and P["E"]["type"] == "area":
            lines += ["        E:", "          type: area", "          area:", f"            We: {vec(P['E'].get('We', [1, 1, 1]))}"]
        elif "E" in P:
            E = P["E"]
            lines += ["        E:", "          type: pinhole", "          pinhole:", f"            We: {vec(E.get('We', [1, 1, 1]))}",
                      "            view:", f"              eye: {vec(E['eye'])}", f"              center: {vec(E['center'])}",
                      f"              up: {vec(E['up'])}", "            perspective:", f"              fov: {float(E['fov'])!r}"]
        if "D" in P:
            lines += ["        D:", f"          R: {vec(P['D']['R'])}" if "R" in P["D"] else f"          TexR: '{write_texture(directory, i, 'D', P['D']['TexR'])}'"]
        if "G" in P:
            G = P["G"]
            lines += ["        G:", f"          R: {vec(G['R'])}" if "R" in G else f"          TexR: '{write_texture(directory, i, 'G', G['TexR'])}'",
                      f"          Eta: {vec(G['Eta'])}", f"          K: {vec(G['K'])}", f"          Roughness: {float(G['Roughness'])!r}"]
        if "S" in P:
            S = P["S"]
            lines += ["        S:", f"          type: {S['type']}", f"          {S['type']}:", f"            R: {vec(S['R'])}"]
            if S["type"] != "reflection":
                lines += [f"            eta1: {float(S['eta1'])!r}", f"            eta2: {float(S['eta2'])!r}"]
        lines.append("")
    path = os.path.join(directory, "scene.yml")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


# ---- C1: Cornell box --------------------------------------------------------------------------
COPPER = {"R": [1, 0.64705882352, 0.60784313725], "Eta": [0.14, 0.129, 0.1585], "K": [4.58625, 3.348125, 2.329375]}


def _cornell_prims(light_le=10.0) -> Spec:
    light = [[[343, 548.75, 227], [343, 548.75, 332], [213, 548.75, 332], [213, 548.75, 227]]]
    back = [[[549.6, 0, 559.2], [0, 0, 559.2], [0, 548.8, 559.2], [556, 548.8, 559.2]]]
    ceiling = [[[556, 548.8, 0], [556, 548.8, 559.2], [0, 548.8, 559.2], [0, 548.8, 0]],
               [[213, 548.8, 227], [213, 548.8, 332], [343, 548.8, 332], [343, 548.8, 227]]]
    floor = [[[552.8, 0, 0], [0, 0, 0], [0, 0, 559.2], [549.6, 0, 559.2]]]
    red = [[[552.8, 0, 0], [549.6, 0, 559.2], [556, 548.8, 559.2], [556, 548.8, 0]]]
    green = [[[0, 0, 559.2], [0, 0, 0], [0, 548.8, 0], [0, 548.8, 559.2]]]
    tall = [[[423, 330, 247], [265, 330, 296], [314, 330, 456], [472, 330, 406]],
            [[423, 0, 247], [423, 330, 247], [472, 330, 406], [472, 0, 406]],
            [[472, 0, 406], [472, 330, 406], [314, 330, 456], [314, 0, 456]],
            [[314, 0, 456], [314, 330, 456], [265, 330, 296], [265, 0, 296]],
            [[265, 0, 296], [265, 330, 296], [423, 330, 247], [423, 0, 247]],
            [[472, 0, 406], [314, 0, 456], [265, 0, 296], [423, 0, 247]]]
    short = [[[130, 165, 65], [82, 165, 225], [240, 165, 272], [290, 165, 114]],
             [[290, 0, 114], [290, 165, 114], [240, 165, 272], [240, 0, 272]],
             [[130, 0, 65], [130, 165, 65], [290, 165, 114], [290, 0, 114]],
             [[82, 0, 225], [82, 165, 225], [130, 165, 65], [130, 0, 65]],
             [[240, 0, 272], [240, 165, 272], [82, 165, 225], [82, 0, 225]],
             [[290, 0, 114], [240, 0, 272], [82, 0, 225], [130, 0, 65]]]
    W = [1, 1, 1]
    return [
        mesh_prim(["L", "D"], quads_to_tris(light), name="light", L={"type": "area", "Le": [light_le] * 3}, D={"R": [0, 0, 0]}),
        mesh_prim(["D"], quads_to_tris(back), name="back", D={"R": W}),
        mesh_prim(["D"], quads_to_tris(ceiling), name="ceiling", D={"R": W}),
        mesh_prim(["D"], quads_to_tris(floor), name="floor", D={"R": W}),
        mesh_prim(["D"], quads_to_tris(red), name="redwall", D={"R": [1, 0, 0]}),
        mesh_prim(["D"], quads_to_tris(green), name="greenwall", D={"R": [0, 1, 0]}),
        mesh_prim(["D"], quads_to_tris(tall), name="largebox", D={"R": W}),
        mesh_prim(["D"], quads_to_tris(short), name="smallbox", D={"R": W}),
    ]


CORNELL_CAMERA = dict(eye=[278, 273, -800], center=[278, 273, -799], up=[0, 1, 0], fov_deg=39.3077)


def cornell_box() -> Spec:
    return _cornell_prims() + [pinhole(**CORNELL_CAMERA)]


def cornell_spheres() -> Spec:
    """C2: Cornell box + a glossy conductor icosphere on the short block and a glass (S fresnel) icosphere."""
    tris, nrm = icosphere(3)  # 1280 triangles each, like the reference's Icosphere fixtures
    spec = _cornell_prims()
    g_c, g_r = np.array([186.0, 165.0 + 80.0, 169.0]), 80.0
    s_c, s_r = np.array([400.0, 90.0, 120.0]), 90.0
    spec.append(mesh_prim(["G"], tris * g_r + g_c, nrm, name="glossy", G=dict(COPPER, Roughness=0.1)))
    spec.append(mesh_prim(["S"], tris * s_r + s_c, nrm, name="glass",
                          S={"type": "fresnel", "R": [0.60784313725, 0.80392156862, 1], "eta1": 1.0, "eta2": 2.0}))
    spec.append(pinhole(**CORNELL_CAMERA))
    return spec


def cornell_raw_sensor(spheres: bool = False) -> Spec:
    """Cornell box seen by an E.area ("raw") sensor instead of the pinhole: a 200 x 200 quad just inside the open front,
    facing the back wall (the reference authors' verification device for lt / ltdirect / bdpt, rt.hpp:412-420)."""
    spec = [p for p in (cornell_spheres() if spheres else cornell_box()) if "E" not in p["params"]]
    z, lo, hi = 20.0, 178.0, 378.0
    spec.append(area_sensor([[lo, lo, z], [hi, lo, z], [hi, hi, z], [lo, hi, z]]))
    return spec


def cornell_mixed_lights() -> Spec:
    """C2's scene lit by all three light kinds: the area light, a point light and a directional light (inert in pt / ptdirect,
    rt.hpp:934-937; a path origin in lt / ltdirect, rt.hpp:549-562)."""
    spec = cornell_spheres()
    cam = spec.pop()
    spec.append({"type": ["L"], "mesh": None, "params": {"L": {"type": "point", "Le": [3.0e4, 2.0e4, 1.0e4], "position": [120.0, 400.0, 300.0]}}})
    spec.append({"type": ["L"], "mesh": None, "params": {"L": {"type": "directional", "Le": [2.0, 2.0, 3.0], "direction": [0.3, -0.5, 0.81240384]}}})
    spec.append(cam)
    return spec


def checker_texture(n: int = 8, res: int = 64, a=(0.9, 0.2, 0.2), b=(0.2, 0.3, 0.9)) -> np.ndarray:
    y, x = np.mgrid[0:res, 0:res]
    m = ((x * n // res) + (y * n // res)) % 2
    return np.where(m[..., None] == 0, np.array(a, np.float32), np.array(b, np.float32)).astype(np.float32)


def cornell_textured() -> Spec:
    """the Cornell box with a TexR checker on the floor (D) and on the glossy sphere (G); uv of the floor
    span [0, 2]^2 (exercises the fract wrap), of the sphere a lat-long map."""
    spec = cornell_spheres()
    tex_d, tex_g = checker_texture(8, 64), checker_texture(6, 32, (1.0, 0.8, 0.5), (0.6, 0.6, 0.7))
    for pr in spec:
        m = pr.get("mesh")
        if m is None:
            continue
        if m["name"] == "floor":
            t = m["tris"]
            lo, hi = t.reshape(-1, 3).min(0), t.reshape(-1, 3).max(0)
            m["uv"] = np.stack([(t[..., 0] - lo[0]) / (hi[0] - lo[0]) * 2, (t[..., 2] - lo[2]) / (hi[2] - lo[2]) * 2], axis=-1)
            pr["params"]["D"] = {"TexR": tex_d}
        if m["name"] == "glossy":
            t = m["tris"]
            c = t.reshape(-1, 3).mean(0)
            d = t - c
            d /= np.linalg.norm(d, axis=-1, keepdims=True)
            m["uv"] = np.stack([np.arctan2(d[..., 2], d[..., 0]) / (2 * math.pi) + 0.5, np.arccos(np.clip(d[..., 1], -1, 1)) / math.pi], axis=-1)
            G = dict(pr["params"]["G"])
            G.pop("R")
            G["TexR"] = tex_g
            pr["params"]["G"] = G
    return spec


# ---- known-answer scenes ----------------------------------------------------------------------
def furnace(rho: float = 0.5, le: float = 1.0) -> Spec:
    """Closed box, every wall [L, D] with Le = le and grey albedo rho, camera inside.
    pt and ptdirect converge to le * sum_k rho^k (= le/(1-rho) with -m -1)."""
    q = box_quads([-1, -1, -1], [1, 1, 1], inward=True)
    spec = []
    for i in range(6):
        spec.append(mesh_prim(["L", "D"], quads_to_tris(q[i:i + 1]), name=f"wall{i}", L={"type": "area", "Le": [le] * 3}, D={"R": [rho] * 3}))
    spec.append(pinhole(eye=[0.1, 0.05, 0.2], center=[0.3, 0.1, -1.0], up=[0, 1, 0], fov_deg=60))
    return spec


def light_over_plane(le: float = 5.0, height: float = 1.0, half: float = 0.5, rho: float = 0.8) -> Spec:
    """One square light (side 2*half, facing down) at y = height over a large diffuse floor. A narrow-fov camera
    below the light looks at the floor point under the light's centre, where the radiance with -m 3 is
    rho * Le * F,  F = 4/pi * X * atan(X),  X = half / sqrt(half^2 + height^2)  (parallel-square form factor)."""
    light = [[[half, height, -half], [half, height, half], [-half, height, half], [-half, height, -half]]]
    floor = [[[-20, 0, -20], [-20, 0, 20], [20, 0, 20], [20, 0, -20]]]
    return [
        mesh_prim(["L", "D"], quads_to_tris(light), name="light", L={"type": "area", "Le": [le] * 3}, D={"R": [0, 0, 0]}),
        mesh_prim(["D"], quads_to_tris(floor), name="floor", D={"R": [rho] * 3}),
        pinhole(eye=[2.0, 0.8 * height, 0.0], center=[0, 0, 0], up=[0, 1, 0], fov_deg=2),
    ]


# ---- branch coverage: the primitive kinds no BASELINE scene contains -----------------------------------------------
def cornell_branches(light_res: int = 0) -> Spec:
    """Cornell box whose blocks are replaced by primitives that exercise the branches C1-C4 never take:
      * a [D, G] and a [G, S] primitive — the lobe is resolved by precedence D > G > S (rt.hpp:808-859: the `if (type & D) ... else if
        (type & G) ...` chains; src/nanogi.cpp:601 strips only the emitter bits), so the first renders as D, the second as G;
      * S.reflection (a mirror block) and S.refraction (a sphere), rt.hpp:808-859, :1066-1098;
      * a pure [L] mesh: an emitter with no BSDF — a path that hits it from outside a direct-light connection finds no lobe to
        sample and ends there (rt.hpp:909, :1146, :1334);
      * light_res > 0: the ceiling light tessellated into 2 * light_res^2 triangles (131 072 at 256): a large area CDF
        (basic.hpp:440-497, rt.hpp:1747-1765)."""
    spec = [p for p in _cornell_prims() if p["mesh"]["name"] not in ("largebox", "smallbox", "light")]
    lx0, lx1, lz0, lz1, ly = 213.0, 343.0, 227.0, 332.0, 548.75
    if light_res > 0:
        xs = np.linspace(lx1, lx0, light_res + 1); zs = np.linspace(lz0, lz1, light_res + 1)
        X, Z = np.meshgrid(xs, zs, indexing="ij")
        P = np.stack([X, np.full_like(X, ly), Z], axis=-1)
        q = np.stack([P[:-1, :-1], P[:-1, 1:], P[1:, 1:], P[1:, :-1]], axis=2).reshape(-1, 4, 3)
        light = quads_to_tris(q)
    else:
        light = quads_to_tris([[[lx1, ly, lz0], [lx1, ly, lz1], [lx0, ly, lz1], [lx0, ly, lz0]]])
    spec.insert(0, mesh_prim(["L", "D"], light, name="light", L={"type": "area", "Le": [10.0] * 3}, D={"R": [0, 0, 0]}))
    tris, nrm = icosphere(2)
    spec.append(mesh_prim(["D", "G"], tris * 60.0 + np.array([120.0, 60.0, 150.0]), nrm, name="dg", D={"R": [0.2, 0.7, 0.9]}, G=dict(COPPER, Roughness=0.2)))
    spec.append(mesh_prim(["G", "S"], tris * 60.0 + np.array([420.0, 60.0, 150.0]), nrm, name="gs", G=dict(COPPER, Roughness=0.15),
                          S={"type": "fresnel", "R": [1, 1, 1], "eta1": 1.0, "eta2": 1.5}))
    spec.append(mesh_prim(["S"], tris * 70.0 + np.array([278.0, 70.0, 330.0]), nrm, name="refr",
                          S={"type": "refraction", "R": [0.95, 0.95, 1.0], "eta1": 1.0, "eta2": 1.5}))
    mirror = box_quads(np.array([60.0, 0.0, 380.0]), np.array([200.0, 300.0, 520.0]), inward=False)
    spec.append(mesh_prim(["S"], quads_to_tris(mirror), name="mirror", S={"type": "reflection", "R": [0.9, 0.9, 0.9]}))
    # a small emissive quad on the floor with NO BSDF lobe at all
    ex0, ex1, ez0, ez1, ey = 380.0, 470.0, 380.0, 470.0, 0.5
    spec.append(mesh_prim(["L"], quads_to_tris([[[ex0, ey, ez0], [ex0, ey, ez1], [ex1, ey, ez1], [ex1, ey, ez0]]]), name="bare_light",
                          L={"type": "area", "Le": [4.0, 3.0, 2.0]}))
    spec.append(pinhole(**CORNELL_CAMERA))
    return spec


# ---- C3: ~1M-triangle flattened "instanced" spheres -------------------------------------------
def instanced_spheres(seed: int = 1, subdiv: int = 5, grid=(4, 4, 3)) -> Spec:
    rng = np.random.default_rng(seed)
    base, base_n = icosphere(subdiv)
    room_lo, room_hi = np.array([-10.0, 0.0, -10.0]), np.array([10.0, 9.0, 10.0])
    spec: Spec = []
    # 4 quad lights just below the ceiling, facing down
    for i, (cx, cz) in enumerate([(-5, -5), (5, -5), (-5, 5), (5, 5)]):
        y, h = 8.95, 1.0
        q = [[[cx + h, y, cz - h], [cx + h, y, cz + h], [cx - h, y, cz + h], [cx - h, y, cz - h]]]
        spec.append(mesh_prim(["L", "D"], quads_to_tris(q), name=f"light{i}", L={"type": "area", "Le": [20, 20, 20]}, D={"R": [0, 0, 0]}))
    walls = box_quads(room_lo, room_hi, inward=True)
    cols = [[0.75, 0.75, 0.75], [0.75, 0.75, 0.75], [0.7, 0.25, 0.25], [0.25, 0.7, 0.25], [0.75, 0.75, 0.75], [0.75, 0.75, 0.75]]
    for i in range(6):
        spec.append(mesh_prim(["D"], quads_to_tris(walls[i:i + 1]), name=f"wall{i}", D={"R": cols[i]}))
    # 16-triangle floor detail: 8 low slabs (top quads)
    det = []
    for k in range(8):
        x0 = -9 + k * 2.3
        det.append([[x0, 0.05, -9.5], [x0, 0.05, -8.5], [x0 + 1.5, 0.05, -8.5], [x0 + 1.5, 0.05, -9.5]])
    spec.append(mesh_prim(["D"], quads_to_tris(det), name="floordetail", D={"R": [0.5, 0.5, 0.6]}))
    gx, gy, gz = grid
    idx = 0
    for ix in range(gx):
        for iz in range(gy):
            for iy in range(gz):
                cell = np.array([(ix + 0.5) / gx, (iy + 0.5) / gz, (iz + 0.5) / gy])
                c = room_lo + cell * (room_hi - room_lo) * np.array([1, 0.85, 1]) + rng.uniform(-0.4, 0.4, 3)
                r = rng.uniform(0.9, 1.3)
                tris = base * r + c
                kind = idx % 3
                if kind == 0:
                    spec.append(mesh_prim(["D"], tris, base_n, name=f"sphere{idx}", D={"R": list(rng.uniform(0.2, 0.8, 3))}))
                elif kind == 1:
                    spec.append(mesh_prim(["G"], tris, base_n, name=f"sphere{idx}", G=dict(COPPER, Roughness=float([0.05, 0.1, 0.3][(idx // 3) % 3]))))
                else:
                    spec.append(mesh_prim(["S"], tris, base_n, name=f"sphere{idx}", S={"type": "fresnel", "R": [1, 1, 1], "eta1": 1.0, "eta2": 1.5}))
                idx += 1
    spec.append(pinhole(eye=[0, 4.5, 9.5], center=[0, 4.0, 0], up=[0, 1, 0], fov_deg=45))
    return spec


# ---- C4: ~10M-triangle interior ----------------------------------------------------------------
def _displaced_grid(origin, du, dv, nu: int, nv: int, amp: float, rng) -> np.ndarray:
    """A (nu x nv)-quad grid spanning origin + s*du + t*dv, displaced along its normal by smooth noise."""
    origin, du, dv = (np.asarray(a, dtype=np.float64) for a in (origin, du, dv))
    n = np.cross(du, dv)
    n /= np.linalg.norm(n)
    s = np.linspace(0, 1, nu + 1)
    t = np.linspace(0, 1, nv + 1)
    S, T = np.meshgrid(s, t, indexing="ij")
    ph = rng.uniform(0, 6.28, 4)
    disp = amp * (np.sin(37 * S + ph[0]) * np.sin(29 * T + ph[1]) + 0.5 * np.sin(113 * S + ph[2]) * np.sin(97 * T + ph[3]))
    disp[0, :] = disp[-1, :] = 0
    disp[:, 0] = disp[:, -1] = 0
    P = origin + S[..., None] * du + T[..., None] * dv + disp[..., None] * n
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    q = np.stack([a, b, c, d], axis=2).reshape(-1, 4, 3)
    return quads_to_tris(q)


def interior(seed: int = 2, target_tris: int = 10_600_000, n_lights: int = 256) -> Spec:
    """Nave + aisles + columns (Sibenik-like layout); walls are displaced subdivided grids to reach target_tris."""
    rng = np.random.default_rng(seed)
    L, Wd, H = 60.0, 24.0, 16.0
    spec: Spec = []
    # lights: small quads under the ceiling, each its own [L, D] primitive (uniform light pick)
    side = int(math.ceil(math.sqrt(n_lights)))
    k = 0
    for i in range(side):
        for j in range(side):
            if k >= n_lights:
                break
            cx = -L / 2 + (i + 0.5) * L / side
            cz = -Wd / 2 + (j + 0.5) * Wd / side
            y, h = H - 0.3, 0.15
            q = [[[cx + h, y, cz - h], [cx + h, y, cz + h], [cx - h, y, cz + h], [cx - h, y, cz - h]]]
            spec.append(mesh_prim(["L", "D"], quads_to_tris(q), name=f"light{k}", L={"type": "area", "Le": [400, 380, 340]}, D={"R": [0, 0, 0]}))
            k += 1
    n_cols = 16
    # budget: 6 room surfaces + columns (each an 8-sided prism of grids)
    col_tris = target_tris // 4
    wall_tris = target_tris - col_tris
    per_wall = wall_tris // 6
    res = max(2, int(math.sqrt(per_wall / 2)))
    lo, hi = np.array([-L / 2, 0, -Wd / 2]), np.array([L / 2, H, Wd / 2])
    surfaces = [  # origin, du, dv chosen so the normal (du x dv) points into the room
        (lo, [0, 0, Wd], [L, 0, 0], [0.6, 0.55, 0.5]),                      # floor  (+y)
        ([lo[0], hi[1], lo[2]], [L, 0, 0], [0, 0, Wd], [0.7, 0.7, 0.7]),     # ceiling (-y)
        (lo, [0, H, 0], [0, 0, Wd], [0.65, 0.6, 0.55]),                      # x = lo (+x)
        ([hi[0], lo[1], lo[2]], [0, 0, Wd], [0, H, 0], [0.65, 0.6, 0.55]),   # x = hi (-x)
        (lo, [L, 0, 0], [0, H, 0], [0.6, 0.6, 0.65]),                        # z = lo (+z)
        ([lo[0], lo[1], hi[2]], [0, H, 0], [L, 0, 0], [0.6, 0.6, 0.65]),     # z = hi (-z)
    ]
    for i, (o, du, dv, colr) in enumerate(surfaces):
        tris = _displaced_grid(o, du, dv, res, res, 0.05, rng)
        if i % 3 == 2:
            spec.append(mesh_prim(["G"], tris, name=f"wall{i}", G=dict(COPPER, R=[0.9, 0.9, 0.9], Roughness=0.3)))
        else:
            spec.append(mesh_prim(["D"], tris, name=f"wall{i}", D={"R": colr}))
    per_col = col_tris // n_cols
    sides = 8
    cres = max(2, int(math.sqrt(2 * per_col / sides)))   # each side is a cres x cres/4 grid = cres^2 / 2 triangles
    for c in range(n_cols):
        cx = -L / 2 + (c // 2 + 0.5) * L / (n_cols // 2)
        cz = -Wd / 4 if c % 2 == 0 else Wd / 4
        r = 0.9
        parts = []
        for s in range(sides):
            a0, a1 = 2 * math.pi * s / sides, 2 * math.pi * (s + 1) / sides
            p0 = np.array([cx + r * math.cos(a0), 0, cz + r * math.sin(a0)])
            p1 = np.array([cx + r * math.cos(a1), 0, cz + r * math.sin(a1)])
            parts.append(_displaced_grid(p0, [0, H - 0.6, 0], p1 - p0, cres, max(2, cres // 4), 0.02, rng))
        spec.append(mesh_prim(["D"], np.concatenate(parts, axis=0), name=f"column{c}", D={"R": [0.7, 0.68, 0.6]}))
    spec.append(pinhole(eye=[-L / 2 + 2, 5.0, 0.5], center=[L / 2, 6.0, 0], up=[0, 1, 0], fov_deg=60))
    return spec


# ---- C5: ray batches --------------------------------------------------------------------------
def camera_rays(scene: capi.SceneData, width: int, height: int, tmin: float = 1e-4, tmax: float = 3.4028234663852886e38) -> np.ndarray:
    """Coherent batch: one pinhole ray through every pixel centre (SampleDirection E.pinhole, rt.hpp:733-740)."""
    p = scene.prims[scene.sensor_prim()]
    vx, vy, vz = (np.array(list(v)) for v in (p.e_vx, p.e_vy, p.e_vz))
    tan_f = math.tan(p.e_fov * 0.5)
    aspect = width / height
    xs = (np.arange(width) + 0.5) / width * 2 - 1
    ys = (np.arange(height) + 0.5) / height * 2 - 1
    X, Y = np.meshgrid(xs, ys)
    d_eye = np.stack([aspect * tan_f * X, tan_f * Y, -np.ones_like(X)], axis=-1).reshape(-1, 3)
    d_eye /= np.linalg.norm(d_eye, axis=1, keepdims=True)
    d = d_eye[:, 0:1] * vx + d_eye[:, 1:2] * vy + d_eye[:, 2:3] * vz
    rays = np.zeros(d.shape[0], capi.RAY_DTYPE)
    rays["o"] = np.array(list(p.e_position), dtype=np.float32)
    rays["d"] = d.astype(np.float32)
    rays["tmin"] = tmin
    rays["tmax"] = tmax
    return rays


def random_rays(scene: capi.SceneData, n: int, seed: int = 3, occlusion: bool = False) -> np.ndarray:
    """Incoherent batch: origins uniform in the scene AABB, directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    lo = scene.positions.reshape(-1, 3).min(axis=0).astype(np.float64)
    hi = scene.positions.reshape(-1, 3).max(axis=0).astype(np.float64)
    rays = np.zeros(n, capi.RAY_DTYPE)
    rays["o"] = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    z = 1 - 2 * rng.random(n)
    r = np.sqrt(np.maximum(0, 1 - z * z))
    ph = 2 * math.pi * rng.random(n)
    rays["d"] = np.stack([r * np.cos(ph), r * np.sin(ph), z], axis=1).astype(np.float32)
    rays["tmin"] = 1e-4
    if occlusion:
        rays["tmax"] = (rng.uniform(0.1, 1.0, n) * np.linalg.norm(hi - lo)).astype(np.float32)
    else:
        rays["tmax"] = np.float32(3.4028234663852886e38)
    return rays
