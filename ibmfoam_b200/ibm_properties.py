"""constant/IBMProperties: reader, writer and generator (the data format either side of the IB path).

Format (read by particleCloud.C:989-1005,1034-1036,1070-1117 and particleIBM.C:279-297, written by
applications/utilities/initializeIBM: particleSizeDistribution.C:74-138, monodisperse.C:124-300):

    nParticles <label>;  moving <Switch>;  rotating <Switch>;  coupled <Switch>;  e <scalar>;
    flow { UInf <vector>; [rho rhoInf; rhoInf <scalar>;]  [p pRef; pRef <scalar>;] }
    particle.<i> { particleShape sphere|cylinder|ellipse;  position <vector>;  d <scalar> | a <scalar> b <scalar>;
                   nTheta <label>;  nk <label>;  delta <scalar>;  rho <scalar>;  [theta <vector>] [v <vector>] [omega <vector>] }

`read_ibm_properties` turns such a file into the (CloudSpec, ParticleTable) pair the C ABI takes; `write_ibm_properties`
writes one the reference's solver reads; `generate_monodisperse` is the `initializeIBM` pre-processor (random sequential
placement with the reference's non-overlap rule |xi - xj| > ri + rj, monodisperse.C:162-178) with a reproducible seed -
the reference seeds from the wall clock (particleSizeDistribution.C:57).  Host-side, runs once per case: pure Python.
"""
from __future__ import annotations

import math
import re

import numpy as np

from . import cases

_SHAPES = {"sphere": cases.SPHERE, "cylinder": cases.CYLINDER, "ellipse": cases.ELLIPSE}
_SHAPE_NAMES = {v: k for k, v in _SHAPES.items()}
_TRUE = {"true", "on", "yes", "y", "t"}
_FALSE = {"false", "off", "no", "n", "f", "none"}


class FoamDictError(ValueError):
    """what the reference reports through FatalIOError: a missing keyword or a malformed entry"""


def _tokenize(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return re.findall(r"[{}();]|\"[^\"]*\"|[^\s{}();]+", text)


def parse_foam_dict(text):
    """Minimal OpenFOAM dictionary parser: `key value;`, `key (a b c);`, `name { ... }`.  Returns nested dicts whose leaves
    are strings or lists of strings (lists = parenthesised values).  The FoamFile header is returned like any sub-dict."""
    toks = _tokenize(text)
    pos = 0

    def parse_list():
        nonlocal pos
        out = []
        while True:
            if pos >= len(toks):
                raise FoamDictError("unterminated list")
            t = toks[pos]
            pos += 1
            if t == ")":
                return out
            out.append(parse_list() if t == "(" else t)

    def parse_block(top):
        nonlocal pos
        d = {}
        while pos < len(toks):
            key = toks[pos]
            pos += 1
            if key == "}":
                if top:
                    raise FoamDictError("unbalanced '}'")
                return d
            if key in "{();":
                raise FoamDictError(f"unexpected '{key}'")
            if pos >= len(toks):
                raise FoamDictError(f"entry '{key}' is not terminated")
            if toks[pos] == "{":
                pos += 1
                d[key] = parse_block(False)
                continue
            vals = []
            while True:
                if pos >= len(toks):
                    raise FoamDictError(f"entry '{key}' is not terminated by ';'")
                t = toks[pos]
                pos += 1
                if t == ";":
                    break
                vals.append(parse_list() if t == "(" else t)
            d[key] = vals[0] if len(vals) == 1 else vals
        if not top:
            raise FoamDictError("unterminated '{'")
        return d

    return parse_block(True)


def _lookup(d, key, where):
    if key not in d:
        raise FoamDictError(f"keyword {key} is undefined in dictionary \"{where}\"")   # the reference's FatalIOError
    return d[key]


def _switch(v, key):
    s = str(v).lower()
    if s in _TRUE:
        return True
    if s in _FALSE:
        return False
    raise FoamDictError(f"bad Switch value '{v}' for {key}")


def _vector(v, key):
    if not isinstance(v, list) or len(v) != 3:
        raise FoamDictError(f"entry {key} is not a vector")
    return [float(x) for x in v]


def read_ibm_properties(text, g=(0.0, 0.0, 0.0)):
    """-> (CloudSpec, ParticleTable).  Missing mandatory keys raise FoamDictError (the reference aborts on them:
    `coupled`, `e`, `moving`, `rotating`, `nParticles`, and per particle `particleShape`, `position`, `rho`, `nTheta`,
    `delta`, `d` | `a`,`b`).  Keys the code never reads (`rhoRef`, `interpolationSchemes`, ... App. B #16) are ignored."""
    d = parse_foam_dict(text)
    top = "IBMProperties"
    n = int(_lookup(d, "nParticles", top))
    flow = d.get("flow")
    if not isinstance(flow, dict):
        raise FoamDictError(f"keyword flow is undefined in dictionary \"{top}\"")
    cloud = cases.CloudSpec(
        moving=_switch(_lookup(d, "moving", top), "moving"), rotating=_switch(_lookup(d, "rotating", top), "rotating"),
        coupled=_switch(_lookup(d, "coupled", top), "coupled"), e=float(_lookup(d, "e", top)),
        UInf=tuple(_vector(flow["UInf"], "UInf")) if "UInf" in flow else (0.0, 0.0, 0.0),
        rhoRef=float(flow.get("rhoInf", 1.0)), pRef=float(flow.get("pInf", 0.0)), g=tuple(float(x) for x in g))
    kind, pos, dd, aa, bb, nT, nk, dl, th, vv, om, rho = ([] for _ in range(12))
    for i in range(n):
        name = f"particle.{i}"
        p = d.get(name)
        if not isinstance(p, dict):
            raise FoamDictError(f"keyword {name} is undefined in dictionary \"{top}\"")
        shape = str(_lookup(p, "particleShape", name))
        if shape not in _SHAPES:
            raise FoamDictError(f"Unknown particle shape type {shape}; valid particleShapes are: {sorted(_SHAPES)}")
        k = _SHAPES[shape]
        kind.append(k)
        pos.append(_vector(_lookup(p, "position", name), "position"))
        if k == cases.ELLIPSE:
            aa.append(float(_lookup(p, "a", name))); bb.append(float(_lookup(p, "b", name))); dd.append(0.0)
        else:
            dd.append(float(_lookup(p, "d", name))); aa.append(0.0); bb.append(0.0)
        nT.append(int(_lookup(p, "nTheta", name)))
        nk.append(int(p.get("nk", 1)))
        dl.append(float(_lookup(p, "delta", name)))
        rho.append(float(_lookup(p, "rho", name)))
        th.append(_vector(p["theta"], "theta") if "theta" in p else [0.0, 0.0, 0.0])
        vv.append(_vector(p["v"], "v") if "v" in p else [0.0, 0.0, 0.0])
        om.append(_vector(p["omega"], "omega") if "omega" in p else [0.0, 0.0, 0.0])
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    table = cases.ParticleTable(kind=np.asarray(kind, np.int32), position=f64(pos).reshape(n, 3), d=f64(dd), a=f64(aa), b=f64(bb),
                                nTheta=np.asarray(nT, np.int32), nk=np.asarray(nk, np.int32), delta=f64(dl),
                                theta=f64(th).reshape(n, 3), v=f64(vv).reshape(n, 3), omega=f64(om).reshape(n, 3), rho=f64(rho))
    return cloud, table


_HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
  =========                 |
  \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox
   \\\\    /   O peration     | Website:  https://openfoam.org
    \\\\  /    A nd           | Version:  6
     \\\\/     M anipulation  |
\\*---------------------------------------------------------------------------*/
FoamFile
{
    version     2.0;
    format      ascii;
    class       dictionary;
    location    "constant";
    object      IBMProperties;
}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def _num(x):
    return repr(float(x)) if float(x) != int(x) or abs(x) >= 1e15 else str(int(x))


def _vec(v):
    return "(" + " ".join(_num(x) for x in v) + ")"


def write_ibm_properties(cloud: "cases.CloudSpec", table: "cases.ParticleTable") -> str:
    """Text of constant/IBMProperties in the layout initializeIBM writes (one `key<TAB>value;` per line).  Numbers are
    written with repr(): reading the file back gives the same doubles."""
    sw = lambda b: "true" if b else "false"
    out = [_HEADER, f"nParticles\t{table.n};\n", f"moving\t{sw(cloud.moving)};\n", f"rotating\t{sw(cloud.rotating)};\n",
           f"coupled\t{sw(cloud.coupled)};\n", f"e\t{_num(cloud.e)};\n", "flow\n{\n", f"UInf\t{_vec(cloud.UInf)};\n",
           "rho\trhoInf;\n", f"rhoInf\t{_num(cloud.rhoRef)};\n"]
    if cloud.pRef != 0.0:
        out.append(f"pInf\t{_num(cloud.pRef)};\n")
    out.append("}\n")
    for i in range(table.n):
        k = int(table.kind[i])
        out.append(f"particle.{i}\n{{\nparticleShape\t{_SHAPE_NAMES[k]};\nposition\t{_vec(table.position[i])};\n")
        if k == cases.ELLIPSE:
            out.append(f"a\t{_num(table.a[i])};\nb\t{_num(table.b[i])};\n")
        else:
            out.append(f"d\t{_num(table.d[i])};\n")
        out.append(f"nTheta\t{int(table.nTheta[i])};\nnk\t{int(table.nk[i])};\ndelta\t{_num(table.delta[i])};\nrho\t{_num(table.rho[i])};\n")
        for key, arr in (("theta", table.theta), ("v", table.v), ("omega", table.omega)):
            if np.any(arr[i] != 0.0):
                out.append(f"{key}\t{_vec(arr[i])};\n")
        out.append("}\n")
    out.append("\n")
    return "".join(out)


def generate_monodisperse(box_lo, box_hi, d, *, n_particles=None, volume_fraction=None, shape="sphere", rho=1500.0, nTheta=50,
                          nk=1, delta=None, seed=12345, periodic=True, velocity_sigma=0.0, omega_sigma=0.0, length=None):
    """The `initializeIBM` pre-processor for one monodisperse particle type (monodisperse.C:55-300): the particle count from
    `nParticles` or from `volumeFraction` (count = fraction * box volume / particle volume, :64-110), positions by random
    sequential placement with the non-overlap rule of :162-178, optional Gaussian v and omega.  Returns a ParticleTable."""
    lo, hi = np.asarray(box_lo, float), np.asarray(box_hi, float)
    if (n_particles is None) == (volume_fraction is None):
        raise ValueError("give exactly one of n_particles / volume_fraction")
    kind = _SHAPES[shape]
    if kind == cases.ELLIPSE:
        raise ValueError("monodisperse placement uses one diameter: spheres and cylinders only")
    if n_particles is None:
        if kind == cases.SPHERE:
            vp, vbox = math.pi * d ** 3 / 6.0, float(np.prod(hi - lo))
        else:
            L = float(length if length is not None else hi[2] - lo[2])
            vp, vbox = math.pi * (d / 2.0) ** 2 * L, float(np.prod((hi - lo)[:2])) * L
        n_particles = int(round(volume_fraction * vbox / vp))
    if kind == cases.SPHERE:
        pos = cases.place_spheres(n_particles, lo, hi, d / 2.0, seed=seed, periodic=periodic)
    else:  # cylinders: placement in the plane, centre on the mid-plane z (the reference's 2-D cases)
        zc = 0.5 * (lo[2] + hi[2])
        pts, rng = [], np.random.default_rng(seed)
        tries = 0
        while len(pts) < n_particles:
            tries += 1
            if tries > 10_000_000:
                raise RuntimeError("cylinder placement did not converge")
            p = rng.random(2) * (hi - lo)[:2] + lo[:2]
            if not periodic and (np.any(p - d / 2 < lo[:2]) or np.any(p + d / 2 > hi[:2])):
                continue
            ok = True
            for q in pts:
                dv = p - q
                if periodic:
                    dv -= (hi - lo)[:2] * np.round(dv / (hi - lo)[:2])
                if not math.hypot(dv[0], dv[1]) > d:
                    ok = False
                    break
            if ok:
                pts.append(p)
        pos = np.column_stack([np.asarray(pts), np.full(n_particles, zc)])
    rng = np.random.default_rng(seed + 1)
    v = rng.normal(0.0, velocity_sigma, (n_particles, 3)) if velocity_sigma > 0 else np.zeros((n_particles, 3))
    om = rng.normal(0.0, omega_sigma, (n_particles, 3)) if omega_sigma > 0 else np.zeros((n_particles, 3))
    if kind == cases.CYLINDER:
        v[:, 2] = 0.0
        om[:, :2] = 0.0
    return cases.ParticleTable.uniform(n_particles, kind, pos, d=d, nTheta=nTheta, nk=nk, delta=delta if delta is not None else d / 8.0,
                                       rho=rho, v=v, omega=om)
