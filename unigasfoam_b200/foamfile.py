"""OpenFOAM on-disk formats of the particle loop's state - host side, ASCII.

What uniGasFoam leaves in a time directory and reads back on restart, written / read here so that a cloud driven
through libugf can be checkpointed, resumed and exchanged with the reference solver:

  <time>/lagrangian/<cloud>/positions                      Cloud<uniGasParcel>: "(x y z) celli" per parcel
  <time>/lagrangian/<cloud>/{U,cellWeight,radialWeight,ERot,ELevel,typeId,newParcel,vibLevel}
                                                          IOField<vector|scalar|label|labelField>
                                                          (U/parcels/uniGasParcelIO.C:89-181)
  <time>/<cloud>{SigmaTcRMax,CellWeightFactor,SubCellLevels,CollisionModelId}
                                                          volScalarField / volVectorField, zeroGradient patches
                                                          (U/clouds/uniGasCloud.C:433-488)
  <time>/uniform/time                                      deltaT, index (U/clouds/uniGasCloud.C:581-593)

ASCII only (`format ascii`; the reference's controlDict default); numbers are written with 17 significant digits so
that a write / read cycle returns the same bits.  Parsing is a small tokenizer for the OpenFOAM dictionary syntax plus
numpy for the long lists; it reads the reference's own tutorial fields (tests/golden/openfoam).
"""
import os
import re

import numpy as np

BANNER = ("/*--------------------------------*- C++ -*----------------------------------*\\\n"
          "| =========                 |                                                 |\n"
          "| \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox           |\n"
          "|  \\\\    /   O peration     | Version:  2212                                  |\n"
          "|   \\\\  /    A nd           | Website:  www.openfoam.com                      |\n"
          "|    \\\\/     M anipulation  |                                                 |\n"
          "\\*---------------------------------------------------------------------------*/\n")
RULE = "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n"


class FoamFormatError(ValueError):
    pass


# ---- writing ---------------------------------------------------------------------------------------------
def _header(cls, location, obj):
    return (BANNER + "FoamFile\n{\n    version     2.0;\n    format      ascii;\n"
            '    arch        "LSB;label=32;scalar=64";\n'
            f"    class       {cls};\n    location    \"{location}\";\n    object      {obj};\n}}\n" + RULE + "\n")


def _fmt_scalars(a):
    return "\n".join(repr(float(v)) for v in a)


def _fmt_vectors(a):
    return "\n".join("(" + " ".join(repr(float(c)) for c in v) + ")" for v in a)


def _fmt_labels(a):
    return "\n".join(str(int(v)) for v in a)


def write_io_field(path, cls, location, values, kind):
    """IOField<T>: `N ( ... )`.  kind: scalar | vector | label | labelField (list of label lists)."""
    n = len(values)
    if kind == "scalar":
        body = _fmt_scalars(values)
    elif kind == "vector":
        body = _fmt_vectors(values)
    elif kind == "label":
        body = _fmt_labels(values)
    elif kind == "labelField":
        body = "\n".join(f"{len(v)}(" + " ".join(str(int(x)) for x in v) + ")" for v in values)
    else:
        raise ValueError(kind)
    with open(path, "w") as f:
        f.write(_header(cls, location, os.path.basename(path)))
        f.write(f"{n}\n(\n{body}\n)\n" if n else "0()\n")
        f.write("\n\n// ************************************************************************* //\n")


def write_lagrangian(case_dir, time_name, parcels, cloud="uniGas"):
    """uniGasParcel::writeFields (U/parcels/uniGasParcelIO.C:141-181) + particle::writeFields positions.
    parcels: dict with position [n,3], U [n,3], cell [n] and optionally typeId, ERot, cellWeight, radialWeight,
    ELevel, newParcel, vibLevel (defaults: 0 / 1.0 / empty)."""
    n = len(parcels["cell"])
    loc = f"{time_name}/lagrangian/{cloud}"
    d = os.path.join(case_dir, time_name, "lagrangian", cloud)
    os.makedirs(d, exist_ok=True)
    pos, cell = np.asarray(parcels["position"], float), np.asarray(parcels["cell"])
    with open(os.path.join(d, "positions"), "w") as f:
        f.write(_header(f"Cloud<{cloud}Parcel>", loc, "positions"))
        if n:
            f.write(f"{n}\n(\n")
            f.write("\n".join("(" + " ".join(repr(float(c)) for c in p) + f") {int(c_)}" for p, c_ in zip(pos, cell)))
            f.write("\n)\n")
        else:
            f.write("0()\n")
        f.write("\n\n// ************************************************************************* //\n")
    get = lambda k, default: np.asarray(parcels[k]) if parcels.get(k) is not None else np.full(n, default)
    write_io_field(os.path.join(d, "U"), "vectorField", loc, np.asarray(parcels["U"], float), "vector")
    write_io_field(os.path.join(d, "cellWeight"), "scalarField", loc, get("cellWeight", 1.0), "scalar")
    write_io_field(os.path.join(d, "radialWeight"), "scalarField", loc, get("radialWeight", 1.0), "scalar")
    write_io_field(os.path.join(d, "ERot"), "scalarField", loc, get("ERot", 0.0), "scalar")
    write_io_field(os.path.join(d, "ELevel"), "labelField", loc, get("ELevel", 0), "label")
    write_io_field(os.path.join(d, "typeId"), "labelField", loc, get("typeId", 0), "label")
    write_io_field(os.path.join(d, "newParcel"), "labelField", loc, get("newParcel", 0), "label")
    vib = parcels.get("vibLevel")
    write_io_field(os.path.join(d, "vibLevel"), "labelFieldField", loc, vib if vib is not None else [[]] * n, "labelField")
    return d


def write_vol_field(path, location, dimensions, internal, patches, vector=False):
    """volScalarField / volVectorField with zeroGradient (or given) patch types; patches: list of names or
    dict name -> type word, or name -> {"type": word, "value": scalar / vector / array per face} (what a `calculated`
    patch of an output field carries)."""
    internal = np.asarray(internal, float)
    cls = "volVectorField" if vector else "volScalarField"
    typ = "vector" if vector else "scalar"
    if isinstance(patches, dict):
        ptypes = patches
    else:
        ptypes = {p: "zeroGradient" for p in patches}
    with open(path, "w") as f:
        f.write(_header(cls, location, os.path.basename(path)))
        f.write("dimensions      [" + " ".join(str(int(x)) for x in dimensions) + "];\n\n")
        flat = internal.reshape(len(internal), -1) if internal.ndim > 1 or vector else internal
        uniform = len(internal) > 0 and (flat == flat[0]).all()
        if uniform:
            v = "(" + " ".join(repr(float(c)) for c in flat[0]) + ")" if vector else repr(float(flat[0]))
            f.write(f"internalField   uniform {v};\n\n")
        else:
            f.write(f"internalField   nonuniform List<{typ}> \n{len(internal)}\n(\n")
            f.write(_fmt_vectors(internal) if vector else _fmt_scalars(internal))
            f.write("\n)\n;\n\n")
        f.write("boundaryField\n{\n")
        for name, t in ptypes.items():
            if isinstance(t, dict):
                f.write(f"    {name}\n    {{\n        type            {t['type']};\n")
                if "value" in t:
                    f.write("        value           " + _patch_value(t["value"], vector) + ";\n")
                f.write("    }\n")
                continue
            f.write(f"    {name}\n    {{\n        type            {t};\n    }}\n")
        f.write("}\n\n\n// ************************************************************************* //\n")


def _patch_value(v, vector):
    v = np.asarray(v, float)
    one = "(" + " ".join(repr(float(c)) for c in v) + ")" if vector and v.ndim == 1 else (repr(float(v)) if v.ndim == 0 else None)
    if one is not None:
        return "uniform " + one
    if len(v) == 0:
        return f"nonuniform List<{'vector' if vector else 'scalar'}> 0()"
    body = " ".join("(" + " ".join(repr(float(c)) for c in r) + ")" for r in v) if vector else " ".join(repr(float(x)) for x in v)
    return f"nonuniform List<{'vector' if vector else 'scalar'}> {len(v)}({body})"


def boundary_values(field, patch, n_faces):
    """The `value` entry of one patch of a field read by read_vol_field -> array [n_faces] / [n_faces,3] (None if the
    patch has no value entry)."""
    e = field["boundary"][patch]
    if "value" not in e:
        return None
    vector = field["class"] == "volVectorField"
    txt = e["value"]
    if txt.startswith("uniform"):
        v = _numbers(txt[len("uniform"):])
        return np.tile(v, (n_faces, 1)) if vector else np.full(n_faces, float(v[0]))
    m = re.match(r"nonuniform\s+List<\w+>\s*(\d+)\s*\((.*)\)\s*$", txt, re.S)
    if not m:
        raise FoamFormatError(f"{patch}: uniform or nonuniform List<T> value expected")
    a = _numbers(m.group(2))
    a = a.reshape(-1, 3) if vector else a
    if len(a) != int(m.group(1)) or len(a) != n_faces:
        raise FoamFormatError(f"{patch}: {len(a)} values for {n_faces} faces")
    return a


# ---- reading ---------------------------------------------------------------------------------------------
_COMMENT = re.compile(r"/\*.*?\*/|//[^\n]*", re.S)


def _strip(text):
    return _COMMENT.sub(" ", text)


def _split_header(text):
    m = re.search(r"FoamFile\s*\{(.*?)\}", text, re.S)
    if not m:
        raise FoamFormatError("no FoamFile header")
    hdr = dict(re.findall(r"(\w+)\s+([^;]+);", m.group(1)))
    hdr = {k: v.strip().strip('"') for k, v in hdr.items()}
    if hdr.get("format", "ascii") != "ascii":
        raise FoamFormatError("only `format ascii` is supported")
    return hdr, text[m.end():]


def _numbers(s, dtype=float):
    s = s.replace("(", " ").replace(")", " ")
    a = np.array(s.split(), dtype=dtype) if s.strip() else np.empty(0, dtype)
    return a


def _list_body(rest):
    """`N ( ... )` or `N{v}` or `0()`: returns (n, inner text or None, uniform value text or None, text after)."""
    m = re.match(r"\s*(\d+)\s*", rest)
    if not m:
        raise FoamFormatError("list size expected")
    n = int(m.group(1))
    rest = rest[m.end():]
    if rest.startswith("{"):
        j = rest.index("}")
        return n, None, rest[1:j], rest[j + 1:]
    if not rest.startswith("("):
        raise FoamFormatError("( expected")
    depth, j = 0, 0
    for j, ch in enumerate(rest):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                break
    return n, rest[1:j], None, rest[j + 1:]


def read_io_field(path, kind):
    hdr, rest = _split_header(_strip(open(path).read()))
    n, inner, uni, _ = _list_body(rest)
    if kind == "labelField":
        if inner is None:
            return [list(_numbers(uni, int))[1:] if uni.strip() else [] for _ in range(n)]
        out = []
        for m in re.finditer(r"(\d+)\s*\(([^)]*)\)", inner):
            out.append([int(x) for x in m.group(2).split()])
        if len(out) != n:
            raise FoamFormatError(f"{path}: {len(out)} entries, header says {n}")
        return out
    dtype = int if kind == "label" else float
    width = 3 if kind == "vector" else 1
    if inner is None:
        a = np.tile(_numbers(uni, dtype), n)
    else:
        a = _numbers(inner, dtype)
    if a.size != n * width:
        raise FoamFormatError(f"{path}: {a.size} numbers for {n} x {width}")
    return a.reshape(n, 3) if kind == "vector" else a


def read_lagrangian(case_dir, time_name, cloud="uniGas"):
    """Inverse of write_lagrangian (uniGasParcel::readFields, U/parcels/uniGasParcelIO.C:89-138)."""
    d = os.path.join(case_dir, time_name, "lagrangian", cloud)
    hdr, rest = _split_header(_strip(open(os.path.join(d, "positions")).read()))
    if not hdr.get("class", "").startswith("Cloud<"):
        raise FoamFormatError("positions: class Cloud<...> expected")
    n, inner, _, _ = _list_body(rest)
    a = _numbers(inner or "")
    if a.size != 4 * n:
        raise FoamFormatError("positions: `(x y z) celli` per parcel expected (barycentric `coordinates` files are not read)")
    a = a.reshape(n, 4)
    out = {"position": np.ascontiguousarray(a[:, :3]), "cell": a[:, 3].astype(np.int32)}
    if not np.array_equal(out["cell"], a[:, 3]):
        raise FoamFormatError("positions: non-integer cell label")
    for name, kind in (("U", "vector"), ("cellWeight", "scalar"), ("radialWeight", "scalar"), ("ERot", "scalar"), ("ELevel", "label"),
                       ("typeId", "label"), ("newParcel", "label"), ("vibLevel", "labelField")):
        v = read_io_field(os.path.join(d, name), kind)
        if len(v) != n:
            raise FoamFormatError(f"{name}: {len(v)} entries for {n} parcels")  # Cloud::checkFieldIOobject
        out[name] = v
    out["typeId"] = out["typeId"].astype(np.int32)
    return out


def read_vol_field(path):
    """-> dict(class, dimensions, internal (float or array [n] / [n,3], or a uniform value), uniform: bool, boundary:
    {patch: {key: word}})."""
    hdr, rest = _split_header(_strip(open(path).read()))
    vector = hdr.get("class") == "volVectorField"
    m = re.search(r"dimensions\s*\[([^\]]*)\]\s*;", rest)
    if not m:
        raise FoamFormatError("dimensions missing")
    dims = [int(float(x)) for x in m.group(1).split()]
    m = re.search(r"internalField\s+", rest)
    if not m:
        raise FoamFormatError("internalField missing")
    tail = rest[m.end():]
    if tail.startswith("uniform"):
        j = tail.index(";")
        val = _numbers(tail[len("uniform"):j])
        internal, uniform = (val if vector else float(val[0])), True
        tail = tail[j + 1:]
    else:
        m2 = re.match(r"nonuniform\s+List<\w+>\s*", tail)
        if not m2:
            raise FoamFormatError("internalField: uniform or nonuniform List<T> expected")
        n, inner, uni, tail = _list_body(tail[m2.end():])
        a = _numbers(inner) if inner is not None else np.tile(_numbers(uni), n)
        internal, uniform = (a.reshape(n, 3) if vector else a), False
        if len(internal) != n:
            raise FoamFormatError("internalField size mismatch")
    boundary = {}
    m = re.search(r"boundaryField\s*\{", tail)
    if m:
        body = tail[m.end():]
        for pm in re.finditer(r"([\w.]+)\s*\{([^{}]*)\}", body):
            boundary[pm.group(1)] = {k: v.strip() for k, v in re.findall(r"(\w+)\s+([^;]+);", pm.group(2))}
    return {"class": hdr.get("class"), "object": hdr.get("object"), "dimensions": dims, "internal": internal, "uniform": uniform,
            "boundary": boundary}


def expand_internal(field, n_cells):
    v = field["internal"]
    if field["uniform"]:
        return np.tile(np.asarray(v, float), (n_cells, 1)) if np.ndim(v) else np.full(n_cells, float(v))
    if len(v) != n_cells:
        raise FoamFormatError(f"{field.get('object')}: {len(v)} values for {n_cells} cells")
    return v


# ---- the cloud's time directory ------------------------------------------------------------------------------
def write_cloud_time(case_dir, time_name, mesh, parcels, sigmaTcRMax, cellWeightFactor=None, subCellLevels=None, collisionModelId=None,
                     deltaT=None, index=0, cloud="uniGas"):
    """What uniGasFoam writes for the cloud at a write time: lagrangian/<cloud>/*, the four cell-state fields and
    uniform/time."""
    t = os.path.join(case_dir, time_name)
    os.makedirs(os.path.join(t, "uniform"), exist_ok=True)
    write_lagrangian(case_dir, time_name, parcels, cloud)
    nC = mesh.n_cells
    patches = [p.name for p in mesh.patches]
    full = lambda v, d: np.full(nC, d, float) if v is None else np.broadcast_to(np.asarray(v, float), (nC,))
    write_vol_field(os.path.join(t, cloud + "SigmaTcRMax"), time_name, [0, 3, -1, 0, 0, 0, 0], full(sigmaTcRMax, 0.0), patches)
    write_vol_field(os.path.join(t, cloud + "CellWeightFactor"), time_name, [0] * 7, full(cellWeightFactor, 1.0), patches)
    write_vol_field(os.path.join(t, cloud + "CollisionModelId"), time_name, [0] * 7, full(collisionModelId, 0.0), patches)
    lv = np.ones((nC, 3)) if subCellLevels is None else np.broadcast_to(np.asarray(subCellLevels, float), (nC, 3))
    write_vol_field(os.path.join(t, cloud + "SubCellLevels"), time_name, [0] * 7, lv, patches, vector=True)
    with open(os.path.join(t, "uniform", "time"), "w") as f:
        f.write(_header("dictionary", f"{time_name}/uniform", "time"))
        f.write(f"value           {time_name};\n\nname            \"{time_name}\";\n\nindex           {int(index)};\n\n")
        if deltaT is not None:
            f.write(f"deltaT          {float(deltaT)!r};\n\ndeltaT0         {float(deltaT)!r};\n\n")
        f.write("\n// ************************************************************************* //\n")
    return t


def read_cloud_time(case_dir, time_name, n_cells, cloud="uniGas"):
    t = os.path.join(case_dir, time_name)
    out = {"parcels": read_lagrangian(case_dir, time_name, cloud)}
    for key, name in (("sigmaTcRMax", "SigmaTcRMax"), ("cellWeightFactor", "CellWeightFactor"), ("collisionModelId", "CollisionModelId"),
                      ("subCellLevels", "SubCellLevels")):
        p = os.path.join(t, cloud + name)
        out[key] = expand_internal(read_vol_field(p), n_cells) if os.path.exists(p) else None
    tp = os.path.join(t, "uniform", "time")
    if os.path.exists(tp):
        _, rest = _split_header(_strip(open(tp).read()))
        kv = dict(re.findall(r"(\w+)\s+([^;]+);", rest))
        out["index"] = int(kv.get("index", 0))
        out["deltaT"] = float(kv["deltaT"]) if "deltaT" in kv else None
    return out


# ---- constant/polyMesh -----------------------------------------------------------------------------------------
def write_polymesh(case_dir, mesh, region="constant/polyMesh"):
    """points, faces, owner, neighbour, boundary of a PolyMesh in OpenFOAM's ASCII layout (what blockMesh writes)."""
    d = os.path.join(case_dir, region)
    os.makedirs(d, exist_ok=True)
    nP, nF, nI, nC = len(mesh.points), mesh.n_faces, mesh.n_internal, mesh.n_cells
    note = f'    note        "nPoints:{nP}  nCells:{nC}  nFaces:{nF}  nInternalFaces:{nI}";\n'

    def hdr(cls, obj, with_note=False):
        h = _header(cls, region, obj)
        return h.replace(f"    class       {cls};\n", f"    class       {cls};\n" + (note if with_note else ""))

    with open(os.path.join(d, "points"), "w") as f:
        f.write(hdr("vectorField", "points"))
        f.write(f"{nP}\n(\n" + _fmt_vectors(mesh.points) + "\n)\n")
    off, fp = mesh.face_point_offsets, mesh.face_points
    with open(os.path.join(d, "faces"), "w") as f:
        f.write(hdr("faceList", "faces"))
        f.write(f"{nF}\n(\n")
        f.write("\n".join(f"{off[i + 1] - off[i]}(" + " ".join(str(int(v)) for v in fp[off[i]:off[i + 1]]) + ")" for i in range(nF)))
        f.write("\n)\n")
    for name, arr in (("owner", mesh.owner), ("neighbour", mesh.neighbour)):
        with open(os.path.join(d, name), "w") as f:
            f.write(hdr("labelList", name, with_note=True))
            f.write(f"{len(arr)}\n(\n" + _fmt_labels(arr) + "\n)\n")
    with open(os.path.join(d, "boundary"), "w") as f:
        f.write(hdr("polyBoundaryMesh", "boundary"))
        f.write(f"{len(mesh.patches)}\n(\n")
        for p in mesh.patches:
            f.write(f"    {p.name}\n    {{\n        type            {p.kind};\n")
            if p.kind == "cyclic":
                f.write(f"        neighbourPatch  {mesh.patches[p.partner].name};\n")
            if p.kind == "processor":
                f.write(f"        myProcNo        0;\n        neighbProcNo    {p.partner};\n")
            f.write(f"        nFaces          {p.size};\n        startFace       {p.start};\n    }}\n")
        f.write(")\n")
    return d


def read_polymesh(case_dir, region="constant/polyMesh"):
    """constant/polyMesh -> PolyMesh with its derived geometry (primitiveMesh: face areas / centres, cell centres /
    volumes, cell -> faces).  Patch types as in the boundary file; cyclic partners from neighbourPatch; directions whose
    faces are all on `empty` patches are not solved (polyMesh::solutionD)."""
    from . import foamdict, mesh as _mesh
    d = os.path.join(case_dir, region)

    def body(name):
        hdr, rest = _split_header(_strip(open(os.path.join(d, name)).read()))
        return _list_body(rest)

    nP, inner, _, _ = body("points")
    pts = _numbers(inner).reshape(nP, 3)
    nF, inner, _, _ = body("faces")
    sizes, flat = [], []
    for m in re.finditer(r"(\d+)\s*\(([^()]*)\)", inner):
        ids = m.group(2).split()
        if len(ids) != int(m.group(1)):
            raise FoamFormatError("faces: entry size mismatch")
        sizes.append(len(ids)); flat.extend(ids)
    if len(sizes) != nF:
        raise FoamFormatError(f"faces: {len(sizes)} entries, header says {nF}")
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    fpt = np.array(flat, dtype=np.int32)
    nO, inner, uni, _ = body("owner")
    owner = _numbers(inner, int).astype(np.int32) if inner is not None else np.full(nO, int(uni), np.int32)
    nN, inner, uni, _ = body("neighbour")
    neigh = _numbers(inner or "", int).astype(np.int32)
    if len(owner) != nF or len(neigh) != nN:
        raise FoamFormatError("owner / neighbour size mismatch")
    text = _strip(open(os.path.join(d, "boundary")).read())
    _, rest = _split_header(text)
    m = re.match(r"\s*(\d+)\s*\(", rest)
    if not m:
        raise FoamFormatError("boundary: list expected")
    ent = foamdict.parse(rest[m.end():rest.rindex(")")])  # the entries of the list are `name { ... }`: a dictionary body
    if len(ent) != int(m.group(1)):
        raise FoamFormatError("boundary: patch count mismatch")
    names = list(ent)
    patches = []
    for name in names:
        e = ent[name]
        kind = e["type"]
        if kind in ("symmetryPlane", "symmetry", "wedge"):
            kind = "symmetryPlane" if kind != "symmetry" else "symmetry"
        partner = -1
        if kind == "cyclic":
            partner = names.index(e["neighbourPatch"])
        if kind == "processor":
            partner = int(e["neighbProcNo"])
        patches.append(_mesh.Patch(name, kind, int(e["startFace"]), int(e["nFaces"]), partner))
    pm = _mesh.PolyMesh(pts, off, fpt, owner, neigh, patches)
    pm.compute_geometry()
    sol = [1, 1, 1]
    for p in patches:  # polyMesh::calcDirections: the normals of the empty patches mark the unsolved direction
        if p.kind == "empty" and p.size:
            S = np.abs(pm.face_areas[p.start:p.start + p.size]).sum(0)
            sol[int(np.argmax(S))] = 0
    pm.solution_d = tuple(sol)
    for p in patches:  # cyclic separation: offset between the two patches' face centres
        if p.kind == "cyclic":
            q = patches[p.partner]
            sep = pm.face_centres[q.start:q.start + q.size].mean(0) - pm.face_centres[p.start:p.start + p.size].mean(0)
            p.separation = tuple(float(v) for v in sep)
    return pm
