"""`<time>/uniform/volFieldsMethod_<fieldName>`: the accumulator dictionary of uniGasVolFields (SURVEY.md §8f rank 3,
Appendix E).

The reference keeps the time-weighted sums behind its output fields in member lists and, at every write time, dumps them
into an IOdictionary (`uniGasVolFields::writeOut`, U/macroscopicProperties/derived/volumetric/uniGasVolFields/
uniGasVolFields.C:609-669); with `averagingAcrossManyRuns` a restarted run reads them back (`readIn`, :549-605) and the
averages continue.  libugf (and the oracle) hold the same sums in `acc [nCells][16]`, `accSpecies [nCells][nSpecies]`
and `bacc [nBoundaryFaces][16]` of the state array (`ugf_state_save`, layout in csrc/ugf_api.cu); this module maps them
to the reference's entry names and back:

    cells   rhoNMean 0, rhoMMean 1, linearKEMean 2 (= sum m|U|^2, cellMeasurements.C:452), momentumMean 3-5,
            rotationalEMean 6, rotationalDofMean 7, rhoNMeanXnParticle 8, rhoMMeanXnParticle 9,
            momentumMeanXnParticle 10-12, linearKEMeanXnParticle 13, rhoNMeanInt 14;
            slot 15 = 5 rhoNMean + rotationalDofMean is rebuilt on reading (no reference entry);
            per species: nParcelsXnParticle = accSpecies; one species: nParcels = rhoNMean, mccSpecies = linearKEMean
    faces   [patch][face of the patch]: rhoNBF 0, rhoMBF 1, linearKEBF 2 (= 1/2 m|U|^2, uniGasPatchBoundary.C:170),
            momentumBF 3-5, rotationalEBF 6, rotationalDofBF 7, qBF 8, fDBF 9-11, speciesRhoNIntBF 12, electronicEBF 14;
            one species: speciesRhoNBF = rhoNBF, mccSpeciesBF = 2 linearKEBF
    scalars nTimeSteps = nAvTimeSteps, timeCounter = timeAvCounter

Entries of modes libugf does not carry (vibrational, electronic levels; mfp / mcr scratch lists, which the reference zeroes
after every write, uniGasVolFields.C:1197-1201) are written as zeros of the right shape so that `readIfPresent` finds
lists of the size it expects.  ASCII only.
"""
import os
import re

import numpy as np

from . import foamfile

STATE_MAGIC = 1431783237.0
NACC = 16
NBM = 16


class VolFieldsFormatError(ValueError):
    pass


# ---- the state array of ugf_state_save --------------------------------------------------------------------------
def state_views(buf):
    """Views into a state array (no copies): scalars [6] = step, timeAvCounter, nAvTimeSteps, sampleCounter, decTimeSteps,
    decTimeAv; acc [nC,16]; accSpecies [nC,nS]; bacc [nB,16]."""
    buf = np.asarray(buf)
    if buf.dtype != np.float64 or buf.ndim != 1 or len(buf) < 14 or buf[0] != STATE_MAGIC or buf[1] != 1.0:
        raise VolFieldsFormatError("not a ugf state array (magic / version)")
    nC, nS, nB = int(buf[2]), int(buf[3]), int(buf[4])
    o = 14 + nC * (1 + 1 + 1 + 3 + 6)
    need = o + nC * (NACC + nS) + nB * NBM
    if len(buf) < need:
        raise VolFieldsFormatError("state array shorter than its header says")
    acc = buf[o:o + nC * NACC].reshape(nC, NACC)
    o += nC * NACC
    accS = buf[o:o + nC * nS].reshape(nC, nS)
    o += nC * nS
    bacc = buf[o:o + nB * NBM].reshape(nB, NBM)
    return {"nCells": nC, "nSpecies": nS, "nBoundaryFaces": nB, "scalars": buf[8:14], "acc": acc, "accSpecies": accS, "bacc": bacc}


_CELL_SCALARS = (("rhoNMean", 0), ("rhoNMeanXnParticle", 8), ("rhoNMeanInt", 14), ("rhoMMean", 1), ("rhoMMeanXnParticle", 9),
                 ("linearKEMean", 2), ("linearKEMeanXnParticle", 13), ("rotationalEMean", 6), ("rotationalDofMean", 7))
_CELL_VECTORS = (("momentumMean", 3), ("momentumMeanXnParticle", 10))
_FACE_SCALARS = (("rhoNBF", 0), ("rhoMBF", 1), ("linearKEBF", 2), ("rotationalEBF", 6), ("rotationalDofBF", 7), ("qBF", 8),
                 ("speciesRhoNIntBF", 12))
_FACE_VECTORS = (("momentumBF", 3), ("fDBF", 9))


# ---- writer -------------------------------------------------------------------------------------------------------
def _scalar_list(a):
    a = np.asarray(a, float).ravel()
    return f"{len(a)}(" + " ".join(repr(float(v)) for v in a) + ")" if len(a) else "0()"


def _vector_list(a):
    a = np.asarray(a, float).reshape(-1, 3)
    return f"{len(a)}(" + " ".join("(" + " ".join(repr(float(c)) for c in v) + ")" for v in a) + ")" if len(a) else "0()"


def _list_of(items):
    return f"{len(items)}(" + " ".join(items) + ")" if items else "0()"


def volfields_entries(mesh, views):
    """state views -> ordered {entry name: OpenFOAM text of the value} in the order of writeOut."""
    nC, nS = views["nCells"], views["nSpecies"]
    acc, accS, bacc = views["acc"], views["accSpecies"], views["bacc"]
    if nC != mesh.n_cells or views["nBoundaryFaces"] != mesh.n_boundary_faces:
        raise VolFieldsFormatError("state array belongs to another mesh")
    nI = mesh.n_internal
    per_patch = lambda fmt, cols: _list_of([fmt(bacc[p.start - nI:p.start - nI + p.size, cols]) for p in mesh.patches])
    zero_cells = _scalar_list(np.zeros(nC))
    zero_faces = _list_of([_scalar_list(np.zeros(p.size)) for p in mesh.patches])
    per_species_cells = lambda one: _list_of([one if nS == 1 else zero_cells for _ in range(nS)])
    e = {"nTimeSteps": str(int(views["scalars"][2])), "timeCounter": repr(float(views["scalars"][1]))}
    cs = {k: _scalar_list(acc[:, s]) for k, s in _CELL_SCALARS}
    cv = {k: _vector_list(acc[:, s:s + 3]) for k, s in _CELL_VECTORS}
    fs = {k: per_patch(_scalar_list, s) for k, s in _FACE_SCALARS}
    fv = {k: per_patch(_vector_list, slice(s, s + 3)) for k, s in _FACE_VECTORS}
    e["rhoNMean"] = cs["rhoNMean"]
    e["rhoNMeanXnParticle"] = cs["rhoNMeanXnParticle"]
    e["rhoNMeanInt"] = cs["rhoNMeanInt"]
    e["molsElec"] = zero_cells
    e["rhoMMean"] = cs["rhoMMean"]
    e["rhoMMeanXnParticle"] = cs["rhoMMeanXnParticle"]
    e["linearKEMean"] = cs["linearKEMean"]
    e["linearKEMeanXnParticle"] = cs["linearKEMeanXnParticle"]
    e["rotationalEMean"] = cs["rotationalEMean"]
    e["rotationalDofMean"] = cs["rotationalDofMean"]
    e["momentumMean"] = cv["momentumMean"]
    e["momentumMeanXnParticle"] = cv["momentumMeanXnParticle"]
    e["vibrationalETotal"] = _list_of(["0()"] * nS)  # no vibrational modes
    e["electronicETotal"] = _list_of([zero_cells] * nS)
    e["nParcels"] = per_species_cells(cs["rhoNMean"])
    e["nParcelsXnParticle"] = _list_of([_scalar_list(accS[:, s]) for s in range(nS)])
    e["mccSpecies"] = per_species_cells(cs["linearKEMean"])
    e["nGroundElectronicLevel"] = _list_of([zero_cells] * nS)
    e["nFirstElectronicLevel"] = _list_of([zero_cells] * nS)
    e["mfp"] = _list_of([zero_cells] * nS)
    e["mcr"] = _list_of([zero_cells] * nS)
    e["rhoNBF"] = fs["rhoNBF"]
    e["rhoMBF"] = fs["rhoMBF"]
    e["linearKEBF"] = fs["linearKEBF"]
    e["rotationalEBF"] = fs["rotationalEBF"]
    e["rotationalDofBF"] = fs["rotationalDofBF"]
    e["qBF"] = fs["qBF"]
    e["totalvDofBF"] = zero_faces
    e["speciesRhoNIntBF"] = fs["speciesRhoNIntBF"]
    e["speciesRhoNElecBF"] = zero_faces
    e["momentumBF"] = fv["momentumBF"]
    e["fDBF"] = fv["fDBF"]
    e["vibrationalEBF"] = _list_of([zero_faces] * nS)
    e["electronicEBF"] = _list_of([per_patch(_scalar_list, 14) if nS == 1 else zero_faces for _ in range(nS)])
    e["speciesRhoNBF"] = _list_of([fs["rhoNBF"] if nS == 1 else zero_faces for _ in range(nS)])
    e["mccSpeciesBF"] = _list_of([per_patch(lambda a: _scalar_list(2.0 * np.asarray(a)), 2) if nS == 1 else zero_faces for _ in range(nS)])
    return e


def write_volfields_method(case_dir, time_name, field_name, mesh, state):
    """Writes <case>/<time>/uniform/volFieldsMethod_<field_name> from a state array; returns the path."""
    e = volfields_entries(mesh, state_views(state))
    d = os.path.join(case_dir, time_name, "uniform")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "volFieldsMethod_" + field_name)
    with open(path, "w") as f:
        f.write(foamfile._header("dictionary", f"{time_name}/uniform", os.path.basename(path)))
        for k, v in e.items():
            f.write(f"{k:<23} {v};\n\n")
        f.write("\n// ************************************************************************* //\n")
    return path


# ---- reader: `key value;` entries whose values are (nested) counted lists ---------------------------------------------
_WS = re.compile(r"\s*")
_INT = re.compile(r"\d+")
_WORD = re.compile(r"[^\s;(){}]+")
_VEC_END = re.compile(r"\)\s*\)")


def _value(s, i):
    """Parses one value at s[i:]: a number, `(x y z)`, `N(...)`, `N{v}` or `(...)`; -> (value, next index).  Flat lists of
    numbers -> float array [N]; lists of `(x y z)` -> [N,3]; anything nested deeper -> Python list of values."""
    i = _WS.match(s, i).end()
    n = None
    m = _INT.match(s, i)
    if m and m.end() < len(s) and s[m.end()] in "({":
        n = int(m.group())
        i = m.end()
    if s[i] == "{":  # uniform list N{v}
        v, j = _value(s, i + 1)
        j = _WS.match(s, j).end()
        if s[j] != "}" or n is None:
            raise VolFieldsFormatError("malformed N{value} list")
        return (np.full(n, v) if np.ndim(v) == 0 else np.tile(np.asarray(v, float), (n, 1))), j + 1
    if s[i] != "(":
        m = _WORD.match(s, i)
        if not m:
            raise VolFieldsFormatError(f"value expected at offset {i}")
        try:
            return float(m.group()), m.end()
        except ValueError:
            return m.group(), m.end()
    close, inner = s.find(")", i + 1), s.find("(", i + 1)
    if close < 0:
        raise VolFieldsFormatError("missing )")
    if inner < 0 or inner > close:  # flat
        a = np.array(s[i + 1:close].split(), float)
        if n is not None and len(a) != n:
            raise VolFieldsFormatError(f"list announces {n} entries, holds {len(a)}")
        return a, close + 1
    first = _WS.match(s, i + 1).end()
    if s[first] == "(":  # elements without a count: vectors
        m = _VEC_END.search(s, first)
        if not m:
            raise VolFieldsFormatError("missing ) after a vector list")
        a = np.array(s[first:m.end() - 1].replace("(", " ").replace(")", " ").split(), float)
        if len(a) % 3:
            raise VolFieldsFormatError("vector list with a length that is no multiple of 3")
        a = a.reshape(-1, 3)
        if n is not None and len(a) != n:
            raise VolFieldsFormatError(f"list announces {n} entries, holds {len(a)}")
        return a, m.end()
    items, j = [], i + 1
    while True:
        j = _WS.match(s, j).end()
        if j >= len(s):
            raise VolFieldsFormatError("missing )")
        if s[j] == ")":
            break
        v, j = _value(s, j)
        items.append(v)
    if n is not None and len(items) != n:
        raise VolFieldsFormatError(f"list announces {n} entries, holds {len(items)}")
    return items, j + 1


def parse_volfields_method(text):
    hdr, text = foamfile._split_header(foamfile._strip(text))  # ascii only
    out, i = {}, 0
    while True:
        i = _WS.match(text, i).end()
        if i >= len(text):
            return out
        m = _WORD.match(text, i)
        if not m:
            raise VolFieldsFormatError(f"keyword expected at offset {i}")
        key, i = m.group(), m.end()
        j = _WS.match(text, i).end()
        if text[j] == "{":  # FoamFile header
            k = text.find("}", j)
            if k < 0:
                raise VolFieldsFormatError("missing }")
            i = k + 1
            continue
        v, i = _value(text, i)
        i = _WS.match(text, i).end()
        if i >= len(text) or text[i] != ";":
            raise VolFieldsFormatError(f"missing ; after {key}")
        i += 1
        out[key] = v


def read_volfields_method(path):
    with open(path) as f:
        return parse_volfields_method(f.read())


def _per_patch(mesh, v, name, width):
    """[patch][face] lists -> [nBoundaryFaces(, 3)]"""
    if len(v) != len(mesh.patches):
        raise VolFieldsFormatError(f"{name}: {len(v)} patches, the mesh has {len(mesh.patches)}")
    out = np.zeros((mesh.n_boundary_faces, width))
    nI = mesh.n_internal
    for p, a in zip(mesh.patches, v):
        a = np.asarray(a, float).reshape(-1, width)
        if len(a) != p.size:
            raise VolFieldsFormatError(f"{name}: patch {p.name} has {p.size} faces, the list {len(a)}")
        out[p.start - nI:p.start - nI + p.size] = a
    return out


def apply_volfields_method(d, mesh, state):
    """Puts the sums of a volFieldsMethod dictionary (read_volfields_method) into a copy of a state array of the same
    set-up; entries that are absent leave the state as it is (readIfPresent)."""
    state = np.array(state, dtype=np.float64)
    v = state_views(state)
    nC, nS = v["nCells"], v["nSpecies"]
    acc, accS, bacc = v["acc"], v["accSpecies"], v["bacc"]
    if "nTimeSteps" in d:
        v["scalars"][2] = float(d["nTimeSteps"])
    if "timeCounter" in d:
        v["scalars"][1] = float(d["timeCounter"])
    for k, s in _CELL_SCALARS:
        if k in d:
            a = np.asarray(d[k], float)
            if a.shape != (nC,):
                raise VolFieldsFormatError(f"{k}: {a.shape} values for {nC} cells")
            acc[:, s] = a
    for k, s in _CELL_VECTORS:
        if k in d:
            a = np.asarray(d[k], float)
            if a.shape != (nC, 3):
                raise VolFieldsFormatError(f"{k}: {a.shape} values for {nC} cells")
            acc[:, s:s + 3] = a
    if "rhoNMean" in d or "rotationalDofMean" in d:
        acc[:, 15] = 5.0 * acc[:, 0] + acc[:, 7]
    if "nParcelsXnParticle" in d:
        lst = d["nParcelsXnParticle"]
        if len(lst) != nS:
            raise VolFieldsFormatError(f"nParcelsXnParticle: {len(lst)} species, the cloud has {nS}")
        for s in range(nS):
            accS[:, s] = np.asarray(lst[s], float)
    for k, s in _FACE_SCALARS:
        if k in d:
            bacc[:, s] = _per_patch(mesh, d[k], k, 1)[:, 0]
    for k, s in _FACE_VECTORS:
        if k in d:
            bacc[:, s:s + 3] = _per_patch(mesh, d[k], k, 3)
    if "electronicEBF" in d and nS == 1:  # mixtures: the state holds the sum over species only, the lists are written as zeros
        tot = np.zeros(mesh.n_boundary_faces)
        for sp in d["electronicEBF"]:
            tot += _per_patch(mesh, sp, "electronicEBF", 1)[:, 0]
        bacc[:, 14] = tot
    return state
