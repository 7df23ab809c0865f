"""OpenFOAM dictionary files -> Python dicts, and a uniGasFoam case directory -> the inputs of the host mirror.

The reference is configured through `constant/uniGasProperties`, `system/{controlDict, boundariesDict,
fieldPropertiesDict, hybridDecompositionDict, uniGasInitialisationDict}` (tutorials/uniGasFoam/*).
`UniGasCloud` takes exactly those dictionaries with the reference's key names, so reading the files is all it takes to
run a reference case on libugf: `load_case(case_dir)`.

Parser: the subset of the OpenFOAM dictionary grammar those files use - `key value ... ;`, `key { ... }`, lists
`( ... )` whose items may be named sub-dictionaries (`boundary { ... }` -> the dictionary), `key{...}` without
white space, comments, the FoamFile header (dropped), switches (yes/no/on/off/true/false -> bool).  No `#include`, no
`$macro` expansion.
"""
import os
import re

_TOKEN = re.compile(r'"[^"]*"|[{}();]|[^\s{}();"]+')
_SWITCH = {"yes": True, "on": True, "true": True, "no": False, "off": False, "false": False}
_COMMENT = re.compile(r"/\*.*?\*/|//[^\n]*", re.S)


class FoamDictError(ValueError):
    pass


def _atom(tok):
    if tok.startswith('"'):
        return tok[1:-1]
    if tok in _SWITCH:
        return _SWITCH[tok]
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


class _Parser:
    def __init__(self, text):
        self.t = _TOKEN.findall(_COMMENT.sub(" ", text))
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        tok = self.peek()
        if tok is None:
            raise FoamDictError("unexpected end of file")
        self.i += 1
        return tok

    def dict_body(self, top=False):
        out = {}
        while True:
            tok = self.peek()
            if tok is None:
                if top:
                    return out
                raise FoamDictError("missing }")
            if tok == "}":
                if top:
                    raise FoamDictError("unbalanced }")
                self.i += 1
                return out
            key = self.next()
            if key in "{();":
                raise FoamDictError(f"keyword expected, got {key!r}")
            if self.peek() == "{":
                self.i += 1
                out[key] = self.dict_body()
                if self.peek() == ";":  # `numberDensities {Ar 4.2e20;};` - a stray semicolon after a sub-dictionary
                    self.i += 1
                continue
            vals = []
            while self.peek() != ";":
                if self.peek() is None:
                    raise FoamDictError(f"missing ; after {key}")
                vals.append(self.value())
            self.i += 1
            out[key] = None if not vals else (vals[0] if len(vals) == 1 else vals)

    def value(self):
        tok = self.next()
        if tok == "(":
            return self.list_body()
        if tok == "{":
            return self.dict_body()
        if tok in ");}":
            raise FoamDictError(f"value expected, got {tok!r}")
        return _atom(tok)

    def list_body(self):
        items = []
        while True:
            tok = self.peek()
            if tok is None:
                raise FoamDictError("missing )")
            if tok == ")":
                self.i += 1
                return items
            if tok == "(":
                self.i += 1
                items.append(self.list_body())
            elif tok == "{":
                self.i += 1
                items.append(self.dict_body())
            else:
                self.i += 1
                if self.peek() == "{":  # named entry of a list of dictionaries: `boundary { ... }`
                    self.i += 1
                    items.append(self.dict_body())
                else:
                    items.append(_atom(tok))


def parse(text):
    d = _Parser(text).dict_body(top=True)
    d.pop("FoamFile", None)
    return d


def read(path):
    with open(path) as f:
        return parse(f.read())


# ---- a uniGasFoam case -----------------------------------------------------------------------------------------
def _mol(m):
    """moleculeProperties entry: empty lists come as [] - keep the keys cloud._species_struct reads."""
    out = dict(m)
    for k in ("electronicEnergyList", "degeneracyList"):
        if k in out and not isinstance(out[k], list):
            out[k] = [out[k]]
    return out


def load_case(case_dir, overrides=None):
    """-> dict with uniGasProperties, boundariesDict, deltaT, fieldPropertiesDict, hybridDecompositionDict (or None),
    uniGasInitialisationDict (or None), controlDict.  `overrides` is merged into uniGasProperties (one level deep for
    sub-dictionaries), e.g. {"collisionProperties": {"macroInterpolation": False}}."""
    sysd, cst = os.path.join(case_dir, "system"), os.path.join(case_dir, "constant")
    props = read(os.path.join(cst, "uniGasProperties"))
    props["typeIdList"] = list(props["typeIdList"]) if isinstance(props["typeIdList"], list) else [props["typeIdList"]]
    props["moleculeProperties"] = {k: _mol(v) for k, v in props["moleculeProperties"].items()}
    for k, v in (overrides or {}).items():
        if isinstance(v, dict) and isinstance(props.get(k), dict):
            props[k] = dict(props[k], **v)
        else:
            props[k] = v
    out = {"uniGasProperties": props}
    ctl = read(os.path.join(sysd, "controlDict"))
    out["controlDict"] = ctl
    out["deltaT"] = float(ctl["deltaT"])
    bd = read(os.path.join(sysd, "boundariesDict"))
    for key in ("uniGasPatchBoundaries", "uniGasGeneralBoundaries", "uniGasCyclicBoundaries"):
        bd[key] = bd.get(key) or []
    for e in bd["uniGasGeneralBoundaries"]:  # typeIds (Ar) -> ["Ar"]
        pr = e.get(e["boundaryModel"] + "Properties", {})
        if "typeIds" in pr and not isinstance(pr["typeIds"], list):
            pr["typeIds"] = [pr["typeIds"]]
    out["boundariesDict"] = bd
    opt = lambda name: read(os.path.join(sysd, name)) if os.path.exists(os.path.join(sysd, name)) else None
    out["fieldPropertiesDict"] = opt("fieldPropertiesDict")
    out["hybridDecompositionDict"] = opt("hybridDecompositionDict")
    out["uniGasInitialisationDict"] = opt("uniGasInitialisationDict")
    return out


def sample_interval(fieldPropertiesDict):
    """timeProperties.sampleInterval of the first uniGasVolFields entry (default 1)."""
    for f in (fieldPropertiesDict or {}).get("uniGasFields", []):
        if f.get("fieldModel") == "uniGasVolFields":
            return int(f.get("timeProperties", {}).get("sampleInterval", 1))
    return 1


def vol_field_names(fieldPropertiesDict, averaging_only=False):
    """The `field` words of the uniGasVolFields entries (uniGasVolFields.C:54): the names behind
    uniform/volFieldsMethod_<field>; with averaging_only, just the entries that set `averagingAcrossManyRuns`."""
    out = []
    for f in (fieldPropertiesDict or {}).get("uniGasFields", []):
        if f.get("fieldModel") != "uniGasVolFields":
            continue
        pr = f.get("uniGasVolFieldsProperties", {})
        if "field" in pr and (not averaging_only or pr.get("averagingAcrossManyRuns", False)):
            out.append(str(pr["field"]))
    return out


def partition_from_dict(mesh, decomposeParDict, weights=None, axis=0):
    """system/decomposeParDict -> cell-to-rank map for `mesh.decompose`.  `numberOfSubdomains` ranks; `method simple` with
    `simpleCoeffs { n (nx ny nz); }` cuts along the direction with more than one part (one direction only here);
    `scotch` / `hierarchical` / anything else falls back to slabs along `axis`.  With `weightField <name>` the caller
    passes that field's cell values as `weights` and the slabs are cut at equal weight (the commented
    `weightField uniGasRhoNMean_Ar` of the hypersonicCylinder tutorial).  -> (cell_rank, n_ranks)"""
    from . import mesh as ugmesh
    n = int(decomposeParDict["numberOfSubdomains"])
    if decomposeParDict.get("method") == "simple":
        parts = [int(v) for v in (decomposeParDict.get("simpleCoeffs") or decomposeParDict.get("coeffs") or {}).get("n", [n, 1, 1])]
        if parts[0] * parts[1] * parts[2] != n:
            raise FoamDictError(f"simpleCoeffs n {parts} does not multiply to numberOfSubdomains {n}")
        split = [i for i, k in enumerate(parts) if k > 1]
        if len(split) > 1:
            if "weightField" in decomposeParDict:
                raise FoamDictError(f"simpleCoeffs n {parts} with a weightField: only a split along one direction is supported")
            return ugmesh.block_partition(mesh, parts), n
        if split:
            axis = split[0]
    if "weightField" in decomposeParDict:
        if weights is None:
            raise FoamDictError(f"decomposeParDict names weightField {decomposeParDict['weightField']}: pass its cell values as `weights`")
        return ugmesh.weighted_slab_partition(mesh, n, weights, axis), n
    return ugmesh.slab_partition(mesh, n, axis), n
