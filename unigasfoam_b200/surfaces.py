"""Surface field models of fieldPropertiesDict on top of the device tallies (host side, evaluated at write time).

uniGasMassFluxSurface  (U/macroscopicProperties/derived/massFlux/uniGasMassFluxSurface/uniGasMassFluxSurface.C:213-330):
    flux of molecules, mass, momentum and energy through a face zone along `fluxDirection`, from uniGasFaceTracker's
    per-face tallies (ugf_set_face_tracker / ugf_download_face_tracker).
uniGasForceSurface     (.../force/uniGasForceSurface/uniGasForceSurface.C:130-200):
    force on a wall patch, sum over its faces of fD |Sf| from the boundary measurements (ugf_download_fields).
Both average over the time since the last reset; with sampleInterval 1 that is what the reference accumulates.
"""
import numpy as np


class UniGasMassFluxSurface:
    def __init__(self, cloud, properties):
        """properties: field, faceZone (list of face labels of the zone - the reference looks the name up in mesh.faceZones()),
        fluxDirection, typeIds (names)."""
        self.cloud, self.mesh = cloud, cloud.mesh
        self.fieldName = properties["field"]
        self.faces = np.asarray(properties["faceZone"], np.int32)
        d = np.asarray(properties["fluxDirection"], float)
        self.fluxDirection = d / np.sqrt((d * d).sum())
        self.typeIds = [cloud.typeIdList.index(n) for n in properties.get("typeIds", cloud.typeIdList)]
        S = self.mesh.face_areas[self.faces]
        self.magSf = np.sqrt((S * S).sum(1))
        self.nF = S / self.magSf[:, None]
        self.zoneSurfaceArea = float(self.magSf.sum())
        self.timeAvCounter = 0.0
        self._t0 = None
        cloud.setFaceTracker(self.faces)

    def begin(self):
        self.cloud.faceTracker(reset=True)
        self._t0 = self.cloud.counters()["step"]

    def calculateField(self, resetAtOutput=False):
        """-> dict molFlux, massFlux, momentumFlux, energyFlux (per unit area and time) over the steps since begin()."""
        if self._t0 is None:
            raise RuntimeError("begin() first")
        t = self.cloud.faceTracker(reset=resetAtOutput)[:, self.typeIds, :].sum(1)  # [faces, 6]
        steps = self.cloud.counters()["step"] - self._t0
        self.timeAvCounter = steps * self.cloud.cfg.deltaT
        FN = self.cloud.cfg.nParticle
        proj = self.nF @ self.fluxDirection
        out = {
            "molFlux": float((t[:, 0] * FN * proj).sum()),
            "massFlux": float((t[:, 1] * FN * proj).sum()),
            "momentumFlux": float((t[:, 2:5] * FN @ self.fluxDirection).sum()),
            "energyFlux": float((t[:, 5] * FN * proj).sum()),
        }
        den = self.timeAvCounter * self.zoneSurfaceArea
        out = {k: (v / den if den > 0 else 0.0) for k, v in out.items()}
        if resetAtOutput:
            self._t0 = self.cloud.counters()["step"]
        return out


class UniGasForceSurface:
    def __init__(self, cloud, properties):
        self.cloud, self.mesh = cloud, cloud.mesh
        self.fieldName = properties["field"]
        self.patch = self.mesh.patch_index(properties["patch"])

    def calculateField(self, resetAtOutput=False):
        """Time-averaged force vector on the patch: sum of fD |Sf| over its faces."""
        p = self.mesh.patches[self.patch]
        f = self.cloud.fields(resetAtOutput=resetAtOutput)
        b0 = p.start - self.mesh.n_internal
        S = self.mesh.face_areas[p.start:p.start + p.size]
        A = np.sqrt((S * S).sum(1))
        return (f["fD"][b0:b0 + p.size] * A[:, None]).sum(0)
