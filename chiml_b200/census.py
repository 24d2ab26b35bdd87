"""Cell census and algorithmic-bytes model.

Algorithmic bytes per step count every TIME-VARYING state array once read + once written, plus one
read of each field array by the other half-step's stencil; static coefficient / id data excluded:
  * every updated field-component cell:         16 B (RW) + 8 B (cross-read)         = 24 B
      -> 3-D vacuum cell = 6 x 24 = 144 B, 2-D cell = 3 x 24 = 72 B
  * every D-type cell, per E component:          + 16 B (D RW)
  * every pole of a dispersive cell, per comp:   + 24 B (P read, P_prev read, P_new written)
  * every CPML psi touched:                      + 16 B
  * every emitter, per level system of N levels: + 96 N^2 B (rho RW, 4 histories read, 1 written) + 96 B
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np

from . import plan as P


@dataclass
class Census:
    cells: int                 # grid cells of this slab (no ghosts), the unit of "cell-updates"
    comp_cells: int            # updated field-component cells (curl or CPML grid term)
    d_cells: int               # E-component cells that carry D
    pole_cells: int            # sum over dispersive component cells of their pole count
    psi_cells: int             # psi values updated per step
    emitter_bytes: int = 0

    @property
    def bytes_per_step(self) -> int:
        return 24 * self.comp_cells + 16 * self.d_cells + 24 * self.pole_cells + 16 * self.psi_cells + self.emitter_bytes

    def as_dict(self) -> Dict[str, int]:
        return {"cells": self.cells, "component_cells": self.comp_cells, "d_cells": self.d_cells, "pole_cells": self.pole_cells,
                "psi_cells": self.psi_cells, "emitter_bytes": self.emitter_bytes, "bytes_per_step": self.bytes_per_step,
                "bytes_per_cell": round(self.bytes_per_step / max(self.cells, 1), 2)}


def census(plan: P.Plan) -> Census:
    lnx, lny, lnz = plan.ln
    cells = (lnx - 2) * (lny - 2) * (lnz - 2 if lnz > 1 else 1)
    npoles = {o.obj: o.npoles for o in plan.objects}
    comp_cells = 0
    d_cells = 0
    pole_cells = 0
    for (kind, comp), runs in plan.lists.items():
        if len(runs) == 0:
            continue
        n = int(runs["n"].sum())
        if kind in (P.LIST_U, P.LIST_D):
            comp_cells += n
        if kind in (P.LIST_LORD, P.LIST_ORDIPD):
            d_cells += n
        if kind == P.LIST_LORD:
            pole_cells += int((runs["n"] * np.array([npoles.get(int(o), 0) for o in runs["obj"]])).sum())
        if kind == P.LIST_ORDIPP:
            ncomp = 3 if plan.mode == P.MODE_3D else (2 if plan.mode == P.MODE_TE else 1)
            pole_cells += ncomp * int((runs["n"] * np.array([npoles.get(int(o), 0) for o in runs["obj"]])).sum())
    psi_cells = 0
    for c in plan.cpml:
        # every CPML grid-list cell is an updated component cell (counted once: part 0 and part 1 cover the same cells)
        if c.part == 0 or not any(o.comp == c.comp and o.part == 0 for o in plan.cpml):
            comp_cells += int(c.grid["nAx"].sum())
        if c.has_psi:
            psi_cells += int(c.psi["transSz"].sum())
    emitter_bytes = sum(e.nemit * (e.nsys * 96 * e.nlevel * e.nlevel + 96) for e in plan.emitters)
    return Census(cells, comp_cells, d_cells, pole_cells, psi_cells, emitter_bytes)
