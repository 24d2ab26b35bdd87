"""y-slab decomposition: the ghost-row exchange protocol of one time step (reference GRID/parallelGrid.hpp:738-770 transferDat,
called from applBCH_/applBCE_/applBCOrDip_, FDTD_MANAGER/parallelFDTDField.hpp:1267-1269,1285-1287,1362-1363, and the E/P box
transfers of ML/parallelQE.hpp:618-645,694-715), reduced to the rows the stencils actually read.

Local rows: 0 = lower ghost, 1 .. ny = owned, ny+1 = upper ghost.  A step is four phases; after each phase the listed rows move:

  phase 0  H half step (+ CPML) and sources        ->  Hx, Hz  row ny  UP   (E update of the upper slab reads H at y-1)
  phase 1  oriented-dipole poles at the nodes      ->  node P_y (every pole)  row 1  DOWN   (D->E averages P_y[r], P_y[r+y])
  phase 2  E half step (+ CPML, D->E), emitter addP ->  Ex, Ez  row 1  DOWN  (H update of the lower slab reads E at y+1)
                                                       Ey      row ny UP    (node poles / emitters average Ey[r], Ey[r-y])
  phase 3  emitter density update                  ->  emitter P_y first row  DOWN  (addP of the lower slab's top row)

Periodic runs (CompCell.PBC, real fields) close the slabs into a ring: slab 0's lower neighbour is slab nranks-1 and vice versa, every
exchange above also crosses that seam, with two differences that come from the reference's wrap rows (applyBC1Proc: F[0] <- F[ymax-1],
F[ymax] <- F[1] with ymax = ln_y + 1, or ln_y for the components that are one row short in y: Ey, Hx, Hz):
  * the last slab's top owned row of Hx, Hz (and Ey) is ny - 1, not ny: that is the row it sends UP across the seam;
  * its row ny of Hx, Hz is the wrap image of slab 0's row 1 -- read by its own E update of row ny -- so after phase 0 slab 0 also
    sends Hx, Hz row 1 DOWN across the seam into that row.
The x / z ghost cells are wrapped inside every slab (applyBCProcMid; the checker and the engine do it after the H half step and at the end of
the step).  Emitters on a ring: the reference updates them BEFORE it wraps E, so across the seam their averages of Ey see the wrap rows of the
step before -- the two seam rows of Ey (last slab's row ny - 1 -> slab 0's row 0, slab 0's row 1 -> last slab's row ny) travel after phase 3;
emitter polarisation boxes do not cross the seam (the reference does not wrap them).

The CUDA engine implements this protocol natively (peer-to-peer stores + flags, chiml_gpu_halo_*); this module states it once in
host terms, drives the CPU checker through it in the world_size-2 gloo tests, and gathers slab results onto rank 0.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

from . import plan as P

EX, EY, EZ, HX, HY, HZ = range(6)
UP, DOWN = +1, -1

# (phase, kind, what, direction)
PROTOCOL: List[Tuple[int, str, Tuple[int, ...], int]] = [
    (0, "field", (HX, HZ), UP),
    (1, "ordip_py", (), DOWN),
    (2, "field", (EX, EZ), DOWN),
    (2, "field", (EY,), UP),
    (3, "emitter_py", (), DOWN),
]


def emitter_sends_down(plan: P.Plan, e: P.PlanEmitter) -> bool:
    """The slab's first owned row lies inside the emitter object's row range: its P_y row is the top rim of the slab below."""
    return plan.rank > 0 and e.box_lo[1] == 0


def emitter_receives_from_above(plan: P.Plan, e: P.PlanEmitter) -> bool:
    return plan.rank < plan.nranks - 1 and e.box_lo[1] + e.box_n[1] + 1 == plan.ln[1] - 1


def step_slab(sim, plan: P.Plan, amp_step: np.ndarray, send: Callable, recv: Callable) -> None:
    """One time step of one slab.  `sim` offers step_phase(phase, amp), field(f) -> (ly, lz, lx) view, ordip_pole(c, p, prev),
    emitter_P(slot, c) views (the CPU checker does); send(dst_rank, array) / recv(src_rank, out_array) move one row."""
    ny = plan.ln[1] - 2
    lower, upper = plan.rank - 1, plan.rank + 1
    has_lower, has_upper = lower >= 0, upper < plan.nranks
    ring = bool(plan.periodic) and plan.nranks > 1
    last = plan.rank == plan.nranks - 1
    if ring:
        lower, upper, has_lower, has_upper = lower % plan.nranks, upper % plan.nranks, True, True
    present = plan.fields_present()
    for phase in range(4):
        sim.step_phase(phase, amp_step)
        if ring and phase == 0:
            # the seam: slab 0's Hx, Hz row 1 is the wrap image the last slab's E update reads in its row ny
            for f in (HX, HZ):
                if f not in present:
                    continue
                if plan.rank == 0:
                    send(plan.nranks - 1, np.ascontiguousarray(sim.field(f)[1]))
                if last:
                    buf = np.empty_like(np.ascontiguousarray(sim.field(f)[ny]))
                    recv(0, buf)
                    sim.field(f)[ny] = buf
        if ring and phase == 3 and EY in present and plan.emitters:
            # emitters on a ring: the reference updates the emitters BEFORE it wraps E (step() items 16, 17), so their averages Ey[r], Ey[r - y] see
            # the wrap rows of Ey as of the step before: the two seam rows of Ey travel at the end of the step, not with the E half step
            if last:
                send(0, np.ascontiguousarray(sim.field(EY)[ny - 1]))
                buf = np.empty_like(np.ascontiguousarray(sim.field(EY)[ny])); recv(0, buf); sim.field(EY)[ny] = buf
            if plan.rank == 0:
                buf = np.empty_like(np.ascontiguousarray(sim.field(EY)[0])); recv(plan.nranks - 1, buf); sim.field(EY)[0] = buf
                send(plan.nranks - 1, np.ascontiguousarray(sim.field(EY)[1]))
        for ph, kind, fields, direction in PROTOCOL:
            if ph != phase:
                continue
            rows = []   # (array_view, send_row, recv_row)
            if kind == "field":
                # (ring: the last slab's components that are one row short in y end at row ny - 1)
                rows = [(sim.field(f), (ny - 1 if ring and last and f in (EY, HX, HZ) else ny) if direction == UP else 1, 0 if direction == UP else ny + 1)
                        for f in fields if f in present]
            elif kind == "ordip_py":
                if 1 in present:
                    rows = [(sim.ordip_pole(1, p, 0), 1, ny + 1) for p in range(plan.n_ordip_poles)]
            elif kind == "emitter_py":
                if 1 in present:
                    for q, e in enumerate(plan.emitters):
                        if emitter_sends_down(plan, e) or emitter_receives_from_above(plan, e):
                            rows.append((sim.emitter_P(q, 1), 1 if emitter_sends_down(plan, e) else None,
                                         e.box_n[1] + 1 if emitter_receives_from_above(plan, e) else None))
            seam_ey = ring and kind == "field" and fields == (EY,) and bool(plan.emitters)
            for arr, srow, rrow in rows:
                to, frm = (upper, lower) if direction == UP else (lower, upper)
                can_send = (has_upper if direction == UP else has_lower) and srow is not None and not (seam_ey and last)
                can_recv = (has_lower if direction == UP else has_upper) and rrow is not None and not (seam_ey and plan.rank == 0)
                # even ranks send first, odd ranks receive first: no deadlock with blocking point-to-point calls
                ops = [("s", can_send), ("r", can_recv)] if plan.rank % 2 == 0 else [("r", can_recv), ("s", can_send)]
                for op, ok in ops:
                    if not ok:
                        continue
                    if op == "s":
                        send(to, np.ascontiguousarray(arr[srow]))
                    else:
                        buf = np.empty_like(np.ascontiguousarray(arr[rrow]))
                        recv(frm, buf)
                        arr[rrow] = buf
