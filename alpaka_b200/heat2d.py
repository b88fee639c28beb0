"""example/heatEquation2D through the C ABI: the host-side mirror of the reference driver
(reference: example/heatEquation2D/src/heatEquation2D.cpp:34-203). One fused kernel per FTCS step instead of the
reference's Stencil + Boundary pair; the boundary's transcendental factors are computed HERE on the host (numpy ->
glibc sin/exp, the libm the reference CPU back-end uses) and only multiplied/added on the device."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import B200Error, check
from .runtime import Buf, Dev, Queue, memcpy

EDGE_TOP, EDGE_BOTTOM, EDGE_LEFT, EDGE_RIGHT, EDGE_ALL = 1, 2, 4, 8, 15


def exact_solution(x, y, t):
    """analyticalSolution.hpp:17-21. NB the evaluation order: exp(((-pi)*pi)*t) * (sin(pi*x) + sin(pi*y))."""
    pi = math.pi
    return math.exp(-pi * pi * t) * (math.sin(pi * x) + math.sin(pi * y))


def boundary_tables(ny: int, nx: int, dx: float, dy: float, j_offset: int = 0, i_offset: int = 0):
    """sx[i] = sin(pi*(i*dx)), sy[j] = sin(pi*(j*dy)) for the padded index range (optionally a sub-domain window).
    math.sin is glibc's sin: the same bits the reference's std::sin produces on this host."""
    pi = math.pi
    sx = np.array([math.sin(pi * ((i + i_offset) * dx)) for i in range(nx + 2)], dtype=np.float64)
    sy = np.array([math.sin(pi * ((j + j_offset) * dy)) for j in range(ny + 2)], dtype=np.float64)
    return sx, sy


def time_factor(step: int, dt: float) -> float:
    pi = math.pi
    return math.exp(-pi * pi * (step * dt))


def initial_field(ny: int, nx: int, dx: float, dy: float) -> np.ndarray:
    """initalizeBuffer (analyticalSolution.hpp:58-71): exactSolution(i*dx, j*dy, 0) = 1 * (sx[i] + sy[j])."""
    sx, sy = boundary_tables(ny, nx, dx, dy)
    e0 = math.exp(-math.pi * math.pi * 0.0)
    return e0 * (sx[None, :] + sy[:, None])


def stability_ratio(dx: float, dy: float, dt: float) -> float:
    """heatEquation2D.cpp:67: must be <= 1."""
    return 2 * dt / ((dx * dx * dy * dy) / (dx * dx + dy * dy))


class Heat2D:
    """Ping-pong pair of (ny+2) x (nx+2) device fields and the step loop of heatEquation2D.cpp:141-182."""

    def __init__(self, queue: Queue, ny: int, nx: int, dx: float, dy: float, dt: float, *, edges: int = EDGE_ALL,
                 j_offset: int = 0, i_offset: int = 0, ipc: bool = False):
        if ny < 1 or nx < 1:
            raise B200Error(-1, "heat2d: extent must be at least 1 x 1")
        r = stability_ratio(dx, dy, dt)
        if r > 1.0:
            raise B200Error(-1, f"Stability condition check failed: dt/min(dx^2,dy^2) = {r}, it is required to be <= 0.5")
        self.queue, self.dev = queue, queue.dev
        self.ny, self.nx, self.dx, self.dy, self.dt = ny, nx, dx, dy, dt
        self.rx = dt / (dx * dx)  # StencilKernel.hpp:70
        self.ry = dt / (dy * dy)  # StencilKernel.hpp:71
        self.bufs = [Buf(self.dev, np.float64, (ny + 2, nx + 2), queue, ipc=ipc) for _ in range(2)]
        self.cur = 0
        self.step_index = 0  # completed steps; the driver numbers steps from 1
        sx, sy = boundary_tables(ny, nx, dx, dy, j_offset, i_offset)
        plan = C.c_void_p()
        check(
            _lib.load().b200_heat2d_plan_create(
                self.dev.idx, self.bufs[0].ptr, self.bufs[1].ptr, self.bufs[0].pitch_bytes, ny, nx,
                sx.ctypes.data, sy.ctypes.data, edges, C.byref(plan),
            )
        )
        self.plan = plan.value
        self.edges = edges

    def upload(self, field: np.ndarray) -> None:
        """Both buffers start as copies of the field (see oracle/ref_heat2d.cpp on corners)."""
        from .runtime import HostBuf

        if not isinstance(field, HostBuf):  # a pinned HostBuf goes to the device as it is (asynchronous copy)
            field = np.ascontiguousarray(field, dtype=np.float64)
        shape = tuple(field.extent) if isinstance(field, HostBuf) else field.shape
        if shape != (self.ny + 2, self.nx + 2):
            raise B200Error(-1, "heat2d: field must be (ny+2) x (nx+2)")
        memcpy(self.queue, self.bufs[0], field)  # one trip over PCIe ...
        memcpy(self.queue, self.bufs[1], self.bufs[0])  # ... the second copy is device to device
        self.queue.wait()
        self.cur = 0

    #: time levels per launch that `step(n, fuse=True)` aims for (decomp.SUPPORTED_DEPTHS); tests and tools override it per call.
    #: Measured over ~1000 steps under the power cap, us per step (profiles/r02/heat_depth_probe.log, heat_walk_probe.log):
    #:   16384^2      tile kernel 4 levels 204-215 | walker 4 levels 186-195, 6 levels 182-189, 8 levels 174-185
    #:   2048 x 16384 tile kernel 4 levels 28.5    | walker 4 levels 26.1-26.7, 6 levels 29.0-29.5, 8 levels 36-39
    #: -> four levels by default, eight on fields tall enough for long walks (DEEP_FUSE from DEEP_MIN_ROWS rows on)
    DEFAULT_FUSE = 4
    DEEP_FUSE = 8
    DEEP_MIN_ROWS = 12288

    def default_depth(self) -> int:
        return self.DEEP_FUSE if self.ny >= self.DEEP_MIN_ROWS and self.nx >= 4096 else self.DEFAULT_FUSE

    def step(self, n: int = 1, *, fuse=True) -> None:
        """n FTCS steps. `fuse` (stand-alone fields): up to `fuse` steps (True = default_depth(), False = 1) go through ONE
        launch that keeps the intermediate time levels in registers -- b200_heat2d_step2_f64 (2 levels) or
        b200_heat2d_stepn_f64 (3, 4, 6, 8): HBM traffic of one step, same bits. A remainder runs in shallower launches.
        NB after a fused launch the current field is in the buffer one swap away, whatever the number of levels."""
        lib = _lib.load()
        depth = self.default_depth() if fuse is True else (1 if fuse is False else int(fuse))
        if not 1 <= depth <= 8:
            raise B200Error(-1, "heat2d: between 1 and 8 time levels per launch")
        if self.edges != EDGE_ALL:
            depth = 1
        from .decomp import launch_schedule

        for k in launch_schedule(n, depth):  # e.g. 3, 3, ..., then 4 = 2 + 2 rather than 3 + 1
            tfs = [time_factor(self.step_index + 1 + l, self.dt) for l in range(k)]
            if k == 1:
                check(lib.b200_heat2d_step_f64(self.plan, self.queue.handle, self.cur, self.rx, self.ry, tfs[0]))
            elif k == 2:
                check(lib.b200_heat2d_step2_f64(self.plan, self.queue.handle, self.cur, self.rx, self.ry, tfs[0], tfs[1]))
            else:
                arr = (C.c_double * k)(*tfs)
                check(lib.b200_heat2d_stepn_f64(self.plan, self.queue.handle, self.cur, self.rx, self.ry, k, arr))
            self.step_index += k
            self.cur ^= 1  # std::swap(uNextBufAcc, uCurrBufAcc), heatEquation2D.cpp:181
        self.queue._after_enqueue()

    def step_window(self, j0: int, j1: int, i0: int, i1: int, *, advance: bool, queue: Queue | None = None) -> None:
        """One step restricted to output cells [j0,j1) x [i0,i1); `advance` swaps buffers and counts the step."""
        q = queue or self.queue
        tf = time_factor(self.step_index + 1, self.dt)
        check(_lib.load().b200_heat2d_step_window_f64(self.plan, q.handle, self.cur, self.rx, self.ry, tf, j0, j1, i0, i1))
        if advance:
            self.step_index += 1
            self.cur ^= 1

    def current(self) -> Buf:
        return self.bufs[self.cur]

    def download(self, out=None):
        """The current field on the host; `out` may be a pinned HostBuf of the field's shape (returned as numpy)."""
        from .runtime import HostBuf

        if out is None:
            out = np.empty((self.ny + 2, self.nx + 2), dtype=np.float64)
        self.queue.wait()
        memcpy(self.queue, out, self.current())
        self.queue.wait()
        return out.array if isinstance(out, HostBuf) else out

    def close(self) -> None:
        if getattr(self, "plan", None):
            _lib.load().b200_heat2d_plan_destroy(self.plan)
            self.plan = None
        for b in getattr(self, "bufs", []):
            b.free()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def validate_solution(field: np.ndarray, dx: float, dy: float, t_max: float) -> float:
    """validateSolution (analyticalSolution.hpp:31-51): max-abs error over core cells vs the analytic field."""
    ny, nx = field.shape[0] - 2, field.shape[1] - 2
    pi = math.pi
    sx = np.sin(pi * (np.arange(nx + 2) * dx))
    sy = np.sin(pi * (np.arange(ny + 2) * dy))
    exact = math.exp(-pi * pi * t_max) * (sx[None, :] + sy[:, None])
    return float(np.max(np.abs(field[1:-1, 1:-1] - exact[1:-1, 1:-1])))
