"""Multi-GPU layer: slab-sharded streams/Dot/reduce and the 2-D decomposed heatEquation2D with the halo exchange
fused into the step kernel (the reference has no multi-device driver).

One process per GPU (torchrun); `torch.distributed` is plumbing only: it carries the CUDA-IPC handles once at set-up
and the single Dot/reduce scalar per rank. The heat halos never touch NCCL or the host: each rank's step kernel stores
its border cells straight into the neighbours' ghost cells through IPC-mapped peer pointers (NVLink) and publishes
the time level in a flag word (b200_heat2d_step_halo_f64, include/b200/b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib, decomp, heat2d
from ._lib import B200Error, Heat2dHalo, check
from .decomp import OPPOSITE, SIDES, Tile
from .runtime import Buf, Queue, memcpy


class _HaloWiring:
    """What HeatTile and HeatSlab share: exporting / mapping the two field buffers and the flag words, and handing the
    neighbours' pointers to the plan (b200_heat2d_plan_set_halo). Subclasses provide `tile` (geometry with `neighbours`,
    `shape`, `rank`), `flags`, `dev`, `_field_bufs()` and `_plan_handle()`."""

    def _field_bufs(self):
        raise NotImplementedError

    def _plan_handle(self):
        raise NotImplementedError

    def local_pointers(self) -> dict:
        u0, u1 = self._field_bufs()
        return {"u0": u0.ptr, "u1": u1.ptr, "flags": self.flags.ptr, "dev": self.dev.idx, "shape": self.tile.shape,
                "pitch": u0.pitch_bytes}

    def export(self) -> dict:
        """CUDA-IPC handles of the two field buffers and the flag words (picklable)."""
        lib = _lib.load()
        u0, u1 = self._field_bufs()
        out = {"shape": self.tile.shape, "pitch": u0.pitch_bytes, "rank": self.tile.rank}
        for name, ptr in (("u0", u0.ptr), ("u1", u1.ptr), ("flags", self.flags.ptr)):
            hb = C.create_string_buffer(64)
            check(lib.b200_ipc_get_mem_handle(self.dev.idx, ptr, hb))
            out[name] = hb.raw
        return out

    def open_peer(self, exported: dict) -> dict:
        """Maps a neighbour's exported buffers into this process; returns raw pointers like local_pointers()."""
        lib = _lib.load()
        out = {"shape": tuple(exported["shape"]), "pitch": exported["pitch"]}
        for name in ("u0", "u1", "flags"):
            p = C.c_void_p()
            check(lib.b200_ipc_open_mem_handle(self.dev.idx, exported[name], C.byref(p)))
            out[name] = p.value
            self._opened.append(p.value)
        return out

    def connect(self, peers: dict) -> None:
        """peers: side -> pointers dict of the neighbour on that side (None / missing on physical boundaries)."""
        halo = Heat2dHalo()
        pitch = self._field_bufs()[0].pitch_bytes
        for k, side in enumerate(SIDES):
            nb = peers.get(side)
            if (nb is None) != (self.tile.neighbours[side] is None):
                raise B200Error(-1, f"heat tile {self.tile.rank}: neighbour on side '{side}' does not match the decomposition")
            if nb is None:
                continue
            if tuple(nb["shape"]) != tuple(self.tile.shape) or nb["pitch"] != pitch:
                raise B200Error(-1, "heat tiles must have identical extents and pitches")
            halo.peer_u[k][0] = nb["u0"]
            halo.peer_u[k][1] = nb["u1"]
            halo.peer_flag[k] = nb["flags"] + 4 * SIDES.index(OPPOSITE[side])
        halo.my_flags = self.flags.ptr
        check(_lib.load().b200_heat2d_plan_set_halo(self._plan_handle(), C.byref(halo)))
        self.connected = True

    def status(self) -> int:
        """0, or 1 + side of the first flag wait that timed out."""
        s = C.c_uint32(0)
        check(_lib.load().b200_heat2d_halo_status(self._plan_handle(), C.byref(s)))
        return int(s.value)

    def raise_on_timeout(self) -> None:
        """A neighbour's flag did not arrive within `exchange.timeout_ms`: ghost cells were read stale, the field is invalid."""
        s = self.status()
        if s != 0:
            raise B200Error(-1, f"heat halo exchange on rank {self.tile.rank}: the neighbour on side '{SIDES[s - 1]}' never "
                                f"published its time level (flag wait timed out); the field is invalid")

    def _close_peers(self) -> None:
        lib = _lib.load()
        for p in self._opened:
            lib.b200_ipc_close_mem_handle(self.dev.idx, p)
        self._opened = []


class HeatTile(_HaloWiring):
    """One rank's tile of the decomposed field: ping-pong buffers + flag words (both IPC-exportable), the plan, and the
    step loop. Usage: construct on every rank, exchange `export()` (or `local_pointers()` inside one process), call
    `connect()`, `upload()` the tile's window of the initial field, then `step(n)`."""

    def __init__(self, queue: Queue, tile: Tile, NY: int, NX: int, dt: Optional[float] = None):
        self.queue, self.dev, self.tile = queue, queue.dev, tile
        self.NY, self.NX = NY, NX
        self.dx, self.dy = 1.0 / (NX + 1), 1.0 / (NY + 1)  # heatEquation2D.cpp:62-63 on the GLOBAL grid
        self.dt = 0.2 * min(self.dx * self.dx, self.dy * self.dy) if dt is None else dt
        self.h = heat2d.Heat2D(queue, tile.ny, tile.nx, self.dx, self.dy, self.dt, edges=tile.edges,
                               j_offset=tile.j_offset, i_offset=tile.i_offset, ipc=True)
        self.flags = Buf(self.dev, np.uint32, 16, ipc=True)
        lib = _lib.load()
        check(lib.b200_memset_async(self.dev.idx, self.flags.ptr, 0, 64, queue.handle))
        queue.wait()
        self._opened: list[int] = []
        self.connected = False

    def _field_bufs(self):
        return self.h.bufs

    def _plan_handle(self):
        return self.h.plan

    # ---- data
    def upload(self, local_field: np.ndarray) -> None:
        self.h.upload(local_field)

    def initial_field(self) -> np.ndarray:
        """This tile's window of the global initial field exactSolution(i*dx, j*dy, 0), ghosts included."""
        import math

        sx, sy = heat2d.boundary_tables(self.tile.ny, self.tile.nx, self.dx, self.dy, self.tile.j_offset, self.tile.i_offset)
        return math.exp(-math.pi * math.pi * 0.0) * (sx[None, :] + sy[:, None])

    def step(self, n: int = 1) -> None:
        lib = _lib.load()
        h = self.h
        if not self.connected:
            raise B200Error(-1, "HeatTile.step before connect()")
        for _ in range(n):
            h.step_index += 1
            tf = heat2d.time_factor(h.step_index, h.dt)
            check(lib.b200_heat2d_step_halo_f64(h.plan, self.queue.handle, h.cur, h.rx, h.ry, tf, h.step_index))
            h.cur ^= 1
        self.queue._after_enqueue()

    def download(self) -> np.ndarray:
        out = self.h.download()
        self.raise_on_timeout()
        return out

    def close(self) -> None:
        self._close_peers()
        self.h.close()
        self.flags.free()


class HeatSlab(_HaloWiring):
    """One rank's ROW SLAB of the field, advanced `levels` (2, 3 or 4) time levels per launch and per exchange
    (b200_heat2d_slab_plan_create + b200_heat2d_step2_halo_f64 / b200_heat2d_stepn_halo_f64, include/b200/b200.h): the
    temporal blocking of the stand-alone kernels carried to several GPUs. Ghost rows are G = `levels` deep, so the
    local array is (ny+2G) x (nx+2): rows 0..G-1 / ny+G.. are ghosts (or ring + unused rows on a physical side), core
    rows are G..ny+G-1. Same wiring calls as HeatTile (export / open_peer / connect, connect_over_process_group,
    connect_in_process)."""

    DEFAULT_LEVELS = 4

    def __init__(self, queue: Queue, rank: int, world: int, NY: int, NX: int, dt: Optional[float] = None,
                 levels: Optional[int] = None):
        import math

        G = self.DEFAULT_LEVELS if levels is None else int(levels)
        if G not in (2, 3, 4, 6, 8):
            raise B200Error(-1, "heat slabs advance 2, 3, 4, 6 or 8 time levels per launch")
        try:
            self.tile = decomp.slab_for(rank, world, NY, NX, G)  # geometry: pure host logic, tested on CPU
        except ValueError as e:
            raise B200Error(-1, str(e)) from None
        self.queue, self.dev = queue, queue.dev
        self.NY, self.NX, self.levels = NY, NX, G
        ny, nx = NY // world, NX
        self.dx, self.dy = 1.0 / (NX + 1), 1.0 / (NY + 1)
        self.dt = 0.2 * min(self.dx * self.dx, self.dy * self.dy) if dt is None else dt
        if heat2d.stability_ratio(self.dx, self.dy, self.dt) > 1.0:
            raise B200Error(-1, "Stability condition check failed")
        self.rx, self.ry = self.dt / (self.dx * self.dx), self.dt / (self.dy * self.dy)
        self.bufs = [Buf(self.dev, np.float64, (ny + 2 * G, nx + 2), queue, ipc=True) for _ in range(2)]
        self.cur, self.step_index, self.launch_index = 0, 0, 0
        pi = math.pi
        self.g0 = self.tile.g0  # global padded row of local row 0 (may be negative: unused rows)
        self.sx = np.array([math.sin(pi * (i * self.dx)) for i in range(nx + 2)], dtype=np.float64)
        self.sy = np.array([math.sin(pi * ((self.g0 + j) * self.dy)) for j in range(ny + 2 * G)], dtype=np.float64)
        plan = C.c_void_p()
        lib = _lib.load()
        check(lib.b200_heat2d_slab_plan_create(self.dev.idx, self.bufs[0].ptr, self.bufs[1].ptr, self.bufs[0].pitch_bytes,
                                               ny, nx, self.sx.ctypes.data, self.sy.ctypes.data, self.tile.edges, G,
                                               C.byref(plan)))
        self.plan = plan.value
        self.flags = Buf(self.dev, np.uint32, 16, ipc=True)
        check(lib.b200_memset_async(self.dev.idx, self.flags.ptr, 0, 64, queue.handle))
        queue.wait()
        self._opened = []
        self.connected = False

    def _field_bufs(self):
        return self.bufs

    def _plan_handle(self):
        return self.plan

    # ---- data
    def window(self, global_field: np.ndarray) -> np.ndarray:
        """This slab's (ny+2G) x (nx+2) window of a global (NY+2) x (NX+2) padded field; rows outside the field are 0."""
        return self.tile.window(global_field)

    def initial_field(self) -> np.ndarray:
        import math

        return math.exp(-math.pi * math.pi * 0.0) * (self.sx[None, :] + self.sy[:, None])

    def upload(self, local_field: np.ndarray) -> None:
        local_field = np.ascontiguousarray(local_field, dtype=np.float64)
        if local_field.shape != self.tile.shape:
            raise B200Error(-1, "heat slab: field must be (ny+2G) x (nx+2)")
        memcpy(self.queue, self.bufs[0], local_field)
        memcpy(self.queue, self.bufs[1], local_field)
        self.queue.wait()
        self.cur = 0

    def step(self, n: Optional[int] = None) -> None:
        """n FTCS steps (default: one launch of `levels`): launches of `levels` time levels each, the remainder in
        shallower ones (every launch refreshes all G ghost rows, so depths can be mixed; 4 = 2 + 2 rather than 3 + 1).
        A slab cannot advance a single level: n = 1, or an odd n at levels = 2, is refused."""
        G = self.levels
        n = G if n is None else n
        try:
            schedule = decomp.launch_schedule(n, G, min_depth=2)
        except ValueError as e:
            raise B200Error(-1, f"heat slab with ghost rows {G} deep: {e}") from None
        if not self.connected:
            raise B200Error(-1, "HeatSlab.step before connect()")
        lib = _lib.load()
        for k in schedule:
            self.launch_index += 1
            tfs = [heat2d.time_factor(self.step_index + 1 + l, self.dt) for l in range(k)]
            if k == 2:
                check(lib.b200_heat2d_step2_halo_f64(self.plan, self.queue.handle, self.cur, self.rx, self.ry, tfs[0], tfs[1],
                                                     self.launch_index))
            else:
                arr = (C.c_double * k)(*tfs)
                check(lib.b200_heat2d_stepn_halo_f64(self.plan, self.queue.handle, self.cur, self.rx, self.ry, k, arr,
                                                     self.launch_index))
            self.step_index += k
            self.cur ^= 1
        self.queue._after_enqueue()

    def download(self) -> np.ndarray:
        out = np.empty(self.tile.shape, dtype=np.float64)
        self.queue.wait()
        memcpy(self.queue, out, self.bufs[self.cur])
        self.queue.wait()
        self.raise_on_timeout()
        return out

    def owned_rows(self) -> tuple[int, int]:
        """[j0, j1) of the local rows this slab OWNS: its core rows, plus the physical ring row on a boundary side."""
        return self.tile.owned_rows()

    def stitch(self, global_out: np.ndarray, local_field: np.ndarray) -> None:
        """Writes the rows this slab owns into the global (NY+2) x (NX+2) field."""
        self.tile.stitch(global_out, local_field)

    def close(self) -> None:
        self._close_peers()
        if getattr(self, "plan", None):
            _lib.load().b200_heat2d_plan_destroy(self.plan)
            self.plan = None
        for b in self.bufs:
            b.free()
        self.bufs = []
        self.flags.free()


class HeatTileDeep(_HaloWiring):
    """One rank's tile of a Py x Px decomposition advanced `levels` (4, 6 or 8) time levels per launch: ghost cells
    `levels` deep on all four sides (b200_heat2d_tile_plan_create + b200_heat2d_stepn_tile_f64, include/b200/b200.h). The
    exchange has two phases per launch: the rows travel inside the walker launch (peer stores from its strips, as for
    HeatSlab), then a small column kernel stores the first / last G core columns -- over all rows, the freshly received
    ghost rows included, so the corner blocks reach the diagonal neighbours through the vertical ones -- into the left /
    right neighbours and publishes the launch index in their column flags. Same wiring calls as HeatTile."""

    def __init__(self, queue: Queue, rank: int, world: int, NY: int, NX: int, dt: Optional[float] = None, levels: int = 4,
                 grid: Optional[tuple] = None):
        import math

        G = int(levels)
        if G not in (4, 6, 8):
            raise B200Error(-1, "deep heat tiles advance 4, 6 or 8 time levels per launch")
        try:
            self.tile = decomp.deep_tile_for(rank, world, NY, NX, G, grid)
        except ValueError as e:
            raise B200Error(-1, str(e)) from None
        self.queue, self.dev = queue, queue.dev
        self.NY, self.NX, self.levels = NY, NX, G
        ny, nx = self.tile.ny, self.tile.nx
        self.dx, self.dy = 1.0 / (NX + 1), 1.0 / (NY + 1)
        self.dt = 0.2 * min(self.dx * self.dx, self.dy * self.dy) if dt is None else dt
        if heat2d.stability_ratio(self.dx, self.dy, self.dt) > 1.0:
            raise B200Error(-1, "Stability condition check failed")
        self.rx, self.ry = self.dt / (self.dx * self.dx), self.dt / (self.dy * self.dy)
        self.bufs = [Buf(self.dev, np.float64, (ny + 2 * G, nx + 2 * G), queue, ipc=True) for _ in range(2)]
        self.cur, self.step_index, self.launch_index = 0, 0, 0
        pi = math.pi
        self.sx = np.array([math.sin(pi * ((self.tile.gi0 + i) * self.dx)) for i in range(nx + 2 * G)], dtype=np.float64)
        self.sy = np.array([math.sin(pi * ((self.tile.gj0 + j) * self.dy)) for j in range(ny + 2 * G)], dtype=np.float64)
        plan = C.c_void_p()
        lib = _lib.load()
        check(lib.b200_heat2d_tile_plan_create(self.dev.idx, self.bufs[0].ptr, self.bufs[1].ptr, self.bufs[0].pitch_bytes,
                                               ny, nx, self.sx.ctypes.data, self.sy.ctypes.data, self.tile.edges, G,
                                               C.byref(plan)))
        self.plan = plan.value
        self.flags = Buf(self.dev, np.uint32, 16, ipc=True)
        check(lib.b200_memset_async(self.dev.idx, self.flags.ptr, 0, 64, queue.handle))
        queue.wait()
        self._opened = []
        self.connected = False

    def _field_bufs(self):
        return self.bufs

    def _plan_handle(self):
        return self.plan

    def window(self, global_field: np.ndarray) -> np.ndarray:
        return self.tile.window(global_field)

    def initial_field(self) -> np.ndarray:
        import math

        return math.exp(-math.pi * math.pi * 0.0) * (self.sx[None, :] + self.sy[:, None])

    def upload(self, local_field: np.ndarray) -> None:
        local_field = np.ascontiguousarray(local_field, dtype=np.float64)
        if local_field.shape != self.tile.shape:
            raise B200Error(-1, "deep heat tile: field must be (ny+2G) x (nx+2G)")
        memcpy(self.queue, self.bufs[0], local_field)
        memcpy(self.queue, self.bufs[1], local_field)
        self.queue.wait()
        self.cur = 0

    def step(self, n: Optional[int] = None) -> None:
        """n FTCS steps (default: one launch of `levels`) in launches of 4, 6 or 8 levels, none deeper than the ghost cells."""
        G = self.levels
        n = G if n is None else n
        try:
            schedule = decomp.launch_schedule(n, G, min_depth=4, depths=(4, 6, 8))
        except ValueError as e:
            raise B200Error(-1, f"deep heat tile with ghost cells {G} deep: {e}") from None
        if not self.connected:
            raise B200Error(-1, "HeatTileDeep.step before connect()")
        lib = _lib.load()
        for k in schedule:
            self.launch_index += 1
            tfs = [heat2d.time_factor(self.step_index + 1 + l, self.dt) for l in range(k)]
            arr = (C.c_double * k)(*tfs)
            check(lib.b200_heat2d_stepn_tile_f64(self.plan, self.queue.handle, self.cur, self.rx, self.ry, k, arr, self.launch_index))
            self.step_index += k
            self.cur ^= 1
        self.queue._after_enqueue()

    def download(self) -> np.ndarray:
        out = np.empty(self.tile.shape, dtype=np.float64)
        self.queue.wait()
        memcpy(self.queue, out, self.bufs[self.cur])
        self.queue.wait()
        self.raise_on_timeout()
        return out

    def stitch(self, global_out: np.ndarray, local_field: np.ndarray) -> None:
        self.tile.stitch(global_out, local_field)

    def close(self) -> None:
        self._close_peers()
        if getattr(self, "plan", None):
            _lib.load().b200_heat2d_plan_destroy(self.plan)
            self.plan = None
        for b in self.bufs:
            b.free()
        self.bufs = []
        self.flags.free()


def connect_over_process_group(tile_runner: HeatTile, dist) -> None:
    """One process per GPU: all-gather the IPC handles (objects, once) and map the neighbours' buffers."""
    world = dist.get_world_size()
    exported = [None] * world
    dist.all_gather_object(exported, tile_runner.export())
    peers = {}
    for side in SIDES:
        r = tile_runner.tile.neighbours[side]
        peers[side] = None if r is None else tile_runner.open_peer(exported[r])
    tile_runner.connect(peers)
    dist.barrier()


def connect_in_process(runners: list) -> None:
    """One process driving several tiles (several devices with peer access enabled, or -- in tests -- several tiles on
    one device): plain device pointers, no IPC."""
    ptrs = [r.local_pointers() for r in runners]
    for r in runners:
        peers = {side: (None if r.tile.neighbours[side] is None else ptrs[r.tile.neighbours[side]]) for side in SIDES}
        r.connect(peers)


class ScalarExchange:
    """One rank's end of the Dot / reduce exchange FUSED into the reduction launch (b200_dot_allranks_* /
    b200_reduce_sum_allranks_*, include/b200/b200.h): the last block of the single-pass reduction stores this device's
    scalar into every rank's slot array through peer pointers, publishes the call number, waits for the others and folds
    the slots in rank order -- the all-ranks value comes out of the same launch, bit-identical on every rank, without NCCL
    or the host. Usage: construct on every rank, connect_exchange_over_process_group(ex, dist) (or
    connect_exchange_in_process([...])), then ex.dot(queue, a, b) / ex.reduce_sum(queue, x) collectively."""

    _ALLRANKS = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32", np.dtype(np.uint32): "u32"}

    def __init__(self, queue: Queue, rank: int, world: int):
        if not (1 <= world <= 16 and 0 <= rank < world):
            raise B200Error(-1, "ScalarExchange: between 1 and 16 ranks")
        self.dev, self.rank, self.world = queue.dev, rank, world
        self.buf = Buf(self.dev, np.uint8, _lib.EXCHANGE_BYTES, ipc=True)
        check(_lib.load().b200_memset_async(self.dev.idx, self.buf.ptr, 0, _lib.EXCHANGE_BYTES, queue.handle))
        queue.wait()
        self.step = 0
        self._opened: list[int] = []
        self.desc = None

    def export(self) -> bytes:
        hb = C.create_string_buffer(64)
        check(_lib.load().b200_ipc_get_mem_handle(self.dev.idx, self.buf.ptr, hb))
        return hb.raw

    def open_peer(self, handle: bytes) -> int:
        p = C.c_void_p()
        check(_lib.load().b200_ipc_open_mem_handle(self.dev.idx, handle, C.byref(p)))
        self._opened.append(p.value)
        return p.value

    def connect(self, bases: list) -> None:
        """bases[r]: rank r's exchange buffer as seen from this device (own pointer at index `rank`)."""
        if len(bases) != self.world or bases[self.rank] != self.buf.ptr:
            raise B200Error(-1, "ScalarExchange.connect: one pointer per rank, the own buffer at the own rank")
        d = _lib.Exchange()
        for r, ptr in enumerate(bases):
            d.base[r] = ptr
        d.world, d.rank = self.world, self.rank
        self.desc = d

    def _call(self, name: str, queue: Queue, *args) -> None:
        if self.desc is None:
            raise B200Error(-1, "ScalarExchange used before connect()")
        self.step += 1
        check(getattr(_lib.load(), name)(queue.handle, *args, queue.reduce_scratch(), C.byref(self.desc), self.step))
        queue._after_enqueue()

    def dot_async(self, queue: Queue, a: Buf, b: Buf, out: Buf, n: Optional[int] = None) -> None:
        """Collective: out[0] (device) = sum over all ranks of sum_i a[i]*b[i] of the rank's slab."""
        sfx = self._ALLRANKS.get(a.dtype)
        if sfx not in ("f64", "f32") or b.dtype != a.dtype or out.dtype != a.dtype:
            raise B200Error(-1, "dot over all ranks: float32 or float64 buffers of one type")
        n = a.extent[0] if n is None else n
        self._call(f"b200_dot_allranks_{sfx}", queue, a.ptr, b.ptr, n, out.ptr)

    def reduce_sum_async(self, queue: Queue, source: Buf, out: Buf, n: Optional[int] = None) -> None:
        sfx = self._ALLRANKS.get(source.dtype)
        if sfx is None or out.dtype != source.dtype:
            raise B200Error(-1, "reduce over all ranks: uint32, float32 or float64 buffers of one type")
        n = source.extent[0] if n is None else n
        self._call(f"b200_reduce_sum_allranks_{sfx}", queue, source.ptr, n, out.ptr)

    def _to_host(self, queue: Queue, out: Buf):
        host = np.empty(1, dtype=out.dtype)
        memcpy(queue, host, out)
        queue.wait()
        out.free()
        self.raise_on_timeout()
        return host[0]

    def raise_on_timeout(self) -> None:
        """A peer's flag did not arrive within `exchange.timeout_ms`: the value on this rank is not the all-ranks sum."""
        s = self.status()
        if s != 0:
            raise B200Error(-1, f"fused Dot/reduce exchange: rank {s - 1}'s scalar never arrived on rank {self.rank} "
                                f"(flag wait timed out); the result is invalid")

    def dot(self, queue: Queue, a: Buf, b: Buf, n: Optional[int] = None):
        from .runtime import alloc_buf

        out = alloc_buf(queue.dev, a.dtype, 1, queue)
        self.dot_async(queue, a, b, out, n)
        return self._to_host(queue, out)

    def reduce_sum(self, queue: Queue, source: Buf, n: Optional[int] = None):
        from .runtime import alloc_buf

        out = alloc_buf(queue.dev, source.dtype, 1, queue)
        self.reduce_sum_async(queue, source, out, n)
        return self._to_host(queue, out)

    def status(self) -> int:
        s = C.c_uint32(0)
        check(_lib.load().b200_exchange_status(self.dev.idx, self.buf.ptr, C.byref(s)))
        return int(s.value)

    def close(self) -> None:
        lib = _lib.load()
        for p in self._opened:
            lib.b200_ipc_close_mem_handle(self.dev.idx, p)
        self._opened = []
        self.buf.free()


def connect_exchange_over_process_group(ex: ScalarExchange, dist) -> None:
    """One process per GPU: all-gather the IPC handles once and map every other rank's exchange buffer."""
    handles = [None] * ex.world
    dist.all_gather_object(handles, ex.export())
    ex.connect([ex.buf.ptr if r == ex.rank else ex.open_peer(handles[r]) for r in range(ex.world)])
    dist.barrier()


def connect_exchange_in_process(exchanges: list) -> None:
    """One process driving all ranks (tests: several ranks on one device; or several peer-enabled devices)."""
    ptrs = [e.buf.ptr for e in exchanges]
    for e in exchanges:
        e.connect(list(ptrs))


def dot_all_ranks(local_value: float, dist, device) -> float:
    """Dot's exchange step: one double per rank, gathered (NCCL all_gather of 8 bytes) and summed in rank order."""
    import torch

    mine = torch.tensor([local_value], dtype=torch.float64, device=device)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return float(decomp.combine_in_rank_order([float(p.item()) for p in parts]))
