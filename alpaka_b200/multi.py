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


class HeatTile:
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

    # ---- wiring
    def local_pointers(self) -> dict:
        return {"u0": self.h.bufs[0].ptr, "u1": self.h.bufs[1].ptr, "flags": self.flags.ptr, "dev": self.dev.idx,
                "shape": self.tile.shape, "pitch": self.h.bufs[0].pitch_bytes}

    def export(self) -> dict:
        """CUDA-IPC handles of the two field buffers and the flag words (picklable)."""
        lib = _lib.load()
        out = {"shape": self.tile.shape, "pitch": self.h.bufs[0].pitch_bytes, "rank": self.tile.rank}
        for name, ptr in (("u0", self.h.bufs[0].ptr), ("u1", self.h.bufs[1].ptr), ("flags", self.flags.ptr)):
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
        for k, side in enumerate(SIDES):
            nb = peers.get(side)
            if (nb is None) != (self.tile.neighbours[side] is None):
                raise B200Error(-1, f"heat tile {self.tile.rank}: neighbour on side '{side}' does not match the decomposition")
            if nb is None:
                continue
            if tuple(nb["shape"]) != self.tile.shape or nb["pitch"] != self.h.bufs[0].pitch_bytes:
                raise B200Error(-1, "heat tiles must have identical extents and pitches")
            halo.peer_u[k][0] = nb["u0"]
            halo.peer_u[k][1] = nb["u1"]
            halo.peer_flag[k] = nb["flags"] + 4 * SIDES.index(OPPOSITE[side])
        halo.my_flags = self.flags.ptr
        check(_lib.load().b200_heat2d_plan_set_halo(self.h.plan, C.byref(halo)))
        self.connected = True

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

    def status(self) -> int:
        s = C.c_uint32(0)
        check(_lib.load().b200_heat2d_halo_status(self.h.plan, C.byref(s)))
        return int(s.value)

    def download(self) -> np.ndarray:
        return self.h.download()

    def close(self) -> None:
        lib = _lib.load()
        for p in self._opened:
            lib.b200_ipc_close_mem_handle(self.dev.idx, p)
        self._opened = []
        self.h.close()
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


def dot_all_ranks(local_value: float, dist, device) -> float:
    """Dot's exchange step: one double per rank, gathered (NCCL all_gather of 8 bytes) and summed in rank order."""
    import torch

    mine = torch.tensor([local_value], dtype=torch.float64, device=device)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return float(decomp.combine_in_rank_order([float(p.item()) for p in parts]))
