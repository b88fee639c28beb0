"""Host-side partitioning logic of the multi-GPU layer -- pure Python, no device needed, so the
world_size > 1 logic is testable on CPU with the gloo backend (tests/test_multi_rank_cpu.py).

  * streams / Dot / reduce: contiguous slabs, rank r owns [lo, hi) of every array; no exchange except one scalar.
  * heatEquation2D: Py x Px process grid over the NY x NX core cells; every rank owns an equal (ny+2) x (nx+2) tile
    whose outer ring is either the physical boundary (BoundaryKernel semantics) or ghost cells owned by a neighbour.
The reference has nothing of this (every driver uses device 0); parity is defined as "decomposed result == undecomposed
oracle result", bit for bit."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

EDGE_TOP, EDGE_BOTTOM, EDGE_LEFT, EDGE_RIGHT = 1, 2, 4, 8
SIDES = ("top", "bottom", "left", "right")
OPPOSITE = {"top": "bottom", "bottom": "top", "left": "right", "right": "left"}


def slab_bounds(n: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous slab of n elements; slab starts are multiples of `align` elements (16-byte
    boundaries for the vector loads), the last rank takes the remainder; lower ranks never get fewer elements."""
    if world < 1 or not (0 <= rank < world) or n < 0 or align < 1:
        raise ValueError("slab_bounds: bad arguments")
    per = -(-n // world)  # ceil
    per = -(-per // align) * align
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi


def process_grid(world: int) -> tuple[int, int]:
    """(Py, Px) with Py * Px == world, as square as possible, Py >= Px: rows are split first because a row halo is
    contiguous in memory and a column halo is strided. 1 -> 1x1, 2 -> 2x1, 4 -> 2x2, 8 -> 4x2."""
    if world < 1:
        raise ValueError("process_grid: world must be >= 1")
    px = int(world**0.5)
    while world % px != 0:
        px -= 1
    return world // px, px


@dataclass(frozen=True)
class Tile:
    """One rank's share of the heat field. Local padded coordinates [0, ny+2) x [0, nx+2); local [j][i] is global
    padded [j + j_offset][i + i_offset]."""

    rank: int
    py: int
    px: int
    cy: int
    cx: int
    ny: int
    nx: int
    j_offset: int
    i_offset: int
    edges: int
    neighbours: dict = field(default_factory=dict)  # side -> rank or None

    @property
    def shape(self) -> tuple[int, int]:
        return self.ny + 2, self.nx + 2


def tile_for(rank: int, world: int, NY: int, NX: int, grid: Optional[tuple[int, int]] = None) -> Tile:
    py, px = grid if grid is not None else process_grid(world)
    if py * px != world:
        raise ValueError("tile_for: process grid does not match the world size")
    if NY % py != 0 or NX % px != 0:
        raise ValueError(f"tile_for: {NY} x {NX} core cells do not divide over a {py} x {px} process grid")
    cy, cx = divmod(rank, px)
    ny, nx = NY // py, NX // px
    edges = 0
    nb: dict = {}
    nb["top"] = None if cy == 0 else (cy - 1) * px + cx
    nb["bottom"] = None if cy == py - 1 else (cy + 1) * px + cx
    nb["left"] = None if cx == 0 else cy * px + (cx - 1)
    nb["right"] = None if cx == px - 1 else cy * px + (cx + 1)
    for side, bit in zip(SIDES, (EDGE_TOP, EDGE_BOTTOM, EDGE_LEFT, EDGE_RIGHT)):
        if nb[side] is None:
            edges |= bit
    return Tile(rank, py, px, cy, cx, ny, nx, cy * ny, cx * nx, edges, nb)


def tile_view(global_field, tile: Tile):
    """The tile's (ny+2) x (nx+2) window of a global (NY+2) x (NX+2) padded field (ghosts included)."""
    return global_field[tile.j_offset : tile.j_offset + tile.ny + 2, tile.i_offset : tile.i_offset + tile.nx + 2]


def stitch(global_out, tile: Tile, local_field) -> None:
    """Writes the cells the tile OWNS into the global field: its core cells plus the physical ring on its boundary
    sides (ghost cells belong to the neighbours and are skipped)."""
    j0 = 0 if tile.edges & EDGE_TOP else 1
    j1 = tile.ny + 2 if tile.edges & EDGE_BOTTOM else tile.ny + 1
    i0 = 0 if tile.edges & EDGE_LEFT else 1
    i1 = tile.nx + 2 if tile.edges & EDGE_RIGHT else tile.nx + 1
    global_out[tile.j_offset + j0 : tile.j_offset + j1, tile.i_offset + i0 : tile.i_offset + i1] = local_field[j0:j1, i0:i1]


def halo_slices(tile: Tile, side: str):
    """(send, recv) index expressions in local padded coordinates for the exchange with the neighbour on `side`:
    `send` = my border core cells, `recv` = my ghost cells on that side."""
    ny, nx = tile.ny, tile.nx
    if side == "top":
        return (slice(1, 2), slice(1, nx + 1)), (slice(0, 1), slice(1, nx + 1))
    if side == "bottom":
        return (slice(ny, ny + 1), slice(1, nx + 1)), (slice(ny + 1, ny + 2), slice(1, nx + 1))
    if side == "left":
        return (slice(1, ny + 1), slice(1, 2)), (slice(1, ny + 1), slice(0, 1))
    if side == "right":
        return (slice(1, ny + 1), slice(nx, nx + 1)), (slice(1, ny + 1), slice(nx + 1, nx + 2))
    raise ValueError(side)


@dataclass(frozen=True)
class Slab:
    """One rank's ROW SLAB of the heat field for launches that advance several time levels: all NX columns, NY/world core
    rows, ghost rows `ghost` deep. Local array (ny + 2*ghost) x (nx+2); local row j is global padded row g0 + j (g0 may be
    negative on the first slab: rows outside the field are unused). Core rows are ghost .. ny+ghost-1; on a physical
    side the row next to them is the ring. Field names shared with Tile where the wiring code needs them."""

    rank: int
    world: int
    ny: int
    nx: int
    ghost: int
    edges: int
    neighbours: dict = field(default_factory=dict)

    @property
    def shape(self) -> tuple[int, int]:
        return self.ny + 2 * self.ghost, self.nx + 2

    @property
    def j_offset(self) -> int:
        return self.rank * self.ny

    @property
    def g0(self) -> int:
        return self.rank * self.ny - (self.ghost - 1)

    def owned_rows(self) -> tuple[int, int]:
        """[j0, j1) of the local rows this slab OWNS: its core rows, plus the physical ring row on a boundary side."""
        j0 = self.ghost - 1 if self.edges & EDGE_TOP else self.ghost
        j1 = self.ny + self.ghost + 1 if self.edges & EDGE_BOTTOM else self.ny + self.ghost
        return j0, j1

    def window(self, global_field):
        """This slab's local array cut out of a global (NY+2) x (NX+2) padded field; rows outside the field are 0."""
        import numpy as np

        rows, NYp = self.shape[0], self.ny * self.world + 2
        out = np.zeros((rows, self.nx + 2))
        lo, hi = max(self.g0, 0), min(self.g0 + rows, NYp)
        out[lo - self.g0 : hi - self.g0, :] = global_field[lo:hi, :]
        return out

    def stitch(self, global_out, local_field) -> None:
        j0, j1 = self.owned_rows()
        global_out[self.g0 + j0 : self.g0 + j1, :] = local_field[j0:j1, :]

    def send_rows(self, side: str) -> slice:
        """My `ghost` border core rows that become the ghost rows of the neighbour on `side` ('top' / 'bottom')."""
        return slice(self.ghost, 2 * self.ghost) if side == "top" else slice(self.ny, self.ny + self.ghost)

    def recv_rows(self, side: str) -> slice:
        """My ghost rows on `side`."""
        return slice(0, self.ghost) if side == "top" else slice(self.ny + self.ghost, self.ny + 2 * self.ghost)


def slab_for(rank: int, world: int, NY: int, NX: int, ghost: int) -> Slab:
    if ghost < 1 or world < 1 or not (0 <= rank < world):
        raise ValueError("slab_for: bad arguments")
    if NY % world != 0 or NY // world < 2 * ghost:
        raise ValueError(f"slab_for: {NY} core rows do not divide into {world} slabs of at least {2 * ghost} rows")
    nb = {"top": None if rank == 0 else rank - 1, "bottom": None if rank == world - 1 else rank + 1, "left": None, "right": None}
    edges = EDGE_LEFT | EDGE_RIGHT | (EDGE_TOP if rank == 0 else 0) | (EDGE_BOTTOM if rank == world - 1 else 0)
    return Slab(rank, world, NY // world, NX, ghost, edges, nb)


@dataclass(frozen=True)
class DeepTile:
    """One rank's tile of a Py x Px decomposition for launches that advance several time levels: ghost cells `ghost` deep
    on all four sides. Local array (ny + 2*ghost) x (nx + 2*ghost); local [j][i] is global padded [gj0 + j][gi0 + i] (negative
    or beyond the field on a physical side: those cells are unused). Core cells are [ghost, ny+ghost) x [ghost, nx+ghost); on
    a physical side the row / column next to them is the ring."""

    rank: int
    py: int
    px: int
    cy: int
    cx: int
    ny: int
    nx: int
    ghost: int
    edges: int
    neighbours: dict = field(default_factory=dict)

    @property
    def shape(self) -> tuple[int, int]:
        return self.ny + 2 * self.ghost, self.nx + 2 * self.ghost

    @property
    def gj0(self) -> int:
        return self.cy * self.ny - (self.ghost - 1)

    @property
    def gi0(self) -> int:
        return self.cx * self.nx - (self.ghost - 1)

    def owned(self) -> tuple[int, int, int, int]:
        """[j0, j1) x [i0, i1) of the local cells this tile OWNS: its core cells plus the physical ring on boundary sides."""
        g = self.ghost
        j0 = g - 1 if self.edges & EDGE_TOP else g
        j1 = self.ny + g + 1 if self.edges & EDGE_BOTTOM else self.ny + g
        i0 = g - 1 if self.edges & EDGE_LEFT else g
        i1 = self.nx + g + 1 if self.edges & EDGE_RIGHT else self.nx + g
        return j0, j1, i0, i1

    def window(self, global_field):
        """This tile's local array cut out of a global (NY+2) x (NX+2) padded field; cells outside the field are 0."""
        import numpy as np

        rows, cols = self.shape
        NYp, NXp = global_field.shape
        out = np.zeros((rows, cols))
        jl, jh = max(self.gj0, 0), min(self.gj0 + rows, NYp)
        il, ih = max(self.gi0, 0), min(self.gi0 + cols, NXp)
        out[jl - self.gj0 : jh - self.gj0, il - self.gi0 : ih - self.gi0] = global_field[jl:jh, il:ih]
        return out

    def stitch(self, global_out, local_field) -> None:
        j0, j1, i0, i1 = self.owned()
        global_out[self.gj0 + j0 : self.gj0 + j1, self.gi0 + i0 : self.gi0 + i1] = local_field[j0:j1, i0:i1]


def deep_tile_for(rank: int, world: int, NY: int, NX: int, ghost: int, grid: Optional[tuple[int, int]] = None) -> DeepTile:
    """Geometry of rank's tile with ghost cells `ghost` deep (the Py x Px grid of tile_for)."""
    t = tile_for(rank, world, NY, NX, grid)
    if ghost < 1 or t.ny < 2 * ghost or t.nx < 2 * ghost:
        raise ValueError(f"deep_tile_for: tiles of {t.ny} x {t.nx} core cells are smaller than 2 x {ghost} ghost cells")
    return DeepTile(rank, t.py, t.px, t.cy, t.cx, t.ny, t.nx, ghost, t.edges, t.neighbours)


#: time levels one launch can advance: 1 (one-step kernel), 2, 3, 4 (tile kernels), 4, 6, 8 (walker kernel)
SUPPORTED_DEPTHS = (1, 2, 3, 4, 6, 8)


def launch_schedule(n: int, depth: int, min_depth: int = 1, depths: tuple = SUPPORTED_DEPTHS) -> list[int]:
    """Time levels per launch for n steps with at most `depth` levels per launch, out of the depths the kernels support
    (`depths`, by default SUPPORTED_DEPTHS; deep 2-D tiles: the walker's 4, 6, 8): the fewest launches, and among those
    schedules the one whose shallowest launch is deepest (4 = 2 + 2 rather than 3 + 1), deep launches first. min_depth = 2
    for slabs, which cannot advance a single level."""
    if n < 0 or depth < max(1, min_depth):
        raise ValueError("launch_schedule: bad arguments")
    allowed = [d for d in depths if min_depth <= d <= depth]
    if not allowed:
        raise ValueError(f"no supported launch depth in {min_depth}..{depth}")
    top = allowed[-1]
    bulk = max(0, (n - 4 * top) // top)  # whole launches of the deepest kind; the tail is planned exactly
    m = n - bulk * top
    # best[k] = (launches, -shallowest, first depth) for k steps, or None
    best: list = [None] * (m + 1)
    best[0] = (0, -top, 0)
    for k in range(1, m + 1):
        for d in allowed:
            if d <= k and best[k - d] is not None:
                cand = (best[k - d][0] + 1, max(best[k - d][1], -d), d)
                if best[k] is None or cand[:2] < best[k][:2]:
                    best[k] = cand
    if best[m] is None:
        raise ValueError(f"{n} step(s) cannot be covered by launches of {allowed} time levels")
    tail, k = [], m
    while k > 0:
        tail.append(best[k][2])
        k -= best[k][2]
    return [top] * bulk + sorted(tail, reverse=True)


def combine_in_rank_order(parts):
    """The Dot / float-reduce exchange step: one scalar per rank, summed left to right in rank order so the result
    does not depend on the collective's internal order."""
    total = parts[0]
    for p in parts[1:]:
        total = total + p
    return total
