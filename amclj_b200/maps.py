"""Occupancy-grid ingest: the map the hot path reads.

The reference receives its map as a ``nav_msgs/OccupancyGrid`` message produced by the ROS
``map_server`` node from ``launch/lgfloor.pgm`` + ``launch/lgfloor.yaml``
(reference launch/stage_pioneer.launch:10; message shape nav_msgs.clj:25-38: ``data`` is a
``height x width`` row-major matrix of signed bytes, -1 unknown / 0 free / 100 occupied,
``get-cell map x y = (sel data y x)``, map.clj:19-24).  ``map_server`` itself is external
to the reference; its published trinary conversion is restated in
:func:`occupancy_from_image`.

This module is host-side ingest only (not on the timed path).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_DATA = Path(__file__).resolve().parent / "data"


@dataclass
class MapMetaData:
    """nav_msgs/MapMetaData as amclj uses it (nav_msgs.clj:8-23).

    ``resolution`` is a ROS float32; ``origin`` is carried but ignored by
    ``world->map`` (map.clj:36-46), exactly as in the reference.
    """

    resolution: np.float32
    width: int
    height: int
    origin: tuple = (0.0, 0.0, 0.0)


@dataclass
class OccupancyGrid:
    """nav_msgs/OccupancyGrid (nav_msgs.clj:25-38): ``data[y, x]`` int8."""

    info: MapMetaData
    data: np.ndarray
    header: dict = field(default_factory=lambda: {"frameId": "map"})

    def get_cell(self, x: int, y: int) -> int:
        """map.clj:19-24 with the documented bounds contract (x >= W, y >= H => -1)."""
        if x < 0 or y < 0 or x >= self.info.width or y >= self.info.height:
            return -1
        return int(self.data[y, x])

    def inhabitable(self, x: int, y: int) -> bool:
        """map.clj:48-55."""
        return self.get_cell(x, y) == 0


def read_pgm(path) -> np.ndarray:
    """Binary (P5) PGM -> uint8/uint16 image, row 0 = top of the picture."""
    raw = Path(path).read_bytes()
    m = re.match(rb"P5\s+(?:#[^\n]*\n\s*)*(\d+)\s+(\d+)\s+(\d+)\s", raw)
    if m is None:
        raise ValueError(f"{path}: not a binary PGM (P5)")
    w, h, maxval = (int(g) for g in m.groups())
    dt = np.uint8 if maxval < 256 else np.dtype(">u2")
    img = np.frombuffer(raw, dtype=dt, count=w * h, offset=m.end()).reshape(h, w)
    return img


def write_pgm(path, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    Path(path).write_bytes(b"P5\n%d %d\n255\n" % (w, h) + img.tobytes())


def read_map_yaml(path) -> dict:
    """The six keys of a map_server YAML (launch/lgfloor.yaml:1-6)."""
    import yaml

    meta = yaml.safe_load(Path(path).read_text())
    for key in ("image", "resolution", "origin", "negate", "occupied_thresh", "free_thresh"):
        if key not in meta:
            raise ValueError(f"{path}: missing key {key!r}")
    return meta


def occupancy_from_image(img, resolution, negate=0, occupied_thresh=0.65, free_thresh=0.196,
                         origin=(0.0, 0.0, 0.0), maxval=255) -> OccupancyGrid:
    """map_server's trinary conversion: ``occ = (255 - v)/255`` (``v/255`` when negated);
    ``occ > occupied_thresh`` -> 100, ``occ < free_thresh`` -> 0, else -1; image rows are
    flipped so grid row 0 is the picture's bottom row."""
    v = np.asarray(img).astype(np.float64)
    occ = v / float(maxval) if negate else (float(maxval) - v) / float(maxval)
    grid = np.full(v.shape, -1, dtype=np.int8)
    grid[occ > occupied_thresh] = 100
    grid[occ < free_thresh] = 0
    grid = np.ascontiguousarray(grid[::-1])
    h, w = grid.shape
    return OccupancyGrid(MapMetaData(np.float32(resolution), w, h, tuple(origin)), grid)


def load_map(yaml_path) -> OccupancyGrid:
    """Load ``<name>.yaml`` + the PGM it names, the way ``map_server <yaml>`` would."""
    yaml_path = Path(yaml_path)
    meta = read_map_yaml(yaml_path)
    img = read_pgm(yaml_path.parent / meta["image"])
    return occupancy_from_image(img, meta["resolution"], meta["negate"],
                                meta["occupied_thresh"], meta["free_thresh"], meta["origin"])


def lgfloor() -> OccupancyGrid:
    """The reference's map fixture (launch/lgfloor.pgm/.yaml, 662 x 639 @ 0.05 m),
    re-encoded by tools/import_lgfloor.py so it travels to machines without the reference."""
    z = np.load(_DATA / "lgfloor_map.npz")
    return occupancy_from_image(z["image"], float(z["resolution"]), int(z["negate"]),
                                float(z["occupied_thresh"]), float(z["free_thresh"]),
                                tuple(z["origin"]), int(z["maxval"]))


def synthetic_map(width=8192, height=8192, resolution=0.05, seed=5, target_occupied=0.03):
    """BASELINE config C5: 1-cell border wall plus seeded axis-aligned wall segments and
    room outlines until ~``target_occupied`` of the cells are occupied; no unknown cells."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.zeros((height, width), dtype=np.int8)
    g[0, :] = g[-1, :] = 100
    g[:, 0] = g[:, -1] = 100
    target = int(target_occupied * width * height)
    occupied = int((g != 0).sum())
    while occupied < target:
        kind = rng.integers(0, 3)
        t = int(rng.integers(1, 4))  # wall thickness in cells
        if kind == 0:  # horizontal segment
            y = int(rng.integers(1, height - t))
            x0 = int(rng.integers(1, width - 2))
            ln = int(rng.integers(40, 600))
            blk = g[y:y + t, x0:min(x0 + ln, width - 1)]
        elif kind == 1:  # vertical segment
            x = int(rng.integers(1, width - t))
            y0 = int(rng.integers(1, height - 2))
            ln = int(rng.integers(40, 600))
            blk = g[y0:min(y0 + ln, height - 1), x:x + t]
        else:  # room outline with a door gap
            rw, rh = int(rng.integers(60, 400)), int(rng.integers(60, 400))
            x0 = int(rng.integers(1, max(2, width - rw - 1)))
            y0 = int(rng.integers(1, max(2, height - rh - 1)))
            before = int((g[y0:y0 + rh, x0:x0 + rw] != 0).sum())
            g[y0, x0:x0 + rw] = 100
            g[min(y0 + rh - 1, height - 1), x0:x0 + rw] = 100
            g[y0:y0 + rh, x0] = 100
            g[y0:y0 + rh, min(x0 + rw - 1, width - 1)] = 100
            door = int(rng.integers(x0 + 5, x0 + rw - 25)) if rw > 40 else x0 + 5
            g[y0, door:door + 20] = 0
            occupied += int((g[y0:y0 + rh, x0:x0 + rw] != 0).sum()) - before
            continue
        before = int((blk != 0).sum())
        blk[...] = 100
        occupied += blk.size - before
    g[0, :] = g[-1, :] = 100
    g[:, 0] = g[:, -1] = 100
    return OccupancyGrid(MapMetaData(np.float32(resolution), width, height), g)
