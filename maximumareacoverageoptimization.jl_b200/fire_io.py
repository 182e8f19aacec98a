"""Fire-point table I/O: the data format on the input side of the coverage path.

The reference hands fire points from its cellular automaton (src/DynamicArea.jl:100-108) to the
cell store (src/CellFunctions.jl:20-79) through an .xlsx sheet: one row per timestep, each row a
flat run of 5-tuples `[x, y, area, weight, covered]`.  This module reads that sheet without any
spreadsheet dependency (an .xlsx is a zip of XML) and reads/writes the same rows as a compact
.npz (used for the committed test fixture).  All citations relative to /root/reference/.
"""
from __future__ import annotations

import re
import zipfile

import numpy as np

_ROW = re.compile(r"<row [^>]*>(.*?)</row>", re.S)
_CELL = re.compile(r'<c r="([A-Z]+)(\d+)"([^>]*?)(?:/>|>(.*?)</c>)', re.S)
_VAL = re.compile(r"<v>(.*?)</v>", re.S)


def _col_index(letters: str) -> int:
    n = 0
    for ch in letters:
        n = n * 26 + (ord(ch) - 64)
    return n


def load_fire_rows_xlsx(path: str, sheet: str = "xl/worksheets/sheet1.xml") -> list[np.ndarray]:
    """Rows of the sheet as (n_k x 5) float64 arrays; missing cells are dropped like the reference's
    `filter!(!ismissing, ...)` (src/CellFunctions.jl:36)."""
    with zipfile.ZipFile(path) as z:
        xml = z.read(sheet).decode("utf-8")
    rows = []
    for m in _ROW.finditer(xml):
        cells = []
        for c in _CELL.finditer(m.group(1)):
            body = c.group(4)
            if body is None:
                continue
            v = _VAL.search(body)
            if v is None:
                continue
            if 't="s"' in c.group(3) or 't="str"' in c.group(3):
                raise ValueError("string cell in a fire-point sheet")
            cells.append((_col_index(c.group(1)), float(v.group(1))))
        cells.sort()
        vals = np.array([v for _, v in cells], dtype=np.float64)
        if vals.size % 5:
            raise ValueError("row length is not a multiple of 5")
        rows.append(vals.reshape(-1, 5))
    return rows


def save_fire_rows_npz(path: str, rows) -> None:
    lens = np.array([len(r) for r in rows], dtype=np.int64)
    flat = np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 5) for r in rows], axis=0) if len(rows) else np.zeros((0, 5))
    np.savez_compressed(path, lens=lens, flat=flat)


def load_fire_rows_npz(path: str) -> list[np.ndarray]:
    with np.load(path) as f:
        lens, flat = f["lens"], f["flat"]
    out, k = [], 0
    for n in lens:
        out.append(flat[k:k + n].copy())
        k += n
    return out


def load_fire_rows(path: str) -> list[np.ndarray]:
    return load_fire_rows_xlsx(path) if path.endswith(".xlsx") else load_fire_rows_npz(path)


def rows_from_points(steps) -> list[np.ndarray]:
    """DynamicArea's export (src/DynamicArea.jl:100-108): one row per step."""
    return [np.asarray(s, dtype=np.float64).reshape(-1, 5) for s in steps]
