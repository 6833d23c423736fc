"""Readers for the reference's three input formats (unchanged, so existing files keep working):
  csv_config      key = value, CSVconfig.h:13-98
  parameter file  name = init[, step[, lower, upper]], Parameters.h:50-85, :140-172
  input csv       read_data, moma_input.h:401-527 (+ build_cell_genealogy :125-151)
Host-side input preparation; nothing numerical happens here except log() of the length column, which goes
through the C library's log like the reference's (math.log, not numpy's vectorised log).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from .forest import LineageData, PARAM_NAMES


@dataclass
class CSVConfig:
    time_col: str = "time"
    rescale_time: float = 1.0
    length_col: str = "length"
    length_islog: bool = False
    fp_col: str = "gfp"
    fp_auto: float = 0.0
    delm: str = ","
    segment_col: str = ""
    filter_col: str = ""
    cell_tags: list = field(default_factory=lambda: ["cell_id"])
    parent_tags: list = field(default_factory=lambda: ["parent_id"])


def _string2bool(s):
    """utils.h:20-29: exactly these spellings, anything else is an input error"""
    if s in ("True", "true", "TRUE", "1"):
        return True
    if s in ("False", "false", "FALSE", "0"):
        return False
    raise ValueError(f"(string2bool) no valid bool conversion of {s!r}")


def read_csv_config(path):
    cfg = CSVConfig()
    if path is None:
        return cfg
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if not line or line[0] == "#":
                continue
            parts = line.split("=")
            if len(parts) < 2:
                continue
            key, val = parts[0].strip(), parts[1].strip()
            if key in ("time_col", "length_col", "fp_col", "delm", "segment_col", "filter_col"):
                setattr(cfg, key, val)
            elif key in ("rescale_time", "fp_auto"):
                setattr(cfg, key, float(val))
            elif key == "length_islog":
                cfg.length_islog = _string2bool(val)
            elif key in ("cell_tags", "parent_tags"):
                setattr(cfg, key, [v.strip() for v in val.split(",")])
    return cfg


@dataclass
class Parameter:
    name: str
    init: float = 0.0
    step: float = 0.0
    lower: float = 0.0
    upper: float = 0.0
    kind: str = "unset"   # fixed | free | bound


def read_parameter_file(path):
    """list of 11 Parameter in the fixed order (Parameters.h:175)"""
    params = {n: Parameter(n) for n in PARAM_NAMES}
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if not line or line[0] == "#":
                continue
            parts = line.split("=")
            name = parts[0].strip()
            if name not in params or len(parts) < 2:
                continue
            vals = [float(v.strip()) for v in parts[1].split(",")]
            if any(math.isnan(v) for v in vals):
                raise ValueError(f"parameter {name}: nan")
            p = params[name]
            if len(vals) == 4:
                p.init, p.step, p.lower, p.upper, p.kind = vals[0], vals[1], vals[2], vals[3], "bound"
            elif len(vals) == 1:
                p.init, p.kind = vals[0], "fixed"
            elif len(vals) == 2:
                p.init, p.step, p.kind = vals[0], vals[1], "free"
            else:
                raise ValueError(f"parameter {name}: invalid number of arguments")
    missing = [n for n in PARAM_NAMES if params[n].kind == "unset"]
    if missing:
        raise ValueError("parameters not set: " + ", ".join(missing))
    return [params[n] for n in PARAM_NAMES]


def _remove_last_decimal(s):
    """moma_input.h:327-347: only a purely numeric tag whose part after the last '.' is all zeros loses its decimals
    ("7.0" -> "7"); "2.1", "20150624.0.1.5" and "a.b" stay as they are"""
    if any(not (ch.isdigit() and ch.isascii()) and ch != "." for ch in s):
        return s
    if any(ch != "0" for ch in s.split(".")[-1]):
        return s
    return str(int(float(s)))   # std::to_string(std::stoi(str)): the leading integer


def read_data(path, cfg: CSVConfig = None, noise_model="scaled", division_model="binomial"):
    """returns (LineageData, cell_ids).  Rows are grouped into a new cell whenever the composed id changes."""
    cfg = cfg or CSVConfig()
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split(cfg.delm)
        idx = {}
        for i, h in enumerate(header):   # get_header_indices (moma_input.h:366-381): trimmed tags, the first occurrence wins
            idx.setdefault(h.strip(" \t\n\v\f\r"), i)
        for col in [cfg.time_col, cfg.length_col, cfg.fp_col] + cfg.cell_tags + cfg.parent_tags + \
                ([cfg.segment_col] if cfg.segment_col else []) + ([cfg.filter_col] if cfg.filter_col else []):
            if col not in idx:
                raise ValueError(f"(read_data) {col} is not a column in input file")
        cell_ids, parent_ids, offsets = [], [], []
        time, x, g, seg = [], [], [], []
        last = None
        for line in fh:
            parts = line.rstrip("\n").split(cfg.delm)
            if cfg.filter_col and not _string2bool(parts[idx[cfg.filter_col]]):
                continue
            cid = ".".join(_remove_last_decimal(parts[idx[t]]) for t in cfg.cell_tags)
            if cid != last:
                cell_ids.append(cid)
                parent_ids.append(".".join(_remove_last_decimal(parts[idx[t]]) for t in cfg.parent_tags))
                offsets.append(len(time))
                last = cid
            vals = (float(parts[idx[cfg.time_col]]), float(parts[idx[cfg.length_col]]), float(parts[idx[cfg.fp_col]]))
            if any(math.isnan(v) for v in vals):
                raise ValueError("nan in input")
            time.append(vals[0] / cfg.rescale_time)
            x.append(vals[1] if cfg.length_islog else math.log(vals[1]))
            g.append(vals[2])
            seg.append(int(parts[idx[cfg.segment_col]]) if cfg.segment_col else 0)
    offsets.append(len(time))
    # genealogy with a hash map instead of the reference's O(N^2) scan; same result for unique ids
    first_index = {}
    for i, cid in enumerate(cell_ids):
        if cid in first_index:
            raise ValueError(f"cell id {cid} appears in two separate blocks of rows")
        first_index[cid] = i
    parent = np.array([first_index.get(p, -1) for p in parent_ids], dtype=np.int32)
    data = LineageData(cell_offset=np.array(offsets), parent=parent, time=np.array(time), log_length=np.array(x),
                       fp=np.array(g), segment=np.array(seg, dtype=np.int32), noise_model=noise_model,
                       division_model=division_model, fp_auto=cfg.fp_auto)
    return data, cell_ids


# ---- binary forest file (layout in host/ggp_data.hpp::read_binary_forest) -------------------------------
_MAGIC = b"GGPFORE1"


def write_forest_binary(path, data, cell_ids=None, parent_ids=None):
    """LineageData -> the binary forest file `gfp_gaussian -i` reads instead of a csv.  cell_ids / parent_ids: the id strings the
    output files print (default: the cell's number from 1, the parent's number or 0 for a root)."""
    import struct
    n_cells, n_ctp = int(data.n_cells), int(data.n_ctp)
    if cell_ids is None:
        cell_ids = [str(c + 1) for c in range(n_cells)]
    if parent_ids is None:
        parent_ids = [str(int(p) + 1) if p >= 0 else "0" for p in data.parent]
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(struct.pack("<qq", n_cells, n_ctp))
        f.write(np.ascontiguousarray(data.cell_offset, dtype="<i8").tobytes())
        for a in (data.time, data.log_length, data.fp):
            f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(data.segment, dtype="<i4").tobytes())
        for c, p in zip(cell_ids, parent_ids):
            for s in (c, p):
                b = str(s).encode()
                f.write(struct.pack("<I", len(b)))
                f.write(b)


def read_forest_binary(path, noise_model="scaled", division_model="binomial"):
    """binary forest file -> (LineageData, cell_ids, parent_ids); parents are resolved like the csv reader does (a cell's parent is
    the cell whose id equals its parent id)"""
    import struct
    from .forest import LineageData
    with open(path, "rb") as f:
        if f.read(8) != _MAGIC:
            raise ValueError("not a binary forest file: " + str(path))
        n_cells, n_ctp = struct.unpack("<qq", f.read(16))
        off = np.frombuffer(f.read(8 * (n_cells + 1)), dtype="<i8").astype(np.int64)
        time, x, g = (np.frombuffer(f.read(8 * n_ctp), dtype="<f8").astype(np.float64) for _ in range(3))
        seg = np.frombuffer(f.read(4 * n_ctp), dtype="<i4").astype(np.int32)
        ids = []
        for _ in range(2 * n_cells):
            (n,) = struct.unpack("<I", f.read(4))
            ids.append(f.read(n).decode())
    cell_ids, parent_ids = ids[0::2], ids[1::2]
    index = {c: i for i, c in enumerate(cell_ids)}
    parent = np.array([index.get(p, -1) for p in parent_ids], dtype=np.int64)
    data = LineageData(cell_offset=off, parent=parent, time=time, log_length=x, fp=g, segment=seg, noise_model=noise_model,
                       division_model=division_model)
    return data, cell_ids, parent_ids
