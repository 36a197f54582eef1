"""Reader of the MGICCHK1 container mgic_hier_write_checkpoint writes (the GRChombo checkpoint of Source/WriteOutput.H:127-227
without HDF5) and its conversion to Chombo's HDF5 layout where h5py exists (tools/mgic2hdf5.py is the command line).

Container: b"MGICCHK1", uint64 header length, JSON header, zero padding to a multiple of 8 bytes, float64 data.  Data order =
Chombo's write(handle, LevelData, "data"): level after level, box after box, per box component after component, x fastest
over the box grown by the ghost vector."""
import json
import struct

import numpy as np


def read(path, load_data=True):
    """-> (header dict, [per level: list of arrays [32, nz+6, ny+6, nx+6] per box] or None)"""
    with open(path, "rb") as f:
        if f.read(8) != b"MGICCHK1":
            raise ValueError(f"{path} is not an MGICCHK1 container")
        (hl,) = struct.unpack("<Q", f.read(8))
        hdr = json.loads(f.read(hl).decode())
        f.read((8 - hl % 8) % 8)
        start = f.tell()
        if not load_data:
            return hdr, None
        g = hdr["ghost"]
        ncomp = hdr["root"]["ints"]["num_components"]
        levels = []
        for lv in hdr["levels"]:
            boxes = []
            for b, off in zip(lv["boxes"], lv["offsets"]):
                shape = (ncomp, b[5] - b[2] + 1 + 2 * g[2], b[4] - b[1] + 1 + 2 * g[1], b[3] - b[0] + 1 + 2 * g[0])
                f.seek(start + 8 * (lv["data_start_double"] + off))
                boxes.append(np.fromfile(f, dtype="<f8", count=int(np.prod(shape))).reshape(shape))
            levels.append(boxes)
    return hdr, levels


def to_hdf5(path, out):
    """Write Chombo's checkpoint layout (what HDF5HeaderData::writeToFile, write(handle, DisjointBoxLayout) and
    write(handle, LevelData, "data") produce) from the container.  Needs h5py."""
    import h5py   # not in the build image; present wherever GRChombo's tools are
    hdr, levels = read(path)
    g = hdr["ghost"]
    with h5py.File(out, "w") as h:
        for k, v in hdr["root"]["ints"].items():
            h.attrs[k] = np.int32(v)
        for k, v in hdr["root"]["reals"].items():
            h.attrs[k] = np.float64(v)
        for k, v in hdr["root"]["strings"].items():
            h.attrs[k] = np.bytes_(v)
        cg = h.create_group("Chombo_global")
        cg.attrs["SpaceDim"] = np.int32(3)
        cg.attrs["testReal"] = np.float64(0.0)
        box_t = np.dtype([("lo_i", "<i4"), ("lo_j", "<i4"), ("lo_k", "<i4"), ("hi_i", "<i4"), ("hi_j", "<i4"), ("hi_k", "<i4")])
        iv_t = np.dtype([("intvecti", "<i4"), ("intvectj", "<i4"), ("intvectk", "<i4")])
        for lv, data in zip(hdr["levels"], levels):
            grp = h.create_group(lv["group"])
            for k, v in lv["ints"].items():
                grp.attrs[k] = np.int32(v)
            for k, v in lv["reals"].items():
                grp.attrs[k] = np.float64(v)
            grp.attrs["prob_domain"] = np.array(tuple(lv["prob_domain"]), dtype=box_t)
            grp.create_dataset("boxes", data=np.array([tuple(b) for b in lv["boxes"]], dtype=box_t))
            grp.create_dataset("Processors", data=np.zeros(len(lv["boxes"]), dtype="<i4"))
            flat = np.concatenate([d.ravel() for d in data]) if data else np.zeros(0)
            grp.create_dataset("data:datatype=0", data=flat)
            grp.create_dataset("data:offsets=0", data=np.array(lv["offsets"], dtype="<i8"))
            da = grp.create_group("data_attributes")
            da.attrs["comps"] = np.int32(hdr["root"]["ints"]["num_components"])
            da.attrs["ghost"] = np.array(tuple(g), dtype=iv_t)
            da.attrs["outputGhost"] = np.array(tuple(g), dtype=iv_t)
            da.attrs["objectType"] = np.bytes_("FArrayBox")
    return out
