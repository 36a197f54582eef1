#!/usr/bin/env python
"""python tools/mgic2hdf5.py vcPoissonFinal.3d.mgic [vcPoissonFinal.3d.hdf5]: the GRChombo checkpoint the reference writes
(Source/WriteOutput.H:127-227) from the container this build writes instead (no HDF5 library here).  Needs h5py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    from mg_ic_code_b200 import checkpoint
    src = sys.argv[1]
    hdr, _ = checkpoint.read(src, load_data=False)
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(src) or ".", hdr["filename"])
    print(checkpoint.to_hdf5(src, dst))
