"""Grid3D / .xyz format (Grid3D.h, GridData.h): product mirror vs a file written by the REFERENCE's own Grid3D::Save
(tests/golden/grid_ref_saved.xyz) and what the reference's Load makes of it (grid_ref_loaded.json)."""
import json
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = os.path.join(HERE, "golden", "grid_ref_saved.xyz")
REF = json.load(open(os.path.join(HERE, "golden", "grid_ref_loaded.json")))


def test_load_matches_reference_loader(crv):
    g = crv.Grid3D.Load(REF_FILE)
    assert list(g.GetVoxelCount()) == REF["counts"] and list(g.GetVoxelSize()) == REF["voxel_size"]
    assert g.data.tolist() == REF["loaded"]
    assert g.GetMajorant() == REF["majorant"]
    assert g.GetSize() == 60 and g.GetByteSize() == 240 and g.GetElementSize() == 4


def test_save_is_byte_identical_to_reference_save(crv, tmp_path):
    g = crv.Grid3D(5, 4, 3, 0.01, 0.02, 0.04)
    g.data[:] = np.array(REF["saved_input"], np.float32)
    out = tmp_path / "mine.xyz"
    g.Save(str(out))
    assert out.read_bytes() == open(REF_FILE, "rb").read()


def test_file_layout_and_zfastest_roundtrip(crv, tmp_path):
    """header = 3 x u32 + 3 x f64, payload z-fastest (GridData.h:12); SaveZFastest round-trips through Load while the
    reference's Save (x-fastest dump) does not -- the asymmetry is kept (Grid3D.h:61-85 vs :167-178)"""
    nx, ny, nz = 4, 3, 5
    g = crv.Grid3D(nx, ny, nz, 0.5, 0.25, 0.125)
    g.data[:] = np.arange(nx * ny * nz, dtype=np.float32)
    p = tmp_path / "z.xyz"
    g.SaveZFastest(str(p))
    raw = p.read_bytes()
    assert struct.unpack("<3I3d", raw[:36]) == (nx, ny, nz, 0.5, 0.25, 0.125)
    payload = np.frombuffer(raw[36:], np.float32)
    assert payload[1] == g(0, 0, 1) and payload[nz] == g(0, 1, 0) and payload[nz * ny] == g(1, 0, 0)
    back = crv.Grid3D.Load(str(p))
    assert np.array_equal(back.data, g.data)
    q = tmp_path / "x.xyz"
    g.Save(str(q))
    assert not np.array_equal(crv.Grid3D.Load(str(q)).data, g.data)


def test_error_behaviour(crv, tmp_path):
    assert crv.Grid3D.Load(str(tmp_path / "missing.xyz")) is None         # nullptr (Grid3D.h:54-58)
    g = crv.Grid3D(2, 2, 2)
    with pytest.raises(IndexError):                                       # std::out_of_range (Grid3D.h:119)
        g(2, 0, 0)
    assert g.GetMajorant() == 0.0
    g.data[:] = -1.0
    assert g.GetMajorant() == 0.0                                          # the scan starts from T() = 0
