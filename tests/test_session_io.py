"""CPU tests of the session file format (liloc_b200/host/session_io.cpp): the PCD and g2o files the reference writes in
save_map / save_session (src/mapOptmization.cpp:420-635) and reads in dataManager::Session (dataLoader.hpp:207-304).
PCL and GTSAM are not available here, so the files are produced byte by byte the way those libraries lay them out."""
import os

import numpy as np
import pytest


def _pcd_header(fields, sizes, types, counts, n, data):
    return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT 1\n"
            "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (" ".join(fields), " ".join(map(str, sizes)), " ".join(types), " ".join(map(str, counts)), n, n, data)).encode()


def test_pcd_binary_xyzi_as_pcl_writes_it(tmp_path):
    from liloc_b200 import host_api
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(777, 4)).astype(np.float32) * 30
    p = tmp_path / "0.pcd"
    p.write_bytes(_pcd_header(["x", "y", "z", "intensity"], [4] * 4, ["F"] * 4, [1] * 4, len(pts), "binary") + pts.tobytes())
    assert np.array_equal(host_api.load_pcd(str(p)), pts)
    # our writer produces the same bytes, and reads back
    q = tmp_path / "1.pcd"
    host_api.save_pcd(str(q), pts)
    assert q.read_bytes() == p.read_bytes()
    assert np.array_equal(host_api.load_pcd(str(q)), pts)


def test_pcd_padding_extra_fields_and_field_order(tmp_path):
    from liloc_b200 import host_api
    rng = np.random.default_rng(2)
    n = 100
    pts = rng.normal(size=(n, 4)).astype(np.float32)
    # old PCL layouts keep the struct padding as "_" fields: x y z _ intensity _(x3)  = the 32-byte in-memory PointXYZI
    rec = np.zeros((n, 8), np.float32)
    rec[:, 0:3] = pts[:, :3]; rec[:, 3] = 1.0; rec[:, 4] = pts[:, 3]
    p = tmp_path / "pad.pcd"
    p.write_bytes(_pcd_header(["x", "y", "z", "_", "intensity", "_"], [4, 4, 4, 4, 4, 4], ["F"] * 6, [1, 1, 1, 1, 1, 3], n, "binary") + rec.tobytes())
    assert np.array_equal(host_api.load_pcd(str(p)), pts)
    # other field order + an integer ring field + a double time field (Velodyne-style clouds)
    dt = np.dtype([("intensity", "<f4"), ("ring", "<u2"), ("z", "<f4"), ("time", "<f8"), ("y", "<f4"), ("x", "<f4")])
    r2 = np.zeros(n, dt)
    r2["x"], r2["y"], r2["z"], r2["intensity"], r2["ring"], r2["time"] = pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3], 7, 1.5
    p2 = tmp_path / "order.pcd"
    p2.write_bytes(_pcd_header(["intensity", "ring", "z", "time", "y", "x"], [4, 2, 4, 8, 4, 4], ["F", "U", "F", "F", "F", "F"], [1] * 6, n, "binary") + r2.tobytes())
    assert np.array_equal(host_api.load_pcd(str(p2)), pts)
    # xyz only -> intensity 0
    p3 = tmp_path / "xyz.pcd"
    p3.write_bytes(_pcd_header(["x", "y", "z"], [4] * 3, ["F"] * 3, [1] * 3, n, "binary") + np.ascontiguousarray(pts[:, :3]).tobytes())
    got = host_api.load_pcd(str(p3))
    assert np.array_equal(got[:, :3], pts[:, :3]) and not got[:, 3].any()


def test_pcd_ascii_and_errors(tmp_path):
    from liloc_b200 import host_api
    pts = np.array([[1.5, -2.25, 3.0, 10.0], [0.0, 0.125, -7.5, 255.0]], np.float32)
    p = tmp_path / "a.pcd"
    p.write_bytes(_pcd_header(["x", "y", "z", "intensity"], [4] * 4, ["F"] * 4, [1] * 4, 2, "ascii") + b"1.5 -2.25 3 10\n0 0.125 -7.5 255\n")
    assert np.array_equal(host_api.load_pcd(str(p)), pts)
    with pytest.raises(IOError):
        host_api.load_pcd(str(tmp_path / "missing.pcd"))
    bad = tmp_path / "short.pcd"
    bad.write_bytes(_pcd_header(["x", "y", "z", "intensity"], [4] * 4, ["F"] * 4, [1] * 4, 10, "binary") + b"\0" * 40)
    with pytest.raises(IOError, match="truncated"):
        host_api.load_pcd(str(bad))
    comp = tmp_path / "comp.pcd"
    comp.write_bytes(_pcd_header(["x", "y", "z", "intensity"], [4] * 4, ["F"] * 4, [1] * 4, 1, "binary_zstd") + b"\0" * 16)
    with pytest.raises(IOError, match="not supported"):
        host_api.load_pcd(str(comp))


def test_g2o_vertices_round_trip(tmp_path):
    from liloc_b200 import host_api
    # a file in the reference's own format (README example line, dataLoader.hpp:60) with edges mixed in and ids out of order
    txt = ("VERTEX_SE3:QUAT 1 1.000000 2.000000 0.500000 0.000000 0.000000 0.382683 0.923880\n"
           "VERTEX_SE3:QUAT 0 0.000000 0.000000 0.000000 0.000000 0.000000 0.000000 1.000000\n"
           "EDGE_SE3:QUAT 0 1 1.000000 2.000000 0.500000 0.000000 0.000000 0.382683 0.923880\n"
           "VERTEX_SE3:QUAT 99 -61.332581 -9.253125 0.131973 -0.004256 -0.005810 -0.625732 0.780005\n")
    p = tmp_path / "g.g2o"
    p.write_text(txt)
    kp = host_api.load_g2o(str(p))
    assert kp.shape == (3, 7) and list(kp[:, 6]) == [0, 1, 99]                    # ascending node id, id in `intensity`
    assert np.allclose(kp[1, :6], [0, 0, np.pi / 4, 1, 2, 0.5], atol=1e-5)
    q = np.array([-0.004256, -0.005810, -0.625732, 0.780005]); q /= np.linalg.norm(q)
    yaw = np.arctan2(2 * (q[0] * q[1] + q[3] * q[2]), 1 - 2 * (q[1] ** 2 + q[2] ** 2))
    assert abs(kp[2, 2] - yaw) < 1e-6 and np.allclose(kp[2, 3:6], [-61.332581, -9.253125, 0.131973], atol=1e-5)
    # writer -> reader: poses survive the 6-decimal text format
    rng = np.random.default_rng(3)
    poses = np.c_[rng.uniform(-0.3, 0.3, (20, 2)), rng.uniform(-3, 3, 20), rng.uniform(-50, 50, (20, 3)), np.zeros(20)]
    out = tmp_path / "w.g2o"
    host_api.save_g2o(str(out), poses)
    lines = out.read_text().splitlines()
    assert sum(l.startswith("VERTEX_SE3:QUAT") for l in lines) == 20 and sum(l.startswith("EDGE_SE3:QUAT") for l in lines) == 19
    back = host_api.load_g2o(str(out))
    assert np.abs(back[:, :3] - poses[:, :3]).max() < 5e-6 and np.abs(back[:, 3:6] - poses[:, 3:6]).max() < 5e-6   # float32 + 6 decimals


def _lzf_compress(data: bytes) -> bytes:
    """A small LZF encoder (liblzf stream format): greedy matches of 3..264 bytes within the last 8192 bytes, found through
    a dictionary of 3-byte prefixes; everything else as literal runs of at most 32 bytes."""
    out, lit, last, i, n = bytearray(), bytearray(), {}, 0, len(data)

    def flush():
        for a in range(0, len(lit), 32):
            run = lit[a:a + 32]
            out.append(len(run) - 1); out.extend(run)
        lit.clear()
    while i < n:
        key = data[i:i + 3]
        j = last.get(key, -1) if len(key) == 3 else -1
        if len(key) == 3:
            last[key] = i
        if j >= 0 and i - j <= 8192:
            ln = 3
            while i + ln < n and ln < 264 and data[j + ln] == data[i + ln]:
                ln += 1
            flush()
            dist, l2 = i - j - 1, ln - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8)); out.append(l2 - 7)
            out.append(dist & 0xFF)
            i += ln
        else:
            lit.append(data[i]); i += 1
    flush()
    return bytes(out)


def test_transformations_pcd_layout(tmp_path):
    """transformations.pcd (src/mapOptmization.cpp:438, :564): PointXYZIRPYT (include/utility.h:87-103) the way
    pcl::io::savePCDFileBinary lays a registered point struct out -- raw 48-byte records, padding declared as `_` fields."""
    from liloc_b200 import host_api
    rng = np.random.default_rng(5)
    kp = np.c_[rng.uniform(-3, 3, (7, 3)), rng.uniform(-50, 50, (7, 3)), 1.7e9 + np.arange(7) * 0.1]
    path = str(tmp_path / "transformations.pcd")
    assert host_api.lib().liloc_host_save_poses_pcd(path.encode(), kp.ctypes.data_as(host_api.C.POINTER(host_api.C.c_double)), len(kp)) == 0
    raw = open(path, "rb").read()
    head, body = raw.split(b"DATA binary\n", 1)
    lines = dict(l.split(" ", 1) for l in head.decode().splitlines() if not l.startswith("#"))
    assert lines["FIELDS"] == "x y z _ intensity roll pitch yaw time _" and lines["SIZE"] == "4 4 4 1 4 4 4 4 8 1"
    assert lines["TYPE"] == "F F F U F F F F F U" and lines["COUNT"] == "1 1 1 4 1 1 1 1 1 8"
    assert lines["POINTS"] == "7" and lines["WIDTH"] == "7" and lines["HEIGHT"] == "1"
    assert sum(int(a) * int(b) for a, b in zip(lines["SIZE"].split(), lines["COUNT"].split())) == 48 and len(body) == 48 * 7
    rec = np.frombuffer(body, np.dtype([("xyz", "<f4", 3), ("one", "<f4"), ("irpy", "<f4", 4), ("time", "<f8"), ("pad", "V8")]))
    assert np.array_equal(rec["xyz"], kp[:, 3:6].astype(np.float32)) and np.array_equal(rec["irpy"][:, 1:], kp[:, :3].astype(np.float32))
    assert np.array_equal(rec["time"], kp[:, 6]) and np.array_equal(rec["irpy"][:, 0], np.arange(7, dtype=np.float32))
    # the PCD reader of this repo takes it as a cloud of the key positions (x y z intensity; the rest skipped)
    pts = host_api.load_pcd(path)
    assert np.array_equal(pts[:, :3], kp[:, 3:6].astype(np.float32)) and np.array_equal(pts[:, 3], np.arange(7, dtype=np.float32))


def test_pcd_binary_compressed(tmp_path):
    """pcl::io::savePCDFileBinaryCompressed: sizes + LZF stream over the field-major block."""
    import struct
    from liloc_b200 import host_api
    rng = np.random.default_rng(3)
    n = 1500
    pts = np.round(rng.normal(size=(n, 4)) * 20, 1).astype(np.float32)      # few distinct values: real back references
    pts[200:900, 2] = 1.25; pts[:, 3] = 7.0                                    # long runs (overlapping copies)
    ring = rng.integers(0, 16, n).astype(np.uint16)
    block = pts[:, 0].tobytes() + pts[:, 1].tobytes() + pts[:, 2].tobytes() + pts[:, 3].tobytes() + ring.tobytes()
    comp = _lzf_compress(block)
    assert len(comp) < 0.8 * len(block)
    p = tmp_path / "c.pcd"
    p.write_bytes(_pcd_header(["x", "y", "z", "intensity", "ring"], [4, 4, 4, 4, 2], ["F", "F", "F", "F", "U"], [1] * 5, n, "binary_compressed")
                  + struct.pack("<II", len(comp), len(block)) + comp)
    assert np.array_equal(host_api.load_pcd(str(p)), pts)
    # literal-only stream (what an encoder emits for incompressible data)
    lit = b"".join(bytes([len(block[a:a + 32]) - 1]) + block[a:a + 32] for a in range(0, len(block), 32))
    p.write_bytes(_pcd_header(["x", "y", "z", "intensity", "ring"], [4, 4, 4, 4, 2], ["F", "F", "F", "F", "U"], [1] * 5, n, "binary_compressed")
                  + struct.pack("<II", len(lit), len(block)) + lit)
    assert np.array_equal(host_api.load_pcd(str(p)), pts)
    # corrupt streams are rejected, not read out of bounds
    for bad in (comp[:-3], b"\xe0\xff\xff" + comp, comp + b"\x05abc"):
        p.write_bytes(_pcd_header(["x", "y", "z", "intensity", "ring"], [4, 4, 4, 4, 2], ["F", "F", "F", "F", "U"], [1] * 5, n, "binary_compressed")
                      + struct.pack("<II", len(bad), len(block)) + bad)
        with pytest.raises(Exception):
            host_api.load_pcd(str(p))
