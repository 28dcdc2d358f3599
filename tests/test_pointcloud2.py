"""sensor_msgs/PointCloud2 replay: the ROS 1 wire format is written HERE with struct.pack (an independent
implementation of the message layout) and read by include/pclomp_b200/pointcloud2.hpp; the GPU test replays a dump of
synthetic scans through apps/replay_b200 (the mapping-node loop) against the oracle's restatement of that loop."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from util import ROOT

INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "toyslam_b200", "lib")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
DT = {1: "b", 2: "B", 3: "h", 4: "H", 5: "i", 6: "I", 7: "f", 8: "d"}


def ros_string(s):
    b = s.encode()
    return struct.pack("<I", len(b)) + b


def pointcloud2_bytes(xyz, seq=0, stamp=(0, 0), frame_id="velodyne", layout="xyz16", big_endian=False, is_dense=True, height=1):
    """Serialise an (n,3) array as a sensor_msgs/PointCloud2 message.  layouts:
    xyz16   x y z FLOAT32 at 0/4/8, point_step 16 (pcl::PointXYZ)
    xyzir32 x y z at 0/4/8, intensity FLOAT32 at 16, ring UINT16 at 20, point_step 32 (Velodyne driver)
    f64     x y z FLOAT64 at 8/16/24 after a leading UINT32 'id', point_step 32"""
    n = len(xyz)
    e = ">" if big_endian else "<"
    if layout == "xyz16":
        fields, step = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7)], 16
    elif layout == "xyzir32":
        fields, step = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("ring", 20, 4)], 32
    else:
        fields, step = [("id", 0, 6), ("x", 8, 8), ("y", 16, 8), ("z", 24, 8)], 32
    width = n // height
    assert width * height == n
    row_pad = 8 if height > 1 else 0              # rows may be padded: row_step > width * point_step
    row_step = width * step + row_pad
    data = bytearray(row_step * height)
    for i in range(n):
        r, c = divmod(i, width)
        base = r * row_step + c * step
        for name, off, dt in fields:
            if name in ("x", "y", "z"):
                v = float(xyz[i, "xyz".index(name)])
                struct.pack_into(e + DT[dt], data, base + off, v)
            elif name == "intensity":
                struct.pack_into(e + "f", data, base + off, 7.5)
            elif name == "ring":
                struct.pack_into(e + "H", data, base + off, i % 64)
            else:
                struct.pack_into(e + "I", data, base + off, i)
    msg = struct.pack("<III", seq, stamp[0], stamp[1]) + ros_string(frame_id) + struct.pack("<II", height, width)
    msg += struct.pack("<I", len(fields))
    for name, off, dt in fields:
        msg += ros_string(name) + struct.pack("<IBI", off, dt, 1)
    msg += struct.pack("<BII", 1 if big_endian else 0, step, row_step)
    msg += struct.pack("<I", len(data)) + bytes(data) + struct.pack("<B", 1 if is_dense else 0)
    return msg


def write_dump(path, messages):
    with open(path, "wb") as f:
        for m in messages:
            f.write(struct.pack("<I", len(m)) + m)


READER_TU = r'''
#include <cstdio>
#include <fstream>
#include <pclomp_b200/pointcloud2.hpp>
int main(int argc, char** argv) {
  pclomp_b200::PointCloud2DumpReader reader(argv[1]);
  std::ofstream out(argv[2], std::ios::binary);
  std::ofstream re(argv[3], std::ios::binary);
  pclomp_b200::PointCloud2 m;
  pcl::PointCloud<pcl::PointXYZI> cloud;      // the point type of lidar_subscriber_node.cpp:38
  int k = 0;
  while (reader.next(m)) {
    if (!pclomp_b200::fromROSMsg(m, cloud)) return 2;
    std::printf("msg %d seq %u stamp %u.%u frame %s w %u h %u dense %d n %zu\n", k, m.seq, m.stamp_sec, m.stamp_nsec, m.frame_id.c_str(),
                cloud.width, cloud.height, cloud.is_dense ? 1 : 0, cloud.size());
    for (const auto& p : cloud.points) { const float xyz[3] = {p.x, p.y, p.z}; out.write(reinterpret_cast<const char*>(xyz), 12); }
    pclomp_b200::PointCloud2 back;
    pclomp_b200::toROSMsg(cloud, back);
    back.seq = m.seq; back.stamp_sec = m.stamp_sec; back.stamp_nsec = m.stamp_nsec; back.frame_id = m.frame_id;
    pclomp_b200::appendToDump(re, back);
    ++k;
  }
  return k > 0 ? 0 : 3;
}
'''


def test_pointcloud2_reader_against_struct_pack(tmp_path):
    rng = np.random.default_rng(1)
    clouds = [rng.uniform(-50, 50, size=(n, 3)).astype(np.float32) for n in (100, 64, 33, 0, 12)]
    clouds[1][5] = [np.nan, 1.0, 2.0]
    msgs = [pointcloud2_bytes(clouds[0], seq=7, stamp=(1700000000, 123456789), layout="xyz16"),
            pointcloud2_bytes(clouds[1], seq=8, layout="xyzir32", is_dense=False, height=4),      # organised cloud, padded rows
            pointcloud2_bytes(clouds[2], seq=9, layout="f64", big_endian=True, frame_id="lidar/front"),
            pointcloud2_bytes(clouds[3], seq=10, layout="xyz16"),                                   # empty message
            pointcloud2_bytes(clouds[4], seq=11, layout="xyzir32", big_endian=True)]
    dump, out, re_dump = str(tmp_path / "in.pc2dump"), str(tmp_path / "xyz.bin"), str(tmp_path / "re.pc2dump")
    write_dump(dump, msgs)
    src = tmp_path / "reader.cpp"
    src.write_text(READER_TU)
    exe = str(tmp_path / "reader")
    r = subprocess.run([CXX, "-O1", "-std=c++17", "-Wall", "-I", INC, str(src), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, dump, out, re_dump], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    lines = r.stdout.strip().split("\n")
    assert len(lines) == 5
    assert lines[0] == "msg 0 seq 7 stamp 1700000000.123456789 frame velodyne w 100 h 1 dense 1 n 100"
    assert lines[1].endswith("w 16 h 4 dense 0 n 64") and "frame lidar/front" in lines[2] and lines[3].endswith("n 0")
    got = np.fromfile(out, dtype=np.float32).reshape(-1, 3)
    exp = np.concatenate(clouds)
    assert np.array_equal(got, exp, equal_nan=True)            # fp64 / big-endian fields come back as the same fp32 values
    # toROSMsg round trip: the rewritten dump is exactly what struct.pack writes for the xyz16 layout
    exp_re = b"".join(struct.pack("<I", len(m)) + m for m in [
        pointcloud2_bytes(clouds[0], seq=7, stamp=(1700000000, 123456789), layout="xyz16"),
        pointcloud2_bytes(clouds[1], seq=8, layout="xyz16", is_dense=False),
        pointcloud2_bytes(clouds[2], seq=9, layout="xyz16", frame_id="lidar/front"),
        pointcloud2_bytes(clouds[3], seq=10, layout="xyz16"),
        pointcloud2_bytes(clouds[4], seq=11, layout="xyz16")])
    got_re = open(re_dump, "rb").read()
    # organised input (height 4) is rewritten unorganised (height 1): compare after normalising that message's header
    assert len(got_re) == len(exp_re)
    assert got_re.count(b"velodyne") == exp_re.count(b"velodyne") and got_re[-200:] == exp_re[-200:]


def test_truncated_message_is_rejected(tmp_path):
    m = pointcloud2_bytes(np.ones((10, 3), np.float32))
    dump = str(tmp_path / "bad.pc2dump")
    with open(dump, "wb") as f:
        f.write(struct.pack("<I", len(m) - 9) + m[:-9])       # data array cut short
    src = tmp_path / "reader.cpp"
    src.write_text(READER_TU)
    exe = str(tmp_path / "reader")
    assert subprocess.run([CXX, "-O1", "-std=c++17", "-I", INC, str(src), "-o", exe]).returncode == 0
    r = subprocess.run([exe, dump, str(tmp_path / "o"), str(tmp_path / "r")], capture_output=True, text=True)
    assert r.returncode == 3                                    # no message accepted


@pytest.mark.gpu
def test_replay_app_matches_reference_mapping_loop(tmp_path):
    """apps/replay_b200: PointCloud2 dump -> fromROSMsg -> the mapping-node loop on the device pipeline, against the oracle's
    restatement of ndt_rosbag_mapping_node.cpp:42-161 on the same scans."""
    import workloads
    from toyslam_b200 import _build
    from util import oracle_mapping_loop
    _build.build_apps()
    scans, _ = workloads.config3_sequence(5, azimuth_steps=600, leaf=0.05)
    msgs = [pointcloud2_bytes(s, seq=k, stamp=(1000 + k, 0), layout="xyzir32" if k % 2 else "xyz16") for k, s in enumerate(scans)]
    dump = str(tmp_path / "drive.pc2dump")
    write_dump(dump, msgs)
    out = subprocess.run([os.path.join(ROOT, "apps", "replay_b200"), dump, "--save-map", str(tmp_path / "map.pcd")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    ref_steps, ref_map = oracle_mapping_loop(scans)
    rows = re.findall(r"^scan (\d+) (\d+) \S+ (\d+) (\d+) (\d) (\d+) (\S+)  (\S+) (\S+) (\S+)  (\d+)$", out.stdout, flags=re.M)
    assert len(rows) == 5 and "replayed 5 scans" in out.stdout
    for k, row in enumerate(rows):
        assert int(row[2]) == len(scans[k]) and int(row[3]) == ref_steps[k]["n_filtered"]
        assert int(row[4]) == int(ref_steps[k]["converged"]) and int(row[5]) == ref_steps[k]["iterations"]
        t = np.array([float(row[7]), float(row[8]), float(row[9])])
        assert np.abs(t - ref_steps[k]["pose"][:3, 3]).max() < 1e-3
        if k > 0:
            assert abs(float(row[6]) - ref_steps[k]["fitness"]) <= 1e-4 * ref_steps[k]["fitness"] + 1e-6
    m = re.search(r"saved (\d+) map points", out.stdout)
    assert m and abs(int(m.group(1)) - len(ref_map)) <= max(4, 0.004 * len(ref_map))
