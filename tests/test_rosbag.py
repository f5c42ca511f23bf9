"""rosbag v2.0 ingest: the bag mode of the reference's main loop (main.cpp:26-35,60-76) without
ROS.  The reference ships no bag, so the fixtures are written here by a minimal bag writer that follows the published
"Bag format 2.0" record layout (bag header padded to 4096 bytes, chunks with connection + message records, index data,
chunk info) and the sensor_msgs/PointCloud2 serialisation; the reader (ll_bag_*) must return exactly the messages of
one topic, in record-time order, with the header fields ll_set_scans_pointcloud2_host needs."""
import struct

import numpy as np
import pytest


def _field(name, value):
    b = name.encode() + b"=" + value
    return struct.pack("<I", len(b)) + b


def _record(fields, data):
    h = b"".join(_field(n, v) for n, v in fields)
    return struct.pack("<I", len(h)) + h + struct.pack("<I", len(data)) + data


def _string(s):
    b = s.encode() if isinstance(s, str) else s
    return struct.pack("<I", len(b)) + b


def lz4_block_compress(src):
    """Greedy LZ4 block encoder (published block format: last 5 bytes literal, no match starting in the last 12)."""
    n, out, anchor, i, table = len(src), bytearray(), 0, 0, {}

    def lengths(v):
        b = bytearray()
        while v >= 255:
            b.append(255)
            v -= 255
        b.append(v)
        return b

    def emit(lit, offset=None, mlen=0):
        token = (min(len(lit), 15) << 4) | (min(mlen - 4, 15) if offset else 0)
        out.append(token)
        if len(lit) >= 15:
            out.extend(lengths(len(lit) - 15))
        out.extend(lit)
        if offset:
            out.extend(struct.pack("<H", offset))
            if mlen - 4 >= 15:
                out.extend(lengths(mlen - 4 - 15))

    while n >= 13 and i <= n - 12:
        key = bytes(src[i:i + 4])
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            m = 4
            while i + m < n - 5 and src[cand + m] == src[i + m]:
                m += 1
            emit(src[anchor:i], i - cand, m)
            i += m
            anchor = i
        else:
            i += 1
    emit(src[anchor:])
    return bytes(out)


def lz4_frame_liblz4(src):
    """The same frame written by the real liblz4 (through pyarrow's codec), an encoder independent of this repo."""
    import pyarrow as pa
    return pa.compress(bytes(src), codec="lz4", asbytes=True)


def lz4_frame(src, block=3000):
    """LZ4 frame (magic 0x184D2204) as roslz4 writes bag chunks: independent blocks, content checksum present (its value is
    not checked by the reader, nor is the header checksum); every third block is stored uncompressed."""
    out = bytearray(struct.pack("<I", 0x184D2204) + bytes([0x64, 0x70, 0x00]))
    for k, a in enumerate(range(0, len(src), block)):
        part = bytes(src[a:a + block])
        if k % 3 == 2:
            out += struct.pack("<I", 0x80000000 | len(part)) + part
        else:
            c = lz4_block_compress(part)
            out += struct.pack("<I", len(c)) + c
    out += struct.pack("<I", 0) + b"\xde\xad\xbe\xef"
    return bytes(out)


def pointcloud2_message(seq, stamp, frame_id, fields, point_step, data, n_points, is_dense):
    """fields: list of (name, offset, datatype, count)"""
    out = struct.pack("<III", seq, stamp[0], stamp[1]) + _string(frame_id)
    out += struct.pack("<II", 1, n_points)  # height, width
    out += struct.pack("<I", len(fields))
    for name, off, dt, cnt in fields:
        out += _string(name) + struct.pack("<IBI", off, dt, cnt)
    out += struct.pack("<BII", 0, point_step, point_step * n_points)
    out += _string(bytes(data)) + struct.pack("<B", 1 if is_dense else 0)
    return out


def write_bag(path, connections, chunks, compression=b"none"):
    """connections: {conn: (topic, type)}; chunks: list of lists of (conn, (sec, nsec), payload)"""
    def conn_record(c):
        topic, typ = connections[c]
        data = _field("topic", topic.encode()) + _field("type", typ.encode()) + _field("md5sum", b"0" * 32) + \
            _field("message_definition", b"# not needed by the reader\n")
        return _record([("op", b"\x07"), ("conn", struct.pack("<I", c)), ("topic", topic.encode())], data)

    body = b""
    chunk_infos = []
    for msgs in chunks:
        inner, seen, index = b"", set(), {}
        for c, (sec, nsec), payload in msgs:
            if c not in seen:
                inner += conn_record(c)
                seen.add(c)
            index.setdefault(c, []).append((sec, nsec, len(inner)))
            inner += _record([("op", b"\x02"), ("conn", struct.pack("<I", c)), ("time", struct.pack("<II", sec, nsec))], payload)
        chunk_pos = 13 + 4096 + len(body)
        if compression == b"bz2":
            import bz2
            stored = bz2.compress(inner, 1 if len(body) % 2 else 9)
        else:
            stored = lz4_frame(inner) if compression == b"lz4" else (lz4_frame_liblz4(inner) if compression == b"liblz4" else inner)
        body += _record([("op", b"\x05"), ("compression", b"lz4" if compression == b"liblz4" else compression),
                         ("size", struct.pack("<I", len(inner)))], stored)
        for c, entries in index.items():
            body += _record([("op", b"\x04"), ("ver", struct.pack("<I", 1)), ("conn", struct.pack("<I", c)), ("count", struct.pack("<I", len(entries)))],
                            b"".join(struct.pack("<III", s, n, o) for s, n, o in entries))
        times = [t for _, t, _ in msgs]
        chunk_infos.append((chunk_pos, min(times), max(times), {c: len(e) for c, e in index.items()}))
    index_pos = 13 + 4096 + len(body)
    tail = b"".join(conn_record(c) for c in connections)
    for pos, t0, t1, counts in chunk_infos:
        tail += _record([("op", b"\x06"), ("ver", struct.pack("<I", 1)), ("chunk_pos", struct.pack("<Q", pos)),
                         ("start_time", struct.pack("<II", *t0)), ("end_time", struct.pack("<II", *t1)), ("count", struct.pack("<I", len(counts)))],
                        b"".join(struct.pack("<II", c, n) for c, n in counts.items()))
    hdr_fields = [("op", b"\x03"), ("index_pos", struct.pack("<Q", index_pos)), ("conn_count", struct.pack("<I", len(connections))),
                  ("chunk_count", struct.pack("<I", len(chunks)))]
    h = b"".join(_field(n, v) for n, v in hdr_fields)
    pad = 4096 - 4 - len(h) - 4
    bag_header = struct.pack("<I", len(h)) + h + struct.pack("<I", pad) + b" " * pad
    assert len(bag_header) == 4096
    with open(path, "wb") as f:
        f.write(b"#ROSBAG V2.0\n" + bag_header + body + tail)


VELO_FIELDS = [("x", 0, 7, 1), ("y", 4, 7, 1), ("z", 8, 7, 1), ("intensity", 12, 7, 1), ("ring", 16, 4, 1), ("time", 18, 7, 1)]


def velodyne_bytes(xyzi, rng):
    rec = np.zeros(len(xyzi), np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                                        "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22}))
    rec["x"], rec["y"], rec["z"], rec["intensity"] = xyzi[:, 0], xyzi[:, 1], xyzi[:, 2], xyzi[:, 3]
    rec["ring"] = rng.integers(0, 16, len(xyzi))
    return np.frombuffer(rec.tobytes(), np.uint8)


def make_bag(path, clouds, rng, **kw):
    """three topics; the lidar messages are spread over two chunks and stored out of time order"""
    conns = {0: ("/imu/data", "sensor_msgs/Imu"), 1: ("/velodyne_points", "sensor_msgs/PointCloud2"), 2: ("/other_cloud", "sensor_msgs/PointCloud2")}
    datas = [velodyne_bytes(c, rng) for c in clouds]
    lidar = [(1, (100 + i // 2, 500000000 * (i % 2)), pointcloud2_message(i, (100 + i // 2, 500000000 * (i % 2) + 7), "velodyne", VELO_FIELDS, 22, d, len(c), i % 2 == 0))
             for i, (c, d) in enumerate(zip(clouds, datas))]
    imu = [(0, (100 + i, 1), bytes(rng.integers(0, 255, 40, dtype=np.uint8))) for i in range(3)]
    other = [(2, (100, 3), pointcloud2_message(0, (100, 3), "x", VELO_FIELDS[:3], 12, np.zeros(24, np.uint8), 2, True))]
    chunk0 = [imu[0]] + other + [lidar[2], lidar[0]] + [imu[1]]
    chunk1 = [lidar[3], imu[2], lidar[1]] + lidar[4:]
    write_bag(path, conns, [chunk0, chunk1], **kw)
    return datas


def test_reader_returns_topic_in_time_order(built, tmp_path):
    from lego_loam_bor_b200.capi import RosBag
    rng = np.random.default_rng(2)
    clouds = [rng.normal(0, 10, (n, 4)).astype(np.float32) for n in (50, 1, 0, 333, 17)]
    path = str(tmp_path / "t.bag")
    datas = make_bag(path, clouds, rng)
    bag = RosBag(path, "/velodyne_points")
    assert len(bag) == len(clouds) and bag.topic == "/velodyne_points"
    for i, (c, d) in enumerate(zip(clouds, datas)):
        v, data = bag.message(i)
        assert (v.width, v.height, v.point_step, v.row_step) == (len(c), 1, 22, 22 * len(c))
        assert (v.off_x, v.off_y, v.off_z, v.off_intensity) == (0, 4, 8, 12)
        assert (v.stamp_sec, v.stamp_nsec) == (100 + i // 2, 500000000 * (i % 2) + 7)
        assert v.bag_time_ns == (100 + i // 2) * 10**9 + 500000000 * (i % 2)
        assert v.is_dense == (1 if i % 2 == 0 else 0) and v.is_bigendian == 0
        assert np.array_equal(data, d)
    # no topic given: the first PointCloud2 topic of the file
    first = RosBag(path)
    assert first.topic == "/other_cloud" and len(first) == 1
    v, data = first.message(0)
    assert (v.off_x, v.off_y, v.off_z, v.off_intensity, v.point_step, v.width) == (0, 4, 8, -1, 12, 2)
    assert len(RosBag(path, "/no_such_topic")) == 0


def test_compressed_chunks(built, tmp_path):
    """`rosbag record --lz4` / `-j`: the same messages from LZ4-framed chunks (compressed, stored and self-overlapping
    matches; written by the test's encoder and by liblz4) and from bzip2 chunks (written by libbz2)."""
    from lego_loam_bor_b200.capi import LegoLoamError, RosBag
    rng = np.random.default_rng(6)
    clouds = [rng.normal(0, 10, (n, 4)).astype(np.float32) for n in (400, 1, 0, 900, 17)]
    clouds[3][100:600] = 0  # long runs of equal bytes: matches that overlap their own output
    sizes = {}
    codecs = [b"none", b"lz4", b"bz2"]  # bz2 chunks are written by libbz2 (Python's bz2 module)
    try:
        import pyarrow as pa
        if pa.Codec.is_available("lz4"):
            codecs.append(b"liblz4")  # chunks compressed by the real library: pins the decoder, not just its round trip
    except ImportError:
        pass
    for comp in codecs:
        path = tmp_path / (comp.decode() + ".bag")
        datas = make_bag(str(path), clouds, np.random.default_rng(7), compression=comp)
        sizes[comp] = path.stat().st_size
        bag = RosBag(str(path), "/velodyne_points")
        assert len(bag) == len(clouds)
        for i, d in enumerate(datas):
            v, data = bag.message(i)
            assert v.width == len(clouds[i]) and np.array_equal(data, d), f"{comp} message {i}"
    assert sizes[b"lz4"] < sizes[b"none"]
    # a damaged frame is reported, not read past
    whole = bytearray((tmp_path / "lz4.bag").read_bytes())
    at = whole.index(struct.pack("<I", 0x184D2204))
    whole[at + 4] = 0x24  # frame version 00
    (tmp_path / "bad.bag").write_bytes(bytes(whole))
    with pytest.raises(LegoLoamError, match="corrupt lz4 chunk"):
        RosBag(str(tmp_path / "bad.bag"))


def test_compressed_chunks_large(built, tmp_path):
    """Chunks larger than a bzip2 block (100 kB at level 1) and than an LZ4 frame block of the test encoder; long runs of
    equal bytes (bzip2's initial run-length stage: four equal bytes + a count) and incompressible noise."""
    from lego_loam_bor_b200.capi import RosBag
    rng = np.random.default_rng(8)
    big = rng.normal(0, 30, (14000, 4)).astype(np.float32)
    big[2000:9000] = 0
    big[9000:9300, 0] = 1.5
    clouds = [big, rng.normal(0, 1, (3000, 4)).astype(np.float32), np.zeros((5000, 4), np.float32), big[::-1].copy()]
    codecs = [b"bz2", b"lz4"]
    try:
        import pyarrow as pa
        if pa.Codec.is_available("lz4"):
            codecs.append(b"liblz4")
    except ImportError:
        pass
    for comp in codecs:
        path = tmp_path / (comp.decode() + "_big.bag")
        datas = make_bag(str(path), clouds, np.random.default_rng(9), compression=comp)
        bag = RosBag(str(path), "/velodyne_points")
        assert len(bag) == len(clouds)
        for i, d in enumerate(datas):
            v, data = bag.message(i)
            assert np.array_equal(data, d), f"{comp} message {i}"


def test_reader_errors(built, tmp_path):
    from lego_loam_bor_b200.capi import LegoLoamError, RosBag
    rng = np.random.default_rng(3)
    clouds = [rng.normal(0, 10, (n, 4)).astype(np.float32) for n in (5, 6, 7, 8, 9)]
    with pytest.raises(LegoLoamError, match="cannot open"):
        RosBag(str(tmp_path / "missing.bag"))
    p = tmp_path / "junk.bag"
    p.write_bytes(b"#ROSBAG V1.2\n" + b"\0" * 100)
    with pytest.raises(LegoLoamError, match="not a rosbag v2.0"):
        RosBag(str(p))
    make_bag(str(tmp_path / "xz.bag"), clouds, rng, compression=b"xz")
    with pytest.raises(LegoLoamError, match="compression 'xz' is not supported"):
        RosBag(str(tmp_path / "xz.bag"))
    make_bag(str(tmp_path / "bz2.bag"), clouds, rng, compression=b"bz2")
    whole = bytearray((tmp_path / "bz2.bag").read_bytes())
    at = whole.index(b"BZh")
    whole[at + 4] ^= 0xff  # block magic
    (tmp_path / "bad_bz2.bag").write_bytes(bytes(whole))
    with pytest.raises(LegoLoamError, match="corrupt bz2 chunk"):
        RosBag(str(tmp_path / "bad_bz2.bag"))
    make_bag(str(tmp_path / "ok.bag"), clouds, rng)
    whole = (tmp_path / "ok.bag").read_bytes()
    (tmp_path / "cut.bag").write_bytes(whole[:13 + 4096 + 300])
    with pytest.raises(LegoLoamError, match="truncated"):
        RosBag(str(tmp_path / "cut.bag"))


@pytest.mark.gpu
def test_bag_to_device(built, tmp_path):
    """bag -> ll_set_scans_pointcloud2_host -> projection, against the oracle fed with the same clouds"""
    from lego_loam_bor_b200.capi import LegoLoam, RosBag
    from oracle.oracle_py import Oracle
    from parity_utils import make_scans, same_bits
    rng = np.random.default_rng(4)
    p, cfg, scans = make_scans("A", [0], range(5))
    clouds = [scans[(0, f)] for f in range(5)]
    path = str(tmp_path / "seq.bag")
    make_bag(path, clouds, rng)
    bag = RosBag(path, "/velodyne_points")
    gpu, o = LegoLoam(p, batch=1), Oracle(p)
    for i in range(len(bag)):
        v, data = bag.message(i)
        gpu.set_scans_pointcloud2([data], v.point_step, v.off_x, v.off_y, v.off_z, v.off_intensity, is_dense=bool(v.is_dense))
        gpu.process_scans()
        o.image_projection(clouds[i])
        o.feature_association()
        for name in ("RANGE_MAT", "LABEL_MAT", "SEG_COL_IND", "SURF_LAST", "TRANSFORM_SUM"):
            assert same_bits(gpu.download(name, 0), o.download(name)), f"message {i}: {name}"


def test_reader_survives_corruption(built, tmp_path):
    """A bag is an untrusted file: truncated or bit-flipped bags (all three chunk codecs) must end in LegoLoamError or in
    a readable bag, never in a crash.  The loop runs in a child process so that a crash would show as its exit code."""
    import subprocess
    import sys
    rng = np.random.default_rng(12)
    clouds = [rng.normal(0, 10, (n, 4)).astype(np.float32) for n in (300, 40, 0, 700, 9)]
    clouds[3][100:500] = 0
    for comp in (b"none", b"lz4", b"bz2"):
        make_bag(str(tmp_path / (comp.decode() + ".bag")), clouds, np.random.default_rng(13), compression=comp)
    child = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from lego_loam_bor_b200.capi import LegoLoamError, RosBag
rng = np.random.default_rng(14)
ok = bad = 0
for comp in ("none", "lz4", "bz2"):
    whole = open(sys.argv[2] + "/" + comp + ".bag", "rb").read()
    for trial in range(150):
        b = bytearray(whole)
        if trial % 3 == 0:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(13 + 4096, len(b)))] ^= 1 << int(rng.integers(0, 8))
        path = sys.argv[2] + "/fuzz.bag"
        open(path, "wb").write(bytes(b))
        try:
            bag = RosBag(path, "/velodyne_points")
            for i in range(len(bag)):
                try:
                    bag.message(i)
                except LegoLoamError:
                    bad += 1
            bag.close()
            ok += 1
        except LegoLoamError:
            bad += 1
print(ok, bad)
"""
    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", child, root, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    ok, bad = map(int, r.stdout.split())
    assert ok + bad >= 450 and bad > 0
