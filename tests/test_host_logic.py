"""CPU tests of the host side: the C ABI library exports what include/lego_loam_b200.h declares, the
parameter struct mirrors loam_config.yaml, there is no silent CPU fallback, and the multi-GPU sharding
logic of bench.py works across 2 gloo ranks."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ll_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from lego_loam_bor_b200 import capi
    from lego_loam_bor_b200._paths import LIB_CUDA
    names = declared_functions(os.path.join(ROOT, "include", "lego_loam_b200.h"))
    assert len(names) >= 25
    lib = C.CDLL(LIB_CUDA)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS must list exactly the header's entry points"


def test_default_params_match_reference_yaml(built):
    """ll_default_params == LeGO-LOAM/config/loam_config.yaml:5-35; LegoLoamParams layout == the C struct."""
    from lego_loam_bor_b200 import capi, default_params
    from lego_loam_bor_b200.params import LegoLoamParams
    lib = capi.load_library()
    p = LegoLoamParams()
    lib.ll_default_params(C.addressof(p))
    assert p.as_dict() == default_params().as_dict()
    assert (p.num_vertical_scans, p.num_horizontal_scans, p.ground_scan_index) == (16, 1800, 7)
    assert p.segment_theta == 60.0 and p.mapping_frequency_divider == 5 and p.nearest_feature_search_distance == 5.0
    assert C.sizeof(LegoLoamParams) == 21 * 4


def test_yaml_loader_uses_reference_key_tree(tmp_path):
    from lego_loam_bor_b200.params import load_yaml
    y = tmp_path / "loam_config.yaml"
    y.write_text("lego_loam:\n  laser:\n    num_vertical_scans: 64\n    num_horizontal_scans: 2048\n    ground_scan_index: 31\n"
                 "    vertical_angle_bottom: -16.6\n    vertical_angle_top: 16.6\n  mapping:\n    mapping_frequency_divider: 3\n")
    p = load_yaml(str(y))
    assert (p.num_vertical_scans, p.num_horizontal_scans, p.ground_scan_index, p.mapping_frequency_divider) == (64, 2048, 31, 3)
    assert abs(p.vertical_angle_bottom + 16.6) < 1e-6 and p.segment_theta == 60.0
    bad = tmp_path / "bad.yaml"
    bad.write_text("lego_loam:\n  laser:\n    no_such_key: 1\n")
    with pytest.raises(KeyError):
        load_yaml(str(bad))


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path must fail loudly, never route through the oracle."""
    import torch
    from lego_loam_bor_b200 import capi, default_params
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.LegoLoamError):
        capi.LegoLoam(default_params(), batch=1)
    # and nothing under the package imports the oracle
    pkg = os.path.join(ROOT, "lego_loam_bor_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                hit = re.search(r"(^\s*(import|from)\s+oracle|#include\s+\"[^\"]*oracle|oracle_py|liblego_oracle|\blo_[a-z_]+\()", text, flags=re.M)
                assert not hit, f"{fn} uses the oracle: {hit.group(0)!r}"


def test_invalid_arguments_are_rejected(built):
    from lego_loam_bor_b200 import capi, default_params
    lib = capi.load_library()
    h = C.c_void_p()
    p = default_params()
    assert lib.ll_create(None, 1, 10, 0, None, C.byref(h)) == -1
    assert lib.ll_create(C.addressof(p), 0, 10, 0, None, C.byref(h)) == -1
    p.num_vertical_scans = 1
    assert lib.ll_create(C.addressof(p), 1, 10, 0, None, C.byref(h)) == -1
    assert lib.ll_image_projection(None) == -1 and lib.ll_destroy(None) == -1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import torch
    seqs = bench.shard_sequences(rank, 8)
    local_ms = 10.0 + 5.0 * rank  # rank 1 is slower

    def reduce_max(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    value, worst = bench.aggregate_throughput(world, 8, 20, local_ms, reduce_max)
    gathered = [None] * world
    dist.all_gather_object(gathered, seqs)
    q.put((rank, seqs, value, worst, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_strong_scaling_split_covers_every_sequence_once():
    """--total-seqs: SURVEY 8e's 64 / 32 / 16 / 8 sequences per GPU at 1 / 2 / 4 / 8 GPUs."""
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8, 3):
        shares = [bench.shard_total(r, world, 64) for r in range(world)]
        assert sorted(sum(shares, [])) == list(range(64))
        assert max(map(len, shares)) - min(map(len, shares)) <= 1
        if 64 % world == 0:
            assert all(len(x) == 64 // world for x in shares)
    v, worst = bench.aggregate_throughput(8, 64 / 8, 20, 2.0)
    assert abs(v - 64 * 20 / 2e-3) < 1e-6


def test_sharding_over_two_gloo_ranks():
    """Independent sequences are split over ranks without overlap; the job throughput is all scans over the
    slowest rank's time (the contract's max-over-ranks)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    (r0, s0, v0, w0, g0), (r1, s1, v1, w1, g1) = res
    assert set(s0).isdisjoint(s1) and len(set(s0)) == 8 and len(set(s1)) == 8
    assert g0 == g1 == [s0, s1]
    assert w0 == w1 == 15.0
    assert v0 == v1 == 2 * 8 * 20 / 15e-3


def test_shard_sequences_replication():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.shard_sequences(0, 4) == [0, 1, 2, 3]
    assert bench.shard_sequences(3, 4) == [12, 13, 14, 15]
    assert bench.shard_sequences(1, 8, unique=2) == [8, 9, 8, 9, 8, 9, 8, 9]


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/lego_loam_b200.h must compile as strict C99 (no C++ types in the
    signatures)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "lego_loam_b200.h"\n'
                   "int main(void) { LegoLoamParams p; ll_pointcloud2_view v; double o[7]; float t[6] = {0};\n"
                   "  (void)v; ll_default_params(&p); ll_transform_to_odometry(t, o); return o[6] == 1.0 ? 0 : 1; }\n")
    obj = tmp_path / "abi.o"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(obj)])
